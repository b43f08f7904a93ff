// context (context.hpp:79-98 of the reference): here simply "the primary CUDA context of one device".
#ifndef B200_BOOST_COMPUTE_CONTEXT_HPP
#define B200_BOOST_COMPUTE_CONTEXT_HPP

#include <boost/compute/device.hpp>

namespace boost {
namespace compute {

class context
{
public:
    context() : m_device() {}
    explicit context(const device &d) : m_device(d) {}

    device get_device() const { return m_device; }
    bool operator==(const context &other) const { return m_device == other.m_device; }
    bool operator!=(const context &other) const { return !(*this == other); }

private:
    device m_device;
};

} // namespace compute
} // namespace boost

#endif
