// device (device.hpp:49-157 of the reference): identifies one CUDA device; only the queries the
// sort / scan / reduce path and its tests use.
#ifndef B200_BOOST_COMPUTE_DEVICE_HPP
#define B200_BOOST_COMPUTE_DEVICE_HPP

#include <cstddef>
#include <string>

#include <boost/compute/exception/opencl_error.hpp>

namespace boost {
namespace compute {

class device
{
public:
    enum type {
        cpu = (1 << 1),
        gpu = (1 << 2),
        accelerator = (1 << 3)
    };

    device() : m_id(-1) {}
    explicit device(int ordinal) : m_id(ordinal) {}

    int id() const { return m_id; }
    int get() const { return m_id; }

    // every device behind this library is a GPU, so dispatch_sort & co. always take the GPU branch
    // (algorithm/sort.hpp:117-121)
    unsigned long long type() const { return gpu; }

    std::string name() const
    {
        char buf[256] = {0};
        detail::check(bcb_device_info(m_id, buf, sizeof(buf), 0, 0, 0, 0));
        return std::string(buf);
    }
    std::string vendor() const { return "NVIDIA Corporation"; }

    unsigned int compute_units() const
    {
        int cu = 0;
        detail::check(bcb_device_info(m_id, 0, 0, &cu, 0, 0, 0));
        return static_cast<unsigned int>(cu);
    }

    std::size_t global_memory_size() const
    {
        std::size_t bytes = 0;
        detail::check(bcb_device_info(m_id, 0, 0, 0, &bytes, 0, 0));
        return bytes;
    }

    bool supports_extension(const std::string &name) const
    {
        return name == "cl_khr_fp64"; // doubles are native
    }

    bool operator==(const device &other) const { return m_id == other.m_id; }
    bool operator!=(const device &other) const { return m_id != other.m_id; }

private:
    int m_id;
};

} // namespace compute
} // namespace boost

#endif
