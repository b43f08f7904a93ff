#ifndef B200_BOOST_COMPUTE_DETAIL_DEFAULT_QUEUE_HPP
#define B200_BOOST_COMPUTE_DETAIL_DEFAULT_QUEUE_HPP

#include <boost/compute/system.hpp>

namespace boost {
namespace compute {
namespace detail {

inline command_queue& default_queue_ref()
{
    return system::default_queue();
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
