// fill / fill_n (algorithm/fill.hpp, fill_n.hpp of the reference): one trivial kernel behind bcb_fill.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_FILL_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_FILL_HPP

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {

template<class T, class V>
inline void fill(buffer_iterator<T> first, buffer_iterator<T> last, const V &value,
                 command_queue &queue = system::default_queue())
{
    const std::size_t n = detail::iterator_range_size(first, last);
    if(n == 0){
        return;
    }
    const T v = static_cast<T>(value);
    queue.make_current();
    detail::check(bcb_fill(queue.get(), first.device_ptr(), n, &v, sizeof(T)));
}

template<class T, class Size, class V>
inline void fill_n(buffer_iterator<T> first, Size count, const V &value,
                   command_queue &queue = system::default_queue())
{
    ::boost::compute::fill(first, first + count, value, queue);
}

} // namespace compute
} // namespace boost

#endif
