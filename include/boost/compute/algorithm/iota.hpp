// iota (algorithm/iota.hpp of the reference): first[i] = value + i.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_IOTA_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_IOTA_HPP

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {

template<class T, class V>
inline void iota(buffer_iterator<T> first, buffer_iterator<T> last, const V &value,
                 command_queue &queue = system::default_queue())
{
    static_assert(detail::dtype_of<T>::supported, "iota(): scalar value types only");
    const std::size_t n = detail::iterator_range_size(first, last);
    if(n == 0){
        return;
    }
    const T start = static_cast<T>(value);
    queue.make_current();
    detail::check(bcb_iota(queue.get(), detail::dtype_of<T>::value, first.device_ptr(), n, &start));
}

} // namespace compute
} // namespace boost

#endif
