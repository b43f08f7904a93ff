// is_sorted (algorithm/is_sorted.hpp:39-68 of the reference): true when no adjacent pair is out of order
// under less<T> (default) or greater<T>.  Blocks (returns a host bool).
#ifndef B200_BOOST_COMPUTE_ALGORITHM_IS_SORTED_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_IS_SORTED_HPP

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/functional/field.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T>
inline bool is_sorted_impl(buffer_iterator<T> first, buffer_iterator<T> last, bool descending, command_queue &queue)
{
    static_assert(dtype_of<T>::supported, "is_sorted(): scalar key types only");
    int result = 1;
    queue.make_current();
    check(bcb_is_sorted(queue.get(), dtype_of<T>::value, descending ? 1 : 0, first.device_ptr(),
                        iterator_range_size(first, last), &result));
    return result != 0;
}

} // namespace detail

template<class T>
inline bool is_sorted(buffer_iterator<T> first, buffer_iterator<T> last, command_queue &queue = system::default_queue())
{
    return detail::is_sorted_impl(first, last, false, queue);
}

template<class T>
inline bool is_sorted(buffer_iterator<T> first, buffer_iterator<T> last, less<T>, command_queue &queue = system::default_queue())
{
    return detail::is_sorted_impl(first, last, false, queue);
}

template<class T>
inline bool is_sorted(buffer_iterator<T> first, buffer_iterator<T> last, greater<T>, command_queue &queue = system::default_queue())
{
    return detail::is_sorted_impl(first, last, true, queue);
}

// with a field comparator (functional/field.hpp): no adjacent pair with compare(x[i+1], x[i])
template<class T, class Compare>
inline typename std::enable_if<is_field_compare<Compare>::value, bool>::type
is_sorted(buffer_iterator<T> first, buffer_iterator<T> last, Compare compare, command_queue &queue = system::default_queue())
{
    const field_spec f = compare.template resolve<T>();
    int result = 1;
    queue.make_current();
    detail::check(bcb_is_sorted_by_field(queue.get(), first.device_ptr(), detail::iterator_range_size(first, last), sizeof(T), f.offset, f.dtype,
                                 f.unary, f.descending ? 1 : 0, &result));
    return result != 0;
}

} // namespace compute
} // namespace boost

#endif
