// detail::serial_insertion_sort(_by_key) (algorithm/detail/insertion_sort.hpp:25-159 of the reference):
// single-thread stable insertion sort with the native compare, used below the radix-sort size thresholds.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_DETAIL_INSERTION_SORT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_DETAIL_INSERTION_SORT_HPP

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T>
inline void serial_insertion_sort_impl(buffer_iterator<T> first, buffer_iterator<T> last, bool use_greater,
                                       void *values, std::size_t value_bytes, command_queue &queue)
{
    static_assert(dtype_of<T>::supported, "serial_insertion_sort(): scalar key types only");
    queue.make_current();
    check(bcb_insertion_sort(queue.get(), dtype_of<T>::value, use_greater ? 1 : 0, first.device_ptr(),
                             iterator_range_size(first, last), values, value_bytes));
}

template<class T>
inline void serial_insertion_sort(buffer_iterator<T> first, buffer_iterator<T> last, less<T>, command_queue &queue)
{
    serial_insertion_sort_impl(first, last, false, 0, 0, queue);
}
template<class T>
inline void serial_insertion_sort(buffer_iterator<T> first, buffer_iterator<T> last, greater<T>, command_queue &queue)
{
    serial_insertion_sort_impl(first, last, true, 0, 0, queue);
}
template<class T>
inline void serial_insertion_sort(buffer_iterator<T> first, buffer_iterator<T> last, command_queue &queue)
{
    serial_insertion_sort_impl(first, last, false, 0, 0, queue);
}

template<class T, class T2>
inline void serial_insertion_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                                         buffer_iterator<T2> values_first, less<T>, command_queue &queue)
{
    serial_insertion_sort_impl(keys_first, keys_last, false, values_first.device_ptr(), sizeof(T2), queue);
}
template<class T, class T2>
inline void serial_insertion_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                                         buffer_iterator<T2> values_first, greater<T>, command_queue &queue)
{
    serial_insertion_sort_impl(keys_first, keys_last, true, values_first.device_ptr(), sizeof(T2), queue);
}
template<class T, class T2>
inline void serial_insertion_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                                         buffer_iterator<T2> values_first, command_queue &queue)
{
    serial_insertion_sort_impl(keys_first, keys_last, false, values_first.device_ptr(), sizeof(T2), queue);
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
