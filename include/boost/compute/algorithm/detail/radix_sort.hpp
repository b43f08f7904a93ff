// detail::radix_sort / radix_sort_by_key (algorithm/detail/radix_sort.hpp:428-462 of the reference).
// The reference builds and runs 3 OpenCL kernels per 4-bit pass here; this header only turns the iterator
// arguments into (device pointer, count, dtype code) and calls the ahead-of-time compiled onesweep sort.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_DETAIL_RADIX_SORT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_DETAIL_RADIX_SORT_HPP

#include <iterator>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T>
inline void radix_sort(buffer_iterator<T> first, buffer_iterator<T> last, const bool ascending, command_queue &queue)
{
    static_assert(is_radix_sortable<T>::value, "radix_sort(): key type is not radix sortable");
    queue.make_current();
    check(bcb_radix_sort(queue.get(), dtype_of<T>::value, ascending ? 1 : 0, first.device_ptr(),
                         iterator_range_size(first, last), 0, 0));
}

template<class T>
inline void radix_sort(buffer_iterator<T> first, buffer_iterator<T> last, command_queue &queue)
{
    radix_sort(first, last, true, queue);
}

template<class T, class T2>
inline void radix_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                              buffer_iterator<T2> values_first, const bool ascending, command_queue &queue)
{
    static_assert(is_radix_sortable<T>::value, "radix_sort_by_key(): key type is not radix sortable");
    queue.make_current();
    check(bcb_radix_sort(queue.get(), dtype_of<T>::value, ascending ? 1 : 0, keys_first.device_ptr(),
                         iterator_range_size(keys_first, keys_last), values_first.device_ptr(), sizeof(T2)));
}

template<class T, class T2>
inline void radix_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                              buffer_iterator<T2> values_first, command_queue &queue)
{
    radix_sort_by_key(keys_first, keys_last, values_first, true, queue);
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
