// sort_by_key() (algorithm/sort_by_key.hpp:135-163 of the reference) and dispatch_gpu_sort_by_key (:33-86):
// fewer than 32 keys -> serial insertion sort by key, otherwise stable radix sort carrying the values.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_SORT_BY_KEY_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_SORT_BY_KEY_HPP

#include <iterator>

#include <boost/compute/algorithm/detail/insertion_sort.hpp>
#include <boost/compute/algorithm/detail/radix_sort.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/functional/operator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T, class T2>
inline void dispatch_gpu_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                                     buffer_iterator<T2> values_first, less<T> compare, command_queue &queue)
{
    if(iterator_range_size(keys_first, keys_last) < 32){
        serial_insertion_sort_by_key(keys_first, keys_last, values_first, compare, queue);
    } else {
        radix_sort_by_key(keys_first, keys_last, values_first, true, queue);
    }
}

template<class T, class T2>
inline void dispatch_gpu_sort_by_key(buffer_iterator<T> keys_first, buffer_iterator<T> keys_last,
                                     buffer_iterator<T2> values_first, greater<T> compare, command_queue &queue)
{
    if(iterator_range_size(keys_first, keys_last) < 32){
        serial_insertion_sort_by_key(keys_first, keys_last, values_first, compare, queue);
    } else {
        radix_sort_by_key(keys_first, keys_last, values_first, false, queue);
    }
}

template<class T, class T2, class Compare>
inline void dispatch_gpu_sort_by_key(buffer_iterator<T>, buffer_iterator<T>, buffer_iterator<T2>, Compare, command_queue &)
{
    static_assert(sizeof(T) == 0, "sort_by_key(): only less<T> and greater<T> are supported on this path");
}

} // namespace detail

template<class KeyIterator, class ValueIterator, class Compare>
inline void sort_by_key(KeyIterator keys_first, KeyIterator keys_last, ValueIterator values_first, Compare compare,
                        command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<KeyIterator>::value, "sort_by_key(): keys must be a device range");
    static_assert(is_device_iterator<ValueIterator>::value, "sort_by_key(): values must be a device range");
    detail::dispatch_gpu_sort_by_key(keys_first, keys_last, values_first, compare, queue);
}

template<class KeyIterator, class ValueIterator>
inline void sort_by_key(KeyIterator keys_first, KeyIterator keys_last, ValueIterator values_first,
                        command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<KeyIterator>::value_type key_type;
    ::boost::compute::sort_by_key(keys_first, keys_last, values_first, less<key_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
