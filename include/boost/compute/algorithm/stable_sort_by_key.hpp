// stable_sort_by_key() (algorithm/stable_sort_by_key.hpp:29-163 of the reference): radix_sort_by_key for
// less<T> / greater<T> on radix-sortable keys.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_STABLE_SORT_BY_KEY_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_STABLE_SORT_BY_KEY_HPP

#include <iterator>

#include <boost/compute/algorithm/detail/radix_sort.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/functional/operator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T, class T2>
inline void dispatch_gpu_ssort_by_key(buffer_iterator<T> kf, buffer_iterator<T> kl, buffer_iterator<T2> vf, less<T>,
                                      command_queue &queue)
{
    radix_sort_by_key(kf, kl, vf, true, queue);
}
template<class T, class T2>
inline void dispatch_gpu_ssort_by_key(buffer_iterator<T> kf, buffer_iterator<T> kl, buffer_iterator<T2> vf, greater<T>,
                                      command_queue &queue)
{
    radix_sort_by_key(kf, kl, vf, false, queue);
}
template<class T, class T2, class Compare>
inline void dispatch_gpu_ssort_by_key(buffer_iterator<T>, buffer_iterator<T>, buffer_iterator<T2>, Compare, command_queue &)
{
    static_assert(sizeof(T) == 0, "stable_sort_by_key(): only less<T> and greater<T> are supported on this path");
}

} // namespace detail

template<class KeyIterator, class ValueIterator, class Compare>
inline void stable_sort_by_key(KeyIterator keys_first, KeyIterator keys_last, ValueIterator values_first,
                               Compare compare, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<KeyIterator>::value, "stable_sort_by_key(): keys must be a device range");
    static_assert(is_device_iterator<ValueIterator>::value, "stable_sort_by_key(): values must be a device range");
    detail::dispatch_gpu_ssort_by_key(keys_first, keys_last, values_first, compare, queue);
}

template<class KeyIterator, class ValueIterator>
inline void stable_sort_by_key(KeyIterator keys_first, KeyIterator keys_last, ValueIterator values_first,
                               command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<KeyIterator>::value_type key_type;
    ::boost::compute::stable_sort_by_key(keys_first, keys_last, values_first, less<key_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
