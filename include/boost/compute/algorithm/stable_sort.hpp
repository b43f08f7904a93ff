// stable_sort() (algorithm/stable_sort.hpp:52-110 of the reference): with less<T> / greater<T> on a
// radix-sortable T it goes straight to the (stable) radix sort, with no small-n branch.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_STABLE_SORT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_STABLE_SORT_HPP

#include <iterator>

#include <boost/compute/algorithm/detail/merge_sort_on_gpu.hpp>
#include <boost/compute/algorithm/detail/radix_sort.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/functional/operator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T>
inline void dispatch_gpu_stable_sort(buffer_iterator<T> first, buffer_iterator<T> last, less<T>, command_queue &queue)
{
    radix_sort(first, last, true, queue);
}
template<class T>
inline void dispatch_gpu_stable_sort(buffer_iterator<T> first, buffer_iterator<T> last, greater<T>, command_queue &queue)
{
    radix_sort(first, last, false, queue);
}
// custom comparators (stable_sort.hpp:34-50 of the reference: merge_sort_on_gpu with stable = true)
template<class T, class Compare>
inline typename std::enable_if<is_field_compare<Compare>::value>::type
dispatch_gpu_stable_sort(buffer_iterator<T> first, buffer_iterator<T> last, Compare compare, command_queue &queue)
{
    merge_sort_on_gpu(first, last, compare, true, queue);
}
template<class T, class Compare>
inline typename std::enable_if<!is_field_compare<Compare>::value>::type
dispatch_gpu_stable_sort(buffer_iterator<T>, buffer_iterator<T>, Compare, command_queue &)
{
    static_assert(sizeof(T) == 0, "stable_sort(): less<T>, greater<T> and the field comparators of functional/field.hpp are supported");
}

} // namespace detail

template<class Iterator, class Compare>
inline void stable_sort(Iterator first, Iterator last, Compare compare, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<Iterator>::value, "stable_sort(): device range required");
    detail::dispatch_gpu_stable_sort(first, last, compare, queue);
}

template<class Iterator>
inline void stable_sort(Iterator first, Iterator last, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<Iterator>::value_type value_type;
    ::boost::compute::stable_sort(first, last, less<value_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
