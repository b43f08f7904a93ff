// sort() (algorithm/sort.hpp:182-202 of the reference) and its dispatch (sort.hpp:34-148).
//   device range, less<T> / greater<T>, radix-sortable T:  n < 2 nothing; n <= 32 serial insertion sort;
//                                                          otherwise radix sort (ascending / descending)
//   host range (any random-access contiguous range of T):  copied to the device, sorted, copied back -- the
//                                                          reference maps the range with a mapped_view (:125-148)
//   device range, field comparator (functional/field.hpp):  detail::merge_sort_on_gpu (stable)
// Arbitrary comparison functions need run-time OpenCL code generation in the reference: they fail to compile with a
// clear message.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_SORT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_SORT_HPP

#include <iterator>
#include <type_traits>

#include <boost/compute/algorithm/detail/insertion_sort.hpp>
#include <boost/compute/algorithm/detail/merge_sort_on_gpu.hpp>
#include <boost/compute/algorithm/detail/radix_sort.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/functional/operator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T>
inline void dispatch_gpu_sort(buffer_iterator<T> first, buffer_iterator<T> last, less<T> compare, command_queue &queue)
{
    const std::size_t count = iterator_range_size(first, last);
    if(count < 2){
        return;
    }
    if(count <= 32){
        serial_insertion_sort(first, last, compare, queue);
    } else {
        radix_sort(first, last, true, queue);
    }
}

template<class T>
inline void dispatch_gpu_sort(buffer_iterator<T> first, buffer_iterator<T> last, greater<T> compare, command_queue &queue)
{
    const std::size_t count = iterator_range_size(first, last);
    if(count < 2){
        return;
    }
    if(count <= 32){
        serial_insertion_sort(first, last, compare, queue);
    } else {
        radix_sort(first, last, false, queue);
    }
}

// custom comparators (sort.hpp:83-106 of the reference: merge_sort_on_gpu): the field-comparator family of
// functional/field.hpp, as a stable key-value radix sort with the records as payload
template<class T, class Compare>
inline typename std::enable_if<is_field_compare<Compare>::value>::type
dispatch_gpu_sort(buffer_iterator<T> first, buffer_iterator<T> last, Compare compare, command_queue &queue)
{
    merge_sort_on_gpu(first, last, compare, queue);
}

template<class T, class Compare>
inline typename std::enable_if<!is_field_compare<Compare>::value>::type
dispatch_gpu_sort(buffer_iterator<T>, buffer_iterator<T>, Compare, command_queue &)
{
    static_assert(sizeof(T) == 0, "sort(): less<T>, greater<T> and the field comparators of functional/field.hpp (less_by, "
                                  "less_by_component, get<N>(_1) < get<N>(_2), abs(_1) < abs(_2)) are supported; an arbitrary "
                                  "comparison function needs the reference's run-time OpenCL code generation");
}

// device iterators
template<class Iterator, class Compare>
inline void dispatch_sort(Iterator first, Iterator last, Compare compare, command_queue &queue, std::true_type)
{
    dispatch_gpu_sort(first, last, compare, queue);
}

template<class T> inline int host_sort_order(less<T>) { return 0; }
template<class T> inline int host_sort_order(greater<T>) { return 1; }

// host iterators
template<class Iterator, class Compare>
inline void dispatch_sort(Iterator first, Iterator last, Compare compare, command_queue &queue, std::false_type)
{
    typedef typename std::iterator_traits<Iterator>::value_type T;
    static_assert(dtype_of<T>::supported, "sort(): scalar key types only");
    const std::size_t count = iterator_range_size(first, last);
    if(count < 2){
        return;
    }
    queue.make_current();
    check(bcb_sort_host(queue.get(), dtype_of<T>::value, host_sort_order<T>(compare), &*first, count));
}

} // namespace detail

template<class Iterator, class Compare>
inline void sort(Iterator first, Iterator last, Compare compare, command_queue &queue = system::default_queue())
{
    detail::dispatch_sort(first, last, compare, queue, typename is_device_iterator<Iterator>::type());
}

template<class Iterator>
inline void sort(Iterator first, Iterator last, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<Iterator>::value_type value_type;
    ::boost::compute::sort(first, last, ::boost::compute::less<value_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
