// Shared dispatch of set_union / set_intersection / set_difference / set_symmetric_difference -> bcb_set_operation.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_DETAIL_SET_OPERATION_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_DETAIL_SET_OPERATION_HPP

#include <iterator>
#include <type_traits>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class InputIterator1, class InputIterator2, class OutputIterator>
inline OutputIterator set_operation(int which, InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, InputIterator2 last2,
                                    OutputIterator result, command_queue &queue)
{
    static_assert(is_device_iterator<InputIterator1>::value && is_device_iterator<InputIterator2>::value &&
                  is_device_iterator<OutputIterator>::value, "set operations: device ranges required");
    typedef typename std::iterator_traits<InputIterator1>::value_type T;
    typedef typename std::iterator_traits<InputIterator2>::value_type T2;
    typedef typename std::iterator_traits<OutputIterator>::value_type R;
    static_assert(dtype_of<T>::supported, "set operations: scalar value types only");
    static_assert(std::is_same<T, T2>::value, "set operations: both ranges must have the same value type");
    // the result may be another integer type of the same width (test_set_union.cpp:24-42 writes int_ into a uint_
    // vector): a same-width integer conversion keeps the bits
    static_assert(std::is_same<T, R>::value || (std::is_integral<T>::value && std::is_integral<R>::value && sizeof(T) == sizeof(R)),
                  "set operations: the result range must have the inputs' value type");
    size_t count = 0;
    queue.make_current();
    check(bcb_set_operation(queue.get(), dtype_of<T>::value, which, first1.device_ptr(), iterator_range_size(first1, last1), first2.device_ptr(),
                            iterator_range_size(first2, last2), result.device_ptr(), &count));
    return result + static_cast<typename std::iterator_traits<OutputIterator>::difference_type>(count);
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
