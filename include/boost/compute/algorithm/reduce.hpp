// reduce() (algorithm/reduce.hpp:275-305 of the reference): result = first[0] op ... op first[n-1].
//   * an empty range leaves *result untouched (:283-285);
//   * result may be a host iterator/pointer (blocks until the value is there, like the copy_n at :225) or a
//     device iterator (enqueue-and-return);
//   * the arithmetic type is the functor's type U for plus<U> etc. (result_of<F(T,T)>), so plus<float> over a
//     uchar range accumulates in float (test_reduce.cpp:269-277).
// One kernel launch (vectorised loads + warp shuffles + last-block fold) replaces reduce_on_gpu's three.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_REDUCE_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_REDUCE_HPP

#include <iterator>
#include <type_traits>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

// result is a device iterator
template<class T, class U, class BinaryFunction>
inline void dispatch_reduce(buffer_iterator<T> first, buffer_iterator<T> last, buffer_iterator<U> result,
                            BinaryFunction, command_queue &queue, std::true_type)
{
    typedef typename BinaryFunction::result_type R;
    static_assert(std::is_same<R, U>::value, "reduce(): the device result range must have the functor's value type");
    queue.make_current();
    check(bcb_reduce(queue.get(), dtype_of<T>::value, dtype_of<R>::value, BinaryFunction::op_code, first.device_ptr(),
                     iterator_range_size(first, last), result.device_ptr(), 1));
}

// result is a host iterator / pointer
template<class T, class OutputIterator, class BinaryFunction>
inline void dispatch_reduce(buffer_iterator<T> first, buffer_iterator<T> last, OutputIterator result,
                            BinaryFunction, command_queue &queue, std::false_type)
{
    typedef typename BinaryFunction::result_type R;
    R value;
    queue.make_current();
    check(bcb_reduce(queue.get(), dtype_of<T>::value, dtype_of<R>::value, BinaryFunction::op_code, first.device_ptr(),
                     iterator_range_size(first, last), &value, 0));
    *result = value;
}

} // namespace detail

template<class InputIterator, class OutputIterator, class BinaryFunction>
inline typename std::enable_if<!std::is_same<BinaryFunction, command_queue>::value>::type
reduce(InputIterator first, InputIterator last, OutputIterator result, BinaryFunction function,
       command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "reduce(): device input range required");
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "reduce(): scalar value types only");
    if(first == last){
        return;
    }
    detail::dispatch_reduce(first, last, result, function, queue, typename is_device_iterator<OutputIterator>::type());
}

template<class InputIterator, class OutputIterator>
inline void reduce(InputIterator first, InputIterator last, OutputIterator result,
                   command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    ::boost::compute::reduce(first, last, result, plus<T>(), queue);
}

} // namespace compute
} // namespace boost

#endif
