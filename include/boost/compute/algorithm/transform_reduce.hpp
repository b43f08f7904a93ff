// transform_reduce() (algorithm/transform_reduce.hpp:40-90 of the reference) and inner_product()
// (inner_product.hpp:40-97).  The reference feeds transform iterators into reduce / accumulate; here one fused
// load-transform-fold kernel runs (same one-launch structure as reduce()).  Unary form: transform is a unary tag
// (identity / negate / abs / square); binary form: transform is an operator tag (multiplies, plus, minus, min, max).
// `result` is a host pointer (blocks) or a device iterator (enqueue-and-return); an empty range leaves it untouched.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_REDUCE_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_REDUCE_HPP

#include <iterator>
#include <type_traits>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/lambda/placeholders.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T, class OutputIterator>
inline void run_transform_reduce(const void *in1, const void *in2, size_t n, int transform, int reduce_op, OutputIterator result,
                                 command_queue &queue, std::true_type /* device result */)
{
    queue.make_current();
    check(bcb_transform_reduce(queue.get(), dtype_of<T>::value, in1, in2, n, transform, reduce_op, result.device_ptr(), 1));
}

template<class T, class OutputIterator>
inline void run_transform_reduce(const void *in1, const void *in2, size_t n, int transform, int reduce_op, OutputIterator result,
                                 command_queue &queue, std::false_type /* host result */)
{
    if(n == 0){
        return;
    }
    T value;
    queue.make_current();
    check(bcb_transform_reduce(queue.get(), dtype_of<T>::value, in1, in2, n, transform, reduce_op, &value, 0));
    *result = value;
}

} // namespace detail

// unary
template<class InputIterator, class OutputIterator, class UnaryTransformFunction, class BinaryReduceFunction>
inline typename std::enable_if<is_device_iterator<InputIterator>::value && !is_device_iterator<UnaryTransformFunction>::value>::type
transform_reduce(InputIterator first, InputIterator last, OutputIterator result, UnaryTransformFunction, BinaryReduceFunction,
                 command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "transform_reduce(): scalar value types only");
    detail::run_transform_reduce<T>(first.device_ptr(), nullptr, detail::iterator_range_size(first, last), UnaryTransformFunction::unary_code,
                                    BinaryReduceFunction::op_code, result, queue, typename is_device_iterator<OutputIterator>::type());
}

// binary
template<class InputIterator1, class InputIterator2, class OutputIterator, class BinaryTransformFunction, class BinaryReduceFunction>
inline typename std::enable_if<is_device_iterator<InputIterator1>::value && is_device_iterator<InputIterator2>::value>::type
transform_reduce(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, OutputIterator result, BinaryTransformFunction,
                 BinaryReduceFunction, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator1>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "transform_reduce(): scalar value types only");
    static_assert(std::is_same<T, typename std::iterator_traits<InputIterator2>::value_type>::value,
                  "transform_reduce(): both ranges must have the same value type");
    detail::run_transform_reduce<T>(first1.device_ptr(), first2.device_ptr(), detail::iterator_range_size(first1, last1),
                                    BinaryTransformFunction::op_code, BinaryReduceFunction::op_code, result, queue,
                                    typename is_device_iterator<OutputIterator>::type());
}

// inner_product(first1, last1, first2, init): init + sum of x_i * y_i, returned as T = decltype(init)
template<class InputIterator1, class InputIterator2, class T>
inline T inner_product(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, T init,
                       command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator1>::value_type V;
    V sum = V();
    ::boost::compute::transform_reduce(first1, last1, first2, &sum, multiplies<V>(), plus<V>(), queue);
    return first1 == last1 ? init : static_cast<T>(plus<V>()(static_cast<V>(init), sum));
}

// inner_product with explicit accumulate / transform functions (inner_product.hpp:66-97)
template<class InputIterator1, class InputIterator2, class T, class BinaryAccumulateFunction, class BinaryTransformFunction>
inline T inner_product(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, T init, BinaryAccumulateFunction accumulate_function,
                       BinaryTransformFunction transform_function, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator1>::value_type V;
    V folded = V();
    ::boost::compute::transform_reduce(first1, last1, first2, &folded, transform_function, accumulate_function, queue);
    return first1 == last1 ? init : static_cast<T>(accumulate_function(static_cast<V>(init), folded));
}

} // namespace compute
} // namespace boost

#endif
