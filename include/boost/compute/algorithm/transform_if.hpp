// transform_if() (algorithm/transform_if.hpp:42-117 of the reference) and copy_if() (copy_if.hpp:28-52): copies
// function(x) for every x of [first, last) that satisfies the predicate to result, preserving order; returns the end
// of the output.  The reference runs three sweeps (flags into an n-element index vector, exclusive_scan, scatter);
// here it is ONE kernel with a decoupled look-back over the per-tile counts.  The returned iterator needs the count
// on the host, so the call blocks (the reference reads the last scanned index back the same way, :68-73).
#ifndef B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_IF_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_IF_HPP

#include <iterator>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/lambda/placeholders.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class OutputIterator, class UnaryFunction>
inline OutputIterator transform_if(InputIterator first, InputIterator last, OutputIterator result, UnaryFunction,
                                   const lambda::predicate_expr &predicate, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value && is_device_iterator<OutputIterator>::value,
                  "transform_if(): device ranges required");
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "transform_if(): scalar value types only");
    static_assert(std::is_same<T, typename std::iterator_traits<OutputIterator>::value_type>::value,
                  "transform_if(): input and output value types must match");
    const bcb_pred p = predicate.encode<T>();
    size_t count = 0;
    queue.make_current();
    detail::check(bcb_transform_if(queue.get(), detail::dtype_of<T>::value, first.device_ptr(), detail::iterator_range_size(first, last),
                                   UnaryFunction::unary_code, &p, result.device_ptr(), &count));
    return result + static_cast<typename std::iterator_traits<OutputIterator>::difference_type>(count);
}

template<class InputIterator, class OutputIterator>
inline OutputIterator copy_if(InputIterator first, InputIterator last, OutputIterator result,
                              const lambda::predicate_expr &predicate, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    return ::boost::compute::transform_if(first, last, result, identity<T>(), predicate, queue);
}

} // namespace compute
} // namespace boost

#endif
