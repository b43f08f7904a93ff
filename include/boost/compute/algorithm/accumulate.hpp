// accumulate() (algorithm/accumulate.hpp:165-188 of the reference): returns init op first[0] op first[1] ...
// as a host value of init's type T.  The reference only takes the parallel reduce() path for integer
// plus / multiplies with the identity as init and for min / max with extreme inits (:62-98), and otherwise runs
// a single-work-item left fold.  Here every associative functor whose type equals T runs the parallel kernel
// with init folded in afterwards (exact for integers; float sums differ from the serial fold by summation
// order only); minus / divides and mixed T keep the strict serial left fold (serial_accumulate.hpp:22-50).
#ifndef B200_BOOST_COMPUTE_ALGORITHM_ACCUMULATE_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_ACCUMULATE_HPP

#include <iterator>
#include <type_traits>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class T, class BinaryFunction>
inline typename std::enable_if<!std::is_same<BinaryFunction, command_queue>::value, T>::type
accumulate(InputIterator first, InputIterator last, T init, BinaryFunction,
           command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "accumulate(): device input range required");
    typedef typename std::iterator_traits<InputIterator>::value_type IT;
    typedef typename BinaryFunction::argument_type F;
    static_assert(detail::dtype_of<IT>::supported && detail::dtype_of<T>::supported && detail::dtype_of<F>::supported,
                  "accumulate(): scalar value types only");
    T result = init;
    queue.make_current();
    detail::check(bcb_accumulate(queue.get(), detail::dtype_of<IT>::value, detail::dtype_of<F>::value,
                                 detail::dtype_of<T>::value, BinaryFunction::op_code, first.device_ptr(),
                                 detail::iterator_range_size(first, last), &init, &result));
    return result;
}

template<class InputIterator, class T>
inline T accumulate(InputIterator first, InputIterator last, T init, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputIterator>::value_type IT;
    return ::boost::compute::accumulate(first, last, init, plus<IT>(), queue);
}

} // namespace compute
} // namespace boost

#endif
