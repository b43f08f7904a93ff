// detail::scan (algorithm/detail/scan.hpp:22-39 of the reference): the single entry both scans funnel into.
// The reference switches between scan_on_cpu and the recursive scan_on_gpu; here it is one launch of the
// single-pass decoupled look-back kernel behind bcb_scan.  Arithmetic happens in the OUTPUT value type
// (exclusive_scan.hpp:80-85); `first == result` (in place) is allowed.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_DETAIL_SCAN_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_DETAIL_SCAN_HPP

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class TIn, class TOut, class T, class BinaryOperator>
inline buffer_iterator<TOut> scan(buffer_iterator<TIn> first, buffer_iterator<TIn> last, buffer_iterator<TOut> result,
                                  bool exclusive, T init, BinaryOperator, command_queue &queue)
{
    static_assert(dtype_of<TIn>::supported && dtype_of<TOut>::supported, "scan(): scalar value types only");
    const std::size_t n = iterator_range_size(first, last);
    if(n == 0){
        return result; // scan_on_gpu.hpp:316-318
    }
    const TOut init_value = static_cast<TOut>(init);
    queue.make_current();
    check(bcb_scan(queue.get(), dtype_of<TIn>::value, dtype_of<TOut>::value, BinaryOperator::op_code, exclusive ? 1 : 0,
                   first.device_ptr(), result.device_ptr(), n, &init_value));
    return result + n;
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
