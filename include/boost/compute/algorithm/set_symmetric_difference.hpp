// set_symmetric_difference() (algorithm/set_symmetric_difference.hpp:121-199): the elements present in exactly one of the ranges.
// Both input ranges must be sorted; std::set_symmetric_difference multiset semantics (equal elements: first range first).  The reference
// tiles the two ranges by balanced path, flags, scans and scatters; here: flags by binary search -> the library's single-
// pass scan -> scatter (compute_b200/csrc/set_ops.cu).  Returns result + count, a host value: blocks.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_SET_SYMMETRIC_DIFFERENCE_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_SET_SYMMETRIC_DIFFERENCE_HPP

#include <boost/compute/algorithm/detail/set_operation.hpp>

namespace boost {
namespace compute {

template<class InputIterator1, class InputIterator2, class OutputIterator>
inline OutputIterator set_symmetric_difference(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, InputIterator2 last2,
                             OutputIterator result, command_queue &queue = system::default_queue())
{
    return detail::set_operation(BCB_SET_SYMMETRIC_DIFFERENCE, first1, last1, first2, last2, result, queue);
}

} // namespace compute
} // namespace boost

#endif
