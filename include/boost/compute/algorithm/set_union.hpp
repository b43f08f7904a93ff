// set_union() (algorithm/set_union.hpp:120-199): every element that is in either range (max of the two multiplicities).
// Both input ranges must be sorted; std::set_union multiset semantics (equal elements: first range first).  The reference
// tiles the two ranges by balanced path, flags, scans and scatters; here: flags by binary search -> the library's single-
// pass scan -> scatter (compute_b200/csrc/set_ops.cu).  Returns result + count, a host value: blocks.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_SET_UNION_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_SET_UNION_HPP

#include <boost/compute/algorithm/detail/set_operation.hpp>

namespace boost {
namespace compute {

template<class InputIterator1, class InputIterator2, class OutputIterator>
inline OutputIterator set_union(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, InputIterator2 last2,
                             OutputIterator result, command_queue &queue = system::default_queue())
{
    return detail::set_operation(BCB_SET_UNION, first1, last1, first2, last2, result, queue);
}

} // namespace compute
} // namespace boost

#endif
