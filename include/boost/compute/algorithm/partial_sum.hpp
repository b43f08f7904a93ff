// partial_sum() (algorithm/partial_sum.hpp:31-41 of the reference) == inclusive_scan with plus.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_PARTIAL_SUM_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_PARTIAL_SUM_HPP

#include <boost/compute/algorithm/inclusive_scan.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class OutputIterator>
inline OutputIterator partial_sum(InputIterator first, InputIterator last, OutputIterator result,
                                  command_queue &queue = system::default_queue())
{
    return ::boost::compute::inclusive_scan(first, last, result, queue);
}

} // namespace compute
} // namespace boost

#endif
