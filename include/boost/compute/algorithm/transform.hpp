// transform() (algorithm/transform.hpp:30-75 of the reference), unary form with the closed function set
// (identity / negate / abs / square): result[i] = function(first[i]).  Runs as transform_if with an always-true
// predicate: one pass, order preserved.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_TRANSFORM_HPP

#include <boost/compute/algorithm/transform_if.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class OutputIterator, class UnaryFunction>
inline OutputIterator transform(InputIterator first, InputIterator last, OutputIterator result, UnaryFunction function,
                                command_queue &queue = system::default_queue())
{
    lambda::predicate_expr always = { BCB_AR_NONE, 0, BCB_CMP_TRUE, 0 };
    return ::boost::compute::transform_if(first, last, result, function, always, queue);
}

} // namespace compute
} // namespace boost

#endif
