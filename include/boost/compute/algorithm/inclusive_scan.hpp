// inclusive_scan() (algorithm/inclusive_scan.hpp:53-87 of the reference): result[i] = first[0] op ... op first[i].
#ifndef B200_BOOST_COMPUTE_ALGORITHM_INCLUSIVE_SCAN_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_INCLUSIVE_SCAN_HPP

#include <iterator>

#include <boost/compute/algorithm/detail/scan.hpp>
#include <boost/compute/detail/default_queue.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class OutputIterator, class BinaryOperator>
inline typename std::enable_if<!std::is_same<BinaryOperator, command_queue>::value, OutputIterator>::type
inclusive_scan(InputIterator first, InputIterator last, OutputIterator result, BinaryOperator binary_op,
               command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "inclusive_scan(): device input range required");
    static_assert(is_device_iterator<OutputIterator>::value, "inclusive_scan(): device output range required");
    typedef typename std::iterator_traits<OutputIterator>::value_type output_type;
    return detail::scan(first, last, result, false, output_type(0), binary_op, queue);
}

template<class InputIterator, class OutputIterator>
inline OutputIterator inclusive_scan(InputIterator first, InputIterator last, OutputIterator result,
                                     command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<OutputIterator>::value_type output_type;
    return ::boost::compute::inclusive_scan(first, last, result, plus<output_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
