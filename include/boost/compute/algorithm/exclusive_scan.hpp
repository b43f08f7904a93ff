// exclusive_scan() (algorithm/exclusive_scan.hpp:55-104 of the reference):
// result[i] = init op first[0] op ... op first[i-1]; default init 0, default op plus<output value type>.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_EXCLUSIVE_SCAN_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_EXCLUSIVE_SCAN_HPP

#include <iterator>

#include <boost/compute/algorithm/detail/scan.hpp>
#include <boost/compute/detail/default_queue.hpp>

namespace boost {
namespace compute {

template<class InputIterator, class OutputIterator, class T, class BinaryOperator>
inline OutputIterator exclusive_scan(InputIterator first, InputIterator last, OutputIterator result, T init,
                                     BinaryOperator binary_op, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "exclusive_scan(): device input range required");
    static_assert(is_device_iterator<OutputIterator>::value, "exclusive_scan(): device output range required");
    return detail::scan(first, last, result, true, init, binary_op, queue);
}

template<class InputIterator, class OutputIterator, class T>
inline typename std::enable_if<!std::is_same<T, command_queue>::value, OutputIterator>::type
exclusive_scan(InputIterator first, InputIterator last, OutputIterator result, T init,
               command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<OutputIterator>::value_type output_type;
    return ::boost::compute::exclusive_scan(first, last, result, init, plus<output_type>(), queue);
}

template<class InputIterator, class OutputIterator>
inline OutputIterator exclusive_scan(InputIterator first, InputIterator last, OutputIterator result,
                                     command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<OutputIterator>::value_type output_type;
    return ::boost::compute::exclusive_scan(first, last, result, output_type(0), plus<output_type>(), queue);
}

} // namespace compute
} // namespace boost

#endif
