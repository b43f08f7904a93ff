// minmax_element() (algorithm/minmax_element.hpp:33-66): (min_element, max_element) of a device range.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_MINMAX_ELEMENT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_MINMAX_ELEMENT_HPP

#include <utility>

#include <boost/compute/algorithm/max_element.hpp>
#include <boost/compute/algorithm/min_element.hpp>

namespace boost {
namespace compute {

template<class InputIterator>
inline std::pair<InputIterator, InputIterator> minmax_element(InputIterator first, InputIterator last,
                                                              command_queue &queue = system::default_queue())
{
    return std::make_pair(::boost::compute::min_element(first, last, queue), ::boost::compute::max_element(first, last, queue));
}

template<class InputIterator, class T>
inline std::pair<InputIterator, InputIterator> minmax_element(InputIterator first, InputIterator last, less<T>,
                                                              command_queue &queue = system::default_queue())
{
    return ::boost::compute::minmax_element(first, last, queue);
}

} // namespace compute
} // namespace boost

#endif
