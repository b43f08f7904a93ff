// min_element() (algorithm/min_element.hpp:36-80; detail/find_extrema_with_reduce.hpp:77-316): iterator to the FIRST smallest
// element of a device range (ties: the smaller index, :156-158); `first` for an empty range (test_extrema.cpp:39-51).
// One (value, index) reduction launch; the iterator is a host value: blocks.  The comparison is less<T> (the default and
// the only one compiled ahead of time).
#ifndef B200_BOOST_COMPUTE_ALGORITHM_MIN_ELEMENT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_MIN_ELEMENT_HPP

#include <iterator>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {

template<class InputIterator>
inline InputIterator min_element(InputIterator first, InputIterator last, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "min_element(): device range required");
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "min_element(): scalar value types only");
    size_t index = 0;
    queue.make_current();
    detail::check(bcb_find_extremum(queue.get(), detail::dtype_of<T>::value, first.device_ptr(), detail::iterator_range_size(first, last), 0, &index));
    return first + static_cast<typename std::iterator_traits<InputIterator>::difference_type>(index);
}

template<class InputIterator, class T>
inline InputIterator min_element(InputIterator first, InputIterator last, less<T>, command_queue &queue = system::default_queue())
{
    return ::boost::compute::min_element(first, last, queue);
}

} // namespace compute
} // namespace boost

#endif
