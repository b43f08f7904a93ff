// copy / copy_n (algorithm/copy.hpp:178-735, copy_n.hpp of the reference) for the three directions the path
// needs: host -> device, device -> host (both blocking, like enqueue_*_buffer in the reference) and
// device -> device (enqueued).  Host ranges that are not plain pointers are staged through a std::vector.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_COPY_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_COPY_HPP

#include <iterator>
#include <type_traits>
#include <vector>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

// host -> device
template<class HostIterator, class T>
inline buffer_iterator<T> copy_to_device(HostIterator first, HostIterator last, buffer_iterator<T> result,
                                         command_queue &queue)
{
    typedef typename std::iterator_traits<HostIterator>::value_type H;
    const std::size_t n = iterator_range_size(first, last);
    if(n == 0){
        return result;
    }
    if(std::is_pointer<HostIterator>::value && std::is_same<typename std::remove_cv<H>::type, T>::value){
        queue.enqueue_write_buffer(result.get_buffer(), result.get_index() * sizeof(T), n * sizeof(T), &*first);
    } else {
        std::vector<T> staging(first, last); // converts element type if needed
        queue.enqueue_write_buffer(result.get_buffer(), result.get_index() * sizeof(T), n * sizeof(T), &staging[0]);
    }
    return result + n;
}

// device -> host
template<class T, class HostIterator>
inline HostIterator copy_to_host(buffer_iterator<T> first, buffer_iterator<T> last, HostIterator result,
                                 command_queue &queue)
{
    typedef typename std::iterator_traits<HostIterator>::value_type H;
    const std::size_t n = iterator_range_size(first, last);
    if(n == 0){
        return result;
    }
    if(std::is_pointer<HostIterator>::value && std::is_same<H, T>::value){
        queue.enqueue_read_buffer(first.get_buffer(), first.get_index() * sizeof(T), n * sizeof(T), &*result);
        std::advance(result, n);
        return result;
    }
    std::vector<T> staging(n);
    queue.enqueue_read_buffer(first.get_buffer(), first.get_index() * sizeof(T), n * sizeof(T), &staging[0]);
    return std::copy(staging.begin(), staging.end(), result);
}

// device -> device
template<class T>
inline buffer_iterator<T> copy_on_device(buffer_iterator<T> first, buffer_iterator<T> last, buffer_iterator<T> result,
                                         command_queue &queue)
{
    const std::size_t n = iterator_range_size(first, last);
    if(n != 0){
        queue.enqueue_copy_buffer(first.get_buffer(), result.get_buffer(), first.get_index() * sizeof(T),
                                  result.get_index() * sizeof(T), n * sizeof(T));
    }
    return result + n;
}

template<class In, class Out>
inline Out dispatch_copy(In first, In last, Out result, command_queue &queue, std::false_type, std::true_type)
{
    return copy_to_device(first, last, result, queue);
}
template<class In, class Out>
inline Out dispatch_copy(In first, In last, Out result, command_queue &queue, std::true_type, std::false_type)
{
    return copy_to_host(first, last, result, queue);
}
template<class In, class Out>
inline Out dispatch_copy(In first, In last, Out result, command_queue &queue, std::true_type, std::true_type)
{
    return copy_on_device(first, last, result, queue);
}

} // namespace detail

template<class InputIterator, class OutputIterator>
inline OutputIterator copy(InputIterator first, InputIterator last, OutputIterator result,
                           command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value || is_device_iterator<OutputIterator>::value,
                  "copy(): at least one side must be a device iterator");
    return detail::dispatch_copy(first, last, result, queue,
                                 typename is_device_iterator<InputIterator>::type(),
                                 typename is_device_iterator<OutputIterator>::type());
}

template<class InputIterator, class Size, class OutputIterator>
inline OutputIterator copy_n(InputIterator first, Size count, OutputIterator result,
                             command_queue &queue = system::default_queue())
{
    return ::boost::compute::copy(first, first + count, result, queue);
}

} // namespace compute
} // namespace boost

#endif
