// count_if() / count() (algorithm/count_if.hpp:31-58, count.hpp:32-59, detail/count_if_with_reduce.hpp:27-80): the
// number of elements satisfying the predicate (count: equal to a value).  Host return value: blocks.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_COUNT_IF_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_COUNT_IF_HPP

#include <iterator>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/lambda/placeholders.hpp>

namespace boost {
namespace compute {

template<class InputIterator>
inline size_t count_if(InputIterator first, InputIterator last, const lambda::predicate_expr &predicate,
                       command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator>::value, "count_if(): device range required");
    typedef typename std::iterator_traits<InputIterator>::value_type T;
    static_assert(detail::dtype_of<T>::supported, "count_if(): scalar value types only");
    const bcb_pred p = predicate.encode<T>();
    unsigned long long count = 0;
    queue.make_current();
    detail::check(bcb_count_if(queue.get(), detail::dtype_of<T>::value, first.device_ptr(), detail::iterator_range_size(first, last), &p, &count));
    return static_cast<size_t>(count);
}

template<class InputIterator, class T>
inline size_t count(InputIterator first, InputIterator last, const T &value, command_queue &queue = system::default_queue())
{
    return ::boost::compute::count_if(first, last, lambda::_1 == value, queue);
}

} // namespace compute
} // namespace boost

#endif
