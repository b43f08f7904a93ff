// experimental::sort_by_transform() (experimental/sort_by_transform.hpp:26-63 of the reference): sorts [first, last) by
// the key transform(x): keys = transform(range), then sort_by_key(keys, range, compare).  Transform: the closed unary
// set (identity / negate / abs / square); compare: less / greater.
#ifndef B200_BOOST_COMPUTE_EXPERIMENTAL_SORT_BY_TRANSFORM_HPP
#define B200_BOOST_COMPUTE_EXPERIMENTAL_SORT_BY_TRANSFORM_HPP

#include <iterator>

#include <boost/compute/algorithm/sort_by_key.hpp>
#include <boost/compute/algorithm/transform.hpp>
#include <boost/compute/container/vector.hpp>

namespace boost {
namespace compute {
namespace experimental {

template<class Iterator, class Transform, class Compare>
inline void sort_by_transform(Iterator first, Iterator last, Transform transform, Compare compare,
                              command_queue &queue = system::default_queue())
{
    typedef typename Transform::result_type key_type;
    const size_t n = detail::iterator_range_size(first, last);
    if(n < 2){
        return;
    }
    ::boost::compute::vector<key_type> keys(n, queue.get_context());
    ::boost::compute::transform(first, last, keys.begin(), transform, queue);
    ::boost::compute::sort_by_key(keys.begin(), keys.end(), first, compare, queue);
}

} // namespace experimental
} // namespace compute
} // namespace boost

#endif
