// detail::merge_sort_on_gpu(first, last, compare, [stable,] queue) (algorithm/detail/merge_sort_on_gpu.hpp:523-572 of the
// reference): the comparison sort behind sort() / stable_sort() with a custom comparator.  The reference generates a
// block-wise merge sort around the comparator's OpenCL source; here the comparator comes from the closed family
// f(a.field) < f(b.field) (functional/field.hpp), for which the stable comparison sort IS a stable key-value radix sort:
// bcb_sort_by_field projects the field into a key array and sorts the keys with the records as payload.  Always stable
// (the `stable` flag of the reference is accepted and ignored).  less<T> / greater<T> go to the radix sort directly.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_DETAIL_MERGE_SORT_ON_GPU_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_DETAIL_MERGE_SORT_ON_GPU_HPP

#include <type_traits>

#include <boost/compute/algorithm/detail/radix_sort.hpp>
#include <boost/compute/command_queue.hpp>
#include <boost/compute/functional/field.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {
namespace detail {

template<class T, class Compare>
inline typename std::enable_if<is_field_compare<Compare>::value>::type
merge_sort_on_gpu(buffer_iterator<T> first, buffer_iterator<T> last, Compare compare, bool /*stable*/, command_queue &queue)
{
    static_assert(std::is_trivially_copyable<T>::value, "records are moved as bytes");
    const field_spec f = compare.template resolve<T>();
    queue.make_current();
    check(bcb_sort_by_field(queue.get(), first.device_ptr(), iterator_range_size(first, last), sizeof(T), f.offset, f.dtype, f.unary,
                            f.descending ? 1 : 0));
}

template<class T>
inline void merge_sort_on_gpu(buffer_iterator<T> first, buffer_iterator<T> last, less<T>, bool, command_queue &queue)
{
    radix_sort(first, last, true, queue);
}

template<class T>
inline void merge_sort_on_gpu(buffer_iterator<T> first, buffer_iterator<T> last, greater<T>, bool, command_queue &queue)
{
    radix_sort(first, last, false, queue);
}

template<class Iterator, class Compare>
inline void merge_sort_on_gpu(Iterator first, Iterator last, Compare compare, command_queue &queue)
{
    merge_sort_on_gpu(first, last, compare, false, queue);
}

} // namespace detail
} // namespace compute
} // namespace boost

#endif
