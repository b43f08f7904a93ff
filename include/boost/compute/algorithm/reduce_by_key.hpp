// reduce_by_key() (algorithm/reduce_by_key.hpp:60-118 of the reference): every run of consecutive equal keys becomes one
// (key, fold of its values) pair; returns the pair of output end iterators (a host value: blocks).  The reference
// (detail/reduce_by_key_with_scan.hpp:48-97) computes head flags, an inclusive scan by key and a scatter with O(2n)
// temporaries; here ONE kernel folds segments through a segmented decoupled look-back.  The key predicate is equality
// (equal_to<K>, the reference's default); functions: plus, multiplies, min, max, bit_and / bit_or / bit_xor.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_REDUCE_BY_KEY_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_REDUCE_BY_KEY_HPP

#include <iterator>
#include <utility>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/functional/operator.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/lambda/placeholders.hpp>

namespace boost {
namespace compute {

template<class InputKeyIterator, class InputValueIterator, class OutputKeyIterator, class OutputValueIterator, class BinaryFunction>
inline typename std::enable_if<!std::is_same<BinaryFunction, command_queue>::value, std::pair<OutputKeyIterator, OutputValueIterator> >::type
reduce_by_key(InputKeyIterator keys_first, InputKeyIterator keys_last, InputValueIterator values_first, OutputKeyIterator keys_result,
              OutputValueIterator values_result, BinaryFunction, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputKeyIterator>::value && is_device_iterator<InputValueIterator>::value &&
                  is_device_iterator<OutputKeyIterator>::value && is_device_iterator<OutputValueIterator>::value,
                  "reduce_by_key(): device ranges required");
    typedef typename std::iterator_traits<InputKeyIterator>::value_type K;
    typedef typename std::iterator_traits<InputValueIterator>::value_type V;
    static_assert(detail::dtype_of<K>::supported && detail::dtype_of<V>::supported, "reduce_by_key(): scalar key and value types only");
    size_t count = 0;
    queue.make_current();
    detail::check(bcb_reduce_by_key(queue.get(), detail::dtype_of<K>::value, detail::dtype_of<V>::value, keys_first.device_ptr(),
                                    values_first.device_ptr(), detail::iterator_range_size(keys_first, keys_last), keys_result.device_ptr(),
                                    values_result.device_ptr(), BinaryFunction::op_code, &count));
    typedef typename std::iterator_traits<OutputKeyIterator>::difference_type D;
    return std::make_pair(keys_result + static_cast<D>(count), values_result + static_cast<D>(count));
}

// with an explicit key predicate: only equality is compiled (the reference's default)
template<class InputKeyIterator, class InputValueIterator, class OutputKeyIterator, class OutputValueIterator, class BinaryFunction, class K>
inline std::pair<OutputKeyIterator, OutputValueIterator>
reduce_by_key(InputKeyIterator keys_first, InputKeyIterator keys_last, InputValueIterator values_first, OutputKeyIterator keys_result,
              OutputValueIterator values_result, BinaryFunction function, equal_to<K>, command_queue &queue = system::default_queue())
{
    return ::boost::compute::reduce_by_key(keys_first, keys_last, values_first, keys_result, values_result, function, queue);
}

template<class InputKeyIterator, class InputValueIterator, class OutputKeyIterator, class OutputValueIterator>
inline std::pair<OutputKeyIterator, OutputValueIterator>
reduce_by_key(InputKeyIterator keys_first, InputKeyIterator keys_last, InputValueIterator values_first, OutputKeyIterator keys_result,
              OutputValueIterator values_result, command_queue &queue = system::default_queue())
{
    typedef typename std::iterator_traits<InputValueIterator>::value_type V;
    return ::boost::compute::reduce_by_key(keys_first, keys_last, values_first, keys_result, values_result, plus<V>(), queue);
}

} // namespace compute
} // namespace boost

#endif
