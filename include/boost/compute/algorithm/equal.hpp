// equal() (algorithm/equal.hpp:30-47 of the reference: mismatch(...) == last) for ranges of the same scalar type:
// one fused pass, OR-reduction of first1[i] XOR first2[i] over the bit patterns (so -0.0 and +0.0 differ and a NaN
// equals itself, unlike operator==; the callers on this path -- is_permutation -- compare sorted integer / char ranges).
#ifndef B200_BOOST_COMPUTE_ALGORITHM_EQUAL_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_EQUAL_HPP

#include <iterator>

#include <boost/compute/algorithm/transform_reduce.hpp>

namespace boost {
namespace compute {

template<class InputIterator1, class InputIterator2>
inline bool equal(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator1>::value && is_device_iterator<InputIterator2>::value, "equal(): device ranges required");
    typedef typename std::iterator_traits<InputIterator1>::value_type T;
    static_assert(std::is_same<T, typename std::iterator_traits<InputIterator2>::value_type>::value, "equal(): same value type required");
    static_assert(detail::dtype_of<T>::supported, "equal(): scalar value types only");
    const size_t n = detail::iterator_range_size(first1, last1);
    if(n == 0){
        return true;
    }
    // compare as unsigned integers of the same width
    const int code = sizeof(T) == 1 ? BCB_UCHAR : (sizeof(T) == 2 ? BCB_USHORT : (sizeof(T) == 4 ? BCB_UINT : BCB_ULONG));
    unsigned long long diff = 0;
    queue.make_current();
    detail::check(bcb_transform_reduce(queue.get(), code, first1.device_ptr(), first2.device_ptr(), n, BCB_BIT_XOR, BCB_BIT_OR, &diff, 0));
    return diff == 0;
}

} // namespace compute
} // namespace boost

#endif
