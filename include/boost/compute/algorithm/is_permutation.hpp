// is_permutation() (algorithm/is_permutation.hpp:43-67 of the reference): copies of both ranges are sorted and compared.
#ifndef B200_BOOST_COMPUTE_ALGORITHM_IS_PERMUTATION_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_IS_PERMUTATION_HPP

#include <iterator>

#include <boost/compute/algorithm/equal.hpp>
#include <boost/compute/algorithm/sort.hpp>
#include <boost/compute/container/vector.hpp>

namespace boost {
namespace compute {

template<class InputIterator1, class InputIterator2>
inline bool is_permutation(InputIterator1 first1, InputIterator1 last1, InputIterator2 first2, InputIterator2 last2,
                           command_queue &queue = system::default_queue())
{
    static_assert(is_device_iterator<InputIterator1>::value && is_device_iterator<InputIterator2>::value,
                  "is_permutation(): device ranges required");
    typedef typename std::iterator_traits<InputIterator1>::value_type value_type1;
    typedef typename std::iterator_traits<InputIterator2>::value_type value_type2;
    if(detail::iterator_range_size(first1, last1) != detail::iterator_range_size(first2, last2)){
        return false;
    }
    vector<value_type1> temp1(first1, last1, queue);
    vector<value_type2> temp2(first2, last2, queue);
    ::boost::compute::sort(temp1.begin(), temp1.end(), queue);
    ::boost::compute::sort(temp2.begin(), temp2.end(), queue);
    return ::boost::compute::equal(temp1.begin(), temp1.end(), temp2.begin(), queue);
}

} // namespace compute
} // namespace boost

#endif
