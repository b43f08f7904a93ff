#ifndef B200_BOOST_COMPUTE_ALGORITHM_COUNT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_COUNT_HPP
#include <boost/compute/algorithm/count_if.hpp>  // count(first, last, value) = count_if(_1 == value) (count.hpp:32-59)
#endif
