#ifndef B200_BOOST_COMPUTE_ALGORITHM_FILL_N_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_FILL_N_HPP
#include <boost/compute/algorithm/fill.hpp>
#endif
