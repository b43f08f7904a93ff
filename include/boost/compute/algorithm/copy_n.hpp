#ifndef B200_BOOST_COMPUTE_ALGORITHM_COPY_N_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_COPY_N_HPP
#include <boost/compute/algorithm/copy.hpp>
#endif
