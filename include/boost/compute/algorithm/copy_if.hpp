#ifndef B200_BOOST_COMPUTE_ALGORITHM_COPY_IF_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_COPY_IF_HPP
#include <boost/compute/algorithm/transform_if.hpp>  // copy_if = transform_if with identity (copy_if.hpp:28-52)
#endif
