#ifndef B200_BOOST_COMPUTE_FUNCTIONAL_HPP
#define B200_BOOST_COMPUTE_FUNCTIONAL_HPP
#include <boost/compute/functional/field.hpp>
#include <boost/compute/functional/operator.hpp>
#endif
