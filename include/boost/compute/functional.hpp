#ifndef BOOST_COMPUTE_FUNCTIONAL_HPP
#define BOOST_COMPUTE_FUNCTIONAL_HPP
#include <boost/compute/functional/operator.hpp>
#endif
