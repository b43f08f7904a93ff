#ifndef B200_BOOST_COMPUTE_EXCEPTION_HPP
#define B200_BOOST_COMPUTE_EXCEPTION_HPP
#include <boost/compute/exception/opencl_error.hpp>
#endif
