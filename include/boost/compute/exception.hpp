#ifndef BOOST_COMPUTE_EXCEPTION_HPP
#define BOOST_COMPUTE_EXCEPTION_HPP
#include <boost/compute/exception/opencl_error.hpp>
#endif
