#ifndef B200_BOOST_COMPUTE_CONTAINER_HPP
#define B200_BOOST_COMPUTE_CONTAINER_HPP
#include <boost/compute/container/array.hpp>
#include <boost/compute/container/mapped_view.hpp>
#include <boost/compute/container/valarray.hpp>
#include <boost/compute/container/vector.hpp>
#endif
