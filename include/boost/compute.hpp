// boost/compute.hpp -- umbrella header of the B200-native subset: core + vector + the sort / scan / reduce path.
#ifndef B200_BOOST_COMPUTE_HPP
#define B200_BOOST_COMPUTE_HPP
#include <boost/compute/algorithm.hpp>
#include <boost/compute/container.hpp>
#include <boost/compute/core.hpp>
#include <boost/compute/functional.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/types/fundamental.hpp>
#endif
