#ifndef B200_BOOST_COMPUTE_CORE_HPP
#define B200_BOOST_COMPUTE_CORE_HPP
#include <boost/compute/buffer.hpp>
#include <boost/compute/command_queue.hpp>
#include <boost/compute/context.hpp>
#include <boost/compute/device.hpp>
#include <boost/compute/exception.hpp>
#include <boost/compute/system.hpp>
#endif
