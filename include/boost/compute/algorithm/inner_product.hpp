#ifndef B200_BOOST_COMPUTE_ALGORITHM_INNER_PRODUCT_HPP
#define B200_BOOST_COMPUTE_ALGORITHM_INNER_PRODUCT_HPP
#include <boost/compute/algorithm/transform_reduce.hpp>  // inner_product (inner_product.hpp:40-97) lives next to transform_reduce
#endif
