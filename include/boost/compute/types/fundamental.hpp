// Scalar type aliases of the reference (boost/compute/types/fundamental.hpp:30-39): char_ ... double_,
// here plain fixed-width C++ types instead of cl_* typedefs.
#ifndef B200_BOOST_COMPUTE_TYPES_FUNDAMENTAL_HPP
#define B200_BOOST_COMPUTE_TYPES_FUNDAMENTAL_HPP

#include <cstdint>

namespace boost {
namespace compute {

typedef std::int8_t char_;
typedef std::uint8_t uchar_;
typedef std::int16_t short_;
typedef std::uint16_t ushort_;
typedef std::int32_t int_;
typedef std::uint32_t uint_;
typedef std::int64_t long_;
typedef std::uint64_t ulong_;
typedef float float_;
typedef double double_;

} // namespace compute
} // namespace boost

#endif
