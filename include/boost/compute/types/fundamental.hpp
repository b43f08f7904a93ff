// Scalar type aliases of the reference (boost/compute/types/fundamental.hpp:30-39): char_ ... double_,
// here plain fixed-width C++ types instead of cl_* typedefs.
#ifndef B200_BOOST_COMPUTE_TYPES_FUNDAMENTAL_HPP
#define B200_BOOST_COMPUTE_TYPES_FUNDAMENTAL_HPP

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

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

// Vector types (types/fundamental.hpp:41-172 of the reference): plain aggregates of N scalars with the (x, y), (x, y, z, w),
// (s0..s7), (s0..sf) members.  On this path they are record types: containers hold them, copies move them, and the
// sorts with a field comparator (functional/field.hpp) order them by one component.
namespace detail {

template<class Scalar, std::size_t N> struct vector_type_desc;
template<class Scalar> struct vector_type_desc<Scalar, 2> { Scalar x, y; };
template<class Scalar> struct vector_type_desc<Scalar, 4> { Scalar x, y, z, w; };
template<class Scalar> struct vector_type_desc<Scalar, 8> { Scalar s0, s1, s2, s3, s4, s5, s6, s7; };
template<class Scalar> struct vector_type_desc<Scalar, 16> { Scalar s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, sa, sb, sc, sd, se, sf; };

} // namespace detail

template<class Scalar, std::size_t N>
class vector_type : public detail::vector_type_desc<Scalar, N>
{
public:
    typedef Scalar scalar_type;

    vector_type() { for(std::size_t i = 0; i < N; i++) (*this)[i] = Scalar(); }
    explicit vector_type(const Scalar s) { for(std::size_t i = 0; i < N; i++) (*this)[i] = s; }
    template<class... Rest, class = typename std::enable_if<sizeof...(Rest) + 2 == N>::type>
    vector_type(const Scalar a, const Scalar b, const Rest... rest)
    {
        const Scalar v[N] = { a, b, static_cast<Scalar>(rest)... };
        for(std::size_t i = 0; i < N; i++) (*this)[i] = v[i];
    }

    std::size_t size() const { return N; }
    Scalar &operator[](std::size_t i) { return reinterpret_cast<Scalar *>(this)[i]; }
    Scalar operator[](std::size_t i) const { return reinterpret_cast<const Scalar *>(this)[i]; }
    bool operator==(const vector_type &o) const { return std::memcmp(this, &o, sizeof(Scalar) * N) == 0; }
    bool operator!=(const vector_type &o) const { return !(*this == o); }
};

#define BOOST_COMPUTE_B200_VECTOR_TYPES(scalar)                      \
    typedef vector_type<scalar##_, 2> scalar##2_;                    \
    typedef vector_type<scalar##_, 4> scalar##4_;                    \
    typedef vector_type<scalar##_, 8> scalar##8_;                    \
    typedef vector_type<scalar##_, 16> scalar##16_;
BOOST_COMPUTE_B200_VECTOR_TYPES(char)
BOOST_COMPUTE_B200_VECTOR_TYPES(uchar)
BOOST_COMPUTE_B200_VECTOR_TYPES(short)
BOOST_COMPUTE_B200_VECTOR_TYPES(ushort)
BOOST_COMPUTE_B200_VECTOR_TYPES(int)
BOOST_COMPUTE_B200_VECTOR_TYPES(uint)
BOOST_COMPUTE_B200_VECTOR_TYPES(long)
BOOST_COMPUTE_B200_VECTOR_TYPES(ulong)
BOOST_COMPUTE_B200_VECTOR_TYPES(float)
BOOST_COMPUTE_B200_VECTOR_TYPES(double)
#undef BOOST_COMPUTE_B200_VECTOR_TYPES

template<class T> struct is_vector_type : std::false_type {};
template<class Scalar, std::size_t N> struct is_vector_type<vector_type<Scalar, N> > : std::true_type {};

} // namespace compute
} // namespace boost

#endif
