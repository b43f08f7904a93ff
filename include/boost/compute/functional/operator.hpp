// Functor tags (functional/operator.hpp:57-96 of the reference).  In the reference these generate OpenCL C
// source at run time; here they are empty tag types that select an ahead-of-time compiled kernel through
// the bcb_op code, and remain callable on the host.
#ifndef B200_BOOST_COMPUTE_FUNCTIONAL_OPERATOR_HPP
#define B200_BOOST_COMPUTE_FUNCTIONAL_OPERATOR_HPP

#include <compute_b200.h>

namespace boost {
namespace compute {

#define BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(name, code, expr)            \
    template<class T>                                                      \
    struct name                                                            \
    {                                                                      \
        typedef T result_type;                                             \
        typedef T argument_type;                                           \
        static const int op_code = code;                                   \
        T operator()(const T &x, const T &y) const { return expr; }        \
    };

BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(plus, BCB_PLUS, static_cast<T>(x + y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(minus, BCB_MINUS, static_cast<T>(x - y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(multiplies, BCB_MULTIPLIES, static_cast<T>(x * y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(divides, BCB_DIVIDES, static_cast<T>(x / y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(bit_and, BCB_BIT_AND, static_cast<T>(x & y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(bit_or, BCB_BIT_OR, static_cast<T>(x | y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(bit_xor, BCB_BIT_XOR, static_cast<T>(x ^ y))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(min, BCB_MIN, (y < x ? y : x))
BOOST_COMPUTE_B200_DECLARE_BINARY_TAG(max, BCB_MAX, (x < y ? y : x))

#undef BOOST_COMPUTE_B200_DECLARE_BINARY_TAG

template<class T>
struct less
{
    typedef bool result_type;
    bool operator()(const T &x, const T &y) const { return x < y; }
};

template<class T>
struct greater
{
    typedef bool result_type;
    bool operator()(const T &x, const T &y) const { return x > y; }
};

} // namespace compute
} // namespace boost

#endif
