// Placeholder expressions (lambda/placeholders.hpp, lambda/functional.hpp of the reference) restricted to the closed,
// ahead-of-time compiled predicate form of the C ABI:  ((_1 ARITH a) CMP b)  with ARITH in {*, %, +, -, &} or absent.
// `_1 < 5`, `_1 * 2 >= 10`, `_1 % 2 == 1`, `_1 % 2 != 0` (test_copy_if.cpp:35-66, test_transform_if.cpp:34,
// test_count.cpp:69) build a predicate_expr; anything else is a compile error (the reference would generate OpenCL C for
// arbitrary expressions at run time; there is no run-time compiler here).
#ifndef B200_BOOST_COMPUTE_LAMBDA_PLACEHOLDERS_HPP
#define B200_BOOST_COMPUTE_LAMBDA_PLACEHOLDERS_HPP

#include <cstring>

#include <compute_b200.h>

#include <boost/compute/exception/opencl_error.hpp>
#include <boost/compute/functional/field.hpp>

namespace boost {
namespace compute {
namespace lambda {

// ((x ARITH a) CMP b); operands are kept as long double and narrowed to the element type when the kernel is chosen
struct predicate_expr
{
    int arith;
    long double a;
    int cmp;
    long double b;

    template<class T>
    bcb_pred encode() const
    {
        bcb_pred p;
        p.arith = arith;
        p.cmp = cmp;
        p.a_bits = 0;
        p.b_bits = 0;
        const T av = static_cast<T>(a), bv = static_cast<T>(b);
        std::memcpy(&p.a_bits, &av, sizeof(T));
        std::memcpy(&p.b_bits, &bv, sizeof(T));
        return p;
    }
};

struct arith_expr
{
    int arith;
    long double a;
};

struct placeholder1
{
};

#define BOOST_COMPUTE_B200_ARITH(op, code)                                                                  \
    template<class S> inline arith_expr operator op(placeholder1, S s) { arith_expr e = { code, static_cast<long double>(s) }; return e; }
BOOST_COMPUTE_B200_ARITH(*, BCB_AR_MUL)
BOOST_COMPUTE_B200_ARITH(%, BCB_AR_MOD)
BOOST_COMPUTE_B200_ARITH(+, BCB_AR_ADD)
BOOST_COMPUTE_B200_ARITH(-, BCB_AR_SUB)
BOOST_COMPUTE_B200_ARITH(&, BCB_AR_AND)
#undef BOOST_COMPUTE_B200_ARITH

#define BOOST_COMPUTE_B200_CMP(op, code)                                                                     \
    template<class S> inline predicate_expr operator op(placeholder1, S s)                                   \
    { predicate_expr p = { BCB_AR_NONE, 0, code, static_cast<long double>(s) }; return p; }                   \
    template<class S> inline predicate_expr operator op(arith_expr e, S s)                                   \
    { predicate_expr p = { e.arith, e.a, code, static_cast<long double>(s) }; return p; }
BOOST_COMPUTE_B200_CMP(==, BCB_CMP_EQ)
BOOST_COMPUTE_B200_CMP(!=, BCB_CMP_NE)
BOOST_COMPUTE_B200_CMP(<, BCB_CMP_LT)
BOOST_COMPUTE_B200_CMP(<=, BCB_CMP_LE)
BOOST_COMPUTE_B200_CMP(>, BCB_CMP_GT)
BOOST_COMPUTE_B200_CMP(>=, BCB_CMP_GE)
#undef BOOST_COMPUTE_B200_CMP

static const placeholder1 _1 = placeholder1();

// binary comparator expressions of the field-comparator family (functional/field.hpp):  f(_1) < f(_2),  f(_1) > f(_2)  with
// f = get<N>, abs, or both sides plain (lambda/get.hpp, lambda/functional.hpp of the reference)
struct placeholder2
{
};
static const placeholder2 _2 = placeholder2();

template<int Arg> struct projected  // get<N>(abs(_Arg)) pieces collected so far
{
    int index;
    int unary;
};
template<int N> inline projected<1> get(placeholder1) { projected<1> p = { N, BCB_UN_IDENTITY }; return p; }
template<int N> inline projected<2> get(placeholder2) { projected<2> p = { N, BCB_UN_IDENTITY }; return p; }
inline projected<1> abs(placeholder1) { projected<1> p = { 0, BCB_UN_ABS }; return p; }
inline projected<2> abs(placeholder2) { projected<2> p = { 0, BCB_UN_ABS }; return p; }
template<int Arg> inline projected<Arg> abs(projected<Arg> p) { p.unary = BCB_UN_ABS; return p; }

namespace detail {
inline component_compare projected_compare(int i1, int u1, int i2, int u2, bool descending)
{
    if(i1 != i2 || u1 != u2){
        throw opencl_error(BCB_EUNSUPPORTED);  // both sides must look at the same field through the same function
    }
    component_compare c = { i1, u1, descending };
    return c;
}
} // namespace detail
inline component_compare operator<(projected<1> a, projected<2> b) { return detail::projected_compare(a.index, a.unary, b.index, b.unary, false); }
inline component_compare operator>(projected<1> a, projected<2> b) { return detail::projected_compare(a.index, a.unary, b.index, b.unary, true); }
inline component_compare operator<(placeholder1, placeholder2) { component_compare c = { 0, BCB_UN_IDENTITY, false }; return c; }
inline component_compare operator>(placeholder1, placeholder2) { component_compare c = { 0, BCB_UN_IDENTITY, true }; return c; }

} // namespace lambda

using lambda::_1;  // (lambda.hpp: `using lambda::_1`)
using lambda::_2;

// unary function tags usable with transform_if / transform_reduce (functional/identity.hpp, functional/math.hpp: abs;
// negate<T> of functional/operator.hpp; `_1 * _1` of the lambda layer -> square)
template<class T> struct identity { typedef T result_type; static const int unary_code = BCB_UN_IDENTITY; T operator()(const T &x) const { return x; } };
template<class T> struct negate { typedef T result_type; static const int unary_code = BCB_UN_NEGATE; T operator()(const T &x) const { return static_cast<T>(-x); } };
template<class T> struct abs { typedef T result_type; static const int unary_code = BCB_UN_ABS; T operator()(const T &x) const { return x < T(0) ? static_cast<T>(-x) : x; } };
template<class T> struct square { typedef T result_type; static const int unary_code = BCB_UN_SQUARE; T operator()(const T &x) const { return static_cast<T>(x * x); } };
template<class T> struct equal_to { typedef bool result_type; bool operator()(const T &x, const T &y) const { return x == y; } };

} // namespace compute
} // namespace boost

#endif
