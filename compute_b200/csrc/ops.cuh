// ops.cuh -- the binary functors the path accepts (functional/operator.hpp:73-96 of the reference) as
// compile-time tags, plus dtype <-> C++ type dispatch helpers.
#pragma once

#include "common.cuh"
#include <cfloat>
#include <climits>
#include <cmath>

namespace bcb {

template <typename T> struct is_fp { static constexpr bool value = false; };
template <> struct is_fp<float> { static constexpr bool value = true; };
template <> struct is_fp<double> { static constexpr bool value = true; };

template <typename T> struct limits;
template <> struct limits<signed char> { static __host__ __device__ signed char lo() { return SCHAR_MIN; } static __host__ __device__ signed char hi() { return SCHAR_MAX; } };
template <> struct limits<unsigned char> { static __host__ __device__ unsigned char lo() { return 0; } static __host__ __device__ unsigned char hi() { return UCHAR_MAX; } };
template <> struct limits<short> { static __host__ __device__ short lo() { return SHRT_MIN; } static __host__ __device__ short hi() { return SHRT_MAX; } };
template <> struct limits<unsigned short> { static __host__ __device__ unsigned short lo() { return 0; } static __host__ __device__ unsigned short hi() { return USHRT_MAX; } };
template <> struct limits<int> { static __host__ __device__ int lo() { return INT_MIN; } static __host__ __device__ int hi() { return INT_MAX; } };
template <> struct limits<unsigned> { static __host__ __device__ unsigned lo() { return 0; } static __host__ __device__ unsigned hi() { return UINT_MAX; } };
template <> struct limits<long long> { static __host__ __device__ long long lo() { return LLONG_MIN; } static __host__ __device__ long long hi() { return LLONG_MAX; } };
template <> struct limits<unsigned long long> { static __host__ __device__ unsigned long long lo() { return 0; } static __host__ __device__ unsigned long long hi() { return ULLONG_MAX; } };
// floating point: the identities of min / max are +-infinity
__host__ __device__ inline float inf_f()
{
#ifdef __CUDA_ARCH__
    return __int_as_float(0x7f800000);
#else
    return HUGE_VALF;
#endif
}
__host__ __device__ inline double inf_d()
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff0000000000000LL);
#else
    return HUGE_VAL;
#endif
}
template <> struct limits<float> { static __host__ __device__ float lo() { return -inf_f(); } static __host__ __device__ float hi() { return inf_f(); } };
template <> struct limits<double> { static __host__ __device__ double lo() { return -inf_d(); } static __host__ __device__ double hi() { return inf_d(); } };

// unsigned type of the same width, for wrap-around integer arithmetic
template <typename T> struct wrap_type { typedef T type; };
template <> struct wrap_type<signed char> { typedef unsigned type; };
template <> struct wrap_type<unsigned char> { typedef unsigned type; };
template <> struct wrap_type<short> { typedef unsigned type; };
template <> struct wrap_type<unsigned short> { typedef unsigned type; };
template <> struct wrap_type<int> { typedef unsigned type; };
template <> struct wrap_type<long long> { typedef unsigned long long type; };

template <int OP, typename T> struct Op;

template <typename T> struct Op<BCB_PLUS, T> {
    typedef typename wrap_type<T>::type W;
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return (T)((W)a + (W)b); }
    // -0.0 is the exact identity of floating-point addition (+0.0 would turn a -0.0 input into +0.0)
    static __host__ __device__ __forceinline__ T identity() { return is_fp<T>::value ? (T)(-0.0) : (T)0; }
};
template <typename T> struct Op<BCB_MULTIPLIES, T> {
    typedef typename wrap_type<T>::type W;
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return (T)((W)a * (W)b); }
    static __host__ __device__ __forceinline__ T identity() { return (T)1; }
};
// OpenCL C: min(x, y) = y < x ? y : x ; max(x, y) = x < y ? y : x
template <typename T> struct Op<BCB_MIN, T> {
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return b < a ? b : a; }
    static __host__ __device__ __forceinline__ T identity() { return limits<T>::hi(); }
};
template <typename T> struct Op<BCB_MAX, T> {
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return a < b ? b : a; }
    static __host__ __device__ __forceinline__ T identity() { return limits<T>::lo(); }
};
template <typename T> struct Op<BCB_BIT_AND, T> {
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return (T)(a & b); }
    static __host__ __device__ __forceinline__ T identity() { return (T)~(T)0; }
};
template <typename T> struct Op<BCB_BIT_OR, T> {
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return (T)(a | b); }
    static __host__ __device__ __forceinline__ T identity() { return (T)0; }
};
template <typename T> struct Op<BCB_BIT_XOR, T> {
    static __host__ __device__ __forceinline__ T apply(T a, T b) { return (T)(a ^ b); }
    static __host__ __device__ __forceinline__ T identity() { return (T)0; }
};

inline bool op_is_associative(int op) { return op >= BCB_PLUS && op <= BCB_BIT_XOR; }
inline bool op_is_bitwise(int op) { return op == BCB_BIT_AND || op == BCB_BIT_OR || op == BCB_BIT_XOR; }

// element i of a buffer of runtime dtype, converted to A (C++ conversion, like the implicit
// conversion OpenCL C applies when a plus<U> functor is fed T values)
template <typename A>
__host__ __device__ __forceinline__ A load_as(const void *p, size_t i, int dtype)
{
    switch (dtype) {
    case BCB_CHAR: return (A)((const signed char *)p)[i];
    case BCB_UCHAR: return (A)((const unsigned char *)p)[i];
    case BCB_SHORT: return (A)((const short *)p)[i];
    case BCB_USHORT: return (A)((const unsigned short *)p)[i];
    case BCB_INT: return (A)((const int *)p)[i];
    case BCB_UINT: return (A)((const unsigned *)p)[i];
    case BCB_LONG: return (A)((const long long *)p)[i];
    case BCB_ULONG: return (A)((const unsigned long long *)p)[i];
    case BCB_FLOAT: return (A)((const float *)p)[i];
    default: return (A)((const double *)p)[i];
    }
}

// dtype -> type dispatch: X(DTYPE_CODE, CPP_TYPE)
#define BCB_FOR_EACH_INT_TYPE(X)                                                             \
    X(BCB_CHAR, signed char) X(BCB_UCHAR, unsigned char) X(BCB_SHORT, short) X(BCB_USHORT, unsigned short) \
    X(BCB_INT, int) X(BCB_UINT, unsigned) X(BCB_LONG, long long) X(BCB_ULONG, unsigned long long)
#define BCB_FOR_EACH_FP_TYPE(X) X(BCB_FLOAT, float) X(BCB_DOUBLE, double)
#define BCB_FOR_EACH_TYPE(X) BCB_FOR_EACH_INT_TYPE(X) BCB_FOR_EACH_FP_TYPE(X)

#ifdef __CUDACC__
// fixed-shape (hence run-to-run deterministic) warp / block folds shared by reduce.cu and stream_ops.cu
template <typename A, int OP>
__device__ __forceinline__ A warp_reduce(A v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        A o;
        if constexpr (sizeof(A) < 4) o = (A)__shfl_down_sync(0xffffffffu, (int)v, off);
        else o = __shfl_down_sync(0xffffffffu, v, off);
        v = Op<OP, A>::apply(v, o);
    }
    return v;
}

template <typename A, int OP>
__device__ __forceinline__ A block_reduce(A v, A *smem /* [32] */)
{
    v = warp_reduce<A, OP>(v);
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const unsigned nwarps = blockDim.x >> 5;
        A w = lane < nwarps ? smem[lane] : Op<OP, A>::identity();
        w = warp_reduce<A, OP>(w);
        if (lane == 0) smem[0] = w;
    }
    __syncthreads();
    A r = smem[0];
    __syncthreads();
    return r;
}

#endif  // __CUDACC__

}  // namespace bcb
