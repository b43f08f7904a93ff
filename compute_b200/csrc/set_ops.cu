// set_ops.cu -- callers of scan and reduce, second batch for sm_100a:
//   * set_union / set_intersection / set_difference / set_symmetric_difference on two sorted ranges
//     (algorithm/set_union.hpp:120-199, set_intersection.hpp:104-175, set_difference.hpp:112-186,
//     set_symmetric_difference.hpp:121-199: balanced-path tiles -> flags -> exclusive_scan -> scatter).  Here: one pass
//     decides for every element of A and of B whether it survives (multiset rule: the k-th occurrence of v in A
//     survives an intersection iff k < count_B(v), a difference iff k >= count_B(v); B's k-th occurrence survives a
//     union / symmetric difference iff k >= count_A(v)), by binary search in the other range; the two flag arrays go
//     through the library's single-pass scan; a scatter pass writes every survivor at
//         kept_A_before(i) + kept_B_before(lower_bound_B(v))        for A[i]
//         kept_B_before(j) + kept_A_before(upper_bound_A(v))        for B[j]
//     which is the std::set_* order (equal elements: first range first).
//   * min_element / max_element (algorithm/detail/find_extrema_with_reduce.hpp:77-316): one-launch (value, index)
//     reduction; ties go to the smaller index, for the minimum and for the maximum alike (:156-158).
#include "ops.cuh"

#include <cstring>

namespace bcb {

int scratch_reserve(StreamState *st, size_t bytes, void **out);

template <typename T>
__device__ __forceinline__ unsigned lower_bound_dev(const T *a, unsigned n, T v)
{
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
template <typename T>
__device__ __forceinline__ unsigned upper_bound_dev(const T *a, unsigned n, T v)
{
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        if (v < a[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

// flags + the position each element needs in the OTHER range's kept-prefix array
template <typename T>
__global__ void __launch_bounds__(256)
set_flags_kernel(const T *__restrict__ a, unsigned na, const T *__restrict__ b, unsigned nb, int which, unsigned *__restrict__ flag_a,
                 unsigned *__restrict__ other_a, unsigned *__restrict__ flag_b, unsigned *__restrict__ other_b)
{
    const size_t total = (size_t)na + nb + 2;  // one extra slot per range: its scan then also yields the range's total
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (i <= na) {
            if (i == na) { flag_a[na] = 0; continue; }
            const unsigned ia = (unsigned)i;
            const T v = a[ia];
            const unsigned lb_b = lower_bound_dev(b, nb, v);
            unsigned keep = 1;  // union: all of A
            if (which != BCB_SET_UNION) {
                const unsigned k = ia - lower_bound_dev(a, na, v);
                const unsigned cnt_b = upper_bound_dev(b, nb, v) - lb_b;
                keep = which == BCB_SET_INTERSECTION ? (k < cnt_b) : (k >= cnt_b);
            }
            flag_a[ia] = keep;
            other_a[ia] = lb_b;
        } else {
            const size_t j = i - na - 1;
            if (j == nb) { flag_b[nb] = 0; continue; }
            const unsigned jb = (unsigned)j;
            unsigned keep = 0;  // intersection, difference: nothing of B
            const T v = b[jb];
            const unsigned ub_a = upper_bound_dev(a, na, v);
            if (which == BCB_SET_UNION || which == BCB_SET_SYMMETRIC_DIFFERENCE) {
                const unsigned k = jb - lower_bound_dev(b, nb, v);
                const unsigned cnt_a = ub_a - lower_bound_dev(a, na, v);
                keep = k >= cnt_a;
            }
            flag_b[jb] = keep;
            other_b[jb] = ub_a;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
set_scatter_kernel(const T *__restrict__ a, unsigned na, const T *__restrict__ b, unsigned nb, const unsigned *__restrict__ kept_a,
                   const unsigned *__restrict__ other_a, const unsigned *__restrict__ kept_b, const unsigned *__restrict__ other_b,
                   T *__restrict__ out, unsigned long long *total)
{
    // kept_x[i] = number of survivors of range x before position i (exclusive scan of the flags, one slot past the end)
    const size_t n = (size_t)na + nb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (i < na) {
            const unsigned ia = (unsigned)i;
            if (kept_a[ia + 1] != kept_a[ia]) out[(size_t)kept_a[ia] + kept_b[other_a[ia]]] = a[ia];
        } else {
            const unsigned jb = (unsigned)(i - na);
            if (kept_b[jb + 1] != kept_b[jb]) out[(size_t)kept_b[jb] + kept_a[other_b[jb]]] = b[jb];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = (unsigned long long)kept_a[na] + kept_b[nb];
}

template <typename T>
static int set_operation_typed(StreamState *st, int which, const void *a, size_t na, const void *b, size_t nb, void *out)
{
    // scratch: flags / kept prefix (in place) and cross positions of both ranges
    const size_t la = na + 1, lb = nb + 1;
    void *mem;
    BCB_TRY(scratch_reserve(st, (2 * la + 2 * lb) * sizeof(unsigned) + 64, &mem));
    unsigned *flag_a = (unsigned *)mem, *other_a = flag_a + la, *flag_b = other_a + la, *other_b = flag_b + lb;
    const size_t total = na + nb + 2;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)st->sm_count * 16;
    if (blocks > cap) blocks = cap;
    set_flags_kernel<T><<<(unsigned)blocks, 256, 0, st->stream>>>((const T *)a, (unsigned)na, (const T *)b, (unsigned)nb, which, flag_a, other_a, flag_b,
                                                                 other_b);
    BCB_CUDA_TRY(cudaGetLastError());
    const unsigned zero = 0;
    BCB_TRY(bcb_scan((bcb_stream)st->stream, BCB_UINT, BCB_UINT, BCB_PLUS, 1, flag_a, flag_a, la, &zero));
    BCB_TRY(bcb_scan((bcb_stream)st->stream, BCB_UINT, BCB_UINT, BCB_PLUS, 1, flag_b, flag_b, lb, &zero));
    set_scatter_kernel<T><<<(unsigned)blocks, 256, 0, st->stream>>>((const T *)a, (unsigned)na, (const T *)b, (unsigned)nb, flag_a, other_a, flag_b,
                                                                   other_b, (T *)out, (unsigned long long *)st->pinned_slot_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// ---- extrema -------------------------------------------------------------------------------------------------
template <typename T> struct Extremum { T v; unsigned long long i; };

template <typename T, bool MAX>
__device__ __forceinline__ Extremum<T> better(Extremum<T> x, Extremum<T> y)
{
    // the candidate that compares strictly better wins; equal candidates: the smaller index
    const bool y_wins = MAX ? (x.v < y.v) : (y.v < x.v);
    const bool x_wins = MAX ? (y.v < x.v) : (x.v < y.v);
    if (y_wins || (!x_wins && y.i < x.i)) return y;
    return x;
}

template <typename T, bool MAX>
__device__ __forceinline__ Extremum<T> warp_best(Extremum<T> e)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Extremum<T> o;
        if constexpr (sizeof(T) < 4) o.v = (T)__shfl_down_sync(0xffffffffu, (int)e.v, off);
        else o.v = __shfl_down_sync(0xffffffffu, e.v, off);
        o.i = __shfl_down_sync(0xffffffffu, e.i, off);
        e = better<T, MAX>(e, o);
    }
    return e;
}

constexpr int kExtThreads = 256;

template <typename T, bool MAX>
__global__ void __launch_bounds__(kExtThreads)
find_extremum_kernel(const T *__restrict__ in, size_t n, T *part_v, unsigned long long *part_i, unsigned *done_counter, unsigned long long *result)
{
    __shared__ T sv[kExtThreads / 32];
    __shared__ unsigned long long si[kExtThreads / 32];
    __shared__ bool is_last;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    auto block_best = [&](Extremum<T> e) {
        e = warp_best<T, MAX>(e);
        if (lane == 0) { sv[warp] = e.v; si[warp] = e.i; }
        __syncthreads();
        if (warp == 0) {
            Extremum<T> w;
            w.v = sv[lane < kExtThreads / 32 ? lane : 0];
            w.i = si[lane < kExtThreads / 32 ? lane : 0];
            e = warp_best<T, MAX>(w);
        }
        __syncthreads();
        return e;  // valid in thread 0
    };
    // every thread starts from element 0 (always a valid candidate: n > 0)
    Extremum<T> e{in[0], 0ull};
    for (size_t i = (size_t)blockIdx.x * kExtThreads + tid; i < n; i += (size_t)gridDim.x * kExtThreads)
        e = better<T, MAX>(e, Extremum<T>{in[i], (unsigned long long)i});
    e = block_best(e);
    if (tid == 0) {
        part_v[blockIdx.x] = e.v;
        part_i[blockIdx.x] = e.i;
        __threadfence();
        is_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Extremum<T> f{((volatile T *)part_v)[0], ((volatile unsigned long long *)part_i)[0]};
    for (unsigned b = tid; b < gridDim.x; b += kExtThreads)
        f = better<T, MAX>(f, Extremum<T>{((volatile T *)part_v)[b], ((volatile unsigned long long *)part_i)[b]});
    f = block_best(f);
    if (tid == 0) {
        *result = f.i;
        *done_counter = 0;
    }
}

template <typename T>
static int find_extremum_typed(StreamState *st, const void *in, size_t n, int want_max)
{
    size_t blocks = (n + kExtThreads * 8 - 1) / (kExtThreads * 8);
    const size_t cap = (size_t)st->sm_count * 8;
    if (blocks > cap) blocks = cap;
    void *mem;
    BCB_TRY(scratch_reserve(st, blocks * 16 + 64, &mem));
    unsigned long long *part_i = (unsigned long long *)mem;
    T *part_v = (T *)(part_i + blocks);
    unsigned *done = (unsigned *)(st->control + kControlReduceDone);
    unsigned long long *result = (unsigned long long *)st->pinned_slot_dev;
    if (want_max) find_extremum_kernel<T, true><<<(unsigned)blocks, kExtThreads, 0, st->stream>>>((const T *)in, n, part_v, part_i, done, result);
    else find_extremum_kernel<T, false><<<(unsigned)blocks, kExtThreads, 0, st->stream>>>((const T *)in, n, part_v, part_i, done, result);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_set_operation(bcb_stream stream, int dtype, int which, const void *a, size_t na, const void *b, size_t nb, void *out, size_t *count_host)
{
    if (!count_host) return BCB_EINVAL;
    *count_host = 0;
    if (!dtype_size(dtype)) return BCB_EINVAL;
    if (which < BCB_SET_UNION || which > BCB_SET_SYMMETRIC_DIFFERENCE) return BCB_EINVAL;
    if (na + nb == 0) return BCB_SUCCESS;
    if ((na && !a) || (nb && !b) || !out) return BCB_EINVAL;
    if (na >= 0xfffffff0ull || nb >= 0xfffffff0ull || na + nb >= 0xfffffff0ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    int rc;
    switch (dtype) {
#define X(DT, T) case DT: rc = set_operation_typed<T>(st, which, a, na, b, nb, out); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_TRY(rc);
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));  // the returned end iterator is a host value
    *count_host = (size_t)(*(volatile unsigned long long *)st->pinned_slot);
    return BCB_SUCCESS;
}

int bcb_find_extremum(bcb_stream stream, int dtype, const void *in, size_t n, int want_max, size_t *index_host)
{
    if (!index_host) return BCB_EINVAL;
    *index_host = 0;
    if (!dtype_size(dtype)) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;  // empty or one element: first (test_extrema.cpp:39-51)
    if (!in) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    int rc;
    switch (dtype) {
#define X(DT, T) case DT: rc = find_extremum_typed<T>(st, in, n, want_max); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_TRY(rc);
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));  // the returned iterator is a host value
    *index_host = (size_t)(*(volatile unsigned long long *)st->pinned_slot);
    return BCB_SUCCESS;
}

}  // extern "C"
