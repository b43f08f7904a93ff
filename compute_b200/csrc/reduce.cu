// reduce.cu -- reduce / accumulate for sm_100a.
//
// Replaces reduce_on_gpu + generic_reduce + inplace_reduce of the reference
// (algorithm/detail/reduce_on_gpu.hpp:169-280, algorithm/reduce.hpp:160-236: three launches, ping/pong buffers,
// scalar loads) by ONE launch: every thread streams 128-bit vectors with four independent loads in flight,
// folds them in registers, warps combine with shuffles, and the last block to finish (atomic ticket) folds the
// per-block partials in a fixed order and writes the result -- to device memory, or straight into a pinned,
// device-mapped host slot so a host-returning call costs one launch + one stream sync.
// HBM-bound: 1 x sizeof(T) bytes per element.
#include "ops.cuh"
#include "tma.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace bcb {

constexpr int kReduceThreads = 256;
constexpr int kReduceUnroll = 4;
constexpr int kMaxReduceBlocks = 148 * 8 * 2;

template <typename T, int OP, int VEC>
__device__ __forceinline__ T fold_vec(T acc, const uint4 &v)
{
    const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
    for (int k = 0; k < VEC; k++) acc = Op<OP, T>::apply(acc, e[k]);
    return acc;
}

// Same-type kernel: 128-bit streaming loads over the 16-byte-aligned body, scalar head/tail.
template <typename T, int OP, int UNROLL, bool CONTIG>
__global__ void __launch_bounds__(kReduceThreads)
reduce_kernel(const T *__restrict__ in, size_t n, T *partials, unsigned *done_counter, T *result)
{
    constexpr int kReduceUnroll = UNROLL;
    constexpr int VEC = 16 / sizeof(T);
    __shared__ T smem[32];
    __shared__ bool is_last;

    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;

    size_t head = ((16 - ((uintptr_t)in & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    const size_t nvec = (n - head) / VEC;
    const size_t tail_start = head + nvec * VEC;
    const uint4 *vin = reinterpret_cast<const uint4 *>(in + head);

    T acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = Op<OP, T>::identity();

    if constexpr (CONTIG) {
        // each CTA streams contiguous UNROLL x 4 KiB chunks (thread t reads vectors t, t+256, ... of the chunk)
        const size_t chunk = (size_t)kReduceThreads * kReduceUnroll;
        const size_t nchunks = nvec / chunk;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const uint4 *p = vin + c * chunk + threadIdx.x;
            uint4 x[kReduceUnroll];
#pragma unroll
            for (int u = 0; u < kReduceUnroll; u++) x[u] = ld_stream_v4(p + u * kReduceThreads);
#pragma unroll
            for (int u = 0; u < kReduceUnroll; u++) acc[u % 4] = fold_vec<T, OP, VEC>(acc[u % 4], x[u]);
        }
        for (size_t v = nchunks * chunk + gid; v < nvec; v += gthreads) acc[0] = fold_vec<T, OP, VEC>(acc[0], ld_stream_v4(vin + v));
    } else {
        size_t v = gid;
        for (; v + (kReduceUnroll - 1) * gthreads < nvec; v += kReduceUnroll * gthreads) {
            uint4 x[kReduceUnroll];
#pragma unroll
            for (int u = 0; u < kReduceUnroll; u++) x[u] = ld_stream_v4(vin + v + u * gthreads);
#pragma unroll
            for (int u = 0; u < kReduceUnroll; u++) acc[u % 4] = fold_vec<T, OP, VEC>(acc[u % 4], x[u]);
        }
        for (; v < nvec; v += gthreads) acc[0] = fold_vec<T, OP, VEC>(acc[0], ld_stream_v4(vin + v));
    }
    // scalar head and tail (< 2 * VEC elements in total)
    if (gid < head) acc[1] = Op<OP, T>::apply(acc[1], in[gid]);
    if (tail_start + gid < n) acc[2] = Op<OP, T>::apply(acc[2], in[tail_start + gid]);

    T a = Op<OP, T>::apply(Op<OP, T>::apply(acc[0], acc[1]), Op<OP, T>::apply(acc[2], acc[3]));
    a = block_reduce<T, OP>(a, smem);

    if (gridDim.x == 1) {
        if (threadIdx.x == 0) *result = a;
        return;
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = a;
        __threadfence();
        const unsigned ticket = atomicAdd(done_counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // last block: fold the partials in a fixed (index) order -> run-to-run deterministic for floats
    T p = Op<OP, T>::identity();
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) p = Op<OP, T>::apply(p, ((volatile T *)partials)[i]);
    p = block_reduce<T, OP>(p, smem);
    if (threadIdx.x == 0) {
        *result = p;
        *done_counter = 0;  // ready for the next call on this stream
    }
}

// Large, 16-byte aligned ranges: the bulk-copy engine streams the input.  Persistent CTAs (2 per SM) own a ring of
// shared-memory stages filled by cp.async.bulk (completion on an mbarrier); the threads only read shared memory
// (128-bit, conflict-free) and fold.  128 KiB of loads are in flight per SM without a single register tied up, which is
// what a read-only stream needs to reach the HBM read ceiling (bench/tma_scatter.cu: bulk loads alone 7.1-7.3 TB/s,
// CUB DeviceReduce 6.9 TB/s, the register-staged kernel above 6.6 TB/s).  Tiles are dealt round-robin (no inter-CTA
// dependency: nothing can deadlock); the ragged tail is folded by the last CTA with plain loads.
template <typename T, int OP, int kReduceStageBytes, int kReduceStages>
__global__ void __launch_bounds__(kReduceThreads, 2)
reduce_tma_kernel(const T *__restrict__ in, size_t n, T *partials, unsigned *done_counter, T *result)
{
    constexpr int VEC = 16 / sizeof(T);
    constexpr int S = kReduceStages;
    constexpr size_t TILE = kReduceStageBytes / sizeof(T);
    constexpr int VPT = kReduceStageBytes / 16 / kReduceThreads;  // vectors per thread and stage
    extern __shared__ __align__(128) unsigned char ring[];
    unsigned long long *full_bar = reinterpret_cast<unsigned long long *>(ring + (size_t)S * kReduceStageBytes);
    __shared__ T smem[32];
    __shared__ bool is_last;
    const unsigned tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const size_t tiles = n / TILE;
    auto issue = [&](int s, size_t tile) {
        if (tile < tiles) {
            mbar_expect_tx(&full_bar[s], (unsigned)kReduceStageBytes);
            tma_load_1d(ring + (size_t)s * kReduceStageBytes, in + tile * TILE, (unsigned)kReduceStageBytes, &full_bar[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
        mbar_init_fence();
        for (int s = 0; s < S; s++) issue(s, (size_t)b + (size_t)s * G);
    }
    __syncthreads();
    T acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = Op<OP, T>::identity();
    unsigned it = 0;
    for (size_t tile = b; tile < tiles; tile += G, ++it) {
        const int s = (int)(it % S);
        mbar_wait(&full_bar[s], (it / S) & 1u);
        const uint4 *v = reinterpret_cast<const uint4 *>(ring + (size_t)s * kReduceStageBytes) + tid;
#pragma unroll
        for (int j = 0; j < VPT; j++) acc[j % 4] = fold_vec<T, OP, VEC>(acc[j % 4], v[j * kReduceThreads]);
        __syncthreads();  // everybody has read the stage: refill it with the tile S rounds ahead
        if (tid == 0) issue(s, tile + (size_t)S * G);
    }
    // ragged tail (< one tile): last CTA, plain loads
    if (b == G - 1) {
        for (size_t i = tiles * TILE + tid; i < n; i += kReduceThreads) acc[0] = Op<OP, T>::apply(acc[0], in[i]);
    }
    T a = Op<OP, T>::apply(Op<OP, T>::apply(acc[0], acc[1]), Op<OP, T>::apply(acc[2], acc[3]));
    a = block_reduce<T, OP>(a, smem);
    if (tid == 0) {
        partials[b] = a;
        __threadfence();
        is_last = (atomicAdd(done_counter, 1u) == G - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    T p = Op<OP, T>::identity();  // fixed (index) order: deterministic for floats
    for (unsigned i = tid; i < G; i += kReduceThreads) p = Op<OP, T>::apply(p, ((volatile T *)partials)[i]);
    p = block_reduce<T, OP>(p, smem);
    if (tid == 0) {
        *result = p;
        *done_counter = 0;
    }
}

// Mixed-type kernel (plus<U> over a T range, test_reduce.cpp:269-277): scalar converting loads.
template <typename A, int OP>
__global__ void __launch_bounds__(kReduceThreads)
reduce_cast_kernel(const void *__restrict__ in, int in_dtype, size_t n, A *partials, unsigned *done_counter, A *result)
{
    __shared__ A smem[32];
    __shared__ bool is_last;
    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    A a = Op<OP, A>::identity();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gthreads)
        a = Op<OP, A>::apply(a, load_as<A>(in, i, in_dtype));
    a = block_reduce<A, OP>(a, smem);
    if (gridDim.x == 1) {
        if (threadIdx.x == 0) *result = a;
        return;
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = a;
        __threadfence();
        is_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    A p = Op<OP, A>::identity();
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) p = Op<OP, A>::apply(p, ((volatile A *)partials)[i]);
    p = block_reduce<A, OP>(p, smem);
    if (threadIdx.x == 0) {
        *result = p;
        *done_counter = 0;
    }
}

// serial_accumulate (algorithm/detail/serial_accumulate.hpp:22-50): one thread, strict left fold,
// result = (A) op<F>((F)result, (F)x) -- kept for non-associative ops and mixed accumulator types.
template <typename F>
__device__ __forceinline__ F apply_runtime_op(int op, F a, F b)
{
    switch (op) {
    case BCB_PLUS: return Op<BCB_PLUS, F>::apply(a, b);
    case BCB_MULTIPLIES: return Op<BCB_MULTIPLIES, F>::apply(a, b);
    case BCB_MIN: return Op<BCB_MIN, F>::apply(a, b);
    case BCB_MAX: return Op<BCB_MAX, F>::apply(a, b);
    case BCB_MINUS: {
        typedef typename wrap_type<F>::type W;
        return (F)((W)a - (W)b);
    }
    case BCB_DIVIDES:
        if constexpr (is_fp<F>::value) return a / b;
        else return b == (F)0 ? (F)0 : (F)(a / b);
    default: break;
    }
    if constexpr (!is_fp<F>::value) {
        if (op == BCB_BIT_AND) return (F)(a & b);
        if (op == BCB_BIT_OR) return (F)(a | b);
        if (op == BCB_BIT_XOR) return (F)(a ^ b);
    }
    return a;
}

template <typename A, typename F>
__global__ void serial_accumulate_kernel(const void *in, int in_dtype, size_t n, int op, A init, A *result)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    A r = init;
    for (size_t i = 0; i < n; i++) r = (A)apply_runtime_op<F>(op, (F)r, load_as<F>(in, i, in_dtype));
    *result = r;
}

static int reduce_grid(size_t n, size_t elem_bytes, int sm_count)
{
    // one block per 256 threads x 4 vectors x 16 B = 16 KiB, capped at 8 resident blocks per SM
    size_t bytes = n * elem_bytes;
    size_t blocks = (bytes + (size_t)kReduceThreads * kReduceUnroll * 16 - 1) / ((size_t)kReduceThreads * kReduceUnroll * 16);
    size_t cap = (size_t)sm_count * 8;
    if (cap > (size_t)kMaxReduceBlocks) cap = kMaxReduceBlocks;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T, int OP>
static int launch_reduce_same(StreamState *st, const void *in, size_t n, void *result_dev)
{
    void *partials;
    BCB_TRY(scratch_reserve(st, (size_t)kMaxReduceBlocks * sizeof(T), &partials));
    unsigned *counter = reinterpret_cast<unsigned *>(st->control + kControlReduceDone);
    LaunchTimer timer(st, BCB_K_REDUCE);
    if ((((uintptr_t)in) & 15) == 0 && n * sizeof(T) >= ((size_t)64 << 20)) {
        // >= 64 MiB, 16-byte aligned: bulk-copy ring, 2 persistent CTAs per SM
        // ring shape: 4 stages of 16 KiB (measured 16-48 KiB x 2-6 stages: 6.63-6.68 TB/s per step, all within 1 %;
        // the kernel itself runs at 6.9-6.95 TB/s = the read ceiling of this part, CUB DeviceReduce: 6.92)
        auto go = [&](auto kernel, size_t smem) -> int {
            static std::atomic<unsigned long long> configured{0};  // bit per device
            const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
            if (!(configured.load(std::memory_order_acquire) & bit) || !bit) {
                BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured.fetch_or(bit, std::memory_order_release);
            }
            kernel<<<st->sm_count * 2, kReduceThreads, smem, st->stream>>>((const T *)in, n, (T *)partials, counter, (T *)result_dev);
            BCB_CUDA_TRY(cudaGetLastError());
            return BCB_SUCCESS;
        };
        return go(reduce_tma_kernel<T, OP, 16 * 1024, 4>, 4 * 16 * 1024 + 64);
    }
    {
        // default (measured best of the variants on B200): contiguous 8 KiB chunks per CTA, 2 vectors in flight per
        // thread, 16 CTAs per SM
        int g = st->sm_count * 16 < kMaxReduceBlocks ? st->sm_count * 16 : kMaxReduceBlocks;
        const size_t chunks = (n * sizeof(T) + (size_t)kReduceThreads * 32 - 1) / ((size_t)kReduceThreads * 32);
        if ((size_t)g > chunks) g = (int)(chunks ? chunks : 1);
        reduce_kernel<T, OP, 2, true><<<g, kReduceThreads, 0, st->stream>>>((const T *)in, n, (T *)partials, counter, (T *)result_dev);
    }
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename A, int OP>
static int launch_reduce_cast(StreamState *st, const void *in, int in_dtype, size_t n, void *result_dev)
{
    void *partials;
    BCB_TRY(scratch_reserve(st, (size_t)kMaxReduceBlocks * sizeof(A), &partials));
    unsigned *counter = reinterpret_cast<unsigned *>(st->control + kControlReduceDone);
    int grid = reduce_grid(n, dtype_size(in_dtype), st->sm_count);
    LaunchTimer timer(st, BCB_K_REDUCE);
    reduce_cast_kernel<A, OP><<<grid, kReduceThreads, 0, st->stream>>>(in, in_dtype, n, (A *)partials, counter, (A *)result_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename A>
static int dispatch_reduce_op(StreamState *st, int in_dtype, int res_dtype, int op, const void *in, size_t n, void *result_dev)
{
    const bool same = (in_dtype == res_dtype);
#define OP_CASE(OPC)                                                                    \
    case OPC:                                                                           \
        return same ? launch_reduce_same<A, OPC>(st, in, n, result_dev)                 \
                    : launch_reduce_cast<A, OPC>(st, in, in_dtype, n, result_dev);
    switch (op) {
        OP_CASE(BCB_PLUS) OP_CASE(BCB_MULTIPLIES) OP_CASE(BCB_MIN) OP_CASE(BCB_MAX)
    default: break;
    }
    if constexpr (!is_fp<A>::value) {
        switch (op) {
            OP_CASE(BCB_BIT_AND) OP_CASE(BCB_BIT_OR) OP_CASE(BCB_BIT_XOR)
        default: break;
        }
    }
#undef OP_CASE
    return BCB_EUNSUPPORTED;
}

static int reduce_to_device(StreamState *st, int in_dtype, int res_dtype, int op, const void *in, size_t n, void *result_dev)
{
    switch (res_dtype) {
#define X(DT, T) case DT: return dispatch_reduce_op<T>(st, in_dtype, res_dtype, op, in, n, result_dev);
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
}

template <typename A>
static int launch_serial_acc(StreamState *st, int in_dtype, int op_dtype, int op, const void *in, size_t n, const void *init_host, void *result_dev)
{
    A init;
    std::memcpy(&init, init_host, sizeof(A));
    switch (op_dtype) {
#define X(DT, F)                                                                                                  \
    case DT:                                                                                                      \
        serial_accumulate_kernel<A, F><<<1, 32, 0, st->stream>>>(in, in_dtype, n, op, init, (A *)result_dev);     \
        break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// host-side init (+) r for the associative fast path
template <typename A>
static void host_combine(int op, const void *init_host, const void *r_host, void *out)
{
    A a, b, r;
    std::memcpy(&a, init_host, sizeof(A));
    std::memcpy(&b, r_host, sizeof(A));
    switch (op) {
    case BCB_PLUS: r = Op<BCB_PLUS, A>::apply(a, b); break;
    case BCB_MULTIPLIES: r = Op<BCB_MULTIPLIES, A>::apply(a, b); break;
    case BCB_MIN: r = Op<BCB_MIN, A>::apply(a, b); break;
    case BCB_MAX: r = Op<BCB_MAX, A>::apply(a, b); break;
    default:
        r = a;
        if constexpr (!is_fp<A>::value) {
            if (op == BCB_BIT_AND) r = (A)(a & b);
            if (op == BCB_BIT_OR) r = (A)(a | b);
            if (op == BCB_BIT_XOR) r = (A)(a ^ b);
        }
        break;
    }
    std::memcpy(out, &r, sizeof(A));
}

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_reduce(bcb_stream stream, int in_dtype, int result_dtype, int op, const void *in, size_t n, void *result,
               int result_is_device)
{
    if (n == 0) return BCB_SUCCESS;  // reduce.hpp:283-285: result untouched
    if (!in || !result) return BCB_EINVAL;
    const size_t rw = dtype_size(result_dtype);
    if (!dtype_size(in_dtype) || !rw) return BCB_EINVAL;
    if (!op_is_associative(op)) return BCB_EUNSUPPORTED;
    if (op_is_bitwise(op) && dtype_is_float(result_dtype)) return BCB_EUNSUPPORTED;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    if (result_is_device) return reduce_to_device(st, in_dtype, result_dtype, op, in, n, result);
    BCB_TRY(reduce_to_device(st, in_dtype, result_dtype, op, in, n, st->pinned_slot_dev));
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    std::memcpy(result, st->pinned_slot, rw);
    return BCB_SUCCESS;
}

int bcb_accumulate(bcb_stream stream, int in_dtype, int op_dtype, int acc_dtype, int op, const void *in, size_t n,
                   const void *init_host, void *result_host)
{
    const size_t aw = dtype_size(acc_dtype);
    if (!aw || !dtype_size(op_dtype) || !dtype_size(in_dtype) || !init_host || !result_host) return BCB_EINVAL;
    if (op < BCB_PLUS || op > BCB_DIVIDES) return BCB_EINVAL;
    if (op_is_bitwise(op) && dtype_is_float(op_dtype)) return BCB_EUNSUPPORTED;
    if (n == 0) { std::memcpy(result_host, init_host, aw); return BCB_SUCCESS; }  // accumulate.hpp:109-112
    if (!in) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    if (op_is_associative(op) && acc_dtype == op_dtype) {
        // parallel path: r = x0 op ... op x(n-1) on the device, then init op r on the host
        BCB_TRY(reduce_to_device(st, in_dtype, op_dtype, op, in, n, st->pinned_slot_dev));
        BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
        switch (acc_dtype) {
#define X(DT, T) case DT: host_combine<T>(op, init_host, st->pinned_slot, result_host); break;
            BCB_FOR_EACH_TYPE(X)
#undef X
        default: return BCB_EINVAL;
        }
        return BCB_SUCCESS;
    }
    // serial left fold
    int rc;
    switch (acc_dtype) {
#define X(DT, T) case DT: rc = launch_serial_acc<T>(st, in_dtype, op_dtype, op, in, n, init_host, st->pinned_slot_dev); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_TRY(rc);
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    std::memcpy(result_host, st->pinned_slot, aw);
    return BCB_SUCCESS;
}

}  // extern "C"
