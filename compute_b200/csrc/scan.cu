// scan.cu -- single-pass inclusive / exclusive scan with decoupled look-back for sm_100a.
//
// Replaces scan_impl / local_scan_kernel / write_scanned_output_kernel of the reference
// (algorithm/detail/scan_on_gpu.hpp:26-324: recursive 256-wide Hillis-Steele block scans, two full read+write
// sweeps = 16 B/elem for 4-byte types, ~11 launches + 4 allocations for 2^28 elements, +8 B/elem when in place)
// by ONE kernel that reads every element once and writes it once (8 B/elem, HBM-bound):
//   * tiles of 4096 elements take their id from an atomic ticket (forward progress does not depend on the
//     order the hardware schedules CTAs in);
//   * a tile is loaded with fully coalesced 128-bit loads (lane-striped vectors), scanned in registers with
//     warp shuffles, and its aggregate is published in a tile descriptor;
//   * warp 0 looks back over the 32 preceding descriptors at a time (decoupled look-back) to get the tile's
//     exclusive prefix, publishes the inclusive prefix, and the tile writes its output with 128-bit stores.
// Semantics follow the operator-generic serial_scan (algorithm/detail/serial_scan.hpp:26-97); descriptors carry
// an epoch tag so no per-call initialisation launch is needed.  In-place (in == out) is safe: a tile reads all
// of its input before it writes, and tiles are disjoint.
// Floating-point prefixes are folded strictly in tile order, so results are run-to-run deterministic.
#include "ops.cuh"

#include <cstring>

namespace bcb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanItems = 16;                         // elements per thread
constexpr int kScanTile = kScanThreads * kScanItems;   // 4096
constexpr int kScanMaxWindows = 40;                    // look-back windows buffered for the ordered fp fold

enum : unsigned { kInvalid = 0u, kPartial = 1u, kInclusive = 2u };

template <typename T>
__device__ __forceinline__ T shfl_up_t(T v, int d)
{
    if constexpr (sizeof(T) < 4) return (T)__shfl_up_sync(0xffffffffu, (int)v, d);
    else return __shfl_up_sync(0xffffffffu, v, d);
}
template <typename T>
__device__ __forceinline__ T shfl_t(T v, int src)
{
    if constexpr (sizeof(T) < 4) return (T)__shfl_sync(0xffffffffu, (int)v, src);
    else return __shfl_sync(0xffffffffu, v, src);
}
template <typename T>
__device__ __forceinline__ T shfl_down_t(T v, int d)
{
    if constexpr (sizeof(T) < 4) return (T)__shfl_down_sync(0xffffffffu, (int)v, d);
    else return __shfl_down_sync(0xffffffffu, v, d);
}

// ---- tile descriptors --------------------------------------------------------------------------
// T up to 4 bytes: one 64-bit word {tag = epoch<<2 | status : 32, value bits : 32}, single-copy atomic.
// 8-byte T: status word + separate partial / inclusive value arrays, ordered with release / acquire.
template <typename T, bool SMALL = (sizeof(T) <= 4)> struct TileState;

template <typename T> struct TileState<T, true> {
    unsigned long long *words;
    static size_t bytes(size_t tiles) { return tiles * sizeof(unsigned long long); }
    __host__ __device__ void bind(void *mem, size_t) { words = (unsigned long long *)mem; }
    __device__ __forceinline__ void post(size_t tile, unsigned epoch, unsigned status, T v) const
    {
        unsigned bits = 0;
        memcpy(&bits, &v, sizeof(T));
        st_relaxed_u64(words + tile, ((unsigned long long)((epoch << 2) | status) << 32) | bits);
    }
    // returns status (kInvalid if the slot does not carry this epoch yet)
    __device__ __forceinline__ unsigned peek(size_t tile, unsigned epoch, T &v) const
    {
        const unsigned long long w = ld_relaxed_u64(words + tile);
        const unsigned tag = (unsigned)(w >> 32);
        if ((tag >> 2) != epoch) return kInvalid;
        const unsigned bits = (unsigned)w;
        memcpy(&v, &bits, sizeof(T));
        return tag & 3u;
    }
};

template <typename T> struct TileState<T, false> {
    unsigned *status;
    T *partial;
    T *inclusive;
    static size_t bytes(size_t tiles) { return ((tiles * 4 + 15) & ~(size_t)15) + 2 * tiles * sizeof(T); }
    __host__ __device__ void bind(void *mem, size_t tiles)
    {
        status = (unsigned *)mem;
        partial = (T *)((char *)mem + ((tiles * 4 + 15) & ~(size_t)15));
        inclusive = partial + tiles;
    }
    __device__ __forceinline__ void post(size_t tile, unsigned epoch, unsigned st, T v) const
    {
        T *dst = (st == kPartial) ? partial : inclusive;
        *((volatile T *)(dst + tile)) = v;
        st_release_u32(status + tile, (epoch << 2) | st);
    }
    __device__ __forceinline__ unsigned peek(size_t tile, unsigned epoch, T &v) const
    {
        const unsigned tag = ld_acquire_u32(status + tile);
        if ((tag >> 2) != epoch) return kInvalid;
        const unsigned st = tag & 3u;
        const T *src = (st == kPartial) ? partial : inclusive;
        v = *((volatile const T *)(src + tile));
        return st;
    }
};

// Exclusive prefix of `tile` (> 0), computed by warp 0; result valid in every lane.
template <typename T, int OP>
__device__ __forceinline__ T lookback_prefix(const TileState<T> &ts, size_t tile, unsigned epoch, T (*window_buf)[32])
{
    typedef Op<OP, T> O;
    const unsigned lane = lane_id();
    long long base = (long long)tile - 1;
    T running = O::identity();  // integers: fold of the windows seen so far (order irrelevant)
    int nwin = 0;               // fp: number of all-partial windows buffered
    while (true) {
        const long long idx = base - (long long)lane;
        T val = O::identity();
        unsigned st = kInclusive;  // tiles "before 0" behave as an inclusive identity
        if (idx >= 0) {
            do { st = ts.peek((size_t)idx, epoch, val); } while (st == kInvalid);
        }
        const unsigned inc = __ballot_sync(0xffffffffu, st == kInclusive);
        if constexpr (!is_fp<T>::value) {
            const int first = inc ? (__ffs(inc) - 1) : 31;
            T v = ((int)lane <= first) ? val : O::identity();
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v = O::apply(v, shfl_down_t(v, off));
            running = O::apply(shfl_t(v, 0), running);
            if (inc) return running;
        } else {
            if (inc) {
                // ordered fold, oldest tile first: inclusive(first), partial(first-1) ... partial(0),
                // then the buffered windows from the most recently buffered (older tiles) to the first one.
                const int first = __ffs(inc) - 1;
                T acc = shfl_t(val, first);
                for (int l = first - 1; l >= 0; --l) acc = O::apply(acc, shfl_t(val, l));
                for (int w = nwin - 1; w >= 0; --w) {
                    const T wv = window_buf[w][lane];
                    for (int l = 31; l >= 0; --l) acc = O::apply(acc, shfl_t(wv, l));
                }
                return O::apply(acc, running);  // running is the identity unless the buffer overflowed
            }
            if (nwin < kScanMaxWindows) {
                window_buf[nwin][lane] = val;
                ++nwin;
            } else {  // > 1280 unresolved predecessors: keep going unordered (still within tolerance)
                T v = val;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v = O::apply(v, shfl_down_t(v, off));
                running = O::apply(shfl_t(v, 0), running);
            }
        }
        base -= 32;
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const T *in, T *out, size_t n, int exclusive, T init, TileState<T> ts, unsigned epoch,
            unsigned long long *ticket, unsigned long long ticket_base)
{
    typedef Op<OP, T> O;
    constexpr int VEC = 16 / sizeof(T);        // elements per 128-bit vector
    constexpr int NV = kScanItems / VEC;       // vectors per thread
    static_assert(NV >= 1, "vector wider than the per-thread item count");

    __shared__ unsigned long long s_tile;
    __shared__ T s_warp_total[kScanWarps];
    __shared__ T s_tile_prefix;
    __shared__ T s_window[is_fp<T>::value ? kScanMaxWindows : 1][32];

    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1ull) - ticket_base;
    __syncthreads();
    const size_t tile = (size_t)s_tile;
    const size_t tile_base = tile * (size_t)kScanTile;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const size_t warp_base = tile_base + (size_t)warp * (32 * kScanItems);
    const bool full = tile_base + kScanTile <= n;

    // ---- load: vector j of lane l covers elements warp_base + (j*32 + l)*VEC .. +VEC ----
    T x[NV][VEC];
    const bool vec_in = full && (((uintptr_t)in & 15) == 0);
    if (vec_in) {
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(in + warp_base + (size_t)(j * 32 + lane) * VEC);
            const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
            for (int k = 0; k < VEC; k++) x[j][k] = e[k];
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; j++) {
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const size_t i = warp_base + (size_t)(j * 32 + lane) * VEC + k;
                x[j][k] = i < n ? in[i] : O::identity();
            }
        }
    }

    // ---- thread-local inclusive scan inside each vector; vsum[j] = vector total ----
    T vsum[NV];
#pragma unroll
    for (int j = 0; j < NV; j++) {
#pragma unroll
        for (int k = 1; k < VEC; k++) x[j][k] = O::apply(x[j][k - 1], x[j][k]);
        vsum[j] = x[j][VEC - 1];
    }
    // ---- warp scan of the vector totals (one independent 5-step scan per vector index) ----
    T vexcl[NV];  // exclusive prefix of this lane's vector j inside the warp's segment
    T carry = O::identity();
#pragma unroll
    for (int j = 0; j < NV; j++) {
        T s = vsum[j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T o = shfl_up_t(s, d);
            if ((int)lane >= d) s = O::apply(o, s);
        }
        T e = shfl_up_t(s, 1);
        if (lane == 0) e = O::identity();
        vexcl[j] = O::apply(carry, e);
        carry = O::apply(carry, shfl_t(s, 31));
    }
    // carry = this warp's total
    if (lane == 0) s_warp_total[warp] = carry;
    __syncthreads();

    T warp_off = O::identity();
    T aggregate = O::identity();
#pragma unroll
    for (int w = 0; w < kScanWarps; w++) {
        const T t = s_warp_total[w];
        if (w < (int)warp) warp_off = O::apply(warp_off, t);
        aggregate = O::apply(aggregate, t);
    }

    // ---- tile descriptor + decoupled look-back (warp 0) ----
    if (warp == 0) {
        T prefix;
        if (tile == 0) {
            prefix = exclusive ? init : O::identity();  // modes 1 (exclusive) and 2 (seeded inclusive) start from init
            if (lane == 0) ts.post(0, epoch, kInclusive, exclusive ? O::apply(init, aggregate) : aggregate);
        } else {
            if (lane == 0) ts.post(tile, epoch, kPartial, aggregate);
            prefix = lookback_prefix<T, OP>(ts, tile, epoch, s_window);
            if (lane == 0) ts.post(tile, epoch, kInclusive, O::apply(prefix, aggregate));
        }
        if (lane == 0) s_tile_prefix = prefix;
    }
    __syncthreads();
    const T base = O::apply(s_tile_prefix, warp_off);
    // tile 0 of an inclusive scan has no prefix at all: `base` is then the identity, which is exact

    // ---- outputs ----
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const T p = O::apply(base, vexcl[j]);
        T y[VEC];
        if (exclusive == 1) {
            y[0] = p;
#pragma unroll
            for (int k = 1; k < VEC; k++) y[k] = O::apply(p, x[j][k - 1]);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; k++) y[k] = O::apply(p, x[j][k]);
        }
        const size_t i0 = warp_base + (size_t)(j * 32 + lane) * VEC;
        if (full && (((uintptr_t)out & 15) == 0)) {
            *reinterpret_cast<uint4 *>(out + i0) = *reinterpret_cast<const uint4 *>(y);
        } else {
#pragma unroll
            for (int k = 0; k < VEC; k++)
                if (i0 + k < n) out[i0 + k] = y[k];
        }
    }
}

template <typename A>
__global__ void convert_kernel(const void *in, int in_dtype, A *out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = load_as<A>(in, i, in_dtype);
}

template <typename T, int OP>
static int launch_scan(StreamState *st, const void *in, void *out, size_t n, int exclusive, const void *init_host)
{
    T init = (T)0;
    if (init_host) std::memcpy(&init, init_host, sizeof(T));
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles > 0x7fffffffull) return BCB_ETOOLARGE;
    void *mem;
    BCB_TRY(lookback_reserve(st, TileState<T>::bytes(tiles), &mem));
    unsigned epoch;
    BCB_TRY(next_epoch(st, &epoch));
    TileState<T> ts;
    ts.bind(mem, tiles);
    const unsigned long long base = st->ticket_base;
    st->ticket_base += tiles;
    LaunchTimer timer(st, BCB_K_SCAN);
    scan_kernel<T, OP><<<(unsigned)tiles, kScanThreads, 0, st->stream>>>(
        (const T *)in, (T *)out, n, exclusive, init, ts, epoch, st->control + kControlTicket, base);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename T>
static int dispatch_scan_op(StreamState *st, int op, const void *in, void *out, size_t n, int exclusive, const void *init_host)
{
    switch (op) {
    case BCB_PLUS: return launch_scan<T, BCB_PLUS>(st, in, out, n, exclusive, init_host);
    case BCB_MULTIPLIES: return launch_scan<T, BCB_MULTIPLIES>(st, in, out, n, exclusive, init_host);
    case BCB_MIN: return launch_scan<T, BCB_MIN>(st, in, out, n, exclusive, init_host);
    case BCB_MAX: return launch_scan<T, BCB_MAX>(st, in, out, n, exclusive, init_host);
    default: break;
    }
    if constexpr (!is_fp<T>::value) {
        switch (op) {
        case BCB_BIT_AND: return launch_scan<T, BCB_BIT_AND>(st, in, out, n, exclusive, init_host);
        case BCB_BIT_OR: return launch_scan<T, BCB_BIT_OR>(st, in, out, n, exclusive, init_host);
        case BCB_BIT_XOR: return launch_scan<T, BCB_BIT_XOR>(st, in, out, n, exclusive, init_host);
        default: break;
        }
    }
    return BCB_EUNSUPPORTED;
}

}  // namespace bcb

using namespace bcb;

extern "C" int bcb_scan(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in, void *out,
                        size_t n, const void *init_host)
{
    if (n == 0) return BCB_SUCCESS;  // scan_on_gpu.hpp:316-318
    if (!in || !out) return BCB_EINVAL;
    const size_t ow = dtype_size(out_dtype);
    if (!ow || !dtype_size(in_dtype)) return BCB_EINVAL;
    if (!op_is_associative(op)) return BCB_EUNSUPPORTED;
    if (op_is_bitwise(op) && dtype_is_float(out_dtype)) return BCB_EUNSUPPORTED;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    const void *src = in;
    if (in_dtype != out_dtype) {
        // arithmetic happens in the OUTPUT type (exclusive_scan.hpp:80-85): convert first, then scan
        void *tmp;
        BCB_TRY(scratch_reserve(st, n * ow, &tmp));
        size_t blocks = (n + 255) / 256;
        const size_t cap = (size_t)st->sm_count * 16;
        if (blocks > cap) blocks = cap;
        switch (out_dtype) {
#define X(DT, T) case DT: convert_kernel<T><<<(unsigned)blocks, 256, 0, st->stream>>>(in, in_dtype, (T *)tmp, n); break;
            BCB_FOR_EACH_TYPE(X)
#undef X
        default: return BCB_EINVAL;
        }
        BCB_CUDA_TRY(cudaGetLastError());
        src = tmp;
    }
    switch (out_dtype) {
#define X(DT, T) case DT: return dispatch_scan_op<T>(st, op, src, out, n, exclusive, init_host);
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
}
