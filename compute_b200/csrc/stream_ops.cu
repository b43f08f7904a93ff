// stream_ops.cu -- the callers of scan and reduce as fused single-pass kernels:
//   transform_if / copy_if   algorithm/transform_if.hpp:42-85 (flags -> exclusive_scan -> scatter: three sweeps and an
//                            n-element index vector in the reference)            -> ONE kernel, 2 x sizeof(T) B/elem at most
//   count_if / count         algorithm/detail/count_if_with_reduce.hpp:27-80 (transform to 0/1 + reduce in ulong)
//   transform_reduce,        algorithm/transform_reduce.hpp:40-90, algorithm/inner_product.hpp:40-97 (transform
//   inner_product            iterators feeding reduce / accumulate)                -> one fused load-transform-reduce kernel
//   reduce_by_key            algorithm/detail/reduce_by_key_with_scan.hpp:48-97 (head flags, inclusive_scan_by_key over
//                            (flag, value) pairs, scatter of the segment ends)    -> ONE kernel with a segmented look-back
// The reference builds these from run-time generated OpenCL C for arbitrary functors; here the functor set is closed
// (include/compute_b200.h: bcb_pred -- ((x ARITH a) CMP b) --, bcb_unary, the bcb_op codes) and compiled ahead of time.
// Order is preserved everywhere (copy_if is stable, reduce_by_key emits segments in input order); integer results are
// exact, floating-point folds have a fixed shape and are run-to-run deterministic.
#include "tile_state.cuh"

#include <cstring>

namespace bcb {

// ---- the closed functor set -----------------------------------------------------------------------------------
struct Pred {
    int arith, cmp;
    unsigned long long a_bits, b_bits;  // operands in the element type's bit pattern
};

template <typename T> __device__ __forceinline__ T from_bits(unsigned long long bits)
{
    T v;
    memcpy(&v, &bits, sizeof(T));
    return v;
}

// ((x ARITH a) CMP b), evaluated like the OpenCL C expression the reference would generate: narrow integers promote to int
template <typename T>
__device__ __forceinline__ bool eval_pred(T x, int arith, int cmp, T a, T b)
{
    typedef decltype(T() * T()) P;
    P y = (P)x;
    switch (arith) {
    case BCB_AR_MUL: y = (P)((P)x * (P)a); break;
    case BCB_AR_ADD: y = (P)((P)x + (P)a); break;
    case BCB_AR_SUB: y = (P)((P)x - (P)a); break;
    case BCB_AR_MOD:
        if constexpr (!is_fp<T>::value) y = a == (T)0 ? (P)0 : (P)((P)x % (P)a);
        break;
    case BCB_AR_AND:
        if constexpr (!is_fp<T>::value) y = (P)((P)x & (P)a);
        break;
    default: break;
    }
    const P c = (P)b;
    switch (cmp) {
    case BCB_CMP_EQ: return y == c;
    case BCB_CMP_NE: return y != c;
    case BCB_CMP_LT: return y < c;
    case BCB_CMP_LE: return y <= c;
    case BCB_CMP_GT: return y > c;
    case BCB_CMP_GE: return y >= c;
    default: return true;
    }
}

template <typename T>
__device__ __forceinline__ T apply_unary(int code, T x)
{
    typedef typename wrap_type<T>::type W;
    switch (code) {
    case BCB_UN_NEGATE:
        if constexpr (is_fp<T>::value) return -x;  // sign flip: -(0.0) is -0.0
        else return (T)((W)0 - (W)x);
    case BCB_UN_ABS:
        if constexpr (is_fp<T>::value) return x < (T)0 ? -x : (x == (T)0 ? (T)0 : x);  // fabs (also clears -0.0)
        else return x < (T)0 ? (T)((W)0 - (W)x) : x;
    case BCB_UN_SQUARE: return (T)((W)x * (W)x);
    default: return x;
    }
}

template <typename T>
__device__ __forceinline__ T apply_binary(int op, T a, T b)
{
    typedef typename wrap_type<T>::type W;
    switch (op) {
    case BCB_PLUS: return (T)((W)a + (W)b);
    case BCB_MINUS: return (T)((W)a - (W)b);
    case BCB_MULTIPLIES: return (T)((W)a * (W)b);
    case BCB_MIN: return b < a ? b : a;
    case BCB_MAX: return a < b ? b : a;
    default: break;
    }
    if constexpr (!is_fp<T>::value) {
        if (op == BCB_BIT_AND) return (T)(a & b);
        if (op == BCB_BIT_OR) return (T)(a | b);
        if (op == BCB_BIT_XOR) return (T)(a ^ b);
    }
    return a;
}

// ---- transform_if / copy_if: one pass, stable --------------------------------------------------------------------
// A tile is 8 warps x 16 x 32 consecutive elements; warp w owns one contiguous chunk and reads it 32 elements (one
// 128-byte line for 4-byte types) at a time, so selected elements are ranked by ballot + popc in input order.  The
// tile's selected count goes through the same epoch-tagged descriptors and warp-parallel look-back as the scan.
constexpr int kCiThreads = 256, kCiItems = 16, kCiTile = kCiThreads * kCiItems;

template <typename T>
__global__ void __launch_bounds__(kCiThreads)
transform_if_kernel(const T *__restrict__ in, T *__restrict__ out, size_t n, int unary, Pred pred, TileState<unsigned> ts, unsigned epoch,
                    unsigned long long *ticket, unsigned long long ticket_base, unsigned long long *total_out)
{
    __shared__ unsigned long long s_tile;
    __shared__ unsigned s_warp_count[kCiThreads / 32];
    __shared__ unsigned s_prefix;
    __shared__ unsigned s_window[1][32];
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1ull) - ticket_base;
    __syncthreads();
    const size_t tile = (size_t)s_tile;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const size_t base = tile * (size_t)kCiTile + (size_t)warp * (kCiItems * 32) + lane;
    const T pa = from_bits<T>(pred.a_bits), pb = from_bits<T>(pred.b_bits);

    T x[kCiItems];
    unsigned keep = 0;  // bit i: element i of this lane is selected
#pragma unroll
    for (int i = 0; i < kCiItems; i++) {
        const size_t idx = base + (size_t)i * 32;
        if (idx < n) {
            x[i] = in[idx];
            if (eval_pred<T>(x[i], pred.arith, pred.cmp, pa, pb)) keep |= 1u << i;
        } else {
            x[i] = T();
        }
    }
    // rank inside the warp's chunk: elements are ordered (i, lane)
    unsigned rank[kCiItems];
    unsigned running = 0;
#pragma unroll
    for (int i = 0; i < kCiItems; i++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (keep >> i) & 1u);
        rank[i] = running + __popc(bal & lanemask_lt());
        running += __popc(bal);
    }
    if (lane == 0) s_warp_count[warp] = running;
    __syncthreads();
    unsigned warp_off = 0, count = 0;
#pragma unroll
    for (int w = 0; w < kCiThreads / 32; w++) {
        const unsigned c = s_warp_count[w];
        if (w < (int)warp) warp_off += c;
        count += c;
    }
    if (warp == 0) {
        unsigned prefix = 0;
        if (tile == 0) {
            if (lane == 0) ts.post(0, epoch, kInclusive, count);
        } else {
            if (lane == 0) ts.post(tile, epoch, kPartial, count);
            prefix = lookback_prefix<unsigned, BCB_PLUS>(ts, tile, epoch, s_window);
            if (lane == 0) ts.post(tile, epoch, kInclusive, prefix + count);
        }
        if (lane == 0) {
            s_prefix = prefix;
            if ((tile + 1) * (size_t)kCiTile >= n) *total_out = (unsigned long long)prefix + count;  // the last tile knows the total
        }
    }
    __syncthreads();
    const size_t dst = (size_t)s_prefix + warp_off;
#pragma unroll
    for (int i = 0; i < kCiItems; i++)
        if ((keep >> i) & 1u) out[dst + rank[i]] = apply_unary<T>(unary, x[i]);
}

// ---- count_if: transform to 0 / 1 and sum (exact) -------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) count_if_kernel(const T *__restrict__ in, size_t n, Pred pred, unsigned long long *counter)
{
    const T pa = from_bits<T>(pred.a_bits), pb = from_bits<T>(pred.b_bits);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) c += eval_pred<T>(in[i], pred.arith, pred.cmp, pa, pb) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ unsigned s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(&s_sum, c);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(counter, (unsigned long long)s_sum);
}

// ---- transform_reduce / inner_product: fused load, transform, fold ---------------------------------------------------
// value(i) = in2 ? binary(tcode, in1[i], in2[i]) : unary(tcode, in1[i]);  result = fold of the values with OP, in T.
template <typename T, int OP>
__global__ void __launch_bounds__(256)
transform_reduce_kernel(const T *__restrict__ in1, const T *__restrict__ in2, size_t n, int tcode, T *partials, unsigned *done_counter, T *result)
{
    __shared__ T smem[32];
    __shared__ bool is_last;
    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    T acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = Op<OP, T>::identity();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * gthreads < n; i += 4 * gthreads) {  // four independent loads per stream in flight
        T a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = in1[i + u * gthreads];
            if (in2) b[u] = in2[i + u * gthreads];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = Op<OP, T>::apply(acc[u], in2 ? apply_binary<T>(tcode, a[u], b[u]) : apply_unary<T>(tcode, a[u]));
    }
    for (; i < n; i += gthreads) acc[0] = Op<OP, T>::apply(acc[0], in2 ? apply_binary<T>(tcode, in1[i], in2[i]) : apply_unary<T>(tcode, in1[i]));
    T a = Op<OP, T>::apply(Op<OP, T>::apply(acc[0], acc[1]), Op<OP, T>::apply(acc[2], acc[3]));
    a = block_reduce<T, OP>(a, smem);
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
    T p = Op<OP, T>::identity();  // fixed (index) order: deterministic for floats
    for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) p = Op<OP, T>::apply(p, ((volatile T *)partials)[j]);
    p = block_reduce<T, OP>(p, smem);
    if (threadIdx.x == 0) {
        *result = p;
        *done_counter = 0;
    }
}

// ---- reduce_by_key: one pass, segmented decoupled look-back ---------------------------------------------------------
// Runs of consecutive equal keys are folded with OP; segment j (in input order) -> keys_out[j], vals_out[j].
// Tile aggregate = (heads in the tile, "tile contains a head", fold of the tile's trailing segment); the prefix of a tile
// is the fold of all earlier aggregates under  (h1,f1,x1) o (h2,f2,x2) = (h1+h2, f1|f2, f2 ? x2 : x1 op x2).
constexpr int kRkThreads = 256, kRkItems = 8, kRkTile = kRkThreads * kRkItems;
struct alignas(16) RkRecord {  // one per tile, its own arena (kArenaSegmented): fields are only ever read as what they were written as
    unsigned status, p_heads, p_flag, i_heads;
    unsigned long long p_val, i_val;
    unsigned long long pad[2];
};
static_assert(sizeof(RkRecord) == 48, "record layout is part of the arena contract");

// flag: 0 = elements but no head, 1 = contains a head, kSegEmpty = no elements at all (the identity of seg_combine)
constexpr unsigned kSegEmpty = 2u;
template <typename V> struct SegAgg { unsigned heads, flag; V x; };
template <typename V> __device__ __forceinline__ SegAgg<V> seg_empty()
{
    SegAgg<V> r;
    r.heads = 0; r.flag = kSegEmpty; r.x = V();
    return r;
}

template <typename V>
__device__ __forceinline__ SegAgg<V> seg_combine(int op, const SegAgg<V> &a, const SegAgg<V> &b)
{
    if (b.flag == kSegEmpty) return a;
    if (a.flag == kSegEmpty) return b;
    SegAgg<V> r;
    r.heads = a.heads + b.heads;
    r.flag = a.flag | b.flag;
    r.x = b.flag ? b.x : apply_binary<V>(op, a.x, b.x);
    return r;
}
template <typename V> __device__ __forceinline__ SegAgg<V> seg_shfl(const SegAgg<V> &a, int src)
{
    SegAgg<V> r;
    r.heads = __shfl_sync(0xffffffffu, a.heads, src);
    r.flag = __shfl_sync(0xffffffffu, a.flag, src);
    r.x = shfl_t(a.x, src);
    return r;
}
template <typename V> __device__ __forceinline__ SegAgg<V> seg_shfl_up(const SegAgg<V> &a, int d)
{
    SegAgg<V> r;
    r.heads = __shfl_up_sync(0xffffffffu, a.heads, d);
    r.flag = __shfl_up_sync(0xffffffffu, a.flag, d);
    r.x = shfl_up_t(a.x, d);
    return r;
}

template <typename K, typename V>
__global__ void __launch_bounds__(kRkThreads)
reduce_by_key_kernel(const K *__restrict__ keys, const V *__restrict__ vals, size_t n, K *__restrict__ keys_out, V *__restrict__ vals_out, int op,
                     RkRecord *rec, unsigned epoch, unsigned long long *ticket, unsigned long long ticket_base, unsigned long long *total_out)
{
    __shared__ unsigned long long s_tile;
    __shared__ K s_keys[kRkTile + 2];  // [0] = key before the tile, [1 .. kRkTile] the tile, [kRkTile + 1] = key after it
    __shared__ V s_vals[kRkTile];
    __shared__ SegAgg<V> s_warp[kRkThreads / 32];
    __shared__ SegAgg<V> s_carry;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1ull) - ticket_base;
    __syncthreads();
    const size_t tile = (size_t)s_tile, tile_base = tile * (size_t)kRkTile;
    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned valid = (unsigned)((n - tile_base) < (size_t)kRkTile ? (n - tile_base) : (size_t)kRkTile);
    // coalesced load into shared memory, then every thread takes kRkItems consecutive elements
#pragma unroll
    for (int i = 0; i < kRkItems; i++) {
        const unsigned j = i * kRkThreads + tid;
        if (j < valid) {
            s_keys[1 + j] = keys[tile_base + j];
            s_vals[j] = vals[tile_base + j];
        }
    }
    if (tid == 0) {
        if (tile_base > 0) s_keys[0] = keys[tile_base - 1];
        if (tile_base + valid < n) s_keys[1 + valid] = keys[tile_base + valid];
    }
    __syncthreads();
    const unsigned first = tid * kRkItems;
    K k[kRkItems + 2];  // k[0] = predecessor, k[kRkItems + 1] = successor
    V v[kRkItems];
#pragma unroll
    for (int i = 0; i < kRkItems + 2; i++) k[i] = (first + i <= valid + 1) ? s_keys[first + i] : K();
#pragma unroll
    for (int i = 0; i < kRkItems; i++) v[i] = first + i < valid ? s_vals[first + i] : V();
    // head flags, thread-local segmented inclusive fold
    unsigned head = 0, heads_upto[kRkItems];
    V s[kRkItems];
    SegAgg<V> mine = seg_empty<V>();
#pragma unroll
    for (int i = 0; i < kRkItems; i++) {
        if (first + i < valid) {
            const bool h = (tile_base + first + i == 0) || !(k[i + 1] == k[i]);
            if (h) head |= 1u << i;
            SegAgg<V> e;
            e.heads = h ? 1u : 0u; e.flag = h ? 1u : 0u; e.x = v[i];
            mine = seg_combine<V>(op, mine, e);
        }
        s[i] = mine.x;
        heads_upto[i] = mine.heads;
    }
    // exclusive segmented scan of the thread aggregates: warp scan by shuffles, then the warps in order
    SegAgg<V> incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const SegAgg<V> o = seg_shfl_up(incl, d);
        if ((int)lane >= d) incl = seg_combine<V>(op, o, incl);
    }
    if (lane == 31) s_warp[warp] = incl;
    SegAgg<V> carry = seg_shfl_up(incl, 1);  // the lanes before this one
    if (lane == 0) carry = seg_empty<V>();
    __syncthreads();
    SegAgg<V> before = seg_empty<V>(), tile_agg = seg_empty<V>();
#pragma unroll
    for (int w = 0; w < kRkThreads / 32; w++) {
        const SegAgg<V> a = s_warp[w];
        if (w < (int)warp) before = seg_combine<V>(op, before, a);
        tile_agg = seg_combine<V>(op, tile_agg, a);
    }
    carry = seg_combine<V>(op, before, carry);  // everything before this thread inside the tile

    // ---- publish the tile aggregate, look back (warp 0), publish the inclusive prefix ----
    if (warp == 0) {
        SegAgg<V> prefix = seg_empty<V>();  // fold of all earlier tiles
        if (tile > 0) {
            if (lane == 0) {
                unsigned long long bits = 0;
                memcpy(&bits, &tile_agg.x, sizeof(V));
                rec[tile].p_heads = tile_agg.heads;
                rec[tile].p_flag = tile_agg.flag;
                *((volatile unsigned long long *)&rec[tile].p_val) = bits;
                st_release_u32(&rec[tile].status, (epoch << 2) | kPartial);
            }
            long long basei = (long long)tile - 1;
            while (true) {
                const long long idx = basei - (long long)lane;
                unsigned st = kInclusive;  // tiles "before 0": inclusive, empty
                SegAgg<V> a = seg_empty<V>();
                if (idx >= 0) {
                    unsigned tag;
                    while (((tag = ld_acquire_u32(&rec[idx].status)) >> 2) != epoch) __nanosleep(kSpinBackoffNs);
                    st = tag & 3u;
                    unsigned long long bits;
                    if (st == kPartial) {
                        a.heads = *((volatile unsigned *)&rec[idx].p_heads);
                        a.flag = *((volatile unsigned *)&rec[idx].p_flag);
                        bits = *((volatile unsigned long long *)&rec[idx].p_val);
                    } else {  // everything up to and including that tile: starts with the head of element 0
                        a.heads = *((volatile unsigned *)&rec[idx].i_heads);
                        a.flag = 1;
                        bits = *((volatile unsigned long long *)&rec[idx].i_val);
                    }
                    memcpy(&a.x, &bits, sizeof(V));
                }
                const unsigned inc = __ballot_sync(0xffffffffu, st == kInclusive);
                const int stop = inc ? (__ffs(inc) - 1) : 31;  // the nearest inclusive descriptor ends the walk
                // fold lanes stop .. 0 (oldest first), then put the window in front of what was gathered so far
                SegAgg<V> win = seg_empty<V>();
                for (int l = stop; l >= 0; --l) win = seg_combine<V>(op, win, seg_shfl(a, l));
                prefix = seg_combine<V>(op, win, prefix);
                if (inc) break;
                basei -= 32;
            }
        }
        const SegAgg<V> inclusive = seg_combine<V>(op, prefix, tile_agg);
        if (lane == 0) {
            unsigned long long bits = 0;
            memcpy(&bits, &inclusive.x, sizeof(V));
            rec[tile].i_heads = inclusive.heads;
            *((volatile unsigned long long *)&rec[tile].i_val) = bits;
            st_release_u32(&rec[tile].status, (epoch << 2) | kInclusive);
            s_carry = prefix;
            if (tile_base + valid >= n) *total_out = inclusive.heads;  // the last tile knows the number of segments
        }
    }
    __syncthreads();
    carry = seg_combine<V>(op, s_carry, carry);  // everything before this thread: earlier tiles, then this tile
    // emit the end of every segment
#pragma unroll
    for (int i = 0; i < kRkItems; i++) {
        if (first + i < valid) {
            const bool last_of_segment = (tile_base + first + i + 1 == n) || !(k[i + 2] == k[i + 1]);
            if (last_of_segment) {
                const unsigned heads_here = carry.heads + heads_upto[i];        // heads up to and including element i
                const bool started_here = (head & ((2u << i) - 1u)) != 0;       // a head among this thread's elements 0 .. i
                const V value = (started_here || carry.flag == kSegEmpty) ? s[i] : apply_binary<V>(op, carry.x, s[i]);
                keys_out[heads_here - 1] = k[i + 1];
                vals_out[heads_here - 1] = value;
            }
        }
    }
}

static int grid_cap(size_t n, int per_thread, int sm_count)
{
    size_t blocks = (n + (size_t)256 * per_thread - 1) / ((size_t)256 * per_thread);
    const size_t cap = (size_t)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static Pred make_pred(const bcb_pred *p)
{
    Pred r;
    r.arith = p->arith; r.cmp = p->cmp; r.a_bits = p->a_bits; r.b_bits = p->b_bits;
    return r;
}

template <typename T>
static int launch_transform_if(StreamState *st, const void *in, void *out, size_t n, int unary, const Pred &pred, unsigned long long *total_dev)
{
    const size_t tiles = (n + kCiTile - 1) / kCiTile;
    if (tiles > 0x7fffffffull) return BCB_ETOOLARGE;
    void *mem;
    BCB_TRY(lookback_reserve(st, kArenaPacked, TileState<unsigned>::bytes(tiles), &mem));
    unsigned epoch;
    BCB_TRY(next_epoch(st, kArenaPacked, &epoch));
    TileState<unsigned> ts;
    ts.bind(mem, tiles);
    const unsigned long long base = ticket_reserve(st, tiles);
    LaunchTimer timer(st, BCB_K_SCAN);
    transform_if_kernel<T><<<(unsigned)tiles, kCiThreads, 0, st->stream>>>((const T *)in, (T *)out, n, unary, pred, ts, epoch,
                                                                         st->control + kControlTicket, base, total_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename T, int OP>
static int launch_transform_reduce(StreamState *st, const void *in1, const void *in2, size_t n, int tcode, void *result_dev)
{
    constexpr int kMaxBlocks = 148 * 16;
    void *partials;
    BCB_TRY(scratch_reserve(st, (size_t)kMaxBlocks * sizeof(T), &partials));
    unsigned *counter = reinterpret_cast<unsigned *>(st->control + kControlReduceDone);
    int grid = grid_cap(n, 8, st->sm_count);
    if (grid > kMaxBlocks) grid = kMaxBlocks;
    LaunchTimer timer(st, BCB_K_REDUCE);
    transform_reduce_kernel<T, OP><<<grid, 256, 0, st->stream>>>((const T *)in1, (const T *)in2, n, tcode, (T *)partials, counter, (T *)result_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename T>
static int dispatch_transform_reduce(StreamState *st, int reduce_op, const void *in1, const void *in2, size_t n, int tcode, void *result_dev)
{
    switch (reduce_op) {
    case BCB_PLUS: return launch_transform_reduce<T, BCB_PLUS>(st, in1, in2, n, tcode, result_dev);
    case BCB_MULTIPLIES: return launch_transform_reduce<T, BCB_MULTIPLIES>(st, in1, in2, n, tcode, result_dev);
    case BCB_MIN: return launch_transform_reduce<T, BCB_MIN>(st, in1, in2, n, tcode, result_dev);
    case BCB_MAX: return launch_transform_reduce<T, BCB_MAX>(st, in1, in2, n, tcode, result_dev);
    default: break;
    }
    if constexpr (!is_fp<T>::value) {
        if (reduce_op == BCB_BIT_OR) return launch_transform_reduce<T, BCB_BIT_OR>(st, in1, in2, n, tcode, result_dev);
    }
    return BCB_EUNSUPPORTED;
}

template <typename K, typename V>
static int launch_reduce_by_key(StreamState *st, const void *keys, const void *vals, size_t n, void *keys_out, void *vals_out, int op,
                                unsigned long long *total_dev)
{
    const size_t tiles = (n + kRkTile - 1) / kRkTile;
    if (tiles > 0x7fffffffull) return BCB_ETOOLARGE;
    void *mem;
    BCB_TRY(lookback_reserve(st, kArenaSegmented, tiles * sizeof(RkRecord), &mem));
    unsigned epoch;
    BCB_TRY(next_epoch(st, kArenaSegmented, &epoch));
    const unsigned long long base = ticket_reserve(st, tiles);
    LaunchTimer timer(st, BCB_K_SCAN);
    reduce_by_key_kernel<K, V><<<(unsigned)tiles, kRkThreads, 0, st->stream>>>((const K *)keys, (const V *)vals, n, (K *)keys_out, (V *)vals_out, op,
                                                                             (RkRecord *)mem, epoch, st->control + kControlTicket, base, total_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename K>
static int dispatch_rbk_values(StreamState *st, int val_dtype, const void *keys, const void *vals, size_t n, void *keys_out, void *vals_out, int op,
                               unsigned long long *total_dev)
{
    switch (val_dtype) {
    case BCB_INT: return launch_reduce_by_key<K, int>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    case BCB_UINT: return launch_reduce_by_key<K, unsigned>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    case BCB_LONG: return launch_reduce_by_key<K, long long>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    case BCB_ULONG: return launch_reduce_by_key<K, unsigned long long>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    case BCB_FLOAT: return launch_reduce_by_key<K, float>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    case BCB_DOUBLE: return launch_reduce_by_key<K, double>(st, keys, vals, n, keys_out, vals_out, op, total_dev);
    default: return BCB_EUNSUPPORTED;  // 1- and 2-byte value types: not built
    }
}

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_transform_if(bcb_stream stream, int dtype, const void *in, size_t n, int unary, const bcb_pred *pred, void *out, size_t *count_host)
{
    if (count_host) *count_host = 0;
    if (!dtype_size(dtype) || !pred || !count_host) return BCB_EINVAL;
    if (unary < BCB_UN_IDENTITY || unary > BCB_UN_SQUARE || pred->cmp < BCB_CMP_EQ || pred->cmp > BCB_CMP_TRUE || pred->arith < BCB_AR_NONE || pred->arith > BCB_AR_AND)
        return BCB_EINVAL;
    if ((pred->arith == BCB_AR_MOD || pred->arith == BCB_AR_AND) && dtype_is_float(dtype)) return BCB_EUNSUPPORTED;
    if (n == 0) return BCB_SUCCESS;
    if (!in || !out) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    unsigned long long *total_dev = (unsigned long long *)st->pinned_slot_dev;
    const Pred p = make_pred(pred);
    int rc;
    switch (dtype) {
#define X(DT, T) case DT: rc = launch_transform_if<T>(st, in, out, n, unary, p, total_dev); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_TRY(rc);
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));  // the returned iterator (result + count) is a host value
    *count_host = (size_t)(*(volatile unsigned long long *)st->pinned_slot);
    return BCB_SUCCESS;
}

int bcb_count_if(bcb_stream stream, int dtype, const void *in, size_t n, const bcb_pred *pred, unsigned long long *count_host)
{
    if (!count_host) return BCB_EINVAL;
    *count_host = 0;
    if (!dtype_size(dtype) || !pred) return BCB_EINVAL;
    if ((pred->arith == BCB_AR_MOD || pred->arith == BCB_AR_AND) && dtype_is_float(dtype)) return BCB_EUNSUPPORTED;
    if (n == 0) return BCB_SUCCESS;
    if (!in) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    unsigned long long *counter = st->control + kControlCount;
    BCB_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st->stream));
    const Pred p = make_pred(pred);
    const int grid = grid_cap(n, 8, st->sm_count);
    switch (dtype) {
#define X(DT, T) case DT: count_if_kernel<T><<<grid, 256, 0, st->stream>>>((const T *)in, n, p, counter); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_CUDA_TRY(cudaGetLastError());
    BCB_CUDA_TRY(cudaMemcpyAsync(st->pinned_slot, counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream));
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    *count_host = *(volatile unsigned long long *)st->pinned_slot;
    return BCB_SUCCESS;
}

int bcb_transform_reduce(bcb_stream stream, int dtype, const void *in1, const void *in2, size_t n, int transform, int reduce_op, void *result,
                         int result_is_device)
{
    const size_t w = dtype_size(dtype);
    if (!w) return BCB_EINVAL;
    if (n == 0) return BCB_SUCCESS;  // like reduce: the result is left untouched
    if (!in1 || !result) return BCB_EINVAL;
    if (in2 ? !(transform == BCB_PLUS || transform == BCB_MINUS || transform == BCB_MULTIPLIES || transform == BCB_MIN || transform == BCB_MAX ||
                (op_is_bitwise(transform) && !dtype_is_float(dtype)))
            : (transform < BCB_UN_IDENTITY || transform > BCB_UN_SQUARE))
        return BCB_EUNSUPPORTED;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    void *dst = result_is_device ? result : st->pinned_slot_dev;
    int rc;
    switch (dtype) {
#define X(DT, T) case DT: rc = dispatch_transform_reduce<T>(st, reduce_op, in1, in2, n, transform, dst); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_TRY(rc);
    if (result_is_device) return BCB_SUCCESS;
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    std::memcpy(result, st->pinned_slot, w);
    return BCB_SUCCESS;
}

int bcb_reduce_by_key(bcb_stream stream, int key_dtype, int val_dtype, const void *keys_in, const void *vals_in, size_t n, void *keys_out,
                      void *vals_out, int op, size_t *count_host)
{
    if (!count_host) return BCB_EINVAL;
    *count_host = 0;
    if (!dtype_size(key_dtype) || !dtype_size(val_dtype)) return BCB_EINVAL;
    if (!(op == BCB_PLUS || op == BCB_MULTIPLIES || op == BCB_MIN || op == BCB_MAX || (!dtype_is_float(val_dtype) && op_is_bitwise(op)))) return BCB_EUNSUPPORTED;
    if (n == 0) return BCB_SUCCESS;
    if (!keys_in || !vals_in || !keys_out || !vals_out) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    unsigned long long *total_dev = (unsigned long long *)st->pinned_slot_dev;
    int rc;
    switch (key_dtype) {  // keys are only compared for equality: integers by width, floating point with ==
    case BCB_CHAR: case BCB_UCHAR: rc = dispatch_rbk_values<unsigned char>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    case BCB_SHORT: case BCB_USHORT: rc = dispatch_rbk_values<unsigned short>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    case BCB_INT: case BCB_UINT: rc = dispatch_rbk_values<unsigned>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    case BCB_LONG: case BCB_ULONG: rc = dispatch_rbk_values<unsigned long long>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    case BCB_FLOAT: rc = dispatch_rbk_values<float>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    default: rc = dispatch_rbk_values<double>(st, val_dtype, keys_in, vals_in, n, keys_out, vals_out, op, total_dev); break;
    }
    BCB_TRY(rc);
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));  // the returned iterator pair is a host value
    *count_host = (size_t)(*(volatile unsigned long long *)st->pinned_slot);
    return BCB_SUCCESS;
}

}  // extern "C"
