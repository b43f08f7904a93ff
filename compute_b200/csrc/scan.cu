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
#include "tile_state.cuh"
#include "scan_ws.cuh"
#include "tma.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>

namespace bcb {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanItems = 16;                         // elements per thread
constexpr int kScanTile = kScanThreads * kScanItems;   // 4096
template <typename T, int OP>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const T *in, T *out, size_t n, int exclusive, T init, TileState<T> ts, unsigned epoch,
            unsigned long long *ticket, unsigned long long ticket_base, const T *__restrict__ init_dev = nullptr)
{
    if (init_dev) init = *init_dev;  // the seed lives on the device (bcb_scan_with_carry)
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

// ---- persistent, TMA-pipelined variant -------------------------------------------------------------------
// The one-tile-per-CTA kernel above exposes every tile's load latency and its look-back latency to the CTA
// (ncu: warps mostly parked at the barrier while warp 0 looks back; 37 % of HBM peak).  Here the grid is sized to
// the resident CTAs, each CTA owns a ring of kStages shared-memory stages that the bulk-copy engine fills ahead of
// time (cp.async.bulk global -> shared, completion on an mbarrier), scans the tile in place in shared memory and
// hands it back to the copy engine (cp.async.bulk shared -> global).  Loads of the next tiles are therefore in
// flight during every phase of the current one, and no registers are tied up by data in flight.
// Tiles are dealt round-robin: CTA b processes tiles b, b + G, b + 2G ... (G = grid size = resident CTAs), so the
// tiles in flight at any moment form one contiguous window of the input and a prefetched tile is never one that
// another CTA is waiting for.  (Drawing tickets ahead of time was measured to be 2x slower: a CTA then HOLDS
// tiles it is not working on yet while their successors spin in the look-back.)
constexpr int kScanWsMinLog2Bytes = 20;               // smallest range (bytes) the warp-specialised kernel takes
constexpr int kRoundThreads = 512;                 // threads per CTA of the round-synchronous kernel
constexpr int kRoundWarps = kRoundThreads / 32;

template <typename T> struct ScanRing {
    static constexpr int kItems = sizeof(T) == 1 ? 16 : (sizeof(T) == 8 ? 4 : 8);  // elements per thread
    static constexpr int kTile = kRoundThreads * kItems;               // 4096 elements (8192 / 2048 for 1- / 8-byte types)
    static constexpr int kTileBytes = kTile * (int)sizeof(T);          // 16 KiB (8 KiB for 1- and 2-byte types)
    static constexpr int kStages = 4;
    static constexpr size_t kBytes = (size_t)kStages * kTileBytes + 1024;  // stages + barriers / bookkeeping
};

// Round-synchronous, software-pipelined look-back.
// With G persistent CTAs and round-robin tiles, round r processes the contiguous window [rG, (r+1)G).  The prefix of
// tile t = rG + b is
//     carry(r)  op  fold(aggregate(rG), ..., aggregate(rG + b - 1)),
// carry(r) being the inclusive prefix published by the LAST tile of round r-1.  All b aggregates are fetched in ONE
// parallel step (thread i polls the descriptor of tile rG + i) and folded with a fixed-shape block reduction: one L2
// round trip instead of b/32, every thread takes part, and a tile's prefix is a pure function of its position and the
// data, so floating-point results are run-to-run deterministic.  Only the last tile of a round publishes an
// inclusive value.
// Each iteration runs phase A of tile `it` (tile -> registers, tile-local scan written back to the stage, aggregate
// published, first poll of the round's descriptors issued) and then phase C of tile `it-1` (finish the poll, fold the
// prefix, add it to the stage, hand the stage to the bulk-copy engine).  The descriptors a tile waits for therefore
// have a whole phase A of slack, and the bulk loads of the next two tiles are in flight all the time.
template <typename T, int OP>
__global__ void __launch_bounds__(kRoundThreads)
scan_tma_kernel(const T *in, T *out, size_t n, int exclusive, T init, TileState<T> ts, unsigned epoch, size_t num_tiles,
                const T *__restrict__ init_dev = nullptr)
{
    if (init_dev) init = *init_dev;  // the seed lives on the device (bcb_scan_with_carry)
    typedef Op<OP, T> O;
    typedef ScanRing<T> R;
    constexpr int VEC = 16 / sizeof(T);
    constexpr int NV = R::kItems / VEC;
    constexpr int S = R::kStages;
    constexpr int TILE = R::kTile;
    static_assert(NV >= 1, "vector wider than the per-thread item count");

    extern __shared__ __align__(128) unsigned char ring_raw[];
    T *stage_base = reinterpret_cast<T *>(ring_raw);
    unsigned long long *full_bar = reinterpret_cast<unsigned long long *>(ring_raw + (size_t)S * R::kTileBytes);  // [S]
    __shared__ T s_warp_total[kRoundWarps];
    __shared__ T s_fold[kRoundWarps];
    __shared__ T s_carry;

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned G = gridDim.x, b = blockIdx.x;
    const bool last_in_round = (b == G - 1);
    const unsigned warp_elem = warp * (32 * R::kItems);
    auto tile_of = [&](unsigned it) { return (size_t)b + (size_t)it * G; };

    auto issue_load = [&](int s, size_t tile) {  // thread 0 only
        if (tile < num_tiles) {
            const size_t base = tile * (size_t)TILE;
            if (base + TILE <= n) {  // full tile: bulk copy; a partial last tile is loaded with guarded loads instead
                mbar_expect_tx(&full_bar[s], (unsigned)R::kTileBytes);
                tma_load_1d(stage_base + (size_t)s * TILE, in + base, (unsigned)R::kTileBytes, &full_bar[s]);
            }
        }
    };

    if (tid == 0) {
        for (int s = 0; s < S; s++) mbar_init(&full_bar[s], 1);
        mbar_init_fence();
        for (int s = 0; s < S; s++) issue_load(s, tile_of((unsigned)s));
    }
    __syncthreads();

    // state of the tile whose phase C is pending
    bool pending = false;
    T pend_aggregate = O::identity();
    T poll_v = O::identity();        // threads < b: aggregate of an earlier tile of the round (if poll_ok)
    T poll_c = O::identity();        // last thread: carry of the round (if poll_ok)
    bool poll_ok = true;

    // ordered fold of the 16 per-warp values in shared memory, done by every warp with shuffles:
    // before = fold of vals[0 .. upto), all = fold of all 16
    auto fold16 = [&](const T *vals, unsigned upto, T &before, T &all) {
        T t = (lane < (unsigned)kRoundWarps) ? vals[lane] : O::identity();
#pragma unroll
        for (int d = 1; d < kRoundWarps; d <<= 1) {
            const T o = shfl_up_t(t, d);
            if ((int)lane >= d) t = O::apply(o, t);
        }
        all = shfl_t(t, kRoundWarps - 1);
        const T prev = shfl_t(t, upto == 0 ? 0 : (int)upto - 1);
        before = upto == 0 ? O::identity() : prev;
    };

    for (unsigned it = 0;; ++it) {
        const size_t tileA = tile_of(it);
        const bool doA = tileA < num_tiles;
        if (!doA && !pending) break;

        T aggregate = O::identity();
        if (doA) {
            // ================= phase A: tile `it` =================
            const int s = (int)(it % S);
            const size_t tile_base = tileA * (size_t)TILE;
            const bool full = tile_base + TILE <= n;
            T *stage = stage_base + (size_t)s * TILE;
            T x[NV][VEC];
            if (full) {
                mbar_wait(&full_bar[s], (it / S) & 1u);
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(stage + warp_elem + (j * 32 + lane) * VEC);
                    const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
                    for (int k = 0; k < VEC; k++) x[j][k] = e[k];
                }
            } else {
#pragma unroll
                for (int j = 0; j < NV; j++) {
#pragma unroll
                    for (int k = 0; k < VEC; k++) {
                        const size_t i = tile_base + warp_elem + (size_t)(j * 32 + lane) * VEC + k;
                        x[j][k] = i < n ? in[i] : O::identity();
                    }
                }
            }
            // vector-local scan, warp scans, block aggregate
            T vsum[NV];
#pragma unroll
            for (int j = 0; j < NV; j++) {
#pragma unroll
                for (int k = 1; k < VEC; k++) x[j][k] = O::apply(x[j][k - 1], x[j][k]);
                vsum[j] = x[j][VEC - 1];
            }
            T vexcl[NV];
            T carry = O::identity();
#pragma unroll
            for (int j = 0; j < NV; j++) {
                T sc = vsum[j];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const T o = shfl_up_t(sc, d);
                    if ((int)lane >= d) sc = O::apply(o, sc);
                }
                T e = shfl_up_t(sc, 1);
                if (lane == 0) e = O::identity();
                vexcl[j] = O::apply(carry, e);
                carry = O::apply(carry, shfl_t(sc, 31));
            }
            if (lane == 0) s_warp_total[warp] = carry;
            __syncthreads();
            T warp_off;
            fold16(s_warp_total, warp, warp_off, aggregate);
            if (tid == 0 && !last_in_round) ts.post(tileA, epoch, kPartial, aggregate);
            // tile-local scan back into the stage (phase C adds the tile prefix)
#pragma unroll
            for (int j = 0; j < NV; j++) {
                const T p = O::apply(warp_off, vexcl[j]);
                T y[VEC];
                if (exclusive == 1) {
                    y[0] = p;
#pragma unroll
                    for (int k = 1; k < VEC; k++) y[k] = O::apply(p, x[j][k - 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; k++) y[k] = O::apply(p, x[j][k]);
                }
                *reinterpret_cast<uint4 *>(stage + warp_elem + (j * 32 + lane) * VEC) = *reinterpret_cast<const uint4 *>(y);
            }
        }

        if (pending) {
            // ================= phase C: tile `it - 1` =================
            const unsigned pit = it - 1;
            const size_t tileC = tile_of(pit);
            const int s = (int)(pit % S);
            const size_t tile_base = tileC * (size_t)TILE;
            const bool full = tile_base + TILE <= n;
            T *stage = stage_base + (size_t)s * TILE;
            // finish the poll that phase A of that tile started
            if (!poll_ok) {
                if (tid < b) {
                    const size_t j = (size_t)pit * G + tid;
                    while (ts.peek(j, epoch, poll_v) == kInvalid) __nanosleep(kSpinBackoffNs);
                } else if (tid == kRoundThreads - 1) {
                    const size_t j = (size_t)pit * G - 1;
                    while (ts.peek(j, epoch, poll_c) != kInclusive) __nanosleep(kSpinBackoffNs);
                }
            }
            if (tid == kRoundThreads - 1) s_carry = poll_c;
            T v = (tid < b) ? poll_v : O::identity();
            // fixed-shape fold: ordered shuffle scan inside each warp, then a left fold over the 16 warp values
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const T o = shfl_up_t(v, off);
                if ((int)lane >= off) v = O::apply(o, v);
            }
            if (lane == 31) s_fold[warp] = v;
            __syncthreads();
            T unused, folded;
            fold16(s_fold, 0, unused, folded);
            const T prefix = O::apply(s_carry, folded);
            if (tid == 0 && last_in_round) ts.post(tileC, epoch, kInclusive, O::apply(prefix, pend_aggregate));
#pragma unroll
            for (int j = 0; j < NV; j++) {
                T *slot = stage + warp_elem + (j * 32 + lane) * VEC;
                uint4 raw = *reinterpret_cast<const uint4 *>(slot);
                T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
                for (int k = 0; k < VEC; k++) e[k] = O::apply(prefix, e[k]);
                if (full) {
                    *reinterpret_cast<uint4 *>(slot) = raw;
                } else {
                    const size_t i0 = tile_base + warp_elem + (size_t)(j * 32 + lane) * VEC;
#pragma unroll
                    for (int k = 0; k < VEC; k++)
                        if (i0 + k < n) out[i0 + k] = e[k];
                }
            }
            fence_proxy_async();  // make the generic-proxy writes to the stage visible to the bulk-copy engine
            __syncthreads();
            if (tid == 0) {
                if (full) tma_store_1d(out + tile_base, stage, (unsigned)R::kTileBytes);
                if (pit > 0) {  // the store issued one iteration ago has drained: recycle its stage
                    tma_store_wait_read<1>();
                    issue_load((int)((pit - 1) % S), tile_of(pit - 1 + S));
                }
            }
        } else if (doA) {
            __syncthreads();  // keep s_warp_total's readers and the next iteration's writers apart
        }

        // first (non-blocking) poll for the tile that just went through phase A
        pending = doA;
        pend_aggregate = aggregate;
        poll_ok = true;
        if (doA) {
            if (tid < b) {
                poll_ok = ts.peek((size_t)it * G + tid, epoch, poll_v) != kInvalid;
            } else if (tid == kRoundThreads - 1) {
                poll_c = exclusive ? init : O::identity();
                if (it > 0) poll_ok = ts.peek((size_t)it * G - 1, epoch, poll_c) == kInclusive;
            }
        }
    }
    if (tid == 0) tma_store_wait_read<0>();  // shared memory must outlive the last bulk stores
}

template <typename A>
__global__ void convert_kernel(const void *in, int in_dtype, A *out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = load_as<A>(in, i, in_dtype);
}

template <typename T, int OP>
static int launch_scan(StreamState *st, const void *in, void *out, size_t n, int exclusive, const void *init_host,
                       const void *init_dev_v = nullptr)
{
    T init = (T)0;
    if (init_host) std::memcpy(&init, init_host, sizeof(T));
    const T *init_dev = (const T *)init_dev_v;  // overrides init when set
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles > 0x7fffffffull) return BCB_ETOOLARGE;
    constexpr int kArena = sizeof(T) <= 4 ? kArenaPacked : kArenaWide;
    static const bool use_tma = [] { const char *e = std::getenv("BCB_SCAN_TMA"); return !(e && e[0] == '0'); }();  // 0: one tile per CTA
    const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    void *mem;
    unsigned epoch;
    TileState<T> ts;
    // Large ranges: the warp-specialised kernel (scan_ws.cuh).  Measured on B200, 2^28 int32: 6.0 TB/s against 4.8 for
    // scan_tma_kernel; faster from 1 MB on (16 us against 25 us for 1-8 MB).
    //   BCB_SCAN_WS=0            keep scan_tma_kernel for large ranges (A/B comparison)
    //   BCB_SCAN_WS_MIN_LOG2=k   test hook: use it from 2^k BYTES on
    static const size_t ws_min_bytes = [] {
        const char *e = std::getenv("BCB_SCAN_WS");
        if (e && e[0] == '0') return (size_t)-1;
        const char *v = std::getenv("BCB_SCAN_WS_MIN_LOG2");
        const int k = v ? std::atoi(v) : kScanWsMinLog2Bytes;
        return (size_t)1 << (k < 10 ? 10 : (k > 40 ? 40 : k));
    }();
    if (aligned && n * sizeof(T) >= ws_min_bytes) {
        constexpr int NV = 3, S = 9, D = 5;
        typedef ScanWsShape<T, NV, S> C;
        const size_t wtiles = (n + C::TILE - 1) / C::TILE;
        // (tagged 64-bit words for every element width: the packed arena, whatever T is)
        BCB_TRY(lookback_reserve(st, kArenaPacked, WsTileState<T>::bytes(wtiles), &mem));
        BCB_TRY(next_epoch(st, kArenaPacked, &epoch));
        WsTileState<T> wts;
        wts.bind(mem);
        auto kernel = scan_ws_kernel<T, OP, NV, S, D>;
        static std::atomic<unsigned long long> configured{0};  // bit per device: > 48 KB dynamic shared memory opted in
        const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
        if (!(configured.load(std::memory_order_acquire) & bit) || !bit) {
            BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
            configured.fetch_or(bit, std::memory_order_release);
        }
        size_t grid = (size_t)st->sm_count;  // one CTA per SM
        if (grid > (size_t)kSwMaxGrid) grid = kSwMaxGrid;
        if (grid > wtiles) grid = wtiles;
        LaunchTimer timer(st, BCB_K_SCAN);
        // every CTA waits for every tile of a round: the whole grid must be resident, which a cooperative launch
        // guarantees (or it fails loudly) whatever else runs on the device
        const T *in_t = (const T *)in;
        T *out_t = (T *)out;
        void *args[] = {(void *)&in_t, (void *)&out_t, (void *)&n, (void *)&exclusive, (void *)&init, (void *)&wts, (void *)&epoch, (void *)&wtiles,
                        (void *)&init_dev};
        BCB_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kernel, dim3((unsigned)grid), dim3(kSwThreads), args, C::SMEM_BYTES, st->stream));
        return BCB_SUCCESS;
    }
    if (use_tma && aligned && n >= (size_t)4 * ScanRing<T>::kTile) {
        typedef ScanRing<T> R;
        const size_t rtiles = (n + R::kTile - 1) / R::kTile;
        // reserve first, then draw the epoch (a reallocation restarts the arena's epoch counter)
        BCB_TRY(lookback_reserve(st, kArena, TileState<T>::bytes(rtiles), &mem));
        BCB_TRY(next_epoch(st, kArena, &epoch));
        ts.bind(mem, rtiles);
        auto kernel = scan_tma_kernel<T, OP>;
        static std::atomic<int> resident[64];
        int per_sm = (st->device < 64) ? resident[st->device].load(std::memory_order_acquire) : 0;
        if (per_sm == 0) {
            BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::kBytes));
            BCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kRoundThreads, R::kBytes));
            if (per_sm < 1) per_sm = 1;
            if (st->device < 64) resident[st->device].store(per_sm, std::memory_order_release);
        }
        size_t grid = (size_t)st->sm_count * (size_t)per_sm;
        if (grid > (size_t)kRoundThreads) grid = kRoundThreads;  // one look-back thread per earlier tile of the round
        if (grid > rtiles) grid = rtiles;
        LaunchTimer timer(st, BCB_K_SCAN);
        // The round-synchronous look-back needs the whole grid resident: a cooperative launch guarantees that (or
        // fails loudly) whatever else runs on the device.
        const T *in_t = (const T *)in;
        T *out_t = (T *)out;
        void *args[] = {(void *)&in_t, (void *)&out_t, (void *)&n, (void *)&exclusive, (void *)&init, (void *)&ts, (void *)&epoch, (void *)&rtiles,
                        (void *)&init_dev};
        BCB_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)kernel, dim3((unsigned)grid), dim3(kRoundThreads), args, R::kBytes, st->stream));
        return BCB_SUCCESS;
    }
    BCB_TRY(lookback_reserve(st, kArena, TileState<T>::bytes(tiles), &mem));
    BCB_TRY(next_epoch(st, kArena, &epoch));
    ts.bind(mem, tiles);
    const unsigned long long base = ticket_reserve(st, tiles);
    LaunchTimer timer(st, BCB_K_SCAN);
    scan_kernel<T, OP><<<(unsigned)tiles, kScanThreads, 0, st->stream>>>(
        (const T *)in, (T *)out, n, exclusive, init, ts, epoch, st->control + kControlTicket, base, init_dev);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// carry of a block-distributed scan, folded on the device: (init op) partial_0 op ... op partial_(rank-1) in rank order over
// the records {value at byte 0, "shard not empty" at byte 8} an all-gather left in `records` (16 bytes per rank)
template <typename T, int OP>
__global__ void fold_carry_kernel(const unsigned char *__restrict__ records, int rank, T init, int have_init, T *__restrict__ carry)
{
    typedef Op<OP, T> O;
    T acc = have_init ? init : O::identity();
    for (int r = 0; r < rank; r++) {
        if (records[r * 16 + 8]) {
            T v;
            unsigned char *q = reinterpret_cast<unsigned char *>(&v);
            for (int b = 0; b < (int)sizeof(T); b++) q[b] = records[r * 16 + b];
            acc = O::apply(acc, v);
        }
    }
    *carry = acc;
}

template <typename T, int OP>
static int launch_scan_with_carry(StreamState *st, const void *in, void *out, size_t n, int exclusive, const void *init_host,
                                  const void *records_dev, int rank)
{
    T init = (T)0;
    if (init_host) std::memcpy(&init, init_host, sizeof(T));
    T *carry = reinterpret_cast<T *>(st->control + kControlCarry);
    fold_carry_kernel<T, OP><<<1, 1, 0, st->stream>>>((const unsigned char *)records_dev, rank, init, exclusive ? 1 : 0, carry);
    BCB_CUDA_TRY(cudaGetLastError());
    // exclusive: seeded with init op carry; inclusive: mode 2 = inclusive scan seeded with a carry (the identity when
    // nothing precedes this rank, which leaves integer results -- and float sums: the identity of plus is -0.0 -- unchanged)
    return launch_scan<T, OP>(st, in, out, n, exclusive ? 1 : 2, nullptr, carry);
}

template <typename T>
static int dispatch_scan_op(StreamState *st, int op, const void *in, void *out, size_t n, int exclusive, const void *init_host,
                            const void *records_dev = nullptr, int rank = 0)
{
    if (records_dev) {
        switch (op) {
        case BCB_PLUS: return launch_scan_with_carry<T, BCB_PLUS>(st, in, out, n, exclusive, init_host, records_dev, rank);
        case BCB_MULTIPLIES: return launch_scan_with_carry<T, BCB_MULTIPLIES>(st, in, out, n, exclusive, init_host, records_dev, rank);
        case BCB_MIN: return launch_scan_with_carry<T, BCB_MIN>(st, in, out, n, exclusive, init_host, records_dev, rank);
        case BCB_MAX: return launch_scan_with_carry<T, BCB_MAX>(st, in, out, n, exclusive, init_host, records_dev, rank);
        default: break;
        }
        if constexpr (!is_fp<T>::value) {
            switch (op) {
            case BCB_BIT_AND: return launch_scan_with_carry<T, BCB_BIT_AND>(st, in, out, n, exclusive, init_host, records_dev, rank);
            case BCB_BIT_OR: return launch_scan_with_carry<T, BCB_BIT_OR>(st, in, out, n, exclusive, init_host, records_dev, rank);
            case BCB_BIT_XOR: return launch_scan_with_carry<T, BCB_BIT_XOR>(st, in, out, n, exclusive, init_host, records_dev, rank);
            default: break;
            }
        }
        return BCB_EUNSUPPORTED;
    }
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

static int scan_impl(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in, void *out, size_t n,
                     const void *init_host, const void *records_dev, int rank);

extern "C" int bcb_scan(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in, void *out,
                        size_t n, const void *init_host)
{
    return scan_impl(stream, in_dtype, out_dtype, op, exclusive, in, out, n, init_host, nullptr, 0);
}

extern "C" int bcb_scan_with_carry(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in, void *out,
                                   size_t n, const void *init_host, const void *records_dev, int rank)
{
    if (!records_dev || rank < 0) return BCB_EINVAL;
    if (exclusive != 0 && exclusive != 1) return BCB_EINVAL;
    return scan_impl(stream, in_dtype, out_dtype, op, exclusive, in, out, n, init_host, records_dev, rank);
}

static int scan_impl(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in, void *out, size_t n,
                     const void *init_host, const void *records_dev, int rank)
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
#define X(DT, T) case DT: return dispatch_scan_op<T>(st, op, src, out, n, exclusive, init_host, records_dev, rank);
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
}
