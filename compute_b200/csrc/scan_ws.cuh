// scan_ws.cuh -- warp-specialised single-pass scan for large, 16-byte aligned ranges on sm_100a.
//
// Replaces scan_impl / local_scan_kernel / write_scanned_output_kernel of the reference
// (algorithm/detail/scan_on_gpu.hpp:26-324) by one launch that reads every element once and writes it once.
//
// Why another kernel: scan_tma_kernel (scan.cu) runs at 0.75 of the copy roofline and ncu shows why -- every thread
// takes part in every per-tile step (look-back poll, two block-wide folds, three block barriers; ~300 instructions
// per thread and 8-element tile, 35 % of the samples parked at a barrier), and a round of tiles (7 MB) lasts only
// ~2 us, the same order as the carry's trip through L2, which sits on the critical path of every CTA.
// Here the roles are split so that nothing on the data path ever waits for another SM:
//   * 16 COMPUTE warps each own a fixed 1/16 of every tile and never synchronise with each other: they wait on
//     mbarriers only.  A tile is visited twice: "R" (as soon as the bulk copy has landed: fold the warp's part,
//     leave the warp total in shared memory) and, D tiles later, "C" (scan in registers, add the warp's offset,
//     write the result back into the stage).  ~10 instructions per element instead of ~37.
//   * one PRODUCER lane drives the bulk-copy engine: cp.async.bulk global->shared S stages ahead, and
//     shared->global for every stage the compute warps have finished.
//   * one AGGREGATE warp folds the 16 warp totals of a tile and publishes the tile aggregate (tile descriptor).
//   * one PREFIX warp per CTA turns aggregates into prefixes: tiles are dealt round-robin (tile = i * G + b), so round i
//     of all CTAs is the contiguous window [i*G, (i+1)*G); the warp polls ALL G aggregates of the round (5 per lane),
//     scans them (fixed shape: floating-point results are run-to-run deterministic and identical in every CTA) and
//     gets its own tile's prefix and the round total at once.  The carry between rounds never leaves the warp's
//     registers, so the serial chain of the whole scan contains no memory round trip at all, and the single hop that
//     remains (aggregate -> L2 -> the other CTAs) has D rounds of slack because R runs D tiles ahead of C.
// All G CTAs must be resident (every CTA waits for every tile of a round): the launcher uses a cooperative launch.
#pragma once

#include "ops.cuh"
#include "tile_state.cuh"
#include "tma.cuh"

namespace bcb {

constexpr int kSwComputeWarps = 16;
constexpr int kSwThreads = (kSwComputeWarps + 3) * 32;  // + producer, aggregate and prefix warps
constexpr int kSwPollPerLane = 5;                       // descriptors per lane of the prefix warp: grids of up to 160 CTAs
constexpr int kSwMaxGrid = 32 * kSwPollPerLane;

template <typename T, int NV, int S> struct ScanWsShape {
    static constexpr int VEC = 16 / (int)sizeof(T);          // elements per 128-bit vector
    static constexpr int WARP_ELEMS = 32 * NV * VEC;         // elements of a tile owned by one compute warp
    static constexpr int TILE = kSwComputeWarps * WARP_ELEMS;
    static constexpr int TILE_BYTES = TILE * (int)sizeof(T);  // NV * 8 KiB
    static constexpr size_t SIDE_BYTES = 2 * (size_t)S * kSwComputeWarps * sizeof(T);  // warp totals + warp offsets
    static constexpr size_t SMEM_BYTES = (size_t)S * TILE_BYTES + SIDE_BYTES + 4 * (size_t)S * sizeof(unsigned long long);
    static_assert(SMEM_BYTES <= 232448, "one CTA per SM: 227 KB of shared memory");
};

// Tile aggregates: 64-bit words {tag = epoch<<2 | kPartial : 32, 32 value bits : 32}, each single-copy atomic and each
// validated by its own tag -- one word for types up to 4 bytes, two (low / high half) for 8-byte types.  No ordering
// between words is needed, so all the loads of a poll are independent and in flight together (the status + value
// records of TileState<T, false> cost two dependent L2 round trips per descriptor).  Same word format as the
// kArenaPacked descriptors of the other kernels: a stale word can never pose as a valid one.
template <typename T> struct WsTileState {
    static constexpr int W = sizeof(T) <= 4 ? 1 : 2;
    unsigned long long *words;
    static size_t bytes(size_t tiles) { return tiles * W * sizeof(unsigned long long); }
    __host__ __device__ void bind(void *mem) { words = (unsigned long long *)mem; }
    __device__ __forceinline__ void post(size_t tile, unsigned epoch, T v) const
    {
        const unsigned long long tag = (unsigned long long)((epoch << 2) | kPartial) << 32;
        if constexpr (W == 1) {
            unsigned bits = 0;
            memcpy(&bits, &v, sizeof(T));
            st_relaxed_u64(words + tile, tag | bits);
        } else {
            unsigned long long bits;
            memcpy(&bits, &v, sizeof(T));
            st_relaxed_u64(words + 2 * tile, tag | (unsigned)bits);
            st_relaxed_u64(words + 2 * tile + 1, tag | (unsigned)(bits >> 32));
        }
    }
    // true once the aggregate of this epoch is there
    __device__ __forceinline__ bool peek(size_t tile, unsigned epoch, T &v) const
    {
        if constexpr (W == 1) {
            const unsigned long long w = ld_relaxed_u64(words + tile);
            const unsigned bits = (unsigned)w;
            memcpy(&v, &bits, sizeof(T));
            return (unsigned)(w >> 34) == epoch;
        } else {
            const unsigned long long lo = ld_relaxed_u64(words + 2 * tile), hi = ld_relaxed_u64(words + 2 * tile + 1);
            const unsigned long long bits = (hi << 32) | (unsigned)lo;
            memcpy(&v, &bits, sizeof(T));
            return (unsigned)(lo >> 34) == epoch && (unsigned)(hi >> 34) == epoch;
        }
    }
};

// MODE 0: the scan.  Diagnostics (bench/scan_lab.cu only): 1 = pipeline without the inter-CTA chain (every prefix is
// the identity: wrong results, shows what the chain costs), 2 = bulk copy in / bulk copy out only.
template <typename T, int OP, int NV, int S, int D, int MODE = 0>
__global__ void __launch_bounds__(kSwThreads, 1)
scan_ws_kernel(const T *in, T *out, size_t n, int exclusive, T init, WsTileState<T> ts, unsigned epoch, size_t num_tiles,
               const T *__restrict__ init_dev = nullptr)
{
    if (init_dev) init = *init_dev;  // the seed lives on the device (multi-GPU scan: the carry of the ranks before this one)
    typedef Op<OP, T> O;
    typedef ScanWsShape<T, NV, S> C;
    constexpr int VEC = C::VEC, TILE = C::TILE;
    static_assert(D >= 1 && D < S - 1, "R runs D tiles ahead of C, and the loads need at least one more stage");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *stage_base = reinterpret_cast<T *>(smem_raw);
    T *wtot = reinterpret_cast<T *>(smem_raw + (size_t)S * C::TILE_BYTES);  // [S][16] warp totals (R)
    T *woff = wtot + S * kSwComputeWarps;                                   // [S][16] tile prefix + warps before (prefix warp)
    unsigned long long *full_bar = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)S * C::TILE_BYTES + C::SIDE_BYTES);
    unsigned long long *red_bar = full_bar + S, *pfx_bar = red_bar + S, *done_bar = pfx_bar + S;

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned G = gridDim.x, b = blockIdx.x;
    const unsigned cnt = b < num_tiles ? (unsigned)((num_tiles - 1 - b) / G + 1) : 0u;  // tiles of this CTA: i * G + b
    auto tile_start = [&](unsigned i) { return ((size_t)i * G + b) * (size_t)TILE; };

    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&red_bar[s], kSwComputeWarps);
            mbar_init(&pfx_bar[s], 1);
            mbar_init(&done_bar[s], kSwComputeWarps);
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp < kSwComputeWarps) {
        // ============================ compute warps ============================
        const unsigned seg = warp * C::WARP_ELEMS + lane * VEC;  // first element of this lane's vector 0 inside a tile
        for (unsigned it = 0; it < cnt + D; ++it) {
            if (it < cnt) {
                // ---- R: fold this warp's part of tile `it` ----
                const int s = (int)(it % S);
                const size_t base = tile_start(it);
                T *st = stage_base + (size_t)s * TILE + seg;
                T acc = O::identity();
                if (base + TILE <= n) {
                    mbar_wait(&full_bar[s], (it / S) & 1u);
                    if (MODE != 2) {
#pragma unroll
                        for (int j = 0; j < NV; j++) {
                            const uint4 v = *reinterpret_cast<const uint4 *>(st + j * 32 * VEC);
                            const T *e = reinterpret_cast<const T *>(&v);
#pragma unroll
                            for (int k = 0; k < VEC; k++) acc = O::apply(acc, e[k]);
                        }
                    }
                } else {  // the ragged last tile (never bulk-copied): guarded loads, here and again in C
#pragma unroll
                    for (int j = 0; j < NV; j++) {
#pragma unroll
                        for (int k = 0; k < VEC; k++) {
                            const size_t i = base + seg + (size_t)j * 32 * VEC + k;
                            if (i < n) acc = O::apply(acc, in[i]);
                        }
                    }
                }
                if (MODE != 2) acc = warp_reduce<T, OP>(acc);
                if (lane == 0) {
                    wtot[s * kSwComputeWarps + warp] = acc;
                    mbar_arrive(&red_bar[s]);
                }
            }
            if (it >= (unsigned)D) {
                // ---- C: scan this warp's part of tile `it - D` ----
                const unsigned i = it - D;
                const int s = (int)(i % S);
                const size_t base = tile_start(i);
                const bool full = base + TILE <= n;
                T *st = stage_base + (size_t)s * TILE + seg;
                mbar_wait(&pfx_bar[s], (i / S) & 1u);
                if (MODE != 2) {
                    T carry = woff[s * kSwComputeWarps + warp];
#pragma unroll
                    for (int j = 0; j < NV; j++) {
                        uint4 raw;
                        T *x = reinterpret_cast<T *>(&raw);
                        if (full) {
                            raw = *reinterpret_cast<const uint4 *>(st + j * 32 * VEC);
                        } else {
#pragma unroll
                            for (int k = 0; k < VEC; k++) {
                                const size_t g = base + seg + (size_t)j * 32 * VEC + k;
                                x[k] = g < n ? in[g] : O::identity();
                            }
                        }
#pragma unroll
                        for (int k = 1; k < VEC; k++) x[k] = O::apply(x[k - 1], x[k]);
                        T sc = x[VEC - 1];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const T o = shfl_up_t(sc, d);
                            if ((int)lane >= d) sc = O::apply(o, sc);
                        }
                        T e = shfl_up_t(sc, 1);
                        if (lane == 0) e = O::identity();
                        const T p = O::apply(carry, e);
                        carry = O::apply(carry, shfl_t(sc, 31));
                        T y[VEC];
                        if (exclusive == 1) {
                            y[0] = p;
#pragma unroll
                            for (int k = 1; k < VEC; k++) y[k] = O::apply(p, x[k - 1]);
                        } else {
#pragma unroll
                            for (int k = 0; k < VEC; k++) y[k] = O::apply(p, x[k]);
                        }
                        if (full) {
                            *reinterpret_cast<uint4 *>(st + j * 32 * VEC) = *reinterpret_cast<const uint4 *>(y);
                        } else {
#pragma unroll
                            for (int k = 0; k < VEC; k++) {
                                const size_t g = base + seg + (size_t)j * 32 * VEC + k;
                                if (g < n) out[g] = y[k];
                            }
                        }
                    }
                    fence_proxy_async();  // the stage is read by the bulk-copy engine next
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&done_bar[s]);
            }
        }
    } else if (warp == kSwComputeWarps) {
        // ============================ producer: the bulk-copy engine ============================
        if (lane == 0) {
            auto issue_load = [&](unsigned i) {
                if (i < cnt) {
                    const size_t base = tile_start(i);
                    if (base + TILE <= n) {
                        const int s = (int)(i % S);
                        mbar_expect_tx(&full_bar[s], (unsigned)C::TILE_BYTES);
                        tma_load_1d(stage_base + (size_t)s * TILE, in + base, (unsigned)C::TILE_BYTES, &full_bar[s]);
                    }
                }
            };
            for (unsigned i = 0; i < (unsigned)S; i++) issue_load(i);
            for (unsigned i = 0; i < cnt; i++) {
                const int s = (int)(i % S);
                mbar_wait(&done_bar[s], (i / S) & 1u);
                const size_t base = tile_start(i);
                if (base + TILE <= n) tma_store_1d(out + base, stage_base + (size_t)s * TILE, (unsigned)C::TILE_BYTES);
                if (i >= 1) {  // the store issued one tile ago has read its stage: refill it
                    tma_store_wait_read<1>();
                    issue_load(i - 1 + S);
                }
            }
            tma_store_wait_read<0>();  // shared memory must outlive the last bulk stores
        }
    } else if (warp == kSwComputeWarps + 1) {
        // ============================ aggregate warp: tile aggregate -> descriptor ============================
        if (MODE == 0) {
            for (unsigned i = 0; i < cnt; i++) {
                const int s = (int)(i % S);
                mbar_wait(&red_bar[s], (i / S) & 1u);
                T v = lane < (unsigned)kSwComputeWarps ? wtot[s * kSwComputeWarps + lane] : O::identity();
#pragma unroll
                for (int d = 1; d < kSwComputeWarps; d <<= 1) {
                    const T o = shfl_up_t(v, d);
                    if ((int)lane >= d) v = O::apply(o, v);
                }
                if (lane == kSwComputeWarps - 1) ts.post((size_t)i * G + b, epoch, v);
            }
        }
    } else if (warp == kSwComputeWarps + 2) {
        // ============================ prefix warp: aggregates of the round -> warp offsets ============================
        // Round i of all CTAs is the contiguous window of tiles [i*G, i*G + G).  Every CTA polls ALL aggregates of the
        // round (K per lane), scans them with a fixed-shape scan and so learns both its own tile's prefix inside the
        // round and the round total; the carry across rounds stays in this warp's registers.  One L2 hop between a
        // tile's aggregate and every prefix that depends on it, no serial chain through memory, and all CTAs compute
        // bit-identical carries (same operations in the same order), so floating-point results are deterministic.
        constexpr int K = kSwPollPerLane;
        T carry = exclusive ? init : O::identity();  // modes 1 (exclusive) and 2 (seeded inclusive) start from init
        // first poll of a round's descriptors: issued one round early, so that its L2 latency overlaps the scan of the
        // round before (R runs ahead: in the steady state the aggregates are there long before they are needed)
        T vn[K];
        bool okn[K];
        auto first_poll = [&](unsigned i) {
            const size_t t0 = (size_t)i * G;
            if (i >= cnt) return;  // (nothing will look at vn / okn again)
            const unsigned m = (unsigned)(num_tiles - t0 < (size_t)G ? num_tiles - t0 : (size_t)G);
            // all K loads are issued before the first result is looked at (a lane without a descriptor re-reads t0)
            bool there[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                okn[k] = lane * K + k >= m;
                there[k] = ts.peek(okn[k] ? t0 : t0 + lane * K + k, epoch, vn[k]);
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (okn[k]) vn[k] = O::identity();
                else okn[k] = there[k];
            }
        };
        if (MODE == 0) first_poll(0);
        for (unsigned i = 0; i < cnt; i++) {
            const int s = (int)(i % S);
            T p = O::identity();
            if (MODE == 0) {
                const size_t t0 = (size_t)i * G;
                T v[K];
                bool ok[K];
                bool all = true;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    v[k] = vn[k];
                    ok[k] = okn[k];
                    all &= ok[k];
                }
                while (!__all_sync(0xffffffffu, all)) {
                    __nanosleep(kSpinBackoffNs);
                    all = true;
                    T got[K];
                    bool there[K];
#pragma unroll
                    for (int k = 0; k < K; k++) there[k] = ts.peek(ok[k] ? t0 : t0 + lane * K + k, epoch, got[k]);
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        if (!ok[k]) {
                            if (there[k]) {
                                v[k] = got[k];
                                ok[k] = true;
                            } else {
                                all = false;
                            }
                        }
                    }
                }
                first_poll(i + 1);
                // lane-local exclusive scan of the lane's K consecutive aggregates, warp scan of the lane totals
                T ex[K];
                T run = O::identity();
#pragma unroll
                for (int k = 0; k < K; k++) {
                    ex[k] = run;
                    run = O::apply(run, v[k]);
                }
                T sc = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const T o = shfl_up_t(sc, d);
                    if ((int)lane >= d) sc = O::apply(o, sc);
                }
                T le = shfl_up_t(sc, 1);
                if (lane == 0) le = O::identity();
                T mine = ex[0];
#pragma unroll
                for (int k = 1; k < K; k++)
                    if ((int)(b % K) == k) mine = ex[k];
                p = O::apply(carry, shfl_t(O::apply(le, mine), (int)(b / K)));
                carry = O::apply(carry, shfl_t(sc, 31));
            }
            mbar_wait(&red_bar[s], (i / S) & 1u);
            T v = lane < (unsigned)kSwComputeWarps ? wtot[s * kSwComputeWarps + lane] : O::identity();
#pragma unroll
            for (int d = 1; d < kSwComputeWarps; d <<= 1) {
                const T o = shfl_up_t(v, d);
                if ((int)lane >= d) v = O::apply(o, v);
            }
            T e = shfl_up_t(v, 1);
            if (lane == 0) e = O::identity();
            if (lane < (unsigned)kSwComputeWarps) woff[s * kSwComputeWarps + lane] = O::apply(p, e);
            __syncwarp();
            if (lane == 0) mbar_arrive(&pfx_bar[s]);
        }
    }
}

}  // namespace bcb
