// radix_pass_ws.cu -- warp-specialised onesweep pass for large keys-only sorts of 32- and 64-bit keys on sm_100a.
//
// Replaces one (count + scan + scatter) round of the reference's radix_sort_impl (algorithm/detail/radix_sort.hpp:
// 186-250, 347-425) with ONE launch per 8-bit digit.  What shaped it (bench/tma_scatter.cu, bench/tma_issue.cu,
// measured on B200):
//   * the scatter phase of a radix pass is bound by the memory system, not by the SM: a write whose ends do not fall on
//     32-byte sector boundaries costs far more than its bytes.  256 runs of 48 keys per tile (12288-key tiles, r01)
//     move 2 x 4.3 GB in 3.0 ms, runs of 96 keys in 2.0 ms, runs of 192 keys in 1.8 ms.  So: ONE tile buffer as large
//     as shared memory allows (49152 u32 keys = 192 KB), and every digit run leaves the SM as one cp.async.bulk
//     shared->global copy of its 16-byte aligned body (only the <= 3 + 3 edge keys of a run are stored by the LSU);
//   * a bulk copy needs 16-byte aligned addresses on both sides, so the digit-sorted tile is laid out in shared memory
//     with every run shifted to the alignment of its destination (3 pad keys per digit at most).  The global position
//     must therefore be known BEFORE the keys are scattered into shared memory: the decoupled look-back runs between
//     the counting sweep and the scatter sweep of a tile, in helper warps, while the workers scatter the previous tile;
//   * issuing a small bulk copy costs a warp ~70-130 cycles (per-lane ELECT / R2UR loop), so the copies are issued by
//     8 helper warps (one digit run per thread), never by the workers;
//   * the counting sweep reads the keys straight from global memory (128-bit loads, nothing kept), the scatter sweep
//     reads them again one tile later (an L2 hit: ~40 MB pass through the 126 MB L2 in between) -- no landing buffer,
//     no registers held across the look-back;
//   * ranking is the two-sweep scheme of radix_sort.cu (sweep 1: shared-memory atomicAdd per (warp, digit), sweep 2: the
//     same atomics in the same order return the position), with 16-bit counters packed two per word.  Like
//     kRankTwoSweep it relies on same-address shared atomics of one warp instruction being applied in lane order;
//     sort_typed verifies the result (verify_sorted_kernel) and falls back on the device if that ever fails.
// One CTA per SM: 24 worker warps + 8 helper warps.  Tiles are drawn from a ticket (atomic counter) so forward
// progress never depends on which CTAs are resident.
//   workers:  count(t0) | count(t1) scatter(t0) | count(t2) scatter(t1) | ...
//   helpers:            | offsets+look-back(t0) | offsets(t1) store(t0) | offsets(t2) store(t1) | ...
#include "radix_common.cuh"
#include "tma.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace bcb {

// -DBCB_WS_PROFILE: per-CTA cycle counters of the phases (thread 0 of each role), printed by the launcher
#ifdef BCB_WS_PROFILE
#define WS_PROF_DECL unsigned long long prof_t0 = 0, prof_acc[8] = {}
#define WS_PROF_BEGIN() prof_t0 = clock64()
#define WS_PROF_END(k) do { const unsigned long long now_ = clock64(); prof_acc[k] += now_ - prof_t0; prof_t0 = now_; } while (0)
#define WS_PROF_DUMP(base) do { for (int q = 0; q < 8; q++) g_ws_prof[(size_t)blockIdx.x * 16 + (base) + q] = prof_acc[q]; } while (0)
__device__ unsigned long long g_ws_prof[148 * 16 * 4];
#else
#define WS_PROF_DECL
#define WS_PROF_BEGIN()
#define WS_PROF_END(k)
#define WS_PROF_DUMP(base)
#endif

constexpr int kWsWorkerWarps = 24;
constexpr int kWsWorkers = kWsWorkerWarps * 32;  // 768
constexpr int kWsHelpers = kRadixSize;           // one per digit value
constexpr int kWsThreads = kWsWorkers + kWsHelpers;
constexpr unsigned kWsNoTile = 0xffffffffu;
// named barriers (0 is __syncthreads); p = parity of the tile's sequence number inside the CTA
enum { kBarCounted = 1 /* +p */, kBarOffsets = 3 /* +p */, kBarScattered = 5, kBarDrained = 6, kBarHelpers = 7, kBarWorkers = 8 };

template <int VB> struct ws_value { typedef unsigned type; };
template <> struct ws_value<8> { typedef unsigned long long type; };

// K: key type, VB: payload bytes (0, 4, 8), DET: deterministic ranking (peer masks by atomicOr; needs a mask table)
template <typename K, int VB, bool DET> struct WsShape {
    static constexpr int KB = (int)sizeof(K);
    static constexpr int MINB = VB ? (VB < KB ? VB : KB) : KB;
    static constexpr int A = 16 / MINB;  // elements per 16-byte chunk of the NARROWER array: run starts aligned to A elements are 16-byte aligned in both
    // elements per worker thread: what fits next to the tables
    static constexpr int ITEMS = VB == 0 ? (DET ? 192 / KB : 224 / KB) : (KB + VB == 8 ? 24 : 16);
    static constexpr int CHUNK = VB == 0 ? (DET ? 48 : 56) / KB : (VB == 4 ? 8 : 4);  // elements a worker holds in registers at a time (+ as many in flight)
    static constexpr int SEG = ITEMS * 32;                 // elements per worker warp (one contiguous segment)
    static constexpr int TILE = kWsWorkers * ITEMS;        // 43008 u32 keys / 21504 u64 keys / 18432 u32+u32 pairs / 12288 u32+u64 pairs
    static constexpr int PAD = (A - 1) * kRadixSize;       // alignment shifts: run d starts (A-1)*d + [0, A) later
    static constexpr size_t KBUF_BYTES = (size_t)(TILE + PAD) * KB;
    static constexpr size_t VBUF_BYTES = (size_t)(TILE + PAD) * VB;
    static constexpr size_t TAB_BYTES = 2 * (size_t)kWsWorkerWarps * kRadixSize * sizeof(unsigned);  // [2][24][256]
    static constexpr size_t MASK_BYTES = DET ? (size_t)kWsWorkerWarps * kRadixSize * sizeof(unsigned) : 0;  // [24][256]
    static constexpr size_t MISC_BYTES = 256;
    static constexpr size_t SMEM_BYTES = KBUF_BYTES + VBUF_BYTES + TAB_BYTES + MASK_BYTES + MISC_BYTES;
    static_assert(KBUF_BYTES % 128 == 0 && VBUF_BYTES % 128 == 0, "the buffers and tables stay aligned");
    static_assert(ITEMS % CHUNK == 0, "whole chunks");
    static_assert(SMEM_BYTES <= 232448, "one CTA per SM: 227 KB of shared memory");
};

template <typename K, int VB, bool DET, int XF, int LB>
__global__ void __launch_bounds__(kWsThreads, 1)
onesweep_ws(const K *__restrict__ keys_in, K *__restrict__ keys_out, const void *__restrict__ vals_in_v, void *__restrict__ vals_out_v,
            const unsigned *__restrict__ digit_base, unsigned long long *lookback, unsigned epoch, size_t n, unsigned num_tiles, int shift,
            const __grid_constant__ Transform tf, unsigned long long *ticket, unsigned long long ticket_base, int flags,
            const unsigned long long *__restrict__ dst_tab, const uint4 *__restrict__ tile_tab, const unsigned *__restrict__ hot)
{
    typedef WsShape<K, VB, DET> C;
    typedef typename ws_value<VB>::type V;
    static_assert(VB == 0 || DET, "a payload needs the deterministic ranking (its stability cannot be verified afterwards)");
    const V *vals_in = reinterpret_cast<const V *>(vals_in_v);
    V *vals_out = reinterpret_cast<V *>(vals_out_v);
    constexpr int ITEMS = C::ITEMS, TILE = C::TILE, A = C::A, SEG = C::SEG;
    constexpr int ROW = kRadixSize;  // words per (warp, tile) table row.  (16-bit counters packed two per word would let
                                     // the tile grow to 49152 keys, but measured 4.0 instead of 3.35 wavefronts per atomic:
                                     // twice as many lanes share a word -- and five more ALU instructions per key)
    extern __shared__ __align__(128) unsigned char smem[];
    K *buf = reinterpret_cast<K *>(smem);
    V *vbuf = reinterpret_cast<V *>(smem + C::KBUF_BYTES);                                      // payload, same positions as the keys
    unsigned *tab = reinterpret_cast<unsigned *>(smem + C::KBUF_BYTES + C::VBUF_BYTES);         // [2][24][ROW]
    unsigned *masks = tab + 2 * kWsWorkerWarps * ROW;                                           // [24][ROW] (DET only)
    volatile unsigned *ring = reinterpret_cast<volatile unsigned *>(smem + C::KBUF_BYTES + C::VBUF_BYTES + C::TAB_BYTES + C::MASK_BYTES);  // [4] tile ids
    unsigned *hscan = const_cast<unsigned *>(ring) + 4;                                        // [2][8] helper warp sums
    // [4] tile descriptors {first key, end (one past the last key), first tile of the tile's segment, segment}.  A plain
    // sort is one segment: tile t covers [t * TILE, min((t + 1) * TILE, n)).  The segmented passes of the multi-GPU sort
    // (bcb_radix_sort_segments) bring a tile table: tiles never straddle two segments, the look-back stops at the
    // segment's first tile, and digit_base is indexed [segment][digit].
    volatile unsigned *ring_d = ring + 20;
    const unsigned tid = threadIdx.x;
    auto tile_start = [&](unsigned i) -> unsigned { return ring_d[(i & 3u) * 4 + 0]; };
    auto tile_end = [&](unsigned i) -> unsigned { return ring_d[(i & 3u) * 4 + 1]; };

    for (unsigned i = tid; i < (DET ? 3 : 2) * kWsWorkerWarps * ROW; i += kWsThreads) tab[i] = 0;  // (the mask table follows the count tables)
    __syncthreads();

    if (hot && __ldg(hot + 1) == kConstDigit) {
        // Every key has the same value of this digit (digit_scan found one bin holding all of them): the stable pass is the
        // identity permutation.  The CTAs stream their tiles from keys_in to keys_out (through the pass's key transform, so
        // that the sortable form between the passes stays what the neighbouring passes expect); no ranking, no look-back.
        typedef typename key_traits<K>::U U;
        if (blockIdx.x == 0 && tid == 0) atomicAdd(ticket, (unsigned long long)num_tiles + gridDim.x);  // (the host's reservation)
        const size_t g = (size_t)blockIdx.x * kWsThreads + tid, nthreads = (size_t)gridDim.x * kWsThreads;
        stream_copy(keys_in, keys_out, n, g, nthreads, [&](K raw) -> K {
            const U srt = (XF == kXfIn || XF == kXfBoth) ? transform_fwd<K>(raw, tf) : (U)raw;
            return XF == kXfOut ? transform_inv<K>(srt, tf) : (XF == kXfIn ? (K)srt : raw);
        });
        if constexpr (VB > 0) stream_copy(vals_in, vals_out, n, g, nthreads, [](V v) { return v; });
        return;
    }

    if (tid < kWsWorkers) {
        // ======================= workers: counting sweep, scatter sweep =======================
        const unsigned w = tid >> 5, lane = tid & 31u;
        const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
        typedef typename key_traits<K>::U U;
        // keys travel between the passes in sortable form: transformed once by the first pass (kXfIn), turned back into
        // the original bit pattern by the last one (kXfOut); in between the digit is a plain bit field
        auto sortable = [&](K raw) -> U { return (XF == kXfIn || XF == kXfBoth) ? transform_fwd<K>(raw, tf) : (U)raw; };
        auto digit = [&](U t) -> unsigned { return (unsigned)(t >> shift) & (kRadixSize - 1); };
        auto stored = [&](K raw, U t) -> K { return XF == kXfOut ? transform_inv<K>(t, tf) : (XF == kXfIn ? (K)t : raw); };
        // Digit values that hold a large share of the keys (constant high bytes of small integers, the exponent byte of
        // floats): 32 lanes adding to ONE shared-memory counter serialise, a pass over such a digit took 2-8x as long.
        // digit_scan names up to two such values; their lanes are ranked by ballot (one atomic per warp instruction and
        // value, by the first of them), which is also the lane order the speculative ranking assumes anyway.
        const unsigned h0 = hot ? __ldg(hot) : kNoHotDigit, h1 = hot ? __ldg(hot + 1) : kNoHotDigit;
        const bool hot_mode = h0 != kNoHotDigit;
        // hot lanes touch no shared memory at all: their counts / positions are kept in two (warp-uniform) registers per
        // sweep -- the count of the hot value is added to the table once per tile, the positions are start-of-run +
        // keys of that value seen so far + lanes below in the ballot
        auto count_one = [&](unsigned *row, unsigned d, unsigned &acc0, unsigned &acc1, auto hot_tag) {
            if constexpr (!decltype(hot_tag)::value) {
                atomicAdd(&row[d], 1u);
            } else {
                const bool a = d == h0, b = d == h1;
                acc0 += __popc(__ballot_sync(0xffffffffu, a));
                acc1 += __popc(__ballot_sync(0xffffffffu, b));
                if (!(a || b)) atomicAdd(&row[d], 1u);
            }
        };
        auto rank_one = [&](unsigned *row, unsigned d, unsigned &run0, unsigned &run1, auto hot_tag) -> unsigned {
            if constexpr (!decltype(hot_tag)::value) return atomicAdd(&row[d], 1u);  // start of the (warp, digit) run + rank: the same adds as count_one
            const bool a = d == h0, b = d == h1;
            const unsigned ma = __ballot_sync(0xffffffffu, a), mb = __ballot_sync(0xffffffffu, b);
            unsigned pos = (a ? run0 : run1) + __popc((a ? ma : mb) & lanemask_lt());
            if (!(a || b)) pos = atomicAdd(&row[d], 1u);
            run0 += __popc(ma);
            run1 += __popc(mb);
            return pos;
        };
        WS_PROF_DECL;
        auto draw = [&](unsigned i) {  // tile id of the CTA's i-th tile -> ring[i & 3], visible after the workers' barrier
            if (tid == 0) {
                // (Spreading the CTAs' first draws over one tile period -- so that the SMs do not all read their next tile
                // from HBM at the same moment -- was measured: 9.64 against 9.69 ms for the four passes of 2^30 keys, noise.)
                const unsigned long long t = atomicAdd(ticket, 1ull) - ticket_base;
                ring[i & 3u] = t < num_tiles ? (unsigned)t : kWsNoTile;
                if (t < num_tiles) {
                    uint4 desc;
                    if (tile_tab) {
                        desc = __ldg(tile_tab + t);
                    } else {
                        const unsigned long long s0 = t * TILE, e0 = s0 + TILE;
                        desc = make_uint4((unsigned)s0, (unsigned)(e0 < n ? e0 : n), 0u, 0u);
                    }
                    volatile unsigned *rd = ring_d + (i & 3u) * 4;
                    rd[0] = desc.x; rd[1] = desc.y; rd[2] = desc.z; rd[3] = desc.w;
                }
                // (An experiment that pulled the tile one round ahead into L2 with cp.async.bulk.prefetch was measured
                // 5 % SLOWER: the pass is bound by the scattered writes, whose drain the prefetch only delays.)
            }
            named_bar_sync(kBarWorkers, kWsWorkers);
            return ring[i & 3u];
        };
        auto count = [&](unsigned p, unsigned i, auto hot_tag) {
            const unsigned t_first = tile_start(i), t_end = tile_end(i);
            const size_t base = (size_t)t_first + (size_t)w * SEG;
            unsigned *row = tab + (p * kWsWorkerWarps + w) * ROW;
            unsigned acc0 = 0, acc1 = 0;  // keys of the hot digit values in this warp's segment
            if (t_end - t_first == (unsigned)TILE) {
                // any order inside the warp's segment: 128-bit loads, a window of WIN per lane in flight.  The lines are
                // asked to STAY in L2 (evict_last): the scatter sweep reads them again one tile later.
                const uint4 *src = reinterpret_cast<const uint4 *>(keys_in + base) + lane;
                constexpr int VECK = 16 / (int)sizeof(K);                                     // keys per 128-bit vector
                constexpr int NV = SEG / VECK / 32;                                           // vectors per lane
                constexpr int WIN0 = NV % 7 == 0 ? 7 : (NV % 6 == 0 ? 6 : (NV % 4 == 0 ? 4 : 1));  // vectors in flight per lane
                constexpr int WIN = (decltype(hot_tag)::value && WIN0 > 4) ? (NV % 4 == 0 ? 4 : (NV % 3 == 0 ? 3 : 2)) : WIN0;  // (the ballot ranking needs registers)
                uint4 v[WIN];
#pragma unroll
                for (int u = 0; u < WIN; u++) v[u] = ld_hint_v4(src + u * 32, keep);
#pragma unroll
                for (int j = 0; j < NV; j++) {
                    const K *e = reinterpret_cast<const K *>(&v[j % WIN]);
#pragma unroll
                    for (int c = 0; c < VECK; c++) count_one(row, digit(sortable(e[c])), acc0, acc1, hot_tag);
                    if (j + WIN < NV) v[j % WIN] = ld_hint_v4(src + (j + WIN) * 32, keep);
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < ITEMS; i++) {
                    const size_t idx = base + (size_t)i * 32 + lane;
                    // padding counts as digit 255: sorts last
                    count_one(row, idx < t_end ? digit(sortable(__ldg(keys_in + idx))) : (unsigned)(kRadixSize - 1), acc0, acc1, hot_tag);
                }
            }
            if constexpr (decltype(hot_tag)::value) {
                if (lane == 0 && acc0) atomicAdd(&row[h0], acc0);
                if (lane == 1 && acc1) atomicAdd(&row[h1], acc1);  // (acc1 stays 0 when there is no second hot value)
            }
            named_bar_arrive(kBarCounted + p, kWsThreads);
        };
        auto scatter_tile = [&](unsigned p, unsigned i_tile, bool first, auto full_tag, auto hot_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const size_t n = tile_end(i_tile);  // (shadows the range length: the bound of THIS tile)
            const size_t base = (size_t)tile_start(i_tile) + (size_t)w * SEG + lane;
            unsigned *row = tab + (p * kWsWorkerWarps + w) * ROW;
            constexpr int CHUNK = (decltype(hot_tag)::value && C::CHUNK % 2 == 0) ? C::CHUNK / 2 : C::CHUNK;  // (the ballot ranking needs registers)
            K cur[CHUNK], nxt[CHUNK];
            V vcur[VB ? CHUNK : 1], vnxt[VB ? CHUNK : 1];
            auto load = [&](K (&dst)[CHUNK], V (&vdst)[VB ? CHUNK : 1], int c) {
#pragma unroll
                for (int i = 0; i < CHUNK; i++) {
                    const size_t idx = base + (size_t)(c * CHUNK + i) * 32;
                    dst[i] = (FULL || idx < n) ? ld_hint(keys_in + idx, drop) : (K)0;  // last use: first in line for eviction
                    if constexpr (VB > 0) vdst[i] = (FULL || idx < n) ? __ldg(vals_in + idx) : V();
                }
            };
            load(cur, vcur, 0);  // second read of the tile (the counting sweep was the first): flies during the waits below
            WS_PROF_BEGIN();
            named_bar_sync(kBarOffsets + p, kWsThreads);          // the tables hold the start of every (warp, digit) run
            WS_PROF_END(2);
            if (!first) named_bar_sync(kBarDrained, kWsThreads);  // the previous tile's bulk copies have read the buffer
            WS_PROF_END(3);
            [[maybe_unused]] unsigned *mrow = masks + w * ROW;  // (deterministic ranking only)
            unsigned run0 = 0, run1 = 0;  // next position of the hot digit values in this warp's runs
            if constexpr (decltype(hot_tag)::value) {
                run0 = row[h0];
                run1 = row[h1 != kNoHotDigit ? h1 : h0];
            }
#pragma unroll 1
            for (int c = 0; c < ITEMS / CHUNK; c++) {
                if (c + 1 < ITEMS / CHUNK) load(nxt, vnxt, c + 1);
#pragma unroll
                for (int i = 0; i < CHUNK; i++) {
                    const U t = sortable(cur[i]);
                    unsigned d = digit(t);
                    if (!FULL && base + (size_t)(c * CHUNK + i) * 32 >= n) d = kRadixSize - 1;
                    unsigned pos;
                    if constexpr (DET) {
                        // deterministic by construction: the lanes OR their bit into the mask of their digit; after a warp
                        // barrier the mask holds the complete peer set whatever order the atomics were applied in; the
                        // highest peer clears it and advances the run's cursor.  (Lanes of a hot digit value take no part:
                        // their peer set is a ballot and their cursor a register.)
                        bool cold = true;
                        unsigned hot_pos = 0;
                        if constexpr (decltype(hot_tag)::value) {
                            const bool a = d == h0, b = d == h1;
                            const unsigned ma = __ballot_sync(0xffffffffu, a), mb = __ballot_sync(0xffffffffu, b);
                            hot_pos = (a ? run0 : run1) + __popc((a ? ma : mb) & lanemask_lt());
                            run0 += __popc(ma);
                            run1 += __popc(mb);
                            cold = !(a || b);
                        }
                        if (cold) atomicOr(&mrow[d], 1u << lane);
                        __syncwarp();
                        const unsigned m = cold ? mrow[d] : 0u, o = cold ? row[d] : 0u;
                        __syncwarp();
                        if (cold && (m >> lane) == 1u) {
                            mrow[d] = 0;
                            row[d] = o + __popc(m);
                        }
                        __syncwarp();
                        pos = cold ? o + __popc(m & lanemask_lt()) : hot_pos;
                    } else {
                        // same atomics, same order as count(): start of the run + rank.  (Issuing all atomics of the chunk
                        // before the first store was measured 3 % slower: the LSU queue, not the latency, is the limit.)
                        pos = rank_one(row, d, run0, run1, hot_tag);
                    }
                    buf[pos] = stored(cur[i], t);
                    if constexpr (VB > 0) vbuf[pos] = vcur[i];
                }
#pragma unroll
                for (int i = 0; i < CHUNK; i++) {
                    cur[i] = nxt[i];
                    if constexpr (VB > 0) vcur[i] = vnxt[i];
                }
            }
            __syncwarp();
            {  // the row is this warp's own: zero it for the warp's next tile
                uint4 *z = reinterpret_cast<uint4 *>(row);
                z[lane] = make_uint4(0, 0, 0, 0);
                z[lane + 32] = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async();
            named_bar_arrive(kBarScattered, kWsThreads);
            WS_PROF_END(4);
        };
        auto scatter = [&](unsigned p, unsigned i_tile, bool first, auto hot_tag) {
            if (tile_end(i_tile) - tile_start(i_tile) == (unsigned)TILE) scatter_tile(p, i_tile, first, std::true_type(), hot_tag);
            else scatter_tile(p, i_tile, first, std::false_type(), hot_tag);
        };
        // (two copies of the loop, chosen once: the ballot ranking costs nothing -- not even registers -- where no digit
        // value is hot)
        auto work = [&](auto hot_tag) {
            unsigned t_cur = draw(0);
            if (t_cur != kWsNoTile) count(0, 0, hot_tag);
            else named_bar_arrive(kBarCounted + 0, kWsThreads);
            for (unsigned i = 0; t_cur != kWsNoTile; ++i) {
                WS_PROF_BEGIN();
                const unsigned t_next = draw(i + 1);
                WS_PROF_END(0);
                if (t_next != kWsNoTile) count((i + 1) & 1u, i + 1, hot_tag);
                else named_bar_arrive(kBarCounted + ((i + 1) & 1u), kWsThreads);  // the helpers learn from the ring that it is void
                WS_PROF_END(1);
                scatter(i & 1u, i, i == 0, hot_tag);
                t_cur = t_next;
            }
        };
        if (hot_mode) work(std::true_type());
        else work(std::false_type());
        if (tid == 0) WS_PROF_DUMP(0);
    } else {
        // ======================= helpers: digit scan, look-back, bulk stores (one digit value per thread) =======================
        const unsigned d = tid - kWsWorkers, lane = d & 31u, hw = d >> 5;
        // dst_tab (multi-GPU exchange pass, bcb_radix_sort_exchange): every digit value has its own destination array --
        // [d] keys, [256 + d] values, 16-byte aligned, possibly another GPU's memory -- and digit_base[d] counts from there
        K *const kout_d = dst_tab ? reinterpret_cast<K *>(__ldg(dst_tab + d)) : keys_out;
        V *const vout_d = (VB > 0 && dst_tab) ? reinterpret_cast<V *>(__ldg(dst_tab + kRadixSize + d)) : vals_out;
        const unsigned long long drop = l2_policy_evict_first();  // written lines are not read again in this pass
        const unsigned long long tag_p = (unsigned long long)((epoch << 2) | kLbPartial) << 32;
        const unsigned long long tag_i = (unsigned long long)((epoch << 2) | kLbInclusive) << 32;
        struct Run { unsigned g, s, c; };  // first index in keys_out, first index in the tile buffer, length
        WS_PROF_DECL;
        struct Scan { unsigned pub, e, gbase, first_tile; };
        // part 1: digit counts of the tile, exclusive scan over the digit values, publish the tile's count of this digit
        auto scan_publish = [&](unsigned p, unsigned i_tile, unsigned t, Scan &sc) {
            WS_PROF_BEGIN();
            const unsigned t_first = tile_start(i_tile), t_end = tile_end(i_tile);
            sc.first_tile = ring_d[(i_tile & 3u) * 4 + 2];
            sc.gbase = __ldg(digit_base + (size_t)ring_d[(i_tile & 3u) * 4 + 3] * kRadixSize + d);  // (in flight during the scan)
            unsigned *col = tab + p * kWsWorkerWarps * ROW + d;
            unsigned count = 0;
#pragma unroll
            for (int w = 0; w < kWsWorkerWarps; w++) count += col[w * ROW];
            unsigned pub = count;  // without the padding of a partial tile (counted as digit 255)
            if (d == kRadixSize - 1) pub -= (unsigned)TILE - (t_end - t_first);
            // exclusive scan over the 256 digit values -> start of each run in the digit-sorted tile
            unsigned incl = count;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += o;
            }
            if (lane == 31) hscan[p * 8 + hw] = incl;
            named_bar_sync(kBarHelpers, kWsHelpers);
            unsigned add = 0;
#pragma unroll
            for (int j = 0; j < 7; j++) add += (j < (int)hw) ? hscan[p * 8 + j] : 0u;
            sc.e = incl - count + add;
            sc.pub = pub;
            st_relaxed_u64(lookback + (size_t)t * kRadixSize + d, (t == sc.first_tile ? tag_i : tag_p) | pub);
            WS_PROF_END(0);
        };
        // part 2: walk back over the earlier tiles (LB descriptors in flight per step), publish the inclusive count, turn
        // the (warp, digit) counts into the start of every (warp, digit) run inside the tile buffer
        auto resolve = [&](unsigned p, unsigned t, const Scan &sc) -> Run {
            WS_PROF_BEGIN();
            unsigned *col = tab + p * kWsWorkerWarps * ROW + d;
            unsigned long long *mine = lookback + (size_t)t * kRadixSize + d;
            unsigned excl = 0;
            if (t != sc.first_tile) {
                const long long lowest = (long long)sc.first_tile;  // the walk ends at the segment's first tile at the latest
                long long j = (long long)t - 1;
                bool done = false;
                while (!done) {
                    unsigned long long v[LB];
#pragma unroll
                    for (int k = 0; k < LB; k++) {
                        const long long idx = j - k;
                        v[k] = idx >= lowest ? ld_relaxed_u64(lookback + (size_t)idx * kRadixSize + d) : tag_i;  // before the first tile: inclusive zero
                    }
                    int consumed = 0;
#pragma unroll
                    for (int k = 0; k < LB; k++) {
                        if (!done && consumed == k) {
                            const unsigned tag = (unsigned)(v[k] >> 32);
                            if ((tag >> 2) == epoch) {  // published
                                excl += (unsigned)v[k];
                                consumed = k + 1;
                                done = (tag & 3u) == kLbInclusive;
                            }
                        }
                    }
                    j -= consumed;
                    if (consumed == 0) __nanosleep(40);
                }
                st_relaxed_u64(mine, tag_i | (unsigned)(excl + sc.pub));
            }
            WS_PROF_END(1);
            // shared-memory start of the run: after the runs before it, shifted to its destination's 16-byte phase
            Run r;
            r.g = sc.gbase + excl;
            const unsigned nat = sc.e + (A - 1) * d;
            r.s = nat + ((r.g - nat) & (A - 1));
            r.c = sc.pub;
            unsigned run = r.s;
#pragma unroll
            for (int w = 0; w < kWsWorkerWarps; w++) {  // (the table still holds the counts scan_publish read)
                const unsigned c = col[w * ROW];
                col[w * ROW] = run;
                run += c;
            }
            named_bar_arrive(kBarOffsets + p, kWsThreads);
            WS_PROF_END(2);
            return r;
        };
        auto store = [&](const Run r, bool more) {
            WS_PROF_BEGIN();
            named_bar_sync(kBarScattered, kWsThreads);
            WS_PROF_END(3);
            // body: the 16-byte chunks that lie entirely inside the run, as one bulk copy
            const unsigned m = r.g & (A - 1), end = m + r.c;  // the run covers keys [m, end) counted from the chunk-aligned base
            const K *src0 = buf + (r.s - m);
            K *dst0 = kout_d + ((size_t)r.g - m);
            const unsigned first = (m + A - 1) / A, last = end / A;  // full chunks [first, last)
            if (r.c && last > first) {
                // (first * A keys / values are a whole number of 16-byte chunks in BOTH arrays: A counts elements of the narrower one)
                if (flags & 2) tma_store_issue(dst0 + first * A, src0 + first * A, (last - first) * A * (unsigned)sizeof(K));
                else tma_store_issue_hint(dst0 + first * A, src0 + first * A, (last - first) * A * (unsigned)sizeof(K), drop);
                if constexpr (VB > 0)
                    tma_store_issue_hint(vout_d + ((size_t)r.g - m) + first * A, vbuf + (r.s - m) + first * A, (last - first) * A * (unsigned)VB, drop);
            }
            WS_PROF_END(4);
            tma_commit();
            // edges: keys [m, min(first * A, end)) and [last * A, end), at most A - 1 each, by the LSU.  Lane = (run, slot):
            // 32 / (2 * (A - 1)) runs per instruction, so a warp store touches a handful of sectors instead of 32 lines.
            // (As byte-masked bulk copies -- cp.async.bulk .cp_mask, two more copies per run -- the pass was 7 % slower:
            // issuing a bulk copy costs a warp 70-300 cycles.)
            if constexpr (A > 1) {
                constexpr int SLOTS = 2 * (A - 1), PER = 32 / SLOTS;  // 6 slots x 5 runs (u32), 2 slots x 16 runs (u64)
                const unsigned sub = lane / SLOTS, slot = lane - sub * SLOTS;
                for (unsigned r0 = 0; r0 < 32; r0 += PER) {
                    const unsigned srcl = (r0 + sub) & 31u;
                    const unsigned og = __shfl_sync(0xffffffffu, r.g, srcl), os = __shfl_sync(0xffffffffu, r.s, srcl),
                                   oc = __shfl_sync(0xffffffffu, r.c, srcl);
                    K *const okout = reinterpret_cast<K *>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)kout_d, srcl));
                    V *const ovout = reinterpret_cast<V *>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)vout_d, srcl));
                    if (sub < (unsigned)PER && r0 + sub < 32 && oc) {
                        const unsigned om = og & (A - 1), oend = om + oc;
                        const unsigned ofirst = (om + A - 1) / A, olast = oend / A;
                        unsigned k;  // position of this lane's key, counted from the chunk-aligned base
                        bool on;
                        if (slot < (unsigned)(A - 1)) {  // head: [om, min(ofirst * A, oend))
                            k = om + slot;
                            on = k < ofirst * A && k < oend;
                        } else {                         // tail: [olast * A, oend), unless the run ends inside its first chunk
                            k = olast * A + (slot - (A - 1));
                            on = olast >= ofirst && k < oend;
                        }
                        if (on) {
                            okout[(size_t)og - om + k] = buf[os - om + k];
                            if constexpr (VB > 0) ovout[(size_t)og - om + k] = vbuf[os - om + k];
                        }
                    }
                }
            }
            WS_PROF_END(5);
            tma_store_wait_read<0>();
            if (more) named_bar_arrive(kBarDrained, kWsThreads);  // the buffer can be scattered into again
            WS_PROF_END(6);
        };
        named_bar_sync(kBarCounted + 0, kWsThreads);
        unsigned t_cur = ring[0];
        Run r_cur{0, 0, 0};
        Scan sc;
        if (t_cur != kWsNoTile) {
            scan_publish(0, 0, t_cur, sc);
            r_cur = resolve(0, t_cur, sc);
        }
        // Per iteration: offsets (scan, publish, look-back) of the next tile, then the stores of the current one.
        // (BCB_WS_FLAGS=4, experiment: publish, store, and only then walk the look-back -- its predecessors have had the
        // whole store phase to publish, so the walk spins half as long (cycle counters: 214K -> 107K per CTA), but the
        // bulk copies are then issued while the workers are in their scatter sweep and take longer to get through the
        // LSU queue (603K -> 771K): 2^30 keys 2.47 ms per pass against 2.41.)
        for (unsigned i = 0; t_cur != kWsNoTile; ++i) {
            WS_PROF_BEGIN();
            named_bar_sync(kBarCounted + ((i + 1) & 1u), kWsThreads);
            WS_PROF_END(7);
            const unsigned t_next = ring[(i + 1) & 3u];
            Run r_next{0, 0, 0};
            if (t_next != kWsNoTile) {
                scan_publish((i + 1) & 1u, i + 1, t_next, sc);
                if (!(flags & 4)) r_next = resolve((i + 1) & 1u, t_next, sc);
            }
            store(r_cur, t_next != kWsNoTile);
            if (t_next != kWsNoTile && (flags & 4)) r_next = resolve((i + 1) & 1u, t_next, sc);
            t_cur = t_next;
            r_cur = r_next;
        }
        if (d == 0) WS_PROF_DUMP(8);
    }
}

template <typename K, int VB, bool DET, int XF>
static int ws_launch_typed(StreamState *st, const void *kin, void *kout, const void *vin, void *vout, const unsigned *base,
                           unsigned long long *lookback, size_t n, int shift, const Transform &tf, const unsigned long long *dst_tab,
                           const uint4 *tile_tab, size_t tab_tiles, const unsigned *hot)
{
    typedef WsShape<K, VB, DET> C;
    auto kernel = onesweep_ws<K, VB, DET, XF, 8>;
    static std::atomic<unsigned long long> configured{0};  // bit per device: > 48 KB dynamic shared memory opted in
    const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
    if (!(configured.load(std::memory_order_acquire) & bit) || !bit) {
        BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        configured.fetch_or(bit, std::memory_order_release);
    }
    const size_t tiles = tile_tab ? tab_tiles : (n + C::TILE - 1) / C::TILE;
    size_t grid = (size_t)st->sm_count;
    if (grid > tiles) grid = tiles;
    unsigned epoch;
    BCB_TRY(next_epoch(st, kArenaPacked, &epoch));
    const unsigned long long ticket_base = ticket_reserve(st, tiles + grid);  // every CTA draws one void ticket
    static const int ws_flags = [] { const char *e = std::getenv("BCB_WS_FLAGS"); return e ? std::atoi(e) : 0; }();  // experiments
    LaunchTimer timer(st, dst_tab ? BCB_K_EXCHANGE_PASS : BCB_K_ONESWEEP_PASS);
    kernel<<<(unsigned)grid, kWsThreads, C::SMEM_BYTES, st->stream>>>((const K *)kin, (K *)kout, vin, vout, base, lookback, epoch, n,
                                                                      (unsigned)tiles, shift, tf, st->control + kControlTicket, ticket_base, ws_flags,
                                                                      dst_tab, tile_tab, hot);
    BCB_CUDA_TRY(cudaGetLastError());
#ifdef BCB_WS_PROFILE
    {   // mean cycles per CTA and phase (workers 0-7, helpers 8-15); see the WS_PROF_END(k) sites for what k is
        static unsigned long long host[148 * 16 * 4];
        cudaStreamSynchronize(st->stream);
        cudaMemcpyFromSymbol(host, g_ws_prof, sizeof(host));
        double mean[16] = {};
        for (size_t b = 0; b < grid; b++)
            for (int q = 0; q < 16; q++) mean[q] += (double)host[b * 16 + q] / (double)grid;
        fprintf(stderr, "ws_prof tiles/CTA %.1f | workers: draw %.0f count %.0f wait_offsets %.0f wait_drained %.0f scatter %.0f | helpers: scan %.0f lookback %.0f "
                        "offsets_out %.0f wait_scattered %.0f issue %.0f edges %.0f drain %.0f wait_counted %.0f\n",
                (double)tiles / (double)grid, mean[0], mean[1], mean[2], mean[3], mean[4], mean[8], mean[9], mean[10], mean[11], mean[12], mean[13], mean[14], mean[15]);
    }
#endif
    return BCB_SUCCESS;
}

template <typename K, int VB, bool DET>
static int ws_launch_xf(StreamState *st, const void *kin, void *kout, const void *vin, void *vout, const unsigned *base,
                        unsigned long long *lookback, size_t n, int shift, const Transform &tf, int xf, const unsigned long long *dst_tab,
                        const uint4 *tile_tab, size_t tab_tiles, const unsigned *hot)
{
    switch (xf) {
    case kXfNone: return ws_launch_typed<K, VB, DET, kXfNone>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, dst_tab, tile_tab, tab_tiles, hot);
    case kXfIn: return ws_launch_typed<K, VB, DET, kXfIn>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, dst_tab, tile_tab, tab_tiles, hot);
    case kXfOut: return ws_launch_typed<K, VB, DET, kXfOut>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, dst_tab, tile_tab, tab_tiles, hot);
    default: return ws_launch_typed<K, VB, DET, kXfBoth>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, dst_tab, tile_tab, tab_tiles, hot);
    }
}

size_t ws_tile_size(int key_bytes, int value_bytes, bool deterministic)
{
    if (value_bytes == 4) return (size_t)WsShape<unsigned, 4, true>::TILE;
    if (value_bytes == 8) return (size_t)WsShape<unsigned, 8, true>::TILE;
    if (key_bytes == 8) return (size_t)WsShape<unsigned long long, 0, false>::TILE;
    return deterministic ? (size_t)WsShape<unsigned, 0, true>::TILE : (size_t)WsShape<unsigned, 0, false>::TILE;
}

bool ws_supports(int key_bytes, int value_bytes, bool deterministic)
{
    if (value_bytes) return key_bytes == 4 && (value_bytes == 4 || value_bytes == 8);
    if (deterministic) return key_bytes == 4;
    return key_bytes == 4 || key_bytes == 8;
}

int ws_launch_pass(StreamState *st, int key_bytes, const void *kin, void *kout, const void *vin, void *vout, int value_bytes,
                   const unsigned *base, unsigned long long *lookback, size_t n, int shift, const Transform &tf, int xf, bool deterministic,
                   const unsigned long long *dst_tab, const uint4 *tile_tab, size_t tab_tiles, const unsigned *hot)
{
    // bulk copies need 16-byte aligned arrays (with dst_tab the caller has checked the destinations it holds)
    if ((((uintptr_t)kin | (uintptr_t)kout | (uintptr_t)vin | (uintptr_t)vout) & 15) != 0) return BCB_EUNSUPPORTED;
    if (!ws_supports(key_bytes, value_bytes, deterministic)) return BCB_EUNSUPPORTED;
    if (value_bytes == 4) return ws_launch_xf<unsigned, 4, true>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, xf, dst_tab, tile_tab, tab_tiles, hot);
    if (value_bytes == 8) return ws_launch_xf<unsigned, 8, true>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, xf, dst_tab, tile_tab, tab_tiles, hot);
    if (key_bytes == 8) return ws_launch_xf<unsigned long long, 0, false>(st, kin, kout, nullptr, nullptr, base, lookback, n, shift, tf, xf, dst_tab, tile_tab, tab_tiles, hot);
    return deterministic ? ws_launch_xf<unsigned, 0, true>(st, kin, kout, nullptr, nullptr, base, lookback, n, shift, tf, xf, dst_tab, tile_tab, tab_tiles, hot)
                         : ws_launch_xf<unsigned, 0, false>(st, kin, kout, nullptr, nullptr, base, lookback, n, shift, tf, xf, dst_tab, tile_tab, tab_tiles, hot);
}

}  // namespace bcb
