// radix_exchange_ws.cu -- warp-specialised exchange pass of the multi-GPU sort for sm_100a: ONE stable partition of a
// rank's unsorted shard into <= 8 buckets (bucket = number of splitters <= transformed key), each bucket written
// straight to its destination -- another GPU's receive buffer over NVLink, or local memory -- with bulk copies.
//
// Same pipeline as onesweep_ws (radix_pass_ws.cu): one CTA per SM, ticketed tiles, 24 worker warps that count tile
// t+1 and scatter tile t into a tile buffer in shared memory, a helper warp that turns counts into positions (digit
// scan over the buckets + decoupled look-back) and hands the bucket runs to the bulk-copy engine.  What differs:
//   * with <= 8 buckets same-address shared atomics would serialise (the first version of the r01 exchange kernel took
//     8.5 ms for 2^30 keys that way), so ranking uses no shared memory at all: sweep 1 counts in per-lane packed
//     registers; in sweep 2 one to three ballots over the bits of the bucket give every lane the peer mask of its
//     bucket, lane j < 8 keeps the cursor of bucket j in a register and the other lanes fetch it by shuffle.  Ballot
//     order is lane order: the partition is stable and deterministic by construction, payloads included.
//   * a run is 1/8 of a tile (~7000 keys, 27 KB): one bulk copy carries it, which is also what NVLink likes (the r01
//     kernel wrote 128-byte lines from the LSU: 4.9 ms of SM time per 2^30 keys even with every destination local);
//   * every bucket has its own destination pointer; a run's position in the tile buffer is shifted to the 16-byte phase
//     of its destination address so that its body is one aligned bulk copy (<= 3 + 3 edge elements by the LSU).
// Replaces nothing in the reference (it has no multi-device path); the result of the multi-GPU sort
// it serves is bit-identical to the single-GPU radix_sort_impl (algorithm/detail/radix_sort.hpp:252-426).
#include "radix_common.cuh"
#include "tma.cuh"

#include <atomic>
#include <cstdlib>
#include <type_traits>

namespace bcb {

constexpr int kExWorkerWarps = 24;
constexpr int kExWorkers = kExWorkerWarps * 32;
constexpr int kExThreads = kExWorkers + 32;  // + one helper warp (lane b < 8 owns bucket b)
constexpr int kExBuckets = kMaxSplitters + 1;
constexpr unsigned kExNoTile = 0xffffffffu;
enum { kExBarCounted = 1 /* +p */, kExBarOffsets = 3 /* +p */, kExBarScattered = 5, kExBarDrained = 6, kExBarWorkers = 7 };

template <int VB> struct ex_value { typedef unsigned type; };
template <> struct ex_value<8> { typedef unsigned long long type; };

template <typename K, int VB> struct ExShape {
    static constexpr int KB = (int)sizeof(K);
    static constexpr int MINB = VB ? (VB < KB ? VB : KB) : KB;
    static constexpr int A = 16 / MINB;                    // elements per 16-byte chunk of the narrower array
    static constexpr int PER32 = 32 / KB;                  // keys per 32-byte sector
    // keys per worker thread: ONE CONTIGUOUS range of whole sectors (64 u32 / 32 u64 keys, 32 u32+u32 / 16 u32+u64 pairs)
    static constexpr int ITEMS = 196 * 1024 / (KB + VB) / kExWorkers / PER32 * PER32;
    static constexpr int NS = ITEMS / PER32;               // key sectors per lane
    static constexpr int W = VB == 0 ? 4 : (VB == 4 ? 2 : 1);  // key sectors (+ their values) a lane holds in registers at a time
    static constexpr int SEG = ITEMS * 32;
    static constexpr int TILE = kExWorkers * ITEMS;        // 49152 u32 keys / 24576 u64 keys or u32+u32 pairs / 12288 u32+u64 pairs
    static constexpr int PAD = (A - 1) * kExBuckets + A;
    static constexpr size_t KBUF_BYTES = ((size_t)(TILE + PAD) * KB + 127) / 128 * 128;
    static constexpr size_t VBUF_BYTES = ((size_t)(TILE + PAD) * VB + 127) / 128 * 128;
    static constexpr size_t CUR_BYTES = (size_t)kExBuckets * kExWorkers * sizeof(unsigned);          // [8][768] per-thread cursors (bank = thread)
    static constexpr size_t TAB_BYTES = 2 * (size_t)kExWorkerWarps * kExBuckets * sizeof(unsigned);  // [2][24][8]
    static constexpr size_t MISC_BYTES = 64 + 256 + 64;  // tile-id ring, bucket look-up table, splitters
    static constexpr size_t SMEM_BYTES = KBUF_BYTES + VBUF_BYTES + CUR_BYTES + TAB_BYTES + MISC_BYTES;
    static_assert(NS % W == 0 && ITEMS < 256 && TILE + PAD < 65536, "whole rounds; 8-bit per-lane counts; 16-bit cursors");
    static_assert(SMEM_BYTES <= 232448, "one CTA per SM: 227 KB of shared memory");
};

// 32-byte sector load (LDG.256, sm_100+) with an L2 eviction-priority hint; 32-byte aligned address
struct Sector { unsigned w[8]; };
__device__ __forceinline__ Sector ld_sector(const void *p, unsigned long long policy)
{
    Sector s;
    asm volatile("ld.global.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(s.w[0]), "=r"(s.w[1]), "=r"(s.w[2]), "=r"(s.w[3]), "=r"(s.w[4]), "=r"(s.w[5]), "=r"(s.w[6]), "=r"(s.w[7])
                 : "l"(p), "l"(policy));
    return s;
}

template <typename K, int VB, bool IDENT>
__global__ void __launch_bounds__(kExThreads, 1)
exchange_ws(const K *__restrict__ keys_in, const void *__restrict__ vals_in_v, unsigned long long *lookback, unsigned epoch, size_t n,
            unsigned num_tiles, const __grid_constant__ Transform tf, unsigned long long *ticket, unsigned long long ticket_base, int flags)
{
    typedef ExShape<K, VB> C;
    typedef typename ex_value<VB>::type V;
    const V *vals_in = reinterpret_cast<const V *>(vals_in_v);
    constexpr int ITEMS = C::ITEMS, TILE = C::TILE, A = C::A, SEG = C::SEG, NB = kExBuckets, PER32 = C::PER32, NS = C::NS, W = C::W;
    extern __shared__ __align__(128) unsigned char smem[];
    K *buf = reinterpret_cast<K *>(smem);
    V *vbuf = reinterpret_cast<V *>(smem + C::KBUF_BYTES);
    unsigned *cursor = reinterpret_cast<unsigned *>(smem + C::KBUF_BYTES + C::VBUF_BYTES);  // [8][768]: thread-private next slot per bucket
    unsigned *tab = reinterpret_cast<unsigned *>(smem + C::KBUF_BYTES + C::VBUF_BYTES + C::CUR_BYTES);  // [2][24][8]: counts, then run starts
    volatile unsigned *ring = reinterpret_cast<volatile unsigned *>(smem + C::KBUF_BYTES + C::VBUF_BYTES + C::CUR_BYTES + C::TAB_BYTES);  // [4] tile ids
    unsigned char *lut = smem + C::KBUF_BYTES + C::VBUF_BYTES + C::CUR_BYTES + C::TAB_BYTES + 64;  // [256]: bucket of the first key of a top-byte bin | 8 if a splitter lies inside the bin
    typedef typename key_traits<K>::U U;
    U *ssplit = reinterpret_cast<U *>(lut + 256);                                    // [8] splitters in the key's compute width
    const unsigned tid = threadIdx.x, lane = tid & 31u;

    // Bucket of a key = number of splitters <= transformed key.  Seven compares per key made the sweeps instruction
    // bound (27 + 52 instructions per key), so the bucket comes from a 256-entry table over the most significant byte:
    // the bucket of the bin's first key, plus one compare against the single splitter that may lie inside the bin (the
    // host checks that no bin holds two; the splitters of the histogram plan sit on bin edges: no compare ever counts).
    constexpr int TOPSHIFT = (int)sizeof(K) * 8 - 8;
    if (tid < 256) {
        const U first = (U)tid << TOPSHIFT, last = first | (((U)1 << TOPSHIFT) - 1);
        unsigned lo = 0, hi = 0;
        for (int j = 0; j < tf.nsplit; j++) {
            lo += (first >= (U)tf.split[j]) ? 1u : 0u;
            hi += (last >= (U)tf.split[j]) ? 1u : 0u;
        }
        lut[tid] = (unsigned char)(lo | (hi > lo ? 8u : 0u));
        if (tid < kExBuckets) ssplit[tid] = tid < (unsigned)tf.nsplit ? (U)tf.split[tid] : (U)~(U)0;
    }
    __syncthreads();

    if (tid < kExWorkers) {
        // ======================= workers =======================
        const unsigned w = tid >> 5;
        const unsigned long long keep = l2_policy_evict_last(), drop = l2_policy_evict_first();
        // branch-free: the second look-up is done even where no splitter lies inside the bin (ssplit[lo] is then a
        // splitter beyond the bin or all-ones, and the flag masks the compare out)
        // Default: the splitter compares in registers (3.5 / 3.6 / 4.0 ms per 2^30 keys for 1 / 3 / 7 splitters).  The table
        // look-up (BCB_SPLIT_WS_FLAGS=1) needs fewer instructions but puts two dependent shared-memory loads in front of
        // every key: 4.3 ms whatever the splitter count -- the sweeps are latency bound, not issue bound.
        const bool by_compares = (flags & 1) == 0;
        auto bucket = [&](K raw) -> unsigned {
            if (by_compares) return pass_digit<K, kDigitSplit>(raw, 0, tf);
            const U t = IDENT ? (U)raw : transform_fwd<K>(raw, tf);
            const unsigned e = lut[(unsigned)(t >> TOPSHIFT)];
            return (e & 7u) + ((e >> 3) & (t >= ssplit[e & 7u] ? 1u : 0u));
        };
        auto draw = [&](unsigned i) {
            if (tid == 0) {
                const unsigned long long t = atomicAdd(ticket, 1ull) - ticket_base;
                ring[i & 3u] = t < num_tiles ? (unsigned)t : kExNoTile;
            }
            named_bar_sync(kExBarWorkers, kExWorkers);
            return ring[i & 3u];
        };
        // Ranking without a single cross-lane operation per key: a lane owns ONE CONTIGUOUS range of ITEMS keys of its
        // warp's segment (whole 32-byte sectors, read with 256-bit loads), so the stable order inside a bucket is
        // (warp, lane, position in the lane's range).  Sweep 1 counts the lane's keys per bucket in packed registers;
        // a warp scan of the packed counts gives the lane its offset inside the warp's (warp, bucket) runs; in sweep 2
        // every thread bumps its OWN eight cursors (16-bit, shared memory, bank = thread: conflict-free).  (The first
        // version ranked keys that were interleaved across the lanes with 3-4 ballots + 2 shuffles per key: 106
        // instructions per 32 keys in all, issue bound at 4.6 ms per 2^30 keys.)
        struct LaneOffs { unsigned long long lo, hi; };  // exclusive prefix over the lower lanes, buckets 0-3 / 4-7, 16-bit fields
        auto spread = [](unsigned x) -> unsigned long long {  // four 8-bit fields -> four 16-bit fields
            unsigned long long v = x;
            v = (v | (v << 16)) & 0x0000ffff0000ffffull;
            v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
            return v;
        };
        auto count = [&](unsigned p, unsigned t) -> LaneOffs {
            const size_t base = (size_t)t * TILE + (size_t)w * SEG + (size_t)lane * ITEMS;
            unsigned pc_lo = 0, pc_hi = 0;  // this lane's keys per bucket, 8-bit fields (ITEMS < 256): buckets 0-3 / 4-7
            auto tally = [&](unsigned b) {
                const unsigned inc = 1u << ((b & 3u) * 8u);
                if (b & 4u) pc_hi += inc;
                else pc_lo += inc;
            };
            if ((size_t)t * TILE + TILE <= n) {
                const K *src = keys_in + base;
#pragma unroll 1
                for (int r = 0; r < NS / W; r++) {
                    Sector sec[W];
#pragma unroll
                    for (int j = 0; j < W; j++) sec[j] = ld_sector(src + (r * W + j) * PER32, keep);  // stays in L2 for sweep 2
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        const K *e = reinterpret_cast<const K *>(sec[j].w);
#pragma unroll
                        for (int c = 0; c < PER32; c++) tally(bucket(e[c]));
                    }
                }
            } else {
#pragma unroll 4
                for (int i = 0; i < ITEMS; i++)
                    if (base + i < n) tally(bucket(__ldg(keys_in + base + i)));
            }
            // widen to 16-bit fields (a warp's segment has < 65536 keys), inclusive scan over the lanes
            const unsigned long long own_lo = spread(pc_lo), own_hi = spread(pc_hi);
            unsigned long long lo = own_lo, hi = own_hi;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long a = __shfl_up_sync(0xffffffffu, lo, off), c = __shfl_up_sync(0xffffffffu, hi, off);
                if ((int)lane >= off) { lo += a; hi += c; }
            }
            const unsigned long long tot_lo = __shfl_sync(0xffffffffu, lo, 31), tot_hi = __shfl_sync(0xffffffffu, hi, 31);
            if (lane < (unsigned)NB) tab[(p * kExWorkerWarps + w) * NB + lane] = (unsigned)(((lane & 4u) ? tot_hi : tot_lo) >> (16 * (lane & 3u))) & 0xffffu;
            named_bar_arrive(kExBarCounted + p, kExThreads);
            return LaneOffs{lo - own_lo, hi - own_hi};
        };
        auto scatter_tile = [&](unsigned p, unsigned t, bool first, const LaneOffs lx, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const size_t base = (size_t)t * TILE + (size_t)w * SEG + (size_t)lane * ITEMS;
            const K *ksrc = keys_in + base;
            [[maybe_unused]] const V *vsrc = vals_in + base;
            Sector sec[W];
            constexpr int VS = VB ? W * VB / (int)sizeof(K) : 1;  // value sectors per round
            [[maybe_unused]] Sector vsec[VS];
            auto load_round = [&](int r) {
#pragma unroll
                for (int j = 0; j < W; j++) sec[j] = ld_sector(ksrc + (r * W + j) * PER32, drop);  // second read of the tile: an L2 hit, last use
                if constexpr (VB > 0) {
#pragma unroll
                    for (int j = 0; j < VS; j++) vsec[j] = ld_sector(vsrc + (r * W * PER32) + j * (32 / VB), drop);
                }
            };
            if constexpr (FULL) load_round(0);
            named_bar_sync(kExBarOffsets + p, kExThreads);          // the table holds the start of every (warp, bucket) run
            if (!first) named_bar_sync(kExBarDrained, kExThreads);  // the previous tile's bulk copies have read the buffer
            {   // this thread's cursors: start of the (warp, bucket) run + the keys of that bucket in the lower lanes
                const uint4 a = *reinterpret_cast<const uint4 *>(tab + (p * kExWorkerWarps + w) * NB);
                const uint4 c = *reinterpret_cast<const uint4 *>(tab + (p * kExWorkerWarps + w) * NB + 4);
                cursor[0 * kExWorkers + tid] = a.x + (unsigned)(lx.lo & 0xffffu);
                cursor[1 * kExWorkers + tid] = a.y + (unsigned)((lx.lo >> 16) & 0xffffu);
                cursor[2 * kExWorkers + tid] = a.z + (unsigned)((lx.lo >> 32) & 0xffffu);
                cursor[3 * kExWorkers + tid] = a.w + (unsigned)(lx.lo >> 48);
                cursor[4 * kExWorkers + tid] = c.x + (unsigned)(lx.hi & 0xffffu);
                cursor[5 * kExWorkers + tid] = c.y + (unsigned)((lx.hi >> 16) & 0xffffu);
                cursor[6 * kExWorkers + tid] = c.z + (unsigned)((lx.hi >> 32) & 0xffffu);
                cursor[7 * kExWorkers + tid] = c.w + (unsigned)(lx.hi >> 48);
            }
            auto place = [&](K key, [[maybe_unused]] V val) {
                unsigned *cp = cursor + bucket(key) * kExWorkers + tid;
                const unsigned pos = *cp;
                *cp = pos + 1;
                buf[pos] = key;
                if constexpr (VB > 0) vbuf[pos] = val;
            };
            if constexpr (FULL) {
#pragma unroll 1
                for (int r = 0; r < NS / W; r++) {
                    if (r) load_round(r);
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        const K *e = reinterpret_cast<const K *>(sec[j].w);
#pragma unroll
                        for (int c = 0; c < PER32; c++) {
                            V val = V();
                            if constexpr (VB > 0) {
                                constexpr int VPS = 32 / (VB ? VB : 1);  // values per sector
                                val = reinterpret_cast<const V *>(vsec[(j * PER32 + c) / VPS].w)[(j * PER32 + c) % VPS];
                            }
                            place(e[c], val);
                        }
                    }
                }
            } else {
#pragma unroll 2
                for (int i = 0; i < ITEMS; i++) {
                    if (base + i < n) {
                        V val = V();
                        if constexpr (VB > 0) val = __ldg(vals_in + base + i);
                        place(__ldg(keys_in + base + i), val);
                    }
                }
            }
            fence_proxy_async();
            named_bar_arrive(kExBarScattered, kExThreads);
        };
        auto scatter = [&](unsigned p, unsigned t, bool first, const LaneOffs lx) {
            if ((size_t)t * TILE + TILE <= n) scatter_tile(p, t, first, lx, std::true_type());
            else scatter_tile(p, t, first, lx, std::false_type());
        };
        unsigned t_cur = draw(0);
        LaneOffs lx_cur{0, 0};
        if (t_cur != kExNoTile) lx_cur = count(0, t_cur);
        else named_bar_arrive(kExBarCounted + 0, kExThreads);
        for (unsigned i = 0; t_cur != kExNoTile; ++i) {
            const unsigned t_next = draw(i + 1);
            LaneOffs lx_next{0, 0};
            if (t_next != kExNoTile) lx_next = count((i + 1) & 1u, t_next);
            else named_bar_arrive(kExBarCounted + ((i + 1) & 1u), kExThreads);
            scatter(i & 1u, t_cur, i == 0, lx_cur);
            t_cur = t_next;
            lx_cur = lx_next;
        }
    } else {
        // ======================= helper warp: lane b < 8 owns bucket b =======================
        const unsigned b = lane;
        const bool owner = b <= (unsigned)tf.nsplit;
        const unsigned long long drop = l2_policy_evict_first();
        const unsigned long long tag_p = (unsigned long long)((epoch << 2) | kLbPartial) << 32;
        const unsigned long long tag_i = (unsigned long long)((epoch << 2) | kLbInclusive) << 32;
        // destination of the bucket, re-based to a 16-byte aligned address: element g0 of that array is dst[0]
        const unsigned long long dk = owner ? tf.dst_keys[b] : 0ull;
        const unsigned g0 = (unsigned)((dk & 15ull) / sizeof(K));
        K *kbase = reinterpret_cast<K *>(dk - (unsigned long long)g0 * sizeof(K));
        V *vbase = reinterpret_cast<V *>((VB && owner) ? tf.dst_vals[b] - (unsigned long long)g0 * VB : 0ull);
        struct Run { unsigned long long g; unsigned s, c; };  // first element (re-based index), first slot in the tile buffer, length
        constexpr int LB = 16;  // descriptors in flight per look-back step
        struct Scan { unsigned count, e; };
        // part 1: bucket counts of the tile, exclusive scan over the buckets, publish the counts (successors never wait for more)
        auto scan_publish = [&](unsigned p, unsigned t) -> Scan {
            const unsigned *col = tab + p * kExWorkerWarps * NB + (b & 7u);
            unsigned count = 0;
            if (b < (unsigned)NB) {
#pragma unroll
                for (int w = 0; w < kExWorkerWarps; w++) count += col[w * NB];
            }
            unsigned incl = count;  // exclusive scan over the buckets -> start of each run in the bucket-sorted tile
#pragma unroll
            for (int off = 1; off < NB; off <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
                if ((int)lane >= off) incl += o;
            }
            if (owner) st_relaxed_u64(lookback + (size_t)t * NB + b, (t == 0 ? tag_i : tag_p) | count);
            return Scan{count, incl - count};
        };
        // part 2: decoupled look-back, publish the inclusive count, turn the (warp, bucket) counts into run starts
        auto resolve = [&](unsigned p, unsigned t, const Scan sc) -> Run {
            unsigned *col = tab + p * kExWorkerWarps * NB + (b & 7u);
            unsigned long long excl = 0;
            if (owner && t != 0) {
                long long j = (long long)t - 1;
                bool done = false;
                while (!done) {
                    unsigned long long v[LB];
#pragma unroll
                    for (int k = 0; k < LB; k++) {
                        const long long idx = j - k;
                        v[k] = idx >= 0 ? ld_relaxed_u64(lookback + (size_t)idx * NB + b) : tag_i;  // before tile 0: inclusive zero
                    }
                    int consumed = 0;
#pragma unroll
                    for (int k = 0; k < LB; k++) {
                        if (!done && consumed == k) {
                            const unsigned tag = (unsigned)(v[k] >> 32);
                            if ((tag >> 2) == epoch) {
                                excl += (unsigned)v[k];
                                consumed = k + 1;
                                done = (tag & 3u) == kLbInclusive;
                            }
                        }
                    }
                    j -= consumed;
                    if (consumed == 0) __nanosleep(40);
                }
                // (a bucket of one source holds < 2^32 keys: n does)
                st_relaxed_u64(lookback + (size_t)t * NB + b, tag_i | (unsigned)(excl + sc.count));
            }
            Run r;
            r.g = (unsigned long long)g0 + excl;
            const unsigned nat = sc.e + (A - 1) * b;
            r.s = nat + (((unsigned)r.g - nat) & (A - 1));
            r.c = sc.count;
            if (b < (unsigned)NB) {
                unsigned run = r.s;
#pragma unroll
                for (int w = 0; w < kExWorkerWarps; w++) {
                    const unsigned c = col[w * NB];
                    col[w * NB] = run;
                    run += c;
                }
            }
            __syncwarp();
            named_bar_arrive(kExBarOffsets + p, kExThreads);
            return r;
        };
        auto store_issue = [&](const Run r) {
            named_bar_sync(kExBarScattered, kExThreads);
            if (owner && r.c) {
                const unsigned m = (unsigned)(r.g & (A - 1)), end = m + r.c;  // the run covers [m, end) counted from its chunk-aligned start
                const K *src0 = buf + (r.s - m);
                K *dst0 = kbase + (r.g - m);
                const unsigned first = (m + A - 1) / A, last = end / A;  // whole 16-byte chunks [first, last)
                if (last > first) {
                    tma_store_issue_hint(dst0 + first * A, src0 + first * A, (last - first) * A * (unsigned)sizeof(K), drop);
                    if constexpr (VB > 0)
                        tma_store_issue_hint(vbase + (r.g - m) + first * A, vbuf + (r.s - m) + first * A, (last - first) * A * (unsigned)VB, drop);
                }
                // edges: [m, min(first * A, end)) and [max(last, first) * A, end), at most A - 1 elements each
                const unsigned head_end = first * A < end ? first * A : end;
                for (unsigned k = m; k < head_end; k++) {
                    dst0[k] = src0[k];
                    if constexpr (VB > 0) (vbase + (r.g - m))[k] = (vbuf + (r.s - m))[k];
                }
                if (last >= first) {
                    for (unsigned k = last * A; k < end; k++) {
                        dst0[k] = src0[k];
                        if constexpr (VB > 0) (vbase + (r.g - m))[k] = (vbuf + (r.s - m))[k];
                    }
                }
            }
            tma_commit();
        };
        auto store_wait = [&](bool more) {
            tma_store_wait_read<0>();
            __syncwarp();
            if (more) named_bar_arrive(kExBarDrained, kExThreads);  // the buffer can be scattered into again
        };
        named_bar_sync(kExBarCounted + 0, kExThreads);
        unsigned t_cur = ring[0];
        Run r_cur{0, 0, 0};
        if (t_cur != kExNoTile) r_cur = resolve(0, t_cur, scan_publish(0, t_cur));
        // Per iteration: publish the next tile's counts, hand the current tile to the copy engine, and walk the next
        // tile's look-back WHILE the copies drain (issuing eight copies is cheap here, unlike the 256 of a radix pass)
        for (unsigned i = 0; t_cur != kExNoTile; ++i) {
            named_bar_sync(kExBarCounted + ((i + 1) & 1u), kExThreads);
            const unsigned t_next = ring[(i + 1) & 3u];
            Scan sc{0, 0};
            if (t_next != kExNoTile) sc = scan_publish((i + 1) & 1u, t_next);
            store_issue(r_cur);
            Run r_next{0, 0, 0};
            if (t_next != kExNoTile) r_next = resolve((i + 1) & 1u, t_next, sc);
            store_wait(t_next != kExNoTile);
            t_cur = t_next;
            r_cur = r_next;
        }
    }
}

template <typename K, int VB>
static int exchange_launch_typed(StreamState *st, const void *kin, const void *vin, size_t n, const Transform &tf)
{
    typedef ExShape<K, VB> C;
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0;  // unsigned ascending keys: the key is its own sortable form
    auto kernel = ident ? exchange_ws<K, VB, true> : exchange_ws<K, VB, false>;
    static std::atomic<unsigned long long> configured[2];  // bit per device: > 48 KB dynamic shared memory opted in
    const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
    if (!(configured[ident].load(std::memory_order_acquire) & bit) || !bit) {
        BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        configured[ident].fetch_or(bit, std::memory_order_release);
    }
    const size_t tiles = (n + C::TILE - 1) / C::TILE;
    size_t grid = (size_t)st->sm_count;
    if (grid > tiles) grid = tiles;
    void *lb;
    BCB_TRY(lookback_reserve(st, kArenaPacked, tiles * kExBuckets * sizeof(unsigned long long), &lb));
    unsigned epoch;
    BCB_TRY(next_epoch(st, kArenaPacked, &epoch));
    const unsigned long long ticket_base = ticket_reserve(st, tiles + grid);  // every CTA draws one void ticket
    const char *fe = std::getenv("BCB_SPLIT_WS_FLAGS");  // experiments
    const int ex_flags = fe ? std::atoi(fe) : 0;
    LaunchTimer timer(st, BCB_K_EXCHANGE_PASS);
    kernel<<<(unsigned)grid, kExThreads, C::SMEM_BYTES, st->stream>>>((const K *)kin, vin, (unsigned long long *)lb, epoch, n, (unsigned)tiles, tf,
                                                                      st->control + kControlTicket, ticket_base, ex_flags);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// Returns BCB_EUNSUPPORTED for shapes this kernel does not cover (the caller then uses the LSU exchange kernel of
// radix_sort.cu): key widths other than 4 / 8 bytes, 64-bit keys with a payload, sources that are not 16-byte aligned,
// sources that are not 32-byte aligned, destinations whose key and value
// addresses do not share a 16-byte phase, two splitters strictly inside one top-byte bin.
int ws_exchange_pass(StreamState *st, int key_bytes, const void *kin, const void *vin, int value_bytes, size_t n, const Transform &tf)
{
    if ((((uintptr_t)kin | (uintptr_t)vin) & 31) != 0) return BCB_EUNSUPPORTED;  // 256-bit loads
    if (!(key_bytes == 4 || (key_bytes == 8 && value_bytes == 0))) return BCB_EUNSUPPORTED;
    if (!(value_bytes == 0 || value_bytes == 4 || value_bytes == 8)) return BCB_EUNSUPPORTED;
    {   // the bucket look-up table resolves at most one splitter strictly inside a top-byte bin (the splitters are sorted)
        const int shift = key_bytes * 8 - 8;
        const unsigned long long low = (1ull << shift) - 1;
        for (int j = 0; j + 1 < tf.nsplit; j++)
            if ((tf.split[j] & low) && (tf.split[j + 1] & low) && (tf.split[j] >> shift) == (tf.split[j + 1] >> shift)) return BCB_EUNSUPPORTED;
    }
    for (int b = 0; b <= tf.nsplit; b++) {
        const unsigned long long g0 = (tf.dst_keys[b] & 15ull) / (unsigned)key_bytes;
        if (value_bytes && ((tf.dst_vals[b] - g0 * (unsigned)value_bytes) & 15ull)) return BCB_EUNSUPPORTED;
    }
    if (key_bytes == 8) return exchange_launch_typed<unsigned long long, 0>(st, kin, nullptr, n, tf);
    if (value_bytes == 4) return exchange_launch_typed<unsigned, 4>(st, kin, vin, n, tf);
    if (value_bytes == 8) return exchange_launch_typed<unsigned, 8>(st, kin, vin, n, tf);
    return exchange_launch_typed<unsigned, 0>(st, kin, nullptr, n, tf);
}

}  // namespace bcb
