// radix_sort.cu -- stable LSD radix sort (keys and key-value pairs) for sm_100a: onesweep, 8-bit digits.
//
// Replaces radix_sort_impl of the reference (algorithm/detail/radix_sort.hpp:252-426): 4-bit digits, per pass a
// count kernel, a recursive exclusive_scan over per-block counters, a 16-entry scan task and a scatter kernel
// whose in-block rank is an O(block) serial loop -- about 100 B/key of DRAM traffic and 48+ launches for 32-bit
// keys.  Here a 32-bit sort is 6 launches and 36 B/key (K + P*2*(K+V), P = key bytes):
//   1. radix_histogram: ONE read of the keys builds the 256-bin histogram of every digit position
//      (128-bit loads, shared-memory atomics, one flush per CTA);
//   2. digit_scan: exclusive scan of each 256-bin histogram -> global base of every digit value;
//   3. onesweep_pass, once per digit: persistent CTAs take tiles round-robin (the next tile's keys are prefetched
//      into the dead key registers), rank the keys of each warp's contiguous segment against per-warp
//      shared-memory digit tables (shared-memory atomics, see "Ranking modes" below -- hardware match.any is far too
//      slow on this part), publish the tile's 256 digit counts, obtain the counts of all earlier tiles by batched
//      decoupled look-back (one thread per digit value), reorder the tile through shared memory and write each
//      digit's run to its final position.  The scatter is stable (ranks follow the original order), so the result
//      is identical to the reference's stable 4-bit LSD sort by the same transformed key.
//   4. large keys-only sorts use the faster "two-sweep" ranking speculatively and verify the result
//      (verify_sorted_kernel, one more read); see "Speculative ranking" below.
// Keys stay in their original bit pattern in memory; the reference's order-preserving transform
// (radix_sort.hpp:100-127, asc and desc, incl. its descending quirks) is applied in registers when a digit is
// extracted, because the descending float transform is not invertible (-0.0 and +denorm_min collide).
// Look-back words are 64-bit {epoch<<2|status, count} so no per-pass initialisation is needed and counts up to
// 2^32-1 fit.
#include "radix_common.cuh"

#include <atomic>
#include <cstdlib>
#include <type_traits>
#include <cstring>

namespace bcb {

// ---- 1. histogram of all digit positions in one read ------------------------------------------
template <typename K>
__global__ void __launch_bounds__(kHistThreads)
radix_histogram(const K *__restrict__ keys, size_t n, unsigned *__restrict__ hist, Transform tf, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;  // fallback launch of a speculative sort whose verification passed
    constexpr int NPASS = sizeof(K);
    constexpr int VEC = 16 / sizeof(K);
    __shared__ unsigned sh[NPASS][kRadixSize];
    for (int i = threadIdx.x; i < NPASS * kRadixSize; i += kHistThreads) (&sh[0][0])[i] = 0;
    __syncthreads();

    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t head = ((16 - ((uintptr_t)keys & 15)) & 15) / sizeof(K);
    if (head > n) head = n;
    const size_t nvec = (n - head) / VEC;
    const size_t tail_start = head + nvec * VEC;
    const uint4 *vin = reinterpret_cast<const uint4 *>(keys + head);

    auto count_key = [&](K k) {
#pragma unroll
        for (int p = 0; p < NPASS; p++) atomicAdd(&sh[p][digit_of<K>(k, p * kRadixBits, tf)], 1u);
    };
    size_t v = gid;
    for (; v + gthreads < nvec; v += 2 * gthreads) {  // two independent 128-bit loads in flight
        const uint4 a = ld_stream_v4(vin + v);
        const uint4 b = ld_stream_v4(vin + v + gthreads);
        const K *ea = reinterpret_cast<const K *>(&a);
        const K *eb = reinterpret_cast<const K *>(&b);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(ea[k]);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(eb[k]);
    }
    for (; v < nvec; v += gthreads) {
        const uint4 a = ld_stream_v4(vin + v);
        const K *ea = reinterpret_cast<const K *>(&a);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(ea[k]);
    }
    if (gid < head) count_key(keys[gid]);
    if (tail_start + gid < n) count_key(keys[tail_start + gid]);

    __syncthreads();
    for (int i = threadIdx.x; i < NPASS * kRadixSize; i += kHistThreads) {
        const unsigned c = (&sh[0][0])[i];
        if (c) atomicAdd(hist + i, c);
    }
}

// Variant with one counter COLUMN per lane: bin (p, d) of lane l lives at word (p*256 + d)*COLS + (l % COLS), i.e. in
// bank l -- the 32 atomics of a warp instruction never share a bank, whatever the digits are (the plain layout above
// pays ~3 wavefronts per instruction for random digits).  One 1024-thread CTA per SM owns NPASS x 256 x COLS counters
// (128 KB for 32- and 64-bit keys); the columns are summed when the CTA flushes.
template <typename K, int COLS, bool IDENT>
__global__ void __launch_bounds__(1024, 1)
radix_histogram_columns(const K *__restrict__ keys, size_t n, unsigned *__restrict__ hist, Transform tf, const int *__restrict__ gate)
{
    if (gate && *gate == 0) return;
    constexpr int NPASS = sizeof(K);
    constexpr int VEC = 16 / sizeof(K);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *sh = reinterpret_cast<unsigned *>(smem_raw);
    for (int i = threadIdx.x; i < NPASS * kRadixSize * COLS / 4; i += blockDim.x) reinterpret_cast<uint4 *>(sh)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    unsigned *mine = sh + (threadIdx.x & (COLS - 1));

    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t head = ((16 - ((uintptr_t)keys & 15)) & 15) / sizeof(K);
    if (head > n) head = n;
    const size_t nvec = (n - head) / VEC;
    const size_t tail_start = head + nvec * VEC;
    const uint4 *vin = reinterpret_cast<const uint4 *>(keys + head);

    auto count_key = [&](K k) {
#pragma unroll
        for (int p = 0; p < NPASS; p++) {
            const unsigned d = IDENT ? ((unsigned)(k >> (p * kRadixBits)) & (kRadixSize - 1)) : digit_of<K>(k, p * kRadixBits, tf);
            atomicAdd(mine + (p * kRadixSize + d) * COLS, 1u);
        }
    };
    size_t v = gid;
    for (; v + gthreads < nvec; v += 2 * gthreads) {  // two independent 128-bit loads in flight
        const uint4 a = ld_stream_v4(vin + v);
        const uint4 b = ld_stream_v4(vin + v + gthreads);
        const K *ea = reinterpret_cast<const K *>(&a);
        const K *eb = reinterpret_cast<const K *>(&b);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(ea[k]);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(eb[k]);
    }
    for (; v < nvec; v += gthreads) {
        const uint4 a = ld_stream_v4(vin + v);
        const K *ea = reinterpret_cast<const K *>(&a);
#pragma unroll
        for (int k = 0; k < VEC; k++) count_key(ea[k]);
    }
    if (gid < head) count_key(keys[gid]);
    if (tail_start + gid < n) count_key(keys[tail_start + gid]);

    __syncthreads();
    for (int bin = threadIdx.x; bin < NPASS * kRadixSize; bin += blockDim.x) {
        unsigned c = 0;
#pragma unroll
        for (int j = 0; j < COLS; j++) c += sh[bin * COLS + ((j + threadIdx.x) & (COLS - 1))];  // rotated: conflict-free
        if (c) atomicAdd(hist + bin, c);
    }
}

// ---- 2. exclusive scan of each digit histogram: hist[p][d] -> base[p][d] -------------------------
__global__ void __launch_bounds__(kRadixSize) digit_scan(const unsigned *__restrict__ hist, unsigned *__restrict__ base,
                                                         const int *__restrict__ gate = nullptr, unsigned long long *fallbacks = nullptr,
                                                         unsigned *__restrict__ hot = nullptr)
{
    if (gate && *gate == 0) return;
    if (fallbacks && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(fallbacks, 1ull);  // a gated launch that runs IS a fallback
    __shared__ unsigned wsum[kRadixSize / 32];
    __shared__ unsigned long long wmax[kRadixSize / 32];
    const unsigned d = threadIdx.x, lane = d & 31u, warp = d >> 5;
    const unsigned c = hist[blockIdx.x * kRadixSize + d];
    unsigned s = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, s, off);
        if ((int)lane >= off) s += o;
    }
    if (lane == 31) wsum[warp] = s;
    __syncthreads();
    unsigned add = 0, total = 0;
    for (unsigned w = 0; w < kRadixSize / 32; w++) {
        if (w < warp) add += wsum[w];
        total += wsum[w];
    }
    base[blockIdx.x * kRadixSize + d] = s - c + add;
    if (!hot) return;
    // a digit value that holds at least half of the keys (and then a second one with at least a quarter): the
    // warp-specialised pass ranks those by ballot instead of same-address shared atomics, which serialise
    // (hot[2 * pass + {0, 1}], kNoHotDigit when there is none).  Measured on B200, 2^28 u32 keys: keys < 2^16 61 -> 83
    // Gkeys/s, keys < 1000 59 -> 76; with shares of 30-40 % (perf_sort_float's exponent byte) the ballots cost what they
    // save (81.4 against 79.9 Gkeys/s), hence the thresholds.
    unsigned long long mine = ((unsigned long long)c << 8) | d, first = 0;
    for (int round = 0; round < 2; round++) {
        unsigned long long m = mine;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, off);
            m = o > m ? o : m;
        }
        __syncthreads();  // (wmax of the previous round has been read)
        if (lane == 0) wmax[warp] = m;
        __syncthreads();
        unsigned long long best = 0;
        for (unsigned w = 0; w < kRadixSize / 32; w++) best = wmax[w] > best ? wmax[w] : best;
        const bool is_hot = total != 0 && (round == 0 ? (best >> 8) * 2 >= (unsigned long long)total
                                                      : ((first >> 8) * 2 >= (unsigned long long)total && (best >> 8) * 4 >= (unsigned long long)total));
        if (d == 0) hot[blockIdx.x * 2 + round] = is_hot ? (unsigned)(best & 0xffu) : kNoHotDigit;
        if (round == 1 && d == 0 && total != 0 && (first >> 8) == (unsigned long long)total) hot[blockIdx.x * 2 + 1] = kConstDigit;
        if (round == 0) {
            first = best;
            if (mine == first) mine = 0;  // the runner-up comes from the others
        }
    }
}

// ---- 3. one onesweep pass ---------------------------------------------------------------------------
// Ranking modes.  Measured on B200 (bench/microbench.cu, cycles per warp-instruction per SM): hardware match.any
// with ~30 distinct values costs 62, an 8-ballot software match 27, shared-memory atomics / loads about 1.5-1.8.
// At HBM speed an SM must retire 32 keys every ~11 cycles, so both match flavours are out; ranking is built from
// shared-memory atomics instead:
//   kRankAtomicOr     (default for pairs and small inputs): every lane ORs its lane bit into a per-warp {mask, count} entry of its digit;
//                     after a warp barrier the entry holds the full peer mask (order independent), the rank inside
//                     the round is popc(mask & lanes below), the highest peer clears the mask and bumps the count.
//                     Deterministic by construction: 3 shared-memory instructions per key.
//   kRankBallot       splitter mode only (at most 8 buckets + padding): the peer mask of a key comes from three
//                     ballots over the bits of its bucket, the running count of bucket b lives in a register of lane
//                     b.  No shared memory, no atomics (same-address atomics of 2..8 buckets would serialise).
//   kRankTwoSweep     sweep 1 only counts (atomicAdd, result unused), the tables turn into offsets, sweep 2 repeats
//                     the same atomicAdds in the same order and takes the returned value as the key's position.  One
//                     instruction per key and sweep and no rank registers, but stable only if same-address atomics of
//                     one warp instruction are applied in lane order, which CUDA does not promise -- usable only where
//                     the result is verified (keys-only sorts, below).  Keys only.
enum { kRankAtomicOr = 0, kRankBallot = 2, kRankTwoSweep = 3 };

template <int VB> struct value_type;
template <> struct value_type<0> { typedef unsigned char type; };
template <> struct value_type<1> { typedef unsigned char type; };
template <> struct value_type<2> { typedef unsigned short type; };
template <> struct value_type<4> { typedef unsigned type; };
template <> struct value_type<8> { typedef unsigned long long type; };
template <> struct value_type<16> { typedef uint4 type; };

template <typename K, int VB, int THREADS, int ITEMS, int RANK>
struct PassSmem {
    static constexpr int WARPS = THREADS / 32;
    // lanes per "virtual warp" (one digit table and one contiguous key segment each): 16 for the atomic-OR ranking,
    // whose {count:16 | mask:16} entries need a 16-bit peer mask; a full warp for the ordered-atomics ranking
    static constexpr int VWL = (RANK == kRankAtomicOr) ? 16 : 32;
    static constexpr int VWARPS = THREADS / VWL;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr size_t kElem = (sizeof(K) > (size_t)VB) ? sizeof(K) : (size_t)VB;
    static constexpr size_t kWarpTab = (size_t)VWARPS * kRadixSize * sizeof(unsigned);
    static constexpr size_t kSmall = kRadixSize * sizeof(unsigned) + 64;  // out_base + misc
    static constexpr size_t kBytes = kWarpTab + kSmall + (size_t)TILE * kElem + 16;
};

// Tile layout: 16-lane virtual warp v = tid/16 owns the contiguous segment [v*ITEMS*16, (v+1)*ITEMS*16) of the tile;
// its lane h holds items i*16 + h.  One warp instruction therefore reads two 64-byte pieces (full 32-byte sectors).
template <int ITEMS, int VWL>
__device__ __forceinline__ unsigned tile_offset_of_item0()
{
    return (threadIdx.x / VWL) * (ITEMS * VWL) + (threadIdx.x % VWL);
}

template <typename K, int THREADS, int ITEMS, int VWL>
__device__ __forceinline__ void load_tile_keys(const K *__restrict__ keys_in, size_t n, size_t tile, K (&key)[ITEMS])
{
    constexpr int TILE = THREADS * ITEMS;
    const size_t tile_base = tile * (size_t)TILE;
    const unsigned off0 = tile_offset_of_item0<ITEMS, VWL>();
    if (tile_base + TILE <= n) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = __ldg(keys_in + tile_base + off0 + i * VWL);
    } else {
        const unsigned valid = (unsigned)(n - tile_base);
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned t = off0 + i * VWL;
            key[i] = t < valid ? __ldg(keys_in + tile_base + t) : (K)0;
        }
    }
}

template <typename K, int VB, int THREADS, int ITEMS, int LBATCH, int RANK, int IDENT, bool FULL>
__device__ __forceinline__ void
pass_tile(const K *__restrict__ keys_in, K *__restrict__ keys_out, const void *__restrict__ vals_in_v,
          void *__restrict__ vals_out_v, const unsigned *__restrict__ digit_base, unsigned long long *lookback,
          unsigned epoch, size_t n, int shift, const Transform &tf, size_t tile, unsigned char *smem_raw,
          K (&key)[ITEMS], size_t num_tiles, volatile unsigned *next_tile)
{
    typedef PassSmem<K, VB, THREADS, ITEMS, RANK> L;
    typedef typename value_type<VB>::type V;
    constexpr int TILE = L::TILE;

    unsigned *out_base = reinterpret_cast<unsigned *>(smem_raw + L::kWarpTab);
    unsigned *misc = out_base + kRadixSize;  // [0..1] tile ids (onesweep_pass), [2..9] warp sums of the digit scan
    unsigned char *elem_buf = smem_raw + L::kWarpTab + L::kSmall;
    K *keys_sorted = reinterpret_cast<K *>(elem_buf);
    // digit table of each 16-lane virtual warp.  kRankAtomicOr packs {count : 16 | peer mask : 16} into one word
    // (a 32-bit entry keeps every access a single-wavefront-per-bank operation); the other modes store the count.
    unsigned *tab = reinterpret_cast<unsigned *>(smem_raw);  // [VWARPS][256]
    constexpr int CSHIFT = (RANK == kRankAtomicOr) ? 16 : 0;
    constexpr int VWARPS = L::VWARPS, VWL = L::VWL;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, vwarp = tid / VWL;
    const size_t tile_base = tile * (size_t)TILE;
    const unsigned valid = FULL ? (unsigned)TILE : (unsigned)(n - tile_base);
    const unsigned off0 = tile_offset_of_item0<ITEMS, VWL>();  // in-tile index of this thread's item 0; item i is off0 + VWL*i
    unsigned *wt = tab + vwarp * kRadixSize;

    // key[] already holds this tile's keys (loaded by the caller / prefetched during the previous tile).

    // ---- rank inside the warp (stable: item-major, lane-minor == memory order) ----
    static_assert(VB == 0 || RANK != kRankTwoSweep, "the two-sweep ranking keeps no ranks for a payload");
    // with a payload, rank[i] later also carries (in its upper half) the digit of the sorted-tile position this thread
    // writes in round i: one register per item instead of two
    static_assert(TILE <= 65536, "in-tile positions are 16-bit");
    typename std::conditional<(VB > 0), unsigned, unsigned short>::type rank[RANK == kRankTwoSweep ? 1 : ITEMS];
    unsigned dpk[IDENT == kDigitSplit ? (ITEMS + 7) / 8 : 1] = {};  // splitter mode: packed buckets of this thread's keys
    if constexpr (RANK == kRankAtomicOr) {
        const unsigned hbit = 1u << (lane & 15u);
        const unsigned hlt = hbit - 1u;
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            unsigned d = pass_digit<K, IDENT>(key[i], shift, tf);
            if (!FULL && off0 + i * VWL >= valid) d = kRadixSize - 1;  // padding sorts last within the tile
            atomicOr(&wt[d], hbit);
            __syncwarp();
            const unsigned e = wt[d];  // {keys of digit d in earlier rounds : 16 | peers in this round : 16}
            const unsigned below = __popc(e & hlt);
            rank[i] = (unsigned short)((e >> 16) + below);
            __syncwarp();
            if (((e & 0xffffu) >> (lane & 15u)) == 1u) wt[d] = (e & 0xffff0000u) + ((below + 1u) << 16);  // highest peer
            __syncwarp();
        }
    } else if constexpr (RANK == kRankBallot) {
        static_assert(IDENT == kDigitSplit || RANK != kRankBallot, "ballot ranking covers the splitter buckets only");
        const unsigned lt = (1u << lane) - 1u;
        unsigned cnt = 0;  // lane b < 8: keys of bucket b seen so far by this warp; lane 8: padding
        // lane_inv[b] = 0 if bit b of my lane id is set, else all-ones: (ballot ^ lane_inv[b]) = lanes agreeing with it
        const unsigned lane_inv0 = (lane & 1u) ? 0u : ~0u, lane_inv1 = (lane & 2u) ? 0u : ~0u, lane_inv2 = (lane & 4u) ? 0u : ~0u,
                       lane_inv3 = (lane & 8u) ? 0u : ~0u;
        const int nbits = tf.nsplit > 3 ? 3 : (tf.nsplit > 1 ? 2 : 1);  // uniform: bits that tell the buckets apart
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            unsigned d = pass_digit<K, IDENT>(key[i], shift, tf);
            const bool pad = !FULL && off0 + i * VWL >= valid;
            if (pad) d = 8;
            dpk[i / 8] |= d << (4 * (i % 8));  // the bucket costs up to 7 compares: evaluate it once, keep 4 bits per key
            // peers = lanes whose key has my key's bucket; mine = lanes whose key's bucket is my lane id
            unsigned bal = __ballot_sync(0xffffffffu, d & 1u);
            unsigned peers = bal ^ ((d & 1u) - 1u), mine = bal ^ lane_inv0;
            if (nbits > 1) {
                bal = __ballot_sync(0xffffffffu, d & 2u);
                peers &= bal ^ (((d >> 1) & 1u) - 1u);
                mine &= bal ^ lane_inv1;
                if (nbits > 2) {
                    bal = __ballot_sync(0xffffffffu, d & 4u);
                    peers &= bal ^ (((d >> 2) & 1u) - 1u);
                    mine &= bal ^ lane_inv2;
                }
            }
            if (!FULL) {
                bal = __ballot_sync(0xffffffffu, d & 8u);
                peers &= bal ^ (((d >> 3) & 1u) - 1u);
                mine &= bal ^ lane_inv3;
            }
            const unsigned before = __shfl_sync(0xffffffffu, cnt, d);
            rank[i] = (unsigned short)(before + __popc(peers & lt));
            cnt += __popc(mine);  // (lanes that stand for no bucket count garbage that is never read)
        }
        if ((int)lane <= tf.nsplit) wt[lane] = cnt;  // (higher lanes stand for no bucket: with fewer ballot bits they hold aliases)
        if (!FULL && lane == 8) wt[kRadixSize - 1] = cnt;  // padding sorts last within the tile, as in the other modes
    } else if constexpr (RANK == kRankTwoSweep) {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            unsigned d = pass_digit<K, IDENT>(key[i], shift, tf);
            if (!FULL && off0 + i * VWL >= valid) d = kRadixSize - 1;
            atomicAdd(&wt[d], 1u);  // count only; the position comes from the second sweep
        }
    }
    __syncthreads();

    // ---- per digit: exclusive prefix over warps, tile count, exclusive scan over the 256 digit values ----
    unsigned count = 0;
    if (tid < kRadixSize) {
#pragma unroll
        for (int v = 0; v < VWARPS; v++) count += tab[v * kRadixSize + tid] >> CSHIFT;
    }
    unsigned incl = count;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
        if ((int)lane >= off) incl += o;
    }
    if (tid < kRadixSize && lane == 31) misc[2 + warp] = incl;
    __syncthreads();
    unsigned my_start = 0;
    if (tid < kRadixSize) {
        unsigned add = 0;
#pragma unroll
        for (int w = 0; w < kRadixSize / 32; w++) add += (w < (int)warp) ? misc[2 + w] : 0u;
        my_start = incl - count + add;
        // position of a key in the sorted tile = entry of (its warp, its digit) + its rank inside the warp
        unsigned run = my_start;
#pragma unroll
        for (int v = 0; v < VWARPS; v++) {
            const unsigned c = tab[v * kRadixSize + tid] >> CSHIFT;
            tab[v * kRadixSize + tid] = run << CSHIFT;
            run += c;
        }
        if (!FULL && tid == kRadixSize - 1) count -= (unsigned)TILE - valid;  // drop the padding from the published count
        const unsigned status = (tile == 0) ? kLbInclusive : kLbPartial;
        st_relaxed_u64(lookback + tile * kRadixSize + tid, ((unsigned long long)((epoch << 2) | status) << 32) | count);
    }
    __syncthreads();

    // ---- reorder the tile through shared memory ----
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        unsigned d;
        if constexpr (IDENT == kDigitSplit) d = (dpk[i / 8] >> (4 * (i % 8))) & 15u;
        else d = pass_digit<K, IDENT>(key[i], shift, tf);
        if (!FULL && off0 + i * VWL >= valid) d = kRadixSize - 1;
        unsigned pos;
        if constexpr (RANK == kRankTwoSweep) {
            pos = atomicAdd(&wt[d], 1u);  // same atomics, same order as the counting sweep: offset + rank
        } else {
            pos = (wt[d] >> CSHIFT) + rank[i];
            rank[i] = (unsigned short)pos;  // (zero-extends when the element type is 32-bit)
        }
        keys_sorted[pos] = key[i];
    }

    // ---- prefetch the next tile's keys into the (now dead) key registers: the loads fly during the look-back and
    // the store phase of this tile, so a tile never waits for its own input ----
    {
        const size_t next = *next_tile;  // drawn by thread 0 at the top of this tile (see onesweep_pass)
        if (next < num_tiles) load_tile_keys<K, THREADS, ITEMS, PassSmem<K, VB, THREADS, ITEMS, RANK>::VWL>(keys_in, n, next, key);
    }

    // ---- decoupled look-back (threads 0..255, one digit each).  The keys already sit in shared memory, so
    // the key registers are dead here and the batched descriptor loads do not raise register pressure ----
    if (tid < kRadixSize) {
        unsigned excl = 0;
        if (tile > 0) {
            constexpr int LB = LBATCH;  // independent descriptor loads in flight per step of the walk
            long long j = (long long)tile - 1;
            bool done = false;
            while (!done) {
                unsigned long long w[LB];
#pragma unroll
                for (int b = 0; b < LB; b++) {
                    const long long idx = j - b;
                    w[b] = idx >= 0 ? ld_relaxed_u64(lookback + (size_t)idx * kRadixSize + tid)
                                    : ((unsigned long long)((epoch << 2) | kLbInclusive) << 32);
                }
                int consumed = 0;
#pragma unroll
                for (int b = 0; b < LB; b++) {
                    if (!done && consumed == b) {
                        const unsigned tag = (unsigned)(w[b] >> 32);
                        if ((tag >> 2) == epoch) {  // published
                            excl += (unsigned)w[b];
                            consumed = b + 1;
                            done = (tag & 3u) == kLbInclusive;
                        }
                    }
                }
                j -= consumed;
                if (consumed == 0) __nanosleep(40);  // nothing new was published: back off before polling again
            }
            st_relaxed_u64(lookback + tile * kRadixSize + tid,
                           ((unsigned long long)((epoch << 2) | kLbInclusive) << 32) | (unsigned)(excl + count));
        }
        if constexpr (IDENT != kDigitSplit) {
            out_base[tid] = __ldg(digit_base + tid) + excl - my_start;  // global index = out_base[d] + position in the sorted tile
        } else {
            // splitter mode: every bucket has its own destination (possibly another GPU's memory).  out_base is not
            // needed as an index table, so it holds, per bucket: the byte address of the first element of this tile's
            // run (keys, values), the run's start inside the sorted tile and its length.
            if (tid <= kMaxSplitters) {
                const unsigned long long first = (unsigned long long)__ldg(digit_base + tid) + excl;  // elements before the run
                unsigned long long *addr = reinterpret_cast<unsigned long long *>(out_base);
                addr[tid] = tf.dst_keys[tid] + first * sizeof(K);
                if constexpr (VB > 0) addr[kMaxSplitters + 1 + tid] = tf.dst_vals[tid] + first * (unsigned long long)VB;
                out_base[32 + tid] = my_start;
                out_base[40 + tid] = count;
            }
        }
    }
    __syncthreads();

    // the digit tables are dead from here on (ranking and reorder are complete in every warp): zero them for the next
    // tile now, overlapped with the memory-bound write phase, instead of behind an extra barrier at the top of the loop
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        for (unsigned i = tid; i < L::kWarpTab / 16; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }

    // ---- write keys: consecutive threads -> consecutive addresses inside each digit run ----
    if constexpr (IDENT == kDigitSplit) {
        // Few, long runs (<= 8 buckets).  Walk them one by one with the lanes aligned to the DESTINATION: lane l of a
        // warp always writes an address whose element index is l mod 32, so every warp store is one aligned line
        // instead of straddling two -- over NVLink a straddling store is two packets (measured: 7.2-8.0 ms -> see
        // DESIGN.md for the 8-rank exchange pass).
        const unsigned long long *addr = reinterpret_cast<const unsigned long long *>(out_base);
        for (int b = 0; b <= tf.nsplit; b++) {
            const unsigned s = out_base[32 + b], c = out_base[40 + b];
            if (c == 0) continue;
            const unsigned long long a0 = addr[b];
            const unsigned mis = (unsigned)(a0 / sizeof(K)) & 31u;  // element index of the run start inside its line
            for (unsigned v = tid; v < c + mis; v += THREADS) {
                if (v >= mis) {
                    const unsigned j = v - mis;
                    *reinterpret_cast<K *>(a0 + (unsigned long long)j * sizeof(K)) = keys_sorted[s + j];
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned p = i * THREADS + tid;
            if (FULL || p < valid) {
                const K k = keys_sorted[p];
                const unsigned d = pass_digit<K, IDENT>(k, shift, tf);
                if constexpr (VB > 0) rank[i] |= d << 16;
                keys_out[(size_t)(out_base[d] + p)] = k;
            }
        }
    }

    if constexpr (VB > 0) {
        const V *vals_in = reinterpret_cast<const V *>(vals_in_v);
        V *vals_out = reinterpret_cast<V *>(vals_out_v);
        V *vals_sorted = reinterpret_cast<V *>(elem_buf);
        V val[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned t = off0 + i * VWL;
            if (FULL || t < valid) val[i] = vals_in[tile_base + t];
        }
        __syncthreads();  // everyone is done reading keys_sorted
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned t = off0 + i * VWL;
            if (FULL || t < valid) vals_sorted[rank[i] & 0xffffu] = val[i];
        }
        __syncthreads();
        if constexpr (IDENT == kDigitSplit) {
            const unsigned long long *addr = reinterpret_cast<const unsigned long long *>(out_base);
            for (int b = 0; b <= tf.nsplit; b++) {
                const unsigned s = out_base[32 + b], c = out_base[40 + b];
                if (c == 0) continue;
                const unsigned long long a0 = addr[kMaxSplitters + 1 + b];
                const unsigned mis = (unsigned)(a0 / VB) & 31u;
                for (unsigned v = tid; v < c + mis; v += THREADS) {
                    if (v >= mis) {
                        const unsigned j = v - mis;
                        *reinterpret_cast<V *>(a0 + (unsigned long long)j * VB) = vals_sorted[s + j];
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < ITEMS; i++) {
                const unsigned p = i * THREADS + tid;
                if (FULL || p < valid) vals_out[(size_t)(out_base[rank[i] >> 16] + p)] = vals_sorted[p];
            }
        }
    }
}

// Tile ids come from a ticket (atomic counter), drawn one tile ahead so that the next tile's keys can be prefetched:
// a tile's predecessors were all drawn earlier, by CTAs that are running, so the look-back never waits for a CTA that
// has not been scheduled -- forward progress does not depend on the whole grid being resident (other streams may hold
// SMs).  `gate`: fallback launches of a speculative sort run only if the verification flag is set.
__device__ __forceinline__ unsigned draw_tile(unsigned long long *ticket, unsigned long long ticket_base, size_t num_tiles)
{
    const unsigned long long t = atomicAdd(ticket, 1ull) - ticket_base;
    return t < num_tiles ? (unsigned)t : 0xffffffffu;
}

template <typename K, int VB, int THREADS, int ITEMS, int LBATCH, int RANK, int IDENT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
onesweep_pass(const K *__restrict__ keys_in, K *__restrict__ keys_out, const void *__restrict__ vals_in_v,
              void *__restrict__ vals_out_v, const unsigned *__restrict__ digit_base, unsigned long long *lookback,
              unsigned epoch, size_t n, size_t num_tiles, int shift, unsigned long long *ticket, unsigned long long ticket_base,
              const int *__restrict__ gate, const __grid_constant__ Transform tf, const unsigned *__restrict__ hot)
{
    // Persistent CTAs: the grid is sized to the number of CTAs that fit the device; the tiles in flight form one
    // contiguous window of the input (their scattered writes merge in L2), and the next tile's keys are fetched while
    // the current tile is still being processed.
    typedef PassSmem<K, VB, THREADS, ITEMS, RANK> L;
    static_assert(THREADS >= kRadixSize && THREADS % 32 == 0, "one look-back thread per digit value");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    volatile unsigned *slot = reinterpret_cast<unsigned *>(smem_raw + L::kWarpTab) + kRadixSize;  // misc[0..1]: tile ids

    if (gate && *gate == 0) {  // not needed: keep the ticket counter in step with the host's reservation and leave
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(ticket, (unsigned long long)num_tiles + gridDim.x);
        return;
    }
    if (IDENT != kDigitSplit && hot && __ldg(hot + 1) == kConstDigit) {
        // every key has the same value of this digit: the stable pass is the identity permutation -- a streaming copy
        // (keys stay raw in memory between the passes of this kernel, so nothing is transformed)
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(ticket, (unsigned long long)num_tiles + gridDim.x);
        typedef typename value_type<VB>::type V;
        const size_t g = (size_t)blockIdx.x * THREADS + threadIdx.x, nthreads = (size_t)gridDim.x * THREADS;
        stream_copy(keys_in, keys_out, n, g, nthreads, [](K k) { return k; });
        if constexpr (VB > 0) stream_copy(reinterpret_cast<const V *>(vals_in_v), reinterpret_cast<V *>(vals_out_v), n, g, nthreads, [](V v) { return v; });
        return;
    }
    if (threadIdx.x == 0) slot[0] = draw_tile(ticket, ticket_base, num_tiles);
    // the digit tables start out zero; every tile zeroes them again once it is done with them (during its write phase)
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        for (unsigned i = threadIdx.x; i < L::kWarpTab / 16; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    size_t tile = slot[0];
    K key[ITEMS];
    if (tile < num_tiles) load_tile_keys<K, THREADS, ITEMS, L::VWL>(keys_in, n, tile, key);
    for (unsigned it = 0; tile < num_tiles; ++it) {
        volatile unsigned *next = slot + ((it + 1) & 1u);
        if (threadIdx.x == 0) *next = draw_tile(ticket, ticket_base, num_tiles);  // read after the tile's first barrier
        if ((tile + 1) * (size_t)L::TILE <= n)
            pass_tile<K, VB, THREADS, ITEMS, LBATCH, RANK, IDENT, true>(keys_in, keys_out, vals_in_v, vals_out_v, digit_base, lookback,
                                                                         epoch, n, shift, tf, tile, smem_raw, key,
                                                                         num_tiles, next);
        else
            pass_tile<K, VB, THREADS, ITEMS, LBATCH, RANK, IDENT, false>(keys_in, keys_out, vals_in_v, vals_out_v, digit_base, lookback,
                                                                          epoch, n, shift, tf, tile, smem_raw, key,
                                                                          num_tiles, next);
        __syncthreads();  // all stores of this tile issued, shared memory free for the next one
        tile = *next;
    }
}

// ---- helpers for payloads whose size is not 1/2/4/8/16 bytes: sort (key, index), then gather ----
__global__ void iota_u32_kernel(unsigned *p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = (unsigned)i;
}
__global__ void gather_bytes_kernel(const unsigned char *src, const unsigned *idx, unsigned char *dst, size_t n, size_t w)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n * w; i += stride) {
        const size_t e = i / w, b = i - e * w;
        dst[i] = src[(size_t)idx[e] * w + b];
    }
}

// serial_insertion_sort(_by_key): algorithm/detail/insertion_sort.hpp:25-159 -- one thread, native compare
template <typename T>
__global__ void insertion_sort_kernel(T *keys, size_t n, int greater, unsigned char *vals, size_t vb)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (size_t i = 1; i < n; i++) {
        const T key = keys[i];
        unsigned char tmp[256];
        if (vals) for (size_t b = 0; b < vb; b++) tmp[b] = vals[i * vb + b];
        size_t pos = i;
        while (pos > 0 && (greater ? (key > keys[pos - 1]) : (key < keys[pos - 1]))) {
            keys[pos] = keys[pos - 1];
            if (vals) for (size_t b = 0; b < vb; b++) vals[pos * vb + b] = vals[(pos - 1) * vb + b];
            pos--;
        }
        keys[pos] = key;
        if (vals) for (size_t b = 0; b < vb; b++) vals[pos * vb + b] = tmp[b];
    }
}

// ---- launch plumbing -----------------------------------------------------------------------------------
constexpr int default_min_blocks(int threads) { return threads <= 256 ? 3 : (threads <= 512 ? 2 : 1); }

// per (kernel instantiation, device) launch facts, computed once; safe to race (all writers store the same values)
struct PassLaunchCache {
    std::atomic<int> per_sm[64];
    PassLaunchCache() { for (auto &v : per_sm) v.store(0, std::memory_order_relaxed); }
};

template <typename K, int VB, int THREADS, int ITEMS, int LBATCH, int RANK, int IDENT, int MINB = default_min_blocks(THREADS)>
static int launch_pass_impl(StreamState *st, const void *kin, void *kout, const void *vin, void *vout, const unsigned *base,
                            unsigned long long *lookback, size_t n, int shift, const Transform &tf, const int *gate = nullptr,
                            const unsigned *hot = nullptr)
{
    typedef PassSmem<K, VB, THREADS, ITEMS, RANK> L;
    static_assert((IDENT == kDigitSplit) == (RANK == kRankBallot), "the splitter pass and the ballot ranking go together");
    constexpr size_t kSmemBytes = L::kBytes;
    auto kernel = onesweep_pass<K, VB, THREADS, ITEMS, LBATCH, RANK, IDENT, MINB>;
    static PassLaunchCache cache;  // CTAs of this kernel that fit one SM (0 = not configured yet on that device)
    int per_sm = st->device < 64 ? cache.per_sm[st->device].load(std::memory_order_acquire) : 0;
    if (per_sm == 0) {
        BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        BCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, kSmemBytes));
        if (per_sm < 1) per_sm = 1;
        if (st->device < 64) cache.per_sm[st->device].store(per_sm, std::memory_order_release);
    }
    const size_t tiles = (n + L::TILE - 1) / L::TILE;
    size_t grid = (size_t)st->sm_count * (size_t)per_sm;
    if (grid > tiles) grid = tiles;
    unsigned epoch;
    BCB_TRY(next_epoch(st, kArenaPacked, &epoch));
    const unsigned long long ticket_base = ticket_reserve(st, tiles + grid);  // every CTA draws one void ticket
    // (gated fallback launches of a speculative sort return at once: timed as "other", not as pass kernels)
    LaunchTimer timer(st, gate ? BCB_K_OTHER : (IDENT == kDigitSplit ? BCB_K_EXCHANGE_PASS : BCB_K_ONESWEEP_PASS));
    kernel<<<(unsigned)grid, THREADS, kSmemBytes, st->stream>>>((const K *)kin, (K *)kout, vin, vout, base, lookback, epoch, n, tiles,
                                                               shift, st->control + kControlTicket, ticket_base, gate, tf, hot);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// Speculative ranking for large keys-only sorts (32- and 64-bit keys).
// The two-sweep pass kernels (kRankTwoSweep here, onesweep_ws in radix_pass_ws.cu) are much faster than the atomic-OR
// one but stable only if same-address shared atomics of one warp instruction are applied in lane order (true on every
// B200 measured, not promised by CUDA).  For a KEYS-ONLY sort that assumption can be CHECKED after the fact: every pass
// is a permutation whatever order the atomics took, so the output is the correct result if and only if it is sorted by
// the transformed key.  sort_typed therefore runs the fast passes, verifies sortedness in one extra read (4 B/key)
// into a DEVICE flag, and enqueues the deterministic sort of the same buffer behind it with every kernel gated on
// that flag (the launches return at once when the flag is clear) -- re-sorting a permutation of the input gives the
// same bytes.  Nothing waits on the host: the sort stays enqueue-and-return.  Key-value sorts never speculate
// (stability of the payload cannot be verified from the keys), and neither do descending float sorts (their key
// transform is not injective, see sort_typed).
//   BCB_SORT_SPECULATIVE=0       always use the deterministic atomic-OR kernel
//   BCB_SORT_FORCE_FALLBACK=1    test hook: treat every verification as failed
//   BCB_SORT_WS=0                keep the r01 two-sweep kernel for large sorts (A/B comparison)
//   BCB_SORT_HOT=0               warp-specialised kernel: no ballot ranking of frequent digit values (A/B comparison)
//   BCB_SORT_SPEC_MIN_LOG2=k, BCB_SORT_WS_MIN_LOG2=k   test hooks: move the size thresholds below
// Size thresholds, measured on B200 (2^k uint32 keys, ms per sort: deterministic / two-sweep + verification / onesweep_ws +
// verification): 2^23 0.26 / - / 0.33; 2^24 0.45 / 0.34 / -; 2^25 0.70 / 0.52 / 0.61; 2^26 1.19 / 0.90 / 0.95; 2^27 2.19 / - / 1.63;
// 2^30 19 / 12.7 / 11.2.  Verification and the gated fallback launches cost ~70 us, the warp-specialised pipeline needs
// >= ~20 tiles per SM to fill.
constexpr int kSpeculativeMinLog2 = 24;
constexpr int kWsMinLog2 = 27;
struct SortEnv {
    bool speculative, force_fallback, ws, hot;
    size_t spec_min, ws_min;
    SortEnv()
    {
        auto log2_of = [](const char *name, int dflt) {
            const char *v = std::getenv(name);
            const int k = v ? std::atoi(v) : dflt;
            return (size_t)1 << (k < 10 ? 10 : (k > 40 ? 40 : k));
        };
        spec_min = log2_of("BCB_SORT_SPEC_MIN_LOG2", kSpeculativeMinLog2);
        ws_min = log2_of("BCB_SORT_WS_MIN_LOG2", kWsMinLog2);
        const char *e = std::getenv("BCB_SORT_SPECULATIVE");
        speculative = !(e && e[0] == '0');
        e = std::getenv("BCB_SORT_FORCE_FALLBACK");
        force_fallback = e && e[0] == '1';
        e = std::getenv("BCB_SORT_WS");
        ws = !(e && e[0] == '0');
        e = std::getenv("BCB_SORT_HOT");  // 0: no ballot ranking of frequent digit values (A/B comparison)
        hot = !(e && e[0] == '0');
    }
};
static const SortEnv &sort_env()
{
    static const SortEnv env;  // thread-safe initialisation (C++11)
    return env;
}

// sortedness by the transformed key (the order the sort is defined by), 128-bit loads
template <typename K, bool IDENT>
__device__ __forceinline__ typename key_traits<K>::U verify_key(K raw, const Transform &tf)
{
    typedef typename key_traits<K>::U U;
    if constexpr (IDENT) return (U)raw;
    else return (U)transformed_key<K>(raw, tf);
}

// coalesced 128-bit loads: a warp takes kVerifyUnroll x 32 consecutive vectors per iteration (lane l reads vectors
// base + j*32 + l, all loads issued before the first compare); the seam to the next vector comes from the neighbouring
// lane by shuffle (lane 31: from lane 0 of the next group), only the seam after the warp's last vector touches memory again
constexpr int kVerifyUnroll = 4;
template <typename K, bool IDENT>
__global__ void __launch_bounds__(256) verify_sorted_kernel(const K *__restrict__ keys, size_t n, Transform tf, int *flag, int force)
{
    typedef typename key_traits<K>::U U;
    constexpr int VEC = 16 / sizeof(K);
    constexpr int UNR = kVerifyUnroll;
    const unsigned lane = threadIdx.x & 31u;
    const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t warp_id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int bad = 0;
    size_t done = 0;
    if ((((uintptr_t)keys) & 15) == 0) {
        const size_t nvec = n / VEC;
        for (size_t base = warp_id * (32 * UNR); base < nvec; base += warps * (32 * UNR)) {  // warp-uniform trip count
            uint4 x[UNR];
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const size_t v = base + j * 32 + lane;
                x[j] = make_uint4(0, 0, 0, 0);
                if (v < nvec) x[j] = ld_stream_v4(keys + v * VEC);
            }
            // first key after the warp's chunk (the next chunk's first vector); lane 31 only
            const size_t after = (base + (size_t)UNR * 32) * VEC;
            U after_key = 0;
            if (lane == 31 && after < nvec * VEC) after_key = verify_key<K, IDENT>(__ldg(keys + after), tf);
            U first[UNR], last[UNR];
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const K *e = reinterpret_cast<const K *>(&x[j]);
                U k[VEC];
#pragma unroll
                for (int i = 0; i < VEC; i++) k[i] = verify_key<K, IDENT>(e[i], tf);
                const bool in = base + j * 32 + lane < nvec;
#pragma unroll
                for (int i = 1; i < VEC; i++) bad |= in && (k[i - 1] > k[i]);
                first[j] = k[0];
                last[j] = k[VEC - 1];
            }
#pragma unroll
            for (int j = 0; j < UNR; j++) {
                const size_t v = base + j * 32 + lane;
                U next = __shfl_down_sync(0xffffffffu, first[j], 1);
                const U wrap = __shfl_sync(0xffffffffu, j + 1 < UNR ? first[(j + 1) % UNR] : after_key, j + 1 < UNR ? 0 : 31);
                if (lane == 31) next = wrap;
                bad |= (v + 1 < nvec) && (last[j] > next);  // the seam into the scalar tail is checked below
            }
        }
        done = nvec * VEC;
    }
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (done ? done - 1 : 0) + gid; i + 1 < n; i += stride)  // scalar tail, including the seam into it
        bad |= verify_key<K, IDENT>(__ldg(keys + i), tf) > verify_key<K, IDENT>(__ldg(keys + i + 1), tf);
    if (bad || force) *flag = 1;
}

template <typename K, int VB, int THREADS, int ITEMS, int LBATCH = kLookbackBatch, int MINB = default_min_blocks(THREADS)>
static int launch_pass(StreamState *st, const void *kin, void *kout, const void *vin, void *vout, const unsigned *base,
                       unsigned long long *lookback, size_t n, int shift, const Transform &tf, const int *gate, const unsigned *hot)
{
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0;  // unsigned ascending keys: the digit is a plain bit field
    return ident ? launch_pass_impl<K, VB, THREADS, ITEMS, LBATCH, kRankAtomicOr, kDigitIdent, MINB>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, gate, hot)
                 : launch_pass_impl<K, VB, THREADS, ITEMS, LBATCH, kRankAtomicOr, kDigitTransform, MINB>(st, kin, kout, vin, vout, base, lookback, n, shift, tf, gate, hot);
}

// the speculative two-sweep pass (keys only, 32- and 64-bit keys)
template <typename K, int THREADS, int ITEMS, int LBATCH, int MINB>
static int launch_two_sweep(StreamState *st, const void *kin, void *kout, const unsigned *base, unsigned long long *lookback, size_t n,
                            int shift, const Transform &tf, const unsigned *hot)
{
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0;
    return ident ? launch_pass_impl<K, 0, THREADS, ITEMS, LBATCH, kRankTwoSweep, kDigitIdent, MINB>(st, kin, kout, nullptr, nullptr, base, lookback, n, shift, tf, nullptr, hot)
                 : launch_pass_impl<K, 0, THREADS, ITEMS, LBATCH, kRankTwoSweep, kDigitTransform, MINB>(st, kin, kout, nullptr, nullptr, base, lookback, n, shift, tf, nullptr, hot);
}

// tile shapes of the deterministic pass: (key bytes, value bytes) -> THREADS x ITEMS
template <typename K, int VB> struct PassConfig { static constexpr int THREADS = 384, ITEMS = 16; };
template <> struct PassConfig<unsigned, 0> { static constexpr int THREADS = 384, ITEMS = 20; };
template <> struct PassConfig<unsigned short, 0> { static constexpr int THREADS = 384, ITEMS = 20; };
template <> struct PassConfig<unsigned char, 0> { static constexpr int THREADS = 384, ITEMS = 20; };
template <> struct PassConfig<unsigned long long, 0> { static constexpr int THREADS = 384, ITEMS = 12; };
template <typename K> struct PassConfig<K, 8> { static constexpr int THREADS = 384, ITEMS = 12; };
template <typename K> struct PassConfig<K, 16> { static constexpr int THREADS = 384, ITEMS = 8; };
template <> struct PassConfig<unsigned long long, 4> { static constexpr int THREADS = 384, ITEMS = 12; };

// tile shapes of the two-sweep pass: without rank registers 32 keys per thread fit 80 registers (2 CTAs of 384 threads
// per SM); the per-tile costs (digit scan, look-back, barriers) are spread over 12288 keys.  Measured on B200,
// 2^30 u32 keys, 4 passes: 384x20 14.0 ms, 384x24 12.5 ms, 384x32 11.0 ms, 384x40 11.5 ms, 512x24 11.6 ms.
template <typename K> struct SpecConfig { static constexpr int THREADS = 384, ITEMS = 32, LB = 4, MINB = 2; };
// (2^28 u64 keys, 8 passes: 384x12 10.9 ms, 384x16 9.2 ms, 384x20 8.7 ms, 384x24 8.7 ms)
template <> struct SpecConfig<unsigned long long> { static constexpr int THREADS = 384, ITEMS = 20, LB = 4, MINB = 2; };

// which pass kernel a sort uses: the r01 deterministic atomic-OR kernel, the r01 speculative two-sweep kernel, the
// warp-specialised kernel with the speculative ranking (keys only), or with the deterministic ranking (payloads, keys
// with a non-injective transform)
enum { kPassDeterministic = 0, kPassTwoSweep = 1, kPassWs = 2, kPassWsDet = 3 };

template <typename K, int VB>
static size_t tile_size_for(int pass_kind)
{
    if (pass_kind == kPassWs || pass_kind == kPassWsDet) return ws_tile_size((int)sizeof(K), VB, pass_kind == kPassWsDet);
    if constexpr (VB == 0 && (sizeof(K) == 4 || sizeof(K) == 8)) {
        if (pass_kind == kPassTwoSweep) return (size_t)SpecConfig<K>::THREADS * SpecConfig<K>::ITEMS;
    }
    return (size_t)PassConfig<K, VB>::THREADS * PassConfig<K, VB>::ITEMS;
}

template <typename K, int VB>
static int run_pass(StreamState *st, const void *kin, void *kout, const void *vin, void *vout, const unsigned *base,
                    unsigned long long *lookback, size_t n, int shift, const Transform &tf, int pass_kind, const int *gate,
                    const unsigned *hot = nullptr)
{
    if (pass_kind == kPassWs || pass_kind == kPassWsDet) {
        // keys travel between the passes in transformed form (one transform in the first pass, its inverse in the last)
        // unless the transform cannot be inverted (descending float / double): then every pass transforms for the digit only
        const bool ident = (tf.nm | tf.xc | tf.fa) == 0, injective = !(tf.fa != 0 && tf.nm != 0);
        constexpr int kLastShift = ((int)sizeof(K) - 1) * kRadixBits;
        const int xf = ident ? kXfNone : (!injective ? kXfBoth : (shift == 0 ? kXfIn : (shift == kLastShift ? kXfOut : kXfNone)));
        return ws_launch_pass(st, (int)sizeof(K), kin, kout, vin, vout, VB, base, lookback, n, shift, tf, xf, pass_kind == kPassWsDet,
                              nullptr, nullptr, 0, hot);
    }
    if constexpr (VB == 0 && (sizeof(K) == 4 || sizeof(K) == 8)) {
        if (pass_kind == kPassTwoSweep)
            return launch_two_sweep<K, SpecConfig<K>::THREADS, SpecConfig<K>::ITEMS, SpecConfig<K>::LB, SpecConfig<K>::MINB>(
                st, kin, kout, base, lookback, n, shift, tf, hot);
    }
    return launch_pass<K, VB, PassConfig<K, VB>::THREADS, PassConfig<K, VB>::ITEMS>(st, kin, kout, vin, vout, base, lookback, n,
                                                                                     shift, tf, gate, hot);
}

// ---- multi-GPU partition pass: bucket histogram by splitters ---------------------------------------------
// ge[j] = number of keys whose transformed key is >= splitter j (7 compare-and-add per key, 128-bit loads); the bucket
// sizes are the differences: bucket b = ge[b-1] - ge[b] with ge[-1] = n, ge[nsplit] = 0
template <typename K>
__global__ void __launch_bounds__(256) split_histogram(const K *__restrict__ keys, size_t n, unsigned *__restrict__ hist, Transform tf)
{
    typedef typename key_traits<K>::U U;
    constexpr int VEC = 16 / sizeof(K);
    unsigned ge[kMaxSplitters];
    U sp[kMaxSplitters];
#pragma unroll
    for (int j = 0; j < kMaxSplitters; j++) { ge[j] = 0; sp[j] = (U)tf.split[j]; }
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if ((((uintptr_t)keys) & 15) == 0) {
        const size_t nvec = n / VEC;
        size_t v = gid;
        for (; v + stride < nvec; v += 2 * stride) {  // two independent 128-bit loads in flight
            const uint4 x0 = ld_stream_v4(keys + v * VEC), x1 = ld_stream_v4(keys + (v + stride) * VEC);
            const K *e0 = reinterpret_cast<const K *>(&x0), *e1 = reinterpret_cast<const K *>(&x1);
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const U t0 = (U)transformed_key<K>(e0[i], tf), t1 = (U)transformed_key<K>(e1[i], tf);
#pragma unroll
                for (int j = 0; j < kMaxSplitters; j++) ge[j] += (unsigned)(t0 >= sp[j]) + (unsigned)(t1 >= sp[j]);
            }
        }
        for (; v < nvec; v += stride) {
            const uint4 x0 = ld_stream_v4(keys + v * VEC);
            const K *e0 = reinterpret_cast<const K *>(&x0);
#pragma unroll
            for (int i = 0; i < VEC; i++) {
                const U t0 = (U)transformed_key<K>(e0[i], tf);
#pragma unroll
                for (int j = 0; j < kMaxSplitters; j++) ge[j] += (unsigned)(t0 >= sp[j]);
            }
        }
        done = nvec * VEC;
    }
    for (size_t i = done + gid; i < n; i += stride) {
        const U t = (U)transformed_key<K>(__ldg(keys + i), tf);
#pragma unroll
        for (int j = 0; j < kMaxSplitters; j++) ge[j] += (unsigned)(t >= sp[j]);
    }
    __shared__ unsigned block_ge[kMaxSplitters];
    if (threadIdx.x < kMaxSplitters) block_ge[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxSplitters; j++) {
        const unsigned v = __reduce_add_sync(0xffffffffu, ge[j]);
        if ((threadIdx.x & 31u) == 0 && v) atomicAdd(block_ge + j, v);
    }
    __syncthreads();
    if ((int)threadIdx.x < tf.nsplit && block_ge[threadIdx.x]) atomicAdd(hist + kRadixSize + threadIdx.x, block_ge[threadIdx.x]);
}

// hist[256 + j] = ge[j]  ->  hist[b] = size of bucket b
__global__ void split_counts_kernel(unsigned *hist, unsigned n, int nsplit)
{
    const int b = threadIdx.x;
    if (b > nsplit) return;
    const unsigned hi = b == 0 ? n : hist[kRadixSize + b - 1];
    const unsigned lo = b == nsplit ? 0u : hist[kRadixSize + b];
    hist[b] = hi - lo;
}

// bucket sizes of the splitter partition: hist[0..nsplit] on the device, copied to counts_host (blocks)
template <typename K>
static int partition_counts_typed(StreamState *st, const void *kin, size_t n, const Transform &tf, unsigned long long *counts_host)
{
    unsigned *hist = st->hist;
    BCB_CUDA_TRY(cudaMemsetAsync(hist, 0, 2 * kRadixSize * sizeof(unsigned), st->stream));
    size_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const size_t cap = (size_t)st->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    {
        LaunchTimer timer(st, BCB_K_RADIX_HISTOGRAM);
        split_histogram<K><<<(unsigned)blocks, 256, 0, st->stream>>>((const K *)kin, n, hist, tf);
    }
    BCB_CUDA_TRY(cudaGetLastError());
    split_counts_kernel<<<1, 32, 0, st->stream>>>(hist, (unsigned)n, tf.nsplit);
    BCB_CUDA_TRY(cudaGetLastError());
    unsigned host_counts[kMaxSplitters + 1];
    BCB_CUDA_TRY(cudaMemcpyAsync(host_counts, hist, sizeof(host_counts), cudaMemcpyDeviceToHost, st->stream));
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    for (int j = 0; j <= tf.nsplit; j++) counts_host[j] = host_counts[j];
    return BCB_SUCCESS;
}

// one stable onesweep pass whose "digit" is the splitter bucket; bucket b is written contiguously from tf.dst_keys[b]
// (+ base[b] elements).  Asynchronous.
template <typename K, int VB, int THREADS, int ITEMS, int MINB>
static int partition_scatter_shape(StreamState *st, const void *kin, const void *vin, size_t n, const Transform &tf, const unsigned *base)
{
    const size_t tile = (size_t)THREADS * ITEMS;
    const size_t tiles = (n + tile - 1) / tile;
    void *lb;
    BCB_TRY(lookback_reserve(st, kArenaPacked, tiles * kRadixSize * sizeof(unsigned long long), &lb));
    return launch_pass_impl<K, VB, THREADS, ITEMS, 4, kRankBallot, kDigitSplit, MINB>(st, kin, nullptr, vin, nullptr, base,
                                                                                                  (unsigned long long *)lb, n, 0, tf);
}

template <typename K, int VB>
static int partition_scatter_typed(StreamState *st, const void *kin, const void *vin, size_t n, const Transform &tf, const unsigned *base)
{
    return partition_scatter_shape<K, VB, PassConfig<K, VB>::THREADS, PassConfig<K, VB>::ITEMS, default_min_blocks(PassConfig<K, VB>::THREADS)>(
        st, kin, vin, n, tf, base);
}

template <typename K>
static int partition_scatter_by_value_size(StreamState *st, const void *kin, const void *vin, size_t vb, size_t n, const Transform &tf,
                                           const unsigned *base)
{
    if (!vin || vb == 0) return partition_scatter_typed<K, 0>(st, kin, nullptr, n, tf, base);
    switch (vb) {
    case 4: return partition_scatter_typed<K, 4>(st, kin, vin, n, tf, base);
    case 8: return partition_scatter_typed<K, 8>(st, kin, vin, n, tf, base);
    default: return BCB_EUNSUPPORTED;  // callers fall back to sort-then-cut (bcb_partition_points)
    }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// sets *flag to 1 if the range is not sorted by the transformed key (or if `force`); the caller clears it first
template <typename K>
static int run_verify(StreamState *st, const void *keys, size_t n, const Transform &tf, int *flag, int force = 0)
{
    size_t blocks = (n + 256 * 16 - 1) / (256 * 16);
    const size_t cap = (size_t)st->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if ((tf.nm | tf.xc | tf.fa) == 0)
        verify_sorted_kernel<K, true><<<(unsigned)blocks, 256, 0, st->stream>>>((const K *)keys, n, tf, flag, force);
    else
        verify_sorted_kernel<K, false><<<(unsigned)blocks, 256, 0, st->stream>>>((const K *)keys, n, tf, flag, force);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// device words of the per-stream control block used by the speculative sort
static int *spec_flag(StreamState *st) { return reinterpret_cast<int *>(st->control + kControlSpecFlag); }
static unsigned long long *spec_fallback_counter(StreamState *st) { return st->control + kControlSpecFallbacks; }

// histogram of every digit position + per-digit exclusive scan, then one pass per digit
// histogram of every digit position of the transformed keys, one read (hist: [sizeof(K)][256], zeroed by the caller)
template <typename K>
static int launch_histogram(StreamState *st, const void *src_keys, size_t n, unsigned *hist, const Transform &tf, const int *gate)
{
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0;
    if (n >= ((size_t)1 << 22)) {  // one counter column per lane (conflict-free), one CTA per SM
        constexpr int COLS = sizeof(K) == 8 ? 16 : 32;
        constexpr size_t kSmem = sizeof(K) * kRadixSize * COLS * sizeof(unsigned);
        auto kernel = ident ? radix_histogram_columns<K, COLS, true> : radix_histogram_columns<K, COLS, false>;
        static std::atomic<unsigned long long> configured[2];  // bit per device
        const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
        if (!(configured[ident].load(std::memory_order_acquire) & bit) || !bit) {
            BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
            configured[ident].fetch_or(bit, std::memory_order_release);
        }
        LaunchTimer timer(st, gate ? BCB_K_OTHER : BCB_K_RADIX_HISTOGRAM);
        kernel<<<(unsigned)st->sm_count, 1024, kSmem, st->stream>>>((const K *)src_keys, n, hist, tf, gate);
    } else {
        size_t blocks = (n * sizeof(K) + (size_t)kHistThreads * 32 - 1) / ((size_t)kHistThreads * 32);
        const size_t cap = (size_t)st->sm_count * (2048 / kHistThreads);
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        LaunchTimer timer(st, gate ? BCB_K_OTHER : BCB_K_RADIX_HISTOGRAM);
        radix_histogram<K><<<(unsigned)blocks, kHistThreads, 0, st->stream>>>((const K *)src_keys, n, hist, tf, gate);
    }
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename K, int VB>
static int sort_passes(StreamState *st, void *keys, void *values, size_t n, const Transform &tf, const void *src_keys,
                       const void *src_vals, int pass_kind, const int *gate)
{
    constexpr int NPASS = sizeof(K);
    const size_t kbytes = align_up(n * sizeof(K), 256);
    const size_t vbytes = align_up(n * (size_t)VB, 256);
    void *scratch;
    BCB_TRY(scratch_reserve(st, kbytes + vbytes, &scratch));
    void *tmp_keys = scratch;
    void *tmp_vals = VB ? (void *)((char *)scratch + kbytes) : nullptr;

    const size_t tile = tile_size_for<K, VB>(pass_kind);
    const size_t tiles = (n + tile - 1) / tile;
    void *lb;
    BCB_TRY(lookback_reserve(st, kArenaPacked, tiles * kRadixSize * sizeof(unsigned long long), &lb));

    unsigned *hist = st->hist;
    unsigned *base = st->hist + 8 * kRadixSize;
    unsigned *hot = st->hist + kHistHotOffset;  // [8 passes][2] digit values the warp-specialised pass ranks by ballot
    const bool ws_kind = sort_env().hot;  // (every pass kernel takes the hint: the r01 kernels use the constant-digit case only)
    BCB_CUDA_TRY(cudaMemsetAsync(hist, 0, NPASS * kRadixSize * sizeof(unsigned), st->stream));
    {
        BCB_TRY((launch_histogram<K>(st, src_keys, n, hist, tf, gate)));
        {
            LaunchTimer timer(st, gate ? BCB_K_OTHER : BCB_K_DIGIT_SCAN);
            digit_scan<<<NPASS, kRadixSize, 0, st->stream>>>(hist, base, gate, gate ? spec_fallback_counter(st) : nullptr, ws_kind ? hot : nullptr);
        }
        BCB_CUDA_TRY(cudaGetLastError());
    }
    void *kin = keys, *kout = tmp_keys, *vin = values, *vout = tmp_vals;
    for (int p = 0; p < NPASS; p++) {
        BCB_TRY((run_pass<K, VB>(st, p == 0 ? src_keys : kin, kout, p == 0 ? src_vals : vin, vout, base + p * kRadixSize,
                                 (unsigned long long *)lb, n, p * kRadixBits, tf, pass_kind, gate, ws_kind ? hot + 2 * p : nullptr)));
        void *t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (kin != keys) {  // odd pass count (8-bit keys): result is in the temporary (never gated: those sorts do not speculate)
        BCB_CUDA_TRY(cudaMemcpyAsync(keys, kin, n * sizeof(K), cudaMemcpyDeviceToDevice, st->stream));
        if (VB) BCB_CUDA_TRY(cudaMemcpyAsync(values, vin, n * (size_t)VB, cudaMemcpyDeviceToDevice, st->stream));
    }
    return BCB_SUCCESS;
}

// second half of a speculative keys-only sort: verify the speculation on the device; the deterministic sort of the
// (permuted) buffer follows, every launch gated on the flag
template <typename K>
static int verify_and_fix(StreamState *st, void *keys, size_t n, const Transform &tf)
{
    int *flag = spec_flag(st);
    BCB_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st->stream));
    {
        LaunchTimer timer(st, BCB_K_OTHER);
        BCB_TRY(run_verify<K>(st, keys, n, tf, flag, sort_env().force_fallback ? 1 : 0));
    }
    st->spec_runs++;
    return sort_passes<K, 0>(st, keys, nullptr, n, tf, keys, nullptr, kPassDeterministic, flag);
}

// ---- small ranges: the whole sort in ONE launch ---------------------------------------------------------------
// A small sort is all latency in the multi-launch path: six launches and a memset, ~60 us however few keys there are.
// Here every tile has its own CTA, all CTAs are resident (cooperative launch), and
// a pass is: rank the tile in shared memory (the deterministic atomic-OR ranking, keys kept in registers) -> publish the
// tile's 256 digit counts -> grid barrier -> every CTA sums the counts of ALL tiles itself (a few hundred coalesced
// L2 reads per digit thread: no serial chain) -> scatter straight to the other buffer (the whole range lives in L2)
// -> grid barrier.  No histogram pass, no descriptors, no look-back.  Keys stay raw in memory; stable by construction.
constexpr int kSmThreads = 512, kSmWarps = kSmThreads / 32;
// largest launch, measured on B200 (us per sort, one launch / six launches): u32 keys 2^10 23 / 61, 2^14 44 / 60, 2^16 48 / 58,
// 2^18 (43 tiles) 54 / 61; with a u32 payload 2^10 36 / 63, 2^14 66 / 70, 2^16 68 / 69, 2^18 79 / 71; 2^20 keys 124 / 68 --
// the two grid barriers and the all-tiles sum of every pass grow with the tile count, the launches saved do not
constexpr size_t kSmallMaxTilesKeys = 48, kSmallMaxTilesPairs = 8;
template <typename K> struct SmallShape { static constexpr int ITEMS = sizeof(K) == 8 ? 8 : 12; static constexpr int TILE = kSmThreads * ITEMS; };

__device__ __forceinline__ void grid_barrier(unsigned long long *counter, unsigned long long target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // this CTA's global writes before the arrival
        atomicAdd(counter, 1ull);
        while (ld_acquire_u64(counter) < target) { }
    }
    __syncthreads();
}

template <typename K, int VB>
__global__ void __launch_bounds__(kSmThreads, 2)
small_sort_kernel(K *buf_a, K *buf_b, void *val_a_v, void *val_b_v, const K *src, const void *src_vals_v, unsigned *tile_counts, unsigned n,
                  int npass, const __grid_constant__ Transform tf, unsigned long long *bar, unsigned long long bar_base)
{
    typedef typename value_type<VB>::type V;
    constexpr int ITEMS = SmallShape<K>::ITEMS, TILE = SmallShape<K>::TILE;
    __shared__ unsigned mask[kSmWarps][kRadixSize];  // peers of the current round, per (warp, digit)
    __shared__ unsigned cnt[kSmWarps][kRadixSize];   // keys of (warp, digit) so far; after the ranking: start of the warp's run in the tile's digit run
    __shared__ unsigned base[kRadixSize];            // where this tile's run of each digit value starts in the output
    __shared__ unsigned wsum[kRadixSize / 32];
    __shared__ unsigned cnt_total[kRadixSize];       // single-tile launches: the tile's digit counts never leave the CTA
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, tile = blockIdx.x, G = gridDim.x;
    const unsigned seg = tile * TILE + w * (32 * ITEMS) + lane;  // this thread's item i is element seg + 32 * i: memory order = (i, lane) order
    unsigned long long target = bar_base;
    for (int p = 0; p < npass; p++) {
        const K *in = p == 0 ? src : ((p & 1) ? buf_b : buf_a);
        K *out = (p & 1) ? buf_a : buf_b;
        const V *vin = reinterpret_cast<const V *>(p == 0 ? src_vals_v : ((p & 1) ? val_b_v : val_a_v));
        V *vout = reinterpret_cast<V *>((p & 1) ? val_a_v : val_b_v);
        const int shift = p * kRadixBits;
        for (unsigned i = tid; i < kSmWarps * kRadixSize; i += kSmThreads) { (&mask[0][0])[i] = 0; (&cnt[0][0])[i] = 0; }
        __syncthreads();
        K key[ITEMS];
        unsigned short rank[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; i++) key[i] = (seg + 32 * i < n) ? __ldcg(in + seg + 32 * i) : (K)0;  // (L2 only: other SMs wrote this buffer a pass ago)
        // rank inside the warp: the lanes OR their bit into the mask of their digit; after a warp barrier the mask holds the
        // complete peer set whatever order the atomics were applied in; the highest peer clears it and bumps the count
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const bool valid = seg + 32 * i < n;
            const unsigned d = digit_of<K>(key[i], shift, tf);
            if (valid) atomicOr(&mask[w][d], 1u << lane);
            __syncwarp();
            const unsigned m = valid ? mask[w][d] : 0u, c = valid ? cnt[w][d] : 0u;
            __syncwarp();
            if (valid && (m >> lane) == 1u) {
                mask[w][d] = 0;
                cnt[w][d] = c + __popc(m);
            }
            __syncwarp();
            rank[i] = (unsigned short)(c + __popc(m & lanemask_lt()));
        }
        __syncthreads();
        if (tid < kRadixSize) {  // per digit: exclusive prefix over the warps, the tile's count to global memory
            unsigned run = 0;
#pragma unroll
            for (int v = 0; v < kSmWarps; v++) {
                const unsigned c = cnt[v][tid];
                cnt[v][tid] = run;
                run += c;
            }
            if (G > 1) tile_counts[tile * kRadixSize + tid] = run;
            else cnt_total[tid] = run;
        }
        if (G > 1) {
            target += G;
            grid_barrier(bar, target);
        } else {
            __syncthreads();
        }
        {
            // counts of this digit value in ALL tiles (total) and in the tiles before this one (below): both halves of the
            // CTA take every other tile, eight loads in flight per thread -- a handful of L2 round trips, no serial chain
            const unsigned d = tid & (kRadixSize - 1), half = tid >> 8;
            unsigned total = 0, below = 0;
            if (G > 1) {
                unsigned t = half;
                for (; t + 14 < G; t += 16) {
                    unsigned c[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) c[u] = __ldcg(tile_counts + (t + 2 * u) * kRadixSize + d);
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        total += c[u];
                        below += (t + 2 * u < tile) ? c[u] : 0u;
                    }
                }
                for (; t < G; t += 2) {
                    const unsigned c = __ldcg(tile_counts + t * kRadixSize + d);
                    total += c;
                    below += t < tile ? c : 0u;
                }
            } else if (half == 0) {
                total = cnt_total[d];
            }
            if (half == 1) { mask[0][d] = total; mask[1][d] = below; }  // (the mask table is idle between the ranking and the next pass)
            __syncthreads();
            if (half == 0) {
                total += mask[0][d];
                below += mask[1][d];
                unsigned incl = total;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
                    if ((int)lane >= off) incl += o;
                }
                if (lane == 31) wsum[w] = incl;
                base[d] = incl - total + below;  // (+ the digit values of the earlier warps, below)
            }
        }
        __syncthreads();
        if (tid < kRadixSize) {
            unsigned add = 0;
            for (unsigned v = 0; v < w; v++) add += wsum[v];
            base[tid] += add;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; i++) {
            const unsigned idx = seg + 32 * i;
            if (idx < n) {
                const unsigned d = digit_of<K>(key[i], shift, tf);
                const unsigned pos = base[d] + cnt[w][d] + rank[i];
                out[pos] = key[i];
                if constexpr (VB > 0) vout[pos] = __ldcg(vin + idx);
            }
        }
        if (G > 1) {
            target += G;
            grid_barrier(bar, target);  // the pass's output is complete (and the tile counts have been read) before anyone goes on
        } else {
            __threadfence_block();
            __syncthreads();
        }
    }
}

// resident CTAs of the small-sort kernel on the device (0 = not asked yet)
template <typename K, int VB>
static int small_sort_capacity(StreamState *st)
{
    static std::atomic<int> cached[64];
    int cap = st->device < 64 ? cached[st->device].load(std::memory_order_acquire) : 0;
    if (cap == 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, small_sort_kernel<K, VB>, kSmThreads, 0) != cudaSuccess) { (void)cudaGetLastError(); per_sm = 0; }
        cap = per_sm > 0 ? per_sm * st->sm_count : -1;
        if (st->device < 64) cached[st->device].store(cap, std::memory_order_release);
    }
    return cap;
}

// BCB_EUNSUPPORTED: not a small range (or switched off): the caller goes on with the multi-launch sort
template <typename K, int VB>
static int small_sort(StreamState *st, void *keys, void *values, size_t n, const Transform &tf, const void *src_keys, const void *src_vals)
{
    {
        const char *e = std::getenv("BCB_SORT_SMALL");  // 0: always the multi-launch sort (A/B comparison, test hook; read per call)
        if (e && e[0] == '0') return BCB_EUNSUPPORTED;
    }
    constexpr size_t TILE = SmallShape<K>::TILE;
    const size_t tiles = (n + TILE - 1) / TILE;
    const int cap = small_sort_capacity<K, VB>(st);
    if (cap <= 0 || tiles > (size_t)cap || tiles > (VB ? kSmallMaxTilesPairs : kSmallMaxTilesKeys)) return BCB_EUNSUPPORTED;
    constexpr int NPASS = sizeof(K);
    const size_t kbytes = align_up(n * sizeof(K), 256), vbytes = align_up(n * (size_t)VB, 256);
    void *scratch;
    BCB_TRY(scratch_reserve(st, kbytes + vbytes + tiles * kRadixSize * sizeof(unsigned), &scratch));
    K *tmp_keys = (K *)scratch;
    void *tmp_vals = VB ? (void *)((char *)scratch + kbytes) : nullptr;
    unsigned *tile_counts = (unsigned *)((char *)scratch + kbytes + vbytes);
    K *buf_a = (K *)keys;
    const K *src = (const K *)src_keys;
    unsigned n32 = (unsigned)n;
    int npass = NPASS;
    unsigned long long *bar = st->control + kControlGridBar;
    unsigned long long bar_base = st->gridbar_base;
    {
        LaunchTimer timer(st, BCB_K_ONESWEEP_PASS);
        void *args[] = {(void *)&buf_a, (void *)&tmp_keys, (void *)&values, (void *)&tmp_vals, (void *)&src, (void *)&src_vals, (void *)&tile_counts,
                        (void *)&n32, (void *)&npass, (void *)&tf, (void *)&bar, (void *)&bar_base};
        BCB_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)small_sort_kernel<K, VB>, dim3((unsigned)tiles), dim3(kSmThreads), args, 0, st->stream));
    }
    if (tiles > 1) st->gridbar_base += 2ull * NPASS * tiles;  // (only once the launch is in: the counter and the base must stay in step)
    if (NPASS & 1) {  // 8-bit keys: one pass, the result is in the temporary
        BCB_CUDA_TRY(cudaMemcpyAsync(keys, tmp_keys, n * sizeof(K), cudaMemcpyDeviceToDevice, st->stream));
        if (VB) BCB_CUDA_TRY(cudaMemcpyAsync(values, tmp_vals, n * (size_t)VB, cudaMemcpyDeviceToDevice, st->stream));
    }
    return BCB_SUCCESS;
}

template <typename K, int VB>
static int sort_typed(StreamState *st, void *keys, void *values, size_t n, const Transform &tf, const void *src_keys = nullptr,
                      const void *src_vals = nullptr)
{
    // src_keys / src_vals: read the input from there instead (sorted copy; the source is left untouched)
    if (!src_keys) { src_keys = keys; src_vals = values; }
    {
        const int rc = small_sort<K, VB>(st, keys, values, n, tf, src_keys, src_vals);
        if (rc != BCB_EUNSUPPORTED) return rc;
    }
    int pass_kind = kPassDeterministic;
    {   // large sorts the warp-specialised kernel covers with its deterministic ranking: payloads, non-injective transforms
        const SortEnv &env = sort_env();
        const bool aligned = ((((uintptr_t)keys) | ((uintptr_t)src_keys) | ((uintptr_t)values) | ((uintptr_t)src_vals)) & 15) == 0;
        if (env.ws && aligned && n >= env.ws_min && ws_supports((int)sizeof(K), VB, true)) pass_kind = kPassWsDet;
    }
    if constexpr (VB == 0 && (sizeof(K) == 4 || sizeof(K) == 8)) {
        // The verification argument needs an INJECTIVE key transform (sorted permutation => unique bytes).  The
        // reference's descending float transform is not (radix_sort.hpp:100-127: -0.0 / +denorm_min and
        // +0.0 / -denorm_min collide, and their relative order is then decided by stability alone), so
        // descending float / double sorts always take the deterministic kernel.
        const bool injective = !(tf.fa != 0 && tf.nm != 0);
        const SortEnv &env = sort_env();
        if (injective && env.speculative && n >= env.spec_min) {
            pass_kind = kPassTwoSweep;
            // bulk copies need 16-byte aligned arrays (the scratch buffer always is)
            const bool aligned = ((((uintptr_t)keys) | ((uintptr_t)src_keys)) & 15) == 0;
            // (64-bit keys: measured 23.7 Gkeys/s with onesweep_ws against 29.2 with the two-sweep kernel -- 21504-key
            // tiles give each CTA too few keys per digit run to pay for the helper-warp pipeline)
            if (env.ws && aligned && n >= env.ws_min && sizeof(K) == 4) pass_kind = kPassWs;
        }
    }
    if (pass_kind == kPassDeterministic || pass_kind == kPassWsDet) return sort_passes<K, VB>(st, keys, values, n, tf, src_keys, src_vals, pass_kind, nullptr);

    if constexpr (VB == 0 && (sizeof(K) == 4 || sizeof(K) == 8)) {
        BCB_TRY((sort_passes<K, VB>(st, keys, values, n, tf, src_keys, src_vals, pass_kind, nullptr)));
        return verify_and_fix<K>(st, keys, n, tf);
    }
    return BCB_EINVAL;
}

template <typename K>
static int sort_by_value_size(StreamState *st, void *keys, void *values, size_t vb, size_t n, const Transform &tf,
                              const void *src_keys = nullptr, const void *src_vals = nullptr)
{
    if (!values || vb == 0) return sort_typed<K, 0>(st, keys, nullptr, n, tf, src_keys, nullptr);
    const bool aligned = ((uintptr_t)values % (vb <= 16 ? vb : 1)) == 0 && (!src_vals || ((uintptr_t)src_vals % (vb <= 16 ? vb : 1)) == 0);
    if (aligned) {
        switch (vb) {
        case 1: return sort_typed<K, 1>(st, keys, values, n, tf, src_keys, src_vals);
        case 2: return sort_typed<K, 2>(st, keys, values, n, tf, src_keys, src_vals);
        case 4: return sort_typed<K, 4>(st, keys, values, n, tf, src_keys, src_vals);
        case 8: return sort_typed<K, 8>(st, keys, values, n, tf, src_keys, src_vals);
        case 16: return sort_typed<K, 16>(st, keys, values, n, tf, src_keys, src_vals);
        default: break;
        }
    }
    if (src_keys) {  // generic payload: copy first, then sort in place
        BCB_CUDA_TRY(cudaMemcpyAsync(keys, src_keys, n * sizeof(K), cudaMemcpyDeviceToDevice, st->stream));
        BCB_CUDA_TRY(cudaMemcpyAsync(values, src_vals, n * vb, cudaMemcpyDeviceToDevice, st->stream));
    }
    // generic payload: stable-sort (key, original index) pairs, then gather the payload bytes
    unsigned *idx;
    unsigned char *gathered;
    BCB_CUDA_TRY(cudaMallocAsync((void **)&idx, n * sizeof(unsigned), st->stream));
    cudaError_t e = cudaMallocAsync((void **)&gathered, n * vb, st->stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); (void)cudaFreeAsync(idx, st->stream); return (int)e; }
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)st->sm_count * 16;
    if (blocks > cap) blocks = cap;
    iota_u32_kernel<<<(unsigned)blocks, 256, 0, st->stream>>>(idx, n);
    int rc = sort_typed<K, 4>(st, keys, idx, n, tf);
    if (rc == BCB_SUCCESS) {
        gather_bytes_kernel<<<(unsigned)blocks, 256, 0, st->stream>>>((const unsigned char *)values, idx, gathered, n, vb);
        if (cudaMemcpyAsync(values, gathered, n * vb, cudaMemcpyDeviceToDevice, st->stream) != cudaSuccess) rc = (int)cudaGetLastError();
    }
    (void)cudaFreeAsync(idx, st->stream);
    (void)cudaFreeAsync(gathered, st->stream);
    if (rc == BCB_SUCCESS) {
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) rc = (int)le;
    }
    return rc;
}

static int radix_sort_impl(StreamState *st, int key_dtype, int ascending, void *keys, size_t n, void *values, size_t vb,
                           const void *src_keys = nullptr, const void *src_vals = nullptr)
{
    const Transform tf = make_transform(key_dtype, ascending != 0);
    switch (dtype_size(key_dtype)) {
    case 1: return sort_by_value_size<unsigned char>(st, keys, values, vb, n, tf, src_keys, src_vals);
    case 2: return sort_by_value_size<unsigned short>(st, keys, values, vb, n, tf, src_keys, src_vals);
    case 4: return sort_by_value_size<unsigned>(st, keys, values, vb, n, tf, src_keys, src_vals);
    case 8: return sort_by_value_size<unsigned long long>(st, keys, values, vb, n, tf, src_keys, src_vals);
    default: return BCB_EINVAL;
    }
}

static int insertion_sort_impl(StreamState *st, int key_dtype, int greater, void *keys, size_t n, void *values, size_t vb)
{
    unsigned char *v = (values && vb) ? (unsigned char *)values : nullptr;
    switch (key_dtype) {
#define X(DT, T) case DT: insertion_sort_kernel<T><<<1, 32, 0, st->stream>>>((T *)keys, n, greater, v, vb); break;
        BCB_FOR_EACH_TYPE(X)
#undef X
    default: return BCB_EINVAL;
    }
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// ---- what the other radix translation units (radix_exchange.cu, radix_field.cu) use from this one ----
int radix_sort_device(StreamState *st, int key_dtype, int ascending, void *keys, size_t n, void *values, size_t value_bytes)
{
    return radix_sort_impl(st, key_dtype, ascending, keys, n, values, value_bytes);
}

int radix_verify_and_fix(StreamState *st, int key_bytes, void *keys, size_t n, const Transform &tf)
{
    return key_bytes == 4 ? verify_and_fix<unsigned>(st, keys, n, tf) : verify_and_fix<unsigned long long>(st, keys, n, tf);
}

bool radix_speculation_enabled() { return sort_env().speculative; }

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_radix_sort(bcb_stream stream, int key_dtype, int ascending, void *keys, size_t n, void *values, size_t value_bytes)
{
    if (!dtype_size(key_dtype)) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;
    if (!keys) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    return radix_sort_impl(st, key_dtype, ascending, keys, n, values, values ? value_bytes : 0);
}

int bcb_radix_sort_copy(bcb_stream stream, int key_dtype, int ascending, const void *keys_in, void *keys_out, size_t n,
                        const void *values_in, void *values_out, size_t value_bytes)
{
    const size_t w = dtype_size(key_dtype);
    if (!w) return BCB_EINVAL;
    if (n == 0) return BCB_SUCCESS;
    if (!keys_in || !keys_out || ((values_in != nullptr) != (values_out != nullptr))) return BCB_EINVAL;
    if ((keys_in == keys_out) != (values_in == values_out) && values_in) return BCB_EINVAL;  // in place means both
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    const size_t vb = values_in ? value_bytes : 0;
    if (n < 2 || keys_in == keys_out) {
        if (keys_in != keys_out) {
            BCB_CUDA_TRY(cudaMemcpyAsync(keys_out, keys_in, n * w, cudaMemcpyDeviceToDevice, st->stream));
            if (vb) BCB_CUDA_TRY(cudaMemcpyAsync(values_out, values_in, n * vb, cudaMemcpyDeviceToDevice, st->stream));
        }
        if (n < 2) return BCB_SUCCESS;
        return radix_sort_impl(st, key_dtype, ascending, keys_out, n, values_out, vb);
    }
    return radix_sort_impl(st, key_dtype, ascending, keys_out, n, values_out, vb, keys_in, values_in);
}

int bcb_insertion_sort(bcb_stream stream, int key_dtype, int greater, void *keys, size_t n, void *values, size_t value_bytes)
{
    if (!dtype_size(key_dtype)) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;  // insertion_sort.hpp:34-37
    if (!keys) return BCB_EINVAL;
    if (n > 4096 || (values && value_bytes > 256)) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    return insertion_sort_impl(st, key_dtype, greater, keys, n, values, value_bytes);
}

int bcb_partition_points(bcb_stream stream, int key_dtype, int ascending, const void *sorted_keys, size_t n,
                         const unsigned long long *splitters_host, size_t num_splitters, unsigned long long *points_host)
{
    if (!dtype_size(key_dtype)) return BCB_EINVAL;
    if (num_splitters == 0) return BCB_SUCCESS;
    if (!splitters_host || !points_host || (n && !sorted_keys)) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    const Transform tf = make_transform(key_dtype, ascending != 0);
    unsigned long long *dev;
    BCB_CUDA_TRY(cudaMallocAsync((void **)&dev, 2 * num_splitters * sizeof(unsigned long long), st->stream));
    int rc = BCB_SUCCESS;
    cudaError_t e = cudaMemcpyAsync(dev, splitters_host, num_splitters * sizeof(unsigned long long), cudaMemcpyHostToDevice, st->stream);
    if (e == cudaSuccess) {
        const unsigned blocks = (unsigned)((num_splitters + 63) / 64);
        switch (dtype_size(key_dtype)) {
        case 1: partition_points_kernel<unsigned char><<<blocks, 64, 0, st->stream>>>((const unsigned char *)sorted_keys, n, dev, (unsigned)num_splitters, dev + num_splitters, tf); break;
        case 2: partition_points_kernel<unsigned short><<<blocks, 64, 0, st->stream>>>((const unsigned short *)sorted_keys, n, dev, (unsigned)num_splitters, dev + num_splitters, tf); break;
        case 4: partition_points_kernel<unsigned><<<blocks, 64, 0, st->stream>>>((const unsigned *)sorted_keys, n, dev, (unsigned)num_splitters, dev + num_splitters, tf); break;
        default: partition_points_kernel<unsigned long long><<<blocks, 64, 0, st->stream>>>((const unsigned long long *)sorted_keys, n, dev, (unsigned)num_splitters, dev + num_splitters, tf); break;
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(points_host, dev + num_splitters, num_splitters * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream);
    if (e != cudaSuccess) rc = (int)e;
    (void)cudaFreeAsync(dev, st->stream);
    e = cudaStreamSynchronize(st->stream);
    if (rc == BCB_SUCCESS && e != cudaSuccess) rc = (int)e;
    if (rc != BCB_SUCCESS) (void)cudaGetLastError();
    return rc;
}

int bcb_is_sorted_by_radix_key(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n, int *result_host)
{
    if (!result_host || !dtype_size(key_dtype)) return BCB_EINVAL;
    *result_host = 1;
    if (n < 2) return BCB_SUCCESS;
    if (!keys) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    const Transform tf = make_transform(key_dtype, ascending != 0);
    int *flag = (int *)st->pinned_slot_dev;
    *(volatile int *)st->pinned_slot = 0;
    switch (dtype_size(key_dtype)) {
    case 1: BCB_TRY(run_verify<unsigned char>(st, keys, n, tf, flag)); break;
    case 2: BCB_TRY(run_verify<unsigned short>(st, keys, n, tf, flag)); break;
    case 4: BCB_TRY(run_verify<unsigned>(st, keys, n, tf, flag)); break;
    default: BCB_TRY(run_verify<unsigned long long>(st, keys, n, tf, flag)); break;
    }
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    *result_host = (*(volatile int *)st->pinned_slot) ? 0 : 1;
    return BCB_SUCCESS;
}

int bcb_sort_speculation_stats(bcb_stream stream, unsigned long long *runs, unsigned long long *fallbacks)
{
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    if (runs) *runs = st->spec_runs;
    if (fallbacks) {  // counted on the device by the gated fallback launches
        BCB_CUDA_TRY(cudaMemcpyAsync(fallbacks, spec_fallback_counter(st), sizeof(unsigned long long), cudaMemcpyDeviceToHost, st->stream));
        BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    }
    return BCB_SUCCESS;
}

static int split_transform(int key_dtype, int ascending, const unsigned long long *splitters_host, size_t num_splitters, Transform *tf)
{
    if (!dtype_size(key_dtype)) return BCB_EINVAL;
    if (num_splitters > (size_t)kMaxSplitters) return BCB_EUNSUPPORTED;
    if (num_splitters && !splitters_host) return BCB_EINVAL;
    *tf = make_transform(key_dtype, ascending != 0);
    tf->nsplit = (int)num_splitters;
    for (size_t j = 0; j < (size_t)kMaxSplitters; j++) tf->split[j] = j < num_splitters ? splitters_host[j] : ~0ull;
    return BCB_SUCCESS;
}

int bcb_radix_top_histogram(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n, unsigned long long *counts_host)
{
    if (!counts_host) return BCB_EINVAL;
    const size_t w = dtype_size(key_dtype);
    if (!w) return BCB_EINVAL;
    for (int d = 0; d < kRadixSize; d++) counts_host[d] = 0;
    if (n == 0) return BCB_SUCCESS;
    if (!keys) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    const Transform tf = make_transform(key_dtype, ascending != 0);
    unsigned *hist = st->hist;
    BCB_CUDA_TRY(cudaMemsetAsync(hist, 0, w * kRadixSize * sizeof(unsigned), st->stream));
    int rc;
    switch (w) {
    case 1: rc = launch_histogram<unsigned char>(st, keys, n, hist, tf, nullptr); break;
    case 2: rc = launch_histogram<unsigned short>(st, keys, n, hist, tf, nullptr); break;
    case 4: rc = launch_histogram<unsigned>(st, keys, n, hist, tf, nullptr); break;
    default: rc = launch_histogram<unsigned long long>(st, keys, n, hist, tf, nullptr); break;
    }
    BCB_TRY(rc);
    unsigned host[kRadixSize];
    BCB_CUDA_TRY(cudaMemcpyAsync(host, hist + (w - 1) * kRadixSize, sizeof(host), cudaMemcpyDeviceToHost, st->stream));
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    for (int d = 0; d < kRadixSize; d++) counts_host[d] = host[d];
    return BCB_SUCCESS;
}

int bcb_partition_counts(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n,
                         const unsigned long long *splitters_host, size_t num_splitters, unsigned long long *counts_host)
{
    if (!counts_host) return BCB_EINVAL;
    Transform tf;
    BCB_TRY(split_transform(key_dtype, ascending, splitters_host, num_splitters, &tf));
    for (size_t j = 0; j <= num_splitters; j++) counts_host[j] = 0;
    if (n == 0) return BCB_SUCCESS;
    if (!keys) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    switch (dtype_size(key_dtype)) {
    case 1: return partition_counts_typed<unsigned char>(st, keys, n, tf, counts_host);
    case 2: return partition_counts_typed<unsigned short>(st, keys, n, tf, counts_host);
    case 4: return partition_counts_typed<unsigned>(st, keys, n, tf, counts_host);
    default: return partition_counts_typed<unsigned long long>(st, keys, n, tf, counts_host);
    }
}

static int partition_scatter_dispatch(StreamState *st, int key_dtype, const void *keys_in, const void *values_in, size_t vb, size_t n,
                                      const Transform &tf, const unsigned *base)
{
    switch (dtype_size(key_dtype)) {
    case 1: return partition_scatter_by_value_size<unsigned char>(st, keys_in, values_in, vb, n, tf, base);
    case 2: return partition_scatter_by_value_size<unsigned short>(st, keys_in, values_in, vb, n, tf, base);
    case 4: return partition_scatter_by_value_size<unsigned>(st, keys_in, values_in, vb, n, tf, base);
    default: return partition_scatter_by_value_size<unsigned long long>(st, keys_in, values_in, vb, n, tf, base);
    }
}

int bcb_partition_scatter(bcb_stream stream, int key_dtype, int ascending, const void *keys_in, const void *values_in,
                          size_t value_bytes, size_t n, const unsigned long long *splitters_host, size_t num_splitters,
                          void *const *dst_keys, void *const *dst_values)
{
    Transform tf;
    BCB_TRY(split_transform(key_dtype, ascending, splitters_host, num_splitters, &tf));
    if (n == 0) return BCB_SUCCESS;
    if (!keys_in || !dst_keys) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    const size_t vb = values_in ? value_bytes : 0;
    if (vb && !dst_values) return BCB_EINVAL;
    if (vb != 0 && vb != 4 && vb != 8) return BCB_EUNSUPPORTED;
    for (size_t j = 0; j <= num_splitters; j++) {
        tf.dst_keys[j] = (unsigned long long)(uintptr_t)dst_keys[j];
        tf.dst_vals[j] = vb ? (unsigned long long)(uintptr_t)dst_values[j] : 0ull;
        // destinations must be element aligned (the kernel derives the lane alignment of a run from its address)
        if (!dst_keys[j] || tf.dst_keys[j] % dtype_size(key_dtype) || (vb && (!dst_values[j] || tf.dst_vals[j] % vb))) return BCB_EINVAL;
    }
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    // BCB_SPLIT_WS=1 (opt-in, read per call): the warp-specialised exchange kernel of radix_exchange_ws.cu (lane-private
    // ranking, one bulk copy per bucket run) for shards of >= 2^BCB_SPLIT_WS_MIN_LOG2 (default 24) keys.  Measured on
    // one B200 with local destinations, 2^30 u32 keys: 3.5 / 3.6 / 4.0 ms for 1 / 3 / 7 splitters against 3.5 / 4.0 / 5.0 ms
    // for the LSU kernel below -- ahead for 4 and 8 ranks, but its bulk copies into PEER memory are not yet measured on
    // more than one GPU, so the LSU kernel stays the default.
    {
        const char *e = std::getenv("BCB_SPLIT_WS");
        if (e && e[0] == '1') {
            const char *v = std::getenv("BCB_SPLIT_WS_MIN_LOG2");
            const int k = v ? std::atoi(v) : 24;
            if (n >= ((size_t)1 << (k < 10 ? 10 : (k > 40 ? 40 : k)))) {
                const int rc = ws_exchange_pass(st, (int)dtype_size(key_dtype), keys_in, values_in, (int)vb, n, tf);
                if (rc != BCB_EUNSUPPORTED) return rc;
            }
        }
    }
    unsigned *base = st->hist + 8 * kRadixSize;  // every bucket starts at its own destination pointer
    BCB_CUDA_TRY(cudaMemsetAsync(base, 0, kRadixSize * sizeof(unsigned), st->stream));
    return partition_scatter_dispatch(st, key_dtype, keys_in, values_in, vb, n, tf, base);
}

int bcb_partition_by_splitters(bcb_stream stream, int key_dtype, int ascending, const void *keys_in, void *keys_out,
                               const void *values_in, void *values_out, size_t value_bytes, size_t n,
                               const unsigned long long *splitters_host, size_t num_splitters, unsigned long long *counts_host)
{
    if (!counts_host) return BCB_EINVAL;
    Transform tf;
    BCB_TRY(split_transform(key_dtype, ascending, splitters_host, num_splitters, &tf));
    for (size_t j = 0; j <= num_splitters; j++) counts_host[j] = 0;
    if (n == 0) return BCB_SUCCESS;
    if (!keys_in || !keys_out) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    const size_t vb = (values_in && values_out) ? value_bytes : 0;
    if (vb != 0 && vb != 4 && vb != 8) return BCB_EUNSUPPORTED;
    BCB_TRY(bcb_partition_counts(stream, key_dtype, ascending, keys_in, n, splitters_host, num_splitters, counts_host));
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    // one output buffer: bucket b starts at the exclusive prefix of the counts (digit_scan of the histogram)
    unsigned *base = st->hist + 8 * kRadixSize;
    digit_scan<<<1, kRadixSize, 0, st->stream>>>(st->hist, base);
    BCB_CUDA_TRY(cudaGetLastError());
    for (size_t j = 0; j <= num_splitters; j++) {
        tf.dst_keys[j] = (unsigned long long)(uintptr_t)keys_out;
        tf.dst_vals[j] = (unsigned long long)(uintptr_t)values_out;
    }
    BCB_TRY(partition_scatter_dispatch(st, key_dtype, keys_in, values_in, vb, n, tf, base));
    BCB_CUDA_TRY(cudaStreamSynchronize(st->stream));
    return BCB_SUCCESS;
}

int bcb_sort_host(bcb_stream stream, int key_dtype, int descending, void *host_keys, size_t n)
{
    const size_t w = dtype_size(key_dtype);
    if (!w) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;  // sort.hpp:45-48
    if (!host_keys) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    void *dev;
    BCB_CUDA_TRY(cudaMallocAsync(&dev, n * w, st->stream));
    int rc = BCB_SUCCESS;
    cudaError_t e = cudaSuccess;
    // a large range in pageable memory (sort(v.begin(), v.end()) on a std::vector) is staged through pinned slots by
    // several host threads (runtime.cu); anything else is one DMA on the stream
    bool staged = false;
    if (n * w >= ((size_t)32 << 20)) {
        e = cudaStreamSynchronize(st->stream);  // the allocation is ready for the staging streams
        if (e != cudaSuccess) rc = (int)e;
        if (rc == BCB_SUCCESS) {
            const int s = staged_copy_pageable(dev, host_keys, n * w, true);
            if (s == BCB_SUCCESS) staged = true;
            else if (s != BCB_EUNSUPPORTED) rc = s;
        }
    }
    if (rc == BCB_SUCCESS && !staged) {
        e = cudaMemcpyAsync(dev, host_keys, n * w, cudaMemcpyHostToDevice, st->stream);
        if (e != cudaSuccess) rc = (int)e;
    }
    if (rc == BCB_SUCCESS) {
        // dispatch_gpu_sort, sort.hpp:34-81
        rc = (n <= 32) ? insertion_sort_impl(st, key_dtype, descending, dev, n, nullptr, 0)
                       : radix_sort_impl(st, key_dtype, !descending, dev, n, nullptr, 0);
    }
    if (rc == BCB_SUCCESS && staged) {
        e = cudaStreamSynchronize(st->stream);
        if (e != cudaSuccess) rc = (int)e;
        if (rc == BCB_SUCCESS) rc = staged_copy_pageable(dev, host_keys, n * w, false);
    } else if (rc == BCB_SUCCESS) {
        e = cudaMemcpyAsync(host_keys, dev, n * w, cudaMemcpyDeviceToHost, st->stream);
        if (e != cudaSuccess) rc = (int)e;
    }
    (void)cudaFreeAsync(dev, st->stream);
    e = cudaStreamSynchronize(st->stream);
    if (rc == BCB_SUCCESS && e != cudaSuccess) rc = (int)e;
    if (rc != BCB_SUCCESS) (void)cudaGetLastError();
    return rc;
}

}  // extern "C"
