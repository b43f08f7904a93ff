// radix_sort_ns.cu -- onesweep pass whose tile-local 8-bit sort is two stable 4-bit splits ("nibble split").
//
// Why: on B200 the pass kernel is bound by the shared-memory / LSU data pipe (ncu: 77 % busy, ~29 wavefronts per 32
// keys with the atomic-OR ranking of radix_sort.cu, most of them bank-conflict replays of random accesses to the
// per-warp digit tables).  A 4-bit digit has only 16 values, so a thread can count and rank its own keys entirely
// in REGISTERS (sixteen 4-bit counters in one 64-bit word), the per-thread counts are combined with packed shuffle
// scans (16 byte-wide fields in four registers, indexed with PRMT), and the only shared-memory traffic left per
// split is one 16-bit table lookup and the scatter itself.  Two such splits (low nibble, then high nibble of the
// digit) sort the tile stably by the 8-bit digit with about half the wavefronts and no atomics at all -- hence
// deterministic by construction.  Everything around it (global histogram, decoupled look-back with 64-bit epoch-tagged
// descriptors, run-wise coalesced scatter, persistent round-robin tiles with register prefetch) is the same as in
// radix_sort.cu.
//
// Layout: 512 threads x 12 keys, blocked (thread t owns tile positions 12t .. 12t+11, so its own keys are already in
// stable order); 16-lane virtual warps keep every packed byte counter <= 192.
#include "radix_common.cuh"

namespace bcb {

constexpr int kNsThreads = 512;
constexpr int kNsItems = 12;
constexpr int kNsTile = kNsThreads * kNsItems;  // 6144
constexpr int kNsVWarps = kNsThreads / 16;      // 32

size_t ns_tile_size() { return kNsTile; }

template <typename K> struct NsSmem {
    static constexpr size_t kKeys = (size_t)kNsTile * sizeof(K);
    static constexpr size_t kTab = (size_t)kNsVWarps * 16 * sizeof(unsigned short);  // [vwarp][nibble]
    static constexpr size_t kSmall = 3 * kRadixSize * sizeof(unsigned) + 256;        // dstart, dend, out_base, misc
    static constexpr size_t kBytes = kKeys + kTab + kSmall;
};

// blocked tile load: thread t reads keys [12t, 12t+12) of the tile; 128-bit loads when the addresses allow it
template <typename K>
__device__ __forceinline__ void ns_load_tile(const K *__restrict__ keys_in, size_t n, size_t tile, K (&key)[kNsItems])
{
    const size_t tile_base = tile * (size_t)kNsTile;
    const size_t first = tile_base + (size_t)threadIdx.x * kNsItems;
    constexpr int VEC = 16 / sizeof(K);
    if (tile_base + kNsTile <= n) {
        if constexpr (sizeof(K) >= 4) {
            if ((((uintptr_t)keys_in) & 15) == 0) {
#pragma unroll
                for (int v = 0; v < kNsItems / VEC; v++) {
                    const uint4 x = ld_stream_v4(keys_in + first + v * VEC);
                    const K *e = reinterpret_cast<const K *>(&x);
#pragma unroll
                    for (int k = 0; k < VEC; k++) key[v * VEC + k] = e[k];
                }
                return;
            }
        }
#pragma unroll
        for (int i = 0; i < kNsItems; i++) key[i] = __ldg(keys_in + first + i);
    } else {
#pragma unroll
        for (int i = 0; i < kNsItems; i++) key[i] = (first + i < n) ? __ldg(keys_in + first + i) : (K)0;
    }
}

// One stable 4-bit split of the tile.  Input: this thread's 12 consecutive elements (valid while their tile position
// is < valid).  Output: every valid element scattered to its position in `sk`, ordered by nibble, ties in input order.
template <typename K, int DMODE, bool HI, bool FULL>
__device__ __forceinline__ void ns_split(const K (&key)[kNsItems], K *sk, unsigned short *tab /*[32][16]*/, unsigned *dtot /*[16]*/,
                                         unsigned valid, int shift, const Transform &tf)
{
    const unsigned tid = threadIdx.x, lane = tid & 31u, hl = tid & 15u, vw = tid >> 4, warp = tid >> 5;
    const unsigned p0 = tid * kNsItems;

    // ---- count this thread's keys per nibble in registers: sixteen 4-bit counters (max 12) in one 64-bit word;
    //      the count seen by each key BEFORE it is added is its rank among the thread's equal nibbles ----
    unsigned long long cnt = 0, rloc = 0;
    unsigned nibs = 0;  // the 12 nibbles, 4 bits each... 48 bits needed -> two words
    unsigned nibs_hi = 0;
#pragma unroll
    for (int i = 0; i < kNsItems; i++) {
        const unsigned d = pass_digit<K, DMODE>(key[i], shift, tf);
        const unsigned nib = HI ? (d >> 4) : (d & 15u);
        if (i < 8) nibs |= nib << (4 * i);
        else nibs_hi |= nib << (4 * (i - 8));
        if (FULL || p0 + i < valid) {
            const unsigned sh = nib * 4;
            rloc |= ((cnt >> sh) & 15ull) << (4 * i);
            cnt += 1ull << sh;
        }
    }
    // ---- expand to sixteen byte-wide counters in four registers (nibbles 4q .. 4q+3 in e[q]) ----
    unsigned e[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const unsigned x = (unsigned)(cnt >> (16 * q)) & 0xffffu;
        e[q] = (x & 0xfu) | ((x & 0xf0u) << 4) | ((x & 0xf00u) << 8) | ((x & 0xf000u) << 12);
    }
    // ---- inclusive scan over the 16 lanes of the virtual warp (packed byte adds: every field stays <= 192) ----
    unsigned inc[4] = {e[0], e[1], e[2], e[3]};
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned o = __shfl_up_sync(0xffffffffu, inc[q], d, 16);
            if ((int)hl >= d) inc[q] += o;
        }
    }
    if (hl == 15) {  // virtual-warp totals per nibble
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int j = 0; j < 4; j++) tab[vw * 16 + q * 4 + j] = (unsigned short)((inc[q] >> (8 * j)) & 0xffu);
        }
    }
    unsigned exc[4];  // keys with the same nibble in lower lanes of this virtual warp
#pragma unroll
    for (int q = 0; q < 4; q++) exc[q] = inc[q] - e[q];
    __syncthreads();

    // ---- exclusive scan of the 16 x 32 (nibble, virtual warp) totals in nibble-major order:
    //      warp v (< 16) scans nibble v over the 32 virtual warps ----
    unsigned my_tot = 0, my_inc = 0;
    if (warp < 16) {
        my_tot = tab[lane * 16 + warp];
        my_inc = my_tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, my_inc, d);
            if ((int)lane >= d) my_inc += o;
        }
        if (lane == 31) dtot[warp] = my_inc;
    }
    __syncthreads();
    if (warp < 16) {
        unsigned base = 0;
#pragma unroll
        for (int v = 0; v < 16; v++) base += (v < (int)warp) ? dtot[v] : 0u;
        tab[lane * 16 + warp] = (unsigned short)(base + my_inc - my_tot);
    }
    __syncthreads();

    // ---- scatter: position = base of (nibble, virtual warp) + equal nibbles in lower lanes + rank inside the thread ----
    const unsigned short *my_tab = tab + vw * 16;
#pragma unroll
    for (int i = 0; i < kNsItems; i++) {
        if (FULL || p0 + i < valid) {
            const unsigned nib = ((i < 8) ? (nibs >> (4 * i)) : (nibs_hi >> (4 * (i - 8)))) & 15u;
            const unsigned lo = __byte_perm(exc[0], exc[1], nib & 7u);
            const unsigned hi = __byte_perm(exc[2], exc[3], nib & 7u);
            const unsigned below = ((nib & 8u) ? hi : lo) & 0xffu;
            const unsigned pos = (unsigned)my_tab[nib] + below + (unsigned)((rloc >> (4 * i)) & 15ull);
            sk[pos] = key[i];
        }
    }
}

template <typename K, int DMODE, int LBATCH, bool FULL>
__device__ __forceinline__ void ns_tile(const K *__restrict__ keys_in, K *__restrict__ keys_out, const unsigned *__restrict__ digit_base,
                                        unsigned long long *lookback, unsigned epoch, size_t n, size_t num_tiles, int shift,
                                        const Transform &tf, size_t tile, unsigned char *smem_raw, K (&key)[kNsItems])
{
    typedef NsSmem<K> L;
    K *sk = reinterpret_cast<K *>(smem_raw);
    unsigned short *tab = reinterpret_cast<unsigned short *>(smem_raw + L::kKeys);
    unsigned *dstart = reinterpret_cast<unsigned *>(smem_raw + L::kKeys + L::kTab);
    unsigned *dend = dstart + kRadixSize;
    unsigned *out_base = dend + kRadixSize;
    unsigned *dtot = out_base + kRadixSize;  // [16]

    const unsigned tid = threadIdx.x;
    const size_t tile_base = tile * (size_t)kNsTile;
    const unsigned valid = FULL ? (unsigned)kNsTile : (unsigned)(n - tile_base);
    const unsigned p0 = tid * kNsItems;

    if (tid < kRadixSize) { dstart[tid] = 0; dend[tid] = 0; }

    // ---- split 1: by the low nibble of the digit ----
    ns_split<K, DMODE, false, FULL>(key, sk, tab, dtot, valid, shift, tf);
    __syncthreads();
    // ---- split 2: by the high nibble (stable, so the tile ends up sorted by the whole digit) ----
#pragma unroll
    for (int i = 0; i < kNsItems; i++) key[i] = sk[p0 + i];
    // (ns_split synchronises before anybody scatters, so every thread has re-read its keys by then)
    ns_split<K, DMODE, true, FULL>(key, sk, tab, dtot, valid, shift, tf);
    __syncthreads();

    // ---- prefetch the next tile of this CTA into the key registers (they are dead until the next iteration) ----
    {
        const size_t next = tile + gridDim.x;
        K peek[kNsItems];
        // digit runs of the sorted tile: a run starts where the digit differs from its predecessor
        unsigned dprev = 0xffffffffu;
        if (tid > 0 && (FULL || p0 - 1 < valid)) dprev = pass_digit<K, DMODE>(sk[p0 - 1], shift, tf);
#pragma unroll
        for (int i = 0; i < kNsItems; i++) peek[i] = sk[p0 + i];
#pragma unroll
        for (int i = 0; i < kNsItems; i++) {
            if (FULL || p0 + i < valid) {
                const unsigned d = pass_digit<K, DMODE>(peek[i], shift, tf);
                if (d != dprev) {
                    dstart[d] = p0 + i;
                    if (dprev != 0xffffffffu) dend[dprev] = p0 + i;
                }
                dprev = d;
                if (p0 + i == valid - 1) dend[d] = valid;
            }
        }
        if (next < num_tiles) ns_load_tile<K>(keys_in, n, next, key);
    }
    __syncthreads();

    // ---- publish the tile's digit counts, batched decoupled look-back (one thread per digit value) ----
    if (tid < kRadixSize) {
        const unsigned my_start = dstart[tid];
        const unsigned count = dend[tid] - my_start;
        const unsigned status = (tile == 0) ? kLbInclusive : kLbPartial;
        st_relaxed_u64(lookback + tile * kRadixSize + tid, ((unsigned long long)((epoch << 2) | status) << 32) | count);
        unsigned excl = 0;
        if (tile > 0) {
            constexpr int LB = LBATCH;
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
                        if ((tag >> 2) == epoch) {
                            excl += (unsigned)w[b];
                            consumed = b + 1;
                            done = (tag & 3u) == kLbInclusive;
                        }
                    }
                }
                j -= consumed;
                if (consumed == 0) __nanosleep(40);
            }
            st_relaxed_u64(lookback + tile * kRadixSize + tid,
                           ((unsigned long long)((epoch << 2) | kLbInclusive) << 32) | (unsigned)(excl + count));
        }
        out_base[tid] = __ldg(digit_base + tid) + excl - my_start;
    }
    __syncthreads();

    // ---- write keys: consecutive threads -> consecutive addresses inside each digit run ----
#pragma unroll
    for (int i = 0; i < kNsItems; i++) {
        const unsigned p = i * kNsThreads + tid;
        if (FULL || p < valid) {
            const K k = sk[p];
            const unsigned d = pass_digit<K, DMODE>(k, shift, tf);
            keys_out[(size_t)(out_base[d] + p)] = k;
        }
    }
}

template <typename K, int DMODE, int LBATCH>
__global__ void __launch_bounds__(kNsThreads, 2)
onesweep_pass_ns(const K *__restrict__ keys_in, K *__restrict__ keys_out, const unsigned *__restrict__ digit_base,
                 unsigned long long *lookback, unsigned epoch, size_t n, size_t num_tiles, int shift, Transform tf)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t tile = blockIdx.x;
    K key[kNsItems];
    if (tile < num_tiles) ns_load_tile<K>(keys_in, n, tile, key);
    for (; tile < num_tiles; tile += gridDim.x) {
        if ((tile + 1) * (size_t)kNsTile <= n)
            ns_tile<K, DMODE, LBATCH, true>(keys_in, keys_out, digit_base, lookback, epoch, n, num_tiles, shift, tf, tile, smem_raw, key);
        else
            ns_tile<K, DMODE, LBATCH, false>(keys_in, keys_out, digit_base, lookback, epoch, n, num_tiles, shift, tf, tile, smem_raw, key);
        __syncthreads();  // the store phase is done with shared memory
    }
}

template <typename K, int DMODE>
static int ns_launch_typed(StreamState *st, const void *kin, void *kout, const unsigned *base, unsigned long long *lookback, size_t n,
                           int shift, const Transform &tf)
{
    typedef NsSmem<K> L;
    auto kernel = onesweep_pass_ns<K, DMODE, kLookbackBatch>;
    static int resident[64] = {};
    int per_sm = (st->device < 64) ? resident[st->device] : 0;
    if (per_sm == 0) {
        BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kBytes));
        BCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kNsThreads, L::kBytes));
        if (per_sm < 1) per_sm = 1;
        if (st->device < 64) resident[st->device] = per_sm;
    }
    const size_t tiles = (n + kNsTile - 1) / kNsTile;
    size_t grid = (size_t)st->sm_count * (size_t)per_sm;
    if (grid > tiles) grid = tiles;
    unsigned epoch;
    BCB_TRY(next_epoch(st, &epoch));
    LaunchTimer timer(st, BCB_K_ONESWEEP_PASS);
    kernel<<<(unsigned)grid, kNsThreads, L::kBytes, st->stream>>>((const K *)kin, (K *)kout, base, lookback, epoch, n, tiles, shift, tf);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

template <typename K>
static int ns_launch_mode(StreamState *st, const void *kin, void *kout, const unsigned *base, unsigned long long *lookback, size_t n,
                          int shift, const Transform &tf, int digit_mode)
{
    switch (digit_mode) {
    case kDigitIdent: return ns_launch_typed<K, kDigitIdent>(st, kin, kout, base, lookback, n, shift, tf);
    case kDigitTransform: return ns_launch_typed<K, kDigitTransform>(st, kin, kout, base, lookback, n, shift, tf);
    case kDigitSplit: return ns_launch_typed<K, kDigitSplit>(st, kin, kout, base, lookback, n, shift, tf);
    default: return BCB_EINVAL;
    }
}

int ns_launch_pass(StreamState *st, int key_bytes, const void *kin, void *kout, const unsigned *base, unsigned long long *lookback,
                   size_t n, int shift, const Transform &tf, int digit_mode)
{
    switch (key_bytes) {
    case 4: return ns_launch_mode<unsigned>(st, kin, kout, base, lookback, n, shift, tf, digit_mode);
    case 8: return ns_launch_mode<unsigned long long>(st, kin, kout, base, lookback, n, shift, tf, digit_mode);
    default: return BCB_EUNSUPPORTED;  // 8- and 16-bit keys stay on the atomic-OR kernel
    }
}

}  // namespace bcb
