// radix_exchange.cu -- the multi-GPU sort whose exchange is one of its radix passes (bcb_radix_exchange_scatter on the
// source, bcb_radix_sort_segments on the owner; compute_b200/distributed.py plans the placement).  New functionality: the
// reference has no multi-device path; the single-GPU counterpart is radix_sort_impl (algorithm/detail/radix_sort.hpp:
// 252-426).  Both halves run the warp-specialised pass kernel of radix_pass_ws.cu.
#include "radix_common.cuh"

#include <atomic>
#include <cstdlib>

namespace bcb {

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- multi-GPU sort whose exchange is ONE of its radix passes ---------------------------------------------------
// Pass over the MOST significant digit first, on the source: digit value d is owned by one rank, and the all-gathered
// top-digit histograms tell every source where its run of d starts inside the owner's receive buffer (after the runs of
// the lower ranks: equal keys keep their global input order), so the pass writes every digit run straight into peer
// memory.  The destination then holds one SEGMENT per digit value it owns and sorts every segment by the remaining
// digits with stable LSD passes -- all segments in one launch per digit (tiles never straddle two segments, the
// look-back stops at a segment's first tile, digit bases are per segment).  As many passes over the data as on one
// GPU; the exchange costs NVLink time and one more histogram read, not a pass.
// Both halves use the warp-specialised kernel (one bulk copy per digit run, also into peer memory).

// is the shape covered?  (decided from types and the environment only: every rank of a collective call agrees)
static int exchange_shape(int key_dtype, size_t vb, const Transform &tf, bool *speculative)
{
    const size_t w = dtype_size(key_dtype);
    if (!w) return BCB_EINVAL;
    if (w < 4 || (vb != 0 && vb != 4 && vb != 8) || (vb && w != 4)) return BCB_EUNSUPPORTED;
    const bool injective = !(tf.fa != 0 && tf.nm != 0);
    *speculative = vb == 0 && injective && radix_speculation_enabled();  // verified at the end of bcb_radix_sort_segments
    if (!ws_supports((int)w, (int)vb, !*speculative)) return BCB_EUNSUPPORTED;  // (64-bit keys: speculative flavour only)
    return BCB_SUCCESS;
}

// histograms of the digits below the most significant one, per segment: hist[s][p][256].  One CTA per SM, one counter
// column per lane (conflict-free shared atomics, see radix_histogram_columns); a CTA takes a contiguous share of the
// 16-byte vectors of all segments and flushes its counters whenever it moves on to another segment.
template <typename K, int COLS, bool IDENT>
__global__ void __launch_bounds__(1024, 1)
segment_histogram(const K *__restrict__ keys, const unsigned *__restrict__ seg_begin, const unsigned *__restrict__ seg_len, int num_segments,
                  unsigned *__restrict__ hist, Transform tf)
{
    constexpr int NP = (int)sizeof(K) - 1;
    constexpr int VEC = 16 / (int)sizeof(K);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned *sh = reinterpret_cast<unsigned *>(smem_raw);
    unsigned *mine = sh + (threadIdx.x & (COLS - 1));
    auto count_key = [&](K k) {
#pragma unroll
        for (int p = 0; p < NP; p++) {
            const unsigned d = IDENT ? ((unsigned)(k >> (p * kRadixBits)) & (kRadixSize - 1)) : digit_of<K>(k, p * kRadixBits, tf);
            atomicAdd(mine + (p * kRadixSize + d) * COLS, 1u);
        }
    };
    // this CTA's share [v0, v1) of the vectors of all segments (a segment of len keys has ceil(len / VEC) vectors)
    unsigned long long total = 0;
    for (int s = 0; s < num_segments; s++) total += (seg_len[s] + VEC - 1) / VEC;
    const unsigned long long v0 = total * blockIdx.x / gridDim.x, v1 = total * (blockIdx.x + 1) / gridDim.x;
    unsigned long long seg_v0 = 0;
    for (int s = 0; s < num_segments && seg_v0 < v1; s++) {
        const unsigned len = seg_len[s];
        const unsigned long long nv = (len + VEC - 1) / VEC;
        const unsigned long long a = v0 > seg_v0 ? v0 - seg_v0 : 0, b = (v1 - seg_v0) < nv ? (v1 - seg_v0) : nv;
        seg_v0 += nv;
        if (a >= b) continue;
        for (int i = threadIdx.x; i < NP * kRadixSize * COLS / 4; i += blockDim.x) reinterpret_cast<uint4 *>(sh)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const K *base = keys + seg_begin[s];  // 16-byte aligned (checked by the launcher)
        const unsigned long long full = len / VEC;  // vectors [0, full) are whole
        unsigned long long v = a + threadIdx.x;
        const unsigned long long bf = b < full ? b : full;
        for (; v + blockDim.x < bf; v += 2 * blockDim.x) {  // two independent 128-bit loads in flight
            const uint4 x = ld_stream_v4(base + v * VEC), y = ld_stream_v4(base + (v + blockDim.x) * VEC);
            const K *ex = reinterpret_cast<const K *>(&x), *ey = reinterpret_cast<const K *>(&y);
#pragma unroll
            for (int k = 0; k < VEC; k++) count_key(ex[k]);
#pragma unroll
            for (int k = 0; k < VEC; k++) count_key(ey[k]);
        }
        for (; v < bf; v += blockDim.x) {
            const uint4 x = ld_stream_v4(base + v * VEC);
            const K *ex = reinterpret_cast<const K *>(&x);
#pragma unroll
            for (int k = 0; k < VEC; k++) count_key(ex[k]);
        }
        if (b > full && threadIdx.x < len - full * VEC) count_key(base[full * VEC + threadIdx.x]);  // the partial last vector
        __syncthreads();
        for (int bin = threadIdx.x; bin < NP * kRadixSize; bin += blockDim.x) {
            unsigned c = 0;
#pragma unroll
            for (int j = 0; j < COLS; j++) c += sh[bin * COLS + ((j + threadIdx.x) & (COLS - 1))];  // rotated: conflict-free
            if (c) atomicAdd(hist + (size_t)s * NP * kRadixSize + bin, c);
        }
        __syncthreads();
    }
}

// hist[s][p][d] -> base[p][s][d] = where the run of digit value d of segment s starts in the output of pass p: inside the
// segment's own (aligned) slot for the passes in between, at its compact position in the result for the last pass
__global__ void __launch_bounds__(kRadixSize) segment_digit_scan(const unsigned *__restrict__ hist, unsigned *__restrict__ base,
                                                                 const unsigned *__restrict__ seg_begin, const unsigned *__restrict__ seg_out,
                                                                 int num_segments, int num_passes)
{
    __shared__ unsigned wsum[kRadixSize / 32];
    const unsigned d = threadIdx.x, lane = d & 31u, warp = d >> 5;
    const int p = blockIdx.x, sgm = blockIdx.y;
    const unsigned c = hist[((size_t)sgm * num_passes + p) * kRadixSize + d];
    unsigned incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
        if ((int)lane >= off) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned add = 0;
    for (unsigned w = 0; w < warp; w++) add += wsum[w];
    base[((size_t)p * num_segments + sgm) * kRadixSize + d] = (p == num_passes - 1 ? seg_out[sgm] : seg_begin[sgm]) + incl - c + add;
}

// tile table of a segmented pass: {first key, end, first tile of the segment, segment} per tile
__global__ void segment_tiles(const unsigned *__restrict__ seg_begin, const unsigned *__restrict__ seg_len, const unsigned *__restrict__ seg_tile0,
                              int num_segments, unsigned tile, uint4 *__restrict__ tab)
{
    const int sgm = blockIdx.x;
    if (sgm >= num_segments) return;
    const unsigned begin = seg_begin[sgm], len = seg_len[sgm], t0 = seg_tile0[sgm];
    const unsigned tiles = (len + tile - 1) / tile;
    for (unsigned j = threadIdx.x; j < tiles; j += blockDim.x) {
        const unsigned first = begin + j * tile, last = (j + 1) * tile < len ? first + tile : begin + len;
        tab[t0 + j] = make_uint4(first, last, t0, (unsigned)sgm);
    }
}

template <typename K, int VB>
static int exchange_scatter_typed(StreamState *st, const void *keys, const void *values, size_t n, const Transform &tf, bool spec,
                                  const unsigned long long *dst_tab_host, const unsigned *dst_first_host)
{
    constexpr int NPASS = sizeof(K);
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0, injective = !(tf.fa != 0 && tf.nm != 0);
    const size_t tile = ws_tile_size((int)sizeof(K), VB, !spec);
    void *lb;
    BCB_TRY(lookback_reserve(st, kArenaPacked, ((n + tile - 1) / tile) * kRadixSize * sizeof(unsigned long long), &lb));
    unsigned *base = st->hist + 8 * kRadixSize;
    unsigned long long *dst_tab = reinterpret_cast<unsigned long long *>(st->hist + 16 * kRadixSize);
    // (pageable sources: the runtime stages them before returning, the caller's arrays are free again)
    BCB_CUDA_TRY(cudaMemcpyAsync(base, dst_first_host, kRadixSize * sizeof(unsigned), cudaMemcpyHostToDevice, st->stream));
    BCB_CUDA_TRY(cudaMemcpyAsync(dst_tab, dst_tab_host, 2 * kRadixSize * sizeof(unsigned long long), cudaMemcpyHostToDevice, st->stream));
    // the keys leave in sortable form (the segment passes extract plain bit fields and the last one restores the original
    // bit pattern) unless the transform cannot be inverted: then they stay raw everywhere
    const int xf = ident ? kXfNone : (injective ? kXfIn : kXfBoth);
    return ws_launch_pass(st, (int)sizeof(K), keys, nullptr, values, nullptr, VB, base, (unsigned long long *)lb, n, (NPASS - 1) * kRadixBits, tf,
                          xf, !spec, dst_tab);
}

template <typename K, int VB>
static int sort_segments_typed(StreamState *st, void *recv_keys, void *recv_values, void *out_keys, void *out_values, const Transform &tf,
                               bool spec, const unsigned *seg_host /* [4][S]: begin, len, out, tile0 */, int S, size_t span, size_t n_out,
                               size_t tiles)
{
    constexpr int NP = (int)sizeof(K) - 1;  // passes over the digits below the most significant one
    const bool ident = (tf.nm | tf.xc | tf.fa) == 0, injective = !(tf.fa != 0 && tf.nm != 0);
    const size_t tile = ws_tile_size((int)sizeof(K), VB, !spec);
    // scratch: ping-pong partner of the receive buffer (same segment layout) + the tables
    const size_t kbytes = align_up(span * sizeof(K), 256), vbytes = align_up(span * (size_t)VB, 256);
    const size_t hist_bytes = align_up((size_t)S * NP * kRadixSize * sizeof(unsigned), 256), seg_bytes = align_up((size_t)4 * S * sizeof(unsigned), 256);
    const size_t tab_bytes = align_up(tiles * sizeof(uint4), 256);
    void *scratch;
    BCB_TRY(scratch_reserve(st, kbytes + vbytes + 2 * hist_bytes + seg_bytes + tab_bytes, &scratch));
    char *q = (char *)scratch;
    void *tmp_keys = q; q += kbytes;
    void *tmp_vals = VB ? (void *)q : nullptr; q += vbytes;
    unsigned *hist = (unsigned *)q; q += hist_bytes;
    unsigned *base = (unsigned *)q; q += hist_bytes;
    unsigned *seg = (unsigned *)q; q += seg_bytes;
    uint4 *tab = (uint4 *)q;
    void *lb;
    BCB_TRY(lookback_reserve(st, kArenaPacked, tiles * kRadixSize * sizeof(unsigned long long), &lb));
    BCB_CUDA_TRY(cudaMemcpyAsync(seg, seg_host, (size_t)4 * S * sizeof(unsigned), cudaMemcpyHostToDevice, st->stream));
    BCB_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)S * NP * kRadixSize * sizeof(unsigned), st->stream));
    const unsigned *seg_begin = seg, *seg_len = seg + S, *seg_out = seg + 2 * S, *seg_tile0 = seg + 3 * S;
    {
        // the keys arrive in sortable form unless the transform is not injective (see exchange_scatter_typed)
        const bool plain = ident || injective;
        constexpr int COLS = sizeof(K) == 8 ? 16 : 32;
        constexpr size_t kSmem = (size_t)NP * kRadixSize * COLS * sizeof(unsigned);
        auto kernel = plain ? segment_histogram<K, COLS, true> : segment_histogram<K, COLS, false>;
        static std::atomic<unsigned long long> configured[2];  // bit per device
        const unsigned long long bit = st->device < 64 ? (1ull << st->device) : 0ull;
        if (!(configured[plain].load(std::memory_order_acquire) & bit) || !bit) {
            BCB_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
            configured[plain].fetch_or(bit, std::memory_order_release);
        }
        LaunchTimer timer(st, BCB_K_RADIX_HISTOGRAM);
        kernel<<<(unsigned)st->sm_count, 1024, kSmem, st->stream>>>((const K *)recv_keys, seg_begin, seg_len, S, hist, tf);
    }
    BCB_CUDA_TRY(cudaGetLastError());
    {
        LaunchTimer timer(st, BCB_K_DIGIT_SCAN);
        segment_digit_scan<<<dim3(NP, S), kRadixSize, 0, st->stream>>>(hist, base, seg_begin, seg_out, S, NP);
        segment_tiles<<<S, 128, 0, st->stream>>>(seg_begin, seg_len, seg_tile0, S, (unsigned)tile, tab);
    }
    BCB_CUDA_TRY(cudaGetLastError());
    const void *kin = recv_keys, *vin = recv_values;
    for (int p = 0; p < NP; p++) {
        const bool last = p == NP - 1;
        void *kout = last ? out_keys : (kin == recv_keys ? tmp_keys : recv_keys);
        void *vout = last ? out_values : (kin == recv_keys ? tmp_vals : recv_values);
        const int xf = ident ? kXfNone : (!injective ? kXfBoth : (last ? kXfOut : kXfNone));
        BCB_TRY(ws_launch_pass(st, (int)sizeof(K), kin, kout, vin, vout, VB, base + (size_t)p * S * kRadixSize, (unsigned long long *)lb, span,
                               p * kRadixBits, tf, xf, !spec, nullptr, tab, tiles));
        kin = kout;
        vin = vout;
    }
    if (spec) return radix_verify_and_fix(st, (int)sizeof(K), out_keys, n_out, tf);  // (keys only: exchange_shape)
    return BCB_SUCCESS;
}

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_radix_exchange_scatter(bcb_stream stream, int key_dtype, int ascending, const void *keys, const void *values, size_t value_bytes,
                               size_t n, void *const *dst_keys, void *const *dst_values, const unsigned long long *dst_first)
{
    const size_t vb = values ? value_bytes : 0;
    const Transform tf = make_transform(key_dtype, ascending != 0);
    bool spec = false;
    BCB_TRY(exchange_shape(key_dtype, vb, tf, &spec));
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    if (!dst_keys || !dst_first || (vb && !dst_values)) return BCB_EINVAL;
    if (n && !keys) return BCB_EINVAL;
    if ((((uintptr_t)keys) | ((uintptr_t)values)) & 15) return BCB_EINVAL;  // (bulk copies; the host layer realigns)
    unsigned long long tab[2 * kRadixSize];
    unsigned first[kRadixSize];
    for (int d = 0; d < kRadixSize; d++) {
        tab[d] = (unsigned long long)(uintptr_t)dst_keys[d];
        tab[kRadixSize + d] = vb ? (unsigned long long)(uintptr_t)dst_values[d] : 0ull;
        // the digit runs leave the SM as bulk copies: 16-byte aligned destination arrays
        if (!tab[d] || (tab[d] & 15) || (vb && (!tab[kRadixSize + d] || (tab[kRadixSize + d] & 15)))) return BCB_EINVAL;
        if (dst_first[d] >= 0xffff0000ull) return BCB_ETOOLARGE;
        first[d] = (unsigned)dst_first[d];
    }
    if (n == 0) return BCB_SUCCESS;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    if (dtype_size(key_dtype) == 8) return exchange_scatter_typed<unsigned long long, 0>(st, keys, nullptr, n, tf, spec, tab, first);
    switch (vb) {
    case 0: return exchange_scatter_typed<unsigned, 0>(st, keys, nullptr, n, tf, spec, tab, first);
    case 4: return exchange_scatter_typed<unsigned, 4>(st, keys, values, n, tf, spec, tab, first);
    default: return exchange_scatter_typed<unsigned, 8>(st, keys, values, n, tf, spec, tab, first);
    }
}

int bcb_radix_sort_segments(bcb_stream stream, int key_dtype, int ascending, void *recv_keys, void *recv_values, size_t value_bytes,
                            void *out_keys, void *out_values, const unsigned long long *seg_begin, const unsigned long long *seg_len,
                            size_t num_segments)
{
    const size_t vb = recv_values ? value_bytes : 0;
    const Transform tf = make_transform(key_dtype, ascending != 0);
    bool spec = false;
    BCB_TRY(exchange_shape(key_dtype, vb, tf, &spec));
    if (num_segments > (size_t)kRadixSize) return BCB_EINVAL;
    if (num_segments && (!seg_begin || !seg_len)) return BCB_EINVAL;
    const size_t w = dtype_size(key_dtype);
    const size_t tile = ws_tile_size((int)w, (int)vb, !spec);
    // drop the empty segments; begin / len / compact output position / first tile of each of the others
    unsigned seg[4 * kRadixSize];
    int S = 0;
    for (size_t i = 0; i < num_segments; i++) S += seg_len[i] != 0;
    size_t n_out = 0, tiles = 0, span = 0;
    for (size_t i = 0, j = 0; i < num_segments; i++) {
        if (!seg_len[i]) continue;
        if (seg_begin[i] + seg_len[i] >= 0xffff0000ull || n_out + seg_len[i] >= 0xffff0000ull) return BCB_ETOOLARGE;
        if (seg_begin[i] < span) return BCB_EINVAL;             // ascending, not overlapping
        if ((seg_begin[i] * w) & 15 || (seg_begin[i] * vb) & 15) return BCB_EINVAL;  // the passes read whole 16-byte vectors
        seg[j] = (unsigned)seg_begin[i];
        seg[S + j] = (unsigned)seg_len[i];
        seg[2 * S + j] = (unsigned)n_out;
        seg[3 * S + j] = (unsigned)tiles;
        n_out += seg_len[i];
        tiles += (seg_len[i] + tile - 1) / tile;
        span = seg_begin[i] + seg_len[i];
        j++;
    }
    if (n_out == 0) return BCB_SUCCESS;
    if (!recv_keys || !out_keys || (vb && !out_values)) return BCB_EINVAL;
    if ((((uintptr_t)recv_keys) | ((uintptr_t)recv_values) | ((uintptr_t)out_keys) | ((uintptr_t)out_values)) & 15) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    if (w == 8) return sort_segments_typed<unsigned long long, 0>(st, recv_keys, nullptr, out_keys, nullptr, tf, spec, seg, S, span, n_out, tiles);
    switch (vb) {
    case 0: return sort_segments_typed<unsigned, 0>(st, recv_keys, nullptr, out_keys, nullptr, tf, spec, seg, S, span, n_out, tiles);
    case 4: return sort_segments_typed<unsigned, 4>(st, recv_keys, recv_values, out_keys, out_values, tf, spec, seg, S, span, n_out, tiles);
    default: return sort_segments_typed<unsigned, 8>(st, recv_keys, recv_values, out_keys, out_values, tf, spec, seg, S, span, n_out, tiles);
    }
}

}  // extern "C"
