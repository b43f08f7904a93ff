// radix_common.cuh -- pieces shared by the radix-sort translation units (radix_sort.cu, radix_pass_ws.cu)
#pragma once

#include "ops.cuh"

namespace bcb {

constexpr int kRadixBits = 8;
constexpr int kRadixSize = 1 << kRadixBits;
constexpr int kHistThreads = 512;
constexpr int kLookbackBatch = 8;  // default look-back batch

enum : unsigned { kLbInvalid = 0u, kLbPartial = 1u, kLbInclusive = 2u };

// order-preserving transform parameters (uniform): key' = ((x ^ nm) - nm) ^ xc ^ (asr(x) & fa)
constexpr int kMaxSplitters = 7;  // multi-GPU partition pass: up to 8 destinations
struct Transform {
    unsigned long long nm;  // all-ones: negate x first (descending signed / float)
    unsigned long long xc;  // xor constant: sign bit (signed, float) or all-ones (descending unsigned)
    unsigned long long fa;  // float only: bits below the sign, selected when x is negative
    // splitter mode (digit = number of splitters <= transformed key): used by the multi-GPU partition pass
    unsigned long long split[kMaxSplitters];
    int nsplit;
    // splitter mode only: where bucket b goes (device addresses; other GPUs' memory when the multi-GPU sort scatters
    // straight into its peers' receive buffers over NVLink)
    unsigned long long dst_keys[kMaxSplitters + 1];
    unsigned long long dst_vals[kMaxSplitters + 1];
};

inline Transform make_transform(int dtype, bool ascending)
{
    const unsigned w = (unsigned)dtype_size(dtype) * 8;
    const unsigned long long ones = (w == 64) ? ~0ull : ((1ull << w) - 1);
    const unsigned long long sign = 1ull << (w - 1);
    Transform t{};
    const bool sgn = dtype_is_signed_int(dtype), flt = dtype_is_float(dtype);
    if (sgn || flt) {
        t.xc = sign;
        if (!ascending) t.nm = ~0ull;
        if (flt) t.fa = ones & ~sign;
    } else if (!ascending) {
        t.xc = ones;
    }
    return t;
}

template <typename K> struct key_traits;
template <> struct key_traits<unsigned char> { typedef unsigned U; typedef int S; };
template <> struct key_traits<unsigned short> { typedef unsigned U; typedef int S; };
template <> struct key_traits<unsigned> { typedef unsigned U; typedef int S; };
template <> struct key_traits<unsigned long long> { typedef unsigned long long U; typedef long long S; };

template <typename K>
__device__ __forceinline__ unsigned digit_of(K raw, int shift, const Transform &tf)
{
    typedef typename key_traits<K>::U U;
    typedef typename key_traits<K>::S S;
    const U x = (U)raw;
    const U nm = (U)tf.nm;
    // asr over the compute width: only meaningful (fa != 0) for float / double keys, whose width IS the compute width
    const U neg = (U)((S)x >> (sizeof(U) * 8 - 1));
    const U t = ((x ^ nm) - nm) ^ (U)tf.xc ^ (neg & (U)tf.fa);
    return (unsigned)(t >> shift) & (kRadixSize - 1);
}

// full transformed key (all digits), for comparisons in the common unsigned order
template <typename K>
__device__ __forceinline__ unsigned long long transformed_key(K raw, const Transform &tf)
{
    typedef typename key_traits<K>::U U;
    typedef typename key_traits<K>::S S;
    const U x = (U)raw;
    const U nm = (U)tf.nm;
    const U neg = (U)((S)x >> (sizeof(U) * 8 - 1));
    const U t = ((x ^ nm) - nm) ^ (U)tf.xc ^ (neg & (U)tf.fa);
    const unsigned long long ones = sizeof(K) == 8 ? ~0ull : ((1ull << (sizeof(K) * 8 % 64)) - 1);
    return (unsigned long long)t & ones;
}

// one thread per splitter: lower bound of the splitter in a sorted range, compared in the transformed space
template <typename K>
__global__ void partition_points_kernel(const K *__restrict__ keys, size_t n, const unsigned long long *__restrict__ splitters,
                                        unsigned num, unsigned long long *__restrict__ points, Transform tf)
{
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num) return;
    const unsigned long long s = splitters[j];
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        if (transformed_key<K>(keys[mid], tf) < s) lo = mid + 1;
        else hi = mid;
    }
    points[j] = lo;
}

// transformed key in the key's own compute width, and its inverse (all transforms except descending float / double,
// which is not injective -- radix_sort.hpp:100-127 -- and never takes this path)
template <typename K>
__device__ __forceinline__ typename key_traits<K>::U transform_fwd(K raw, const Transform &tf)
{
    typedef typename key_traits<K>::U U;
    typedef typename key_traits<K>::S S;
    const U x = (U)raw;
    const U nm = (U)tf.nm;
    const U neg = (U)((S)x >> (sizeof(U) * 8 - 1));
    return ((x ^ nm) - nm) ^ (U)tf.xc ^ (neg & (U)tf.fa);
}
template <typename K>
__device__ __forceinline__ K transform_inv(typename key_traits<K>::U t, const Transform &tf)
{
    typedef typename key_traits<K>::U U;
    if (tf.fa) {  // ascending float: originally non-negative keys carry the sign bit now
        const U sign = (U)tf.xc;
        return (K)((t & sign) ? (t ^ sign) : ~t);
    }
    const U nm = (U)tf.nm;
    return (K)(((t ^ (U)tf.xc) + nm) ^ nm);
}
// what a pass does with the key transform: kXfNone -- the keys in memory are already in sortable (transformed or
// unsigned-ascending) form; kXfIn -- first pass: transform on load, store transformed; kXfOut -- last pass: store the
// original bit pattern again; kXfBoth -- transform only to extract the digit (keys stay raw in memory)
enum { kXfNone = 0, kXfIn = 1, kXfOut = 2, kXfBoth = 3 };

// digit modes of the pass kernel: plain bit field (unsigned ascending), transformed bit field, splitter bucket
enum { kDigitIdent = 1, kDigitTransform = 0, kDigitSplit = 2 };

template <typename K, int IDENT>
__device__ __forceinline__ unsigned pass_digit(K raw, int shift, const Transform &tf)
{
    if constexpr (IDENT == kDigitIdent) {
        return (unsigned)(raw >> shift) & (kRadixSize - 1);
    } else if constexpr (IDENT == kDigitSplit) {
        // compares in the key's own compute width (splitters of narrower keys fit it by construction)
        typedef typename key_traits<K>::U U;
        const U t = (U)transformed_key<K>(raw, tf);
        // unused splitters are all-ones (split_transform): they can only count for t == all-ones, whose bucket is the
        // last one anyway, hence the clamp instead of a "j < nsplit" test per compare
        // (nsplit is uniform: 2 ranks pay for one compare, 4 ranks for three, 8 ranks for seven)
        unsigned d = (t >= (U)tf.split[0]) ? 1u : 0u;
        if (tf.nsplit > 1) {
            d += ((t >= (U)tf.split[1]) ? 1u : 0u) + ((t >= (U)tf.split[2]) ? 1u : 0u);
            if (tf.nsplit > 3) {
#pragma unroll
                for (int j = 3; j < kMaxSplitters; j++) d += (t >= (U)tf.split[j]) ? 1u : 0u;
            }
        }
        return min(d, (unsigned)tf.nsplit);
    } else {
        return digit_of<K>(raw, shift, tf);
    }
}


// streaming copy of n elements by `nthreads` threads (this one is number `g`), every element through `f`: 128-bit accesses,
// four independent loads in flight per thread when both arrays are 16-byte aligned -- the constant-digit "pass"
template <typename T, typename F>
__device__ __forceinline__ void stream_copy(const T *__restrict__ in, T *__restrict__ out, size_t n, size_t g, size_t nthreads, F f)
{
    constexpr int VEC = 16 / (int)sizeof(T);
    size_t done = 0;
    if (((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0) {
        const size_t nvec = n / VEC;
        for (size_t v = g; v < nvec; v += 4 * nthreads) {
            uint4 x[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (v + u * nthreads < nvec) x[u] = ld_stream_v4(in + (v + u * nthreads) * VEC);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (v + u * nthreads < nvec) {
                    T *e = reinterpret_cast<T *>(&x[u]);
#pragma unroll
                    for (int k = 0; k < VEC; k++) e[k] = f(e[k]);
                    st_stream_v4(out + (v + u * nthreads) * VEC, x[u]);
                }
            }
        }
        done = nvec * VEC;
    }
    for (size_t i = done + g; i < n; i += nthreads) out[i] = f(__ldg(in + i));
}

// warp-specialised bulk-copy pass kernel (radix_pass_ws.cu): large sorts.  Keys only (32- / 64-bit) with the speculative
// two-sweep ranking; 32-bit keys + 4- / 8-byte payload, or keys only with a non-injective transform, with the
// deterministic atomic-OR ranking.  Returns BCB_EUNSUPPORTED for shapes it does not cover (arrays not 16-byte aligned).
int ws_launch_pass(StreamState *st, int key_bytes, const void *kin, void *kout, const void *vin, void *vout, int value_bytes,
                   const unsigned *base, unsigned long long *lookback, size_t n, int shift, const Transform &tf, int xf, bool deterministic,
                   const unsigned long long *dst_tab = nullptr, const uint4 *tile_tab = nullptr, size_t tab_tiles = 0,
                   const unsigned *hot = nullptr);
// hot (device, [2]): digit values of this pass that hold so many keys that same-address shared atomics would serialise
// (digit_scan finds them); the kernel ranks those by ballot.  kNoHotDigit = none.
constexpr unsigned kNoHotDigit = 0xffffffffu;
// hot[1] == kConstDigit: EVERY key has the digit value hot[0] -- the stable pass over such a digit is the identity
// permutation, and both pass kernels turn into a streaming copy (keys < 2^16 in a 32-bit type: two of the four passes)
constexpr unsigned kConstDigit = 0xfffffffeu;
// layout of StreamState::hist (unsigned words): [0, 2048) digit counts, [2048, 4096) digit bases, [4096, 5120) destination
// table of the exchange pass (512 x u64), [5120, 5136) hot digits
constexpr int kHistHotOffset = 5120;
// tile_tab (device, one uint4 per tile: {first key, end, first tile of the segment, segment}): segmented pass -- every
// segment is sorted on its own, base is indexed [segment][256] (see bcb_radix_sort_segments).
// dst_tab (device, [2][256]): the pass is the exchange pass of the multi-GPU sort -- the run of digit value d is written to
// the array at dst_tab[d] (values: dst_tab[256 + d]) starting at element base[d], instead of kout / vout (may be null).
size_t ws_tile_size(int key_bytes, int value_bytes, bool deterministic);
bool ws_supports(int key_bytes, int value_bytes, bool deterministic);

// radix_sort.cu, for the other radix translation units: the device sort behind bcb_radix_sort, the second half of a
// speculative keys-only sort (verification + gated deterministic re-sort; key_bytes 4 or 8), BCB_SORT_SPECULATIVE
int radix_sort_device(StreamState *st, int key_dtype, int ascending, void *keys, size_t n, void *values, size_t value_bytes);
int radix_verify_and_fix(StreamState *st, int key_bytes, void *keys, size_t n, const Transform &tf);
bool radix_speculation_enabled();

// warp-specialised exchange pass of the multi-GPU sort (radix_exchange_ws.cu): bucket b of the stable partition by the
// splitters in tf goes to tf.dst_keys[b] / tf.dst_vals[b].  BCB_EUNSUPPORTED for shapes it does not cover.
int ws_exchange_pass(StreamState *st, int key_bytes, const void *kin, const void *vin, int value_bytes, size_t n, const Transform &tf);

}  // namespace bcb
