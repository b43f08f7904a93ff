// tile_state.cuh -- tile descriptors and the warp-parallel decoupled look-back shared by the single-pass kernels
// (scan.cu, stream_ops.cu): a tile publishes its aggregate, warp 0 walks back over the preceding descriptors 32 at a time.
#pragma once

#include "ops.cuh"

#include <cstring>

namespace bcb {

constexpr int kScanMaxWindows = 40;
constexpr unsigned kSpinBackoffNs = 40;                // pause between polls of a descriptor that is not published yet                    // look-back windows buffered for the ordered fp fold

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
    // one 32-byte record per tile, the same for every 8-byte type and every tile count: a slot is only ever read
    // as what it was written as (its own arena, see StreamState::arena), so stale bytes can never pose as a tag
    struct Record { unsigned status; unsigned pad; T partial; T inclusive; unsigned long long pad2; };
    static_assert(sizeof(Record) == 32, "record layout is part of the arena contract");
    Record *rec;
    static size_t bytes(size_t tiles) { return tiles * sizeof(Record); }
    __host__ __device__ void bind(void *mem, size_t) { rec = (Record *)mem; }
    __device__ __forceinline__ void post(size_t tile, unsigned epoch, unsigned st, T v) const
    {
        T *dst = (st == kPartial) ? &rec[tile].partial : &rec[tile].inclusive;
        *((volatile T *)dst) = v;
        st_release_u32(&rec[tile].status, (epoch << 2) | st);
    }
    __device__ __forceinline__ unsigned peek(size_t tile, unsigned epoch, T &v) const
    {
        const unsigned tag = ld_acquire_u32(&rec[tile].status);
        if ((tag >> 2) != epoch) return kInvalid;
        const unsigned st = tag & 3u;
        const T *src = (st == kPartial) ? &rec[tile].partial : &rec[tile].inclusive;
        v = *((volatile const T *)src);
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
            while ((st = ts.peek((size_t)idx, epoch, val)) == kInvalid) __nanosleep(kSpinBackoffNs);
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


}  // namespace bcb
