// common.cuh -- shared helpers for the sm_100a kernels behind include/compute_b200.h
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/compute_b200.h"

namespace bcb {

#define BCB_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t _e = (expr);                             \
        if (_e != cudaSuccess) { (void)cudaGetLastError(); return (int)_e; } \
    } while (0)

#define BCB_TRY(expr)                      \
    do {                                   \
        int _s = (expr);                   \
        if (_s != BCB_SUCCESS) return _s;  \
    } while (0)

inline size_t dtype_size(int dtype)
{
    switch (dtype) {
    case BCB_CHAR: case BCB_UCHAR: return 1;
    case BCB_SHORT: case BCB_USHORT: return 2;
    case BCB_INT: case BCB_UINT: case BCB_FLOAT: return 4;
    case BCB_LONG: case BCB_ULONG: case BCB_DOUBLE: return 8;
    default: return 0;
    }
}
inline bool dtype_is_float(int d) { return d == BCB_FLOAT || d == BCB_DOUBLE; }
inline bool dtype_is_signed_int(int d) { return d == BCB_CHAR || d == BCB_SHORT || d == BCB_INT || d == BCB_LONG; }

// ---- per-stream scratch (runtime.cu) -------------------------------------------------------
// Grow-only, stream-ordered (cudaMallocAsync / cudaFreeAsync on the owning stream), so a
// launcher never frees memory a still-running kernel of an earlier call may touch.
struct StreamState {
    int device = 0;
    cudaStream_t stream = nullptr;
    // bulk scratch (temporary key/value buffers, partials ...)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    // small persistent control block, zeroed once at creation:
    //   [0]   u64 ticket counter (monotonic across calls; launchers pass the base)
    //   [1]   int flag of the speculative sort: set by the verification kernel when the result is not sorted
    //   [2]   u64 number of speculative sorts that fell back to the deterministic kernel
    //   [3]   u32 blocks-done counter of the one-launch reductions (reset by the last block)
    //   [4]   u64 element counter of count_if
    //   [5]   carry of a block-distributed scan, folded on the device (bcb_scan_with_carry)
    //   [6]   u64 arrival counter of the grid barriers of the one-launch small sort (monotonic; launchers pass the base)
    unsigned long long *control = nullptr;
    unsigned long long ticket_base = 0;
    unsigned long long gridbar_base = 0;  // arrivals handed out so far on the grid-barrier counter (control[6])
    // decoupled look-back descriptors (scan + sort); zeroed at (re)allocation, validated by epoch tags.  One arena per
    // descriptor LAYOUT, each with its own epoch counter: a slot is only ever read under the layout it was written in,
    // so a stale word can never alias a valid tag of another layout (kArenaPacked: u64 {tag:32 | payload:32} words of
    // the sort and the <= 4-byte scans; kArenaWide: 32-byte {status, partial, inclusive} records of the 8-byte scans)
    struct LookbackArena {
        void *mem = nullptr;
        size_t bytes = 0;
        uint32_t epoch = 0;  // 30-bit generation tag, bumped once per launch that uses the arena
    };
    LookbackArena arena[3];
    // radix digit histograms / bases
    uint32_t *hist = nullptr;  // [8 passes][256]
    // pinned, device-mapped result slot for host-returning calls
    void *pinned_slot = nullptr;      // host address
    void *pinned_slot_dev = nullptr;  // device alias
    int sm_count = 0;
    unsigned long long spec_runs = 0;  // speculative keys-only sorts enqueued (fallbacks are counted on the device)
    // optional per-kernel timing (bcb_timing_*): CUDA event pairs recorded around each launch
    bool timing = false;
    struct TimedLaunch { int kind; cudaEvent_t start, stop; };
    TimedLaunch *timed = nullptr;
    int timed_count = 0, timed_capacity = 0;
};

// RAII helper used by the launchers: records an event pair around a kernel launch when timing is on
struct LaunchTimer {
    StreamState *st;
    int slot = -1;
    LaunchTimer(StreamState *s, int kind);
    ~LaunchTimer();
};

int stream_state(cudaStream_t stream, StreamState **out);
// Copies between PAGEABLE host memory and the device, staged by the library instead of the driver: several host threads
// move chunks through their own pinned slots and streams, so the memcpy into / out of pinned memory runs at the speed of
// several cores and overlaps the DMA (the driver's own staging of a pageable cudaMemcpy is single-threaded: ~14 GB/s
// against ~50 GB/s of the link).  Blocking; the device range must be ready (the caller has synchronised with whatever
// produced it).  Returns BCB_EUNSUPPORTED when the range is not pageable or too small to be worth it: the caller then
// uses cudaMemcpyAsync.
int staged_copy_pageable(void *device_ptr, void *host_ptr, size_t bytes, bool to_device);
bool staged_copy_eligible(const void *host_ptr, size_t bytes);
int scratch_reserve(StreamState *st, size_t bytes, void **out);
enum { kArenaPacked = 0, kArenaWide = 1, kArenaSegmented = 2, kArenaCount = 3 };
// Reserve FIRST, then draw the epoch: a (re)allocation zeroes the arena and restarts its epoch counter, so an epoch
// drawn before the last reserve of a launch could be handed out again later.
int lookback_reserve(StreamState *st, int arena, size_t bytes, void **out);
// next epoch tag in [1, 2^30) of the arena: wraps by zeroing the arena
int next_epoch(StreamState *st, int arena, uint32_t *epoch);
// persistent kernels draw tile ids from the per-stream ticket counter: returns the base of `draws` fresh tickets
unsigned long long ticket_reserve(StreamState *st, unsigned long long draws);

constexpr int kControlTicket = 0, kControlSpecFlag = 1, kControlSpecFallbacks = 2, kControlReduceDone = 3, kControlCount = 4,
              kControlCarry = 5 /* carry of a block-distributed scan (bcb_scan_with_carry) */,
              kControlGridBar = 6 /* monotonic arrival counter of the one-launch small sort's grid barriers */;

// ---- device helpers ------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// single-copy-atomic 64-bit accesses that bypass the (incoherent) L1
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// streaming 128-bit load: read-once data, do not pollute L1
__device__ __forceinline__ uint4 ld_stream_v4(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}


// L2 eviction-priority policies for loads that carry a cache hint (createpolicy, PTX 7.4+)
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ld_hint_v4(const void *p, unsigned long long policy)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ unsigned ld_hint(const unsigned *p, unsigned long long policy)
{
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
    return r;
}
__device__ __forceinline__ unsigned long long ld_hint(const unsigned long long *p, unsigned long long policy)
{
    unsigned long long r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(policy));
    return r;
}

#endif  // __CUDACC__

}  // namespace bcb
