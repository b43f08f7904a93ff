// peer_bulk.cu -- what bounds the exchange pass of the multi-GPU sort over NVLink: bandwidth of writes from GPU 0 into GPU 1's
// HBM (and, for comparison, into GPU 0's own) as a function of HOW they are issued -- cp.async.bulk shared->global copies of
// run-sized pieces at 16-byte alignment (what onesweep_ws issues: ~672 bytes per digit run), the same pieces aligned to
// 128-byte lines, larger pieces, and warp-wide LSU stores of whole lines.  One process, two devices, peer access enabled.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench/peer_bulk bench/peer_bulk.cu && ./bench/peer_bulk
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../compute_b200/csrc/tma.cuh"

using namespace bcb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

// every thread of the first 8 warps issues `per_thread` bulk copies of `bytes` from the CTA's shared buffer to pseudo-random
// places of dst (a region of `span` bytes); align = 16 or 128; phase16 adds a random multiple of 16 below 128
__global__ void __launch_bounds__(256, 1) bulk_writer(char *dst, size_t span, unsigned bytes, unsigned align, int per_thread)
{
    extern __shared__ __align__(128) unsigned char smem[];
    for (unsigned i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned *>(smem)[i] = i;
    __syncthreads();
    fence_proxy_async();
    const unsigned long long slots = (span - 2 * (size_t)bytes - 256) / align;
    unsigned long long h = mix(((unsigned long long)blockIdx.x << 20) | threadIdx.x);
    for (int k = 0; k < per_thread; k++) {
        h = mix(h + k);
        char *p = dst + (h % slots) * align;
        const unsigned soff = (unsigned)((h >> 40) % ((160 * 1024 - bytes) / 16)) * 16;
        tma_store_issue(p, smem + soff, bytes);
        tma_commit();
        if ((k & 7) == 7) tma_store_wait_read<4>();
    }
    tma_store_wait_read<0>();
}

// a warp stores `lines` consecutive 128-byte lines (uint4 per lane, 4 instructions per 512 bytes ... here: one uint per lane
// per line, the pattern of the round-1 exchange kernel) at pseudo-random line-aligned places
__global__ void __launch_bounds__(1024, 1) lsu_writer(char *dst, size_t span, unsigned lines, int per_warp)
{
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long slots = (span - (size_t)lines * 128 - 256) / 128;
    unsigned long long h = mix(((unsigned long long)blockIdx.x << 20) | warp);
    for (int k = 0; k < per_warp; k++) {
        h = mix(h + k);
        unsigned *p = reinterpret_cast<unsigned *>(dst + (h % slots) * 128);
        for (unsigned l = 0; l < lines; l++) p[l * 32 + lane] = (unsigned)h + l;
    }
}

static float time_kernel(int which, char *dst, size_t span, unsigned a, unsigned b, int reps, size_t *bytes_out)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148;
    float best = 1e30f;
    for (int it = 0; it < 3; it++) {
        CK(cudaEventRecord(e0));
        if (which == 0) {
            bulk_writer<<<grid, 256, 160 * 1024>>>(dst, span, a, b, reps);
            *bytes_out = (size_t)grid * 256 * reps * a;
        } else {
            lsu_writer<<<grid, 1024>>>(dst, span, a, reps);
            *bytes_out = (size_t)grid * 32 * reps * a * 128;
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    const size_t span = (size_t)2 << 30;
    char *local = nullptr, *peer = nullptr;
    CK(cudaSetDevice(0));
    CK(cudaMalloc(&local, span));
    if (ndev > 1) {
        CK(cudaSetDevice(1));
        CK(cudaMalloc(&peer, span));
        CK(cudaSetDevice(0));
        CK(cudaDeviceEnablePeerAccess(1, 0));
    }
    CK(cudaFuncSetAttribute(bulk_writer, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    struct Case { const char *name; int which; unsigned a, b; int reps; };
    const Case cases[] = {
        {"bulk 672 B, 16-byte aligned (a digit run of onesweep_ws)", 0, 672, 16, 2048},
        {"bulk 640 B, 128-byte aligned", 0, 640, 128, 2048},
        {"bulk 1344 B, 16-byte aligned", 0, 1344, 16, 1024},
        {"bulk 4096 B, 128-byte aligned", 0, 4096, 128, 512},
        {"bulk 32768 B, 128-byte aligned", 0, 32768, 128, 64},
        {"LSU warp stores, 1 line (128 B) per place", 1, 1, 0, 4096},
        {"LSU warp stores, 5 lines per place", 1, 5, 0, 1024},
        {"LSU warp stores, 64 lines per place", 1, 64, 0, 128},
    };
    for (int target = 0; target < (peer ? 2 : 1); target++) {
        char *dst = target ? peer : local;
        for (const Case &c : cases) {
            size_t bytes = 0;
            const float ms = time_kernel(c.which, dst, span, c.a, c.b, c.reps, &bytes);
            printf("%-6s %-58s %8.1f GB/s  (%.1f MB in %.3f ms)\n", target ? "peer" : "local", c.name, bytes / 1e6 / ms, bytes / 1e6, ms);
        }
    }
    return 0;
}
