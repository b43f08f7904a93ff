// tma_issue.cu -- what does it cost a warp to ISSUE small cp.async.bulk shared->global copies?  (cycles per copy)
//   A: every lane issues its own copy (divergent operands: the compiler emits a per-lane ELECT / R2UR loop)
//   B: lane 0 issues 32 copies in a loop, operands read from shared memory
//   C: lane 0 issues 32 copies whose operands are computed from the loop counter (provably warp-uniform)
// for W = 1, 4, 8, 24 warps of one CTA per SM issuing at the same time.  384-byte copies to contiguous, aligned addresses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_issue tma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store(void *gmem_dst, unsigned smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes) : "memory");
}
constexpr int kPiece = 384, kReps = 6;
template <int MODE>
__global__ void __launch_bounds__(768) issue(unsigned char *out, unsigned long long *cycles, int warps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint4 *ops = (uint4 *)(smem + 98304);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *mine = out + ((size_t)blockIdx.x * 768 + tid) * kPiece * kReps;
    for (int r = 0; r < kReps; r++) {
        const unsigned long long d = (unsigned long long)(mine + (size_t)r * kPiece);
        ops[r * 768 + tid] = make_uint4((unsigned)d, (unsigned)(d >> 32), smem_u32(smem) + (tid * kPiece) % 98304, kPiece);
    }
    __syncthreads();
    if ((int)warp >= warps) return;
    const unsigned long long t0 = clock64();
    for (int r = 0; r < kReps; r++) {
        if (MODE == 0) {
            const uint4 o = ops[r * 768 + tid];
            tma_store((void *)(((unsigned long long)o.y << 32) | o.x), o.z, o.w);
        } else if (MODE == 1) {
            for (int i = 0; i < 32; i++) {
                const uint4 o = ops[r * 768 + warp * 32 + i];
                if (lane == 0) tma_store((void *)(((unsigned long long)o.y << 32) | o.x), o.z, o.w);
            }
        } else {
            for (int i = 0; i < 32; i++) {
                const unsigned idx = warp * 32 + i;
                if (lane == 0) tma_store(out + ((size_t)blockIdx.x * 768 + idx) * kPiece * kReps + (size_t)r * kPiece, smem_u32(smem) + (idx * kPiece) % 98304, kPiece);
            }
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    const unsigned long long t1 = clock64();
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    const unsigned long long t2 = clock64();
    if (lane == 0 && warp == 0) { cycles[blockIdx.x * 2] = t1 - t0; cycles[blockIdx.x * 2 + 1] = t2 - t0; }
}
template <int MODE> void run(const char *name, unsigned char *out, unsigned long long *cyc, int sms)
{
    cudaFuncSetAttribute(issue<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304 + 768 * kReps * 16);
    for (int warps : {1, 4, 8, 24}) {
        issue<MODE><<<sms, 768, 98304 + 768 * kReps * 16>>>(out, cyc, warps);
        cudaDeviceSynchronize();
        unsigned long long h[296];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double a = 0, b = 0;
        for (int i = 0; i < sms; i++) { a += h[2 * i]; b += h[2 * i + 1]; }
        printf("%-28s warps %2d: %7.1f cycles per copy per warp to issue, %7.1f incl. drain   (%s)\n", name, warps, a / sms / (32.0 * kReps), b / sms / (32.0 * kReps),
               cudaGetErrorString(cudaGetLastError()));
    }
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned char *out; unsigned long long *cyc;
    cudaMalloc(&out, (size_t)sms * 768 * kPiece * kReps + 4096);
    cudaMalloc(&cyc, 296 * 8);
    run<0>("A per-lane (divergent)", out, cyc, sms);
    run<1>("B lane 0, operands from smem", out, cyc, sms);
    run<2>("C lane 0, uniform operands", out, cyc, sms);
    return 0;
}
