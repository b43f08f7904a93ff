// microbench.cu -- per-SM throughput of the warp-level / shared-memory primitives a radix-sort ranking step can be
// built from, measured on the target GPU.  Prints cycles per warp-instruction per SM (all 4 SMSPs busy).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512;
constexpr int kIters = 2048;
constexpr int kUnroll = 8;

template <int OP>
__global__ void __launch_bounds__(kThreads) kern(unsigned *out, unsigned seed)
{
    __shared__ unsigned sm[8192];
    for (int i = threadIdx.x; i < 8192; i += kThreads) sm[i] = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
    unsigned acc = 0;
    unsigned *row = sm + warp * 256;  // per-warp 256-word row (16 warps x 256 = 4096 words)
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            x = x * 1664525u + 1013904223u;  // LCG: per-lane pseudo-random digit
            const unsigned d = x >> 24;
            if (OP == 0) { acc += d; }                                                     // baseline: LCG only
            if (OP == 1) { acc ^= __ballot_sync(0xffffffffu, d & 1); }                     // vote.ballot
            if (OP == 2) { acc ^= __match_any_sync(0xffffffffu, d); }                      // match.any, ~30 distinct
            if (OP == 3) { acc ^= __match_any_sync(0xffffffffu, d & 3); }                  // match.any, 4 distinct
            if (OP == 4) { acc ^= __match_any_sync(0xffffffffu, d & 0); }                  // match.any, 1 distinct
            if (OP == 5) { acc ^= __shfl_sync(0xffffffffu, x, d & 31); }                   // shfl idx
            if (OP == 6) { atomicAdd(&row[d], 1u); }                                       // ATOMS/RED add, random bin
            if (OP == 7) { acc += atomicAdd(&row[d], 1u); }                                // ATOMS add w/ return
            if (OP == 8) { atomicOr(&row[d], 1u << lane); }                                // ATOMS or
            if (OP == 9) { acc += row[d]; }                                                // LDS random
            if (OP == 10) { row[d] = x; }                                                  // STS random
            if (OP == 11) { acc += __popc(x) + __ffs(x); }                                 // popc + ffs
            if (OP == 12) { acc += __reduce_add_sync(0xffffffffu, d); }                    // redux
            if (OP == 13) { acc += row[lane + (d & 7) * 32]; }                             // LDS conflict-free
            if (OP == 14) { atomicAdd(&row[lane + (d & 7) * 32], 1u); }                    // ATOMS conflict-free
            if (OP == 15) { atomicAdd(&row[7], 1u); }                                      // ATOMS same address
            if (OP == 16) {                                                                // 8-ballot match (CUB style)
                unsigned peers = 0xffffffffu;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1);
                    peers &= ((d >> b) & 1) ? m : ~m;
                }
                acc ^= peers;
            }
            if (OP == 17) { atomicAdd((unsigned long long *)&row[(d & 127) * 2], 1ull << 32 | (1ull << lane)); }  // ATOMS.64 add
            if (OP == 18) { acc += ((unsigned long long *)row)[d & 127] >> 32; }           // LDS.64 random
            if (OP == 19) { __syncwarp(); acc += d; }                                      // syncwarp
        }
    }
    if (acc == 0x12345678u) out[0] = acc + sm[threadIdx.x];
}

template <int OP>
void run(const char *name, unsigned *out, int sms)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = sms * 4;  // 4 x 512 threads = full occupancy
    kern<OP><<<blocks, kThreads>>>(out, 1);
    cudaEventRecord(a);
    kern<OP><<<blocks, kThreads>>>(out, 2);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;                       // at nominal max clock
    const double warp_instr_per_sm = 4.0 * (kThreads / 32) * (double)kIters * kUnroll;
    printf("%-34s %8.3f ms  %7.2f cycles per warp-op per SM\n", name, ms, cycles / warp_instr_per_sm);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned *out;
    cudaMalloc(&out, 1024);
    printf("SMs %d; numbers include the LCG baseline (subtract row 0)\n", sms);
    run<0>("baseline (LCG only)", out, sms);
    run<1>("vote.ballot", out, sms);
    run<2>("match.any (~30 distinct)", out, sms);
    run<3>("match.any (4 distinct)", out, sms);
    run<4>("match.any (1 distinct)", out, sms);
    run<5>("shfl.idx", out, sms);
    run<6>("atoms.add random bin (no return)", out, sms);
    run<7>("atoms.add random bin (return)", out, sms);
    run<8>("atoms.or random bin", out, sms);
    run<9>("lds random", out, sms);
    run<10>("sts random", out, sms);
    run<11>("popc+ffs", out, sms);
    run<12>("redux.add", out, sms);
    run<13>("lds conflict-free", out, sms);
    run<14>("atoms.add conflict-free", out, sms);
    run<15>("atoms.add same address", out, sms);
    run<16>("8-ballot match", out, sms);
    run<17>("atoms.add.64 random", out, sms);
    run<18>("lds.64 random", out, sms);
    run<19>("syncwarp", out, sms);
    return 0;
}
