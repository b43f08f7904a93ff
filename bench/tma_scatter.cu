// tma_scatter.cu -- what bounds the write phase of a radix pass on B200, and can the bulk-copy engine carry it?
// Skeleton of an onesweep pass WITHOUT the ranking: persistent CTAs stage tiles into shared memory with cp.async.bulk,
// pretend the stage is the digit-sorted tile (D runs of R keys) and write every run to its own region of the output
// -- the access pattern of the scatter phase.  Parameters (runtime unless noted):
//   THREADS x TILE (template)   384 x 12288 (2 CTAs / SM) or 768 x 24576 (1 CTA / SM)
//   D                           runs per tile (R = TILE / D keys each)
//   scatter                     1: run (t, d) goes to region d (the radix pattern); 0: runs stay in input order (contiguous)
//   misalign                    1: run starts are misaligned by (d + t) % 4 keys, 0: every run starts 16-byte aligned
//   store                       0 none, 1 LSU (LDS key + LDS out_base + STG per key, the r01 write phase),
//                               2 bulk bodies only, 3 bulk bodies + edges by LSU (lane = digit x key slot),
//                               4 bulk bodies + edges as byte-masked 16-byte bulk copies (cp.async.bulk ... .cp_mask)
//   load                        0 none, 1 bulk loads
// Prints GB/s over the bytes actually moved.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tma_scatter tma_scatter.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store(void *gmem_dst, const void *smem_src, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_masked16(void *gmem_dst, const void *smem_src, unsigned short mask)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Params {
    unsigned D, R;
    int scatter, misalign, store, load;
    size_t region;  // keys per digit region (scatter)
};

template <int THREADS, int TILE>
__global__ void __launch_bounds__(THREADS) skeleton(const unsigned *__restrict__ in, unsigned *__restrict__ out, size_t tiles, Params P)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int SLOT = TILE + 64;  // keys per buffer (slack: the emulated run shift reads past the tile)
    unsigned *buf[2] = {reinterpret_cast<unsigned *>(smem), reinterpret_cast<unsigned *>(smem) + SLOT};
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 2 * SLOT * 4);
    unsigned *out_base = reinterpret_cast<unsigned *>(bar + 2);  // [256]
    const unsigned tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const unsigned D = P.D, R = P.R;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (b < tiles && P.load) {
            mbar_expect_tx(&bar[0], TILE * 4);
            tma_load_1d(buf[0], in + (size_t)b * TILE, TILE * 4, &bar[0]);
        }
    }
    __syncthreads();
    unsigned it = 0;
    for (size_t tile = b; tile < tiles; tile += G, ++it) {
        const int s = it & 1;
        if (P.store >= 2) tma_wait_read0();  // the other buffer's stores have read it
        __syncthreads();
        if (tid == 0 && tile + G < tiles && P.load) {
            mbar_expect_tx(&bar[s ^ 1], TILE * 4);
            tma_load_1d(buf[s ^ 1], in + (tile + G) * TILE, TILE * 4, &bar[s ^ 1]);
        }
        if (P.load) mbar_wait(&bar[s], (it >> 1) & 1u);
        if (P.store == 0) continue;
        const unsigned *sorted = buf[s];
        // 16-byte aligned slot of run (tile, d) in the output, in keys; the run itself starts m keys later
        auto slot_of = [&](unsigned d) -> size_t { return P.scatter ? (size_t)d * P.region + tile * R : (tile * D + d) * (size_t)R; };
        auto mis_of = [&](unsigned d) -> unsigned { return P.misalign ? (d & 3u) : 0u; };  // constant per digit: runs of consecutive tiles abut exactly
        if (P.store == 1) {
            if (tid < D) out_base[tid] = (unsigned)(slot_of(tid) + mis_of(tid)) - tid * R;
            __syncthreads();
            const unsigned inv = 0xffffffffu / R + 1;
#pragma unroll 8
            for (int i = 0; i < TILE / THREADS; i++) {
                const unsigned p = i * THREADS + tid;
                const unsigned k = sorted[p];
                const unsigned d = __umulhi(p, inv);  // p / R (the real kernel derives it from the key)
                out[(size_t)out_base[d] + p + (k & 0)] = k;
            }
            __syncthreads();
            continue;
        }
        fence_proxy_async();  // (the real kernel has generic-proxy writes to order before the bulk reads)
        {
            // one op per thread: thread i -> digit i / 3, part i % 3 (0 body, 1 head, 2 tail), so that every warp issues
            // at most 32 bulk copies per tile (a divergent cp.async.bulk is a per-lane loop of ~130 cycles per copy)
            const unsigned d = tid / 3, part = tid - d * 3;
            if (d < D) {
                const unsigned m = mis_of(d), e = m + R;
                unsigned *dst = out + slot_of(d);         // aligned base; the run covers keys [m, e) from here
                const unsigned *src = sorted + d * R;     // congruent (emulated) shared-memory position of the same base
                const unsigned first = (m + 3) >> 2, last = e >> 2;  // full 16-byte chunks [first, last)
                if (part == 0 && last > first) tma_store(dst + first * 4, src + first * 4, (last - first) * 16);
                if (P.store == 4) {
                    if (part == 1 && m) tma_store_masked16(dst, src, (unsigned short)(0xffffu << (4 * m)));
                    if (part == 2 && (e & 3u)) tma_store_masked16(dst + (e & ~3u), src + (e & ~3u), (unsigned short)((1u << (4 * (e & 3u))) - 1u));
                }
                tma_commit();
            }
        }
        if (P.store == 3) {
            // edge keys by LSU: lane -> (digit, key slot), 6 slots per digit, 5 digits per warp instruction
            const unsigned warp = tid >> 5, lane = tid & 31u;
            const unsigned sub = lane / 6, slot = lane - sub * 6;
            for (unsigned d0 = warp * 5; d0 < D; d0 += (THREADS / 32) * 5) {
                const unsigned d = d0 + sub;
                if (sub < 5 && d < D) {
                    const unsigned m = mis_of(d), e = m + R;
                    unsigned *dst = out + slot_of(d);
                    const unsigned *src = sorted + d * R;
                    if (slot < 3) { if (m && m + slot < 4) dst[m + slot] = src[m + slot]; }
                    else if (slot - 3 < (e & 3u)) dst[(e & ~3u) + slot - 3] = src[(e & ~3u) + slot - 3];
                }
            }
        }
    }
    if (P.store >= 2) tma_wait_read0();
}

template <int THREADS, int TILE>
static void run(const unsigned *in, unsigned *out, size_t n, int sms, unsigned D, int scatter, int misalign, int store, int load)
{
    const size_t tiles = n / TILE;
    Params P;
    P.D = D; P.R = TILE / D; P.scatter = scatter; P.misalign = misalign; P.store = store; P.load = load;
    P.region = ((tiles * P.R + 8 + 63) / 64) * 64;
    const size_t smem = 2 * (TILE + 64) * 4 + 16 + 1024 + 64;
    cudaFuncSetAttribute(skeleton<THREADS, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, skeleton<THREADS, TILE>, THREADS, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a);
        skeleton<THREADS, TILE><<<sms * per_sm, THREADS, smem>>>(in, out, tiles, P);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double bytes = (load ? 1.0 : 0.0) * tiles * TILE * 4 + (store ? 1.0 : 0.0) * tiles * TILE * 4;
    static const char *names[] = {"none", "LSU", "bulk-bodies", "bulk+LSU-edges", "bulk+masked-edges"};
    printf("tile %5d x%d/SM D=%3u R=%3u scatter=%d misalign=%d load=%d store=%-18s %8.3f ms %8.1f GB/s %s\n", TILE, per_sm, D, P.R, scatter, misalign, load,
           names[store], best, bytes / best * 1e-6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char **argv)
{
    const size_t n = (size_t)1 << (argc > 1 ? atoi(argv[1]) : 28);
    unsigned *in, *out;
    cudaMalloc(&in, n * 4);
    cudaMalloc(&out, (n + (1 << 22)) * 4);
    cudaMemset(in, 1, n * 4);
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("n = %zu keys, %d SMs\n", n, sms);
    // needs THREADS >= 3 * D for the one-op-per-thread mapping: 768 threads, D = 256
    for (int store : {1, 3, 4}) run<768, 24576>(in, out, n, sms, 256, 1, 1, store, 1);
    for (int store : {1, 3, 4}) run<768, 24576>(in, out, n, sms, 256, 1, 1, store, 0);
    for (int store : {1, 3, 4}) run<768, 12288>(in, out, n, sms, 256, 1, 1, store, 1);
    for (int store : {3, 4}) run<768, 12288>(in, out, n, sms, 256, 1, 1, store, 0);
    for (int store : {2}) run<768, 24576>(in, out, n, sms, 256, 0, 0, store, 0);
    for (int store : {2}) run<768, 24576>(in, out, n, sms, 256, 1, 0, store, 0);
    return 0;
}
