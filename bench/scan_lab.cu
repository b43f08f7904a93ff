// scan_lab.cu -- shape sweep of the warp-specialised scan kernel (compute_b200/csrc/scan_ws.cuh) on one GPU.
// For every <vectors per thread, stages, lag> shape: the scan itself (checked in full against a host prefix sum), the
// same pipeline without the inter-CTA prefix chain (MODE 1) and the bare bulk-copy pipeline (MODE 2) -- so a number
// below the copy roofline can be attributed to the chain, the arithmetic or the copy engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I include -o bench/scan_lab bench/scan_lab.cu
// Run:   bench/scan_lab [log2n = 28] [reps = 10]
#include "../compute_b200/csrc/scan_ws.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace bcb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static unsigned g_epoch = 0;

template <typename T, int NV, int S, int D, int MODE>
static float run(const T *in, T *out, size_t n, void *desc, int sms, int reps, int exclusive, T init)
{
    typedef ScanWsShape<T, NV, S> C;
    auto kernel = scan_ws_kernel<T, BCB_PLUS, NV, S, D, MODE>;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
    size_t tiles = (n + C::TILE - 1) / C::TILE;
    WsTileState<T> ts;
    ts.bind(desc);
    size_t grid = (size_t)sms < tiles ? (size_t)sms : tiles;
    if (grid > (size_t)kSwMaxGrid) grid = kSwMaxGrid;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = 0; r < reps + 2; r++) {
        unsigned epoch = ++g_epoch;
        const T *init_dev = nullptr;
        void *args[] = {(void *)&in, (void *)&out, (void *)&n, (void *)&exclusive, (void *)&init, (void *)&ts, (void *)&epoch, (void *)&tiles, (void *)&init_dev};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)kernel, dim3((unsigned)grid), dim3(kSwThreads), args, C::SMEM_BYTES, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    const double gb = 2.0 * n * sizeof(T) / 1e9;
    printf("NV=%d S=%d D=%d mode=%d tile=%6d B  smem=%6zu  avg %.4f ms %7.1f GB/s   best %.4f ms %7.1f GB/s", NV, S, D, MODE, C::TILE_BYTES,
           (size_t)C::SMEM_BYTES, sum / reps, gb / (sum / reps) * 1e3, best, gb / best * 1e3);
    return sum / reps;
}

template <typename T> static bool check(const T *dev_out, const std::vector<T> &ref, size_t n)
{
    std::vector<T> got(n);
    CK(cudaMemcpy(got.data(), dev_out, n * sizeof(T), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++)
        if (got[i] != ref[i]) {
            printf("  MISMATCH at %zu: got %lld want %lld\n", i, (long long)got[i], (long long)ref[i]);
            return false;
        }
    printf("  ok\n");
    return true;
}

template <int NV, int S, int D>
static bool shape(const int *in, int *out, size_t n, void *desc, int sms, int reps, const std::vector<int> &ref_excl)
{
    run<int, NV, S, D, 0>(in, out, n, desc, sms, reps, 1, 7);
    bool ok = check(out, ref_excl, n);
    run<int, NV, S, D, 1>(in, out, n, desc, sms, reps, 1, 7);
    printf("\n");
    return ok;
}

int main(int argc, char **argv)
{
    const int log2n = argc > 1 ? atoi(argv[1]) : 28;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const size_t n = ((size_t)1 << log2n);
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    printf("n = 2^%d int32, %d SMs\n", log2n, sms);
    std::vector<int> h(n), ref(n);
    unsigned long long x = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        h[i] = (int)(x % 25);
    }
    unsigned acc = 7;
    for (size_t i = 0; i < n; i++) { ref[i] = (int)acc; acc += (unsigned)h[i]; }
    int *in, *out;
    void *desc;
    CK(cudaMalloc(&in, n * sizeof(int)));
    CK(cudaMalloc(&out, n * sizeof(int)));
    const size_t desc_bytes = (n / 1024 + 64) * 32;
    CK(cudaMalloc(&desc, desc_bytes));
    CK(cudaMemset(desc, 0, desc_bytes));
    CK(cudaMemcpy(in, h.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    bool ok = true;
    ok &= shape<4, 7, 3>(in, out, n, desc, sms, reps, ref);
    ok &= shape<3, 9, 4>(in, out, n, desc, sms, reps, ref);
    ok &= shape<3, 9, 5>(in, out, n, desc, sms, reps, ref);
    ok &= shape<2, 14, 6>(in, out, n, desc, sms, reps, ref);
    ok &= shape<2, 14, 8>(in, out, n, desc, sms, reps, ref);
    ok &= shape<2, 13, 7>(in, out, n, desc, sms, reps, ref);
    // ragged size + in place
    {
        const size_t m = n - 12345;
        acc = 7;
        CK(cudaMemcpy(out, h.data(), m * sizeof(int), cudaMemcpyHostToDevice));
        run<int, 3, 9, 4, 0>(out, out, m, desc, sms, -1, 1, 7);  // (one launch: it is in place)
        ok &= check(out, ref, m);
    }
    printf(ok ? "ALL_OK\n" : "FAILED\n");
    return ok ? 0 : 1;
}
