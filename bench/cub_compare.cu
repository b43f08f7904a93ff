// cub_compare.cu -- same-GPU competitor numbers (CUB from the CUDA toolkit), the modern equivalent of the
// reference's own Thrust comparisons (perf/perf_thrust_sort.cu:33-41).  Not part of the product; bench only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cub_compare cub_compare.cu
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void fill_random(unsigned *p, size_t n, unsigned seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long z = (i + seed) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        p[i] = (unsigned)(z ^ (z >> 31));
    }
}

template <class F> float time_min(F f, int reps, void (*reset)(void *), void *ctx)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        if (reset) reset(ctx);
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv)
{
    const int log2n = argc > 1 ? atoi(argv[1]) : 30;
    const size_t n = (size_t)1 << log2n;
    unsigned *src, *a, *b;
    CK(cudaMalloc(&src, n * 4)); CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
    fill_random<<<148 * 8, 256>>>(src, n, 12345);
    void *tmp = nullptr; size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, a, b, (long long)n));
    CK(cudaMalloc(&tmp, tmp_bytes));
    struct Ctx { unsigned *src, *a; size_t n; } ctx{src, a, n};
    auto reset = [](void *c) { Ctx *x = (Ctx *)c; cudaMemcpy(x->a, x->src, x->n * 4, cudaMemcpyDeviceToDevice); };
    float ms = time_min([&] { cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, a, b, (long long)n); }, 5, reset, &ctx);
    printf("{\"cub\": \"DeviceRadixSort::SortKeys u32\", \"log2n\": %d, \"ms\": %.4f, \"Gkeys_s\": %.3f}\n", log2n, ms, n / ms / 1e6);
    cudaFree(tmp);

    // pairs u32/u32 at 2^28, scan / reduce int32 at 2^28
    const size_t m = (size_t)1 << (log2n < 28 ? log2n : 28);
    unsigned *va = b, *vb = nullptr, *kb = nullptr;
    CK(cudaMalloc(&vb, m * 4)); CK(cudaMalloc(&kb, m * 4));
    tmp = nullptr; tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, a, kb, va, vb, (long long)m));
    CK(cudaMalloc(&tmp, tmp_bytes));
    Ctx ctx2{src, a, m};
    ms = time_min([&] { cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, a, kb, va, vb, (long long)m); }, 5, reset, &ctx2);
    printf("{\"cub\": \"DeviceRadixSort::SortPairs u32+u32\", \"log2n\": 28, \"ms\": %.4f, \"Gkeys_s\": %.3f}\n", ms, m / ms / 1e6);
    cudaFree(tmp);

    int *in = (int *)src, *out = (int *)a;
    tmp = nullptr; tmp_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (long long)m));
    CK(cudaMalloc(&tmp, tmp_bytes));
    ms = time_min([&] { cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (long long)m); }, 6, nullptr, nullptr);
    printf("{\"cub\": \"DeviceScan::ExclusiveSum i32\", \"log2n\": 28, \"ms\": %.4f, \"GB_s\": %.1f}\n", ms, m * 8 / ms / 1e6);
    cudaFree(tmp);
    tmp = nullptr; tmp_bytes = 0;
    CK(cub::DeviceReduce::Sum(nullptr, tmp_bytes, in, out, (long long)m));
    CK(cudaMalloc(&tmp, tmp_bytes));
    ms = time_min([&] { cub::DeviceReduce::Sum(tmp, tmp_bytes, in, out, (long long)m); }, 6, nullptr, nullptr);
    printf("{\"cub\": \"DeviceReduce::Sum i32\", \"log2n\": 28, \"ms\": %.4f, \"GB_s\": %.1f}\n", ms, m * 4 / ms / 1e6);
    return 0;
}
