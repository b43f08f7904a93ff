// stl_baseline.cpp -- TEST / BENCH INFRASTRUCTURE, never shipped: the single-threaded STL baselines the reference
// times next to its own algorithms:
//   C1  perf/perf_stl_sort.cpp:22-30          std::sort over a vector of random ints
//   C2  perf/perf_stl_partial_sum.cpp:31-47   std::partial_sum over ints in [0, 25)
//   C3  perf/perf_stl_accumulate.cpp:34-38    std::accumulate over ints in [0, 25)
// Each function runs the STL call on the caller's buffer and returns its duration in seconds (the reference's
// perf_timer brackets exactly the STL call); min over trials is taken by the caller, as perf_timer::min_time does.
#include <algorithm>
#include <chrono>
#include <cstddef>
#include <cstdint>
#include <numeric>

static inline double seconds_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

extern "C" {

double orc_stl_sort_u32(uint32_t *keys, size_t n)
{
    const auto t0 = std::chrono::steady_clock::now();
    std::sort(keys, keys + n);
    return seconds_since(t0);
}

double orc_stl_partial_sum_i32(const int32_t *in, int32_t *out, size_t n)
{
    const auto t0 = std::chrono::steady_clock::now();
    std::partial_sum(in, in + n, out);
    return seconds_since(t0);
}

double orc_stl_accumulate_i32(const int32_t *in, size_t n, int32_t *result)
{
    const auto t0 = std::chrono::steady_clock::now();
    *result = std::accumulate(in, in + n, int32_t(0));
    return seconds_since(t0);
}

}  // extern "C"
