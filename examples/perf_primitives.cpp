// Timing harness over the C++ layer, following the reference's perf protocol (perf/perf.hpp:49-97, perf_sort.cpp:27-47,
// perf_exclusive_scan.cpp:54-73, perf_accumulate.cpp:32-45): per trial the input is re-uploaded outside the timed region,
// the timed region is "algorithm(...); queue.finish();", the minimum over the trials is reported, and the result is
// checked afterwards.  Usage: perf_primitives [log2_n] [trials]
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include <boost/compute.hpp>

namespace compute = boost::compute;

template<class F>
static double min_seconds(int trials, F &&trial)
{
    double best = 1e30;
    for(int t = 0; t < trials; t++){
        best = std::min(best, trial());
    }
    return best;
}

static double seconds_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char **argv)
{
    const int log2n = argc > 1 ? std::atoi(argv[1]) : 24;
    const int trials = argc > 2 ? std::atoi(argv[2]) : 3;
    const size_t n = size_t(1) << log2n;
    compute::command_queue &queue = compute::system::default_queue();
    std::printf("device: %s, n = 2^%d, min of %d trials\n", queue.get_device().name().c_str(), log2n, trials);

    std::mt19937 rng(12345);
    std::vector<unsigned> keys(n);
    for(size_t i = 0; i < n; i++) keys[i] = rng();
    compute::vector<unsigned> d_keys(n, queue.get_context());
    bool ok = true;

    const double t_sort = min_seconds(trials, [&]{
        compute::copy(keys.begin(), keys.end(), d_keys.begin(), queue);
        queue.finish();
        const auto t0 = std::chrono::steady_clock::now();
        compute::sort(d_keys.begin(), d_keys.end(), queue);
        queue.finish();
        return seconds_since(t0);
    });
    ok = ok && compute::is_sorted(d_keys.begin(), d_keys.end(), queue);
    std::printf("sort<uint>          %9.3f ms  %8.2f Gkeys/s\n", t_sort * 1e3, n / t_sort / 1e9);

    std::vector<int> ints(n);
    for(size_t i = 0; i < n; i++) ints[i] = int(rng() % 25);  // perf_exclusive_scan.cpp:22-25
    compute::vector<int> d_in(ints.begin(), ints.end(), queue), d_out(n, queue.get_context());
    const double t_scan = min_seconds(trials, [&]{
        const auto t0 = std::chrono::steady_clock::now();
        compute::exclusive_scan(d_in.begin(), d_in.end(), d_out.begin(), queue);
        queue.finish();
        return seconds_since(t0);
    });
    const long long host_sum = std::accumulate(ints.begin(), ints.end() - 1, 0LL);
    ok = ok && int(d_out.back()) == int(host_sum);  // perf_exclusive_scan.cpp:75-94 checks the last element
    std::printf("exclusive_scan<int> %9.3f ms  %8.1f GB/s\n", t_scan * 1e3, 8.0 * n / t_scan / 1e9);

    int total = 0;
    const double t_acc = min_seconds(trials, [&]{
        const auto t0 = std::chrono::steady_clock::now();
        total = compute::accumulate(d_in.begin(), d_in.end(), 0, queue);
        return seconds_since(t0);
    });
    ok = ok && total == int(host_sum + ints.back());
    std::printf("accumulate<int>     %9.3f ms  %8.1f GB/s\n", t_acc * 1e3, 4.0 * n / t_acc / 1e9);

    std::printf(ok ? "results ok\n" : "RESULTS WRONG\n");
    return ok ? 0 : 1;
}
