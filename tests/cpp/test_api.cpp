// Exercises the header-only boost::compute layer (include/boost/compute) end to end on a CUDA device, the way the
// reference's Boost.Test files use it: vectors from host ranges, algorithms on begin()/end() with a queue, results
// checked against the reference's golden vectors (file:line cited per case) and against std:: algorithms.
#include <algorithm>
#include <array>
#include <stdexcept>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include <boost/compute.hpp>
#include <boost/compute/experimental/sort_by_transform.hpp>

namespace compute = boost::compute;

static int g_failures = 0;
static int g_checks = 0;

#define CHECK(cond)                                                                     \
    do {                                                                                \
        ++g_checks;                                                                     \
        if(!(cond)){                                                                    \
            ++g_failures;                                                               \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);               \
        }                                                                               \
    } while(0)

template<class T>
static std::vector<T> to_host(const compute::vector<T> &v, compute::command_queue &queue)
{
    std::vector<T> h(v.size());
    compute::copy(v.begin(), v.end(), h.begin(), queue);
    return h;
}

template<class T, size_t N>
static bool equals(const std::vector<T> &got, const T (&expected)[N])
{
    return got.size() == N && std::equal(got.begin(), got.end(), expected);
}

static void test_core(compute::command_queue &queue)
{
    compute::device device = compute::system::default_device();
    CHECK(compute::system::device_count() >= 1);
    CHECK(!device.name().empty());
    CHECK(device.type() & compute::device::gpu);
    CHECK(device.compute_units() > 0);
    CHECK(queue.get_device() == device);
    CHECK(queue.get_context() == compute::system::default_context());
    CHECK(compute::system::find_device(device.name()) == device);
    bool threw = false;
    try { compute::system::find_device("no such device"); } catch(compute::no_device_found &) { threw = true; }
    CHECK(threw);
    compute::opencl_error err(BCB_EINVAL);
    CHECK(err.error_code() == BCB_EINVAL);
    CHECK(std::string(err.what()).find("invalid") != std::string::npos);
}

static void test_vector(compute::command_queue &queue)
{
    // test_vector.cpp:42-167 style checks
    compute::vector<int> v(compute::system::default_context());
    CHECK(v.size() == 0 && v.empty());
    v.push_back(1, queue);
    v.push_back(3, queue);
    v.push_back(5, queue);
    v.push_back(7, queue);
    v.push_back(9, queue);
    CHECK(v.size() == 5);
    CHECK(int(v[0]) == 1 && int(v[4]) == 9 && int(v.front()) == 1 && int(v.back()) == 9);
    v.resize(3, queue);
    CHECK(v.size() == 3 && int(v.back()) == 5);
    v.resize(1000, queue);
    CHECK(v.size() == 1000 && v.capacity() >= 1000 && int(v[2]) == 5);
    v[2] = 42;
    CHECK(int(v[2]) == 42);
    compute::vector<float> f(size_t(10), 9.f, queue);
    CHECK(float(f[0]) == 9.f && float(f[9]) == 9.f);
    compute::fill(f.begin(), f.end(), 2.5f, queue);
    CHECK(float(f[5]) == 2.5f);
    compute::vector<int> io(8, compute::system::default_context());
    compute::iota(io.begin(), io.end(), 3, queue);
    const int expected_iota[] = {3, 4, 5, 6, 7, 8, 9, 10};
    CHECK(equals(to_host(io, queue), expected_iota));
    compute::vector<int> copy_of(io);
    CHECK(equals(to_host(copy_of, queue), expected_iota));
    compute::vector<int> dst(8, compute::system::default_context());
    compute::copy(io.begin() + 2, io.end(), dst.begin(), queue);   // device -> device
    CHECK(int(dst[0]) == 5 && int(dst[5]) == 10);
    CHECK((io.begin() + 3).read(queue) == 6);
    CHECK(io.end() - io.begin() == 8);
    {   // a large std::vector (pageable memory, > 32 MB): the copies are staged by the library; sort in between
        const size_t n = (size_t(9) << 20) + 12345;
        std::vector<unsigned> big(n);
        unsigned x = 12345u;
        for (size_t i = 0; i < n; i++) { x = x * 1664525u + 1013904223u; big[i] = x; }
        compute::vector<unsigned> dev(big.begin(), big.end(), queue);
        std::vector<unsigned> back(n);
        compute::copy(dev.begin(), dev.end(), back.begin(), queue);
        queue.finish();
        CHECK(back == big);
        compute::sort(dev.begin(), dev.end(), queue);
        compute::copy(dev.begin(), dev.end(), back.begin(), queue);
        queue.finish();
        std::sort(big.begin(), big.end());
        CHECK(back == big);
    }
}

static void test_sort(compute::command_queue &queue)
{
    {   // test_sort.cpp:135-145
        int data[] = {-4, 152, -5000, 963, 75321, -456, 0, 1112};
        compute::vector<int> v(data, data + 8, queue);
        CHECK(!compute::is_sorted(v.begin(), v.end(), queue));
        compute::sort(v.begin(), v.end(), queue);
        CHECK(compute::is_sorted(v.begin(), v.end(), queue));
        const int expected[] = {-5000, -456, -4, 0, 152, 963, 1112, 75321};
        CHECK(equals(to_host(v, queue), expected));
        // test_sort.cpp:227-236
        compute::copy(data, data + 8, v.begin(), queue);
        compute::sort(v.begin(), v.end(), compute::greater<int>(), queue);
        const int expected_desc[] = {75321, 1112, 963, 152, 0, -4, -456, -5000};
        CHECK(equals(to_host(v, queue), expected_desc));
        CHECK(compute::is_sorted(v.begin(), v.end(), compute::greater<int>(), queue));
    }
    {   // test_radix_sort.cpp:175-202 float keys incl. +-0
        float data[] = {-6023.0f, 152.5f, -63.0f, 1234567.0f, 11.2f, -5000.1f, 0.0f, 14.0f, -8.25f, -0.0f};
        compute::vector<float> v(data, data + 10, queue);
        compute::detail::radix_sort(v.begin(), v.end(), queue);
        const float expected[] = {-6023.0f, -5000.1f, -63.0f, -8.25f, -0.0f, 0.0f, 11.2f, 14.0f, 152.5f, 1234567.0f};
        std::vector<float> got = to_host(v, queue);
        CHECK(equals(got, expected));
        CHECK(std::signbit(got[4]) && !std::signbit(got[5]));  // -0.0 before +0.0 under the radix transform
    }
    {   // test_radix_sort.cpp:530-541 sub-range
        int data[] = {9, 8, 7, 6, 5, 4, 3, 2, 1, 0};
        compute::vector<int> v(data, data + 10, queue);
        compute::detail::radix_sort(v.begin() + 2, v.end() - 2, queue);
        const int expected[] = {9, 8, 2, 3, 4, 5, 6, 7, 1, 0};
        CHECK(equals(to_host(v, queue), expected));
    }
    {   // test_sort.cpp:286-292 host range
        int data[] = {5, 2, 3, 6, 7, 4, 0, 1};
        std::vector<int> host(data, data + 8);
        compute::sort(host.begin(), host.end(), queue);
        const int expected[] = {0, 1, 2, 3, 4, 5, 6, 7};
        CHECK(equals(host, expected));
    }
    {   // large random sort vs std::sort, all key widths (perf_sort.cpp:84-130 style self-check)
        std::mt19937_64 rng(12345);
        const size_t n = 1000003;
        std::vector<compute::uint_> a(n);
        std::vector<compute::ulong_> b(n);
        std::vector<compute::short_> c(n);
        std::vector<double> d(n);
        for(size_t i = 0; i < n; i++){
            a[i] = compute::uint_(rng()); b[i] = rng(); c[i] = compute::short_(rng());
            d[i] = (double(rng() >> 11) / 9007199254740992.0 - 0.5) * 1e5;
        }
        compute::vector<compute::uint_> da(a.begin(), a.end(), queue);
        compute::vector<compute::ulong_> db(b.begin(), b.end(), queue);
        compute::vector<compute::short_> dc(c.begin(), c.end(), queue);
        compute::vector<double> dd(d.begin(), d.end(), queue);
        compute::sort(da.begin(), da.end(), queue);
        compute::sort(db.begin(), db.end(), compute::greater<compute::ulong_>(), queue);
        compute::stable_sort(dc.begin(), dc.end(), queue);
        compute::sort(dd.begin(), dd.end(), queue);
        queue.finish();
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end(), std::greater<compute::ulong_>());
        std::sort(c.begin(), c.end());
        std::sort(d.begin(), d.end());
        CHECK(to_host(da, queue) == a);
        CHECK(to_host(db, queue) == b);
        CHECK(to_host(dc, queue) == c);
        CHECK(to_host(dd, queue) == d);
    }
}

struct payload16 { int x; int y; float z; float w; };

static void test_sort_by_key(compute::command_queue &queue)
{
    {   // test_radix_sort_by_key.cpp:30-61 stability
        int keys_data[] = {10, 9, 2, 7, 6, -1, 4, 2, 2, 10};
        int values_data[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
        compute::vector<int> keys(keys_data, keys_data + 10, queue);
        compute::vector<int> values(values_data, values_data + 10, queue);
        compute::detail::radix_sort_by_key(keys.begin(), keys.end(), values.begin(), queue);
        const int ek[] = {-1, 2, 2, 2, 4, 6, 7, 9, 10, 10};
        const int ev[] = {6, 3, 8, 9, 7, 5, 4, 2, 1, 10};
        CHECK(equals(to_host(keys, queue), ek));
        CHECK(equals(to_host(values, queue), ev));
        // :63-100 descending
        compute::copy(keys_data, keys_data + 10, keys.begin(), queue);
        compute::copy(values_data, values_data + 10, values.begin(), queue);
        compute::stable_sort_by_key(keys.begin(), keys.end(), values.begin(), compute::greater<int>(), queue);
        const int dk[] = {10, 10, 9, 7, 6, 4, 2, 2, 2, -1};
        const int dv[] = {1, 10, 2, 4, 5, 7, 3, 8, 9, 6};
        CHECK(equals(to_host(keys, queue), dk));
        CHECK(equals(to_host(values, queue), dv));
    }
    {   // test_sort_by_key.cpp:79-91 char payload (small-n insertion path)
        int keys_data[] = {6, 2, 1, 3, 4, 7, 5, 0};
        compute::char_ values_data[] = {'g', 'c', 'b', 'd', 'e', 'h', 'f', 'a'};
        compute::vector<int> keys(keys_data, keys_data + 8, queue);
        compute::vector<compute::char_> values(values_data, values_data + 8, queue);
        compute::sort_by_key(keys.begin(), keys.end(), values.begin(), queue);
        const compute::char_ ev[] = {'a', 'b', 'c', 'd', 'e', 'f', 'g', 'h'};
        CHECK(equals(to_host(values, queue), ev));
    }
    {   // test_sort_by_key.cpp:175-205 16-byte struct payload, n = 1024 reversed keys
        const int n = 1024;
        std::vector<int> hk(n);
        std::vector<payload16> hv(n);
        for(int i = 0; i < n; i++){
            hk[i] = n - i;
            hv[i].x = n - i; hv[i].y = n - i; hv[i].z = hv[i].w = (n - i) / 0.5f;
        }
        compute::vector<int> keys(hk.begin(), hk.end(), queue);
        compute::vector<payload16> values(n, compute::system::default_context());
        compute::copy(&hv[0], &hv[0] + n, values.begin(), queue);
        compute::sort_by_key(keys.begin(), keys.end(), values.begin(), queue);
        std::vector<payload16> out(n);
        compute::copy(values.begin(), values.end(), &out[0], queue);
        bool ok = compute::is_sorted(keys.begin(), keys.end(), queue);
        for(int i = 0; i < n; i++) ok = ok && out[i].x == i + 1 && out[i].w == (i + 1) / 0.5f;
        CHECK(ok);
    }
    {   // perf_sort_by_key.cpp:39-46: int keys + 8-byte values, checked against a stable host sort
        std::mt19937 rng(7);
        const size_t n = 300007;
        std::vector<int> hk(n);
        std::vector<compute::long_> hv(n);
        for(size_t i = 0; i < n; i++){ hk[i] = int(rng() % 1000) - 500; hv[i] = compute::long_(i); }
        compute::vector<int> keys(hk.begin(), hk.end(), queue);
        compute::vector<compute::long_> values(hv.begin(), hv.end(), queue);
        compute::sort_by_key(keys.begin(), keys.end(), values.begin(), queue);
        std::vector<size_t> order(n);
        std::iota(order.begin(), order.end(), size_t(0));
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b){ return hk[a] < hk[b]; });
        std::vector<compute::long_> got = to_host(values, queue);
        bool ok = true;
        for(size_t i = 0; i < n; i++) ok = ok && got[i] == compute::long_(order[i]);
        CHECK(ok);
    }
}

static void test_scan(compute::command_queue &queue)
{
    compute::context context = queue.get_context();
    {   // test_scan.cpp:41-138
        int data[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
        compute::vector<int> v(data, data + 12, queue);
        compute::vector<int> r(12, context);
        compute::vector<int>::iterator end = compute::inclusive_scan(v.begin(), v.end(), r.begin(), queue);
        CHECK(end == r.end());
        const int inc[] = {0, 1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 66};
        CHECK(equals(to_host(r, queue), inc));
        compute::exclusive_scan(v.begin(), v.end(), r.begin(), queue);
        const int exc[] = {0, 0, 1, 3, 6, 10, 15, 21, 28, 36, 45, 55};
        CHECK(equals(to_host(r, queue), exc));
        compute::exclusive_scan(v.begin(), v.end(), v.begin(), queue);  // in place
        CHECK(equals(to_host(v, queue), exc));
    }
    {   // test_scan.cpp:360-390 multiplies with init
        int data[] = {1, 2, 1, 2, 3};
        compute::vector<int> v(data, data + 5, queue);
        compute::vector<int> r(5, context);
        compute::exclusive_scan(v.begin(), v.end(), r.begin(), int(10), compute::multiplies<int>(), queue);
        const int e[] = {10, 10, 20, 20, 40};
        CHECK(equals(to_host(r, queue), e));
        compute::inclusive_scan(v.begin(), v.end(), r.begin(), compute::multiplies<int>(), queue);
        const int i[] = {1, 2, 2, 4, 12};
        CHECK(equals(to_host(r, queue), i));
    }
    {   // test_partial_sum.cpp:28-41
        int data[] = {1, 2, 5, 3, 9, 1, 4, 2};
        compute::vector<int> a(data, data + 8, queue);
        compute::vector<int> b(8, context);
        CHECK(compute::partial_sum(a.begin(), a.end(), b.begin(), queue) == b.end());
        const int e[] = {1, 3, 8, 11, 20, 21, 25, 27};
        CHECK(equals(to_host(b, queue), e));
    }
    {   // perf_exclusive_scan.cpp:54-94: ints in [0,25), checked against std::partial_sum
        std::mt19937 rng(3);
        const size_t n = 5000011;
        std::vector<int> h(n);
        for(size_t i = 0; i < n; i++) h[i] = int(rng() % 25);
        compute::vector<int> v(h.begin(), h.end(), queue);
        compute::vector<int> r(n, context);
        compute::exclusive_scan(v.begin(), v.end(), r.begin(), queue);
        std::vector<int> ref(n);
        std::partial_sum(h.begin(), h.end() - 1, ref.begin() + 1);
        ref[0] = 0;
        CHECK(to_host(r, queue) == ref);
    }
}

static void test_reduce_accumulate(compute::command_queue &queue)
{
    compute::context context = queue.get_context();
    {   // test_reduce.cpp:29-40
        int data[] = {1, 5, 9, 13, 17};
        compute::vector<int> v(data, data + 5, queue);
        int sum = 0, product = 0;
        compute::reduce(v.begin(), v.end(), &sum, compute::plus<int>(), queue);
        compute::reduce(v.begin(), v.end(), &product, compute::multiplies<int>(), queue);
        CHECK(sum == 45 && product == 9945);
        int mn = 0, mx = 0;
        compute::reduce(v.begin(), v.end(), &mn, compute::min<int>(), queue);
        compute::reduce(v.begin(), v.end(), &mx, compute::max<int>(), queue);
        CHECK(mn == 1 && mx == 17);
    }
    {   // test_reduce.cpp:42-49 empty range leaves the result untouched
        compute::vector<short> v(context);
        short sum = 7;
        compute::reduce(v.begin(), v.end(), &sum, queue);
        CHECK(sum == 7);
    }
    {   // test_reduce.cpp:80-88 device result iterators
        int data[] = {1, 2, 3, 4, 5, 6, 7, 8};
        compute::vector<int> in(data, data + 8, queue);
        compute::vector<int> res(2, context);
        compute::reduce(in.begin(), in.begin() + 4, res.begin(), queue);
        compute::reduce(in.begin() + 4, in.end(), res.end() - 1, queue);
        const int e[] = {10, 26};
        CHECK(equals(to_host(res, queue), e));
    }
    {   // test_reduce.cpp:269-277 uchar range accumulated in float
        compute::vector<compute::uchar_> v(context);
        v.push_back(250, queue);
        v.push_back(250, queue);
        float sum = 0;
        compute::reduce(v.begin(), v.end(), &sum, compute::plus<float>(), queue);
        CHECK(sum == 500.f);
    }
    {   // test_accumulate.cpp:25-85,150-206
        int data[] = {2, 4, 6, 8};
        compute::vector<int> v(data, data + 4, queue);
        CHECK(compute::accumulate(v.begin(), v.end(), 0, queue) == 20);
        CHECK(compute::accumulate(v.begin(), v.end(), -10, queue) == 10);
        CHECK(compute::accumulate(v.begin(), v.end(), 2, compute::multiplies<int>(), queue) == 768);
        int q[] = {2, 8, 16};
        compute::vector<int> dq(q, q + 3, queue);
        CHECK(compute::accumulate(dq.begin(), dq.end(), 1024, compute::divides<int>(), queue) == 4);
        compute::vector<int> io(1025, context);
        compute::iota(io.begin(), io.end(), 0, queue);
        CHECK(compute::accumulate(io.begin(), io.end(), 2, queue) == 524802);
        compute::vector<int> empty(context);
        CHECK(compute::accumulate(empty.begin(), empty.end(), 4, queue) == 4);
    }
    {   // test_accumulate.cpp:258-300: int-typed init over float data == std::accumulate
        std::vector<float> h(10000, 1.01f);
        compute::vector<float> v(h.begin(), h.end(), queue);
        CHECK(compute::accumulate(v.begin(), v.end(), 0, queue) == std::accumulate(h.begin(), h.end(), 0));
    }
    {   // perf_accumulate.cpp:27-45 + float reduce tolerance
        std::mt19937 rng(5);
        const size_t n = 4000037;
        std::vector<int> h(n);
        std::vector<float> f(n);
        for(size_t i = 0; i < n; i++){ h[i] = int(rng() % 25); f[i] = float(rng() % 1000) / 1000.f; }
        compute::vector<int> v(h.begin(), h.end(), queue);
        CHECK(compute::accumulate(v.begin(), v.end(), 0, queue) == std::accumulate(h.begin(), h.end(), 0));
        compute::vector<float> df(f.begin(), f.end(), queue);
        float s = 0;
        compute::reduce(df.begin(), df.end(), &s, queue);
        double ref = std::accumulate(f.begin(), f.end(), 0.0);
        CHECK(std::fabs(double(s) - ref) <= 4 * 22 * std::ldexp(1.0, -24) * ref);
    }
}

// container/array.hpp and container/mapped_view.hpp (the containers either side of the path)
static void test_array_and_mapped_view(compute::command_queue &queue)
{
    // test_array.cpp:33-52 -- construct from a host array, read back, fill, element access
    std::array<int, 6> host = {{5, -1, 9, 0, 7, 2}};
    compute::array<int, 6> a(host, queue);
    CHECK(a.size() == 6 && !a.empty());
    CHECK(int(a[2]) == 9 && int(a.front()) == 5 && int(a.back()) == 2);
    compute::sort(a.begin(), a.end(), queue);
    std::vector<int> got(6);
    compute::copy(a.begin(), a.end(), got.begin(), queue);
    const int sorted[] = {-1, 0, 2, 5, 7, 9};
    CHECK(equals(got, sorted));
    CHECK(compute::accumulate(a.begin(), a.end(), 0, queue) == 22);
    compute::array<int, 6> b(a);
    CHECK(int(b[5]) == 9);
    a.fill(3, queue);
    CHECK(compute::accumulate(a.begin(), a.end(), 0, queue) == 18);
    bool threw = false;
    try { a.at(6); } catch(std::out_of_range &) { threw = true; }
    CHECK(threw);

    // test_mapped_view.cpp:29-75 -- algorithms run on host memory in place
    std::vector<unsigned> keys(100000);
    std::mt19937 rng(7);
    for(size_t i = 0; i < keys.size(); i++) keys[i] = rng();
    std::vector<unsigned> expected(keys);
    std::sort(expected.begin(), expected.end());
    {
        compute::mapped_view<unsigned> view(&keys[0], keys.size(), queue.get_context());
        CHECK(view.size() == keys.size());
        compute::sort(view.begin(), view.end(), queue);
        view.map(queue);
        CHECK(keys == expected);
        unsigned long long total = 0;
        for(size_t i = 0; i < keys.size(); i++) total += keys[i];
        unsigned device_total = 0;
        compute::reduce(view.begin(), view.end(), &device_total, queue);
        CHECK(device_total == static_cast<unsigned>(total));
        view.unmap(queue);
    }
    std::vector<int> ones(1000, 1), sums(1000);
    {
        compute::mapped_view<int> in(&ones[0], ones.size()), out(&sums[0], sums.size());
        compute::inclusive_scan(in.begin(), in.end(), out.begin(), queue);
        out.map(queue);
    }
    CHECK(sums[0] == 1 && sums[999] == 1000);
}

// callers of scan / reduce: the reference's own test literals
static void test_scan_and_reduce_callers(compute::command_queue &queue)
{
    using compute::_1;
    compute::context context = queue.get_context();
    {   // test_copy_if.cpp:24-46
        int data[] = { 1, 6, 3, 5, 8, 2, 4 };
        compute::vector<int> input(data, data + 7, queue);
        compute::vector<int> output(input.size(), context);
        compute::fill(output.begin(), output.end(), -1, queue);
        compute::vector<int>::iterator iter = compute::copy_if(input.begin(), input.end(), output.begin(), _1 < 5, queue);
        CHECK(iter == output.begin() + 4);
        CHECK(to_host(output, queue) == (std::vector<int>{1, 3, 2, 4, -1, -1, -1}));
        compute::fill(output.begin(), output.end(), 42, queue);
        iter = compute::copy_if(input.begin(), input.end(), output.begin(), _1 * 2 >= 10, queue);
        CHECK(iter == output.begin() + 3);
        CHECK(to_host(output, queue) == (std::vector<int>{6, 5, 8, 42, 42, 42, 42}));
    }
    {   // test_copy_if.cpp:48-68
        int data[] = { 1, 2, 3, 4, 5, 1, 2, 3, 4, 5 };
        compute::vector<int> input(data, data + 10, queue);
        compute::vector<int> odds(input.size(), context);
        CHECK(compute::copy_if(input.begin(), input.end(), odds.begin(), _1 % 2 == 1, queue) == odds.begin() + 6);
        std::vector<int> h = to_host(odds, queue);
        CHECK((std::vector<int>(h.begin(), h.begin() + 6) == std::vector<int>{1, 3, 5, 1, 3, 5}));
    }
    {   // test_transform_if.cpp:23-39
        int data[] = { -2, -3, -4, -5, -6, -7, -8, -9 };
        compute::vector<int> input(data, data + 8, queue);
        compute::vector<int> output(input.size(), context);
        compute::vector<int>::iterator end = compute::transform_if(input.begin(), input.end(), output.begin(), compute::abs<int>(), _1 % 2 != 0, queue);
        CHECK(end - output.begin() == 4);
        std::vector<int> h = to_host(output, queue);
        CHECK((std::vector<int>(h.begin(), h.begin() + 4) == std::vector<int>{3, 5, 7, 9}));
    }
    {   // test_count.cpp:32-42, :63-72
        int data[] = { 1, 2, 1, 2, 3 };
        compute::vector<int> v(data, data + 5, queue);
        CHECK(compute::count(v.begin(), v.end(), 1, queue) == 2u);
        CHECK(compute::count(v.begin(), v.end(), 3, queue) == 1u);
        CHECK(compute::count(v.begin() + 1, v.end(), 1, queue) == 1u);
        CHECK(compute::count(v.begin() + 1, v.end() - 1, 2, queue) == 2u);
        float fdata[] = { 1.0f, 2.5f, -1.0f, 3.0f, 5.0f };
        compute::vector<float> f(fdata, fdata + 5, queue);
        CHECK(compute::count_if(f.begin(), f.end(), _1 > 2.0f, queue) == 3u);
    }
    {   // test_inner_product.cpp:23-37, test_transform_reduce.cpp:24-40
        int d1[] = { 1, 2, 3, 4 }, d2[] = { 10, 20, 30, 40 };
        compute::vector<int> a(d1, d1 + 4, queue), b(d2, d2 + 4, queue);
        CHECK(compute::inner_product(a.begin(), a.end(), b.begin(), 0, queue) == 300);
        int d3[] = { 1, -2, -3, -4, 5 };
        compute::vector<int> c(d3, d3 + 5, queue);
        int sum = 0;
        compute::transform_reduce(c.begin(), c.end(), &sum, compute::abs<int>(), compute::plus<int>(), queue);
        CHECK(sum == 15);
        compute::vector<int> dev_result(1, context);
        compute::transform_reduce(c.begin(), c.end(), dev_result.begin(), compute::square<int>(), compute::plus<int>(), queue);
        CHECK(to_host(dev_result, queue)[0] == 55);
    }
    {   // test_reduce_by_key.cpp:26-46, :108-132
        int keys[] = { 0, 2, -3, -3, -3, -3, -3, 4 };
        int data[] = { 1, 1, 1, 1, 1, 2, 5, 1 };
        compute::vector<int> k(keys, keys + 8, queue), v(data, data + 8, queue), ko(8, context), vo(8, context);
        std::pair<compute::vector<int>::iterator, compute::vector<int>::iterator> r =
            compute::reduce_by_key(k.begin(), k.end(), v.begin(), ko.begin(), vo.begin(), queue);
        CHECK(r.first == ko.begin() + 4 && r.second == vo.begin() + 4);
        std::vector<int> hk = to_host(ko, queue), hv = to_host(vo, queue);
        CHECK((std::vector<int>(hk.begin(), hk.begin() + 4) == std::vector<int>{0, 2, -3, 4}));
        CHECK((std::vector<int>(hv.begin(), hv.begin() + 4) == std::vector<int>{1, 1, 10, 1}));
        int keys2[] = { 0, 2, 2, 3, 3, 3, 3, 3, 4 };
        int data2[] = { 1, 2, 1, -3, 1, 4, 2, 5, 77 };
        compute::vector<int> k2(keys2, keys2 + 9, queue), v2(data2, data2 + 9, queue), ko2(9, context), vo2(9, context);
        compute::reduce_by_key(k2.begin(), k2.end(), v2.begin(), ko2.begin(), vo2.begin(), compute::min<int>(), compute::equal_to<int>(), queue);
        hv = to_host(vo2, queue);
        CHECK((std::vector<int>(hv.begin(), hv.begin() + 4) == std::vector<int>{1, 1, -3, 77}));
    }
}

// callers of sort
static void test_sort_callers(compute::command_queue &queue)
{
    {   // test_is_permutation.cpp:25-46, :48-69
        int d1[] = {1, 3, 1, 2, 5}, d2[] = {3, 1, 5, 1, 2};
        compute::vector<int> v1(d1, d1 + 5, queue), v2(d2, d2 + 5, queue);
        CHECK(compute::is_permutation(v1.begin(), v1.begin() + 5, v2.begin(), v2.begin() + 5, queue));
        int one = 1;
        compute::copy(&one, &one + 1, v2.begin(), queue);
        CHECK(!compute::is_permutation(v1.begin(), v1.begin() + 5, v2.begin(), v2.begin() + 5, queue));
        const char s1[] = "abade", s2[] = "aadeb";
        compute::vector<char> c1(s1, s1 + 5, queue), c2(s2, s2 + 5, queue);
        CHECK(compute::is_permutation(c1.begin(), c1.end(), c2.begin(), c2.end(), queue));
        CHECK(!compute::is_permutation(c1.begin(), c1.end(), c2.begin(), c2.begin() + 4, queue));
    }
    {   // test_sort_by_transform.cpp:24-39
        int data[] = { 1, -2, 4, -3, 0, 5, -8, -9 };
        compute::vector<int> v(data, data + 8, queue);
        compute::experimental::sort_by_transform(v.begin(), v.end(), compute::abs<int>(), compute::less<int>(), queue);
        CHECK(to_host(v, queue) == (std::vector<int>{0, 1, -2, -3, 4, 5, -8, -9}));
        // a range long enough for the radix path: stable order by |x|
        std::vector<int> big(5000);
        for (size_t i = 0; i < big.size(); i++) big[i] = (int)((i * 7919u) % 201u) - 100;
        compute::vector<int> dv(big.begin(), big.end(), queue);
        compute::experimental::sort_by_transform(dv.begin(), dv.end(), compute::abs<int>(), compute::less<int>(), queue);
        std::stable_sort(big.begin(), big.end(), [](int a, int b) { return std::abs(a) < std::abs(b); });
        CHECK(to_host(dv, queue) == big);
    }
    {   // transform(): unary closed set
        int data[] = { 3, -4, 5 };
        compute::vector<int> in(data, data + 3, queue), out(3, queue.get_context());
        CHECK(compute::transform(in.begin(), in.end(), out.begin(), compute::square<int>(), queue) == out.end());
        CHECK(to_host(out, queue) == (std::vector<int>{9, 16, 25}));
    }
}

// second batch of callers: set operations on sorted ranges, extrema, valarray reductions
// sorts with a custom comparator: the reference's own cases with the comparator spelt as a field comparator
struct Particle
{
    Particle() : x(0.f), y(0.f) {}
    Particle(float _x, float _y) : x(_x), y(_y) {}
    float x;
    float y;
};

static void test_comparator_sorts(compute::command_queue &queue)
{
    using compute::int2_;
    namespace lambda = compute::lambda;
    const compute::context &context = compute::system::default_context();
    {   // test_stable_sort.cpp:41-90: int2_ by the first, then by the second component
        compute::vector<int2_> vec(context);
        vec.push_back(int2_(2, 1), queue);
        vec.push_back(int2_(2, 2), queue);
        vec.push_back(int2_(1, 2), queue);
        vec.push_back(int2_(1, 1), queue);
        CHECK(compute::is_sorted(vec.begin(), vec.end(), compute::less_by_component<0>(), queue) == false);
        compute::stable_sort(vec.begin(), vec.end(), compute::less_by_component<0>(), queue);
        CHECK(compute::is_sorted(vec.begin(), vec.end(), compute::less_by_component<0>(), queue) == true);
        std::vector<int2_> result(vec.size());
        compute::copy(vec.begin(), vec.end(), result.begin(), queue);
        queue.finish();
        CHECK(result[0] == int2_(1, 2) && result[1] == int2_(1, 1) && result[2] == int2_(2, 1) && result[3] == int2_(2, 2));
        compute::stable_sort(vec.begin(), vec.end(), lambda::get<1>(lambda::_1) < lambda::get<1>(lambda::_2), queue);  // a.y < b.y
        compute::copy(vec.begin(), vec.end(), result.begin(), queue);
        queue.finish();
        CHECK(result[0] == int2_(1, 1) && result[1] == int2_(2, 1) && result[2] == int2_(1, 2) && result[3] == int2_(2, 2));
    }
    {   // test_merge_sort_gpu.cpp:330-378: stable merge sort of int2_ by .x
        int2_ data[] = { int2_(8, 3), int2_(5, 1), int2_(2, 1), int2_(6, 1), int2_(8, 1), int2_(7, 1), int2_(4, 1), int2_(8, 2) };
        compute::vector<int2_> vector(data, data + 8, queue);
        CHECK(!compute::is_sorted(vector.begin(), vector.end(), compute::less_by(&int2_::x), queue));
        compute::detail::merge_sort_on_gpu(vector.begin(), vector.end(), compute::less_by(&int2_::x), true /*stable*/, queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by(&int2_::x), queue));
        const int2_ expected[] = { int2_(2, 1), int2_(4, 1), int2_(5, 1), int2_(6, 1), int2_(7, 1), int2_(8, 3), int2_(8, 1), int2_(8, 2) };
        std::vector<int2_> h(8);
        compute::copy(vector.begin(), vector.end(), h.begin(), queue);
        queue.finish();
        CHECK(std::equal(h.begin(), h.end(), expected));
    }
    {   // test_sort.cpp:294-326: a struct by its x member
        std::vector<Particle> particles;
        particles.push_back(Particle(0.1f, 0.f));
        particles.push_back(Particle(-0.4f, 0.f));
        particles.push_back(Particle(10.0f, 0.f));
        particles.push_back(Particle(0.001f, 0.f));
        compute::vector<Particle> vector(4, context);
        compute::copy(particles.begin(), particles.end(), vector.begin(), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by(&Particle::x), queue) == false);
        compute::sort(vector.begin(), vector.end(), compute::less_by(&Particle::x), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by(&Particle::x), queue) == true);
        compute::copy(vector.begin(), vector.end(), particles.begin(), queue);
        queue.finish();
        CHECK(particles[0].x == -0.4f && particles[1].x == 0.001f && particles[2].x == 0.1f && particles[3].x == 10.0f);
        compute::sort(vector.begin(), vector.end(), compute::greater_by(&Particle::x), queue);
        compute::copy(vector.begin(), vector.end(), particles.begin(), queue);
        queue.finish();
        CHECK(particles[0].x == 10.0f && particles[3].x == -0.4f);
    }
    {   // test_sort.cpp:328-360: 100 int2_ by .x, only the ends are pinned
        const size_t size = 100;
        std::vector<int2_> host(size, int2_(0, 0));
        host[0] = int2_(100, 0);
        host[size / 4] = int2_(20, 0);
        host[(size * 3) / 4] = int2_(9, 0);
        host[size - 3] = int2_(-10, 0);
        host[size / 2 + 1] = int2_(-10, -1);
        compute::vector<int2_> vector(size, context);
        compute::copy(host.begin(), host.end(), vector.begin(), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by_component<0>(), queue) == false);
        compute::sort(vector.begin(), vector.end(), compute::less_by_component<0>(), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by_component<0>(), queue) == true);
        compute::copy(vector.begin(), vector.end(), host.begin(), queue);
        queue.finish();
        CHECK(host[0][0] == -10 && host[1][0] == -10 && host[size - 3][0] == 9 && host[size - 2][0] == 20 && host[size - 1][0] == 100);
        CHECK(host[0] != host[1]);
    }
    {   // test_merge_sort_gpu.cpp:223-256: 1024 ints by abs()
        const int size = 1024;
        std::vector<int> data(size);
        for(int i = 0; i < size; i++) data[i] = i % 2 ? size - i : i - size;
        compute::vector<int> vector(data.begin(), data.end(), queue);
        CHECK(!compute::is_sorted(vector.begin(), vector.end(), lambda::abs(lambda::_1) < lambda::abs(lambda::_2), queue));
        compute::detail::merge_sort_on_gpu(vector.begin(), vector.end(), lambda::abs(lambda::_1) < lambda::abs(lambda::_2), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_abs<int>(), queue));
        std::vector<int> h = to_host(vector, queue);
        std::stable_sort(data.begin(), data.end(), [](int a, int b) { return std::abs(a) < std::abs(b); });
        CHECK(h == data);
        // less / greater go to the radix sort (test_merge_sort_gpu.cpp:65-100)
        compute::detail::merge_sort_on_gpu(vector.begin(), vector.end(), compute::greater<int>(), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::greater<int>(), queue));
    }
    {   // long8_ by one component (test_merge_sort_gpu.cpp:294-328 compares a.s0 < b.s3, not an ordering; by s0 here)
        using compute::long8_;
        std::vector<long8_> data(256);
        for(long i = 0; i < 256; i++) data[i] = i % 2 ? long8_(i) : long8_(i * i);
        compute::vector<long8_> vector(data.begin(), data.end(), queue);
        compute::detail::merge_sort_on_gpu(vector.begin(), vector.end(), compute::less_by(&long8_::s0), queue);
        CHECK(compute::is_sorted(vector.begin(), vector.end(), compute::less_by(&long8_::s0), queue));
        std::vector<long8_> h(256);
        compute::copy(vector.begin(), vector.end(), h.begin(), queue);
        queue.finish();
        std::stable_sort(data.begin(), data.end(), [](const long8_ &a, const long8_ &b) { return a.s0 < b.s0; });
        CHECK(std::equal(h.begin(), h.end(), data.begin()));
    }
}

static void test_set_operations_and_extrema(compute::command_queue &queue)
{
    compute::context context = queue.get_context();
    {   // test_set_union.cpp:24-42, test_set_intersection.cpp:24-41, test_set_difference.cpp:24-41, test_set_symmetric_difference.cpp:24-42
        int dataset1[] = {1, 1, 2, 2, 2, 2, 3, 3, 4, 5, 6, 10};
        int dataset2[] = {0, 2, 2, 4, 5, 6, 8, 8, 9, 9, 9, 13};
        compute::vector<int> set1(dataset1, dataset1 + 12, queue), set2(dataset2, dataset2 + 12, queue);
        compute::vector<unsigned> result(19, context);  // (the reference's tests write int_ sets into a uint_ vector)
        compute::vector<unsigned>::iterator iter =
            compute::set_union(set1.begin(), set1.begin() + 12, set2.begin(), set2.begin() + 12, result.begin(), queue);
        CHECK(iter == result.begin() + 19);
        CHECK(to_host(result, queue) == (std::vector<unsigned>{0, 1, 1, 2, 2, 2, 2, 3, 3, 4, 5, 6, 8, 8, 9, 9, 9, 10, 13}));
        compute::vector<int> r2(24, context);
        compute::fill(r2.begin(), r2.end(), -1, queue);
        CHECK(compute::set_intersection(set1.begin(), set1.end(), set2.begin(), set2.end(), r2.begin(), queue) == r2.begin() + 5);
        std::vector<int> h = to_host(r2, queue);
        CHECK((std::vector<int>(h.begin(), h.begin() + 6) == std::vector<int>{2, 2, 4, 5, 6, -1}));
        CHECK(compute::set_difference(set1.begin(), set1.end(), set2.begin(), set2.end(), r2.begin(), queue) == r2.begin() + 7);
        h = to_host(r2, queue);
        CHECK((std::vector<int>(h.begin(), h.begin() + 7) == std::vector<int>{1, 1, 2, 2, 3, 3, 10}));
        CHECK(compute::set_symmetric_difference(set1.begin(), set1.end(), set2.begin(), set2.end(), r2.begin(), queue) == r2.begin() + 14);
        h = to_host(r2, queue);
        CHECK((std::vector<int>(h.begin(), h.begin() + 14) == std::vector<int>{0, 1, 1, 2, 2, 3, 3, 8, 8, 9, 9, 9, 10, 13}));
    }
    {   // test_set_union.cpp:44-62 (strings), and against std::set_* on longer ranges with many duplicates
        const char s1[] = "abcccdddeeff", s2[] = "bccdfgh";
        compute::vector<char> c1(s1, s1 + 12, queue), c2(s2, s2 + 7, queue), out(19, context);
        CHECK(compute::set_union(c1.begin(), c1.end(), c2.begin(), c2.end(), out.begin(), queue) == out.begin() + 14);
        std::vector<char> h = to_host(out, queue);
        CHECK(std::string(h.begin(), h.begin() + 14) == "abcccdddeeffgh");
        std::vector<int> a(70001), b(50003);
        for (size_t i = 0; i < a.size(); i++) a[i] = (int)((i * 2654435761u) % 30011u);
        for (size_t i = 0; i < b.size(); i++) b[i] = (int)((i * 40503u + 17u) % 30011u);
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        compute::vector<int> da(a.begin(), a.end(), queue), db(b.begin(), b.end(), queue), dr(a.size() + b.size(), context);
        std::vector<int> expect(a.size() + b.size());
        expect.resize(std::set_union(a.begin(), a.end(), b.begin(), b.end(), expect.begin()) - expect.begin());
        size_t n = compute::set_union(da.begin(), da.end(), db.begin(), db.end(), dr.begin(), queue) - dr.begin();
        std::vector<int> got = to_host(dr, queue);
        CHECK(n == expect.size() && std::equal(expect.begin(), expect.end(), got.begin()));
        expect.assign(a.size() + b.size(), 0);
        expect.resize(std::set_symmetric_difference(a.begin(), a.end(), b.begin(), b.end(), expect.begin()) - expect.begin());
        n = compute::set_symmetric_difference(da.begin(), da.end(), db.begin(), db.end(), dr.begin(), queue) - dr.begin();
        got = to_host(dr, queue);
        CHECK(n == expect.size() && std::equal(expect.begin(), expect.end(), got.begin()));
    }
    {   // test_extrema.cpp:39-51, :53-75, :259-306
        compute::vector<int> v(size_t(4096), 0, queue);
        CHECK(compute::min_element(v.begin(), v.begin(), queue) == v.begin());
        CHECK(compute::min_element(v.begin(), v.begin() + 1, queue) == v.begin());
        compute::iota(v.begin(), v.begin() + 512, 1, queue);
        compute::fill(v.end() - 512, v.end(), 513, queue);
        CHECK(compute::min_element(v.begin(), v.end(), queue) == v.begin() + 512);
        CHECK(compute::max_element(v.begin(), v.end(), queue) == v.end() - 512);
        std::pair<compute::vector<int>::iterator, compute::vector<int>::iterator> mm = compute::minmax_element(v.begin(), v.end(), compute::less<int>(), queue);
        CHECK(mm.first.read(queue) == 0 && mm.second.read(queue) == 513);
        compute::vector<int> w(5000, context);
        compute::iota(w.begin(), w.end(), 0, queue);
        CHECK(compute::max_element(w.begin(), w.end(), queue) == w.end() - 1);
        CHECK(compute::min_element(w.begin() + 1000, w.end() - 1000, queue) == w.begin() + 1000);
        CHECK(compute::max_element(w.begin() + 1000, w.end() - 1000, queue) == w.begin() + 3999);
    }
    {   // test_valarray.cpp:44-62
        int data[] = { 5, 2, 3, 7, 1, 9, 6, 5 };
        compute::valarray<int> array(data, 8);
        CHECK(array.size() == 8u);
        CHECK((array.min)() == 1 && (array.max)() == 9);
        int d2[] = { 1, 2, 3, 4 };
        compute::valarray<int> a2(d2, 4);
        CHECK(a2.sum() == 10);
    }
}

int main()
{
    try {
        compute::command_queue &queue = compute::system::default_queue();
        std::printf("device: %s\n", queue.get_device().name().c_str());
        test_core(queue);
        test_vector(queue);
        test_sort(queue);
        test_sort_by_key(queue);
        test_scan(queue);
        test_reduce_accumulate(queue);
        test_array_and_mapped_view(queue);
        test_scan_and_reduce_callers(queue);
        test_sort_callers(queue);
        test_comparator_sorts(queue);
        test_set_operations_and_extrema(queue);
        queue.finish();
    } catch(std::exception &e) {
        std::printf("EXCEPTION: %s\n", e.what());
        return 2;
    }
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures ? 1 : 0;
}
