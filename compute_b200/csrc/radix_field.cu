// radix_field.cu -- sorts of records by a projected field: bcb_sort_by_field / bcb_is_sorted_by_field, the ahead-of-time
// counterpart of the reference's sorts with a custom comparator (algorithm/sort.hpp:83-106, stable_sort.hpp:34-50,
// detail/merge_sort_on_gpu.hpp:523-572, is_sorted.hpp:39-68).
#include "radix_common.cuh"

#include <type_traits>

namespace bcb {

// ---- sorts of records by a projected field: the ahead-of-time counterpart of the reference's custom comparators ----
// The reference compiles an arbitrary compare(a, b) into its merge sort at run time (sort.hpp:83-106 ->
// detail/merge_sort_on_gpu.hpp:523-572).  There is no run-time compiler here; the comparators its own tests use are all of the
// form  f(a.field) < f(b.field)  (int2_ by .x / .y, a struct by its x member, ints by abs(): test_sort.cpp:294-360,
// test_stable_sort.cpp:41-90, test_merge_sort_gpu.cpp:223-380), which is a stable key-value radix sort: project the
// field into a key array, sort the keys with the records as payload.
// F: field type as stored, K: key type (unsigned counterpart for abs() of a signed integer, as OpenCL's abs() returns)
// reverse: the key of a descending sort -- an order-REVERSING image of the projection (~k for integers, -f for floats), so
// that the stable ascending radix sort of the keys is the stable sort by ">" (the radix sort's own descending transform
// is not used: it reproduces the reference's radix quirks, e.g. INT_MIN first, which a comparator does not have)
template <typename F, typename K, int UN>
__global__ void project_field_kernel(const unsigned char *__restrict__ records, size_t n, size_t stride, size_t offset, K *__restrict__ keys, int aligned,
                                     int reverse)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const unsigned char *p = records + i * stride + offset;
        F f;
        if (aligned) {
            f = *reinterpret_cast<const F *>(p);
        } else {
            unsigned char *q = reinterpret_cast<unsigned char *>(&f);
#pragma unroll
            for (int b = 0; b < (int)sizeof(F); b++) q[b] = p[b];
        }
        K k;
        if constexpr (UN == BCB_UN_ABS) {
            if constexpr (std::is_floating_point<F>::value) k = f < F(0) ? -f : f;
            else if constexpr (std::is_signed<F>::value) k = f < F(0) ? (K)(K(0) - (K)f) : (K)f;  // |INT_MIN| = 2^(w-1), unsigned
            else k = f;
        } else {
            // identity: the key keeps the field's bit pattern (and the caller its dtype)
            unsigned char *kq = reinterpret_cast<unsigned char *>(&k);
            const unsigned char *fq = reinterpret_cast<const unsigned char *>(&f);
#pragma unroll
            for (int b = 0; b < (int)sizeof(F); b++) kq[b] = fq[b];
        }
        if (reverse) {
            if constexpr (std::is_floating_point<K>::value) k = -k;
            else k = (K)~k;
        }
        keys[i] = k;
    }
}

template <typename F, typename K>
static int project_field(StreamState *st, const void *records, size_t n, size_t stride, size_t offset, int unary, void *keys, int reverse)
{
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)st->sm_count * 16;
    if (blocks > cap) blocks = cap;
    const int aligned = (((uintptr_t)records | stride | offset) % sizeof(F)) == 0;
    if (unary == BCB_UN_ABS)
        project_field_kernel<F, K, BCB_UN_ABS><<<(unsigned)blocks, 256, 0, st->stream>>>((const unsigned char *)records, n, stride, offset, (K *)keys, aligned, reverse);
    else
        project_field_kernel<F, K, BCB_UN_IDENTITY><<<(unsigned)blocks, 256, 0, st->stream>>>((const unsigned char *)records, n, stride, offset, (K *)keys, aligned, reverse);
    BCB_CUDA_TRY(cudaGetLastError());
    return BCB_SUCCESS;
}

// projects into a fresh key array (stream-ordered allocation, freed by the caller); *key_dtype = type of the keys
static int project_field_keys(StreamState *st, const void *records, size_t n, size_t stride, size_t offset, int field_dtype, int unary,
                              void **keys, int *key_dtype, int reverse)
{
    const size_t w = dtype_size(field_dtype);
    if (!w || (unary != BCB_UN_IDENTITY && unary != BCB_UN_ABS)) return BCB_EINVAL;
    if (offset + w > stride) return BCB_EINVAL;
    *key_dtype = field_dtype;
    if (unary == BCB_UN_ABS) {
        switch (field_dtype) {
        case BCB_CHAR: *key_dtype = BCB_UCHAR; break;
        case BCB_SHORT: *key_dtype = BCB_USHORT; break;
        case BCB_INT: *key_dtype = BCB_UINT; break;
        case BCB_LONG: *key_dtype = BCB_ULONG; break;
        default: break;
        }
    }
    BCB_CUDA_TRY(cudaMallocAsync(keys, n * w, st->stream));
    int rc;
    switch (field_dtype) {
    case BCB_CHAR: rc = project_field<signed char, unsigned char>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_UCHAR: rc = project_field<unsigned char, unsigned char>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_SHORT: rc = project_field<short, unsigned short>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_USHORT: rc = project_field<unsigned short, unsigned short>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_INT: rc = project_field<int, unsigned>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_UINT: rc = project_field<unsigned, unsigned>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_LONG: rc = project_field<long long, unsigned long long>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_ULONG: rc = project_field<unsigned long long, unsigned long long>(st, records, n, stride, offset, unary, *keys, reverse); break;
    case BCB_FLOAT: rc = project_field<float, float>(st, records, n, stride, offset, unary, *keys, reverse); break;
    default: rc = project_field<double, double>(st, records, n, stride, offset, unary, *keys, reverse); break;
    }
    if (rc != BCB_SUCCESS) (void)cudaFreeAsync(*keys, st->stream);
    return rc;
}

}  // namespace bcb

using namespace bcb;

extern "C" {

int bcb_sort_by_field(bcb_stream stream, void *records, size_t n, size_t record_bytes, size_t field_offset, int field_dtype,
                      int unary, int descending)
{
    if (!dtype_size(field_dtype) || record_bytes == 0) return BCB_EINVAL;
    if (unary != BCB_UN_IDENTITY && unary != BCB_UN_ABS) return BCB_EUNSUPPORTED;
    if (field_offset + dtype_size(field_dtype) > record_bytes) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;
    if (!records) return BCB_EINVAL;
    if (n >= 0xffff0000ull) return BCB_ETOOLARGE;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    void *keys;
    int key_dtype;
    BCB_TRY(project_field_keys(st, records, n, record_bytes, field_offset, field_dtype, unary, &keys, &key_dtype, descending != 0));
    const int rc = radix_sort_device(st, key_dtype, 1, keys, n, records, record_bytes);
    (void)cudaFreeAsync(keys, st->stream);
    return rc;
}

int bcb_is_sorted_by_field(bcb_stream stream, const void *records, size_t n, size_t record_bytes, size_t field_offset, int field_dtype,
                           int unary, int descending, int *result_host)
{
    if (!result_host) return BCB_EINVAL;
    *result_host = 1;
    if (!dtype_size(field_dtype) || record_bytes == 0) return BCB_EINVAL;
    if (unary != BCB_UN_IDENTITY && unary != BCB_UN_ABS) return BCB_EUNSUPPORTED;
    if (field_offset + dtype_size(field_dtype) > record_bytes) return BCB_EINVAL;
    if (n < 2) return BCB_SUCCESS;
    if (!records) return BCB_EINVAL;
    StreamState *st;
    BCB_TRY(stream_state((cudaStream_t)stream, &st));
    void *keys;
    int key_dtype;
    BCB_TRY(project_field_keys(st, records, n, record_bytes, field_offset, field_dtype, unary, &keys, &key_dtype, 0));
    const int rc = bcb_is_sorted(stream, key_dtype, descending, keys, n, result_host);  // native compare of the projections; blocks
    (void)cudaFreeAsync(keys, st->stream);
    return rc;
}

}  // extern "C"
