/*
 * compute_b200.h -- C ABI of the B200-native sort / scan / reduce path.
 *
 * This is the drop-in boundary: plain pointers, sizes and integer codes only.  The
 * header-only C++ layer under include/boost/compute/ (same spellings as the reference)
 * and the Python mirror compute_b200/ both sit on top of exactly these entry points.
 * Every entry point cites the reference interface it replaces; paths are relative to
 * boostorg/compute include/boost/compute/.
 *
 * Conventions
 *  - All device work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the
 *    legacy default stream) and the call returns without waiting, exactly like the
 *    reference's enqueue-and-return algorithms (perf/perf_sort.cpp:38-39); calls that
 *    hand a value back to the host block until it is there (accumulate.hpp:178-188,
 *    reduce.hpp:225).
 *  - Return value: 0 on success, otherwise a cudaError_t value (< 10000) or one of the
 *    BCB_E* codes; nothing throws or aborts.  bcb_error_string() describes either kind.
 *    The C++ layer maps nonzero codes to boost::compute::opencl_error
 *    (exception/opencl_error.hpp:30-61).
 *  - Scratch memory is owned by the library, cached per stream and stream-ordered; the
 *    caller owns keys / values / in / out.
 *  - Element counts are size_t; sorts accept n < 2^32 (the reference narrows to uint_:
 *    algorithm/detail/radix_sort.hpp:351), scan/reduce accept any n.
 *  - There is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef COMPUTE_B200_H
#define COMPUTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* scalar types of types/fundamental.hpp:30-39 (char_ ... double_) */
typedef enum bcb_dtype {
    BCB_CHAR = 0, BCB_UCHAR = 1, BCB_SHORT = 2, BCB_USHORT = 3, BCB_INT = 4,
    BCB_UINT = 5, BCB_LONG = 6, BCB_ULONG = 7, BCB_FLOAT = 8, BCB_DOUBLE = 9
} bcb_dtype;

/* functors of functional/operator.hpp:73-96 that the path accepts */
typedef enum bcb_op {
    BCB_PLUS = 0, BCB_MULTIPLIES = 1, BCB_MIN = 2, BCB_MAX = 3,
    BCB_BIT_AND = 4, BCB_BIT_OR = 5, BCB_BIT_XOR = 6,
    BCB_MINUS = 7, BCB_DIVIDES = 8 /* non-associative: bcb_accumulate (serial fold) only */
} bcb_op;

#define BCB_SUCCESS 0
#define BCB_EINVAL 10001       /* bad argument (null pointer, unknown dtype/op ...) */
#define BCB_EUNSUPPORTED 10002 /* combination outside the hot path (e.g. bit op on float) */
#define BCB_ETOOLARGE 10003    /* n beyond the supported range of the entry point */
#define BCB_ENODEVICE 10004    /* no CUDA device (system.hpp:238-241 no_device_found) */

typedef void *bcb_stream; /* cudaStream_t */

#if defined(__GNUC__)
#define BCB_API __attribute__((visibility("default")))
#else
#define BCB_API
#endif

BCB_API const char *bcb_error_string(int status);
BCB_API int bcb_version(void);

/* ---- device / queue / buffer plumbing: the subset of the L1 core the path touches ---- */
/* system::devices / device::name / compute_units / global_memory_size (system.hpp:92-196, device.hpp) */
BCB_API int bcb_device_count(int *count);
BCB_API int bcb_device_info(int device, char *name, size_t name_capacity, int *compute_units,
                    size_t *global_mem_bytes, int *cc_major, int *cc_minor);
BCB_API int bcb_set_device(int device);
BCB_API int bcb_get_device(int *device);
/* command_queue ctor / dtor / finish (command_queue.hpp:125-162, :1564-1572) */
BCB_API int bcb_stream_create(int device, bcb_stream *stream);
BCB_API int bcb_stream_destroy(bcb_stream stream);
BCB_API int bcb_stream_synchronize(bcb_stream stream);
/* buffer ctor / dtor (buffer.hpp:74-89); pinned host staging for copies */
BCB_API int bcb_malloc(void **device_ptr, size_t bytes);
BCB_API int bcb_free(void *device_ptr);
BCB_API int bcb_host_alloc(void **host_ptr, size_t bytes);
BCB_API int bcb_host_free(void *host_ptr);
/* mapped_view (container/mapped_view.hpp:217-240, CL_MEM_USE_HOST_PTR): register an existing host range so kernels
 * can address it in place (zero copy over PCIe); *device_ptr receives the device alias.  Unregister (which first waits
 * for the device, like bcb_free) before freeing the host memory. */
BCB_API int bcb_host_register(void *host_ptr, size_t bytes, void **device_ptr);
BCB_API int bcb_host_unregister(void *host_ptr);

/* enqueue_write_buffer / enqueue_read_buffer / enqueue_copy_buffer (command_queue.hpp:297-675); async on stream
 * (a host range of >= 32 MB in pageable memory is staged by the library, see bcb_sort_host: such a copy blocks) */
BCB_API int bcb_memcpy_h2d(bcb_stream stream, void *device_dst, const void *host_src, size_t bytes);
BCB_API int bcb_memcpy_d2h(bcb_stream stream, void *host_dst, const void *device_src, size_t bytes);
BCB_API int bcb_memcpy_d2d(bcb_stream stream, void *device_dst, const void *device_src, size_t bytes);
/* fill / iota / is_sorted (algorithm/fill.hpp, iota.hpp, is_sorted.hpp:39-68) -- the helpers either side of the path */
BCB_API int bcb_fill(bcb_stream stream, void *device_ptr, size_t n, const void *value_host, size_t value_bytes);
BCB_API int bcb_iota(bcb_stream stream, int dtype, void *device_ptr, size_t n, const void *start_host);
BCB_API int bcb_is_sorted(bcb_stream stream, int dtype, int descending, const void *keys, size_t n, int *result_host);
/* per-kernel device timing for benchmarks: when enabled, launchers bracket each kernel with CUDA events on the
 * stream.  bcb_timing_read waits for the stream, returns the summed duration and launch count of one kernel kind
 * since the last read of that kind, and forgets them. */
typedef enum bcb_kernel_kind {
    BCB_K_RADIX_HISTOGRAM = 0, BCB_K_DIGIT_SCAN = 1, BCB_K_ONESWEEP_PASS = 2, BCB_K_SCAN = 3, BCB_K_REDUCE = 4,
    BCB_K_OTHER = 5, BCB_K_EXCHANGE_PASS = 6, BCB_K_COUNT = 7
} bcb_kernel_kind;
BCB_API int bcb_timing_enable(bcb_stream stream, int enable);
BCB_API int bcb_timing_read(bcb_stream stream, int kind, double *total_ms, unsigned long long *launches);
/* scratch cache of a stream */
BCB_API int bcb_workspace_bytes(bcb_stream stream, size_t *bytes);
BCB_API int bcb_workspace_release(bcb_stream stream);

/* ---- sort ---- */
/* detail::radix_sort / radix_sort_by_key (algorithm/detail/radix_sort.hpp:428-462 -> radix_sort_impl :252-426).
 * Stable LSD radix sort by the reference's key transform (:100-127), in place.  values == NULL for keys only;
 * value_bytes is sizeof(T2), any size >= 1.  ascending != 0 sorts by less<T>, 0 by greater<T>. */
BCB_API int bcb_radix_sort(bcb_stream stream, int key_dtype, int ascending, void *keys, size_t n,
                   void *values, size_t value_bytes);
/* bcb_radix_sort as a sorted copy: reads keys_in (and values_in), leaves them untouched, writes the sorted range to
 * keys_out (values_out).  Same kernels and pass count (the first pass reads the source).  In place when the pointers
 * are equal. */
BCB_API int bcb_radix_sort_copy(bcb_stream stream, int key_dtype, int ascending, const void *keys_in, void *keys_out,
                                size_t n, const void *values_in, void *values_out, size_t value_bytes);

/* Keys-only sorts of >= 2^24 32/64-bit keys (except descending float / double keys, whose reference transform is not
 * injective) run a faster, speculatively stable pass kernel and verify the result on the device (sorted by the
 * transformed key <=> correct, because every pass is a permutation); the deterministic sort of the same buffer is
 * enqueued behind the check with every launch gated on its flag, so nothing waits on the host and the call stays
 * enqueue-and-return.  These counters report how often that happened on the stream (the read waits for the stream).
 * BCB_SORT_SPECULATIVE=0 in the environment disables speculation. */
BCB_API int bcb_sort_speculation_stats(bcb_stream stream, unsigned long long *verified_runs, unsigned long long *fallbacks);
/* is the range sorted by the transformed radix key of radix_sort.hpp:100-127 (the order radix_sort defines, which for
 * floats differs from operator< on +-0 / NaN)?  This is the check the speculative sort runs on its own output. Blocks. */
BCB_API int bcb_is_sorted_by_radix_key(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n,
                                       int *result_host);
/* detail::serial_insertion_sort / _by_key (algorithm/detail/insertion_sort.hpp:25-159): one thread, native compare.
 * greater != 0 uses ">" (descending).  n <= 4096. */
BCB_API int bcb_insertion_sort(bcb_stream stream, int key_dtype, int greater, void *keys, size_t n,
                       void *values, size_t value_bytes);
/* sort() on a host range (algorithm/sort.hpp:125-148: maps the range, sorts, unmaps): copies host_keys to the
 * device, applies the sort() dispatch of sort.hpp:34-81, copies back, and waits.  A range of >= 32 MB in PAGEABLE
 * memory (what sort(v.begin(), v.end()) on a std::vector hands over) is staged by the library itself -- up to 16 host
 * threads move 2 MB chunks through pinned slots on their own streams, memcpy and DMA overlapped -- instead of by the driver's
 * single-threaded path: 2^30 uint32 keys 200 ms against 690 ms (pinned memory: 171 ms).  BCB_STAGED_COPY=0 disables. */
BCB_API int bcb_sort_host(bcb_stream stream, int key_dtype, int descending, void *host_keys, size_t n);

/* sort() / stable_sort() / detail::merge_sort_on_gpu with a custom comparator (algorithm/sort.hpp:83-106,
 * stable_sort.hpp:34-50, detail/merge_sort_on_gpu.hpp:523-572).  The reference compiles an arbitrary compare(a, b) at run
 * time; ahead of time the family  f(a.field) < f(b.field)  ("> " with descending != 0) is provided, which covers every
 * comparator of the reference's own tests: records of record_bytes bytes each, a scalar field of field_dtype at
 * field_offset, f = BCB_UN_IDENTITY or BCB_UN_ABS (abs of a signed integer compares as unsigned, like OpenCL's abs).
 * Implemented as a STABLE key-value radix sort (project the field, sort the keys with the records as payload), in place.
 * Float fields order by the radix key (-0.0 before +0.0), as stable_sort(less<float>) does in the reference.
 * bcb_is_sorted_by_field is is_sorted(first, last, compare) for the same comparators (native compare; blocks). */
BCB_API int bcb_sort_by_field(bcb_stream stream, void *records, size_t n, size_t record_bytes, size_t field_offset,
                              int field_dtype, int unary, int descending);
BCB_API int bcb_is_sorted_by_field(bcb_stream stream, const void *records, size_t n, size_t record_bytes, size_t field_offset,
                                   int field_dtype, int unary, int descending, int *result_host);

/* Partition points of an already sorted range against splitters given in the transformed key space of
 * radix_sort.hpp:100-127 (uint64, non-decreasing): points_host[j] = first index whose transformed key is
 * >= splitters_host[j].  Used by the multi-GPU sample sort to cut a sorted shard into per-destination slices
 * (new functionality: the reference has no multi-device path).  Blocks. */
BCB_API int bcb_partition_points(bcb_stream stream, int key_dtype, int ascending, const void *sorted_keys, size_t n,
                                 const unsigned long long *splitters_host, size_t num_splitters,
                                 unsigned long long *points_host);

/* Stable partition of an UNSORTED range into num_splitters + 1 buckets (bucket of a key = number of splitters <= its
 * transformed key; splitters as for bcb_partition_points, at most 7): one onesweep-style pass, out of place
 * (keys_in -> keys_out, values likewise; value_bytes 0, 4 or 8).  counts_host receives the bucket sizes.  This is the
 * first step of the multi-GPU sample sort: the buckets are the per-destination slices of the all-to-all.  Blocks. */
BCB_API int bcb_partition_by_splitters(bcb_stream stream, int key_dtype, int ascending, const void *keys_in, void *keys_out,
                                       const void *values_in, void *values_out, size_t value_bytes, size_t n,
                                       const unsigned long long *splitters_host, size_t num_splitters,
                                       unsigned long long *counts_host);

/* 256-bin histogram of the MOST SIGNIFICANT 8-bit digit of the transformed keys (the order the sort is defined by), one
 * read of the range -- the first step of the multi-GPU sort when the key distribution lets whole digit values be dealt
 * to the ranks: the all-gathered histograms give the splitters AND the P x P count matrix in one exchange (no sampling,
 * no separate count pass).  Blocks. */
BCB_API int bcb_radix_top_histogram(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n,
                                    unsigned long long *counts_host /* [256] */);

/* The two halves of the partition as separate calls, for the multi-GPU sort that scatters straight into its peers'
 * receive buffers: bcb_partition_counts returns the bucket sizes (blocks); the host layer exchanges them, derives where
 * this rank's slice of every bucket starts inside each destination rank's buffer, and bcb_partition_scatter then writes
 * bucket b contiguously from dst_keys[b] (and dst_values[b]) -- plain device addresses, which may be another GPU's
 * memory opened with bcb_ipc_open, so the exchange is fused into the pass (NVLink stores, no all-to-all).
 * bcb_partition_scatter is asynchronous; value_bytes 0, 4 or 8 (else BCB_EUNSUPPORTED). */
BCB_API int bcb_partition_counts(bcb_stream stream, int key_dtype, int ascending, const void *keys, size_t n,
                                 const unsigned long long *splitters_host, size_t num_splitters,
                                 unsigned long long *counts_host);
BCB_API int bcb_partition_scatter(bcb_stream stream, int key_dtype, int ascending, const void *keys_in,
                                  const void *values_in, size_t value_bytes, size_t n,
                                  const unsigned long long *splitters_host, size_t num_splitters, void *const *dst_keys,
                                  void *const *dst_values);

/* The multi-GPU sort whose exchange is ONE of its radix passes (new functionality; the single-GPU counterpart is
 * radix_sort_impl, algorithm/detail/radix_sort.hpp:308-425): as many passes over the data as on one GPU, instead of a
 * partition pass plus a full local sort.
 *   1. bcb_radix_top_histogram on every rank, all-gathered by the host layer: whole values of the most significant digit
 *      are dealt to the ranks; inside the owner's receive buffer every digit value has a SEGMENT (starting on a 16-byte
 *      boundary), in which rank r's keys of that digit follow those of the ranks < r.
 *   2. bcb_radix_exchange_scatter (source side, asynchronous): one stable pass over the most significant digit that
 *      writes the run of digit d to dst_keys[d] (dst_values[d]) -- host arrays of 256 device addresses, 16-byte aligned,
 *      possibly another GPU's memory opened with bcb_ipc_open -- from element dst_first[d] on.  Every digit run of a
 *      tile leaves the SM as one bulk copy (cp.async.bulk), also over NVLink.  The input is not modified.  Keys with an
 *      invertible transform travel in their sortable form.
 *   3. after a stream-ordered barrier across the ranks, bcb_radix_sort_segments (destination side, asynchronous): every
 *      segment [seg_begin[i], seg_begin[i] + seg_len[i]) (elements; ascending, disjoint, 16-byte aligned starts) of the
 *      receive buffer is sorted on its own by the remaining digits -- stable LSD passes, ALL segments in one launch per
 *      digit -- and the last pass writes the segments back to back into out_keys / out_values.  recv_* are scratch.
 *      Keys-only sorts with an injective transform rank speculatively and are verified on out_keys (gated deterministic
 *      re-sort on the device, as in bcb_radix_sort).
 * Keys: 32-bit with value_bytes 0, 4 or 8, or 64-bit keys only with an injective transform; anything else returns
 * BCB_EUNSUPPORTED (decided from the types alone: every rank of a collective call gets the same answer). */
BCB_API int bcb_radix_exchange_scatter(bcb_stream stream, int key_dtype, int ascending, const void *keys, const void *values,
                                       size_t value_bytes, size_t n, void *const *dst_keys /* [256] */,
                                       void *const *dst_values /* [256] or NULL */, const unsigned long long *dst_first /* [256] */);
BCB_API int bcb_radix_sort_segments(bcb_stream stream, int key_dtype, int ascending, void *recv_keys, void *recv_values,
                                    size_t value_bytes, void *out_keys, void *out_values, const unsigned long long *seg_begin,
                                    const unsigned long long *seg_len, size_t num_segments);

/* Peer memory for the above: export a bcb_malloc'ed buffer as a 64-byte handle, open a peer's handle in this process
 * (cudaIpcGetMemHandle / cudaIpcOpenMemHandle with lazy peer access), close it again. */
BCB_API int bcb_ipc_export(void *device_ptr, unsigned char *handle64);
BCB_API int bcb_ipc_open(const unsigned char *handle64, void **device_ptr);
BCB_API int bcb_ipc_close(void *device_ptr);

/* ---- scan ---- */
/* detail::scan (algorithm/detail/scan.hpp:22-39) with the operator-generic semantics of serial_scan.hpp:26-97:
 * inclusive: out[i] = x0 op ... op xi;  exclusive: out[i] = init op x0 op ... op x(i-1).  Arithmetic in out_dtype
 * (exclusive_scan.hpp:80-85).  in == out (in place) is allowed.  init_host points at one out_dtype value
 * (NULL = 0); it is ignored for inclusive scans.  exclusive == 2 is an inclusive scan seeded with init
 * (out[i] = init op x0 op ... op xi), the per-rank step of the multi-GPU inclusive scan. */
BCB_API int bcb_scan(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive,
             const void *in, void *out, size_t n, const void *init_host);
/* The scan of ONE BLOCK of a block-distributed range (multi-GPU, new functionality): rank r scans its block seeded with
 * init op partial_0 op ... op partial_(r-1), the partials of the blocks before it, folded in rank order ON THE DEVICE
 * from records_dev -- `rank` records of 16 bytes as an all-gather leaves them: the partial (out_dtype) at byte 0, a
 * "block not empty" flag at byte 8 -- so the whole distributed scan (local reduce, all-gather, this call) is enqueued
 * without a host synchronisation.  exclusive: 0 or 1; init_host only counts for exclusive scans. */
BCB_API int bcb_scan_with_carry(bcb_stream stream, int in_dtype, int out_dtype, int op, int exclusive, const void *in,
                                void *out, size_t n, const void *init_host, const void *records_dev, int rank);

/* ---- reduce / accumulate ---- */
/* reduce (algorithm/reduce.hpp:275-305): result = x0 op ... op x(n-1) in result_dtype (= result_of<F(T,T)>,
 * the functor's type).  n == 0 leaves *result untouched (:283-285).  result_is_device selects a device
 * destination (enqueue-and-return) or a host destination (blocks until written). */
BCB_API int bcb_reduce(bcb_stream stream, int in_dtype, int result_dtype, int op, const void *in, size_t n,
               void *result, int result_is_device);
/* accumulate (algorithm/accumulate.hpp:102-188): returns init op x0 op ... as a host value of acc_dtype (the type
 * of init); op_dtype is the functor's argument type.  Associative (op, type) pairs run the parallel reduce
 * kernel and fold init in afterwards (exact for integers; float sums differ from the reference's serial fold
 * only by summation order); MINUS / DIVIDES and mixed acc/op types run the single-thread left fold of
 * detail/serial_accumulate.hpp:22-50.  Always blocks. */
BCB_API int bcb_accumulate(bcb_stream stream, int in_dtype, int op_dtype, int acc_dtype, int op, const void *in,
                   size_t n, const void *init_host, void *result_host);


/* ---- callers of scan and reduce with a closed, ahead-of-time compiled functor set ---- */
/* predicate ((x ARITH a) CMP b): what `_1 < 5`, `_1 * 2 >= 10`, `_1 % 2 == 1` of the reference's lambda placeholders
 * (lambda/placeholders.hpp) expand to; a_bits / b_bits hold a and b in the ELEMENT type's bit pattern (low bytes) */
typedef enum bcb_arith { BCB_AR_NONE = 0, BCB_AR_MUL = 1, BCB_AR_MOD = 2, BCB_AR_ADD = 3, BCB_AR_SUB = 4, BCB_AR_AND = 5 } bcb_arith;
typedef enum bcb_cmp { BCB_CMP_EQ = 0, BCB_CMP_NE = 1, BCB_CMP_LT = 2, BCB_CMP_LE = 3, BCB_CMP_GT = 4, BCB_CMP_GE = 5, BCB_CMP_TRUE = 6 } bcb_cmp;
typedef struct bcb_pred { int arith; int cmp; unsigned long long a_bits; unsigned long long b_bits; } bcb_pred;
/* unary functions (functional/: identity, negate; abs<T> = math builtin; square = _1 * _1) */
typedef enum bcb_unary { BCB_UN_IDENTITY = 0, BCB_UN_NEGATE = 1, BCB_UN_ABS = 2, BCB_UN_SQUARE = 3 } bcb_unary;

/* transform_if (algorithm/transform_if.hpp:42-85) / copy_if (copy_if.hpp: unary = identity): stable compaction of the
 * elements satisfying pred, transformed by `unary`, in ONE pass.  *count_host = number written (the reference returns
 * result + count, a host value: blocks).  in and out must not overlap. */
BCB_API int bcb_transform_if(bcb_stream stream, int dtype, const void *in, size_t n, int unary, const bcb_pred *pred,
                             void *out, size_t *count_host);
/* count_if / count (algorithm/detail/count_if_with_reduce.hpp:27-80; count(v) = pred {NONE, EQ, v}); blocks */
BCB_API int bcb_count_if(bcb_stream stream, int dtype, const void *in, size_t n, const bcb_pred *pred,
                         unsigned long long *count_host);
/* transform_reduce (algorithm/transform_reduce.hpp:40-90) and inner_product (inner_product.hpp:40-97): in2 == NULL ->
 * reduce_op over unary(transform, in1[i]); else reduce_op over binary(transform = a bcb_op code, in1[i], in2[i]).
 * Arithmetic in `dtype`; n == 0 leaves the result untouched; result on the host (blocks) or on the device. */
BCB_API int bcb_transform_reduce(bcb_stream stream, int dtype, const void *in1, const void *in2, size_t n, int transform,
                                 int reduce_op, void *result, int result_is_device);
/* reduce_by_key (algorithm/reduce_by_key.hpp:60-118, detail/reduce_by_key_with_scan.hpp:48-97): every run of consecutive
 * equal keys -> one (key, op-fold of its values) pair, in input order, in ONE pass.  Values: 4- and 8-byte types.
 * *count_host = number of runs (the reference returns the pair of end iterators: blocks). */
BCB_API int bcb_reduce_by_key(bcb_stream stream, int key_dtype, int val_dtype, const void *keys_in, const void *vals_in,
                              size_t n, void *keys_out, void *vals_out, int op, size_t *count_host);

/* set operations on two SORTED ranges (algorithm/set_union.hpp:120-199, set_intersection.hpp:104-175, set_difference.hpp:
 * 112-186, set_symmetric_difference.hpp:121-199; std::set_* multiset semantics, equal elements: first range first):
 * flags by binary search -> the library's scan -> scatter.  out must hold the result (na + nb elements always do) and
 * must not overlap a or b.  *count_host = number written (the reference returns result + count: blocks). */
typedef enum bcb_set_op { BCB_SET_UNION = 0, BCB_SET_INTERSECTION = 1, BCB_SET_DIFFERENCE = 2, BCB_SET_SYMMETRIC_DIFFERENCE = 3 } bcb_set_op;
BCB_API int bcb_set_operation(bcb_stream stream, int dtype, int which, const void *a, size_t na, const void *b, size_t nb,
                              void *out, size_t *count_host);
/* min_element / max_element (algorithm/min_element.hpp, max_element.hpp with less<T>;
 * detail/find_extrema_with_reduce.hpp:77-316): index of the FIRST smallest / largest element (ties: smaller index,
 * :156-158); n < 2 -> 0.  The reference returns an iterator, a host value: blocks. */
BCB_API int bcb_find_extremum(bcb_stream stream, int dtype, const void *in, size_t n, int want_max, size_t *index_host);

#ifdef __cplusplus
}
#endif
#endif /* COMPUTE_B200_H */
