/* glu_b200.h — C ABI of the B200-native replacement for loryruta/gl-radix-sort's hot path.
 *
 * The reference ("GLU") is a header-only C++17/OpenGL library; its hot path is three classes whose
 * operator() takes GL shader-storage-buffer handles:
 *     glu::Reduce(DataType, ReduceOperator)   void operator()(GLuint buffer, size_t count)
 *                                                              glu/Reduce.hpp:62,111
 *     glu::BlellochScan(DataType)             void operator()(GLuint buffer, size_t count, size_t num_partitions = 1)
 *                                                              glu/BlellochScan.hpp:91,130
 *     glu::RadixSort()                        void prepare_internal_buffers(size_t count)
 *                                             void operator()(GLuint key_buffer, GLuint val_buffer, size_t count,
 *                                                             size_t num_steps = 0)
 *                                                              glu/RadixSort.hpp:205,237,273
 * There is no FFI layer in the reference; this header is what one would bind.  The only change of
 * meaning is that a GLuint SSBO handle becomes a CUDA device pointer, and the scratch the reference
 * classes own (glu/RadixSort.hpp:193-200) becomes caller-provided temporary storage sized by the
 * matching *_tmp_bytes query.  include/glu/{Reduce,BlellochScan,RadixSort}.hpp re-create the classes
 * on top of these entry points (same names, arguments and print-and-exit error behaviour);
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns a glu_status (0 = success); nothing here prints or exits;
 *  - all hot-path calls only ENQUEUE work on `stream` (a cudaStream_t; NULL = default stream) of the
 *    current CUDA device and return without synchronising — like the reference's glDispatchCompute +
 *    glMemoryBarrier sequences (glu/Reduce.hpp:131-133);
 *  - results are produced IN PLACE in the caller's device buffers, as in the reference;
 *  - d_tmp must be 256-byte aligned (any cudaMalloc pointer) and at least *_tmp_bytes large; its
 *    contents are undefined before and after a call;
 *  - no CPU fallback exists: without a CUDA device every compute call returns GLU_ERROR_CUDA.
 */
#ifndef GLU_B200_H
#define GLU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GLU_API
#else
#define GLU_API __attribute__((visibility("default")))
#endif

typedef void* glu_stream_t; /* cudaStream_t */
typedef void* glu_event_t;  /* cudaEvent_t  */

typedef enum glu_status
{
    GLU_SUCCESS = 0,
    GLU_ERROR_INVALID_ARGUMENT = 1,  /* null buffer, count == 0 where the reference rejects it, ... */
    GLU_ERROR_INVALID_DATA_TYPE = 2, /* glu/data_types.hpp:40 "Invalid data type"                   */
    GLU_ERROR_INVALID_OPERATOR = 3,  /* glu/Reduce.hpp:93 "Invalid reduction operator"              */
    GLU_ERROR_TMP_TOO_SMALL = 4,
    GLU_ERROR_MISALIGNED = 5,        /* buffer not aligned to its scalar type / tmp not 256 B aligned */
    GLU_ERROR_COUNT_TOO_LARGE = 6,   /* see the per-function limits below                            */
    GLU_ERROR_CUDA = 7               /* a CUDA runtime call failed; see glu_last_cuda_error()        */
} glu_status;

/* glu/data_types.hpp:8-22 — same names, same values */
typedef enum glu_data_type
{
    GLU_DATA_TYPE_FLOAT = 0,
    GLU_DATA_TYPE_DOUBLE,
    GLU_DATA_TYPE_INT,
    GLU_DATA_TYPE_UINT,
    GLU_DATA_TYPE_VEC2,
    GLU_DATA_TYPE_VEC4,
    GLU_DATA_TYPE_DVEC2,
    GLU_DATA_TYPE_DVEC4,
    GLU_DATA_TYPE_UVEC2,
    GLU_DATA_TYPE_UVEC4,
    GLU_DATA_TYPE_IVEC2,
    GLU_DATA_TYPE_IVEC4
} glu_data_type;

/* glu/Reduce.hpp:42-48 */
typedef enum glu_reduce_operator
{
    GLU_REDUCE_OPERATOR_SUM = 0,
    GLU_REDUCE_OPERATOR_MUL,
    GLU_REDUCE_OPERATOR_MIN,
    GLU_REDUCE_OPERATOR_MAX
} glu_reduce_operator;

GLU_API int glu_version(void);
GLU_API const char* glu_status_string(int status);
/* cudaGetErrorString of the CUDA error behind the calling thread's last GLU_ERROR_CUDA */
GLU_API const char* glu_last_cuda_error(void);
/* size in bytes of one element of `data_type` (std430 array stride: 4/8/16/32), 0 if invalid */
GLU_API size_t glu_data_type_size(int data_type);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
GLU_API uint64_t glu_kernel_launch_count(void);

/* Per-kernel device timing for bench.py's roofline (the measure_gl_elapsed_time role, glu/gl_utils.hpp:249-265).
 * While enabled, every kernel launch of the hot path is bracketed by a pair of CUDA events on its own
 * stream.  glu_profile_collect() synchronises, sums the elapsed time and the launch count of one kernel
 * family since the last collect, and clears them.  Off by default (no events, no overhead). */
typedef enum glu_kernel_id
{
    GLU_KERNEL_REDUCE = 0,
    GLU_KERNEL_SCAN = 1,
    GLU_KERNEL_SORT_HISTOGRAM = 2,
    GLU_KERNEL_SORT_ONESWEEP = 3,
    GLU_KERNEL_SORT_PARTITION = 4, /* the multi-GPU partition / exchange pass */
    GLU_KERNEL_COUNT_ = 5
} glu_kernel_id;
GLU_API int glu_profile_enable(int on);
GLU_API int glu_profile_collect(int kernel_id, double* total_ms, uint64_t* launches);

/* ---------------------------------------------------------------------------------------------- hot path */

/* Replaces glu::Reduce::operator() (glu/Reduce.hpp:111-135).
 * Folds d_data[0..count) with `op` (component-wise for vector types) and leaves the result in
 * element 0.  Elements 1..count-1 are left untouched (the reference clobbers some of them with
 * partials; callers can rely on element 0 only, in both).  count == 1 is a no-op, count == 0 is
 * GLU_ERROR_INVALID_ARGUMENT (glu/Reduce.hpp:114).  Integer types wrap mod 2^32; floating types are
 * reduced in a fixed order (deterministic run to run). */
GLU_API size_t glu_reduce_tmp_bytes(size_t count, int data_type);
GLU_API int glu_reduce(void* d_data, size_t count, int data_type, int op, void* d_tmp, size_t tmp_bytes,
                       glu_stream_t stream);

/* Replaces glu::BlellochScan::operator() (glu/BlellochScan.hpp:130-139).
 * In-place exclusive prefix sum (operator +, identity 0) over each of `num_partitions` adjacent
 * segments of `count` elements.  Unlike the reference (glu/BlellochScan.hpp:134) `count` need not be
 * a power of two; on powers of two the results are identical.  count == 0 or num_partitions == 0 is
 * GLU_ERROR_INVALID_ARGUMENT.  count * num_partitions must be < 2^40 / element size. */
GLU_API size_t glu_scan_exclusive_tmp_bytes(size_t count, size_t num_partitions, int data_type);
GLU_API int glu_scan_exclusive(void* d_data, size_t count, size_t num_partitions, int data_type, void* d_tmp,
                               size_t tmp_bytes, glu_stream_t stream);

/* Replaces glu::RadixSort::operator() (glu/RadixSort.hpp:273-334).
 * Stable ascending sort of (key, value) uint32 pairs by key, in place in d_keys / d_vals.
 * num_steps keeps the reference's meaning (number of 4-bit steps: only the low 4*num_steps key bits
 * take part; 0 or >= 8 = all 32 bits).  Deviation: the result always lands in d_keys / d_vals (the
 * reference leaves an odd-num_steps result in its internal scratch).  count <= 1 is a no-op
 * (glu/RadixSort.hpp:278).  count must be < 2^31 (the reference's own limit: 32-bit uniforms). */
GLU_API size_t glu_radix_sort_u32kv_tmp_bytes(size_t count);
GLU_API int glu_radix_sort_u32kv(uint32_t* d_keys, uint32_t* d_vals, size_t count, size_t num_steps, void* d_tmp,
                                 size_t tmp_bytes, glu_stream_t stream);

/* The same stable LSD sort with the knobs the reference's README lists as limitations (README.md:88-89: "the value
 * buffer is mandatory") and CUB-style callers expect — SURVEY.md §8(f) row 3.  Same kernels, compile-time flavours:
 *  - d_vals == NULL: key-only sort (no value array is read, written or allocated: 36 B of HBM traffic per key
 *    instead of 68 B per pair);
 *  - only key bits [begin_bit, end_bit) take part (0 <= begin_bit <= end_bit <= 32; pairs whose keys agree in those
 *    bits keep their input order); ceil((end_bit - begin_bit) / 8) passes.  glu_radix_sort_u32kv's num_steps is
 *    begin_bit = 0, end_bit = 4 * num_steps;  begin_bit == end_bit is a no-op;
 *  - descending != 0: largest key first, equal keys still in input order (std::stable_sort with greater<>) — the
 *    passes partition by the complemented digit.
 * d_tmp is sized by glu_radix_sort_u32_ex_tmp_bytes(count, with_values) (with_values = 0 for d_vals == NULL). */
GLU_API size_t glu_radix_sort_u32_ex_tmp_bytes(size_t count, int with_values);
GLU_API int glu_radix_sort_u32_ex(uint32_t* d_keys, uint32_t* d_vals, size_t count, unsigned begin_bit, unsigned end_bit,
                                  int descending, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);

/* 64-bit keys and payloads wider than 32 bits, stable, ascending or descending, in place:
 * key_bytes is 4 or 8 (unsigned integers), value_bytes is 0 (d_vals == NULL: keys only), 4, 8 or 16 (opaque
 * elements, e.g. uint64 / double / a vec4).  Built on the 32-bit sort: (key word, element index) pairs are sorted —
 * twice for 8-byte keys, low word then high word — and the wide keys and values are moved once through the
 * resulting permutation (gl-radix-sort_b200/csrc/glu_radix_sort_wide.cu).  key_bytes == 4 with value_bytes <= 4 is
 * glu_radix_sort_u32_ex itself.  d_keys / d_vals must be aligned to their element size; count < 2^31. */
GLU_API size_t glu_radix_sort_wide_tmp_bytes(size_t count, size_t key_bytes, size_t value_bytes);
GLU_API int glu_radix_sort_wide(void* d_keys, size_t key_bytes, void* d_vals, size_t value_bytes, size_t count,
                                int descending, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);

/* ------------------------------------------------------------------ building blocks of the multi-GPU path
 * Not in the reference (it is single-GPU); these are what gl-radix-sort_b200/distributed.py composes with
 * NCCL / NVLink peer memory (DESIGN.md "Multi-GPU").  Same conventions as above. */

/* glu_reduce with the result written to d_result[0] (one element) instead of d_data[0]; d_data is not modified.
 * count == 1 copies the element. */
GLU_API int glu_reduce_into(const void* d_data, size_t count, int data_type, int op, void* d_result, void* d_tmp,
                            size_t tmp_bytes, glu_stream_t stream);

/* glu_scan_exclusive starting every partition from *d_init (one element in device memory, read when the kernel
 * runs: the std::exclusive_scan `init`) instead of 0; d_init == NULL means 0.  A rank's base in a sharded scan. */
GLU_API int glu_scan_exclusive_init(void* d_data, size_t count, size_t num_partitions, int data_type,
                                    const void* d_init, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);

/* d_hist[0 .. 256) <- number of keys whose digit ((key >> shift) & ((1 << bits) - 1)) has each value
 * (bins >= 1 << bits stay 0); 1 <= bits <= 8.  Reads the keys once. */
GLU_API int glu_radix_histogram_u32(const uint32_t* d_keys, size_t count, unsigned shift, unsigned bits,
                                    uint32_t* d_hist, glu_stream_t stream);

/* Stable partition of the pairs by one digit — one onesweep pass whose output is a table of destinations:
 * the i-th pair (in input order) whose digit is d goes to d_key_dst[d][i] / d_val_dst[d][i].  d_key_dst and
 * d_val_dst are DEVICE arrays of 1 << bits pointers; the pointers may address peer GPUs' memory (NVLink P2P),
 * which turns the pass into a fused partition + all-to-all.  Inputs are not modified. */
GLU_API size_t glu_radix_partition_u32kv_tmp_bytes(size_t count);
GLU_API int glu_radix_partition_u32kv(const uint32_t* d_keys, const uint32_t* d_vals, size_t count, unsigned shift,
                                      unsigned bits, uint32_t* const* d_key_dst, uint32_t* const* d_val_dst,
                                      void* d_tmp, size_t tmp_bytes, glu_stream_t stream);

/* The same pass partitioning by DESTINATION instead of by digit: the i-th pair whose digit maps to destination
 * g = d_dest_of_digit[digit] (a device array of 256 bytes, every g < 16) goes to d_key_dst[g][i] / d_val_dst[g][i].
 * d_key_dst / d_val_dst are device arrays of at least 16 pointers: the first 16 entries are read, only the first
 * max(g)+1 are dereferenced.  d_dest_of_digit need not be monotone (padding of the last tile never depends on it).  A tile then
 * leaves the SM as a few long runs (tile / #destinations pairs each) instead of 256 short ones, which is what
 * remote stores over NVLink need to run near link speed: the fused partition + all-to-all of the multi-GPU sort. */
GLU_API int glu_radix_partition_by_dest_u32kv(const uint32_t* d_keys, const uint32_t* d_vals, size_t count,
                                              unsigned shift, unsigned bits, const uint8_t* d_dest_of_digit,
                                              uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp,
                                              size_t tmp_bytes, glu_stream_t stream);

/* Device-resident counts: the same kernels reading the number of pairs from *d_count (a uint32 in device memory written
 * by an earlier operation of the stream; *d_count <= max_count) — what lets the multi-GPU sort run without a host
 * round trip between its exchange plan and the passes that depend on it.  Grids and d_tmp are sized for max_count
 * (glu_radix_sort_u32kv_tmp_bytes(max_count) / glu_radix_partition_u32kv_tmp_bytes(max_count)); after the sort
 * elements [*d_count, max_count) of d_keys / d_vals are undefined. */
GLU_API int glu_radix_sort_u32kv_dyn(uint32_t* d_keys, uint32_t* d_vals, const uint32_t* d_count, size_t max_count,
                                     size_t num_steps, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);
GLU_API int glu_radix_partition_u32kv_dyn(const uint32_t* d_keys, const uint32_t* d_vals, const uint32_t* d_count,
                                          size_t max_count, unsigned shift, unsigned bits, uint32_t* const* d_key_dst,
                                          uint32_t* const* d_val_dst, void* d_tmp, size_t tmp_bytes, glu_stream_t stream);
GLU_API int glu_radix_partition_by_dest_u32kv_dyn(const uint32_t* d_keys, const uint32_t* d_vals,
                                                  const uint32_t* d_count, size_t max_count, unsigned shift,
                                                  unsigned bits, const uint8_t* d_dest_of_digit,
                                                  uint32_t* const* d_key_dst, uint32_t* const* d_val_dst, void* d_tmp,
                                                  size_t tmp_bytes, glu_stream_t stream);

/* SEGMENTED stable sort: `num_segments` (<= 256) independent (key, value) sequences sorted by key bits
 * [begin_bit, end_bit) in one histogram launch + ceil(bits / 8) digit passes — what the multi-GPU sort runs on the
 * top-digit buckets a rank received ("the local onesweep on the remaining 24 bits", BASELINE.json north_star): 4 + 16 B
 * per pair and pass instead of a full 32-bit sort of the received range.
 *   Input (arrays A): segment s holds d_seg_count[s] pairs (device-resident counts) starting at element
 *   first_tile[s] * glu_radix_sort_segment_tile(), first_tile = exclusive scan of ceil(count / tile): segments start at
 *   tile boundaries, the gap behind a segment's last pair is padding (never read as data, never written).
 *   Output: compact — segment s at the sum of the counts before it — in arrays B when the number of passes
 *   (ceil((end_bit - begin_bit) / 8)) is odd, in arrays A when it is even; *result_in_b says which (host value, may be
 *   NULL).  The other pair of arrays is scratch.  All four arrays hold max_tiles * tile elements and are 16-byte
 *   aligned.  If the segments need more than max_tiles tiles nothing is sorted (the caller's plan checks capacity).
 * Keys with equal bits keep their order inside the segment (stable); segments never mix. */
GLU_API size_t glu_radix_sort_segment_tile(void);
GLU_API size_t glu_radix_sort_u32kv_segmented_tmp_bytes(size_t max_tiles);
GLU_API int glu_radix_sort_u32kv_segmented(uint32_t* d_keys_a, uint32_t* d_vals_a, uint32_t* d_keys_b, uint32_t* d_vals_b,
                                           const uint32_t* d_seg_count, size_t num_segments, size_t max_tiles,
                                           unsigned begin_bit, unsigned end_bit, void* d_tmp, size_t tmp_bytes,
                                           glu_stream_t stream, int* result_in_b);

/* The same sort on an input made of RUNS (the multi-GPU sort whose all-to-all is one DMA copy per peer: a rank receives
 * one chunk per source, and a bucket is spread over the chunks).  The input of the first pass is `num_runs` runs in
 * segment order — a segment is one or more consecutive runs, concatenated in run order (this is what keeps the global
 * sort stable: the runs of a bucket in source-rank order) —, run r = count[r] pairs starting at TILE phys[r] of the A
 * arrays (runs start at tile boundaries, anywhere in the arrays; the slots behind a run's last pair are padding).
 *   d_runs: 5 rows of (num_runs + 1) uint32 — [0] first tile of the run in the first pass's own tile numbering =
 *   exclusive scan of ceil(count / tile), entry num_runs = the total; [1] phys; [2] count; [3] segment of the run;
 *   [4] first tile (numbering of row 0) of the run's segment.  d_seg_count[s] = sum of the counts of segment s's runs.
 * Later passes and the output are those of glu_radix_sort_u32kv_segmented (tile-aligned segments, compact result);
 * max_tiles must also cover the first pass's tiles (d_runs[0][num_runs]). */
GLU_API int glu_radix_sort_u32kv_segmented_runs(uint32_t* d_keys_a, uint32_t* d_vals_a, uint32_t* d_keys_b,
                                                uint32_t* d_vals_b, const uint32_t* d_seg_count, size_t num_segments,
                                                size_t max_tiles, unsigned begin_bit, unsigned end_bit,
                                                const uint32_t* d_runs, size_t num_runs, void* d_tmp, size_t tmp_bytes,
                                                glu_stream_t stream, int* result_in_b);

/* The exchange plan of the multi-GPU sort, computed on the device from the all-gathered split-digit histograms
 * d_hist_all[world][256] (world <= 16): bucket -> destination rank by balanced prefix (contiguous bucket ranges),
 * and this rank's destination table for glu_radix_partition_by_dest_u32kv_dyn —
 *   d_dest_of_digit[256]            destination rank of every bucket,
 *   d_key_dst / d_val_dst[256]      entry g < world: d_peer_keys[g] / d_peer_vals[g] (base address of rank g's receive
 *                                   arrays as mapped in THIS process) advanced past what ranks < `rank` send to g,
 *   d_counts[0]                     pairs this rank sends (= send_count), d_counts[1] pairs this rank receives,
 *   d_info[world + 2]               pairs every rank receives, then d_counts[1], then an overflow flag.
 * If any rank would receive more than `capacity` pairs the flag is set, both counts are 0 and every destination is
 * NULL, so the dependent passes do nothing; the host reads d_info and reports.  One tiny kernel, no host sync. */
GLU_API int glu_radix_exchange_plan(const uint32_t* d_hist_all, int world, int rank, size_t send_count, size_t capacity,
                                    const uint64_t* d_peer_keys, const uint64_t* d_peer_vals, uint32_t** d_key_dst,
                                    uint32_t** d_val_dst, uint8_t* d_dest_of_digit, uint32_t* d_counts,
                                    uint64_t* d_info, glu_stream_t stream);

/* The BUCKET-major exchange plan (MSD split followed by the segmented local sort): same bucket -> rank assignment, but
 * rank g receives its buckets one after the other in increasing order, each starting at a tile boundary of
 * glu_radix_sort_u32kv_segmented (glu_radix_sort_segment_tile() pairs), inside a bucket the sources in rank order —
 *   d_key_dst / d_val_dst[256]   entry b: where THIS rank's run of bucket b goes (for glu_radix_partition_u32kv_dyn),
 *   d_seg_count[256]             pairs of bucket b if this rank receives it, else 0: the segment table of the local sort,
 *   d_counts / d_info            as glu_radix_exchange_plan (d_counts[1], d_info[g] = pairs received, without padding).
 * capacity_tiles: tiles a rank's receive arrays hold; if any rank needs more the overflow flag is set, nothing is sent
 * and nothing is sorted. */
GLU_API int glu_radix_exchange_plan_buckets(const uint32_t* d_hist_all, int world, int rank, size_t send_count,
                                            size_t capacity_tiles, const uint64_t* d_peer_keys,
                                            const uint64_t* d_peer_vals, uint32_t** d_key_dst, uint32_t** d_val_dst,
                                            uint32_t* d_seg_count, uint32_t* d_counts, uint64_t* d_info,
                                            glu_stream_t stream);

/* The STAGED form of the bucket-major exchange: (1) a local MSD pass (glu_radix_partition_u32kv_dyn with the table
 * written by glu_radix_exchange_stage_tables) brings this rank's pairs into bucket-major order in its own staging arrays
 * at full HBM speed; (2) glu_radix_exchange_copy_u32kv moves every bucket's run — count[b] pairs at d_stage_off[b] — to
 * d_key_dst[b] / d_val_dst[b] (the tables of glu_radix_exchange_plan_buckets; peer memory over NVLink).  The runs are
 * long, so every warp-wide store is one full aligned 128-byte line, and a grid of one CTA per SM (num_ctas <= 0) keeps
 * the links busy while the rest of the GPU sorts.  d_my_hist: this rank's 256 split-digit counts (its row of the
 * all-gathered histograms).  Buckets the rank keeps for itself (d_seg_count[b] != 0, the plan's segment table; may be
 * NULL) skip the staging arrays: their table entry is the final place d_key_dst[b] / d_val_dst[b] and d_copy_count[b]
 * is 0; for the others d_copy_count[b] = d_my_hist[b] — the `d_count` of glu_radix_exchange_copy_u32kv.  Buckets whose
 * destination pointer is NULL (overflow) are skipped. */
GLU_API int glu_radix_exchange_stage_tables(const uint32_t* d_my_hist, uint32_t* d_stage_keys, uint32_t* d_stage_vals,
                                            const uint32_t* d_seg_count, uint32_t* const* d_key_dst,
                                            uint32_t* const* d_val_dst, uint32_t** d_stage_key_dst,
                                            uint32_t** d_stage_val_dst, uint32_t* d_stage_off, uint32_t* d_copy_count,
                                            glu_stream_t stream);
GLU_API int glu_radix_exchange_copy_u32kv(const uint32_t* d_stage_keys, const uint32_t* d_stage_vals,
                                          const uint32_t* d_stage_off, const uint32_t* d_count, uint32_t* const* d_key_dst,
                                          uint32_t* const* d_val_dst, int num_ctas, glu_stream_t stream);

/* CUDA IPC plumbing for one-process-per-GPU peer access: export a glu_malloc'ed allocation, map a peer's. */
#define GLU_IPC_HANDLE_BYTES 64
GLU_API int glu_ipc_get_handle(void* d_ptr, unsigned char handle[GLU_IPC_HANDLE_BYTES]);
GLU_API int glu_ipc_open_handle(const unsigned char handle[GLU_IPC_HANDLE_BYTES], void** d_ptr);
GLU_API int glu_ipc_close_handle(void* d_ptr);

/* ------------------------------------------------------------------------ host-buffer entry points (e2e)
 * Same semantics on HOST arrays: upload -> hot path -> download, synchronous.  These are what the
 * reference's test-suite does around every call with ShaderStorageBuffer(data) ... get_data<T>()
 * (glu/gl_utils.hpp:157-171,226-235; test/radix_sort_tests.cpp:100-106). */
GLU_API int glu_reduce_host(void* h_data, size_t count, int data_type, int op);
GLU_API int glu_scan_exclusive_host(void* h_data, size_t count, size_t num_partitions, int data_type);
GLU_API int glu_radix_sort_u32kv_host(uint32_t* h_keys, uint32_t* h_vals, size_t count, size_t num_steps);

/* A queue of host-buffer sorts for callers that sort one batch after another (the loop of
 * test/radix_sort_tests.cpp:160-193 with real data): up to `depth` jobs are in flight, each on its own stream with
 * its own device arrays and scratch, so the upload of job k+1 overlaps the sort and the download of job k (PCIe is
 * full duplex; a single synchronous call leaves each direction idle half of the time).  submit() enqueues
 * upload -> glu_radix_sort_u32kv -> download and returns; it blocks only while the slot it needs (the job submitted
 * `depth` submits ago) is still running.  The host arrays must stay valid until wait() and should be pinned
 * (glu_malloc_host).  Results land in place in the host arrays, exactly as with glu_radix_sort_u32kv_host. */
typedef struct glu_host_sort_queue glu_host_sort_queue_t;
GLU_API int glu_host_sort_queue_create(glu_host_sort_queue_t** queue, size_t max_count, int depth);
GLU_API int glu_host_sort_queue_submit(glu_host_sort_queue_t* queue, uint32_t* h_keys, uint32_t* h_vals, size_t count,
                                       size_t num_steps);
GLU_API int glu_host_sort_queue_wait(glu_host_sort_queue_t* queue); /* every submitted job is complete */
GLU_API int glu_host_sort_queue_destroy(glu_host_sort_queue_t* queue);

/* ------------------------------------------------------- device plumbing for the C++ classes and test runner
 * (the ShaderStorageBuffer / measure_gl_elapsed_time roles, glu/gl_utils.hpp:146-265) */
GLU_API int glu_device_count(int* count);
GLU_API int glu_set_device(int device);
GLU_API int glu_device_info(int device, char* name, size_t name_cap, int* sm_count, int* cc_major, int* cc_minor,
                            size_t* total_mem_bytes, int* warp_size);
GLU_API int glu_malloc(void** d_ptr, size_t bytes);
GLU_API int glu_free(void* d_ptr);
GLU_API int glu_malloc_host(void** h_ptr, size_t bytes); /* pinned */
GLU_API int glu_free_host(void* h_ptr);
GLU_API int glu_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, glu_stream_t stream);
GLU_API int glu_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, glu_stream_t stream);
GLU_API int glu_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, glu_stream_t stream);
GLU_API int glu_memset_u32(void* d_dst, uint32_t value, size_t count, glu_stream_t stream);
/* Stream-ordered signalling between GPUs (the device-side barrier of the multi-GPU sort's "dma" exchange, no
 * collective involved): glu_signal_peers_u32 stores `value` (an epoch number) into the words h_flag_addrs[0..count)
 * (count <= 16, addresses in peer memory as mapped in this process, 0 = skip) after everything enqueued on `stream`
 * before it — e.g. the copies into those peers; glu_stream_wait_flags_u32 makes `stream` wait until the local words
 * d_flags[i] (i < count <= 32, i != skip) have all reached `value` (wrap-around safe; traps after ~4 s instead of
 * hanging if a peer never signals). */
GLU_API int glu_signal_peers_u32(const uint64_t* h_flag_addrs, int count, uint32_t value, glu_stream_t stream);
GLU_API int glu_stream_wait_flags_u32(const uint32_t* d_flags, int count, int skip, uint32_t value, glu_stream_t stream);
GLU_API int glu_stream_create(glu_stream_t* stream); /* a blocking stream: ordered with the legacy default stream */
GLU_API int glu_stream_destroy(glu_stream_t stream);
GLU_API int glu_stream_synchronize(glu_stream_t stream);
GLU_API int glu_event_create(glu_event_t* event);
GLU_API int glu_event_destroy(glu_event_t event);
GLU_API int glu_event_record(glu_event_t event, glu_stream_t stream);
GLU_API int glu_event_synchronize(glu_event_t event);
GLU_API int glu_event_elapsed_ms(float* ms, glu_event_t start, glu_event_t stop);

#ifdef __cplusplus
}
#endif

#endif /* GLU_B200_H */
