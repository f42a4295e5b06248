/*
 * gcnb.h -- C ABI of libgcn_b200.so, the Blackwell (sm_100a) GCN training kernels.
 *
 * This is the drop-in boundary of the hot path: plain pointers and sizes, no C++/torch types.  The
 * reference (davide-gurrieri/parallel-GCN) has no FFI of its own; each entry point below replaces the CUDA
 * kernel(s) behind one method of the reference's C++ module API and cites it (paths relative to the
 * reference repo).  The C++ classes in parallel-gcn_b200/host/ (same names/signatures as the reference:
 * Variable, Dropout, SparseMatmul, GraphSum, ReLU, Matmul, CrossEntropyLoss, Adam, GCN, Parser) call ONLY
 * these functions; tests and bench.py call them through ctypes.
 *
 * Conventions
 *   - every `d_*` pointer is a device pointer; tensors are fp32 row-major, indices uint32 (the reference's
 *     `natural`), labels/truth int32, sizes int64;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - return value: 0 on success, otherwise a cudaError_t (or GCNB_E_* below); gcnb_error_string() explains;
 *   - no hidden global state: per-graph metadata lives in an explicit plan object; all reductions are
 *     fixed-order (no floating-point atomics), so results are bit-reproducible run to run;
 *   - there is NO CPU fallback: every function fails (cudaErrorNoDevice / cudaErrorInsufficientDriver ...)
 *     when no sm_100 device is usable.
 */
#ifndef GCNB_H
#define GCNB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCNB_API __attribute__((visibility("default")))
#define GCNB_E_BADARG 10001
#define GCNB_E_UNSUPPORTED 10002
#define GCNB_E_COMM 10003 /* NCCL missing or a collective failed */

typedef void *gcnb_stream_t; /* cudaStream_t */

GCNB_API const char *gcnb_error_string(int code);

GCNB_API int gcnb_version(void);
/* host threads the plan builders (window staging, bit tiles, ELL) may use; 0 = default = min(16, hardware threads) or
 * GCNB_HOST_THREADS.  The row-partitioned engine sets hardware threads / ranks: the ranks of a job share one host. */
GCNB_API int gcnb_set_host_threads(int n);
GCNB_API int gcnb_host_threads(void);
/* device sanity: fails unless the current device is compute capability 10.x; fills SM count. */
GCNB_API int gcnb_device_check(int *sm_count);

/* ---------------------------------------------------------------------------------------------------
 * CSR x dense plan.  One plan per sparse index (graph adjacency, feature matrix, or the transposed
 * feature matrix); built once, reused by every launch.  Holds the load-balancing metadata the reference
 * does not have: rows are cut into segments of <= seg_nnz entries, segments are grouped into per-SM queues of
 * equal nnz (contiguous rows per SM => neighbour rows are re-used from that SM's L1), long rows are combined
 * in a fixed order from per-segment partials.  Replaces the launch-shape knobs CudaParams::N_BLOCKS/N_THREADS
 * (include/utils.cuh:17-23, src/module.cu:71-72 pattern).
 * The plan BORROWS d_indptr/d_indices (they must outlive it), exactly like DevSparseIndex users do
 * (include/sparse.cuh:21-29).
 * ------------------------------------------------------------------------------------------------- */
typedef struct gcnb_spmm_plan gcnb_spmm_plan;
GCNB_API int gcnb_spmm_plan_create(const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows,
                                   int64_t n_cols, int seg_nnz /*0 = default*/, gcnb_stream_t stream,
                                   gcnb_spmm_plan **out);
GCNB_API int gcnb_spmm_plan_destroy(gcnb_spmm_plan *plan);
/* d_out[i] = value of the first entry (i, i + col0) of CSR row i, 0 when the row has none: the diagonal of a row block of
 * the normalised adjacency (src/parser.cpp:164-181: 1 / deg_i), whose square roots are the scales of a bit-tile plan */
GCNB_API int gcnb_csr_diagonal_f32(const uint32_t *d_indptr, const uint32_t *d_indices, const float *d_values,
                                   int64_t n_rows, int64_t col0, float *d_out, gcnb_stream_t stream);
/* number of segments / split rows / queues (introspection for tests & DESIGN numbers) */
GCNB_API int gcnb_spmm_plan_info(const gcnb_spmm_plan *plan, int64_t out[8]);

/* Optional window-staged fast path for a product whose values never change between launches -- GraphSum's
 * graph_value (src/parser.cpp:164-181, GraphSum::forward/backward src/module.cu:188-210).  Builds (once, on the host,
 * multi-threaded) a second representation of the index: entries of rows with many neighbours inside one column window
 * of B (a window = as many rows of B as fit in an SM's shared memory at feature width `dim`) become packed segments
 * with 16-bit window-local column ids and a copy of their values; the rest stays a (smaller) CSR.  Later
 * gcnb_spmm_f32 calls with the SAME d_values pointer, this dim and no permutation run the staged kernel + the generic
 * kernel on the remainder; any other call is unaffected.  h_indptr / h_indices: host copies of the plan's CSR arrays,
 * or NULL to have them copied back from the device.  The representation is built for dim == 16 (other dims: no-op,
 * returns 0); products of width 16 and of any width >= 64 then use it (wide operands run as 16-column slabs, see
 * gcnb_spmm_ld_f32).  Graphs without column locality (< 25 % of the entries stageable) are left on the generic kernel
 * (gcnb_reorder_communities, include/gcnb_engine.h, renumbers a graph so that its communities become contiguous).
 * Calling it again with a new value array re-gathers the packed values only. */
GCNB_API int gcnb_spmm_plan_stage(gcnb_spmm_plan *plan, const uint32_t *h_indptr, const uint32_t *h_indices,
                                  const float *d_values, int dim, gcnb_stream_t stream);
/* same with the builder's knobs exposed (0 = default): window_rows <= 3072, min_seg = fewest entries of a row inside a
 * window worth a segment (16), seg_cap = longest segment (512), min_window_nnz = fewest staged entries that justify
 * copying a window (8 * window_rows) */
GCNB_API int gcnb_spmm_plan_stage_ex(gcnb_spmm_plan *plan, const uint32_t *h_indptr, const uint32_t *h_indices,
                                     const float *d_values, int dim, int window_rows, int min_seg, int seg_cap,
                                     int64_t min_window_nnz, gcnb_stream_t stream);
/* Background staging: the same build and upload on a helper thread with its own stream (host copies of the index are read
 * back from the device there), while the plan keeps serving products on the generic kernel.  _finish joins the helper
 * and attaches the result; call it at a point of the caller's choosing (results before / after differ in summation
 * order only), from the thread that uses the plan, while no product on this plan is being enqueued.  *job = NULL: nothing
 * to do (other dim, empty matrix, already staged).  _done polls without blocking. */
typedef struct gcnb_stage_job gcnb_stage_job;
GCNB_API int gcnb_spmm_plan_stage_async_begin(gcnb_spmm_plan *plan, const float *d_values, int dim, gcnb_stage_job **job);
GCNB_API int gcnb_spmm_plan_stage_async_done(const gcnb_stage_job *job);
GCNB_API int gcnb_spmm_plan_stage_async_finish(gcnb_spmm_plan *plan, gcnb_stage_job *job);
/* out = {staged?, window_rows, staged entries, remainder entries, segments, runs, chunks, partial slots} */
GCNB_API int gcnb_spmm_plan_stage_info(const gcnb_spmm_plan *plan, int64_t out[8]);
/* number of 16-column slabs a product of width `dim` runs as through the staged kernels (0: generic kernel).  Widths
 * 16 and >= 64 are staged; each slab launches the staged kernel, the remainder kernel and the merge kernel (+ one
 * packing kernel when B is not 16 floats wide). */
GCNB_API int gcnb_spmm_plan_stage_slabs(const gcnb_spmm_plan *plan, int dim);
/* Row-partitioned GraphSum (one rank = one row block of A, B = the slabs of all ranks gathered over NVLink): columns
 * [col0, col1) of A -- rows of B -- are the rank's OWN slab, which exists before the exchange.  Call before
 * gcnb_spmm_plan_stage: staged column windows that lie entirely inside the range get their own run list.  Then, per
 * product, gcnb_spmm_stage_own_f32 launches those runs from the rank's slab (d_B_own = its first row, 16 floats per row)
 * while the exchange is still in flight, and the following gcnb_spmm_f32 call on the gathered matrix does the rest
 * (other windows, remainder entries, merge).  *launched = 0 means nothing was launched (no staged plan, no window inside
 * the range, other dim / value array): the gcnb_spmm_f32 call then does everything, as without this call. */
GCNB_API int gcnb_spmm_plan_set_own_cols(gcnb_spmm_plan *plan, int64_t col0, int64_t col1);
GCNB_API int gcnb_spmm_stage_own_f32(gcnb_spmm_plan *plan, const float *d_values, const float *d_B_own, int dim,
                                     gcnb_stream_t stream, int *launched);

/* The staging builder on its own, host memory only, no CUDA call: lets the plan layout be verified on a machine
 * without a GPU (tests/test_stage_cpu.py).  Arrays are described in parallel-gcn_b200/csrc/spmm_plan.cuh. */
typedef struct gcnb_stage_host gcnb_stage_host;
GCNB_API int gcnb_stage_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols,
                                   int dim, int window_rows /*0 = default*/, int min_seg /*0 = default*/,
                                   int seg_cap /*0 = default*/, int64_t min_window_nnz /*0 = default*/,
                                   int n_cta /*0 = 148*/, int n_threads /*0 = auto*/, gcnb_stage_host **out);
/* same, with the own column range of a row-partitioned product (gcnb_spmm_plan_set_own_cols) */
GCNB_API int gcnb_stage_host_build_own(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols,
                                       int dim, int window_rows, int min_seg, int seg_cap, int64_t min_window_nnz, int n_cta,
                                       int n_threads, int64_t own_col0, int64_t own_col1, gcnb_stage_host **out);
GCNB_API int gcnb_stage_host_sizes(const gcnb_stage_host *h, int64_t out[13]);
GCNB_API int gcnb_stage_host_copy(const gcnb_stage_host *h, int which, void *dst, int64_t bytes);
GCNB_API int gcnb_stage_host_destroy(gcnb_stage_host *h);

/* Bit-tile GraphSum (parallel-gcn_b200/csrc/spmm_bittile.cu): the dense blocks of a product whose values factor as
 * value[i,j] = row_scale[i] * col_scale[j] -- GraphSum's graph_value = 1/sqrt(deg_i deg_j), src/parser.cpp:164-181 --
 * run on the tcgen05 tensor cores.  Rows are cut into blocks of 128, columns into chunks of 64; a (block, chunk) tile
 * with >= min_tile_nnz (0 = 1.6 % of the cells) entries becomes a 128 x 64 bit map (chunk_cols = 128: 128 x 128), B is pre-scaled and split into three exact bf16
 * pieces per launch, the 0/1 x bf16 products are exact and accumulate in fp32 (TMEM); every other entry (sparse tiles,
 * duplicates, values that are not the product of the scales within 1e-6 relative) stays in a remainder CSR with its
 * original value and goes through the generic kernel on a second stream.  h_* are HOST arrays (the plan copies what it
 * needs).  h_row_scale / h_col_scale: both NULL = derive them from the diagonal entries (sqrt(value[i,i]); square
 * matrices only, rows without a diagonal entry stay in the remainder).  Width is fixed at 16 columns, B and C
 * contiguous (row stride 16).  Result: fixed summation order (bit-reproducible), within ~1e-6 relative of the CSR
 * product.  Needs compute capability 10.x (GCNB_E_UNSUPPORTED otherwise). */
typedef struct gcnb_bittile_plan gcnb_bittile_plan;
GCNB_API int gcnb_bittile_plan_create(const uint32_t *h_indptr, const uint32_t *h_indices, const float *h_values,
                                      int64_t n_rows, int64_t n_cols, const float *h_row_scale, const float *h_col_scale,
                                      int min_tile_nnz, int chunk_cols /*0 = 64; 64 or 128*/,
                                      int row_blocks /*0 = 1; 2 = items of 256 rows (chunk 64 only)*/, gcnb_stream_t stream,
                                      gcnb_bittile_plan **out);
GCNB_API int gcnb_bittile_plan_destroy(gcnb_bittile_plan *plan);
/* The same plan built ON THE DEVICE from device-resident arrays (parallel-gcn_b200/csrc/spmm_bittile_build.cu): the CSR never
 * travels to the host; bit-identical to gcnb_bittile_plan_create on the same matrix.  GCNB_E_UNSUPPORTED when the matrix needs
 * the host builder: entries that do not factor (they keep their values there), no tile at all, more column chunks than the
 * shared-memory histogram holds (n_cols > 4.19 M at 64 columns per chunk), scales to be derived from a non-square matrix. */
GCNB_API int gcnb_bittile_plan_create_device(const uint32_t *d_indptr, const uint32_t *d_indices, const float *d_values,
                                             int64_t n_rows, int64_t n_cols, const float *d_row_scale, const float *d_col_scale,
                                             int min_tile_nnz, int chunk_cols, int row_blocks, gcnb_stream_t stream,
                                             gcnb_bittile_plan **out);
GCNB_API int gcnb_bittile_device_build_fits(int64_t n_cols, int chunk_cols /*64 or 128*/);
/* test aid: element counts and contents of a plan's device arrays, whichever builder made them.  which: 0 tile_chunk (u32),
 * 1 bits (u64), 2 cta_tile_ptr, 3 cta_item_ptr, 4 items (u32 x 2), 5 row_scale (f32), 6 col_scale, 7 ELL idx (u32), 8 ELL off,
 * 9 ELL steps, 10 ELL rows, 11 ELL split_row, 12 ELL split_ptr; sizes[13] = ELL slots, [14] = entries in tiles, [15] = remainder */
GCNB_API int gcnb_bittile_plan_sizes(const gcnb_bittile_plan *plan, int64_t sizes[16]);
GCNB_API int gcnb_bittile_plan_copy(const gcnb_bittile_plan *plan, int which, void *h_dst, int64_t bytes);
/* 1 when the current device can run the bit-tile kernels (tcgen05 / TMEM: compute capability 10.x) */
GCNB_API int gcnb_bittile_supported(void);
/* out = {tiles, entries in tiles, remainder entries, items' row blocks, columns per tile + 1000 * row_blocks
 * (+ 100000 when the remainder runs on the pattern-only ELL kernel below), CTAs, bit-map bytes, packed-B bytes} */
GCNB_API int gcnb_bittile_plan_info(const gcnb_bittile_plan *plan, int64_t out[8]);
GCNB_API int gcnb_bittile_spmm16_f32(gcnb_bittile_plan *plan, const float *d_B, float *d_C, gcnb_stream_t stream);
/* A plan built from a RENUMBERED square matrix (rows and columns permuted alike -- gcnb_reorder_communities +
 * gcnb_permute_csr of include/gcnb_engine.h): plan index k is row h_old_of_new[k] of the caller's B and C.  The pack
 * kernel gathers B through it and both halves of the product add their rows to C through it: the caller's operands keep
 * their numbering.  Needs a plan whose remainder is the ELL kernel (every entry factors).  h_values of
 * gcnb_bittile_plan_create may be NULL when both scale arrays are given: the matrix is then the PATTERN scaled by them. */
GCNB_API int64_t gcnb_bittile_plan_unfactored(const gcnb_bittile_plan *plan); /* entries with value != row_scale * col_scale */
GCNB_API int gcnb_bittile_plan_set_permutation(gcnb_bittile_plan *plan, const uint32_t *h_old_of_new, gcnb_stream_t stream);
/* kernels launched per 16-column product (pack, MMA kernel, remainder [+ combine], [+ final add]) */
GCNB_API int gcnb_bittile_plan_launches(const gcnb_bittile_plan *plan);
/* Routes later gcnb_spmm_f32 / gcnb_spmm_ld_f32 calls on `plan` that use exactly this d_values pointer, no permutation,
 * 16 contiguous columns (ldb == ldc == 16) through the bit-tile plan (built from the same CSR and the same values);
 * every other call is unaffected.  The bit-tile plan is borrowed (destroy it after the spmm plan); bt = NULL detaches. */
GCNB_API int gcnb_spmm_plan_attach_bittile(gcnb_spmm_plan *plan, gcnb_bittile_plan *bt, const float *d_values);
/* the same product on column slabs of wider row-major matrices (row strides ldb / ldc floats, 16 <= dim <= ldb, ldc): runs
 * 16 columns at a time (pack + MMA kernel + remainder + add per slab), the last slab shifted left to end at dim */
GCNB_API int gcnb_bittile_spmm_ld_f32(gcnb_bittile_plan *plan, const float *d_B, int64_t ldb, float *d_C, int64_t ldc, int dim,
                                      gcnb_stream_t stream);
/* debugging aid: bit mask of the steps gcnb_bittile_spmm16_f32 runs (1 pack, 2 MMA kernel, 4 remainder, 8 final add; default 15) */
GCNB_API int gcnb_bittile_debug_parts(gcnb_bittile_plan *plan, int parts);
/* debugging aid: runs the packing kernel alone and copies the bf16 operand image of B (6144 bytes per 64 rows) to h_out */
GCNB_API int gcnb_bittile_debug_pack(gcnb_bittile_plan *plan, const float *d_B, void *h_out, int64_t bytes,
                                     gcnb_stream_t stream);
/* The builder on its own, host memory only, no CUDA call (tests/test_bittile_cpu.py consumes the arrays exactly as the
 * kernel does).  sizes: {n_rows, n_cols, nnz, blocks of 128 * row_blocks rows, tiles, entries in tiles, items, CTAs, columns per tile,
 * row_blocks, remainder entries that do not factor, remainder entries}; copy: which = 0
 * tile_chunk, 1 bits (uint64 x row_blocks x 128 x columns/64 per tile), 2 cta_tile_ptr, 3 cta_item_ptr, 4 items (uint32 x 2), 5 r_indptr,
 * 6 r_indices, 7 r_values, 8 row_scale, 9 col_scale (layouts: BitTileHost in spmm_bittile.cu). */
typedef struct gcnb_bittile_host gcnb_bittile_host;
GCNB_API int gcnb_bittile_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, const float *h_values,
                                     int64_t n_rows, int64_t n_cols, const float *h_row_scale, const float *h_col_scale,
                                     int min_tile_nnz, int chunk_cols /*0 = 64*/, int row_blocks /*0 = 1*/,
                                     int n_cta /*0 = 148*/, int n_threads /*0 = auto*/, gcnb_bittile_host **out);
GCNB_API int gcnb_bittile_host_sizes(const gcnb_bittile_host *h, int64_t out[12]);
GCNB_API int gcnb_bittile_host_copy(const gcnb_bittile_host *h, int which, void *dst, int64_t bytes);
GCNB_API int gcnb_bittile_host_destroy(gcnb_bittile_host *h);

/* Pattern-only row gather at width 16 (csrc/spmm_ell.cu): R[i][:] = row_scale[i] * sum over the CSR entries (i, j) of
 * B2[j][:].  The remainder kernel of a bit-tile plan whose entries all factor (value = row_scale * col_scale): the pack
 * kernel writes B2 = diag(col_scale) * B once per launch, so entries carry no value.  Replaces the part of
 * graphsum_kernel (src/module.cu:172-186) the tensor-core tiles do not cover.  d_B2 holds n_cols + 1 rows of 16
 * floats, the LAST ONE ALL ZERO (padding entries point at it).  Deterministic (fixed summation order).
 * gcnb_ell_host_*: the host builder alone, for CPU tests that consume the arrays as the kernel does;
 *   sizes: {rows, cols, entries, bundles, index words, split rows, partial slots, wide-bundle minimum length}
 *   copy which: 0 idx (uint32), 1 off (bundles + 1), 2 steps, 3 rows (bundles x 8), 4 split_row, 5 split_ptr */
typedef struct gcnb_ell_host gcnb_ell_host;
typedef struct gcnb_ell_plan gcnb_ell_plan;
GCNB_API int gcnb_ell_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols,
                                 int n_threads, gcnb_ell_host **out);
GCNB_API int gcnb_ell_host_sizes(const gcnb_ell_host *h, int64_t out[8]);
GCNB_API int gcnb_ell_host_copy(const gcnb_ell_host *h, int which, void *dst, int64_t bytes);
GCNB_API int gcnb_ell_host_destroy(gcnb_ell_host *h);
GCNB_API int gcnb_ell_plan_create(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols,
                                  gcnb_stream_t stream, gcnb_ell_plan **out);
GCNB_API int gcnb_ell_plan_destroy(gcnb_ell_plan *plan);
GCNB_API int gcnb_ell_gather16_f32(gcnb_ell_plan *plan, const float *d_B2, const float *d_row_scale, float *d_R,
                                   gcnb_stream_t stream);

/* C[n_rows x dim] = A_csr * B[n_cols x dim],  A_csr values = d_values[e] (or d_values[d_perm[e]] if d_perm).
 *   GraphSum::forward/backward + graphsum_kernel            src/module.cu:172-210  (values = graph_value)
 *   SparseMatmul::forward + sparse_matmul_kernel_forward     src/module.cu:108-132  (values = input Variable)
 * Fully overwrites C.  Summation: fixed order (segment-local lane tree, then ascending segments). */
GCNB_API int gcnb_spmm_f32(gcnb_spmm_plan *plan, const float *d_values, const uint32_t *d_perm, const float *d_B,
                           float *d_C, int dim, gcnb_stream_t stream);
/* Same product on COLUMN SLABS of wider row-major matrices: B has row stride ldb floats, C row stride ldc floats
 * (both >= dim); only columns [0, dim) of the two pointers are read / written.  This is how a wide GraphSum
 * (hidden 600, parameters/parameters_reddit.txt:5) runs: 16 columns at a time, so that the window-staged kernels apply
 * and the slab of B being gathered stays on chip.  A staged plan does this slab loop by itself for any
 * dim >= 64, and the generic kernel cuts operands much larger than L2 into L2-sized slabs. */
GCNB_API int gcnb_spmm_ld_f32(gcnb_spmm_plan *plan, const float *d_values, const uint32_t *d_perm, const float *d_B,
                              int64_t ldb, float *d_C, int64_t ldc, int dim, gcnb_stream_t stream);

/* Transposed-CSR companion for SparseMatmul::backward (src/module.cu:136-163; atomicAdd there, fixed-order here):
 * builds on the device the CSC of a CSR (column pointers, row ids, and the permutation into the CSR value array
 * so that dropped-out input values are picked up without rebuilding).  Outputs are allocated by the call and
 * owned by the returned handle. */
typedef struct gcnb_csc gcnb_csc;
GCNB_API int gcnb_csc_create(const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows, int64_t n_cols,
                             gcnb_stream_t stream, gcnb_csc **out);
GCNB_API int gcnb_csc_destroy(gcnb_csc *csc);
GCNB_API int gcnb_csc_arrays(const gcnb_csc *csc, const uint32_t **d_colptr, const uint32_t **d_rowidx,
                             const uint32_t **d_perm, int *is_dense /* every row holds columns 0..n_cols-1 */);

/* ---------------------------------------------------------------------------------------------------
 * Dense tall-skinny products (Matmul module, src/module.cu:270-472).  Row-major fp32, FFMA (no TF32: the
 * parity bar is 1e-5).  `ws` = caller workspace for the split-K partials of the TN product
 * (gcnb_matmul_tn_workspace() bytes); fixed-order second pass => deterministic (reference: atomicAdd :389).
 *   NN: C[m x p]  = A[m x n]   * B[n x p]      Matmul::forward            :319-328
 *   NT: dA[m x n] = dC[m x p]  * B[n x p]^T    matmul_kernel_backward_1   :332-374
 *   TN: dB[n x p] = A[m x n]^T * dC[m x p]     matmul_kernel_backward_2   :377-391
 * ------------------------------------------------------------------------------------------------- */
GCNB_API int gcnb_matmul_nn_f32(const float *d_A, const float *d_B, float *d_C, int64_t m, int n, int p,
                                gcnb_stream_t stream);
GCNB_API int gcnb_matmul_nt_f32(const float *d_dC, const float *d_B, float *d_dA, int64_t m, int n, int p,
                                gcnb_stream_t stream);
GCNB_API int64_t gcnb_matmul_tn_workspace(int64_t m, int n, int p);
GCNB_API int gcnb_matmul_tn_f32(const float *d_A, const float *d_dC, float *d_dB, int64_t m, int n, int p,
                                void *d_ws, int64_t ws_bytes, gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Elementwise / RNG.
 * Stateless Philox4x32-10 reproducing the reference's cuRAND streams with no state array:
 * element 4g+k of an RNG op = lane k of Philox(ctr=(t,0,g,0), key=(seed,0)), t = number of earlier RNG ops that
 * covered state g (src/variable.cu:5-11,44-61; src/module.cu:16-63).  The caller describes "earlier ops" by
 * up to GCNB_MAX_RNG_HIST (n_groups, count) pairs: t(g) = sum(count_i for n_groups_i > g).
 * ------------------------------------------------------------------------------------------------- */
#define GCNB_MAX_RNG_HIST 48 /* an L-layer model has up to 2L distinct RNG consumer sizes: models up to 24 layers */
typedef struct {
  uint32_t seed;
  uint32_t group_offset; /* row-partitioned ranks: a local slab whose element 0 is GLOBAL element e0 passes        */
  uint32_t elem_lead;    /* group_offset = e0 / 4 and elem_lead = e0 % 4 (masks then do not depend on the partition) */
  int n_hist;
  uint32_t hist_groups[GCNB_MAX_RNG_HIST]; /* ceil(size/4) of an earlier RNG consumer */
  uint32_t hist_count[GCNB_MAX_RNG_HIST];  /* how many times it has run */
} gcnb_rng_t;

/* Variable::glorot (src/variable.cu:44-83): w = (u - 0.5) * 2*sqrtf(6/(rows+cols)), double arithmetic. */
GCNB_API int gcnb_glorot_f32(float *d_w, int64_t size, uint32_t rows, uint32_t cols, const gcnb_rng_t *rng,
                             gcnb_stream_t stream);
/* Dropout::forward (src/module.cu:16-76): x *= (u >= p) ? scale : 0 in place; d_mask (1 byte/element) optional.
 * If d_ext_mask != NULL the keep decisions are READ from it instead of drawn (injected masks). */
GCNB_API int gcnb_dropout_fwd_f32(float *d_x, uint8_t *d_mask, const uint8_t *d_ext_mask, int64_t size, float p,
                                  const gcnb_rng_t *rng, gcnb_stream_t stream);
/* Same, out of place (d_src is left intact): lets the GCN driver keep the feature values pristine instead of the
 * reference's in-place drop + set_input restore copy every eval (src/gcn.cu:181-200). */
GCNB_API int gcnb_dropout_fwd_oop_f32(const float *d_src, float *d_dst, uint8_t *d_mask, const uint8_t *d_ext_mask,
                                      int64_t size, float p, const gcnb_rng_t *rng, gcnb_stream_t stream);
/* Dropout::backward (src/module.cu:80-99): g *= mask ? scale : 0. */
GCNB_API int gcnb_dropout_bwd_f32(float *d_g, const uint8_t *d_mask, int64_t size, float p, gcnb_stream_t stream);
/* ReLU::forward/backward (src/module.cu:222-265); mask written only when training. */
GCNB_API int gcnb_relu_fwd_f32(float *d_x, uint8_t *d_mask, int64_t size, int training, gcnb_stream_t stream);
GCNB_API int gcnb_relu_bwd_f32(float *d_g, const uint8_t *d_mask, int64_t size, gcnb_stream_t stream);
/* Fused pair used by the GCN driver: ReLU then Dropout in one pass, one packed mask byte per element
 * (bit0 = relu keep, bit1 = dropout keep); backward applies dropout-bwd then relu-bwd like the module chain. */
GCNB_API int gcnb_relu_dropout_fwd_f32(float *d_x, uint8_t *d_mask, const uint8_t *d_ext_mask, int64_t size, float p,
                                       int training, const gcnb_rng_t *rng, gcnb_stream_t stream);
GCNB_API int gcnb_relu_dropout_bwd_f32(float *d_g, const uint8_t *d_mask, int64_t size, float p, gcnb_stream_t stream);
/* Parser::calculateGraphValues (src/parser.cpp:164-181) on the device: d_out[e] = 1. / sqrtf(deg(src) * deg(dst)) with
 * the reference's arithmetic (unsigned product -> float -> sqrtf -> double divide -> fp32), bit-identical to the host
 * loop.  Square graphs only (degrees of both endpoints come from d_indptr). */
GCNB_API int gcnb_graph_values_f32(const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows, float *d_out,
                                   gcnb_stream_t stream);
/* GCN::set_truth (src/gcn.cu:204-226). */
GCNB_API int gcnb_set_truth(int32_t *d_truth, const uint32_t *d_split, const int32_t *d_label, int64_t n,
                            uint32_t current_split, gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * First layer on a DENSE feature matrix (every svmlight row lists all F columns, e.g. Reddit's 602): the
 * SparseMatmul::forward/backward products (src/module.cu:108-163) with Dropout(input) (src/module.cu:16-76) applied
 * on the fly from a 1-bit/element keep mask -- X is streamed once per product, never rewritten, never restored
 * (src/gcn.cu:181-200 copies it back before every eval).  d_bits == NULL means "no dropout" (eval).
 * Supported when gcnb_dense_feat_supported(F, P): P in {8,16,32}, F <= 1024; otherwise use the generic products.
 * ------------------------------------------------------------------------------------------------- */
GCNB_API int gcnb_dense_feat_supported(int f, int p);
/* Keep decisions (u >= p) of the n_rows*f elements, same Philox stream and lane order as gcnb_dropout_fwd_f32, stored
 * 1 bit per element in the row-tile layout the two products bulk-copy: rows in tiles of 32, element (r, k) is bit
 * (r % 32) * f + k of tile r / 32, every tile padded to a multiple of 16 bytes (gcnb_dropout_maskbits_words words). */
GCNB_API int64_t gcnb_dropout_maskbits_words(int64_t n_rows, int f);
GCNB_API int gcnb_dropout_maskbits(uint32_t *d_bits, int64_t n_rows, int f, float p, const gcnb_rng_t *rng,
                                   gcnb_stream_t stream);
/* out[n x p] = (X .* keep/(1-p_drop))[n x f] * W[f x p] */
GCNB_API int gcnb_dense_feat_fwd_f32(const float *d_X, const uint32_t *d_bits, float p_drop, const float *d_W,
                                     float *d_out, int64_t n, int f, int p, gcnb_stream_t stream);
/* dW[f x p] = (X .* keep/(1-p_drop))^T * dH[n x p]; per-CTA row-slab partials in d_ws, reduced in slab order */
GCNB_API int64_t gcnb_dense_feat_tn_workspace(int64_t n, int f, int p);
GCNB_API int gcnb_dense_feat_tn_f32(const float *d_X, const uint32_t *d_bits, float p_drop, const float *d_dH,
                                    float *d_dW, int64_t n, int f, int p, void *d_ws, int64_t ws_bytes,
                                    gcnb_stream_t stream);

/* Exact-split tcgen05 GEMM for a WIDE first layer (hidden 600, parameters/parameters_reddit.txt): out[n x p] = X[n x f] * W
 * with X and W each carried as three bf16 pieces and the six piece products of weight >= 2^-24 accumulated in fp32 in TMEM
 * (parallel-gcn_b200/csrc/dense_tc.cu).  X is packed once per dataset into the operand image the MMA reads
 * (gcnb_dense_tc_x_bytes / gcnb_dense_tc_pack_x); W is packed per call into d_ws (gcnb_dense_tc_w_bytes).  No dropout on X
 * (the wide configuration has none; evaluation never has).  Status: written in round 1, not yet run on a GPU; nothing in
 * the engine calls it yet. */
GCNB_API int gcnb_dense_tc_supported(int f, int p);
GCNB_API int64_t gcnb_dense_tc_x_bytes(int64_t n, int f);
GCNB_API int64_t gcnb_dense_tc_w_bytes(int f, int p);
GCNB_API int gcnb_dense_tc_pack_x(const float *d_X, void *d_img, int64_t n, int f, gcnb_stream_t stream);
GCNB_API int gcnb_dense_tc_fwd_f32(const void *d_x_img, const float *d_W, float *d_out, int64_t n, int f, int p, void *d_ws,
                                   int64_t ws_bytes, gcnb_stream_t stream);
/* weight gradient dW[f x p] = X^T * dH (SparseMatmul::backward on a dense X, src/module.cu:136-163; atomicAdd there): X^T is
 * packed once (gcnb_dense_tc_xt_bytes / gcnb_dense_tc_pack_xt), dH per call; K (the nodes) is cut into slices whose partial
 * tiles are added in ascending order.  d_ws: gcnb_dense_tc_tn_workspace(n, f, p) bytes. */
GCNB_API int64_t gcnb_dense_tc_xt_bytes(int64_t n, int f);
GCNB_API int gcnb_dense_tc_pack_xt(const float *d_X, void *d_img, int64_t n, int f, gcnb_stream_t stream);
GCNB_API int64_t gcnb_dense_tc_tn_workspace(int64_t n, int f, int p);
GCNB_API int gcnb_dense_tc_tn_f32(const void *d_xt_img, const float *d_dH, float *d_dW, int64_t n, int f, int p, void *d_ws,
                                  int64_t ws_bytes, gcnb_stream_t stream);
GCNB_API int gcnb_dense_tc_debug_pack_w(const float *d_W, void *d_ws, int f, int p, gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * CrossEntropyLoss::forward (src/module.cu:484-541) + GCN::get_accuracy (src/gcn.cu:264-289) in one pass.
 * In-place logits -= rowmax for rows with truth >= 0 (API-visible side effect kept); if training:
 * grad = softmax/num_samples, grad[truth] -= 1.0/num_samples (double), rows with truth < 0 get grad 0.
 * d_result[0] = un-normalised loss sum (float), d_result[1] = bit pattern of the uint32 wrong count,
 * d_result[2] = bit pattern of the uint32 labelled-row count.  Deterministic two-level reduction.
 * d_ws: gcnb_ce_workspace(n) bytes, zero-filled ONCE by the caller (cudaMemset) before the first launch; the
 *       kernel leaves it ready for the next launch (same rule for gcnb_sumsq_f32).
 * ------------------------------------------------------------------------------------------------- */
GCNB_API int64_t gcnb_ce_workspace(int64_t n);
GCNB_API int gcnb_softmax_ce_f32(float *d_logits, float *d_grad, const int32_t *d_truth, int64_t n, int num_classes,
                                 uint32_t num_samples, int training, float *d_result, void *d_ws,
                                 gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Output head in one kernel (csrc/head.cu): the last layer's Matmul::forward (src/module.cu:274-328) on the narrow
 * GraphSum result y [n x in_dim], CrossEntropyLoss::forward (:484-541) + get_accuracy (src/gcn.cu:264-289) and, when
 * training, Matmul::backward (src/module.cu:332-391): d_dy = dz W^T and the partial sums of dW = y^T dz.  Same element
 * arithmetic as gcnb_matmul_nn_f32 / gcnb_softmax_ce_f32 / gcnb_matmul_nt_f32 (bit-identical logits, loss, dy);
 * d_logits receives the CE-shifted logits; d_grad (dz, [n x classes]) may be NULL: it then never reaches memory.
 * d_result as for gcnb_softmax_ce_f32.  gcnb_head_reduce_dw_f32 adds the partial sums in a fixed order (any stream that
 * is ordered after the head kernel).  in_dim 8 / 16 / 32, classes <= 64 (gcnb_head_supported); d_y / d_dy 16-byte aligned.
 * d_ws: gcnb_head_workspace() bytes, zero-filled ONCE by the caller before the first launch.
 * ------------------------------------------------------------------------------------------------- */
GCNB_API int gcnb_head_supported(int in_dim, int num_classes);
GCNB_API int64_t gcnb_head_workspace(int64_t n, int in_dim, int num_classes);
GCNB_API int gcnb_head_f32(const float *d_y, const float *d_w, const int32_t *d_truth, int64_t n, int in_dim, int num_classes,
                           uint32_t num_samples, int training, float *d_logits, float *d_grad, float *d_dy, float *d_result,
                           void *d_ws, int64_t ws_bytes, gcnb_stream_t stream);
/* the same on the tensor cores (mma.sync TF32, x = hi + lo with all four piece products: fp32-level accuracy, other bits than
 * the FMA chains); in_dim 16, classes <= 48; same workspace and gcnb_head_reduce_dw_f32 */
GCNB_API int gcnb_head_tc_supported(int in_dim, int num_classes);
GCNB_API int gcnb_head_tc_f32(const float *d_y, const float *d_w, const int32_t *d_truth, int64_t n, int in_dim, int num_classes,
                              uint32_t num_samples, int training, float *d_logits, float *d_grad, float *d_dy, float *d_result,
                              void *d_ws, int64_t ws_bytes, gcnb_stream_t stream);
GCNB_API int gcnb_head_reduce_dw_f32(const void *d_ws, float *d_dw, int64_t n, int in_dim, int num_classes,
                                     gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Adam::step (src/optim.cu:42-95) for up to GCNB_MAX_TENSORS weights in ONE launch, and
 * GCN::get_l2_penalty (src/gcn.cu:230-260) as a fixed-order sum of squares.
 * ------------------------------------------------------------------------------------------------- */
#define GCNB_MAX_TENSORS 48
typedef struct {
  int n_tensors;
  float *w[GCNB_MAX_TENSORS];
  const float *g[GCNB_MAX_TENSORS];
  float *m[GCNB_MAX_TENSORS];
  float *v[GCNB_MAX_TENSORS];
  int64_t size[GCNB_MAX_TENSORS];
  int decay[GCNB_MAX_TENSORS];
} gcnb_adam_tensors_t;
GCNB_API int gcnb_adam_step_f32(const gcnb_adam_tensors_t *t, float weight_decay, float beta1, float beta2, float eps,
                                float step_size, gcnb_stream_t stream);
/* CUDA-graph replay: `node` is a kernel node of a captured graph that was produced by one of the library's calls whose
 * arguments change from epoch to epoch -- the dropout family (gcnb_dropout_fwd[_oop]_f32, gcnb_relu_dropout_fwd_f32,
 * gcnb_dropout_maskbits: the Philox descriptor) and gcnb_adam_step_f32 (step_size).  Rewrites that argument of the node
 * inside the instantiated graph `exec` (cudaGraphExec_t / cudaGraphNode_t passed as void*), so an epoch captured once
 * is replayed with the next epoch's randomness.  GCNB_E_UNSUPPORTED if the node is not such a kernel. */
GCNB_API int gcnb_graph_patch_node(void *exec, void *node, const gcnb_rng_t *rng, const float *step_size);
/* d_out[0] = sum_i w[i]^2 (ascending fixed tree).  d_ws: gcnb_sumsq_workspace(n) bytes. */
GCNB_API int64_t gcnb_sumsq_workspace(int64_t n);
GCNB_API int gcnb_sumsq_f32(const float *d_w, int64_t n, float *d_out, void *d_ws, gcnb_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Communicator of the row-partitioned multi-GPU engine (one process per GPU).  The reference has no
 * distributed path at all (single GPU: SURVEY 5.8), so there is no reference line to cite; the collectives below are
 * the ones the partitioned epoch needs: all-gather of the [N/R x d] GraphSum input slab, sum all-reduce of the
 * replicated weight gradients and of the loss / count scalars.  NCCL is loaded at run time (libnccl.so.2).
 * Bootstrap: rank 0 calls gcnb_comm_unique_id and ships the GCNB_COMM_ID_BYTES to every rank by any side channel
 * (bench.py: torch.distributed broadcast), then every rank calls gcnb_comm_create with its CUDA device current.
 * world == 1 needs no id and no NCCL.
 * ------------------------------------------------------------------------------------------------- */
#define GCNB_COMM_ID_BYTES 128
typedef struct gcnb_comm gcnb_comm;
GCNB_API int gcnb_comm_unique_id(void *out_id /* GCNB_COMM_ID_BYTES */);
GCNB_API int gcnb_comm_create(int rank, int world, const void *id_bytes, gcnb_comm **out);
GCNB_API int gcnb_comm_destroy(gcnb_comm *c);
GCNB_API int gcnb_comm_rank(const gcnb_comm *c);
GCNB_API int gcnb_comm_world(const gcnb_comm *c);
/* d_recv[world x count_per_rank] = concatenation of every rank's d_send[count_per_rank], in rank order */
GCNB_API int gcnb_comm_all_gather_f32(gcnb_comm *c, const float *d_send, float *d_recv, int64_t count_per_rank,
                                      gcnb_stream_t stream);
/* in-place sum over ranks of `count` float32 (is_u32 = 0) or uint32 (is_u32 = 1) values */
GCNB_API int gcnb_comm_all_reduce_sum(gcnb_comm *c, void *d_buf, int64_t count, int is_u32, gcnb_stream_t stream);
/* Slab gather for the row-partitioned GraphSum (the all-gather of the [N/R x d] input slab) over NVLink
 * peer memory: gcnb_comm_gather_setup (collective, once) allocates two gather buffers of `gather_floats` per rank and
 * opens every peer's through CUDA IPC; gcnb_comm_gather_slabs_f32 then stores this rank's slab into slot `rank` of
 * every rank's buffer and waits for the peers' flags (falls back to ncclAllGather, collectively, when peer memory is
 * unavailable).  *d_full_out (owned by the communicator, valid until the next-but-one gather) holds the world slabs
 * in rank order.  gcnb_comm_gather_mode: 0 single rank, 1 NCCL all-gather, 2 peer-memory push. */
GCNB_API int gcnb_comm_gather_setup(gcnb_comm *c, int64_t gather_floats);
GCNB_API int gcnb_comm_gather_slabs_f32(gcnb_comm *c, const float *d_slab, int64_t count_per_rank,
                                        const float **d_full_out, gcnb_stream_t stream);
/* same; overlapped != 0 tells the communicator that the caller computes on another stream meanwhile: large slabs then
 * travel through the copy engines (peer cudaMemcpyAsync + a flag kernel) instead of the SM push kernel */
GCNB_API int gcnb_comm_gather_slabs_ex_f32(gcnb_comm *c, const float *d_slab, int64_t count_per_rank,
                                           const float **d_full_out, int overlapped, gcnb_stream_t stream);
GCNB_API int gcnb_comm_gather_mode(const gcnb_comm *c);
/* Halo exchange: after this call the slab gathers whose slab is [block x dim] ship to every peer only
 * the rows of this rank's block that the peer's row block references (same buffers, flags and slot layout: column ids
 * need no translation, nothing downstream changes; unreferenced rows of a peer's slot are never written nor read).
 * d_indices / nnz: the column ids of this rank's CSR block in the slot layout (column c = row c % block of rank
 * c / block); rows_local <= block.  Collective (one all-gather of n_global / 8 bytes per rank).  The lists are used when
 * they hold at most max_fraction of what the full push sends (GCNB_HALO=0 / 1 forces it), peer-memory mode only.
 * info: {active, rows sent per exchange, rows of the full push, rows this rank references in other ranks' blocks}.
 * gcnb_halo_lists_from_masks: the host logic behind it (masks = [world][words] bits over the column space; rows_out is
 * malloc'ed, off_out has world + 1 entries) -- exported for CPU tests. */
GCNB_API int gcnb_comm_halo_setup(gcnb_comm *c, const uint32_t *d_indices, int64_t nnz, int64_t rows_local, int64_t block,
                                  double max_fraction, int64_t info[4], gcnb_stream_t stream);
GCNB_API int gcnb_comm_halo_active(const gcnb_comm *c);
GCNB_API int gcnb_halo_lists_from_masks(const uint32_t *masks, int world, int rank, int64_t words, int64_t block,
                                        int64_t rows_local, uint32_t **rows_out, int64_t *off_out);
GCNB_API int gcnb_comm_group_start(gcnb_comm *c);
GCNB_API int gcnb_comm_group_end(gcnb_comm *c);

#ifdef __cplusplus
}
#endif
#endif /* GCNB_H */
