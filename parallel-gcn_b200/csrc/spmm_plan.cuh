// spmm_plan.cuh -- the CSR x dense plan object shared by spmm.cu (generic segment kernel) and spmm_stage.cu
// (window-staged GraphSum kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <vector>

namespace gcnb {

// ---- window staging (spmm_stage.cu) ----------------------------------------------------------------------------
// The columns (= rows of the dense operand B) are cut into windows of `window_rows` rows, one window fitting in an
// SM's shared memory.  Entries (i, j) of rows that hold >= min_seg entries inside a window are STAGED: the CTA that
// owns the window copies it into shared memory once (TMA bulk copy) and gathers the neighbour rows from there with
// conflict-free LDS.128 -- half an LSU wavefront per neighbour row instead of one L1 wavefront, and 2-byte
// window-local column indices instead of 4-byte ones.  Everything else (the REMAINDER) stays a plain CSR and goes
// through the generic kernel.
struct StageParams {
  int dim = 16;                  // feature width the plan is built for (window_rows * dim * 4 bytes of shared memory)
  int window_rows = 0;           // 0 = as many rows as fit (<= 65536: local column ids are 16-bit)
  int min_seg = 16;              // (row, window) pairs with fewer entries are not worth a segment
  int seg_cap = 256;            // longer (row, window) runs are cut into pieces of <= seg_cap entries
  int64_t min_window_nnz = 0;    // 0 = 8 * window_rows: a window must be re-used to pay for its copy
  int n_cta = 148;               // persistent CTAs (one per SM)
  int n_threads = 0;             // host threads for the build (0 = hardware concurrency, <= 16)
  int runs_per_queue = 1;        // > 1 cuts every CTA's share of a window into several runs
  int min_avg_seg = 32;          // staging is skipped when the average segment is shorter (per-segment work dominates)
  // row-partitioned products: columns [own_col0, own_col1) are the rank's own slab of B; windows entirely inside go to
  // a second run list (own_runs) that can be processed before the peers' slabs have arrived
  int64_t own_col0 = 0, own_col1 = 0;
};

// large plan arrays: malloc'ed and NOT value-initialised (a std::vector would zero-fill ~1 GB on one thread before the
// builder's threads overwrite it; the builder touches every element it needs itself, in parallel)
template <class T>
struct HostArray {
  T *p = nullptr;
  size_t n = 0;
  HostArray() = default;
  HostArray(const HostArray &) = delete;
  HostArray &operator=(const HostArray &) = delete;
  HostArray(HostArray &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  HostArray &operator=(HostArray &&o) noexcept {
    if (this != &o) { free(p); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~HostArray() { free(p); }
  void alloc(size_t count) {
    free(p);
    n = count;
    p = count ? static_cast<T *>(malloc(count * sizeof(T))) : nullptr;
  }
  T *data() { return p; }
  const T *data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

constexpr int kStageLanes = 32;                 // segments per bundle (one per lane)
constexpr int kStageBlock = 4;                  // steps per packed block (one 8-byte id load + one 16-byte value load)
constexpr uint32_t kStagePad = 0xffffffffu;     // padding marker in the value permutation / unused slot

// Host-side result of the build (pure CPU, unit-tested without a GPU through gcnb_stage_host_*).
// A SEGMENT is the run of one row's entries inside one window (cut into pieces of <= seg_cap).  Segments of a
// window are sorted by length and dealt 32 at a time into BUNDLES: lane l of the warp that processes a bundle owns
// segment l entirely (private accumulators, no cross-lane reduction), step k of the bundle handles entry k of all 32
// segments (ELL layout, so ids / values of a step are one coalesced load).
//   bundles[b] = (first block, steps L = longest segment, shortest segment (0 if a lane is unused), 0)
//   lens[b*32 + l] = entries of lane l's segment
//   block = 4 steps x 32 lanes: pidx[(block*32 + l)*4 + k%4] = window-local column of entry k of lane l;
//           pperm (same addressing) = position of that entry in the original CSR value array, kStagePad for padding
//   entry order inside a segment alternates even/odd local columns, and lanes with bit 2 set start on odd: the two
//   lanes of a quarter-warp that read the same 16-byte column chunk then always hit opposite bank halves
//   lane_slot[b*32 + l] = partial slot written by lane l (kStagePad if the lane is unused); the slots of row r are the
//   contiguous range [row_slot[r], row_slot[r+1]) in ascending (window, piece) order
//   runs[r] = (window, first bundle, end bundle, 0);  CTA q processes runs [run_begin[q], run_begin[q+1])
//   r_* : remainder CSR (original entry order inside a row) and its permutation into the original value array
struct StagedHost {
  int dim = 0, window_rows = 0, n_win = 0, n_cta = 0;
  int64_t n_rows = 0, n_cols = 0, nnz = 0, staged_nnz = 0, n_blocks = 0, n_slots = 0, n_segs = 0;
  std::vector<uint4> bundles, runs, own_runs;
  std::vector<uint16_t> lens;
  std::vector<uint32_t> run_begin, own_run_begin;
  HostArray<uint16_t> pidx;
  HostArray<uint32_t> pperm;
  std::vector<uint32_t> row_slot, lane_slot;
  std::vector<uint32_t> r_indptr;
  HostArray<uint32_t> r_indices, r_perm;
};

int stage_build_host(const uint32_t *indptr, const uint32_t *indices, int64_t n_rows, int64_t n_cols,
                     const StageParams &params, StagedHost &out);

// ---- pattern-only row gather at width 16 (spmm_ell.cu): the remainder of a bit-tile plan whose entries all factor --------
constexpr int kEllWideMin = 256;   // rows with more entries are shared by the 8 lane groups of a warp (wide bundle)
constexpr int kEllFewRows = 65536;          // plans with fewer rows cannot fill the machine with 8 rows per warp:
constexpr int kEllWideMinFewRows = 64;      // there rows of more than 64 entries are shared already
constexpr int kEllWideMax = 8192;  // longer rows are cut into parts with partial slots
// bundles in ticket order; off[b] = first uint4-row of bundle b (a uint4-row = 8 groups x 4 indices = 128 bytes);
// steps[b] = uint4-rows of the bundle | 0x80000000 for a wide bundle; rows[b*8 + g] = row owned by lane group g
// (0xffffffff: unused) or, wide, rows[b*8] = the row and rows[b*8 + 1] = its partial slot (0xffffffff: writes R directly);
// split_row[k] = k-th row assembled from the slots [split_ptr[k], split_ptr[k+1])
struct EllHost {
  int64_t n_rows = 0, n_cols = 0, nnz = 0, n_bundles = 0, n_slots = 0;
  int wide_min = kEllWideMin;  // rows with more entries are wide bundles (chosen by the layout)
  HostArray<uint32_t> idx;
  std::vector<uint32_t> off, steps, rows, split_row, split_ptr;
};
struct EllDev {
  int64_t n_rows = 0, n_cols = 0, nnz = 0, n_bundles = 0, n_split = 0, n_slots = 0;
  uint32_t *d_idx = nullptr, *d_off = nullptr, *d_steps = nullptr, *d_rows = nullptr, *d_split_row = nullptr,
           *d_split_ptr = nullptr, *d_counter = nullptr;
  float *d_slots = nullptr;
  bool static_schedule = false;  // few-rows plans: bundles dealt round-robin to the warps instead of through the ticket counter
};
int ell_build_host(const uint32_t *indptr, const uint32_t *indices, int64_t n_rows, int64_t n_cols, int n_threads, EllHost &out);
int ell_upload_plan(const EllHost &H, cudaStream_t stream, EllDev **out);
void ell_destroy(EllDev *e);
// the same plan from device arrays (spmm_ell.cu): d_src[d_row_beg[i] + k] = column of entry k of row i, k < d_row_len[i]
int ell_build_device(const uint32_t *d_row_beg, const uint32_t *d_row_len, const uint32_t *d_src, int64_t n_rows, int64_t n_cols,
                     cudaStream_t stream, EllDev **out);
// R[n_rows x 16] (row stride ldr floats, 16-byte aligned rows) = or += diag(row_scale) * pattern * B2 (B2: n_cols + 1 rows,
// the last one zero); accumulate: vector reductions into R instead of stores; d_row_map (optional): plan row k is written
// to row d_row_map[k] of R; ctas_per_sm 0 = default
int ell_launch(EllDev *e, const float *d_B2, const float *d_row_scale, float *d_R, int64_t ldr, int accumulate,
               const uint32_t *d_row_map, int ctas_per_sm, cudaStream_t stream);

struct StagedDev;  // device mirror, spmm_stage.cu
void stage_destroy(StagedDev *s);

}  // namespace gcnb

struct gcnb_spmm_plan {
  const uint32_t *d_indptr = nullptr;
  const uint32_t *d_indices = nullptr;
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  int seg_nnz = 0;
  int64_t n_seg = 0, n_split_rows = 0, n_slots = 0;
  int n_queues = 0;
  uint4 *d_segs = nullptr;         // (row, begin, end, slot or 0xffffffff)
  uint32_t *d_queue_begin = nullptr;  // n_queues + 1
  uint32_t *d_counters = nullptr;     // n_queues (+1 done counter)
  uint32_t *d_split_row = nullptr;    // n_split_rows
  uint32_t *d_split_slot = nullptr;   // n_split_rows + 1
  float *d_scratch = nullptr;
  int64_t scratch_dim = 0;
  int64_t max_deg = 0;
  int batch = 1;                      // segments claimed per atomic ticket (short-row graphs: > 1)
  int max_cta_per_sm = 0;             // > 0 caps the persistent grid (a co-resident kernel needs the rest of the SM)
  gcnb::StagedDev *staged = nullptr;  // optional window-staged fast path (gcnb_spmm_plan_stage)
  int64_t own_col0 = 0, own_col1 = 0;  // gcnb_spmm_plan_set_own_cols (before staging)
  // optional bit-tile representation (spmm_bittile.cu), borrowed: used for 16-column contiguous products with this value array
  struct gcnb_bittile_plan *bittile = nullptr;
  const float *bittile_values = nullptr;
};

// spmm_stage.cu: runs the staged path if it applies to this call (same values pointer, same dim, no permutation) and
// sets *handled = 1; otherwise *handled = 0 and the caller runs the generic kernel.  Returns 0 or an error code.
// generic segment kernel on any plan (never the staged path), spmm.cu
namespace gcnb {
int spmm_generic_launch(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, int ldb,
                        float *d_C, int ldc, int dim, cudaStream_t stream);
}
// B and C may be column slabs of wider matrices (row strides ldb / ldc floats); dims that are multiples of the staged
// width run slab by slab.
int gcnb_stage_try_spmm(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, int ldb,
                        float *d_C, int ldc, int dim, cudaStream_t stream, int *handled);
