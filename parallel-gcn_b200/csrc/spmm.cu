// spmm.cu -- CSR x dense product for sm_100a: GraphSum (normalised adjacency x activations, forward and
// backward) and SparseMatmul forward (svmlight features x W0), plus the transposed (CSC) view used by
// SparseMatmul backward.  Replaces graphsum_kernel / sparse_matmul_kernel_forward / _backward of the reference
// (src/module.cu:108-186), which run one thread per OUTPUT ELEMENT with no load balancing and fp32 atomics.
//
// Design (B200): HBM streams (indices, values) are read exactly once with L1::no_allocate; the gathered
// neighbour rows go through L1/L2.  Work unit = row segment of <= seg_nnz entries handled by one warp:
// the 32 lanes are LPR lanes across the feature dimension (VEC floats each, float4 when dim % 4 == 0) times
// G = 32/LPR neighbours in flight; a 32-entry chunk of indices/values is loaded coalesced, one per lane, and
// broadcast with shuffles.  Segments are dealt to per-SM queues holding CONTIGUOUS rows with equal nnz, so an
// SM's gathers concentrate on one neighbourhood (L1 reuse on community-ordered graphs) and degree skew is
// absorbed by segment granularity; idle SMs steal from the other queues.  Rows longer than one segment are
// combined from per-segment partials in ascending order by a second tiny kernel: no floating-point atomics.
#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <thread>
#include <vector>

#include "common.cuh"
#include "spmm_plan.cuh"

namespace gcnb {

const DeviceInfo &device_info() {
  static DeviceInfo info = [] {
    DeviceInfo d;
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      d.sm_count = p.multiProcessorCount;
      d.cc_major = p.major;
      d.ok = 1;
    }
    return d;
  }();
  return info;
}

static int g_host_threads = 0;
int host_threads() {
  if (g_host_threads > 0) return g_host_threads;
  if (const char *e = getenv("GCNB_HOST_THREADS"))
    if (atoi(e) > 0) return atoi(e);
  return (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
}

struct Patchable {
  const void *kernel;
  int rng_arg, step_arg;
};
static std::vector<Patchable> &patch_table() {
  static std::vector<Patchable> t;
  return t;
}

void register_patchable(const void *kernel, int rng_arg, int step_arg) { patch_table().push_back({kernel, rng_arg, step_arg}); }
bool lookup_patchable(const void *kernel, int *rng_arg, int *step_arg) {
  for (const Patchable &p : patch_table())
    if (p.kernel == kernel) {
      *rng_arg = p.rng_arg;
      *step_arg = p.step_arg;
      return true;
    }
  return false;
}

}  // namespace gcnb

using namespace gcnb;

namespace {

constexpr int kThreads = 256;
#ifndef GCNB_BLOCKS_NARROW
#define GCNB_BLOCKS_NARROW 4
#endif
constexpr int kBlocksNarrow = GCNB_BLOCKS_NARROW;  // resident CTAs per SM asked of ptxas for the one-tile kernels
constexpr uint32_t kNoSlot = 0xffffffffu;
constexpr int64_t kStaticMaxSegs = 32768;  // up to this many segments: one warp per segment, no queues (see spmm_seg_kernel)

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const char *p) { v = __ldg(reinterpret_cast<const float4 *>(p)); }
  __device__ __forceinline__ void fma(float a, const Vec<4> &x) {
    v.x = fmaf(a, x.v.x, v.x);
    v.y = fmaf(a, x.v.y, v.y);
    v.z = fmaf(a, x.v.z, v.z);
    v.w = fmaf(a, x.v.w, v.w);
  }
  __device__ __forceinline__ void xor_add(int o) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  __device__ __forceinline__ void store(float *p) const { *reinterpret_cast<float4 *>(p) = v; }
};
template <>
struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const char *p) { v = __ldg(reinterpret_cast<const float *>(p)); }
  __device__ __forceinline__ void fma(float a, const Vec<1> &x) { v = fmaf(a, x.v, v); }
  __device__ __forceinline__ void xor_add(int o) { v += __shfl_xor_sync(0xffffffffu, v, o); }
  __device__ __forceinline__ void store(float *p) const { *p = v; }
};

// One warp per claimed segment.  KT = column tiles held per lane (dim <= VEC*LPR*KT handled in one pass);
// EXACT: dim == VEC*LPR*KT, so no column predicate is needed.
template <int VEC, int LPR, int KT, bool EXACT>
struct SegArgs {
  const uint4 *__restrict__ segs;
  const uint32_t *__restrict__ indices;
  const float *__restrict__ values;
  const uint32_t *__restrict__ perm;
  const float *__restrict__ B;
  float *__restrict__ C;
  float *__restrict__ scratch;
  int dim;
  int ldb, ldc;  // row strides of B and C in floats (>= dim): a column slab of a wider matrix is a valid operand
};

// gathers + FMAs of one 32-entry chunk held one-per-lane in (idx, val).  Loads are issued in batches of UB
// independent gathers before any FMA consumes them (memory-level parallelism); TAIL chunks predicate the loads
// (never multiply a foreign row by a padding zero: 0 * inf would poison the sum).
template <int VEC, int LPR, int KT, bool EXACT, bool TAIL>
__device__ __forceinline__ void chunk_fma(Vec<VEC> (&acc)[KT], const char *__restrict__ Bl, uint32_t row_bytes,
                                          uint32_t idx, float val, int cnt, int g, const bool (&colok)[KT]) {
  constexpr int G = 32 / LPR;
  constexpr int U = 32 / G;  // gather steps per chunk == LPR
  constexpr int W = VEC * LPR;
  constexpr int UB0 = 32 / (VEC * KT);
  constexpr int UB = UB0 < 1 ? 1 : (UB0 > 8 ? (U < 8 ? U : 8) : (UB0 > U ? U : UB0));
#pragma unroll
  for (int u0 = 0; u0 < U; u0 += UB) {
    if (TAIL && u0 * G >= cnt) break;  // warp-uniform
    float a[UB];
    Vec<VEC> x[UB][KT];
#pragma unroll
    for (int u = 0; u < UB; u++) {
      const int j = (u0 + u) * G + g;
      const uint32_t c = __shfl_sync(0xffffffffu, idx, j);
      a[u] = __shfl_sync(0xffffffffu, val, j);
      const bool ok = !TAIL || j < cnt;
      const char *p = Bl + (uint64_t)c * row_bytes;
#pragma unroll
      for (int t = 0; t < KT; t++) {
        x[u][t].zero();
        if (ok && (EXACT || colok[t])) x[u][t].load(p + (size_t)t * W * sizeof(float));
      }
    }
#pragma unroll
    for (int u = 0; u < UB; u++)
#pragma unroll
      for (int t = 0; t < KT; t++) acc[t].fma(a[u], x[u][t]);
  }
}

// indices / values of entries [e0 + lane] of a segment (zero beyond its end)
template <int VEC, int LPR, int KT, bool EXACT>
__device__ __forceinline__ void load_chunk(const SegArgs<VEC, LPR, KT, EXACT> &A, uint32_t e, uint32_t end,
                                           uint32_t &idx, float &val) {
  idx = 0;
  val = 0.f;
  if (e < end) {
    idx = ld_stream_u32(A.indices + e);
    val = A.perm ? __ldg(A.values + __ldg(A.perm + e)) : ld_stream_f32(A.values + e);
  }
}

// One segment.  (idx_n, val_n) arrive holding the segment's first 32 entries and leave holding the first 32 entries
// of `next` (the segment this warp runs afterwards, if any): while the last chunk's rows are gathered, the next
// segment's indices are already in flight -- short rows are otherwise a chain of dependent memory latencies.
template <int VEC, int LPR, int KT, bool EXACT>
__device__ __forceinline__ void run_segment(const SegArgs<VEC, LPR, KT, EXACT> &A, const uint4 sg, const int lane,
                                            uint32_t &idx_n, float &val_n, const bool has_next, const uint4 next) {
  constexpr int W = VEC * LPR;
  const int g = lane / LPR, l = lane % LPR;
  const int dim = A.dim;
  const uint32_t row = sg.x, beg = sg.y, end = sg.z, slot = sg.w;
  const char *Bl = reinterpret_cast<const char *>(A.B + l * VEC);
  const uint32_t row_bytes = (uint32_t)A.ldb * (uint32_t)sizeof(float);
  bool colok[KT];
#pragma unroll
  for (int t = 0; t < KT; t++) colok[t] = EXACT || (t * W + l * VEC < dim);

  Vec<VEC> acc[KT];
#pragma unroll
  for (int t = 0; t < KT; t++) acc[t].zero();

  // software pipeline: indices/values of the next 32 entries are in flight while the current chunk's rows
  // are gathered
  uint32_t base = beg;
  if (base >= end && has_next) load_chunk(A, next.y + lane, next.z, idx_n, val_n);
  while (base < end) {  // warp-uniform
    const uint32_t idx = idx_n;
    const float val = val_n;
    const uint32_t nb = base + 32;
    if (nb < end) load_chunk(A, nb + lane, end, idx_n, val_n);
    else if (has_next) load_chunk(A, next.y + lane, next.z, idx_n, val_n);
    if (nb <= end) chunk_fma<VEC, LPR, KT, EXACT, false>(acc, Bl, row_bytes, idx, val, 32, g, colok);
    else chunk_fma<VEC, LPR, KT, EXACT, true>(acc, Bl, row_bytes, idx, val, (int)(end - base), g, colok);
    base = nb;
  }

  // fixed-order tree over the G neighbour groups
#pragma unroll
  for (int t = 0; t < KT; t++)
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) acc[t].xor_add(o);
  if (g == 0) {
    float *dst = (slot == kNoSlot) ? A.C + (size_t)row * A.ldc : A.scratch + (size_t)slot * dim;
#pragma unroll
    for (int t = 0; t < KT; t++)
      if (colok[t]) acc[t].store(dst + t * W + l * VEC);
  }
}

// Drains one queue.  Two tickets ahead: while segment n runs, the header of segment n+1 is already loaded (its
// first chunk gets prefetched by run_segment) and the atomic that claims segment n+2 is in flight.
template <int VEC, int LPR, int KT, bool EXACT>
__device__ __forceinline__ void drain_queue_lookahead(const SegArgs<VEC, LPR, KT, EXACT> &A, const uint32_t *__restrict__ queue_begin,
                                            uint32_t *__restrict__ counters, int q, int lane) {
  const uint32_t qb = __ldg(queue_begin + q), qn = __ldg(queue_begin + q + 1) - qb;
  if (*((volatile uint32_t *)(counters + q)) >= qn) return;  // already drained: do not pay for an atomic
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(counters + q, 1u);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t >= qn) return;
  uint4 sg = __ldg(A.segs + qb + t);
  if (lane == 0) t = atomicAdd(counters + q, 1u);
  t = __shfl_sync(0xffffffffu, t, 0);
  bool has_next = t < qn;
  uint4 sgn = make_uint4(0, 0, 0, 0);
  if (has_next) sgn = __ldg(A.segs + qb + t);
  uint32_t idx_n;
  float val_n;
  load_chunk(A, sg.y + lane, sg.z, idx_n, val_n);
  for (;;) {
    uint32_t t2 = 0;
    if (has_next && lane == 0) t2 = atomicAdd(counters + q, 1u);  // claims the segment after next; used below
    run_segment<VEC, LPR, KT, EXACT>(A, sg, lane, idx_n, val_n, has_next, sgn);
    if (!has_next) break;
    sg = sgn;
    t2 = __shfl_sync(0xffffffffu, t2, 0);
    has_next = t2 < qn;
    if (has_next) sgn = __ldg(A.segs + qb + t2);
  }
}

// Short rows: drains one queue in batches of kBatch (<= 32) consecutive segments per atomic ticket: the batch's headers arrive in ONE
// coalesced load (lane k holds segment k) and the next segment's first chunk is prefetched by run_segment, so the
// ticket -> header -> indices latency chain is paid once per batch instead of once per (short) row.
template <int VEC, int LPR, int KT, bool EXACT>
__device__ __forceinline__ void drain_queue_batched(const SegArgs<VEC, LPR, KT, EXACT> &A, const uint32_t *__restrict__ queue_begin,
                                            uint32_t *__restrict__ counters, int q, int lane, const uint32_t kBatch) {
  const uint32_t qb = __ldg(queue_begin + q), qn = __ldg(queue_begin + q + 1) - qb;
  uint32_t t = 0;
  if (lane == 0) t = atomicAdd(counters + q, kBatch);
  t = __shfl_sync(0xffffffffu, t, 0);
  while (t < qn) {
    const uint32_t nb = min(kBatch, qn - t);
    uint4 h = make_uint4(0, 0, 0, 0);
    if ((uint32_t)lane < nb) h = __ldg(A.segs + qb + t + lane);
    uint32_t tn = 0;
    if (lane == 0) tn = atomicAdd(counters + q, kBatch);  // next batch's ticket is in flight during this batch
    uint4 sg;
    sg.x = __shfl_sync(0xffffffffu, h.x, 0); sg.y = __shfl_sync(0xffffffffu, h.y, 0);
    sg.z = __shfl_sync(0xffffffffu, h.z, 0); sg.w = __shfl_sync(0xffffffffu, h.w, 0);
    uint32_t idx_n;
    float val_n;
    load_chunk(A, sg.y + lane, sg.z, idx_n, val_n);
    for (uint32_t k = 0; k < nb; k++) {
      const bool has_next = k + 1 < nb;
      uint4 sgn;
      sgn.x = __shfl_sync(0xffffffffu, h.x, (int)((k + 1) & 31)); sgn.y = __shfl_sync(0xffffffffu, h.y, (int)((k + 1) & 31));
      sgn.z = __shfl_sync(0xffffffffu, h.z, (int)((k + 1) & 31)); sgn.w = __shfl_sync(0xffffffffu, h.w, (int)((k + 1) & 31));
      run_segment<VEC, LPR, KT, EXACT>(A, sg, lane, idx_n, val_n, has_next, sgn);
      sg = sgn;
    }
    t = __shfl_sync(0xffffffffu, tn, 0);
  }
}

template <int VEC, int LPR, int KT, bool EXACT>
__device__ __forceinline__ void drain_queue(const SegArgs<VEC, LPR, KT, EXACT> &A, const uint32_t *__restrict__ queue_begin,
                                            uint32_t *__restrict__ counters, int q, int lane, int batch) {
  if (batch <= 1) drain_queue_lookahead<VEC, LPR, KT, EXACT>(A, queue_begin, counters, q, lane);
  else drain_queue_batched<VEC, LPR, KT, EXACT>(A, queue_begin, counters, q, lane, (uint32_t)batch);
}

template <int VEC, int LPR, int KT, bool EXACT>
__global__ void __launch_bounds__(kThreads, (KT == 1 ? kBlocksNarrow : 4))
spmm_seg_kernel(const uint4 *__restrict__ segs, const uint32_t *__restrict__ queue_begin,
                uint32_t *__restrict__ counters, int n_queues, const uint32_t *__restrict__ indices,
                const float *__restrict__ values, const uint32_t *__restrict__ perm, const float *__restrict__ B,
                float *__restrict__ C, float *__restrict__ scratch, int dim, int ldb, int ldc, int batch) {
  const SegArgs<VEC, LPR, KT, EXACT> A{segs, indices, values, perm, B, C, scratch, dim, ldb, ldc};
  const int lane = threadIdx.x & 31;
  if (batch == 0) {
    // small matrices (cora: 13 k entries): warp w runs segment w -- no tickets, no stealing, no counter reset.  The queue
    // machinery below costs ~15 us per launch whatever the size, nine SpMM launches per epoch.  Same run_segment, same bits.
    const uint32_t w = (uint32_t)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (w < (uint32_t)n_queues) {  // n_queues carries the segment count in this mode
      const uint4 sg = __ldg(segs + w);
      uint32_t idx_n;
      float val_n;
      load_chunk(A, sg.y + lane, sg.z, idx_n, val_n);
      run_segment<VEC, LPR, KT, EXACT>(A, sg, lane, idx_n, val_n, false, make_uint4(0, 0, 0, 0));
    }
    return;
  }
  uint32_t smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const int q0 = (int)(smid % (uint32_t)n_queues);
  // 1) this SM's own queue (contiguous rows => L1 reuse of gathered neighbour rows)
  drain_queue<VEC, LPR, KT, EXACT>(A, queue_begin, counters, q0, lane, batch);
  // 2) steal: look at ALL other queues at once -- every lane probes up to 8 of them with independent loads (one L2 latency
  // instead of one per 32 queues: on cora-sized graphs the serial probe loop was most of the kernel) -- then drain the ones
  // that still hold segments
  for (int base = 1; base < n_queues; base += 256) {
    uint32_t mine = 0;  // bit k: queue (q0 + base + 32 k + lane) % n_queues still holds segments
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int off = base + k * 32 + lane;
      if (off < n_queues) {
        const int q = (q0 + off) % n_queues;
        const uint32_t taken = *((volatile uint32_t *)(counters + q));
        if (taken < __ldg(queue_begin + q + 1) - __ldg(queue_begin + q)) mine |= 1u << k;
      }
    }
#pragma unroll 1
    for (int k = 0; k < 8 && base + k * 32 < n_queues; k++) {
      uint32_t mask = __ballot_sync(0xffffffffu, (mine >> k) & 1u);
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        drain_queue<VEC, LPR, KT, EXACT>(A, queue_begin, counters, (q0 + base + k * 32 + src) % n_queues, lane, batch);
      }
    }
  }
  // the last CTA to finish leaves the tickets zeroed for the next launch (no memset per launch: on small graphs a
  // memset node costs as much as the product itself).  Every other CTA has left its work loops when it counts itself
  // done, so nobody reads the tickets any more.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counters + n_queues, 1u) == gridDim.x - 1) {
      for (int q = 0; q <= n_queues; q++) counters[q] = 0u;
      __threadfence();
    }
  }
}

// out[i] = value of the first entry (i, i + col0) of row i, 0 if the row has none (the scales of a bit-tile plan are the
// square roots of the diagonal of the normalised adjacency)
__global__ void csr_diagonal_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                                    const float *__restrict__ values, int64_t n_rows, int64_t col0, float *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
    float d = 0.f;
    for (uint32_t e = indptr[i]; e < indptr[i + 1]; e++)
      if ((int64_t)indices[e] == i + col0) {
        d = values[e];
        break;
      }
    out[i] = d;
  }
}

// rows cut into several segments: out[row] = partial[s0] + partial[s0+1] + ... in ascending order
__global__ void spmm_combine_kernel(const uint32_t *__restrict__ split_row, const uint32_t *__restrict__ split_slot,
                                    const float *__restrict__ scratch, float *__restrict__ C, int64_t n_split,
                                    int dim, int ldc) {
  const int64_t total = n_split * dim;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int c = (int)(i - r * dim);
    const uint32_t s0 = __ldg(split_slot + r), s1 = __ldg(split_slot + r + 1);
    float sum = 0.f;
    for (uint32_t s = s0; s < s1; s++) sum += __ldg(scratch + (size_t)s * dim + c);
    C[(size_t)__ldg(split_row + r) * ldc + c] = sum;
  }
}

// all-columns-present check of a feature CSR (Reddit: every row holds columns 0..n_cols-1), done on the device so
// that 4*nnz bytes of indices are not pulled back over PCIe just to be scanned
__global__ void dense_check_kernel(const uint32_t *__restrict__ indices, int64_t nnz, uint32_t n_cols,
                                   unsigned int *__restrict__ mismatch) {
  bool bad = false;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
    bad |= __ldg(indices + e) != (uint32_t)(e % n_cols);
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(mismatch, 1u);
}

using KernelFn = void (*)(const uint4 *, const uint32_t *, uint32_t *, int, const uint32_t *, const float *,
                          const uint32_t *, const float *, float *, float *, int, int, int, int);

template <int VEC, int LPR, int KT>
KernelFn kfn(int dim) {
  if (dim == VEC * LPR * KT) return spmm_seg_kernel<VEC, LPR, KT, true>;
  return spmm_seg_kernel<VEC, LPR, KT, false>;
}

template <int VEC>
KernelFn pick_kernel(int dim) {
  const int chunks = (dim + VEC - 1) / VEC;  // lanes needed across the feature dimension
  if (chunks <= 1) return kfn<VEC, 1, 1>(dim);
  if (chunks <= 2) return kfn<VEC, 2, 1>(dim);
  if (chunks <= 4) return kfn<VEC, 4, 1>(dim);
  if (chunks <= 8) return kfn<VEC, 8, 1>(dim);
  if (chunks <= 16) return kfn<VEC, 16, 1>(dim);
  if (chunks <= 32) return kfn<VEC, 32, 1>(dim);
  if (chunks <= 64) return kfn<VEC, 32, 2>(dim);
  if (chunks <= 128) return kfn<VEC, 32, 4>(dim);
  if (chunks <= 256) return kfn<VEC, 32, 8>(dim);
  return nullptr;
}

}  // namespace

extern "C" {

int gcnb_spmm_plan_create(const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows, int64_t n_cols,
                          int seg_nnz, gcnb_stream_t stream_, gcnb_spmm_plan **out) {
  if (!d_indptr || !out || n_rows < 0) return GCNB_E_BADARG;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  cudaStream_t stream = as_stream(stream_);
  if (seg_nnz <= 0) seg_nnz = 1024;
  std::vector<uint32_t> indptr((size_t)n_rows + 1);
  GCNB_CHECK(cudaMemcpyAsync(indptr.data(), d_indptr, indptr.size() * 4, cudaMemcpyDeviceToHost, stream));
  GCNB_CHECK(cudaStreamSynchronize(stream));

  auto *p = new gcnb_spmm_plan();
  p->d_indptr = d_indptr;
  p->d_indices = d_indices;
  p->n_rows = n_rows;
  p->n_cols = n_cols;
  p->nnz = n_rows ? indptr[n_rows] : 0;
  p->seg_nnz = seg_nnz;

  std::vector<uint4> segs;
  segs.reserve((size_t)n_rows + (size_t)(p->nnz / seg_nnz) + 1);
  std::vector<uint32_t> split_row, split_slot;
  uint32_t slots = 0;
  for (int64_t r = 0; r < n_rows; r++) {
    const uint32_t b = indptr[r], e = indptr[r + 1];
    const uint32_t deg = e - b;
    p->max_deg = std::max<int64_t>(p->max_deg, deg);
    if (deg <= (uint32_t)seg_nnz) {
      segs.push_back(make_uint4((uint32_t)r, b, e, kNoSlot));
    } else {
      // equal pieces (not seg_nnz + remainder) so that the last piece is not a straggler
      const uint32_t pieces = (deg + seg_nnz - 1) / seg_nnz;
      split_row.push_back((uint32_t)r);
      split_slot.push_back(slots);
      for (uint32_t k = 0; k < pieces; k++) {
        const uint32_t sb = b + (uint32_t)((uint64_t)deg * k / pieces);
        const uint32_t se = b + (uint32_t)((uint64_t)deg * (k + 1) / pieces);
        segs.push_back(make_uint4((uint32_t)r, sb, se, slots++));
      }
    }
  }
  split_slot.push_back(slots);
  p->n_seg = (int64_t)segs.size();
  // short rows: several segments per atomic ticket (see drain_queue_batched); long rows: one, with look-ahead
  p->batch = (int)std::min<int64_t>(8, std::max<int64_t>(1, 384 / std::max<int64_t>(1, p->nnz / std::max<int64_t>(1, p->n_seg))));
  if (const char *e = getenv("GCNB_SPMM_BATCH")) p->batch = std::max(1, std::min(32, atoi(e)));  // tuning probe
  p->n_split_rows = (int64_t)split_row.size();
  p->n_slots = slots;

  // per-SM queues: contiguous segments, equal cost (nnz + fixed per-segment overhead)
  p->n_queues = std::max(1, di.sm_count);
  std::vector<uint32_t> qbeg((size_t)p->n_queues + 1, 0);
  {
    const double seg_overhead = 24.0;
    double total = 0;
    for (auto &s : segs) total += (s.z - s.y) + seg_overhead;
    double acc = 0;
    int q = 1;
    for (size_t i = 0; i < segs.size() && q < p->n_queues; i++) {
      acc += (segs[i].z - segs[i].y) + seg_overhead;
      while (q < p->n_queues && acc >= total * q / p->n_queues) qbeg[q++] = (uint32_t)(i + 1);
    }
    for (; q <= p->n_queues; q++) qbeg[q] = (uint32_t)segs.size();
    qbeg[p->n_queues] = (uint32_t)segs.size();
  }

  auto upload = [&](void **dst, const void *src, size_t bytes) -> int {
    if (bytes == 0) bytes = 4;
    GCNB_CHECK(cudaMalloc(dst, bytes));
    if (src) GCNB_CHECK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return 0;
  };
  int rc = 0;
  if (!rc) rc = upload((void **)&p->d_segs, segs.empty() ? nullptr : segs.data(), segs.size() * sizeof(uint4));
  if (!rc) rc = upload((void **)&p->d_queue_begin, qbeg.data(), qbeg.size() * 4);
  if (!rc) rc = upload((void **)&p->d_counters, nullptr, ((size_t)p->n_queues + 1) * 4);
  if (!rc) rc = (int)cudaMemsetAsync(p->d_counters, 0, ((size_t)p->n_queues + 1) * 4, stream);  // kernels re-zero them
  if (!rc) rc = upload((void **)&p->d_split_row, split_row.empty() ? nullptr : split_row.data(), split_row.size() * 4);
  if (!rc) rc = upload((void **)&p->d_split_slot, split_slot.data(), split_slot.size() * 4);
  if (!rc) rc = (int)cudaStreamSynchronize(stream);
  if (rc) {
    gcnb_spmm_plan_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}

int gcnb_csr_diagonal_f32(const uint32_t *d_indptr, const uint32_t *d_indices, const float *d_values, int64_t n_rows,
                          int64_t col0, float *d_out, gcnb_stream_t stream) {
  if (!d_indptr || !d_out || n_rows < 0 || (n_rows > 0 && (!d_indices || !d_values))) return GCNB_E_BADARG;
  if (n_rows == 0) return 0;
  const int blocks = (int)std::min<int64_t>((n_rows + 255) / 256, (int64_t)std::max(1, device_info().sm_count) * 8);
  csr_diagonal_kernel<<<blocks, 256, 0, as_stream(stream)>>>(d_indptr, d_indices, d_values, n_rows, col0, d_out);
  GCNB_LAUNCH_CHECK();
  return 0;
}

int gcnb_spmm_plan_destroy(gcnb_spmm_plan *p) {
  if (!p) return 0;
  cudaFree(p->d_segs);
  cudaFree(p->d_queue_begin);
  cudaFree(p->d_counters);
  cudaFree(p->d_split_row);
  cudaFree(p->d_split_slot);
  cudaFree(p->d_scratch);
  gcnb::stage_destroy(p->staged);
  delete p;
  return 0;
}

int gcnb_spmm_plan_info(const gcnb_spmm_plan *p, int64_t out[8]) {
  if (!p || !out) return GCNB_E_BADARG;
  out[0] = p->n_rows; out[1] = p->nnz; out[2] = p->n_seg; out[3] = p->n_split_rows;
  out[4] = p->n_slots; out[5] = p->n_queues; out[6] = p->seg_nnz; out[7] = p->max_deg;
  return 0;
}

int gcnb_spmm_plan_attach_bittile(gcnb_spmm_plan *p, gcnb_bittile_plan *bt, const float *d_values) {
  if (!p || (bt && !d_values)) return GCNB_E_BADARG;
  p->bittile = bt;
  p->bittile_values = bt ? d_values : nullptr;
  return 0;
}

int gcnb_spmm_f32(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, float *d_C,
                  int dim, gcnb_stream_t stream_) {
  return gcnb_spmm_ld_f32(p, d_values, d_perm, d_B, dim, d_C, dim, dim, stream_);
}

int gcnb_spmm_ld_f32(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, int64_t ldb_,
                     float *d_C, int64_t ldc_, int dim, gcnb_stream_t stream_) {
  if (!p || !d_values || !d_B || !d_C || dim <= 0 || ldb_ < dim || ldc_ < dim || ldb_ > (1 << 28) || ldc_ > (1 << 28))
    return GCNB_E_BADARG;
  if (p->n_rows == 0) return 0;
  const int ldb = (int)ldb_, ldc = (int)ldc_;
  cudaStream_t stream = as_stream(stream_);
  if (p->bittile && d_values == p->bittile_values && !d_perm) {  // tensor-core bit tiles + remainder CSR, spmm_bittile.cu
    if (dim == 16 && ldb == 16 && ldc == 16 && (((uintptr_t)d_B | (uintptr_t)d_C) % 16 == 0))
      return gcnb_bittile_spmm16_f32(p->bittile, d_B, d_C, stream_);
    // 16-column slabs for every wider operand (the last slab shifted left to end at dim): width 41 on the Reddit-shape graph
    // = 3 slabs of ~0.25 ms against 2.0 ms for one pass of the generic kernel
    if (dim > 16) return gcnb_bittile_spmm_ld_f32(p->bittile, d_B, ldb, d_C, ldc, dim, stream_);
  }
  if (p->staged) {  // window-staged fast path (static values, 16-column slabs), spmm_stage.cu
    int handled = 0;
    const int rc = gcnb_stage_try_spmm(p, d_values, d_perm, d_B, ldb, d_C, ldc, dim, stream, &handled);
    if (rc || handled) return rc;
  }
  // Column slabs: (1) the kernel holds at most 256 column chunks per row (1024 floats vectorised, 256 otherwise);
  // (2) a dense operand much larger than L2 is gathered from HBM at sector granularity, while a slab that fits in L2
  // is gathered on chip and re-streaming the index (8 bytes per entry per slab) is cheap next to row gathers of
  // 4 * width bytes per entry -- Reddit-shape GraphSum at width 600: 75 ms as one pass, ~4x less in 32-column slabs.
  const bool vec4 = (dim % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && (((uintptr_t)d_B | (uintptr_t)d_C) % 16 == 0);
  int width = std::min(dim, vec4 ? 1024 : 256);
  static const int64_t l2_budget = [] {
    const char *e = getenv("GCNB_SPMM_SLAB_BYTES");  // tuning probe; 0 disables the L2 slabs
    return e ? atoll(e) : (int64_t)32 << 20;
  }();
  if (l2_budget > 0 && dim > 32 && p->n_cols * (int64_t)dim * 4 > 2 * l2_budget) {
    int64_t w = l2_budget / (p->n_cols * 4);
    w = std::max<int64_t>(16, w / 16 * 16);
    width = (int)std::min<int64_t>(width, w);
  }
  for (int c0 = 0; c0 < dim; c0 += width) {
    const int w = std::min(width, dim - c0);
    const int rc = spmm_generic_launch(p, d_values, d_perm, d_B + c0, ldb, d_C + c0, ldc, w, stream);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"

int gcnb::spmm_generic_launch(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, int ldb,
                              float *d_C, int ldc, int dim, cudaStream_t stream) {
  if (p->n_rows == 0) return 0;
  const bool vec4 = (dim % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && (((uintptr_t)d_B | (uintptr_t)d_C) % 16 == 0);
  KernelFn fn = vec4 ? pick_kernel<4>(dim) : pick_kernel<1>(dim);
  if (!fn) return GCNB_E_UNSUPPORTED;
  if (p->n_slots > 0 && p->scratch_dim < dim) {
    // grows once per new (larger) feature width; steady-state launches never allocate
    GCNB_CHECK(cudaStreamSynchronize(stream));
    cudaFree(p->d_scratch);
    p->d_scratch = nullptr;
    GCNB_CHECK(cudaMalloc((void **)&p->d_scratch, (size_t)p->n_slots * dim * sizeof(float)));
    p->scratch_dim = dim;
  }
  if (p->n_slots > 0 && (uintptr_t)p->d_scratch % 16 != 0) return GCNB_E_UNSUPPORTED;
  int blocks_per_sm = 0;
  {
    static std::mutex mu;
    static std::unordered_map<const void *, int> cache;  // occupancy query costs microseconds: once per kernel
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find((const void *)fn);
    if (it == cache.end()) {
      GCNB_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, kThreads, 0));
      if (blocks_per_sm < 1) blocks_per_sm = 1;
      cache[(const void *)fn] = blocks_per_sm;
    } else {
      blocks_per_sm = it->second;
    }
  }
  if (p->n_seg <= kStaticMaxSegs && p->max_cta_per_sm == 0) {
    // launch-bound regime: one warp per segment, statically
    const int64_t grid = std::max<int64_t>(1, (p->n_seg + (kThreads / 32) - 1) / (kThreads / 32));
    fn<<<(unsigned)grid, kThreads, 0, stream>>>(p->d_segs, p->d_queue_begin, p->d_counters, (int)p->n_seg, p->d_indices, d_values,
                                                d_perm, d_B, d_C, p->d_scratch, dim, ldb, ldc, /*batch = static*/ 0);
    GCNB_LAUNCH_CHECK();
    if (p->n_split_rows > 0) {
      const int64_t total = p->n_split_rows * dim;
      const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->n_queues * 8);
      spmm_combine_kernel<<<blocks, 256, 0, stream>>>(p->d_split_row, p->d_split_slot, p->d_scratch, d_C, p->n_split_rows, dim,
                                                      ldc);
      GCNB_LAUNCH_CHECK();
    }
    return 0;
  }
  if (p->max_cta_per_sm > 0) blocks_per_sm = std::min(blocks_per_sm, p->max_cta_per_sm);
  const int64_t warps_needed = p->n_seg;
  int64_t grid = (int64_t)p->n_queues * blocks_per_sm;
  const int64_t min_grid = (warps_needed + (kThreads / 32) - 1) / (kThreads / 32);
  if (grid > min_grid) grid = std::max<int64_t>(1, min_grid);
  fn<<<(unsigned)grid, kThreads, 0, stream>>>(p->d_segs, p->d_queue_begin, p->d_counters, p->n_queues, p->d_indices,
                                              d_values, d_perm, d_B, d_C, p->d_scratch, dim, ldb, ldc, p->batch);
  GCNB_LAUNCH_CHECK();
  if (p->n_split_rows > 0) {
    const int64_t total = p->n_split_rows * dim;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)p->n_queues * 8);
    spmm_combine_kernel<<<blocks, 256, 0, stream>>>(p->d_split_row, p->d_split_slot, p->d_scratch, d_C,
                                                    p->n_split_rows, dim, ldc);
    GCNB_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" {

// ---- CSC companion -----------------------------------------------------------------------------------
struct gcnb_csc {
  uint32_t *d_colptr = nullptr, *d_rowidx = nullptr, *d_perm = nullptr;
  int is_dense = 0;
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
};

static int csc_build(gcnb_csc *c, const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows, int64_t n_cols,
                     cudaStream_t stream) {
  std::vector<uint32_t> indptr((size_t)n_rows + 1);
  GCNB_CHECK(cudaMemcpyAsync(indptr.data(), d_indptr, indptr.size() * 4, cudaMemcpyDeviceToHost, stream));
  GCNB_CHECK(cudaStreamSynchronize(stream));
  const int64_t nnz = n_rows ? indptr[n_rows] : 0;
  for (int64_t r = 0; r < n_rows; r++)
    if (indptr[r + 1] < indptr[r]) return GCNB_E_BADARG;  // not a CSR offset array
  c->n_rows = n_rows;
  c->n_cols = n_cols;
  c->nnz = nnz;
  // dense detection needs only indptr first: every row must hold exactly n_cols entries
  bool dense = n_cols > 0 && nnz == n_rows * n_cols;
  for (int64_t r = 0; dense && r < n_rows; r++) dense = (indptr[r + 1] - indptr[r]) == (uint32_t)n_cols;
  if (dense && nnz) {
    unsigned int *d_flag = nullptr, h_flag = 0;
    GCNB_CHECK(cudaMalloc((void **)&d_flag, 4));
    int rc = (int)cudaMemsetAsync(d_flag, 0, 4, stream);
    if (!rc) {
      const int blocks = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)std::max(1, device_info().sm_count) * 16);
      dense_check_kernel<<<blocks, 256, 0, stream>>>(d_indices, nnz, (uint32_t)n_cols, d_flag);
      rc = (int)cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, stream);
    }
    if (!rc) rc = (int)cudaStreamSynchronize(stream);
    cudaFree(d_flag);
    if (rc) return rc;
    dense = (h_flag == 0);
  }
  c->is_dense = dense ? 1 : 0;
  if (!dense) {
    std::vector<uint32_t> indices((size_t)nnz);
    if (nnz) {
      GCNB_CHECK(cudaMemcpyAsync(indices.data(), d_indices, (size_t)nnz * 4, cudaMemcpyDeviceToHost, stream));
      GCNB_CHECK(cudaStreamSynchronize(stream));
    }
    // a column index beyond n_cols (a negative svmlight index cast to uint32, a wrong input_dim) would overrun the
    // counting sort below: an error, not a heap overflow
    for (int64_t e = 0; e < nnz; e++)
      if (indices[e] >= (uint64_t)n_cols) return GCNB_E_BADARG;
    // stable counting sort by column: entries of a column stay in ascending row order => fixed summation order
    std::vector<uint32_t> colptr((size_t)n_cols + 1, 0), rowidx((size_t)nnz), perm((size_t)nnz);
    for (int64_t e = 0; e < nnz; e++) colptr[indices[e] + 1]++;
    for (int64_t j = 0; j < n_cols; j++) colptr[j + 1] += colptr[j];
    std::vector<uint32_t> cur(colptr.begin(), colptr.end() - 1);
    for (int64_t r = 0; r < n_rows; r++)
      for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
        const uint32_t pos = cur[indices[e]]++;
        rowidx[pos] = (uint32_t)r;
        perm[pos] = e;
      }
    GCNB_CHECK(cudaMalloc((void **)&c->d_colptr, colptr.size() * 4));
    GCNB_CHECK(cudaMalloc((void **)&c->d_rowidx, std::max<size_t>(4, rowidx.size() * 4)));
    GCNB_CHECK(cudaMalloc((void **)&c->d_perm, std::max<size_t>(4, perm.size() * 4)));
    GCNB_CHECK(cudaMemcpyAsync(c->d_colptr, colptr.data(), colptr.size() * 4, cudaMemcpyHostToDevice, stream));
    if (nnz) {
      GCNB_CHECK(cudaMemcpyAsync(c->d_rowidx, rowidx.data(), rowidx.size() * 4, cudaMemcpyHostToDevice, stream));
      GCNB_CHECK(cudaMemcpyAsync(c->d_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice, stream));
    }
    GCNB_CHECK(cudaStreamSynchronize(stream));
  }
  return 0;
}

int gcnb_csc_create(const uint32_t *d_indptr, const uint32_t *d_indices, int64_t n_rows, int64_t n_cols,
                    gcnb_stream_t stream_, gcnb_csc **out) {
  if (!d_indptr || !out || n_rows < 0 || n_cols < 0) return GCNB_E_BADARG;
  *out = nullptr;
  auto *c = new gcnb_csc();
  const int rc = csc_build(c, d_indptr, d_indices, n_rows, n_cols, as_stream(stream_));
  if (rc) {  // nothing leaks on an error path
    gcnb_csc_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

int gcnb_csc_destroy(gcnb_csc *c) {
  if (!c) return 0;
  cudaFree(c->d_colptr);
  cudaFree(c->d_rowidx);
  cudaFree(c->d_perm);
  delete c;
  return 0;
}

int gcnb_csc_arrays(const gcnb_csc *c, const uint32_t **d_colptr, const uint32_t **d_rowidx, const uint32_t **d_perm,
                    int *is_dense) {
  if (!c) return GCNB_E_BADARG;
  if (d_colptr) *d_colptr = c->d_colptr;
  if (d_rowidx) *d_rowidx = c->d_rowidx;
  if (d_perm) *d_perm = c->d_perm;
  if (is_dense) *is_dense = c->is_dense;
  return 0;
}

// ---- CUDA-graph replay support -------------------------------------------------------------------------------------------
int gcnb_graph_patch_node(void *exec_, void *node_, const gcnb_rng_t *rng, const float *step_size) {
  if (!exec_ || !node_ || (!rng && !step_size)) return GCNB_E_BADARG;
  cudaGraphExec_t exec = (cudaGraphExec_t)exec_;
  cudaGraphNode_t node = (cudaGraphNode_t)node_;
  cudaKernelNodeParams kp{};
  GCNB_CHECK(cudaGraphKernelNodeGetParams(node, &kp));
  int rng_arg = -1, step_arg = -1;
  if (!gcnb::lookup_patchable(kp.func, &rng_arg, &step_arg)) return GCNB_E_UNSUPPORTED;
  const int n_args = std::max(rng_arg, step_arg) + 1;
  if ((rng && rng_arg < 0) || (step_size && step_arg < 0) || !kp.kernelParams || n_args > 16) return GCNB_E_BADARG;
  void *args[16];
  for (int i = 0; i < n_args; i++) args[i] = kp.kernelParams[i];
  gcnb_rng_t r;
  float st;
  if (rng) {
    r = *rng;
    args[rng_arg] = &r;
  }
  if (step_size) {
    st = *step_size;
    args[step_arg] = &st;
  }
  // every patchable kernel keeps the patched argument LAST, so the first n_args pointers are its complete argument list
  kp.kernelParams = args;
  GCNB_CHECK(cudaGraphExecKernelNodeSetParams(exec, node, &kp));
  return 0;
}

const char *gcnb_error_string(int code) {
  if (code == 0) return "success";
  if (code == GCNB_E_BADARG) return "gcnb: bad argument";
  if (code == GCNB_E_UNSUPPORTED) return "gcnb: unsupported shape/alignment";
  if (code == GCNB_E_COMM) return "gcnb: NCCL unavailable or collective failed";
  return cudaGetErrorString((cudaError_t)code);
}
int gcnb_version(void) { return 100; }
int gcnb_set_host_threads(int n) {
  gcnb::g_host_threads = n > 0 ? n : 0;
  return 0;
}
int gcnb_host_threads(void) { return gcnb::host_threads(); }
int gcnb_device_check(int *sm_count) {
  int n = 0;
  GCNB_CHECK(cudaGetDeviceCount(&n));
  if (n == 0) return (int)cudaErrorNoDevice;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  if (sm_count) *sm_count = di.sm_count;
  if (di.cc_major != 10) return (int)cudaErrorInvalidDevice;
  return 0;
}

}  // extern "C"
