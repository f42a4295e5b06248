// spmm_ell.cu -- pattern-only row gather at width 16:  R[i][:] = row_scale[i] * sum over the entries (i, j) of B2[j][:]
//
// The REMAINDER of a bit-tile GraphSum plan (spmm_bittile.cu): the entries outside the dense blocks, ~26 % of the
// Reddit-shape bench graph, mostly its uniformly random inter-community edges.  Every one of them costs one random
// 64-byte row of an L2-resident matrix; nothing can be shared or staged, so the kernel's only job is to keep
// as many of those row gathers in flight as the LSU pipe and L2 take, with nothing else on that pipe.
//
// GraphSum's values factor (graph_value[i,j] = s_i * s_j, src/parser.cpp:164-181), so with B2 = diag(s) * B written once
// per launch by the pack kernel the entries carry NO value: 4 bytes of index per entry instead of 8, adds instead of
// FMAs.  (A plan with entries that do not factor keeps the generic valued kernel, spmm.cu.)
//
// Layout (host-built once, gcnb::ell_build_host): rows are sorted by length and dealt 8 at a time into BUNDLES; a warp
// processes a bundle with its 8 lane groups (4 lanes x float4 = one 64-byte row) each owning ONE row: private sums, no
// shuffle, no cross-lane reduction -- the generic kernel spends 0.25 LSU wavefronts per entry on index / value
// broadcasts and a 12-shuffle tree per row on the same pipe as the gathers.  Indices are stored [bundle][step / 4][group]
// as uint4: the four lanes of a group load the same 16 bytes (one LDG.128 per warp = 128 contiguous bytes = the next four
// gather steps of all 8 rows).  Rows are padded to the bundle's length with the index of an all-zero row (n_cols).
// Rows longer than kEllWideMin entries (64 when the plan has few rows, see ell_layout) become WIDE bundles: the 8 groups share the row (entry e -> group e % 8) and
// finish with a 3-step shuffle tree; rows longer than kEllWideMax are cut into parts whose partial sums go to slots
// that ell_combine_kernel adds in ascending order.  Bundles are claimed longest first through one atomic ticket
// (two tickets and one header ahead).  Fixed summation order => bit-reproducible.
//
// Reference being replaced: graphsum_kernel, src/module.cu:172-186 (the part of it the tensor-core tiles do not cover).
#include <algorithm>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "common.cuh"
#include "spmm_plan.cuh"

using namespace gcnb;

namespace gcnb {

constexpr uint32_t kEllNone = 0xffffffffu;

// One wide bundle (a row of more than kEllWideMin entries, or one part of a row of more than kEllWideMax): `beg` is the
// part's first entry RELATIVE to the row's first entry
struct EllWide {
  uint32_t row, beg, len, slot;
};

// Bundle layout from the rows' lengths alone: wide parts first (longest first), then the short rows sorted by length (longest
// first, ties by row id), 8 per bundle.  Fills H.off / steps / rows / split_* / n_slots / n_bundles; returns the number of
// uint4-rows of H.idx in *n_idx_rows.  Shared by the host builder and the device builder (ell_build_device).
static int ell_layout(const uint32_t *len_of_row, int64_t n_rows, EllHost &H, std::vector<EllWide> &wide,
                      std::vector<uint32_t> &narrow, uint64_t *n_idx_rows) {
  wide.clear();
  H.split_row.clear();
  H.split_ptr.assign(1, 0u);
  uint32_t n_slots = 0;
  // narrow rows by counting sort (descending length, ascending row id inside a length == a stable sort by length)
  // Few rows (a rank's block of a partitioned graph, a small graph): 8 rows per warp leave most of the machine idle and the
  // longest row of a bundle is one lane group's serial chain (B200, one of 8 row blocks of the bench graph: 75 us for
  // 3.8 M entries, 4x the full-size rate).  There the 8 groups share every row of more than 64 entries.
  const uint32_t wide_min = (uint32_t)(n_rows < kEllFewRows ? kEllWideMinFewRows : kEllWideMin);
  H.wide_min = (int)wide_min;
  std::vector<uint32_t> bucket((size_t)wide_min + 2, 0u);
  for (int64_t i = 0; i < n_rows; i++) {
    const uint32_t len = len_of_row[i];
    if (len <= wide_min) {
      bucket[(size_t)wide_min - len + 1]++;
    } else if (len <= (uint32_t)kEllWideMax) {
      wide.push_back(EllWide{(uint32_t)i, 0u, len, kEllNone});
    } else {
      const uint32_t parts = (len + kEllWideMax - 1) / kEllWideMax;
      for (uint32_t p = 0; p < parts; p++) {
        const uint32_t b = (uint32_t)((uint64_t)len * p / parts), e = (uint32_t)((uint64_t)len * (p + 1) / parts);
        wide.push_back(EllWide{(uint32_t)i, b, e - b, n_slots++});
      }
      H.split_row.push_back((uint32_t)i);
      H.split_ptr.push_back(n_slots);
    }
  }
  for (size_t k = 1; k < bucket.size(); k++) bucket[k] += bucket[k - 1];
  narrow.assign((size_t)bucket.back(), 0u);
  for (int64_t i = 0; i < n_rows; i++) {
    const uint32_t len = len_of_row[i];
    if (len <= wide_min) narrow[bucket[(size_t)wide_min - len]++] = (uint32_t)i;
  }
  H.n_slots = n_slots;
  std::stable_sort(wide.begin(), wide.end(), [](const EllWide &a, const EllWide &b) { return a.len > b.len; });
  const size_t n_narrow_b = (narrow.size() + 7) / 8;
  const size_t n_b = wide.size() + n_narrow_b;
  if (n_b > 0x7ffffff0ull) return GCNB_E_BADARG;
  H.off.assign(n_b + 1, 0u);
  H.steps.assign(n_b, 0u);
  H.rows.assign(n_b * 8, kEllNone);
  uint64_t acc = 0;
  for (size_t b = 0; b < n_b; b++) {
    uint32_t s4;
    if (b < wide.size()) {
      const EllWide &w = wide[b];
      s4 = ((w.len + 7) / 8 + 3) / 4;
      H.steps[b] = s4 | 0x80000000u;
      H.rows[b * 8 + 0] = w.row;
      H.rows[b * 8 + 1] = w.slot;
    } else {
      const size_t k0 = (b - wide.size()) * 8;
      s4 = (len_of_row[narrow[k0]] + 3) / 4;  // the bundle's first row is its longest
      H.steps[b] = s4;
      for (size_t g = 0; g < 8 && k0 + g < narrow.size(); g++) H.rows[b * 8 + g] = narrow[k0 + g];
    }
    H.off[b] = (uint32_t)acc;
    acc += s4;
    if (acc > 0xfffffff0ull) return GCNB_E_BADARG;
  }
  H.off[n_b] = (uint32_t)acc;
  H.n_bundles = (int64_t)n_b;
  *n_idx_rows = acc;
  return 0;
}

int ell_build_host(const uint32_t *indptr, const uint32_t *indices, int64_t n_rows, int64_t n_cols, int n_threads,
                   EllHost &H) {
  if (!indptr || n_rows < 0 || n_cols < 0 || n_rows > 0xfffffff0ll || n_cols > 0xfffffff0ll) return GCNB_E_BADARG;
  H.n_rows = n_rows;
  H.n_cols = n_cols;
  H.nnz = indptr[n_rows];
  if (H.nnz > 0 && !indices) return GCNB_E_BADARG;
  std::vector<uint32_t> len_of_row((size_t)n_rows);
  for (int64_t i = 0; i < n_rows; i++) len_of_row[(size_t)i] = indptr[i + 1] - indptr[i];
  std::vector<EllWide> wide;
  std::vector<uint32_t> narrow;
  uint64_t acc = 0;
  const int lrc = ell_layout(len_of_row.data(), n_rows, H, wide, narrow, &acc);
  if (lrc) return lrc;
  const size_t n_b = (size_t)H.n_bundles;
  H.idx.alloc((size_t)acc * 32);
  // ---- fill: idx[((off + k / 4) * 8 + g) * 4 + k % 4] = column of entry k of group g, n_cols (the zero row) as padding
  int T = n_threads > 0 ? n_threads : host_threads();
  T = (int)std::max<size_t>(1, std::min<size_t>((size_t)T, n_b / 64 + 1));
  const uint32_t pad = (uint32_t)n_cols;
  uint32_t *out = H.idx.data();
  auto fill = [&](int t) {
    for (size_t b = (size_t)t; b < n_b; b += (size_t)T) {
      const uint32_t s4 = H.steps[b] & 0x7fffffffu;
      uint32_t *base = out + (size_t)H.off[b] * 32;
      if (H.steps[b] & 0x80000000u) {
        const EllWide &w = wide[b];
        const uint32_t cap = s4 * 32, beg = indptr[w.row] + w.beg;
        for (uint32_t e = 0; e < cap; e++) {
          const uint32_t g = e & 7u, k = e >> 3;
          base[((size_t)(k >> 2) * 8 + g) * 4 + (k & 3u)] = e < w.len ? indices[beg + e] : pad;
        }
      } else {
        for (uint32_t g = 0; g < 8; g++) {
          const uint32_t r = H.rows[b * 8 + g];
          const uint32_t beg = r == kEllNone ? 0u : indptr[r], len = r == kEllNone ? 0u : indptr[r + 1] - beg;
          for (uint32_t k = 0; k < s4 * 4; k++)
            base[((size_t)(k >> 2) * 8 + g) * 4 + (k & 3u)] = k < len ? indices[beg + k] : pad;
        }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(fill, t);
  fill(0);
  for (auto &x : th) x.join();
  return 0;
}

namespace {

__device__ __forceinline__ uint4 ell_ld_idx(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ell_row(const float *__restrict__ B2l, uint32_t j) {
  return __ldg(reinterpret_cast<const float4 *>(B2l + (size_t)j * 16));
}
__device__ __forceinline__ void ell_add(float4 &a, const float4 &x) {
  a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
}

// out[0..4) = v, or out += v with one vector reduction (REDG.ADD.F32x4): the bit-tile plan lets the MMA kernel's epilogue and
// this kernel add their halves of a row into a zeroed C in whichever order they finish -- 0 + a + b == 0 + b + a in
// floating point, so the result does not depend on the order
__device__ __forceinline__ void ell_out(float *dst, const float4 v, const bool accumulate) {
  if (accumulate)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  else
    *reinterpret_cast<float4 *>(dst) = v;
}

struct EllHeader {
  uint32_t off, steps, row, slot;
};
__device__ __forceinline__ EllHeader ell_header(const uint32_t *__restrict__ off, const uint32_t *__restrict__ steps,
                                                const uint32_t *__restrict__ rows, uint32_t t, int g) {
  EllHeader h;
  h.off = __ldg(off + t);
  h.steps = __ldg(steps + t);
  const bool wide = (h.steps & 0x80000000u) != 0u;
  h.row = __ldg(rows + (size_t)t * 8 + (wide ? 0 : g));
  h.slot = wide ? __ldg(rows + (size_t)t * 8 + 1) : kEllNone;
  return h;
}

// persistent; one bundle per warp at a time.  counter[0] = ticket, counter[1] = CTAs done (the last one re-zeroes both)
__global__ void __launch_bounds__(256, 4)
ell_gather16_kernel(const uint4 *__restrict__ idx4, const uint32_t *__restrict__ off, const uint32_t *__restrict__ steps,
                    const uint32_t *__restrict__ rows, const float *__restrict__ B2, const float *__restrict__ row_scale,
                    float *__restrict__ R, int64_t ldr, int accumulate, const uint32_t *__restrict__ row_map,
                    float *__restrict__ slots, uint32_t *__restrict__ counter, uint32_t n_bundles, uint32_t static_stride) {
  const int lane = threadIdx.x & 31, g = lane >> 2, l = lane & 3;
  const float *B2l = B2 + l * 4;
  // static_stride > 0 (plans with few rows: many short bundles): warp w takes bundles w, w + stride, ... -- the ticket
  // counter is ONE address, and tens of thousands of atomics on it cost more than the short bundles they hand out
  uint32_t t = 0, tn = 0;
  if (static_stride) {
    t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    tn = t + static_stride;
  } else {
    if (lane == 0) {
      t = atomicAdd(counter, 1u);
      tn = atomicAdd(counter, 1u);
    }
    t = __shfl_sync(0xffffffffu, t, 0);
    tn = __shfl_sync(0xffffffffu, tn, 0);
  }
  EllHeader h{0, 0, kEllNone, kEllNone}, hn{0, 0, kEllNone, kEllNone};
  if (t < n_bundles) h = ell_header(off, steps, rows, t, g);
  while (t < n_bundles) {
    uint32_t tnn = tn + static_stride;
    if (!static_stride && lane == 0) tnn = atomicAdd(counter, 1u);     // two tickets ahead
    if (tn < n_bundles) hn = ell_header(off, steps, rows, tn, g);      // one header ahead
    const uint32_t n4 = h.steps & 0x7fffffffu;
    const bool wide = (h.steps & 0x80000000u) != 0u;
    const uint4 *p = idx4 + (size_t)h.off * 8 + g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t s = 0;
    if (n4 >= 2) {
      uint4 a = ell_ld_idx(p), b = ell_ld_idx(p + 8);
      for (; s + 4 <= n4; s += 2) {  // the indices of the next 8 steps are in flight while these 8 rows are gathered
        const uint4 an = ell_ld_idx(p + (size_t)(s + 2) * 8), bn = ell_ld_idx(p + (size_t)(s + 3) * 8);
        const float4 x0 = ell_row(B2l, a.x), x1 = ell_row(B2l, a.y), x2 = ell_row(B2l, a.z), x3 = ell_row(B2l, a.w);
        const float4 x4 = ell_row(B2l, b.x), x5 = ell_row(B2l, b.y), x6 = ell_row(B2l, b.z), x7 = ell_row(B2l, b.w);
        ell_add(acc, x0); ell_add(acc, x1); ell_add(acc, x2); ell_add(acc, x3);
        ell_add(acc, x4); ell_add(acc, x5); ell_add(acc, x6); ell_add(acc, x7);
        a = an;
        b = bn;
      }
      {
        const float4 x0 = ell_row(B2l, a.x), x1 = ell_row(B2l, a.y), x2 = ell_row(B2l, a.z), x3 = ell_row(B2l, a.w);
        const float4 x4 = ell_row(B2l, b.x), x5 = ell_row(B2l, b.y), x6 = ell_row(B2l, b.z), x7 = ell_row(B2l, b.w);
        ell_add(acc, x0); ell_add(acc, x1); ell_add(acc, x2); ell_add(acc, x3);
        ell_add(acc, x4); ell_add(acc, x5); ell_add(acc, x6); ell_add(acc, x7);
        s += 2;
      }
    }
    if (s < n4) {  // odd count: one more block of four steps
      const uint4 a = ell_ld_idx(p + (size_t)s * 8);
      const float4 x0 = ell_row(B2l, a.x), x1 = ell_row(B2l, a.y), x2 = ell_row(B2l, a.z), x3 = ell_row(B2l, a.w);
      ell_add(acc, x0); ell_add(acc, x1); ell_add(acc, x2); ell_add(acc, x3);
    }
    if (wide) {
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
      }
      if (g == 0) {
        if (h.slot == kEllNone) {
          const float sc = __ldg(row_scale + h.row);
          const size_t ro = row_map ? __ldg(row_map + h.row) : h.row;  // renumbered plan: its row k is the caller's row_map[k]
          ell_out(R + ro * ldr + l * 4, make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc), accumulate != 0);
        } else {
          *reinterpret_cast<float4 *>(slots + (size_t)h.slot * 16 + l * 4) = acc;
        }
      }
    } else if (h.row != kEllNone) {
      const float sc = __ldg(row_scale + h.row);
      const size_t ro = row_map ? __ldg(row_map + h.row) : h.row;
      ell_out(R + ro * ldr + l * 4, make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc), accumulate != 0);
    }
    t = tn;
    h = hn;
    tn = static_stride ? tnn : __shfl_sync(0xffffffffu, tnn, 0);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(counter + 1, 1u) == gridDim.x - 1) {
      counter[0] = 0u;
      counter[1] = 0u;
      __threadfence();
    }
  }
}

// rows cut into parts: R[row] = row_scale[row] * (slot[s0] + slot[s0 + 1] + ...), ascending; 4 lanes per row
__global__ void __launch_bounds__(256) ell_combine_kernel(const uint32_t *__restrict__ split_row, const uint32_t *__restrict__ split_ptr,
                                                          const float *__restrict__ slots, const float *__restrict__ row_scale,
                                                          float *__restrict__ R, int64_t ldr, int accumulate,
                                                          const uint32_t *__restrict__ row_map, int64_t n_split) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t k = tid >> 2;
  const int l = (int)(tid & 3);
  if (k >= n_split) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t s = split_ptr[k]; s < split_ptr[k + 1]; s++) ell_add(acc, *reinterpret_cast<const float4 *>(slots + (size_t)s * 16 + l * 4));
  const uint32_t row = split_row[k];
  const float sc = row_scale[row];
  const size_t ro = row_map ? row_map[row] : row;
  ell_out(R + ro * ldr + l * 4, make_float4(acc.x * sc, acc.y * sc, acc.z * sc, acc.w * sc), accumulate != 0);
}

// Device-side fill of the index array (ell_build_device): one warp per bundle writes its uint4-rows; output word o of a
// bundle holds step k = (o / 32) * 4 + o % 4 of lane group g = (o % 32) / 4 -- a lane keeps its group for the whole bundle.
// src[row_beg[r] + k] = k-th entry of row r (k < row_len[r]); wide[b] = (first entry relative to the row, entries) of wide
// bundle b < n_wide (its row is rows[b * 8])
__global__ void __launch_bounds__(256) ell_fill_kernel(const uint32_t *__restrict__ off, const uint32_t *__restrict__ steps,
                                                       const uint32_t *__restrict__ rows, const uint2 *__restrict__ wide,
                                                       const uint32_t *__restrict__ row_beg, const uint32_t *__restrict__ row_len,
                                                       const uint32_t *__restrict__ src, uint32_t pad, uint32_t *__restrict__ idx,
                                                       uint32_t n_bundles) {
  const int lane = threadIdx.x & 31, g = lane >> 2, kk = lane & 3;
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_bundles; b += n_warps) {
    const uint32_t st = steps[b], s4 = st & 0x7fffffffu;
    uint32_t *base = idx + (size_t)off[b] * 32;
    if (st & 0x80000000u) {
      const uint2 w = wide[b];
      const uint32_t beg = row_beg[rows[(size_t)b * 8]] + w.x;
      for (uint32_t k4 = 0; k4 < s4; k4++) {
        const uint32_t e = (k4 * 4 + kk) * 8 + g;
        base[k4 * 32 + lane] = e < w.y ? src[beg + e] : pad;
      }
    } else {
      const uint32_t r = rows[(size_t)b * 8 + g];
      const uint32_t beg = r == kEllNone ? 0u : row_beg[r], len = r == kEllNone ? 0u : row_len[r];
      for (uint32_t k4 = 0; k4 < s4; k4++) {
        const uint32_t k = k4 * 4 + kk;
        base[k4 * 32 + lane] = k < len ? src[beg + k] : pad;
      }
    }
  }
}

template <class T>
int ell_upload(T **dst, const T *src, size_t n, cudaStream_t stream) {
  *dst = nullptr;
  GCNB_CHECK(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) GCNB_CHECK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

}  // namespace

void ell_destroy(EllDev *e) {
  if (!e) return;
  cudaFree(e->d_idx); cudaFree(e->d_off); cudaFree(e->d_steps); cudaFree(e->d_rows); cudaFree(e->d_split_row);
  cudaFree(e->d_split_ptr); cudaFree(e->d_slots); cudaFree(e->d_counter);
  delete e;
}

int ell_upload_plan(const EllHost &H, cudaStream_t stream, EllDev **out) {
  *out = nullptr;
  auto *e = new EllDev();
  e->n_rows = H.n_rows; e->n_cols = H.n_cols; e->nnz = H.nnz; e->n_bundles = H.n_bundles;
  e->n_split = (int64_t)H.split_row.size(); e->n_slots = H.n_slots;
  e->static_schedule = H.wide_min != kEllWideMin;
  int rc = 0;
  if (!rc) rc = ell_upload(&e->d_idx, H.idx.data(), H.idx.size(), stream);
  if (!rc) rc = ell_upload(&e->d_off, H.off.data(), H.off.size(), stream);
  if (!rc) rc = ell_upload(&e->d_steps, H.steps.data(), H.steps.size(), stream);
  if (!rc) rc = ell_upload(&e->d_rows, H.rows.data(), H.rows.size(), stream);
  if (!rc) rc = ell_upload(&e->d_split_row, H.split_row.data(), H.split_row.size(), stream);
  if (!rc) rc = ell_upload(&e->d_split_ptr, H.split_ptr.data(), H.split_ptr.size(), stream);
  if (!rc) rc = (int)cudaMalloc((void **)&e->d_slots, std::max<size_t>((size_t)H.n_slots * 16 * sizeof(float), 16));
  if (!rc) rc = (int)cudaMalloc((void **)&e->d_counter, 2 * sizeof(uint32_t));
  if (!rc) rc = (int)cudaMemsetAsync(e->d_counter, 0, 2 * sizeof(uint32_t), stream);
  if (!rc) rc = (int)cudaStreamSynchronize(stream);  // the host arrays may go out of scope
  if (rc) {
    ell_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

// The same plan built where the entries already are: d_src[d_row_beg[i] + k] = column of the k-th entry of row i,
// k < d_row_len[i] (device arrays; the bit-tile device builder leaves its remainder like this).  Only the row lengths travel to
// the host (4 bytes per row) for the bundle layout; the index array is filled by ell_fill_kernel.  Bit-identical to
// ell_build_host + ell_upload_plan on the compacted CSR.
int ell_build_device(const uint32_t *d_row_beg, const uint32_t *d_row_len, const uint32_t *d_src, int64_t n_rows, int64_t n_cols,
                     cudaStream_t stream, EllDev **out) {
  *out = nullptr;
  if (n_rows < 0 || n_cols < 0 || n_rows > 0xfffffff0ll || n_cols > 0xfffffff0ll) return GCNB_E_BADARG;
  std::vector<uint32_t> len_of_row((size_t)n_rows);
  if (n_rows) GCNB_CHECK(cudaMemcpyAsync(len_of_row.data(), d_row_len, (size_t)n_rows * 4, cudaMemcpyDeviceToHost, stream));
  GCNB_CHECK(cudaStreamSynchronize(stream));
  EllHost H;
  H.n_rows = n_rows;
  H.n_cols = n_cols;
  uint64_t nnz = 0;
  for (uint32_t l : len_of_row) nnz += l;
  if (nnz > 0xfffffff0ull) return GCNB_E_BADARG;
  H.nnz = (int64_t)nnz;
  std::vector<EllWide> wide;
  std::vector<uint32_t> narrow;
  uint64_t acc = 0;
  const int lrc = ell_layout(len_of_row.data(), n_rows, H, wide, narrow, &acc);
  if (lrc) return lrc;
  std::vector<uint2> wide_part(wide.size());
  for (size_t b = 0; b < wide.size(); b++) wide_part[b] = make_uint2(wide[b].beg, wide[b].len);
  auto *e = new EllDev();
  e->n_rows = H.n_rows; e->n_cols = H.n_cols; e->nnz = H.nnz; e->n_bundles = H.n_bundles;
  e->n_split = (int64_t)H.split_row.size(); e->n_slots = H.n_slots;
  e->static_schedule = H.wide_min != kEllWideMin;
  uint2 *d_wide = nullptr;
  int rc = 0;
  if (!rc) rc = (int)cudaMalloc((void **)&e->d_idx, std::max<size_t>((size_t)acc * 32, 1) * 4);
  if (!rc) rc = ell_upload(&e->d_off, H.off.data(), H.off.size(), stream);
  if (!rc) rc = ell_upload(&e->d_steps, H.steps.data(), H.steps.size(), stream);
  if (!rc) rc = ell_upload(&e->d_rows, H.rows.data(), H.rows.size(), stream);
  if (!rc) rc = ell_upload(&e->d_split_row, H.split_row.data(), H.split_row.size(), stream);
  if (!rc) rc = ell_upload(&e->d_split_ptr, H.split_ptr.data(), H.split_ptr.size(), stream);
  if (!rc) rc = ell_upload(&d_wide, wide_part.data(), wide_part.size(), stream);
  if (!rc) rc = (int)cudaMalloc((void **)&e->d_slots, std::max<size_t>((size_t)H.n_slots * 16 * sizeof(float), 16));
  if (!rc) rc = (int)cudaMalloc((void **)&e->d_counter, 2 * sizeof(uint32_t));
  if (!rc) rc = (int)cudaMemsetAsync(e->d_counter, 0, 2 * sizeof(uint32_t), stream);
  if (!rc && H.n_bundles > 0) {
    const int64_t want = (H.n_bundles + 7) / 8;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)std::max(1, device_info().sm_count) * 16));
    ell_fill_kernel<<<grid, 256, 0, stream>>>(e->d_off, e->d_steps, e->d_rows, d_wide, d_row_beg, d_row_len, d_src, (uint32_t)n_cols,
                                              e->d_idx, (uint32_t)H.n_bundles);
    rc = (int)cudaPeekAtLastError();
  }
  if (!rc) rc = (int)cudaStreamSynchronize(stream);  // the host arrays may go out of scope
  cudaFree(d_wide);
  if (rc) {
    ell_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

int ell_launch(EllDev *e, const float *d_B2, const float *d_row_scale, float *d_R, int64_t ldr, int accumulate,
               const uint32_t *d_row_map, int ctas_per_sm, cudaStream_t stream) {
  if (e->n_bundles == 0) return 0;
  const DeviceInfo &di = device_info();
  const int per_sm = ctas_per_sm > 0 ? ctas_per_sm : 2;
  const int64_t want = (e->n_bundles + 7) / 8;  // 8 warps per CTA
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)std::max(1, di.sm_count) * per_sm));
  ell_gather16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(e->d_idx), e->d_off, e->d_steps, e->d_rows,
                                                d_B2, d_row_scale, d_R, ldr, accumulate, d_row_map, e->d_slots, e->d_counter,
                                                (uint32_t)e->n_bundles, e->static_schedule ? (uint32_t)grid * 8u : 0u);
  GCNB_LAUNCH_CHECK();
  if (e->n_split > 0) {
    ell_combine_kernel<<<(unsigned)((e->n_split * 4 + 255) / 256), 256, 0, stream>>>(e->d_split_row, e->d_split_ptr, e->d_slots,
                                                                                     d_row_scale, d_R, ldr, accumulate, d_row_map,
                                                                                     e->n_split);
    GCNB_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace gcnb

// =====================================================================================================================
// C ABI: the host builder alone (CPU tests consume the arrays exactly as the kernel does) and a stand-alone device plan
// =====================================================================================================================
struct gcnb_ell_host {
  EllHost H;
};
struct gcnb_ell_plan {
  EllDev *dev = nullptr;
};

extern "C" {

int gcnb_ell_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols, int n_threads,
                        gcnb_ell_host **out) {
  if (!out) return GCNB_E_BADARG;
  auto *h = new gcnb_ell_host();
  const int rc = ell_build_host(h_indptr, h_indices, n_rows, n_cols, n_threads, h->H);
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

// out = {rows, cols, entries, bundles, index words (uint32), split rows, partial slots, wide-bundle minimum length}
int gcnb_ell_host_sizes(const gcnb_ell_host *h, int64_t out[8]) {
  if (!h || !out) return GCNB_E_BADARG;
  const EllHost &H = h->H;
  out[0] = H.n_rows; out[1] = H.n_cols; out[2] = H.nnz; out[3] = H.n_bundles; out[4] = (int64_t)H.idx.size();
  out[5] = (int64_t)H.split_row.size(); out[6] = H.n_slots; out[7] = H.wide_min;
  return 0;
}

// which: 0 idx, 1 off (bundles + 1), 2 steps, 3 rows (bundles x 8), 4 split_row, 5 split_ptr.  Copies min(bytes, size).
int gcnb_ell_host_copy(const gcnb_ell_host *h, int which, void *dst, int64_t bytes) {
  if (!h || !dst || bytes < 0) return GCNB_E_BADARG;
  const EllHost &H = h->H;
  const void *src = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: src = H.idx.data(); n = H.idx.size() * 4; break;
    case 1: src = H.off.data(); n = H.off.size() * 4; break;
    case 2: src = H.steps.data(); n = H.steps.size() * 4; break;
    case 3: src = H.rows.data(); n = H.rows.size() * 4; break;
    case 4: src = H.split_row.data(); n = H.split_row.size() * 4; break;
    case 5: src = H.split_ptr.data(); n = H.split_ptr.size() * 4; break;
    default: return GCNB_E_BADARG;
  }
  if (n) memcpy(dst, src, std::min<size_t>(n, (size_t)bytes));
  return 0;
}

int gcnb_ell_host_destroy(gcnb_ell_host *h) {
  delete h;
  return 0;
}

int gcnb_ell_plan_create(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols,
                         gcnb_stream_t stream, gcnb_ell_plan **out) {
  if (!out) return GCNB_E_BADARG;
  *out = nullptr;
  if (!device_info().ok) return (int)cudaErrorNoDevice;
  EllHost H;
  int rc = ell_build_host(h_indptr, h_indices, n_rows, n_cols, 0, H);
  if (rc) return rc;
  auto *p = new gcnb_ell_plan();
  rc = ell_upload_plan(H, as_stream(stream), &p->dev);
  if (rc) {
    delete p;
    return rc;
  }
  *out = p;
  return 0;
}

int gcnb_ell_plan_destroy(gcnb_ell_plan *p) {
  if (!p) return 0;
  ell_destroy(p->dev);
  delete p;
  return 0;
}

// R[n_rows x 16] = diag(row_scale) * pattern * B2; d_B2 holds n_cols + 1 rows of 16 floats, the LAST ONE ALL ZERO
int gcnb_ell_gather16_f32(gcnb_ell_plan *p, const float *d_B2, const float *d_row_scale, float *d_R, gcnb_stream_t stream) {
  if (!p || !p->dev || !d_B2 || !d_row_scale || !d_R) return GCNB_E_BADARG;
  return ell_launch(p->dev, d_B2, d_row_scale, d_R, 16, 0, nullptr, 0, as_stream(stream));
}

}  // extern "C"
