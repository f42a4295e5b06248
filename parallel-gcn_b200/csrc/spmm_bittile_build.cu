// spmm_bittile_build.cu -- the bit-tile GraphSum plan (spmm_bittile.cu) built ON THE GPU from the device-resident CSR.
//
// Why: the host builder needs the graph on the host (the engine reads 0.9 GB back over PCIe for the Reddit-shape graph), walks
// 115 M entries with 16 threads and uploads 260 MB of bit maps and indices: ~0.3 s, during which training epochs ran on the
// generic kernel (the background build of round 2; 20 epochs of an end-to-end run never saw the fast path).  Here the CSR
// never leaves HBM; the host only sees 4 bytes per row block (tile counts, for the CTA schedule) and 4 bytes per row (remainder
// lengths, for the ELL bundle layout).  The result is BIT-IDENTICAL to gcnb_bittile_plan_create on the same matrix
// (tests/test_zz_bittile_gpu.py compares every array of the two plans), so which builder ran never shows in a result.
//
//   btb_scales_kernel   s_i = sqrt(first diagonal entry of row i)                      (when the caller passes no scales)
//   btb_count_kernel    one CTA per row block: shared-memory histogram of the block's entries over the column chunks,
//                       tiles of the block = chunks holding >= min_tile_nnz entries
//   (host)              CTA schedule from the tile counts (bittile_schedule, the host builder's)
//   btb_fill_kernel     one CTA per row block: histogram again, chunk -> tile slot by a block-wide scan, then one warp per row
//                       walks the row IN ORDER, 32 entries at a time: entries of selected chunks whose value factors set
//                       their bit (64-bit atomicOr: the old word tells a duplicate entry, which cannot be a second bit), the
//                       others are compacted in order into the row's remainder
//   ell_build_device    remainder -> ELL bundles (spmm_ell.cu)
// Restrictions (the caller falls back to the host builder on GCNB_E_UNSUPPORTED): the chunk histogram must fit shared memory
// (32-bit counters up to ~3.2 M columns at 64 columns per chunk, saturating 16-bit counters up to 4.19 M) and every entry must
// factor as row_scale * col_scale (GraphSum's always do).
//
// Reference being replaced: none (set-up of the kernels that replace graphsum_kernel, src/module.cu:172-186).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "spmm_bittile.cuh"

using namespace gcnb;

namespace {

constexpr int kBtbThreads = 512;
constexpr size_t kBtbSmemMax = 200 * 1024;  // dynamic shared memory of the two block kernels: 4 bytes per column chunk ...
constexpr int64_t kBtbMaxChunks16 = 65534;   // ... or 2 (saturating 16-bit counters; a tile slot must fit 16 bits too)

__device__ __forceinline__ int btb_bit_of_col(uint32_t c) {  // = bt_bit_of_col (spmm_bittile.cu), c in [0, 64)
  const uint32_t cc = c & 31u;
  return (int)((c & 32u) + (cc >> 1) + 16u * (cc & 1u));
}

// s[i] = sqrtf(value of the FIRST entry (i, i)) when that value is positive, else NaN; one warp per row
__global__ void __launch_bounds__(256) btb_scales_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                                                         const float *__restrict__ values, int64_t n_rows, float *__restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_rows; i += n_warps) {
    const uint32_t e0 = indptr[i], e1 = indptr[i + 1];
    float out = nanf("");
    for (uint32_t base = e0; base < e1; base += 32) {
      const uint32_t e = base + lane;
      const bool hit = e < e1 && indices[e] == (uint32_t)i;
      const uint32_t bal = __ballot_sync(0xffffffffu, hit);
      if (bal) {
        const float v = values[base + (uint32_t)(__ffs(bal) - 1)];
        if (v > 0.f) out = sqrtf(v);
        break;
      }
    }
    if (lane == 0) s[i] = out;
  }
}

// rows / columns without a usable scale own no bit and no entry: the kernels multiply by the scales unconditionally
__global__ void btb_clean_scales_kernel(float *__restrict__ s, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!(fabsf(s[i]) <= 3.0e38f)) s[i] = 0.f;
}

struct BtbArgs {
  const uint32_t *indptr, *indices;
  const float *values;  // may be NULL: a pattern scaled by row_scale x col_scale
  const float *row_scale, *col_scale;
  int64_t n_rows;
  uint32_t n_chunks, thr;
  int bh, shift, chunk_cols, wpr;
  // count pass
  uint32_t *tiles_of_block;
  // fill pass
  const unsigned long long *tile_base;  // first tile of every row block
  uint32_t *tile_chunk;
  unsigned long long *bits;
  uint32_t *rem_idx;   // rem_idx[indptr[i] + k] = column of the k-th remainder entry of row i
  uint32_t *rem_len;   // remainder entries of row i
  unsigned long long *counters;  // [0] entries in tiles, [1] entries that do not factor
};

// Counters: 32-bit, or -- for column ranges whose 32-bit histogram would not fit shared memory -- 16-bit halves of 32-bit
// words that SATURATE: a count is only ever compared with the threshold (<= 0x7fff), and an increment is skipped once the
// half reads 0x7fff or more; the at most 512 threads that pass that test together leave it below 0x8200, so a half never
// carries into its neighbour.
template <class CT>
struct BtbCounter;
template <>
struct BtbCounter<uint32_t> {
  static constexpr uint32_t kNoTile = 0xffffffffu;
  static __device__ __forceinline__ void inc(uint32_t *cnt, uint32_t c) { atomicAdd(&cnt[c], 1u); }
};
template <>
struct BtbCounter<uint16_t> {
  static constexpr uint32_t kNoTile = 0xffffu;
  static __device__ __forceinline__ void inc(uint16_t *cnt, uint32_t c) {
    uint32_t *w = reinterpret_cast<uint32_t *>(cnt) + (c >> 1);
    const uint32_t sh = (c & 1u) * 16u;
    if (((*reinterpret_cast<volatile uint32_t *>(w) >> sh) & 0xffffu) < 0x7fffu) atomicAdd(w, 1u << sh);
  }
};

template <class CT>
__device__ __forceinline__ void btb_histogram(const BtbArgs &a, CT *cnt, int64_t r0, int64_t r1) {
  const uint32_t n_zero = sizeof(CT) == 2 ? a.n_chunks + (a.n_chunks & 1u) : a.n_chunks;  // whole 32-bit words
  for (uint32_t c = threadIdx.x; c < n_zero; c += blockDim.x) cnt[c] = (CT)0;
  __syncthreads();
  const uint32_t e1 = a.indptr[r1];
  for (uint32_t e = a.indptr[r0] + threadIdx.x; e < e1; e += blockDim.x) {
    const uint32_t c = a.indices[e] >> a.shift;
    if (c < a.n_chunks) BtbCounter<CT>::inc(cnt, c);
  }
  __syncthreads();
}

template <class CT>
__global__ void __launch_bounds__(kBtbThreads) btb_count_kernel(BtbArgs a) {
  extern __shared__ uint32_t btb_smem[];
  CT *cnt = reinterpret_cast<CT *>(btb_smem);
  __shared__ uint32_t total;
  const int64_t r0 = (int64_t)blockIdx.x * a.bh, r1 = min(a.n_rows, r0 + a.bh);
  if (threadIdx.x == 0) total = 0u;
  btb_histogram(a, cnt, r0, r1);
  uint32_t mine = 0;
  for (uint32_t c = threadIdx.x; c < a.n_chunks; c += blockDim.x) mine += cnt[c] >= a.thr ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&total, mine);
  __syncthreads();
  if (threadIdx.x == 0) a.tiles_of_block[blockIdx.x] = total;
}

template <class CT>
__global__ void __launch_bounds__(kBtbThreads) btb_fill_kernel(BtbArgs a) {
  extern __shared__ uint32_t btb_smem[];
  CT *cnt = reinterpret_cast<CT *>(btb_smem);  // histogram, then chunk -> tile slot of this block (kNoTile: not a tile)
  constexpr uint32_t kBtbNoTile = BtbCounter<CT>::kNoTile;
  __shared__ uint32_t warp_tot[kBtbThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * a.bh, r1 = min(a.n_rows, r0 + a.bh);
  btb_histogram(a, cnt, r0, r1);
  // ---- slots in ascending chunk order: thread t owns chunks [t * per, (t + 1) * per)
  const uint32_t per = (a.n_chunks + blockDim.x - 1) / blockDim.x;
  const uint32_t c0 = min(a.n_chunks, threadIdx.x * per), c1 = min(a.n_chunks, c0 + per);
  uint32_t mine = 0;
  for (uint32_t c = c0; c < c1; c++) mine += cnt[c] >= a.thr ? 1u : 0u;
  uint32_t incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t slot = incl - mine;
  for (int w = 0; w < warp; w++) slot += warp_tot[w];
  const unsigned long long tb = a.tile_base[blockIdx.x];
  for (uint32_t c = c0; c < c1; c++) {
    if (cnt[c] >= a.thr) {
      a.tile_chunk[tb + slot] = c;
      cnt[c] = (CT)slot++;
    } else {
      cnt[c] = (CT)kBtbNoTile;
    }
  }
  __syncthreads();
  // ---- rows: one warp per row, the row's entries in order
  unsigned long long n_tile = 0, n_unf = 0;
  const uint32_t col_mask = (uint32_t)a.chunk_cols - 1u;
  for (int64_t i = r0 + warp; i < r1; i += (blockDim.x >> 5)) {
    const uint32_t rl = (uint32_t)(i - r0);
    const float si = a.row_scale[i];
    const uint32_t e0 = a.indptr[i], e1 = a.indptr[i + 1];
    uint32_t rc = 0;
    for (uint32_t base = e0; base < e1; base += 32) {
      const uint32_t e = base + lane;
      const bool act = e < e1;
      uint32_t j = 0;
      bool factors = false, cand = false;
      uint32_t li = kBtbNoTile;
      if (act) {
        j = a.indices[e];
        const float p = __fmul_rn(si, a.col_scale[j]);
        const float v = a.values ? a.values[e] : p;
        factors = fabsf(__fsub_rn(v, p)) <= __fmul_rn(1e-6f, fabsf(v));  // false for NaN scales
        const uint32_t c = j >> a.shift;
        li = c < a.n_chunks ? cnt[c] : kBtbNoTile;
        cand = factors && li != kBtbNoTile;
      }
      // a duplicate entry (i, j) cannot be a second bit: inside these 32 entries the lowest lane owns the bit, across groups
      // the old word returned by the atomic tells
      const unsigned long long key = act ? (unsigned long long)j : (0x100000000ull | (unsigned long long)lane);
      const uint32_t same = __match_any_sync(0xffffffffu, key);
      bool in_tile = false;
      if (cand && lane == __ffs(same) - 1) {
        const uint32_t cc = j & col_mask;
        const unsigned long long m = 1ull << btb_bit_of_col(cc & 63u);
        unsigned long long *w = a.bits + ((tb + li) * (unsigned long long)a.bh + rl) * (unsigned long long)a.wpr + (cc >> 6);
        in_tile = (atomicOr(w, m) & m) == 0ull;
      }
      const uint32_t rem = __ballot_sync(0xffffffffu, act && !in_tile);
      if (act && !in_tile) a.rem_idx[e0 + rc + __popc(rem & ((1u << lane) - 1u))] = j;
      rc += __popc(rem);
      n_tile += __popc(__ballot_sync(0xffffffffu, in_tile));  // (every lane counts; lane 0 reports)
      n_unf += __popc(__ballot_sync(0xffffffffu, act && !factors));
    }
    if (lane == 0) a.rem_len[i] = rc;
  }
  if (lane == 0) {
    if (n_tile) atomicAdd(a.counters, n_tile);
    if (n_unf) atomicAdd(a.counters + 1, n_unf);
  }
}

template <class T>
int btb_upload(T **dst, const T *src, size_t n, cudaStream_t stream) {
  *dst = nullptr;
  GCNB_CHECK(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) GCNB_CHECK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

struct Scratch {  // freed on every exit path
  std::vector<void *> ptrs;
  ~Scratch() { release(); }
  void release() {
    for (void *p : ptrs) cudaFree(p);
    ptrs.clear();
  }
  template <class T>
  int alloc(T **dst, size_t n) {
    *dst = nullptr;
    GCNB_CHECK(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
    ptrs.push_back(*dst);
    return 0;
  }
};

}  // namespace

extern "C" {

// 1 when a matrix of n_cols columns fits the device builder's shared-memory histogram (one counter per column chunk)
int gcnb_bittile_device_build_fits(int64_t n_cols, int chunk_cols) {
  if (chunk_cols != 64 && chunk_cols != 128) chunk_cols = 64;
  const int64_t n_chunks = (n_cols + chunk_cols - 1) / chunk_cols;
  return n_cols > 0 && ((size_t)n_chunks * 4 <= kBtbSmemMax || n_chunks <= kBtbMaxChunks16);
}

// gcnb_bittile_plan_create with every input array ON THE DEVICE (d_values / the two scale arrays may be NULL as there: scales
// from the diagonal of a square matrix; a pattern when only the scales are given).  GCNB_E_UNSUPPORTED: this matrix needs the
// host builder (entries that do not factor, or more column chunks than the shared-memory histogram takes).
int gcnb_bittile_plan_create_device(const uint32_t *d_indptr, const uint32_t *d_indices, const float *d_values, int64_t n_rows,
                                    int64_t n_cols, const float *d_row_scale, const float *d_col_scale, int min_tile_nnz,
                                    int chunk_cols, int row_blocks, gcnb_stream_t stream_, gcnb_bittile_plan **out) {
  if (!out) return GCNB_E_BADARG;
  *out = nullptr;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  if (di.cc_major != 10) return GCNB_E_UNSUPPORTED;  // tcgen05 / TMEM
  if (!d_indptr || n_rows <= 0 || n_cols <= 0 || n_rows > 0xfffffff0ll || n_cols > 0xfffffff0ll) return GCNB_E_BADARG;
  if ((d_row_scale == nullptr) != (d_col_scale == nullptr)) return GCNB_E_BADARG;
  if (!d_values && !d_row_scale) return GCNB_E_BADARG;
  if (!d_row_scale && n_rows != n_cols) return GCNB_E_UNSUPPORTED;  // no diagonal to take the scales from
  cudaStream_t stream = as_stream(stream_);
  // same defaults and probes as gcnb_bittile_plan_create
  if (chunk_cols == 0)
    if (const char *e = getenv("GCNB_BT_CHUNK")) chunk_cols = atoi(e);
  if (row_blocks == 0)
    if (const char *e = getenv("GCNB_BT_RB")) row_blocks = atoi(e);
  if (chunk_cols == 0 && row_blocks == 0) {
    chunk_cols = 64;
    row_blocks = 2;
  }
  if (chunk_cols == 0) chunk_cols = row_blocks == 2 ? 64 : 128;
  if (row_blocks == 0) row_blocks = 1;
  if (chunk_cols != 64 && chunk_cols != 128) return GCNB_E_BADARG;
  if (row_blocks != 1 && !(row_blocks == 2 && chunk_cols == 64)) return GCNB_E_BADARG;
  const int64_t BH = (int64_t)kBtRows * row_blocks;
  const int64_t n_blk = (n_rows + BH - 1) / BH;
  const int64_t n_chunks = (n_cols + chunk_cols - 1) / chunk_cols;
  // 32-bit counters while they fit shared memory, else saturating 16-bit ones (GCNB_BTB_COUNTER16=1: test probe, always 16)
  bool counter16 = (size_t)n_chunks * 4 > kBtbSmemMax;
  if (const char *e = getenv("GCNB_BTB_COUNTER16")) counter16 = counter16 || atoi(e) != 0;
  const uint32_t thr_arg = (uint32_t)(min_tile_nnz > 0 ? min_tile_nnz : 2 * chunk_cols * row_blocks);
  if (counter16 && (n_chunks > kBtbMaxChunks16 || thr_arg > 0x7fffu)) return GCNB_E_UNSUPPORTED;
  const size_t smem = counter16 ? ((size_t)n_chunks * 2 + 3) / 4 * 4 : (size_t)n_chunks * 4;
  const bool verbose = getenv("GCNB_SETUP_VERBOSE") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    cudaStreamSynchronize(stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bittile device build] %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
    tp = now;
  };

  auto *p = new gcnb_bittile_plan();
  auto fail = [&](int code) {
    gcnb_bittile_plan_destroy(p);
    if (code == (int)cudaErrorMemoryAllocation) {  // no room for the scratch (4 bytes per entry): the host builder's business
      cudaGetLastError();
      return GCNB_E_UNSUPPORTED;
    }
    return code;
  };
  Scratch scratch;
  int rc = 0;
  uint32_t nnz32 = 0;
  if ((rc = (int)cudaMemcpyAsync(&nnz32, d_indptr + n_rows, 4, cudaMemcpyDeviceToHost, stream))) return fail(rc);
  // ---- scales (owned by the plan)
  if ((rc = (int)cudaMalloc((void **)&p->d_row_scale, (size_t)n_rows * 4))) return fail(rc);
  if ((rc = (int)cudaMalloc((void **)&p->d_col_scale, (size_t)n_cols * 4))) return fail(rc);
  if (d_row_scale) {
    if ((rc = (int)cudaMemcpyAsync(p->d_row_scale, d_row_scale, (size_t)n_rows * 4, cudaMemcpyDeviceToDevice, stream))) return fail(rc);
    if ((rc = (int)cudaMemcpyAsync(p->d_col_scale, d_col_scale, (size_t)n_cols * 4, cudaMemcpyDeviceToDevice, stream))) return fail(rc);
  } else {
    if (!d_indices) return fail(GCNB_E_BADARG);
    btb_scales_kernel<<<(unsigned)std::min<int64_t>((n_rows + 7) / 8, (int64_t)di.sm_count * 16), 256, 0, stream>>>(
        d_indptr, d_indices, d_values, n_rows, p->d_row_scale);
    if ((rc = (int)cudaPeekAtLastError())) return fail(rc);
    if ((rc = (int)cudaMemcpyAsync(p->d_col_scale, p->d_row_scale, (size_t)n_rows * 4, cudaMemcpyDeviceToDevice, stream))) return fail(rc);
  }
  lap("scales");
  // ---- tiles per row block
  BtbArgs a{};
  a.indptr = d_indptr; a.indices = d_indices; a.values = d_values; a.row_scale = p->d_row_scale; a.col_scale = p->d_col_scale;
  a.n_rows = n_rows; a.n_chunks = (uint32_t)n_chunks;
  a.thr = thr_arg;
  a.bh = (int)BH; a.shift = chunk_cols == 128 ? 7 : 6; a.chunk_cols = chunk_cols; a.wpr = chunk_cols / 64;
  if ((rc = scratch.alloc(&a.tiles_of_block, (size_t)n_blk))) return fail(rc);
  auto count_fn = counter16 ? btb_count_kernel<uint16_t> : btb_count_kernel<uint32_t>;
  auto fill_fn = counter16 ? btb_fill_kernel<uint16_t> : btb_fill_kernel<uint32_t>;
  if ((rc = (int)cudaFuncSetAttribute(count_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return fail(rc);
  if ((rc = (int)cudaFuncSetAttribute(fill_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))) return fail(rc);
  count_fn<<<(unsigned)n_blk, kBtbThreads, smem, stream>>>(a);
  if ((rc = (int)cudaPeekAtLastError())) return fail(rc);
  std::vector<uint32_t> tiles_of_block((size_t)n_blk);
  if ((rc = (int)cudaMemcpyAsync(tiles_of_block.data(), a.tiles_of_block, (size_t)n_blk * 4, cudaMemcpyDeviceToHost, stream))) return fail(rc);
  if ((rc = (int)cudaStreamSynchronize(stream))) return fail(rc);
  lap("tile counts");
  // ---- CTA schedule (host, 4 bytes per row block)
  std::vector<uint32_t> cta_tile_ptr, cta_item_ptr;
  std::vector<uint2> items;
  std::vector<uint64_t> tile_base;
  const int n_cta = std::max(1, di.sm_count);
  const int64_t n_tiles = bittile_schedule(tiles_of_block.data(), n_blk, n_cta, chunk_cols, row_blocks, cta_tile_ptr, cta_item_ptr, items,
                                           tile_base);
  if (n_tiles < 0) return fail(GCNB_E_BADARG);
  p->n_rows = n_rows; p->n_cols = n_cols; p->nnz = nnz32; p->n_blk = n_blk; p->n_tiles = n_tiles; p->n_cta = n_cta;
  p->chunk = chunk_cols; p->rb = row_blocks;
  p->n_chunks = (n_cols + 127) / 128 * 2;
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "tile_base is uploaded as 64-bit words");
  unsigned long long *d_tile_base = nullptr;
  if ((rc = scratch.alloc(&d_tile_base, (size_t)n_blk))) return fail(rc);
  if ((rc = (int)cudaMemcpyAsync(d_tile_base, tile_base.data(), (size_t)n_blk * 8, cudaMemcpyHostToDevice, stream))) return fail(rc);
  if ((rc = btb_upload(&p->d_cta_tile_ptr, cta_tile_ptr.data(), cta_tile_ptr.size(), stream))) return fail(rc);
  if ((rc = btb_upload(&p->d_cta_item_ptr, cta_item_ptr.data(), cta_item_ptr.size(), stream))) return fail(rc);
  if ((rc = btb_upload(&p->d_items, items.data(), items.size(), stream))) return fail(rc);
  const size_t n_words = (size_t)n_tiles * (size_t)BH * (size_t)a.wpr;
  if ((rc = (int)cudaMalloc((void **)&p->d_tile_chunk, std::max<size_t>((size_t)n_tiles, 1) * 4))) return fail(rc);
  if ((rc = (int)cudaMalloc((void **)&p->d_bits, std::max<size_t>(n_words, 1) * 8))) return fail(rc);
  if ((rc = (int)cudaMemsetAsync(p->d_bits, 0, std::max<size_t>(n_words, 1) * 8, stream))) return fail(rc);
  if ((rc = scratch.alloc(&a.rem_idx, (size_t)nnz32))) return fail(rc);
  if ((rc = scratch.alloc(&a.rem_len, (size_t)n_rows))) return fail(rc);
  if ((rc = scratch.alloc(&a.counters, 2))) return fail(rc);
  if ((rc = (int)cudaMemsetAsync(a.counters, 0, 16, stream))) return fail(rc);
  a.tile_base = d_tile_base; a.tile_chunk = p->d_tile_chunk; a.bits = reinterpret_cast<unsigned long long *>(p->d_bits);
  fill_fn<<<(unsigned)n_blk, kBtbThreads, smem, stream>>>(a);
  if ((rc = (int)cudaPeekAtLastError())) return fail(rc);
  unsigned long long counters[2] = {0, 0};
  if ((rc = (int)cudaMemcpyAsync(counters, a.counters, 16, cudaMemcpyDeviceToHost, stream))) return fail(rc);
  if ((rc = (int)cudaStreamSynchronize(stream))) return fail(rc);  // (also: the schedule's host arrays go out of scope)
  lap("bit maps + remainder");
  if (counters[1] != 0) return fail(GCNB_E_UNSUPPORTED);  // entries that do not factor keep their values: host builder
  p->tile_nnz = (int64_t)counters[0];
  p->rem_nnz = (int64_t)nnz32 - p->tile_nnz;
  p->n_unfactored = 0;
  bool use_ell = n_tiles > 0;
  if (const char *e = getenv("GCNB_BT_ELL")) use_ell = use_ell && atoi(e) != 0;
  if (!use_ell) return fail(GCNB_E_UNSUPPORTED);  // no tiles at all / the valued remainder probe: the host builder's business
  if ((rc = ell_build_device(d_indptr, a.rem_len, a.rem_idx, n_rows, n_cols, stream, &p->ell))) return fail(rc);
  lap("ELL remainder");
  const size_t b2_bytes = ((size_t)n_cols + 1) * 16 * sizeof(float);
  if ((rc = (int)cudaMalloc((void **)&p->d_B2, b2_bytes))) return fail(rc);
  if ((rc = (int)cudaMemsetAsync(p->d_B2, 0, b2_bytes, stream))) return fail(rc);  // the padding row stays zero
  btb_clean_scales_kernel<<<(unsigned)std::min<int64_t>((n_rows + 255) / 256, 1024), 256, 0, stream>>>(p->d_row_scale, n_rows);
  btb_clean_scales_kernel<<<(unsigned)std::min<int64_t>((n_cols + 255) / 256, 1024), 256, 0, stream>>>(p->d_col_scale, n_cols);
  if ((rc = (int)cudaPeekAtLastError())) return fail(rc);
  lap("B2 + scale clean-up");
  if ((rc = bittile_finish_plan(p, stream))) return fail(rc);
  lap("buffers");
  scratch.release();
  lap("scratch freed");
  *out = p;
  return 0;
}

// Debugging / test aid: the sizes (in elements) and contents of a plan's device arrays, whichever builder made them.
// which: 0 tile_chunk (u32), 1 bits (u64), 2 cta_tile_ptr, 3 cta_item_ptr, 4 items (u32 x 2), 5 row_scale (f32), 6 col_scale,
// 7 ELL idx (u32), 8 ELL off, 9 ELL steps, 10 ELL rows, 11 ELL split_row, 12 ELL split_ptr.  sizes[13] = n_slots, [14] = tile_nnz,
// [15] = rem_nnz.
int gcnb_bittile_plan_sizes(const gcnb_bittile_plan *p, int64_t sizes[16]) {
  if (!p || !sizes) return GCNB_E_BADARG;
  const int64_t wpr = p->chunk / 64;
  const int64_t n_items = [&] {
    uint32_t v = 0;
    if (p->d_cta_item_ptr) cudaMemcpy(&v, p->d_cta_item_ptr + p->n_cta, 4, cudaMemcpyDeviceToHost);
    return (int64_t)v;
  }();
  sizes[0] = p->n_tiles; sizes[1] = p->n_tiles * kBtRows * p->rb * wpr; sizes[2] = p->n_cta + 1; sizes[3] = p->n_cta + 1;
  sizes[4] = n_items; sizes[5] = p->n_rows; sizes[6] = p->n_cols;
  const EllDev *e = p->ell;
  uint32_t idx_rows = 0;
  if (e && e->d_off) cudaMemcpy(&idx_rows, e->d_off + e->n_bundles, 4, cudaMemcpyDeviceToHost);
  sizes[7] = e ? (int64_t)idx_rows * 32 : 0; sizes[8] = e ? e->n_bundles + 1 : 0; sizes[9] = e ? e->n_bundles : 0;
  sizes[10] = e ? e->n_bundles * 8 : 0; sizes[11] = e ? e->n_split : 0; sizes[12] = e ? e->n_split + 1 : 0;
  sizes[13] = e ? e->n_slots : 0; sizes[14] = p->tile_nnz; sizes[15] = p->rem_nnz;
  return 0;
}

int gcnb_bittile_plan_copy(const gcnb_bittile_plan *p, int which, void *h_dst, int64_t bytes) {
  if (!p || !h_dst || bytes < 0) return GCNB_E_BADARG;
  int64_t sizes[16];
  const int rc = gcnb_bittile_plan_sizes(p, sizes);
  if (rc) return rc;
  const EllDev *e = p->ell;
  const void *src = nullptr;
  int64_t elem = 4;
  switch (which) {
    case 0: src = p->d_tile_chunk; break;
    case 1: src = p->d_bits; elem = 8; break;
    case 2: src = p->d_cta_tile_ptr; break;
    case 3: src = p->d_cta_item_ptr; break;
    case 4: src = p->d_items; elem = 8; break;
    case 5: src = p->d_row_scale; break;
    case 6: src = p->d_col_scale; break;
    case 7: src = e ? e->d_idx : nullptr; break;
    case 8: src = e ? e->d_off : nullptr; break;
    case 9: src = e ? e->d_steps : nullptr; break;
    case 10: src = e ? e->d_rows : nullptr; break;
    case 11: src = e ? e->d_split_row : nullptr; break;
    case 12: src = e ? e->d_split_ptr : nullptr; break;
    default: return GCNB_E_BADARG;
  }
  const int64_t n = std::min<int64_t>(bytes, sizes[which] * elem);
  if (n > 0 && !src) return GCNB_E_BADARG;
  if (n > 0) GCNB_CHECK(cudaMemcpy(h_dst, src, (size_t)n, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
