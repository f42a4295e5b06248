// spmm_bittile.cu -- bit-tile GraphSum for sm_100a: the DENSE BLOCKS of the normalised adjacency go through the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM) as 0/1 bit maps; the rest stays a CSR on the generic kernel.
//
// Why: at 16 columns every gather formulation of  C = A_csr * B  is bound by the SM's LSU data pipe (half
// a wavefront per gathered 64-byte row from shared memory, one from L1), not by HBM: 0.30 of the HBM roofline on the
// Reddit-shape graph.  GraphSum's values factor: graph_value[i,j] = s_i * s_j with s = 1/sqrt(deg)
// (src/parser.cpp:164-181), so  C = diag(s) * M * (diag(s) * B)  with M the 0/1 pattern.  Inside a community M is
// dense enough (8 % on the bench graph) that a BIT MAP is the smaller encoding (1 bit per cell = 12 bits per entry
// against 48 for a 16-bit id + fp32 value), and a 0/1 operand is exact in bf16, so the product  M * B'  can run on
// the tensor cores WITHOUT losing fp32 accuracy: B' = s_j * B_j (fp32) is split into three bf16 pieces
// (8 + 8 + 8 significand bits, hi + mid + lo == B' exactly), every product 1.0 * piece is exact and the sums are
// accumulated in fp32.  No neighbour row is gathered at all.
//
// How: rows are cut into blocks of 128, columns into chunks of 64 or 128 (default 128).  A TILE (row block x chunk) holding at
// least min_tile_nnz entries that satisfy the factorisation becomes a bit map (1 or 2 KB); all other entries (sparse
// tiles, duplicate entries, values that are not s_i * s_j) form the REMAINDER.  Per launch:
//   bt_pack_kernel     B' = s_j * B_j split into bf16 hi | mid | lo, stored chunk by chunk as the K-major core-matrix
//                      image tcgen05.mma reads from shared memory (48 x 16 per k-step, no swizzle; 96 bytes per row of B),
//                      plus the fp32 copy of B' the pattern-only remainder kernel gathers from
//   bt_mma_wide_kernel<RB, W>  one persistent CTA per SM, warp-specialised (see the comment above it):
//                        warps 0-7   expand bit-map words into bf16 0/1 A operands and write them to TMEM (tcgen05.st):
//                                    thread = row, 64 bits -> 32 packed registers with 2 integer ops per register
//                        warp  12    streams the packed chunks of B' into a shared-memory ring (cp.async.bulk)
//                        warp  13    converged, one elected lane issues tcgen05.mma.kind::f16, M = 128, N = 48 (3 pieces
//                                    x 16 columns), K = 16, A from TMEM, B from shared memory, D in TMEM; tcgen05.commit
//                                    frees stages
//                        warps 8-11  epilogue: tcgen05.ld the accumulators, add pieces small to large, scale by s_i,
//                                    store the block's partial rows
//   remainder          second stream, concurrently (it lives on the LSU pipe, the MMA path does not): the pattern-only
//                      ELL gather of spmm_ell.cu when every remainder entry factors (GraphSum always), else the
//                      valued CSR on the generic kernel (spmm.cu)
//   bt_add_kernel      C = P + R
// Tile shapes: W = 2 -> 128 x 128 tiles (default; 128 us for the MMA kernel on the bench graph), RB = 2 -> items of 256
// rows x 64 columns whose two halves share a B' stage (134 us), W = RB = 1 -> 128 x 64 (190 us).
// Summation order is fixed by the plan => bit-reproducible run to run.  Relative to the CSR product the result differs
// by the rounding of s_i * s_j against 1/sqrtf(deg_i * deg_j) and by the summation order (~1e-6 relative; parity bar 1e-5).
//
// Reference being replaced: graphsum_kernel, src/module.cu:172-186.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <thread>
#include <vector>

#include "bulk.cuh"
#include "common.cuh"
#include "spmm_plan.cuh"
#include "spmm_bittile.cuh"
#include "tcgen05.cuh"

using namespace gcnb;

namespace gcnb {


// Host-side plan (pure CPU; unit-tested without a GPU through gcnb_bittile_host_*).
//   CTA q processes tiles [cta_tile_ptr[q], cta_tile_ptr[q+1]) in order; they belong to its items
//   [cta_item_ptr[q], cta_item_ptr[q+1]);  items[k] = (row block, end position of the block's tiles RELATIVE to the
//   CTA's first tile).  tile_chunk[t] = column chunk of tile t; bits[t*128 + r] = the 64 cells of row r of tile t
//   (in general bits[((t*rb + sub)*128 + r)*(chunk/64) + w]: sub = 128-row half of a 256-row item, w = column / 64;
//   items[k].x counts blocks of 128*rb rows):
//   low word = columns 0..31, high word = columns 32..63; inside a word column c sits at bit (c >> 1) + 16 * (c & 1),
//   which lets register q of the expansion (columns 2q, 2q+1 as a bf16 pair) be  (word & (0x00010001 << q)) * (0x3F80 >> q).
struct BitTileHost {
  int64_t n_rows = 0, n_cols = 0, nnz = 0, n_blk = 0, n_tiles = 0, tile_nnz = 0;
  int n_cta = 0, min_tile_nnz = 0;
  int chunk = kBtChunk;  // columns per tile: 64 (one 64-bit word per row) or 128 (two adjacent words per row)
  int rb = 1;            // row blocks of 128 per item (2: super-tiles of 256 rows whose two MMAs share one B' stage)
  std::vector<uint32_t> tile_chunk, cta_tile_ptr, cta_item_ptr;
  std::vector<uint2> items;
  HostArray<uint64_t> bits;
  std::vector<uint32_t> r_indptr;
  HostArray<uint32_t> r_indices;
  HostArray<float> r_values;
  std::vector<float> row_scale, col_scale;
  int64_t n_unfactored = 0;  // remainder entries whose value is NOT row_scale * col_scale (0: the pattern-only ELL kernel applies)
};

namespace {

template <class F>
void bt_run_threads(int T, F f) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back([&f, t] { f(t); });
  f(0);
  for (auto &x : th) x.join();
}

inline int bt_bit_of_col(uint32_t c) {  // c in [0, 64)
  const uint32_t cc = c & 31u;
  return (int)((c & 32u) + (cc >> 1) + 16u * (cc & 1u));
}

struct BlockOut {
  std::vector<uint32_t> chunks;
  std::vector<uint64_t> bits;
  std::vector<uint32_t> ridx;
  std::vector<float> rval;
  uint32_t rcount[2 * kBtRows];
  int64_t tile_nnz = 0, unfactored = 0;
};

}  // namespace

int64_t bittile_schedule(const uint32_t *tiles_of_block, int64_t n_blk, int n_cta, int chunk_cols, int row_blocks,
                         std::vector<uint32_t> &cta_tile_ptr, std::vector<uint32_t> &cta_item_ptr, std::vector<uint2> &items,
                         std::vector<uint64_t> &tile_base) {
  std::vector<uint32_t> order;
  for (int64_t b = 0; b < n_blk; b++)
    if (tiles_of_block[b]) order.push_back((uint32_t)b);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tiles_of_block[a] > tiles_of_block[b]; });
  std::vector<std::vector<uint32_t>> per_cta((size_t)n_cta);
  {
    typedef std::pair<uint64_t, int> Load;  // (load, cta): smallest load first, ties by CTA index
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> pq;
    for (int q = 0; q < n_cta; q++) pq.push(Load(0, q));
    for (uint32_t b : order) {
      Load l = pq.top();
      pq.pop();
      per_cta[(size_t)l.second].push_back(b);
      l.first += ((size_t)tiles_of_block[b] + (size_t)(chunk_cols == 128 ? 3 : 6)) * (size_t)row_blocks;
      pq.push(l);
    }
  }
  cta_tile_ptr.assign((size_t)n_cta + 1, 0u);
  cta_item_ptr.assign((size_t)n_cta + 1, 0u);
  items.clear();
  tile_base.assign((size_t)n_blk, 0ull);
  uint64_t tiles = 0;
  for (int q = 0; q < n_cta; q++) {
    cta_tile_ptr[(size_t)q] = (uint32_t)tiles;
    cta_item_ptr[(size_t)q] = (uint32_t)items.size();
    uint32_t pos = 0;
    for (uint32_t b : per_cta[(size_t)q]) {
      tile_base[b] = tiles + pos;
      pos += tiles_of_block[b];
      items.push_back(make_uint2(b, pos));
    }
    tiles += pos;
    if (tiles > 0xfffffff0ull) return -1;
  }
  cta_tile_ptr[(size_t)n_cta] = (uint32_t)tiles;
  cta_item_ptr[(size_t)n_cta] = (uint32_t)items.size();
  return (int64_t)tiles;
}

int bittile_build_host(const uint32_t *indptr, const uint32_t *indices, const float *values, int64_t n_rows, int64_t n_cols,
                       const float *row_scale, const float *col_scale, int min_tile_nnz, int chunk_cols, int row_blocks,
                       int n_cta, int n_threads, BitTileHost &H) {
  if (chunk_cols == 0) chunk_cols = kBtChunk;
  if (row_blocks == 0) row_blocks = 1;
  if (chunk_cols != 64 && chunk_cols != 128) return GCNB_E_BADARG;
  if (row_blocks != 1 && !(row_blocks == 2 && chunk_cols == 64)) return GCNB_E_BADARG;
  const int64_t BH = (int64_t)kBtRows * row_blocks;  // rows per item
  if (!indptr || (!indices && indptr[n_rows] > 0) || n_rows < 0 || n_cols < 0 || n_rows > 0xfffffff0ll || n_cols > 0xfffffff0ll)
    return GCNB_E_BADARG;
  if ((row_scale == nullptr) != (col_scale == nullptr)) return GCNB_E_BADARG;
  // values == NULL (explicit scales only): a PATTERN -- every entry (i, j) has the value row_scale[i] * col_scale[j]
  if (!values && indptr[n_rows] > 0 && !row_scale) return GCNB_E_BADARG;
  H.n_rows = n_rows;
  H.n_cols = n_cols;
  H.nnz = indptr[n_rows];
  H.n_cta = n_cta > 0 ? n_cta : 148;
  H.chunk = chunk_cols;
  H.rb = row_blocks;
  H.min_tile_nnz = min_tile_nnz > 0 ? min_tile_nnz : 2 * chunk_cols * row_blocks;  // 1.6 % of the cells
  H.n_blk = (n_rows + BH - 1) / BH;
  const int64_t n_chunks = (n_cols + chunk_cols - 1) / chunk_cols;
  const int shift = chunk_cols == 128 ? 7 : 6;
  const size_t wpr = (size_t)chunk_cols / 64;  // bit-map words per row of a tile
  int T = n_threads > 0 ? n_threads : host_threads();
  T = (int)std::max<int64_t>(1, std::min<int64_t>(T, H.n_blk));

  const bool verbose = getenv("GCNB_SETUP_VERBOSE") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    const auto now = std::chrono::steady_clock::now();
    if (verbose) fprintf(stderr, "[bittile build] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
    tp = now;
  };
  // scales: given, or the square roots of the diagonal entries (GraphSum: value[i,i] = 1/deg_i, s_i = 1/sqrt(deg_i));
  // a row without a usable diagonal gets NaN, which fails every factorisation check => its entries stay in the remainder
  H.row_scale.assign((size_t)n_rows, 0.f);
  H.col_scale.assign((size_t)n_cols, 0.f);
  if (row_scale) {
    std::copy(row_scale, row_scale + n_rows, H.row_scale.begin());
    std::copy(col_scale, col_scale + n_cols, H.col_scale.begin());
  } else {
    const float nan = std::nanf("");
    std::fill(H.col_scale.begin(), H.col_scale.end(), nan);
    if (n_rows == n_cols) {
      bt_run_threads(T, [&](int t) {
        for (int64_t i = t; i < n_rows; i += T) {
          float s = nan;
          for (uint32_t e = indptr[i]; e < indptr[i + 1]; e++)
            if (indices[e] == (uint32_t)i) {
              if (values[e] > 0.f) s = sqrtf(values[e]);
              break;
            }
          H.col_scale[(size_t)i] = s;
        }
      });
    }
    H.row_scale = H.col_scale;
    H.row_scale.resize((size_t)n_rows, nan);
  }

  lap("scales");
  std::vector<BlockOut> blocks((size_t)H.n_blk);
  std::atomic<int64_t> next{0};
  const uint32_t thr = (uint32_t)H.min_tile_nnz;
  bt_run_threads(T, [&](int) {
    std::vector<uint32_t> cnt((size_t)n_chunks, 0u), touched;
    std::vector<int32_t> sel((size_t)n_chunks, -1);
    std::vector<uint32_t> sr_idx;  // remainder of the current block: thread-local scratch, copied out at its exact size
    std::vector<float> sr_val;     // (growing one vector per block by push_back made 16 threads fight over the allocator)
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= H.n_blk) break;
      BlockOut &o = blocks[(size_t)b];
      const int64_t r0 = b * BH, r1 = std::min<int64_t>(n_rows, r0 + BH);
      touched.clear();
      for (uint32_t e = indptr[r0]; e < indptr[r1]; e++) {
        const uint32_t c = indices[e] >> shift;
        if (cnt[c]++ == 0) touched.push_back(c);
      }
      std::sort(touched.begin(), touched.end());
      for (uint32_t c : touched) {
        if (cnt[c] >= thr) {
          sel[c] = (int32_t)o.chunks.size();
          o.chunks.push_back(c);
        }
      }
      o.bits.assign(o.chunks.size() * (size_t)BH * wpr, 0ull);
      const size_t block_entries = indptr[r1] - indptr[r0];
      if (sr_idx.size() < block_entries) {
        sr_idx.resize(block_entries);
        sr_val.resize(block_entries);
      }
      size_t n_rem = 0;
      for (int64_t i = r0; i < r1; i++) {
        const int rl = (int)(i - r0);
        uint32_t rc = 0;
        const float si = H.row_scale[(size_t)i];
        for (uint32_t e = indptr[i]; e < indptr[i + 1]; e++) {
          const uint32_t j = indices[e];
          const float p = si * H.col_scale[j];
          const float v = values ? values[e] : p;
          const int32_t li = sel[j >> shift];
          bool in_tile = false;
          const bool factors = fabsf(v - p) <= 1e-6f * fabsf(v);  // false for NaN scales
          if (!factors) o.unfactored++;
          if (li >= 0) {
            if (factors) {
              const uint32_t cc = j & (uint32_t)(chunk_cols - 1);
              uint64_t &w = o.bits[((size_t)li * BH + rl) * wpr + (cc >> 6)];  // rl = sub * 128 + row
              const uint64_t m = 1ull << bt_bit_of_col(cc & 63u);
              if (!(w & m)) {  // a duplicate entry cannot be a second bit: remainder
                w |= m;
                in_tile = true;
              }
            }
          }
          if (in_tile) {
            o.tile_nnz++;
          } else {
            sr_idx[n_rem] = j;
            sr_val[n_rem] = v;
            n_rem++;
            rc++;
          }
        }
        o.rcount[rl] = rc;
      }
      o.ridx.assign(sr_idx.begin(), sr_idx.begin() + (ptrdiff_t)n_rem);
      o.rval.assign(sr_val.begin(), sr_val.begin() + (ptrdiff_t)n_rem);
      for (uint32_t c : touched) {
        cnt[c] = 0;
        sel[c] = -1;
      }
    }
  });

  lap("tiles + remainder per block");
  // CTA schedule (bittile_schedule: shared with the device builder)
  std::vector<uint32_t> tiles_of_block((size_t)H.n_blk);
  for (int64_t b = 0; b < H.n_blk; b++) tiles_of_block[(size_t)b] = (uint32_t)blocks[(size_t)b].chunks.size();
  std::vector<uint64_t> tile_base;
  const int64_t tiles = bittile_schedule(tiles_of_block.data(), H.n_blk, H.n_cta, chunk_cols, row_blocks, H.cta_tile_ptr,
                                         H.cta_item_ptr, H.items, tile_base);
  if (tiles < 0) return GCNB_E_BADARG;
  H.n_tiles = (int64_t)tiles;
  H.tile_chunk.assign((size_t)tiles, 0u);
  H.bits.alloc((size_t)tiles * BH * wpr);

  // remainder CSR offsets
  H.r_indptr.assign((size_t)n_rows + 1, 0u);
  uint64_t racc = 0;
  H.tile_nnz = 0;
  H.n_unfactored = 0;
  for (int64_t b = 0; b < H.n_blk; b++) {
    const BlockOut &o = blocks[(size_t)b];
    const int64_t r0 = b * BH, r1 = std::min<int64_t>(n_rows, r0 + BH);
    for (int64_t i = r0; i < r1; i++) {
      H.r_indptr[(size_t)i] = (uint32_t)racc;
      racc += o.rcount[i - r0];
    }
    H.tile_nnz += o.tile_nnz;
    H.n_unfactored += o.unfactored;
  }
  H.r_indptr[(size_t)n_rows] = (uint32_t)racc;
  H.r_indices.alloc((size_t)racc);
  H.r_values.alloc((size_t)racc);
  lap("schedule + offsets");
  next = 0;
  bt_run_threads(T, [&](int) {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= H.n_blk) break;
      BlockOut &o = blocks[(size_t)b];
      if (!o.chunks.empty()) {
        std::copy(o.chunks.begin(), o.chunks.end(), H.tile_chunk.begin() + (ptrdiff_t)tile_base[(size_t)b]);
        memcpy(H.bits.data() + tile_base[(size_t)b] * BH * wpr, o.bits.data(), o.bits.size() * sizeof(uint64_t));
      }
      if (!o.ridx.empty()) {
        const size_t at = H.r_indptr[(size_t)(b * BH)];
        memcpy(H.r_indices.data() + at, o.ridx.data(), o.ridx.size() * 4);
        memcpy(H.r_values.data() + at, o.rval.data(), o.rval.size() * 4);
      }
      BlockOut().chunks.swap(o.chunks);
      std::vector<uint64_t>().swap(o.bits);
      std::vector<uint32_t>().swap(o.ridx);
      std::vector<float>().swap(o.rval);
    }
  });
  lap("flatten");
  // the kernels multiply by the scales unconditionally: rows / columns without a usable scale own no bit, give them 0
  for (float &x : H.row_scale)
    if (!(fabsf(x) <= 3.0e38f)) x = 0.f;
  for (float &x : H.col_scale)
    if (!(fabsf(x) <= 3.0e38f)) x = 0.f;
  return 0;
}

// =====================================================================================================================
// device
// =====================================================================================================================

// instruction descriptor of tcgen05.mma.kind::f16: D fp32, A and B bf16, both K-major, N = 48, M = 128
constexpr uint32_t kBtIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBtN >> 3) << 17) | ((uint32_t)(kBtRows >> 4) << 24);
// shared-memory matrix descriptor of one 48 x 16 bf16 operand, K-major, no swizzle: core matrices (8 rows x 16 bytes) of
// one k-half are contiguous (stride 128 bytes between 8-row groups = SBO), the second k-half follows 768 bytes later (LBO)
__device__ __forceinline__ uint64_t bt_b_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(768 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

struct BtArgs {
  const uint32_t *tile_chunk;
  const uint64_t *bits;
  const uint32_t *cta_tile_ptr;
  const uint32_t *cta_item_ptr;
  const uint2 *items;
  const uint8_t *packed;    // chunks of kBtChunkBytes
  const float *row_scale;
  float *P;                 // [n_blk * 128][16]: the tiles' partial rows (C == nullptr)
  float *C;                 // or: the product itself, zeroed by the pack kernel; partial rows are ADDED (row stride ldc)
  int64_t ldc;
  int64_t n_rows;
  const uint32_t *perm;     // optional: plan row / column k is row perm[k] of the caller's B and C (renumbered plans)
};

// B' = col_scale[j] * B[j][:] split into bf16 hi | mid | lo (truncation, exact sum) in the operand layout of the MMA:
// chunk c, k-step ks, element (n = piece*16 + col, k) at  c*6144 + ks*1536 + (k/8)*768 + (n/8)*128 + (n%8)*16 + (k%8)*2.
// One thread = 8 consecutive rows of B x one column: three 16-byte stores.
__global__ void __launch_bounds__(256) bt_pack_kernel(const float *__restrict__ B, const float *__restrict__ col_scale,
                                                      uint8_t *__restrict__ packed, float *__restrict__ B2, int64_t n_cols,
                                                      int64_t n_groups, float *__restrict__ Cz, int64_t ldc, int64_t n_rows,
                                                      const uint32_t *__restrict__ perm) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = tid >> 4;
  const int col = (int)(tid & 15);
  if (g >= n_groups) return;
  const int64_t j0 = g * 8;
  if (Cz)  // merge by reduction (bt_slab16): the product starts from zero
    for (int i = 0; i < 8; i++)
      if (j0 + i < n_rows) Cz[(j0 + i) * ldc + col] = 0.f;
  uint32_t hi[4], mid[4], lo[4];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t j = j0 + i;
    float x = 0.f;
    if (j < n_cols) {
      const int64_t jb = perm ? (int64_t)__ldg(perm + j) : j;  // renumbered plan: its column j is the caller's row perm[j]
      x = __ldg(col_scale + j) * __ldg(B + jb * 16 + col);
      if (B2) B2[j * 16 + col] = x;  // fp32 copy of B' for the pattern-only remainder kernel (spmm_ell.cu)
    }
    const uint32_t xb = __float_as_uint(x);
    const uint32_t hb = xb & 0xffff0000u;
    const float r1 = x - __uint_as_float(hb);
    const uint32_t mb = __float_as_uint(r1) & 0xffff0000u;
    const float r2 = r1 - __uint_as_float(mb);
    const uint32_t lb = __float_as_uint(r2) & 0xffff0000u;
    if (i & 1) {
      hi[i >> 1] |= hb;
      mid[i >> 1] |= mb;
      lo[i >> 1] |= lb;
    } else {
      hi[i >> 1] = hb >> 16;
      mid[i >> 1] = mb >> 16;
      lo[i >> 1] = lb >> 16;
    }
  }
  const int64_t c = j0 >> 6;
  const int kl = (int)(j0 & 63);
  uint8_t *base = packed + c * kBtChunkBytes + (kl >> 4) * kBtKStepBytes + ((kl >> 3) & 1) * 768;
  const int n0 = col;  // piece p: n = p*16 + col
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int n = p * 16 + n0;
    uint4 v;
    const uint32_t *src = p == 0 ? hi : (p == 1 ? mid : lo);
    v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3];
    *reinterpret_cast<uint4 *>(base + (n >> 3) * 128 + (n & 7) * 16) = v;
  }
}

// The same packing for a 16-column slab of a wider row-major matrix (row stride ldb floats)
__global__ void __launch_bounds__(256) bt_pack_ld_kernel(const float *__restrict__ B, int64_t ldb,
                                                         const float *__restrict__ col_scale, uint8_t *__restrict__ packed,
                                                         float *__restrict__ B2, int64_t n_cols, int64_t n_groups,
                                                         float *__restrict__ Cz, int64_t ldc, int64_t n_rows,
                                                         const uint32_t *__restrict__ perm) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t g = tid >> 4;
  const int col = (int)(tid & 15);
  if (g >= n_groups) return;
  const int64_t j0 = g * 8;
  if (Cz)
    for (int i = 0; i < 8; i++)
      if (j0 + i < n_rows) Cz[(j0 + i) * ldc + col] = 0.f;
  uint32_t hi[4], mid[4], lo[4];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t j = j0 + i;
    float x = 0.f;
    if (j < n_cols) {
      const int64_t jb = perm ? (int64_t)__ldg(perm + j) : j;
      x = __ldg(col_scale + j) * __ldg(B + jb * ldb + col);
      if (B2) B2[j * 16 + col] = x;
    }
    const uint32_t xb = __float_as_uint(x);
    const uint32_t hb = xb & 0xffff0000u;
    const float r1 = x - __uint_as_float(hb);
    const uint32_t mb = __float_as_uint(r1) & 0xffff0000u;
    const float r2 = r1 - __uint_as_float(mb);
    const uint32_t lb = __float_as_uint(r2) & 0xffff0000u;
    if (i & 1) {
      hi[i >> 1] |= hb;
      mid[i >> 1] |= mb;
      lo[i >> 1] |= lb;
    } else {
      hi[i >> 1] = hb >> 16;
      mid[i >> 1] = mb >> 16;
      lo[i >> 1] = lb >> 16;
    }
  }
  const int64_t c = j0 >> 6;
  const int kl = (int)(j0 & 63);
  uint8_t *base = packed + c * kBtChunkBytes + (kl >> 4) * kBtKStepBytes + ((kl >> 3) & 1) * 768;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int n = p * 16 + col;
    uint4 v;
    const uint32_t *src = p == 0 ? hi : (p == 1 ? mid : lo);
    v.x = src[0]; v.y = src[1]; v.z = src[2]; v.w = src[3];
    *reinterpret_cast<uint4 *>(base + (n >> 3) * 128 + (n & 7) * 16) = v;
  }
}

// C[:, 0..16) (row stride ldc) = P + R
// (plan row k is row perm[k] of C for a renumbered plan)
__global__ void __launch_bounds__(256) bt_add_ld_kernel(const float *__restrict__ P, const float *__restrict__ R,
                                                        float *__restrict__ C, int64_t ldc, int64_t n,
                                                        const uint32_t *__restrict__ perm) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = perm ? (int64_t)__ldg(perm + (i >> 4)) : (i >> 4);
    C[row * ldc + (i & 15)] = P[i] + R[i];
  }
}

__device__ __forceinline__ uint64_t bt_ld_bits(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// 32 bits -> 16 registers of two bf16 (0.0 or 1.0) each; bit q and bit q+16 of the word are the pair of register q
__device__ __forceinline__ void bt_expand_word(uint32_t w, uint32_t *r) {
  const uint32_t w8 = w >> 8;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    r[q] = (w & (0x00010001u << q)) * (0x3F80u >> q);
    r[q + 8] = (w8 & (0x00010001u << q)) * (0x3F80u >> q);
  }
}

// ---- the MMA kernel ---------------------------------------------------------------------------------------------------
// Round 1's first kernel (separate A / B rings, a single issuing thread in a divergent branch; deleted) spent ~650 clocks
// per 64-column tile against 96 of tensor time: 303 us on the bench graph.  This one (a) unifies the rings: stage s = B' chunk in shared memory + A operand in TMEM, ONE
// full[s] barrier (4 expander-warp arrivals + the producer's arrive.expect_tx + the copy's bytes) and ONE free[s]
// barrier (one commit) per tile; (b) takes tiles of W * 64 columns (W = 2: 8 MMAs, 12 KB of B', 128 bits per row per
// tile), halving the round trips per unit of work again.  Two accumulators per set (the measured error of 73-MMA chains
// is 0.5 ulp: the TMEM accumulate does not drift), two sets so the epilogue overlaps the next row block.
// TMEM: stages 8/W x 32W columns = 256, accumulators 2 sets x 2 x 48 = 192.
template <int RB, int W>
struct BtWide {
  static_assert(RB == 1 || (RB == 2 && W == 1), "256-row items take 64-column tiles");
  static constexpr int kStages = RB == 1 ? 8 / W : 5;
  static constexpr int kACols = 32 * W * RB;           // TMEM columns of one stage's A operand(s)
  static constexpr int kKSteps = 4 * W;
  static constexpr int kChunkBytes = kKSteps * kBtKStepBytes;
  static constexpr int kAcc = RB == 1 ? 2 : 1;         // accumulators per (set, 128-row half)
  static constexpr int kAccCol0 = kStages * kACols;    // 256 / 320
  static constexpr int kSetCols = RB * kAcc * kBtN;    // 96
  static_assert(kAccCol0 + 2 * kSetCols <= 512, "TMEM columns");
  static constexpr size_t kSmemBytes = (size_t)kStages * kChunkBytes + (2 * kStages + 4) * 8 + 16;
};

// RB = 2: an item is 256 rows, a tile 256 x 64 bits; the two 128-row halves are two MMAs (two TMEM accumulators, two A
// operands in the stage) against the SAME B' stage, which halves the B' traffic through L2 and lets the plan select
// tiles by the density of 256 x 64 cells.  One accumulator per half and set (chains of ~300 MMAs; measured: no drift).
template <int RB, int W>
__global__ void __maxnreg__(64) bt_mma_wide_kernel(BtArgs a) {  // 448 x 64 registers: two 256-thread x 64-register remainder CTAs stay co-resident (at 72 they do not: +60 us)
  using K = BtWide<RB, W>;
  extern __shared__ __align__(128) uint8_t bt_smem[];
  uint8_t *smem_b = bt_smem;
  uint64_t *bars = reinterpret_cast<uint64_t *>(bt_smem + K::kStages * K::kChunkBytes);
  uint64_t *full = bars, *free_ = bars + K::kStages;
  uint64_t *acc_full = free_ + K::kStages, *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x;
  const uint32_t tile0 = a.cta_tile_ptr[q];
  const uint32_t T = a.cta_tile_ptr[q + 1] - tile0;
  const uint32_t item0 = a.cta_item_ptr[q], item1 = a.cta_item_ptr[q + 1];

  if (threadIdx.x == 0) {
    for (int i = 0; i < K::kStages; i++) {
      mbar_init(&full[i], 5);   // 4 expander warps + the producer's arrive.expect_tx
      mbar_init(&free_[i], 1);  // one tcgen05.commit
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

  if (warp < 8) {
    // ---- expanders: group g = warp >> 2 takes tiles t = g, g+2, ...; thread = row of each 128-row half
    constexpr int NW = RB * W;                                  // 64-bit words per thread per tile
    constexpr size_t kHalfWords = (size_t)kBtRows * W;          // words between the two halves of a tile
    constexpr size_t kTileWords = (size_t)RB * kHalfWords;
    const int quarter = warp & 3, g = warp >> 2;
    const uint64_t *bp = a.bits + (size_t)tile0 * kTileWords + (size_t)(quarter * 32 + lane) * W;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    auto word_at = [&](uint32_t t, int i) {  // i = half * W + w
      return bt_ld_bits(bp + (size_t)t * kTileWords + (size_t)(i / W) * kHalfWords + (i % W));
    };
    uint64_t w0[NW], w1[NW], w2[NW];
    uint32_t t = (uint32_t)g;
#pragma unroll
    for (int i = 0; i < NW; i++) {
      w0[i] = t < T ? word_at(t, i) : 0ull;
      w1[i] = t + 2 < T ? word_at(t + 2, i) : 0ull;
      w2[i] = t + 4 < T ? word_at(t + 4, i) : 0ull;
    }
    for (; t < T; t += 2) {
      uint64_t w3[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) w3[i] = t + 6 < T ? word_at(t + 6, i) : 0ull;
      const uint32_t s = t % K::kStages, use = t / K::kStages;
      if (use > 0) mbar_wait(&free_[s], (use - 1) & 1);  // the MMAs that read this stage's previous content are done
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < NW; i++) {
        uint32_t r[32];
        bt_expand_word((uint32_t)w0[i], r);
        bt_expand_word((uint32_t)(w0[i] >> 32), r + 16);
        tc_st32(tmem + lane_base + s * K::kACols + i * 32, r);  // half h occupies columns [h * 32 W, (h + 1) * 32 W)
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
#pragma unroll
      for (int i = 0; i < NW; i++) {
        w0[i] = w1[i];
        w1[i] = w2[i];
        w2[i] = w3[i];
      }
    }
  } else if (warp < 12) {
    // ---- epilogue: thread = row (TMEM lane) of each 128-row half
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    uint32_t prev_end = 0;
    for (uint32_t it = item0; it < item1; it++) {
      const uint2 item = a.items[it];
      const uint32_t n_tiles = item.y - prev_end;
      prev_end = item.y;
      const uint32_t k = it - item0, set = k & 1, use = k >> 1;
      mbar_wait(&acc_full[set], use & 1);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < RB; h++) {
        const int64_t row = ((int64_t)item.x * RB + h) * kBtRows + quarter * 32 + lane;
        const float sc = row < a.n_rows ? __ldg(a.row_scale + row) : 0.f;
        const uint32_t acc0 = tmem + lane_base + K::kAccCol0 + set * K::kSetCols + h * (K::kAcc * kBtN);
        float tot[16];
#pragma unroll
        for (int p = 2; p >= 0; p--) {  // lo, + mid, + hi
          float s0[16];
          tc_ld16(acc0 + p * 16, s0);
          if (K::kAcc > 1 && n_tiles > 1) {
            float s1[16];
            tc_ld16(acc0 + kBtN + p * 16, s1);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i++) s0[i] += s1[i];
          } else {
            tc_wait_ld();
          }
#pragma unroll
          for (int i = 0; i < 16; i++) tot[i] = p == 2 ? s0[i] : tot[i] + s0[i];
        }
        if (a.C) {  // merge by reduction: whichever of this kernel and the remainder kernel finishes a row first, 0 + p + r
          if (row < a.n_rows) {
            float *dst = a.C + (a.perm ? (int64_t)__ldg(a.perm + row) : row) * a.ldc;
#pragma unroll
            for (int i = 0; i < 4; i++)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * i), "f"(sc * tot[4 * i]),
                           "f"(sc * tot[4 * i + 1]), "f"(sc * tot[4 * i + 2]), "f"(sc * tot[4 * i + 3])
                           : "memory");
          }
        } else {
          float4 *dst = reinterpret_cast<float4 *>(a.P + row * 16);
#pragma unroll
          for (int i = 0; i < 4; i++)
            dst[i] = make_float4(sc * tot[4 * i], sc * tot[4 * i + 1], sc * tot[4 * i + 2], sc * tot[4 * i + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  } else if (warp == 12) {
    // ---- producer of B'
    for (uint32_t base = 0; base < T; base += 32) {
      const uint32_t mine = base + lane < T ? __ldg(a.tile_chunk + tile0 + base + lane) : 0u;
      const uint32_t cnt = min(32u, T - base);
      for (uint32_t k = 0; k < cnt; k++) {
        const uint32_t c = __shfl_sync(0xffffffffu, mine, (int)k);
        if (lane == 0) {
          const uint32_t t = base + k, s = t % K::kStages, use = t / K::kStages;
          if (use > 0) mbar_wait(&free_[s], (use - 1) & 1);
          mbar_expect_tx(&full[s], K::kChunkBytes);
          bulk_g2s(smem_b + s * K::kChunkBytes, a.packed + (size_t)c * K::kChunkBytes, K::kChunkBytes, &full[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---- MMA issuer: one wait, RB * 4W MMAs, one commit per tile.  The whole warp walks the loop (warp-uniform control
    // flow keeps addresses and descriptors in uniform registers); one elected lane issues.
    uint32_t t = 0;
    uint2 item = item0 < item1 ? a.items[item0] : make_uint2(0, 0);
    for (uint32_t it = item0; it < item1; it++) {
      const uint2 next_item = it + 1 < item1 ? a.items[it + 1] : make_uint2(0, 0);
      const uint32_t k = it - item0, set = k & 1, use = k >> 1;
      if (use > 0) mbar_wait(&acc_empty[set], (use - 1) & 1);
      tc_fence_after();
      const uint32_t t_begin = t;
      for (; t < item.y; t++) {
        const uint32_t s = t % K::kStages, u = t / K::kStages;
        mbar_wait(&full[s], u & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idx = t - t_begin;
          const uint64_t bd = bt_b_desc(smem_u32(smem_b + s * K::kChunkBytes));
          const uint32_t acc_flag = idx >= (uint32_t)K::kAcc ? 1u : 0u;
#pragma unroll
          for (int h = 0; h < RB; h++) {
            const uint32_t d = tmem + K::kAccCol0 + set * K::kSetCols + (h * K::kAcc + idx % K::kAcc) * kBtN;
            const uint32_t a0 = tmem + s * K::kACols + h * (32 * W);
#pragma unroll
            for (int ks = 0; ks < K::kKSteps; ks++)  // the next k-step of B' is 1536 bytes = 96 descriptor units further
              tc_mma_ts(d, a0 + ks * 8, bd + (uint64_t)(ks * (kBtKStepBytes >> 4)), kBtIdesc, ks > 0 ? 1u : acc_flag);
          }
          tc_commit(&free_[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&acc_full[set]);
      __syncwarp();
      item = next_item;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

__global__ void __launch_bounds__(256) bt_add_kernel(const float4 *__restrict__ P, const float4 *__restrict__ R,
                                                     float4 *__restrict__ C, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = P[i], r = R[i];
    C[i] = make_float4(p.x + r.x, p.y + r.y, p.z + r.z, p.w + r.w);
  }
}

}  // namespace gcnb

// =====================================================================================================================
// C ABI
// =====================================================================================================================
struct gcnb_bittile_host {
  BitTileHost H;
};


namespace {
template <class T>
int bt_upload(T **dst, const T *src, size_t n, cudaStream_t stream) {
  *dst = nullptr;
  GCNB_CHECK(cudaMalloc((void **)dst, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) GCNB_CHECK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}
}  // namespace

namespace gcnb {
int bittile_finish_plan(gcnb_bittile_plan *p, cudaStream_t stream) {
  const bool verbose = getenv("GCNB_SETUP_VERBOSE") != nullptr;
  auto tp = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[bittile finish] %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tp).count());
    tp = now;
  };
  const size_t packed_bytes = std::max<size_t>((size_t)p->n_chunks * kBtChunkBytes, 16);
  const size_t p_bytes = std::max<size_t>((size_t)p->n_blk * p->rb * kBtRows * 16 * sizeof(float), 16);
  GCNB_CHECK(cudaMalloc((void **)&p->d_packed, packed_bytes));
  GCNB_CHECK(cudaMalloc((void **)&p->d_P, p_bytes));
  GCNB_CHECK(cudaMalloc((void **)&p->d_R, std::max<size_t>((size_t)p->n_rows * 16 * sizeof(float), 16)));
  lap("operand / partial buffers");
  GCNB_CHECK(cudaMemsetAsync(p->d_P, 0, p_bytes, stream));  // blocks without tiles stay 0 for ever
  GCNB_CHECK(cudaStreamSynchronize(stream));
  lap("synchronise");
  // tuning probe: cap the remainder kernel's CTAs per SM so that, whichever kernel the block scheduler sees first, the MMA
  // kernel's CTA (448 threads x 68 registers) still fits on every SM (first measurements: launched at the same instant the
  // two kernels took 772 us instead of 502)
  if (const char *e = getenv("GCNB_BT_MERGE")) p->merge_by_reduction = atoi(e) != 0;
  if (const char *e = getenv("GCNB_BT_REM_CTAS")) p->rem_ctas = std::max(0, atoi(e));
  GCNB_CHECK(cudaStreamCreateWithFlags(&p->aux, cudaStreamNonBlocking));
  GCNB_CHECK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  GCNB_CHECK(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  lap("stream + events");
  GCNB_CHECK(cudaFuncSetAttribute(bt_mma_wide_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BtWide<1, 1>::kSmemBytes));
  GCNB_CHECK(cudaFuncSetAttribute(bt_mma_wide_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BtWide<1, 2>::kSmemBytes));
  GCNB_CHECK(cudaFuncSetAttribute(bt_mma_wide_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BtWide<2, 1>::kSmemBytes));
  lap("kernel attributes");
  return 0;
}
}  // namespace gcnb

extern "C" {

int gcnb_bittile_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, const float *h_values, int64_t n_rows,
                            int64_t n_cols, const float *h_row_scale, const float *h_col_scale, int min_tile_nnz,
                            int chunk_cols, int row_blocks, int n_cta, int n_threads, gcnb_bittile_host **out) {
  if (!out) return GCNB_E_BADARG;
  auto *h = new gcnb_bittile_host();
  const int rc = bittile_build_host(h_indptr, h_indices, h_values, n_rows, n_cols, h_row_scale, h_col_scale, min_tile_nnz,
                                    chunk_cols, row_blocks, n_cta, n_threads, h->H);
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int gcnb_bittile_host_sizes(const gcnb_bittile_host *h, int64_t out[12]) {
  if (!h || !out) return GCNB_E_BADARG;
  const BitTileHost &H = h->H;
  out[0] = H.n_rows; out[1] = H.n_cols; out[2] = H.nnz; out[3] = H.n_blk; out[4] = H.n_tiles; out[5] = H.tile_nnz;
  out[6] = (int64_t)H.items.size(); out[7] = H.n_cta; out[8] = H.chunk; out[9] = H.rb;
  out[10] = H.n_unfactored; out[11] = (int64_t)H.r_indices.size();
  return 0;
}

// which: 0 tile_chunk, 1 bits (uint64), 2 cta_tile_ptr, 3 cta_item_ptr, 4 items (uint32 x2), 5 r_indptr, 6 r_indices,
// 7 r_values, 8 row_scale, 9 col_scale.  Copies min(bytes, size) bytes.
int gcnb_bittile_host_copy(const gcnb_bittile_host *h, int which, void *dst, int64_t bytes) {
  if (!h || !dst || bytes < 0) return GCNB_E_BADARG;
  const BitTileHost &H = h->H;
  const void *src = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: src = H.tile_chunk.data(); n = H.tile_chunk.size() * 4; break;
    case 1: src = H.bits.data(); n = H.bits.size() * 8; break;
    case 2: src = H.cta_tile_ptr.data(); n = H.cta_tile_ptr.size() * 4; break;
    case 3: src = H.cta_item_ptr.data(); n = H.cta_item_ptr.size() * 4; break;
    case 4: src = H.items.data(); n = H.items.size() * sizeof(uint2); break;
    case 5: src = H.r_indptr.data(); n = H.r_indptr.size() * 4; break;
    case 6: src = H.r_indices.data(); n = H.r_indices.size() * 4; break;
    case 7: src = H.r_values.data(); n = H.r_values.size() * 4; break;
    case 8: src = H.row_scale.data(); n = H.row_scale.size() * 4; break;
    case 9: src = H.col_scale.data(); n = H.col_scale.size() * 4; break;
    default: return GCNB_E_BADARG;
  }
  if (n) memcpy(dst, src, std::min<size_t>(n, (size_t)bytes));
  return 0;
}

int gcnb_bittile_host_destroy(gcnb_bittile_host *h) {
  delete h;
  return 0;
}

// 1 when the current device can run the bit-tile kernels (tcgen05 / TMEM: compute capability 10.x)
int gcnb_bittile_supported(void) {
  const DeviceInfo &di = device_info();
  return di.ok && di.cc_major == 10;
}

int gcnb_bittile_plan_destroy(gcnb_bittile_plan *p) {
  if (!p) return 0;
  if (p->rem) gcnb_spmm_plan_destroy(p->rem);
  gcnb::ell_destroy(p->ell);
  cudaFree(p->d_B2);
  cudaFree(p->d_perm);
  cudaFree(p->d_tile_chunk); cudaFree(p->d_cta_tile_ptr); cudaFree(p->d_cta_item_ptr); cudaFree(p->d_items);
  cudaFree(p->d_bits); cudaFree(p->d_r_indptr); cudaFree(p->d_r_indices); cudaFree(p->d_r_values);
  cudaFree(p->d_row_scale); cudaFree(p->d_col_scale); cudaFree(p->d_packed); cudaFree(p->d_P); cudaFree(p->d_R);
  if (p->aux) cudaStreamDestroy(p->aux);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  delete p;
  return 0;
}

int gcnb_bittile_plan_create(const uint32_t *h_indptr, const uint32_t *h_indices, const float *h_values, int64_t n_rows,
                             int64_t n_cols, const float *h_row_scale, const float *h_col_scale, int min_tile_nnz,
                             int chunk_cols, int row_blocks, gcnb_stream_t stream_, gcnb_bittile_plan **out) {
  if (!out) return GCNB_E_BADARG;
  *out = nullptr;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  if (di.cc_major != 10) return GCNB_E_UNSUPPORTED;  // tcgen05 / TMEM
  cudaStream_t stream = as_stream(stream_);
  BitTileHost H;
  if (chunk_cols == 0)
    if (const char *e = getenv("GCNB_BT_CHUNK")) chunk_cols = atoi(e);    // tuning probe: 64 or 128
  if (row_blocks == 0)
    if (const char *e = getenv("GCNB_BT_RB")) row_blocks = atoi(e);       // tuning probe: 1 or 2
  // default shape: items of 256 rows x 64 columns whose two halves share a B' stage (B200, bench graph, GraphSum in the
  // bench step: 235 us; 128 x 128 tiles 242 us; alone the MMA kernels take 131 / 126 us, 128 x 64 tiles 190 us)
  if (chunk_cols == 0 && row_blocks == 0) {
    chunk_cols = 64;
    row_blocks = 2;
  }
  if (chunk_cols == 0) chunk_cols = row_blocks == 2 ? 64 : 128;
  int rc = bittile_build_host(h_indptr, h_indices, h_values, n_rows, n_cols, h_row_scale, h_col_scale, min_tile_nnz,
                              chunk_cols, row_blocks, di.sm_count, 0, H);
  if (rc) return rc;
  auto *p = new gcnb_bittile_plan();
  auto fail = [&](int code) {
    gcnb_bittile_plan_destroy(p);
    return code;
  };
  p->n_rows = n_rows; p->n_cols = n_cols; p->nnz = H.nnz; p->n_blk = H.n_blk; p->n_tiles = H.n_tiles;
  p->tile_nnz = H.tile_nnz; p->rem_nnz = (int64_t)H.r_indices.size(); p->n_cta = H.n_cta;
  p->n_unfactored = H.n_unfactored;
  p->chunk = H.chunk;
  p->rb = H.rb;
  p->n_chunks = (n_cols + 127) / 128 * 2;  // 64-row units of the packed B' image, padded to whole 128-column chunks
  if ((rc = bt_upload(&p->d_tile_chunk, H.tile_chunk.data(), H.tile_chunk.size(), stream))) return fail(rc);
  if ((rc = bt_upload(&p->d_bits, H.bits.data(), H.bits.size(), stream))) return fail(rc);
  if ((rc = bt_upload(&p->d_cta_tile_ptr, H.cta_tile_ptr.data(), H.cta_tile_ptr.size(), stream))) return fail(rc);
  if ((rc = bt_upload(&p->d_cta_item_ptr, H.cta_item_ptr.data(), H.cta_item_ptr.size(), stream))) return fail(rc);
  if ((rc = bt_upload(&p->d_items, H.items.data(), H.items.size(), stream))) return fail(rc);
  bool use_ell = H.n_unfactored == 0 && H.n_tiles > 0;
  if (const char *e = getenv("GCNB_BT_ELL")) use_ell = use_ell && atoi(e) != 0;  // tuning probe: 0 keeps the valued generic kernel
  if (use_ell) {
    gcnb::EllHost E;
    if ((rc = gcnb::ell_build_host(H.r_indptr.data(), H.r_indices.data(), n_rows, n_cols, 0, E))) return fail(rc);
    if ((rc = gcnb::ell_upload_plan(E, stream, &p->ell))) return fail(rc);
    const size_t b2_bytes = ((size_t)n_cols + 1) * 16 * sizeof(float);
    if ((rc = (int)cudaMalloc((void **)&p->d_B2, b2_bytes))) return fail(rc);
    if ((rc = (int)cudaMemsetAsync(p->d_B2, 0, b2_bytes, stream))) return fail(rc);  // the padding row stays zero
  } else {
    if ((rc = bt_upload(&p->d_r_indptr, H.r_indptr.data(), H.r_indptr.size(), stream))) return fail(rc);
    if ((rc = bt_upload(&p->d_r_indices, H.r_indices.data(), H.r_indices.size(), stream))) return fail(rc);
    if ((rc = bt_upload(&p->d_r_values, H.r_values.data(), H.r_values.size(), stream))) return fail(rc);
  }
  if ((rc = bt_upload(&p->d_row_scale, H.row_scale.data(), H.row_scale.size(), stream))) return fail(rc);
  if ((rc = bt_upload(&p->d_col_scale, H.col_scale.data(), H.col_scale.size(), stream))) return fail(rc);
  if ((rc = bittile_finish_plan(p, stream))) return fail(rc);  // (synchronises: the host arrays go out of scope)
  if (!p->ell) {
    if ((rc = gcnb_spmm_plan_create(p->d_r_indptr, p->d_r_indices, n_rows, n_cols, 0, stream_, &p->rem))) return fail(rc);
    p->rem->max_cta_per_sm = p->rem_ctas;
  }
  *out = p;
  return 0;
}

// out = {tiles, entries in tiles, remainder entries, row blocks, items (blocks with tiles), CTAs, bit-map bytes, packed B' bytes}
int gcnb_bittile_plan_info(const gcnb_bittile_plan *p, int64_t out[8]) {
  if (!p || !out) return GCNB_E_BADARG;
  out[0] = p->n_tiles; out[1] = p->tile_nnz; out[2] = p->rem_nnz; out[3] = p->n_blk; out[4] = p->chunk + 1000 * p->rb + (p->ell ? 100000 : 0);
  out[5] = p->n_cta;
  out[6] = p->n_tiles * (int64_t)kBtRows * p->rb * (p->chunk / 8); out[7] = p->n_chunks * (int64_t)kBtChunkBytes;
  return 0;
}

// entries of the matrix whose value is not row_scale * col_scale (0: a renumbered, pattern-only plan may replace this one)
int64_t gcnb_bittile_plan_unfactored(const gcnb_bittile_plan *p) { return p ? p->n_unfactored : -1; }

// A plan built from a RENUMBERED matrix (rows and columns permuted alike, e.g. community by community): plan index k is row
// h_old_of_new[k] of the caller's operands.  The pack kernel gathers B through it and both halves of the product add their
// rows to C through it, so the caller never sees the renumbering.  Square plans whose remainder is the ELL kernel only.
int gcnb_bittile_plan_set_permutation(gcnb_bittile_plan *p, const uint32_t *h_old_of_new, gcnb_stream_t stream_) {
  if (!p || !h_old_of_new) return GCNB_E_BADARG;
  if (p->n_rows != p->n_cols || !p->ell || p->n_tiles == 0 || !p->merge_by_reduction) return GCNB_E_UNSUPPORTED;
  std::vector<uint8_t> seen((size_t)p->n_rows, 0);
  for (int64_t k = 0; k < p->n_rows; k++) {
    if (h_old_of_new[k] >= (uint64_t)p->n_rows || seen[h_old_of_new[k]]) return GCNB_E_BADARG;
    seen[h_old_of_new[k]] = 1;
  }
  cudaStream_t stream = as_stream(stream_);
  if (!p->d_perm) GCNB_CHECK(cudaMalloc((void **)&p->d_perm, std::max<size_t>((size_t)p->n_rows, 1) * 4));
  GCNB_CHECK(cudaMemcpyAsync(p->d_perm, h_old_of_new, (size_t)p->n_rows * 4, cudaMemcpyHostToDevice, stream));
  GCNB_CHECK(cudaStreamSynchronize(stream));
  return 0;
}

// kernels launched per 16-column product on aligned operands: pack, MMA kernel, remainder (+ its combine kernel when rows are
// cut), and the final add unless the two halves are merged by reduction
int gcnb_bittile_plan_launches(const gcnb_bittile_plan *p) {
  if (!p) return 0;
  const bool merge = p->n_tiles > 0 && p->ell && p->merge_by_reduction;
  int n = (p->n_tiles > 0 ? 2 : 0) + 1 + (merge ? 0 : 1);
  if (p->ell && p->ell->n_split > 0) n++;
  if (p->rem && p->rem->n_split_rows > 0) n++;
  return n;
}

// debugging aid: switch steps of gcnb_bittile_spmm16_f32 off (bit 0 pack, 1 MMA kernel, 2 remainder, 3 final add) to time them apart
int gcnb_bittile_debug_parts(gcnb_bittile_plan *p, int parts) {
  if (!p) return GCNB_E_BADARG;
  p->parts = parts & 15;
  return 0;
}

// debugging aid: run the pack kernel alone and copy the operand image of B' back to the host
int gcnb_bittile_debug_pack(gcnb_bittile_plan *p, const float *d_B, void *h_out, int64_t bytes, gcnb_stream_t stream_) {
  if (!p || !d_B || !h_out || bytes < 0) return GCNB_E_BADARG;
  cudaStream_t stream = as_stream(stream_);
  const int64_t n_groups = p->n_chunks * (kBtChunk / 8);
  if (n_groups == 0) return 0;
  bt_pack_kernel<<<(unsigned)((n_groups * 16 + 255) / 256), 256, 0, stream>>>(d_B, p->d_col_scale, p->d_packed, nullptr,
                                                                              p->n_cols, n_groups, nullptr, 0, 0, nullptr);
  GCNB_LAUNCH_CHECK();
  GCNB_CHECK(cudaMemcpyAsync(h_out, p->d_packed, (size_t)std::min<int64_t>(bytes, p->n_chunks * (int64_t)kBtChunkBytes),
                             cudaMemcpyDeviceToHost, stream));
  GCNB_CHECK(cudaStreamSynchronize(stream));
  return 0;
}

// one 16-column slab: B row stride ldb, C row stride ldc (16 / 16: the contiguous kernels measured in round 1)
static int bt_slab16(gcnb_bittile_plan *p, const float *d_B, int64_t ldb, float *d_C, int64_t ldc, cudaStream_t stream) {
  const int parts = p->parts;  // 15 unless a probe switched steps off (gcnb_bittile_debug_parts)
  const bool tiles = p->n_tiles > 0;
  const bool contiguous = ldb == 16 && ldc == 16 && (((uintptr_t)d_B | (uintptr_t)d_C) % 16 == 0);
  const int64_t n_groups = p->n_chunks * (kBtChunk / 8);
  // Merge by reduction (the default whenever the remainder is the pattern-only ELL kernel): the pack kernel zeroes C, then
  // the MMA kernel's epilogue and the remainder kernel each ADD their half of a row (REDG.ADD.F32x4).  0 + a + b == 0 + b + a
  // in floating point, so the result does not depend on which finishes first; no partial buffers, no final add kernel.
  // Otherwise (valued remainder, unaligned slabs, C overlapping B, probes): partial buffers P and R + bt_add_kernel.
  const bool overlap = d_C < d_B + (p->n_cols - 1) * ldb + 16 && d_B < d_C + (p->n_rows - 1) * ldc + 16;
  const bool merge = tiles && p->ell && parts == 15 && p->merge_by_reduction && (uintptr_t)d_C % 16 == 0 && ldc % 4 == 0 &&
                     p->n_rows <= n_groups * 8 && !overlap;
  if (tiles && (parts & 1)) {
    const int64_t threads = n_groups * 16;
    float *cz = merge ? d_C : nullptr;
    if (ldb == 16)
      bt_pack_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_B, p->d_col_scale, p->d_packed, p->d_B2,
                                                                           p->n_cols, n_groups, cz, ldc, p->n_rows, p->d_perm);
    else
      bt_pack_ld_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_B, ldb, p->d_col_scale, p->d_packed,
                                                                              p->d_B2, p->n_cols, n_groups, cz, ldc, p->n_rows,
                                                                              p->d_perm);
    GCNB_LAUNCH_CHECK();
  }
  // The MMA kernel goes first (one CTA per SM, half the register file), then the remainder on the second stream
  // fills what is left of every SM: launched the other way round the remainder's persistent CTAs own all registers
  // and the two kernels run one after the other.
  GCNB_CHECK(cudaEventRecord(p->ev_fork, stream));
  if (tiles && (parts & 2)) {
    BtArgs a;
    a.tile_chunk = p->d_tile_chunk; a.bits = p->d_bits; a.cta_tile_ptr = p->d_cta_tile_ptr;
    a.cta_item_ptr = p->d_cta_item_ptr; a.items = p->d_items; a.packed = p->d_packed; a.row_scale = p->d_row_scale;
    a.P = p->d_P; a.C = merge ? d_C : nullptr; a.ldc = ldc; a.n_rows = p->n_rows; a.perm = p->d_perm;
    if (p->rb == 2) bt_mma_wide_kernel<2, 1><<<p->n_cta, kBtThreads, BtWide<2, 1>::kSmemBytes, stream>>>(a);
    else if (p->chunk == 128) bt_mma_wide_kernel<1, 2><<<p->n_cta, kBtThreads, BtWide<1, 2>::kSmemBytes, stream>>>(a);
    else bt_mma_wide_kernel<1, 1><<<p->n_cta, kBtThreads, BtWide<1, 1>::kSmemBytes, stream>>>(a);
    GCNB_LAUNCH_CHECK();
  }
  GCNB_CHECK(cudaStreamWaitEvent(p->aux, p->ev_fork, 0));
  if (parts & 4) {
    const int rc = p->ell ? gcnb::ell_launch(p->ell, p->d_B2, p->d_row_scale, merge ? d_C : p->d_R, merge ? ldc : 16, merge ? 1 : 0,
                                             merge ? p->d_perm : nullptr, p->rem_ctas, p->aux)
                          : gcnb::spmm_generic_launch(p->rem, p->d_r_values, nullptr, d_B, (int)ldb, p->d_R, 16, 16, p->aux);
    if (rc) return rc;
  }
  GCNB_CHECK(cudaEventRecord(p->ev_join, p->aux));
  GCNB_CHECK(cudaStreamWaitEvent(stream, p->ev_join, 0));
  if ((parts & 8) && !merge) {
    if (contiguous && !p->d_perm) {
      const int64_t n4 = p->n_rows * 4;
      const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)device_info().sm_count * 8);
      bt_add_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4 *>(p->d_P),
                                               reinterpret_cast<const float4 *>(p->d_R), reinterpret_cast<float4 *>(d_C), n4);
    } else {
      const int64_t n = p->n_rows * 16;
      const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)device_info().sm_count * 16);
      bt_add_ld_kernel<<<blocks, 256, 0, stream>>>(p->d_P, p->d_R, d_C, ldc, n, p->d_perm);
    }
  }
  GCNB_LAUNCH_CHECK();
  return 0;
}

int gcnb_bittile_spmm16_f32(gcnb_bittile_plan *p, const float *d_B, float *d_C, gcnb_stream_t stream_) {
  if (!p || !d_B || !d_C) return GCNB_E_BADARG;
  if (p->n_rows == 0) return 0;
  return bt_slab16(p, d_B, 16, d_C, 16, as_stream(stream_));
}

// C[:, 0..dim) = A * B[:, 0..dim) on column slabs of wider row-major matrices (row strides ldb / ldc floats, dim >= 16):
// 16 columns at a time, the last slab shifted left so that it ends at dim (its overlap is computed twice, identically)
int gcnb_bittile_spmm_ld_f32(gcnb_bittile_plan *p, const float *d_B, int64_t ldb, float *d_C, int64_t ldc, int dim,
                             gcnb_stream_t stream_) {
  if (!p || !d_B || !d_C || dim < 16 || ldb < dim || ldc < dim || ldb > (1 << 28) || ldc > (1 << 28)) return GCNB_E_BADARG;
  if (p->n_rows == 0) return 0;
  for (int c0 = 0; c0 < dim; c0 += 16) {
    const int c = std::min(c0, dim - 16);
    const int rc = bt_slab16(p, d_B + c, ldb, d_C + c, ldc, as_stream(stream_));
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
