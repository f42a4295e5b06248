// dense_tc.cu -- exact-split tcgen05 GEMM for the wide first layer:  out[n x p] = X[n x f] * W[f x p]  (hidden 600 of
// parameters/parameters_reddit.txt; SparseMatmul::forward on an all-columns feature matrix, src/module.cu:108-132).
//
// It re-uses exactly the tcgen05 pieces of the bit-tile GraphSum (spmm_bittile.cu): no-swizzle K-major shared-memory
// descriptors fed by cp.async.bulk of pre-packed operand images, fp32 accumulators in TMEM, tcgen05.commit / mbarrier stage
// recycling.  First B200 run (round 2): 1.0 ms against 8.0 ms on the SIMT kernel at 232965 x 602 x 600, but 1e-5 off: the
// tensor core's fp32 accumulate TRUNCATES, and a chain of 228 MMAs into one accumulator drifts by ~228 * 2^-25 of its
// magnitude.  (The bit-tile product does not drift: its addends are 8-bit-significand pieces whose partial sums stay
// exactly representable.)  Hence the accumulator CLASSES below.
//
// Why: at hidden 600 the product is compute bound (168 GFLOP per call) and the fp32 SIMT kernel runs at ~18 TFLOP/s =
// 15 ms.  Exactness on bf16 tensor cores: x and w are split into three bf16 pieces each (8 + 8 + 8
// significand bits, exact), a product of two pieces is exact in fp32, and of the nine piece products the six with
// (piece_x + piece_w) <= 2 are kept -- hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid -- the dropped ones are below
// 2^-24 of the product.  Six bf16 MMAs per k-step instead of one: ~1 PFLOP-equivalent, ~0.5 ms at the B200's bf16 rate.
//
// X is a constant of the training run: it is packed ONCE (gcnb_dense_tc_pack_x) into the operand image the MMA reads
// from shared memory -- per (block of 128 rows, k-step of 16 features, piece) a 4 KB K-major tile -- so the kernel has no
// operand transformation at all: one thread streams A and B tiles with bulk copies, one thread issues MMAs, four warps
// drain the accumulators.  W (f x p) is packed per call (gcnb_dense_tc_pack_w, 2 MB).  The p columns are cut into parts
// of <= 160 columns (600 -> 4 x 160); a part owns THREE accumulators of that width in TMEM, one per magnitude class of
// the piece products -- hi*hi | hi*mid + mid*hi | hi*lo + lo*hi + mid*mid -- so that the only chain whose truncation
// matters (hi*hi) is one MMA per k-step long (38 at f = 602: <= 38 * 2^-24 of the sum), and the epilogue adds the
// classes small to large with rounded fp32 adds.  The weight gradient cuts K (the nodes) into slices of <= 32 k-steps for
// the same reason; their partial tiles are added in ascending order.  The parts of a row block are consecutive items of
// one CTA, so X's tiles are re-read from L2.
// Dropout on X is not supported here (the wide configuration has input dropout 0; evaluation never has one).
#include <algorithm>
#include <cstdint>

#include "bulk.cuh"
#include "common.cuh"
#include "tcgen05.cuh"

using namespace gcnb;

namespace gcnb {

constexpr int kTcRows = 128;
constexpr int kTcATile = kTcRows * 16 * 2;  // 4096 bytes: 128 x 16 bf16
constexpr int kTcStages = 4;
constexpr int kTcThreads = 6 * 32;          // warps 0-3 epilogue, 4 producer, 5 MMA issuer

__device__ __forceinline__ void tcg_split3(float x, uint32_t &hi, uint32_t &mid, uint32_t &lo) {  // bf16 bit patterns
  const uint32_t hb = __float_as_uint(x) & 0xffff0000u;
  const float r1 = x - __uint_as_float(hb);
  const uint32_t mb = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(mb);
  hi = hb >> 16;
  mid = mb >> 16;
  lo = (__float_as_uint(r2) & 0xffff0000u) >> 16;
}

// A image: tile (row block b, k-step ks, piece pc) at ((b * KS + ks) * 3 + pc) * 4096; thread = (row, k-half): 8 features
__global__ void __launch_bounds__(256) tc_pack_x_kernel(const float *__restrict__ X, uint8_t *__restrict__ img, int64_t n, int f,
                                                        int KS, int64_t n_blk) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n_blk * 128 rows x KS k-steps x 2 halves
  const int64_t total = n_blk * kTcRows * (int64_t)KS * 2;
  if (tid >= total) return;
  const int half = (int)(tid & 1);
  const int64_t t2 = tid >> 1;
  const int r = (int)(t2 % kTcRows);
  const int64_t t3 = t2 / kTcRows;
  const int ks = (int)(t3 % KS);
  const int64_t b = t3 / KS;
  const int64_t row = b * kTcRows + r;
  uint32_t h[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = ks * 16 + half * 8 + i;
    const float x = (row < n && k < f) ? __ldg(X + row * f + k) : 0.f;
    uint32_t a, c, d;
    tcg_split3(x, a, c, d);
    const int sh = (i & 1) * 16;
    h[i >> 1] |= a << sh;
    m[i >> 1] |= c << sh;
    l[i >> 1] |= d << sh;
  }
  uint8_t *tile = img + ((b * KS + ks) * 3) * (int64_t)kTcATile + half * 2048 + (r >> 3) * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4 *>(tile) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(tile + kTcATile) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4 *>(tile + 2 * kTcATile) = make_uint4(l[0], l[1], l[2], l[3]);
}

// B image: tile (part q, k-step ks, piece pc) at ((q * KS + ks) * 3 + pc) * pcols * 32; operand row = output column
__global__ void __launch_bounds__(256) tc_pack_w_kernel(const float *__restrict__ W, uint8_t *__restrict__ img, int f, int p, int KS,
                                                        int n_parts, int pcols) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over parts x KS x pcols columns x 2 halves
  const int64_t total = (int64_t)n_parts * KS * pcols * 2;
  if (tid >= total) return;
  const int half = (int)(tid & 1);
  const int64_t t2 = tid >> 1;
  const int c = (int)(t2 % pcols);
  const int64_t t3 = t2 / pcols;
  const int ks = (int)(t3 % KS);
  const int q = (int)(t3 / KS);
  const int col = q * pcols + c;
  uint32_t h[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int k = ks * 16 + half * 8 + i;
    const float x = (col < p && k < f) ? __ldg(W + (int64_t)k * p + col) : 0.f;
    uint32_t a, b2, d;
    tcg_split3(x, a, b2, d);
    const int sh = (i & 1) * 16;
    h[i >> 1] |= a << sh;
    m[i >> 1] |= b2 << sh;
    l[i >> 1] |= d << sh;
  }
  const int64_t piece_bytes = (int64_t)pcols * 32;
  uint8_t *tile = img + (((int64_t)q * KS + ks) * 3) * piece_bytes + (int64_t)half * (pcols * 16) + (c >> 3) * 128 + (c & 7) * 16;
  *reinterpret_cast<uint4 *>(tile) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(tile + piece_bytes) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4 *>(tile + 2 * piece_bytes) = make_uint4(l[0], l[1], l[2], l[3]);
}

struct TcArgs {
  const uint8_t *a_img, *b_img;
  float *out;        // k_slices == 1: out[n x p];  k_slices > 1: partial tiles [slice][block][part][128][pcols]
  int64_t n, n_blk;  // rows of the A operand (output rows) and their blocks of 128
  int p, KS, n_parts, pcols, k_slices;
};

// item j of a CTA -> (row block, column part, k range).  One k slice: the parts of a row block are consecutive items of one
// CTA (the A tiles are re-read from L2).  Split K (the transposed product: few output tiles, long K): flat round robin.
struct TcItem {
  int64_t blk;
  int q, ks0, ks1, slice;
};
__device__ __forceinline__ int64_t tc_item_count(const TcArgs &a) {
  if (a.k_slices == 1) {
    const int64_t my_blocks = a.n_blk > (int64_t)blockIdx.x ? (a.n_blk - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    return my_blocks * a.n_parts;
  }
  const int64_t total = a.n_blk * a.n_parts * a.k_slices;
  return total > (int64_t)blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
}
__device__ __forceinline__ TcItem tc_item(const TcArgs &a, int64_t j) {
  TcItem it;
  if (a.k_slices == 1) {
    it.blk = blockIdx.x + (j / a.n_parts) * gridDim.x;
    it.q = (int)(j % a.n_parts);
    it.slice = 0;
    it.ks0 = 0;
    it.ks1 = a.KS;
  } else {
    const int64_t id = blockIdx.x + j * gridDim.x;
    it.q = (int)(id % a.n_parts);
    it.blk = (id / a.n_parts) % a.n_blk;
    it.slice = (int)(id / (a.n_parts * a.n_blk));
    it.ks0 = (int)((int64_t)a.KS * it.slice / a.k_slices);
    it.ks1 = (int)((int64_t)a.KS * (it.slice + 1) / a.k_slices);
  }
  return it;
}

// persistent: CTA c takes row blocks c, c + grid, ...; the n_parts column parts of a block are consecutive items
__global__ void __launch_bounds__(kTcThreads, 1) tc_gemm_kernel(TcArgs a) {
  extern __shared__ __align__(128) uint8_t tc_smem[];
  const uint32_t b_stage_bytes = 3u * (uint32_t)a.pcols * 32u;
  const uint32_t stage_bytes = 3u * kTcATile + b_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tc_smem + (size_t)kTcStages * stage_bytes);
  uint64_t *full = bars, *free_ = bars + kTcStages, *acc_full = free_ + kTcStages, *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&free_[i], 1);
    }
    for (int i = 0; i < 2; i++) {  // one accumulator set (three classes x pcols columns); slot 1 is unused
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);
  const int64_t n_items = tc_item_count(a);

  if (warp < 4) {
    // ---- epilogue: thread = row
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int64_t k = 0; k < n_items; k++) {
      const TcItem it = tc_item(a, k);
      const int64_t blk = it.blk;
      const int q = it.q;
      const uint32_t set = 0u, use = (uint32_t)k;
      const int r_in_blk = warp * 32 + lane;
      const int64_t row = blk * kTcRows + r_in_blk;
      mbar_wait(&acc_full[set], use & 1);
      tc_fence_after();
      const uint32_t acc0 = tmem + lane_base;
      float *tile = a.k_slices > 1
                        ? a.out + ((((int64_t)it.slice * a.n_blk + blk) * a.n_parts + q) * kTcRows + r_in_blk) * (int64_t)a.pcols
                        : nullptr;
      for (int c0 = 0; c0 < a.pcols; c0 += 16) {
        float v[16], v1[16], v2[16];
        tc_ld16(acc0 + 2 * a.pcols + c0, v2);  // hi*lo + lo*hi + mid*mid
        tc_ld16(acc0 + a.pcols + c0, v1);      // hi*mid + mid*hi
        tc_ld16(acc0 + c0, v);                 // hi*hi
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (v2[i] + v1[i]) + v[i];  // small to large, rounded adds
        if (tile) {  // partial tile of this k slice: every element is written (padding rows / columns hold exact zeros)
#pragma unroll
          for (int i = 0; i < 4; i++)
            reinterpret_cast<float4 *>(tile + c0)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else if (row < a.n) {
          const int col0 = q * a.pcols + c0;
          float *dst = a.out + row * a.p + col0;
#pragma unroll
          for (int i = 0; i < 16; i++)
            if (col0 + i < a.p) dst[i] = v[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
  } else if (warp == 4) {
    // ---- producer: one A stage (3 x 4 KB) + one B stage (3 x pcols x 32 bytes) per k-step
    if (lane == 0) {
      uint64_t t = 0;
      for (int64_t k = 0; k < n_items; k++) {
        const TcItem it = tc_item(a, k);
        const int64_t blk = it.blk;
        const int q = it.q;
        for (int ks = it.ks0; ks < it.ks1; ks++, t++) {
          const uint32_t s = (uint32_t)(t % kTcStages), use = (uint32_t)(t / kTcStages);
          if (use > 0) mbar_wait(&free_[s], (use - 1) & 1);
          uint8_t *stage = tc_smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
          bulk_g2s(stage, a.a_img + ((blk * a.KS + ks) * 3) * (int64_t)kTcATile, 3 * kTcATile, &full[s]);
          bulk_g2s(stage + 3 * kTcATile, a.b_img + (((int64_t)q * a.KS + ks) * 3) * (int64_t)(a.pcols * 32), b_stage_bytes,
                   &full[s]);
        }
      }
    }
  } else {
    // ---- MMA issuer (warp converged, one elected lane issues): six piece products per k-step
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a.pcols >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
    uint64_t t = 0;
    for (int64_t k = 0; k < n_items; k++) {
      const uint32_t set = 0u, use = (uint32_t)k;
      if (use > 0) mbar_wait(&acc_empty[set], (use - 1) & 1);
      tc_fence_after();
      const uint32_t d0 = tmem, d1 = tmem + (uint32_t)a.pcols, d2 = tmem + 2u * (uint32_t)a.pcols;  // accumulator classes
      const TcItem it = tc_item(a, k);
      for (int ks = it.ks0; ks < it.ks1; ks++, t++) {
        const uint32_t s = (uint32_t)(t % kTcStages), u = (uint32_t)(t / kTcStages);
        mbar_wait(&full[s], u & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(tc_smem + (size_t)s * stage_bytes);
          const uint32_t sb = sa + 3 * kTcATile;
          const uint32_t pb = (uint32_t)a.pcols * 32u;
          const uint64_t a0 = tc_smem_desc(sa, 2048), a1 = tc_smem_desc(sa + kTcATile, 2048), a2 = tc_smem_desc(sa + 2 * kTcATile, 2048);
          const uint64_t b0 = tc_smem_desc(sb, (uint32_t)a.pcols * 16u), b1 = tc_smem_desc(sb + pb, (uint32_t)a.pcols * 16u),
                         b2 = tc_smem_desc(sb + 2 * pb, (uint32_t)a.pcols * 16u);
          const uint32_t first = ks > it.ks0 ? 1u : 0u;
          tc_mma_ss(d2, a2, b0, idesc, first);  // lo*hi    class 2 (with the next two): below 2^-16 of the product
          tc_mma_ss(d2, a0, b2, idesc, 1u);     // hi*lo
          tc_mma_ss(d2, a1, b1, idesc, 1u);     // mid*mid
          tc_mma_ss(d1, a1, b0, idesc, first);  // mid*hi   class 1 (with the next one): below 2^-8
          tc_mma_ss(d1, a0, b1, idesc, 1u);     // hi*mid
          tc_mma_ss(d0, a0, b0, idesc, first);  // hi*hi    class 0: one MMA per k-step => the shortest chain
          tc_commit(&free_[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&acc_full[set]);
      __syncwarp();
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// A image of X^T for the weight gradient dW[f x p] = X^T * dH: operand row = feature, k = node.  Same tile format as
// tc_pack_x_kernel; thread = (feature, k-half) of one (feature block, k-step)
__global__ void __launch_bounds__(256) tc_pack_xt_kernel(const float *__restrict__ X, uint8_t *__restrict__ img, int64_t n, int f,
                                                         int KS, int64_t f_blk) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = f_blk * kTcRows * (int64_t)KS * 2;
  if (tid >= total) return;
  const int r = (int)(tid % kTcRows);          // adjacent threads = adjacent features: coalesced reads of X's rows
  const int64_t t2 = tid / kTcRows;
  const int half = (int)(t2 & 1);
  const int64_t t3 = t2 >> 1;
  const int ks = (int)(t3 % KS);
  const int64_t b = t3 / KS;
  const int64_t feat = b * kTcRows + r;
  uint32_t h[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int64_t node = (int64_t)ks * 16 + half * 8 + i;
    const float x = (feat < f && node < n) ? __ldg(X + node * f + feat) : 0.f;
    uint32_t a, c, d;
    tcg_split3(x, a, c, d);
    const int sh = (i & 1) * 16;
    h[i >> 1] |= a << sh;
    m[i >> 1] |= c << sh;
    l[i >> 1] |= d << sh;
  }
  uint8_t *tile = img + ((b * KS + ks) * 3) * (int64_t)kTcATile + half * 2048 + (r >> 3) * 128 + (r & 7) * 16;
  *reinterpret_cast<uint4 *>(tile) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4 *>(tile + kTcATile) = make_uint4(m[0], m[1], m[2], m[3]);
  *reinterpret_cast<uint4 *>(tile + 2 * kTcATile) = make_uint4(l[0], l[1], l[2], l[3]);
}

// out[rows x p] = sum over the k slices, ascending, of the partial tiles
__global__ void __launch_bounds__(256) tc_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out, int64_t rows,
                                                        int p, int64_t n_blk, int n_parts, int pcols, int k_slices) {
  const int64_t total = rows * p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / p;
    const int col = (int)(i % p);
    const int64_t blk = row / kTcRows;
    const int r = (int)(row % kTcRows), q = col / pcols, c = col % pcols;
    float s = 0.f;
    for (int sl = 0; sl < k_slices; sl++)
      s += partial[((((int64_t)sl * n_blk + blk) * n_parts + q) * kTcRows + r) * (int64_t)pcols + c];
    out[i] = s;
  }
}

// 4 stages x (12 KB of A + at most 24 KB of B) + barriers: allow it once
static cudaError_t tc_allow_smem() {
  static cudaError_t rc = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  return rc;
}

static void tc_shape(int f, int p, int *KS, int *n_parts, int *pcols) {
  *KS = (f + 15) / 16;
  const int p_pad = (p + 15) / 16 * 16;
  *n_parts = (p_pad + 159) / 160;
  *pcols = ((p_pad + *n_parts - 1) / *n_parts + 15) / 16 * 16;  // <= 160: three accumulator classes in 512 TMEM columns
}

}  // namespace gcnb

extern "C" {

int gcnb_dense_tc_supported(int f, int p) { return f >= 1 && p >= 16 && p <= 4096; }

// bytes of the packed X image (n rows, f features) and of the per-call workspace for W (f x p)
int64_t gcnb_dense_tc_x_bytes(int64_t n, int f) {
  const int64_t n_blk = (n + kTcRows - 1) / kTcRows;
  return n_blk * ((f + 15) / 16) * 3 * (int64_t)kTcATile;
}
int64_t gcnb_dense_tc_w_bytes(int f, int p) {
  int KS, n_parts, pcols;
  tc_shape(f, p, &KS, &n_parts, &pcols);
  return (int64_t)n_parts * KS * 3 * pcols * 32;
}

int gcnb_dense_tc_pack_x(const float *d_X, void *d_img, int64_t n, int f, gcnb_stream_t stream_) {
  if (!d_X || !d_img || n < 0 || f < 1) return GCNB_E_BADARG;
  if (n == 0) return 0;
  const int64_t n_blk = (n + kTcRows - 1) / kTcRows;
  const int KS = (f + 15) / 16;
  const int64_t total = n_blk * kTcRows * (int64_t)KS * 2;
  tc_pack_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream_)>>>(d_X, (uint8_t *)d_img, n, f, KS, n_blk);
  GCNB_LAUNCH_CHECK();
  return 0;
}

// out[n x p] = X * W with X given as its packed image (gcnb_dense_tc_pack_x); d_ws: gcnb_dense_tc_w_bytes(f, p) bytes
int gcnb_dense_tc_fwd_f32(const void *d_x_img, const float *d_W, float *d_out, int64_t n, int f, int p, void *d_ws,
                          int64_t ws_bytes, gcnb_stream_t stream_) {
  if (!d_x_img || !d_W || !d_out || !d_ws || n < 0 || !gcnb_dense_tc_supported(f, p)) return GCNB_E_BADARG;
  if (ws_bytes < gcnb_dense_tc_w_bytes(f, p)) return GCNB_E_BADARG;
  if (n == 0) return 0;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  if (di.cc_major != 10) return GCNB_E_UNSUPPORTED;
  cudaStream_t stream = as_stream(stream_);
  TcArgs a;
  tc_shape(f, p, &a.KS, &a.n_parts, &a.pcols);
  a.a_img = (const uint8_t *)d_x_img;
  a.b_img = (const uint8_t *)d_ws;
  a.out = d_out;
  a.n = n;
  a.n_blk = (n + kTcRows - 1) / kTcRows;
  a.p = p;
  a.k_slices = 1;
  const int64_t wt = (int64_t)a.n_parts * a.KS * a.pcols * 2;
  tc_pack_w_kernel<<<(unsigned)((wt + 255) / 256), 256, 0, stream>>>(d_W, (uint8_t *)d_ws, f, p, a.KS, a.n_parts, a.pcols);
  GCNB_LAUNCH_CHECK();
  const size_t smem = (size_t)kTcStages * (3 * kTcATile + 3 * (size_t)a.pcols * 32) + (2 * kTcStages + 4) * 8 + 16;
  GCNB_CHECK(tc_allow_smem());
  const int grid = (int)std::min<int64_t>(a.n_blk, std::max(1, di.sm_count));
  tc_gemm_kernel<<<grid, kTcThreads, smem, stream>>>(a);
  GCNB_LAUNCH_CHECK();
  return 0;
}

// ---- weight gradient  dW[f x p] = X^T[f x n] * dH[n x p]  (SparseMatmul::backward on a dense X, src/module.cu:136-163) --------
// X^T is packed once (operand rows = features, K = nodes); dH is packed per call; K is cut into slices so that all SMs have
// work (5 feature blocks x 3 column parts only), the slices' partial tiles are added in ascending order.
static int tc_tn_slices(int64_t n, int f, int p, int sms) {
  (void)f; (void)p; (void)sms;
  const int64_t ks = (n + 15) / 16;
  return (int)std::max<int64_t>(1, (ks + 31) / 32);  // <= 32 k-steps per accumulator chain (see the header)
}
int64_t gcnb_dense_tc_xt_bytes(int64_t n, int f) {
  return ((f + kTcRows - 1) / kTcRows) * ((n + 15) / 16) * 3 * (int64_t)kTcATile;
}
int gcnb_dense_tc_pack_xt(const float *d_X, void *d_img, int64_t n, int f, gcnb_stream_t stream_) {
  if (!d_X || !d_img || n < 0 || f < 1 || n > (1ll << 30)) return GCNB_E_BADARG;
  if (n == 0) return 0;
  const int64_t f_blk = (f + kTcRows - 1) / kTcRows;
  const int KS = (int)((n + 15) / 16);
  const int64_t total = f_blk * kTcRows * (int64_t)KS * 2;
  tc_pack_xt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream_)>>>(d_X, (uint8_t *)d_img, n, f, KS, f_blk);
  GCNB_LAUNCH_CHECK();
  return 0;
}
// workspace: packed dH image + partial tiles (148 SMs assumed for sizing when no device is present)
int64_t gcnb_dense_tc_tn_workspace(int64_t n, int f, int p) {
  if (n < 1 || n > (1ll << 30) || !gcnb_dense_tc_supported(f, p)) return 0;
  int KS, n_parts, pcols;
  tc_shape((int)n, p, &KS, &n_parts, &pcols);
  const DeviceInfo &di = device_info();
  const int slices = tc_tn_slices(n, f, p, di.ok ? std::max(1, di.sm_count) : 148);
  const int64_t img = (int64_t)n_parts * KS * 3 * pcols * 32;
  const int64_t partial = (int64_t)slices * ((f + kTcRows - 1) / kTcRows) * n_parts * kTcRows * pcols * 4;
  return img + partial + 256;
}
int gcnb_dense_tc_tn_f32(const void *d_xt_img, const float *d_dH, float *d_dW, int64_t n, int f, int p, void *d_ws,
                         int64_t ws_bytes, gcnb_stream_t stream_) {
  if (!d_xt_img || !d_dH || !d_dW || !d_ws || n < 1 || n > (1ll << 30) || !gcnb_dense_tc_supported(f, p)) return GCNB_E_BADARG;
  if (ws_bytes < gcnb_dense_tc_tn_workspace(n, f, p)) return GCNB_E_BADARG;
  const DeviceInfo &di = device_info();
  if (!di.ok) return (int)cudaErrorNoDevice;
  if (di.cc_major != 10) return GCNB_E_UNSUPPORTED;
  cudaStream_t stream = as_stream(stream_);
  TcArgs a;
  tc_shape((int)n, p, &a.KS, &a.n_parts, &a.pcols);  // K = nodes
  a.k_slices = tc_tn_slices(n, f, p, std::max(1, di.sm_count));
  a.n = f;
  a.n_blk = (f + kTcRows - 1) / kTcRows;
  a.p = p;
  const int64_t img_bytes = ((int64_t)a.n_parts * a.KS * 3 * a.pcols * 32 + 255) / 256 * 256;
  a.a_img = (const uint8_t *)d_xt_img;
  a.b_img = (const uint8_t *)d_ws;
  float *partial = reinterpret_cast<float *>((uint8_t *)d_ws + img_bytes);
  a.out = a.k_slices > 1 ? partial : d_dW;
  const int64_t wt = (int64_t)a.n_parts * a.KS * a.pcols * 2;
  tc_pack_w_kernel<<<(unsigned)((wt + 255) / 256), 256, 0, stream>>>(d_dH, (uint8_t *)d_ws, (int)n, p, a.KS, a.n_parts, a.pcols);
  GCNB_LAUNCH_CHECK();
  const size_t smem = (size_t)kTcStages * (3 * kTcATile + 3 * (size_t)a.pcols * 32) + (2 * kTcStages + 4) * 8 + 16;
  GCNB_CHECK(tc_allow_smem());
  const int64_t items = a.n_blk * a.n_parts * a.k_slices;
  const int grid = (int)std::min<int64_t>(items, std::max(1, di.sm_count));
  tc_gemm_kernel<<<grid, kTcThreads, smem, stream>>>(a);
  GCNB_LAUNCH_CHECK();
  if (a.k_slices > 1) {
    const int64_t total = (int64_t)f * p;
    tc_reduce_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)di.sm_count * 8), 256, 0, stream>>>(
        partial, d_dW, f, p, a.n_blk, a.n_parts, a.pcols, a.k_slices);
    GCNB_LAUNCH_CHECK();
  }
  return 0;
}

// debugging aid: copies of the packed images for the CPU-side layout check
int gcnb_dense_tc_debug_pack_w(const float *d_W, void *d_ws, int f, int p, gcnb_stream_t stream_) {
  if (!d_W || !d_ws || !gcnb_dense_tc_supported(f, p)) return GCNB_E_BADARG;
  int KS, n_parts, pcols;
  tc_shape(f, p, &KS, &n_parts, &pcols);
  const int64_t wt = (int64_t)n_parts * KS * pcols * 2;
  tc_pack_w_kernel<<<(unsigned)((wt + 255) / 256), 256, 0, as_stream(stream_)>>>(d_W, (uint8_t *)d_ws, f, p, KS, n_parts, pcols);
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
