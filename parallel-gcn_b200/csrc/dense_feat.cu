// dense_feat.cu -- the first layer when the feature matrix is dense (every svmlight row lists all F columns, as
// Reddit's 602 do): X[N x F] * W0[F x P] and its weight gradient X^T * dH, with the input dropout applied on the fly
// from a 1-bit/element keep mask instead of materialising a dropped copy of X.
//
// Reference path being replaced (per training epoch, Reddit): Dropout::forward rewrites the 561 MB input in place and
// pulls 2.2 GB of Philox state through HBM (src/module.cu:16-76), SparseMatmul::forward re-reads values + 561 MB of
// column indices (src/module.cu:108-132), SparseMatmul::backward issues Fnnz*H atomicAdds (src/module.cu:136-152), and
// every eval first restores the input with a 1.1 GB copy (src/gcn.cu:181-200).  Here: one Philox pass writes 17.5 MB of
// mask bits; forward and backward each stream X exactly once (561 MB, the HBM roofline of these two kernels).
//
//   dropout_maskbits_kernel : bit j of the mask = keep decision of element j (same Philox stream/lanes as Dropout)
//   dense_feat_fwd_kernel   : warp = R rows x all P columns, lanes stride over k; W0 staged in shared memory (padded
//                             rows: conflict-free LDS.128); 64 accumulators per lane, transposing butterfly reduction
//   dense_feat_tn_kernel    : thread = FPT features x P columns, CTAs own row slabs; partials reduced in slab order
//                             (deterministic, no atomics)
#include <algorithm>

#include "common.cuh"
#include "philox.cuh"
#include "bulk.cuh"

using namespace gcnb;

// opt a kernel in to `bytes` of dynamic shared memory, once per (kernel, larger size): steady-state launches make no
// attribute call (and none inside a CUDA-graph capture)
#define GCNB_SMEM_OPT_IN(bytes, ...)                                                                                 \
  do {                                                                                                               \
    static size_t opted_in = 0;                                                                                      \
    if ((size_t)(bytes) > opted_in) {                                                                                \
      GCNB_CHECK(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      opted_in = (size_t)(bytes);                                                                                    \
    }                                                                                                                \
  } while (0)

namespace {

constexpr int kT = 256;

constexpr int TR = 32;  // rows per tile (8 warps x 4 rows in the forward kernel)

// ---- 1 bit per element keep mask, in the tile layout the two products consume ----------------------------------------
// Rows are grouped in tiles of TR = 32; inside a tile element (r, k) is bit idx = (r % 32) * F + k, i.e. the tile's
// elements in memory order, and every tile starts on a 16-byte boundary (WPT words per tile, padded) so that a tile's
// mask is one bulk copy.  Global element of (tile, idx) = tile * 32 * F + idx.
__host__ __device__ inline int64_t mask_words_per_tile(int F) { return ((((int64_t)TR * F + 31) / 32) + 3) & ~(int64_t)3; }

__global__ void dropout_maskbits_kernel(uint32_t *__restrict__ bits, int64_t size, int F, int64_t wpt, int64_t total_words,
                                        float p, gcnb_rng_t rng) {
  const int64_t tile_elems = (int64_t)TR * F;
  const uint32_t lead = rng.elem_lead;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total_words; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tile = w / wpt;
    const int64_t idx0 = (w - tile * wpt) * 32;            // first bit of this word inside the tile
    int64_t nvalid = min((int64_t)32, tile_elems - idx0);  // padding words / bits stay zero
    const int64_t e0 = tile * tile_elems + idx0;           // local element index of bit 0
    nvalid = min(nvalid, size - e0);
    uint32_t out = 0;
    if (nvalid > 0) {
      const uint64_t pos0 = (uint64_t)e0 + lead;  // position counted from the first global group of this call
      const uint32_t g0 = (uint32_t)(pos0 >> 2);
      const int shift = (int)(pos0 & 3);          // bit b of the word is lane (b + shift) & 3 of group g0 + (b + shift) / 4
      if (shift == 0 && nvalid == 32) {
        // the common case (full word on a group boundary): eight independent Philox chains in flight -- a Philox call is
        // ten dependent multiply rounds, one chain at a time leaves the integer pipes idle
        float u[8][4];
#pragma unroll
        for (int gi = 0; gi < 8; gi++) rng_uniform4(rng, g0 + gi, u[gi]);
#pragma unroll
        for (int gi = 0; gi < 8; gi++)
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (u[gi][k] >= p) out |= 1u << (gi * 4 + k);
      } else {
#pragma unroll 1
        for (int gi = 0; gi < 9; gi++) {
          const int b_first = gi * 4 - shift;       // word bit of lane 0 of this group
          if (b_first >= nvalid) break;
          float u[4];
          rng_uniform4(rng, g0 + gi, u);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int b = b_first + k;
            if (b >= 0 && b < nvalid && u[k] >= p) out |= 1u << b;
          }
        }
      }
    }
    bits[w] = out;
  }
}

// keep-test on a tile's mask staged in shared memory (idx = position of the element inside the tile)
__device__ __forceinline__ float masked(float x, const uint32_t *sbits, int idx, float scale) {
  if (sbits == nullptr) return x;
  return ((sbits[idx >> 5] >> (idx & 31)) & 1u) ? x * scale : 0.f;  // Dropout: x *= keep ? scale : 0
}

// ---- TMA bulk-copy plumbing (cp.async.bulk + mbarrier): a tile of TR consecutive rows of X is ONE contiguous block
// of HBM, so each stage is a single bulk copy of up to 77 KB issued by one thread -- 150 KB in flight per SM, which
// is what it takes to stream at HBM speed (per-lane LDG streams from 16 warps keep ~8 KB in flight and crawl).
// thread 0: start the copy of rows [r0, r0+nrows) of a row-major [.. x F] matrix into `dst`; the 16-byte-multiple
// prefix goes through the bulk engine, a possible 4/8/12-byte tail by plain stores (visible after the next barrier)
__device__ __forceinline__ void issue_tile(float *dst, const float *__restrict__ src, int64_t r0, int nrows, int F,
                                           uint64_t *bar) {
  const size_t bytes = (size_t)nrows * F * sizeof(float);
  const uint32_t bulk = (uint32_t)(bytes & ~(size_t)15);
  const float *g = src + (size_t)r0 * F;
  mbar_expect_tx(bar, bulk);
  if (bulk) bulk_g2s(dst, g, bulk, bar);
  for (size_t i = bulk / 4; i < bytes / 4; i++) dst[i] = __ldg(g + i);
}

// ---- forward: out[N x P] = (X .* mask*scale)[N x F] * W[F x P] ---------------------------------------------------
// persistent CTA (1 per SM): W0 resident in shared memory (padded rows: conflict-free LDS.128), X tiles double-buffered
// by bulk copies; each warp owns R = 64/P... here TR/8 = 4 rows of the tile and all P columns, lanes stride over k.
template <int P>
__global__ void __launch_bounds__(kT, 1)
dense_feat_fwd_kernel(const float *__restrict__ X, const uint32_t *__restrict__ bits, float scale,
                      const float *__restrict__ W, float *__restrict__ out, int64_t N, int F) {
  constexpr int R = TR / (kT / 32);  // rows per warp per tile
  constexpr int WP = P + 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile0 = reinterpret_cast<float *>(smem_raw);
  float *tile1 = tile0 + (size_t)TR * F;
  float *Ws = tile1 + (size_t)TR * F;
  const int wpt = (int)mask_words_per_tile(F);
  uint32_t *mb0 = reinterpret_cast<uint32_t *>(Ws + (((size_t)F * WP + 3) & ~(size_t)3));
  uint32_t *mb1 = mb0 + wpt;
  uint64_t *bars = reinterpret_cast<uint64_t *>(mb1 + wpt);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ntiles = (N + TR - 1) / TR;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int nbar = bits ? 2 : 1;  // arrivals per stage: X tile (+ mask tile)
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], nbar);
    mbar_init(&bars[1], nbar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int64_t t) {
    issue_tile(s ? tile1 : tile0, X, t * TR, (int)min((int64_t)TR, N - t * TR), F, &bars[s]);
    if (bits) {
      mbar_expect_tx(&bars[s], (uint32_t)wpt * 4);
      bulk_g2s(s ? mb1 : mb0, bits + t * wpt, (uint32_t)wpt * 4, &bars[s]);
    }
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {
      const int64_t t = first + s * stride;
      if (t < ntiles) issue(s, t);
    }
  }
  for (int i = threadIdx.x; i < F * P; i += kT) Ws[(i / P) * WP + (i % P)] = __ldg(W + i);
  __syncthreads();
  int it = 0;
  for (int64_t t = first; t < ntiles; t += stride, it++) {
    const int s = it & 1;
    const float *tile = s ? tile1 : tile0;
    const uint32_t *sbits = bits ? (s ? mb1 : mb0) : nullptr;
    mbar_wait(&bars[s], (uint32_t)((it >> 1) & 1));
    const int64_t r0 = t * TR + wib * R;
    float acc[R * P];
#pragma unroll
    for (int i = 0; i < R * P; i++) acc[i] = 0.f;
    const float *xrow = tile + (size_t)(wib * R) * F;
#pragma unroll 2
    for (int k = lane; k < F; k += 32) {
      float x[R];
#pragma unroll
      for (int q = 0; q < R; q++) {
        const int64_t r = r0 + q;
        x[q] = (r < N) ? masked(xrow[(size_t)q * F + k], sbits, (wib * R + q) * F + k, scale) : 0.f;
      }
      const float4 *wr = reinterpret_cast<const float4 *>(Ws + k * WP);
#pragma unroll
      for (int c4 = 0; c4 < P / 4; c4++) {
        const float4 w = wr[c4];
#pragma unroll
        for (int q = 0; q < R; q++) {
          acc[q * P + c4 * 4 + 0] = fmaf(x[q], w.x, acc[q * P + c4 * 4 + 0]);
          acc[q * P + c4 * 4 + 1] = fmaf(x[q], w.y, acc[q * P + c4 * 4 + 1]);
          acc[q * P + c4 * 4 + 2] = fmaf(x[q], w.z, acc[q * P + c4 * 4 + 2]);
          acc[q * P + c4 * 4 + 3] = fmaf(x[q], w.w, acc[q * P + c4 * 4 + 3]);
        }
      }
    }
    // transposing butterfly: R*P partial sums per lane -> every lane ends with R*P/32 complete outputs
    constexpr int V = R * P;  // 64 for P=16, 128 for P=32, 32 for P=8
#pragma unroll
    for (int m = 16, n = V; m >= 1; m >>= 1, n >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < n / 2; i++) {
        const float lo = acc[i], hi = acc[i + n / 2];
        const float send = up ? lo : hi;
        const float keep = up ? hi : lo;
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
    constexpr int PER = V / 32;  // contiguous outputs per lane: indices lane*PER .. lane*PER+PER-1 (= q*P + c)
    const int i0 = lane * PER;
    const int q = i0 / P, c = i0 % P;
    const int64_t r = r0 + q;
    if (r < N) {
#pragma unroll
      for (int i = 0; i < PER; i++) out[r * P + c + i] = acc[i];
    }
    __syncthreads();  // everyone is done with this stage's tile
    if (threadIdx.x == 0) {
      const int64_t tn = t + 2 * stride;
      if (tn < ntiles) issue(s, tn);
    }
  }
}

// ---- weight gradient: dW[F x P] = (X .* mask*scale)^T * dH[N x P] --------------------------------------------------
// persistent CTA: thread t owns features t, t+256, ... (FPT of them) x all P columns in registers; X and dH tiles of
// TR rows arrive by bulk copy; per-CTA partial written once at the end, reduced in CTA order by a second kernel.
template <int P, int FPT>
__global__ void __launch_bounds__(kT, 1)
dense_feat_tn_kernel(const float *__restrict__ X, const uint32_t *__restrict__ bits, float scale,
                     const float *__restrict__ dH, float *__restrict__ ws, int64_t N, int F) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float *tile0 = reinterpret_cast<float *>(smem_raw);
  float *tile1 = tile0 + (size_t)TR * F;
  float *dh0 = tile1 + (size_t)TR * F;
  float *dh1 = dh0 + TR * P;
  const int wpt = (int)mask_words_per_tile(F);
  uint32_t *mb0 = reinterpret_cast<uint32_t *>(dh1 + TR * P);
  uint32_t *mb1 = mb0 + wpt;
  uint64_t *bars = reinterpret_cast<uint64_t *>(mb1 + wpt);
  const int tid = threadIdx.x;
  const int64_t ntiles = (N + TR - 1) / TR;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int nbar = bits ? 3 : 2;  // expect_tx arrivals per stage: X tile, dH tile (+ mask tile)
  if (tid == 0) {
    mbar_init(&bars[0], nbar);
    mbar_init(&bars[1], nbar);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int64_t t) {
    const int nr = (int)min((int64_t)TR, N - t * TR);
    issue_tile(s ? tile1 : tile0, X, t * TR, nr, F, &bars[s]);
    issue_tile(s ? dh1 : dh0, dH, t * TR, nr, P, &bars[s]);
    if (bits) {
      mbar_expect_tx(&bars[s], (uint32_t)wpt * 4);
      bulk_g2s(s ? mb1 : mb0, bits + t * wpt, (uint32_t)wpt * 4, &bars[s]);
    }
  };
  if (tid == 0) {
    for (int s = 0; s < 2; s++) {
      const int64_t t = first + s * stride;
      if (t < ntiles) issue(s, t);
    }
  }
  __syncthreads();
  float acc[FPT][P];
#pragma unroll
  for (int f = 0; f < FPT; f++)
#pragma unroll
    for (int c = 0; c < P; c++) acc[f][c] = 0.f;
  int it = 0;
  for (int64_t t = first; t < ntiles; t += stride, it++) {
    const int s = it & 1;
    const float *tile = s ? tile1 : tile0;
    const float4 *dh = reinterpret_cast<const float4 *>(s ? dh1 : dh0);
    const uint32_t *sbits = bits ? (s ? mb1 : mb0) : nullptr;
    mbar_wait(&bars[s], (uint32_t)((it >> 1) & 1));
    const int nrows = (int)min((int64_t)TR, N - t * TR);
#pragma unroll 4
    for (int rr = 0; rr < nrows; rr++) {
      float x[FPT];
#pragma unroll
      for (int f = 0; f < FPT; f++) {
        const int j = tid + f * kT;
        x[f] = (j < F) ? masked(tile[(size_t)rr * F + j], sbits, rr * F + j, scale) : 0.f;
      }
#pragma unroll
      for (int c4 = 0; c4 < P / 4; c4++) {
        const float4 g = dh[rr * (P / 4) + c4];
#pragma unroll
        for (int f = 0; f < FPT; f++) {
          acc[f][c4 * 4 + 0] = fmaf(x[f], g.x, acc[f][c4 * 4 + 0]);
          acc[f][c4 * 4 + 1] = fmaf(x[f], g.y, acc[f][c4 * 4 + 1]);
          acc[f][c4 * 4 + 2] = fmaf(x[f], g.z, acc[f][c4 * 4 + 2]);
          acc[f][c4 * 4 + 3] = fmaf(x[f], g.w, acc[f][c4 * 4 + 3]);
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      const int64_t tn = t + 2 * stride;
      if (tn < ntiles) issue(s, tn);
    }
  }
  float *dst = ws + (size_t)blockIdx.x * F * P;
#pragma unroll
  for (int f = 0; f < FPT; f++) {
    const int j = tid + f * kT;
    if (j < F) {
#pragma unroll
      for (int c4 = 0; c4 < P / 4; c4++)
        *reinterpret_cast<float4 *>(dst + (size_t)j * P + c4 * 4) =
            make_float4(acc[f][c4 * 4], acc[f][c4 * 4 + 1], acc[f][c4 * 4 + 2], acc[f][c4 * 4 + 3]);
    }
  }
}

// =====================================================================================================================
// Tensor-core variants for P = 16 (hidden 16, the Part-1 model): mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, hi = tf32(x), lo = tf32(x - hi);  x*w ~= lo*w_hi + hi*w_lo + hi*w_hi, fp32 accumulation), which keeps
// fp32-level accuracy (relative error of a product ~2^-21 before accumulation).  Why: the FMA kernels above are
// issue-bound -- 16 FFMA per element of X plus mask and address work, ~0.9-1.3 warp instructions per element at 2 warps
// per scheduler (ncu: 200-375 us for 561 MB, 0.23-0.44 of the HBM roofline); one mma replaces 32 FFMA warp
// instructions, so the instruction stream shrinks ~3x and the kernels approach the copy rate of X.
// X streams through shared memory in half-tiles of 16 rows (one bulk copy each, kStages deep); a mask tile (32 rows)
// is fetched with each of its two half-tiles.
constexpr int HR = 16;       // rows per half-tile = m (forward) / 2 k-steps (weight gradient)
// pipeline shape: two CTAs per SM with two half-tiles in flight each when that fits (while one CTA sits at its
// per-tile barrier the other one computes), else one CTA with three (measured: deeper pipelines do not help)

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// W / dH side (once per kernel / once per k-step): round-to-nearest split
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
// X side (once per element, the hot path): hi = x with the 13 low mantissa bits cleared, lo = x - hi exactly (the
// tensor core reads the upper 19 bits of lo).  Two instructions instead of the eight a rounded split compiles to;
// a product is then off by at most ~2^-20 relative, within the fp32 parity bar of 1e-5.
__device__ __forceinline__ void split_trunc(float x, uint32_t &hi, uint32_t &lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 32 keep bits starting at element idx of a mask tile (one spare zero word follows every tile)
__device__ __forceinline__ uint32_t mask_window(const uint32_t *sbits, int idx) {
  const int w = idx >> 5;
  return __funnelshift_r(sbits[w], sbits[w + 1], idx & 31);
}
// dropped elements become +0; the 1/(1-p) scale is applied once to the accumulated result
__device__ __forceinline__ float keep_if(float x, uint32_t window, uint32_t bit) { return (window & bit) ? x : 0.f; }

// forward, P = 16: out[N x 16] = (X .* mask*scale) * W.  8 warps split K (F padded to a multiple of 8) of one 16-row
// half-tile: warp w always owns k-steps [w*KSP, (w+1)*KSP), so its B fragments (hi and lo halves of W, split once) stay
// in registers for the whole kernel and the inner loop is 4 LDS + split + 6 mma per k-step.  The 8 partial 16x16 results
// are added in warp order through shared memory (double-buffered: one __syncthreads per half-tile).
// Padding instead of predicates: columns >= F meet zero rows of W, rows >= N are computed but not stored, and the
// stage buffers are zero-filled once so that no stale NaN pattern can meet a zero.
template <bool MASK, int KSP>
__global__ void __launch_bounds__(kT, 2)
dense_feat_fwd_mma_kernel(const float *__restrict__ X, const uint32_t *__restrict__ bits, float scale,
                          const float *__restrict__ W, float *__restrict__ out, int64_t N, int F, const int kStages) {
  constexpr int P = 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int wpt = (int)mask_words_per_tile(F);
  const size_t tile_floats = ((size_t)HR * F + 3) & ~(size_t)3;
  float *tiles = reinterpret_cast<float *>(smem_raw);
  float *pad = tiles + kStages * tile_floats;             // 8 zero floats: the last row's reads past column F
  uint32_t *mbits = reinterpret_cast<uint32_t *>(pad + 8);  // kStages x (wpt + 4): spare zero words for the window
  float *red = reinterpret_cast<float *>(mbits + (size_t)kStages * (wpt + 4));  // 2 x 8 warps x 256
  uint64_t *bars = reinterpret_cast<uint64_t *>(red + 2 * 8 * 256);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t nht = (N + HR - 1) / HR;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  for (size_t i = threadIdx.x; i < kStages * tile_floats + 8; i += kT) tiles[i] = 0.f;
  if (MASK)
    for (int i = threadIdx.x; i < kStages * (wpt + 4); i += kT) mbits[i] = 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], MASK ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int64_t ht) {
    issue_tile(tiles + s * tile_floats, X, ht * HR, (int)min((int64_t)HR, N - ht * HR), F, &bars[s]);
    if (MASK) {
      mbar_expect_tx(&bars[s], (uint32_t)wpt * 4);
      bulk_g2s(mbits + (size_t)s * (wpt + 4), bits + (ht >> 1) * wpt, (uint32_t)wpt * 4, &bars[s]);
    }
  };
  if (threadIdx.x == 0) {
    fence_proxy_async();  // the zero fill above (generic proxy) is ordered before the bulk copies (async proxy)
    for (int s = 0; s < kStages; s++) {
      const int64_t ht = first + s * stride;
      if (ht < nht) issue(s, ht);
    }
  }
  // this warp's B fragments: W rows k0 + t and k0 + t + 4, columns j*8 + g, for its KSP k-steps
  uint32_t bh[KSP][2][2], bl[KSP][2][2];
#pragma unroll
  for (int q = 0; q < KSP; q++) {
    const int ka = (wib * KSP + q) * 8 + t, kb = ka + 4;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      split_tf32(ka < F ? __ldg(W + (size_t)ka * P + j * 8 + g) : 0.f, bh[q][j][0], bl[q][j][0]);
      split_tf32(kb < F ? __ldg(W + (size_t)kb * P + j * 8 + g) : 0.f, bh[q][j][1], bl[q][j][1]);
    }
  }
  int it = 0;
  for (int64_t ht = first; ht < nht; ht += stride, it++) {
    const int s = it % kStages;
    const float *tile = tiles + s * tile_floats;
    const uint32_t *sb = mbits + (size_t)s * (wpt + 4);
    mbar_wait(&bars[s], (uint32_t)((it / kStages) & 1));
    const int nrows = (int)min((int64_t)HR, N - ht * HR);
    const float *xa = tile + (size_t)g * F + wib * KSP * 8 + t;  // (row g, first k of this warp)
    const float *xb = xa + (size_t)8 * F;                        // row g + 8
    const int ia = ((int)(ht & 1) * HR + g) * F + wib * KSP * 8 + t, ib = ia + 8 * F;  // the same elements' mask bits
    // The tensor core adds into its fp32 accumulator with truncation, so long in-place chains drift.  Here a chain is
    // at most 3 * KSP = 30 mma (the accumulators restart from zero for every half-tile and the eight warps' partial sums
    // are added with rounded FADDs below): <= ~2e-6 relative in the worst case, measured well inside the 1e-5 bar.
    float c[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    const int nq = min(KSP, (F + 7) / 8 - wib * KSP);  // k-steps of this warp that start below F (warp-uniform)
#pragma unroll
    for (int q = 0; q < KSP; q++) {
      if (q >= nq) break;
      float x[4] = {xa[q * 8], xb[q * 8], xa[q * 8 + 4], xb[q * 8 + 4]};
      if (MASK) {
        const uint32_t m0 = mask_window(sb, ia + q * 8), m1 = mask_window(sb, ib + q * 8);
        x[0] = keep_if(x[0], m0, 1u);
        x[1] = keep_if(x[1], m1, 1u);
        x[2] = keep_if(x[2], m0, 16u);
        x[3] = keep_if(x[3], m1, 16u);
      }
      uint32_t ah[4], al[4];
#pragma unroll
      for (int e = 0; e < 4; e++) split_trunc(x[e], ah[e], al[e]);
#pragma unroll
      for (int j = 0; j < 2; j++) {
        mma_tf32(c[j], al, bh[q][j][0], bh[q][j][1]);
        mma_tf32(c[j], ah, bl[q][j][0], bl[q][j][1]);
        mma_tf32(c[j], ah, bh[q][j][0], bh[q][j][1]);
      }
    }
    float *my = red + ((size_t)(it & 1) * 8 + wib) * 256;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      *reinterpret_cast<float2 *>(my + g * P + j * 8 + 2 * t) = make_float2(c[j][0], c[j][1]);
      *reinterpret_cast<float2 *>(my + (g + 8) * P + j * 8 + 2 * t) = make_float2(c[j][2], c[j][3]);
    }
    __syncthreads();  // partials visible; every warp is done with this stage's tile
    if (threadIdx.x == 0) {
      const int64_t hn = ht + (int64_t)kStages * stride;
      if (hn < nht) issue(s, hn);
    }
    {
      const float *r = red + (size_t)(it & 1) * 8 * 256 + threadIdx.x;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) sum += r[w * 256];  // ascending k ranges: fixed order
      const int row = threadIdx.x / P;
      if (row < nrows) out[(ht * HR + row) * P + (threadIdx.x % P)] = MASK ? sum * scale : sum;
    }
  }
}

// weight gradient, P = 16: dW[F x 16] = (X .* mask*scale)^T * dH.  m = features (warp w owns the 16-feature tiles
// w, w + 8, ... in registers for the whole kernel), k = rows (2 k-steps per half-tile), n = 16 columns.  Rows beyond N
// are neutralised on the dH side (zero B fragments); features >= F land in accumulators that are never stored.
template <bool MASK, int MT>
__global__ void __launch_bounds__(kT, 2)
dense_feat_tn_mma_kernel(const float *__restrict__ X, const uint32_t *__restrict__ bits, float scale,
                         const float *__restrict__ dH, float *__restrict__ ws, int64_t N, int F, const int kStages) {
  constexpr int P = 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int wpt = (int)mask_words_per_tile(F);
  const size_t tile_floats = ((size_t)HR * F + 3) & ~(size_t)3;
  float *tiles = reinterpret_cast<float *>(smem_raw);
  float *pad = tiles + kStages * tile_floats;              // 16 zero floats: the last row's reads past feature F
  float *dhs = pad + 16;                                   // kStages x 16 x 16
  uint32_t *mbits = reinterpret_cast<uint32_t *>(dhs + kStages * HR * P);
  uint64_t *bars = reinterpret_cast<uint64_t *>(mbits + (size_t)kStages * (wpt + 4));
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t nht = (N + HR - 1) / HR;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  for (size_t i = threadIdx.x; i < kStages * tile_floats + 16; i += kT) tiles[i] = 0.f;
  if (MASK)
    for (int i = threadIdx.x; i < kStages * (wpt + 4); i += kT) mbits[i] = 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&bars[s], MASK ? 3 : 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int s, int64_t ht) {
    const int nr = (int)min((int64_t)HR, N - ht * HR);
    issue_tile(tiles + s * tile_floats, X, ht * HR, nr, F, &bars[s]);
    issue_tile(dhs + s * HR * P, dH, ht * HR, nr, P, &bars[s]);
    if (MASK) {
      mbar_expect_tx(&bars[s], (uint32_t)wpt * 4);
      bulk_g2s(mbits + (size_t)s * (wpt + 4), bits + (ht >> 1) * wpt, (uint32_t)wpt * 4, &bars[s]);
    }
  };
  if (threadIdx.x == 0) {
    fence_proxy_async();
    for (int s = 0; s < kStages; s++) {
      const int64_t ht = first + s * stride;
      if (ht < nht) issue(s, ht);
    }
  }
  const int n_mt = (F + 15) / 16;
  float c[MT][2][4];
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) c[m][j][q] = 0.f;
  int it = 0;
  for (int64_t ht = first; ht < nht; ht += stride, it++) {
    const int s = it % kStages;
    const float *tile = tiles + s * tile_floats;
    const float *dh = dhs + s * HR * P;
    const uint32_t *sb = mbits + (size_t)s * (wpt + 4);
    mbar_wait(&bars[s], (uint32_t)((it / kStages) & 1));
    const int nrows = (int)min((int64_t)HR, N - ht * HR);
    const int mrow = (int)(ht & 1) * HR;
    // per half-tile partial sums (six mma per fragment from zero), then one rounded FADD into the running sums: the
    // tensor core's truncating accumulation would drift over the ~100 half-tiles a CTA walks
    float part[MT][2][4];
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) part[m][j][q] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      const int ra = ks * 8 + t, rb = ra + 4;  // rows (k) of this lane's fragment elements
      // B fragments: dH rows ra / rb, columns j*8 + g; rows beyond N contribute nothing
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int j = 0; j < 2; j++) {
        split_tf32(ra < nrows ? dh[ra * P + j * 8 + g] : 0.f, bh[j][0], bl[j][0]);
        split_tf32(rb < nrows ? dh[rb * P + j * 8 + g] : 0.f, bh[j][1], bl[j][1]);
      }
      const float *xa = tile + (size_t)ra * F + wib * 16 + g;  // (row ra, feature wib*16 + g)
      const float *xb = xa + (size_t)4 * F;                    // row rb
      const int ia = (mrow + ra) * F + wib * 16 + g, ib = ia + 4 * F;
#pragma unroll
      for (int m = 0; m < MT; m++) {
        if (wib + m * 8 < n_mt) {  // warp-uniform
          float x[4] = {xa[m * 128], xa[m * 128 + 8], xb[m * 128], xb[m * 128 + 8]};  // features fa, fa + 8
          if (MASK) {
            const uint32_t wa = mask_window(sb, ia + m * 128), wb = mask_window(sb, ib + m * 128);
            x[0] = keep_if(x[0], wa, 1u);
            x[1] = keep_if(x[1], wa, 256u);
            x[2] = keep_if(x[2], wb, 1u);
            x[3] = keep_if(x[3], wb, 256u);
          }
          uint32_t ah[4], al[4];
#pragma unroll
          for (int q = 0; q < 4; q++) split_trunc(x[q], ah[q], al[q]);
#pragma unroll
          for (int j = 0; j < 2; j++) {
            mma_tf32(part[m][j], al, bh[j][0], bh[j][1]);
            mma_tf32(part[m][j], ah, bl[j][0], bl[j][1]);
            mma_tf32(part[m][j], ah, bh[j][0], bh[j][1]);
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int q = 0; q < 4; q++) c[m][j][q] += part[m][j][q];
    __syncthreads();  // every warp is done with this stage
    if (threadIdx.x == 0) {
      const int64_t hn = ht + (int64_t)kStages * stride;
      if (hn < nht) issue(s, hn);
    }
  }
  float *dst = ws + (size_t)blockIdx.x * F * P;
  const float sc = MASK ? scale : 1.f;
#pragma unroll
  for (int m = 0; m < MT; m++) {
    const int mt = wib + m * 8;
    if (mt >= n_mt) continue;
    const int fa = mt * 16 + g, fb = fa + 8;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      if (fa < F)
        *reinterpret_cast<float2 *>(dst + (size_t)fa * P + j * 8 + 2 * t) = make_float2(c[m][j][0] * sc, c[m][j][1] * sc);
      if (fb < F)
        *reinterpret_cast<float2 *>(dst + (size_t)fb * P + j * 8 + 2 * t) = make_float2(c[m][j][2] * sc, c[m][j][3] * sc);
    }
  }
}

size_t fwd_mma_smem(int f, int ns) {
  const size_t tile_floats = ((size_t)HR * f + 3) & ~(size_t)3;
  return (ns * tile_floats + 8 + ns * (mask_words_per_tile(f) + 4) + 2 * 8 * 256) * 4 + ns * 8 + 16;
}
size_t tn_mma_smem(int f, int ns) {
  const size_t tile_floats = ((size_t)HR * f + 3) & ~(size_t)3;
  return (ns * tile_floats + 16 + ns * HR * 16 + ns * (mask_words_per_tile(f) + 4)) * 4 + ns * 8 + 16;
}
struct MmaShape {
  int stages = 0, ctas_per_sm = 0;
};
MmaShape mma_shape(int f) {
  MmaShape m;
  const size_t cap = 227 * 1024, per_cta_reserved = 1024;
  if (2 * (fwd_mma_smem(f, 2) + per_cta_reserved) <= cap && 2 * (tn_mma_smem(f, 2) + per_cta_reserved) <= cap) {
    m.stages = 2;
    m.ctas_per_sm = 2;
  } else if (fwd_mma_smem(f, 3) <= cap && tn_mma_smem(f, 3) <= cap) {
    m.stages = 3;
    m.ctas_per_sm = 1;
  }
  return m;
}
int mma_ctas(int64_t n, int f) {
  const int sm = std::max(1, device_info().sm_count);
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + HR - 1) / HR, (int64_t)sm * mma_shape(f).ctas_per_sm));
}
// the tensor-core kernels cover hidden 16 with up to 1024 features (MT <= 8 feature tiles per warp); GCNB_DENSE_MMA=0
// keeps the FMA kernels (A/B measurements)
bool use_mma(int f, int p) {
  static const bool on = [] {
    const char *e = getenv("GCNB_DENSE_MMA");
    return !(e && atoi(e) == 0);
  }();
  return on && p == 16 && f >= 8 && f <= 768 && mma_shape(f).stages > 0;
}

__global__ void __launch_bounds__(256) cta_partial_reduce_kernel(const float *__restrict__ ws, float *__restrict__ out, int64_t elems, int parts) {
  // 32 outputs x 8 groups per CTA: group g adds the partials g, g + 8, ... in ascending order, then the groups are added in
  // order -- a fixed tree (deterministic) with 8x the parallelism and 1/8 of the dependent chain of one thread per output
  // (9632 outputs x ~300 partials took 52 us that way)
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < elems; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + lane;
    float s = 0.f;
    if (i < elems)
      for (int z = g; z < parts; z += 8) s += __ldg(ws + (size_t)z * elems + i);
    red[g][lane] = s;
    __syncthreads();
    if (g == 0 && i < elems) {
      float t = red[0][lane];
#pragma unroll
      for (int k = 1; k < 8; k++) t += red[k][lane];
      out[i] = t;
    }
    __syncthreads();
  }
}

int persistent_ctas(int64_t n) {
  const int sm = std::max(1, device_info().sm_count);
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + TR - 1) / TR, (int64_t)sm));
}
size_t fwd_smem(int f, int p) {
  return ((size_t)2 * TR * f + (size_t)f * (p + 4) + 4 + 2 * (size_t)mask_words_per_tile(f)) * sizeof(float) + 2 * sizeof(uint64_t) + 16;
}
size_t tn_smem(int f, int p) {
  return ((size_t)2 * TR * f + (size_t)2 * TR * p + 2 * (size_t)mask_words_per_tile(f)) * sizeof(float) + 2 * sizeof(uint64_t) + 16;
}

}  // namespace

namespace {
const int patchables_registered = [] {
  register_patchable((const void *)dropout_maskbits_kernel, 6, -1);
  return 0;
}();
}  // namespace

extern "C" {

int gcnb_dense_feat_supported(int f, int p) {
  return (p == 8 || p == 16 || p == 32) && f >= 1 && f <= 4 * kT && fwd_smem(f, p) <= 227 * 1024 && tn_smem(f, p) <= 227 * 1024;
}

int64_t gcnb_dropout_maskbits_words(int64_t n_rows, int f) { return ((n_rows + TR - 1) / TR) * mask_words_per_tile(f); }

int gcnb_dropout_maskbits(uint32_t *d_bits, int64_t n_rows, int f, float p, const gcnb_rng_t *rng, gcnb_stream_t s) {
  if (!d_bits || !rng || n_rows < 0 || f <= 0 || rng->elem_lead > 3) return GCNB_E_BADARG;
  if (((uintptr_t)d_bits % 16) != 0) return GCNB_E_UNSUPPORTED;
  if (n_rows == 0) return 0;
  const int64_t wpt = mask_words_per_tile(f);
  const int64_t words = ((n_rows + TR - 1) / TR) * wpt;
  const int sm = std::max(1, device_info().sm_count);
  const int blocks = (int)std::min<int64_t>((words + kT - 1) / kT, (int64_t)sm * 16);
  dropout_maskbits_kernel<<<blocks, kT, 0, as_stream(s)>>>(d_bits, n_rows * f, f, wpt, words, p, *rng);
  GCNB_LAUNCH_CHECK();
  return 0;
}

int gcnb_dense_feat_fwd_f32(const float *d_X, const uint32_t *d_bits, float p_drop, const float *d_W, float *d_out,
                            int64_t n, int f, int p, gcnb_stream_t s) {
  if (!d_X || !d_W || !d_out || n < 0) return GCNB_E_BADARG;
  if (!gcnb_dense_feat_supported(f, p) || ((uintptr_t)d_X % 16) != 0) return GCNB_E_UNSUPPORTED;
  if (n == 0) return 0;
  const float scale = (float)(1.0 / (1.0 - p_drop));
  const int blocks = persistent_ctas(n);
  cudaStream_t st = as_stream(s);
  if (use_mma(f, p)) {
    const int ns = mma_shape(f).stages;
    const size_t sm = fwd_mma_smem(f, ns);
    const int mblocks = mma_ctas(n, f);
    const int ksp = (((f + 7) / 8) + 7) / 8;  // k-steps per warp
#define FWDM(MASKED, KK)                                                                                                   \
  do {                                                                                                                     \
    GCNB_SMEM_OPT_IN(sm, dense_feat_fwd_mma_kernel<MASKED, KK>); \
    dense_feat_fwd_mma_kernel<MASKED, KK><<<mblocks, kT, sm, st>>>(d_X, d_bits, scale, d_W, d_out, n, f, ns);                  \
  } while (0)
#define FWDMM(KK)                  \
  do {                             \
    if (d_bits) FWDM(true, KK);    \
    else FWDM(false, KK);          \
  } while (0)
    if (ksp <= 1) FWDMM(1);
    else if (ksp <= 2) FWDMM(2);
    else if (ksp <= 4) FWDMM(4);
    else if (ksp <= 6) FWDMM(6);
    else if (ksp <= 8) FWDMM(8);
    else if (ksp <= 10) FWDMM(10);
    else FWDMM(12);
#undef FWDMM
#undef FWDM
    GCNB_LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = fwd_smem(f, p);
#define FWD(PP)                                                                                                   \
  do {                                                                                                            \
    GCNB_SMEM_OPT_IN(smem, dense_feat_fwd_kernel<PP>); \
    dense_feat_fwd_kernel<PP><<<blocks, kT, smem, st>>>(d_X, d_bits, scale, d_W, d_out, n, f);                    \
  } while (0)
  if (p == 8) FWD(8);
  else if (p == 16) FWD(16);
  else FWD(32);
#undef FWD
  GCNB_LAUNCH_CHECK();
  return 0;
}

int64_t gcnb_dense_feat_tn_workspace(int64_t n, int f, int p) {
  const int ctas = use_mma(f, p) ? std::max(persistent_ctas(n), mma_ctas(n, f)) : persistent_ctas(n);
  return (int64_t)ctas * f * p * (int64_t)sizeof(float);
}

int gcnb_dense_feat_tn_f32(const float *d_X, const uint32_t *d_bits, float p_drop, const float *d_dH, float *d_dW,
                           int64_t n, int f, int p, void *d_ws, int64_t ws_bytes, gcnb_stream_t s) {
  if (!d_X || !d_dH || !d_dW || n < 0) return GCNB_E_BADARG;
  if (!gcnb_dense_feat_supported(f, p) || ((uintptr_t)d_dH % 16) != 0 || ((uintptr_t)d_X % 16) != 0)
    return GCNB_E_UNSUPPORTED;
  cudaStream_t st = as_stream(s);
  if (n == 0) {
    GCNB_CHECK(cudaMemsetAsync(d_dW, 0, (size_t)f * p * 4, st));
    return 0;
  }
  const int ctas = persistent_ctas(n);
  if (!d_ws || ws_bytes < (int64_t)ctas * f * p * 4 || ((uintptr_t)d_ws % 16) != 0) return GCNB_E_BADARG;
  const float scale = (float)(1.0 / (1.0 - p_drop));
  if (use_mma(f, p)) {
    const int ns = mma_shape(f).stages;
    const size_t sm = tn_mma_smem(f, ns);
    const int mctas = mma_ctas(n, f);
    if (ws_bytes < (int64_t)mctas * f * p * 4) return GCNB_E_BADARG;
    const int mt = ((f + 15) / 16 + 7) / 8;  // 16-feature tiles per warp
#define TNM(MASKED, MTT)                                                                                                    \
  do {                                                                                                                      \
    GCNB_SMEM_OPT_IN(sm, dense_feat_tn_mma_kernel<MASKED, MTT>); \
    dense_feat_tn_mma_kernel<MASKED, MTT><<<mctas, kT, sm, st>>>(d_X, d_bits, scale, d_dH, (float *)d_ws, n, f, ns);          \
  } while (0)
#define TNMM(MTT)                    \
  do {                               \
    if (d_bits) TNM(true, MTT);      \
    else TNM(false, MTT);            \
  } while (0)
    if (mt <= 1) TNMM(1);
    else if (mt <= 2) TNMM(2);
    else if (mt <= 3) TNMM(3);
    else if (mt <= 5) TNMM(5);
    else TNMM(8);
#undef TNMM
#undef TNM
    GCNB_LAUNCH_CHECK();
    const int64_t elems = (int64_t)f * p;
    cta_partial_reduce_kernel<<<(int)std::min<int64_t>((elems + 31) / 32, 4096), 256, 0, st>>>((const float *)d_ws, d_dW,
                                                                                                   elems, mctas);
    GCNB_LAUNCH_CHECK();
    return 0;
  }
  const int fpt = (f + kT - 1) / kT;
  const size_t smem = tn_smem(f, p);
#define TN(PP, FF)                                                                                                      \
  do {                                                                                                                  \
    GCNB_SMEM_OPT_IN(smem, dense_feat_tn_kernel<PP, FF>); \
    dense_feat_tn_kernel<PP, FF><<<ctas, kT, smem, st>>>(d_X, d_bits, scale, d_dH, (float *)d_ws, n, f);                  \
  } while (0)
#define TNP(PP)                 \
  do {                          \
    if (fpt == 1) TN(PP, 1);    \
    else if (fpt == 2) TN(PP, 2); \
    else if (fpt == 3) TN(PP, 3); \
    else TN(PP, 4);             \
  } while (0)
  if (p == 8) TNP(8);
  else if (p == 16) TNP(16);
  else TNP(32);
#undef TNP
#undef TN
  GCNB_LAUNCH_CHECK();
  const int64_t elems = (int64_t)f * p;
  cta_partial_reduce_kernel<<<(int)std::min<int64_t>((elems + 31) / 32, 4096), 256, 0, st>>>((const float *)d_ws, d_dW,
                                                                                                 elems, ctas);
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
