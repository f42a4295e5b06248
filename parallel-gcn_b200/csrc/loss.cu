// loss.cu -- masked softmax cross-entropy + gradient + accuracy in one pass (sm_100a).
// Replaces cross_entropy_loss_kernel (src/module.cu:484-524: one THREAD per node walking C floats at stride C,
// loss via warp shuffle + atomicAdd) and get_accuracy_kernel (src/gcn.cu:264-278, integer atomicAdd), plus the
// cudaMemsetAsync of the whole gradient (src/module.cu:528-529): here one WARP per node row (coalesced row
// reads), unlabelled rows get their zero gradient written by the same kernel, and loss / wrong / labelled
// counts are reduced per block and then in ascending block order by the last block to finish (deterministic).
#include <algorithm>

#include "common.cuh"

using namespace gcnb;

namespace {

constexpr int kT = 256;
constexpr int kWarpsPerBlock = kT / 32;
constexpr int kMaxPerLane = 8;  // supports num_classes <= 256 per pass; larger falls back to the looped kernel

template <int PER_LANE>
__global__ void __launch_bounds__(kT)
softmax_ce_kernel(float *__restrict__ logits, float *__restrict__ grad, const int32_t *__restrict__ truth, int64_t n,
                  int C, uint32_t num_samples, int training, float *__restrict__ result,
                  float *__restrict__ part_loss, uint32_t *__restrict__ part_cnt, unsigned int *__restrict__ ticket) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float loss = 0.f;
  uint32_t wrong = 0, labelled = 0;
  for (int64_t row = warp0; row < n; row += nwarps) {
    const int t = __ldg(truth + row);
    float *lg = logits + (size_t)row * C;
    if (t < 0) {
      if (training) {
        float *gr = grad + (size_t)row * C;
        for (int j = lane; j < C; j += 32) gr[j] = 0.f;
      }
      continue;
    }
    float x[PER_LANE];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < PER_LANE; k++) {
      const int j = lane + 32 * k;
      x[k] = (j < C) ? lg[j] : -INFINITY;
      mx = fmaxf(mx, x[k]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < PER_LANE; k++) {
      const int j = lane + 32 * k;
      if (j < C) {
        x[k] -= mx;  // numerical stability; written back (API-visible side effect of the reference)
        lg[j] = x[k];
        se += expf(x[k]);
      }
    }
    se = warp_sum(se);
    // shifted truth logit lives in lane t % 32, slot t / 32
    float xt = 0.f;
#pragma unroll
    for (int k = 0; k < PER_LANE; k++)
      if (k == (t >> 5)) xt = x[k];
    xt = __shfl_sync(0xffffffffu, xt, t & 31);
    if (lane == 0) {
      loss += logf(se) - xt;
      labelled++;
      if (xt < 0.f) wrong++;  // src/gcn.cu:273-276
    }
    if (training) {
      float *gr = grad + (size_t)row * C;
#pragma unroll
      for (int k = 0; k < PER_LANE; k++) {
        const int j = lane + 32 * k;
        if (j < C) {
          const float prob = expf(x[k]) / se;
          float gv = prob / (float)num_samples;
          if (j == t) gv = (float)(gv - 1.0 / (double)num_samples);  // double literal in the reference (:517)
          gr[j] = gv;
        }
      }
    }
  }
  // block reduction in warp order, then ascending block order by the last block
  __shared__ float s_loss[kWarpsPerBlock];
  __shared__ uint32_t s_wrong[kWarpsPerBlock], s_lab[kWarpsPerBlock];
  __shared__ bool last;
  if (lane == 0) {
    s_loss[wib] = loss;
    s_wrong[wib] = wrong;
    s_lab[wib] = labelled;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bl = 0.f;
    uint32_t bw = 0, bb = 0;
    for (int i = 0; i < kWarpsPerBlock; i++) {
      bl += s_loss[i];
      bw += s_wrong[i];
      bb += s_lab[i];
    }
    part_loss[blockIdx.x] = bl;
    part_cnt[2 * blockIdx.x] = bw;
    part_cnt[2 * blockIdx.x + 1] = bb;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float tl = 0.f;
    uint32_t tw = 0, tb = 0;
    for (unsigned i = 0; i < gridDim.x; i++) {
      tl += ((volatile float *)part_loss)[i];
      tw += ((volatile uint32_t *)part_cnt)[2 * i];
      tb += ((volatile uint32_t *)part_cnt)[2 * i + 1];
    }
    result[0] = tl;
    result[1] = __uint_as_float(tw);
    result[2] = __uint_as_float(tb);
    *ticket = 0;
  }
}


// ---- lane-per-row variant for narrow outputs (C <= kMaxRowC): a warp stages a tile of 32 consecutive rows in shared
// memory with coalesced loads, lane l then owns row l (row stride odd: conflict-free), and the shifted logits / the
// gradient tile leave through coalesced stores again.  ~12x fewer instructions per row than the warp-per-row kernel
// at C = 41 (no shuffles, no idle lanes, one exponential per element).
constexpr int kMaxRowC = 64;

__global__ void __launch_bounds__(kT)
softmax_ce_rows_kernel(float *__restrict__ logits, float *__restrict__ grad, const int32_t *__restrict__ truth, int64_t n,
                       int C, uint32_t num_samples, int training, int aligned16, float *__restrict__ result,
                       float *__restrict__ part_loss, uint32_t *__restrict__ part_cnt, unsigned int *__restrict__ ticket) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int CP = C | 1;  // odd row stride
  const bool contiguous = CP == C && aligned16;  // tile is one 16-byte-aligned block on both sides: 128-bit copies
  float *sx = smem + (size_t)wib * 2 * 32 * CP;  // shifted logits of the warp's tile
  float *se_ = sx + 32 * CP;                     // exponentials -> gradient
  const int64_t ntiles = (n + 31) / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const float inv_ns = 1.0f / (float)num_samples;
  float loss = 0.f;
  uint32_t wrong = 0, labelled = 0;
  for (int64_t tile = warp0; tile < ntiles; tile += nwarps) {
    const int64_t r0 = tile * 32;
    const int rows = (int)min((int64_t)32, n - r0);
    const int total = rows * C;
    float *lg = logits + (size_t)r0 * C;
    // coalesced copy-in (global element i of the tile -> row i / C, column i % C); independent loads, unrolled
    if (contiguous) {  // odd C: the tile is one contiguous block in shared memory too
      const float4 *src = reinterpret_cast<const float4 *>(lg);
      float4 *dst = reinterpret_cast<float4 *>(sx);
      const int n4 = total >> 2;
#pragma unroll 4
      for (int i = lane; i < n4; i += 32) dst[i] = src[i];
      for (int i = (n4 << 2) + lane; i < total; i += 32) sx[i] = lg[i];
    } else {
#pragma unroll 4
      for (int i = lane; i < total; i += 32) {
        const int row = i / C;
        sx[row * CP + (i - row * C)] = lg[i];
      }
    }
    __syncwarp();
    const int t = (lane < rows) ? __ldg(truth + r0 + lane) : -1;
    float *xr = sx + lane * CP, *er = se_ + lane * CP;
    if (t >= 0) {
      float mx = -INFINITY;
      for (int j = 0; j < C; j++) mx = fmaxf(mx, xr[j]);
      float sum = 0.f;
      for (int j = 0; j < C; j++) {
        const float x = xr[j] - mx;  // numerical stability; written back (API-visible side effect of the reference)
        xr[j] = x;
        const float e = expf(x);
        er[j] = e;
        sum += e;
      }
      const float xt = xr[t];
      loss += logf(sum) - xt;
      labelled++;
      if (xt < 0.f) wrong++;  // src/gcn.cu:273-276
      if (training) {
        const float inv = 1.0f / sum;
        for (int j = 0; j < C; j++) er[j] = (er[j] * inv) * inv_ns;
        er[t] = (float)((double)er[t] - 1.0 / (double)num_samples);  // double literal in the reference (:517)
      }
    } else if (training && lane < rows) {
      for (int j = 0; j < C; j++) er[j] = 0.f;  // unlabelled rows: zero gradient (the reference memsets it)
    }
    __syncwarp();
    // coalesced copy-out: shifted logits (rows without a label are rewritten unchanged), then the gradient tile
    {
      float *gr = training ? grad + (size_t)r0 * C : nullptr;
      if (contiguous) {
        const int n4 = total >> 2;
        float4 *d0 = reinterpret_cast<float4 *>(lg);
        const float4 *s0 = reinterpret_cast<const float4 *>(sx);
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) d0[i] = s0[i];
        for (int i = (n4 << 2) + lane; i < total; i += 32) lg[i] = sx[i];
        if (training) {
          float4 *d1 = reinterpret_cast<float4 *>(gr);
          const float4 *s1 = reinterpret_cast<const float4 *>(se_);
#pragma unroll 4
          for (int i = lane; i < n4; i += 32) d1[i] = s1[i];
          for (int i = (n4 << 2) + lane; i < total; i += 32) gr[i] = se_[i];
        }
      } else {
#pragma unroll 4
        for (int i = lane; i < total; i += 32) {
          const int row = i / C;
          const int o = row * CP + (i - row * C);
          lg[i] = sx[o];
          if (training) gr[i] = se_[o];
        }
      }
    }
    __syncwarp();
  }
  // lanes -> warp (fixed shuffle tree) -> block in warp order -> ascending block order by the last block
  loss = warp_sum(loss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
    labelled += __shfl_xor_sync(0xffffffffu, labelled, o);
  }
  __shared__ float s_loss[kWarpsPerBlock];
  __shared__ uint32_t s_wrong[kWarpsPerBlock], s_lab[kWarpsPerBlock];
  __shared__ bool last;
  if (lane == 0) {
    s_loss[wib] = loss;
    s_wrong[wib] = wrong;
    s_lab[wib] = labelled;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bl = 0.f;
    uint32_t bw = 0, bb = 0;
    for (int i = 0; i < kWarpsPerBlock; i++) {
      bl += s_loss[i];
      bw += s_wrong[i];
      bb += s_lab[i];
    }
    part_loss[blockIdx.x] = bl;
    part_cnt[2 * blockIdx.x] = bw;
    part_cnt[2 * blockIdx.x + 1] = bb;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float tl = 0.f;
    uint32_t tw = 0, tb = 0;
    for (unsigned i = 0; i < gridDim.x; i++) {
      tl += ((volatile float *)part_loss)[i];
      tw += ((volatile uint32_t *)part_cnt)[2 * i];
      tb += ((volatile uint32_t *)part_cnt)[2 * i + 1];
    }
    result[0] = tl;
    result[1] = __uint_as_float(tw);
    result[2] = __uint_as_float(tb);
    *ticket = 0;
  }
}

int ce_blocks(int64_t n) {
  const int sm = std::max(1, device_info().sm_count);
  const int64_t want = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm * 8));
}
constexpr int kMaxCeBlocks = 148 * 8 * 2;

}  // namespace

extern "C" {

int64_t gcnb_ce_workspace(int64_t n) {
  (void)n;
  return 16 + (int64_t)kMaxCeBlocks * 12;
}

int gcnb_softmax_ce_f32(float *d_logits, float *d_grad, const int32_t *d_truth, int64_t n, int num_classes,
                        uint32_t num_samples, int training, float *d_result, void *d_ws, gcnb_stream_t s) {
  if (!d_logits || !d_truth || !d_result || !d_ws || n < 0 || num_classes <= 0 || (training && !d_grad))
    return GCNB_E_BADARG;
  if (num_classes > 32 * kMaxPerLane) return GCNB_E_UNSUPPORTED;
  int blocks = ce_blocks(n);
  if (blocks > kMaxCeBlocks) blocks = kMaxCeBlocks;
  unsigned int *ticket = (unsigned int *)d_ws;
  float *part_loss = (float *)d_ws + 4;
  uint32_t *part_cnt = (uint32_t *)d_ws + 4 + kMaxCeBlocks;
  cudaStream_t st = as_stream(s);
  if (num_classes <= kMaxRowC) {
    const size_t smem = (size_t)kWarpsPerBlock * 2 * 32 * (num_classes | 1) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      GCNB_CHECK(cudaFuncSetAttribute(softmax_ce_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      kWarpsPerBlock * 2 * 32 * (kMaxRowC | 1) * (int)sizeof(float)));
      attr_set = true;
    }
    const int sm = std::max(1, device_info().sm_count);
    const int64_t want = ((n + 31) / 32 + kWarpsPerBlock - 1) / kWarpsPerBlock;
    blocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm * 2));
    const int aligned16 = (((uintptr_t)d_logits | (uintptr_t)d_grad) % 16) == 0;
    softmax_ce_rows_kernel<<<blocks, kT, smem, st>>>(d_logits, d_grad, d_truth, n, num_classes, num_samples, training,
                                                     aligned16, d_result, part_loss, part_cnt, ticket);
    GCNB_LAUNCH_CHECK();
    return 0;
  }
#define CE_LAUNCH(PL)                                                                                            \
  softmax_ce_kernel<PL><<<blocks, kT, 0, st>>>(d_logits, d_grad, d_truth, n, num_classes, num_samples, training, \
                                               d_result, part_loss, part_cnt, ticket)
  if (num_classes <= 32) CE_LAUNCH(1);
  else if (num_classes <= 64) CE_LAUNCH(2);
  else if (num_classes <= 128) CE_LAUNCH(4);
  else CE_LAUNCH(8);
#undef CE_LAUNCH
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
