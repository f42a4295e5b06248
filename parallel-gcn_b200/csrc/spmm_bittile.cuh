// spmm_bittile.cuh -- the device plan of the bit-tile GraphSum (spmm_bittile.cu) and the pieces its two builders share:
// the host builder (bittile_build_host, CPU arrays in, unit-tested without a GPU) and the device builder
// (spmm_bittile_build.cu: the CSR never leaves the GPU).
#pragma once
#include <vector>

#include "common.cuh"
#include "spmm_plan.cuh"

namespace gcnb {

constexpr int kBtRows = 128;                 // rows per block = MMA M = TMEM lanes
constexpr int kBtChunk = 64;                 // columns per tile = 4 MMA k-steps of 16
constexpr int kBtN = 48;                     // MMA N: 3 bf16 pieces x 16 columns
constexpr int kBtKStepBytes = kBtN * 16 * 2; // 1536: one 48 x 16 bf16 operand
constexpr int kBtChunkBytes = 4 * kBtKStepBytes;  // 6144 bytes of packed B' per chunk
constexpr int kBtThreads = 14 * 32;

// CTA schedule of the MMA kernel: longest-processing-time greedy over the row blocks that own tiles (cost = tiles + a
// per-block constant for the epilogue and the pipeline drain); deterministic.  tiles_of_block[b] = tiles of row block b.
// Out: cta_tile_ptr / cta_item_ptr (n_cta + 1 each), items (block, end position of its tiles relative to the CTA's first
// tile), tile_base[b] = index of block b's first tile.  Returns the number of tiles, or -1 when it exceeds 32 bits.
int64_t bittile_schedule(const uint32_t *tiles_of_block, int64_t n_blk, int n_cta, int chunk_cols, int row_blocks,
                         std::vector<uint32_t> &cta_tile_ptr, std::vector<uint32_t> &cta_item_ptr, std::vector<uint2> &items,
                         std::vector<uint64_t> &tile_base);

}  // namespace gcnb

struct gcnb_bittile_plan {
  int64_t n_rows = 0, n_cols = 0, nnz = 0, n_blk = 0, n_tiles = 0, tile_nnz = 0, rem_nnz = 0, n_chunks = 0;
  int n_cta = 0;
  int chunk = gcnb::kBtChunk;  // columns per tile (64: bt_mma_wide_kernel<1, 1> / <2, 1>; 128: <1, 2>)
  int rb = 1;            // 128-row blocks per item (2: bt_mma_wide_kernel<2, 1>)
  int parts = 15;  // debugging: bit 0 pack, 1 MMA kernel, 2 remainder, 3 final add
  uint32_t *d_tile_chunk = nullptr, *d_cta_tile_ptr = nullptr, *d_cta_item_ptr = nullptr;
  uint2 *d_items = nullptr;
  uint64_t *d_bits = nullptr;
  uint32_t *d_r_indptr = nullptr, *d_r_indices = nullptr;
  float *d_r_values = nullptr, *d_row_scale = nullptr, *d_col_scale = nullptr;
  uint8_t *d_packed = nullptr;
  float *d_P = nullptr, *d_R = nullptr;
  gcnb_spmm_plan *rem = nullptr;  // valued remainder CSR on the generic kernel (entries that do not factor exist, or GCNB_BT_ELL=0)
  gcnb::EllDev *ell = nullptr;    // pattern-only remainder (spmm_ell.cu): every remainder entry factors
  float *d_B2 = nullptr;          // [n_cols + 1][16]: diag(col_scale) * B of the current launch, last row zero
  int rem_ctas = 0;               // CTAs per SM of the remainder kernel (0 = its default)
  int64_t n_unfactored = 0;       // entries whose value is not row_scale * col_scale (0: the matrix is a scaled pattern)
  uint32_t *d_perm = nullptr;     // gcnb_bittile_plan_set_permutation: plan index k = caller's row perm[k] (needs the merge path)
  int merge_by_reduction = 1;     // GCNB_BT_MERGE=0 (tuning probe): partial buffers + bt_add_kernel even with the ELL remainder
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

namespace gcnb {
// everything of gcnb_bittile_plan_create that does not depend on where the tiles were built: operand / partial buffers,
// second stream, events, kernel attributes.  Synchronises `stream`.
int bittile_finish_plan(gcnb_bittile_plan *p, cudaStream_t stream);
}  // namespace gcnb
