// spmm_stage.cu -- window-staged GraphSum for sm_100a: the part of  C = A_csr * B  whose neighbour rows cluster in
// column windows is served from SHARED MEMORY instead of L1/L2.
//
// Why (ncu, profiles/): at dim = 16 the generic segment kernel (spmm.cu) reads HBM exactly once (949 MB per launch
// on the Reddit-shape graph) but runs at 0.22 of the HBM roofline, because every gathered 64-byte neighbour row costs
// one L1 wavefront and the LSU data pipe (1 wavefront / clock / SM) saturates at ~80 %: 115 M gathers over 148 SMs is
// >= 410 us before any other instruction, plus the shuffles that broadcast ids and values (same pipe).  A row fetched
// from shared memory with conflict-free LDS.128 costs HALF a wavefront, the 2-byte window-local id and the value
// arrive in coalesced per-lane loads (no shuffles), and nothing is reduced across lanes.
//
// How: columns are cut into windows of window_rows rows of B (192 KB at dim 16).  The plan (host, built once,
// spmm_plan.cuh) collects per window the (row, window) runs with >= min_seg entries as SEGMENTS, sorts them by length
// and deals them 32 at a time into BUNDLES stored ELL-style.  A persistent CTA per SM walks its runs: a TMA bulk copy
// brings the window into shared memory, then each of its 32 warps takes bundles; lane l owns segment l: per step it
// reads its next (id, value), fetches the 64-byte neighbour row with four LDS.128 (column chunk c ^ (l & 3), so a
// quarter-warp touches all 32 banks once) and accumulates 16 private sums; at the end it stores its partial row.
// Whatever is not staged (the REMAINDER CSR) goes through the generic kernel, and a last kernel adds each row's
// partials in a fixed order -- no floating-point atomics, bit-reproducible.
//
// Reference being replaced: graphsum_kernel, src/module.cu:172-186 (one thread per output element, serial over the
// row, every lane re-reading the row's indices).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "bulk.cuh"
#include "common.cuh"
#include "spmm_plan.cuh"

using namespace gcnb;

// =====================================================================================================================
// host build
// =====================================================================================================================
namespace gcnb {

namespace stage_detail {

template <class F>
void run_threads(int T, F f) {
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back([&f, t] { f(t); });
  f(0);
  for (auto &x : th) x.join();
}

struct WinDiv {  // exact j / d for 32-bit j (64-bit magic multiplier)
  unsigned __int128 magic;
  explicit WinDiv(uint32_t d) : magic((((unsigned __int128)1) << 64) / d + 1) {}
  uint32_t operator()(uint32_t j) const { return (uint32_t)(((unsigned __int128)j * magic) >> 64); }
};

inline uint32_t pieces_of(uint32_t k, uint32_t cap) { return (k + cap - 1) / cap; }
inline uint32_t piece_begin(uint32_t k, uint32_t pieces, uint32_t p) { return (uint32_t)((uint64_t)k * p / pieces); }

struct TmpSeg {  // segment before bundling
  uint32_t row, len, ordinal, ent_off;
};

}  // namespace stage_detail
using namespace stage_detail;

int stage_build_host(const uint32_t *indptr, const uint32_t *indices, int64_t n_rows, int64_t n_cols,
                     const StageParams &P, StagedHost &H) {
  if (!indptr || n_rows < 0 || n_cols <= 0 || P.dim <= 0 || P.dim % 4) return GCNB_E_BADARG;
  int wc = P.window_rows;
  if (wc <= 0) wc = (int)std::min<int64_t>(65536, (192 * 1024) / ((int64_t)P.dim * 4));
  if (wc <= 0 || wc > 65536) return GCNB_E_BADARG;
  const int64_t n_win64 = (n_cols + wc - 1) / wc;
  if (n_win64 > (1 << 22)) return GCNB_E_UNSUPPORTED;
  const int n_win = (int)n_win64;
  const int64_t nnz = n_rows ? indptr[n_rows] : 0;
  const uint32_t min_seg = (uint32_t)std::max(1, P.min_seg);
  const uint32_t seg_cap = (uint32_t)std::min(65535, std::max(4, P.seg_cap));
  const int64_t min_window_nnz = P.min_window_nnz > 0 ? P.min_window_nnz : (int64_t)8 * wc;
  int T = P.n_threads > 0 ? P.n_threads : host_threads();
  if (nnz < (1 << 16)) T = 1;
  const WinDiv win_of((uint32_t)wc);

  const bool verbose = getenv("GCNB_STAGE_VERBOSE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!verbose) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[stage] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  H = StagedHost();
  H.dim = P.dim;
  H.window_rows = wc;
  H.n_win = n_win;
  H.n_cta = std::max(1, P.n_cta);
  H.n_rows = n_rows;
  H.n_cols = n_cols;
  H.nnz = nnz;

  // contiguous row ranges of equal nnz per thread
  std::vector<int64_t> rr((size_t)T + 1, n_rows);
  rr[0] = 0;
  for (int t = 1; t < T; t++) {
    const uint32_t target = (uint32_t)((uint64_t)nnz * t / T);
    rr[t] = std::lower_bound(indptr, indptr + n_rows + 1, target) - indptr;
    rr[t] = std::min<int64_t>(std::max(rr[t], rr[t - 1]), n_rows);
  }

  // ---- pass A: which windows are worth a copy --------------------------------------------------------------------
  std::vector<int64_t> win_tot((size_t)T * n_win, 0);
  run_threads(T, [&](int t) {
    std::vector<uint32_t> cnt((size_t)n_win, 0), touched;
    int64_t *tot = win_tot.data() + (size_t)t * n_win;
    for (int64_t r = rr[t]; r < rr[t + 1]; r++) {
      touched.clear();
      for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
        const uint32_t w = win_of(indices[e]);
        if (cnt[w]++ == 0) touched.push_back(w);
      }
      for (uint32_t w : touched) {
        if (cnt[w] >= min_seg) tot[w] += cnt[w];
        cnt[w] = 0;
      }
    }
  });
  lap("pass A (window totals)");
  std::vector<uint8_t> enabled((size_t)n_win, 0);
  bool any = false;
  for (int w = 0; w < n_win; w++) {
    int64_t s = 0;
    for (int t = 0; t < T; t++) s += win_tot[(size_t)t * n_win + w];
    enabled[w] = s >= min_window_nnz;
    any |= enabled[w] != 0;
  }

  // ---- pass B: sizes ---------------------------------------------------------------------------------------------
  H.row_slot.assign((size_t)n_rows + 1, 0);
  H.r_indptr.assign((size_t)n_rows + 1, 0);
  std::vector<uint32_t> nseg((size_t)T * n_win, 0);
  std::vector<uint64_t> nent((size_t)T * n_win, 0);
  run_threads(T, [&](int t) {
    std::vector<uint32_t> cnt((size_t)n_win, 0), touched;
    uint32_t *ns = nseg.data() + (size_t)t * n_win;
    uint64_t *ne = nent.data() + (size_t)t * n_win;
    for (int64_t r = rr[t]; r < rr[t + 1]; r++) {
      touched.clear();
      const uint32_t deg = indptr[r + 1] - indptr[r];
      if (any)
        for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
          const uint32_t w = win_of(indices[e]);
          if (cnt[w]++ == 0) touched.push_back(w);
        }
      uint32_t staged = 0, slots = 0;
      for (uint32_t w : touched) {
        const uint32_t k = cnt[w];
        cnt[w] = 0;
        if (!enabled[w] || k < min_seg) continue;
        const uint32_t pieces = pieces_of(k, seg_cap);
        ns[w] += pieces;
        ne[w] += k;
        slots += pieces;
        staged += k;
      }
      H.row_slot[(size_t)r + 1] = slots;
      H.r_indptr[(size_t)r + 1] = deg - staged;
    }
  });
  lap("pass B (sizes)");
  for (int64_t r = 0; r < n_rows; r++) {
    H.row_slot[(size_t)r + 1] += H.row_slot[(size_t)r];
    H.r_indptr[(size_t)r + 1] += H.r_indptr[(size_t)r];
  }
  H.n_segs = n_rows ? H.row_slot[(size_t)n_rows] : 0;
  const int64_t rem_nnz = n_rows ? H.r_indptr[(size_t)n_rows] : 0;
  H.staged_nnz = nnz - rem_nnz;
  if (H.staged_nnz > 0xfffffff0ll) return GCNB_E_UNSUPPORTED;

  // offsets: window-major, thread-minor (threads own ascending row ranges => rows ascending inside a window)
  std::vector<uint32_t> seg_off((size_t)T * n_win), ent_off((size_t)T * n_win), win_seg_begin((size_t)n_win + 1, 0);
  {
    uint64_t so = 0, eo = 0;
    for (int w = 0; w < n_win; w++) {
      win_seg_begin[w] = (uint32_t)so;
      for (int t = 0; t < T; t++) {
        seg_off[(size_t)t * n_win + w] = (uint32_t)so;
        ent_off[(size_t)t * n_win + w] = (uint32_t)eo;
        so += nseg[(size_t)t * n_win + w];
        eo += nent[(size_t)t * n_win + w];
      }
    }
    win_seg_begin[n_win] = (uint32_t)so;
  }
  HostArray<TmpSeg> tseg;
  HostArray<uint16_t> tcol;
  HostArray<uint32_t> tent;
  tseg.alloc((size_t)H.n_segs);
  tcol.alloc((size_t)H.staged_nnz);
  tent.alloc((size_t)H.staged_nnz);
  H.r_indices.alloc((size_t)rem_nnz);
  H.r_perm.alloc((size_t)rem_nnz);

  lap("alloc temporaries");
  // ---- pass C: segments (entries in even/odd alternating order) + remainder CSR -------------------------------------
  run_threads(T, [&](int t) {
    std::vector<uint32_t> cnt((size_t)n_win, 0), start((size_t)n_win, 0), touched, staged_w;
    std::vector<uint32_t> buf_col, buf_e, ev, od;
    std::vector<uint32_t> so(seg_off.begin() + (size_t)t * n_win, seg_off.begin() + (size_t)(t + 1) * n_win);
    std::vector<uint32_t> eo(ent_off.begin() + (size_t)t * n_win, ent_off.begin() + (size_t)(t + 1) * n_win);
    for (int64_t r = rr[t]; r < rr[t + 1]; r++) {
      const uint32_t b = indptr[r], e_end = indptr[r + 1];
      uint32_t rpos = H.r_indptr[(size_t)r];
      if (H.row_slot[(size_t)r + 1] == H.row_slot[(size_t)r]) {  // nothing staged: the row is copied as is
        for (uint32_t e = b; e < e_end; e++, rpos++) {
          H.r_indices[rpos] = indices[e];
          H.r_perm[rpos] = e;
        }
        continue;
      }
      touched.clear();
      for (uint32_t e = b; e < e_end; e++) {
        const uint32_t w = win_of(indices[e]);
        if (cnt[w]++ == 0) touched.push_back(w);
      }
      staged_w.clear();
      for (uint32_t w : touched)
        if (enabled[w] && cnt[w] >= min_seg) staged_w.push_back(w);
      std::sort(staged_w.begin(), staged_w.end());
      uint32_t tot = 0;
      for (uint32_t w : staged_w) {
        start[w] = tot;
        tot += cnt[w];
      }
      buf_col.resize(tot);
      buf_e.resize(tot);
      // staged entries grouped by window (original order inside), the rest to the remainder CSR
      for (uint32_t e = b; e < e_end; e++) {
        const uint32_t j = indices[e];
        const uint32_t w = win_of(j);
        if (enabled[w] && cnt[w] >= min_seg) {
          const uint32_t pos = start[w]++;
          buf_col[pos] = j - w * (uint32_t)wc;
          buf_e[pos] = e;
        } else {
          H.r_indices[rpos] = j;
          H.r_perm[rpos] = e;
          rpos++;
        }
      }
      uint32_t ordinal = 0;
      for (uint32_t w : staged_w) {
        const uint32_t k = cnt[w];
        const uint32_t first = start[w] - k;  // start[] was advanced by k
        const uint32_t pieces = pieces_of(k, seg_cap);
        for (uint32_t p = 0; p < pieces; p++) {
          const uint32_t pb = piece_begin(k, pieces, p), pe = piece_begin(k, pieces, p + 1), len = pe - pb;
          ev.clear();
          od.clear();
          for (uint32_t i = first + pb; i < first + pe; i++) ((buf_col[i] & 1u) ? od : ev).push_back(i);
          const uint32_t o0 = eo[w];
          size_t ie = 0, io = 0;
          for (uint32_t q = 0; q < len; q++) {
            bool take_even;
            if (ie < ev.size() && io < od.size()) take_even = (q & 1u) == 0;
            else take_even = ie < ev.size();
            const uint32_t i = take_even ? ev[ie++] : od[io++];
            tcol[o0 + q] = (uint16_t)buf_col[i];
            tent[o0 + q] = buf_e[i];
          }
          tseg[so[w]++] = TmpSeg{(uint32_t)r, len, ordinal++, o0};
          eo[w] += len;
        }
      }
      for (uint32_t w : touched) cnt[w] = 0;
    }
  });

  lap("pass C (segments+remainder)");
  // ---- phase 2a: per window, sort segments by length and count bundles / blocks --------------------------------------
  std::vector<uint32_t> order((size_t)H.n_segs);
  std::iota(order.begin(), order.end(), 0u);
  std::vector<uint32_t> win_bundles((size_t)n_win + 1, 0);
  std::vector<uint64_t> win_blocks((size_t)n_win + 1, 0);
  {
    std::atomic<int> next(0);
    run_threads(T, [&](int) {
      for (;;) {
        const int w = next.fetch_add(1);
        if (w >= n_win) break;
        const uint32_t sb = win_seg_begin[w], se = win_seg_begin[w + 1];
        if (sb == se) continue;
        std::sort(order.begin() + sb, order.begin() + se, [&](uint32_t a, uint32_t c) {
          return tseg[a].len != tseg[c].len ? tseg[a].len > tseg[c].len : a < c;
        });
        uint64_t blocks = 0;
        for (uint32_t i = sb; i < se; i += kStageLanes) blocks += (tseg[order[i]].len + kStageBlock - 1) / kStageBlock;
        win_bundles[w + 1] = (se - sb + kStageLanes - 1) / kStageLanes;
        win_blocks[w + 1] = blocks;
      }
    });
  }
  for (int w = 0; w < n_win; w++) {
    win_bundles[w + 1] += win_bundles[w];
    win_blocks[w + 1] += win_blocks[w];
  }
  lap("phase 2a (sort)");
  const uint64_t n_bundles = win_bundles[n_win];
  H.n_blocks = (int64_t)win_blocks[n_win];
  if ((uint64_t)H.n_blocks > 0xfffffff0ull || n_bundles * kStageLanes > 0xfffffff0ull) return GCNB_E_UNSUPPORTED;
  H.n_slots = H.n_segs;
  H.bundles.assign((size_t)n_bundles, make_uint4(0, 0, 0, 0));
  H.lens.assign((size_t)n_bundles * kStageLanes, 0);
  H.lane_slot.assign((size_t)n_bundles * kStageLanes, kStagePad);
  H.pidx.alloc((size_t)H.n_blocks * kStageLanes * kStageBlock);   // padding written per bundle in phase 2b
  H.pperm.alloc((size_t)H.n_blocks * kStageLanes * kStageBlock);

  lap("alloc outputs");
  // ---- phase 2b: fill bundles -----------------------------------------------------------------------------------------
  {
    std::atomic<int> next(0);
    run_threads(T, [&](int) {
      for (;;) {
        const int w = next.fetch_add(1);
        if (w >= n_win) break;
        const uint32_t sb = win_seg_begin[w], se = win_seg_begin[w + 1];
        uint64_t blk = win_blocks[w];
        uint32_t bundle = win_bundles[w];
        for (uint32_t i0 = sb; i0 < se; i0 += kStageLanes, bundle++) {
          const uint32_t nl = std::min<uint32_t>(kStageLanes, se - i0);
          const uint32_t L = tseg[order[i0]].len;
          const uint32_t minL = nl == kStageLanes ? tseg[order[i0 + nl - 1]].len : 0;
          H.bundles[bundle] = make_uint4((uint32_t)blk, L, minL, 0);
          {
            const size_t b0 = (size_t)blk * kStageLanes * kStageBlock;
            const size_t cnt = (size_t)((L + kStageBlock - 1) / kStageBlock) * kStageLanes * kStageBlock;
            std::fill(H.pidx.data() + b0, H.pidx.data() + b0 + cnt, (uint16_t)0);
            std::fill(H.pperm.data() + b0, H.pperm.data() + b0 + cnt, kStagePad);
          }
          // lane assignment: lanes l and l + 4 of a quarter-warp read the same 16-byte column chunk, so they should
          // fetch rows of opposite parity.  Inside a segment entries alternate even/odd (pass C) and lanes with bit 2
          // set start on odd; the unpaired tail of a segment is all-even or all-odd, so segments with an even surplus
          // go to bit2 = 0 lanes and odd-surplus ones to their bit2 = 1 partners (similar lengths: both lists are
          // length-sorted).
          uint32_t lane_of[kStageLanes], n_even_of[kStageLanes];
          {
            uint32_t elist[kStageLanes], olist[kStageLanes], ne = 0, no = 0;
            for (uint32_t i = 0; i < nl; i++) {
              const TmpSeg &sg = tseg[order[i0 + i]];
              uint32_t n_even = 0;
              for (uint32_t k = 0; k < sg.len; k++) n_even += (tcol[sg.ent_off + k] & 1u) == 0;
              n_even_of[i] = n_even;
              if (2 * n_even >= sg.len) elist[ne++] = i; else olist[no++] = i;
            }
            bool used[kStageLanes] = {false};
            auto lane0 = [](uint32_t i) { return (i >> 2) * 8 + (i & 3); };  // i-th lane with bit 2 clear
            uint32_t ie = 0, io = 0;
            for (uint32_t i = 0; i < 16 && ie < ne; i++, ie++) { lane_of[elist[ie]] = lane0(i); used[lane0(i)] = true; }
            for (uint32_t i = 0; i < 16 && io < no; i++, io++) { lane_of[olist[io]] = lane0(i) + 4; used[lane0(i) + 4] = true; }
            uint32_t free_lane = 0;
            auto next_free = [&]() { while (used[free_lane]) free_lane++; used[free_lane] = true; return free_lane; };
            for (; ie < ne; ie++) lane_of[elist[ie]] = next_free();
            for (; io < no; io++) lane_of[olist[io]] = next_free();
          }
          for (uint32_t i = 0; i < nl; i++) {
            const TmpSeg &sg = tseg[order[i0 + i]];
            const uint32_t l = lane_of[i];
            const size_t bl = (size_t)bundle * kStageLanes + l;
            H.lens[bl] = (uint16_t)sg.len;
            H.lane_slot[bl] = H.row_slot[sg.row] + sg.ordinal;
            const uint32_t paired = 2 * std::min(n_even_of[i], sg.len - n_even_of[i]);
            const bool flip = (l >> 2) & 1u;  // start on the other parity: swap the members of every (even, odd) pair
            for (uint32_t k = 0; k < sg.len; k++) {
              const uint32_t src = (flip && k < paired) ? (k ^ 1u) : k;
              const size_t pos = ((size_t)(blk + k / kStageBlock) * kStageLanes + l) * kStageBlock + k % kStageBlock;
              H.pidx[pos] = tcol[sg.ent_off + src];
              H.pperm[pos] = tent[sg.ent_off + src];
            }
          }
          blk += (L + kStageBlock - 1) / kStageBlock;
        }
      }
    });
  }

  lap("phase 2b (fill bundles)");
  // ---- runs and per-CTA queues: contiguous, equal estimated cycles ---------------------------------------------------
  // Two classes of windows when the caller names an OWN column range (row-partitioned GraphSum: the rank's own slab of B
  // is available before the slabs of its peers have arrived): windows entirely inside it form a second run list that
  // can be processed while the exchange is still in flight.
  auto build_runs = [&](bool own_class, std::vector<uint4> &runs, std::vector<uint32_t> &run_begin) {
    const double c_step = 18.0, c_bundle = 80.0, c_load = 6000.0;
    auto bundle_cost = [&](const uint4 &b) { return c_bundle + c_step * b.y; };
    auto in_class = [&](int w) {
      const int64_t c0 = (int64_t)w * wc, c1 = std::min<int64_t>(n_cols, c0 + wc);
      const bool own = P.own_col1 > P.own_col0 && c0 >= P.own_col0 && c1 <= P.own_col1;
      return own == own_class;
    };
    double total = 0;
    int n_used = 0;
    for (int w = 0; w < n_win; w++) {
      if (win_bundles[w + 1] == win_bundles[w] || !in_class(w)) continue;
      n_used++;
      for (uint32_t i = win_bundles[w]; i < win_bundles[w + 1]; i++) total += bundle_cost(H.bundles[i]);
    }
    const int Q = H.n_cta;
    total += c_load * (n_used + Q);
    // a queue is cut into ~kRunsPerQueue runs so that a CTA that finishes early can take over whole runs of another
    const double max_run = std::max(4.0 * c_load, total / Q / std::max(1, P.runs_per_queue));
    runs.clear();
    run_begin.assign((size_t)Q + 1, 0);
    int q = 0;
    double acc = 0;
    for (int w = 0; w < n_win; w++) {
      uint32_t sb = win_bundles[w];
      const uint32_t se = win_bundles[w + 1];
      if (sb == se || !in_class(w)) continue;
      acc += c_load;
      double run_acc = 0;
      for (uint32_t i = sb; i < se; i++) {
        const double c = bundle_cost(H.bundles[i]);
        acc += c;
        run_acc += c;
        const bool cut_queue = q < Q - 1 && acc >= total * (q + 1) / Q && i + 1 < se;
        if (cut_queue || (run_acc >= max_run && i + 1 < se)) {
          runs.push_back(make_uint4((uint32_t)w, sb, i + 1, 0));
          sb = i + 1;
          run_acc = 0;
          if (cut_queue) {
            run_begin[++q] = (uint32_t)runs.size();
            acc += c_load;
          }
        }
      }
      runs.push_back(make_uint4((uint32_t)w, sb, se, 0));
      if (q < Q - 1 && acc >= total * (q + 1) / Q) run_begin[++q] = (uint32_t)runs.size();
    }
    while (q < Q) run_begin[++q] = (uint32_t)runs.size();
  };
  build_runs(false, H.runs, H.run_begin);
  build_runs(true, H.own_runs, H.own_run_begin);
  return 0;
}

}  // namespace gcnb

// =====================================================================================================================
// device
// =====================================================================================================================
namespace gcnb {

struct StagedDev {
  int dim = 0, window_rows = 0, n_cta = 0;
  int64_t n_rows = 0, n_cols = 0, n_blocks = 0, n_slots = 0, n_runs = 0, staged_nnz = 0, rem_nnz = 0, n_segs = 0;
  uint4 *d_bundles = nullptr, *d_runs = nullptr, *d_own_runs = nullptr;
  uint32_t *d_own_run_begin = nullptr;
  int64_t n_own_runs = 0, own_col0 = 0;
  bool own_pending = false;  // the own-window runs of the next product have already been launched (gcnb_spmm_stage_own_f32)
  uint16_t *d_lens = nullptr;
  uint32_t *d_run_begin = nullptr, *d_row_slot = nullptr, *d_lane_slot = nullptr, *d_counters = nullptr;
  uint16_t *d_pidx = nullptr;
  uint32_t *d_pperm = nullptr;
  float *d_pval = nullptr;
  float *d_partial = nullptr;
  uint32_t *d_r_indptr = nullptr, *d_r_indices = nullptr, *d_r_perm = nullptr;
  float *d_r_val = nullptr;
  gcnb_spmm_plan *rem = nullptr;      // generic plan over the remainder CSR
  cudaStream_t aux = nullptr;         // the remainder product runs here, concurrently with the staged kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  const float *values_src = nullptr;  // the value array the packed copies were gathered from
  // column-slab mode (operands wider than the staged width): one 16-column slab of B packed contiguously (so that the
  // window copy stays one bulk copy and the remainder's random gathers hit a 15 MB L2-resident matrix instead of a
  // strided one in HBM), and the remainder product of the slab before it is merged into the strided C
  float *d_pack = nullptr, *d_remout = nullptr;
};

void stage_destroy(StagedDev *s) {
  if (!s) return;
  cudaFree(s->d_bundles); cudaFree(s->d_runs); cudaFree(s->d_lens); cudaFree(s->d_run_begin); cudaFree(s->d_row_slot);
  cudaFree(s->d_lane_slot); cudaFree(s->d_counters); cudaFree(s->d_pidx); cudaFree(s->d_pperm); cudaFree(s->d_pval); cudaFree(s->d_partial);
  cudaFree(s->d_r_indptr); cudaFree(s->d_r_indices); cudaFree(s->d_r_perm); cudaFree(s->d_r_val);
  cudaFree(s->d_pack); cudaFree(s->d_remout); cudaFree(s->d_own_runs); cudaFree(s->d_own_run_begin);
  if (s->rem) gcnb_spmm_plan_destroy(s->rem);
  if (s->aux) cudaStreamDestroy(s->aux);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  delete s;
}

}  // namespace gcnb

namespace {

constexpr int kStageMaxWindow = 3072;  // rows of 64 bytes: 192 KB of the SM's 227 KB

__device__ __forceinline__ uint2 ld_stream_u2(const uint2 *p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void fma4(float4 &acc, float a, const float4 x) {
  acc.x = fmaf(a, x.x, acc.x);
  acc.y = fmaf(a, x.y, acc.y);
  acc.z = fmaf(a, x.z, acc.z);
  acc.w = fmaf(a, x.w, acc.w);
}

// one entry of this lane's segment: the 64-byte row `lcol` of the window, column chunk c read at c ^ (lane & 3)
__device__ __forceinline__ void stage_entry(float4 (&acc)[4], const char *__restrict__ win, const uint32_t (&xo)[4],
                                            uint32_t lcol, float a) {
  const uint32_t base = lcol * 64u;
  float4 x[4];
#pragma unroll
  for (int c = 0; c < 4; c++) x[c] = *reinterpret_cast<const float4 *>(win + (base ^ xo[c]));
#pragma unroll
  for (int c = 0; c < 4; c++) fma4(acc[c], a, x[c]);
}

// dim == 16.  Lane l owns segment l of the bundle; acc[c] holds column chunk c ^ (l & 3).
// Runs are claimed CTA-wide with one atomic ticket per run: first from the CTA's own queue (contiguous runs, so the
// window usually stays), then from whichever queue still has runs (work stealing evens out the SMs).
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
spmm_staged16_kernel(const uint4 *__restrict__ runs, const uint32_t *__restrict__ run_begin,
                     uint32_t *__restrict__ counters, int n_queues, const uint4 *__restrict__ bundles,
                     const uint16_t *__restrict__ lens, const uint32_t *__restrict__ lane_slot,
                     const uint2 *__restrict__ pidx, const float4 *__restrict__ pval, const float *__restrict__ B,
                     float *__restrict__ partial, int window_rows, int64_t n_cols) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const char *win = reinterpret_cast<const char *>(smem_raw);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)window_rows * 64);
  volatile uint32_t *s_claim = reinterpret_cast<volatile uint32_t *>(bar + 1);  // [0] queue, [1] ticket
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t xo[4];
#pragma unroll
  for (int c = 0; c < 4; c++) xo[c] = (uint32_t)((c ^ (lane & 3)) << 4);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  uint32_t phase = 0, cur_win = 0xffffffffu;
  int q = (int)(blockIdx.x % (unsigned)n_queues);
  bool own = true;
  for (;;) {
    __syncthreads();  // every warp is done with the previous run (window and s_claim are free again)
    if (warp == 0) {
      uint32_t cq = (uint32_t)q, ticket = 0xffffffffu;
      if (own) {
        if (lane == 0) ticket = atomicAdd(counters + q, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket >= __ldg(run_begin + q + 1) - __ldg(run_begin + q)) ticket = 0xffffffffu;
      }
      if (ticket == 0xffffffffu) {  // own queue is drained: take a run from the first queue that still has some
        for (int base = 1; base < n_queues && ticket == 0xffffffffu; base += 32) {
          const int off = base + lane;
          bool has = false;
          int qq = 0;
          if (off < n_queues) {
            qq = (q + off) % n_queues;
            has = *((volatile uint32_t *)(counters + qq)) < __ldg(run_begin + qq + 1) - __ldg(run_begin + qq);
          }
          uint32_t mask = __ballot_sync(0xffffffffu, has);
          while (mask && ticket == 0xffffffffu) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int vq = __shfl_sync(0xffffffffu, qq, src);
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(counters + vq, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t < __ldg(run_begin + vq + 1) - __ldg(run_begin + vq)) {
              ticket = t;
              cq = (uint32_t)vq;
            }
          }
        }
      }
      if (lane == 0) {
        s_claim[0] = cq;
        s_claim[1] = ticket;
      }
    }
    __syncthreads();
    const uint32_t cq = s_claim[0], ticket = s_claim[1];
    if (ticket == 0xffffffffu) break;
    own = own && cq == (uint32_t)q;
    const uint4 run = __ldg(runs + __ldg(run_begin + cq) + ticket);
    if (run.x != cur_win) {
      cur_win = run.x;
      if (threadIdx.x == 0) {
        const int64_t row0 = (int64_t)run.x * window_rows;
        const uint32_t rows = (uint32_t)min((int64_t)window_rows, n_cols - row0);
        const uint32_t bytes = rows * 64u;
        fence_proxy_async();
        mbar_expect_tx(bar, bytes);
        const char *src = reinterpret_cast<const char *>(B + row0 * 16);
        for (uint32_t off = 0; off < bytes; off += 32768u)
          bulk_g2s(smem_raw + off, src + off, min(32768u, bytes - off), bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
    }

    uint32_t b = run.y + warp;
    if (b >= run.z) continue;
    uint4 bd = __ldg(bundles + b);
    uint32_t mylen = __ldg(lens + (size_t)b * 32 + lane);
    uint32_t myslot = __ldg(lane_slot + (size_t)b * 32 + lane);
    const uint2 *ip = pidx + (size_t)bd.x * 32 + lane;
    const float4 *vp = pval + (size_t)bd.x * 32 + lane;
    uint2 I = ld_stream_u2(ip);
    float4 V = ld_stream_f4(vp);
    for (;;) {
      const uint32_t bnext = b + NT / 32;
      const bool more = bnext < run.z;
      uint4 bdn = make_uint4(0, 0, 0, 0);
      uint32_t lenn = 0, slotn = 0;
      if (more) {
        bdn = __ldg(bundles + bnext);
        lenn = __ldg(lens + (size_t)bnext * 32 + lane);
        slotn = __ldg(lane_slot + (size_t)bnext * 32 + lane);
      }
      const uint32_t L = bd.y, nblk = (L + 3) >> 2, full_blk = bd.z >> 2;  // blocks in which every lane is active
      float4 acc[4];
#pragma unroll
      for (int c = 0; c < 4; c++) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t kb = 0; kb < nblk; kb++) {
        const uint2 cI = I;
        const float4 cV = V;
        if (kb + 1 < nblk) {
          ip += 32;
          vp += 32;
          I = ld_stream_u2(ip);
          V = ld_stream_f4(vp);
        } else if (more) {  // first block of this warp's next bundle
          ip = pidx + (size_t)bdn.x * 32 + lane;
          vp = pval + (size_t)bdn.x * 32 + lane;
          I = ld_stream_u2(ip);
          V = ld_stream_f4(vp);
        }
        const uint32_t lc[4] = {cI.x & 0xffffu, cI.x >> 16, cI.y & 0xffffu, cI.y >> 16};
        const float a[4] = {cV.x, cV.y, cV.z, cV.w};
        if (kb < full_blk) {
#pragma unroll
          for (int u = 0; u < 4; u++) stage_entry(acc, win, xo, lc[u], a[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (kb * 4 + u < mylen) stage_entry(acc, win, xo, lc[u], a[u]);
        }
      }
      if (mylen) {
        float *dst = partial + (size_t)myslot * 16;
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<float4 *>(dst + (xo[c] >> 2)) = acc[c];
      }
      if (!more) break;
      bd = bdn;
      mylen = lenn;
      myslot = slotn;
      b = bnext;
    }
  }
}

// C[row] = R[row] + partial[slot] over the row's slots [row_slot[r], row_slot[r+1]) in ascending order, R = the remainder
// product (in place in C when R == nullptr, else a contiguous [n_rows x 16] buffer and C is a 16-column slab with row
// stride ldc).  Thread = (row, 16-byte chunk): consecutive threads read consecutive 16-byte pieces; the loads of four
// slots are issued together, the additions stay in slot order.  VW = widest aligned access to C (4, 2 or 1 floats).
template <int VW>
__global__ void __launch_bounds__(256)
stage_add16_kernel(const uint32_t *__restrict__ row_slot, const float *__restrict__ partial, const float *__restrict__ R,
                   float *__restrict__ C, int64_t n_rows, int64_t ldc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_rows * 4) return;
  const int64_t r = i >> 2;
  const int c = (int)(i & 3);
  const uint32_t s0 = __ldg(row_slot + r), s1 = __ldg(row_slot + r + 1);
  float *dst = C + r * ldc + c * 4;
  float4 acc;
  if (R) {
    acc = __ldg(reinterpret_cast<const float4 *>(R + r * 16 + c * 4));
  } else {
    if (s0 == s1) return;
    acc = *reinterpret_cast<const float4 *>(dst);  // in-place mode is only used with ldc == 16 (VW == 4)
  }
  const float4 *p = reinterpret_cast<const float4 *>(partial + (size_t)s0 * 16 + c * 4);
  for (uint32_t sl = s0; sl < s1; sl += 4, p += 16) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sl + u < s1) v[u] = __ldg(p + u * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      acc.x += v[u].x;
      acc.y += v[u].y;
      acc.z += v[u].z;
      acc.w += v[u].w;
    }
  }
  if (VW == 4) {
    *reinterpret_cast<float4 *>(dst) = acc;
  } else if (VW == 2) {
    *reinterpret_cast<float2 *>(dst) = make_float2(acc.x, acc.y);
    *reinterpret_cast<float2 *>(dst + 2) = make_float2(acc.z, acc.w);
  } else {
    dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
  }
}

// one 16-column slab of a wider row-major matrix, packed contiguously: out[r][0..16) = B[r * ldb + 0..16)
template <int VW>
__global__ void __launch_bounds__(256)
stage_pack16_kernel(const float *__restrict__ B, int64_t ldb, float *__restrict__ out, int64_t n_rows) {
  constexpr int PER = 16 / VW;  // pieces per row
  const int64_t total = n_rows * PER;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / PER;
    const int c = (int)(i - r * PER) * VW;
    const float *src = B + r * ldb + c;
    if (VW == 4) *reinterpret_cast<float4 *>(out + r * 16 + c) = __ldg(reinterpret_cast<const float4 *>(src));
    else if (VW == 2) *reinterpret_cast<float2 *>(out + r * 16 + c) = __ldg(reinterpret_cast<const float2 *>(src));
    else out[r * 16 + c] = __ldg(src);
  }
}

__global__ void stage_gather_values_kernel(const uint32_t *__restrict__ perm, const float *__restrict__ values,
                                           float *__restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t e = __ldg(perm + i);
    out[i] = e == kStagePad ? 0.f : __ldg(values + e);
  }
}

template <class T>
int upload_vec(T **dst, const HostArray<T> &v, cudaStream_t stream) {
  const size_t bytes = std::max<size_t>(16, v.size() * sizeof(T));
  GCNB_CHECK(cudaMalloc((void **)dst, bytes));
  if (!v.empty()) GCNB_CHECK(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

template <class T>
int upload_vec(T **dst, const std::vector<T> &v, cudaStream_t stream) {
  const size_t bytes = std::max<size_t>(16, v.size() * sizeof(T));
  GCNB_CHECK(cudaMalloc((void **)dst, bytes));
  if (!v.empty()) GCNB_CHECK(cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
  return 0;
}

int gather_values(StagedDev *s, const float *d_values, cudaStream_t stream) {
  const int blocks = std::max(1, device_info().sm_count) * 8;
  const int64_t np = s->n_blocks * kStageLanes * kStageBlock;
  if (np) stage_gather_values_kernel<<<blocks, 256, 0, stream>>>(s->d_pperm, d_values, s->d_pval, np);
  GCNB_LAUNCH_CHECK();
  if (s->rem_nnz) stage_gather_values_kernel<<<blocks, 256, 0, stream>>>(s->d_r_perm, d_values, s->d_r_val, s->rem_nnz);
  GCNB_LAUNCH_CHECK();
  s->values_src = d_values;
  return 0;
}

}  // namespace

namespace {

// widest vector access (floats) valid for every row of a slab that starts at `p` with row stride `ld`
int slab_vec_width(const void *p, int64_t ld) {
  if ((uintptr_t)p % 16 == 0 && ld % 4 == 0) return 4;
  if ((uintptr_t)p % 8 == 0 && ld % 2 == 0) return 2;
  return 1;
}

// one launch of the persistent staged kernel over a run list (`counters`: n_cta tickets, zeroed here)
int launch_staged_runs(StagedDev *s, const uint4 *d_runs, const uint32_t *d_run_begin, uint32_t *counters, const float *d_B,
                       cudaStream_t stream) {
  const size_t smem = (size_t)s->window_rows * s->dim * sizeof(float) + 32;
  static const int nt = [] {
    const char *e = getenv("GCNB_STAGE_THREADS");  // tuning probe
    return e ? atoi(e) : 564;
  }();
  GCNB_CHECK(cudaMemsetAsync(counters, 0, (size_t)s->n_cta * 4, stream));
#define GCNB_STAGE_LAUNCH(NT, MINB)                                                                                    \
  spmm_staged16_kernel<NT, MINB><<<s->n_cta, NT, smem, stream>>>(                                                      \
      d_runs, d_run_begin, counters, s->n_cta, s->d_bundles, s->d_lens, s->d_lane_slot,                                \
      reinterpret_cast<const uint2 *>(s->d_pidx), reinterpret_cast<const float4 *>(s->d_pval), d_B, s->d_partial,     \
      s->window_rows, s->n_cols)
  if (nt == 1024) GCNB_STAGE_LAUNCH(1024, 1);
  else if (nt == 512) GCNB_STAGE_LAUNCH(512, 1);
  else if (nt == 564) GCNB_STAGE_LAUNCH(512, 2);  // 512 threads capped at 64 registers
  else if (nt == 384) GCNB_STAGE_LAUNCH(384, 1);
  else GCNB_STAGE_LAUNCH(256, 1);
#undef GCNB_STAGE_LAUNCH
  GCNB_LAUNCH_CHECK();
  return 0;
}

// C[:, 0..16) (row stride ldc) = A * B[:, 0..16) (row stride ldb) through the staged representation
int staged_slab16(StagedDev *s, const float *d_B, int64_t ldb, float *d_C, int64_t ldc, cudaStream_t stream) {
  const int sms = std::max(1, device_info().sm_count);
  if (ldb != 16 || (uintptr_t)d_B % 16 != 0) {
    if (!s->d_pack) GCNB_CHECK(cudaMalloc((void **)&s->d_pack, std::max<size_t>(16, (size_t)s->n_cols * 16 * sizeof(float))));
    const int vw = slab_vec_width(d_B, ldb);
    const int blocks = (int)std::min<int64_t>((s->n_cols * (16 / vw) + 255) / 256, (int64_t)sms * 16);
    if (vw == 4) stage_pack16_kernel<4><<<blocks, 256, 0, stream>>>(d_B, ldb, s->d_pack, s->n_cols);
    else if (vw == 2) stage_pack16_kernel<2><<<blocks, 256, 0, stream>>>(d_B, ldb, s->d_pack, s->n_cols);
    else stage_pack16_kernel<1><<<blocks, 256, 0, stream>>>(d_B, ldb, s->d_pack, s->n_cols);
    GCNB_LAUNCH_CHECK();
    d_B = s->d_pack;
  }
  const bool in_place = ldc == 16 && (uintptr_t)d_C % 16 == 0;
  float *rem_out = d_C;
  if (!in_place) {
    if (!s->d_remout) GCNB_CHECK(cudaMalloc((void **)&s->d_remout, std::max<size_t>(16, (size_t)s->n_rows * 16 * sizeof(float))));
    rem_out = s->d_remout;
  }
  static const int concurrent = [] {
    const char *e = getenv("GCNB_STAGE_CONCURRENT");  // tuning probe
    return e ? atoi(e) : 1;
  }();
  // The remainder product (latency-bound L2 gathers) runs on the plan's second stream while the staged kernel
  // (LSU-bound shared-memory gathers, one persistent CTA per SM) runs here: the two co-reside on every SM.
  cudaStream_t rs = stream;
  if (concurrent) {
    GCNB_CHECK(cudaEventRecord(s->ev_fork, stream));
    GCNB_CHECK(cudaStreamWaitEvent(s->aux, s->ev_fork, 0));
    rs = s->aux;
  }
  const bool own_done = s->own_pending;  // launched ahead of the exchange from the rank's own slab
  s->own_pending = false;
  if (s->n_runs > 0) {
    const int rc = launch_staged_runs(s, s->d_runs, s->d_run_begin, s->d_counters, d_B, stream);
    if (rc) return rc;
  }
  if (!own_done && s->n_own_runs > 0) {
    const int rc = launch_staged_runs(s, s->d_own_runs, s->d_own_run_begin, s->d_counters + s->n_cta, d_B, stream);
    if (rc) return rc;
  }
  int rc = spmm_generic_launch(s->rem, s->d_r_val, nullptr, d_B, 16, rem_out, 16, 16, rs);
  if (rc) return rc;
  if (concurrent) {
    GCNB_CHECK(cudaEventRecord(s->ev_join, s->aux));
    GCNB_CHECK(cudaStreamWaitEvent(stream, s->ev_join, 0));
  }
  const int64_t total = s->n_rows * 4;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (in_place) {
    stage_add16_kernel<4><<<blocks, 256, 0, stream>>>(s->d_row_slot, s->d_partial, nullptr, d_C, s->n_rows, 16);
  } else {
    const int vw = slab_vec_width(d_C, ldc);
    if (vw == 4) stage_add16_kernel<4><<<blocks, 256, 0, stream>>>(s->d_row_slot, s->d_partial, rem_out, d_C, s->n_rows, ldc);
    else if (vw == 2) stage_add16_kernel<2><<<blocks, 256, 0, stream>>>(s->d_row_slot, s->d_partial, rem_out, d_C, s->n_rows, ldc);
    else stage_add16_kernel<1><<<blocks, 256, 0, stream>>>(s->d_row_slot, s->d_partial, rem_out, d_C, s->n_rows, ldc);
  }
  GCNB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Smallest operand width (other than the staged width itself) that is cut into 16-column slabs; narrower products stay
// on the generic kernel, which reads the index once for all columns.
static int stage_slab_min_dim() {
  static const int v = [] {
    const char *e = getenv("GCNB_STAGE_SLAB_MIN");  // tuning probe
    return e ? atoi(e) : 64;
  }();
  return v;
}

int gcnb_stage_try_spmm(gcnb_spmm_plan *p, const float *d_values, const uint32_t *d_perm, const float *d_B, int ldb,
                        float *d_C, int ldc, int dim, cudaStream_t stream, int *handled) {
  *handled = 0;
  StagedDev *s = p->staged;
  if (!s || d_perm || s->dim != 16 || d_values != s->values_src) return 0;
  if (dim != 16 && dim < stage_slab_min_dim()) return 0;
  // slabs of 16 columns; the last one is shifted left to end at `dim` (its overlap with the previous slab is computed
  // twice with identical results) so that every slab takes the staged path
  for (int c0 = 0; c0 < dim; c0 += 16) {
    const int c = std::min(c0, dim - 16);
    const int rc = staged_slab16(s, d_B + c, ldb, d_C + c, ldc, stream);
    if (rc) return rc;
  }
  *handled = 1;
  return 0;
}

extern "C" {

int gcnb_spmm_plan_stage(gcnb_spmm_plan *p, const uint32_t *h_indptr, const uint32_t *h_indices, const float *d_values,
                         int dim, gcnb_stream_t stream_) {
  return gcnb_spmm_plan_stage_ex(p, h_indptr, h_indices, d_values, dim, 0, 0, 0, 0, stream_);
}

// Builds the staged representation of `p` WITHOUT attaching it (*out = nullptr when staging is not worthwhile): reads only
// fields of the plan that never change after gcnb_spmm_plan_create / gcnb_spmm_plan_set_own_cols, so it may run on a
// helper thread (own stream) while the plan is in use on the generic kernel.
static int stage_make(const gcnb_spmm_plan *p, const uint32_t *h_indptr, const uint32_t *h_indices, const float *d_values,
                      int dim, int window_rows, int min_seg, int seg_cap, int64_t min_window_nnz, gcnb_stream_t stream_,
                      StagedDev **out) {
  *out = nullptr;
  cudaStream_t stream = as_stream(stream_);
  std::vector<uint32_t> indptr_copy, indices_copy;
  if (!h_indptr) {
    indptr_copy.resize((size_t)p->n_rows + 1);
    GCNB_CHECK(cudaMemcpyAsync(indptr_copy.data(), p->d_indptr, indptr_copy.size() * 4, cudaMemcpyDeviceToHost, stream));
    h_indptr = indptr_copy.data();
  }
  if (!h_indices) {
    indices_copy.resize((size_t)p->nnz);
    GCNB_CHECK(cudaMemcpyAsync(indices_copy.data(), p->d_indices, indices_copy.size() * 4, cudaMemcpyDeviceToHost, stream));
    h_indices = indices_copy.data();
  }
  GCNB_CHECK(cudaStreamSynchronize(stream));
  StageParams P;
  P.dim = dim;
  P.n_cta = std::max(1, device_info().sm_count);
  if (window_rows > 0) P.window_rows = std::min(window_rows, kStageMaxWindow);
  if (min_seg > 0) P.min_seg = min_seg;
  if (seg_cap > 0) P.seg_cap = seg_cap;
  P.min_window_nnz = min_window_nnz;
  if (window_rows > 0) P.min_avg_seg = 1;  // explicit knobs (tests): stage whatever qualifies
  P.own_col0 = p->own_col0;
  P.own_col1 = p->own_col1;
  StagedHost H;
  int rc = stage_build_host(h_indptr, h_indices, p->n_rows, p->n_cols, P, H);
  if (rc == GCNB_E_UNSUPPORTED) return 0;  // too many windows / too large: stay on the generic kernel
  if (rc) return rc;
  // little to gain (few stageable entries, or segments so short that per-segment work dominates): generic kernel
  if (H.staged_nnz * 4 < H.nnz || H.staged_nnz < (int64_t)P.min_avg_seg * H.n_segs) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    const int max_smem = kStageMaxWindow * 16 * (int)sizeof(float) + 32;
    GCNB_CHECK(cudaFuncSetAttribute(spmm_staged16_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    GCNB_CHECK(cudaFuncSetAttribute(spmm_staged16_kernel<384, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    GCNB_CHECK(cudaFuncSetAttribute(spmm_staged16_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    GCNB_CHECK(cudaFuncSetAttribute(spmm_staged16_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    GCNB_CHECK(cudaFuncSetAttribute(spmm_staged16_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  auto *s = new StagedDev();
  s->dim = dim;
  s->window_rows = H.window_rows;
  s->n_cta = H.n_cta;
  s->n_rows = H.n_rows;
  s->n_cols = H.n_cols;
  s->n_blocks = H.n_blocks;
  s->n_slots = H.n_slots;
  s->n_runs = (int64_t)H.runs.size();
  s->n_own_runs = (int64_t)H.own_runs.size();
  s->own_col0 = P.own_col0;
  s->n_segs = H.n_segs;
  s->staged_nnz = H.staged_nnz;
  s->rem_nnz = H.nnz - H.staged_nnz;
  auto fail = [&](int code) {
    stage_destroy(s);
    return code;
  };
  if ((rc = upload_vec(&s->d_bundles, H.bundles, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_runs, H.runs, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_lens, H.lens, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_run_begin, H.run_begin, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_own_runs, H.own_runs, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_own_run_begin, H.own_run_begin, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_row_slot, H.row_slot, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_lane_slot, H.lane_slot, stream))) return fail(rc);
  if ((rc = (int)cudaMalloc((void **)&s->d_counters, (size_t)H.n_cta * 2 * 4))) return fail(rc);
  if ((rc = (int)cudaStreamCreateWithFlags(&s->aux, cudaStreamNonBlocking))) return fail(rc);
  if ((rc = (int)cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming))) return fail(rc);
  if ((rc = (int)cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming))) return fail(rc);
  if ((rc = upload_vec(&s->d_pidx, H.pidx, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_pperm, H.pperm, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_r_indptr, H.r_indptr, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_r_indices, H.r_indices, stream))) return fail(rc);
  if ((rc = upload_vec(&s->d_r_perm, H.r_perm, stream))) return fail(rc);
  const size_t np = (size_t)H.n_blocks * kStageLanes * kStageBlock;
  if ((rc = (int)cudaMalloc((void **)&s->d_pval, std::max<size_t>(16, np * 4)))) return fail(rc);
  if ((rc = (int)cudaMalloc((void **)&s->d_r_val, std::max<size_t>(16, (size_t)s->rem_nnz * 4)))) return fail(rc);
  if ((rc = (int)cudaMalloc((void **)&s->d_partial, std::max<size_t>(16, (size_t)H.n_slots * dim * 4)))) return fail(rc);
  if ((rc = (int)cudaStreamSynchronize(stream))) return fail(rc);
  if ((rc = gcnb_spmm_plan_create(s->d_r_indptr, s->d_r_indices, p->n_rows, p->n_cols, p->seg_nnz, stream_, &s->rem)))
    return fail(rc);
  if ((rc = gather_values(s, d_values, stream))) return fail(rc);
  if ((rc = (int)cudaStreamSynchronize(stream))) return fail(rc);
  *out = s;
  return 0;
}

int gcnb_spmm_plan_stage_ex(gcnb_spmm_plan *p, const uint32_t *h_indptr, const uint32_t *h_indices,
                            const float *d_values, int dim, int window_rows, int min_seg, int seg_cap,
                            int64_t min_window_nnz, gcnb_stream_t stream_) {
  if (!p || !d_values) return GCNB_E_BADARG;
  if (dim != 16 || p->n_rows == 0 || p->nnz == 0) return 0;  // only the dim-16 kernel exists; not an error
  if (p->staged && p->staged->dim == dim) return gather_values(p->staged, d_values, as_stream(stream_));  // re-gather
  StagedDev *s = nullptr;
  const int rc = stage_make(p, h_indptr, h_indices, d_values, dim, window_rows, min_seg, seg_cap, min_window_nnz, stream_, &s);
  if (rc) return rc;
  if (s) p->staged = s;
  return 0;
}

// ---- background staging ----------------------------------------------------------------------------------------------
// The build (0.2 s of host threads at Reddit scale) and the upload of the packed arrays (0.7 GB) run on a helper thread
// with its own stream while the caller already trains on the generic kernel; _finish joins and attaches.  WHEN the
// caller finishes is the caller's decision (the engine does it at a fixed epoch), so results do not depend on timing.
struct gcnb_stage_job {
  std::thread th;
  int rc = 0;
  int device = 0;
  StagedDev *s = nullptr;
  std::atomic<int> done{0};
};

int gcnb_spmm_plan_stage_async_begin(gcnb_spmm_plan *p, const float *d_values, int dim, gcnb_stage_job **out) {
  if (!p || !d_values || !out) return GCNB_E_BADARG;
  *out = nullptr;
  if (dim != 16 || p->n_rows == 0 || p->nnz == 0 || p->staged) return 0;  // nothing to do: no job
  int device = 0;
  GCNB_CHECK(cudaGetDevice(&device));
  auto *job = new gcnb_stage_job();
  job->device = device;
  const gcnb_spmm_plan *cp = p;
  job->th = std::thread([job, cp, d_values, dim] {
    cudaStream_t st = nullptr;
    int rc = (int)cudaSetDevice(job->device);
    if (!rc) rc = (int)cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    // host copies of the index are read back from the device here (the caller's host arrays may be gone by now)
    if (!rc) rc = stage_make(cp, nullptr, nullptr, d_values, dim, 0, 0, 0, 0, (gcnb_stream_t)st, &job->s);
    if (st) cudaStreamDestroy(st);
    job->rc = rc;
    job->done.store(1, std::memory_order_release);
  });
  *out = job;
  return 0;
}

int gcnb_spmm_plan_stage_async_done(const gcnb_stage_job *job) { return !job || job->done.load(std::memory_order_acquire); }

int gcnb_spmm_plan_stage_async_finish(gcnb_spmm_plan *p, gcnb_stage_job *job) {
  if (!job) return 0;
  if (job->th.joinable()) job->th.join();
  const int rc = job->rc;
  if (!rc && job->s) {
    if (p && !p->staged) p->staged = job->s;
    else stage_destroy(job->s);
  }
  delete job;
  return rc;
}

int gcnb_spmm_plan_set_own_cols(gcnb_spmm_plan *p, int64_t col0, int64_t col1) {
  if (!p || col0 < 0 || col1 < col0 || col1 > p->n_cols) return GCNB_E_BADARG;
  p->own_col0 = col0;
  p->own_col1 = col1;
  return 0;
}

int gcnb_spmm_stage_own_f32(gcnb_spmm_plan *p, const float *d_values, const float *d_B_own, int dim, gcnb_stream_t stream_,
                            int *launched) {
  if (!p || !d_values || !d_B_own || !launched) return GCNB_E_BADARG;
  *launched = 0;
  StagedDev *s = p->staged;
  if (!s || dim != 16 || s->dim != 16 || d_values != s->values_src || s->n_own_runs == 0 || (uintptr_t)d_B_own % 16 != 0)
    return 0;
  // the own-window runs only touch rows [own_col0, own_col1) of B: address them through the slab
  const float *base = d_B_own - s->own_col0 * 16;
  const int rc = launch_staged_runs(s, s->d_own_runs, s->d_own_run_begin, s->d_counters + s->n_cta, base, as_stream(stream_));
  if (rc) return rc;
  s->own_pending = true;
  *launched = 1;
  return 0;
}

int gcnb_spmm_plan_stage_slabs(const gcnb_spmm_plan *p, int dim) {
  const StagedDev *s = p ? p->staged : nullptr;
  if (!s || s->dim != 16 || (dim != 16 && dim < stage_slab_min_dim())) return 0;
  return (dim + 15) / 16;
}

int gcnb_spmm_plan_stage_info(const gcnb_spmm_plan *p, int64_t out[8]) {
  if (!p || !out) return GCNB_E_BADARG;
  for (int i = 0; i < 8; i++) out[i] = 0;
  const StagedDev *s = p->staged;
  if (!s) return 0;
  out[0] = 1; out[1] = s->window_rows; out[2] = s->staged_nnz; out[3] = s->rem_nnz;
  out[4] = s->n_segs; out[5] = s->n_runs + s->n_own_runs; out[6] = s->n_blocks; out[7] = s->n_slots;
  return 0;
}

// ---- host-only access to the builder (CPU unit tests: no CUDA call is made) --------------------------------------------
struct gcnb_stage_host {
  StagedHost H;
};

int gcnb_stage_host_build(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols, int dim,
                          int window_rows, int min_seg, int seg_cap, int64_t min_window_nnz, int n_cta, int n_threads,
                          gcnb_stage_host **out) {
  return gcnb_stage_host_build_own(h_indptr, h_indices, n_rows, n_cols, dim, window_rows, min_seg, seg_cap, min_window_nnz,
                                   n_cta, n_threads, 0, 0, out);
}

int gcnb_stage_host_build_own(const uint32_t *h_indptr, const uint32_t *h_indices, int64_t n_rows, int64_t n_cols, int dim,
                              int window_rows, int min_seg, int seg_cap, int64_t min_window_nnz, int n_cta, int n_threads,
                              int64_t own_col0, int64_t own_col1, gcnb_stage_host **out) {
  if (!out) return GCNB_E_BADARG;
  StageParams P;
  P.dim = dim;
  P.window_rows = window_rows;
  if (min_seg > 0) P.min_seg = min_seg;
  if (seg_cap > 0) P.seg_cap = seg_cap;
  P.min_window_nnz = min_window_nnz;
  if (n_cta > 0) P.n_cta = n_cta;
  P.n_threads = n_threads;
  P.own_col0 = own_col0;
  P.own_col1 = own_col1;
  auto *h = new gcnb_stage_host();
  const int rc = stage_build_host(h_indptr, h_indices, n_rows, n_cols, P, h->H);
  if (rc) {
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int gcnb_stage_host_sizes(const gcnb_stage_host *h, int64_t out[13]) {
  if (!h || !out) return GCNB_E_BADARG;
  const StagedHost &H = h->H;
  out[0] = H.window_rows; out[1] = H.n_win; out[2] = H.n_cta; out[3] = H.staged_nnz;
  out[4] = (int64_t)H.bundles.size(); out[5] = (int64_t)H.runs.size(); out[6] = H.n_blocks; out[7] = H.n_slots;
  out[8] = (int64_t)H.r_indices.size(); out[9] = H.n_rows; out[10] = H.nnz; out[11] = H.n_segs;
  out[12] = (int64_t)H.own_runs.size();
  return 0;
}

// which: 0 bundles (uint32 x4), 1 runs (uint32 x4), 2 run_begin, 3 pidx (uint16), 4 pperm, 5 row_slot, 6 r_indptr,
// 7 r_indices, 8 r_perm, 9 lens (uint16), 10 lane_slot.  Copies min(bytes, size) bytes.
int gcnb_stage_host_copy(const gcnb_stage_host *h, int which, void *dst, int64_t bytes) {
  if (!h || !dst) return GCNB_E_BADARG;
  const StagedHost &H = h->H;
  const void *src = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: src = H.bundles.data(); n = H.bundles.size() * sizeof(uint4); break;
    case 1: src = H.runs.data(); n = H.runs.size() * sizeof(uint4); break;
    case 2: src = H.run_begin.data(); n = H.run_begin.size() * 4; break;
    case 3: src = H.pidx.data(); n = H.pidx.size() * 2; break;
    case 4: src = H.pperm.data(); n = H.pperm.size() * 4; break;
    case 5: src = H.row_slot.data(); n = H.row_slot.size() * 4; break;
    case 6: src = H.r_indptr.data(); n = H.r_indptr.size() * 4; break;
    case 7: src = H.r_indices.data(); n = H.r_indices.size() * 4; break;
    case 8: src = H.r_perm.data(); n = H.r_perm.size() * 4; break;
    case 9: src = H.lens.data(); n = H.lens.size() * 2; break;
    case 10: src = H.lane_slot.data(); n = H.lane_slot.size() * 4; break;
    case 11: src = H.own_runs.data(); n = H.own_runs.size() * sizeof(uint4); break;
    case 12: src = H.own_run_begin.data(); n = H.own_run_begin.size() * 4; break;
    default: return GCNB_E_BADARG;
  }
  memcpy(dst, src, std::min<size_t>(n, (size_t)bytes));
  return 0;
}

int gcnb_stage_host_destroy(gcnb_stage_host *h) {
  delete h;
  return 0;
}

}  // extern "C"
