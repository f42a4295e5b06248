// reorder.cpp -- locality reordering of the graph ahead of the engine (nothing like it exists in the
// reference, whose script/ordering.py sorts result lines).  The window-staged GraphSum (csrc/spmm_stage.cu) serves the
// entries of a row that fall into one 3072-row column window from shared memory; that only pays when the nodes of a
// community sit next to each other.  Files rarely come that way, so this module finds communities by label propagation
// and renumbers the nodes community by community.  The renumbered dataset is an ordinary dataset: everything downstream
// keeps the reference's semantics on it, and gcnb_unpermute_rows maps per-node outputs back.
//
// Label propagation, synchronous and deterministic (independent of the number of threads): every sweep, node i adopts
// the label that occurs most often among its neighbours' labels of the previous sweep (its own row entry i included);
// ties are broken by a hash of (label, sweep) -- breaking them by the smallest label would let a handful of ids flood
// the graph through its random long-range edges.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

#include "../../../include/gcnb_engine.h"

namespace {

inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

unsigned n_threads(int64_t work) {
  if (work < (1 << 16)) return 1;
  return std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
}

template <class F>
void parallel_rows(int64_t n, const uint32_t *indptr, F fn) {  // contiguous row ranges of equal nnz
  const int64_t nnz = n ? indptr[n] : 0;
  const unsigned nt = n_threads(nnz);
  std::vector<int64_t> cut(nt + 1, n);
  cut[0] = 0;
  for (unsigned t = 1; t < nt; t++) {
    const uint32_t target = (uint32_t)((uint64_t)nnz * t / nt);
    cut[t] = std::max<int64_t>(cut[t - 1], std::lower_bound(indptr, indptr + n + 1, target) - indptr);
    cut[t] = std::min<int64_t>(cut[t], n);
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back([=] { fn(cut[t], cut[t + 1]); });
  for (auto &x : th) x.join();
}

// label[i] = community of node i (a node id of that community)
void propagate_labels(int64_t n, const uint32_t *indptr, const uint32_t *indices, int max_sweeps, uint64_t seed,
                      std::vector<uint32_t> &label) {
  if (max_sweeps <= 0) max_sweeps = 8;
  std::vector<uint32_t> next((size_t)n);
  label.resize((size_t)n);
  std::iota(label.begin(), label.end(), 0u);
  for (int sweep = 0; sweep < max_sweeps; sweep++) {
    std::atomic<int64_t> changed_total(0);
    const uint64_t salt = mix64(seed ^ ((uint64_t)sweep << 32));
    parallel_rows(n, indptr, [&](int64_t a, int64_t b) {
      std::vector<uint32_t> buf;
      int64_t changed = 0;
      for (int64_t i = a; i < b; i++) {
        const uint32_t rb = indptr[i], re = indptr[i + 1];
        if (rb == re) {
          next[i] = label[i];
          continue;
        }
        buf.resize(re - rb);
        for (uint32_t e = rb; e < re; e++) buf[e - rb] = label[indices[e]];
        std::sort(buf.begin(), buf.end());
        uint32_t best = buf[0], best_cnt = 0;
        uint64_t best_h = 0;
        for (size_t k = 0; k < buf.size();) {
          size_t k2 = k + 1;
          while (k2 < buf.size() && buf[k2] == buf[k]) k2++;
          const uint32_t cnt = (uint32_t)(k2 - k);
          const uint64_t h = mix64(salt ^ buf[k]);
          if (cnt > best_cnt || (cnt == best_cnt && h > best_h)) {
            best = buf[k];
            best_cnt = cnt;
            best_h = h;
          }
          k = k2;
        }
        next[i] = best;
        changed += best != label[i];
      }
      changed_total.fetch_add(changed, std::memory_order_relaxed);
    });
    label.swap(next);
    if (changed_total.load() * 200 < n) break;  // < 0.5 % of the nodes moved
  }
}

// nodes community by community: communities in order of their smallest member, members in their original order
// (stable => deterministic); order[k] = old id of the k-th node
void community_order(int64_t n, const std::vector<uint32_t> &label, std::vector<uint32_t> &order) {
  std::vector<uint32_t> first((size_t)n, 0xffffffffu);
  for (int64_t i = 0; i < n; i++)
    if (first[label[i]] == 0xffffffffu) first[label[i]] = (uint32_t)i;
  order.resize((size_t)n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return first[label[x]] < first[label[y]]; });
}

}  // namespace

extern "C" {

int gcnb_reorder_communities(int64_t n, const uint32_t *indptr, const uint32_t *indices, int max_sweeps, uint64_t seed,
                             uint32_t *new_of_old, int64_t *n_communities) {
  if (n < 0 || n > 0xffffffffll || !indptr || (!indices && n && indptr[n]) || !new_of_old) return GCNB_E_BADARG;
  std::vector<uint32_t> label;
  propagate_labels(n, indptr, indices, max_sweeps, seed, label);
  std::vector<uint32_t> first((size_t)n, 0xffffffffu);
  for (int64_t i = 0; i < n; i++)
    if (first[label[i]] == 0xffffffffu) first[label[i]] = (uint32_t)i;
  std::vector<uint32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return first[label[x]] < first[label[y]]; });
  int64_t comms = 0;
  for (int64_t k = 0; k < n; k++) {
    new_of_old[order[k]] = (uint32_t)k;
    if (k == 0 || label[order[k]] != label[order[k - 1]]) comms++;
  }
  if (n_communities) *n_communities = comms;
  return 0;
}

// Balanced, community-aligned row partition for the multi-GPU engine (north_star: "METIS-style").  The
// engine's ranks own contiguous row blocks and exchange the rows other ranks reference, so a good partition (i) gives every
// rank the same amount of GraphSum work -- CSR entries, not rows: degrees are skewed -- and (ii) cuts few edges, i.e. puts
// rank boundaries on community borders.  Communities come from the label propagation above; nodes are laid out community by
// community and the stream is cut at the entry-count targets k * nnz / world, each cut snapped to the nearest community
// border when one lies within `tolerance` (fraction of a rank's share, 0 = 0.1) of the target and placed inside the
// community otherwise (a community larger than a rank's share has to be split).  Rank r receives the ids
// [r * block, r * block + rows[r]) with block = the largest row count rounded up to a multiple of 4 (the engine's slab
// size); the ids in between are padding (isolated dummy nodes when the caller materialises them).
// stats: {communities, entries whose two ends lie on different ranks, the same count for the plain contiguous partition into
// equal row blocks of the GIVEN numbering, largest entry count of a rank}.
int gcnb_partition_communities(int64_t n, const uint32_t *indptr, const uint32_t *indices, int world, int max_sweeps,
                               uint64_t seed, double tolerance, uint32_t *new_of_old, int64_t *block_out,
                               int64_t *rows_out /* [world] */, int64_t stats[4]) {
  if (n < 1 || n > 0xffffffffll || world < 1 || !indptr || (!indices && indptr[n]) || !new_of_old || !block_out || !rows_out)
    return GCNB_E_BADARG;
  if (tolerance <= 0) tolerance = 0.1;
  std::vector<uint32_t> label, order;
  propagate_labels(n, indptr, indices, max_sweeps, seed, label);
  community_order(n, label, order);
  const uint64_t nnz = indptr[n];
  // prefix entry counts in the new order; community borders
  std::vector<uint64_t> pre((size_t)n + 1, 0);
  std::vector<int64_t> border;  // positions k where a community starts (k = 0 included), plus n
  for (int64_t k = 0; k < n; k++) {
    pre[k + 1] = pre[k] + (indptr[order[k] + 1] - indptr[order[k]]);
    if (k == 0 || label[order[k]] != label[order[k - 1]]) border.push_back(k);
  }
  const int64_t comms = (int64_t)border.size();
  border.push_back(n);
  std::vector<int64_t> cut((size_t)world + 1, n);
  cut[0] = 0;
  const double share = (double)nnz / world;
  for (int r = 1; r < world; r++) {
    const uint64_t target = (uint64_t)((double)nnz * r / world);
    int64_t k = std::lower_bound(pre.begin(), pre.end(), target) - pre.begin();  // first position with pre >= target
    k = std::min<int64_t>(std::max<int64_t>(k, cut[r - 1]), n);
    // nearest community border
    auto it = std::lower_bound(border.begin(), border.end(), k);
    int64_t best = -1;
    double best_d = 0;
    for (auto c : {it, it == border.begin() ? it : it - 1}) {
      if (c == border.end()) continue;
      const double d = std::fabs((double)pre[*c] - (double)target);
      if (*c >= cut[r - 1] && (best < 0 || d < best_d)) best = *c, best_d = d;
    }
    cut[r] = (best >= 0 && best_d <= tolerance * share) ? best : k;
  }
  int64_t max_rows = 0;
  uint64_t max_nnz = 0;
  for (int r = 0; r < world; r++) {
    rows_out[r] = cut[r + 1] - cut[r];
    max_rows = std::max(max_rows, rows_out[r]);
    max_nnz = std::max<uint64_t>(max_nnz, pre[cut[r + 1]] - pre[cut[r]]);
  }
  const int64_t block = (max_rows + 3) / 4 * 4;
  if ((uint64_t)block * (uint64_t)world > 0xffffffffull) return GCNB_E_UNSUPPORTED;
  std::vector<uint32_t> rank_of((size_t)n);
  for (int r = 0; r < world; r++)
    for (int64_t k = cut[r]; k < cut[r + 1]; k++) {
      new_of_old[order[k]] = (uint32_t)(r * block + (k - cut[r]));
      rank_of[order[k]] = (uint32_t)r;
    }
  *block_out = block;
  if (stats) {
    const int64_t eq = (n + world - 1) / world;  // the plain partition: equal contiguous row blocks of the given numbering
    std::atomic<int64_t> cut_new(0), cut_old(0);
    parallel_rows(n, indptr, [&](int64_t a, int64_t b) {
      int64_t cn = 0, co = 0;
      for (int64_t i = a; i < b; i++)
        for (uint32_t e = indptr[i]; e < indptr[i + 1]; e++) {
          const uint32_t j = indices[e];
          cn += rank_of[i] != rank_of[j];
          co += (i / eq) != (int64_t)(j / eq);
        }
      cut_new.fetch_add(cn);
      cut_old.fetch_add(co);
    });
    stats[0] = comms;
    stats[1] = cut_new.load();
    stats[2] = cut_old.load();
    stats[3] = (int64_t)max_nnz;
  }
  return 0;
}

// CSR of the renumbered graph: row new_of_old[i] holds { new_of_old[j] : j in row i }, the self entry first and the
// neighbours ascending (the parser's row convention when the file lists neighbours in order).  Duplicates are kept.
int gcnb_permute_csr(int64_t n, const uint32_t *indptr, const uint32_t *indices, const uint32_t *new_of_old,
                     uint32_t *out_indptr, uint32_t *out_indices) {
  if (n < 0 || !indptr || !new_of_old || !out_indptr || (n && indptr[n] && (!indices || !out_indices))) return GCNB_E_BADARG;
  std::vector<uint32_t> old_of_new((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    if (new_of_old[i] >= (uint64_t)n) return GCNB_E_BADARG;
    old_of_new[new_of_old[i]] = (uint32_t)i;
  }
  out_indptr[0] = 0;
  for (int64_t k = 0; k < n; k++) {
    const uint32_t i = old_of_new[k];
    out_indptr[k + 1] = out_indptr[k] + (indptr[i + 1] - indptr[i]);
  }
  parallel_rows(n, out_indptr, [&](int64_t a, int64_t b) {
    for (int64_t k = a; k < b; k++) {
      const uint32_t i = old_of_new[k];
      uint32_t *dst = out_indices + out_indptr[k];
      const uint32_t len = indptr[i + 1] - indptr[i];
      for (uint32_t e = 0; e < len; e++) dst[e] = new_of_old[indices[indptr[i] + e]];
      // keep a leading self entry in place (parser convention), sort the rest
      uint32_t lead = (len > 0 && indices[indptr[i]] == i) ? 1u : 0u;
      std::sort(dst + lead, dst + len);
    }
  });
  return 0;
}

// out[new_of_old[i]] = in[i] for rows of row_bytes bytes (features, labels, split); gcnb_unpermute_rows is the inverse
// (per-node outputs of a model trained on the renumbered dataset, back in the original numbering)
int gcnb_permute_rows(int64_t n, int64_t row_bytes, const uint32_t *new_of_old, const void *in, void *out) {
  if (n < 0 || row_bytes <= 0 || !new_of_old || !in || !out || in == out) return GCNB_E_BADARG;
  const unsigned nt = n_threads(n * row_bytes / 64);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      for (int64_t i = n * t / nt; i < n * (t + 1) / nt; i++)
        std::memcpy((char *)out + (size_t)new_of_old[i] * row_bytes, (const char *)in + (size_t)i * row_bytes, (size_t)row_bytes);
    });
  for (auto &x : th) x.join();
  return 0;
}

int gcnb_unpermute_rows(int64_t n, int64_t row_bytes, const uint32_t *new_of_old, const void *in, void *out) {
  if (n < 0 || row_bytes <= 0 || !new_of_old || !in || !out || in == out) return GCNB_E_BADARG;
  const unsigned nt = n_threads(n * row_bytes / 64);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      for (int64_t i = n * t / nt; i < n * (t + 1) / nt; i++)
        std::memcpy((char *)out + (size_t)i * row_bytes, (const char *)in + (size_t)new_of_old[i] * row_bytes, (size_t)row_bytes);
    });
  for (auto &x : th) x.join();
  return 0;
}

}  // extern "C"
