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

}  // namespace

extern "C" {

int gcnb_reorder_communities(int64_t n, const uint32_t *indptr, const uint32_t *indices, int max_sweeps, uint64_t seed,
                             uint32_t *new_of_old, int64_t *n_communities) {
  if (n < 0 || n > 0xffffffffll || !indptr || (!indices && n && indptr[n]) || !new_of_old) return GCNB_E_BADARG;
  if (max_sweeps <= 0) max_sweeps = 8;
  std::vector<uint32_t> label((size_t)n), next((size_t)n);
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
  // communities in order of their smallest member, members in their original order (stable => deterministic)
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
