// synth.cpp -- synthetic workloads of BASELINE.json configs 3-5 (the Reddit files are not shipped with the
// reference: .MISSING_LARGE_BLOBS).  Deterministic in `seed` and independent of the number of host threads:
// every random quantity is a pure hash of (seed, stream, index).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "../../../include/gcnb_engine.h"

namespace {

inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t i) { return mix64(mix64(seed ^ (stream << 56)) + i); }
inline double u01(uint64_t r) { return ((r >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
inline double normal(uint64_t seed, uint64_t stream, uint64_t i) {
  const double a = u01(rnd(seed, stream, 2 * i)), b = u01(rnd(seed, stream, 2 * i + 1));
  return std::sqrt(-2.0 * std::log(a)) * std::cos(6.283185307179586 * b);
}

unsigned n_threads() { return std::max(1u, std::min(64u, std::thread::hardware_concurrency())); }

template <class F>
void parallel_for(int64_t n, F fn) {
  const unsigned nt = n < (1 << 16) ? 1 : n_threads();
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) th.emplace_back([=] { fn(n * t / nt, n * (t + 1) / nt); });
  for (auto &x : th) x.join();
}

// parallel sort + unique of 64-bit keys by bucketing on the high part (keys are (a << 32) | b with a < n)
void sort_keys(std::vector<uint64_t> &keys, int64_t n) {
  const unsigned nb = keys.size() < (1u << 18) ? 1 : n_threads();
  if (nb == 1) {
    std::sort(keys.begin(), keys.end());
    return;
  }
  auto bucket_of = [&](uint64_t k) { return (unsigned)(((k >> 32) * nb) / (uint64_t)n); };
  std::vector<size_t> cnt(nb + 1, 0);
  for (uint64_t k : keys) cnt[bucket_of(k) + 1]++;
  for (unsigned b = 0; b < nb; b++) cnt[b + 1] += cnt[b];
  std::vector<uint64_t> tmp(keys.size());
  {
    std::vector<size_t> cur(cnt.begin(), cnt.end() - 1);
    for (uint64_t k : keys) tmp[cur[bucket_of(k)]++] = k;
  }
  std::vector<std::thread> th;
  for (unsigned b = 0; b < nb; b++)
    th.emplace_back([&, b] { std::sort(tmp.begin() + cnt[b], tmp.begin() + cnt[b + 1]); });
  for (auto &x : th) x.join();
  keys.swap(tmp);
}

}  // namespace

extern "C" {

void gcnb_host_free(void *p) { std::free(p); }

int gcnb_synth_graph(int64_t n, int64_t m_target, int n_blocks, double intra, double sigma, int64_t max_deg,
                     uint64_t seed, uint32_t **indptr_out, uint32_t **indices_out, int64_t *nnz_out) {
  if (n <= 1 || m_target < 0 || n_blocks < 1 || !indptr_out || !indices_out || !nnz_out) return GCNB_E_BADARG;
  if ((double)m_target > 0.4 * (double)n * (double)(n - 1) / 2.0) return GCNB_E_BADARG;  // too dense to top up
  const int64_t bs = (n + n_blocks - 1) / n_blocks;
  // expected-degree weights: lognormal, scaled so that sum = 2m, clipped to [., max_deg] (two rescale rounds)
  std::vector<double> w((size_t)n);
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) w[i] = std::exp(sigma * normal(seed, 1, (uint64_t)i));
  });
  for (int round = 0; round < 4; round++) {
    double s = 0;
    for (double x : w) s += x;
    const double k = 2.0 * (double)m_target / s;
    for (double &x : w) x = std::min(x * k, (double)max_deg);
  }
  // cumulative weights, global and per block
  std::vector<double> cdf((size_t)n + 1, 0.0);
  for (int64_t i = 0; i < n; i++) cdf[i + 1] = cdf[i] + w[i];
  const double total = cdf[n];
  auto pick = [&](double lo, double hi, double u) -> int64_t {  // node whose cdf interval holds lo + u*(hi-lo)
    const double x = lo + u * (hi - lo);
    int64_t i = (int64_t)(std::upper_bound(cdf.begin(), cdf.end(), x) - cdf.begin()) - 1;
    return std::max<int64_t>(0, std::min<int64_t>(n - 1, i));
  };
  std::vector<uint64_t> keys;
  keys.reserve((size_t)(m_target * 1.02) + 16);
  int64_t next_edge_id = 0;
  for (int round = 0; round < 64 && (int64_t)keys.size() < m_target; round++) {
    const int64_t need = m_target - (int64_t)keys.size();
    const int64_t batch = need + need / 50 + 16;
    const size_t old = keys.size();
    keys.resize(old + (size_t)batch);
    const int64_t id0 = next_edge_id;
    parallel_for(batch, [&](int64_t a, int64_t b) {
      for (int64_t e = a; e < b; e++) {
        const uint64_t id = (uint64_t)(id0 + e);
        int64_t u = 0, v = 0;
        for (uint64_t attempt = 0;; attempt++) {
          u = pick(0.0, total, u01(rnd(seed, 2, id * 8 + attempt * 3)));
          const bool local = u01(rnd(seed, 2, id * 8 + attempt * 3 + 1)) < intra;
          const double ur = u01(rnd(seed, 2, id * 8 + attempt * 3 + 2));
          if (local) {
            const int64_t b0 = (u / bs) * bs, b1 = std::min(n, b0 + bs);
            v = pick(cdf[b0], cdf[b1], ur);
          } else {
            v = pick(0.0, total, ur);
          }
          if (u != v || attempt > 16) break;
        }
        if (u == v) v = (u + 1) % n;
        const uint64_t lo = (uint64_t)std::min(u, v), hi = (uint64_t)std::max(u, v);
        keys[old + (size_t)e] = (lo << 32) | hi;
      }
    });
    next_edge_id += batch;
    sort_keys(keys, n);
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    if ((int64_t)keys.size() > m_target) {
      // drop a deterministic subset of the surplus: the highest-hash edges
      const size_t surplus = keys.size() - (size_t)m_target;
      std::vector<std::pair<uint64_t, uint64_t>> h;
      h.reserve(keys.size());
      for (uint64_t k : keys) h.emplace_back(mix64(k ^ seed), k);
      std::nth_element(h.begin(), h.begin() + surplus, h.end(), std::greater<std::pair<uint64_t, uint64_t>>());
      std::vector<uint64_t> drop;
      drop.reserve(surplus);
      for (size_t i = 0; i < surplus; i++) drop.push_back(h[i].second);
      std::sort(drop.begin(), drop.end());
      std::vector<uint64_t> kept;
      kept.reserve((size_t)m_target);
      std::set_difference(keys.begin(), keys.end(), drop.begin(), drop.end(), std::back_inserter(kept));
      keys.swap(kept);
    }
  }
  if ((int64_t)keys.size() != m_target) return GCNB_E_UNSUPPORTED;
  // symmetric CSR with the parser's convention: row i = [i, sorted neighbours]
  const int64_t m = m_target;
  const int64_t nnz = 2 * m + n;
  uint32_t *indptr = (uint32_t *)std::malloc(((size_t)n + 1) * 4);
  uint32_t *indices = (uint32_t *)std::malloc((size_t)nnz * 4);
  if (!indptr || !indices) return GCNB_E_UNSUPPORTED;
  std::vector<uint32_t> deg((size_t)n, 1);
  for (uint64_t k : keys) {
    deg[k >> 32]++;
    deg[k & 0xffffffffu]++;
  }
  indptr[0] = 0;
  for (int64_t i = 0; i < n; i++) indptr[i + 1] = indptr[i] + deg[i];
  std::vector<uint32_t> cur((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    indices[indptr[i]] = (uint32_t)i;
    cur[i] = indptr[i] + 1;
  }
  // keys are sorted by (lo, hi): pass 1 writes, for every node, its smaller neighbours in ascending order
  // (pairs (lo, hi) visit a fixed hi with ascending lo), pass 2 appends the larger neighbours (ascending hi).
  for (uint64_t k : keys) indices[cur[k & 0xffffffffu]++] = (uint32_t)(k >> 32);
  for (uint64_t k : keys) indices[cur[k >> 32]++] = (uint32_t)(k & 0xffffffffu);
  *indptr_out = indptr;
  *indices_out = indices;
  *nnz_out = nnz;
  return 0;
}

int gcnb_synth_dense_features(int64_t n, int f, uint64_t seed, uint32_t *indptr, uint32_t *indices, float *values) {
  if (n < 0 || f <= 0 || !indptr || !indices || !values) return GCNB_E_BADARG;
  if ((uint64_t)n * (uint64_t)f > 0xffffffffull) return GCNB_E_UNSUPPORTED;
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) {
      indptr[i] = (uint32_t)(i * f);
      for (int j = 0; j < f; j++) {
        const size_t e = (size_t)i * f + j;
        indices[e] = (uint32_t)j;
        values[e] = (float)normal(seed, 3, e);
      }
    }
  });
  indptr[n] = (uint32_t)(n * f);
  return 0;
}

int gcnb_synth_labels(int64_t n, int classes, double frac_train, double frac_val, uint64_t seed, int32_t *label,
                      uint32_t *split) {
  if (n < 0 || classes <= 0 || !label || !split) return GCNB_E_BADARG;
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) {
      label[i] = (int32_t)(rnd(seed, 4, (uint64_t)i) % (uint64_t)classes);
      const double u = u01(rnd(seed, 5, (uint64_t)i));
      split[i] = u < frac_train ? 1u : (u < frac_train + frac_val ? 2u : 3u);
    }
  });
  return 0;
}

}  // extern "C"
