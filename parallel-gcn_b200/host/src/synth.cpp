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

// ---- row-local symmetric generator (scale-out workloads: every rank builds ONLY its row block) -------------------------
// Whether the undirected edge {i, j} exists is a pure function of (seed, min(i,j), max(i,j)), so any row can be
// generated without looking at any other row and the union of all rows is a symmetric simple graph:
//   * inside a community (contiguous blocks of `block_size` nodes) every pair is a candidate, accepted with
//     probability p_in * w_i * w_j;
//   * across communities the candidates of node i are its images under `n_reflect` fixed reflections
//     j = (c_k - i) mod n (an involution: i is then the k-th candidate of j), accepted with p_out * w_i * w_j;
//   * w = lognormal(sigma) expected-degree weights with mean ~1, clipped so that no probability exceeds 1.
int gcnb_synth_sym_rows(int64_t n, int64_t row0, int64_t rows, int64_t block_size, double mean_intra, double mean_inter,
                        int n_reflect, double sigma, uint64_t seed, uint32_t **indptr_out, uint32_t **indices_out,
                        int64_t *nnz_out) {
  return gcnb_synth_sym_rows_local(n, row0, rows, block_size, mean_intra, mean_inter, n_reflect, 0, sigma, seed, indptr_out,
                                   indices_out, nnz_out);
}

// ... with inter_window > 0 the edges that leave a community stay NEAR it instead: the candidates of node i across
// communities are i +- d_k for n_reflect / 2 fixed shifts d_k in [block_size, inter_window] (symmetric: i is the
// candidate of i + d_k under -d_k), so a row block references only the rows within inter_window of its borders -- the
// structure a METIS-style partition of a citation / co-purchase graph has, and the case the halo exchange is for.
int gcnb_synth_sym_rows_local(int64_t n, int64_t row0, int64_t rows, int64_t block_size, double mean_intra, double mean_inter,
                              int n_reflect, int64_t inter_window, double sigma, uint64_t seed, uint32_t **indptr_out,
                              uint32_t **indices_out, int64_t *nnz_out) {
  if (n <= 1 || n > 0xffffffffll || row0 < 0 || rows < 0 || row0 + rows > n || block_size < 2 || n_reflect < 0 ||
      mean_intra < 0 || mean_inter < 0 || !indptr_out || !indices_out || !nnz_out ||
      (inter_window != 0 && inter_window < block_size))
    return GCNB_E_BADARG;
  const double p_in = std::min(1.0, mean_intra / (double)block_size);
  const double p_out = n_reflect > 0 ? std::min(1.0, mean_inter / (double)n_reflect) : 0.0;
  const double w_max = std::sqrt(1.0 / std::max({p_in, p_out, 1e-12}));
  std::vector<float> w((size_t)n);
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++)
      w[i] = (float)std::min(w_max, std::exp(sigma * normal(seed, 11, (uint64_t)i) - 0.5 * sigma * sigma));
  });
  std::vector<int64_t> refl((size_t)n_reflect);
  for (int k = 0; k < n_reflect; k++) refl[k] = (int64_t)(rnd(seed, 12, (uint64_t)k) % (uint64_t)n);
  std::vector<int64_t> shift;  // inter_window > 0: +-d for n_reflect / 2 distinct shifts d
  if (inter_window > 0) {
    const uint64_t span = (uint64_t)(inter_window - block_size + 1);
    for (int k = 0; k < n_reflect / 2; k++) shift.push_back(block_size + (int64_t)(rnd(seed, 14, (uint64_t)k) % span));
    std::sort(shift.begin(), shift.end());
    shift.erase(std::unique(shift.begin(), shift.end()), shift.end());
  }
  const uint64_t pair_seed = mix64(seed ^ 0x5eed5eedull);
  auto accept = [&](uint64_t i, uint64_t j, double p) {
    const uint64_t lo = std::min(i, j), hi = std::max(i, j);
    return u01(mix64(pair_seed ^ ((lo << 32) | hi))) < p * (double)w[lo] * (double)w[hi];
  };
  const unsigned nt = rows < 1024 ? 1 : n_threads();
  std::vector<std::vector<uint32_t>> t_idx(nt);
  std::vector<uint32_t> deg((size_t)rows, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      std::vector<uint32_t> &out = t_idx[t];
      std::vector<uint32_t> row;
      for (int64_t r = rows * t / nt; r < rows * (t + 1) / nt; r++) {
        const int64_t i = row0 + r;
        const int64_t b0 = (i / block_size) * block_size, b1 = std::min(n, b0 + block_size);
        row.clear();
        for (int64_t j = b0; j < b1; j++)
          if (j != i && accept((uint64_t)i, (uint64_t)j, p_in)) row.push_back((uint32_t)j);
        const size_t n_in = row.size();
        if (inter_window > 0) {
          for (const int64_t d : shift)
            for (const int64_t j : {i - d, i + d}) {
              if (j < 0 || j >= n || (j >= b0 && j < b1)) continue;
              if (accept((uint64_t)i, (uint64_t)j, p_out)) row.push_back((uint32_t)j);
            }
        } else {
          for (int k = 0; k < n_reflect; k++) {
            int64_t j = refl[k] - i;
            if (j < 0) j += n;
            if (j == i || (j >= b0 && j < b1)) continue;
            if (accept((uint64_t)i, (uint64_t)j, p_out)) row.push_back((uint32_t)j);
          }
        }
        std::sort(row.begin() + n_in, row.end());
        row.erase(std::unique(row.begin() + n_in, row.end()), row.end());
        std::inplace_merge(row.begin(), row.begin() + n_in, row.end());
        deg[r] = (uint32_t)row.size() + 1;
        out.push_back((uint32_t)i);
        out.insert(out.end(), row.begin(), row.end());
      }
    });
  for (auto &x : th) x.join();
  uint64_t nnz = 0;
  for (auto &v : t_idx) nnz += v.size();
  if (nnz > 0xfffffff0ull) return GCNB_E_UNSUPPORTED;
  uint32_t *indptr = (uint32_t *)std::malloc(((size_t)rows + 1) * 4);
  uint32_t *indices = (uint32_t *)std::malloc(std::max<size_t>(4, (size_t)nnz * 4));
  if (!indptr || !indices) return GCNB_E_UNSUPPORTED;
  indptr[0] = 0;
  for (int64_t r = 0; r < rows; r++) indptr[r + 1] = indptr[r] + deg[r];
  th.clear();
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      if (!t_idx[t].empty()) std::memcpy(indices + indptr[rows * t / nt], t_idx[t].data(), t_idx[t].size() * 4);
    });
  for (auto &x : th) x.join();
  *indptr_out = indptr;
  *indices_out = indices;
  *nnz_out = (int64_t)nnz;
  return 0;
}

// graph_value of a row block from GLOBAL degrees, the reference's arithmetic (src/parser.cpp:164-181):
// 1. / sqrtf(unsigned(deg_src * deg_dst)), the divide in double, then rounded to fp32
int gcnb_synth_graph_values(const uint32_t *indptr, const uint32_t *indices, int64_t rows, int64_t row0,
                            const uint32_t *deg_global, float *out) {
  if (!indptr || !indices || !deg_global || !out || rows < 0) return GCNB_E_BADARG;
  parallel_for(rows, [&](int64_t a, int64_t b) {
    for (int64_t r = a; r < b; r++) {
      const uint32_t ds = deg_global[row0 + r];
      for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) out[e] = (float)(1. / sqrtf((float)(ds * deg_global[indices[e]])));
    }
  });
  return 0;
}

// cheap dense features for the large workloads: uniform on [-sqrt(3), sqrt(3)) (unit variance), one hash per two values
int gcnb_synth_dense_features_uniform(int64_t n, int f, uint64_t seed, uint64_t elem_offset, uint32_t *indptr,
                                      uint32_t *indices, float *values) {
  if (n < 0 || f <= 0 || !indptr || !indices || !values) return GCNB_E_BADARG;
  if ((uint64_t)n * (uint64_t)f > 0xffffffffull) return GCNB_E_UNSUPPORTED;
  parallel_for(n, [&](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; i++) {
      indptr[i] = (uint32_t)(i * f);
      for (int j = 0; j < f; j++) {
        const size_t e = (size_t)i * f + j;
        const uint64_t g = elem_offset + e;
        const uint64_t h = rnd(seed, 13, g >> 1);
        const uint32_t bits = (g & 1) ? (uint32_t)(h >> 32) : (uint32_t)h;
        indices[e] = (uint32_t)j;
        values[e] = ((float)(bits >> 8) * (1.0f / 16777216.0f) - 0.5f) * 3.4641016f;
      }
    }
  });
  indptr[n] = (uint32_t)(n * f);
  return 0;
}

}  // extern "C"
