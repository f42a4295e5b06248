// parser.cpp -- .graph / .svmlight / .split -> host CSR, bit-compatible with the reference's Parser
// (src/parser.cpp:15-132,164-209) but written as a multi-threaded single-pass scanner over the whole file
// (std::from_chars, no istringstream): the reference takes minutes at Reddit scale.
// Semantics kept (all pinned by tests/test_parser.py against the reference's own parser):
//   * a line counts only when terminated by '\n' (getline + eof() check drops an unterminated last line);
//   * graph row i = [i, neighbours in file order...], duplicates kept, a non-integer token ends the line;
//   * svmlight "label k:v ...": a line without a leading integer gives label -1 and no features;
//     input_dim = max k + 1, output_dim = max label + 1;
//   * split: leading integer of each line; counts of 1/2/3 -> train_dim/val_dim/test_dim;
//   * graph_value[e] = 1. / sqrtf(deg(src) * deg(dst)) with the 32-bit unsigned product.
#include "../include/parser.h"
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace {

struct FileBuf {
  std::vector<char> data;
  bool ok = false;
};

FileBuf slurp(const std::string &path) {
  FileBuf fb;
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return fb;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  fb.data.resize(static_cast<size_t>(n));
  const size_t got = n ? std::fread(fb.data.data(), 1, static_cast<size_t>(n), f) : 0;
  fb.data.resize(got);
  std::fclose(f);
  fb.ok = true;
  return fb;
}

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// `stream >> int`: skip blanks, optional sign, digits.  Returns false (and leaves cur) when no integer starts here.
inline bool scan_int(const char *&cur, const char *end, long &out) {
  const char *p = cur;
  while (p < end && is_ws(*p)) p++;
  if (p >= end) return false;
  const char *q = p;
  if (*q == '+') q++;
  long v = 0;
  auto r = std::from_chars(q, end, v);
  if (r.ec != std::errc() || r.ptr == q) return false;
  cur = r.ptr;
  out = v;
  return true;
}

// chunk [begin, end) of complete lines per worker
struct Chunk {
  const char *begin, *end;
  size_t first_line = 0, n_lines = 0;
};

std::vector<Chunk> split_lines(const std::vector<char> &buf) {
  // usable region ends after the last '\n'
  const char *b = buf.data();
  const char *e = b + buf.size();
  while (e > b && e[-1] != '\n') e--;
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (static_cast<size_t>(e - b) < (1u << 20)) nt = 1;
  std::vector<Chunk> chunks;
  const char *cur = b;
  for (unsigned t = 0; t < nt && cur < e; t++) {
    const char *stop = (t + 1 == nt) ? e : b + (e - b) * (t + 1) / nt;
    if (stop < cur) stop = cur;
    while (stop < e && stop[-1] != '\n') stop++;
    if (stop > cur) chunks.push_back({cur, stop});
    cur = stop;
  }
  std::vector<std::thread> th;
  for (auto &c : chunks)
    th.emplace_back([&c] {
      size_t n = 0;
      for (const char *p = c.begin; p < c.end;) {
        const char *nl = static_cast<const char *>(std::memchr(p, '\n', c.end - p));
        n++;
        p = nl + 1;
      }
      c.n_lines = n;
    });
  for (auto &t : th) t.join();
  size_t acc = 0;
  for (auto &c : chunks) {
    c.first_line = acc;
    acc += c.n_lines;
  }
  return chunks;
}

template <class F>
void for_each_chunk(std::vector<Chunk> &chunks, F fn) {
  std::vector<std::thread> th;
  for (size_t i = 0; i < chunks.size(); i++) th.emplace_back([&, i] { fn(i, chunks[i]); });
  for (auto &t : th) t.join();
}

}  // namespace

Parser::Parser(GCNParams *gcnParams_, GCNData *gcnData_, std::string graph_name, bool no_feature_, bool quiet_)
    : graph_path("data/" + graph_name + ".graph"), split_path("data/" + graph_name + ".split"),
      svmlight_path("data/" + graph_name + ".svmlight"), gcnParams(gcnParams_), gcnData(gcnData_),
      no_feature(no_feature_), quiet(quiet_) {}

bool Parser::isValidInput() {
  for (const auto *p : {&graph_path, &split_path, &svmlight_path}) {
    FILE *f = std::fopen(p->c_str(), "rb");
    if (!f) return false;
    std::fclose(f);
  }
  return true;
}

void Parser::parseGraph() {
  auto &graph = gcnData->graph;
  FileBuf fb = slurp(graph_path);
  auto chunks = split_lines(fb.data);
  size_t n_nodes = 0;
  for (auto &c : chunks) n_nodes += c.n_lines;
  std::vector<std::vector<natural>> idx(chunks.size());
  std::vector<std::vector<natural>> deg(chunks.size());
  for_each_chunk(chunks, [&](size_t ci, Chunk &c) {
    auto &ix = idx[ci];
    auto &dg = deg[ci];
    dg.reserve(c.n_lines);
    ix.reserve(static_cast<size_t>(c.end - c.begin) / 4);
    size_t node = c.first_line;
    for (const char *p = c.begin; p < c.end; node++) {
      const char *nl = static_cast<const char *>(std::memchr(p, '\n', c.end - p));
      natural d = 1;
      ix.push_back(static_cast<natural>(node));  // implicit self connection first (src/parser.cpp:28-31)
      long nb;
      const char *cur = p;
      while (scan_int(cur, nl, nb)) {
        ix.push_back(static_cast<natural>(nb));
        d++;
      }
      dg.push_back(d);
      p = nl + 1;
    }
  });
  graph.indptr.assign(n_nodes + 1, 0);
  size_t total = 0, row = 0;
  for (size_t ci = 0; ci < chunks.size(); ci++)
    for (natural d : deg[ci]) {
      total += d;
      graph.indptr[++row] = static_cast<natural>(total);
    }
  graph.indices.resize(total);
  size_t off = 0;
  for (auto &ix : idx) {
    std::memcpy(graph.indices.data() + off, ix.data(), ix.size() * sizeof(natural));
    off += ix.size();
  }
  gcnParams->num_nodes = static_cast<natural>(n_nodes);
}

void Parser::parseNode() {
  auto &fidx = gcnData->feature_index;
  auto &fval = gcnData->feature_value;
  auto &labels = gcnData->label;
  FileBuf fb = slurp(svmlight_path);
  auto chunks = split_lines(fb.data);
  struct Part {
    std::vector<natural> idx, cnt;
    std::vector<real> val;
    std::vector<integer> lab;
    long max_idx = 0, max_label = 0;
  };
  std::vector<Part> parts(chunks.size());
  const bool nofeat = no_feature;
  for_each_chunk(chunks, [&](size_t ci, Chunk &c) {
    Part &pt = parts[ci];
    pt.cnt.reserve(c.n_lines);
    pt.lab.reserve(c.n_lines);
    for (const char *p = c.begin; p < c.end;) {
      const char *nl = static_cast<const char *>(std::memchr(p, '\n', c.end - p));
      const char *cur = p;
      long label = -1;
      const bool ok = scan_int(cur, nl, label);
      pt.lab.push_back(ok ? static_cast<integer>(label) : -1);
      natural n = 0;
      if (ok) {
        pt.max_label = std::max(pt.max_label, label);
        for (;;) {
          while (cur < nl && is_ws(*cur)) cur++;
          if (cur >= nl) break;
          const char *tok_end = cur;
          while (tok_end < nl && !is_ws(*tok_end)) tok_end++;
          // token "k:v" -- `kv_ss >> k >> col >> v` (src/parser.cpp:93-98)
          long k = 0;
          auto r = std::from_chars(*cur == '+' ? cur + 1 : cur, tok_end, k);
          float v = 0.f;
          if (r.ec == std::errc() && r.ptr < tok_end) {
            const char *vp = r.ptr + 1;  // skip the separator character
            if (vp < tok_end && *vp == '+') vp++;
            std::from_chars(vp, tok_end, v);
          }
          pt.val.push_back(nofeat ? 1.0f : v);
          pt.idx.push_back(static_cast<natural>(k));
          pt.max_idx = std::max(pt.max_idx, k);
          n++;
          cur = tok_end;
        }
      }
      pt.cnt.push_back(n);
      p = nl + 1;
    }
  });
  size_t rows = 0, nnz = 0;
  long max_idx = 0, max_label = 0;
  for (auto &pt : parts) {
    rows += pt.cnt.size();
    nnz += pt.idx.size();
    max_idx = std::max(max_idx, pt.max_idx);
    max_label = std::max(max_label, pt.max_label);
  }
  fidx.indptr.assign(rows + 1, 0);
  fidx.indices.resize(nnz);
  fval.resize(nnz);
  labels.resize(rows);
  size_t r = 0, off = 0, acc = 0;
  for (auto &pt : parts) {
    for (size_t i = 0; i < pt.cnt.size(); i++) {
      acc += pt.cnt[i];
      labels[r] = pt.lab[i];
      fidx.indptr[++r] = static_cast<natural>(acc);
    }
    std::memcpy(fidx.indices.data() + off, pt.idx.data(), pt.idx.size() * sizeof(natural));
    std::memcpy(fval.data() + off, pt.val.data(), pt.val.size() * sizeof(real));
    off += pt.idx.size();
  }
  gcnParams->input_dim = static_cast<natural>(max_idx + 1);
  gcnParams->output_dim = static_cast<natural>(max_label + 1);
}

void Parser::parseSplit() {
  auto &split = gcnData->split;
  FileBuf fb = slurp(split_path);
  auto chunks = split_lines(fb.data);
  size_t n = 0;
  for (auto &c : chunks) n += c.n_lines;
  split.assign(n, 0);
  for_each_chunk(chunks, [&](size_t, Chunk &c) {
    size_t line = c.first_line;
    for (const char *p = c.begin; p < c.end; line++) {
      const char *nl = static_cast<const char *>(std::memchr(p, '\n', c.end - p));
      long v = 0;
      const char *cur = p;
      scan_int(cur, nl, v);
      split[line] = static_cast<natural>(v);
      p = nl + 1;
    }
  });
  for (natural v : split) {
    if (v == 1) gcnParams->train_dim++;
    else if (v == 2) gcnParams->val_dim++;
    else if (v == 3) gcnParams->test_dim++;
  }
}

void Parser::calculateGraphValues() {
  const auto &g = gcnData->graph;
  auto &val = gcnData->graph_value;
  val.resize(g.indices.size());
  const size_t n = g.indptr.size() - 1;
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (g.indices.size() < (1u << 20)) nt = 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      for (size_t src = n * t / nt; src < n * (t + 1) / nt; src++)
        for (natural e = g.indptr[src]; e < g.indptr[src + 1]; e++) {
          const natural dst = g.indices[e];
          val[e] = 1. / sqrtf((g.indptr[src + 1] - g.indptr[src]) * (g.indptr[dst + 1] - g.indptr[dst]));
        }
    });
  for (auto &t : th) t.join();
}

bool Parser::parse() {
  if (!isValidInput()) return false;
  if (!quiet) std::cout << "PARSING DATA ..." << std::endl;
  parseGraph();
  if (!quiet) std::cout << "Parse Graph Succeeded." << std::endl;
  parseNode();
  if (!quiet) std::cout << "Parse Node Succeeded." << std::endl;
  parseSplit();
  if (!quiet) std::cout << "Parse Split Succeeded." << std::endl;
  calculateGraphValues();
  return true;
}
