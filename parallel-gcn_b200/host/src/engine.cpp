// engine.cpp -- engine-level C ABI (include/gcnb_engine.h): Parser + GCN behind plain C with host buffers.
#include <cstring>
#include <memory>
#include "../../../include/gcnb_engine.h"
#include "../include/gcn.cuh"
#include "../include/parser.h"
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <thread>
#include <type_traits>

struct gcnb_dataset {
  GCNParams params;
  GCNData data;
};

struct gcnb_gcn {
  GCNParams params;
  AdamParams adam;
  std::unique_ptr<GCN> gcn;
  std::vector<std::vector<unsigned char>> unused;
  std::vector<const unsigned char *> masks;
  double last_exchange_ms = 0;  // of the last gcnb_gcn_timed_epochs call with GraphSum timing
};

extern "C" {

int gcnb_dataset_parse(const char *root, const char *name, int no_feature, gcnb_dataset **out) {
  if (!root || !name || !out) return GCNB_E_BADARG;
  *out = nullptr;
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return GCNB_E_BADARG;
  if (chdir(root) != 0) return GCNB_E_BADARG;  // Parser opens data/<name>.* relative to the CWD (src/parser.cpp:3-9)
  auto d = std::make_unique<gcnb_dataset>();
  bool ok;
  {
    Parser parser(&d->params, &d->data, name, no_feature != 0, true);
    ok = parser.parse();
  }
  if (chdir(cwd) != 0) ok = false;
  if (!ok) return GCNB_E_BADARG;
  *out = d.release();
  return 0;
}

int gcnb_dataset_dims(const gcnb_dataset *d, int64_t dims[10]) {
  if (!d || !dims) return GCNB_E_BADARG;
  dims[0] = d->params.num_nodes;
  dims[1] = (int64_t)d->data.graph.indices.size();
  dims[2] = (int64_t)d->data.feature_index.indptr.size() - 1;
  dims[3] = (int64_t)d->data.feature_index.indices.size();
  dims[4] = d->params.input_dim;
  dims[5] = d->params.output_dim;
  dims[6] = (int64_t)d->data.split.size();
  dims[7] = d->params.train_dim;
  dims[8] = d->params.val_dim;
  dims[9] = d->params.test_dim;
  return 0;
}

int gcnb_dataset_copy(const gcnb_dataset *d, int which, void *dst) {
  if (!d || !dst) return GCNB_E_BADARG;
  auto cp = [&](const void *src, size_t bytes) { std::memcpy(dst, src, bytes); };
  switch (which) {
    case 0: cp(d->data.graph.indptr.data(), d->data.graph.indptr.size() * 4); break;
    case 1: cp(d->data.graph.indices.data(), d->data.graph.indices.size() * 4); break;
    case 2: cp(d->data.feature_index.indptr.data(), d->data.feature_index.indptr.size() * 4); break;
    case 3: cp(d->data.feature_index.indices.data(), d->data.feature_index.indices.size() * 4); break;
    case 4: cp(d->data.feature_value.data(), d->data.feature_value.size() * 4); break;
    case 5: cp(d->data.label.data(), d->data.label.size() * 4); break;
    case 6: cp(d->data.split.data(), d->data.split.size() * 4); break;
    case 7: cp(d->data.graph_value.data(), d->data.graph_value.size() * 4); break;
    default: return GCNB_E_BADARG;
  }
  return 0;
}

int gcnb_dataset_free(gcnb_dataset *d) {
  delete d;
  return 0;
}

// ---- binary container -------------------------------------------------------------------------------------
// The text formats (.graph / .split / .svmlight) cost a tokenising pass per load -- minutes at Reddit scale with the
// reference's istringstream parser, seconds with this library's -- so a parsed dataset can be stored once and mapped back:
// a 128-byte header followed by the eight arrays of gcnb_dataset_copy, each starting on a 64-byte boundary, exactly the
// bytes the parser produced (loading a stored dataset is bit-identical to parsing the text again).
namespace {
struct BinHeader {
  char magic[8];       // "GCNBDS1\0"
  uint64_t version;    // 1
  uint64_t num_nodes, input_dim, output_dim, train_dim, val_dim, test_dim;
  uint64_t count[8];   // elements of: graph indptr, graph indices, feat indptr, feat indices, feat value, label, split, graph value
};
static_assert(sizeof(BinHeader) == 128, "header layout");
constexpr char kMagic[8] = {'G', 'C', 'N', 'B', 'D', 'S', '1', '\0'};
inline size_t pad64(size_t x) { return (x + 63) & ~(size_t)63; }
}  // namespace

int gcnb_dataset_save(const gcnb_dataset *d, const char *path) {
  if (!d || !path) return GCNB_E_BADARG;
  BinHeader h{};
  std::memcpy(h.magic, kMagic, 8);
  h.version = 1;
  h.num_nodes = d->params.num_nodes;
  h.input_dim = d->params.input_dim;
  h.output_dim = d->params.output_dim;
  h.train_dim = d->params.train_dim;
  h.val_dim = d->params.val_dim;
  h.test_dim = d->params.test_dim;
  const void *ptr[8] = {d->data.graph.indptr.data(), d->data.graph.indices.data(), d->data.feature_index.indptr.data(),
                        d->data.feature_index.indices.data(), d->data.feature_value.data(), d->data.label.data(),
                        d->data.split.data(), d->data.graph_value.data()};
  const size_t cnt[8] = {d->data.graph.indptr.size(), d->data.graph.indices.size(), d->data.feature_index.indptr.size(),
                         d->data.feature_index.indices.size(), d->data.feature_value.size(), d->data.label.size(),
                         d->data.split.size(), d->data.graph_value.size()};
  for (int k = 0; k < 8; k++) h.count[k] = cnt[k];
  FILE *f = fopen(path, "wb");
  if (!f) return GCNB_E_BADARG;
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  static const char zeros[64] = {0};
  for (int k = 0; k < 8 && ok; k++) {
    const size_t bytes = cnt[k] * 4;
    if (bytes) ok = fwrite(ptr[k], 1, bytes, f) == bytes;
    const size_t fill = pad64(bytes) - bytes;
    if (ok && fill) ok = fwrite(zeros, 1, fill, f) == fill;
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? 0 : GCNB_E_BADARG;
}

int gcnb_dataset_load(const char *path, gcnb_dataset **out) {
  if (!path || !out) return GCNB_E_BADARG;
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return GCNB_E_BADARG;
  struct stat sb;
  if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < sizeof(BinHeader)) {
    close(fd);
    return GCNB_E_BADARG;
  }
  void *map = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) return GCNB_E_BADARG;
  const BinHeader *h = static_cast<const BinHeader *>(map);
  size_t need = sizeof(BinHeader);
  bool ok = std::memcmp(h->magic, kMagic, 8) == 0 && h->version == 1;
  for (int k = 0; k < 8 && ok; k++) {
    ok = h->count[k] <= ((size_t)1 << 40);
    need += pad64((size_t)h->count[k] * 4);
  }
  ok = ok && need <= (size_t)sb.st_size && h->count[0] == h->num_nodes + 1 && h->count[2] == h->count[5] + 1 &&
       h->count[3] == h->count[4] && (h->count[7] == 0 || h->count[7] == h->count[1]);
  if (!ok) {
    munmap(map, (size_t)sb.st_size);
    return GCNB_E_BADARG;
  }
  auto d = std::make_unique<gcnb_dataset>();
  d->params.num_nodes = (natural)h->num_nodes;
  d->params.input_dim = (natural)h->input_dim;
  d->params.output_dim = (natural)h->output_dim;
  d->params.train_dim = (natural)h->train_dim;
  d->params.val_dim = (natural)h->val_dim;
  d->params.test_dim = (natural)h->test_dim;
  const char *p = static_cast<const char *>(map) + sizeof(BinHeader);
  auto take = [&](auto &vec, int k) {
    using T = typename std::remove_reference<decltype(vec)>::type::value_type;
    static_assert(sizeof(T) == 4, "32-bit arrays");
    vec.resize((size_t)h->count[k]);
    if (h->count[k]) std::memcpy(vec.data(), p, (size_t)h->count[k] * 4);
    p += pad64((size_t)h->count[k] * 4);
  };
  take(d->data.graph.indptr, 0);
  take(d->data.graph.indices, 1);
  take(d->data.feature_index.indptr, 2);
  take(d->data.feature_index.indices, 3);
  take(d->data.feature_value, 4);
  take(d->data.label, 5);
  take(d->data.split, 6);
  take(d->data.graph_value, 7);
  munmap(map, (size_t)sb.st_size);
  // a short or crafted file must be an error here, not an out-of-bounds read on the device later: label / split per node,
  // monotone offsets that end at the entry counts, indices inside their dimensions
  const size_t n = d->params.num_nodes;
  auto csr_ok = [](const std::vector<natural> &ip, const std::vector<natural> &ix, size_t dim) {
    if (ip.empty() || ip.front() != 0 || ip.back() != ix.size()) return false;
    for (size_t i = 0; i + 1 < ip.size(); i++)
      if (ip[i + 1] < ip[i]) return false;
    for (natural c : ix)
      if (c >= dim) return false;
    return true;
  };
  if (d->data.label.size() != n || d->data.split.size() != n || !csr_ok(d->data.graph.indptr, d->data.graph.indices, n) ||
      !csr_ok(d->data.feature_index.indptr, d->data.feature_index.indices, d->params.input_dim))
    return GCNB_E_BADARG;
  for (integer l : d->data.label)
    if (l >= (integer)d->params.output_dim) return GCNB_E_BADARG;
  *out = d.release();
  return 0;
}

static int fill_params(const gcnb_gcn_config *cfg, gcnb_gcn *g) {
  if (cfg->n_layers < 1 || (cfg->n_layers > 1 && !cfg->hidden_dims) || !cfg->dropouts) return GCNB_E_BADARG;
  g->params.num_nodes = (natural)cfg->num_nodes;
  g->params.input_dim = (natural)cfg->input_dim;
  g->params.output_dim = (natural)cfg->output_dim;
  g->params.n_layers = (natural)cfg->n_layers;
  g->params.hidden_dims.assign(cfg->hidden_dims, cfg->hidden_dims + (cfg->n_layers - 1));
  g->params.dropouts.assign(cfg->dropouts, cfg->dropouts + cfg->n_layers);
  g->params.epochs = cfg->epochs;
  g->params.early_stopping = cfg->early_stopping;
  g->adam.learning_rate = cfg->learning_rate;
  g->adam.beta1 = cfg->beta1;
  g->adam.beta2 = cfg->beta2;
  g->adam.eps = cfg->eps;
  g->adam.weight_decay = cfg->weight_decay;
  CudaParams::SEED = cfg->seed;
  return 0;
}

int gcnb_gcn_create(const gcnb_gcn_config *cfg, const gcnb_gcn_data *data, gcnb_gcn **out) {
  if (!cfg || !data || !out) return GCNB_E_BADARG;
  int sm = 0;
  const int dc = gcnb_device_check(&sm);
  if (dc) return dc;  // no GPU => error, never a CPU path
  auto g = std::make_unique<gcnb_gcn>();
  const int rc = fill_params(cfg, g.get());
  if (rc) return rc;
  const size_t n = (size_t)cfg->num_nodes;
  for (size_t i = 0; i < n; i++) {
    if (data->split[i] == 1) g->params.train_dim++;
    else if (data->split[i] == 2) g->params.val_dim++;
    else if (data->split[i] == 3) g->params.test_dim++;
  }
  GCNDataView v{};
  v.graph_indptr = data->graph_indptr;
  v.graph_indices = data->graph_indices;
  v.graph_nnz = (size_t)data->graph_nnz;
  v.graph_value = data->graph_value;
  v.feat_indptr = data->feat_indptr;
  v.feat_indices = data->feat_indices;
  v.feat_value = data->feat_value;
  v.feat_nnz = (size_t)data->feat_nnz;
  v.label = data->label;
  v.split = data->split;
  v.num_nodes = n;
  g->gcn = std::make_unique<GCN>(&g->params, &g->adam, v, cfg->quiet != 0);
  if (!cfg->reorder) g->gcn->set_reorder(false);
  g->masks.assign(g->params.n_layers, nullptr);
  *out = g.release();
  return 0;
}

int gcnb_gcn_create_partitioned(const gcnb_gcn_config *cfg, const gcnb_gcn_data *data, const gcnb_gcn_partition *part,
                                gcnb_gcn **out) {
  if (!cfg || !data || !part || !part->comm || !out || !data->graph_value) return GCNB_E_BADARG;
  int sm = 0;
  const int dc = gcnb_device_check(&sm);
  if (dc) return dc;
  auto g = std::make_unique<gcnb_gcn>();
  int rc = fill_params(cfg, g.get());
  if (rc) return rc;
  const size_t n = (size_t)cfg->num_nodes;
  // global split counts (CE normalisation, src/gcn.cu:214-219): local counts summed over the ranks
  uint32_t counts[3] = {0, 0, 0};
  for (size_t i = 0; i < n; i++)
    if (data->split[i] >= 1 && data->split[i] <= 3) counts[data->split[i] - 1]++;
  {
    uint32_t *d = nullptr;
    if (cudaMalloc((void **)&d, sizeof counts) != cudaSuccess) return (int)cudaGetLastError();
    cudaMemcpy(d, counts, sizeof counts, cudaMemcpyHostToDevice);
    rc = gcnb_comm_all_reduce_sum(part->comm, d, 3, 1, nullptr);
    if (!rc) rc = (int)cudaMemcpy(counts, d, sizeof counts, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (rc) return rc;
  }
  g->params.train_dim = counts[0];
  g->params.val_dim = counts[1];
  g->params.test_dim = counts[2];
  GCNDataView v{};
  v.graph_indptr = data->graph_indptr;
  v.graph_indices = data->graph_indices;
  v.graph_nnz = (size_t)data->graph_nnz;
  v.graph_value = data->graph_value;
  v.feat_indptr = data->feat_indptr;
  v.feat_indices = data->feat_indices;
  v.feat_value = data->feat_value;
  v.feat_nnz = (size_t)data->feat_nnz;
  v.label = data->label;
  v.split = data->split;
  v.num_nodes = n;
  GCNPartition p;
  p.comm = part->comm;
  p.n_global = (size_t)part->n_global;
  p.row_offset = (size_t)part->row_offset;
  p.block = (size_t)part->block;
  p.feat_elem_offset = (size_t)part->feat_elem_offset;
  p.feat_nnz_global = (size_t)part->feat_nnz_global;
  g->gcn = std::make_unique<GCN>(&g->params, &g->adam, v, p, cfg->quiet != 0);
  if (!cfg->reorder) g->gcn->set_reorder(false);
  g->masks.assign(g->params.n_layers, nullptr);
  *out = g.release();
  return 0;
}

int gcnb_gcn_create_from_dataset(const gcnb_gcn_config *cfg, const gcnb_dataset *d, gcnb_gcn **out) {
  if (!cfg || !d || !out) return GCNB_E_BADARG;
  gcnb_gcn_config c = *cfg;
  c.num_nodes = d->params.num_nodes;
  c.input_dim = d->params.input_dim;
  c.output_dim = d->params.output_dim;
  gcnb_gcn_data gd{};
  gd.graph_indptr = d->data.graph.indptr.data();
  gd.graph_indices = d->data.graph.indices.data();
  gd.graph_nnz = (int64_t)d->data.graph.indices.size();
  gd.graph_value = d->data.graph_value.data();
  gd.feat_indptr = d->data.feature_index.indptr.data();
  gd.feat_indices = d->data.feature_index.indices.data();
  gd.feat_value = d->data.feature_value.data();
  gd.feat_nnz = (int64_t)d->data.feature_index.indices.size();
  gd.label = d->data.label.data();
  gd.split = d->data.split.data();
  return gcnb_gcn_create(&c, &gd, out);
}

int gcnb_sweep_run(const gcnb_dataset *d, const gcnb_sweep_trial *trials, int64_t n_trials, int workers,
                   gcnb_sweep_result *results, double *wall_s) {
  if (!d || (n_trials > 0 && (!trials || !results)) || n_trials < 0) return GCNB_E_BADARG;
  for (int64_t i = 0; i < n_trials; i++)
    if (trials[i].n_layers < 1 || trials[i].n_layers > 8) return GCNB_E_BADARG;
  int sm = 0;
  const int dc = gcnb_device_check(&sm);
  if (dc) return dc;  // no GPU => error, never a CPU path
  const auto t0 = std::chrono::steady_clock::now();
  int device = 0;
  CHECK_CUDA_ERROR(cudaGetDevice(&device));
  const DevGCNData shared(d->data);  // one upload for every trial
  if (workers <= 0) workers = (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  workers = (int)std::min<int64_t>(workers, std::max<int64_t>(1, n_trials));
  std::atomic<int64_t> next{0};
  auto work = [&]() {
    CHECK_CUDA_ERROR(cudaSetDevice(device));
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_trials) return;
      const gcnb_sweep_trial &t = trials[i];
      GCNParams params = d->params;  // num_nodes, dims and split counts parsed from the data
      params.n_layers = (natural)t.n_layers;
      params.hidden_dims.assign(t.hidden_dims, t.hidden_dims + (t.n_layers - 1));
      params.dropouts.assign(t.dropouts, t.dropouts + t.n_layers);
      params.epochs = t.epochs;
      params.early_stopping = t.early_stopping;
      AdamParams adam;
      adam.learning_rate = t.learning_rate;
      adam.weight_decay = t.weight_decay;
      GCN gcn(&params, &adam, shared, t.seed, true);
      gcn.run();
      gcnb_sweep_result &r = results[i];
      r.last_val_accuracy = gcn.last_val_accuracy;
      r.last_val_loss = gcn.last_val_loss;
      r.last_train_loss = gcn.last_train_loss;
      r.avg_epoch_ms = gcn.avg_epoch_time;
      r.total_s = gcn.total_time;
      r.epochs_run = gcn.epochs_run();
    }
  };
  if (workers == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; w++) pool.emplace_back(work);
    for (auto &th : pool) th.join();
  }
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  if (wall_s) *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

int gcnb_gcn_halo_info(const gcnb_gcn *g, int64_t out[4]) {
  if (!g || !out) return GCNB_E_BADARG;
  g->gcn->halo_info(out);
  return 0;
}

int gcnb_gcn_destroy(gcnb_gcn *g) {
  delete g;
  return 0;
}

int gcnb_gcn_train_epoch(gcnb_gcn *g, float out[2]) {
  if (!g || !out) return GCNB_E_BADARG;
  auto r = g->gcn->train_epoch();
  out[0] = r.first;
  out[1] = r.second;
  return 0;
}

int gcnb_gcn_eval(gcnb_gcn *g, int split, float out[2]) {
  if (!g || !out || split < 1 || split > 3) return GCNB_E_BADARG;
  auto r = g->gcn->eval((natural)split);
  out[0] = r.first;
  out[1] = r.second;
  return 0;
}

int gcnb_gcn_run(gcnb_gcn *g, float out[4]) {
  if (!g) return GCNB_E_BADARG;
  reset_timer();
  g->gcn->run();
  if (out) {
    out[0] = g->gcn->avg_epoch_time;
    out[1] = g->gcn->total_time;
    out[2] = g->gcn->last_val_accuracy;
    out[3] = (float)g->gcn->epochs_run();
  }
  return 0;
}

int64_t gcnb_gcn_weight_size(const gcnb_gcn *g, int layer) {
  if (!g || layer < 0 || layer >= (int)g->gcn->n_layers()) return -1;
  return g->gcn->weight(layer)->size;
}
int gcnb_gcn_get_weight(const gcnb_gcn *g, int layer, float *dst) {
  if (!g || !dst || layer < 0 || layer >= (int)g->gcn->n_layers()) return GCNB_E_BADARG;
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  g->gcn->weight(layer)->dev_data.copy_to_host(dst);
  return 0;
}
int gcnb_gcn_set_weight(gcnb_gcn *g, int layer, const float *src) {
  if (!g || !src || layer < 0 || layer >= (int)g->gcn->n_layers()) return GCNB_E_BADARG;
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  g->gcn->weight(layer)->dev_data.copy_to_device(src);
  return 0;
}
int gcnb_gcn_get_weight_grad(const gcnb_gcn *g, int layer, float *dst) {
  if (!g || !dst || layer < 0 || layer >= (int)g->gcn->n_layers()) return GCNB_E_BADARG;
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  g->gcn->weight(layer)->dev_grad.copy_to_host(dst);
  return 0;
}
int gcnb_gcn_get_logits(const gcnb_gcn *g, float *dst) {
  if (!g || !dst) return GCNB_E_BADARG;
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  g->gcn->logits()->dev_data.copy_to_host(dst);
  return 0;
}
int gcnb_gcn_set_mask(gcnb_gcn *g, int site, const uint8_t *host_mask) {
  if (!g || site < 0 || site >= (int)g->gcn->n_layers()) return GCNB_E_BADARG;
  g->masks[site] = host_mask;
  g->gcn->set_external_masks(g->masks);  // uploads now; the host buffer is not referenced afterwards
  return 0;
}
int64_t gcnb_gcn_launches_per_epoch(const gcnb_gcn *g) { return g ? (int64_t)g->gcn->launches_per_epoch() : -1; }
int gcnb_gcn_set_cuda_graph(gcnb_gcn *g, int on) {
  if (!g) return GCNB_E_BADARG;
  g->gcn->set_use_cuda_graph(on != 0);
  return 0;
}
int gcnb_gcn_finish_setup(gcnb_gcn *g) {
  if (!g) return GCNB_E_BADARG;
  g->gcn->finish_setup();
  return 0;
}

// host-only access to the Philox consumption bookkeeping (Variable::rng_history), for tests: no CUDA call
int gcnb_rng_history_reset(void) {
  Variable::initialize_random();
  return 0;
}
int gcnb_rng_history_consume(uint64_t n_elements) {
  Variable::rng_consume((size_t)n_elements);
  return 0;
}
int gcnb_rng_history_descriptor(gcnb_rng_t *out) {
  if (!out) return GCNB_E_BADARG;
  *out = Variable::rng_descriptor();
  return 0;
}
int gcnb_gcn_uses_cuda_graph(const gcnb_gcn *g) { return g ? (int)g->gcn->uses_cuda_graph() : -1; }
int gcnb_gcn_graph_staged(const gcnb_gcn *g) { return g ? (int)g->gcn->graph_staged() : -1; }
int gcnb_gcn_graph_bittile(const gcnb_gcn *g) { return g ? (int)g->gcn->graph_bittile() : -1; }
int gcnb_gcn_path_info(const gcnb_gcn *g, int out[8]) {
  if (!g || !out) return GCNB_E_BADARG;
  g->gcn->path_info(out);
  return 0;
}
int64_t gcnb_gcn_launches_total(const gcnb_gcn *g) { return g ? (int64_t)g->gcn->launches_total() : -1; }
int gcnb_gcn_timed_epochs(gcnb_gcn *g, int n_epochs, int with_eval, int time_graphsum, float out[4]) {
  if (!g || !out || n_epochs < 0) return GCNB_E_BADARG;
  g->gcn->set_time_graphsum(time_graphsum != 0);
  const size_t l0 = g->gcn->launches_total();
  out[0] = g->gcn->timed_epochs((natural)n_epochs, with_eval != 0);
  double ms = 0;
  size_t calls = 0;
  g->gcn->graphsum_timing(&ms, &calls);
  out[1] = (float)ms;
  out[2] = (float)calls;
  out[3] = (float)(g->gcn->launches_total() - l0);
  g->last_exchange_ms = g->gcn->graphsum_exchange_ms();
  g->gcn->set_time_graphsum(false);
  return 0;
}
double gcnb_gcn_graphsum_exchange_ms(const gcnb_gcn *g) { return g ? g->last_exchange_ms : 0.0; }

}  // extern "C"
