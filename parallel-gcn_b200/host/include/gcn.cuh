// gcn.cuh -- model driver, mirror of include/gcn.cuh:23-122 (GCNSmartObjects, GCNParams, GCNData, DevGCNData, GCN
// with the same constructor, run() and public result fields).  The private part is a B200-first redesign:
//   * one stream, no per-module event choreography; the per-epoch launch sequence is fixed, so small datasets can
//     be replayed from a CUDA graph (launch-bound regime);
//   * the feature matrix is never overwritten: input dropout writes a second buffer (no set_input restore copy,
//     src/gcn.cu:181-200), and an all-columns-present feature CSR (Reddit) is routed to the dense kernels;
//   * per layer the cheaper association is used: (A_hat * a) * W when in_dim < out_dim (GraphSum runs at the
//     narrower width), A_hat * (a * W) otherwise -- same math as the module chain, different rounding order;
//   * ReLU+Dropout fused, loss + accuracy fused, multi-tensor Adam, fixed-order reductions.
#ifndef GCN_CUH
#define GCN_CUH
#include <cuda_runtime.h>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "../include/module.cuh"
#include "../include/optim.cuh"
#include "../include/reduction.cuh"
#include "../include/shared_ptr.cuh"
#include "../include/smart_object.cuh"
#include "../include/sparse.cuh"
#include "../include/utils.cuh"
#include "../include/variable.cuh"

using std::shared_ptr;
using std::unique_ptr;

class GCNSmartObjects {
 public:
  smart_stream forward_training_stream;
  smart_stream forward_evaluation_stream;
  std::vector<smart_stream> backward_streams;  // 2
  smart_event start_backward;
  smart_event start_set_input;
  std::vector<smart_event> start_matmul_backward;  // L - 1
  std::vector<smart_event> start_matmul_forward;   // L
  explicit GCNSmartObjects(const natural n_layers);
};

struct GCNParams {
  natural num_nodes, input_dim, output_dim;
  std::vector<natural> hidden_dims = {16};
  std::vector<real> dropouts = {0.5, 0.5};
  natural epochs{100}, early_stopping{0};
  natural train_dim{0}, val_dim{0}, test_dim{0};
  natural n_layers{2};
  void print_info() const;
};

struct GCNData {
  SparseIndex feature_index, graph;
  std::vector<natural> split;
  std::vector<integer> label;
  std::vector<real> graph_value;
  std::vector<real> feature_value;
};

// extension: non-owning view of the same data as raw host arrays (C ABI / synthetic graphs)
struct GCNDataView {
  const natural *graph_indptr, *graph_indices;
  size_t graph_nnz;
  const real *graph_value;  // may be nullptr: computed as 1./sqrtf(deg_src*deg_dst)
  const natural *feat_indptr, *feat_indices;
  const real *feat_value;
  size_t feat_nnz;
  const integer *label;
  const natural *split;
  size_t num_nodes;
};

class DevGCNData {
 public:
  DevSparseIndex dev_graph_index;    // adjacency matrix
  DevSparseIndex dev_feature_index;  // feature
  dev_shared_ptr<real> dev_feature_value;
  dev_shared_ptr<real> dev_graph_value;
  dev_shared_ptr<natural> dev_split;
  dev_shared_ptr<integer> dev_label;
  natural label_size;
  DevGCNData(const GCNData &gcn_data);
  DevGCNData(const GCNDataView &view);
};

struct GCNEngineState;  // plans, workspaces, captured graphs (src/gcn.cpp)

// extension: row block of one rank in the multi-GPU engine (see gcnb_gcn_partition in include/gcnb_engine.h)
struct gcnb_comm;
struct GCNPartition {
  gcnb_comm *comm = nullptr;
  size_t n_global = 0, row_offset = 0, block = 0, feat_elem_offset = 0, feat_nnz_global = 0;
};

class GCN {
  GCNSmartObjects smart_objects;
  natural L;
  const GCNData *data;
  DevGCNData dev_data;
  std::vector<shared_ptr<Variable>> variables;
  shared_ptr<Variable> input, output;
  std::vector<shared_ptr<Variable>> weights;
  std::vector<bool> decays;
  Adam optimizer;
  dev_shared_ptr<integer> dev_truth;
  std::string variables_info;
  shared_ptr<GCNEngineState> st;

  void set_truth(const natural current_split, cudaStream_t stream) const;
  void forward_pass(bool training, natural split, cudaStream_t stream);
  void train_body(cudaStream_t stream);
  void backward_pass(cudaStream_t stream);
  std::pair<real, real> finalize(cudaStream_t stream, int slot) const;
  std::pair<real, real> read_result(int slot) const;
  // returns false when sync == false and the passes were only enqueued (graph replays): results via read_result after a sync
  bool train_and_eval(natural split, std::pair<real, real> &train, std::pair<real, real> &val, bool sync = true);
  void print_variable_info() const;
  void run_concurrent();
  void init(bool quiet, const natural *h_graph_indptr = nullptr, const natural *h_graph_indices = nullptr,
            const GCNPartition *part = nullptr, const natural *seed = nullptr);

 public:
  real avg_epoch_time;
  real total_time;
  real last_val_accuracy;
  const GCNParams *params;
  const AdamParams *adam_params;
  // The reference silences its output at compile time (-DNO_OUTPUT, Makefile:40-63).  The library is compiled once,
  // so the flag is sampled in the including translation unit and passed down.
#ifdef NO_OUTPUT
  GCN(GCNParams const *params_, AdamParams const *adam_params_, GCNData const *data_)
      : GCN(params_, adam_params_, data_, true) {}
#else
  GCN(GCNParams const *params_, AdamParams const *adam_params_, GCNData const *data_)
      : GCN(params_, adam_params_, data_, false) {}
#endif
  GCN(GCNParams const *params_, AdamParams const *adam_params_, GCNData const *data_, bool quiet);
  GCN(GCNParams const *params_, AdamParams const *adam_params_, const GCNDataView &view, bool quiet);
  // row-partitioned rank: params_->num_nodes = LOCAL rows, train/val/test_dim = GLOBAL counts, view = the row block
  GCN(GCNParams const *params_, AdamParams const *adam_params_, const GCNDataView &view, const GCNPartition &part,
      bool quiet);
  // several models over ONE device-resident dataset (the tuning sweeps of test/tuning_accuracy.cpp construct 20 models per
  // parameter combination and upload the dataset 20 times): the buffers of `shared` are referenced, not copied; the seed
  // is the model's own (CudaParams::SEED is process-wide, concurrent constructors would race on it)
  GCN(GCNParams const *params_, AdamParams const *adam_params_, const DevGCNData &shared, natural seed, bool quiet);
  ~GCN();
  void run();
  void set_concurrent(bool on);  // several models run at once from several host threads: run() keeps its timers private

  // ---- extensions (not in the reference's public surface; used by the C ABI engine, tests and bench) ----
  std::pair<real, real> train_epoch();                       // reference: private, src/gcn.cu:307-343
  std::pair<real, real> eval(const natural current_split);   // reference: private, src/gcn.cu:293-303
  natural n_layers() const { return L; }
  const shared_ptr<Variable> &weight(natural l) const { return weights[l]; }
  const shared_ptr<Variable> &logits() const { return output; }
  // injected randomness for parity against the reference CPU implementation: keep-masks (1 byte per element)
  // for the next training pass, one per dropout site (0 = input, l = hidden layer l-1); nullptr = draw with Philox
  void set_external_masks(const std::vector<const unsigned char *> &host_masks);
  void set_quiet(bool q);
  void set_reorder(bool on);  // allow the (A*a)*W association (default on)
  // replay captured epochs as CUDA graphs (default: on for small datasets, which are launch-bound); results are
  // bit-identical to eager launches
  void set_use_cuda_graph(bool on);
  bool uses_cuda_graph() const;
  size_t launches_per_epoch() const;
  // GCNB_ASYNC_STAGE=1 builds the staged GraphSum representation in the background and attaches it before a fixed training
  // epoch (GCNB_STAGE_SWITCH_EPOCH, default 128); finish_setup() attaches it now (waits for the helper if need be)
  void finish_setup();
  bool graph_bittile() const;  // GraphSum at width 16 runs the tcgen05 bit-tile path (default when the graph has dense blocks)
  // which fast paths are active: {window-staged GraphSum, bit-tile GraphSum, dense-feature first layer, evaluation through
  // the propagated features A_hat X, CUDA-graph replay, background set-up still pending, exact-split tcgen05 GEMM,
  // 1 = partitioned / 2 = graph renumbered for the bit tiles}
  void path_info(int out[8]) const;
  bool graph_staged() const;  // GraphSum at widths 16 / >= 64 runs the window-staged kernels (csrc/spmm_stage.cu)
  size_t launches_total() const;
  void set_time_graphsum(bool on);                        // event pair around every GraphSum launch
  void graphsum_timing(double *ms_total, size_t *calls) const;
  void halo_info(int64_t out[4]) const;  // partitioned: {halo exchange active, rows sent per exchange, rows of a full push, rows needed}
  double graphsum_exchange_ms() const;  // partitioned: summed time from the start of a GraphSum call until the peers' slabs have landed
  float timed_epochs(natural n_epochs, bool with_eval);   // ms between CUDA events on the engine stream
  natural epochs_run() const;
  real last_val_loss = 0, last_train_loss = 0;  // of the last epoch run() completed
};
#endif
