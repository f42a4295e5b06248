/*
 * gcnb_engine.h -- engine-level C ABI of libgcn_b200.so: the reference's *driver* surface (Parser, GCN ctor,
 * GCN::run / train_epoch / eval) behind plain C, taking HOST buffers.  This is what a foreign-language caller (or
 * bench.py's end-to-end leg) binds; include/gcnb.h is the kernel-level ABI underneath it.
 *
 *   Parser(GCNParams*, GCNData*, name).parse()         include/parser.h:12-19, src/parser.cpp:189-209
 *   GCN(GCNParams const*, AdamParams const*, GCNData const*)   include/gcn.cuh:114-121, src/gcn.cu:146-177
 *   GCN::run()                                          src/gcn.cu:347-436
 *   GCN::train_epoch() / eval(split)  (private there)   src/gcn.cu:293-343
 *
 * All functions return 0 on success or a cudaError_t / GCNB_E_* code; unrecoverable CUDA failures inside the C++
 * classes follow the reference's convention (message on stderr, exit).  Host pointers are read during the call
 * only (the engine uploads them; pinned memory makes that upload run at full PCIe speed).
 */
#ifndef GCNB_ENGINE_H
#define GCNB_ENGINE_H
#include "gcnb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int64_t num_nodes, input_dim, output_dim;
  int32_t n_layers;
  const uint32_t *hidden_dims; /* n_layers - 1 entries */
  const float *dropouts;       /* n_layers entries     */
  uint32_t epochs, early_stopping;
  float learning_rate, beta1, beta2, eps, weight_decay;
  uint32_t seed;    /* CudaParams::SEED */
  int32_t quiet;    /* runtime form of -DNO_OUTPUT */
  int32_t reorder;  /* 1: allow (A*a)*W when in_dim < out_dim; 0: always the module chain's A*(a*W) */
} gcnb_gcn_config;

typedef struct {
  const uint32_t *graph_indptr, *graph_indices; /* [num_nodes+1], [graph_nnz]; row i starts with i (self) */
  int64_t graph_nnz;
  const float *graph_value;                     /* [graph_nnz] or NULL => 1/sqrtf(deg*deg) computed */
  const uint32_t *feat_indptr, *feat_indices;   /* [num_nodes+1], [feat_nnz] */
  const float *feat_value;                      /* [feat_nnz] */
  int64_t feat_nnz;
  const int32_t *label;                         /* [num_nodes], -1 = unlabelled */
  const uint32_t *split;                        /* [num_nodes], 1 train / 2 val / 3 test */
} gcnb_gcn_data;

/* ---- datasets (Parser) ---- */
typedef struct gcnb_dataset gcnb_dataset;
/* reads <root>/data/<name>.{graph,split,svmlight}; NULL in *out + nonzero return if a file is missing */
GCNB_API int gcnb_dataset_parse(const char *root, const char *name, int no_feature, gcnb_dataset **out);
/* dims: num_nodes, graph_nnz, feat_rows, feat_nnz, input_dim, output_dim, n_split, train_dim, val_dim, test_dim */
GCNB_API int gcnb_dataset_dims(const gcnb_dataset *d, int64_t dims[10]);
/* which: 0 graph_indptr 1 graph_indices 2 feat_indptr 3 feat_indices 4 feat_value 5 label 6 split 7 graph_value */
GCNB_API int gcnb_dataset_copy(const gcnb_dataset *d, int which, void *dst);
GCNB_API int gcnb_dataset_free(gcnb_dataset *d);
/* binary container of a parsed dataset (a 128-byte header + the eight arrays of gcnb_dataset_copy, 64-byte aligned):
 * written once, loading it back (mmap + copy) is bit-identical to parsing the text files again and takes milliseconds */
GCNB_API int gcnb_dataset_save(const gcnb_dataset *d, const char *path);
GCNB_API int gcnb_dataset_load(const char *path, gcnb_dataset **out);

/* Row partition of one rank (multi-GPU).  The rank owns the contiguous global rows
 * [row_offset, row_offset + cfg.num_nodes) of the adjacency, of the features, labels, split and of every activation;
 * `block` (a multiple of 4, the same on every rank, >= every rank's row count, world * block >= n_global) is the slab
 * size of the all-gather.  Graph column ids stay GLOBAL; graph_value must be given (it needs global degrees).
 * cfg.num_nodes is the LOCAL row count; train/val/test sample counts are global and computed by all-reduce. */
typedef struct {
  gcnb_comm *comm;         /* from gcnb_comm_create; borrowed, must outlive the model */
  int64_t n_global;        /* rows of the whole graph */
  int64_t row_offset;      /* first global row of this rank */
  int64_t block;           /* all-gather slab rows */
  int64_t feat_elem_offset;/* global position of this rank's first feature value (sum of earlier ranks' feat_nnz) */
  int64_t feat_nnz_global; /* feature values of the whole dataset */
} gcnb_gcn_partition;

/* ---- model (GCN) ---- */
typedef struct gcnb_gcn gcnb_gcn;
GCNB_API int gcnb_gcn_create(const gcnb_gcn_config *cfg, const gcnb_gcn_data *data, gcnb_gcn **out);
/* same on one row block; every rank of the communicator must call it (collectives inside) */
GCNB_API int gcnb_gcn_create_partitioned(const gcnb_gcn_config *cfg, const gcnb_gcn_data *data,
                                         const gcnb_gcn_partition *part, gcnb_gcn **out);
GCNB_API int gcnb_gcn_create_from_dataset(const gcnb_gcn_config *cfg, const gcnb_dataset *d, gcnb_gcn **out);
GCNB_API int gcnb_gcn_destroy(gcnb_gcn *g);
GCNB_API int gcnb_gcn_train_epoch(gcnb_gcn *g, float out_loss_acc[2]);
GCNB_API int gcnb_gcn_eval(gcnb_gcn *g, int split, float out_loss_acc[2]);
/* GCN::run(): out[0] = avg_epoch_time (ms, reference's TMR_TRAIN/(epochs+1)), out[1] = total_time (s),
 * out[2] = last_val_accuracy, out[3] = epochs actually run */
GCNB_API int gcnb_gcn_run(gcnb_gcn *g, float out[4]);
GCNB_API int64_t gcnb_gcn_weight_size(const gcnb_gcn *g, int layer);
GCNB_API int gcnb_gcn_get_weight(const gcnb_gcn *g, int layer, float *host_dst);
GCNB_API int gcnb_gcn_set_weight(gcnb_gcn *g, int layer, const float *host_src);
GCNB_API int gcnb_gcn_get_weight_grad(const gcnb_gcn *g, int layer, float *host_dst);
GCNB_API int gcnb_gcn_get_logits(const gcnb_gcn *g, float *host_dst); /* [num_nodes x output_dim], CE-shifted */
/* injected keep-masks (1 byte/element) for the following training passes; site 0 = input features
 * ([feat_nnz]), site l = hidden layer l-1 ([num_nodes x hidden_dims[l-1]]); NULL = back to Philox */
GCNB_API int gcnb_gcn_set_mask(gcnb_gcn *g, int site, const uint8_t *host_mask);
GCNB_API int64_t gcnb_gcn_launches_per_epoch(const gcnb_gcn *g);
/* 1 if GraphSum at feature width 16 uses the window-staged kernels (graph with column locality), else 0 */
GCNB_API int gcnb_gcn_graph_staged(const gcnb_gcn *g);
/* 1 if GraphSum at width 16 runs the tcgen05 bit-tile path (parallel-gcn_b200/csrc/spmm_bittile.cu; the default when a
 * quarter of the adjacency's entries sit in dense blocks, GCNB_BITTILE=0 turns it off) */
GCNB_API int gcnb_gcn_graph_bittile(const gcnb_gcn *g);
/* which fast paths are active right now: out = {window-staged GraphSum, bit-tile GraphSum, dense-feature first layer
 * (tensor-core X W0 / X^T dH kernels), evaluation through the propagated features A_hat X, CUDA-graph replay usable,
 * background set-up still pending (gcnb_gcn_finish_setup), exact-split tcgen05 GEMM packed, 1 = row-partitioned /
 * 2 = the bit tiles were built from the graph renumbered community by community (transparent: GCNB_RENUMBER=0 disables)} */
GCNB_API int gcnb_gcn_path_info(const gcnb_gcn *g, int out[8]);
GCNB_API int64_t gcnb_gcn_launches_total(const gcnb_gcn *g);
/* CUDA-graph replay of the training epoch and of the evaluation passes (small datasets are launch-bound).  Default: on
 * when graph + feature entries <= 8 Mi (GCNB_CUDA_GRAPH=0/1 overrides); never used by a partitioned model, with injected
 * masks, or while GraphSum launches are being timed.  Results are bit-identical to eager launches.
 * gcnb_gcn_uses_cuda_graph: 1 if the next passes may be replayed. */
/* Large single-GPU models (graph + feature entries > 8 Mi; GCNB_ASYNC_STAGE=0 builds synchronously instead):
 * gcnb_gcn_create returns as soon as the dataset is on the device and the first epochs run on the generic GraphSum kernel
 * while the static GraphSum representation (bit tiles, or window staging when the graph has no dense blocks) is built and
 * uploaded by a helper thread; it is attached before training epoch GCNB_STAGE_SWITCH_EPOCH (default 128) -- a fixed point, so results do not
 * depend on timing -- or by this call (which waits for the helper if it is still busy).  No-op otherwise. */
GCNB_API int gcnb_gcn_finish_setup(gcnb_gcn *g);
GCNB_API int gcnb_gcn_set_cuda_graph(gcnb_gcn *g, int on);
GCNB_API int gcnb_gcn_uses_cuda_graph(const gcnb_gcn *g);
/* measurement hook (bench.py): n_epochs x {train_epoch [+ eval(2)]} bracketed by CUDA events on the engine's stream.
 * out[0] = total ms, out[1] = summed ms of the GraphSum SpMM launches (event pair per launch, if time_graphsum),
 * out[2] = number of GraphSum launches, out[3] = CUDA kernels launched in the region */
GCNB_API int gcnb_gcn_timed_epochs(gcnb_gcn *g, int n_epochs, int with_eval, int time_graphsum, float out[4]);
/* row-partitioned models, after a gcnb_gcn_timed_epochs call with time_graphsum: the part of the summed GraphSum time
 * (out[1]) that passed between the start of a call and the moment every peer's slab had landed -- the slab exchange, and
 * with window staging the own-slab windows that overlap it */
GCNB_API double gcnb_gcn_graphsum_exchange_ms(const gcnb_gcn *g);
/* row-partitioned models: {halo exchange active (gcnb_comm_halo_setup), rows this rank ships per exchange, rows a full
 * slab push would ship, rows of other ranks' blocks this rank references} */
GCNB_API int gcnb_gcn_halo_info(const gcnb_gcn *g, int64_t out[4]);

/* ---- tuning sweep as a throughput workload (test/tuning_accuracy.cpp:56-196, test/tuning_cuda.cpp of the reference) ----
 * The reference tunes by constructing one model after the other (20 seeds x every parameter combination, 1000 epochs with
 * early stopping each): every constructor uploads the dataset again and a cora-sized epoch leaves most of the GPU idle.
 * gcnb_sweep_run uploads the dataset ONCE and runs the trials on `workers` host threads, each model on its own streams with
 * its own Philox context and (small datasets) its own CUDA-graph replays, so the epochs of several models overlap on the
 * device.  A trial is exactly `CudaParams::SEED = seed; GCN gcn{params, adam_params, data}; gcn.run();` in the quiet build:
 * results do not depend on `workers` or on which trials run side by side (bit-identical to running them one by one).
 * workers <= 0: min(8, hardware threads).  wall_s: wall-clock seconds of the whole sweep (upload included). */
typedef struct {
  int32_t n_layers;          /* 1..8 */
  uint32_t hidden_dims[7];   /* n_layers - 1 entries used */
  float dropouts[8];         /* n_layers entries used */
  uint32_t epochs, early_stopping;
  float learning_rate, weight_decay; /* beta1 0.9, beta2 0.999, eps 1e-8 as the reference's AdamParams defaults */
  uint32_t seed;
} gcnb_sweep_trial;
typedef struct {
  float last_val_accuracy, last_val_loss, last_train_loss;
  float avg_epoch_ms;  /* the reference's TMR_TRAIN average: wall time of the trial / (epochs run + 1) */
  float total_s;
  uint32_t epochs_run;
} gcnb_sweep_result;
GCNB_API int gcnb_sweep_run(const gcnb_dataset *d, const gcnb_sweep_trial *trials, int64_t n_trials, int workers,
                            gcnb_sweep_result *results, double *wall_s);

/* Host-only view of the stateless-Philox bookkeeping behind Variable (src/variable.cu:13-26 keeps a state array; here the
 * consumption history IS the state): reset = Variable::initialize_random(), consume = "an RNG op over n_elements ran"
 * (64-bit: GLOBAL element counts of row-partitioned models), descriptor = what the next RNG kernel receives.  No CUDA. */
GCNB_API int gcnb_rng_history_reset(void);
GCNB_API int gcnb_rng_history_consume(uint64_t n_elements);
GCNB_API int gcnb_rng_history_descriptor(gcnb_rng_t *out);

/* ---- synthetic workloads (BASELINE.json configs 3-5; deterministic in seed, independent of thread count) ----
 * Symmetric simple graph with `n_blocks` contiguous planted communities: undirected edges are drawn with endpoint
 * probability proportional to a lognormal(sigma) weight (expected-degree model, weights clipped so that no expected
 * degree exceeds max_deg); with probability `intra` the second endpoint comes from the first one's community.
 * Output is the parser's CSR convention (row i = [i, sorted neighbours]); arrays are malloc'ed, free with
 * gcnb_host_free.  n_undirected_edges is met exactly. */
GCNB_API int gcnb_synth_graph(int64_t n, int64_t n_undirected_edges, int n_blocks, double intra, double sigma,
                              int64_t max_deg, uint64_t seed, uint32_t **indptr_out, uint32_t **indices_out,
                              int64_t *nnz_out);
GCNB_API void gcnb_host_free(void *p);
/* dense feature CSR (every row holds columns 0..f-1, as svmlight-Reddit parses), values ~ N(0,1) */
GCNB_API int gcnb_synth_dense_features(int64_t n, int f, uint64_t seed, uint32_t *indptr, uint32_t *indices,
                                       float *values);
/* uniform labels in [0, classes), split 1/2/3 with the given train/val fractions (rest = test) */
GCNB_API int gcnb_synth_labels(int64_t n, int classes, double frac_train, double frac_val, uint64_t seed,
                               int32_t *label, uint32_t *split);

/* Row-local symmetric generator for the scale-out workloads (BASELINE.json config 5): rows [row0, row0 + rows) of a
 * symmetric simple graph on n nodes whose edge {i, j} is a pure function of (seed, min, max) -- every rank of a
 * row-partitioned job generates only its block, and the blocks of all ranks form one consistent symmetric graph.
 * Communities are contiguous blocks of block_size nodes (every pair inside is a candidate, expected mean_intra
 * neighbours); across communities node i's candidates are its images under n_reflect fixed reflections (c_k - i) mod n
 * (expected mean_inter neighbours); degrees are skewed by lognormal(sigma) weights.  Column ids are global; row i =
 * [i, sorted neighbours] (the parser's convention).  Arrays are malloc'ed: gcnb_host_free. */
GCNB_API int gcnb_synth_sym_rows(int64_t n, int64_t row0, int64_t rows, int64_t block_size, double mean_intra,
                                 double mean_inter, int n_reflect, double sigma, uint64_t seed, uint32_t **indptr_out,
                                 uint32_t **indices_out, int64_t *nnz_out);
/* same with LOCAL inter-community edges (inter_window > 0; 0 = the reflections above): across communities node i's candidates
 * are i +- d_k for n_reflect / 2 fixed shifts d_k in [block_size, inter_window], so a row block references only rows within
 * inter_window of its borders -- the structure the halo exchange (gcnb_comm_halo_setup) is for */
GCNB_API int gcnb_synth_sym_rows_local(int64_t n, int64_t row0, int64_t rows, int64_t block_size, double mean_intra,
                                       double mean_inter, int n_reflect, int64_t inter_window, double sigma, uint64_t seed,
                                       uint32_t **indptr_out, uint32_t **indices_out, int64_t *nnz_out);
/* graph_value of a row block from the GLOBAL degree array (Parser::calculateGraphValues arithmetic, src/parser.cpp:164-181) */
GCNB_API int gcnb_synth_graph_values(const uint32_t *indptr, const uint32_t *indices, int64_t rows, int64_t row0,
                                     const uint32_t *deg_global, float *out);
/* dense feature CSR with cheap uniform(-sqrt 3, sqrt 3) values; elem_offset = global position of the block's first value */
GCNB_API int gcnb_synth_dense_features_uniform(int64_t n, int f, uint64_t seed, uint64_t elem_offset, uint32_t *indptr,
                                               uint32_t *indices, float *values);

/* ---- locality reordering (host side, multi-threaded, deterministic) ----
 * Communities by synchronous label propagation (ties broken by a hash, at most max_sweeps sweeps, 0 = 8), nodes
 * renumbered community by community: new_of_old[i] = new id of node i.  The window-staged GraphSum needs the nodes of a
 * community next to each other; datasets rarely come that way.  The renumbered dataset (gcnb_permute_csr for the
 * graph, gcnb_permute_rows for features / labels / split) is an ordinary dataset for everything downstream;
 * gcnb_unpermute_rows maps per-node outputs (logits) back to the original numbering. */
GCNB_API int gcnb_reorder_communities(int64_t n, const uint32_t *indptr, const uint32_t *indices, int max_sweeps,
                                      uint64_t seed, uint32_t *new_of_old, int64_t *n_communities);
/* Balanced, community-aligned row partition for `world` ranks: equal CSR-entry counts per rank, rank
 * boundaries on community borders where one lies within `tolerance` (fraction of a rank's share; 0 = 0.1) of the balanced
 * cut.  new_of_old[i] in [0, world * block): rank r owns the ids [r * block, r * block + rows_out[r]), block = a multiple of
 * 4 (the engine's slab size); the remaining ids are padding.  stats: {communities, cut entries, cut entries of the plain
 * equal-row-block partition of the given numbering, largest entry count of a rank}. */
GCNB_API int gcnb_partition_communities(int64_t n, const uint32_t *indptr, const uint32_t *indices, int world, int max_sweeps,
                                        uint64_t seed, double tolerance, uint32_t *new_of_old, int64_t *block_out,
                                        int64_t *rows_out, int64_t stats[4]);
GCNB_API int gcnb_permute_csr(int64_t n, const uint32_t *indptr, const uint32_t *indices, const uint32_t *new_of_old,
                              uint32_t *out_indptr, uint32_t *out_indices);
GCNB_API int gcnb_permute_rows(int64_t n, int64_t row_bytes, const uint32_t *new_of_old, const void *in, void *out);
GCNB_API int gcnb_unpermute_rows(int64_t n, int64_t row_bytes, const uint32_t *new_of_old, const void *in, void *out);

#ifdef __cplusplus
}
#endif
#endif /* GCNB_ENGINE_H */
