// module_chain.cpp -- test program (not product code): composes the product's C++ module classes BY HAND, the way the
// reference's GCN constructor composes its own (/root/reference/src/gcn.cu:47-142: Dropout -> SparseMatmul -> GraphSum ->
// ReLU -> Dropout -> Matmul -> GraphSum -> CrossEntropyLoss over shared Variables), runs one training forward + backward
// pass and one evaluation forward pass as GCN::train_epoch / GCN::eval do (src/gcn.cu:293-343, :441-470), and dumps the
// tensors for tests/test_dropin_gpu.py to compare with the oracle's module outputs.
//
//   module_chain <dataset name> <output dir> [seed]        (run from a directory that holds data/<name>.graph|.split|.svmlight)
//
// Built on the spot by tests/native/build.py against parallel-gcn_b200/host/include + libgcn_b200.so only.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "gcn.cuh"
#include "module.cuh"
#include "optim.cuh"
#include "parser.h"

namespace {

template <class T>
void dump(const std::string &dir, const char *name, const T *dev, size_t n) {
  std::vector<T> h(n);
  CHECK_CUDA_ERROR(cudaMemcpy(h.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost));
  std::ofstream f(dir + "/" + name, std::ios::binary);
  f.write(reinterpret_cast<const char *>(h.data()), (std::streamsize)(n * sizeof(T)));
}

void set_truth(const GCNData &data, natural split, const dev_shared_ptr<integer> &dev_truth) {
  std::vector<integer> t(data.label.size());
  for (size_t i = 0; i < t.size(); i++) t[i] = data.split[i] == split ? data.label[i] : -1;  // set_truth_kernel, src/gcn.cu:204-226
  dev_truth.copy_to_device(t.data());
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: module_chain <dataset> <out dir> [seed]\n");
    return 2;
  }
  const std::string name = argv[1], out = argv[2];
  if (argc > 3) CudaParams::SEED = (natural)strtoul(argv[3], nullptr, 10);
  GCNParams params;
  GCNData data;
  Parser parser(&params, &data, name, false, true);
  if (!parser.parse()) {
    fprintf(stderr, "cannot read %s\n", name.c_str());
    return 3;
  }
  const natural N = params.num_nodes, F = params.input_dim, H = params.hidden_dims.front(), C = params.output_dim;
  DevGCNData dev(data);

  smart_stream fwd, bwd0, bwd1;
  smart_event weights_ready0, weights_ready1, input_free, unused, dgrad_ready, loss_ready;
  dev_shared_ptr<integer> dev_truth(N);
  pinned_host_ptr<real> loss(1);

  // variables in the reference's order: input, {var1, weight, var2} per layer
  auto input = std::make_shared<Variable>(data.feature_index.indices.size(), false, true);
  auto l1_var1 = std::make_shared<Variable>(N * H);
  auto l1_weight = std::make_shared<Variable>(F * H, true, true, F, H);
  auto l1_var2 = std::make_shared<Variable>(N * H, true, true);
  auto l2_var1 = std::make_shared<Variable>(N * C);
  auto l2_weight = std::make_shared<Variable>(H * C, true, true, H, C);
  auto output = std::make_shared<Variable>(N * C);

  std::vector<std::unique_ptr<Module>> modules;
  modules.push_back(std::make_unique<Dropout>(input, params.dropouts.front()));
  modules.push_back(std::make_unique<SparseMatmul>(input, l1_weight, l1_var1, &dev.dev_feature_index, N, F, H, weights_ready0, input_free));
  modules.push_back(std::make_unique<GraphSum>(l1_var1, l1_var2, &dev.dev_graph_index, dev.dev_graph_value, H, false, unused));
  modules.push_back(std::make_unique<ReLU>(l1_var2));
  modules.push_back(std::make_unique<Dropout>(l1_var2, params.dropouts.back()));
  modules.push_back(std::make_unique<Matmul>(l1_var2, l2_weight, l2_var1, N, H, C, weights_ready1, dgrad_ready, bwd1));
  modules.push_back(std::make_unique<GraphSum>(l2_var1, output, &dev.dev_graph_index, dev.dev_graph_value, C, true, dgrad_ready));
  modules.push_back(std::make_unique<CrossEntropyLoss>(output, dev_truth, loss, C, loss_ready));

  Variable::initialize_random();
  l1_weight->glorot();
  l2_weight->glorot();
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  dump(out, "w0_init.f32", l1_weight->dev_data.get(), l1_weight->size);
  dump(out, "w1_init.f32", l2_weight->dev_data.get(), l2_weight->size);

  // ---- training pass: set_input, set_truth(train), forward(true), backward in reverse order
  CHECK_CUDA_ERROR(cudaMemcpy(input->dev_data.get(), dev.dev_feature_value.get(), input->size * sizeof(real), cudaMemcpyDeviceToDevice));
  set_truth(data, 1, dev_truth);
  modules.back()->set_num_samples(params.train_dim);
  for (const auto &m : modules) m->forward(true, fwd);
  CHECK_CUDA_ERROR(cudaStreamSynchronize(fwd.get()));
  const real train_loss_sum = *loss;
  dump(out, "train_logits.f32", output->dev_data.get(), output->size);  // shifted in place for labelled rows (CE side effect)
  dump(out, "train_hidden.f32", l1_var2->dev_data.get(), l1_var2->size);
  for (auto it = modules.rbegin(); it != modules.rend(); ++it) (*it)->backward(bwd0);
  CHECK_CUDA_ERROR(cudaDeviceSynchronize());
  dump(out, "dw0.f32", l1_weight->dev_grad.get(), l1_weight->size);
  dump(out, "dw1.f32", l2_weight->dev_grad.get(), l2_weight->size);
  dump(out, "dlogits.f32", output->dev_grad.get(), output->size);

  // ---- evaluation pass on the validation split: pristine input, forward(false)
  CHECK_CUDA_ERROR(cudaMemcpy(input->dev_data.get(), dev.dev_feature_value.get(), input->size * sizeof(real), cudaMemcpyDeviceToDevice));
  set_truth(data, 2, dev_truth);
  modules.back()->set_num_samples(params.val_dim);
  for (const auto &m : modules) m->forward(false, fwd);
  CHECK_CUDA_ERROR(cudaStreamSynchronize(fwd.get()));
  const real val_loss_sum = *loss;
  dump(out, "val_logits.f32", output->dev_data.get(), output->size);

  std::ofstream meta(out + "/meta.txt");
  meta << N << " " << F << " " << H << " " << C << " " << params.train_dim << " " << params.val_dim << " "
       << std::hexfloat << train_loss_sum << " " << val_loss_sum << "\n";
  printf("module_chain ok: train loss sum %.6f (%u samples), val loss sum %.6f (%u samples)\n", train_loss_sum,
         params.train_dim, val_loss_sum, params.val_dim);
  return 0;
}
