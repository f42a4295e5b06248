// utils.cuh -- mirror of the reference's include/utils.cuh:10-139 (types, CudaParams, CHECK_CUDA_ERROR, CEIL,
// print_gpu_info, string2vec).  CudaParams::N_THREADS / N_BLOCKS are kept because the reference's drivers assign
// them (test/performance_gpu.cpp:37-49, src/parser.cpp:245-247); the B200 kernels size their own grids from the
// SM count, so the two are accepted and ignored.  randState is an empty tag: the engine's Philox is stateless.
#ifndef UTILS_CUH
#define UTILS_CUH
#include <cmath>
#include <cstdlib>
#include <cuda_runtime.h>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

using natural = unsigned;
using integer = int;
using real = float;
struct randState {};  // reference: curandStatePhilox4_32_10_t (64 B of HBM per 4 elements); here: no state

namespace CudaParams {
inline natural N_THREADS;
inline natural N_BLOCKS;
constexpr natural TILE_DIM = 16;
inline natural SEED = 19990304;
}  // namespace CudaParams

#define CHECK_CUDA_ERROR(val) check((val), #val, __FILE__, __LINE__)
template <typename T>
void check(T err, const char *const func, const char *const file, const int line) {
  if (err != cudaSuccess) {
    std::cerr << "CUDA Runtime Error at: " << file << ":" << line << std::endl;
    std::cerr << cudaGetErrorString(static_cast<cudaError_t>(err)) << " " << func << std::endl;
    std::exit(EXIT_FAILURE);
  }
}
// libgcn_b200 kernels return an int status; any failure is fatal (the reference's error convention:
// print to std::cerr and exit, include/utils.cuh:29-40).  No CPU fallback exists.
void gcnb_check_fatal(int code, const char *what, const char *file, int line);
#define GCNB_CALL(expr) gcnb_check_fatal((expr), #expr, __FILE__, __LINE__)

#define CEIL(M, N) (((M) + (N)-1) / (N))

void print_gpu_info();

template <class T>
std::vector<T> string2vec(const std::string &str, char sep = ',') {
  std::vector<T> values;
  std::istringstream iss(str);
  std::string token;
  while (std::getline(iss, token, sep)) {
    T value;
    if (std::is_same<T, int>::value) value = static_cast<T>(std::stoi(token));
    else if (std::is_same<T, float>::value) value = static_cast<T>(std::stof(token));
    else if (std::is_same<T, double>::value) value = static_cast<T>(std::stod(token));
    else if (std::is_same<T, unsigned>::value) value = static_cast<T>(std::stoul(token));
    else {
      std::cerr << "ERROR: type not supported" << std::endl;
      exit(EXIT_FAILURE);
    }
    values.push_back(value);
  }
  return values;
}
#endif
