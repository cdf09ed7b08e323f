// devbuf.hpp -- a growing device buffer and the CUDA error macro shared by the translation units
// of the samplers (qb200_diagk.cu, qb200_exact.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <string>

#include "ctx_access.hpp"

#define QD_CUDA(call)                                                                   \
  do {                                                                                  \
    const cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess)                                                              \
      return qb200::set_error(-100, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace qb200 {

struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DBuf() {
    if (p) cudaFree(p);
  }
  int reserve(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    QD_CUDA(cudaMalloc(&p, n));
    bytes = n;
    return 0;
  }
  template <class T>
  T* as() const {
    return (T*)p;
  }
};

}  // namespace qb200
