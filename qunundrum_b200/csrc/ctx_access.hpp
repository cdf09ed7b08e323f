// ctx_access.hpp -- what the other translation units of the library may see of a
// qb200_context (defined in qb200.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

struct qb200_context;

namespace qb200 {

struct TextState;  // qb200_text.cu

struct CtxView {
  int device;
  int sm_count;
  cudaStream_t stream;
  uint64_t* launches;
  TextState** text;
};

CtxView ctx_view(qb200_context* ctx);
int set_error(int code, const std::string& msg);  // becomes qb200_last_error()
void text_state_destroy(TextState* st);

}  // namespace qb200
