// exact_api.hpp -- what qb200_diagk.cu uses of the exact sampler (qb200_exact.cu) for the
// diagonal pipeline that keeps j on the device (qb200_diagk_sample_drawn).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/qunundrum_b200.h"

namespace qb200 {

uint32_t exact_chunk(const qb200_exact* s);
uint32_t exact_j_limbs(const qb200_exact* s);
int exact_device(const qb200_exact* s);
int exact_upload_stream(qb200_exact* s, const uint8_t* stream, uint64_t stream_len, const uint8_t** d_stream,
                        cudaStream_t st);
int exact_draw_j_tiles(qb200_exact* s, uint32_t B, const qb200_exact_region* regions, const uint32_t* t,
                       const uint8_t* d_stream, uint64_t stream_len, const uint32_t** d_jT,
                       const int32_t** d_status, cudaStream_t st);

}  // namespace qb200
