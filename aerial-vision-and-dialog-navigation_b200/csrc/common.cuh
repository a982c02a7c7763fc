// Shared host/device helpers for libavdn.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/avdn.h"

namespace avdn {

// thread-local error text behind avdn_last_error_string()
char* err_buf();
int set_err(int code, const char* fmt, ...);

// cudaGetLastError() -> AVDN_ERR_LAUNCH with the CUDA message
int check_launch(const char* what);

inline cudaStream_t to_cuda(avdn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

}  // namespace avdn

// Stateless dropout: element `idx` of dropout site `site` in the step with seed `seed` is kept iff a
// 24-bit hash of (seed, site, idx) is >= p * 2^24.  Forward and backward kernels re-evaluate the same
// hash, so no mask is ever stored.  (nn.Dropout semantics: kept values are scaled by 1/(1-p).)
__device__ __forceinline__ bool avdn_drop_keep(unsigned long long seed, unsigned int site, unsigned long long idx,
                                               unsigned int thresh24) {
  unsigned long long x = idx * 0x9E3779B97F4A7C15ull + (seed ^ ((unsigned long long)site * 0xD1B54A32D192ED03ull));
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  return (unsigned int)(x >> 40) >= thresh24;
}
inline unsigned int avdn_drop_thresh(float p) { return (unsigned int)(p * 16777216.0f); }

#define AVDN_REQUIRE(cond, ...)                                     \
  do {                                                              \
    if (!(cond)) return avdn::set_err(AVDN_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)
