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

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains (its CTAs
// become resident and run their prologue as SM resources free up) and blocks in avdn_pdl_wait() until the
// predecessor has completed and flushed, so the fill/drain gap between two dependent launches -- ~550 kernels per
// training step -- overlaps with the tail of the previous kernel.  OFF by default: measured on the B = 64 training
// step (power-capped at ~1.7 GHz) it changes nothing, 58.3-58.9 ms with and without;
// AVDN_PDL=1 in the environment turns the attribute on for the tcgen05 GEMM and the BatchNorm kernels.  Kernels
// launched this way must call avdn_pdl_wait() before their first global-memory access that depends on earlier work.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace avdn

// first statement of a kernel launched with avdn::launch_pdl: let the successor start its own prologue, then wait
// for the predecessor's results (both are no-ops in a plain launch)
__device__ __forceinline__ void avdn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void avdn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
