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

#define AVDN_REQUIRE(cond, ...)                                     \
  do {                                                              \
    if (!(cond)) return avdn::set_err(AVDN_ERR_BAD_ARG, __VA_ARGS__); \
  } while (0)
