// Block 0 of the trunk on the tcgen05 tensor cores (conv0_tc.cu); dispatched from the C ABI in conv0.cu.
#pragma once
#include "common.cuh"

namespace avdn {

constexpr long long CONV0_TC_MAX_PIXELS = (1ll << 31) - 256;

int conv0_tc_fwd_stats(const void* x, const float* w, int N, int H, int W, double* stats, float* zw, double* xs9,
                       cudaStream_t s);
// mask: NULL (eval) or [N*H*W] u32 receiving the sign bits of a for conv0_tc_bwd
int conv0_tc_apply(const void* x, const float* w, const float* scale, const float* shift, float slope, void* a,
                   uint32_t* mask, int N, int H, int W, cudaStream_t s);
int conv0_tc_bwd(const void* x, const float* w, const void* da, const uint32_t* mask, const float* scale,
                 const float* mean, const float* rstd, float slope, int N, int H, int W, const float* zw,
                 const double* xs9, double* sums, float* gw, float* dw, float* dgamma, float* dbeta, cudaStream_t s);

}  // namespace avdn
