// Halo-tile 3x3 convolution for the thin layers of the trunk (conv3_halo.cu) and the tensor-map helper it shares
// with gemm.cu.
#pragma once
#include "common.cuh"

#include <cuda.h>

namespace avdn {

// rank-4 tensor map with a swizzled box whose inner extent is exactly 128 or 64 bytes (gemm.cu)
int encode_tensor_map_4d(const void* ptr, int elem_bytes, const int64_t* dim, const int64_t* stride,
                         const int32_t* boxdim, CUtensorMap* out);

bool conv3_halo_supported(int H, int W, int Cin, int Cout);
int conv3_halo_fwd(const void* x, const void* wf, void* z, int N, int H, int W, int Cin, int Cout, double* stats,
                   cudaStream_t s);
bool conv3_halo_dgrad_supported(int H, int W, int Cin, int Cout);
int conv3_halo_dgrad(const void* dz, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout, cudaStream_t s);

int conv3_halo_wgrad(const void* dz, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, cudaStream_t s);

}  // namespace avdn
