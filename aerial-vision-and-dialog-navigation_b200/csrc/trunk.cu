// Stage 2 of the AVDN hot path: the xview-yolov3 Darknet trunk
// (src/models/dark_net.py:7-64,212-240) around the tensor-core convolutions.
//
// CUDA-core kernels that are HBM-bound by nature: train-mode BatchNorm statistics and
// application fused with LeakyReLU(0.01) and the shortcut add, their backward
// counterparts, and the weight (un)packing between the reference's
// [Cout,Cin,kh,kw] fp32 parameters and the bf16 GEMM operand layouts.
//
// Activations are NHWC bf16 with channels padded to a multiple of 64; a tensor
// is addressed as rows x C (rows = N*H*W).  Every elementwise kernel moves 16
// bytes (8 channels) per thread per access.
#include "common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    f[2 * i] = __low2float(b);
    f[2 * i + 1] = __high2float(b);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ------------------------------------------------------------ BN statistics
// Per-channel reductions over the R rows of a [R,C] bf16 tensor.  Every thread owns ONE group of
// 8 channels for its whole life (the grid stride is a multiple of C/8), so the per-channel
// constants sit in registers; UNR (4 or 8) independent 16-byte loads per tensor are in flight per
// thread.  Partial sums are fp32 per thread, reduced through shared memory per block and
// committed with one fp64 atomic per channel per block.
//   forward : sums[0][c] = sum z          sums[1][c] = sum z^2
//   backward: sums[0][c] = sum g          sums[1][c] = sum g * (z - mean)     g = da * leaky'(y)
constexpr int BN_THREADS = 256;
constexpr int BN_UNROLL = 4;

template <bool BWD, int UNR = BN_UNROLL>
__global__ void __launch_bounds__(BN_THREADS, UNR > 4 ? 2 : 0) bn_reduce_kernel(const uint4* __restrict__ z, const uint4* __restrict__ da,
                                                               const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ mean, float slope,
                                                               long long n8, int C, double* __restrict__ sums,
                                                               int rev) {
  avdn_pdl_trigger();
  avdn_pdl_wait();
  extern __shared__ float sred[];           // [2][RY][C]
  const int C8 = C >> 3;
  const long long tid = blockIdx.x * (long long)BN_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * BN_THREADS;      // multiple of C8 (host guarantees)
  // rev: walk the tensor back to front (element n8-1-k instead of k; n8 is a multiple of C8, so the
  // thread's channel group becomes C8-1-cx) -- see bn_order_flag() below
  const int cx = rev ? C8 - 1 - (int)(tid % C8) : (int)(tid % C8);
  const long long last = n8 - 1;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  float sc[8], sh[8], mu[8];
  if (BWD) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[cx * 8 + j]; sh[j] = shift[cx * 8 + j]; mu[j] = mean[cx * 8 + j]; }
  }
  for (long long i = tid; i < n8; i += stride * UNR) {
    uint4 zv[UNR], gv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long k = i + u * stride;
      if (k < n8) {
        const long long kk = rev ? last - k : k;
        zv[u] = __ldg(z + kk);
        if (BWD) gv[u] = __ldg(da + kk);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (i + u * stride < n8) {
        float f[8];
        unpack8(zv[u], f);
        if (!BWD) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s1[j] += f[j]; s2[j] = fmaf(f[j], f[j], s2[j]); }
        } else {
          float g[8];
          unpack8(gv[u], g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = fmaf(f[j], sc[j], sh[j]);
            const float gg = y > 0.f ? g[j] : g[j] * slope;
            s1[j] += gg;
            s2[j] = fmaf(gg, f[j] - mu[j], s2[j]);
          }
        }
      }
    }
  }
  // block reduction: threads with the same cx are BN_THREADS/C8 "row lanes" apart
  const int RY = BN_THREADS / C8, ry = threadIdx.x / C8;
  float* a1 = sred;
  float* a2 = sred + RY * C;
  // note: threadIdx.x % C8 == cx (C8-1-cx when rev) because BN_THREADS % C8 == 0: a bijection per row lane either way
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[ry * C + cx * 8 + j] = s1[j]; a2[ry * C + cx * 8 + j] = s2[j]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    float t1 = 0.f, t2 = 0.f;
    for (int y = 0; y < RY; ++y) { t1 += a1[y * C + c]; t2 += a2[y * C + c]; }
    atomicAdd(&sums[c], (double)t1);
    atomicAdd(&sums[C + c], (double)t2);
  }
}

// mean/var -> per-channel affine (scale, shift); running statistics update
// (nn.BatchNorm2d: momentum 0.1, eps 1e-5, unbiased variance for the running buffer).
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long R, int C, int C_real,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ rstd) {
  avdn_pdl_trigger();
  avdn_pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= C_real) { scale[c] = 0.f; shift[c] = 0.f; mean[c] = 0.f; rstd[c] = 0.f; return; }
  const double m = sums[c] / (double)R;
  double var = sums[C + c] / (double)R - m * m;
  if (var < 0.0) var = 0.0;
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rs;
  scale[c] = sc;
  shift[c] = beta[c] - (float)m * sc;
  mean[c] = (float)m;
  rstd[c] = rs;
  if (running_mean) {
    const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: affine from the running statistics
__global__ void bn_eval_coeffs_kernel(int C, int C_real, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ running_mean,
                                      const float* __restrict__ running_var, float eps, float* __restrict__ scale,
                                      float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= C_real) { scale[c] = 0.f; shift[c] = 0.f; return; }
  const float rs = 1.f / sqrtf(running_var[c] + eps);
  scale[c] = gamma[c] * rs;
  shift[c] = beta[c] - running_mean[c] * gamma[c] * rs;
}

// a = leaky(z*scale + shift) (+ residual)
template <int UNR>
__global__ void __launch_bounds__(BN_THREADS, UNR > 4 ? 2 : 0) bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const uint4* __restrict__ residual, uint4* __restrict__ a,
                                                              long long n8, int C8, float slope, int rev) {
  avdn_pdl_trigger();
  avdn_pdl_wait();
  const long long tid = blockIdx.x * (long long)BN_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * BN_THREADS;      // multiple of C8
  const int cx = rev ? C8 - 1 - (int)(tid % C8) : (int)(tid % C8);
  const long long last = n8 - 1;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[cx * 8 + j]; sh[j] = shift[cx * 8 + j]; }
  const bool has_res = residual != nullptr;
  for (long long i = tid; i < n8; i += stride * UNR) {
    uint4 zv[UNR], rv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long k = i + u * stride;
      if (k < n8) {
        const long long kk = rev ? last - k : k;
        zv[u] = __ldg(z + kk);
        if (has_res) rv[u] = __ldg(residual + kk);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long k = i + u * stride;
      if (k < n8) {
        float f[8], r[8];
        unpack8(zv[u], f);
        if (has_res) unpack8(rv[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = fmaf(f[j], sc[j], sh[j]);
          y = y > 0.f ? y : y * slope;
          f[j] = has_res ? y + r[j] : y;
        }
        a[rev ? last - k : k] = pack8(f);
      }
    }
  }
}

// per-channel coefficients of the backward apply pass:
//   dz = scale*(g - S1/R - xhat*S2/R),  xhat = (z-mean)*rstd,  S2 = rstd * sum g*(z-mean)
//      = scale*g + A*z + B,   A = -scale*rstd^2*sums1/R,   B = -scale*sums0/R - A*mean
// coef [4][C]: scale, shift, A, B.  Also dgamma += S2, dbeta += S1 for the real channels.
__global__ void bn_bwd_coef_kernel(const double* __restrict__ sums, double invR, int C, int C_real,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                   float* __restrict__ coef, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  avdn_pdl_trigger();
  avdn_pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sc = scale[c], rs = rstd[c], mu = mean[c];
  const double A = -sc * rs * rs * sums[C + c] * invR;
  const double B = -sc * sums[c] * invR - A * mu;
  coef[c] = scale[c];
  coef[C + c] = shift[c];
  coef[2 * C + c] = (float)A;
  coef[3 * C + c] = (float)B;
  if (c < C_real && dgamma && dbeta) {
    dbeta[c] += (float)sums[c];
    dgamma[c] += (float)(rs * sums[C + c]);
  }
}

template <int UNR>
__global__ void __launch_bounds__(BN_THREADS, UNR > 4 ? 2 : 0) bn_bwd_apply_kernel(const uint4* __restrict__ da, const uint4* __restrict__ z,
                                                                  const float* __restrict__ coef,
                                                                  uint4* __restrict__ dz, long long n8, int C,
                                                                  float slope, int rev) {
  avdn_pdl_trigger();
  avdn_pdl_wait();
  const int C8 = C >> 3;
  const long long tid = blockIdx.x * (long long)BN_THREADS + threadIdx.x;
  const long long stride = (long long)gridDim.x * BN_THREADS;      // multiple of C8
  const int cx = rev ? C8 - 1 - (int)(tid % C8) : (int)(tid % C8);
  const long long last = n8 - 1;
  float sc[8], sh[8], cA[8], cB[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = coef[cx * 8 + j]; sh[j] = coef[C + cx * 8 + j];
    cA[j] = coef[2 * C + cx * 8 + j]; cB[j] = coef[3 * C + cx * 8 + j];
  }
  for (long long i = tid; i < n8; i += stride * UNR) {
    uint4 zv[UNR], gv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long k = i + u * stride;
      if (k < n8) {
        const long long kk = rev ? last - k : k;
        zv[u] = __ldg(z + kk);
        gv[u] = __ldg(da + kk);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long k = i + u * stride;
      if (k < n8) {
        float f[8], g[8];
        unpack8(zv[u], f);
        unpack8(gv[u], g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float y = fmaf(f[j], sc[j], sh[j]);
          const float gg = y > 0.f ? g[j] : g[j] * slope;
          f[j] = fmaf(sc[j], gg, fmaf(cA[j], f[j], cB[j]));
        }
        dz[rev ? last - k : k] = pack8(f);
      }
    }
  }
}

// eval-mode backward is not needed (the trunk is only differentiated in train mode).

// ------------------------------------------------------- weight (un)packing
// w [Cout,Cin,k,k] fp32 -> wf [Cout_p][k*k][Cin_p] bf16 and wd [Cin_p][k*k][Cout_p] bf16 (zero padded)
__device__ __forceinline__ void pack_conv_weight_body(const float* __restrict__ w, int Cout, int Cin, int k, int Cout_p,
                                                      int Cin_p, __nv_bfloat16* __restrict__ wf,
                                                      __nv_bfloat16* __restrict__ wd) {
  const int kk = k * k;
  const long long n = (long long)Cout_p * kk * Cin_p;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin_p), t = (int)((i / Cin_p) % kk), co = (int)(i / ((long long)Cin_p * kk));
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((long long)co * Cin + ci) * kk + t];
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    wf[i] = b;
    wd[((long long)ci * kk + t) * Cout_p + co] = b;
  }
}
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int Cout_p, int Cin_p,
                                        __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  pack_conv_weight_body(w, Cout, Cin, k, Cout_p, Cin_p, wf, wd);
}
// every layer of the trunk in one launch: blockIdx.y = table entry.  Both outputs are transposes of the
// [Cout x Cin] matrix per tap, so a CTA moves 32(co) x 32(ci) x k*k tiles through shared memory: the fp32
// read is contiguous along (ci, tap), the wf write along ci, the wd write along co.
template <int KK>
__device__ __forceinline__ void pack_tiles(const avdn_conv_item& it, __nv_bfloat16 (*tile)[32][34]) {
  constexpr int kk = KK;
  const int tco = it.Cout_p / 32, tci = (it.Cin_p + 31) / 32;
  __nv_bfloat16* __restrict__ wf = reinterpret_cast<__nv_bfloat16*>(it.wf);
  __nv_bfloat16* __restrict__ wd = reinterpret_cast<__nv_bfloat16*>(it.wd);
  const int tid = threadIdx.x;
  for (int tl = blockIdx.x; tl < tco * tci; tl += gridDim.x) {
    const int co0 = (tl / tci) * 32, ci0 = (tl % tci) * 32;
    __syncthreads();
    // load: row r = co, contiguous run of 32*kk floats (ci, tap)
    constexpr int run = 32 * kk;
    for (int e = tid; e < 32 * run; e += 256) {
      const int r = e / run, q = e - r * run, ci = q / kk, t = q - ci * kk;
      const int co = co0 + r, cg = ci0 + ci;
      float v = 0.f;
      if (co < it.Cout && cg < it.Cin) v = __ldg(it.w + ((long long)co * it.Cin + cg) * kk + t);
      tile[t][r][ci] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    // wf[co][t][ci]: 32 contiguous ci per (co, t)
    for (int e = tid; e < 32 * kk * 32; e += 256) {
      const int ci = e & 31, rt = e >> 5, t = rt % kk, r = rt / kk;
      if (ci0 + ci < it.Cin_p)
        wf[((long long)(co0 + r) * kk + t) * it.Cin_p + ci0 + ci] = tile[t][r][ci];
    }
    // wd[ci][t][co]: 32 contiguous co per (ci, t)
    for (int e = tid; e < 32 * kk * 32; e += 256) {
      const int r = e & 31, ct = e >> 5, t = ct % kk, ci = ct / kk;
      if (ci0 + ci < it.Cin_p)
        wd[((long long)(ci0 + ci) * kk + t) * it.Cout_p + co0 + r] = tile[t][r][ci];
    }
  }
}
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const avdn_conv_item* __restrict__ items) {
  __shared__ __nv_bfloat16 tile[9][32][34];
  const avdn_conv_item it = items[blockIdx.y];
  if (it.k == 3) pack_tiles<9>(it, tile);
  else pack_tiles<1>(it, tile);
}
// a range of layers in one launch: blockIdx.y = table entry (relative to `items`).  Plain layout: tiles of
// 32(co) x 32(ci) x k*k through shared memory (dwf is contiguous along ci, grad along (ci, tap)).
__device__ __forceinline__ void unpack_conv_wgrad_body(const float* __restrict__ dwf, int Cout, int Cin, int k,
                                                       int Cin_p, float* __restrict__ grad) {
  const int kk = k * k;
  const long long n = (long long)Cout * Cin * kk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % kk), ci = (int)((i / kk) % Cin), co = (int)(i / ((long long)kk * Cin));
    grad[i] += dwf[((long long)co * kk + t) * Cin_p + ci];
  }
}

// Weight gradient computed on PIXEL-PAIR views (32-channel activations: two adjacent pixels form one
// 128-byte row, so the MN-major wgrad operands keep 64-element rows).
//   stride 1: both operands paired.  D'[(a,co)][blk(kh,s)][(b,ci)] = sum_P dZ[2P+a][co] X[2(P+s)+b (+kh row)][ci]
//             tap dw of pixel parity a lands in (s,b) = f(a+dw): -1->(-1,1) 0->(0,0) 1->(0,1) 2->(1,0)
//   stride 2: only X paired (pair index = output pixel).  D'[co][blk(kh,s)][(b,ci)], dw -> (s,b): -1->(-1,1) 0->(0,0) 1->(0,1)
// grad [Cout,Cin,k,k] += the blocks that make up each filter tap.
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwf, int Cout, int Cin, int k, int Cin_p,
                                         float* __restrict__ grad) {
  unpack_conv_wgrad_body(dwf, Cout, Cin, k, Cin_p, grad);
}
__device__ __forceinline__ void unpack_conv_wgrad_pairs_body(const float* __restrict__ dwp, int Cout, int Cin, int k,
                                                             int stride, int Cout_p, int Cin_p,
                                                             float* __restrict__ grad) {
  const int kk = k * k, pad = (k - 1) / 2;
  const int Np = 2 * Cin_p;
  const int nshift = (stride == 2) ? 2 : (k == 3 ? 3 : 1);
  const long long ld = (long long)k * nshift * Np;
  const long long n = (long long)Cout * Cin * kk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % kk), ci = (int)((i / kk) % Cin), co = (int)(i / ((long long)kk * Cin));
    const int kh = t / k, dw = t % k - pad;
    float v = 0.f;
    if (stride == 2) {
      const int s = (dw < 0) ? -1 : 0, b = (dw == 0) ? 0 : 1;
      v = dwp[(long long)co * ld + (long long)(kh * 2 + (s + 1)) * Np + b * Cin_p + ci];
    } else {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int u = a + dw;                         // -1 .. 2
        const int s = (u < 0) ? -1 : (u > 1 ? 1 : 0), b = u & 1;
        const int blk = kh * nshift + (k == 3 ? s + 1 : 0);
        v += dwp[(long long)(a * Cout_p + co) * ld + (long long)blk * Np + b * Cin_p + ci];
      }
    }
    grad[i] += v;
  }
}
__global__ void unpack_conv_wgrad_pairs_kernel(const float* __restrict__ dwp, int Cout, int Cin, int k, int stride,
                                               int Cout_p, int Cin_p, float* __restrict__ grad) {
  unpack_conv_wgrad_pairs_body(dwp, Cout, Cin, k, stride, Cout_p, Cin_p, grad);
}
template <int KK>
__device__ __forceinline__ void unpack_tiles(const avdn_conv_item& it, float (*tile)[32][33]) {
  constexpr int kk = KK;
  const int tco = (it.Cout + 31) / 32, tci = (it.Cin + 31) / 32;
  const int tid = threadIdx.x;
  for (int tl = blockIdx.x; tl < tco * tci; tl += gridDim.x) {
    const int co0 = (tl / tci) * 32, ci0 = (tl % tci) * 32;
    __syncthreads();
    for (int e = tid; e < 32 * kk * 32; e += 256) {
      const int ci = e & 31, rt = e >> 5, t = rt % kk, r = rt / kk;
      float v = 0.f;
      if (co0 + r < it.Cout && ci0 + ci < it.Cin)
        v = it.dwf[((long long)(co0 + r) * kk + t) * it.Cin_p + ci0 + ci];
      tile[t][r][ci] = v;
    }
    __syncthreads();
    constexpr int run = 32 * kk;
    for (int e = tid; e < 32 * run; e += 256) {
      const int r = e / run, q = e - r * run, ci = q / kk, t = q - ci * kk;
      const int co = co0 + r, cg = ci0 + ci;
      if (co < it.Cout && cg < it.Cin) it.grad[((long long)co * it.Cin + cg) * kk + t] += tile[t][r][ci];
    }
  }
}
__global__ void __launch_bounds__(256) unpack_conv_wgrads_kernel(const avdn_conv_item* __restrict__ items) {
  __shared__ float tile[9][32][33];
  const avdn_conv_item it = items[blockIdx.y];
  if (it.dwf == nullptr) return;       // this block's weight gradient was accumulated into grad directly
  if (it.pairs) {
    unpack_conv_wgrad_pairs_body(it.dwf, it.Cout, it.Cin, it.k, it.stride, it.Cout_p, it.Cin_p, it.grad);
    return;
  }
  if (it.k == 3) unpack_tiles<9>(it, tile);
  else unpack_tiles<1>(it, tile);
}

// fp32 -> bf16 cast (linear-layer weights, features)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// NHWC bf16 [N,HW,C] -> NCHW-flattened fp32 [N,C,HW] (trunk output -> ET `frames`) and its adjoint
__global__ void nhwc_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int N, int HW,
                                        int C) {
  const long long n = (long long)N * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW), c = (int)((i / HW) % C);
    const long long b = i / ((long long)HW * C);
    out[i] = __bfloat162float(in[(b * HW + p) * C + c]);
  }
}
__global__ void nchw_f32_to_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int N, int HW,
                                        int C) {
  const long long n = (long long)N * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), p = (int)((i / C) % HW);
    const long long b = i / ((long long)HW * C);
    out[i] = __float2bfloat16_rn(in[(b * C + c) * HW + p]);
  }
}

inline int grid_for(long long n, int block = 256, int waves = 8) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)avdn::sm_count() * waves;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

// ===================================================================== C ABI
// Traversal order (avdn_bn_set_order / AVDN_BN_ORDER, a bit mask; 0 = every pass front to back in a multi-wave
// grid, the round-1 behaviour).  Measured inside the config-2 training step on one B200 (tools/bn_order_ab.py,
// profiles/r02_bn_order_ab.txt): one-wave grids (only co-resident CTAs, each thread looping over the tensor) take the
// step's BatchNorm calls from 16.5 to 14.8 ms -- the backward passes ran 16 waves of CTAs that each loaded their
// per-channel coefficients for ONE batch of four loads.  Direction: the producer of a tensor leaves its TAIL in L2
// (the convolutions walk their tiles front to back) and the consumer of the pass's output starts at its HEAD, so a
// pass that walks back to front takes its first bytes from L2 and leaves the head of what it writes there; that is
// worth little inside the BatchNorm calls themselves (14.8 -> 14.6..14.8 ms) and ~0.5-1 ms of the step in the
// convolutions that follow a reversed forward apply.  Eight loads per tensor in flight per thread (bit 16, two CTAs of
// <= 128 registers per SM) take the backward passes from 9.97 to 9.4 ms.  Default 25 = 16 | 8 | 1.
//   1 = forward apply back to front      2 = backward reduce back to front      4 = backward apply back to front
//   8 = one-wave grids (so that the grid-stride sweep is monotone in time; implied by 1|2|4)
//  16 = eight instead of four 16-byte loads per tensor in flight per thread (one-wave grids; <= 128 registers)
constexpr int BN_ORDER_DEFAULT = 25;
static int& bn_order_flag() {
  static int flag = [] {
    const char* e = getenv("AVDN_BN_ORDER");
    return e ? atoi(e) & 31 : BN_ORDER_DEFAULT;
  }();
  return flag;
}
extern "C" int avdn_bn_set_order(int mask) {
  int& f = bn_order_flag();
  const int old = f;
  if (mask >= 0) f = mask & 31;
  return old;
}
template <typename K>
static int bn_resident_per_sm(K kernel, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, BN_THREADS, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}
// grid of BN_THREADS-wide blocks whose total thread count is a multiple of C/8 (so that every
// thread keeps its channel group) and that fills the machine without exceeding the work
static int bn_grid(long long n8, int waves, int unroll = BN_UNROLL) {
  long long blocks = (n8 + (long long)BN_THREADS * unroll - 1) / ((long long)BN_THREADS * unroll);
  const long long cap = (long long)avdn::sm_count() * waves;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

static void bn_bwd_apply_launch(const void* da, const void* z, const float* coef, void* dz, long long n8, int C,
                                float slope, cudaStream_t s) {
  const int order = bn_order_flag();
  const int rev = (order & 4) ? 1 : 0;
  if (order & 16) {
    static const int occ = bn_resident_per_sm(bn_bwd_apply_kernel<8>, 0);
    avdn::launch_pdl(bn_bwd_apply_kernel<8>, dim3(bn_grid(n8, occ, 8)), dim3(BN_THREADS), 0, s,
                     reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(z), coef,
                     reinterpret_cast<uint4*>(dz), n8, C, slope, rev);
    return;
  }
  int waves = 16;
  if (order) {
    static const int occ = bn_resident_per_sm(bn_bwd_apply_kernel<BN_UNROLL>, 0);
    waves = occ;
  }
  avdn::launch_pdl(bn_bwd_apply_kernel<BN_UNROLL>, dim3(bn_grid(n8, waves)), dim3(BN_THREADS), 0, s,
                   reinterpret_cast<const uint4*>(da), reinterpret_cast<const uint4*>(z), coef,
                   reinterpret_cast<uint4*>(dz), n8, C, slope, rev);
}

static int bn_reduce_launch(bool bwd, const void* z, const void* da, const float* scale, const float* shift,
                            const float* mean, float slope, long long R, int C, double* sums, cudaStream_t s) {
  AVDN_REQUIRE(C % 8 == 0 && C >= 8 && BN_THREADS % (C / 8) == 0,
               "bn reduce: C=%d must be 8 * (a divisor of %d)", C, BN_THREADS);
  const int C8 = C / 8;
  const int RY = BN_THREADS / C8;
  const long long n8 = R * C8;
  const size_t smem = (size_t)2 * RY * C * sizeof(float);      // = 2 * 256 * 8 * 4 = 16 KB
  const int order = bn_order_flag();
  const int rev = bwd && (order & 2) ? 1 : 0;
  int waves = 8;
  if (order && bwd) {      // the forward statistics (not on the training path: the convolutions produce them) keep their grid
    static const int occ_b = bn_resident_per_sm(bn_reduce_kernel<true>, 16384);
    waves = occ_b;
  }
  const int blocks = bn_grid(n8, waves);
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, s) != cudaSuccess) return avdn::check_launch("bn reduce memset");
  if (bwd && (order & 16)) {
    static const int occ8 = bn_resident_per_sm(bn_reduce_kernel<true, 8>, 16384);
    avdn::launch_pdl(bn_reduce_kernel<true, 8>, dim3(bn_grid(n8, occ8, 8)), dim3(BN_THREADS), smem, s,
                     reinterpret_cast<const uint4*>(z), reinterpret_cast<const uint4*>(da), scale, shift, mean, slope,
                     n8, C, sums, rev);
  } else if (bwd)
    avdn::launch_pdl(bn_reduce_kernel<true>, dim3(blocks), dim3(BN_THREADS), smem, s, reinterpret_cast<const uint4*>(z),
                     reinterpret_cast<const uint4*>(da), scale, shift, mean, slope, n8, C, sums, rev);
  else
    avdn::launch_pdl(bn_reduce_kernel<false>, dim3(blocks), dim3(BN_THREADS), smem, s, reinterpret_cast<const uint4*>(z),
                     nullptr, nullptr, nullptr, nullptr, slope, n8, C, sums, 0);
  return avdn::check_launch("bn_reduce_kernel");
}

extern "C" int avdn_bn_stats(const void* z, long long R, int C, int C_real, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, float momentum, float eps, double* sums,
                             float* scale, float* shift, float* mean, float* rstd, avdn_stream_t stream) {
  AVDN_REQUIRE(z && gamma && beta && sums && scale && shift && mean && rstd && R > 0, "avdn_bn_stats: bad argument");
  cudaStream_t s = avdn::to_cuda(stream);
  int r = bn_reduce_launch(false, z, nullptr, nullptr, nullptr, nullptr, 0.f, R, C, sums, s);
  if (r) return r;
  avdn::launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, R, C, C_real, gamma, beta,
                   running_mean, running_var, momentum, eps, scale, shift, mean, rstd);
  return avdn::check_launch("bn_finalize_kernel");
}

extern "C" int avdn_bn_finalize(const double* sums, long long R, int C, int C_real, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum,
                                float eps, float* scale, float* shift, float* mean, float* rstd,
                                avdn_stream_t stream) {
  AVDN_REQUIRE(sums && gamma && beta && scale && shift && mean && rstd && R > 0, "avdn_bn_finalize: bad argument");
  avdn::launch_pdl(bn_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, avdn::to_cuda(stream), sums, R, C, C_real,
                   gamma, beta, running_mean, running_var, momentum, eps, scale, shift, mean, rstd);
  return avdn::check_launch("bn_finalize_kernel");
}

extern "C" int avdn_bn_eval_coeffs(int C, int C_real, const float* gamma, const float* beta,
                                   const float* running_mean, const float* running_var, float eps, float* scale,
                                   float* shift, avdn_stream_t stream) {
  AVDN_REQUIRE(gamma && beta && running_mean && running_var && scale && shift, "avdn_bn_eval_coeffs: null pointer");
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, avdn::to_cuda(stream)>>>(C, C_real, gamma, beta, running_mean,
                                                                           running_var, eps, scale, shift);
  return avdn::check_launch("bn_eval_coeffs_kernel");
}

extern "C" int avdn_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, void* a,
                             long long R, int C, float slope, avdn_stream_t stream) {
  AVDN_REQUIRE(z && scale && shift && a && R > 0 && C % 8 == 0 && BN_THREADS % (C / 8) == 0,
               "avdn_bn_apply: bad argument (C=%d)", C);
  const long long n8 = R * (C / 8);
  const int order = bn_order_flag();
  if (order & 16) {
    static const int occ8 = bn_resident_per_sm(bn_apply_kernel<8>, 0);
    avdn::launch_pdl(bn_apply_kernel<8>, dim3(bn_grid(n8, occ8, 8)), dim3(BN_THREADS), 0, avdn::to_cuda(stream),
                     reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<const uint4*>(residual),
                     reinterpret_cast<uint4*>(a), n8, C / 8, slope, order & 1);
    return avdn::check_launch("avdn_bn_apply");
  }
  int waves = 16;
  if (order) {
    static const int occ = bn_resident_per_sm(bn_apply_kernel<BN_UNROLL>, 0);
    waves = occ;
  }
  avdn::launch_pdl(bn_apply_kernel<BN_UNROLL>, dim3(bn_grid(n8, waves)), dim3(BN_THREADS), 0, avdn::to_cuda(stream),
                   reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<const uint4*>(residual),
                   reinterpret_cast<uint4*>(a), n8, C / 8, slope, order & 1);
  return avdn::check_launch("avdn_bn_apply");
}

extern "C" int avdn_bn_backward(const void* da, const void* z, const float* scale, const float* shift,
                                const float* mean, const float* rstd, long long R, int C, int C_real, float slope,
                                double* sums, void* dz, float* dgamma, float* dbeta, avdn_stream_t stream) {
  AVDN_REQUIRE(da && z && scale && shift && mean && rstd && sums && dz && R > 0, "avdn_bn_backward: bad argument");
  cudaStream_t s = avdn::to_cuda(stream);
  int r = bn_reduce_launch(true, z, da, scale, shift, mean, slope, R, C, sums, s);
  if (r) return r;
  float* coef = reinterpret_cast<float*>(sums + 2 * C);         // second half of the [4,C] f64 scratch
  avdn::launch_pdl(bn_bwd_coef_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, 1.0 / (double)R, C, C_real, scale,
                   shift, mean, rstd, coef, dgamma, dbeta);
  r = avdn::check_launch("bn_bwd_coef_kernel");
  if (r) return r;
  const long long n8 = R * (C / 8);
  bn_bwd_apply_launch(da, z, coef, dz, n8, C, slope, s);
  return avdn::check_launch("bn_bwd_apply_kernel");
}

extern "C" int avdn_bn_backward_apply(const void* da, const void* z, const float* scale, const float* shift,
                                      const float* mean, const float* rstd, long long R, int C, int C_real,
                                      float slope, const double* sums, float* coef, void* dz, float* dgamma,
                                      float* dbeta, avdn_stream_t stream) {
  AVDN_REQUIRE(da && z && scale && shift && mean && rstd && sums && coef && dz && R > 0,
               "avdn_bn_backward_apply: bad argument");
  AVDN_REQUIRE(C % 8 == 0 && C >= 8 && BN_THREADS % (C / 8) == 0,
               "avdn_bn_backward_apply: C=%d must be 8 * (a divisor of %d)", C, BN_THREADS);
  cudaStream_t s = avdn::to_cuda(stream);
  avdn::launch_pdl(bn_bwd_coef_kernel, dim3((C + 127) / 128), dim3(128), 0, s, sums, 1.0 / (double)R, C, C_real, scale,
                   shift, mean, rstd, coef, dgamma, dbeta);
  int r = avdn::check_launch("bn_bwd_coef_kernel");
  if (r) return r;
  const long long n8 = R * (C / 8);
  bn_bwd_apply_launch(da, z, coef, dz, n8, C, slope, s);
  return avdn::check_launch("bn_bwd_apply_kernel");
}

extern "C" int avdn_pack_conv_weight(const float* w, int Cout, int Cin, int k, int Cout_p, int Cin_p, void* wf,
                                     void* wd, avdn_stream_t stream) {
  AVDN_REQUIRE(w && wf && wd && Cout_p >= Cout && Cin_p >= Cin, "avdn_pack_conv_weight: bad argument");
  const long long n = (long long)Cout_p * k * k * Cin_p;
  pack_conv_weight_kernel<<<grid_for(n), 256, 0, avdn::to_cuda(stream)>>>(
      w, Cout, Cin, k, Cout_p, Cin_p, reinterpret_cast<__nv_bfloat16*>(wf), reinterpret_cast<__nv_bfloat16*>(wd));
  return avdn::check_launch("avdn_pack_conv_weight");
}

extern "C" int avdn_pack_conv_weights(const avdn_conv_item* items_dev, int n_items, avdn_stream_t stream) {
  AVDN_REQUIRE(n_items >= 0 && n_items <= 65535, "avdn_pack_conv_weights: bad item count");
  if (n_items == 0) return AVDN_OK;
  AVDN_REQUIRE(items_dev, "avdn_pack_conv_weights: null table");
  pack_conv_weights_kernel<<<dim3(2 * avdn::sm_count(), n_items), 256, 0, avdn::to_cuda(stream)>>>(items_dev);
  return avdn::check_launch("avdn_pack_conv_weights");
}

extern "C" int avdn_unpack_conv_wgrads(const avdn_conv_item* items_dev, int first, int count, avdn_stream_t stream) {
  AVDN_REQUIRE(first >= 0 && count >= 0 && count <= 65535, "avdn_unpack_conv_wgrads: bad range");
  if (count == 0) return AVDN_OK;
  AVDN_REQUIRE(items_dev, "avdn_unpack_conv_wgrads: null table");
  unpack_conv_wgrads_kernel<<<dim3(2 * avdn::sm_count(), count), 256, 0, avdn::to_cuda(stream)>>>(items_dev + first);
  return avdn::check_launch("avdn_unpack_conv_wgrads");
}

extern "C" int avdn_unpack_conv_wgrad(const float* dwf, int Cout, int Cin, int k, int Cin_p, float* grad,
                                      avdn_stream_t stream) {
  AVDN_REQUIRE(dwf && grad, "avdn_unpack_conv_wgrad: null pointer");
  const long long n = (long long)Cout * Cin * k * k;
  unpack_conv_wgrad_kernel<<<grid_for(n), 256, 0, avdn::to_cuda(stream)>>>(dwf, Cout, Cin, k, Cin_p, grad);
  return avdn::check_launch("avdn_unpack_conv_wgrad");
}

extern "C" int avdn_unpack_conv_wgrad_pairs(const float* dwp, int Cout, int Cin, int k, int stride, int Cout_p,
                                            int Cin_p, float* grad, avdn_stream_t stream) {
  AVDN_REQUIRE(dwp && grad && (k == 1 || k == 3) && (stride == 1 || (stride == 2 && k == 3)),
               "avdn_unpack_conv_wgrad_pairs: bad argument");
  const long long n = (long long)Cout * Cin * k * k;
  unpack_conv_wgrad_pairs_kernel<<<grid_for(n), 256, 0, avdn::to_cuda(stream)>>>(dwp, Cout, Cin, k, stride, Cout_p,
                                                                                  Cin_p, grad);
  return avdn::check_launch("avdn_unpack_conv_wgrad_pairs");
}

extern "C" int avdn_cast_f32_bf16(const float* in, void* out, long long n, avdn_stream_t stream) {
  AVDN_REQUIRE(n >= 0, "avdn_cast_f32_bf16: n < 0");
  if (n == 0) return AVDN_OK;
  AVDN_REQUIRE(in && out, "avdn_cast_f32_bf16: null pointer");
  cast_f32_bf16_kernel<<<grid_for(n), 256, 0, avdn::to_cuda(stream)>>>(in, reinterpret_cast<__nv_bfloat16*>(out), n);
  return avdn::check_launch("avdn_cast_f32_bf16");
}

extern "C" int avdn_nhwc_to_nchw_f32(const void* in, float* out, int N, int HW, int C, avdn_stream_t stream) {
  AVDN_REQUIRE(in && out && N > 0, "avdn_nhwc_to_nchw_f32: bad argument");
  nhwc_to_nchw_f32_kernel<<<grid_for((long long)N * HW * C), 256, 0, avdn::to_cuda(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), out, N, HW, C);
  return avdn::check_launch("avdn_nhwc_to_nchw_f32");
}

extern "C" int avdn_nchw_f32_to_nhwc(const float* in, void* out, int N, int HW, int C, avdn_stream_t stream) {
  AVDN_REQUIRE(in && out && N > 0, "avdn_nchw_f32_to_nhwc: bad argument");
  nchw_f32_to_nhwc_kernel<<<grid_for((long long)N * HW * C), 256, 0, avdn::to_cuda(stream)>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), N, HW, C);
  return avdn::check_launch("avdn_nchw_f32_to_nhwc");
}
