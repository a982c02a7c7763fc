// First convolution of the Darknet trunk: 3 -> 32 channels, 3x3, pad 1 (module_list.0.conv_0,
// src/models/dark_net.py:22-28) -- forward and weight gradient.
//
// K = 27 is too thin for a tcgen05 tile (the 128 x N x 16 UMMA atom would run at < 20 % useful
// work and the im2col operand cannot be described to TMA with 8-byte pixels), and both passes are
// HBM-bound by the 64-channel-padded bf16 activation they write / read.  They run on warp-level
// mma.sync (m16n8k16, bf16 -> fp32): the 3x3x(R,G,B,0) patch of a pixel is three contiguous
// 12-element windows of the input rows, so with k = kw*4 + c (padded 12 -> 16) per filter row the
// A / B fragments are plain 32-bit / 16-bit shared-memory loads of the staged input rows.
//
//   forward : Z[pixel][co]   = sum_kh  P_kh[pixel][16] . W_kh[16][co]      (+ fused BN statistics)
//   wgrad   : dW_kh[co][16] += sum_pixel dZ^T[co][pixel] . P_kh[pixel][16]
#include "common.cuh"
#include "conv0_tc.cuh"

#include <stdlib.h>

namespace {

constexpr int C0_OUT = 32;     // real output channels
constexpr int C0_PAD = 32;     // stored as 32 channels (64-byte pixels; block 1 reads 32-element k-blocks)
constexpr int WARPS = 7;       // 7 warps x 16 pixels = half a 224-pixel row
constexpr int THREADS = WARPS * 32;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// stage `rows` input rows [h_first, h_first+rows) of image n into smem with a zero halo pixel on each
// side: xs[r][(W+2)] pixels of 8 bytes (R,G,B,0 bf16).  Rows outside the image are zero.
__device__ __forceinline__ void stage_x(uint2* xs, const uint2* __restrict__ x, int n, int h_first, int rows, int H,
                                        int W) {
  const int pitch = W + 2;
  for (int i = threadIdx.x; i < rows * pitch; i += THREADS) {
    const int r = i / pitch, q = i - r * pitch;            // q = pixel + 1
    const int h = h_first + r, w = q - 1;
    uint2 v = make_uint2(0u, 0u);
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(x + ((size_t)n * H + h) * W + w);
    xs[i] = v;
  }
}

// ------------------------------------------------------------------ forward
// One CTA iteration = 4 output rows of one image; a warp owns 16-pixel tiles.
constexpr int F_ROWS = 4;
constexpr int OUT_PITCH = 80;          // bytes per pixel row of the per-warp output staging (64 + pad)

__global__ void __launch_bounds__(THREADS) conv0_fwd_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                            uint4* __restrict__ z, int N, int H, int W,
                                                            double* __restrict__ stats,
                                                            const float* __restrict__ aff_scale,
                                                            const float* __restrict__ aff_shift, float slope,
                                                            int round_first) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* xs = reinterpret_cast<uint2*>(smem);                                  // [(F_ROWS+2)][W+2]
  uint8_t* outs = smem + ((size_t)(F_ROWS + 2) * (W + 2) + 2) * 8;              // [WARPS][16][OUT_PITCH]
  // the 16-wide patch window of the last pixels runs 2 pixels past the last staged row (zero weights)
  if (threadIdx.x < 2) xs[(F_ROWS + 2) * (W + 2) + threadIdx.x] = make_uint2(0u, 0u);
  __shared__ float s_stat[WARPS][2 * C0_OUT];          // one set per warp: summed in warp order (reproducible)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  // weight fragments: B[k = kw*4 + ci][n = co] per filter row kh, n-tile nt (co = nt*8 + g)
  uint32_t wf[3][4][2];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * t + 8 * r + e, kw = k >> 2, ci = k & 3, co = nt * 8 + g;
          v[e] = (kw < 3 && ci < 3) ? w[((co * 3 + ci) * 3 + kh) * 3 + kw] : 0.f;
        }
        wf[kh][nt][r] = pack_bf16(v[0], v[1]);
      }
  float st1[4][2], st2[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { st1[nt][0] = st1[nt][1] = st2[nt][0] = st2[nt][1] = 0.f; }
  // eval mode: BatchNorm affine + LeakyReLU folded into the store (this lane's channels nt*8 + 2t, +1)
  float asc[4][2], ash[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      asc[nt][e] = aff_scale ? aff_scale[nt * 8 + 2 * t + e] : 1.f;
      ash[nt][e] = aff_scale ? aff_shift[nt * 8 + 2 * t + e] : 0.f;
    }

  const int strips_per_img = H / F_ROWS, tiles_per_row = W / 16;
  const int pitch = W + 2;
  const long long total = (long long)N * strips_per_img;
  uint8_t* my_out = outs + warp * 16 * OUT_PITCH;
  for (long long s = blockIdx.x; s < total; s += gridDim.x) {
    const int n = (int)(s / strips_per_img), h0 = (int)(s % strips_per_img) * F_ROWS;
    __syncthreads();                                   // previous strip fully consumed
    stage_x(xs, x, n, h0 - 1, F_ROWS + 2, H, W);
    __syncthreads();
    for (int tile = warp; tile < F_ROWS * tiles_per_row; tile += WARPS) {
      const int hr = tile / tiles_per_row, w0 = (tile % tiles_per_row) * 16;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        // A[m = pixel][k]: the 16 values starting at input pixel (w0 + m - 1), i.e. staged pixel w0 + m
        const uint8_t* rowp = reinterpret_cast<const uint8_t*>(xs + (size_t)(hr + kh) * pitch + w0);
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(rowp + g * 8 + t * 4);
        a[1] = *reinterpret_cast<const uint32_t*>(rowp + (g + 8) * 8 + t * 4);
        a[2] = *reinterpret_cast<const uint32_t*>(rowp + g * 8 + t * 4 + 16);
        a[3] = *reinterpret_cast<const uint32_t*>(rowp + (g + 8) * 8 + t * 4 + 16);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a, wf[kh][nt][0], wf[kh][nt][1]);
      }
      // C fragment: (pixel g, co nt*8+2t..+1), (pixel g+8, ...) -> bf16 -> per-warp staging
      if (aff_scale) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // train mode (second pass of the recompute path): the affine sees z as it would have been stored
            if (round_first) acc[nt][e] = __bfloat162float(__float2bfloat16_rn(acc[nt][e]));
            const float y = fmaf(acc[nt][e], asc[nt][e & 1], ash[nt][e & 1]);
            acc[nt][e] = fmaxf(y, y * slope);
          }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t lo = pack_bf16(acc[nt][0], acc[nt][1]), hi = pack_bf16(acc[nt][2], acc[nt][3]);
        *reinterpret_cast<uint32_t*>(my_out + g * OUT_PITCH + (nt * 8 + 2 * t) * 2) = lo;
        *reinterpret_cast<uint32_t*>(my_out + (g + 8) * OUT_PITCH + (nt * 8 + 2 * t) * 2) = hi;
        if (stats) {
          const __nv_bfloat162 l2 = *reinterpret_cast<const __nv_bfloat162*>(&lo);
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&hi);
          const float v0 = __low2float(l2), v1 = __high2float(l2), v2 = __low2float(h2), v3 = __high2float(h2);
          st1[nt][0] += v0 + v2; st1[nt][1] += v1 + v3;
          st2[nt][0] = fmaf(v0, v0, fmaf(v2, v2, st2[nt][0]));
          st2[nt][1] = fmaf(v1, v1, fmaf(v3, v3, st2[nt][1]));
        }
      }
      __syncwarp();
      // 16 pixels x 64 bytes, 512 contiguous bytes per instruction
      uint4* dst = z + (((size_t)n * H + h0 + hr) * W + w0) * (C0_PAD / 8);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qd = i * 32 + lane, px = qd >> 2, part = qd & 3;
        dst[qd] = *reinterpret_cast<const uint4*>(my_out + px * OUT_PITCH + part * 16);
      }
      __syncwarp();
    }
  }
  if (stats) {
    // lanes with the same t hold the same channels: reduce over g, then one slot per (warp, channel)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = st1[nt][e], b = st2[nt][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (g == 0) { s_stat[warp][nt * 8 + 2 * t + e] = a; s_stat[warp][C0_OUT + nt * 8 + 2 * t + e] = b; }
      }
    __syncthreads();
    if (threadIdx.x < 2 * C0_OUT) {
      const int which = threadIdx.x / C0_OUT, c = threadIdx.x % C0_OUT;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) v += s_stat[w][threadIdx.x];
      atomicAdd(stats + which * C0_PAD + c, (double)v);
    }
  }
}

// ------------------------------------------------------------------- wgrad
// One CTA iteration = 2 output rows of one image.  D_kh[co][k] += dZ^T[co][pixel] . P_kh[pixel][k].
constexpr int G_ROWS = 2;
constexpr int DZ_PITCH = 80;           // bytes per pixel of the staged dZ (32 channels = 64 B + pad)

__global__ void __launch_bounds__(THREADS) conv0_wgrad_kernel(const uint8_t* __restrict__ dz, const uint2* __restrict__ x,
                                                              float* __restrict__ dw, int N, int H, int W) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint2* xs = reinterpret_cast<uint2*>(smem);                                  // [(G_ROWS+2)][W+2]
  uint8_t* dzs = smem + ((size_t)(G_ROWS + 2) * (W + 2) + 2) * 8;               // [G_ROWS*W][DZ_PITCH]
  if (threadIdx.x < 2) xs[(G_ROWS + 2) * (W + 2) + threadIdx.x] = make_uint2(0u, 0u);
  __shared__ float s_dw[27 * C0_OUT];
  for (int i = threadIdx.x; i < 27 * C0_OUT; i += THREADS) s_dw[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[3][2][2][4];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { acc[kh][mt][nt][0] = acc[kh][mt][nt][1] = acc[kh][mt][nt][2] = acc[kh][mt][nt][3] = 0.f; }

  const int strips_per_img = H / G_ROWS, tiles_per_row = W / 16, pitch = W + 2;
  const long long total = (long long)N * strips_per_img;
  for (long long s = blockIdx.x; s < total; s += gridDim.x) {
    const int n = (int)(s / strips_per_img), h0 = (int)(s % strips_per_img) * G_ROWS;
    __syncthreads();
    // dZ rows: 64 bytes (32 channels) per pixel
    const uint8_t* src = dz + ((size_t)n * H + h0) * W * (C0_PAD * 2);
    for (int i = threadIdx.x; i < G_ROWS * W * 4; i += THREADS) {
      const int px = i >> 2, part = i & 3;
      cp_async16(dzs + (size_t)px * DZ_PITCH + part * 16, src + (size_t)px * (C0_PAD * 2) + part * 16);
    }
    stage_x(xs, x, n, h0 - 1, G_ROWS + 2, H, W);
    cp_async_wait_all();
    __syncthreads();
    for (int tile = warp; tile < G_ROWS * tiles_per_row; tile += WARPS) {
      const int hr = tile / tiles_per_row, w0 = (tile % tiles_per_row) * 16;
      // A = dZ^T: [m = co][k = pixel] from dzs[pixel][co] via ldmatrix.trans (two m-tiles of 16 channels)
      uint32_t a[2][4];
      {
        const int mi = lane >> 3, r = lane & 7;
        const int kpx = r + 8 * (mi >> 1), cooff = 8 * (mi & 1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t addr = (uint32_t)__cvta_generic_to_shared(dzs + (size_t)(hr * W + w0 + kpx) * DZ_PITCH +
                                                                    (16 * mt + cooff) * 2);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3])
                       : "r"(addr));
        }
      }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        // B[k = pixel][n = j]: element j of the 16-window starting at staged pixel (w0 + pixel)
        const __nv_bfloat16* rowp = reinterpret_cast<const __nv_bfloat16*>(xs + (size_t)(hr + kh) * pitch + w0);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int j = nt * 8 + g;
          const uint16_t e0 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t) * 4 + j);
          const uint16_t e1 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 1) * 4 + j);
          const uint16_t e2 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 8) * 4 + j);
          const uint16_t e3 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 9) * 4 + j);
          const uint32_t b0 = (uint32_t)e0 | ((uint32_t)e1 << 16), b1 = (uint32_t)e2 | ((uint32_t)e3 << 16);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(acc[kh][mt][nt], a[mt], b0, b1);
        }
      }
    }
  }
  // C fragment: (co = 16mt + g (+8), j = 8nt + 2t (+1)), j = kw*4 + ci
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = 16 * mt + g + 8 * (e >> 1), j = 8 * nt + 2 * t + (e & 1), kw = j >> 2, ci = j & 3;
          if (kw < 3 && ci < 3) atomicAdd(&s_dw[((co * 3 + ci) * 3 + kh) * 3 + kw], acc[kh][mt][nt][e]);
        }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * C0_OUT; i += THREADS) atomicAdd(&dw[i], s_dw[i]);
}


// =====================================================================================================
// Train-mode recompute path.  Layer 0 moves the largest tensors of the trunk (N x 224 x 224 x 32 bf16 = 2 GB at
// N = 640) for 0.6 % of its FLOPs, so its pre-activation z and its gradient dz are never stored:
//
//   forward  pass 1  conv0_train_kernel<FWD1>  x -> BatchNorm batch statistics of z (rounded to bf16, exactly what a
//                    stored z would hold) and Zw[co][ci][kh][kw] = sum_px z[px][co] * x[px+tap][ci]
//            pass 2  conv0_fwd_kernel (round_first)  x -> a = leaky(z*scale + shift), z recomputed
//   backward pass    conv0_train_kernel<BWD>   x, dA -> S1 = sum g, S2 = sum g*(z-mean), g = dA*leaky'(z*scale+shift),
//                    and Gw[co][ci][kh][kw] = sum_px g[px][co] * x[px+tap][ci]
//            finish  dW = scale*Gw + A*Zw + B*Xw,  dgamma += rstd*S2,  dbeta += S1
// because dz = scale*g + A*z + B (A, B per channel from S1, S2: bn_bwd_coef in trunk.cu) is linear in (g, z, 1) and
// the weight gradient is linear in dz.  Xw[ci][kh][kw] = sum_px x[px+tap][ci] comes from the total and the border
// sums of x (conv0_xsum_kernel).  HBM traffic: x three times + a once + dA once, instead of z, a, dA, dz twice each.
//
// Strips of output rows are double-buffered with cp.async: the rows of the next strip land while the warps work on
// the current one.
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem),
               "r"(n)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// rows [h_first, h_first+rows) of image n -> xs[r][1 + w] (the halo pixels xs[r][0], xs[r][W+1] are zeroed once by
// the caller); rows outside the image are zero-filled
__device__ __forceinline__ void stage_x_async(uint2* xs, const uint2* __restrict__ x, int n, int h_first, int rows,
                                              int H, int W) {
  const int pitch = W + 2;
  for (int i = threadIdx.x; i < rows * W; i += THREADS) {
    const int r = i / W, w = i - r * W, h = h_first + r;
    const bool ok = (h >= 0 && h < H);
    cp_async8_zfill(xs + r * pitch + 1 + w, ok ? x + ((size_t)n * H + h) * W + w : x, ok);
  }
}
__device__ __forceinline__ void zero_halo(uint2* xs, int rows, int W) {
  const int pitch = W + 2;
  for (int i = threadIdx.x; i < rows * 2; i += THREADS) xs[(i >> 1) * pitch + ((i & 1) ? W + 1 : 0)] = make_uint2(0u, 0u);
  if (threadIdx.x < 2) xs[rows * pitch + threadIdx.x] = make_uint2(0u, 0u);   // the last 16-window overruns by 2 pixels
}

// weight fragments B[k = kw*4 + ci][n = co] per filter row kh and n-tile nt (co = nt*8 + g)
__device__ __forceinline__ void load_wfrag(const float* __restrict__ w, int g, int t, uint32_t (&wf)[3][4][2]) {
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = 2 * t + 8 * r + e, kw = k >> 2, ci = k & 3, co = nt * 8 + g;
          v[e] = (kw < 3 && ci < 3) ? w[((co * 3 + ci) * 3 + kh) * 3 + kw] : 0.f;
        }
        wf[kh][nt][r] = pack_bf16(v[0], v[1]);
      }
}
// z tile: 16 pixels (row hr of the strip, starting at w0) x 32 channels
__device__ __forceinline__ void conv_tile(const uint2* xs, int pitch, int hr, int w0, int g, int t,
                                          const uint32_t (&wf)[3][4][2], float (&acc)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const uint8_t* rowp = reinterpret_cast<const uint8_t*>(xs + (size_t)(hr + kh) * pitch + w0);
    uint32_t a[4];
    a[0] = *reinterpret_cast<const uint32_t*>(rowp + g * 8 + t * 4);
    a[1] = *reinterpret_cast<const uint32_t*>(rowp + (g + 8) * 8 + t * 4);
    a[2] = *reinterpret_cast<const uint32_t*>(rowp + g * 8 + t * 4 + 16);
    a[3] = *reinterpret_cast<const uint32_t*>(rowp + (g + 8) * 8 + t * 4 + 16);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a, wf[kh][nt][0], wf[kh][nt][1]);
  }
}
// D_kh[co][j] += M^T[co][pixel] . P_kh[pixel][j]: M (16 pixels x 32 channels, bf16) sits in the warp's staging
// buffer `stg` (pixel pitch OUT_PITCH); P_kh are the 16-windows of the staged input rows
__device__ __forceinline__ void wgrad_tile(const uint8_t* stg, const uint2* xs, int pitch, int hr, int w0, int lane,
                                           float (&acc)[3][2][2][4]) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t a[2][4];
  {
    const int mi = lane >> 3, r = lane & 7;
    const int kpx = r + 8 * (mi >> 1), cooff = 8 * (mi & 1);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t addr = (uint32_t)__cvta_generic_to_shared(stg + (size_t)kpx * OUT_PITCH + (16 * mt + cooff) * 2);
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3])
                   : "r"(addr));
    }
  }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const __nv_bfloat16* rowp = reinterpret_cast<const __nv_bfloat16*>(xs + (size_t)(hr + kh) * pitch + w0);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int j = nt * 8 + g;
      const uint16_t e0 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t) * 4 + j);
      const uint16_t e1 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 1) * 4 + j);
      const uint16_t e2 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 8) * 4 + j);
      const uint16_t e3 = *reinterpret_cast<const uint16_t*>(rowp + (2 * t + 9) * 4 + j);
      const uint32_t b0 = (uint32_t)e0 | ((uint32_t)e1 << 16), b1 = (uint32_t)e2 | ((uint32_t)e3 << 16);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(acc[kh][mt][nt], a[mt], b0, b1);
    }
  }
}

constexpr int TR_ROWS = 2;             // output rows per strip of the recompute kernels

// BWD = false: forward pass 1 (statistics of z + Zw).  BWD = true: backward pass (S1, S2 + Gw).
//   sums [2][32] f64 (+=), wacc [32*27] fp32 (+=, nn.Conv2d weight layout)
template <bool BWD>
__global__ void __launch_bounds__(THREADS, 2) conv0_train_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                                 const uint8_t* __restrict__ da,
                                                                 const float* __restrict__ bn_scale,
                                                                 const float* __restrict__ bn_shift,
                                                                 const float* __restrict__ bn_mean, float slope, int N,
                                                                 int H, int W, double* __restrict__ sums,
                                                                 float* __restrict__ wacc) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pitch = W + 2;
  const size_t xs_bytes = ((size_t)(TR_ROWS + 2) * pitch + 2) * 8;
  const size_t da_bytes = BWD ? (size_t)TR_ROWS * W * DZ_PITCH : 0;
  const size_t buf_bytes = (xs_bytes + da_bytes + 15) / 16 * 16;
  uint8_t* stg_all = smem + 2 * buf_bytes;                                     // [WARPS][16][OUT_PITCH]
  __shared__ float s_acc[27 * C0_OUT];
  __shared__ float s_stat[WARPS][2 * C0_OUT];
  __shared__ float4 s_coef[C0_OUT];          // (scale, shift, mean, -) per channel: one 16-byte load each
  for (int i = threadIdx.x; i < 27 * C0_OUT; i += THREADS) s_acc[i] = 0.f;
  if (BWD)
    for (int i = threadIdx.x; i < C0_OUT; i += THREADS) s_coef[i] = make_float4(bn_scale[i], bn_shift[i], bn_mean[i], 0.f);
  for (int b = 0; b < 2; ++b) zero_halo(reinterpret_cast<uint2*>(smem + b * buf_bytes), TR_ROWS + 2, W);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t wf[3][4][2];
  load_wfrag(w, g, t, wf);
  float wa[3][2][2][4];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { wa[kh][mt][nt][0] = wa[kh][mt][nt][1] = wa[kh][mt][nt][2] = wa[kh][mt][nt][3] = 0.f; }
  float st1[4][2], st2[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { st1[nt][0] = st1[nt][1] = st2[nt][0] = st2[nt][1] = 0.f; }

  const int strips_per_img = H / TR_ROWS, tiles_per_row = W / 16;
  const long long total = (long long)N * strips_per_img;
  uint8_t* my_stg = stg_all + warp * 16 * OUT_PITCH;
  auto issue = [&](long long s, int b) {
    const int n = (int)(s / strips_per_img), h0 = (int)(s % strips_per_img) * TR_ROWS;
    uint8_t* base = smem + b * buf_bytes;
    stage_x_async(reinterpret_cast<uint2*>(base), x, n, h0 - 1, TR_ROWS + 2, H, W);
    if (BWD) {
      uint8_t* dzs = base + xs_bytes;
      const uint8_t* src = da + ((size_t)n * H + h0) * W * (C0_PAD * 2);
      for (int i = threadIdx.x; i < TR_ROWS * W * 4; i += THREADS) {
        const int px = i >> 2, part = i & 3;
        cp_async16(dzs + (size_t)px * DZ_PITCH + part * 16, src + (size_t)px * (C0_PAD * 2) + part * 16);
      }
    }
  };
  __syncthreads();                                     // halos / accumulators initialised
  long long s = blockIdx.x;
  int buf = 0;
  if (s < total) issue(s, 0);
  cp_async_commit();
  for (; s < total; s += gridDim.x, buf ^= 1) {
    const long long s2 = s + gridDim.x;
    if (s2 < total) issue(s2, buf ^ 1);                // the previous iteration's trailing barrier freed that buffer
    cp_async_commit();
    cp_async_wait_1();                                 // this strip's group has landed (per thread) ...
    __syncthreads();                                   // ... for every thread
    const uint2* xs = reinterpret_cast<const uint2*>(smem + buf * buf_bytes);
    const uint8_t* dzs = smem + buf * buf_bytes + xs_bytes;
    // this warp's 16-pixel tiles of the strip: (row hr, tile column tc), advanced without divisions
    int hr = 0, tc = warp;
    while (tc >= tiles_per_row) { tc -= tiles_per_row; ++hr; }
    for (; hr < TR_ROWS; tc += WARPS) {
      while (tc >= tiles_per_row) { tc -= tiles_per_row; ++hr; }
      if (hr >= TR_ROWS) break;
      const int w0 = tc * 16;
      float acc[4][4];
      conv_tile(xs, pitch, hr, w0, g, t, wf, acc);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        // z as stored: rounded to bf16.  lo = (pixel g; channels nt*8+2t, +1), hi = (pixel g+8; same channels)
        uint32_t lo = pack_bf16(acc[nt][0], acc[nt][1]), hi = pack_bf16(acc[nt][2], acc[nt][3]);
        const float z0 = __uint_as_float(lo << 16), z1 = __uint_as_float(lo & 0xFFFF0000u);
        const float z2 = __uint_as_float(hi << 16), z3 = __uint_as_float(hi & 0xFFFF0000u);
        if (!BWD) {
          st1[nt][0] += z0 + z2; st1[nt][1] += z1 + z3;
          st2[nt][0] = fmaf(z0, z0, fmaf(z2, z2, st2[nt][0]));
          st2[nt][1] = fmaf(z1, z1, fmaf(z3, z3, st2[nt][1]));
        } else {
          const int c0 = nt * 8 + 2 * t;
          const float4 k0 = s_coef[c0], k1 = s_coef[c0 + 1];
          const float sc0 = k0.x, sc1 = k1.x, sh0 = k0.y, sh1 = k1.y, mu0 = k0.z, mu1 = k1.z;
          const uint32_t dlo = *reinterpret_cast<const uint32_t*>(dzs + (size_t)(hr * W + w0 + g) * DZ_PITCH + c0 * 2);
          const uint32_t dhi = *reinterpret_cast<const uint32_t*>(dzs + (size_t)(hr * W + w0 + g + 8) * DZ_PITCH + c0 * 2);
          const float g0 = __uint_as_float(dlo << 16) * (fmaf(z0, sc0, sh0) > 0.f ? 1.f : slope);
          const float g1 = __uint_as_float(dlo & 0xFFFF0000u) * (fmaf(z1, sc1, sh1) > 0.f ? 1.f : slope);
          const float g2 = __uint_as_float(dhi << 16) * (fmaf(z2, sc0, sh0) > 0.f ? 1.f : slope);
          const float g3 = __uint_as_float(dhi & 0xFFFF0000u) * (fmaf(z3, sc1, sh1) > 0.f ? 1.f : slope);
          st1[nt][0] += g0 + g2; st1[nt][1] += g1 + g3;
          st2[nt][0] = fmaf(g0, z0 - mu0, fmaf(g2, z2 - mu0, st2[nt][0]));
          st2[nt][1] = fmaf(g1, z1 - mu1, fmaf(g3, z3 - mu1, st2[nt][1]));
          lo = pack_bf16(g0, g1); hi = pack_bf16(g2, g3);            // the tensor-core operand: g rounded to bf16
        }
        *reinterpret_cast<uint32_t*>(my_stg + g * OUT_PITCH + (nt * 8 + 2 * t) * 2) = lo;
        *reinterpret_cast<uint32_t*>(my_stg + (g + 8) * OUT_PITCH + (nt * 8 + 2 * t) * 2) = hi;
      }
      __syncwarp();
      wgrad_tile(my_stg, xs, pitch, hr, w0, lane, wa);
      __syncwarp();
    }
    __syncthreads();                                   // the strip is consumed: its buffer may be refilled
  }
  // ---- per-CTA reduction of the weight-shaped accumulator and of the per-channel sums ----
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int co = 16 * mt + g + 8 * (e >> 1), j = 8 * nt + 2 * t + (e & 1), kw = j >> 2, ci = j & 3;
          if (kw < 3 && ci < 3) atomicAdd(&s_acc[((co * 3 + ci) * 3 + kh) * 3 + kw], wa[kh][mt][nt][e]);
        }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = st1[nt][e], b = st2[nt][e];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (g == 0) { s_stat[warp][nt * 8 + 2 * t + e] = a; s_stat[warp][C0_OUT + nt * 8 + 2 * t + e] = b; }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * C0_OUT; i += THREADS) atomicAdd(&wacc[i], s_acc[i]);
  if (threadIdx.x < 2 * C0_OUT) {
    const int which = threadIdx.x / C0_OUT, c = threadIdx.x % C0_OUT;
    float v = 0.f;
#pragma unroll
    for (int wi = 0; wi < WARPS; ++wi) v += s_stat[wi][threadIdx.x];
    atomicAdd(sums + which * C0_PAD + c, (double)v);
  }
}

// Forward pass that writes the activation: a = leaky(z*scale + shift), z recomputed (train mode pass 2: z rounded
// to bf16 first, exactly what a stored z would hold; eval mode: the fp32 accumulator).  Lean (no statistics, the
// affine coefficients in shared memory) so that four CTAs fit an SM, strips double-buffered with cp.async.
constexpr int AP_ROWS = 4;
__global__ void __launch_bounds__(THREADS, 4) conv0_apply_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                                 uint4* __restrict__ out, int N, int H, int W,
                                                                 const float* __restrict__ aff_scale,
                                                                 const float* __restrict__ aff_shift, float slope,
                                                                 int round_first) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int pitch = W + 2;
  const size_t buf_bytes = ((((size_t)(AP_ROWS + 2) * pitch + 2) * 8) + 15) / 16 * 16;
  uint8_t* stg_all = smem + 2 * buf_bytes;                                     // [WARPS][16][OUT_PITCH]
  __shared__ float2 s_aff[C0_OUT];                                              // (scale, shift) per channel
  for (int i = threadIdx.x; i < C0_OUT; i += THREADS) s_aff[i] = make_float2(aff_scale[i], aff_shift[i]);
  for (int b = 0; b < 2; ++b) zero_halo(reinterpret_cast<uint2*>(smem + b * buf_bytes), AP_ROWS + 2, W);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  uint32_t wf[3][4][2];
  load_wfrag(w, g, t, wf);
  const int strips_per_img = H / AP_ROWS, tiles_per_row = W / 16;
  const long long total = (long long)N * strips_per_img;
  uint8_t* my_stg = stg_all + warp * 16 * OUT_PITCH;
  auto issue = [&](long long s, int b) {
    const int n = (int)(s / strips_per_img), h0 = (int)(s % strips_per_img) * AP_ROWS;
    stage_x_async(reinterpret_cast<uint2*>(smem + b * buf_bytes), x, n, h0 - 1, AP_ROWS + 2, H, W);
  };
  __syncthreads();
  long long s = blockIdx.x;
  int buf = 0;
  if (s < total) issue(s, 0);
  cp_async_commit();
  for (; s < total; s += gridDim.x, buf ^= 1) {
    const long long s2 = s + gridDim.x;
    if (s2 < total) issue(s2, buf ^ 1);
    cp_async_commit();
    cp_async_wait_1();
    __syncthreads();
    const uint2* xs = reinterpret_cast<const uint2*>(smem + buf * buf_bytes);
    const int n = (int)(s / strips_per_img), h0 = (int)(s % strips_per_img) * AP_ROWS;
    int hr = 0, tc = warp;
    while (tc >= tiles_per_row) { tc -= tiles_per_row; ++hr; }
    for (; hr < AP_ROWS; tc += WARPS) {
      while (tc >= tiles_per_row) { tc -= tiles_per_row; ++hr; }
      if (hr >= AP_ROWS) break;
      const int w0 = tc * 16;
      float acc[4][4];
      conv_tile(xs, pitch, hr, w0, g, t, wf, acc);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 a0 = s_aff[nt * 8 + 2 * t], a1 = s_aff[nt * 8 + 2 * t + 1];
        float v[4] = {acc[nt][0], acc[nt][1], acc[nt][2], acc[nt][3]};
        if (round_first) {
          const uint32_t lo = pack_bf16(v[0], v[1]), hi = pack_bf16(v[2], v[3]);
          v[0] = __uint_as_float(lo << 16); v[1] = __uint_as_float(lo & 0xFFFF0000u);
          v[2] = __uint_as_float(hi << 16); v[3] = __uint_as_float(hi & 0xFFFF0000u);
        }
        const float y0 = fmaf(v[0], a0.x, a0.y), y1 = fmaf(v[1], a1.x, a1.y);
        const float y2 = fmaf(v[2], a0.x, a0.y), y3 = fmaf(v[3], a1.x, a1.y);
        *reinterpret_cast<uint32_t*>(my_stg + g * OUT_PITCH + (nt * 8 + 2 * t) * 2) =
            pack_bf16(fmaxf(y0, y0 * slope), fmaxf(y1, y1 * slope));
        *reinterpret_cast<uint32_t*>(my_stg + (g + 8) * OUT_PITCH + (nt * 8 + 2 * t) * 2) =
            pack_bf16(fmaxf(y2, y2 * slope), fmaxf(y3, y3 * slope));
      }
      __syncwarp();
      uint4* dst = out + (((size_t)n * H + h0 + hr) * W + w0) * (C0_PAD / 8);      // 16 pixels x 64 bytes, contiguous
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int qd = i * 32 + lane, px = qd >> 2, part = qd & 3;
        dst[qd] = *reinterpret_cast<const uint4*>(my_stg + px * OUT_PITCH + part * 16);
      }
      __syncwarp();
    }
    __syncthreads();
  }
}

// Totals and border sums of x over all images, per input channel (f64 [9][4]: total, top row, bottom row, left
// column, right column, corners TL, TR, BL, BR; component 3 unused): Xw[ci][kh][kw] follows by inclusion-exclusion.
__global__ void __launch_bounds__(256) conv0_xsum_kernel(const uint2* __restrict__ x, int N, int H, int W,
                                                         double* __restrict__ xs9) {
  __shared__ float s_part[9][3];
  if (threadIdx.x < 27) (&s_part[0][0])[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)N * H;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float tot[3] = {0.f, 0.f, 0.f}, rtop[3] = {0.f, 0.f, 0.f}, rbot[3] = {0.f, 0.f, 0.f};
  for (long long r = warp_id; r < rows; r += n_warps) {           // one image row per warp iteration
    const int hq = (int)(r % H);
    const bool top = (hq == 0), bot = (hq == H - 1);
    const uint2* xr = x + r * W;
    float rs[3] = {0.f, 0.f, 0.f};
    for (int wq = lane; wq < W; wq += 32) {
      const uint2 v = __ldg(xr + wq);
      const float c[3] = {__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u), __uint_as_float(v.y << 16)};
      rs[0] += c[0]; rs[1] += c[1]; rs[2] += c[2];
      if (wq == 0 || wq == W - 1) {
        const int col = (wq == 0) ? 3 : 4;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          atomicAdd(&s_part[col][k], c[k]);
          if (top) atomicAdd(&s_part[wq == 0 ? 5 : 6][k], c[k]);
          if (bot) atomicAdd(&s_part[wq == 0 ? 7 : 8][k], c[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      tot[k] += rs[k];
      if (top) rtop[k] += rs[k];
      if (bot) rbot[k] += rs[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = tot[k], b = rtop[k], c2 = rbot[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (lane == 0) { atomicAdd(&s_part[0][k], a); atomicAdd(&s_part[1][k], b); atomicAdd(&s_part[2][k], c2); }
  }
  __syncthreads();
  if (threadIdx.x < 27) atomicAdd(xs9 + (threadIdx.x / 3) * 4 + threadIdx.x % 3, (double)s_part[threadIdx.x / 3][threadIdx.x % 3]);
}

// dW = scale*Gw + A*Zw + B*Xw ; dgamma += rstd*S2 ; dbeta += S1      (one thread per weight element)
__global__ void conv0_bwd_finish_kernel(const double* __restrict__ sums, const float* __restrict__ gw,
                                        const float* __restrict__ zw, const double* __restrict__ xs9, double invR,
                                        const float* __restrict__ scale, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, float* __restrict__ dw,
                                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * C0_OUT) return;
  const int kw = i % 3, kh = (i / 3) % 3, ci = (i / 9) % 3, co = i / 27;
  const double sc = scale[co], rs = rstd[co], mu = mean[co];
  const double S1 = sums[co], S2 = sums[C0_PAD + co];
  const double A = -sc * rs * rs * S2 * invR;
  const double B = -sc * S1 * invR - A * mu;
  // input pixel = output pixel + (kh-1, kw-1): kh = 2 never reads the top row, kh = 0 never the bottom row, ...
  const int rex = (kh == 2) ? 1 : (kh == 0 ? 2 : 0), cex = (kw == 2) ? 3 : (kw == 0 ? 4 : 0);
  double xw = xs9[ci];
  if (rex) xw -= xs9[rex * 4 + ci];
  if (cex) xw -= xs9[cex * 4 + ci];
  if (rex && cex) xw += xs9[(5 + (rex - 1) * 2 + (cex - 3)) * 4 + ci];
  dw[i] += (float)(sc * (double)gw[i] + A * (double)zw[i] + B * xw);
  if (ci == 0 && kh == 0 && kw == 0) {
    dbeta[co] += (float)S1;
    dgamma[co] += (float)(rs * S2);
  }
}

}  // namespace

// ===================================================================== C ABI
// Which kernels serve the recompute path and the eval forward: 1 = tcgen05 (conv0_tc.cu), 0 = warp-level mma.sync
// (this file).  AVDN_CONV0_TC in the environment sets the initial value; avdn_conv0_set_tensor_path() changes it.
static int& conv0_tc_flag() {
  static int flag = [] {
    const char* e = getenv("AVDN_CONV0_TC");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return flag;
}
static bool conv0_use_tc(int N, int H, int W) {
  return conv0_tc_flag() && (long long)N * H * W <= avdn::CONV0_TC_MAX_PIXELS;
}
extern "C" int avdn_conv0_set_tensor_path(int on) {
  int& f = conv0_tc_flag();
  const int old = f;
  if (on >= 0) f = on ? 1 : 0;
  return old;
}

static int conv0_fwd_launch(const char* who, const void* x_nhwc4, const float* w, void* out, int N, int H, int W,
                            double* stats, const float* scale, const float* shift, float slope,
                            avdn_stream_t stream, int round_first = 0) {
  AVDN_REQUIRE(x_nhwc4 && w && out && N > 0 && H > 0 && W > 0, "%s: bad argument", who);
  if (W % 16 != 0 || H % F_ROWS != 0 || W > 2048)
    return avdn::set_err(AVDN_ERR_UNSUPPORTED, "%s: W %% 16 == 0, H %% 4 == 0, W <= 2048 required (%dx%d)", who, H, W);
  cudaStream_t s = avdn::to_cuda(stream);
  if (stats && cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C0_PAD, s) != cudaSuccess)
    return avdn::check_launch("avdn_conv0_fwd memset");
  const size_t smem = ((size_t)(F_ROWS + 2) * (W + 2) + 2) * 8 + (size_t)WARPS * 16 * OUT_PITCH;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (cudaFuncSetAttribute(conv0_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return avdn::check_launch("avdn_conv0_fwd smem attribute");
    attr = smem;
  }
  const long long strips = (long long)N * (H / F_ROWS);
  const long long cap = (long long)avdn::sm_count() * 6;
  conv0_fwd_kernel<<<(unsigned)(strips < cap ? strips : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x_nhwc4), w, reinterpret_cast<uint4*>(out), N, H, W, stats, scale, shift, slope,
      round_first);
  return avdn::check_launch(who);
}

extern "C" int avdn_conv0_fwd(const void* x_nhwc4, const float* w, void* z, int N, int H, int W, double* stats,
                              avdn_stream_t stream) {
  return conv0_fwd_launch("avdn_conv0_fwd", x_nhwc4, w, z, N, H, W, stats, nullptr, nullptr, 0.f, stream);
}

static int conv0_apply_launch(const char* who, const void* x_nhwc4, const float* w, const float* scale, const float* shift,
                              float slope, void* a, int N, int H, int W, int round_first, avdn_stream_t stream) {
  AVDN_REQUIRE(x_nhwc4 && w && a && scale && shift && N > 0 && H > 0 && W > 0, "%s: bad argument", who);
  if (W % 16 != 0 || H % AP_ROWS != 0 || W > 1024)
    return avdn::set_err(AVDN_ERR_UNSUPPORTED, "%s: W %% 16 == 0, H %% 4 == 0, W <= 1024 required (%dx%d)", who, H, W);
  const size_t buf_bytes = ((((size_t)(AP_ROWS + 2) * (W + 2) + 2) * 8) + 15) / 16 * 16;
  const size_t smem = 2 * buf_bytes + (size_t)WARPS * 16 * OUT_PITCH;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (cudaFuncSetAttribute(conv0_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return avdn::check_launch("conv0_apply_kernel smem attribute");
    attr = smem;
  }
  const long long strips = (long long)N * (H / AP_ROWS);
  const long long cap = (long long)avdn::sm_count() * 4;
  conv0_apply_kernel<<<(unsigned)(strips < cap ? strips : cap), THREADS, smem, avdn::to_cuda(stream)>>>(
      reinterpret_cast<const uint2*>(x_nhwc4), w, reinterpret_cast<uint4*>(a), N, H, W, scale, shift, slope, round_first);
  return avdn::check_launch(who);
}

extern "C" int avdn_conv0_fwd_eval(const void* x_nhwc4, const float* w, const float* scale, const float* shift,
                                   float slope, void* a, int N, int H, int W, avdn_stream_t stream) {
  if (conv0_use_tc(N, H, W)) {
    AVDN_REQUIRE(x_nhwc4 && w && a && scale && shift && N > 0 && H > 0 && W > 0, "avdn_conv0_fwd_eval: bad argument");
    return avdn::conv0_tc_apply(x_nhwc4, w, scale, shift, slope, a, nullptr, N, H, W, avdn::to_cuda(stream));
  }
  return conv0_apply_launch("avdn_conv0_fwd_eval", x_nhwc4, w, scale, shift, slope, a, N, H, W, 0, stream);
}

extern "C" int avdn_conv0_wgrad(const void* dz, const void* x_nhwc4, float* dw, int N, int H, int W,
                                avdn_stream_t stream) {
  AVDN_REQUIRE(dz && x_nhwc4 && dw && N > 0 && H > 0 && W > 0, "avdn_conv0_wgrad: bad argument");
  if (W % 16 != 0 || H % G_ROWS != 0 || W > 1024)
    return avdn::set_err(AVDN_ERR_UNSUPPORTED, "avdn_conv0_wgrad: W %% 16 == 0, H %% 2 == 0, W <= 1024 required (%dx%d)", H, W);
  const size_t smem = ((size_t)(G_ROWS + 2) * (W + 2) + 2) * 8 + (size_t)G_ROWS * W * DZ_PITCH;
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    if (cudaFuncSetAttribute(conv0_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return avdn::check_launch("avdn_conv0_wgrad smem attribute");
    attr = smem;
  }
  const long long strips = (long long)N * (H / G_ROWS);
  const long long cap = (long long)avdn::sm_count() * 4;
  conv0_wgrad_kernel<<<(unsigned)(strips < cap ? strips : cap), THREADS, smem, avdn::to_cuda(stream)>>>(
      reinterpret_cast<const uint8_t*>(dz), reinterpret_cast<const uint2*>(x_nhwc4), dw, N, H, W);
  return avdn::check_launch("avdn_conv0_wgrad");
}

// ---- train-mode recompute path (no stored z / dz) ----
static int conv0_train_launch(bool bwd, const void* x, const float* w, const void* da, const float* scale,
                              const float* shift, const float* mean, float slope, int N, int H, int W, double* sums,
                              float* wacc, cudaStream_t s) {
  if (W % 16 != 0 || H % TR_ROWS != 0 || W > 1024)
    return avdn::set_err(AVDN_ERR_UNSUPPORTED, "conv0 train path: W %% 16 == 0, H %% 2 == 0, W <= 1024 required (%dx%d)", H, W);
  const size_t xs_bytes = ((size_t)(TR_ROWS + 2) * (W + 2) + 2) * 8;
  const size_t da_bytes = bwd ? (size_t)TR_ROWS * W * DZ_PITCH : 0;
  const size_t buf_bytes = (xs_bytes + da_bytes + 15) / 16 * 16;
  const size_t smem = 2 * buf_bytes + (size_t)WARPS * 16 * OUT_PITCH;
  static size_t attr[2] = {0, 0};
  if (smem > 48 * 1024 && smem > attr[bwd]) {
    cudaError_t e = bwd ? cudaFuncSetAttribute(conv0_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(conv0_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return avdn::check_launch("conv0 train path smem attribute");
    attr[bwd] = smem;
  }
  const long long strips = (long long)N * (H / TR_ROWS);
  const long long cap = (long long)avdn::sm_count() * 2;
  const unsigned grid = (unsigned)(strips < cap ? strips : cap);
  if (bwd)
    conv0_train_kernel<true><<<grid, THREADS, smem, s>>>(reinterpret_cast<const uint2*>(x), w,
                                                         reinterpret_cast<const uint8_t*>(da), scale, shift, mean, slope,
                                                         N, H, W, sums, wacc);
  else
    conv0_train_kernel<false><<<grid, THREADS, smem, s>>>(reinterpret_cast<const uint2*>(x), w, nullptr, nullptr, nullptr,
                                                          nullptr, 0.f, N, H, W, sums, wacc);
  return avdn::check_launch(bwd ? "conv0_train_kernel<BWD>" : "conv0_train_kernel<FWD1>");
}

extern "C" int avdn_conv0_fwd_stats(const void* x_nhwc4, const float* w, int N, int H, int W, double* stats, float* zw,
                                    double* xs9, avdn_stream_t stream) {
  AVDN_REQUIRE(x_nhwc4 && w && stats && zw && xs9 && N > 0 && H > 0 && W > 0, "avdn_conv0_fwd_stats: bad argument");
  cudaStream_t s = avdn::to_cuda(stream);
  if (conv0_use_tc(N, H, W)) return avdn::conv0_tc_fwd_stats(x_nhwc4, w, N, H, W, stats, zw, xs9, s);
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C0_PAD, s) != cudaSuccess ||
      cudaMemsetAsync(zw, 0, sizeof(float) * 27 * C0_OUT, s) != cudaSuccess ||
      cudaMemsetAsync(xs9, 0, sizeof(double) * 9 * 4, s) != cudaSuccess)
    return avdn::check_launch("avdn_conv0_fwd_stats memset");
  const long long rows = (long long)N * H;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)avdn::sm_count() * 8;
  conv0_xsum_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(reinterpret_cast<const uint2*>(x_nhwc4), N, H,
                                                                           W, xs9);
  int r = avdn::check_launch("conv0_xsum_kernel");
  if (r) return r;
  return conv0_train_launch(false, x_nhwc4, w, nullptr, nullptr, nullptr, nullptr, 0.f, N, H, W, stats, zw, s);
}

extern "C" int avdn_conv0_fwd_apply(const void* x_nhwc4, const float* w, const float* scale, const float* shift,
                                    float slope, void* a, void* mask, int N, int H, int W, avdn_stream_t stream) {
  if (conv0_use_tc(N, H, W)) {
    AVDN_REQUIRE(x_nhwc4 && w && a && scale && shift && N > 0 && H > 0 && W > 0, "avdn_conv0_fwd_apply: bad argument");
    return avdn::conv0_tc_apply(x_nhwc4, w, scale, shift, slope, a, reinterpret_cast<uint32_t*>(mask), N, H, W,
                                avdn::to_cuda(stream));
  }
  return conv0_apply_launch("avdn_conv0_fwd_apply", x_nhwc4, w, scale, shift, slope, a, N, H, W, 1, stream);
}

extern "C" int avdn_conv0_bwd(const void* x_nhwc4, const float* w, const void* da, const void* mask, const float* scale,
                              const float* shift, const float* mean, const float* rstd, float slope, int N, int H, int W,
                              const float* zw, const double* xs9, double* sums, float* gw, float* dw, float* dgamma,
                              float* dbeta, avdn_stream_t stream) {
  AVDN_REQUIRE(x_nhwc4 && w && da && scale && shift && mean && rstd && zw && xs9 && sums && gw && dw && dgamma && dbeta &&
                   N > 0 && H > 0 && W > 0,
               "avdn_conv0_bwd: bad argument");
  cudaStream_t s = avdn::to_cuda(stream);
  if (conv0_use_tc(N, H, W)) {
    AVDN_REQUIRE(mask, "avdn_conv0_bwd: the tensor-core path needs the sign mask avdn_conv0_fwd_apply wrote");
    return avdn::conv0_tc_bwd(x_nhwc4, w, da, reinterpret_cast<const uint32_t*>(mask), scale, mean, rstd, slope, N, H, W,
                              zw, xs9, sums, gw, dw, dgamma, dbeta, s);
  }
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C0_PAD, s) != cudaSuccess ||
      cudaMemsetAsync(gw, 0, sizeof(float) * 27 * C0_OUT, s) != cudaSuccess)
    return avdn::check_launch("avdn_conv0_bwd memset");
  int r = conv0_train_launch(true, x_nhwc4, w, da, scale, shift, mean, slope, N, H, W, sums, gw, s);
  if (r) return r;
  conv0_bwd_finish_kernel<<<(27 * C0_OUT + 127) / 128, 128, 0, s>>>(sums, gw, zw, xs9, 1.0 / ((double)N * H * W), scale, mean,
                                                                    rstd, dw, dgamma, dbeta);
  return avdn::check_launch("conv0_bwd_finish_kernel");
}
