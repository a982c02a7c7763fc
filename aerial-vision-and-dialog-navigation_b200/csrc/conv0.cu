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
                                                            const float* __restrict__ aff_shift, float slope) {
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

}  // namespace

// ===================================================================== C ABI
static int conv0_fwd_launch(const char* who, const void* x_nhwc4, const float* w, void* out, int N, int H, int W,
                            double* stats, const float* scale, const float* shift, float slope,
                            avdn_stream_t stream) {
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
      reinterpret_cast<const uint2*>(x_nhwc4), w, reinterpret_cast<uint4*>(out), N, H, W, stats, scale, shift, slope);
  return avdn::check_launch(who);
}

extern "C" int avdn_conv0_fwd(const void* x_nhwc4, const float* w, void* z, int N, int H, int W, double* stats,
                              avdn_stream_t stream) {
  return conv0_fwd_launch("avdn_conv0_fwd", x_nhwc4, w, z, N, H, W, stats, nullptr, nullptr, 0.f, stream);
}

extern "C" int avdn_conv0_fwd_eval(const void* x_nhwc4, const float* w, const float* scale, const float* shift,
                                   float slope, void* a, int N, int H, int W, avdn_stream_t stream) {
  AVDN_REQUIRE(scale && shift, "avdn_conv0_fwd_eval: scale/shift are required");
  return conv0_fwd_launch("avdn_conv0_fwd_eval", x_nhwc4, w, a, N, H, W, nullptr, scale, shift, slope, stream);
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
