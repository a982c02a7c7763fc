// Block 0 of the Darknet trunk (module_list.0: nn.Conv2d(3, 32, 3, pad 1) + nn.BatchNorm2d + nn.LeakyReLU,
// src/models/dark_net.py:22-33) on the tcgen05 tensor cores -- the three passes of the train-mode recompute path
// and the eval-mode forward (include/avdn.h: avdn_conv0_fwd_stats / _fwd_apply / _bwd / _fwd_eval).
//
// The block moves the trunk's largest tensors (640 x 224 x 224 pixels) for 0.6 % of its FLOPs: every pass should
// cost what HBM charges for the bytes it must touch.  The warp-level mma.sync kernels of conv0.cu are bound by
// instruction issue instead (256-438 warp instructions per 16 pixels of fragment loads, shuffles and
// accumulator handling); here the arithmetic is a handful of UMMA instructions per 128 pixels and the threads only
// move data:
//
//   im2col tile   one thread per pixel gathers its 3x3 patch (9 x 8-byte pixels R,G,B,0 -> 36 bf16, element
//                 e = (kh*3 + kw)*4 + c) into ONE 128-byte row of a 128-row shared-memory tile, written in the
//                 128-byte-swizzle layout tcgen05 reads (16-byte chunk j of row r at chunk j ^ (r & 7)); element
//                 63 of every row is 1.0.  The SAME 16 KB image is a K-major A operand [128 px x 64 k] for the
//                 convolution and an MN-major operand [K = 128 px x 64] for the pixel reductions.
//   convolution   Z[128 px x 32] = X . W^T : 3 UMMAs (M 128, N 32, K 16) into TMEM.
//   reductions    every sum over pixels the block needs is a Gram product on the same tile:
//                   pass 1   G  += X^T X  (64 x 64): sum z = W.G[:,63], sum z^2 = w^T G w, Zw = W.G, Xw = G[:,63]
//                   backward Gw += G'^T X (32 x 64), G' = da * leaky'(bn(z)) written to a second tile by the
//                            epilogue: sum g' = Gw[:,63], sum g' z = <w, Gw>, and Gw itself is the raw weight gradient
//                 8 UMMAs (M 128, N 64, K 16) per tile, accumulated in TMEM over the CTA's tiles and flushed to f64
//                 every FLUSH tiles.  Pass 1 therefore has NO per-pixel epilogue at all.
//
// Per step at B = 64 (32.1 M pixels): pass 1 reads x (257 MB), pass 2 reads x and writes a (2.06 GB), the
// backward reads x and da (2.06 GB).
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv0_tc.cuh"

namespace {

using namespace avdn_tc;

constexpr int C0 = 32;                 // output channels (= stored channels)
constexpr int TILE = 128;              // pixels per tile = UMMA M
constexpr int TILE_BYTES = TILE * 128; // 16 KB: 128 rows of 64 bf16
constexpr int THREADS = 128;           // one thread per pixel of the tile; warp w owns TMEM lanes 32w..32w+31
constexpr int NE = 36;                 // patch elements that carry data (9 taps x 4 channels, the 4th is zero)
constexpr int ONE = 63;                // the element of every patch row that is 1.0 (column sums)
constexpr int FLUSH = 64;              // tiles between two f64 flushes of a TMEM accumulator (8192 pixels in fp32)

// byte offset of 16-byte chunk `c` of row `r` in a 128-byte-swizzled tile of 128-byte rows
__device__ __forceinline__ uint32_t sw128(uint32_t r, uint32_t c) { return r * 128u + ((c ^ (r & 7u)) << 4); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void umma1(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  umma_bf16<1>(tmem_d, ad, bd, idesc, acc);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// instruction descriptors: D f32, A = B = bf16; bit 15 / 16 = A / B MN-major; N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC_CONV = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C0 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
constexpr uint32_t IDESC_GRAM = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                                ((uint32_t)(128 >> 4) << 24);

struct Pix {              // the pixel a thread owns in the current tile
  long long p;            // flat index n*H*W + y*W + x   (< 2^31: checked by the host)
  int y, x;
  bool valid;
};
__device__ __forceinline__ Pix locate(long long tile, long long P, int H, int W) {
  Pix q;
  q.p = tile * TILE + threadIdx.x;
  q.valid = q.p < P;
  const uint32_t p32 = (uint32_t)q.p, row = p32 / (uint32_t)W;
  q.x = (int)(p32 - row * (uint32_t)W);
  q.y = (int)(row % (uint32_t)H);
  return q;
}
// the 3x3 neighbourhood of the pixel (zero outside the image): v[kh*3 + kw] = x[y + kh - 1][x + kw - 1]
__device__ __forceinline__ void gather(const uint2* __restrict__ x, const Pix& q, int H, int W, uint2 (&v)[9]) {
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int y2 = q.y + kh - 1;
    const bool rok = q.valid && y2 >= 0 && y2 < H;
    const uint2* r = x + (q.p + (long long)(kh - 1) * W - 1);
    const uint2 z = make_uint2(0u, 0u);
    v[kh * 3 + 0] = (rok && q.x > 0) ? __ldg(r) : z;
    v[kh * 3 + 1] = rok ? __ldg(r + 1) : z;
    v[kh * 3 + 2] = (rok && q.x < W - 1) ? __ldg(r + 2) : z;
  }
}
// patch row of this thread's pixel: chunks 0..4 (elements 0..39) and chunk 7 (the ones column; 0 for a pixel
// past the end so that it drops out of every sum).  Chunks 5, 6 stay zero from the set-up.
__device__ __forceinline__ void store_patch(uint32_t tile_addr, const uint2 (&v)[9], bool valid) {
  const uint32_t r = threadIdx.x;
  sts128(tile_addr + sw128(r, 0), v[0].x, v[0].y, v[1].x, v[1].y);
  sts128(tile_addr + sw128(r, 1), v[2].x, v[2].y, v[3].x, v[3].y);
  sts128(tile_addr + sw128(r, 2), v[4].x, v[4].y, v[5].x, v[5].y);
  sts128(tile_addr + sw128(r, 3), v[6].x, v[6].y, v[7].x, v[7].y);
  sts128(tile_addr + sw128(r, 4), v[8].x, v[8].y, 0u, 0u);
  sts128(tile_addr + sw128(r, 7), 0u, 0u, 0u, valid ? 0x3F800000u : 0u);      // element 63 = bf16(1.0)
}
__device__ __forceinline__ void zero_chunks(uint32_t tile_addr, int c_lo, int c_hi) {
  for (int c = c_lo; c <= c_hi; ++c) sts128(tile_addr + sw128(threadIdx.x, c), 0u, 0u, 0u, 0u);
}
// W tile: K-major [32 co x 64 e] bf16, 128-byte swizzle; e = (kh*3 + kw)*4 + ci from w [co][ci][kh][kw] fp32
__device__ __forceinline__ void build_w_tile(uint8_t* wt, const float* __restrict__ w) {
  for (int i = threadIdx.x; i < C0 * 64; i += THREADS) {
    const int co = i >> 6, e = i & 63;
    const int tap = e >> 2, ci = e & 3;
    float v = 0.f;
    if (e < NE && ci < 3) v = w[co * 27 + ci * 9 + tap];
    *reinterpret_cast<__nv_bfloat16*>(wt + sw128(co, e >> 3) + (e & 7) * 2) = __float2bfloat16_rn(v);
  }
}
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot, uint32_t cols) {
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  return *slot;
}
__device__ __forceinline__ void tmem_free(uint32_t base, uint32_t cols) {
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ uint8_t* align1k(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
}

// G (+)= [A0 | A1]^T X over the 128 pixels of a tile: A = two MN-major atoms (64 columns each, 16 KB apart),
// B = X (MN-major, 64 columns): 8 UMMAs of K = 16 pixels.  `acc` = 0 overwrites the accumulator with the first.
__device__ __forceinline__ void gram_mma(uint32_t tmem_d, uint32_t a_addr, uint32_t x_addr, uint32_t acc) {
  const uint64_t ad = make_smem_desc(a_addr, TILE_BYTES, 1024), bd = make_smem_desc(x_addr, TILE_BYTES, 1024);
#pragma unroll
  for (int k = 0; k < TILE / 16; ++k) {
    umma1(tmem_d, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), IDESC_GRAM, acc);     // 16 rows x 128 B = 2 KB
    acc = 1u;
  }
}
// Z = X . W^T : K = 48 covers the 36 patch elements
__device__ __forceinline__ void conv_mma(uint32_t tmem_d, uint32_t x_addr, uint32_t w_addr) {
  const uint64_t ad = make_smem_desc(x_addr, 16, 1024), bd = make_smem_desc(w_addr, 16, 1024);
#pragma unroll
  for (int k = 0; k < 3; ++k) umma1(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_CONV, k ? 1u : 0u);
}

// ------------------------------------------------------------------ pass 1: Gram matrix of the patches
// gram [64][64] f64 += sum over pixels x_patch x_patch^T (rows / columns < 36 and 63 are written).
__global__ void __launch_bounds__(THREADS, 4) conv0_tc_gram_kernel(const uint2* __restrict__ x, int H, int W, long long P,
                                                                   long long n_tiles, double* __restrict__ gram) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t xs = smem_u32(sm);                    // [2 tiles][16 KB] + 16 KB that only the unused rows read
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int b = 0; b < 3; ++b) zero_chunks(xs + b * TILE_BYTES, b < 2 ? 5 : 0, b < 2 ? 6 : 7);
  const uint32_t tmem = tmem_alloc(&tmem_slot, 64);
  const int warp = threadIdx.x >> 5;

  uint32_t it = 0, since_flush = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const uint32_t b = it & 1u;
    const Pix q = locate(tile, P, H, W);
    uint2 v[9];
    gather(x, q, H, W, v);
    if (it >= 2) mbar_wait(bar0 + 8 * b, ((it >> 1) - 1u) & 1u);        // the UMMAs of tile it-2 have read buffer b
    store_patch(xs + b * TILE_BYTES, v, q.valid);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      gram_mma(tmem, xs + b * TILE_BYTES, xs + b * TILE_BYTES, since_flush ? 1u : 0u);
      tcgen05_commit<1>(bar0 + 8 * b);
    }
    __syncwarp();
    ++since_flush;
    const bool last = tile + gridDim.x >= n_tiles;
    if (since_flush == FLUSH || last) {
      mbar_wait(bar0 + 8 * b, (it >> 1) & 1u);                          // everything issued so far has completed
      tcgen05_fence_after();
      if (warp < 2) {                                                   // rows 0..63 of the accumulator
        uint32_t lo[32], hi[32];
        tmem_ld32_async(tmem + ((uint32_t)(warp * 32) << 16), lo);
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 32, hi);
        const int r = threadIdx.x;
        if (r < NE || r == ONE) {
          double* g = gram + r * 64;
#pragma unroll
          for (int c = 0; c < 32; ++c) atomicAdd(g + c, (double)__uint_as_float(lo[c]));
#pragma unroll
          for (int c = 0; c < NE - 32; ++c) atomicAdd(g + 32 + c, (double)__uint_as_float(hi[c]));
          atomicAdd(g + ONE, (double)__uint_as_float(hi[ONE - 32]));
        }
      }
      tcgen05_fence_before();
      __syncthreads();
      since_flush = 0;
    }
  }
  tmem_free(tmem, 64);
}

// stats [2][32] = (sum z, sum z^2), zw [32][3][3][3] = sum z * x_patch, xw [27] = sum x_patch -- all from the Gram matrix
__global__ void __launch_bounds__(256) conv0_tc_stats_finish_kernel(const double* __restrict__ gram, const float* __restrict__ w,
                                                                    double* __restrict__ stats, float* __restrict__ zw,
                                                                    double* __restrict__ xw) {
  __shared__ double s_w[C0][NE];           // bf16-rounded weights by patch element
  __shared__ double s_zw[C0][NE + 1];      // W . G  (last column: against the ones element)
  for (int i = threadIdx.x; i < C0 * NE; i += blockDim.x) {
    const int co = i / NE, e = i % NE, tap = e >> 2, ci = e & 3;
    s_w[co][e] = ci < 3 ? (double)bf16_round(w[co * 27 + ci * 9 + tap]) : 0.0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C0 * (NE + 1); i += blockDim.x) {
    const int co = i / (NE + 1), j = i % (NE + 1), e2 = j < NE ? j : ONE;
    double a = 0.0;
    for (int e = 0; e < NE; ++e) a += s_w[co][e] * gram[e * 64 + e2];
    s_zw[co][j] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C0 * 27; i += blockDim.x) {
    const int co = i / 27, ci = (i / 9) % 3, tap = i % 9;
    zw[i] = (float)s_zw[co][tap * 4 + ci];
  }
  if (threadIdx.x < 27) xw[threadIdx.x] = gram[((threadIdx.x % 9) * 4 + threadIdx.x / 9) * 64 + ONE];
  if (threadIdx.x < C0) {
    const int co = threadIdx.x;
    double a = 0.0;
    for (int e = 0; e < NE; ++e) a += s_w[co][e] * s_zw[co][e];
    stats[co] = s_zw[co][NE];
    stats[C0 + co] = a;
  }
}

// ------------------------------------------------------------------ pass 2 / eval: a = leaky(bn(conv(x)))
__global__ void __launch_bounds__(THREADS, 6) conv0_tc_apply_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift, float slope,
                                                                    uint4* __restrict__ a, int H, int W, long long P,
                                                                    long long n_tiles, int round_first) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float2 s_coef[C0];
  const uint32_t xs = smem_u32(sm), ws = xs + TILE_BYTES, outs = ws + C0 * 128;     // X 16 KB | W 4 KB | staging 8 KB
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < C0) s_coef[threadIdx.x] = make_float2(scale[threadIdx.x], shift[threadIdx.x]);
  zero_chunks(xs, 5, 6);
  build_w_tile(sm + TILE_BYTES, w);
  const uint32_t tmem = tmem_alloc(&tmem_slot, 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t my_out = outs + warp * 2048;          // this warp's 32 pixels x 64 B

  long long tile = blockIdx.x;
  Pix q = locate(tile, P, H, W);
  uint2 v[9];
  if (tile < n_tiles) gather(x, q, H, W, v);
  for (uint32_t it = 0; tile < n_tiles; ++it) {
    store_patch(xs, v, q.valid);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      conv_mma(tmem, xs, ws);
      tcgen05_commit<1>(bar_a);
    }
    __syncwarp();
    const long long cur = tile;
    tile += gridDim.x;
    if (tile < n_tiles) {                               // the next tile's loads fly under this tile's epilogue
      q = locate(tile, P, H, W);
      gather(x, q, H, W, v);
    }
    mbar_wait(bar_a, it & 1u);
    tcgen05_fence_after();
    uint32_t z[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), z);
    tcgen05_fence_before();
    uint32_t o[16];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float z0 = __uint_as_float(z[c]), z1 = __uint_as_float(z[c + 1]);
      if (round_first) { z0 = bf16_round(z0); z1 = bf16_round(z1); }
      const float2 k0 = s_coef[c], k1 = s_coef[c + 1];
      float a0 = fmaf(z0, k0.x, k0.y), a1 = fmaf(z1, k1.x, k1.y);
      a0 = a0 > 0.f ? a0 : a0 * slope;
      a1 = a1 > 0.f ? a1 : a1 * slope;
      o[c >> 1] = pack2(a0, a1);
    }
    // 64-byte pixel rows -> the warp's 2 KB staging slab (chunk c of row l at c ^ ((l >> 1) & 3)) -> 4 coalesced
    // 512-byte stores
    const uint32_t sx = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      sts128(my_out + lane * 64 + (((uint32_t)c ^ sx) << 4), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    __syncwarp();
    const long long base16 = (cur * TILE + warp * 32) * 4;          // in 16-byte units
    const long long end16 = P * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t off = (uint32_t)j * 512u + (uint32_t)lane * 16u;
      const uint32_t row = off >> 6, c = (off >> 4) & 3u;
      const uint4 val = lds128(my_out + row * 64 + ((c ^ ((row >> 1) & 3u)) << 4));
      const long long g = base16 + j * 32 + lane;
      if (g < end16) a[g] = val;
    }
    __syncwarp();
  }
  tmem_free(tmem, 32);
}

// ------------------------------------------------------------------ backward
// gwacc [32][64] f64 += sum over pixels g'[co] * x_patch[e], g' = da * leaky'(scale * bf16(z) + shift).
__global__ void __launch_bounds__(THREADS, 4) conv0_tc_bwd_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                                  const uint4* __restrict__ da,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, float slope, int H,
                                                                  int W, long long P, long long n_tiles,
                                                                  double* __restrict__ gwacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float2 s_coef[C0];
  const uint32_t xs = smem_u32(sm), gs = xs + TILE_BYTES, ws = gs + TILE_BYTES;     // X 16 KB | G' 16 KB | W 4 KB
  const uint32_t bar1 = smem_u32(&bars[0]), bar2 = bar1 + 8;
  if (threadIdx.x == 0) {
    mbar_init(bar1, 1); mbar_init(bar2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < C0) s_coef[threadIdx.x] = make_float2(scale[threadIdx.x], shift[threadIdx.x]);
  zero_chunks(xs, 5, 6);
  zero_chunks(gs, 4, 7);
  build_w_tile(sm + 2 * TILE_BYTES, w);
  const uint32_t tmem = tmem_alloc(&tmem_slot, 128);
  const uint32_t tmem_z = tmem, tmem_g = tmem + 64;
  const int warp = threadIdx.x >> 5;

  uint32_t it = 0, since_flush = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const Pix q = locate(tile, P, H, W);
    uint2 v[9];
    gather(x, q, H, W, v);
    uint4 d[4];
    {
      const uint4* dp = da + q.p * 4;
      const uint4 zz = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int c = 0; c < 4; ++c) d[c] = q.valid ? __ldg(dp + c) : zz;
    }
    if (it > 0) mbar_wait(bar2, (it - 1u) & 1u);      // the Gram UMMAs of the previous tile have read X and G'
    store_patch(xs, v, q.valid);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      conv_mma(tmem_z, xs, ws);
      tcgen05_commit<1>(bar1);
    }
    __syncwarp();
    mbar_wait(bar1, it & 1u);
    tcgen05_fence_after();
    uint32_t z[32];
    tmem_ld32(tmem_z + ((uint32_t)(warp * 32) << 16), z);
    tcgen05_fence_before();
    const uint32_t dw_[16] = {d[0].x, d[0].y, d[0].z, d[0].w, d[1].x, d[1].y, d[1].z, d[1].w,
                              d[2].x, d[2].y, d[2].z, d[2].w, d[3].x, d[3].y, d[3].z, d[3].w};
    uint32_t o[16];
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      const float z0 = bf16_round(__uint_as_float(z[c])), z1 = bf16_round(__uint_as_float(z[c + 1]));
      const float2 k0 = s_coef[c], k1 = s_coef[c + 1];
      const float g0 = __uint_as_float(dw_[c >> 1] << 16), g1 = __uint_as_float(dw_[c >> 1] & 0xFFFF0000u);
      o[c >> 1] = pack2(fmaf(z0, k0.x, k0.y) > 0.f ? g0 : g0 * slope, fmaf(z1, k1.x, k1.y) > 0.f ? g1 : g1 * slope);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) sts128(gs + sw128(threadIdx.x, c), o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      gram_mma(tmem_g, xs, xs, since_flush ? 1u : 0u);               // rows 64..95 of the accumulator = G'^T X
      tcgen05_commit<1>(bar2);
    }
    __syncwarp();
    ++since_flush;
    const bool last = tile + gridDim.x >= n_tiles;
    if (since_flush == FLUSH || last) {
      mbar_wait(bar2, it & 1u);
      tcgen05_fence_after();
      if (warp == 2) {                                                // TMEM lanes 64..95 = output channel
        uint32_t lo[32], hi[32];
        tmem_ld32_async(tmem_g + ((uint32_t)64 << 16), lo);
        tmem_ld32(tmem_g + ((uint32_t)64 << 16) + 32, hi);
        double* g = gwacc + (threadIdx.x - 64) * 64;
#pragma unroll
        for (int c = 0; c < 32; ++c) atomicAdd(g + c, (double)__uint_as_float(lo[c]));
#pragma unroll
        for (int c = 0; c < NE - 32; ++c) atomicAdd(g + 32 + c, (double)__uint_as_float(hi[c]));
        atomicAdd(g + ONE, (double)__uint_as_float(hi[ONE - 32]));
      }
      tcgen05_fence_before();
      __syncthreads();
      since_flush = 0;
    }
  }
  tmem_free(tmem, 128);
}

// dW = scale*Gw + A*Zw + B*Xw ; dgamma += rstd*S2 ; dbeta += S1, with S1 = sum g' = Gw[:,63] and
// S2 = sum g' (z - mean) = <w, Gw> - mean * S1                       (one thread per weight element)
__global__ void conv0_tc_bwd_finish_kernel(const double* __restrict__ gwacc, const float* __restrict__ w,
                                           const float* __restrict__ zw, const double* __restrict__ xw, double invR,
                                           const float* __restrict__ scale, const float* __restrict__ mean,
                                           const float* __restrict__ rstd, double* __restrict__ sums,
                                           float* __restrict__ gw, float* __restrict__ dw, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 27 * C0) return;
  const int co = i / 27, ci = (i / 9) % 3, tap = i % 9;
  const double* g = gwacc + co * 64;
  double gz = 0.0;
  for (int j = 0; j < 27; ++j) gz += (double)bf16_round(w[co * 27 + j]) * g[(j % 9) * 4 + j / 9];
  const double sc = scale[co], rs = rstd[co], mu = mean[co];
  const double S1 = g[ONE], S2 = gz - mu * S1;
  const double A = -sc * rs * rs * S2 * invR;
  const double B = -sc * S1 * invR - A * mu;
  const double gwi = g[tap * 4 + ci];
  gw[i] = (float)gwi;
  dw[i] += (float)(sc * gwi + A * (double)zw[i] + B * xw[ci * 9 + tap]);
  if (ci == 0 && tap == 0) {
    sums[co] = S1;
    sums[C0 + co] = S2;
    dbeta[co] += (float)S1;
    dgamma[co] += (float)(rs * S2);
  }
}

// per-device f64 scratch: the Gram matrix [64][64] and the backward accumulator [32][64]
double* scratch(cudaStream_t s) {
  static double* buf[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!buf[dev] && cudaMalloc(&buf[dev], sizeof(double) * 64 * 64) != cudaSuccess) return nullptr;
  if (cudaMemsetAsync(buf[dev], 0, sizeof(double) * 64 * 64, s) != cudaSuccess) return nullptr;
  return buf[dev];
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : 1;
}

}  // namespace

namespace avdn {

int conv0_tc_fwd_stats(const void* x, const float* w, int N, int H, int W, double* stats, float* zw, double* xs9,
                       cudaStream_t s) {
  const long long P = (long long)N * H * W, n_tiles = (P + TILE - 1) / TILE;
  double* gram = scratch(s);
  if (!gram) return set_err(AVDN_ERR_LAUNCH, "conv0 tensor path: no scratch");
  const size_t smem = 3 * TILE_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    if (set_smem(conv0_tc_gram_kernel, smem)) return check_launch("conv0_tc_gram_kernel smem attribute");
    attr = true;
  }
  const long long cap = (long long)sm_count() * 4;
  conv0_tc_gram_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x), H, W, P, n_tiles, gram);
  int r = check_launch("conv0_tc_gram_kernel");
  if (r) return r;
  conv0_tc_stats_finish_kernel<<<1, 256, 0, s>>>(gram, w, stats, zw, xs9);
  return check_launch("conv0_tc_stats_finish_kernel");
}

int conv0_tc_apply(const void* x, const float* w, const float* scale, const float* shift, float slope, void* a, int N,
                   int H, int W, int round_first, cudaStream_t s) {
  const long long P = (long long)N * H * W, n_tiles = (P + TILE - 1) / TILE;
  const size_t smem = TILE_BYTES + C0 * 128 + 4 * 2048 + 1024;
  static bool attr = false;
  if (!attr) {
    if (set_smem(conv0_tc_apply_kernel, smem)) return check_launch("conv0_tc_apply_kernel smem attribute");
    attr = true;
  }
  const long long cap = (long long)sm_count() * 6;
  conv0_tc_apply_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x), w, scale, shift, slope, reinterpret_cast<uint4*>(a), H, W, P, n_tiles,
      round_first);
  return check_launch("conv0_tc_apply_kernel");
}

int conv0_tc_bwd(const void* x, const float* w, const void* da, const float* scale, const float* shift,
                 const float* mean, const float* rstd, float slope, int N, int H, int W, const float* zw,
                 const double* xs9, double* sums, float* gw, float* dw, float* dgamma, float* dbeta, cudaStream_t s) {
  const long long P = (long long)N * H * W, n_tiles = (P + TILE - 1) / TILE;
  double* gwacc = scratch(s);
  if (!gwacc) return set_err(AVDN_ERR_LAUNCH, "conv0 tensor path: no scratch");
  const size_t smem = 2 * TILE_BYTES + C0 * 128 + 1024;
  static bool attr = false;
  if (!attr) {
    if (set_smem(conv0_tc_bwd_kernel, smem)) return check_launch("conv0_tc_bwd_kernel smem attribute");
    attr = true;
  }
  const long long cap = (long long)sm_count() * 4;
  conv0_tc_bwd_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x), w, reinterpret_cast<const uint4*>(da), scale, shift, slope, H, W, P, n_tiles,
      gwacc);
  int r = check_launch("conv0_tc_bwd_kernel");
  if (r) return r;
  conv0_tc_bwd_finish_kernel<<<(27 * C0 + 127) / 128, 128, 0, s>>>(gwacc, w, zw, xs9, 1.0 / (double)P, scale, mean, rstd,
                                                                   sums, gw, dw, dgamma, dbeta);
  return check_launch("conv0_tc_bwd_finish_kernel");
}

}  // namespace avdn
