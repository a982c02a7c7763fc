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
//                 39 of every row is 1.0.  The SAME 16 KB image is a K-major A operand [128 px x 48 k] for the
//                 convolution and an MN-major operand [K = 128 px x 48] for the pixel reductions.
//   convolution   Z[128 px x 32] = X . W^T : 3 UMMAs (M 128, N 32, K 16) into TMEM (pass 2 / eval only).
//   reductions    every sum over pixels the block needs is a Gram product on the same tile:
//                   pass 1   G  += X^T X  (48 x 48): sum z = W.G[:,39], sum z^2 = w^T G w, Zw = W.G, Xw = G[:,39]
//                   backward Gw += G'^T X (32 x 48), G' = da * leaky' from the sign mask pass 2 left (one bit per
//                            element): sum g' = Gw[:,39], sum g' z = <w, Gw>, and Gw itself is the raw weight gradient
//                 8 UMMAs (M 64, N 48, K 16) per tile, accumulated in TMEM over the CTA's tiles and flushed to f64
//                 every FLUSH tiles.  Neither reduction pass has a per-pixel epilogue, and the backward never
//                 recomputes z.
//
// All three kernels are bound by shared-memory bandwidth (LSU wavefronts + the tensor core's operand reads add up
// to ~100 % in ncu), so the data movement is what is trimmed: BatchNorm coefficients sit in the constant bank, the
// activation leaves through a linear staging slab and one bulk copy per warp, da arrives by bulk copy, and the next
// tile's gather is in flight while the current one is processed.
//
// Per step at B = 64 (32.1 M pixels): pass 1 reads x (257 MB), pass 2 reads x and writes a (2.06 GB) + the mask
// (128 MB), the backward reads x, da (2.06 GB) and the mask.
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv0_tc.cuh"

namespace {

using namespace avdn_tc;

constexpr int C0 = 32;                 // output channels (= stored channels)
constexpr int TILE = 128;              // pixels per tile = UMMA M of the convolution, K of the reductions
constexpr int TILE_BYTES = TILE * 128; // 16 KB: 128 rows of 64 bf16
constexpr int THREADS = 128;           // one thread per pixel of the tile; warp w owns TMEM lanes 32w..32w+31
constexpr int NE = 36;                 // patch elements that carry data (9 taps x 4 channels, the 4th is zero)
constexpr int ONE = 39;                // the element of every patch row that is 1.0 (column sums)
constexpr int GN = 48;                 // columns of the Gram accumulators (elements 0..39 used)
constexpr int FLUSH = 64;              // tiles between two f64 flushes of a TMEM accumulator (8192 pixels in fp32)
constexpr int OUT_BYTES = TILE * 64;   // one tile of a / da: 128 pixels x 32 bf16

__constant__ float c_scale[C0];        // BatchNorm scale / shift of the running apply kernel (constant bank operands)
__constant__ float c_shift[C0];

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
// linear shared -> global / global -> shared bulk copies (16-byte granularity)
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

// instruction descriptors: D f32, A = B = bf16; bit 15 / 16 = A / B MN-major; N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC_CONV = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C0 >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
constexpr uint32_t IDESC_GRAM = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(GN >> 3) << 17) |
                                ((uint32_t)(64 >> 4) << 24);

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
// the same pixel slot one grid stride (gridDim.x tiles) further on, without divisions
struct Stride {
  int dp, dx, dy;         // gridDim.x * TILE pixels = dy' rows + dx pixels, dy = dy' % H
};
__device__ __forceinline__ Stride make_stride(int H, int W) {
  Stride st;
  st.dp = (int)gridDim.x * TILE;
  const int rows = st.dp / W;
  st.dx = st.dp - rows * W;
  st.dy = rows % H;
  return st;
}
__device__ __forceinline__ void advance(Pix& q, const Stride& st, long long P, int H, int W) {
  q.p += st.dp;
  q.valid = q.p < P;
  q.x += st.dx;
  const int carry = q.x >= W ? 1 : 0;
  q.x -= carry * W;
  q.y += st.dy + carry;
  q.y -= q.y >= H ? H : 0;
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
// patch row of this thread's pixel: chunks 0..4 = elements 0..39; element 39 is the ones column (0 for a pixel past
// the end, which therefore drops out of every sum).  Chunk 5 stays zero from the set-up, 6 and 7 are never read.
__device__ __forceinline__ void store_patch(uint32_t tile_addr, const uint2 (&v)[9], bool valid) {
  const uint32_t r = threadIdx.x;
  sts128(tile_addr + sw128(r, 0), v[0].x, v[0].y, v[1].x, v[1].y);
  sts128(tile_addr + sw128(r, 1), v[2].x, v[2].y, v[3].x, v[3].y);
  sts128(tile_addr + sw128(r, 2), v[4].x, v[4].y, v[5].x, v[5].y);
  sts128(tile_addr + sw128(r, 3), v[6].x, v[6].y, v[7].x, v[7].y);
  sts128(tile_addr + sw128(r, 4), v[8].x, v[8].y, 0u, valid ? 0x3F800000u : 0u);      // element 39 = bf16(1.0)
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

// D[64 x 48] (+)= A^T X over the 128 pixels of a tile: A and B = X are MN-major 64-column atoms (one 128-byte row per
// pixel): 8 UMMAs of K = 16 pixels.  `acc` = 0 overwrites the accumulator with the first.
// TMEM layout of an M = 64 accumulator: row m sits in lane (m % 16) + 32 * (m / 16).
__device__ __forceinline__ void gram_mma(uint32_t tmem_d, uint32_t a_addr, uint32_t x_addr, uint32_t acc) {
  const uint64_t ad = make_smem_desc(a_addr, TILE_BYTES, 1024), bd = make_smem_desc(x_addr, TILE_BYTES, 1024);
#pragma unroll
  for (int k = 0; k < TILE / 16; ++k) {
    umma1(tmem_d, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 128), IDESC_GRAM, acc);     // 16 rows x 128 B = 2 KB
    acc = 1u;
  }
}
// rows of an M = 64 accumulator this warp holds: lane l < 16 of warp w has row 16 w + l; columns 0..35 and ONE of it
// are added to dst[row][.] (row stride 64)
__device__ __forceinline__ void flush_rows(uint32_t tmem_d, double* __restrict__ dst, int n_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t lo[32], hi[32];
  tmem_ld32_async(tmem_d + ((uint32_t)(warp * 32) << 16), lo);            // columns 0..31
  tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + 16, hi);             // columns 16..47
  const int r = warp * 16 + lane;
  if (lane < 16 && r < n_rows && (r < NE || r == ONE || n_rows == C0)) {
    double* g = dst + r * 64;
#pragma unroll
    for (int c = 0; c < 32; ++c) atomicAdd(g + c, (double)__uint_as_float(lo[c]));
#pragma unroll
    for (int c = 32; c < NE; ++c) atomicAdd(g + c, (double)__uint_as_float(hi[c - 16]));
    atomicAdd(g + ONE, (double)__uint_as_float(hi[ONE - 16]));
  }
}
// Z = X . W^T : K = 48 covers the 36 patch elements (the weights of elements 36..47 are zero)
__device__ __forceinline__ void conv_mma(uint32_t tmem_d, uint32_t x_addr, uint32_t w_addr) {
  const uint64_t ad = make_smem_desc(x_addr, 16, 1024), bd = make_smem_desc(w_addr, 16, 1024);
#pragma unroll
  for (int k = 0; k < 3; ++k) umma1(tmem_d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC_CONV, k ? 1u : 0u);
}

// ------------------------------------------------------------------ pass 1: Gram matrix of the patches
// gram [64][64] f64 += sum over pixels x_patch x_patch^T (rows / columns < 36 and 39 are written).
__global__ void __launch_bounds__(THREADS, 4) conv0_tc_gram_kernel(const uint2* __restrict__ x, int H, int W, long long P,
                                                                   long long n_tiles, double* __restrict__ gram) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const uint32_t xs = smem_u32(sm);                    // 2 tiles of 16 KB
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int b = 0; b < 2; ++b) zero_chunks(xs + b * TILE_BYTES, 5, 7);
  const uint32_t tmem = tmem_alloc(&tmem_slot, 64);

  long long tile = blockIdx.x;
  Pix q = locate(tile, P, H, W);
  const Stride st = make_stride(H, W);
  uint2 v[9];
  if (tile < n_tiles) gather(x, q, H, W, v);
  uint32_t since_flush = 0;
  for (uint32_t it = 0; tile < n_tiles; ++it) {
    const uint32_t b = it & 1u;
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
    tile += gridDim.x;
    const bool last = tile >= n_tiles;
    if (!last) {                                                        // the next tile's loads fly under the UMMAs
      advance(q, st, P, H, W);
      gather(x, q, H, W, v);
    }
    if (since_flush == FLUSH || last) {
      mbar_wait(bar0 + 8 * b, (it >> 1) & 1u);                          // everything issued so far has completed
      tcgen05_fence_after();
      flush_rows(tmem, gram, ONE + 1);
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
// mask (NULL or [P] u32): bit 31 - c of mask[p] = (a[p][c] > 0), for the backward.
__global__ void __launch_bounds__(THREADS, 5) conv0_tc_apply_kernel(const uint2* __restrict__ x, const float* __restrict__ w,
                                                                    float slope, uint8_t* __restrict__ a,
                                                                    uint32_t* __restrict__ mask, int H, int W, long long P,
                                                                    long long n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t xs = smem_u32(sm), ws = xs + TILE_BYTES, outs = ws + C0 * 128;     // X 16 KB | W 4 KB | 2 x staging 8 KB
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  zero_chunks(xs, 5, 5);
  build_w_tile(sm + TILE_BYTES, w);
  const uint32_t tmem = tmem_alloc(&tmem_slot, 32);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rot = (uint32_t)(lane >> 1) & 3u;     // chunk rotation that makes the 64-byte-row stores conflict-free

  long long tile = blockIdx.x;
  Pix q = locate(tile, P, H, W);
  const Stride st = make_stride(H, W);
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
    const bool cur_valid = q.valid;
    tile += gridDim.x;
    if (tile < n_tiles) {                               // the next tile's loads fly under this tile's epilogue
      advance(q, st, P, H, W);
      gather(x, q, H, W, v);
    }
    mbar_wait(bar_a, it & 1u);
    tcgen05_fence_after();
    uint32_t z[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), z);
    tcgen05_fence_before();
    uint32_t o[16], sign = 0u;
#pragma unroll
    for (int c = 0; c < 32; c += 2) {
      float a0 = fmaf(__uint_as_float(z[c]), c_scale[c], c_shift[c]);
      float a1 = fmaf(__uint_as_float(z[c + 1]), c_scale[c + 1], c_shift[c + 1]);
      a0 = fmaxf(a0, a0 * slope);                       // leaky, 0 < slope < 1
      a1 = fmaxf(a1, a1 * slope);
      sign = __funnelshift_l(__float_as_uint(a0), sign, 1);
      sign = __funnelshift_l(__float_as_uint(a1), sign, 1);
      o[c >> 1] = pack2(a0, a1);
    }
    const long long pix = cur * TILE + threadIdx.x;
    if (mask != nullptr && cur_valid) mask[pix] = ~sign;
    // rotate the four 16-byte chunks of the pixel row left by `rot`: chunk (j + rot) & 3 ends up in slot j
    if (rot & 1u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t t = o[k];
        o[k] = o[4 + k]; o[4 + k] = o[8 + k]; o[8 + k] = o[12 + k]; o[12 + k] = t;
      }
    }
    if (rot & 2u) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint32_t t = o[k]; o[k] = o[8 + k]; o[8 + k] = t;
        t = o[4 + k]; o[4 + k] = o[12 + k]; o[12 + k] = t;
      }
    }
    // the warp's 32 pixels x 64 B go to its linear 2 KB slab and leave with one bulk copy
    const uint32_t slab = outs + (it & 1u) * OUT_BYTES + warp * 2048;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(slab + lane * 64 + ((((uint32_t)j + rot) & 3u) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const long long first = cur * TILE + warp * 32;
      long long n = P - first;
      n = n > 32 ? 32 : n;
      if (n > 0) bulk_store(a + first * 64, slab, (uint32_t)n * 64u);
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");    // the other slab (tile it-1) has been read
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tmem_free(tmem, 32);
}

// ------------------------------------------------------------------ backward
// gwacc [32][64] f64 += sum over pixels g'[co] * x_patch[e], g' = da where mask says a > 0, slope * da elsewhere.
__global__ void __launch_bounds__(THREADS, 4) conv0_tc_bwd_kernel(const uint2* __restrict__ x, const uint8_t* __restrict__ da,
                                                                  const uint32_t* __restrict__ mask, float slope, int H,
                                                                  int W, long long P, long long n_tiles,
                                                                  double* __restrict__ gwacc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = align1k(smem_raw);
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_slot;
  const uint32_t xs = smem_u32(sm), gs = xs + TILE_BYTES, ds = gs + TILE_BYTES;     // X 16 KB | G' 16 KB | 2 x da 8 KB
  const uint32_t bar_mma = smem_u32(&bars[0]), bar_da = bar_mma + 8;
  if (threadIdx.x == 0) {
    mbar_init(bar_mma, 1); mbar_init(bar_da, 1); mbar_init(bar_da + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  zero_chunks(xs, 5, 7);
  zero_chunks(gs, 4, 7);                               // channels 32..63 of the A atom: accumulator rows 32..63 = 0
  const uint32_t tmem = tmem_alloc(&tmem_slot, 64);
  const int lane = threadIdx.x & 31;
  const uint32_t rot = (uint32_t)(lane >> 1) & 3u;

  auto load_da = [&](long long t, uint32_t stage) {    // thread 0: one bulk copy of the tile's 128 x 64 B
    long long n = P - t * TILE;
    n = n > TILE ? TILE : n;
    mbar_expect_tx(bar_da + 8 * stage, (uint32_t)n * 64u);
    bulk_load(ds + stage * OUT_BYTES, da + t * TILE * 64, (uint32_t)n * 64u, bar_da + 8 * stage);
  };

  long long tile = blockIdx.x;
  Pix q = locate(tile, P, H, W);
  const Stride st = make_stride(H, W);
  uint2 v[9];
  uint32_t m = 0u;
  if (tile < n_tiles) {
    if (threadIdx.x == 0) load_da(tile, 0);
    gather(x, q, H, W, v);
    if (q.valid) m = __ldg(mask + q.p);
  }
  uint32_t since_flush = 0;
  for (uint32_t it = 0; tile < n_tiles; ++it) {
    const uint32_t s = it & 1u;
    if (it > 0) mbar_wait(bar_mma, (it - 1u) & 1u);    // the UMMAs of the previous tile have read X and G'
    store_patch(xs, v, q.valid);
    mbar_wait(bar_da + 8 * s, (it >> 1) & 1u);         // this tile's da has landed
    const uint32_t row = ds + s * OUT_BYTES + threadIdx.x * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (uint32_t)j ^ rot;            // permuted chunk order: neither the loads from the linear
                                                       // 64-byte rows nor the swizzled stores below conflict
      uint4 d = q.valid ? lds128(row + (c << 4)) : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t bits = m << (8u * c);             // bit 31 = channel 8c, bit 30 = channel 8c + 1, ...
      uint32_t dw_[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float g0 = __uint_as_float(dw_[k] << 16), g1 = __uint_as_float(dw_[k] & 0xFFFF0000u);
        const bool p0 = (bits << (2 * k)) & 0x80000000u, p1 = (bits << (2 * k + 1)) & 0x80000000u;
        dw_[k] = pack2(p0 ? g0 : g0 * slope, p1 ? g1 : g1 * slope);
      }
      sts128(gs + sw128(threadIdx.x, c), dw_[0], dw_[1], dw_[2], dw_[3]);
    }
    fence_proxy_async_smem();
    __syncthreads();
    tile += gridDim.x;
    const bool last = tile >= n_tiles;
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      gram_mma(tmem, gs, xs, since_flush ? 1u : 0u);
      tcgen05_commit<1>(bar_mma);
      if (!last) load_da(tile, s ^ 1u);                // stage s^1 was consumed before the barrier above
    }
    __syncwarp();
    ++since_flush;
    if (!last) {
      advance(q, st, P, H, W);
      gather(x, q, H, W, v);
      m = q.valid ? __ldg(mask + q.p) : 0u;
    }
    if (since_flush == FLUSH || last) {
      mbar_wait(bar_mma, it & 1u);
      tcgen05_fence_after();
      flush_rows(tmem, gwacc, C0);
      tcgen05_fence_before();
      __syncthreads();
      since_flush = 0;
    }
  }
  tmem_free(tmem, 64);
}

// dW = scale*Gw + A*Zw + B*Xw ; dgamma += rstd*S2 ; dbeta += S1, with S1 = sum g' = Gw[:,39] and
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

// per-device f64 scratch: the Gram matrix [64][64] / the backward accumulator [32][64]
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
  const size_t smem = 2 * TILE_BYTES + 1024;
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

int conv0_tc_apply(const void* x, const float* w, const float* scale, const float* shift, float slope, void* a,
                   uint32_t* mask, int N, int H, int W, cudaStream_t s) {
  if (!(slope > 0.f && slope < 1.f)) return set_err(AVDN_ERR_UNSUPPORTED, "conv0 tensor path: 0 < slope < 1 required");
  const long long P = (long long)N * H * W, n_tiles = (P + TILE - 1) / TILE;
  if (cudaMemcpyToSymbolAsync(c_scale, scale, sizeof(float) * C0, 0, cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
      cudaMemcpyToSymbolAsync(c_shift, shift, sizeof(float) * C0, 0, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
    return check_launch("conv0 tensor path: coefficient upload");
  const size_t smem = TILE_BYTES + C0 * 128 + 2 * OUT_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    if (set_smem(conv0_tc_apply_kernel, smem)) return check_launch("conv0_tc_apply_kernel smem attribute");
    attr = true;
  }
  const long long cap = (long long)sm_count() * 5;
  conv0_tc_apply_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x), w, slope, reinterpret_cast<uint8_t*>(a), mask, H, W, P, n_tiles);
  return check_launch("conv0_tc_apply_kernel");
}

int conv0_tc_bwd(const void* x, const float* w, const void* da, const uint32_t* mask, const float* scale,
                 const float* mean, const float* rstd, float slope, int N, int H, int W, const float* zw,
                 const double* xs9, double* sums, float* gw, float* dw, float* dgamma, float* dbeta, cudaStream_t s) {
  const long long P = (long long)N * H * W, n_tiles = (P + TILE - 1) / TILE;
  double* gwacc = scratch(s);
  if (!gwacc) return set_err(AVDN_ERR_LAUNCH, "conv0 tensor path: no scratch");
  const size_t smem = 2 * TILE_BYTES + 2 * OUT_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    if (set_smem(conv0_tc_bwd_kernel, smem)) return check_launch("conv0_tc_bwd_kernel smem attribute");
    attr = true;
  }
  const long long cap = (long long)sm_count() * 4;
  conv0_tc_bwd_kernel<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, s>>>(
      reinterpret_cast<const uint2*>(x), reinterpret_cast<const uint8_t*>(da), mask, slope, H, W, P, n_tiles, gwacc);
  int r = check_launch("conv0_tc_bwd_kernel");
  if (r) return r;
  conv0_tc_bwd_finish_kernel<<<(27 * C0 + 127) / 128, 128, 0, s>>>(gwacc, w, zw, xs9, 1.0 / (double)P, scale, mean, rstd,
                                                                   sums, gw, dw, dgamma, dbeta);
  return check_launch("conv0_tc_bwd_finish_kernel");
}

}  // namespace avdn
