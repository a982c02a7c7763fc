// 3x3, stride-1, pad-1 convolution for the THIN layers of the trunk (<= 64 input channels, e.g. module_list.3:
// 32 -> 64 at 112 x 112; nn.Conv2d at src/models/dark_net.py:22-28) with halo-tile reuse of the input.
//
// In the general implicit-GEMM kernel (gemm.cu) every one of the nine taps of a 128-pixel tile is its own TMA box:
// the input travels L2 -> shared memory nine times, in 64-byte rows, and those layers are bound by that fill
// (38 B/clk/SM delivered, epilogue and filter traffic irrelevant).  Here an output tile is 8 pixels wide
// and 16 rows tall and the input arrives as THREE boxes of 8 x 18 pixels -- one per horizontal tap kw, shifted by
// kw - 1 pixels -- so that
//   * a box row (8 pixels x CIN bf16) is exactly one swizzle atom (8 rows of 64 / 128 bytes), and
//   * the A operand of tap (kh, kw) is box kw advanced by kh whole atoms: a plain K-major descriptor, no sub-atom
//     start address.
// The input is read 3.4x (3 boxes x 18/16 rows) instead of 9x, the whole filter bank (9 x COUT x CIN bf16) stays
// resident in shared memory, and the tile leaves through a 128-byte-swizzled staging slab and one TMA store; the
// BatchNorm batch statistics (sum, sum of squares of the bf16-rounded outputs per channel) are read back from that
// slab, as in gemm.cu.  One CTA = 128 threads = 128 TMEM lanes = the 128 pixels of a tile; two CTAs per SM overlap
// each other's load / UMMA / epilogue phases.
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv3_halo.cuh"

namespace {

using namespace avdn_tc;

constexpr int TW = 8, TH = 16;            // output tile: 8 pixels x 16 rows = 128 = UMMA M
constexpr int THREADS = 128;

template <int CIN, int COUT, int XBUFS>
struct Cfg {
  static constexpr int ROW = CIN * 2;                    // bytes per pixel = K-major row (64 or 128)
  static constexpr int ATOM = 8 * ROW;                   // 8 pixels = one box row = one swizzle atom
  static constexpr int XBOX = (TH + 2) * ATOM;           // one horizontal-tap box: 18 rows
  static constexpr int XBUF = 3 * XBOX;                  // the three boxes of a tile
  static constexpr int WTAP = COUT * ROW;                // filters of one tap: COUT rows
  static constexpr int WBYTES = 9 * WTAP;
  static constexpr int OUT_ROW = COUT * 2;               // bytes per output pixel (64 or 128)
  static constexpr int OUT_BYTES = 128 * OUT_ROW;
  static constexpr int NCH = OUT_ROW / 16;               // 16-byte chunks per output pixel
  static constexpr uint32_t LAYOUT = (CIN == 64) ? 2u : 4u;  // SWIZZLE_128B / SWIZZLE_64B
  static constexpr int SMEM = XBUFS * XBUF + WBYTES + 2 * OUT_BYTES + 1024;     // two staging slabs
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  static constexpr uint32_t TMEM_COLS = COUT < 32 ? 32 : COUT;
  static_assert((CIN == 32 || CIN == 64) && (COUT == 32 || COUT == 64), "thin layers only");
};

struct alignas(64) Params {
  CUtensorMap tmX, tmW, tmZ;
  int32_t tiles_x, tiles_y, n_tiles;
  double* stats;          // NULL or [2][COUT] f64, accumulated
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}
// byte offset of 16-byte chunk j of pixel row m in the staging slab, in the swizzle the TMA store expects
// (128-byte rows: chunk ^ (m & 7); 64-byte rows: chunk ^ ((m >> 1) & 3))
template <int OUT_ROW>
__device__ __forceinline__ uint32_t out_chunk(uint32_t m, uint32_t j) {
  return m * OUT_ROW + ((OUT_ROW == 128 ? (j ^ (m & 7u)) : (j ^ ((m >> 1) & 3u))) << 4);
}

// FLIP: tap (kh, kw) of the input patch multiplies filter tap 8 - (3 kh + kw) -- the data gradient of a stride-1
// 3x3 convolution is that convolution of dz with the filters mirrored (and w_d = [ci][tap][co] as the B operand).
// STATS: per-channel sum / sum of squares of the rounded outputs.  XBUFS: 2 = the next tile's boxes load under this
// tile's UMMAs and epilogue; 1 = they load under the epilogue only (128-byte pixels: three boxes are 54 KB).
template <int CIN, int COUT, bool FLIP, bool STATS, int XBUFS>
__global__ void __launch_bounds__(THREADS, 2) conv3_halo_kernel(const __grid_constant__ Params p) {
  using C = Cfg<CIN, COUT, XBUFS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xs = base, ws = xs + XBUFS * C::XBUF, outs = ws + C::WBYTES;
  __shared__ __align__(8) uint64_t bars[5];              // X buffer 0 / 1, filters, UMMAs of accumulator 0 / 1 done
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_x = smem_u32(&bars[0]), bar_w = bar_x + 16, bar_mma = bar_x + 24;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_x, 1); mbar_init(bar_x + 8, 1); mbar_init(bar_w, 1); mbar_init(bar_mma, 1); mbar_init(bar_mma + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(2 * C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;

  const int per_img = p.tiles_x * p.tiles_y;
  auto load_x = [&](int tile, uint32_t buf) {            // thread 0: the three boxes of a tile
    const int n = tile / per_img, r = tile - n * per_img, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    mbar_expect_tx(bar_x + 8 * buf, (uint32_t)C::XBUF);
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
      tma_load_4d<1>(xs + buf * C::XBUF + kw * C::XBOX, &p.tmX, bar_x + 8 * buf, 0, tx * TW + kw - 1, ty * TH - 1, n);
  };
  const int n_my = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (tid == 0 && n_my > 0) {
    mbar_expect_tx(bar_w, (uint32_t)C::WBYTES);
    for (int t = 0; t < 9; ++t) tma_load_4d<1>(ws + t * C::WTAP, &p.tmW, bar_w, t * CIN, 0, 0, 0);
    load_x(blockIdx.x, 0);
  }
  // statistics: thread owns 16-byte chunk tid % NCH (8 channels) of the rows tid / NCH + (128 / NCH) i
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;

  // Software pipeline over the CTA's tiles: iteration `it` issues the UMMAs of tile it into accumulator it & 1 and
  // then runs the epilogue of tile it-1 from the other accumulator while they execute; the boxes of tile it+1 are
  // in flight throughout (two X buffers) or from the moment the UMMAs of tile it have read the single buffer.
  for (int it = 0; it <= n_my; ++it) {
    if (it < n_my && tid == 0) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const uint32_t b = (XBUFS == 2) ? ((uint32_t)it & 1u) : 0u;
      const bool more = it + 1 < n_my;
      if (XBUFS == 2 && more) {
        // buffer (it+1) & 1 was read by the UMMAs of tile it-1
        if (it >= 1) mbar_wait(bar_mma + 8 * ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
        load_x(tile + (int)gridDim.x, b ^ 1u);
      }
      if (it == 0) mbar_wait(bar_w, 0);
      mbar_wait(bar_x + 8 * b, (uint32_t)(XBUFS == 2 ? (it >> 1) : it) & 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem + (uint32_t)(it & 1) * C::TMEM_COLS;
      uint32_t acc = 0u;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int kh = t / 3, kw = t - kh * 3;
        const uint64_t ad = make_smem_desc(xs + b * C::XBUF + kw * C::XBOX + kh * C::ATOM, 16, C::ATOM, C::LAYOUT);
        const uint64_t bd = make_smem_desc(ws + (FLIP ? 8 - t : t) * C::WTAP, 16, C::ATOM, C::LAYOUT);
#pragma unroll
        for (int k = 0; k < CIN / 16; ++k) {
          umma_bf16<1>(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), C::IDESC, acc);
          acc = 1u;
        }
      }
      tcgen05_commit<1>(bar_mma + 8 * (it & 1));
      if (XBUFS == 1 && more) {                          // one buffer: reload once these UMMAs have read it
        mbar_wait(bar_mma + 8 * (it & 1), (uint32_t)(it >> 1) & 1u);
        load_x(tile + (int)gridDim.x, 0);
      }
    }
    __syncwarp();
    if (it >= 1) {
      const int j = it - 1, tile = (int)blockIdx.x + j * (int)gridDim.x;
      mbar_wait(bar_mma + 8 * (j & 1), (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem + (uint32_t)(j & 1) * C::TMEM_COLS;
      // slab j & 1: the TMA store of tile j-2 has finished reading it (wait_group.read 1 below); a bulk store takes
      // ~2500 cycles until its source has been read, which one slab would put on every tile's critical path
      const uint32_t slab = outs + (uint32_t)(j & 1) * C::OUT_BYTES;
#pragma unroll
      for (int h = 0; h < COUT / 32; ++h) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((warp * 32u) << 16) + 32 * h, v);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          sts128(slab + out_chunk<C::OUT_ROW>(tid, (uint32_t)(4 * h + c)),
                 pack2(__uint_as_float(v[8 * c]), __uint_as_float(v[8 * c + 1])),
                 pack2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                 pack2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])),
                 pack2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
      }
      tcgen05_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      if (tid == 0) {
        const int n = tile / per_img, r = tile - n * per_img, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        tma_store_4d(&p.tmZ, slab, 0, tx * TW, ty * TH, n);
        tma_commit_group();
      }
      if (STATS) {
        const uint32_t cj = tid % C::NCH;
#pragma unroll
        for (int i = 0; i < C::NCH; ++i) {
          const uint32_t m = tid / C::NCH + (128u / C::NCH) * i;
          const uint4 v = lds128(slab + out_chunk<C::OUT_ROW>(m, cj));
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float a = __uint_as_float(w4[k] << 16), c2 = __uint_as_float(w4[k] & 0xFFFF0000u);
            s1[2 * k] += a; s2[2 * k] = fmaf(a, a, s2[2 * k]);
            s1[2 * k + 1] += c2; s2[2 * k + 1] = fmaf(c2, c2, s2[2 * k + 1]);
          }
        }
      }
      if (tid == 0) tma_wait_group_read<1>();            // the OTHER slab (tile j-1's store) may be rewritten
    }
    __syncthreads();                                     // next slab free; accumulator (it-1) & 1 drained before tile it+1
  }
  if (tid == 0) tma_wait_group_read<0>();
  __syncthreads();
  if (STATS && p.stats != nullptr) {
    // block reduction through the (now free) staging slab: [2][row groups][COUT] floats
    constexpr int G = 128 / C::NCH;
    float* s_red = reinterpret_cast<float*>(smem_raw + (outs - smem_u32(smem_raw)));
    const int j = tid % C::NCH, g = tid / C::NCH;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_red[g * COUT + j * 8 + k] = s1[k]; s_red[(G + g) * COUT + j * 8 + k] = s2[k]; }
    __syncthreads();
    if (tid < 2 * COUT) {
      const int which = tid / COUT, c = tid % COUT;
      float a = 0.f;
      for (int g2 = 0; g2 < G; ++g2) a += s_red[(which * G + g2) * COUT + c];
      atomicAdd(p.stats + which * COUT + c, (double)a);
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * C::TMEM_COLS) : "memory");
}

template <int CIN, int COUT, bool FLIP, bool STATS, int XBUFS>
int launch(const void* x, const void* w, void* z, int N, int H, int W, double* stats, cudaStream_t s) {
  using C = Cfg<CIN, COUT, XBUFS>;
  Params p;
  const int64_t dx[4] = {CIN, W, H, N}, sx[4] = {1, CIN, (int64_t)W * CIN, (int64_t)H * W * CIN};
  const int32_t bx[4] = {CIN, TW, TH + 2, 1};
  const int64_t K = 9ll * CIN;                           // w: [COUT rows][9 taps x CIN], tap-major
  const int64_t dw[4] = {K, COUT, 1, 1}, sw[4] = {1, K, K * COUT, K * COUT};
  const int32_t bw[4] = {CIN, COUT, 1, 1};
  const int64_t dz[4] = {COUT, W, H, N}, sz[4] = {1, COUT, (int64_t)W * COUT, (int64_t)H * W * COUT};
  const int32_t bz[4] = {COUT, TW, TH, 1};
  int r = avdn::encode_tensor_map_4d(x, 2, dx, sx, bx, &p.tmX);
  if (!r) r = avdn::encode_tensor_map_4d(w, 2, dw, sw, bw, &p.tmW);
  if (!r) r = avdn::encode_tensor_map_4d(z, 2, dz, sz, bz, &p.tmZ);
  if (r) return r;
  p.tiles_x = W / TW;
  p.tiles_y = H / TH;
  const long long nt = (long long)N * p.tiles_x * p.tiles_y;
  if (nt >= (1ll << 31)) return avdn::set_err(AVDN_ERR_UNSUPPORTED, "conv3_halo: too many tiles");
  p.n_tiles = (int32_t)nt;
  p.stats = stats;
  if (STATS && stats && cudaMemsetAsync(stats, 0, sizeof(double) * 2 * COUT, s) != cudaSuccess)
    return avdn::check_launch("conv3_halo stats memset");
  auto kfn = conv3_halo_kernel<CIN, COUT, FLIP, STATS, XBUFS>;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess)
      return avdn::check_launch("conv3_halo smem attribute");
    attr = true;
  }
  const long long cap = (long long)avdn::sm_count() * 2;
  kfn<<<(unsigned)(nt < cap ? nt : cap), THREADS, C::SMEM, s>>>(p);
  return avdn::check_launch("conv3_halo_kernel");
}

// ------------------------------------------------------------------ weight gradient
// dW[co][ci][kh][kw] += sum over pixels dz[p][co] * x[p + (kh-1, kw-1)][ci]  for the 32 -> 64 block: the pixel
// dimension is K.  Per 8 x 16-pixel tile: A = the dz tile, MN-major (one 128-byte row of 64 co per pixel, M = 64);
// B = the three horizontal-tap boxes of x read MN-major (one 64-byte row of 32 ci per pixel, 64-byte swizzle) as ONE
// N = 96 operand -- its three 32-wide MN atoms are the three boxes, LBO = one box apart -- advanced by kh whole atoms.
// 3 filter rows x 8 K-steps = 24 UMMAs (M 64, N 96, K 16) per tile into three accumulators (288 TMEM columns) that
// live for the whole kernel: no per-tile epilogue; a producer thread and an issuer thread run a 3-stage ring.
constexpr int WG_STAGES = 3;
struct alignas(64) WgParams {
  CUtensorMap tmX, tmDz;
  int32_t tiles_x, tiles_y, n_tiles;
  float* dw;              // [64][32][3][3] fp32, accumulated
};

__global__ void __launch_bounds__(THREADS, 1) conv3_halo_wgrad_kernel(const __grid_constant__ WgParams p) {
  constexpr int CIN = 32, COUT = 64;
  constexpr int XBOX = (TH + 2) * 8 * CIN * 2;           // 9216
  constexpr int XBUF = 3 * XBOX, DZ = 128 * COUT * 2;    // 27648 + 16384 per stage
  constexpr int STAGE = XBUF + DZ;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2 * WG_STAGES + 1];
  __shared__ uint32_t tmem_slot;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_free = bar_full + 8 * WG_STAGES, bar_done = bar_free + 8 * WG_STAGES;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_free + 8 * s, 1); }
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  const int per_img = p.tiles_x * p.tiles_y;
  const int n_my = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (tid == 0) {
    // ---- producer: the dz tile and the three x boxes of every tile of this CTA ----
    for (int it = 0; it < n_my; ++it) {
      const int s = it % WG_STAGES, use = it / WG_STAGES;
      if (use > 0) mbar_wait(bar_free + 8 * s, (uint32_t)(use - 1) & 1u);
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const int n = tile / per_img, r = tile - n * per_img, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
      const uint32_t sx = base + s * STAGE, sd = sx + XBUF;
      mbar_expect_tx(bar_full + 8 * s, (uint32_t)STAGE);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
        tma_load_4d<1>(sx + kw * XBOX, &p.tmX, bar_full + 8 * s, 0, tx * TW + kw - 1, ty * TH - 1, n);
      tma_load_4d<1>(sd, &p.tmDz, bar_full + 8 * s, 0, tx * TW, ty * TH, n);
    }
  } else if (tid == 32) {
    // ---- issuer ----
    // D f32, A = B = bf16, both MN-major; N = 96, M = 64
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(96 >> 3) << 17) |
                               ((uint32_t)(64 >> 4) << 24);
    for (int it = 0; it < n_my; ++it) {
      const int s = it % WG_STAGES, use = it / WG_STAGES;
      mbar_wait(bar_full + 8 * s, (uint32_t)use & 1u);
      tcgen05_fence_after();
      const uint32_t sx = base + s * STAGE, sd = sx + XBUF;
      const uint64_t ad = make_smem_desc(sd, 16384, 1024, 2);            // dz: 128-byte rows, 8-row groups 1 KB apart
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        // x: 64-byte rows, 8-row groups 512 B apart, the three 32-wide MN atoms (kw) one box apart
        const uint64_t bd = make_smem_desc(sx + kh * 512, XBOX, 512, 4);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16<1>(tmem + 96 * kh, ad + (uint64_t)(k * 128), bd + (uint64_t)(k * 64), IDESC, (it > 0 || k > 0) ? 1u : 0u);
      }
      tcgen05_commit<1>(bar_free + 8 * s);
      if (it == n_my - 1) tcgen05_commit<1>(bar_done);
    }
  }
  __syncwarp();
  if (n_my > 0) {
    mbar_wait(bar_done, 0);
    tcgen05_fence_after();
    // M = 64 accumulator: row co sits in lane (co % 16) + 32 * (co / 16); column 96 kh + 32 kw + ci
    const int co = (int)warp * 16 + (int)lane;
#pragma unroll
    for (int h = 0; h < 9; ++h) {
      uint32_t v[32];
      tmem_ld32(tmem + ((warp * 32u) << 16) + 32 * h, v);
      if (lane < 16) {
        const int kh = h / 3, kw = h % 3;
#pragma unroll
        for (int ci = 0; ci < 32; ++ci) atomicAdd(p.dw + ((co * CIN + ci) * 3 + kh) * 3 + kw, __uint_as_float(v[ci]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

namespace avdn {

static bool tile_ok(int H, int W) { return H % TH == 0 && W % TW == 0 && H >= TH && W >= TW; }

bool conv3_halo_supported(int H, int W, int Cin, int Cout) { return Cin == 32 && Cout == 64 && tile_ok(H, W); }
bool conv3_halo_dgrad_supported(int H, int W, int Cin, int Cout) { return Cin == 32 && Cout == 64 && tile_ok(H, W); }

int conv3_halo_fwd(const void* x, const void* wf, void* z, int N, int H, int W, int Cin, int Cout, double* stats,
                   cudaStream_t s) {
  if (!conv3_halo_supported(H, W, Cin, Cout))
    return set_err(AVDN_ERR_UNSUPPORTED, "conv3_halo: shape %dx%d, %d -> %d channels not covered", H, W, Cin, Cout);
  return stats ? launch<32, 64, false, true, 1>(x, wf, z, N, H, W, stats, s)
               : launch<32, 64, false, false, 1>(x, wf, z, N, H, W, nullptr, s);
}

// dx [N,H,W,Cin] = data gradient of the Cin -> Cout layer: a 3x3 convolution of dz [N,H,W,Cout] with the mirrored
// filters w_d [Cin][9][Cout]
int conv3_halo_dgrad(const void* dz, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout, cudaStream_t s) {
  if (!conv3_halo_dgrad_supported(H, W, Cin, Cout))
    return set_err(AVDN_ERR_UNSUPPORTED, "conv3_halo dgrad: shape %dx%d, %d -> %d channels not covered", H, W, Cin, Cout);
  return launch<64, 32, true, false, 1>(dz, wd, dx, N, H, W, nullptr, s);
}

int conv3_halo_wgrad(const void* dz, const void* x, float* dw, int N, int H, int W, int Cin, int Cout, cudaStream_t s) {
  if (!conv3_halo_supported(H, W, Cin, Cout))
    return set_err(AVDN_ERR_UNSUPPORTED, "conv3_halo wgrad: shape %dx%d, %d -> %d channels not covered", H, W, Cin, Cout);
  WgParams p;
  const int64_t dx[4] = {Cin, W, H, N}, sx[4] = {1, Cin, (int64_t)W * Cin, (int64_t)H * W * Cin};
  const int32_t bx[4] = {Cin, TW, TH + 2, 1};
  const int64_t dd[4] = {Cout, W, H, N}, sd[4] = {1, Cout, (int64_t)W * Cout, (int64_t)H * W * Cout};
  const int32_t bd[4] = {Cout, TW, TH, 1};
  int r = encode_tensor_map_4d(x, 2, dx, sx, bx, &p.tmX);
  if (!r) r = encode_tensor_map_4d(dz, 2, dd, sd, bd, &p.tmDz);
  if (r) return r;
  p.tiles_x = W / TW;
  p.tiles_y = H / TH;
  const long long nt = (long long)N * p.tiles_x * p.tiles_y;
  if (nt >= (1ll << 31)) return set_err(AVDN_ERR_UNSUPPORTED, "conv3_halo wgrad: too many tiles");
  p.n_tiles = (int32_t)nt;
  p.dw = dw;
  const int smem = WG_STAGES * (3 * (TH + 2) * 8 * 64 + 128 * 128) + 1024;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(conv3_halo_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return check_launch("conv3_halo wgrad smem attribute");
    attr = true;
  }
  const long long cap = sm_count();
  conv3_halo_wgrad_kernel<<<(unsigned)(nt < cap ? nt : cap), THREADS, smem, s>>>(p);
  return check_launch("conv3_halo_wgrad_kernel");
}

}  // namespace avdn

extern "C" int avdn_conv3x3_thin_wgrad(const void* dz, const void* x_nhwc, float* dw, int N, int H, int W, int Cin,
                                       int Cout, avdn_stream_t stream) {
  AVDN_REQUIRE(dz && x_nhwc && dw && N > 0, "avdn_conv3x3_thin_wgrad: bad argument");
  return avdn::conv3_halo_wgrad(dz, x_nhwc, dw, N, H, W, Cin, Cout, avdn::to_cuda(stream));
}

extern "C" int avdn_conv3x3_thin_fwd(const void* x_nhwc, const void* w_f, void* z, int N, int H, int W, int Cin,
                                     int Cout, double* stats, avdn_stream_t stream) {
  AVDN_REQUIRE(x_nhwc && w_f && z && N > 0, "avdn_conv3x3_thin_fwd: bad argument");
  return avdn::conv3_halo_fwd(x_nhwc, w_f, z, N, H, W, Cin, Cout, stats, avdn::to_cuda(stream));
}

extern "C" int avdn_conv3x3_thin_dgrad(const void* dz, const void* w_d, void* dx, int N, int H, int W, int Cin,
                                       int Cout, avdn_stream_t stream) {
  AVDN_REQUIRE(dz && w_d && dx && N > 0, "avdn_conv3x3_thin_dgrad: bad argument");
  return avdn::conv3_halo_dgrad(dz, w_d, dx, N, H, W, Cin, Cout, avdn::to_cuda(stream));
}

extern "C" int avdn_conv3x3_thin_supported(int H, int W, int Cin, int Cout) {
  return avdn::conv3_halo_supported(H, W, Cin, Cout) ? 1 : 0;
}
