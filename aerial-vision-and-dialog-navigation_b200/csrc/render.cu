// Stage 1 of the AVDN hot path: batched view rendering (src/env.py:254-332).
//
// Replaces, per pose, cv2.getPerspectiveTransform + two cv2.warpPerspective
// calls (map and attention map) + the agent's image normalisation
// (src/xview_et/agent.py:586-592) with three kernels:
//
//   pack_tile_kernel     one-off per map: BGR u8 HWC (+ attention) -> one 8-byte
//                        record per pixel holding the pixel AND the one below it
//                        (each B|G<<8|R<<16|ATT<<24), with a 1-px zero border
//   homography_kernel    one thread per pose: OpenCV's 8x8 LU + 3x3 adjugate
//                        inverse in float64, same operation order, no FMA
//   render_kernel        one CTA per (pose, 32-row band): fixed-point bilinear
//                        gather (INTER_BITS=5), bit-exact with OpenCV
//
// The warp is a gather: every output pixel reads a 2x2 footprint of the source
// at a data-dependent position.  It is bound by L2->SM sector traffic, not by
// tensor cores; the design rules that matter are (i) one aligned 64-bit load
// per tap COLUMN (both rows, all four channels at once): the 2x2 footprint is
// two adjacent records = 16 contiguous bytes, inside one 32-byte sector three
// times out of four (a row-major 4-byte tile needs 2.25 sectors per footprint),
// (ii) lanes of a warp arranged as an 8x4 output patch so that a warp's
// footprint stays compact under any rotation, (iii) the blend runs on packed
// 16-bit lanes (vertical) and dp2a (horizontal), (iv) the band is staged in
// shared memory as BGRA words and leaves the SM as full 16-byte coalesced
// stores (a 32-row band of a 224x224x3 view is one contiguous 21 504-byte
// range of HBM).
//
// Exactness: OpenCV evaluates, in float64 and per 64-pixel-wide block,
//   X0 = (M0*xb + M1*y) + M2 ... W = W0 + M6*x1 ; W = W ? 32/W : 0
//   X  = rint((X0 + M0*x1) * W)                                   
// A correctly rounded f64 division per pixel is the most expensive part, so the
// fast path uses a Newton reciprocal (error ~1e-15 relative) and extracts
// floor/fraction with a magic-number add; whenever the fractional part is
// within 2^-17 of a rounding tie (or anything is out of range) the lane
// re-evaluates with exactly OpenCV's sequence (__dmul_rn/__dadd_rn/__ddiv_rn,
// rint).  The fast path can therefore never round differently from OpenCV.
//
// This file is compiled with -fmad=false.
#include "common.cuh"

namespace {

constexpr int VIEW = AVDN_VIEW;       // 224
constexpr int BAND = 32;              // rows per CTA
constexpr int NBAND = VIEW / BAND;    // 7
constexpr int THREADS = 256;
constexpr int SPITCH = 232;            // shared-memory row pitch of a band, in pixels (words)
constexpr int INTER_BITS = 5;
constexpr int INTER_TAB = 1 << INTER_BITS;

// ---------------------------------------------------------------- pack tile
__device__ __forceinline__ uint32_t packed_pixel(const uint8_t* __restrict__ map_bgr,
                                                 const uint8_t* __restrict__ att, int att_ch,
                                                 int H, int W, int yy, int xx) {
  // (xx, yy) in padded coordinates: source pixel (xx-1, yy-1), zero outside
  if (yy < 1 || yy > H || xx < 1 || xx > W) return 0u;
  const long long s = (long long)(yy - 1) * W + (xx - 1);
  const uint8_t* p = map_bgr + s * 3;
  uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
  if (att) v |= (uint32_t)att[s * att_ch] << 24;
  return v;
}

__global__ void pack_tile_kernel(const uint8_t* __restrict__ map_bgr,
                                 const uint8_t* __restrict__ att, int att_ch,
                                 int H, int W, uint2* __restrict__ tile8) {
  const int pitch = W + 2;
  const long long n = (long long)(H + 1) * pitch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int yy = (int)(i / pitch), xx = (int)(i - (long long)yy * pitch);
    uint2 v;
    v.x = packed_pixel(map_bgr, att, att_ch, H, W, yy, xx);        // row yy
    v.y = packed_pixel(map_bgr, att, att_ch, H, W, yy + 1, xx);    // row yy + 1
    tile8[i] = v;
  }
}

// ------------------------------------------------------- map preparation (N2)
// cv2.resize(im, (new_w, H), INTER_AREA) for a horizontal shrink (src/env.py:221).  OpenCV's ResizeArea_
// accumulates src * alpha in float, entry by entry of the decimation table (built on the host in float64 exactly
// as computeResizeAreaTab does), then saturate_cast<uchar> (round half to even).  This file is compiled with
// -fmad=false: the multiply and the add round separately, as in OpenCV's scalar code.
__global__ void resize_area_width_kernel(const uint8_t* __restrict__ src, int H, int W, int new_w,
                                         const int32_t* __restrict__ ofs, const int32_t* __restrict__ sidx,
                                         const float* __restrict__ alpha, uint8_t* __restrict__ dst) {
  const long long n = (long long)H * new_w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / new_w), dx = (int)(i - (long long)y * new_w);
    const uint8_t* row = src + (size_t)y * W * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int e = ofs[dx]; e < ofs[dx + 1]; ++e) {
      const uint8_t* p = row + (size_t)sidx[e] * 3;
      const float w = alpha[e];
      a0 = __fadd_rn(a0, __fmul_rn((float)p[0], w));
      a1 = __fadd_rn(a1, __fmul_rn((float)p[1], w));
      a2 = __fadd_rn(a2, __fmul_rn((float)p[2], w));
    }
    uint8_t* o = dst + i * 3;
    o[0] = (uint8_t)min(255, max(0, __float2int_rn(a0)));
    o[1] = (uint8_t)min(255, max(0, __float2int_rn(a1)));
    o[2] = (uint8_t)min(255, max(0, __float2int_rn(a2)));
  }
}

// cv2.circle(att, center, radius, 255, thickness=-1) (src/env.py:226-230): OpenCV's Circle() walks the midpoint
// algorithm and fills rows cy +- dy with half-width dx and rows cy +- dx with half-width dy.  One thread per spot
// replays that walk into a half-width table hw[spot][0..r]; the raster kernel tests every pixel against the spots.
__global__ void circle_spans_kernel(const int32_t* __restrict__ spots, int n, int rmax, int32_t* __restrict__ hw) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int radius = spots[3 * s + 2];
  int32_t* t = hw + (size_t)s * (rmax + 1);
  for (int d = 0; d <= rmax; ++d) t[d] = -1;
  if (radius < 0 || radius > rmax) return;
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    t[dy] = max(t[dy], dx);
    t[dx] = max(t[dx], dy);
    ++dy;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

__global__ void raster_attention_kernel(const int32_t* __restrict__ spots, int n, int rmax,
                                        const int32_t* __restrict__ hw, int H, int W, int ch, uint8_t* __restrict__ att) {
  const long long np = (long long)H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    uint8_t v = 0;
    for (int s = 0; s < n; ++s) {
      const int cx = spots[3 * s], cy = spots[3 * s + 1], r = spots[3 * s + 2];
      const int d = abs(y - cy);
      if (d <= r && r <= rmax) {
        const int h = hw[(size_t)s * (rmax + 1) + d];
        if (h >= 0 && abs(x - cx) <= h) { v = 255; break; }
      }
    }
    for (int c = 0; c < ch; ++c) att[i * ch + c] = v;
  }
}

// ------------------------------------------------------------ gps -> pixels
__global__ void gps_to_pixels_kernel(const double* __restrict__ corners_gps,
                                     const double* __restrict__ geo, int P,
                                     int32_t* __restrict__ corners_px) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // corner index
  if (i >= P * 4) return;
  const int p = i >> 2;
  const double lat = corners_gps[2 * i], lng = corners_gps[2 * i + 1];
  const double bl_lng = geo[5 * p + 1], tr_lat = geo[5 * p + 2], ratio = geo[5 * p + 4];
  // src/env.py:196 — both axes use lat_ratio; round() is half-to-even
  corners_px[2 * i]     = __double2int_rn(__ddiv_rn(__dsub_rn(lng, bl_lng), ratio));
  corners_px[2 * i + 1] = __double2int_rn(__ddiv_rn(__dsub_rn(tr_lat, lat), ratio));
}

// --------------------------------------------------------------- homography
__global__ void homography_kernel(const int32_t* __restrict__ corners_px, int P,
                                  double* __restrict__ minv) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double a[8][8], b[8];
  const double dstx[4] = {0.0, VIEW - 1.0, VIEW - 1.0, 0.0};   // src/env.py:275-278
  const double dsty[4] = {0.0, 0.0, VIEW - 1.0, VIEW - 1.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // int corners -> float32 (src/env.py:284) -> float64 inside OpenCV
    const double sx = (double)(float)corners_px[p * 8 + 2 * i];
    const double sy = (double)(float)corners_px[p * 8 + 2 * i + 1];
    a[i][0] = a[i + 4][3] = sx;
    a[i][1] = a[i + 4][4] = sy;
    a[i][2] = a[i + 4][5] = 1.0;
    a[i][3] = a[i][4] = a[i][5] = a[i + 4][0] = a[i + 4][1] = a[i + 4][2] = 0.0;
    a[i][6] = __dmul_rn(-sx, dstx[i]);
    a[i][7] = __dmul_rn(-sy, dstx[i]);
    a[i + 4][6] = __dmul_rn(-sx, dsty[i]);
    a[i + 4][7] = __dmul_rn(-sy, dsty[i]);
    b[i] = dstx[i];
    b[i + 4] = dsty[i];
  }
  bool singular = false;
  const double eps = 2.220446049250313e-16 * 100;
  for (int i = 0; i < 8 && !singular; ++i) {
    int k = i;
    for (int j = i + 1; j < 8; ++j)
      if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
    if (fabs(a[k][i]) < eps) { singular = true; break; }
    if (k != i) {
      for (int j = i; j < 8; ++j) { double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
      double t = b[i]; b[i] = b[k]; b[k] = t;
    }
    const double d = __ddiv_rn(-1.0, a[i][i]);
    for (int j = i + 1; j < 8; ++j) {
      const double alpha = __dmul_rn(a[j][i], d);
      for (int kk = i + 1; kk < 8; ++kk)
        a[j][kk] = __dadd_rn(a[j][kk], __dmul_rn(alpha, a[i][kk]));
      b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
    }
  }
  double m[9];
  if (!singular) {
    for (int i = 7; i >= 0; --i) {
      double s = b[i];
      for (int kk = i + 1; kk < 8; ++kk) s = __dsub_rn(s, __dmul_rn(a[i][kk], b[kk]));
      b[i] = __ddiv_rn(s, a[i][i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = b[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = 0.0;     // cv::solve failure -> zero solution
  }
  m[8] = 1.0;
  // cv::invert, n == 3 closed form
#define MUL(x, y) __dmul_rn((x), (y))
#define SUB(x, y) __dsub_rn((x), (y))
  const double c0 = SUB(MUL(m[4], m[8]), MUL(m[5], m[7]));
  const double c1 = SUB(MUL(m[3], m[8]), MUL(m[5], m[6]));
  const double c2 = SUB(MUL(m[3], m[7]), MUL(m[4], m[6]));
  const double det = __dadd_rn(SUB(MUL(m[0], c0), MUL(m[1], c1)), MUL(m[2], c2));
  double* o = minv + (size_t)p * 9;
  if (det != 0.0) {
    const double d = __ddiv_rn(1.0, det);
    o[0] = MUL(SUB(MUL(m[4], m[8]), MUL(m[5], m[7])), d);
    o[1] = MUL(SUB(MUL(m[2], m[7]), MUL(m[1], m[8])), d);
    o[2] = MUL(SUB(MUL(m[1], m[5]), MUL(m[2], m[4])), d);
    o[3] = MUL(SUB(MUL(m[5], m[6]), MUL(m[3], m[8])), d);
    o[4] = MUL(SUB(MUL(m[0], m[8]), MUL(m[2], m[6])), d);
    o[5] = MUL(SUB(MUL(m[2], m[3]), MUL(m[0], m[5])), d);
    o[6] = MUL(SUB(MUL(m[3], m[7]), MUL(m[4], m[6])), d);
    o[7] = MUL(SUB(MUL(m[1], m[6]), MUL(m[0], m[7])), d);
    o[8] = MUL(SUB(MUL(m[0], m[4]), MUL(m[1], m[3])), d);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = 0.0;     // cv::invert failure -> zero matrix
  }
#undef MUL
#undef SUB
}

// ------------------------------------------------------------------- render
// Exactly OpenCV's per-pixel sequence (WarpPerspectiveInvoker), used by the
// rare lanes whose fast-path result sits next to a rounding tie.
__device__ __noinline__ int2 exact_coords(const double* __restrict__ m, int xb, int x1, int y) {
  const double dxb = (double)xb, dx1 = (double)x1, dy = (double)y;
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dxb), __dmul_rn(m[1], dy)), m[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dxb), __dmul_rn(m[4], dy)), m[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dxb), __dmul_rn(m[7], dy)), m[8]);
  double W = __dadd_rn(W0, __dmul_rn(m[6], dx1));
  W = (W != 0.0) ? __ddiv_rn((double)INTER_TAB, W) : 0.0;
  double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], dx1)), W);
  double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], dx1)), W);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  return make_int2(__double2int_rn(fX), __double2int_rn(fY));
}

__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

// v = f + 1.5*2^32 ; returns true when the fast result is trustworthy.
__device__ __forceinline__ bool fast_round(double f, int* out) {
  const double v = __dadd_rn(f, 6442450944.0);
  const uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  const uint32_t fl = __funnelshift_r(lo, hi, 20) ^ 0x80000000u;   // floor(f)
  const int frac = (int)(lo & 0xFFFFFu);                              // 20-bit fraction
  *out = (int)fl + (frac > 0x80000 ? 1 : 0);
  const int d = frac - 0x80000;
  // exponent of v must be that of [2^32, 2^33): rejects NaN/inf/|f| >= 2^31
  return ((hi >> 20) == 0x41Fu) && (d > 8 || d < -8);
}

// Blend of one output pixel from the two records of its footprint.
//   r0 = (p00 | p10 << 32) at column sx, r1 = (p01 | p11 << 32) at column sx + 1
// out_c = (sum_taps p_c * wy * wx + 512) >> 10  (SURVEY App. A), evaluated as a vertical
// blend on packed 16-bit lanes (two channels per IMAD; <= 255 * 32 fits 13 bits) followed by
// one dp2a per channel for the horizontal blend.  Pure integer refactoring: bit-exact.
template <bool ATT>
__device__ __forceinline__ uint32_t blend(uint2 r0, uint2 r1, uint32_t ax, uint32_t ay) {
  const uint32_t wy0 = INTER_TAB - ay, wy1 = ay;
  const uint32_t wx = (INTER_TAB - ax) | (ax << 8);
  const uint32_t vBR0 = __byte_perm(r0.x, 0, 0x4240) * wy0 + __byte_perm(r0.y, 0, 0x4240) * wy1;
  const uint32_t vGA0 = __byte_perm(r0.x, 0, 0x4341) * wy0 + __byte_perm(r0.y, 0, 0x4341) * wy1;
  const uint32_t vBR1 = __byte_perm(r1.x, 0, 0x4240) * wy0 + __byte_perm(r1.y, 0, 0x4240) * wy1;
  const uint32_t vGA1 = __byte_perm(r1.x, 0, 0x4341) * wy0 + __byte_perm(r1.y, 0, 0x4341) * wy1;
  const uint32_t oB = __dp2a_lo(__byte_perm(vBR0, vBR1, 0x5410), wx, 512u) >> 10;
  const uint32_t oR = __dp2a_lo(__byte_perm(vBR0, vBR1, 0x7632), wx, 512u) >> 10;
  const uint32_t oG = __dp2a_lo(__byte_perm(vGA0, vGA1, 0x5410), wx, 512u) >> 10;
  uint32_t o = oB | (oG << 8) | (oR << 16);
  if (ATT) o |= (__dp2a_lo(__byte_perm(vGA0, vGA1, 0x7632), wx, 512u) >> 10) << 24;
  return o;
}

// Word index of output pixel x inside a staged band row: the 16-byte chunks of every 16-pixel group are XOR-permuted
// by the group's index so that the repack below (a thread reads the four chunks of ONE group, neighbouring threads
// neighbouring groups: a 64-byte stride) is free of bank conflicts; the writers' pattern is unaffected.
__device__ __forceinline__ int swz_px(int x) { return (((x >> 2) ^ ((x >> 5) & 3)) << 2) | (x & 3); }

// 64-bit gather from the packed tile with an L2 evict-last policy: the tile (72 MB for a 3000 x 3000 map) is re-read
// by every pose while 0.6 GB of views stream out through the same L2
__device__ __forceinline__ uint2 ldg_tile(const uint2* p, uint64_t policy) {
  uint2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
  return v;
}

template <bool ATT>
__global__ void __launch_bounds__(THREADS, 4)
render_kernel(const avdn_tile_desc* __restrict__ tiles, const int32_t* __restrict__ tile_idx,
              const double* __restrict__ minv, int P,
              uint8_t* __restrict__ views, uint8_t* __restrict__ att,
              float* __restrict__ norm_nchw, __nv_bfloat16* __restrict__ norm_nhwc,
              const float* __restrict__ norm_lut) {
  __shared__ double s_m[9];
  __shared__ double s_tab[4][BAND][3];                       // 32*X0, 32*Y0, W0
  // BGRA per pixel; row pitch 232 words: the 4 rows of a warp's 8x4 patch fall in distinct banks
  __shared__ __align__(16) uint32_t s_px[BAND * SPITCH];
  __shared__ float s_lut[3 * 256];

  const int p = blockIdx.x / NBAND, band = blockIdx.x - p * NBAND;
  const int t = threadIdx.x;
  if (t < 9) s_m[t] = minv[(size_t)p * 9 + t];
  const bool want_norm = (norm_nchw != nullptr) || (norm_nhwc != nullptr);
  if (want_norm)
    for (int i = t; i < 768; i += THREADS) s_lut[i] = norm_lut[i];
  const avdn_tile_desc td = tiles[tile_idx ? tile_idx[p] : 0];
  __syncthreads();
  if (t < 4 * BAND) {
    const int xbi = t >> 5, r = t & 31;
    const double dxb = (double)(xbi * 64), dy = (double)(band * BAND + r);
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(s_m[0], dxb), __dmul_rn(s_m[1], dy)), s_m[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(s_m[3], dxb), __dmul_rn(s_m[4], dy)), s_m[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(s_m[6], dxb), __dmul_rn(s_m[7], dy)), s_m[8]);
    s_tab[xbi][r][0] = X0 * 32.0;     // exact power-of-two scaling
    s_tab[xbi][r][1] = Y0 * 32.0;
    s_tab[xbi][r][2] = W0;
  }
  __syncthreads();

  const int warp = t >> 5, lane = t & 31;
  const int x1 = warp * 8 + (lane & 7);           // column inside the 64-wide block
  const int ly = lane >> 3;                       // 0..3
  const double dx1 = (double)x1;
  const double c0s = __dmul_rn(s_m[0], dx1) * 32.0;
  const double c3s = __dmul_rn(s_m[3], dx1) * 32.0;
  const double c6 = __dmul_rn(s_m[6], dx1);
  const int pitch = td.W + 2;
  const uint2* __restrict__ tile = reinterpret_cast<const uint2*>(td.tile8);
  uint64_t keep_policy;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep_policy));

  const int n_xb = (x1 < 32) ? 4 : 3;             // 224 = 3*64 + 32 (warp-uniform)
  for (int xbi = 0; xbi < n_xb; ++xbi) {
    const int x = xbi * 64 + x1;
#pragma unroll 2
    for (int rg = 0; rg < BAND / 4; ++rg) {
      const int r = rg * 4 + ly;
      const double X0s = s_tab[xbi][r][0], Y0s = s_tab[xbi][r][1], W0 = s_tab[xbi][r][2];
      const double W = __dadd_rn(W0, c6);
      double rr = rcp_seed(W);
      rr = fma(rr, fma(-W, rr, 1.0), rr);
      rr = fma(rr, fma(-W, rr, 1.0), rr);
      const double fX = __dmul_rn(__dadd_rn(X0s, c0s), rr);
      const double fY = __dmul_rn(__dadd_rn(Y0s, c3s), rr);
      int X, Y;
      const bool okx = fast_round(fX, &X);
      const bool oky = fast_round(fY, &Y);
      if (!(okx && oky)) {   // rare: re-evaluate exactly as OpenCV does (matrix re-read from global)
        const int2 e = exact_coords(minv + (size_t)p * 9, xbi * 64, x1, band * BAND + r);
        X = e.x;
        Y = e.y;
      }
      const int sx = X >> INTER_BITS, sy = Y >> INTER_BITS;
      uint2 r0 = make_uint2(0u, 0u), r1 = make_uint2(0u, 0u);
      if ((unsigned)(sx + 1) <= (unsigned)td.W && (unsigned)(sy + 1) <= (unsigned)td.H) {
        const uint2* q = tile + (size_t)(sy + 1) * pitch + (sx + 1);
        r0 = ldg_tile(q, keep_policy);
        r1 = ldg_tile(q + 1, keep_policy);
      }
      s_px[r * SPITCH + swz_px(x)] = blend<ATT>(r0, r1, (uint32_t)(X & (INTER_TAB - 1)),
                                              (uint32_t)(Y & (INTER_TAB - 1)));
    }
  }
  __syncthreads();

  const size_t band_px = (size_t)p * VIEW * VIEW + (size_t)band * BAND * VIEW;
  if (views || (ATT && att)) {
    // 16 pixels per work item: 4 x 16 B of BGRA words -> 3 x 16 B of BGR bytes (+ 16 B of attention)
    uint4* vdst = reinterpret_cast<uint4*>(views + band_px * 3);
    uint4* adst = reinterpret_cast<uint4*>(att + band_px);
    const uint4* src = reinterpret_cast<const uint4*>(s_px);
    for (int g = t; g < BAND * VIEW / 16; g += THREADS) {
      const int row = g / (VIEW / 16), c16 = g - row * (VIEW / 16);
      uint4 q[4];
      const int sw = (c16 >> 1) & 3;                       // swz_px: chunk k of group c16 sits at k ^ sw
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = src[row * (SPITCH / 4) + c16 * 4 + (k ^ sw)];
      if (views) {
        uint32_t w[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          w[3 * k + 0] = __byte_perm(q[k].x, q[k].y, 0x4210);
          w[3 * k + 1] = __byte_perm(q[k].y, q[k].z, 0x5421);
          w[3 * k + 2] = __byte_perm(q[k].z, q[k].w, 0x6542);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)      // streaming stores: the views are not read back by this kernel
          __stcs(vdst + g * 3 + k, make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]));
      }
      if (ATT && att) {
        uint32_t a4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          a4[k] = __byte_perm(__byte_perm(q[k].x, q[k].y, 0x0073), __byte_perm(q[k].z, q[k].w, 0x0073), 0x5410);
        __stcs(adst + g, make_uint4(a4[0], a4[1], a4[2], a4[3]));
      }
    }
  }
  if (norm_nhwc) {
    // [P,224,224,4] bf16: R,G,B,0 -> one 8-byte store per pixel, coalesced
    uint2* dst = reinterpret_cast<uint2*>(norm_nhwc) + band_px;
    for (int i = t; i < BAND * VIEW; i += THREADS) {
      const int row = i / VIEW;
      const uint32_t v = s_px[row * SPITCH + swz_px(i - row * VIEW)];
      const float r_ = s_lut[(v >> 16) & 255u];
      const float g_ = s_lut[256 + ((v >> 8) & 255u)];
      const float b_ = s_lut[512 + (v & 255u)];
      const __nv_bfloat162 rg = __floats2bfloat162_rn(r_, g_);
      const __nv_bfloat162 b0 = __floats2bfloat162_rn(b_, 0.f);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&rg);
      o.y = *reinterpret_cast<const uint32_t*>(&b0);
      dst[i] = o;
    }
  }
  if (norm_nchw) {
    // [P,3,224,224] f32, channel c = RGB -> BGR byte 2-c; 4 pixels per store
    const uint4* src = reinterpret_cast<const uint4*>(s_px);
    for (int i = t; i < 3 * BAND * VIEW / 4; i += THREADS) {
      const int c = i / (BAND * VIEW / 4);
      const int j4 = i - c * (BAND * VIEW / 4);          // group of 4 pixels inside the band
      const int rowq = j4 / (VIEW / 4), chq = j4 - rowq * (VIEW / 4);
      const uint4 q = src[rowq * (SPITCH / 4) + (chq ^ ((chq >> 3) & 3))];
      const int sh = 8 * (2 - c);
      float4 o;
      o.x = s_lut[c * 256 + ((q.x >> sh) & 255u)];
      o.y = s_lut[c * 256 + ((q.y >> sh) & 255u)];
      o.z = s_lut[c * 256 + ((q.z >> sh) & 255u)];
      o.w = s_lut[c * 256 + ((q.w >> sh) & 255u)];
      float* base = norm_nchw + ((size_t)p * 3 + c) * VIEW * VIEW + (size_t)band * BAND * VIEW + j4 * 4;
      *reinterpret_cast<float4*>(base) = o;
    }
  }
}

}  // namespace

// ===================================================================== C ABI
extern "C" int avdn_pack_tile(const uint8_t* map_bgr, const uint8_t* att, int att_ch, int H,
                              int W, void* tile8, avdn_stream_t stream) {
  AVDN_REQUIRE(map_bgr && tile8, "avdn_pack_tile: null pointer");
  AVDN_REQUIRE(((uintptr_t)tile8 & 7) == 0, "avdn_pack_tile: tile8 must be 8-byte aligned");
  AVDN_REQUIRE(H > 0 && W > 0 && H <= 32766 && W <= 32766, "avdn_pack_tile: bad size %dx%d", H, W);
  AVDN_REQUIRE(!att || att_ch >= 1, "avdn_pack_tile: att_ch must be >= 1");
  const long long n = (long long)(H + 1) * (W + 2);
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  pack_tile_kernel<<<blocks, 256, 0, avdn::to_cuda(stream)>>>(map_bgr, att, att_ch, H, W,
                                                              reinterpret_cast<uint2*>(tile8));
  return avdn::check_launch("avdn_pack_tile");
}

extern "C" int avdn_resize_area_width(const uint8_t* src, int H, int W, int new_w, const int32_t* ofs,
                                      const int32_t* sidx, const float* alpha, uint8_t* dst, avdn_stream_t stream) {
  AVDN_REQUIRE(src && dst && ofs && sidx && alpha && H > 0 && W > 0 && new_w > 0 && new_w <= W,
               "avdn_resize_area_width: bad argument (horizontal shrink only)");
  const long long n = (long long)H * new_w;
  const int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  resize_area_width_kernel<<<blocks, 256, 0, avdn::to_cuda(stream)>>>(src, H, W, new_w, ofs, sidx, alpha, dst);
  return avdn::check_launch("avdn_resize_area_width");
}

extern "C" int avdn_raster_attention(const int32_t* spots, int n_spots, int rmax, int32_t* hw_scratch, int H, int W,
                                     int ch, uint8_t* att, avdn_stream_t stream) {
  AVDN_REQUIRE(att && H > 0 && W > 0 && ch >= 1 && n_spots >= 0 && rmax >= 0, "avdn_raster_attention: bad argument");
  AVDN_REQUIRE(n_spots == 0 || (spots && hw_scratch), "avdn_raster_attention: null spots / scratch");
  cudaStream_t s = avdn::to_cuda(stream);
  if (n_spots > 0) {
    circle_spans_kernel<<<(n_spots + 63) / 64, 64, 0, s>>>(spots, n_spots, rmax, hw_scratch);
    int r = avdn::check_launch("avdn_raster_attention (spans)");
    if (r) return r;
  }
  const long long np = (long long)H * W;
  const int blocks = (int)((np + 255) / 256 < 148 * 16 ? (np + 255) / 256 : 148 * 16);
  raster_attention_kernel<<<blocks, 256, 0, s>>>(spots, n_spots, rmax, hw_scratch, H, W, ch, att);
  return avdn::check_launch("avdn_raster_attention");
}

extern "C" int avdn_gps_to_pixels(const double* corners_gps, const double* geo, int P,
                                  int32_t* corners_px, avdn_stream_t stream) {
  AVDN_REQUIRE(P >= 0, "avdn_gps_to_pixels: P < 0");
  if (P == 0) return AVDN_OK;
  AVDN_REQUIRE(corners_gps && geo && corners_px, "avdn_gps_to_pixels: null pointer");
  gps_to_pixels_kernel<<<(P * 4 + 127) / 128, 128, 0, avdn::to_cuda(stream)>>>(corners_gps, geo, P,
                                                                                corners_px);
  return avdn::check_launch("avdn_gps_to_pixels");
}

extern "C" int avdn_homography_from_corners(const int32_t* corners_px, int P, double* minv,
                                            avdn_stream_t stream) {
  AVDN_REQUIRE(P >= 0, "avdn_homography_from_corners: P < 0");
  if (P == 0) return AVDN_OK;
  AVDN_REQUIRE(corners_px && minv, "avdn_homography_from_corners: null pointer");
  homography_kernel<<<(P + 63) / 64, 64, 0, avdn::to_cuda(stream)>>>(corners_px, P, minv);
  return avdn::check_launch("avdn_homography_from_corners");
}

extern "C" int avdn_render_views(const avdn_tile_desc* tiles, int n_tiles, const int32_t* tile_idx,
                                 const double* minv, int P, uint8_t* views, uint8_t* att,
                                 float* norm_nchw, void* norm_nhwc, const float* norm_lut,
                                 avdn_stream_t stream) {
  AVDN_REQUIRE(P >= 0, "avdn_render_views: P < 0");
  if (P == 0) return AVDN_OK;
  AVDN_REQUIRE(tiles && n_tiles >= 1 && minv, "avdn_render_views: null tiles/minv");
  AVDN_REQUIRE(views || att || norm_nchw || norm_nhwc, "avdn_render_views: no output requested");
  AVDN_REQUIRE(!(norm_nchw || norm_nhwc) || norm_lut, "avdn_render_views: norm output needs norm_lut");
  AVDN_REQUIRE(((uintptr_t)views & 15) == 0 && ((uintptr_t)att & 15) == 0 &&
                   ((uintptr_t)norm_nchw & 15) == 0 && ((uintptr_t)norm_nhwc & 15) == 0,
               "avdn_render_views: outputs must be 16-byte aligned");
  AVDN_REQUIRE((long long)P * NBAND < 2147483647LL, "avdn_render_views: too many poses");
  if (att)
    render_kernel<true><<<P * NBAND, THREADS, 0, avdn::to_cuda(stream)>>>(
        tiles, tile_idx, minv, P, views, att, norm_nchw,
        reinterpret_cast<__nv_bfloat16*>(norm_nhwc), norm_lut);
  else
    render_kernel<false><<<P * NBAND, THREADS, 0, avdn::to_cuda(stream)>>>(
        tiles, tile_idx, minv, P, views, att, norm_nchw,
        reinterpret_cast<__nv_bfloat16*>(norm_nhwc), norm_lut);
  return avdn::check_launch("avdn_render_views");
}
