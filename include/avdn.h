/*
 * avdn.h — C ABI of libavdn.so, the sm_100a device library behind the AVDN
 * (Aerial Vision-and-Dialog Navigation) HAA-Transformer hot path.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name ends in `_host`;
 *   - the caller owns every buffer (inputs, outputs, workspace); nothing is
 *     allocated, retained or freed by the library;
 *   - work is enqueued on `stream`; the library never synchronises the device;
 *   - return 0 on success, a negative avdn_status on failure;
 *     avdn_last_error_string() gives the reason (thread-local);
 *   - there is no CPU fallback: on a machine without an sm_100 device every
 *     compute entry point returns AVDN_ERR_NO_DEVICE / AVDN_ERR_LAUNCH.
 *
 * Each declaration cites the reference code (file:line under /root/reference)
 * that it replaces.
 */
#ifndef AVDN_H_
#define AVDN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* avdn_stream_t; /* == cudaStream_t */

typedef enum avdn_status {
  AVDN_OK = 0,
  AVDN_ERR_BAD_ARG = -1,    /* null pointer, bad shape, bad alignment */
  AVDN_ERR_UNSUPPORTED = -2,/* shape / mode outside what the kernels implement */
  AVDN_ERR_LAUNCH = -3,     /* cudaGetLastError() after a launch */
  AVDN_ERR_NO_DEVICE = -4,  /* no CUDA device / not sm_100 */
  AVDN_ERR_DRIVER = -5      /* driver entry point (tensor-map encode) failed */
} avdn_status;

const char* avdn_last_error_string(void);
/* ABI version; bumped whenever a signature changes. */
int avdn_abi_version(void);
/* 1 if device 0 is compute capability 10.x, 0 otherwise (no exception). */
int avdn_device_supported(void);

#define AVDN_VIEW 224 /* src/env.py:273-274 */

/* ------------------------------------------------------------------------
 * Stage 1 — view renderer (src/env.py:254-332)
 * ---------------------------------------------------------------------- */

/* Map preparation: packs a BGR u8 HWC satellite tile and (optionally) its
 * human-attention map into the renderer's HBM layout: one u32 per pixel
 * (B | G<<8 | R<<16 | ATT<<24) with a 1-pixel zero border, row pitch W+2.
 * The zero border implements BORDER_CONSTANT(0) of cv2.warpPerspective
 * (src/env.py:290,292) without per-tap predicates.
 *   map_bgr  [H,W,3] u8      (self.map_batch[name],           src/env.py:217-221)
 *   att      [H,W,att_ch] u8 or NULL; channel 0 is taken
 *                            (self.attention_map_batch[name], src/env.py:224-231)
 *   tile4    [(H+2)*(W+2)] u32 (out)                                            */
int avdn_pack_tile(const uint8_t* map_bgr, const uint8_t* att, int att_ch,
                   int H, int W, uint32_t* tile4, avdn_stream_t stream);

/* gps_to_img_coords (src/env.py:189-196), batched and bit-exact:
 *   x = rint((lng - bl_lng) / lat_ratio), y = rint((tr_lat - lat) / lat_ratio)
 * in float64 with round-half-even.
 *   corners_gps [P,4,2] f64 (lat,lng);  geo [P,5] f64 = (bl_lat, bl_lng,
 *   tr_lat, tr_lng, lat_ratio) per pose;  corners_px [P,4,2] i32 (x,y) (out)   */
int avdn_gps_to_pixels(const double* corners_gps, const double* geo, int P,
                       int32_t* corners_px, avdn_stream_t stream);

/* cv2.getPerspectiveTransform (src/env.py:287) followed by the matrix
 * inversion cv2.warpPerspective performs internally (src/env.py:290):
 *   corners_px [P,4,2] i32 (FL,FR,BR,BL; x,y)  ->  minv [P,9] f64 row-major.
 * float64, same elimination / rounding order as OpenCV (no FMA contraction);
 * a singular system yields the all-zero matrix, as OpenCV does.                */
int avdn_homography_from_corners(const int32_t* corners_px, int P, double* minv,
                                 avdn_stream_t stream);

/* One packed tile as the renderer sees it. */
typedef struct avdn_tile_desc {
  const uint32_t* tile4; /* avdn_pack_tile output */
  int32_t H, W;          /* un-padded size */
} avdn_tile_desc;

/* cv2.warpPerspective(map, M, (224,224)) and the same warp of the attention
 * map (src/env.py:290-293), fixed-point bilinear, bit-exact, for P poses.
 *   tiles     [n_tiles] avdn_tile_desc (device array)
 *   tile_idx  [P] i32 or NULL (all poses use tiles[0])
 *   minv      [P,9] f64 from avdn_homography_from_corners
 * Outputs, each may be NULL:
 *   views     [P,224,224,3] u8 BGR HWC     = obs['current_view']  (env.py:308)
 *   att       [P,224,224]   u8             = 255 * obs['gt_saliency'] (env.py:293)
 *   norm_nchw [P,3,224,224] f32 RGB, (x-mean)/std (src/xview_et/agent.py:586-592)
 *   norm_nhwc [P,224,224,4] bf16 RGB0, same normalisation rounded to bf16
 *             (channel 3 is zero; this is the trunk's input layout)
 *   norm_lut  [3,256] f32, lut[c][v] = (float(v) - mean[c]) / std[c] for RGB
 *             channel c, computed on the host in float32 exactly as the
 *             reference does; required when norm_nchw or norm_nhwc is given.  */
int avdn_render_views(const avdn_tile_desc* tiles, int n_tiles,
                      const int32_t* tile_idx, const double* minv, int P,
                      uint8_t* views, uint8_t* att, float* norm_nchw,
                      void* norm_nhwc, const float* norm_lut,
                      avdn_stream_t stream);

/* ------------------------------------------------------------------------
 * Tensor-core primitive (tcgen05 / TMEM / TMA): every dense contraction of
 * the path runs through one planned kernel.  It replaces the cuDNN / cuBLAS
 * library calls behind
 *   nn.Conv2d fwd/dgrad/wgrad        src/models/dark_net.py:22-28
 *   nn.Linear / MHA projections      src/models/enc_vl.py:16-22,68
 *   QK^T, PV and their backward      (inside nn.TransformerEncoderLayer)
 *   ET heads / fc2                   src/models/ET_haa.py:98-119,144-167
 * A plan is a caller-owned HOST blob (avdn_gemm_plan_bytes() bytes) holding the
 * encoded TMA descriptors and launch geometry for fixed device pointers; it can
 * be re-run any number of times (and captured in a CUDA graph).
 * ---------------------------------------------------------------------- */
enum { AVDN_GEMM_PLAIN = 0, AVDN_GEMM_CONV = 1, AVDN_GEMM_WGRAD = 2 };
enum { AVDN_DT_BF16 = 0, AVDN_DT_F32 = 1 };

/* One k-step source of a convolution: which A (CONV) / B (WGRAD) view, the
 * offset added to the box's W and H coordinates, and the operand-B k offset
 * (CONV) or the output column offset (WGRAD) of this filter tap. */
typedef struct avdn_tap { int32_t map, d1, d2, bk; } avdn_tap;

/* A bf16 operand as a rank-4 strided tensor; dim[0] is contiguous.  Strides in
 * elements.  box = TMA box (box[0] must be 64). */
typedef struct avdn_operand {
  const void* ptr;
  int64_t dim[4];
  int64_t stride[4];
  int32_t box[4];
} avdn_operand;

typedef struct avdn_gemm_core {
  int32_t mode;               /* AVDN_GEMM_* */
  int32_t M, N;               /* logical output extent per batch (PLAIN/WGRAD); N only (CONV) */
  int32_t num_kb;             /* k-steps of 64: ceil(K/64) | taps*cblocks | pixel tiles */
  int32_t split_k;            /* >1 only with accumulate == 2 */
  int32_t batch0, batch1;     /* PLAIN: grid.z = batch0*batch1*split_k -> coords 2,3 */
  int32_t b_batched;          /* PLAIN: 1 if operand B has the batch dims too */
  int32_t cblocks, n_taps;    /* CONV: channel blocks per tap */
  avdn_tap taps[9];
  int32_t tiles_w, tiles_h, tiles_n;   /* spatial tiling of the box over (W,H,N) */
  int32_t box_w, box_h, box_n;
  int32_t valid_w, valid_h, valid_n;   /* CONV: extent of valid output coords */
  int32_t out_H, out_W, out_sh, out_sw, out_oh, out_ow; /* CONV: out pixel (n, h*sh+oh, w*sw+ow) */
  int32_t out_dtype;          /* AVDN_DT_* */
  int32_t accumulate;         /* 0 store, 1 out += (read-modify-write), 2 atomicAdd (fp32) */
  int32_t relu;
  float alpha;
  uint32_t tx_bytes;          /* filled by avdn_gemm_plan */
  int32_t pad_;
  int64_t ldc, out_bs0, out_bs1;       /* output row pitch / batch strides, elements */
  void* out;
  const float* bias;          /* [N] fp32 or NULL */
} avdn_gemm_core;

typedef struct avdn_gemm_desc {
  avdn_gemm_core core;
  int32_t bn;                 /* N tile: 64, 128 or 256 */
  int32_t a_mn, b_mn;         /* 0 = K-major operand, 1 = MN-major operand */
  int32_t n_a, n_b;           /* number of A / B views (parity views of stride-2 convs) */
  int32_t grid_m, grid_n, grid_z;
  avdn_operand a[4];
  avdn_operand b[4];
} avdn_gemm_desc;

size_t avdn_gemm_plan_bytes(void);
int avdn_gemm_plan(const avdn_gemm_desc* desc, void* plan_host, size_t plan_bytes);
int avdn_gemm_run(const void* plan_host, avdn_stream_t stream);

/* ------------------------------------------------------------------------
 * Stage 2 — Darknet trunk (src/models/dark_net.py).  Activations are NHWC
 * bf16, channels padded to a multiple of 64, addressed as R rows x C.
 * ---------------------------------------------------------------------- */

/* First convolution 3->32, 3x3, pad 1 (module_list.0.conv_0; K = 27 is too thin
 * for a tensor-core tile).  x [N,H,W,4] bf16 (R,G,B,0: avdn_render_views'
 * norm_nhwc), w [32,3,3,3] fp32 (nn.Conv2d layout), z [N,H,W,64] bf16 (ch 32..63 = 0). */
int avdn_conv0_fwd(const void* x_nhwc4, const float* w, void* z, int N, int H, int W, avdn_stream_t stream);
/* dw [32,3,3,3] fp32 += sum_pixels dz * x (weight gradient of the same layer). */
int avdn_conv0_wgrad(const void* dz, const void* x_nhwc4, float* dw, int N, int H, int W, avdn_stream_t stream);

/* nn.BatchNorm2d in train mode (dark_net.py:31; eps 1e-5, momentum 0.1): batch
 * statistics of z [R,C] bf16 -> per-channel affine scale = gamma*rstd,
 * shift = beta - mean*scale, plus mean/rstd for the backward pass; updates the
 * running buffers (NULL to skip).  sums [2,C] f64 is scratch.  Channels >= C_real
 * are padding (scale = shift = 0).                                            */
int avdn_bn_stats(const void* z, long long R, int C, int C_real, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, double* sums,
                  float* scale, float* shift, float* mean, float* rstd, avdn_stream_t stream);
/* eval mode: the same affine from the running statistics. */
int avdn_bn_eval_coeffs(int C, int C_real, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, avdn_stream_t stream);
/* a = LeakyReLU_slope(z*scale + shift) (+ residual): BN apply + nn.LeakyReLU()
 * (dark_net.py:33, slope 0.01) + the shortcut add (dark_net.py:224-226).       */
int avdn_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, void* a,
                  long long R, int C, float slope, avdn_stream_t stream);
/* Backward of the above (train mode): dz [R,C] bf16 from da; dgamma/dbeta (+=, the
 * C_real real channels).  The residual branch receives da unchanged (caller).   */
int avdn_bn_backward(const void* da, const void* z, const float* scale, const float* shift, const float* mean,
                     const float* rstd, long long R, int C, int C_real, float slope, double* sums, void* dz,
                     float* dgamma, float* dbeta, avdn_stream_t stream);

/* nn.Conv2d weight [Cout,Cin,k,k] fp32 -> GEMM operands (bf16, zero padded):
 * wf [Cout_p][k*k][Cin_p] (forward / wgrad layout), wd [Cin_p][k*k][Cout_p] (dgrad). */
int avdn_pack_conv_weight(const float* w, int Cout, int Cin, int k, int Cout_p, int Cin_p, void* wf, void* wd,
                          avdn_stream_t stream);
/* grad [Cout,Cin,k,k] fp32 += dwf [Cout_p][k*k][Cin_p] fp32 (WGRAD output layout). */
int avdn_unpack_conv_wgrad(const float* dwf, int Cout, int Cin, int k, int Cin_p, float* grad,
                           avdn_stream_t stream);
int avdn_cast_f32_bf16(const float* in, void* out, long long n, avdn_stream_t stream);
/* Trunk output [N,HW,C] bf16 NHWC -> `frames` [N,C,HW] fp32 (the .view at
 * src/xview_et/agent.py:594) and its adjoint for the backward pass.           */
int avdn_nhwc_to_nchw_f32(const void* in, float* out, int N, int HW, int C, avdn_stream_t stream);
int avdn_nchw_f32_to_nhwc(const float* in, void* out, int N, int HW, int C, avdn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVDN_H_ */
