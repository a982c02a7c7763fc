// Agent slice of the AVDN hot path (src/xview_et/agent.py): the per-step loss
// (four MSE-sum terms, the angular term and the NSS human-attention loss with
// the 8x8 -> 224x224 bilinear upsample fused in, forward AND backward in one
// kernel), waypoint post-processing / discretisation, and the optimiser step
// (global-norm clip + AdamW over a flat parameter arena).
#include "common.cuh"

namespace {

constexpr int VIEW = AVDN_VIEW;
constexpr int NPX = VIEW * VIEW;
constexpr float PI_REF = 3.14159f;      // the reference's pi (agent.py:606,666,745)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// F.interpolate(bilinear, align_corners=False) source taps of destination index d (8 -> 224)
__device__ __forceinline__ void bilin_tap(int d, int* i0, int* i1, float* l1) {
  const float scale = 8.0f / 224.0f;
  float src = scale * ((float)d + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  const int a = (int)src;
  *i0 = a;
  *i1 = a + (a < 7 ? 1 : 0);
  *l1 = src - (float)a;
}

__device__ __forceinline__ float ang_ref(float v0, float v1) {
  // ((atan2(v0, v1) / 3.14159 + 2) / 2) % 1   in float32, as torch evaluates it
  const float a = (atan2f(v0, v1) / PI_REF + 2.0f) / 2.0f;
  return a - floorf(a);
}

// One CTA per sample: loss_i and its gradient w.r.t. output[i,:] and h_sali[i,:].
//   agent.py:663-669  4 x MSELoss(reduction='sum') + angular term
//   agent.py:673-681  nss_w * NSS(pred_saliency[i], gt_saliency[i]) when sum(gt) > 0
//   agent.py:883-885  loss += ml_loss * train_ml / batch_size   (== `scale`)
//
// NSS without ever forming the 224x224 prediction: the upsample is linear and separable,
//   p[y,x] = sum_ab Wy[y,a] Wx[x,b] h[a,b]        (Wy = Wx = W, the 224x8 bilinear tap matrix)
// so with c[a] = sum_y W[y,a], G[a,a'] = sum_y W[y,a] W[y,a'] (constants) and ONE pass over the
// fixation map f for F[a,b] = sum_yx W[y,a] W[x,b] f[y,x] and SF = sum f:
//   sum p   = sum_ab h[a,b] c[a] c[b]          sum p^2 = sum_ab h[a,b] (G h G^T)[a,b]
//   sum p f = sum_ab h[a,b] F[a,b]
//   dL/dh[a,b] = k1 (F[a,b] - fm c[a] c[b]) - k2 ((G h G^T)[a,b] - m c[a] c[b])
// (the adjoint of the upsample applied to the per-pixel gradient k1 (f - fm) - k2 (p - m)).
__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ output, const float* __restrict__ h_sali,
                                                   const float* __restrict__ gt_xy, const float* __restrict__ gt_alt,
                                                   const float* __restrict__ gt_prog, const uint8_t* __restrict__ att,
                                                   const float* __restrict__ jitter, float nss_w, int nss_r,
                                                   double scale, double* __restrict__ loss_total,
                                                   double* __restrict__ loss_i, float* __restrict__ d_output,
                                                   float* __restrict__ d_h_sali) {
  __shared__ float s_h[64];
  __shared__ float s_dh[64];
  __shared__ double s_c[8], s_G[8][8], s_F[8][8], s_GhG[8][8], s_T[8][8];
  __shared__ double s_col[8][VIEW];          // per-column partial sums A[a][x] = sum_y W[y,a] f[y,x]
  __shared__ double s_red[8];
  __shared__ double s_stat[6];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 64) {
    s_h[tid] = h_sali[b * 64 + tid]; s_dh[tid] = 0.f;
    (&s_G[0][0])[tid] = 0.0; (&s_F[0][0])[tid] = 0.0;
  }
  if (tid < 8) s_c[tid] = 0.0;
  __syncthreads();
  double nss_term = 0.0;
  bool use_nss = false;
  if (att != nullptr && nss_w != 0.f) {
    const uint8_t* f = att + (size_t)b * NPX;
    int t0 = 0, t1 = 0; float tl = 0.f;
    double sf = 0.0;
    if (tid < VIEW) {
      bilin_tap(tid, &t0, &t1, &tl);           // this thread's taps, as a row (for c, G) and as a column
      const double w0 = (double)(1.f - tl), w1 = (double)tl;
      atomicAdd(&s_c[t0], w0); atomicAdd(&s_c[t1], w1);
      atomicAdd(&s_G[t0][t0], w0 * w0); atomicAdd(&s_G[t0][t1], w0 * w1);
      atomicAdd(&s_G[t1][t0], w1 * w0); atomicAdd(&s_G[t1][t1], w1 * w1);
#pragma unroll
      for (int a = 0; a < 8; ++a) s_col[a][tid] = 0.0;
      // column pass: coalesced byte loads (thread = x), weights of row y into its two tap cells
      for (int y = 0; y < VIEW; ++y) {
        int y0, y1; float ly;
        bilin_tap(y, &y0, &y1, &ly);           // warp-uniform
        const double fv = (double)f[y * VIEW + tid] / 255.0;
        sf += fv;
        s_col[y0][tid] += (double)(1.f - ly) * fv;
        s_col[y1][tid] += (double)ly * fv;
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const double v = s_col[a][tid];
        atomicAdd(&s_F[a][t0], v * w0);
        atomicAdd(&s_F[a][t1], v * w1);
      }
    }
    sf = warp_sum_d(sf);
    if (lane == 0) s_red[warp] = sf;
    __syncthreads();
    if (tid < 64) {                            // T = G h  (rows), then GhG = T G^T (columns)
      const int a = tid >> 3, c2 = tid & 7;
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += s_G[a][k] * (double)s_h[k * 8 + c2];
      s_T[a][c2] = t;
    }
    __syncthreads();
    if (tid < 64) {
      const int a = tid >> 3, c2 = tid & 7;
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += s_T[a][k] * s_G[c2][k];
      s_GhG[a][c2] = t;
    }
    __syncthreads();
    if (tid == 0) {
      double SF = 0.0;
      for (int w = 0; w < 8; ++w) SF += s_red[w];
      double sp = 0.0, spp = 0.0, spf = 0.0;
      for (int a = 0; a < 8; ++a)
        for (int c2 = 0; c2 < 8; ++c2) {
          const double h = (double)s_h[a * 8 + c2];
          sp += h * s_c[a] * s_c[c2];
          spp += h * s_GhG[a][c2];
          spf += h * s_F[a][c2];
        }
      const double m = sp / NPX;
      double var = (spp - NPX * m * m) / (NPX - 1);
      if (var < 0) var = 0;
      s_stat[0] = m; s_stat[1] = sqrt(var); s_stat[2] = SF; s_stat[3] = spf;
    }
    __syncthreads();
    const double m = s_stat[0], sd = s_stat[1], SF = s_stat[2], SPF = s_stat[3];
    if (SF > 0.0 && sd > 0.0) {
      use_nss = true;
      const double half = (nss_r == 0) ? 1.0 : 0.5;
      const double A = SPF - m * SF;
      const double Fe = SF + 0.001;
      double sum_nf = half * A / sd;
      if (nss_r == 1) sum_nf += SF;
      else if (nss_r == -1) sum_nf -= SF;
      nss_term = -(double)nss_w * sum_nf / Fe;
      // per pixel: d/dp = c * [ (f - SF/N)/sd - A (p - m) / ((N-1) sd^3) ]; its upsample adjoint:
      const double c = -(double)nss_w * half / Fe * scale;
      const double k1 = c / sd, k2 = c * A / ((NPX - 1.0) * sd * sd * sd), fm = SF / NPX;
      if (tid < 64) {
        const int a = tid >> 3, c2 = tid & 7;
        const double cc = s_c[a] * s_c[c2];
        s_dh[tid] = (float)(k1 * (s_F[a][c2] - fm * cc) - k2 * (s_GhG[a][c2] - m * cc));
      }
    }
  }
  __syncthreads();
  if (tid < 64) d_h_sali[b * 64 + tid] = s_dh[tid];
  if (tid == 0) {
    const float px = output[b * 4], py = output[b * 4 + 1], pa = output[b * 4 + 2], pp = output[b * 4 + 3];
    const float gx = gt_xy[b * 2], gy = gt_xy[b * 2 + 1];
    const float j = jitter ? jitter[b] : 0.f;
    const float pyj = py + j;
    const float dax = px - gx, day = py - gy, dalt = pa - gt_alt[b], dpr = pp - gt_prog[b];
    const float dang = ang_ref(px, pyj) - ang_ref(gx, gy);
    const float mse = dax * dax + day * day;
    double l = (double)mse + (double)(dang * dang) + (double)(dalt * dalt) + (double)(dpr * dpr);
    if (use_nss) l += nss_term;
    loss_i[b] = l;
    atomicAdd(loss_total, l * scale);
    const float r2 = px * px + pyj * pyj;
    const float k = r2 > 0.f ? (2.f * dang) / (2.f * PI_REF * r2) : 0.f;
    const float sc = (float)scale;
    d_output[b * 4] = sc * (2.f * dax + k * pyj);
    d_output[b * 4 + 1] = sc * (2.f * day - k * px);
    d_output[b * 4 + 2] = sc * 2.f * dalt;
    d_output[b * 4 + 3] = sc * 2.f * dpr;
  }
}

// 8x8 -> 224x224 bilinear upsample (pred_saliency of the reference API)
__global__ void upsample_kernel(const float* __restrict__ h_sali, int B, float* __restrict__ pred) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)B * NPX) return;
  const int b = (int)(i / NPX), r = (int)(i % NPX), y = r / VIEW, x = r - y * VIEW;
  int y0, y1, x0, x1; float ly, lx;
  bilin_tap(y, &y0, &y1, &ly);
  bilin_tap(x, &x0, &x1, &lx);
  const float* h = h_sali + b * 64;
  pred[i] = (1.f - ly) * ((1.f - lx) * h[y0 * 8 + x0] + lx * h[y0 * 8 + x1]) +
            ly * ((1.f - lx) * h[y1 * 8 + x0] + lx * h[y1 * 8 + x1]);
}

// adjoint of the upsample: d_h_sali[b][64] = sum_pixels w * d_pred.  One CTA per sample.
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ d_pred,
                                                           float* __restrict__ d_h_sali) {
  __shared__ float s_dh[64];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < 64) s_dh[tid] = 0.f;
  __syncthreads();
  const float* g = d_pred + (size_t)b * NPX;
  for (int i = tid; i < NPX; i += 256) {
    const int y = i / VIEW, x = i - y * VIEW;
    int y0, y1, x0, x1; float ly, lx;
    bilin_tap(y, &y0, &y1, &ly);
    bilin_tap(x, &x0, &x1, &lx);
    const float v = g[i];
    atomicAdd(&s_dh[y0 * 8 + x0], v * (1.f - ly) * (1.f - lx));
    atomicAdd(&s_dh[y0 * 8 + x1], v * (1.f - ly) * lx);
    atomicAdd(&s_dh[y1 * 8 + x0], v * ly * (1.f - lx));
    atomicAdd(&s_dh[y1 * 8 + x1], v * ly * lx);
  }
  __syncthreads();
  if (tid < 64) d_h_sali[b * 64 + tid] = s_dh[tid];
}

// agent.py:637-653,738,745-752: normalise, clamp, discretise.  Integers are bit-exact with the
// host code: float32 element arithmetic, then float64 for the literal-constant expressions.
__global__ void postprocess_kernel(const float* __restrict__ output, const double* __restrict__ edge_len, int B,
                                   float stop_threshold, int* __restrict__ angle_deg, double* __restrict__ dist,
                                   int* __restrict__ altitude_m, uint8_t* __restrict__ stop,
                                   float* __restrict__ xy_norm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float x = output[i * 4], y = output[i * 4 + 1];
  const float m = fmaxf(fmaxf(fabsf(x), fabsf(y)), 1.f);
  x = __fdiv_rn(x, m);
  y = __fdiv_rn(y, m);
  const float a = fminf(1.f, fmaxf(0.f, output[i * 4 + 2]));
  const float p = fminf(1.f, fmaxf(0.f, output[i * 4 + 3]));
  // np.arctan2 on float32 scalars -> float32; the python-float constants promote to float64
  const double at = (double)atan2f(x, y);
  double a_dir = (at / 3.14159 + 2.0) / 2.0;
  a_dir = a_dir - floor(a_dir);
  angle_deg[i] = __double2int_rn(a_dir * 360.0);
  // np.linalg.norm of a float32 pair -> float32
  const float nrm = sqrtf(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
  dist[i] = (double)nrm * (edge_len[i] / 2.0);
  altitude_m[i] = __double2int_rn((double)a * 360.0) + 40;
  stop[i] = p > stop_threshold ? 1 : 0;
  if (xy_norm) { xy_norm[i * 2] = x; xy_norm[i * 2 + 1] = y; }
}

// ----------------------------------------------------------------- optimiser
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  double a = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = g[i];
    a += (double)v * v;
  }
  a = warp_sum_d(a);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    atomicAdd(out, t);
  }
}

// torch.optim.AdamW step over a flat arena; grads are first scaled by
// min(1, max_norm / (sqrt(*sumsq) + 1e-6)) when sumsq != NULL (clip_grad_norm_).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                    float beta1, float beta2, float eps, float wd, float bias1,
                                                    float bias2_sqrt, const double* __restrict__ sumsq, float max_norm,
                                                    float grad_scale) {
  float coef = grad_scale;
  if (sumsq) {
    const float nrm = (float)sqrt(*sumsq) * grad_scale;
    const float c = max_norm / (nrm + 1e-6f);
    if (c < 1.f) coef *= c;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias2_sqrt + eps;
    pi -= (lr / bias1) * (mi / denom);
    p[i] = pi;
  }
}

}  // namespace

extern "C" int avdn_loss(const float* output, const float* h_sali, const float* gt_xy, const float* gt_alt,
                         const float* gt_prog, const uint8_t* att, const float* jitter, int B, float nss_w, int nss_r,
                         double scale, double* loss_total, double* loss_i, float* d_output, float* d_h_sali,
                         avdn_stream_t stream) {
  AVDN_REQUIRE(output && h_sali && gt_xy && gt_alt && gt_prog && loss_total && loss_i && d_output && d_h_sali,
               "avdn_loss: null pointer");
  AVDN_REQUIRE(nss_r >= -1 && nss_r <= 1, "avdn_loss: nss_r must be -1, 0 or 1");
  if (B == 0) return AVDN_OK;
  loss_kernel<<<B, 256, 0, avdn::to_cuda(stream)>>>(output, h_sali, gt_xy, gt_alt, gt_prog, att, jitter, nss_w, nss_r,
                                                  scale, loss_total, loss_i, d_output, d_h_sali);
  return avdn::check_launch("avdn_loss");
}

extern "C" int avdn_upsample_saliency(const float* h_sali, int B, float* pred, avdn_stream_t stream) {
  AVDN_REQUIRE(h_sali && pred, "avdn_upsample_saliency: null pointer");
  if (B == 0) return AVDN_OK;
  const long long n = (long long)B * NPX;
  upsample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, avdn::to_cuda(stream)>>>(h_sali, B, pred);
  return avdn::check_launch("avdn_upsample_saliency");
}

extern "C" int avdn_upsample_saliency_bwd(const float* d_pred, int B, float* d_h_sali, avdn_stream_t stream) {
  AVDN_REQUIRE(d_pred && d_h_sali, "avdn_upsample_saliency_bwd: null pointer");
  if (B == 0) return AVDN_OK;
  upsample_bwd_kernel<<<B, 256, 0, avdn::to_cuda(stream)>>>(d_pred, d_h_sali);
  return avdn::check_launch("avdn_upsample_saliency_bwd");
}

extern "C" int avdn_postprocess_waypoints(const float* output, const double* edge_len, int B, float stop_threshold,
                                          int* angle_deg, double* dist, int* altitude_m, uint8_t* stop,
                                          float* xy_norm, avdn_stream_t stream) {
  AVDN_REQUIRE(output && edge_len && angle_deg && dist && altitude_m && stop, "avdn_postprocess_waypoints: null pointer");
  if (B == 0) return AVDN_OK;
  postprocess_kernel<<<(B + 127) / 128, 128, 0, avdn::to_cuda(stream)>>>(output, edge_len, B, stop_threshold,
                                                                        angle_deg, dist, altitude_m, stop, xy_norm);
  return avdn::check_launch("avdn_postprocess_waypoints");
}

extern "C" int avdn_sumsq(const float* g, long long n, double* out, avdn_stream_t stream) {
  AVDN_REQUIRE(g && out && n >= 0, "avdn_sumsq: bad argument");
  if (n == 0) return AVDN_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)avdn::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, avdn::to_cuda(stream)>>>(g, n, out);
  return avdn::check_launch("avdn_sumsq");
}

extern "C" int avdn_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                          float beta2, float eps, float wd, int step, const double* sumsq, float max_norm,
                          float grad_scale, avdn_stream_t stream) {
  AVDN_REQUIRE(p && g && m && v && n >= 0 && step >= 1, "avdn_adamw: bad argument");
  if (n == 0) return AVDN_OK;
  const float bias1 = 1.f - powf(beta1, (float)step);
  const float bias2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)avdn::sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<(unsigned)blocks, 256, 0, avdn::to_cuda(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bias1,
                                                                  bias2_sqrt, sumsq, max_norm, grad_scale);
  return avdn::check_launch("avdn_adamw");
}
