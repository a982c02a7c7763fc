// Language encoder feeding the AVDN hot path: the warp-level kernels of
// `CustomBERTModel` (src/models/vln_model.py:128-159 = HuggingFace BertModel + a 2-layer head) that are not
// already in encoder.cu.  The dense contractions run through the tcgen05 GEMM primitive (gemm.cu), LayerNorm
// (eps 1e-12) and the masked softmax (key-padding mode, T = 0) through encoder.cu's kernels.
//
//   bert_embed_ln   word[ids] + position[0..S) + token_type[0] -> LayerNorm  (BertEmbeddings)
//   gelu_fwd / bwd  erf-GELU of the FFN's intermediate activation (hidden_act = "gelu")
//   bert_embed_bwd  scatter-add of the embedding gradients
#include "common.cuh"

namespace {

constexpr int E = 768;
constexpr int EPL = E / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per token row
__global__ void __launch_bounds__(256) bert_embed_ln_kernel(const long long* __restrict__ ids,
                                                            const float* __restrict__ word, const float* __restrict__ pos,
                                                            const float* __restrict__ type0,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            long long M, int S, int vocab, float eps,
                                                            float* __restrict__ v_out, float* __restrict__ y,
                                                            __nv_bfloat16* __restrict__ y16, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  long long id = ids[row];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const int s = (int)(row % S);
  float x[EPL];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int j = lane + 32 * i;
    const float t = (word[id * E + j] + type0[j]) + pos[(size_t)s * E + j];   // HF: inputs + token_type, then + position
    x[i] = t;
    sum += t;
  }
  const float mu = warp_sum(sum) * (1.f / E);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) { const float d = x[i] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) * (1.f / E) + eps);
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int j = lane + 32 * i;
    const float o = (x[i] - mu) * rs * gamma[j] + beta[j];
    if (v_out) v_out[row * E + j] = x[i];
    if (y) y[row * E + j] = o;
    if (y16) y16[row * E + j] = __float2bfloat16_rn(o);
  }
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// d_word[ids[row]] += dv[row], d_pos[row % S] += dv[row], d_type0 += dv[row]
__global__ void __launch_bounds__(256) bert_embed_bwd_kernel(const long long* __restrict__ ids,
                                                             const float* __restrict__ dv, long long M, int S, int vocab,
                                                             float* __restrict__ d_word, float* __restrict__ d_pos,
                                                             float* __restrict__ d_type0) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  long long id = ids[row];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const int s = (int)(row % S);
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int j = lane + 32 * i;
    const float g = dv[row * E + j];
    atomicAdd(&d_word[id * E + j], g);
    atomicAdd(&d_pos[(size_t)s * E + j], g);
    atomicAdd(&d_type0[j], g);
  }
}

__device__ __forceinline__ float gelu_erf(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const __nv_bfloat162* __restrict__ u, __nv_bfloat162* __restrict__ h,
                                                       long long n2) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n2; i += (long long)gridDim.x * 256) {
    const __nv_bfloat162 v = u[i];
    h[i] = __floats2bfloat162_rn(gelu_erf(__low2float(v)), gelu_erf(__high2float(v)));
  }
}

// du = dh * gelu'(u),  gelu'(u) = Phi(u) + u * phi(u)
__device__ __forceinline__ float gelu_grad(float u) {
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat162* __restrict__ u,
                                                       const __nv_bfloat162* __restrict__ dh,
                                                       __nv_bfloat162* __restrict__ du, long long n2) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n2; i += (long long)gridDim.x * 256) {
    const __nv_bfloat162 v = u[i], g = dh[i];
    du[i] = __floats2bfloat162_rn(__low2float(g) * gelu_grad(__low2float(v)),
                                  __high2float(g) * gelu_grad(__high2float(v)));
  }
}

// Backward of y = act(x w^T + b) for the tiny fp32 heads (pooler, linears): g = dy * act'(y);
// dx[m][k] (+)= sum_n g[m][n] w[n][k];  dw[n][k] += sum_m g[m][n] x[m][k];  db[n] += sum_m g[m][n].
// act: 0 none, 1 ReLU (y > 0), 2 tanh (1 - y^2).  One thread per dx element, then one per dw element.
__device__ __forceinline__ float act_grad(float dy, float y, int act) {
  if (act == 1) return y > 0.f ? dy : 0.f;
  if (act == 2) return dy * (1.f - y * y);
  return dy;
}
__global__ void __launch_bounds__(256) linear_f32_bwd_kernel(const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ w, const float* __restrict__ y,
                                                             const float* __restrict__ dy, int M, int N, int K, int act,
                                                             float* __restrict__ dx, long long lddx, int dx_accumulate,
                                                             float* __restrict__ dw, float* __restrict__ db) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  const long long n_dx = (long long)M * K, n_dw = (long long)N * K;
  if (i < n_dx) {
    if (dx) {
      const int m = (int)(i / K), k = (int)(i % K);
      float a = 0.f;
      for (int n = 0; n < N; ++n) a = fmaf(act_grad(dy[(long long)m * N + n], y[(long long)m * N + n], act), w[(long long)n * K + k], a);
      float* o = dx + (long long)m * lddx + k;
      *o = dx_accumulate ? *o + a : a;
    }
  } else if (i < n_dx + n_dw) {
    const long long j = i - n_dx;
    const int n = (int)(j / K), k = (int)(j % K);
    float a = 0.f, bsum = 0.f;
    for (int m = 0; m < M; ++m) {
      const float g = act_grad(dy[(long long)m * N + n], y[(long long)m * N + n], act);
      a = fmaf(g, x[(long long)m * ldx + k], a);
      bsum += g;
    }
    dw[j] += a;
    if (k == 0 && db) db[n] += bsum;
  }
}

inline unsigned grid_for2(long long n2) {
  long long b = (n2 + 255) / 256;
  const long long cap = (long long)avdn::sm_count() * 16;
  return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int avdn_bert_embed_ln(const long long* ids, const float* word, const float* pos, const float* type0,
                                  const float* gamma, const float* beta, int B, int S, int vocab, float eps,
                                  float* v_out, float* y, void* y16, float* mean, float* rstd, avdn_stream_t stream) {
  AVDN_REQUIRE(ids && word && pos && type0 && gamma && beta && mean && rstd && B > 0 && S > 0 && vocab > 0,
               "avdn_bert_embed_ln: bad argument");
  const long long M = (long long)B * S;
  bert_embed_ln_kernel<<<(unsigned)((M + 7) / 8), 256, 0, avdn::to_cuda(stream)>>>(
      ids, word, pos, type0, gamma, beta, M, S, vocab, eps, v_out, y, reinterpret_cast<__nv_bfloat16*>(y16), mean, rstd);
  return avdn::check_launch("avdn_bert_embed_ln");
}

extern "C" int avdn_bert_embed_bwd(const long long* ids, const float* dv, int B, int S, int vocab, float* d_word,
                                   float* d_pos, float* d_type0, avdn_stream_t stream) {
  AVDN_REQUIRE(ids && dv && d_word && d_pos && d_type0 && B > 0 && S > 0, "avdn_bert_embed_bwd: bad argument");
  const long long M = (long long)B * S;
  bert_embed_bwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, avdn::to_cuda(stream)>>>(ids, dv, M, S, vocab, d_word, d_pos,
                                                                                  d_type0);
  return avdn::check_launch("avdn_bert_embed_bwd");
}

extern "C" int avdn_gelu_fwd(const void* u, void* h, long long n, avdn_stream_t stream) {
  AVDN_REQUIRE(u && h && n >= 0 && (n % 2) == 0, "avdn_gelu_fwd: bad argument (n must be even)");
  if (n == 0) return AVDN_OK;
  gelu_fwd_kernel<<<grid_for2(n / 2), 256, 0, avdn::to_cuda(stream)>>>(
      reinterpret_cast<const __nv_bfloat162*>(u), reinterpret_cast<__nv_bfloat162*>(h), n / 2);
  return avdn::check_launch("avdn_gelu_fwd");
}

extern "C" int avdn_gelu_bwd(const void* u, const void* dh, void* du, long long n, avdn_stream_t stream) {
  AVDN_REQUIRE(u && dh && du && n >= 0 && (n % 2) == 0, "avdn_gelu_bwd: bad argument (n must be even)");
  if (n == 0) return AVDN_OK;
  gelu_bwd_kernel<<<grid_for2(n / 2), 256, 0, avdn::to_cuda(stream)>>>(
      reinterpret_cast<const __nv_bfloat162*>(u), reinterpret_cast<const __nv_bfloat162*>(dh),
      reinterpret_cast<__nv_bfloat162*>(du), n / 2);
  return avdn::check_launch("avdn_gelu_bwd");
}

extern "C" int avdn_linear_f32_bwd(const float* x, long long ldx, const float* w, const float* y, const float* dy, int M,
                                   int N, int K, int act, float* dx, long long lddx, int dx_accumulate, float* dw,
                                   float* db, avdn_stream_t stream) {
  AVDN_REQUIRE(x && w && y && dy && dw && M > 0 && N > 0 && K > 0 && act >= 0 && act <= 2,
               "avdn_linear_f32_bwd: bad argument");
  const long long n = (long long)M * K + (long long)N * K;
  linear_f32_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, avdn::to_cuda(stream)>>>(x, ldx, w, y, dy, M, N, K, act, dx,
                                                                                     lddx, dx_accumulate, dw, db);
  return avdn::check_launch("avdn_linear_f32_bwd");
}
