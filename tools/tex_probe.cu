// Probe: is the texture unit's bilinear filter exact enough to reproduce OpenCV's
// fixed-point (INTER_BITS=5) bilinear blend bit-for-bit, and how fast is it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tex_probe tools/tex_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// exactness: one thread per sample; sample = (sx, sy, ax, ay)
__global__ void exact_kernel(cudaTextureObject_t tex, const uint32_t* __restrict__ tile, int H, int W, int pitch_px,
                             long long n, unsigned long long* mism, float* maxdev, int mode) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = hash32((uint32_t)i * 2654435761u + 12345u);
  int sx = (int)(h % (uint32_t)(W + 8)) - 4; h = hash32(h);
  int sy = (int)(h % (uint32_t)(H + 8)) - 4; h = hash32(h);
  int ax = h & 31, ay = (h >> 5) & 31;
  if (mode == 1) { ax = (int)(i & 31); ay = (int)((i >> 5) & 31); }
  const float u = (float)sx + (float)ax * (1.f / 32.f) + 0.5f;
  const float v = (float)sy + (float)ay * (1.f / 32.f) + 0.5f;
  const float4 r = tex2D<float4>(tex, u, v);
  uint32_t p[4];
  for (int k = 0; k < 4; ++k) {
    const int x = sx + (k & 1), y = sy + (k >> 1);
    p[k] = (x >= 0 && x < W && y >= 0 && y < H) ? tile[(size_t)y * pitch_px + x] : 0u;
  }
  const int w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;
  const float rr[4] = {r.x, r.y, r.z, r.w};
  float md = 0.f;
  bool bad = false;
  for (int c = 0; c < 4; ++c) {
    const int sh = 8 * c;
    const int s = (int)((p[0] >> sh) & 255) * w00 + (int)((p[1] >> sh) & 255) * w01 + (int)((p[2] >> sh) & 255) * w10 +
                  (int)((p[3] >> sh) & 255) * w11;
    const int ref = (s + 512) >> 10;
    // candidate reconstruction: floor(r*255 + 513/1024)
    const int got = (int)floorf(fmaf(rr[c], 255.0f, 513.0f / 1024.0f));
    const float dev = fabsf(rr[c] * 261120.0f - (float)s);
    md = fmaxf(md, dev);
    if (got != ref) bad = true;
  }
  if (bad) atomicAdd(mism, 1ull);
  // float atomic max via int compare (non-negative floats)
  atomicMax(reinterpret_cast<int*>(maxdev), __float_as_int(md));
}

// throughput: render 224x224 views with an affine map (fp32 coords), TEX path
__global__ void __launch_bounds__(256) render_tex_kernel(cudaTextureObject_t tex, const float* __restrict__ aff, int P,
                                                         uint8_t* __restrict__ views) {
  __shared__ __align__(16) uint8_t s_view[32 * 224 * 3];
  const int p = blockIdx.x / 7, band = blockIdx.x % 7;
  const float a0 = aff[p * 6 + 0], a1 = aff[p * 6 + 1], a2 = aff[p * 6 + 2];
  const float a3 = aff[p * 6 + 3], a4 = aff[p * 6 + 4], a5 = aff[p * 6 + 5];
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  // warp handles an 8x4 patch; 28 patches across, 8 down per band; 8 warps
  for (int patch = warp; patch < 28 * 8; patch += 8) {
    const int px = (patch % 28) * 8 + (lane & 7);
    const int py = (patch / 28) * 4 + (lane >> 3);
    const float x = (float)px, y = (float)(band * 32 + py);
    float fx = fmaf(a0, x, fmaf(a1, y, a2));
    float fy = fmaf(a3, x, fmaf(a4, y, a5));
    // quantise to 1/32 like the real kernel will
    fx = rintf(fx * 32.f) * (1.f / 32.f) + 0.5f;
    fy = rintf(fy * 32.f) * (1.f / 32.f) + 0.5f;
    const float4 r = tex2D<float4>(tex, fx, fy);
    const int o = (py * 224 + px) * 3;
    s_view[o + 0] = (uint8_t)(int)fmaf(r.x, 255.0f, 513.0f / 1024.0f);
    s_view[o + 1] = (uint8_t)(int)fmaf(r.y, 255.0f, 513.0f / 1024.0f);
    s_view[o + 2] = (uint8_t)(int)fmaf(r.z, 255.0f, 513.0f / 1024.0f);
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(views + ((size_t)p * 224 * 224 + (size_t)band * 32 * 224) * 3);
  const uint4* src = reinterpret_cast<const uint4*>(s_view);
  for (int i = t; i < 32 * 224 * 3 / 16; i += 256) dst[i] = src[i];
}

// same access pattern with 4 LDG gathers (what the current kernel does), no fp64
__global__ void __launch_bounds__(256) render_ldg_kernel(const uint32_t* __restrict__ tile, int pitch, int H, int W,
                                                         const float* __restrict__ aff, int P,
                                                         uint8_t* __restrict__ views) {
  __shared__ __align__(16) uint8_t s_view[32 * 224 * 3];
  const int p = blockIdx.x / 7, band = blockIdx.x % 7;
  const float a0 = aff[p * 6 + 0], a1 = aff[p * 6 + 1], a2 = aff[p * 6 + 2];
  const float a3 = aff[p * 6 + 3], a4 = aff[p * 6 + 4], a5 = aff[p * 6 + 5];
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  for (int patch = warp; patch < 28 * 8; patch += 8) {
    const int px = (patch % 28) * 8 + (lane & 7);
    const int py = (patch / 28) * 4 + (lane >> 3);
    const float x = (float)px, y = (float)(band * 32 + py);
    const int X = __float2int_rn(fmaf(a0, x, fmaf(a1, y, a2)) * 32.f);
    const int Y = __float2int_rn(fmaf(a3, x, fmaf(a4, y, a5)) * 32.f);
    const int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
    uint32_t p00 = 0, p01 = 0, p10 = 0, p11 = 0;
    if ((unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1)) {
      const uint32_t* q = tile + (size_t)sy * pitch + sx;
      p00 = __ldg(q); p01 = __ldg(q + 1); p10 = __ldg(q + pitch); p11 = __ldg(q + pitch + 1);
    }
    const int w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;
    const int o = (py * 224 + px) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int sh = 8 * c;
      const int v = (int)((p00 >> sh) & 255u) * w00 + (int)((p01 >> sh) & 255u) * w01 +
                    (int)((p10 >> sh) & 255u) * w10 + (int)((p11 >> sh) & 255u) * w11;
      s_view[o + c] = (uint8_t)((v + 512) >> 10);
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(views + ((size_t)p * 224 * 224 + (size_t)band * 32 * 224) * 3);
  const uint4* src = reinterpret_cast<const uint4*>(s_view);
  for (int i = t; i < 32 * 224 * 3 / 16; i += 256) dst[i] = src[i];
}

__global__ void fill_kernel(uint32_t* t, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    t[i] = hash32((uint32_t)i + 77u);
}

static cudaTextureObject_t make_tex_pitch(uint32_t* d, int H, int W, int pitch_px) {
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr = d;
  rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
  rd.res.pitch2D.width = W;
  rd.res.pitch2D.height = H;
  rd.res.pitch2D.pitchInBytes = (size_t)pitch_px * 4;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
  return t;
}
static cudaTextureObject_t make_tex_array(cudaArray_t arr) {
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 0;
  cudaTextureObject_t t = 0;
  CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
  return t;
}

int main() {
  const int H = 3000, W = 3000, pitch_px = 3008;   // 12032 B, multiple of 32/128
  int align = 0;
  CK(cudaDeviceGetAttribute(&align, cudaDevAttrTexturePitchAlignment, 0));
  printf("texturePitchAlignment %d\n", align);
  uint32_t* tile;
  CK(cudaMalloc(&tile, (size_t)H * pitch_px * 4));
  fill_kernel<<<1024, 256>>>(tile, (long long)H * pitch_px);
  cudaTextureObject_t tp = make_tex_pitch(tile, H, W, pitch_px);
  cudaArray_t arr;
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
  CK(cudaMallocArray(&arr, &cd, W, H));
  CK(cudaMemcpy2DToArray(arr, 0, 0, tile, (size_t)pitch_px * 4, (size_t)W * 4, H, cudaMemcpyDeviceToDevice));
  cudaTextureObject_t ta = make_tex_array(arr);

  unsigned long long* mism; float* maxdev;
  CK(cudaMalloc(&mism, 8)); CK(cudaMalloc(&maxdev, 4));
  for (int which = 0; which < 2; ++which)
    for (int mode = 0; mode < 2; ++mode) {
      CK(cudaMemset(mism, 0, 8)); CK(cudaMemset(maxdev, 0, 4));
      const long long n = 1ll << 28;
      exact_kernel<<<(unsigned)((n + 255) / 256), 256>>>(which ? ta : tp, tile, H, W, pitch_px, n, mism, maxdev, mode);
      CK(cudaDeviceSynchronize());
      unsigned long long hm; float hd;
      CK(cudaMemcpy(&hm, mism, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&hd, maxdev, 4, cudaMemcpyDeviceToHost));
      printf("exactness %s mode %d: samples %lld mismatches %llu max |r*261120 - sum| = %g\n", which ? "array" : "pitch2D",
             mode, n, hm, hd);
    }

  // throughput
  const int P = 4096;
  std::vector<float> aff(P * 6);
  srand(1);
  for (int p = 0; p < P; ++p) {
    const double cx = 400 + 2200.0 * rand() / RAND_MAX, cy = 400 + 2200.0 * rand() / RAND_MAX;
    const double side = 133 + 1200.0 * rand() / RAND_MAX, th = 6.2831853 * rand() / RAND_MAX;
    const double s = side / 223.0, c_ = cos(th) * s, s_ = sin(th) * s;
    aff[p * 6 + 0] = (float)c_; aff[p * 6 + 1] = (float)-s_; aff[p * 6 + 2] = (float)(cx - 111.5 * c_ + 111.5 * s_);
    aff[p * 6 + 3] = (float)s_; aff[p * 6 + 4] = (float)c_;  aff[p * 6 + 5] = (float)(cy - 111.5 * s_ - 111.5 * c_);
  }
  float* daff; uint8_t* views;
  CK(cudaMalloc(&daff, aff.size() * 4)); CK(cudaMemcpy(daff, aff.data(), aff.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&views, (size_t)P * 224 * 224 * 3));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int which = 0; which < 3; ++which) {
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      CK(cudaEventRecord(e0));
      if (which == 0) render_tex_kernel<<<P * 7, 256>>>(tp, daff, P, views);
      else if (which == 1) render_tex_kernel<<<P * 7, 256>>>(ta, daff, P, views);
      else render_ldg_kernel<<<P * 7, 256>>>(tile, pitch_px, H, W, daff, P, views);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2 && ms < best) best = ms;
    }
    const double bytes = (double)P * 150528 + 27e6;
    printf("render %s: %.3f ms  -> %.1f GB/s algorithmic, %.2f Gpx/s\n",
           which == 0 ? "tex pitch2D" : which == 1 ? "tex array" : "4x LDG fp32", best, bytes / best / 1e6,
           (double)P * 224 * 224 / best / 1e6);
  }
  return 0;
}
