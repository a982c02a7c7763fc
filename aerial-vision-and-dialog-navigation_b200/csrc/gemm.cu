// tcgen05 / TMEM / TMA tensor-core primitive of libavdn.so (sm_100a).
//
// One warp-specialised kernel covers every dense contraction of the hot path:
//
//   PLAIN  D[M,N] = alpha * op(A) . op(B)^T (+bias)(relu)   batched over 2 extra dims
//          -> nn.Linear fwd / dgrad (src/models/enc_vl.py:16-22, ET_haa.py:98-119),
//             QK^T, PV and their backward products inside nn.TransformerEncoderLayer
//   CONV   implicit-GEMM convolution over NHWC bf16: the A tile of every (tap,
//          channel-block) k-step is ONE 4-D TMA box (64 ch, bw, bh, bn) of the
//          activation tensor shifted by the tap offset; TMA out-of-bounds zero
//          fill IS the padding.  Stride-2 layers read one of four parity views.
//          -> nn.Conv2d fwd and dgrad (src/models/dark_net.py:22-28)
//   WGRAD  D[Cout,Cin] += sum_pixels dZ[pix,co] * X[pix+tap,ci]: both operands
//          MN-major 4-D boxes, split-K over pixel tiles, fp32 atomic epilogue
//          -> nn.Conv2d / nn.Linear weight gradients
//
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (single elected lane,
// tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM), warps 2..5 = epilogue
// (tcgen05.ld 32x32b, one TMEM lane quarter each).  smem ring of STAGES x
// (A 16 KB + B BN*128 B), 128-byte swizzle, mbarrier full/empty pairs,
// tcgen05.commit releases a stage and finally signals the epilogue.  Two CTAs
// fit per SM (<= 100 KB smem, <= 256 TMEM columns each) so one CTA's epilogue
// overlaps the other's main loop.
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

namespace {

constexpr int BM = 128;         // UMMA M (cta_group::1)
constexpr int BK = 64;          // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;

// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// smem matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1)
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused
//   MN-major: 64 MN elements per 128-B row, one row per k; 8-k groups SBO apart,
//             64-wide MN atoms LBO apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

struct alignas(64) KernelParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  avdn_gemm_core c;     // plain-data description shared with the host (gemm.h)
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;   // + alignment slack
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_kernel(const __grid_constant__ KernelParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFF + 8 * (2 * STAGES + 1));

  const avdn_gemm_core& c = p.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates -------------------------------------------------
  const int mt = blockIdx.x, nt = blockIdx.y;
  int z = blockIdx.z;
  const int split = z % c.split_k;  z /= c.split_k;
  int z0 = 0, z1 = 0, tap_fixed = 0;
  if (c.mode == AVDN_GEMM_WGRAD) tap_fixed = z;
  else { z0 = z % c.batch0; z1 = z / c.batch0; }
  const int n0 = nt * BN;
  // conv modes: spatial tile of this CTA's rows (CONV) -- WGRAD walks tiles in the k loop
  int w0 = 0, h0 = 0, i0 = 0;
  if (c.mode == AVDN_GEMM_CONV) {
    w0 = (mt % c.tiles_w) * c.box_w;
    h0 = ((mt / c.tiles_w) % c.tiles_h) * c.box_h;
    i0 = (mt / (c.tiles_w * c.tiles_h)) * c.box_n;
  }
  const int m0 = mt * BM;
  // k range of this split
  const int kb_per = (c.num_kb + c.split_k - 1) / c.split_k;
  const int kb_begin = split * kb_per;
  const int kb_end = min(c.num_kb, kb_begin + kb_per);
  const int my_kb = max(0, kb_end - kb_begin);

  // ---- one-time setup -----------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32((const void*)tmem_slot)),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===============================
    if (lane == 0 && my_kb > 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
        mbar_expect_tx(full_bar(stage), c.tx_bytes);
        if (c.mode == AVDN_GEMM_PLAIN) {
          const int k0 = kb * BK;
          if (!A_MN) tma_load_4d(sa, &p.tmA[0], full_bar(stage), k0, m0, z0, z1);
          else {
            tma_load_4d(sa, &p.tmA[0], full_bar(stage), m0, k0, z0, z1);
            tma_load_4d(sa + 8192, &p.tmA[0], full_bar(stage), m0 + 64, k0, z0, z1);
          }
          if (!B_MN) tma_load_4d(sb, &p.tmB[0], full_bar(stage), k0, n0, z0 * c.b_batched, z1 * c.b_batched);
          else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * 8192, &p.tmB[0], full_bar(stage), n0 + 64 * j, k0, z0 * c.b_batched,
                          z1 * c.b_batched);
          }
        } else if (c.mode == AVDN_GEMM_CONV) {
          const int t = kb / c.cblocks, cb = kb - t * c.cblocks;
          const avdn_tap tp = c.taps[t];
          tma_load_4d(sa, &p.tmA[tp.map], full_bar(stage), cb * BK, w0 + tp.d1, h0 + tp.d2, i0);
          tma_load_4d(sb, &p.tmB[0], full_bar(stage), tp.bk + cb * BK, n0, 0, 0);
        } else {  // WGRAD: k-step = one pixel tile (box_w*box_h*box_n == 64 pixels)
          const avdn_tap tp = c.taps[tap_fixed];
          const int pw = (kb % c.tiles_w) * c.box_w;
          const int ph = ((kb / c.tiles_w) % c.tiles_h) * c.box_h;
          const int pn = (kb / (c.tiles_w * c.tiles_h)) * c.box_n;
          tma_load_4d(sa, &p.tmA[0], full_bar(stage), m0, pw, ph, pn);
          tma_load_4d(sa + 8192, &p.tmA[0], full_bar(stage), m0 + 64, pw, ph, pn);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sb + j * 8192, &p.tmB[tp.map], full_bar(stage), n0 + 64 * j, pw + tp.d1, ph + tp.d2, pn);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ================================
    if (lane == 0 && my_kb > 0) {
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < my_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t ad = A_MN ? make_smem_desc(sa + k * (UMMA_K * 128), 8192, 1024)
                                   : make_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t bd = B_MN ? make_smem_desc(sb + k * (UMMA_K * 128), 8192, 1024)
                                   : make_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(empty_bar(stage));            // frees the smem stage when the MMAs retire
        if (kb == my_kb - 1) tcgen05_commit(tmem_full_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ============================== epilogue ================================
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int row = q * 32 + lane;               // row of the 128-row tile
    if (my_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
    }
    // ---- where does this row go? ----
    bool row_ok;
    size_t row_off;                              // element offset of (row, col 0)
    if (c.mode == AVDN_GEMM_CONV) {
      const int dw = row % c.box_w, dh = (row / c.box_w) % c.box_h, dn = row / (c.box_w * c.box_h);
      const int w = w0 + dw, h = h0 + dh, n = i0 + dn;
      row_ok = (dn < c.box_n) && (w < c.valid_w) && (h < c.valid_h) && (n < c.valid_n);
      row_off = (((size_t)n * c.out_H + (size_t)h * c.out_sh + c.out_oh) * c.out_W + (size_t)w * c.out_sw +
                 c.out_ow) * (size_t)c.ldc;
    } else {
      row_ok = (m0 + row) < c.M;
      row_off = (size_t)(m0 + row) * (size_t)c.ldc + (size_t)z0 * c.out_bs0 + (size_t)z1 * c.out_bs1;
      if (c.mode == AVDN_GEMM_WGRAD) row_off += (size_t)c.taps[tap_fixed].bk;
    }
    const float alpha = c.alpha;
#pragma unroll 1
    for (int cc = 0; cc < BN; cc += 32) {
      uint32_t v[32];
      if (my_kb > 0) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
      else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int col0 = n0 + cc;
      if (!row_ok || col0 >= c.N) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]) * alpha;
        if (c.bias && (col0 + j) < c.N) x += __ldg(c.bias + col0 + j);
        if (c.relu) x = fmaxf(x, 0.f);
        f[j] = x;
      }
      if (c.relu_mask) {
        const __nv_bfloat16* mk = reinterpret_cast<const __nv_bfloat16*>(c.relu_mask) + row_off + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((col0 + j) < c.N && !(__bfloat162float(mk[j]) > 0.f)) f[j] = 0.f;
      }
      const bool full = (col0 + 32) <= c.N;
      if (c.out_dtype == AVDN_DT_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(c.out) + row_off + col0;
        if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t w[4];
            if (c.accumulate == 1) {
              const uint4 old = o4[g];
              const uint32_t ow[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 ob = *reinterpret_cast<const __nv_bfloat162*>(&ow[j]);
                f[g * 8 + 2 * j] += __low2float(ob);
                f[g * 8 + 2 * j + 1] += __high2float(ob);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
              w[j] = *reinterpret_cast<const uint32_t*>(&b2);
            }
            o4[g] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
          for (int j = 0; j < 32 && (col0 + j) < c.N; ++j) {
            float x = f[j];
            if (c.accumulate == 1) x += __bfloat162float(o[j]);
            o[j] = __float2bfloat16_rn(x);
          }
        }
      } else {
        float* o = reinterpret_cast<float*>(c.out) + row_off + col0;
        if (c.accumulate == 2) {
          for (int j = 0; j < 32 && (col0 + j) < c.N; ++j) atomicAdd(o + j, f[j]);
        } else if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
          float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 x = make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
            if (c.accumulate == 1) { const float4 old = o4[g]; x.x += old.x; x.y += old.y; x.z += old.z; x.w += old.w; }
            o4[g] = x;
          }
        } else {
          for (int j = 0; j < 32 && (col0 + j) < c.N; ++j) {
            float x = f[j];
            if (c.accumulate == 1) x += o[j];
            o[j] = x;
          }
        }
      }
    }
  }

  // ---- teardown -------------------------------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BN < 32 ? 32 : BN))
                 : "memory");
  }
}

// ----------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int encode_operand(const avdn_operand& o, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return avdn::set_err(AVDN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (o.stride[0] != 1) return avdn::set_err(AVDN_ERR_BAD_ARG, "operand dim 0 must be contiguous");
  if ((reinterpret_cast<uintptr_t>(o.ptr) & 15) != 0)
    return avdn::set_err(AVDN_ERR_BAD_ARG, "operand pointer must be 16-byte aligned");
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    dims[i] = (cuuint64_t)o.dim[i];
    box[i] = (cuuint32_t)o.box[i];
    if (o.dim[i] < 1) return avdn::set_err(AVDN_ERR_BAD_ARG, "operand dim %d < 1", i);
    if (box[i] < 1 || box[i] > 256) return avdn::set_err(AVDN_ERR_BAD_ARG, "TMA box dim %d = %u out of range", i, box[i]);
  }
  if (box[0] != 64) return avdn::set_err(AVDN_ERR_BAD_ARG, "TMA box dim 0 must be 64 bf16 (128-byte swizzle)");
  for (int i = 1; i < 4; ++i) {
    const cuuint64_t s = (cuuint64_t)o.stride[i] * 2;
    if (s % 16 != 0 || s == 0)
      return avdn::set_err(AVDN_ERR_BAD_ARG, "operand stride %d (%llu B) must be a non-zero multiple of 16", i, (unsigned long long)s);
    strides[i - 1] = s;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(o.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return avdn::set_err(AVDN_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return AVDN_OK;
}

struct Plan {
  uint32_t magic;
  int32_t bn, stages, a_mn, b_mn;
  dim3 grid;
  int smem;
  KernelParams kp;
};
constexpr uint32_t PLAN_MAGIC = 0xA7D17C05u;

template <int BN, int STAGES, bool A_MN, bool B_MN>
int launch_t(const Plan& pl, cudaStream_t s) {
  auto kfn = gemm_kernel<BN, STAGES, A_MN, B_MN>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN, STAGES>::TOTAL) !=
        cudaSuccess)
      return avdn::check_launch("cudaFuncSetAttribute(gemm_kernel)");
    attr_done = true;
  }
  kfn<<<pl.grid, NUM_THREADS, SmemLayout<BN, STAGES>::TOTAL, s>>>(pl.kp);
  return avdn::check_launch("gemm_kernel");
}

template <int BN, int STAGES>
int launch_bn(const Plan& pl, cudaStream_t s) {
  if (!pl.a_mn && !pl.b_mn) return launch_t<BN, STAGES, false, false>(pl, s);
  if (!pl.a_mn && pl.b_mn) return launch_t<BN, STAGES, false, true>(pl, s);
  if (pl.a_mn && !pl.b_mn) return launch_t<BN, STAGES, true, false>(pl, s);
  return launch_t<BN, STAGES, true, true>(pl, s);
}

}  // namespace

extern "C" size_t avdn_gemm_plan_bytes(void) { return sizeof(Plan); }

extern "C" int avdn_gemm_plan(const avdn_gemm_desc* d, void* plan_host, size_t plan_bytes) {
  AVDN_REQUIRE(d && plan_host, "avdn_gemm_plan: null pointer");
  AVDN_REQUIRE(plan_bytes >= sizeof(Plan), "avdn_gemm_plan: plan buffer too small (%zu < %zu)", plan_bytes, sizeof(Plan));
  AVDN_REQUIRE(d->core.mode >= AVDN_GEMM_PLAIN && d->core.mode <= AVDN_GEMM_WGRAD, "avdn_gemm_plan: bad mode");
  AVDN_REQUIRE(d->bn == 64 || d->bn == 128 || d->bn == 256, "avdn_gemm_plan: bn must be 64/128/256");
  AVDN_REQUIRE(d->n_a >= 1 && d->n_a <= 4 && d->n_b >= 1 && d->n_b <= 4, "avdn_gemm_plan: 1..4 operand maps");
  AVDN_REQUIRE(d->core.out, "avdn_gemm_plan: null output");
  AVDN_REQUIRE(d->core.split_k >= 1 && d->core.num_kb >= 0, "avdn_gemm_plan: bad k split");
  AVDN_REQUIRE(d->core.split_k == 1 || d->core.accumulate == 2, "avdn_gemm_plan: split-K needs the atomic epilogue");
  AVDN_REQUIRE(d->core.accumulate != 2 || d->core.out_dtype == AVDN_DT_F32, "avdn_gemm_plan: atomic epilogue is fp32 only");
  Plan* pl = reinterpret_cast<Plan*>(plan_host);
  memset(pl, 0, sizeof(Plan));
  pl->bn = d->bn;
  pl->a_mn = d->a_mn;
  pl->b_mn = d->b_mn;
  pl->kp.c = d->core;
  for (int i = 0; i < d->n_a; ++i) {
    int r = encode_operand(d->a[i], &pl->kp.tmA[i]);
    if (r) return r;
  }
  for (int i = 0; i < d->n_b; ++i) {
    int r = encode_operand(d->b[i], &pl->kp.tmB[i]);
    if (r) return r;
  }
  // bytes one k-step deposits in a stage: full boxes, OOB elements are zero-filled and counted
  auto box_bytes = [](const avdn_operand& o) {
    long long b = 2;
    for (int i = 0; i < 4; ++i) b *= o.box[i];
    return b;
  };
  const long long a_bytes = box_bytes(d->a[0]) * (d->a_mn ? 2 : 1);
  const long long b_bytes = box_bytes(d->b[0]) * (d->b_mn ? d->bn / 64 : 1);
  AVDN_REQUIRE(a_bytes <= BM * BK * 2 && b_bytes <= (long long)d->bn * BK * 2, "avdn_gemm_plan: boxes exceed the stage");
  pl->kp.c.tx_bytes = (uint32_t)(a_bytes + b_bytes);
  // stages: keep two CTAs per SM where the tile allows it
  pl->stages = (d->bn == 256) ? 4 : (d->bn == 128 ? 3 : 4);
  pl->smem = 0;
  pl->grid = dim3((unsigned)d->grid_m, (unsigned)d->grid_n, (unsigned)d->grid_z);
  AVDN_REQUIRE(d->grid_m >= 1 && d->grid_n >= 1 && d->grid_z >= 1 && d->grid_n <= 65535 && d->grid_z <= 65535,
               "avdn_gemm_plan: bad grid %d x %d x %d", d->grid_m, d->grid_n, d->grid_z);
  pl->magic = PLAN_MAGIC;
  return AVDN_OK;
}

extern "C" int avdn_gemm_run(const void* plan_host, avdn_stream_t stream) {
  const Plan* pl = reinterpret_cast<const Plan*>(plan_host);
  AVDN_REQUIRE(pl && pl->magic == PLAN_MAGIC, "avdn_gemm_run: not a plan");
  cudaStream_t s = avdn::to_cuda(stream);
  switch (pl->bn) {
    case 64: return launch_bn<64, 4>(*pl, s);
    case 128: return launch_bn<128, 3>(*pl, s);
    case 256: return launch_bn<256, 4>(*pl, s);
  }
  return avdn::set_err(AVDN_ERR_UNSUPPORTED, "avdn_gemm_run: bn %d", pl->bn);
}
