// tcgen05 / TMEM / TMA tensor-core primitive of libavdn.so (sm_100a).
//
// One persistent, warp-specialised kernel covers every dense contraction of the hot path:
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
//          MN-major 4-D boxes, split-K over pixel tiles, fp32 reduce-add epilogue
//          -> nn.Conv2d / nn.Linear weight gradients
//
// Structure (one CTA -- or one CTA pair, cta_group::2 -- per SM, looping over tiles):
//   warp 0      TMA producer: fills a ring of smem stages (128-byte swizzle), running ahead across tile
//               boundaries.  A stage holds up to KPS k-blocks (A 16 KB + B each): one mbarrier round trip and
//               one tcgen05.commit cost ~300 cycles of the issuing thread, so thin k-blocks are batched.
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma kind::f16 (bf16 x bf16 -> fp32)
//               into one of TWO TMEM accumulators; tcgen05.commit frees smem stages and hands
//               the finished accumulator to the epilogue
//   warps 2..5  epilogue: tcgen05.ld (one TMEM lane quarter each) -> alpha/bias/ReLU -> bf16|fp32
//               -> 128-byte-swizzled smem slab -> TMA store (or TMA reduce-add for the
//               accumulating modes); the epilogue of tile i overlaps the main loop of tile i+1.
//               Slabs rotate through up to 4 buffers: a bulk store needs ~2500 cycles until its
//               shared-memory source has been read.
//               Optional fused BatchNorm statistics: per-channel sum / sum of squares of the
//               rounded outputs, accumulated per CTA and committed with fp64 atomics
//               (nn.BatchNorm2d batch statistics, dark_net.py:31).
//   cta_group::2: the pair computes a 256 x BN tile; each CTA loads its own 128 rows of A and
//               HALF of B, so the L2->SM operand traffic per FLOP drops by a third.
#include "common.cuh"
#include "tcgen05.cuh"
#include "conv3_halo.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

namespace {

using namespace avdn_tc;

constexpr int BM = 128;         // rows per CTA (UMMA M = 128 * CTAS)
constexpr int BK = 64;          // default k-block: 64 bf16 = one 128-byte swizzle row (BKT = 32: 64-byte rows,
                                // 64-byte swizzle, for the 32-channel activations of trunk blocks 0 and 2)
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int SLAB_BYTES = BM * 128;      // one epilogue slab: 128 rows x 128 bytes

// ------------------------------------------------------------------ PTX glue
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

constexpr int MAX_STAGES = 12;
constexpr int MAX_SLABS = 4;

struct alignas(64) KernelParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB[4];
  CUtensorMap tmC[4];   // output map (out_tma == 1); one per output class (avdn_gemm_core.n_classes)
  avdn_gemm_core c;     // plain-data description shared with the host (avdn.h)
  int32_t grid_m, grid_n, grid_z;   // tile space (grid_m counts 128-row tiles)
  int32_t out_tma;      // 1: epilogue goes through smem slabs + TMA store / reduce-add
  int32_t rows_in_box;  // rows of a tile that exist (CONV: bw*bh*bn <= 128; otherwise 128)
  int32_t stages;       // smem ring depth (<= MAX_STAGES)
  int32_t kps;          // k-blocks per stage (1..4)
  int32_t slabs;        // epilogue slab buffers (2..MAX_SLABS)
  int32_t dbg;          // timing experiments (AVDN_GEMM_DBG): 1 no bulk store, 2 no TMEM load, 4 no staging stores
  long long* dbg_out;   // NULL, or a device buffer receiving a clock64 trace of CTA 0 (AVDN_GEMM_DBG_BUF;
                        // tools/gemm_epilogue_probe.py)
};

// smem: [stages x kps x (A|B)] [slabs x 16 KB] [float 2 x BN] [barriers] [tmem slot] [BNB: uint2 row table x 128]
template <int BN, int CTAS, int BKT>
struct Tiles {
  static constexpr int A_BYTES = BM * BKT * 2;                 // 16 KB (8 KB for 32-element k-blocks)
  static constexpr int B_BYTES = (BN / CTAS) * BKT * 2;        // this CTA's part of B
  static constexpr int SUB_BYTES = A_BYTES + B_BYTES;          // one k-block
  static constexpr int NUM_BARS = 2 * MAX_STAGES + 4;
  static constexpr int STAT_GRP = (BN == 32) ? 8 : 4;          // row groups of the statistics pass: one accumulator set each
  static constexpr int STAT_BYTES = STAT_GRP * 2 * BN * 4;
  static constexpr int TAIL_BYTES = STAT_BYTES + NUM_BARS * 8 + 16 + 1024;   // stats + barriers + slot + align slack
};

// Thin tiles (BN <= 64: the <= 64-channel blocks of the trunk) are bound by the latency of one CTA's
// producer -> MMA -> epilogue chain, not by any throughput: two CTAs per SM (half the shared memory each, 2 x 2 x BN
// TMEM columns) interleave two such chains.
// BNB: the epilogue also accumulates the BatchNorm-backward reductions of the block that produced the tensor this
// launch is the data gradient of (avdn_gemm_core.bnb_*).
template <int BN, bool A_MN, bool B_MN, int CTAS, int BKT, bool BNB>
__global__ void __launch_bounds__(NUM_THREADS, (BN <= 128) ? 2 : 1)
gemm_kernel(const __grid_constant__ KernelParams p) {
  static_assert(BKT == 64 || (BKT == 32 && !A_MN && !B_MN), "32-element k-blocks: K-major operands only");
  static_assert(!BNB || (!A_MN && !B_MN), "fused BatchNorm-backward statistics: K-major (CONV) operands only");
  using L = Tiles<BN, CTAS, BKT>;
  constexpr int BNH = BN / CTAS;                                // B columns this CTA loads
  constexpr uint32_t KLAYOUT = (BKT == 64) ? 2u : 4u;          // swizzle mode of the K-major operands
  constexpr uint32_t KSBO = 8u * BKT * 2u;                     // 8 rows of BKT bf16
  avdn_pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int STAGES = p.stages, KPS = p.kps, NSLAB = p.slabs;
  const uint32_t stage_bytes = (uint32_t)KPS * L::SUB_BYTES;
  const uint32_t slab_off = (uint32_t)STAGES * stage_bytes;
  const uint32_t stat_off = slab_off + (uint32_t)NSLAB * SLAB_BYTES;
  const uint32_t bar_off = stat_off + L::STAT_BYTES;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * L::NUM_BARS);
  float* s_stat = reinterpret_cast<float*>(smem_gen + stat_off);

  const avdn_gemm_core& c = p.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);

  // ---- tile space: pair-tiles when CTAS == 2 (two adjacent 128-row tiles share one B) -----
  // Tile index t = nt + grid_n * (mtp + pm * z), visited as t = t_begin, t_begin + t_step, ...
  // The three roles walk it with carry arithmetic (no divisions on the per-tile path: the roles are
  // single threads / single warps and a 64-bit division costs more than a whole k-step).
  const uint32_t pm = (uint32_t)(p.grid_m + CTAS - 1) / CTAS;
  const uint32_t gn = (uint32_t)p.grid_n;
  const uint32_t total = pm * gn * (uint32_t)p.grid_z;
  // Output classes (CONV, n_classes > 1): the launch index of a CTA (pair) is class-fastest, NC * r + q.  The host
  // makes the number of CTA (pair)s a multiple of NC, so q is a constant of the CTA and all NC tiles of a spatial
  // tile r are visited in the same iteration by NC neighbouring CTAs; the class of the tile is (q + iteration) % NC:
  // a bijection for every r, and every CTA cycles through the classes (their work differs: 1, 2, 2, 4 taps).
  const uint32_t NC = (c.mode == AVDN_GEMM_CONV && c.n_classes > 1) ? (uint32_t)c.n_classes : 1u;
  const uint32_t cq = (blockIdx.x / CTAS) % NC;
  const uint32_t t_begin = (blockIdx.x / CTAS) / NC, t_step = (gridDim.x / CTAS) / NC;
  const int kb_per = (c.num_kb + c.split_k - 1) / c.split_k;

  struct Tile { int mt, nt, z0, z1, tap, kb_begin, my_kb, cls; };
  struct Walk {
    uint32_t t, nt, mtp, z;          // current tile
    uint32_t dn, dm, dz;             // t_step decomposed
    uint32_t zc; int z0, z1, tap, kb_begin, my_kb;   // cached decomposition of z
    uint32_t it;                     // iteration count (class rotation)
  };
  auto walk_init = [&]() {
    Walk w;
    w.t = t_begin;
    w.nt = t_begin % gn;
    uint32_t r = t_begin / gn;
    w.mtp = r % pm;
    w.z = r / pm;
    w.dn = t_step % gn;
    r = t_step / gn;
    w.dm = r % pm;
    w.dz = r / pm;
    w.zc = 0xFFFFFFFFu;
    w.z0 = w.z1 = w.tap = w.kb_begin = w.my_kb = 0;
    w.it = 0;
    return w;
  };
  auto walk_tile = [&](Walk& w) {
    if (w.z != w.zc) {               // z changes once per (pm * grid_n) tiles
      w.zc = w.z;
      int z = (int)w.z, split;
      w.z0 = w.z1 = w.tap = 0;
      if (c.mode == AVDN_GEMM_WGRAD) {
        // taps fastest: the CTAs running side by side sweep the SAME pixel range for all filter
        // taps, so dZ and X are fetched from HBM once and re-read from L2 by the other taps
        w.tap = z % c.n_taps;
        split = z / c.n_taps;
      } else {
        split = z % c.split_k;  z /= c.split_k;
        w.z0 = z % c.batch0; w.z1 = z / c.batch0;
      }
      w.kb_begin = split * kb_per;
      const int kb_end = min(c.num_kb, w.kb_begin + kb_per);
      w.my_kb = max(0, kb_end - w.kb_begin);
    }
    Tile T;
    T.nt = (int)w.nt;
    T.mt = (int)w.mtp * CTAS + (int)rank;
    T.z0 = w.z0; T.z1 = w.z1; T.tap = w.tap; T.kb_begin = w.kb_begin; T.my_kb = w.my_kb;
    T.cls = 0;
    if (NC > 1) {                    // this class's taps: T.tap is their first index, all of them in one pass
      T.cls = (int)((cq + w.it) % NC);
      T.tap = c.cls_tap0[T.cls];
      T.kb_begin = 0;
      T.my_kb = (c.cls_tap0[T.cls + 1] - T.tap) * c.cblocks;
    }
    return T;
  };
  auto walk_next = [&](Walk& w) {
    ++w.it;
    w.t += t_step;
    w.nt += w.dn;
    uint32_t carry = (w.nt >= gn) ? 1u : 0u;
    w.nt -= carry * gn;
    w.mtp += w.dm + carry;
    carry = (w.mtp >= pm) ? 1u : 0u;
    w.mtp -= carry * pm;
    w.z += w.dz + carry;
  };

  // ---- one-time setup -----------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * CTAS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32((const void*)tmem_slot)),
                   "r"((uint32_t)(2 * BN))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32((const void*)tmem_slot)),
                   "r"((uint32_t)(2 * BN))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < L::STAT_GRP * 2 * BN; i += EPI_THREADS) s_stat[i] = 0.f;
  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();          // peer barriers initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  avdn_pdl_wait();                             // barriers, TMEM and tables are set up: now wait for the inputs

  if (warp == 0) {
    // =========================== TMA producer ===============================
    // the whole warp walks the loops (uniform control flow); one elected lane issues
    {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_kb = c.tx_bytes;                    // bytes both CTAs deposit per k-block
      const uint32_t twh = (uint32_t)(c.tiles_w * c.tiles_h);
      for (Walk wk = walk_init(); wk.t < total; walk_next(wk)) {
        const Tile T = walk_tile(wk);
        if (T.my_kb == 0) continue;
        const int n0 = T.nt * BN + (int)rank * BNH;         // this CTA's slice of B
        const int m0 = T.mt * BM;
        int w0 = 0, h0 = 0, i0 = 0;
        // k-block cursors (advanced incrementally: no division inside the k loop)
        int cb = 0, tp_i = 0, k0 = 0, pw = 0, ph = 0, pn = 0;
        avdn_tap tp = c.taps[0];
        if (c.mode == AVDN_GEMM_CONV) {
          const uint32_t mt = (uint32_t)T.mt;
          const uint32_t q = mt / (uint32_t)c.tiles_w, in = mt / twh;
          w0 = (int)(mt - q * c.tiles_w) * c.box_w;
          h0 = (int)(q - in * c.tiles_h) * c.box_h;
          i0 = (int)in * c.box_n;                           // phantom tile of an odd pair: out of bounds -> zeros
          tp_i = T.kb_begin / c.cblocks;
          cb = T.kb_begin - tp_i * c.cblocks;
          if (NC > 1) tp_i += T.tap;                        // first tap of this tile's output class
          tp = c.taps[tp_i];
        } else if (c.mode == AVDN_GEMM_WGRAD) {
          tp = c.taps[T.tap];
          const uint32_t kb = (uint32_t)T.kb_begin;
          const uint32_t q = kb / (uint32_t)c.tiles_w, in = kb / twh;
          pw = (int)(kb - q * c.tiles_w) * c.box_w;
          ph = (int)(q - in * c.tiles_h) * c.box_h;
          pn = (int)in * c.box_n;
        } else {
          k0 = T.kb_begin * BKT;
        }
        const int zb0 = T.z0 * c.b_batched, zb1 = T.z1 * c.b_batched;
        for (int rem = T.my_kb; rem > 0;) {
          const int nkb = rem < KPS ? rem : KPS;
          rem -= nkb;
          mbar_wait(empty_bar(stage), phase ^ 1);
          uint32_t fb = full_bar(stage);
          if (CTAS == 2) fb = mapa_rank(fb, 0);
          const bool issuer = elect_one();
          if (leader && issuer) mbar_expect_tx(full_bar(stage), tx_kb * (uint32_t)nkb);
          uint32_t sa = smem_base + stage * stage_bytes;
          for (int j = 0; j < nkb; ++j, sa += L::SUB_BYTES) {
            const uint32_t sb = sa + L::A_BYTES;
            if (c.mode == AVDN_GEMM_PLAIN) {
              if (!A_MN) {
                if (issuer) tma_load_4d<CTAS>(sa, &p.tmA[0], fb, k0, m0, T.z0, T.z1);
              } else {
                if (issuer) tma_load_4d<CTAS>(sa, &p.tmA[0], fb, m0, k0, T.z0, T.z1);
                if (issuer) tma_load_4d<CTAS>(sa + 8192, &p.tmA[0], fb, m0 + 64, k0, T.z0, T.z1);
              }
              if (!B_MN) {
                if (issuer) tma_load_4d<CTAS>(sb, &p.tmB[0], fb, k0, n0, zb0, zb1);
              } else {
#pragma unroll
                for (int jj = 0; jj < BNH / 64; ++jj)
                  if (issuer) tma_load_4d<CTAS>(sb + jj * 8192, &p.tmB[0], fb, n0 + 64 * jj, k0, zb0, zb1);
              }
              k0 += BKT;
            } else if (c.mode == AVDN_GEMM_CONV) {
              if (issuer) tma_load_4d<CTAS>(sa, &p.tmA[tp.map], fb, cb * BKT, w0 + tp.d1, h0 + tp.d2, i0);
              if (issuer) tma_load_4d<CTAS>(sb, &p.tmB[0], fb, tp.bk + cb * BKT, n0, 0, 0);
              if (++cb == c.cblocks) { cb = 0; ++tp_i; tp = c.taps[tp_i < c.n_taps ? tp_i : 0]; }
            } else {  // WGRAD: k-block = one pixel tile (box_w*box_h*box_n == 64 pixels)
              if (issuer) tma_load_4d<CTAS>(sa, &p.tmA[0], fb, m0, pw, ph, pn);
              if (issuer) tma_load_4d<CTAS>(sa + 8192, &p.tmA[0], fb, m0 + 64, pw, ph, pn);
#pragma unroll
              for (int jj = 0; jj < BNH / 64; ++jj)
                if (issuer) tma_load_4d<CTAS>(sb + jj * 8192, &p.tmB[tp.map], fb, n0 + 64 * jj, pw + tp.d1, ph + tp.d2, pn);
              pw += c.box_w;
              if (pw >= c.tiles_w * c.box_w) {
                pw = 0; ph += c.box_h;
                if (ph >= c.tiles_h * c.box_h) { ph = 0; pn += c.box_n; }
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ================================
    // the whole warp of the leader CTA walks the loops; one elected lane issues the MMAs and commits
    if (leader) {
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((BM * CTAS) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      int trace_n = 0;
      // descriptors of stage 0, k = 0; the start-address field (bits 0..13, 16-byte units) is advanced by
      // plain additions: + stage * stage_bytes / 16, + j * SUB_BYTES / 16, + k * (bytes per UMMA_K) / 16
      const uint64_t ad0 = A_MN ? make_smem_desc(smem_base, 8192, 1024)
                                : make_smem_desc(smem_base, 16, KSBO, KLAYOUT);
      const uint64_t bd0 = B_MN ? make_smem_desc(smem_base + L::A_BYTES, 8192, 1024)
                                : make_smem_desc(smem_base + L::A_BYTES, 16, KSBO, KLAYOUT);
      constexpr uint32_t AK = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      constexpr uint32_t BKS = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      const uint32_t stage_units = stage_bytes >> 4;
      for (Walk wk = walk_init(); wk.t < total; walk_next(wk)) {
        const Tile T = walk_tile(wk);
        if (T.my_kb == 0) continue;
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        ++it;
        mbar_wait(tempty_bar(as), aphase ^ 1);              // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        uint32_t first = 0u;                                 // 0 for the very first MMA of the tile
        for (int rem = T.my_kb; rem > 0;) {
          const int nkb = rem < KPS ? rem : KPS;
          rem -= nkb;
          const bool trace = p.dbg_out && blockIdx.x == 0 && trace_n < 256;   // (all lanes count, lane 0 writes)
          long long tc0 = 0, tc1 = 0;
          if (trace) tc0 = clock64();
          mbar_wait(full_bar(stage), phase);
          if (trace) tc1 = clock64();
          tcgen05_fence_after();
          if (elect_one()) {
            uint64_t ad = ad0 + (uint64_t)((uint32_t)stage * stage_units), bd = bd0 + (uint64_t)((uint32_t)stage * stage_units);
            uint32_t acc = first;
            for (int j = 0; j < nkb; ++j) {
#pragma unroll
              for (int k = 0; k < BKT / UMMA_K; ++k) {
                umma_bf16<CTAS>(tmem_d, ad + (uint64_t)(k * AK), bd + (uint64_t)(k * BKS), idesc, acc);
                acc = 1u;
              }
              ad += (uint64_t)(L::SUB_BYTES >> 4);
              bd += (uint64_t)(L::SUB_BYTES >> 4);
            }
            tcgen05_commit<CTAS>(empty_bar(stage));         // frees the smem stage when the MMAs retire
            if (rem == 0) tcgen05_commit<CTAS>(tfull_bar(as));
          }
          __syncwarp();
          first = 1u;
          if (trace && lane == 0) {
            p.dbg_out[trace_n * 4 + 0] = tc0; p.dbg_out[trace_n * 4 + 1] = tc1;
            p.dbg_out[trace_n * 4 + 2] = clock64(); p.dbg_out[trace_n * 4 + 3] = nkb;
          }
          if (trace) ++trace_n;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ============================== epilogue ================================
    const int q = warp & 3;                      // TMEM lane quarter of this warp
    const int row = q * 32 + lane;               // row of the 128-row tile
    const int et = threadIdx.x - 64;             // 0..127
    const bool is_bf16 = (c.out_dtype == AVDN_DT_BF16);
    const float alpha = c.alpha;
    const bool has_alpha = (alpha != 1.0f);
    const float* __restrict__ bias = c.bias;
    uint32_t it = 0, slab_ctr = 0;
    int stat_nt = -1;
    int etrace_n = 0;
    const bool etrace = p.dbg_out && blockIdx.x == 0 && et == 0;
    long long* eout = p.dbg_out ? p.dbg_out + 2048 : nullptr;
#define ETR(slot) do { if (etrace && etrace_n < 120) eout[etrace_n * 16 + (slot)] = clock64(); } while (0)
    auto flush_stats = [&]() {
      // all epilogue threads: commit the CTA's running column sums of tile column `stat_nt`
      epi_bar_sync();
      for (int i = et; i < 2 * BN; i += EPI_THREADS) {
        const int j = i % BN, which = i / BN;
        const int col = stat_nt * BN + j;
        float v = 0.f;                               // the row groups' partial sums, always in the same order
#pragma unroll
        for (int g = 0; g < L::STAT_GRP; ++g) { v += s_stat[g * 2 * BN + i]; s_stat[g * 2 * BN + i] = 0.f; }
        if (col < c.N) atomicAdd((BNB ? c.bnb_sums : c.stats) + (size_t)which * c.N + col, (double)v);
      }
      epi_bar_sync();
    };
    // The epilogue warps run one per scheduler: their speed is the single-warp issue rate (~4 cycles per
    // dependent instruction), so the common case (alpha = 1, no bias, no ReLU: every convolution) must not
    // pay per-element predicates.  Rows of a tile beyond rows_in_box hold garbage: they are neither stored
    // (the TMA box ends before them) nor counted by the statistics pass.
    const bool need_fin = has_alpha || (bias != nullptr) || (c.relu != 0);
    // Fused eval-mode BatchNorm + LeakyReLU (+ shortcut): per-column scale/shift of the current tile column
    // live in s_stat (the statistics scratch; the two uses exclude each other).
    const bool has_aff = (c.col_scale != nullptr);
    const float slope = c.leaky_slope;
    const __nv_bfloat16* res_row = nullptr;
    int aff_nt = -1;
    auto finish_aff = [&](uint32_t (&v)[32], int col0, int cc) {
      const float4* sc4 = reinterpret_cast<const float4*>(s_stat + cc);
      const float4* sh4 = reinterpret_cast<const float4*>(s_stat + BN + cc);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 sc = sc4[g], sh = sh4[g];
        float x0 = fmaf(__uint_as_float(v[4 * g + 0]), sc.x, sh.x);
        float x1 = fmaf(__uint_as_float(v[4 * g + 1]), sc.y, sh.y);
        float x2 = fmaf(__uint_as_float(v[4 * g + 2]), sc.z, sh.z);
        float x3 = fmaf(__uint_as_float(v[4 * g + 3]), sc.w, sh.w);
        v[4 * g + 0] = __float_as_uint(fmaxf(x0, x0 * slope));
        v[4 * g + 1] = __float_as_uint(fmaxf(x1, x1 * slope));
        v[4 * g + 2] = __float_as_uint(fmaxf(x2, x2 * slope));
        v[4 * g + 3] = __float_as_uint(fmaxf(x3, x3 * slope));
      }
      if (res_row) {
        const uint4* rp = reinterpret_cast<const uint4*>(res_row + col0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 q = __ldg(rp + g);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[8 * g + 2 * j] = __float_as_uint(__uint_as_float(v[8 * g + 2 * j]) + __uint_as_float(w[j] << 16));
            v[8 * g + 2 * j + 1] =
                __float_as_uint(__uint_as_float(v[8 * g + 2 * j + 1]) + __uint_as_float(w[j] & 0xFFFF0000u));
          }
        }
      }
    };
    auto finish32 = [&](uint32_t (&v)[32], int col0) {
      const bool full = (col0 + 32) <= c.N;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]) * alpha;
        if (bias && (full || (col0 + j) < c.N)) x += __ldg(bias + col0 + j);
        if (c.relu) x = fmaxf(x, 0.f);
        v[j] = __float_as_uint(x);
      }
    };
    const uint32_t twh = (uint32_t)(c.tiles_w * c.tiles_h);
    // ---- fused BatchNorm-backward statistics (BNB): tile-invariant part ----
    // Element offset of output pixel (n, h, w) = OB + n*SN + h*SH + w*SW (+ column); `bnb_z` is addressed like `out`.
    // The position (wi, hi, ni) of a tile row inside the box does not depend on the tile: s_row[r] holds its element
    // offset relative to the box origin and its packed coordinates (0xFFFFFFFF: the row is not part of the box).
    // Statistics mapping: a warp owns 32 rows; lane -> (16-byte chunk cg of the slab row = 8 columns, row set rs);
    // the thread visits rows q*32 + ST_RS*i + rs, i < ST_NI, so one warp instruction reads ST_RS whole rows.
    constexpr int ST_CH = (BN == 32) ? 4 : 8;                    // 16-byte chunks per slab row
    constexpr int ST_RS = 32 / ST_CH;                            // rows per warp instruction (4 or 8)
    constexpr int ST_NI = 32 / ST_RS;                            // rows per thread (8 or 4)
    const __nv_bfloat16* bnb_z = reinterpret_cast<const __nv_bfloat16*>(c.bnb_z);
    const __nv_bfloat16* bnb_o = reinterpret_cast<const __nv_bfloat16*>(c.out);
    uint2* s_row = reinterpret_cast<uint2*>(smem_gen + bar_off + 8 * L::NUM_BARS + 16);
    uint32_t SW = 0, SH = 0, SN = 0, OB = 0;
    if (BNB) {
      SW = (uint32_t)c.out_sw * (uint32_t)c.ldc;
      SH = (uint32_t)c.out_sh * (uint32_t)c.out_W * (uint32_t)c.ldc;
      SN = (uint32_t)c.out_H * (uint32_t)c.out_W * (uint32_t)c.ldc;
      OB = ((uint32_t)c.out_oh * (uint32_t)c.out_W + (uint32_t)c.out_ow) * (uint32_t)c.ldc;
      const int t2 = et / c.box_w;
      const int wi = et - t2 * c.box_w, ni = t2 / c.box_h, hi = t2 - ni * c.box_h;
      s_row[et] = (et < p.rows_in_box)
                      ? make_uint2((uint32_t)ni * SN + (uint32_t)hi * SH + (uint32_t)wi * SW,
                                   (uint32_t)wi | ((uint32_t)hi << 10) | ((uint32_t)ni << 20))
                      : make_uint2(0u, 0xFFFFFFFFu);
      epi_bar_sync();
    }
    // row r of the box whose origin is (w0, h0, i0): inside the tensor?
    auto row_ok = [&](uint32_t xy, int w0_, int h0_, int i0_) {
      return xy != 0xFFFFFFFFu && w0_ + (int)(xy & 1023u) < c.valid_w && h0_ + (int)((xy >> 10) & 1023u) < c.valid_h &&
             i0_ + (int)(xy >> 20) < c.valid_n;
    };
    // origin of the output class of a tile (its parity offset inside the strided output)
    auto cls_origin = [&](int cls) -> uint32_t {
      return (NC > 1) ? ((uint32_t)c.cls_oh[cls] * (uint32_t)c.out_W + (uint32_t)c.cls_ow[cls]) * (uint32_t)c.ldc : OB;
    };
    auto prefetch_tile = [&](int mt_, int nt_, int cls_) {
      const uint32_t mt = (uint32_t)mt_;
      const uint32_t qq = mt / (uint32_t)c.tiles_w, in = mt / twh;
      const int w0_ = (int)(mt - qq * c.tiles_w) * c.box_w, h0_ = (int)(qq - in * c.tiles_h) * c.box_h;
      const int i0_ = (int)in * c.box_n;
      const uint2 rw = s_row[et];
      if (row_ok(rw.y, w0_, h0_, i0_)) {
        const int n0_ = nt_ * BN;
        const uint32_t off = cls_origin(cls_) + (uint32_t)i0_ * SN + (uint32_t)h0_ * SH + (uint32_t)w0_ * SW + rw.x + (uint32_t)n0_;
        const int ncols = min(BN, c.N - n0_);
        for (int b = 0; b < ncols; b += 64) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(bnb_z + off + b));
          if (c.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(bnb_o + off + b));
        }
      }
    };
    for (Walk wk = walk_init(); wk.t < total; walk_next(wk)) {
      const Tile T = walk_tile(wk);
      if (T.my_kb == 0) continue;
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      ++it;
      const int n0 = T.nt * BN, m0 = T.mt * BM;
      int w0 = 0, h0 = 0, i0 = 0;
      if (c.mode == AVDN_GEMM_CONV) {
        const uint32_t mt = (uint32_t)T.mt;
        const uint32_t qq = mt / (uint32_t)c.tiles_w, in = mt / twh;
        w0 = (int)(mt - qq * c.tiles_w) * c.box_w;
        h0 = (int)(qq - in * c.tiles_h) * c.box_h;
        i0 = (int)in * c.box_n;
      }
      if ((BNB || c.stats) && stat_nt != T.nt) {
        if (stat_nt >= 0) flush_stats();
        stat_nt = T.nt;
      }
      if (BNB) {
        // pull the rows of z (and of the old dA) of the NEXT tile into L2 while this one is processed (the first
        // tile prefetches itself as well): the statistics loop below then waits for L2 hits, not for HBM
        Walk wn = wk;
        if (it == 1) prefetch_tile(T.mt, T.nt, T.cls);
        walk_next(wn);
        if (wn.t < total) {
          const Tile Tn = walk_tile(wn);
          prefetch_tile(Tn.mt, Tn.nt, Tn.cls);
        }
      }
      if (has_aff) {
        if (aff_nt != T.nt) {
          epi_bar_sync();                                  // readers of the previous tile column are done
          for (int i = et; i < 2 * BN; i += EPI_THREADS) {
            const int j = (i < BN) ? i : i - BN, col = n0 + j;
            s_stat[i] = (col < c.N) ? __ldg((i < BN ? c.col_scale : c.col_shift) + col) : 0.f;
          }
          epi_bar_sync();
          aff_nt = T.nt;
        }
        res_row = nullptr;
        if (c.residual) {
          const int wi = row % c.box_w, t2 = row / c.box_w;
          const int hi = t2 % c.box_h, ni = t2 / c.box_h;
          const int w = w0 + wi, h = h0 + hi, n = i0 + ni;
          if (row < p.rows_in_box && w < c.valid_w && h < c.valid_h && n < c.valid_n)
            res_row = reinterpret_cast<const __nv_bfloat16*>(c.residual) +
                      ((size_t)((size_t)n * c.out_H + h) * c.out_W + w) * (size_t)c.ldc;
        }
      }
      ETR(0);
      mbar_wait(tfull_bar(as), aphase);
      tcgen05_fence_after();
      ETR(1);
      const uint32_t tmem_acc = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);

      if (p.out_tma) {
        // ---- TMEM -> registers -> swizzled smem slab -> TMA store / reduce-add ----
        constexpr bool SLAB64 = (BN == 32);       // bf16 tile of 32 columns: 64-byte slab rows, 64-byte swizzle
        // one slab = 128 bytes per row: 64 bf16 columns (two 32-column TMEM chunks) or 32 fp32 columns
        auto issue_store = [&](uint8_t* slab, int scol) {
          if (et == 0) {
            const uint32_t srca = smem_u32(slab);
            int c0 = scol, c1, c2, c3;
            if (c.mode == AVDN_GEMM_CONV) { c1 = w0; c2 = h0; c3 = i0; }
            else if (c.mode == AVDN_GEMM_WGRAD) { c0 = scol + c.taps[T.tap].bk; c1 = m0; c2 = 0; c3 = 0; }
            else { c1 = m0; c2 = T.z0; c3 = T.z1; }
            if (!(p.dbg & 1)) {
              if (c.accumulate && !BNB) tma_reduce_add_4d(&p.tmC[T.cls], srca, c0, c1, c2, c3);
              else tma_store_4d(&p.tmC[T.cls], srca, c0, c1, c2, c3);
            }
            tma_commit_group();
          }
        };
        auto acquire_slab = [&]() {
          // the bulk store that read this buffer NSLAB slabs ago must have finished reading it
          uint8_t* slab = smem_gen + slab_off + slab_ctr * SLAB_BYTES;
          if (et == 0 && !(p.dbg & 32)) tma_wait_group_read_n(NSLAB - 1);
          if (!(p.dbg & 16)) epi_bar_sync();
          return slab;
        };
        if (is_bf16) {
#pragma unroll 1
          for (int cc = 0; cc < BN; cc += 64) {
            const int col0 = n0 + cc;
            if (col0 >= c.N) break;                                 // uniform: whole slab out of range
            const bool two = (!SLAB64) && (col0 + 32 < c.N);
            uint32_t va[32], vb[32];
            if (!(p.dbg & 2)) {
              tmem_ld32_async(tmem_acc + (uint32_t)cc, va);           // both chunks in flight while we wait for
              if (two) tmem_ld32_async(tmem_acc + (uint32_t)cc + 32u, vb);   // the slab buffer
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { va[j] = 0x3f800000u + j + cc; vb[j] = 0x3f900000u + j; }
            }
            if (cc == 0) ETR(2);
            uint8_t* slab = acquire_slab();
            if (cc == 0) ETR(3);
            tmem_wait_ld();
            if (cc == 0) ETR(4);
            if (has_aff) {
              finish_aff(va, col0, cc);
              if (two) finish_aff(vb, col0 + 32, cc + 32);
            } else if (need_fin) {
              finish32(va, col0);
              if (two) finish32(vb, col0 + 32);
            }
            if (p.dbg & 4) {
              if (va[0] == 12345u && vb[5] == 77u) slab[row] = 1;
            } else if (SLAB64) {
              uint8_t* rp = slab + row * 64;
              const int s4 = (row >> 1) & 3;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(va[g * 8 + 2 * j]),
                                                                  __uint_as_float(va[g * 8 + 2 * j + 1]));
                  w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(rp + ((g ^ s4) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            } else {
              uint8_t* rowp = slab + row * 128;
              const int sw = row & 7;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int e = (g & 3) * 8 + 2 * j;
                  const float x0 = __uint_as_float(g < 4 ? va[e] : vb[e]);
                  const float x1 = __uint_as_float(g < 4 ? va[e + 1] : vb[e + 1]);
                  const __nv_bfloat162 b2 = __floats2bfloat162_rn(x0, x1);
                  w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                const uint4 val = (g < 4 || two) ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4*>(rowp + ((g ^ sw) << 4)) = val;
              }
            }
            if (cc == 0) ETR(5);
            if (!(p.dbg & 8)) fence_proxy_async_smem();
            if (!(p.dbg & 16)) epi_bar_sync();
            if (cc == 0) ETR(6);
            if (!(BNB && c.accumulate)) issue_store(slab, col0);
            if (BNB) {
              // Fused BatchNorm-backward reductions of the producing block.  dA comes from the slab (bf16: what is
              // stored), z -- and with accumulate the old dA, which is added here in fp32 and written back to the
              // slab before it is stored -- from global memory: 16 bytes (8 columns) per thread and row, all of a
              // thread's rows in flight at once.
              const int cg = lane & (ST_CH - 1), rs = lane / ST_CH;
              const int colc = col0 + 8 * cg;
              const bool col_ok = colc < c.N;                       // N is a multiple of 8
              const bool acc = (c.accumulate != 0);
              const bool tile_full = (w0 + c.box_w <= c.valid_w) && (h0 + c.box_h <= c.valid_h) &&
                                     (i0 + c.box_n <= c.valid_n);
              const uint32_t tile_off = cls_origin(T.cls) + (uint32_t)i0 * SN + (uint32_t)h0 * SH + (uint32_t)w0 * SW + (uint32_t)colc;
              if (cc == 0) ETR(8);
              uint4 zq[ST_NI], oq[ST_NI];
              uint32_t okm = 0u;
#pragma unroll
              for (int i = 0; i < ST_NI; ++i) {
                const int r = q * 32 + ST_RS * i + rs;
                const uint2 rw = s_row[r];
                const bool ok = col_ok && (tile_full ? (r < p.rows_in_box) : row_ok(rw.y, w0, h0, i0));
                zq[i] = make_uint4(0u, 0u, 0u, 0u); oq[i] = make_uint4(0u, 0u, 0u, 0u);
                if (ok) {
                  zq[i] = *reinterpret_cast<const uint4*>(bnb_z + tile_off + rw.x);
                  if (acc) oq[i] = *reinterpret_cast<const uint4*>(bnb_o + tile_off + rw.x);
                  okm |= 1u << i;
                }
              }
              if (cc == 0) ETR(9);
              float sc[8], sh[8], s1[8], s2[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) { sc[j] = 0.f; sh[j] = 0.f; s1[j] = 0.f; s2[j] = 0.f; }
              if (col_ok) {
                const float4* ps = reinterpret_cast<const float4*>(c.bnb_scale + colc);
                const float4* ph = reinterpret_cast<const float4*>(c.bnb_shift + colc);
                const float4 a0 = ps[0], a1 = ps[1], b0 = ph[0], b1 = ph[1];
                sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
              }
#pragma unroll
              for (int i = 0; i < ST_NI; ++i) {
                if ((okm >> i) & 1u) {
                  const int r = q * 32 + ST_RS * i + rs;            // (r & 7) = (ST_RS*i + rs) & 7: q*32 drops out
                  uint8_t* qp = SLAB64 ? slab + r * 64 + ((cg ^ ((r >> 1) & 3)) << 4)
                                       : slab + r * 128 + ((cg ^ (r & 7)) << 4);
                  uint4 dq = *reinterpret_cast<const uint4*>(qp);
                  uint32_t dw[4] = {dq.x, dq.y, dq.z, dq.w};
                  const uint32_t zw[4] = {zq[i].x, zq[i].y, zq[i].z, zq[i].w};
                  if (acc) {
                    const uint32_t ow[4] = {oq[i].x, oq[i].y, oq[i].z, oq[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const float lo = __uint_as_float(dw[j] << 16) + __uint_as_float(ow[j] << 16);
                      const float hi = __uint_as_float(dw[j] & 0xFFFF0000u) + __uint_as_float(ow[j] & 0xFFFF0000u);
                      const __nv_bfloat162 b2 = __floats2bfloat162_rn(lo, hi);
                      dw[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(qp) = make_uint4(dw[0], dw[1], dw[2], dw[3]);
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float d0 = __uint_as_float(dw[j] << 16), d1 = __uint_as_float(dw[j] & 0xFFFF0000u);
                    const float z0 = __uint_as_float(zw[j] << 16), z1 = __uint_as_float(zw[j] & 0xFFFF0000u);
                    const float g0 = d0 * (fmaf(z0, sc[2 * j], sh[2 * j]) > 0.f ? 1.f : slope);
                    const float g1 = d1 * (fmaf(z1, sc[2 * j + 1], sh[2 * j + 1]) > 0.f ? 1.f : slope);
                    s1[2 * j] += g0; s2[2 * j] = fmaf(g0, z0, s2[2 * j]);
                    s1[2 * j + 1] += g1; s2[2 * j + 1] = fmaf(g1, z1, s2[2 * j + 1]);
                  }
                }
              }
              if (cc == 0) ETR(10);
              // the warp's 32 rows: sum over the row sets (lanes that share cg), in a fixed order
#pragma unroll
              for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int o = ST_CH; o < 32; o <<= 1) {
                  s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                  s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
                }
              }
              if (rs == 0 && col_ok) {
                // centre: sum g*(z - mean) = sum g*z - mean * sum g (32 rows of fp32 partial sums)
                const float4* pm = reinterpret_cast<const float4*>(c.bnb_mean + colc);
                const float4 m0 = pm[0], m1 = pm[1];
                const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                float* sg = s_stat + q * 2 * BN + cc + 8 * cg;        // one accumulator set per warp: plain adds
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  sg[j] += s1[j];
                  sg[BN + j] += fmaf(-mu[j], s1[j], s2[j]);
                }
              }
              if (cc == 0) ETR(11);
              if (acc) {                       // the slab now holds old + new: publish it to the async proxy, store
                fence_proxy_async_smem();
                epi_bar_sync();
                issue_store(slab, col0);
              }
              if (cc == 0) ETR(12);
            } else if (c.stats) {
              // fused BatchNorm statistics of the ROUNDED outputs: thread -> (column pair, row group); one
              // 32-bit shared-memory read per row brings two columns
              const int npair = SLAB64 ? 16 : 32;                 // column pairs per slab row
              const int jp = et & (npair - 1), grp = et / npair;  // 4 (8) row groups
              const int rows_per = 128 / (EPI_THREADS / npair);
              const int r0 = grp * rows_per, r1 = min(p.rows_in_box, r0 + rows_per);
              float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll 4
              for (int r = r0; r < r1; ++r) {
                const uint8_t* qp = SLAB64 ? slab + r * 64 + (((jp >> 2) ^ ((r >> 1) & 3)) << 4) + (jp & 3) * 4
                                           : slab + r * 128 + (((jp >> 2) ^ (r & 7)) << 4) + (jp & 3) * 4;
                const uint32_t wv = *reinterpret_cast<const uint32_t*>(qp);
                const float x0 = __uint_as_float(wv << 16), x1 = __uint_as_float(wv & 0xFFFF0000u);
                s1a += x0; s2a = fmaf(x0, x0, s2a);
                s1b += x1; s2b = fmaf(x1, x1, s2b);
              }
              // (column pair, row group) belongs to this thread alone: plain adds, and a summation order that
              // does not depend on warp timing (run-to-run reproducible statistics up to the f64 global adds)
              float* sg = s_stat + grp * 2 * BN + cc + 2 * jp;
              sg[0] += s1a;
              sg[1] += s1b;
              sg[BN] += s2a;
              sg[BN + 1] += s2b;
            }
            if (++slab_ctr == (uint32_t)NSLAB) slab_ctr = 0;
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < BN; cc += 32) {
            const int col0 = n0 + cc;
            if (col0 >= c.N) break;
            uint32_t va[32];
            tmem_ld32_async(tmem_acc + (uint32_t)cc, va);
            uint8_t* slab = acquire_slab();
            tmem_wait_ld();
            if (need_fin) finish32(va, col0);
            uint8_t* rowp = slab + row * 128;
            const int sw = row & 7;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(rowp + ((g ^ sw) << 4)) = make_uint4(va[4 * g], va[4 * g + 1], va[4 * g + 2], va[4 * g + 3]);
            fence_proxy_async_smem();
            epi_bar_sync();
            issue_store(slab, col0);
            if (++slab_ctr == (uint32_t)NSLAB) slab_ctr = 0;
          }
        }
        tcgen05_fence_before();
        if (lane == 0) {
          if (CTAS == 1) mbar_arrive(tempty_bar(as));
          else mbar_arrive_cluster(tempty_bar(as), 0);
        }
        ETR(7);
        if (etrace) ++etrace_n;
      } else {
        // ---- direct path: each thread stores its own row (relu_mask, unaligned or tiny outputs) ----
        const bool row_ok = (m0 + row) < c.M;
        size_t row_off = (size_t)(m0 + row) * (size_t)c.ldc + (size_t)T.z0 * c.out_bs0 + (size_t)T.z1 * c.out_bs1;
        if (c.mode == AVDN_GEMM_WGRAD) row_off += (size_t)c.taps[T.tap].bk;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
          const int col0 = n0 + cc;
          if (col0 >= c.N) break;
          uint32_t v[32];
          tmem_ld32(tmem_acc + (uint32_t)cc, v);
          if (!row_ok) continue;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]) * alpha;
            if (bias && (col0 + j) < c.N) x += __ldg(bias + col0 + j);
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
          if (is_bf16) {
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
        tcgen05_fence_before();
        if (lane == 0) {
          if (CTAS == 1) mbar_arrive(tempty_bar(as));
          else mbar_arrive_cluster(tempty_bar(as), 0);
        }
      }
    }
    if ((BNB || c.stats) && stat_nt >= 0) flush_stats();
    if (et == 0) tma_wait_group_read<0>();       // smem slabs must outlive the bulk stores reading them
#undef ETR
  }

  // ---- teardown -------------------------------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();             // the peer may still be signalling our barriers / reading our smem
  if (warp == 1) {
    tcgen05_fence_after();
    if (CTAS == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN))
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

// rank-4 tensor map with a swizzled box whose inner extent is exactly 128 bytes (128-byte swizzle) or
// 64 bytes (64-byte swizzle)
int encode_map(const void* ptr, int elem_bytes, const int64_t* dim, const int64_t* stride, const int32_t* boxdim,
               CUtensorMap* out, bool check_only = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return avdn::set_err(AVDN_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (stride[0] != 1) return avdn::set_err(AVDN_ERR_BAD_ARG, "operand dim 0 must be contiguous");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0)
    return avdn::set_err(AVDN_ERR_BAD_ARG, "operand pointer must be 16-byte aligned");
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    dims[i] = (cuuint64_t)dim[i];
    box[i] = (cuuint32_t)boxdim[i];
    if (dim[i] < 1) return avdn::set_err(AVDN_ERR_BAD_ARG, "operand dim %d < 1", i);
    if (box[i] < 1 || box[i] > 256) return avdn::set_err(AVDN_ERR_BAD_ARG, "TMA box dim %d = %u out of range", i, box[i]);
  }
  const int inner = (int)box[0] * elem_bytes;
  if (inner != 128 && inner != 64)
    return avdn::set_err(AVDN_ERR_BAD_ARG, "TMA box dim 0 must span 128 or 64 bytes (swizzled rows)");
  for (int i = 1; i < 4; ++i) {
    const cuuint64_t s = (cuuint64_t)stride[i] * (cuuint64_t)elem_bytes;
    if (s % 16 != 0 || s == 0)
      return avdn::set_err(AVDN_ERR_BAD_ARG, "operand stride %d (%llu B) must be a non-zero multiple of 16", i, (unsigned long long)s);
    strides[i - 1] = s;
  }
  if (check_only) return AVDN_OK;
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   inner == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return avdn::set_err(AVDN_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return AVDN_OK;
}

}  // namespace

int avdn::encode_tensor_map_4d(const void* ptr, int elem_bytes, const int64_t* dim, const int64_t* stride,
                               const int32_t* boxdim, CUtensorMap* out) {
  return encode_map(ptr, elem_bytes, dim, stride, boxdim, out);
}

namespace {

int encode_operand(const avdn_operand& o, CUtensorMap* out) {
  return encode_map(o.ptr, 2, o.dim, o.stride, o.box, out);
}

struct Plan {
  uint32_t magic;
  int32_t bn, a_mn, b_mn, ctas, bk;
  int32_t grid;
  int32_t smem;
  int32_t occ2;          // 1: thin tile planned for two CTAs per SM
  int32_t bnb;           // 1: fused BatchNorm-backward statistics (gemm_kernel<..., BNB = true>)
  KernelParams kp;
};
constexpr uint32_t PLAN_MAGIC = 0xA7D17C06u;

constexpr int SMEM_MAX = 232448;       // 227 KB: the dynamic shared memory a CTA can opt in to on sm_100

template <int BN, bool A_MN, bool B_MN, int CTAS, int BKT = BK, bool BNB = false>
int launch_t(const Plan& pl, cudaStream_t s) {
  auto kfn = gemm_kernel<BN, A_MN, B_MN, CTAS, BKT, BNB>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX) != cudaSuccess)
      return avdn::check_launch("cudaFuncSetAttribute(gemm_kernel)");
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl.grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = (size_t)pl.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;     // see avdn::launch_pdl (common.cuh)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = avdn::pdl_enabled() ? 2 : 1;
  if (cudaLaunchKernelEx(&cfg, kfn, pl.kp) != cudaSuccess) return avdn::check_launch("gemm_kernel launch");
  return avdn::check_launch("gemm_kernel");
}

template <int BN, int CTAS>
int launch_bn(const Plan& pl, cudaStream_t s) {
  if (!pl.a_mn && !pl.b_mn) return launch_t<BN, false, false, CTAS>(pl, s);
  if (!pl.a_mn && pl.b_mn) return launch_t<BN, false, true, CTAS>(pl, s);
  if (pl.a_mn && !pl.b_mn) return launch_t<BN, true, false, CTAS>(pl, s);
  return launch_t<BN, true, true, CTAS>(pl, s);
}

}  // namespace

extern "C" size_t avdn_gemm_plan_bytes(void) { return sizeof(Plan); }

extern "C" int avdn_gemm_plan(const avdn_gemm_desc* d, void* plan_host, size_t plan_bytes) {
  AVDN_REQUIRE(d && plan_host, "avdn_gemm_plan: null pointer");
  AVDN_REQUIRE(plan_bytes >= sizeof(Plan), "avdn_gemm_plan: plan buffer too small (%zu < %zu)", plan_bytes, sizeof(Plan));
  AVDN_REQUIRE(d->core.mode >= AVDN_GEMM_PLAIN && d->core.mode <= AVDN_GEMM_WGRAD, "avdn_gemm_plan: bad mode");
  AVDN_REQUIRE(d->bn == 32 || d->bn == 64 || d->bn == 128 || d->bn == 256, "avdn_gemm_plan: bn must be 32/64/128/256");
  AVDN_REQUIRE(d->bk == 64 || d->bk == 32, "avdn_gemm_plan: bk must be 64 or 32");
  AVDN_REQUIRE(d->bk == 64 || (!d->a_mn && !d->b_mn && d->bn == 64 && d->ctas == 1),
               "avdn_gemm_plan: 32-element k-blocks need K-major operands, bn = 64, single CTAs");
  AVDN_REQUIRE(d->bn != 32 || (!d->a_mn && !d->b_mn && d->bk == 64 && d->core.out_dtype == AVDN_DT_BF16),
               "avdn_gemm_plan: bn = 32 is the K-major bf16-output variant only");
  AVDN_REQUIRE(d->n_a >= 1 && d->n_a <= 4 && d->n_b >= 1 && d->n_b <= 4, "avdn_gemm_plan: 1..4 operand maps");
  AVDN_REQUIRE(d->core.out, "avdn_gemm_plan: null output");
  AVDN_REQUIRE(d->core.split_k >= 1 && d->core.num_kb >= 1, "avdn_gemm_plan: bad k split");
  AVDN_REQUIRE(d->core.split_k == 1 || d->core.accumulate == 2, "avdn_gemm_plan: split-K needs the atomic epilogue");
  AVDN_REQUIRE(d->core.accumulate != 2 || d->core.out_dtype == AVDN_DT_F32, "avdn_gemm_plan: atomic epilogue is fp32 only");
  AVDN_REQUIRE(d->ctas == 1 || d->ctas == 2, "avdn_gemm_plan: ctas must be 1 or 2");
  AVDN_REQUIRE(d->ctas == 1 || d->bn >= 128, "avdn_gemm_plan: cta pairs need bn >= 128");
  AVDN_REQUIRE(!d->core.stats || (d->core.out_dtype == AVDN_DT_BF16 && d->core.accumulate == 0),
               "avdn_gemm_plan: fused statistics need a plain bf16 store");
  Plan* pl = reinterpret_cast<Plan*>(plan_host);
  memset(pl, 0, sizeof(Plan));
  pl->bn = d->bn;
  pl->a_mn = d->a_mn;
  pl->b_mn = d->b_mn;
  pl->ctas = d->ctas;
  pl->bk = d->bk;
  pl->kp.c = d->core;
  const avdn_gemm_core& c = d->core;
  for (int i = 0; i < d->n_a; ++i) {
    int r = encode_operand(d->a[i], &pl->kp.tmA[i]);
    if (r) return r;
  }
  for (int i = 0; i < d->n_b; ++i) {
    avdn_operand b = d->b[i];
    if (d->ctas == 2) {            // each CTA of a pair loads half of the B tile
      if (!d->b_mn) b.box[1] = d->bn / 2;
    }
    int r = encode_operand(b, &pl->kp.tmB[i]);
    if (r) return r;
  }
  // bytes one k-step deposits in a stage (both CTAs of a pair): full boxes, OOB elements are
  // zero-filled and counted
  auto box_bytes = [](const avdn_operand& o) {
    long long b = 2;
    for (int i = 0; i < 4; ++i) b *= o.box[i];
    return b;
  };
  const long long a_bytes = box_bytes(d->a[0]) * (d->a_mn ? 2 : 1);
  const long long b_bytes = box_bytes(d->b[0]) * (d->b_mn ? d->bn / 64 : 1);
  AVDN_REQUIRE(a_bytes <= BM * d->bk * 2 && b_bytes <= (long long)d->bn * d->bk * 2, "avdn_gemm_plan: boxes exceed the stage");
  pl->kp.c.tx_bytes = (uint32_t)(a_bytes * d->ctas + b_bytes);
  pl->kp.grid_m = d->grid_m;
  pl->kp.grid_n = d->grid_n;
  pl->kp.grid_z = d->grid_z;
  AVDN_REQUIRE(d->grid_m >= 1 && d->grid_n >= 1 && d->grid_z >= 1, "avdn_gemm_plan: bad tile space %d x %d x %d",
               d->grid_m, d->grid_n, d->grid_z);
  pl->kp.rows_in_box = BM;
  {
    const char* e = getenv("AVDN_GEMM_DBG");
    pl->kp.dbg = e ? atoi(e) : 0;
    const char* b = getenv("AVDN_GEMM_DBG_BUF");      // device address of a >= 64 KB trace buffer
    pl->kp.dbg_out = b ? reinterpret_cast<long long*>(strtoull(b, nullptr, 0)) : nullptr;
  }
  if (c.mode == AVDN_GEMM_CONV) {
    pl->kp.rows_in_box = c.box_w * c.box_h * c.box_n;
    AVDN_REQUIRE(pl->kp.rows_in_box >= 1 && pl->kp.rows_in_box <= BM, "avdn_gemm_plan: conv box of %d rows", pl->kp.rows_in_box);
  }
  // ---- output tensor map: the TMA epilogue is used whenever the output can be described ----
  pl->kp.out_tma = 0;
  if (!c.relu_mask) {
    const int eb = (c.out_dtype == AVDN_DT_BF16) ? 2 : 4;
    const int slab_cols = (d->bn == 32) ? 32 : 128 / eb;
    int64_t dim[4], str[4];
    int32_t box[4];
    const uint8_t* base = reinterpret_cast<const uint8_t*>(c.out);
    bool ok = true;
    if (c.mode == AVDN_GEMM_CONV) {
      // out pixel (n, h*sh+oh, w*sw+ow), channel-contiguous rows of pitch ldc
      base += ((int64_t)c.out_oh * c.out_W + c.out_ow) * c.ldc * eb;
      dim[0] = c.N; dim[1] = c.valid_w; dim[2] = c.valid_h; dim[3] = c.valid_n;
      str[0] = 1; str[1] = (int64_t)c.out_sw * c.ldc; str[2] = (int64_t)c.out_sh * c.out_W * c.ldc;
      str[3] = (int64_t)c.out_H * c.out_W * c.ldc;
      box[0] = slab_cols; box[1] = c.box_w; box[2] = c.box_h; box[3] = c.box_n;
    } else if (c.mode == AVDN_GEMM_WGRAD) {
      dim[0] = c.ldc; dim[1] = c.M; dim[2] = 1; dim[3] = 1;
      str[0] = 1; str[1] = c.ldc; str[2] = (int64_t)c.ldc * c.M; str[3] = str[2];
      box[0] = slab_cols; box[1] = BM; box[2] = 1; box[3] = 1;
      // a tap's columns [bk, bk+N) must not spill into the next tap: N must fill whole slabs
      ok = (c.N % slab_cols) == 0;
    } else {
      dim[0] = c.N; dim[1] = c.M; dim[2] = c.batch0; dim[3] = c.batch1;
      str[0] = 1; str[1] = c.ldc;
      str[2] = c.out_bs0 ? c.out_bs0 : (int64_t)c.ldc * c.M;
      str[3] = c.out_bs1 ? c.out_bs1 : str[2] * c.batch0;
      box[0] = slab_cols; box[1] = BM; box[2] = 1; box[3] = 1;
    }
    if (ok && encode_map(base, eb, dim, str, box, nullptr, true) == AVDN_OK) {
      int r = encode_map(base, eb, dim, str, box, &pl->kp.tmC[0]);
      if (r) return r;
      pl->kp.out_tma = 1;
      if (c.mode == AVDN_GEMM_CONV && c.n_classes > 1) {       // one map per output class (parity offset)
        const uint8_t* base0 = reinterpret_cast<const uint8_t*>(c.out);
        for (int k = 0; k < c.n_classes; ++k) {
          const uint8_t* bk = base0 + ((int64_t)c.cls_oh[k] * c.out_W + c.cls_ow[k]) * c.ldc * eb;
          r = encode_map(bk, eb, dim, str, box, &pl->kp.tmC[k]);
          if (r) return r;
        }
      }
    }
  }
  AVDN_REQUIRE(!c.stats || pl->kp.out_tma, "avdn_gemm_plan: fused statistics need the TMA epilogue");
  if (c.n_classes > 1) {
    AVDN_REQUIRE(c.mode == AVDN_GEMM_CONV && c.n_classes <= 4 && pl->kp.out_tma && c.split_k == 1 && !c.stats && !c.col_scale,
                 "avdn_gemm_plan: output classes need a CONV launch with the TMA epilogue (no stats / affine / split-K)");
    AVDN_REQUIRE(c.cls_tap0[0] == 0 && c.cls_tap0[c.n_classes] == c.n_taps, "avdn_gemm_plan: class tap ranges must cover the taps");
    for (int k = 0; k < c.n_classes; ++k)
      AVDN_REQUIRE(c.cls_tap0[k + 1] > c.cls_tap0[k] && c.cls_oh[k] >= 0 && c.cls_oh[k] < c.out_sh && c.cls_ow[k] >= 0 &&
                       c.cls_ow[k] < c.out_sw,
                   "avdn_gemm_plan: bad output class %d", k);
  }
  AVDN_REQUIRE((c.col_scale == nullptr) == (c.col_shift == nullptr), "avdn_gemm_plan: col_scale and col_shift go together");
  AVDN_REQUIRE(!c.col_scale || (pl->kp.out_tma && c.out_dtype == AVDN_DT_BF16 && !c.stats && c.accumulate == 0 &&
                                !c.bias && !c.relu && c.alpha == 1.0f),
               "avdn_gemm_plan: the affine epilogue needs a plain bf16 TMA store (no stats/accumulate/bias/relu)");
  AVDN_REQUIRE(!c.residual || (c.col_scale && c.mode == AVDN_GEMM_CONV && c.out_sh == 1 && c.out_sw == 1 &&
                               c.out_oh == 0 && c.out_ow == 0 && (c.ldc % 8) == 0 && ((uintptr_t)c.residual & 15) == 0),
               "avdn_gemm_plan: residual needs the affine epilogue of a unit-stride CONV output");
  AVDN_REQUIRE(pl->kp.out_tma || c.mode != AVDN_GEMM_CONV, "avdn_gemm_plan: conv output cannot be described to TMA");
  pl->bnb = (c.bnb_z != nullptr) ? 1 : 0;
  AVDN_REQUIRE((c.bnb_z != nullptr) == (c.bnb_scale != nullptr) && (c.bnb_z != nullptr) == (c.bnb_shift != nullptr) &&
                   (c.bnb_z != nullptr) == (c.bnb_mean != nullptr) && (c.bnb_z != nullptr) == (c.bnb_sums != nullptr),
               "avdn_gemm_plan: bnb_z / bnb_scale / bnb_shift / bnb_mean / bnb_sums go together");
  AVDN_REQUIRE(!pl->bnb || (c.mode == AVDN_GEMM_CONV && !d->a_mn && !d->b_mn && pl->kp.out_tma &&
                            c.out_dtype == AVDN_DT_BF16 && !c.stats && !c.col_scale && !c.bias && !c.relu &&
                            c.alpha == 1.0f && c.accumulate <= 1 && d->bn <= 128 && (c.N % 2) == 0 &&
                            ((uintptr_t)c.bnb_z & 3) == 0 && ((uintptr_t)c.bnb_scale & 7) == 0 &&
                            ((uintptr_t)c.bnb_shift & 7) == 0 && ((uintptr_t)c.bnb_mean & 7) == 0),
               "avdn_gemm_plan: fused BatchNorm-backward statistics need a K-major CONV launch with a plain bf16 "
               "TMA-store epilogue, bn <= 128 and aligned coefficient vectors");
  AVDN_REQUIRE(!pl->bnb || (long long)c.valid_n * c.out_H * c.out_W * c.ldc < (1ll << 32),
               "avdn_gemm_plan: fused BatchNorm-backward statistics address the output with 32-bit element offsets");
  // ---- shared-memory plan: slabs, k-blocks per stage, ring depth ----
  {
    const int sub = BM * d->bk * 2 + (d->bn / d->ctas) * d->bk * 2;         // one k-block (this CTA)
    // stats + barriers + tmem slot + align slack (+ the row table of the fused BatchNorm-backward statistics)
    const int tail = (d->bn == 32 ? 8 : 4) * 2 * d->bn * 4 + (2 * MAX_STAGES + 4) * 8 + 16 + 1024 + (c.bnb_z ? 1024 : 0);
    // k-blocks per stage: one barrier round trip + commit costs the issuing thread ~300 cycles, a k-block of
    // MMAs (bn/256 * 512 cycles at bk = 64) should not be much shorter than that
    int kps = 1;
    const int mma_cycles = d->bn * 2 * d->bk / 64;                           // tensor time of one k-block
    const int budget = SMEM_MAX - tail - MAX_SLABS * SLAB_BYTES;
    while (kps < 4 && kps * mma_cycles < 1024 && (budget / ((kps + 1) * sub)) * (kps + 1) >= 4 &&
           budget / ((kps + 1) * sub) >= 2)
      ++kps;
    if (c.mode == AVDN_GEMM_CONV && c.n_taps == 9 && kps == 4) kps = 3;     // whole filter rows
    if (kps > c.num_kb) kps = c.num_kb;
    const char* ek = getenv("AVDN_GEMM_KPS");
    if (ek && atoi(ek) >= 1 && atoi(ek) <= 4) kps = atoi(ek);
    int slabs = MAX_SLABS;
    const char* es = getenv("AVDN_GEMM_SLABS");
    if (es && atoi(es) >= 2 && atoi(es) <= MAX_SLABS) slabs = atoi(es);
    // two CTAs per SM for thin tiles: each gets half of the SM's shared memory (228 KB - 1 KB reserved per CTA)
    int smem_cap = SMEM_MAX;
    const char* eo = getenv("AVDN_GEMM_OCC2");
    pl->occ2 = 0;
    const int occ_bn = eo ? atoi(eo) : 128;       // AVDN_GEMM_OCC2 = widest tile that runs two CTAs per SM (0 = off)
    if (d->bn <= occ_bn && d->bn <= 128) {
      const int cap2 = (228 * 1024) / 2 - 1024;
      int k2 = kps;
      while (k2 > 1 && (cap2 - tail - 2 * SLAB_BYTES) / (k2 * sub) < 2) --k2;      // fewer k-blocks per stage if needed
      if ((cap2 - tail - 2 * SLAB_BYTES) / (k2 * sub) >= 2) { smem_cap = cap2; slabs = 2; kps = k2; pl->occ2 = 1; }
    }
    int stages = (smem_cap - tail - slabs * SLAB_BYTES) / (kps * sub);
    while (stages < 2 && slabs > 2) { --slabs; stages = (smem_cap - tail - slabs * SLAB_BYTES) / (kps * sub); }
    while (stages < 2 && kps > 1) { --kps; stages = (smem_cap - tail - slabs * SLAB_BYTES) / (kps * sub); }
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    AVDN_REQUIRE(stages >= 2, "avdn_gemm_plan: shared memory plan failed (bn %d bk %d)", d->bn, d->bk);
    pl->kp.stages = stages;
    pl->kp.kps = kps;
    pl->kp.slabs = slabs;
    pl->smem = stages * kps * sub + slabs * SLAB_BYTES + tail;
  }
  // ---- launch geometry: persistent, one CTA (pair) per SM ----
  const long long pm = (d->grid_m + d->ctas - 1) / d->ctas;
  const long long nc = (c.mode == AVDN_GEMM_CONV && c.n_classes > 1) ? c.n_classes : 1;
  const long long tiles = pm * d->grid_n * d->grid_z * nc;
  AVDN_REQUIRE(tiles < (1ll << 31), "avdn_gemm_plan: too many tiles");
  const long long slots = (long long)avdn::sm_count() * (pl->occ2 ? 2 : 1) / d->ctas;
  long long g = tiles < slots ? tiles : slots;
  g = g / nc * nc;                     // class-fastest walk: a whole number of class groups (tiles is a multiple of nc)
  pl->grid = (int)(g * d->ctas);
  pl->magic = PLAN_MAGIC;
  return AVDN_OK;
}

extern "C" int avdn_gemm_run(const void* plan_host, avdn_stream_t stream) {
  const Plan* pl = reinterpret_cast<const Plan*>(plan_host);
  AVDN_REQUIRE(pl && pl->magic == PLAN_MAGIC, "avdn_gemm_run: not a plan");
  cudaStream_t s = avdn::to_cuda(stream);
  if (pl->kp.c.stats) {
    if (cudaMemsetAsync(pl->kp.c.stats, 0, sizeof(double) * 2 * pl->kp.c.N, s) != cudaSuccess)
      return avdn::check_launch("gemm stats memset");
  }
  if (pl->bnb) {
    if (pl->ctas == 2) {
      if (pl->bn == 128) return launch_t<128, false, false, 2, 64, true>(*pl, s);
    } else if (pl->bk == 32) {
      if (pl->bn == 64) return launch_t<64, false, false, 1, 32, true>(*pl, s);
    } else {
      switch (pl->bn) {
        case 32: return launch_t<32, false, false, 1, 64, true>(*pl, s);
        case 64: return launch_t<64, false, false, 1, 64, true>(*pl, s);
        case 128: return launch_t<128, false, false, 1, 64, true>(*pl, s);
      }
    }
    return avdn::set_err(AVDN_ERR_UNSUPPORTED, "avdn_gemm_run: bnb with bn %d ctas %d bk %d", pl->bn, pl->ctas, pl->bk);
  }
  if (pl->ctas == 2) {
    switch (pl->bn) {
      case 128: return launch_bn<128, 2>(*pl, s);
      case 256: return launch_bn<256, 2>(*pl, s);
    }
  } else {
    if (pl->bk == 32) return launch_t<64, false, false, 1, 32>(*pl, s);
    switch (pl->bn) {
      case 32: return launch_t<32, false, false, 1, 64>(*pl, s);
      case 64: return launch_bn<64, 1>(*pl, s);
      case 128: return launch_bn<128, 1>(*pl, s);
      case 256: return launch_bn<256, 1>(*pl, s);
    }
  }
  return avdn::set_err(AVDN_ERR_UNSUPPORTED, "avdn_gemm_run: bn %d ctas %d", pl->bn, pl->ctas);
}
