// Tensor-core path for the genuinely dense layers (include/b200audio.h: b2a_prep_bf16, b2a_conv1d_tc).
//
// A stride-1 1-D convolution over channels-last activations is a sum of row-shifted GEMMs:
//     Y[l, n] = sum_tap sum_ci  A[l + shift_tap, ci] * W[tap][n][ci]
// so one tcgen05 kernel serves every dense conv (and every Linear: one tap, shift 0).  TMA fetches the
// shifted A tile for each tap straight from the activation matrix -- rows outside [0, L) are zero-filled by
// the TMA unit, which IS the convolution's zero padding -- and the weight tile, both into 128B-swizzled
// shared memory that tcgen05.mma consumes through shared-memory descriptors; the fp32 accumulator lives in
// TMEM and is drained by four epilogue warps (tcgen05.ld) that fuse bias / activation / LayerScale /
// residual / scale / accumulate.
//
// Precision: activations are stored as TWO bf16 planes (hi = bf16(a), lo = bf16(a - hi)) written by the
// prologue kernel together with the AdaIN/Snake/LeakyReLU input transform; weights are bf16-exact
// (bf16 checkpoint), so  (hi + lo) * w  reproduces the fp32 product to ~2^-17 while running on the bf16
// tensor pipe (2 MMAs per tile instead of 1).  `planes = 1` drops the lo plane (pure bf16 activations).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (warp_id % 4 selects the TMEM lane quarter).  One 128 x BN output tile per CTA.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int TM = 128;            // rows (positions) per CTA tile == UMMA_M
constexpr int TK = 64;             // bf16 elements per 128-byte swizzle row == K extent of one stage

__host__ __device__ __forceinline__ uint32_t pow2_cols(uint32_t n) { uint32_t c = 32; while (c < n) c <<= 1; return c; }   // TMEM allocations are powers of two >= 32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, 128B-swizzled operand tile whose rows are 128 bytes: 8-row groups 1024 B apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);              // start address
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// start address shifted by whole 128-byte rows inside a swizzled tile: bo_mode 1 also records the row phase in the descriptor's
// base-offset field (bits 49-51 = (addr >> 7) & 7)
__device__ __forceinline__ uint64_t umma_desc_sw128_rows(uint32_t saddr, int bo_mode) {
  uint64_t d = umma_desc_sw128(saddr);
  if (bo_mode) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  return d;
}
// Warp-convergent issue (all 32 lanes call with identical operands; one elected lane issues) of the four K=16 steps of a 64-wide K chunk.
// Issuing from inside `if (lane == 0)` makes ptxas guard every UTCHMMA operand with an ELECT + R2UR.BROADCAST sequence; with the ring-slot
// divisions and per-MMA descriptor rebuilds one stage (8 MMAs = 0.27 us of tensor time) cost ~500 instructions = ~1 us of issue time
// (round-2 measurement on conv_fused.cu, whose issuer was a copy of this one).
__device__ __forceinline__ void umma_x4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred e, p, q;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "add.s64 a1, %1, 2;\n\tadd.s64 b1, %2, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 b2, %2, 4;\n\tadd.s64 a3, %1, 6;\n\tadd.s64 b3, %2, 6;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, q;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, q;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, q;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
               ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __noinline__ float act_noinline(float v, int act, float p0) { return b2a_act(v, act, p0, 1.f, 1.f); }

struct TcParams {
  int B, L, Lout, Cout, cin_pad, taps, planes, wplanes, BN, stages, f16;
  int Mrows, up_s, up_crop, C;     // transposed-conv mode: N = up_s * C, GEMM row m & column (r, co) -> output row m*up_s + r - up_crop
  int reuse, R, shift_min, wst, bo_mode;   // A-reuse mode: one (128 + span)-row A tile per K chunk serves every tap (row-shifted descriptors)
  int shift[32];
  const float* bias; int post_act; float post_p0;
  const float* cscale; int64_t cscale_bs;
  const float* res; int64_t res_bs, res_ld; int res_div;
  float out_scale; int accumulate;
  float* y; int64_t y_bs, y_ld;
  long long* dbg;                 // optional: 8 clock64 stamps from CTA (0,0,0) (b2a_conv1d_tc_debug)
  double* stats; int stats_slots; // optional InstanceNorm partials of the OUTPUT: [B][stats_slots][C][2] = (sum, sum of squares) per 32-row group
};

// smem: [stages] x { A_hi 16 KB | A_lo 16 KB (planes==2) | W BN*128 B }, then barriers
__global__ void __launch_bounds__(192, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = TM * 128, w_bytes = p.BN * 128;
  const int stage_bytes = a_bytes * p.planes + w_bytes * p.wplanes;
  uint64_t* full = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* tmem_full = empty + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();
  const int l0 = blockIdx.x * TM, n0 = blockIdx.y * p.BN, b = blockIdx.z;
  const int kchunks = p.cin_pad / TK;
  const int iters = p.taps * kchunks;

  uint32_t tmem_base = 0;
  if (!p.reuse) {
    if (warp == 0 && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
      for (int s = 0; s < p.stages; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      mbar_init(tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {                                    // TMEM: BN fp32 accumulator columns (power of two >= 32)
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(pow2_cols((uint32_t)p.BN)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();
  }

  if (p.reuse) {
    // ================= A-reuse mode =================
    // smem: [planes] x A super-tile (R rows x 128 B) | [wst] x { W BN*128 B (x wplanes) } | barriers.  For each 64-channel K chunk the
    // (128 + span)-row activation tile is fetched ONCE; tap t multiplies rows [shift_t - shift_min, +128) of it, addressed by
    // advancing the shared-memory descriptor by whole 128-byte rows, so the L2 -> SM operand traffic per chunk drops from
    // taps * (A + W) to A' + taps * W.
    const int a_rbytes = p.R * 128;
    const int a_total = a_rbytes * p.planes;
    const int w_stage = w_bytes * p.wplanes;
    uint8_t* wbase = smem + a_total;
    uint64_t* bars = (uint64_t*)(wbase + (size_t)p.wst * w_stage);
    uint64_t* a_full = bars; uint64_t* a_empty = bars + 1;
    uint64_t* w_full = bars + 2; uint64_t* w_empty = w_full + p.wst;
    uint64_t* r_tmem_full = w_empty + p.wst;
    uint32_t* r_slot = (uint32_t*)(r_tmem_full + 1);
    if (warp == 0 && lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
      mbar_init(a_full, 1); mbar_init(a_empty, 1);
      for (int s = 0; s < p.wst; s++) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
      mbar_init(r_tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(r_slot)), "r"(pow2_cols((uint32_t)p.BN)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *r_slot;
    if (dbg && threadIdx.x == 0) p.dbg[1] = clock64();
    if (warp == 0) {
      if (lane == 0) {                                  // W producer
        for (int it = 0; it < iters; it++) {
          const int kc = it / p.taps, tap = it - kc * p.taps;
          const int s = it % p.wst, ph = (it / p.wst) & 1;
          mbar_wait(w_empty + s, ph ^ 1);
          uint8_t* st = wbase + (size_t)s * w_stage;
          mbar_expect_tx(w_full + s, (uint32_t)w_stage);
          tma_load_2d(st, &map_w, w_full + s, kc * TK, tap * p.Cout + n0);
          if (p.wplanes == 2) tma_load_2d(st + w_bytes, &map_wlo, w_full + s, kc * TK, tap * p.Cout + n0);
        }
      }
    } else if (warp == 2) {
      if (lane == 0) {                                  // A producer: its own warp, so it never shares a scheduler slot with the W ring
        for (int kc = 0; kc < kchunks; kc++) {
          mbar_wait(a_empty, (kc & 1) ^ 1);
          mbar_expect_tx(a_full, (uint32_t)a_total);
          tma_load_3d(smem, &map_hi, a_full, kc * TK, l0 + p.shift_min, b);
          if (p.planes == 2) tma_load_3d(smem + a_rbytes, &map_lo, a_full, kc * TK, l0 + p.shift_min, b);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      const uint32_t fmt = p.f16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
      const uint32_t a0 = smem_u32(smem);
      uint32_t s = 0, ph = 0, accum = 0;
      const bool two_a = p.planes == 2, two_w = p.wplanes == 2;
      for (int kc = 0; kc < kchunks; kc++) {
        mbar_wait(a_full, kc & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int tap = 0; tap < p.taps; tap++) {
          const int it = kc * p.taps + tap;
          mbar_wait(w_full + s, ph);
          if (dbg && lane == 0 && it == 0) p.dbg[2] = clock64();
          if (dbg && lane == 0 && it == iters - 1) p.dbg[3] = clock64();
          const uint32_t wst_addr = smem_u32(wbase + (size_t)s * w_stage);
          const uint32_t rowoff = (uint32_t)(p.shift[tap] - p.shift_min) * 128u;
          const uint64_t wd = umma_desc_sw128(wst_addr), ad = umma_desc_sw128_rows(a0 + rowoff, p.bo_mode);
          umma_x4(tbase, ad, wd, idesc, accum);
          accum = 1;
          if (two_a) umma_x4(tbase, umma_desc_sw128_rows(a0 + a_rbytes + rowoff, p.bo_mode), wd, idesc, 1u);
          if (two_w) umma_x4(tbase, ad, umma_desc_sw128(wst_addr + w_bytes), idesc, 1u);
          umma_commit_elect(w_empty + s);
          if (tap == p.taps - 1) umma_commit_elect(a_empty);
          if (it == iters - 1) umma_commit_elect(r_tmem_full);
          if (++s == (uint32_t)p.wst) { s = 0; ph ^= 1; }
        }
      }
    }
    // epilogue warps fall through to the shared epilogue below with these aliases
    tmem_full = r_tmem_full;
    tmem_base = tbase;
  } else
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < iters; it++) {
        const int s = it % p.stages, ph = (it / p.stages) & 1;
        mbar_wait(empty + s, ph ^ 1);
        const int tap = it / kchunks, kc = it % kchunks;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(full + s, (uint32_t)stage_bytes);
        tma_load_3d(st, &map_hi, full + s, kc * TK, l0 + p.shift[tap], b);
        if (p.planes == 2) tma_load_3d(st + a_bytes, &map_lo, full + s, kc * TK, l0 + p.shift[tap], b);
        tma_load_2d(st + (size_t)a_bytes * p.planes, &map_w, full + s, kc * TK, tap * p.Cout + n0);
        if (p.wplanes == 2) tma_load_2d(st + (size_t)a_bytes * p.planes + w_bytes, &map_wlo, full + s, kc * TK, tap * p.Cout + n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both, N>>3 @17, M>>4 @24
    const uint32_t fmt = p.f16 ? 0u : 1u;                 // F16F32Format: 0 = f16, 1 = bf16
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    uint32_t s = 0, ph = 0;
    const bool two_a = p.planes == 2, two_w = p.wplanes == 2;
    for (int it = 0; it < iters; it++) {
      mbar_wait(full + s, ph);
      if (dbg && lane == 0 && it == 0) p.dbg[2] = clock64();
      if (dbg && lane == 0 && it == iters - 1) p.dbg[3] = clock64();
      const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
      // products kept: a_hi*w_hi, a_lo*w_hi, and (fp32 checkpoints: weights split too) a_hi*w_lo; a_lo*w_lo ~ 2^-17*2^-9 is dropped
      const uint64_t ad = umma_desc_sw128(st), wd = umma_desc_sw128(st + a_bytes * p.planes);
      umma_x4(tmem_base, ad, wd, idesc, it != 0);
      if (two_a) umma_x4(tmem_base, umma_desc_sw128(st + a_bytes), wd, idesc, 1u);
      if (two_w) umma_x4(tmem_base, ad, umma_desc_sw128(st + a_bytes * p.planes + w_bytes), idesc, 1u);
      umma_commit_elect(empty + s);                        // frees the stage when these MMAs retire
      if (it == iters - 1) umma_commit_elect(tmem_full);   // accumulator complete
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
    }
  }
  if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> (per-warp shared-memory transpose) -> coalesced global =====
    // tcgen05.ld hands lane i the 32 columns of ROW i; writing that straight out makes every global access touch 32
    // different rows (32 sectors per instruction, latencies exposed one after another: measured 9k cycles per 32-column
    // chunk, 5x the main loop).  Each warp therefore transposes its 32x32 chunk through a padded smem tile and then
    // reads/writes whole 128-byte row segments: residual and previous-output loads for all 32 rows are issued first.
    const int quarter = warp & 3;
    // the transpose staging tile reuses pipeline stage 0: every TMA write and MMA read of it has retired once tmem_full fires
    float* stage = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
    // while the main loop runs, pull this warp's residual / previous-output rows towards L2 (lane = row, one request per 128 B)
    if (!p.up_s) {
      const int prow = l0 + quarter * 32 + lane;
      if (prow < p.Lout) {
        if (p.res) {
          const float* q = p.res + (int64_t)b * p.res_bs + (int64_t)(p.res_div == 2 ? (prow >> 1) : prow) * p.res_ld + n0;
          for (int c = 0; c < p.BN; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + c));
        }
        if (p.accumulate) {
          const float* q = p.y + (int64_t)b * p.y_bs + (int64_t)prow * p.y_ld + n0;
          for (int c = 0; c < p.BN; c += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + c));
        }
      }
    }
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (dbg && warp == 2 && lane == 0) p.dbg[4] = clock64();
    const int mrow0 = l0 + quarter * 32;                           // first GEMM row of this warp's quarter
    const int mul = p.up_s ? p.up_s : 1;
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
      if (dbg && warp == 2 && lane == 0 && c0 == 0) p.dbg[8] = clock64();
      if (mrow0 >= p.Mrows) continue;
#pragma unroll
      for (int j = 0; j < 32; j++) stage[lane * 33 + j] = __uint_as_float(r[j]);
      __syncwarp();
      if (dbg && warp == 2 && lane == 0 && c0 == 0) p.dbg[9] = clock64();
      const int n = n0 + c0 + lane;                               // GEMM column of this lane for every row below
      const int ph = p.up_s ? n / p.C : 0;                         // up-sampling phase (uniform over the 32-column chunk)
      const int co = n - ph * p.C;                                 // output channel
      const int add = p.up_s ? ph - p.up_crop : 0;
      const float bias = p.bias ? __ldg(p.bias + co) : 0.f;
      const float cs = p.cscale ? __ldg(p.cscale + (int64_t)b * p.cscale_bs + co) : 1.f;
      float* ycol = p.y + (int64_t)b * p.y_bs + co;
      const float* rcol = p.res ? p.res + (int64_t)b * p.res_bs + co : nullptr;
      // The four epilogue warps are the only warps on their schedulers, so this code is issue-latency bound: per-row work is
      // kept to one load / one FMA chain / one store off incrementally advanced pointers.  Valid rows form a contiguous range
      // [i_lo, i_hi) of the 32 (row = row0 + i*mul is monotonic); full tiles take the predicate-free path.
      const int row0 = mrow0 * mul + add;
      const int mvalid = min(32, p.Mrows - mrow0);
      int i_lo = 0, i_hi = mvalid;
      if (row0 < 0) i_lo = (-row0 + mul - 1) / mul;
      if (row0 + (mvalid - 1) * mul >= p.Lout) i_hi = p.Lout > row0 ? (p.Lout - row0 + mul - 1) / mul : 0;
      const int64_t ystride = (int64_t)mul * p.y_ld;
      float* yp = ycol + (int64_t)row0 * p.y_ld;
      const bool half_res = p.res_div == 2;                        // nearest x2 shortcut (istftnet.py:838-850); only with mul == 1
      const int odd = row0 & 1;
      const float* rp = rcol ? rcol + (int64_t)(half_res ? (row0 >> 1) : row0) * p.res_ld : nullptr;
      const int64_t rstride = (int64_t)mul * p.res_ld;
      const float osc = p.out_scale;
      if (dbg && warp == 2 && lane == 0 && c0 == 0) p.dbg[10] = clock64();
      if (i_lo == 0 && i_hi == 32) {
        float rr[32];
        if (rp) {
          if (!half_res) {
#pragma unroll
            for (int i = 0; i < 32; i++) rr[i] = __ldg(rp + i * rstride);
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) rr[i] = __ldg(rp + (int64_t)((i + odd) >> 1) * p.res_ld);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) rr[i] = 0.f;
        }
        if (p.accumulate) {
          float oo[32];
#pragma unroll
          for (int i = 0; i < 32; i++) oo[i] = yp[i * ystride];
#pragma unroll
          for (int i = 0; i < 32; i++) rr[i] = rr[i] * osc + oo[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) rr[i] *= osc;
        }
        const float cso = cs * osc;
        if (p.post_act) {
#pragma unroll
          for (int i = 0; i < 32; i++) yp[i * ystride] = act_noinline(stage[i * 33 + lane] + bias, p.post_act, p.post_p0) * cso + rr[i];
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) yp[i * ystride] = (stage[i * 33 + lane] + bias) * cso + rr[i];
        }
      } else {
        for (int i = i_lo; i < i_hi; i++) {                        // ragged edge tiles: plain loop
          const int row = row0 + i * mul;
          float t = stage[i * 33 + lane] + bias;
          if (p.post_act) t = act_noinline(t, p.post_act, p.post_p0);
          float r = rcol ? __ldg(rcol + (int64_t)(half_res ? (row >> 1) : row) * p.res_ld) : 0.f;
          float o = p.accumulate ? ycol[(int64_t)row * p.y_ld] : 0.f;
          ycol[(int64_t)row * p.y_ld] = (t * cs + r) * osc + o;
        }
      }
      __syncwarp();
      if (dbg && warp == 2 && lane == 0 && c0 == 0) p.dbg[11] = clock64();
    }
  }
  if (dbg && warp == 2 && lane == 0) p.dbg[5] = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dbg && threadIdx.x == 0) p.dbg[6] = clock64();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(pow2_cols((uint32_t)p.BN)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks the output tiles (static round-robin), the accumulator is DOUBLE-BUFFERED in TMEM
// (2 x BN columns) and eight epilogue warps (two per TMEM lane quarter, splitting the column chunks) drain tile i while the
// TMA / MMA warps already run tile i+1.  Against the one-tile-per-CTA kernel this removes (a) the wave-quantisation tail
// (366 tiles on 296 CTA slots = 2 waves), (b) the per-tile prologue (barrier init, TMEM alloc, descriptor prefetch) and
// (c) the serialisation of main loop and epilogue inside a CTA; with one CTA per SM the operand ring is as deep as shared
// memory allows.  smem: [stages] x { A planes | W planes } | 8 x 32x33 fp32 transpose tiles | barriers.
constexpr int PERSIST_THREADS = 352;     // warp 0 TMA (W / classic), warp 1 MMA, warps 2..9 epilogue, warp 10 A-tile producer (reuse mode)
constexpr int PERSIST_STAGING = 8 * 32 * 33 * 4;

__global__ void __launch_bounds__(PERSIST_THREADS, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                       const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_wlo, const TcParams p,
                       int ntm, int ntn, int ntiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = TM * 128, w_bytes = p.BN * 128;
  const int stage_bytes = a_bytes * p.planes + w_bytes * p.wplanes;
  // classic: [stages] x {A planes | W planes};  reuse: [2] x A super-tile planes (R rows) | [wst] x W planes.  Then staging, barriers.
  const int a_rbytes = p.R * 128, a_buf = a_rbytes * p.planes, w_stage = w_bytes * p.wplanes;
  const size_t ring_bytes = p.reuse ? (size_t)2 * a_buf + (size_t)p.wst * w_stage : (size_t)p.stages * stage_bytes;
  uint8_t* wbase = smem + (size_t)2 * a_buf;
  const int nring = p.reuse ? p.wst : p.stages;
  float* staging = reinterpret_cast<float*>(smem + ring_bytes);
  uint64_t* full = (uint64_t*)(smem + ring_bytes + PERSIST_STAGING);
  uint64_t* empty = full + nring;
  uint64_t* tfull = empty + nring;             // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint64_t* a_full = tempty + 2;               // [2] (reuse mode)
  uint64_t* a_empty = a_full + 2;              // [2]
  uint32_t* tmem_slot = (uint32_t*)(a_empty + 2);
  const int kchunks = p.cin_pad / TK;
  const int iters = p.taps * kchunks;
  const uint32_t tbuf_stride = pow2_cols((uint32_t)p.BN);      // second accumulator starts on a power-of-two column
  const uint32_t tmem_cols = 2 * tbuf_stride;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
    for (int s = 0; s < nring; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(tfull, 1); mbar_init(tfull + 1, 1); mbar_init(tempty, 8); mbar_init(tempty + 1, 8);
    mbar_init(a_full, 1); mbar_init(a_full + 1, 1); mbar_init(a_empty, 1); mbar_init(a_empty + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && p.reuse) {
    // ===== W producer (reuse mode): one weight tile per (K chunk, tap) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n0 = (tile % ntn) * p.BN;
        for (int i = 0; i < iters; i++, it++) {
          const int kc = i / p.taps, tap = i - kc * p.taps;
          const int s = it % p.wst, ph = (it / p.wst) & 1;
          mbar_wait(empty + s, ph ^ 1);
          uint8_t* st = wbase + (size_t)s * w_stage;
          mbar_expect_tx(full + s, (uint32_t)w_stage);
          tma_load_2d(st, &map_w, full + s, kc * TK, tap * p.Cout + n0);
          if (p.wplanes == 2) tma_load_2d(st + w_bytes, &map_wlo, full + s, kc * TK, tap * p.Cout + n0);
        }
      }
    }
  } else if (warp == 10) {
    // ===== A producer (reuse mode): one (128 + span)-row activation tile per K chunk, double-buffered =====
    if (lane == 0 && p.reuse) {
      uint32_t cg = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = (tile / ntn) % ntm, b = tile / (ntn * ntm);
        const int l0 = mt * TM;
        for (int kc = 0; kc < kchunks; kc++, cg++) {
          const uint32_t ab = cg & 1;
          mbar_wait(a_empty + ab, ((cg >> 1) & 1) ^ 1);
          uint8_t* dst = smem + (size_t)ab * a_buf;
          mbar_expect_tx(a_full + ab, (uint32_t)a_buf);
          tma_load_3d(dst, &map_hi, a_full + ab, kc * TK, l0 + p.shift_min, b);
          if (p.planes == 2) tma_load_3d(dst + a_rbytes, &map_lo, a_full + ab, kc * TK, l0 + p.shift_min, b);
        }
      }
    }
  } else if (warp == 0) {
    // ===== TMA producer: runs ahead across tile boundaries, bounded only by the operand ring =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int nt = tile % ntn, mt = (tile / ntn) % ntm, b = tile / (ntn * ntm);
        const int l0 = mt * TM, n0 = nt * p.BN;
        for (int i = 0; i < iters; i++, it++) {
          const int s = it % p.stages, ph = (it / p.stages) & 1;
          mbar_wait(empty + s, ph ^ 1);
          const int tap = i / kchunks, kc = i % kchunks;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(full + s, (uint32_t)stage_bytes);
          tma_load_3d(st, &map_hi, full + s, kc * TK, l0 + p.shift[tap], b);
          if (p.planes == 2) tma_load_3d(st + a_bytes, &map_lo, full + s, kc * TK, l0 + p.shift[tap], b);
          tma_load_2d(st + (size_t)a_bytes * p.planes, &map_w, full + s, kc * TK, tap * p.Cout + n0);
          if (p.wplanes == 2) tma_load_2d(st + (size_t)a_bytes * p.planes + w_bytes, &map_wlo, full + s, kc * TK, tap * p.Cout + n0);
        }
      }
    }
  } else if (warp == 1 && p.reuse) {
    // ===== MMA issuer (reuse mode): tap t reads rows [shift_t - shift_min, +128) of the resident A tile =====
    const uint32_t fmt = p.f16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    const uint32_t a0 = smem_u32(smem);
    uint32_t rs = 0, rph = 0, lt = 0, cg = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      mbar_wait(tempty + buf, (use & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + buf * tbuf_stride;
      uint32_t accum = 0;
      for (int kc = 0; kc < kchunks; kc++, cg++) {
        const uint32_t ab = cg & 1;
        mbar_wait(a_full + ab, (cg >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int tap = 0; tap < p.taps; tap++) {
          mbar_wait(full + rs, rph);
          const uint32_t wst_addr = smem_u32(wbase + (size_t)rs * w_stage);
          const uint32_t abase = a0 + ab * (uint32_t)a_buf + (uint32_t)(p.shift[tap] - p.shift_min) * 128u;
          const uint64_t wd = umma_desc_sw128(wst_addr), ad = umma_desc_sw128(abase);
          umma_x4(tacc, ad, wd, idesc, accum);
          accum = 1;
          if (p.planes == 2) umma_x4(tacc, umma_desc_sw128(abase + a_rbytes), wd, idesc, 1u);
          if (p.wplanes == 2) umma_x4(tacc, ad, umma_desc_sw128(wst_addr + w_bytes), idesc, 1u);
          umma_commit_elect(empty + rs);
          if (tap == p.taps - 1) umma_commit_elect(a_empty + ab);
          if (kc == kchunks - 1 && tap == p.taps - 1) umma_commit_elect(tfull + buf);
          if (++rs == (uint32_t)p.wst) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t fmt = p.f16 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
    uint32_t rs = 0, rph = 0, lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
      const uint32_t buf = lt & 1, use = lt >> 1;
      mbar_wait(tempty + buf, (use & 1) ^ 1);                 // the epilogue has drained this accumulator (first use: free)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + buf * tbuf_stride;
      for (int i = 0; i < iters; i++) {
        mbar_wait(full + rs, rph);
        const uint32_t st = smem_u32(smem + (size_t)rs * stage_bytes);
        const uint64_t ad = umma_desc_sw128(st), wd = umma_desc_sw128(st + a_bytes * p.planes);
        umma_x4(tacc, ad, wd, idesc, i != 0);
        if (p.planes == 2) umma_x4(tacc, umma_desc_sw128(st + a_bytes), wd, idesc, 1u);
        if (p.wplanes == 2) umma_x4(tacc, ad, umma_desc_sw128(st + a_bytes * p.planes + w_bytes), idesc, 1u);
        umma_commit_elect(empty + rs);
        if (i == iters - 1) umma_commit_elect(tfull + buf);
        if (++rs == (uint32_t)p.stages) { rs = 0; rph ^= 1; }
      }
    }
  } else if (warp >= 2 && warp <= 9) {
    // ===== epilogue (8 warps): quarter = warp % 4 selects the TMEM lanes, sub = (warp - 2) / 4 the 32-column chunks it owns =====
    const int quarter = warp & 3, sub = (warp - 2) >> 2;
    float* stage = staging + (warp - 2) * (32 * 33);
    const int mul = p.up_s ? p.up_s : 1;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, lt++) {
      const int nt = tile % ntn, mt = (tile / ntn) % ntm, b = tile / (ntn * ntm);
      const int l0 = mt * TM, n0 = nt * p.BN;
      const uint32_t buf = lt & 1, use = lt >> 1;
      if (!p.up_s) {                                          // pull residual / previous-output rows towards L2 while the MMAs run
        const int prow = l0 + quarter * 32 + lane;
        if (prow < p.Lout) {
          if (p.res) {
            const float* q = p.res + (int64_t)b * p.res_bs + (int64_t)(p.res_div == 2 ? (prow >> 1) : prow) * p.res_ld + n0;
            for (int c = sub * 32; c < p.BN; c += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + c));
          }
          if (p.accumulate) {
            const float* q = p.y + (int64_t)b * p.y_bs + (int64_t)prow * p.y_ld + n0;
            for (int c = sub * 32; c < p.BN; c += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + c));
          }
        }
      }
      mbar_wait(tfull + buf, use & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int mrow0 = l0 + quarter * 32;
      for (int c0 = sub * 32; c0 < p.BN; c0 += 64) {
        uint32_t r[32];
        tmem_ld32(tmem_base + buf * tbuf_stride + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, r);
        if (mrow0 >= p.Mrows) {                                // rows past the end: their statistics slot is an explicit zero
          if (p.stats) {
            const int n_ = n0 + c0 + lane, ph_ = p.up_s ? n_ / p.C : 0;
            double* w = p.stats + ((((int64_t)b * p.stats_slots + (int64_t)(mt * 4 + quarter) * mul + ph_) * p.C) + (n_ - ph_ * p.C)) * 2;
            w[0] = 0.0; w[1] = 0.0;
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j++) stage[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        const int n = n0 + c0 + lane;
        const int ph = p.up_s ? n / p.C : 0;
        const int co = n - ph * p.C;
        const int add = p.up_s ? ph - p.up_crop : 0;
        const float bias = p.bias ? __ldg(p.bias + co) : 0.f;
        const float cs = p.cscale ? __ldg(p.cscale + (int64_t)b * p.cscale_bs + co) : 1.f;
        float* ycol = p.y + (int64_t)b * p.y_bs + co;
        const float* rcol = p.res ? p.res + (int64_t)b * p.res_bs + co : nullptr;
        const int row0 = mrow0 * mul + add;
        const int mvalid = min(32, p.Mrows - mrow0);
        int i_lo = 0, i_hi = mvalid;
        if (row0 < 0) i_lo = (-row0 + mul - 1) / mul;
        if (row0 + (mvalid - 1) * mul >= p.Lout) i_hi = p.Lout > row0 ? (p.Lout - row0 + mul - 1) / mul : 0;
        const int64_t ystride = (int64_t)mul * p.y_ld;
        float* yp = ycol + (int64_t)row0 * p.y_ld;
        const bool half_res = p.res_div == 2;
        const int odd = row0 & 1;
        const float* rp = rcol ? rcol + (int64_t)(half_res ? (row0 >> 1) : row0) * p.res_ld : nullptr;
        const int64_t rstride = (int64_t)mul * p.res_ld;
        const float osc = p.out_scale;
        float st1 = 0.f, st2 = 0.f;                            // InstanceNorm partials of the values written below (this lane's column)
        if (i_lo == 0 && i_hi == 32) {
          float rr[32];
          if (rp) {
            if (!half_res) {
#pragma unroll
              for (int i = 0; i < 32; i++) rr[i] = __ldg(rp + i * rstride);
            } else {
#pragma unroll
              for (int i = 0; i < 32; i++) rr[i] = __ldg(rp + (int64_t)((i + odd) >> 1) * p.res_ld);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) rr[i] = 0.f;
          }
          if (p.accumulate) {
            float oo[32];
#pragma unroll
            for (int i = 0; i < 32; i++) oo[i] = yp[i * ystride];
#pragma unroll
            for (int i = 0; i < 32; i++) rr[i] = rr[i] * osc + oo[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) rr[i] *= osc;
          }
          const float cso = cs * osc;
          if (p.post_act) {
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float v = act_noinline(stage[i * 33 + lane] + bias, p.post_act, p.post_p0) * cso + rr[i];
              yp[i * ystride] = v; st1 += v; st2 = fmaf(v, v, st2);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float v = (stage[i * 33 + lane] + bias) * cso + rr[i];
              yp[i * ystride] = v; st1 += v; st2 = fmaf(v, v, st2);
            }
          }
        } else {
          for (int i = i_lo; i < i_hi; i++) {
            const int row = row0 + i * mul;
            float t = stage[i * 33 + lane] + bias;
            if (p.post_act) t = act_noinline(t, p.post_act, p.post_p0);
            float rv = rcol ? __ldg(rcol + (int64_t)(half_res ? (row >> 1) : row) * p.res_ld) : 0.f;
            float o = p.accumulate ? ycol[(int64_t)row * p.y_ld] : 0.f;
            const float v = (t * cs + rv) * osc + o;
            ycol[(int64_t)row * p.y_ld] = v; st1 += v; st2 = fmaf(v, v, st2);
          }
        }
        if (p.stats) {
          double* w = p.stats + ((((int64_t)b * p.stats_slots + (int64_t)(mt * 4 + quarter) * mul + ph) * p.C) + co) * 2;
          w[0] = (double)st1; w[1] = (double)st2;
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + buf);                // 8 arrivals free the accumulator for tile lt + 2
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- prologue: fp32 activations -> (hi, lo) bf16 planes with the fused input transform; pad channels are zeroed
template <typename T> __device__ __forceinline__ T to16(float v);
template <> __device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }
__device__ __forceinline__ float from16(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float from16(__half v) { return __half2float(v); }

// 8 channels per thread: 2 x 16-byte reads, one 16-byte write per plane
// ACT >= 0: the activation is a compile-time constant (no per-element switch, a third of the code: the generic instantiation was
// instruction-cache- and branch-bound, see profiles/r01_mimi_launches.md); ACT = -1: runtime `act`.
template <typename T16, int ACT>
__global__ void prep_bf16_kernel(const float* __restrict__ x, int64_t x_bs, int64_t x_ld, int B, int L, int C, int cpad,
                                 const float* __restrict__ scale, const float* __restrict__ shift, int act, float p0,
                                 const float* __restrict__ a, const float* __restrict__ bb, T16* __restrict__ hi,
                                 T16* __restrict__ lo) {
  const int cp8 = cpad / 8;
  const int64_t total = (int64_t)B * L * cp8;
  const bool vec = (x_ld % 4 == 0) && (x_bs % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  // Measured alternatives that were SLOWER than this plain grid-stride loop (kept out): a channel group pinned per thread with the
  // constants in registers and no divisions (Whisper prep 5.0 -> 7.7 ms), and two index slots per iteration with both loads in
  // flight (vocoder prep 24 -> 30 ms: the extra registers cost more occupancy than the second load buys).
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cp8) * 8;
    const int64_t r = idx / cp8;
    const int l = (int)(r % L), b = (int)(r / L);
    const float* xp = x + (int64_t)b * x_bs + (int64_t)l * x_ld + c;
    float v[8];
    if (vec && c + 8 <= C) {
      float4 t0 = __ldg(reinterpret_cast<const float4*>(xp)), t1 = __ldg(reinterpret_cast<const float4*>(xp) + 1);
      v[0] = t0.x; v[1] = t0.y; v[2] = t0.z; v[3] = t0.w; v[4] = t1.x; v[5] = t1.y; v[6] = t1.z; v[7] = t1.w;
    } else {
#pragma unroll
      for (int q = 0; q < 8; q++) v[q] = (c + q < C) ? __ldg(xp + q) : 0.f;
    }
    __align__(16) T16 h[8];
    __align__(16) T16 lw[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int cc = c + q;
      float t = 0.f;
      if (cc < C) {
        t = v[q];
        if (scale) t = fmaf(t, __ldg(scale + (int64_t)b * C + cc), __ldg(shift + (int64_t)b * C + cc));
        if constexpr (ACT == B2A_ACT_SNAKE) { const float sn = b2a_sin(__ldg(a + cc) * t); t = fmaf(__ldg(bb + cc), sn * sn, t); }
        else if constexpr (ACT == B2A_ACT_ELU) t = t > 0.f ? t : expm1f(t);
        else if constexpr (ACT == B2A_ACT_LRELU) t = t > 0.f ? t : t * p0;
        else if constexpr (ACT == 0) { }
        else if (act) t = b2a_act(t, act, p0, a ? __ldg(a + cc) : 1.f, bb ? __ldg(bb + cc) : 1.f);
      }
      h[q] = to16<T16>(t);
      lw[q] = to16<T16>(t - from16(h[q]));
    }
    *reinterpret_cast<uint4*>(hi + r * cpad + c) = *reinterpret_cast<uint4*>(h);
    if (lo) *reinterpret_cast<uint4*>(lo + r * cpad + c) = *reinterpret_cast<uint4*>(lw);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
long long* g_dbg = nullptr;

int get_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int f16) {
  cuuint64_t gd[3]; cuuint64_t gs[2]; cuuint32_t bx[3]; cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i < rank - 1; i++) gs[i] = strides_bytes[i];
  CUresult r = g_encode(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace

/* debug aid: device buffer of 8 int64 that the next b2a_conv1d_tc launches stamp with clock64() at their phase boundaries */
extern "C" int32_t b2a_conv1d_tc_debug(void* dbg8) { g_dbg = (long long*)dbg8; return B2A_OK; }

extern "C" int32_t b2a_prep_bf16(const float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t L, int32_t C, int32_t cpad,
                                 const float* scale, const float* shift, int32_t act, float p0, const float* a, const float* b,
                                 void* hi, void* lo, int32_t f16, void* stream) {
  B2A_CHECK_ARG(x && hi && B > 0 && L > 0 && C > 0 && cpad >= C && cpad % 64 == 0, "bad pointers/shape (cpad must be a multiple of 64)");
  B2A_CHECK_ARG((scale == nullptr) == (shift == nullptr), "scale and shift come together");
  int64_t total = (int64_t)B * L * (cpad / 8);
  int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
#define B2A_PREP_LAUNCH(T, A) prep_bf16_kernel<T, A><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, x_bs, x_ld, B, L, C, cpad, scale, shift, act, p0, a, b, (T*)hi, (T*)lo)
  if (f16) {
    if (act == 0) B2A_PREP_LAUNCH(__half, 0); else B2A_PREP_LAUNCH(__half, -1);
  } else if (act == 0) B2A_PREP_LAUNCH(__nv_bfloat16, 0);
  else if (act == B2A_ACT_SNAKE && a && b) B2A_PREP_LAUNCH(__nv_bfloat16, B2A_ACT_SNAKE);
  else if (act == B2A_ACT_ELU) B2A_PREP_LAUNCH(__nv_bfloat16, B2A_ACT_ELU);
  else if (act == B2A_ACT_LRELU) B2A_PREP_LAUNCH(__nv_bfloat16, B2A_ACT_LRELU);
  else B2A_PREP_LAUNCH(__nv_bfloat16, -1);
#undef B2A_PREP_LAUNCH
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}

extern "C" int32_t b2a_conv1d_tc(const void* a_hi, const void* a_lo, int32_t f16, int32_t B, int32_t L, int32_t cin_pad, const void* w_bf16,
                                 const void* w_lo,
                                 int32_t taps, const int32_t* shifts_host, int32_t Cout, int32_t Lout, const float* bias,
                                 int32_t post_act, float post_p0, const float* cscale, int64_t cscale_bs, const float* res,
                                 int64_t res_bs, int64_t res_ld, int32_t res_div, float out_scale, int32_t accumulate, float* y,
                                 int64_t y_bs, int64_t y_ld, int32_t up_stride, int32_t up_crop, double* stats_ws, int32_t stats_slots,
                                 void* stream) {
  B2A_CHECK_ARG(a_hi && w_bf16 && y && shifts_host, "null pointer");
  B2A_CHECK_ARG(up_stride >= 0 && up_crop >= 0 && (up_stride == 0 || (Cout % up_stride == 0 && (Cout / up_stride) % 32 == 0)),
                "transposed mode: Cout = up_stride * C with C a multiple of 32");
  B2A_CHECK_ARG(B > 0 && L > 0 && Lout > 0 && taps > 0 && taps <= 32 && cin_pad % 64 == 0 && (res_div == 1 || res_div == 2), "bad shape (res_div must be 1 or 2)");
  B2A_CHECK_ARG(Cout % 32 == 0 && y_ld % 4 == 0 && (res == nullptr || res_ld % 4 == 0), "Cout must be a multiple of 32; row strides multiples of 4");
  if (get_encode() != 0) { b2a_set_error("b2a_conv1d_tc: cuTensorMapEncodeTiled entry point not found"); return B2A_E_CUDA; }
  TcParams p;
  p.f16 = f16 ? 1 : 0;
  p.wplanes = w_lo ? 2 : 1;
  p.up_s = up_stride; p.up_crop = up_crop; p.C = up_stride ? Cout / up_stride : Cout;
  p.Mrows = up_stride ? L + taps - 1 : Lout;
  p.B = B; p.L = L; p.Lout = Lout; p.Cout = Cout; p.cin_pad = cin_pad; p.taps = taps; p.planes = a_lo ? 2 : 1;
  // N tile: 256 halves the A re-reads per output but only pays once the grid fills the machine; otherwise 128 (more CTAs)
  const int mtiles = cdiv(up_stride ? L + taps - 1 : Lout, TM) * B;
  // N tile = a divisor of Cout, multiple of 32 (the epilogue's chunk), <= 256 (UMMA N): the largest one >= 128 that still gives every
  // SM a tile; when the problem is too small for that, the widest one <= 128 (more CTAs; narrower tiles measured slower on Kokoro).  96 / 192 matter: the Qwen3 vocoder's
  // 96-, 192- and 384-channel blocks would otherwise run as 32-/64-/128-wide tiles and re-read A three times.
  {
    int best = 0, fallback = 0;
    for (int bn = 256; bn >= 32; bn -= 32) {
      if (Cout % bn) continue;
      if (!best && bn >= 128 && (int64_t)mtiles * (Cout / bn) >= 148) best = bn;   // wide tile that still gives every SM work
      if (!fallback && bn <= 128) fallback = bn;                                    // else the widest tile <= 128 (more CTAs)
    }
    p.BN = best ? best : fallback;
  }
  for (int i = 0; i < taps; i++) p.shift[i] = shifts_host[i];
  p.bias = bias; p.post_act = post_act; p.post_p0 = post_p0; p.cscale = cscale; p.cscale_bs = cscale_bs;
  p.res = res; p.res_bs = res_bs; p.res_ld = res_ld; p.res_div = res_div; p.out_scale = out_scale; p.accumulate = accumulate;
  p.y = y; p.y_bs = y_bs; p.y_ld = y_ld;
  p.dbg = g_dbg;
  p.stats = stats_ws; p.stats_slots = stats_slots;
  const int stage_bytes = TM * 128 * p.planes + p.BN * 128 * p.wplanes;
  // Two CTAs per SM whenever two 2-stage pipelines fit (<= ~113 KB each): one CTA's epilogue then overlaps the other's main
  // loop.  Otherwise one CTA per SM with as many stages as fit.  (Stage 0 doubles as the 17 KB epilogue staging tile.)
  // Grids that cannot even give every SM one CTA gain nothing from co-residency: they get the deep pipeline instead (the serial
  // K loop of a decoder conv -- 54 iterations on 32 CTAs -- is then bound by L2 bandwidth, not by TMA latency).
  const int per_cta_2 = 2 * stage_bytes + 1024 + 128;
  static int deep_max = -1;
  if (deep_max < 0) { const char* e = getenv("B2A_TC_DEEP_MAX"); deep_max = e ? atoi(e) : 80; }
  const int total_ctas = cdiv(p.Mrows, TM) * (Cout / p.BN) * B;
  if (total_ctas > deep_max && 2 * per_cta_2 <= 226 * 1024 && 2 * stage_bytes >= 4 * 32 * 33 * 4) p.stages = 2;
  else { p.stages = (212 * 1024) / stage_bytes; if (p.stages > 6) p.stages = 6; if (p.stages < 1) p.stages = 1; }
  if ((size_t)p.stages * stage_bytes < 4 * 32 * 33 * 4) { b2a_set_error("b2a_conv1d_tc: tile too small for the epilogue staging"); return B2A_E_UNSUPPORTED; }
  size_t smem = (size_t)p.stages * stage_bytes + 1024 /*align slack*/ + (2 * p.stages + 1) * 8 + 64;
  // A-reuse mode (multi-tap stride-1 convs whose taps span <= 64 rows): see the kernel.  B2A_TC_REUSE=0 disables it,
  // B2A_TC_BO=1 selects the descriptor base-offset variant.
  static int reuse_on = -1, bo_mode = -1;
  if (reuse_on < 0) { const char* e = getenv("B2A_TC_REUSE"); reuse_on = (e && e[0] == '1') ? 1 : 0; }   // measured slower (single A buffer): opt-in
  if (bo_mode < 0) { const char* e = getenv("B2A_TC_BO"); bo_mode = (e && e[0] == '1') ? 1 : 0; }
  p.reuse = 0; p.R = TM; p.shift_min = 0; p.wst = 0; p.bo_mode = bo_mode;
  uint32_t a_box_rows = TM;
  if (reuse_on && !up_stride && taps >= 2) {
    int smin = p.shift[0], smax = p.shift[0];
    for (int i = 1; i < taps; i++) { smin = p.shift[i] < smin ? p.shift[i] : smin; smax = p.shift[i] > smax ? p.shift[i] : smax; }
    const int span = smax - smin;
    if (span <= 64) {
      const int R = (TM + span + 7) / 8 * 8;
      const int a_total = R * 128 * p.planes, w_stage = p.BN * 128 * p.wplanes;
      int wst = 3;
      size_t need = (size_t)a_total + (size_t)wst * w_stage;
      if (2 * (need + 1024 + 256) > 226 * 1024) {                       // one CTA per SM anyway: deepen the weight ring
        while (wst < 6 && (size_t)a_total + (size_t)(wst + 1) * w_stage + 2048 <= 200 * 1024) wst++;
        need = (size_t)a_total + (size_t)wst * w_stage;
      }
      if (need + 2048 <= 220 * 1024 && a_total >= 4 * 32 * 33 * 4) {
        p.reuse = 1; p.R = R; p.shift_min = smin; p.wst = wst;
        a_box_rows = (uint32_t)R;
        smem = need + 1024 + (2 * wst + 3) * 8 + 64;
      }
    }
  }

  // ---- persistent kernel (default): stage count / A-reuse layout that fit one CTA per SM
  static int persist_on = -1, preuse_on = -1, nsm = 0;
  if (persist_on < 0) {
    const char* e = getenv("B2A_TC_PERSIST"); persist_on = (e && e[0] == '0') ? 0 : 1;
    const char* r = getenv("B2A_TC_PREUSE"); preuse_on = (r && r[0] == '1') ? 1 : 0;   // A-reuse: opt-in (measured 1-2 % slower: the loop is not L2-bound)
    int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm <= 0) nsm = 148;
  }
  bool use_persist = false;
  TcParams pq = p;
  size_t psm = 0;
  if (persist_on && !p.reuse && !p.dbg && 2 * p.BN <= 512) {
    const size_t budget = (size_t)227 * 1024 - PERSIST_STAGING - 2048;
    if (preuse_on && !up_stride && taps >= 2) {
      int smin = p.shift[0], smax = p.shift[0];
      for (int i = 1; i < taps; i++) { smin = p.shift[i] < smin ? p.shift[i] : smin; smax = p.shift[i] > smax ? p.shift[i] : smax; }
      const int span = smax - smin;
      const int R = (TM + span + 7) / 8 * 8;
      const size_t a2 = (size_t)2 * R * 128 * p.planes, w_stage = (size_t)p.BN * 128 * p.wplanes;
      if (span <= 64 && a2 + 2 * w_stage <= budget) {
        int wst = (int)((budget - a2) / w_stage);
        if (wst > 8) wst = 8;
        pq.reuse = 1; pq.R = R; pq.shift_min = smin; pq.wst = wst;
        a_box_rows = (uint32_t)R;
        psm = a2 + (size_t)wst * w_stage + PERSIST_STAGING + 1024 + (2 * wst + 8) * 8 + 64;
        use_persist = true;
      }
    }
    if (!use_persist) {
      int pst = (int)(budget / stage_bytes);
      if (pst > 6) pst = 6;
      if (pst >= 2) {
        pq.stages = pst; pq.reuse = 0; pq.R = TM; pq.wst = 0;
        psm = (size_t)pst * stage_bytes + PERSIST_STAGING + 1024 + (2 * pst + 8) * 8 + 64;
        use_persist = true;
      }
    }
  }

  if (stats_ws) {
    const int need = cdiv(p.Mrows, TM) * 4 * (up_stride ? up_stride : 1);
    if (!use_persist || stats_slots != need) {
      b2a_set_error("b2a_conv1d_tc: fused output statistics need the persistent kernel and %d slots (got %d)", need, stats_slots);
      return B2A_E_UNSUPPORTED;
    }
  }

  CUtensorMap mh, ml, mw, mwl;
  uint64_t adims[3] = {(uint64_t)cin_pad, (uint64_t)L, (uint64_t)B};
  uint64_t astr[2] = {(uint64_t)cin_pad * 2, (uint64_t)cin_pad * 2 * (uint64_t)L};
  uint32_t abox[3] = {TK, a_box_rows, 1};
  int e = make_map(&mh, a_hi, 3, adims, astr, abox, p.f16);
  if (!e) e = make_map(&ml, a_lo ? a_lo : a_hi, 3, adims, astr, abox, p.f16);
  uint64_t wdims[2] = {(uint64_t)cin_pad, (uint64_t)taps * Cout};
  uint64_t wstr[1] = {(uint64_t)cin_pad * 2};
  uint32_t wbox[2] = {TK, (uint32_t)p.BN};
  if (!e) e = make_map(&mw, w_bf16, 2, wdims, wstr, wbox, p.f16);
  if (!e) e = make_map(&mwl, w_lo ? w_lo : w_bf16, 2, wdims, wstr, wbox, p.f16);
  if (e) { b2a_set_error("b2a_conv1d_tc: cuTensorMapEncodeTiled failed (%d)", e); return B2A_E_CUDA; }

  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(conv_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  if (use_persist) {
    const int ntm = cdiv(p.Mrows, TM), ntn = Cout / p.BN, ntiles = ntm * ntn * B;
    const int g = ntiles < nsm ? ntiles : nsm;
    conv_tc_persist_kernel<<<g, PERSIST_THREADS, psm, (cudaStream_t)stream>>>(mh, ml, mw, mwl, pq, ntm, ntn, ntiles);
    B2A_CHECK_LAUNCH();
    return B2A_OK;
  }
  dim3 grid(cdiv(p.Mrows, TM), Cout / p.BN, B);
  conv_tc_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(mh, ml, mw, mwl, p);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}
