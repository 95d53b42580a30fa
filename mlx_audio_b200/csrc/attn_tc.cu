// Flash attention on the 5th-gen tensor cores (include/b200audio.h: b2a_attention_tc) for head_dim 64:
// replaces mx.fast.scaled_dot_product_attention in the Whisper encoder / decoder prefill, Kokoro's ALBERT, Mimi and the
// Qwen3 vocoder transformer (whisper.py:369-385, modules.py:519-560, mimi/modules/transformer.py:79-112,
// speech_tokenizer.py:265-303).
//
// One CTA = 128 queries of one (batch, head); key tiles of 64.  Per key tile:
//   S = Q K^T        tcgen05.mma  M128 x N64 x K64, operands fp16 hi/lo planes (3 products: hi*hi, lo*hi, hi*lo -> fp32-grade
//                    scores), accumulator double-buffered in TMEM columns [0,64) / [128,192): S(t+1) runs under softmax(t)
//   softmax          thread r of the four softmax warps owns ROW r (tcgen05.ld hands a lane its row), so the online-softmax
//                    statistics are thread-local: no shuffles, no shared memory; P is written back to shared memory as fp16
//                    hi/lo planes in the 128-byte-swizzled K-major layout the MMA reads
//   O_tile = P V     tcgen05.mma  M128 x N64 x K64 against V^T tiles (K-major, prepared by the prologue kernel), accumulator in
//                    TMEM columns [64,128); the owning thread folds it into its 64 fp32 output registers with the rescale factor
// Q (pre-scaled by scale*log2 e so that exp2 is the only transcendental), K and V^T are converted once per call by
// attn_tc_prep_kernel; K/V tiles arrive by TMA (out-of-range keys zero-filled by the TMA unit and masked in the softmax).
// Warp roles (192 threads, 2 CTAs per SM): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = softmax.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128, BN = 64, HD = 64;

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
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// warp-convergent issue of four K=16 steps (see gemm_tc.cu: issuing from inside `if (lane == 0)` costs ~25 instructions per MMA)
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct AtcParams {
  int B, H, Tq, Tk;
  int causal, q_offset, window;
  float* o; int64_t o_bs, o_ld;          // [B, Tq, H*64] fp32
};

// ---- prologue: fp32 [B,T,H*64] -> fp16 hi/lo planes.  q,k: [B*H, T, 64];  v: transposed [B*H, 64, Tkp] (zero-padded keys)
__global__ void attn_tc_prep_qk_kernel(const float* x, int64_t x_bs, int64_t x_ld, int B, int H, int T, float mul, __half* hi, __half* lo) {
  const int64_t total = (int64_t)B * H * T * (HD / 8);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % (HD / 8));
    const int64_t r = idx / (HD / 8);
    const int t = (int)(r % T);
    const int h = (int)((r / T) % H), b = (int)(r / ((int64_t)T * H));
    const float* src = x + (int64_t)b * x_bs + (int64_t)t * x_ld + h * HD + c8 * 8;
    float4 a = *reinterpret_cast<const float4*>(src), c = *reinterpret_cast<const float4*>(src + 4);
    float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    __align__(16) __half hh[8], ll[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { float f = v[j] * mul; hh[j] = __float2half_rn(f); ll[j] = __float2half_rn(f - __half2float(hh[j])); }
    const int64_t dst = (((int64_t)b * H + h) * T + t) * HD + c8 * 8;
    *reinterpret_cast<uint4*>(hi + dst) = *reinterpret_cast<uint4*>(hh);
    *reinterpret_cast<uint4*>(lo + dst) = *reinterpret_cast<uint4*>(ll);
  }
}
// tile 64 keys x 64 dims through shared memory: out[bh][d][t]
__global__ void attn_tc_prep_vt_kernel(const float* v, int64_t v_bs, int64_t v_ld, int B, int H, int T, int Tp, __half* hi, __half* lo) {
  __shared__ float tile[64][65];
  const int t0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int d = i & 63, tt = i >> 6;
    const int t = t0 + tt;
    tile[tt][d] = t < T ? v[(int64_t)b * v_bs + (int64_t)t * v_ld + h * HD + d] : 0.f;
  }
  __syncthreads();
  const int64_t base = ((int64_t)b * H + h) * HD;
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int tt = i & 63, d = i >> 6;
    const int t = t0 + tt;
    if (t < Tp) {
      const float f = tile[tt][d];
      const __half hh = __float2half_rn(f);
      hi[(base + d) * Tp + t] = hh;
      lo[(base + d) * Tp + t] = __float2half_rn(f - __half2float(hh));
    }
  }
}

// smem (1024-aligned): Qh 16K | Ql 16K | Kh 8K | Kl 8K | Vh 8K | Vl 8K | Ph 16K | Pl 16K | barriers
constexpr int OFF_QH = 0, OFF_QL = 16384, OFF_KH = 32768, OFF_KL = 40960, OFF_VH = 49152, OFF_VL = 57344, OFF_PH = 65536, OFF_PL = 81920,
              OFF_BAR = 98304, SMEM_BYTES = OFF_BAR + 128 + 1024;

__global__ void __launch_bounds__(192, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
               const __grid_constant__ CUtensorMap map_kh, const __grid_constant__ CUtensorMap map_kl,
               const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl, const AtcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + OFF_BAR);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 2, *v_full = bars + 3, *v_empty = bars + 4, *s_full = bars + 5 /* [2] */,
           *p_full = bars + 7, *o_full = bars + 8;
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, bh = blockIdx.y;

  // key-tile range of this query tile (uniform over its rows: conservative superset, exact mask applied per element)
  int k_hi = p.Tk;                                               // exclusive
  if (p.causal) { const int last = q0 + BM - 1 + p.q_offset + 1; if (last < k_hi) k_hi = last; }
  int k_lo = 0;
  if (p.window > 0) { const int first = q0 + p.q_offset - p.window + 1; if (first > 0) k_lo = first; }
  const int t_lo = k_lo / BN, t_hi = k_hi > 0 ? (k_hi + BN - 1) / BN : 0;
  const int nt = t_hi > t_lo ? t_hi - t_lo : 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_qh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vh) : "memory");
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_full + 1, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      mbar_expect_tx(q_full, 32768);
      tma_load_3d(smem + OFF_QH, &map_qh, q_full, 0, q0, bh);
      tma_load_3d(smem + OFF_QL, &map_ql, q_full, 0, q0, bh);
      for (int t = 0; t < nt; t++) {
        const int kt = (t_lo + t) * BN;
        mbar_wait(k_empty, (t & 1) ^ 1);
        mbar_expect_tx(k_full, 16384);
        tma_load_3d(smem + OFF_KH, &map_kh, k_full, 0, kt, bh);
        tma_load_3d(smem + OFF_KL, &map_kl, k_full, 0, kt, bh);
        mbar_wait(v_empty, (t & 1) ^ 1);
        mbar_expect_tx(v_full, 16384);
        tma_load_3d(smem + OFF_VH, &map_vh, v_full, kt, 0, bh);
        tma_load_3d(smem + OFF_VL, &map_vl, v_full, kt, 0, bh);
      }
    }
  } else if (warp == 1) {
    // D = f32, A = B = f16 (format 0), K-major both, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    // S is double-buffered in TMEM (columns [0,64) and [128,192)): S(t+1) = Q K(t+1)^T is issued BEFORE waiting for P(t), so the
    // score MMA of the next key tile runs underneath the softmax of the current one.
    static_assert(HD / 16 == 4 && BN / 16 == 4, "umma_x4 issues four K steps");
    const uint64_t dqh = umma_desc_sw128(sb + OFF_QH), dql = umma_desc_sw128(sb + OFF_QL), dkh = umma_desc_sw128(sb + OFF_KH),
                   dkl = umma_desc_sw128(sb + OFF_KL), dph = umma_desc_sw128(sb + OFF_PH), dpl = umma_desc_sw128(sb + OFF_PL),
                   dvh = umma_desc_sw128(sb + OFF_VH), dvl = umma_desc_sw128(sb + OFF_VL);
    auto issue_s = [&](int t) {
      mbar_wait(k_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d = tmem_base + ((t & 1) ? 128u : 0u);
      umma_x4(d, dqh, dkh, idesc, 0u);                         // S = Qh Kh^T + Ql Kh^T + Qh Kl^T
      umma_x4(d, dql, dkh, idesc, 1u);
      umma_x4(d, dqh, dkl, idesc, 1u);
      umma_commit_elect(k_empty);
      umma_commit_elect(s_full + (t & 1));
    };
    if (nt > 0) { mbar_wait(q_full, 0); issue_s(0); }
    for (int t = 0; t < nt; t++) {
      if (t + 1 < nt) issue_s(t + 1);
      mbar_wait(p_full, t & 1);
      mbar_wait(v_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // O_tile = Ph Vh + Pl Vh + Ph Vl   (A = P [128 x 64 keys], B = V^T tile [64 dims x 64 keys])
      umma_x4(tmem_base + 64, dph, dvh, idesc, 0u);
      umma_x4(tmem_base + 64, dpl, dvh, idesc, 1u);
      umma_x4(tmem_base + 64, dph, dvl, idesc, 1u);
      umma_commit_elect(v_empty);
      umma_commit_elect(o_full);
    }
  } else {
    // ===== softmax / accumulate: thread <-> query row =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                        // row inside the tile == TMEM lane
    const int qi = q0 + row;                                    // query index
    const int qpos = qi + p.q_offset;                           // its position on the key axis (causal / window)
    const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_o = t_s + 64;
    float acc[HD];
#pragma unroll
    for (int j = 0; j < HD; j++) acc[j] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow_h = smem + OFF_PH + row * 128;
    uint8_t* prow_l = smem + OFF_PL + row * 128;
    const int sw = row & 7;
    for (int t = 0; t < nt; t++) {
      const int kt = (t_lo + t) * BN;
      mbar_wait(s_full + (t & 1), (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float s[BN];
      const uint32_t t_sb = t_s + ((t & 1) ? 128u : 0u);
      tmem_ld32_nowait(t_sb, s);
      tmem_ld32_nowait(t_sb + 32, s + 32);
      tmem_wait_ld();
      // mask + row max
      int kmax = p.Tk - 1;                                      // last allowed key
      if (p.causal && qpos < kmax) kmax = qpos;
      const int kmin = p.window > 0 ? qpos - p.window + 1 : 0;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < BN; j++) {
        const int kk = kt + j;
        if (kk > kmax || kk < kmin) s[j] = -INFINITY;
        mx = fmaxf(mx, s[j]);
      }
      const float m_new = fmaxf(m, mx);
      const float alpha = (m_new == -INFINITY) ? 1.f : exp2f(m - m_new);     // m = -inf -> 0 (acc is 0 anyway)
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < BN / 8; c++) {
        __align__(16) __half hh[8], ll[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float pv = (m_new == -INFINITY) ? 0.f : exp2f(s[c * 8 + j] - m_new);
          rs += pv;
          hh[j] = __float2half_rn(pv);
          ll[j] = __float2half_rn(pv - __half2float(hh[j]));
        }
        const int pc = (c ^ sw) << 4;                            // 128-byte swizzle: 16-byte chunk index XOR (row mod 8)
        *reinterpret_cast<uint4*>(prow_h + pc) = *reinterpret_cast<uint4*>(hh);
        *reinterpret_cast<uint4*>(prow_l + pc) = *reinterpret_cast<uint4*>(ll);
      }
      l = l * alpha + rs;
      m = m_new;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async-proxy reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // fold the previous accumulator scale while the P V MMA runs
#pragma unroll
      for (int j = 0; j < HD; j++) acc[j] *= alpha;
      mbar_wait(o_full, t & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_ld32_nowait(t_o, s);
      tmem_ld32_nowait(t_o + 32, s + 32);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < HD; j++) acc[j] += s[j];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    // ---- write out: stage the 128 x 64 fp32 tile through shared memory (P planes are free now) for coalesced rows
    const float inv = l > 0.f ? 1.f / l : 0.f;
    float* stage = reinterpret_cast<float*>(smem + OFF_VH);      // 128 rows x 68 floats = 34 816 B over the idle V / P buffers
#pragma unroll
    for (int j = 0; j < HD; j += 4)
      *reinterpret_cast<float4*>(stage + row * 68 + j) = make_float4(acc[j] * inv, acc[j + 1] * inv, acc[j + 2] * inv, acc[j + 3] * inv);
    __syncwarp();
    const int b = bh / p.H, h = bh % p.H;
    // each warp wrote its own 32 rows: it streams them out two rows per instruction (16 lanes x float4 = 256 B per row)
    const int half = lane >> 4, l16 = lane & 15;
    for (int rr = 0; rr < 32; rr += 2) {
      const int r = quarter * 32 + rr + half;
      const int q = q0 + r;
      if (q < p.Tq) {
        const float4 v4 = *reinterpret_cast<const float4*>(stage + r * 68 + l16 * 4);
        *reinterpret_cast<float4*>(p.o + (int64_t)b * p.o_bs + (int64_t)q * p.o_ld + h * HD + l16 * 4) = v4;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;

int map3(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1) {
  if (!g_enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    g_enc = (EncodeTiledFn)fn;
  }
  cuuint64_t gd[3] = {d0, d1, d2}; cuuint64_t gs[2] = {s1, s2}; cuuint32_t bx[3] = {b0, b1, 1}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

inline int64_t tk_pad(int Tk) { return ((int64_t)Tk + 7) / 8 * 8; }

}  // namespace

extern "C" int64_t b2a_attention_tc_ws_bytes(int32_t B, int32_t H, int32_t Tq, int32_t Tk) {
  const int64_t bh = (int64_t)B * H;
  return 2 * 2 * (bh * Tq * HD + bh * Tk * HD + bh * HD * tk_pad(Tk)) + 1024;
}

extern "C" int32_t b2a_attention_tc(const b2a_attn_t* a, void* ws, void* stream) {
  B2A_CHECK_ARG(a && ws && a->q && a->k && a->v && a->o, "null pointer");
  B2A_CHECK_ARG(a->D == 64 && a->H == a->Hkv && a->k_len == nullptr, "tensor-core attention: head_dim 64, no GQA, no per-row key length");
  B2A_CHECK_ARG(a->B > 0 && a->Tq > 0 && a->Tk > 0 && a->H > 0, "bad shape");
  B2A_CHECK_ARG(a->q_ld % 4 == 0 && a->k_ld % 4 == 0 && a->o_ld % 4 == 0 && a->q_bs % 4 == 0 && a->k_bs % 4 == 0 && a->o_bs % 4 == 0 &&
                ((uintptr_t)a->q & 15) == 0 && ((uintptr_t)a->k & 15) == 0 && ((uintptr_t)a->o & 15) == 0, "q/k/o rows must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = a->B, H = a->H, Tq = a->Tq, Tk = a->Tk;
  const int64_t bh = (int64_t)B * H, Tkp = tk_pad(Tk);
  __half* base = (__half*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  __half* qh = base; __half* ql = qh + bh * Tq * HD;
  __half* kh = ql + bh * Tq * HD; __half* kl = kh + bh * Tk * HD;
  __half* vh = kl + bh * Tk * HD; __half* vl = vh + bh * HD * Tkp;
  {
    int64_t tq = bh * Tq * (HD / 8), tk = bh * Tk * (HD / 8);
    int gq = (int)((tq + 255) / 256); if (gq > 148 * 16) gq = 148 * 16;
    int gk = (int)((tk + 255) / 256); if (gk > 148 * 16) gk = 148 * 16;
    attn_tc_prep_qk_kernel<<<gq, 256, 0, st>>>(a->q, a->q_bs, a->q_ld, B, H, Tq, a->scale * 1.4426950408889634f, qh, ql);
    attn_tc_prep_qk_kernel<<<gk, 256, 0, st>>>(a->k, a->k_bs, a->k_ld, B, H, Tk, 1.f, kh, kl);
    dim3 gv((unsigned)((Tkp + 63) / 64), H, B);
    attn_tc_prep_vt_kernel<<<gv, 256, 0, st>>>(a->v, a->v_bs, a->v_ld, B, H, Tk, (int)Tkp, vh, vl);
  }
  CUtensorMap mqh, mql, mkh, mkl, mvh, mvl;
  int e = map3(&mqh, qh, HD, Tq, bh, HD * 2, (uint64_t)Tq * HD * 2, HD, BM);
  if (!e) e = map3(&mql, ql, HD, Tq, bh, HD * 2, (uint64_t)Tq * HD * 2, HD, BM);
  if (!e) e = map3(&mkh, kh, HD, Tk, bh, HD * 2, (uint64_t)Tk * HD * 2, HD, BN);
  if (!e) e = map3(&mkl, kl, HD, Tk, bh, HD * 2, (uint64_t)Tk * HD * 2, HD, BN);
  if (!e) e = map3(&mvh, vh, Tkp, HD, bh, Tkp * 2, (uint64_t)HD * Tkp * 2, BN, HD);
  if (!e) e = map3(&mvl, vl, Tkp, HD, bh, Tkp * 2, (uint64_t)HD * Tkp * 2, BN, HD);
  if (e) { b2a_set_error("b2a_attention_tc: cuTensorMapEncodeTiled failed (%d)", e); return B2A_E_CUDA; }
  AtcParams p{B, H, Tq, Tk, a->causal, a->q_offset, a->window, a->o, a->o_bs, a->o_ld};
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); attr = true; }
  dim3 grid((Tq + BM - 1) / BM, (unsigned)bh);
  attn_tc_kernel<<<grid, 192, SMEM_BYTES, st>>>(mqh, mql, mkh, mkl, mvh, mvl, p);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}
