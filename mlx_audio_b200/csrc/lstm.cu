// Bidirectional LSTM recurrence (include/b200audio.h: b2a_lstm_bidir; reference modules.py:93-285).
// The recurrence is latency-bound: T sequential steps of a 1024x256 mat-vec.  One thread-block
// CLUSTER of 8 CTAs serves one (direction, batch) pair: CTA r owns hidden units [32r, 32r+32), i.e.
// 128 gate rows whose 128x256 slice of W_h lives entirely in REGISTERS (128 per thread, 256 threads);
// every step each CTA publishes its 32 new h values into all 8 CTAs' shared memory through
// distributed shared memory and the cluster barrier orders the steps.  W_h is read from HBM once.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int LH = 256, NCTA = 8, UPC = LH / NCTA;   // 32 units per CTA

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(256, 1)
lstm_bidir_kernel(const float* __restrict__ xproj, const float* __restrict__ wh, float* __restrict__ out,
                  int64_t out_ld, int T) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, half = tid & 1, r = tid >> 1;      // r: local gate row 0..127
  const int gate = r >> 5, unit = r & 31;
  const int grow = gate * LH + rank * UPC + unit;                   // row of W_h / column of xproj
  __shared__ __align__(16) float hbuf[2][LH];
  __shared__ float gates[4 * UPC];

  float w[128];
  {
    const float* wp = wh + ((int64_t)dir * 4 * LH + grow) * LH + half * 128;
#pragma unroll
    for (int j = 0; j < 128; j += 4) {
      float4 t = *reinterpret_cast<const float4*>(wp + j);
      w[j] = t.x; w[j + 1] = t.y; w[j + 2] = t.z; w[j + 3] = t.w;
    }
  }
  for (int i = tid; i < 2 * LH; i += 256) (&hbuf[0][0])[i] = 0.f;
  float c = 0.f;
  const float* xp_base = xproj + (int64_t)b * T * 2 * 4 * LH + (int64_t)dir * 4 * LH + grow;
  float xp_cur = 0.f;
  if (half == 0 && T > 0) xp_cur = __ldg(xp_base + (int64_t)(dir == 0 ? 0 : T - 1) * 2 * 4 * LH);
  cluster.sync();

  for (int step = 0; step < T; step++) {
    const int t = dir == 0 ? step : T - 1 - step;
    const int cur = step & 1;
    float xp_next = 0.f;                                            // prefetch next step's input projection
    if (half == 0 && step + 1 < T) xp_next = __ldg(xp_base + (int64_t)(dir == 0 ? t + 1 : t - 1) * 2 * 4 * LH);
    const float* hp = &hbuf[cur][half * 128];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 128; j += 8) {
      float4 a = *reinterpret_cast<const float4*>(hp + j);
      float4 e = *reinterpret_cast<const float4*>(hp + j + 4);
      s0 = fmaf(w[j], a.x, s0); s0 = fmaf(w[j + 1], a.y, s0); s0 = fmaf(w[j + 2], a.z, s0); s0 = fmaf(w[j + 3], a.w, s0);
      s1 = fmaf(w[j + 4], e.x, s1); s1 = fmaf(w[j + 5], e.y, s1); s1 = fmaf(w[j + 6], e.z, s1); s1 = fmaf(w[j + 7], e.w, s1);
    }
    float s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (half == 0) gates[r] = s + xp_cur;
    __syncthreads();
    if (tid < UPC) {
      const float gi = 1.f / (1.f + expf(-gates[tid]));
      const float gf = 1.f / (1.f + expf(-gates[UPC + tid]));
      const float gg = tanhf(gates[2 * UPC + tid]);
      const float go = 1.f / (1.f + expf(-gates[3 * UPC + tid]));
      c = fmaf(gf, c, gi * gg);
      const float hval = go * tanhf(c);
      out[((int64_t)b * T + t) * out_ld + dir * LH + rank * UPC + tid] = hval;
#pragma unroll
      for (int rr = 0; rr < NCTA; rr++) {
        float* remote = cluster.map_shared_rank(&hbuf[0][0], rr);
        remote[(cur ^ 1) * LH + rank * UPC + tid] = hval;
      }
    }
    xp_cur = xp_next;
    cluster.sync();
  }
}

}  // namespace

extern "C" int32_t b2a_lstm_bidir(const float* xproj, const float* wh, float* out, int64_t out_ld, int32_t B, int32_t T,
                                  int32_t H, void* stream) {
  B2A_CHECK_ARG(xproj && wh && out && B > 0 && T > 0, "bad pointers/shape");
  if (H != LH) { b2a_set_error("b2a_lstm_bidir: hidden size %d not supported (256)", H); return B2A_E_UNSUPPORTED; }
  dim3 grid(NCTA, 2, B);
  lstm_bidir_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xproj, wh, out, out_ld, T);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}
