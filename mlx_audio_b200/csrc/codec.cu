// Codec token decode (include/b200audio.h: b2a_rvq_decode, b2a_snac_from_codes): residual-VQ codebook
// gathers summed in registers.  Index handling is exact integer work; out-of-range codes raise a flag.
#include "common.cuh"

namespace {

__global__ void rvq_decode_kernel(const int64_t* __restrict__ codes, int64_t codes_bs, int64_t codes_qs, int nq, int64_t T,
                                  const float* __restrict__ cb, int bins, int dim, float* __restrict__ out, int64_t out_ld,
                                  int* __restrict__ err, int B) {
  // one warp per (b,t): lanes stride the embedding dimension, the nq gathers accumulate in registers
  const int64_t total = (int64_t)B * T;
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < total;
       row += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    int b = (int)(row / T); int64_t t = row % T;
    for (int d0 = 0; d0 < dim; d0 += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int d = d0 + lane * 4;
      for (int q = 0; q < nq; q++) {
        int64_t code = codes[(int64_t)b * codes_bs + (int64_t)q * codes_qs + t];
        if (code < 0 || code >= bins) { if (lane == 0) atomicExch(err, 1); code = 0; }
        if (d < dim) {
          float4 e = *reinterpret_cast<const float4*>(cb + ((int64_t)q * bins + code) * dim + d);
          acc.x += e.x; acc.y += e.y; acc.z += e.z; acc.w += e.w;
        }
      }
      if (d < dim) *reinterpret_cast<float4*>(out + row * out_ld + d) = acc;
    }
  }
}

struct SnacLevels {
  const int64_t* codes[4]; const float* emb[4]; const float* w[4]; const float* bias[4]; int stride[4]; int n;
};

__global__ void snac_from_codes_kernel(SnacLevels lv, int B, int64_t T, int bins, int cd, int dim, float* __restrict__ out,
                                       int* __restrict__ err) {
  // CTA = 8 frames; thread = output channel(s).  e[level][cd] staged in shared memory.
  __shared__ float es[8][4][16];
  const int64_t t0 = (int64_t)blockIdx.x * 8; const int b = blockIdx.y;
  for (int idx = threadIdx.x; idx < 8 * lv.n * cd; idx += blockDim.x) {
    int j = idx % cd, l = (idx / cd) % lv.n, f = idx / (cd * lv.n);
    int64_t t = t0 + f;
    float v = 0.f;
    if (t < T) {
      int64_t Tl = T / lv.stride[l];
      int64_t code = lv.codes[l][(int64_t)b * Tl + t / lv.stride[l]];
      if (code < 0 || code >= bins) { atomicExch(err, 1); code = 0; }
      v = lv.emb[l][code * cd + j];
    }
    es[f][l][j] = v;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int f = 0; f < 8; f++) acc[f] = 0.f;
    for (int l = 0; l < lv.n; l++) {
      float bl = lv.bias[l] ? lv.bias[l][c] : 0.f;
#pragma unroll
      for (int f = 0; f < 8; f++) acc[f] += bl;
      for (int j = 0; j < cd; j++) {
        float wv = lv.w[l][(int64_t)j * dim + c];
#pragma unroll
        for (int f = 0; f < 8; f++) acc[f] = fmaf(wv, es[f][l][j], acc[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < 8; f++) if (t0 + f < T) out[((int64_t)b * T + t0 + f) * dim + c] = acc[f];
  }
}

}  // namespace

extern "C" int32_t b2a_rvq_decode(const int64_t* codes, int64_t codes_bs, int64_t codes_qs, int32_t B, int32_t nq, int64_t T,
                                  const float* codebooks, int32_t bins, int32_t dim, float* out, int64_t out_ld,
                                  int32_t* err_flag_dev, void* stream) {
  B2A_CHECK_ARG(codes && codebooks && out && err_flag_dev && B > 0 && nq > 0 && T > 0 && bins > 0, "bad pointers/shape");
  B2A_CHECK_ARG(dim > 0 && dim % 4 == 0 && out_ld % 4 == 0, "dim and out_ld must be multiples of 4");
  int64_t rows = (int64_t)B * T;
  int blocks = (int)((rows + 7) / 8); if (blocks > 148 * 16) blocks = 148 * 16;
  rvq_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(codes, codes_bs, codes_qs, nq, T, codebooks, bins, dim, out, out_ld,
                                                               err_flag_dev, B);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}

extern "C" int32_t b2a_snac_from_codes(const int64_t* const* codes_host_ptrs, const int32_t* strides_host, int32_t n_levels,
                                       const float* const* emb_host_ptrs, const float* const* w_host_ptrs,
                                       const float* const* bias_host_ptrs, int32_t B, int64_t T, int32_t bins, int32_t cd,
                                       int32_t dim, float* out, int32_t* err_flag_dev, void* stream) {
  B2A_CHECK_ARG(codes_host_ptrs && strides_host && emb_host_ptrs && w_host_ptrs && bias_host_ptrs && out && err_flag_dev, "null pointer");
  B2A_CHECK_ARG(n_levels > 0 && n_levels <= 4 && cd > 0 && cd <= 16 && B > 0 && T > 0 && dim > 0, "bad shape (levels<=4, codebook_dim<=16)");
  SnacLevels lv; lv.n = n_levels;
  for (int i = 0; i < n_levels; i++) {
    B2A_CHECK_ARG(strides_host[i] > 0 && T % strides_host[i] == 0, "T must be a multiple of every vq stride");
    lv.codes[i] = codes_host_ptrs[i]; lv.emb[i] = emb_host_ptrs[i]; lv.w[i] = w_host_ptrs[i]; lv.bias[i] = bias_host_ptrs[i];
    lv.stride[i] = strides_host[i];
  }
  dim3 grid(cdiv(T, 8), B);
  snac_from_codes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lv, B, T, bins, cd, dim, out, err_flag_dev);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}
