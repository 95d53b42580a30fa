// Fused Whisper decode step (include/b200audio.h: b2a_whisper_greedy_step): SuppressBlank + SuppressTokens +
// ApplyTimestampRules + GreedyDecoder.update (stt/models/whisper/decoding.py:307-325,349-442) in one launch, token
// history read on the device -- the reference builds the timestamp mask on the host from tokens.tolist() every step.
#include "common.cuh"

namespace {

struct GreedyParams {
  const float* logits; int64_t logits_bs;       // [B, V] (last position)
  const int64_t* tokens; int64_t tokens_bs;     // [B, >= cur_len] history incl. the sot sequence
  int cur_len, sample_begin, V;
  const float* suppress;                        // [V] additive mask (0 / -inf) or NULL
  const float* blank;                           // [V] additive mask for the first sampled position or NULL
  int eot, no_timestamps, timestamp_begin, max_initial_ts, without_timestamps;
  int64_t* next_out;                            // [B]
  float* sum_logprobs;                          // [B] in/out
  int* not_done;                                // [1] incremented when a row's next token is not eot
};

__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); i++) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) r += sh[i];
  __syncthreads();
  return r;
}

// one CTA per batch row
__global__ void whisper_greedy_kernel(const GreedyParams p) {
  __shared__ float sh[32];
  __shared__ int s_flags[4];                    // last_ts, pen_ts, ts_limit (exclusive upper end of the forbidden timestamp range), first
  __shared__ unsigned long long s_best;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* lg = p.logits + (int64_t)b * p.logits_bs;
  const int64_t* tk = p.tokens + (int64_t)b * p.tokens_bs;
  const int tb = p.timestamp_begin;
  const bool first = p.cur_len == p.sample_begin;
  if (tid == 0) {
    const int n = p.cur_len - p.sample_begin;
    const int64_t* seq = tk + p.sample_begin;
    const int last_ts = n >= 1 && seq[n - 1] >= tb;
    const int pen_ts = n < 2 || seq[n - 2] >= tb;
    // The reference's "timestamps must not decrease" mask is `mask[k, timestamp_begin : last_timestamp]` where
    // last_timestamp is the POSITION of the last timestamp token in the sampled sequence (decoding.py:400-408), i.e. a slice
    // [50364 : small index) that is always empty.  Reproduced as the no-op it is: limit = timestamp_begin.
    const int limit = tb;
    s_flags[0] = last_ts; s_flags[1] = pen_ts; s_flags[2] = limit; s_best = 0ull;
  }
  __syncthreads();
  const int last_ts = s_flags[0], pen_ts = s_flags[1], limit = s_flags[2];
  const float NEG = -INFINITY;
  // filtered logit BEFORE the timestamp-rule mask (what ApplyTimestampRules sees), and the rule mask without the probability test
  auto pre = [&](int v) -> float {
    float x = lg[v];
    if (first && p.blank) x += p.blank[v];
    if (p.suppress) x += p.suppress[v];
    return x;
  };
  auto rule_masked = [&](int v) -> bool {
    if (p.without_timestamps) return false;
    if (v == p.no_timestamps) return true;
    if (last_ts) { if (pen_ts) { if (v >= tb) return true; } else { if (v < p.eot) return true; } }
    if (v >= tb && v < limit) return true;
    if (first) { if (v < tb) return true; if (p.max_initial_ts >= 0 && v > tb + p.max_initial_ts) return true; }
    return false;
  };
  bool text_masked = false;
  if (!p.without_timestamps) {
    // logsumexp over timestamps vs max over text of the pre-mask logits (the normaliser cancels in the comparison)
    float mt = NEG, mx_ts = NEG;
    for (int v = tid; v < p.V; v += nt) { float x = pre(v); if (v < tb) mt = fmaxf(mt, x); else mx_ts = fmaxf(mx_ts, x); }
    mt = block_max(mt, sh); mx_ts = block_max(mx_ts, sh);
    float se = 0.f;
    if (mx_ts > NEG) for (int v = tb + tid; v < p.V; v += nt) se += expf(pre(v) - mx_ts);
    se = block_sum(se, sh);
    const float ts_lse = mx_ts > NEG ? mx_ts + logf(se) : NEG;
    text_masked = ts_lse > mt;
  }
  // final logits: argmax (lowest index on ties) and logsumexp
  float best = NEG; int besti = 0x7fffffff;
  for (int v = tid; v < p.V; v += nt) {
    float x = (rule_masked(v) || (text_masked && v < tb)) ? NEG : pre(v);
    if (x > best) { best = x; besti = v; }
  }
  const float gmax = block_max(best, sh);
  if (tid == 0) s_best = gmax > NEG ? 0xffffffffull : 0ull;        // everything masked: argmax of all -inf is index 0 (mx.argmax)
  __syncthreads();
  if (gmax > NEG && best == gmax) atomicMin(&s_best, (unsigned long long)(unsigned)besti);   // lowest index on ties
  __syncthreads();
  float se = 0.f;
  if (gmax > NEG)
    for (int v = tid; v < p.V; v += nt) {
      float x = (rule_masked(v) || (text_masked && v < tb)) ? NEG : pre(v);
      se += expf(x - gmax);
    }
  se = block_sum(se, sh);
  __syncthreads();
  if (tid == 0) {
    int nxt = (int)s_best;
    const float cur_lp = gmax > NEG ? -logf(se) : NAN;             // logit[nxt] - logsumexp = gmax - (gmax + log se)
    const bool was_eot = tk[p.cur_len - 1] == p.eot;
    if (!was_eot) p.sum_logprobs[b] += cur_lp;
    if (was_eot) nxt = p.eot;
    p.next_out[b] = nxt;
    if (nxt != p.eot) atomicAdd(p.not_done, 1);
  }
}

}  // namespace

extern "C" int32_t b2a_whisper_greedy_step(const float* logits, int64_t logits_bs, const int64_t* tokens, int64_t tokens_bs,
                                           int32_t B, int32_t cur_len, int32_t sample_begin, int32_t V, const float* suppress_mask,
                                           const float* blank_mask, int32_t eot, int32_t no_timestamps, int32_t timestamp_begin,
                                           int32_t max_initial_ts, int32_t without_timestamps, int64_t* next_out,
                                           float* sum_logprobs, int32_t* not_done, void* stream) {
  B2A_CHECK_ARG(logits && tokens && next_out && sum_logprobs && not_done, "null pointer");
  B2A_CHECK_ARG(B > 0 && V > 0 && cur_len >= sample_begin && cur_len >= 1 && timestamp_begin > 0 && timestamp_begin <= V, "bad shape");
  GreedyParams p{logits, logits_bs, tokens, tokens_bs, cur_len, sample_begin, V, suppress_mask, blank_mask, eot, no_timestamps,
                 timestamp_begin, max_initial_ts, without_timestamps, next_out, sum_logprobs, not_done};
  whisper_greedy_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(p);
  B2A_CHECK_LAUNCH();
  return B2A_OK;
}
