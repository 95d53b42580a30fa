/* b200audio -- C ABI of the B200-native speech-inference hot path.
 *
 * The reference (Blaizzy/mlx-audio) is pure Python on Apple MLX: it has no FFI for this path.
 * Each entry point below replaces the MLX primitive call sites listed beside it (paths relative
 * to /root/reference/mlx_audio); INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions (SURVEY.md section 8b "C-ABI layer"):
 *   - every function returns 0 on success or a negative B2A_E_* code; b2a_last_error() gives a
 *     thread-local message; nothing is ever allocated -- the caller owns every buffer, including
 *     the workspaces, whose sizes the *_ws_bytes helpers report;
 *   - all pointers are DEVICE pointers unless the name ends in _host; activations are float32,
 *     channels-last [B, L, C] with explicit batch / row strides in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; kernels are asynchronous on it.
 */
#ifndef B200AUDIO_H
#define B200AUDIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2A_OK 0
#define B2A_E_INVALID (-1)   /* bad argument (ValueError on the Python side) */
#define B2A_E_CUDA (-2)      /* CUDA runtime error (RuntimeError) */
#define B2A_E_UNSUPPORTED (-3)

/* activation codes used by prologues / epilogues */
enum {
  B2A_ACT_NONE = 0,
  B2A_ACT_LRELU = 1,      /* p0 = negative slope                (istftnet.py:721, nn.LeakyReLU) */
  B2A_ACT_SNAKE = 2,      /* x + b[c]*sin(a[c]*x)^2             (istftnet.py:382, snac/layers.py:124, speech_tokenizer.py:123) */
  B2A_ACT_ELU = 3,        /* alpha = 1                          (mimi/modules/seanet.py:102) */
  B2A_ACT_GELU = 4,       /* exact erf                          (whisper.py:415, modules.py:536) */
  B2A_ACT_GELU_TANH = 5,  /* tanh approximation                 (mimi/modules/transformer.py:141) */
  B2A_ACT_TANH = 6,
  B2A_ACT_SIGMOID = 7,
  B2A_ACT_SILU = 8,
  B2A_ACT_CLIP1 = 9       /* clip to [-1, 1]                    (speech_tokenizer.py:880) */
};

const char* b2a_last_error(void);
int32_t b2a_version(void);
int32_t b2a_device_sm_count(void);

/* ---- 1-D convolution family --------------------------------------------------------------
 * Replaces mx.conv1d / mx.conv_transpose1d / nn.Linear call sites:
 *   tts/models/kokoro/istftnet.py:128-166,811,915; codec/models/mimi/modules/conv.py:41-48,103-109;
 *   codec/models/snac/layers.py:58,111-118; stt/models/whisper/whisper.py:430-431; every nn.Linear
 *   (a Linear is the K=1 case with L = rows).
 * y[b,l,co] = epilogue( bias[co] + sum_{k,ci} W[k][ci][co] * pre(x[b, l*stride - pad_left + k*dilation, ci]) )
 *   pre(v)   = act_pre( v*pre_scale[b,ci] + pre_shift[b,ci] ), and 0 outside [0,L) (pad_mode 0) or
 *              the clamped edge sample (pad_mode 1)
 *   epilogue(v) = ( act_post(v) * post_cscale[b?,co] + res[b, l/res_div, co] ) * out_scale   (+= y if accumulate)
 * Transposed form (scatter rule of mx.conv_transpose1d, no kernel flip), with q = l + pad_left:
 *   y[b,l,co] = epilogue( bias[co] + sum_{ci} sum_{k: (q-k)%stride==0} W[k][ci][co] * pre(x[b,(q-k)/stride,ci]) )
 * Weights are pre-packed by the host as float32 [K][Cin/groups][Cout] (bf16-exact values).
 * groups must be 1 (dense) or == Cin == Cout (depthwise, W is [K][C]).
 */
typedef struct {
  const float* x; int64_t x_bs; int64_t x_ld;
  int32_t B, L, Cin;
  const float* w; const float* bias;
  float* y; int64_t y_bs; int64_t y_ld;
  int32_t Lout, Cout;
  int32_t K, stride, dilation, pad_left, groups, pad_mode;
  const float* pre_scale; const float* pre_shift;      /* [B,Cin] or NULL */
  int32_t pre_act; float pre_p0; const float* pre_a; const float* pre_b;   /* per-Cin snake params */
  int32_t post_act; float post_p0;
  const float* post_cscale; int64_t post_cscale_bs;    /* [Cout] (bs 0) or [B,Cout] or NULL */
  const float* res; int64_t res_bs; int64_t res_ld; int32_t res_div;
  float out_scale; int32_t accumulate;
  /* optional: emit the NEXT tensor-core layer's operand instead of y -- act_emit(value) split into bf16 planes hi / lo
   * [B, Lout, emit_ld] (what b2a_prep_bf16 would produce from y); depthwise stride-1 layers with Cout % 64 == 0 only
   * (snac/layers.py:208-231: Snake -> depthwise conv -> Snake -> 1x1 conv).  y may be NULL then. */
  void* emit_hi; void* emit_lo; int64_t emit_ld;
  int32_t emit_act; float emit_p0; const float* emit_a; const float* emit_b;
} b2a_conv1d_t;

int32_t b2a_conv1d_cl(const b2a_conv1d_t* p, void* stream);
int32_t b2a_convtr1d_cl(const b2a_conv1d_t* p, void* stream);

/* ---- tensor-core path for dense stride-1 convs / Linears (csrc/gemm_tc.cu) --------------------------------
 * b2a_prep_bf16: the conv prologue (pre_scale/shift + activation, as in b2a_conv1d_t) evaluated once per element and
 * stored as two bf16 planes hi = bf16(v), lo = bf16(v - hi), each [B, L, cpad] (cpad multiple of 64, pad channels zero);
 * lo == NULL stores hi only.
 * b2a_conv1d_tc: Y[b,l,n] = epilogue( sum_tap sum_ci (hi+lo)[b, l + shifts[tap], ci] * W[tap][n][ci] ), rows outside
 * [0,L) read as zero (TMA out-of-bounds fill == the conv's zero padding).  W is bf16 [taps][Cout][cin_pad]; Cout % 32 == 0.
 * tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands staged by TMA; epilogue fields as in b2a_conv1d_t.
 * f16 != 0: planes and weights are IEEE fp16 instead of bf16 (fp16 checkpoints such as Whisper's: weights stay exact).
 * up_stride > 0: TRANSPOSED conv with K = taps * up_stride in polyphase form -- Cout = up_stride * C, W[tap j][r*C + co][ci] =
 * w[k = r + j*up_stride][ci][co], shifts[j] = -j; GEMM row m / column (r, co) lands on output row m*up_stride + r - up_crop
 * (rows outside [0, Lout) dropped), i.e. the GEMM output IS the up-sampled signal, no col2im pass. */
int32_t b2a_prep_bf16(const float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t L, int32_t C, int32_t cpad,
                      const float* scale, const float* shift, int32_t act, float p0, const float* a, const float* b,
                      void* hi, void* lo, int32_t f16, void* stream);
int32_t b2a_conv1d_tc(const void* a_hi, const void* a_lo, int32_t f16, int32_t B, int32_t L, int32_t cin_pad, const void* w_bf16,
                      const void* w_lo, /* NULL, or the low plane bf16(w - w_hi) of an fp32 checkpoint's weights */
                      int32_t taps, const int32_t* shifts_host, int32_t Cout, int32_t Lout, const float* bias,
                      int32_t post_act, float post_p0, const float* cscale, int64_t cscale_bs, const float* res,
                      int64_t res_bs, int64_t res_ld, int32_t res_div, float out_scale, int32_t accumulate, float* y,
                      int64_t y_bs, int64_t y_ld, int32_t up_stride, int32_t up_crop, double* stats_ws,
                      int32_t stats_slots, void* stream);

/* profiling aid: CTA (0,0,0) of subsequent b2a_conv1d_tc launches stamps clock64() at its phase boundaries into dbg8[0..6]
 * (entry, setup done, first operands landed, last operands landed, accumulator ready, epilogue done, exit); NULL disables. */
int32_t b2a_conv1d_tc_debug(void* dbg8);

/* strided 2-D copy (concat without torch.cat): dst[r, c] = src[r, c] */
int32_t b2a_copy2d(const float* src, int64_t src_ld, float* dst, int64_t dst_ld, int64_t rows, int32_t cols, void* stream);
/* dst[r, :] = src[idx[r], :] (+ add[r % add_period, :] if add != NULL) -- the alignment expansion `x @ pred_aln_trg` of
 * kokoro.py:148-170, nn.Embedding, and token + positional embedding (whisper.py:483-486) */
int32_t b2a_gather_rows(const float* src, int64_t src_ld, const int64_t* idx, float* dst, int64_t dst_ld,
                        int64_t rows, int32_t cols, int64_t n_src_rows, const float* add, int64_t add_ld, int64_t add_period,
                        void* stream);
/* Duration head + alignment (kokoro.py:140-164): if dur_f != NULL, pred_dur[t] = clip(round_half_even(dur_f[t] / speed), 1, 100)
 * with nan->1, +inf->100, -inf->1 (mx.nan_to_num / mx.round / mx.clip); else pred_dur = dur_i.  Then idx_out[f] = token of
 * frame f (device prefix sum, frames beyond max_frames dropped) and *total_dev = sum(pred_dur). */
int32_t b2a_durations_to_index(const float* dur_f, const int64_t* dur_i, int32_t T, float speed, int64_t* pred_dur_out,
                               int64_t* idx_out, int64_t max_frames, int64_t* total_dev, void* stream);

/* ---- normalisation -----------------------------------------------------------------------
 * InstanceNorm statistics folded with the AdaIN affine (istftnet.py:216-268,327-338):
 *   scale[b,c] = (1+gamma[b,c]) * rstd[b,c],  shift[b,c] = beta[b,c] - scale[b,c]*mean[b,c]
 * with gb = [B, 2C] (gamma | beta) or NULL for plain InstanceNorm; biased variance, eps inside sqrt.
 * ws: float64 workspace of b2a_adain_ws_bytes(B,L,C) bytes. */
int64_t b2a_adain_ws_bytes(int32_t B, int32_t L, int32_t C);
int32_t b2a_adain_coeffs(const float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t L, int32_t C,
                         const float* gb, float eps, float* scale, float* shift, void* ws, void* stream);
/* Same coefficients from partial sums a producer already wrote: partials [B][nslots][C][2] float64 (sum, sum of squares per row group).
 * b2a_conv1d_tc(stats_ws != NULL) emits them from its epilogue -- one slot per 32 output rows (times the up-sampling factor in
 * transposed mode; stats_slots = ceil(Mrows/128)*4*max(1, up_stride)) -- so AdaIN-conv chains skip the statistics pass over HBM. */
int32_t b2a_adain_coeffs_from_partials(const double* partials, int32_t nslots, int32_t B, int32_t L, int32_t C, const float* gb,
                                       float eps, float* scale, float* shift, void* stream);
/* (sum, sumsq) over L of every channel of x [B, L, C], ADDED to n_dst (1..4) binned accumulators laid out [B][.][2][4] int64; dst[i] points
 * at the first channel's bins, dst_bs[i] int64 elements separate batches.  This is the statistics format b2a_conv1d_fused consumes (pre_mode 2)
 * and produces (stats_out); the stand-alone kernel covers tensors no fused conv produced (LSTM outputs, concatenated side channels). */
int32_t b2a_channel_stats(const float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t L, int32_t C, int64_t* const* dst,
                          const int64_t* dst_bs, int32_t n_dst, void* stream);
/* AdaIN (scale, shift) [B, C] from binned statistics [B, C, 2, 4] (for consumers outside the fused conv: depthwise layers). */
int32_t b2a_coeffs_from_stats(const int64_t* stats, int32_t B, int32_t L, int32_t C, const float* gb, float eps, float* scale, float* shift,
                              void* stream);
/* y[r,:] = LN(x[r,:] + res[r,:]) * w + b, or (1+ada[c])*LN + ada[C+c] when ada != NULL
 * (nn.LayerNorm, modules.py:71-90 AdaLayerNorm). rms != 0 -> RMSNorm (no mean, talker.py:267). */
int32_t b2a_layernorm(const float* x, int64_t x_ld, const float* res, int64_t res_ld, float* y, int64_t y_ld,
                      int64_t rows, int32_t C, const float* w, const float* b, const float* ada, float eps,
                      int32_t rms, int32_t post_act, float post_p0, void* stream);

/* ---- attention -----------------------------------------------------------------------------
 * softmax(scale * q k^T + mask) v with fp32 softmax; replaces mx.fast.scaled_dot_product_attention
 * (mimi/modules/transformer.py:109, talker.py:307) and the unfused form (whisper.py:371-385, modules.py:497-505).
 * q [B,Tq,H,D], k/v [B,Tk,Hkv,D] with token strides *_ld and batch strides *_bs (elements); head h at +h*D.
 * causal != 0: key j visible to query i iff j <= i + q_offset; window > 0: also i + q_offset - j < window. */
typedef struct {
  const float* q; const float* k; const float* v; float* o;
  int64_t q_bs, q_ld, k_bs, k_ld, v_bs, v_ld, o_bs, o_ld;
  int32_t B, Tq, Tk, H, Hkv, D;
  float scale; int32_t causal, q_offset, window;
  const int32_t* k_len;     /* [B] valid key count or NULL */
} b2a_attn_t;
int32_t b2a_attention(const b2a_attn_t* p, void* stream);
/* Same contract on the tensor cores (tcgen05) for D == 64, H == Hkv, k_len == NULL: S = QK^T and PV as fp16 hi/lo-plane MMAs
 * (fp32-grade products), online softmax with one thread per query row.  ws: device scratch of b2a_attention_tc_ws_bytes bytes
 * (fp16 planes of q, k and the transposed v). */
int64_t b2a_attention_tc_ws_bytes(int32_t B, int32_t H, int32_t Tq, int32_t Tk);
int32_t b2a_attention_tc(const b2a_attn_t* p, void* ws, void* stream);
/* in-place rotary embedding on x [B,T,H,D] (token stride ld): traditional != 0 rotates pairs (2i,2i+1)
 * (nn.RoPE(traditional=True), mimi/modules/transformer.py:75-77), else (i, i+D/2) (talker.py:14-18). */
int32_t b2a_rope(float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t T, int32_t H, int32_t D,
                 int32_t offset, float base, int32_t traditional, void* stream);

/* ---- bidirectional LSTM recurrence (modules.py:93-285) ------------------------------------
 * xproj [B,T,2,4H] = x @ Wx^T + b_ih + b_hh for (forward|backward), gate order i,f,g,o;
 * wh [2,4H,H]; out [B,T,2H] (forward | backward).  H must be 256 (Kokoro) -- one 8-CTA cluster
 * per (direction, batch) keeps Wh in registers and exchanges h through distributed shared memory. */
int32_t b2a_lstm_bidir(const float* xproj, const float* wh, float* out, int64_t out_ld, int32_t B, int32_t T, int32_t H, void* stream);

/* ---- DSP ------------------------------------------------------------------------------------
 * b2a_stft: dsp.py:385-433 on a batch of 1-D signals. out_re/out_im [B, frames, n_fft/2+1].
 * pad_mode: 0 none (center=False), 1 reflect, 2 constant.  window [n_fft] (already zero-padded). */
int32_t b2a_stft(const float* x, int64_t x_bs, int32_t B, int64_t n, const float* window, int32_t n_fft, int32_t hop,
                 int32_t pad_mode, int64_t frames, float* out_re, float* out_im, void* stream);
/* Whisper log-mel (stt/models/whisper/audio.py:41-82): frames-major [B, frames, n_mels]; `frames` excludes the
 * dropped last STFT frame; n = samples per signal (right zero padding of `padding` samples is virtual).
 * gmax [B] scratch for the per-utterance max. */
int32_t b2a_whisper_logmel(const float* x, int64_t x_bs, int32_t B, int64_t n, int64_t padding, const float* window,
                           const float* filters, int32_t n_mels, int64_t frames, float* out, float* gmax, void* stream);
/* dsp.py:436-513 / 663-738: inverse rFFT per frame, synthesis window, overlap-add, divide by sum(w^2) (norm_sq) or
 * sum(w), clamp denominators as the reference does (mode 0: where(wsum>1e-10); mode 1: max(wsum,1e-10)).
 * re/im [B, n_freq, T]; out [B, out_len] starting at sample `trim` of the OLA buffer. ws: B*T*n_fft floats. */
int32_t b2a_istft(const float* re, const float* im, int32_t B, int32_t n_fft, int32_t T, int32_t hop, const float* window,
                  int32_t norm_sq, int32_t clamp_mode, int64_t trim, int64_t out_len, float* out, float* ws, void* stream);

/* Polyphase resampling, the arithmetic of scipy.signal.resample_poly(x, up, down, window=h, padtype="edge") that the reference
 * calls on the host (resample.py:29-47): x [B, n_in] float32, h float64 FIR ALREADY multiplied by `up`, n_pre_pad / n_pre_remove
 * as SciPy derives them (host side: mlx_audio_b200/resample.py), out [B, n_out] float32, float64 accumulation. */
int32_t b2a_resample_poly(const float* x, int64_t x_bs, int32_t B, int64_t n_in, const double* h, int32_t n_h, int32_t up,
                          int32_t down, int64_t n_pre_pad, int64_t n_pre_remove, float* out, int64_t n_out, void* stream);

/* ---- Kokoro hn-NSF source + iSTFT head (istftnet.py:548-709, 453-545, 826-835) ------------
 * f0 [B, n_frames] (the F0 curve, one value per 300 samples); noise [B, n_frames*300, 9] injected N(0,1) (or NULL);
 * lin_w [9], lin_b [1] = m_source.l_linear.  har [B, n_frames*60+1, 22] = (|STFT| , angle) of the tanh-merged source,
 * n_fft 20 hop 5 periodic Hann, reflect-centred.  n_down = ceil(float(300*n_frames) * float(1/300)) is the length of the
 * reference's down-sampled phase track (interpolate.py:43-50; n_frames or n_frames+1 by floating-point rounding -- a parity
 * quirk the host evaluates exactly as the reference does).  src_ws: float [B, n_frames*300]; ph_ws: double [B, n_down, 9]. */
int32_t b2a_kokoro_source(const float* f0, int32_t B, int32_t n_frames, int32_t n_down, const float* noise, const float* lin_w,
                          const float* lin_b, float* har, float* src_ws, double* ph_ws, void* stream);
/* x [B, T, 22] = conv_post output -> audio [B, (T-1)*5] : exp / sin heads, cos/sin, 20-point inverse rFFT, periodic Hann,
 * overlap-add, / sum(w^2), trim 10 samples each side (phase in [-1,1] so mlx_unwrap is the identity). */
int32_t b2a_kokoro_istft_head(const float* x, int64_t x_bs, int64_t x_ld, int32_t B, int32_t T, float* audio, void* stream);

/* ---- fused dense conv1d / nn.Linear / polyphase ConvTranspose1d on tcgen05 (csrc/conv_fused.cu) ------------------------------
 * One launch = [InstanceNorm / AdaIN coefficients from the producer's (sum, sumsq)] + input activation + bf16/fp16 hi(+lo) split +
 * sum over taps of row-shifted GEMMs + bias / activation / channel scale / residual / scale / accumulate (+ polyphase scatter) +
 * (sum, sumsq) of the output for the NEXT layer's InstanceNorm.  Replaces, per layer, the call sites of mx.conv1d / conv_transpose1d /
 * nn.Linear TOGETHER WITH the elementwise chain in front of them: AdaIN1d + Snake / LeakyReLU in AdaINResBlock1 and AdainResBlk1d
 * (tts/models/kokoro/istftnet.py:216-396, 853-933), the generator's ups / conv_post (:725-835), ALBERT's Linear layers
 * (tts/models/kokoro/modules.py:434-645).  Up to B2A_CONVF_MAX_PROBLEMS independent problems share one persistent grid (the three
 * parallel resblocks of a generator stage).  x: fp32 [B, L, Cin] (16-byte aligned, row stride % 4 == 0); optional x1 / x2 (same
 * strides) are added to x first (the resblock average of the previous stage), then * in_scale.  Weights as for b2a_conv1d_tc:
 * 16-bit [taps][N][cin_pad] (+ optional lo plane).  y: fp32 [B, Lout, C].  pre_mode 0: no affine; 1: x*scale[b,c]+shift[b,c];
 * 2: scale/shift derived in-kernel from pre_stats [B, Cin, 2, 4] = (sum, sumsq) over L (biased variance, eps) and gamma|beta rows
 * pre_gb [B, 2 Cin] (NULL: plain InstanceNorm).  stats_out [B, C, 2, 4] is ADDED to (zero it before the launch).  A statistic is four
 * int64 bins counting multiples of 2^(-100 + 40 k): integer atomics make the accumulation order-independent, so results are
 * bit-reproducible across runs and between eager launches and graph replays (csrc/common.cuh: repro_add / repro_value).
 * ws: zero-initialised scratch (>= 16 MiB recommended) for split-K partial tiles; NULL disables K splitting. */
#define B2A_CONVF_MAX_PROBLEMS 4
typedef struct {
  const float* x; const float* x1; const float* x2; int64_t x_bs, x_ld; float in_scale;
  int32_t B, L, Cin;
  int32_t pre_mode; const float* pre_scale; const float* pre_shift; const int64_t* pre_stats; const float* pre_gb; int64_t pre_gb_bs; float pre_eps;
  int32_t pre_act; float pre_p0; const float* pre_a; const float* pre_b;
  const void* w_hi; const void* w_lo; int32_t cin_pad, taps, N; int32_t shifts[32];
  int32_t Lout; const float* bias; int32_t post_act; float post_p0; const float* cscale; int64_t cscale_bs;
  const float* res; int64_t res_bs, res_ld; int32_t res_div; float out_scale; int32_t accumulate;
  float* y; int64_t y_bs, y_ld;
  int32_t up_stride, up_crop;
  int64_t* stats_out;
} b2a_convf_t;
int32_t b2a_conv1d_fused_debug(void* stamps /* device uint64 [grid][16] or NULL: phase time stamps of the next launches */);
int32_t b2a_conv1d_fused(const b2a_convf_t* problems, int32_t n_problems, int32_t planes, int32_t f16, void* ws, int64_t ws_bytes, void* stream);

/* out[i] ~ N(0,1), i < n: Philox4x32-10 keyed by `seed`, counter `offset + i/4`, Box-Muller.  The production replacement for
 * mx.random.normal in SineGen / NoiseBlock (istftnet.py:649, snac/layers.py:263); parity tests inject the noise instead. */
int32_t b2a_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, void* stream);
/* Same draws with {seed, offset} read from DEVICE memory (state[0], state[1]); a second one-thread launch then advances state[1] by the
 * (n + 3) / 4 counters used, so a captured CUDA graph draws fresh noise on every replay the way mx.random's global state advances. */
int32_t b2a_randn_dev(float* out, int64_t n, uint64_t* state, void* stream);

/* ---- Whisper decode step (stt/models/whisper/decoding.py:307-325,349-442) -----------------------------------
 * One launch = SuppressBlank + SuppressTokens + ApplyTimestampRules + GreedyDecoder.update(temperature 0) for every row:
 * next_out[b] = argmax of the filtered logits (eot once a row has ended), sum_logprobs[b] += its log-probability while the row
 * is live, *not_done += 1 per row whose next token is not eot.  tokens [B, >= cur_len] is the device-resident history
 * (no per-step tolist()); suppress_mask / blank_mask are additive 0/-inf vectors [V] (NULL = none); max_initial_ts < 0 = off.
 * temperature > 0 (the fallback temperatures of whisper.py:957-995): the next token is a categorical draw from softmax(filtered / temperature)
 * -- inverse CDF in index order driven by u[b] in [0,1) -- and the log-probability bookkeeping uses the un-tempered logits, as
 * GreedyDecoder.update does (decoding.py:307-325). */
int32_t b2a_whisper_greedy_step(const float* logits, int64_t logits_bs, const int64_t* tokens, int64_t tokens_bs,
                                int32_t B, int32_t cur_len, int32_t sample_begin, int32_t V, const float* suppress_mask,
                                const float* blank_mask, int32_t eot, int32_t no_timestamps, int32_t timestamp_begin,
                                int32_t max_initial_ts, int32_t without_timestamps, int64_t* next_out,
                                float* sum_logprobs, int32_t* not_done, float temperature, const float* u, void* stream);

/* Fused LM sampler (tts/models/qwen3_tts/qwen3_tts.py:805-860 over lm/sample_utils.py:131-239,279): additive suppress mask ->
 * sign-aware repetition penalty on the `seen` set -> temperature (<= 0: argmax) -> top-k -> top-p -> min-p -> categorical draw
 * by inverse CDF in index order with the caller's uniform u[b] (MLX's PRNG is not reproducible; tests inject u).  V <= 4096.
 * out[b * out_stride] receives the token; mark_seen != 0 also sets seen[b][token] (generated_token_ids.append, :1402).
 * finished (optional, uint8 [B], in/out) implements the batch loop's bookkeeping (:1880-1887,1923-1929): a finished row emits eos
 * and is not marked; a row that samples eos becomes finished.
 * filtered_out (optional) receives the filtered, temperature-scaled logits the draw is made from. */
int32_t b2a_sample_token(const float* logits, int64_t logits_bs, int32_t B, int32_t V, const float* suppress_mask,
                         uint8_t* seen, int64_t seen_bs, int32_t mark_seen, float repetition_penalty, float temperature,
                         int32_t top_k, float top_p, float min_p, const float* u, int64_t* out, int64_t out_stride,
                         float* filtered_out, uint8_t* finished, int32_t eos, void* stream);

/* ---- autoregressive LM step (Qwen3-TTS talker / code predictor, tts/models/qwen3_tts/talker.py) -----------------------
 * All position-dependent scalars may come from device memory (base_dev, step_dev) so that one captured CUDA graph replays
 * every frame of Model.generate's loop (qwen3_tts.py:1323-1404) without host round trips.
 *
 * b2a_gemv_bf16: y[m, n] = sum_k W[n,k] xn[m,k] (+ bias[n]) (+ res[m,n]) for M <= 8 activation rows -- nn.Linear at decode
 * time (talker.py:284-286,314,335).  W bf16 row-major [N, w_ld].  norm_w != NULL fuses the preceding nn.RMSNorm
 * (talker.py:388,395; x * rsqrt(mean(x^2) + eps) * norm_w).  mode 1 fuses SwiGLU (talker.py:319-321): W rows interleaved
 * (gate_0, up_0, gate_1, ...), y[m, n/2] = silu(gate) * up, y has N/2 columns.  prefetch (optional): the NEXT projection's
 * weights, prefetch_bytes of them are pulled into L2 (prefetch.global.L2) while this kernel runs, so the dependent launch that
 * follows finds them on chip. */
int32_t b2a_gemv_bf16(const float* x, int64_t x_ld, int32_t M, int32_t K, const void* w_bf16, int64_t w_ld, int32_t N,
                      const float* bias, const float* norm_w, float norm_eps, int32_t mode, const float* res, int64_t res_ld,
                      float* y, int64_t y_ld, const void* prefetch, int64_t prefetch_bytes, void* stream);
/* TalkerAttention / CodePredictorAttention / DecoderAttention up to the cache update (talker.py:288-307,558-572;
 * speech_tokenizer.py:291-296): qkv [B,S,(Hq+2Hkv) D] -> per-head RMSNorm of q and k (weights [D], NULL = none), rotary
 * embedding in the rotate_half convention, q_out [B,S,Hq,D], k/v appended to the caches [B,Smax,Hkv,D] at row base + s,
 * base = *base_dev (or base_host when base_dev is NULL).  Rotary position of frequency slot i: pos3[axis,b,s] with the
 * interleaved-MRoPE axis rule of talker.py:139-184 (axis 1 if i%3==1 && i<3*sec_h, axis 2 if i%3==2 && i<3*sec_w, else 0);
 * pos3 NULL = base + s on every axis (sec_h = sec_w = 0 gives the standard RoPE of talker.py:68-113), minus pos_shift[b] (clamped at
 * 0) when pos_shift != NULL: the cumsum(attention_mask) - 1 positions of left-padded batches (talker.py:452-457). */
int32_t b2a_qknorm_rope_cache(const float* qkv, int64_t qkv_bs, int64_t qkv_ss, int32_t B, int32_t S, int32_t Hq, int32_t Hkv,
                              int32_t D, const float* q_norm_w, const float* k_norm_w, float eps, const int32_t* pos3,
                              const int32_t* base_dev, int32_t base_host, int32_t sec_h, int32_t sec_w, float theta,
                              float* q_out, int64_t qo_bs, int64_t qo_ss, float* k_cache, float* v_cache, int64_t c_bs,
                              int64_t c_ss, int32_t smax, const int32_t* pos_shift, void* stream);
/* mx.fast.scaled_dot_product_attention against the KV cache with GQA (talker.py:309-312): query s attends cache rows
 * [kv_start[b], base + s] (causal inside the new block; kv_start NULL = 0, else the left-padding count of
 * qwen3_tts.py:486-604's batches).  out [B,S,Hq*D].  max_k bounds base + S (shared-memory sizing). */
int32_t b2a_attn_decode(const float* q, int64_t q_bs, int64_t q_ss, const float* k_cache, const float* v_cache, int64_t c_bs,
                        int64_t c_ss, float* out, int64_t o_bs, int64_t o_ss, int32_t B, int32_t S, int32_t Hq, int32_t Hkv,
                        int32_t D, float scale, const int32_t* base_dev, int32_t base_host, const int32_t* kv_start,
                        int32_t max_k, void* stream);
/* Single-token decode (S = 1, GQA group of 2): b2a_qknorm_rope_cache + b2a_attn_decode in one launch, one CTA per (kv head, batch)
 * -- each cache row is read once for both query heads of the group.  qkv [B, (Hq+2Hkv) D]; out [B, Hq*D]; pos3 [3,B] or NULL. */
int32_t b2a_attn_decode_fused(const float* qkv, int64_t qkv_bs, int32_t B, int32_t Hq, int32_t Hkv, int32_t D, const float* q_norm_w,
                              const float* k_norm_w, float eps, const int32_t* pos3, const int32_t* base_dev, int32_t base_host,
                              int32_t sec_h, int32_t sec_w, float theta, float* k_cache, float* v_cache, int64_t c_bs, int64_t c_ss,
                              int32_t smax, float scale, const int32_t* kv_start, float* out, int64_t o_bs, void* stream);
/* y[r, i] = silu(gate) * up (talker.py:319-321, speech_tokenizer.py:321-322) for the batched (prefill) path: x [rows, 2I] holds
 * (gate | up) halves, or interleaved (gate_0, up_0, gate_1, ...) pairs -- the row order b2a_gemv_bf16 mode 1 uses. */
int32_t b2a_swiglu(const float* x, int64_t x_ld, int64_t rows, int32_t I, int32_t interleaved, float* y, int64_t y_ld, void* stream);
/* Next talker input (qwen3_tts.py:1383-1398): out[b] = text(b) + sum_g tables[g][codes[b,g]], text(b) = text[b, step] while
 * step = *step_dev - step_sub < n_text, else pad (tts_pad_embed); text/pad NULL = 0.  tables_dev / bins_dev are DEVICE arrays
 * of G table pointers / table sizes; an out-of-range code sets *err_flag_dev.  tidx != NULL selects the batch rule of
 * _next_batch_input_embeds(pad_when_index_clamped=True) (qwen3_tts.py:993-1015): text row min(tidx[b], n_text-1), replaced by pad
 * when that is >= n_text-1; afterwards tidx[b] += 1 for rows that are not finished. */
int32_t b2a_embed_sum(const int64_t* codes, int64_t codes_bs, int32_t B, int32_t G, int32_t dim, const float* const* tables_dev,
                      const int32_t* bins_dev, const float* text, int64_t text_bs, int64_t text_ss, int32_t n_text,
                      const float* pad, const int32_t* step_dev, int32_t step_sub, float* out, int64_t out_bs,
                      int32_t* err_flag_dev, int32_t* tidx, const uint8_t* finished, void* stream);
/* *p += v on the stream (KVCache.offset bookkeeping, lm/models/cache.py:112-155, kept on the device). */
int32_t b2a_incr_i32(int32_t* p, int32_t v, void* stream);

/* ---- codec (RVQ decode) ---------------------------------------------------------------------
 * out[b,t,:] (+)= sum_q codebooks[q][codes[b,q,t]][:]   (mimi/modules/quantization.py:47-49,103-108;
 * speech_tokenizer.py:431-490).  codes int64 [B, nq, T]; codebooks [nq, bins, dim]. Returns B2A_E_INVALID if any code
 * is out of range (checked on device, reported via *err_flag_dev != 0). */
int32_t b2a_rvq_decode(const int64_t* codes, int64_t codes_bs, int64_t codes_qs, int32_t B, int32_t nq, int64_t T,
                       const float* codebooks, int32_t bins, int32_t dim, float* out, int64_t out_ld, int32_t* err_flag_dev, void* stream);
/* Residual-VQ ENCODE: codes[row, q] = nearest entry of codebooks[q] to the row's residual, q = 0 .. nq-1, the residual shrinking by the
 * chosen entry after every level (mimi/modules/quantization.py:37-45, 90-101; speech_tokenizer.py:957-1058 uses the same quantizer).
 * x [rows, dim] fp32 (row stride x_ld); codebooks [nq, bins, dim]; c2 [nq, bins] float64: mode 0 (Euclidean) |e|^2 / 2, score = c2 - x.e;
 * mode 1 (SNAC, snac/vq.py:56-73: one level, L2-normalised rows -- codebooks must hold the NORMALISED table) |en|^2, score = |xn|^2 - 2 xn.en
 * + c2.  Scores accumulate in float64; ties -> lowest index.  codes int64, element (row, q) at row * codes_row_stride + q * codes_level_stride. */
int32_t b2a_rvq_encode(const float* x, int64_t x_ld, int64_t rows, int32_t dim, const float* codebooks, const double* c2, int32_t bins,
                       int32_t nq, int32_t mode, int64_t* codes, int64_t codes_row_stride, int64_t codes_level_stride, void* stream);
/* SNAC from_codes (snac/vq.py:111-131): z[b,t,:] = sum_l ( W_l @ E_l[codes_l[b, t / stride_l]] + bias_l ),
 * codes_l int64 [B, T/stride_l]; E_l [bins, cd]; W_l [cd][dim] (packed, K=1); out [B,T,dim]. */
int32_t b2a_snac_from_codes(const int64_t* const* codes_host_ptrs, const int32_t* strides_host, int32_t n_levels,
                            const float* const* emb_host_ptrs, const float* const* w_host_ptrs, const float* const* bias_host_ptrs,
                            int32_t B, int64_t T, int32_t bins, int32_t cd, int32_t dim, float* out, int32_t* err_flag_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200AUDIO_H */
