/*
 * funasr_b200 — C ABI of the B200-native (sm_100a) backend for FunASR's offline Paraformer hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*); no torch / C++
 * types cross the boundary.  All functions are stream-ordered, re-entrant, hold no global state, never
 * synchronise the device, and return FA_OK (0) or a negative FaStatus.  The caller owns every buffer
 * (inputs, outputs, workspace); weights are borrowed pointers.
 *
 * Each entry point names the reference interface it replaces (paths relative to the FunASR tree,
 * commit 3c58cb5 / funasr 1.4.3).  The tensor-level operator boundary mirrors the reference's own
 * export signature  export_forward(speech[B,T,560], speech_lengths[B]) -> (logits[B,N,V], token_num[B])
 * (funasr/models/paraformer/export_meta.py) extended upward by the frontend and downward by greedy ids.
 *
 * Layout conventions: row-major, innermost dimension contiguous; a "row" is one (utterance, frame) or
 * (utterance, token) pair; nn.Linear weights keep their [out_features, in_features] layout.
 */
#ifndef FUNASR_B200_H_
#define FUNASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fa_stream_t; /* cudaStream_t */

typedef enum {
  FA_OK = 0,
  FA_ERR_ARG = -1,         /* null pointer / bad size */
  FA_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
  FA_ERR_WORKSPACE = -3,   /* workspace too small */
  FA_ERR_UNSUPPORTED = -4  /* shape outside what the kernels are built for */
} FaStatus;

/* GEMM arithmetic modes (FaLinear contractions). fp32 accumulate everywhere. */
typedef enum {
  FA_GEMM_F32_SIMT = 0, /* fp32 FFMA tiles: the parity reference path */
  FA_GEMM_F16X1 = 1,   /* tcgen05 kind::f16, one fp16 pass (fast mode) */
  FA_GEMM_F16X3 = 3,   /* tcgen05, fp16 planes x = hi + lo: hi*hi + hi*lo + lo*hi (~2^-22 relative per product) */
  FA_GEMM_F16X6 = 6    /* tcgen05, three fp16 planes, six products (~fp32) */
} FaGemmMode;

/* nn.Linear: y = x W^T + b.  w_planes (optional) holds the fp16 planes made by fa_split_planes for the
 * tcgen05 path: [3][out_f][in_pad] fp16 (hi, mid, lo), in_pad = in_f rounded up to 64. */
typedef struct {
  const float* w;        /* [out_f, in_f] */
  const float* b;        /* [out_f] or NULL */
  const void* w_planes;  /* or NULL */
  int32_t out_f, in_f, in_pad, _pad;
} FaLinear;

typedef struct {
  const float* g; /* weight [n] */
  const float* b; /* bias   [n] */
  int32_t n;
  float eps;
} FaNorm;

/* EncoderLayerSANM (funasr/models/sanm/encoder.py:44-148) */
typedef struct {
  FaNorm norm1;          /* over in_size (560 for encoders0, 512 otherwise) */
  FaLinear qkv;          /* self_attn.linear_q_k_v  [1536, in_size] */
  const float* fsmn_w;   /* self_attn.fsmn_block.weight [512, 11] (depthwise) */
  FaLinear out;          /* self_attn.linear_out    [512, 512] */
  FaNorm norm2;
  FaLinear w1;           /* feed_forward.w_1 [2048, 512] */
  FaLinear w2;           /* feed_forward.w_2 [512, 2048] */
} FaEncLayer;

/* SANMEncoder (funasr/models/sanm/encoder.py:188-461), input_layer == "pe" */
typedef struct {
  const FaEncLayer* layers;  /* host array, n_layers entries; [0] is encoders0.0 */
  int32_t n_layers;
  int32_t heads;
  int32_t fsmn_k;            /* 11 */
  int32_t _pad;
  FaNorm after_norm;
  const float* pe_inv_timescales; /* [in_size/2] device, transformer/embedding.py:409-414 */
} FaEncoder;

/* CifPredictorV2 (funasr/models/paraformer/cif_predictor.py:209-314) */
typedef struct {
  FaLinear conv;      /* cif_conv1d as a GEMM: weight repacked to [512, 3*512], W[n, k*512+c] = w[n,c,k] */
  const float* out_w; /* cif_output.weight [512] */
  const float* out_b; /* cif_output.bias   [1]   */
  float threshold;      /* 1.0 */
  float tail_threshold; /* 0.45 */
  float smooth_factor;  /* 1.0 */
  float noise_threshold;/* 0.0 */
  int32_t cif_variant;  /* 0: CifPredictorV2 `cif_v1` (fp64 prefix sums, paraformer/cif_predictor.py:818-908);
                         * 1: CifPredictorV3 `cif` (sequential fp32, bicif_paraformer/cif_predictor.py:37-84) */
  int32_t _pad;
} FaPredictor;

/* DecoderLayerSANM (funasr/models/paraformer/decoder.py:26-121) */
typedef struct {
  FaNorm norm1;
  FaLinear ffn_w1;       /* feed_forward.w_1 [2048, 512] */
  FaNorm ffn_norm;       /* feed_forward.norm over 2048 */
  FaLinear ffn_w2;       /* feed_forward.w_2 [512, 2048], no bias */
  FaNorm norm2;
  const float* fsmn_w;   /* self_attn.fsmn_block.weight [512, 11]; NULL for decoders3 */
  FaNorm norm3;
  FaLinear q;            /* src_attn.linear_q   [512, 512] */
  FaLinear kv;           /* src_attn.linear_k_v [1024, 512] */
  FaLinear out;          /* src_attn.linear_out [512, 512] */
} FaDecLayer;

/* ParaformerSANMDecoder (funasr/models/paraformer/decoder.py:234-449) */
typedef struct {
  const FaDecLayer* layers; /* host array, n_layers entries (decoders) */
  int32_t n_layers;
  int32_t heads;
  int32_t fsmn_k;
  int32_t vocab;
  FaDecLayer last;          /* decoders3.0: norm1 + ffn only */
  FaNorm after_norm;
  FaLinear output;          /* output_layer [vocab, 512] */
  /* ContextualParaformerDecoder (funasr/models/contextual_paraformer/decoder.py:133-352); has_bias == 0 for the plain
   * Paraformer decoder.  With has_bias: `layers` are decoders.{0..att_layer_num-2}, `bias_last` is last_decoder, and
   * x = x_self_attn + bias_output([x_src_attn ; clas_scale * bias_decoder(x_self_attn, hotword memory)]) (:330-340). */
  int32_t has_bias;
  int32_t n_hotwords;       /* rows of hw_embed, <= t_max */
  FaDecLayer bias_last;     /* last_decoder (ContextualDecoderLayer :22-100) */
  FaNorm bias_norm3;        /* bias_decoder.norm3 */
  FaLinear bias_q, bias_kv, bias_out; /* bias_decoder.src_attn.linear_q / linear_k_v / linear_out */
  FaLinear bias_output;     /* bias_output Conv1d(1024->512, k=1, no bias) as a [512, 1024] linear */
  const float* hw_embed;    /* [n_hotwords, 512] hotword memory (LSTM last hidden states, model.py:350-372) */
  const int32_t* hw_lens;   /* [batch] device, every entry == n_hotwords */
  float clas_scale;         /* 1.0 */
  int32_t _pad2;
} FaDecoder;

/* ---------------------------------------------------------------------------------------------
 * Library info
 * ------------------------------------------------------------------------------------------- */
const char* fa_version(void);
/* Monotone count of kernel launches issued by this library in this process (bench "gpu_launches"). */
uint64_t fa_launch_count(void);
const char* fa_status_string(int status);

/* ---------------------------------------------------------------------------------------------
 * Frontend — replaces WavFrontend.forward (funasr/frontends/wav_frontend.py:149-196), i.e.
 * torchaudio.compliance.kaldi.fbank (dither=0, hamming, 25 ms / 10 ms, 80 mel, snip_edges) + apply_lfr
 * (:63-86, m=7 n=6) + apply_cmvn (:46-60), fused.  wav is float32 in [-1,1] (x32768 applied inside,
 * :169).  Rows t >= feat_lens[b] of feats are zero-filled (pad_sequence(..., 0.0), :195).
 *   wav [B, wav_stride], wav_lens[B] (samples, >= 400), cmvn [2,560] or NULL,
 *   mel_banks [80,257] (kaldi.py get_mel_banks + zero column), window [400] (hamming),
 *   feats [B, t_max, 560], feat_lens [B].
 * ------------------------------------------------------------------------------------------- */
int fa_fbank_lfr_cmvn(const float* wav, const int32_t* wav_lens, int32_t batch, int64_t wav_stride,
                      const float* cmvn, const float* mel_banks, const float* window,
                      float* feats, int32_t* feat_lens, int32_t t_max, fa_stream_t stream);
/* Same, writing utterance b at feats + b * feats_batch_stride_rows * 560 (t_max rows each): lets the caller leave room
 * for prepended frames, e.g. SenseVoiceSmall's 4 query frames (funasr/models/sense_voice/model.py:971-995). */
int fa_fbank_lfr_cmvn_strided(const float* wav, const int32_t* wav_lens, int32_t batch, int64_t wav_stride,
                              const float* cmvn, const float* mel_banks, const float* window, float* feats,
                              int64_t feats_batch_stride_rows, int32_t* feat_lens, int32_t t_max, fa_stream_t stream);
/* The same kernel fed with PRECOMPUTED constants (what the engines use): fa_fbank_make_tables derives, once per configuration,
 * the sparse support of the 80 mel filters, the FFT twiddles and the window into `tables` (fa_fbank_tables_bytes() bytes of
 * device memory), so that a CTA copies 9 KB instead of re-scanning the [80, 257] filter matrix.  lfr_m / lfr_n select the
 * low-frame-rate stacking: 7 / 6 (Paraformer, SenseVoice: feats [B, t_max, 560]) or 5 / 1 (the FSMN-VAD frontend,
 * fsmn_vad_streaming/template.yaml:54-62: feats [B, t_max, 400], one row per 10 ms frame); cmvn is [2, 80 * lfr_m] or NULL. */
size_t fa_fbank_tables_bytes(void);
int fa_fbank_make_tables(const float* mel_banks, const float* window, float* tables, fa_stream_t stream);
int fa_fbank_lfr_cmvn_tables(const float* wav, const int32_t* wav_lens, int32_t batch, int64_t wav_stride, const float* cmvn,
                             const float* tables, int32_t lfr_m, int32_t lfr_n, float* feats, int64_t feats_batch_stride_rows,
                             int32_t* feat_lens, int32_t t_max, fa_stream_t stream);
/* One utterance shorter than a 25 ms frame (2 <= n_samples < 400): WavFrontend.forward passes frame_length = min(25 ms, len / fs)
 * (funasr/frontends/wav_frontend.py:174), so kaldi.fbank uses the whole utterance as ONE window of n_samples (hamming, `window`),
 * zero-padded to padded_fft = the next power of two, with mel_banks [80, padded_fft / 2 + 1] built for that FFT size; the single
 * log-mel frame is repeated lfr_m times and CMVN'd into feats_row [80 * lfr_m].  (The batched kernels above give such rows
 * feat_lens = 0.) */
int fa_fbank_short(const float* wav, int32_t n_samples, const float* window, const float* mel_banks, int32_t padded_fft,
                   const float* cmvn, int32_t lfr_m, float* feats_row, fa_stream_t stream);
/* dst[b, r, :] = rows[r, :] for r < n_rows (dst rows of `cols` floats, utterances dst_batch_stride_rows apart):
 * the query-frame prepend `torch.cat((input_query, speech), dim=1)` of sense_voice/model.py:985-995. */
int fa_broadcast_rows(const float* rows, int32_t n_rows, int32_t cols, float* dst, int64_t dst_batch_stride_rows,
                      int32_t batch, fa_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Operator-level entry points (each is used by the model-level calls below and exposed for parity tests)
 * ------------------------------------------------------------------------------------------- */
/* LayerNorm over the last dim (funasr/models/transformer/layer_norm.py:13-39).  If pe_inv != NULL the
 * input row (b,t) is first mapped x*xscale + PE(t+1) (SANMEncoder.forward encoder.py:409,428 with
 * SinusoidalPositionEncoder embedding.py:396-432); rows_per_batch gives t = row % rows_per_batch. */
int fa_layernorm(const float* x, int64_t rows, const FaNorm* norm, float* y,
                 const float* pe_inv, float xscale, int32_t rows_per_batch, fa_stream_t stream);

/* y[rows, out_f] = act(x[rows, in_f(ldx)] W^T + b) (+ res1) (+ res2); replaces torch.nn.Linear calls
 * (attention.py:256,306; positionwise_feed_forward.py:34).  relu != 0 applies ReLU before residuals. */
int fa_linear(const float* x, int64_t ldx, int64_t rows, const FaLinear* lin, int32_t relu,
              const float* res1, int64_t ld_res1, const float* res2, int64_t ld_res2,
              float* y, int64_t ldy, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream);

/* The tensor-core GEMM alone, A operand already split into fp16 planes [npl][rows][lin->in_pad] (npl = 1 / 2 / 3 for
 * F16X1 / X3 / X6) by fa_split_rows: what the model-level calls launch between fused producers and consumers. */
int fa_split_rows(const float* x, int64_t ldx, int64_t rows, int32_t cols, int32_t cols_pad, int32_t nplanes, void* planes,
                  fa_stream_t stream);
int fa_linear_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, const float* res1, int64_t ld_res1,
                     const float* res2, int64_t ld_res2, float* y, int64_t ldy, int32_t gemm_mode, fa_stream_t stream);
/* Same GEMM with the plane-emitting epilogue: out_planes [npl][rows][ld_out] fp16 (hi, lo, ...) of act(A W^T + b) — the launch the
 * encoder makes for FFN w_1, whose ReLU output feeds w_2 without an fp32 round trip (bench.py times exactly this launch). */
int fa_linear_planes_to_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, void* out_planes,
                               int64_t ld_out, int32_t gemm_mode, fa_stream_t stream);

/* FSMN memory block: out = m * (v*m + dwconv_k(v*m)) (+ res); m[t] = t < lens[b]
 * (MultiHeadedAttentionSANM.forward_fsmn attention.py:216-239; decoder variant :583-631).  out must not overlap v or res (the staged
 * kernel reads whole tiles ahead of its stores). */
int fa_fsmn(const float* v, int64_t ldv, const int32_t* lens, int32_t batch, int32_t t_max, int32_t channels,
            const float* w, int32_t ksize, const float* res, int64_t ld_res, float* out, int64_t ld_out,
            fa_stream_t stream);
/* The same memory block through the TMA-staged, warp-specialised kernel (persistent CTAs, cp.async.bulk.tensor ring of
 * [64 + k - 1] x 128-channel boxes, results bit-identical to fa_fsmn).  FA_ERR_UNSUPPORTED unless ksize is 11 or 21, channels is a
 * multiple of 128 and v / res are 16-byte aligned with pitches that are multiples of 4 floats.  fa_fsmn and the model-level calls
 * take this route by default when the shape allows it and t_max >= 64 (FA_FSMN_TMA=0: SIMT kernel only). */
int fa_fsmn_tma(const float* v, int64_t ldv, const int32_t* lens, int32_t batch, int32_t t_max, int32_t channels,
                const float* w, int32_t ksize, const float* res, int64_t ld_res, float* out, int64_t ld_out,
                fa_stream_t stream);
/* The SIMT strip kernel for every shape (ksize 11 / 21 / 31, any channel count or alignment): the A/B partner of fa_fsmn_tma. */
int fa_fsmn_simt(const float* v, int64_t ldv, const int32_t* lens, int32_t batch, int32_t t_max, int32_t channels,
                 const float* w, int32_t ksize, const float* res, int64_t ld_res, float* out, int64_t ld_out,
                 fa_stream_t stream);

/* Multi-head scaled-dot attention with key-padding mask: masked_fill(-inf) -> softmax -> masked_fill(0)
 * -> @V, heads merged (attention.py:288-304 self, :760-794 cross).  head_dim is 128.
 *   q [B, tq, ldq], k/v [B, tk, ldk/ldv] (head h at column h*128), ctx [B, tq, ld_ctx]. */
int fa_attention(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                 const int32_t* key_lens, int32_t batch, int32_t heads, int32_t tq, int32_t tk,
                 float* ctx, int64_t ld_ctx, fa_stream_t stream);

/* Same contract on the tcgen05 tensor cores (fp16 operand planes, fp32 accumulation in TMEM):
 * gemm_mode FA_GEMM_F16X1 (one plane) or FA_GEMM_F16X3/X6 (hi+lo planes, three MMA terms).  workspace holds the
 * operand planes (size from fa_attention_tc_workspace_bytes). */
size_t fa_attention_tc_workspace_bytes(int32_t batch, int32_t heads, int32_t tq, int32_t tk, int32_t gemm_mode);
int fa_attention_tc(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                    const int32_t* key_lens, int32_t batch, int32_t heads, int32_t tq, int32_t tk,
                    float* ctx, int64_t ld_ctx, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Model-level entry points
 * ------------------------------------------------------------------------------------------- */
/* SANMEncoder.forward (encoder.py:392-461): feats [B,T,560], lens[B] -> enc [B,T,512].
 * If enc->pe_inv_timescales == NULL the struct describes a plain stack of 512->512 SAN-M layers applied to an
 * existing [B,T,512] stream (no x*sqrt(d)+PE, every layer has its residual): SenseVoiceEncoderSmall's `tp_encoders`
 * + `tp_norm` (sense_voice/model.py:650-655). */
size_t fa_sanm_encoder_workspace_bytes(int32_t batch, int32_t t_max, int32_t gemm_mode);
int fa_sanm_encoder_forward(const FaEncoder* enc, const float* feats, const int32_t* lens, int32_t batch,
                            int32_t t_max, float* out, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                            fa_stream_t stream);

/* CifPredictorV2.forward, inference branch (cif_predictor.py:253-314 + tail_process_fn :414-446 + cif_v1
 * :853-908).  enc [B,T,512], lens[B] ->
 *   acoustic [B, n_cap, 512] (rows >= fires zero-filled; n_cap >= 1, tokens beyond n_cap are dropped),
 *   token_num [B] int32 (= floor(sum alpha'), the reference's pre_token_length),
 *   alphas [B, T+1], peaks [B, T+1] (cif_peak / "fires"). */
size_t fa_cif_predictor_workspace_bytes(int32_t batch, int32_t t_max, int32_t gemm_mode);
int fa_cif_predictor_forward(const FaPredictor* pred, const float* enc, const int32_t* lens, int32_t batch,
                             int32_t t_max, float* acoustic, int32_t n_cap, int32_t* token_num, float* alphas,
                             float* peaks, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                             fa_stream_t stream);

/* out[r] = x[r, :n].sum() in fp32 with the summation order of torch's CPU kernel (ATen SumKernel.cpp: 8 SIMD lanes x 4 ILP
 * accumulators, 4-level cascade): the CIF predictor's integer token count is floor(alphas.sum(-1)) (cif_predictor.py:443-444),
 * so the order decides an integer outcome.  x [rows, ld] fp32 on the device, out [rows]. */
int fa_row_sum_f32(const float* x, int64_t ld, int32_t rows, int32_t n, float* out, fa_stream_t stream);

/* Timestamp head of CifPredictorV3.get_upsample_timestamp (bicif_paraformer/cif_predictor.py:331-352), after the
 * ConvTranspose1d upsampling (a fa_linear with the [3*512, 512] repacked weight) and the BLSTM:
 *   alphas2 = relu(sigmoid(feat . w + b) * smooth2 - noise2) * mask;  alphas2 *= token_num / sum(alphas2);
 *   us_peaks = cif_wo_hidden(alphas2, threshold - 1e-4).
 * feat [B, t_up, dz] (dz = 1024), lens_up[B] = 3 * encoder lengths, token_num[B]; us_alphas / us_peaks [B, t_up]. */
int fa_cif_upsample_alphas(const float* feat, int32_t dz, const float* w, const float* b, const int32_t* lens_up,
                           const int32_t* token_num, int32_t batch, int32_t t_up, float smooth2, float noise2,
                           float threshold, float* us_alphas, float* us_peaks, fa_stream_t stream);

/* One-layer bidirectional LSTM recurrence (torch.nn.LSTM(512, 512, 1, batch_first=True, bidirectional=True), the `blstm` of
 * CifPredictorV3, bicif_paraformer/cif_predictor.py:187-190) as a persistent weight-stationary kernel.  The caller supplies
 * the input projections of ALL steps (one fa_linear): xproj [B*T, 4096] = x [W_ih_fwd; W_ih_bwd]^T + (b_ih + b_hh), gate order
 * i,f,g,o per direction.  w_hh_* [2048, 512].  out [B, T, 1024] (forward | reverse).  batch <= 256, hidden == 512.
 * sync_scratch8: 8 bytes of device memory (zeroed by the call) for the per-direction step barrier. */
int fa_blstm_forward(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, int32_t batch, int32_t t_len,
                     int32_t hidden, float* out, void* sync_scratch8, fa_stream_t stream);

/* Tensor-core variant (default in the plugin): the per-step [B,512] x [512,2048] product on warp-level fp16 MMAs with the
 * 3-product operand split (fp32 accumulate), h exchanged between CTAs as fp16 hi / lo planes.  Same contract as
 * fa_blstm_forward; scratch >= fa_blstm_tc_scratch_bytes(batch) bytes of device memory (zeroed by the call). */
size_t fa_blstm_tc_scratch_bytes(int32_t batch);
int fa_blstm_forward_tc(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, int32_t batch, int32_t t_len,
                        int32_t hidden, float* out, void* scratch, size_t scratch_bytes, fa_stream_t stream);

/* Measurement aid for fa_blstm_forward: skip_mask bit 0 drops the recurrent dot products, bit 1 the h gather, bit 2 the step
 * barrier (results are then meaningless); used by tools/bicif_probe.py to attribute the step time. */
int fa_debug_blstm_variant(int32_t skip_mask, const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, int32_t batch,
                           int32_t t_len, float* out, void* sync_scratch8, fa_stream_t stream);

/* ParaformerSANMDecoder.forward (decoder.py:397-449) + greedy argmax (paraformer/model.py:642-644).
 *   enc [B,T,512], enc_lens[B]; acoustic [B, ld_acoustic_rows, 512] of which the first n_max rows are used;
 *   tok_lens[B].  Outputs: argmax_ids [B, n_max] int32, argmax_logp [B, n_max] (log-softmax value of the
 *   arg-max, :643), and — if logits != NULL — the full pre-softmax logits [B, n_max, vocab].
 *   If log_softmax != 0 the logits buffer is converted in place to log_softmax (model.py:345). */
size_t fa_paraformer_decoder_workspace_bytes(int32_t batch, int32_t t_max, int32_t n_max, int32_t vocab,
                                             int32_t gemm_mode);
/* Same, for a contextual decoder (FaDecoder.has_bias) with n_hotwords entries in its hotword memory; the plain query above
 * assumes n_hotwords <= t_max. */
size_t fa_paraformer_decoder_workspace_bytes_hw(int32_t batch, int32_t t_max, int32_t n_max, int32_t vocab,
                                                int32_t gemm_mode, int32_t n_hotwords);
int fa_paraformer_decoder_forward(const FaDecoder* dec, const float* enc, const int32_t* enc_lens, int32_t batch,
                                  int32_t t_max, const float* acoustic, int64_t ld_acoustic_rows,
                                  const int32_t* tok_lens, int32_t n_max, int32_t* argmax_ids, float* argmax_logp,
                                  float* logits, int32_t log_softmax, int32_t gemm_mode, void* workspace,
                                  size_t ws_bytes, fa_stream_t stream);

/* The same with return_hidden + return_both (decoder.py:441-449): additionally writes the after_norm output `hidden`
 * [B, n_max, 512] — the `decoder_hidden` SeacoParaformer feeds to its hotword decoder (seaco_paraformer/model.py:290-297). */
int fa_paraformer_decoder_forward_hidden(const FaDecoder* dec, const float* enc, const int32_t* enc_lens, int32_t batch,
                                         int32_t t_max, const float* acoustic, int64_t ld_acoustic_rows,
                                         const int32_t* tok_lens, int32_t n_max, int32_t* argmax_ids, float* argmax_logp,
                                         float* logits, int32_t log_softmax, float* hidden, int32_t gemm_mode,
                                         void* workspace, size_t ws_bytes, fa_stream_t stream);

/* A SAN-M decoder stack WITHOUT input / output layer over an arbitrary memory: the SeACo decoder of SeacoParaformer
 * (seaco_paraformer/model.py:100-110: ParaformerSANMDecoder(use_output_layer=False, wo_input_layer=True), FFN 1024, FSMN k=21,
 * 6 attention layers) attending over the hotword embeddings.  Uses dec->layers / n_layers / heads / fsmn_k / last / after_norm.
 *   memory [t_mem, 512] when mem_shared != 0 (the same hotword memory for every utterance, model.py:306-308), else
 *   [B, t_mem, 512]; mem_lens[B]; x [B, ld_x_rows, 512] of which n_max rows are used; tok_lens[B].
 *   n_run attention layers are run (<= dec->n_layers), then
 *     finish != 0     : decoders3 + after_norm -> hidden [B, n_max, 512]  (ParaformerSANMDecoder.forward, decoder.py:397-449)
 *     attn_probs != 0 : layer n_run-1 stops at its cross-attention and writes utterance 0's probability matrix
 *                       [heads, n_max, t_mem] (forward_asf6 / get_attn_mat, decoder.py:485-513, :123-146); hidden untouched. */
size_t fa_sanm_decoder_stack_workspace_bytes(int32_t batch, int32_t t_mem, int32_t n_max, int32_t gemm_mode);
int fa_sanm_decoder_stack_forward(const FaDecoder* dec, const float* memory, const int32_t* mem_lens, int32_t mem_shared,
                                  int32_t batch, int32_t t_mem, const float* x, int64_t ld_x_rows, const int32_t* tok_lens,
                                  int32_t n_max, int32_t n_run, int32_t finish, float* hidden, float* attn_probs,
                                  int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream);

/* ids / best_logp [rows] = arg-max and its log-softmax value of (a (+ b)) W^T + bias — SeACo's hotword_output_layer over
 * cif_attended + dec_attended (seaco_paraformer/model.py:351-355); logp != NULL receives the full log-softmax rows [rows, out_f]. */
size_t fa_linear_argmax_workspace_bytes(int64_t rows, int32_t vocab, int32_t gemm_mode);
int fa_linear_argmax(const FaLinear* lin, const float* a, const float* b_or_null, int64_t rows, int32_t* ids, float* best_logp,
                     float* logp, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream);

/* SeACo merge with seaco_weight = 1 (seaco_paraformer/model.py:357-378): per token row, the decoder's arg-max where the hotword
 * decoder's arg-max is NO_BIAS, else the hotword decoder's.  merged != NULL additionally receives the merged log-prob rows
 * (dec_logp / dha_logp: full log-softmax rows [rows, vocab]). */
int fa_seaco_merge(const int32_t* dec_ids, const float* dec_best, const int32_t* dha_ids, const float* dha_best, int64_t rows,
                   int32_t no_bias, int32_t* out_ids, float* out_best, const float* dec_logp, const float* dha_logp,
                   float* merged, int32_t vocab, fa_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * FSMN-VAD (funasr/models/fsmn_vad_streaming): the encoder FSMN.forward (encoder.py:355-377) over the LFR-5/1 features of a whole
 * waveform, reduced to the silence posterior per 10 ms frame that the end-point detector reads (model.py:789-792), and the frame
 * energies of ComputeDecibel (model.py:458-529).  The detector itself is sequential threshold logic and stays on the host
 * (funasr_b200/vad.py), as in the reference.
 *   Weights are fp32 [out, in_padded] with the input dimension zero-padded to a multiple of 16 (FaLinear.in_f = padded width):
 *   in1 400->140, in2 140(144)->250 + ReLU, per layer lin 250(256)->128 (no bias), conv_w [128, lorder] (causal depthwise taps,
 *   tap lorder-1 = the current frame), affine 128->250 + ReLU, out1 250(256)->140, out2 140(144)->248.
 *   feats [t, ld_feats] -> sil_prob [t]; scores != NULL additionally receives the full softmax [t, out2.out_f].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  FaLinear lin;
  const float* conv_w;
  FaLinear affine;
} FaVadLayer;
typedef struct {
  FaLinear in1, in2;
  const FaVadLayer* layers;
  int32_t n_layers, lorder;
  FaLinear out1, out2;
  int32_t sil_ids[4];
  int32_t n_sil, _pad;
} FaVadEncoder;
size_t fa_fsmn_vad_workspace_bytes(const FaVadEncoder* enc, int32_t t);
int fa_fsmn_vad_forward(const FaVadEncoder* enc, const float* feats, int64_t ld_feats, int32_t t, float* sil_prob, float* scores,
                        void* workspace, size_t ws_bytes, fa_stream_t stream);
/* Host-side end-point detector (no GPU work): the per-frame silence posteriors and frame energies of ONE whole recording ->
 * [start_ms, end_ms] segments, as FsmnVADStreaming.inference produces them over its 60 s chunks (fsmn_vad_streaming/model.py:
 * GetFrameState :761-823, WindowDetector :218-320, DetectOneFrame :1158-1302, dynamic end-silence schedule :1003-1067).  The fields
 * are VADXOptions' (model.py:71-174).  schedule: n_schedule pairs {accumulated-speech limit in ms (< 0 = no limit), end silence in ms},
 * used when dynamic_silence != 0 (the reference's default when no max_end_silence_time is given); speech_noise_thres: NaN = the
 * options' value.  Returns the number of segments found (segments receives the first max_segments {start, end} pairs), or a negative
 * status (FA_ERR_ARG also for posteriors outside (0, 1), where the reference's math.log raises). */
typedef struct {
  int32_t sample_rate, detect_mode, max_end_silence_time, max_start_silence_time, window_size_ms, sil_to_speech_time_thres,
      speech_to_sil_time_thres, do_extend, lookback_time_start_point, lookahead_time_end_point, max_single_segment_time,
      noise_frame_num_used_for_snr, frame_in_ms, frame_length_ms;
  double speech_2_noise_ratio, snr_thres, decibel_thres, speech_noise_thres, fe_prior_thres;
} FaVadOptions;
int64_t fa_vad_detect_segments(const double* sil_prob, const double* decibel, int64_t frames, int64_t n_samples, const FaVadOptions* opts,
                               int32_t chunk_ms, int32_t dynamic_silence, const double* schedule, int32_t n_schedule,
                               double speech_noise_thres, int32_t* segments, int64_t max_segments);
/* Host only: the integrate-and-fire trace of one utterance's weights, funasr/utils/timestamp_tools.py:14-34 (cif_wo_hidden: fp32 running
 * sum, reduced by `threshold` after every frame that reaches it; trace[t] = the value before the reduction) — the re-integration step of
 * ts_prediction_lfr6_standard (:67-72). */
int fa_cif_wo_hidden_host(const float* alphas, int64_t n, float threshold, float* trace);
/* decibel[f] = 10 log10(sum_{j<400} wav[160 f + j]^2 + 1e-6), f < frames (ComputeDecibel, model.py:516-525). */
int fa_frame_decibels(const float* wav, int64_t n_samples, int32_t frames, float* decibel, fa_stream_t stream);

/* Greedy post-filter (paraformer/model.py:655-666): keep argmax_ids[b, k] for k < tok_lens[b] that are not
 * in {blank=0, sos=1, eos=2}; out_ids [B, n_max] (padded with -1), out_lens [B]. */
int fa_greedy_filter(const int32_t* argmax_ids, const int32_t* tok_lens, int32_t batch, int32_t n_max,
                     int32_t sos, int32_t eos, int32_t blank, int32_t* out_ids, int32_t* out_lens,
                     fa_stream_t stream);

/* CTC greedy head of SenseVoiceSmall (sense_voice/model.py:1003-1025, ctc/ctc.py:192-203): logits = enc W^T + b,
 * log_softmax, arg-max per frame, torch.unique_consecutive, drop blank.
 *   enc [B,T,512], lens[B] -> out_ids [B,T] (padded with -1), out_lens [B]; if logp != NULL it receives the
 *   full log_softmax [B,T,vocab] (parity checks); argmax_ids [B,T] is scratch/diagnostic output. */
size_t fa_ctc_greedy_workspace_bytes(int32_t batch, int32_t t_max, int32_t vocab, int32_t gemm_mode);
int fa_ctc_greedy_forward(const FaLinear* ctc_lo, const float* enc, const int32_t* lens, int32_t batch, int32_t t_max,
                          int32_t blank, int32_t* argmax_ids, int32_t* out_ids, int32_t* out_lens, float* logp,
                          int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream);

/* Polyphase sinc resampler (torchaudio.functional.resample semantics, used by FunASR's loader when the input rate differs from
 * the model's: funasr/utils/load_utils.py:176-178).  orig / nnew are the rates divided by their gcd, table [nnew, 2*width + orig]
 * is torchaudio's _get_sinc_resample_kernel (funasr_b200/resample.py builds it).  x [B, x_stride] with lens[B] valid samples ->
 * y [B, y_stride], first y_cap columns written (zero beyond each row's length), out_lens[b] = min(ceil(nnew*lens[b]/orig), y_cap). */
int fa_resample(const float* x, const int32_t* lens, int32_t batch, int64_t x_stride, const float* table, int32_t orig,
                int32_t nnew, int32_t width, float* y, int64_t y_stride, int32_t y_cap, int32_t* out_lens, fa_stream_t stream);

/* out[i, :] = table[ids[i], :]  (torch.nn.Embedding forward; CTTransformer.embed, ct_transformer/model.py:120).  ids are clamped to the table. */
int fa_embedding(const int32_t* ids, const float* table, int32_t dim, int32_t vocab, int64_t n, float* out, fa_stream_t stream);

/* Sample decode of the audio loader on the device (funasr/utils/load_utils.py:48-179; torchaudio.load(normalize=True) scaling): interleaved
 * PCM frames (device memory) -> mono fp32 [frames] in [-1, 1), channels averaged.  sample_format: 0 = f32, 1 = s16le, 2 = s24le packed,
 * 3 = s32le, 4 = u8.  The container (RIFF WAV header) is parsed on the host (funasr_b200/audio.py). */
int fa_pcm_decode(const void* pcm, int32_t sample_format, int32_t channels, int64_t frames, float* out, fa_stream_t stream);

/* Split fp32 [rows, cols] into three fp16 planes [3][rows][cols_pad] (hi, mid, lo; zero padded columns):
 * weight repack for the tcgen05 GEMM path (called once per weight after load_pretrained_model). */
int fa_split_planes(const float* src, int64_t ld_src, int64_t rows, int32_t cols, int32_t cols_pad,
                  void* planes, fa_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Handle-style offline recogniser (no Python, no torch) — the counterpart of FunASR's C++ runtime surface
 * runtime/onnxruntime/include/funasrruntime.h:100-116:
 *   fa_offline_init          <-> FunOfflineInit(model_path, thread_num, use_gpu, batch_size)            :100
 *   fa_offline_infer         <-> FunOfflineInferBuffer(handle, sz_buf, n_len, ...) ("pcm" = s16le)       :103-106
 *   fa_offline_result_ids    <-> FunASRGetResult(result, n_index)   (token ids; the tokenizer stays with the caller) :69
 *   fa_offline_result_audio_seconds <-> FunASRGetRetSnippetTime                                         :77
 *   fa_offline_free_result / fa_offline_uninit <-> FunASRFreeResult :75 / FunOfflineUninit              :116
 * model_file: flat tensor file written by funasr_b200/pack.py from a FunASR state_dict (same tensor names as model.pt) plus
 * am.mvn and the kaldi mel/window tables.  pcm_format: 0 = float32 in [-1,1], 1 = int16 little endian (converted on the
 * device, halves the H2D bytes).  bufs are HOST pointers, n_samples[i] >= 400.  Returns NULL on error
 * (fa_offline_last_error() says why); there is no CPU fallback.
 * ------------------------------------------------------------------------------------------- */
void* fa_offline_init(const char* model_file, int32_t device, int32_t gemm_mode);
void* fa_offline_infer(void* handle, const void* const* bufs, const int64_t* n_samples, int32_t batch, int32_t pcm_format);
/* Same with a hotword memory for a ContextualParaformer model file (FunOfflineInferBuffer's `hw_emb`, funasrruntime.h:104): hw_embed
 * is HOST memory [n_hotwords, 512], the rows CompileHotwordEmbedding produces (last entry = the <s> hotword).  Required when
 * fa_offline_is_contextual(handle), ignored otherwise. */
void* fa_offline_infer_hw(void* handle, const void* const* bufs, const int64_t* n_samples, int32_t batch, int32_t pcm_format,
                          const float* hw_embed, int32_t n_hotwords);
int32_t fa_offline_is_contextual(const void* handle);
/* Host copy of a tensor of the model file by its FunASR state_dict name (e.g. "bias_embed.weight" for the hotword encoder that
 * runs on the host); owned by the handle.  NULL if absent. */
const float* fa_offline_host_tensor(void* handle, const char* name, int64_t* numel);
int32_t fa_offline_result_count(const void* result);
const int32_t* fa_offline_result_ids(const void* result, int32_t index, int32_t* n_ids);
float fa_offline_result_audio_seconds(const void* result);
void fa_offline_free_result(void* result);
void fa_offline_uninit(void* handle);
const char* fa_offline_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FUNASR_B200_H_ */
