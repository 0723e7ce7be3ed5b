// Model-level C-ABI entry points: the kernel sequences of SANMEncoder.forward, CifPredictorV2.forward and
// ParaformerSANMDecoder.forward (+ greedy arg-max), stream-ordered over a caller-provided workspace.
#include "common.cuh"
#include <stdlib.h>
#include "kernels.h"
#include <string.h>
#include <math.h>
#include <map>
#include <mutex>
#include <utility>

namespace fa {

std::atomic<unsigned long long> g_launch_count{0};

// Side stream for the encoder's FSMN memory branch: it depends only on the QKV GEMM (like the attention kernel) and is HBM
// bound with a tiny footprint, so it runs concurrently with the latency-bound attention kernel and joins before the
// out-projection.  One lazily created (stream, fork event, join event) per device; FA_OVERLAP_FSMN=0 keeps everything on the
// caller's stream.
// The (side stream, fork event, join event) triple belongs to ONE caller stream on one device: two host threads that run
// encoders on different streams (two fa_offline handles, one worker thread per model) get different triples, so one thread's
// fork record can never be consumed by the other's side stream.  Calls that share a caller stream are ordered by that stream.
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream* side_stream(cudaStream_t caller) {
  static const bool enabled = [] { const char* e = getenv("FA_OVERLAP_FSMN"); return !(e && e[0] == '0'); }();
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, SideStream*> pool;
  std::lock_guard<std::mutex> lock(mu);
  SideStream*& slot = pool[std::make_pair(dev, caller)];
  if (!slot) {
    SideStream* s = new SideStream();
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) != cudaSuccess) {
      delete s;
      return nullptr;
    }
    slot = s;
  }
  return slot;
}

bool pdl_enabled() {
  static const bool v = [] { const char* e = getenv("FA_PDL"); return e && e[0] == '1'; }();   // opt-in: measured neutral (46.0 vs 46.1 ms)
  return v;
}

static int linear(const float* x, int64_t ldx, int64_t rows, const FaLinear& lin, int relu, const float* r1, int64_t ld1,
                  const float* r2, int64_t ld2, float* y, int64_t ldy, int mode, Arena* scratch, cudaStream_t st) {
  if (!lin.w) return FA_ERR_ARG;
  if (mode == FA_GEMM_F32_SIMT)
    return gemm_f32_launch(x, ldx, rows, lin.w, lin.out_f, lin.in_f, lin.b, relu, r1, ld1, r2, ld2, y, ldy, st);
  return gemm_tc_launch(x, ldx, rows, lin, relu, r1, ld1, r2, ld2, y, ldy, mode, scratch, st);
}

static inline int npl_for(int mode) { return mode == FA_GEMM_BF16X1 ? 1 : (mode == FA_GEMM_BF16X3 ? 2 : 3); }
static inline size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------------ encoder
static size_t enc_scratch_bytes(int batch, int t_max, int heads, int mode) {
  const int64_t M = (int64_t)batch * t_max;
  return max_sz(gemm_tc_scratch_bytes(M, 2048, mode), attention_tc_scratch_bytes(batch, heads, t_max, t_max, mode));
}
static size_t enc_plan(int batch, int t_max, int din, int mode) {
  const int64_t M = (int64_t)batch * t_max;
  ArenaSizer s;
  s.take(M * (size_t)din * 4);   // u
  s.take(M * 1536ull * 4);       // qkv
  s.take(M * 512ull * 4);        // mem
  s.take(M * 512ull * 4);        // ctx
  s.take(M * 512ull * 4);        // xa
  s.take(M * 512ull * 4);        // xb
  s.take(M * 2048ull * 4);       // h
  if (mode != FA_GEMM_F32_SIMT) {
    s.take(3ull * M * 512 * 2);    // ctx planes
    s.take(3ull * M * 2048 * 2);   // h planes
    s.take(3ull * M * 576 * 2);    // LN output planes
    s.take(2ull * M * 512 * 2);    // q planes (scaled)
    s.take(2ull * M * 512 * 2);    // k planes
    s.take(2ull * batch * 512 * (size_t)((t_max + 63) / 64 * 64) * 2);   // v planes, transposed per head
  }
  s.take(enc_scratch_bytes(batch, t_max, 4, mode));
  return s.off + 256;
}

}  // namespace fa

using namespace fa;

extern "C" size_t fa_sanm_encoder_workspace_bytes(int32_t batch, int32_t t_max, int32_t gemm_mode) {
  return enc_plan(batch, t_max, 560, gemm_mode);
}

extern "C" int fa_sanm_encoder_forward(const FaEncoder* enc, const float* feats, const int32_t* lens, int32_t batch,
                                       int32_t t_max, float* out, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                                       fa_stream_t stream) {
  if (!enc || !enc->layers || !feats || !lens || !out || batch <= 0 || t_max <= 0 || enc->n_layers < 1) return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = (int64_t)batch * t_max;
  const int D = enc->after_norm.n;
  const int din = enc->layers[0].norm1.n;
  const bool embed = enc->pe_inv_timescales != nullptr;   // false: plain stack over an existing [B,T,512] stream
  if (D != 512 || enc->heads * 128 != D || din > 560 || (!embed && din != D)) return FA_ERR_UNSUPPORTED;
  Arena a(workspace, ws_bytes);
  float* u = a.take<float>(M * (size_t)560);
  float* qkv = a.take<float>(M * 1536ull);
  float* mem = a.take<float>(M * 512ull);
  float* ctx = a.take<float>(M * 512ull);
  float* xa = a.take<float>(M * 512ull);
  float* xb = a.take<float>(M * 512ull);
  float* h = a.take<float>(M * 2048ull);
  const bool tc = gemm_mode != FA_GEMM_F32_SIMT;
  const int npl = npl_for(gemm_mode);
  __nv_bfloat16* ctx_planes = tc ? a.take<__nv_bfloat16>(3ull * M * 512) : nullptr;
  __nv_bfloat16* h_planes = tc ? a.take<__nv_bfloat16>(3ull * M * 2048) : nullptr;
  __nv_bfloat16* u_planes = tc ? a.take<__nv_bfloat16>(3ull * M * 576) : nullptr;
  const int t_pad = (t_max + 63) / 64 * 64;
  __nv_bfloat16* q_planes = tc ? a.take<__nv_bfloat16>(2ull * M * 512) : nullptr;
  __nv_bfloat16* k_planes = tc ? a.take<__nv_bfloat16>(2ull * M * 512) : nullptr;
  __nv_bfloat16* vt_planes = tc ? a.take<__nv_bfloat16>(2ull * batch * 512 * (size_t)t_pad) : nullptr;
  const size_t sb = enc_scratch_bytes(batch, t_max, enc->heads, gemm_mode);
  char* sp = a.take<char>(sb);
  if (!a.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);

  const float* x = embed ? nullptr : feats;  // residual stream (with PE input: undefined before layer 0, in_size != size)
  for (int l = 0; l < enc->n_layers; ++l) {
    const FaEncLayer& L = enc->layers[l];
    const int in = L.norm1.n;
    if (L.qkv.in_f != in || L.qkv.out_f != 3 * D || L.w1.in_f != D || L.w2.out_f != D || L.w2.in_f != L.w1.out_f) return FA_ERR_ARG;
    if (L.w1.out_f > 2048) return FA_ERR_UNSUPPORTED;      // the workspace plan sizes the FFN hidden slice for linear_units <= 2048
    // x = x*sqrt(D) + PE is folded into the first LayerNorm (encoder.py:409,428)
    // tensor-core path: LayerNorm writes the bf16 planes the QKV GEMM consumes (no fp32 round trip, no split pass)
    if (l > 0 && in != D) return FA_ERR_UNSUPPORTED;
    const bool first = embed && l == 0;
    FA_RETURN_IF_ERR(layernorm_launch(first ? feats : x, M, L.norm1, tc ? nullptr : u, first ? enc->pe_inv_timescales : nullptr,
                                      first ? sqrtf((float)D) : 1.f, t_max, st, u_planes, npl, L.qkv.in_pad));
    if (tc) {
      // QKV GEMM epilogue emits the attention operands directly: q (x d_k^-0.5) / k as bf16 planes, v transposed per head
      // as bf16 planes plus fp32 v (the only fp32 columns written) for the FSMN branch
      AttnSinks sk;
      sk.q0 = 0; sk.k0 = D; sk.v0 = 2 * D; sk.width = D; sk.npl = npl < 2 ? npl : 2; sk.t_rows = t_max; sk.t_pad = t_pad;
      sk.qscale = (float)(1.0 / sqrt(128.0)); sk.q_planes = q_planes; sk.k_planes = k_planes; sk.vt_planes = vt_planes;
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(u_planes, M, L.qkv, 0, nullptr, 0, nullptr, 0, qkv, 3 * D, nullptr, 0, gemm_mode, st, &sk));
    } else {
      FA_RETURN_IF_ERR(linear(u, in, M, L.qkv, 0, nullptr, 0, nullptr, 0, qkv, 3 * D, gemm_mode, &scratch, st));
    }
    SideStream* side = tc ? side_stream(st) : nullptr;
    if (side) {                                     // FSMN memory branch runs beside the attention kernel (both need only QKV)
      FA_CUDA_OK(cudaEventRecord(side->fork, st));
      FA_CUDA_OK(cudaStreamWaitEvent(side->st, side->fork, 0));
      FA_RETURN_IF_ERR(fsmn_launch(qkv + 2 * D, 3 * D, lens, batch, t_max, D, L.fsmn_w, enc->fsmn_k, nullptr, 0, mem, D, side->st));
      FA_CUDA_OK(cudaEventRecord(side->join, side->st));
    } else {
      FA_RETURN_IF_ERR(fsmn_launch(qkv + 2 * D, 3 * D, lens, batch, t_max, D, L.fsmn_w, enc->fsmn_k, nullptr, 0, mem, D, st));
    }
    // x2 = (residual if in_size == size) + (linear_out(ctx) + fsmn_memory)     encoder.py:120-137, attention.py:327
    float* x2 = (x == xa) ? xb : xa;
    const float* res = (in == D && x != nullptr) ? x : nullptr;
    float* x3 = (x2 == xa) ? xb : xa;
    if (!tc) {
      FA_RETURN_IF_ERR(attention_f32_launch(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, lens, batch, enc->heads, t_max,
                                            t_max, ctx, D, st));
      FA_RETURN_IF_ERR(linear(ctx, D, M, L.out, 0, mem, D, res, D, x2, D, gemm_mode, &scratch, st));
      FA_RETURN_IF_ERR(layernorm_launch(x2, M, L.norm2, u, nullptr, 1.f, t_max, st));
      FA_RETURN_IF_ERR(linear(u, D, M, L.w1, 1, nullptr, 0, nullptr, 0, h, L.w1.out_f, gemm_mode, &scratch, st));
      FA_RETURN_IF_ERR(linear(h, L.w1.out_f, M, L.w2, 0, x2, D, nullptr, 0, x3, D, gemm_mode, &scratch, st));
    } else {
      // tensor-core path: attention emits the context as bf16 planes (A operand of linear_out); FFN w_1 emits its
      // ReLU output as planes for w_2 — neither intermediate makes an fp32 round trip through HBM
      FA_RETURN_IF_ERR(attention_tc_planes_launch(q_planes, k_planes, vt_planes, lens, batch, enc->heads, t_max, t_max, nullptr, 0,
                                                  ctx_planes, D, npl, gemm_mode, st));
      if (side) FA_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));          // join: linear_out adds the FSMN memory
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(ctx_planes, M, L.out, 0, mem, D, res, D, x2, D, nullptr, 0, gemm_mode, st));
      if (L.w1.out_f != L.w2.in_pad || L.w1.in_pad != D) return FA_ERR_UNSUPPORTED;
      FA_RETURN_IF_ERR(layernorm_launch(x2, M, L.norm2, nullptr, nullptr, 1.f, t_max, st, u_planes, npl, D));
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(u_planes, M, L.w1, 1, nullptr, 0, nullptr, 0, nullptr, 0, h_planes, L.w1.out_f, gemm_mode, st));
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(h_planes, M, L.w2, 0, x2, D, nullptr, 0, x3, D, nullptr, 0, gemm_mode, st));
    }
    x = x3;
  }
  return layernorm_launch(x, M, enc->after_norm, out, nullptr, 1.f, t_max, st);
}

// ---------------------------------------------------------------------------------------------- predictor
extern "C" size_t fa_cif_predictor_workspace_bytes(int32_t batch, int32_t t_max, int32_t gemm_mode) {
  const int64_t M = (int64_t)batch * t_max;
  ArenaSizer s;
  s.take(M * 1536ull * 4);
  s.take(M * 512ull * 4);
  s.take(M * 4ull);
  s.take(gemm_tc_scratch_bytes(M, 1536, gemm_mode));
  return s.off + 256;
}

extern "C" int fa_cif_predictor_forward(const FaPredictor* pred, const float* enc, const int32_t* lens, int32_t batch,
                                        int32_t t_max, float* acoustic, int32_t n_cap, int32_t* token_num,
                                        float* alphas, float* peaks, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                                        fa_stream_t stream) {
  if (!pred || !enc || !lens || !acoustic || !token_num || !alphas || !peaks || batch <= 0 || t_max <= 0 || n_cap <= 0)
    return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 512;
  if (pred->conv.out_f != D || pred->conv.in_f != 3 * D) return FA_ERR_UNSUPPORTED;
  const int64_t M = (int64_t)batch * t_max;
  Arena a(workspace, ws_bytes);
  float* xc = a.take<float>(M * 1536ull);
  float* c = a.take<float>(M * 512ull);
  float* alpha_rows = a.take<float>(M);
  const size_t sb = gemm_tc_scratch_bytes(M, 1536, gemm_mode);
  char* sp = a.take<char>(sb);
  if (!a.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);
  FA_RETURN_IF_ERR(cif_im2col_launch(enc, M, t_max, D, xc, st));
  FA_RETURN_IF_ERR(linear(xc, 3 * D, M, pred->conv, 1, nullptr, 0, nullptr, 0, c, D, gemm_mode, &scratch, st));
  FA_RETURN_IF_ERR(cif_alpha_launch(c, D, pred->out_w, pred->out_b, lens, t_max, M, pred->smooth_factor,
                                    pred->noise_threshold, alpha_rows, st));
  FA_CUDA_OK(cudaMemsetAsync(acoustic, 0, (size_t)batch * n_cap * D * sizeof(float), st));
  if (pred->cif_variant == 1)     // CifPredictorV3 (BiCifParaformer): sequential fp32 `cif`
    return cif_fire_loop_launch(enc, alpha_rows, lens, batch, t_max, D, pred->tail_threshold, pred->threshold, acoustic, n_cap,
                                token_num, alphas, peaks, st);
  if (pred->cif_variant != 0) return FA_ERR_ARG;
  return cif_fire_launch(enc, alpha_rows, lens, batch, t_max, D, pred->tail_threshold, acoustic, n_cap, token_num, alphas,
                         peaks, st);
}

// CifPredictorV3.get_upsample_timestamp after the BLSTM (bicif_paraformer/cif_predictor.py:331-352)
extern "C" int fa_cif_upsample_alphas(const float* feat, int32_t dz, const float* w, const float* b, const int32_t* lens_up,
                                      const int32_t* token_num, int32_t batch, int32_t t_up, float smooth2, float noise2,
                                      float threshold, float* us_alphas, float* us_peaks, fa_stream_t stream) {
  if (!feat || !w || !b || !lens_up || !token_num || !us_alphas || !us_peaks || batch <= 0 || t_up <= 0 || dz <= 0 || (dz & 3))
    return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  FA_RETURN_IF_ERR(cif_alpha_launch(feat, dz, w, b, lens_up, t_up, (int64_t)batch * t_up, smooth2, noise2, us_alphas, st));
  return cif_upsample_scan_launch(us_alphas, token_num, batch, t_up, (float)((double)threshold - 1e-4), us_peaks, st);
}

// ------------------------------------------------------------------------------------------------ decoder
static size_t dec_scratch_bytes(int batch, int t_max, int n_max, int mode) {
  const int64_t Mq = (int64_t)batch * n_max, Mk = (int64_t)batch * t_max;
  return max_sz(gemm_tc_scratch_bytes(Mq > Mk ? Mq : Mk, 2048, mode), attention_tc_scratch_bytes(batch, 4, n_max, t_max, mode));
}
// hotword region of the contextual bias decoder: k|v rows of the (shared) hotword memory + GEMM / attention scratch
static size_t hw_region_bytes(int batch, int n_max, int nh, int mode) {
  if (nh <= 0) return 0;
  const int64_t Mq = (int64_t)batch * n_max;
  ArenaSizer s;
  s.take((size_t)nh * 1024 * 4);
  s.take(max_sz(gemm_tc_scratch_bytes(Mq > nh ? Mq : nh, 1024, mode), attention_tc_scratch_bytes(batch, 4, n_max, nh, mode)));
  return s.off + 256;
}
static size_t dec_plan(int batch, int t_max, int n_max, int vocab, int mode, size_t dec_scratch, int n_hotwords) {
  const int64_t Mq = (int64_t)batch * n_max, Mk = (int64_t)batch * t_max;
  ArenaSizer s;
  s.take(Mq * 512ull * 4);   // ya
  s.take(Mq * 512ull * 4);   // yb
  s.take(Mq * 512ull * 4);   // t1
  s.take(Mq * 2048ull * 4);  // hq
  s.take(Mq * 512ull * 4);   // f
  s.take(Mq * 512ull * 4);   // qd
  s.take(Mq * 512ull * 4);   // ctx
  s.take(Mk * 1024ull * 4);  // kv
  s.take(Mq * (size_t)vocab * 4);  // logits (used when the caller passes none)
  s.take(Mq * 1024ull * 4);        // [x_src_attn ; cx] of the contextual decoder
  s.take(hw_region_bytes(batch, n_max, n_hotwords, mode));
  if (mode != FA_GEMM_F32_SIMT) {
    s.take(3ull * Mq * 512 * 2);   // ctx planes
    s.take(3ull * Mk * 512 * 2);   // enc planes (split once, reused by the 16 kv GEMMs)
    s.take(3ull * Mq * 512 * 2);   // LN output planes
    s.take(3ull * Mq * 2048 * 2);  // FFN hidden planes
    s.take(2ull * Mq * 512 * 2);   // q planes
    s.take(2ull * Mk * 512 * 2);   // k planes
    s.take(2ull * batch * 512 * (size_t)((t_max + 63) / 64 * 64) * 2);   // v planes, transposed per head
  }
  s.take(dec_scratch);
  return s.off + 256;
}

extern "C" size_t fa_paraformer_decoder_workspace_bytes(int32_t batch, int32_t t_max, int32_t n_max, int32_t vocab,
                                                        int32_t gemm_mode) {
  return dec_plan(batch, t_max, n_max, vocab, gemm_mode, dec_scratch_bytes(batch, t_max, n_max, gemm_mode), t_max);
}
extern "C" size_t fa_paraformer_decoder_workspace_bytes_hw(int32_t batch, int32_t t_max, int32_t n_max, int32_t vocab,
                                                           int32_t gemm_mode, int32_t n_hotwords) {
  return dec_plan(batch, t_max, n_max, vocab, gemm_mode, dec_scratch_bytes(batch, t_max, n_max, gemm_mode), n_hotwords);
}

static int dec_ffn(const FaDecLayer& L, const float* y, int64_t Mq, float* t1, float* hq, float* f, int mode,
                   Arena* scratch, cudaStream_t st, __nv_bfloat16* t1_planes, __nv_bfloat16* hq_planes) {
  // f = w_2( LN_2048( relu( w_1( LN1(y) ) ) ) )   decoder.py:97-100, sanm/positionwise_feed_forward.py:33
  if (L.ffn_w1.in_f != 512 || L.ffn_w2.out_f != 512 || L.ffn_w2.in_f != L.ffn_w1.out_f || L.ffn_norm.n != L.ffn_w1.out_f) return FA_ERR_ARG;
  if (L.ffn_w1.out_f > 2048) return FA_ERR_UNSUPPORTED;    // dec_plan sizes hq / hq_planes for linear_units <= 2048
  if (mode != FA_GEMM_F32_SIMT) {
    const int npl = npl_for(mode);
    if (L.ffn_w1.in_pad != 512 || L.ffn_w2.in_pad != L.ffn_w1.out_f) return FA_ERR_UNSUPPORTED;
    FA_RETURN_IF_ERR(layernorm_launch(y, Mq, L.norm1, nullptr, nullptr, 1.f, 1, st, t1_planes, npl, 512));
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(t1_planes, Mq, L.ffn_w1, 1, nullptr, 0, nullptr, 0, hq, L.ffn_w1.out_f, nullptr, 0, mode, st));
    FA_RETURN_IF_ERR(layernorm_launch(hq, Mq, L.ffn_norm, nullptr, nullptr, 1.f, 1, st, hq_planes, npl, L.ffn_w1.out_f));
    return gemm_tc_planes_launch(hq_planes, Mq, L.ffn_w2, 0, nullptr, 0, nullptr, 0, f, 512, nullptr, 0, mode, st);
  }
  FA_RETURN_IF_ERR(layernorm_launch(y, Mq, L.norm1, t1, nullptr, 1.f, 1, st));
  FA_RETURN_IF_ERR(linear(t1, 512, Mq, L.ffn_w1, 1, nullptr, 0, nullptr, 0, hq, L.ffn_w1.out_f, mode, scratch, st));
  FA_RETURN_IF_ERR(layernorm_launch(hq, Mq, L.ffn_norm, hq, nullptr, 1.f, 1, st));
  return linear(hq, L.ffn_w1.out_f, Mq, L.ffn_w2, 0, nullptr, 0, nullptr, 0, f, 512, mode, scratch, st);
}

extern "C" int fa_paraformer_decoder_forward(const FaDecoder* dec, const float* enc, const int32_t* enc_lens,
                                             int32_t batch, int32_t t_max, const float* acoustic,
                                             int64_t ld_acoustic_rows, const int32_t* tok_lens, int32_t n_max,
                                             int32_t* argmax_ids, float* argmax_logp, float* logits, int32_t log_softmax,
                                             int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream) {
  if (!dec || !enc || !enc_lens || !acoustic || !tok_lens || !argmax_ids || !argmax_logp || batch <= 0 || t_max <= 0 ||
      n_max <= 0 || ld_acoustic_rows < n_max)
    return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 512;
  if (dec->after_norm.n != D || dec->heads * 128 != D) return FA_ERR_UNSUPPORTED;
  const int64_t Mq = (int64_t)batch * n_max, Mk = (int64_t)batch * t_max;
  const int V = dec->vocab;
  Arena a(workspace, ws_bytes);
  float* ya = a.take<float>(Mq * 512ull);
  float* yb = a.take<float>(Mq * 512ull);
  float* t1 = a.take<float>(Mq * 512ull);
  float* hq = a.take<float>(Mq * 2048ull);
  float* f = a.take<float>(Mq * 512ull);
  float* qd = a.take<float>(Mq * 512ull);
  float* ctx = a.take<float>(Mq * 512ull);
  float* kv = a.take<float>(Mk * 1024ull);
  float* lg = a.take<float>(Mq * (size_t)V);
  float* cat = a.take<float>(Mq * 1024ull);
  const size_t hwb = hw_region_bytes(batch, n_max, dec->has_bias ? dec->n_hotwords : 0, gemm_mode);
  char* hw_region = a.take<char>(hwb);
  const bool tc = gemm_mode != FA_GEMM_F32_SIMT;
  const int npl = npl_for(gemm_mode);
  __nv_bfloat16* ctx_planes = tc ? a.take<__nv_bfloat16>(3ull * Mq * 512) : nullptr;
  __nv_bfloat16* enc_planes = tc ? a.take<__nv_bfloat16>(3ull * Mk * 512) : nullptr;
  __nv_bfloat16* t1_planes = tc ? a.take<__nv_bfloat16>(3ull * Mq * 512) : nullptr;
  __nv_bfloat16* hq_planes = tc ? a.take<__nv_bfloat16>(3ull * Mq * 2048) : nullptr;
  const int t_pad = (t_max + 63) / 64 * 64;
  __nv_bfloat16* q_planes = tc ? a.take<__nv_bfloat16>(2ull * Mq * 512) : nullptr;
  __nv_bfloat16* k_planes = tc ? a.take<__nv_bfloat16>(2ull * Mk * 512) : nullptr;
  __nv_bfloat16* vt_planes = tc ? a.take<__nv_bfloat16>(2ull * batch * 512 * (size_t)t_pad) : nullptr;
  const size_t sb = dec_scratch_bytes(batch, t_max, n_max, gemm_mode);
  char* sp = a.take<char>(sb);
  if (!a.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);
  if (logits) lg = logits;
  if (tc) FA_RETURN_IF_ERR(split_rows_launch(enc, D, Mk, D, D, npl, enc_planes, st));   // memory is layer-invariant

  // tgt = acoustic[:, :n_max]  (decoder.py:424)
  FA_CUDA_OK(cudaMemcpy2DAsync(ya, (size_t)n_max * D * 4, acoustic, (size_t)ld_acoustic_rows * D * 4, (size_t)n_max * D * 4,
                               batch, cudaMemcpyDeviceToDevice, st));
  fa::count_launch();
  float* y = ya;
  // One attention decoder layer.  x_self_out receives `residual + fsmn(...)` (x_self_attn); if src_out != nullptr the
  // cross-attention output is written there WITHOUT the residual (x_src_attn, leading dim ld_src) and *y_next is not
  // produced — the ContextualDecoderLayer contract (contextual_paraformer/decoder.py:60-100).
  auto attention_layer = [&](const FaDecLayer& L, float* yin, float** x_self_out, float* src_out, int64_t ld_src, float** y_next) -> int {
    FA_RETURN_IF_ERR(dec_ffn(L, yin, Mq, t1, hq, f, gemm_mode, &scratch, st, t1_planes, hq_planes));
    // x = residual + fsmn(LN2(f), tgt_mask)     decoder.py:103-107
    FA_RETURN_IF_ERR(layernorm_launch(f, Mq, L.norm2, t1, nullptr, 1.f, 1, st));
    float* x2 = (yin == ya) ? yb : ya;
    FA_RETURN_IF_ERR(fsmn_launch(t1, D, tok_lens, batch, n_max, D, L.fsmn_w, dec->fsmn_k, yin, D, x2, D, st));
    // x = residual + src_attn(LN3(x), memory)    decoder.py:109-118, attention.py:796-813
    if (tc) {
      FA_RETURN_IF_ERR(layernorm_launch(x2, Mq, L.norm3, nullptr, nullptr, 1.f, 1, st, t1_planes, npl, D));
      AttnSinks sq;                       // q -> scaled bf16 planes only (no fp32 round trip)
      sq.q0 = 0; sq.width = D; sq.npl = npl < 2 ? npl : 2; sq.t_rows = n_max; sq.qscale = (float)(1.0 / sqrt(128.0)); sq.q_planes = q_planes;
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(t1_planes, Mq, L.q, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, gemm_mode, st, &sq));
    } else {
      FA_RETURN_IF_ERR(layernorm_launch(x2, Mq, L.norm3, t1, nullptr, 1.f, 1, st));
      FA_RETURN_IF_ERR(linear(t1, D, Mq, L.q, 0, nullptr, 0, nullptr, 0, qd, D, gemm_mode, &scratch, st));
    }
    float* y2 = (x2 == ya) ? yb : ya;
    float* dst = src_out ? src_out : y2;
    const int64_t ldd = src_out ? ld_src : D;
    const float* res = src_out ? nullptr : x2;
    if (!tc) {
      FA_RETURN_IF_ERR(linear(enc, D, Mk, L.kv, 0, nullptr, 0, nullptr, 0, kv, 2 * D, gemm_mode, &scratch, st));
      FA_RETURN_IF_ERR(attention_f32_launch(qd, D, kv, 2 * D, kv + D, 2 * D, enc_lens, batch, dec->heads, n_max, t_max, ctx,
                                            D, st));
      FA_RETURN_IF_ERR(linear(ctx, D, Mq, L.out, 0, res, D, nullptr, 0, dst, ldd, gemm_mode, &scratch, st));
    } else {
      AttnSinks skv;                      // k -> planes, v -> transposed planes; nothing in fp32
      skv.k0 = 0; skv.v0 = D; skv.width = D; skv.npl = npl < 2 ? npl : 2; skv.t_rows = t_max; skv.t_pad = t_pad;
      skv.k_planes = k_planes; skv.vt_planes = vt_planes;
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(enc_planes, Mk, L.kv, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, gemm_mode, st, &skv));
      FA_RETURN_IF_ERR(attention_tc_planes_launch(q_planes, k_planes, vt_planes, enc_lens, batch, dec->heads, n_max, t_max, nullptr, 0,
                                                  ctx_planes, D, npl, gemm_mode, st));
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(ctx_planes, Mq, L.out, 0, res, D, nullptr, 0, dst, ldd, nullptr, 0, gemm_mode, st));
    }
    *x_self_out = x2;
    if (y_next) *y_next = y2;
    return FA_OK;
  };
  for (int l = 0; l < dec->n_layers; ++l) {
    float* xs = nullptr;
    FA_RETURN_IF_ERR(attention_layer(dec->layers[l], y, &xs, nullptr, 0, &y));
  }
  if (dec->has_bias) {
    // ContextualParaformerDecoder.forward decoder.py:325-340
    const int nh = dec->n_hotwords;
    if (!dec->hw_embed || !dec->hw_lens || nh <= 0 || dec->clas_scale != 1.0f) return FA_ERR_UNSUPPORTED;
    float* x_self = nullptr;
    FA_RETURN_IF_ERR(attention_layer(dec->bias_last, y, &x_self, cat, 2 * D, nullptr));      // cat[:, :512] = x_src_attn
    // bias decoder: cross attention of LN3(x_self_attn) over the hotword memory (identical for every utterance)
    FA_RETURN_IF_ERR(layernorm_launch(x_self, Mq, dec->bias_norm3, t1, nullptr, 1.f, 1, st));
    FA_RETURN_IF_ERR(linear(t1, D, Mq, dec->bias_q, 0, nullptr, 0, nullptr, 0, qd, D, gemm_mode, &scratch, st));
    // the hotword k | v rows [nh, 1024] are the same for every utterance: one copy, attended with kv_shared (no per-utterance
    // replication, so the hotword count is independent of t_max)
    Arena a2(hw_region, hwb);
    float* kvh = a2.take<float>((size_t)nh * 1024);
    const size_t sb2 = max_sz(gemm_tc_scratch_bytes(Mq > nh ? Mq : nh, 1024, gemm_mode), attention_tc_scratch_bytes(batch, 4, n_max, nh, gemm_mode));
    char* sp2 = a2.take<char>(sb2);
    if (!a2.ok()) return FA_ERR_WORKSPACE;
    Arena scratch2(sp2, sb2);
    FA_RETURN_IF_ERR(linear(dec->hw_embed, D, nh, dec->bias_kv, 0, nullptr, 0, nullptr, 0, kvh, 2 * D, gemm_mode, &scratch2, st));
    if (!tc) {
      FA_RETURN_IF_ERR(attention_f32_launch(qd, D, kvh, 2 * D, kvh + D, 2 * D, dec->hw_lens, batch, dec->heads, n_max, nh, ctx, D, st, 1));
    } else {
      FA_RETURN_IF_ERR(attention_tc_launch(qd, D, kvh, 2 * D, kvh + D, 2 * D, dec->hw_lens, batch, dec->heads, n_max, nh, ctx, D,
                                           nullptr, 0, 0, gemm_mode, &scratch2, st, 1));
    }
    FA_RETURN_IF_ERR(linear(ctx, D, Mq, dec->bias_out, 0, nullptr, 0, nullptr, 0, cat + D, 2 * D, gemm_mode, &scratch, st));   // cat[:, 512:] = cx
    float* y2 = (x_self == ya) ? yb : ya;
    FA_RETURN_IF_ERR(linear(cat, 2 * D, Mq, dec->bias_output, 0, x_self, D, nullptr, 0, y2, D, gemm_mode, &scratch, st));
    y = y2;
  }
  // decoders3: FFN only, no residual (decoder.py:97-102,121); after_norm; output_layer
  FA_RETURN_IF_ERR(dec_ffn(dec->last, y, Mq, t1, hq, f, gemm_mode, &scratch, st, t1_planes, hq_planes));
  if (tc) {
    FA_RETURN_IF_ERR(layernorm_launch(f, Mq, dec->after_norm, nullptr, nullptr, 1.f, 1, st, t1_planes, npl, D));
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(t1_planes, Mq, dec->output, 0, nullptr, 0, nullptr, 0, lg, V, nullptr, 0, gemm_mode, st));
  } else {
    FA_RETURN_IF_ERR(layernorm_launch(f, Mq, dec->after_norm, t1, nullptr, 1.f, 1, st));
    FA_RETURN_IF_ERR(linear(t1, D, Mq, dec->output, 0, nullptr, 0, nullptr, 0, lg, V, gemm_mode, &scratch, st));
  }
  return argmax_lse_launch(lg, Mq, V, V, argmax_ids, argmax_logp, (logits && log_softmax) ? 1 : 0, st);
}

// ------------------------------------------------------------------------------------------ CTC greedy head
extern "C" size_t fa_ctc_greedy_workspace_bytes(int32_t batch, int32_t t_max, int32_t vocab, int32_t gemm_mode) {
  const int64_t M = (int64_t)batch * t_max;
  ArenaSizer s;
  s.take(M * (size_t)vocab * 4);
  s.take(M * 4ull);
  s.take(gemm_tc_scratch_bytes(M, 512, gemm_mode));
  return s.off + 256;
}

extern "C" int fa_ctc_greedy_forward(const FaLinear* ctc_lo, const float* enc, const int32_t* lens, int32_t batch,
                                     int32_t t_max, int32_t blank, int32_t* argmax_ids, int32_t* out_ids, int32_t* out_lens,
                                     float* logp, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream) {
  if (!ctc_lo || !enc || !lens || !argmax_ids || !out_ids || !out_lens || batch <= 0 || t_max <= 0) return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = (int64_t)batch * t_max;
  const int V = ctc_lo->out_f;
  Arena a(workspace, ws_bytes);
  float* lg = a.take<float>(M * (size_t)V);
  float* best = a.take<float>(M);
  const size_t sb = gemm_tc_scratch_bytes(M, 512, gemm_mode);
  char* sp = a.take<char>(sb);
  if (!a.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);
  if (logp) lg = logp;
  FA_RETURN_IF_ERR(linear(enc, ctc_lo->in_f, M, *ctc_lo, 0, nullptr, 0, nullptr, 0, lg, V, gemm_mode, &scratch, st));
  FA_RETURN_IF_ERR(argmax_lse_launch(lg, M, V, V, argmax_ids, best, logp ? 1 : 0, st));
  return ctc_filter_launch(argmax_ids, lens, batch, t_max, blank, out_ids, out_lens, st);
}

// ------------------------------------------------------------------------------------------ op-level + info
extern "C" int fa_linear(const float* x, int64_t ldx, int64_t rows, const FaLinear* lin, int32_t relu, const float* res1,
                         int64_t ld_res1, const float* res2, int64_t ld_res2, float* y, int64_t ldy, int32_t gemm_mode,
                         void* workspace, size_t ws_bytes, fa_stream_t stream) {
  if (!lin || !x || !y) return FA_ERR_ARG;
  Arena scratch(workspace, ws_bytes);
  return linear(x, ldx, rows, *lin, relu, res1, ld_res1, res2, ld_res2, y, ldy, gemm_mode, &scratch, (cudaStream_t)stream);
}

extern "C" int fa_split_rows(const float* x, int64_t ldx, int64_t rows, int32_t cols, int32_t cols_pad, int32_t nplanes, void* planes,
                             fa_stream_t stream) {
  if (!x || !planes || nplanes < 1 || nplanes > 3) return FA_ERR_ARG;
  return split_rows_launch(x, ldx, rows, cols, cols_pad, nplanes, reinterpret_cast<__nv_bfloat16*>(planes), (cudaStream_t)stream);
}

extern "C" int fa_linear_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, const float* res1,
                                int64_t ld_res1, const float* res2, int64_t ld_res2, float* y, int64_t ldy, int32_t gemm_mode,
                                fa_stream_t stream) {
  if (!a_planes || !lin || !y || gemm_mode == FA_GEMM_F32_SIMT) return FA_ERR_ARG;
  return gemm_tc_planes_launch(reinterpret_cast<const __nv_bfloat16*>(a_planes), rows, *lin, relu, res1, ld_res1, res2, ld_res2, y, ldy,
                               nullptr, 0, gemm_mode, (cudaStream_t)stream);
}

extern "C" int fa_linear_planes_to_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, void* out_planes,
                                          int64_t ld_out, int32_t gemm_mode, fa_stream_t stream) {
  if (!a_planes || !lin || !out_planes || gemm_mode == FA_GEMM_F32_SIMT) return FA_ERR_ARG;
  return gemm_tc_planes_launch(reinterpret_cast<const __nv_bfloat16*>(a_planes), rows, *lin, relu, nullptr, 0, nullptr, 0, nullptr, 0,
                               reinterpret_cast<__nv_bfloat16*>(out_planes), ld_out, gemm_mode, (cudaStream_t)stream);
}

extern "C" const char* fa_version(void) { return "funasr_b200 0.1.0 (sm_100a)"; }
extern "C" uint64_t fa_launch_count(void) { return (uint64_t)fa::g_launch_count.load(); }
extern "C" const char* fa_status_string(int status) {
  switch (status) {
    case FA_OK: return "ok";
    case FA_ERR_ARG: return "bad argument";
    case FA_ERR_CUDA: return "CUDA error";
    case FA_ERR_WORKSPACE: return "workspace too small";
    case FA_ERR_UNSUPPORTED: return "unsupported shape";
    default: return "unknown";
  }
}
