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

static inline int npl_for(int mode) { return mode == FA_GEMM_F16X1 ? 1 : (mode == FA_GEMM_F16X3 ? 2 : 3); }
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
  const int hd = enc->heads > 0 ? D / enc->heads : 0;
  // the tcgen05 kernels are built for the Paraformer / SenseVoice shape (d = 512, 4 x 128); the fp32 path also runs the small
  // SAN-M stacks around the hot path (CT-Transformer punctuation: d = 256, 8 x 32)
  if (D < 64 || D > 512 || (D & 15) || enc->heads < 1 || enc->heads * hd != D || hd < 32 || hd > 128 || (hd & 31) || din > 560 || (din & 15) ||
      (!embed && din != D))
    return FA_ERR_UNSUPPORTED;
  if (gemm_mode != FA_GEMM_F32_SIMT && (D != 512 || hd != 128)) return FA_ERR_UNSUPPORTED;
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
  plane_t* ctx_planes = tc ? a.take<plane_t>(3ull * M * 512) : nullptr;
  plane_t* h_planes = tc ? a.take<plane_t>(3ull * M * 2048) : nullptr;
  plane_t* u_planes = tc ? a.take<plane_t>(3ull * M * 576) : nullptr;
  const int t_pad = (t_max + 63) / 64 * 64;
  plane_t* q_planes = tc ? a.take<plane_t>(2ull * M * 512) : nullptr;
  plane_t* k_planes = tc ? a.take<plane_t>(2ull * M * 512) : nullptr;
  plane_t* vt_planes = tc ? a.take<plane_t>(2ull * batch * 512 * (size_t)t_pad) : nullptr;
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
    // tensor-core path: LayerNorm writes the fp16 planes the QKV GEMM consumes (no fp32 round trip, no split pass)
    if (l > 0 && in != D) return FA_ERR_UNSUPPORTED;
    const bool first = embed && l == 0;
    // a first layer with in_size == size keeps its residual (encoder.py:120-126): the embedded rows x*sqrt(d) + PE are then needed
    // beside their LayerNorm (CT-Transformer: 256 -> 256; Paraformer / SenseVoice: 560 -> 512, no residual)
    float* emb = (first && in == D) ? xa : nullptr;
    FA_RETURN_IF_ERR(layernorm_launch(first ? feats : x, M, L.norm1, tc ? nullptr : u, first ? enc->pe_inv_timescales : nullptr,
                                      first ? sqrtf((float)D) : 1.f, t_max, st, u_planes, npl, L.qkv.in_pad, emb));
    if (emb) x = emb;
    if (tc) {
      // QKV GEMM epilogue emits the attention operands directly: q (x d_k^-0.5) / k as fp16 planes, v transposed per head
      // as fp16 planes plus fp32 v (the only fp32 columns written) for the FSMN branch
      AttnSinks sk;
      sk.q0 = 0; sk.k0 = D; sk.v0 = 2 * D; sk.width = D; sk.npl = npl < 2 ? npl : 2; sk.t_rows = t_max; sk.t_pad = t_pad;
      sk.qscale = (float)(1.0 / sqrt(128.0)); sk.q_planes = q_planes; sk.k_planes = k_planes; sk.vt_planes = vt_planes;
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(u_planes, M, L.qkv, 0, nullptr, 0, nullptr, 0, qkv, 3 * D, nullptr, 0, gemm_mode, st, &sk));
    } else {
      FA_RETURN_IF_ERR(linear(u, in, M, L.qkv, 0, nullptr, 0, nullptr, 0, qkv, 3 * D, gemm_mode, &scratch, st));
    }
    SideStream* side = tc ? side_stream(st) : nullptr;
    // x2 = (residual if in_size == size) + (linear_out(ctx) + fsmn_memory)     encoder.py:120-137, attention.py:327
    float* x2 = (x == xa) ? xb : xa;
    const float* res = (in == D && x != nullptr) ? x : nullptr;
    float* x3 = (x2 == xa) ? xb : xa;
    // tensor-core path: the FSMN kernel (HBM bound, on the side stream beside the latency-bound attention kernel) also adds the
    // layer's residual, so the out-projection epilogue reads ONE fp32 stream instead of two (it was bound by those reads:
    // 198 MB in 69 us, tensor pipe 43 %).  fp32 path: the reference's own association ((att + mem) + residual) is kept.
    const float* fsmn_res = tc ? res : nullptr;
    if (side) {                                     // FSMN memory branch runs beside the attention kernel (both need only QKV)
      FA_CUDA_OK(cudaEventRecord(side->fork, st));
      FA_CUDA_OK(cudaStreamWaitEvent(side->st, side->fork, 0));
      FA_RETURN_IF_ERR(fsmn_launch(qkv + 2 * D, 3 * D, lens, batch, t_max, D, L.fsmn_w, enc->fsmn_k, fsmn_res, D, mem, D, side->st));
      FA_CUDA_OK(cudaEventRecord(side->join, side->st));
    } else {
      FA_RETURN_IF_ERR(fsmn_launch(qkv + 2 * D, 3 * D, lens, batch, t_max, D, L.fsmn_w, enc->fsmn_k, fsmn_res, D, mem, D, st));
    }
    if (!tc) {
      if (hd == 128) {
        FA_RETURN_IF_ERR(attention_f32_launch(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, lens, batch, enc->heads, t_max,
                                              t_max, ctx, D, st));
      } else {
        FA_RETURN_IF_ERR(attention_small_launch(qkv, 3 * D, qkv + D, 3 * D, qkv + 2 * D, 3 * D, lens, batch, enc->heads, hd, t_max, t_max, ctx, D, st));
      }
      FA_RETURN_IF_ERR(linear(ctx, D, M, L.out, 0, mem, D, res, D, x2, D, gemm_mode, &scratch, st));
      FA_RETURN_IF_ERR(layernorm_launch(x2, M, L.norm2, u, nullptr, 1.f, t_max, st));
      FA_RETURN_IF_ERR(linear(u, D, M, L.w1, 1, nullptr, 0, nullptr, 0, h, L.w1.out_f, gemm_mode, &scratch, st));
      FA_RETURN_IF_ERR(linear(h, L.w1.out_f, M, L.w2, 0, x2, D, nullptr, 0, x3, D, gemm_mode, &scratch, st));
    } else {
      // tensor-core path: attention emits the context as fp16 planes (A operand of linear_out); FFN w_1 emits its
      // ReLU output as planes for w_2 — neither intermediate makes an fp32 round trip through HBM
      FA_RETURN_IF_ERR(attention_tc_planes_launch(q_planes, k_planes, vt_planes, lens, batch, enc->heads, t_max, t_max, nullptr, 0,
                                                  ctx_planes, D, npl, gemm_mode, st));
      if (side) FA_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));          // join: linear_out adds the FSMN memory
      FA_RETURN_IF_ERR(gemm_tc_planes_launch(ctx_planes, M, L.out, 0, mem, D, nullptr, 0, x2, D, nullptr, 0, gemm_mode, st));   // mem already holds residual + memory
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
  if (gemm_mode != FA_GEMM_F32_SIMT && pred->conv.w_planes && pred->conv.in_pad == 3 * D) {
    // tensor-core path: no im2col copy — the GEMM's A operand is the overlapping view of the zero-padded encoder planes (cif.cu)
    const int npl = npl_for(gemm_mode);
    const int64_t Mp = (int64_t)batch * (t_max + 2), rows_alloc = Mp + 2;
    plane_t* pp = scratch.take<plane_t>((size_t)npl * rows_alloc * D);          // <= the im2col planes this scratch was sized for
    if (!scratch.ok()) return FA_ERR_WORKSPACE;
    float* cp = xc;                                                              // [Mp, D] fits the unused im2col buffer (3 D per row)
    FA_RETURN_IF_ERR(cif_pad_planes_launch(enc, batch, t_max, D, npl, rows_alloc, pp, st));
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(pp, Mp, pred->conv, 1, nullptr, 0, nullptr, 0, cp, D, nullptr, 0, gemm_mode, st, nullptr, D, rows_alloc));
    FA_RETURN_IF_ERR(cif_alpha_launch(cp, D, pred->out_w, pred->out_b, lens, t_max, M, pred->smooth_factor,
                                      pred->noise_threshold, alpha_rows, st, t_max + 2));
  } else {
    FA_RETURN_IF_ERR(cif_im2col_launch(enc, M, t_max, D, xc, st));
    FA_RETURN_IF_ERR(linear(xc, 3 * D, M, pred->conv, 1, nullptr, 0, nullptr, 0, c, D, gemm_mode, &scratch, st));
    FA_RETURN_IF_ERR(cif_alpha_launch(c, D, pred->out_w, pred->out_b, lens, t_max, M, pred->smooth_factor,
                                      pred->noise_threshold, alpha_rows, st));
  }
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
                   Arena* scratch, cudaStream_t st, plane_t* t1_planes, plane_t* hq_planes) {
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

// Cross-attention probabilities of ONE utterance (DecoderLayerSANM.get_attn_mat, decoder.py:123-146 ->
// MultiHeadedAttentionCrossAtt.forward_attention with ret_attn, sanm/attention.py:760-794): probs[h, n, t] =
// softmax_t( (q[n, h] * d_k^-0.5) . k[t, h] ) with keys t >= klen masked to -inf before and to 0 after the softmax.
// SeACo's attention-score filtering sums this matrix over heads and tokens on the host exactly like the reference
// (seaco_paraformer/model.py:325-328), so the matrix itself is the output.  One warp per (head, query); tiny (N x n_hotwords).
__global__ void __launch_bounds__(128)
attn_probs_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk, int heads, int n_q, int t_k,
                  const int32_t* __restrict__ key_lens, float qscale, float* __restrict__ probs) {
  extern __shared__ float s_sc[];                       // [4 warps][t_k]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + warp;                // h * n_q + n
  if (row >= heads * n_q) return;
  const int h = row / n_q, n = row - h * n_q;
  const int klen = min(key_lens[0], t_k);                // utterance 0's key count
  float* sc = s_sc + warp * t_k;
  const float* qr = q + (int64_t)n * ldq + h * 128;
  float qv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) qv[j] = __fmul_rn(qr[lane + 32 * j], qscale);
  float mx = -INFINITY;
  for (int t = 0; t < t_k; ++t) {
    const float* kr = k + (int64_t)t * ldk + h * 128;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = fmaf(qv[j], kr[lane + 32 * j], acc);
    acc = warp_sum(acc);
    const float sv = t < klen ? acc : -INFINITY;
    if (lane == 0) sc[t] = sv;
    mx = fmaxf(mx, sv);
  }
  __syncwarp();
  float sum = 0.f;
  for (int t = lane; t < t_k; t += 32) { const float e = t < klen ? expf(sc[t] - mx) : 0.f; sc[t] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  for (int t = lane; t < t_k; t += 32) probs[(int64_t)row * t_k + t] = t < klen ? sc[t] / sum : 0.f;
}

// Workspace slices shared by every decoder-stack entry point (same carve order as dec_plan).
struct DecBuf {
  float *ya, *yb, *t1, *hq, *f, *qd, *ctx, *kv, *lg, *cat;
  char* hw_region; size_t hwb;
  plane_t *ctx_planes, *mem_planes, *t1_planes, *hq_planes, *q_planes, *k_planes, *vt_planes;
  char* sp; size_t sb;
};
static bool dec_carve(Arena& a, DecBuf& b, int batch, int t_max, int n_max, int vocab, int mode, int n_hotwords) {
  const int64_t Mq = (int64_t)batch * n_max, Mk = (int64_t)batch * t_max;
  const bool tc = mode != FA_GEMM_F32_SIMT;
  b.ya = a.take<float>(Mq * 512ull); b.yb = a.take<float>(Mq * 512ull); b.t1 = a.take<float>(Mq * 512ull);
  b.hq = a.take<float>(Mq * 2048ull); b.f = a.take<float>(Mq * 512ull); b.qd = a.take<float>(Mq * 512ull);
  b.ctx = a.take<float>(Mq * 512ull); b.kv = a.take<float>(Mk * 1024ull); b.lg = a.take<float>(Mq * (size_t)vocab);
  b.cat = a.take<float>(Mq * 1024ull);
  b.hwb = hw_region_bytes(batch, n_max, n_hotwords, mode);
  b.hw_region = a.take<char>(b.hwb);
  const int t_pad = (t_max + 63) / 64 * 64;
  b.ctx_planes = tc ? a.take<plane_t>(3ull * Mq * 512) : nullptr;
  b.mem_planes = tc ? a.take<plane_t>(3ull * Mk * 512) : nullptr;
  b.t1_planes = tc ? a.take<plane_t>(3ull * Mq * 512) : nullptr;
  b.hq_planes = tc ? a.take<plane_t>(3ull * Mq * 2048) : nullptr;
  b.q_planes = tc ? a.take<plane_t>(2ull * Mq * 512) : nullptr;
  b.k_planes = tc ? a.take<plane_t>(2ull * Mk * 512) : nullptr;
  b.vt_planes = tc ? a.take<plane_t>(2ull * batch * 512 * (size_t)t_pad) : nullptr;
  b.sb = dec_scratch_bytes(batch, t_max, n_max, mode);
  b.sp = a.take<char>(b.sb);
  return a.ok();
}

// One run of a SAN-M decoder stack over a cross-attention memory.  memory [mem_batch * t_mem, 512] with mem_batch = batch, or 1
// when mem_shared (the SeACo / contextual hotword memory: every utterance attends over the same rows — one k/v projection, one
// copy).  tgt [batch, n_max, 512] lives in b.ya on entry.
struct DecRun {
  int batch, n_max, t_mem, heads, fsmn_k, mode, mem_shared;
  const float* memory; const int32_t* mem_lens; const int32_t* tok_lens;
  cudaStream_t st;
  DecBuf* b; Arena* scratch;
  int64_t Mq() const { return (int64_t)batch * n_max; }
  int64_t Mk() const { return (int64_t)(mem_shared ? 1 : batch) * t_mem; }
};

// One attention decoder layer (DecoderLayerSANM.forward, paraformer/decoder.py:78-121).  *x_self_out receives `residual +
// fsmn(...)` (x_self_attn); if src_out != nullptr the cross-attention output is written there WITHOUT the residual (x_src_attn,
// leading dim ld_src) and *y_next is not produced — the ContextualDecoderLayer contract (contextual_paraformer/decoder.py:60-100).
// attn_probs != nullptr: stop at the cross-attention and write utterance 0's probability matrix [heads, n_max, t_mem] instead
// (get_attn_mat, decoder.py:123-146).
static int dec_attention_layer(const DecRun& r, const FaDecLayer& L, float* yin, float** x_self_out, float* src_out, int64_t ld_src,
                               float** y_next, float* attn_probs) {
  DecBuf& b = *r.b;
  const int D = 512;
  const int64_t Mq = r.Mq(), Mk = r.Mk();
  const bool tc = r.mode != FA_GEMM_F32_SIMT;
  const int npl = npl_for(r.mode);
  const int t_pad = (r.t_mem + 63) / 64 * 64;
  cudaStream_t st = r.st;
  FA_RETURN_IF_ERR(dec_ffn(L, yin, Mq, b.t1, b.hq, b.f, r.mode, r.scratch, st, b.t1_planes, b.hq_planes));
  // x = residual + fsmn(LN2(f), tgt_mask)     decoder.py:103-107
  FA_RETURN_IF_ERR(layernorm_launch(b.f, Mq, L.norm2, b.t1, nullptr, 1.f, 1, st));
  float* x2 = (yin == b.ya) ? b.yb : b.ya;
  FA_RETURN_IF_ERR(fsmn_launch(b.t1, D, r.tok_lens, r.batch, r.n_max, D, L.fsmn_w, r.fsmn_k, yin, D, x2, D, st));
  *x_self_out = x2;
  if (attn_probs) {
    // q / k in fp32 through the mode's GEMM; only utterance 0's rows are needed (seaco_paraformer/model.py:325: hotword_scores[0])
    FA_RETURN_IF_ERR(layernorm_launch(x2, r.n_max, L.norm3, b.t1, nullptr, 1.f, 1, st));
    FA_RETURN_IF_ERR(linear(b.t1, D, r.n_max, L.q, 0, nullptr, 0, nullptr, 0, b.qd, D, r.mode, r.scratch, st));
    FA_RETURN_IF_ERR(linear(r.memory, D, r.t_mem, L.kv, 0, nullptr, 0, nullptr, 0, b.kv, 2 * D, r.mode, r.scratch, st));
    const int rows = r.heads * r.n_max;
    const size_t smem = (size_t)4 * r.t_mem * sizeof(float);
    if (smem > 96 * 1024) return FA_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) FA_CUDA_OK(cudaFuncSetAttribute(attn_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_probs_kernel<<<(rows + 3) / 4, 128, smem, st>>>(b.qd, D, b.kv, 2 * D, r.heads, r.n_max, r.t_mem, r.mem_lens,
                                                         (float)(1.0 / sqrt(128.0)), attn_probs);
    FA_CHECK_LAUNCH();
    return FA_OK;
  }
  // x = residual + src_attn(LN3(x), memory)    decoder.py:109-118, attention.py:796-813
  if (tc) {
    FA_RETURN_IF_ERR(layernorm_launch(x2, Mq, L.norm3, nullptr, nullptr, 1.f, 1, st, b.t1_planes, npl, D));
    AttnSinks sq;                       // q -> scaled fp16 planes only (no fp32 round trip)
    sq.q0 = 0; sq.width = D; sq.npl = npl < 2 ? npl : 2; sq.t_rows = r.n_max; sq.qscale = (float)(1.0 / sqrt(128.0)); sq.q_planes = b.q_planes;
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(b.t1_planes, Mq, L.q, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, r.mode, st, &sq));
  } else {
    FA_RETURN_IF_ERR(layernorm_launch(x2, Mq, L.norm3, b.t1, nullptr, 1.f, 1, st));
    FA_RETURN_IF_ERR(linear(b.t1, D, Mq, L.q, 0, nullptr, 0, nullptr, 0, b.qd, D, r.mode, r.scratch, st));
  }
  float* y2 = (x2 == b.ya) ? b.yb : b.ya;
  float* dst = src_out ? src_out : y2;
  const int64_t ldd = src_out ? ld_src : D;
  const float* res = src_out ? nullptr : x2;
  if (!tc) {
    FA_RETURN_IF_ERR(linear(r.memory, D, Mk, L.kv, 0, nullptr, 0, nullptr, 0, b.kv, 2 * D, r.mode, r.scratch, st));
    FA_RETURN_IF_ERR(attention_f32_launch(b.qd, D, b.kv, 2 * D, b.kv + D, 2 * D, r.mem_lens, r.batch, r.heads, r.n_max, r.t_mem, b.ctx, D, st,
                                          r.mem_shared));
    FA_RETURN_IF_ERR(linear(b.ctx, D, Mq, L.out, 0, res, D, nullptr, 0, dst, ldd, r.mode, r.scratch, st));
  } else {
    AttnSinks skv;                      // k -> planes, v -> transposed planes; nothing in fp32
    skv.k0 = 0; skv.v0 = D; skv.width = D; skv.npl = npl < 2 ? npl : 2; skv.t_rows = r.t_mem; skv.t_pad = t_pad;
    skv.k_planes = b.k_planes; skv.vt_planes = b.vt_planes;
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(b.mem_planes, Mk, L.kv, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, r.mode, st, &skv));
    FA_RETURN_IF_ERR(attention_tc_planes_launch(b.q_planes, b.k_planes, b.vt_planes, r.mem_lens, r.batch, r.heads, r.n_max, r.t_mem, nullptr, 0,
                                                b.ctx_planes, D, npl, r.mode, st, r.mem_shared));
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(b.ctx_planes, Mq, L.out, 0, res, D, nullptr, 0, dst, ldd, nullptr, 0, r.mode, st));
  }
  if (y_next) *y_next = y2;
  return FA_OK;
}

// decoders3 (FFN only, no residual: decoder.py:97-102,121) + after_norm -> hidden fp32 (optional) and/or planes for output_layer
static int dec_finish(const DecRun& r, const FaDecoder* dec, float* y, float* hidden_out) {
  DecBuf& b = *r.b;
  const bool tc = r.mode != FA_GEMM_F32_SIMT;
  FA_RETURN_IF_ERR(dec_ffn(dec->last, y, r.Mq(), b.t1, b.hq, b.f, r.mode, r.scratch, r.st, b.t1_planes, b.hq_planes));
  if (tc) return layernorm_launch(b.f, r.Mq(), dec->after_norm, hidden_out, nullptr, 1.f, 1, r.st, b.t1_planes, npl_for(r.mode), 512);
  return layernorm_launch(b.f, r.Mq(), dec->after_norm, hidden_out ? hidden_out : b.t1, nullptr, 1.f, 1, r.st);
}

static int decoder_forward_impl(const FaDecoder* dec, const float* enc, const int32_t* enc_lens, int32_t batch, int32_t t_max,
                                const float* acoustic, int64_t ld_acoustic_rows, const int32_t* tok_lens, int32_t n_max,
                                int32_t* argmax_ids, float* argmax_logp, float* logits, int32_t log_softmax, float* hidden_out,
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
  DecBuf b;
  if (!dec_carve(a, b, batch, t_max, n_max, V, gemm_mode, dec->has_bias ? dec->n_hotwords : 0)) return FA_ERR_WORKSPACE;
  Arena scratch(b.sp, b.sb);
  const bool tc = gemm_mode != FA_GEMM_F32_SIMT;
  const int npl = npl_for(gemm_mode);
  float* lg = logits ? logits : b.lg;
  if (tc) FA_RETURN_IF_ERR(split_rows_launch(enc, D, Mk, D, D, npl, b.mem_planes, st));   // memory is layer-invariant

  // tgt = acoustic[:, :n_max]  (decoder.py:424)
  FA_CUDA_OK(cudaMemcpy2DAsync(b.ya, (size_t)n_max * D * 4, acoustic, (size_t)ld_acoustic_rows * D * 4, (size_t)n_max * D * 4,
                               batch, cudaMemcpyDeviceToDevice, st));
  fa::count_launch();
  DecRun r{batch, n_max, t_max, dec->heads, dec->fsmn_k, gemm_mode, 0, enc, enc_lens, tok_lens, st, &b, &scratch};
  float* y = b.ya;
  for (int l = 0; l < dec->n_layers; ++l) {
    float* xs = nullptr;
    FA_RETURN_IF_ERR(dec_attention_layer(r, dec->layers[l], y, &xs, nullptr, 0, &y, nullptr));
  }
  if (dec->has_bias) {
    // ContextualParaformerDecoder.forward decoder.py:325-340
    const int nh = dec->n_hotwords;
    if (!dec->hw_embed || !dec->hw_lens || nh <= 0 || dec->clas_scale != 1.0f) return FA_ERR_UNSUPPORTED;
    float* x_self = nullptr;
    FA_RETURN_IF_ERR(dec_attention_layer(r, dec->bias_last, y, &x_self, b.cat, 2 * D, nullptr, nullptr));      // cat[:, :512] = x_src_attn
    // bias decoder: cross attention of LN3(x_self_attn) over the hotword memory (identical for every utterance)
    FA_RETURN_IF_ERR(layernorm_launch(x_self, Mq, dec->bias_norm3, b.t1, nullptr, 1.f, 1, st));
    FA_RETURN_IF_ERR(linear(b.t1, D, Mq, dec->bias_q, 0, nullptr, 0, nullptr, 0, b.qd, D, gemm_mode, &scratch, st));
    // the hotword k | v rows [nh, 1024] are the same for every utterance: one copy, attended with kv_shared (no per-utterance
    // replication, so the hotword count is independent of t_max)
    Arena a2(b.hw_region, b.hwb);
    float* kvh = a2.take<float>((size_t)nh * 1024);
    const size_t sb2 = max_sz(gemm_tc_scratch_bytes(Mq > nh ? Mq : nh, 1024, gemm_mode), attention_tc_scratch_bytes(batch, 4, n_max, nh, gemm_mode));
    char* sp2 = a2.take<char>(sb2);
    if (!a2.ok()) return FA_ERR_WORKSPACE;
    Arena scratch2(sp2, sb2);
    FA_RETURN_IF_ERR(linear(dec->hw_embed, D, nh, dec->bias_kv, 0, nullptr, 0, nullptr, 0, kvh, 2 * D, gemm_mode, &scratch2, st));
    if (!tc) {
      FA_RETURN_IF_ERR(attention_f32_launch(b.qd, D, kvh, 2 * D, kvh + D, 2 * D, dec->hw_lens, batch, dec->heads, n_max, nh, b.ctx, D, st, 1));
    } else {
      FA_RETURN_IF_ERR(attention_tc_launch(b.qd, D, kvh, 2 * D, kvh + D, 2 * D, dec->hw_lens, batch, dec->heads, n_max, nh, b.ctx, D,
                                           nullptr, 0, 0, gemm_mode, &scratch2, st, 1));
    }
    FA_RETURN_IF_ERR(linear(b.ctx, D, Mq, dec->bias_out, 0, nullptr, 0, nullptr, 0, b.cat + D, 2 * D, gemm_mode, &scratch, st));   // cat[:, 512:] = cx
    float* y2 = (x_self == b.ya) ? b.yb : b.ya;
    FA_RETURN_IF_ERR(linear(b.cat, 2 * D, Mq, dec->bias_output, 0, x_self, D, nullptr, 0, y2, D, gemm_mode, &scratch, st));
    y = y2;
  }
  // decoders3, after_norm (-> hidden), output_layer
  FA_RETURN_IF_ERR(dec_finish(r, dec, y, hidden_out));
  if (tc) {
    FA_RETURN_IF_ERR(gemm_tc_planes_launch(b.t1_planes, Mq, dec->output, 0, nullptr, 0, nullptr, 0, lg, V, nullptr, 0, gemm_mode, st));
  } else {
    FA_RETURN_IF_ERR(linear(hidden_out ? hidden_out : b.t1, D, Mq, dec->output, 0, nullptr, 0, nullptr, 0, lg, V, gemm_mode, &scratch, st));
  }
  return argmax_lse_launch(lg, Mq, V, V, argmax_ids, argmax_logp, (logits && log_softmax) ? 1 : 0, st);
}

extern "C" int fa_paraformer_decoder_forward(const FaDecoder* dec, const float* enc, const int32_t* enc_lens,
                                             int32_t batch, int32_t t_max, const float* acoustic,
                                             int64_t ld_acoustic_rows, const int32_t* tok_lens, int32_t n_max,
                                             int32_t* argmax_ids, float* argmax_logp, float* logits, int32_t log_softmax,
                                             int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream) {
  return decoder_forward_impl(dec, enc, enc_lens, batch, t_max, acoustic, ld_acoustic_rows, tok_lens, n_max, argmax_ids, argmax_logp,
                              logits, log_softmax, nullptr, gemm_mode, workspace, ws_bytes, stream);
}

// return_hidden + return_both (decoder.py:441-449): additionally writes the after_norm output [B, n_max, 512]
extern "C" int fa_paraformer_decoder_forward_hidden(const FaDecoder* dec, const float* enc, const int32_t* enc_lens,
                                                    int32_t batch, int32_t t_max, const float* acoustic,
                                                    int64_t ld_acoustic_rows, const int32_t* tok_lens, int32_t n_max,
                                                    int32_t* argmax_ids, float* argmax_logp, float* logits, int32_t log_softmax,
                                                    float* hidden, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                                                    fa_stream_t stream) {
  if (!hidden) return FA_ERR_ARG;
  return decoder_forward_impl(dec, enc, enc_lens, batch, t_max, acoustic, ld_acoustic_rows, tok_lens, n_max, argmax_ids, argmax_logp,
                              logits, log_softmax, hidden, gemm_mode, workspace, ws_bytes, stream);
}

// A SAN-M decoder stack WITHOUT input / output layer over an arbitrary memory — the SeACo decoder of SeacoParaformer
// (seaco_paraformer/model.py:100-110: ParaformerSANMDecoder(use_output_layer=False, wo_input_layer=True), FFN 1024, FSMN k = 21,
// 6 attention layers) attending over the hotword embeddings.  mem_shared != 0: memory is [t_mem, 512], the same for every
// utterance (model.py:306-308 repeats `selected` over the batch).  n_run = attention layers to run (<= dec->n_layers);
//   finish != 0      -> decoders3 + after_norm, hidden [B, n_max, 512] (ParaformerSANMDecoder.forward, decoder.py:397-449)
//   attn_probs != 0  -> instead, layer n_run - 1 stops at its cross-attention and writes utterance 0's probability matrix
//                       [heads, n_max, t_mem] (forward_asf6 / get_attn_mat, decoder.py:485-513,123-146); hidden is not written.
extern "C" size_t fa_sanm_decoder_stack_workspace_bytes(int32_t batch, int32_t t_mem, int32_t n_max, int32_t gemm_mode) {
  return dec_plan(batch, t_mem, n_max, 0, gemm_mode, dec_scratch_bytes(batch, t_mem, n_max, gemm_mode), 0);
}

extern "C" int fa_sanm_decoder_stack_forward(const FaDecoder* dec, const float* memory, const int32_t* mem_lens, int32_t mem_shared,
                                             int32_t batch, int32_t t_mem, const float* x, int64_t ld_x_rows,
                                             const int32_t* tok_lens, int32_t n_max, int32_t n_run, int32_t finish, float* hidden,
                                             float* attn_probs, int32_t gemm_mode, void* workspace, size_t ws_bytes,
                                             fa_stream_t stream) {
  if (!dec || !memory || !mem_lens || !x || !tok_lens || batch <= 0 || t_mem <= 0 || n_max <= 0 || ld_x_rows < n_max || n_run < 0 ||
      n_run > dec->n_layers || (!hidden && !attn_probs) || (attn_probs && n_run < 1))
    return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 512;
  if (dec->after_norm.n != D || dec->heads * 128 != D) return FA_ERR_UNSUPPORTED;
  Arena a(workspace, ws_bytes);
  DecBuf b;
  if (!dec_carve(a, b, batch, t_mem, n_max, 0, gemm_mode, 0)) return FA_ERR_WORKSPACE;
  Arena scratch(b.sp, b.sb);
  const bool tc = gemm_mode != FA_GEMM_F32_SIMT;
  DecRun r{batch, n_max, t_mem, dec->heads, dec->fsmn_k, gemm_mode, mem_shared ? 1 : 0, memory, mem_lens, tok_lens, st, &b, &scratch};
  if (tc) FA_RETURN_IF_ERR(split_rows_launch(memory, D, r.Mk(), D, D, npl_for(gemm_mode), b.mem_planes, st));
  FA_CUDA_OK(cudaMemcpy2DAsync(b.ya, (size_t)n_max * D * 4, x, (size_t)ld_x_rows * D * 4, (size_t)n_max * D * 4, batch,
                               cudaMemcpyDeviceToDevice, st));
  fa::count_launch();
  float* y = b.ya;
  for (int l = 0; l < n_run; ++l) {
    float* xs = nullptr;
    const bool last_probs = attn_probs && l == n_run - 1;
    FA_RETURN_IF_ERR(dec_attention_layer(r, dec->layers[l], y, &xs, nullptr, 0, last_probs ? nullptr : &y, last_probs ? attn_probs : nullptr));
  }
  if (attn_probs) return FA_OK;
  if (finish) return dec_finish(r, dec, y, hidden);
  FA_CUDA_OK(cudaMemcpyAsync(hidden, y, (size_t)r.Mq() * D * 4, cudaMemcpyDeviceToDevice, st));
  fa::count_launch();
  return FA_OK;
}

// SeACo merge (seaco_paraformer/model.py:357-378 with seaco_weight = 1): per token row, if argmax(dha_pred) == NO_BIAS keep the
// decoder's distribution, else take the hotword decoder's -> merged arg-max id and its log-probability; optionally the full
// merged log-prob rows (both inputs must then be log-softmax rows).
__global__ void seaco_merge_kernel(const int32_t* __restrict__ dec_ids, const float* __restrict__ dec_best, const int32_t* __restrict__ dha_ids,
                                   const float* __restrict__ dha_best, int64_t rows, int no_bias, int32_t* __restrict__ out_ids,
                                   float* __restrict__ out_best, const float* __restrict__ dec_logp, const float* __restrict__ dha_logp,
                                   float* __restrict__ merged, int vocab) {
  const int64_t row = blockIdx.x;
  const bool keep_dec = dha_ids[row] == no_bias;
  if (threadIdx.x == 0) {
    out_ids[row] = keep_dec ? dec_ids[row] : dha_ids[row];
    out_best[row] = keep_dec ? dec_best[row] : dha_best[row];
  }
  if (merged) {
    const float* src = (keep_dec ? dec_logp : dha_logp) + row * vocab;
    // dec * mask + dha * (1 - mask) with mask in {0, 1}: x * 1 + y * 0 — the reference's arithmetic keeps x bit for bit (finite y)
    for (int c = threadIdx.x; c < vocab; c += blockDim.x) merged[row * vocab + c] = src[c];
  }
}

extern "C" int fa_seaco_merge(const int32_t* dec_ids, const float* dec_best, const int32_t* dha_ids, const float* dha_best, int64_t rows,
                              int32_t no_bias, int32_t* out_ids, float* out_best, const float* dec_logp, const float* dha_logp,
                              float* merged, int32_t vocab, fa_stream_t stream) {
  if (!dec_ids || !dec_best || !dha_ids || !dha_best || !out_ids || !out_best || rows < 0) return FA_ERR_ARG;
  if (merged && (!dec_logp || !dha_logp || vocab <= 0)) return FA_ERR_ARG;
  if (rows == 0) return FA_OK;
  seaco_merge_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(dec_ids, dec_best, dha_ids, dha_best, rows, no_bias, out_ids, out_best,
                                                                      dec_logp, dha_logp, merged, vocab);
  FA_CHECK_LAUNCH();
  return FA_OK;
}

// hotword_output_layer + log-softmax arg-max over hidden rows (seaco_paraformer/model.py:352-355): logits = (a + b) W^T + bias for
// the two attended streams a = cif_attended, b = dec_attended (model.py:351: merged = cif_attended + dec_attended).
extern "C" size_t fa_linear_argmax_workspace_bytes(int64_t rows, int32_t vocab, int32_t gemm_mode) {
  ArenaSizer s;
  s.take((size_t)rows * 512 * 4);
  s.take((size_t)rows * vocab * 4);
  s.take(gemm_tc_scratch_bytes(rows, 512, gemm_mode));
  return s.off + 256;
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
  reinterpret_cast<float4*>(o)[i] = make_float4(__fadd_rn(x.x, y.x), __fadd_rn(x.y, y.y), __fadd_rn(x.z, y.z), __fadd_rn(x.w, y.w));
}

extern "C" int fa_linear_argmax(const FaLinear* lin, const float* a, const float* b_or_null, int64_t rows, int32_t* ids, float* best_logp,
                                float* logp, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream) {
  if (!lin || !a || !ids || !best_logp || rows <= 0 || lin->in_f <= 0 || lin->in_f > 512 || (lin->in_f & 3)) return FA_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int V = lin->out_f, K = lin->in_f;
  Arena ar(workspace, ws_bytes);
  float* sum = ar.take<float>((size_t)rows * 512);
  float* lg = ar.take<float>((size_t)rows * V);
  const size_t sb = gemm_tc_scratch_bytes(rows, 512, gemm_mode);
  char* sp = ar.take<char>(sb);
  if (!ar.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);
  const float* x = a;
  if (b_or_null) {
    const int64_t n4 = rows * (K / 4);
    add_rows_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(a, b_or_null, sum, n4);
    FA_CHECK_LAUNCH();
    x = sum;
  }
  if (logp) lg = logp;
  FA_RETURN_IF_ERR(linear(x, K, rows, *lin, 0, nullptr, 0, nullptr, 0, lg, V, gemm_mode, &scratch, st));
  return argmax_lse_launch(lg, rows, V, V, ids, best_logp, logp ? 1 : 0, st);
}

// ------------------------------------------------------------------------------------------ CTC greedy head
extern "C" size_t fa_ctc_greedy_workspace_bytes(int32_t batch, int32_t t_max, int32_t vocab, int32_t gemm_mode) {
  const int64_t M = (int64_t)batch * t_max;
  ArenaSizer s;
  s.take(M * (size_t)((vocab + 3) & ~3) * 4);
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
  // internal logits rows are pitched to a multiple of 4 floats (25055 -> 25056) so the GEMM epilogue and the arg-max sweep use
  // 16-byte accesses; a caller-provided log-prob tensor keeps the dense [M, V] layout
  const int64_t ldv = logp ? V : ((V + 3) & ~3);
  float* lg = a.take<float>(M * (size_t)((V + 3) & ~3));
  float* best = a.take<float>(M);
  const size_t sb = gemm_tc_scratch_bytes(M, 512, gemm_mode);
  char* sp = a.take<char>(sb);
  if (!a.ok()) return FA_ERR_WORKSPACE;
  Arena scratch(sp, sb);
  if (logp) lg = logp;
  FA_RETURN_IF_ERR(linear(enc, ctc_lo->in_f, M, *ctc_lo, 0, nullptr, 0, nullptr, 0, lg, ldv, gemm_mode, &scratch, st));
  FA_RETURN_IF_ERR(argmax_lse_launch(lg, M, V, ldv, argmax_ids, best, logp ? 1 : 0, st));
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
  return split_rows_launch(x, ldx, rows, cols, cols_pad, nplanes, reinterpret_cast<plane_t*>(planes), (cudaStream_t)stream);
}

extern "C" int fa_linear_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, const float* res1,
                                int64_t ld_res1, const float* res2, int64_t ld_res2, float* y, int64_t ldy, int32_t gemm_mode,
                                fa_stream_t stream) {
  if (!a_planes || !lin || !y || gemm_mode == FA_GEMM_F32_SIMT) return FA_ERR_ARG;
  return gemm_tc_planes_launch(reinterpret_cast<const plane_t*>(a_planes), rows, *lin, relu, res1, ld_res1, res2, ld_res2, y, ldy,
                               nullptr, 0, gemm_mode, (cudaStream_t)stream);
}

extern "C" int fa_linear_planes_to_planes(const void* a_planes, int64_t rows, const FaLinear* lin, int32_t relu, void* out_planes,
                                          int64_t ld_out, int32_t gemm_mode, fa_stream_t stream) {
  if (!a_planes || !lin || !out_planes || gemm_mode == FA_GEMM_F32_SIMT) return FA_ERR_ARG;
  return gemm_tc_planes_launch(reinterpret_cast<const plane_t*>(a_planes), rows, *lin, relu, nullptr, 0, nullptr, 0, nullptr, 0,
                               reinterpret_cast<plane_t*>(out_planes), ld_out, gemm_mode, (cudaStream_t)stream);
}

// rows of an embedding table: out[i, :] = table[ids[i], :] (torch.nn.Embedding forward, e.g. CTTransformer.embed ct_transformer/model.py:120)
__global__ void embedding_kernel(const int32_t* __restrict__ ids, const float* __restrict__ table, int dim, int vocab, int64_t n, float* __restrict__ out) {
  const int64_t i = blockIdx.x;
  const int id = min(max(ids[i], 0), vocab - 1);
  for (int c = threadIdx.x; c < dim; c += blockDim.x) out[i * dim + c] = table[(int64_t)id * dim + c];
}
extern "C" int fa_embedding(const int32_t* ids, const float* table, int32_t dim, int32_t vocab, int64_t n, float* out, fa_stream_t stream) {
  if (!ids || !table || !out || dim <= 0 || vocab <= 0 || n < 0) return FA_ERR_ARG;
  if (n == 0) return FA_OK;
  embedding_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(ids, table, dim, vocab, n, out);
  FA_CHECK_LAUNCH();
  return FA_OK;
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
