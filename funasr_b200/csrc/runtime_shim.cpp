// FunOffline* — the OfflineStream entry points of FunASR's C++ runtime (runtime/onnxruntime/include/funasrruntime.h:100-116) with
// their exact C++ signatures, over this library's handle API (offline.cu: fa_offline_*).  Host-only C++: file / buffer decoding
// (raw s16le PCM, RIFF WAV PCM16 / float32), the hotword encoder of ContextualParaformer (Embedding + 1-layer LSTM, O(#hotwords):
// the reference runs it on the CPU too — model_eb.onnx, runtime/onnxruntime/src/paraformer.cpp CompileHotwordEmbedding) and the
// ids -> text join.  Everything per audio frame runs in fa_offline_infer_hw on the GPU.
#include "../../include/funasrruntime_b200.h"
#include "../../include/funasr_b200.h"

#include <math.h>
#include <stdio.h>
#include <string.h>
#include <fstream>
#include <sstream>

namespace {

thread_local std::string g_shim_err;

struct OfflineStream {
  void* h = nullptr;
  std::vector<std::string> vocab;
  std::unordered_map<std::string, int> token_id;
  int batch = 1;
};

struct ShimResult {
  std::vector<std::string> msgs;
  std::string stamp, stamp_sents;
  float snippet_time = 0.f;
};

bool is_ascii_word(const std::string& s) {
  if (s.empty()) return false;
  for (unsigned char c : s) if (c >= 0x80) return false;
  return true;
}

// tokens -> text: word pieces ending in "@@" are glued to the next token, consecutive ASCII words are separated by one space,
// CJK tokens are concatenated.  A plain join in the spirit of the runtime's Vocab::Vector2StringV2 (runtime/onnxruntime/src/vocab.cpp:
// 164-278) for Paraformer's char / word vocabulary — NOT a restatement of it: that function also drops <s> / </s> / <unk>, keeps
// runs of single letters unspaced, repairs "xx@@" before a CJK token and has an en-bpe mode.  Text post-processing is CPU string
// work outside the accelerated path; the ids this shim returns are the parity-checked output.
std::string join_tokens(const OfflineStream& s, const int32_t* ids, int n) {
  std::string out;
  bool prev_ascii = false, glue = false;
  for (int i = 0; i < n; ++i) {
    std::string tok = (ids[i] >= 0 && ids[i] < (int)s.vocab.size()) ? s.vocab[ids[i]] : std::to_string(ids[i]);
    if (s.vocab.empty()) tok = std::to_string(ids[i]);
    bool next_glue = false;
    if (tok.size() > 2 && tok.compare(tok.size() - 2, 2, "@@") == 0) { tok.resize(tok.size() - 2); next_glue = true; }
    const bool ascii = is_ascii_word(tok);
    if (!out.empty() && !glue && ascii && prev_ascii) out += ' ';
    out += tok;
    prev_ascii = ascii;
    glue = next_glue;
  }
  return out;
}

// UTF-8 aware split of one hotword into vocabulary units: ASCII runs are one (lower-cased) unit, every other code point is one unit
std::vector<std::string> split_units(const std::string& w) {
  std::vector<std::string> u;
  size_t i = 0;
  while (i < w.size()) {
    const unsigned char c = (unsigned char)w[i];
    if (c < 0x80) {
      std::string a;
      while (i < w.size() && (unsigned char)w[i] < 0x80) { a += (char)tolower(w[i]); ++i; }
      u.push_back(a);
    } else {
      const int len = c >= 0xF0 ? 4 : (c >= 0xE0 ? 3 : 2);
      u.push_back(w.substr(i, len));
      i += len;
    }
  }
  return u;
}

inline float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// RIFF WAVE: returns the PCM payload and its format (1 = s16le, 0 = float32); mono or the first channel layout is required
bool parse_wav(const std::string& bytes, const char** data, size_t* n_bytes, int* fmt, int* rate) {
  if (bytes.size() < 44 || memcmp(bytes.data(), "RIFF", 4) != 0 || memcmp(bytes.data() + 8, "WAVE", 4) != 0) return false;
  size_t pos = 12;
  int channels = 1, bits = 16, tag = 1;
  *rate = 16000;
  while (pos + 8 <= bytes.size()) {
    uint32_t sz;
    memcpy(&sz, bytes.data() + pos + 4, 4);
    if (memcmp(bytes.data() + pos, "fmt ", 4) == 0 && pos + 8 + 16 <= bytes.size()) {
      uint16_t t, ch, b;
      uint32_t r;
      memcpy(&t, bytes.data() + pos + 8, 2); memcpy(&ch, bytes.data() + pos + 10, 2); memcpy(&r, bytes.data() + pos + 12, 4);
      memcpy(&b, bytes.data() + pos + 22, 2);
      tag = t; channels = ch; bits = b; *rate = (int)r;
    } else if (memcmp(bytes.data() + pos, "data", 4) == 0) {
      if (channels != 1) return false;
      *data = bytes.data() + pos + 8;
      *n_bytes = sz <= bytes.size() - pos - 8 ? sz : bytes.size() - pos - 8;
      if (tag == 1 && bits == 16) { *fmt = 1; return true; }
      if (tag == 3 && bits == 32) { *fmt = 0; return true; }
      return false;
    }
    pos += 8 + sz + (sz & 1);
  }
  return false;
}

FUNASR_RESULT infer_pcm(OfflineStream* s, const char* data, size_t n_bytes, int fmt, const std::vector<std::vector<float>>& hw_emb) {
  const int64_t n = (int64_t)(n_bytes / (fmt == 1 ? 2 : 4));
  std::vector<float> hw;
  int n_hw = 0;
  if (fa_offline_is_contextual(s->h)) {
    for (const auto& row : hw_emb) if (row.size() == 512) { hw.insert(hw.end(), row.begin(), row.end()); ++n_hw; }
    if (n_hw == 0) { g_shim_err = "contextual model: hw_emb must hold [n, 512] rows from CompileHotwordEmbedding"; return nullptr; }
  }
  const void* bufs[1] = {data};
  const int64_t lens[1] = {n};
  void* r = fa_offline_infer_hw(s->h, bufs, lens, 1, fmt, n_hw ? hw.data() : nullptr, n_hw);
  if (!r) { g_shim_err = fa_offline_last_error(); return nullptr; }
  ShimResult* out = new ShimResult();
  const int cnt = fa_offline_result_count(r);
  for (int i = 0; i < cnt; ++i) {
    int32_t k = 0;
    const int32_t* ids = fa_offline_result_ids(r, i, &k);
    out->msgs.push_back(join_tokens(*s, ids, k));
  }
  out->snippet_time = fa_offline_result_audio_seconds(r);
  fa_offline_free_result(r);
  return out;
}

}  // namespace

const char* FunB200LastError() { return g_shim_err.c_str(); }

FUNASR_HANDLE FunOfflineInit(std::map<std::string, std::string>& model_path, int thread_num, bool use_gpu, int batch_size) {
  (void)thread_num; (void)use_gpu;
  g_shim_err.clear();
  auto it = model_path.find("model-dir");
  if (it == model_path.end()) { g_shim_err = "model_path[\"model-dir\"] is missing"; return nullptr; }
  const std::string dir = it->second;
  int mode = FA_GEMM_F16X3, device = 0;
  auto gm = model_path.find("gemm-mode");
  if (gm != model_path.end()) {
    if (gm->second == "fp32") mode = FA_GEMM_F32_SIMT;
    else if (gm->second == "fp16") mode = FA_GEMM_F16X1;
    else if (gm->second == "fp16x6") mode = FA_GEMM_F16X6;
    else if (gm->second != "fp16x3") { g_shim_err = "unknown gemm-mode " + gm->second; return nullptr; }
  }
  auto gi = model_path.find("gpu-id");
  if (gi != model_path.end()) device = atoi(gi->second.c_str());
  OfflineStream* s = new OfflineStream();
  s->batch = batch_size > 0 ? batch_size : 1;
  s->h = fa_offline_init((dir + "/model.fab2").c_str(), device, mode);
  if (!s->h) { g_shim_err = fa_offline_last_error(); delete s; return nullptr; }
  std::ifstream tf(dir + "/tokens.txt");
  std::string line;
  while (tf && std::getline(tf, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    s->token_id[line] = (int)s->vocab.size();
    s->vocab.push_back(line);
  }
  return s;
}

void FunOfflineReset(FUNASR_HANDLE, FUNASR_DEC_HANDLE) {}

void FunOfflineUninit(FUNASR_HANDLE handle) {
  OfflineStream* s = static_cast<OfflineStream*>(handle);
  if (!s) return;
  fa_offline_uninit(s->h);
  delete s;
}

FUNASR_RESULT FunOfflineInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE, QM_CALLBACK fn_callback,
                                    const std::vector<std::vector<float>>& hw_emb, int sampling_rate, std::string wav_format, bool,
                                    FUNASR_DEC_HANDLE, std::string, bool) {
  g_shim_err.clear();
  OfflineStream* s = static_cast<OfflineStream*>(handle);
  if (!s || !sz_buf || n_len <= 0) { g_shim_err = "bad argument"; return nullptr; }
  const char* data = sz_buf;
  size_t nb = (size_t)n_len;
  int fmt = 1, rate = sampling_rate;
  std::string holder;
  if (wav_format == "wav") {
    holder.assign(sz_buf, (size_t)n_len);
    if (!parse_wav(holder, &data, &nb, &fmt, &rate)) { g_shim_err = "unsupported WAV (need mono PCM16 or float32)"; return nullptr; }
  } else if (wav_format != "pcm") { g_shim_err = "wav_format must be \"pcm\" (s16le) or \"wav\""; return nullptr; }
  if (rate != 16000) { g_shim_err = "audio must be 16 kHz"; return nullptr; }
  FUNASR_RESULT r = infer_pcm(s, data, nb, fmt, hw_emb);
  if (fn_callback) fn_callback(1, 1);
  return r;
}

FUNASR_RESULT FunOfflineInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                              const std::vector<std::vector<float>>& hw_emb, int sampling_rate, bool itn, FUNASR_DEC_HANDLE dec_handle) {
  g_shim_err.clear();
  if (!sz_filename) { g_shim_err = "bad argument"; return nullptr; }
  std::ifstream f(sz_filename, std::ios::binary);
  if (!f) { g_shim_err = std::string("cannot open ") + sz_filename; return nullptr; }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string bytes = ss.str();
  const std::string name = sz_filename;
  const bool wav = name.size() > 4 && (name.compare(name.size() - 4, 4, ".wav") == 0 || name.compare(name.size() - 4, 4, ".WAV") == 0);
  return FunOfflineInferBuffer(handle, bytes.data(), (int)bytes.size(), mode, fn_callback, hw_emb, sampling_rate, wav ? "wav" : "pcm", itn, dec_handle);
}

// bias_embed -> 1-layer LSTM (gate order i, f, g, o) -> the hidden state after each hotword's last token
// (contextual_paraformer/model.py:350-372); the hotword list is followed by the <s> entry (model.py:606-607)
const std::vector<std::vector<float>> CompileHotwordEmbedding(FUNASR_HANDLE handle, std::string& hotwords, ASR_TYPE) {
  g_shim_err.clear();
  std::vector<std::vector<float>> out;
  OfflineStream* s = static_cast<OfflineStream*>(handle);
  if (!s || !fa_offline_is_contextual(s->h)) return out;
  int64_t n_emb = 0, n_ih = 0, n_hh = 0, n_bi = 0, n_bh = 0;
  const float* emb = fa_offline_host_tensor(s->h, "bias_embed.weight", &n_emb);
  const float* w_ih = fa_offline_host_tensor(s->h, "bias_encoder.weight_ih_l0", &n_ih);
  const float* w_hh = fa_offline_host_tensor(s->h, "bias_encoder.weight_hh_l0", &n_hh);
  const float* b_ih = fa_offline_host_tensor(s->h, "bias_encoder.bias_ih_l0", &n_bi);
  const float* b_hh = fa_offline_host_tensor(s->h, "bias_encoder.bias_hh_l0", &n_bh);
  const int D = 512;
  if (!emb || !w_ih || !w_hh || !b_ih || !b_hh || n_ih != 4 * D * D || n_hh != 4 * D * D) { g_shim_err = "model file has no hotword encoder"; return out; }
  const int vocab = (int)(n_emb / D);
  std::vector<std::vector<int>> lists;
  std::stringstream ss(hotwords);
  std::string w;
  while (ss >> w) {
    std::vector<int> ids;
    bool ok = true;
    for (const std::string& u : split_units(w)) {
      int id = -1;
      auto it = s->token_id.find(u);
      if (it != s->token_id.end()) id = it->second;
      else if (s->vocab.empty()) id = atoi(u.c_str());                  // no vocabulary file: decimal token ids
      if (id < 0 || id >= vocab) { ok = false; break; }
      ids.push_back(id);
    }
    if (ok && !ids.empty()) lists.push_back(ids);                      // hotwords with out-of-vocabulary units are dropped
  }
  lists.push_back({1});                                                // <s>
  std::vector<float> h(D), c(D), gates(4 * D);
  for (const auto& ids : lists) {
    std::fill(h.begin(), h.end(), 0.f);
    std::fill(c.begin(), c.end(), 0.f);
    for (int id : ids) {
      const float* x = emb + (size_t)id * D;
      for (int g = 0; g < 4 * D; ++g) {
        const float* wi = w_ih + (size_t)g * D;
        const float* wh = w_hh + (size_t)g * D;
        float a = b_ih[g] + b_hh[g];
        float acc1 = 0.f, acc2 = 0.f;
        for (int k = 0; k < D; ++k) { acc1 += wi[k] * x[k]; acc2 += wh[k] * h[k]; }
        gates[g] = a + acc1 + acc2;
      }
      for (int k = 0; k < D; ++k) {
        const float ig = sigm(gates[k]), fg = sigm(gates[D + k]), gg = tanhf(gates[2 * D + k]), og = sigm(gates[3 * D + k]);
        c[k] = fg * c[k] + ig * gg;
      }
      for (int k = 0; k < D; ++k) h[k] = sigm(gates[3 * D + k]) * tanhf(c[k]);
    }
    out.push_back(h);
  }
  return out;
}

const char* FunASRGetResult(FUNASR_RESULT result, int n_index) {
  ShimResult* r = static_cast<ShimResult*>(result);
  if (!r || n_index < 0 || n_index >= (int)r->msgs.size()) return nullptr;
  return r->msgs[n_index].c_str();
}
const char* FunASRGetStamp(FUNASR_RESULT result) { return result ? static_cast<ShimResult*>(result)->stamp.c_str() : nullptr; }
const char* FunASRGetStampSents(FUNASR_RESULT result) { return result ? static_cast<ShimResult*>(result)->stamp_sents.c_str() : nullptr; }
const int FunASRGetRetNumber(FUNASR_RESULT result) { return result ? (int)static_cast<ShimResult*>(result)->msgs.size() : 0; }
void FunASRFreeResult(FUNASR_RESULT result) { delete static_cast<ShimResult*>(result); }
const float FunASRGetRetSnippetTime(FUNASR_RESULT result) { return result ? static_cast<ShimResult*>(result)->snippet_time : 0.f; }

FUNASR_DEC_HANDLE FunASRWfstDecoderInit(FUNASR_HANDLE, int, float, float, float) { return nullptr; }
void FunASRWfstDecoderUninit(FUNASR_DEC_HANDLE) {}
void FunWfstDecoderLoadHwsRes(FUNASR_DEC_HANDLE, int, std::unordered_map<std::string, int>&) {}
void FunWfstDecoderUnloadHwsRes(FUNASR_DEC_HANDLE) {}
