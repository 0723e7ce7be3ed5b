// Fused Fbank + LFR + CMVN frontend (one kernel): wav[B, N] -> feats[B, T, 560].
//
// Replaces WavFrontend.forward (funasr/frontends/wav_frontend.py:149-196): per utterance
//   x32768 (:169) -> torchaudio.compliance.kaldi.fbank(dither=0, hamming, 25/10 ms, 80 mel, snip_edges)
//   (kaldi.py:514-647) -> apply_lfr m=7 n=6 (:63-86) -> apply_cmvn (:46-60) -> pad_sequence(0.0) (:195).
//
// HBM-bound by design: algorithmic bytes = 4 B/sample in + 4*560 B/LFR-row out (3.04 MB per 30 s).
// One CTA owns kRows consecutive LFR rows of one utterance: it stages the (6*kRows+1) frames' worth of
// samples into shared memory once (coalesced), each warp turns frames into log-mel rows with a
// shared-memory 256-point complex Stockham FFT (real 512-point FFT by even/odd packing), and the CTA
// then writes its LFR rows (7 stacked log-mel frames, CMVN applied) with fully coalesced stores.
// Adjacent CTAs recompute one overlapping frame (1/48 redundancy) instead of round-tripping log-mel
// through HBM.
#include "common.cuh"

namespace fa {

constexpr int kWin = 400, kShift = 160, kFft = 512, kBins = 257, kMel = 80;
constexpr int kLfrM = 7, kLfrN = 6, kFeat = kMel * kLfrM;
constexpr int kRows = 8;                               // LFR rows per CTA
constexpr int kFrames = kLfrN * (kRows - 1) + kLfrM;   // 49 frames feed 8 rows
constexpr int kSpan = (kFrames - 1) * kShift + kWin;   // 8080 samples
constexpr int kWarps = 8;
constexpr int kMelPackMax = 1024;

struct FbankSmem {
  float wav[kSpan + 8];
  float logmel[kFrames * kMel];
  float2 fft[kWarps][2][256];
  float2 tw[256];
  float win[kWin];
  float melw[kMelPackMax];
  int mel_start[kMel], mel_len[kMel], mel_off[kMel];
};

__global__ void __launch_bounds__(kWarps * 32)
fbank_lfr_cmvn_kernel(const float* __restrict__ wav, const int32_t* __restrict__ wav_lens, int64_t wav_stride,
                      const float* __restrict__ cmvn, const float* __restrict__ mel_banks,
                      const float* __restrict__ window, float* __restrict__ feats, int32_t* __restrict__ feat_lens,
                      int t_max, int64_t batch_stride_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FbankSmem& s = *reinterpret_cast<FbankSmem*>(smem_raw);
  const int b = blockIdx.y, i0 = blockIdx.x * kRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = wav_lens[b];
  const int m = n >= kWin ? 1 + (n - kWin) / kShift : 0;      // kaldi.py:_get_strided, snip_edges
  const int t_b = (m + kLfrN - 1) / kLfrN;                    // wav_frontend.py:73
  if (blockIdx.x == 0 && tid == 0) feat_lens[b] = t_b;

  float* out = feats + ((int64_t)b * batch_stride_rows + i0) * kFeat;
  if (i0 >= t_b) {  // pure padding rows
    const int rows = min(kRows, t_max - i0);
    for (int idx = tid; idx < rows * kFeat; idx += blockDim.x) out[idx] = 0.f;
    return;
  }
  const int f_lo = max(0, kLfrN * i0 - (kLfrM - 1) / 2);
  const int f_hi = min(m - 1, kLfrN * (i0 + kRows - 1) + (kLfrM - 1) / 2);
  const int nfr = f_hi - f_lo + 1;

  // ---- stage samples, window, twiddles and the sparse mel filters ----
  {
    const float* src = wav + (int64_t)b * wav_stride + (int64_t)f_lo * kShift;
    const int span = (nfr - 1) * kShift + kWin;
    for (int j = tid; j < span; j += blockDim.x) s.wav[j] = __ldg(src + j) * 32768.0f;   // exact scaling
    for (int j = tid; j < kWin; j += blockDim.x) s.win[j] = window[j];
    for (int k = tid; k < 256; k += blockDim.x) {
      float sn, cs;
      sincospif((float)k * (1.0f / 256.0f), &sn, &cs);        // e^{-2 pi i k / 512}
      s.tw[k] = make_float2(cs, -sn);
    }
    // non-zero support of each triangular filter (kaldi.py:get_mel_banks): warp-parallel coalesced scan
    for (int j = warp; j < kMel; j += kWarps) {
      const float* row = mel_banks + j * kBins;
      int st = kBins, en = -1;
      for (int k = lane; k < kBins; k += 32) {
        if (__ldg(row + k) != 0.f) { st = min(st, k); en = max(en, k); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        st = min(st, __shfl_xor_sync(0xffffffffu, st, o));
        en = max(en, __shfl_xor_sync(0xffffffffu, en, o));
      }
      if (lane == 0) { s.mel_start[j] = en < 0 ? 0 : st; s.mel_len[j] = en < 0 ? 0 : en - st + 1; }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int off = 0;
    for (int j = 0; j < kMel; ++j) { s.mel_off[j] = off; off += s.mel_len[j]; if (off > kMelPackMax) { s.mel_len[j] = 0; off = s.mel_off[j]; } }
  }
  __syncthreads();
  if (tid < kMel) {
    const float* row = mel_banks + tid * kBins + s.mel_start[tid];
    float* dst = s.melw + s.mel_off[tid];
    for (int k = 0; k < s.mel_len[tid]; ++k) dst[k] = row[k];
  }
  __syncthreads();

  // ---- one warp per frame ----
  float2* A = s.fft[warp][0];
  float2* Bf = s.fft[warp][1];
  for (int f = warp; f < nfr; f += kWarps) {
    const float* x = s.wav + f * kShift;
    float part = 0.f;
    for (int j = lane; j < kWin; j += 32) part += x[j];
    const float mean = warp_sum(part) / (float)kWin;                // remove_dc_offset kaldi.py:183-186
    float* Af = reinterpret_cast<float*>(A);
    for (int j = lane; j < kFft; j += 32) {
      float y = 0.f;
      if (j < kWin) {
        const float cur = __fsub_rn(x[j], mean);
        const float prev = __fsub_rn(x[j > 0 ? j - 1 : 0], mean);   // replicate pad, :193-198
        y = __fmul_rn(__fsub_rn(cur, __fmul_rn(0.97f, prev)), s.win[j]);
      }
      Af[j] = y;                                                    // z[n] = y[2n] + i y[2n+1]
    }
    __syncwarp();
    float2* in = A;
    float2* outb = Bf;
#pragma unroll 1
    for (int p = 1; p < 256; p <<= 1) {                             // radix-2 Stockham, 8 passes
      const int tws = 256 / p;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = lane + 32 * q;
        const int k = i & (p - 1);
        const int j = ((i - k) << 1) + k;
        const float2 w = s.tw[k * tws];
        const float2 u0 = in[i], v = in[i + 128];
        const float2 u1 = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
        outb[j] = make_float2(u0.x + u1.x, u0.y + u1.y);
        outb[j + p] = make_float2(u0.x - u1.x, u0.y - u1.y);
      }
      __syncwarp();
      float2* t = in; in = outb; outb = t;
    }
    // after 8 passes the spectrum of z sits in `in` (== A); power spectrum of the real signal -> P[0..256]
    float* P = reinterpret_cast<float*>(outb);
    for (int k = lane; k < kBins; k += 32) {
      float re, im;
      if (k == 0) { re = in[0].x + in[0].y; im = 0.f; }
      else if (k == 256) { re = in[0].x - in[0].y; im = 0.f; }
      else {
        const float2 zk = in[k], zc = in[256 - k];
        const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
        const float orr = 0.5f * (zk.y + zc.y), oi = -0.5f * (zk.x - zc.x);
        const float2 w = s.tw[k];
        re = er + (w.x * orr - w.y * oi);
        im = ei + (w.x * oi + w.y * orr);
      }
      const float mag = sqrtf(re * re + im * im);                   // rfft(..).abs().pow(2.0) :616-618
      P[k] = mag * mag;
    }
    __syncwarp();
    for (int j = lane; j < kMel; j += 32) {
      const float* wts = s.melw + s.mel_off[j];
      const float* pp = P + s.mel_start[j];
      float acc = 0.f;
      for (int k = 0; k < s.mel_len[j]; ++k) acc = fmaf(pp[k], wts[k], acc);
      s.logmel[f * kMel + j] = logf(fmaxf(acc, 1.1920929e-07f));    // :632-633
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- LFR stacking + CMVN, coalesced row stores ----
  const int rows = min(kRows, t_max - i0);
  for (int idx = tid; idx < rows * kFeat; idx += blockDim.x) {
    const int r = idx / kFeat, col = idx - r * kFeat;
    const int i = i0 + r;
    float v = 0.f;
    if (i < t_b) {
      const int j = col / kMel, c = col - j * kMel;
      int fsrc = kLfrN * i - (kLfrM - 1) / 2 + j;
      fsrc = min(max(fsrc, 0), m - 1);
      v = s.logmel[(fsrc - f_lo) * kMel + c];
      if (cmvn != nullptr) v = __fmul_rn(__fadd_rn(v, __ldg(cmvn + col)), __ldg(cmvn + kFeat + col));
    }
    out[idx] = v;
  }
}

}  // namespace fa

extern "C" int fa_fbank_lfr_cmvn_strided(const float* wav, const int32_t* wav_lens, int32_t batch, int64_t wav_stride,
                                         const float* cmvn, const float* mel_banks, const float* window, float* feats,
                                         int64_t feats_batch_stride_rows, int32_t* feat_lens, int32_t t_max,
                                         fa_stream_t stream) {
  if (!wav || !wav_lens || !mel_banks || !window || !feats || !feat_lens || batch <= 0 || t_max <= 0 ||
      feats_batch_stride_rows < t_max)
    return FA_ERR_ARG;
  const size_t smem = sizeof(fa::FbankSmem);
  static fa::PerDeviceOnce once;
  FA_RETURN_IF_ERR(fa::ensure_dyn_smem(fa::fbank_lfr_cmvn_kernel, smem, once));
  dim3 grid((t_max + fa::kRows - 1) / fa::kRows, batch);
  fa::fbank_lfr_cmvn_kernel<<<grid, fa::kWarps * 32, smem, (cudaStream_t)stream>>>(
      wav, wav_lens, wav_stride, cmvn, mel_banks, window, feats, feat_lens, t_max, feats_batch_stride_rows);
  FA_CHECK_LAUNCH();
  return FA_OK;
}

extern "C" int fa_fbank_lfr_cmvn(const float* wav, const int32_t* wav_lens, int32_t batch, int64_t wav_stride,
                                 const float* cmvn, const float* mel_banks, const float* window, float* feats,
                                 int32_t* feat_lens, int32_t t_max, fa_stream_t stream) {
  return fa_fbank_lfr_cmvn_strided(wav, wav_lens, batch, wav_stride, cmvn, mel_banks, window, feats, t_max, feat_lens, t_max, stream);
}
