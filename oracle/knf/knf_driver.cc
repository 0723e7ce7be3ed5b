// TEST INFRASTRUCTURE (oracle side, never linked into the product).
// Thin C-ABI driver over the reference's vendored kaldi-native-fbank, compiled FROM the reference tree where it lies
// (oracle/knf/Makefile -> oracle/_ref/libknf_ref.so).  It configures the extractor exactly as the reference's C++
// runtime does for the offline Paraformer (runtime/onnxruntime/src/paraformer.cpp:24-31: dither 0, 80 bins, 16 kHz,
// hamming, 25 ms / 10 ms, energy_floor 0) and feeds it like Paraformer::FbankKaldi (paraformer.cpp:298-312: samples
// scaled by 32768, one AcceptWaveform call).  Second, independent pin of the Fbank arithmetic next to torchaudio.
#include <cstdint>
#include <vector>

#include "kaldi-native-fbank/csrc/online-feature.h"

extern "C" {

// -> number of frames written (<= max_frames), or the number of frames ready when out == nullptr
int knf_ref_fbank(const float* wav, int n_samples, float sample_rate, int n_mels, float frame_length_ms, float frame_shift_ms,
                  float* out, int max_frames) {
  knf::FbankOptions o;
  o.frame_opts.dither = 0;
  o.mel_opts.num_bins = n_mels;
  o.frame_opts.samp_freq = sample_rate;
  o.frame_opts.window_type = "hamming";
  o.frame_opts.frame_shift_ms = frame_shift_ms;
  o.frame_opts.frame_length_ms = frame_length_ms;
  o.energy_floor = 0;
  o.mel_opts.debug_mel = false;
  knf::OnlineFbank fb(o);
  std::vector<float> buf(n_samples);
  for (int i = 0; i < n_samples; ++i) buf[i] = wav[i] * 32768;
  fb.AcceptWaveform(sample_rate, buf.data(), static_cast<int32_t>(buf.size()));
  int frames = fb.NumFramesReady();
  if (!out) return frames;
  if (frames > max_frames) frames = max_frames;
  for (int i = 0; i < frames; ++i) {
    const float* f = fb.GetFrame(i);
    for (int j = 0; j < n_mels; ++j) out[(int64_t)i * n_mels + j] = f[j];
  }
  return frames;
}

}  // extern "C"
