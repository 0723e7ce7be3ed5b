// TEST INFRASTRUCTURE (oracle side, never linked into the product).
// C-ABI driver over the end-point detector of the reference's C++ runtime (runtime/onnxruntime/src/e2e-vad.h, header-only),
// compiled FROM the reference tree (oracle/knf/Makefile).  Called the way FsmnVad::Infer calls it
// (runtime/onnxruntime/src/fsmn-vad.cpp:245-249): one E2EVadModel per recording, all scores and the whole waveform at once,
// is_final = true, offline mode.  An independent implementation of the state machine that funasr_b200/vad.py restates from the
// Python model (funasr/models/fsmn_vad_streaming/model.py); it has no dynamic end-silence schedule, so it is compared on
// configurations with a fixed max_end_silence_time.
#include <cstdint>
#include <vector>

#include "e2e-vad.h"

extern "C" {

// sil_prob[frames]: posterior of the silence pdf (pdf id 0); wav[n_samples] in [-1, 1].  Writes up to max_segments {start_ms, end_ms}
// pairs into segments and returns how many the detector produced.
int vad_ref_segments(const float* sil_prob, int frames, const float* wav, int n_samples, int max_end_silence_ms, int max_single_segment_ms,
                     float speech_noise_thres, int sample_rate, int32_t* segments, int max_segments) {
  std::vector<std::vector<float>> scores(frames, std::vector<float>(1));
  for (int i = 0; i < frames; ++i) scores[i][0] = sil_prob[i];
  std::vector<float> w(wav, wav + n_samples);
  funasr::E2EVadModel vad;
  std::vector<std::vector<int>> seg = vad(scores, w, true, false, max_end_silence_ms, max_single_segment_ms, speech_noise_thres, sample_rate);
  int n = 0;
  for (const auto& s : seg) {
    if (n < max_segments) { segments[2 * n] = s[0]; segments[2 * n + 1] = s[1]; }
    ++n;
  }
  return n;
}

}  // extern "C"
