// Small host-side (no CUDA) sequential routines of the timestamp post-processing, in the library so that the per-utterance host
// work keeps up with the GPU (64 utterances per ~40 ms step).
#include "../../include/funasr_b200.h"

// Integrate-and-fire trace of one utterance, funasr/utils/timestamp_tools.py:14-34 (`cif_wo_hidden`): fp32 running sum of the
// weights, reduced by `threshold` right after every frame where it reaches it; trace[t] holds the value BEFORE the reduction.
// ts_prediction_lfr6_standard re-integrates the renormalised weights with it whenever the fire count differs from tokens + 1 (:67-72)
// — for BiCif / SeACo Paraformer that is every utterance (the head fires once per token).  Plain fp32 adds in program order
// (compiled with -ffp-contract=off), identical to the numpy loop in funasr_b200/timestamps.py (1 ms per 1500 frames there).
extern "C" int fa_cif_wo_hidden_host(const float* alphas, int64_t n, float threshold, float* trace) {
  if (n < 0 || (n > 0 && (!alphas || !trace))) return FA_ERR_ARG;
  float level = 0.0f;
  for (int64_t t = 0; t < n; ++t) {
    level = level + alphas[t];
    trace[t] = level;
    if (level >= threshold) level = level - 1.0f * threshold;
  }
  return FA_OK;
}
