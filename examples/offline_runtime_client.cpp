// The call sequence of FunASR's own offline client (runtime/onnxruntime/bin/funasr-onnx-offline.cpp:113-242) against this
// library: FunOfflineInit -> FunASRWfstDecoderInit -> CompileHotwordEmbedding -> FunOfflineInfer (file) / FunOfflineInferBuffer
// -> FunASRGetResult / FunASRGetStamp / FunASRGetRetSnippetTime -> FunASRFreeResult -> ... -> FunOfflineUninit.
// Build (the header can be the reference's own funasrruntime.h: the signatures are identical):
//   g++ -std=c++17 -DFUNASR_RUNTIME_HEADER='"funasrruntime_b200.h"' -Iinclude examples/offline_runtime_client.cpp -Lfunasr_b200 -lfunasr_b200
// usage: offline_runtime_client <model-dir> <audio.wav|audio.pcm> [gemm-mode] [hotwords separated by spaces]
#ifndef FUNASR_RUNTIME_HEADER
#define FUNASR_RUNTIME_HEADER "funasrruntime_b200.h"
#endif
#include <stdint.h>
#include <stdio.h>
#include <fstream>
#include <sstream>
#include <string>      // before the runtime header: funasrruntime.h uses std::string without including <string> itself
#include FUNASR_RUNTIME_HEADER

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <model-dir> <audio.wav|audio.pcm> [gemm-mode] [hotwords]\n", argv[0]); return 2; }
  std::map<std::string, std::string> model_path;
  model_path.insert({"model-dir", argv[1]});
  model_path.insert({"quantize", "false"});
  if (argc > 3) model_path.insert({"gemm-mode", argv[3]});
  FUNASR_HANDLE asr_handle = FunOfflineInit(model_path, 1, true, 1);
  if (!asr_handle) { fprintf(stderr, "FunASR init failed\n"); return 1; }
  FUNASR_DEC_HANDLE decoder_handle = FunASRWfstDecoderInit(asr_handle, ASR_OFFLINE, 3.0f, 3.0f, 10.0f);
  std::unordered_map<std::string, int> hws_map;
  FunWfstDecoderLoadHwsRes(decoder_handle, 20, hws_map);
  std::string nn_hotwords = argc > 4 ? argv[4] : "";
  std::vector<std::vector<float>> hotwords_embedding = CompileHotwordEmbedding(asr_handle, nn_hotwords);
  printf("hotword_rows %zu\n", hotwords_embedding.size());
  float snippet_time = 0.f;
  // 1) the file entry point
  FUNASR_RESULT result = FunOfflineInfer(asr_handle, argv[2], RASR_NONE, nullptr, hotwords_embedding, 16000, true, decoder_handle);
  if (!result) { fprintf(stderr, "no return data!\n"); return 1; }
  printf("file_result %s\n", FunASRGetResult(result, 0));
  printf("stamp [%s] sents [%s] n %d\n", FunASRGetStamp(result), FunASRGetStampSents(result), FunASRGetRetNumber(result));
  snippet_time += FunASRGetRetSnippetTime(result);
  FunASRFreeResult(result);
  // 2) the buffer entry point with the same bytes
  std::ifstream f(argv[2], std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string bytes = ss.str();
  const std::string name = argv[2];
  const bool wav = name.size() > 4 && name.compare(name.size() - 4, 4, ".wav") == 0;
  result = FunOfflineInferBuffer(asr_handle, bytes.data(), (int)bytes.size(), RASR_NONE, nullptr, hotwords_embedding, 16000, wav ? "wav" : "pcm", true,
                                 decoder_handle);
  if (!result) { fprintf(stderr, "no return data!\n"); return 1; }
  printf("buffer_result %s\n", FunASRGetResult(result, 0));
  snippet_time += FunASRGetRetSnippetTime(result);
  FunASRFreeResult(result);
  printf("audio_seconds %.3f\n", snippet_time);
  FunWfstDecoderUnloadHwsRes(decoder_handle);
  FunASRWfstDecoderUninit(decoder_handle);
  FunOfflineUninit(asr_handle);
  return 0;
}
