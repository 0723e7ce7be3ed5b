/* The OfflineStream surface of FunASR's C++ runtime — runtime/onnxruntime/include/funasrruntime.h:21-77,100-116 — with the SAME
 * names, argument lists and defaults, implemented over this library's handle API (funasr_b200.h: fa_offline_*).  A server written
 * against funasrruntime.h (runtime/websocket, runtime/http, bin/funasr-onnx-offline.cpp) links against libfunasr_b200.so with this
 * header in place of the original for the offline ASR calls it makes.  Differences, all at run time, none in the signatures:
 *   model_path["model-dir"] names a directory holding `model.fab2` (funasr_b200/pack.py, written from an unmodified model.pt +
 *   am.mvn) and optionally `tokens.txt` (one token per line; without it results carry the token ids in decimal);
 *   optional keys "gemm-mode" (fp32 | fp16 | fp16x3 | fp16x6, default fp16x3) and "gpu-id" (default 0);
 *   there is no CPU path (use_gpu is ignored); audio must be 16 kHz; the WFST / LM decoder entry points are accepted and ignored
 *   (greedy decoding, like the reference without --lm-dir).
 */
#pragma once
#include <stdint.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#define _FUNASRAPI

typedef void* FUNASR_HANDLE;
typedef void* FUNASR_RESULT;
typedef void* FUNASR_DEC_HANDLE;
typedef unsigned char FUNASR_BOOL;

#define FUNASR_TRUE 1
#define FUNASR_FALSE 0
#define QM_DEFAULT_THREAD_NUM 4

typedef enum { RASR_NONE = -1, RASRM_CTC_GREEDY_SEARCH = 0, RASRM_CTC_RPEFIX_BEAM_SEARCH = 1, RASRM_ATTENSION_RESCORING = 2 } FUNASR_MODE;
typedef enum { ASR_OFFLINE = 0, ASR_ONLINE = 1, ASR_TWO_PASS = 2 } ASR_TYPE;

typedef void (*QM_CALLBACK)(int cur_step, int n_total);

// OfflineStream (funasrruntime.h:100-116)
_FUNASRAPI FUNASR_HANDLE FunOfflineInit(std::map<std::string, std::string>& model_path, int thread_num, bool use_gpu = false, int batch_size = 1);
_FUNASRAPI void FunOfflineReset(FUNASR_HANDLE handle, FUNASR_DEC_HANDLE dec_handle = nullptr);
_FUNASRAPI FUNASR_RESULT FunOfflineInferBuffer(FUNASR_HANDLE handle, const char* sz_buf, int n_len, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                                               const std::vector<std::vector<float>>& hw_emb, int sampling_rate = 16000,
                                               std::string wav_format = "pcm", bool itn = true, FUNASR_DEC_HANDLE dec_handle = nullptr,
                                               std::string svs_lang = "auto", bool svs_itn = true);
_FUNASRAPI FUNASR_RESULT FunOfflineInfer(FUNASR_HANDLE handle, const char* sz_filename, FUNASR_MODE mode, QM_CALLBACK fn_callback,
                                         const std::vector<std::vector<float>>& hw_emb, int sampling_rate = 16000, bool itn = true,
                                         FUNASR_DEC_HANDLE dec_handle = nullptr);
_FUNASRAPI const std::vector<std::vector<float>> CompileHotwordEmbedding(FUNASR_HANDLE handle, std::string& hotwords, ASR_TYPE mode = ASR_OFFLINE);
_FUNASRAPI void FunOfflineUninit(FUNASR_HANDLE handle);

// result accessors shared with the other streams (funasrruntime.h:67-77)
_FUNASRAPI const char* FunASRGetResult(FUNASR_RESULT result, int n_index);
_FUNASRAPI const char* FunASRGetStamp(FUNASR_RESULT result);
_FUNASRAPI const char* FunASRGetStampSents(FUNASR_RESULT result);
_FUNASRAPI const int FunASRGetRetNumber(FUNASR_RESULT result);
_FUNASRAPI void FunASRFreeResult(FUNASR_RESULT result);
_FUNASRAPI const float FunASRGetRetSnippetTime(FUNASR_RESULT result);

// WFST decoder (funasrruntime.h:134-138): accepted, no effect (greedy decoding)
_FUNASRAPI FUNASR_DEC_HANDLE FunASRWfstDecoderInit(FUNASR_HANDLE handle, int asr_type, float glob_beam, float lat_beam, float am_scale);
_FUNASRAPI void FunASRWfstDecoderUninit(FUNASR_DEC_HANDLE handle);
_FUNASRAPI void FunWfstDecoderLoadHwsRes(FUNASR_DEC_HANDLE handle, int inc_bias, std::unordered_map<std::string, int>& hws_map);
_FUNASRAPI void FunWfstDecoderUnloadHwsRes(FUNASR_DEC_HANDLE handle);

// extension (not in funasrruntime.h): why the last call on this thread failed
const char* FunB200LastError();
