// FSMN memory block: out[b,t,c] = m[t] * ( v[t]m[t] + sum_j w[c,j] v[t+j-L]m[t+j-L] ) (+ res[b,t,c]),
// m[t] = 1[t < lens[b]], L = (k-1)/2, zero outside [0, t_max) — the depthwise Conv1d(groups=C, no bias)
// of MultiHeadedAttentionSANM.forward_fsmn (sanm/attention.py:216-239) and of the decoder's
// MultiHeadedAttentionSANMDecoder.forward (:583-631); `res` fuses DecoderLayerSANM's `residual + x`
// (paraformer/decoder.py:107).  HBM-bound (reads v once + halo, writes once): channels are the
// coalesced axis, each thread slides a k-wide register window down a strip of time steps.
#include "common.cuh"

namespace fa {

constexpr int FSMN_TT = 32;     // time steps per thread strip
constexpr int FSMN_KMAX = 31;

// K taps, the window of output t covers inputs t - L .. t - L + K - 1: L = (K-1)/2 is the centred SAN-M memory block, L = K - 1
// the causal memory of the FSMN-VAD encoder (FSMNBlock.conv_left over zero LEFT padding, fsmn_vad_streaming/encoder.py:136-160).
template <int K, int L>
__global__ void __launch_bounds__(128)
fsmn_kernel(const float* __restrict__ v, int64_t ldv, const int32_t* __restrict__ lens, int t_max, int channels,
            const float* __restrict__ w, const float* __restrict__ res, int64_t ldr, float* __restrict__ out,
            int64_t ldo) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = blockIdx.y * FSMN_TT;
  const int b = blockIdx.z;
  pdl_wait();
  pdl_trigger();
  if (c >= channels) return;
  const int len = min(lens[b], t_max);
  float wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = __ldg(w + c * K + j);
  const float* vb = v + (int64_t)b * t_max * ldv + c;
  auto load = [&](int t) -> float { return (t >= 0 && t < len) ? __ldg(vb + (int64_t)t * ldv) : 0.f; };  // v*m, zero pad
  float win[K];
#pragma unroll
  for (int j = 0; j < K - 1; ++j) win[j + 1] = load(t0 - L + j);
  const int t_end = min(t0 + FSMN_TT, t_max);
  for (int t = t0; t < t_end; ++t) {
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = win[j + 1];
    win[K - 1] = load(t + K - 1 - L);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fmaf(wk[j], win[j], acc);
    float o = t < len ? __fadd_rn(acc, win[L]) : 0.f;     // (conv + inputs) * mask
    const int64_t row = (int64_t)b * t_max + t;
    if (res) o = __fadd_rn(__ldg(res + row * ldr + c), o);
    out[row * ldo + c] = o;
  }
}

int fsmn_launch(const float* v, int64_t ldv, const int32_t* lens, int batch, int t_max, int channels, const float* w,
                int ksize, const float* res, int64_t ldr, float* out, int64_t ldo, cudaStream_t st, int causal) {
  if (batch <= 0 || t_max <= 0) return FA_OK;
  if (!v || !lens || !w || !out) return FA_ERR_ARG;
  dim3 grid((channels + 127) / 128, (t_max + FSMN_TT - 1) / FSMN_TT, batch);
  if (causal) {
    if (ksize != 20) return FA_ERR_UNSUPPORTED;
    FA_CUDA_OK(launch_pdl(fsmn_kernel<20, 19>, grid, dim3(128), 0, st, 1, v, ldv, lens, t_max, channels, w, res, ldr, out, ldo));
    FA_CHECK_LAUNCH();
    return FA_OK;
  }
  switch (ksize) {
    case 11: FA_CUDA_OK(launch_pdl(fsmn_kernel<11, 5>, grid, dim3(128), 0, st, 1, v, ldv, lens, t_max, channels, w, res, ldr, out, ldo)); break;
    case 21: FA_CUDA_OK(launch_pdl(fsmn_kernel<21, 10>, grid, dim3(128), 0, st, 1, v, ldv, lens, t_max, channels, w, res, ldr, out, ldo)); break;
    case 31: FA_CUDA_OK(launch_pdl(fsmn_kernel<31, 15>, grid, dim3(128), 0, st, 1, v, ldv, lens, t_max, channels, w, res, ldr, out, ldo)); break;
    default: return FA_ERR_UNSUPPORTED;
  }
  FA_CHECK_LAUNCH();
  return FA_OK;
}

}  // namespace fa

extern "C" int fa_fsmn(const float* v, int64_t ldv, const int32_t* lens, int32_t batch, int32_t t_max, int32_t channels,
                       const float* w, int32_t ksize, const float* res, int64_t ld_res, float* out, int64_t ld_out,
                       fa_stream_t stream) {
  return fa::fsmn_launch(v, ldv, lens, batch, t_max, channels, w, ksize, res, ld_res, out, ld_out, (cudaStream_t)stream, 0);
}
