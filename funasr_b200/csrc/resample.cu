// Polyphase sinc resampler: the GPU counterpart of torchaudio.functional.resample as FunASR's loader applies it when the input
// rate differs from the model's (funasr/utils/load_utils.py:176-178, torchaudio.transforms.Resample defaults: sinc_interp_hann,
// lowpass_filter_width 6, rolloff 0.99).  torchaudio pads the waveform by (width, width + orig) zeros and runs a conv1d with
// `new` output channels and stride `orig` (orig / new already divided by their gcd); output sample n = i*new + j is therefore
//   y[n] = sum_k x[i*orig + k - width] * kernel[j][k],   k < 2*width + orig,   x = 0 outside [0, len),
// truncated to ceil(new * len / orig) samples.  The kernel table is computed on the host (funasr_b200/resample.py restates
// torchaudio's _get_sinc_resample_kernel).  One thread per output sample, table rows through the read-only cache.
#include "common.cuh"

namespace fa {

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ x, const int32_t* __restrict__ lens, int64_t x_stride, const float* __restrict__ table,
                int orig, int nnew, int width, int taps, float* __restrict__ y, int64_t y_stride, int y_cap, int32_t* __restrict__ out_lens) {
  const int b = blockIdx.y;
  const int len = lens[b];
  const int64_t out_len64 = ((int64_t)nnew * len + orig - 1) / orig;         // ceil(new * len / orig)
  const int out_len = (int)(out_len64 < y_cap ? out_len64 : y_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) out_lens[b] = out_len;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= y_cap) return;
  float acc = 0.f;
  if (n < out_len) {
    const int i = n / nnew, j = n - i * nnew;
    const float* xr = x + (int64_t)b * x_stride;
    const float* tr = table + (int64_t)j * taps;
    const int base = i * orig - width;
    const int k_lo = base < 0 ? -base : 0;
    const int k_hi = min(taps, len - base);
    for (int k = k_lo; k < k_hi; ++k) acc = fmaf(__ldg(xr + base + k), __ldg(tr + k), acc);
  }
  y[(int64_t)b * y_stride + n] = acc;                                         // rows are zero filled beyond out_len
}

// Interleaved PCM frames -> mono fp32 in [-1, 1): the sample decode of the loader (funasr/utils/load_utils.py:48-179 ->
// torchaudio.load(normalize=True) semantics: u8 -> (x - 128) / 128, s16 -> x / 2^15, s24 (packed, little endian) -> x / 2^23,
// s32 -> x / 2^31, f32 unchanged; load_utils.py:168-170 / extract_fbank average the channels).  One thread per frame.
__global__ void __launch_bounds__(256)
pcm_decode_kernel(const unsigned char* __restrict__ pcm, int fmt, int channels, int64_t frames, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames) return;
  float acc = 0.f;
  for (int c = 0; c < channels; ++c) {
    const int64_t k = i * channels + c;
    float v;
    switch (fmt) {
      case 0: v = reinterpret_cast<const float*>(pcm)[k]; break;
      case 1: v = (float)reinterpret_cast<const int16_t*>(pcm)[k] * (1.0f / 32768.0f); break;
      case 2: {
        const unsigned char* p = pcm + 3 * k;
        int32_t x = (int32_t)p[0] | ((int32_t)p[1] << 8) | ((int32_t)(signed char)p[2] << 16);
        v = (float)x * (1.0f / 8388608.0f);
        break;
      }
      case 3: v = (float)reinterpret_cast<const int32_t*>(pcm)[k] * (1.0f / 2147483648.0f); break;
      default: v = ((float)pcm[k] - 128.0f) * (1.0f / 128.0f); break;
    }
    acc += v;
  }
  out[i] = channels > 1 ? acc / (float)channels : acc;
}

}  // namespace fa

extern "C" int fa_pcm_decode(const void* pcm, int32_t sample_format, int32_t channels, int64_t frames, float* out, fa_stream_t stream) {
  if (!pcm || !out || sample_format < 0 || sample_format > 4 || channels < 1 || channels > 64 || frames < 0) return FA_ERR_ARG;
  if (frames == 0) return FA_OK;
  fa::pcm_decode_kernel<<<(unsigned)((frames + 255) / 256), 256, 0, (cudaStream_t)stream>>>(static_cast<const unsigned char*>(pcm), sample_format, channels,
                                                                                          frames, out);
  FA_CHECK_LAUNCH();
  return FA_OK;
}

extern "C" int fa_resample(const float* x, const int32_t* lens, int32_t batch, int64_t x_stride, const float* table, int32_t orig,
                           int32_t nnew, int32_t width, float* y, int64_t y_stride, int32_t y_cap, int32_t* out_lens, fa_stream_t stream) {
  if (!x || !lens || !table || !y || !out_lens || batch <= 0 || orig <= 0 || nnew <= 0 || width <= 0 || y_cap <= 0 || y_stride < y_cap)
    return FA_ERR_ARG;
  const int taps = 2 * width + orig;
  dim3 grid((y_cap + 255) / 256, batch);
  fa::resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, lens, x_stride, table, orig, nnew, width, taps, y, y_stride, y_cap, out_lens);
  FA_CHECK_LAUNCH();
  return FA_OK;
}
