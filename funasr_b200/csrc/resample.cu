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

}  // namespace fa

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
