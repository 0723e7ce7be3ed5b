// Persistent bidirectional LSTM recurrence (one layer, batch_first, no packing) — the BLSTM of CifPredictorV3's timestamp head
// (funasr/models/bicif_paraformer/cif_predictor.py:187-190, :318-320: `output2, _ = self.blstm(output2)` on the 3x upsampled
// encoder output).  cuDNN needs 34 ms for [64, 1500, 512] in fp32; the recurrence is 1500 strictly sequential steps of a
// [B,512] x [512,2048] product per direction, so the design goal is the shortest possible step:
//   * the input projections x W_ih^T + b_ih + b_hh of ALL steps are one tcgen05 GEMM of this library (fa_linear), outside;
//   * weight-stationary recurrence: 2 directions x 64 CTAs, CTA c keeps the 4 gate rows of hidden units [8c, 8c+8) of W_hh
//     (32 x 512 fp32 = 64 KB) in shared memory for the whole sequence, plus one [64, 512] fp32 copy of h_{t-1} (128 KB);
//   * per step: gather h_{t-1} (written by the 64 CTAs of this direction into the OUTPUT tensor itself) -> 64 x 32 dot products
//     of length 512 in fp32 (thread = one sequence x two hidden units x four gates, cell state in registers) -> gates, c, h ->
//     write the h slice -> per-direction grid barrier (monotone atomic counter).
// Exact fp32 arithmetic (no tensor cores): the timestamps are thresholded downstream.  Launched cooperatively so all 128 CTAs
// are co-resident (the barrier would deadlock otherwise).
#include "common.cuh"

namespace fa {

constexpr int LS_H = 512;          // hidden size
constexpr int LS_UNITS = 8;        // hidden units per CTA
constexpr int LS_NC = LS_H / LS_UNITS;   // 64 CTAs per direction
constexpr int LS_ROWS = 4 * LS_UNITS;    // 32 gate rows per CTA
constexpr int LS_BT = 64;          // sequences per batch tile (256 threads = 64 sequences x 4 unit pairs)
constexpr int LS_WLD = LS_H + 4;   // padded weight row pitch (floats): rows of different unit pairs hit different banks
constexpr int LS_HLD = LS_H + 4;   // padded h row pitch: the 8 sequences of a warp read their float4 from 8 x 4 distinct banks

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256, 1)
blstm_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh_f, const float* __restrict__ w_hh_b, int batch, int T,
             float* __restrict__ out, unsigned int* __restrict__ counters) {
  extern __shared__ __align__(16) float smf[];
  float* sW = smf;                                  // [32][LS_WLD]
  float* sH = smf + LS_ROWS * LS_WLD;               // [LS_BT][LS_HLD]
  const int dir = blockIdx.x / LS_NC, c = blockIdx.x % LS_NC;
  const int tid = threadIdx.x;
  const float* whh = dir == 0 ? w_hh_f : w_hh_b;
  // local row r = g*8 + u  <->  W_hh row g*512 + (8c + u)   (PyTorch gate order i, f, g, o)
  for (int idx = tid; idx < LS_ROWS * (LS_H / 4); idx += blockDim.x) {
    const int r = idx / (LS_H / 4), k4 = idx % (LS_H / 4);
    const int g = r / LS_UNITS, u = r % LS_UNITS;
    const float4 w = __ldg(reinterpret_cast<const float4*>(whh + ((int64_t)g * LS_H + c * LS_UNITS + u) * LS_H) + k4);
    *reinterpret_cast<float4*>(sW + r * LS_WLD + 4 * k4) = w;
  }
  const int bl = tid >> 2, q = tid & 3;              // sequence within the tile, unit pair {2q, 2q+1}
  const int n_tiles = (batch + LS_BT - 1) / LS_BT;   // <= 4 (checked by the launcher)
  float cst[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) { cst[i][0] = 0.f; cst[i][1] = 0.f; }
  __syncthreads();
  unsigned int* counter = counters + dir;
  const int64_t out_ld = 2 * LS_H;                   // out [B, T, 2H]
  const int64_t xp_ld = 2 * 4 * LS_H;                // xproj [B*T, 2 dirs x 4H]
  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    const int tp = dir == 0 ? t - 1 : t + 1;         // time index of h_{t-1} in processing order
    // x projections (+ both biases) of this thread's 8 gate rows for every batch tile: independent of h, so they are fetched
    // before the barrier wait and their latency hides behind it
    float2 xin[4][4];
#pragma unroll
    for (int bt = 0; bt < 4; ++bt) {
      const int b = bt * LS_BT + bl;
      if (bt < n_tiles && b < batch) {
        const float* xp = xproj + ((int64_t)b * T + t) * xp_ld + dir * 4 * LS_H + c * LS_UNITS + 2 * q;
#pragma unroll
        for (int g = 0; g < 4; ++g) xin[bt][g] = __ldg(reinterpret_cast<const float2*>(xp + g * LS_H));
      }
    }
    if (step > 0) {
      // every CTA of this direction has published its slice of h for the previous step
      if (tid == 0) {
        const unsigned int want = (unsigned int)step * LS_NC;
        while (*reinterpret_cast<volatile unsigned int*>(counter) < want) { }
        __threadfence();
      }
      __syncthreads();
    }
#pragma unroll
    for (int bt = 0; bt < 4; ++bt) {                 // compile-time trip count keeps the cell states in registers
      if (bt >= n_tiles) break;
      const int b0 = bt * LS_BT;
      const int nb = min(LS_BT, batch - b0);
      if (step > 0) {
        // gather h_{t-1} [nb, 512] of this direction, L2 -> shared memory with 16-byte async copies (cp.async.cg bypasses L1:
        // the lines were written by other SMs during this launch); all 32 copies of a thread are in flight at once
        const float* src0 = out + ((int64_t)b0 * T + tp) * out_ld + dir * LS_H;
        for (int idx = tid; idx < nb * (LS_H / 4); idx += blockDim.x) {
          const int b = idx / (LS_H / 4), k4 = idx % (LS_H / 4);
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sH + b * LS_HLD + 4 * k4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src0 + (int64_t)b * T * out_ld + 4 * k4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const int b = b0 + bl;
      if (bl < nb) {
        float acc[4][2];
#pragma unroll
        for (int g = 0; g < 4; ++g) { acc[g][0] = xin[bt][g].x; acc[g][1] = xin[bt][g].y; }
        if (step > 0) {
          const float* hrow = sH + bl * LS_HLD;
          float dot[4][2];
#pragma unroll
          for (int g = 0; g < 4; ++g) { dot[g][0] = 0.f; dot[g][1] = 0.f; }
#pragma unroll 4
          for (int k = 0; k < LS_H; k += 4) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrow + k);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const float4 w4 = *reinterpret_cast<const float4*>(sW + (g * LS_UNITS + 2 * q + u) * LS_WLD + k);
                float d = dot[g][u];
                d = fmaf(h4.x, w4.x, d); d = fmaf(h4.y, w4.y, d); d = fmaf(h4.z, w4.z, d); d = fmaf(h4.w, w4.w, d);
                dot[g][u] = d;
              }
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) { acc[g][0] += dot[g][0]; acc[g][1] += dot[g][1]; }
        }
        float hv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float ig = sigmoidf_(acc[0][u]), fg = sigmoidf_(acc[1][u]), gg = tanhf(acc[2][u]), og = sigmoidf_(acc[3][u]);
          const float cn = fg * cst[bt][u] + ig * gg;
          cst[bt][u] = cn;
          hv[u] = og * tanhf(cn);
        }
        *reinterpret_cast<float2*>(out + ((int64_t)b * T + t) * out_ld + dir * LS_H + c * LS_UNITS + 2 * q) = make_float2(hv[0], hv[1]);
      }
      __syncthreads();                               // sH is reused by the next batch tile
    }
    // publish: all threads' stores -> device scope, then one arrival per CTA
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicAdd(counter, 1u);
  }
}

int blstm_launch(const float* xproj, const float* w_hh_f, const float* w_hh_b, int batch, int T, int hidden, float* out,
                 unsigned int* counters, cudaStream_t st) {
  if (batch <= 0 || T <= 0) return FA_OK;
  if (!xproj || !w_hh_f || !w_hh_b || !out || !counters) return FA_ERR_ARG;
  if (hidden != LS_H || batch > 4 * LS_BT) return FA_ERR_UNSUPPORTED;
  const size_t smem = (size_t)(LS_ROWS * LS_WLD + LS_BT * LS_HLD) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    FA_CUDA_OK(cudaFuncSetAttribute(blstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  FA_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned int), st));
  void* args[] = {(void*)&xproj, (void*)&w_hh_f, (void*)&w_hh_b, (void*)&batch, (void*)&T, (void*)&out, (void*)&counters};
  FA_CUDA_OK(cudaLaunchCooperativeKernel((const void*)blstm_kernel, dim3(2 * LS_NC), dim3(256), args, smem, st));
  count_launch();
  return FA_OK;
}

}  // namespace fa

// One-layer bidirectional LSTM over [B, T, 512] given the input projections of both directions:
//   xproj [B*T, 2*2048] = x W_ih^T + b_ih + b_hh, columns [0,2048) forward gates (i,f,g,o), [2048,4096) reverse.
extern "C" int fa_blstm_forward(const float* xproj, const float* w_hh_fwd, const float* w_hh_bwd, int32_t batch, int32_t t_len,
                                int32_t hidden, float* out, void* sync_scratch8, fa_stream_t stream) {
  return fa::blstm_launch(xproj, w_hh_fwd, w_hh_bwd, batch, t_len, hidden, out, static_cast<unsigned int*>(sync_scratch8), (cudaStream_t)stream);
}
