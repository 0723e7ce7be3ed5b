// Internal launch helpers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fa { typedef __half plane_t; }   // 16-bit operand plane element (tc_common.cuh)

namespace fa {

// y (fp32) and/or planes (fp16 [nplanes][rows][cols_pad], the A operand of a following tcgen05 GEMM)
int layernorm_launch(const float* x, int64_t rows, const FaNorm& nm, float* y, const float* pe_inv, float xscale,
                     int rows_per_batch, cudaStream_t st, plane_t* planes = nullptr, int nplanes = 0, int cols_pad = 0,
                     float* emb_out = nullptr);   // emb_out: with pe_inv, also write the embedded (pre-norm) rows there
int gemm_f32_launch(const float* A, int64_t lda, int64_t M, const float* W, int N, int K, const float* bias, int relu,
                    const float* r1, int64_t ldr1, const float* r2, int64_t ldr2, float* C, int64_t ldc,
                    cudaStream_t st);
// tcgen05 fp16-split GEMM (gemm_tc.cu)
size_t gemm_tc_scratch_bytes(int64_t max_rows, int max_k, int mode);
int gemm_tc_launch(const float* x, int64_t ldx, int64_t rows, const FaLinear& lin, int relu, const float* r1,
                   int64_t ld1, const float* r2, int64_t ld2, float* y, int64_t ldy, int mode, Arena* scratch,
                   cudaStream_t st);
// Column-range sinks of a GEMM epilogue feeding the tensor-core attention (gemm_tc.cu): columns [q0, q0+width) -> q planes
// (scaled), [k0, k0+width) -> k planes, [v0, v0+width) -> transposed v planes (+ fp32 into the GEMM's C when C != null).
// A range is disabled by placing it outside [0, N) (e.g. -1000000).
struct AttnSinks {
  int enabled = 0;
  int q0 = -1000000, k0 = -1000000, v0 = -1000000;
  int width = 512;             // heads * 128
  int npl = 2;                 // planes written
  int t_rows = 1, t_pad = 64;  // rows per utterance (v rows = keys), padded key pitch of the transposed planes
  float qscale = 1.f;
  plane_t* q_planes = nullptr;   // [npl][M][width]
  plane_t* k_planes = nullptr;   // [npl][M][width]
  plane_t* vt_planes = nullptr;  // [npl][B*width][t_pad]
};
// tcgen05 attention (attention_tc.cu); ctx fp32 and/or fp16 planes [npl][B*tq][ldp]
int attention_tc_planes_launch(const plane_t* qp, const plane_t* kp, const plane_t* vt, const int32_t* key_lens,
                               int batch, int heads, int tq, int tk, float* ctx, int64_t ldc, plane_t* ctx_planes,
                               int64_t ldp, int out_nplanes, int mode, cudaStream_t st, int kv_shared = 0);
size_t attention_tc_scratch_bytes(int batch, int heads, int tq, int tk, int mode);
int attention_tc_launch(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        const int32_t* key_lens, int batch, int heads, int tq, int tk, float* ctx, int64_t ldc,
                        plane_t* ctx_planes, int64_t ldp, int out_nplanes, int mode, Arena* scratch, cudaStream_t st,
                        int kv_shared = 0);   // kv_shared: k / v hold ONE batch entry that every utterance attends over
int gemm_tc_planes_launch(const plane_t* a_planes, int64_t M, const FaLinear& lin, int relu, const float* r1, int64_t ld1,
                          const float* r2, int64_t ld2, float* y, int64_t ldy, plane_t* out_planes, int64_t ldo,
                          int mode, cudaStream_t st, const AttnSinks* att = nullptr, int64_t a_ld = 0, int64_t a_plane_rows = 0);
// a_ld / a_plane_rows (0 = dense: K_pad / M): row pitch of the A planes and rows between planes when A is an overlapping view
int split_rows_launch(const float* x, int64_t ldx, int64_t rows, int cols, int cols_pad, int nplanes, plane_t* planes,
                      cudaStream_t st);
int attention_f32_launch(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         const int32_t* key_lens, int batch, int heads, int tq, int tk, float* ctx, int64_t ldc,
                         cudaStream_t st, int kv_shared = 0);
int attention_small_launch(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const int32_t* key_lens,
                           int batch, int heads, int head_dim, int tq, int tk, float* ctx, int64_t ldc, cudaStream_t st, int kv_shared = 0);
int fsmn_launch(const float* v, int64_t ldv, const int32_t* lens, int batch, int t_max, int channels, const float* w,
                int ksize, const float* res, int64_t ldr, float* out, int64_t ldo, cudaStream_t st, int causal = 0);
int cif_im2col_launch(const float* enc, int64_t rows, int t_max, int d, float* xc, cudaStream_t st);
int cif_fire_loop_launch(const float* enc, const float* alpha_rows, const int32_t* lens, int batch, int t_max, int d,
                         float tail, float threshold, float* acoustic, int n_cap, int32_t* token_num, float* alphas, float* peaks,
                         cudaStream_t st);
int cif_upsample_scan_launch(float* alphas2, const int32_t* token_num, int batch, int t3, float thr, float* us_peaks, cudaStream_t st);
int cif_alpha_launch(const float* c, int d, const float* w, const float* b0, const int32_t* lens, int t_max,
                     int64_t rows, float smooth, float noise, float* alpha_rows, cudaStream_t st, int c_rows_per_batch = 0);
int cif_pad_planes_launch(const float* enc, int batch, int t_max, int d, int nplanes, int64_t rows_alloc, plane_t* planes, cudaStream_t st);
int cif_fire_launch(const float* enc, const float* alpha_rows, const int32_t* lens, int batch, int t_max, int d,
                    float tail, float* acoustic, int n_cap, int32_t* token_num, float* alphas, float* peaks,
                    cudaStream_t st);
int ctc_filter_launch(const int32_t* ids, const int32_t* lens, int batch, int t_max, int blank, int32_t* out_ids,
                      int32_t* out_lens, cudaStream_t st);
int argmax_lse_launch(float* logits, int64_t rows, int vocab, int64_t ld, int32_t* ids, float* best_logp,
                      int write_log_softmax, cudaStream_t st);

}  // namespace fa
