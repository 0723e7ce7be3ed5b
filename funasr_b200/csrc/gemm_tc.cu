// tcgen05 / TMEM / TMA GEMM with fp16 operand splitting and a fused nn.Linear epilogue (sm_100a only).
//
//   Y[M,N] = act( sum_{(p,q) in terms} A_p[M,K] * W_q[N,K]^T + b ) (+ res1) (+ res2)
//
// A_p / W_q are the fp16 planes of the fp32 operands (x = hi + mid + lo, made by split kernels), accumulation is
// fp32 in TMEM.  Modes: F16X1 = {(0,0)} (fast), F16X3 = {(0,0),(0,1),(1,0)} (~2^-17 relative, the default
// parity mode on tensor cores), F16X6 = X3 + {(1,1),(0,2),(2,0)} (~fp32).  This replaces the torch.nn.Linear calls
// of the hot path (sanm/attention.py:256,306, transformer/positionwise_feed_forward.py:34,
// sanm/positionwise_feed_forward.py:33, paraformer/decoder.py:444, cif conv as GEMM).
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0 : TMA producer  — cp.async.bulk.tensor.2d (SWIZZLE_128B) of the A/W plane tiles into a smem ring
//   warp 1 : MMA issuer    — one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (128xBNx16) per K=16
//                            slice and per split term; tcgen05.commit releases ring slots / publishes accumulators
//   warp 2 : TMEM allocator (2 accumulators x BN columns, double buffered so the epilogue of tile i overlaps
//            the MMAs of tile i+1)
//   warps 4-7 : epilogue   — tcgen05.ld (32 lanes x 32 columns per warp and step) -> bias/ReLU/residuals ->
//            fp32 rows to HBM, or fp16 planes for a following GEMM.
// Tensor-pipe bound when operand tiles are reused from L2; roofline notes in DESIGN.md.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <stdlib.h>
#include <unordered_map>

namespace fa {

constexpr int TC_BM = 128;      // UMMA M (cta_group::1)
constexpr int TC_BK = 64;       // one 128-byte swizzle span of fp16
constexpr int TC_UK = 16;       // UMMA K for 16-bit inputs
constexpr uint32_t TC_TILE_BYTES_A = TC_BM * TC_BK * 2;   // 16 KB per plane tile

// ------------------------------------------------------------------------------------------------ kernel
struct TcParams {
  int64_t M;
  int N, Kp;            // Kp: K padded to a multiple of 64 (planes are zero padded)
  int64_t a_plane_rows; // rows between consecutive A planes in the 2D tensor map (= M)
  int w_plane_rows;     // = N
  int n_terms;          // 1, 3 or 6
  int relu;
  const float* bias;
  const float* r1; int64_t ldr1;
  const float* r2; int64_t ldr2;
  float* C; int64_t ldc;                 // fp32 output (or null)
  plane_t* out_planes;             // fp16 plane output [3][M][ldo] (or null)
  int64_t ldo; int out_nplanes;
  int tiles_m, tiles_n;
  // cta_group::2 kernel only — schedule of the 256 x 256 pair tiles over the P CTA pairs: full_rounds rounds of P whole tiles, then
  // the tail_tiles leftover tiles cut into tail_sub (1 | 2 | 4) column slices of 256 / tail_sub so that the last, partial wave
  // costs ceil(tail_tiles * tail_sub / P) / tail_sub of a round instead of a whole one (gemm_tc2_kernel, pick_tail_sub)
  int full_rounds, tail_tiles, tail_sub;
  AttnSinks att;                          // optional: route column ranges to attention operand planes
  // tcgen05 accumulates in fp32 with round-toward-zero: every 16-wide k-step shrinks the running sum by a fraction of an ulp, a
  // SYSTEMATIC relative error that grows linearly in K (tools/noise_probe.py, Gaussian data: -1.70e-6 at K = 512 and -6.35e-6 at
  // K = 2048 with the x3 split, identical with x6; -1.10e-6 at K = 512 single pass; the fp32 SIMT GEMM shows -3e-10).  The epilogue
  // multiplies the accumulator by 1 + (K / 16) * c (c = 5.3e-8 per k-step for the split modes, 3.4e-8 for one pass) before the
  // bias: the expected shrink is undone, what remains is the random part (~1e-6).  FA_RZ_COMP=0 disables it (A/B).
  float acc_scale;
};

static float rz_comp_scale(int kp, int n_terms) {
  static const bool on = [] { const char* e = getenv("FA_RZ_COMP"); return !(e && e[0] == '0'); }();
  if (!on) return 1.0f;
  return 1.0f + (float)(kp / 16) * (n_terms >= 3 ? 5.3e-8f : 3.4e-8f);
}

__constant__ int c_term_a[6] = {0, 0, 1, 1, 0, 2};
__constant__ int c_term_w[6] = {0, 1, 0, 1, 2, 0};

// Epilogue of one accumulator tile.  Eight epilogue warps per CTA (two per SM sub-partition, so one warp's TMEM / shared /
// global latencies are hidden by the other): warp w drains TMEM lane quarter w%4 and every second 16-column chunk.
//   phase 1  tcgen05.ld 32x32b.x16 (thread = row) -> raw fp32 accumulators into a padded shared-memory tile [32][20]
//   phase 2  re-read with the warp laid out as 8 rows x 4 float4 columns, so bias / residual loads and every store are
//            coalesced 16-byte (fp32) or 8-byte (fp16 plane) accesses; V columns of the attention sink take a
//            column-per-lane path that writes the per-head transposed planes as 4 consecutive keys (8 bytes) per store.
// History (profiles/README.md): v1 stored straight from the row-per-thread layout (4-byte stores to 32 lines per
// instruction); v2 staged through smem but kept 4 epilogue warps and measured SLOWER — ncu showed tensor pipe 17-25 %,
// warps_active 12 %, 12-36 cycles per issued instruction: the epilogue was instruction-latency bound, not store bound.
constexpr int EPI_CH = 16;                       // columns per chunk
constexpr int EPI_LD = 20;                       // padded row pitch (floats): 16-byte aligned, conflict-free float4 phases
constexpr int EPI_WARP_FLOATS = 32 * EPI_LD;     // 2.5 KB per epilogue warp
constexpr int EPI_WARPS = 8;

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = hi + mid + lo split of 4 values into fp16 planes: packed converts (cvt.rn.satfinite.f16x2.f32), packed unpack.  NPL is a compile-
// time constant: with a runtime plane count the loop compiled to a branchy 4x-unrolled body (ncu: 47 % of all executed
// instructions of the FFN-w_1 GEMM sat in this function).
template <int NPL>
__device__ __forceinline__ void store_planes4(plane_t* dst, int64_t plane_stride, float x0, float x1, float x2, float x3) {
#pragma unroll
  for (int pl = 0; pl < NPL; ++pl) {
    uint2 pk;
    pk.x = pack_planes2(x0, x1);
    pk.y = pack_planes2(x2, x3);
    *reinterpret_cast<uint2*>(dst) = pk;
    if (pl + 1 < NPL) {
      dst += plane_stride;
      const float2 a = unpack_planes2(pk.x), b = unpack_planes2(pk.y);
      x0 -= a.x; x1 -= a.y; x2 -= b.x; x3 -= b.y;
    }
  }
}

// per-head transposed V planes straight from the row-per-lane registers: for a fixed head dim the 32 lanes hold 32
// consecutive keys, so every 2-byte store instruction covers one 64-byte run
template <int NPL>
__device__ __forceinline__ void store_vt16(plane_t* dst, int64_t t_pad, int64_t plane, const uint32_t (&r)[16], const float* bias, float acc_scale) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    float x0 = __uint_as_float(r[j]) * acc_scale, x1 = __uint_as_float(r[j + 1]) * acc_scale;
    if (bias) { x0 += __ldg(bias + j); x1 += __ldg(bias + j + 1); }
    plane_t* d0 = dst + (int64_t)j * t_pad;
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl) {
      const uint32_t hb = pack_planes2(x0, x1);
      d0[0] = __ushort_as_half((unsigned short)(hb & 0xFFFFu));
      d0[t_pad] = __ushort_as_half((unsigned short)(hb >> 16));
      if (pl + 1 < NPL) { d0 += plane; const float2 a = unpack_planes2(hb); x0 -= a.x; x1 -= a.y; }
    }
  }
}

// The same through shared memory, for key counts that are a multiple of 4 (rows of one utterance then start at a multiple of 4,
// so groups of four consecutive keys stay inside one utterance and are 8-byte aligned in the transposed planes): the warp's
// 32 x 16 chunk is staged at pitch 17 (conflict free for the row-per-lane writes AND for the column reads below), then lane
// (g = lane & 7, c = lane >> 3) packs keys 4g..4g+3 of columns c, c+4, c+8, c+12 and stores 8 bytes per plane — 8 store
// instructions per lane and chunk instead of 32 two-byte ones (ncu of the QKV GEMM, round 2: the V third of the tiles spent
// ~180 instructions per chunk in store_vt16 and its tensor pipe sat at 74 % against 89 % for the plane-emitting FFN w_1).
constexpr int VT_LD = 17;
template <int NPL>
__device__ __forceinline__ void store_vt_staged(plane_t* __restrict__ dst_g /* this lane's key group, column 0 of the chunk */, int64_t t_pad,
                                                int64_t plane, const float (&x)[16], float* stage, int lane) {
  float* srow = stage + lane * VT_LD;
#pragma unroll
  for (int j = 0; j < 16; ++j) srow[j] = x[j];
  __syncwarp();
  const int g = lane & 7, c = lane >> 3;
  const float* sp = stage + (4 * g) * VT_LD + c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float k0 = sp[4 * i], k1 = sp[4 * i + VT_LD], k2 = sp[4 * i + 2 * VT_LD], k3 = sp[4 * i + 3 * VT_LD];
    plane_t* d = dst_g + (int64_t)(c + 4 * i) * t_pad;
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl) {
      uint2 pk;
      pk.x = pack_planes2(k0, k1);
      pk.y = pack_planes2(k2, k3);
      *reinterpret_cast<uint2*>(d) = pk;
      if (pl + 1 < NPL) {
        d += plane;
        const float2 a = unpack_planes2(pk.x), b = unpack_planes2(pk.y);
        k0 -= a.x; k1 -= a.y; k2 -= b.x; k3 -= b.y;
      }
    }
  }
  __syncwarp();
}

// EPI selects the output kind at compile time so the inner loops carry no runtime branching on it:
//   EPI_F32    fp32 rows (+ bias, ReLU, up to two residuals; ragged N tail supported)
//   EPI_PLANES fp16 planes for a following GEMM (+ bias, ReLU)
//   EPI_ATT    attention operands: scaled q planes / k planes / per-head transposed v planes (+ fp32 v for FSMN)
//   EPI_F32R2  EPI_F32 with TWO residuals (only reachable through the C ABI; its interior path keeps the one-chunk-ahead pipeline of
//              both residual streams in a kernel instantiation of its own, so the common one-residual kernel can spend those
//              registers on a deeper ring)
constexpr int EPI_F32 = 0, EPI_PLANES = 1, EPI_ATT = 2, EPI_F32R2 = 3;

// fp32-output interior tiles: the residual rows do not depend on the accumulator, so their loads are software pipelined one
// chunk ahead and the first chunk's are issued BEFORE waiting for the accumulator (out-projection / FFN-w_2 epilogues were
// bound by memory-level parallelism: 8 warps x 8 float4 loads in flight per SM sustain ~4 TB/s, measured 262 MB in 66 us).
__device__ __forceinline__ void epilogue_fast_f32(const TcParams& p, const int BN, uint32_t tmem_acc, int64_t row0, int tile_col0, float* stage, int lane,
                                                  int half, uint64_t* full_bar, uint32_t full_phase) {
  const int rr0 = lane >> 2, c4 = (lane & 3) * 4;
  const int64_t rfirst = row0 + rr0;
  float* srow = stage + lane * EPI_LD;
  const float* sp = stage + rr0 * EPI_LD + c4;
  float* c_row = p.C + rfirst * p.ldc + c4 + tile_col0;
  const float* r1_row = p.r1 ? p.r1 + rfirst * p.ldr1 + c4 + tile_col0 : nullptr;
  const float* r2_row = p.r2 ? p.r2 + rfirst * p.ldr2 + c4 + tile_col0 : nullptr;
  const int64_t sc = 8 * p.ldc, s1 = 8 * p.ldr1, s2 = 8 * p.ldr2;
  const bool c_vec = (p.ldc & 3) == 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 nv1[4], nv2[4];
  auto prefetch = [&](int c0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      nv1[it] = r1_row ? __ldg(reinterpret_cast<const float4*>(r1_row + c0 + it * s1)) : z4;
      nv2[it] = r2_row ? __ldg(reinterpret_cast<const float4*>(r2_row + c0 + it * s2)) : z4;
    }
  };
  prefetch(half * EPI_CH);
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
#pragma unroll 2
  for (int c0 = half * EPI_CH; c0 < BN; c0 += 2 * EPI_CH) {
    float4 rv1[4], rv2[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) { rv1[it] = nv1[it]; rv2[it] = nv2[it]; }
    if (c0 + 2 * EPI_CH < BN) prefetch(c0 + 2 * EPI_CH);
    {
      uint32_t r[16];
      tmem_ld_32x16(tmem_acc + c0, r);
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
    __syncwarp();
    float4 bias4 = z4;
    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + tile_col0 + c0 + c4));
    float* pc = c_row + c0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
      float v0 = fmaf(acc.x, p.acc_scale, bias4.x), v1 = fmaf(acc.y, p.acc_scale, bias4.y), v2 = fmaf(acc.z, p.acc_scale, bias4.z), v3 = fmaf(acc.w, p.acc_scale, bias4.w);
      if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
      v0 += rv1[it].x; v1 += rv1[it].y; v2 += rv1[it].z; v3 += rv1[it].w;
      v0 += rv2[it].x; v1 += rv2[it].y; v2 += rv2[it].z; v3 += rv2[it].w;
      if (c_vec) *reinterpret_cast<float4*>(pc) = make_float4(v0, v1, v2, v3);
      else { pc[0] = v0; pc[1] = v1; pc[2] = v2; pc[3] = v3; }
      pc += sc;
    }
    __syncwarp();
  }
}

// The same for at most ONE residual (every fp32-output GEMM of the model: out-projection + x, FFN w_2 + x): the registers the second
// residual's pipeline would hold become a 5-deep ring of the first's, so each warp keeps five 16-column chunks (10 KB) of residual
// rows in flight instead of one.  Measured: out-projection (K = 512, alone, L2 flushed) 63.9 us with one chunk in flight, 60.6 with
// three, 57.3 with five, 46.1 WITHOUT its residual stream — so only part of the gap to the 26 us MMA floor was memory-level
// parallelism; the rest is the L2 -> SM bandwidth the operand tiles (512 KB per tile and CTA for K = 512) and the fp32 rows
// (256 KB in + out) share: 386 MB in 57 us = 6.8 TB/s (tools/gemm_shapes.py, profiles/r2_gemm_shapes_*.json).
__device__ __forceinline__ void epilogue_fast_f32_r1(const TcParams& p, const int BN, uint32_t tmem_acc, int64_t row0, int tile_col0, float* stage,
                                                     int lane, int half, uint64_t* full_bar, uint32_t full_phase) {
  const int rr0 = lane >> 2, c4 = (lane & 3) * 4;
  const int64_t rfirst = row0 + rr0;
  float* srow = stage + lane * EPI_LD;
  const float* sp = stage + rr0 * EPI_LD + c4;
  float* c_row = p.C + rfirst * p.ldc + c4 + tile_col0;
  const float* r_row = p.r1 ? p.r1 + rfirst * p.ldr1 + c4 + tile_col0 : nullptr;
  const int64_t sc = 8 * p.ldc, s1 = 8 * p.ldr1;
  const bool c_vec = (p.ldc & 3) == 0;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int cb = half * EPI_CH;                       // this warp's chunk i covers columns [cb + 32 i, +16)
  constexpr int RING = 5;
  float4 ring[RING][4];
  auto fetch = [&](float4 (&dst)[4], int c0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) dst[it] = r_row ? __ldg(reinterpret_cast<const float4*>(r_row + c0 + it * s1)) : z4;
  };
#pragma unroll
  for (int i = 0; i < RING; ++i)
    if (cb + 32 * i < BN) fetch(ring[i], cb + 32 * i);              // issued BEFORE waiting for the accumulator
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
#pragma unroll
  for (int i = 0; i < 8; ++i) {                                      // BN <= 256: at most 8 chunks per warp
    const int c0 = cb + 32 * i;
    if (c0 < BN) {
      {
        uint32_t r[16];
        tmem_ld_32x16(tmem_acc + c0, r);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      __syncwarp();
      float4 bias4 = z4;
      if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + tile_col0 + c0 + c4));
      float* pc = c_row + c0;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
        const float4 rv = ring[i % RING][it];
        float v0 = fmaf(acc.x, p.acc_scale, bias4.x), v1 = fmaf(acc.y, p.acc_scale, bias4.y), v2 = fmaf(acc.z, p.acc_scale, bias4.z), v3 = fmaf(acc.w, p.acc_scale, bias4.w);
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        v0 += rv.x; v1 += rv.y; v2 += rv.z; v3 += rv.w;
        if (c_vec) *reinterpret_cast<float4*>(pc) = make_float4(v0, v1, v2, v3);
        else { pc[0] = v0; pc[1] = v1; pc[2] = v2; pc[3] = v3; }
        pc += sc;
      }
      __syncwarp();
      if (c0 + 32 * RING < BN) fetch(ring[i % RING], c0 + 32 * RING);
    }
  }
}

// Interior tiles (all 32 rows and all BN columns in range): straight-line code, no bounds predicates, every row base
// computed once per tile and advanced by constant strides.
template <int EPI, int NPL>
__device__ __forceinline__ void epilogue_fast(const TcParams& p, const int BN, uint32_t tmem_acc, int64_t row0, int tile_col0, float* stage, int lane,
                                              int half) {
  const AttnSinks& a = p.att;
  constexpr int QPL = NPL < 2 ? NPL : 2;             // attention operands carry at most two planes
  const int rr0 = lane >> 2, c4 = (lane & 3) * 4;
  const int64_t rfirst = row0 + rr0;
  float* srow = stage + lane * EPI_LD;
  const float* sp = stage + rr0 * EPI_LD + c4;
  float* c_row = (EPI != EPI_PLANES && p.C) ? p.C + rfirst * p.ldc + c4 : nullptr;
  const float* r1_row = (EPI == EPI_F32 && p.r1) ? p.r1 + rfirst * p.ldr1 + c4 : nullptr;
  const float* r2_row = (EPI == EPI_F32 && p.r2) ? p.r2 + rfirst * p.ldr2 + c4 : nullptr;
  plane_t* o_row = EPI == EPI_PLANES ? p.out_planes + rfirst * p.ldo + c4 : nullptr;
  plane_t* q_row = EPI == EPI_ATT ? a.q_planes + rfirst * a.width + c4 : nullptr;
  plane_t* k_row = EPI == EPI_ATT ? a.k_planes + rfirst * a.width + c4 : nullptr;
  const int64_t sc = 8 * p.ldc, s1 = 8 * p.ldr1, s2 = 8 * p.ldr2, so = 8 * p.ldo, sq = 8 * (int64_t)a.width;
  const int64_t plane_o = p.M * p.ldo, plane_q = p.M * (int64_t)a.width;
  const bool c_vec = (p.ldc & 3) == 0;
  plane_t* vt_row = nullptr;
  plane_t* vt_grp = nullptr;                          // staged path: this lane's group of four keys (rows row0 + 4 (lane & 7) ..)
  int64_t vt_plane = 0;
  const bool vt_staged = EPI == EPI_ATT && (a.t_rows & 3) == 0 && (a.t_pad & 3) == 0 && (reinterpret_cast<uintptr_t>(a.vt_planes) & 7) == 0;
  if (EPI == EPI_ATT) {
    const int64_t rw = row0 + lane;
    const int b2 = (int)(rw / a.t_rows), t2 = (int)(rw - (int64_t)b2 * a.t_rows);
    vt_row = a.vt_planes + (int64_t)b2 * a.width * a.t_pad + t2;
    vt_plane = (p.M / a.t_rows) * (int64_t)a.width * a.t_pad;
    const int64_t rg = row0 + 4 * (lane & 7);
    const int bg = (int)(rg / a.t_rows), tg = (int)(rg - (int64_t)bg * a.t_rows);
    vt_grp = a.vt_planes + (int64_t)bg * a.width * a.t_pad + tg;
  }
#pragma unroll 1
  for (int c0 = half * EPI_CH; c0 < BN; c0 += 2 * EPI_CH) {
    const int col0 = tile_col0 + c0;
    const bool v_sink = EPI == EPI_ATT && col0 >= a.v0 && col0 < a.v0 + a.width;
    {
      uint32_t r[16];
      tmem_ld_32x16(tmem_acc + c0, r);
      if (EPI == EPI_ATT && v_sink && vt_staged) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));      // same address in every lane: one broadcast
          x[j] = fmaf(__uint_as_float(r[j]), p.acc_scale, b4.x); x[j + 1] = fmaf(__uint_as_float(r[j + 1]), p.acc_scale, b4.y);
          x[j + 2] = fmaf(__uint_as_float(r[j + 2]), p.acc_scale, b4.z); x[j + 3] = fmaf(__uint_as_float(r[j + 3]), p.acc_scale, b4.w);
        }
        store_vt_staged<QPL>(vt_grp + (int64_t)(col0 - a.v0) * a.t_pad, a.t_pad, vt_plane, x, stage, lane);   // ends with __syncwarp: stage is free again
      }
      if (EPI != EPI_ATT || !v_sink || p.C != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
      if (EPI == EPI_ATT && v_sink && !vt_staged)
        store_vt16<QPL>(vt_row + (int64_t)(col0 - a.v0) * a.t_pad, a.t_pad, vt_plane, r, p.bias ? p.bias + col0 : nullptr, p.acc_scale);
    }
    __syncwarp();
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c4));
    if (EPI == EPI_F32) {
      float* pc = c_row + col0;
      const float* pr1 = r1_row ? r1_row + col0 : nullptr;
      const float* pr2 = r2_row ? r2_row + col0 : nullptr;
      // residual rows of all four passes are fetched up front (memory-level parallelism)
      float4 rv1[4], rv2[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        rv1[it] = pr1 ? __ldg(reinterpret_cast<const float4*>(pr1 + it * s1)) : make_float4(0.f, 0.f, 0.f, 0.f);
        rv2[it] = pr2 ? __ldg(reinterpret_cast<const float4*>(pr2 + it * s2)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
        float v0 = fmaf(acc.x, p.acc_scale, bias4.x), v1 = fmaf(acc.y, p.acc_scale, bias4.y), v2 = fmaf(acc.z, p.acc_scale, bias4.z), v3 = fmaf(acc.w, p.acc_scale, bias4.w);
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        v0 += rv1[it].x; v1 += rv1[it].y; v2 += rv1[it].z; v3 += rv1[it].w;
        v0 += rv2[it].x; v1 += rv2[it].y; v2 += rv2[it].z; v3 += rv2[it].w;
        if (c_vec) *reinterpret_cast<float4*>(pc) = make_float4(v0, v1, v2, v3);
        else { pc[0] = v0; pc[1] = v1; pc[2] = v2; pc[3] = v3; }
        pc += sc;
      }
    } else if (EPI == EPI_PLANES) {
      plane_t* po = o_row + col0;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
        float v0 = fmaf(acc.x, p.acc_scale, bias4.x), v1 = fmaf(acc.y, p.acc_scale, bias4.y), v2 = fmaf(acc.z, p.acc_scale, bias4.z), v3 = fmaf(acc.w, p.acc_scale, bias4.w);
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
        store_planes4<NPL>(po, plane_o, v0, v1, v2, v3);
        po += so;
      }
    } else {
      const bool q_sink = col0 >= a.q0 && col0 < a.q0 + a.width;
      const bool k_sink = col0 >= a.k0 && col0 < a.k0 + a.width;
      if (q_sink || k_sink) {
        plane_t* pq = q_sink ? q_row + (col0 - a.q0) : k_row + (col0 - a.k0);
        const float qs = q_sink ? a.qscale : 1.0f;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
          store_planes4<QPL>(pq, plane_q, __fmul_rn(fmaf(acc.x, p.acc_scale, bias4.x), qs), __fmul_rn(fmaf(acc.y, p.acc_scale, bias4.y), qs),
                             __fmul_rn(fmaf(acc.z, p.acc_scale, bias4.z), qs), __fmul_rn(fmaf(acc.w, p.acc_scale, bias4.w), qs));
          pq += sq;
        }
      } else if (v_sink && c_row) {                    // fp32 V rows feed the FSMN memory block
        float* pc = c_row + col0;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
          const float4 o = make_float4(fmaf(acc.x, p.acc_scale, bias4.x), fmaf(acc.y, p.acc_scale, bias4.y), fmaf(acc.z, p.acc_scale, bias4.z),
                                       fmaf(acc.w, p.acc_scale, bias4.w));
          if (c_vec) *reinterpret_cast<float4*>(pc) = o;
          else { pc[0] = o.x; pc[1] = o.y; pc[2] = o.z; pc[3] = o.w; }
          pc += sc;
        }
      }
    }
    __syncwarp();
  }
}

// Edge tiles (row tail of M, ragged N such as vocab 8404 / 25055): every access bounds checked.
template <int EPI, int NPL>
__device__ __noinline__ void epilogue_edge(const TcParams& p, const int BN, uint32_t tmem_acc, int64_t row0, int tile_col0, float* stage, int lane,
                                           int half) {
  const AttnSinks& a = p.att;
  constexpr int QPL = NPL < 2 ? NPL : 2;
#pragma unroll 1
  for (int c0 = half * EPI_CH; c0 < BN; c0 += 2 * EPI_CH) {
    const int col0 = tile_col0 + c0;
    {
      uint32_t r[16];
      tmem_ld_32x16(tmem_acc + c0, r);
      float* srow = stage + lane * EPI_LD;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      if (EPI == EPI_ATT && col0 >= a.v0 && col0 < a.v0 + a.width && row0 + lane < p.M) {
        const int64_t rw = row0 + lane;
        const int b2 = (int)(rw / a.t_rows), t2 = (int)(rw - (int64_t)b2 * a.t_rows);
        store_vt16<QPL>(a.vt_planes + ((int64_t)b2 * a.width + (col0 - a.v0)) * a.t_pad + t2, a.t_pad,
                        (p.M / a.t_rows) * (int64_t)a.width * a.t_pad, r, p.bias ? p.bias + col0 : nullptr, p.acc_scale);
      }
    }
    __syncwarp();
    if (col0 < p.N && row0 < p.M) {
      const bool full = col0 + EPI_CH <= p.N;
      const bool v_sink = EPI == EPI_ATT && col0 >= a.v0 && col0 < a.v0 + a.width;
      const bool q_sink = EPI == EPI_ATT && col0 >= a.q0 && col0 < a.q0 + a.width;
      const bool k_sink = EPI == EPI_ATT && col0 >= a.k0 && col0 < a.k0 + a.width;
      const bool want_c = EPI == EPI_F32 ? true : (EPI == EPI_ATT ? (p.C != nullptr && v_sink) : false);
      if (EPI != EPI_ATT || want_c || q_sink || k_sink) {
        const int rr0 = lane >> 2, c4 = (lane & 3) * 4;
        const int col = col0 + c4;
        const int64_t rfirst = row0 + rr0;
        const int rows_left = (int)((p.M - rfirst + 7) >> 3);        // passes (of 8 rows) with a valid row for this lane
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) {
          if (full) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
          else { float* bp = reinterpret_cast<float*>(&bias4); for (int e = 0; e < 4; ++e) if (col + e < p.N) bp[e] = __ldg(p.bias + col + e); }
        }
        const float* sp = stage + rr0 * EPI_LD + c4;
        const bool c_vec = (p.ldc & 3) == 0;
        for (int it = 0; it < 4 && it < rows_left; ++it) {
          const int64_t row = rfirst + 8 * it;
          const float4 acc = *reinterpret_cast<const float4*>(sp + it * 8 * EPI_LD);
          float vv[4] = {fmaf(acc.x, p.acc_scale, bias4.x), fmaf(acc.y, p.acc_scale, bias4.y), fmaf(acc.z, p.acc_scale, bias4.z), fmaf(acc.w, p.acc_scale, bias4.w)};
          if (p.relu) { for (int e = 0; e < 4; ++e) vv[e] = fmaxf(vv[e], 0.f); }
          if (full) {
            if (EPI == EPI_F32) {
              for (int e = 0; e < 4; ++e) {
                if (p.r1) vv[e] += __ldg(p.r1 + row * p.ldr1 + col + e);
                if (p.r2) vv[e] += __ldg(p.r2 + row * p.ldr2 + col + e);
              }
            }
            if (want_c) {
              float* pc = p.C + row * p.ldc + col;
              if (c_vec) *reinterpret_cast<float4*>(pc) = make_float4(vv[0], vv[1], vv[2], vv[3]);
              else { pc[0] = vv[0]; pc[1] = vv[1]; pc[2] = vv[2]; pc[3] = vv[3]; }
            }
            if (EPI == EPI_PLANES) store_planes4<NPL>(p.out_planes + row * p.ldo + col, p.M * p.ldo, vv[0], vv[1], vv[2], vv[3]);
            if (q_sink) store_planes4<QPL>(a.q_planes + row * a.width + (col - a.q0), p.M * (int64_t)a.width, __fmul_rn(vv[0], a.qscale),
                                           __fmul_rn(vv[1], a.qscale), __fmul_rn(vv[2], a.qscale), __fmul_rn(vv[3], a.qscale));
            if (k_sink) store_planes4<QPL>(a.k_planes + row * a.width + (col - a.k0), p.M * (int64_t)a.width, vv[0], vv[1], vv[2], vv[3]);
          } else {                                                   // ragged N tail: scalar (fp32 output only)
            for (int e = 0; e < 4; ++e) {
              if (col + e >= p.N) break;
              float x = vv[e];
              if (p.r1) x += __ldg(p.r1 + row * p.ldr1 + col + e);
              if (p.r2) x += __ldg(p.r2 + row * p.ldr2 + col + e);
              if (want_c) p.C[row * p.ldc + col + e] = x;
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

// BN: columns of this accumulator tile (a compile-time constant in the single-CTA kernel; 256 / 128 / 64 in the pair kernel)
template <int EPI, int NPL>
__device__ __forceinline__ void epilogue_warp(const TcParams& p, const int BN, uint32_t tmem_acc, int64_t row0, int tile_col0, float* stage, int lane,
                                              int half, uint64_t* full_bar, uint32_t full_phase) {
  constexpr int EPIB = EPI == EPI_F32R2 ? EPI_F32 : EPI;
  const bool interior = row0 + 32 <= p.M && tile_col0 + BN <= p.N;
  if (EPI == EPI_F32 && interior) {                                                                  // both wait for the accumulator themselves
    epilogue_fast_f32_r1(p, BN, tmem_acc, row0, tile_col0, stage, lane, half, full_bar, full_phase);
    return;
  }
  if (EPI == EPI_F32R2 && interior) {
    epilogue_fast_f32(p, BN, tmem_acc, row0, tile_col0, stage, lane, half, full_bar, full_phase);
    return;
  }
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
  if (interior) epilogue_fast<EPIB, NPL>(p, BN, tmem_acc, row0, tile_col0, stage, lane, half);
  else epilogue_edge<EPIB, NPL>(p, BN, tmem_acc, row0, tile_col0, stage, lane, half);
}

template <int BN, int STAGES, int APL, int WPL, int EPI>  // APL / WPL: A / W planes resident per stage
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr uint32_t TILE_W_BYTES = BN * TC_BK * 2;
  constexpr uint32_t STAGE_BYTES = APL * TC_TILE_BYTES_A + WPL * TILE_W_BYTES;
  // align to 1024 B WITHOUT leaving the shared address space (a uintptr_t round trip makes every access a generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // 8 x [32][20] floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.tiles_m * p.tiles_n;
  const int k_blocks = p.Kp / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();                                          // prologue above is global-memory free; operands are read below
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one_sync()) {          // single elected lane: no waterfall loops around the uniform-datapath TMA / MMA instructions
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          unsigned char* st = smem + stage * STAGE_BYTES;
#pragma unroll
          for (int pl = 0; pl < APL; ++pl)
            tma_load_2d(st + pl * TC_TILE_BYTES_A, &map_a, &full_bar[stage], kb * TC_BK, (int)(pl * p.a_plane_rows + (int64_t)tm * TC_BM));
#pragma unroll
          for (int pl = 0; pl < WPL; ++pl)
#pragma unroll
            for (int hb = 0; hb < BN / 128; ++hb)   // W box = 128 rows; a 256-wide tile is two boxes, contiguous in smem
              tma_load_2d(st + APL * TC_TILE_BYTES_A + pl * TILE_W_BYTES + hb * (128 * TC_BK * 2), &map_w, &full_bar[stage], kb * TC_BK,
                          pl * p.w_plane_rows + tn * BN + hb * 128);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {          // single elected lane: no waterfall loops around the uniform-datapath TMA / MMA instructions
      constexpr uint32_t idesc = make_idesc_f16(TC_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          for (int t = 0; t < p.n_terms; ++t) {
            const uint64_t da = make_sw128_desc(st + c_term_a[t] * TC_TILE_BYTES_A);
            const uint64_t dw = make_sw128_desc(st + APL * TC_TILE_BYTES_A + c_term_w[t] * TILE_W_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UK; ++k) {
              // advance 32 bytes (16 fp16) inside the 128-byte swizzle span: +2 in 16-byte units
              umma_f16(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | t | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);                  // ring slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                      // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (warps 4..11: TMEM lane quarter warp%4, alternate 16-column chunks) =====================
    const int q = warp & 3, half = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tm = tile / p.tiles_n, tn = tile - tm * p.tiles_n;
      epilogue_warp<EPI, APL>(p, BN, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, (int64_t)tm * TC_BM + q * 32, tn * BN,
                                  epi_stage + (warp - 4) * EPI_WARP_FLOATS, lane, half, &tmem_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);        // 4 arrivals (one per epilogue warp) free the accumulator
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}


// ------------------------------------------------------------------------------------------------ cta_group::2 variant
// CTA pair (cluster 2x1x1) computes a 256 x 256 output tile: each CTA stages ITS 128 rows of A and ITS 128 rows of W
// (half of the N tile) per k-block, so operand bytes per MMA cycle are half those of the single-CTA 128x256 tile
// (64 KB per k-block per SM for the three x3 terms) and three stages fit.  The leader issues
// tcgen05.mma.cta_group::2 (M=256); each CTA drains its own 128 TMEM lanes in the epilogue.
//
// Schedule: pair q of P runs the whole tiles q, q + P, ... of the first full_rounds * P tiles, then the column slices q, q + P, ...
// of the leftover tiles (item_of).  A 256 x 256 tile takes ~9 us (K = 512, x3) to ~35 us (K = 2048); with 250 tiles on 74 pairs
// the last round kept 28 pairs busy for a whole tile time while 46 idled (16 % of the FFN-w_2 / out-projection launch).  Slices
// of 128 / 64 columns (tcgen05.mma N = 128 / 64, the W box of each CTA 64 / 32 rows through a second tensor map) spread those 28
// tiles over 56 / 112 items.  Narrow slices re-read the A rows from L2 more often, so only the tail uses them.
struct TcItem { int tm, col0, bn; };
__device__ __forceinline__ bool item_of(const TcParams& p, int pair, int n_pairs, int it, TcItem& o) {
  int tile, part = 0;
  o.bn = 256;
  if (it < p.full_rounds) {
    tile = pair + it * n_pairs;
  } else {
    const int j = pair + (it - p.full_rounds) * n_pairs;
    if (j >= p.tail_tiles * p.tail_sub) return false;
    tile = p.full_rounds * n_pairs + j / p.tail_sub;
    part = j - (j / p.tail_sub) * p.tail_sub;
    o.bn = 256 / p.tail_sub;
  }
  o.tm = tile / p.tiles_n;
  o.col0 = (tile - o.tm * p.tiles_n) * 256 + part * o.bn;
  return true;
}

template <int STAGES, int PL, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_tail,
                const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int BN = 256;
  constexpr uint32_t TILE_BYTES = 128 * TC_BK * 2;                 // 16 KB: 128 rows x 64 fp16
  constexpr uint32_t STAGE_BYTES = 2 * PL * TILE_BYTES;            // A planes + W-half planes of this CTA
  // align to 1024 B WITHOUT leaving the shared address space (a uintptr_t round trip makes every access a generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2] (used in the leader only: 8 arrivals = 4 epilogue warps x 2 CTAs)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // 8 x [32][20] floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int k_blocks = p.Kp / TC_BK;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_w_tail); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_base_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // barriers of both CTAs are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  pdl_wait();                                          // prologue above is global-memory free; operands are read below
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one_sync()) {          // single elected lane: no waterfall loops around the uniform-datapath TMA / MMA instructions
      int stage = 0; uint32_t phase = 0;
      TcItem it;
      for (int i = 0; item_of(p, pair, n_pairs, i, it); ++i) {
        const CUtensorMap* mw = it.bn == BN ? &map_w : &map_w_tail;            // W box: bn / 2 rows per CTA
        const uint32_t w_bytes = (uint32_t)(it.bn / 2) * (TC_BK * 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * PL * (TILE_BYTES + w_bytes));   // bytes of BOTH CTAs land on the leader's barrier
          unsigned char* st = smem + stage * STAGE_BYTES;
#pragma unroll
          for (int pl = 0; pl < PL; ++pl)
            tma_load_2d_2sm(st + pl * TILE_BYTES, &map_a, &full_bar[stage], kb * TC_BK,
                            (int)(pl * p.a_plane_rows + (int64_t)it.tm * 256 + rank * 128));
#pragma unroll
          for (int pl = 0; pl < PL; ++pl)
            tma_load_2d_2sm(st + (PL + pl) * TILE_BYTES, mw, &full_bar[stage], kb * TC_BK,
                            pl * p.w_plane_rows + it.col0 + (int)rank * (it.bn / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one_sync()) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      TcItem it;
      for (int i = 0; item_of(p, pair, n_pairs, i, it); ++i) {
        const uint32_t idesc = make_idesc_f16(256, it.bn);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);       // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          for (int t = 0; t < p.n_terms; ++t) {
            const uint64_t da = make_sw128_desc(st + c_term_a[t] * TILE_BYTES);
            const uint64_t dw = make_sw128_desc(st + (PL + c_term_w[t]) * TILE_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UK; ++k) umma_f16_2sm(d_tmem, da + 2 * k, dw + 2 * k, idesc, (kb | t | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);              // frees the slot in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tmem_full[acc]);                  // accumulators complete in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (each CTA: its 128 rows) =====================
    const int q = warp & 3, half = (warp - 4) >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    TcItem it;
    for (int i = 0; item_of(p, pair, n_pairs, i, it); ++i) {
      epilogue_warp<EPI, PL>(p, it.bn, tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN, (int64_t)it.tm * 256 + rank * 128 + q * 32, it.col0,
                             epi_stage + (warp - 4) * EPI_WARP_FLOATS, lane, half, &tmem_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer may still arrive on / read this CTA's shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

// fp32 rows [rows, cols] (ld) -> fp16 planes [nplanes][rows][cols_pad]; 4 elements per thread.
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols, int cols_pad, int nplanes,
                  plane_t* __restrict__ planes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = cols_pad >> 2;
  const int64_t total = rows * c4n;
  if (i >= total) return;
  const int64_t r = i / c4n;
  const int c = (int)(i - r * c4n) * 4;
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c + 3 < cols) x = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
  else {
    float* e = reinterpret_cast<float*>(&x);
    for (int k = 0; k < 4; ++k) if (c + k < cols) e[k] = __ldg(src + r * ld + c + k);
  }
  float v[4] = {x.x, x.y, x.z, x.w};
  const int64_t plane = rows * cols_pad;
  for (int pl = 0; pl < nplanes; ++pl) {
    uint2 pk;
    pk.x = pack_planes2(v[0], v[1]);
    pk.y = pack_planes2(v[2], v[3]);
    *reinterpret_cast<uint2*>(planes + pl * plane + r * cols_pad + c) = pk;
    const float2 a = unpack_planes2(pk.x), b = unpack_planes2(pk.y);
    v[0] -= a.x; v[1] -= a.y; v[2] -= b.x; v[3] -= b.y;
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static const PFN_encodeTiled fn = []() -> PFN_encodeTiled {       // thread-safe one-time lookup
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<PFN_encodeTiled>(p);
    return nullptr;
  }();
  return fn;
}

// 2D fp16 tensor [rows, cols] (row pitch ld elements), box {64 cols, box_rows}, 128B swizzle.
// A forward pass encodes ~570 maps over a few dozen distinct (pointer, shape) pairs — the workspace slices and the weight
// planes are the same every layer and every step — so the encoded descriptors are kept in a small per-thread cache
// (no locking; a descriptor depends only on the key).
struct MapKey {
  const void* base; uint64_t rows, cols, ld; uint32_t box;
  bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = (uint64_t)(uintptr_t)k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 31 + k.ld * 131 + k.box + (h << 6) + (h >> 2));
    return (size_t)h;
  }
};
int make_plane_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{base, rows, cols, ld, box_rows};
  auto it = cache.find(key);
  if (it != cache.end()) { *m = it->second; return FA_OK; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) return FA_ERR_CUDA;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return FA_ERR_CUDA;
  if (cache.size() >= 4096) cache.clear();
  cache.emplace(key, *m);
  return FA_OK;
}

static int planes_for_mode(int mode) { return mode == FA_GEMM_F16X1 ? 1 : (mode == FA_GEMM_F16X3 ? 2 : 3); }

size_t gemm_tc_scratch_bytes(int64_t max_rows, int max_k, int mode) {
  if (mode == FA_GEMM_F32_SIMT) return 0;
  const int kp = (max_k + 63) / 64 * 64;
  return (size_t)planes_for_mode(mode) * (size_t)max_rows * kp * 2 + 1024;
}

template <int BN, int STAGES, int APL, int WPL, int EPI>
static int launch_cfg_e(const CUtensorMap& ma, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (APL * TC_TILE_BYTES_A + WPL * BN * TC_BK * 2) + 1024 + 256 + EPI_WARPS * EPI_WARP_FLOATS * 4;
  static PerDeviceOnce once;
  FA_RETURN_IF_ERR(ensure_dyn_smem(gemm_tc_kernel<BN, STAGES, APL, WPL, EPI>, smem, once));
  const int n_sm = sm_count();
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < n_sm ? tiles : n_sm;
  FA_CUDA_OK(launch_pdl(gemm_tc_kernel<BN, STAGES, APL, WPL, EPI>, dim3(grid), dim3(384), smem, st, 1, ma, mw, p));
  FA_CHECK_LAUNCH();
  return FA_OK;
}

static inline int epi_kind(const TcParams& p) { return p.att.enabled ? EPI_ATT : (p.out_planes ? EPI_PLANES : (p.r2 ? EPI_F32R2 : EPI_F32)); }

template <int BN, int STAGES, int APL, int WPL>
static int launch_cfg(const CUtensorMap& ma, const CUtensorMap& mw, const TcParams& p, cudaStream_t st) {
  switch (epi_kind(p)) {
    case EPI_ATT: return launch_cfg_e<BN, STAGES, APL, WPL, EPI_ATT>(ma, mw, p, st);
    case EPI_PLANES: return launch_cfg_e<BN, STAGES, APL, WPL, EPI_PLANES>(ma, mw, p, st);
    case EPI_F32R2: return launch_cfg_e<BN, STAGES, APL, WPL, EPI_F32R2>(ma, mw, p, st);
    default: return launch_cfg_e<BN, STAGES, APL, WPL, EPI_F32>(ma, mw, p, st);
  }
}

template <int STAGES, int PL, int EPI>
static int launch_cfg2_e(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mwt, const TcParams& p, int pairs, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * PL * 128 * TC_BK * 2) + 1024 + 256 + EPI_WARPS * EPI_WARP_FLOATS * 4;
  static PerDeviceOnce once;
  FA_RETURN_IF_ERR(ensure_dyn_smem(gemm_tc2_kernel<STAGES, PL, EPI>, smem, once));
  FA_CUDA_OK(launch_pdl(gemm_tc2_kernel<STAGES, PL, EPI>, dim3(2 * pairs), dim3(384), smem, st, 1, ma, mw, mwt, p));   // cluster dims are compile-time (__cluster_dims__)
  FA_CHECK_LAUNCH();
  return FA_OK;
}

template <int STAGES, int PL>
static int launch_cfg2(const CUtensorMap& ma, const CUtensorMap& mw, const CUtensorMap& mwt, const TcParams& p, int pairs, cudaStream_t st) {
  switch (epi_kind(p)) {
    case EPI_ATT: return launch_cfg2_e<STAGES, PL, EPI_ATT>(ma, mw, mwt, p, pairs, st);
    case EPI_PLANES: return launch_cfg2_e<STAGES, PL, EPI_PLANES>(ma, mw, mwt, p, pairs, st);
    case EPI_F32R2: return launch_cfg2_e<STAGES, PL, EPI_F32R2>(ma, mw, mwt, p, pairs, st);
    default: return launch_cfg2_e<STAGES, PL, EPI_F32>(ma, mw, mwt, p, pairs, st);
  }
}

// Column slices per leftover tile (1 | 2 | 4): the one that minimises the tail's duration ceil(tail * sub / P) / sub, the coarser
// one on ties (wider tiles move fewer operand bytes per flop).  FA_GEMM_TAIL=0 keeps whole tiles (A/B runs).
static int pick_tail_sub(int tail_tiles, int pairs) {
  static const bool on = [] { const char* e = getenv("FA_GEMM_TAIL"); return !(e && e[0] == '0'); }();
  if (!on || tail_tiles == 0) return 1;
  int best = 1, best_q = 4 * ((tail_tiles + pairs - 1) / pairs);           // duration in quarter rounds
  for (int sub = 2; sub <= 4; sub *= 2) {
    const int q = ((tail_tiles * sub + pairs - 1) / pairs) * (4 / sub);
    if (q < best_q) { best_q = q; best = sub; }
  }
  return best;
}

static bool use_2cta() {
  static const bool on = [] { const char* e = getenv("FA_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  return on;
}

// A planes already split: a_planes [npl][M][Kp]
int gemm_tc_planes_launch(const plane_t* a_planes, int64_t M, const FaLinear& lin, int relu, const float* r1, int64_t ld1,
                          const float* r2, int64_t ld2, float* y, int64_t ldy, plane_t* out_planes, int64_t ldo,
                          int mode, cudaStream_t st, const AttnSinks* att, int64_t a_ld, int64_t a_plane_rows) {
  if (M <= 0) return FA_OK;
  const uint64_t lda = a_ld > 0 ? (uint64_t)a_ld : (uint64_t)lin.in_pad;          // A row pitch (overlapping view: < K_pad)
  const int64_t apr = a_plane_rows > 0 ? a_plane_rows : M;                         // rows between consecutive A planes
  if (lda & 7) return FA_ERR_UNSUPPORTED;
  if (!lin.w_planes || !a_planes) return FA_ERR_ARG;
  const int N = lin.out_f, Kp = lin.in_pad;
  if (Kp % TC_BK != 0 || M * 3 > 0x7fffffffLL) return FA_ERR_UNSUPPORTED;
  if (y && (ldy & 3) == 0 && (((uintptr_t)y) & 15)) return FA_ERR_UNSUPPORTED;
  if ((r1 && (ld1 & 3)) || (r2 && (ld2 & 3))) return FA_ERR_UNSUPPORTED;
  const int npl = planes_for_mode(mode);
  // 128x256 tiles halve the operand bytes per MMA cycle (the 128x128 tile is L2-bandwidth bound); used when N splits
  // evenly and there are enough tiles to fill the machine.  x6 keeps 128x128 (three planes per operand do not fit twice).
  // Ragged N (the vocabulary projections: 8404, 25055) also runs on pair tiles when the output is plain fp32 rows: the last column
  // tile's W box reaches past row N of a plane — into the next plane's first rows or, for the last plane, out of the tensor map
  // (zero fill) — so its surplus accumulator columns hold finite garbage that the bounds-checked edge epilogue never stores.
  const bool ragged_ok = N >= 1024 && !out_planes && !att;
  if (use_2cta() && npl <= 2 && (N % 256 == 0 || ragged_ok) && M >= 256) {
    // cta_group::2: 256 x 256 pair tiles (see gemm_tc2_kernel)
    CUtensorMap ma2, mw2;
    // rows of the map: every row of every plane whose K_pad elements lie inside the allocation (an overlapping view's last rows do not)
    const uint64_t a_rows = (uint64_t)apr * npl - (lda < (uint64_t)Kp ? ((uint64_t)Kp - lda + lda - 1) / lda : 0);
    FA_RETURN_IF_ERR(make_plane_map(&ma2, a_planes, a_rows, (uint64_t)Kp, lda, 128));
    FA_RETURN_IF_ERR(make_plane_map(&mw2, lin.w_planes, (uint64_t)N * 3, (uint64_t)Kp, (uint64_t)Kp, 128));
    TcParams p2;
    p2.M = M; p2.N = N; p2.Kp = Kp; p2.a_plane_rows = apr; p2.w_plane_rows = N;
    p2.n_terms = mode == FA_GEMM_F16X1 ? 1 : 3;
    p2.relu = relu; p2.bias = lin.b; p2.r1 = r1; p2.ldr1 = ld1; p2.r2 = r2; p2.ldr2 = ld2; p2.C = y; p2.ldc = ldy;
    p2.out_planes = out_planes; p2.ldo = ldo; p2.out_nplanes = npl;
    p2.tiles_m = (int)((M + 255) / 256); p2.tiles_n = (N + 255) / 256;
    p2.acc_scale = rz_comp_scale(Kp, p2.n_terms);
    if (att) { p2.att = *att; p2.att.enabled = 1; } else { p2.att = AttnSinks{}; }
    if (att && (att->width % 32 != 0 || att->t_rows <= 0 || M % att->t_rows != 0)) return FA_ERR_UNSUPPORTED;
    const int tiles = p2.tiles_m * p2.tiles_n, half_sms = sm_count() / 2;
    p2.full_rounds = tiles / half_sms;
    p2.tail_tiles = tiles - p2.full_rounds * half_sms;
    p2.tail_sub = pick_tail_sub(p2.tail_tiles, half_sms);
    const int pairs = p2.full_rounds > 0 ? half_sms : (p2.tail_tiles * p2.tail_sub < half_sms ? p2.tail_tiles * p2.tail_sub : half_sms);
    CUtensorMap mwt = mw2;
    if (p2.tail_sub > 1) FA_RETURN_IF_ERR(make_plane_map(&mwt, lin.w_planes, (uint64_t)N * 3, (uint64_t)Kp, (uint64_t)Kp, 128 / p2.tail_sub));
    return npl == 1 ? launch_cfg2<6, 1>(ma2, mw2, mwt, p2, pairs, st) : launch_cfg2<3, 2>(ma2, mw2, mwt, p2, pairs, st);
  }
  const bool wide = (npl <= 2) && (N % 256 == 0) && (N >= 1024);
  const int BN = wide ? 256 : 128;
  CUtensorMap ma, mw;
  const uint64_t a_rows1 = (uint64_t)apr * npl - (lda < (uint64_t)Kp ? ((uint64_t)Kp - lda + lda - 1) / lda : 0);
  FA_RETURN_IF_ERR(make_plane_map(&ma, a_planes, a_rows1, (uint64_t)Kp, lda, TC_BM));
  FA_RETURN_IF_ERR(make_plane_map(&mw, lin.w_planes, (uint64_t)N * 3, (uint64_t)Kp, (uint64_t)Kp, wide ? 128 : BN));
  TcParams p;
  p.M = M; p.N = N; p.Kp = Kp; p.a_plane_rows = apr; p.w_plane_rows = N;
  p.n_terms = mode == FA_GEMM_F16X1 ? 1 : (mode == FA_GEMM_F16X3 ? 3 : 6);
  p.relu = relu; p.bias = lin.b; p.r1 = r1; p.ldr1 = ld1; p.r2 = r2; p.ldr2 = ld2; p.C = y; p.ldc = ldy;
  p.out_planes = out_planes; p.ldo = ldo; p.out_nplanes = npl;
  p.tiles_m = (int)((M + TC_BM - 1) / TC_BM); p.tiles_n = (N + BN - 1) / BN;
  p.full_rounds = 0; p.tail_tiles = 0; p.tail_sub = 1;
  p.acc_scale = rz_comp_scale(Kp, p.n_terms);
  if (att) { p.att = *att; p.att.enabled = 1; } else { p.att = AttnSinks{}; }
  if (att && (N % 32 != 0 || att->width % 32 != 0 || att->t_rows <= 0 || M % att->t_rows != 0)) return FA_ERR_UNSUPPORTED;
  if (out_planes && (N % 32 != 0)) return FA_ERR_UNSUPPORTED;
  if (wide) return npl == 1 ? launch_cfg<256, 4, 1, 1>(ma, mw, p, st) : launch_cfg<256, 2, 2, 2>(ma, mw, p, st);
  switch (npl) {
    case 1: return launch_cfg<128, 6, 1, 1>(ma, mw, p, st);
    case 2: return launch_cfg<128, 3, 2, 2>(ma, mw, p, st);
    default: return launch_cfg<128, 2, 3, 3>(ma, mw, p, st);
  }
}

int split_rows_launch(const float* x, int64_t ldx, int64_t rows, int cols, int cols_pad, int nplanes, plane_t* planes,
                      cudaStream_t st) {
  if (rows <= 0) return FA_OK;
  if ((ldx & 3) || (((uintptr_t)x) & 15) || (cols_pad & 3)) return FA_ERR_UNSUPPORTED;
  const int64_t total = rows * (cols_pad / 4);
  split_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, ldx, rows, cols, cols_pad, nplanes, planes);
  FA_CHECK_LAUNCH();
  return FA_OK;
}

int gemm_tc_launch(const float* x, int64_t ldx, int64_t rows, const FaLinear& lin, int relu, const float* r1, int64_t ld1,
                   const float* r2, int64_t ld2, float* y, int64_t ldy, int mode, Arena* scratch, cudaStream_t st) {
  if (rows <= 0) return FA_OK;
  if (mode != FA_GEMM_F16X1 && mode != FA_GEMM_F16X3 && mode != FA_GEMM_F16X6) return FA_ERR_ARG;
  if (!scratch) return FA_ERR_WORKSPACE;
  const int npl = planes_for_mode(mode);
  Arena local(scratch->base, scratch->cap);   // scratch is reused by every call (stream ordered)
  plane_t* planes = local.take<plane_t>((size_t)npl * rows * lin.in_pad);
  if (!local.ok()) return FA_ERR_WORKSPACE;
  FA_RETURN_IF_ERR(split_rows_launch(x, ldx, rows, lin.in_f, lin.in_pad, npl, planes, st));
  return gemm_tc_planes_launch(planes, rows, lin, relu, r1, ld1, r2, ld2, y, ldy, nullptr, 0, mode, st, nullptr);
}

}  // namespace fa
