// tcgen05 / TMEM / TMA fused multi-head attention with bf16 operand splitting (sm_100a).
//
// Same contract as attention_f32.cu (sanm/attention.py:288-304 with scores from :324-325; cross-attention :760-794,
// :811-812): ctx = softmax(mask(q d_k^-0.5 . k^T)) v per head, key-padding mask, heads merged.  Score and context
// contractions run on the 5th-gen tensor cores with fp32 accumulation in TMEM; operands are bf16 planes (hi, lo) of the
// fp32 tensors so that S = Qh.Kh + Qh.Kl + Ql.Kh and O = Ph.Vh + Ph.Vl + Pl.Vh carry ~2^-17 relative error (x3 mode),
// or one plane (x1 mode).  The [B,H,Tq,Tk] score tensor never leaves the SM.
//
// One CTA = 128 queries of one (utterance, head), keys in chunks of 64, two passes over the keys:
//   pass A: S~ = Qh.Kh (one MMA term) -> per-row max m   (softmax is invariant to the choice of m; an approximate
//           maximum only has to keep exp(s - m) in range, so one bf16 term suffices)
//   pass B: S (all terms) -> p = exp(s - m), l += sum p, P planes -> smem, O += P.V accumulated in TMEM with no
//           rescaling traffic; finally O / l -> bf16 planes (A operand of the out-projection GEMM) and/or fp32.
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 softmax + epilogue:
// TMEM lane == query row; two threads per row (warp w owns lane quarter w%4 and column half (w-4)/4) so that every SM
// sub-partition has two softmax warps to interleave.  S is double buffered in TMEM so the MMAs of chunk j+1 overlap
// the softmax of chunk j.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include <math.h>

namespace fa {

constexpr int AT_BQ = 128, AT_BKEY = 64, AT_D = 128;
constexpr uint32_t AT_Q_KBLK = AT_BQ * 128;      // 16 KB: 128 rows x 64 bf16
constexpr uint32_t AT_K_KBLK = AT_BKEY * 128;    // 8 KB : 64 keys x 64 bf16
constexpr uint32_t AT_V_TILE = AT_D * 128;       // 16 KB: 128 d-rows x 64 keys
constexpr uint32_t AT_P_TILE = AT_BQ * 128;      // 16 KB: 128 queries x 64 keys

struct AttTcParams {
  int tq, tk, heads, batch;
  const int32_t* key_lens;
  int64_t q_plane_rows, k_plane_rows, v_plane_rows;   // rows between planes in the respective 2D maps
  float* ctx; int64_t ldc;                             // fp32 output (or null)
  __nv_bfloat16* ctx_planes; int64_t ldp; int out_nplanes;   // bf16 planes [npl][B*tq][ldp] (or null)
};

__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 softmax warps

template <int NPL, int OPL>   // NPL: operand planes (1 | 2); OPL: context planes written (0 = fp32 context only)
__global__ void __launch_bounds__(384, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const AttTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr uint32_t Q_BYTES = NPL * 2 * AT_Q_KBLK;
  constexpr uint32_t K_BYTES = NPL * 2 * AT_K_KBLK;
  constexpr uint32_t V_BYTES = NPL * AT_V_TILE;
  constexpr uint32_t P_BYTES = NPL * AT_P_TILE;
  constexpr int NT = NPL == 1 ? 1 : 3;
  // align to 1024 B WITHOUT leaving the shared address space (a uintptr_t round trip makes every access a generic LD/ST)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sQ = smem;
  // K and V chunks share ONE 3-slot ring, filled in exactly the order the MMA warp consumes them
  //   pass A: Khi(0) .. Khi(nc-1)          pass B: K(0), K(1), V(0), K(2), V(1), ..., K(nc-1), V(nc-2), V(nc-1)
  // (a K chunk is dead once its score MMAs retire, its V twin only at P.V time, and S(t+1) is issued before P.V(t)), which
  // frees 32 KB against separate 2+2 rings — spent on double buffering P, so softmax(t+1) writes its probabilities while
  // P.V(t) is still running instead of waiting for it (v2 was bound by that hand-off: tensor pipe 34 %).
  constexpr uint32_t SLOT_BYTES = K_BYTES > V_BYTES ? K_BYTES : V_BYTES;
  unsigned char* sRing = sQ + Q_BYTES;
  unsigned char* sP = sRing + 3 * SLOT_BYTES;        // [2][P_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* r_full = bars + 1;        // [3]
  uint64_t* r_empty = bars + 4;       // [3]
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* s_empty = bars + 9;       // [2]
  uint64_t* p_full = bars + 11;       // [2]
  uint64_t* p_empty = bars + 13;      // [2]
  uint64_t* o_full = bars + 15;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* s_red = reinterpret_cast<float*>(bars + 17);   // [2][128] row max / row sum exchange between column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int klen = min(p.key_lens[b], p.tk);
  const int nc = (klen + AT_BKEY - 1) / AT_BKEY;     // key chunks with at least one valid key

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 3; ++s) { mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 8);
      mbar_init(&p_full[s], 8); mbar_init(&p_empty[s], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && nc > 0) {
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int pl = 0; pl < NPL; ++pl)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(sQ + (pl * 2 + kb) * AT_Q_KBLK, &map_q, q_full, h * AT_D + kb * 64,
                      (int)(pl * p.q_plane_rows + (int64_t)b * p.tq + q0));
      uint32_t n = 0;                                       // ring sequence number
      auto load_k = [&](int j, int npl_load) {
        const uint32_t slot = n % 3u;
        mbar_wait(&r_empty[slot], ((n / 3u) & 1u) ^ 1u);
        mbar_expect_tx(&r_full[slot], (uint32_t)npl_load * 2u * AT_K_KBLK);
        unsigned char* dst = sRing + slot * SLOT_BYTES;
        for (int pl = 0; pl < npl_load; ++pl)
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d(dst + (pl * 2 + kb) * AT_K_KBLK, &map_k, &r_full[slot], h * AT_D + kb * 64,
                        (int)(pl * p.k_plane_rows + (int64_t)b * p.tk + j * AT_BKEY));
        ++n;
      };
      auto load_v = [&](int t) {
        const uint32_t slot = n % 3u;
        mbar_wait(&r_empty[slot], ((n / 3u) & 1u) ^ 1u);
        mbar_expect_tx(&r_full[slot], V_BYTES);
        unsigned char* dst = sRing + slot * SLOT_BYTES;
#pragma unroll
        for (int pl = 0; pl < NPL; ++pl)
          tma_load_2d(dst + pl * AT_V_TILE, &map_v, &r_full[slot], t * AT_BKEY,
                      (int)(pl * p.v_plane_rows + ((int64_t)b * p.heads + h) * AT_D));
        ++n;
      };
      for (int i = 0; i < nc; ++i) load_k(i, 1);            // pass A: hi plane only
      load_k(0, NPL);
      for (int t = 0; t < nc; ++t) {
        if (t + 1 < nc) load_k(t + 1, NPL);
        load_v(t);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nc > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(AT_BQ, AT_BKEY);
      constexpr uint32_t idesc_o = make_idesc_bf16(AT_BQ, AT_D);
      const int ta[3] = {0, 0, 1}, tb[3] = {0, 1, 0};
      const uint32_t q_addr = smem_u32(sQ), p_addr0 = smem_u32(sP), ring_addr = smem_u32(sRing);
      mbar_wait(q_full, 0);
      tc_fence_after();
      uint32_t n = 0, sj = 0;                               // ring sequence number, score-tile job number
      auto issue_qk = [&](int nterm) {
        const uint32_t slot = n % 3u, st = sj & 1u;
        mbar_wait(&r_full[slot], (n / 3u) & 1u);
        mbar_wait(&s_empty[st], ((sj >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t k_addr = ring_addr + slot * SLOT_BYTES;
        const uint32_t d_s = tmem_base + st * AT_BKEY;
        for (int term = 0; term < nterm; ++term) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) {
            const uint64_t da = make_sw128_desc(q_addr + (ta[term] * 2 + (k >> 2)) * AT_Q_KBLK) + 2 * (k & 3);
            const uint64_t db = make_sw128_desc(k_addr + (tb[term] * 2 + (k >> 2)) * AT_K_KBLK) + 2 * (k & 3);
            umma_bf16(d_s, da, db, idesc_s, (term | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&s_full[st]);
        umma_commit(&r_empty[slot]);                        // K chunk is dead once its score MMAs retire
        ++n; ++sj;
      };
      auto issue_pv = [&](int t) {
        const uint32_t slot = n % 3u, pb = (uint32_t)t & 1u;
        mbar_wait(&p_full[pb], ((uint32_t)t >> 1) & 1u);
        mbar_wait(&r_full[slot], (n / 3u) & 1u);
        tc_fence_after();
        const uint32_t v_addr = ring_addr + slot * SLOT_BYTES, p_addr = p_addr0 + pb * P_BYTES;
        for (int term = 0; term < NT; ++term) {
          const uint64_t da = make_sw128_desc(p_addr + ta[term] * AT_P_TILE);
          const uint64_t db = make_sw128_desc(v_addr + tb[term] * AT_V_TILE);
#pragma unroll
          for (int k = 0; k < AT_BKEY / 16; ++k) umma_bf16(tmem_o, da + 2 * k, db + 2 * k, idesc_o, (t | term | k) != 0 ? 1u : 0u);
        }
        umma_commit(&r_empty[slot]);
        umma_commit(&p_empty[pb]);
        ++n;
      };
      for (int i = 0; i < nc; ++i) issue_qk(1);
      issue_qk(NT);
      for (int t = 0; t < nc; ++t) {
        if (t + 1 < nc) issue_qk(NT);
        issue_pv(t);
      }
      umma_commit(o_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue: two threads per query row =====================
    const int qw = warp & 3;                   // TMEM lane quarter this warp may access
    const int hf = (warp - 4) >> 2;            // column half of a 64-key chunk / of the 128 output dims
    const int r = qw * 32 + lane;              // row in tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(qw * 32) << 16;
    float m = -INFINITY, l = 0.f;
    if (nc > 0) {
      // ---- pass A: approximate row max over this thread's 32 columns of every chunk
      for (int i = 0; i < nc; ++i) {
        const int st = i & 1;
        mbar_wait(&s_full[st], ((uint32_t)i >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_addr + st * AT_BKEY + hf * 32, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);           // values are in registers: release the score tile first
        const int kbase = i * AT_BKEY + hf * 32;
        if (kbase + 32 <= klen) {                       // whole half-chunk valid (warp-uniform): no per-element predicates
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) m = fmaxf(m, __uint_as_float(v[jj]));
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) if (kbase + jj < klen) m = fmaxf(m, __uint_as_float(v[jj]));
        }
      }
      s_red[hf * 128 + r] = m;
      softmax_bar();
      m = fmaxf(m, s_red[(hf ^ 1) * 128 + r]);     // finite: key 0 is always valid when nc > 0
      // ---- pass B: probabilities
      for (int t = 0; t < nc; ++t) {
        const int i = nc + t, st = i & 1, pb = t & 1;
        mbar_wait(&s_full[st], ((uint32_t)i >> 1) & 1);
        tc_fence_after();
        float pr[32];
        {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + lane_addr + st * AT_BKEY + hf * 32, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[st]);
          const int kbase = t * AT_BKEY + hf * 32;
          if (kbase + 32 <= klen) {                     // fully valid half-chunk: straight-line exp / accumulate
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) { const float pv = __expf(__uint_as_float(v[jj]) - m); pr[jj] = pv; l += pv; }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
              const float pv = (kbase + jj < klen) ? __expf(__uint_as_float(v[jj]) - m) : 0.f;
              pr[jj] = pv;
              l += pv;
            }
          }
        }
        // P planes -> smem, K-major SWIZZLE_128B: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
        uint32_t* prow = reinterpret_cast<uint32_t*>(sP + pb * P_BYTES + (size_t)r * 128);
        mbar_wait(&p_empty[pb], (((uint32_t)t >> 1) & 1) ^ 1);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // packed cvt.rn.bf16x2.f32 (ALU pipe) instead of two scalar F2F.BF16 (quarter-rate XU pipe, shared with EX2);
            // bf16 -> fp32 is a 16-bit shift
            const float a = pr[cc * 8 + 2 * e], bb = pr[cc * 8 + 2 * e + 1];
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
            if (NPL > 1) {
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hi[e] << 16), bb - __uint_as_float(hi[e] & 0xFFFF0000u));
              lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
          }
          const int pc = ((hf * 4 + cc) ^ (r & 7)) * 4;
          *reinterpret_cast<uint4*>(prow + pc) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (NPL > 1) *reinterpret_cast<uint4*>(prow + AT_P_TILE / 4 + pc) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[pb]);
      }
      softmax_bar();                   // everyone has read the exchanged maxima before the slots are reused
      s_red[hf * 128 + r] = l;
      softmax_bar();
      l += s_red[(hf ^ 1) * 128 + r];
      mbar_wait(o_full, 0);
      tc_fence_after();
    }
    // ---- epilogue: O / l for this warp's 32 rows x 64 head dims, staged through shared memory (the K/V ring is dead once
    //      o_full has fired) so that the context rows leave as coalesced 8-byte (bf16 planes) / 16-byte (fp32) stores
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    float* stage = reinterpret_cast<float*>(sRing) + (warp - 4) * (32 * 36);
    const int64_t grow0 = (int64_t)b * p.tq + q0 + qw * 32;
#pragma unroll 1
    for (int c0 = hf * 64; c0 < hf * 64 + 64; c0 += 32) {
      {
        uint32_t v[32];
        if (nc > 0) tmem_ld_32x32(tmem_o + lane_addr + c0, v);
        float* srow = stage + lane * 36;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          float4 o4;
          o4.x = nc > 0 ? __uint_as_float(v[jj]) * inv : 0.f;
          o4.y = nc > 0 ? __uint_as_float(v[jj + 1]) * inv : 0.f;
          o4.z = nc > 0 ? __uint_as_float(v[jj + 2]) * inv : 0.f;
          o4.w = nc > 0 ? __uint_as_float(v[jj + 3]) * inv : 0.f;
          *reinterpret_cast<float4*>(srow + jj) = o4;
        }
      }
      __syncwarp();
      const int rr0 = lane >> 3, c4 = (lane & 7) * 4;
      const int col = h * AT_D + c0 + c4;
      const int rows_ok = p.tq - (q0 + qw * 32);                      // valid rows of this warp's 32
      float* pc = p.ctx ? p.ctx + (grow0 + rr0) * p.ldc + col : nullptr;
      __nv_bfloat16* pp = OPL > 0 ? p.ctx_planes + (grow0 + rr0) * p.ldp + col : nullptr;
      const int64_t plane = (int64_t)p.batch * p.tq * p.ldp;
      const float* sp = stage + rr0 * 36 + c4;
#pragma unroll 2
      for (int it = 0; it < 8; ++it) {
        if (it * 4 + rr0 < rows_ok) {
          const float4 o4 = *reinterpret_cast<const float4*>(sp + it * 4 * 36);
          if (pc) *reinterpret_cast<float4*>(pc + (int64_t)it * 4 * p.ldc) = o4;
          if (OPL > 0) {
            float x0 = o4.x, x1 = o4.y, x2 = o4.z, x3 = o4.w;
            __nv_bfloat16* dst = pp + (int64_t)it * 4 * p.ldp;
#pragma unroll
            for (int pl = 0; pl < OPL; ++pl) {
              const __nv_bfloat162 p01 = __floats2bfloat162_rn(x0, x1), p23 = __floats2bfloat162_rn(x2, x3);
              uint2 pk;
              pk.x = *reinterpret_cast<const uint32_t*>(&p01);
              pk.y = *reinterpret_cast<const uint32_t*>(&p23);
              *reinterpret_cast<uint2*>(dst) = pk;
              if (pl + 1 < OPL) {
                dst += plane;
                x0 -= __uint_as_float(pk.x << 16); x1 -= __uint_as_float(pk.x & 0xFFFF0000u);
                x2 -= __uint_as_float(pk.y << 16); x3 -= __uint_as_float(pk.y & 0xFFFF0000u);
              }
            }
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// V [B, tk, ldv] (head h at column h*128) -> Vt planes [npl][B*H*128][tkp] (keys contiguous), via a 64x64 smem transpose.
__global__ void __launch_bounds__(256)
vt_planes_kernel(const float* __restrict__ v, int64_t ldv, int tk, int tkp, int heads, int nplanes, int64_t plane_elems,
                 __nv_bfloat16* __restrict__ vt) {
  __shared__ float tile[64][65];
  const int t0 = blockIdx.x * 64, dblk = blockIdx.y, b = blockIdx.z;   // dblk over heads*2 (64-wide d blocks)
  const int tid = threadIdx.x;
  for (int idx = tid; idx < 64 * 16; idx += 256) {
    const int tr = idx >> 4, c4 = idx & 15;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + tr < tk) x = __ldg(reinterpret_cast<const float4*>(v + ((int64_t)b * tk + t0 + tr) * ldv + dblk * 64 + 4 * c4));
    tile[tr][4 * c4 + 0] = x.x; tile[tr][4 * c4 + 1] = x.y; tile[tr][4 * c4 + 2] = x.z; tile[tr][4 * c4 + 3] = x.w;
  }
  __syncthreads();
  // each thread writes 2 consecutive keys for one d: 64 d x 32 key-pairs = 2048 items / 256 threads
  for (int idx = tid; idx < 64 * 32; idx += 256) {
    const int d = idx >> 5, kp = (idx & 31) * 2;
    if (t0 + kp >= tkp) continue;
    float a = tile[kp][d], bb = tile[kp + 1][d];
    __nv_bfloat16* dst = vt + ((int64_t)b * heads * AT_D + dblk * 64 + d) * tkp + t0 + kp;
    for (int pl = 0; pl < nplanes; ++pl) {
      const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(bb);
      *reinterpret_cast<__nv_bfloat162*>(dst + pl * plane_elems) = __halves2bfloat162(ha, hb);
      a -= __bfloat162float(ha); bb -= __bfloat162float(hb);
    }
  }
}

// fp32 [rows, cols] (ld) * scale -> bf16 planes [nplanes][rows][cols]
__global__ void __launch_bounds__(256)
scale_split_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols, float scale, int nplanes,
                   __nv_bfloat16* __restrict__ planes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4n = cols >> 2;
  if (i >= rows * c4n) return;
  const int64_t r = i / c4n;
  const int c = (int)(i - r * c4n) * 4;
  const float4 x = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
  float v[4] = {__fmul_rn(x.x, scale), __fmul_rn(x.y, scale), __fmul_rn(x.z, scale), __fmul_rn(x.w, scale)};
  const int64_t plane = rows * cols;
  for (int pl = 0; pl < nplanes; ++pl) {
    __nv_bfloat16 hh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { hh[k] = __float2bfloat16_rn(v[k]); v[k] -= __bfloat162float(hh[k]); }
    __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(planes + pl * plane + r * cols + c);
    dst[0] = __halves2bfloat162(hh[0], hh[1]);
    dst[1] = __halves2bfloat162(hh[2], hh[3]);
  }
}

size_t attention_tc_scratch_bytes(int batch, int heads, int tq, int tk, int mode) {
  if (mode == FA_GEMM_F32_SIMT) return 0;
  const int npl = mode == FA_GEMM_BF16X1 ? 1 : 2;
  const int tkp = (tk + 63) / 64 * 64;
  const int d = heads * AT_D;
  return (size_t)npl * 2 * ((size_t)batch * tq * d + (size_t)batch * tk * d + (size_t)batch * d * tkp) + 4096;
}

int attention_tc_launch(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        const int32_t* key_lens, int batch, int heads, int tq, int tk, float* ctx, int64_t ldc,
                        __nv_bfloat16* ctx_planes, int64_t ldp, int out_nplanes, int mode, Arena* scratch, cudaStream_t st) {
  if (batch <= 0 || tq <= 0) return FA_OK;
  if (!q || !k || !v || !key_lens || tk <= 0 || !scratch) return FA_ERR_ARG;
  if ((ldq | ldk | ldv) & 3) return FA_ERR_UNSUPPORTED;
  const int npl = mode == FA_GEMM_BF16X1 ? 1 : 2;
  const int d = heads * AT_D;
  const int tkp = (tk + 63) / 64 * 64;
  const int64_t mq = (int64_t)batch * tq, mk = (int64_t)batch * tk, mv = (int64_t)batch * d;
  Arena local(scratch->base, scratch->cap);
  __nv_bfloat16* qp = local.take<__nv_bfloat16>((size_t)npl * mq * d);
  __nv_bfloat16* kp = local.take<__nv_bfloat16>((size_t)npl * mk * d);
  __nv_bfloat16* vt = local.take<__nv_bfloat16>((size_t)npl * mv * tkp);
  if (!local.ok()) return FA_ERR_WORKSPACE;
  const float qscale = (float)(1.0 / sqrt((double)AT_D));
  {
    const int64_t tot = mq * (d / 4);
    scale_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(q, ldq, mq, d, qscale, npl, qp);
    FA_CHECK_LAUNCH();
    const int64_t totk = mk * (d / 4);
    scale_split_kernel<<<(unsigned)((totk + 255) / 256), 256, 0, st>>>(k, ldk, mk, d, 1.0f, npl, kp);
    FA_CHECK_LAUNCH();
    dim3 g((tkp + 63) / 64, heads * 2, batch);
    vt_planes_kernel<<<g, 256, 0, st>>>(v, ldv, tk, tkp, heads, npl, mv * tkp, vt);
    FA_CHECK_LAUNCH();
  }
  return attention_tc_planes_launch(qp, kp, vt, key_lens, batch, heads, tq, tk, ctx, ldc, ctx_planes, ldp, out_nplanes, mode, st);
}

template <int NPL, int OPL>
static int launch_att(dim3 grid, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttTcParams& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)NPL * (2 * AT_Q_KBLK + 3 * AT_V_TILE + 2 * AT_P_TILE) + 1024 + 256 + 1024;
  static bool done = false;
  if (!done) { FA_CUDA_OK(cudaFuncSetAttribute(attention_tc_kernel<NPL, OPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
  attention_tc_kernel<NPL, OPL><<<grid, 384, smem, st>>>(mq, mk, mv, p);
  return FA_OK;
}

// Operand planes already in place (written by the producing GEMMs' epilogues, gemm_tc.cu AttnSinks):
// qp [npl][B*tq][H*128] (scaled), kp [npl][B*tk][H*128], vt [npl][B*H*128][round_up(tk,64)].
int attention_tc_planes_launch(const __nv_bfloat16* qp, const __nv_bfloat16* kp, const __nv_bfloat16* vt, const int32_t* key_lens,
                               int batch, int heads, int tq, int tk, float* ctx, int64_t ldc, __nv_bfloat16* ctx_planes,
                               int64_t ldp, int out_nplanes, int mode, cudaStream_t st) {
  if (batch <= 0 || tq <= 0) return FA_OK;
  if (!qp || !kp || !vt || !key_lens || tk <= 0) return FA_ERR_ARG;
  const int npl = mode == FA_GEMM_BF16X1 ? 1 : 2;
  const int d = heads * AT_D;
  const int tkp = (tk + 63) / 64 * 64;
  const int64_t mq = (int64_t)batch * tq, mk = (int64_t)batch * tk, mv = (int64_t)batch * d;
  CUtensorMap mq_map, mk_map, mv_map;
  FA_RETURN_IF_ERR(make_bf16_map(&mq_map, qp, (uint64_t)mq * npl, (uint64_t)d, (uint64_t)d, AT_BQ));
  FA_RETURN_IF_ERR(make_bf16_map(&mk_map, kp, (uint64_t)mk * npl, (uint64_t)d, (uint64_t)d, AT_BKEY));
  FA_RETURN_IF_ERR(make_bf16_map(&mv_map, vt, (uint64_t)mv * npl, (uint64_t)tk, (uint64_t)tkp, AT_D));
  AttTcParams p;
  p.tq = tq; p.tk = tk; p.heads = heads; p.batch = batch; p.key_lens = key_lens;
  p.q_plane_rows = mq; p.k_plane_rows = mk; p.v_plane_rows = mv;
  p.ctx = ctx; p.ldc = ldc; p.ctx_planes = ctx_planes; p.ldp = ldp; p.out_nplanes = out_nplanes;
  dim3 grid((tq + AT_BQ - 1) / AT_BQ, heads, batch);
  const int opl = ctx_planes ? out_nplanes : 0;
  if (ctx_planes && (opl < 1 || opl > 3)) return FA_ERR_ARG;
  if (npl == 1) {
    if (opl > 1) return FA_ERR_UNSUPPORTED;
    FA_RETURN_IF_ERR(opl == 0 ? (launch_att<1, 0>(grid, mq_map, mk_map, mv_map, p, st)) : (launch_att<1, 1>(grid, mq_map, mk_map, mv_map, p, st)));
  } else {
    if (opl == 1) return FA_ERR_UNSUPPORTED;
    FA_RETURN_IF_ERR(opl == 0 ? (launch_att<2, 0>(grid, mq_map, mk_map, mv_map, p, st))
                     : opl == 2 ? (launch_att<2, 2>(grid, mq_map, mk_map, mv_map, p, st))
                                : (launch_att<2, 3>(grid, mq_map, mk_map, mv_map, p, st)));
  }
  FA_CHECK_LAUNCH();
  return FA_OK;
}

}  // namespace fa

extern "C" size_t fa_attention_tc_workspace_bytes(int32_t batch, int32_t heads, int32_t tq, int32_t tk, int32_t gemm_mode) {
  return fa::attention_tc_scratch_bytes(batch, heads, tq, tk, gemm_mode);
}

extern "C" int fa_attention_tc(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                               const int32_t* key_lens, int32_t batch, int32_t heads, int32_t tq, int32_t tk, float* ctx,
                               int64_t ld_ctx, int32_t gemm_mode, void* workspace, size_t ws_bytes, fa_stream_t stream) {
  if (!ctx || heads * fa::AT_D > 4096 || (ld_ctx & 3)) return FA_ERR_ARG;
  if (gemm_mode == FA_GEMM_F32_SIMT) return FA_ERR_ARG;
  fa::Arena scratch(workspace, ws_bytes);
  return fa::attention_tc_launch(q, ldq, k, ldk, v, ldv, key_lens, batch, heads, tq, tk, ctx, ld_ctx, nullptr, 0, 0, gemm_mode,
                                 &scratch, (cudaStream_t)stream);
}
