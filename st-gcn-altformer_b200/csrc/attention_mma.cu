// Tensor-core small-sequence attention (bf16, mma.sync.m16n8k16 + ldmatrix), forward and backward.
// Reference: model/AltFormer/model_ST.py:49-67.  Sequences are 22..64 tokens long, so one warp owns one
// (sequence, head): S = QK^T, softmax, PV -- and in backward dP, dS, dQ, dK, dV -- all stay in registers
// and shared memory; a CTA owns HG heads of one sequence so that global traffic is whole 256/512-byte row
// segments (coalesced 16-byte accesses), read once and written once.
//
// Fragment conventions (PTX ISA, m16n8k16, g = lane / 4, t = lane % 4):
//   A (16x16 row):  a0 (g, 2t..2t+1)  a1 (g+8, 2t..)  a2 (g, 2t+8..)  a3 (g+8, 2t+8..)
//   B (16x8  col):  b0 (k = 2t..2t+1, n = g)          b1 (k = 2t+8.., n = g)
//   C (16x8):       c0,c1 (g, 2t..2t+1)               c2,c3 (g+8, 2t..2t+1)
// ldmatrix (non-transposed) of an 8x8 b16 tile gives thread (row g, cols 2t..2t+1); .trans gives
// (rows 2t..2t+1, col g).  An accumulator tile pair (two n8 tiles) is therefore already an A fragment.
#include <stdlib.h>

#include "common.cuh"

namespace afb {
namespace attn_mma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}

// 2^x on the SFU without the libm range fix-ups (arguments are <= 0 here; -inf -> 0, denormal results flush)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Geometry of one CTA's shared memory (all pitches are an odd number of 16-byte chunks => conflict-free
// ldmatrix and 16-byte row copies).
template <int DH, int LP, int HG> struct Geo {
  static constexpr int W = HG * DH;             // columns per q / k / v part handled by this CTA
  static constexpr int QKV_PITCH = 3 * W + 8;   // elements
  static constexpr int O_PITCH = W + 8;
  static constexpr int S_PITCH = LP + 8;
  static constexpr int kThreads = HG * 32;
  static constexpr size_t fwd_bytes = (size_t)LP * QKV_PITCH * 2;   // O is staged over the consumed Q rows
  static constexpr size_t bwd_tile_bytes = ((size_t)LP * QKV_PITCH + (size_t)LP * O_PITCH) * 2;   // q|k|v + dO
  static constexpr size_t bwd_bytes = bwd_tile_bytes + (size_t)HG * 2 * LP * S_PITCH * 2;            // + P / dS staging
};

__host__ __device__ constexpr int bwd_wph(int LP) { return LP >= 32 ? 2 : 1; }   // backward: warps per head

// rows [0, L) of PARTS W-wide column windows of a [.., ld] matrix -> smem tile; rows [L, LP) zeroed.
// One warp per (row, part): the pair index is warp-uniform, so the address arithmetic is a handful of uniform
// instructions per 512 contiguous bytes (the previous per-16-byte-chunk div/mod made the copy loop the top
// issue-slot consumer of the forward kernel, ncu r01).
template <int W, int PITCH, int PARTS, int LP, int NT>
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int ld, int part_stride, int L) {
  constexpr int C16 = W / 8;  // 16-byte chunks per part row
  constexpr int NW = NT / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 32-bit offsets from two bases (a tile spans < 2^31 elements): one 64-bit add per copy instead of 64-bit multiplies
  const uint32_t d0 = smem_u32(dst) + (uint32_t)lane * 16u;
  const bf16* s0 = src + lane * 8;
#pragma unroll
  for (int j0 = 0; j0 < LP * PARTS; j0 += NW) {
    const int j = j0 + warp;
    if ((LP * PARTS) % NW != 0 && j >= LP * PARTS) break;
    const int row = j / PARTS, part = j - row * PARTS;
    const uint32_t d = d0 + (uint32_t)((row * PITCH + part * W) * 2);
    const int soff = row * ld + part * part_stride;
#pragma unroll
    for (int c = 0; c < C16; c += 32) {
      if (C16 % 32 != 0 && lane + c >= C16) break;
      if (row < L)   // asynchronous 16-byte copies: every chunk of the tile is in flight at once (cp_async_wait_all below)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + c * 16), "l"(s0 + soff + c * 8) : "memory");
      else
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(d + c * 16), "r"(0) : "memory");
    }
  }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
template <int W, int PITCH>
__device__ __forceinline__ void store_tile(bf16* dst, int ld, const bf16* src, int L, int nthreads) {
  constexpr int C16 = W / 8;
  for (int idx = threadIdx.x; idx < L * C16; idx += nthreads) {
    const int row = idx / C16, c = idx % C16;
    *reinterpret_cast<uint4*>(dst + (row * ld + c * 8)) = *reinterpret_cast<const uint4*>(src + (row * PITCH + c * 8));
  }
}

// Lane-dependent shared-memory byte offsets of the four access patterns, computed ONCE per thread; every ldmatrix /
// store below is then base + lane offset + compile-time constant (the per-call pointer arithmetic used to be ~40 % of
// the forward kernel's instructions, ncu r01).
//   a : A operand, non-transposed   rows r0 + (lane & 15),                 cols c0 + (lane >> 4) * 8
//   k : B operand from [n][k] rows  rows n0 + (lane & 7) + (id >> 1) * 8,  cols c0 + (id & 1) * 8        (id = lane >> 3)
//   bt: B operand from [k][n] rows (transposed load)  rows k0 + (lane & 7) + (id & 1) * 8, cols c0 + (id >> 1) * 8
//   c : C fragment element (g, 2t)   rows r0 + g (+8), cols c0 + 2t
template <int PITCH> struct LaneOff {
  uint32_t a, k, bt, c;
  __device__ __forceinline__ LaneOff() {
    const int lane = threadIdx.x & 31, id = lane >> 3;
    a = (uint32_t)(((lane & 15) * PITCH + (lane >> 4) * 8) * 2);
    k = (uint32_t)((((lane & 7) + (id >> 1) * 8) * PITCH + (id & 1) * 8) * 2);
    bt = (uint32_t)((((lane & 7) + (id & 1) * 8) * PITCH + (id >> 1) * 8) * 2);
    c = (uint32_t)(((lane >> 2) * PITCH + 2 * (lane & 3)) * 2);
  }
};
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// S tile for query rows [mi*16, mi*16+16): s[nj][0..3] = (Q K^T) for key tile nj, then masked softmax.
// Returns the normalised probabilities in s (fp32), columns >= L are exactly 0.
// qbase / kbase: shared-memory byte addresses of the head's Q / K columns with the lane offsets (a / k) already added.
template <int DH, int LP, int PITCH>
__device__ __forceinline__ void scores_softmax(uint32_t qbase, uint32_t kbase, int mi, int L, float scale_log2, float (&s)[LP / 8][4]) {
  const int t = threadIdx.x & 3;
  uint32_t qa[DH / 16][4];
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) ldsm_x4(qbase + (uint32_t)((mi * 16 * PITCH + kk * 16) * 2), qa[kk]);
#pragma unroll
  for (int nj = 0; nj < LP / 8; ++nj)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[nj][e] = 0.f;
  const int nt_valid = (L + 7) >> 3;   // key tiles holding at least one real key (warp-uniform)
#pragma unroll
  for (int n2 = 0; n2 < LP / 16; ++n2) {
    if (2 * n2 >= nt_valid) break;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t kb[4];
      ldsm_x4(kbase + (uint32_t)((n2 * 16 * PITCH + kk * 16) * 2), kb);
      mma(s[2 * n2], qa[kk], kb[0], kb[1]);
      if (2 * n2 + 1 < nt_valid) mma(s[2 * n2 + 1], qa[kk], kb[2], kb[3]);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nj = 0; nj < LP / 8; ++nj) {
    const int c = nj * 8 + 2 * t;
    if (c >= L) { s[nj][0] = -INFINITY; s[nj][2] = -INFINITY; }
    if (c + 1 >= L) { s[nj][1] = -INFINITY; s[nj][3] = -INFINITY; }
    mx0 = fmaxf(mx0, fmaxf(s[nj][0], s[nj][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nj][2], s[nj][3]));
  }
  mx0 = quad_max(mx0);
  mx1 = quad_max(mx1);
  float sum0 = 0.f, sum1 = 0.f;
  const float nm0 = -mx0 * scale_log2, nm1 = -mx1 * scale_log2;   // one FFMA + one MUFU.EX2 per element
#pragma unroll
  for (int nj = 0; nj < LP / 8; ++nj) {
    s[nj][0] = ex2_fast(fmaf(s[nj][0], scale_log2, nm0));
    s[nj][1] = ex2_fast(fmaf(s[nj][1], scale_log2, nm0));
    s[nj][2] = ex2_fast(fmaf(s[nj][2], scale_log2, nm1));
    s[nj][3] = ex2_fast(fmaf(s[nj][3], scale_log2, nm1));
    sum0 += s[nj][0] + s[nj][1];
    sum1 += s[nj][2] + s[nj][3];
  }
  const float inv0 = 1.0f / quad_sum(sum0), inv1 = 1.0f / quad_sum(sum1);
#pragma unroll
  for (int nj = 0; nj < LP / 8; ++nj) {
    s[nj][0] *= inv0; s[nj][1] *= inv0;
    s[nj][2] *= inv1; s[nj][3] *= inv1;
  }
}

// acc[nd][..] += A(16 x 16*NK16) * B, B[k][n] = rows k, cols n of a row-major tile (transposed load).
// bbase: byte address of the tile's (row 0, first column) with the lane offset (bt) already added.
template <int NK16, int ND8, int PITCH>
__device__ __forceinline__ void mma_a_regs_bt(const uint32_t (&a)[NK16][4], uint32_t bbase, float (&acc)[ND8][4]) {
#pragma unroll
  for (int kk = 0; kk < NK16; ++kk)
#pragma unroll
    for (int n2 = 0; n2 < ND8 / 2; ++n2) {
      uint32_t b[4];
      ldsm_x4_t(bbase + (uint32_t)((kk * 16 * PITCH + n2 * 16) * 2), b);
      mma(acc[2 * n2], a[kk], b[0], b[1]);
      mma(acc[2 * n2 + 1], a[kk], b[2], b[3]);
    }
}

// cbase: byte address of (row 0, first column) of the destination with the lane offset (c) already added
template <int ND8, int PITCH>
__device__ __forceinline__ void store_acc(uint32_t cbase, int row0, const float (&acc)[ND8][4], float m0, float m1) {
#pragma unroll
  for (int nd = 0; nd < ND8; ++nd) {
    sts32(cbase + (uint32_t)((row0 * PITCH + nd * 8) * 2), pack2(acc[nd][0] * m0, acc[nd][1] * m0));
    sts32(cbase + (uint32_t)(((row0 + 8) * PITCH + nd * 8) * 2), pack2(acc[nd][2] * m1, acc[nd][3] * m1));
  }
}

// ---------------------------------------------------------------------------------------------
template <int DH, int LP, int HG>
__global__ void __launch_bounds__(HG * 32) attn_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, int L, int heads, float scale,
                                                               const float* __restrict__ out_scale) {
  using G = Geo<DH, LP, HG>;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sq = reinterpret_cast<bf16*>(smraw);
  const int groups = heads / HG;
  const int64_t b = blockIdx.x / groups;
  const int h0 = (blockIdx.x % groups) * HG;
  const int D = heads * DH;
  const float os = out_scale != nullptr ? out_scale[b] : 1.f;   // DropPath keep factor of this sequence
  if (os == 0.f) {   // dropped sequence: zeros, nothing to compute
    bf16* dst = o + b * L * D + h0 * DH;
    for (int idx = threadIdx.x; idx < L * (G::W / 8); idx += G::kThreads)
      *reinterpret_cast<uint4*>(dst + (idx / (G::W / 8)) * (int64_t)D + (idx % (G::W / 8)) * 8) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  load_tile<G::W, G::QKV_PITCH, 3, LP, G::kThreads>(sq, qkv + b * L * 3 * D + h0 * DH, 3 * D, D, L);
  cp_async_wait_all();
  __syncthreads();
  const int hw = threadIdx.x >> 5;
  const int qcol = hw * DH, kcol = G::W + hw * DH, vcol = 2 * G::W + hw * DH;
  const float scale_log2 = scale * 1.4426950408889634f;
  const LaneOff<G::QKV_PITCH> lo;
  const uint32_t sq32 = smem_u32(sq);
  const uint32_t qbase = sq32 + lo.a + qcol * 2, kbase = sq32 + lo.k + kcol * 2, vbase = sq32 + lo.bt + vcol * 2;
  const uint32_t obase = sq32 + lo.c + qcol * 2;
#pragma unroll
  for (int mi = 0; mi < LP / 16; ++mi) {   // unrolled: every ldmatrix / store address becomes base + immediate
    if (mi * 16 >= L) break;
    float s[LP / 8][4];
    scores_softmax<DH, LP, G::QKV_PITCH>(qbase, kbase, mi, L, scale_log2, s);
    uint32_t pa[LP / 16][4];
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk) {
      pa[kk][0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      pa[kk][1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      pa[kk][2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[kk][3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
    float acc[DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < DH / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    mma_a_regs_bt<LP / 16, DH / 8, G::QKV_PITCH>(pa, vbase, acc);
    __syncwarp();   // every lane has consumed this tile's Q fragments: the rows can take the output
    store_acc<DH / 8, G::QKV_PITCH>(obase, mi * 16, acc, os, os);
  }
  __syncthreads();
  store_tile<G::W, G::QKV_PITCH>(o + b * L * D + h0 * DH, D, sq, L, G::kThreads);
}

// ---------------------------------------------------------------------------------------------
// Two warps per head (LP >= 32): each warp owns every other 16-row tile of every phase, so a CTA carries twice the
// warps for the same shared memory (the kernel is latency-bound at 2 CTAs x 8 warps per SM); the pair meets at a
// named barrier wherever one phase overwrites an operand the partner may still be reading.
template <int DH, int LP, int HG>
__global__ void __launch_bounds__(HG * 32 * bwd_wph(LP), bwd_wph(LP)) attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dO, bf16* __restrict__ dqkv,
                                                               int L, int heads, float scale) {
  using G = Geo<DH, LP, HG>;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sq = reinterpret_cast<bf16*>(smraw);           // q | k | v   ->  q | dK | dV
  bf16* sdo = sq + LP * G::QKV_PITCH;                  // dO          ->  dQ
  constexpr int WPH = bwd_wph(LP), NT = HG * 32 * WPH;
  const int hw = (threadIdx.x >> 5) / WPH, half = (threadIdx.x >> 5) % WPH;
  auto pair_sync = [&]() {
    if (WPH == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + hw), "r"(32 * WPH) : "memory");
  };
  bf16* sP = sdo + LP * G::O_PITCH + hw * 2 * LP * G::S_PITCH;   // per-warp staging of P and dS (bf16)
  bf16* sdS = sP + LP * G::S_PITCH;
  const int groups = heads / HG;
  const int64_t b = blockIdx.x / groups;
  const int h0 = (blockIdx.x % groups) * HG;
  const int D = heads * DH;
  load_tile<G::W, G::QKV_PITCH, 3, LP, NT>(sq, qkv + b * L * 3 * D + h0 * DH, 3 * D, D, L);
  load_tile<G::W, G::O_PITCH, 1, LP, NT>(sdo, dO + b * L * D + h0 * DH, D, 0, L);
  cp_async_wait_all();
  __syncthreads();
  const int qcol = hw * DH, kcol = G::W + hw * DH, vcol = 2 * G::W + hw * DH, ocol = hw * DH;
  const float scale_log2 = scale * 1.4426950408889634f;
  const LaneOff<G::QKV_PITCH> lq;
  const LaneOff<G::O_PITCH> ld;
  const LaneOff<G::S_PITCH> ls;
  const uint32_t sq32 = smem_u32(sq), sdo32 = smem_u32(sdo), sP32 = smem_u32(sP), sdS32 = smem_u32(sdS);
  // This warp's 16-row tiles are half, half + WPH, ...: the runtime part (half) is folded into the bases once, the
  // loops below run over a compile-time index, so every shared-memory address is base + immediate.
  constexpr int NTILE = (LP / 16 + WPH - 1) / WPH;
  const uint32_t hq = (uint32_t)(half * 16 * G::QKV_PITCH * 2), ho = (uint32_t)(half * 16 * G::O_PITCH * 2);
  const uint32_t hs = (uint32_t)(half * 16 * G::S_PITCH * 2), hc = (uint32_t)(half * 16 * 2);
  auto live = [&](int i) { return (LP / 16) % WPH == 0 || half + i * WPH < LP / 16; };

  // ---- phase A: P and dS = P * (dP - delta) * scale for every query tile -> staging ----
#pragma unroll
  for (int i = 0; i < NTILE; ++i) {
    if (!live(i)) break;
    const int mi = i * WPH;   // tile offset beyond `half` (compile-time after unrolling)
    float s[LP / 8][4];
    scores_softmax<DH, LP, G::QKV_PITCH>(sq32 + lq.a + qcol * 2 + hq, sq32 + lq.k + kcol * 2, mi, L, scale_log2, s);
    uint32_t da[DH / 16][4];
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk)
      ldsm_x4(sdo32 + ld.a + ho + (uint32_t)((mi * 16 * G::O_PITCH + ocol + kk * 16) * 2), da[kk]);
    float dp[LP / 8][4];
#pragma unroll
    for (int nj = 0; nj < LP / 8; ++nj)
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[nj][e] = 0.f;
    const int nt_valid = (L + 7) >> 3;   // key tiles with at least one real key; P is exactly 0 beyond them
#pragma unroll
    for (int n2 = 0; n2 < LP / 16; ++n2) {
      if (2 * n2 >= nt_valid) break;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {  // dP = dO V^T : B[k = d][n = key] = V[key][d] (non-transposed load)
        uint32_t vb[4];
        ldsm_x4(sq32 + lq.k + (uint32_t)((n2 * 16 * G::QKV_PITCH + vcol + kk * 16) * 2), vb);
        mma(dp[2 * n2], da[kk], vb[0], vb[1]);
        if (2 * n2 + 1 < nt_valid) mma(dp[2 * n2 + 1], da[kk], vb[2], vb[3]);
      }
    }
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nj = 0; nj < LP / 8; ++nj) {
      d0 += s[nj][0] * dp[nj][0] + s[nj][1] * dp[nj][1];
      d1 += s[nj][2] * dp[nj][2] + s[nj][3] * dp[nj][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
#pragma unroll
    for (int nj = 0; nj < LP / 8; ++nj) {
      const uint32_t off0 = ls.c + hs + (uint32_t)((mi * 16 * G::S_PITCH + nj * 8) * 2), off1 = off0 + 8 * G::S_PITCH * 2;
      sts32(sP32 + off0, pack2(s[nj][0], s[nj][1]));
      sts32(sP32 + off1, pack2(s[nj][2], s[nj][3]));
      sts32(sdS32 + off0, pack2(s[nj][0] * (dp[nj][0] - d0) * scale, s[nj][1] * (dp[nj][1] - d0) * scale));
      sts32(sdS32 + off1, pack2(s[nj][2] * (dp[nj][2] - d1) * scale, s[nj][3] * (dp[nj][3] - d1) * scale));
    }
  }
  pair_sync();

  // ---- phase B1: dV = P^T dO  (A = P^T via transposed loads of the staged P) -> v columns ----
#pragma unroll
  for (int i = 0; i < NTILE; ++i) {
    if (!live(i)) break;
    const int mj = i * WPH;
    uint32_t a[LP / 16][4];
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk)  // a0:(key 0-7, q 0-7) a1:(key 8-15, q 0-7) a2:(key 0-7, q 8-15) a3:(key 8-15, q 8-15)
      ldsm_x4_t(sP32 + ls.k + hc + (uint32_t)((kk * 16 * G::S_PITCH + mj * 16) * 2), a[kk]);
    float acc[DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < DH / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    mma_a_regs_bt<LP / 16, DH / 8, G::O_PITCH>(a, sdo32 + ld.bt + ocol * 2, acc);
    store_acc<DH / 8, G::QKV_PITCH>(sq32 + lq.c + vcol * 2 + hq, mj * 16, acc, 1.f, 1.f);
  }
  pair_sync();
  // ---- phase B2: dQ = dS K  -> the (now free) dO columns of this head ----
#pragma unroll
  for (int i = 0; i < NTILE; ++i) {
    if (!live(i)) break;
    const int mi = i * WPH;
    uint32_t a[LP / 16][4];
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk)
      ldsm_x4(sdS32 + ls.a + hs + (uint32_t)((mi * 16 * G::S_PITCH + kk * 16) * 2), a[kk]);
    float acc[DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < DH / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    mma_a_regs_bt<LP / 16, DH / 8, G::QKV_PITCH>(a, sq32 + lq.bt + kcol * 2, acc);
    store_acc<DH / 8, G::O_PITCH>(sdo32 + ld.c + ocol * 2 + ho, mi * 16, acc, 1.f, 1.f);
  }
  pair_sync();
  // ---- phase B3: dK = dS^T Q  -> k columns (K is no longer needed by this warp) ----
#pragma unroll
  for (int i = 0; i < NTILE; ++i) {
    if (!live(i)) break;
    const int mj = i * WPH;
    uint32_t a[LP / 16][4];
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk)
      ldsm_x4_t(sdS32 + ls.k + hc + (uint32_t)((kk * 16 * G::S_PITCH + mj * 16) * 2), a[kk]);
    float acc[DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < DH / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    mma_a_regs_bt<LP / 16, DH / 8, G::QKV_PITCH>(a, sq32 + lq.bt + qcol * 2, acc);
    store_acc<DH / 8, G::QKV_PITCH>(sq32 + lq.c + kcol * 2 + hq, mj * 16, acc, 1.f, 1.f);
  }
  __syncthreads();
  bf16* out = dqkv + b * L * 3 * D + h0 * DH;
  store_tile<G::W, G::O_PITCH>(out, 3 * D, sdo, L, NT);                      // dQ
  store_tile<G::W, G::QKV_PITCH>(out + D, 3 * D, sq + G::W, L, NT);          // dK
  store_tile<G::W, G::QKV_PITCH>(out + 2 * D, 3 * D, sq + 2 * G::W, L, NT);  // dV
}

// ---------------------------------------------------------------------------------------------
// Backward for LP = 32 (L <= 32: the SHREC / DHG spatial and temporal stages), one warp per head, NO staging buffers:
//   * both 16-row query tiles are processed together, so every K / V / dO / Q operand fragment is loaded from shared
//     memory once and feeds the MMAs of both tiles (the two-warp variant above loads each fragment per tile);
//   * P and dS stay in registers: as A operands they are the packed C fragments (FA2 reuse, for dQ = dS K), and their
//     transposes (for dV = P^T dO, dK = dS^T Q) come from movmatrix on the same registers.
// The older kernel's bound was the shared-memory data pipe (66 % of peak, ncu r01); this one moves ~40 % fewer
// wavefronts and needs 41 KB less shared memory per CTA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t movm_t(uint32_t v) {   // transpose of an 8x8 b16 tile held in mma fragment layout
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

template <int DH, int HG>
__global__ void __launch_bounds__(HG * 32, 2) attn_bwd32_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dO, bf16* __restrict__ dqkv,
                                                                int L, int heads, float scale) {
  constexpr int LP = 32;
  using G = Geo<DH, LP, HG>;
  constexpr int NT = HG * 32, KD = DH / 16, ND = DH / 8;
  constexpr int QP = G::QKV_PITCH, OP = G::O_PITCH;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sq = reinterpret_cast<bf16*>(smraw);           // q | k | v   ->  q | dK | dV
  bf16* sdo = sq + LP * QP;                            // dO          ->  dQ
  const int hw = threadIdx.x >> 5, t = threadIdx.x & 3;
  const int groups = heads / HG;
  const int64_t b = blockIdx.x / groups;
  const int h0 = (blockIdx.x % groups) * HG;
  const int D = heads * DH;
  load_tile<G::W, QP, 3, LP, NT>(sq, qkv + b * L * 3 * D + h0 * DH, 3 * D, D, L);
  load_tile<G::W, OP, 1, LP, NT>(sdo, dO + b * L * D + h0 * DH, D, 0, L);
  cp_async_wait_all();
  __syncthreads();
  const int qcol = hw * DH, kcol = G::W + hw * DH, vcol = 2 * G::W + hw * DH, ocol = hw * DH;
  const float scale_log2 = scale * 1.4426950408889634f;
  const LaneOff<QP> lq;
  const LaneOff<OP> ld;
  const uint32_t sq32 = smem_u32(sq), sdo32 = smem_u32(sdo);
  const int nt_valid = (L + 7) >> 3;   // key tiles holding at least one real key (warp-uniform)

  // ---- phase A: S = Q K^T and dP = dO V^T for both query tiles; one K / V fragment load serves both ----
  float s[2][4][4], dp[2][4][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[m][n][e] = 0.f; dp[m][n][e] = 0.f; }
#pragma unroll
  for (int kd = 0; kd < KD; ++kd) {
    uint32_t qa[2][4], da[2][4];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      ldsm_x4(sq32 + lq.a + (uint32_t)((m * 16 * QP + qcol + kd * 16) * 2), qa[m]);
      ldsm_x4(sdo32 + ld.a + (uint32_t)((m * 16 * OP + ocol + kd * 16) * 2), da[m]);
    }
#pragma unroll
    for (int n2 = 0; n2 < 2; ++n2) {
      if (2 * n2 >= nt_valid) break;
      uint32_t kb[4], vb[4];
      ldsm_x4(sq32 + lq.k + (uint32_t)((n2 * 16 * QP + kcol + kd * 16) * 2), kb);
      ldsm_x4(sq32 + lq.k + (uint32_t)((n2 * 16 * QP + vcol + kd * 16) * 2), vb);
      const bool second = 2 * n2 + 1 < nt_valid;
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma(s[m][2 * n2], qa[m], kb[0], kb[1]);
        mma(dp[m][2 * n2], da[m], vb[0], vb[1]);
        if (second) {
          mma(s[m][2 * n2 + 1], qa[m], kb[2], kb[3]);
          mma(dp[m][2 * n2 + 1], da[m], vb[2], vb[3]);
        }
      }
    }
  }
  // softmax rows, delta = sum_j P dP, dS = P (dP - delta) scale; P / dS packed as A fragments (keys = K dimension):
  //   pa[m][kk][0..3] = { n-tile 2kk rows g, n-tile 2kk rows g+8, n-tile 2kk+1 rows g, n-tile 2kk+1 rows g+8 }
  uint32_t pa[2][2][4], sa[2][2][4];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int c = n * 8 + 2 * t;
      if (c >= L) { s[m][n][0] = -INFINITY; s[m][n][2] = -INFINITY; }
      if (c + 1 >= L) { s[m][n][1] = -INFINITY; s[m][n][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[m][n][0], s[m][n][1]));
      mx1 = fmaxf(mx1, fmaxf(s[m][n][2], s[m][n][3]));
    }
    mx0 = quad_max(mx0);
    mx1 = quad_max(mx1);
    const float nm0 = -mx0 * scale_log2, nm1 = -mx1 * scale_log2;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[m][n][0] = ex2_fast(fmaf(s[m][n][0], scale_log2, nm0));
      s[m][n][1] = ex2_fast(fmaf(s[m][n][1], scale_log2, nm0));
      s[m][n][2] = ex2_fast(fmaf(s[m][n][2], scale_log2, nm1));
      s[m][n][3] = ex2_fast(fmaf(s[m][n][3], scale_log2, nm1));
      sum0 += s[m][n][0] + s[m][n][1];
      sum1 += s[m][n][2] + s[m][n][3];
    }
    const float inv0 = 1.0f / quad_sum(sum0), inv1 = 1.0f / quad_sum(sum1);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[m][n][0] *= inv0; s[m][n][1] *= inv0;
      s[m][n][2] *= inv1; s[m][n][3] *= inv1;
      d0 += s[m][n][0] * dp[m][n][0] + s[m][n][1] * dp[m][n][1];
      d1 += s[m][n][2] * dp[m][n][2] + s[m][n][3] * dp[m][n][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = 2 * kk + j;
        pa[m][kk][2 * j] = pack2(s[m][n][0], s[m][n][1]);
        pa[m][kk][2 * j + 1] = pack2(s[m][n][2], s[m][n][3]);
        sa[m][kk][2 * j] = pack2(s[m][n][0] * (dp[m][n][0] - d0) * scale, s[m][n][1] * (dp[m][n][1] - d0) * scale);
        sa[m][kk][2 * j + 1] = pack2(s[m][n][2] * (dp[m][n][2] - d1) * scale, s[m][n][3] * (dp[m][n][3] - d1) * scale);
      }
  }

  // out[kt] (16 keys x DH) = sum over both query tiles kq of X[kq]^T (16 keys x 16 queries) * B[kq] (16 queries x DH);
  // X = P (dV, B = dO) or dS (dK, B = Q).  A fragment of X^T for key tile kt from the packed fragments of X:
  //   a0 = T(x[kq][kt][0]) a1 = T(x[kq][kt][2]) a2 = T(x[kq][kt][1]) a3 = T(x[kq][kt][3])
  auto xt_times = [&](const uint32_t (&x)[2][2][4], uint32_t bbase, int bpitch, uint32_t cbase) {
    float acc[2][ND][4];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[kt][nd][e] = 0.f;
#pragma unroll
    for (int kq = 0; kq < 2; ++kq) {
      uint32_t a[2][4];
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        a[kt][0] = movm_t(x[kq][kt][0]);
        a[kt][1] = movm_t(x[kq][kt][2]);
        a[kt][2] = movm_t(x[kq][kt][1]);
        a[kt][3] = movm_t(x[kq][kt][3]);
      }
#pragma unroll
      for (int n2 = 0; n2 < ND / 2; ++n2) {
        uint32_t bf[4];
        ldsm_x4_t(bbase + (uint32_t)((kq * 16 * bpitch + n2 * 16) * 2), bf);
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
          mma(acc[kt][2 * n2], a[kt], bf[0], bf[1]);
          mma(acc[kt][2 * n2 + 1], a[kt], bf[2], bf[3]);
        }
      }
    }
    __syncwarp();   // all lanes are done reading the operand columns this result may overwrite
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) store_acc<ND, QP>(cbase, kt * 16, acc[kt], 1.f, 1.f);
  };

  // ---- B1: dV = P^T dO -> v columns (V is dead after phase A) ----
  xt_times(pa, sdo32 + ld.bt + ocol * 2, OP, sq32 + lq.c + vcol * 2);
  // ---- B2: dQ = dS K -> the dO columns of this head (dO is dead after B1); one K^T fragment load serves both tiles ----
  {
    float acc[2][ND][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nd = 0; nd < ND; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int n2 = 0; n2 < ND / 2; ++n2) {
        uint32_t bf[4];
        ldsm_x4_t(sq32 + lq.bt + (uint32_t)((kk * 16 * QP + kcol + n2 * 16) * 2), bf);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          mma(acc[m][2 * n2], sa[m][kk], bf[0], bf[1]);
          mma(acc[m][2 * n2 + 1], sa[m][kk], bf[2], bf[3]);
        }
      }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 2; ++m) store_acc<ND, OP>(sdo32 + ld.c + ocol * 2, m * 16, acc[m], 1.f, 1.f);
  }
  // ---- B3: dK = dS^T Q -> k columns (K is dead after B2) ----
  xt_times(sa, sq32 + lq.bt + qcol * 2, QP, sq32 + lq.c + kcol * 2);
  __syncthreads();
  bf16* out = dqkv + b * L * 3 * D + h0 * DH;
  store_tile<G::W, OP>(out, 3 * D, sdo, L, NT);                      // dQ
  store_tile<G::W, QP>(out + D, 3 * D, sq + G::W, L, NT);            // dK
  store_tile<G::W, QP>(out + 2 * D, 3 * D, sq + 2 * G::W, L, NT);    // dV
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("attention_mma: cudaFuncSetAttribute(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <int DH, int LP, int HG>
int launch_fwd(const void* qkv, void* o, int64_t B, int L, int heads, float scale, const float* out_scale, cudaStream_t st) {
  using G = Geo<DH, LP, HG>;
  int rc = set_smem(attn_fwd_mma_kernel<DH, LP, HG>, G::fwd_bytes);
  if (rc) return rc;
  attn_fwd_mma_kernel<DH, LP, HG><<<(unsigned)(B * (heads / HG)), G::kThreads, G::fwd_bytes, st>>>((const bf16*)qkv, (bf16*)o, L, heads, scale,
                                                                                                    out_scale);
  return check_launch("attention_fwd_mma");
}
template <int DH, int LP, int HG>
int launch_bwd(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, float scale, cudaStream_t st) {
  using G = Geo<DH, LP, HG>;
  static const bool staged = getenv("AFB_ATTN_BWD_STAGED") != nullptr;   // debugging knob: the older staged kernel
  if (LP == 32 && DH == 32 && !staged) {   // register-resident P / dS, shared operand fragments
    constexpr size_t bytes = G::bwd_tile_bytes;
    int rc = set_smem(attn_bwd32_kernel<DH, HG>, bytes);
    if (rc) return rc;
    attn_bwd32_kernel<DH, HG><<<(unsigned)(B * (heads / HG)), HG * 32, bytes, st>>>((const bf16*)qkv, (const bf16*)dO, (bf16*)dqkv, L, heads, scale);
    return check_launch("attention_bwd32");
  }
  int rc = set_smem(attn_bwd_mma_kernel<DH, LP, HG>, G::bwd_bytes);
  if (rc) return rc;
  attn_bwd_mma_kernel<DH, LP, HG><<<(unsigned)(B * (heads / HG)), G::kThreads * bwd_wph(LP), G::bwd_bytes, st>>>((const bf16*)qkv, (const bf16*)dO, (bf16*)dqkv, L, heads,
                                                                                                    scale);
  return check_launch("attention_bwd_mma");
}

}  // namespace attn_mma

// true if the tensor-core path covers this shape
bool attention_mma_supported(int L, int heads, int dh) {
  return (dh == 32 || dh == 64) && L >= 1 && L <= 64 && heads % 4 == 0;
}

#define AFB_ATTN_DISPATCH(FN, ...)                                                              \
  do {                                                                                          \
    const int LP = L <= 16 ? 16 : (L <= 32 ? 32 : (L <= 48 ? 48 : 64));                         \
    const bool hg8 = heads % 8 == 0 && (bwd ? (dh == 32 && LP <= 48) : (dh == 32 || LP <= 32)); \
    if (dh == 32) {                                                                             \
      if (hg8) {                                                                                \
        if (LP == 16) return attn_mma::FN<32, 16, 8>(__VA_ARGS__);                              \
        if (LP == 32) return attn_mma::FN<32, 32, 8>(__VA_ARGS__);                              \
        if (LP == 48) return attn_mma::FN<32, 48, 8>(__VA_ARGS__);                              \
        return attn_mma::FN<32, 64, 8>(__VA_ARGS__);                                            \
      }                                                                                         \
      if (LP == 16) return attn_mma::FN<32, 16, 4>(__VA_ARGS__);                                \
      if (LP == 32) return attn_mma::FN<32, 32, 4>(__VA_ARGS__);                                \
      if (LP == 48) return attn_mma::FN<32, 48, 4>(__VA_ARGS__);                                \
      return attn_mma::FN<32, 64, 4>(__VA_ARGS__);                                              \
    }                                                                                           \
    if (LP == 16) return attn_mma::FN<64, 16, 4>(__VA_ARGS__);                                  \
    if (LP == 32) return attn_mma::FN<64, 32, 4>(__VA_ARGS__);                                  \
    if (LP == 48) return attn_mma::FN<64, 48, 4>(__VA_ARGS__);                                  \
    return attn_mma::FN<64, 64, 4>(__VA_ARGS__);                                                \
  } while (0)

int attention_fwd_mma(const void* qkv, void* o, int64_t B, int L, int heads, int dh, float scale, const float* out_scale, cudaStream_t st) {
  const bool bwd = false;
  AFB_ATTN_DISPATCH(launch_fwd, qkv, o, B, L, heads, scale, out_scale, st);
}
int attention_bwd_mma(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, int dh, float scale, cudaStream_t st) {
  const bool bwd = true;
  AFB_ATTN_DISPATCH(launch_bwd, qkv, dO, dqkv, B, L, heads, scale, st);
}

}  // namespace afb
