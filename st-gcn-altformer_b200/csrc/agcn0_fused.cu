// gcn0 forward as ONE persistent cooperative kernel (bf16 performance mode; reference: model/unit_agcn.py:73-93 with
// C_in = 3, constructed at model/AltFormer/ST_GCN_AltFormer.py:43-48).  Replaces gcn0_scores + gcn0_apply_mma of agcn0.cu
// for the shapes it supports (V even, V <= 24, Cout = 128, the per-sample stage fits shared memory); agcn0.cu stays the
// fp32 parity path and the fallback.
//
//   phase 1  (one CTA per sample, 8 warps)         everything per-sample sits on mma.sync with hi/lo bf16 split operands
//     x[n] -> X2[(t,a)][u] and Y_i[(t,a)][v] = C_i x + e_i   (C_i = Wa_i^T Wb_i / (IC T), e_i = Wa_i^T bb_i / (IC T))
//     S_i^T[v][u] = sum_(t,a) Y_i[(t,a)][v] X2[(t,a)][u]     (rank-3T form of theta^T phi: no 1024-long contraction)
//     M_i = softmax_u(S_i) + A_i + PA_i  -> Mmat (fp32, backward), Mop (bf16 hi|lo, [(i,v)][u]) and shared memory
//     z = M^T x per 4-frame group -> r = (z, x, 1) as bf16 hi/lo, stored slot-major RT[slot][position] in shared memory
//     -> second moments R^T R by MMA (training) -> fp64 slot atomics
//   grid barrier (training only; sense-reversing, bounded spin)
//   phase 2  (every CTA, redundantly)              E[r], Cov(r) -> batch statistics of BOTH BatchNorms -> BN-folded weights
//                                                  as hi/lo mma B fragments in shared memory; CTA 0 also writes the
//                                                  statistics the backward needs and updates the running buffers
//   expansion                                      y = relu(Wfold [r; 1]) per 16-position tile, three-term hi/lo product so
//                                                  the ReLU mask is decided at ~2^-16, not 2^-9, relative precision
//                                                  -> 128B-swizzled staging -> TMA tensor stores (64 channels x 16 rows)
// Modes: training with N <= #CTAs ("resident": the CTA that owns a sample keeps its r rows in shared memory across the
// barrier and expands them itself); eval (no batch dependency: samples stream through phase 1 + expansion, no barrier);
// training with more samples than CTAs ("generic": M goes to global memory as bf16 hi|lo and warp pairs recompute z per
// 8-frame unit after the barrier).
// The 128-channel activation is written exactly once.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "mma_utils.cuh"

namespace afb {

int make_tensor_map_bf16(void* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                         uint32_t box0, uint32_t box1);

namespace {

using namespace mmau;

constexpr int NR = AFB_GCN0_NR;
constexpr int NMOM = AFB_GCN0_NMOM;
constexpr int NSTAT = AFB_GCN0_NSTAT_BASE;
constexpr int kSlots = AFB_GCN0_SLOTS;    // stride of the workspace halves
constexpr int kUseSlots = 8;              // slots the fused kernel spreads its fp64 atomics over
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTeams = kWarps / 2;      // phase 3: one unit (8 frames of one sample) per warp pair
constexpr int VP = 32;                  // joints padded to two k-steps
constexpr int UP = VP + 8;              // pitch (elements) of arrays whose contiguous index is the joint u: 80 B rows
constexpr int COUT = 128;
constexpr int kUnitFrames = 8;          // phase-3 unit; 8 V positions = whole 16-row tiles for even V
constexpr int kZFrames = 4;             // phase-1 z group: 12 (t, a) rows of one 16-row tile
constexpr unsigned long long kSpinTimeoutNs = 2000000000ull;

__host__ __device__ constexpr int r16u(int n) { return (n + 15) & ~15; }
__host__ __device__ constexpr int tri(int j, int k) { return NR + j * NR - (j * (j - 1)) / 2 + (k - j); }  // j <= k

// CBT: 8-column tiles per adjacency subset (V <= 8 CBT).  Columns c = i * CB + v.
template <int CBT> struct Cfg {
  static constexpr int CB = 8 * CBT, NC = 3 * CB, NT = 3 * CBT;
  static constexpr int MT = (NC + 15) / 16;     // 16-row tiles over c in the scores step
  static constexpr int YP = MT * 16 + 8;        // pitch of Y[(t,a)][c]: odd multiple of 16 B -> conflict-free ldmatrix
};

struct SmemPlan {   // byte offsets inside the dynamic shared memory (all multiples of 16; `work` is 1024-aligned)
  int wfrag_h, wfrag_l, ap, work;   // ap: fp32 [3][V][V] = A + PA, staged once per CTA (every CTA reading the same 12 KB
                                    // from L2 with scattered 4-byte loads serialises on a handful of L2 lines)
  int x2h, x2l, msh, msl, yh, yl, rth, rtl, pp, p1_end;   // phase 1 (the r rows overlay the Y operand); pp = RT pitch (elements)
  int stage;                                              // 8 x 4 KB TMA staging (overlays X2 / Ms in the resident modes)
  int team, team_bytes, aop_l, xt_h, xt_l, pup;           // generic mode; pup = pitch of the team's RT (elements)
  int fin;                                                // phase 2 scratch
  int xraw, xraw_bytes;                                   // the sample's (T, V, 3) fp32 block, landed by one bulk copy
  int total;
};
// resident: the r rows of the CTA's sample stay in shared memory until they are expanded (one sample per CTA, or eval)
template <int CBT>
__host__ __device__ inline SmemPlan make_plan(int T, int V, bool resident) {
  using C = Cfg<CBT>;
  SmemPlan s;
  s.wfrag_h = 0;
  s.wfrag_l = 4096;
  s.ap = 8192;
  s.work = 8192 + 7168;             // 3 * 24 * 24 floats rounded up; multiple of 1024
  const int K16 = r16u(3 * T), XR = K16 + 16;
  s.x2h = s.work;
  s.x2l = s.x2h + XR * UP * 2;
  s.msh = s.x2l + XR * UP * 2;
  s.msl = s.msh + C::MT * 16 * UP * 2;
  int lo = s.msl + C::MT * 16 * UP * 2;
  if (lo < s.work + kWarps * 4096) lo = s.work + kWarps * 4096;   // the staging buffers overlay X2 / Ms only
  lo = (lo + 1023) & ~1023;
  s.yh = lo;
  s.yl = s.yh + K16 * C::YP * 2;
  s.pp = r16u(T * V) + 8;           // (pp * 2) bytes = odd multiple of 16: 8-row ldmatrix phases hit 8 distinct bank groups
  s.rth = lo;
  s.rtl = s.rth + 16 * s.pp * 2;
  const int ybytes = 2 * K16 * C::YP * 2, rbytes = 2 * 16 * s.pp * 2;
  s.p1_end = lo + (ybytes > rbytes ? ybytes : rbytes);
  s.stage = s.work;
  s.team = s.stage + kWarps * 4096;
  s.pup = r16u(kUnitFrames * V) + 8;
  s.aop_l = 16 * s.pup * 2;
  s.xt_h = 2 * s.aop_l;
  s.xt_l = s.xt_h + 32 * UP * 2;
  s.team_bytes = s.xt_l + 32 * UP * 2;
  const int p3_end = s.team + kTeams * s.team_bytes;
  s.fin = s.work;
  s.total = (resident || s.p1_end > p3_end) ? s.p1_end : p3_end;
  s.xraw = (s.total + 127) & ~127;
  s.xraw_bytes = T * V * 12;
  s.total = s.xraw + ((s.xraw_bytes + 15) & ~15);
  return s;
}

struct FusedArgs {
  uint32_t* mop;          // bf16 pairs: [N][2 (hi, lo)][NC][UP]  (generic mode only)
  double* moments;        // [2][kSlots][NMOM]
  unsigned int* ctrl;     // [0] barrier arrivals (monotonic), [1] its value at the start of the current launch,
                          // [2] parity selecting the moment buffer of the current launch
  int resident;           // training: N <= gridDim.x, every CTA keeps its sample's r rows in shared memory
};

// phase time stamps of the last launch (globaltimer ns, thread 0 of every CTA): debugging / profiling aid
constexpr int kStamps = 16, kStampCtas = 512;
__device__ unsigned long long g_stamps[kStampCtas * kStamps];

// ---- small helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf_lo_f(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi_f(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// (a, b) -> bf16x2 hi parts and bf16x2 of the residuals (a in the low half)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack2(a, b);
  lo = pack2(a - bf_lo_f(hi), b - bf_hi_f(hi));
}
__device__ __forceinline__ void split1(float a, uint16_t& hi, uint16_t& lo) {
  const bf16 h = __float2bfloat16_rn(a);
  const bf16 l = __float2bfloat16_rn(a - __bfloat162float(h));
  hi = *reinterpret_cast<const uint16_t*>(&h);
  lo = *reinterpret_cast<const uint16_t*>(&l);
}
__device__ __forceinline__ uint32_t relu_pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
// coherent (never .nc) loads of data another CTA wrote earlier in this launch
__device__ __forceinline__ uint32_t ld_global_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// (compiled in with -DAFB_GCN0_STAMPS only: a %globaltimer read costs ~100 ns of the issuing warp's time and the other
// warps meet it at the next barrier -- 13 stamps were 7 % of the kernel's stall samples)
__device__ __forceinline__ void stamp(int k) {
#ifdef AFB_GCN0_STAMPS
  if (threadIdx.x == 0 && blockIdx.x < kStampCtas) g_stamps[blockIdx.x * kStamps + k] = global_ns();
#else
  (void)k;
#endif
}
__device__ __forceinline__ void team_sync(int team) {   // the two warps of a generic-mode team
  asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && (++spins > (1u << 24))) __trap();   // bounded: a lost copy traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Grid barrier on a monotonically increasing arrival counter: ctrl[0] counts arrivals over all launches, ctrl[1] holds
// the counter value at the start of this launch (`base`, read at kernel entry before any CTA can have arrived; the last
// arriver publishes base + gridDim.x for the next launch, off the critical path).  Waiters poll the counter itself, so
// they wake one L2 round trip after the last arrival.  Needs all CTAs co-resident (cooperative launch); the spin is
// bounded so a bug traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* ctrl, unsigned int base) {
  __syncthreads();   // CTA-scope ordering of every thread's prior writes before thread 0's release (cumulativity)
  if (threadIdx.x == 0) {
    const unsigned int target = base + gridDim.x;
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctrl) : "memory");
    if (old + 1 == target) {
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(ctrl + 1), "r"(target) : "memory");
      ctrl[2] ^= 1u;   // moment-buffer parity of the next launch
    } else {
      const unsigned long long t0 = global_ns();
      unsigned int spins = 0;
      while ((int)(ld_acquire_u32(ctrl) - target) < 0) {
        __nanosleep(20);
        if ((++spins & 255u) == 0 && global_ns() - t0 > kSpinTimeoutNs) __trap();
      }
    }
  }
  __syncthreads();
}

struct Lane {   // per-thread indices used everywhere
  int tid, lane, warp, g, tq, lj, lr;
};

__device__ __forceinline__ void zero_region(uint8_t* base, int bytes, int tid) {
  uint4* z4 = reinterpret_cast<uint4*>(base);
  for (int i = tid; i < (bytes >> 4); i += kThreads) z4[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------------------------
// phase 1 for one sample: operands, scores, M, r rows (hi/lo, slot-major RT[16][pp]) in shared memory; training: R^T R
// accumulated into this warp's D0 / D1.  The phase-1 region must be zero on entry (zero_region).  Ends with every r row
// written and visible (__syncthreads inside).
// ------------------------------------------------------------------------------------------------------------------
template <int CBT>
__device__ __forceinline__ void phase1_sample(const afb_gcn0_fwd_t& p, const FusedArgs& fa, const SmemPlan& pl, uint8_t* sm,
                                              const float (*coef_s)[12], int n, const Lane& L, bool write_mop, bool moments,
                                              float (&D0)[4], float (&D1)[4], const float* xs) {
  using C = Cfg<CBT>;
  constexpr int CB = C::CB, NC = C::NC, NT = C::NT, MT = C::MT, YP = C::YP;
  const int T = p.T, V = p.V;
  const int tid = L.tid, lane = L.lane, warp = L.warp, g = L.g, tq = L.tq, lj = L.lj, lr = L.lr;
  const int K16 = r16u(3 * T), PP = pl.pp;
  bf16* X2h = reinterpret_cast<bf16*>(sm + pl.x2h);
  bf16* X2l = reinterpret_cast<bf16*>(sm + pl.x2l);
  bf16* Yh = reinterpret_cast<bf16*>(sm + pl.yh);
  bf16* Yl = reinterpret_cast<bf16*>(sm + pl.yl);
  bf16* Msh = reinterpret_cast<bf16*>(sm + pl.msh);
  bf16* Msl = reinterpret_cast<bf16*>(sm + pl.msl);
  bf16* RTh = reinterpret_cast<bf16*>(sm + pl.rth);
  bf16* RTl = reinterpret_cast<bf16*>(sm + pl.rtl);
  const float* APs = reinterpret_cast<const float*>(sm + pl.ap);
  // S1: x[n] -> X2 (hi/lo) and Y (hi/lo); thread = (frame, pair of adjacent joints).  Two pairs per thread per round, both
  // pairs' loads issued before either is used.
  {
    const float* xg = xs != nullptr ? xs : p.x + (size_t)n * T * V * 3;   // xs: the sample already sits in shared memory
    const int VH = V >> 1, npairs = T * VH;
    for (int pr0 = tid; pr0 < npairs; pr0 += 2 * kThreads) {
      float2 q[2][3];
      int tt[2], vv[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pr = pr0 + h * kThreads;
        const bool ok = pr < npairs;
        tt[h] = ok ? pr / VH : -1;
        vv[h] = ok ? (pr - tt[h] * VH) * 2 : 0;
        const float2* xp = reinterpret_cast<const float2*>(xg + (size_t)((ok ? tt[h] : 0) * V + vv[h]) * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) q[h][k] = ok ? xp[k] : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (tt[h] < 0) continue;
        const int t = tt[h], v = vv[h];
        const float xa[3] = {q[h][0].x, q[h][0].y, q[h][1].x}, xb[3] = {q[h][1].y, q[h][2].x, q[h][2].y};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          uint32_t hi, lo;
          split2(xa[a], xb[a], hi, lo);
          *reinterpret_cast<uint32_t*>(X2h + (3 * t + a) * UP + v) = hi;
          *reinterpret_cast<uint32_t*>(X2l + (3 * t + a) * UP + v) = lo;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const float* cf = coef_s[i];
            const float ya = fmaf(cf[a * 3], xa[0], fmaf(cf[a * 3 + 1], xa[1], fmaf(cf[a * 3 + 2], xa[2], cf[9 + a])));
            const float yb = fmaf(cf[a * 3], xb[0], fmaf(cf[a * 3 + 1], xb[1], fmaf(cf[a * 3 + 2], xb[2], cf[9 + a])));
            uint32_t hi, lo;
            split2(ya, yb, hi, lo);
            *reinterpret_cast<uint32_t*>(Yh + (3 * t + a) * YP + i * CB + v) = hi;
            *reinterpret_cast<uint32_t*>(Yl + (3 * t + a) * YP + i * CB + v) = lo;
          }
      }
    }
  }
  __syncthreads();
  stamp(3);
  // S2: S^T[c][u] = sum_k Y[k][c] X2[k][u]; softmax over u; M = P + A + PA
  for (int mt = warp; mt < MT; mt += kWarps) {
    const int c0 = mt * 16 + g, c1 = c0 + 8;
    const int i0 = c0 / CB, v0 = c0 - i0 * CB, i1 = c1 / CB, v1 = c1 - i1 * CB;
    const bool ok0 = c0 < NC && v0 < V, ok1 = c1 < NC && v1 < V;
    float ap[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int u = nt * 8 + 2 * tq + j;
        const int e0 = (i0 * V + u) * V + v0, e1 = (i1 * V + u) * V + v1;
        ap[nt][j] = (ok0 && u < V) ? APs[e0] : 0.f;
        ap[nt][2 + j] = (ok1 && u < V) ? APs[e1] : 0.f;
      }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll 2
    for (int ks = 0; ks < K16 / 16; ++ks) {
      uint32_t ah[4], al[4];
      const int a_off = (ks * 16 + (lj >> 1) * 8 + lr) * YP + mt * 16 + (lj & 1) * 8;
      ldsm_x4_t(smem_u32(Yh + a_off), ah);
      ldsm_x4_t(smem_u32(Yl + a_off), al);
#pragma unroll
      for (int ntp = 0; ntp < 2; ++ntp) {
        uint32_t bh[4], bl[4];
        const int b_off = (ks * 16 + (lj & 1) * 8 + lr) * UP + ntp * 16 + (lj >> 1) * 8;
        ldsm_x4_t(smem_u32(X2h + b_off), bh);
        ldsm_x4_t(smem_u32(X2l + b_off), bl);
        mma(acc[2 * ntp], ah, bh[0], bh[1]);
        mma(acc[2 * ntp + 1], ah, bh[2], bh[3]);
        mma(acc[2 * ntp], ah, bl[0], bl[1]);
        mma(acc[2 * ntp + 1], ah, bl[2], bl[3]);
        mma(acc[2 * ntp], al, bh[0], bh[1]);
        mma(acc[2 * ntp + 1], al, bh[2], bh[3]);
      }
    }
    // softmax over u: this thread holds u = nt*8 + 2tq + {0,1} of rows c0 (elements 0,1) and c1 (elements 2,3)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool uv = nt * 8 + 2 * tq + j < V;
        if (!uv) { acc[nt][j] = -INFINITY; acc[nt][2 + j] = -INFINITY; }
        mx0 = fmaxf(mx0, acc[nt][j]);
        mx1 = fmaxf(mx1, acc[nt][2 + j]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        acc[nt][j] = __expf(acc[nt][j] - mx0);        // exp(-inf) = 0 for the padded joints
        acc[nt][2 + j] = __expf(acc[nt][2 + j] - mx1);
        s0 += acc[nt][j];
        s1 += acc[nt][2 + j];
      }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float r0 = 1.0f / s0, r1 = 1.0f / s1;
    float* Mg = p.Mmat + (size_t)n * 3 * V * V;
    uint32_t* moph = fa.mop + ((size_t)n * 2 * NC) * (UP / 2);
    uint32_t* mopl = moph + (size_t)NC * (UP / 2);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int u = nt * 8 + 2 * tq;
      float m00 = 0.f, m01 = 0.f, m10 = 0.f, m11 = 0.f;
      if (ok0) {
        if (u < V) { m00 = acc[nt][0] * r0 + ap[nt][0]; Mg[(i0 * V + u) * V + v0] = m00; }
        if (u + 1 < V) { m01 = acc[nt][1] * r0 + ap[nt][1]; Mg[(i0 * V + u + 1) * V + v0] = m01; }
      }
      if (ok1) {
        if (u < V) { m10 = acc[nt][2] * r1 + ap[nt][2]; Mg[(i1 * V + u) * V + v1] = m10; }
        if (u + 1 < V) { m11 = acc[nt][3] * r1 + ap[nt][3]; Mg[(i1 * V + u + 1) * V + v1] = m11; }
      }
      uint32_t h0, l0, h1, l1;
      split2(m00, m01, h0, l0);
      split2(m10, m11, h1, l1);
      *reinterpret_cast<uint32_t*>(Msh + c0 * UP + u) = h0;
      *reinterpret_cast<uint32_t*>(Msl + c0 * UP + u) = l0;
      *reinterpret_cast<uint32_t*>(Msh + c1 * UP + u) = h1;
      *reinterpret_cast<uint32_t*>(Msl + c1 * UP + u) = l1;
      if (write_mop) {
        if (c0 < NC) { moph[(c0 * UP + u) >> 1] = h0; mopl[(c0 * UP + u) >> 1] = l0; }
        if (c1 < NC) { moph[(c1 * UP + u) >> 1] = h1; mopl[(c1 * UP + u) >> 1] = l1; }
      }
    }
  }
  __syncthreads();   // M complete; the Y operand is dead from here on (the r rows overlay it)
  stamp(4);
  // S4: z per 4-frame group -> RT[slot][pos] (hi/lo).  Every element of RT is written exactly once: slots 0..8 by the z
  // scatter (column pairs -> 4-byte stores), 9..11 by the x copy, 12 (ones) .. 15 and the padded positions by the fills.
  {
    const int ngroups = (T + kZFrames - 1) / kZFrames;
    for (int zg = warp; zg < ngroups; zg += kWarps) {
      const int t0 = zg * kZFrames, nfr = min(kZFrames, T - t0);
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int a_off = (3 * t0 + (lj & 1) * 8 + lr) * UP + ks * 16 + (lj >> 1) * 8;
        ldsm_x4(smem_u32(X2h + a_off), ah[ks]);
        ldsm_x4(smem_u32(X2l + a_off), al[ks]);
      }
      // rows of this thread's C fragments: m = g and g + 8 -> (frame, channel)
      const int m0 = g, m1 = g + 8;
      const int tl0 = m0 / 3, a0 = m0 - 3 * tl0, tl1 = m1 / 3, a1 = m1 - 3 * tl1;
      const bool rv0 = tl0 < nfr, rv1 = m1 < 3 * kZFrames && tl1 < nfr;
      // element offsets of (slot = a, position = frame start + 2 tq); the subset adds 3 i PP, the column tile 8 per tile
      const int d0 = a0 * PP + (t0 + tl0) * V + 2 * tq, d1 = a1 * PP + (t0 + tl1) * V + 2 * tq;
#pragma unroll
      for (int nt0 = 0; nt0 < NT; nt0 += 3) {   // three column tiles in flight: independent MMA chains
        float z[3][4];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          uint32_t bh[4], bl[4];
          const int b_off = ((nt0 + k) * 8 + lr) * UP + lj * 8;
          ldsm_x4(smem_u32(Msh + b_off), bh);
          ldsm_x4(smem_u32(Msl + b_off), bl);
          z[k][0] = z[k][1] = z[k][2] = z[k][3] = 0.f;
          mma(z[k], ah[0], bh[0], bh[1]);
          mma(z[k], ah[1], bh[2], bh[3]);
          mma(z[k], ah[0], bl[0], bl[1]);
          mma(z[k], ah[1], bl[2], bl[3]);
          mma(z[k], al[0], bh[0], bh[1]);
          mma(z[k], al[1], bh[2], bh[3]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int nt = nt0 + k;
          const int i = nt / CBT, vt = (nt - i * CBT) * 8;
          if (vt + 2 * tq < V) {   // V is even: both columns of the pair are valid together
            uint32_t hi, lo;
            if (rv0) {
              split2(z[k][0], z[k][1], hi, lo);
              *reinterpret_cast<uint32_t*>(RTh + 3 * i * PP + vt + d0) = hi;
              *reinterpret_cast<uint32_t*>(RTl + 3 * i * PP + vt + d0) = lo;
            }
            if (rv1) {
              split2(z[k][2], z[k][3], hi, lo);
              *reinterpret_cast<uint32_t*>(RTh + 3 * i * PP + vt + d1) = hi;
              *reinterpret_cast<uint32_t*>(RTl + 3 * i * PP + vt + d1) = lo;
            }
          }
        }
      }
      // slots 9..11 = x (already split in X2); slot 12 = 1 (hi) / 0 (lo); 13..15 = 0
      const int VH = V >> 1;
      for (int e = lane; e < 3 * nfr * VH; e += 32) {
        const int row = e / VH, vp = (e - row * VH) * 2, tl = row / 3, a = row - 3 * tl;
        const int src = (3 * (t0 + tl) + a) * UP + vp, dst = (9 + a) * PP + (t0 + tl) * V + vp;
        *reinterpret_cast<uint32_t*>(RTh + dst) = *reinterpret_cast<const uint32_t*>(X2h + src);
        *reinterpret_cast<uint32_t*>(RTl + dst) = *reinterpret_cast<const uint32_t*>(X2l + src);
      }
      for (int e = lane; e < nfr * VH; e += 32) {
        const int dst = 12 * PP + t0 * V + 2 * e;
        *reinterpret_cast<uint32_t*>(RTh + dst) = 0x3f803f80u;
        *reinterpret_cast<uint32_t*>(RTl + dst) = 0u;
#pragma unroll
        for (int s = 1; s < 4; ++s) {
          *reinterpret_cast<uint32_t*>(RTh + dst + s * PP) = 0u;
          *reinterpret_cast<uint32_t*>(RTl + dst + s * PP) = 0u;
        }
      }
    }
    // padded positions (T*V .. PP) of all 16 slot rows
    const int padw = (PP - T * V) >> 1;
    for (int e = tid; e < 16 * padw; e += kThreads) {
      const int row = e / padw, c = (e - row * padw) * 2;
      *reinterpret_cast<uint32_t*>(RTh + row * PP + T * V + c) = 0u;
      *reinterpret_cast<uint32_t*>(RTl + row * PP + T * V + c) = 0u;
    }
  }
  __syncthreads();
  stamp(5);
  if (moments) {
    for (int ks = warp; ks < r16u(T * V) / 16; ks += kWarps) {
      uint32_t rh[4];
      ldsm_x4(smem_u32(RTh + ((lj & 1) * 8 + lr) * PP + ks * 16 + (lj >> 1) * 8), rh);
      // the statistics average ~1e5 products: unbiased bf16 rounding of the factors cancels to ~1e-5 relative, so the
      // hi x hi term alone carries them (the lo parts exist for the ReLU-mask precision of the expansion).  The same
      // fragments serve as A (slots x positions) and as both column tiles of B.
      mma(D0, rh, rh[0], rh[2]);
      mma(D1, rh, rh[1], rh[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// expansion of one 16-position tile from slot-major r rows: y = relu(Wfold [r; 1]), three-term hi/lo product; the rows
// leave through the warp's two 2 KB staging halves as TMA tensor stores (64 channels x 16 rows each)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void expand_tile(const bf16* RH, const bf16* RL, int pitch, int col0, const uint2* wfh, const uint2* wfl,
                                            uint32_t stage_u32, const CUtensorMap* tmY, int out_row, int n, const Lane& L) {
  uint32_t ah[4], al[4];
  {
    const int off = ((L.lj >> 1) * 8 + L.lr) * pitch + col0 + (L.lj & 1) * 8;
    ldsm_x4_t(smem_u32(RH + off), ah);
    ldsm_x4_t(smem_u32(RL + off), al);
  }
  const int srow = (L.lj & 1) * 8 + L.lr;   // staging row this lane addresses in stmatrix
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    // the store issued from this half-buffer one tile ago must have finished reading it
    if (L.lane == 0) bulk_wait_read1();
    __syncwarp();
    const uint32_t sbase = stage_u32 + half * 2048 + srow * 128;
#pragma unroll
    for (int q = 0; q < 2; ++q) {   // four n-tiles at a time: four independent 3-MMA chains in flight
      const int nt = half * 8 + 4 * q;
      float d[4][4];
      uint2 fh[4], fl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        fh[k] = wfh[(nt + k) * 32 + L.lane];
        fl[k] = wfl[(nt + k) * 32 + L.lane];
        d[k][0] = d[k][1] = d[k][2] = d[k][3] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) mma(d[k], ah, fh[k].x, fh[k].y);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma(d[k], al, fh[k].x, fh[k].y);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma(d[k], ah, fl[k].x, fl[k].y);
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int chunk = 4 * q + 2 * k2 + (L.lj >> 1);
        stsm_x4(sbase + ((chunk ^ (srow & 7)) << 4), relu_pack2(d[2 * k2][0], d[2 * k2][1]), relu_pack2(d[2 * k2][2], d[2 * k2][3]),
                relu_pack2(d[2 * k2 + 1][0], d[2 * k2 + 1][1]), relu_pack2(d[2 * k2 + 1][2], d[2 * k2 + 1][3]));
      }
    }
    fence_async_smem();
    __syncwarp();
    if (L.lane == 0) {
      tma_store_3d(tmY, stage_u32 + half * 2048, half * 64, out_row, n);
      bulk_commit();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <int CBT>
__global__ void __launch_bounds__(kThreads, 2) gcn0_fused_kernel(const __grid_constant__ afb_gcn0_fwd_t p,
                                                                 const __grid_constant__ CUtensorMap tmY, const FusedArgs fa) {
  using C = Cfg<CBT>;
  constexpr int NC = C::NC, NT = C::NT;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ float coef_s[3][12];
  __shared__ float momS[16 * 16];
  __shared__ unsigned int s_gen, s_par;

  const int T = p.T, V = p.V, N = p.N;
  Lane L;
  L.tid = threadIdx.x; L.lane = L.tid & 31; L.warp = L.tid >> 5; L.g = L.lane >> 2; L.tq = L.lane & 3;
  L.lj = L.lane >> 3; L.lr = L.lane & 7;
  const int tid = L.tid, lane = L.lane, warp = L.warp, g = L.g, tq = L.tq, lj = L.lj, lr = L.lr;
  const bool training = p.training != 0;
  const bool resident = !training || fa.resident != 0;
  const SmemPlan pl = make_plan<CBT>(T, V, resident);
  stamp(0);

  // ---- prologue: every global read of the first sample is requested up front (one exposed miss latency) -----------
  // (the two control words were published by the previous launch; two different warps read them so neither load waits for
  // the other, and thread 0 is left free to issue the sample's bulk copy first)
  if (tid == 32) s_gen = training ? ld_acquire_u32(fa.ctrl + 1) : 0u;
  if (tid == 64) s_par = training ? ld_acquire_u32(fa.ctrl + 2) : 0u;
  // x[n] (T*V*12 contiguous bytes) -> shared memory by ONE bulk copy issued before anything else: S1 then reads it at
  // shared-memory latency instead of paying an L2 round trip per round.  (16-byte size / alignment: else direct loads.)
  __shared__ __align__(8) unsigned long long xbar;
  const uint32_t xbar_u32 = smem_u32(&xbar);
  const bool use_bulk = ((T * V * 12) & 15) == 0 && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  float* xraw = reinterpret_cast<float*>(sm + pl.xraw);
  uint32_t xphase = 0;
  if (use_bulk) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xbar_u32));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      if (blockIdx.x < N) bulk_load(smem_u32(xraw), p.x + (size_t)blockIdx.x * T * V * 3, (uint32_t)(T * V * 12), xbar_u32);
    }
  } else if (blockIdx.x < N) {   // x[n] -> L2 (consumed two barriers from here)
    const char* xg = reinterpret_cast<const char*>(p.x + (size_t)blockIdx.x * T * V * 3);
    for (int i = tid; i < (T * V * 12 + 127) >> 7; i += kThreads) prefetch_l2(xg + i * 128);
  }
  // per-sample: wait for the sample's copy (issuing it first for every sample after the CTA's first)
  auto acquire_x = [&](int n, bool first) -> const float* {
    if (!use_bulk) return nullptr;
    if (!first) {
      // (the previous sample's S1 reads of xraw are behind at least one __syncthreads)
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulk_load(smem_u32(xraw), p.x + (size_t)n * T * V * 3, (uint32_t)(T * V * 12), xbar_u32);
      }
    }
    mbar_wait_parity(xbar_u32, xphase);
    xphase ^= 1u;
    return xraw;
  };
  // A + PA and the score weights -> shared memory.  All loads are issued before the first use (fixed trip counts, fully
  // unrolled): a load -> store loop would pay one cold-miss latency per iteration.
  const int IC = p.IC, per = 7 * IC;
  float* wst = reinterpret_cast<float*>(sm);   // [3][7 * IC]: Wa (3 IC) | Wb (3 IC) | bb (IC); overlays the (unwritten) W fragments
  {
    constexpr int kApIt = (3 * 24 * 24 + kThreads - 1) / kThreads;   // V <= 24
    constexpr int kWsIt = 3;                                         // 3 * 7 * IC <= 768 (IC <= 36; checked on the host)
    float va[kApIt], vb[kApIt], vw[kWsIt];
#pragma unroll
    for (int k = 0; k < kApIt; ++k) {
      const int e = tid + k * kThreads;
      va[k] = e < 3 * V * V ? p.A[e] : 0.f;
      vb[k] = e < 3 * V * V ? p.PA[e] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kWsIt; ++k) {
      const int e = tid + k * kThreads;
      float v = 0.f;
      if (e < 3 * per) {
        const int i = e / per, r = e - i * per;
        v = r < 3 * IC ? p.Wa[i][r] : (r < 6 * IC ? p.Wb[i][r - 3 * IC] : p.bb[i][r - 6 * IC]);
      }
      vw[k] = v;
    }
    float* APw = reinterpret_cast<float*>(sm + pl.ap);
#pragma unroll
    for (int k = 0; k < kApIt; ++k) {
      const int e = tid + k * kThreads;
      if (e < 3 * V * V) APw[e] = va[k] + vb[k];
    }
#pragma unroll
    for (int k = 0; k < kWsIt; ++k) {
      const int e = tid + k * kThreads;
      if (e < 3 * per) wst[e] = vw[k];
    }
  }
  // weights of this thread's output channel (phase 2): requested here so their latency hides behind phase 1
  const int o = tid & (COUT - 1);
  float w[NR], bsum = 0.f, bd1 = 0.f, bd2 = 0.f, bdn = 0.f, gam_h = 0.f, bet_h = 0.f, gam_d = 0.f, bet_d = 0.f, rm_h = 0.f, rv_h = 1.f, rm_d = 0.f, rv_d = 1.f;
  if (tid < COUT) {
#pragma unroll
    for (int j = 0; j < 9; ++j) w[j] = p.Wd[j / 3][o * 3 + j % 3];
#pragma unroll
    for (int j = 9; j < 12; ++j) w[j] = p.Wdn[o * 3 + (j - 9)];
    bsum = p.bd[0][o]; bd1 = p.bd[1][o]; bd2 = p.bd[2][o];   // summed in phase 2 (nothing here may wait on these loads)
    bdn = p.bdn[o];
    gam_h = p.bn_g[o]; bet_h = p.bn_b[o]; gam_d = p.dn_g[o]; bet_d = p.dn_b[o];
    rm_h = p.bn_rm[o]; rv_h = p.bn_rv[o]; rm_d = p.dn_rm[o]; rv_d = p.dn_rv[o];
  } else {
#pragma unroll
    for (int j = 0; j < NR; ++j) w[j] = 0.f;
  }
  momS[tid] = 0.f;
  zero_region(sm + pl.work, pl.p1_end - pl.work, tid);   // padding of every phase-1 operand must be exact zeros
  __syncthreads();
  stamp(1);
  // scaled score coefficients: 4 lanes per coefficient over the inner channels
  {
    const bool act = tid < 36 * 4;     // (every lane takes part in the shuffles)
    const int q = act ? tid >> 2 : 0, part = tid & 3, i = q / 12, e = q - 12 * i;
    const float* Wa = wst + i * per;
    const float* Wb = Wa + 3 * IC;
    const float* bb = Wa + 6 * IC;
    const int ja = e < 9 ? e / 3 : e - 9, jb = e % 3;
    float acc = 0.f;
    for (int c = part; c < IC; c += 4) acc = fmaf(Wa[c * 3 + ja], e < 9 ? Wb[c * 3 + jb] : bb[c], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (act && part == 0) coef_s[i][e] = acc / (float)(IC * T);
  }
  __syncthreads();
  stamp(2);
  const unsigned int gen = s_gen;            // arrival counter at the start of this launch
  const int buf = (int)(s_par & 1u);         // moment buffer of this launch (flipped by the last arriver for the next one)
  double* mom_cur = fa.moments + (size_t)buf * kSlots * NMOM;
  if (training) {   // re-arm the buffer the NEXT launch will use (last read by the previous launch, which has completed)
    double* mom_next = fa.moments + (size_t)(buf ^ 1) * kSlots * NMOM;
    for (int s = blockIdx.x; s < kUseSlots; s += gridDim.x)
      if (tid < NMOM) mom_next[s * NMOM + tid] = 0.0;
  }

  // phase 2 as a lambda: E[r], Cov(r) (training: from the fp64 slot sums) -> BN statistics -> folded hi/lo B fragments
  auto fold_weights = [&]() {
    double* tot = reinterpret_cast<double*>(sm + pl.fin);      // [96]
    float* Ef = reinterpret_cast<float*>(tot + NMOM);          // [12] + pad
    float* Covf = Ef + 16;                                      // [12][12]
    const double m = (double)N * T * V;
    if (training) {
      const double inv_m = 1.0 / m;
      if (tid < NMOM) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < kUseSlots; ++l) s += __ldcg(mom_cur + l * NMOM + tid);
        tot[tid] = s * inv_m;
      }
      __syncthreads();
      stamp(9);
      if (tid < NR) Ef[tid] = (float)tot[tid];
      if (tid < NR * NR) {
        const int j = tid / NR, k = tid % NR;
        const int a = j < k ? j : k, b = j < k ? k : j;
        Covf[tid] = (float)(tot[tri(a, b)] - tot[j] * tot[k]);
      }
    } else {
      if (tid < NR) Ef[tid] = 0.f;
      if (tid < NR * NR) Covf[tid] = 0.f;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
      if (tid < NR) p.stats[tid] = Ef[tid];
      if (tid < NR * NR) p.stats[NR + tid] = Covf[tid];
    }
    if (tid < COUT) {
      bsum += bd1 + bd2;
      bd1 = bd2 = 0.f;
      float mean_h = bsum, mean_d = bdn, var_h = 0.f, var_d = 0.f;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        mean_h = fmaf(w[j], Ef[j], mean_h);
        float row = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) row = fmaf(Covf[j * NR + k], w[k], row);
        var_h = fmaf(w[j], row, var_h);
      }
#pragma unroll
      for (int j = 9; j < 12; ++j) {
        mean_d = fmaf(w[j], Ef[j], mean_d);
        float row = 0.f;
#pragma unroll
        for (int k = 9; k < 12; ++k) row = fmaf(Covf[j * NR + k], w[k], row);
        var_d = fmaf(w[j], row, var_d);
      }
      var_h = fmaxf(var_h, 0.f);
      var_d = fmaxf(var_d, 0.f);
      if (training) {
        if (blockIdx.x == 0) {
          const float unb = m > 1.0 ? (float)(m / (m - 1.0)) : 1.0f;
          p.bn_rm[o] = (1.0f - p.momentum) * rm_h + p.momentum * mean_h;
          p.bn_rv[o] = (1.0f - p.momentum) * rv_h + p.momentum * var_h * unb;
          p.dn_rm[o] = (1.0f - p.momentum) * rm_d + p.momentum * mean_d;
          p.dn_rv[o] = (1.0f - p.momentum) * rv_d + p.momentum * var_d * unb;
        }
      } else {
        mean_h = rm_h; var_h = rv_h; mean_d = rm_d; var_d = rv_d;
      }
      const float rstd_h = rsqrtf(var_h + p.eps), rstd_d = rsqrtf(var_d + p.eps);
      const float sh = gam_h * rstd_h, sd = gam_d * rstd_d;
      float wf[16];
#pragma unroll
      for (int j = 0; j < 9; ++j) wf[j] = sh * w[j];
#pragma unroll
      for (int j = 9; j < 12; ++j) wf[j] = sd * w[j];
      // y = sh (w.z + b - mean_h) + beta_h + sd (wdn.x + bdn - mean_d) + beta_d: r is NOT centred here (its hi/lo split
      // carries 16 mantissa bits), the constant rides on the "1" slot
      wf[12] = fmaf(sh, bsum - mean_h, bet_h) + fmaf(sd, bdn - mean_d, bet_d);
      wf[13] = wf[14] = wf[15] = 0.f;
      uint2* fh = reinterpret_cast<uint2*>(sm + pl.wfrag_h) + ((o >> 3) * 32 + (o & 7) * 4);
      uint2* fl = reinterpret_cast<uint2*>(sm + pl.wfrag_l) + ((o >> 3) * 32 + (o & 7) * 4);
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) {
        uint32_t h0, l0, h1, l1;
        split2(wf[2 * t4], wf[2 * t4 + 1], h0, l0);
        split2(wf[2 * t4 + 8], wf[2 * t4 + 9], h1, l1);
        fh[t4] = make_uint2(h0, h1);
        fl[t4] = make_uint2(l0, l1);
      }
      if (blockIdx.x == 0) {
        // centred form kept for the record: Wfold[12] is the pre-ReLU value at the batch centre E[r]
        float4* wfg = reinterpret_cast<float4*>(p.Wfold + o * 16);
        float ctr = wf[12];
#pragma unroll
        for (int j = 0; j < NR; ++j) ctr = fmaf(wf[j], Ef[j], ctr);
        wfg[0] = make_float4(wf[0], wf[1], wf[2], wf[3]);
        wfg[1] = make_float4(wf[4], wf[5], wf[6], wf[7]);
        wfg[2] = make_float4(wf[8], wf[9], wf[10], wf[11]);
        wfg[3] = make_float4(ctr, 0.f, 0.f, 0.f);
        p.stats[NSTAT + o] = mean_h;
        p.stats[NSTAT + COUT + o] = rstd_h;
        p.stats[NSTAT + 2 * COUT + o] = mean_d;
        p.stats[NSTAT + 3 * COUT + o] = rstd_d;
      }
    }
    __syncthreads();
  };

  const uint2* wfh = reinterpret_cast<const uint2*>(sm + pl.wfrag_h);
  const uint2* wfl = reinterpret_cast<const uint2*>(sm + pl.wfrag_l);
  const uint32_t stage_u32 = smem_u32(sm + pl.stage + warp * 4096);
  const bf16* RTh = reinterpret_cast<const bf16*>(sm + pl.rth);
  const bf16* RTl = reinterpret_cast<const bf16*>(sm + pl.rtl);
  float D0[4] = {0.f, 0.f, 0.f, 0.f}, D1[4] = {0.f, 0.f, 0.f, 0.f};   // this warp's share of R^T R (slots 0-7 | 8-15)
  const int ntiles_all = r16u(T * V) / 16;

  // ------------------------------------------------------------------------------------------------------------
  // eval: running statistics, no batch dependency -> samples stream through phase 1 + expansion, no grid barrier
  // ------------------------------------------------------------------------------------------------------------
  if (!training) {
    // the fold scratch and the staging buffers overlay X2 / Ms; the W fragments overlay the coefficient staging (dead)
    bool first = true;
    for (int n = blockIdx.x; n < N; n += gridDim.x) {
      if (!first) {
        if (lane == 0) bulk_wait_read0();   // the previous sample's stores have read the staging buffers
        __syncthreads();
        zero_region(sm + pl.work, pl.p1_end - pl.work, tid);
        __syncthreads();
      }
      phase1_sample<CBT>(p, fa, pl, sm, coef_s, n, L, false, false, D0, D1, acquire_x(n, first));
      if (first) fold_weights();   // after phase 1: its scratch overlays the (now dead) X2 operand
      first = false;
      for (int mt = warp; mt < ntiles_all; mt += kWarps)
        expand_tile(RTh, RTl, pl.pp, mt * 16, wfh, wfl, stage_u32, &tmY, mt * 16, n, L);
    }
    if (lane == 0) bulk_wait0();
    return;
  }

  // ------------------------------------------------------------------------------------------------------------
  // training, phase 1
  // ------------------------------------------------------------------------------------------------------------
  bool had_sample = false;
  int my_n = -1;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    if (had_sample) {
      __syncthreads();   // previous sample's readers are done
      zero_region(sm + pl.work, pl.p1_end - pl.work, tid);
      __syncthreads();
    }
    const float* xs = acquire_x(n, !had_sample);
    had_sample = true;
    my_n = n;
    phase1_sample<CBT>(p, fa, pl, sm, coef_s, n, L, !resident, true, D0, D1, xs);
  }
  stamp(6);
  if (had_sample) {
    // D0: rows g / g+8, columns 2tq, 2tq+1;  D1: columns + 8
    atomicAdd(&momS[g * 16 + 2 * tq], D0[0]);
    atomicAdd(&momS[g * 16 + 2 * tq + 1], D0[1]);
    atomicAdd(&momS[(g + 8) * 16 + 2 * tq], D0[2]);
    atomicAdd(&momS[(g + 8) * 16 + 2 * tq + 1], D0[3]);
    atomicAdd(&momS[g * 16 + 8 + 2 * tq], D1[0]);
    atomicAdd(&momS[g * 16 + 8 + 2 * tq + 1], D1[1]);
    atomicAdd(&momS[(g + 8) * 16 + 8 + 2 * tq], D1[2]);
    atomicAdd(&momS[(g + 8) * 16 + 8 + 2 * tq + 1], D1[3]);
    __syncthreads();
    if (tid < NR + NR * (NR + 1) / 2) {
      float val;
      if (tid < NR) {
        val = momS[tid * 16 + 12];          // sum r_j * 1
      } else {
        int rem = tid - NR, pa = 0;
        while (rem >= NR - pa) { rem -= NR - pa; ++pa; }
        val = momS[pa * 16 + pa + rem];
      }
      atomicAdd(mom_cur + (blockIdx.x % kUseSlots) * NMOM + tid, (double)val);
    }
  }
  stamp(7);
  grid_barrier(fa.ctrl, gen);
  stamp(8);
  fold_weights();     // (scratch overlays X2, dead by now)
  stamp(10);

  if (resident) {
    if (my_n >= 0)
      for (int mt = warp; mt < ntiles_all; mt += kWarps)
        expand_tile(RTh, RTl, pl.pp, mt * 16, wfh, wfl, stage_u32, &tmY, mt * 16, my_n, L);
    stamp(11);
    if (lane == 0) bulk_wait0();
    stamp(12);
    return;
  }

  // ------------------------------------------------------------------------------------------------------------
  // training, generic mode (more samples than CTAs): warp pairs recompute z per 8-frame unit from Mop
  // ------------------------------------------------------------------------------------------------------------
  const int team = warp >> 1, wt = warp & 1, l64 = wt * 32 + lane;
  uint8_t* tbase = sm + pl.team + team * pl.team_bytes;
  bf16* AopH = reinterpret_cast<bf16*>(tbase);                 // RT layout: [16 slots][pup]
  bf16* AopL = reinterpret_cast<bf16*>(tbase + pl.aop_l);
  bf16* XTh = reinterpret_cast<bf16*>(tbase + pl.xt_h);
  bf16* XTl = reinterpret_cast<bf16*>(tbase + pl.xt_l);
  const int PUP = pl.pup;
  {   // team-private operands: zero everything once, then the ones row (slot 12, hi)
    uint4* z4 = reinterpret_cast<uint4*>(tbase);
    for (int i = l64; i < pl.team_bytes >> 4; i += 64) z4[i] = make_uint4(0u, 0u, 0u, 0u);
    team_sync(team);
    for (int c = l64; c < PUP / 2; c += 64) *reinterpret_cast<uint32_t*>(AopH + 12 * PUP + 2 * c) = 0x3f803f80u;
  }
  const int upn = (T + kUnitFrames - 1) / kUnitFrames;
  const int nunits = N * upn;
  for (int uid = blockIdx.x + gridDim.x * team; uid < nunits; uid += gridDim.x * kTeams) {
    const int n = uid / upn, f0 = (uid - n * upn) * kUnitFrames;
    const int nf = min(kUnitFrames, T - f0), P = nf * V;
    team_sync(team);   // the team's previous unit is fully consumed (and the one-time init above is visible)
    {   // prep: x of the unit -> XT[(t,a)][u] (stage-1 A operand) and r slots 9..11
      const float* xg = p.x + ((size_t)n * T + f0) * V * 3;
      for (int e = l64; e < P * 3; e += 64) {
        const int pos = e / 3, a = e - 3 * pos, tl = pos / V, u = pos - tl * V;
        uint16_t hi, lo;
        split1(xg[e], hi, lo);
        *reinterpret_cast<uint16_t*>(XTh + (3 * tl + a) * UP + u) = hi;
        *reinterpret_cast<uint16_t*>(XTl + (3 * tl + a) * UP + u) = lo;
        *reinterpret_cast<uint16_t*>(AopH + (9 + a) * PUP + pos) = hi;
        *reinterpret_cast<uint16_t*>(AopL + (9 + a) * PUP + pos) = lo;
      }
    }
    team_sync(team);
    // stage 1: z[(t,a)][c] = sum_u XT[(t,a)][u] M[u][c]; warp wt owns rows 16 wt .. 16 wt + 15
    if (wt * 16 < 3 * nf) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int a_off = (wt * 16 + (lj & 1) * 8 + lr) * UP + ks * 16 + (lj >> 1) * 8;
        ldsm_x4(smem_u32(XTh + a_off), ah[ks]);
        ldsm_x4(smem_u32(XTl + a_off), al[ks]);
      }
      const int m0 = wt * 16 + g, m1 = m0 + 8;
      const int tl0 = m0 / 3, a0 = m0 - 3 * tl0, tl1 = m1 / 3, a1 = m1 - 3 * tl1;
      const bool rv0 = tl0 < nf, rv1 = tl1 < nf;
      const int d0 = a0 * PUP + tl0 * V + 2 * tq, d1 = a1 * PUP + tl1 * V + 2 * tq;
      const uint32_t* moph = fa.mop + ((size_t)n * 2 * NC) * (UP / 2);
      const uint32_t* mopl = moph + (size_t)NC * (UP / 2);
      uint32_t bh[2][4], bl[2][4];   // double-buffered B fragments: tile nt + 1 is in flight while nt is consumed
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bh[0][q] = ld_global_u32(moph + ((g * UP) >> 1) + q * 4 + tq);
        bl[0][q] = ld_global_u32(mopl + ((g * UP) >> 1) + q * 4 + tq);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int cur = nt & 1, nxt = cur ^ 1;
        if (nt + 1 < NT) {
          const int row = (((nt + 1) * 8 + g) * UP) >> 1;
#pragma unroll
          for (int q = 0; q < 4; ++q) {   // k = q*8 + 2tq, 2tq+1  (q = 2 ks + {0, 1})
            bh[nxt][q] = ld_global_u32(moph + row + q * 4 + tq);
            bl[nxt][q] = ld_global_u32(mopl + row + q * 4 + tq);
          }
        }
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        mma(z, ah[0], bh[cur][0], bh[cur][1]);
        mma(z, ah[1], bh[cur][2], bh[cur][3]);
        mma(z, ah[0], bl[cur][0], bl[cur][1]);
        mma(z, ah[1], bl[cur][2], bl[cur][3]);
        mma(z, al[0], bh[cur][0], bh[cur][1]);
        mma(z, al[1], bh[cur][2], bh[cur][3]);
        const int i = nt / CBT, vt = (nt - i * CBT) * 8;
        if (vt + 2 * tq < V) {
          uint32_t hi, lo;
          if (rv0) {
            split2(z[0], z[1], hi, lo);
            *reinterpret_cast<uint32_t*>(AopH + 3 * i * PUP + vt + d0) = hi;
            *reinterpret_cast<uint32_t*>(AopL + 3 * i * PUP + vt + d0) = lo;
          }
          if (rv1) {
            split2(z[2], z[3], hi, lo);
            *reinterpret_cast<uint32_t*>(AopH + 3 * i * PUP + vt + d1) = hi;
            *reinterpret_cast<uint32_t*>(AopL + 3 * i * PUP + vt + d1) = lo;
          }
        }
      }
    }
    team_sync(team);
    // expansion: 16-position tiles alternate between the two warps
    const int ntiles = (P + 15) >> 4;
    for (int mt = wt; mt < ntiles; mt += 2) expand_tile(AopH, AopL, PUP, mt * 16, wfh, wfl, stage_u32, &tmY, f0 * V + mt * 16, n, L);
  }
  if (lane == 0) bulk_wait0();
  stamp(12);
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", what, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <int CBT>
int launch_fused(const afb_gcn0_fwd_t* p, cudaStream_t st) {
  auto kern = gcn0_fused_kernel<CBT>;
  int rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // resident plan first (eval always; training when every sample gets its own CTA), else the generic plan
  SmemPlan pl = make_plan<CBT>(p->T, p->V, true);
  if (pl.total > 220 * 1024) return -1;
  if ((rc = set_smem(kern, 220 * 1024, "gcn0_fused"))) return rc;
  int occ = 0;
  auto query = [&](const SmemPlan& q) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, (size_t)q.total) != cudaSuccess || occ < 1) return false;
    if (occ > 2) occ = 2;
    return true;
  };
  if (!query(pl)) { set_error("gcn0_fused: occupancy query failed"); return AFB_ERR_DRIVER; }
  int resident = 1;
  if (p->training && p->N > sms * occ) {
    resident = 0;
    pl = make_plan<CBT>(p->T, p->V, false);
    if (pl.total > 220 * 1024) return -1;
    if (!query(pl)) { set_error("gcn0_fused: occupancy query failed"); return AFB_ERR_DRIVER; }
  }
  CUtensorMap tm;
  if ((rc = make_tensor_map_bf16(&tm, p->y, COUT, (uint64_t)p->T * p->V, (uint64_t)p->N, COUT, (uint64_t)p->T * p->V * COUT, 64, 16)))
    return rc;
  FusedArgs fa;
  fa.mop = reinterpret_cast<uint32_t*>(p->Aop);
  fa.ctrl = reinterpret_cast<unsigned int*>(p->counter) + 4;   // words 4..6: the two-kernel path owns word 0
  fa.moments = p->moments + (size_t)kSlots * NMOM;            // [1], [2] of the [3][slots][NMOM] workspace
  fa.resident = resident;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sms * occ));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = (size_t)pl.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // the training modes hold a grid barrier: all CTAs must be co-resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p->training ? 1 : 0;
  if (!p->training) {   // eval streams samples without any inter-CTA dependency: plain launch, as many CTAs as help
    const int want = p->N < sms * occ ? p->N : sms * occ;
    cfg.gridDim = dim3((unsigned)(want > 0 ? want : 1));
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *p, tm, fa);
  if (e != cudaSuccess) {
    set_error("gcn0_fused: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return (int)e;
  }
  return 0;
}

}  // namespace

// 0: launched; -1: shape not covered (caller falls back to the two-kernel path); > 0: error (message set)
int gcn0_fused_launch(const afb_gcn0_fwd_t* p, cudaStream_t st) {
  static const bool off = getenv("AFB_GCN0_FUSED") != nullptr && getenv("AFB_GCN0_FUSED")[0] == '0';
  if (off || p->precise == 1 || p->y_dtype != AFB_BF16 || p->Cout != COUT || p->Aop == nullptr) return -1;
  if (p->V < 2 || (p->V & 1) || p->V > 24 || p->IC <= 0 || p->IC > 36) return -1;
  if (((uintptr_t)p->y & 15) != 0 || ((uintptr_t)p->x & 7) != 0) return -1;
  const int cbt = (p->V + 7) / 8;
  if (cbt == 3) return launch_fused<3>(p, st);
  if (cbt == 2) return launch_fused<2>(p, st);
  if (cbt == 1) return launch_fused<1>(p, st);
  return -1;
}

// copies the phase time stamps of the last fused launch (kStamps per CTA, first `ctas` CTAs) to the host
int gcn0_fused_stamps(unsigned long long* out, int ctas) {
  if (ctas > kStampCtas) ctas = kStampCtas;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_stamps, sizeof(unsigned long long) * kStamps * ctas);
  return e == cudaSuccess ? 0 : (int)e;
}

size_t gcn0_fused_mop_bytes(int N, int V) {
  const int cb = ((V + 7) / 8) * 8;
  return (size_t)N * 2 * 3 * cb * UP * 2;
}

}  // namespace afb
