// Tensor-core versions of the per-sample graph pieces of the general unit_agcn(C -> C_out) (model/unit_agcn.py:80-88;
// the layer of the TCN_GCN_unit stack, model/ST_TR/ST_TR_new.py:355-385).  bf16 activations, mma.sync.m16n8k16, V <= 48.
// Per frame every piece is a small matrix product with the joints as one dimension -- far below a tcgen05 tile (64/128
// rows with the accumulator in TMEM), so these are warp-level MMAs on ldmatrix fragments like the attention kernels:
//
//   aggregate_fwd   z_t,i [v][c] = sum_u M_i[u][v] x_t[u][c]          A = M_i^T (shared, per sample), B = x_t (ldmatrix.trans)
//                   EXACT: M as bf16 hi + lo and z written as (hi | lo | hi) K-concatenated slabs -- the A operand of the
//                   3-term conv_d GEMM that forms the BatchNorm input at fp32 accuracy (exact ReLU masks), no split pass
//   aggregate_bwd   dx_t [u][c] (+)= sum_i sum_v M_i[u][v] dz_t,i[v][c]   A = M_i, B = dz_t,i (ldmatrix.trans), K = 3 VP
//                   dM_i [u][v]   = sum_t sum_c x_t[u][c] dz_t,i[v][c]   A = x_t, B = dz_t,i ([n][k] rows, plain ldmatrix);
//                   each warp owns one (i, 16-row) output tile over all frames of its CTA
// The fp32 parity mode keeps the CUDA-core kernels of agcn.cu.
#include "common.cuh"
#include "mma_utils.cuh"

namespace afb {
namespace {

using namespace mmau;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int CC = 64;          // channel chunk staged per round
constexpr int CP = CC + 8;      // its shared-memory pitch (144 B: odd multiple of 16 -> conflict-free ldmatrix rows)

template <typename K>
int set_smem(K kernel, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", what, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// one warp copies rows [0, nrows) x CH channels (16-byte pieces) of a bf16 matrix into its [VP][CH + 8] tile
template <int CH>
__device__ __forceinline__ void warp_stage(bf16* dst, const bf16* src, int64_t ld, int nrows, int lane) {
  constexpr int kPieces = CH / 8;
  for (int e = lane; e < nrows * kPieces; e += 32) {
    const int r = e / kPieces, q = e - r * kPieces;
    cp_async16(smem_u32(dst + r * (CH + 8) + q * 8), src + (int64_t)r * ld + q * 8);
  }
}
constexpr int kFwdFramesPerWarp = 4;   // aggregate forward: frames per warp (amortises the per-CTA staging of M)
constexpr int CX = 32;          // channel chunk of the dx kernel (three dz slabs staged per warp: 64 would allow one CTA per SM only)
constexpr int CXP = CX + 8;

// ------------------------------------------------------------------------------------------------------------------
// aggregate forward: grid (N * ceil(T / kWarps)), one frame per warp
// ------------------------------------------------------------------------------------------------------------------
template <int VP, bool EXACT>
__global__ void __launch_bounds__(kThreads, 2) agcn_aggr_fwd_mma_kernel(const bf16* __restrict__ x, const float* __restrict__ Mmat,
                                                                       bf16* __restrict__ z, int T, int V, int C, int groups) {
  constexpr int UP = VP + 8, KS = VP / 16, MT = VP / 16;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Ah = reinterpret_cast<bf16*>(smraw);                 // [3][VP v][UP u] = M_i^T (hi)
  bf16* Al = Ah + 3 * VP * UP;                               // (lo; EXACT only)
  bf16* Xs = Ah + (EXACT ? 2 : 1) * 3 * VP * UP;             // [kWarps][VP][CP]
  bf16* Os = Xs + kWarps * VP * CP;                          // [kWarps][EXACT ? 2 : 1][16][CP] output staging (hi, lo)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3, lj = lane >> 3, lr = lane & 7;
  const int n = blockIdx.x / groups, tbase = (blockIdx.x % groups) * (kWarps * kFwdFramesPerWarp);
  // zero everything once (operand padding must be exact zeros), then M^T
  {
    const int words = ((EXACT ? 2 : 1) * 3 * VP * UP + kWarps * VP * CP) / 2;
    uint32_t* w = reinterpret_cast<uint32_t*>(smraw);
    for (int e = tid; e < words; e += kThreads) w[e] = 0u;
  }
  __syncthreads();
  const float* Mg = Mmat + (int64_t)n * 3 * V * V;
  for (int e = tid; e < 3 * V * V; e += kThreads) {
    const int i = e / (V * V), r = e - i * V * V, u = r / V, v = r - u * V;
    const float m = Mg[e];
    const bf16 h = __float2bfloat16_rn(m);
    Ah[(i * VP + v) * UP + u] = h;
    if (EXACT) Al[(i * VP + v) * UP + u] = __float2bfloat16_rn(m - __bfloat162float(h));
  }
  __syncthreads();
  bf16* Ow = Os + warp * (EXACT ? 2 : 1) * 16 * CP;
  // this warp's work items: (frame, channel chunk) pairs
  const int chunks = C / CC;
  int nfr = 0;
  for (int k = 0; k < kFwdFramesPerWarp; ++k) nfr += (tbase + warp + k * kWarps) < T ? 1 : 0;
  const int items = nfr * chunks;
  if (items == 0) return;
  const int ldz = (EXACT ? 9 : 3) * C;
  bf16* Xw = Xs + warp * VP * CP;
  for (int it = 0; it < items; ++it) {
    const int t = tbase + warp + (it / chunks) * kWarps, c0 = (it % chunks) * CC;
    const int64_t row0 = ((int64_t)n * T + t) * V;
    warp_stage<CC>(Xw, x + row0 * C + c0, C, V, lane);   // (the previous item's ldmatrix reads are warp-synchronous: done)
    cp_async_wait_all();
    __syncwarp();
    uint32_t b[KS][8][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t r4[4];
        ldsm_x4_t(smem_u32(Xw + (ks * 16 + (lj & 1) * 8 + lr) * CP + q * 16 + (lj >> 1) * 8), r4);
        b[ks][2 * q][0] = r4[0]; b[ks][2 * q][1] = r4[1]; b[ks][2 * q + 1][0] = r4[2]; b[ks][2 * q + 1][1] = r4[3];
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt * 16 >= V) continue;
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a[4];
          const int a_off = (i * VP + mt * 16 + (lj & 1) * 8 + lr) * UP + ks * 16 + (lj >> 1) * 8;
          ldsm_x4(smem_u32(Ah + a_off), a);
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) mma(acc[nt], a, b[ks][nt][0], b[ks][nt][1]);
          if (EXACT) {
            ldsm_x4(smem_u32(Al + a_off), a);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) mma(acc[nt], a, b[ks][nt][0], b[ks][nt][1]);
          }
        }
        // fragments -> staging tile(s) [16][CP] (conflict-free 4-byte stores) -> 16-byte row-contiguous global stores
        __syncwarp();   // the previous tile's copy-out has finished reading the staging tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint32_t h0 = pack2(acc[nt][0], acc[nt][1]), h1 = pack2(acc[nt][2], acc[nt][3]);
          *reinterpret_cast<uint32_t*>(Ow + g * CP + nt * 8 + 2 * tq) = h0;
          *reinterpret_cast<uint32_t*>(Ow + (g + 8) * CP + nt * 8 + 2 * tq) = h1;
          if (EXACT) {
            *reinterpret_cast<uint32_t*>(Ow + (16 + g) * CP + nt * 8 + 2 * tq) =
                pack2(acc[nt][0] - __uint_as_float(h0 << 16), acc[nt][1] - __uint_as_float(h0 & 0xffff0000u));
            *reinterpret_cast<uint32_t*>(Ow + (16 + g + 8) * CP + nt * 8 + 2 * tq) =
                pack2(acc[nt][2] - __uint_as_float(h1 << 16), acc[nt][3] - __uint_as_float(h1 & 0xffff0000u));
          }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 16 rows x 8 pieces of 16 bytes: lane -> (row = lane / 8 + 4 k, piece = lane % 8)
          const int r = (lane >> 3) + 4 * k, q = lane & 7, v = mt * 16 + r;
          if (v < V) {
            bf16* dst = z + (row0 + v) * ldz + i * C + c0 + q * 8;
            const uint4 hi = *reinterpret_cast<const uint4*>(Ow + r * CP + q * 8);
            *reinterpret_cast<uint4*>(dst) = hi;
            if (EXACT) {
              *reinterpret_cast<uint4*>(dst + 3 * C) = *reinterpret_cast<const uint4*>(Ow + (16 + r) * CP + q * 8);
              *reinterpret_cast<uint4*>(dst + 6 * C) = hi;
            }
          }
        }
      }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// aggregate backward, dx: grid (N * ceil(T / kWarps)), one frame per warp; the three dz slabs of a channel chunk staged
// ------------------------------------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(kThreads, 2) agcn_aggr_bwd_dx_mma_kernel(const bf16* __restrict__ dz, const float* __restrict__ Mmat,
                                                                          bf16* __restrict__ dx, int accumulate, int T, int V, int C,
                                                                          int groups) {
  constexpr int UP = VP + 8, KS = VP / 16, MT = VP / 16;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Ah = reinterpret_cast<bf16*>(smraw);      // [3][VP u][UP v] = M_i (hi)
  bf16* Zs = Ah + 3 * VP * UP;                    // [kWarps][3][VP][CXP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3, lj = lane >> 3, lr = lane & 7;
  const int n = blockIdx.x / groups, t = (blockIdx.x % groups) * kWarps + warp;
  {
    const int words = (3 * VP * UP + kWarps * 3 * VP * CXP) / 2;
    uint32_t* w = reinterpret_cast<uint32_t*>(smraw);
    for (int e = tid; e < words; e += kThreads) w[e] = 0u;
  }
  __syncthreads();
  const float* Mg = Mmat + (int64_t)n * 3 * V * V;
  for (int e = tid; e < 3 * V * V; e += kThreads) {
    const int i = e / (V * V), r = e - i * V * V, u = r / V, v = r - u * V;
    Ah[(i * VP + u) * UP + v] = __float2bfloat16_rn(Mg[e]);
  }
  __syncthreads();
  if (t >= T) return;
  bf16* Zw = Zs + warp * 3 * VP * CXP;
  const int64_t row0 = ((int64_t)n * T + t) * V;
  for (int c0 = 0; c0 < C; c0 += CX) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; ++i) warp_stage<CX>(Zw + i * VP * CXP, dz + row0 * 3 * C + i * C + c0, 3 * C, V, lane);
    cp_async_wait_all();
    __syncwarp();
    float acc[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t b[4][2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t r4[4];
          ldsm_x4_t(smem_u32(Zw + i * VP * CXP + (ks * 16 + (lj & 1) * 8 + lr) * CXP + q * 16 + (lj >> 1) * 8), r4);
          b[2 * q][0] = r4[0]; b[2 * q][1] = r4[1]; b[2 * q + 1][0] = r4[2]; b[2 * q + 1][1] = r4[3];
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a[4];
          ldsm_x4(smem_u32(Ah + (i * VP + mt * 16 + (lj & 1) * 8 + lr) * UP + ks * 16 + (lj >> 1) * 8), a);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma(acc[mt][nt], a, b[nt][0], b[nt][1]);
        }
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int u0 = mt * 16 + g, u1 = u0 + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (u0 < V) {
          uint32_t* d = reinterpret_cast<uint32_t*>(dx + (row0 + u0) * C + c0 + nt * 8 + 2 * tq);
          float a0 = acc[mt][nt][0], a1 = acc[mt][nt][1];
          if (accumulate) { const uint32_t o = *d; a0 += __uint_as_float(o << 16); a1 += __uint_as_float(o & 0xffff0000u); }
          *d = pack2(a0, a1);
        }
        if (u1 < V) {
          uint32_t* d = reinterpret_cast<uint32_t*>(dx + (row0 + u1) * C + c0 + nt * 8 + 2 * tq);
          float a0 = acc[mt][nt][2], a1 = acc[mt][nt][3];
          if (accumulate) { const uint32_t o = *d; a0 += __uint_as_float(o << 16); a1 += __uint_as_float(o & 0xffff0000u); }
          *d = pack2(a0, a1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// aggregate backward, dM: grid (N, splits); the CTA stages FG frames x one channel chunk (x and the three dz slabs) per
// round; warp w owns output tiles (i, mt) = w, w + kWarps, ... and keeps them in registers over all rounds
// ------------------------------------------------------------------------------------------------------------------
constexpr int FG = 4;

template <int VP>
__global__ void __launch_bounds__(kThreads, 2) agcn_aggr_bwd_dm_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dz,
                                                                          float* __restrict__ dM, int T, int V, int C, int splits) {
  constexpr int MT = VP / 16, NT = VP / 8, ITEMS = 3 * MT, PER_WARP = (ITEMS + kWarps - 1) / kWarps;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* Xs = reinterpret_cast<bf16*>(smraw);      // [FG][VP][CP]
  bf16* Zs = Xs + FG * VP * CP;                   // [FG][3][VP][CP]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3, lj = lane >> 3, lr = lane & 7;
  const int n = blockIdx.x, sp = blockIdx.y;
  const int per = (T + splits - 1) / splits, t_begin = sp * per, t_end = min(T, t_begin + per);
  {
    const int words = (FG * 4 * VP * CP) / 2;
    uint32_t* w = reinterpret_cast<uint32_t*>(smraw);
    for (int e = tid; e < words; e += kThreads) w[e] = 0u;
  }
  float acc[PER_WARP][NT][4];
#pragma unroll
  for (int q = 0; q < PER_WARP; ++q)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[q][nt][0] = acc[q][nt][1] = acc[q][nt][2] = acc[q][nt][3] = 0.f;
  constexpr int kPieces = CC / 8;
  for (int t0 = t_begin; t0 < t_end; t0 += FG) {
    const int nf = min(FG, t_end - t0);
    for (int c0 = 0; c0 < C; c0 += CC) {
      __syncthreads();   // previous round's readers are done (and the zero fill, first time)
      // stage: (frame, slab 0 = x / 1..3 = dz_i, row, piece)
      const int total = nf * 4 * V * kPieces;
      for (int e = tid; e < total; e += kThreads) {
        const int q = e % kPieces, r = (e / kPieces) % V, sl = (e / (kPieces * V)) % 4, f = e / (kPieces * V * 4);
        const int64_t row = ((int64_t)n * T + t0 + f) * V + r;
        if (sl == 0) cp_async16(smem_u32(Xs + (f * VP + r) * CP + q * 8), x + row * C + c0 + q * 8);
        else cp_async16(smem_u32(Zs + ((f * 3 + sl - 1) * VP + r) * CP + q * 8), dz + row * 3 * C + (sl - 1) * C + c0 + q * 8);
      }
      cp_async_wait_all();
      __syncthreads();
#pragma unroll
      for (int q = 0; q < PER_WARP; ++q) {
        const int item = warp + q * kWarps;
        if (item >= ITEMS) continue;
        const int i = item / MT, mt = item - i * MT;
        if (mt * 16 >= V) continue;
        for (int f = 0; f < nf; ++f) {
#pragma unroll
          for (int ks = 0; ks < CC / 16; ++ks) {
            uint32_t a[4];
            ldsm_x4(smem_u32(Xs + (f * VP + mt * 16 + (lj & 1) * 8 + lr) * CP + ks * 16 + (lj >> 1) * 8), a);
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
              uint32_t r4[4];   // rows = v (n index), columns = c (k index): plain ldmatrix gives the col-major B fragments
              ldsm_x4(smem_u32(Zs + ((f * 3 + i) * VP + np * 16 + (lj >> 1) * 8 + lr) * CP + ks * 16 + (lj & 1) * 8), r4);
              mma(acc[q][2 * np], a, r4[0], r4[1]);
              mma(acc[q][2 * np + 1], a, r4[2], r4[3]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PER_WARP; ++q) {
    const int item = warp + q * kWarps;
    if (item >= ITEMS) continue;
    const int i = item / MT, mt = item - i * MT;
    const int u0 = mt * 16 + g, u1 = u0 + 8;
    float* out = dM + ((int64_t)n * 3 + i) * V * V;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int v = nt * 8 + 2 * tq;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (v + j < V) {
          if (u0 < V) { if (splits > 1) atomicAdd(out + u0 * V + v + j, acc[q][nt][j]); else out[u0 * V + v + j] = acc[q][nt][j]; }
          if (u1 < V) { if (splits > 1) atomicAdd(out + u1 * V + v + j, acc[q][nt][2 + j]); else out[u1 * V + v + j] = acc[q][nt][2 + j]; }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// scores forward: grid (N, 3).  S_i[u][v] = sum_t sum_ic theta_t[u][ic] phi_t[v][ic] / (IC T); warps take frames round robin
// and keep a private S in registers; one shared-memory reduction, then softmax over u, M = P + A + PA.
// TIN = float: theta/phi arrive at fp32 accuracy (exact-mask forward) and are applied as bf16 hi + lo (3 MMA terms).
// ------------------------------------------------------------------------------------------------------------------
constexpr int ICC = 32;   // inner-channel chunk staged per round (IC = 16 runs with a chunk of 16)

template <int VP, typename TIN>
__global__ void __launch_bounds__(kThreads, 2) agcn_scores_fwd_mma_kernel(const TIN* __restrict__ thph, int ld, const float* __restrict__ A,
                                                                         const float* __restrict__ PA, float* __restrict__ P,
                                                                         float* __restrict__ Mmat, int T, int V, int IC) {
  constexpr bool EXACT = sizeof(TIN) == 4;
  constexpr int MT = VP / 16, NT = VP / 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int icc = IC < ICC ? IC : ICC, pitch = icc + 8;
  const int tile = VP * pitch;                                    // elements of one [VP][pitch] operand tile
  bf16* base = reinterpret_cast<bf16*>(smraw);                    // per warp: th_hi, ph_hi, (th_lo, ph_lo)
  const int per_warp = (EXACT ? 4 : 2) * tile;
  float* S = reinterpret_cast<float*>(base + kWarps * per_warp);  // [VP][VP] reduction target
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3, lj = lane >> 3, lr = lane & 7;
  const int n = blockIdx.x, i = blockIdx.y;
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smraw);
    const int words = (kWarps * per_warp) / 2 + VP * VP;
    for (int e = tid; e < words; e += kThreads) w[e] = 0u;
  }
  __syncthreads();
  bf16* thh = base + warp * per_warp;
  bf16* phh = thh + tile;
  bf16* thl = phh + tile;
  bf16* phl = thl + tile;
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  for (int t = warp; t < T; t += kWarps) {
    const TIN* src = thph + ((int64_t)n * T + t) * V * ld;
    for (int c0 = 0; c0 < IC; c0 += icc) {
      __syncwarp();
      if (EXACT) {
        // six 16-byte loads per lane are issued before the first is converted: the one-load-at-a-time loop kept 512 B per
        // warp in flight and ran the whole kernel at 1.2 TB/s (469 us for 553 MB at C = 128, T = 32, N = 1024)
        const int quads = icc / 4, total = 2 * V * quads;
        constexpr int kBatch = 6;
        for (int e0 = lane; e0 < total; e0 += 32 * kBatch) {
          float4 v4[kBatch];
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            const int e = e0 + 32 * b;
            if (e < total) {
              const int which = e / (V * quads), r = (e / quads) % V, q = e % quads;
              v4[b] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (int64_t)r * ld + (which ? 3 + i : i) * IC + c0 + q * 4);
            }
          }
#pragma unroll
          for (int b = 0; b < kBatch; ++b) {
            const int e = e0 + 32 * b;
            if (e < total) {
              const int which = e / (V * quads), r = (e / quads) % V, q = e % quads;
              const uint32_t h0 = pack2(v4[b].x, v4[b].y), h1 = pack2(v4[b].z, v4[b].w);
              const uint32_t l0 = pack2(v4[b].x - __uint_as_float(h0 << 16), v4[b].y - __uint_as_float(h0 & 0xffff0000u));
              const uint32_t l1 = pack2(v4[b].z - __uint_as_float(h1 << 16), v4[b].w - __uint_as_float(h1 & 0xffff0000u));
              bf16* dh = (which ? phh : thh) + r * pitch + q * 4;
              bf16* dl = (which ? phl : thl) + r * pitch + q * 4;
              *reinterpret_cast<uint2*>(dh) = make_uint2(h0, h1);
              *reinterpret_cast<uint2*>(dl) = make_uint2(l0, l1);
            }
          }
        }
      } else {
        const int pieces = icc / 8;
        for (int e = lane; e < 2 * V * pieces; e += 32) {
          const int which = e / (V * pieces), r = (e / pieces) % V, q = e % pieces;
          cp_async16(smem_u32((which ? phh : thh) + r * pitch + q * 8),
                     reinterpret_cast<const bf16*>(src) + (int64_t)r * ld + (which ? 3 + i : i) * IC + c0 + q * 8);
        }
        cp_async_wait_all();
      }
      __syncwarp();
      for (int ks = 0; ks < icc / 16; ++ks) {
        uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
          uint32_t r4[4];   // phi rows = v (n index), columns = ic (k index)
          const int off = (np * 16 + (lj >> 1) * 8 + lr) * pitch + ks * 16 + (lj & 1) * 8;
          ldsm_x4(smem_u32(phh + off), r4);
          bh[2 * np][0] = r4[0]; bh[2 * np][1] = r4[1]; bh[2 * np + 1][0] = r4[2]; bh[2 * np + 1][1] = r4[3];
          if (EXACT) {
            ldsm_x4(smem_u32(phl + off), r4);
            bl[2 * np][0] = r4[0]; bl[2 * np][1] = r4[1]; bl[2 * np + 1][0] = r4[2]; bl[2 * np + 1][1] = r4[3];
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          uint32_t a[4];
          const int off = (mt * 16 + (lj & 1) * 8 + lr) * pitch + ks * 16 + (lj >> 1) * 8;
          ldsm_x4(smem_u32(thh + off), a);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            mma(acc[mt][nt], a, bh[nt][0], bh[nt][1]);
            if (EXACT) mma(acc[mt][nt], a, bl[nt][0], bl[nt][1]);
          }
          if (EXACT) {
            ldsm_x4(smem_u32(thl + off), a);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma(acc[mt][nt], a, bh[nt][0], bh[nt][1]);
          }
        }
      }
    }
  }
  // reduce the warps' partial S
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int u0 = mt * 16 + g, v = nt * 8 + 2 * tq;
      atomicAdd(S + u0 * VP + v, acc[mt][nt][0]);
      atomicAdd(S + u0 * VP + v + 1, acc[mt][nt][1]);
      atomicAdd(S + (u0 + 8) * VP + v, acc[mt][nt][2]);
      atomicAdd(S + (u0 + 8) * VP + v + 1, acc[mt][nt][3]);
    }
  __syncthreads();
  const float inv = 1.0f / (float)(IC * T);
  // softmax over u for column v: one warp per column, lanes over u
  for (int v = warp; v < V; v += kWarps) {
    float s0 = lane < V ? S[lane * VP + v] * inv : -INFINITY;
    float s1 = lane + 32 < V ? S[(lane + 32) * VP + v] * inv : -INFINITY;
    float mx = fmaxf(s0, s1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e0 = lane < V ? __expf(s0 - mx) : 0.f, e1 = lane + 32 < V ? __expf(s1 - mx) : 0.f;
    float den = e0 + e1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    const float rden = 1.0f / den;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u = lane + 32 * h;
      if (u < V) {
        const float pr = (h ? e1 : e0) * rden;
        const int idx = (i * V + u) * V + v;
        const int64_t gi = (int64_t)n * 3 * V * V + idx;
        P[gi] = pr;
        Mmat[gi] = pr + A[idx] + PA[idx];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// scores backward: grid (N, 3).  dPA += dM; dS = P (dM - colsum_u(P dM)) / (IC T);
// d(theta)_t [u][ic] = sum_v dS[u][v] phi_t[v][ic],  d(phi)_t [v][ic] = sum_u dS[u][v] theta_t[u][ic]   (one frame per warp round)
// ------------------------------------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(kThreads, 2) agcn_scores_bwd_mma_kernel(const bf16* __restrict__ thph, int ld, const float* __restrict__ P,
                                                                         const float* __restrict__ dM, float* __restrict__ dPA,
                                                                         bf16* __restrict__ dthph, int T, int V, int IC) {
  constexpr int UP = VP + 8, MT = VP / 16, KS = VP / 16;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int icc = IC < ICC ? IC : ICC, pitch = icc + 8, tile = VP * pitch;
  float* dSf = reinterpret_cast<float*>(smraw);                   // [V][V] scratch
  bf16* dS = reinterpret_cast<bf16*>(dSf + VP * VP);              // [VP u][UP v]
  bf16* dSt = dS + VP * UP;                                       // [VP v][UP u]
  bf16* stage = dSt + VP * UP;                                    // per warp: theta tile, phi tile
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3, lj = lane >> 3, lr = lane & 7;
  const int n = blockIdx.x, i = blockIdx.y;
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smraw);
    const int words = VP * VP + (2 * VP * UP + kWarps * 2 * tile) / 2;
    for (int e = tid; e < words; e += kThreads) w[e] = 0u;
  }
  __syncthreads();
  const int64_t g0 = ((int64_t)n * 3 + i) * V * V;
  for (int e = tid; e < V * V; e += kThreads) {
    const float d = dM[g0 + e];
    dSf[e] = d;
    atomicAdd(dPA + i * V * V + e, d);
  }
  __syncthreads();
  const float inv = 1.0f / (float)(IC * T);
  for (int v = warp; v < V; v += kWarps) {   // column v: dot over u, then dS
    float part = 0.f;
    for (int u = lane; u < V; u += 32) part += P[g0 + u * V + v] * dSf[u * V + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    for (int u = lane; u < V; u += 32) {
      const bf16 val = __float2bfloat16_rn(P[g0 + u * V + v] * (dSf[u * V + v] - part) * inv);
      dS[u * UP + v] = val;
      dSt[v * UP + u] = val;
    }
  }
  __syncthreads();
  bf16* ths = stage + warp * 2 * tile;
  bf16* phs = ths + tile;
  const int pieces = icc / 8;
  for (int t = warp; t < T; t += kWarps) {
    const int64_t row0 = ((int64_t)n * T + t) * V;
    for (int c0 = 0; c0 < IC; c0 += icc) {
      __syncwarp();
      for (int e = lane; e < 2 * V * pieces; e += 32) {
        const int which = e / (V * pieces), r = (e / pieces) % V, q = e % pieces;
        cp_async16(smem_u32((which ? phs : ths) + r * pitch + q * 8), thph + (row0 + r) * ld + (which ? 3 + i : i) * IC + c0 + q * 8);
      }
      cp_async_wait_all();
      __syncwarp();
#pragma unroll
      for (int which = 0; which < 2; ++which) {   // 0: d(theta) = dS phi ; 1: d(phi) = dS^T theta
        const bf16* Aop = which ? dSt : dS;
        const bf16* Bop = which ? ths : phs;
        for (int nq = 0; nq < icc / 16; ++nq) {   // 16 output channels per pass
          float acc[MT][2][4];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[mt][j][0] = acc[mt][j][1] = acc[mt][j][2] = acc[mt][j][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t r4[4];
            ldsm_x4_t(smem_u32(Bop + (ks * 16 + (lj & 1) * 8 + lr) * pitch + nq * 16 + (lj >> 1) * 8), r4);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              uint32_t a[4];
              ldsm_x4(smem_u32(Aop + (mt * 16 + (lj & 1) * 8 + lr) * UP + ks * 16 + (lj >> 1) * 8), a);
              mma(acc[mt][0], a, r4[0], r4[1]);
              mma(acc[mt][1], a, r4[2], r4[3]);
            }
          }
          const int col = (which ? 3 + i : i) * IC + c0 + nq * 16 + 2 * tq;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (r0 < V) *reinterpret_cast<uint32_t*>(dthph + (row0 + r0) * ld + col + j * 8) = pack2(acc[mt][j][0], acc[mt][j][1]);
              if (r1 < V) *reinterpret_cast<uint32_t*>(dthph + (row0 + r1) * ld + col + j * 8) = pack2(acc[mt][j][2], acc[mt][j][3]);
            }
          }
        }
      }
    }
  }
}

}  // namespace
}  // namespace afb

using namespace afb;

// z: split == 0 -> [M, 3C] bf16; split == 1 -> [M, 9C] = (hi | lo | hi) slabs of 3C columns each (exact-mask forward)
extern "C" int afb_agcn_aggregate_fwd_mma(const void* x, const float* Mmat, void* z, int split, int N, int T, int V, int C, afb_stream s) {
  AFB_REQUIRE(x && Mmat && z && N > 0 && T > 0 && V > 0 && V <= 48 && C > 0 && C % 64 == 0, "agcn_aggregate_fwd_mma: bad args (V <= 48, C %% 64 == 0)");
  const int VP = V <= 32 ? 32 : 48, UP = VP + 8, per_cta = kWarps * kFwdFramesPerWarp, groups = (T + per_cta - 1) / per_cta;
  const size_t smem = ((size_t)(split ? 2 : 1) * 3 * VP * UP + (size_t)kWarps * VP * CP + (size_t)kWarps * (split ? 2 : 1) * 16 * CP) * 2;
  int rc;
#define LAUNCH(VP_, EX_)                                                                                              \
  do {                                                                                                                \
    if ((rc = set_smem(agcn_aggr_fwd_mma_kernel<VP_, EX_>, smem, "agcn_aggregate_fwd_mma"))) return rc;               \
    agcn_aggr_fwd_mma_kernel<VP_, EX_><<<N * groups, kThreads, smem, as_stream(s)>>>((const bf16*)x, Mmat, (bf16*)z, T, V, C, groups); \
  } while (0)
  if (VP == 32) { if (split) LAUNCH(32, true); else LAUNCH(32, false); }
  else { if (split) LAUNCH(48, true); else LAUNCH(48, false); }
#undef LAUNCH
  return check_launch("agcn_aggregate_fwd_mma");
}

extern "C" int afb_agcn_aggregate_bwd_mma(const void* x, const void* dz, const float* Mmat, void* dx, int accumulate, float* dM, int N,
                                          int T, int V, int C, afb_stream s) {
  AFB_REQUIRE(x && dz && Mmat && dx && dM && N > 0 && T > 0 && V > 0 && V <= 48 && C > 0 && C % 64 == 0,
              "agcn_aggregate_bwd_mma: bad args (V <= 48, C %% 64 == 0)");
  AFB_REQUIRE(x != dx && dz != dx, "agcn_aggregate_bwd_mma: dx must not alias an input");
  const int VP = V <= 32 ? 32 : 48, UP = VP + 8, groups = (T + kWarps - 1) / kWarps;
  int rc;
  {
    const size_t smem = ((size_t)3 * VP * UP + (size_t)kWarps * 3 * VP * CXP) * 2;
    if (VP == 32) {
      if ((rc = set_smem(agcn_aggr_bwd_dx_mma_kernel<32>, smem, "agcn_aggregate_bwd_dx"))) return rc;
      agcn_aggr_bwd_dx_mma_kernel<32><<<N * groups, kThreads, smem, as_stream(s)>>>((const bf16*)dz, Mmat, (bf16*)dx, accumulate, T, V, C, groups);
    } else {
      if ((rc = set_smem(agcn_aggr_bwd_dx_mma_kernel<48>, smem, "agcn_aggregate_bwd_dx"))) return rc;
      agcn_aggr_bwd_dx_mma_kernel<48><<<N * groups, kThreads, smem, as_stream(s)>>>((const bf16*)dz, Mmat, (bf16*)dx, accumulate, T, V, C, groups);
    }
    if ((rc = check_launch("agcn_aggregate_bwd_dx"))) return rc;
  }
  {
    int splits = 1;
    if (N < 296) { splits = (296 + N - 1) / N; const int mx = (T + FG - 1) / FG; if (splits > mx) splits = mx; if (splits < 1) splits = 1; }
    if (splits > 1) {
      cudaError_t e = cudaMemsetAsync(dM, 0, sizeof(float) * (size_t)N * 3 * V * V, as_stream(s));
      if (e != cudaSuccess) { set_error("agcn_aggregate_bwd_mma: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const size_t smem = (size_t)FG * 4 * VP * CP * 2;
    dim3 grid(N, splits);
    if (VP == 32) {
      if ((rc = set_smem(agcn_aggr_bwd_dm_mma_kernel<32>, smem, "agcn_aggregate_bwd_dm"))) return rc;
      agcn_aggr_bwd_dm_mma_kernel<32><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)x, (const bf16*)dz, dM, T, V, C, splits);
    } else {
      if ((rc = set_smem(agcn_aggr_bwd_dm_mma_kernel<48>, smem, "agcn_aggregate_bwd_dm"))) return rc;
      agcn_aggr_bwd_dm_mma_kernel<48><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)x, (const bf16*)dz, dM, T, V, C, splits);
    }
  }
  return check_launch("agcn_aggregate_bwd_dm");
}

// thph: [M, ld] bf16 (dtype AFB_BF16) or fp32 (AFB_F32: applied as bf16 hi + lo); V <= 48, IC % 16 == 0
extern "C" int afb_agcn_scores_fwd_mma(const void* thph, int dt, int ld, const float* A, const float* PA, float* P, float* Mmat, int N, int T,
                                       int V, int IC, afb_stream s) {
  AFB_REQUIRE(thph && A && PA && P && Mmat && N > 0 && T > 0 && V > 0 && V <= 48 && IC % 16 == 0 && ld >= 6 * IC && ld % 8 == 0,
              "agcn_scores_fwd_mma: bad args (V <= 48, IC %% 16 == 0)");
  AFB_REQUIRE(IC <= ICC || IC % ICC == 0, "agcn_scores_fwd_mma: IC=%d unsupported", IC);
  const int VP = V <= 32 ? 32 : 48, icc = IC < ICC ? IC : ICC, tile = VP * (icc + 8);
  const bool f32 = dt == AFB_F32;
  const size_t smem = (size_t)kWarps * (f32 ? 4 : 2) * tile * 2 + (size_t)VP * VP * 4;
  dim3 grid(N, 3);
  int rc;
#define LAUNCH(VP_, T_)                                                                                                  \
  do {                                                                                                                   \
    if ((rc = set_smem(agcn_scores_fwd_mma_kernel<VP_, T_>, smem, "agcn_scores_fwd_mma"))) return rc;                    \
    agcn_scores_fwd_mma_kernel<VP_, T_><<<grid, kThreads, smem, as_stream(s)>>>((const T_*)thph, ld, A, PA, P, Mmat, T, V, IC); \
  } while (0)
  if (VP == 32) { if (f32) LAUNCH(32, float); else LAUNCH(32, bf16); }
  else { if (f32) LAUNCH(48, float); else LAUNCH(48, bf16); }
#undef LAUNCH
  return check_launch("agcn_scores_fwd_mma");
}

extern "C" int afb_agcn_scores_bwd_mma(const void* thph, int ld, const float* P, const float* dM, float* dPA, void* dthph, int N, int T, int V,
                                       int IC, afb_stream s) {
  AFB_REQUIRE(thph && P && dM && dPA && dthph && N > 0 && T > 0 && V > 0 && V <= 48 && IC % 16 == 0 && ld >= 6 * IC && ld % 8 == 0,
              "agcn_scores_bwd_mma: bad args (V <= 48, IC %% 16 == 0)");
  AFB_REQUIRE(IC <= ICC || IC % ICC == 0, "agcn_scores_bwd_mma: IC=%d unsupported", IC);
  const int VP = V <= 32 ? 32 : 48, icc = IC < ICC ? IC : ICC, tile = VP * (icc + 8);
  const size_t smem = (size_t)VP * VP * 4 + ((size_t)2 * VP * (VP + 8) + (size_t)kWarps * 2 * tile) * 2;
  dim3 grid(N, 3);
  int rc;
  if (VP == 32) {
    if ((rc = set_smem(agcn_scores_bwd_mma_kernel<32>, smem, "agcn_scores_bwd_mma"))) return rc;
    agcn_scores_bwd_mma_kernel<32><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)thph, ld, P, dM, dPA, (bf16*)dthph, T, V, IC);
  } else {
    if ((rc = set_smem(agcn_scores_bwd_mma_kernel<48>, smem, "agcn_scores_bwd_mma"))) return rc;
    agcn_scores_bwd_mma_kernel<48><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)thph, ld, P, dM, dPA, (bf16*)dthph, T, V, IC);
  }
  return check_launch("agcn_scores_bwd_mma");
}
