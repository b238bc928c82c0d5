// gcn0 = unit_agcn(3 -> C_out), the bandwidth-bound adaptive graph convolution that opens the model
// (reference: model/unit_agcn.py:73-93, constructed at model/AltFormer/ST_GCN_AltFormer.py:43-48).
//
// With C_in = 3 everything before the 128-channel expansion is tiny, so the layer is restructured as
//   r[pos] = ( z_0, z_1, z_2, x ) in R^12,  z_i[n,t,v,:] = sum_u x[n,t,u,:] * M_i[n,u,v]
//   y[pos] = relu( BN(sum_i Wd_i z_i + b) + BN_down(Wdn x + bdn) ) = relu( Wfold * [r - E r ; 1] )
// and both BatchNorms' batch statistics follow from E[r] and Cov(r) (12 + 78 numbers) without ever
// materialising the (N,128,T,V) pre-activation.  Attention scores use the Gram form
//   S_i[u,v] = ( sum_ab (Wa_i^T Wb_i)[a,b] G[a,u,b,v] + ... ) / (IC*T),  G = sum_t x[t,u,a] x[t,v,b]
// which removes the 1024-long theta/phi contraction.  The (N,T,V,3) input is read directly; the
// 128-channel output is produced by one bf16 mma.sync k-step per 16 positions and written once.
//
//   gcn0_scores_kernel   per sample: M_i = softmax_u(S_i) + A_i + PA_i, per-sample moments of r
//   gcn0_finalize_kernel 1 CTA: fp64 reduce of moments -> E, Cov -> BN stats -> folded weights
//   gcn0_apply_kernel    per (sample, frame chunk): z -> A operand -> MMA -> ReLU -> coalesced store
// Backward (parameter gradients only): gcn0_bwd_q / fin1 / dz / fin2, see afb_gcn0_bwd.
#include <stdlib.h>

#include "common.cuh"
#include "mma_utils.cuh"

namespace afb {
namespace {

constexpr int NR = AFB_GCN0_NR;
constexpr int NMOM = AFB_GCN0_NMOM;
constexpr int NSTAT = AFB_GCN0_NSTAT_BASE;
constexpr int kThreads = 256;

__host__ __device__ constexpr int a4(int n) { return (n + 3) & ~3; }  // keep smem sections 16-byte aligned
__host__ __device__ constexpr int tri(int j, int k) { return NR + j * NR - (j * (j - 1)) / 2 + (k - j); }  // j <= k

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// relu(a), relu(b) -> bf16x2 (a in the low half), one instruction
__device__ __forceinline__ uint32_t relu_pack2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// four 8x8 b16 tiles from mma C-fragment registers to shared memory; lane l supplies the row address of tile l / 8, row l % 8
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

struct Coef {  // per subset: C (3x3), e (3), f (3)
  float v[3][16];
};

// one warp per coefficient, lanes over the inner channels (a serial 32-long chain of dependent global loads per
// thread used to cost ~3 us at the head of every CTA)
__device__ void compute_coef(const afb_gcn0_fwd_t& p, float (*coef)[16]) {
  const int lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int t = threadIdx.x >> 5; t < 45; t += nw) {
    const int i = t / 15, q = t % 15;
    const float* Wa = p.Wa[i];
    const float* Wb = p.Wb[i];
    float acc = 0.f;
    for (int c = lane; c < p.IC; c += 32) {
      if (q < 9) acc += Wa[c * 3 + q / 3] * Wb[c * 3 + q % 3];
      else if (q < 12) acc += Wa[c * 3 + (q - 9)] * p.bb[i][c];
      else acc += p.ba[i][c] * Wb[c * 3 + (q - 12)];
    }
    acc = warp_sum(acc);
    if (lane == 0) coef[i][q] = acc;
  }
}

// two adjacent elements (even index) as floats
template <typename T> __device__ __forceinline__ void ld2f(const T* p, float& a, float& b);
template <> __device__ __forceinline__ void ld2f<float>(const float* p, float& a, float& b) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  a = v.x; b = v.y;
}
template <> __device__ __forceinline__ void ld2f<bf16>(const bf16* p, float& a, float& b) {
  const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
  a = __uint_as_float(v << 16); b = __uint_as_float(v & 0xffff0000u);
}

// g[a][b] = sum_t x[t,u,a] x[t,v,b], su[a] = sum_t x[t,u,a], sv[b] = sum_t x[t,v,b]
__device__ __forceinline__ void gram_pair(const float* xs, int T, int V, int u, int v, float (&g)[9], float (&su)[3], float (&sv)[3]) {
#pragma unroll
  for (int q = 0; q < 9; ++q) g[q] = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) { su[a] = 0.f; sv[a] = 0.f; }
  for (int t = 0; t < T; ++t) {
    const float* xu = xs + (t * V + u) * 3;
    const float* xv = xs + (t * V + v) * 3;
    const float u0 = xu[0], u1 = xu[1], u2 = xu[2], v0 = xv[0], v1 = xv[1], v2 = xv[2];
    g[0] += u0 * v0; g[1] += u0 * v1; g[2] += u0 * v2;
    g[3] += u1 * v0; g[4] += u1 * v1; g[5] += u1 * v2;
    g[6] += u2 * v0; g[7] += u2 * v1; g[8] += u2 * v2;
    su[0] += u0; su[1] += u1; su[2] += u2;
    sv[0] += v0; sv[1] += v1; sv[2] += v2;
  }
}

// r[0..8] = z (subset-major, channel-minor), r[9..11] = x   for position (t, v) of the staged frames
__device__ __forceinline__ void position_r(const float* xs, const float* Ms, int V, int tl, int v, float (&r)[NR]) {
#pragma unroll
  for (int j = 0; j < 9; ++j) r[j] = 0.f;
  const float* xt = xs + tl * V * 3;
  for (int u = 0; u < V; ++u) {
    const float x0 = xt[u * 3], x1 = xt[u * 3 + 1], x2 = xt[u * 3 + 2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float m = Ms[(i * V + u) * V + v];
      r[i * 3 + 0] += x0 * m; r[i * 3 + 1] += x1 * m; r[i * 3 + 2] += x2 * m;
    }
  }
  r[9] = xt[v * 3]; r[10] = xt[v * 3 + 1]; r[11] = xt[v * 3 + 2];
}

// ---------------------------------------------------------------------------------------------
// forward 1: mixing matrices, moments of r, and (last CTA) statistics + folded weights
// ---------------------------------------------------------------------------------------------
constexpr int kSlots = AFB_GCN0_SLOTS;   // fp64 accumulation slots (spreads atomic contention)
constexpr int kPosChunk = 704;           // positions staged per moment pass (33 KB of r vectors)

// E[r], Cov(r) -> batch statistics of both BatchNorms -> BN-folded weights.  Runs in the last CTA to finish.
__device__ __noinline__ void gcn0_finalize(const afb_gcn0_fwd_t& p, double* dsm /* >= (96 + 12 + 144 + 4 * 96) doubles + Cout * 12 floats */) {
  double* tot = dsm;
  double* E = dsm + NMOM;
  double* Cov = E + NR;   // [NR][NR]
  const int tid = threadIdx.x;
  // slot sums with every thread holding at most a few loads in flight (a 32-long dependent load chain per moment
  // used to put ~5 us on the tail of the launch): thread (j, group) sums its group's slots, then the groups are added
  {
    double* partial = Cov + NR * NR;          // [kSlotGroups][NMOM] scratch behind Cov
    constexpr int kSlotGroups = 4;
    const int j = tid % NMOM, grp = tid / NMOM;
    if (grp < kSlotGroups) {
      double sacc = 0.0;
#pragma unroll
      for (int l = grp; l < kSlots; l += kSlotGroups) sacc += __ldcg(p.moments + l * NMOM + j);
      partial[grp * NMOM + j] = sacc;
    }
    __syncthreads();
    if (tid < NMOM) {
      double sacc = 0.0;
#pragma unroll
      for (int q = 0; q < kSlotGroups; ++q) sacc += partial[q * NMOM + tid];
      tot[tid] = sacc;
    }
  }
  __syncthreads();
  const double m = (double)p.N * p.T * p.V;
  if (tid < NR) E[tid] = tot[tid] / m;
  __syncthreads();
  for (int e = tid; e < NR * NR; e += blockDim.x) {
    const int j = e / NR, k = e % NR;
    const int a = j < k ? j : k, b = j < k ? k : j;
    Cov[e] = tot[tri(a, b)] / m - E[j] * E[k];
  }
  __syncthreads();
  if (tid < NR) p.stats[tid] = (float)E[tid];
  for (int e = tid; e < NR * NR; e += blockDim.x) p.stats[NR + e] = (float)Cov[e];
  // the 12 mixing weights of every output channel -> shared memory with all threads (the per-channel loop below used to
  // start with 12 + 3 dependent global loads and a dynamically indexed local array: 6-8 us on the tail of the launch)
  float* wsm = reinterpret_cast<float*>(Cov + NR * NR + 4 * NMOM);   // [Cout][12] behind the slot partials
  float* Ef = wsm + p.Cout * NR;                                     // [12] + [144] fp32 copies of E and Cov
  float* Covf = Ef + 16;
  if (tid < NR) Ef[tid] = (float)E[tid];
  for (int e = tid; e < NR * NR; e += blockDim.x) Covf[e] = (float)Cov[e];
  for (int e = tid; e < p.Cout * NR; e += blockDim.x) {
    const int o = e / NR, j = e % NR;
    wsm[e] = j < 9 ? p.Wd[j / 3][o * 3 + j % 3] : p.Wdn[o * 3 + (j - 9)];
  }
  __syncthreads();
  // 4 lanes per output channel split the rows of the two quadratic forms (fp64 chains are the critical path here)
  for (int o4 = tid; o4 < p.Cout * 4; o4 += blockDim.x) {
    const int o = o4 >> 2, part = o4 & 3;
    const float* w = wsm + o * NR;
    double b = 0.0;
    if (part == 0) b = (double)p.bd[0][o] + (double)p.bd[1][o] + (double)p.bd[2][o];
    // E and Cov were formed in fp64 (that is where the cancellation is); the per-channel contractions run in fp32 like
    // the reference's own BatchNorm statistics -- fp64 chains here were the serial tail of the launch
    float mean_hf = 0.f, mean_df = 0.f, var_hf = 0.f, var_df = 0.f;
    for (int j = part; j < 9; j += 4) {
      mean_hf = fmaf(w[j], Ef[j], mean_hf);
      float row = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) row = fmaf(Covf[j * NR + k], w[k], row);
      var_hf = fmaf(w[j], row, var_hf);
    }
    if (part < 3) {
      const int j = 9 + part;
      mean_df = w[j] * Ef[j];
      float row = 0.f;
#pragma unroll
      for (int k = 9; k < 12; ++k) row = fmaf(Covf[j * NR + k], w[k], row);
      var_df = w[j] * row;
    }
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {   // the 4 lanes of a channel are adjacent; blockDim and Cout*4 are multiples of 32
      mean_hf += __shfl_xor_sync(0xffffffffu, mean_hf, off);
      mean_df += __shfl_xor_sync(0xffffffffu, mean_df, off);
      var_hf += __shfl_xor_sync(0xffffffffu, var_hf, off);
      var_df += __shfl_xor_sync(0xffffffffu, var_df, off);
    }
    double mean_h = mean_hf, mean_d = mean_df, var_h = var_hf, var_d = var_df;
    if (part != 0) continue;
    mean_h += b;
    mean_d += p.bdn[o];
    if (var_h < 0.0) var_h = 0.0;
    if (var_d < 0.0) var_d = 0.0;
    const double ctr_h = mean_h, ctr_d = mean_d;  // pre-BN activations at the batch centre E[r]
    if (p.training) {
      const double unb = m > 1.0 ? m / (m - 1.0) : 1.0;
      p.bn_rm[o] = (float)((1.0 - p.momentum) * p.bn_rm[o] + p.momentum * mean_h);
      p.bn_rv[o] = (float)((1.0 - p.momentum) * p.bn_rv[o] + p.momentum * var_h * unb);
      p.dn_rm[o] = (float)((1.0 - p.momentum) * p.dn_rm[o] + p.momentum * mean_d);
      p.dn_rv[o] = (float)((1.0 - p.momentum) * p.dn_rv[o] + p.momentum * var_d * unb);
    } else {
      mean_h = p.bn_rm[o]; var_h = p.bn_rv[o];
      mean_d = p.dn_rm[o]; var_d = p.dn_rv[o];
    }
    const double rstd_h = 1.0 / sqrt(var_h + (double)p.eps), rstd_d = 1.0 / sqrt(var_d + (double)p.eps);
    const double sh = p.bn_g[o] * rstd_h, sd = p.dn_g[o] * rstd_d;
    float wf[16];   // folded weights stay in registers: stored as float4s, packed into the MMA fragments from here
#pragma unroll
    for (int j = 0; j < 9; ++j) wf[j] = (float)(sh * w[j]);
#pragma unroll
    for (int j = 9; j < 12; ++j) wf[j] = (float)(sd * w[j]);
    wf[12] = (float)(sh * (ctr_h - mean_h) + p.bn_b[o] + sd * (ctr_d - mean_d) + p.dn_b[o]);
    wf[13] = wf[14] = wf[15] = 0.f;
    float4* wfg = reinterpret_cast<float4*>(p.Wfold + o * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) wfg[q] = make_float4(wf[4 * q], wf[4 * q + 1], wf[4 * q + 2], wf[4 * q + 3]);
    if (p.Wfrag != nullptr) {   // mma.m16n8k16 B fragments of Wfold^T, one uint2 per (n-tile, lane)
      uint2* frag = reinterpret_cast<uint2*>(p.Wfrag) + ((o >> 3) * 32 + (o & 7) * 4);
#pragma unroll
      for (int t = 0; t < 4; ++t) frag[t] = make_uint2(pack_bf16(wf[2 * t], wf[2 * t + 1]), pack_bf16(wf[2 * t + 8], wf[2 * t + 9]));
    }
    p.stats[NSTAT + o] = (float)mean_h;
    p.stats[NSTAT + p.Cout + o] = (float)rstd_h;
    p.stats[NSTAT + 2 * p.Cout + o] = (float)mean_d;
    p.stats[NSTAT + 3 * p.Cout + o] = (float)rstd_d;
  }
  __syncthreads();
  for (int j = tid; j < kSlots * NMOM; j += blockDim.x) p.moments[j] = 0.0;  // re-arm for the next launch
  if (tid == 0) *p.counter = 0;
}

// One CTA of 512 threads per sample: the per-sample work (Gram matrix, 3 softmaxed mixing matrices, r vectors,
// 90 moments) is a chain of small dependent phases, so the kernel is latency-bound -- a wide CTA shortens every
// phase (only N CTAs exist, 256 for the benchmark batch, so wide CTAs also fill the SMs).
constexpr int kScoreThreads = 512;    // 2 CTAs per SM: the benchmark's 256 samples run as one wave
constexpr int kMomSegs = kScoreThreads / 10;   // position segments per 3 x 3 moment block (51 x 10 threads)

// (__grid_constant__: gcn0_finalize takes the parameter block by reference; without it every thread would first
// copy the 464-byte struct to local memory -- 120 MB of local stores per launch, the top stall in the r01 profile)
__global__ void __launch_bounds__(kScoreThreads, 2) gcn0_scores_kernel(const __grid_constant__ afb_gcn0_fwd_t p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T, V = p.V, n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  float* xs = sm;                        // [T*V*3]
  float* Ms = xs + a4(T * V * 3);        // [3][V][V]
  float* rs = Ms + a4(3 * V * V);        // [kPosChunk][12]   r vectors of one chunk of positions; then [segs][10][12] partial moments
  __shared__ float coef[3][16];
  __shared__ int is_last;
  const float* xg = p.x + (int64_t)n * T * V * 3;
  if ((T * V * 3) % 4 == 0) {            // sample base is 16-byte aligned then: one float4 per thread, all in flight
    const float4* x4 = reinterpret_cast<const float4*>(xg);
    for (int i = tid; i < T * V * 3 / 4; i += nthr) reinterpret_cast<float4*>(xs)[i] = x4[i];
  } else {
    for (int i = tid; i < T * V * 3; i += nthr) xs[i] = xg[i];
  }
  float* APs = rs + kPosChunk * NR;      // [3][V][V]  A + PA (same for every sample; staged while x arrives)
  for (int i = tid; i < 3 * V * V; i += nthr) APs[i] = p.A[i] + p.PA[i];
  compute_coef(p, coef);
  __syncthreads();
  const float inv = 1.0f / (float)(p.IC * T);
  // Gram form of the scores.  G[a,u,b,v] = G[b,v,a,u]: only the pairs u <= v are accumulated (two threads per pair, each
  // over half of the frames, combined by one shuffle); the (v, u) score uses the transposed 3x3 block and swapped sums.
  const int npairs = V * (V + 1) / 2;
  for (int it = tid; it < ((npairs * 2 + 31) & ~31); it += nthr) {
    const int pr = it >> 1, half = it & 1;
    const bool live = pr < npairs;
    int u = 0, v = 0;
    if (live) {   // pr = u*V - u(u-1)/2 + (v - u)
      u = (int)((2.f * V + 1.f - sqrtf((2.f * V + 1.f) * (2.f * V + 1.f) - 8.f * pr)) * 0.5f);
      while (u * V - (u * (u - 1)) / 2 > pr) --u;
      while ((u + 1) * V - ((u + 1) * u) / 2 <= pr) ++u;
      v = u + pr - (u * V - (u * (u - 1)) / 2);
    }
    float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, su[3] = {0.f, 0.f, 0.f}, sv[3] = {0.f, 0.f, 0.f};
    const int t0 = half ? (T + 1) / 2 : 0, t1 = half ? T : (T + 1) / 2;
#pragma unroll 4
    for (int t = t0; t < t1; ++t) {
      const float* xu = xs + (t * V + u) * 3;
      const float* xv = xs + (t * V + v) * 3;
      const float u0 = xu[0], u1 = xu[1], u2 = xu[2], v0 = xv[0], v1 = xv[1], v2 = xv[2];
      g[0] += u0 * v0; g[1] += u0 * v1; g[2] += u0 * v2;
      g[3] += u1 * v0; g[4] += u1 * v1; g[5] += u1 * v2;
      g[6] += u2 * v0; g[7] += u2 * v1; g[8] += u2 * v2;
      su[0] += u0; su[1] += u1; su[2] += u2;
      sv[0] += v0; sv[1] += v1; sv[2] += v2;
    }
    float sc[3], sct[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {   // the score is linear in (g, su, sv): each half contributes its partial
      float acc = 0.f, acct = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          acc += coef[i][a * 3 + c] * g[a * 3 + c];
          acct += coef[i][a * 3 + c] * g[c * 3 + a];
        }
        acc += coef[i][9 + a] * su[a] + coef[i][12 + a] * sv[a];
        acct += coef[i][9 + a] * sv[a] + coef[i][12 + a] * su[a];
      }
      sc[i] = acc + __shfl_xor_sync(0xffffffffu, acc, 1);
      sct[i] = acct + __shfl_xor_sync(0xffffffffu, acct, 1);
    }
    if (live && half == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        Ms[(i * V + u) * V + v] = sc[i] * inv;
        Ms[(i * V + v) * V + u] = sct[i] * inv;
      }
    }
  }
  __syncthreads();
  // softmax over u for fixed (i, v): one warp per column, lanes over u (V <= 64: two rows per lane); four columns per
  // warp step so four independent reduce / exp / reduce chains are in flight
  {
    const int nw = nthr >> 5, ncol = 3 * V;
    for (int c0 = warp; c0 < ncol; c0 += 4 * nw) {
      float s0[4], s1[4], mx[4], e0[4], e1[4], den[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = c0 + k * nw;
        const int i = col / V, v = col % V;
        s0[k] = (col < ncol && lane < V) ? Ms[(i * V + lane) * V + v] : -INFINITY;
        s1[k] = (col < ncol && lane + 32 < V) ? Ms[(i * V + lane + 32) * V + v] : -INFINITY;
        mx[k] = fmaxf(s0[k], s1[k]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], off));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e0[k] = s0[k] > -INFINITY ? __expf(s0[k] - mx[k]) : 0.f;
        e1[k] = s1[k] > -INFINITY ? __expf(s1[k] - mx[k]) : 0.f;
        den[k] = e0[k] + e1[k];
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) den[k] += __shfl_xor_sync(0xffffffffu, den[k], off);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = c0 + k * nw;
        if (col >= ncol) continue;
        const int i = col / V, v = col % V;
        const float rden = 1.0f / den[k];
        if (lane < V) {
          const int idx = (i * V + lane) * V + v;
          const float m = e0[k] * rden + APs[idx];
          Ms[idx] = m;
          p.Mmat[(int64_t)n * 3 * V * V + idx] = m;
        }
        if (lane + 32 < V) {
          const int idx = (i * V + lane + 32) * V + v;
          const float m = e1[k] * rden + APs[idx];
          Ms[idx] = m;
          p.Mmat[(int64_t)n * 3 * V * V + idx] = m;
        }
      }
    }
  }
  __syncthreads();
  if (p.Aop != nullptr) {   // operands of the tensor-core apply pass: A_i[v][u] = M_i[u][v] (bf16, zero padded) + colsum
    const int VP = p.V <= 16 ? 16 : (p.V <= 32 ? 32 : 48), AP = VP + 8;
    bf16* Ag = reinterpret_cast<bf16*>(p.Aop) + (int64_t)n * 3 * VP * AP;
    for (int e = tid; e < 3 * VP * AP; e += nthr) {
      const int i = e / (VP * AP), r = e % (VP * AP), v = r / AP, u = r % AP;
      Ag[e] = __float2bfloat16_rn((v < V && u < V) ? Ms[(i * V + u) * V + v] : 0.f);
    }
    for (int e = tid; e < 3 * VP; e += nthr) {
      const int i = e / VP, v = e % VP;
      float cs = 0.f;
      if (v < V)
        for (int u = 0; u < V; ++u) cs += Ms[(i * V + u) * V + v];
      p.colsum[(int64_t)n * 3 * VP + e] = cs;
    }
  }
  // moments: the 12 x 12 second-moment matrix in 3 x 3 blocks (10 blocks of the upper triangle).  Thread (block, seg)
  // accumulates its block's 9 products (+ the 3 first moments on diagonal blocks) over every kMomSegs-th position:
  // 6 loads per 9 FMAs instead of 2 per 1, no shuffles.
  const int mblk = tid % 10, seg = tid / 10;
  int bj = 0, bk = 0;
  {
    int rem = mblk;
    while (rem >= 4 - bj) { rem -= 4 - bj; ++bj; }
    bk = bj + rem;
  }
  float macc[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) macc[e] = 0.f;
  const int chunk_pos = (kPosChunk / V) * V;   // whole frames per chunk
  for (int c0 = 0; c0 < T * V; c0 += chunk_pos) {
    const int np = min(chunk_pos, T * V - c0);
    // r vectors of this chunk: thread = (frame, pair of adjacent joints) computes all nine z for both joints, so every
    // x value is loaded once for 18 FMAs (chunks hold whole frames: chunk_pos is a multiple of V)
    {
      const int vp_n = (V + 1) / 2;
      const int frames = np / V;
      for (int it = tid; it < frames * vp_n; it += nthr) {
        const int tl = it / vp_n, v0 = (it % vp_n) * 2;
        const bool two = v0 + 1 < V;
        const int pos0 = c0 + tl * V + v0;
        const float* xt = xs + (pos0 / V) * V * 3;
        float z[2][9];
#pragma unroll
        for (int j = 0; j < 9; ++j) { z[0][j] = 0.f; z[1][j] = 0.f; }
#pragma unroll 2
        for (int u = 0; u < V; ++u) {
          const float x0 = xt[u * 3], x1 = xt[u * 3 + 1], x2 = xt[u * 3 + 2];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float* mrow = Ms + (i * V + u) * V + v0;
            const float ma = mrow[0], mb = two ? mrow[1] : 0.f;
            z[0][i * 3] = fmaf(x0, ma, z[0][i * 3]); z[0][i * 3 + 1] = fmaf(x1, ma, z[0][i * 3 + 1]); z[0][i * 3 + 2] = fmaf(x2, ma, z[0][i * 3 + 2]);
            z[1][i * 3] = fmaf(x0, mb, z[1][i * 3]); z[1][i * 3 + 1] = fmaf(x1, mb, z[1][i * 3 + 1]); z[1][i * 3 + 2] = fmaf(x2, mb, z[1][i * 3 + 2]);
          }
        }
        float* dst = rs + (pos0 - c0) * NR;
#pragma unroll
        for (int j = 0; j < 9; ++j) dst[j] = z[0][j];
        dst[9] = xt[v0 * 3]; dst[10] = xt[v0 * 3 + 1]; dst[11] = xt[v0 * 3 + 2];
        if (two) {
#pragma unroll
          for (int j = 0; j < 9; ++j) dst[NR + j] = z[1][j];
          dst[NR + 9] = xt[v0 * 3 + 3]; dst[NR + 10] = xt[v0 * 3 + 4]; dst[NR + 11] = xt[v0 * 3 + 5];
        }
      }
    }
    __syncthreads();
    if (seg < kMomSegs) {
#pragma unroll 2
      for (int pl = seg; pl < np; pl += kMomSegs) {
        const float* r = rs + pl * NR;
        const float a0 = r[3 * bj], a1 = r[3 * bj + 1], a2 = r[3 * bj + 2];
        const float b0 = r[3 * bk], b1 = r[3 * bk + 1], b2 = r[3 * bk + 2];
        macc[0] = fmaf(a0, b0, macc[0]); macc[1] = fmaf(a0, b1, macc[1]); macc[2] = fmaf(a0, b2, macc[2]);
        macc[3] = fmaf(a1, b0, macc[3]); macc[4] = fmaf(a1, b1, macc[4]); macc[5] = fmaf(a1, b2, macc[5]);
        macc[6] = fmaf(a2, b0, macc[6]); macc[7] = fmaf(a2, b1, macc[7]); macc[8] = fmaf(a2, b2, macc[8]);
        macc[9] += a0; macc[10] += a1; macc[11] += a2;   // first moments (used from the diagonal blocks)
      }
    }
    __syncthreads();
  }
  // per-thread partials -> rs (free now): [seg][block][12]; then moment j gathers its element over the segments
  if (seg < kMomSegs) {
#pragma unroll
    for (int e = 0; e < 12; ++e) rs[(seg * 10 + mblk) * 12 + e] = macc[e];
  }
  __syncthreads();
  if (tid < 90) {
    int blk, elem;
    if (tid < NR) {   // first moment of component tid: diagonal block (tid / 3, tid / 3)
      const int d = tid / 3;
      blk = d * 4 - (d * (d - 1)) / 2;
      elem = 9 + tid % 3;
    } else {          // second moment (pa <= pb) in the tri() order of the moments buffer
      int rem = tid - NR, pa = 0;
      while (rem >= NR - pa) { rem -= NR - pa; ++pa; }
      const int pb = pa + rem, ja = pa / 3, jb = pb / 3;
      blk = ja * 4 - (ja * (ja - 1)) / 2 + (jb - ja);
      elem = (pa % 3) * 3 + (pb % 3);
    }
    double tot = 0.0;
    for (int sg = 0; sg < kMomSegs; ++sg) tot += (double)rs[(sg * 10 + blk) * 12 + elem];
    atomicAdd(p.moments + (blockIdx.x % kSlots) * NMOM + tid, tot);
  }
  // last CTA to finish turns the accumulated moments into statistics and folded weights
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(p.counter, 1) == (int)gridDim.x - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    gcn0_finalize(p, reinterpret_cast<double*>(sm));
  }
}

// ---------------------------------------------------------------------------------------------
// forward 3: apply
// ---------------------------------------------------------------------------------------------
constexpr int kARow = 24;  // bf16 elements per A-operand row (16 used): 48 B rows keep LDS.32 conflict-free

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f));
}

// COUT: output channels (compile time for the MMA path).  TT frames per CTA, P = TT*V positions.
template <bool MMA, typename TY, int COUT>
__global__ void __launch_bounds__(kThreads) gcn0_apply_kernel(const afb_gcn0_fwd_t p, int TT, int chunks) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int V = p.V, T = p.T;
  const int n = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TT;
  const int tt = min(TT, T - t0);
  const int P = tt * V, P16 = (TT * V + 15) / 16 * 16;
  const int Cout = MMA ? COUT : p.Cout;
  float* Ms = reinterpret_cast<float*>(smraw);        // [3*V*V]
  float* xs = Ms + a4(3 * V * V);                     // [TT*V*3]
  float* ctr = xs + a4(TT * V * 3);                   // [16]
  uint8_t* after = reinterpret_cast<uint8_t*>(ctr + 16);
  const float* Mg = p.Mmat + (int64_t)n * 3 * V * V;
  for (int i = threadIdx.x; i < 3 * V * V; i += blockDim.x) Ms[i] = Mg[i];
  const float* xg = p.x + ((int64_t)n * T + t0) * V * 3;
  for (int i = threadIdx.x; i < P * 3; i += blockDim.x) xs[i] = xg[i];
  if (threadIdx.x < 16) ctr[threadIdx.x] = threadIdx.x < NR ? p.stats[threadIdx.x] : 0.f;
  const int64_t row0 = ((int64_t)n * T + t0) * V;

  if constexpr (MMA) {
    bf16* Aop = reinterpret_cast<bf16*>(after);                        // [P16][kARow]
    bf16* tile = Aop + P16 * kARow;                                    // [P16][COUT + 8]
    constexpr int kTileRow = COUT + 8;
    // B fragments of all COUT/8 n-tiles live in registers for the whole kernel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
    uint32_t bfrag[COUT / 8][2];
#pragma unroll
    for (int nt = 0; nt < COUT / 8; ++nt) {
      const float* wf = p.Wfold + (nt * 8 + g) * 16;
      bfrag[nt][0] = pack_bf16(wf[2 * tq], wf[2 * tq + 1]);
      bfrag[nt][1] = pack_bf16(wf[2 * tq + 8], wf[2 * tq + 9]);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < P16 * 4; it += blockDim.x) {
      const int pos = it >> 2, i = it & 3;
      bf16* dst = Aop + pos * kARow + i * 4;  // slots: [0..8] z, [9..11] x, [12] one, [13..15] zero
      if (pos >= P) {
        if (i < 3) { dst = Aop + pos * kARow + i * 3; dst[0] = dst[1] = dst[2] = __float2bfloat16_rn(0.f); }
        else { dst = Aop + pos * kARow + 9; for (int q = 0; q < 7; ++q) dst[q] = __float2bfloat16_rn(0.f); }
        continue;
      }
      const int tl = pos / V, v = pos % V;
      const float* xt = xs + tl * V * 3;
      if (i < 3) {
        float z0 = 0.f, z1 = 0.f, z2 = 0.f;
        for (int u = 0; u < V; ++u) {
          const float m = Ms[(i * V + u) * V + v];
          z0 += xt[u * 3] * m; z1 += xt[u * 3 + 1] * m; z2 += xt[u * 3 + 2] * m;
        }
        dst = Aop + pos * kARow + i * 3;
        dst[0] = __float2bfloat16_rn(z0 - ctr[i * 3]);
        dst[1] = __float2bfloat16_rn(z1 - ctr[i * 3 + 1]);
        dst[2] = __float2bfloat16_rn(z2 - ctr[i * 3 + 2]);
      } else {
        dst = Aop + pos * kARow + 9;
        dst[0] = __float2bfloat16_rn(xt[v * 3] - ctr[9]);
        dst[1] = __float2bfloat16_rn(xt[v * 3 + 1] - ctr[10]);
        dst[2] = __float2bfloat16_rn(xt[v * 3 + 2] - ctr[11]);
        dst[3] = __float2bfloat16_rn(1.f);
        dst[4] = dst[5] = dst[6] = __float2bfloat16_rn(0.f);
      }
    }
    __syncthreads();
    for (int mt = warp; mt < P16 / 16; mt += kThreads / 32) {
      uint32_t a[4];
      const bf16* ar = Aop + (mt * 16 + g) * kARow;
      a[0] = *reinterpret_cast<const uint32_t*>(ar + 2 * tq);
      a[1] = *reinterpret_cast<const uint32_t*>(ar + 8 * kARow + 2 * tq);
      a[2] = *reinterpret_cast<const uint32_t*>(ar + 2 * tq + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(ar + 8 * kARow + 2 * tq + 8);
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt) {
        float d[4];
        mma_bf16_16816(d, a, bfrag[nt][0], bfrag[nt][1]);
        bf16* o0 = tile + (mt * 16 + g) * kTileRow + nt * 8 + 2 * tq;
        *reinterpret_cast<uint32_t*>(o0) = pack_bf16(fmaxf(d[0], 0.f), fmaxf(d[1], 0.f));
        *reinterpret_cast<uint32_t*>(o0 + 8 * kTileRow) = pack_bf16(fmaxf(d[2], 0.f), fmaxf(d[3], 0.f));
      }
    }
    __syncthreads();
    bf16* yg = reinterpret_cast<bf16*>(p.y) + row0 * COUT;
    for (int idx = threadIdx.x; idx < P * (COUT / 8); idx += blockDim.x) {
      const int r = idx / (COUT / 8), c8 = idx % (COUT / 8);
      *reinterpret_cast<uint4*>(yg + (int64_t)r * COUT + c8 * 8) = *reinterpret_cast<const uint4*>(tile + r * kTileRow + c8 * 8);
    }
  } else {
    float* Aop = reinterpret_cast<float*>(after);   // [TT*V][13]
    float* Wf = Aop + TT * V * 13;                  // [Cout][13]
    for (int i = threadIdx.x; i < Cout * 13; i += blockDim.x) Wf[i] = p.Wfold[(i / 13) * 16 + (i % 13)];
    __syncthreads();
    for (int pos = threadIdx.x; pos < P; pos += blockDim.x) {
      float r[NR];
      position_r(xs, Ms, V, pos / V, pos % V, r);
#pragma unroll
      for (int j = 0; j < NR; ++j) Aop[pos * 13 + j] = r[j] - ctr[j];
      Aop[pos * 13 + 12] = 1.f;
    }
    __syncthreads();
    TY* yg = reinterpret_cast<TY*>(p.y) + row0 * Cout;
    for (int idx = threadIdx.x; idx < P * Cout; idx += blockDim.x) {
      const int pos = idx / Cout, c = idx % Cout;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 13; ++j) acc += Aop[pos * 13 + j] * Wf[c * 13 + j];
      stf<TY>(yg + idx, fmaxf(acc, 0.f));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward 3 (bf16 performance path): both stages on the tensor cores, one warp per frame.
//   stage 1  z[v, slot] = c[slot, v] + sum_{i,u} M_i[u,v] * (x[u,a] - E x_a)        (slot = 3i + a)
//            A = [M_0^T | M_1^T | M_2^T] (bf16, per sample, staged once per CTA), B = block-sparse copy of
//            the centred frame; the fp32 accumulator starts at c = E[x_a] * colsum(M_i)[v] - E[z_slot], so
//            the result is already centred and bf16 rounding only touches deviations.
//   stage 2  y[v, :] = relu(Wfold * [z ; x - E x ; 1])  -- the stage-1 accumulators ARE the A fragments.
// The frame's 128-channel rows are staged in the warp's shared-memory tile and leave as 16-byte coalesced
// stores (a frame is V consecutive 256-byte rows of the output).
// ---------------------------------------------------------------------------------------------
template <int COUT, int VP>   // VP = V rounded up to 16
__global__ void __launch_bounds__(kThreads) gcn0_apply_mma_kernel(const afb_gcn0_fwd_t p, int TT, int chunks) {
  using namespace mmau;
  constexpr int AP = VP + 8;          // A-operand pitch (elements)
  constexpr int OP = COUT + 8;        // output staging pitch
  constexpr int MT = VP / 16;
  extern __shared__ __align__(16) uint8_t smraw[];
  const int V = p.V, T = p.T;
  const int n = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TT;
  const int tt = min(TT, T - t0);
  float* xs = reinterpret_cast<float*>(smraw);                 // [TT*V*3]
  float* cs = xs + a4(TT * V * 3);                             // [9][VP] accumulator initialisers
  float* colsum = cs + 9 * VP;                                 // [3][VP]
  float* ctr = colsum + 3 * VP;                                // [16]
  bf16* Asm = reinterpret_cast<bf16*>(ctr + 16);               // [3][VP][AP]  (built by gcn0_scores)
  bf16* stage = Asm + 3 * VP * AP;                             // [8 warps][VP][OP]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;

  {  // operand staging with cp.async: every 16-byte (A) / 4-byte (x) piece is in flight at once
    const uint4* Ag = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.Aop) + (int64_t)n * 3 * VP * AP);
    uint4* As4 = reinterpret_cast<uint4*>(Asm);
    for (int i = threadIdx.x; i < 3 * VP * AP / 8; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(As4 + i)), "l"(Ag + i) : "memory");
    const float* xg = p.x + ((int64_t)n * T + t0) * V * 3;
    for (int i = threadIdx.x; i < tt * V * 3; i += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(xs + i)), "l"(xg + i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (threadIdx.x < 16) ctr[threadIdx.x] = threadIdx.x < NR ? p.stats[threadIdx.x] : 0.f;
  if (threadIdx.x < 3 * VP) colsum[threadIdx.x] = p.colsum[(int64_t)n * 3 * VP + threadIdx.x];
  uint32_t bfrag[COUT / 8][2];
#pragma unroll
  for (int nt = 0; nt < COUT / 8; ++nt) {
    const uint2 f = reinterpret_cast<const uint2*>(p.Wfrag)[nt * 32 + lane];
    bfrag[nt][0] = f.x;
    bfrag[nt][1] = f.y;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * VP; e += blockDim.x) {
    const int slot = e / VP, v = e % VP;
    cs[e] = v < V ? ctr[9 + slot % 3] * colsum[(slot / 3) * VP + v] - ctr[slot] : 0.f;
  }
  __syncthreads();

  const int slot_i = g / 3, slot_a = g % 3;      // z slot owned by this lane's B column in n-tile 0
  bf16* mystage = stage + warp * VP * OP;
  for (int fl = warp; fl < tt; fl += kThreads / 32) {
    const float* xt = xs + fl * V * 3;
    auto xc = [&](int u, int a) { return u < V ? xt[u * 3 + a] - ctr[9 + a] : 0.f; };
    // B fragments of the centred frame: n-tile 0 (slots 0-7) and the single z slot (8) of n-tile 1
    uint32_t bx[MT][2], bx1[MT][2];
#pragma unroll
    for (int ku = 0; ku < MT; ++ku) {
      const int u0 = ku * 16 + 2 * t;
      bx[ku][0] = pack2(xc(u0, slot_a), xc(u0 + 1, slot_a));
      bx[ku][1] = pack2(xc(u0 + 8, slot_a), xc(u0 + 9, slot_a));
      bx1[ku][0] = g == 0 ? pack2(xc(u0, 2), xc(u0 + 1, 2)) : 0u;
      bx1[ku][1] = g == 0 ? pack2(xc(u0 + 8, 2), xc(u0 + 9, 2)) : 0u;
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int v0 = mt * 16 + g, v1 = v0 + 8;
      float z0[4], z1[4];   // n-tile 0 (slots 2t, 2t+1) and n-tile 1 (slots 8+2t, 9+2t) for rows v0 / v1
      z0[0] = cs[(2 * t) * VP + v0]; z0[1] = cs[(2 * t + 1) * VP + v0];
      z0[2] = cs[(2 * t) * VP + v1]; z0[3] = cs[(2 * t + 1) * VP + v1];
      z1[0] = t == 0 ? cs[8 * VP + v0] : 0.f; z1[1] = 0.f;
      z1[2] = t == 0 ? cs[8 * VP + v1] : 0.f; z1[3] = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int ku = 0; ku < MT; ++ku) {
          uint32_t a[4];
          ldsm_x4(smem_u32(Asm + (i * VP + mt * 16 + (lane & 15)) * AP + ku * 16 + (lane >> 4) * 8), a);
          const bool mine = slot_i == i;
          mma(z0, a, mine ? bx[ku][0] : 0u, mine ? bx[ku][1] : 0u);
          if (i == 2) mma(z1, a, bx1[ku][0], bx1[ku][1]);
        }
      // stage-2 A fragment: slots 0-7 from z0, slot 8 from z1, slots 9-11 = centred x, slot 12 = 1
      uint32_t a2[4];
      a2[0] = pack2(z0[0], z0[1]);
      a2[1] = pack2(z0[2], z0[3]);
      if (t == 0) {
        a2[2] = pack2(z1[0], xc(v0, 0));
        a2[3] = pack2(z1[2], xc(v1, 0));
      } else if (t == 1) {
        a2[2] = pack2(xc(v0, 1), xc(v0, 2));
        a2[3] = pack2(xc(v1, 1), xc(v1, 2));
      } else if (t == 2) {
        a2[2] = pack2(1.f, 0.f);
        a2[3] = a2[2];
      } else {
        a2[2] = 0u;
        a2[3] = 0u;
      }
      // two n-tiles per step: ReLU folded into the bf16 conversion (cvt.rn.relu), the four 8x8 tiles (rows g / g+8 of
      // both n-tiles) leave through one stmatrix instead of four 4-byte stores
      const uint32_t st_row = smem_u32(mystage + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * OP + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < COUT / 8; nt += 2) {
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
        mma(d0, a2, bfrag[nt][0], bfrag[nt][1]);
        mma(d1, a2, bfrag[nt + 1][0], bfrag[nt + 1][1]);
        stsm_x4(st_row + nt * 16, relu_pack2(d0[0], d0[1]), relu_pack2(d0[2], d0[3]), relu_pack2(d1[0], d1[1]), relu_pack2(d1[2], d1[3]));
      }
    }
    __syncwarp();
    bf16* yg = reinterpret_cast<bf16*>(p.y) + (((int64_t)n * T + t0 + fl) * V) * COUT;
    for (int idx = lane; idx < V * (COUT / 8); idx += 32) {
      const int r = idx / (COUT / 8), c8 = idx % (COUT / 8);
      *reinterpret_cast<uint4*>(yg + (int64_t)r * COUT + c8 * 8) = *reinterpret_cast<const uint4*>(mystage + r * OP + c8 * 8);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// workspace layout (floats): Q [Cout][16] | U [Cout][16] | cvec [16] | Kmat [9][9 -> 96] | gram [3][16]
__host__ __device__ inline int ws_Q(int) { return 0; }
__host__ __device__ inline int ws_U(int cout) { return 16 * cout; }
__host__ __device__ inline int ws_c(int cout) { return 32 * cout; }
__host__ __device__ inline int ws_K(int cout) { return 32 * cout + 16; }
__host__ __device__ inline int ws_gram(int cout) { return 32 * cout + 16 + 96; }

// pass 1: Q[o][j] = sum_pos g1[pos,o] * (r_j - E_j)  (j < 12),  Q[o][12] = sum_pos g1[pos,o]
template <typename TY>
__global__ void __launch_bounds__(kThreads) gcn0_bwd_q_kernel(const afb_gcn0_bwd_t b, int TT, int chunks) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const afb_gcn0_fwd_t& p = b.f;
  const int V = p.V, T = p.T, Cout = p.Cout;
  const int n = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TT;
  const int tt = min(TT, T - t0);
  const int P = tt * V;
  float* Ms = reinterpret_cast<float*>(smraw);
  float* xs = Ms + 3 * V * V;
  float* ctr = xs + TT * V * 3;
  float* Aop = ctr + 16;             // [TT*V][13]
  float* part = Aop + TT * V * 13;   // [parts][Cout][13]  (parts * Cout = 2 * kThreads)
  const float* Mg = p.Mmat + (int64_t)n * 3 * V * V;
  for (int i = threadIdx.x; i < 3 * V * V; i += blockDim.x) Ms[i] = Mg[i];
  const float* xg = p.x + ((int64_t)n * T + t0) * V * 3;
  for (int i = threadIdx.x; i < P * 3; i += blockDim.x) xs[i] = xg[i];
  if (threadIdx.x < 16) ctr[threadIdx.x] = threadIdx.x < NR ? p.stats[threadIdx.x] : 0.f;
  __syncthreads();
  for (int pos = threadIdx.x; pos < P; pos += blockDim.x) {
    float r[NR];
    position_r(xs, Ms, V, pos / V, pos % V, r);
#pragma unroll
    for (int j = 0; j < NR; ++j) Aop[pos * 13 + j] = r[j] - ctr[j];
    Aop[pos * 13 + 12] = 1.f;
  }
  __syncthreads();
  const int64_t row0 = ((int64_t)n * T + t0) * V;
  const TY* dy = reinterpret_cast<const TY*>(b.dy) + row0 * Cout;
  const TY* y = reinterpret_cast<const TY*>(p.y) + row0 * Cout;
  // thread = 2 adjacent channels x one slice of the positions; the y / dy loads of 4 positions are issued
  // before any of them is used (the old one-position-at-a-time loop with the dy load behind the ReLU test was
  // a chain of dependent global loads)
  const int half_c = Cout / 2;
  const int o2 = (threadIdx.x % half_c) * 2, part_id = threadIdx.x / half_c, nparts = kThreads / half_c;
  float acc0[13], acc1[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }
  if (part_id < nparts) {
    for (int pos0 = part_id; pos0 < P; pos0 += 4 * nparts) {
      float ya[4], yb[4], ga[4], gb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pos = pos0 + u * nparts;
        ya[u] = yb[u] = ga[u] = gb[u] = 0.f;
        if (pos < P) {
          ld2f<TY>(y + (int64_t)pos * Cout + o2, ya[u], yb[u]);
          ld2f<TY>(dy + (int64_t)pos * Cout + o2, ga[u], gb[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pos = pos0 + u * nparts;
        if (pos < P) {
          const float g0 = ya[u] > 0.f ? ga[u] : 0.f, g1 = yb[u] > 0.f ? gb[u] : 0.f;
#pragma unroll
          for (int j = 0; j < 13; ++j) {
            const float a = Aop[pos * 13 + j];
            acc0[j] = fmaf(g0, a, acc0[j]);
            acc1[j] = fmaf(g1, a, acc1[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      part[(part_id * Cout + o2) * 13 + j] = acc0[j];
      part[(part_id * Cout + o2 + 1) * 13 + j] = acc1[j];
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < Cout * 13; e += blockDim.x) {
    const int oo = e / 13, j = e % 13;
    float s = 0.f;
    for (int q = 0; q < nparts; ++q) s += part[(q * Cout + oo) * 13 + j];
    atomicAdd(b.ws + ws_Q(Cout) + oo * 16 + j, s);
  }
}

// finalize 1: BN / conv_d / down gradients, and U, c, K for the dz pass (single CTA, thread per channel)
__global__ void gcn0_bwd_fin1_kernel(const afb_gcn0_bwd_t b) {
  const afb_gcn0_fwd_t& p = b.f;
  const int Cout = p.Cout;
  __shared__ float E[NR];
  __shared__ float Cov[NR][NR];
  extern __shared__ float shm[];  // per channel: u[9], coefK, dbeta  -> [Cout][12]
  if (threadIdx.x < NR) E[threadIdx.x] = p.stats[threadIdx.x];
  for (int i = threadIdx.x; i < NR * NR; i += blockDim.x) Cov[i / NR][i % NR] = p.stats[NR + i];
  __syncthreads();
  const double m = (double)p.N * p.T * p.V;
  for (int o = threadIdx.x; o < Cout; o += blockDim.x) {
    float w[NR];
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 3; ++a) w[i * 3 + a] = p.Wd[i][o * 3 + a];
    for (int a = 0; a < 3; ++a) w[9 + a] = p.Wdn[o * 3 + a];
    const float rstd_h = p.stats[NSTAT + Cout + o], rstd_d = p.stats[NSTAT + 3 * Cout + o];
    const float ga = p.bn_g[o], gd = p.dn_g[o];
    const float* Q = b.ws + ws_Q(Cout) + o * 16;
    const float dbeta = Q[12];
    float dgam = 0.f, dgam_d = 0.f;
    for (int j = 0; j < 9; ++j) dgam += w[j] * Q[j];
    for (int j = 9; j < 12; ++j) dgam_d += w[j] * Q[j];
    if (!p.training) {
      // running-statistics BatchNorm (eval): h - rm = w.(r - E) + (ctr - rm) with ctr = w.E + b the pre-BN value at the
      // centre the Q sums were taken around (E = 0 after the fused forward, the batch mean after the two-kernel one);
      // mean and variance are constants, so no statistics terms: dW_j = g rstd sum g1 r_j, db = g rstd dbeta
      float ctr_h = p.bd[0][o] + p.bd[1][o] + p.bd[2][o], ctr_d = p.bdn[o];
      for (int j = 0; j < 9; ++j) ctr_h += w[j] * E[j];
      for (int j = 9; j < 12; ++j) ctr_d += w[j] * E[j];
      const float mean_h = p.stats[NSTAT + o], mean_d = p.stats[NSTAT + 2 * Cout + o];
      dgam = (dgam + (ctr_h - mean_h) * dbeta) * rstd_h;
      dgam_d = (dgam_d + (ctr_d - mean_d) * dbeta) * rstd_d;
      atomicAdd(b.dbn_g + o, dgam);
      atomicAdd(b.dbn_b + o, dbeta);
      atomicAdd(b.ddn_g + o, dgam_d);
      atomicAdd(b.ddn_b + o, dbeta);
      for (int j = 0; j < 9; ++j) atomicAdd(b.dWd[j / 3] + o * 3 + (j % 3), ga * rstd_h * (Q[j] + E[j] * dbeta));
      for (int j = 9; j < 12; ++j) atomicAdd(b.dWdn + o * 3 + (j - 9), gd * rstd_d * (Q[j] + E[j] * dbeta));
      for (int i = 0; i < 3; ++i) atomicAdd(b.dbd[i] + o, ga * rstd_h * dbeta);
      atomicAdd(b.dbdn + o, gd * rstd_d * dbeta);
      float* U = b.ws + ws_U(Cout) + o * 16;
      for (int j = 0; j < 9; ++j) {
        U[j] = ga * rstd_h * w[j];
        shm[o * 12 + j] = w[j];
      }
      for (int j = 9; j < 16; ++j) U[j] = 0.f;
      shm[o * 12 + 9] = 0.f;    // no K / c corrections: the statistics do not depend on the batch
      shm[o * 12 + 10] = 0.f;
      continue;
    }
    dgam *= rstd_h;
    dgam_d *= rstd_d;
    atomicAdd(b.dbn_g + o, dgam);
    atomicAdd(b.dbn_b + o, dbeta);
    atomicAdd(b.ddn_g + o, dgam_d);
    atomicAdd(b.ddn_b + o, dbeta);
    for (int j = 0; j < 9; ++j) {
      float cw = 0.f;
      for (int k = 0; k < 9; ++k) cw += Cov[j][k] * w[k];
      atomicAdd(b.dWd[j / 3] + o * 3 + (j % 3), ga * rstd_h * (Q[j] - dgam * rstd_h * cw));
    }
    for (int j = 9; j < 12; ++j) {
      float cw = 0.f;
      for (int k = 9; k < 12; ++k) cw += Cov[j][k] * w[k];
      atomicAdd(b.dWdn + o * 3 + (j - 9), gd * rstd_d * (Q[j] - dgam_d * rstd_d * cw));
    }
    // conv biases feeding a batch-stat BN have exactly zero gradient (dbd, dbdn += 0)
    float* U = b.ws + ws_U(Cout) + o * 16;
    for (int j = 0; j < 9; ++j) {
      U[j] = ga * rstd_h * w[j];
      shm[o * 12 + j] = w[j];
    }
    for (int j = 9; j < 16; ++j) U[j] = 0.f;
    shm[o * 12 + 9] = (float)(ga * rstd_h * rstd_h * dgam / m);   // K coefficient
    shm[o * 12 + 10] = (float)(ga * rstd_h * dbeta / m);          // c coefficient
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 + 81; e += blockDim.x) {
    float s = 0.f;
    if (e < 9) {
      for (int o = 0; o < Cout; ++o) s += shm[o * 12 + e] * shm[o * 12 + 10];
      b.ws[ws_c(Cout) + e] = s;
    } else {
      const int k = (e - 9) / 9, j = (e - 9) % 9;
      for (int o = 0; o < Cout; ++o) s += shm[o * 12 + k] * shm[o * 12 + 9] * shm[o * 12 + j];
      b.ws[ws_K(Cout) + k * 9 + j] = s;
    }
  }
}

// pass 2 (one CTA per sample): dz -> dM -> dPA, dS -> Gram-form gradients of theta/phi
template <typename TY, int CPL>
__global__ void __launch_bounds__(kThreads) gcn0_bwd_dz_kernel(const afb_gcn0_bwd_t b, int TCK) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const afb_gcn0_fwd_t& p = b.f;
  const int V = p.V, T = p.T, Cout = p.Cout, n = blockIdx.x;
  float* xs = reinterpret_cast<float*>(smraw);   // [T*V*3]   the whole sample (the Gram pass at the end needs every frame)
  float* Ms = xs + T * V * 3;                    // [3VV]
  float* dM = Ms + 3 * V * V;                    // [3VV]     accumulated over the frame chunks
  float* dz = dM + 3 * V * V;                    // [TCK*V][9] one chunk of TCK frames at a time (T = 180, V = 46 would need 300 KB)
  float* cK = dz + TCK * V * 9;                  // c[9], pad to 16, K[81], E[12]
  float* red = cK + 16 + 96 + 16;                // [8][48]
  const float* xg = p.x + (int64_t)n * T * V * 3;
  for (int i = threadIdx.x; i < T * V * 3; i += blockDim.x) xs[i] = xg[i];
  const float* Mg = p.Mmat + (int64_t)n * 3 * V * V;
  for (int i = threadIdx.x; i < 3 * V * V; i += blockDim.x) { Ms[i] = Mg[i]; dM[i] = 0.f; }
  for (int i = threadIdx.x; i < 16 + 96; i += blockDim.x) cK[i] = b.ws[ws_c(Cout) + i];
  if (threadIdx.x < 16) cK[16 + 96 + threadIdx.x] = threadIdx.x < NR ? p.stats[threadIdx.x] : 0.f;
  const float* cvec = cK;
  const float* Kmat = cK + 16;
  const float* E = cK + 16 + 96;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float U[CPL][9];
#pragma unroll
  for (int q = 0; q < CPL; ++q)
#pragma unroll
    for (int j = 0; j < 9; ++j) U[q][j] = b.ws[ws_U(Cout) + (lane * CPL + q) * 16 + j];
  __syncthreads();
  const TY* dy = reinterpret_cast<const TY*>(b.dy) + (int64_t)n * T * V * Cout;
  const TY* y = reinterpret_cast<const TY*>(p.y) + (int64_t)n * T * V * Cout;
  for (int t0 = 0; t0 < T; t0 += TCK) {
  const int tc = min(TCK, T - t0), npos = tc * V;
  // centred z of every position of the chunk, all threads; it is parked in dz[] and replaced by the finished dz below
  for (int it = threadIdx.x; it < npos * 3; it += blockDim.x) {
    const int pos = it / 3, i = it % 3, tl = pos / V, v = pos % V;
    const float* xt = xs + (t0 + tl) * V * 3;
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
    for (int u = 0; u < V; ++u) {
      const float m = Ms[(i * V + u) * V + v];
      z0 += xt[u * 3] * m; z1 += xt[u * 3 + 1] * m; z2 += xt[u * 3 + 2] * m;
    }
    dz[pos * 9 + i * 3] = z0 - E[i * 3]; dz[pos * 9 + i * 3 + 1] = z1 - E[i * 3 + 1]; dz[pos * 9 + i * 3 + 2] = z2 - E[i * 3 + 2];
  }
  __syncthreads();
  // dz[pos][j] = sum_c relu'(y) dy[pos][c] U[c][j] - c_j - sum_k zc_k K[k][j]; two positions per warp step so both
  // rows' loads are in flight together
  for (int pos0 = warp * 2; pos0 < npos; pos0 += 2 * (kThreads / 32)) {
    float part[2][9];
    float yv[2][CPL], gv[2][CPL];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t gpos = (int64_t)t0 * V + min(pos0 + h, npos - 1);
#pragma unroll
      for (int q = 0; q < CPL; q += 2) {
        ld2f<TY>(y + gpos * Cout + lane * CPL + q, yv[h][q], yv[h][q + 1]);
        ld2f<TY>(dy + gpos * Cout + lane * CPL + q, gv[h][q], gv[h][q + 1]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 9; ++j) part[h][j] = 0.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const float g1 = yv[h][q] > 0.f ? gv[h][q] : 0.f;
#pragma unroll
        for (int j = 0; j < 9; ++j) part[h][j] = fmaf(g1, U[q][j], part[h][j]);
      }
#pragma unroll
      for (int j = 0; j < 9; ++j) part[h][j] = warp_sum(part[h][j]);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pos = pos0 + h;
      if (pos < npos && lane < 9) {
        float corr = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) corr = fmaf(dz[pos * 9 + k], Kmat[k * 9 + lane], corr);
        float mine = part[h][0];
#pragma unroll
        for (int j = 1; j < 9; ++j) mine = lane == j ? part[h][j] : mine;
        __syncwarp(0x1ffu);   // all 9 lanes have read the parked z of this position before it is overwritten
        dz[pos * 9 + lane] = mine - cvec[lane] - corr;
      }
    }
  }
  __syncthreads();
  // dM_i[u][v] += sum_{t in chunk, a} x[t,u,a] * dz[(t,v)][3i+a]
  for (int e = threadIdx.x; e < 3 * V * V; e += blockDim.x) {
    const int i = e / (V * V), u = (e / V) % V, v = e % V;
    float s = 0.f;
    for (int t = 0; t < tc; ++t) {
      const float* xu = xs + ((t0 + t) * V + u) * 3;
      const float* d = dz + (t * V + v) * 9 + i * 3;
      s += xu[0] * d[0] + xu[1] * d[1] + xu[2] * d[2];
    }
    dM[e] += s;
  }
  __syncthreads();   // the next chunk overwrites dz
  }
  for (int e = threadIdx.x; e < 3 * V * V; e += blockDim.x) atomicAdd(b.dPA + e, dM[e]);
  __syncthreads();
  // dS = P * (dM - sum_u P*dM) per column (i, v);  P = M - (A + PA)
  for (int col = threadIdx.x; col < 3 * V; col += blockDim.x) {
    const int i = col / V, v = col % V;
    float dot = 0.f;
    for (int u = 0; u < V; ++u) {
      const int idx = (i * V + u) * V + v;
      const float pr = Ms[idx] - p.A[idx] - p.PA[idx];
      dot += pr * dM[idx];
    }
    for (int u = 0; u < V; ++u) {
      const int idx = (i * V + u) * V + v;
      const float pr = Ms[idx] - p.A[idx] - p.PA[idx];
      dM[idx] = pr * (dM[idx] - dot);
    }
  }
  __syncthreads();
  float acc[45];
#pragma unroll
  for (int q = 0; q < 45; ++q) acc[q] = 0.f;
  for (int pr = threadIdx.x; pr < V * V; pr += blockDim.x) {
    const int u = pr / V, v = pr % V;
    float g[9], su[3], sv[3];
    gram_pair(xs, T, V, u, v, g, su, sv);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float ds = dM[(i * V + u) * V + v];
#pragma unroll
      for (int q = 0; q < 9; ++q) acc[i * 15 + q] += ds * g[q];
#pragma unroll
      for (int a = 0; a < 3; ++a) { acc[i * 15 + 9 + a] += ds * su[a]; acc[i * 15 + 12 + a] += ds * sv[a]; }
    }
  }
#pragma unroll
  for (int q = 0; q < 45; ++q) {
    const float s = warp_sum(acc[q]);
    if (lane == 0) red[warp * 48 + q] = s;
  }
  __syncthreads();
  if (threadIdx.x < 45) {
    float s = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) s += red[w * 48 + threadIdx.x];
    const float inv = 1.0f / (float)(p.IC * T);
    atomicAdd(b.ws + ws_gram(Cout) + (threadIdx.x / 15) * 16 + threadIdx.x % 15, s * inv);
  }
}

// finalize 2: chain rule through C = Wa^T Wb, e = Wa^T bb, f = Wb^T ba
__global__ void gcn0_bwd_fin2_kernel(const afb_gcn0_bwd_t b) {
  const afb_gcn0_fwd_t& p = b.f;
  for (int e = threadIdx.x; e < 3 * p.IC; e += blockDim.x) {
    const int i = e / p.IC, c = e % p.IC;
    const float* G = b.ws + ws_gram(p.Cout) + i * 16;  // dC[9], de[3], df[3]
    const float* Wa = p.Wa[i] + c * 3;
    const float* Wb = p.Wb[i] + c * 3;
    float dbb = 0.f, dba = 0.f;
    for (int a = 0; a < 3; ++a) {
      float s = 0.f;
      for (int q = 0; q < 3; ++q) s += Wb[q] * G[a * 3 + q];
      atomicAdd(b.dWa[i] + c * 3 + a, s + p.bb[i][c] * G[9 + a]);
      dbb += Wa[a] * G[9 + a];
    }
    for (int q = 0; q < 3; ++q) {
      float s = 0.f;
      for (int a = 0; a < 3; ++a) s += Wa[a] * G[a * 3 + q];
      atomicAdd(b.dWb[i] + c * 3 + q, s + p.ba[i][c] * G[12 + q]);
      dba += Wb[q] * G[12 + q];
    }
    atomicAdd(b.dba[i] + c, dba);
    atomicAdd(b.dbb[i] + c, dbb);
  }
}

int pick_tt(int T, int V) {
  int tt = 192 / V;
  if (tt < 1) tt = 1;
  if (tt > T) tt = T;
  return tt;
}

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

int check_fwd_args(const afb_gcn0_fwd_t* p) {
  AFB_REQUIRE(p && p->x && p->A && p->PA && p->Mmat && p->moments && p->counter && p->stats && p->Wfold && p->y, "gcn0: null pointer");
  for (int i = 0; i < 3; ++i)
    AFB_REQUIRE(p->Wa[i] && p->ba[i] && p->Wb[i] && p->bb[i] && p->Wd[i] && p->bd[i], "gcn0: null weight pointer");
  AFB_REQUIRE(p->Wdn && p->bdn && p->bn_g && p->bn_b && p->dn_g && p->dn_b && p->bn_rm && p->bn_rv && p->dn_rm && p->dn_rv,
              "gcn0: null BN pointer");
  AFB_REQUIRE(p->N > 0 && p->T > 0 && p->V > 0 && p->V <= 64, "gcn0: bad shape N=%d T=%d V=%d (V<=64)", p->N, p->T, p->V);
  AFB_REQUIRE(p->Cout % 32 == 0 && p->Cout <= 256, "gcn0: Cout=%d unsupported", p->Cout);
  return 0;
}

}  // namespace
}  // namespace afb

namespace afb {
int gcn0_fused_launch(const afb_gcn0_fwd_t* p, cudaStream_t st);   // agcn0_fused.cu
size_t gcn0_fused_mop_bytes(int N, int V);
int gcn0_fused_stamps(unsigned long long* out, int ctas);
}

using namespace afb;

extern "C" int64_t afb_gcn0_aop_bytes(int N, int V) {
  const int VP = V <= 16 ? 16 : (V <= 32 ? 32 : 48);
  const size_t two_kernel = (size_t)N * 3 * VP * (VP + 8) * 2, fused = gcn0_fused_mop_bytes(N, V);
  return (int64_t)(two_kernel > fused ? two_kernel : fused);
}

extern "C" int afb_gcn0_fused_stamps(uint64_t* out, int ctas) {
  AFB_REQUIRE(out && ctas > 0, "gcn0_fused_stamps: bad args");
  return gcn0_fused_stamps(reinterpret_cast<unsigned long long*>(out), ctas);
}

extern "C" int afb_gcn0_fwd(const afb_gcn0_fwd_t* p, afb_stream s) {
  int rc = check_fwd_args(p);
  if (rc) return rc;
  cudaStream_t st = as_stream(s);
  rc = gcn0_fused_launch(p, st);   // one cooperative kernel (bf16 mode, shapes it covers); -1 = not covered
  if (rc >= 0) return rc;
  const int T = p->T, V = p->V;
  {
    size_t smem = ((size_t)a4(T * V * 3) + 2 * a4(3 * V * V) + (size_t)kPosChunk * NR) * sizeof(float);
    const size_t fin_bytes = 636 * sizeof(double) + ((size_t)p->Cout * NR + 16 + NR * NR) * sizeof(float) + 16;
    if (smem < fin_bytes) smem = fin_bytes;   // the finalize step reuses the buffer (636 doubles + Cout x 12 floats)
    AFB_REQUIRE(smem <= 220 * 1024, "gcn0: T*V too large for the per-sample shared-memory stage (%zu B)", smem);
    if ((rc = set_smem(gcn0_scores_kernel, smem, "gcn0_scores"))) return rc;
    gcn0_scores_kernel<<<p->N, kScoreThreads, smem, st>>>(*p);
    if ((rc = check_launch("gcn0_scores"))) return rc;
  }
  const int TT = pick_tt(T, V), chunks = ceil_div(T, TT);
  const int P16 = (TT * V + 15) / 16 * 16;
  const size_t head = ((size_t)a4(3 * V * V) + a4(TT * V * 3) + 16) * sizeof(float);
  const bool mma = p->precise == 0 && p->y_dtype == AFB_BF16 && p->Cout == 128;   // precise 2 (exact ReLU masks) on a shape the fused kernel
                                                                                 // does not cover: fp32 FMA apply, bf16 output
  static const bool old_apply = getenv("AFB_GCN0_APPLY_V1") != nullptr;
  if (mma && !old_apply && V <= 48 && p->Aop != nullptr && p->colsum != nullptr && p->Wfrag != nullptr) {
    const int TT2 = T < 8 ? T : 8, chunks2 = ceil_div(T, TT2);   // one frame per warp; 2 CTAs per SM overlap each other's setup
    const int VP = V <= 16 ? 16 : (V <= 32 ? 32 : 48);
    const size_t smem2 = ((size_t)a4(TT2 * V * 3) + 12 * VP + 16) * sizeof(float) + (size_t)3 * VP * (VP + 8) * 2 +
                         (size_t)(kThreads / 32) * VP * (128 + 8) * 2;
#define LAUNCH_APPLY_MMA(VP_)                                                                           \
  do {                                                                                                  \
    if ((rc = set_smem(gcn0_apply_mma_kernel<128, VP_>, smem2, "gcn0_apply_mma"))) return rc;           \
    gcn0_apply_mma_kernel<128, VP_><<<p->N * chunks2, kThreads, smem2, st>>>(*p, TT2, chunks2);         \
  } while (0)
    if (VP == 16) LAUNCH_APPLY_MMA(16); else if (VP == 32) LAUNCH_APPLY_MMA(32); else LAUNCH_APPLY_MMA(48);
#undef LAUNCH_APPLY_MMA
  } else if (mma) {
    const size_t smem = head + (size_t)P16 * kARow * 2 + (size_t)P16 * (128 + 8) * 2;
    if ((rc = set_smem(gcn0_apply_kernel<true, bf16, 128>, smem, "gcn0_apply"))) return rc;
    gcn0_apply_kernel<true, bf16, 128><<<p->N * chunks, kThreads, smem, st>>>(*p, TT, chunks);
  } else {
    const size_t smem = head + ((size_t)TT * V * 13 + (size_t)p->Cout * 13) * sizeof(float);
    if (p->y_dtype == AFB_BF16) {
      if ((rc = set_smem(gcn0_apply_kernel<false, bf16, 0>, smem, "gcn0_apply"))) return rc;
      gcn0_apply_kernel<false, bf16, 0><<<p->N * chunks, kThreads, smem, st>>>(*p, TT, chunks);
    } else {
      if ((rc = set_smem(gcn0_apply_kernel<false, float, 0>, smem, "gcn0_apply"))) return rc;
      gcn0_apply_kernel<false, float, 0><<<p->N * chunks, kThreads, smem, st>>>(*p, TT, chunks);
    }
  }
  return check_launch("gcn0_apply");
}

extern "C" int afb_gcn0_bwd(const afb_gcn0_bwd_t* b, afb_stream s) {
  AFB_REQUIRE(b && b->dy && b->ws && b->dPA && b->dWdn && b->dbdn && b->dbn_g && b->dbn_b && b->ddn_g && b->ddn_b, "gcn0_bwd: null pointer");
  for (int i = 0; i < 3; ++i)
    AFB_REQUIRE(b->dWa[i] && b->dba[i] && b->dWb[i] && b->dbb[i] && b->dWd[i] && b->dbd[i], "gcn0_bwd: null gradient pointer");
  const afb_gcn0_fwd_t* p = &b->f;
  int rc = check_fwd_args(p);
  if (rc) return rc;
  AFB_REQUIRE(kThreads % p->Cout == 0 || p->Cout == 256, "gcn0_bwd: Cout=%d unsupported", p->Cout);
  cudaStream_t st = as_stream(s);
  const int T = p->T, V = p->V, Cout = p->Cout;
  cudaError_t e = cudaMemsetAsync(b->ws, 0, sizeof(float) * AFB_GCN0_BWD_WS(Cout), st);
  if (e != cudaSuccess) { set_error("gcn0_bwd: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
  const int TT = pick_tt(T, V), chunks = ceil_div(T, TT);
  {
    const size_t smem = ((size_t)3 * V * V + TT * V * 3 + 16 + (size_t)TT * V * 13 + 2 * kThreads * 13) * sizeof(float);
    if (p->y_dtype == AFB_BF16) {
      if ((rc = set_smem(gcn0_bwd_q_kernel<bf16>, smem, "gcn0_bwd_q"))) return rc;
      gcn0_bwd_q_kernel<bf16><<<p->N * chunks, kThreads, smem, st>>>(*b, TT, chunks);
    } else {
      if ((rc = set_smem(gcn0_bwd_q_kernel<float>, smem, "gcn0_bwd_q"))) return rc;
      gcn0_bwd_q_kernel<float><<<p->N * chunks, kThreads, smem, st>>>(*b, TT, chunks);
    }
    if ((rc = check_launch("gcn0_bwd_q"))) return rc;
  }
  gcn0_bwd_fin1_kernel<<<1, 256, (size_t)Cout * 12 * sizeof(float), st>>>(*b);
  if ((rc = check_launch("gcn0_bwd_fin1"))) return rc;
  {
    // the sample's x stays whole in shared memory; dz is processed in chunks of TCK frames sized to the budget
    const size_t fixed = ((size_t)T * V * 3 + 6 * V * V + 16 + 96 + 16 + 8 * 48) * sizeof(float);
    AFB_REQUIRE(fixed + (size_t)V * 9 * sizeof(float) <= 200 * 1024, "gcn0_bwd: T*V too large for the per-sample shared-memory stage (%zu B)", fixed);
    int TCK = (int)((200 * 1024 - fixed) / ((size_t)V * 9 * sizeof(float)));
    if (TCK > T) TCK = T;
    const size_t smem = fixed + (size_t)TCK * V * 9 * sizeof(float);
#define LAUNCH_DZ(TYPE, CPL_)                                                              \
  do {                                                                                     \
    if ((rc = set_smem(gcn0_bwd_dz_kernel<TYPE, CPL_>, smem, "gcn0_bwd_dz"))) return rc;   \
    gcn0_bwd_dz_kernel<TYPE, CPL_><<<p->N, kThreads, smem, st>>>(*b, TCK);                 \
  } while (0)
    const int cpl = Cout / 32;
    AFB_REQUIRE(cpl == 2 || cpl == 4 || cpl == 8, "gcn0_bwd: Cout=%d unsupported (64, 128 or 256)", Cout);
    if (p->y_dtype == AFB_BF16) {
      if (cpl == 2) LAUNCH_DZ(bf16, 2); else if (cpl == 4) LAUNCH_DZ(bf16, 4); else LAUNCH_DZ(bf16, 8);
    } else {
      if (cpl == 2) LAUNCH_DZ(float, 2); else if (cpl == 4) LAUNCH_DZ(float, 4); else LAUNCH_DZ(float, 8);
    }
#undef LAUNCH_DZ
    if ((rc = check_launch("gcn0_bwd_dz"))) return rc;
  }
  gcn0_bwd_fin2_kernel<<<1, 128, 0, st>>>(*b);
  return check_launch("gcn0_bwd_fin2");
}
