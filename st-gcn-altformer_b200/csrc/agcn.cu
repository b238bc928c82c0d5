// General unit_agcn(C -> C_out), C % 8 == 0 (the TCN_GCN_unit stack, model/ST_TR/ST_TR_new.py:355-385;
// op sequence of model/unit_agcn.py:80-88).  The 1x1 convolutions (theta/phi, conv_d, down) run on the
// tcgen05 GEMM; this file holds the per-sample graph pieces around them:
//   scores_fwd    S_i = theta_i^T phi_i / (IC*T) -> P_i = softmax_u -> M_i = P_i + A_i + PA_i
//   aggregate_fwd z[n,t,v,(i,c)] = sum_u x[n,t,u,c] M_i[n,u,v]           (adjacency staged in smem)
//   aggregate_bwd dx += sum_{i,v} dz M_i ;  dM_i = sum_{t,c} x dz
//   scores_bwd    dPA, dS = P (dM - colsum(P dM)), d(theta), d(phi)
// Activations are channels-last token matrices; all math fp32 on CUDA cores.
#include "common.cuh"

namespace afb {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxPairs = 16;  // ceil(64*64 / 256)

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

// grid (N, 3)
template <typename T>
__global__ void __launch_bounds__(kThreads) agcn_scores_fwd_kernel(const T* __restrict__ thph, int ld, const float* __restrict__ A,
                                                                   const float* __restrict__ PA, float* __restrict__ P,
                                                                   float* __restrict__ Mmat, int Tn, int V, int IC, int TC) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, i = blockIdx.y, st = IC + 1;
  float* th = sm;                  // [TC*V][IC+1]
  float* ph = th + TC * V * st;    // [TC*V][IC+1]
  float* S = ph + TC * V * st;     // [V][V]
  float acc[kMaxPairs];
#pragma unroll
  for (int q = 0; q < kMaxPairs; ++q) acc[q] = 0.f;
  const T* base = thph + (int64_t)n * Tn * V * ld;
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int rows = min(TC, Tn - t0) * V;
    for (int e = threadIdx.x; e < rows * IC; e += blockDim.x) {
      const int r = e / IC, c = e % IC;
      const T* src = base + (int64_t)(t0 * V + r) * ld;
      th[r * st + c] = ldf<T>(src + i * IC + c);
      ph[r * st + c] = ldf<T>(src + (3 + i) * IC + c);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kMaxPairs; ++q) {
      const int pr = threadIdx.x + q * kThreads;
      if (pr < V * V) {
        const int u = pr / V, v = pr % V;
        float a = 0.f;
        for (int r = 0; r < rows; r += V) {
          const float* tu = th + (r + u) * st;
          const float* pv = ph + (r + v) * st;
          for (int c = 0; c < IC; ++c) a += tu[c] * pv[c];
        }
        acc[q] += a;
      }
    }
    __syncthreads();
  }
  const float inv = 1.0f / (float)(IC * Tn);
#pragma unroll
  for (int q = 0; q < kMaxPairs; ++q) {
    const int pr = threadIdx.x + q * kThreads;
    if (pr < V * V) S[pr] = acc[q] * inv;
  }
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float mx = -INFINITY;
    for (int u = 0; u < V; ++u) mx = fmaxf(mx, S[u * V + v]);
    float den = 0.f;
    for (int u = 0; u < V; ++u) den += __expf(S[u * V + v] - mx);
    const float rden = 1.0f / den;
    for (int u = 0; u < V; ++u) {
      const float pr = __expf(S[u * V + v] - mx) * rden;
      const int idx = (i * V + u) * V + v;
      const int64_t g = (int64_t)n * 3 * V * V + idx;
      P[g] = pr;
      Mmat[g] = pr + A[idx] + PA[idx];
    }
  }
}

// grid (N * chunks)
template <typename T>
__global__ void __launch_bounds__(kThreads) agcn_aggregate_fwd_kernel(const T* __restrict__ x, const float* __restrict__ Mmat,
                                                                      T* __restrict__ z, int Tn, int V, int C, int TC, int chunks) {
  extern __shared__ float sm[];
  const int n = blockIdx.x / chunks, t0 = (blockIdx.x % chunks) * TC;
  const int rows = min(TC, Tn - t0) * V;
  float* Ms = sm;               // [3VV]
  float* xs = Ms + 3 * V * V;   // [TC*V][C]
  const float* Mg = Mmat + (int64_t)n * 3 * V * V;
  for (int e = threadIdx.x; e < 3 * V * V; e += blockDim.x) Ms[e] = Mg[e];
  const int64_t row0 = ((int64_t)n * Tn + t0) * V;
  for (int e = threadIdx.x; e < rows * C; e += blockDim.x) xs[e] = ldf<T>(x + row0 * C + e);
  __syncthreads();
  for (int e = threadIdx.x; e < rows * C; e += blockDim.x) {
    const int r = e / C, c = e % C;
    const int tl = r / V, v = r % V;
    const float* xt = xs + tl * V * C + c;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int u = 0; u < V; ++u) {
      const float xv = xt[u * C];
      a0 += xv * Ms[(0 * V + u) * V + v];
      a1 += xv * Ms[(1 * V + u) * V + v];
      a2 += xv * Ms[(2 * V + u) * V + v];
    }
    T* dst = z + (row0 + r) * 3 * C + c;
    stf<T>(dst, a0);
    stf<T>(dst + C, a1);
    stf<T>(dst + 2 * C, a2);
  }
}

// grid (N): one CTA per sample loops over frame chunks
template <typename T>
__global__ void __launch_bounds__(kThreads) agcn_aggregate_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dz,
                                                                      const float* __restrict__ Mmat, T* __restrict__ dx,
                                                                      int accumulate, float* __restrict__ dM, int Tn, int V, int C,
                                                                      int TC) {
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int sx = C + 1, sz = 3 * C + 1;
  float* Ms = sm;                     // [3VV]
  float* xs = Ms + 3 * V * V;         // [TC*V][C+1]
  float* zs = xs + TC * V * sx;       // [TC*V][3C+1]
  const float* Mg = Mmat + (int64_t)n * 3 * V * V;
  for (int e = threadIdx.x; e < 3 * V * V; e += blockDim.x) Ms[e] = Mg[e];
  constexpr int kMaxItems = 3 * kMaxPairs;
  float acc[kMaxItems];
#pragma unroll
  for (int q = 0; q < kMaxItems; ++q) acc[q] = 0.f;
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int rows = min(TC, Tn - t0) * V;
    const int64_t row0 = ((int64_t)n * Tn + t0) * V;
    __syncthreads();
    for (int e = threadIdx.x; e < rows * C; e += blockDim.x) xs[(e / C) * sx + e % C] = ldf<T>(x + row0 * C + e);
    for (int e = threadIdx.x; e < rows * 3 * C; e += blockDim.x) zs[(e / (3 * C)) * sz + e % (3 * C)] = ldf<T>(dz + row0 * 3 * C + e);
    __syncthreads();
    // dx[t,u,c] = sum_i sum_v dz[t,v,(i,c)] * M_i[u,v]
    for (int e = threadIdx.x; e < rows * C; e += blockDim.x) {
      const int r = e / C, c = e % C;
      const int tl = r / V, u = r % V;
      float a = 0.f;
      for (int v = 0; v < V; ++v) {
        const float* zr = zs + (tl * V + v) * sz + c;
        a += zr[0] * Ms[(0 * V + u) * V + v] + zr[C] * Ms[(1 * V + u) * V + v] + zr[2 * C] * Ms[(2 * V + u) * V + v];
      }
      T* dst = dx + row0 * C + e;
      if (accumulate) a += ldf<T>(dst);
      stf<T>(dst, a);
    }
    // dM_i[u,v] += sum_{t,c} x[t,u,c] * dz[t,v,(i,c)]
#pragma unroll
    for (int q = 0; q < kMaxItems; ++q) {
      const int it = threadIdx.x + q * kThreads;
      if (it < 3 * V * V) {
        const int i = it / (V * V), u = (it / V) % V, v = it % V;
        float a = 0.f;
        for (int r = 0; r < rows; r += V) {
          const float* xr = xs + (r + u) * sx;
          const float* zr = zs + (r + v) * sz + i * C;
          for (int c = 0; c < C; ++c) a += xr[c] * zr[c];
        }
        acc[q] += a;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kMaxItems; ++q) {
    const int it = threadIdx.x + q * kThreads;
    if (it < 3 * V * V) dM[(int64_t)n * 3 * V * V + it] = acc[q];
  }
}

// grid (N, 3)
template <typename T>
__global__ void __launch_bounds__(kThreads) agcn_scores_bwd_kernel(const T* __restrict__ thph, int ld, const float* __restrict__ P,
                                                                   const float* __restrict__ dM, float* __restrict__ dPA,
                                                                   T* __restrict__ dthph, int Tn, int V, int IC, int TC) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, i = blockIdx.y, st = IC + 1;
  float* th = sm;
  float* ph = th + TC * V * st;
  float* dS = ph + TC * V * st;  // [V][V]
  const int64_t g0 = ((int64_t)n * 3 + i) * V * V;
  for (int e = threadIdx.x; e < V * V; e += blockDim.x) {
    const float d = dM[g0 + e];
    dS[e] = d;
    atomicAdd(dPA + i * V * V + e, d);
  }
  __syncthreads();
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float dot = 0.f;
    for (int u = 0; u < V; ++u) dot += P[g0 + u * V + v] * dS[u * V + v];
    for (int u = 0; u < V; ++u) dS[u * V + v] = P[g0 + u * V + v] * (dS[u * V + v] - dot);
  }
  __syncthreads();
  const float inv = 1.0f / (float)(IC * Tn);
  const T* base = thph + (int64_t)n * Tn * V * ld;
  T* obase = dthph + (int64_t)n * Tn * V * ld;
  for (int t0 = 0; t0 < Tn; t0 += TC) {
    const int rows = min(TC, Tn - t0) * V;
    __syncthreads();
    for (int e = threadIdx.x; e < rows * IC; e += blockDim.x) {
      const int r = e / IC, c = e % IC;
      const T* src = base + (int64_t)(t0 * V + r) * ld;
      th[r * st + c] = ldf<T>(src + i * IC + c);
      ph[r * st + c] = ldf<T>(src + (3 + i) * IC + c);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < rows * IC; e += blockDim.x) {
      const int r = e / IC, c = e % IC;
      const int tl = r / V, w = r % V;  // w plays u for d(theta) and v for d(phi)
      float a = 0.f, b = 0.f;
      for (int k = 0; k < V; ++k) {
        a += dS[w * V + k] * ph[(tl * V + k) * st + c];   // dth[t,u=w,c] = sum_v dS[u,v] ph[t,v,c]
        b += dS[k * V + w] * th[(tl * V + k) * st + c];   // dph[t,v=w,c] = sum_u dS[u,v] th[t,u,c]
      }
      T* dst = obase + (int64_t)(t0 * V + r) * ld;
      stf<T>(dst + i * IC + c, a * inv);
      stf<T>(dst + (3 + i) * IC + c, b * inv);
    }
  }
}

int pick_tc(int per_frame_bytes, int fixed_bytes, int Tn, int budget = 96 * 1024) {
  int tc = (budget - fixed_bytes) / per_frame_bytes;
  if (tc < 1) tc = 1;
  if (tc > Tn) tc = Tn;
  return tc;
}

}  // namespace
}  // namespace afb

using namespace afb;

extern "C" int afb_agcn_scores_fwd(const void* thph, int dt, int ld, const float* A, const float* PA, float* P, float* Mmat,
                                   int N, int T, int V, int IC, afb_stream s) {
  AFB_REQUIRE(thph && A && PA && P && Mmat && N > 0 && T > 0, "agcn_scores_fwd: bad args");
  AFB_REQUIRE(V <= 64 && ld >= 6 * IC, "agcn_scores_fwd: V=%d ld=%d unsupported", V, ld);
  const int per_frame = 2 * V * (IC + 1) * 4, fixed = V * V * 4;
  const int TC = pick_tc(per_frame, fixed, T);
  const size_t smem = (size_t)TC * per_frame + fixed;
  dim3 grid(N, 3);
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = set_smem(agcn_scores_fwd_kernel<bf16>, smem, "agcn_scores_fwd"))) return rc;
    agcn_scores_fwd_kernel<bf16><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)thph, ld, A, PA, P, Mmat, T, V, IC, TC);
  } else {
    if ((rc = set_smem(agcn_scores_fwd_kernel<float>, smem, "agcn_scores_fwd"))) return rc;
    agcn_scores_fwd_kernel<float><<<grid, kThreads, smem, as_stream(s)>>>((const float*)thph, ld, A, PA, P, Mmat, T, V, IC, TC);
  }
  return check_launch("agcn_scores_fwd");
}

extern "C" int afb_agcn_aggregate_fwd(const void* x, const float* Mmat, void* z, int dt, int N, int T, int V, int C, afb_stream s) {
  AFB_REQUIRE(x && Mmat && z && N > 0 && T > 0 && V <= 64, "agcn_aggregate_fwd: bad args");
  const int per_frame = V * C * 4, fixed = 3 * V * V * 4;
  const int TC = pick_tc(per_frame, fixed, T, 64 * 1024);
  const int chunks = ceil_div(T, TC);
  const size_t smem = (size_t)TC * per_frame + fixed;
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = set_smem(agcn_aggregate_fwd_kernel<bf16>, smem, "agcn_aggregate_fwd"))) return rc;
    agcn_aggregate_fwd_kernel<bf16><<<N * chunks, kThreads, smem, as_stream(s)>>>((const bf16*)x, Mmat, (bf16*)z, T, V, C, TC, chunks);
  } else {
    if ((rc = set_smem(agcn_aggregate_fwd_kernel<float>, smem, "agcn_aggregate_fwd"))) return rc;
    agcn_aggregate_fwd_kernel<float><<<N * chunks, kThreads, smem, as_stream(s)>>>((const float*)x, Mmat, (float*)z, T, V, C, TC, chunks);
  }
  return check_launch("agcn_aggregate_fwd");
}

extern "C" int afb_agcn_aggregate_bwd(const void* x, const void* dz, const float* Mmat, void* dx, int accumulate, float* dM, int dt,
                                      int N, int T, int V, int C, afb_stream s) {
  AFB_REQUIRE(x && dz && Mmat && dx && dM && N > 0 && T > 0 && V <= 64, "agcn_aggregate_bwd: bad args");
  const int per_frame = V * (4 * C + 2) * 4, fixed = 3 * V * V * 4;
  const int TC = pick_tc(per_frame, fixed, T, 160 * 1024);
  const size_t smem = (size_t)TC * per_frame + fixed;
  AFB_REQUIRE(smem <= 220 * 1024, "agcn_aggregate_bwd: V*C too large for shared memory");
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = set_smem(agcn_aggregate_bwd_kernel<bf16>, smem, "agcn_aggregate_bwd"))) return rc;
    agcn_aggregate_bwd_kernel<bf16><<<N, kThreads, smem, as_stream(s)>>>((const bf16*)x, (const bf16*)dz, Mmat, (bf16*)dx, accumulate, dM, T, V, C, TC);
  } else {
    if ((rc = set_smem(agcn_aggregate_bwd_kernel<float>, smem, "agcn_aggregate_bwd"))) return rc;
    agcn_aggregate_bwd_kernel<float><<<N, kThreads, smem, as_stream(s)>>>((const float*)x, (const float*)dz, Mmat, (float*)dx, accumulate, dM, T, V, C, TC);
  }
  return check_launch("agcn_aggregate_bwd");
}

extern "C" int afb_agcn_scores_bwd(const void* thph, int ld, const float* P, const float* dM, float* dPA, void* dthph, int dt, int N,
                                   int T, int V, int IC, afb_stream s) {
  AFB_REQUIRE(thph && P && dM && dPA && dthph && N > 0 && T > 0 && V <= 64, "agcn_scores_bwd: bad args");
  const int per_frame = 2 * V * (IC + 1) * 4, fixed = V * V * 4;
  const int TC = pick_tc(per_frame, fixed, T);
  const size_t smem = (size_t)TC * per_frame + fixed;
  dim3 grid(N, 3);
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = set_smem(agcn_scores_bwd_kernel<bf16>, smem, "agcn_scores_bwd"))) return rc;
    agcn_scores_bwd_kernel<bf16><<<grid, kThreads, smem, as_stream(s)>>>((const bf16*)thph, ld, P, dM, dPA, (bf16*)dthph, T, V, IC, TC);
  } else {
    if ((rc = set_smem(agcn_scores_bwd_kernel<float>, smem, "agcn_scores_bwd"))) return rc;
    agcn_scores_bwd_kernel<float><<<grid, kThreads, smem, as_stream(s)>>>((const float*)thph, ld, P, dM, dPA, (float*)dthph, T, V, IC, TC);
  }
  return check_launch("agcn_scores_bwd");
}
