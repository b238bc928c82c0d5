// Multi-head attention over short sequences (joints per frame or frames per joint; tuned for L <= 64, plain
// coverage kernels up to L = 256).
// Reference: model/AltFormer/model_ST.py:49-67.  One warp owns one (sequence, head): q/k/v slices
// live in that warp's shared memory, scores and probabilities never leave the SM (the reference
// materialises a [B, heads, L, L] tensor per block).  fp32 math on CUDA cores; the layout of qkv is the
// one nn.Linear(dim, 3*dim) + reshape(B, L, 3, heads, dh) implies: feature = s*D + h*dh + d.
#include "common.cuh"

#include <stdlib.h>

namespace afb {
// tensor-core path (attention_mma.cu)
bool attention_mma_supported(int L, int heads, int dh);
int attention_fwd_mma(const void* qkv, void* o, int64_t B, int L, int heads, int dh, float scale, const float* out_scale, cudaStream_t st);
bool attention_tc_supported(int L, int heads, int dh);
bool attention_bwd_tc_supported(int L, int heads, int dh);
int attention_bwd_tc(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, int dh, float scale, cudaStream_t st);
int attention_fwd_tc(const void* qkv, void* o, int64_t B, int L, int heads, int dh, float scale, const float* out_scale, int variant,
                     cudaStream_t st);
int attention_bwd_mma(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, int dh, float scale, cudaStream_t st);

namespace {

constexpr int kSmemBudget = 200 * 1024;

// bf16 goes to the mma.sync kernels; fp32 (parity mode), odd head sizes, or AFB_ATTN_SIMT=1 (cross-check)
// use the fp32 CUDA-core kernels below.
bool use_mma(int dt, int L, int heads, int dh) {
  static const bool force_simt = getenv("AFB_ATTN_SIMT") != nullptr;
  return dt == AFB_BF16 && !force_simt && attention_mma_supported(L, heads, dh);
}

template <typename T>
__device__ __forceinline__ void load_slice(float* dst, const T* src, int L, int dh, int ld, int lane) {
  // dst [L][dh+1] fp32; src points at (row 0, first feature of the head)
  for (int l = 0; l < L; ++l)
    for (int d = lane; d < dh; d += 32) dst[l * (dh + 1) + d] = ldf<T>(src + (int64_t)l * ld + d);
}

template <typename T>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o, int64_t items, int L, int heads,
                                                       int dh, float scale, int per_warp_floats, const float* __restrict__ out_scale) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= items) return;
  const int h = (int)(item % heads);
  const int64_t b = item / heads;
  const int D = heads * dh, ld = 3 * D, st = dh + 1;
  float* q = sm + (size_t)warp * per_warp_floats;
  float* k = q + L * st;
  float* v = k + L * st;
  float* prow = v + L * st;
  const T* base = qkv + (b * L) * ld + h * dh;
  load_slice<T>(q, base, L, dh, ld, lane);
  load_slice<T>(k, base + D, L, dh, ld, lane);
  load_slice<T>(v, base + 2 * D, L, dh, ld, lane);
  __syncwarp();
  const int j0 = lane, j1 = lane + 32;
  const float os = out_scale != nullptr ? out_scale[b] : 1.f;   // DropPath keep factor of this sequence
  for (int i = 0; i < L; ++i) {
    float s0 = -INFINITY, s1 = -INFINITY;
    if (j0 < L) {
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a += q[i * st + d] * k[j0 * st + d];
      s0 = a * scale;
    }
    if (j1 < L) {
      float a = 0.f;
      for (int d = 0; d < dh; ++d) a += q[i * st + d] * k[j1 * st + d];
      s1 = a * scale;
    }
    const float mx = warp_max(fmaxf(s0, s1));
    const float e0 = j0 < L ? __expf(s0 - mx) : 0.f, e1 = j1 < L ? __expf(s1 - mx) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    if (j0 < L) prow[j0] = e0 * inv;
    if (j1 < L) prow[j1] = e1 * inv;
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < L; ++j) a += prow[j] * v[j * st + d];
      stf<T>(o + (b * L + i) * D + h * dh + d, a * os);
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dO, T* __restrict__ dqkv,
                                                       int64_t items, int L, int heads, int dh, float scale, int per_warp_floats) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (item >= items) return;
  const int h = (int)(item % heads);
  const int64_t b = item / heads;
  const int D = heads * dh, ld = 3 * D, st = dh + 1, ps = L + 1;
  float* q = sm + (size_t)warp * per_warp_floats;
  float* k = q + L * st;
  float* v = k + L * st;
  float* go = v + L * st;
  float* P = go + L * st;   // [L][L+1]
  float* dS = P + L * ps;   // [L][L+1], already multiplied by `scale`
  const T* base = qkv + (b * L) * ld + h * dh;
  load_slice<T>(q, base, L, dh, ld, lane);
  load_slice<T>(k, base + D, L, dh, ld, lane);
  load_slice<T>(v, base + 2 * D, L, dh, ld, lane);
  load_slice<T>(go, dO + (b * L) * D + h * dh, L, dh, D, lane);
  __syncwarp();
  const int j0 = lane, j1 = lane + 32;
  for (int i = 0; i < L; ++i) {
    float s0 = -INFINITY, s1 = -INFINITY, g0 = 0.f, g1 = 0.f;
    if (j0 < L) {
      float a = 0.f, c = 0.f;
      for (int d = 0; d < dh; ++d) { a += q[i * st + d] * k[j0 * st + d]; c += go[i * st + d] * v[j0 * st + d]; }
      s0 = a * scale; g0 = c;
    }
    if (j1 < L) {
      float a = 0.f, c = 0.f;
      for (int d = 0; d < dh; ++d) { a += q[i * st + d] * k[j1 * st + d]; c += go[i * st + d] * v[j1 * st + d]; }
      s1 = a * scale; g1 = c;
    }
    const float mx = warp_max(fmaxf(s0, s1));
    const float e0 = j0 < L ? __expf(s0 - mx) : 0.f, e1 = j1 < L ? __expf(s1 - mx) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    const float delta = warp_sum(p0 * g0 + p1 * g1);
    if (j0 < L) { P[i * ps + j0] = p0; dS[i * ps + j0] = scale * p0 * (g0 - delta); }
    if (j1 < L) { P[i * ps + j1] = p1; dS[i * ps + j1] = scale * p1 * (g1 - delta); }
  }
  __syncwarp();
  T* obase = dqkv + (b * L) * ld + h * dh;
  for (int d = lane; d < dh; d += 32) {
    for (int i = 0; i < L; ++i) {  // dQ[i][d] = sum_j dS[i][j] k[j][d]
      float a = 0.f;
      for (int j = 0; j < L; ++j) a += dS[i * ps + j] * k[j * st + d];
      stf<T>(obase + (int64_t)i * ld + d, a);
    }
    for (int j = 0; j < L; ++j) {  // dK[j][d] = sum_i dS[i][j] q[i][d];  dV[j][d] = sum_i P[i][j] dO[i][d]
      float a = 0.f, c = 0.f;
      for (int i = 0; i < L; ++i) { a += dS[i * ps + j] * q[i * st + d]; c += P[i * ps + j] * go[i * st + d]; }
      stf<T>(obase + D + (int64_t)j * ld + d, a);
      stf<T>(obase + 2 * D + (int64_t)j * ld + d, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Long sequences (64 < L <= 256; the reference's default num_frame = 180): one CTA of 4 warps per (sequence, head),
// the head's q / k / v (and dO) slices in shared memory as fp32, flash-style recomputation in backward so nothing of
// size L x L is ever stored.  CUDA-core fp32 math: this path is for coverage, the tuned kernels handle L <= 64.
// ---------------------------------------------------------------------------------------------
constexpr int kLongWarps = 4;
constexpr int kLongMaxL = 256;

template <typename T>
__device__ __forceinline__ void load_slice_cta(float* dst, const T* src, int L, int dh, int ld) {
  for (int e = threadIdx.x; e < L * dh; e += blockDim.x) {
    const int l = e / dh, d = e % dh;
    dst[l * (dh + 1) + d] = ldf<T>(src + (int64_t)l * ld + d);
  }
}

template <typename T>
__global__ void __launch_bounds__(kLongWarps * 32) attn_long_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ o, int L, int heads, int dh,
                                                                        float scale, const float* __restrict__ out_scale) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % heads;
  const int64_t b = blockIdx.x / heads;
  const int D = heads * dh, ld = 3 * D, st = dh + 1;
  float* ks = sm;
  float* vs = ks + L * st;
  float* qs = vs + L * st;                 // [warps][dh]
  float* prow = qs + kLongWarps * dh;      // [warps][L]
  const T* base = qkv + (b * L) * ld + h * dh;
  load_slice_cta<T>(ks, base + D, L, dh, ld);
  load_slice_cta<T>(vs, base + 2 * D, L, dh, ld);
  __syncthreads();
  const float os = out_scale != nullptr ? out_scale[b] : 1.f;
  float* q = qs + warp * dh;
  float* pr = prow + warp * L;
  for (int i = warp; i < L; i += kLongWarps) {
    for (int d = lane; d < dh; d += 32) q[d] = ldf<T>(base + (int64_t)i * ld + d);
    __syncwarp();
    float sc[kLongMaxL / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY;
      if (j < L) {
        a = 0.f;
        for (int d = 0; d < dh; ++d) a = fmaf(q[d], ks[j * st + d], a);
        a *= scale;
      }
      sc[jj] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = j < L ? __expf(sc[jj] - mx) : 0.f;
      sum += sc[jj];
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      if (j < L) pr[j] = sc[jj] * inv;
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {
      float a = 0.f;
      for (int j = 0; j < L; ++j) a = fmaf(pr[j], vs[j * st + d], a);
      stf<T>(o + (b * L + i) * D + h * dh + d, a * os);
    }
    __syncwarp();
  }
}

// Backward: phase 1 (warp per query row) -> dQ, row statistics; phase 2 (warp per key row) recomputes the column of
// P / dS from the statistics -> dK, dV.  No accumulators, no atomics.
template <typename T>
__global__ void __launch_bounds__(kLongWarps * 32) attn_long_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dO, T* __restrict__ dqkv,
                                                                        int L, int heads, int dh, float scale) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % heads;
  const int64_t b = blockIdx.x / heads;
  const int D = heads * dh, ld = 3 * D, st = dh + 1;
  float* qs = sm;
  float* ks = qs + L * st;
  float* vs = ks + L * st;
  float* gs = vs + L * st;                 // dO
  float* lse = gs + L * st;                // [L]  row max + log(row sum) of the scaled scores
  float* dlt = lse + L;                    // [L]  sum_j P_ij dP_ij
  float* wbuf = dlt + L;                   // [warps][2][L]  per-warp row / column of dS (and P)
  const T* base = qkv + (b * L) * ld + h * dh;
  load_slice_cta<T>(qs, base, L, dh, ld);
  load_slice_cta<T>(ks, base + D, L, dh, ld);
  load_slice_cta<T>(vs, base + 2 * D, L, dh, ld);
  load_slice_cta<T>(gs, dO + (b * L) * D + h * dh, L, dh, D);
  __syncthreads();
  float* w0 = wbuf + warp * 2 * L;
  float* w1 = w0 + L;
  T* obase = dqkv + (b * L) * ld + h * dh;
  // ---- phase 1: rows ----
  for (int i = warp; i < L; i += kLongWarps) {
    float sc[kLongMaxL / 32], dp[kLongMaxL / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      float a = -INFINITY, c = 0.f;
      if (j < L) {
        a = 0.f;
        for (int d = 0; d < dh; ++d) {
          a = fmaf(qs[i * st + d], ks[j * st + d], a);
          c = fmaf(gs[i * st + d], vs[j * st + d], c);
        }
        a *= scale;
      }
      sc[jj] = a;
      dp[jj] = c;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = j < L ? __expf(sc[jj] - mx) : 0.f;
      sum += sc[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float delta = 0.f;
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      sc[jj] *= inv;
      delta = fmaf(sc[jj], dp[jj], delta);
    }
    delta = warp_sum(delta);
    if (lane == 0) {
      lse[i] = mx + __logf(sum);
      dlt[i] = delta;
    }
#pragma unroll
    for (int jj = 0; jj < kLongMaxL / 32; ++jj) {
      const int j = lane + 32 * jj;
      if (j < L) w0[j] = scale * sc[jj] * (dp[jj] - delta);   // dS_ij (already times scale)
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {   // dQ_i = sum_j dS_ij K_j
      float a = 0.f;
      for (int j = 0; j < L; ++j) a = fmaf(w0[j], ks[j * st + d], a);
      stf<T>(obase + (int64_t)i * ld + d, a);
    }
    __syncwarp();
  }
  __syncthreads();
  // ---- phase 2: columns ----
  for (int j = warp; j < L; j += kLongWarps) {
#pragma unroll
    for (int ii = 0; ii < kLongMaxL / 32; ++ii) {
      const int i = lane + 32 * ii;
      if (i < L) {
        float a = 0.f, c = 0.f;
        for (int d = 0; d < dh; ++d) {
          a = fmaf(qs[i * st + d], ks[j * st + d], a);
          c = fmaf(gs[i * st + d], vs[j * st + d], c);
        }
        const float pij = __expf(a * scale - lse[i]);
        w0[i] = pij;                                   // P_ij
        w1[i] = scale * pij * (c - dlt[i]);            // dS_ij
      }
    }
    __syncwarp();
    for (int d = lane; d < dh; d += 32) {   // dK_j = sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i
      float a = 0.f, c = 0.f;
      for (int i = 0; i < L; ++i) {
        a = fmaf(w1[i], qs[i * st + d], a);
        c = fmaf(w0[i], gs[i * st + d], c);
      }
      stf<T>(obase + D + (int64_t)j * ld + d, a);
      stf<T>(obase + 2 * D + (int64_t)j * ld + d, c);
    }
    __syncwarp();
  }
}

template <typename K>
int configure_smem(K kernel, int bytes) {
  if (bytes <= 48 * 1024) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("attention: cudaFuncSetAttribute(%d) failed: %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace
}  // namespace afb

using namespace afb;

extern "C" int afb_attention_fwd(const void* qkv, void* o, int dt, int64_t B, int L, int heads, int dh, float scale,
                                 const float* out_scale, afb_stream s) {
  AFB_REQUIRE(qkv && o && B > 0, "attention_fwd: bad args");
  AFB_REQUIRE(L >= 1 && L <= kLongMaxL && dh >= 1 && dh <= 64, "attention: L=%d dh=%d unsupported (L<=256, dh<=64)", L, dh);
  if (L > 64) {
    const int smem = (2 * L * (dh + 1) + kLongWarps * (dh + L)) * 4;
    AFB_REQUIRE(smem <= 220 * 1024, "attention: L=%d dh=%d does not fit shared memory", L, dh);
    AFB_REQUIRE(B * heads < (1ll << 31), "attention: too many (sequence, head) items");
    int rc;
    if (dt == AFB_BF16) {
      if ((rc = configure_smem(attn_long_fwd_kernel<bf16>, smem))) return rc;
      attn_long_fwd_kernel<bf16><<<(unsigned)(B * heads), kLongWarps * 32, smem, as_stream(s)>>>((const bf16*)qkv, (bf16*)o, L, heads, dh, scale, out_scale);
    } else {
      if ((rc = configure_smem(attn_long_fwd_kernel<float>, smem))) return rc;
      attn_long_fwd_kernel<float><<<(unsigned)(B * heads), kLongWarps * 32, smem, as_stream(s)>>>((const float*)qkv, (float*)o, L, heads, dh, scale, out_scale);
    }
    return check_launch("attention_long_fwd");
  }
  // Forward on the Blackwell paths (attention_tc.cu: TMA, tcgen05.mma, scores / probabilities / output in tensor memory)
  // wherever it is the faster kernel: dh = 32 (B200, B*L = 180k tokens: 71.5 vs 79.8 us at L = 22, 65 vs 100 us at L = 64);
  // at dh = 64 the two tie at the HBM floor and the warp-level kernel stays.  AFB_ATTN_TC (read per call so one process
  // can compare the paths): 0 = warp-level MMA kernel, 1 = tcgen05 kernel for every shape it covers, 2 = its variant with
  // the probabilities staged in shared memory instead of tensor memory.
  if (dt == AFB_BF16 && attention_tc_supported(L, heads, dh)) {
    const char* tc = getenv("AFB_ATTN_TC");
    const char mode = tc != nullptr ? tc[0] : '\0';
    if (mode == '1' || mode == '2' || (mode != '0' && dh == 32))
      return attention_fwd_tc(qkv, o, B, L, heads, dh, scale, out_scale, mode == '2' ? 1 : 0, as_stream(s));
  }
  if (use_mma(dt, L, heads, dh)) return attention_fwd_mma(qkv, o, B, L, heads, dh, scale, out_scale, as_stream(s));
  const int per_warp = 3 * L * (dh + 1) + L + 3;
  int warps = kSmemBudget / (per_warp * 4);
  if (warps > 8) warps = 8;
  AFB_REQUIRE(warps >= 1, "attention: tile does not fit shared memory");
  const int64_t items = B * heads;
  const int smem = warps * per_warp * 4;
  const int grid = ceil_div(items, warps);
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = configure_smem(attn_fwd_kernel<bf16>, smem))) return rc;
    attn_fwd_kernel<bf16><<<grid, warps * 32, smem, as_stream(s)>>>((const bf16*)qkv, (bf16*)o, items, L, heads, dh, scale, per_warp,
                                                                     out_scale);
  } else {
    if ((rc = configure_smem(attn_fwd_kernel<float>, smem))) return rc;
    attn_fwd_kernel<float><<<grid, warps * 32, smem, as_stream(s)>>>((const float*)qkv, (float*)o, items, L, heads, dh, scale, per_warp,
                                                                       out_scale);
  }
  return check_launch("attention_fwd");
}

extern "C" int afb_attention_bwd(const void* qkv, const void* dO, void* dqkv, int dt, int64_t B, int L, int heads, int dh,
                                 float scale, afb_stream s) {
  AFB_REQUIRE(qkv && dO && dqkv && B > 0, "attention_bwd: bad args");
  AFB_REQUIRE(L >= 1 && L <= kLongMaxL && dh >= 1 && dh <= 64, "attention: L=%d dh=%d unsupported (L<=256, dh<=64)", L, dh);
  if (L > 64) {
    const int smem = (4 * L * (dh + 1) + 2 * L + kLongWarps * 2 * L) * 4;
    AFB_REQUIRE(smem <= 227 * 1024, "attention_bwd: L=%d dh=%d does not fit shared memory (L*dh too large)", L, dh);
    AFB_REQUIRE(B * heads < (1ll << 31), "attention: too many (sequence, head) items");
    int rc;
    if (dt == AFB_BF16) {
      if ((rc = configure_smem(attn_long_bwd_kernel<bf16>, smem))) return rc;
      attn_long_bwd_kernel<bf16><<<(unsigned)(B * heads), kLongWarps * 32, smem, as_stream(s)>>>((const bf16*)qkv, (const bf16*)dO, (bf16*)dqkv, L, heads, dh, scale);
    } else {
      if ((rc = configure_smem(attn_long_bwd_kernel<float>, smem))) return rc;
      attn_long_bwd_kernel<float><<<(unsigned)(B * heads), kLongWarps * 32, smem, as_stream(s)>>>((const float*)qkv, (const float*)dO, (float*)dqkv, L, heads, dh, scale);
    }
    return check_launch("attention_long_bwd");
  }
  // Backward on the tcgen05 / TMEM / TMA kernel (attention_tc.cu) at dh = 32: 150 vs 182 us at L = 22, 145 vs 357 us at L = 64
  // (B*L = 180k tokens).  AFB_ATTN_TC_BWD=0 forces the warp-level kernel.
  if (dt == AFB_BF16 && attention_bwd_tc_supported(L, heads, dh)) {
    const char* tc = getenv("AFB_ATTN_TC_BWD");
    if (tc == nullptr || tc[0] != '0') return attention_bwd_tc(qkv, dO, dqkv, B, L, heads, dh, scale, as_stream(s));
  }
  if (use_mma(dt, L, heads, dh)) return attention_bwd_mma(qkv, dO, dqkv, B, L, heads, dh, scale, as_stream(s));
  const int per_warp = 4 * L * (dh + 1) + 2 * L * (L + 1) + 2;
  int warps = kSmemBudget / (per_warp * 4);
  if (warps > 4) warps = 4;
  AFB_REQUIRE(warps >= 1, "attention: tile does not fit shared memory");
  const int64_t items = B * heads;
  const int smem = warps * per_warp * 4;
  const int grid = ceil_div(items, warps);
  int rc;
  if (dt == AFB_BF16) {
    if ((rc = configure_smem(attn_bwd_kernel<bf16>, smem))) return rc;
    attn_bwd_kernel<bf16><<<grid, warps * 32, smem, as_stream(s)>>>((const bf16*)qkv, (const bf16*)dO, (bf16*)dqkv, items, L, heads, dh,
                                                                   scale, per_warp);
  } else {
    if ((rc = configure_smem(attn_bwd_kernel<float>, smem))) return rc;
    attn_bwd_kernel<float><<<grid, warps * 32, smem, as_stream(s)>>>((const float*)qkv, (const float*)dO, (float*)dqkv, items, L, heads,
                                                                     dh, scale, per_warp);
  }
  return check_launch("attention_bwd");
}
