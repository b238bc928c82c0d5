// Generic strided CUDA-core GEMM (fp32 accumulate).  Used for the tiny classifier head
// (512 -> num_class, N not a multiple of 64) and as the on-device cross-check of the tcgen05 GEMMs.
#include "common.cuh"

namespace afb {
namespace {

constexpr int TS = 32;  // tile edge; 256 threads, each computes a 2x2 micro-tile

__device__ __forceinline__ void atomic_add_f32(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_f32(bf16*, float) {}   // split-K is never selected for bf16 outputs

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B, TC* __restrict__ C,
                                                        const float* __restrict__ bias, int M, int N, int K, int64_t sai, int64_t sak,
                                                        int64_t sbj, int64_t sbk, int64_t sci, int64_t scj, float alpha, float beta) {
  __shared__ float As[TS][TS + 1];
  __shared__ float Bs[TS][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  // split-K (gridDim.z > 1, fp32 C only): each z slice owns a range of 32-wide k-steps and adds its partial with
  // atomics; the host has already applied beta (zeroed C when beta == 0), slice 0 adds the bias.
  const int ksteps = (K + TS - 1) / TS, per = (ksteps + gridDim.z - 1) / gridDim.z;
  const int kbeg = blockIdx.z * per * TS, kend = min(K, (int)(blockIdx.z + 1) * per * TS);
  const bool split = gridDim.z > 1;
  for (int k0 = kbeg; k0 < kend; k0 += TS) {
    for (int e = threadIdx.x; e < TS * TS; e += 256) {
      // pick the faster-varying index along whichever stride is 1 so loads coalesce
      int r, c;
      if (sak == 1) { r = e / TS; c = e % TS; } else { c = e / TS; r = e % TS; }
      const int i = i0 + r, k = k0 + c;
      As[r][c] = (i < M && k < K) ? ldf<TA>(A + i * sai + k * sak) : 0.f;
    }
    for (int e = threadIdx.x; e < TS * TS; e += 256) {
      int r, c;
      if (sbk == 1) { r = e / TS; c = e % TS; } else { c = e / TS; r = e % TS; }
      const int j = j0 + r, k = k0 + c;
      Bs[r][c] = (j < N && k < K) ? ldf<TB>(B + j * sbj + k * sbk) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < TS; ++k) {
      const float a0 = As[ty][k], a1 = As[ty + 16][k];
      const float b0 = Bs[tx][k], b1 = Bs[tx + 16][k];
      acc[0][0] += a0 * b0; acc[0][1] += a0 * b1;
      acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int i = i0 + ty + 16 * u, j = j0 + tx + 16 * w;
      if (i < M && j < N) {
        float v = alpha * acc[u][w] + ((bias && blockIdx.z == 0) ? bias[j] : 0.f);
        TC* dst = C + i * sci + j * scj;
        if (split) {
          atomic_add_f32(dst, v);
        } else {
          if (beta != 0.f) v += beta * ldf<TC>(dst);
          stf<TC>(dst, v);
        }
      }
    }
}

template <typename TA, typename TB>
int launch_c(const afb_gemm_simt_t* p, cudaStream_t st) {
  dim3 grid(ceil_div(p->N, TS), ceil_div(p->M, TS));
  // The classifier-head shapes give 1-16 tiles with 8-16 serial k-steps each: split K over gridDim.z until ~64 CTAs
  // exist.  Needs an fp32, contiguous-or-strided but non-aliased C and beta in {0, 1} (beta = 0: C is zeroed first).
  const int ksteps = ceil_div(p->K, TS);
  int splits = 1;
  const bool dense = p->scj == 1 && p->sci == p->N;
  if (p->c_dtype == AFB_F32 && (p->beta == 1.f || (p->beta == 0.f && dense)) && ksteps >= 4) {
    while (splits * 2 <= ksteps / 2 && (int)(grid.x * grid.y) * splits < 64) splits *= 2;
  }
  if (splits > 1) {
    grid.z = splits;
    if (p->beta == 0.f) cudaMemsetAsync(p->C, 0, (size_t)p->M * p->N * sizeof(float), st);
  }
  if (p->c_dtype == AFB_BF16)
    gemm_simt_kernel<TA, TB, bf16><<<grid, 256, 0, st>>>((const TA*)p->A, (const TB*)p->B, (bf16*)p->C, p->bias, p->M, p->N, p->K, p->sai,
                                                         p->sak, p->sbj, p->sbk, p->sci, p->scj, p->alpha, p->beta);
  else
    gemm_simt_kernel<TA, TB, float><<<grid, 256, 0, st>>>((const TA*)p->A, (const TB*)p->B, (float*)p->C, p->bias, p->M, p->N, p->K,
                                                          p->sai, p->sak, p->sbj, p->sbk, p->sci, p->scj, p->alpha, p->beta);
  return check_launch("gemm_simt");
}

}  // namespace
}  // namespace afb

using namespace afb;

extern "C" int afb_gemm_simt(const afb_gemm_simt_t* p, afb_stream s) {
  AFB_REQUIRE(p && p->A && p->B && p->C && p->M > 0 && p->N > 0 && p->K > 0, "gemm_simt: bad args");
  cudaStream_t st = as_stream(s);
  if (p->a_dtype == AFB_BF16 && p->b_dtype == AFB_BF16) return launch_c<bf16, bf16>(p, st);
  if (p->a_dtype == AFB_BF16) return launch_c<bf16, float>(p, st);
  if (p->b_dtype == AFB_BF16) return launch_c<float, bf16>(p, st);
  return launch_c<float, float>(p, st);
}
