// bf16 LayerNorm fast path (D = 256 / 512, the two AltFormer widths): one warp per token row, every global
// access is a 16-byte vector (8 bf16 per lane), statistics in fp32.  nn.LayerNorm at model_ST.py:75,80,96.
// Backward fuses the residual-gradient add (dx = dres + LN'(dy)) and accumulates dgamma / dbeta per CTA in
// shared memory before one atomicAdd per column.
#include "common.cuh"

namespace afb {
namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ void unpack8(const uint4& r, float (&o)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = __uint_as_float(w[j] << 16);
    o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void load8f(const float* p, float (&o)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// NC = D / 256 chunks of 8 elements per lane
template <int NC>
__global__ void __launch_bounds__(kBlock) ln_fwd_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, bf16* __restrict__ y, float* __restrict__ mean,
                                                             float* __restrict__ rstd, int64_t rows, float eps, int rev) {
  constexpr int D = NC * 256;
  pdl_launch_dependents();
  pdl_wait();   // programmatic dependent launch (common.cuh): hides the launch latency, there is no prologue to overlap
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float g[NC][8], b[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    load8f(gamma + c * 256 + lane * 8, g[c]);
    load8f(beta + c * 256 + lane * 8, b[c]);
  }
  // two rows per iteration: both rows' loads are issued before either reduction starts (twice the bytes in flight)
  // rev: rows visited in descending order (ping-pong traversal, api.cu:next_stream_dir)
  for (int64_t ri = warp0; ri < rows; ri += 2 * nwarps) {
    const bool has2 = ri + nwarps < rows;
    const int64_t row = rev ? rows - 1 - ri : ri;
    const int64_t row2 = has2 ? (rev ? rows - 1 - (ri + nwarps) : ri + nwarps) : row;
    uint4 ra[NC], rb[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      ra[c] = *reinterpret_cast<const uint4*>(x + row * D + c * 256 + lane * 8);
      rb[c] = has2 ? *reinterpret_cast<const uint4*>(x + row2 * D + c * 256 + lane * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
    float va[NC][8], vb[NC][8];
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      unpack8(ra[c], va[c]);
      unpack8(rb[c], vb[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) { sa += va[c][i]; sb += vb[c][i]; }
    }
    const float mua = warp_sum(sa) * (1.0f / D), mub = warp_sum(sb) * (1.0f / D);
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        va[c][i] -= mua;
        vb[c][i] -= mub;
        qa += va[c][i] * va[c][i];
        qb += vb[c][i] * vb[c][i];
      }
    const float rsa = rsqrtf(warp_sum(qa) * (1.0f / D) + eps), rsb = rsqrtf(warp_sum(qb) * (1.0f / D) + eps);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(va[c][i] * rsa, g[c][i], b[c][i]);
      *reinterpret_cast<uint4*>(y + row * D + c * 256 + lane * 8) = pack8(o);
      if (has2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(vb[c][i] * rsb, g[c][i], b[c][i]);
        *reinterpret_cast<uint4*>(y + row2 * D + c * 256 + lane * 8) = pack8(o);
      }
    }
    if (lane == 0) {
      if (mean) { mean[row] = mua; if (has2) mean[row2] = mub; }
      if (rstd) { rstd[row] = rsa; if (has2) rstd[row2] = rsb; }
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(kBlock) ln_bwd_bf16_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                             const float* __restrict__ gamma, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, const bf16* __restrict__ dres,
                                                             bf16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             int64_t rows, int rev) {
  constexpr int D = NC * 256;
  __shared__ float red[2][D];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) (&red[0][0])[i] = 0.f;
  float g[NC][8], ag[NC][8], ab[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    load8f(gamma + c * 256 + lane * 8, g[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { ag[c][i] = 0.f; ab[c][i] = 0.f; }
  }
  __syncthreads();
  for (int64_t ri = warp0; ri < rows; ri += nwarps) {
    const int64_t row = rev ? rows - 1 - ri : ri;
    uint4 rdy[NC], rx[NC], rres[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {   // all loads of the row first (memory-level parallelism)
      rdy[c] = *reinterpret_cast<const uint4*>(dy + row * D + c * 256 + lane * 8);
      rx[c] = *reinterpret_cast<const uint4*>(x + row * D + c * 256 + lane * 8);
      if (dres != nullptr) rres[c] = *reinterpret_cast<const uint4*>(dres + row * D + c * 256 + lane * 8);
    }
    const float mu = mean[row], rs = rstd[row];
    float xh[NC][8], gy[NC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float d[8], xv[8];
      unpack8(rdy[c], d);
      unpack8(rx[c], xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[c][i] = (xv[i] - mu) * rs;
        ab[c][i] += d[i];
        ag[c][i] = fmaf(d[i], xh[c][i], ag[c][i]);
        gy[c][i] = d[i] * g[c][i];
        s1 += gy[c][i];
        s2 = fmaf(gy[c][i], xh[c][i], s2);
      }
    }
    const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = rs * (gy[c][i] - m1 - xh[c][i] * m2);
      if (dres != nullptr) {
        float r[8];
        unpack8(rres[c], r);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += r[i];
      }
      *reinterpret_cast<uint4*>(dx + row * D + c * 256 + lane * 8) = pack8(o);
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&red[0][c * 256 + lane * 8 + i], ag[c][i]);
      atomicAdd(&red[1][c * 256 + lane * 8 + i], ab[c][i]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dgamma + c, red[0][c]);
    atomicAdd(dbeta + c, red[1][c]);
  }
}

// rows_per_warp: minimum rows each warp should own (amortises per-block setup and the backward's dgamma/dbeta atomics)
inline int grid_rows(int64_t rows, int max_blocks, int rows_per_warp = 1) {
  const int64_t per_block = (int64_t)(kBlock / 32) * rows_per_warp;
  int64_t g = (rows + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace

int layernorm_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows, int D,
                       float eps, cudaStream_t st) {
  const int grid = grid_rows(rows, 148 * 8, 2);
  const int rev = next_stream_dir();
  const bool small = pdl_force_all() || rows <= 65536;
  cudaError_t le = D == 256 ? launch_pdl(small, ln_fwd_bf16_kernel<1>, dim3(grid), dim3(kBlock), 0, st, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, rows, eps, rev)
                            : launch_pdl(small, ln_fwd_bf16_kernel<2>, dim3(grid), dim3(kBlock), 0, st, (const bf16*)x, gamma, beta, (bf16*)y, mean, rstd, rows, eps, rev);
  if (le != cudaSuccess) {
    set_error("layernorm_fwd_bf16: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("layernorm_fwd_bf16");
}

int layernorm_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
                       void* dx, float* dgamma, float* dbeta, int64_t rows, int D, cudaStream_t st) {
  const int grid = grid_rows(rows, 148 * 6, 8);   // every block ends with 2*D global atomics
  const int rev = next_stream_dir();
  const bool small = pdl_force_all() || rows <= 65536;
  cudaError_t le = D == 256 ? launch_pdl(small, ln_bwd_bf16_kernel<1>, dim3(grid), dim3(kBlock), 0, st, (const bf16*)dy, (const bf16*)x, gamma, mean, rstd,
                                         (const bf16*)dres, (bf16*)dx, dgamma, dbeta, rows, rev)
                            : launch_pdl(small, ln_bwd_bf16_kernel<2>, dim3(grid), dim3(kBlock), 0, st, (const bf16*)dy, (const bf16*)x, gamma, mean, rstd,
                                         (const bf16*)dres, (bf16*)dx, dgamma, dbeta, rows, rev);
  if (le != cudaSuccess) {
    set_error("layernorm_bwd_bf16: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("layernorm_bwd_bf16");
}

}  // namespace afb
