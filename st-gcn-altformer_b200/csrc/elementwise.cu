// Bandwidth-bound support kernels: casts, LayerNorm, BatchNorm(+ReLU/residual), column reductions,
// pooling, softmax-CE, AdamW, stream transforms.  All math in fp32 (fp64 for BN statistics).
#include "common.cuh"

namespace afb {
namespace {

constexpr int kBlock = 256;
inline int grid_for(int64_t work, int per_block, int max_blocks = 148 * 16) {
  int64_t g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// casts / packing
// ---------------------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    st4<D>(dst + 4 * i, ld4<S>(src + 4 * i));
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    stf<D>(dst + i, ldf<S>(src + i));
  }
}

__global__ void cast_transpose_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// w (co, ci, k) -> fwd [co][k][ci], bwd [ci][k-1-tap][co]
__global__ void conv_pack_kernel(const float* __restrict__ w, bf16* __restrict__ fwd, bf16* __restrict__ bwd, int co, int ci, int k) {
  const int64_t total = (int64_t)co * ci * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % k);
    const int c_in = (int)((i / k) % ci);
    const int c_out = (int)(i / ((int64_t)k * ci));
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (fwd) fwd[((int64_t)c_out * k + tap) * ci + c_in] = v;
    if (bwd) bwd[((int64_t)c_in * k + (k - 1 - tap)) * co + c_out] = v;
  }
}

// dW (co, ci, k) += tmp [k][co][ci]
__global__ void conv_dw_unpack_kernel(const float* __restrict__ tmp, float* __restrict__ dW, int co, int ci, int k) {
  const int64_t total = (int64_t)co * ci * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % k);
    const int64_t oc = i / k;   // c_out * ci + c_in
    dW[i] += tmp[(int64_t)tap * co * ci + oc];
  }
}

template <typename S, typename D>
__global__ void copy2d_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst, int64_t ldd, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    stf<D>(dst + r * ldd + c, ldf<S>(src + r * lds + c));
  }
}

__global__ void split3_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int64_t rows, int cols, int which) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i % cols);
    const float v = src[i];
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    if (which == 2) {  // row-stacked (hi ; hi ; lo): the MN-major B operand of dX = dY * W
      dst[i] = hi;
      dst[total + i] = hi;
      dst[2 * total + i] = lo;
    } else {
      bf16* o = dst + r * 3 * cols;
      o[c] = hi;
      o[cols + c] = which == 0 ? lo : hi;
      o[2 * cols + c] = which == 0 ? hi : lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, D = 128*NV (NV float4 chunks per lane)
// ---------------------------------------------------------------------------------------------
template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(kBlock) ln_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, TO* __restrict__ y,
                                                        float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[NV], b[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    g[j] = *reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4);
    b[j] = *reinterpret_cast<const float4*>(beta + j * 128 + lane * 4);
  }
  for (int64_t row = warp0; row < rows; row += nwarps) {
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j] = ld4<TI>(x + row * D + j * 128 + lane * 4);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j].x -= mu; v[j].y -= mu; v[j].z -= mu; v[j].w -= mu;
      q += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = v[j].x * rs * g[j].x + b[j].x;
      o.y = v[j].y * rs * g[j].y + b[j].y;
      o.z = v[j].z * rs * g[j].z + b[j].z;
      o.w = v[j].w * rs * g[j].w + b[j].w;
      st4<TO>(y + row * D + j * 128 + lane * 4, o);
    }
    if (lane == 0) {
      if (mean) mean[row] = mu;
      if (rstd) rstd[row] = rs;
    }
  }
}

template <typename TG, typename TX, typename TR, typename TO, int NV>
__global__ void __launch_bounds__(kBlock) ln_bwd_kernel(const TG* __restrict__ dy, const TX* __restrict__ x,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const TR* __restrict__ dres,
                                                        TO* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                        int64_t rows) {
  constexpr int D = NV * 128;
  __shared__ float red[2][kBlock / 32][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    g[j] = *reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4);
    ag[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[NV], gy[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 xv = ld4<TX>(x + row * D + j * 128 + lane * 4);
      const float4 d = ld4<TG>(dy + row * D + j * 128 + lane * 4);
      xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      ab[j].x += d.x; ab[j].y += d.y; ab[j].z += d.z; ab[j].w += d.w;
      ag[j].x += d.x * xh[j].x; ag[j].y += d.y * xh[j].y; ag[j].z += d.z * xh[j].z; ag[j].w += d.w * xh[j].w;
      gy[j] = make_float4(d.x * g[j].x, d.y * g[j].y, d.z * g[j].z, d.w * g[j].w);
      s1 += gy[j].x + gy[j].y + gy[j].z + gy[j].w;
      s2 += gy[j].x * xh[j].x + gy[j].y * xh[j].y + gy[j].z * xh[j].z + gy[j].w * xh[j].w;
    }
    const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = rs * (gy[j].x - m1 - xh[j].x * m2);
      o.y = rs * (gy[j].y - m1 - xh[j].y * m2);
      o.z = rs * (gy[j].z - m1 - xh[j].z * m2);
      o.w = rs * (gy[j].w - m1 - xh[j].w * m2);
      if (dres != nullptr) {
        const float4 r = ld4<TR>(dres + row * D + j * 128 + lane * 4);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      st4<TO>(dx + row * D + j * 128 + lane * 4, o);
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    *reinterpret_cast<float4*>(&red[0][warp][j * 128 + lane * 4]) = ag[j];
    *reinterpret_cast<float4*>(&red[1][warp][j * 128 + lane * 4]) = ab[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) { sg += red[0][w][c]; sb += red[1][w][c]; }
    atomicAdd(dgamma + c, sg);
    atomicAdd(dbeta + c, sb);
  }
}

// ---------------------------------------------------------------------------------------------
// column reductions over token rows: block = (C/4 channel groups) x (row lanes)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t perm_row(int64_t m, int T, int V) {  // (n,t,v) -> (n,v,t)
  const int v = (int)(m % V);
  const int64_t nt = m / V;
  const int t = (int)(nt % T);
  const int64_t n = nt / T;
  return (n * V + v) * T + t;
}

// Column tiling shared by the three reductions: a block covers 128 columns (32 float4 lanes) x 8 row
// lanes and a contiguous slab of rows; grid = (row slabs, ceil(C/128)).
constexpr int kColTile = 128;
constexpr int kRowLanes = kBlock / (kColTile / 4);

struct ColTile {
  int c4, rl;
  bool active;
  int64_t r_begin, r_end;
};
__device__ __forceinline__ ColTile col_tile(int64_t M, int C) {
  ColTile t;
  t.c4 = blockIdx.y * kColTile + (threadIdx.x & 31) * 4;
  t.rl = threadIdx.x >> 5;
  t.active = t.c4 < C;
  const int64_t rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  t.r_begin = blockIdx.x * rows_per_block;
  t.r_end = t.r_begin + rows_per_block < M ? t.r_begin + rows_per_block : M;
  return t;
}

template <typename T>
__global__ void __launch_bounds__(kBlock) colstats_kernel(const T* __restrict__ x, int64_t M, int C, int ldx, double* __restrict__ sum,
                                                          double* __restrict__ sumsq) {
  __shared__ float sred[2][kRowLanes][kColTile];
  const ColTile t = col_tile(M, C);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  if (t.active) {
    int64_t r = t.r_begin + t.rl;
    for (; r + 3 * kRowLanes < t.r_end; r += 4 * kRowLanes) {   // four independent row loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ld4<T>(x + (r + k * kRowLanes) * ldx + t.c4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w;
        q.x += v[k].x * v[k].x; q.y += v[k].y * v[k].y; q.z += v[k].z * v[k].z; q.w += v[k].w * v[k].w;
      }
    }
    for (; r < t.r_end; r += kRowLanes) {
      const float4 v = ld4<T>(x + r * ldx + t.c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
    }
  }
  *reinterpret_cast<float4*>(&sred[0][t.rl][(threadIdx.x & 31) * 4]) = s;
  *reinterpret_cast<float4*>(&sred[1][t.rl][(threadIdx.x & 31) * 4]) = q;
  __syncthreads();
  if (threadIdx.x < kColTile) {
    const int c = blockIdx.y * kColTile + threadIdx.x;
    if (c < C) {
      double a = 0.0, b = 0.0;
      for (int l = 0; l < kRowLanes; ++l) { a += sred[0][l][threadIdx.x]; b += sred[1][l][threadIdx.x]; }
      atomicAdd(sum + c, a);
      atomicAdd(sumsq + c, b);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) colsum_kernel(const T* __restrict__ x, int64_t M, int C, int ldx,
                                                        const float* __restrict__ row_scale, int div, float* __restrict__ out) {
  __shared__ float sred[kRowLanes][kColTile];
  const ColTile t = col_tile(M, C);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t.active)
    for (int64_t r = t.r_begin + t.rl; r < t.r_end; r += kRowLanes) {
      float4 v = ld4<T>(x + r * ldx + t.c4);
      if (row_scale != nullptr) {
        const float sc = row_scale[r / div];
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      }
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  *reinterpret_cast<float4*>(&sred[t.rl][(threadIdx.x & 31) * 4]) = s;
  __syncthreads();
  if (threadIdx.x < kColTile) {
    const int c = blockIdx.y * kColTile + threadIdx.x;
    if (c < C) {
      float a = 0.f;
      for (int l = 0; l < kRowLanes; ++l) a += sred[l][threadIdx.x];
      atomicAdd(out + c, a);
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, int64_t M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rm,
                                   float* __restrict__ rv, float momentum, float eps, int training, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mu, var;
  if (training) {
    mu = sum[c] / (double)M;
    var = sumsq[c] / (double)M - mu * mu;
    if (var < 0.0) var = 0.0;
    if (rm != nullptr) {
      const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
      rm[c] = (float)((1.0 - momentum) * rm[c] + momentum * mu);
      rv[c] = (float)((1.0 - momentum) * rv[c] + momentum * unb);
    }
  } else {
    mu = rm[c];
    var = rv[c];
  }
  const float rs = (float)(1.0 / sqrt(var + (double)eps));
  if (mean) mean[c] = (float)mu;
  if (rstd) rstd[c] = rs;
  const float sc = gamma[c] * rs;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mu * sc;
}

template <typename TX, typename TR, typename TY>
__global__ void __launch_bounds__(kBlock) bn_act_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const TR* __restrict__ res_pre,
                                                            const TR* __restrict__ res_post, int relu, TY* __restrict__ y,
                                                            TY* __restrict__ y2, int64_t M, int C, int T, int V) {
  const int cg = C >> 2;
  const int64_t total = M * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / cg;
    const int c4 = (int)(i % cg) * 4;
    const float4 v = ld4<TX>(x + m * C + c4);
    const float4 sc = *reinterpret_cast<const float4*>(scale + c4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c4);
    float4 o = make_float4(v.x * sc.x + sh.x, v.y * sc.y + sh.y, v.z * sc.z + sh.z, v.w * sc.w + sh.w);
    if (res_pre != nullptr) {
      const float4 r = ld4<TR>(res_pre + m * C + c4);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (res_post != nullptr) {
      const float4 r = ld4<TR>(res_post + m * C + c4);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    if (y != nullptr) st4<TY>(y + m * C + c4, o);
    if (y2 != nullptr) st4<TY>(y2 + perm_row(m, T, V) * C + c4, o);
  }
}

// g = (dy[m] + dy2[perm(m)]) * [pre-activation > 0]; the relu mask is recomputed from x, scale/shift
// (+res_pre) so the forward output does not have to be kept alive.
template <typename TG, typename TX, typename TR>
__device__ __forceinline__ float4 bn_masked_grad(const TG* dy, const TG* dy2, const TX* x, const TR* res_pre, const float4 mu,
                                                 const float4 rs, const float4 ga, const float4 be, int relu, int64_t m, int c4, int C,
                                                 int T, int V, float4& xh) {
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (dy != nullptr) g = ld4<TG>(dy + m * C + c4);
  if (dy2 != nullptr) {
    const float4 h = ld4<TG>(dy2 + perm_row(m, T, V) * C + c4);
    g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
  }
  const float4 xv = ld4<TX>(x + m * C + c4);
  xh = make_float4((xv.x - mu.x) * rs.x, (xv.y - mu.y) * rs.y, (xv.z - mu.z) * rs.z, (xv.w - mu.w) * rs.w);
  if (relu) {
    float4 pre = make_float4(xh.x * ga.x + be.x, xh.y * ga.y + be.y, xh.z * ga.z + be.z, xh.w * ga.w + be.w);
    if (res_pre != nullptr) {
      const float4 r = ld4<TR>(res_pre + m * C + c4);
      pre.x += r.x; pre.y += r.y; pre.z += r.z; pre.w += r.w;
    }
    if (pre.x <= 0.f) g.x = 0.f;
    if (pre.y <= 0.f) g.y = 0.f;
    if (pre.z <= 0.f) g.z = 0.f;
    if (pre.w <= 0.f) g.w = 0.f;
  }
  return g;
}

template <typename TG, typename TX, typename TR>
__global__ void __launch_bounds__(kBlock) bn_bwd_reduce_kernel(const TG* __restrict__ dy, const TG* __restrict__ dy2,
                                                               const TX* __restrict__ x, const TR* __restrict__ res_pre,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int C,
                                                               int T, int V) {
  __shared__ float sred[2][kRowLanes][kColTile];
  const ColTile t = col_tile(M, C);
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
  if (t.active) {
    const int c4 = t.c4;
    const float4 mu = *reinterpret_cast<const float4*>(mean + c4), rs = *reinterpret_cast<const float4*>(rstd + c4);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c4), be = *reinterpret_cast<const float4*>(beta + c4);
    for (int64_t r = t.r_begin + t.rl; r < t.r_end; r += kRowLanes) {
      float4 xh;
      const float4 g = bn_masked_grad<TG, TX, TR>(dy, dy2, x, res_pre, mu, rs, ga, be, relu, r, c4, C, T, V, xh);
      sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
      sg.x += g.x * xh.x; sg.y += g.y * xh.y; sg.z += g.z * xh.z; sg.w += g.w * xh.w;
    }
  }
  *reinterpret_cast<float4*>(&sred[0][t.rl][(threadIdx.x & 31) * 4]) = sg;
  *reinterpret_cast<float4*>(&sred[1][t.rl][(threadIdx.x & 31) * 4]) = sb;
  __syncthreads();
  if (threadIdx.x < kColTile) {
    const int c = blockIdx.y * kColTile + threadIdx.x;
    if (c < C) {
      float a = 0.f, b = 0.f;
      for (int l = 0; l < kRowLanes; ++l) { a += sred[0][l][threadIdx.x]; b += sred[1][l][threadIdx.x]; }
      atomicAdd(dgamma + c, a);
      atomicAdd(dbeta + c, b);
    }
  }
}

template <typename TG, typename TX, typename TR, typename TO>
__global__ void __launch_bounds__(kBlock) bn_bwd_apply_kernel(const TG* __restrict__ dy, const TG* __restrict__ dy2,
                                                              const TX* __restrict__ x, const TR* __restrict__ res_pre,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu,
                                                              int training, TO* __restrict__ dx, TO* __restrict__ dres, int64_t M, int C,
                                                              int T, int V) {
  const int cg = C >> 2;
  const int64_t total = M * cg;
  const float inv_m = training ? 1.0f / (float)M : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / cg;
    const int c4 = (int)(i % cg) * 4;
    const float4 mu = *reinterpret_cast<const float4*>(mean + c4), rs = *reinterpret_cast<const float4*>(rstd + c4);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + c4), be = *reinterpret_cast<const float4*>(beta + c4);
    const float4 dg = *reinterpret_cast<const float4*>(dgamma + c4), db = *reinterpret_cast<const float4*>(dbeta + c4);
    float4 xh;
    const float4 g = bn_masked_grad<TG, TX, TR>(dy, dy2, x, res_pre, mu, rs, ga, be, relu, m, c4, C, T, V, xh);
    float4 o;
    o.x = ga.x * rs.x * (g.x - db.x * inv_m - xh.x * dg.x * inv_m);
    o.y = ga.y * rs.y * (g.y - db.y * inv_m - xh.y * dg.y * inv_m);
    o.z = ga.z * rs.z * (g.z - db.z * inv_m - xh.z * dg.z * inv_m);
    o.w = ga.w * rs.w * (g.w - db.w * inv_m - xh.w * dg.w * inv_m);
    st4<TO>(dx + m * C + c4, o);
    if (dres != nullptr) st4<TO>(dres + m * C + c4, g);
  }
}

// ---------------------------------------------------------------------------------------------
// pooling
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_mean_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t B, int L, int D) {
  const int d2 = D >> 1;
  const int64_t total = B * d2;
  const float inv = 1.0f / L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d2;
    const int d = (int)(i % d2) * 2;
    float a0 = 0.f, a1 = 0.f;
    for (int l = 0; l < L; ++l) {
      const float2 v = ld2<T>(x + (b * L + l) * D + d);
      a0 += v.x; a1 += v.y;
    }
    st2<T>(y + b * D + d, a0 * inv, a1 * inv);
  }
}
template <typename T>
__global__ void pool_mean_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t B, int L, int D) {
  const int d2 = D >> 1;
  const int64_t total = B * L * d2;
  const float inv = 1.0f / L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % d2) * 2;
    const int64_t bl = i / d2;
    const int64_t b = bl / L;
    const float2 g = ld2<T>(dy + b * D + d);
    st2<T>(dx + bl * D + d, g.x * inv, g.y * inv);
  }
}
// bf16 fast paths (D % 8 == 0): 16-byte accesses, 8 channels per thread
__global__ void pool_mean_fwd_bf16x8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t B, int L, int D) {
  const int d8 = D >> 3;
  const int64_t total = B * d8;
  const float inv = 1.0f / L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d8;
    const int d = (int)(i % d8) * 8;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bf16* src = x + b * L * D + d;
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const uint4 r = *reinterpret_cast<const uint4*>(src + (int64_t)l * D);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[2 * j] += __uint_as_float(w[j] << 16);
        a[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * j] * inv, a[2 * j + 1] * inv);
      o[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(y + b * D + d) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
__global__ void pool_mean_bwd_bf16x8_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dx, int64_t B, int L, int D) {
  const int d8 = D >> 3;
  const int64_t total = B * d8;
  const float inv = 1.0f / L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / d8;
    const int d = (int)(i % d8) * 8;
    const uint4 r = *reinterpret_cast<const uint4*>(dy + b * D + d);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(w[j] << 16) * inv, __uint_as_float(w[j] & 0xffff0000u) * inv);
      o[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
    bf16* dst = dx + b * L * D + d;
    for (int l = 0; l < L; ++l) *reinterpret_cast<uint4*>(dst + (int64_t)l * D) = v;
  }
}
template <typename T>
__global__ void pool_max_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int32_t* __restrict__ arg, int64_t B, int L, int D) {
  const int64_t total = B * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / D;
    const int d = (int)(i % D);
    float best = ldf<T>(x + (b * L) * D + d);
    int bi = 0;
    for (int l = 1; l < L; ++l) {
      const float v = ldf<T>(x + (b * L + l) * D + d);
      if (v > best) { best = v; bi = l; }
    }
    stf<T>(y + i, best);
    arg[i] = bi;
  }
}
template <typename T>
__global__ void pool_max_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ arg, T* __restrict__ dx, int64_t B, int L,
                                    int D) {
  const int64_t total = B * L * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const int64_t bl = i / D;
    const int l = (int)(bl % L);
    const int64_t b = bl / L;
    stf<T>(dx + i, arg[b * D + d] == l ? ldf<T>(dy + b * D + d) : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// softmax cross-entropy (mean reduction), one warp per sample
// ---------------------------------------------------------------------------------------------
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ loss,
                                  float* __restrict__ dlogits, int N, int C) {
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= N) return;
  const float* z = logits + (int64_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(z[c] - mx);
  s = warp_sum(s);
  const int y = (int)labels[row];
  const float inv_n = 1.0f / N;
  if (dlogits != nullptr)
    for (int c = lane; c < C; c += 32) dlogits[(int64_t)row * C + c] = (expf(z[c] - mx) / s - (c == y ? 1.f : 0.f)) * inv_n;
  if (lane == 0 && loss != nullptr) atomicAdd(loss, (logf(s) + mx - z[y]) * inv_n);
}

// ---------------------------------------------------------------------------------------------
// AdamW over a flat buffer (torch.optim.AdamW semantics)
// ---------------------------------------------------------------------------------------------
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             bf16* __restrict__ pb, int64_t n, const int32_t* __restrict__ step, float lr, float b1, float b2, float eps,
                             float wd, float grad_scale) {
  const float t = (float)(*step);
  const float bc1 = 1.0f - powf(b1, t), bc2 = 1.0f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.0f - lr * wd;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pv = *reinterpret_cast<float4*>(p + 4 * i);
    const float4 gv = *reinterpret_cast<const float4*>(g + 4 * i);
    float4 mv = *reinterpret_cast<float4*>(m + 4 * i), vv = *reinterpret_cast<float4*>(v + 4 * i);
    float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k] * grad_scale;
      mp[k] = b1 * mp[k] + (1.0f - b1) * gk;
      vp[k] = b2 * vp[k] + (1.0f - b2) * gk * gk;
      pp[k] = pp[k] * decay - step_size * mp[k] / (sqrtf(vp[k]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + 4 * i) = pv;
    *reinterpret_cast<float4*>(m + 4 * i) = mv;
    *reinterpret_cast<float4*>(v + 4 * i) = vv;
    if (pb != nullptr) st4<bf16>(pb + 4 * i, pv);
  }
}
__global__ void step_inc_kernel(int32_t* step) { *step += 1; }

template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t M, int C, const float* __restrict__ rs, int div) {
  const int cg = C >> 2;
  const int64_t total = M * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / cg;
    const float sc = rs[m / div];
    float4 v = ld4<T>(x + 4 * i);
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    st4<T>(y + 4 * i, v);
  }
}

template <typename T>
__global__ void scale_rows_any_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t M, int C, const float* __restrict__ rs, int div) {
  const int64_t total = M * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    stf<T>(y + i, ldf<T>(x + i) * rs[(i / C) / div]);
}

// ---------------------------------------------------------------------------------------------
// input streams / ensemble
// ---------------------------------------------------------------------------------------------
__global__ void bone_kernel(const float* __restrict__ x, const int32_t* __restrict__ parent, float* __restrict__ y, int64_t NT, int V) {
  const int64_t total = NT * V * 3;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    const int64_t jv = i / 3;
    const int v = (int)(jv % V);
    const int64_t nt = jv / V;
    y[i] = x[i] - x[(nt * V + parent[v]) * 3 + c];
  }
}
__global__ void motion_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int T, int V) {
  const int64_t frame = (int64_t)V * 3, total = (int64_t)N * T * frame;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)((i / frame) % T);
    y[i] = t + 1 < T ? x[i + frame] - x[i] : 0.f;
  }
}
// strided temporal convolution = sum over phases of stride-1 convolutions on phase tensors:
//   gather : y[n, j, v, :] = x[n, j * stride + phase, v, :]  (zero when that frame is >= T)      y: [N, To, V, C]
//   scatter: x[n, j * stride + phase, v, :] = y[n, j, v, :]  for the frames that exist           (inverse, backward pass)
template <typename T, bool SCATTER>
__global__ void __launch_bounds__(kBlock) frame_phase_kernel(const T* __restrict__ src, T* __restrict__ dst, int N, int Tn, int To, int V,
                                                             int C, int stride, int phase) {
  const int c4n = C >> 2;
  const int64_t total = (int64_t)N * To * V * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % c4n);
    const int64_t r = i / c4n;                 // (n, j, v) row of the phase tensor
    const int v = (int)(r % V);
    const int j = (int)((r / V) % To);
    const int n = (int)(r / ((int64_t)V * To));
    const int f = j * stride + phase;
    const int64_t full = (((int64_t)n * Tn + f) * V + v) * C + 4 * c4, ph = r * C + 4 * c4;
    if (SCATTER) {
      if (f < Tn) st4<T>(dst + full, ld4<T>(src + ph));
    } else {
      st4<T>(dst + ph, f < Tn ? ld4<T>(src + full) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// y = a * b elementwise, 4 elements per thread (dropout mask application, net.py:48)
template <typename T>
__global__ void __launch_bounds__(kBlock) mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 u = ld4<T>(a + 4 * i), v = ld4<T>(b + 4 * i);
    st4<T>(y + 4 * i, make_float4(u.x * v.x, u.y * v.y, u.z * v.z, u.w * v.w));
  }
}
__global__ void axpby_kernel(const float* __restrict__ a, float wa, const float* __restrict__ b, float wb, float* __restrict__ out,
                             int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = wa * a[i] + wb * b[i];
}

}  // namespace
}  // namespace afb

namespace afb {
// bf16 fast paths (layernorm_bf16.cu)
int layernorm_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int64_t rows, int D,
                       float eps, cudaStream_t st);
int layernorm_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, const void* dres,
                       void* dx, float* dgamma, float* dbeta, int64_t rows, int D, cudaStream_t st);
}  // namespace afb

using namespace afb;

#define DISPATCH_DT(dt, T, ...)                        \
  if ((dt) == AFB_BF16) { using T = bf16; __VA_ARGS__; } \
  else { using T = float; __VA_ARGS__; }

extern "C" int afb_cast(const void* src, int sd, void* dst, int dd, int64_t n, afb_stream s) {
  AFB_REQUIRE(src && dst && n >= 0, "cast: bad args");
  if (n == 0) return 0;
  AFB_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "cast: unaligned");
  const int g = grid_for(n / 4 + 1, kBlock);
  DISPATCH_DT(sd, S, DISPATCH_DT(dd, D, (cast_kernel<S, D><<<g, kBlock, 0, as_stream(s)>>>((const S*)src, (D*)dst, n))));
  return check_launch("cast");
}

extern "C" int afb_cast_transpose(const float* src, void* dst, int rows, int cols, afb_stream s) {
  AFB_REQUIRE(src && dst && rows > 0 && cols > 0, "cast_transpose: bad args");
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  cast_transpose_kernel<<<grid, block, 0, as_stream(s)>>>(src, (bf16*)dst, rows, cols);
  return check_launch("cast_transpose");
}

extern "C" int afb_conv_weight_pack(const float* w, void* fwd, void* bwd, int co, int ci, int k, afb_stream s) {
  AFB_REQUIRE(w && (fwd || bwd), "conv_weight_pack: bad args");
  conv_pack_kernel<<<grid_for((int64_t)co * ci * k, kBlock), kBlock, 0, as_stream(s)>>>(w, (bf16*)fwd, (bf16*)bwd, co, ci, k);
  return check_launch("conv_weight_pack");
}

extern "C" int afb_conv_dw_unpack(const float* tmp, float* dW, int co, int ci, int k, afb_stream s) {
  AFB_REQUIRE(tmp && dW && co > 0 && ci > 0 && k > 0, "conv_dw_unpack: bad args");
  conv_dw_unpack_kernel<<<grid_for((int64_t)co * ci * k, kBlock), kBlock, 0, as_stream(s)>>>(tmp, dW, co, ci, k);
  return check_launch("conv_dw_unpack");
}

extern "C" int afb_copy2d(const void* src, int sd, int64_t lds, void* dst, int dd, int64_t ldd, int64_t rows, int cols, afb_stream s) {
  AFB_REQUIRE(src && dst && rows > 0 && cols > 0, "copy2d: bad args");
  const int g = grid_for(rows * cols, kBlock);
  DISPATCH_DT(sd, S, DISPATCH_DT(dd, D, (copy2d_kernel<S, D><<<g, kBlock, 0, as_stream(s)>>>((const S*)src, lds, (D*)dst, ldd, rows, cols))));
  return check_launch("copy2d");
}

extern "C" int afb_split3(const float* src, void* dst, int64_t rows, int cols, int which, afb_stream s) {
  AFB_REQUIRE(src && dst, "split3: bad args");
  split3_kernel<<<grid_for(rows * cols, kBlock), kBlock, 0, as_stream(s)>>>(src, (bf16*)dst, rows, cols, which);
  return check_launch("split3");
}

template <typename TI, typename TO>
static int ln_fwd_dispatch(const void* x, const float* g, const float* b, void* y, float* mean, float* rstd, int64_t rows, int D,
                           float eps, cudaStream_t st) {
  const int grid = grid_for(rows, kBlock / 32, 148 * 8);
  switch (D / 128) {
    case 1: ln_fwd_kernel<TI, TO, 1><<<grid, kBlock, 0, st>>>((const TI*)x, g, b, (TO*)y, mean, rstd, rows, eps); break;
    case 2: ln_fwd_kernel<TI, TO, 2><<<grid, kBlock, 0, st>>>((const TI*)x, g, b, (TO*)y, mean, rstd, rows, eps); break;
    case 4: ln_fwd_kernel<TI, TO, 4><<<grid, kBlock, 0, st>>>((const TI*)x, g, b, (TO*)y, mean, rstd, rows, eps); break;
    default: set_error("layernorm: D=%d unsupported (128, 256 or 512)", D); return AFB_ERR_UNSUPPORTED;
  }
  return check_launch("layernorm_fwd");
}

extern "C" int afb_layernorm_fwd(const void* x, int xd, const float* gamma, const float* beta, void* y, int yd, float* mean,
                                 float* rstd, int64_t rows, int D, float eps, afb_stream s) {
  AFB_REQUIRE(x && gamma && beta && y && rows > 0, "layernorm_fwd: bad args");
  AFB_REQUIRE(D == 128 || D == 256 || D == 512, "layernorm: D=%d unsupported (128, 256 or 512)", D);
  cudaStream_t st = as_stream(s);
  if (xd == AFB_BF16 && yd == AFB_BF16 && (D == 256 || D == 512) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0)
    return layernorm_fwd_bf16(x, gamma, beta, y, mean, rstd, rows, D, eps, st);
  if (xd == AFB_BF16 && yd == AFB_BF16) return ln_fwd_dispatch<bf16, bf16>(x, gamma, beta, y, mean, rstd, rows, D, eps, st);
  if (xd == AFB_F32 && yd == AFB_BF16) return ln_fwd_dispatch<float, bf16>(x, gamma, beta, y, mean, rstd, rows, D, eps, st);
  if (xd == AFB_F32 && yd == AFB_F32) return ln_fwd_dispatch<float, float>(x, gamma, beta, y, mean, rstd, rows, D, eps, st);
  return ln_fwd_dispatch<bf16, float>(x, gamma, beta, y, mean, rstd, rows, D, eps, st);
}

template <typename T>
static int ln_bwd_dispatch(const void* dy, const void* x, const float* g, const float* mean, const float* rstd, const void* dres,
                           void* dx, float* dgamma, float* dbeta, int64_t rows, int D, cudaStream_t st) {
  const int grid = grid_for(rows, kBlock / 32, 148 * 4);
  switch (D / 128) {
    case 1: ln_bwd_kernel<T, T, T, T, 1><<<grid, kBlock, 0, st>>>((const T*)dy, (const T*)x, g, mean, rstd, (const T*)dres, (T*)dx, dgamma, dbeta, rows); break;
    case 2: ln_bwd_kernel<T, T, T, T, 2><<<grid, kBlock, 0, st>>>((const T*)dy, (const T*)x, g, mean, rstd, (const T*)dres, (T*)dx, dgamma, dbeta, rows); break;
    case 4: ln_bwd_kernel<T, T, T, T, 4><<<grid, kBlock, 0, st>>>((const T*)dy, (const T*)x, g, mean, rstd, (const T*)dres, (T*)dx, dgamma, dbeta, rows); break;
    default: set_error("layernorm: D=%d unsupported", D); return AFB_ERR_UNSUPPORTED;
  }
  return check_launch("layernorm_bwd");
}

extern "C" int afb_layernorm_bwd(const void* dy, int dyd, const void* x, int xd, const float* gamma, const float* mean,
                                 const float* rstd, const void* dres, int drd, void* dx, int dxd, float* dgamma, float* dbeta,
                                 int64_t rows, int D, afb_stream s) {
  AFB_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0, "layernorm_bwd: bad args");
  AFB_REQUIRE(D == 128 || D == 256 || D == 512, "layernorm: D=%d unsupported (128, 256 or 512)", D);
  AFB_REQUIRE(dyd == xd && xd == dxd && (dres == nullptr || drd == xd), "layernorm_bwd: all activations must share one dtype");
  if (xd == AFB_BF16 && (D == 256 || D == 512) && (((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0)
    return layernorm_bwd_bf16(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, as_stream(s));
  if (xd == AFB_BF16) return ln_bwd_dispatch<bf16>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, as_stream(s));
  return ln_bwd_dispatch<float>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, D, as_stream(s));
}

static bool col_shape_ok(int C) { return C % 4 == 0 && C >= 4; }
static dim3 col_grid(int64_t M, int C) {
  const int col_tiles = ceil_div(C, kColTile);
  // every block ends with one atomic per column: ~4 blocks per SM keep the same-address atomics (fp64 for the BatchNorm
  // statistics) off the critical path -- 1184 blocks spent most of a 35 us launch queueing 1184 atomics per address
  int row_blocks = grid_for(M, kRowLanes * 16, (148 * 4) / col_tiles > 0 ? (148 * 4) / col_tiles : 1);
  return dim3(row_blocks, col_tiles);
}

extern "C" int afb_colstats(const void* x, int dt, int64_t M, int C, int ldx, double* sum, double* sumsq, afb_stream s) {
  AFB_REQUIRE(x && sum && sumsq && M > 0, "colstats: bad args");
  AFB_REQUIRE(col_shape_ok(C) && ldx % 4 == 0, "colstats: C=%d unsupported", C);
  DISPATCH_DT(dt, T, (colstats_kernel<T><<<col_grid(M, C), kBlock, 0, as_stream(s)>>>((const T*)x, M, C, ldx, sum, sumsq)));
  return check_launch("colstats");
}

extern "C" int afb_colsum(const void* x, int dt, int64_t M, int C, int ldx, const float* row_scale, int div, float* out,
                          afb_stream s) {
  AFB_REQUIRE(x && out && M > 0, "colsum: bad args");
  AFB_REQUIRE(col_shape_ok(C) && ldx % 4 == 0, "colsum: C=%d unsupported", C);
  DISPATCH_DT(dt, T, (colsum_kernel<T><<<col_grid(M, C), kBlock, 0, as_stream(s)>>>((const T*)x, M, C, ldx, row_scale, div > 0 ? div : 1, out)));
  return check_launch("colsum");
}

extern "C" int afb_bn_finalize(const double* sum, const double* sumsq, int64_t M, int C, const float* gamma, const float* beta,
                               float* rm, float* rv, float momentum, float eps, int training, float* mean, float* rstd,
                               float* scale, float* shift, afb_stream s) {
  AFB_REQUIRE(gamma && beta && scale && shift, "bn_finalize: bad args");
  AFB_REQUIRE(training ? (sum && sumsq) : (rm && rv), "bn_finalize: missing statistics");
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(s)>>>(sum, sumsq, M, C, gamma, beta, rm, rv, momentum, eps, training,
                                                                 mean, rstd, scale, shift);
  return check_launch("bn_finalize");
}

extern "C" int afb_bn_act_fwd(const void* x, int xd, const float* scale, const float* shift, const void* res_pre,
                              const void* res_post, int rd, int relu, void* y, void* y2, int yd, int64_t M, int C, int T, int V,
                              afb_stream s) {
  AFB_REQUIRE(x && scale && shift && (y || y2) && M > 0 && C % 4 == 0, "bn_act_fwd: bad args");
  // one dtype for everything, or the "exact mask" layout of the bf16 mode: fp32 pre-activation x, bf16 residual / outputs
  // (mixed2: the residual is itself an fp32 BatchNorm output -- unit_agcn's `down` branch -- and must reach the mask unrounded)
  const bool has_res = res_pre || res_post;
  const bool mixed2 = xd == AFB_F32 && yd == AFB_BF16 && has_res && rd == AFB_F32;
  const bool mixed = xd == AFB_F32 && yd == AFB_BF16 && !mixed2;
  AFB_REQUIRE((xd == yd || mixed || mixed2) && (!has_res || rd == yd || mixed2), "bn_act_fwd: unsupported dtype combination");
  AFB_REQUIRE(y2 == nullptr || (T > 0 && V > 0 && M % ((int64_t)T * V) == 0), "bn_act_fwd: permuted copy needs T,V");
  const int grid = grid_for(M * (C / 4), kBlock);
  if (mixed2) {
    bn_act_fwd_kernel<float, float, bf16><<<grid, kBlock, 0, as_stream(s)>>>((const float*)x, scale, shift, (const float*)res_pre,
                                                                             (const float*)res_post, relu, (bf16*)y, (bf16*)y2, M, C, T, V);
    return check_launch("bn_act_fwd");
  }
  if (mixed) {
    bn_act_fwd_kernel<float, bf16, bf16><<<grid, kBlock, 0, as_stream(s)>>>((const float*)x, scale, shift, (const bf16*)res_pre,
                                                                            (const bf16*)res_post, relu, (bf16*)y, (bf16*)y2, M, C, T, V);
    return check_launch("bn_act_fwd");
  }
  DISPATCH_DT(xd, T_, (bn_act_fwd_kernel<T_, T_, T_><<<grid, kBlock, 0, as_stream(s)>>>(
                          (const T_*)x, scale, shift, (const T_*)res_pre, (const T_*)res_post, relu, (T_*)y, (T_*)y2, M, C, T, V)));
  return check_launch("bn_act_fwd");
}

extern "C" int afb_bn_bwd_reduce(const void* dy, const void* dy2, int gd, const void* x, int xd, const void* res_pre, int rd,
                                 const float* mean, const float* rstd, const float* gamma, const float* beta, int relu,
                                 float* dgamma, float* dbeta, int64_t M, int C, int T, int V, afb_stream s) {
  AFB_REQUIRE((dy || dy2) && x && mean && rstd && gamma && beta && dgamma && dbeta && M > 0, "bn_bwd_reduce: bad args");
  AFB_REQUIRE(col_shape_ok(C), "bn_bwd_reduce: C=%d unsupported", C);
  const bool mixed2 = xd == AFB_F32 && gd == AFB_BF16 && res_pre != nullptr && rd == AFB_F32;
  const bool mixed = xd == AFB_F32 && gd == AFB_BF16 && !mixed2;
  AFB_REQUIRE((gd == xd || mixed || mixed2) && (res_pre == nullptr || rd == gd || mixed2), "bn_bwd_reduce: unsupported dtype combination");
  AFB_REQUIRE(dy2 == nullptr || (T > 0 && V > 0 && M % ((int64_t)T * V) == 0), "bn_bwd_reduce: permuted grad needs T,V");
  if (mixed2) {
    bn_bwd_reduce_kernel<bf16, float, float><<<col_grid(M, C), kBlock, 0, as_stream(s)>>>(
        (const bf16*)dy, (const bf16*)dy2, (const float*)x, (const float*)res_pre, mean, rstd, gamma, beta, relu, dgamma, dbeta, M, C, T, V);
    return check_launch("bn_bwd_reduce");
  }
  if (mixed) {
    bn_bwd_reduce_kernel<bf16, float, bf16><<<col_grid(M, C), kBlock, 0, as_stream(s)>>>(
        (const bf16*)dy, (const bf16*)dy2, (const float*)x, (const bf16*)res_pre, mean, rstd, gamma, beta, relu, dgamma, dbeta, M, C, T, V);
    return check_launch("bn_bwd_reduce");
  }
  DISPATCH_DT(xd, T_, (bn_bwd_reduce_kernel<T_, T_, T_><<<col_grid(M, C), kBlock, 0, as_stream(s)>>>(
                          (const T_*)dy, (const T_*)dy2, (const T_*)x, (const T_*)res_pre, mean, rstd, gamma, beta, relu, dgamma, dbeta,
                          M, C, T, V)));
  return check_launch("bn_bwd_reduce");
}

extern "C" int afb_bn_bwd_apply(const void* dy, const void* dy2, int gd, const void* x, int xd, const void* res_pre, int rd,
                                const float* mean, const float* rstd, const float* gamma, const float* beta, const float* dgamma,
                                const float* dbeta, int relu, int training, void* dx, void* dres, int od, int64_t M, int C, int T,
                                int V, afb_stream s) {
  AFB_REQUIRE((dy || dy2) && x && mean && rstd && gamma && beta && dgamma && dbeta && dx && M > 0 && C % 4 == 0,
              "bn_bwd_apply: bad args");
  const bool mixed2 = xd == AFB_F32 && gd == AFB_BF16 && res_pre != nullptr && rd == AFB_F32;
  const bool mixed = xd == AFB_F32 && gd == AFB_BF16 && !mixed2;
  AFB_REQUIRE((gd == xd || mixed || mixed2) && od == gd && (res_pre == nullptr || rd == gd || mixed2),
              "bn_bwd_apply: unsupported dtype combination");
  const int grid = grid_for(M * (C / 4), kBlock);
  if (mixed2) {
    bn_bwd_apply_kernel<bf16, float, float, bf16><<<grid, kBlock, 0, as_stream(s)>>>(
        (const bf16*)dy, (const bf16*)dy2, (const float*)x, (const float*)res_pre, mean, rstd, gamma, beta, dgamma, dbeta, relu, training,
        (bf16*)dx, (bf16*)dres, M, C, T, V);
    return check_launch("bn_bwd_apply");
  }
  if (mixed) {
    bn_bwd_apply_kernel<bf16, float, bf16, bf16><<<grid, kBlock, 0, as_stream(s)>>>(
        (const bf16*)dy, (const bf16*)dy2, (const float*)x, (const bf16*)res_pre, mean, rstd, gamma, beta, dgamma, dbeta, relu, training,
        (bf16*)dx, (bf16*)dres, M, C, T, V);
    return check_launch("bn_bwd_apply");
  }
  DISPATCH_DT(xd, T_, (bn_bwd_apply_kernel<T_, T_, T_, T_><<<grid, kBlock, 0, as_stream(s)>>>(
                          (const T_*)dy, (const T_*)dy2, (const T_*)x, (const T_*)res_pre, mean, rstd, gamma, beta, dgamma, dbeta, relu,
                          training, (T_*)dx, (T_*)dres, M, C, T, V)));
  return check_launch("bn_bwd_apply");
}

extern "C" int afb_pool_mean_fwd(const void* x, void* y, int dt, int64_t B, int L, int D, afb_stream s) {
  AFB_REQUIRE(x && y && B > 0 && L > 0 && D % 2 == 0, "pool_mean_fwd: bad args");
  if (dt == AFB_BF16 && D % 8 == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    pool_mean_fwd_bf16x8_kernel<<<grid_for(B * D / 8, kBlock), kBlock, 0, as_stream(s)>>>((const bf16*)x, (bf16*)y, B, L, D);
    return check_launch("pool_mean_fwd");
  }
  DISPATCH_DT(dt, T, (pool_mean_fwd_kernel<T><<<grid_for(B * D / 2, kBlock), kBlock, 0, as_stream(s)>>>((const T*)x, (T*)y, B, L, D)));
  return check_launch("pool_mean_fwd");
}
extern "C" int afb_pool_mean_bwd(const void* dy, void* dx, int dt, int64_t B, int L, int D, afb_stream s) {
  AFB_REQUIRE(dy && dx && B > 0 && L > 0 && D % 2 == 0, "pool_mean_bwd: bad args");
  if (dt == AFB_BF16 && D % 8 == 0 && (((uintptr_t)dy | (uintptr_t)dx) & 15) == 0) {
    pool_mean_bwd_bf16x8_kernel<<<grid_for(B * D / 8, kBlock), kBlock, 0, as_stream(s)>>>((const bf16*)dy, (bf16*)dx, B, L, D);
    return check_launch("pool_mean_bwd");
  }
  DISPATCH_DT(dt, T, (pool_mean_bwd_kernel<T><<<grid_for(B * L * D / 2, kBlock), kBlock, 0, as_stream(s)>>>((const T*)dy, (T*)dx, B, L, D)));
  return check_launch("pool_mean_bwd");
}
extern "C" int afb_pool_max_fwd(const void* x, void* y, int32_t* arg, int dt, int64_t B, int L, int D, afb_stream s) {
  AFB_REQUIRE(x && y && arg && B > 0 && L > 0, "pool_max_fwd: bad args");
  DISPATCH_DT(dt, T, (pool_max_fwd_kernel<T><<<grid_for(B * D, kBlock), kBlock, 0, as_stream(s)>>>((const T*)x, (T*)y, arg, B, L, D)));
  return check_launch("pool_max_fwd");
}
extern "C" int afb_pool_max_bwd(const void* dy, const int32_t* arg, void* dx, int dt, int64_t B, int L, int D, afb_stream s) {
  AFB_REQUIRE(dy && dx && arg && B > 0 && L > 0, "pool_max_bwd: bad args");
  DISPATCH_DT(dt, T, (pool_max_bwd_kernel<T><<<grid_for(B * L * D, kBlock), kBlock, 0, as_stream(s)>>>((const T*)dy, arg, (T*)dx, B, L, D)));
  return check_launch("pool_max_bwd");
}

extern "C" int afb_softmax_ce(const float* logits, const int64_t* labels, float* loss, float* dlogits, int N, int C, afb_stream s) {
  AFB_REQUIRE(logits && labels && N > 0 && C > 0, "softmax_ce: bad args");
  softmax_ce_kernel<<<ceil_div((int64_t)N * 32, kBlock), kBlock, 0, as_stream(s)>>>(logits, labels, loss, dlogits, N, C);
  return check_launch("softmax_ce");
}

extern "C" int afb_adamw(float* p, const float* g, float* m, float* v, void* pb, int64_t n, const int32_t* step, float lr,
                         float b1, float b2, float eps, float wd, float grad_scale, afb_stream s) {
  AFB_REQUIRE(p && g && m && v && step && n > 0 && n % 4 == 0, "adamw: bad args (n must be a multiple of 4)");
  adamw_kernel<<<grid_for(n / 4, kBlock), kBlock, 0, as_stream(s)>>>(p, g, m, v, (bf16*)pb, n, step, lr, b1, b2, eps, wd, grad_scale);
  return check_launch("adamw");
}
extern "C" int afb_step_inc(int32_t* step, afb_stream s) {
  AFB_REQUIRE(step, "step_inc: null");
  step_inc_kernel<<<1, 1, 0, as_stream(s)>>>(step);
  return check_launch("step_inc");
}

extern "C" int afb_scale_rows(const void* x, void* y, int dt, int64_t M, int C, const float* rs, int div, afb_stream s) {
  AFB_REQUIRE(x && y && rs && M > 0 && C > 0 && div > 0, "scale_rows: bad args");
  if (C % 4 == 0) {
    DISPATCH_DT(dt, T, (scale_rows_kernel<T><<<grid_for(M * C / 4, kBlock), kBlock, 0, as_stream(s)>>>((const T*)x, (T*)y, M, C, rs, div)));
  } else {
    DISPATCH_DT(dt, T, (scale_rows_any_kernel<T><<<grid_for(M * C, kBlock), kBlock, 0, as_stream(s)>>>((const T*)x, (T*)y, M, C, rs, div)));
  }
  return check_launch("scale_rows");
}

extern "C" int afb_bone_stream(const float* x, const int32_t* parent, float* y, int64_t NT, int V, afb_stream s) {
  AFB_REQUIRE(x && parent && y && NT > 0, "bone_stream: bad args");
  bone_kernel<<<grid_for(NT * V * 3, kBlock), kBlock, 0, as_stream(s)>>>(x, parent, y, NT, V);
  return check_launch("bone_stream");
}
extern "C" int afb_motion_stream(const float* x, float* y, int N, int T, int V, afb_stream s) {
  AFB_REQUIRE(x && y && N > 0, "motion_stream: bad args");
  motion_kernel<<<grid_for((int64_t)N * T * V * 3, kBlock), kBlock, 0, as_stream(s)>>>(x, y, N, T, V);
  return check_launch("motion_stream");
}
// palm-centre normalisation (Hand_Dataset.py:61): every joint of every frame minus joint 1 of frame 0 of its sequence
__global__ void __launch_bounds__(kBlock) palm_center_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total,
                                                             int64_t per_seq, int joint) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / per_seq;
    const int c = (int)(i % 3);
    y[i] = x[i] - __ldg(x + n * per_seq + joint * 3 + c);
  }
}
// Hand_Dataset.data_aug (data_process/Hand_Dataset.py:84-157) for a whole batch on the device: one of four transforms per
// sample, chosen and parameterised by the caller (kind[n], params[n][16]):
//   0 scale            y = x * p[0]                                            (:86-96,  factor ~ U(0.8, 1.2))
//   1 shift            y = x + p[0..2]                                         (:98-107, offset ~ U(-0.1, 0.1)^3)
//   2 noise            y[:, j_k, :] = x + p[4 + 3k .. 6 + 3k] for the four joints j_k = p[k]   (:109-123)
//   3 time_interpolate y[t] = x[t] + p[0] (x[t+1] - x[t]) for t < T - 1, last frame = the one before it   (:125-142)
//   anything else      y = x
__global__ void augment_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int T, int V,
                               const int* __restrict__ kind, const float* __restrict__ params) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % 3);
  const int v = (int)((i / 3) % V);
  const int t = (int)((i / (3 * (int64_t)V)) % T);
  const int64_t n = i / (3 * (int64_t)V * T);
  const float* p = params + n * 16;
  const int k = kind[n];
  float val = x[i];
  if (k == 0) {
    val *= p[0];
  } else if (k == 1) {
    val += p[c];
  } else if (k == 2) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if ((int)p[q] == v) val += p[4 + 3 * q + c];
  } else if (k == 3 && T > 1) {
    const int ts = t < T - 1 ? t : T - 2;                 // the padded last frame repeats frame T - 2's result
    const int64_t j = i + (int64_t)(ts - t) * V * 3;
    const float a = x[j], b = x[j + (int64_t)V * 3];
    val = a + (b - a) * p[0];
  }
  y[i] = val;
}
extern "C" int afb_augment(const float* x, float* y, int64_t N, int T, int V, const int* kind, const float* params, afb_stream s) {
  AFB_REQUIRE(x && y && x != y && kind && params && N > 0 && T > 0 && V > 0, "augment: bad args (out of place only)");
  const int64_t total = N * T * V * 3;
  augment_kernel<<<grid_for(total, kBlock), kBlock, 0, as_stream(s)>>>(x, y, total, T, V, kind, params);
  return check_launch("augment");
}
extern "C" int afb_palm_center(const float* x, float* y, int N, int T, int V, int joint, afb_stream s) {
  AFB_REQUIRE(x && y && x != y && N > 0 && T > 0 && joint >= 0 && joint < V, "palm_center: bad args (out of place only)");
  const int64_t per_seq = (int64_t)T * V * 3;
  palm_center_kernel<<<grid_for(N * per_seq, kBlock), kBlock, 0, as_stream(s)>>>(x, y, N * per_seq, per_seq, joint);
  return check_launch("palm_center");
}
extern "C" int afb_frame_phase(const void* src, void* dst, int dt, int scatter, int N, int T, int To, int V, int C, int stride, int phase,
                               afb_stream s) {
  AFB_REQUIRE(src && dst && N > 0 && T > 0 && To > 0 && V > 0 && C % 4 == 0 && stride >= 1 && phase >= 0 && phase < stride,
              "frame_phase: bad args");
  const int g = grid_for((int64_t)N * To * V * (C / 4), kBlock);
  if (scatter) {
    DISPATCH_DT(dt, T_, (frame_phase_kernel<T_, true><<<g, kBlock, 0, as_stream(s)>>>((const T_*)src, (T_*)dst, N, T, To, V, C, stride, phase)));
  } else {
    DISPATCH_DT(dt, T_, (frame_phase_kernel<T_, false><<<g, kBlock, 0, as_stream(s)>>>((const T_*)src, (T_*)dst, N, T, To, V, C, stride, phase)));
  }
  return check_launch("frame_phase");
}
extern "C" int afb_mul(const void* a, const void* b, void* y, int dt, int64_t n, afb_stream s) {
  AFB_REQUIRE(a && b && y && n > 0 && n % 4 == 0, "mul: bad args (n must be a multiple of 4)");
  DISPATCH_DT(dt, T, (mul_kernel<T><<<grid_for(n / 4, kBlock), kBlock, 0, as_stream(s)>>>((const T*)a, (const T*)b, (T*)y, n / 4)));
  return check_launch("mul");
}
extern "C" int afb_axpby(const float* a, float wa, const float* b, float wb, float* out, int64_t n, afb_stream s) {
  AFB_REQUIRE(a && b && out && n > 0, "axpby: bad args");
  axpby_kernel<<<grid_for(n, kBlock), kBlock, 0, as_stream(s)>>>(a, wa, b, wb, out, n);
  return check_launch("axpby");
}
