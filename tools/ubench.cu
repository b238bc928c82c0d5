// Micro-benchmarks that size the gcn0 fused kernel (run on the B200 box: tools/ubench):
//   1. mma.sync.m16n8k16 bf16 issue rate / latency per SM (legacy tensor path used by the warp-level kernels)
//   2. write-only streaming of a 46 MB tensor: st.global.v4 vs 1-D bulk store (UBLKCP) from shared memory
//   3. cost of a grid-wide barrier (one atomic ticket + spin) in a cooperative launch of 148 x k CTAs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int ILP>
__global__ void mma_rate(float* out, int iters) {
  float acc[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma(acc[i], a0, a1, a2, a3, b0, b1);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 12345.678f) out[0] = s;
}

__global__ void store_v4(uint4* dst, size_t n16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}

// every warp owns a 4 KB staging buffer, fills it, and hands it to the bulk-copy engine; chunk = 4 KB of the output
__global__ void store_bulk(uint8_t* dst, size_t nchunks, int fill) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* buf = sm + warp * 4096;
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(buf);
  for (size_t c = (size_t)blockIdx.x * nw + warp; c < nchunks; c += (size_t)gridDim.x * nw) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    if (fill) {
#pragma unroll
      for (int q = 0; q < 8; ++q) reinterpret_cast<uint4*>(buf)[q * 32 + lane] = make_uint4(lane, q, (uint32_t)c, 1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst + c * 4096), "r"(saddr) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void grid_barrier_loop(unsigned int* counter, int rounds, long long* cycles) {
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int target = (unsigned int)(r + 1) * gridDim.x;
      atomicAdd(counter, 1u);
      while (*((volatile unsigned int*)counter) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1e3);
  float* out; CK(cudaMalloc(&out, 1024));
  const int iters = 4096;
  for (int warps = 4; warps <= 32; warps *= 2) {
    const int grid = prop.multiProcessorCount;
    float ms1 = time_ms([&] { mma_rate<1><<<grid, warps * 32>>>(out, iters); }, 3);
    float ms4 = time_ms([&] { mma_rate<4><<<grid, warps * 32>>>(out, iters / 4); }, 3);
    float ms8 = time_ms([&] { mma_rate<8><<<grid, warps * 32>>>(out, iters / 8); }, 3);
    const double mmas = (double)iters * warps;   // per SM
    printf("mma.sync m16n8k16 bf16, %2d warps/SM: ILP1 %.3f ms (%.1f ns/mma/SM)  ILP4 %.3f ms (%.2f ns/mma/SM)  ILP8 %.3f ms (%.2f ns/mma/SM = %.0f TFLOP/s chip)\n",
           warps, ms1, ms1 * 1e6 / mmas, ms4, ms4 * 1e6 / mmas, ms8, ms8 * 1e6 / mmas, 2.0 * 4096 * mmas * grid / (ms8 * 1e-3) / 1e12);
  }
  CK(cudaGetLastError());
  const size_t bytes = 256ull * 32 * 22 * 128 * 2;   // 46.1 MB (gcn0 output at cfg2)
  uint8_t* dst; CK(cudaMalloc(&dst, bytes));
  uint8_t* flush; CK(cudaMalloc(&flush, 512u << 20));
  for (int rep = 0; rep < 2; ++rep) {
    for (int cps = 1; cps <= 8; cps *= 2) {
      cudaMemset(flush, rep, 512u << 20);
      float ms = time_ms([&] { store_v4<<<148 * cps, 256>>>((uint4*)dst, bytes / 16); }, 5);
      printf("store_v4   %d CTAs/SM x256 thr: %.2f us = %.0f GB/s\n", cps, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
      for (int cps = 1; cps <= 2; ++cps) {
        for (int fill = 0; fill <= 1; ++fill) {
          CK(cudaFuncSetAttribute(store_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, warps * 4096));
          float ms = time_ms([&] { store_bulk<<<148 * cps, warps * 32, warps * 4096>>>(dst, bytes / 4096, fill); }, 5);
          printf("store_bulk %d CTAs/SM x%2d warps fill=%d: %.2f us = %.0f GB/s\n", cps, warps, fill, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
        }
      }
    }
  }
  CK(cudaGetLastError());
  unsigned int* counter; CK(cudaMalloc(&counter, 4));
  long long* cyc; CK(cudaMalloc(&cyc, 8));
  for (int cps = 1; cps <= 2; ++cps) {
    const int rounds = 200;
    CK(cudaMemset(counter, 0, 4));
    int grid = 148 * cps, threads = 256;
    void* args[] = {&counter, (void*)&rounds, &cyc};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchCooperativeKernel((void*)grid_barrier_loop, dim3(grid), dim3(threads), args, 0, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("grid barrier, %d CTAs: %.2f us per barrier (%lld cycles)\n", grid, ms * 1e3 / rounds, h / rounds);
  }
  // empty-kernel launch + event overhead
  {
    float ms = time_ms([&] { mma_rate<1><<<148, 256>>>(out, 0); }, 200);
    printf("empty kernel back-to-back: %.2f us per launch\n", ms * 1e3);
  }
  printf("done\n");
  return 0;
}
