// tcgen05 / TMEM / TMA GEMMs for sm_100a.
//
//   gemm_tn_kernel : C[M,N] = epilogue(A[M,K] * B[N,K]^T)   nn.Linear forward, dX, and the 9x1
//                    temporal conv as an implicit GEMM (per-tap row-shifted TMA loads).
//   gemm_dw_kernel : dW[N1,N2] += G[M,N1]^T * X[M,N2]       weight gradients, contraction over
//                    token rows (both operands MN-major), split over CTAs, fp32 atomics.
//
// Structure of gemm_tn_kernel (one CTA per SM, persistent over tiles; 10 warps):
//   warp 0    TMA producer  : cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx; in B-stationary
//                             mode the [BN x K] weight panel is loaded once and only A tiles stream
//   warp 1    MMA issuer    : one elected lane issues tcgen05.mma (M=128, N=BN, K=16) into one of
//                             two TMEM accumulator stages; tcgen05.commit releases smem / signals
//   warps 2-9 epilogue      : thread = output row (tcgen05.ld 32x32b.x32), fused epilogue (bias, pos-embed, GELU,
//                             GELU', DropPath scale, residual) -> 128B-swizzled staging -> TMA store; specialised
//                             straight-line variants for the hot bf16 combinations (epi_fast_loop)
// gemm_dw_kernel: producer + MMA issuer + 4 warps that sum the bias gradient from the G tiles while the MMAs run and
// then drain the accumulators with vector reductions.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace afb {
namespace {

constexpr int BM = 128;       // rows per tile == TMEM lanes
constexpr int BK = 64;        // bf16 elements per k-block row == 128 bytes == swizzle span
constexpr int kThreads = 192;      // dW kernel: producer + MMA + 4 epilogue warps
constexpr int kEpiWarps = 4;
constexpr int kTnEpiWarps = 8;     // TN kernel: two epilogue warps per TMEM lane quarter
constexpr int kTnThreads = 64 + 32 * kTnEpiWarps;
constexpr int kScratchStride = 36;  // floats; 144 B rows keep float4 writes conflict-free
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;  // 4 s: a protocol bug traps instead of hanging

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (visible as a launch error) instead of hanging the GPU.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait(bar, parity)) return;
    if ((spins & 1023u) == 0 && global_ns() - t0 > kWaitTimeoutNs) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane/row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// same load without the wait: lets several loads (and global loads) be in flight before one tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 format): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 |
// SWIZZLE_128B (2) <<61.   K-major tiles: rows of 128 B, 8-row groups 1024 B apart (SBO).
// MN-major tiles: 64-element (128 B) MN runs, 8 k-rows per 1024 B atom (SBO), next 64 MN at LBO.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D fp32, A/B bf16, majors, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 32 consecutive elements of one row <-> registers, 16-byte vector accesses (row start 16-byte aligned)
__device__ __forceinline__ void load32_dyn(const void* base, int64_t idx, int dt, float (&o)[32]) {
  if (dt == AFB_BF16) {
    const uint4* ptr = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + idx);
    uint4 raw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) raw[q] = ptr[q];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {raw[q].x, raw[q].y, raw[q].z, raw[q].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        o[8 * q + 2 * j] = __uint_as_float(w[j] << 16);
        o[8 * q + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  } else {
    const float4* ptr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 f = ptr[q];
      o[4 * q] = f.x; o[4 * q + 1] = f.y; o[4 * q + 2] = f.z; o[4 * q + 3] = f.w;
    }
  }
}
__device__ __forceinline__ void store32_dyn(void* base, int64_t idx, int dt, const float (&v)[32]) {
  if (dt == AFB_BF16) {
    uint4* ptr = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(base) + idx);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
        w[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      ptr[q] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  } else {
    float4* ptr = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
#pragma unroll
    for (int q = 0; q < 8; ++q) ptr[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}
__device__ __forceinline__ void prefetch_row(const void* base, int64_t idx, int dt, int ncols) {
  const int esz = dt == AFB_BF16 ? 2 : 4;
  const char* ptr = reinterpret_cast<const char*>(base) + idx * esz;
  for (int off = 0; off < ncols * esz; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + off));
}

// TMA store of one staged [32 rows x 128 bytes] box (128B-swizzled smem) into a 3-D tensor (cols, rows, batch)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Shared-memory plan of the TN kernel.  STAT ("B-stationary"): the whole [BN x K] weight panel stays resident
// (<= 128 KB) and only A tiles stream through the ring -- cuts the L2->SM fill per tile by 3x for the
// K <= 256/512 linears whose weight panel would otherwise be re-fetched for every 128-row tile.
template <int BN, bool STAT> struct TnCfg {
  static constexpr int kStages = STAT ? 8 : (BN >= 256 ? 4 : (BN >= 128 ? 6 : 8));   // STAT: upper bound, see TnArgs::stages
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;                       // one k-block of the B panel
  static constexpr int kResKBlocks = (128 * 1024) / kBBytes;        // resident k-blocks (STAT)
  static constexpr int kBRegion = STAT ? 128 * 1024 : kStages * kBBytes;
  static constexpr int kStageTx = STAT ? kABytes : kABytes + kBBytes;
  static constexpr int kTmemCols = BN >= 256 ? 512 : (BN >= 128 ? 256 : 128);
  static constexpr int kStagingBytes = kTnEpiWarps * 4096;          // one 32 x 128 B box per epilogue warp
  static constexpr int kFixedBytes = kStagingBytes + 256 /*barriers*/ + 1024 /*align*/;
  static constexpr int kSmemMax = 232448;  // 227 KB opt-in limit per CTA
  // streaming: fixed ring.  stationary: the ring takes whatever the resident panel leaves (host: stat_stages()).
  static constexpr int kSmemBytes = STAT ? kSmemMax : kStages * kABytes + kBRegion + kFixedBytes;
  static int stat_stages(int k_blocks) {
    int st = (kSmemMax - kFixedBytes - k_blocks * kBBytes) / kABytes;
    return st > 8 ? 8 : st;
  }
};

struct TnArgs {
  int m_tiles_per_batch, n_tiles, total_m_tiles, k_blocks;
  int stages;   // ring depth actually used (stationary mode: as many A stages as fit beside the weight panel)
  int rev;      // visit the m-tiles in descending order (ping-pong traversal, api.cu:next_stream_dir)
  int rows_per_batch, batches, N;
  int kb_per_tap, tap_row_stride, tap_pad;
  int out_dtype, act, has_c2;
  int epi_kind;   // EPI_*: compile-time specialised epilogue for the hot bf16 combinations, else EPI_GENERIC
  float alpha;
  const float* bias;
  const float* pos;
  int pos_rows;
  const void* aux;
  int aux_dtype, ldaux;
  const void* residual;
  int res_dtype, ldres;
  const float* row_scale;
  int row_scale_div;
  int rs_bias;   // AFB_RS_BIAS: the row scale multiplies only the bias term (A rows are pre-scaled)
};

// Tile walk of one CTA.  Streaming: tiles blockIdx.x, +grid, ... over (m-tile, n-block) with n fastest, so
// CTAs working on the same A tile run side by side (A is fetched from HBM once, then hits L2).  Stationary:
// the CTA keeps n-block (blockIdx.x % n_tiles) and walks m-tiles with stride grid / n_tiles.
template <bool STAT>
struct TileWalk {
  // k = position of the m-tile in visiting order, mt = the m-tile itself (descending when p.rev: see next_stream_dir())
  int n_blk, mt, k, k_stride, k_end, n_tiles, lin, lin_stride, lin_end, last;
  __device__ __forceinline__ TileWalk(const TnArgs& p) {
    n_tiles = p.n_tiles;
    last = p.rev ? p.total_m_tiles - 1 : -1;
    if (STAT) {
      n_blk = blockIdx.x % p.n_tiles;
      k = blockIdx.x / p.n_tiles;
      k_stride = gridDim.x / p.n_tiles;
      k_end = p.total_m_tiles;
    } else {
      lin = blockIdx.x;
      lin_stride = gridDim.x;
      lin_end = p.total_m_tiles * p.n_tiles;
      n_blk = lin % n_tiles;
      k = lin / n_tiles;
    }
    mt = last >= 0 ? last - k : k;
  }
  __device__ __forceinline__ bool valid() const { return STAT ? k < k_end : lin < lin_end; }
  __device__ __forceinline__ void next() {
    if (STAT) {
      k += k_stride;
    } else {
      lin += lin_stride;
      n_blk = lin % n_tiles;
      k = lin / n_tiles;
    }
    mt = last >= 0 ? last - k : k;
  }
};

// Specialised epilogues.  The runtime-flag epilogue in the kernel body costs ~290 instructions per 32-column
// chunk (ncu r01: 11% FFMA, >40% flag tests / branches / uniform moves) and, with two epilogue warps per
// scheduler, made every K <= 512 linear epilogue-issue bound.  The hot bf16 combinations get a straight-line
// version: one 128-byte output unit (64 columns) per step, both TMEM loads and the residual / aux row loads in
// flight together, no per-chunk flag tests.
enum { EPI_GENERIC = 0, EPI_BIAS, EPI_BIAS_RES, EPI_BIAS_GELU_C2, EPI_GELU_BWD, EPI_PLAIN };

__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float* o) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    o[2 * j] = __uint_as_float(w[j] << 16);
    o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// 8 consecutive columns of this thread's row -> bf16 -> 16-byte chunk q of the 128B-swizzled staging row
__device__ __forceinline__ void sts_pack8(uint8_t* rowp, int q, int lane, const float* x) {
  uint32_t wv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
    wv[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
  *reinterpret_cast<uint4*>(rowp + ((q ^ (lane & 7)) * 16)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
}
// the previous TMA store of this warp must have finished reading the staging buffer before it is rewritten
__device__ __forceinline__ void staging_acquire(int lane) {
  if (lane == 0) bulk_wait_read0();
  __syncwarp();
}
// hand the staged [32 rows x 128 B] unit to the TMA store engine
__device__ __forceinline__ void staging_store(uint32_t stage_addr, int lane, const CUtensorMap* map, int col, int row0, int batch) {
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(map, stage_addr, col, row0, batch);
    bulk_commit();
  }
}

// The whole epilogue-warp loop for one specialised kind.  Residual / aux rows are software-pipelined one output
// unit ahead in registers (raw_next), on top of the L2 prefetch one tile ahead: the row loads of a unit are in
// flight while the previous unit is converted, staged and stored, instead of being waited for right after issue.
template <int BN, bool STAT, int KIND>
__device__ __forceinline__ void epi_fast_loop(const TnArgs& p, const CUtensorMap* tmC, const CUtensorMap* tmC2, uint32_t tmem_base,
                                              uint32_t tfull0, uint32_t tempty0, uint8_t* stage_ptr, uint32_t stage_addr, int lane,
                                              int lane_grp, int half) {
  constexpr bool kBias = KIND == EPI_BIAS || KIND == EPI_BIAS_RES || KIND == EPI_BIAS_GELU_C2;
  constexpr bool kRes = KIND == EPI_BIAS_RES;
  constexpr bool kAux = KIND == EPI_GELU_BWD;
  constexpr bool kRow = kRes || kAux;   // reads one bf16 row operand per output element
  constexpr bool kScale = KIND == EPI_BIAS_RES || KIND == EPI_GELU_BWD || KIND == EPI_PLAIN;
  constexpr int UNITS = BN / 64;
  uint8_t* rowp = stage_ptr + lane * 128;
  const bf16* rsrc = reinterpret_cast<const bf16*>(kRes ? p.residual : p.aux);
  const int ldr = kRes ? p.ldres : p.ldaux;

  // Row operand (residual / aux) of unit u of tile t, loaded COALESCED: load i of lane l covers 16-byte chunk (l & 7) of
  // row 4 i + (l >> 3) of the warp's 32 rows, so one warp-wide LDG.128 reads four full 128-byte lines (thread-per-row
  // loads touched 32 different lines per instruction and capped these GEMMs well below the qkv-type ones).  The chunks
  // are handed to their row's thread through the warp's staging buffer right before use.
  // (Only for the residual kinds: their math before the addition is short.  GELU_BWD keeps thread-per-row loads, its
  // long per-element math must not wait for the staging buffer -- measured +20 % when routed through it.)
  constexpr bool kViaStage = kRes;
  const int sub_row = lane >> 3, sub_chunk = lane & 7;
  auto row_ptr = [&](const TileWalk<STAT>& t, int u) -> const uint4* {   // thread-per-row variant
    const int rl = (t.mt % p.m_tiles_per_batch) * BM + lane_grp * 32 + lane;
    if (rl >= p.rows_per_batch) return nullptr;
    const int64_t r = (int64_t)(t.mt / p.m_tiles_per_batch) * p.rows_per_batch + rl;
    return reinterpret_cast<const uint4*>(rsrc + r * ldr + t.n_blk * BN + u * 64);
  };
  auto unit_base = [&](const TileWalk<STAT>& t, int u, int& rows_valid) -> const bf16* {
    const int rl0 = (t.mt % p.m_tiles_per_batch) * BM + lane_grp * 32;
    rows_valid = p.rows_per_batch - rl0;
    const int64_t r0 = (int64_t)(t.mt / p.m_tiles_per_batch) * p.rows_per_batch + rl0;
    return rsrc + r0 * ldr + t.n_blk * BN + u * 64 + sub_chunk * 8;
  };
  auto l2_prefetch = [&](const TileWalk<STAT>& t) {
    if (!kRow) return;
    for (int u = half; u < UNITS; u += 2) {
      if (!kViaStage) {
        const uint4* ptr = row_ptr(t, u);
        if (ptr != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        continue;
      }
      int rv;
      const bf16* b0 = unit_base(t, u, rv);
      if (sub_chunk == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (4 * i + sub_row < rv) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + (int64_t)(4 * i + sub_row) * ldr));
      }
    }
  };
  auto issue = [&](uint4* r, const TileWalk<STAT>& t, int u) {
    if (!kViaStage) {
      const uint4* ptr = row_ptr(t, u);
#pragma unroll
      for (int q = 0; q < 8; ++q) r[q] = ptr != nullptr ? __ldg(ptr + q) : make_uint4(0u, 0u, 0u, 0u);
      return;
    }
    int rv;
    const bf16* b0 = unit_base(t, u, rv);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      r[i] = 4 * i + sub_row < rv ? __ldg(reinterpret_cast<const uint4*>(b0 + (int64_t)(4 * i + sub_row) * ldr)) : make_uint4(0u, 0u, 0u, 0u);
  };

  TileWalk<STAT> w(p);
  uint4 raw_next[kRow ? 8 : 1];
  if (kRow && w.valid() && half < UNITS) {
    l2_prefetch(w);
    issue(raw_next, w, half);
  }
  int acc = 0;
  uint32_t acc_phase = 0;
  for (; w.valid(); w.next()) {
    const int batch = w.mt / p.m_tiles_per_batch;
    const int row_in_batch0 = (w.mt % p.m_tiles_per_batch) * BM + lane_grp * 32;
    const bool valid = row_in_batch0 + lane < p.rows_per_batch;
    const int64_t row = (int64_t)batch * p.rows_per_batch + row_in_batch0 + lane;
    const int ncol0 = w.n_blk * BN;
    TileWalk<STAT> wn = w;
    wn.next();
    if (wn.valid()) l2_prefetch(wn);
    float rscale = 1.f;
    if ((kScale || KIND == EPI_BIAS_GELU_C2) && p.row_scale != nullptr && valid) rscale = p.row_scale[row / p.row_scale_div];
    const float bscale = p.rs_bias ? rscale : 1.f;    // scale of the bias term
    const float vscale = p.rs_bias ? 1.f : rscale;    // scale of the finished value
    const uint32_t tmem_tile = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN);
    mbar_wait(tfull0 + 8u * acc, acc_phase);
    tc_fence_after();
    bool arrived = false;
#pragma unroll 1
    for (int u = half; u < UNITS; u += 2) {
      const int n0 = ncol0 + u * 64;
      uint4 raw[kRow ? 8 : 1];
      if (kRow) {
#pragma unroll
        for (int q = 0; q < 8; ++q) raw[q] = raw_next[q];
        if (u + 2 < UNITS) issue(raw_next, w, u + 2);
        else if (wn.valid()) issue(raw_next, wn, half);
      }
      float pre[KIND == EPI_BIAS_GELU_C2 ? 64 : 1];
#pragma unroll
      for (int c = 0; c < 2; ++c) {   // two 32-column TMEM loads per unit keep the live register set small
        uint32_t av[32];
        tmem_ld32_issue(tmem_tile + (uint32_t)(u * 64 + c * 32), av);
        tmem_wait_ld();
        if (c == 1 && u + 2 >= UNITS) {   // this warp's last TMEM read of the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
          arrived = true;
        }
        if (c == 0) {
          staging_acquire(lane);
          if (kViaStage) {   // coalesced chunks -> staging (same 128B-swizzled layout as the output) -> own row below
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + sub_row;
              *reinterpret_cast<uint4*>(stage_ptr + r * 128 + ((sub_chunk ^ (r & 7)) << 4)) = raw[i];
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int q = c * 4 + q4;
          uint4 rowq;
          if (kViaStage) rowq = *reinterpret_cast<const uint4*>(rowp + ((q ^ (lane & 7)) << 4));
          else if (kRow) rowq = raw[q];
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(av[8 * q4 + i]);
          if (kBias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * q);
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * q + 1);
            if (kRes) {   // the only biased kind that takes a row scale
              x[0] = fmaf(b0.x, bscale, x[0]); x[1] = fmaf(b0.y, bscale, x[1]); x[2] = fmaf(b0.z, bscale, x[2]); x[3] = fmaf(b0.w, bscale, x[3]);
              x[4] = fmaf(b1.x, bscale, x[4]); x[5] = fmaf(b1.y, bscale, x[5]); x[6] = fmaf(b1.z, bscale, x[6]); x[7] = fmaf(b1.w, bscale, x[7]);
            } else {
              x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
              x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
          }
          if (KIND == EPI_BIAS_GELU_C2) {   // pre-activation goes out first; keep it for the GELU round below
#pragma unroll
            for (int i = 0; i < 8; ++i) pre[8 * q + i] = x[i];
          }
          if (kAux) {
            float t[8];
            unpack_bf16x8(rowq, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] *= gelu_fast_grad_f(t[i]);
          }
          if (kScale && p.row_scale != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] *= vscale;
          }
          if (kRes) {
            float t[8];
            unpack_bf16x8(rowq, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] += t[i];
          }
          sts_pack8(rowp, q, lane, x);
        }
      }
      staging_store(stage_addr, lane, KIND == EPI_BIAS_GELU_C2 ? tmC2 : tmC, n0, row_in_batch0, batch);
      if (KIND == EPI_BIAS_GELU_C2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) pre[i] = gelu_fast_f(pre[i]) * vscale;   // overlaps the store's smem read
        staging_acquire(lane);
#pragma unroll
        for (int q = 0; q < 8; ++q) sts_pack8(rowp, q, lane, pre + 8 * q);
        staging_store(stage_addr, lane, tmC, n0, row_in_batch0, batch);
      }
    }
    if (!arrived) {   // BN == 64: the odd warp of each lane quarter has no unit
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8u * acc);
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
  }
}

template <int BN, bool B_MN, bool STAT>
__global__ void __launch_bounds__(kTnThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const __grid_constant__ CUtensorMap tmC2, const TnArgs p) {
  using Cfg = TnCfg<BN, STAT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int nstages = STAT ? p.stages : Cfg::kStages;
  const uint32_t sA = base;
  const uint32_t sB = base + nstages * Cfg::kABytes;
  const uint32_t b_region = STAT ? (uint32_t)p.k_blocks * Cfg::kBBytes : (uint32_t)Cfg::kBRegion;
  const uint32_t sStage = sB + b_region;
  const uint32_t kBarOff = nstages * Cfg::kABytes + b_region + Cfg::kStagingBytes;
  const uint32_t bar0 = base + kBarOff;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * Cfg::kStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * Cfg::kStages + 2 + a); };
  const uint32_t bres_bar = bar0 + 8u * (2 * Cfg::kStages + 4);
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + kBarOff + 8 * (2 * Cfg::kStages + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kTnEpiWarps);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // everything above touched only this CTA's shared / tensor memory; global memory from here on

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------------------
    if (lane == 0) {
      TileWalk<STAT> w(p);
      if (STAT && w.valid()) {  // the weight panel of this CTA's n-block, once
        mbar_expect_tx(bres_bar, (uint32_t)p.k_blocks * Cfg::kBBytes);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          if (!B_MN) {
            tma_load_2d(&tmB, bres_bar, sB + kb * Cfg::kBBytes, kb * BK, w.n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(&tmB, bres_bar, sB + kb * Cfg::kBBytes + j * (64 * 128), w.n_blk * BN + j * 64, kb * BK);
          }
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (; w.valid(); w.next()) {
        const int batch = w.mt / p.m_tiles_per_batch;
        const int m0 = (w.mt % p.m_tiles_per_batch) * BM;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::kStageTx);
          const int tap = kb / p.kb_per_tap;
          const int kc = (kb - tap * p.kb_per_tap) * BK;
          tma_load_3d(&tmA, full_bar(stage), sA + stage * Cfg::kABytes, kc, m0 + (tap - p.tap_pad) * p.tap_row_stride, batch);
          if (!STAT) {
            if (!B_MN) {
              tma_load_2d(&tmB, full_bar(stage), sB + stage * Cfg::kBBytes, kb * BK, w.n_blk * BN);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)  // 64(k) x 64(n) boxes, N contiguous
                tma_load_2d(&tmB, full_bar(stage), sB + stage * Cfg::kBBytes + j * (64 * 128), w.n_blk * BN + j * 64, kb * BK);
            }
          }
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TileWalk<STAT> w(p);
      if (STAT && w.valid()) mbar_wait(bres_bar, 0);
      for (; w.valid(); w.next()) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = sA + stage * Cfg::kABytes;
          const uint32_t b_addr = sB + (STAT ? kb : stage) * Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_desc(b_addr + k * 2048, 64 * 128, 1024) : make_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------ epilogue warps ----------------------------------------
    // 8 warps: warp w reads TMEM lane quarter (w & 3); the two warps of a quarter alternate over 128-byte
    // output units (64 bf16 / 32 fp32 columns).  Thread = one output row: tcgen05.ld hands it 32 consecutive
    // columns, residual / aux / pos come in as 16-byte vectors of that row (prefetched to L2 while the MMA
    // runs), and the finished unit is staged in 128B-swizzled shared memory and written with one TMA store
    // (full, coalesced lines; rows beyond the batch are clipped by the tensor map).
    const int ew = warp - 2;
    const int lane_grp = warp & 3;
    const int half = ew >> 2;
    const uint32_t my_stage = sStage + ew * 4096;
    uint8_t* my_stage_ptr = smem + nstages * Cfg::kABytes + b_region + ew * 4096;
    const bool out_bf16 = p.out_dtype == AFB_BF16;
    const int ncl = out_bf16 ? 2 : 1;          // tcgen05.ld chunks per 128-byte staging row
    const int units = BN / (32 * ncl);
    if (p.epi_kind != EPI_GENERIC) {
#define AFB_EPI_FAST(KIND) \
  epi_fast_loop<BN, STAT, KIND>(p, &tmC, &tmC2, tmem_base, tfull_bar(0), tempty_bar(0), my_stage_ptr, my_stage, lane, lane_grp, half)
      switch (p.epi_kind) {
        case EPI_BIAS: AFB_EPI_FAST(EPI_BIAS); break;
        case EPI_BIAS_RES: AFB_EPI_FAST(EPI_BIAS_RES); break;
        case EPI_BIAS_GELU_C2: AFB_EPI_FAST(EPI_BIAS_GELU_C2); break;
        case EPI_GELU_BWD: AFB_EPI_FAST(EPI_GELU_BWD); break;
        default: AFB_EPI_FAST(EPI_PLAIN); break;
      }
#undef AFB_EPI_FAST
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    bool first_tile = true;
    auto prefetch_tile = [&](const TileWalk<STAT>& t) {
      if (p.residual == nullptr && p.aux == nullptr) return;
      const int rl = (t.mt % p.m_tiles_per_batch) * BM + lane_grp * 32 + lane;
      if (rl >= p.rows_per_batch) return;
      const int64_t r = (int64_t)(t.mt / p.m_tiles_per_batch) * p.rows_per_batch + rl;
      for (int u = half; u < units; u += 2) {
        const int n0 = t.n_blk * BN + u * 32 * ncl;
        if (p.residual != nullptr) prefetch_row(p.residual, r * p.ldres + n0, p.res_dtype, 32 * ncl);
        if (p.aux != nullptr) prefetch_row(p.aux, r * p.ldaux + n0, p.aux_dtype, 32 * ncl);
      }
    };
    for (TileWalk<STAT> w(p); p.epi_kind == EPI_GENERIC && w.valid(); w.next()) {
      const int batch = w.mt / p.m_tiles_per_batch;
      const int m0 = (w.mt % p.m_tiles_per_batch) * BM;
      const int row_in_batch0 = m0 + lane_grp * 32;
      const int row_local = row_in_batch0 + lane;
      const bool valid = row_local < p.rows_per_batch;
      const int64_t row = (int64_t)batch * p.rows_per_batch + row_local;
      const int ncol0 = w.n_blk * BN;
      // Warm L2 with the epilogue operands (residual / aux rows) ONE TILE AHEAD: the per-thread row loads below
      // only keep ~16 KB in flight per SM, which caps an HBM-latency stream at ~1.2 TB/s; as L2 hits they run
      // at several TB/s.  Each warp prefetches the 128-byte units it will read itself.
      if (first_tile) {
        prefetch_tile(w);
        first_tile = false;
      }
      {
        TileWalk<STAT> wn = w;
        wn.next();
        if (wn.valid()) prefetch_tile(wn);
      }
      float rscale = 1.f;
      if (p.row_scale != nullptr && valid) rscale = p.row_scale[row / p.row_scale_div];
      const float* pos_row = p.pos != nullptr ? p.pos + (int64_t)(row % p.pos_rows) * p.N : nullptr;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      bool arrived = false;
#pragma unroll 1
      for (int u = half; u < units; u += 2) {
#pragma unroll 1
        for (int pass = 0; pass < (p.has_c2 ? 2 : 1); ++pass) {   // pass 0 of a GELU+preact GEMM stores the pre-activation
#pragma unroll 1
          for (int c = 0; c < ncl; ++c) {
            const int ch = u * ncl + c;
            const int n0 = ncol0 + ch * 32;
            float v[32], t[32];
            const bool last_pass = pass == (p.has_c2 ? 1 : 0);
            const bool has_res = p.residual != nullptr && valid && last_pass;
            const bool has_aux = p.act == AFB_ACT_GELU_BWD && valid;
            // epilogue operands are requested before the TMEM wait so their latency overlaps it
            if (has_res) load32_dyn(p.residual, row * p.ldres + n0, p.res_dtype, t);
            else if (has_aux) load32_dyn(p.aux, row * p.ldaux + n0, p.aux_dtype, t);
            float4 bq[8];
            if (p.bias != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q) bq[q] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + q);
            }
            tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
            if (last_pass && u + 2 >= units && c == ncl - 1) {  // this warp's last TMEM read of the accumulator
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(acc));
              arrived = true;
            }
            if (p.bias != nullptr) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 b4 = bq[q];
                const float bs = p.rs_bias ? rscale : 1.f;
                v[4 * q] = v[4 * q] * p.alpha + b4.x * bs; v[4 * q + 1] = v[4 * q + 1] * p.alpha + b4.y * bs;
                v[4 * q + 2] = v[4 * q + 2] * p.alpha + b4.z * bs; v[4 * q + 3] = v[4 * q + 3] * p.alpha + b4.w * bs;
              }
            } else if (p.alpha != 1.f) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= p.alpha;
            }
            if (pos_row != nullptr && valid) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 e4 = __ldg(reinterpret_cast<const float4*>(pos_row + n0) + q);
                v[4 * q] += e4.x; v[4 * q + 1] += e4.y; v[4 * q + 2] += e4.z; v[4 * q + 3] += e4.w;
              }
            }
            if (last_pass) {
              if (p.act == AFB_ACT_GELU) {
                if (out_bf16) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = gelu_fast_f(v[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = gelu_f(v[i]);
                }
              } else if (p.act == AFB_ACT_GELU_BWD) {
                if (has_aux) {
                  if (out_bf16) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= gelu_fast_grad_f(t[i]);
                  } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= gelu_grad_f(t[i]);
                  }
                }
              } else if (p.act == AFB_ACT_RELU) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (p.row_scale != nullptr && !p.rs_bias) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= rscale;
              }
              if (has_res) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += t[i];
              }
            }
            // stage: row = lane, 128-byte rows, 16-byte chunk index XOR (row & 7)  (SWIZZLE_128B).  The previous
            // TMA store of this warp must have finished READING the buffer; waiting here (not right after
            // issuing it) lets that read overlap the TMEM load and the epilogue math of this unit.
            if (c == 0) {
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
            }
            uint8_t* rowp = my_stage_ptr + lane * 128;
            if (out_bf16) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t wv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
                  wv[j] = *reinterpret_cast<uint32_t*>(&h2);
                }
                const int chunk = (c * 4 + q) ^ (lane & 7);
                *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int chunk = q ^ (lane & 7);
                *reinterpret_cast<float4*>(rowp + chunk * 16) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int col = ncol0 + u * 32 * ncl;
            tma_store_3d((p.has_c2 && pass == 0) ? &tmC2 : &tmC, my_stage, col, row_in_batch0, batch);
            bulk_commit();
          }
        }
      }
      if (!arrived) {  // BN == 64 with bf16 output: the odd warps of each quarter have no unit in this tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// dW kernel: contraction over rows.  A = G^T (MN-major, M-dim = N1 tile), B = X^T (MN-major, N-dim = N2 tile
// of BN2).  Each CTA owns one (n1 tile, n2 tile, row-split), accumulates its row blocks in TMEM and adds its
// partial into dW with fp32 atomics.
//   BN1 = 256 (N1 % 256 == 0): two M=128 accumulators share every X tile -- the kernel is bound by the L2->SM
//   fill (~6300 B/clk chip-wide), and a 256 x 256 tile moves 64 KB per 64-row block instead of 2 x 48 KB.
//   TMEM is then full (2 x 256 columns), so the bias gradient (column sums of G) is accumulated by the
//   otherwise idle epilogue warps straight from the G tiles in shared memory.
//   BN1 = 128: one accumulator; the bias gradient rides on one extra N=16 MMA against an all-ones tile.
// ---------------------------------------------------------------------------------------------
template <int BN1, int BN2, int TAPS> struct DwCfg {
  static constexpr int kABytes = 64 * BN1 * 2;      // 64 rows x BN1 n1
  static constexpr int kTapBytes = 64 * BN2 * 2;    // 64 rows x BN2 n2 of one tap
  static constexpr int kBBytes = TAPS * kTapBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (192 * 1024) / kStageBytes > 6 ? 6 : (192 * 1024) / kStageBytes;
  static constexpr bool kOnesTrick = BN1 == 128;
  // BN1 == 128: accumulator columns [0, BN2) hold dW, [BN2, BN2 + 16) hold G^T * ones (the bias gradient)
  static constexpr int kUsedCols = kOnesTrick ? TAPS * BN2 + 16 : 2 * BN2;
  static constexpr int kTmemCols = kUsedCols > 256 ? 512 : (kUsedCols > 128 ? 256 : 128);
  static constexpr int kScratchBytes = kEpiWarps * 32 * kScratchStride * 4;
  static constexpr int kOnesBytes = 8192;           // one all-ones [64 x 64] bf16 tile (MN-major B operand)
  static constexpr int kSmemBytes = kStages * kStageBytes + kScratchBytes + kOnesBytes + 256 + 1024;
};

struct DwArgs {
  int n1_tiles, n2_tiles, splits;
  int row_blocks_per_batch, total_row_blocks;
  int N1, N2;
  int x_row_shift;
  float* dW;
  long long ld1, ld2;
  float alpha;
  float* dbias;   // optional: dbias[n1] += alpha * sum_m G[m, n1]
  int tap_row_stride;        // TAPS > 1: tap t reads X rows shifted by x_row_shift + t * tap_row_stride ...
  long long tap_dw_stride;   // ... and accumulates into dW + t * tap_dw_stride
  int vec4;       // ld2 == 1 and 16-byte aligned rows: the epilogue uses red.global.add.v4.f32
  const float* dbias_rs;   // optional per-row factor of the bias gradient (smem column-sum path only)
  int rs_div;
  long long rows_per_batch;
  int rev;   // row blocks visited in descending order (ping-pong traversal)
};

template <int BN1, int BN2, int TAPS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_dw_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const DwArgs p) {
  using Cfg = DwCfg<BN1, BN2, TAPS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t sA = base;
  const uint32_t sB = base + Cfg::kStages * Cfg::kABytes;
  float* scratch = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  const uint32_t sOnes = base + Cfg::kStages * Cfg::kStageBytes + Cfg::kScratchBytes;
  constexpr uint32_t kDwBarOff = Cfg::kStages * Cfg::kStageBytes + Cfg::kScratchBytes + Cfg::kOnesBytes;
  const uint32_t bar0 = base + kDwBarOff;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (Cfg::kStages + s); };
  const uint32_t tfull_bar = bar0 + 8u * (2 * Cfg::kStages);
  const uint32_t tmem_slot = bar0 + 8u * (2 * Cfg::kStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + kDwBarOff + 8 * (2 * Cfg::kStages + 1));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  // work decomposition: blockIdx.x = (tile, split)
  const int split = blockIdx.x % p.splits;
  const int tile = blockIdx.x / p.splits;
  const int n2_blk = tile % p.n2_tiles;
  const int n1_blk = tile / p.n2_tiles;
  const int per = (p.total_row_blocks + p.splits - 1) / p.splits;
  const int rb_begin = split * per;
  const int rb_end = min(rb_begin + per, p.total_row_blocks);
  const int n_rb = max(rb_end - rb_begin, 0);
  // the CTAs of the first n2 tile also reduce G over rows (bias gradient)
  const bool do_bias = p.dbias != nullptr && n2_blk == 0;
  const bool smem_bias = do_bias && !Cfg::kOnesTrick;
  if (do_bias && Cfg::kOnesTrick) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kScratchBytes);
    for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x) ones[i] = 0x3f803f80u;  // bf16 1.0 pairs
    fence_async_smem();   // generic-proxy writes -> visible to the tensor-core (async) proxy
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), smem_bias ? 1 + kEpiWarps : 1);   // MMA commit (+ the column-sum warps)
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // global memory (operands, the dW / dbias accumulators) only from here on

  if (n_rb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int rb_i = rb_begin; rb_i < rb_end; ++rb_i) {
          const int rb = p.rev ? p.total_row_blocks - 1 - rb_i : rb_i;
          const int batch = rb / p.row_blocks_per_batch;
          const int r0 = (rb % p.row_blocks_per_batch) * 64;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
#pragma unroll
          for (int j = 0; j < BN1 / 64; ++j)
            tma_load_3d(&tmG, full_bar(stage), sA + stage * Cfg::kABytes + j * 8192, n1_blk * BN1 + j * 64, r0, batch);
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
            for (int j = 0; j < BN2 / 64; ++j)
              tma_load_3d(&tmX, full_bar(stage), sB + stage * Cfg::kBBytes + tap * Cfg::kTapBytes + j * 8192, n2_blk * BN2 + j * 64,
                          r0 + p.x_row_shift + tap * p.tap_row_stride, batch);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(128, BN2, 1, 1);
        constexpr uint32_t idesc_ones = make_idesc(128, 16, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_rb; ++i) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = sA + stage * Cfg::kABytes;
          const uint32_t b_addr = sB + stage * Cfg::kBBytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 16 contraction rows per MMA = two 8-row atoms
#pragma unroll
            for (int h = 0; h < BN1 / 128; ++h) {   // 128 n1 columns = two 64-column boxes per accumulator
              const uint64_t adesc = make_desc(a_addr + h * 16384 + k * 2048, 8192, 1024);
#pragma unroll
              for (int tap = 0; tap < TAPS; ++tap) {   // the taps of a k x 1 conv share the G tile
                const uint64_t bdesc = make_desc(b_addr + tap * Cfg::kTapBytes + k * 2048, 8192, 1024);
                umma_bf16(tmem_base + (uint32_t)((h * TAPS + tap) * BN2), adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
              }
              if (Cfg::kOnesTrick && do_bias)
                umma_bf16(tmem_base + TAPS * BN2, adesc, make_desc(sOnes, 8192, 1024), idesc_ones, (i | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar);
      }
    } else {
      const int ew = warp - 2;
      const int lane_grp = warp & 3;
      if (smem_bias) {
        // column sums of the G tiles while the MMAs run: thread = one bf16 pair (2 of the 256 n1 columns),
        // 64 rows per stage, 128-byte rows with the 16-byte chunk index XOR (row & 7) (SWIZZLE_128B)
        const int box = ew, word = lane;           // 4 warps x 32 lanes x 2 columns = 256
        float s0 = 0.f, s1 = 0.f;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_rb; ++i) {
          float sc0 = 1.f, sc1 = 1.f;   // factors of tile rows `lane` and `lane + 32` (rows past the batch are zero-filled)
          if (p.dbias_rs != nullptr) {
            const int rb = p.rev ? p.total_row_blocks - 1 - (rb_begin + i) : rb_begin + i;
            const long long bat = rb / p.row_blocks_per_batch;
            const long long r0 = (long long)(rb % p.row_blocks_per_batch) * 64 + lane;
            const long long ra = bat * p.rows_per_batch + min(r0, p.rows_per_batch - 1);
            const long long rb2 = bat * p.rows_per_batch + min(r0 + 32, p.rows_per_batch - 1);
            sc0 = p.dbias_rs[ra / p.rs_div];
            sc1 = p.dbias_rs[rb2 / p.rs_div];
          }
          mbar_wait(full_bar(stage), phase);
          const uint8_t* g = smem + stage * Cfg::kABytes + box * 8192 + (word & 3) * 4;
          if (p.dbias_rs != nullptr) {
#pragma unroll 16
            for (int r = 0; r < 64; ++r) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(g + r * 128 + (((word >> 2) ^ (r & 7)) << 4));
              const float f = __shfl_sync(0xffffffffu, r < 32 ? sc0 : sc1, r & 31);
              s0 = fmaf(f, __uint_as_float(v << 16), s0);
              s1 = fmaf(f, __uint_as_float(v & 0xffff0000u), s1);
            }
          } else {
#pragma unroll 16
            for (int r = 0; r < 64; ++r) {
              const uint32_t v = *reinterpret_cast<const uint32_t*>(g + r * 128 + (((word >> 2) ^ (r & 7)) << 4));
              s0 += __uint_as_float(v << 16);
              s1 += __uint_as_float(v & 0xffff0000u);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1u; }
        }
        const int n1 = n1_blk * BN1 + box * 64 + word * 2;
        if (n1 < p.N1) atomicAdd(p.dbias + n1, p.alpha * s0);
        if (n1 + 1 < p.N1) atomicAdd(p.dbias + n1 + 1, p.alpha * s1);
      }
      float* my = scratch + ew * 32 * kScratchStride;
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int acc = 0; acc < (BN1 / 128) * TAPS; ++acc) {
        const int h = acc / TAPS;
        float* dWt = p.dW + (acc % TAPS) * p.tap_dw_stride;
        for (int ch = 0; ch < BN2 / 32; ++ch) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * BN2 + ch * 32), v);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(my + lane * kScratchStride + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
          if (p.vec4) {   // dW rows contiguous in n2: 16-byte vector reductions, 8 lanes per row, 4 rows per instruction
            const int c4 = (lane & 7) * 4;
            const int n2 = n2_blk * BN2 + ch * 32 + c4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              const int n1 = n1_blk * BN1 + h * 128 + lane_grp * 32 + r;
              const float4 v4 = *reinterpret_cast<const float4*>(my + r * kScratchStride + c4);
              if (n1 < p.N1)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dWt + n1 * p.ld1 + n2), "f"(p.alpha * v4.x),
                             "f"(p.alpha * v4.y), "f"(p.alpha * v4.z), "f"(p.alpha * v4.w)
                             : "memory");
            }
          } else {
            const int n2 = n2_blk * BN2 + ch * 32 + lane;
#pragma unroll 4
            for (int r = 0; r < 32; ++r) {
              const int n1 = n1_blk * BN1 + h * 128 + lane_grp * 32 + r;
              if (n1 < p.N1 && n2 < p.N2) atomicAdd(dWt + n1 * p.ld1 + n2 * p.ld2, p.alpha * my[r * kScratchStride + lane]);
            }
          }
          __syncwarp();
        }
      }
      if (Cfg::kOnesTrick && do_bias) {   // column BN2 of the accumulator = sum over this CTA's rows of G[:, n1]
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(TAPS * BN2), v);
        const int n1 = n1_blk * BN1 + lane_grp * 32 + lane;
        if (n1 < p.N1) atomicAdd(p.dbias + n1, p.alpha * v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// tensor [d2][d1][d0] (d0 contiguous), strides in elements; 128B swizzle; OOB reads give zero, OOB writes are clipped.
int make_map(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
             uint32_t box0, uint32_t box1, int rank, int dtype = AFB_BF16, uint32_t box2 = 1, int swizzle_bytes = 128) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled unavailable (driver too old?)");
    return AFB_ERR_DRIVER;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  const uint64_t esz = dtype == AFB_BF16 ? 2 : 4;
  cuuint64_t strides[2] = {stride1 * esz, stride2 * esz};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dtype == AFB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (ptr %p dims %llu,%llu,%llu strides %llu,%llu box %u,%u)", (int)r, ptr,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)stride1,
              (unsigned long long)stride2, box0, box1);
    return AFB_ERR_DRIVER;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool B_MN, bool STAT>
int launch_tn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2, const TnArgs& a,
              cudaStream_t st) {
  using Cfg = TnCfg<BN, STAT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN, B_MN, STAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm_tn: cudaFuncSetAttribute(%d B) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  int grid;
  if (STAT) {
    grid = (num_sms() / a.n_tiles) * a.n_tiles;
    const int max_useful = a.total_m_tiles * a.n_tiles;
    if (grid > max_useful) grid = max_useful;
  } else {
    const int tiles = a.total_m_tiles * a.n_tiles;
    grid = tiles < num_sms() ? tiles : num_sms();
  }
  const bool small = pdl_force_all() || a.total_m_tiles * a.n_tiles <= 4 * num_sms();
  cudaError_t le = launch_pdl(small, gemm_tn_kernel<BN, B_MN, STAT>, dim3(grid), dim3(kTnThreads), (size_t)Cfg::kSmemBytes, st, tmA, tmB, tmC, tmC2, a);
  if (le != cudaSuccess) {
    set_error("gemm_tn: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("gemm_tn");
}

template <int BN, bool B_MN>
int launch_tn_stat(bool stat, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmC2,
                   const TnArgs& a, cudaStream_t st) {
  return stat ? launch_tn<BN, B_MN, true>(tmA, tmB, tmC, tmC2, a, st) : launch_tn<BN, B_MN, false>(tmA, tmB, tmC, tmC2, a, st);
}

template <int BN1, int BN2, int TAPS = 1>
int launch_dw(const CUtensorMap& tmG, const CUtensorMap& tmX, const DwArgs& a, cudaStream_t st) {
  using Cfg = DwCfg<BN1, BN2, TAPS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_dw_kernel<BN1, BN2, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("gemm_dw: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured = true;
  }
  const int grid = a.n1_tiles * a.n2_tiles * a.splits;
  const bool small = pdl_force_all() || a.total_row_blocks <= 16 * num_sms();
  cudaError_t le = launch_pdl(small, gemm_dw_kernel<BN1, BN2, TAPS>, dim3(grid), dim3(kThreads), (size_t)Cfg::kSmemBytes, st, tmG, tmX, a);
  if (le != cudaSuccess) {
    set_error("gemm_dw: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("gemm_dw");
}

}  // namespace

// bf16 3-D tensor map (cols, rows, batch) with 128B swizzle for kernels in other translation units (gcn0 fused store)
int make_tensor_map_bf16(void* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                         uint32_t box0, uint32_t box1) {
  return make_map(reinterpret_cast<CUtensorMap*>(map), ptr, d0, d1, d2, stride1, stride2, box0, box1, 3);
}
// same with a box that spans several entries of the outermost dimension (packed short sequences: attention_tc.cu)
int make_tensor_map_bf16_box3(void* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                              uint32_t box0, uint32_t box1, uint32_t box2, int swizzle_bytes) {
  return make_map(reinterpret_cast<CUtensorMap*>(map), ptr, d0, d1, d2, stride1, stride2, box0, box1, 3, AFB_BF16, box2, swizzle_bytes);
}
}  // namespace afb

using namespace afb;

extern "C" int afb_gemm_tn(const afb_gemm_tn_t* p, afb_stream s) {
  AFB_REQUIRE(p && p->A && p->B && p->C, "gemm_tn: null operand");
  AFB_REQUIRE(p->taps >= 1 && p->k_per_tap > 0, "gemm_tn: bad taps/k");
  AFB_REQUIRE(p->N % 64 == 0, "gemm_tn: N=%d must be a multiple of 64", p->N);
  AFB_REQUIRE(p->lda % 8 == 0 && p->ldb % 8 == 0 && p->ldc % 8 == 0, "gemm_tn: leading dims must be 16-byte aligned");
  AFB_REQUIRE((p->residual == nullptr || p->ldres % 8 == 0) && (p->aux == nullptr || p->ldaux % 8 == 0),
              "gemm_tn: residual / aux leading dims must be multiples of 8");
  AFB_REQUIRE(((uintptr_t)p->C & 15) == 0 && ((uintptr_t)p->C2 & 15) == 0 && ((uintptr_t)p->residual & 15) == 0 &&
                  ((uintptr_t)p->aux & 15) == 0 && ((uintptr_t)p->bias & 15) == 0 && ((uintptr_t)p->pos & 15) == 0,
              "gemm_tn: epilogue operands must be 16-byte aligned");
  AFB_REQUIRE(p->taps == 1 || p->k_per_tap % BK == 0, "gemm_tn: k_per_tap %% 64 != 0 with taps > 1");
  AFB_REQUIRE(p->k_per_tap % 8 == 0, "gemm_tn: k_per_tap %% 8 != 0");
  AFB_REQUIRE(((uintptr_t)p->A & 15) == 0 && ((uintptr_t)p->B & 15) == 0, "gemm_tn: operands must be 16-byte aligned");
  AFB_REQUIRE(p->rows_per_batch > 0 && p->batches > 0, "gemm_tn: empty problem");

  const int K = p->taps * p->k_per_tap;
  const int kb_per_tap = ceil_div(p->k_per_tap, BK);
  const int k_blocks = kb_per_tap * p->taps;
  // Tile width and mode.  B-stationary when the whole [BN x K] weight panel fits 128 KB of shared memory
  // (K <= 256 at BN 256, <= 512 at BN 128, <= 1024 at BN 64) and there are enough CTAs per n-block.
  int BN = (p->N % 256 == 0) ? 256 : (p->N % 128 == 0 ? 128 : 64);
  bool stat = false;
  static const bool no_stat = getenv("AFB_GEMM_NO_STAT") != nullptr;
  if (!no_stat && p->taps == 1) {
    // L2->SM fill per 128-row m-tile: streaming re-fetches A and the B panel for each of its N/BN tiles,
    // stationary(bn) re-fetches only A, N/bn times.  Pick the variant that moves fewer bytes.
    const long a_tile = 128L * K * 2;
    long best = (long)(p->N / BN) * (a_tile + (long)BN * K * 2);
    static const int max_stat_bn = getenv("AFB_GEMM_STAT_BN") ? atoi(getenv("AFB_GEMM_STAT_BN")) : 256;
    for (int bn = max_stat_bn; bn >= 64; bn >>= 1) {
      if (p->N % bn != 0 || (long)k_blocks * bn * 128 > 128 * 1024 || p->N / bn > num_sms()) continue;
      // the panel load is only amortised over enough m-tiles per CTA (measured: 64 m-tiles of the D=512 stage
      // run 10-25% faster streaming)
      const long total_m_tiles = (long)ceil_div(p->rows_per_batch, BM) * p->batches;
      if (total_m_tiles < 8L * (num_sms() / (p->N / bn))) continue;
      const long fill = (long)(p->N / bn) * a_tile;
      if (fill < best) {
        best = fill;
        BN = bn;
        stat = true;
      }
    }
  }
  TnArgs a;
  a.m_tiles_per_batch = ceil_div(p->rows_per_batch, BM);
  a.n_tiles = p->N / BN;
  a.total_m_tiles = a.m_tiles_per_batch * p->batches;
  a.kb_per_tap = kb_per_tap;
  a.k_blocks = k_blocks;
  a.stages = 0;
  a.rev = next_stream_dir();
  if (stat) {
    a.stages = BN == 256 ? TnCfg<256, true>::stat_stages(k_blocks) : (BN == 128 ? TnCfg<128, true>::stat_stages(k_blocks) : TnCfg<64, true>::stat_stages(k_blocks));
    AFB_REQUIRE(a.stages >= 2, "gemm_tn: internal: stationary panel leaves no room for the A ring");
  }
  a.rows_per_batch = (int)p->rows_per_batch;
  a.batches = p->batches;
  a.N = p->N;
  a.tap_row_stride = p->tap_row_stride;
  a.tap_pad = p->tap_pad;
  a.out_dtype = p->out_dtype; a.act = p->act; a.alpha = p->alpha;
  a.has_c2 = (p->C2 != nullptr && p->act == AFB_ACT_GELU) ? 1 : 0;
  a.bias = p->bias; a.pos = p->pos; a.pos_rows = p->pos_rows > 0 ? p->pos_rows : 1;
  a.aux = p->aux; a.aux_dtype = p->aux_dtype; a.ldaux = p->ldaux;
  a.residual = p->residual; a.res_dtype = p->res_dtype; a.ldres = p->ldres;
  a.row_scale = p->row_scale; a.row_scale_div = p->row_scale_div > 0 ? p->row_scale_div : 1;
  a.rs_bias = (p->row_scale != nullptr && p->row_scale_mode == AFB_RS_BIAS) ? 1 : 0;
  AFB_REQUIRE(!a.rs_bias || p->act == AFB_ACT_NONE, "gemm_tn: AFB_RS_BIAS cannot be combined with an activation");
  a.epi_kind = EPI_GENERIC;
  static const bool no_fast_epi = getenv("AFB_GEMM_GENERIC_EPI") != nullptr;
  if (!no_fast_epi && p->out_dtype == AFB_BF16 && p->alpha == 1.f && p->pos == nullptr) {
    const bool res = p->residual != nullptr, bias = p->bias != nullptr, rs = p->row_scale != nullptr;
    if (p->act == AFB_ACT_NONE && bias && !res && !rs) a.epi_kind = EPI_BIAS;
    else if (p->act == AFB_ACT_NONE && bias && res && p->res_dtype == AFB_BF16) a.epi_kind = EPI_BIAS_RES;
    else if (p->act == AFB_ACT_GELU && bias && a.has_c2 && !res) a.epi_kind = EPI_BIAS_GELU_C2;
    else if (p->act == AFB_ACT_GELU_BWD && !bias && !res && p->aux != nullptr && p->aux_dtype == AFB_BF16) a.epi_kind = EPI_GELU_BWD;
    else if (p->act == AFB_ACT_NONE && !bias && !res && !a.rs_bias) a.epi_kind = EPI_PLAIN;
  }
  AFB_REQUIRE(p->act != AFB_ACT_GELU_BWD || p->aux != nullptr, "gemm_tn: GELU_BWD needs aux");
  AFB_REQUIRE(p->act != AFB_ACT_GELU_BWD || p->residual == nullptr, "gemm_tn: GELU_BWD cannot be combined with a residual");

  CUtensorMap tmA, tmB, tmC, tmC2;
  int rc = make_map(&tmA, p->A, (uint64_t)p->k_per_tap, (uint64_t)p->rows_per_batch, (uint64_t)p->batches, (uint64_t)p->lda,
                    (uint64_t)p->rows_per_batch * p->lda, BK, BM, 3);
  if (rc) return rc;
  if (!p->b_mn_major)
    rc = make_map(&tmB, p->B, (uint64_t)K, (uint64_t)p->N, 1, (uint64_t)p->ldb, 0, BK, (uint32_t)BN, 2);
  else
    rc = make_map(&tmB, p->B, (uint64_t)p->N, (uint64_t)K, 1, (uint64_t)p->ldb, 0, 64, 64, 2);
  if (rc) return rc;
  const uint32_t out_box = p->out_dtype == AFB_BF16 ? 64 : 32;  // 128 bytes of columns x 32 rows per TMA store
  rc = make_map(&tmC, p->C, (uint64_t)p->N, (uint64_t)p->rows_per_batch, (uint64_t)p->batches, (uint64_t)p->ldc,
                (uint64_t)p->rows_per_batch * p->ldc, out_box, 32, 3, p->out_dtype);
  if (rc) return rc;
  rc = make_map(&tmC2, a.has_c2 ? p->C2 : p->C, (uint64_t)p->N, (uint64_t)p->rows_per_batch, (uint64_t)p->batches, (uint64_t)p->ldc,
                (uint64_t)p->rows_per_batch * p->ldc, out_box, 32, 3, p->out_dtype);
  if (rc) return rc;
  cudaStream_t st = as_stream(s);
  if (!p->b_mn_major) {
    if (BN == 256) return launch_tn_stat<256, false>(stat, tmA, tmB, tmC, tmC2, a, st);
    if (BN == 128) return launch_tn_stat<128, false>(stat, tmA, tmB, tmC, tmC2, a, st);
    return launch_tn_stat<64, false>(stat, tmA, tmB, tmC, tmC2, a, st);
  }
  if (BN == 256) return launch_tn_stat<256, true>(stat, tmA, tmB, tmC, tmC2, a, st);
  if (BN == 128) return launch_tn_stat<128, true>(stat, tmA, tmB, tmC, tmC2, a, st);
  return launch_tn_stat<64, true>(stat, tmA, tmB, tmC, tmC2, a, st);
}

extern "C" int afb_gemm_dw(const afb_gemm_dw_t* p, afb_stream s) {
  AFB_REQUIRE(p && p->G && p->X && p->dW, "gemm_dw: null operand");
  AFB_REQUIRE(p->N2 % 64 == 0, "gemm_dw: N2=%d must be a multiple of 64", p->N2);
  AFB_REQUIRE(p->N1 % 8 == 0, "gemm_dw: N1=%d must be a multiple of 8", p->N1);
  AFB_REQUIRE(p->ldg % 8 == 0 && p->ldx % 8 == 0, "gemm_dw: leading dims must be 16-byte aligned");
  AFB_REQUIRE(((uintptr_t)p->G & 15) == 0 && ((uintptr_t)p->X & 15) == 0, "gemm_dw: operands must be 16-byte aligned");
  const int BN2 = (p->N2 % 256 == 0) ? 256 : (p->N2 % 128 == 0 ? 128 : 64);
  static const bool no_wide = getenv("AFB_DW_BN1_128") != nullptr;
  const int taps = p->taps > 1 ? p->taps : 1;
  AFB_REQUIRE(taps <= 3 && (taps == 1 || (p->N1 <= 128 && p->N2 <= 128)), "gemm_dw: taps=%d needs taps <= 3, N1 <= 128, N2 <= 128", taps);
  const int BN1 = (!no_wide && taps == 1 && p->N1 % 256 == 0) ? 256 : 128;
  DwArgs a;
  a.n1_tiles = ceil_div(p->N1, BN1);
  a.n2_tiles = p->N2 / BN2;
  a.row_blocks_per_batch = ceil_div(p->rows_per_batch, 64);
  a.total_row_blocks = a.row_blocks_per_batch * p->batches;
  const int tiles = a.n1_tiles * a.n2_tiles;
  // one CTA per SM (the smem ring takes ~200 KB): 256-wide tiles run as ONE wave, each CTA paying the pipeline
  // fill and the atomic epilogue once; the narrower tiles keep two waves (shorter tails for the conv shapes)
  int splits = (((BN1 == 256 || taps > 1) ? 1 : 2) * num_sms()) / tiles;
  if (splits < 1) splits = 1;
  if (splits > a.total_row_blocks) splits = a.total_row_blocks;
  // keep at least 8 row blocks per split so the atomic epilogue is amortised
  const int max_splits = a.total_row_blocks / 8 > 0 ? a.total_row_blocks / 8 : 1;
  if (splits > max_splits) splits = max_splits;
  a.splits = splits;
  a.N1 = p->N1; a.N2 = p->N2; a.x_row_shift = p->x_row_shift;
  a.dW = p->dW; a.ld1 = p->ld1; a.ld2 = p->ld2; a.alpha = p->alpha; a.dbias = p->dbias;
  a.tap_row_stride = p->tap_row_stride; a.tap_dw_stride = p->tap_dw_stride;
  a.vec4 = (p->ld2 == 1 && p->ld1 % 4 == 0 && ((uintptr_t)p->dW & 15) == 0 && (taps == 1 || p->tap_dw_stride % 4 == 0)) ? 1 : 0;
  a.dbias_rs = p->dbias != nullptr ? p->dbias_row_scale : nullptr;
  a.rs_div = p->row_scale_div > 0 ? p->row_scale_div : 1;
  a.rows_per_batch = p->rows_per_batch;
  a.rev = next_stream_dir();
  AFB_REQUIRE(a.dbias_rs == nullptr || BN1 == 256, "gemm_dw: dbias_row_scale needs N1 %% 256 == 0 (N1=%d)", p->N1);
  CUtensorMap tmG, tmX;
  int rc = make_map(&tmG, p->G, (uint64_t)p->N1, (uint64_t)p->rows_per_batch, (uint64_t)p->batches, (uint64_t)p->ldg,
                    (uint64_t)p->rows_per_batch * p->ldg, 64, 64, 3);
  if (rc) return rc;
  rc = make_map(&tmX, p->X, (uint64_t)p->N2, (uint64_t)p->rows_per_batch, (uint64_t)p->batches, (uint64_t)p->ldx,
                (uint64_t)p->rows_per_batch * p->ldx, 64, 64, 3);
  if (rc) return rc;
  cudaStream_t st = as_stream(s);
  if (taps > 1) {   // N2 <= 128: BN2 is 128 or 64
    if (BN2 == 128) return taps == 3 ? launch_dw<128, 128, 3>(tmG, tmX, a, st) : launch_dw<128, 128, 2>(tmG, tmX, a, st);
    return taps == 3 ? launch_dw<128, 64, 3>(tmG, tmX, a, st) : launch_dw<128, 64, 2>(tmG, tmX, a, st);
  }
  if (BN1 == 256) {
    if (BN2 == 256) return launch_dw<256, 256>(tmG, tmX, a, st);
    if (BN2 == 128) return launch_dw<256, 128>(tmG, tmX, a, st);
    return launch_dw<256, 64>(tmG, tmX, a, st);
  }
  if (BN2 == 256) return launch_dw<128, 256>(tmG, tmX, a, st);
  if (BN2 == 128) return launch_dw<128, 128>(tmG, tmX, a, st);
  return launch_dw<128, 64>(tmG, tmX, a, st);
}
