// Short-sequence multi-head attention, forward (below) and backward (second half of the file), on the Blackwell paths: TMA
// loads, tcgen05.mma with the scores and the outputs in TMEM, softmax out of tcgen05.ld, TMA stores.  (reference: model/AltFormer/model_ST.py:49-67, q @ k^T -> softmax
// -> @ v per (sequence, head); L = 22 joints / 32 frames on SHREC, 46 / 64 on LMDHG.)
//
// One problem is far smaller than an MMA (22 x 22 scores per head), so G = 128 / LP sequences are PACKED into one
// 128-row tile, each padded to LP = 32 or 64 rows:
//   * the TMA box is (64 feature columns = one 128-byte swizzle span = two heads at dh 32, LP rows, G sequences) of the
//     3-D view (3D features, L tokens, B sequences) of qkv: rows l >= L of every sequence are out of bounds in the token
//     dimension and arrive as ZEROS, so the padded tile needs no masking of q / k / v and the token rows of sequence g sit
//     at tile rows g*LP .. g*LP + L - 1;
//   * S = Q K^T is ONE tcgen05.mma chain per head (M = 128, N = 128, K = dh): only the G diagonal LP x LP blocks mean
//     anything (the tensor pipe is otherwise idle in this HBM-bound kernel, the G-fold redundant FLOPs are free);
//   * tile row r = TMEM lane r, and the keys of ITS sequence are the LP consecutive TMEM columns (r / LP) * LP ..: each
//     softmax warp owns one lane quarter, so its 32 rows belong to one sequence and ONE warp-uniform tcgen05.ld hands
//     every thread exactly its own row of scores -- the softmax is thread-local (no shuffles, no shared memory);
//   * P goes back to TENSOR memory (tcgen05.st, bf16 pairs; PTMEM = true, the default) as the block-diagonal A operand
//     [128 x 128] whose off-diagonal blocks are zeroed once per CTA and never written -- or, PTMEM = false, to shared memory
//     (K-major, 128B swizzle); O = P V is a second MMA chain (N = dh, K = 128, A from TMEM, V read in place from its TMA
//     box as an MN-major B operand whose 32-column slice starts mid-swizzle-span at dh 32), the finished 128 x dh tile is
//     read back with tcgen05.ld, scaled by 1 / rowsum (and the optional DropPath factor), staged and written with a TMA
//     store whose box clips the padded rows;
//   * the sequence length is a template constant (LC) for the dataset lengths 22 / 32 / 46 / 64: masked key columns then
//     cost nothing (the compare + select per column was a sixth of the instructions); other lengths run the LC = 0 variant.
// Warp roles: 0 = TMA producer, 1 = S issuer (+ TMEM owner), 10 = P V issuer, 2..5 and 6..9 = two softmax / epilogue groups (thread = tile
// row; group b owns the items i = b mod 2 and the S / P / O buffers b), so two heads are in the softmax at any time while
// the tensor pipe works on the products either side of them.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace afb {

int make_tensor_map_bf16_box3(void* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1, uint64_t stride2,
                              uint32_t box0, uint32_t box1, uint32_t box2, int swizzle_bytes);

namespace attn_tc {

constexpr int kThreads = 64 + 256 + 32;        // producer, S issuer, two softmax groups of four warps, P V issuer
constexpr int kBoxBytes = 128 * 128;          // one (64 columns x 128 rows) bf16 box
constexpr int kStageBytes = 3 * kBoxBytes;    // q | k | v boxes of one 64-column box (two heads at dh 32, one at dh 64)
constexpr int kPBytes = 128 * 128 * 2;        // block-diagonal probabilities in shared memory (PTMEM = false): two 64-key k-blocks
constexpr unsigned long long kWaitTimeoutNs = 4000000000ull;

// ---- PTX wrappers (same conventions as gemm_tcgen05.cu) ---------------------------------------------------------
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
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded: a protocol bug traps (a launch error) instead of hanging the GPU
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
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane = tile row); no wait
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
// 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (lane = row, two bf16 per 32-bit column) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// shared-memory matrix descriptor, 128B swizzle (see gemm_tcgen05.cu:make_desc)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
  long long B;
  int L, heads, D;
  int tiles;              // ceil(B / G)
  float scale_log2;       // softmax scale * log2(e)
  float scale;            // softmax scale (backward: folded into dS)
  const float* out_scale; // optional [B]
  int rev;                // tiles visited in descending order (ping-pong traversal, api.cu:next_stream_dir)
};

// PTMEM: the probabilities go back to TENSOR memory (tcgen05.st) and the P V product reads its A operand from there;
// otherwise they are staged in shared memory as a K-major 128B-swizzled operand.
template <int LP, int DH, bool PTMEM>
struct Cfg {
  static constexpr int G = 128 / LP;            // sequences per tile
  static constexpr int HB = 64 / DH;            // heads per 128-byte box
  static constexpr int KS_S = DH / 16;          // k-steps of the scores MMA
  static constexpr int kOutBytes = 128 * DH * 2;   // one head's output tile, staged per softmax group
  static constexpr int kOutBufs = (DH == 32 && PTMEM) ? 2 : 1;   // staging tiles per group (two: the previous store may still be reading)
  static constexpr int kPSmem = PTMEM ? 0 : 2 * kPBytes;
  static constexpr int kStages = PTMEM ? 4 : (DH == 32 ? 3 : 2);
  // tensor memory: S[2] at 0 / 128 (fp32 scores, 128 columns), O at 256 + 64 group (+ 32 sub-buffer at dh 32), P[2] at 384 / 448
  // (128 keys as bf16 pairs = 64 columns, off-diagonal blocks zeroed once)
  static constexpr int kTmemCols = 512;
  static constexpr int kSmem = 1024 /*align*/ + kStages * kStageBytes + kPSmem + 2 * kOutBufs * kOutBytes + 256 /*barriers*/;
};

// LC: the sequence length as a compile-time constant (0 = read it from the arguments): the masked key columns then cost
// nothing instead of a compare + select each, which was a sixth of the kernel's instructions.
template <int LP, int DH, bool PTMEM, int LC>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmO, const TcArgs a) {
  using C = Cfg<LP, DH, PTMEM>;
  constexpr int G = C::G, HB = C::HB, kStages = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStage = base;
  const uint32_t sP = base + kStages * kStageBytes;
  const uint32_t sOut = sP + C::kPSmem;
  const uint32_t sBar = sOut + 2 * C::kOutBufs * C::kOutBytes;
  uint8_t* pP = smem + kStages * kStageBytes;
  uint8_t* pOut = pP + C::kPSmem;
  // barriers (8 bytes each)
  auto full_bar = [&](int s) { return sBar + 8 * s; };
  auto empty_bar = [&](int s) { return sBar + 8 * (kStages + s); };
  auto sfull_bar = [&](int b) { return sBar + 8 * (2 * kStages + b); };
  auto sfree_bar = [&](int b) { return sBar + 8 * (2 * kStages + 2 + b); };
  auto pfull_bar = [&](int b) { return sBar + 8 * (2 * kStages + 4 + b); };
  auto pfree_bar = [&](int b) { return sBar + 8 * (2 * kStages + 6 + b); };
  // O is double-buffered PER GROUP at dh 32 (four 32-column buffers: the P V product of item i must not wait until the group
  // has read item i - 2 back, which it only does after publishing P(i) -- 19 % of all stall samples sat in that chain)
  constexpr int NSUB = DH == 32 ? 2 : 1;
  auto ofull_bar = [&](int b, int sub) { return sBar + 8 * (2 * kStages + 8 + 2 * b + sub); };
  auto ofree_bar = [&](int b, int sub) { return sBar + 8 * (2 * kStages + 12 + 2 * b + sub); };
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  const int nbox = a.heads / HB;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < a.tiles; t += gridDim.x) ++my_tiles;
  const int total = my_tiles * a.heads;   // (tile, head) items of this CTA, in order

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sfull_bar(b), 1); mbar_init(sfree_bar(b), 4);
      mbar_init(pfull_bar(b), 4); mbar_init(pfree_bar(b), 1);
      for (int sub = 0; sub < 2; ++sub) { mbar_init(ofull_bar(b, sub), 1); mbar_init(ofree_bar(b, sub), 4); }
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), C::kTmemCols);
  if (!PTMEM) {   // the off-diagonal blocks of both P buffers stay zero for the whole kernel
    uint4* z = reinterpret_cast<uint4*>(pP);
    for (int i = tid; i < C::kPSmem >> 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  pdl_wait();   // programmatic dependent launch (common.cuh): global memory only from here on

  if (warp == 0) {
    // ------------------------------------------ TMA producer ---------------------------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = blockIdx.x; ti < a.tiles; ti += gridDim.x) {
        const int t = a.rev ? a.tiles - 1 - ti : ti;
        for (int bx = 0; bx < nbox; ++bx) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), kStageBytes);
          const uint32_t dst = sStage + stage * kStageBytes;
#pragma unroll
          for (int s = 0; s < 3; ++s) tma_load_3d(&tmQ, full_bar(stage), dst + s * kBoxBytes, s * a.D + bx * 64, 0, t * G);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ------------------------------------------ MMA issuers ----------------------------------------------------
    // Two issuing threads so that neither product waits behind the other's barrier: warp 1 issues the scores of an item
    // as soon as its S buffer has been read out (two items ahead of the P V product at most), warp 10 issues P V when
    // the probabilities are in place.  They write disjoint TMEM regions; each commit covers its own thread's MMAs.
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc(128, DH, 0, 1);
      auto stage_of = [&](int item, uint32_t& ph) {
        const int box = item / HB;
        ph = (uint32_t)((box / kStages) & 1);
        return box % kStages;
      };
      if (warp == 1) {
        for (int i = 0; i < total; ++i) {   // S(i) = Q K^T of item i
          uint32_t ph;
          const int stage = stage_of(i, ph);
          const int slot = i % HB, b = i & 1;
          if (slot == 0) mbar_wait(full_bar(stage), ph);
          mbar_wait(sfree_bar(b), (uint32_t)(((i >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t q_addr = sStage + stage * kStageBytes + slot * (DH * 2);
          const uint32_t k_addr = q_addr + kBoxBytes;
#pragma unroll
          for (int k = 0; k < C::KS_S; ++k)
            umma_bf16(tmem + (uint32_t)(b * 128), make_desc(q_addr + k * 32, 16, 1024), make_desc(k_addr + k * 32, 16, 1024), idesc_s,
                      k != 0 ? 1u : 0u);
          umma_commit(sfull_bar(b));
        }
      } else {
        for (int j = 0; j < total; ++j) {   // O(j) = P(j) V
          uint32_t ph;
          const int stage = stage_of(j, ph);
          const int slot = j % HB, b = j & 1;
          mbar_wait(pfull_bar(b), (uint32_t)((j >> 1) & 1));
          const int sub = (j >> 1) % NSUB, ou = (j >> 1) / NSUB;     // O buffer of the group and how often it has been used
          mbar_wait(ofree_bar(b, sub), (uint32_t)((ou & 1) ^ 1));
          tc_fence_after();
          const uint32_t v_addr = sStage + stage * kStageBytes + 2 * kBoxBytes + slot * (DH * 2);
          const uint32_t d_addr = tmem + 256u + (uint32_t)(b * 64 + sub * 32);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bdesc = make_desc(v_addr + ks * 2048, 8192, 1024);
            if (PTMEM)
              umma_bf16_ts(d_addr, tmem + 384u + (uint32_t)(b * 64 + ks * 8), bdesc, idesc_o, ks != 0 ? 1u : 0u);
            else
              umma_bf16(d_addr, make_desc(sP + b * kPBytes + (ks >> 2) * kBoxBytes + (ks & 3) * 32, 16, 1024), bdesc, idesc_o, ks != 0 ? 1u : 0u);
          }
          umma_commit(ofull_bar(b, sub));
          umma_commit(pfree_bar(b));
          // the q | k | v boxes of this head group are consumed: every S product of the box finished before its softmax,
          // which finished before the P V products this commit covers
          if (slot == HB - 1) umma_commit(empty_bar(stage));
        }
      }
    }
  } else {
    // ------------------------------------------ softmax / epilogue (thread = tile row) -------------------------
    const int grp = (warp - 2) >> 2;            // softmax group = buffer index of its items
    const int quarter = warp & 3;               // TMEM lane quarter this warp may read
    const int r = quarter * 32 + lane;          // tile row
    const int g = r / LP;                       // sequence inside the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int L = LC != 0 ? LC : a.L;
    const int key0 = g * LP;                    // first key (= S column) of this row's sequence
    // P destination of this row.  shared memory: keys key0 .. key0 + LP - 1 of k-block key0 / 64;  tensor memory:
    // columns key0 / 2 .. of the group's P region
    uint8_t* prow = pP + grp * kPBytes + (key0 >> 6) * kBoxBytes + r * 128;
    const int pchunk0 = (key0 & 63) >> 3;
    const uint32_t p_taddr = lane_addr + 384u + (uint32_t)(grp * 64 + (key0 >> 1));
    uint8_t* orow = pOut + grp * C::kOutBufs * C::kOutBytes + r * (DH * 2);
    const int oswz = DH == 64 ? (r & 7) : ((r >> 1) & 3);     // 128B / 64B swizzle of the staging rows
    const bool elected = ((warp - 2) & 3) == 0 && lane == 0;  // one store issuer per group
    auto grp_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
    const uint32_t sfull = sfull_bar(grp), sfree = sfree_bar(grp), pfull = pfull_bar(grp), pfree = pfree_bar(grp);

    if (PTMEM) {   // zero the group's whole P region once: only the diagonal block of each row is ever rewritten
      uint32_t z[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(lane_addr + 384u + (uint32_t)(grp * 64 + c * 16), z);
      tmem_wait_st();
    }
    // O phase of item j: read the finished 128 x dh tile back, scale by inv (1 / rowsum of that item), stage, store
    // (tile, head) of the group's next O phase, advanced incrementally (items grp, grp + 2, ... in order)
    int o_tile = blockIdx.x, o_head = grp, o_cnt = 0;
    auto o_phase = [&](int j, float inv) {
      const int tile = a.rev ? a.tiles - 1 - o_tile : o_tile, h = o_head, ob = C::kOutBufs == 2 ? (o_cnt & 1) : 0;
      ++o_cnt;
      o_head += 2;
      if (o_head >= a.heads) { o_head -= a.heads; o_tile += gridDim.x; }
      const int sub = (j >> 1) % NSUB, ou = (j >> 1) / NSUB;
      mbar_wait(ofull_bar(grp, sub), (uint32_t)(ou & 1));
      tc_fence_after();
      uint32_t orr[DH];
#pragma unroll
      for (int c = 0; c < DH / 32; ++c) tmem_ld32_issue(lane_addr + 256u + (uint32_t)(grp * 64 + sub * 32 + c * 32), orr + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofree_bar(grp, sub));
      float sc = inv;
      if (a.out_scale != nullptr) {   // DropPath keep factor of the row's sequence
        const long long bq = (long long)tile * G + g;
        sc *= bq < a.B ? a.out_scale[bq] : 0.f;
      }
      // the store that last used this staging tile must have finished reading it
      if (elected) {
        if (C::kOutBufs == 2) bulk_wait_read1(); else bulk_wait_read0();
      }
      grp_bar();
      uint8_t* dst = orow + ob * C::kOutBytes;
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q] = pack_bf16x2(__uint_as_float(orr[8 * c + 2 * q]) * sc, __uint_as_float(orr[8 * c + 2 * q + 1]) * sc);
        *reinterpret_cast<uint4*>(dst + ((c ^ oswz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_async_smem();
      grp_bar();
      if (elected) {
        tma_store_3d(&tmO, sOut + (grp * C::kOutBufs + ob) * C::kOutBytes, h * DH, 0, tile * G);
        bulk_commit();
      }
    };
    float inv_prev = 0.f;
    for (int i = grp; i < total; i += 2) {
      mbar_wait(sfull, (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      uint32_t sr[LP];
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) tmem_ld32_issue(lane_addr + (uint32_t)(grp * 128 + key0 + c * 32), sr + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sfree);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < LP; ++j)
        if (j < L) mx = fmaxf(mx, __uint_as_float(sr[j]));
      const float off = mx * a.scale_log2;
      float sum = 0.f;
      uint32_t pk[LP / 2];
#pragma unroll
      for (int j = 0; j < LP; j += 2) {
        const float p0 = j < L ? ex2(fmaf(__uint_as_float(sr[j]), a.scale_log2, -off)) : 0.f;
        const float p1 = j + 1 < L ? ex2(fmaf(__uint_as_float(sr[j + 1]), a.scale_log2, -off)) : 0.f;
        sum += p0 + p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
      }
      mbar_wait(pfree, (uint32_t)(((i >> 1) & 1) ^ 1));
      if (PTMEM) {
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < LP / 32; ++c) tmem_st16(p_taddr + (uint32_t)(c * 16), pk + c * 16);
        tmem_wait_st();
        tc_fence_before();
      } else {
#pragma unroll
        for (int c = 0; c < LP / 8; ++c)
          *reinterpret_cast<uint4*>(prow + (((pchunk0 + c) ^ (r & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull);
      // the P V product of this item now runs on the tensor pipe: read back the group's PREVIOUS item meanwhile
      if (i >= 2) o_phase(i - 2, inv_prev);
      inv_prev = 1.0f / sum;
    }
    if (total > grp) o_phase(((total - 1 - grp) & ~1) + grp, inv_prev);
    if (elected) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

template <int LP, int DH, bool PTMEM, int LC>
int launch(const void* qkv, void* o, int64_t B, int L, int heads, float scale, const float* out_scale, cudaStream_t st) {
  using C = Cfg<LP, DH, PTMEM>;
  auto kern = attn_fwd_tc_kernel<LP, DH, PTMEM, LC>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) {
      set_error("attention_tc: cudaFuncSetAttribute(%d) failed: %s", C::kSmem, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return (int)e;
    }
    configured = true;
  }
  const int D = heads * DH;
  CUtensorMap tmQ, tmO;
  int rc;
  if ((rc = make_tensor_map_bf16_box3(&tmQ, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D, (uint64_t)L * 3 * D, 64, LP, C::G, 128)))
    return rc;
  // output boxes are one head wide: 64-byte rows (64B swizzle) at dh 32, 128-byte rows at dh 64
  if ((rc = make_tensor_map_bf16_box3(&tmO, o, (uint64_t)D, (uint64_t)L, (uint64_t)B, (uint64_t)D, (uint64_t)L * D, DH, LP, C::G, DH * 2)))
    return rc;
  TcArgs a;
  a.B = B; a.L = L; a.heads = heads; a.D = D;
  a.tiles = (int)((B + C::G - 1) / C::G);
  a.scale_log2 = scale * 1.4426950408889634f;
  a.scale = scale;
  a.out_scale = out_scale;
  a.rev = next_stream_dir();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.tiles < sms ? a.tiles : sms;
  cudaError_t le = launch_pdl(pdl_force_all() || a.tiles <= 4 * sms, kern, dim3(grid), dim3(kThreads), (size_t)C::kSmem, st, tmQ, tmO, a);
  if (le != cudaSuccess) {
    set_error("attention_fwd_tc: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("attention_fwd_tc");
}


// ==================================================================================================================
// Backward (dh = 32), same packed tile: per (tile, head) item five products on the tensor pipe
//     S = Q K^T and dP = dO V^T          [128 x 128] each, K = dh                 -> TMEM (single buffer, drained at once)
//     dV = P^T dO,  dK = dS^T Q,  dQ = dS K   [128 x dh] each, K = 128             -> TMEM (double-buffered per group)
// Thread = tile row, so the softmax backward is thread-local too: P_j = exp2(c s_j - c m) / sum, delta = sum_j P_j dP_j,
// dS_j = scale P_j (dP_j - delta).  P and dS are needed TRANSPOSED (keys as the M dimension): one bf16 copy in shared
// memory read through MN-major A descriptors.  Only the diagonal blocks are non-zero -- tile rows 0..63 hold keys 0..63
// only, rows 64..127 keys 64..127 -- so each operand keeps just the two [64 rows x 64 keys] blocks plus ONE 2 KB zero
// block that the descriptor's leading-dimension offset substitutes for the all-zero half of every k-step (18 KB instead of
// 32 KB per operand).  dQ = dS K needs dS un-transposed: that copy lives in tensor memory (A operand from TMEM).
// Warp roles as in the forward: 0 = TMA producer (q|k|v|dO boxes), 1 = S / dP issuer, 10 = dV / dK / dQ issuer,
// 2..5 and 6..9 = two softmax / epilogue groups owning alternate heads.
// ==================================================================================================================
constexpr int kBwdStageBytes = 4 * kBoxBytes;      // q | k | v | dO
constexpr int kBwdStages = 2;
constexpr int kTBlock = 64 * 128;                  // one [64 rows x 64 keys] block of a transposed operand
constexpr int kTOperand = 2 * kTBlock + 2048;      // kb0 | zero block | kb1
constexpr int kBwdOutTile = 128 * 64;              // one staged [128 rows x 32 channels] output tile

template <int LP>
struct BwdCfg {
  static constexpr int G = 128 / LP;
  static constexpr int kSmem = 1024 + kBwdStages * kBwdStageBytes + 2 * kTOperand + 2 * 3 * kBwdOutTile + 256;
};

template <int LP, int LC>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmDQ, const TcArgs a) {
  using C = BwdCfg<LP>;
  constexpr int G = C::G, DH = 32, HB = 2, kStages = kBwdStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStage = base;
  const uint32_t sPT = base + kStages * kBwdStageBytes;     // P  (transposed use): kb0 | Z | kb1
  const uint32_t sST = sPT + kTOperand;                     // dS (transposed use)
  const uint32_t sOut = sST + kTOperand;                    // [group][dq | dk | dv] staging tiles
  const uint32_t sBar = sOut + 2 * 3 * kBwdOutTile;
  uint8_t* pPT = smem + kStages * kBwdStageBytes;
  uint8_t* pST = pPT + kTOperand;
  uint8_t* pOut = pST + kTOperand;
  auto full_bar = [&](int st) { return sBar + 8 * st; };
  auto empty_bar = [&](int st) { return sBar + 8 * (kStages + st); };
  // S / dP and the P / dS operands are single buffers used by the two groups in turn, but every barrier exists once PER
  // GROUP (index = item & 1): a waiter is then never more than one phase away from the barrier it polls
  auto sfull_bar = [&](int b) { return sBar + 8 * (2 * kStages + b); };
  auto sfree_bar = [&](int b) { return sBar + 8 * (2 * kStages + 2 + b); };
  auto pfull_bar = [&](int b) { return sBar + 8 * (2 * kStages + 4 + b); };
  auto pfree_bar = [&](int b) { return sBar + 8 * (2 * kStages + 6 + b); };
  auto ofull_bar = [&](int b) { return sBar + 8 * (2 * kStages + 8 + b); };
  auto ofree_bar = [&](int b) { return sBar + 8 * (2 * kStages + 10 + b); };
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();
  const int nbox = a.heads / HB;
  int my_tiles = 0;
  for (int t = blockIdx.x; t < a.tiles; t += gridDim.x) ++my_tiles;
  const int total = my_tiles * a.heads;

  if (tid == 0) {
    for (int st = 0; st < kStages; ++st) { mbar_init(full_bar(st), 1); mbar_init(empty_bar(st), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(sfull_bar(b), 1); mbar_init(sfree_bar(b), 4);
      mbar_init(pfull_bar(b), 4); mbar_init(pfree_bar(b), 1);
      mbar_init(ofull_bar(b), 1); mbar_init(ofree_bar(b), 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
  {   // transposed operands: everything outside the diagonal blocks (and the zero blocks) stays zero for the whole kernel
    uint4* z = reinterpret_cast<uint4*>(pPT);
    for (int i = tid; i < (2 * kTOperand) >> 4; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  pdl_wait();   // programmatic dependent launch (common.cuh): global memory only from here on
  // tensor memory: S 0..127 | dP 128..255 | OUT[b] 256 + 96 b: dQ, dK, dV (32 columns each) | dS (bf16 pairs) 448..511

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ti = blockIdx.x; ti < a.tiles; ti += gridDim.x) {
        const int t = a.rev ? a.tiles - 1 - ti : ti;
        for (int bx = 0; bx < nbox; ++bx) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), kBwdStageBytes);
          const uint32_t dst = sStage + stage * kBwdStageBytes;
#pragma unroll
          for (int q = 0; q < 3; ++q) tma_load_3d(&tmQ, full_bar(stage), dst + q * kBoxBytes, q * a.D + bx * 64, 0, t * G);
          tma_load_3d(&tmDO, full_bar(stage), dst + 3 * kBoxBytes, bx * 64, 0, t * G);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(128, 128, 0, 0);     // A, B K-major
      constexpr uint32_t idesc_t = make_idesc(128, DH, 1, 1);      // A MN-major (transposed P / dS), B MN-major
      constexpr uint32_t idesc_q = make_idesc(128, DH, 0, 1);      // A from TMEM (K-major), B MN-major
      auto stage_of = [&](int item, uint32_t& ph) {
        const int box = item / HB;
        ph = (uint32_t)((box / kStages) & 1);
        return box % kStages;
      };
      if (warp == 1) {
        for (int i = 0; i < total; ++i) {   // S(i) = Q K^T, dP(i) = dO V^T
          uint32_t ph;
          const int stage = stage_of(i, ph);
          const int slot = i % HB;
          if (slot == 0) mbar_wait(full_bar(stage), ph);
          if (i >= 1) mbar_wait(sfree_bar((i - 1) & 1), (uint32_t)(((i - 1) >> 1) & 1));   // item i - 1 has been read out
          tc_fence_after();
          const uint32_t q_addr = sStage + stage * kBwdStageBytes + slot * (DH * 2);
          const uint32_t k_addr = q_addr + kBoxBytes, v_addr = q_addr + 2 * kBoxBytes, do_addr = q_addr + 3 * kBoxBytes;
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem, make_desc(q_addr + k * 32, 16, 1024), make_desc(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem + 128u, make_desc(do_addr + k * 32, 16, 1024), make_desc(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          umma_commit(sfull_bar(i & 1));
        }
      } else {
        for (int j = 0; j < total; ++j) {   // dV = P^T dO, dK = dS^T Q, dQ = dS K
          uint32_t ph;
          const int stage = stage_of(j, ph);
          const int slot = j % HB, b = j & 1;
          mbar_wait(pfull_bar(b), (uint32_t)((j >> 1) & 1));
          mbar_wait(ofree_bar(b), (uint32_t)(((j >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t q_addr = sStage + stage * kBwdStageBytes + slot * (DH * 2);
          const uint32_t k_addr = q_addr + kBoxBytes, do_addr = q_addr + 3 * kBoxBytes;
          const uint32_t out = tmem + 256u + (uint32_t)(b * 96);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            // A operand of k-step ks (16 tile rows): M group 0 = keys 0..63, group 1 = keys 64..127; the group that is all
            // zero for these rows is the shared zero block (start / leading offset chosen accordingly)
            const uint32_t t_start = ks < 4 ? (uint32_t)(ks * 2048) : (uint32_t)kTBlock;
            const uint32_t t_lbo = ks < 4 ? (uint32_t)(kTBlock - ks * 2048) : (uint32_t)(2048 + (ks - 4) * 2048);
            const uint64_t b_do = make_desc(do_addr + ks * 2048, 8192, 1024);
            const uint64_t b_q = make_desc(q_addr + ks * 2048, 8192, 1024);
            const uint64_t b_k = make_desc(k_addr + ks * 2048, 8192, 1024);
            umma_bf16(out + 64u, make_desc(sPT + t_start, t_lbo, 1024), b_do, idesc_t, ks != 0 ? 1u : 0u);       // dV
            umma_bf16(out + 32u, make_desc(sST + t_start, t_lbo, 1024), b_q, idesc_t, ks != 0 ? 1u : 0u);        // dK
            umma_bf16_ts(out, tmem + 448u + (uint32_t)(ks * 8), b_k, idesc_q, ks != 0 ? 1u : 0u);                 // dQ
          }
          umma_commit(ofull_bar(b));
          umma_commit(pfree_bar(b));
          if (slot == HB - 1) umma_commit(empty_bar(stage));
        }
      }
    }
  } else {
    const int grp = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int g = r / LP;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int L = LC != 0 ? LC : a.L;
    const int key0 = g * LP;
    // transposed operands: tile row r lives in block r / 64 (kb0 at 0, kb1 behind the zero block), its keys at (key0 & 63)
    const int trow = (r >> 6) * (kTBlock + 2048) + (r & 63) * 128;
    const int pchunk0 = (key0 & 63) >> 3;
    const uint32_t ds_taddr = lane_addr + 448u + (uint32_t)(key0 >> 1);
    uint8_t* orow = pOut + grp * 3 * kBwdOutTile + r * 64;
    const int oswz = (r >> 1) & 3;
    const bool elected = ((warp - 2) & 3) == 0 && lane == 0;
    auto grp_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory"); };
    const uint32_t ofull = ofull_bar(grp), ofree = ofree_bar(grp), sfull = sfull_bar(grp), sfree = sfree_bar(grp), pfull = pfull_bar(grp);
    if (grp == 0) {   // zero the dS region of tensor memory once (each warp its lane quarter)
      uint32_t z[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) z[k] = 0u;
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(lane_addr + 448u + (uint32_t)(c * 16), z);
      tmem_wait_st();
      tc_fence_before();
    }
    asm volatile("bar.sync 3, 256;" ::: "memory");   // both groups: the zeroing precedes every dS write
    tc_fence_after();
    int o_tile = blockIdx.x, o_head = grp;
    // epilogue of item j: dQ | dK | dV rows -> bf16 -> staging -> three TMA stores into dqkv
    auto o_phase = [&](int j) {
      const int tile = a.rev ? a.tiles - 1 - o_tile : o_tile, h = o_head;
      o_head += 2;
      if (o_head >= a.heads) { o_head -= a.heads; o_tile += gridDim.x; }
      mbar_wait(ofull, (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t orr[96];
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld32_issue(lane_addr + 256u + (uint32_t)(grp * 96 + c * 32), orr + c * 32);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofree);
      if (elected) bulk_wait_read0();   // the group's previous stores have read the staging tiles
      grp_bar();
#pragma unroll
      for (int t3 = 0; t3 < 3; ++t3)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            w[q] = pack_bf16x2(__uint_as_float(orr[t3 * 32 + 8 * c + 2 * q]), __uint_as_float(orr[t3 * 32 + 8 * c + 2 * q + 1]));
          *reinterpret_cast<uint4*>(orow + t3 * kBwdOutTile + ((c ^ oswz) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      fence_async_smem();
      grp_bar();
      if (elected) {
#pragma unroll
        for (int t3 = 0; t3 < 3; ++t3)
          tma_store_3d(&tmDQ, sOut + (grp * 3 + t3) * kBwdOutTile, t3 * a.D + h * DH, 0, tile * G);
        bulk_commit();
      }
    };
    for (int i = grp; i < total; i += 2) {
      if (i >= 2) o_phase(i - 2);   // drains OUT[grp] before the products of item i need it
      mbar_wait(sfull, (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      uint32_t sr[LP], dr[LP];
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) {
        tmem_ld32_issue(lane_addr + (uint32_t)(key0 + c * 32), sr + c * 32);
        tmem_ld32_issue(lane_addr + 128u + (uint32_t)(key0 + c * 32), dr + c * 32);
      }
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sfree);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < LP; ++j)
        if (j < L) mx = fmaxf(mx, __uint_as_float(sr[j]));
      const float off = mx * a.scale_log2;
      float sum = 0.f, dot = 0.f;
      float pv[LP];
#pragma unroll
      for (int j = 0; j < LP; ++j) {
        pv[j] = j < L ? ex2(fmaf(__uint_as_float(sr[j]), a.scale_log2, -off)) : 0.f;
        sum += pv[j];
        dot = fmaf(pv[j], __uint_as_float(dr[j]), dot);
      }
      const float inv = 1.0f / sum;
      const float delta = dot * inv;                 // sum_j P_j dP_j
      const float sinv = inv * a.scale;              // softmax scale folded into dS
      uint32_t pk[LP / 2], dk[LP / 2];
#pragma unroll
      for (int j = 0; j < LP; j += 2) {
        pk[j >> 1] = pack_bf16x2(pv[j] * inv, pv[j + 1] * inv);
        dk[j >> 1] = pack_bf16x2(pv[j] * sinv * (__uint_as_float(dr[j]) - delta), pv[j + 1] * sinv * (__uint_as_float(dr[j + 1]) - delta));
      }
      if (i >= 1) mbar_wait(pfree_bar((i - 1) & 1), (uint32_t)(((i - 1) >> 1) & 1));   // the products of item i - 1 have read P / dS
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < LP / 8; ++c) {
        const int off16 = ((pchunk0 + c) ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(pPT + trow + off16) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        *reinterpret_cast<uint4*>(pST + trow + off16) = make_uint4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]);
      }
#pragma unroll
      for (int c = 0; c < LP / 32; ++c) tmem_st16(ds_taddr + (uint32_t)(c * 16), dk + c * 16);
      tmem_wait_st();
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull);
    }
    // the group's last item (every earlier one was drained at the top of a later iteration)
    const int last = total > grp ? ((total - 1 - grp) & ~1) + grp : -1;
    if (last >= 0) o_phase(last);
    if (elected) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int LP, int LC>
int launch_bwd(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, float scale, cudaStream_t st) {
  using C = BwdCfg<LP>;
  auto kern = attn_bwd_tc_kernel<LP, LC>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    if (e != cudaSuccess) {
      set_error("attention_bwd_tc: cudaFuncSetAttribute(%d) failed: %s", C::kSmem, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return (int)e;
    }
    configured = true;
  }
  const int D = heads * 32;
  CUtensorMap tmQ, tmDO, tmDQ;
  int rc;
  if ((rc = make_tensor_map_bf16_box3(&tmQ, qkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D, (uint64_t)L * 3 * D, 64, LP, C::G, 128)))
    return rc;
  if ((rc = make_tensor_map_bf16_box3(&tmDO, dO, (uint64_t)D, (uint64_t)L, (uint64_t)B, (uint64_t)D, (uint64_t)L * D, 64, LP, C::G, 128)))
    return rc;
  if ((rc = make_tensor_map_bf16_box3(&tmDQ, dqkv, (uint64_t)3 * D, (uint64_t)L, (uint64_t)B, (uint64_t)3 * D, (uint64_t)L * 3 * D, 32, LP, C::G, 64)))
    return rc;
  TcArgs a;
  a.B = B; a.L = L; a.heads = heads; a.D = D;
  a.tiles = (int)((B + C::G - 1) / C::G);
  a.scale_log2 = scale * 1.4426950408889634f;
  a.scale = scale;
  a.out_scale = nullptr;
  a.rev = next_stream_dir();
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.tiles < sms ? a.tiles : sms;
  cudaError_t le = launch_pdl(pdl_force_all() || a.tiles <= 4 * sms, kern, dim3(grid), dim3(kThreads), (size_t)C::kSmem, st, tmQ, tmDO, tmDQ, a);
  if (le != cudaSuccess) {
    set_error("attention_bwd_tc: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return (int)le;
  }
  return check_launch("attention_bwd_tc");
}

}  // namespace attn_tc

// shapes the tcgen05 forward covers: bf16, dh 32 / 64, L <= 64, whole 128-byte boxes of heads, 16-byte aligned rows
bool attention_tc_supported(int L, int heads, int dh) {
  return (dh == 32 || dh == 64) && L >= 1 && L <= 64 && heads >= 2 && (heads * dh) % 64 == 0;
}

// variant: 0 = probabilities through tensor memory (A operand of P V read from TMEM), 1 = through shared memory
int attention_fwd_tc(const void* qkv, void* o, int64_t B, int L, int heads, int dh, float scale, const float* out_scale, int variant,
                     cudaStream_t st) {
  if (((uintptr_t)qkv & 15) != 0 || ((uintptr_t)o & 15) != 0) {
    set_error("attention_fwd_tc: qkv / o must be 16-byte aligned");
    return AFB_ERR_INVALID;
  }
#define AFB_TC_LAUNCH(LP, DH, LC)                                                                          \
  return variant == 1 ? attn_tc::launch<LP, DH, false, LC>(qkv, o, B, L, heads, scale, out_scale, st)      \
                      : attn_tc::launch<LP, DH, true, LC>(qkv, o, B, L, heads, scale, out_scale, st)
  // the dataset lengths get their own instantiation (22 joints / 32 frames on SHREC and DHG, 46 / 64 on LMDHG)
  if (dh == 32) {
    if (L == 22) AFB_TC_LAUNCH(32, 32, 22);
    if (L == 32) AFB_TC_LAUNCH(32, 32, 32);
    if (L < 32) AFB_TC_LAUNCH(32, 32, 0);
    if (L == 46) AFB_TC_LAUNCH(64, 32, 46);
    if (L == 64) AFB_TC_LAUNCH(64, 32, 64);
    AFB_TC_LAUNCH(64, 32, 0);
  }
  if (L == 32) AFB_TC_LAUNCH(32, 64, 32);
  if (L < 32) AFB_TC_LAUNCH(32, 64, 0);
  if (L == 64) AFB_TC_LAUNCH(64, 64, 64);
  AFB_TC_LAUNCH(64, 64, 0);
#undef AFB_TC_LAUNCH
}

// backward on the same paths: dh 32, even head count (two heads per 128-byte box)
bool attention_bwd_tc_supported(int L, int heads, int dh) { return dh == 32 && L >= 1 && L <= 64 && heads >= 2 && heads % 2 == 0; }

int attention_bwd_tc(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int heads, int dh, float scale, cudaStream_t st) {
  if (((uintptr_t)qkv & 15) != 0 || ((uintptr_t)dO & 15) != 0 || ((uintptr_t)dqkv & 15) != 0) {
    set_error("attention_bwd_tc: operands must be 16-byte aligned");
    return AFB_ERR_INVALID;
  }
  (void)dh;
  if (L == 22) return attn_tc::launch_bwd<32, 22>(qkv, dO, dqkv, B, L, heads, scale, st);
  if (L == 32) return attn_tc::launch_bwd<32, 32>(qkv, dO, dqkv, B, L, heads, scale, st);
  if (L < 32) return attn_tc::launch_bwd<32, 0>(qkv, dO, dqkv, B, L, heads, scale, st);
  if (L == 46) return attn_tc::launch_bwd<64, 46>(qkv, dO, dqkv, B, L, heads, scale, st);
  if (L == 64) return attn_tc::launch_bwd<64, 64>(qkv, dO, dqkv, B, L, heads, scale, st);
  return attn_tc::launch_bwd<64, 0>(qkv, dO, dqkv, B, L, heads, scale, st);
}

}  // namespace afb
