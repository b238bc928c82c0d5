/* altformer_b200 -- C ABI of the B200-native (sm_100a) hot path of ST-GCN-AltFormer.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (as void*); none takes a
 * torch type.  All functions return 0 on success, a negative AFB_ERR_* code for argument errors, or
 * a positive cudaError_t for launch errors; afb_last_error() returns a human readable message.
 * There is no CPU fallback: every pointer must be device memory of an sm_100 GPU.
 *
 * The reference is pure Python/PyTorch, so "the interface each entry replaces" is the ATen call
 * sequence of a reference nn.Module method.  Paths are relative to the reference root.
 *
 * Activations are channels-last token matrices [M, C] with M = N*T*V in (n, t, v) order (the
 * layout `rearrange(x, 'b c f p -> (b f) p c')` of model/AltFormer/model_ST.py:152 produces), in
 * AFB_BF16 (performance mode) or AFB_F32 (parity mode).
 */
#ifndef ALTFORMER_B200_H
#define ALTFORMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFB_F32 0
#define AFB_BF16 1

#define AFB_ERR_INVALID (-1)
#define AFB_ERR_UNSUPPORTED (-2)
#define AFB_ERR_DRIVER (-3)

#define AFB_ACT_NONE 0
#define AFB_ACT_GELU 1     /* out = gelu_erf(v); if C2 != NULL the pre-activation v is stored there */
#define AFB_ACT_GELU_BWD 2 /* out = v * gelu'(aux) */
#define AFB_ACT_RELU 3
#define AFB_RS_VALUE 0
#define AFB_RS_BIAS 1

typedef void* afb_stream;

int afb_version(void);
const char* afb_last_error(void);
/* 0 if device `dev` is an sm_100 part, else AFB_ERR_UNSUPPORTED. */
int afb_device_ok(int dev);

/* ------------------------------------------------------------------------------------------ *
 * Tensor-core GEMMs (tcgen05.mma + TMEM accumulators + TMA operand staging).
 * ------------------------------------------------------------------------------------------ */

/* C[m, n] = epilogue( sum_k A[m (+tap shift), k] * B[n, k] )            (nn.Linear / Conv2d (k x 1))
 * Replaces: F.linear at model/AltFormer/model_ST.py:27-31,51,64,153,182,204 and the temporal
 * Conv2d of model/net.py:50 (taps > 1: implicit GEMM, tap t reads A rows shifted by
 * (t - tap_pad) * tap_row_stride inside each batch, zero outside -- the conv's zero padding).
 *   A: bf16 [batches][rows_per_batch][lda]     B: bf16 [N][ldb]  (K-major), or, if b_mn_major,
 *   B: bf16 [K][ldb] with N contiguous (used for dX = dY * W with W stored [N_out, K_in]).
 * K = taps * k_per_tap; k_per_tap % 64 == 0 when taps > 1; N % 64 == 0; lda, ldb % 8 == 0.
 * epilogue(v): v = alpha*acc + bias[n] + pos[(m % pos_rows), n]; activation; v *= row_scale[m / row_scale_div];
 *              v += residual[m, n]; store as out_dtype.  Optional pointers may be NULL. */
typedef struct {
  const void* A;
  const void* B;
  void* C;
  void* C2; /* optional pre-activation output (same dtype/ld as C) */
  int64_t rows_per_batch;
  int32_t batches;
  int32_t N;
  int32_t k_per_tap;
  int32_t taps;
  int32_t tap_row_stride;
  int32_t tap_pad;
  int32_t lda, ldb, ldc;
  int32_t b_mn_major;
  int32_t out_dtype;
  int32_t act;
  float alpha;
  const float* bias;
  const float* pos;
  int32_t pos_rows;
  const void* aux;
  int32_t aux_dtype;
  int32_t ldaux;
  const void* residual;
  int32_t res_dtype;
  int32_t ldres;
  const float* row_scale;
  int32_t row_scale_div;
  int32_t row_scale_mode; /* AFB_RS_VALUE: v *= row_scale (above).  AFB_RS_BIAS: the A rows already carry the scale
                             (DropPath folded into the saved activation), so only the bias term is scaled:
                             v = alpha*acc + row_scale*bias[n] (+ residual). */
} afb_gemm_tn_t;
int afb_gemm_tn(const afb_gemm_tn_t* p, afb_stream s);

/* dW[n1*ld1 + n2*ld2] += alpha * sum_m G[m, n1] * X[m (+ x_row_shift), n2]      (weight gradients)
 * Replaces: the autograd weight-gradient GEMMs of nn.Linear / Conv2d.  Contraction runs over token
 * rows (both operands MN-major for the MMA), split across CTAs, accumulated with fp32 atomics into
 * dW (which the caller zeroes, cf. model.zero_grad() at SHREC/ST_TS/train_sttran.py:187).
 *   G: bf16 [batches][rows_per_batch][ldg], X: bf16 [batches][rows_per_batch][ldx];
 *   N1 % 64 == 0 (or N1 <= 128), N2 % 64 == 0. */
typedef struct {
  const void* G;
  const void* X;
  float* dW;
  int64_t rows_per_batch;
  int32_t batches;
  int32_t N1, N2;
  int32_t ldg, ldx;
  int64_t ld1, ld2;
  int32_t x_row_shift;
  float alpha;
  float* dbias; /* optional: dbias[n1] += alpha * sum_m G[m, n1] (bias gradient: column sums of the G tiles taken
                   from shared memory while the MMAs run, or one extra N=16 MMA against an all-ones operand) */
  int32_t taps;           /* > 1 (<= 3, needs N1 <= 128 and N2 <= 128): one launch produces `taps` gradients that share G:  */
  int32_t tap_row_stride; /*   dW[t * tap_dw_stride + ...] uses X rows shifted by x_row_shift + t * tap_row_stride          */
  int64_t tap_dw_stride;  /*   (the k x 1 temporal conv: taps of one weight read the same dY tile). 0 / 1 = single tap.     */
  const float* dbias_row_scale; /* optional (needs N1 % 256 == 0): dbias[n1] += alpha * sum_m rs[m / row_scale_div] * G[m, n1]
                                   -- the DropPath keep factor of the bias path; dW itself is NOT scaled */
  int32_t row_scale_div;
} afb_gemm_dw_t;
int afb_gemm_dw(const afb_gemm_dw_t* p, afb_stream s);

/* Generic strided CUDA-core GEMM for tiny / odd shapes (classifier head, test cross-checks):
 * C[i*sci + j*scj] = alpha * sum_k A[i*sai + k*sak] * B[j*sbj + k*sbk] (+ bias[j]) (+ beta*C). */
typedef struct {
  const void* A;
  const void* B;
  void* C;
  const float* bias;
  int32_t M, N, K;
  int64_t sai, sak, sbj, sbk, sci, scj;
  int32_t a_dtype, b_dtype, c_dtype;
  float alpha, beta;
} afb_gemm_simt_t;
int afb_gemm_simt(const afb_gemm_simt_t* p, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * Parameter preparation
 * ------------------------------------------------------------------------------------------ */
int afb_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, afb_stream s);
/* strided 2-D cast-copy: dst[r*ldd + c] = (dst_dtype) src[r*lds + c]  (weight stacking) */
int afb_copy2d(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd, int64_t rows,
               int cols, afb_stream s);
/* dst[c*rows + r] = bf16(src[r*cols + c]) */
int afb_cast_transpose(const float* src, void* dst_bf16, int rows, int cols, afb_stream s);
/* Conv2d weight (co, ci, k, 1) fp32 -> fwd bf16 [co][k][ci] and bwd bf16 [ci][k'][co] with k' = k-1-tap
 * (the flipped kernel of the transposed conv that yields dX).  Either output may be NULL. */
int afb_conv_weight_pack(const float* w, void* fwd_bf16, void* bwd_bf16, int co, int ci, int k, afb_stream s);
/* dW (co, ci, k) += tap-major weight gradient tmp [k][co][ci] (what afb_gemm_dw with taps > 1 produces with contiguous
 * rows, so its epilogue can use 16-byte vector reductions); Conv2d weight gradient of model/net.py:50. */
int afb_conv_dw_unpack(const float* tmp, float* dW, int co, int ci, int k, afb_stream s);
/* hi/lo bf16 split of an fp32 matrix for the 3-pass fp32-parity GEMM: dst is [rows][3*cols] holding
 * (hi | lo | hi) when which == 0 (A side) or (hi | hi | lo) when which == 1 (B side); which == 2 stacks
 * (hi ; hi ; lo) along rows instead ([3*rows][cols], the MN-major B operand of dX = dY * W). */
int afb_split3(const float* src, void* dst_bf16, int64_t rows, int cols, int which, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * LayerNorm (nn.LayerNorm at model_ST.py:75,80,96,141), one warp per token row.
 * ------------------------------------------------------------------------------------------ */
int afb_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta, void* y, int y_dtype,
                      float* mean, float* rstd, int64_t rows, int D, float eps, afb_stream s);
/* dx = (dres ? dres : 0) + LN'(dy);  dgamma/dbeta are ACCUMULATED (atomics) -- caller zeroes. */
int afb_layernorm_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, const float* gamma,
                      const float* mean, const float* rstd, const void* dres, int dres_dtype, void* dx,
                      int dx_dtype, float* dgamma, float* dbeta, int64_t rows, int D, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * Small-sequence multi-head attention (model_ST.py:49-67): qkv [B*L, 3*D] with feature index
 * s*D + h*dh + d -> o [B*L, D].  L <= 256 (tensor-core kernels for L <= 64 with dh 32 / 64 in bf16, CUDA-core kernels
 * otherwise; backward needs 16 * L * (dh + 1) bytes of shared memory <= ~220 KB), dh <= 64.  Softmax probabilities never leave
 * the SM.  out_scale (optional, [B]) multiplies whole output sequences (DropPath keep factor; a
 * sequence with factor 0 is written as zeros without being computed).
 * ------------------------------------------------------------------------------------------ */
int afb_attention_fwd(const void* qkv, void* o, int dtype, int64_t B, int L, int heads, int dh, float scale,
                      const float* out_scale, afb_stream s);
int afb_attention_bwd(const void* qkv, const void* dO, void* dqkv, int dtype, int64_t B, int L, int heads,
                      int dh, float scale, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * BatchNorm2d (+ReLU, + residuals) on token matrices -- model/net.py:52-54, model/unit_agcn.py:91-93.
 * ------------------------------------------------------------------------------------------ */
/* sum[c] += sum_m x[m,c]; sumsq[c] += sum_m x[m,c]^2   (fp64 accumulators, caller zeroes) */
int afb_colstats(const void* x, int dtype, int64_t M, int C, int ldx, double* sum, double* sumsq, afb_stream s);
/* colsum[c] += sum_m row_scale[m/div] * x[m,c]  (bias gradients; fp32 atomics, caller zeroes) */
int afb_colsum(const void* x, int dtype, int64_t M, int C, int ldx, const float* row_scale, int row_scale_div,
               float* out, afb_stream s);
/* From fp64 sums: batch mean / biased var -> mean, rstd; scale = gamma*rstd, shift = beta - mean*scale;
 * running stats updated with momentum and the unbiased variance (training != 0).  With
 * training == 0, mean/var come from the running buffers. num_batches_tracked is host-side. */
int afb_bn_finalize(const double* sum, const double* sumsq, int64_t M, int C, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                    int training, float* mean, float* rstd, float* scale, float* shift, afb_stream s);
/* y = relu?( x*scale + shift + res_pre ) + res_post;  y2 (optional) = same values with rows permuted
 * from (n,t,v) to (n,v,t) order (TS stage input, model_TS.py:161). */
int afb_bn_act_fwd(const void* x, int x_dtype, const float* scale, const float* shift, const void* res_pre,
                   const void* res_post, int res_dtype, int relu, void* y, void* y2, int y_dtype, int64_t M,
                   int C, int T, int V, afb_stream s);
/* Backward.  g = (dy[m] + dy2[perm(m)]) * [pre-activation > 0]; the ReLU mask is recomputed from
 * x, mean/rstd/gamma/beta (+ res_pre), so the forward output need not be kept.  dy or dy2 may be NULL
 * (dy2 is the gradient arriving in (n,v,t) row order from the TS stage).
 * pass 1: dgamma[c] += sum g*xhat, dbeta[c] += sum g   (fp32 atomics, caller zeroes) */
int afb_bn_bwd_reduce(const void* dy, const void* dy2, int dy_dtype, const void* x, int x_dtype, const void* res_pre,
                      int res_dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                      int relu, float* dgamma, float* dbeta, int64_t M, int C, int T, int V, afb_stream s);
/* pass 2: dx = gamma*rstd*(g - dbeta/M - xhat*dgamma/M) (training) or gamma*rstd*g (eval);
 * dres (optional) = g, the gradient of res_pre. */
int afb_bn_bwd_apply(const void* dy, const void* dy2, int dy_dtype, const void* x, int x_dtype, const void* res_pre,
                     int res_dtype, const float* mean, const float* rstd, const float* gamma, const float* beta,
                     const float* dgamma, const float* dbeta, int relu, int training, void* dx, void* dres,
                     int dx_dtype, int64_t M, int C, int T, int V, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * Pooling between / after the stages (model_ST.py:167-172,194-202; model_TS.py:169-174,194-199)
 * ------------------------------------------------------------------------------------------ */
int afb_pool_mean_fwd(const void* x, void* y, int dtype, int64_t B, int L, int D, afb_stream s);
int afb_pool_mean_bwd(const void* dy, void* dx, int dtype, int64_t B, int L, int D, afb_stream s);
int afb_pool_max_fwd(const void* x, void* y, int32_t* argmax, int dtype, int64_t B, int L, int D, afb_stream s);
int afb_pool_max_bwd(const void* dy, const int32_t* argmax, void* dx, int dtype, int64_t B, int L, int D,
                     afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * Loss and optimizer (SHREC/ST_TS/train_sttran.py:157,161,185-191)
 * ------------------------------------------------------------------------------------------ */
/* loss[0] += mean CE; dlogits = (softmax - onehot) / N.  logits fp32 [N, C]. */
int afb_softmax_ce(const float* logits, const int64_t* labels, float* loss, float* dlogits, int N, int C,
                   afb_stream s);
/* Fused multi-tensor AdamW over one flat fp32 buffer; also refreshes the bf16 shadow copy.
 * `step` is a device counter (1-based value used; incremented by afb_step_inc). */
int afb_adamw(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const int32_t* step,
              float lr, float beta1, float beta2, float eps, float wd, float grad_scale, afb_stream s);
int afb_step_inc(int32_t* step, afb_stream s);
/* y = a*x (elementwise), and row-scaled copies for DropPath backward */
int afb_scale_rows(const void* x, void* y, int dtype, int64_t M, int C, const float* row_scale, int div,
                   afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * gcn0 = unit_agcn(3 -> C_out): model/unit_agcn.py:73-93 with C_in <= 4.
 * x fp32 [N, T, V, Cin] (the (N,T,V,3) skeleton batch itself; no permute copy).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const float* x;      /* [N][T][V][3] */
  const float* A;      /* [3][V][V] static adjacency */
  const float* PA;     /* [3][V][V] learned */
  const float* Wa[3];  /* conv_a[i].weight [IC][3] */
  const float* ba[3];  /* [IC] */
  const float* Wb[3];  /* conv_b[i].weight [IC][3] */
  const float* bb[3];
  const float* Wd[3];  /* conv_d[i].weight [Cout][3] */
  const float* bd[3];  /* [Cout] */
  const float* Wdn;    /* down.0.weight [Cout][3] */
  const float* bdn;    /* [Cout] */
  const float* bn_g;   /* bn.weight / bn.bias [Cout] */
  const float* bn_b;
  const float* dn_g;   /* down.1.weight / bias */
  const float* dn_b;
  float* bn_rm;        /* running stats (updated when training) */
  float* bn_rv;
  float* dn_rm;
  float* dn_rv;
  int32_t N, T, V, Cout, IC;
  int32_t training;
  float momentum, eps;
  float* Mmat;         /* out [N][3][V][V]: softmax_u(S_i) + A_i + PA_i (saved for backward) */
  double* moments;     /* persistent workspace [3][AFB_GCN0_SLOTS][AFB_GCN0_NMOM], ZERO before the first call; the
                          kernels re-zero it ([0]: two-kernel path; [1], [2]: the fused kernel's ping-pong halves),
                          so one buffer serves every launch on a stream */
  int32_t* counter;    /* persistent, zero-initialised int32[8]: CTA ticket (word 0) / fused kernel's grid-barrier words (4..6)
                          (re-armed by the kernels) */
  float* stats;        /* out [AFB_GCN0_NSTAT_BASE + 4*Cout]: E[r] (12), Cov(r) (144), mean_h, rstd_h, mean_d, rstd_d */
  float* Wfold;        /* out [Cout][16]: BN-folded weights of the apply pass */
  void* Aop;           /* out (optional) bf16 operand copy of M for the tensor-core passes, afb_gcn0_aop_bytes(N, V)
                          bytes: two-kernel path [N][3][VP][VP+8] (A_i[v][u] = M_i[u][v], VP = V rounded up to 16);
                          fused kernel [N][2 (hi, lo)][3*CB][40] with column c = i*CB + v, CB = V rounded up to 8 */
  float* colsum;       /* out (optional) [N][3][VP]: sum_u M_i[u][v] */
  void* Wfrag;         /* out (optional) uint32 [Cout/8][32][2]: Wfold as mma.m16n8k16 B fragments (bf16 pairs);
                          the tensor-core apply pass needs Aop, colsum and Wfrag (V <= 48) */
  void* y;             /* out [N*T*V][Cout] */
  int32_t y_dtype;
  int32_t precise;     /* 0: bf16 tensor-core apply (fused kernel when it covers the shape, else mma.sync two-kernel path);
                          1: fp32 FMA apply (parity mode); 2: exact ReLU masks wanted in bf16 mode -- fused kernel (hi/lo
                          operands) when it covers the shape, else the fp32 FMA apply with bf16 output */
} afb_gcn0_fwd_t;
#define AFB_GCN0_NR 12          /* r = (z_0, z_1, z_2, x): 9 + 3 */
#define AFB_GCN0_NMOM 96        /* 12 first + 78 second moments (upper triangle), padded */
#define AFB_GCN0_SLOTS 32       /* fp64 accumulation slots */
#define AFB_GCN0_NSTAT_BASE 160 /* 12 + 144, padded */
int afb_gcn0_fwd(const afb_gcn0_fwd_t* p, afb_stream s);
/* bytes the caller allocates for afb_gcn0_fwd_t::Aop */
int64_t afb_gcn0_aop_bytes(int N, int V);
/* profiling aid: globaltimer stamps (ns) of the last fused-kernel launch, 8 per CTA for the first `ctas` CTAs:
 * entry, operands staged, M done, r rows done, moments posted, barrier passed, weights folded, stores issued */
int afb_gcn0_fused_stamps(uint64_t* out, int ctas);

/* Backward (parameter gradients only: gcn0's input is the data tensor, model/AltFormer/
 * ST_GCN_AltFormer.py:70, so dx is never required).  Training-mode BN only. */
typedef struct {
  afb_gcn0_fwd_t f;     /* same pointers as forward (x, weights, Mmat, stats, y = forward output) */
  const void* dy;       /* [N*T*V][Cout], dtype f.y_dtype */
  float* ws;            /* workspace [AFB_GCN0_BWD_WS(Cout)] floats, zeroed by the call */
  /* gradient outputs, ACCUMULATED (caller zeroes) */
  float* dPA;
  float* dWa[3]; float* dba[3]; float* dWb[3]; float* dbb[3]; float* dWd[3]; float* dbd[3];
  float* dWdn; float* dbdn; float* dbn_g; float* dbn_b; float* ddn_g; float* ddn_b;
} afb_gcn0_bwd_t;
#define AFB_GCN0_BWD_WS(cout) (32 * (cout) + 256)
int afb_gcn0_bwd(const afb_gcn0_bwd_t* p, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * general unit_agcn(C -> C_out) pieces (model/unit_agcn.py:80-88), C % 8 == 0:
 * theta/phi come from afb_gemm_tn on x [M,C] with the stacked weight [6*IC, C];
 * ------------------------------------------------------------------------------------------ */
/* S_i[n,u,v] = sum_{c,t} th_i[n,t,u,c]*ph_i[n,t,v,c]/(IC*T); M_i = softmax_u(S_i) + A_i + PA_i.
 * thph: [N*T*V][ld >= 6*IC] = (th_0, th_1, th_2, ph_0, ph_1, ph_2, pad).  Writes P (softmax) and Mmat. */
int afb_agcn_scores_fwd(const void* thph, int dtype, int ld, const float* A, const float* PA, float* P,
                        float* Mmat, int N, int T, int V, int IC, afb_stream s);
/* z[n,t,v,(i,c)] = sum_u x[n,t,u,c] * M_i[n,u,v]   -> [M][3*C] */
int afb_agcn_aggregate_fwd(const void* x, const float* Mmat, void* z, int dtype, int N, int T, int V, int C,
                           afb_stream s);
/* dx[n,t,u,c] (+)= sum_{i,v} dz[n,t,v,(i,c)] * M_i[n,u,v];  dM_i[n,u,v] = sum_{t,c} x[n,t,u,c]*dz[n,t,v,(i,c)] */
int afb_agcn_aggregate_bwd(const void* x, const void* dz, const float* Mmat, void* dx, int accumulate,
                           float* dM, int dtype, int N, int T, int V, int C, afb_stream s);
/* dPA += sum_n dM; dS = P*(dM - colsum(P*dM)); dthph from dS (scaled by 1/(IC*T)). */
int afb_agcn_scores_bwd(const void* thph, int ld, const float* P, const float* dM, float* dPA, void* dthph,
                        int dtype, int N, int T, int V, int IC, afb_stream s);

/* Tensor-core (mma.sync) versions of the two aggregate steps for bf16 activations, V <= 48, C % 64 == 0 (csrc/agcn_mma.cu).
 * aggregate_fwd_mma: split == 0 -> z [M, 3C] bf16; split == 1 -> z [M, 9C] = (hi | lo | hi) slabs of 3C columns, the A operand
 * of the 3-term (hi/lo) conv_d GEMM of the exact-mask forward (M is applied as bf16 hi + lo).
 * aggregate_bwd_mma: dx (+)= sum_{i,v} dz M_i ; dM_i = sum_{t,c} x dz (written, not accumulated). */
int afb_agcn_aggregate_fwd_mma(const void* x, const float* Mmat, void* z, int split, int N, int T, int V, int C, afb_stream s);
int afb_agcn_aggregate_bwd_mma(const void* x, const void* dz, const float* Mmat, void* dx, int accumulate, float* dM,
                               int N, int T, int V, int C, afb_stream s);
/* scores on tensor cores (V <= 48, IC % 16 == 0): thph bf16, or fp32 applied as bf16 hi + lo (exact-mask forward); the backward
 * takes / produces bf16. */
int afb_agcn_scores_fwd_mma(const void* thph, int dtype, int ld, const float* A, const float* PA, float* P, float* Mmat,
                            int N, int T, int V, int IC, afb_stream s);
int afb_agcn_scores_bwd_mma(const void* thph, int ld, const float* P, const float* dM, float* dPA, void* dthph,
                            int N, int T, int V, int IC, afb_stream s);

/* ------------------------------------------------------------------------------------------ *
 * Input streams + ensemble (data_process/Hand_Dataset.py:183-217; SHREC/ST_TS/emsemble.py:217-218)
 * ------------------------------------------------------------------------------------------ */
int afb_bone_stream(const float* x, const int32_t* parent, float* y, int64_t NT, int V, afb_stream s);
int afb_motion_stream(const float* x, float* y, int N, int T, int V, afb_stream s);
/* y = x - x[:, 0, joint, :] per sequence: palm-centre normalisation (data_process/Hand_Dataset.py:61, joint = 1); out of place */
int afb_palm_center(const float* x, float* y, int N, int T, int V, int joint, afb_stream s);
/* Hand_Dataset.data_aug for a batch on the device (data_process/Hand_Dataset.py:84-157): kind[n] in {0 scale, 1 shift, 2 noise on
 * four joints, 3 time_interpolate, other = copy}, params[n][16] = factor | offset xyz | 4 joint ids + 4 x xyz | r; out of place */
int afb_augment(const float* x, float* y, int64_t N, int T, int V, const int* kind, const float* params, afb_stream s);
int afb_axpby(const float* a, float wa, const float* b, float wb, float* out, int64_t n, afb_stream s);
/* Phase tensors of a strided temporal convolution (Unit2D stride (s, 1), model/net.py:24-27; TCN_GCN_unit stride 2 and its
 * down1, model/ST_TR/ST_TR_new.py:362-374).  scatter == 0: dst [N, To, V, C] = src[n, j * stride + phase, v, :] (zero past T);
 * scatter == 1: dst [N, T, V, C] (frames j * stride + phase) = src [N, To, V, C]. */
int afb_frame_phase(const void* src, void* dst, int dtype, int scatter, int N, int T, int To, int V, int C, int stride, int phase,
                    afb_stream s);
/* y = a * b elementwise (same dtype; n % 4 == 0): applies an explicit dropout mask, nn.Dropout of model/net.py:40,48 */
int afb_mul(const void* a, const void* b, void* y, int dtype, int64_t n, afb_stream s);

#ifdef __cplusplus
}
#endif
#endif /* ALTFORMER_B200_H */
