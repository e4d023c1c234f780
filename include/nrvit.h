/* libnrvit — C ABI of the B200-native ViT encoder hot path.
 *
 * The reference (RandallBalestriero/noise-robust-vit) has no FFI: its hot path is the Python
 * nn.Module API of vit_pytorch_robust/simple_vit.py and vit_pytorch_robust/vit.py, which bottoms
 * out in ATen library calls.  This header is the boundary a maintainer binds instead (ctypes
 * stub shown in INTEGRATION.md): every entry point names the reference op(s) it replaces.
 *
 * Conventions
 *  - plain C, no torch / C++ types; every entry returns 0 on success, <0 = NRV_E* and
 *    nrv_last_error() (thread-local) says why.  There is NO CPU fallback: without an sm_100
 *    device nrv_init fails and every compute entry returns NRV_ENOTINIT.
 *  - all compute entries are asynchronous on the `stream` argument (a cudaStream_t passed as
 *    void*), allocate nothing on the device, never synchronise, and are CUDA-graph capturable.
 *  - the caller owns every buffer (activations, stash, workspace, parameters, gradients).
 *  - activations are NRV_BF16 (production: bf16 operands, fp32 accumulate) or NRV_F32 (check mode:
 *    fp32 activations and weights, GEMMs on the same tcgen05 pipeline with 3xTF32 split operands,
 *    i.e. fp32-grade products with fp32 accumulation).  LayerNorm statistics, biases, affine
 *    parameters, gradients of parameters and optimiser state are always fp32.
 *  - matrices are row-major with an explicit leading dimension (in elements).
 */
#ifndef NRVIT_H_
#define NRVIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRV_ABI_VERSION 7

/* status codes */
#define NRV_OK 0
#define NRV_EINVAL (-1)
#define NRV_ECUDA (-2)
#define NRV_EARCH (-3)
#define NRV_ENOTINIT (-4)
#define NRV_ENOTIMPL (-5)

/* dtypes */
#define NRV_BF16 0
#define NRV_F32 1

/* operand layouts for nrv_gemm */
#define NRV_K_MAJOR 0  /* row-major [rows, K]  (K contiguous)        */
#define NRV_MN_MAJOR 1 /* row-major [K, rows]  (M or N contiguous)   */

/* GEMM epilogues */
#define NRV_EPI_STORE 0      /* out = alpha*acc (+bias) (+pos) (+residual)                    */
#define NRV_EPI_GELU 1       /* out2 = u = alpha*acc+bias ; out = gelu_erf(u) (+residual)      */
#define NRV_EPI_DGELU 2      /* out = (alpha*acc) * gelu_erf'(aux)                             */
#define NRV_EPI_GELU_GRAD 4  /* out = gelu_erf(u) ; out2 = gelu_erf'(u)   (training forward)   */
#define NRV_EPI_MUL 5        /* out = (alpha*acc) * aux                   (its backward)       */
#define NRV_EPI_ATOMIC_F32 3 /* out(fp32) += alpha*acc, split-K with red.global.add            */

/* attention normalisation (reference: nn.Softmax vs utils.SinkhornAttention, simple_vit.py:56-59) */
#define NRV_ATTN_SOFTMAX 0
#define NRV_ATTN_SINKHORN3 1

/* attention implementation selector */
#define NRV_ATTN_IMPL_AUTO 0  /* tcgen05 kernel for bf16 when the shape is supported, else SIMT */
#define NRV_ATTN_IMPL_SIMT 1  /* fp32 CUDA-core kernel (check mode / cross-check)               */
#define NRV_ATTN_IMPL_TC 2    /* tcgen05 + TMEM kernel (bf16 only)                              */

/* pooling before the head (simple_vit.py:146 mean ; vit.py:347 class token) */
#define NRV_POOL_MEAN 0
#define NRV_POOL_CLS 1

/* patch flattening order: SimpleViT '(p1 p2 c)' (simple_vit.py:127-129) vs Conv2d weight
 * '(c p1 p2)' (vit.py:237-242) */
#define NRV_PATCH_P1P2C 0
#define NRV_PATCH_CP1P2 1

/* ---------------------------------------------------------------------------------------------
 * runtime
 * ------------------------------------------------------------------------------------------- */
int nrv_abi_version(void);
/* Binds the library to CUDA device `device`; fails with NRV_EARCH unless it is compute capability
 * 10.x.  Idempotent. */
int nrv_init(int device);
const char* nrv_last_error(void);
int nrv_num_sms(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
long long nrv_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GEMM: out[M,N] = epilogue(alpha * A[M,K] * B[N,K]^T), tcgen05 + TMEM + TMA.
 * Replaces aten::mm/addmm behind every nn.Linear fwd/bwd on the path
 * (simple_vit.py:37-42,61-62,76,130,136 ; vit.py:41-47,237-242,265 ; utils.py:115,579).
 * ------------------------------------------------------------------------------------------- */
typedef struct nrv_gemm_desc {
  int M, N, K;
  int dtype;          /* NRV_BF16 operands, or NRV_F32 operands (3xTF32 split; needs workspace) */
  int out_dtype;      /* NRV_BF16 or NRV_F32 (ATOMIC_F32 implies F32) */
  const void* a; long long lda; int a_layout;
  const void* b; long long ldb; int b_layout;
  int epi;
  float alpha;
  void* out; long long ldo;
  void* out2;                        /* GELU: pre-activation (same ld/dtype as out), may be NULL */
  const float* bias;                 /* [N] fp32 or NULL */
  const void* residual; long long ldr; /* [M,N] same dtype as out, or NULL */
  const void* aux; long long ldaux;  /* DGELU: pre-activation [M,N], same dtype as out */
  /* optional token-row remap + positional table (patch embedding):
   *   GEMM row r = (b, p) with p < pos_rows_in  ->  out row b*pos_rows_out + p + pos_row_off,
   *   out += pos[p + pos_row_off, :]   (pos may be NULL: remap only) ; disabled if pos_rows_in==0 */
  const float* pos; long long ldpos; int pos_rows_in, pos_rows_out, pos_row_off;
  int splits;        /* ATOMIC_F32 only: K splits, 0 = auto */
  int force_bn128;   /* testing / tuning: use the 128-wide N tile */
  int force_single_cta; /* testing / tuning: never use the CTA-pair (cta_group::2) kernel */
  void* workspace; size_t workspace_bytes; /* NRV_F32 only: >= nrv_gemm_workspace_bytes(M,N,K) */
  float* colsum;     /* EPI_MUL with bf16 output only, may be NULL: colsum[N] += column sums of the stored output
                        (fp32 reds from the epilogue: the bias gradient of the Linear whose dX this GEMM computes) */
  int tile_mode;     /* testing / tuning: 0 = auto, 1 = 256x256 units (one accumulator, double-buffered in TMEM),
                        2 = 512x256 units (two row blocks share the B tile; both accumulators live) */
  /* LayerNorm folded into the product (the "fused LayerNorm + projection GEMM" of the forward pass; reference ops
   * simple_vit.py:65-67,38-39 ; vit.py:123,128).  With A = the raw rows x and B = the ROW-CENTRED gamma o W
   * (nrv_ln_fold_weights: B[n,k] = gamma_k W[n,k] - mean_k(gamma o W[n,:]), every row summing to exactly zero), the
   * accumulator already is sum_k (x_k - mu) gamma_k W[n,k], so
   *   LayerNorm(x) W^T + b  =  rstd_m acc_mn + c_n,   c_n = sum_k beta_k W[n,k] + b_n   (pass c as `bias`)
   * ln_stats: fp64 [M][2] = (sum_k x, sum_k x^2) of every row of A over its K_ln columns (nrv_rowstats, or the
   * stats_out of the GEMM that produced A).  STORE / GELU / GELU_GRAD epilogues, TMA path. */
  const double* ln_stats; float ln_eps; int K_ln;
  float* ln_mean_out; float* ln_rstd_out; /* optional fp32 [M]: mean and rstd of every row (needed by nrv_layernorm_bwd) */
  /* optional fp64 [M][2], STORE epilogue (with or without residual): += (sum, sum of squares) of every output row (the
   * fp32 values before the store rounds them), i.e. the ln_stats of the next product.  The caller zeroes it.  fp64 so
   * that the order in which the column blocks arrive cannot change the fp32 mean / rstd derived from it. */
  double* stats_out;
} nrv_gemm_desc;

int nrv_gemm(const nrv_gemm_desc* d, void* stream);
size_t nrv_gemm_workspace_bytes(int M, int N, int K, int dtype);
/* Measurement hook (bench.py roofline): while enabled, every GEMM launch is bracketed by CUDA events
 * on its own stream; nrv_gemm_timing_read waits for them and returns the summed kernel time (ms),
 * the summed 2*M*N*K FLOPs and the number of launches since enabling.  Not capturable in a graph. */
int nrv_gemm_timing(int enable);
int nrv_gemm_timing_read(double* ms, double* flops, long long* launches);
/* per-launch records {M, N, K, epi | a_layout<<4 | b_layout<<5, microseconds}; returns the count */
int nrv_gemm_timing_detail(long long* out, int max_records);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm (aten::native_layer_norm fwd/bwd; simple_vit.py:38,54,136 ; vit.py:104,115,167)
 * x,y,dy,dx,dres: `dtype` [rows, dim]; gamma/beta/dgamma/dbeta fp32 [dim]; mean/rstd fp32 [rows].
 * bwd: dx = LN'(dy) (+ dres, the residual-branch gradient, may be NULL);
 *      dgamma/dbeta are ACCUMULATED (+=) ; if colsum != NULL, colsum[dim] += column sums of the
 *      produced dx (the bias gradient of the Linear that feeds the residual stream).
 * ------------------------------------------------------------------------------------------- */
int nrv_layernorm_fwd(const void* x, const float* gamma, const float* beta, float eps, void* y,
                      float* mean, float* rstd, long long rows, int dim, int dtype, void* stream);
/* xn_out (optional, `dtype` [rows, dim], needs beta): also writes the normalised rows (x - mean) rstd gamma + beta, which
 * the forward pass does not keep when its LayerNorm is folded into the projection GEMM (nrv_gemm_desc.ln_stats) and
 * which the weight-gradient GEMM of that projection reads. */
int nrv_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                      const float* gamma, const void* dres, void* dx, float* dgamma, float* dbeta,
                      float* colsum, const float* beta, void* xn_out, long long rows, int dim, int dtype,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Pieces of the folded LayerNorm (see nrv_gemm_desc.ln_stats):
 *   nrv_rowstats: stats fp64 [rows][2] = (sum_k x, sum_k x^2) of every row of x `dtype` [rows, dim]  (overwrites)
 *   nrv_ln_fold_weights: Wf[n,k] = gamma_k W[n,k] - mean_k(gamma o W[n,:]) (`dtype` [rows, ldw], K = dim columns; after
 *   rounding to `dtype` one small element per row absorbs the remainder so that the STORED row sums to zero to ~1e-6 of
 *   its scale: the mean of x then cancels inside the accumulation whatever |mean| / std is),
 *   c[n] = sum_k beta_k W[n,k] + bias[n] (bias may be NULL) */
int nrv_rowstats(const void* x, long long rows, int dim, int dtype, double* stats, void* stream);
int nrv_ln_fold_weights(const void* W, const float* gamma, const float* beta, const float* bias, void* Wf, float* c,
                        int rows, int dim, long long ldw, int dtype, void* stream);
size_t nrv_layernorm_bwd_workspace(long long rows, int dim);

/* out[cols] += sum over rows of x[rows, cols] (`dtype` in, fp32 out): Linear bias gradients */
int nrv_colsum(const void* x, long long ldx, long long rows, int cols, int dtype, float* out,
               void* workspace, size_t workspace_bytes, void* stream);
size_t nrv_colsum_workspace(long long rows, int cols);

/* ---------------------------------------------------------------------------------------------
 * Patch extraction (einops Rearrange 'b c (h p1) (w p2) -> b h w (p1 p2 c)', simple_vit.py:127 ;
 * the im2col implicit in Conv2d(k=s=P), vit.py:237-242,323): img [B,C,H,W] fp32 or bf16 ->
 * patches `out_dtype` [B*nh*nw, ld] with zero padding in columns [C*ph*pw, ld).
 * ------------------------------------------------------------------------------------------- */
int nrv_im2col(const void* img, int img_dtype, int B, int C, int H, int W, int ph, int pw,
               int order, void* patches, int out_dtype, long long ld, void* stream);
/* Patch embedding with the im2col FUSED into the projection GEMM through TMA (reference ops: vit.py:237-242 conv_proj,
 * :323-331 reshape / permute, :174 + pos_embedding): no patch matrix is materialised.  img: bf16 [B, C, H, W]
 * (16-byte aligned), w: bf16 [D, C*ph*pw] in the conv weight's (c p1 p2) order, out: bf16 token rows
 *   out[b * tokens_per_img + tok_off + py*(W/pw) + px, :] = patch(b, py, px) . w^T + bias + pos[tok_off + py*(W/pw) + px, :]
 * (bias, pos fp32, either may be NULL; rows outside [tok_off, tok_off + patches) are not written: the class token).
 * nrv_patch_embed_bwd_weight: dw[D, C*ph*pw] (fp32) += dx^T . patches with dx: bf16 token rows laid out like `out`.
 * Shapes: nrv_patch_embed_supported() (bf16, NRV_PATCH_CP1P2, patch width 16/32/64 with ph a multiple of 64/pw,
 * C*ph*pw > 128, D > 128); everything else goes through nrv_im2col + nrv_gemm. */
int nrv_patch_embed_supported(int C, int H, int W, int ph, int pw, int patch_order, int dtype, int img_dtype, int D);
int nrv_patch_embed_fwd(const void* img, int B, int C, int H, int W, int ph, int pw, const void* w, long long ldw,
                        const float* bias, const float* pos, long long ldpos, int tokens_per_img, int tok_off, void* out,
                        long long ldo, int D, void* stream);
int nrv_patch_embed_bwd_weight(const void* img, int B, int C, int H, int W, int ph, int pw, const void* dx, long long lddx,
                               int tokens_per_img, int tok_off, float* dw, long long lddw, int D, void* stream);
/* nn.Dropout of the encoder (vit.py:45,47,109,125,166 ; README ViT dropout / emb_dropout), training mode:
 * out[i] = x[i] * keep_i / (1 - p) (+ residual[i]); keep_i is a pure function of (seed, layer, site, i)
 * (Philox4x32-10), so the backward pass regenerates the mask.  Sites: NRV_DROP_*; layer = -1 for the embedding.
 * x, residual, out: `dtype` [n] (n % 8 == 0), residual may be NULL, out may alias x. */
#define NRV_DROP_ATTN_OUT 0 /* after the attention output projection (vit.py:125)  */
#define NRV_DROP_FC1 1      /* after GELU (vit.py:45)                              */
#define NRV_DROP_FC2 2      /* after the second MLP Linear (vit.py:47)             */
#define NRV_DROP_EMB 3      /* after the positional embedding (vit.py:174)         */
#define NRV_DROP_ATTN_PROB 4 /* attention probabilities [B,H,N,N] (vit.py:105-110) */
int nrv_dropout(const void* x, const void* residual, void* out, long long n, int dtype, float p,
                unsigned long long seed, int layer, int site, void* stream);
/* Noisy-input objective of the examples (x + std * randn_like(x), examples/nowak.py:153,196) in one pass:
 * out[i] = x[i] + stddev * N(0,1), Philox4x32-10 + Box-Muller keyed by (seed, i).  x, out: `dtype` [n], n % 8 == 0,
 * out may alias x. */
int nrv_add_gaussian_noise(const void* x, void* out, long long n, int dtype, float stddev, unsigned long long seed,
                           void* stream);
/* Fixed 2-D sin/cos table of SimpleViT (posemb_sincos_2d, simple_vit.py:15-28): out fp32 [h*w, dim],
 * token t = y*w + x, out[t] = [sin(x w_j), cos(x w_j), sin(y w_j), cos(y w_j)], w_j =
 * temperature^(-j/(dim/4-1)); dim % 4 == 0 and dim > 4 required (the reference asserts the former
 * and divides by zero on dim == 4). */
int nrv_posemb_sincos_2d(float* out, int h, int w, int dim, float temperature, void* stream);
/* x[b, 0, :] = cls[:] + pos[0, :]  (vit.py:341-342,174) */
int nrv_cls_token_fwd(const float* cls, const float* pos, void* x, int B, int tokens, int dim,
                      int dtype, void* stream);
/* dpos[t,:] += sum_b dx[b,t,:] (t in [0,tokens)), dcls[:] += sum_b dx[b,0,:] ; either may be NULL */
int nrv_posemb_bwd(const void* dx, int B, int tokens, int dim, int dtype, float* dpos, float* dcls,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Attention core (simple_vit.py:70-75 ; utils.py:207-232 as intended = softmax(QK^T/sqrt(dh))V).
 * qkv: `dtype` [B, N, 3, H, dh] (the packed projection output, q|k|v then head-major);
 * out: `dtype` [B, N, H*dh]; lse: fp32 [B, H, N] (log-sum-exp of the scaled scores) for
 * NRV_ATTN_SOFTMAX, fp32 [B, H, 8, N] (lse + the 7 Sinkhorn normalisation vectors) for
 * NRV_ATTN_SINKHORN3 — nrv_attn_stats_elems() floats.
 * bwd: dqkv [B, N, 3, H, dh] from dout, recomputing P from q,k and lse.
 * impl: NRV_ATTN_IMPL_AUTO picks a tcgen05 kernel for bf16 (dh == 64 up to 208 / 256 tokens: the fused training kernels;
 * forward dh <= 128 up to 384 tokens and backward dh <= 80 up to 1024 tokens: the general kernels), else the SIMT one.
 * ------------------------------------------------------------------------------------------- */
int nrv_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale,
                 int mode, int dtype, int impl, void* workspace, size_t workspace_bytes, void* stream);
int nrv_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                 int B, int N, int H, int dh, float scale, int mode, int dtype, int impl,
                 void* workspace, size_t workspace_bytes, void* stream);
/* Profiling aid: when non-NULL, CTA 0 of the tcgen05 attention kernels writes clock64() phase
 * timestamps of its first 64 tiles to device_buf[tile][8] (int64). NULL disables. */
int nrv_attn_debug_timestamps(long long* device_buf);
/* scratch for nrv_attn_bwd: softmax delta = rowsum(dO o O), fp32 [B, H, N]; Sinkhorn: per-CTA N x N gradient matrix
 * (plus the probability matrix when it does not fit in shared memory); general tcgen05 backward: per-CTA running dQ */
size_t nrv_attn_bwd_workspace(int B, int N, int H, int dh);
/* scratch for nrv_attn_fwd / nrv_attn_probs: 0 unless mode is NRV_ATTN_SINKHORN3 (or probabilities are requested)
 * and the N x N fp32 matrix of a head does not fit in shared memory (N > ~204 at dh = 64); workspace may then be NULL */
size_t nrv_attn_fwd_workspace(int B, int N, int H, int dh, int mode);
size_t nrv_attn_stats_elems(int B, int N, int H, int mode);
/* Introspection (reference: recorder.py:28-31 hooks the output of Attention.attend): probs fp32 [B, H, N, N] =
 * softmax(q k^T * scale) for NRV_ATTN_SOFTMAX, followed by the 3 Sinkhorn iterations for NRV_ATTN_SINKHORN3
 * (utils.py:1031-1037), from the packed projection output qkv [B, N, 3, H, dh].  stats: scratch of
 * nrv_attn_stats_elems() floats; workspace: nrv_attn_fwd_workspace(B, N, H, dh, NRV_ATTN_SINKHORN3) bytes.
 * Debug path on CUDA cores: not used by forward / backward. */
int nrv_attn_probs(const void* qkv, float* probs, float* stats, int B, int N, int H, int dh, float scale, int mode,
                   int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pooling + loss.
 *  pool:   x [B, N, D] -> pooled [B, D]  (mean over tokens, or token 0)
 *  pool bwd: dpooled [B, D] -> dx [B, N, D] (overwrites; CLS: zeros elsewhere)
 *  softmax-CE with label smoothing (F.cross_entropy, examples/baseline.py:70):
 *          logits fp32 [B, ldl>=C]; labels int64 [B]; loss_mean (fp32 scalar, overwritten, may be
 *          NULL); dlogits `dl_dtype` [B, ldd>=C] = grad_scale * dloss/dlogits, zero in columns >= C
 *          (may be NULL).
 * ------------------------------------------------------------------------------------------- */
int nrv_pool_fwd(const void* x, void* pooled, int B, int N, int D, int pool, int dtype, void* stream);
int nrv_pool_bwd(const void* dpooled, void* dx, int B, int N, int D, int pool, int dtype,
                 void* stream);
int nrv_softmax_ce(const float* logits, long long ldl, const long long* labels,
                   float label_smoothing, float* loss_mean, void* dlogits, int dl_dtype,
                   long long ldd, float grad_scale, int B, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused AdamW over flat fp32 buffers (torch.optim.AdamW semantics, examples/CIFAR100.py:90-97)
 * + bf16 shadow refresh.  p,m,v,g: fp32 [n]; shadow: bf16 [n] or NULL.
 * g is multiplied by grad_scale (1/world, loss-scale) and, if grad_scale_dev != NULL, by
 * *grad_scale_dev (device scalar: the global-norm clip coefficient).  step >= 1.
 * ------------------------------------------------------------------------------------------- */
int nrv_adamw(float* p, float* m, float* v, const float* g, void* shadow, long long n, float lr,
              float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
              const float* grad_scale_dev, void* stream);
/* fp32 -> bf16 cast of a flat buffer (shadow refresh after load_state_dict) */
int nrv_cast_bf16(const float* src, void* dst, long long n, void* stream);
/* the inverse: bf16 -> fp32 (gradient buckets exchanged in bf16, parallel.py) */
int nrv_cast_f32(const void* src, float* dst, long long n, void* stream);
/* out[0] += sum(g^2) (for clip_grad_norm_); coef[0] = min(1, max_norm/(sqrt(sumsq)*extra+1e-6)) */
int nrv_sumsq(const float* g, long long n, float* out, void* stream);
int nrv_clip_coef(const float* sumsq, float max_norm, float extra_scale, float* coef, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-encoder entries: one C call runs every kernel of a forward pass, or of a range of
 * backward stages.  (Transformer.forward simple_vit.py:93-97 ; Encoder.forward vit.py:169-175 ;
 * SimpleViT.forward simple_vit.py:138-149 ; VisionTransformer.forward vit.py:335-351 up to the
 * classifier, and their autograd.)
 * ------------------------------------------------------------------------------------------- */
typedef struct nrv_vit_config {
  int batch, channels, img_h, img_w, patch_h, patch_w;
  int dim, depth, heads, dim_head, mlp_dim;
  int cls_token;     /* 1: prepend class token (VisionTransformer), 0: SimpleViT */
  int pool;          /* NRV_POOL_* */
  int patch_order;   /* NRV_PATCH_* */
  int qkv_bias;      /* 1: in_proj_bias / out_proj.bias present (vit.py) ; 0: SimpleViT */
  float ln_eps;      /* 1e-5 SimpleViT, 1e-6 VisionTransformer */
  int attn_mode;     /* NRV_ATTN_* */
  int attn_impl;     /* NRV_ATTN_IMPL_* */
  int img_dtype;     /* NRV_F32 or NRV_BF16 input images */
  int dtype;         /* activation / weight-matrix dtype: NRV_BF16 or NRV_F32 (check mode) */
  int training;      /* 1: fill the activation stash for backward */
  /* dropout (active only when training = 1 and the probability is > 0; masks: see nrv_dropout) */
  float p_drop;      /* after out-proj, GELU and FC2 (VisionTransformer `dropout`, README ViT `dropout`) */
  float p_emb_drop;  /* after the positional embedding (VisionTransformer `dropout`, README ViT `emb_dropout`) */
  float p_attn_drop; /* on the attention probabilities (VisionTransformer `attention_dropout`, README ViT `dropout`):
                        softmax attention only; bf16: the general tcgen05 attention kernels draw the mask (dh <= 80,
                        <= 384 tokens), otherwise (fp32 check mode, larger shapes) the fp32 CUDA-core attention kernels */
  unsigned long long drop_seed;
  int ln_mode;       /* NRV_LN_FOLDED (0, default): the LayerNorms in front of the QKV and FC1 projections are folded into
                        those GEMMs (nrv_gemm_desc.ln_stats; no LayerNorm kernel and no normalised copy of the stream in the
                        forward layer loop); NRV_LN_SEPARATE (1): stand-alone LayerNorm kernels.  Training with p_drop > 0
                        always runs NRV_LN_SEPARATE (the dropout kernels produce the stream there). */
} nrv_vit_config;
#define NRV_LN_FOLDED 0
#define NRV_LN_SEPARATE 1

/* Per-layer parameters: weight matrices in cfg.dtype (bf16 shadows refreshed by nrv_adamw, or the
 * fp32 masters in check mode), vectors fp32.  The same struct carries gradients (all fp32,
 * accumulated with +=). */
typedef struct nrv_vit_layer {
  void* w_qkv;  /* [3*I, D] */
  void* w_out;  /* [D, I]   */
  void* w_fc1;  /* [M, D]   */
  void* w_fc2;  /* [D, M]   */
  float *ln1_g, *ln1_b, *b_qkv, *b_out, *ln2_g, *ln2_b, *b_fc1, *b_fc2; /* b_qkv/b_out may be NULL */
} nrv_vit_layer;

typedef struct nrv_vit_params {
  void* w_patch;   /* params: cfg.dtype [D, patch_ld] (patch_ld = patch_dim rounded up to 8, zero
                      padded) ; grads: fp32 [D, patch_ld] */
  float* b_patch;  /* [D] */
  float* pos;      /* fp32 [tokens, D]: learned pos_embedding or the sincos table; grads: NULL for
                      the fixed sincos table */
  float* cls;      /* fp32 [D] or NULL */
  float *lnf_g, *lnf_b; /* final LayerNorm (linear_head.0 / encoder.ln) */
  nrv_vit_layer* layers; /* [depth] */
} nrv_vit_params;

/* bytes of caller-provided scratch: `stash` persists from forward to backward (training),
 * `workspace` is transient within one call */
size_t nrv_vit_stash_bytes(const nrv_vit_config* cfg);
size_t nrv_vit_workspace_bytes(const nrv_vit_config* cfg);
/* Where a training=1 forward left a tensor inside its stash (introspection: extractor.py:50-59 reads the
 * transformer's tokens, recorder.py:28-31 the attention probabilities, recomputed from the stashed projections
 * with nrv_attn_probs):  NRV_STASH_STREAM, index k in [0, 2*depth]: residual stream [B*N, D] entering layer k/2
 * (even k), between its two branches (odd k), index 2*depth = the transformer's output;
 * NRV_STASH_QKV, index = layer: packed projections [B, N, 3, H, dh].  Element type cfg.dtype. */
#define NRV_STASH_STREAM 0
#define NRV_STASH_QKV 1
int nrv_vit_stash_tensor(const nrv_vit_config* cfg, int what, int index, size_t* offset, size_t* bytes);

/* img [B,C,H,W] -> feat cfg.dtype [B, D]: final-LayerNorm'ed pooled token (mean over tokens, or the
 * class-token row).  LayerNorm is row-wise, so VisionTransformer's encoder.ln followed by x[:,0]
 * (vit.py:175,347) equals LN of the class-token row; SimpleViT: x.mean(1) then linear_head.0
 * (simple_vit.py:136,146-149).  The classifier Linear is one nrv_gemm on feat. */
int nrv_vit_forward(const nrv_vit_config* cfg, const nrv_vit_params* params, const void* img,
                    void* feat, void* stash, void* workspace, void* stream);
/* Backward stages, run from stage_hi down to stage_lo (inclusive):
 *   depth   : dfeat [B, D] -> final LN bwd + pool bwd -> gradient of the last layer's output
 *   depth-1 .. 0 : transformer layers
 *   -1      : patch embedding / class token / positional embedding (`img`: the forward pass's images -- the weight
 *             gradient gathers its patches from them by TMA when the forward did (nrv_patch_embed_supported shapes); for the
 *             other layouts the patch matrix of the forward pass is kept in the stash and `img` may be NULL)
 * Splitting the range lets the caller start the gradient all-reduce of finished layers while
 * earlier layers are still running.  Parameter gradients are accumulated into grads->*. */
int nrv_vit_backward(const nrv_vit_config* cfg, const nrv_vit_params* params,
                     const nrv_vit_params* grads, const void* img, const void* dfeat, void* stash,
                     void* workspace, int stage_hi, int stage_lo, void* stream);
/* One-shot: the NEXT nrv_vit_backward call of this thread records `cuda_event` (a cudaEvent_t) on its stream at the point
 * where every parameter gradient of stages >= stage_lo is final EXCEPT the ln1 gamma / beta of layer stage_lo (their
 * LayerNorm backward, the last kernel of the stage, is still to come) -- or at the end of the call when there is no such
 * point (stage_lo = -1, folded LayerNorm).  Data-parallel training makes the bucket's all-reduce wait for this event, so
 * the collective starts under that LayerNorm backward instead of beside the next stage's persistent GEMM.  NULL clears. */
int nrv_vit_backward_marker(void* cuda_event);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange: in-place SUM all-reduce of one bucket of the flat gradient buffer over
 * the GPUs of the box, NCCL over NVLink / NVSwitch, asynchronous on `stream`.  The reference has no counterpart (an
 * external trainer wraps the model in torch DDP: examples/evaluation.py:137-138); parallel.py launches one call per
 * finished group of backward stages on a side stream.  NCCL is bound at run time (the libnccl.so.2 already in the
 * process), so single-GPU users never need it.  Rendezvous: rank 0 calls nrv_comm_get_unique_id and ships the
 * nrv_comm_unique_id_bytes() bytes to the other ranks by any means (parallel.py: torch.distributed.broadcast).
 * max_ctas > 0 caps the CTAs NCCL may use per collective (the backward kernels are persistent and fill the SMs).
 * nrv_comm_register: registers a buffer once so that NCCL can use it in place (NVLS / zero-copy paths).
 * ------------------------------------------------------------------------------------------- */
typedef struct nrv_comm nrv_comm;
int nrv_comm_unique_id_bytes(void);
int nrv_comm_get_unique_id(void* id_out, int bytes);
int nrv_comm_init(const void* id, int bytes, int nranks, int rank, int max_ctas, nrv_comm** out);
int nrv_comm_register(nrv_comm* c, void* buf, size_t bytes, void** handle);
int nrv_comm_deregister(nrv_comm* c, void* handle);
int nrv_comm_allreduce_bucket(nrv_comm* c, void* buf, long long count, int dtype, void* stream);
int nrv_comm_nccl_version(void);
int nrv_comm_destroy(nrv_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* NRVIT_H_ */
