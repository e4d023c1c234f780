// Whole-encoder orchestration: one C call issues every kernel of a ViT forward pass, or of a range
// of backward stages, on one stream.  No allocation, no synchronisation: CUDA-graph capturable.
//
// Reference graph being replaced (forward; backward is its autograd):
//   SimpleViT.forward   simple_vit.py:138-149   (Transformer.forward :93-97, Attention.forward :64-76,
//                                                FeedForward :37-42)
//   VisionTransformer.forward vit.py:335-351    (_process_input :308-333, Encoder.forward :169-175,
//                                                EncoderBlock.forward :118-130, MLPBlock :35-47)
//
// Kernel sequence per layer (forward): LN -> GEMM(qkv,+bias) -> attention -> GEMM(out,+bias,+residual)
//   -> LN -> GEMM(fc1,+bias,GELU; keeps pre-activation) -> GEMM(fc2,+bias,+residual)
// (backward): dW2 | dX2 with GELU' epilogue | dW1 | dX1 | LN-bwd(+residual grad, +bias colsum)
//   | dWo | dXo | attention-bwd | dWqkv | dXqkv | LN-bwd(+residual grad, +bias colsum)
// Weight gradients accumulate straight into the caller's fp32 gradient buffers (split-K red.add).
#include "common.cuh"
#include "nrvit_internal.h"

namespace nrv {

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Dims {
  int B, n, N, D, I, M, L, H, dh, pdim, pld, esz, dtype, attn_mode;
  long long T;
  float p_drop, p_emb, p_attn;  // 0 unless training
  unsigned long long seed;
  bool fold;                    // LayerNorm folded into the QKV / FC1 GEMMs (ln_fold.cu): no xn1 / xn2 in the stash
};

static int make_dims(const nrv_vit_config* c, Dims* d) {
  NRV_REQUIRE(c != nullptr, "nrv_vit: null config");
  NRV_REQUIRE(c->dtype == NRV_BF16 || c->dtype == NRV_F32, "nrv_vit: cfg.dtype must be NRV_BF16 or NRV_F32");
  NRV_REQUIRE(c->img_dtype == NRV_BF16 || c->img_dtype == NRV_F32, "nrv_vit: cfg.img_dtype must be NRV_BF16 or NRV_F32");
  NRV_REQUIRE(c->batch > 0 && c->channels > 0 && c->patch_h > 0 && c->patch_w > 0 && c->depth > 0 && c->heads > 0,
              "nrv_vit: batch, channels, patch size, depth and heads must be positive");
  NRV_REQUIRE(c->img_h % c->patch_h == 0 && c->img_w % c->patch_w == 0,
              "Image dimensions must be divisible by the patch size.");
  NRV_REQUIRE(c->dim % 8 == 0 && c->mlp_dim % 8 == 0 && (c->heads * c->dim_head) % 8 == 0,
              "nrv_vit: dim, mlp_dim and heads*dim_head must be multiples of 8");
  d->B = c->batch;
  d->n = (c->img_h / c->patch_h) * (c->img_w / c->patch_w);
  d->N = d->n + (c->cls_token ? 1 : 0);
  d->D = c->dim;
  d->H = c->heads;
  d->dh = c->dim_head;
  d->I = c->heads * c->dim_head;
  d->M = c->mlp_dim;
  d->L = c->depth;
  d->pdim = c->channels * c->patch_h * c->patch_w;
  d->pld = (d->pdim + 7) / 8 * 8;
  d->dtype = c->dtype;
  d->attn_mode = c->attn_mode;
  d->esz = c->dtype == NRV_BF16 ? 2 : 4;
  d->T = (long long)d->B * d->N;
  NRV_REQUIRE(c->p_drop >= 0.f && c->p_drop < 1.f && c->p_emb_drop >= 0.f && c->p_emb_drop < 1.f &&
              c->p_attn_drop >= 0.f && c->p_attn_drop < 1.f, "nrv_vit: dropout probabilities must be in [0, 1)");
  d->p_drop = c->training ? c->p_drop : 0.f;
  d->p_emb = c->training ? c->p_emb_drop : 0.f;
  d->p_attn = c->training ? c->p_attn_drop : 0.f;
  d->seed = c->drop_seed;
  NRV_REQUIRE(c->ln_mode == NRV_LN_FOLDED || c->ln_mode == NRV_LN_SEPARATE, "nrv_vit: bad ln_mode %d", c->ln_mode);
  // with dropout the residual stream is produced by the dropout kernels (no GEMM epilogue to emit the row statistics);
  // a pure function of the configuration, so forward and backward agree on the stash layout
  d->fold = c->ln_mode == NRV_LN_FOLDED && d->p_drop == 0.f;
  return NRV_OK;
}

// ---- stash layout (training): everything backward needs -------------------------------------
struct LayerStash {
  size_t xn1, mean1, rstd1, qkv, o, lse, xn2, mean2, rstd2, u, h;
};
struct StashPlan {
  size_t xs0;          // residual stream: xs[k], k = 0..2L, each [T, D]
  size_t xs_stride;
  size_t layer0, layer_stride;
  LayerStash l;        // offsets within a layer block
  size_t pooled, meanf, rstdf;
  size_t patches;      // [T, pld] patch matrix in token-row layout (class-token rows zero): A of the forward GEMM, B of dW
  size_t total;
};

static StashPlan plan_stash(const Dims& d) {
  StashPlan p{};
  size_t off = 0;
  const size_t TD = align_up((size_t)d.T * d.D * d.esz);
  p.xs0 = off; p.xs_stride = TD; off += TD * (2 * (size_t)d.L + 1);
  size_t lo = 0;
  auto take = [&](size_t bytes) { size_t r = lo; lo += align_up(bytes); return r; };
  p.l.xn1 = take(d.fold ? 0 : (size_t)d.T * d.D * d.esz);
  p.l.mean1 = take((size_t)d.T * 4);
  p.l.rstd1 = take((size_t)d.T * 4);
  p.l.qkv = take((size_t)d.T * 3 * d.I * d.esz);
  p.l.o = take((size_t)d.T * d.I * d.esz);
  p.l.lse = take(nrv_attn_stats_elems(d.B, d.N, d.H, d.attn_mode) * 4);
  p.l.xn2 = take(d.fold ? 0 : (size_t)d.T * d.D * d.esz);
  p.l.mean2 = take((size_t)d.T * 4);
  p.l.rstd2 = take((size_t)d.T * 4);
  p.l.u = take((size_t)d.T * d.M * d.esz);
  p.l.h = take((size_t)d.T * d.M * d.esz);
  p.layer0 = off; p.layer_stride = lo; off += lo * (size_t)d.L;
  p.pooled = off; off += align_up((size_t)d.B * d.D * d.esz);
  p.meanf = off; off += align_up((size_t)d.B * 4);
  p.rstdf = off; off += align_up((size_t)d.B * 4);
  p.patches = off; off += align_up((size_t)d.T * d.pld * d.esz);
  p.total = off;
  return p;
}

// ---- workspace layout (transient) -------------------------------------------------------------
struct WorkPlan {
  size_t patches;                       // [T, pld] (bwd uses all T rows, fwd the first B*n)
  size_t dxa, dxb, dxn, dqkv, dob, du;  // backward gradient buffers
  size_t dxm;                           // dropout only: branch output before the mask (forward), masked gradient (backward)
  size_t dpooled;
  size_t attn_ws, attn_ws_bytes;
  size_t red;                           // LN-bwd / colsum partials
  size_t red_bytes;
  size_t gemm_ws, gemm_ws_bytes;        // check-mode operand split
  size_t infer;                         // inference-only: one layer block + 3 residual buffers
  // folded LayerNorm (forward): row statistics of xs[0 .. 2L] ([2L+1][T][2] fp64), per layer the centred gamma o W of the QKV and
  // FC1 projections with their c vectors; (backward) the normalised rows LN-bwd hands to the weight-gradient GEMM
  size_t stats, stats_bytes, fold_w, fold_w_layer, fold_v, fold_v_layer, xnb;
  size_t total;
};

static WorkPlan plan_work(const Dims& d, bool training) {
  WorkPlan w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += align_up(bytes); return r; };
  w.patches = take((size_t)d.T * d.pld * d.esz);
  if (training) {
    w.dxa = take((size_t)d.T * d.D * d.esz);
    w.dxb = take((size_t)d.T * d.D * d.esz);
    w.dxn = take((size_t)d.T * d.D * d.esz);
    w.dqkv = take((size_t)d.T * 3 * d.I * d.esz);
    w.dob = take((size_t)d.T * d.I * d.esz);
    w.du = take((size_t)d.T * d.M * d.esz);
    w.dpooled = take((size_t)d.B * d.D * d.esz);
    if (d.p_drop > 0.f) w.dxm = take((size_t)d.T * d.D * d.esz);
    w.attn_ws_bytes = nrv_attn_bwd_workspace(d.B, d.N, d.H, d.dh);
  }
  {
    // forward scratch of the Sinkhorn kernels when the N x N matrix of a head does not fit in shared memory
    const size_t f = nrv_attn_fwd_workspace(d.B, d.N, d.H, d.dh, d.attn_mode);
    if (f > w.attn_ws_bytes) w.attn_ws_bytes = f;
    w.attn_ws = take(w.attn_ws_bytes);
  }
  size_t red = nrv_layernorm_bwd_workspace(d.T, d.D);
  const int widest = d.M > 3 * d.I ? d.M : 3 * d.I;
  const size_t cs = nrv_colsum_workspace(d.T, widest);
  if (cs > red) red = cs;
  w.red_bytes = red;
  w.red = take(red);
  size_t g = 0;
  if (d.dtype == NRV_F32) {
    auto upd = [&](long long M, long long N, long long K) {
      const size_t b = gemm_workspace_bytes((int)M, (int)N, (int)K, NRV_F32);
      if (b > g) g = b;
    };
    upd(d.T, d.D, d.pld); upd(d.T, 3 * d.I, d.D); upd(d.T, d.D, d.I); upd(d.T, d.M, d.D); upd(d.T, d.D, d.M);
    if (training) {
      upd(d.T, d.I, d.D); upd(d.T, d.D, 3 * d.I);
      upd(d.D, d.pld, d.T); upd(3 * d.I, d.D, d.T); upd(d.D, d.I, d.T); upd(d.M, d.D, d.T); upd(d.D, d.M, d.T);
    }
  }
  w.gemm_ws_bytes = g;
  w.gemm_ws = take(g);
  if (!training) {
    const StashPlan sp = plan_stash(d);
    w.infer = take(sp.layer_stride + 3 * sp.xs_stride + 3 * align_up((size_t)d.B * d.D * d.esz));
  }
  if (d.fold) {
    w.stats_bytes = (size_t)(2 * d.L + 1) * d.T * 2 * sizeof(double);
    w.stats = take(w.stats_bytes);
    w.fold_w_layer = align_up((size_t)3 * d.I * d.D * d.esz) + align_up((size_t)d.M * d.D * d.esz);
    w.fold_w = take(w.fold_w_layer * (size_t)d.L);
    w.fold_v_layer = align_up((size_t)(3 * d.I + d.M) * sizeof(float));
    w.fold_v = take(w.fold_v_layer * (size_t)d.L);
    if (training) w.xnb = take((size_t)d.T * d.D * d.esz);
  }
  w.total = off;
  return w;
}

// resolves buffer addresses for training (stash) and inference (workspace, buffers reused)
struct Bufs {
  uint8_t* stash; uint8_t* work; StashPlan sp; WorkPlan wp; bool training;
  uint8_t* xs(int k) const {
    if (training) return stash + sp.xs0 + sp.xs_stride * (size_t)k;
    return work + wp.infer + sp.layer_stride + sp.xs_stride * (size_t)(k % 3);
  }
  uint8_t* layer(int l, size_t off) const {
    if (training) return stash + sp.layer0 + sp.layer_stride * (size_t)l + off;
    return work + wp.infer + off;
  }
  uint8_t* tail(int which) const {  // 0 pooled, 1 meanf, 2 rstdf
    const size_t slot = align_up((size_t)1);  // 256
    if (training) return stash + (which == 0 ? sp.pooled : (which == 1 ? sp.meanf : sp.rstdf));
    (void)slot;
    uint8_t* base = work + wp.infer + sp.layer_stride + 3 * sp.xs_stride;
    const size_t bd = (sp.meanf - sp.pooled);
    return base + (which == 0 ? 0 : (which == 1 ? bd : 2 * bd));
  }
};

// buffers of the folded LayerNorm inside the workspace
struct Fold {
  const Dims& d; uint8_t* work; const WorkPlan& wp;
  double* stats(int k) const { return reinterpret_cast<double*>(work + wp.stats) + (size_t)k * d.T * 2; }
  void* w_qkv(int l) const { return work + wp.fold_w + wp.fold_w_layer * (size_t)l; }
  void* w_fc1(int l) const { return work + wp.fold_w + wp.fold_w_layer * (size_t)l + align_up((size_t)3 * d.I * d.D * d.esz); }
  float* c_qkv(int l) const { return reinterpret_cast<float*>(work + wp.fold_v + wp.fold_v_layer * (size_t)l); }
  float* c_fc1(int l) const { return c_qkv(l) + 3 * d.I; }
};

struct Gemm {
  nrv_gemm_desc d;
  Gemm(const Dims& dm, const Bufs& bf, long long M, long long N, long long K) {
    memset(&d, 0, sizeof(d));
    d.M = (int)M; d.N = (int)N; d.K = (int)K;
    d.dtype = dm.dtype; d.out_dtype = dm.dtype;
    d.alpha = 1.f;
    d.workspace = bf.work + bf.wp.gemm_ws; d.workspace_bytes = bf.wp.gemm_ws_bytes;
  }
  Gemm& A(const void* p, long long ld, int layout = NRV_K_MAJOR) { d.a = p; d.lda = ld; d.a_layout = layout; return *this; }
  Gemm& Bm(const void* p, long long ld, int layout = NRV_K_MAJOR) { d.b = p; d.ldb = ld; d.b_layout = layout; return *this; }
  Gemm& out(void* p, long long ld) { d.out = p; d.ldo = ld; return *this; }
  Gemm& bias(const float* b) { d.bias = b; return *this; }
  Gemm& residual(const void* r, long long ld) { d.residual = r; d.ldr = ld; return *this; }
  Gemm& gelu(void* pre) { d.epi = NRV_EPI_GELU; d.out2 = pre; return *this; }
  Gemm& dgelu(const void* pre, long long ld) { d.epi = NRV_EPI_DGELU; d.aux = pre; d.ldaux = ld; return *this; }
  Gemm& gelu_grad(void* grad) { d.epi = NRV_EPI_GELU_GRAD; d.out2 = grad; return *this; }
  Gemm& mul(const void* m, long long ld) { d.epi = NRV_EPI_MUL; d.aux = m; d.ldaux = ld; return *this; }
  Gemm& colsum(float* c) { d.colsum = c; return *this; }
  Gemm& atomic() { d.epi = NRV_EPI_ATOMIC_F32; d.out_dtype = NRV_F32; return *this; }
  // LayerNorm folded into this product: A = raw rows, B = gamma o W, bias = c (ln_fold.cu)
  Gemm& ln(const double* stats, float eps, int width, float* mean, float* rstd) {
    d.ln_stats = stats; d.ln_eps = eps; d.K_ln = width; d.ln_mean_out = mean; d.ln_rstd_out = rstd;
    return *this;
  }
  Gemm& stats_out(double* st) { d.stats_out = st; return *this; }
  int run(cudaStream_t st) { return gemm_dispatch(&d, st); }
};

#define NRV_TRY(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

// bias gradient out[cols] += colsum(x): one kernel whose row slices meet in fp32 vector reds (no partials, no
// finalize launch); unaligned outputs take the two-kernel path
static int bias_colsum(const Dims& d, const void* x, int cols, float* out, void* red, size_t red_bytes, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(out) % 16) != 0) return colsum_rows(x, cols, d.T, cols, d.dtype, 1, 0, out, red, red_bytes, st);
  return colsum_atomic(x, cols, d.T, cols, d.dtype, out, st);
}

// One-shot marker of the next nrv_vit_backward call (nrv_vit_backward_marker): data-parallel training starts a bucket's
// all-reduce at this event instead of at the end of the call, so the collective begins under the LayerNorm backward that
// closes the stage (many small CTAs that share the SMs gracefully) instead of beside the next persistent GEMM.
static thread_local cudaEvent_t g_bwd_marker = nullptr;

static int check_cfg_runtime(const nrv_vit_config* c) {
  NRV_REQUIRE(c->pool == NRV_POOL_MEAN || c->pool == NRV_POOL_CLS, "nrv_vit: bad pool mode");
  NRV_REQUIRE(c->pool != NRV_POOL_CLS || c->cls_token, "nrv_vit: class-token pooling needs cls_token=1");
  NRV_REQUIRE(c->patch_order == NRV_PATCH_P1P2C || c->patch_order == NRV_PATCH_CP1P2, "nrv_vit: bad patch_order");
  NRV_REQUIRE(c->attn_mode == NRV_ATTN_SOFTMAX || c->attn_mode == NRV_ATTN_SINKHORN3, "nrv_vit: bad attn_mode");
  if (c->training && c->p_attn_drop > 0.f && c->attn_mode != NRV_ATTN_SOFTMAX) {
    set_error("nrv_vit: dropout on the attention probabilities (p_attn_drop=%g) is implemented for softmax attention only", (double)c->p_attn_drop);
    return NRV_ENOTIMPL;
  }
  if (c->attn_mode == NRV_ATTN_SINKHORN3) {
    const int N = (c->img_h / c->patch_h) * (c->img_w / c->patch_w) + (c->cls_token ? 1 : 0);
    if (!sinkhorn_supported(N, c->dim_head)) {
      set_error("nrv_vit: robust=True (Sinkhorn attention): the K / V tile of one head (N=%d, dh=%d) exceeds shared memory; no fallback", N, c->dim_head);
      return NRV_ENOTIMPL;
    }
  }
  return NRV_OK;
}

}  // namespace nrv

using namespace nrv;

extern "C" {

size_t nrv_vit_stash_bytes(const nrv_vit_config* cfg) {
  Dims d;
  if (make_dims(cfg, &d)) return 0;
  if (!cfg->training) return 0;
  return plan_stash(d).total;
}

int nrv_vit_stash_tensor(const nrv_vit_config* cfg, int what, int index, size_t* offset, size_t* bytes) {
  Dims d;
  NRV_TRY(make_dims(cfg, &d));
  NRV_REQUIRE(cfg->training, "nrv_vit_stash_tensor: the stash exists for training=1 forwards only");
  NRV_REQUIRE(offset && bytes, "nrv_vit_stash_tensor: null pointer");
  const StashPlan sp = plan_stash(d);
  if (what == NRV_STASH_STREAM) {
    NRV_REQUIRE(index >= 0 && index <= 2 * d.L, "nrv_vit_stash_tensor: stream index must be in [0, 2*depth]");
    *offset = sp.xs0 + sp.xs_stride * (size_t)index;
    *bytes = (size_t)d.T * d.D * d.esz;
  } else if (what == NRV_STASH_QKV) {
    NRV_REQUIRE(index >= 0 && index < d.L, "nrv_vit_stash_tensor: layer index out of range");
    *offset = sp.layer0 + sp.layer_stride * (size_t)index + sp.l.qkv;
    *bytes = (size_t)d.T * 3 * d.I * d.esz;
  } else {
    set_error("nrv_vit_stash_tensor: unknown tensor %d", what);
    return NRV_EINVAL;
  }
  return NRV_OK;
}

size_t nrv_vit_workspace_bytes(const nrv_vit_config* cfg) {
  Dims d;
  if (make_dims(cfg, &d)) return 0;
  return plan_work(d, cfg->training != 0).total;
}

int nrv_vit_forward(const nrv_vit_config* cfg, const nrv_vit_params* P, const void* img, void* feat,
                    void* stash, void* workspace, void* stream) {
  NRV_TRY(require_init());
  Dims d;
  NRV_TRY(make_dims(cfg, &d));
  NRV_TRY(check_cfg_runtime(cfg));
  NRV_REQUIRE(P && img && feat && workspace, "nrv_vit_forward: null pointer");
  NRV_REQUIRE(!cfg->training || stash, "nrv_vit_forward: training=1 needs a stash buffer");
  NRV_REQUIRE(P->w_patch && P->b_patch && P->lnf_g && P->lnf_b && P->layers, "nrv_vit_forward: null parameter");
  NRV_REQUIRE(!cfg->cls_token || P->cls, "nrv_vit_forward: cls_token=1 needs params.cls");
  cudaStream_t st = (cudaStream_t)stream;
  Bufs bf{(uint8_t*)stash, (uint8_t*)workspace, plan_stash(d), plan_work(d, cfg->training != 0), cfg->training != 0};
  const int dt = d.dtype;
  const float scale = 1.0f / sqrtf((float)d.dh);

  // ---- patch embedding (simple_vit.py:126-131,141-143 ; vit.py:323-342,174)
  // The patch matrix is laid out on token rows ([B, N, pld]; class-token rows zero) so that the GEMM runs over
  // all T rows with output row = input row, and backward reuses it as the B operand of dW (training: stash).
  uint8_t* patches = cfg->training ? bf.stash + bf.sp.patches : bf.work + bf.wp.patches;
  const int tok_off = cfg->cls_token ? 1 : 0;
  // bf16 images in the conv_proj order with 16/32/64-wide patches: im2col happens inside the GEMM's TMA loads (no patch
  // matrix, no im2col launch; backward gathers the same way).  Forward and backward take the same decision (same pointer).
  const bool patch_tma = patch_tma_shape_ok(cfg->channels, cfg->img_h, cfg->img_w, cfg->patch_h, cfg->patch_w, cfg->patch_order,
                                            dt, cfg->img_dtype, d.D) && (reinterpret_cast<uintptr_t>(img) % 16) == 0;
  if (patch_tma) {
    NRV_TRY(nrv_patch_embed_fwd(img, d.B, cfg->channels, cfg->img_h, cfg->img_w, cfg->patch_h, cfg->patch_w, P->w_patch, d.pld,
                                P->b_patch, P->pos, d.D, d.N, tok_off, bf.xs(0), d.D, d.D, stream));
  } else {
    NRV_TRY(im2col_rows(img, cfg->img_dtype, d.B, cfg->channels, cfg->img_h, cfg->img_w, cfg->patch_h,
                        cfg->patch_w, cfg->patch_order, patches, dt, d.pld, d.N, tok_off, st));
    Gemm g(d, bf, d.T, d.D, d.pld);
    g.A(patches, d.pld).Bm(P->w_patch, d.pld).out(bf.xs(0), d.D).bias(P->b_patch);
    g.d.pos = P->pos; g.d.ldpos = d.D;
    g.d.pos_rows_in = d.N; g.d.pos_rows_out = d.N; g.d.pos_row_off = 0;
    NRV_TRY(g.run(st));
  }
  if (cfg->cls_token)
    NRV_TRY(nrv_cls_token_fwd(P->cls, P->pos, bf.xs(0), d.B, d.N, d.D, dt, stream));
  const long long TD = d.T * d.D, TM = d.T * d.M;
  if (d.p_emb > 0.f)   // vit.py:174  self.dropout(input + pos_embedding)
    NRV_TRY(nrv_dropout(bf.xs(0), nullptr, bf.xs(0), TD, dt, d.p_emb, d.seed, -1, NRV_DROP_EMB, stream));
  void* branch = d.p_drop > 0.f ? bf.work + bf.wp.dxm : nullptr;   // branch output before its dropout

  // ---- folded LayerNorm: statistics of the embedding output, gamma o W of every layer (one launch each); the
  //      statistics of every later stream state are accumulated by the epilogue of the GEMM that writes it
  const Fold fold{d, bf.work, bf.wp};
  if (d.fold) {
    NRV_REQUIRE(d.L * 2 <= 64, "nrv_vit: the folded LayerNorm supports up to 32 layers per call (depth %d); use ln_mode = NRV_LN_SEPARATE", d.L);
    NRV_CUDA(cudaMemsetAsync(bf.work + bf.wp.stats, 0, bf.wp.stats_bytes, st));
    NRV_TRY(rowstats(bf.xs(0), d.T, d.D, dt, fold.stats(0), st));
    LnFoldJob jobs[64];
    for (int l = 0; l < d.L; ++l) {
      const nrv_vit_layer& W = P->layers[l];
      NRV_REQUIRE(W.w_qkv && W.w_fc1 && W.ln1_g && W.ln1_b && W.ln2_g && W.ln2_b && W.b_fc1, "nrv_vit_forward: null parameter in layer %d", l);
      jobs[2 * l] = LnFoldJob{W.w_qkv, fold.w_qkv(l), W.ln1_g, W.ln1_b, W.b_qkv, fold.c_qkv(l), 3 * d.I};
      jobs[2 * l + 1] = LnFoldJob{W.w_fc1, fold.w_fc1(l), W.ln2_g, W.ln2_b, W.b_fc1, fold.c_fc1(l), d.M};
    }
    // rows of different jobs have different counts but one width: QKV and FC1 both read D columns
    NRV_TRY(ln_fold_weights(jobs, 2 * d.L, d.D, d.D, dt, st));
  }

  // ---- transformer layers
  for (int l = 0; l < d.L; ++l) {
    const nrv_vit_layer& W = P->layers[l];
    NRV_REQUIRE(W.w_qkv && W.w_out && W.w_fc1 && W.w_fc2 && W.ln1_g && W.ln1_b && W.ln2_g && W.ln2_b && W.b_fc1 && W.b_fc2,
                "nrv_vit_forward: null parameter in layer %d", l);
    const StashPlan& sp = bf.sp;
    void* x0 = bf.xs(2 * l);
    void* x1 = bf.xs(2 * l + 1);
    void* x2 = bf.xs(2 * l + 2);
    void* xn1 = bf.layer(l, sp.l.xn1);
    void* qkv = bf.layer(l, sp.l.qkv);
    void* o = bf.layer(l, sp.l.o);
    float* lse = (float*)bf.layer(l, sp.l.lse);
    void* xn2 = bf.layer(l, sp.l.xn2);
    void* u = bf.layer(l, sp.l.u);
    void* h = bf.layer(l, sp.l.h);
    // x = attn(x) + x
    float* mean1 = cfg->training ? (float*)bf.layer(l, sp.l.mean1) : nullptr;
    float* rstd1 = cfg->training ? (float*)bf.layer(l, sp.l.rstd1) : nullptr;
    float* mean2 = cfg->training ? (float*)bf.layer(l, sp.l.mean2) : nullptr;
    float* rstd2 = cfg->training ? (float*)bf.layer(l, sp.l.rstd2) : nullptr;
    if (d.fold) {
      NRV_TRY(Gemm(d, bf, d.T, 3 * d.I, d.D).A(x0, d.D).Bm(fold.w_qkv(l), d.D).out(qkv, 3 * d.I).bias(fold.c_qkv(l))
                  .ln(fold.stats(2 * l), cfg->ln_eps, d.D, mean1, rstd1).run(st));
    } else {
      NRV_TRY(nrv_layernorm_fwd(x0, W.ln1_g, W.ln1_b, cfg->ln_eps, xn1, (float*)bf.layer(l, sp.l.mean1),
                                (float*)bf.layer(l, sp.l.rstd1), d.T, d.D, dt, stream));
      NRV_TRY(Gemm(d, bf, d.T, 3 * d.I, d.D).A(xn1, d.D).Bm(W.w_qkv, d.D).out(qkv, 3 * d.I).bias(W.b_qkv).run(st));
    }
    if (d.p_attn > 0.f) {
      // dropout on the probabilities: the general tcgen05 kernel draws the mask (bf16, dh <= 128, <= 384 tokens); the
      // CUDA-core kernel for the fp32 check mode and the shapes beyond (same mask stream)
      if (cfg->attn_impl != NRV_ATTN_IMPL_SIMT && attn_big_supported(d.N, d.dh, dt))
        NRV_TRY(attn_fwd_big(qkv, o, lse, d.B, d.N, d.H, d.dh, scale, st, d.p_attn, d.seed, l));
      else if (cfg->attn_impl == NRV_ATTN_IMPL_TC) {
        set_error("nrv_vit: tcgen05 attention with dropout does not support N=%d dh=%d dtype=%d", d.N, d.dh, dt);
        return NRV_ENOTIMPL;
      } else
        NRV_TRY(attn_fwd_simt(qkv, o, lse, d.B, d.N, d.H, d.dh, scale, dt, st, d.p_attn, d.seed, l));
    } else
      NRV_TRY(nrv_attn_fwd(qkv, o, lse, d.B, d.N, d.H, d.dh, scale, cfg->attn_mode, dt, cfg->attn_impl,
                           bf.work + bf.wp.attn_ws, bf.wp.attn_ws_bytes, stream));
    if (branch) {   // x1 = dropout(out_proj(o)) + x0   (vit.py:124-126)
      NRV_TRY(Gemm(d, bf, d.T, d.D, d.I).A(o, d.I).Bm(W.w_out, d.I).out(branch, d.D).bias(W.b_out).run(st));
      NRV_TRY(nrv_dropout(branch, x0, x1, TD, dt, d.p_drop, d.seed, l, NRV_DROP_ATTN_OUT, stream));
    } else {
      NRV_TRY(Gemm(d, bf, d.T, d.D, d.I).A(o, d.I).Bm(W.w_out, d.I).out(x1, d.D).bias(W.b_out).residual(x0, d.D)
                  .stats_out(d.fold ? fold.stats(2 * l + 1) : nullptr).run(st));
    }
    // x = ff(x) + x
    const void* fc1_in = xn2;
    const void* fc1_w = W.w_fc1;
    const float* fc1_b = W.b_fc1;
    if (d.fold) {
      fc1_in = x1; fc1_w = fold.w_fc1(l); fc1_b = fold.c_fc1(l);
    } else {
      NRV_TRY(nrv_layernorm_fwd(x1, W.ln2_g, W.ln2_b, cfg->ln_eps, xn2, (float*)bf.layer(l, sp.l.mean2),
                                (float*)bf.layer(l, sp.l.rstd2), d.T, d.D, dt, stream));
    }
    // training keeps gelu'(u) (slot `u` of the stash) next to h = gelu(u): the backward epilogue only multiplies
    {
      Gemm g1(d, bf, d.T, d.M, d.D);
      g1.A(fc1_in, d.D).Bm(fc1_w, d.D).out(h, d.M).bias(fc1_b);
      if (d.fold) g1.ln(fold.stats(2 * l + 1), cfg->ln_eps, d.D, mean2, rstd2);
      if (cfg->training) g1.gelu_grad(u); else g1.gelu(nullptr);
      NRV_TRY(g1.run(st));
    }
    if (branch) {
      // dropout after GELU (vit.py:45): the same mask scales h and the stored gelu'(u), so backward needs no extra pass
      NRV_TRY(nrv_dropout(h, nullptr, h, TM, dt, d.p_drop, d.seed, l, NRV_DROP_FC1, stream));
      NRV_TRY(nrv_dropout(u, nullptr, u, TM, dt, d.p_drop, d.seed, l, NRV_DROP_FC1, stream));
      // x2 = dropout(fc2(h)) + x1   (vit.py:46-47,129-130)
      NRV_TRY(Gemm(d, bf, d.T, d.D, d.M).A(h, d.M).Bm(W.w_fc2, d.M).out(branch, d.D).bias(W.b_fc2).run(st));
      NRV_TRY(nrv_dropout(branch, x1, x2, TD, dt, d.p_drop, d.seed, l, NRV_DROP_FC2, stream));
    } else {
      NRV_TRY(Gemm(d, bf, d.T, d.D, d.M).A(h, d.M).Bm(W.w_fc2, d.M).out(x2, d.D).bias(W.b_fc2).residual(x1, d.D)
                  .stats_out(d.fold && l + 1 < d.L ? fold.stats(2 * l + 2) : nullptr).run(st));
    }
  }

  // ---- pool + final LayerNorm (simple_vit.py:146,136 ; vit.py:175,347)
  void* pooled = bf.tail(0);
  NRV_TRY(nrv_pool_fwd(bf.xs(2 * d.L), pooled, d.B, d.N, d.D, cfg->pool, dt, stream));
  NRV_TRY(nrv_layernorm_fwd(pooled, P->lnf_g, P->lnf_b, cfg->ln_eps, feat, (float*)bf.tail(1), (float*)bf.tail(2),
                            d.B, d.D, dt, stream));
  return NRV_OK;
}

int nrv_vit_backward(const nrv_vit_config* cfg, const nrv_vit_params* P, const nrv_vit_params* G,
                     const void* img, const void* dfeat, void* stash, void* workspace, int stage_hi,
                     int stage_lo, void* stream) {
  NRV_TRY(require_init());
  Dims d;
  NRV_TRY(make_dims(cfg, &d));
  NRV_TRY(check_cfg_runtime(cfg));
  NRV_REQUIRE(cfg->training, "nrv_vit_backward: the forward pass must have run with training=1");
  NRV_REQUIRE(P && G && stash && workspace, "nrv_vit_backward: null pointer");
  NRV_REQUIRE(stage_hi <= d.L && stage_lo >= -1 && stage_lo <= stage_hi,
              "nrv_vit_backward: stages must satisfy -1 <= stage_lo <= stage_hi <= depth (got %d..%d)", stage_lo, stage_hi);
  NRV_REQUIRE(P->layers && G->layers, "nrv_vit_backward: null layer table");
  cudaStream_t st = (cudaStream_t)stream;
  Bufs bf{(uint8_t*)stash, (uint8_t*)workspace, plan_stash(d), plan_work(d, true), true};
  const StashPlan& sp = bf.sp;
  const int dt = d.dtype;
  const float scale = 1.0f / sqrtf((float)d.dh);
  uint8_t* W0 = bf.work;
  void* red = W0 + bf.wp.red;
  const size_t red_bytes = bf.wp.red_bytes;
  void* dxn = W0 + bf.wp.dxn;
  void* dqkv = W0 + bf.wp.dqkv;
  void* dob = W0 + bf.wp.dob;
  void* du = W0 + bf.wp.du;
  // The gradient of the residual stream ping-pongs between dxa and dxb.  Which buffer holds the
  // gradient entering stage s is a pure function of s, so a backward split over several calls
  // stays consistent:  grad wrt xs[2l+2] (input of layer-stage l) lives in dxa, the mid-layer
  // gradient (wrt xs[2l+1]) in dxb, and LN1-bwd writes the next stage's input back to dxa.
  void* dxa = W0 + bf.wp.dxa;
  void* dxb = W0 + bf.wp.dxb;
  // With dropout the gradient of a branch output is the stream gradient times that branch's mask (dxm); the bias
  // gradients of fc2 / out_proj are then column sums of dxm instead of by-products of the LayerNorm backward.
  const bool drop = d.p_drop > 0.f;
  void* dxm = drop ? W0 + bf.wp.dxm : nullptr;
  const long long TD = d.T * d.D;

  for (int s = stage_hi; s >= stage_lo; --s) {
    if (s == d.L) {
      // ---- head side: final LN bwd + pool bwd
      NRV_REQUIRE(dfeat != nullptr, "nrv_vit_backward: stage depth needs dfeat");
      void* dpooled = W0 + bf.wp.dpooled;
      NRV_TRY(nrv_layernorm_bwd(dfeat, bf.tail(0), (const float*)bf.tail(1), (const float*)bf.tail(2), P->lnf_g,
                                nullptr, dpooled, G->lnf_g, G->lnf_b, nullptr, nullptr, nullptr, d.B, d.D, dt, red, red_bytes, stream));
      NRV_TRY(nrv_pool_bwd(dpooled, dxa, d.B, d.N, d.D, cfg->pool, dt, stream));
      // bias gradient of the last layer's fc2 (its output gradient is produced here, not by an LN-bwd)
      if (G->layers[d.L - 1].b_fc2 && !drop)
        NRV_TRY(nrv_colsum(dxa, d.D, d.T, d.D, dt, G->layers[d.L - 1].b_fc2, red, red_bytes, stream));
    } else if (s >= 0) {
      const int l = s;
      const nrv_vit_layer& W = P->layers[l];
      const nrv_vit_layer& g = G->layers[l];
      void* x0 = bf.xs(2 * l);
      void* x1 = bf.xs(2 * l + 1);
      void* xn1 = bf.layer(l, sp.l.xn1);
      void* qkv = bf.layer(l, sp.l.qkv);
      void* o = bf.layer(l, sp.l.o);
      float* lse = (float*)bf.layer(l, sp.l.lse);
      void* xn2 = bf.layer(l, sp.l.xn2);
      void* u = bf.layer(l, sp.l.u);
      void* h = bf.layer(l, sp.l.h);
      // ---- MLP branch.  dxa = grad wrt x2 ; dz = grad wrt the fc2 output
      const void* dz = dxa;
      if (drop) {
        NRV_TRY(nrv_dropout(dxa, nullptr, dxm, TD, dt, d.p_drop, d.seed, l, NRV_DROP_FC2, stream));
        dz = dxm;
        if (g.b_fc2) NRV_TRY(bias_colsum(d, dxm, d.D, g.b_fc2, red, red_bytes, st));
      }
      if (g.w_fc2) NRV_TRY(Gemm(d, bf, d.D, d.M, d.T).A(dz, d.D, NRV_MN_MAJOR).Bm(h, d.M, NRV_MN_MAJOR).out(g.w_fc2, d.M).atomic().run(st));
      // du = (dz W2) o gelu'(u) ; in bf16 the same epilogue also reduces du's columns into the fc1 bias gradient
      const bool fused_b1 = dt == NRV_BF16 && g.b_fc1 != nullptr && (reinterpret_cast<uintptr_t>(g.b_fc1) % 8) == 0;
      NRV_TRY(Gemm(d, bf, d.T, d.M, d.D).A(dz, d.D).Bm(W.w_fc2, d.M, NRV_MN_MAJOR).out(du, d.M).mul(u, d.M)
                  .colsum(fused_b1 ? g.b_fc1 : nullptr).run(st));
      // folded LayerNorm: the forward pass kept no normalised rows; the LayerNorm backward below writes them (xnb) for the
      // weight-gradient GEMM, which therefore runs after it
      void* xnb = d.fold ? W0 + bf.wp.xnb : nullptr;
      if (g.w_fc1 && !d.fold) NRV_TRY(Gemm(d, bf, d.M, d.D, d.T).A(du, d.M, NRV_MN_MAJOR).Bm(xn2, d.D, NRV_MN_MAJOR).out(g.w_fc1, d.D).atomic().run(st));
      if (g.b_fc1 && !fused_b1) NRV_TRY(bias_colsum(d, du, d.M, g.b_fc1, red, red_bytes, st));
      NRV_TRY(Gemm(d, bf, d.T, d.D, d.M).A(du, d.M).Bm(W.w_fc1, d.D, NRV_MN_MAJOR).out(dxn, d.D).run(st));
      // dxb = LN2'(dxn) + dxa ; colsum(dxb) = grad of out_proj.bias
      NRV_TRY(nrv_layernorm_bwd(dxn, x1, (const float*)bf.layer(l, sp.l.mean2), (const float*)bf.layer(l, sp.l.rstd2),
                                W.ln2_g, dxa, dxb, g.ln2_g, g.ln2_b, drop ? nullptr : g.b_out,
                                d.fold && g.w_fc1 ? W.ln2_b : nullptr, d.fold && g.w_fc1 ? xnb : nullptr,
                                d.T, d.D, dt, red, red_bytes, stream));
      if (g.w_fc1 && d.fold) NRV_TRY(Gemm(d, bf, d.M, d.D, d.T).A(du, d.M, NRV_MN_MAJOR).Bm(xnb, d.D, NRV_MN_MAJOR).out(g.w_fc1, d.D).atomic().run(st));
      // ---- attention branch.  dxb = grad wrt x1 ; da = grad wrt the out_proj output
      const void* da = dxb;
      if (drop) {
        NRV_TRY(nrv_dropout(dxb, nullptr, dxm, TD, dt, d.p_drop, d.seed, l, NRV_DROP_ATTN_OUT, stream));
        da = dxm;
        if (g.b_out) NRV_TRY(bias_colsum(d, dxm, d.D, g.b_out, red, red_bytes, st));
      }
      if (g.w_out) NRV_TRY(Gemm(d, bf, d.D, d.I, d.T).A(da, d.D, NRV_MN_MAJOR).Bm(o, d.I, NRV_MN_MAJOR).out(g.w_out, d.I).atomic().run(st));
      NRV_TRY(Gemm(d, bf, d.T, d.I, d.D).A(da, d.D).Bm(W.w_out, d.I, NRV_MN_MAJOR).out(dob, d.I).run(st));
      // the fused tcgen05 backward also reduces dqkv's columns into the in_proj bias gradient from its epilogue
      const bool fused_bqkv = d.p_attn == 0.f && g.b_qkv != nullptr && cfg->attn_mode == NRV_ATTN_SOFTMAX &&
                              cfg->attn_impl != NRV_ATTN_IMPL_SIMT && attn_bwd2_supported(d.N, d.dh, dt) &&
                              (reinterpret_cast<uintptr_t>(g.b_qkv) % 8) == 0;
      if (d.p_attn > 0.f) {
        if (cfg->attn_impl != NRV_ATTN_IMPL_SIMT && attn_bwd_big_supported(d.N, d.dh, dt))
          NRV_TRY(attn_bwd_big(qkv, o, dob, lse, dqkv, (float*)(W0 + bf.wp.attn_ws), bf.wp.attn_ws_bytes, d.B, d.N, d.H, d.dh, scale,
                               st, d.p_attn, d.seed, l));
        else if (cfg->attn_impl == NRV_ATTN_IMPL_TC) {
          set_error("nrv_vit: tcgen05 attention backward with dropout does not support N=%d dh=%d dtype=%d", d.N, d.dh, dt);
          return NRV_ENOTIMPL;
        } else
          NRV_TRY(attn_bwd_simt(qkv, o, dob, lse, dqkv, d.B, d.N, d.H, d.dh, scale, dt, st, d.p_attn, d.seed, l));
      } else if (fused_bqkv)
        NRV_TRY(attn_bwd_tc2(qkv, o, dob, lse, dqkv, d.B, d.N, d.H, d.dh, scale, st, g.b_qkv));
      else
        NRV_TRY(nrv_attn_bwd(qkv, o, dob, lse, dqkv, d.B, d.N, d.H, d.dh, scale, cfg->attn_mode, dt, cfg->attn_impl,
                             W0 + bf.wp.attn_ws, bf.wp.attn_ws_bytes, stream));
      if (g.w_qkv && !d.fold) NRV_TRY(Gemm(d, bf, 3 * d.I, d.D, d.T).A(dqkv, 3 * d.I, NRV_MN_MAJOR).Bm(xn1, d.D, NRV_MN_MAJOR).out(g.w_qkv, d.D).atomic().run(st));
      if (g.b_qkv && !fused_bqkv) NRV_TRY(bias_colsum(d, dqkv, 3 * d.I, g.b_qkv, red, red_bytes, st));
      NRV_TRY(Gemm(d, bf, d.T, d.D, 3 * d.I).A(dqkv, 3 * d.I).Bm(W.w_qkv, d.D, NRV_MN_MAJOR).out(dxn, d.D).run(st));
      // every gradient of stages >= stage_lo except this layer's ln1 gamma / beta (and the previous layer's fc2 bias, which
      // belongs to the next bucket anyway) is final here
      if (g_bwd_marker != nullptr && s == stage_lo && !d.fold) {
        NRV_CUDA(cudaEventRecord(g_bwd_marker, st));
        g_bwd_marker = nullptr;
      }
      // dxa = LN1'(dxn) + dxb ; colsum(dxa) = grad of the previous layer's fc2 bias
      float* prev_b_fc2 = (l > 0 && !drop) ? G->layers[l - 1].b_fc2 : nullptr;
      NRV_TRY(nrv_layernorm_bwd(dxn, x0, (const float*)bf.layer(l, sp.l.mean1), (const float*)bf.layer(l, sp.l.rstd1),
                                W.ln1_g, dxb, dxa, g.ln1_g, g.ln1_b, prev_b_fc2,
                                d.fold && g.w_qkv ? W.ln1_b : nullptr, d.fold && g.w_qkv ? xnb : nullptr,
                                d.T, d.D, dt, red, red_bytes, stream));
      if (g.w_qkv && d.fold) NRV_TRY(Gemm(d, bf, 3 * d.I, d.D, d.T).A(dqkv, 3 * d.I, NRV_MN_MAJOR).Bm(xnb, d.D, NRV_MN_MAJOR).out(g.w_qkv, d.D).atomic().run(st));
    } else {
      // ---- embedding: dxa = grad wrt xs[0]  (autograd of simple_vit.py:126-143 / vit.py:323-342,174)
      if (d.p_emb > 0.f)
        NRV_TRY(nrv_dropout(dxa, nullptr, dxa, TD, dt, d.p_emb, d.seed, -1, NRV_DROP_EMB, stream));
      if (G->pos || G->cls)
        NRV_TRY(nrv_posemb_bwd(dxa, d.B, d.N, d.D, dt, G->pos, cfg->cls_token ? G->cls : nullptr, stream));
      const int off = cfg->cls_token ? 1 : 0;
      if (G->b_patch)
        NRV_TRY(colsum_rows(dxa, d.D, d.T, d.D, dt, d.N, off, G->b_patch, red, red_bytes, st));
      if (G->w_patch) {
        const bool patch_tma = patch_tma_shape_ok(cfg->channels, cfg->img_h, cfg->img_w, cfg->patch_h, cfg->patch_w, cfg->patch_order,
                                                  dt, cfg->img_dtype, d.D) && (reinterpret_cast<uintptr_t>(img) % 16) == 0;
        if (patch_tma) {
          // the forward pass kept no patch matrix: dW = dx^T * patches gathers both operands by TMA (K = patch rows)
          NRV_REQUIRE(img != nullptr, "nrv_vit_backward: the embedding stage needs the image");
          NRV_TRY(nrv_patch_embed_bwd_weight(img, d.B, cfg->channels, cfg->img_h, cfg->img_w, cfg->patch_h, cfg->patch_w, dxa, d.D,
                                             d.N, off, (float*)G->w_patch, d.pld, d.D, stream));
        } else {
          // dW = dx^T * patches is one GEMM over K = T (class-token rows of the stashed patch matrix are zero)
          uint8_t* patches = bf.stash + sp.patches;
          NRV_TRY(Gemm(d, bf, d.D, d.pld, d.T).A(dxa, d.D, NRV_MN_MAJOR).Bm(patches, d.pld, NRV_MN_MAJOR).out(G->w_patch, d.pld).atomic().run(st));
        }
      }
    }
  }
  if (g_bwd_marker != nullptr) {   // folded LayerNorm (its dW GEMM follows the LayerNorm backward) / embedding stage: at the end
    NRV_CUDA(cudaEventRecord(g_bwd_marker, st));
    g_bwd_marker = nullptr;
  }
  return NRV_OK;
}

int nrv_vit_backward_marker(void* cuda_event) {
  g_bwd_marker = (cudaEvent_t)cuda_event;
  return NRV_OK;
}

}  // extern "C"
