// Internal declarations shared by the translation units of libnrvit.
#pragma once
#include "../../include/nrvit.h"
#include <cuda_runtime.h>

namespace nrv {
int gemm_dispatch(const nrv_gemm_desc* d, cudaStream_t stream);
// Patch embedding with the im2col fused into the GEMM's operand loads (north_star "patch-embedding kernel fusing im2col with
// the projection GEMM via TMA"; reference ops vit.py:323-331 conv_proj + reshape / permute, :237-242).  The image
// [B, C, H, W] bf16 is addressed as a 5-D tensor (p2, px, image, y, channel); one TMA box = 64 consecutive elements of the
// (c p1 p2) patch vector (64 / pw rows of one patch) for npx neighbouring patches of nb images.  A swizzled TMA box puts every
// inner row (the pw pixels of one patch row) on a line of its own (measured: tools/ubench/tma_box.cu), so the tile is
// 64 / pw sub-tiles [128 patches x pw*2 bytes], one per patch row, in the SWIZZLE_32B / 64B / 128B operand layout for
// pw = 16 / 32 / 64, and the MMA K steps walk the sub-tiles.
struct PatchView {
  const void* img;
  int B, C, H, W, ph, pw, gh, gw;
  int npx, nb;                       // tile shape: npx = largest power of two dividing gw (<= 64), nb = 128 / npx
  int tokens_per_img, tok_off;       // token rows of the output / gradient: row = b * tokens_per_img + tok_off + py * gw + px
  int role;                          // 1: forward (A = image), 2: weight gradient (A = gradient rows gathered alike, B = image)
  int layout;                        // UMMA layout type of the image tiles: 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B for pw = 16 / 32 / 64
};
// shape conditions (pure function of the configuration); the image pointer must also be 16-byte aligned
bool patch_tma_shape_ok(int C, int H, int W, int ph, int pw, int order, int dtype, int img_dtype, int D);
int patch_view_init(PatchView* pv, const void* img, int B, int C, int H, int W, int ph, int pw, int tokens_per_img, int tok_off, int role);
int gemm_dispatch_patch(const nrv_gemm_desc* d, const PatchView* pv, cudaStream_t stream);
size_t gemm_workspace_bytes(int M, int N, int K, int dtype);
int colsum_rows(const void* x, long long ldx, long long rows, int cols, int dtype, int period, int skip,
                float* out, void* workspace, size_t workspace_bytes, cudaStream_t st);
int colsum_atomic(const void* x, long long ldx, long long rows, int cols, int dtype, float* out, cudaStream_t st);
int im2col_rows(const void* img, int img_dtype, int B, int C, int H, int W, int ph, int pw, int order,
                void* patches, int out_dtype, long long ld, int rows_out, int row_off, cudaStream_t st);
// p_drop > 0: dropout on the attention probabilities (mask keyed by seed / layer, site NRV_DROP_ATTN_PROB)
int attn_fwd_simt(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale,
                  int dtype, cudaStream_t st, float p_drop = 0.f, unsigned long long seed = 0, int layer = 0);
int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  int B, int N, int H, int dh, float scale, int dtype, cudaStream_t st, float p_drop = 0.f,
                  unsigned long long seed = 0, int layer = 0);
bool attn_tc_supported(int N, int dh, int dtype);
int attn_fwd_tc2(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale, cudaStream_t st);
bool attn_big_supported(int N, int dh, int dtype);
// p_drop > 0 (general kernels only): dropout on the probabilities, same mask stream as the CUDA-core kernels
int attn_fwd_big(const void* qkv, void* out, float* lse, int B, int N, int H, int dh, float scale, cudaStream_t st,
                 float p_drop = 0.f, unsigned long long seed = 0, int layer = 0);
bool attn_bwd2_supported(int N, int dh, int dtype);
// dbias (optional, fp32 [3*H*dh]): += column sums of the stored dqkv, i.e. the in_proj bias gradient, from the epilogue
int attn_bwd_tc2(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                 int B, int N, int H, int dh, float scale, cudaStream_t st, float* dbias = nullptr);
// general tcgen05 backward (attention_bwd_big.cu): dh 16..80, up to 1024 tokens; dQ summed over key tiles in a per-CTA fp32 scratch
bool attn_bwd_big_supported(int N, int dh, int dtype);
size_t attn_bwd_big_scratch_bytes(int B, int N, int H, int dh);
int attn_bwd_big(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* scratch,
                 size_t scratch_bytes, int B, int N, int H, int dh, float scale, cudaStream_t st, float p_drop = 0.f,
                 unsigned long long seed = 0, int layer = 0);
void gemm_timing_enable(int on);
int gemm_timing_detail(long long* out, int max_records);
int gemm_timing_read(double* ms, double* flops, long long* launches);
void attn_tc_set_debug(long long* buf);
long long* attn_tc_get_debug();
bool sinkhorn_supported(int N, int dh);
size_t sinkhorn_fwd_scratch_bytes(int B, int N, int H, int dh);
size_t sinkhorn_bwd_scratch_bytes(int B, int N, int H, int dh);
int sinkhorn_fwd(const void* qkv, void* out, float* stats, void* scratch, size_t scratch_bytes, int B, int N, int H, int dh,
                 float scale, int dtype, cudaStream_t st);
bool sinkhorn_tc_supported(int N, int dh, int dtype);
int sinkhorn_fwd_tc(const void* qkv, void* out, float* stats, int B, int N, int H, int dh, float scale, cudaStream_t st);
int sinkhorn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* stats, void* dqkv, int B, int N, int H, int dh,
                    float scale, cudaStream_t st);
int attn_probs(const void* qkv, float* probs, float* stats, void* scratch, size_t scratch_bytes, int B, int N, int H, int dh,
               float scale, int sinkhorn, int dtype, cudaStream_t st);
int sinkhorn_bwd(const void* qkv, const void* dout, const float* stats, void* dqkv, float* scratch, int B, int N,
                 int H, int dh, float scale, int dtype, cudaStream_t st);
// LayerNorm folded into the projection GEMMs (ln_fold.cu)
struct LnFoldJob {
  const void* W; void* Wf; const float* gamma; const float* beta; const float* bias; float* c; int rows;
};
int rowstats(const void* x, long long rows, int dim, int dtype, double* stats, cudaStream_t st);
int ln_fold_weights(const LnFoldJob* jobs, int njobs, int dim, int ldw, int dtype, cudaStream_t st);
bool initialised();
int require_init();
}  // namespace nrv
