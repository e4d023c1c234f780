"""ctypes binding of libnrvit.so (include/nrvit.h) + torch-tensor marshalling.

There is deliberately no fallback: if the shared library is missing, or the process has no sm_100
GPU, every compute entry raises.  PyTorch is used only as the owner of device memory and streams.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "..", "lib", "libnrvit.so")

NRV_BF16, NRV_F32 = 0, 1
NRV_K_MAJOR, NRV_MN_MAJOR = 0, 1
EPI_STORE, EPI_GELU, EPI_DGELU, EPI_ATOMIC_F32, EPI_GELU_GRAD, EPI_MUL = 0, 1, 2, 3, 4, 5
ATTN_SOFTMAX, ATTN_SINKHORN3 = 0, 1
ATTN_IMPL_AUTO, ATTN_IMPL_SIMT, ATTN_IMPL_TC = 0, 1, 2
POOL_MEAN, POOL_CLS = 0, 1
DROP_ATTN_OUT, DROP_FC1, DROP_FC2, DROP_EMB, DROP_ATTN_PROB = 0, 1, 2, 3, 4
PATCH_P1P2C, PATCH_CP1P2 = 0, 1

_vp, _ll, _i, _f = C.c_void_p, C.c_longlong, C.c_int, C.c_float


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", _i), ("N", _i), ("K", _i), ("dtype", _i), ("out_dtype", _i),
        ("a", _vp), ("lda", _ll), ("a_layout", _i),
        ("b", _vp), ("ldb", _ll), ("b_layout", _i),
        ("epi", _i), ("alpha", _f),
        ("out", _vp), ("ldo", _ll), ("out2", _vp),
        ("bias", _vp),
        ("residual", _vp), ("ldr", _ll),
        ("aux", _vp), ("ldaux", _ll),
        ("pos", _vp), ("ldpos", _ll), ("pos_rows_in", _i), ("pos_rows_out", _i), ("pos_row_off", _i),
        ("splits", _i), ("force_bn128", _i), ("force_single_cta", _i),
        ("workspace", _vp), ("workspace_bytes", C.c_size_t),
        ("colsum", _vp),
        ("tile_mode", _i),
        ("ln_stats", _vp), ("ln_eps", _f), ("K_ln", _i),
        ("ln_mean_out", _vp), ("ln_rstd_out", _vp),
        ("stats_out", _vp),
    ]


class VitConfig(C.Structure):
    _fields_ = [
        ("batch", _i), ("channels", _i), ("img_h", _i), ("img_w", _i), ("patch_h", _i), ("patch_w", _i),
        ("dim", _i), ("depth", _i), ("heads", _i), ("dim_head", _i), ("mlp_dim", _i),
        ("cls_token", _i), ("pool", _i), ("patch_order", _i), ("qkv_bias", _i), ("ln_eps", _f),
        ("attn_mode", _i), ("attn_impl", _i), ("img_dtype", _i), ("dtype", _i), ("training", _i),
        ("p_drop", _f), ("p_emb_drop", _f), ("p_attn_drop", _f), ("drop_seed", C.c_ulonglong),
        ("ln_mode", _i),
    ]


LAYER_FIELDS = ("w_qkv", "w_out", "w_fc1", "w_fc2",
                "ln1_g", "ln1_b", "b_qkv", "b_out", "ln2_g", "ln2_b", "b_fc1", "b_fc2")


class VitLayer(C.Structure):
    _fields_ = [(n, _vp) for n in LAYER_FIELDS]


class VitParams(C.Structure):
    _fields_ = [("w_patch", _vp), ("b_patch", _vp), ("pos", _vp), ("cls", _vp),
                ("lnf_g", _vp), ("lnf_b", _vp), ("layers", C.POINTER(VitLayer))]


_lib = None
_lock = threading.Lock()
_inited = set()
_sz = C.c_size_t
_cfgp, _parp = C.POINTER(VitConfig), C.POINTER(VitParams)

# name -> (restype, argtypes): one row per declaration in include/nrvit.h (tests check both ways)
SIGNATURES = {
    "nrv_abi_version": (_i, []),
    "nrv_init": (_i, [_i]),
    "nrv_last_error": (C.c_char_p, []),
    "nrv_num_sms": (_i, []),
    "nrv_launch_count": (_ll, []),
    "nrv_gemm": (_i, [C.POINTER(GemmDesc), _vp]),
    "nrv_gemm_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "nrv_gemm_timing": (_i, [_i]),
    "nrv_gemm_timing_detail": (_i, [C.POINTER(_ll), _i]),
    "nrv_gemm_timing_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
    "nrv_add_gaussian_noise": (_i, [_vp, _vp, _ll, _i, _f, C.c_ulonglong, _vp]),
    "nrv_dropout": (_i, [_vp, _vp, _vp, _ll, _i, _f, C.c_ulonglong, _i, _i, _vp]),
    "nrv_layernorm_fwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "nrv_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp, _sz, _vp]),
    "nrv_rowstats": (_i, [_vp, _ll, _i, _i, _vp, _vp]),
    "nrv_ln_fold_weights": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _ll, _i, _vp]),
    "nrv_layernorm_bwd_workspace": (_sz, [_ll, _i]),
    "nrv_colsum": (_i, [_vp, _ll, _ll, _i, _i, _vp, _vp, _sz, _vp]),
    "nrv_colsum_workspace": (_sz, [_ll, _i]),
    "nrv_im2col": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _ll, _vp]),
    "nrv_patch_embed_supported": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "nrv_patch_embed_fwd": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _vp, _ll, _i, _i, _vp, _ll, _i, _vp]),
    "nrv_patch_embed_bwd_weight": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _i, _i, _vp, _ll, _i, _vp]),
    "nrv_cls_token_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "nrv_posemb_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "nrv_posemb_sincos_2d": (_i, [_vp, _i, _i, _i, _f, _vp]),
    "nrv_attn_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _sz, _vp]),
    "nrv_attn_probs": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp, _sz, _vp]),
    "nrv_attn_fwd_workspace": (_sz, [_i, _i, _i, _i, _i]),
    "nrv_attn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _i, _vp, _sz, _vp]),
    "nrv_attn_bwd_workspace": (_sz, [_i, _i, _i, _i]),
    "nrv_vit_stash_tensor": (_i, [C.POINTER(VitConfig), _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "nrv_attn_debug_timestamps": (_i, [_vp]),
    "nrv_attn_stats_elems": (_sz, [_i, _i, _i, _i]),
    "nrv_pool_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nrv_pool_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "nrv_softmax_ce": (_i, [_vp, _ll, _vp, _f, _vp, _vp, _i, _ll, _f, _i, _i, _vp]),
    "nrv_adamw": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp, _vp]),
    "nrv_cast_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "nrv_cast_f32": (_i, [_vp, _vp, _ll, _vp]),
    "nrv_sumsq": (_i, [_vp, _ll, _vp, _vp]),
    "nrv_clip_coef": (_i, [_vp, _f, _f, _vp, _vp]),
    "nrv_comm_unique_id_bytes": (_i, []),
    "nrv_comm_get_unique_id": (_i, [_vp, _i]),
    "nrv_comm_init": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "nrv_comm_register": (_i, [_vp, _vp, _sz, C.POINTER(_vp)]),
    "nrv_comm_deregister": (_i, [_vp, _vp]),
    "nrv_comm_allreduce_bucket": (_i, [_vp, _vp, _ll, _i, _vp]),
    "nrv_comm_nccl_version": (_i, []),
    "nrv_comm_destroy": (_i, [_vp]),
    "nrv_vit_stash_bytes": (_sz, [_cfgp]),
    "nrv_vit_workspace_bytes": (_sz, [_cfgp]),
    "nrv_vit_forward": (_i, [_cfgp, _parp, _vp, _vp, _vp, _vp, _vp]),
    "nrv_vit_backward": (_i, [_cfgp, _parp, _parp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "nrv_vit_backward_marker": (_i, [_vp]),
}


STASH_STREAM, STASH_QKV = 0, 1
LN_FOLDED, LN_SEPARATE = 0, 1


class NrvError(RuntimeError):
    pass


def lib_path():
    return os.path.abspath(_LIB_PATH)


def load():
    """dlopen libnrvit.so and declare every entry point (no device needed)."""
    global _lib
    with _lock:
        if _lib is None:
            path = lib_path()
            if not os.path.exists(path):
                raise NrvError(
                    "libnrvit.so not found at %s: run `python noise-robust-vit_b200/build.py` "
                    "(there is no CPU / PyTorch fallback for the hot path)" % path)
            lib = C.CDLL(path)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().nrv_last_error()
        raise NrvError("%s failed (code %d): %s" % (what or "libnrvit call", rc, (msg or b"").decode()))


def init(device=None):
    """Bind the library to a CUDA device (must be sm_100); raises if there is none."""
    lib = load()
    if not torch.cuda.is_available():
        raise NrvError("libnrvit needs an sm_100 (B200) GPU: torch.cuda.is_available() is False and "
                       "there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited:
        torch.cuda.init()
        with torch.cuda.device(idx):
            torch.zeros(1, device=device)  # make sure the primary context exists
            check(lib.nrv_init(idx), "nrv_init")
        _inited.add(idx)
    return lib


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _dt(t):
    if t.dtype == torch.bfloat16:
        return NRV_BF16
    if t.dtype == torch.float32:
        return NRV_F32
    raise NrvError("unsupported dtype %s (bf16 or fp32 only)" % t.dtype)


def _req(t, name, dtype=None):
    if t is None:
        return
    if not t.is_cuda:
        raise NrvError("%s must be a CUDA tensor (no CPU fallback)" % name)
    if dtype is not None and t.dtype != dtype:
        raise NrvError("%s must be %s, got %s" % (name, dtype, t.dtype))


def gemm(a, b, out, *, a_layout=NRV_K_MAJOR, b_layout=NRV_K_MAJOR, epi=EPI_STORE, alpha=1.0,
         bias=None, residual=None, out2=None, aux=None, pos=None, pos_rows_in=0, pos_rows_out=0,
         pos_row_off=0, splits=0, force_bn128=0, force_single_cta=0, M=None, N=None, K=None, stream=None,
         colsum=None, tile_mode=0, ln_stats=None, ln_eps=1e-5, K_ln=0, ln_mean_out=None,
         ln_rstd_out=None, stats_out=None):
    """out[M,N] = epilogue(alpha * A * B^T).  A: [M,K] (K-major) or [K,M] (MN-major); B likewise."""
    lib = init(a.device)
    for t, n in ((a, "a"), (b, "b"), (out, "out"), (bias, "bias"), (residual, "residual"), (out2, "out2"),
                 (aux, "aux"), (pos, "pos")):
        _req(t, n)
    if a.dtype != b.dtype:
        raise NrvError("gemm operands must share a dtype")
    for t, n in ((a, "a"), (b, "b"), (out, "out")):
        if t.dim() != 2 or t.stride(1) != 1:
            raise NrvError("%s must be 2-D with unit inner stride" % n)
    am, ak = (a.shape if a_layout == NRV_K_MAJOR else (a.shape[1], a.shape[0]))
    bn, bk = (b.shape if b_layout == NRV_K_MAJOR else (b.shape[1], b.shape[0]))
    if K is None and bk != ak:
        raise NrvError("gemm: K mismatch %d vs %d" % (ak, bk))
    M = am if M is None else M
    N = bn if N is None else N
    K = ak if K is None else K
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.dtype = _dt(a)
    d.out_dtype = _dt(out)
    d.a, d.lda, d.a_layout = a.data_ptr(), a.stride(0), a_layout
    d.b, d.ldb, d.b_layout = b.data_ptr(), b.stride(0), b_layout
    d.epi, d.alpha = epi, alpha
    d.out, d.ldo = out.data_ptr(), out.stride(0)
    d.out2 = out2.data_ptr() if out2 is not None else None
    d.bias = bias.data_ptr() if bias is not None else None
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    if aux is not None:
        d.aux, d.ldaux = aux.data_ptr(), aux.stride(0)
    if pos is not None:
        d.pos, d.ldpos = pos.data_ptr(), pos.stride(0)
    d.pos_rows_in, d.pos_rows_out, d.pos_row_off = pos_rows_in, pos_rows_out, pos_row_off
    d.splits, d.force_bn128, d.force_single_cta = splits, force_bn128, force_single_cta
    d.tile_mode = tile_mode
    for t, n in ((ln_mean_out, "ln_mean_out"), (ln_rstd_out, "ln_rstd_out")):
        _req(t, n, torch.float32)
    for t, n in ((ln_stats, "ln_stats"), (stats_out, "stats_out")):
        _req(t, n, torch.float64)
    if ln_stats is not None:
        d.ln_stats, d.ln_eps, d.K_ln = ln_stats.data_ptr(), ln_eps, K_ln or K
        if ln_mean_out is not None:
            d.ln_mean_out, d.ln_rstd_out = ln_mean_out.data_ptr(), ln_rstd_out.data_ptr()
    if stats_out is not None:
        d.stats_out = stats_out.data_ptr()
    if colsum is not None:
        _req(colsum, "colsum")
        d.colsum = colsum.data_ptr()
    ws = None
    if a.dtype == torch.float32:
        nbytes = lib.nrv_gemm_workspace_bytes(M, N, K, NRV_F32)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
        d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    check(lib.nrv_gemm(C.byref(d), stream_ptr(stream)), "nrv_gemm")
    return out
