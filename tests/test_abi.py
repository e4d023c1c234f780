"""The C-ABI library loads and exports exactly what include/nrvit.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "nrvit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrv_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_hot_path_entries():
    names = header_functions()
    for must in ("nrv_init", "nrv_gemm", "nrv_layernorm_fwd", "nrv_layernorm_bwd", "nrv_attn_fwd", "nrv_attn_bwd",
                 "nrv_softmax_ce", "nrv_adamw", "nrv_vit_forward", "nrv_vit_backward", "nrv_im2col"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in header_functions():
        assert hasattr(lib, name), "libnrvit.so does not export %s" % name


def test_ctypes_table_matches_header(lib_built):
    from vit_pytorch_robust import _abi
    assert sorted(_abi.SIGNATURES) == header_functions()
    lib = _abi.load()
    assert lib.nrv_abi_version() == 7


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "nrvit.h")).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "at::" not in src and "#include <torch" not in src


def test_compute_entries_fail_loudly_without_device(lib_built):
    """No CPU fallback: without nrv_init on an sm_100 device every compute entry returns an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from vit_pytorch_robust import _abi
    lib = _abi.load()
    assert lib.nrv_layernorm_fwd(None, None, None, 1e-5, None, None, None, 1, 8, 0, None) == -4  # NRV_ENOTINIT
    assert b"no CPU fallback" in lib.nrv_last_error()
    with pytest.raises(_abi.NrvError):
        _abi.init()
    import vit_pytorch_robust as v
    m = v.SimpleViT(image_size=32, patch_size=8, num_classes=10, dim=64, depth=1, heads=2, mlp_dim=64)
    with pytest.raises(_abi.NrvError):
        m(torch.randn(1, 3, 32, 32))


def test_kernels_are_blackwell_native(lib_built):
    """SASS must contain tcgen05 MMA (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG)."""
    out = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out or "UTCMMA" in out or re.search(r"UTC\w*MMA", out)
    assert "LDTM" in out
    assert "UTMALDG" in out
