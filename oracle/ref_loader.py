"""ORACLE support — test infrastructure only.  Imports the UNMODIFIED reference modules.

The reference package cannot be imported normally (`vit_pytorch_robust/__init__.py:7` imports a
`.datasets` module that does not exist), so a stub package whose __path__ points at the reference
directory is registered under a private name and the sub-modules are imported through it.
Search order: $NRV_REFERENCE, /root/reference (build container), baseline/_ref (travels to the GPU
box if someone installed it).  Returns None when no reference tree is available — callers skip.
"""
import importlib
import os
import sys
import types

_PKG = "_nrv_reference_pkg"


def find_reference_dir():
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("NRV_REFERENCE"), "/root/reference",
             os.path.join(here, "..", "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "vit_pytorch_robust")):
            return os.path.join(c, "vit_pytorch_robust")
    return None


def load_reference():
    """Returns a namespace with .simple_vit, .vit, .utils of the reference, or None."""
    d = find_reference_dir()
    if d is None:
        return None
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [d]
        sys.modules[_PKG] = pkg
    ns = types.SimpleNamespace()
    ns.simple_vit = importlib.import_module(_PKG + ".simple_vit")
    ns.utils = importlib.import_module(_PKG + ".utils")
    try:
        ns.vit = importlib.import_module(_PKG + ".vit")
    except Exception:  # torchvision API drift
        ns.vit = None
    # the one runnable in-tree class with the README `ViT` structure (vit_with_patch_dropout.py:101-152); with
    # patch_dropout = 0 it IS the README model except that the class-token row gets no positional embedding
    ns.vit_with_patch_dropout = importlib.import_module(_PKG + ".vit_with_patch_dropout")
    return ns


def torchvision_twin(**kw):
    """The class vit.py:178-351 was copied from; forward oracle for the VisionTransformer path."""
    from torchvision.models.vision_transformer import VisionTransformer
    return VisionTransformer(**kw)
