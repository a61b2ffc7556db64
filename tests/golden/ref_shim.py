"""Import the reference's Python y-path UNMODIFIED from /root/reference, with stub packages for the
third-party modules that are not installed here.  Used ONLY by the golden
vector generator (tests/golden/make_golden.py) and by tests that are skipped when /root/reference is
absent (it does not exist on the GPU box).  Test scaffolding, not product code."""
import os
import sys
import types

import torch
import torch.nn as nn
from importlib.machinery import ModuleSpec

REF_ROOT = os.environ.get("BASIC_REF_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Auto(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)})
        setattr(self, name, cls)
        return cls


class _Finder:
    roots = ("pytorch_lightning", "compressai", "pytorch_msssim", "entmax", "thop", "ptflops",
             "adabelief_pytorch", "skimage", "zstandard", "brotli", "autograd", "craystack", "survae", "oss2",
             "torchvision")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.roots:
            return ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _Auto(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


_loaded = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "cbench")) and \
        bool([f for f in os.listdir(os.path.join(_REPO, "oracle", "_ref"))
              if f.startswith("ans.")]) if os.path.isdir(os.path.join(_REPO, "oracle", "_ref")) else False


def load():
    """Returns (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder, TopoGroupDynamicMaskConv2dContextModel)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    try:
        import torchvision  # noqa: F401  (real one if present)
        _Finder.roots = tuple(r for r in _Finder.roots if r != "torchvision")
    except Exception:
        pass
    sys.path.insert(0, REF_ROOT)
    sys.meta_path.insert(0, _Finder())
    import compressai.ops.bound_ops as bo

    class LowerBound(nn.Module):
        def __init__(self, b):
            super().__init__()
            self.register_buffer("bound", torch.Tensor([float(b)]))

        def forward(self, x):
            return torch.max(x, self.bound)

    bo.LowerBound = LowerBound
    bo.LowerBoundFunction = type("LowerBoundFunction", (), {"apply": staticmethod(lambda x, b: torch.max(x, b))})
    import cbench
    sys.path.insert(0, _REPO)
    from oracle import ref_loader          # one pybind module instance per process (types register once)
    for ext in ("ans", "rans"):
        mod = ref_loader.load(ext)
        if mod is not None:
            sys.modules["cbench." + ext] = mod
            setattr(cbench, ext, mod)
    from cbench.modules.prior_model.prior_coder.pgm_coder import GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder
    from cbench.nn.layers.masked_conv import TopoGroupDynamicMaskConv2dContextModel
    _loaded = (GaussianChannelGroupMaskConv2DTopoGroupPGMPriorCoder, TopoGroupDynamicMaskConv2dContextModel)
    return _loaded
