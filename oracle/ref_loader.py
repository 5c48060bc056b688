"""Loader for the *executed* reference (test infrastructure, NOT product code).

Finds the reference's files either in the read-only checkout (/root/reference, build container only) or in the
git-ignored copy oracle/_ref/ that oracle/fetch_ref.py makes at build() time and that travels to the GPU box with
the working tree.  The warper source gets the three token substitutions documented in SURVEY.md Appendix B before
it is exec()'d.  No arithmetic is changed:

  'cuda:0'                  -> target device string          (ref :7)
  torch.cuda.FloatTensor    -> torch.FloatTensor (CPU only)  (ref :45,:118-122,...)
  240*320 / 240 * 320       -> self.W*self.H                 (ref :121-122)

On a CUDA device the file therefore runs exactly as shipped, legacy torch.cuda.FloatTensor constructors included
(those allocate on the *current* device, so callers use torch.cuda.set_device first).

Only tests/, oracle/make_golden.py, bench.py's baseline legs and tools/ import this file; nothing under
vi_depth_completion_b200/ does.  Every consumer checks reference_available() and skips when it is False.
"""
import contextlib
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("VIDC_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
_WARPER = os.path.join("networks", "warping_2dof_alignment.py")


def reference_root():
    for root in _CANDIDATES:
        if root and os.path.isfile(os.path.join(root, _WARPER)):
            return root
    return None


REF_ROOT = reference_root() or "/root/reference"


def reference_available() -> bool:
    return reference_root() is not None


def load_reference_module(device: str = "cpu") -> types.ModuleType:
    """networks/warping_2dof_alignment.py with the token substitutions, as a fresh module object."""
    path = os.path.join(reference_root() or REF_ROOT, _WARPER)
    with open(path, "r") as f:
        src = f.read()
    src = src.replace("'cuda:0'", repr(device))
    if device == "cpu":
        src = src.replace("torch.cuda.FloatTensor", "torch.FloatTensor")
    src = src.replace("240*320", "self.W*self.H").replace("240 * 320", "self.W*self.H")
    mod = types.ModuleType("reference_warping_2dof_alignment")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def load_reference_class(device: str = "cpu"):
    return load_reference_module(device).Warping2DOFAlignment


def load_reference_normal_utils() -> types.ModuleType:
    """normal_utils.py unmodified, with the undefined `Normalize` (:12,:24) injected as F.normalize(x, dim=1)
    (inferred intent)."""
    import torch.nn.functional as F
    path = os.path.join(reference_root() or REF_ROOT, "normal_utils.py")
    with open(path, "r") as f:
        src = f.read()
    mod = types.ModuleType("reference_normal_utils")
    mod.__file__ = path
    mod.Normalize = lambda t: F.normalize(t, dim=1)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


@contextlib.contextmanager
def _random_init_resnet():
    """surface_normal.py:13 and depth_completion.py:71 hard-code pretrained=True (a download).  Checkpoints are
    unavailable offline, so torchvision's resnet101 is built with weights=None while the reference modules are
    constructed (BASELINE config 5: random-init CNNs)."""
    import torchvision
    orig = torchvision.models.__dict__["resnet101"]

    def resnet101(*args, **kwargs):
        kwargs.pop("pretrained", None)
        kwargs["weights"] = None
        return orig(*args, **kwargs)

    torchvision.models.__dict__["resnet101"] = resnet101
    try:
        yield
    finally:
        torchvision.models.__dict__["resnet101"] = orig


class ReferenceNetworks:
    """The reference's `networks` package imported UNMODIFIED with `networks.warping_2dof_alignment` aliased to
    `warper_module` -- either the reference's own warper (load_reference_module('cuda:N'))
    or the drop-in (vi_depth_completion_b200.warping_2dof_alignment)."""

    def __init__(self, warper_module):
        root = reference_root()
        if root is None:
            raise RuntimeError("reference sources not available (neither /root/reference nor oracle/_ref)")
        saved = {k: v for k, v in sys.modules.items() if k == "networks" or k.startswith("networks.")}
        for k in saved:
            del sys.modules[k]
        pkg = types.ModuleType("networks")
        pkg.__path__ = [os.path.join(root, "networks")]
        sys.modules["networks"] = pkg
        sys.modules["networks.warping_2dof_alignment"] = warper_module
        try:
            self.surface_normal = importlib.import_module("networks.surface_normal")
            self.depth_completion = importlib.import_module("networks.depth_completion")
        finally:
            for k in [k for k in sys.modules if k == "networks" or k.startswith("networks.")]:
                del sys.modules[k]
            sys.modules.update(saved)

    def build(self, device, use_mask=False, seed=0):
        """(SurfaceNormalPrediction as main.py:243 constructs it, ModifiedFPN as network_run.py:96 does), random
        init, eval mode, on `device`."""
        import numpy as np
        import torch
        torch.manual_seed(seed)
        with _random_init_resnet():
            snp = self.surface_normal.SurfaceNormalPrediction(fc_img=np.array([202., 202.]), use_mask=use_mask)
            fpn = self.depth_completion.ModifiedFPN()
        return snp.to(device).eval(), fpn.to(device).eval()
