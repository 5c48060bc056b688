"""Loader for the *executed* reference (test infrastructure, NOT product code).

Reads networks/warping_2dof_alignment.py from the read-only reference checkout
(/root/reference, present only in the build container -- never on the GPU box),
applies the three token substitutions documented in SURVEY.md Appendix B and
exec()s the result.  No arithmetic is changed:

  'cuda:0'                  -> target device string          (ref :7)
  torch.cuda.FloatTensor    -> torch.FloatTensor (CPU only)  (ref :45,:118-122,...)
  240*320 / 240 * 320       -> self.W*self.H                 (ref :121-122)

Used only by oracle/make_golden.py and by container-only tests that validate the
C restatement (oracle/warp_oracle.c) against the real thing.  Nothing under
tests -m gpu, smoke() or bench.py imports this file.
"""
import os
import types

REF_ROOT = os.environ.get("VIDC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks", "warping_2dof_alignment.py"))


def load_reference_module(device: str = "cpu") -> types.ModuleType:
    path = os.path.join(REF_ROOT, "networks", "warping_2dof_alignment.py")
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
