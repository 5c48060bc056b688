"""Dataset-side gravity conditioning on the GPU.

The reference does this per sample on the host inside its Dataset classes (dataset.py:45-55 for ScanNet,
:334-345 / :472-483 for the Azure-Kinect and Demo loaders): sign flip of y and z, then the choice of the
alignment direction `I_a` from the pitch angle.  Here it is one thread per frame on the device, with the same
roundings (glibc atan2f, MKL VML cos / sin) so `I_g, I_a` are bit-identical to what the DataLoader would feed."""
import ctypes

import torch

from ._cabi import check, lib
from .warping_2dof_alignment import _require_cuda_f32, _stream_ptr

RULES = {"azure": 0, "demo": 0, "scannet": 1}


def condition_gravity(raw_gravity: torch.Tensor, rule: str = "azure"):
    """raw_gravity (B,3) float32 CUDA -> (I_g, I_a), both (B,3)."""
    _require_cuda_f32(raw_gravity, "raw_gravity")
    if rule not in RULES:
        raise RuntimeError(f"rule must be one of {sorted(RULES)}, got {rule!r}")
    raw = raw_gravity.reshape(-1, 3).contiguous()
    B = raw.shape[0]
    I_g = torch.empty_like(raw)
    I_a = torch.empty_like(raw)
    with torch.cuda.device(raw.device):
        check(lib().vidc_condition_gravity(raw.data_ptr(), B, RULES[rule], I_g.data_ptr(), I_a.data_ptr(), _stream_ptr(raw.device)))
    return I_g, I_a
