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


def rasterize_sparse_depth(tracks: torch.Tensor, counts, fc, cc, H: int, W: int):
    """Row f2: (B,N,>=4) float64 KLT tracks [id, x, y, z, ...] -> (B,1,H,W) float32 sparse depth, as dataset.py:496-510 builds it
    per sample on the host (pixel = int(fc * xy / z + cc) in fp64, last point on a pixel wins)."""
    if not tracks.is_cuda or tracks.dtype != torch.float64 or tracks.dim() != 3 or tracks.shape[2] < 4:
        raise RuntimeError("tracks: expected a (B,N,>=4) float64 CUDA tensor")
    tr = tracks.contiguous()
    B, N, cols = tr.shape
    cnt = None
    if counts is not None:
        cnt = torch.as_tensor(counts, dtype=torch.int32, device=tr.device).contiguous()
        if cnt.shape != (B,):
            raise RuntimeError("counts: expected shape (B,)")
    ws = torch.empty((B, H, W), dtype=torch.int32, device=tr.device)
    depth = torch.empty((B, 1, H, W), dtype=torch.float32, device=tr.device)
    with torch.cuda.device(tr.device):
        check(lib().vidc_rasterize_sparse_depth(tr.data_ptr(), cnt.data_ptr() if cnt is not None else None, B, N, cols,
                                                float(fc[0]), float(fc[1]), float(cc[0]), float(cc[1]), H, W,
                                                ws.data_ptr(), depth.data_ptr(), _stream_ptr(tr.device)))
    return depth


def to_tensor_u8(images: torch.Tensor):
    """torchvision's ToTensor on the device (dataset.py:468-471): (B,H,W,C) or (H,W,C) uint8 CUDA, the layout PIL decodes to
    -> (B,C,H,W) / (C,H,W) float32 = x / 255, bit-identical to `transforms.ToTensor()` (one correctly rounded division)."""
    if not images.is_cuda or images.dtype != torch.uint8 or images.dim() not in (3, 4):
        raise RuntimeError("images: expected a (B,H,W,C) or (H,W,C) uint8 CUDA tensor")
    x = images.contiguous()
    single = x.dim() == 3
    if single:
        x = x.unsqueeze(0)
    B, H, W, C = x.shape
    if not 1 <= C <= 4:
        raise RuntimeError(f"images: 1..4 channels, got {C}")
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        for b0 in range(0, B, 65535):
            n = min(65535, B - b0)
            check(lib().vidc_to_tensor_u8(x[b0:b0 + n].data_ptr(), n, H, W, C, out[b0:b0 + n].data_ptr(), _stream_ptr(x.device)))
    return out[0] if single else out
