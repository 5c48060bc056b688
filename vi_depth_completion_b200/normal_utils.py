"""Drop-in replacement for the reference's normal_utils.py (:7-34) plus the mask lines of
networks/surface_normal.py (:150-156, :170), on the same C-ABI kernels.

Same function names and argument meaning as the reference; each loss function is ONE fused pass over HBM
(the reference makes ~8 element-wise passes and two reductions) and does not synchronise with the host
(the reference calls .item() at normal_utils.py:32).  `Normalize` is undefined in the reference
(normal_utils.py:12,24); the evident intent, F.normalize(x, dim=1), is what the kernels implement.
"""
import ctypes

import torch

from . import _cabi
from ._cabi import check, lib
from .warping_2dof_alignment import _image, _require_cuda_f32, _stream_ptr

__all__ = ["Normalize", "compute_normal_vectors_loss_l1", "compute_normal_vectors_loss_l2", "validity_mask",
           "pyramid_masks", "PYRAMID_SIZES"]

PYRAMID_SIZES = ((60, 80), (30, 40), (15, 20), (8, 10))     # networks/surface_normal.py:153-156


def Normalize(x):
    """F.normalize(x, dim=1) for (B,3,H,W) tensors (networks/surface_normal.py:170)."""
    _require_cuda_f32(x, "x")
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError("Normalize: expected (B,3,H,W)")
    out = torch.empty_like(x)
    xi, oi = _image(x), _image(out)
    with torch.cuda.device(x.device):
        check(lib().vidc_normalize3(ctypes.byref(xi), ctypes.byref(oi), _stream_ptr(x.device)))
    return out


def _stats(norm_gt, pred_normals, mask, normalize_prediction):
    _require_cuda_f32(norm_gt, "norm_gt")
    _require_cuda_f32(pred_normals, "pred_normals")
    if not mask.is_cuda:
        raise RuntimeError("mask: expected a CUDA tensor")
    mask = mask.float()                                           # normal_utils.py:21
    if mask.dim() == 3:
        mask = mask.view(mask.shape[0], 1, mask.shape[1], mask.shape[2])
    dev = norm_gt.device
    out = torch.empty(4, dtype=torch.float64, device=dev)
    gi, pi, mi = _image(norm_gt), _image(pred_normals), _image(mask)
    with torch.cuda.device(dev):
        check(lib().vidc_normal_stats(ctypes.byref(gi), ctypes.byref(pi), ctypes.byref(mi), 1 if normalize_prediction else 0,
                                      out.data_ptr(), _stream_ptr(dev)))
    return out        # [sum(angle*mask), sum(mask), sum|n*mask - gt*mask|, sum(cosine_similarity)]


# normal_utils.py:7-17
def compute_normal_vectors_loss_l2(norm_gt, pred_normals, mask):
    s = _stats(norm_gt, pred_normals, mask, True)
    loss = (-s[3] / s[1]).float()
    angle = s[0].float()
    return loss, angle


# normal_utils.py:20-34
def compute_normal_vectors_loss_l1(norm_gt, pred_normals, mask, normalize_prediction=True):
    s = _stats(norm_gt, pred_normals, mask, normalize_prediction)
    loss = (s[2] / s[1]).float()
    angle = s[0].float()
    return loss, angle


# networks/surface_normal.py:151-152
def validity_mask(x1, as_float=True, with_coverage=False):
    """(x1[:,0:1] + x1[:,1:2] + x1[:,2:3] > 1e-2).float() in one pass."""
    _require_cuda_f32(x1, "x1")
    if x1.dim() != 4 or x1.shape[1] < 3:
        raise RuntimeError("validity_mask: expected (B,>=3,H,W)")
    B, _, H, W = x1.shape
    dev = x1.device
    mf = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev) if as_float else None
    mu = torch.empty((B, 1, H, W), dtype=torch.uint8, device=dev) if not as_float else None
    cov = torch.empty((B,), dtype=torch.int32, device=dev) if with_coverage else None
    xi = _image(x1)
    with torch.cuda.device(dev):
        check(lib().vidc_validity_mask(ctypes.byref(xi), mu.data_ptr() if mu is not None else None,
                                       mf.data_ptr() if mf is not None else None,
                                       cov.data_ptr() if cov is not None else None, _stream_ptr(dev)))
    m = mf if as_float else mu
    return (m, cov) if with_coverage else m


# networks/surface_normal.py:153-156
def pyramid_masks(feature_mask, sizes=PYRAMID_SIZES):
    """F.interpolate(feature_mask, size=s, mode='nearest') for every pyramid level, ONE kernel launch.
    feature_mask: (B,1,H,W) float32 (as the reference builds it) or the uint8 mask written by warp_rgbd."""
    if not isinstance(feature_mask, torch.Tensor) or not feature_mask.is_cuda:
        raise RuntimeError("feature_mask: expected a CUDA tensor (this module has no CPU fallback)")
    if feature_mask.dtype not in (torch.float32, torch.uint8):
        raise RuntimeError(f"feature_mask: expected float32 or uint8, got {feature_mask.dtype}")
    if len(sizes) < 1 or len(sizes) > 4:
        raise RuntimeError("pyramid_masks: 1..4 levels")
    m = feature_mask.contiguous()
    B, _, Hin, Win = m.shape
    outs = [torch.empty((B, 1, int(h), int(w)), dtype=torch.float32, device=m.device) for (h, w) in sizes]
    flat = (ctypes.c_int32 * (2 * len(sizes)))(*[int(v) for hw in sizes for v in hw])
    ptrs = (ctypes.c_void_p * len(sizes))(*[o.data_ptr() for o in outs])
    with torch.cuda.device(m.device):
        check(lib().vidc_mask_pyramid(m.data_ptr() if m.dtype == torch.uint8 else None,
                                      m.data_ptr() if m.dtype == torch.float32 else None,
                                      B, Hin, Win, len(sizes), flat, ptrs, _stream_ptr(m.device)))
    return outs
