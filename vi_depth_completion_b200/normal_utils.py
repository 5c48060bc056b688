"""Drop-in replacement for the reference's normal_utils.py (:7-34) plus the mask lines of
networks/surface_normal.py (:150-156, :170), on the same C-ABI kernels.

Same function names and argument meaning as the reference; each loss function is ONE fused pass over HBM
(the reference makes ~8 element-wise passes and two reductions) and does not synchronise with the host
(the reference calls .item() at normal_utils.py:32), and is differentiable w.r.t. pred_normals through one fused backward
pass (the reference back-propagates this loss, network_run.py:186,248).  `Normalize` is undefined in the reference
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


def _mask4(mask):
    if not isinstance(mask, torch.Tensor) or not mask.is_cuda:
        raise RuntimeError("mask: expected a CUDA tensor")
    mask = mask.float()                                           # normal_utils.py:21
    if mask.dim() == 3:
        mask = mask.view(mask.shape[0], 1, mask.shape[1], mask.shape[2])
    return mask


def _stats(norm_gt, pred_normals, mask, normalize_prediction):
    _require_cuda_f32(norm_gt, "norm_gt")
    _require_cuda_f32(pred_normals, "pred_normals")
    mask = _mask4(mask)
    dev = norm_gt.device
    out = torch.empty(4, dtype=torch.float64, device=dev)
    gi, pi, mi = _image(norm_gt), _image(pred_normals), _image(mask)
    with torch.cuda.device(dev):
        check(lib().vidc_normal_stats(ctypes.byref(gi), ctypes.byref(pi), ctypes.byref(mi), 1 if normalize_prediction else 0,
                                      out.data_ptr(), _stream_ptr(dev)))
    return out        # [sum(angle*mask), sum(mask), sum|n*mask - gt*mask|, sum(cosine_similarity)]


_L1_NORMALIZED, _L1_RAW, _L2_COSINE = 0, 1, 2


class _NormalLossFn(torch.autograd.Function):
    """(loss, angle) of normal_utils.py:7-34 with the gradient of the loss w.r.t. pred_normals -- the reference
    back-propagates it (network_run.py:186 -> total_loss.backward() at :248).  Forward: one fused pass
    (vidc_normal_stats); backward: one fused pass (vidc_normal_loss_backward), no host synchronisation in either.
    `angle` is the logged metric (network_run.py:189, `.item()`): it is marked non-differentiable here, while in the
    reference it happens to carry a graph nobody uses.  norm_gt and mask are data and get no gradient."""

    @staticmethod
    def forward(ctx, pred_normals, norm_gt, mask, loss_mode):
        s = _stats(norm_gt, pred_normals, mask, loss_mode != _L1_RAW)
        loss = ((-s[3] if loss_mode == _L2_COSINE else s[2]) / s[1]).float()
        angle = s[0].float()
        ctx.loss_mode = loss_mode
        ctx.save_for_backward(pred_normals, norm_gt, _mask4(mask), s)
        ctx.mark_non_differentiable(angle)
        return loss, angle

    @staticmethod
    def backward(ctx, grad_loss, _grad_angle):
        pred, gt, mask, s = ctx.saved_tensors
        dev = pred.device
        gp = torch.empty_like(pred)
        go = grad_loss.to(device=dev, dtype=torch.float32).contiguous()
        gi, pi, mi, oi = _image(gt), _image(pred), _image(mask), _image(gp)
        with torch.cuda.device(dev):
            check(lib().vidc_normal_loss_backward(ctypes.byref(gi), ctypes.byref(pi), ctypes.byref(mi), ctx.loss_mode,
                                                  s.data_ptr(), go.data_ptr(), ctypes.byref(oi), _stream_ptr(dev)))
        return gp, None, None, None


def _loss(norm_gt, pred_normals, mask, loss_mode):
    if isinstance(pred_normals, torch.Tensor) and pred_normals.dim() != 4:
        raise RuntimeError("pred_normals: expected (B,>=3,H,W)")
    return _NormalLossFn.apply(pred_normals, norm_gt, mask, loss_mode)


# normal_utils.py:7-17
def compute_normal_vectors_loss_l2(norm_gt, pred_normals, mask):
    return _loss(norm_gt, pred_normals, mask, _L2_COSINE)


# normal_utils.py:20-34
def compute_normal_vectors_loss_l1(norm_gt, pred_normals, mask, normalize_prediction=True):
    return _loss(norm_gt, pred_normals, mask, _L1_NORMALIZED if normalize_prediction else _L1_RAW)


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
