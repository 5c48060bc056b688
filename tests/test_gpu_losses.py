"""normal_utils drop-in: the losses are differentiable w.r.t. pred_normals, as the reference's are
(normal_utils.py:7-34 is back-propagated at network_run.py:186 -> total_loss.backward() :248).

Gradients are checked against (i) goldens frozen from the executed reference's autograd on CPU
(oracle/make_golden.py: loss_backward_golden) and (ii), when oracle/_ref travelled to the box, the reference's
normal_utils executed on the same CUDA device.  Tolerance: rounding level -- |d - d_ref| <= 2e-5 * (|d_ref| + the largest
gradient component of the same pixel); the chain holds cancellations, so single components can be far smaller than the
pixel's gradient.
"""
import os

import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(d, ref, rel=2e-5):
    d = np.asarray(d, np.float64); ref = np.asarray(ref, np.float64)
    scale = np.abs(ref) + np.abs(ref[:, :3]).max(1, keepdims=True)
    err = np.abs(d - ref)
    bad = err > rel * scale + 1e-30
    return int(bad.sum()), float((err / (scale + 1e-300)).max())


def _fns():
    from vi_depth_completion_b200 import normal_utils as NU
    return {"l1": lambda p, g, m: NU.compute_normal_vectors_loss_l1(g, p, m),
            "l1_raw": lambda p, g, m: NU.compute_normal_vectors_loss_l1(g, p, m, normalize_prediction=False),
            "l2": lambda p, g, m: NU.compute_normal_vectors_loss_l2(g, p, m)}


@pytest.mark.parametrize("name", ["l1", "l1_raw", "l2"])
def test_loss_gradients_match_reference_autograd_golden(cuda_device, name):
    import torch
    gold = np.load(os.path.join(GOLD, "golden_tiny_loss_backward.npz"))
    pred, gt, maskf, up = C.loss_inputs(3, 48, 64, seed=int(gold["seed"]))
    t = lambda x: torch.from_numpy(x).to(cuda_device)
    p = t(pred).requires_grad_(True)
    loss, angle = _fns()[name](p, t(gt), t(maskf))
    assert loss.grad_fn is not None, "the loss must carry an autograd graph (it is the reference's training loss)"
    assert not angle.requires_grad
    (loss * float(up)).backward()
    assert np.isclose(float(loss.detach()), float(gold[f"{name}_loss"]), rtol=2e-5)
    assert np.isclose(float(angle), float(gold[f"{name}_angle"]), rtol=2e-5)
    nbad, worst = _close(p.grad.cpu().numpy(), gold[f"{name}_grad"])
    assert nbad == 0, f"{name}: {nbad} gradient components off, worst relative error {worst:.3g}"
    assert np.array_equal(p.grad[:, 3].cpu().numpy(), np.zeros_like(pred[:, 3])), "unused channels get exactly zero"


def test_loss_is_part_of_a_training_step(cuda_device):
    """The silent-detach bug of round 1: depth loss + normal loss, total.backward() must reach the normals' producer."""
    import torch
    from vi_depth_completion_b200 import normal_utils as NU
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 3, 3, padding=1).to(cuda_device)
    x = torch.randn(2, 3, 48, 64, device=cuda_device)
    gt = torch.nn.functional.normalize(torch.randn(2, 3, 48, 64, device=cuda_device), dim=1)
    mask = (torch.rand(2, 1, 48, 64, device=cuda_device) > 0.3)
    pred = conv(x)
    loss, angle = NU.compute_normal_vectors_loss_l1(gt, pred, mask)
    total = 0.0
    total += pred.abs().mean() * 0.0          # another loss term in the sum, as network_run.py:241-246 builds it
    total += loss
    total.backward()
    g_ours = conv.weight.grad.clone()
    conv.zero_grad()
    # the same step with the reference's expression in torch (Normalize = F.normalize(dim=1))
    pred = conv(x)
    m = mask.float()
    norms = torch.nn.functional.normalize(pred[:, 0:3], dim=1)
    ref = torch.nn.L1Loss(reduction='sum')(norms * m, gt * m) / torch.sum(m).item()
    ref.backward()
    assert torch.allclose(loss.detach(), ref.detach(), rtol=1e-5)
    assert torch.allclose(g_ours, conv.weight.grad, rtol=1e-4, atol=1e-7)
    assert float(g_ours.abs().max()) > 0


def test_loss_gradients_channels_last_and_no_grad(cuda_device):
    import torch
    pred, gt, maskf, up = C.loss_inputs(2, 40, 56, seed=5)
    t = lambda x: torch.from_numpy(x).to(cuda_device)
    fns = _fns()
    for name in fns:
        p1 = t(pred).requires_grad_(True)
        p2 = t(pred).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        l1, _ = fns[name](p1, t(gt), t(maskf)); l1.backward()
        l2, _ = fns[name](p2, t(gt).contiguous(memory_format=torch.channels_last), t(maskf) > 0); l2.backward()
        assert torch.equal(l1.detach(), l2.detach())
        assert torch.equal(p1.grad, p2.grad), f"{name}: memory format changed the gradient"
        with torch.no_grad():
            l3, _ = fns[name](t(pred), t(gt), t(maskf))
        assert l3.grad_fn is None and torch.equal(l3, l1.detach())


@pytest.mark.parametrize("name", ["l1", "l1_raw", "l2"])
def test_loss_gradients_match_reference_on_cuda(cuda_device, name):
    """Live: the reference's normal_utils.py executed on the same device (oracle/_ref)."""
    import torch
    from oracle import ref_loader as RL
    if not RL.reference_available():
        pytest.skip("reference sources not on this machine (oracle/_ref is made by build())")
    nu = RL.load_reference_normal_utils()
    ref_fns = {"l1": lambda p, g, m: nu.compute_normal_vectors_loss_l1(g, p, m),
               "l1_raw": lambda p, g, m: nu.compute_normal_vectors_loss_l1(g, p, m, normalize_prediction=False),
               "l2": lambda p, g, m: nu.compute_normal_vectors_loss_l2(g, p, m)}
    pred, gt, maskf, up = C.loss_inputs(4, 240, 320, seed=23)
    t = lambda x: torch.from_numpy(x).to(cuda_device)
    p = t(pred).requires_grad_(True); pr = t(pred).requires_grad_(True)
    loss, angle = _fns()[name](p, t(gt), t(maskf)); (loss * 0.37).backward()
    rloss, rangle = ref_fns[name](pr, t(gt), t(maskf)); (rloss * 0.37).backward()
    assert np.isclose(float(loss.detach()), float(rloss.detach()), rtol=2e-5)
    assert np.isclose(float(angle), float(rangle.detach()), rtol=2e-5)
    nbad, worst = _close(p.grad.cpu().numpy(), pr.grad.cpu().numpy(), rel=5e-4)     # fp32 chains of different order
    assert nbad == 0, f"{name}: {nbad} gradient components off, worst relative error {worst:.3g}"
