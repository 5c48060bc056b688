"""BASELINE config 5 as a test: the reference's own networks run UNMODIFIED on top of the drop-in.

networks/surface_normal.py (SurfaceNormalPrediction, constructs the warper at :70 and calls it at :148 / :169) and
networks/depth_completion.py (ModifiedFPN, called as main.py:261-275 does) are imported from oracle/_ref exactly as a user
would after the one-line change of INTEGRATION.md section 2 -- sys.modules['networks.warping_2dof_alignment'] aliased to
vi_depth_completion_b200.warping_2dof_alignment -- with random-init weights (checkpoints are unavailable offline; resnet101 is
built with weights=None).  The same networks with the reference's own warper executed on cuda:0 are the comparison.
Skipped when the reference files are not on the machine.
"""
import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu


def _nets(use_mask, which, device):
    import torch
    from oracle import ref_loader as RL
    if not RL.reference_available():
        pytest.skip("reference sources not on this machine (oracle/_ref is made by __graft_entry__.build())")
    torch.cuda.set_device(0)
    if which == "reference":
        warper = RL.load_reference_module("cuda:0")
    elif which == "reference-cpu-math":
        warper = _cpu_backed_reference_warper(RL, device)
    else:
        import vi_depth_completion_b200.warping_2dof_alignment as warper
    return RL.ReferenceNetworks(warper).build(device, use_mask=use_mask, seed=123)


def _cpu_backed_reference_warper(RL, device):
    """The reference's warper executed on the CPU (the backend the oracle pins, bit-identical to the drop-in) behind a shim that
    moves tensors to the host and back, so the CNNs around it still run on the GPU."""
    import types
    Ref = RL.load_reference_class("cpu")

    class Warping2DOFAlignment:
        def __init__(self, **kw):
            self._w = Ref(**kw)

        def warp_with_gravity_center_aligned(self, x, I_g, I_a, interp_mode='bilinear'):
            h, y = self._w.warp_with_gravity_center_aligned(x.cpu(), I_g.cpu(), I_a.cpu(), interp_mode)
            return h.to(device), y.to(device)

        def inverse_warp_normal_image_with_gravity_center_aligned(self, x, I_g, I_a):
            h, z = self._w.inverse_warp_normal_image_with_gravity_center_aligned(x.cpu(), I_g.cpu(), I_a.cpu())
            return h.to(device), z.to(device)

    return types.SimpleNamespace(Warping2DOFAlignment=Warping2DOFAlignment)


def _inputs(B, device):
    import torch
    I_g, I_a = C.random_gravity(B, seed=77, roll_deg=12, pitch_deg=12)
    rgb = C.smooth_images(B, 240, 320, seed=5)
    depth = C.random_images(B, 240, 320, seed=6, sparse_depth=True)[1][:, None]
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(device)
    return t(rgb), t(depth), t(I_g), t(I_a)


@pytest.mark.parametrize("use_mask", [False, True], ids=["use_mask=False", "use_mask=True"])
def test_reference_networks_run_unmodified_on_the_dropin(cuda_device, use_mask):
    import torch
    torch.backends.cudnn.deterministic = True
    rgb, depth, g, a = _inputs(2, cuda_device)
    snp_ref, fpn_ref = _nets(use_mask, "reference", cuda_device)
    snp_new, fpn_new = _nets(use_mask, "dropin", cuda_device)
    snp_cpu, fpn_cpu = _nets(use_mask, "reference-cpu-math", cuda_device)
    assert type(snp_new.warp_2dof_alignment).__module__ == "vi_depth_completion_b200.warping_2dof_alignment"
    assert type(snp_ref.warp_2dof_alignment).__module__ != type(snp_new.warp_2dof_alignment).__module__
    for p, q in zip(snp_ref.parameters(), snp_new.parameters()):
        assert torch.equal(p, q)                                  # same seed, same random init
    with torch.no_grad():
        n_ref = snp_ref(rgb, g, a)                                # main.py:267-269
        n_new = snp_new(rgb, g, a)
        d_ref = fpn_ref(rgb, n_ref, depth)                        # main.py:274
        d_new = fpn_new(rgb, n_new, depth)
        n_cpu = snp_cpu(rgb, g, a)
        d_cpu = fpn_cpu(rgb, n_cpu, depth)
    # With the reference's warper evaluated by its CPU backend (what the oracle pins) the whole pipeline -- warp, 101-layer CNN,
    # mask pyramid, inverse warp, F.normalize, ModifiedFPN -- is reproduced BIT FOR BIT by the drop-in:
    assert torch.equal(n_new, n_cpu) and torch.equal(d_new, d_cpu)
    assert n_new.shape == (2, 3, 240, 320) and d_new.shape == (2, 1, 240, 320)
    assert torch.isfinite(n_new).all() and torch.isfinite(d_new).all()
    # Against the reference's warper executed on CUDA the CNN sees a warped image that differs by the backends' one-ulp
    # coordinate drift (tests/test_gpu_reference_cuda.py); 101 random-init layers amplify it to (measured: median 0.024,
    # 99th percentile 0.07, max 0.31 degrees):
    ang, _ = C.angular_error_deg(n_new.cpu().numpy(), n_ref.cpu().numpy())
    ang = ang[np.isfinite(ang)]
    assert np.median(ang) <= 0.1 and np.percentile(ang, 99) <= 0.3 and ang.max() <= 2.0, (np.median(ang), np.percentile(ang, 99), ang.max())
    rel = (d_new - d_ref).abs().max().item() / max(d_ref.abs().max().item(), 1e-12)
    assert rel <= 5e-2, rel
    # zero vectors (outside the canvas footprint) are zero in both
    z_ref = (n_ref.abs().sum(1) == 0); z_new = (n_new.abs().sum(1) == 0)
    assert (z_ref != z_new).float().mean().item() <= 1e-4


def test_training_step_through_the_dropin(cuda_device):
    """surface_normal.py forward in train mode, normal_utils L1 loss (network_run.py:186), backward (:248): gradients reach the
    CNN through the drop-in's inverse warp (vidc_warp_backward) and match the reference's autograd through grid_sample / bmm."""
    import torch
    from oracle import ref_loader as RL
    from vi_depth_completion_b200 import normal_utils as NU
    rgb, depth, g, a = _inputs(2, cuda_device)
    gt = torch.nn.functional.normalize(torch.from_numpy(C.random_images(2, 240, 320, seed=9)[2]).to(cuda_device), dim=1)
    mask = torch.ones(2, 1, 240, 320, device=cuda_device)
    grads = {}
    for which in ("reference", "dropin"):
        snp, _ = _nets(False, which, cuda_device)
        snp.eval()                                               # BatchNorm with its initial running statistics: deterministic
        for p in snp.parameters():
            p.requires_grad_(True)
        pred = snp(rgb, g, a)
        if which == "reference":
            loss, _ = RL.load_reference_normal_utils().compute_normal_vectors_loss_l1(gt, pred, mask)
        else:
            loss, _ = NU.compute_normal_vectors_loss_l1(gt, pred, mask)
        loss.backward()
        last = [p for n, p in snp.named_parameters() if n.startswith("feature_concat")][-2]
        assert last.grad is not None and torch.isfinite(last.grad).all() and last.grad.abs().max() > 0
        grads[which] = (loss.item(), last.grad.detach().clone())
    assert abs(grads["dropin"][0] - grads["reference"][0]) <= 1e-3 * abs(grads["reference"][0])
    num = (grads["dropin"][1] - grads["reference"][1]).norm().item()
    den = grads["reference"][1].norm().item()
    assert num <= 2e-2 * den, (num, den)
