"""Full-size checks at BASELINE.json's bench configuration (640x480, 256 frames) through size-independent
properties -- the oracle cannot finish this size in seconds:
  * a strided sample of frames is compared bit-for-bit with the oracle,
  * batch independence: frame i of the 256-batch equals the same frame warped alone,
  * linearity of the resample in the image: warp(a x + b y) ~= a warp(x) + b warp(y),
  * coverage == sum(mask); identity gravity keeps interior pixels of a smooth image within the half-pixel shift,
  * idempotence of the renormalisation; zero vectors stay exactly zero; every non-zero normal has unit length.
"""
import numpy as np
import pytest
import torch

from tests import common as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(cuda_device):
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    w = Warping2DOFAlignment(*C.CAMERAS["S2"])
    B, H, W = 256, int(w.H), int(w.W)
    I_g, I_a = C.random_gravity(B, seed=1234)
    gen = torch.Generator(device=cuda_device).manual_seed(7)
    rgb = torch.rand(B, 3, H, W, device=cuda_device, generator=gen)
    depth = torch.rand(B, 1, H, W, device=cuda_device, generator=gen) * 9.6 + 0.4
    nrm = torch.randn(B, 3, H, W, device=cuda_device, generator=gen)
    g, a = torch.from_numpy(I_g).to(cuda_device), torch.from_numpy(I_a).to(cuda_device)
    return dict(w=w, B=B, H=H, W=W, I_g=I_g, I_a=I_a, rgb=rgb, depth=depth, nrm=nrm, g=g, a=a)


def test_fullsize_sampled_frames_match_oracle(big, oracle_mod):
    w = big["w"]
    _, rgb_w, depth_w, mask, cov = w.warp_rgbd(big["rgb"], big["depth"], big["g"], big["a"], with_coverage=True)
    _, nhat = w.unwarp_normals(big["nrm"], big["g"], big["a"])
    idx = [0, 37, 101, 255]
    o = oracle_mod.Oracle(*C.CAMERAS["S2"])
    sel = lambda t: t[idx].cpu().numpy()
    o_rgb, o_depth, o_mask, o_n = oracle_mod.warp_unwarp_mt(o, sel(big["rgb"]), sel(big["depth"]), sel(big["nrm"]),
                                                          big["I_g"][idx], big["I_a"][idx], 4)
    assert C.count_bit_mismatches(sel(rgb_w), o_rgb) == 0
    assert C.count_bit_mismatches(sel(depth_w)[:, 0], o_depth) == 0
    assert np.array_equal(sel(mask), o_mask)
    assert C.count_bit_mismatches(sel(nhat), o_n) == 0
    assert torch.equal(cov.long(), mask.view(big["B"], -1).sum(1))          # coverage == sum(mask)


def test_fullsize_batch_independence(big):
    w = big["w"]
    _, rgb_w, depth_w, mask = w.warp_rgbd(big["rgb"], big["depth"], big["g"], big["a"])
    _, nhat = w.unwarp_normals(big["nrm"], big["g"], big["a"])
    for i in (3, 200):
        s = slice(i, i + 1)
        _, r1, d1, m1 = w.warp_rgbd(big["rgb"][s], big["depth"][s], big["g"][s], big["a"][s])
        _, n1 = w.unwarp_normals(big["nrm"][s], big["g"][s], big["a"][s])
        assert torch.equal(r1, rgb_w[s]) and torch.equal(d1, depth_w[s]) and torch.equal(m1, mask[s]) and torch.equal(n1, nhat[s])


def test_fullsize_linearity_and_normalisation(big):
    w = big["w"]
    x, y = big["rgb"][:64], big["rgb"][64:128]
    g, a = big["g"][:64], big["a"][:64]
    _, wx = w.warp_with_gravity_center_aligned(x, g, a)
    _, wy = w.warp_with_gravity_center_aligned(y, g, a)
    _, wz = w.warp_with_gravity_center_aligned(0.25 * x + 0.5 * y, g, a)
    assert (wz - (0.25 * wx + 0.5 * wy)).abs().max().item() <= 2e-6          # bilinear resampling is linear in the image
    _, nhat = w.unwarp_normals(big["nrm"][:64], g, a)
    n = nhat.norm(dim=1)
    zero = (nhat == 0).all(dim=1)
    assert ((n - 1).abs()[~zero] <= 2e-6).all()                             # unit length wherever the canvas was hit
    from vi_depth_completion_b200.normal_utils import Normalize
    assert (Normalize(nhat) - nhat).abs().max().item() <= 2e-7               # renormalising again changes nothing


def test_fullsize_identity_gravity(big):
    """g == a: R = I exactly, the warp is the reference's fixed sub-pixel shift (cx vs W/2, align_corners=False);
    interior pixels of a smooth image stay within the shift times the image gradient."""
    w, dev = big["w"], big["rgb"].device
    B, H, W = 8, big["H"], big["W"]
    x = torch.from_numpy(C.smooth_images(B, H, W, seed=3)).to(dev)
    g = torch.tensor([[0.0, 1.0, 0.0]], device=dev).repeat(B, 1)
    Hm, R, _ = w._build_homography(g, g)
    assert torch.equal(R, torch.eye(3, device=dev).expand(B, 3, 3))
    _, y = w.warp_with_gravity_center_aligned(x, g, g)
    assert (y[:, :, 2:-2, 2:-2] - x[:, :, 2:-2, 2:-2]).abs().max().item() < 0.05
