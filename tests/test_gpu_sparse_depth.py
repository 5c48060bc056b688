"""SURVEY 8 row f2, second half: the sparse depth of the KLT tracks warped analytically (vidc_warp_rgb_sparse_depth) must be the
dense path's result bit for bit -- rasterise (dataset.py:496-510) + resample the mostly-zero image
(warping_2dof_alignment.py:108-156) -- on the real demo tracks (golden from the executed reference) and on synthetic tracks
that stress it: many points, adjacent and duplicate pixels (last point wins), points outside the image, steep frames whose
horizon crosses the image, both interpolation modes."""
import hashlib
import os

import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_demo_tracks_match_the_executed_reference(cuda_device):
    import torch
    from vi_depth_completion_b200.gravity import condition_gravity
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    g = np.load(os.path.join(GOLD, "golden_demo.npz"))
    w = Warping2DOFAlignment(*C.CAMERAS["S1"])
    Hh, Ww, B = int(w.H), int(w.W), 8
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    Ig, Ia = condition_gravity(t(g["raw_gravity"]), "azure")
    rgb = t(C.smooth_images(B, Hh, Ww, int(g["rgb_seed"])))
    H, rgb_w, depth_w, mask = w.warp_rgb_sparse_depth(rgb, t(g["tracks"]), g["counts"], g["fc"], g["cc"], Ig, Ia)
    _, _, depth_n, _ = w.warp_rgb_sparse_depth(rgb, t(g["tracks"]), g["counts"], g["fc"], g["cc"], Ig, Ia, depth_mode="nearest")
    assert C.count_bit_mismatches(H.cpu().numpy(), g["Hm"]) == 0
    assert _sha(rgb_w.cpu().numpy()) == str(g["y_rgb_sha256"])
    assert _sha(depth_w.cpu().numpy().reshape(B, Hh, Ww)) == str(g["y_depth_sha256"])
    assert _sha(depth_n.cpu().numpy().reshape(B, Hh, Ww)) == str(g["y_depth_nearest_sha256"])
    assert _sha(mask.cpu().numpy().reshape(B, 1, Hh, Ww)) == str(g["mask_sha256"])
    assert int((depth_w != 0).sum()) > 0


def _synthetic_tracks(B, N, H, W, fc, cc, seed):
    rs = np.random.RandomState(seed)
    tr = np.zeros((B, N, 5), np.float64)
    z = rs.uniform(0.4, 6.0, size=(B, N))
    col = rs.uniform(-20, W + 20, size=(B, N)); row = rs.uniform(-20, H + 20, size=(B, N))     # some land outside the image
    col[:, : N // 8] = np.floor(col[:, N // 8: 2 * (N // 8)]) + 1.3                         # horizontal neighbours
    row[:, : N // 8] = row[:, N // 8: 2 * (N // 8)]
    k = N // 10
    col[:, N - k:] = col[:, :k]; row[:, N - k:] = row[:, :k]                                # duplicates: the last point wins
    tr[..., 0] = np.arange(N)
    tr[..., 1] = (col - cc[0]) / fc[0] * z
    tr[..., 2] = (row - cc[1]) / fc[1] * z
    tr[..., 3] = z
    tr[0, 3, 3] = np.nan; tr[0, 4, 3] = 0.0                                                # rows the loader skips / that divide by zero
    counts = np.full(B, N, np.int32); counts[-1] = N // 2
    return tr, counts


@pytest.mark.parametrize("cam_name,N,grav", [("S1", 150, "random"), ("S2", 700, "random"), ("S2", 2048, "steep"), ("S3", 300, "roll"), ("tiny", 64, "edge")])
def test_sparse_path_is_the_dense_path(cuda_device, cam_name, N, grav):
    import torch
    from vi_depth_completion_b200.gravity import rasterize_sparse_depth
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    w = Warping2DOFAlignment(*C.CAMERAS[cam_name])
    H, W = int(w.H), int(w.W)
    if grav == "random":
        I_g, I_a = C.random_gravity(6, seed=9, roll_deg=35, pitch_deg=35)
    elif grav == "steep":
        I_g, I_a = C.random_gravity(6, seed=10, roll_deg=60, pitch_deg=80)
        ig2, ia2 = C.isolated_nonfinite_gravity(); I_g[:2], I_a[:2] = ig2[:2], ia2[:2]
    elif grav == "roll":
        I_g, I_a = C.extreme_roll_gravity(7, seed=5)
    else:
        I_g, I_a = C.edge_case_gravity()
    B = I_g.shape[0]
    fc, cc = (C.CAMERAS[cam_name][0] * 1.005, C.CAMERAS[cam_name][1] * 1.004), (W / 2 - 0.3, H / 2 + 1.1)
    tr, counts = _synthetic_tracks(B, N, H, W, fc, cc, seed=3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda_device)
    rgb = t(C.random_images(B, H, W, seed=1)[0])
    g, a = t(I_g), t(I_a)
    dense = rasterize_sparse_depth(t(tr), counts, fc, cc, H, W)
    assert int((dense != 0).sum()) > B * N // 4
    for mode in ("bilinear", "nearest"):
        H1, r1, d1, m1, c1 = w.warp_rgbd(rgb, dense, g, a, depth_mode=mode, with_coverage=True)
        H2, r2, d2, m2, c2 = w.warp_rgb_sparse_depth(rgb, t(tr), counts, fc, cc, g, a, depth_mode=mode, with_coverage=True)
        assert torch.equal(H1, H2) and torch.equal(r1, r2) and torch.equal(m1, m2) and torch.equal(c1, c2)
        d1n, d2n = d1.cpu().numpy(), d2.cpu().numpy()
        assert C.count_bit_mismatches(d1n, d2n) == 0, f"{mode}: {int((d1n != d2n).sum())} canvas pixels differ from the dense resample"


def test_sparse_path_errors(cuda_device):
    import torch
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    w = Warping2DOFAlignment(*C.CAMERAS["S1"])
    g = torch.zeros(2, 3, device=cuda_device); g[:, 1] = 1
    rgb = torch.zeros(2, 3, 240, 320, device=cuda_device)
    tr = torch.zeros(2, 10, 5, dtype=torch.float64, device=cuda_device)
    with pytest.raises(RuntimeError, match="2048"):
        w.warp_rgb_sparse_depth(rgb, torch.zeros(2, 3000, 5, dtype=torch.float64, device=cuda_device), None, (200, 200), (160, 120), g, g)
    with pytest.raises(RuntimeError):
        w.warp_rgb_sparse_depth(rgb, tr.float(), None, (200, 200), (160, 120), g, g)
    with pytest.raises(AssertionError):
        w.warp_rgb_sparse_depth(rgb, tr[:1], None, (200, 200), (160, 120), g, g)
    with pytest.raises(RuntimeError, match="canvas size"):
        w.warp_rgb_sparse_depth(rgb[:, :, :100], tr, None, (200, 200), (160, 120), g, g)
