"""The thin torch C++ extension (csrc/torch_ops.cpp, `torch.ops.vidc.*`) in front of the C ABI: same bits as the ctypes
route, reference error behaviour, traceable by torch.compile(fullgraph=True), and cheaper on the host."""
import time

import numpy as np
import pytest

from tests import common as C

pytestmark = pytest.mark.gpu


def _setup(dev, cam="S1", B=4):
    import torch
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    w = Warping2DOFAlignment(*C.CAMERAS[cam])
    H, W = int(w.H), int(w.W)
    I_g, I_a = C.random_gravity(B, seed=3)
    rgb, depth, normals = C.random_images(B, H, W, seed=2, sparse_depth=True)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    return w, t(rgb), t(depth), t(normals), t(I_g), t(I_a)


def test_extension_is_loaded_and_matches_the_ctypes_route(cuda_device):
    import torch
    from vi_depth_completion_b200 import _torchops
    assert _torchops.ops() is not None, "_vidc_torch_ops.so missing: __graft_entry__.build() builds it"
    w, rgb, depth, normals, g, a = _setup(cuda_device)
    outs = {}
    for tag in ("torch", "ctypes"):
        saved = dict(_torchops._state)
        if tag == "ctypes":
            _torchops._state.update(tried=True, ops=None)
        try:
            H1, y = w.warp_with_gravity_center_aligned(rgb, g, a)
            _, yd = w.warp_with_gravity_center_aligned(depth, g, a, interp_mode="nearest")
            _, z = w.inverse_warp_normal_image_with_gravity_center_aligned(normals, g, a)
            _, zn = w.unwarp_normals(normals, g, a)
            _, r2, d2, m2 = w.warp_rgbd(rgb, depth, g, a)
            Hh, R, Hi = w._build_homography(g, a)
            _, ycl = w.warp_with_gravity_center_aligned(rgb.contiguous(memory_format=torch.channels_last), g, a)
            outs[tag] = (H1, y, yd, z, zn, r2, d2, m2, Hh, R, Hi, ycl)
        finally:
            _torchops._state.update(saved)
    for p, q in zip(outs["torch"], outs["ctypes"]):
        assert p.shape == q.shape and p.stride() == q.stride() and torch.equal(p, q)
    assert outs["torch"][-1].is_contiguous(memory_format=torch.channels_last)


def test_extension_error_behaviour(cuda_device):
    import torch
    w, rgb, depth, normals, g, a = _setup(cuda_device)
    with pytest.raises(AssertionError):                                # reference :123
        w.warp_with_gravity_center_aligned(rgb[:3], g, a)
    with pytest.raises(AssertionError):                                # reference :224
        w.inverse_warp_normal_image_with_gravity_center_aligned(normals[:2], g, a)
    with pytest.raises(IndexError):                                    # reference :41
        w.warp_with_gravity_center_aligned(rgb, g, a[:1])
    with pytest.raises(RuntimeError):
        w.warp_with_gravity_center_aligned(rgb.double(), g, a)
    with pytest.raises(RuntimeError):
        torch.ops.vidc.warp_forward(rgb.cpu(), g, a, 202., 202., 159.9, 119.9, 0)      # no CPU kernel is registered


def test_torch_compile_fullgraph(cuda_device):
    import torch
    w, rgb, depth, normals, g, a = _setup(cuda_device)

    def path(x, n, gg, aa):
        _, x1 = w.warp_with_gravity_center_aligned(x, gg, aa)
        _, z = w.inverse_warp_normal_image_with_gravity_center_aligned(n * 2.0, gg, aa)
        return x1 + 1.0, torch.nn.functional.normalize(z, dim=1)

    with torch.no_grad():
        want = path(rgb, normals, g, a)
        got = torch.compile(path, fullgraph=True, backend="aot_eager")(rgb, normals, g, a)
    assert torch.equal(want[0], got[0]) and torch.equal(want[1], got[1])


def test_host_time_per_call(cuda_device):
    """B = 1 at 320x240 (the reference's only documented entry point, demo.sh:7-11): host time of one eager step through the
    extension against the ctypes route (round 1: 88.7 us per step of two calls)."""
    import torch
    from vi_depth_completion_b200 import _torchops
    w, rgb, depth, normals, g, a = _setup(cuda_device, B=1)

    def step():
        w.warp_rgbd(rgb, depth, g, a)
        w.unwarp_normals(normals, g, a)

    res = {}
    for tag in ("torch", "ctypes"):
        saved = dict(_torchops._state)
        if tag == "ctypes":
            _torchops._state.update(tried=True, ops=None)
        try:
            for _ in range(200):
                step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2000):
                step()
            res[tag] = (time.perf_counter() - t0) / 2000 * 1e6
            torch.cuda.synchronize()
        finally:
            _torchops._state.update(saved)
    print(f"host us per step (warp_rgbd + unwarp_normals, B=1): extension {res['torch']:.1f}, ctypes {res['ctypes']:.1f}")
    assert res["torch"] < res["ctypes"]
