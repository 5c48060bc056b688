"""The data format on the DataLoader side of the path (dataset.py:468-471): PIL decodes (H,W,3) uint8, `transforms.ToTensor()`
turns it into (3,H,W) float32 = x / 255.  vidc_to_tensor_u8 does that on the device and vidc_warp_unwarp_host_u8 takes the
uint8 frames as host buffers; both must reproduce ToTensor (executed here by torch / torchvision on the CPU) bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

from tests import common as C

pytestmark = pytest.mark.gpu


def _u8(shape, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, size=shape, dtype=np.uint8)
    n = min(256, x.size)
    x.reshape(-1)[:n] = np.arange(n, dtype=np.uint8)                # every value at least once (where it fits)
    return x


@pytest.mark.parametrize("shape", [(3, 480, 640, 3), (5, 240, 320, 3), (2, 13, 17, 3), (2, 31, 29, 1), (3, 24, 32, 4), (1, 7, 5, 2)],
                         ids=lambda s: "x".join(map(str, s)))
def test_to_tensor_matches_torchvision(cuda_device, shape):
    from PIL import Image
    from torchvision import transforms
    from vi_depth_completion_b200.gravity import to_tensor_u8
    x = _u8(shape, seed=sum(shape))
    got = to_tensor_u8(torch.from_numpy(x).to(cuda_device))
    want = torch.from_numpy(x).permute(0, 3, 1, 2).to(torch.float32).div(255)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert C.count_bit_mismatches(got.cpu().numpy(), want.numpy()) == 0
    if shape[3] == 3:                                                   # the reference's own call: to_tensor(PIL image)
        tv = transforms.ToTensor()(Image.fromarray(x[0]))
        assert torch.equal(got[0].cpu(), tv)
        single = to_tensor_u8(torch.from_numpy(x[0]).to(cuda_device))   # (H,W,C) -> (C,H,W)
        assert torch.equal(single.cpu(), tv)


def test_to_tensor_unaligned_view_and_errors(cuda_device):
    from vi_depth_completion_b200.gravity import to_tensor_u8
    base = torch.from_numpy(_u8((1 + 2 * 24 * 32 * 3,), seed=3)).to(cuda_device)
    x = base[1:].view(2, 24, 32, 3)                                     # data pointer off by one byte: the generic kernel
    want = x.cpu().permute(0, 3, 1, 2).to(torch.float32).div(255)
    assert torch.equal(to_tensor_u8(x).cpu(), want)
    assert to_tensor_u8(torch.empty((0, 4, 4, 3), dtype=torch.uint8, device=cuda_device)).shape == (0, 3, 4, 4)
    with pytest.raises(RuntimeError):
        to_tensor_u8(torch.zeros((2, 4, 4, 3), device=cuda_device))     # float input
    with pytest.raises(RuntimeError):
        to_tensor_u8(torch.zeros((2, 4, 4, 5), dtype=torch.uint8, device=cuda_device))
    with pytest.raises(RuntimeError):
        to_tensor_u8(torch.zeros((2, 4, 4, 3), dtype=torch.uint8))      # host tensor


@pytest.mark.parametrize("cam_name,B", [("S1", 37), ("S2", 5)])
def test_host_entry_with_uint8_frames(cuda_device, oracle_mod, cam_name, B):
    """vidc_warp_unwarp_host_u8: same pipeline, RGB crosses PCIe as (B,H,W,3) uint8.  Bit-identical to the float entry on
    ToTensor(frames) and to the oracle on the same floats."""
    from vi_depth_completion_b200 import _cabi
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    from oracle import oracle as O
    cam = C.CAMERAS[cam_name]
    w, o = Warping2DOFAlignment(*cam), O.Oracle(*cam)
    I_g, I_a = C.random_gravity(B, seed=5)
    u8 = _u8((B, o.H, o.W, 3), seed=21)
    rgb = torch.from_numpy(u8).permute(0, 3, 1, 2).to(torch.float32).div(255).contiguous().numpy()
    _, depth, normals = C.random_images(B, o.H, o.W, seed=12)
    o_rgb, o_depth, o_mask, o_n = oracle_mod.warp_unwarp_mt(o, rgb, depth, normals, I_g, I_a, 4)
    mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    stream = ctypes.c_void_p(torch.cuda.current_stream(cuda_device).cuda_stream)
    results = []
    for entry, first in (("vidc_warp_unwarp_host_u8", u8), ("vidc_warp_unwarp_host", rgb)):
        h = [mk(x) for x in (first, depth, normals, I_g, I_a)]
        outs = [torch.zeros(B, 3, o.H, o.W).pin_memory(), torch.zeros(B, o.H, o.W).pin_memory(),
                torch.zeros(B, 1, o.H, o.W, dtype=torch.uint8).pin_memory(), torch.zeros(B, 3, o.H, o.W).pin_memory()]
        with torch.cuda.device(cuda_device):
            _cabi.check(getattr(_cabi.lib(), entry)(ctypes.byref(w._cam), B, *[t.data_ptr() for t in h], *[t.data_ptr() for t in outs], stream))
        results.append([t.numpy().copy() for t in outs])
    for a, b in zip(*results):
        assert C.count_bit_mismatches(a, b) == 0
    assert C.count_bit_mismatches(results[0][0], o_rgb) == 0
    assert C.count_bit_mismatches(results[0][1], o_depth) == 0
    assert np.array_equal(results[0][2], o_mask)
    assert C.count_bit_mismatches(results[0][3], o_n) == 0
    with torch.cuda.device(cuda_device):                                # null frames: an error, not a crash
        rc = _cabi.lib().vidc_warp_unwarp_host_u8(ctypes.byref(w._cam), B, None, None, None, None, None, None, None, None, None, stream)
    assert rc != 0
    assert _cabi.lib().vidc_release_workspace() == 0
