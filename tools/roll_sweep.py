"""Sensitivity of the kernels to the roll angle (BASELINE config S3: ScanNet intrinsics, 640x480, B=256)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS["S3"]); H, W = int(w.H), int(w.W); B = 256
gen = torch.Generator(device=dev).manual_seed(1)
rgb = torch.rand(B, 3, H, W, device=dev, generator=gen); depth = torch.rand(B, 1, H, W, device=dev, generator=gen)
nrm = torch.randn(B, 3, H, W, device=dev, generator=gen)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for roll in (0, 15, 30, 45, 60, 75, 90, -90):
    rs = np.random.RandomState(0)
    r = np.deg2rad(roll + rs.rand(B).astype(np.float32) * 4 - 2).astype(np.float32)
    I_g = C.gravity_from_angles(r, np.zeros(B, np.float32)); I_a = np.tile(np.array([[0, 1, 0]], np.float32), (B, 1))
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    f = t(lambda: w.warp_rgbd(rgb, depth, g, a)); i = t(lambda: w.unwarp_normals(nrm, g, a))
    valid = float(w.warp_rgbd(rgb, depth, g, a)[3].float().mean())
    print(f"roll {roll:4d}: forward {f:.3f} ms  inverse {i:.3f} ms  -> {B/(f+i)*1e3:8.0f} frames/s   valid fraction {valid:.3f}")
