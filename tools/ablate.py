"""Sensitivity of the forward kernel to its outputs / inputs (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS["S2"]); H, W = int(w.H), int(w.W); B = 256
I_g, I_a = C.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
rgb = torch.rand(B, 3, H, W, device=dev, generator=gen); depth = torch.rand(B, 1, H, W, device=dev, generator=gen) * 9.6 + 0.4
nrm = torch.randn(B, 3, H, W, device=dev, generator=gen)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("rgbd+mask      %.3f" % t(lambda: w.warp_rgbd(rgb, depth, g, a)))
print("rgbd no mask   %.3f" % t(lambda: w.warp_rgbd(rgb, depth, g, a, with_mask=False)))
print("rgb  +mask     %.3f" % t(lambda: w.warp_rgbd(rgb, None, g, a)))
print("rgb  no mask   %.3f" % t(lambda: w.warp_rgbd(rgb, None, g, a, with_mask=False)))
print("rgbd+mask+cov  %.3f" % t(lambda: w.warp_rgbd(rgb, depth, g, a, with_coverage=True)))
print("unwarp norm    %.3f" % t(lambda: w.unwarp_normals(nrm, g, a)))
print("unwarp raw     %.3f" % t(lambda: w.unwarp_normals(nrm, g, a, normalize=False)))
print("unwarp+valid   %.3f" % t(lambda: w.unwarp_normals(nrm, g, a, with_valid=True)))
x = torch.empty_like(rgb); 
print("copy rgb (torch) %.3f  -> GB/s %.0f" % ((lambda ms: (ms, 2*rgb.numel()*4/ms/1e6))(t(lambda: x.copy_(rgb)))))
zero_g = torch.zeros_like(g); zero_g[:, 1] = 1
print("identity frames rgbd+mask %.3f" % t(lambda: w.warp_rgbd(rgb, depth, zero_g, a)))
print("identity frames unwarp    %.3f" % t(lambda: w.unwarp_normals(nrm, zero_g, a)))
xp = torch.cat([rgb, depth], 1).contiguous(memory_format=torch.channels_last)
ms = t(lambda: w.warp_rgbd_packed(xp, g, a))
print("packed NHWC4 rgbd+mask %.3f  -> %.0f GB/s algorithmic (33 B/px)" % (ms, B * H * W * 33 / ms / 1e6))
