"""Other canvas sizes (runtime-geometry kernels): the real Azure Kinect canvas 640x489 and 1280x960."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
for cam in ((404.0, 404.0, 319.87654, 244.3), (808.0, 808.0, 639.9, 479.9)):
    w = Warping2DOFAlignment(*cam); H, W = int(w.H), int(w.W); B = 256 if W < 1000 else 64
    I_g, I_a = C.random_gravity(B, 1234); g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    rgb = torch.rand(B, 3, H, W, device=dev); d = torch.rand(B, 1, H, W, device=dev); n = torch.randn(B, 3, H, W, device=dev)
    def t(fn, k=10):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / k
    f = t(lambda: w.warp_rgbd(rgb, d, g, a)); i = t(lambda: w.unwarp_normals(n, g, a))
    px = B * H * W
    print(f"{W}x{H} B={B} VIDC_SHEAR={os.environ.get('VIDC_SHEAR', '2')}: forward {f:.3f} ms ({px*33/f/1e6:.0f} GB/s) "
          f"inverse {i:.3f} ms ({px*24/i/1e6:.0f} GB/s) -> {B/(f+i)*1e3:.0f} frames/s")
    del rgb, d, n
