"""One forward + one inverse launch at the bench workload (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
cam = sys.argv[1] if len(sys.argv) > 1 else "S2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS[cam])
H, W = int(w.H), int(w.W)
I_g, I_a = C.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
gen = torch.Generator(device=dev).manual_seed(1)
rgb = torch.rand(B, 3, H, W, device=dev, generator=gen)
depth = torch.rand(B, 1, H, W, device=dev, generator=gen) * 9.6 + 0.4
nrm = torch.randn(B, 3, H, W, device=dev, generator=gen)
for _ in range(reps):
    w.warp_rgbd(rgb, depth, g, a)
    w.unwarp_normals(nrm, g, a)
torch.cuda.synchronize()
