"""Development helper: inverse-warp timing with / without renormalisation (ablation runs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS["S2"]); B = 256; H, W = int(w.H), int(w.W)
I_g, I_a = C.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
nrm = torch.randn(B, 3, H, W, device=dev)
for name, fn in (("inv+norm", lambda: w.unwarp_normals(nrm, g, a)), ("inv", lambda: w.unwarp_normals(nrm, g, a, normalize=False))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 20, 4), "ms", flush=True)
