"""Development helper: the sparse-depth forward warp, dense route (rasterise + warp_rgbd) against the analytic route
(warp_rgb_sparse_depth), S2 workload, 150 points per frame like the demo's 89-172."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vi_depth_completion_b200 import synthetic as S
from vi_depth_completion_b200.gravity import rasterize_sparse_depth
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*S.CAMERAS["S2"]); B = 256; H, W = int(w.H), int(w.W)
I_g, I_a = S.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
rgb = torch.rand(B, 3, H, W, device=dev)
rs = np.random.RandomState(0); N = 150
tr = np.zeros((B, N, 5)); z = rs.uniform(0.4, 6, (B, N)); fc, cc = (404.0, 404.0), (319.9, 239.9)
tr[..., 1] = (rs.uniform(0, W, (B, N)) - cc[0]) / fc[0] * z; tr[..., 2] = (rs.uniform(0, H, (B, N)) - cc[1]) / fc[1] * z; tr[..., 3] = z
trd = torch.from_numpy(tr).to(dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
dense = rasterize_sparse_depth(trd, None, fc, cc, H, W)
res = {"rasterise_ms": t(lambda: rasterize_sparse_depth(trd, None, fc, cc, H, W)),
       "warp_rgbd_dense_ms": t(lambda: w.warp_rgbd(rgb, dense, g, a)),
       "warp_rgb_sparse_depth_ms": t(lambda: w.warp_rgb_sparse_depth(rgb, trd, None, fc, cc, g, a)),
       "warp_rgb_only_ms": t(lambda: w.warp_rgbd(rgb, None, g, a))}
res["dense_route_ms"] = res["rasterise_ms"] + res["warp_rgbd_dense_ms"]
res["speedup"] = res["dense_route_ms"] / res["warp_rgb_sparse_depth_ms"]
print(json.dumps(res))
