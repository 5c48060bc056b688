"""Small-batch (launch-bound) timing with and without CUDA-graph replay (BASELINE config S1: 320x240, B=64)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
for cam, B in (("S1", 64), ("S1", 8), ("S2", 16)):
    w = Warping2DOFAlignment(*C.CAMERAS[cam]); H, W = int(w.H), int(w.W)
    I_g, I_a = C.random_gravity(B, 1234)
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    rgb = torch.rand(B, 3, H, W, device=dev); depth = torch.rand(B, 1, H, W, device=dev); nrm = torch.randn(B, 3, H, W, device=dev)
    def step():
        w.warp_rgbd(rgb, depth, g, a); w.unwarp_normals(nrm, g, a)
    for _ in range(5): step()
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n): step()
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / n
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): step()
    for _ in range(5): gr.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): gr.replay()
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / n
    print(f"{cam} B={B}: eager {eager*1e6:.1f} us/step ({B/eager:.0f} frames/s), graph replay {graph*1e6:.1f} us/step ({B/graph:.0f} frames/s)")
