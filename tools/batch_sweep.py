"""Batch sweep of BASELINE config S3 (ScanNet intrinsics 640x480, roll/pitch U(-30,30) deg): device time per step with CUDA
events, eager host time per step, and CUDA-graph replay -- where the path turns from launch-bound to bandwidth-bound."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS["S3"]); H, W = int(w.H), int(w.W)
print(f"| B | device ms/step | frames/s (device) | eager host us/step | frames/s (eager) | graph replay us/step | frames/s (graph) |\n|---|---|---|---|---|---|---|")
for B in (1, 8, 32, 64, 128, 256, 512):
    I_g, I_a = C.random_gravity(B, 1234)
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    rgb = torch.rand(B, 3, H, W, device=dev); depth = torch.rand(B, 1, H, W, device=dev); nrm = torch.randn(B, 3, H, W, device=dev)
    def step():
        w.warp_rgbd(rgb, depth, g, a); w.unwarp_normals(nrm, g, a)
    n = 200 if B <= 64 else 40
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / n
    dev_ms = e0.elapsed_time(e1) / n
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): step()
    for _ in range(5): gr.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): gr.replay()
    torch.cuda.synchronize(); graph = (time.perf_counter() - t0) / n
    print(f"| {B} | {dev_ms:.4f} | {B/dev_ms*1e3:,.0f} | {eager*1e6:.1f} | {B/eager:,.0f} | {graph*1e6:.1f} | {B/graph:,.0f} |", flush=True)
    del rgb, depth, nrm, gr
    torch.cuda.empty_cache()
