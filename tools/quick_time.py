"""Quick device-side timing of the fused kernels (development helper, not the bench contract)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment

def main(cam="S2", B=256, iters=20, roll=30, pitch=30, cl=False):
    dev = torch.device("cuda", 0)
    fx, fy, cx, cy = C.CAMERAS[cam]
    w = Warping2DOFAlignment(fx, fy, cx, cy)
    H, W = int(w.H), int(w.W)
    I_g, I_a = C.random_gravity(B, 1234, roll, pitch)
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    rgb = torch.rand(B, 3, H, W, device=dev, generator=gen)
    depth = torch.rand(B, 1, H, W, device=dev, generator=gen) * 9.6 + 0.4
    nrm = torch.randn(B, 3, H, W, device=dev, generator=gen)
    if cl:
        rgb = rgb.contiguous(memory_format=torch.channels_last); nrm = nrm.contiguous(memory_format=torch.channels_last)
    def fwd(): return w.warp_rgbd(rgb, depth, g, a)
    def inv(): return w.unwarp_normals(nrm, g, a)
    res = {}
    for name, fn in (("forward_rgbd_mask", fwd), ("inverse_rot_norm", inv)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        bpp = 33 if name.startswith("forward") else 24
        res[name] = {"ms": ms, "GBps": B * H * W * bpp / ms / 1e6}
    tot = res["forward_rgbd_mask"]["ms"] + res["inverse_rot_norm"]["ms"]
    res["frames_per_s"] = B / tot * 1e3
    res["frac_of_6546.6"] = B * H * W * 57 / tot / 1e6 / 6546.6
    print(json.dumps({"cam": cam, "B": B, "cl": cl, "roll": roll, **res}), flush=True)

if __name__ == "__main__":
    main()
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        sys.exit(0)
    main(cam="S3", roll=90, pitch=5)
    main(cam="S1", B=64)
    main(cl=True)
