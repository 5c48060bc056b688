"""Device time of the REFERENCE-SHAPED call sequence (surface_normal.py:148-170 unmodified on top of the drop-in class):
warp RGB, warp depth, mask, inverse warp + rotation, F.normalize -- next to the fused entry points."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
from vi_depth_completion_b200 import normal_utils as NU
dev = torch.device("cuda", 0)
for cam, B in (("S2", 256), ("S1", 64)):
    w = Warping2DOFAlignment(*C.CAMERAS[cam]); H, W = int(w.H), int(w.W)
    I_g, I_a = C.random_gravity(B, 1234)
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
    rgb = torch.rand(B, 3, H, W, device=dev); depth = torch.rand(B, H, W, device=dev); nrm = torch.randn(B, 3, H, W, device=dev)
    def t(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    res = {
        "warp_rgb": t(lambda: w.warp_with_gravity_center_aligned(rgb, g, a)),
        "warp_depth_3d": t(lambda: w.warp_with_gravity_center_aligned(depth, g, a)),
        "mask": None, "inverse": t(lambda: w.inverse_warp_normal_image_with_gravity_center_aligned(nrm, g, a)),
    }
    x1 = w.warp_with_gravity_center_aligned(rgb, g, a)[1]
    z = w.inverse_warp_normal_image_with_gravity_center_aligned(nrm, g, a)[1]
    res["mask"] = t(lambda: NU.validity_mask(x1))
    res["mask_torch_as_reference"] = t(lambda: ((x1[:, 0:1] + x1[:, 1:2] + x1[:, 2:3]) > 1e-2).float())
    res["normalize"] = t(lambda: NU.Normalize(z))
    res["normalize_torch_as_reference"] = t(lambda: torch.nn.functional.normalize(z, dim=1))
    res["fused_warp_rgbd"] = t(lambda: w.warp_rgbd(rgb, depth, g, a))
    res["fused_unwarp_normals"] = t(lambda: w.unwarp_normals(nrm, g, a))
    seq = res["warp_rgb"] + res["warp_depth_3d"] + res["mask_torch_as_reference"] + res["inverse"] + res["normalize_torch_as_reference"]
    fused = res["fused_warp_rgbd"] + res["fused_unwarp_normals"]
    print(json.dumps({"cam": cam, "B": B, **{k: round(v, 4) for k, v in res.items()},
                      "reference_shaped_total_ms": round(seq, 4), "frames_per_s_reference_shaped": round(B / seq * 1e3),
                      "fused_total_ms": round(fused, 4), "frames_per_s_fused": round(B / fused * 1e3)}), flush=True)
