#!/usr/bin/env python
"""Run the reference AS SHIPPED on cuda:0 (oracle/_ref via oracle/ref_loader.py) next to the drop-in and the reference's
own CPU backend, on the B200 box.  Writes gpurun_out/ref_cuda_probe.json (difference statistics, timings) and
gpurun_out/ref_cuda_dump_<cfg>.npz (the reference's CUDA-side parameters, grids and outputs for a few frames, so the
accumulate schemes of cuBLAS / ATen CUDA can be searched offline -- SURVEY 7.3 item 2, Appendix C).

    python tools/ref_cuda_probe.py [--time] [--dump]
"""
import argparse
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from tests import common as C  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402


def stats(a, b):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    return {"max_abs": float(np.nanmax(d)) if d.size else 0.0, "bit_mismatches": C.count_bit_mismatches(a, b), "n": int(a.size),
            "n_gt_1e-4": int((d > 1e-4).sum()), "n_gt_1e-5": int((d > 1e-5).sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--dump", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    import torch
    assert RL.reference_available(), "oracle/_ref missing: run python -m oracle.fetch_ref in the build container"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
    RefCuda = RL.load_reference_class("cuda:0")
    RefCpu = RL.load_reference_class("cpu")
    os.makedirs(args.out, exist_ok=True)
    report = {"torch": torch.__version__, "device": torch.cuda.get_device_name(0), "configs": {}}

    cfgs = [("S1", 8, "random"), ("S2", 6, "random"), ("S3", 7, "roll"), ("S1", 12, "edge")]
    for name, B, kind in cfgs:
        cam = C.CAMERAS[name]
        if kind == "random":
            I_g, I_a = C.random_gravity(B, seed=1234)
        elif kind == "roll":
            I_g, I_a = C.extreme_roll_gravity(B, seed=5)
        else:
            I_g, I_a = C.edge_case_gravity(); B = I_g.shape[0]
        ours, rc, rcpu = Warping2DOFAlignment(*cam), RefCuda(*cam), RefCpu(*cam)
        H, W = int(ours.H), int(ours.W)
        rgb, depth, normals = C.random_images(B, H, W, seed=1)
        smooth = C.smooth_images(B, H, W, seed=3)
        t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        g, a = t(I_g), t(I_a)
        gc, ac = torch.from_numpy(I_g), torch.from_numpy(I_a)
        out = {}
        with torch.no_grad():
            # parameters
            Hc, Rc, Hic = [v.cpu().numpy() for v in rc._build_homography(g, a)]
            Ho, Ro, Hio = [v.cpu().numpy() for v in ours._build_homography(g, a)]
            Hp, Rp, Hip = [v.numpy() for v in rcpu._build_homography(gc, ac)]
            out["H refcuda-vs-ours"] = stats(Hc, Ho); out["R refcuda-vs-ours"] = stats(Rc, Ro); out["Hinv refcuda-vs-ours"] = stats(Hic, Hio)
            out["H refcuda-vs-refcpu"] = stats(Hc, Hp); out["R refcuda-vs-refcpu"] = stats(Rc, Rp)
            # grids
            Rt_c, grid_c, igrid_c = [v.cpu().numpy() for v in rc.image_sampler_forward_inverse(g, a)]
            Rt_o, grid_o, igrid_o = [v.cpu().numpy() for v in ours.image_sampler_forward_inverse(g, a)]
            fin = np.isfinite(grid_c) & np.isfinite(grid_o)
            out["grid refcuda-vs-ours"] = stats(np.where(fin, grid_c, 0), np.where(fin, grid_o, 0))
            fin = np.isfinite(igrid_c) & np.isfinite(igrid_o)
            out["inv_grid refcuda-vs-ours"] = stats(np.where(fin, igrid_c, 0), np.where(fin, igrid_o, 0))
            # the reference-shaped call sequence of surface_normal.py:148-170
            res = {}
            for tag, wobj, conv in (("refcuda", rc, t), ("ours", ours, t), ("refcpu", rcpu, lambda x: torch.from_numpy(np.ascontiguousarray(x)))):
                gg, aa = conv(I_g), conv(I_a)
                _, x1 = wobj.warp_with_gravity_center_aligned(conv(rgb), gg, aa)
                _, xs = wobj.warp_with_gravity_center_aligned(conv(smooth), gg, aa)
                _, d1 = wobj.warp_with_gravity_center_aligned(conv(depth), gg, aa)
                _, dn = wobj.warp_with_gravity_center_aligned(conv(depth), gg, aa, interp_mode='nearest')
                mask = (x1[:, 0:1] + x1[:, 1:2] + x1[:, 2:3] > 1e-2)
                _, z = wobj.inverse_warp_normal_image_with_gravity_center_aligned(conv(normals), gg, aa)
                zn = torch.nn.functional.normalize(z, dim=1)
                res[tag] = {k: v.cpu().numpy() for k, v in dict(rgb=x1, smooth=xs, depth=d1, depth_nearest=dn, mask=mask, z=z, zn=zn).items()}
            for other in ("ours", "refcpu"):
                for k in ("rgb", "smooth", "depth", "depth_nearest", "z", "zn"):
                    A, Bm = res["refcuda"][k], res[other][k]
                    ok = np.isfinite(A) & np.isfinite(Bm)
                    out[f"{k} refcuda-vs-{other}"] = stats(np.where(ok, A, 0), np.where(ok, Bm, 0))
                out[f"mask refcuda-vs-{other}"] = {"mismatches": int((res["refcuda"]["mask"] != res[other]["mask"]).sum()),
                                                   "n": int(res[other]["mask"].size)}
                err, _ = C.angular_error_deg(res[other]["zn"], res["refcuda"]["zn"])
                err = err[np.isfinite(err)]
                out[f"angle_deg refcuda-vs-{other}"] = {"max": float(err.max()) if err.size else 0.0, "n_gt_0.01": int((err > 0.01).sum())}
            out["rgb ours-vs-refcpu"] = stats(res["ours"]["rgb"], res["refcpu"]["rgb"])
        report["configs"][f"{name}-{kind}"] = out
        print(name, kind, json.dumps(out, indent=1), flush=True)
        if args.dump:
            nd = 12 if kind == "edge" else (2 if name == "S1" else 1)
            np.savez_compressed(os.path.join(args.out, f"ref_cuda_dump_{name}_{kind}.npz"), cam=np.array(cam), I_g=I_g, I_a=I_a,
                                H=Hc, R=Rc, Hinv=Hic, Rt=Rt_c, grid=grid_c[:nd], inv_grid=igrid_c[:nd], H_cpu=Hp, R_cpu=Rp, Hinv_cpu=Hip,
                                rgb_w=res["refcuda"]["rgb"][:1], z=res["refcuda"]["z"][:1], zn=res["refcuda"]["zn"][:1])

    if args.time:
        # the honest "before" number (reference on one B200) and the reference's CPU path on the box's cores
        cam = C.CAMERAS["S2"]
        rc, rcpu, ours = RefCuda(*cam), RefCpu(*cam), Warping2DOFAlignment(*cam)
        H, W = int(ours.H), int(ours.W)

        def seq(wobj, rgb, depth, normals, g, a):
            _, x1 = wobj.warp_with_gravity_center_aligned(rgb, g, a)
            _, d1 = wobj.warp_with_gravity_center_aligned(depth, g, a)
            mask = (x1[:, 0:1] + x1[:, 1:2] + x1[:, 2:3] > 1e-2).float()
            _, z = wobj.inverse_warp_normal_image_with_gravity_center_aligned(normals, g, a)
            return torch.nn.functional.normalize(z, dim=1), mask, d1

        timing = {}
        for B in (32, 256):
            I_g, I_a = C.random_gravity(B, seed=1234)
            gen = torch.Generator(device=dev).manual_seed(1)
            rgb = torch.rand(B, 3, H, W, device=dev, generator=gen); depth = torch.rand(B, H, W, device=dev, generator=gen) * 9.6 + 0.4
            normals = torch.randn(B, 3, H, W, device=dev, generator=gen)
            g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
            with torch.no_grad():
                seq(rc, rgb[:2], depth[:2], normals[:2], g[:2], a[:2]); torch.cuda.synchronize()
                t0 = time.perf_counter(); seq(rc, rgb, depth, normals, g, a); torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                timing[f"reference_cuda_B{B}"] = {"seconds": dt, "frames_per_s": B / dt}
                for _ in range(3):
                    seq(ours, rgb, depth, normals, g, a)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(10):
                    seq(ours, rgb, depth, normals, g, a)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 10
                timing[f"dropin_same_sequence_B{B}"] = {"seconds": dt, "frames_per_s": B / dt}
            print(json.dumps(timing), flush=True)
        B = 32
        I_g, I_a = C.random_gravity(B, seed=1234)
        rgb, depth, normals = C.random_images(B, H, W, seed=1)
        tt = torch.from_numpy
        with torch.no_grad():
            seq(rcpu, tt(rgb[:2]), tt(depth[:2]), tt(normals[:2]), tt(I_g[:2]), tt(I_a[:2]))
            t0 = time.perf_counter(); seq(rcpu, tt(rgb), tt(depth), tt(normals), tt(I_g), tt(I_a)); dt = time.perf_counter() - t0
        timing["reference_cpu_B32"] = {"seconds": dt, "frames_per_s": B / dt, "torch_threads": torch.get_num_threads(), "cpu_count": os.cpu_count()}
        report["timing"] = timing
        print(json.dumps(timing), flush=True)
    with open(os.path.join(args.out, "ref_cuda_probe.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
