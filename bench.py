#!/usr/bin/env python
"""bench.py -- gravity warp+unwarp frames/sec at 640x480 on N B200s, with HBM-roofline accounting.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frames held by this rank:
    frame parameters (1 thread / frame) -> fused forward warp of RGB + sparse depth + validity mask
    -> fused inverse warp of the normals with R^T rotation and renormalisation.
Workload (BASELINE.json configs[2], the configuration the metric is quoted on): Azure-Kinect-shaped
640x480 camera, 256 frames per GPU, uniform roll/pitch in +-30 deg, I_a = [0,1,0], seeded synthetic
RGB / dense depth / normals.  Frames are independent: every rank owns its own 256 frames (weak scaling),
there is no data-path collective, NCCL is used for the barrier and the max-over-ranks of the timing only.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the same metric
through the C-ABI host-buffer entry point (pinned host buffers, H2D + kernels + D2H inside the timed
region); `roofline` = dominant kernel against the measured HBM copy bandwidth; `cpu_baseline` = the CPU
oracle (a port of the reference's algorithm, bit-identical to the reference run on CPU) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import common as C  # noqa: E402  (seeded synthetic inputs shared with the tests)

WORKLOAD = "S2"          # 640x480 Azure-Kinect-shaped camera
FRAMES_PER_GPU = 256
BYTES_PER_PX_FWD = 12 + 4 + 12 + 4 + 1   # read RGB+depth, write RGB+depth, write u8 mask
BYTES_PER_PX_INV = 12 + 12               # read normals, write rotated+normalised normals
METRIC = "gravity warp+unwarp frames/sec at 640x480"
UNIT = "frames/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM traffic of the dominant kernel from the committed ncu capture, else None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


def make_inputs(rank, B):
    cam = C.CAMERAS[WORKLOAD]
    I_g, I_a = C.random_gravity(B, seed=1234 + rank, roll_deg=30.0, pitch_deg=30.0)   # sharding.rank_seed(1234, rank)
    return cam, I_g, I_a


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port -- the Python
    reference itself cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample = min(256, max(4 * cores, 32))
    cam, I_g, I_a = make_inputs(0, sample)
    orc = O.Oracle(*cam)
    rgb, depth, normals = C.random_images(sample, orc.H, orc.W, seed=1)
    for _ in range(min(args.warmup, 1)):
        O.warp_unwarp_mt(orc, rgb, depth, normals, I_g, I_a, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.warp_unwarp_mt(orc, rgb, depth, normals, I_g, I_a, cores)
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: 640x480, {sample} frames per step (bounded sample of the 256-frame batch), "
                               "warp RGB + warp depth + mask + unwarp normals + renormalise"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames x {args.steps} steps, oracle/warp_oracle.c, {cores} pthreads"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vi_depth_completion_b200 import _cabi, sharding
    from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = FRAMES_PER_GPU
    cam, I_g, I_a = make_inputs(rank, B)
    w = Warping2DOFAlignment(*cam)
    H, W = int(w.H), int(w.W)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    rgb = torch.rand(B, 3, H, W, device=dev, generator=gen)
    depth = torch.rand(B, 1, H, W, device=dev, generator=gen) * 9.6 + 0.4
    normals = torch.randn(B, 3, H, W, device=dev, generator=gen)
    g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)

    lib = _cabi.lib()

    def step():
        w.warp_rgbd(rgb, depth, g, a)          # params + fused forward (RGB, depth, mask)
        w.unwarp_normals(normals, g, a)        # params + fused inverse (gather, R^T, normalise)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: EXACTLY K steps, CUDA events on the launching (current) stream ----------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    launches0 = lib.vidc_launch_count()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for k in range(K):
        ev[k][0].record()
        w.warp_rgbd(rgb, depth, g, a)
        ev[k][1].record()
        w.unwarp_normals(normals, g, a)
        ev[k][2].record()
    e_end.record()
    barrier()
    launches = lib.vidc_launch_count() - launches0
    n_clock_rows_timed = len(sampler.rows) if rank == 0 else 0    # samples taken while the K timed steps ran
    elapsed_ms = e_start.elapsed_time(e_end)
    fwd_ms = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]))
    inv_ms = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(K)]))
    step_ms = [ev[k][0].elapsed_time(ev[k][2]) for k in range(K)]       # report median and min too
    # frames are independent: aggregate = sum(frames) / max(elapsed) over ranks, no data-path collective
    _, elapsed_ms, value = sharding.aggregate_throughput(B * K, elapsed_ms, dev)

    # ---- informational: the opt-in packed layout (channels-last RGBD, one 128-bit load per tap), same frames ----
    packed = torch.cat([rgb, depth], 1).contiguous(memory_format=torch.channels_last)
    for _ in range(3):
        w.warp_rgbd_packed(packed, g, a)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(K):
        w.warp_rgbd_packed(packed, g, a)
    p1.record()
    torch.cuda.synchronize()
    packed_ms = p0.elapsed_time(p1) / K
    del packed

    # ---- e2e: C-ABI host-buffer entry point, pinned host memory, H2D + kernels + D2H timed --------
    hw = H * W
    h_rgb = torch.empty(B, 3, H, W, pin_memory=True); h_rgb.copy_(rgb)
    h_depth = torch.empty(B, 1, H, W, pin_memory=True); h_depth.copy_(depth)
    h_nrm = torch.empty(B, 3, H, W, pin_memory=True); h_nrm.copy_(normals)
    h_g = torch.from_numpy(I_g).pin_memory(); h_a = torch.from_numpy(I_a).pin_memory()
    o_rgb = torch.empty(B, 3, H, W, pin_memory=True); o_depth = torch.empty(B, 1, H, W, pin_memory=True)
    o_mask = torch.empty(B, 1, H, W, dtype=torch.uint8, pin_memory=True); o_nrm = torch.empty(B, 3, H, W, pin_memory=True)
    h2d = (h_rgb.numel() + h_depth.numel() + h_nrm.numel() + h_g.numel() + h_a.numel()) * 4
    d2h = (o_rgb.numel() + o_depth.numel() + o_nrm.numel()) * 4 + o_mask.numel()
    stream = torch.cuda.current_stream(dev).cuda_stream

    def e2e_step():
        _cabi.check(lib.vidc_warp_unwarp_host(ctypes.byref(w._cam), B, h_rgb.data_ptr(), h_depth.data_ptr(), h_nrm.data_ptr(),
                                              h_g.data_ptr(), h_a.data_ptr(), o_rgb.data_ptr(), o_depth.data_ptr(),
                                              o_mask.data_ptr(), o_nrm.data_ptr(), ctypes.c_void_p(stream)))

    Ke = max(3, min(K, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()                  # synchronises its stream before returning
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    _, _, e2e_value = sharding.aggregate_throughput(B * Ke, e2e_s * 1e3, dev)
    clocks = sampler.stop() if rank == 0 else None             # window: timed steps + packed kernel + e2e steps
    if clocks is not None:
        clocks["samples_during_timed_steps"] = n_clock_rows_timed
    # sanity: the e2e path returns the same bits as the resident path
    _, rgb_w, depth_w, mask = w.warp_rgbd(rgb, depth, g, a)
    torch.cuda.synchronize()
    e2e_ok = bool(torch.equal(rgb_w.cpu(), o_rgb) and torch.equal(mask.cpu(), o_mask))

    # ---- cpu baseline (rank 0, N == 1): the oracle port on the host cores, bounded sample ----------
    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle as O
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        sample = max(2 * cores, 32)
        orc = O.Oracle(*cam)
        s_rgb = h_rgb[:sample].numpy(); s_d = h_depth[:sample].numpy(); s_n = h_nrm[:sample].numpy()
        O.warp_unwarp_mt(orc, s_rgb[:cores], s_d[:cores], s_n[:cores], I_g[:cores], I_a[:cores], cores)  # warm
        reps, t0 = 0, time.perf_counter()
        while True:
            ref = O.warp_unwarp_mt(orc, s_rgb, s_d, s_n, I_g[:sample], I_a[:sample], cores)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 50:
                break
        dt = time.perf_counter() - t0
        same = bool(np.array_equal(ref[0], o_rgb[:sample].numpy()) and np.array_equal(ref[2], o_mask[:sample].numpy())
                    and np.array_equal(ref[3], o_nrm[:sample].numpy()) and np.array_equal(ref[1], o_depth[:sample, 0].numpy()))
        cpu = {"value": sample * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} frames x {reps} repetitions of the same step on oracle/warp_oracle.c ({cores} pthreads)",
               "gpu_output_bit_identical_on_sample": same}

    if rank == 0:
        peak, peak_src = measured_peak()
        px = B * H * W
        # kernel families behind the fused entry points (vidc_kernels.cu: shear_level(), VIDC_SHEAR, default 2)
        shear = os.environ.get("VIDC_SHEAR", "2")[:1]
        fwd_name = "warp_rgbd_fast_kernel" if shear == "0" else "warp_rgbd_shear_kernel"
        inv_name = "unwarp_normals_shear_kernel" if shear == "2" else "unwarp_normals_fast_kernel"
        dom = (fwd_name, fwd_ms, BYTES_PER_PX_FWD) if fwd_ms >= inv_ms else (inv_name, inv_ms, BYTES_PER_PX_INV)
        achieved = px * dom[2] / (dom[1] * 1e-3) / 1e9
        traffic = ncu_traffic()
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "peak_source": peak_src,
                    "traffic": ((traffic or {}).get(dom[0]) or {}).get("dram_bytes_total"),
                    "traffic_source": ((traffic or {}).get(dom[0]) or {}).get("source"),
                    "algorithmic_bytes_per_launch": px * dom[2],
                    "kernels": {fwd_name: {"ms": fwd_ms, "GBps": px * BYTES_PER_PX_FWD / fwd_ms / 1e6},
                                inv_name: {"ms": inv_ms, "GBps": px * BYTES_PER_PX_INV / inv_ms / 1e6},
                                "warp_rgbd_nhwc4_kernel (opt-in packed RGBD layout, not part of `value`)": {
                                    "ms": packed_ms, "GBps": px * BYTES_PER_PX_FWD / packed_ms / 1e6,
                                    "frac": px * BYTES_PER_PX_FWD / packed_ms / 1e6 / peak}},
                    "step": {"bytes_per_frame": H * W * (BYTES_PER_PX_FWD + BYTES_PER_PX_INV),
                             "achieved": value / world * H * W * (BYTES_PER_PX_FWD + BYTES_PER_PX_INV) / 1e9,
                             "frac": value / world * H * W * (BYTES_PER_PX_FWD + BYTES_PER_PX_INV) / 1e9 / peak}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / K, "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_min": float(np.min(step_ms)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD}: Azure-Kinect-shaped 640x480 (fx=fy=404), {B} frames per GPU, roll/pitch U(-30,30) deg, "
                                   "RGB+depth forward warp + mask, normals inverse warp + R^T + renormalise",
                       "frames_per_gpu": B, "l2": "inputs (2.2 GB per step) larger than L2, no flush needed",
                       "layout": "NCHW fp32", "sharding": f"batch over {world} GPU(s), no data-path collective"},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "api": "vidc_warp_unwarp_host (C ABI, pinned host buffers)", "matches_resident_path": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
