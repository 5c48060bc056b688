"""Aggregate PCIe ceiling of the box with every GPU busy: one rank per GPU (torchrun), pinned H2D and D2H copies in both
directions at once on all ranks, barrier-aligned; rank 0 prints per-rank and aggregate GB/s.  The host-buffer entry point
(bench.py `e2e`) moves 2.20 GB in and 2.28 GB out per 256 frames and rank."""
import os, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
chunk = 64 << 20
def run(h2d, d2h, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for o in range(0, n, chunk):
            if h2d:
                with torch.cuda.stream(s1): d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
res = torch.tensor([run(True, False), run(False, True), run(True, True)], device=dev, dtype=torch.float64)
if world > 1:
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    m = torch.stack(allr).cpu().numpy()
    print(json.dumps({"gpus": world, "h2d_alone_GBps_per_rank": [round(float(v), 1) for v in m[:, 0]],
                      "d2h_alone_GBps_per_rank": [round(float(v), 1) for v in m[:, 1]],
                      "both_GBps_per_direction_per_rank": [round(float(v), 1) for v in m[:, 2]],
                      "aggregate_h2d_alone": round(float(m[:, 0].sum()), 1), "aggregate_d2h_alone": round(float(m[:, 1].sum()), 1),
                      "aggregate_both_per_direction": round(float(m[:, 2].sum()), 1)}), flush=True)
if world > 1:
    dist.destroy_process_group()
