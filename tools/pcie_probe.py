"""PCIe ceiling for the host-buffer entry point: pinned H2D and D2H copies alone and both directions at once."""
import torch, time
dev = torch.device("cuda", 0)
n = 1 << 30                                            # 1 GiB per direction
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, chunk, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for o in range(0, n, chunk):
            if h2d:
                with torch.cuda.stream(s1): d_in[o:o + chunk].copy_(h_in[o:o + chunk], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out[o:o + chunk].copy_(d_out[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
for chunk in (1 << 30, 64 << 20, 16 << 20):
    print(f"chunk {chunk >> 20:5d} MiB: H2D alone {run(True, False, chunk):5.1f} GB/s, D2H alone {run(False, True, chunk):5.1f} GB/s, "
          f"both at once {run(True, True, chunk):5.1f} GB/s per direction", flush=True)
