"""Device time of vidc_to_tensor_u8 (HWC uint8 -> CHW float, 15 B/px) next to torch's permute + float + div."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vi_depth_completion_b200.gravity import to_tensor_u8
dev = torch.device("cuda", 0)
B, H, W = 256, 480, 640
x = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
ms = t(lambda: to_tensor_u8(x)); ms_t = t(lambda: x.permute(0, 3, 1, 2).to(torch.float32).div(255).contiguous())
print(json.dumps({"to_tensor_u8_ms": round(ms, 4), "GBps": round(B * H * W * 15 / ms / 1e6, 1), "torch_ms": round(ms_t, 4)}))
