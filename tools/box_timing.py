"""Development helper: per-phase clock64 sums of unwarp_normals_box_kernel (library built with -DVIDC_BOX_TIMING)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200 import _cabi
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS["S2"]); B = 256; H, W = int(w.H), int(w.W)
I_g, I_a = C.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
nrm = torch.randn(B, 3, H, W, device=dev)
l = ctypes.CDLL(_cabi.LIB_PATH)
buf = (ctypes.c_ulonglong * 8)()
for _ in range(3): w.unwarp_normals(nrm, g, a)
l.vidc_debug_box_timing(buf, 1)
N = 10
for _ in range(N): w.unwarp_normals(nrm, g, a)
l.vidc_debug_box_timing(buf, 1)
n = buf[7]
names = ["A: prologue+entry", "A: staging+barrier", "A: compute+store", "B: tail+entry", "B: staging+barrier", "B: compute+store"]
for i, nm in enumerate(names):
    print(f"{nm:22s} {buf[i] / n:9.0f} cycles avg per CTA")
print("CTAs", n // N, "total avg", sum(buf[:6]) / n)
