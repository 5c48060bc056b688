"""Host-buffer entry point timing for different pipeline chunk sizes (set VIDC_E2E_CHUNK before the process starts)."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import common as C
from vi_depth_completion_b200 import _cabi
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
w = Warping2DOFAlignment(*C.CAMERAS["S2"]); H, W = int(w.H), int(w.W); B = 256
I_g, I_a = C.random_gravity(B, 1234)
pin = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, pin_memory=True)
h_rgb, h_d, h_n = pin(B, 3, H, W).uniform_(), pin(B, 1, H, W).uniform_(), pin(B, 3, H, W).normal_()
h_g, h_a = torch.from_numpy(I_g).pin_memory(), torch.from_numpy(I_a).pin_memory()
o_rgb, o_d, o_m, o_n = pin(B, 3, H, W), pin(B, 1, H, W), pin(B, 1, H, W, dt=torch.uint8), pin(B, 3, H, W)
lib = _cabi.lib(); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def step():
    _cabi.check(lib.vidc_warp_unwarp_host(ctypes.byref(w._cam), B, h_rgb.data_ptr(), h_d.data_ptr(), h_n.data_ptr(), h_g.data_ptr(), h_a.data_ptr(),
                                          o_rgb.data_ptr(), o_d.data_ptr(), o_m.data_ptr(), o_n.data_ptr(), st))
for _ in range(2): step()
t0 = time.perf_counter()
for _ in range(8): step()
dt = (time.perf_counter() - t0) / 8
print(f"chunk {os.environ.get('VIDC_E2E_CHUNK','16')}: {dt*1e3:.1f} ms/step -> {B/dt:.0f} frames/s, H2D+D2H {(2.202+2.281)/dt:.1f} GB/s")
