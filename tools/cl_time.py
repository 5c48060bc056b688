"""Device time of the entry points on channels-last (NHWC-strided) tensors next to planar NCHW (256 frames of 640x480)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vi_depth_completion_b200 import synthetic as S
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
dev = torch.device("cuda", 0)
B = 256
w = Warping2DOFAlignment(*S.CAMERAS["S2"]); H, W = int(w.H), int(w.W)
I_g, I_a = S.random_gravity(B, 1234)
g, a = torch.from_numpy(I_g).to(dev), torch.from_numpy(I_a).to(dev)
rgb = torch.rand(B, 3, H, W, device=dev); nrm = torch.randn(B, 3, H, W, device=dev)
rgb_cl = rgb.contiguous(memory_format=torch.channels_last); nrm_cl = nrm.contiguous(memory_format=torch.channels_last)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return round(e0.elapsed_time(e1) / n, 4)
res = {"warp_rgb_nchw": t(lambda: w.warp_with_gravity_center_aligned(rgb, g, a)),
       "warp_rgb_channels_last": t(lambda: w.warp_with_gravity_center_aligned(rgb_cl, g, a)),
       "inverse_nchw": t(lambda: w.inverse_warp_normal_image_with_gravity_center_aligned(nrm, g, a)),
       "inverse_channels_last": t(lambda: w.inverse_warp_normal_image_with_gravity_center_aligned(nrm_cl, g, a)),
       "unwarp_normals_nchw": t(lambda: w.unwarp_normals(nrm, g, a)),
       "unwarp_normals_channels_last": t(lambda: w.unwarp_normals(nrm_cl, g, a)),
       "torch_contiguous_copy_of_normals": t(lambda: nrm_cl.contiguous())}
y = w.unwarp_normals(nrm_cl, g, a)[1]
res["output_is_channels_last"] = bool(y.is_contiguous(memory_format=torch.channels_last) and not y.is_contiguous())
res["same_bits"] = bool(torch.equal(y, w.unwarp_normals(nrm, g, a)[1]))
print(json.dumps(res))
