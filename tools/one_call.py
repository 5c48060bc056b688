import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests import common as C
from vi_depth_completion_b200.warping_2dof_alignment import Warping2DOFAlignment
cam = sys.argv[1] if len(sys.argv) > 1 else "S1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
w = Warping2DOFAlignment(*C.CAMERAS[cam]); H, W = int(w.H), int(w.W)
I_g, I_a = C.random_gravity(B, 1234)
rgb, depth, nrm = C.random_images(B, H, W, 1)
t = lambda x: torch.from_numpy(x).to(dev)
out = w.warp_rgbd(t(rgb), t(depth), t(I_g), t(I_a))
torch.cuda.synchronize()
print("ok", float(out[1].sum()))
