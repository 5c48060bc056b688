"""B200-native gravity warp / unwarp path of MARSLab-UMN/vi_depth_completion.

    from vi_depth_completion_b200 import Warping2DOFAlignment       # drop-in for networks/warping_2dof_alignment.py
    from vi_depth_completion_b200 import normal_utils               # drop-in for normal_utils.py

Everything runs on hand-written sm_100a kernels behind the C ABI in include/vidc_b200.h
(libvidc_b200.so, built in-tree by `python -m vi_depth_completion_b200.build`).  There is no CPU fallback.
"""
from .warping_2dof_alignment import Warping2DOFAlignment  # noqa: F401

__all__ = ["Warping2DOFAlignment"]
__version__ = "0.1.0"
