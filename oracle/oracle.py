"""ctypes front end of the CPU oracle (oracle/warp_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(vi_depth_completion_b200) never imports this module.

All arrays are C-contiguous numpy float32 in the reference's NCHW layout.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwarp_oracle.so")
_lib = None


_MODES = {"bilinear": 0, "nearest": 1, "bicubic": 2}      # vidc_oracle_grid_sample mode codes


class OracleCamera(ctypes.Structure):
    _fields_ = [
        ("W", ctypes.c_int32), ("H", ctypes.c_int32),
        ("K", ctypes.c_float * 9), ("Kinv", ctypes.c_float * 9),
        ("cx", ctypes.c_float), ("cy", ctypes.c_float),
        ("inv_half_w", ctypes.c_float), ("inv_half_h", ctypes.c_float),
        ("corners", ctypes.c_float * 12),
    ]


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "warp_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class Oracle:
    """CPU restatement of networks/warping_2dof_alignment.py:Warping2DOFAlignment."""

    def __init__(self, fx=577.87061 * 0.5, fy=577.87061 * 0.5, cx=319.87654 * 0.5, cy=239.87603 * 0.5):
        self.cam = OracleCamera()
        lib().vidc_oracle_camera_init(ctypes.c_double(fx), ctypes.c_double(fy), ctypes.c_double(cx),
                                      ctypes.c_double(cy), ctypes.byref(self.cam))
        self.W, self.H = int(self.cam.W), int(self.cam.H)
        self.K = np.array(self.cam.K, dtype=np.float32).reshape(3, 3)
        self.K_inv = np.array(self.cam.Kinv, dtype=np.float32).reshape(3, 3)

    # ref :35-58
    def build_homography(self, I_g, I_a):
        I_g, I_a = _f32(I_g), _f32(I_a)
        B = I_g.shape[0]
        H = np.empty((B, 3, 3), np.float32); R = np.empty_like(H); Hi = np.empty_like(H)
        lib().vidc_oracle_build_homography(ctypes.byref(self.cam), _p(I_g), _p(I_a), B, _p(H), _p(R), _p(Hi))
        return H, R, Hi

    # ref :125-140 -> px_min, py_min, kw, kh, 1/kw, 1/kh, w_max, h_max per frame
    def frame_scale(self, H):
        H = _f32(H)
        out = np.empty((H.shape[0], 8), np.float32)
        for i in range(H.shape[0]):
            lib().vidc_oracle_frame_scale(ctypes.byref(self.cam), _p(H[i]), _p(out[i]))
        return out

    # ref :158-214
    def image_sampler_forward_inverse(self, I_g, I_a):
        I_g, I_a = _f32(I_g), _f32(I_a)
        B = I_g.shape[0]
        Rt = np.empty((B, 3, 3), np.float32)
        grid = np.empty((B, self.H, self.W, 2), np.float32); inv = np.empty_like(grid)
        lib().vidc_oracle_sampler_forward_inverse(ctypes.byref(self.cam), _p(I_g), _p(I_a), B, _p(Rt), _p(grid), _p(inv))
        return Rt, grid, inv

    # ATen grid_sampler_2d (align_corners=False, zeros)
    def grid_sample(self, x, grid, mode="bilinear"):
        x, grid = _f32(x), _f32(grid)
        B, C, Hin, Win = x.shape
        Ho, Wo = grid.shape[1:3]
        out = np.empty((B, C, Ho, Wo), np.float32)
        lib().vidc_oracle_grid_sample(_p(x), B, C, Hin, Win, _p(grid), Ho, Wo, _MODES[mode], _p(out))
        return out

    # ref :108-156
    def warp_with_gravity_center_aligned(self, x, I_g, I_a, interp_mode="bilinear"):
        x, I_g, I_a = _f32(x), _f32(I_g), _f32(I_a)
        squeeze = x.ndim == 3
        if squeeze:
            x = x[:, None]
        B, C, Hin, Win = x.shape
        assert B == I_g.shape[0]
        H = np.empty((B, 3, 3), np.float32)
        y = np.empty((B, C, self.H, self.W), np.float32)
        lib().vidc_oracle_warp_forward(ctypes.byref(self.cam), _p(x), B, C, Hin, Win, _p(I_g), _p(I_a),
                                       _MODES[interp_mode], _p(H), _p(y))
        return H, (y[:, 0] if squeeze else y)

    # ref :216-255
    def inverse_warp_normal_image_with_gravity_center_aligned(self, x, I_g, I_a):
        x, I_g, I_a = _f32(x), _f32(I_g), _f32(I_a)
        B = x.shape[0]
        assert B == I_g.shape[0] and x.shape[1:] == (3, self.H, self.W)
        H = np.empty((B, 3, 3), np.float32)
        z = np.empty_like(x)
        lib().vidc_oracle_inverse_warp_normals(ctypes.byref(self.cam), _p(x), B, _p(I_g), _p(I_a), _p(H), _p(z))
        return H, z


# networks/surface_normal.py:170
def normalize(z):
    z = _f32(z)
    B, C = z.shape[:2]
    hw = int(np.prod(z.shape[2:]))
    out = np.empty_like(z)
    lib().vidc_oracle_normalize(_p(z), B, C, ctypes.c_size_t(hw), _p(out))
    return out


# networks/surface_normal.py:151
def validity_mask(x1):
    x1 = _f32(x1)
    B = x1.shape[0]
    hw = int(np.prod(x1.shape[2:]))
    m = np.empty((B, 1) + x1.shape[2:], np.uint8)
    lib().vidc_oracle_mask(_p(x1), B, ctypes.c_size_t(hw), _p(m))
    return m


# networks/surface_normal.py:153-156
def mask_nearest(mask, size):
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    B, _, Hin, Win = mask.shape
    out = np.empty((B, 1, size[0], size[1]), np.uint8)
    lib().vidc_oracle_mask_nearest(_p(mask), B, Hin, Win, size[0], size[1], _p(out))
    return out


# normal_utils.py:20-34
def normal_stats(norm_gt, pred, mask, normalize_prediction=True):
    gt, pred, mask = _f32(norm_gt), _f32(pred[:, 0:3]), _f32(mask)
    B = gt.shape[0]
    hw = int(np.prod(gt.shape[2:]))
    out = np.zeros(3, np.float64)
    lib().vidc_oracle_normal_stats(_p(gt), _p(pred), _p(mask), B, ctypes.c_size_t(hw), int(normalize_prediction), _p(out))
    angle_sum, num, l1 = out
    return {"angle_sum": angle_sum, "num": num, "l1_sum": l1, "loss": l1 / num if num else float("nan")}


def warp_unwarp_mt(orc: Oracle, rgb, depth, normals, I_g, I_a, nthreads=1):
    """One frame step of the whole path (what bench.py times): returns rgb_w, depth_w, mask_u8, normals_cam."""
    rgb, normals, I_g, I_a = _f32(rgb), _f32(normals), _f32(I_g), _f32(I_a)
    B = rgb.shape[0]
    H, W = orc.H, orc.W
    depth = _f32(depth).reshape(B, H, W) if depth is not None else None
    rgb_w = np.empty((B, 3, H, W), np.float32)
    depth_w = np.empty((B, H, W), np.float32) if depth is not None else None
    mask = np.empty((B, 1, H, W), np.uint8)
    ncam = np.empty((B, 3, H, W), np.float32)
    lib().vidc_oracle_warp_unwarp_mt(ctypes.byref(orc.cam), B, int(nthreads), _p(rgb), _p(depth) if depth is not None else None,
                                     _p(normals), _p(I_g), _p(I_a), _p(rgb_w), _p(depth_w) if depth is not None else None,
                                     _p(mask), _p(ncam))
    return rgb_w, depth_w, mask, ncam


# dataset.py:45-55 (rule='scannet'), :334-345 / :472-483 (rule='azure')
def condition_gravity(raw, rule="azure"):
    raw = _f32(raw)
    B = raw.shape[0]
    Ig = np.empty((B, 3), np.float32); Ia = np.empty((B, 3), np.float32)
    lib().vidc_oracle_condition_gravity(_p(raw), B, 0 if rule == "azure" else 1, _p(Ig), _p(Ia))
    return Ig, Ia
