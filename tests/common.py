"""Seeded synthetic inputs shared by the CPU tests, the GPU parity tests, the golden generator and bench.py.

numpy RandomState only (bit-stable across machines and torch versions).  Configs follow SURVEY.md section 8(d).
"""
import numpy as np

from vi_depth_completion_b200.synthetic import (CAMERAS, gravity_from_angles, random_gravity, extreme_roll_gravity,  # noqa: F401
                                                random_images, smooth_images)


def edge_case_gravity():
    """Edge cases listed in SURVEY.md section 8(c): g=a, g=-a, near-degenerate, roll +-90, pitch 45,
    un-normalised g, general alignment directions, aspect-guard triggers."""
    g = [
        [0.0, 1.0, 0.0],            # g = a -> R = I
        [0.0, -1.0, 0.0],           # g = -a -> n = 0 exactly -> R = I (not a 180 deg turn)
        [1e-5, -1.0, 0.0],          # theta ~ pi: garbage, non-orthonormal R
        [1.0, 0.0, 0.0],            # roll +90
        [-1.0, 0.0, 0.0],           # roll -90
        [0.0, 0.70710678, 0.70710678],   # pitch 45
        [0.0, 0.70710678, -0.70710678],  # pitch -45
        [0.3, 2.0, -0.4],           # un-normalised gravity (the reference never normalises I_g)
        [0.09, 0.99, -0.19],        # demo_dataset-like
        [0.2, 0.5, 0.84],           # pitch ~57 deg: strong perspective, aspect guard candidate
        [0.05, 0.9, 0.1],
        [0.6, 0.64, 0.48],
    ]
    a = [[0.0, 1.0, 0.0]] * len(g)
    a[10] = [0.0, 0.8, 0.6]         # dataset.py:483 style aligned direction [0, cos p, sin p]
    a[11] = [0.267261, 0.534522, 0.801784]   # fully general direction
    return np.array(g, np.float32), np.array(a, np.float32)


def loss_inputs(B, H, W, seed):
    """Inputs of the normal_utils losses: pred (B,4,H,W) ~ N(0,1) with a few exactly-zero vectors, unit-norm gt, a 0/1 mask
    with ~60 % ones, and the upstream gradient of the scalar loss."""
    rs = np.random.RandomState(seed)
    pred = rs.randn(B, 4, H, W).astype(np.float32)
    pred[0, :, 3::7, 2::5] = 0.0
    gt = rs.randn(B, 3, H, W).astype(np.float32)
    gt = (gt / np.sqrt((gt * gt).sum(1, keepdims=True))).astype(np.float32)
    maskf = (rs.rand(B, 1, H, W) < 0.6).astype(np.float32)
    return pred, gt, maskf, np.float32(1.7)


SPECIAL_SCALES = np.array([0.0, 1e-42, 1e-35, 1e-31, 1e-25, 1e-19, 1e-15, 1e-9, 1.0, 1e9, 1e15, 1e19, 1e25, 1e35, np.inf],
                          dtype=np.float32)


def special_value_images(B, H, W, seed):
    """Images that leave the comfortable range: signed zeros (x * 0 keeps the sign), denormals, huge values, inf and NaN,
    in bands, per pixel and per component.  They pin the behaviour the random tests cannot see: the sign of a zero result
    (ATen pads with +0 and MKL's GEMM accumulates from +0), NaN propagation through clamp_min, the IEEE fallbacks of the
    reciprocal fast paths.  B >= 3."""
    rgb, depth, normals = random_images(B, H, W, seed)
    rs = np.random.RandomState(seed + 77)
    sc = SPECIAL_SCALES
    band = max(H // sc.size, 1)
    with np.errstate(invalid="ignore", over="ignore"):
        for i, s in enumerate(sc):                                           # frame 0: one magnitude per band of rows
            normals[0, :, i * band:(i + 1) * band] *= s
        normals[0, :, :band] *= np.float32(-1.0)                             # the zero band: mixed +-0
        normals[0, :, :band, ::2] *= np.float32(-1.0)
        normals[1] *= sc[rs.randint(0, sc.size, size=(H, W))][None]          # frame 1: per pixel ...
        normals[1, 1] *= sc[rs.randint(0, sc.size, size=(H, W))]             # ... and per component
        normals[2, 0, ::3] = 0.0                                             # frame 2: exact-zero components, some NaN
        normals[2, 1, 1::5, ::2] = -0.0
        normals[2, 2, 7::11, 3::4] = np.nan
        rgb[0, :, : H // 3] *= np.float32(0.0)                               # +0 block (mask must be 0 there)
        rgb[0, :, H // 3: H // 2] *= np.float32(-0.0)                        # -0 block
        rgb[1, 0, 5::7, 2::5] = np.inf
        rgb[1, 1, 3::9, 1::6] = np.nan
        rgb[2] *= sc[rs.randint(5, 12, size=(H, W))][None]
        depth[0, : H // 2] *= np.float32(-0.0)
        depth[1, 4::6, 3::7] = np.nan
        depth[2, 2::5, 1::4] = np.inf
    return rgb, depth, normals


def special_value_gravity(B, seed=8):
    """Frame 0: gravity == alignment axis (R = I exactly, so exact zeros stay exact zeros); the rest moderate tilts."""
    I_g, I_a = random_gravity(B, seed=seed, roll_deg=25, pitch_deg=25)
    I_g[0], I_a[0] = (0.0, 1.0, 0.0), (0.0, 1.0, 0.0)
    return I_g, I_a


def degenerate_gravity():
    """Gravity / alignment vectors no IMU should produce: zero, NaN, inf, overflowing and underflowing magnitudes, signed
    zeros.  The parameter chain (atan2, cos, the bbox min / max) must still follow the reference bit for bit; frames whose
    homography is non-finite have non-finite sampling grids."""
    g = np.array([[0, 0, 0], [np.nan, 1, 0], [0, np.inf, 0], [1e30, 1e30, 0], [1e-30, 1e-30, 1e-30], [0, 1e-45, 0],
                  [3e38, 3e38, 3e38], [0, 1, 0], [0, 1, 0], [0, 1, 0], [-0.0, 1, -0.0], [0, 1, 1e-20], [1e-20, 1, 0],
                  [0.1, -np.inf, 0.2], [1e19, 1e19, 1e19], [1e-23, 1e-23, 1e-23]], np.float32)
    a = np.tile(np.array([[0, 1, 0]], np.float32), (g.shape[0], 1))
    a[7] = [0, 0, 0]; a[8] = [np.nan, 0, 1]; a[9] = [0, 1e20, 0]
    return g, a


def isolated_nonfinite_gravity():
    """Steep-pitch gravity vectors (found by search at 320x240, S1 intrinsics) for which the projective denominator is
    EXACTLY zero at a single pixel of the forward and / or inverse grid: one lane of a warp has a non-finite sampling
    coordinate while its neighbours are finite.  That lane must read as out of bounds (0), like ATen's CUDA kernel."""
    g = np.array([[0.06171473, 0.30808246, -0.94935584], [-0.13590553, 0.21624178, -0.9668346],
                  [-0.04061935, 0.31595406, 0.9479046], [-0.09833453, 0.21313821, 0.9720609]], np.float32)
    a = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (g.shape[0], 1))
    return g, a


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def count_bit_mismatches(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    return int(((bits(a) != bits(b)) & ~both_nan).sum())


def angular_error_deg(a, b):
    """Per-pixel angle between (B,3,H,W) vector fields, where the oracle vector b is non-zero."""
    a = a.astype(np.float64); b = b.astype(np.float64)
    na = np.sqrt((a * a).sum(1)); nb = np.sqrt((b * b).sum(1))
    ok = nb > 0
    cos = np.clip((a * b).sum(1)[ok] / np.maximum(na[ok] * nb[ok], 1e-300), -1, 1)
    return np.degrees(np.arccos(cos)), ok
