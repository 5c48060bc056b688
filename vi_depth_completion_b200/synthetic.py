"""Seeded synthetic inputs of the benchmark configurations: cameras, gravity vectors and images.

numpy RandomState only (bit-stable across machines and torch versions).  Shared by bench.py, the tools and -- through
tests/common.py, which re-exports everything here -- the test-suite and the golden generator.
"""
import numpy as np

# name -> (fx, fy, cx, cy)
CAMERAS = {
    "tiny":  (40.4, 40.4, 31.9, 23.9),                                  # 64x48, for full-tensor goldens
    "S1":    (202.0, 202.0, 159.93827, 119.938015),                     # main.py:243, surface_normal.py:61 -> 320x240
    "S2":    (404.0, 404.0, 319.87654, 239.87603),                      # Azure-Kinect-shaped 640x480
    "S3":    (577.87061, 580.25851, 319.87654, 239.87603),              # ScanNet 640x480 (surface_normal.py:60-61 x2)
    "default": (577.87061 * 0.5, 577.87061 * 0.5, 319.87654 * 0.5, 239.87603 * 0.5),  # constructor defaults, :6
}


def gravity_from_angles(roll, pitch):
    """g = (sin r cos p, cos r cos p, sin p), renormalised in fp32."""
    roll = np.asarray(roll, np.float32); pitch = np.asarray(pitch, np.float32)
    g = np.stack([np.sin(roll) * np.cos(pitch), np.cos(roll) * np.cos(pitch), np.sin(pitch)], 1).astype(np.float32)
    n = np.sqrt((g * g).sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (g / n).astype(np.float32)


def random_gravity(B, seed, roll_deg=30.0, pitch_deg=30.0):
    rs = np.random.RandomState(seed)
    roll = (rs.rand(B).astype(np.float32) * 2 - 1) * np.float32(np.deg2rad(roll_deg))
    pitch = (rs.rand(B).astype(np.float32) * 2 - 1) * np.float32(np.deg2rad(pitch_deg))
    I_g = gravity_from_angles(roll, pitch)
    I_a = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (B, 1))
    return I_g, I_a


def extreme_roll_gravity(B, seed):
    """Config S3: roll in {0, +-45, +-60, +-90} deg + U(-2,2) deg jitter, pitch 0."""
    rs = np.random.RandomState(seed)
    base = np.array([0, 45, -45, 60, -60, 90, -90], np.float32)
    roll = np.deg2rad(base[np.arange(B) % len(base)] + (rs.rand(B).astype(np.float32) * 4 - 2)).astype(np.float32)
    I_g = gravity_from_angles(roll, np.zeros(B, np.float32))
    I_a = np.tile(np.array([[0.0, 1.0, 0.0]], np.float32), (B, 1))
    return I_g, I_a


def random_images(B, H, W, seed, sparse_depth=False):
    """RGB ~ U[0,1), depth ~ U[0.4,10) m (dense) or ~150 non-zero px/frame (sparse), normals ~ N(0,1)."""
    rs = np.random.RandomState(seed)
    rgb = rs.rand(B, 3, H, W).astype(np.float32)
    depth = (rs.rand(B, H, W).astype(np.float32) * np.float32(9.6) + np.float32(0.4)).astype(np.float32)
    if sparse_depth:
        keep = np.zeros((B, H * W), bool)
        for b in range(B):
            keep[b, rs.choice(H * W, size=min(150, H * W), replace=False)] = True
        depth = np.where(keep.reshape(B, H, W), depth, np.float32(0)).astype(np.float32)
    normals = rs.randn(B, 3, H, W).astype(np.float32)
    return rgb, depth, normals


def smooth_images(B, H, W, seed):
    """Low-frequency images (real photographs are smooth): sums of a few sinusoids."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    out = np.zeros((B, 3, H, W), np.float32)
    for b in range(B):
        for c in range(3):
            f = rs.rand(4).astype(np.float32) * np.float32(0.05)
            ph = rs.rand(2).astype(np.float32) * np.float32(6.28)
            out[b, c] = (0.5 + 0.25 * np.sin(f[0] * xx + f[1] * yy + ph[0]) + 0.25 * np.cos(f[2] * xx - f[3] * yy + ph[1])).astype(np.float32)
    return out
