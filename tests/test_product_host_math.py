"""The PRODUCT's own arithmetic (csrc/exact_math.cuh, csrc/frame_params.cuh -- the source the params kernel
compiles for the device) built for the host and pinned against libm, the oracle and the reference goldens.
No GPU, no oracle code in the product: this only checks that the shared __host__ __device__ source is right."""
import ctypes
import os

import numpy as np
import pytest

from tests import common as C
from tests.harness_build import build as build_harness

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def hlib():
    return ctypes.CDLL(build_harness())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_glibc_atan2f_restatement_vs_libm(hlib):
    hlib.host_atan2f_sweep_vs_libm.restype = ctypes.c_size_t
    hlib.host_atan2f_sweep_vs_libm.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_float]
    # x == 1.0 takes the atanf path: every 7th positive float, plus NaN / inf
    assert hlib.host_atan2f_sweep_vs_libm(0, 0x7fc00001, 7, 1.0) == 0
    for x in (0.5, 0.999, 0.7071, 0.1, -0.3, -1.0, 1e-3, 3.0, 0.0, -0.0, float("inf"), 1e30, -1e-30, 1e-38):
        assert hlib.host_atan2f_sweep_vs_libm(0, 0x7f800001, 997, x) == 0
        assert hlib.host_atan2f_sweep_vs_libm(0x80000000, 0xff800001, 1009, x) == 0


def test_mkl_cos_restatement_vs_oracle(hlib, oracle_mod):
    bits = np.arange(0, int(np.float32(1.6).view(np.uint32)) + 1, 61, dtype=np.uint32)
    x = bits.view(np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    hlib.host_mkl_cosf_ha(_p(x), ctypes.c_size_t(x.size), _p(a))
    oracle_mod.lib().vidc_oracle_cosf_array(_p(x), ctypes.c_size_t(x.size), _p(b))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def _product_params(hlib, cam, I_g, I_a):
    from vi_depth_completion_b200._cabi import VidcCamera, lib
    c = VidcCamera()
    assert lib().vidc_camera_init(*[float(v) for v in cam], ctypes.byref(c)) == 0   # host-only entry point
    B = I_g.shape[0]
    out = np.zeros((B, 48), np.float32)
    hlib.host_frame_params(ctypes.byref(c), _p(np.ascontiguousarray(I_g)), _p(np.ascontiguousarray(I_a)), B, _p(out))
    return c, out


@pytest.mark.parametrize("name", ["tiny", "S1", "S2", "S3"])
def test_product_params_match_reference_goldens(hlib, name):
    g = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    c, prm = _product_params(hlib, g["cam"], g["I_g"], g["I_a"])
    assert (c.W, c.H) == (int(g["W"]), int(g["H"]))
    assert C.count_bit_mismatches(np.array(c.K, np.float32), g["K"].reshape(-1)) == 0
    assert C.count_bit_mismatches(np.array(c.Kinv, np.float32), g["K_inv"].reshape(-1)) == 0
    B = prm.shape[0]
    assert C.count_bit_mismatches(prm[:, 0:9].reshape(B, 3, 3), g["Hm"]) == 0
    assert C.count_bit_mismatches(prm[:, 9:18].reshape(B, 3, 3), g["R"]) == 0
    assert C.count_bit_mismatches(prm[:, 18:27].reshape(B, 3, 3), g["Hinv"]) == 0


def test_product_params_match_oracle_random(hlib, oracle_mod):
    for cam_name in ("S1", "S2", "S3", "default"):
        I_g, I_a = C.random_gravity(4096, seed=17, roll_deg=89, pitch_deg=70)
        rs = np.random.RandomState(3)
        I_a[2048:] = rs.randn(2048, 3).astype(np.float32)        # arbitrary, un-normalised alignment directions
        I_g[3072:] = rs.randn(1024, 3).astype(np.float32)
        _, prm = _product_params(hlib, C.CAMERAS[cam_name], I_g, I_a)
        o = oracle_mod.Oracle(*C.CAMERAS[cam_name])
        H, R, Hi = o.build_homography(I_g, I_a)
        sc = o.frame_scale(H)
        B = I_g.shape[0]
        assert C.count_bit_mismatches(prm[:, 0:9].reshape(B, 3, 3), H) == 0
        assert C.count_bit_mismatches(prm[:, 9:18].reshape(B, 3, 3), R) == 0
        assert C.count_bit_mismatches(prm[:, 18:27].reshape(B, 3, 3), Hi) == 0
        assert C.count_bit_mismatches(prm[:, 27:35], sc) == 0


@pytest.mark.parametrize("rule,idx", [("azure", 0), ("scannet", 1)])
def test_product_gravity_conditioning_matches_reference_golden(hlib, rule, idx):
    """dataset.py:45-55 / :472-483: the product's device function, host build."""
    g = np.load(os.path.join(GOLD, "golden_gravity.npz"))
    raw = np.ascontiguousarray(g["raw"])
    Ig, Ia = np.empty_like(raw), np.empty_like(raw)
    hlib.host_condition_gravity(_p(raw), raw.shape[0], idx, _p(Ig), _p(Ia))
    assert C.count_bit_mismatches(Ig, g[rule + "_g"]) == 0
    assert C.count_bit_mismatches(Ia, g[rule + "_a"]) == 0


def test_mkl_sin_restatement_vs_oracle(hlib, oracle_mod):
    bits = np.concatenate([np.arange(0, int(np.float32(3.2).view(np.uint32)) + 1, 97, dtype=np.uint32),
                           np.arange(0x80000000, 0x80000000 + int(np.float32(3.2).view(np.uint32)) + 1, 101, dtype=np.uint32)])
    x = bits.view(np.float32)
    a, b = np.empty_like(x), np.empty_like(x)
    hlib.host_mkl_sinf_ha(_p(x), ctypes.c_size_t(x.size), _p(a))
    oracle_mod.lib().vidc_oracle_sinf_array(_p(x), ctypes.c_size_t(x.size), _p(b))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("cam_name", ["S1", "tiny", "S3"])
def test_exterior_tile_test_is_conservative(hlib, oracle_mod, cam_name):
    """csrc/frame_params.cuh: tile_certainly_exterior -- the test behind the forward kernels' exterior-tile bitmap -- compiled
    for the host and checked against the oracle: over random, extreme, edge-case, degenerate and pole-crossing gravity no
    marked tile may contain a pixel that receives anything from an all-ones source image, and the test must actually
    mark a good share of the truly exterior tiles (otherwise it is useless)."""
    from vi_depth_completion_b200._cabi import VidcCamera, lib
    cam = C.CAMERAS[cam_name]
    o = oracle_mod.Oracle(*cam)
    c = VidcCamera()
    assert lib().vidc_camera_init(*[float(v) for v in cam], ctypes.byref(c)) == 0
    tiles_x, tiles_y = (o.W + 31) // 32, (o.H + 31) // 32
    marked = truly = wrong = 0
    rs = np.random.RandomState(3)
    sets = [C.random_gravity(40, 1, 30, 30), C.random_gravity(40, 2, 89, 75), C.extreme_roll_gravity(14, 3),
            C.edge_case_gravity(), C.degenerate_gravity(), C.isolated_nonfinite_gravity()]
    ga = C.random_gravity(24, 5, 60, 60)
    sets.append((ga[0], rs.randn(24, 3).astype(np.float32)))                     # general alignment directions
    if cam_name == "S3":                                                         # 640x480 (20 x 15 tiles): a lighter selection
        sets = [C.random_gravity(10, 1, 30, 30), C.random_gravity(10, 2, 89, 75), C.extreme_roll_gravity(7, 3),
                (ga[0][:8], sets[-1][1][:8])]
    for I_g, I_a in sets:
        B = I_g.shape[0]
        bits = np.zeros((B, tiles_y, tiles_x), np.uint8)
        hlib.host_exterior_tiles(ctypes.byref(c), _p(np.ascontiguousarray(I_g)), _p(np.ascontiguousarray(I_a)), B, _p(bits))
        with np.errstate(all="ignore"):
            _, y = o.warp_with_gravity_center_aligned(np.ones((B, 1, o.H, o.W), np.float32), I_g, I_a)
        hit = np.zeros((B, tiles_y * 32, tiles_x * 32), bool)
        hit[:, :o.H, :o.W] = (y[:, 0] != 0) | np.isnan(y[:, 0])
        hit_tiles = hit.reshape(B, tiles_y, 32, tiles_x, 32).any(axis=(2, 4))
        wrong += int((bits.astype(bool) & hit_tiles).sum())
        marked += int(bits.sum()); truly += int((~hit_tiles).sum())
    assert wrong == 0
    print(f"{cam_name}: {marked} of {truly} exterior tiles marked")
    if cam_name != "tiny":                # on the 2 x 2 tiles of the tiny canvas few tiles are exterior at all
        assert marked > 0.6 * truly > 0
