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


# ---- round 2: the division proof and the per-tile tables (csrc/frame_params.cuh) on the CPU ---------------------------------
def _f32_fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)   # exact product, one rounding


def test_inverse_division_proof_holds_for_every_pixel(hlib):
    """vidc::inv_division_proven lets the inverse kernels drop their per-pixel window test.  For every frame it accepts, the
    kernels' own fp32 u, v, s (fma(H1, Y, H0 X) + H2) must lie inside the window the shared-reciprocal division is exact in --
    |s| in [2^-40, 2^40], numerators zero or in [2^-80, 2^80] -- at EVERY pixel; frames whose horizon crosses the image, and
    degenerate gravity, must be rejected or pass the same check."""
    cams = ["S1", "S2", "tiny"]
    f = np.float32
    n_proven = 0
    for name in cams:
        g1, a1 = C.random_gravity(24, seed=5, roll_deg=90, pitch_deg=85)
        g2, a2 = C.edge_case_gravity()
        g3, a3 = C.degenerate_gravity()
        g4, a4 = C.isolated_nonfinite_gravity()
        I_g, I_a = np.concatenate([g1, g2, g3, g4]), np.concatenate([a1, a2, a3, a4])
        B = I_g.shape[0]
        from vi_depth_completion_b200._cabi import VidcCamera, lib
        c = VidcCamera()
        assert lib().vidc_camera_init(*[float(v) for v in C.CAMERAS[name]], ctypes.byref(c)) == 0
        flags = np.zeros(B, np.uint8); prm = np.zeros((B, 48), np.float32)
        hlib.host_inv_division_proven(ctypes.byref(c), _p(np.ascontiguousarray(I_g)), _p(np.ascontiguousarray(I_a)), B, _p(flags), _p(prm))
        X, Y = np.meshgrid(np.arange(c.W, dtype=f), np.arange(c.H, dtype=f))
        with np.errstate(all="ignore"):
            for b in range(B):
                Hm = prm[b, :9]
                s = _f32_fma(np.full_like(X, Hm[7]), Y, Hm[6] * X) + Hm[8]
                u = _f32_fma(np.full_like(X, Hm[1]), Y, Hm[0] * X) + Hm[2]
                v = _f32_fma(np.full_like(X, Hm[4]), Y, Hm[3] * X) + Hm[5]
                ok_s = (np.abs(s) >= 2.0 ** -40) & (np.abs(s) <= 2.0 ** 40)
                ok_n = ((u == 0) | ((np.abs(u) >= 2.0 ** -80) & (np.abs(u) <= 2.0 ** 80))) & ((v == 0) | ((np.abs(v) >= 2.0 ** -80) & (np.abs(v) <= 2.0 ** 80)))
                if flags[b]:
                    n_proven += 1
                    assert bool(ok_s.all() and ok_n.all()), f"{name} frame {b}: accepted by the proof but a pixel leaves the window"
    assert n_proven >= 60                                           # the proof is not vacuous: ordinary frames pass it
    # and every frame of the bench workload passes (the kernels run without the window test there)
    I_g, I_a = C.random_gravity(256, seed=1234, roll_deg=30, pitch_deg=30)
    c = VidcCamera()
    lib().vidc_camera_init(*[float(v) for v in C.CAMERAS["S2"]], ctypes.byref(c))
    flags = np.zeros(256, np.uint8)
    hlib.host_inv_division_proven(ctypes.byref(c), _p(I_g), _p(I_a), 256, _p(flags), None)
    assert flags.all()


def test_tile_tables_cover_the_footprints(hlib, oracle_mod):
    """The per-tile tables are hints (no result depends on them) but a useless hint would cost the speed they exist for:
    the staged-inverse boxes must contain the taps of (nearly) every pixel of their tile, the forward source boxes must
    contain every in-image tap of their tile -- checked against the oracle's sampling grids on the S2 workload."""
    from vi_depth_completion_b200._cabi import VidcCamera, lib
    cam = C.CAMERAS["S2"]
    c = VidcCamera()
    lib().vidc_camera_init(*[float(v) for v in cam], ctypes.byref(c))
    B = 6
    I_g, I_a = C.random_gravity(B, seed=1234, roll_deg=30, pitch_deg=30)
    tx, ty = (c.W + 31) // 32, (c.H + 31) // 32
    inv = np.zeros((B, ty, tx, 4), np.uint32); fwd = np.zeros((B, ty, tx, 4), np.uint32)
    hlib.host_tile_tables(ctypes.byref(c), _p(I_g), _p(I_a), B, _p(inv), _p(fwd))
    orc = oracle_mod.Oracle(*cam)
    _, grid, igrid = orc.image_sampler_forward_inverse(I_g, I_a)
    W, H = c.W, c.H

    def taps(gr):
        ix = ((gr[..., 0] + 1) * W - 1) / 2; iy = ((gr[..., 1] + 1) * H - 1) / 2
        return np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)

    x0, y0 = taps(igrid)
    inside = total = 0
    for b in range(B):
        for j in range(ty):
            for i in range(tx):
                e = inv[b, j, i]
                nsub = (int(e[1]) >> 16) & 3
                xs = x0[b, j * 32:(j + 1) * 32, i * 32:(i + 1) * 32]; ys = y0[b, j * 32:(j + 1) * 32, i * 32:(i + 1) * 32]
                total += xs.size
                if nsub == 0:
                    continue
                for sub in range(nsub):
                    w0, w1 = int(e[2 * sub]), int(e[2 * sub + 1])
                    bx0 = ((w0 & 0xffff) ^ 0x8000) - 0x8000; by0 = (w0 >> 16) - (0x10000 if w0 >> 31 else 0)
                    bw, bh = w1 & 0xff, (w1 >> 8) & 0xff
                    rows = slice(0, 32) if nsub == 1 else slice(16 * sub, 16 * sub + 16)
                    ax, ay = xs[rows] - bx0, ys[rows] - by0
                    inside += int(((ax >= 0) & (ax < bw - 1) & (ay >= 0) & (ay < bh - 1)).sum())
                    assert bx0 % 4 == 0 and bw <= 60 and bh <= 40
    assert inside / total > 0.995, inside / total
    x0, y0 = taps(grid)
    missed = seen = 0
    for b in range(B):
        for j in range(ty):
            for i in range(tx):
                bx, by, bw, bh = [int(v) for v in fwd[b, j, i]]
                xs = x0[b, j * 32:(j + 1) * 32, i * 32:(i + 1) * 32]; ys = y0[b, j * 32:(j + 1) * 32, i * 32:(i + 1) * 32]
                live = (xs >= 0) & (xs < W - 1) & (ys >= 0) & (ys < H - 1)
                seen += int(live.sum())
                if bw == 0:
                    missed += int(live.sum())
                    continue
                missed += int((live & ~((xs >= bx) & (xs + 1 < bx + bw) & (ys >= by) & (ys + 1 < by + bh))).sum())
    assert missed / seen < 0.002, missed / seen


def test_u8_to_unit_is_to_tensor(hlib):
    """vidc::u8_to_unit (the ToTensor kernels' x / 255, three fp32 operations) against torch's .div(255) and torchvision's
    ToTensor on a PIL image (dataset.py:468-471) for every uint8 value: bit-identical."""
    import torch
    from PIL import Image
    from torchvision import transforms
    vals = np.arange(256, dtype=np.uint8)
    out = np.empty(256, np.float32)
    hlib.host_u8_to_unit(vals.ctypes.data_as(ctypes.c_void_p), 256, out.ctypes.data_as(ctypes.c_void_p))
    want = torch.from_numpy(vals).to(torch.float32).div(255).numpy()
    assert C.count_bit_mismatches(out, want) == 0
    img = np.stack([vals.reshape(16, 16)] * 3, -1)                                   # (16,16,3) uint8, every value once per channel
    tv = transforms.ToTensor()(Image.fromarray(img)).numpy()
    assert C.count_bit_mismatches(tv[0].reshape(-1), out) == 0
