"""The kernels' shared-reciprocal division helpers must produce the correctly rounded quotient, i.e. the
same bits as the compiler's IEEE division, for every operand -- inside and outside their fast window."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(u, v, s, dev):
    from vi_depth_completion_b200 import _cabi
    n = u.size
    tu, tv, ts = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in (u, v, s))
    out = torch.empty(4 * n, dtype=torch.float32, device=dev)
    _cabi.check(_cabi.lib().vidc_debug_div(tu.data_ptr(), tv.data_ptr(), ts.data_ptr(), n, out.data_ptr(),
                                           ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    o = out.cpu().numpy().view(np.uint32).reshape(4, n)
    nan = lambda a: (a & 0x7fffffff) > 0x7f800000
    bad_u = (o[0] != o[2]) & ~(nan(o[0]) & nan(o[2]))
    bad_v = (o[1] != o[3]) & ~(nan(o[1]) & nan(o[3]))
    return int(bad_u.sum()), int(bad_v.sum())


def test_division_typical_range(cuda_device):
    rs = np.random.RandomState(0)
    n = 1 << 24
    u = (rs.rand(n).astype(np.float32) - 0.5) * 4000
    v = (rs.rand(n).astype(np.float32) - 0.5) * 3000
    s = rs.rand(n).astype(np.float32) * 2 + 0.25
    assert _run(u, v, s, cuda_device) == (0, 0)
    # normalisation-like: components / norm
    z = rs.randn(3, n).astype(np.float32)
    nrm = np.sqrt((z * z).sum(0)).astype(np.float32)
    assert _run(z[0], z[1], nrm, cuda_device) == (0, 0)


def test_division_random_bit_patterns(cuda_device):
    rs = np.random.RandomState(1)
    n = 1 << 24
    bits = lambda: rs.randint(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    assert _run(bits(), bits(), bits(), cuda_device) == (0, 0)


def test_division_special_values(cuda_device):
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.17549435e-38, 3.4e38, 1e-30, 1e30,
                   2.0 ** -80, 2.0 ** 80, 2.0 ** -40, 2.0 ** 40, np.nextafter(np.float32(2.0 ** -80), np.float32(0)),
                   np.nextafter(np.float32(2.0 ** 40), np.float32(np.inf)), 1e-12, 0.5, 3.0, 1 / 3], np.float32)
    u, v, s = np.meshgrid(sp, sp, sp, indexing="ij")
    assert _run(u.ravel(), v.ravel(), s.ravel(), cuda_device) == (0, 0)
