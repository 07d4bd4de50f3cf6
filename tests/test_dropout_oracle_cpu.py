"""Host restatement of the dropout mask definition (oracle/dropout.py): constants, determinism, statistics.  The
bit-for-bit comparison with the kernels is tests/test_dropout_gpu.py::test_masks_equal_host_restatement."""
import numpy as np

from oracle import dropout as od


def test_threshold_and_scale():
    assert od.thresh16(0.1) == 6554 and od.thresh16(0.0) == 0 and od.thresh16(1.0) == 65535
    assert abs(od.scale(0.1) - 1.0 / (1.0 - 6554 / 65536.0)) < 1e-6 and od.scale(0.0) == 1.0
    assert od.site_seed32(0, 0, 0) == ((0x9E3779B97F4A7C15 ^ (0x9E3779B97F4A7C15 >> 32)) & 0xFFFFFFFF)


def test_mask_is_a_pure_function_with_bernoulli_statistics():
    a = od.keep_mask(2048, 768, 0.1, 12345, 3, 2)
    assert a.dtype == np.uint8 and a.shape == (2048, 768)
    assert np.array_equal(a, od.keep_mask(2048, 768, 0.1, 12345, 3, 2))
    assert np.array_equal(a[:100, :64], od.keep_mask(100, 64, 0.1, 12345, 3, 2))       # keyed by (row, column) only
    assert abs(a.mean() - (1 - 6554 / 65536.0)) < 2e-3
    for b in (od.keep_mask(2048, 768, 0.1, 12345, 3, 3), od.keep_mask(2048, 768, 0.1, 12345, 4, 2),
              od.keep_mask(2048, 768, 0.1, 12346, 3, 2), od.keep_mask(2048, 768, 0.1, 12345, 3, 2, row_mul=128)):
        assert not np.array_equal(a, b)
        x, y = a.astype(np.float64) - a.mean(), b.astype(np.float64) - b.mean()
        assert abs((x * y).mean() / (x.std() * y.std())) < 5e-3
    f = a.astype(np.float64)
    for u, v in ((f[:, 0::2], f[:, 1::2]), (f[:, :-2], f[:, 2:]), (f[:, :-8], f[:, 8:]), (f[:-1], f[1:])):
        u, v = u - u.mean(), v - v.mean()
        assert abs((u * v).mean() / (u.std() * v.std())) < 5e-3
    assert od.keep_mask(4, 8, 0.0, 1, 0, 0).all()
