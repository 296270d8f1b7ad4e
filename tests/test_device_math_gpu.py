"""Device-side scalar building blocks vs the host definitions: glibc sinf/cosf port, fastAtan2, std::sort port."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _sweep(orbx_mod, oracle_mod, bits):
    a = bits.view(np.float32)
    s = np.empty_like(a); c = np.empty_like(a)
    assert orbx_mod.lib().orbx_debug_sincos_device(0, _p(a), a.size, _p(s), _p(c)) == 0
    rs, rc = oracle_mod.sincos_array(a, nthreads=16)
    return int((s.view(np.uint32) != rs.view(np.uint32)).sum() + (c.view(np.uint32) != rc.view(np.uint32)).sum())


def test_device_sincosf_exhaustive_sweep_vs_host_libm(orbx_mod, oracle_mod):
    """EVERY float in [0, 2π] (1.09e9 bit patterns, chunks of 2^26): device port == this box's libm sinf/cosf."""
    hi = int(np.float32(6.2831855 * 1.0001).view(np.uint32))
    bad = 0
    step = 1 << 26
    for lo in range(0, hi + 1, step):
        bits = np.arange(lo, min(lo + step, hi + 1), dtype=np.uint32)
        bad += _sweep(orbx_mod, oracle_mod, bits)
    assert bad == 0


def test_device_fast_atan2_vs_oracle(orbx_mod, oracle_mod):
    rng = np.random.default_rng(0)
    n = 2_000_000
    y = rng.integers(-1_300_000, 1_300_000, n).astype(np.float32)
    x = rng.integers(-1_300_000, 1_300_000, n).astype(np.float32)
    y[:8] = [0, 0, 5, -5, -1, 1, -0.0, 7]; x[:8] = [0, 5, 0, 0, 1e6, -1e6, 3, 7]
    out = np.empty(n, np.float32)
    assert orbx_mod.lib().orbx_debug_atan2_device(0, _p(y), _p(x), n, _p(out)) == 0
    L = oracle_mod.lib()
    idx = np.r_[0:8, rng.integers(0, n, 20000)]
    ref = np.array([L.orc_fast_atan2(float(y[i]), float(x[i])) for i in idx], np.float32)
    assert np.array_equal(out[idx].view(np.uint32), ref.view(np.uint32))
    assert out.min() >= 0 and out.max() <= 360


def test_device_stdsort_port_vs_libstdcxx(orbx_mod, oracle_mod):
    rng = np.random.default_rng(1)
    for n in [1, 2, 16, 17, 33, 64, 200, 256, 777, 1737]:
        for rep in range(3):
            sizes = rng.integers(2, 4 + rep * 4, n).astype(np.int32)
            ulx = (rng.integers(0, 6 + rep * 10, n) * 19).astype(np.int32)
            perm = np.zeros(n, np.int32)
            assert orbx_mod.lib().orbx_debug_sort_nodes_device(0, _p(sizes), _p(ulx), n, _p(perm)) == 0
            assert np.array_equal(perm, oracle_mod.sort_nodes(sizes, ulx)), (n, rep)
