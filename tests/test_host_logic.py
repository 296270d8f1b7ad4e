"""Host-side logic of the product that needs no GPU: C-ABI exports, the libstdc++ std::sort port used by
the quadtree kernel (same source compiled for the host), the glibc sinf/cosf port, shard bookkeeping."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_cabi_library_loads_and_exports_every_declared_symbol():
    from dani_slam_b200 import orbx
    L = orbx.lib()
    hdr = open(os.path.join(ROOT, "include", "orbx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", hdr)
    assert len(set(names)) >= 30
    for n in set(names):
        assert hasattr(L, n), f"{n} declared in include/orbx.h but not exported by liborbx.so"


def test_no_cpu_fallback_without_a_device():
    from dani_slam_b200 import orbx
    L = orbx.lib()
    if L.orbx_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    with pytest.raises(orbx.OrbxError) as e:
        orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher(0.7, True)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "dani_slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "liborb_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def _sort_both(oracle_mod, sizes, ulx):
    from dani_slam_b200 import orbx
    sizes = np.ascontiguousarray(sizes, np.int32)
    ulx = np.ascontiguousarray(ulx, np.int32)
    perm = np.zeros(len(sizes), np.int32)
    orbx.lib().orbx_debug_sort_nodes(_p(sizes), _p(ulx), len(sizes), _p(perm))
    return perm, oracle_mod.sort_nodes(sizes, ulx)


def test_stdsort_port_matches_libstdcxx_on_tie_heavy_arrays(oracle_mod):
    rng = np.random.default_rng(0)
    for n in list(range(0, 70)) + [100, 127, 128, 129, 255, 256, 257, 500, 1000, 1024, 1737, 4000]:
        for rep in range(6):
            sizes = rng.integers(2, 2 + max(1, rep * 3 + 1), n)        # few distinct sizes → many ties
            ulx = rng.integers(0, max(1, 4 + rep * 4), n) * 19
            a, b = _sort_both(oracle_mod, sizes, ulx)
            assert np.array_equal(a, b), (n, rep)


def test_stdsort_port_heap_fallback_path(oracle_mod):
    # median-of-3 killer sequences drive introsort into its heap-sort fallback (depth limit 2*lg n)
    def killer(n):
        a = [0] * n
        k = n // 2
        for i in range(1, k + 1):
            if i % 2 == 1:
                a[i - 1] = i
                a[i] = k + i
            a[k + i - 1] = 2 * i
        return a
    for n in [64, 256, 1000, 4096]:
        sizes = np.array(killer(n), np.int32) + 2
        a, b = _sort_both(oracle_mod, sizes, np.zeros(n, np.int32))
        assert np.array_equal(a, b), n
        sizes2 = np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]).astype(np.int32) + 2   # organ pipe
        a, b = _sort_both(oracle_mod, sizes2, (np.arange(n) % 3).astype(np.int32))
        assert np.array_equal(a, b), n


def test_glibc_sincosf_port_exhaustive_host_sweep(tmp_path):
    """Every float in [0, 2π·1.0001] (1.09e9 values): the port compiled for the host == this box's libm."""
    src = tmp_path / "sweep.cpp"
    src.write_text(r'''
#include "glibc_sincosf.h"
#include <cstdio>
#include <cmath>
#include <thread>
#include <vector>
#include <atomic>
int main() {
  float hi = 6.2831855f * 1.0001f; uint32_t hb; memcpy(&hb, &hi, 4);
  const int T = 8; std::atomic<long> bad{0};
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back([&, t] {
    long b = 0;
    for (uint32_t u = t; u <= hb; u += T) { float x; memcpy(&x, &u, 4);
      float s = sinf(x), c = cosf(x), s2 = orbx_libm::sinf_glibc(x), c2 = orbx_libm::cosf_glibc(x);
      if (memcmp(&s, &s2, 4) || memcmp(&c, &c2, 4)) b++; }
    bad += b; });
  for (auto &x : th) x.join();
  printf("%ld\n", bad.load());
  return 0; }
''')
    exe = tmp_path / "sweep"
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-pthread", "-I", os.path.join(ROOT, "dani_slam_b200", "csrc"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], timeout=600).decode().strip()
    assert out == "0", f"{out} floats in [0, 2pi] where the port differs from libm sinf/cosf"


def test_shard_bounds_partition():
    from dani_slam_b200.sharded import shard_bounds
    for n in [0, 1, 7, 8, 9, 1000, 10_000_000]:
        for w in [1, 2, 3, 4, 8]:
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_synth_frames_are_deterministic():
    from dani_slam_b200 import synth
    a, b = synth.throughput_frame(5), synth.throughput_frame(5)
    assert np.array_equal(a, b) and a.shape == (480, 640) and a.dtype == np.uint8
    assert not np.array_equal(a, synth.throughput_frame(6))
    q, db = synth.knn_case(50, 1000, seed=3)
    assert q.shape == (50, 32) and db.shape == (1000, 32)


def test_reference_arm_prints_the_contract_line_with_the_product_arms_config():
    """`bench.py --impl reference` (the reference's own CPU path on the host cores): one JSON line, `impl: reference`, the SAME
    `config` object the product arm prints for this workload (the driver compares them), a `cpu_baseline` describing the run and
    an `e2e` with no transfers.  No GPU involved."""
    import json
    import subprocess
    import sys
    import types

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "orb_frames_per_s" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["dtype"] == "u8" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["gpu_launches"] == 0
    sys.path.insert(0, root)
    import bench
    default_batch = bench.set_workload("tum1")
    assert line["config"] == bench.bench_config(types.SimpleNamespace(workload="tum1"), 1, default_batch)
