"""The C++ drop-in adapters (include/ORBextractor.h with the reference's signatures, include/ORBmatcher_orbx.h,
include/ORBVocabulary_orbx.h):
compile check on the CPU, and on the GPU a C++ program that calls them like Frame::ExtractORB does and
compares with the reference's own ORBextractor.cc (oracle/_ref)."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "adapter_check.cpp")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "tests", "cpp", "shim"),
       "-DORBX_MATCHER_HOT_PATH_ONLY"]      # the out-of-scope ORBmatcher members need the reference's KeyFrame / Sophus headers


def test_adapters_compile_against_the_opencv_shim():
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fsyntax-only", "-Wall"] + INC + [SRC])


@pytest.mark.gpu
def test_cpp_adapter_matches_reference_source(tmp_path, orbx_mod):
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "libref_orb.so")):
        pytest.skip("oracle/_ref not built")
    exe = str(tmp_path / "adapter_check")
    libdirs = [os.path.join(ROOT, "dani_slam_b200"), ref_dir, os.path.join(ROOT, "oracle")]
    with_bow = os.path.exists(os.path.join(ref_dir, "libref_bow.so"))
    with_match = os.path.exists(os.path.join(ref_dir, "libref_match.so"))
    subprocess.check_call(["g++", "-std=c++17", "-O2"] + (["-DWITH_REF_BOW"] if with_bow else []) + (["-DWITH_REF_MATCH"] if with_match else []) + INC +
                          [SRC, "-o", exe] + ["-L" + d for d in libdirs] + ["-lorbx", "-lref_orb"] + (["-lref_bow"] if with_bow else []) +
                          (["-lref_match"] if with_match else []) +
                          ["-lorb_oracle", "-Wl,-rpath," + ":".join(libdirs)])
    from dani_slam_b200 import synth
    voc_path = str(tmp_path / "voc.txt")
    synth.write_vocabulary_text(synth.vocabulary(k=10, L=4, seed=13), voc_path)
    for args in (["640", "480", "1", voc_path], ["752", "480", "2"], ["401", "333", "3", voc_path]):
        r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.strip().startswith("OK"), r.stdout
