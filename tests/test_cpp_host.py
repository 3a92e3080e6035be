"""The C++ host classes (include/btlbf/*.hpp): compile against the C ABI everywhere; run the reference's
unit scenarios through them on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_classes")


def _build():
    import _oracle
    import btl_bloomfilter_b200 as B
    _oracle.build_oracle()
    B.lib()
    pkg = os.path.join(ROOT, "btl_bloomfilter_b200")
    cmd = ["g++", "-std=c++11", "-O1", "-fopenmp", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_host_classes.cpp"), "-o", EXE,
           "-L" + pkg, "-lbtlbf_cuda", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)


def test_cpp_host_classes_compile_and_link():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_reference_unit_scenarios(tmp_path):
    _build()
    out = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout
