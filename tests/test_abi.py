"""The C-ABI shared library loads on a box without a GPU and exports every symbol include/btlbf.h
declares; calls that need a device fail with an error code and a message, never with a CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import btl_bloomfilter_b200 as B
    return B.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "btlbf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(btlbf_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(L):
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "libbtlbf_cuda.so does not export %s" % s


def test_python_binding_covers_the_header():
    from btl_bloomfilter_b200 import _capi
    assert set(declared_symbols()) == set(_capi.SIGNATURES) | {"btlbf_last_error"}


def test_version_and_header_text(L, golden):
    assert L.btlbf_version() == 100
    c = golden["bf_cases"][0]
    buf = C.create_string_buffer(1024)
    n = C.c_size_t()
    assert L.btlbf_format_header(0, c["bits"], c["bits"] // 8, c["h"], c["k"], 0.0, 0, 0, buf, 1024, C.byref(n)) == 0
    assert buf.raw[: n.value].decode() == c["header"]
    c = golden["cbf_cases"][1]
    assert L.btlbf_format_header(1, c["size_rounded"], c["size_rounded"], c["h"], c["k"], 0.0, 0, 0, buf, 1024,
                                 C.byref(n)) == 0
    assert buf.raw[: n.value].decode() == c["header"]


def test_header_double_format_matches_reference(L, ref, tmp_path):
    """dFPR is written by cpptoml with showpoint + 17 significant digits (cpptoml.h:3477-3494)."""
    import _oracle as O
    buf = C.create_string_buffer(1024)
    n = C.c_size_t()
    for dfpr, ne, te in [(0.0, 0, 0), (0.01, 5, 6), (1e-9, 2**40, 3), (0.5, 1, 2), (3.0, 0, 0), (1.5e300, 9, 9),
                         (0.0078125, 10, 11)]:
        f = ref.bf_new(1024, 3, 7)
        ref.L.ref_bf_set_meta(f, dfpr, ne, te)
        p = str(tmp_path / "m.bf")
        ref.L.ref_bf_store(f, p.encode())
        ref.L.ref_bf_free(f)
        raw = open(p, "rb").read()
        want = raw[: raw.index(b"[HeaderEnd]\n") + 12]
        assert L.btlbf_format_header(0, 1024, 128, 3, 7, dfpr, ne, te, buf, 1024, C.byref(n)) == 0
        assert buf.raw[: n.value] == want


def test_no_cpu_fallback_without_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.btlbf_ctx_create(0, C.byref(h))
    assert rc == 2  # BTLBF_ERR_CUDA
    assert L.btlbf_last_error()
    import btl_bloomfilter_b200 as B
    with pytest.raises(B.BtlbfError):
        B.BloomFilter(1024, 4, 5)
