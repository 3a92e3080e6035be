"""The device arithmetic (csrc/nthash_dev.cuh is host+device) checked on the CPU against the oracle and
against Python integers: exact modulo by a launch-invariant divisor, split rotations, seeds, hash mixing."""
import ctypes as C

import numpy as np
import pytest

from _backends import EMU_SO, build_emu


@pytest.fixture(scope="module")
def E():
    build_emu()
    L = C.CDLL(EMU_SO)
    u64 = C.c_uint64
    L.emu_fastmod.restype = u64
    L.emu_fastmod.argtypes = [u64, u64]
    for n in ("emu_srol", "emu_sror"):
        getattr(L, n).restype = u64
        getattr(L, n).argtypes = [u64]
    L.emu_srol_n.restype = u64
    L.emu_srol_n.argtypes = [u64, C.c_uint]
    L.emu_multi_mix.restype = u64
    L.emu_multi_mix.argtypes = [u64, C.c_uint, C.c_uint]
    L.emu_class_seeds.restype = u64
    L.emu_class_seeds.argtypes = [C.c_uint, C.c_int]
    return L


def test_fastmod_is_exact(E):
    rng = np.random.default_rng(0)
    mods = [1, 2, 3, 7, 8, 1000, 8388608, 31_568_113_856, (1 << 35), (1 << 35) - 8, (1 << 32) + 8, (1 << 32) - 8,
            (1 << 32), 16_000_000_000, (1 << 37), (1 << 63), (1 << 63) + 8, (1 << 64) - 8, 4294967297, 999_999_999_989]
    mods += [int(x) for x in rng.integers(1, 1 << 62, 40, dtype=np.uint64)]
    xs = [0, 1, (1 << 64) - 1, (1 << 63), (1 << 32) - 1, (1 << 32), (1 << 32) + 1]
    xs += [int(x) for x in rng.integers(0, 1 << 63, 200, dtype=np.uint64)]
    xs += [int(x) | (1 << 63) for x in rng.integers(0, 1 << 63, 200, dtype=np.uint64)]
    for m in mods:
        xs_m = xs + [m - 1, m, m + 1, 2 * m - 1, 2 * m, (((1 << 64) - 1) // m) * m, (((1 << 64) - 1) // m) * m - 1]
        for x in xs_m:
            x &= (1 << 64) - 1
            assert E.emu_fastmod(x, m) == x % m, (x, m)


def test_rotations_seeds_and_mix_match_the_oracle(E, oracle):
    L = oracle.L
    rng = np.random.default_rng(1)
    for v in [0, 1, (1 << 64) - 1, 1 << 32, 1 << 33, 1 << 63] + [int(x) for x in rng.integers(0, 1 << 63, 300, dtype=np.uint64)]:
        assert E.emu_srol(v) == L.ora_srol(v)
        assert E.emu_sror(v) == L.ora_sror(v)
        for n in (0, 1, 25, 31, 32, 33, 64, 100, 1022, 1023, 65535):
            assert E.emu_srol_n(v, n) == L.ora_srol_n(v, n)
        for i, k in ((1, 25), (3, 32), (5, 4), (63, 100)):
            m = L.ora_multi_mult(i, k)
            t = (v * m) & ((1 << 64) - 1)
            assert E.emu_multi_mix(v, i, k) == t ^ (t >> 27)
    for c in range(256):
        assert E.emu_class_seeds(c, 0) == L.ora_seed(c)
        assert E.emu_class_seeds(c, 1) == L.ora_seed(c & 7) if L.ora_seed(c) else True
