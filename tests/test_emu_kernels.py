"""Kernel logic on the CPU: the per-thread phase functions of the CUDA kernels (tile_core.cuh, the
source nvcc compiles) executed thread by thread by tests/emu/emu.cpp, with capi.cu's chunk / batch /
residual-round orchestration, checked bit-exactly against the oracle and the reference-generated golden
vectors.  This is what can be verified without a GPU; tests/test_gpu_parity.py runs the same suite
through the C ABI on the B200."""
import pytest

import parity_suite as S
from _backends import EmuBackend


@pytest.fixture(scope="module")
def emu():
    return EmuBackend()


@pytest.fixture(scope="module")
def emu_small():
    # tiny chunks / batches / reservation tables: exercises chunk halos, batch boundaries, hashed-slot
    # false conflicts and multi-round residual lists
    return EmuBackend(chunk=4096, batch=4096, resv_log2=10, list_log2=6)


def test_golden_hashes(emu, golden):
    S.check_golden_hashes(emu, golden)


def test_golden_hashes_generic_path(golden):
    S.check_golden_hashes(EmuBackend(force_generic=1), golden)


def test_golden_bf(emu, golden):
    S.check_golden_bf(emu, golden)


def test_golden_bf_early_exit_mode(golden):
    S.check_golden_bf(EmuBackend(query_mode=1), golden)


def test_golden_cbf(emu, golden):
    S.check_golden_cbf(emu, golden)


def test_golden_cbf_small_tables(emu_small, golden):
    S.check_golden_cbf(emu_small, golden)
    S.check_golden_bf(emu_small, golden)


@pytest.mark.parametrize("k,h", [(1, 1), (2, 3), (4, 5), (25, 4), (31, 2), (32, 6), (33, 3), (64, 4), (100, 2)])
def test_random_hashes(emu, oracle, k, h):
    S.check_random_hashes(emu, oracle, k, h, seed=1000 * k + h)


def test_random_hashes_exotic_bytes(emu, oracle):
    # arbitrary byte values, including the raw bytes 1,3,4,5,7 that seedTab accepts (nthash.hpp:195-228)
    S.check_random_hashes(emu, oracle, 5, 3, seed=77, exotic=0.05)
    S.check_random_hashes(emu, oracle, 25, 4, seed=78, exotic=0.01)


def test_random_hashes_multi_tile(emu_small, oracle):
    S.check_random_hashes(emu_small, oracle, 25, 4, seed=5, n_seqs=120, max_len=400)
    S.check_random_hashes(emu_small, oracle, 64, 2, seed=6, n_seqs=3, max_len=9000)


@pytest.mark.parametrize("k,n_seeds,h2", [(5, 2, 2), (31, 2, 1), (16, 3, 3), (40, 1, 4)])
def test_random_spaced(emu, oracle, k, n_seeds, h2):
    S.check_random_spaced(emu, oracle, k, n_seeds, h2, seed=k)


@pytest.mark.parametrize("k,h,bits", [(25, 4, 1 << 16), (32, 6, 8 * 1237), (4, 5, 1024), (21, 3, 8 * 4099)])
def test_random_bf(emu, oracle, k, h, bits):
    S.check_random_bf(emu, oracle, k, h, bits, seed=bits + k)


def test_random_bf_multi_chunk(emu_small, oracle):
    S.check_random_bf(emu_small, oracle, 25, 4, 8 * 3001, seed=11, n_seqs=100, max_len=300)


@pytest.mark.parametrize("k,h,m", [(25, 4, 4096), (8, 5, 100008), (5, 3, 64), (11, 4, 512)])
def test_random_cbf(emu, oracle, k, h, m):
    S.check_random_cbf(emu, oracle, k, h, m, seed=m + k)


def test_random_cbf_small_tables(emu_small, oracle):
    S.check_random_cbf(emu_small, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    d, r = 0, 0
    f = emu_small.filter(1, 64, 3, 5)
    f.insert(["ACGTA" * 40] * 8)
    d, r = f.ordered_stats()
    assert d > 0 and r > 1  # the residual rounds really ran


def test_edge_cases(emu, oracle):
    S.check_edge_cases(emu, oracle)


def test_cfg1(emu, oracle, golden):
    S.check_cfg1(emu, oracle, golden)


# ---------------------------------------------------------------- partitioned (binned) BloomFilter build
@pytest.mark.parametrize("shift,writers,slack", [(10, 3, 20), (8, 5, 0), (12, 1, 300), (5, 2, 50)])
def test_binned_build_equals_direct(oracle, golden, shift, writers, slack):
    be = EmuBackend(bin_shift=shift, bin_writers=writers, bin_slack_pct=slack)
    S.check_golden_bf(be, golden)
    S.check_random_bf(be, oracle, 25, 4, 1 << 16, seed=3)
    S.check_random_bf(be, oracle, 32, 6, 32 * 1237, seed=4)   # non-power-of-two size, partial last partition
    S.check_random_spaced(be, oracle, 16, 3, 3, seed=16)


def test_binned_build_overflow_falls_back_to_direct_atomics(oracle):
    """A skewed batch (one k-mer repeated) overflows its sub-buckets; the overflow goes through the direct path."""
    be = EmuBackend(bin_shift=8, bin_writers=2, bin_slack_pct=0)
    f = be.filter(0, 1 << 14, 4, 11)
    seqs = ["A" * 3000, "ACGT" * 500]
    import numpy as np
    import _oracle as O
    b, off = O.as_batch(seqs)
    filt = np.zeros((1 << 14) // 8, np.uint8)
    assert f.insert(seqs) == oracle.bf_insert_seqs(filt, 1 << 14, 4, 11, b, off)
    assert np.array_equal(f.bytes(), filt)
    assert be.bin_overflow > 0


def test_binned_build_cfg1(oracle, golden):
    S.check_cfg1(EmuBackend(bin_shift=16, bin_writers=7), oracle, golden)
