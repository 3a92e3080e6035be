"""2-bit packed input (include/btlbf.h btlbf_pack_seqs / btlbf_*_seqs_packed*): the packer against a numpy
restatement of the reference's validity rule (vendor/nthash.hpp:189-228 seedTab, :51 code order), and the packed
entry points against the oracle -- i.e. the same results as the ASCII calls on the same sequences, including N runs,
lower case, U, short sequences and chunk halos.  CPU part: the host packer + the kernels' phase functions in the
emulator; GPU part: the product through the C ABI."""
import numpy as np
import pytest

import _oracle as O
import parity_suite as S
from _backends import EmuBackend


def np_pack(bases):
    """numpy restatement of the packed format: codes A0 C1 G2 T/U3 (4 per byte, base i at bits 2*(i&3)), invalid bit
    i&7 of byte i>>3 for every byte that is not one of ACGTUacgtu (an invalid base carries code 0)."""
    lut = np.full(256, 4, np.uint8)
    for ch, c in (("Aa", 0), ("Cc", 1), ("Gg", 2), ("TtUu", 3)):
        for x in ch:
            lut[ord(x)] = c
    cls = lut[bases]
    n = bases.size
    code = np.where(cls < 4, cls, 0).astype(np.uint8)
    pad = np.zeros((-n) % 4, np.uint8)
    c4 = np.concatenate([code, pad]).reshape(-1, 4)
    codes = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    invalid = np.packbits(cls >= 4, bitorder="little")
    return codes, invalid, int((cls >= 4).sum())


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 8, 9, 31, 4097, (1 << 20) + 13, (3 << 20) + 5])
def test_packer_matches_numpy_restatement(n):
    import btl_bloomfilter_b200 as B
    rng = np.random.default_rng(n)
    bases = rng.choice(np.frombuffer(b"ACGTacgtUuNnRYK-*.", np.uint8), size=n)
    pk = B.pack_seqs((bases, np.array([0, n], np.uint64)), keep_invalid=True)
    codes, invalid, bad = np_pack(bases)
    assert pk.n_invalid == bad
    assert np.array_equal(pk.codes, codes)
    assert np.array_equal(pk.invalid, invalid)
    for t in (1, 3):
        pk2 = B.pack_seqs((bases, np.array([0, n], np.uint64)), threads=t, keep_invalid=True)
        assert np.array_equal(pk2.codes, codes) and np.array_equal(pk2.invalid, invalid)


def test_packer_drops_the_plane_of_a_clean_batch_and_refuses_raw_bytes():
    import btl_bloomfilter_b200 as B
    pk = B.pack_seqs(["ACGTTGCA", "acgu"])
    assert pk.invalid is None and pk.n_bases == 12 and pk.n_invalid == 0
    # the raw bytes 1 3 4 5 7 hash in the reference (seedTab) but have no 2-bit form
    with pytest.raises(B.BtlbfError, match="no 2-bit form"):
        B.pack_seqs([b"ACG\x03ACGT"])


@pytest.fixture(scope="module")
def emu_packed():
    return EmuBackend(packed=True)


def test_emu_packed_golden(emu_packed, golden):
    S.check_golden_bf(emu_packed, golden)
    S.check_golden_cbf(emu_packed, golden)


@pytest.mark.parametrize("k,h,bits", [(25, 4, 1 << 16), (32, 6, 8 * 1237), (4, 5, 1024), (1, 2, 64)])
def test_emu_packed_random_bf(emu_packed, oracle, k, h, bits):
    S.check_random_bf(emu_packed, oracle, k, h, bits, seed=bits + k, p_n=0.02)


def test_emu_packed_multi_chunk_spaced_counting_and_edges(oracle):
    small = EmuBackend(packed=True, chunk=4096, batch=4096, resv_log2=10, list_log2=6)
    S.check_random_bf(small, oracle, 25, 4, 8 * 3001, seed=11, n_seqs=100, max_len=300)
    S.check_random_bf(small, oracle, 64, 2, 8 * 3001, seed=12, n_seqs=3, max_len=9000)
    S.check_random_cbf(small, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    S.check_random_spaced(EmuBackend(packed=True), oracle, 31, 2, 1, seed=31)
    S.check_edge_cases(EmuBackend(packed=True), oracle)
    S.check_random_bf(EmuBackend(packed=True, bin_shift=10), oracle, 25, 4, 1 << 16, seed=3)


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu_packed():
    from _backends import GpuBackend
    return GpuBackend(packed=True)


@pytest.mark.gpu
def test_gpu_packed_golden_and_cfg1(gpu_packed, oracle, golden):
    S.check_golden_bf(gpu_packed, golden)
    S.check_golden_cbf(gpu_packed, golden)
    S.check_cfg1(gpu_packed, oracle, golden)


@pytest.mark.gpu
@pytest.mark.parametrize("k,h,bits", [(25, 4, 1 << 16), (32, 6, 8 * 1237), (4, 5, 1024), (1, 2, 64), (100, 2, 1 << 20)])
def test_gpu_packed_random_bf(gpu_packed, oracle, k, h, bits):
    S.check_random_bf(gpu_packed, oracle, k, h, bits, seed=bits + k, p_n=0.02)


@pytest.mark.gpu
def test_gpu_packed_multi_chunk_spaced_counting_and_edges(oracle):
    from _backends import GpuBackend
    small = GpuBackend(packed=True, chunk=4096, batch=4096, resv_log2=10, list_log2=6, drain_threshold=16)
    S.check_random_bf(small, oracle, 25, 4, 8 * 3001, seed=11, n_seqs=100, max_len=300)
    S.check_random_bf(small, oracle, 64, 2, 8 * 3001, seed=12, n_seqs=3, max_len=9000)
    S.check_random_cbf(small, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    S.check_random_spaced(GpuBackend(packed=True), oracle, 31, 2, 1, seed=31)
    S.check_random_spaced(GpuBackend(packed=True), oracle, 16, 3, 3, seed=16)
    S.check_edge_cases(GpuBackend(packed=True), oracle)


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [8, 12, 20])
def test_gpu_packed_partitioned_paths(oracle, shift):
    """sort-bin build + partitioned query reading the packed planes (forced on a small filter)."""
    from _backends import GpuBackend
    be = GpuBackend(packed=True, bin_shift=shift, query_adaptive=0)
    S.check_random_bf(be, oracle, 25, 4, 1 << 22, seed=shift, n_seqs=300, max_len=2000, p_n=0.005)
    S.check_random_bf(be, oracle, 32, 6, 32 * 40009, seed=shift + 1, n_seqs=200, max_len=1500)


@pytest.mark.gpu
def test_gpu_packed_device_and_async_entry_points(oracle):
    """btlbf_*_seqs_packed_dev (device-resident planes) and the _async forms give the bytes of the ASCII calls."""
    import torch
    import btl_bloomfilter_b200 as B
    ctx = B.Context(0)
    rng = np.random.default_rng(5)
    b, off = S.rand_batch(rng, 500, 400, p_n=0.01)
    k, h, bits = 25, 4, 1 << 24
    filt = np.zeros(bits // 8, np.uint8)
    n_ref = oracle.bf_insert_seqs(filt, bits, h, k, b, off)
    e = oracle.bf_contains_seqs(filt, bits, h, k, b, off)
    pk = B.pack_seqs((b, off))
    assert pk.invalid is not None
    # device-resident planes, padded to whole 16 bytes
    def dev(a):
        t = torch.zeros((a.size + 15) // 16 * 16 + 16, dtype=torch.uint8, device="cuda:0")
        t[: a.size] = torch.from_numpy(a).to("cuda:0")
        return t
    d_codes, d_inv = dev(pk.codes), dev(pk.invalid)
    d_off = torch.from_numpy(off.view(np.int64)).to("cuda:0")
    d_stats = torch.zeros(4, dtype=torch.int64, device="cuda:0")
    words = (b.size + 31) // 32
    d_hit = torch.zeros(words + 8, dtype=torch.int32, device="cuda:0")
    d_valid = torch.zeros(words + 8, dtype=torch.int32, device="cuda:0")
    f = B.BloomFilter(bits, h, k, ctx=ctx)
    f.insertSeqsPackedDevice(d_codes.data_ptr(), d_inv.data_ptr(), b.size, d_off.data_ptr(), off.size - 1, d_stats.data_ptr())
    f.containsSeqsPackedDevice(d_codes.data_ptr(), d_inv.data_ptr(), b.size, d_off.data_ptr(), off.size - 1,
                               d_hit.data_ptr(), d_valid.data_ptr(), d_stats[2:].data_ptr())
    ctx.sync()
    st = d_stats.cpu().numpy()
    assert int(st[0]) == n_ref and (int(st[2]), int(st[3])) == e[:2]
    assert np.array_equal(f.to_numpy(), filt)
    nb = (b.size + 31) // 32 * 4
    assert np.array_equal(d_hit.cpu().numpy().view(np.uint8)[:nb], e[2])
    assert np.array_equal(d_valid.cpu().numpy().view(np.uint8)[:nb], e[3])
    # asynchronous host-buffer forms
    f2 = B.BloomFilter(bits, h, k, ctx=ctx)
    counts = np.zeros((2, 2), np.uint64)
    hits = np.zeros(nb, np.uint8)
    valid = np.zeros(nb, np.uint8)
    f2.insertSeqsPackedAsync(pk, counts[0])
    f2.containsSeqsPackedAsync(pk, hits, counts[1], valid_out=valid)
    ctx.sync()
    assert int(counts[0, 0]) == n_ref and (int(counts[1, 0]), int(counts[1, 1])) == e[:2]
    assert np.array_equal(f2.to_numpy(), filt) and np.array_equal(hits, e[2]) and np.array_equal(valid, e[3])


@pytest.mark.gpu
def test_gpu_host_side_packing_of_the_ascii_calls(oracle, golden):
    """Context option host_pack: the ASCII host-buffer calls pack every chunk on host threads before the H2D copy.  Same
    results as without it, including chunks that fall back to ASCII because they hold a raw byte 1 3 4 5 7."""
    from _backends import GpuBackend
    be = GpuBackend(host_pack=1, host_pack_threads=3)
    S.check_golden_bf(be, golden)
    S.check_golden_cbf(be, golden)
    S.check_cfg1(be, oracle, golden)
    small = GpuBackend(host_pack=1, chunk=4096, batch=4096, resv_log2=10, list_log2=6, drain_threshold=16)
    S.check_random_bf(small, oracle, 25, 4, 8 * 3001, seed=11, n_seqs=100, max_len=300)
    S.check_random_cbf(small, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    S.check_edge_cases(GpuBackend(host_pack=1), oracle)
    # exotic bytes: some chunks travel packed, some as ASCII
    rng = np.random.default_rng(9)
    b, off = S.rand_batch(rng, 60, 3000, exotic=0.0005)
    k, h, bits = 21, 3, 1 << 18
    f = small.filter(0, bits, h, k)
    filt = np.zeros(bits // 8, np.uint8)
    assert f.insert((b, off)) == oracle.bf_insert_seqs(filt, bits, h, k, b, off)
    assert np.array_equal(f.bytes(), filt)
    e = oracle.bf_contains_seqs(filt, bits, h, k, b, off)
    g = f.contains((b, off))
    assert e[:2] == g[:2] and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])
    # the partitioned paths on a larger batch (1 Mi-base chunks: the multi-threaded packer)
    big = GpuBackend(host_pack=1, bin_shift=16, query_adaptive=0, chunk=1 << 20)
    S.check_random_bf(big, oracle, 25, 4, 1 << 24, seed=5, n_seqs=600, max_len=8000, p_n=0.003)


def test_packer_property_arbitrary_bytes():
    """Any byte string: the packer either refuses it (exactly when it holds one of the raw bytes 1 3 4 5 7 that seedTab hashes
    but that have no 2-bit form) or produces the planes of the numpy restatement; the vector and the scalar code agree
    whatever the alignment of the input."""
    import btl_bloomfilter_b200 as B
    rng = np.random.default_rng(2026)
    raw = np.array([1, 3, 4, 5, 7], np.uint8)
    for trial in range(300):
        n = int(rng.integers(1, 700))
        kind = trial % 3
        if kind == 0:    # arbitrary bytes
            b = rng.integers(0, 256, n).astype(np.uint8)
        elif kind == 1:  # mostly bases, a few other bytes at or above 8
            b = rng.choice(np.frombuffer(b"ACGTacgtuUNn*-", np.uint8), size=n)
            b[rng.integers(0, n, max(1, n // 50))] = rng.integers(8, 256)
        else:            # bases with bytes 0 2 6 (below 8, but not raw bases: plain invalid)
            b = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
            b[rng.integers(0, n, max(1, n // 40))] = rng.choice(np.array([0, 2, 6], np.uint8))
        lead = int(rng.integers(0, 9))  # misalign the input pointer
        buf = np.concatenate([np.zeros(lead, np.uint8), b])
        view = buf[lead:]
        off = np.array([0, n], np.uint64)
        if np.isin(b, raw).any():
            with pytest.raises(B.BtlbfError, match="no 2-bit form"):
                B.pack_seqs((view, off))
            continue
        pk = B.pack_seqs((view, off), keep_invalid=True)
        codes, invalid, bad = np_pack(b)
        assert pk.n_invalid == bad and np.array_equal(pk.codes, codes) and np.array_equal(pk.invalid, invalid), trial
