"""Parity tests proper: the CUDA path, called through the C ABI (libbtlbf_cuda.so via the Python host
classes), against the oracle on the same seeded inputs and against the golden vectors generated from the
reference.  Bit-exact.  Needs a B200."""
import hashlib
import os

import numpy as np
import pytest

import _oracle as O
import parity_suite as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from _backends import GpuBackend
    return GpuBackend()


@pytest.fixture(scope="module")
def gpu_small():
    from _backends import GpuBackend
    return GpuBackend(chunk=4096, batch=4096, resv_log2=10, list_log2=6, drain_threshold=16)


def test_library_loaded_is_in_tree():
    import btl_bloomfilter_b200 as B
    L = B.lib()
    assert L.btlbf_version() >= 100
    assert os.path.dirname(B._build.LIB).endswith("btl_bloomfilter_b200")


def test_golden_hashes(gpu, golden):
    S.check_golden_hashes(gpu, golden)


def test_golden_hashes_generic_path(golden):
    from _backends import GpuBackend
    S.check_golden_hashes(GpuBackend(force_generic=1), golden)


def test_golden_bf(gpu, golden):
    S.check_golden_bf(gpu, golden)


def test_golden_bf_early_exit_mode(golden):
    from _backends import GpuBackend
    S.check_golden_bf(GpuBackend(query_mode=1), golden)


def test_golden_cbf(gpu, golden):
    S.check_golden_cbf(gpu, golden)


def test_golden_small_tables(gpu_small, golden):
    S.check_golden_cbf(gpu_small, golden)
    S.check_golden_bf(gpu_small, golden)


@pytest.mark.parametrize("k,h", [(1, 1), (2, 3), (4, 5), (25, 4), (31, 2), (32, 6), (33, 3), (64, 4), (100, 2)])
def test_random_hashes(gpu, oracle, k, h):
    S.check_random_hashes(gpu, oracle, k, h, seed=1000 * k + h)


def test_random_hashes_exotic_bytes(gpu, oracle):
    S.check_random_hashes(gpu, oracle, 5, 3, seed=77, exotic=0.05)
    S.check_random_hashes(gpu, oracle, 25, 4, seed=78, exotic=0.01)


def test_random_hashes_multi_tile(gpu_small, oracle):
    S.check_random_hashes(gpu_small, oracle, 25, 4, seed=5, n_seqs=120, max_len=400)
    S.check_random_hashes(gpu_small, oracle, 64, 2, seed=6, n_seqs=3, max_len=9000)


@pytest.mark.parametrize("k,n_seeds,h2", [(5, 2, 2), (31, 2, 1), (16, 3, 3), (40, 1, 4)])
def test_random_spaced(gpu, oracle, k, n_seeds, h2):
    S.check_random_spaced(gpu, oracle, k, n_seeds, h2, seed=k)


@pytest.mark.parametrize("k,h,bits", [(25, 4, 1 << 16), (32, 6, 8 * 1237), (4, 5, 1024), (21, 3, 8 * 4099)])
def test_random_bf(gpu, oracle, k, h, bits):
    S.check_random_bf(gpu, oracle, k, h, bits, seed=bits + k)


def test_random_bf_multi_chunk(gpu_small, oracle):
    S.check_random_bf(gpu_small, oracle, 25, 4, 8 * 3001, seed=11, n_seqs=100, max_len=300)


@pytest.mark.parametrize("k,h,m", [(25, 4, 4096), (8, 5, 100008), (5, 3, 64), (11, 4, 512)])
def test_random_cbf(gpu, oracle, k, h, m):
    S.check_random_cbf(gpu, oracle, k, h, m, seed=m + k)


def test_random_cbf_host_driven_rounds(oracle, golden):
    """the non-cooperative residual path (host-driven rounds + single-CTA drain)"""
    from _backends import GpuBackend
    be = GpuBackend(chunk=4096, batch=4096, resv_log2=10, list_log2=6, drain_threshold=16, ordered_coop=0)
    S.check_random_cbf(be, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    S.check_golden_cbf(be, golden)
    S.check_golden_bf(be, golden)


def test_random_cbf_ungrouped_commit(oracle, golden):
    """pass 2 of the ordered updates one window at a time (the form spaced seeds and h > 8 use)"""
    from _backends import GpuBackend
    be = GpuBackend(chunk=4096, batch=4096, resv_log2=12, list_log2=8, ungrouped_commit=1)
    S.check_random_cbf(be, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    S.check_random_cbf(be, oracle, 25, 6, 100_000, seed=2)
    S.check_golden_cbf(be, golden)
    S.check_golden_bf(be, golden)


@pytest.mark.parametrize("h", [1, 3, 5, 8, 11])
def test_random_cbf_every_group_width(gpu_small, gpu, oracle, h):
    """grouped commit pass for each compile-time h (and the general form beyond 8)"""
    S.check_random_cbf(gpu_small, oracle, 13, h, 40_000, seed=h)
    S.check_random_cbf(gpu, oracle, 21, h, 1 << 20, seed=10 + h, n_seqs=60, max_len=2000)


def test_random_cbf_small_tables(gpu_small, oracle):
    S.check_random_cbf(gpu_small, oracle, 9, 4, 256, seed=1, n_seqs=80, max_len=200)
    f = gpu_small.filter(1, 64, 3, 5)
    f.insert(["ACGTA" * 40] * 8)
    d, r = f.ordered_stats()
    assert d > 0 and r > 1


def test_edge_cases(gpu, oracle):
    S.check_edge_cases(gpu, oracle)


def test_cfg1(gpu, oracle, golden):
    S.check_cfg1(gpu, oracle, golden)


# ---------------------------------------------------------------- larger seeded runs against the oracle
def test_counting_build_200k(gpu, oracle):
    """CountingBloomFilter<uint8_t> build at a size where thousands of k-mers collide inside a batch:
    counter values must equal the single-threaded reference-order result."""
    g = oracle.synth_genome(0, 200_000, 42)
    reads = oracle.synth_reads(0, 1500, 150, g.size, 42, 9)  # overlapping reads: repeated k-mers
    b = np.concatenate([g, reads])
    off = np.concatenate([[0, g.size], g.size + 150 * np.arange(1, 1501)]).astype(np.uint64)
    m = 400_000
    f = gpu.filter(1, m, 4, 25, thr=2)
    cnt = np.zeros(m, np.uint8)
    assert f.insert((b, off)) == oracle.cbf_insert_seqs(cnt, m, 4, 25, b, off)
    assert np.array_equal(f.bytes(), cnt)
    e = oracle.cbf_contains_seqs(cnt, m, 4, 25, 2, reads, (150 * np.arange(1501)).astype(np.uint64))
    gq = f.contains((reads, (150 * np.arange(1501)).astype(np.uint64)))
    assert e[:2] == gq[:2] and np.array_equal(e[2], gq[2])


def test_bf_non_pow2_large_modulus(gpu, oracle):
    """cfg2's filter size (31,568,113,856 bits) is not a power of two: the exact 64-bit modulo must agree.
    Only the touched words are compared (the oracle cannot hold 4 GB cheaply): query parity + popcount."""
    bits = 31_568_113_856
    g = oracle.synth_genome(0, 300_000, 42)
    off = np.array([0, g.size], np.uint64)
    f = gpu.filter(0, bits, 4, 25)
    n = f.insert((g, off))
    assert n == g.size - 24
    # expected set of bit indices from the oracle's hashes
    _, hs, _ = oracle.hash_seqs(4, 25, g, off)
    idx = np.unique(hs[: g.size - 24].reshape(-1) % np.uint64(bits))
    assert f.f.getPop() == idx.size
    nq, nh, hits, valid = f.contains((g, off))
    assert nq == nh == g.size - 24
    miss = oracle.synth_genome(0, 100_000, 43 << 40)
    nq, nh, hits, _ = f.contains((miss, np.array([0, miss.size], np.uint64)))
    _, mh, _ = oracle.hash_seqs(4, 25, miss, np.array([0, miss.size], np.uint64))
    exp = np.isin(mh[: miss.size - 24] % np.uint64(bits), idx).all(axis=1)
    assert nh == int(exp.sum())
    assert np.array_equal(O.bits_to_bool(hits, miss.size)[: miss.size - 24], exp)


def test_full_size_properties(gpu, oracle):
    """Size-independent properties at a BASELINE-scale shape (4 GiB filter, 150 bp reads, k=32, h=6):
    inserted => contained; insert is idempotent (popcount unchanged by re-inserting); the count of
    valid k-mers equals reads * (150 - 32 + 1); hits on an unrelated sequence match a popcount-based FPR."""
    import btl_bloomfilter_b200 as B
    bits = 1 << 35
    f = B.BloomFilter(bits, 6, 32, ctx=gpu.ctx)
    reads = oracle.synth_reads(0, 20000, 150, 3_000_000_000, 42, 1)
    off = (150 * np.arange(20001)).astype(np.uint64)
    n = f.insertSeqs((reads, off))
    assert n == 20000 * 119
    pop = f.getPop()
    assert 0 < pop <= 6 * n
    r = f.containsSeqs((reads, off))
    assert r.n_kmers == n and r.n_hits == n
    assert f.insertSeqs((reads, off)) == n and f.getPop() == pop
    miss = oracle.synth_genome(0, 1_000_000, 43 << 40)
    r = f.containsSeqs((miss, np.array([0, miss.size], np.uint64)))
    assert r.n_hits <= 2  # occupancy ~3e-4 -> FPR ~1e-21
    del f


# ---------------------------------------------------------------- file layout, both directions
def test_store_files_match_reference_bytes(gpu, golden, tmp_path):
    import btl_bloomfilter_b200 as B
    for i, c in enumerate(golden["bf_cases"]):
        f = B.BloomFilter(c["bits"], c["h"], c["k"], ctx=gpu.ctx)
        f.insertSeqs(c["seqs"])
        p = tmp_path / ("bf%d.bf" % i)
        f.storeFilter(str(p))
        raw = p.read_bytes()
        assert hashlib.md5(raw).hexdigest() == c["file_md5"]
        assert raw.startswith(c["header"].encode())
        g = B.BloomFilter(str(p), ctx=gpu.ctx)
        assert (g.getFilterSize(), g.getHashNum(), g.getKmerSize()) == (c["bits"], c["h"], c["k"])
        assert g.to_numpy().tobytes().hex() == c["filter_hex"]
        assert g.getPop() == c["pop"]
    for i, c in enumerate(golden["cbf_cases"]):
        f = B.CountingBloomFilter(c["size"], c["h"], c["k"], c["threshold"], ctx=gpu.ctx)
        assert f.size() == c["size_rounded"] == f.sizeInBytes()
        f.insertSeqs(c["seqs"])
        p = tmp_path / ("cbf%d.bf" % i)
        f.storeFilter(str(p))
        assert hashlib.md5(p.read_bytes()).hexdigest() == c["file_md5"]
        g = B.CountingBloomFilter(str(p), c["threshold"], ctx=gpu.ctx)
        assert g.popCount() == c["popcount"] and g.filtered_popcount() == c["filtered_popcount"]


def test_files_interoperate_with_reference(gpu, ref, oracle, tmp_path):
    """GPU-written file loaded by the reference, reference-written file loaded by the GPU class."""
    import btl_bloomfilter_b200 as B
    rng = np.random.default_rng(8)
    b, off = S.rand_batch(rng, 30, 200)
    f = B.BloomFilter(8 * 5000, 4, 25, ctx=gpu.ctx)
    f.setnEntry(123)
    f.settEntry(456)
    f.insertSeqs((b, off))
    p = str(tmp_path / "gpu.bf")
    f.storeFilter(p)
    rf = ref.L.ref_bf_load(p.encode())
    assert np.array_equal(ref.bf_bytes(rf), f.to_numpy())
    ne, te = O.u64(), O.u64()
    ref.L.ref_bf_get_meta(rf, O.C.byref(ne), O.C.byref(te))
    assert (ne.value, te.value) == (123, 456)
    p2 = str(tmp_path / "ref.bf")
    ref.L.ref_bf_set_meta(rf, 0.015625, 7, 9)
    ref.L.ref_bf_store(rf, p2.encode())
    g = B.BloomFilter(p2, ctx=gpu.ctx)
    assert np.array_equal(g.to_numpy(), f.to_numpy())
    assert (g.m_dFPR, g.getnEntry(), g.gettEntry()) == (0.015625, 7, 9)
    g.storeFilter(str(tmp_path / "again.bf"))
    assert open(p2, "rb").read() == open(str(tmp_path / "again.bf"), "rb").read()
    ref.L.ref_bf_free(rf)
    # counting
    cf = B.CountingBloomFilter(3001, 3, 17, 2, ctx=gpu.ctx)
    cf.insertSeqs((b, off))
    p3 = str(tmp_path / "gpu.cbf")
    cf.storeFilter(p3)
    rc = ref.L.ref_cbf_load(p3.encode(), 2)
    assert np.array_equal(ref.cbf_bytes(rc), cf.to_numpy())
    p4 = str(tmp_path / "ref.cbf")
    ref.L.ref_cbf_store(rc, p4.encode())
    assert open(p3, "rb").read() == open(p4, "rb").read()
    ref.L.ref_cbf_free(rc)


# ---------------------------------------------------------------- legacy per-k-mer interface (reference unit tests)
def test_reference_unit_scenario_bloom(gpu, oracle, tmp_path):
    """Tests/Unit/BloomFilterTests.cpp:53-146 against the GPU class: 1 Gbit filter, h=5, k=4, "ACGTAC"."""
    import btl_bloomfilter_b200 as B
    filterSize, numHashes, k = 1000000000, 5, 4
    f = B.BloomFilter(filterSize, numHashes, k, ctx=gpu.ctx)
    seq = "ACGTAC"
    b, off = O.as_batch([seq])
    _, hs, _ = oracle.hash_seqs(numHashes, k, b, off)
    for p in range(3):
        f.insert(hs[p])
    for p in range(3):
        assert f.contains(hs[p])
    assert f.getPop() == len(np.unique(hs[:3].reshape(-1) % np.uint64(filterSize)))
    p = str(tmp_path / "unit.bf")
    f.storeFilter(p)
    raw = open(p, "rb").read()
    end = raw.index(b"[HeaderEnd]\n") + len(b"[HeaderEnd]\n")
    assert len(raw) - end == f.sizeInBytes() == filterSize // 8
    g = B.BloomFilter(p, ctx=gpu.ctx)
    for q in range(3):
        assert g.contains(hs[q])
    assert g.insertAndCheck(hs[0]) is True
    assert g.insertAndCheck(np.array([1, 2, 3, 4, 5], np.uint64)) is False
    assert g.insertAndCheck(np.array([1, 2, 3, 4, 5], np.uint64)) is True


def test_reference_unit_scenario_counting(gpu, oracle, tmp_path):
    """Tests/Unit/CountingBloomFilterTests.cpp:54-246 for uint8_t: 100001 bytes (-> 100008), h=5, k=8."""
    import btl_bloomfilter_b200 as B
    f = B.CountingBloomFilter(100001, 5, 8, 1, ctx=gpu.ctx)
    assert f.size() == f.sizeInBytes() == 100008
    seq = "ACGTACACTGGACTGAGTCT"
    b, off = O.as_batch([seq])
    n, hs, _ = oracle.hash_seqs(5, 8, b, off)
    for p in range(n):
        f.insert(hs[p])
    for p in range(n):
        assert f.contains(hs[p])
    rng = np.random.default_rng(0)
    rb, ro = O.as_batch(["".join(rng.choice(list("ACGT"), size=60).tolist())])
    rn, rh, _ = oracle.hash_seqs(5, 8, rb, ro)
    cnt = np.zeros(100008, np.uint8)
    oracle.cbf_insert_seqs(cnt, 100008, 5, 8, b, off)
    assert np.array_equal(f.to_numpy(), cnt)
    exp = np.array([cnt[rh[p] % np.uint64(100008)].min() >= 1 for p in range(rn)])
    assert np.array_equal(f.contains(rh[:rn]), exp)
    assert np.array_equal(f.minCount(hs[:n]), np.array([cnt[hs[p] % np.uint64(100008)].min() for p in range(n)]))
    p = str(tmp_path / "unit.cbf")
    f.storeFilter(p)
    g = B.CountingBloomFilter(p, 1, ctx=gpu.ctx)
    assert g.size() == g.sizeInBytes() == 100008 and np.array_equal(g.to_numpy(), cnt)
    # incrementAll + saturation
    h0 = hs[:1]
    for _ in range(300):
        g.incrementAll(h0)
    assert g.minCount(h0[0]) == 255


def test_errors_are_reported_not_fatal(gpu):
    import btl_bloomfilter_b200 as B
    with pytest.raises(ValueError):
        B.BloomFilter(1001, 4, 5, ctx=gpu.ctx)  # "Filter Size ... is not a multiple of 8"
    with pytest.raises(B.BtlbfError):
        B.BloomFilter("/nonexistent/file.bf", ctx=gpu.ctx)
    f = B.BloomFilter(1024, 4, 5, ctx=gpu.ctx)
    with pytest.raises(B.BtlbfError):
        f.setSeeds(["11011", "1101"], 2)  # wrong seed length
    with pytest.raises(B.BtlbfError):
        f.setSeeds(["11011"], 2)  # hashNum != n_seeds*h2


# ---------------------------------------------------------------- partitioned (binned) BloomFilter build
@pytest.mark.parametrize("legacy", [0, 1])
@pytest.mark.parametrize("shift", [8, 12, 20])
def test_binned_build_equals_oracle(oracle, golden, shift, legacy):
    """legacy=0: the sort-bin kernel (sort_bin.cuh) wherever the shape allows; 1: the general-shape kernels"""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=shift, bin_kernel=legacy)
    S.check_golden_bf(be, golden)
    S.check_random_bf(be, oracle, 25, 4, 1 << 16, seed=3)
    S.check_random_bf(be, oracle, 32, 6, 32 * 1237, seed=4)
    S.check_random_spaced(be, oracle, 16, 3, 3, seed=16)
    S.check_cfg1(be, oracle, golden)


@pytest.mark.parametrize("legacy", [0, 1])
def test_binned_build_skewed_input(oracle, legacy):
    """sub-buckets without slack + repetitive input: the overflow path (direct atomics / direct probes)"""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=8, bin_slack_pct=0, bin_kernel=legacy)
    f = be.filter(0, 1 << 14, 4, 11)
    seqs = ["A" * 30000, "ACGT" * 5000, "ACGTTGCA" * 3000]
    b, off = O.as_batch(seqs)
    filt = np.zeros((1 << 14) // 8, np.uint8)
    assert f.insert(seqs) == oracle.bf_insert_seqs(filt, 1 << 14, 4, 11, b, off)
    assert np.array_equal(f.bytes(), filt)
    q = seqs + ["ACGTTGCATTGACCA" * 2000, "C" * 20000]
    qb, qoff = O.as_batch(q)
    nk, nh, hit, valid = f.contains(q)
    onk, onh, ohit, ovalid = oracle.bf_contains_seqs(filt, 1 << 14, 4, 11, qb, qoff)
    assert (nk, nh) == (onk, onh)
    assert np.array_equal(hit, ohit) and np.array_equal(valid, ovalid)


@pytest.mark.parametrize("accum_bytes", [1 << 33, 1 << 16, 0])
def test_binned_build_accumulates_across_calls(oracle, accum_bytes):
    """Pass 1 of successive insert calls appends to the same sub-buckets; pass 2 runs when they are full or the
    filter is read.  Two filters are interleaved, one of them re-hashed with spaced seeds half way."""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=10, bin_accum_bytes=accum_bytes)
    bits, h, k = 1 << 18, 4, 21
    f1, f2 = be.filter(0, bits, h, k), be.filter(0, 3 * 32 * 1031, 3, 15)
    a1, a2 = np.zeros(bits // 8, np.uint8), np.zeros(3 * 32 * 1031 // 8, np.uint8)
    rng = np.random.default_rng(11)
    for i in range(7):
        b, off = S.rand_batch(rng, 5 + i, 3000 + 500 * i, p_n=0.002)
        assert f1.insert((b, off)) == oracle.bf_insert_seqs(a1, bits, h, k, b, off)
        if i % 3 == 2:
            assert f2.insert((b, off)) == oracle.bf_insert_seqs(a2, 3 * 32 * 1031, 3, 15, b, off)
        if i == 4:
            assert np.array_equal(f1.bytes(), a1)  # a read in the middle settles the accumulation
    assert np.array_equal(f1.bytes(), a1)
    assert np.array_equal(f2.bytes(), a2)
    qb, qoff = S.rand_batch(rng, 6, 4000)
    nk, nh, hit, valid = f1.contains((qb, qoff))
    onk, onh, ohit, ovalid = oracle.bf_contains_seqs(a1, bits, h, k, qb, qoff)
    assert (nk, nh) == (onk, onh) and np.array_equal(hit, ohit) and np.array_equal(valid, ovalid)


def test_binned_build_full_size_equals_direct_build(gpu, oracle):
    """cfg2's filter (31,568,113,856 bits, not a power of two), 24 Mbp of the synthetic genome generated in
    HBM: the partitioned build (auto-selected at this size) and the direct-atomics build must produce the
    same 3.95 GB array, and a prefix replayed by the oracle must be fully contained."""
    import torch
    import btl_bloomfilter_b200 as B
    from btl_bloomfilter_b200 import parallel
    bits, n = 31_568_113_856, 24_000_000
    ctx = gpu.ctx
    dev = torch.device("cuda", 0)
    g = torch.empty(n + 64, dtype=torch.uint8, device=dev)
    ctx.synth_genome_device(g.data_ptr(), 5_000_000, n, 42)
    off = torch.tensor([0, n], dtype=torch.int64, device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)
    arrays = []
    for mode in (-1, 1, 0):
        ctx.set_option("bin_mode", mode)
        f = B.BloomFilter(bits, 4, 25, ctx=ctx)
        l0 = ctx.launch_count
        f.insertSeqsDevice(g.data_ptr(), n, off.data_ptr(), 1, stats.data_ptr())
        ctx.sync()
        assert ctx.launch_count - l0 == (1 if mode == -1 else 2)  # auto picks the two-pass build here
        ptr, nbytes = f.device_ptr()
        arrays.append((f, parallel.device_tensor_from_ptr(ptr, nbytes, dev)))
    ctx.set_option("bin_mode", 0)
    assert int(stats[0]) == 3 * (n - 24)
    assert torch.equal(arrays[0][1], arrays[1][1]) and torch.equal(arrays[0][1], arrays[2][1])
    head = oracle.synth_genome(5_000_000, 200_000, 42)
    assert np.array_equal(head, g[:200_000].cpu().numpy())
    r = arrays[1][0].containsSeqs((head, np.array([0, head.size], np.uint64)))
    assert r.n_kmers == r.n_hits == head.size - 24


def test_async_host_calls_match_blocking_calls(gpu, oracle):
    """btlbf_insert_seqs_async / btlbf_contains_seqs_async queued back to back (more calls than tickets),
    results read after one btlbf_ctx_sync: same filter, hit vectors and counts as the oracle."""
    import btl_bloomfilter_b200 as B
    rng = np.random.default_rng(21)
    bits, h, k = 8 * 50_021, 3, 21
    f = B.BloomFilter(bits, h, k, ctx=gpu.ctx)
    filt = np.zeros(bits // 8, np.uint8)
    batches = [S.rand_batch(rng, 30, 400) for _ in range(7)]
    counts = np.zeros((14, 2), np.uint64)
    hits = [np.zeros(O.nbits_bytes(b.size), np.uint8) for b, _ in batches]
    keep = []
    for i, (b, off) in enumerate(batches):
        keep.append(f.insertSeqsAsync((b, off), counts[2 * i]))
        keep.append(f.containsSeqsAsync((b, off), hits[i], counts[2 * i + 1]))
    gpu.ctx.sync()
    for i, (b, off) in enumerate(batches):
        n = oracle.bf_insert_seqs(filt, bits, h, k, b, off)
        assert int(counts[2 * i, 0]) == n
        assert (int(counts[2 * i + 1, 0]), int(counts[2 * i + 1, 1])) == (n, n)
        assert np.array_equal(O.bits_to_bool(hits[i], b.size), O.bits_to_bool(oracle.hash_seqs(h, k, b, off)[2], b.size))
    assert np.array_equal(f.to_numpy(), filt)


# ---------------------------------------------------------------- BASELINE-size property tests (cfg4, cfg5)
def test_cfg4_counting_filter_full_size_properties(gpu, oracle):
    """CountingBloomFilter<uint8_t> with 16e9 counters (BASELINE configs[3]): exact counts on a sample the
    oracle can replay sparsely (every touched counter), threshold-2 query semantics, idempotent queries."""
    import btl_bloomfilter_b200 as B
    m, h, k = 16_000_000_000, 4, 25
    f = B.CountingBloomFilter(m, h, k, 2, ctx=gpu.ctx)
    assert f.size() == f.sizeInBytes() == m
    reads = oracle.synth_reads(0, 4000, 150, 3_000_000_000, 42, 11)
    off = (150 * np.arange(4001)).astype(np.uint64)
    n = f.insertSeqs((reads, off))
    assert n == 4000 * 126
    r1 = f.containsSeqs((reads, off))
    assert r1.n_kmers == n
    # sparse replay: sequential min-increment over a dict of touched counters
    _, hs, valid = oracle.hash_seqs(h, k, reads, off)
    idx = (hs % np.uint64(m))
    vmask = O.bits_to_bool(valid, reads.size)
    cnt = {}
    for p in np.nonzero(vmask)[0]:
        slots = [int(x) for x in idx[p]]
        mn = min(cnt.get(s_, 0) for s_ in slots)
        if mn < 255:
            for s_ in slots:
                if cnt.get(s_, 0) == mn:
                    cnt[s_] = mn + 1
    exp_min = np.array([min(cnt.get(int(x), 0) for x in idx[p]) if vmask[p] else 0 for p in range(reads.size)], np.uint8)
    q = f.minCountSeqs((reads, off))
    assert np.array_equal(q.counts, exp_min)
    assert r1.n_hits == int((exp_min[vmask] >= 2).sum())
    assert f.popCount() == len(cnt)
    n2 = f.insertSeqs((reads, off))
    q2 = f.minCountSeqs((reads, off))
    assert n2 == n and (q2.counts[vmask] >= 2).all()  # second pass: every k-mer now has count >= 2
    r2 = f.containsSeqs((reads, off))
    assert r2.n_hits == n
    del f


@pytest.mark.parametrize("bits", [1 << 29, 1 << 37])
def test_cfg5_spaced_seed_filters_properties(gpu, oracle, bits):
    """stHashIterator, k=31, two symmetric seeds (BASELINE configs[4]) on the L2-resident 64 MB filter and the
    16 GB filter: inserted => contained, bit positions equal the oracle's hashes, idempotent inserts."""
    import btl_bloomfilter_b200 as B
    left = ["111101110111001", "111110110100111"]
    seeds = [x + "1" + x[::-1] for x in left]
    assert all(s == s[::-1] and len(s) == 31 for s in seeds)
    k = 31
    f = B.BloomFilter(bits, 2, k, ctx=gpu.ctx)
    f.setSeeds(seeds, 1)
    reads = oracle.synth_reads(0, 3000, 150, 3_000_000_000, 42, 13)
    off = (150 * np.arange(3001)).astype(np.uint64)
    n = f.insertSeqs((reads, off))
    assert n == 3000 * 120
    _, hs, _, valid = oracle.st_hash_seqs(seeds, 1, k, reads, off)
    vm = O.bits_to_bool(valid, reads.size)
    idx = np.unique((hs[vm] % np.uint64(bits)).reshape(-1))
    assert f.getPop() == idx.size
    r = f.containsSeqs((reads, off))
    assert r.n_kmers == r.n_hits == n
    assert f.insertSeqs((reads, off)) == n and f.getPop() == idx.size
    miss = oracle.synth_genome(0, 200_000, 43 << 40)
    r = f.containsSeqs((miss, np.array([0, miss.size], np.uint64)))
    _, mh, _, mv = oracle.st_hash_seqs(seeds, 1, k, miss, np.array([0, miss.size], np.uint64))
    mvm = O.bits_to_bool(mv, miss.size)
    exp = np.isin(mh % np.uint64(bits), idx).all(axis=1) & mvm
    assert r.n_hits == int(exp.sum())
    assert np.array_equal(O.bits_to_bool(r.hit_bits, miss.size), exp)
    del f


def test_kmer_bloom_filter_matches_reference(gpu, golden):
    """KmerBloomFilter::insert/contains(const char*) (the class SWIG exports as BloomFilter)."""
    import btl_bloomfilter_b200 as B
    for c in golden["kmer_bloom_filter"]:
        f = B.KmerBloomFilter(c["bits"], c["h"], c["k"], ctx=gpu.ctx)
        for km in c["kmers"][:10]:
            f.insert(km)                      # one k-mer per call, as the Perl callers do
        f.insertSeqs(c["kmers"][10:])         # the rest in one batch
        assert f.to_numpy().tobytes().hex() == c["filter_hex"]
        assert [int(f.contains(p)) for p in c["probes"]] == c["contains"]
        assert f.contains("N" * c["k"]) is False


# ---------------------------------------------------------------- MIBF level-1 bit vector
@pytest.mark.parametrize("bits,forced", [(1_000_003, 0), (8 * 300_007, 1), (64 * 4099 + 17, 1)])
def test_mibf_bit_vector_build(oracle, bits, forced):
    """BTLBF_BITVECTOR: insertBV / insertBVColli (MIBFConstructSupport.hpp:55-87) on an sdsl-style bit vector of
    any size; words, k-mer counts and collision counts against the oracle's restatement.  forced=1 runs the
    partitioned build and query on a size that is not a multiple of 32."""
    import btl_bloomfilter_b200 as B
    ctx = B.Context(0)
    if forced:
        ctx.set_option("bin_mode", 1)
        ctx.set_option("bin_query_mode", 1)
        ctx.set_option("bin_part_log2", 12)
    h, k = 3, 19
    rng = np.random.default_rng(bits % 1000)
    bv = B.BitVector(bits, h, k, ctx=ctx)
    words = np.zeros((bits + 63) // 64, np.uint64)
    b1, o1 = S.rand_batch(rng, 30, 900, p_n=0.01)
    assert bv.insertSeqs((b1, o1)) == oracle.mibf_insert_bv_seqs(words, bits, h, k, b1, o1)[0]
    assert np.array_equal(bv.to_numpy(), words.view(np.uint8))
    # second batch shares k-mers with the first: collisions
    b2 = np.concatenate([b1[: b1.size // 2], S.rand_batch(rng, 10, 500)[0]])
    o2 = np.array([0, b2.size], np.uint64)
    assert bv.insertBVColli((b2, o2)) == oracle.mibf_insert_bv_seqs(words, bits, h, k, b2, o2)
    assert np.array_equal(bv.to_numpy(), words.view(np.uint8))
    assert bv.getPop() == int(np.unpackbits(words.view(np.uint8)).sum())
    r = bv.containsSeqs((b1, o1))
    assert r.n_hits == r.n_kmers > 0
    filt = words.view(np.uint8)
    q, qo = S.rand_batch(rng, 20, 700)
    r = bv.containsSeqs((q, qo))
    nk, nh, hit, valid = oracle.bf_contains_seqs(filt, bits, h, k, q, qo)
    assert (r.n_kmers, r.n_hits) == (nk, nh) and np.array_equal(r.hit_bits, hit)


@pytest.mark.parametrize("pct", [0, 20, 101])
def test_adaptive_query_paths_agree(oracle, pct):
    """Partitioned query with device-side path selection: pct=0 always keeps the partitioned kernels, 101 always
    hands the batch to the early-exit kernel, 20 decides from the sampled hit fraction; hit-rich and miss-rich
    read sets must give the oracle's bits on every setting."""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=12, query_adaptive=1, query_adaptive_pct=pct, query_adaptive_min_tiles=1)
    bits, h, k = 1 << 20, 4, 25
    rng = np.random.default_rng(77)
    b, off = S.rand_batch(rng, 30, 3000, p_n=0.005)
    f = be.filter(0, bits, h, k)
    filt = np.zeros(bits // 8, np.uint8)
    assert f.insert((b, off)) == oracle.bf_insert_seqs(filt, bits, h, k, b, off)
    miss = S.rand_batch(rng, 40, 2500)
    for (x, xo) in ((b, off), miss, (np.concatenate([b, miss[0]]), np.concatenate([off, off[-1] + miss[1][1:]]))):
        nk, nh, hit, valid = f.contains((x, xo))
        e = oracle.bf_contains_seqs(filt, bits, h, k, x, xo)
        assert (nk, nh) == e[:2] and np.array_equal(hit, e[2]) and np.array_equal(valid, e[3])


def test_clear_and_upload_drop_parked_kmers(oracle):
    """k-mers parked in the partition buckets (pass 1 done, pass 2 pending) must not survive a clear() or a
    whole-array upload that logically follows them"""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=10)
    bits, h, k = 1 << 18, 4, 21
    rng = np.random.default_rng(5)
    b, off = S.rand_batch(rng, 8, 3000)
    f = be.filter(0, bits, h, k)
    f.insert((b, off))
    f.f.clear()
    assert not f.bytes().any()
    f.insert((b, off))
    other = rng.integers(0, 256, bits // 8).astype(np.uint8)
    f.set_bytes(other)
    assert np.array_equal(f.bytes(), other)
    filt = other.copy()
    assert f.insert((b, off)) == oracle.bf_insert_seqs(filt, bits, h, k, b, off)
    assert np.array_equal(f.bytes(), filt)


@pytest.mark.parametrize("shift", [8, 20, 24])
def test_two_level_pass2_equals_oracle(oracle, golden, shift):
    """pass 2 of the partitioned build through apply2.cu (refine by slice, OR in shared memory), forced for
    small accumulations; includes a skewed input whose level-2 buckets overflow into direct atomics"""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=shift, bin_two_level=1, bin_two_level_min=0)
    S.check_golden_bf(be, golden)
    S.check_random_bf(be, oracle, 25, 4, 1 << 26, seed=3, n_seqs=60, max_len=4000)
    S.check_random_bf(be, oracle, 32, 6, 32 * 1237 * 64 + 8, seed=4)
    S.check_cfg1(be, oracle, golden)
    assert be.ctx.two_level_passes > 0
    f = be.filter(0, 1 << 22, 4, 11)
    seqs = ["A" * 30000, "ACGT" * 5000, "ACGTTGCA" * 3000]
    b, off = O.as_batch(seqs)
    filt = np.zeros((1 << 22) // 8, np.uint8)
    assert f.insert(seqs) == oracle.bf_insert_seqs(filt, 1 << 22, 4, 11, b, off)
    assert np.array_equal(f.bytes(), filt)


def test_per_kmer_updates_from_threads_and_in_order(gpu, oracle):
    """The legacy per-k-mer interface (`bloom.insert(*itr)`, README.md:30-43): (a) eight host threads insert into one
    shared BloomFilter concurrently (the reference's OpenMP pattern, Tests/AdHoc/ParallelFilter.cpp:104-122) -- same bytes
    as the batched build; (b) one thread feeds a CountingBloomFilter k-mer by k-mer, across the boundaries of the
    per-thread (512) and per-filter (65,536) queues, interleaved with reads -- the order-dependent incrementMin result
    of the sequential reference loop (CountingBloomFilter.hpp:134-162), and reads see everything queued before them."""
    import threading
    B = gpu.B
    k, h, bits = 25, 4, 1 << 22
    g = oracle.synth_genome(0, 90_000, 42)
    off = np.array([0, g.size], np.uint64)
    n, hs, valid = oracle.hash_seqs(h, k, g, off)
    hs = np.ascontiguousarray(hs[: g.size - k + 1])
    filt = np.zeros(bits // 8, np.uint8)
    assert oracle.bf_insert_seqs(filt, bits, h, k, g, off) == n
    f = B.BloomFilter(bits, h, k, ctx=gpu.ctx)
    nthreads = 8
    per = (hs.shape[0] + nthreads - 1) // nthreads

    def work(t):
        for row in hs[t * per:(t + 1) * per]:
            f.insert(row)

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert f.getPop() == int(np.unpackbits(filt).sum())  # a read drains every thread's queue
    assert np.array_equal(f.to_numpy(), filt)
    # (b) counting filter, one thread, 70k k-mers with heavy repetition
    m = 50_000
    reps = np.concatenate([g[:30_000], g[:30_000], g[10_000:25_000]])
    roff = np.array([0, 30_000, 60_000, 75_000], np.uint64)
    n2, hs2, valid2 = oracle.hash_seqs(h, k, reps, roff)
    vmask = O.bits_to_bool(valid2, reps.size)
    rows = np.ascontiguousarray(hs2[vmask])
    cnt = np.zeros(m, np.uint8)
    assert oracle.cbf_insert_seqs(cnt, m, h, k, reps, roff) == rows.shape[0]
    c = B.CountingBloomFilter(m, h, k, 2, ctx=gpu.ctx)
    half = rows.shape[0] // 2
    for row in rows[:half]:
        c.insert(row)
    mid = c.minCount(rows[half - 1])  # a read in the middle of the stream sees the first half, all of it
    part = np.zeros(m, np.uint8)      # the sequential loop of CountingBloomFilter.hpp:134-162 over the first half
    for row in rows[:half]:
        slots = (row % np.uint64(m)).astype(np.int64)
        mn = part[slots].min()
        if mn < 255:
            part[slots[part[slots] == mn]] += 1
    assert mid == int(part[(rows[half - 1] % np.uint64(m)).astype(np.int64)].min())
    for row in rows[half:]:
        c.insert(row)
    assert np.array_equal(c.to_numpy(), cnt)
    # (c) one thread alternating between more filters than its queue cache holds: every filter gets exactly its k-mers
    small = [B.BloomFilter(1 << 16, h, k, ctx=gpu.ctx) for _ in range(6)]
    for i, row in enumerate(hs[:3000]):
        small[i % 6].insert(row)
    for j, sf in enumerate(small):
        exp = np.zeros((1 << 16) // 8, np.uint8)
        idx = (hs[j:3000:6].reshape(-1) % np.uint64(1 << 16)).astype(np.int64)
        np.bitwise_or.at(exp, idx >> 3, (1 << (idx & 7)).astype(np.uint8))
        assert np.array_equal(sf.to_numpy(), exp), "filter %d" % j
