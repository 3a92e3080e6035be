"""Parity at PRODUCTION geometry: the partitioned query (sort-bin pass 1 + probe pass 2, with and without the
two-stream overlap) on the BASELINE filter sizes -- 236 / 256 partitions of 16 MiB, batches large enough that the
auto rule takes that path -- compared bit by bit with the oracle on batches that contain hits AND misses.
(The small-geometry variants live in test_gpu_parity.py.)  Needs a B200."""
import numpy as np
import pytest

import _oracle as O
import parity_suite as S

pytestmark = pytest.mark.gpu


def _mixed_batch(oracle, g0, g_len, n_in, n_out, n_rand, seed):
    """150 bp reads: n_in sampled from genome[g0, g0+g_len) (inserted: every k-mer present), n_out from a far region
    of the genome and n_rand of independent random bases (absent but for false positives), shuffled, with a few N."""
    rl = 150
    a = oracle.synth_reads(0, n_in, rl, g_len, 42, seed, g_start=g0).reshape(n_in, rl)
    b = oracle.synth_reads(0, n_out, rl, 50_000_000, 42, seed + 1, g_start=1_500_000_000).reshape(n_out, rl)
    c = oracle.synth_genome(0, n_rand * rl, (43 << 40) + seed).reshape(n_rand, rl)
    reads = np.concatenate([a, b, c])
    rng = np.random.default_rng(seed)
    reads = reads[rng.permutation(reads.shape[0])]
    flat = np.ascontiguousarray(reads).reshape(-1)
    flat[rng.integers(0, flat.size, flat.size // 5000)] = ord("N")
    off = (rl * np.arange(reads.shape[0] + 1)).astype(np.uint64)
    return flat, off


@pytest.mark.parametrize("bits,k,h", [(31_568_113_856, 25, 4), (1 << 35, 32, 6)])
def test_partitioned_query_full_size_equals_oracle(oracle, bits, k, h):
    """cfg2 / cfg3 geometry.  Build: 8 Mbp of the synthetic genome generated in HBM, partitioned build (auto).
    Query: one batch of > 16 M windows mixing present and absent reads, partitioned query forced, adaptive gate
    off, (a) overlapped sub-batches and (b) one pass; expected bits = oracle hashes % m looked up in the set of
    inserted bit indices (BloomFilter.hpp:252-262)."""
    import torch
    import btl_bloomfilter_b200 as B
    ctx = B.Context(0)
    dev = torch.device("cuda", 0)
    g0, n_ins = 5_000_000, 8_000_000
    f = B.BloomFilter(bits, h, k, ctx=ctx)
    g = torch.empty(n_ins + 64, dtype=torch.uint8, device=dev)
    ctx.synth_genome_device(g.data_ptr(), g0, n_ins, 42)
    off_d = torch.tensor([0, n_ins], dtype=torch.int64, device=dev)
    b0 = ctx.counter("binned_launches")
    f.insertSeqsDevice(g.data_ptr(), n_ins, off_d.data_ptr(), 1, 0)
    ctx.sync()
    assert ctx.counter("binned_launches") == b0 + 1, "the build did not take the partitioned path"
    # the inserted index set, from the oracle
    host_g = oracle.synth_genome(g0, n_ins, 42)
    assert np.array_equal(host_g[:100_000], g[:100_000].cpu().numpy())
    _, hs, _ = oracle.hash_seqs(h, k, host_g, np.array([0, n_ins], np.uint64))
    inserted = np.unique(hs[: n_ins - k + 1].reshape(-1) % np.uint64(bits))
    del hs
    assert f.getPop() == inserted.size
    # the query batch and its expected answer
    qb, qoff = _mixed_batch(oracle, g0, n_ins, 60_000, 30_000, 30_000, seed=3)
    assert qb.size >= 16_000_000
    nv, qh, valid = oracle.hash_seqs(h, k, qb, qoff)
    vmask = O.bits_to_bool(valid, qb.size)
    present = np.zeros(qb.size, bool)
    idx = qh[vmask] % np.uint64(bits)
    del qh
    present[vmask] = np.isin(idx.reshape(-1), inserted).reshape(idx.shape).all(axis=1)
    del idx
    n_hit = int(present.sum())
    assert 0.3 * nv < n_hit < 0.7 * nv  # the batch really mixes hits and misses
    ctx.set_option("bin_query_mode", 1)
    ctx.set_option("query_adaptive", 0)
    for sub in (4, 1, 3):
        ctx.set_option("query_sub", sub)
        b0 = ctx.counter("binned_launches")
        r = f.containsSeqs((qb, qoff))
        assert ctx.counter("binned_launches") > b0, "the query did not take the partitioned path"
        assert (r.n_kmers, r.n_hits) == (nv, n_hit), "query_sub=%d" % sub
        assert np.array_equal(r.valid_bits, valid)
        assert np.array_equal(O.bits_to_bool(r.hit_bits, qb.size), present), "query_sub=%d" % sub
    # the auto rule (adaptive gate on, default sub-batching) gives the same bits
    for key, v in (("bin_query_mode", 0), ("query_adaptive", 1), ("query_sub", 0)):
        ctx.set_option(key, v)
    r = f.containsSeqs((qb, qoff))
    assert (r.n_kmers, r.n_hits) == (nv, n_hit)
    assert np.array_equal(O.bits_to_bool(r.hit_bits, qb.size), present)
    # ... and so does the direct gather kernel
    ctx.set_option("bin_query_mode", -1)
    r = f.containsSeqs((qb, qoff))
    assert np.array_equal(O.bits_to_bool(r.hit_bits, qb.size), present)
    # the same batch as 2 bits per base + invalid plane (btlbf_pack_seqs, btlbf_contains_seqs_packed), partitioned path
    pk = B.pack_seqs((qb, qoff))
    assert pk.invalid is not None and pk.n_invalid > 0
    ctx.set_option("bin_query_mode", 1)
    ctx.set_option("query_adaptive", 0)
    b0 = ctx.counter("binned_launches")
    r = f.containsSeqsPacked(pk)
    assert ctx.counter("binned_launches") > b0
    assert (r.n_kmers, r.n_hits) == (nv, n_hit) and np.array_equal(r.valid_bits, valid)
    assert np.array_equal(O.bits_to_bool(r.hit_bits, qb.size), present), "packed query"
    # ... and the packed build of the same genome gives the ASCII build's bytes
    from btl_bloomfilter_b200 import parallel
    f2 = B.BloomFilter(bits, h, k, ctx=ctx)
    b0 = ctx.counter("binned_launches")
    assert f2.insertSeqsPacked(B.pack_seqs((host_g, np.array([0, n_ins], np.uint64)))) == n_ins - k + 1
    assert ctx.counter("binned_launches") > b0, "the packed build did not take the partitioned path"
    va = parallel.device_tensor_from_ptr(*f.device_ptr(), dev)
    vb = parallel.device_tensor_from_ptr(*f2.device_ptr(), dev)
    assert torch.equal(va, vb), "packed build differs from the ASCII build"
    for key, v in (("bin_query_mode", 0), ("query_adaptive", 1)):
        ctx.set_option(key, v)
    del f, f2, va, vb
    ctx.close()


@pytest.mark.parametrize("legacy", [0, 1])
@pytest.mark.parametrize("shift", [8, 12, 20])
def test_partitioned_counting_query_equals_oracle(oracle, golden, shift, legacy):
    """CountingBloomFilter::contains (minCount >= threshold, CountingBloomFilter.hpp:190-196) through the partitioned
    query at small geometry: every partition width, both pass-1 kernel families, thresholds 1..3."""
    from _backends import GpuBackend
    be = GpuBackend(bin_shift=shift, bin_kernel=legacy)
    S.check_golden_cbf(be, golden)
    for thr in (1, 2, 3):
        S.check_random_cbf(be, oracle, 25, 4, 100_008, seed=40 + thr, thr=thr)
    S.check_random_cbf(be, oracle, 9, 6, 4096, seed=7, n_seqs=80, max_len=200)
    assert be.ctx.counter("binned_launches") > 0


def test_partitioned_counting_query_full_size(oracle):
    """cfg4 geometry (16e9 counters): the partitioned threshold query against a sparse sequential replay."""
    import btl_bloomfilter_b200 as B
    ctx = B.Context(0)
    m, h, k, thr = 16_000_000_000, 4, 25, 2
    f = B.CountingBloomFilter(m, h, k, thr, ctx=ctx)
    g = oracle.synth_genome(1000, 60_000, 42)
    # the same region three times over, in pieces: counts 1..3 along it, plus reads that were never inserted
    ins = np.concatenate([g, g[:40_000], g[:20_000]])
    ioff = np.array([0, 60_000, 100_000, 120_000], np.uint64)
    assert f.insertSeqs((ins, ioff)) == (60_000 - 24) + (40_000 - 24) + (20_000 - 24)
    _, hs, valid = oracle.hash_seqs(h, k, ins, ioff)
    vmask = O.bits_to_bool(valid, ins.size)
    cnt = {}
    for p in np.nonzero(vmask)[0]:
        slots = [int(x) for x in hs[p] % np.uint64(m)]
        mn = min(cnt.get(s_, 0) for s_ in slots)
        if mn < 255:
            for s_ in slots:
                if cnt.get(s_, 0) == mn:
                    cnt[s_] = mn + 1
    q = np.concatenate([g, oracle.synth_genome(0, 40_000, 43 << 40)])
    qoff = np.array([0, 60_000, 100_000], np.uint64)
    nv, qh, qvalid = oracle.hash_seqs(h, k, q, qoff)
    qmask = O.bits_to_bool(qvalid, q.size)
    exp = np.zeros(q.size, bool)
    for p in np.nonzero(qmask)[0]:
        exp[p] = min(cnt.get(int(x), 0) for x in qh[p] % np.uint64(m)) >= thr
    assert 30_000 < exp.sum() < 50_000
    direct = f.containsSeqs((q, qoff))
    assert np.array_equal(O.bits_to_bool(direct.hit_bits, q.size), exp)
    ctx.set_option("bin_query_mode", 1)
    for sub in (1, 2):
        ctx.set_option("query_sub", sub)
        b0 = ctx.counter("binned_launches")
        r = f.containsSeqs((q, qoff))
        assert ctx.counter("binned_launches") > b0
        assert (r.n_kmers, r.n_hits) == (nv, int(exp.sum()))
        assert np.array_equal(O.bits_to_bool(r.hit_bits, q.size), exp) and np.array_equal(r.valid_bits, qvalid)
    del f
    ctx.close()


def test_stream_switch_orders_after_parked_kmers(oracle):
    """btlbf_ctx_set_stream while k-mers are parked in the partition buckets: the deferred pass 2 runs on the OLD
    stream, and work queued on the NEW stream right after the switch must see its result (the multi-GPU merge does
    exactly this before handing the array to NCCL)."""
    import torch
    import btl_bloomfilter_b200 as B
    from btl_bloomfilter_b200 import parallel
    ctx = B.Context(0)
    dev = torch.device("cuda", 0)
    ctx.set_option("bin_mode", 1)
    ctx.set_option("bin_part_log2", 20)
    bits, h, k = 1 << 27, 4, 25
    t = torch.zeros(bits // 8, dtype=torch.uint8, device=dev)
    g = oracle.synth_genome(0, 2_000_000, 42)
    off = np.array([0, g.size], np.uint64)
    filt = np.zeros(bits // 8, np.uint8)
    oracle.bf_insert_seqs(filt, bits, h, k, g, off)
    exp = torch.from_numpy(filt).to(dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    # (a) the library's own allocation: parked across the switch
    f = B.BloomFilter(bits, h, k, ctx=ctx)
    ctx.set_stream(s1.cuda_stream)
    f.insertSeqs((g, off))
    ctx.set_stream(s2.cuda_stream)
    ptr, nbytes = f.device_ptr()
    with torch.cuda.stream(s2):
        same = torch.equal(parallel.device_tensor_from_ptr(ptr, nbytes, dev), exp)
    assert same
    # (b) caller-owned memory: nothing stays parked when the insert call returns
    w = B.BloomFilter.from_device_memory(t, bits, h, k, ctx=ctx)
    w.insertSeqs((g, off))
    with torch.cuda.stream(s2):
        same = torch.equal(t, exp)
    assert same
    ctx.set_stream(0)
    del f, w
    ctx.close()
