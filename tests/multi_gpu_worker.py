"""Worker of tests/test_multi_gpu.py (launched by torch.distributed.run, one rank per GPU): read-sharded
build into per-GPU partial filters + NCCL merge, and read-sharded query against the replicated filter,
checked against the oracle on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as O  # noqa: E402
import btl_bloomfilter_b200 as B  # noqa: E402
from btl_bloomfilter_b200 import parallel  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    orc = O.Oracle()
    ctx = B.Context(local)
    k, h = 25, 4
    n_reads, rl, g_len = 6000, 150, 400_000
    reads = orc.synth_reads(0, n_reads, rl, g_len, 42, 5)
    off = (rl * np.arange(n_reads + 1)).astype(np.uint64)
    lo, hi = parallel.shard_range(n_reads, rank, world)
    my = (reads[lo * rl:hi * rl], off[lo:hi + 1] - off[lo])

    # ---- BloomFilter: partial builds + OR merge == single build of all reads (bit-exact for any sharding)
    bits = 8 * 300_007 * 8  # not a power of two, not a multiple of the slice size
    nbytes = bits // 8
    t = torch.zeros(parallel.padded_bytes(nbytes, world), dtype=torch.uint8, device=dev)
    f = B.BloomFilter.from_device_memory(t, bits, h, k, ctx=ctx)
    f.insertSeqs(my)
    parallel.merge_filter(f)
    torch.cuda.synchronize()
    filt = np.zeros(nbytes, np.uint8)
    orc.bf_insert_seqs(filt, bits, h, k, reads, off)
    got = f.to_numpy()
    assert np.array_equal(got, filt), "rank %d: merged BloomFilter differs from the single-build oracle" % rank
    assert not t[nbytes:].any()

    # ---- the same merge as ONE kernel per GPU over peer-mapped memory (CUDA IPC + NVLink loads/stores)
    t2 = torch.zeros(parallel.padded_bytes(nbytes, world), dtype=torch.uint8, device=dev)
    f2 = B.BloomFilter.from_device_memory(t2, bits, h, k, ctx=ctx)
    f2.insertSeqs(my)
    parallel.fused_merge_filter(f2)
    assert np.array_equal(f2.to_numpy(), filt), "rank %d: fused peer-memory merge differs from the oracle" % rank
    f3 = B.BloomFilter(bits, h, k, ctx=ctx)  # the library's own allocation, merged twice (OR is idempotent)
    f3.insertSeqs(my)
    pm = parallel.PeerMerge(ctx, *f3.device_ptr(), f3.KIND)
    pm.merge()
    pm.merge()
    pm.close()
    assert np.array_equal(f3.to_numpy(), filt)

    # ---- the partitioned build parks k-mers until the filter is next touched (forced here on a small filter): the
    # merges must see them -- NCCL path through a stream switch (set_stream), peer path through flush()
    ctx2 = B.Context(local)
    for key, val in (("bin_mode", 1), ("bin_part_log2", 14)):
        ctx2.set_option(key, val)
    f4 = B.BloomFilter(bits, h, k, ctx=ctx2)
    f4.insertSeqs(my)
    assert ctx2.counter("binned_launches") > 0
    pm4 = parallel.PeerMerge(ctx2, *f4.device_ptr(), f4.KIND)
    pm4.merge()
    pm4.close()
    assert np.array_equal(f4.to_numpy(), filt), "rank %d: parked k-mers were lost by the peer merge" % rank
    ctx2.set_option("wrap_accumulate", 1)
    t5 = torch.zeros(parallel.padded_bytes(nbytes, world), dtype=torch.uint8, device=dev)
    f5 = B.BloomFilter.from_device_memory(t5, bits, h, k, ctx=ctx2)
    f5.insertSeqs(my)
    parallel.merge_filter(f5)
    torch.cuda.synchronize()
    assert np.array_equal(f5.to_numpy(), filt), "rank %d: parked k-mers were lost across the stream switch" % rank

    # ---- the OR merge inside NVSwitch (multimem.ld_reduce.or / multimem.st over a multicast mapping of symmetric
    # memory), where the system offers NVLS multicast
    f6, hdl = parallel.symmetric_filter(B.BloomFilter, bits, h, k, ctx2)
    mm_ok = parallel.MultimemMerge.available(hdl)
    if mm_ok:
        f6.insertSeqs(my)
        mm = parallel.MultimemMerge(ctx2, hdl, nbytes)
        mm.merge()
        assert np.array_equal(f6.to_numpy(), filt), "rank %d: in-switch merge differs from the oracle" % rank
        mm.merge()  # idempotent
        assert np.array_equal(f6.to_numpy(), filt)
        mm.close()
        r6 = f6.containsSeqs(my)
        assert r6.n_hits == r6.n_kmers
    # the peer-memory kernel over the symmetric handle's peer pointers (no CUDA IPC), and the calibrated choice
    f7, hdl7 = parallel.symmetric_filter(B.BloomFilter, bits, h, k, ctx2)
    f7.insertSeqs(my)
    m7 = parallel.MultimemMerge(ctx2, hdl7, nbytes, mode="peer")
    m7.merge()
    assert np.array_equal(f7.to_numpy(), filt), "rank %d: peer merge over symmetric memory differs" % rank
    if mm_ok:  # both mechanisms side by side in one kernel
        f8, hdl8 = parallel.symmetric_filter(B.BloomFilter, bits, h, k, ctx2)
        f8.insertSeqs(my)
        m8 = parallel.MultimemMerge(ctx2, hdl8, nbytes, mode="hybrid50")
        m8.merge()
        assert np.array_equal(f8.to_numpy(), filt), "rank %d: hybrid merge differs from the oracle" % rank
        m8.close()
    assert m7.calibrate() in m7.modes
    m7.merge()
    assert np.array_equal(f7.to_numpy(), filt)
    m7.close()

    # ---- pass 2 of the build and the merge pipelined in chunks of partitions (btlbf_filter_flush_parts +
    # btlbf_merge_peers_range): k-mers parked by two insert calls, 1 / 3 / 5 chunks
    ctx2.set_stream(torch.cuda.current_stream().cuda_stream)
    side = torch.cuda.Stream(device=dev)
    tok = torch.zeros(1, dtype=torch.int32, device=dev)
    half = (my[1].size - 1) // 2
    for n_chunks in (1, 3, 5):
        f9, hdl9 = parallel.symmetric_filter(B.BloomFilter, bits, h, k, ctx2)
        b0 = ctx2.counter("binned_launches")
        f9.insertSeqs((my[0][: int(my[1][half])], my[1][: half + 1]))
        f9.insertSeqs((my[0][int(my[1][half]):], my[1][half:] - my[1][half]))
        assert ctx2.counter("binned_launches") >= b0 + 2
        m9 = parallel.MultimemMerge(ctx2, hdl9, nbytes, mode="peer")
        parallel.pipelined_flush_merge(f9, m9, n_chunks, side, tok)
        torch.cuda.synchronize()
        dist.barrier()
        assert np.array_equal(f9.to_numpy(), filt), "rank %d: pipelined merge in %d chunks differs" % (rank, n_chunks)
        m9.close()
    ctx2.set_stream(None)

    # ---- query: filter replicated (after the merge), reads sharded, no collective
    r = f.containsSeqs(my)
    nq, nh, hits, valid = orc.bf_contains_seqs(filt, bits, h, k, my[0], my[1])
    assert (r.n_kmers, r.n_hits) == (nq, nh) and np.array_equal(r.hit_bits, hits)
    miss = orc.synth_genome(rank * 50_000, 50_000, 43 << 40)
    r = f.containsSeqs((miss, np.array([0, miss.size], np.uint64)))
    e = orc.bf_contains_seqs(filt, bits, h, k, miss, np.array([0, miss.size], np.uint64))
    assert (r.n_kmers, r.n_hits) == e[:2] and np.array_equal(r.hit_bits, e[2])

    # ---- CountingBloomFilter: partial incrementMin builds + saturating add; oracle = the same sharded algorithm
    m = 100_008
    tc = torch.zeros(parallel.padded_bytes(m, world), dtype=torch.uint8, device=dev)
    c = B.CountingBloomFilter.from_device_memory(tc, m, h, k, threshold=2, ctx=ctx)
    c.insertSeqs(my)
    parallel.merge_filter(c)
    torch.cuda.synchronize()
    exp = np.zeros(m, np.int32)
    for rr in range(world):
        a, b = parallel.shard_range(n_reads, rr, world)
        part = np.zeros(m, np.uint8)
        orc.cbf_insert_seqs(part, m, h, k, reads[a * rl:b * rl], off[a:b + 1] - off[a])
        exp += part
    exp = np.minimum(exp, 255).astype(np.uint8)
    assert np.array_equal(c.to_numpy(), exp), "rank %d: merged counters differ" % rank
    c2 = B.CountingBloomFilter(m, h, k, 2, ctx=ctx)
    c2.insertSeqs(my)
    parallel.fused_merge_filter(c2)
    assert np.array_equal(c2.to_numpy(), exp), "rank %d: fused saturating-add merge differs" % rank
    dist.barrier()
    if rank == 0:
        print("MULTI-GPU OK world=%d multimem=%s" % (world, "yes" if mm_ok else "unavailable"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
