#!/usr/bin/env python3
"""Device-resident kernel throughput of the other BASELINE.json configs (parity-test shapes, not the bench
headline): cfg3 query-only (k=32, h=6, 4 GiB filter), cfg4 CountingBloomFilter<uint8_t> (16e9 counters, k=25,
h=4, threshold 2), cfg5 spaced seeds (k=31, 2 seeds) on 64 MiB and 16 GiB filters.  One JSON line each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import btl_bloomfilter_b200 as B

dev = torch.device("cuda", 0)
ctx = B.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
for kv in sys.argv[1:]:
    k_, v_ = kv.split("=")
    ctx.set_option(k_, int(v_))
CH, RL = 64 << 20, 150
QF = int(os.environ.get("QUERY_FACTOR", "4"))  # read bases per query batch, in units of CH
NR = CH * QF // RL
RB = NR * RL
roff = torch.arange(0, RB + 1, RL, dtype=torch.int64, device=dev)
goff = torch.tensor([0, CH], dtype=torch.int64, device=dev)
hits = torch.zeros(RB // 32 + 8, dtype=torch.int32, device=dev)


def timed(fn, reps):
    fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(reps):
        fn(i + 1)
    ctx.flush()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(name, make, k, h, reps=6, insert_windows=CH, bytes_ins=None, bytes_qry=None):
    f = make()
    g = [torch.empty(CH + 64, dtype=torch.uint8, device=dev) for _ in range(reps + 1)]
    r = [torch.empty(RB + 64, dtype=torch.uint8, device=dev) for _ in range(3)]
    for i in range(reps + 1):
        ctx.synth_genome_device(g[i].data_ptr(), i * CH, CH, 42)
    for i in range(3):
        ctx.synth_reads_device(r[i].data_ptr(), 0, NR, RL, i * CH, CH, 42, 7 + i)
    st = torch.zeros(4, dtype=torch.int64, device=dev)
    iw = insert_windows
    io = torch.tensor([0, iw], dtype=torch.int64, device=dev)
    ms_i = timed(lambda i: f.insertSeqsDevice(g[i].data_ptr(), iw, io.data_ptr(), 1, st.data_ptr()), reps)
    st.zero_()
    ms_q = timed(lambda i: f.containsSeqsDevice(r[i % 3].data_ptr(), RB, roff.data_ptr(), NR, hits.data_ptr(), 0,
                                                st[2:].data_ptr()), 2)
    kq = NR * (RL - k + 1)
    ki = iw - k + 1
    s = st.cpu().numpy()
    out = {"config": name, "k": k, "hashes": h, "insert_windows": iw, "insert_ms": ms_i,
           "insert_gkmers_s": ki / ms_i / 1e6, "query_kmers": kq, "query_ms": ms_q, "query_gkmers_s": kq / ms_q / 1e6,
           "query_hit_fraction": float(s[3]) / max(1.0, float(s[2])), "options": sys.argv[1:],
           "insert_batches_per_timing": reps, "query_bases_per_batch": RB}
    if bytes_ins:
        out["insert_frac_of_6546GBs"] = out["insert_gkmers_s"] * bytes_ins / 6546.6
    if bytes_qry:
        out["query_frac_of_6546GBs"] = out["query_gkmers_s"] * bytes_qry / 6546.6
    if hasattr(f, "orderedStats"):
        out["ordered_deferred_rounds"] = f.orderedStats()
    print(json.dumps(out), flush=True)
    del f
    torch.cuda.empty_cache()


def spaced(bits):
    left = ["111101110111001", "111110110100111"]
    seeds = [x + "1" + x[::-1] for x in left]
    f = B.BloomFilter(bits, 2, 31, ctx=ctx)
    f.setSeeds(seeds, 1)
    return f


which = os.environ.get("CONFIGS", "cfg3,cfg4,cfg5a,cfg5b").split(",")
if "cfg3" in which:
    run("cfg3: k=32 h=6 2^35-bit BloomFilter", lambda: B.BloomFilter(1 << 35, 6, 32, ctx=ctx), 32, 6,
        bytes_ins=64 * 6 + 1, bytes_qry=32 * 6 + 1)
if "cfg4" in which:
    run("cfg4: CountingBloomFilter<uint8_t> 16e9 counters k=25 h=4 thr=2",
        lambda: B.CountingBloomFilter(16_000_000_000, 4, 25, 2, ctx=ctx), 25, 4, reps=2, insert_windows=16 << 20,
        bytes_ins=64 * 4 + 1, bytes_qry=32 * 4 + 1)
if "cfg5a" in which:
    run("cfg5: spaced seeds k=31 2 seeds, 2^29-bit (64 MiB, L2-resident) BloomFilter", lambda: spaced(1 << 29), 31, 2,
        bytes_ins=64 * 2 + 1, bytes_qry=32 * 2 + 1)
if "cfg5b" in which:
    run("cfg5: spaced seeds k=31 2 seeds, 2^37-bit (16 GiB) BloomFilter", lambda: spaced(1 << 37), 31, 2,
        bytes_ins=64 * 2 + 1, bytes_qry=32 * 2 + 1)
