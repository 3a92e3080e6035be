#!/usr/bin/env python3
"""bench.py -- k-mer insert + query throughput of the BloomFilter hot path on B200.

Workload (BASELINE.json configs[1], "cfg2"): 3 Gbp synthetic genome, k=25, 4 hashes, 31,568,113,856-bit
filter (the reference's calcOptimalSize(3e9, 0.01), 3.95 GB), built in chunks, then 150 bp read queries.
One STEP = one pass of the hot path over one batch: insert one 64 Mi-window genome chunk into the filter
and query one batch of 150 bp reads (4 x 64 MiB of bases by default) sampled from that chunk (all k-mers
present: no early exit).  Like the workload itself (build the filter, then query reads), the timed region
runs the K build batches first and the K query batches after them; both are inside the timed region.

  value     whole-job Gk-mer/s (inserted + queried) with the batches already resident in HBM
            (btlbf_insert_seqs_dev / btlbf_contains_seqs_dev), CUDA events, max over ranks
  e2e       the same steps through the host-buffer C-ABI calls (btlbf_insert_seqs / btlbf_contains_seqs):
            pinned host inputs, H2D + kernels + D2H of the hit bits inside the timed region
  roofline  the phase that dominates the step (the query: pass 1 bin_kernel_sort + pass 2 probe_bins_kernel):
            (32*h + 1) algorithmic bytes per k-mer / its mean duration per step (CUDA events inside the timed
            region) against the measured HBM copy bandwidth; roofline_build (64*h + 1 bytes per k-mer) and
            roofline_step (both phases) follow
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified headers, OpenMP over reads / chunks)
            on a bounded sample of the same workload, same filter size, on this box's host cores

`--impl reference` times only that CPU path, K bounded-sample steps.  N > 1 (torchrun): one rank per GPU,
units (genome chunks / read batches) sharded across ranks, no data-path collective ("weak" scaling);
the partial filters are merged afterwards (all-to-all of 1/N slices + OR + all-gather) and that merge
is timed and reported separately under "merge".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_LEN = 3_000_000_000
K, H = 25, 4
FILTER_BITS = 31_568_113_856  # BloomFilter::calcOptimalSize(3e9, 0.01) with 4 hashes (BloomFilter.hpp:406-413)
CHUNK = 64 << 20              # windows per insert launch
READ_LEN = 150
GENOME_SEED, READ_SEED = 42, 7
WORKLOAD = "cfg2: 3 Gbp synthetic genome build (k=25, h=4, 31.57 Gbit filter) + 150 bp read query"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- CPU reference arm
def cpu_reference_run(n_steps, warmup, sample_bases, threads=0, verbose=False):
    """The reference's CPU path on a bounded sample per step: insert `sample_bases` of the genome
    (64 kb pieces, OpenMP over pieces) + query sample_bases/150 reads sampled from it."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    orc = O.Oracle()
    use_ref = O.Ref.available()
    piece = 65536
    if use_ref:
        R = O.Ref()
        cores = R.L.ref_max_threads()
        filt = R.bf_new(FILTER_BITS, H, K)
    else:
        cores = orc.L.ora_max_threads()
        filt = np.zeros(FILTER_BITS // 8, np.uint8)
    threads = threads or cores
    n_reads = sample_bases // READ_LEN
    roff = (READ_LEN * np.arange(n_reads + 1)).astype(np.uint64)
    tot_k, tot_t, per = 0, 0.0, []
    for s in range(warmup + n_steps):
        g0 = (s * sample_bases) % (G_LEN - sample_bases - K)
        g = orc.synth_genome(g0, sample_bases + K - 1, GENOME_SEED)
        # 64 kb pieces overlapping by k-1 so that every window is inserted exactly once
        starts = np.arange(0, sample_bases, piece, dtype=np.uint64)
        pieces = [g[int(a): int(min(a + piece + K - 1, g.size))] for a in starts]
        pb = np.concatenate(pieces)
        poff = np.concatenate([[0], np.cumsum([p.size for p in pieces])]).astype(np.uint64)
        reads = orc.synth_reads(0, n_reads, READ_LEN, sample_bases, GENOME_SEED, READ_SEED + s, g_start=g0)
        nk, nh = O.u64(), O.u64()
        if use_ref:
            t_i = R.L.ref_bench_bf(filt, O._p8(pb), O._p64(poff), poff.size - 1, 1, threads, C.byref(nk), C.byref(nh))
            k_i = nk.value
            t_q = R.L.ref_bench_bf(filt, O._p8(reads), O._p64(roff), n_reads, 0, threads, C.byref(nk), C.byref(nh))
        else:
            t_i = orc.L.ora_bench_bf(O._p8(filt), FILTER_BITS, H, K, O._p8(pb), O._p64(poff), poff.size - 1, 1, threads,
                                     C.byref(nk), C.byref(nh))
            k_i = nk.value
            t_q = orc.L.ora_bench_bf(O._p8(filt), FILTER_BITS, H, K, O._p8(reads), O._p64(roff), n_reads, 0, threads,
                                     C.byref(nk), C.byref(nh))
        k_q = nk.value
        assert nh.value == k_q, "CPU reference: a k-mer of an inserted region was not found"
        if s >= warmup:
            tot_k += k_i + k_q
            tot_t += t_i + t_q
            per.append((k_i / t_i, k_q / t_q))
        if verbose:
            print("cpu step %d: insert %.2f Mk/s, query %.2f Mk/s" % (s, k_i / t_i / 1e6, k_q / t_q / 1e6), file=sys.stderr)
    if use_ref:
        R.L.ref_bf_free(filt)
    return {"value": tot_k / tot_t / 1e9, "unit": "Gk-mer/s", "cores": int(threads),
            "kind": "reference" if use_ref else "port",
            "sample": "%d steps x (%d bp genome insert in 64 kb pieces + %d reads x %d bp query), %d-bit filter, "
                      "OpenMP over pieces/reads" % (n_steps, sample_bases, n_reads, READ_LEN, FILTER_BITS),
            "insert_gkmers_s": float(np.mean([p[0] for p in per])) / 1e9,
            "query_gkmers_s": float(np.mean([p[1] for p in per])) / 1e9,
            "ms_per_step": tot_t / n_steps * 1e3, "kmers_per_step": tot_k / n_steps}


def bind_to_gpu_numa_node(torch, index):
    """N > 1: run this rank (and first-touch its pinned buffers) on the CPUs next to its GPU, so that eight ranks do
    not pull their host buffers across the socket interconnect.  Best effort; returns what it did."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"pci": bdf, "cpus": len(cpus)}
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:80]}
    return None


# ---------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=42)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8 << 20, help="bases per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--query-factor", type=int, default=4, help="read bases per query batch, in units of --chunk")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (0: leave as is)")
    ap.add_argument("--stream-priority", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (tuning experiments)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": WORKLOAD, "k": K, "hashes": H, "filter_bits": FILTER_BITS, "genome_bp": G_LEN,
              "read_len": READ_LEN, "chunk_windows": args.chunk, "query_bases_per_step": args.chunk * args.query_factor,
              "step": "K build batches, then K query batches, all inside the timed region",
              "l2": "inputs larger than L2 (64 MiB / 256 MiB per batch, 3.95 GB filter); no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, args.steps)
        r = cpu_reference_run(steps, min(args.warmup, 1), args.cpu_sample)
        line = {"impl": "reference", "metric": "k-mers/s inserted+queried", "value": r["value"], "unit": "Gk-mer/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": config, "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "Gk-mer/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import btl_bloomfilter_b200 as B

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = B.Context(local_rank)
    # the library's kernels run on torch's current stream so that torch.cuda.Event brackets them
    # (a non-default stream: the C ABI treats a NULL stream as "use the context's own")
    stream = torch.cuda.Stream(device=dev, priority=args.stream_priority)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if args.l2_fetch:
        ctx.set_option("l2_fetch_granularity", args.l2_fetch)
    for kv in args.opt:
        key, val = kv.split("=")
        ctx.set_option(key, int(val))
        config.setdefault("options", {})[key] = int(val)

    chunk = args.chunk // 4096 * 4096
    n_chunks = (G_LEN + chunk - 1) // chunk
    W, S = args.warmup, args.steps
    n_reads = chunk * args.query_factor // READ_LEN
    read_bases = n_reads * READ_LEN
    filt = B.BloomFilter(FILTER_BITS, H, K, ctx=ctx)

    # ---- synthetic inputs generated in HBM (replayable by the oracle: tests/test_gpu_parity.py)
    n_buf = min(W + S, n_chunks)
    d_genome, d_reads, g_lens = [], [], []
    for i in range(n_buf):
        c = (i * world + rank) % n_chunks
        g0 = c * chunk
        glen = min(chunk + K - 1, G_LEN - g0)
        tg = torch.empty(chunk + 64, dtype=torch.uint8, device=dev)
        ctx.synth_genome_device(tg.data_ptr(), g0, glen, GENOME_SEED)
        tr = torch.empty(read_bases + 64, dtype=torch.uint8, device=dev)
        ctx.synth_reads_device(tr.data_ptr(), 0, n_reads, READ_LEN, g0, glen, GENOME_SEED, READ_SEED + c)
        d_genome.append(tg)
        d_reads.append(tr)
        g_lens.append(glen)
    d_goff = [torch.tensor([0, gl], dtype=torch.int64, device=dev) for gl in sorted(set(g_lens))]
    goff_of = {int(t[1]): t for t in d_goff}
    d_roff = torch.arange(0, read_bases + 1, READ_LEN, dtype=torch.int64, device=dev)
    d_hits = torch.zeros((read_bases + 31) // 32 + 8, dtype=torch.int32, device=dev)
    d_stats = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    def build_dev(i):
        j = i % n_buf
        filt.insertSeqsDevice(d_genome[j].data_ptr(), g_lens[j], goff_of[g_lens[j]].data_ptr(), 1, d_stats.data_ptr())

    def query_dev(i):
        j = i % n_buf
        filt.containsSeqsDevice(d_reads[j].data_ptr(), read_bases, d_roff.data_ptr(), n_reads, d_hits.data_ptr(), 0,
                                d_stats[2:].data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()

    for i in range(W):
        build_dev(i)
    for i in range(W):
        query_dev(i)
    torch.cuda.synchronize()
    d_stats.zero_()
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    ev_q = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    t_start, t_mid, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    barrier()
    torch.cuda.synchronize()
    t_start.record(stream)
    for i in range(S):
        build_dev(W + i)
        ev_b[i].record(stream)
    ctx.flush()  # every k-mer parked in the partition buckets reaches the filter (pass 2) before t_mid
    t_mid.record(stream)
    for i in range(S):
        query_dev(W + i)
        ev_q[i].record(stream)
    t_end.record(stream)
    torch.cuda.synchronize()
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = t_start.elapsed_time(t_end)
    ms_insert = t_start.elapsed_time(t_mid)
    ms_query = t_mid.elapsed_time(t_end)
    # per-call durations on the stream: a build call is pass 1 of its batch, plus pass 2 of the whole
    # accumulation when that call filled the partition buckets (the median is therefore pass 1 alone)
    d_b = [(t_start if i == 0 else ev_b[i - 1]).elapsed_time(ev_b[i]) for i in range(S)]
    d_q = [(t_mid if i == 0 else ev_q[i - 1]).elapsed_time(ev_q[i]) for i in range(S)]
    ms_insert_pass1 = float(np.median(d_b)) * S
    st = d_stats.cpu().numpy()
    k_ins, k_qry, k_hit = int(st[0]), int(st[2]), int(st[3])
    assert k_hit == k_qry, "a k-mer of an inserted chunk was not found (%d of %d)" % (k_hit, k_qry)

    # ---- the miss set: reads of independent random bases, almost every k-mer absent.  Reported next
    # to the headline (which is the hit set: no early exit); outside the timed region of `value`.
    miss = None
    if rank == 0:
        d_miss = torch.empty(read_bases + 64, dtype=torch.uint8, device=dev)
        ctx.synth_genome_device(d_miss.data_ptr(), 0, read_bases, 43 << 40)
        d_ms = torch.zeros(2, dtype=torch.int64, device=dev)
        reps = 3
        filt.containsSeqsDevice(d_miss.data_ptr(), read_bases, d_roff.data_ptr(), n_reads, d_hits.data_ptr(), 0, d_ms.data_ptr())
        torch.cuda.synchronize()
        d_ms.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            filt.containsSeqsDevice(d_miss.data_ptr(), read_bases, d_roff.data_ptr(), n_reads, d_hits.data_ptr(), 0,
                                    d_ms.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
        mk, mh = [int(x) for x in d_ms.cpu().numpy()]
        ms_miss = a.elapsed_time(b) / reps
        miss = {"gkmers_s": mk / reps / (ms_miss * 1e-3) / 1e9, "ms_per_batch": ms_miss, "kmers_per_batch": mk // reps,
                "hit_fraction": mh / max(1, mk),
                "path": "adaptive: sampled hit fraction on the device picks the early-exit kernel for read sets that mostly miss"}
        del d_miss

    # ---- end to end through the host-buffer C ABI (pinned inputs; H2D, kernels and D2H inside the timed region)
    # Streaming form (btlbf_insert_seqs_async / btlbf_contains_seqs_async): every step copies its genome chunk
    # and reads host->device, runs the kernels and copies the hit bits + counts device->host; the calls are
    # queued back to back, so the copies of one step overlap the kernels of another, and the results are
    # checked after the final btlbf_ctx_sync.  "e2e_sync" is the same with the blocking calls.
    e2e = None
    if not args.no_e2e:
        n_host = min(n_buf, 4)
        h_genome = [torch.empty(g_lens[j], dtype=torch.uint8).pin_memory() for j in range(n_host)]
        h_reads = [torch.empty(read_bases, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        for j in range(n_host):
            h_genome[j].copy_(d_genome[j][: g_lens[j]])
            h_reads[j].copy_(d_reads[j][:read_bases])
        h_hits = [torch.zeros((read_bases + 31) // 32 * 4, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        t_roff = torch.arange(0, read_bases + 1, READ_LEN, dtype=torch.int64).pin_memory()
        h_roff = t_roff.numpy().view(np.uint64)
        t_goff = [torch.tensor([0, g_lens[j]], dtype=torch.int64).pin_memory() for j in range(n_host)]
        h_goff = [t.numpy().view(np.uint64) for t in t_goff]
        S2 = min(S, 32)
        h_counts = torch.zeros((S2 + 4, 4), dtype=torch.int64).pin_memory()
        counts = h_counts.numpy().view(np.uint64)
        torch.cuda.synchronize()

        def step_host_sync(i):
            j = i % n_host
            a = filt.insertSeqs((h_genome[j].numpy(), h_goff[j]))
            r = filt.containsSeqs((h_reads[j].numpy(), h_roff), hit_out=h_hits[j].numpy(), want_valid=False)
            return a, r.n_kmers, r.n_hits

        def build_host_async(i, slot):
            j = i % n_host
            filt.insertSeqsAsync((h_genome[j].numpy(), h_goff[j]), counts[slot, 0:2])

        def query_host_async(i, slot):
            j = i % n_host
            filt.containsSeqsAsync((h_reads[j].numpy(), h_roff), h_hits[j].numpy(), counts[slot, 2:4])

        for i in range(min(W, 2)):
            step_host_sync(i)
        # warm-up of the host path: the pinned buffers, the copy engines and the PCIe link (which trains up under
        # traffic) -- a fresh process on a fresh box measures ~20 % low without it
        for rep in range(3):
            for i in range(4):
                build_host_async(i, S2 + (i & 1))
            for i in range(4):
                query_host_async(i, S2 + (i & 1))
            ctx.sync()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(S2):
            build_host_async(i, i)
        for i in range(S2):
            query_host_async(i, i)
        ctx.sync()
        dt = time.perf_counter() - t0
        barrier()
        ke = int(counts[:S2, 0].sum() + counts[:S2, 2].sum())
        assert np.array_equal(counts[:S2, 2], counts[:S2, 3]) and counts[:S2, 2].all(), "e2e: a queried k-mer was not found"
        S3 = min(S, 6)
        t0 = time.perf_counter()
        ks = 0
        for i in range(S3):
            a, q, hq = step_host_sync(i)
            assert hq == q
            ks += a + q
        dts = time.perf_counter() - t0
        barrier()
        e2e = {"kmers": ke, "seconds": dt, "steps": S2, "kmers_sync": ks, "seconds_sync": dts, "steps_sync": S3,
               "h2d": int(np.mean(g_lens[:n_host])) + read_bases + (n_reads + 1) * 8 + 16,
               "d2h": int(h_hits[0].numel()) + 32}

    # ---- reductions over ranks (max time, summed work)
    if world > 1:
        t = torch.tensor([ms_total, ms_insert, ms_query, e2e["seconds"] if e2e else 0.0,
                          e2e["seconds_sync"] if e2e else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_insert, ms_query = float(t[0]), float(t[1]), float(t[2])
        w = torch.tensor([k_ins, k_qry, e2e["kmers"] if e2e else 0, launches, e2e["kmers_sync"] if e2e else 0],
                         dtype=torch.int64, device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        k_ins_all, k_qry_all, ke_all, launches_all, ks_all = [int(x) for x in w]
        if e2e:
            e2e["seconds"], e2e["seconds_sync"] = float(t[3]), float(t[4])
            e2e["kmers"], e2e["kmers_sync"] = ke_all, ks_all
    else:
        k_ins_all, k_qry_all, launches_all = k_ins, k_qry, launches

    merge = None
    if world > 1:
        from btl_bloomfilter_b200 import parallel
        merge = parallel.bench_merge(filt, ctx, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    # rooflines per rank, algorithmic bytes of SURVEY.md 8d: build 64 B per hash (32 B sector read + 32 B dirty
    # write-back) + 1 input byte per k-mer, query 32 B per hash + 1
    ins_bytes = (64 * H + 1) * (k_ins / S)
    ins_ms = ms_insert / S
    qry_bytes = (32 * H + 1) * (k_qry / S)
    qry_ms = ms_query / S
    roof = {"bound": "hbm", "kernel": "BloomFilter build: bin_kernel_sort per batch + apply_bins_kernel per accumulation "
                                      "(all durations of the build phase added)", "achieved": ins_bytes / (ins_ms * 1e-3) / 1e9,
            "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None,
            "bytes_per_kmer": 64 * H + 1, "kmers_per_launch": k_ins / S, "launch_ms": ins_ms,
            "pass1_ms": ms_insert_pass1 / S, "pass2_ms": (ms_insert - ms_insert_pass1) / S,
            "gkmers_s": k_ins / S / (ins_ms * 1e-3) / 1e9}
    roof["frac"] = roof["achieved"] / peak
    roof_q = {"bound": "hbm", "kernel": "BloomFilter query: bin_kernel_sort + probe_bins_kernel + finalize_hits_kernel", "achieved": qry_bytes / (qry_ms * 1e-3) / 1e9,
              "peak": peak, "unit": "GB/s", "bytes_per_kmer": 32 * H + 1, "kmers_per_launch": k_qry / S,
              "launch_ms": qry_ms, "call_ms_median": float(np.median(d_q)),
              "gkmers_s": k_qry / S / (qry_ms * 1e-3) / 1e9}
    roof_q["frac"] = roof_q["achieved"] / peak
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this
    # same command (profiles/r1_dram_bytes_per_launch.json, written by tools/summarize_ncu.py)
    prof = os.path.join(ROOT, "profiles", "r1_dram_bytes_per_step.json")
    if os.path.exists(prof) and chunk == CHUNK and args.query_factor == 4 and not args.opt:
        try:
            t = json.load(open(prof))
            roof["traffic"] = t["build_bytes_per_step"]
            roof_q["traffic"] = t["query_bytes_per_step"]
        except Exception:
            pass
    roof_q["peak_source"] = peak_src
    roof_q["share_of_step"] = ms_query / ms_total
    roof["share_of_step"] = ms_insert / ms_total
    roof["note"] = ("a fraction above 1 is possible: the partitioned build replaces one random 32-byte sector per hash "
                    "(the algorithmic model) by streaming traffic -- compare `traffic` with achieved x launch_ms")
    step_bytes = ins_bytes + qry_bytes
    roof_step = {"bound": "hbm", "kernel": "whole step (build phase + query phase)", "unit": "GB/s", "peak": peak,
                 "achieved": step_bytes / (ms_total / S * 1e-3) / 1e9, "launch_ms": ms_total / S,
                 "traffic": (roof["traffic"] + roof_q["traffic"]) if roof["traffic"] and roof_q["traffic"] else None}
    roof_step["frac"] = roof_step["achieved"] / peak
    # measured hardware ceilings for this access pattern (tools/probe_random_access.py on this GPU type)
    probe_file = os.path.join(ROOT, "profiles", "r1_random_access_probe.jsonl")
    if os.path.exists(probe_file):
        try:
            roof_q["random_access_probe"] = [json.loads(l) for l in open(probe_file) if l.strip()]
        except Exception:
            pass
    line = {"metric": "k-mers/s inserted+queried", "value": (k_ins_all + k_qry_all) / (ms_total * 1e-3) / 1e9,
            "unit": "Gk-mer/s", "n_gpus": world, "steps": S, "warmup": W, "ms_per_step": ms_total / S,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config, "insert_gkmers_s": k_ins_all / (ms_insert * 1e-3) / 1e9,
            "query_gkmers_s": k_qry_all / (ms_query * 1e-3) / 1e9, "kmers_per_step": (k_ins_all + k_qry_all) / S,
            # `roofline` is the phase that dominates the step (the query, ~3/4 of it); the build and the whole step follow
            "roofline": roof_q, "roofline_build": roof, "roofline_step": roof_step,
            "gpu_launches": launches_all, "clocks": clocks}
    if miss:
        line["query_miss_set"] = miss
    if e2e:
        line["e2e"] = {"value": e2e["kmers"] / e2e["seconds"] / 1e9, "unit": "Gk-mer/s", "steps": e2e["steps"],
                       "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "api": "btlbf_insert_seqs_async + btlbf_contains_seqs_async (streaming), pinned host buffers"}
        line["e2e_sync"] = {"value": e2e["kmers_sync"] / e2e["seconds_sync"] / 1e9, "unit": "Gk-mer/s",
                            "steps": e2e["steps_sync"], "api": "btlbf_insert_seqs + btlbf_contains_seqs (blocking)"}
    if merge:
        line["merge"] = merge
    if numa:
        line["config"]["host_numa_binding"] = numa
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(2, 1, args.cpu_sample)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "insert_gkmers_s",
                                                   "query_gkmers_s")}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
