#!/usr/bin/env python3
"""bench.py -- k-mer insert + query throughput of the Bloom-filter hot path on B200.

Headline workload (BASELINE.json configs[1], "cfg2"): 3 Gbp synthetic genome, k=25, 4 hashes, 31,568,113,856-bit
filter (the reference's calcOptimalSize(3e9, 0.01), 3.95 GB), built in 64 Mi-window chunks, then 150 bp read queries.
One STEP = one pass of the hot path over one batch: insert one genome chunk and query one batch of 150 bp reads
(4 x 64 MiB of bases) sampled from an inserted chunk (every k-mer present: no early exit).  Like the workload itself
(build the filter, then query reads) the timed region runs the K build batches, settles the filter -- at N > 1:
MERGES the per-GPU partial filters (one fused kernel per GPU over NVLink peer memory, bracketed by stream-ordered
cross-rank barriers) -- and then runs the K query batches.  Everything is inside the timed region.

  value     whole-job Gk-mer/s (inserted + queried, all ranks) with the batches already resident in HBM
            (btlbf_insert_seqs_dev / btlbf_contains_seqs_dev), CUDA events, max over ranks; at N > 1 the merge is
            part of it (`per_gpu_rate` x N is what N independent GPUs would do without it)
  e2e       the same steps through the host-buffer C-ABI calls: pinned host inputs, H2D + kernels + D2H of the hit
            bits (+ the merge at N > 1) inside the timed region
  roofline  the phase that dominates the step (the query): (32*h + 1) algorithmic bytes per k-mer / its mean
            duration per step against the measured HBM copy bandwidth; roofline_build and roofline_step follow
  job       (cfg2) the complete BASELINE job, strong-scaled: the 45 chunks of the 3 Gbp genome sharded over the N
            ranks -> merge -> 1e8 reads sharded over the ranks, one pair of events, max over ranks
  configs   (default run) the other BASELINE.json configs -- cfg3, cfg4 (counting filter), cfg5a / cfg5b (spaced
            seeds) -- through the same K-step loop with fewer steps: device rate, roofline fraction, e2e and the
            reference's CPU path beside each
  cpu_baseline  the reference's own CPU path (oracle/_ref: unmodified headers, OpenMP over reads / pieces) on a
            bounded sample of the same workload, same filter size, on this box's host cores

`--impl reference` times only that CPU path, K bounded-sample steps of --config.  `--config NAME` makes another
BASELINE config the headline of the line.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# host threads available to this process, taken before anything narrows the affinity mask or torchrun's
# OMP_NUM_THREADS=1 gets a say (the reference arm passes this number to OpenMP explicitly)
HOST_THREADS = len(os.sched_getaffinity(0))

G_LEN = 3_000_000_000
CHUNK = 64 << 20              # windows per insert batch
READ_LEN = 150
KMAX = 32                     # the genome buffers carry KMAX-1 halo bytes: every config inserts CHUNK k-mers per batch
GENOME_SEED, READ_SEED = 42, 7
JOB_READS = 100_000_000       # "then query >= 1e8 150 bp reads"
_SP = ["111101110111001", "111110110100111"]
SEEDS31 = [x + "1" + x[::-1] for x in _SP]  # two symmetric 31-character masks

CONFIGS = {
    "cfg2": dict(kind="bloom", k=25, h=4, size=31_568_113_856,  # BloomFilter::calcOptimalSize(3e9, 0.01), BloomFilter.hpp:406-413
                 workload="cfg2: 3 Gbp synthetic genome build (k=25, h=4, 31.57 Gbit filter) + 150 bp read query"),
    "cfg3": dict(kind="bloom", k=32, h=6, size=1 << 35,
                 workload="cfg3: 150 bp read query vs 4 GiB filter (k=32, h=6, 2^35 bits) built from the same genome"),
    "cfg4": dict(kind="counting", k=25, h=4, size=16_000_000_000, threshold=2,
                 workload="cfg4: CountingBloomFilter<uint8_t>, 16e9 counters, k=25, h=4, threshold-2 read query"),
    "cfg5a": dict(kind="bloom", k=31, h=2, size=1 << 29, seeds=SEEDS31, h2=1,
                  workload="cfg5a: spaced seeds (stHashIterator, k=31, 2 seeds), 64 MiB L2-resident filter"),
    "cfg5b": dict(kind="bloom", k=31, h=2, size=1 << 37, seeds=SEEDS31, h2=1,
                  workload="cfg5b: spaced seeds (stHashIterator, k=31, 2 seeds), 16 GiB filter"),
}


def filter_bytes(cfg):
    return cfg["size"] // 8 if cfg["kind"] == "bloom" else cfg["size"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def source_sha():
    """Hash of the kernel sources: profiles carry it so that a stale ncu figure is never attached to a new binary.
    (pack.cu and ingest.cu are host-only translation units -- the 2-bit packer and the FASTA/FASTQ parser -- and are left out.)"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "btl_bloomfilter_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".hpp")) and name not in ("pack.cu", "ingest.cu"):
            with open(os.path.join(d, name), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()[:16]


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
def mem_available():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 1 << 62


def cpu_reference_run(cfg, n_steps, warmup, sample_bases, threads=0, verbose=False):
    """The reference's CPU path on a bounded sample per step: insert `sample_bases` of the genome (64 kb pieces,
    OpenMP over pieces; README.md:30-43 / 86-113, stHashIterator.hpp:53-57 for spaced seeds) + query
    sample_bases/150 reads sampled from it, against a filter of the config's full size."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    orc = O.Oracle()
    use_ref = O.Ref.available()
    k, h, kind = cfg["k"], cfg["h"], cfg["kind"]
    seeds = cfg.get("seeds")
    if not use_ref and (kind != "bloom" or seeds):
        return {"unavailable": "oracle/_ref (the compiled reference) is needed for the counting / spaced-seed CPU legs"}
    need = filter_bytes(cfg) + (2 << 30)
    if mem_available() < need:
        return {"unavailable": "host has %.1f GB available, the %s-size filter needs %.1f GB" %
                (mem_available() / 1e9, cfg["kind"], need / 1e9)}
    piece = 65536
    threads = int(threads or HOST_THREADS)
    if use_ref:
        R = O.Ref()
        filt = R.L.ref_cbf_new(cfg["size"], h, k, cfg.get("threshold", 1)) if kind == "counting" else R.bf_new(cfg["size"], h, k)
        sp = R._seeds(seeds) if seeds else None
    else:
        filt = np.zeros(cfg["size"] // 8, np.uint8)
    n_reads = sample_bases // READ_LEN
    roff = (READ_LEN * np.arange(n_reads + 1)).astype(np.uint64)
    tot_k, tot_t, per = 0, 0.0, []

    def leg(bases, off, n_seqs, do_insert):
        nk, nh = O.u64(), O.u64()
        if not use_ref:
            t = orc.L.ora_bench_bf(O._p8(filt), cfg["size"], h, k, O._p8(bases), O._p64(off), n_seqs, do_insert, threads,
                                   C.byref(nk), C.byref(nh))
        elif kind == "counting":
            t = R.L.ref_bench_cbf(filt, O._p8(bases), O._p64(off), n_seqs, do_insert, threads, C.byref(nk), C.byref(nh))
        elif seeds:
            t = R.L.ref_bench_st_bf(filt, sp, len(seeds), cfg.get("h2", 1), O._p8(bases), O._p64(off), n_seqs, do_insert,
                                    threads, C.byref(nk), C.byref(nh))
        else:
            t = R.L.ref_bench_bf(filt, O._p8(bases), O._p64(off), n_seqs, do_insert, threads, C.byref(nk), C.byref(nh))
        return t, nk.value, nh.value

    for s in range(warmup + n_steps):
        g0 = (s * sample_bases) % (G_LEN - sample_bases - k)
        g = orc.synth_genome(g0, sample_bases + k - 1, GENOME_SEED)
        # 64 kb pieces overlapping by k-1 so that every window is inserted exactly once
        starts = np.arange(0, sample_bases, piece, dtype=np.uint64)
        pieces = [g[int(a): int(min(a + piece + k - 1, g.size))] for a in starts]
        pb = np.concatenate(pieces)
        poff = np.concatenate([[0], np.cumsum([p.size for p in pieces])]).astype(np.uint64)
        reads = orc.synth_reads(0, n_reads, READ_LEN, sample_bases, GENOME_SEED, READ_SEED + s, g_start=g0)
        t_i, k_i, _ = leg(pb, poff, poff.size - 1, 1)
        t_q, k_q, n_hit = leg(reads, roff, n_reads, 0)
        if kind == "bloom":
            assert n_hit == k_q, "CPU reference: a k-mer of an inserted region was not found"
        if s >= warmup:
            tot_k += k_i + k_q
            tot_t += t_i + t_q
            per.append((k_i / t_i, k_q / t_q))
        if verbose:
            print("cpu step %d: insert %.2f Mk/s, query %.2f Mk/s" % (s, k_i / t_i / 1e6, k_q / t_q / 1e6), file=sys.stderr)
    if use_ref:
        (R.L.ref_cbf_free if kind == "counting" else R.L.ref_bf_free)(filt)
    return {"value": tot_k / tot_t / 1e9, "unit": "Gk-mer/s", "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": "%d steps x (%d bp genome insert in 64 kb pieces + %d reads x %d bp query), %s filter of %d %s, "
                      "OpenMP over pieces/reads" % (n_steps, sample_bases, n_reads, READ_LEN, kind, cfg["size"],
                                                    "bits" if kind == "bloom" else "counters"),
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


# ---------------------------------------------------------------- the GPU side
class Env:
    """Everything the configs share: the device, the stream, the context, the synthetic inputs in HBM."""

    def __init__(self, args, torch, dist, B):
        self.torch, self.dist, self.B, self.args = torch, dist, B, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.numa = bind_to_gpu_numa_node(torch, self.local_rank) if self.world > 1 else None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = B.Context(self.local_rank)
        # the library's kernels run on torch's current stream so that torch.cuda.Event brackets them
        self.stream = torch.cuda.Stream(device=self.dev, priority=args.stream_priority)
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.options = {}
        if args.l2_fetch:
            self.ctx.set_option("l2_fetch_granularity", args.l2_fetch)
        for kv in args.opt:
            key, val = kv.split("=")
            self.ctx.set_option(key, int(val))
            self.options[key] = int(val)
        self.chunk = args.chunk // 16384 * 16384
        self.n_chunks = (G_LEN + self.chunk - 1) // self.chunk
        self.n_reads = self.chunk * args.query_factor // READ_LEN
        self.read_bases = self.n_reads * READ_LEN
        self.token = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.token2 = torch.zeros(1, dtype=torch.int32, device=self.dev)  # barrier word of the side stream
        self.side = torch.cuda.Stream(device=self.dev) if self.world > 1 else None  # merges pipelined behind pass 2
        self.host = None
        self.no_symm = None      # why symmetric memory is not used (set on first failure)
        self.merge_choice = {}   # (kind, size) -> (kernel, calibration) of the N > 1 merge
        if self.world > 1:
            self.ctx.set_option("wrap_accumulate", 1)  # symmetric-memory filters keep the build's accumulation

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def stream_barrier(self):
        """Cross-rank barrier in stream order (no host synchronisation): a one-word NCCL all-reduce."""
        if self.world > 1:
            self.dist.all_reduce(self.token)

    def make_inputs(self, n_genome, n_read):
        """Genome chunks of this rank's shard (chunk c = rank + world * i) and read batches sampled from the first of
        them, generated in HBM (replayable by the oracle: tests/test_gpu_parity.py)."""
        torch, ctx = self.torch, self.ctx
        self.g, self.g_chunk, self.g_len, self.r = [], [], [], []
        for i in range(n_genome):
            c = (self.rank + self.world * i) % self.n_chunks
            g0 = c * self.chunk
            glen = min(self.chunk + KMAX - 1, G_LEN - g0)
            t = torch.empty(self.chunk + 64, dtype=torch.uint8, device=self.dev)
            ctx.synth_genome_device(t.data_ptr(), g0, glen, GENOME_SEED)
            self.g.append(t)
            self.g_chunk.append(c)
            self.g_len.append(glen)
        for j in range(n_read):
            c = self.g_chunk[j]
            t = torch.empty(self.read_bases + 64, dtype=torch.uint8, device=self.dev)
            ctx.synth_reads_device(t.data_ptr(), 0, self.n_reads, READ_LEN, c * self.chunk,
                                   min(self.chunk, G_LEN - c * self.chunk), GENOME_SEED, READ_SEED + c)
            self.r.append(t)
        self.d_roff = torch.arange(0, self.read_bases + 1, READ_LEN, dtype=torch.int64, device=self.dev)
        self.d_hits = torch.zeros((self.read_bases + 31) // 32 + 8, dtype=torch.int32, device=self.dev)
        self._goff = {}
        torch.cuda.synchronize()

    def goff(self, n):
        if n not in self._goff:
            self._goff[n] = self.torch.tensor([0, n], dtype=self.torch.int64, device=self.dev)
        return self._goff[n]

    def insert_len(self, i, k):
        """bases of genome buffer i a config with k-mer size k inserts: chunk windows + the k-1 halo"""
        return min(self.chunk + k - 1, self.g_len[i])

    def make_host_inputs(self, n_host):
        """Pinned host copies of the first n_host genome chunks / read batches (shared by every config's e2e leg)."""
        if self.host is not None:
            return self.host
        torch = self.torch
        H = {"n": n_host}
        H["genome"] = [torch.empty(self.g_len[j], dtype=torch.uint8).pin_memory() for j in range(n_host)]
        H["reads"] = [torch.empty(self.read_bases, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        for j in range(n_host):
            H["genome"][j].copy_(self.g[j][: self.g_len[j]])
            H["reads"][j].copy_(self.r[j][: self.read_bases])
        H["hits"] = [torch.zeros((self.read_bases + 31) // 32 * 4, dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        t_roff = torch.arange(0, self.read_bases + 1, READ_LEN, dtype=torch.int64).pin_memory()
        H["t_roff"] = t_roff
        H["roff"] = t_roff.numpy().view(np.uint64)
        H["counts_t"] = torch.zeros((64, 4), dtype=torch.int64).pin_memory()
        H["counts"] = H["counts_t"].numpy().view(np.uint64)
        # the same batches as 2 bits per base (btlbf_pack_seqs: the library's host packer), in pinned memory; the
        # synthetic batches hold no invalid base, so the invalid plane is dropped (as the packer's caller would)
        if not self.args.no_packed:
            B = self.B
            H["genome_pk_t"], H["reads_pk_t"], H["genome_pk"], H["reads_pk"] = [], [], {}, []
            scratch = np.zeros((max(max(self.g_len[:n_host]), self.read_bases) + 7) // 8, np.uint8)
            dt, nb = 0.0, 0
            for j in range(n_host):
                tr = torch.zeros((self.read_bases + 3) // 4 + 16, dtype=torch.uint8).pin_memory()
                t0 = time.perf_counter()  # (the packer alone: the pinned buffer exists and its pages are touched)
                pk = B.pack_seqs((H["reads"][j].numpy(), H["roff"]), codes_out=tr.numpy(), invalid_out=scratch)
                dt += time.perf_counter() - t0
                assert pk.invalid is None
                H["reads_pk_t"].append(tr)
                H["reads_pk"].append(pk)
                tg = torch.zeros((self.g_len[j] + 3) // 4 + 16, dtype=torch.uint8).pin_memory()
                H["genome_pk_t"].append(tg)
                nb += self.read_bases
            H["pack_reads_gbases_s"] = nb / dt / 1e9
        torch.cuda.synchronize()
        self.host = H
        return H

    def packed_genome(self, j, n):
        """PackedBatch of the first n bases of host genome chunk j (k differs between the configs)."""
        H = self.host
        key = (j, n)
        if key not in H["genome_pk"]:
            scratch = np.zeros((n + 7) // 8, np.uint8)
            pk = self.B.pack_seqs((H["genome"][j].numpy()[:n], np.array([0, n], np.uint64)),
                                  codes_out=H["genome_pk_t"][j].numpy(), invalid_out=scratch)
            assert pk.invalid is None
            H["genome_pk"] = {kk: v for kk, v in H["genome_pk"].items() if kk[0] != j}
            H["genome_pk"][key] = pk
        return H["genome_pk"][key]


def make_filter(env, cfg):
    B = env.B
    if cfg["kind"] == "counting":
        return B.CountingBloomFilter(cfg["size"], cfg["h"], cfg["k"], cfg.get("threshold", 1), ctx=env.ctx)
    f = B.BloomFilter(cfg["size"], cfg["h"], cfg["k"], ctx=env.ctx)
    if cfg.get("seeds"):
        f.setSeeds(cfg["seeds"], cfg.get("h2", 1))
    return f


def make_sharded_filter(env, cfg):
    """N > 1: this rank's partial filter + the object that merges the N of them.  The filters live in symmetric memory
    (torch.distributed._symmetric_memory: peer pointers, and an NVLS multicast mapping where the box has one).
    BloomFilters are merged by whichever kernel the box runs faster at this N -- the OR inside NVSwitch
    (btlbf_merge_multimem) or the peer-memory kernel (btlbf_merge_peers); timed once per filter size, before
    anything else is -- counting filters by the peer-memory kernel (saturating add has no multimem form).  Without
    symmetric memory: the library's own allocation, CUDA IPC and the peer-memory kernel."""
    from btl_bloomfilter_b200 import parallel
    B, want = env.B, env.args.merge
    cls = B.BloomFilter if cfg["kind"] == "bloom" else B.CountingBloomFilter
    if want != "ipc" and not env.no_symm:
        try:
            f, hdl = parallel.symmetric_filter(cls, cfg["size"], cfg["h"], cfg["k"], env.ctx, threshold=cfg.get("threshold", 1))
            if cfg.get("seeds"):
                f.setSeeds(cfg["seeds"], cfg.get("h2", 1))
            pm = parallel.MultimemMerge(env.ctx, hdl, filter_bytes(cfg), f.KIND, mode=None if want == "auto" else want)
            key = (cfg["kind"], cfg["size"])
            if want == "auto" and cfg["kind"] == "bloom":
                if key not in env.merge_choice:
                    pm.calibrate()
                    env.merge_choice[key] = (pm.mode, pm.calibration)
                    f.clear()
                pm.mode, pm.calibration = env.merge_choice[key]
            return f, pm, pm.mode
        except Exception as e:  # noqa: BLE001  symmetric memory unavailable: every rank takes the same branch
            if want in ("multimem", "peer"):
                raise SystemExit("--merge %s: %s" % (want, str(e)[:200]))
            env.no_symm = "%s: %s" % (type(e).__name__, str(e)[:120])
    f = make_filter(env, cfg)
    return f, parallel.PeerMerge(env.ctx, *f.device_ptr(), f.KIND), "peer"


MERGE_HOW = {
    "multimem": "btlbf_merge_multimem: one kernel per GPU, multimem.ld_reduce.or + multimem.st over an NVLS multicast mapping "
                "of the partial filters (symmetric memory); the OR happens inside NVSwitch",
    "peer": "btlbf_merge_peers: one kernel per GPU over NVLink peer memory (reduce-scatter + all-gather in one pass)",
    "hybrid": "btlbf_merge_hybrid: one kernel per GPU, part of its byte range reduced inside NVSwitch (multimem), the rest over "
              "NVLink peer memory, side by side",
}


def filter_view(env, filt):
    from btl_bloomfilter_b200 import parallel
    ptr, nbytes = filt.device_ptr()
    return parallel.device_tensor_from_ptr(ptr, nbytes, env.dev)


def merge_parity(env, cfg):
    """N > 1, before anything is timed: one chunk per rank into per-GPU partial filters of the config's full size,
    fused merge; every rank then builds the same N chunks alone (BloomFilter: into one filter; counting: N partial
    builds + saturating add, the definition of the sharded counting build) and compares the two arrays byte by byte
    on the device.  Single-GPU builds are what tests/ pin to the oracle."""
    from btl_bloomfilter_b200 import parallel
    from btl_bloomfilter_b200._capi import check
    torch, ctx = env.torch, env.ctx
    k = cfg["k"]
    a, pm, how = make_sharded_filter(env, cfg)
    st = torch.zeros(2, dtype=torch.int64, device=env.dev)
    n0 = env.insert_len(0, k)
    a.insertSeqsDevice(env.g[0].data_ptr(), n0, env.goff(n0).data_ptr(), 1, st.data_ptr())
    pm.merge()
    pm.close()
    b = make_filter(env, cfg)
    part = make_filter(env, cfg) if cfg["kind"] == "counting" else None
    tmp = torch.empty(env.chunk + 64, dtype=torch.uint8, device=env.dev)
    for r in range(env.world):
        c = r % env.n_chunks
        glen = min(env.chunk + k - 1, G_LEN - c * env.chunk)
        ctx.synth_genome_device(tmp.data_ptr(), c * env.chunk, glen, GENOME_SEED)
        if part is None:
            b.insertSeqsDevice(tmp.data_ptr(), glen, env.goff(glen).data_ptr(), 1, 0)
        else:
            part.clear()
            part.insertSeqsDevice(tmp.data_ptr(), glen, env.goff(glen).data_ptr(), 1, 0)
            ptr, nbytes = part.device_ptr()
            check(ctx.L.btlbf_filter_merge_from_device(b._h, C.c_void_p(ptr), nbytes))
    same = bool(torch.equal(filter_view(env, a), filter_view(env, b)))
    pop = a.getPop() if cfg["kind"] == "bloom" else a.popCount()
    t = torch.tensor([1 if same else 0, pop, -pop], dtype=torch.int64, device=env.dev)
    env.dist.all_reduce(t, op=env.dist.ReduceOp.MIN)
    del a, b, part, tmp, pm
    return {"merge_parity": bool(t[0] == 1), "merge": how, "identical_popcount_on_all_ranks": bool(int(t[1]) == -int(t[2])),
            "popcount": int(t[1]), "chunks": env.world,
            "how": "sharded build + fused merge == the same chunks built on one GPU (byte-compared on the device)"}


def run_config(env, name, cfg, S, W, headline):
    """The K-step loop of one config: K build batches -> settle (-> merge at N > 1) -> K query batches."""
    torch, dist, ctx, world = env.torch, env.dist, env.ctx, env.world
    k, h, kind = cfg["k"], cfg["h"], cfg["kind"]
    counting = kind == "counting"
    n_g, n_r = len(env.g), len(env.r)
    n_q = max(1, min(n_r, S))  # the read buffers the timed queries cycle over: their chunks are inserted by then
    out = {"workload": cfg["workload"], "k": k, "hashes": h, "filter_bytes": filter_bytes(cfg)}
    pm, merge_how = None, None
    if world > 1:
        if headline or counting:
            out["merge_check"] = merge_parity(env, cfg)
        filt, pm, merge_how = make_sharded_filter(env, cfg)
    else:
        filt = make_filter(env, cfg)
    d_stats = torch.zeros(4, dtype=torch.int64, device=env.dev)

    def build_dev(i):
        j = i % n_g
        n = env.insert_len(j, k)
        filt.insertSeqsDevice(env.g[j].data_ptr(), n, env.goff(n).data_ptr(), 1, d_stats.data_ptr())

    def query_dev(i):
        j = i % n_q
        filt.containsSeqsDevice(env.r[j].data_ptr(), env.read_bases, env.d_roff.data_ptr(), env.n_reads,
                                env.d_hits.data_ptr(), 0, d_stats[2:].data_ptr())

    chunks = env.args.merge_chunks if (pm is not None and hasattr(pm, "launch_range") and kind == "bloom") else 0

    def settle_and_merge(ev_applied=None):
        """pass 2 of whatever the build has parked, then (N > 1) the merge: one after the other, or -- merge_chunks > 0
        -- pipelined: the merge of a chunk of partitions runs over NVLink while the next chunk is applied"""
        if chunks > 0:
            from btl_bloomfilter_b200 import parallel
            parallel.pipelined_flush_merge(filt, pm, chunks, env.side, env.token2, ev_applied)
            return
        ctx.flush()  # every k-mer parked in the partition buckets reaches the filter (pass 2)
        if ev_applied is not None:
            ev_applied.record(env.stream)
        if pm is None:
            return
        env.stream_barrier()  # every rank's partial build is complete
        pm.launch()
        env.stream_barrier()  # every rank's byte range has been written into every filter

    # ---- warm-up.  Counting filters: the chunks behind the read batches are inserted once here and once more in the
    # timed region (2x coverage), so that every queried k-mer reaches the threshold (a hit set, like the BloomFilter's)
    for i in range(n_q if counting else W):
        build_dev(i)
    if counting:
        ctx.flush()  # (counting: one merge only, the timed one -- a second saturating add would count every k-mer N times over)
    else:
        settle_and_merge()
    for i in range(W):
        query_dev(i)
    torch.cuda.synchronize()
    if not counting:
        filt.clear()
    d_stats.zero_()
    ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    ev_q = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    t_start, t_built, t_mid, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    sampler = ClockSampler(env.local_rank)
    if env.rank == 0 and headline:
        sampler.start()
    launches0 = ctx.launch_count
    env.barrier()
    torch.cuda.synchronize()
    t_start.record(env.stream)
    for i in range(S):
        build_dev(i)
        ev_b[i].record(env.stream)
    settle_and_merge(t_built)
    t_mid.record(env.stream)
    for i in range(S):
        query_dev(i)
        ev_q[i].record(env.stream)
    t_end.record(env.stream)
    torch.cuda.synchronize()
    env.barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if (env.rank == 0 and headline) else None
    ms_total = t_start.elapsed_time(t_end)
    ms_build = t_start.elapsed_time(t_built)
    ms_merge = t_built.elapsed_time(t_mid)
    ms_query = t_mid.elapsed_time(t_end)
    d_b = [(t_start if i == 0 else ev_b[i - 1]).elapsed_time(ev_b[i]) for i in range(S)]
    d_q = [(t_mid if i == 0 else ev_q[i - 1]).elapsed_time(ev_q[i]) for i in range(S)]
    st = d_stats.cpu().numpy()
    k_ins, k_qry, k_hit = int(st[0]), int(st[2]), int(st[3])
    if kind == "bloom":
        assert k_hit == k_qry, "%s: a k-mer of an inserted chunk was not found (%d of %d)" % (name, k_hit, k_qry)
    out["query_hit_fraction"] = k_hit / max(1, k_qry)
    if counting:
        out["ordered_deferred_rounds"] = list(filt.orderedStats())

    # ---- the miss set: reads of independent random bases, almost every k-mer absent
    miss = None
    if env.rank == 0 and headline:
        d_miss = torch.empty(env.read_bases + 64, dtype=torch.uint8, device=env.dev)
        ctx.synth_genome_device(d_miss.data_ptr(), 0, env.read_bases, 43 << 40)
        d_ms = torch.zeros(2, dtype=torch.int64, device=env.dev)
        reps = 3
        qm = lambda: filt.containsSeqsDevice(d_miss.data_ptr(), env.read_bases, env.d_roff.data_ptr(), env.n_reads,  # noqa: E731
                                             env.d_hits.data_ptr(), 0, d_ms.data_ptr())
        qm()
        torch.cuda.synchronize()
        d_ms.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(env.stream)
        for _ in range(reps):
            qm()
        b.record(env.stream)
        torch.cuda.synchronize()
        mk, mh = [int(x) for x in d_ms.cpu().numpy()]
        ms_miss = a.elapsed_time(b) / reps
        miss = {"gkmers_s": mk / reps / (ms_miss * 1e-3) / 1e9, "ms_per_batch": ms_miss, "kmers_per_batch": mk // reps,
                "hit_fraction": mh / max(1, mk),
                "path": "adaptive: sampled hit fraction on the device picks the early-exit kernel for read sets that mostly miss"}
        del d_miss

    # ---- end to end through the host-buffer C ABI (pinned inputs; H2D, kernels, D2H -- and the merge -- in the timed
    # region).  Streaming form (btlbf_insert_seqs_async / btlbf_contains_seqs_async): the calls are queued back to back,
    # so the copies of one step overlap the kernels of another; results are checked after the final btlbf_ctx_sync.
    e2e = None
    if not env.args.no_e2e:
        Hh = env.make_host_inputs(min(n_r, 4))
        n_host = Hh["n"]
        counts = Hh["counts"]
        S2 = min(S, 32 if headline else 6)
        nh_q = max(1, min(n_host, S2))
        goffs = [np.array([0, env.insert_len(j, k)], dtype=np.uint64) for j in range(n_host)]

        def build_host_async(i, slot):
            j = i % n_host
            filt.insertSeqsAsync((Hh["genome"][j].numpy()[: int(goffs[j][1])], goffs[j]), counts[slot, 0:2])

        def query_host_async(i, slot):
            j = i % nh_q
            filt.containsSeqsAsync((Hh["reads"][j].numpy(), Hh["roff"]), Hh["hits"][j].numpy(), counts[slot, 2:4])

        def merge_host():
            if pm is not None:
                pm.merge()  # flush + synchronise + barrier | kernel | synchronise + barrier

        # warm-up of the host path: the pinned buffers, the copy engines and the PCIe link (which trains up under
        # traffic) -- a fresh process on a fresh box measures ~20 % low without it
        for rep in range(2 if counting else 3):
            for i in range(min(4, S2)):
                build_host_async(i, 60 + (i & 1))
            ctx.sync()
            merge_host()
            for i in range(min(4, S2)):
                query_host_async(i, 60 + (i & 1))
            ctx.sync()
        counts[:] = 0
        env.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(S2):
            build_host_async(i, i)
        if pm is not None:
            ctx.sync()
            merge_host()
        for i in range(S2):
            query_host_async(i, i)
        ctx.sync()
        dt = time.perf_counter() - t0
        env.barrier()
        ke = int(counts[:S2, 0].sum() + counts[:S2, 2].sum())
        if kind == "bloom":
            assert np.array_equal(counts[:S2, 2], counts[:S2, 3]) and counts[:S2, 2].all(), "e2e: a queried k-mer was not found"
        e2e = {"kmers": ke, "seconds": dt, "steps": S2,
               "h2d": int(goffs[0][1]) + env.read_bases + (env.n_reads + 1) * 8 + 16, "d2h": int(Hh["hits"][0].numel()) + 32}
        if not env.args.no_packed:
            # the same steps from 2-bit packed host buffers (btlbf_*_seqs_packed_async): a quarter of the H2D bytes
            gpk = [env.packed_genome(j, int(goffs[j][1])) for j in range(n_host)]

            def build_host_packed(i, slot):
                filt.insertSeqsPackedAsync(gpk[i % n_host], counts[slot, 0:2])

            def query_host_packed(i, slot):
                j = i % nh_q
                filt.containsSeqsPackedAsync(Hh["reads_pk"][j], Hh["hits"][j].numpy(), counts[slot, 2:4])

            if not counting:
                filt.clear()
            for i in range(min(2, S2)):
                build_host_packed(i, 60 + (i & 1))
            ctx.sync()
            if not counting:
                merge_host()
            for i in range(min(2, S2)):
                query_host_packed(i, 60 + (i & 1))
            ctx.sync()
            if not counting:
                filt.clear()
            counts[:] = 0
            env.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(S2):
                build_host_packed(i, i)
            if pm is not None:
                ctx.sync()
                merge_host()
            for i in range(S2):
                query_host_packed(i, i)
            ctx.sync()
            dtp = time.perf_counter() - t0
            env.barrier()
            kp = int(counts[:S2, 0].sum() + counts[:S2, 2].sum())
            assert kp == ke, "packed e2e: %d k-mers, the ASCII calls saw %d" % (kp, ke)
            if kind == "bloom":
                assert np.array_equal(counts[:S2, 2], counts[:S2, 3]), "packed e2e: a queried k-mer was not found"
            e2e.update({"kmers_packed": kp, "seconds_packed": dtp,
                        "h2d_packed": (int(goffs[0][1]) + 3) // 4 + (env.read_bases + 3) // 4 + (env.n_reads + 1) * 8 + 16})
        if not env.args.no_packed and not counting:
            # the ASCII calls again with the library packing every chunk on host threads before the H2D copy (context
            # option host_pack): same API and buffers as `e2e`, a quarter of the PCIe bytes, host cores spent on packing
            try:
                ctx.set_option("host_pack", 1)
                ctx.set_option("host_pack_threads", max(2, min(16, HOST_THREADS // max(1, world))))
                filt.clear()
                for i in range(min(2, S2)):
                    build_host_async(i, 60 + (i & 1))
                ctx.sync()
                merge_host()
                for i in range(min(2, S2)):
                    query_host_async(i, 60 + (i & 1))
                ctx.sync()
                filt.clear()
                counts[:] = 0
                env.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(S2):
                    build_host_async(i, i)
                if pm is not None:
                    ctx.sync()
                    merge_host()
                for i in range(S2):
                    query_host_async(i, i)
                ctx.sync()
                dth = time.perf_counter() - t0
                env.barrier()
                kh = int(counts[:S2, 0].sum() + counts[:S2, 2].sum())
                assert kh == ke and np.array_equal(counts[:S2, 2], counts[:S2, 3]), "host_pack e2e: results differ"
                e2e.update({"kmers_hostpack": kh, "seconds_hostpack": dth,
                            "threads_hostpack": max(2, min(16, HOST_THREADS // max(1, world)))})
            except Exception as ex:  # noqa: BLE001  (an optional leg must not take the line with it)
                e2e["hostpack_error"] = "%s: %s" % (type(ex).__name__, str(ex)[:200])
            finally:
                ctx.set_option("host_pack", 0)
        if headline:
            def step_host_sync(i):
                j = i % n_host
                a = filt.insertSeqs((Hh["genome"][j].numpy()[: int(goffs[j][1])], goffs[j]))
                r = filt.containsSeqs((Hh["reads"][j % nh_q].numpy(), Hh["roff"]), hit_out=Hh["hits"][j].numpy(), want_valid=False)
                return a, r.n_kmers, r.n_hits
            S3 = min(S, 6)
            t0 = time.perf_counter()
            ks = 0
            for i in range(S3):
                a, q, hq = step_host_sync(i)
                ks += a + q
            e2e.update({"kmers_sync": ks, "seconds_sync": time.perf_counter() - t0, "steps_sync": S3})
            env.barrier()

    # ---- reductions over ranks (max time, summed work)
    if world > 1:
        t = torch.tensor([ms_total, ms_build, ms_merge, ms_query, e2e["seconds"] if e2e else 0.0,
                          e2e.get("seconds_sync", 0.0) if e2e else 0.0, e2e.get("seconds_packed", 0.0) if e2e else 0.0,
                          e2e.get("seconds_hostpack", 0.0) if e2e else 0.0], dtype=torch.float64, device=env.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_build, ms_merge, ms_query = (float(x) for x in t[:4])
        w = torch.tensor([k_ins, k_qry, e2e["kmers"] if e2e else 0, launches, e2e.get("kmers_sync", 0) if e2e else 0,
                          e2e.get("kmers_packed", 0) if e2e else 0, e2e.get("kmers_hostpack", 0) if e2e else 0],
                         dtype=torch.int64, device=env.dev)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        k_ins_all, k_qry_all, ke_all, launches_all, ks_all, kp_all, kh_all = [int(x) for x in w]
        if e2e:
            e2e["seconds"], e2e["kmers"] = float(t[4]), ke_all
            if "seconds_sync" in e2e:
                e2e["seconds_sync"], e2e["kmers_sync"] = float(t[5]), ks_all
            if "seconds_packed" in e2e:
                e2e["seconds_packed"], e2e["kmers_packed"] = float(t[6]), kp_all
            if "seconds_hostpack" in e2e:
                e2e["seconds_hostpack"], e2e["kmers_hostpack"] = float(t[7]), kh_all
    else:
        k_ins_all, k_qry_all, launches_all = k_ins, k_qry, launches

    peak, peak_src = measured_peak()
    # algorithmic bytes of SURVEY.md 8d: build 64 B per hash (32 B sector read + 32 B dirty write-back) + 1 input byte
    # per k-mer, query 32 B per hash + 1
    b_ins, b_qry = 64 * h + 1, 32 * h + 1
    ins_gk = k_ins / (ms_build * 1e-3) / 1e9   # this rank's phases (rank 0 reports; ranks are symmetric)
    qry_gk = k_qry / (ms_query * 1e-3) / 1e9
    out.update({"value": (k_ins_all + k_qry_all) / (ms_total * 1e-3) / 1e9, "unit": "Gk-mer/s", "steps": S, "warmup": W,
                "ms_per_step": ms_total / S, "insert_gkmers_s": k_ins_all / (ms_build * 1e-3) / 1e9,
                "query_gkmers_s": k_qry_all / (ms_query * 1e-3) / 1e9, "kmers_per_step": (k_ins_all + k_qry_all) / S,
                "build_ms": ms_build, "merge_ms": ms_merge, "query_ms": ms_query,
                "insert_frac": ins_gk * b_ins / peak, "query_frac": qry_gk * b_qry / peak,
                "bytes_per_kmer": {"insert": b_ins, "query": b_qry}, "gpu_launches": launches_all})
    if world > 1:
        # what the ranks do when nothing connects them: the same phases without the merge
        out["per_gpu_rate"] = (k_ins + k_qry) / ((ms_build + ms_query) * 1e-3) / 1e9
        mm_share = 1.0 if merge_how == "multimem" else int(merge_how[6:]) / 100.0 if merge_how.startswith("hybrid") else 0.0
        per_dir = filter_bytes(cfg) * (mm_share * (1.0 + 1.0 / world) + (1.0 - mm_share) * 2.0 * (world - 1) / world)
        out["merge"] = {"ms": ms_merge, "filter_bytes": filter_bytes(cfg), "kind": merge_how,
                        "how": MERGE_HOW["hybrid" if merge_how.startswith("hybrid") else merge_how] + ", two stream-ordered barriers",
                        "link_bytes_per_gpu_per_direction": per_dir,
                        "link_GBps_per_gpu_per_direction": per_dir / (ms_merge * 1e-3) / 1e9}
        if chunks > 0:
            out["merge"]["pipelined_chunks"] = chunks
            out["merge"]["how"] = ("pass 2 of the build and btlbf_merge_peers_range pipelined in %d chunks of partitions "
                                   "(btlbf_filter_flush_parts): `ms` is what the merge adds behind the last chunk of pass 2" % chunks)
        if getattr(pm, "calibration", None):
            out["merge"]["calibration_ms"] = pm.calibration
        if env.no_symm:
            out["merge"]["symmetric_memory_unavailable"] = str(env.no_symm)
    if e2e:
        out["e2e"] = {"value": e2e["kmers"] / e2e["seconds"] / 1e9, "unit": "Gk-mer/s", "steps": e2e["steps"],
                      "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                      "api": "btlbf_insert_seqs_async + btlbf_contains_seqs_async (streaming), pinned host buffers"
                             + (", btlbf_merge_peers between the phases" if world > 1 else "")}
        if "seconds_packed" in e2e:
            out["e2e_packed"] = {"value": e2e["kmers_packed"] / e2e["seconds_packed"] / 1e9, "unit": "Gk-mer/s",
                                 "steps": e2e["steps"], "h2d_bytes_per_step": e2e["h2d_packed"],
                                 "d2h_bytes_per_step": e2e["d2h"],
                                 "api": "btlbf_insert_seqs_packed_async + btlbf_contains_seqs_packed_async: the same batches "
                                        "held by the caller as 2 bits per base (packed once, outside the timed region, by "
                                        "btlbf_pack_seqs at %.2f Gbase/s on this host)" % env.host.get("pack_reads_gbases_s", 0.0)}
        if "seconds_hostpack" in e2e:
            out["e2e_hostpack"] = {"value": e2e["kmers_hostpack"] / e2e["seconds_hostpack"] / 1e9, "unit": "Gk-mer/s",
                                   "steps": e2e["steps"], "h2d_bytes_per_step": e2e.get("h2d_packed", 0),
                                   "d2h_bytes_per_step": e2e["d2h"], "host_threads_per_gpu": e2e["threads_hostpack"],
                                   "api": "the calls and ASCII host buffers of `e2e`, with context option host_pack = 1: the "
                                          "library packs every chunk to 2 bits per base on host threads before the H2D copy "
                                          "(packing inside the timed region)"}
        elif "hostpack_error" in e2e:
            out["e2e_hostpack"] = {"error": e2e["hostpack_error"]}
        if "seconds_sync" in e2e:
            out["e2e_sync"] = {"value": e2e["kmers_sync"] / e2e["seconds_sync"] / 1e9, "unit": "Gk-mer/s",
                               "steps": e2e["steps_sync"], "api": "btlbf_insert_seqs + btlbf_contains_seqs (blocking)"}
    if headline:
        ins_bytes, qry_bytes = b_ins * (k_ins / S), b_qry * (k_qry / S)
        ins_ms, qry_ms = ms_build / S, ms_query / S
        pass1 = float(np.median(d_b))
        roof_b = {"bound": "hbm", "kernel": "build: bin_kernel_sort per batch + apply_bins_kernel per accumulation (all "
                  "durations of the build phase added)" if not counting else "counting build: OP_RESV_TOUCH + OP_CBF_COMMIT "
                  "+ list_drain_coop_kernel per 4 Mi-window batch", "achieved": ins_bytes / (ins_ms * 1e-3) / 1e9,
                  "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None, "bytes_per_kmer": b_ins,
                  "kmers_per_launch": k_ins / S, "launch_ms": ins_ms, "pass1_ms": pass1, "pass2_ms": ins_ms - pass1,
                  "gkmers_s": ins_gk, "share_of_step": ms_build / ms_total}
        roof_b["frac"] = roof_b["achieved"] / peak
        roof_b["note"] = ("a fraction above 1 is possible: the partitioned build replaces one random 32-byte sector per hash "
                          "(the algorithmic model) by streaming traffic -- compare `traffic` with achieved x launch_ms")
        roof_q = {"bound": "hbm", "kernel": "query: bin_kernel_sort<...,QUERY> (pass 1: hash + counting sort by filter partition) + "
                  "probe_bins_kernel (pass 2: L2-resident partition probes) + finalize_hits_kernel; all durations of the query "
                  "phase added" if not counting else "counting query: the same two passes with byte counters against the threshold", "achieved": qry_bytes / (qry_ms * 1e-3) / 1e9,
                  "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": None, "bytes_per_kmer": b_qry,
                  "kmers_per_launch": k_qry / S, "launch_ms": qry_ms, "call_ms_median": float(np.median(d_q)),
                  "gkmers_s": qry_gk, "share_of_step": ms_query / ms_total}
        roof_q["frac"] = roof_q["achieved"] / peak
        # dram__bytes_read.sum + dram__bytes_write.sum per step from the ncu --set full capture of this same command
        # (written by tools/summarize_ncu.py); attached only when it was taken from these very kernel sources
        prof = os.path.join(ROOT, "profiles", "r2_dram_bytes_per_step.json")
        if name == "cfg2" and os.path.exists(prof) and not env.options:
            try:
                t = json.load(open(prof))
                if t.get("source_sha") == source_sha():
                    roof_b["traffic"], roof_q["traffic"] = t["build_bytes_per_step"], t["query_bytes_per_step"]
                    roof_b["traffic_source"] = roof_q["traffic_source"] = (
                        "profiles/r2_dram_bytes_per_step.json: ncu --set full of this command at kernel sources %s; "
                        "a constant from that capture, not measured by this run" % t["source_sha"])
                else:
                    roof_q["traffic_source"] = "none: the committed ncu capture is of other kernel sources"
            except Exception:
                pass
        roof_s = {"bound": "hbm", "kernel": "whole step (build phase + merge + query phase)", "unit": "GB/s", "peak": peak,
                  "achieved": (ins_bytes + qry_bytes) / (ms_total / S * 1e-3) / 1e9, "launch_ms": ms_total / S,
                  "traffic": (roof_b["traffic"] + roof_q["traffic"]) if roof_b["traffic"] and roof_q["traffic"] else None}
        roof_s["frac"] = roof_s["achieved"] / peak
        probe_file = os.path.join(ROOT, "profiles", "r1_random_access_probe.jsonl")
        if os.path.exists(probe_file):
            try:
                roof_q["random_access_probe"] = [json.loads(ln) for ln in open(probe_file) if ln.strip()]
            except Exception:
                pass
        out.update({"roofline": roof_q, "roofline_build": roof_b, "roofline_step": roof_s, "clocks": clocks})
        if miss:
            out["query_miss_set"] = miss
        if name == "cfg2" and not env.args.no_job:
            out["job"] = run_job(env, cfg, filt, pm)
    if pm is not None:
        pm.close()
    del filt
    torch.cuda.synchronize()
    return out


def run_job(env, cfg, filt, pm):
    """The complete cfg2 job, strong-scaled: every chunk of the 3 Gbp genome (sharded round-robin over the ranks) ->
    merge -> 1e8 reads (sharded), between one pair of events; max over ranks."""
    torch, dist, ctx, world = env.torch, env.dist, env.ctx, env.world
    k = cfg["k"]
    n_mine = len([c for c in range(env.n_chunks) if c % world == env.rank])
    if n_mine > len(env.g):
        return {"skipped": "needs every chunk of the rank's shard resident (run without --job-chunks limits)"}
    n_batches = (JOB_READS + env.n_reads - 1) // env.n_reads
    q_mine = len([b for b in range(n_batches) if b % world == env.rank])
    d_stats = torch.zeros(4, dtype=torch.int64, device=env.dev)
    best = None
    for rep in range(2):
        filt.clear()
        d_stats.zero_()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        env.barrier()
        torch.cuda.synchronize()
        ev[0].record(env.stream)
        for i in range(n_mine):
            n = env.insert_len(i, k)
            filt.insertSeqsDevice(env.g[i].data_ptr(), n, env.goff(n).data_ptr(), 1, d_stats.data_ptr())
        if pm is not None and env.args.merge_chunks > 0 and hasattr(pm, "launch_range"):
            from btl_bloomfilter_b200 import parallel
            parallel.pipelined_flush_merge(filt, pm, env.args.merge_chunks, env.side, env.token2, ev[1])
        else:
            ctx.flush()
            ev[1].record(env.stream)
            if pm is not None:
                env.stream_barrier()
                pm.launch()
                env.stream_barrier()
        ev[2].record(env.stream)
        for i in range(q_mine):
            filt.containsSeqsDevice(env.r[i % len(env.r)].data_ptr(), env.read_bases, env.d_roff.data_ptr(), env.n_reads,
                                    env.d_hits.data_ptr(), 0, d_stats[2:].data_ptr())
        ev[3].record(env.stream)
        torch.cuda.synchronize()
        env.barrier()
        ms = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])]
        if best is None or ms[3] < best[3]:
            best = ms
    st = d_stats.cpu().numpy()
    assert int(st[2]) == int(st[3]), "job: a k-mer of the genome was not found after the full build"
    pop = filt.getPop()
    kk = torch.tensor([int(st[0]), int(st[2])], dtype=torch.int64, device=env.dev)
    tt = torch.tensor(best, dtype=torch.float64, device=env.dev)
    pp = torch.tensor([pop, -pop], dtype=torch.int64, device=env.dev)
    if world > 1:
        dist.all_reduce(kk, op=dist.ReduceOp.SUM)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pp, op=dist.ReduceOp.MIN)
    occ = pop / cfg["size"]
    return {"what": "3 Gbp genome build sharded over the ranks -> merge -> %d reads x %d bp sharded over the ranks; best of 2"
                    % (n_batches * env.n_reads, READ_LEN), "scaling": "strong",
            "build_ms": float(tt[0]), "merge_ms": float(tt[1]), "query_ms": float(tt[2]), "total_ms": float(tt[3]),
            "kmers_inserted": int(kk[0]), "kmers_queried": int(kk[1]),
            "gkmers_s": float(int(kk[0]) + int(kk[1])) / (float(tt[3]) * 1e-3) / 1e9,
            "build_gkmers_s_incl_merge": int(kk[0]) / ((float(tt[0]) + float(tt[1])) * 1e-3) / 1e9,
            "popcount": int(pp[0]), "identical_popcount_on_all_ranks": int(pp[0]) == -int(pp[1]),
            "occupancy": occ, "expected_occupancy": float(1.0 - np.exp(-cfg["h"] * (G_LEN - k + 1) / cfg["size"]))}


# ---------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs)")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5a,cfg5b", help="which configs the `configs` block covers")
    ap.add_argument("--config-steps", type=int, default=16,
                    help="steps of the other configs (16 chunks = 1 Gi k-mers: one full accumulation of the 16 GiB filter)")
    ap.add_argument("--no-job", action="store_true", help="skip the strong-scaled full cfg2 job")
    ap.add_argument("--cpu-sample", type=int, default=8 << 20, help="bases per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merge", default="auto", choices=["auto", "multimem", "peer", "hybrid30", "hybrid50", "hybrid70", "ipc"],
                    help="N > 1 merge kernel: auto = the faster of the in-switch OR (multimem) and the peer-memory kernel, timed "
                         "once on this box; ipc = own allocations + CUDA IPC + the peer-memory kernel")
    ap.add_argument("--merge-chunks", type=int, default=0,
                    help="N > 1 BloomFilter builds: pipeline the merge behind pass 2 of the build in this many chunks of "
                         "partitions (0: pass 2, then the merge)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the 2-bit packed host-buffer leg (e2e_packed)")
    ap.add_argument("--chunk", type=int, default=CHUNK)
    ap.add_argument("--query-factor", type=int, default=4, help="read bases per query batch, in units of --chunk")
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity (0: leave as is)")
    ap.add_argument("--stream-priority", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="context option key=value (tuning experiments)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = CONFIGS[args.config]
    W, S = max(0, args.warmup), max(1, args.steps)
    config = {"workload": cfg["workload"], "name": args.config, "k": cfg["k"], "hashes": cfg["h"],
              "filter_%s" % ("bits" if cfg["kind"] == "bloom" else "counters"): cfg["size"], "genome_bp": G_LEN,
              "read_len": READ_LEN, "chunk_windows": args.chunk, "query_bases_per_step": args.chunk * args.query_factor,
              "step": "K build batches, settle" + (", merge the per-GPU partial filters" if world > 1 else "") +
                      ", then K query batches, all inside the timed region",
              "l2": "inputs larger than L2 (64 MiB / 256 MiB per batch, multi-GB filter); no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(cfg, S, min(W, 1), args.cpu_sample, HOST_THREADS)
        if "unavailable" in r:
            print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}))
            return
        line = {"impl": "reference", "metric": "k-mers/s inserted+queried", "value": r["value"], "unit": "Gk-mer/s",
                "n_gpus": args.gpus, "steps": S, "warmup": min(W, 1), "ms_per_step": r["ms_per_step"],
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
    env = Env(args, torch, dist, B)
    if env.options:
        config["options"] = env.options
    want_job = args.config == "cfg2" and not args.no_job
    shard = len([c for c in range(env.n_chunks) if c % world == env.rank])
    n_genome = max(min(W + S, env.n_chunks), shard if want_job else 0)
    n_read = max(1, min(W + S, 8, n_genome))
    env.make_inputs(n_genome, n_read)

    line = run_config(env, args.config, cfg, S, W, headline=True)
    head = {"metric": "k-mers/s inserted+queried", "value": line.pop("value"), "unit": line.pop("unit"), "n_gpus": world,
            "steps": S, "warmup": W, "ms_per_step": line.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config}
    for key in ("workload", "k", "hashes", "filter_bytes", "steps", "warmup"):
        line.pop(key, None)
    head.update(line)
    if env.numa:
        head["config"]["host_numa_binding"] = env.numa
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_run(cfg, 2, 1, args.cpu_sample, HOST_THREADS)
        head["cpu_baseline"] = {key: cb[key] for key in ("value", "unit", "cores", "kind", "sample", "insert_gkmers_s",
                                                         "query_gkmers_s", "unavailable") if key in cb}

    # ---- the other BASELINE configs, same loop, fewer steps
    if args.config == "cfg2" and not args.no_configs:
        head["configs"] = {}
        for name in [n for n in args.configs.split(",") if n in CONFIGS and n != args.config]:
            c = CONFIGS[name]
            try:
                r = run_config(env, name, c, max(1, min(S, args.config_steps)), min(W, 3), headline=False)
            except Exception as e:  # noqa: BLE001  (a config that fails must not take the headline line with it)
                r = {"workload": c["workload"], "error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            if world == 1 and not args.no_cpu_baseline and "error" not in r:
                if c["kind"] == "counting":
                    # the exact oracle of the counting build is the single-threaded loop; N threads race (throughput only)
                    one = cpu_reference_run(c, 1, 0, 1 << 20, 1)
                    r["cpu_baseline_1thread"] = {key: one[key] for key in ("value", "unit", "cores", "kind", "sample",
                                                                          "insert_gkmers_s", "query_gkmers_s", "unavailable") if key in one}
                cb = cpu_reference_run(c, 1, 1, args.cpu_sample, HOST_THREADS)
                r["cpu_baseline"] = {key: cb[key] for key in ("value", "unit", "cores", "kind", "sample", "insert_gkmers_s",
                                                              "query_gkmers_s", "unavailable") if key in cb}
            head["configs"][name] = r
    if rank == 0:
        print(json.dumps(head))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
