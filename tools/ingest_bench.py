"""FASTA / FASTQ ingest throughput (btlbf_insert_file / btlbf_query_file): synthetic files in /dev/shm."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import btl_bloomfilter_b200 as B
dev = torch.device("cuda", 0)
ctx = B.Context(0)
G = int(os.environ.get("GENOME_BP", str(1 << 30)))
tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
fa, fq = os.path.join(tmp, "btlbf_g.fa"), os.path.join(tmp, "btlbf_r.fq")
d = torch.empty(G + 64, dtype=torch.uint8, device=dev)
ctx.synth_genome_device(d.data_ptr(), 0, G, 42)
g = d[:G].cpu().numpy()
# FASTA: 8 records, 80-column lines
with open(fa, "wb") as fh:
    per = G // 8 // 80 * 80
    for i in range(8):
        fh.write(b">chr%d synthetic\n" % i)
        m = np.empty((per // 80, 81), np.uint8)
        m[:, :80] = g[i * per:(i + 1) * per].reshape(-1, 80)
        m[:, 80] = 10
        m.tofile(fh)
# FASTQ: 150 bp reads cut from the genome
n_reads = G // 4 // 150
m = np.empty((n_reads, 3 + 150 + 3 + 150 + 1), np.uint8)
m[:, 0:3] = np.frombuffer(b"@r\n", np.uint8)
m[:, 3:153] = g[:n_reads * 150].reshape(-1, 150)
m[:, 153:156] = np.frombuffer(b"\n+\n", np.uint8)
m[:, 156:306] = ord("I")
m[:, 306] = 10
m.tofile(fq)
del m
bits, h, k = 31_568_113_856, 4, 25
for threads in (32, 32, 1, 2, 4, 8, 16, 32):  # the first call creates the pinned staging buffers
    f = B.BloomFilter(bits, h, k, ctx=ctx)
    t0 = time.perf_counter(); ns, nk = f.insertFile(fa, threads); ctx.sync(); t1 = time.perf_counter()
    qs, qk, qh = f.queryFile(fq, threads); t2 = time.perf_counter()
    print(json.dumps({"threads": threads, "fasta_bytes": os.path.getsize(fa), "insert_records": ns, "insert_gkmers_s": nk / (t1 - t0) / 1e9,
                      "insert_GBps_of_file": os.path.getsize(fa) / (t1 - t0) / 1e9, "fastq_bytes": os.path.getsize(fq),
                      "query_records": qs, "query_gkmers_s": qk / (t2 - t1) / 1e9, "query_GBps_of_file": os.path.getsize(fq) / (t2 - t1) / 1e9,
                      "all_found": qh == qk}), flush=True)
    del f
os.remove(fa); os.remove(fq)
