"""Where does the streaming e2e query lose time?  256 MiB read batches through btlbf_contains_seqs_async."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import btl_bloomfilter_b200 as B
dev = torch.device("cuda", 0)
ctx = B.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
bits, H, K, chunk, RL = 31_568_113_856, 4, 25, 256 << 20, 150
f = B.BloomFilter(bits, H, K, ctx=ctx)
nreads = chunk // RL; rb = nreads * RL
NB = 3
dr = torch.empty(rb + 64, dtype=torch.uint8, device=dev)
hr, hh = [], []
for j in range(NB):
    ctx.synth_reads_device(dr.data_ptr(), 0, nreads, RL, j * chunk, chunk, 42, 7 + j)
    torch.cuda.synchronize()
    hr.append(dr[:rb].cpu().pin_memory()); hh.append(torch.zeros((rb + 31) // 32 * 4, dtype=torch.uint8).pin_memory())
roff = torch.arange(0, rb + 1, RL, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
counts = torch.zeros((64, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
doff = torch.arange(0, rb + 1, RL, dtype=torch.int64, device=dev)
dh = torch.zeros((rb + 31) // 32 + 8, dtype=torch.int32, device=dev)
st = torch.zeros(4, dtype=torch.int64, device=dev)
def run(mode, n=8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ctx.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ev[0].record(stream); th = []
    for i in range(n):
        j = i % NB
        if mode == "dev":
            f.containsSeqsDevice(dr.data_ptr(), rb, doff.data_ptr(), nreads, dh.data_ptr(), 0, st[2:].data_ptr())
        elif mode == "nohits":
            check = f._L.btlbf_contains_seqs_async(f._h, hr[j].numpy().ctypes.data, roff.ctypes.data_as(B._capi.u64p), nreads, None, None, counts[i, 2:4].ctypes.data)
            assert check == 0
        else:
            f.containsSeqsAsync((hr[j].numpy(), roff), hh[j].numpy(), counts[i, 2:4])
        ev[i + 1].record(stream)
        th.append(time.perf_counter() - t0)
    ctx.sync(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    spans = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print("%-8s total %.1f ms  per call %.2f ms (%.1f Gk-mer/s)" % (mode, dt * 1e3, dt * 1e3 / n, nreads * 126 / (dt / n) / 1e9), " stream spans", " ".join("%.2f" % x for x in spans), " host enqueue times", " ".join("%.1f" % (x * 1e3) for x in th))
for kv in sys.argv[1:]:
    k_, v_ = kv.split("="); ctx.set_option(k_, int(v_))
for m in ("dev", "async", "async", "nohits"):
    run(m)
