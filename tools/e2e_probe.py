"""Where does the streaming e2e path lose time?  Per-call GPU spans on the active stream."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import btl_bloomfilter_b200 as B
dev = torch.device("cuda", 0)
ctx = B.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
bits, H, K, chunk, RL = 31_568_113_856, 4, 25, 64 << 20, 150
f = B.BloomFilter(bits, H, K, ctx=ctx)
nreads = chunk // RL; rb = nreads * RL
NB = 4
dg = torch.empty(chunk + 64, dtype=torch.uint8, device=dev); dr = torch.empty(rb + 64, dtype=torch.uint8, device=dev)
hg, hr, hh = [], [], []
for j in range(NB):
    ctx.synth_genome_device(dg.data_ptr(), j * chunk, chunk, 42); ctx.synth_reads_device(dr.data_ptr(), 0, nreads, RL, j * chunk, chunk, 42, 7 + j)
    torch.cuda.synchronize()
    hg.append(dg[:chunk].cpu().pin_memory()); hr.append(dr[:rb].cpu().pin_memory()); hh.append(torch.zeros((rb + 31) // 32 * 4, dtype=torch.uint8).pin_memory())
roff = torch.arange(0, rb + 1, RL, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
goff = torch.tensor([0, chunk], dtype=torch.int64).pin_memory().numpy().view(np.uint64)
counts = torch.zeros((64, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
def run(mode, n=12):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n + 1)]
    ctx.sync(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ev[0].record(stream); th = []
    for i in range(n):
        j = i % NB
        if mode in ("both", "ins"):
            f.insertSeqsAsync((hg[j].numpy(), goff), counts[i, 0:2])
        ev[2 * i + 1].record(stream)
        if mode in ("both", "qry"):
            f.containsSeqsAsync((hr[j].numpy(), roff), hh[j].numpy(), counts[i, 2:4])
        ev[2 * i + 2].record(stream)
        th.append(time.perf_counter() - t0)
    ctx.sync(); dt = time.perf_counter() - t0
    ins = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n)]; qry = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(n)]
    print(mode, "total %.1f ms  per step %.2f ms" % (dt * 1e3, dt * 1e3 / n), " ins spans", " ".join("%.2f" % x for x in ins[2:8]), " qry spans", " ".join("%.2f" % x for x in qry[2:8]), " host enqueue done at %.1f ms" % (th[-1] * 1e3))
for m in ("both", "both", "ins", "qry"):
    run(m)
