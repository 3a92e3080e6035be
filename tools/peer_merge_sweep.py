"""torchrun worker: time the fused peer-memory merge (btlbf_merge_peers) for a few kernel shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import btl_bloomfilter_b200 as B
from btl_bloomfilter_b200 import parallel
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = B.Context(local)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
f = B.BloomFilter(31_568_113_856, 4, 25, ctx=ctx)
ptr, nbytes = f.device_ptr()
pm = parallel.PeerMerge(ctx, ptr, nbytes, f.KIND)
one_way = (world - 1) / world * nbytes
for mode in (0, 1, 2):
    for unroll, grid in ((1, 0), (4, 0), (4, 148 * 32)):
        ctx.set_option("peer_unroll", unroll); ctx.set_option("peer_grid", grid); ctx.set_option("peer_mode", mode)
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); dist.barrier()
            a.record(stream); pm.launch(); b.record(stream)
            torch.cuda.synchronize(); dist.barrier()
            best = min(best, a.elapsed_time(b))
        t = torch.tensor([best], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print("mode %d world %d unroll %d grid %5d: %.2f ms  %.0f GB/s per direction" % (mode, world, unroll, grid, float(t[0]), one_way / float(t[0]) / 1e6), flush=True)
pm.close()
dist.destroy_process_group()
