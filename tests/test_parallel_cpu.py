"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo runs of the partial-filter merge
(slice exchange -> local reduce -> all-gather) and of the unit sharding.  The local reduce step on GPUs is the
CUDA OR / saturating-add kernel; here a torch stand-in is injected (test infrastructure only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from btl_bloomfilter_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, nbytes, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pad = parallel.padded_bytes(nbytes, world)
        g = torch.Generator().manual_seed(100 + rank)
        if kind == 0:
            part = torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g)
            part &= torch.randint(0, 256, (pad,), dtype=torch.uint8, generator=g)

            def reduce_fn(dst, src):
                dst |= src
        else:
            part = torch.randint(0, 200, (pad,), dtype=torch.uint8, generator=g)

            def reduce_fn(dst, src):
                dst.copy_(torch.clamp(dst.to(torch.int16) + src.to(torch.int16), max=255).to(torch.uint8))
        part[nbytes:] = 0
        mine = part.clone()
        parallel.merge_partials(part, reduce_fn)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        if kind == 0:
            exp = gathered[0].clone()
            for t in gathered[1:]:
                exp |= t
        else:
            exp = torch.clamp(sum(t.to(torch.int32) for t in gathered), max=255).to(torch.uint8)
        ok = bool(torch.equal(part, exp))
        lo, hi = parallel.shard_range(1001, rank, world)
        out.put((rank, ok, lo, hi))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,nbytes", [(0, 125), (0, 4096 + 8), (1, 100008), (1, 64)])
def test_merge_partials_gloo_world2(kind, nbytes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, nbytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    # shards are contiguous, disjoint and cover all units
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == 1001


def test_slice_geometry():
    for nbytes in (1, 15, 16, 17, 3946014232, 1 << 32):
        for world in (1, 2, 4, 8):
            L = parallel.slice_len(nbytes, world)
            assert L % 16 == 0 and L * world >= nbytes and parallel.padded_bytes(nbytes, world) == L * world
            assert L * world - nbytes < 16 * world + world
    cover = []
    for r in range(8):
        cover.append(parallel.shard_range(45, r, 8))
    assert cover[0][0] == 0 and cover[-1][1] == 45
    assert all(cover[i][1] == cover[i + 1][0] for i in range(7))
    assert max(b - a for a, b in cover) - min(b - a for a, b in cover) <= 1


def test_fused_merge_slices_partition_the_filter():
    """btlbf_merge_slice / parallel.merge_slice: the ranks' byte ranges are 16-byte granules, disjoint, and
    cover the (16-byte padded) filter exactly; the C ABI and the Python mirror agree."""
    import ctypes as C
    from btl_bloomfilter_b200 import lib, parallel
    L = lib()
    for nbytes in (0, 1, 16, 17, 300_007 * 8, 3_946_014_232, 100_008):
        for world in (1, 2, 3, 4, 8, 16):
            end = 0
            for rank in range(world):
                lo, hi = C.c_uint64(), C.c_uint64()
                assert L.btlbf_merge_slice(nbytes, world, rank, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == parallel.merge_slice(nbytes, rank, world)
                assert lo.value % 16 == 0 and hi.value % 16 == 0 and lo.value <= hi.value
                assert lo.value == end
                end = hi.value
            assert end == (nbytes + 15) // 16 * 16
