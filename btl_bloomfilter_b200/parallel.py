"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards by sequence: queries run against a filter replicated on every GPU (no data-path
collective), builds produce one partial filter per GPU.  The partial filters are merged into the
reference layout by   all-to-all of 1/N array slices  ->  local OR (BloomFilter) or saturating add
(CountingBloomFilter<uint8_t>) kernel  ->  all-gather.   OR is commutative, associative and idempotent,
so the merged BloomFilter is bit-identical to a single-GPU build of the same sequences for any sharding.
The counting merge (saturating add of per-shard incrementMin builds) is a different function from the
sequential single-filter build; its oracle is "N sequential partial builds, then saturating add".

Two implementations of that merge:
  merge_partials / merge_filter   NCCL all-to-all + local reduce kernel + NCCL all-gather (also the gloo
                                  path of the CPU tests)
  PeerMerge / fused_merge_filter  ONE kernel per GPU over peer-mapped memory (CUDA IPC + NVLink loads and
                                  stores): each rank reduces its 1/N byte range of all N partial filters
                                  in registers and writes the result into all N of them -- no staging
                                  buffer, no separate reduce pass (btlbf_merge_peers)
  MultimemMerge                   the OR merge INSIDE NVSwitch: the partial filters live in symmetric memory
                                  bound to one multicast object; multimem.ld_reduce.or pulls the OR of all N
                                  replicas of this rank's byte range through the switch and multimem.st fans
                                  the result out to all N (btlbf_merge_multimem).  BloomFilter only.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

ALIGN = 16


def shard_range(n_units, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_units independent units (reads, chunks) for this rank."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_len(nbytes, world):
    """Length of the equal 1/N slices (16-byte aligned) a filter of nbytes is exchanged in."""
    per = (int(nbytes) + world - 1) // world
    return (per + ALIGN - 1) // ALIGN * ALIGN


def padded_bytes(nbytes, world):
    return slice_len(nbytes, world) * world


def device_tensor_from_ptr(ptr, nbytes, device):
    """uint8 tensor view of raw device memory (no copy)."""
    class _Mem:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                    "version": 2}
    return torch.as_tensor(_Mem(), device=device)


def merge_partials(t, reduce_fn, group=None):
    """In-place merge of the per-rank partial arrays `t` (uint8, length world*slice, identical layout on
    every rank): afterwards every rank holds reduce(all partials).  reduce_fn(dst_view, src_view) folds
    src into dst (the CUDA OR / saturating-add kernel on GPUs)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return t
    n = t.numel()
    assert n % world == 0 and (n // world) % ALIGN == 0, "pad the array with padded_bytes()"
    L = n // world
    mine = t[rank * L:(rank + 1) * L]
    recv = torch.empty(n, dtype=t.dtype, device=t.device)
    backend = dist.get_backend(group)
    if backend == "nccl":
        # slice j of my partial goes to rank j; slice `rank` of every peer's partial comes to me
        dist.all_to_all_single(recv, t, group=group)
    else:
        ops = []
        for peer in range(world):
            if peer == rank:
                continue
            ops.append(dist.P2POp(dist.isend, t[peer * L:(peer + 1) * L], peer, group))
            ops.append(dist.P2POp(dist.irecv, recv[peer * L:(peer + 1) * L], peer, group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for peer in range(world):
        if peer != rank:
            reduce_fn(mine, recv[peer * L:(peer + 1) * L])
    del recv
    if backend == "nccl":
        dist.all_gather_into_tensor(t, mine.clone(), group=group)
    else:
        parts = [torch.empty(L, dtype=t.dtype, device=t.device) for _ in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)
        for peer in range(world):
            t[peer * L:(peer + 1) * L].copy_(parts[peer])
    return t


def cuda_reduce_fn(ctx, kind):
    """The local step on the GPU: btlbf_merge_device_buffers (OR for BLOOM, saturating add for COUNTING8)."""
    L = ctx.L
    from ._capi import check

    def fn(dst, src):
        check(L.btlbf_merge_device_buffers(ctx.handle, kind, C.c_void_p(dst.data_ptr()), C.c_void_p(src.data_ptr()),
                                           dst.numel()))
    return fn


def merge_filter(filt, group=None):
    """Merge the per-rank partial filters behind `filt` (created with from_device_memory over a tensor of
    padded_bytes()) in place; returns the tensor."""
    t = filt._tensor
    # kernels of the context and NCCL must be ordered: both run on torch's current stream
    filt._ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    return merge_partials(t, cuda_reduce_fn(filt._ctx, filt.KIND), group)


def bench_merge(filt, ctx, dev, repeats=3):
    """Times the build-side merge of bench.py's per-rank partial filters (max over ranks, CUDA events)."""
    world = dist.get_world_size()
    ptr, nbytes = filt.device_ptr()
    pad = padded_bytes(nbytes, world)
    # the library's allocation is only padded to 16 bytes: merge through a padded staging tensor
    stage = torch.zeros(pad, dtype=torch.uint8, device=dev)
    view = device_tensor_from_ptr(ptr, nbytes, dev)
    stage[:nbytes].copy_(view)
    fn = cuda_reduce_fn(ctx, filt.KIND)
    times = []
    for _ in range(repeats):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        a.record()
        merge_partials(stage, fn)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    tt = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    view.copy_(stage[:nbytes])
    pop = torch.tensor([filt.getPop()], dtype=torch.int64, device=dev)
    pops = [torch.zeros_like(pop) for _ in range(world)]
    dist.all_gather(pops, pop)
    ms = float(tt[0])
    traffic = 2.0 * (world - 1) / world * pad  # bytes sent (= received) per GPU: all-to-all + all-gather
    out = {"nccl_ms": ms, "filter_bytes": int(nbytes), "link_bytes_per_gpu_per_direction": int(traffic),
           "nccl_link_GBps_per_gpu_per_direction": traffic / (ms * 1e-3) / 1e9,
           "identical_popcount_on_all_ranks": len({int(p) for p in pops}) == 1, "popcount": int(pops[0])}
    # the fused kernel over peer memory, on the (already merged, hence idempotent under OR) filters
    if filt.KIND == 0:
        pm = PeerMerge(ctx, ptr, nbytes, filt.KIND)
        times = []
        for _ in range(repeats):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier()
            a.record()
            pm.launch()
            b.record()
            torch.cuda.synchronize()
            dist.barrier()
            times.append(a.elapsed_time(b))
        pm.close()
        tf = torch.tensor([min(times)], dtype=torch.float64, device=dev)
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        pop2 = torch.tensor([filt.getPop()], dtype=torch.int64, device=dev)
        pops2 = [torch.zeros_like(pop2) for _ in range(world)]
        dist.all_gather(pops2, pop2)
        out["fused_ms"] = float(tf[0])
        # every NVLink direction of a GPU carries the payload of its own peer loads / stores plus that of the
        # peers' stores / loads aimed at it: 2 (N-1)/N of the filter per direction, like the NCCL path, but in
        # one phase
        out["fused_link_GBps_per_gpu_per_direction"] = traffic / (float(tf[0]) * 1e-3) / 1e9
        out["fused_popcount_unchanged"] = all(int(p) == int(pops[0]) for p in pops2)
    out["ms"] = min(ms, out.get("fused_ms", ms))
    return out


def merge_slice(nbytes, rank, world):
    """Byte range [lo, hi) of an nbytes filter that `rank` reduces in the fused merge (16-byte granules;
    the same arithmetic as btlbf_merge_slice)."""
    nvec = (int(nbytes) + 15) // 16
    per = (nvec + world - 1) // world
    lo, hi = min(rank * per, nvec), min((rank + 1) * per, nvec)
    return lo * 16, hi * 16


class PeerMerge:
    """Peer mappings of one device array per rank (CUDA IPC), reusable for any number of fused merges.

    ptr / nbytes: this rank's partial filter (raw device memory of identical size on every rank: the
    library's own allocation, f.device_ptr(), or a torch tensor's storage).  Collective: every rank of
    `group` must construct it, merge() and close() together."""

    def __init__(self, ctx, ptr, nbytes, kind, group=None):
        from ._capi import check
        self.ctx, self.kind, self.nbytes, self.group = ctx, kind, int(nbytes), group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        L = ctx.L
        handle = (C.c_uint8 * 64)()
        off = C.c_uint64()
        check(L.btlbf_ipc_export(ctx.handle, C.c_void_p(ptr), handle, C.byref(off)))
        mine = (bytes(handle), int(off.value), int(nbytes), os.getpid())
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        assert all(e[2] == self.nbytes for e in everyone), "partial filters differ in size across ranks"
        self._mapped = []
        bases = (C.c_void_p * self.world)()
        for p, (h, o, _, pid) in enumerate(everyone):
            if p == self.rank:
                bases[p] = ptr
                continue
            hb = (C.c_uint8 * 64).from_buffer_copy(h)
            base = C.c_void_p()
            check(L.btlbf_ipc_open(ctx.handle, hb, C.byref(base)))
            self._mapped.append(base.value)
            bases[p] = base.value + o
        self._bases = bases

    def merge(self):
        """Every rank's array := reduce(all arrays).  Ranks are synchronised before and after."""
        from ._capi import check
        self.ctx.flush()  # k-mers parked by the partitioned build reach the array before any peer reads it
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # all partial builds are complete and visible
        check(self.ctx.L.btlbf_merge_peers(self.ctx.handle, self.kind, self._bases, self.world, self.rank,
                                           self.nbytes))
        self.ctx.sync()
        dist.barrier(group=self.group)  # all ranges have been written everywhere

    def launch(self):
        """The kernel alone (for timing): the caller brackets it with ctx.flush() + synchronize + barrier (the
        flush BEFORE the barrier: parked k-mers must have reached this rank's array when the peers start)."""
        from ._capi import check
        check(self.ctx.L.btlbf_merge_peers(self.ctx.handle, self.kind, self._bases, self.world, self.rank,
                                           self.nbytes))

    def close(self):
        from ._capi import check
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for b in self._mapped:
            check(self.ctx.L.btlbf_ipc_close(self.ctx.handle, C.c_void_p(b)))
        self._mapped = []


def fused_merge_filter(filt, group=None):
    """One-shot fused merge of the per-rank partial filters behind `filt` (any filter of the library)."""
    ptr, nbytes = filt.device_ptr()
    pm = PeerMerge(filt._ctx, ptr, nbytes, filt.KIND, group)
    try:
        pm.merge()
    finally:
        pm.close()


# ---------------------------------------------------------------- in-switch (NVLS multicast) merge
def symmetric_filter(cls, size, hashNum, kmerSize, ctx, group=None, threshold=0):
    """A filter of class `cls` whose array lives in symmetric memory (torch.distributed._symmetric_memory: one
    allocation per rank, mapped into every peer and -- where NVSwitch multicast is available -- bound to one
    multicast object).  Collective.  Returns (filter, rendezvous handle); handle.multicast_ptr is 0 without NVLS.
    The array is zeroed.  Context option wrap_accumulate=1 keeps the partitioned build's accumulation for it."""
    import torch.distributed._symmetric_memory as symm_mem
    nbytes = size // 8 if cls.KIND == 0 else size
    pad = padded_bytes((nbytes + 15) // 16 * 16, dist.get_world_size(group))
    dev = torch.device("cuda", ctx.device)
    t = symm_mem.empty(pad, dtype=torch.uint8, device=dev)
    t.zero_()
    hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
    f = cls.from_device_memory(t, size, hashNum, kmerSize, threshold=threshold, ctx=ctx)
    return f, hdl


class MultimemMerge:
    """Merge of per-rank partial filters held in symmetric memory (symmetric_filter()).  Two kernels serve it:
      "multimem"  BloomFilter only, where the handle carries an NVLS multicast mapping: the OR happens inside NVSwitch
                  (btlbf_merge_multimem: multimem.ld_reduce.or + multimem.st); a GPU moves ~(1 + 1/N) of the filter per
                  NVLink direction
      "peer"      the peer-memory kernel (btlbf_merge_peers) over the handle's peer pointers (no CUDA IPC plumbing
                  needed); 2 (N-1)/N of the filter per direction; also the saturating-add merge of counting filters
      "hybridP"   P per cent of the byte range inside the switch and the rest over peer memory, in one kernel
                  (btlbf_merge_hybrid)
    Which one is faster depends on N (measured on B200 at N = 2: peer 6.5 ms, multimem 12.2 ms for a 3.95 GB filter,
    because multimem pulls the local replica through the switch too); calibrate() times both once on the box and keeps
    the faster.  Collective: every rank of the group constructs it and calls its methods together."""

    def __init__(self, ctx, hdl, nbytes, kind=0, group=None, mode=None):
        self.ctx, self.hdl, self.nbytes, self.kind, self.group = ctx, hdl, int(nbytes), int(kind), group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if kind == 0 else 0
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        assert len(ptrs) == self.world
        self._bases = (C.c_void_p * self.world)(*ptrs)
        self.modes = (["multimem"] if self.mc else []) + ["peer"]  # what calibrate() times
        # both at once, P per cent of the range inside the switch: on request only (measured at N = 2: 8.8 / 9.5 / 10.7 ms
        # for P = 30 / 50 / 70 between peer 6.5 and multimem 12.1 -- the two paths share the links, their times add)
        hybrid = ["hybrid%d" % p for p in (30, 50, 70)] if self.mc and self.world in (2, 4, 8) else []
        if mode is not None and mode not in self.modes + hybrid:
            raise RuntimeError("merge mode %r is not available here (have %s)" % (mode, self.modes + hybrid))
        self.mode = mode or self.modes[0]
        self.calibration = None

    @staticmethod
    def available(hdl):
        return int(getattr(hdl, "multicast_ptr", 0) or 0) != 0

    def launch(self, mode=None):
        """The kernel alone; the caller brackets it (ctx.flush() before the first barrier, see PeerMerge.launch)."""
        from ._capi import check
        L, ctx = self.ctx.L, self.ctx
        mode = mode or self.mode
        if mode == "multimem":
            check(L.btlbf_merge_multimem(ctx.handle, 0, C.c_void_p(self.mc), self.world, self.rank, self.nbytes))
        elif mode.startswith("hybrid"):
            check(L.btlbf_merge_hybrid(ctx.handle, 0, C.c_void_p(self.mc), self._bases, self.world, self.rank, self.nbytes,
                                       int(mode[6:])))
        else:
            check(L.btlbf_merge_peers(ctx.handle, self.kind, self._bases, self.world, self.rank, self.nbytes))

    def launch_range(self, lo, hi, cuda_stream=0):
        """The peer-memory kernel for bytes [lo, hi) of the arrays only, on `cuda_stream` (0: the context's active
        stream), without applying deferred work (btlbf_merge_peers_range)."""
        from ._capi import check
        check(self.ctx.L.btlbf_merge_peers_range(self.ctx.handle, self.kind, self._bases, self.world, self.rank, int(lo), int(hi),
                                                 C.c_void_p(cuda_stream or 0)))

    def calibrate(self, repeats=2):
        """Times every available kernel on this box (max over ranks) and keeps the fastest.  BloomFilter only: OR is
        idempotent, so the trial merges leave a filter that is still correct -- merged -- for whatever was inserted
        before; call it on an empty or throw-away filter state when the unmerged partials matter."""
        if self.kind != 0 or len(self.modes) < 2:
            return self.mode
        dev = torch.device("cuda", self.ctx.device)
        res = {}
        self.ctx.flush()
        for mode in self.modes:
            best = None
            for _ in range(repeats + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
                a.record(torch.cuda.current_stream())
                self.launch(mode)
                b.record(torch.cuda.current_stream())
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
                ms = a.elapsed_time(b)
                best = ms if best is None else min(best, ms)
            t = torch.tensor([best], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            res[mode] = float(t[0])
        self.calibration = res
        self.mode = min(res, key=res.get)
        return self.mode

    def merge(self):
        self.ctx.flush()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.launch()
        self.ctx.sync()
        dist.barrier(group=self.group)

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)


def pipelined_flush_merge(filt, merger, n_chunks, side, token, applied_event=None):
    """Pass 2 of the parked build and the multi-GPU merge, pipelined: the build's pass 2 is partition-major, so chunk j
    of the partitions is merged over NVLink (on the stream `side`) while chunk j+1 is still being applied (on the
    current stream).  Per chunk: btlbf_filter_flush_parts -> event -> cross-rank barrier in stream order (a one-word
    all-reduce of `token`, a tensor used for nothing else) -> btlbf_merge_peers_range of exactly that byte range.
    On return the current stream is ordered after the last merge of every rank; applied_event (optional) is recorded on
    the current stream behind the last chunk of pass 2.  The context's kernels must run on the current torch stream
    (Context.set_stream(torch.cuda.current_stream().cuda_stream)).  Collective.  BloomFilter or counting filter held in symmetric
    memory (merger = MultimemMerge); the result is the one of merger.merge()."""
    main = torch.cuda.current_stream()
    for j in range(n_chunks):
        lo, hi = filt.flushParts(j, n_chunks)
        if hi <= lo:
            continue
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            dist.all_reduce(token, group=merger.group)  # every rank has applied this chunk
            merger.launch_range(lo, hi, side.cuda_stream)
    if applied_event is not None:
        applied_event.record(main)
    with torch.cuda.stream(side):
        dist.all_reduce(token, group=merger.group)      # every rank has written its share of every chunk everywhere
    main.wait_stream(side)
