"""Two drivers with one interface for the parity suite (tests/parity_suite.py):
  EmuBackend  tests/emu/libemu.so -- the kernels' per-thread phase functions run on the CPU (no GPU needed)
  GpuBackend  the product: btl_bloomfilter_b200 over the C ABI of libbtlbf_cuda.so on cuda:0
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libemu.so")
CSRC = os.path.join(ROOT, "btl_bloomfilter_b200", "csrc")


def _batch(seqs):
    if isinstance(seqs, tuple):
        return np.ascontiguousarray(seqs[0], np.uint8), np.ascontiguousarray(seqs[1], np.uint64)
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    return np.frombuffer(b"".join(bs), dtype=np.uint8).copy(), off


def bit_bytes(n):
    return (int(n) + 31) // 32 * 4


def build_emu():
    srcs = [os.path.join(EMU_DIR, "emu.cpp")] + [os.path.join(CSRC, f) for f in
                                                  ("tile_core.cuh", "kernels.cuh", "nthash_dev.cuh", "host_params.hpp")]
    if os.path.exists(EMU_SO) and all(os.path.getmtime(EMU_SO) >= os.path.getmtime(s) for s in srcs):
        return
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I" + cuda_inc, "-I" + CSRC,
                           "-o", EMU_SO, srcs[0]])


class _EmuFilter:
    def __init__(self, be, kind, size, h, k, thr, seeds, h2):
        self.be, self.kind, self.size, self.h, self.k, self.thr = be, kind, size, h, k, thr
        self.seeds, self.h2 = seeds, h2
        self.nbytes = size // 8 if kind == 0 else size
        self.data = np.zeros((self.nbytes + 15) // 16 * 16, np.uint8)
        self.deferred = self.rounds = 0

    def _run(self, op, seqs, want_hit=False, want_valid=False, want_counts=False):
        bases, off = _batch(seqs)
        n = bases.size
        hit = np.zeros(bit_bytes(n) // 4 + 1, np.uint32) if want_hit else None
        valid = np.zeros(bit_bytes(n) // 4 + 1, np.uint32) if want_valid else None
        counts = np.zeros(n + 1, np.uint8) if want_counts else None
        stats = np.zeros(2, np.uint64)
        info = np.zeros(3, np.uint64)
        self.be._call(op, self.kind, self.size, self.h, self.k, self.thr, self.seeds, self.h2, self.data, bases, off,
                      hit, valid, counts, None, None, stats, info, packed=self.be.packed and op in (0, 1))
        self.deferred += int(info[0])
        self.rounds += int(info[1])
        self.be.bin_overflow = int(info[2])
        nb = bit_bytes(n)
        return (n, int(stats[0]), int(stats[1]),
                None if hit is None else hit.view(np.uint8)[:nb].copy(),
                None if valid is None else valid.view(np.uint8)[:nb].copy(),
                None if counts is None else counts[:n].copy())

    def insert(self, seqs):
        return self._run(0, seqs)[1]

    def contains(self, seqs):
        _, nk, nh, hit, valid, _ = self._run(1, seqs, True, True)
        return nk, nh, hit, valid

    def insert_and_check(self, seqs):
        _, nk, nh, hit, valid, _ = self._run(2, seqs, True, True)
        return nk, hit, valid

    def mincount(self, seqs):
        _, nk, nh, _, valid, counts = self._run(3, seqs, False, True, True)
        return nk, counts, valid

    def increment_all(self, seqs):
        return self._run(4, seqs)[1]

    def bytes(self):
        return self.data[: self.nbytes].copy()

    def set_bytes(self, arr):
        self.data[: self.nbytes] = arr
        self.data[self.nbytes:] = 0

    def ordered_stats(self):
        return self.deferred, self.rounds


class EmuBackend:
    name = "emu"

    def __init__(self, chunk=1 << 20, batch=1 << 20, resv_log2=16, list_log2=10, force_generic=0, query_mode=0,
                 bin_shift=0, bin_writers=3, bin_slack_pct=20, packed=False):
        build_emu()
        self.L = C.CDLL(EMU_SO)
        self.packed = packed  # insert / contains read the batch as 2-bit codes + invalid plane (btlbf_pack_seqs)
        self.bin_overflow = 0
        self.opts = dict(chunk=chunk, batch=batch, resv_log2=resv_log2, list_log2=list_log2,
                         force_generic=force_generic, query_mode=query_mode, bin_shift=bin_shift,
                         bin_writers=bin_writers, bin_slack_pct=bin_slack_pct)

    def _call(self, op, kind, size, h, k, thr, seeds, h2, filt, bases, off, hit, valid, counts, hashes, strands, stats,
              info, packed=False):
        def p(a):
            return None if a is None else C.c_void_p(a.ctypes.data)
        invalid = None
        if packed and bases.size:
            from btl_bloomfilter_b200 import pack_seqs  # the product's host packer (host code only)
            pk = pack_seqs((bases, off))
            bases, invalid = pk.codes, pk.invalid
        sp = (C.c_char_p * len(seeds))(*[s.encode() for s in seeds]) if seeds else None
        msg = C.create_string_buffer(256)
        o = self.opts
        rc = self.L.emu_seq_op(C.c_int(op), C.c_int(kind), C.c_uint64(size), C.c_uint(h), C.c_uint(k), C.c_uint(thr),
                               sp, C.c_uint(len(seeds) if seeds else 0), C.c_uint(h2 if seeds else 0), p(filt),
                               p(bases) if bases.size else C.c_void_p(0), p(off), C.c_uint64(off.size - 1), p(hit),
                               p(valid), p(counts), p(hashes), p(strands), p(stats), C.c_int(o["force_generic"]),
                               C.c_int(o["query_mode"]), C.c_uint64(o["chunk"]), C.c_uint64(o["batch"]),
                               C.c_uint(o["resv_log2"]), C.c_uint(o["list_log2"]), p(info), msg, C.c_size_t(256),
                               C.c_uint(o["bin_shift"]), C.c_uint(o["bin_writers"]), C.c_uint(o["bin_slack_pct"]),
                               p(invalid), C.c_int(1 if packed and bases.size else 0))
        if rc != 0:
            raise ValueError(msg.value.decode())

    def hash(self, seqs, h, k, seeds=None, h2=1):
        bases, off = _batch(seqs)
        n = bases.size
        H = len(seeds) * h2 if seeds else h
        hashes = np.zeros((n + 1, H), np.uint64)
        strands = np.zeros((n + 1, H), np.uint8)
        valid = np.zeros(bit_bytes(n) // 4 + 1, np.uint32)
        stats = np.zeros(2, np.uint64)
        self._call(5, 0, 8, h, k, 0, seeds, h2, None, bases, off, None, valid, None, hashes, strands, stats, None)
        return int(stats[0]), hashes[:n], strands[:n], valid.view(np.uint8)[: bit_bytes(n)].copy()

    def filter(self, kind, size, h, k, thr=1, seeds=None, h2=1):
        if kind == 0 and size % 8:
            raise ValueError("not a multiple of 8")
        return _EmuFilter(self, kind, size, h, k, thr, seeds, h2)


class _GpuFilter:
    def __init__(self, f, packed=False):
        self.f = f
        self.packed = packed

    def insert(self, seqs):
        if self.packed:
            return self.f.insertSeqsPacked(self.f._pack(_batch(seqs)))
        return self.f.insertSeqs(_batch(seqs))

    def contains(self, seqs):
        if self.packed:
            r = self.f.containsSeqsPacked(self.f._pack(_batch(seqs)))
        else:
            r = self.f.containsSeqs(_batch(seqs))
        return r.n_kmers, r.n_hits, r.hit_bits, r.valid_bits

    def insert_and_check(self, seqs):
        import ctypes as C
        from btl_bloomfilter_b200 import filters as F
        bases, off = _batch(seqs)
        n = bases.size
        found = np.zeros(bit_bytes(n), np.uint8)
        valid = np.zeros(bit_bytes(n), np.uint8)
        nk = C.c_uint64()
        F.check(self.f._L.btlbf_insert_and_check_seqs(self.f._h, F._ptr(bases), F._p64(off), off.size - 1,
                                                      F._ptr(found), F._ptr(valid), C.byref(nk)))
        return nk.value, found, valid

    def mincount(self, seqs):
        r = self.f.minCountSeqs(_batch(seqs))
        return r.n_kmers, r.counts, r.valid_bits

    def increment_all(self, seqs):
        return self.f.incrementAllSeqs(_batch(seqs))

    def bytes(self):
        return self.f.to_numpy()

    def set_bytes(self, arr):
        self.f.from_numpy(arr)

    def ordered_stats(self):
        return self.f.orderedStats()


class GpuBackend:
    name = "gpu"

    def __init__(self, packed=False, **opts):
        import btl_bloomfilter_b200 as B
        self.B = B
        self.packed = packed  # insert / contains go through btlbf_pack_seqs + the *_seqs_packed entry points
        self.ctx = B.Context(0)
        # translate the emulator's option names
        m = {"chunk": "chunk_bases", "batch": "cbf_batch", "bin_shift": "bin_part_log2"}
        if "bin_shift" in opts:
            opts = dict(opts, bin_mode=1, bin_query_mode=1)
        opts.pop("bin_writers", None)
        for k, v in opts.items():
            self.ctx.set_option(m.get(k, k), v)

    def hash(self, seqs, h, k, seeds=None, h2=1):
        return self.ctx.hash_seqs(_batch(seqs), h, k, seeds, h2)

    def filter(self, kind, size, h, k, thr=1, seeds=None, h2=1):
        B = self.B
        if kind == 0:
            f = B.BloomFilter(size, h, k, ctx=self.ctx)
        else:
            f = B.CountingBloomFilter(size, h, k, thr, ctx=self.ctx)
        if seeds:
            f.setSeeds(seeds, h2)
        f._pack = B.pack_seqs
        return _GpuFilter(f, self.packed)
