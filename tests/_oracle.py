"""ctypes bindings for the TEST ORACLE (oracle/liboracle.so) and, when present, the compiled
reference (oracle/_ref/libbtlref.so).  Test infrastructure only -- the product package never
imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libbtlref.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)
u64 = C.c_uint64
u32 = C.c_uint


def build_oracle():
    src = os.path.join(ORACLE_DIR, "btl_oracle.c")
    stale = (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src)
    if stale or (os.path.isdir("/root/reference") and not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)


def _p8(a):
    return None if a is None else a.ctypes.data_as(u8p)


def _p64(a):
    return None if a is None else a.ctypes.data_as(u64p)


def as_batch(seqs):
    """list of bytes/str -> (bases uint8[n], offsets uint64[n_seqs+1])"""
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    if bases.size == 0:
        bases = np.zeros(1, np.uint8)[:0]
    return bases, off


def nbits_bytes(n):
    return (int(n) + 31) // 32 * 4


def bits_to_bool(bits, n):
    return np.unpackbits(bits, bitorder="little")[: int(n)].astype(bool)


class _SeedSet(C.Structure):
    _fields_ = [("n_seeds", u32), ("h2", u32), ("k", u32), ("n_dc", u32 * 16),
                ("dc", C.POINTER(u32) * 16)]


class Oracle:
    """Plain-C restatement (oracle/btl_oracle.c)."""

    def __init__(self):
        build_oracle()
        L = self.L = C.CDLL(ORACLE_SO)
        L.ora_seed.restype = u64
        L.ora_seed.argtypes = [C.c_ubyte]
        L.ora_srol.restype = u64
        L.ora_srol.argtypes = [u64]
        L.ora_sror.restype = u64
        L.ora_sror.argtypes = [u64]
        L.ora_srol_n.restype = u64
        L.ora_srol_n.argtypes = [u64, u32]
        L.ora_multi_mult.restype = u64
        L.ora_multi_mult.argtypes = [u32, u32]
        L.ora_splitmix64.restype = u64
        L.ora_splitmix64.argtypes = [u64]
        for name in ["ora_hash_seqs", "ora_bf_insert_seqs", "ora_bf_contains_seqs",
                     "ora_bf_insert_and_check_seqs", "ora_cbf_insert_seqs",
                     "ora_cbf_increment_all_seqs", "ora_cbf_mincount_seqs", "ora_cbf_contains_seqs",
                     "ora_st_hash_seqs", "ora_st_bf_insert_seqs", "ora_st_bf_contains_seqs",
                     "ora_st_cbf_insert_seqs", "ora_st_cbf_mincount_seqs", "ora_bf_popcount",
                     "ora_cbf_popcount", "ora_cbf_filtered_popcount"]:
            getattr(L, name).restype = u64
        L.ora_hash_seqs.argtypes = [u32, u32, u8p, u64p, u64, u64p, u8p]
        L.ora_bf_insert_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64]
        L.ora_bf_contains_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64, u8p, u8p, u64p]
        L.ora_bf_insert_and_check_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64, u8p, u8p]
        L.ora_cbf_insert_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64]
        L.ora_cbf_increment_all_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64]
        L.ora_cbf_mincount_seqs.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64, u8p, u8p]
        L.ora_cbf_contains_seqs.argtypes = [u8p, u64, u32, u32, u32, u8p, u64p, u64, u8p, u8p, u64p]
        SS = C.POINTER(_SeedSet)
        L.ora_seedset_parse.argtypes = [SS, C.POINTER(C.c_char_p), u32, u32, u32]
        L.ora_seedset_parse.restype = C.c_int
        L.ora_seedset_free.argtypes = [SS]
        L.ora_st_hash_seqs.argtypes = [SS, u8p, u64p, u64, u64p, u8p, u8p]
        L.ora_st_bf_insert_seqs.argtypes = [u8p, u64, SS, u8p, u64p, u64]
        L.ora_st_bf_contains_seqs.argtypes = [u8p, u64, SS, u8p, u64p, u64, u8p, u8p, u64p]
        L.ora_st_cbf_insert_seqs.argtypes = [u8p, u64, SS, u8p, u64p, u64]
        L.ora_st_cbf_mincount_seqs.argtypes = [u8p, u64, SS, u8p, u64p, u64, u8p, u8p]
        L.ora_bf_popcount.argtypes = [u8p, u64]
        L.ora_cbf_popcount.argtypes = [u8p, u64]
        L.ora_cbf_filtered_popcount.argtypes = [u8p, u64, u32]
        L.ora_bf_header.restype = C.c_int
        L.ora_bf_header.argtypes = [C.c_char_p, C.c_size_t, u64, u64, u32, u32, C.c_double, u64, u64]
        L.ora_cbf_header.restype = C.c_int
        L.ora_cbf_header.argtypes = [C.c_char_p, C.c_size_t, u64, u64, u32, u32, u32]
        L.ora_synth_genome.argtypes = [u8p, u64, u64, u64]
        L.ora_synth_reads.argtypes = [u8p, u64, u64, u32, u64, u64, u64, u64]
        L.ora_bench_bf.restype = C.c_double
        L.ora_bench_bf.argtypes = [u8p, u64, u32, u32, u8p, u64p, u64, C.c_int, C.c_int, u64p, u64p]
        L.ora_max_threads.restype = C.c_int

    # -- helpers
    def _seedset(self, seeds, h2, k):
        ss = _SeedSet()
        arr = (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])
        rc = self.L.ora_seedset_parse(C.byref(ss), arr, len(seeds), h2, k)
        if rc != 0:
            raise ValueError("bad seed set")
        return ss

    def hash_seqs(self, h, k, bases, off):
        n = bases.size
        hashes = np.zeros(n * h, np.uint64)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ora_hash_seqs(h, k, _p8(bases), _p64(off), off.size - 1, _p64(hashes), _p8(valid))
        return int(cnt), hashes.reshape(n, h), valid

    def st_hash_seqs(self, seeds, h2, k, bases, off):
        ss = self._seedset(seeds, h2, k)
        H = len(seeds) * h2
        n = bases.size
        hashes = np.zeros(n * H, np.uint64)
        strands = np.zeros(n * H, np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ora_st_hash_seqs(C.byref(ss), _p8(bases), _p64(off), off.size - 1, _p64(hashes),
                                      _p8(strands), _p8(valid))
        self.L.ora_seedset_free(C.byref(ss))
        return int(cnt), hashes.reshape(n, H), strands.reshape(n, H), valid

    def bf_insert_seqs(self, filt, m, h, k, bases, off):
        return int(self.L.ora_bf_insert_seqs(_p8(filt), m, h, k, _p8(bases), _p64(off), off.size - 1))

    def bf_contains_seqs(self, filt, m, h, k, bases, off):
        n = bases.size
        hits = np.zeros(nbits_bytes(n), np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        nh = u64(0)
        cnt = self.L.ora_bf_contains_seqs(_p8(filt), m, h, k, _p8(bases), _p64(off), off.size - 1,
                                          _p8(hits), _p8(valid), C.byref(nh))
        return int(cnt), int(nh.value), hits, valid

    def bf_insert_and_check_seqs(self, filt, m, h, k, bases, off):
        n = bases.size
        found = np.zeros(nbits_bytes(n), np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ora_bf_insert_and_check_seqs(_p8(filt), m, h, k, _p8(bases), _p64(off),
                                                  off.size - 1, _p8(found), _p8(valid))
        return int(cnt), found, valid

    def mibf_insert_bv_seqs(self, words, m, h, k, bases, off):
        """(n_kmers, n_collisions); words: uint64 array of ceil(m/64) words, updated in place"""
        colli = C.c_uint64()
        self.L.ora_mibf_insert_bv_seqs.restype = C.c_uint64
        n = self.L.ora_mibf_insert_bv_seqs(words.ctypes.data_as(C.c_void_p), C.c_uint64(m), C.c_uint(h), C.c_uint(k),
                                           _p8(bases), _p64(off), C.c_uint64(off.size - 1), C.byref(colli))
        return int(n), int(colli.value)

    def cbf_insert_seqs(self, cntr, m, h, k, bases, off):
        return int(self.L.ora_cbf_insert_seqs(_p8(cntr), m, h, k, _p8(bases), _p64(off), off.size - 1))

    def cbf_increment_all_seqs(self, cntr, m, h, k, bases, off):
        return int(self.L.ora_cbf_increment_all_seqs(_p8(cntr), m, h, k, _p8(bases), _p64(off),
                                                     off.size - 1))

    def cbf_mincount_seqs(self, cntr, m, h, k, bases, off):
        n = bases.size
        counts = np.zeros(max(n, 1), np.uint8)[:n]
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ora_cbf_mincount_seqs(_p8(cntr), m, h, k, _p8(bases), _p64(off), off.size - 1,
                                           _p8(counts), _p8(valid))
        return int(cnt), counts, valid

    def cbf_contains_seqs(self, cntr, m, h, k, thr, bases, off):
        n = bases.size
        hits = np.zeros(nbits_bytes(n), np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        nh = u64(0)
        cnt = self.L.ora_cbf_contains_seqs(_p8(cntr), m, h, k, thr, _p8(bases), _p64(off),
                                           off.size - 1, _p8(hits), _p8(valid), C.byref(nh))
        return int(cnt), int(nh.value), hits, valid

    def st_bf_insert_seqs(self, filt, m, seeds, h2, k, bases, off):
        ss = self._seedset(seeds, h2, k)
        r = int(self.L.ora_st_bf_insert_seqs(_p8(filt), m, C.byref(ss), _p8(bases), _p64(off),
                                             off.size - 1))
        self.L.ora_seedset_free(C.byref(ss))
        return r

    def st_bf_contains_seqs(self, filt, m, seeds, h2, k, bases, off):
        ss = self._seedset(seeds, h2, k)
        n = bases.size
        hits = np.zeros(nbits_bytes(n), np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        nh = u64(0)
        cnt = self.L.ora_st_bf_contains_seqs(_p8(filt), m, C.byref(ss), _p8(bases), _p64(off),
                                             off.size - 1, _p8(hits), _p8(valid), C.byref(nh))
        self.L.ora_seedset_free(C.byref(ss))
        return int(cnt), int(nh.value), hits, valid

    def st_cbf_insert_seqs(self, cntr, m, seeds, h2, k, bases, off):
        ss = self._seedset(seeds, h2, k)
        r = int(self.L.ora_st_cbf_insert_seqs(_p8(cntr), m, C.byref(ss), _p8(bases), _p64(off),
                                              off.size - 1))
        self.L.ora_seedset_free(C.byref(ss))
        return r

    def st_cbf_mincount_seqs(self, cntr, m, seeds, h2, k, bases, off):
        ss = self._seedset(seeds, h2, k)
        n = bases.size
        counts = np.zeros(max(n, 1), np.uint8)[:n]
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ora_st_cbf_mincount_seqs(_p8(cntr), m, C.byref(ss), _p8(bases), _p64(off),
                                              off.size - 1, _p8(counts), _p8(valid))
        self.L.ora_seedset_free(C.byref(ss))
        return int(cnt), counts, valid

    def bf_header(self, bits, nbytes, h, k, dfpr=0.0, nentry=0, tentry=0):
        buf = C.create_string_buffer(1024)
        n = self.L.ora_bf_header(buf, 1024, bits, nbytes, h, k, dfpr, nentry, tentry)
        return buf.raw[:n]

    def cbf_header(self, size, nbytes, h, k, bpc=8):
        buf = C.create_string_buffer(1024)
        n = self.L.ora_cbf_header(buf, 1024, size, nbytes, h, k, bpc)
        return buf.raw[:n]

    def synth_genome(self, start, n, seed):
        out = np.empty(n, np.uint8)
        self.L.ora_synth_genome(_p8(out), start, n, seed)
        return out

    def synth_reads(self, first, n_reads, read_len, g_len, gseed, rseed, g_start=0):
        out = np.empty(n_reads * read_len, np.uint8)
        self.L.ora_synth_reads(_p8(out), first, n_reads, read_len, g_start, g_len, gseed, rseed)
        return out


class Ref:
    """The reference's own headers compiled unmodified (oracle/_ref/libbtlref.so)."""

    @staticmethod
    def available():
        build_oracle()
        return os.path.exists(REF_SO)

    def __init__(self):
        build_oracle()
        L = self.L = C.CDLL(REF_SO)
        vp = C.c_void_p
        for n in ["ref_hash_seqs", "ref_st_hash_seqs", "ref_bf_size_bits", "ref_bf_size_bytes",
                  "ref_bf_pop", "ref_bf_insert_seqs", "ref_bf_contains_seqs",
                  "ref_bf_insert_and_check_seqs", "ref_st_bf_insert_seqs", "ref_st_bf_contains_seqs",
                  "ref_cbf_size", "ref_cbf_size_bytes", "ref_cbf_popcount",
                  "ref_cbf_filtered_popcount", "ref_cbf_insert_seqs", "ref_cbf_increment_all_seqs",
                  "ref_cbf_mincount_seqs", "ref_cbf_contains_seqs", "ref_st_cbf_insert_seqs",
                  "ref_st_cbf_mincount_seqs"]:
            getattr(L, n).restype = u64
        for n in ["ref_bf_new", "ref_bf_load", "ref_cbf_new", "ref_cbf_load"]:
            getattr(L, n).restype = vp
        L.ref_bf_data.restype = u8p
        L.ref_bf_fpr.restype = C.c_double
        cpp = C.POINTER(C.c_char_p)
        L.ref_hash_seqs.argtypes = [u32, u32, u8p, u64p, u64, u64p, u8p]
        L.ref_st_hash_seqs.argtypes = [cpp, u32, u32, u32, u8p, u64p, u64, u64p, u8p, u8p]
        L.ref_bf_new.argtypes = [u64, u32, u32]
        L.ref_bf_load.argtypes = [C.c_char_p]
        for n in ["ref_bf_free", "ref_bf_data", "ref_bf_size_bits", "ref_bf_size_bytes",
                  "ref_bf_hash_num", "ref_bf_kmer_size", "ref_bf_pop", "ref_bf_fpr", "ref_cbf_free",
                  "ref_cbf_size", "ref_cbf_size_bytes", "ref_cbf_hash_num", "ref_cbf_kmer_size",
                  "ref_cbf_popcount", "ref_cbf_filtered_popcount"]:
            getattr(L, n).argtypes = [vp]
        L.ref_bf_set_meta.argtypes = [vp, C.c_double, u64, u64]
        L.ref_bf_get_meta.argtypes = [vp, u64p, u64p]
        L.ref_bf_store.argtypes = [vp, C.c_char_p]
        L.ref_bf_insert_seqs.argtypes = [vp, u8p, u64p, u64]
        L.ref_bf_contains_seqs.argtypes = [vp, u8p, u64p, u64, u8p, u8p, u64p]
        L.ref_bf_insert_and_check_seqs.argtypes = [vp, u8p, u64p, u64, u8p, u8p]
        L.ref_st_bf_insert_seqs.argtypes = [vp, cpp, u32, u32, u8p, u64p, u64]
        L.ref_st_bf_contains_seqs.argtypes = [vp, cpp, u32, u32, u8p, u64p, u64, u8p, u8p, u64p]
        L.ref_cbf_new.argtypes = [u64, u32, u32, u32]
        L.ref_cbf_load.argtypes = [C.c_char_p, u32]
        L.ref_cbf_dump.argtypes = [vp, u8p]
        L.ref_cbf_store.argtypes = [vp, C.c_char_p]
        L.ref_cbf_insert_seqs.argtypes = [vp, u8p, u64p, u64]
        L.ref_cbf_increment_all_seqs.argtypes = [vp, u8p, u64p, u64]
        L.ref_cbf_mincount_seqs.argtypes = [vp, u8p, u64p, u64, u8p, u8p]
        L.ref_cbf_contains_seqs.argtypes = [vp, u8p, u64p, u64, u8p, u8p, u64p]
        L.ref_st_cbf_insert_seqs.argtypes = [vp, cpp, u32, u32, u8p, u64p, u64]
        L.ref_st_cbf_mincount_seqs.argtypes = [vp, cpp, u32, u32, u8p, u64p, u64, u8p, u8p]
        L.ref_kbf_new.restype = vp
        L.ref_kbf_new.argtypes = [u64, u32, u32]
        L.ref_kbf_free.argtypes = [vp]
        L.ref_kbf_data.restype = u8p
        L.ref_kbf_data.argtypes = [vp]
        L.ref_kbf_insert.argtypes = [vp, C.c_char_p]
        L.ref_kbf_contains.argtypes = [vp, C.c_char_p]
        L.ref_kbf_contains.restype = C.c_int
        L.ref_bench_bf.restype = C.c_double
        L.ref_bench_bf.argtypes = [vp, u8p, u64p, u64, C.c_int, C.c_int, u64p, u64p]
        L.ref_bench_cbf.restype = C.c_double
        L.ref_bench_cbf.argtypes = [vp, u8p, u64p, u64, C.c_int, C.c_int, u64p, u64p]
        L.ref_bench_st_bf.restype = C.c_double
        L.ref_bench_st_bf.argtypes = [vp, cpp, u32, u32, u8p, u64p, u64, C.c_int, C.c_int, u64p, u64p]
        L.ref_max_threads.restype = C.c_int

    @staticmethod
    def _seeds(seeds):
        return (C.c_char_p * len(seeds))(*[s.encode() for s in seeds])

    def hash_seqs(self, h, k, bases, off):
        n = bases.size
        hashes = np.zeros(n * h, np.uint64)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ref_hash_seqs(h, k, _p8(bases), _p64(off), off.size - 1, _p64(hashes), _p8(valid))
        return int(cnt), hashes.reshape(n, h), valid

    def st_hash_seqs(self, seeds, h2, k, bases, off):
        H = len(seeds) * h2
        n = bases.size
        hashes = np.zeros(n * H, np.uint64)
        strands = np.zeros(n * H, np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        cnt = self.L.ref_st_hash_seqs(self._seeds(seeds), len(seeds), h2, k, _p8(bases), _p64(off),
                                      off.size - 1, _p64(hashes), _p8(strands), _p8(valid))
        return int(cnt), hashes.reshape(n, H), strands.reshape(n, H), valid

    # BloomFilter handle helpers
    def bf_new(self, bits, h, k):
        return self.L.ref_bf_new(bits, h, k)

    def bf_bytes(self, f):
        n = self.L.ref_bf_size_bytes(f)
        return np.ctypeslib.as_array(self.L.ref_bf_data(f), shape=(n,)).copy() if n else np.zeros(0, np.uint8)

    def bf_set_bytes(self, f, arr):
        n = self.L.ref_bf_size_bytes(f)
        C.memmove(self.L.ref_bf_data(f), arr.ctypes.data, n)

    def bf_contains_seqs(self, f, bases, off):
        n = bases.size
        hits = np.zeros(nbits_bytes(n), np.uint8)
        valid = np.zeros(nbits_bytes(n), np.uint8)
        nh = u64(0)
        cnt = self.L.ref_bf_contains_seqs(f, _p8(bases), _p64(off), off.size - 1, _p8(hits), _p8(valid),
                                          C.byref(nh))
        return int(cnt), int(nh.value), hits, valid

    def cbf_bytes(self, f):
        n = self.L.ref_cbf_size(f)
        out = np.zeros(n, np.uint8)
        self.L.ref_cbf_dump(f, _p8(out))
        return out
