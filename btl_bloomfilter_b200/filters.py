"""Host-side mirror of the reference's filter classes over the CUDA C ABI (include/btlbf.h).

Same names, argument meaning and error behaviour as the reference's C++ classes
(BloomFilter.hpp, CountingBloomFilter.hpp, BloomFilterUtil.h), plus the batched
insertSeqs / containsSeqs entry points.  Every operation on filter contents runs on the GPU
through libbtlbf_cuda.so; there is no numpy / CPU implementation in this module.

The C++ twin of this module (for C++ callers, SWIG) is include/btlbf/*.hpp.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import BLOOM, COUNTING8, BtlbfError, check, lib


# ---------------------------------------------------------------- batches
def as_batch(seqs):
    """list of str/bytes, or (bases uint8[n], offsets uint64[m+1]) -> contiguous (bases, offsets)."""
    if isinstance(seqs, tuple) and len(seqs) == 2:
        bases = np.ascontiguousarray(seqs[0], dtype=np.uint8)
        offsets = np.ascontiguousarray(seqs[1], dtype=np.uint64)
        return bases, offsets
    if isinstance(seqs, (str, bytes, bytearray)):
        seqs = [seqs]
    bs = [s.encode("latin-1") if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8)
    return bases, offsets


def _ptr(a):
    return None if a is None or a.size == 0 else C.c_void_p(a.ctypes.data)


class PackedBatch:
    """A flat batch as 2 bits per base (include/btlbf.h "2-bit packed input"): codes (4 bases per byte, A0 C1 G2 T3,
    the order of vendor/nthash.hpp:51), invalid (one bit per base, or None when every base is one of ACGTUacgtu) and
    the offsets of the ASCII batch (they keep counting bases)."""

    def __init__(self, codes, invalid, offsets, n_invalid=0):
        self.codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.invalid = None if invalid is None else np.ascontiguousarray(invalid, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.n_bases = int(self.offsets[-1]) if self.offsets.size else 0
        self.n_invalid = int(n_invalid)


def pack_seqs(seqs, threads=0, codes_out=None, invalid_out=None, keep_invalid=False):
    """ASCII batch -> PackedBatch through the library's host packer (btlbf_pack_seqs; `threads` 0 = all cores).
    codes_out / invalid_out: optional preallocated (e.g. pinned) uint8 arrays of ceil(n/4) / ceil(n/8) bytes.
    The invalid plane is dropped when the input has no invalid base (unless keep_invalid)."""
    bases, off = as_batch(seqs)
    n = int(bases.size)
    codes = np.zeros((n + 3) // 4, np.uint8) if codes_out is None else codes_out
    invalid = np.zeros((n + 7) // 8, np.uint8) if invalid_out is None else invalid_out
    if codes.size < (n + 3) // 4 or invalid.size < (n + 7) // 8:
        raise ValueError("packed outputs need %d and %d bytes" % ((n + 3) // 4, (n + 7) // 8))
    bad = C.c_uint64()
    check(lib().btlbf_pack_seqs(_ptr(bases), n, _ptr(codes), _ptr(invalid), int(threads), C.byref(bad)))
    return PackedBatch(codes, invalid if (bad.value or keep_invalid) else None, off, bad.value)


def _p64(a):
    return a.ctypes.data_as(_capi.u64p)


def bit_bytes(n):
    return (int(n) + 31) // 32 * 4


def unpack_bits(bits, n):
    """little-endian packed bit array -> bool[n]"""
    return np.unpackbits(bits, bitorder="little")[: int(n)].astype(bool)


class QueryResult:
    """Per-window results of a batched query, indexed by the flat position of the window's first base
    (the reference's ntHashIterator::pos() + the sequence's offset)."""

    def __init__(self, n_bases, offsets, k, hit_bits, valid_bits, n_kmers, n_hits, counts=None):
        self.n_bases = int(n_bases)
        self.offsets = offsets
        self.k = k
        self.hit_bits = hit_bits
        self.valid_bits = valid_bits
        self.n_kmers = int(n_kmers)
        self.n_hits = int(n_hits)
        self.counts = counts

    @property
    def hits(self):
        return unpack_bits(self.hit_bits, self.n_bases)

    @property
    def valid(self):
        return unpack_bits(self.valid_bits, self.n_bases)

    def per_sequence(self):
        """[(positions of valid k-mers, hit flag of each)] per sequence, in iterator order."""
        hits, valid = self.hits, self.valid
        out = []
        for i in range(len(self.offsets) - 1):
            a, b = int(self.offsets[i]), int(self.offsets[i + 1])
            pos = np.nonzero(valid[a:b])[0]
            out.append((pos, hits[a:b][pos]))
        return out


# ---------------------------------------------------------------- context
class Context:
    """One per GPU: streams, staging buffers (btlbf_ctx)."""

    _default = {}

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.btlbf_ctx_create(device, C.byref(h)))
        self.handle = h
        self.device = device

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def set_option(self, key, value):
        check(self.L.btlbf_ctx_set_option(self.handle, key.encode(), int(value)))

    def set_stream(self, cuda_stream):
        """Run the context's kernels on a caller-owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream).
        None restores the context's own stream; 0 means the legacy default stream (cudaStreamLegacy)."""
        if cuda_stream is None:
            ptr = 0
        elif cuda_stream == 0:
            ptr = 1  # cudaStreamLegacy: a NULL handle would mean "the context's own stream" to the C ABI
        else:
            ptr = cuda_stream
        check(self.L.btlbf_ctx_set_stream(self.handle, C.c_void_p(ptr)))

    def sync(self):
        check(self.L.btlbf_ctx_sync(self.handle))

    def flush(self):
        """Queue all deferred work on the active stream: pass 2 of the partitioned BloomFilter build for the k-mers
        parked by earlier insert calls, queued per-k-mer updates, and a wait for the background stream.  Required
        before anything outside this library (torch, NCCL, a peer GPU, your own kernel) reads filter memory
        obtained from device_ptr(); see include/btlbf.h "DEFERRED WORK"."""
        check(self.L.btlbf_ctx_flush(self.handle))

    @property
    def aux_stream(self):
        p = C.c_void_p()
        check(self.L.btlbf_ctx_aux_stream(self.handle, C.byref(p)))
        return p.value

    def counter(self, name):
        n = C.c_uint64()
        check(self.L.btlbf_ctx_counter(self.handle, name.encode(), C.byref(n)))
        return n.value

    @property
    def two_level_passes(self):
        return self.counter("two_level_passes")

    @property
    def launch_count(self):
        n = C.c_uint64()
        check(self.L.btlbf_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def synth_genome_device(self, d_out, start, n, seed):
        check(self.L.btlbf_synth_genome_dev(self.handle, C.c_void_p(d_out), start, n, seed))

    def synth_reads_device(self, d_out, first_read, n_reads, read_len, g_start, g_len, genome_seed, read_seed):
        check(self.L.btlbf_synth_reads_dev(self.handle, C.c_void_p(d_out), first_read, n_reads, read_len, g_start,
                                           g_len, genome_seed, read_seed))

    def random_access_probe(self, d_array, nbytes, n_access, mode):
        ms = C.c_float()
        check(self.L.btlbf_random_access_probe(self.handle, C.c_void_p(d_array), nbytes, n_access, mode, C.byref(ms)))
        return ms.value

    def hash_seqs(self, seqs, hashNum, kmerSize, seeds=None, h2=1):
        """Raw iterator output (ntHashIterator, or stHashIterator when seeds are given):
        (n_kmers, hashes[n_bases, H], strands[n_bases, H], valid_bits)."""
        bases, off = as_batch(seqs)
        n = bases.size
        H = len(seeds) * h2 if seeds else hashNum
        hashes = np.zeros((n, H), np.uint64)
        strands = np.zeros((n, H), np.uint8)
        valid = np.zeros(bit_bytes(n), np.uint8)
        nk = C.c_uint64()
        sp, ns = _seed_array(seeds)
        check(self.L.btlbf_hash_seqs(self.handle, hashNum, kmerSize, sp, ns, h2 if seeds else 0, _ptr(bases),
                                     _p64(off), off.size - 1, _ptr(hashes), _ptr(strands), _ptr(valid),
                                     C.byref(nk)))
        return nk.value, hashes, strands, valid

    def close(self):
        if self.handle:
            self.L.btlbf_ctx_destroy(self.handle)
            self.handle = None


def _seed_array(seeds):
    if not seeds:
        return None, 0
    arr = (C.c_char_p * len(seeds))(*[s.encode() if isinstance(s, str) else s for s in seeds])
    return arr, len(seeds)


# ---------------------------------------------------------------- shared machinery
class _DeviceFilter:
    KIND = BLOOM

    def __init__(self):
        self._ctx = None
        self._h = None
        self._seeds = None
        self._h2 = 0

    # -- lifetime
    def _create(self, ctx, size, hashNum, kmerSize, threshold):
        self._ctx = ctx or Context.default()
        h = C.c_void_p()
        check(self._ctx.L.btlbf_filter_create(self._ctx.handle, self.KIND, size, hashNum, kmerSize, threshold,
                                              C.byref(h)))
        self._h = h

    @classmethod
    def from_device_memory(cls, tensor, size, hashNum, kmerSize, threshold=0, ctx=None):
        """Filter over caller-owned device memory (a torch uint8 tensor, 16-byte aligned, at least
        round_up(bytes, 16) long; used as is, not cleared): btlbf_filter_wrap."""
        self = cls.__new__(cls)
        _DeviceFilter.__init__(self)
        if cls.KIND == BLOOM:
            cls._check_size(size)
            self.m_dFPR, self.m_nEntry, self.m_tEntry, self.m_FPR = 0.0, 0, 0, 0.0
        else:
            self._threshold = int(threshold)
        self._ctx = ctx or Context.default()
        self._tensor = tensor
        h = C.c_void_p()
        check(self._ctx.L.btlbf_filter_wrap(self._ctx.handle, cls.KIND, size, hashNum, kmerSize, threshold,
                                            C.c_void_p(tensor.data_ptr()), tensor.numel() * tensor.element_size(),
                                            C.byref(h)))
        self._h = h
        return self

    def _release(self):
        if self._h is not None and self._ctx is not None and self._ctx.handle:
            self._ctx.L.btlbf_filter_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    @property
    def _L(self):
        return self._ctx.L

    def _info(self):
        kind, size, nbytes = C.c_int(), C.c_uint64(), C.c_uint64()
        h, k, thr = C.c_uint(), C.c_uint(), C.c_uint()
        check(self._L.btlbf_filter_info(self._h, C.byref(kind), C.byref(size), C.byref(nbytes), C.byref(h),
                                        C.byref(k), C.byref(thr)))
        return kind.value, size.value, nbytes.value, h.value, k.value, thr.value

    # -- raw array
    def to_numpy(self):
        """The raw array exactly as the reference keeps it in host memory (m_filter)."""
        out = np.empty(self.sizeInBytes(), np.uint8)
        check(self._L.btlbf_filter_download(self._h, _ptr(out), out.size))
        return out

    def from_numpy(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        check(self._L.btlbf_filter_upload(self._h, _ptr(arr), arr.size))

    def clear(self):
        check(self._L.btlbf_filter_clear(self._h))

    def device_ptr(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(self._L.btlbf_filter_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- spaced seeds (stHashIterator): hashNum must equal len(seeds)*h2
    def setSeeds(self, seeds, h2=1):
        sp, ns = _seed_array(seeds)
        check(self._L.btlbf_filter_set_seeds(self._h, sp, ns, h2))
        self._seeds, self._h2 = (list(seeds), h2) if seeds else (None, 0)

    # -- batched entry points
    def insertSeqs(self, seqs):
        """ntHashIterator (or stHashIterator) over every sequence + insert of every k-mer
        (README.md:30-43 / BloomFilterUtil.h:10-17 for a whole batch).  Returns the k-mer count."""
        bases, off = as_batch(seqs)
        nk = C.c_uint64()
        check(self._L.btlbf_insert_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, C.byref(nk)))
        return nk.value

    def containsSeqs(self, seqs, hit_out=None, valid_out=None, want_valid=True):
        """contains() of every k-mer of every sequence (README.md:46-57): QueryResult.
        hit_out / valid_out: optional preallocated uint8 arrays of bit_bytes(n_bases) bytes (e.g. pinned)."""
        bases, off = as_batch(seqs)
        n = bases.size
        hits = np.zeros(bit_bytes(n), np.uint8) if hit_out is None else hit_out
        valid = valid_out if valid_out is not None else (np.zeros(bit_bytes(n), np.uint8) if want_valid else None)
        if hits.size < bit_bytes(n) or (valid is not None and valid.size < bit_bytes(n)):
            raise ValueError("output arrays need %d bytes" % bit_bytes(n))
        nk, nh = C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_contains_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(hits), _ptr(valid),
                                          C.byref(nk), C.byref(nh)))
        return QueryResult(n, off, self.getKmerSize(), hits, valid, nk.value, nh.value)

    # -- FASTA / FASTQ files (the record loop of swig/writeBloom_rolling.cpp:19-59, parsed by native threads)
    def insertFile(self, path, threads=0):
        """Insert every k-mer of every record of a FASTA / FASTQ file: (n_records, n_kmers)."""
        ns, nk = C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_insert_file(self._h, str(path).encode(), int(threads), C.byref(ns), C.byref(nk)))
        return ns.value, nk.value

    def queryFile(self, path, threads=0):
        """contains() of every k-mer of every record of a FASTA / FASTQ file: (n_records, n_kmers, n_hits)."""
        ns, nk, nh = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_query_file(self._h, str(path).encode(), int(threads), C.byref(ns), C.byref(nk), C.byref(nh)))
        return ns.value, nk.value, nh.value

    # -- streaming (asynchronous) forms: all arrays must stay alive and untouched until Context.sync()
    def insertSeqsAsync(self, seqs, counts_out):
        """Queue insertSeqs; counts_out: uint64[2] (ideally pinned) that receives {n_kmers, n_hits}."""
        bases, off = as_batch(seqs)
        check(self._L.btlbf_insert_seqs_async(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(counts_out)))
        return bases, off  # the caller keeps these alive until sync()

    def containsSeqsAsync(self, seqs, hit_out, counts_out, valid_out=None):
        bases, off = as_batch(seqs)
        if hit_out.size < bit_bytes(bases.size):
            raise ValueError("hit_out needs %d bytes" % bit_bytes(bases.size))
        check(self._L.btlbf_contains_seqs_async(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(hit_out),
                                                _ptr(valid_out), _ptr(counts_out)))
        return bases, off

    # -- 2-bit packed input (btlbf_*_seqs_packed): same results as the ASCII calls on the same sequences
    def insertSeqsPacked(self, pk):
        nk = C.c_uint64()
        check(self._L.btlbf_insert_seqs_packed(self._h, _ptr(pk.codes), _ptr(pk.invalid), _p64(pk.offsets),
                                               pk.offsets.size - 1, C.byref(nk)))
        return nk.value

    def containsSeqsPacked(self, pk, hit_out=None, valid_out=None, want_valid=True):
        n = pk.n_bases
        hits = np.zeros(bit_bytes(n), np.uint8) if hit_out is None else hit_out
        valid = valid_out if valid_out is not None else (np.zeros(bit_bytes(n), np.uint8) if want_valid else None)
        if hits.size < bit_bytes(n) or (valid is not None and valid.size < bit_bytes(n)):
            raise ValueError("output arrays need %d bytes" % bit_bytes(n))
        nk, nh = C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_contains_seqs_packed(self._h, _ptr(pk.codes), _ptr(pk.invalid), _p64(pk.offsets),
                                                 pk.offsets.size - 1, _ptr(hits), _ptr(valid), C.byref(nk), C.byref(nh)))
        return QueryResult(n, pk.offsets, self.getKmerSize(), hits, valid, nk.value, nh.value)

    def insertSeqsPackedAsync(self, pk, counts_out):
        """Queue insertSeqsPacked; the batch and counts_out stay alive and untouched until Context.sync()."""
        check(self._L.btlbf_insert_seqs_packed_async(self._h, _ptr(pk.codes), _ptr(pk.invalid), _p64(pk.offsets),
                                                     pk.offsets.size - 1, _ptr(counts_out)))

    def containsSeqsPackedAsync(self, pk, hit_out, counts_out, valid_out=None):
        if hit_out.size < bit_bytes(pk.n_bases):
            raise ValueError("hit_out needs %d bytes" % bit_bytes(pk.n_bases))
        check(self._L.btlbf_contains_seqs_packed_async(self._h, _ptr(pk.codes), _ptr(pk.invalid), _p64(pk.offsets),
                                                       pk.offsets.size - 1, _ptr(hit_out), _ptr(valid_out),
                                                       _ptr(counts_out)))

    def insertSeqsPackedDevice(self, d_codes, d_invalid, n_bases, d_offsets, n_seqs, d_stats=0):
        check(self._L.btlbf_insert_seqs_packed_dev(self._h, C.c_void_p(d_codes), C.c_void_p(d_invalid or 0), n_bases,
                                                   C.c_void_p(d_offsets), n_seqs, C.c_void_p(d_stats or 0)))

    def containsSeqsPackedDevice(self, d_codes, d_invalid, n_bases, d_offsets, n_seqs, d_hit_bits=0, d_valid_bits=0,
                                 d_stats=0):
        check(self._L.btlbf_contains_seqs_packed_dev(self._h, C.c_void_p(d_codes), C.c_void_p(d_invalid or 0), n_bases,
                                                     C.c_void_p(d_offsets), n_seqs, C.c_void_p(d_hit_bits or 0),
                                                     C.c_void_p(d_valid_bits or 0), C.c_void_p(d_stats or 0)))

    # -- device-resident batches (asynchronous on the context's stream; pointers are raw device addresses)
    def insertSeqsDevice(self, d_bases, n_bases, d_offsets, n_seqs, d_stats=0):
        check(self._L.btlbf_insert_seqs_dev(self._h, C.c_void_p(d_bases), n_bases, C.c_void_p(d_offsets), n_seqs,
                                            C.c_void_p(d_stats or 0)))

    def containsSeqsDevice(self, d_bases, n_bases, d_offsets, n_seqs, d_hit_bits=0, d_valid_bits=0, d_stats=0):
        check(self._L.btlbf_contains_seqs_dev(self._h, C.c_void_p(d_bases), n_bases, C.c_void_p(d_offsets), n_seqs,
                                              C.c_void_p(d_hit_bits or 0), C.c_void_p(d_valid_bits or 0),
                                              C.c_void_p(d_stats or 0)))

    def flushParts(self, chunk, n_chunks):
        """btlbf_filter_flush_parts: pass 2 of the parked build for chunk `chunk` of `n_chunks`; the byte range it covers."""
        lo, hi = C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_filter_flush_parts(self._h, int(chunk), int(n_chunks), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def orderedStats(self):
        d, r = C.c_uint64(), C.c_uint64()
        check(self._L.btlbf_filter_ordered_stats(self._h, C.byref(d), C.byref(r)))
        return d.value, r.value

    def _hashes(self, hashes):
        a = np.ascontiguousarray(hashes, dtype=np.uint64)
        h = self.getHashNum()
        if a.ndim == 1:
            a = a[:h].reshape(1, -1)
        if a.shape[1] < h:
            raise IndexError("need %d hash values per k-mer" % h)
        return np.ascontiguousarray(a[:, :h])


# ---------------------------------------------------------------- BloomFilter
class BloomFilter(_DeviceFilter):
    """BloomFilter.hpp: bit array of `filterSize` bits, `hashNum` hash values per k-mer.

    BloomFilter(filterSize, hashNum, kmerSize)             de novo (BloomFilter.hpp:66-78)
    BloomFilter(expectedElemNum, fpr, hashNum, kmerSize)   sized from an FPR (:85-104)
    BloomFilter(path)                                      from a BTLBloomFilter_v1 file (:106-110)
    """
    KIND = BLOOM

    def __init__(self, *args, ctx=None):
        super().__init__()
        self.m_dFPR = 0.0
        self.m_nEntry = 0
        self.m_tEntry = 0
        self.m_FPR = 0.0
        if len(args) == 1 and isinstance(args[0], (str, bytes)):
            self._ctx = ctx or Context.default()
            self.loadFilter(args[0])
        elif len(args) == 3:
            size, h, k = args
            self._check_size(size)
            self._create(ctx, size, h, k, 0)
        elif len(args) == 4:
            n, fpr, h, k = args
            self.m_dFPR = float(fpr)
            if h == 0:
                h = self.calcOptiHashNum(fpr)
            size = self.calcOptimalSize(n, fpr, h)
            self._create(ctx, size, h, k, 0)
        else:
            raise TypeError("BloomFilter(filterSize, hashNum, kmerSize) | (n, fpr, hashNum, kmerSize) | (path)")

    @staticmethod
    def _check_size(size):
        if size % 8 != 0:  # BloomFilter.hpp:389-394 prints this and exits
            raise ValueError('ERROR: Filter Size "%d" is not a multiple of 8.' % size)

    @staticmethod
    def calcOptimalSize(entries, fpr, hashNum):
        """BloomFilter.hpp:406-413 (returns a multiple of 64)."""
        v = int(-float(entries) * float(hashNum) / math.log(1.0 - math.pow(fpr, 1.0 / float(hashNum))))
        return v + (64 - v % 64)

    @staticmethod
    def calcOptiHashNum(fpr):
        return int(-math.log(fpr) / math.log(2))  # BloomFilter.hpp:419

    # -- per-k-mer interface (precomputed hash values, as in the reference)
    def insert(self, hashes):
        a = self._hashes(hashes)
        check(self._L.btlbf_insert_hashes(self._h, _p64(a), a.shape[0], None))

    def insertAndCheck(self, hashes):
        a = self._hashes(hashes)
        found = np.zeros(a.shape[0], np.uint8)
        check(self._L.btlbf_insert_hashes(self._h, _p64(a), a.shape[0], _ptr(found)))
        return bool(found[0]) if a.shape[0] == 1 else found.astype(bool)

    def contains(self, hashes):
        a = self._hashes(hashes)
        hit = np.zeros(a.shape[0], np.uint8)
        check(self._L.btlbf_contains_hashes(self._h, _p64(a), a.shape[0], _ptr(hit)))
        return bool(hit[0]) if a.shape[0] == 1 else hit.astype(bool)

    # -- batched
    def insertAndCheckSeqs(self, seqs):
        """insertAndCheck of every k-mer in reference order (BloomFilter.hpp:200-214): QueryResult whose
        hits are the k-mers that were already present."""
        bases, off = as_batch(seqs)
        n = bases.size
        found = np.zeros(bit_bytes(n), np.uint8)
        valid = np.zeros(bit_bytes(n), np.uint8)
        nk = C.c_uint64()
        check(self._L.btlbf_insert_and_check_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(found),
                                                  _ptr(valid), C.byref(nk)))
        nh = int(np.unpackbits(found).sum())
        return QueryResult(n, off, self.getKmerSize(), found, valid, nk.value, nh)

    # -- file layout
    def storeFilter(self, path):
        import sys
        sys.stderr.write("Writing a %d byte filter to %s on disk.\n" % (self.sizeInBytes(), path))
        check(self._L.btlbf_filter_store(self._h, str(path).encode(), self.m_dFPR, self.m_nEntry, self.m_tEntry))

    def loadFilter(self, path):
        h = C.c_void_p()
        dfpr, ne, te = C.c_double(), C.c_uint64(), C.c_uint64()
        check(self._ctx.L.btlbf_filter_load(self._ctx.handle, str(path).encode() if isinstance(path, str) else path,
                                            BLOOM, 0, C.byref(h), C.byref(dfpr), C.byref(ne), C.byref(te)))
        self._release()
        self._h = h
        self.m_dFPR, self.m_nEntry, self.m_tEntry = dfpr.value, ne.value, te.value

    def header(self):
        n = C.c_size_t()
        buf = C.create_string_buffer(1024)
        check(self._L.btlbf_format_header(BLOOM, self.getFilterSize(), self.sizeInBytes(), self.getHashNum(),
                                          self.getKmerSize(), self.m_dFPR, self.m_nEntry, self.m_tEntry, buf, 1024,
                                          C.byref(n)))
        return buf.raw[: n.value]

    # -- statistics
    def getPop(self):
        n = C.c_uint64()
        check(self._L.btlbf_filter_popcount(self._h, C.byref(n)))
        return n.value

    def getHashNum(self):
        return self._info()[3]

    def getKmerSize(self):
        return self._info()[4]

    def getFilterSize(self):
        return self._info()[1]

    def sizeInBytes(self):
        return self._info()[2]

    def getFPR(self):
        self.m_FPR = math.pow(float(self.getPop()) / float(self.getFilterSize()), float(self.getHashNum()))
        return self.m_FPR

    def getFPRPrecompute(self):
        return self.m_FPR

    def calcFPR_numInserted(self, numEntr):
        m, h = float(self.getFilterSize()), self.getHashNum()
        return math.pow(1.0 - math.pow(1.0 - 1.0 / m, float(numEntr) * h), float(h))

    def getFPR_numEle(self):
        assert self.m_nEntry > 0
        return self.calcFPR_numInserted(self.m_nEntry)

    def getRedudancyFPR(self):
        assert self.m_nEntry > 0
        total = math.log(self.calcFPR_numInserted(1))
        for i in range(2, self.m_nEntry):
            total = math.log(math.exp(total) + self.calcFPR_numInserted(i))
        return math.exp(total) / self.m_nEntry

    def getnEntry(self):
        return self.m_nEntry

    def gettEntry(self):
        return self.m_tEntry

    def setnEntry(self, v):
        self.m_nEntry = int(v)

    def settEntry(self, v):
        self.m_tEntry = int(v)


class BitVector(_DeviceFilter):
    """Level-1 bit vector of a multi-index Bloom filter (MIBFConstructSupport.hpp:36-46: an sdsl::bit_vector of
    `size` bits, any size): insertSeqs = insertBV (:76-87), insertBVColli (:55-74), containsSeqs, getPop.
    to_numpy() returns the 64-bit words as bytes (little-endian), i.e. m_bv.data()."""
    KIND = 2

    def __init__(self, size, hashNum, kmerSize, ctx=None):
        super().__init__()
        self._create(ctx, int(size), hashNum, kmerSize, 0)

    def insertBVColli(self, seqs):
        """Insert every k-mer; returns (n_kmers, n_collisions): collisions = k-mers whose h bits were all set
        already, in sequence order (the reference's single-threaded result)."""
        bases, off = as_batch(seqs)
        n = bases.size
        found = np.zeros(bit_bytes(n), np.uint8)
        nk = C.c_uint64()
        check(self._L.btlbf_insert_and_check_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(found), None,
                                                  C.byref(nk)))
        return nk.value, int(np.unpackbits(found).sum())

    def getPop(self):
        n = C.c_uint64()
        check(self._L.btlbf_filter_popcount(self._h, C.byref(n)))
        return n.value

    def size(self):
        return self._info()[1]

    def sizeInBytes(self):
        return self._info()[2]

    def getKmerSize(self):
        return self._info()[4]


class KmerBloomFilter(BloomFilter):
    """KmerBloomFilter.hpp:17-74 (what swig/BloomFilter.i exports as "BloomFilter"): insert / contains also take
    one k-mer as text.  The k-mer is hashed on the GPU with the iterator-consistent canonical ntHash (equal to the
    reference's NTC64(kmer,k) for k % 4 != 0; for k % 4 == 0 the reference's value is undefined behaviour)."""

    def _is_text(self, x):
        return isinstance(x, (str, bytes, bytearray))

    def insert(self, kmer_or_hashes):
        if self._is_text(kmer_or_hashes):
            self.insertSeqs([kmer_or_hashes[: self.getKmerSize()]])
        else:
            super().insert(kmer_or_hashes)

    def contains(self, kmer_or_hashes):
        if self._is_text(kmer_or_hashes):
            return self.containsSeqs([kmer_or_hashes[: self.getKmerSize()]]).n_hits == 1
        return super().contains(kmer_or_hashes)


def insertSeq(bloom, seq, hashNum=None, kmerSize=None):
    """BloomFilterUtil.h:10-17: load every k-mer of one sequence into the filter."""
    if hashNum is not None and hashNum != bloom.getHashNum():
        raise ValueError("hashNum does not match the filter")
    if kmerSize is not None and kmerSize != bloom.getKmerSize():
        raise ValueError("kmerSize does not match the filter")
    return bloom.insertSeqs([seq])


# ---------------------------------------------------------------- CountingBloomFilter<uint8_t>
class CountingBloomFilter(_DeviceFilter):
    """CountingBloomFilter.hpp with T = uint8_t.

    CountingBloomFilter(sizeInBytes, hashNum, kmerSize, countThreshold)   (:29-50; size rounded up to x8)
    CountingBloomFilter(path, countThreshold)                              (:260-266)
    """
    KIND = COUNTING8

    def __init__(self, *args, ctx=None):
        super().__init__()
        if len(args) == 2 and isinstance(args[0], (str, bytes)):
            self._ctx = ctx or Context.default()
            self._threshold = int(args[1])
            self.loadFilter(args[0])
        elif len(args) == 4:
            nbytes, h, k, thr = args
            rem = nbytes % 8
            if rem:
                nbytes = nbytes + 8 - rem
            self._threshold = int(thr)
            self._create(ctx, nbytes, h, k, thr)
        else:
            raise TypeError("CountingBloomFilter(sizeInBytes, hashNum, kmerSize, countThreshold) | (path, thr)")

    def __getitem__(self, i):
        return int(self.to_numpy()[i])

    # -- per-k-mer interface
    def minCount(self, hashes):
        a = self._hashes(hashes)
        out = np.zeros(a.shape[0], np.uint8)
        check(self._L.btlbf_mincount_hashes(self._h, _p64(a), a.shape[0], _ptr(out)))
        return int(out[0]) if a.shape[0] == 1 else out

    def contains(self, hashes):
        a = self._hashes(hashes)
        out = np.zeros(a.shape[0], np.uint8)
        check(self._L.btlbf_contains_hashes(self._h, _p64(a), a.shape[0], _ptr(out)))
        return bool(out[0]) if a.shape[0] == 1 else out.astype(bool)

    def insert(self, hashes):
        a = self._hashes(hashes)
        check(self._L.btlbf_insert_hashes(self._h, _p64(a), a.shape[0], None))

    incrementMin = insert

    def insertAndCheck(self, hashes):
        a = self._hashes(hashes)
        found = np.zeros(a.shape[0], np.uint8)
        check(self._L.btlbf_insert_hashes(self._h, _p64(a), a.shape[0], _ptr(found)))
        return bool(found[0]) if a.shape[0] == 1 else found.astype(bool)

    def incrementAll(self, hashes):
        a = self._hashes(hashes)
        check(self._L.btlbf_increment_all_hashes(self._h, _p64(a), a.shape[0]))

    # -- batched
    def minCountSeqs(self, seqs):
        bases, off = as_batch(seqs)
        n = bases.size
        counts = np.zeros(n, np.uint8)
        valid = np.zeros(bit_bytes(n), np.uint8)
        nk = C.c_uint64()
        check(self._L.btlbf_mincount_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, _ptr(counts), _ptr(valid),
                                          C.byref(nk)))
        return QueryResult(n, off, self.getKmerSize(), np.zeros(bit_bytes(n), np.uint8), valid, nk.value, 0,
                           counts=counts)

    def incrementAllSeqs(self, seqs):
        bases, off = as_batch(seqs)
        nk = C.c_uint64()
        check(self._L.btlbf_increment_all_seqs(self._h, _ptr(bases), _p64(off), off.size - 1, C.byref(nk)))
        return nk.value

    # -- accessors
    def getKmerSize(self):
        return self._info()[4]

    def getHashNum(self):
        return self._info()[3]

    def threshold(self):
        return self._info()[5]

    def size(self):
        return self._info()[1]

    def sizeInBytes(self):
        return self._info()[2]

    def popCount(self):
        n = C.c_uint64()
        check(self._L.btlbf_filter_popcount(self._h, C.byref(n)))
        return n.value

    def filtered_popcount(self):
        n = C.c_uint64()
        check(self._L.btlbf_filter_count_ge(self._h, self.threshold(), C.byref(n)))
        return n.value

    def FPR(self):
        return math.pow(float(self.popCount()) / float(self.size()), self.getHashNum())

    def filtered_FPR(self):
        return math.pow(float(self.filtered_popcount()) / float(self.size()), self.getHashNum())

    # -- file layout
    def storeFilter(self, path):
        import sys
        sys.stderr.write("Writing a %d byte filter to %s on disk.\n" % (self.sizeInBytes(), path))
        check(self._L.btlbf_filter_store(self._h, str(path).encode(), 0.0, 0, 0))

    def loadFilter(self, path):
        h = C.c_void_p()
        check(self._ctx.L.btlbf_filter_load(self._ctx.handle, str(path).encode() if isinstance(path, str) else path,
                                            COUNTING8, self._threshold, C.byref(h), None, None, None))
        self._release()
        self._h = h

    def header(self):
        n = C.c_size_t()
        buf = C.create_string_buffer(1024)
        check(self._L.btlbf_format_header(COUNTING8, self.size(), self.sizeInBytes(), self.getHashNum(),
                                          self.getKmerSize(), 0.0, 0, 0, buf, 1024, C.byref(n)))
        return buf.raw[: n.value]
