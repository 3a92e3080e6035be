"""FASTA / FASTQ ingest (btlbf_seqfile_*, btlbf_insert_file, btlbf_query_file): the parser against a plain
Python reader of the same files (CPU), and the file calls against the batched calls on the parsed
sequences (GPU).  Reference behaviour: swig/writeBloom_rolling.cpp:19-59 (one insertSeq per FASTA record)."""
import collections
import ctypes as C

import numpy as np
import pytest

from btl_bloomfilter_b200 import lib


def write_fasta(path, rng, n_records, max_len, width, crlf=False, comments=False, lower=0.05, p_n=0.01):
    seqs = []
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as fh:
        for i in range(n_records):
            n = int(rng.integers(0, max_len))
            s = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
            s = np.where(rng.random(n) < lower, s | 0x20, s)
            s = np.where(rng.random(n) < p_n, ord("N"), s)
            s = s.astype(np.uint8).tobytes().decode()
            seqs.append(s)
            fh.write(">rec%d some description%s" % (i, eol))
            if comments and i % 3 == 0:
                fh.write(";a comment line%s" % eol)
            w = width if width else max(1, n)
            for j in range(0, n, w):
                fh.write(s[j:j + w] + eol)
    return seqs


def write_fastq(path, rng, n_records, max_len):
    seqs = []
    with open(path, "w") as fh:
        for i in range(n_records):
            n = int(rng.integers(1, max_len))
            s = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[.245, .245, .245, .245, .02]).tobytes().decode()
            q = rng.choice(np.frombuffer(b"@+>I5#", np.uint8), size=n).tobytes().decode()  # '@', '+', '>' in qualities
            seqs.append(s)
            fh.write("@read%d\n%s\n+\n%s\n" % (i, s, q))
    return seqs


def parse(path, overlap, n_regions, cap_bases, cap_seqs=1 << 16):
    L = lib()
    pieces, records = [], 0
    for region in range(n_regions):
        r = C.c_void_p()
        assert L.btlbf_seqfile_open(str(path).encode(), overlap, n_regions, region, C.byref(r)) == 0, L.btlbf_last_error()
        bases = np.zeros(cap_bases, np.uint8)
        offs = np.zeros(cap_seqs + 1, np.uint64)
        nb, ns, nr, done = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int(0)
        guard = 0
        while not done.value:
            assert L.btlbf_seqfile_next(r, bases.ctypes.data_as(C.c_void_p), cap_bases,
                                        offs.ctypes.data_as(C.POINTER(C.c_uint64)), cap_seqs, C.byref(nb), C.byref(ns),
                                        C.byref(nr), C.byref(done)) == 0, L.btlbf_last_error()
            assert offs[0] == 0 and offs[ns.value] == nb.value <= cap_bases
            for i in range(ns.value):
                pieces.append(bases[int(offs[i]):int(offs[i + 1])].tobytes().decode())
            records += nr.value
            guard += 1
            assert guard < 100000
        L.btlbf_seqfile_close(r)
    return pieces, records


def kmers(seqs, k):
    c = collections.Counter()
    for s in seqs:
        for i in range(len(s) - k + 1):
            c[s[i:i + k]] += 1
    return c


@pytest.mark.parametrize("width,crlf,comments", [(60, False, False), (7, True, True), (0, False, False), (1, False, True)])
def test_fasta_pieces_carry_every_window_once(tmp_path, width, crlf, comments):
    rng = np.random.default_rng(width + 17)
    path = tmp_path / "a.fa"
    seqs = write_fasta(path, rng, 40, 900, width, crlf, comments)
    k = 11
    want = kmers(seqs, k)
    for n_regions, cap in ((1, 1 << 20), (1, 257), (3, 4096), (7, 300), (16, 64)):
        pieces, records = parse(path, k - 1, n_regions, cap)
        assert records == len(seqs), (n_regions, cap)
        assert kmers(pieces, k) == want, (n_regions, cap)
        assert all(len(p) <= cap for p in pieces)
    pieces, _ = parse(path, 0, 1, 1 << 20)
    assert pieces == [s for s in seqs if s]  # whole records when nothing has to be cut


def test_fasta_one_long_sequence_across_regions_and_batches(tmp_path):
    rng = np.random.default_rng(5)
    path = tmp_path / "chr.fa"
    seqs = write_fasta(path, rng, 1, 1, 80)  # header-only record first
    with open(path, "a") as fh:
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=200_000).tobytes().decode()
        fh.write(">chr\n")
        for j in range(0, len(s), 80):
            fh.write(s[j:j + 80] + "\n")
    k = 25
    want = kmers(seqs + [s], k)
    for n_regions, cap in ((1, 1 << 20), (5, 30_000), (8, 1 << 20)):
        pieces, records = parse(path, k - 1, n_regions, cap)
        assert records == 2
        assert kmers(pieces, k) == want


def test_fastq_records(tmp_path):
    rng = np.random.default_rng(9)
    path = tmp_path / "r.fq"
    seqs = write_fastq(path, rng, 300, 260)
    k = 9
    want = kmers(seqs, k)
    for n_regions, cap in ((1, 1 << 20), (4, 5000), (9, 400)):
        pieces, records = parse(path, k - 1, n_regions, cap)
        assert records == len(seqs)
        assert kmers(pieces, k) == want


def test_bad_files_are_reported(tmp_path):
    L = lib()
    r = C.c_void_p()
    assert L.btlbf_seqfile_open(str(tmp_path / "missing.fa").encode(), 3, 1, 0, C.byref(r)) != 0
    assert b"cannot open" in L.btlbf_last_error()
    p = tmp_path / "x.txt"
    p.write_text("hello\n")
    assert L.btlbf_seqfile_open(str(p).encode(), 3, 1, 0, C.byref(r)) != 0
    assert b"neither FASTA" in L.btlbf_last_error()
    e = tmp_path / "empty.fa"
    e.write_text("")
    pieces, records = parse(e, 3, 2, 100)
    assert pieces == [] and records == 0


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 4])
def test_file_calls_match_batched_calls(tmp_path, threads):
    import btl_bloomfilter_b200 as B
    rng = np.random.default_rng(21)
    fa, fq = tmp_path / "g.fa", tmp_path / "r.fq"
    seqs = write_fasta(fa, rng, 300, 40_000, 70)       # ~6 MB: several parser regions
    reads = write_fastq(fq, rng, 20_000, 300)
    ctx = B.Context(0)
    k, h, bits = 25, 4, 1 << 26
    a, b = B.BloomFilter(bits, h, k, ctx=ctx), B.BloomFilter(bits, h, k, ctx=ctx)
    n_ref = a.insertSeqs(seqs)
    n_seqs, n_kmers = b.insertFile(str(fa), threads)
    assert (n_seqs, n_kmers) == (len(seqs), n_ref)
    assert np.array_equal(a.to_numpy(), b.to_numpy())
    r = a.containsSeqs(reads)
    assert b.queryFile(str(fq), threads) == (len(reads), r.n_kmers, r.n_hits)
    # the counting insert is order-dependent: one reader, file order, same counters as the batched call
    c1, c2 = B.CountingBloomFilter(1 << 20, h, k, 2, ctx=ctx), B.CountingBloomFilter(1 << 20, h, k, 2, ctx=ctx)
    small = seqs[:40]
    sm = tmp_path / "s.fa"
    with open(sm, "w") as fh:
        for i, s in enumerate(small):
            fh.write(">s%d\n%s\n" % (i, s))
    n1 = c1.insertSeqs(small)
    assert c2.insertFile(str(sm), threads) == (len(small), n1)
    assert np.array_equal(c1.to_numpy(), c2.to_numpy())
