#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.json from the UNMODIFIED reference (oracle/_ref/libbtlref.so,
built by oracle/Makefile from the headers under /root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):
    python oracle/make_golden.py
The JSON it writes is committed; tests/test_oracle.py pins the plain-C restatement
(oracle/btl_oracle.c) to it, and the GPU parity tests compare the CUDA path with it, so the
fixtures travel to boxes where the reference tree is absent.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O  # noqa: E402


def hx(a):
    return [format(int(x), "016x") for x in a]


def rand_seq(rng, n, p_n=0.0, p_lower=0.0, alphabet="ACGT"):
    s = rng.choice(list(alphabet), size=n)
    if p_lower:
        low = rng.random(n) < p_lower
        s = np.where(low, np.char.lower(s), s)
    if p_n:
        s = np.where(rng.random(n) < p_n, "N", s)
    return "".join(s.tolist())


def main():
    if not O.Ref.available():
        sys.exit("oracle/_ref/libbtlref.so is missing: run `make -C oracle` where /root/reference exists")
    R = O.Ref()
    rng = np.random.default_rng(20261018)
    out = {"generator": "oracle/make_golden.py", "reference": "bcgsc/btl_bloomfilter v1.2.1 headers, g++ -O3"}

    # ---- ntHashIterator outputs
    hash_cases = []
    fixed = [
        ("TAGAATCACCCAAAGA", 5, 4),
        ("ACGTAC", 4, 5),
        ("ACGTACACTGGACTGAGTCT", 8, 5),
        ("GATTACAGATTACAGATTACAGATTNACAGATTACAGATTACAGATTACAGATTACA", 25, 4),
        ("acgtNNacgtu", 4, 2),
        ("ACG", 4, 3),            # shorter than k: nothing
        ("NNNNNNNN", 3, 2),       # nothing valid
        ("AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", 32, 6),
        ("ACGTNACGTRYACGTTTGACnACGTACGTAC", 4, 1),
        ("AC\x01\x03\x04\x05\x07GTACGT\x02ACGTAC", 5, 3),  # raw bytes that seedTab accepts / rejects
    ]
    for k in (1, 2, 3, 4, 5, 25, 31, 32, 33, 64, 100):
        fixed.append((rand_seq(rng, 140, p_n=0.02, p_lower=0.05), k, 1 + (k % 6)))
    for seq, k, h in fixed:
        bases, off = O.as_batch([seq])
        n, hashes, valid = R.hash_seqs(h, k, bases, off)
        v = O.bits_to_bool(valid, bases.size)
        pos = np.nonzero(v)[0]
        hash_cases.append({"seq": seq, "k": k, "h": h, "n": n, "pos": pos.tolist(),
                           "hashes": [hx(hashes[p]) for p in pos]})
    out["hash_cases"] = hash_cases

    # ---- stHashIterator outputs
    st_cases = []
    st_fixed = [
        ("TAGAATCACCCAAAGA", 5, ["11011", "10101"], 2),
        ("TAGAATCACCCAAAGANNTAGGACCAcgtagctagcattgGGATCGATTTAGC", 7, ["1101011", "1011101", "1111111"], 3),
        (rand_seq(rng, 120, p_n=0.02, p_lower=0.05), 31,
         ["1110110111011011101101101110111", "1011101101110110111011101101101"], 1),
        (rand_seq(rng, 90, p_n=0.03), 12, ["110011001100", "001100110011", "101010101010", "111000111000"], 2),
    ]
    for seq, k, seeds, h2 in st_fixed:
        bases, off = O.as_batch([seq])
        n, hashes, strands, valid = R.st_hash_seqs(seeds, h2, k, bases, off)
        v = O.bits_to_bool(valid, bases.size)
        pos = np.nonzero(v)[0]
        st_cases.append({"seq": seq, "k": k, "seeds": seeds, "h2": h2, "n": n, "pos": pos.tolist(),
                         "hashes": [hx(hashes[p]) for p in pos],
                         "strands": [strands[p].tolist() for p in pos]})
    out["st_cases"] = st_cases

    tmp = tempfile.mkdtemp()

    def md5(path):
        return hashlib.md5(open(path, "rb").read()).hexdigest()

    # ---- BloomFilter scenarios (insert -> bytes, pop, file; contains -> hit vector)
    bf_cases = []
    bf_fixed = [
        (["TAGAATCACCCAAAGA"], 1000, 4, 5, ["TAGAATCACCCAAAGA", "GGGGGCCCCCTTTTTAAAAA"]),
        (["ACGTAC"], 1024, 5, 4, ["ACGTAC", "TTTTTT"]),
        ([rand_seq(rng, 150, p_n=0.01) for _ in range(40)], 65536 + 8, 4, 25,
         [rand_seq(rng, 150) for _ in range(10)]),
        ([rand_seq(rng, 97, p_lower=0.1) for _ in range(25)] + ["", "ACGT", "N"], 8 * 1237, 6, 32,
         [rand_seq(rng, 64) for _ in range(6)]),
    ]
    for seqs, bits, h, k, queries in bf_fixed:
        f = R.bf_new(bits, h, k)
        bases, off = O.as_batch(seqs)
        n_ins = R.L.ref_bf_insert_seqs(f, O._p8(bases), O._p64(off), off.size - 1)
        data = R.bf_bytes(f)
        path = os.path.join(tmp, "bf.bf")
        R.L.ref_bf_store(f, path.encode())
        raw = open(path, "rb").read()
        qb, qo = O.as_batch(queries + seqs[:3])
        nq, nh, hits, valid = R.bf_contains_seqs(f, qb, qo)
        bf_cases.append({"seqs": seqs, "bits": bits, "h": h, "k": k, "n_inserted": int(n_ins),
                         "pop": int(R.L.ref_bf_pop(f)), "filter_hex": data.tobytes().hex(),
                         "file_md5": md5(path), "header": raw[: len(raw) - data.size].decode(),
                         "queries": queries + seqs[:3], "n_queried": nq, "n_hits": nh,
                         "hit_hex": hits.tobytes().hex(), "valid_hex": valid.tobytes().hex()})
        R.L.ref_bf_free(f)
    out["bf_cases"] = bf_cases

    # ---- insertAndCheck (order-dependent which duplicate reports "new"; single-threaded reference order)
    seqs = [rand_seq(rng, 80) for _ in range(6)]
    seqs = seqs + seqs[:2]
    f = R.bf_new(8 * 4096, 3, 21)
    bases, off = O.as_batch(seqs)
    found = np.zeros(O.nbits_bytes(bases.size), np.uint8)
    valid = np.zeros(O.nbits_bytes(bases.size), np.uint8)
    n = R.L.ref_bf_insert_and_check_seqs(f, O._p8(bases), O._p64(off), off.size - 1, O._p8(found), O._p8(valid))
    out["insert_and_check"] = {"seqs": seqs, "bits": 8 * 4096, "h": 3, "k": 21, "n": int(n),
                               "found_hex": found.tobytes().hex(), "valid_hex": valid.tobytes().hex(),
                               "filter_hex": R.bf_bytes(f).tobytes().hex()}
    R.L.ref_bf_free(f)

    # ---- CountingBloomFilter<uint8_t> scenarios (single-threaded, read order)
    cbf_cases = []
    dup = rand_seq(rng, 60)
    cbf_fixed = [
        (["TAGAATCACCCAAAGA"], 1000, 4, 5, 1, ["TAGAATCACCCAAAGA"]),
        (["ACGTACACTGGACTGAGTCT"], 100001, 5, 8, 1, ["ACGTACACTGGACTGAGTCT", rand_seq(rng, 60)]),
        ([rand_seq(rng, 150, p_n=0.01) for _ in range(30)] + [dup] * 5, 4099, 4, 25, 2, [dup, rand_seq(rng, 100)]),
        (["ACGTA"] * 300, 64, 3, 5, 200, ["ACGTA"]),  # saturation at 255
        (["A" * 200, "AC" * 100, "ACG" * 70], 512, 4, 11, 3, ["A" * 30, "ACGACGACGACGACG"]),  # tandem repeats
    ]
    for seqs, size, h, k, thr, queries in cbf_fixed:
        f = R.L.ref_cbf_new(size, h, k, thr)
        bases, off = O.as_batch(seqs)
        n_ins = R.L.ref_cbf_insert_seqs(f, O._p8(bases), O._p64(off), off.size - 1)
        data = R.cbf_bytes(f)
        path = os.path.join(tmp, "cbf.bf")
        R.L.ref_cbf_store(f, path.encode())
        raw = open(path, "rb").read()
        qb, qo = O.as_batch(queries)
        counts = np.zeros(qb.size, np.uint8)
        valid = np.zeros(O.nbits_bytes(qb.size), np.uint8)
        nq = R.L.ref_cbf_mincount_seqs(f, O._p8(qb), O._p64(qo), qo.size - 1, O._p8(counts), O._p8(valid))
        hits = np.zeros(O.nbits_bytes(qb.size), np.uint8)
        nh = O.u64(0)
        R.L.ref_cbf_contains_seqs(f, O._p8(qb), O._p64(qo), qo.size - 1, O._p8(hits), None, O.C.byref(nh))
        cbf_cases.append({"seqs": seqs, "size": size, "size_rounded": int(R.L.ref_cbf_size(f)), "h": h, "k": k,
                          "threshold": thr, "n_inserted": int(n_ins), "popcount": int(R.L.ref_cbf_popcount(f)),
                          "filtered_popcount": int(R.L.ref_cbf_filtered_popcount(f)),
                          "counters_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
                          "counters_hex": data.tobytes().hex() if data.size <= 8192 else None,
                          "file_md5": md5(path), "header": raw[: len(raw) - data.size].decode(),
                          "queries": queries, "n_queried": int(nq), "counts_hex": counts.tobytes().hex(),
                          "valid_hex": valid.tobytes().hex(), "hit_hex": hits.tobytes().hex(),
                          "n_hits": int(nh.value)})
        R.L.ref_cbf_free(f)
    out["cbf_cases"] = cbf_cases

    # ---- incrementAll
    seqs = [rand_seq(rng, 70) for _ in range(8)] + ["ACGTACGTACGTACGTACGT"] * 4
    f = R.L.ref_cbf_new(2048, 3, 9, 1)
    bases, off = O.as_batch(seqs)
    n = R.L.ref_cbf_increment_all_seqs(f, O._p8(bases), O._p64(off), off.size - 1)
    out["increment_all"] = {"seqs": seqs, "size": 2048, "h": 3, "k": 9, "n": int(n),
                            "counters_hex": R.cbf_bytes(f).tobytes().hex()}
    R.L.ref_cbf_free(f)

    # ---- spaced seeds into filters
    seqs = [rand_seq(rng, 120, p_n=0.01) for _ in range(12)]
    seeds = ["1110110111011011101101101110111", "1011101101110110111011101101101"]
    f = R.bf_new(8 * 8192, 2, 31)
    bases, off = O.as_batch(seqs)
    sp = R._seeds(seeds)
    n = R.L.ref_st_bf_insert_seqs(f, sp, 2, 1, O._p8(bases), O._p64(off), off.size - 1)
    q = seqs[:2] + [rand_seq(rng, 100)]
    qb, qo = O.as_batch(q)
    hits = np.zeros(O.nbits_bytes(qb.size), np.uint8)
    valid = np.zeros(O.nbits_bytes(qb.size), np.uint8)
    nh = O.u64(0)
    nq = R.L.ref_st_bf_contains_seqs(f, sp, 2, 1, O._p8(qb), O._p64(qo), qo.size - 1, O._p8(hits), O._p8(valid),
                                     O.C.byref(nh))
    out["st_bf"] = {"seqs": seqs, "seeds": seeds, "h2": 1, "bits": 8 * 8192, "k": 31, "n": int(n),
                    "filter_hex": R.bf_bytes(f).tobytes().hex(), "queries": q, "n_queried": int(nq),
                    "n_hits": int(nh.value), "hit_hex": hits.tobytes().hex(), "valid_hex": valid.tobytes().hex()}
    R.L.ref_bf_free(f)

    # ---- KmerBloomFilter::insert/contains(const char*) (table-driven NTC64 + NTE64), k % 4 != 0 only:
    # for k % 4 == 0 the reference's value is undefined behaviour
    kbf = []
    for k, h, bits in ((5, 4, 8 * 997), (25, 4, 8 * 4099), (31, 3, 1 << 15), (22, 6, 8 * 1237)):
        kmers = [rand_seq(rng, k, p_lower=0.2) for _ in range(40)]
        f = R.L.ref_kbf_new(bits, h, k)
        for km in kmers:
            R.L.ref_kbf_insert(f, km.encode())
        data = np.ctypeslib.as_array(R.L.ref_kbf_data(f), shape=(bits // 8,)).copy()
        probes = kmers[:5] + [rand_seq(rng, k) for _ in range(20)]
        kbf.append({"k": k, "h": h, "bits": bits, "kmers": kmers, "filter_hex": data.tobytes().hex(),
                    "probes": probes, "contains": [int(R.L.ref_kbf_contains(f, p_.encode())) for p_ in probes]})
        R.L.ref_kbf_free(f)
    out["kmer_bloom_filter"] = kbf

    # ---- cfg1 (BASELINE.json configs[0]): 1 Mbp synthetic genome (seed 42), k=25, h=4, 2^23-bit filter
    orc = O.Oracle()
    g = orc.synth_genome(0, 1_000_000, 42)
    off = np.array([0, g.size], np.uint64)
    f = R.bf_new(1 << 23, 4, 25)
    n = R.L.ref_bf_insert_seqs(f, O._p8(g), O._p64(off), 1)
    data = R.bf_bytes(f)
    reads = orc.synth_reads(0, 2000, 150, g.size, 42, 7)
    roff = (np.arange(2001, dtype=np.uint64) * 150)
    nq, nh, hits, valid = R.bf_contains_seqs(f, reads, roff)
    # NB: seeds must differ above bit log2(len/32), else seed^(i>>5) only permutes 32-base blocks
    miss = orc.synth_genome(0, 300_000, 43 << 40)
    moff = np.array([0, miss.size], np.uint64)
    mq, mh, mhits, _ = R.bf_contains_seqs(f, miss, moff)
    out["cfg1"] = {"genome_len": 1_000_000, "genome_seed": 42, "k": 25, "h": 4, "bits": 1 << 23,
                   "n_inserted": int(n), "pop": int(R.L.ref_bf_pop(f)),
                   "filter_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
                   "genome_sha256": hashlib.sha256(g.tobytes()).hexdigest(),
                   "reads": {"n": 2000, "len": 150, "seed": 7, "n_queried": nq, "n_hits": nh,
                             "hit_sha256": hashlib.sha256(hits.tobytes()).hexdigest()},
                   "miss": {"len": 300_000, "seed": 43 << 40, "n_queried": mq, "n_hits": mh,
                            "hit_sha256": hashlib.sha256(mhits.tobytes()).hexdigest()}}
    R.L.ref_bf_free(f)

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    dst = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
