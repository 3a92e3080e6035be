"""Parity checks shared by the CPU emulation run (tests/test_emu_kernels.py) and the GPU run
(tests/test_gpu_parity.py): a backend (tests/_backends.py) against the oracle and the golden vectors
generated from the reference.  Bit-exact everywhere (integer / byte / index work)."""
import hashlib

import numpy as np

import _oracle as O


def _hx(row):
    return [format(int(x), "016x") for x in row]


def rand_batch(rng, n_seqs, max_len, p_n=0.01, p_low=0.05, min_len=0, exotic=0.0):
    seqs = []
    for _ in range(n_seqs):
        n = int(rng.integers(min_len, max_len))
        s = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
        s = np.where(rng.random(n) < p_low, s | 0x20, s)
        s = np.where(rng.random(n) < p_n, ord("N"), s)
        if exotic:
            s = np.where(rng.random(n) < exotic, rng.integers(0, 256, n), s)
        seqs.append(s.astype(np.uint8).tobytes())
    return O.as_batch(seqs)


# ------------------------------------------------------------------ golden vectors (reference outputs)
def check_golden_hashes(be, golden):
    for c in golden["hash_cases"]:
        n, hs, _, valid = be.hash([c["seq"]], c["h"], c["k"])
        assert n == c["n"], c["seq"]
        pos = np.nonzero(O.bits_to_bool(valid, len(c["seq"])))[0].tolist()
        assert pos == c["pos"], (c["seq"], c["k"])
        assert [_hx(hs[p]) for p in pos] == c["hashes"]
        mask = np.ones(len(c["seq"]), bool)
        mask[pos] = False
        assert not hs[mask].any()  # invalid windows carry zeros
    for c in golden["st_cases"]:
        n, hs, st, valid = be.hash([c["seq"]], 0, c["k"], c["seeds"], c["h2"])
        assert n == c["n"]
        pos = np.nonzero(O.bits_to_bool(valid, len(c["seq"])))[0].tolist()
        assert pos == c["pos"]
        assert [_hx(hs[p]) for p in pos] == c["hashes"]
        assert [st[p].tolist() for p in pos] == c["strands"]


def check_golden_bf(be, golden):
    for c in golden["bf_cases"]:
        f = be.filter(0, c["bits"], c["h"], c["k"])
        assert f.insert(c["seqs"]) == c["n_inserted"]
        assert f.bytes().tobytes().hex() == c["filter_hex"]
        nq, nh, hits, valid = f.contains(c["queries"])
        assert (nq, nh) == (c["n_queried"], c["n_hits"])
        assert hits.tobytes().hex() == c["hit_hex"] and valid.tobytes().hex() == c["valid_hex"]
    c = golden["insert_and_check"]
    f = be.filter(0, c["bits"], c["h"], c["k"])
    n, found, valid = f.insert_and_check(c["seqs"])
    assert n == c["n"]
    assert found.tobytes().hex() == c["found_hex"] and valid.tobytes().hex() == c["valid_hex"]
    assert f.bytes().tobytes().hex() == c["filter_hex"]
    c = golden["st_bf"]
    f = be.filter(0, c["bits"], len(c["seeds"]) * c["h2"], c["k"], seeds=c["seeds"], h2=c["h2"])
    assert f.insert(c["seqs"]) == c["n"]
    assert f.bytes().tobytes().hex() == c["filter_hex"]
    nq, nh, hits, valid = f.contains(c["queries"])
    assert (nq, nh) == (c["n_queried"], c["n_hits"])
    assert hits.tobytes().hex() == c["hit_hex"] and valid.tobytes().hex() == c["valid_hex"]


def check_golden_cbf(be, golden):
    for c in golden["cbf_cases"]:
        f = be.filter(1, c["size_rounded"], c["h"], c["k"], thr=c["threshold"])
        assert f.insert(c["seqs"]) == c["n_inserted"]
        data = f.bytes()
        assert hashlib.sha256(data.tobytes()).hexdigest() == c["counters_sha256"], c["seqs"][0][:20]
        nq, counts, valid = f.mincount(c["queries"])
        assert nq == c["n_queried"]
        assert counts.tobytes().hex() == c["counts_hex"] and valid.tobytes().hex() == c["valid_hex"]
        _, nh, hits, _ = f.contains(c["queries"])
        assert nh == c["n_hits"] and hits.tobytes().hex() == c["hit_hex"]
    c = golden["increment_all"]
    f = be.filter(1, c["size"], c["h"], c["k"])
    assert f.increment_all(c["seqs"]) == c["n"]
    assert f.bytes().tobytes().hex() == c["counters_hex"]


def check_cfg1(be, oracle, golden):
    """BASELINE.json configs[0]: 1 Mbp synthetic genome, k=25, 4 hashes, 8 Mbit filter."""
    c = golden["cfg1"]
    g = oracle.synth_genome(0, c["genome_len"], c["genome_seed"])
    off = np.array([0, g.size], np.uint64)
    f = be.filter(0, c["bits"], c["h"], c["k"])
    assert f.insert((g, off)) == c["n_inserted"]
    assert hashlib.sha256(f.bytes().tobytes()).hexdigest() == c["filter_sha256"]
    r = c["reads"]
    reads = oracle.synth_reads(0, r["n"], r["len"], g.size, c["genome_seed"], r["seed"])
    roff = np.arange(r["n"] + 1, dtype=np.uint64) * r["len"]
    nq, nh, hits, _ = f.contains((reads, roff))
    assert (nq, nh) == (r["n_queried"], r["n_hits"])
    assert hashlib.sha256(hits.tobytes()).hexdigest() == r["hit_sha256"]
    m = c["miss"]
    miss = oracle.synth_genome(0, m["len"], m["seed"])
    nq, nh, hits, _ = f.contains((miss, np.array([0, miss.size], np.uint64)))
    assert (nq, nh) == (m["n_queried"], m["n_hits"])
    assert hashlib.sha256(hits.tobytes()).hexdigest() == m["hit_sha256"]


# ------------------------------------------------------------------ randomised, against the oracle
def check_random_hashes(be, oracle, k, h, seed, n_seqs=30, max_len=400, exotic=0.0):
    rng = np.random.default_rng(seed)
    b, off = rand_batch(rng, n_seqs, max_len, exotic=exotic)
    n1, h1, v1 = oracle.hash_seqs(h, k, b, off)
    n2, h2, _, v2 = be.hash((b, off), h, k)
    assert n1 == n2
    assert np.array_equal(v1, v2)
    assert np.array_equal(h1, h2)


def check_random_spaced(be, oracle, k, n_seeds, h2, seed):
    rng = np.random.default_rng(seed)
    seeds = ["".join(rng.choice(["0", "1"], size=k, p=[0.3, 0.7]).tolist()) for _ in range(n_seeds)]
    b, off = rand_batch(rng, 20, 300)
    n1, h1, s1, v1 = oracle.st_hash_seqs(seeds, h2, k, b, off)
    n2, hh, s2, v2 = be.hash((b, off), 0, k, seeds, h2)
    assert n1 == n2 and np.array_equal(v1, v2) and np.array_equal(h1, hh) and np.array_equal(s1, s2)
    H = n_seeds * h2
    bits = 8 * int(rng.integers(100, 4000))
    f = be.filter(0, bits, H, k, seeds=seeds, h2=h2)
    filt = np.zeros(bits // 8, np.uint8)
    assert f.insert((b, off)) == oracle.st_bf_insert_seqs(filt, bits, seeds, h2, k, b, off)
    assert np.array_equal(f.bytes(), filt)
    qb, qo = rand_batch(rng, 10, 300)
    e = oracle.st_bf_contains_seqs(filt, bits, seeds, h2, k, qb, qo)
    g = f.contains((qb, qo))
    assert e[:2] == g[:2] and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])
    m = 8 * int(rng.integers(30, 600))
    cf = be.filter(1, m, H, k, thr=2, seeds=seeds, h2=h2)
    cnt = np.zeros(m, np.uint8)
    assert cf.insert((b, off)) == oracle.st_cbf_insert_seqs(cnt, m, seeds, h2, k, b, off)
    assert np.array_equal(cf.bytes(), cnt)
    e = oracle.st_cbf_mincount_seqs(cnt, m, seeds, h2, k, qb, qo)
    g = cf.mincount((qb, qo))
    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2])


def check_random_bf(be, oracle, k, h, bits, seed, n_seqs=40, max_len=300, p_n=0.01):
    rng = np.random.default_rng(seed)
    b, off = rand_batch(rng, n_seqs, max_len, p_n=p_n)
    f = be.filter(0, bits, h, k)
    filt = np.zeros(bits // 8, np.uint8)
    assert f.insert((b, off)) == oracle.bf_insert_seqs(filt, bits, h, k, b, off)
    assert np.array_equal(f.bytes(), filt)
    # queries: half re-used sequences (hits), half fresh (mostly misses)
    qb, qo = rand_batch(rng, n_seqs // 2 + 1, max_len, p_n=p_n)
    for (x, xo) in ((b, off), (qb, qo)):
        e = oracle.bf_contains_seqs(filt, bits, h, k, x, xo)
        g = f.contains((x, xo))
        assert e[:2] == g[:2]
        assert np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])
    # insertAndCheck on top of the existing content, with duplicates inside the batch
    db, do = O.as_batch([qb[int(qo[i]):int(qo[i + 1])].tobytes() for i in range(min(4, qo.size - 1))] * 2)
    e = oracle.bf_insert_and_check_seqs(filt, bits, h, k, db, do)
    g = f.insert_and_check((db, do))
    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2])
    assert np.array_equal(f.bytes(), filt)


def check_random_cbf(be, oracle, k, h, m, seed, n_seqs=40, max_len=300, dup=3, thr=2):
    rng = np.random.default_rng(seed)
    b, off = rand_batch(rng, n_seqs, max_len)
    seqs = [b[int(off[i]):int(off[i + 1])].tobytes() for i in range(off.size - 1)]
    seqs = seqs + seqs[:dup] * 2 + ["A" * (k + 20), "AC" * (k + 5)]  # repeats -> dependency chains
    b, off = O.as_batch(seqs)
    f = be.filter(1, m, h, k, thr=thr)
    cnt = np.zeros(m, np.uint8)
    assert f.insert((b, off)) == oracle.cbf_insert_seqs(cnt, m, h, k, b, off)
    assert np.array_equal(f.bytes(), cnt)
    qb, qo = rand_batch(rng, 10, max_len)
    for (x, xo) in ((b, off), (qb, qo)):
        e = oracle.cbf_mincount_seqs(cnt, m, h, k, x, xo)
        g = f.mincount((x, xo))
        assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2])
        e = oracle.cbf_contains_seqs(cnt, m, h, k, thr, x, xo)
        g = f.contains((x, xo))
        assert e[:2] == g[:2] and np.array_equal(e[2], g[2])
    # a second insert on top (state carried across calls), then incrementAll
    assert f.insert((qb, qo)) == oracle.cbf_insert_seqs(cnt, m, h, k, qb, qo)
    assert np.array_equal(f.bytes(), cnt)
    assert f.increment_all((qb, qo)) == oracle.cbf_increment_all_seqs(cnt, m, h, k, qb, qo)
    assert np.array_equal(f.bytes(), cnt)


def check_edge_cases(be, oracle):
    # empty batch, empty sequences, sequences shorter than k, k == len, all-N
    for seqs in ([], [""], ["", "", ""], ["ACG"], ["ACGT"], ["NNNNNNNNNN"], ["ACGT", "", "ACGTA", "N", "acgtu"]):
        b, off = O.as_batch(seqs)
        f = be.filter(0, 256, 3, 4)
        filt = np.zeros(32, np.uint8)
        assert f.insert(seqs) == oracle.bf_insert_seqs(filt, 256, 3, 4, b, off)
        assert np.array_equal(f.bytes(), filt)
        e = oracle.bf_contains_seqs(filt, 256, 3, 4, b, off)
        g = f.contains(seqs)
        assert e[:2] == g[:2] and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3])
    # windows must not span sequence boundaries: many reads of exactly k and k+1 bases
    rng = np.random.default_rng(3)
    for k in (5, 32):
        seqs = ["".join(rng.choice(list("ACGT"), size=k + int(rng.integers(0, 2))).tolist()) for _ in range(300)]
        b, off = O.as_batch(seqs)
        n1, h1, v1 = oracle.hash_seqs(2, k, b, off)
        n2, h2, _, v2 = be.hash(seqs, 2, k)
        assert n1 == n2 and np.array_equal(v1, v2) and np.array_equal(h1, h2)
    # smallest filters / non-power-of-two moduli / h = 1
    for bits, h in ((8, 1), (8, 5), (24, 2), (1000, 4), (1 << 16, 7)):
        check_random_bf(be, oracle, 7, h, bits, 99 + bits, n_seqs=6, max_len=80)
    for m, h in ((8, 1), (8, 4), (1000, 3)):
        check_random_cbf(be, oracle, 6, h, m, 7 + m, n_seqs=5, max_len=60)
