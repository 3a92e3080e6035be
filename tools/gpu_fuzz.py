"""Randomised differential run on a GPU: random filter shapes, hashing parameters, batch shapes and context
options (partitioned / direct paths, accumulation budget, adaptive query, two-level pass 2, legacy bin kernels,
ordered-update table sizes) against the oracle.  usage: tools/gpu_fuzz.py <seconds> [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _oracle as O, parity_suite as S
from _backends import GpuBackend

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
orc = O.Oracle()
t_end = time.time() + budget
it = 0
while time.time() < t_end:
    seed = seed0 * 100003 + it
    rng = np.random.default_rng(seed)
    it += 1
    k = int(rng.integers(4, 41)); h = int(rng.integers(1, 9))
    counting = rng.random() < 0.35
    opts = {}
    if counting:
        m = int(rng.integers(32, 200_000)) // 8 * 8 + 8
        opts = dict(chunk=int(rng.choice([4096, 8192, 1 << 20])), batch=int(rng.choice([4096, 8192, 1 << 16])),
                    resv_log2=int(rng.integers(10, 20)), list_log2=int(rng.integers(6, 16)),
                    ungrouped_commit=int(rng.random() < 0.3), ordered_coop=int(rng.random() < 0.8))
    else:
        m = int(rng.integers(64, 1 << 22)) // 8 * 8 + 8
        if rng.random() < 0.75:
            opts = dict(bin_shift=int(rng.integers(8, 21)), bin_kernel=int(rng.random() < 0.25),
                        bin_accum_bytes=int(rng.choice([0, 1 << 12, 1 << 20, 1 << 33])),
                        bin_two_level=int(rng.random() < 0.4), bin_two_level_min=0,
                        bin_slack_pct=int(rng.choice([0, 20, 100])),
                        query_adaptive=int(rng.random() < 0.7), query_adaptive_pct=int(rng.choice([0, 20, 50, 101])),
                        query_adaptive_min_tiles=1, chunk=int(rng.choice([4096, 1 << 16, 1 << 20])))
    if counting and rng.random() < 0.5:  # the partitioned threshold query of counting filters
        opts.update(bin_query_mode=1, bin_part_log2=int(rng.integers(11, 21)), query_adaptive=int(rng.random() < 0.5),
                    query_adaptive_min_tiles=1)
    packed = rng.random() < 0.4  # insert / contains through btlbf_pack_seqs + the 2-bit packed entry points
    be = GpuBackend(packed=bool(packed), **opts)
    seeds, h2 = None, 1
    if rng.random() < 0.3:  # spaced seeds (stHashIterator): masks with cared-for ends, h = n_seeds * h2
        n_seeds, h2 = int(rng.integers(1, 5)), int(rng.integers(1, 3))
        seeds = []
        for _ in range(n_seeds):
            mk = (rng.random(k) < 0.7)
            mk[0] = mk[-1] = True
            seeds.append("".join("1" if x else "0" for x in mk))
        h = n_seeds * h2
    thr = int(rng.integers(1, 4))
    f = be.filter(1 if counting else 0, m, h, k, thr=thr, seeds=seeds, h2=h2)
    if seeds:
        ins = (lambda b, off: orc.st_cbf_insert_seqs(ref, m, seeds, h2, k, b, off)) if counting else \
              (lambda b, off: orc.st_bf_insert_seqs(ref, m, seeds, h2, k, b, off))
    ref = np.zeros(m if counting else m // 8, np.uint8)
    try:
        for rnd in range(int(rng.integers(2, 6))):
            b, off = S.rand_batch(rng, int(rng.integers(1, 40)), int(rng.integers(k + 1, 6000)),
                                  p_n=float(rng.choice([0, 0.002, 0.02])), exotic=0.0 if packed else float(rng.choice([0, 0, 0.001])))
            if rng.random() < 0.3:  # repetitive input: skew, dependency chains
                rep = np.frombuffer((b"ACGT" * 2000 + b"A" * 3000), np.uint8)
                b = np.concatenate([b, rep]); off = np.concatenate([off, [off[-1] + rep.size]]).astype(np.uint64)
            if seeds:
                assert f.insert((b, off)) == ins(b, off), "spaced insert count"
                if counting:
                    e = orc.st_cbf_mincount_seqs(ref, m, seeds, h2, k, b, off); g = f.mincount((b, off))
                    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]), "spaced mincount"
                else:
                    q, qo = (b, off) if rng.random() < 0.5 else S.rand_batch(rng, 20, 3000)
                    e = orc.st_bf_contains_seqs(ref, m, seeds, h2, k, q, qo); g = f.contains((q, qo))
                    assert e[:2] == g[:2] and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3]), "spaced contains"
            elif counting:
                assert f.insert((b, off)) == orc.cbf_insert_seqs(ref, m, h, k, b, off)
                if rng.random() < 0.5:
                    assert np.array_equal(f.bytes(), ref), "counters"
                e = orc.cbf_mincount_seqs(ref, m, h, k, b, off); g = f.mincount((b, off))
                assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]), "mincount"
                e = orc.cbf_contains_seqs(ref, m, h, k, thr, b, off); g = f.contains((b, off))
                assert e[:2] == g[:2] and np.array_equal(e[2], g[2]), "counting contains"
            else:
                if rng.random() < 0.25:
                    e = orc.bf_insert_and_check_seqs(ref, m, h, k, b, off); g = f.insert_and_check((b, off))
                    assert e[0] == g[0] and np.array_equal(e[1], g[1]) and np.array_equal(e[2], g[2]), "insert_and_check"
                else:
                    assert f.insert((b, off)) == orc.bf_insert_seqs(ref, m, h, k, b, off), "insert count"
                if rng.random() < 0.6:
                    q, qo = (b, off) if rng.random() < 0.5 else S.rand_batch(rng, 20, 3000)
                    e = orc.bf_contains_seqs(ref, m, h, k, q, qo); g = f.contains((q, qo))
                    assert e[:2] == g[:2] and np.array_equal(e[2], g[2]) and np.array_equal(e[3], g[3]), "contains"
        assert np.array_equal(f.bytes(), ref), "final array"
    except AssertionError as err:
        print("FUZZ FAILURE seed %d iteration %d: %s  kind=%s m=%d h=%d k=%d packed=%s opts=%s" % (seed, it, err, "cbf" if counting else "bf", m, h, k, packed, opts))
        sys.exit(1)
print("fuzz ok: %d random configurations in %.0f s" % (it, budget))
