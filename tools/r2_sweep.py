#!/usr/bin/env python3
"""Option sweeps on one GPU (device-resident batches, CUDA events): for each config, a list of context-option sets;
prints one JSON line per (config, option set) with build and query rates.  Not part of the product or the tests.

  python tools/r2_sweep.py cfg2 "query_sub=1" "query_sub=4,query_p1_ctas=3" ...
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench as BN  # noqa: E402
import btl_bloomfilter_b200 as B  # noqa: E402


class A:
    stream_priority = 0
    l2_fetch = 0
    opt = []
    chunk = BN.CHUNK
    query_factor = 4
    no_e2e = True


def main():
    name = sys.argv[1]
    sets = sys.argv[2:] or [""]
    cfg = dict(BN.CONFIGS[name])
    if os.environ.get("THRESHOLD"):
        cfg["threshold"] = int(os.environ["THRESHOLD"])
    reps_b = int(os.environ.get("BUILD_REPS", "6"))
    reps_q = int(os.environ.get("QUERY_REPS", "4"))
    env = BN.Env(A, torch, None, B)
    env.make_inputs(reps_b + 1, 3)
    ctx, k = env.ctx, cfg["k"]
    st = torch.zeros(4, dtype=torch.int64, device=env.dev)
    for spec in sets:
        opts = dict(kv.split("=") for kv in spec.split(",") if kv)
        for key, v in opts.items():
            ctx.set_option(key, int(v))
        f = BN.make_filter(env, cfg)

        def build(i):
            n = env.insert_len(i, k)
            f.insertSeqsDevice(env.g[i].data_ptr(), n, env.goff(n).data_ptr(), 1, st.data_ptr())

        def query(i):
            f.containsSeqsDevice(env.r[i % 3].data_ptr(), env.read_bases, env.d_roff.data_ptr(), env.n_reads,
                                 env.d_hits.data_ptr(), 0, st[2:].data_ptr())

        out = {"config": name, "options": opts}
        if os.environ.get("SKIP_BUILD") != "1":
            build(0)
            ctx.flush()
            torch.cuda.synchronize()
            st.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(env.stream)
            for i in range(reps_b):
                build(1 + i)
            ctx.flush()
            b.record(env.stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps_b
            ki = int(st[0]) / reps_b
            out.update({"build_ms": ms, "build_gkmers_s": ki / ms / 1e6,
                        "build_frac": ki / ms / 1e6 * (64 * cfg["h"] + 1) / 6546.6})
        else:
            for i in range(3):
                build(i)
            ctx.flush()
        query(0)
        torch.cuda.synchronize()
        st.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(env.stream)
        for i in range(reps_q):
            query(i)
        b.record(env.stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps_q
        kq = int(st[2]) / reps_q
        out.update({"query_ms": ms, "query_gkmers_s": kq / ms / 1e6, "query_frac": kq / ms / 1e6 * (32 * cfg["h"] + 1) / 6546.6,
                    "hit_fraction": int(st[3]) / max(1, int(st[2]))})
        print(json.dumps(out), flush=True)
        del f
        for key in opts:  # back to the defaults that matter for the next set
            ctx.set_option(key, {"bin_prefetch": -1, "bin_max_parts": 512, "bin_part_log2": 27}.get(key, 0))


if __name__ == "__main__":
    main()
