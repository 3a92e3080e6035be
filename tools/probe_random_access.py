#!/usr/bin/env python3
"""Random-sector microbenchmark: measured ceilings for random 4-byte gathers and random
atomicOr over arrays of several sizes, at each L2 fetch granularity.  Prints one JSON line per case."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import btl_bloomfilter_b200 as B

ctx = B.Context(0)
n_access = 1 << 28
for gran in (0, 32, 64, 128):
    if gran:
        ctx.set_option("l2_fetch_granularity", gran)
    for size in (64 << 20, 4 << 30, 16 << 30):
        t = torch.zeros(size, dtype=torch.uint8, device="cuda")
        for mode, name in ((0, "load"), (1, "atomicOr")):
            ms = min(ctx.random_access_probe(t.data_ptr(), size, n_access, mode) for _ in range(3))
            print(json.dumps({"l2_fetch": gran or "default", "array_bytes": size, "op": name, "accesses": n_access,
                              "ms": ms, "G_access_s": n_access / ms / 1e6,
                              "GBps_32B_sectors": n_access * 32 * (2 if mode else 1) / ms / 1e6}))
        del t
