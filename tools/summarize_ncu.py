#!/usr/bin/env python3
"""Turn ncu captures (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text summaries kept under
profiles/.  Usage: tools/summarize_ncu.py <round tag> <full.ncu-rep>... [--launches launches.csv]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def raw_rows(rep):
    if rep.endswith(".csv"):  # already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv)
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    tag = sys.argv[1]
    args = sys.argv[2:]
    launches = None
    if "--launches" in args:
        i = args.index("--launches")
        launches = args[i + 1]
        args = args[:i] + args[i + 2:]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lines = ["# ncu --set full --clock-control none summaries (%s)" % tag, ""]
    traffic = {}
    for rep in args:
        hdr, units, rows = raw_rows(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        lines.append("## %s" % os.path.basename(rep))
        for r in rows:
            name = r[idx["Kernel Name"]]
            lines.append("")
            lines.append("### %s" % name)
            for m in METRICS:
                if m in idx:
                    lines.append("- %s = %s %s" % (m, r[idx[m]], units[idx[m]]))
            try:
                rd = float(r[idx["dram__bytes_read.sum"]])
                wr = float(r[idx["dram__bytes_write.sum"]])
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
                traffic.setdefault(name.split("(")[0].strip(), []).append(tot)
            except Exception:
                pass
        lines.append("")
    with open(os.path.join(ROOT, "profiles", "%s_ncu_full_summary.md" % tag), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    tj = {k: sum(v) / len(v) for k, v in traffic.items()}
    with open(os.path.join(ROOT, "profiles", "%s_dram_bytes_per_launch.json" % tag), "w") as fh:
        json.dump(tj, fh, indent=1)
    if launches:
        # one row per (launch, metric): ID, ..., kernel name [4], block [7], grid [8], metric [12], unit [13], value [14]
        rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0] != "ID"]
        per = {}
        order = []
        for r in rows:
            i = int(r[0])
            if i not in per:
                per[i] = {"kernel": r[4].split("(")[0][:70], "grid": r[8], "block": r[7]}
                order.append(i)
            per[i][r[12]] = float(r[14].replace(",", ""))
        with open(os.path.join(ROOT, "profiles", "%s_launch_list.csv" % tag), "w") as fh:
            fh.write("id,kernel,grid,block,gpu__time_duration_ns,dram_bytes_read,dram_bytes_write\n")
            for i in order:
                d = per[i]
                fh.write("%d,\"%s\",\"%s\",\"%s\",%.0f,%.0f,%.0f\n" % (
                    i, d["kernel"], d["grid"], d["block"], d.get("gpu__time_duration.sum", 0),
                    d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)))
        # the timed region of `bench.py --steps S --warmup W`: per kernel, the launches after the warm-up ones
        W = int(os.environ.get("BENCH_WARMUP", "3"))
        S = int(os.environ.get("BENCH_STEPS", "8"))
        seen = {}
        tot, dram = {}, {}
        for i in order:
            d = per[i]
            k = d["kernel"]
            if "btl::" not in k or "synth" in k:
                continue
            n = seen.get(k, 0)
            seen[k] = n + 1
            # warm-up launches per kernel: the warm-up builds are settled by one pass 2; every query launches the
            # early-exit kernel twice (strided sample + gated whole-batch form)
            warm = 1 if "apply_bins" in k else 2 * W if "seq_kernel<2" in k else W
            # ... and launches after the timed region (the miss-set measurement) do not belong to the step either
            timed = 10 ** 9 if "apply_bins" in k else 2 * S if "seq_kernel<2" in k else S
            if n < warm or n >= warm + timed:
                continue
            tot[k] = tot.get(k, 0.0) + d.get("gpu__time_duration.sum", 0)
            dram[k] = dram.get(k, 0.0) + d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        s = sum(tot.values())
        with open(os.path.join(ROOT, "profiles", "%s_kernel_shares.md" % tag), "w") as fh:
            fh.write("# share of the timed region per kernel (ncu launch list of `bench.py --steps %d --warmup %d`, "
                     "cold-cache serialised times)\n\n" % (S, W))
            for k, v in sorted(tot.items(), key=lambda x: -x[1]):
                fh.write("- %-60s %8.3f ms total  %5.1f %%   DRAM %7.2f GB\n" % (k, v / 1e6, 100 * v / s, dram[k] / 1e9))
        build = sum(v for k, v in dram.items() if "apply_bins" in k or ("bin_kernel" in k and ", 0>" in k[-6:]))
        query = sum(v for k, v in dram.items() if "probe_bins" in k or "finalize" in k or "seq_kernel<2" in k or
                    "query_gate" in k or ("bin_kernel" in k and ", 1>" in k[-6:]))
        with open(os.path.join(ROOT, "profiles", "%s_dram_bytes_per_step.json" % tag), "w") as fh:
            # the hash of the kernel sources the capture was taken from (recorded on the GPU box by
            # tools/gpu_profile_round.sh next to the launch list; bench.py attaches `traffic` only when it still matches)
            sha = None
            sha_file = os.path.join(os.path.dirname(os.path.abspath(launches)), "source_sha_%s.txt" % os.environ.get("RUN_TAG", tag))
            if os.path.exists(sha_file):
                sha = open(sha_file).read().strip()
            json.dump({"build_bytes_per_step": build / S, "query_bytes_per_step": query / S, "steps": S,
                       "source_sha": sha,
                       "source": "ncu launch list, dram__bytes_read.sum + dram__bytes_write.sum of the timed launches"},
                      fh, indent=1)
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
