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
        rows = [r for r in csv.reader(open(launches)) if len(r) > 12 and r[0] != "ID"]
        with open(os.path.join(ROOT, "profiles", "%s_launch_list.csv" % tag), "w") as fh:
            fh.write("id,kernel,grid,block,gpu__time_duration_ns\n")
            for r in rows:
                fh.write("%s,\"%s\",\"%s\",\"%s\",%s\n" % (r[0], r[4].split("(")[0][:70], r[8], r[7], r[14]))
        # share of the step per kernel
        tot = {}
        for r in rows:
            k = r[4].split("(")[0][:70]
            if "btl::" in k and "synth" not in k:
                tot[k] = tot.get(k, 0.0) + float(r[14])
        s = sum(tot.values())
        with open(os.path.join(ROOT, "profiles", "%s_kernel_shares.md" % tag), "w") as fh:
            fh.write("# share of the timed step per kernel (ncu launch list, cold-cache serialised times)\n\n")
            for k, v in sorted(tot.items(), key=lambda x: -x[1]):
                fh.write("- %-60s %8.3f ms total  %5.1f %%\n" % (k, v / 1e6, 100 * v / s))
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
