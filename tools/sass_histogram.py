#!/usr/bin/env python3
"""Instruction-class histogram of one kernel's SASS.

  static :  tools/sass_histogram.py <object.o|lib.so> <mangled-name substring> [--out profiles/x.md]
            counts the instructions of the function as cuobjdump -sass lists them (every instruction once)
  dynamic:  tools/sass_histogram.py --source-csv <ncu --page source --csv export> [--out ...]
            weights every SASS line by the "# Instructions Executed" (warp-level) column of an
            `ncu --set full --import-source on` capture: where the executed warp-instructions really go

Classes: 64-bit integer multiply-add (IMAD*/IMAD.WIDE: the ntHash multipliers and the exact modulo), other integer ALU
(IADD3, LOP3, SHF, LEA, SEL, ISETP, ...), shared-memory loads / stores / atomics, global and local memory, barriers and
control flow, uniform-datapath instructions, moves."""
import collections
import csv
import re
import subprocess
import sys

CLASSES = [
    ("imad", r"^(IMAD|UIMAD)"),
    ("int_alu", r"^(IADD|IADD3|VIADD|LOP3|LOP|SHF|SHL|SHR|LEA|SEL|ISETP|IABS|POPC|FLO|PRMT|BMSK|SGXT|VABSDIFF|IMNMX|VIMNMX|PLOP3|P2R|R2P|BREV|LOP32I)"),
    ("smem_atomic", r"^ATOMS"),
    ("smem_ld", r"^LDS"),
    ("smem_st", r"^STS"),
    ("global_ld", r"^(LDG|LD\.|LDGMC)"),
    ("global_st", r"^(STG|ST\.|RED|ATOMG|ATOM)"),
    ("local_spill", r"^(LDL|STL)"),
    ("const_ld", r"^(LDC|LDCU|ULDC)"),
    ("barrier_sync", r"^(BAR|BSSY|BSYNC|WARPSYNC|DEPBAR|MEMBAR|ERRBAR|NANOSLEEP|YIELD)"),
    ("branch", r"^(BRA|BRX|JMP|EXIT|RET|CALL|BREAK|KILL)"),
    ("shuffle_vote", r"^(SHFL|VOTE|VOTEU|MATCH|REDUX)"),
    ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR)"),
    ("move", r"^(MOV|S2R|CS2R|UMOV)"),
]


def classify(op):
    for name, pat in CLASSES:
        if re.match(pat, op):
            return name
    return "other"


def static_hist(obj, needle):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    func, take, rows = None, False, []
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            take = needle in m.group(1)
            if take:
                func = m.group(1)
            continue
        if take:
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                ins = m.group(2).strip()
                ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
                rows.append((int(m.group(1), 16), ins.split()[0], 1))
    return func, rows


def dynamic_hist(path):
    rows = []
    with open(path, newline="") as fh:
        rd = csv.reader(fh)
        hdr = next(rd)
        idx = {h: i for i, h in enumerate(hdr)}
        src = idx.get("Source")
        cnt = idx.get("# Instructions Executed", idx.get("Instructions Executed"))
        adr = idx.get("Address")
        for r in rd:
            if len(r) <= max(src, cnt):
                continue
            ins = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip())
            if not ins:
                continue
            try:
                n = int(float(r[cnt]))
            except ValueError:
                continue
            a = int(r[adr], 16) if adr is not None and r[adr] else len(rows)
            rows.append((a, ins.split()[0], n))
    return "(ncu source page)", rows


def report(func, rows, unit, per=None):
    tot = sum(n for _, _, n in rows)
    by = collections.Counter()
    ops = collections.Counter()
    for _, op, n in rows:
        by[classify(op)] += n
        ops[op.split(".")[0]] += n
    lines = ["kernel: %s" % func, "%s: %d" % (unit, tot), ""]
    if per:
        lines.append("per k-mer (%d k-mers in the launch): %.2f %s" % (per, tot / per, unit))
        lines.append("")
    lines.append("| class | %s | share |" % unit)
    lines.append("|---|---|---|")
    for name, n in by.most_common():
        lines.append("| %s | %d | %.1f %% |" % (name, n, 100.0 * n / max(1, tot)))
    lines.append("")
    lines.append("top opcodes: " + ", ".join("%s %.1f %%" % (o, 100.0 * n / max(1, tot)) for o, n in ops.most_common(14)))
    return "\n".join(lines)


def main():
    a = sys.argv[1:]
    out = None
    per = None
    if "--out" in a:
        i = a.index("--out")
        out = a[i + 1]
        a = a[:i] + a[i + 2:]
    if "--kmers" in a:
        i = a.index("--kmers")
        per = int(a[i + 1])
        a = a[:i] + a[i + 2:]
    if a[0] == "--source-csv":
        func, rows = dynamic_hist(a[1])
        txt = report(func, rows, "executed warp-instructions", per)
    else:
        func, rows = static_hist(a[0], a[1])
        txt = report(func, rows, "static instructions")
    print(txt)
    if out:
        with open(out, "a") as fh:
            fh.write(txt + "\n\n")


if __name__ == "__main__":
    main()
