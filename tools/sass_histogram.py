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
    """rows of an `ncu --page source --csv` export: (address, opcode, executed warp-instructions, stall samples)"""
    allrows = list(csv.reader(open(path, newline="")))
    hi = [i for i, r in enumerate(allrows) if r and r[0] == "Address"][0]
    name = allrows[hi - 1][1] if hi > 0 and len(allrows[hi - 1]) > 1 else "(ncu source page)"
    idx = {h: i for i, h in enumerate(allrows[hi])}
    src, cnt, smp = idx["Source"], idx["Instructions Executed"], idx.get("Warp Stall Sampling (All Samples)")
    rows = []
    for r in allrows[hi + 1:]:
        try:
            a, n = int(r[0], 16), int(float(r[cnt]))
        except (ValueError, IndexError):
            continue
        ins = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip())
        if ins:
            rows.append((a, ins.split()[0], n, int(r[smp] or 0) if smp is not None else 0))
    return name, rows


def sections(rows):
    """executed instructions between consecutive CTA barriers (the kernel's phases)"""
    base = rows[0][0]
    tot = sum(r[2] for r in rows)
    out, cur, start = [], 0, rows[0][0]
    smp = 0
    for r in rows:
        cur += r[2]
        smp += r[3] if len(r) > 3 else 0
        if r[1].startswith("BAR"):
            out.append((start - base, r[0] - base, cur, smp))
            cur, smp, start = 0, 0, r[0] + 16
    out.append((start - base, rows[-1][0] - base, cur, smp))
    lines = ["| code range (hex offset) | executed warp-instructions | share | stall samples |", "|---|---|---|---|"]
    for a, b, n, sp in out:
        lines.append("| %x - %x | %d | %.1f %% | %d |" % (a, b, n, 100.0 * n / max(1, tot), sp))
    return "\n".join(lines)


def report(func, rows, unit, per=None):
    tot = sum(r[2] for r in rows)
    by = collections.Counter()
    ops = collections.Counter()
    for r in rows:
        op, n = r[1], r[2]
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
        txt = report(func, rows, "executed warp-instructions", per) + "\n\nbetween CTA barriers:\n\n" + sections(rows)
    else:
        func, rows = static_hist(a[0], a[1])
        txt = report(func, rows, "static instructions")
    print(txt)
    if out:
        with open(out, "a") as fh:
            fh.write(txt + "\n\n")


if __name__ == "__main__":
    main()
