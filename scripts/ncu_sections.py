"""Attribute the executed warp instructions of klt_track_kernel to algorithm sections (ncu source page, cuda+sass view).
Inlined helpers (dp2a wrappers, run_bytes, warp_sum_exact, bilinear_weights, mbarrier / TMA wrappers) are charged to the section of the
surrounding code: SASS instructions are walked in address order and a helper instruction inherits the section of the last non-helper line.
usage: ncu_sections.py <report.ncu-rep> <sections.txt>   (sections.txt: "<first line> <last line> <name>" per row, for klt.cu)"""
import collections
import csv
import os
import re
import subprocess
import sys

rep, secfile = sys.argv[1], sys.argv[2]
sections = []
for ln in open(secfile):
    ln = ln.strip()
    if ln and not ln.startswith("#"):
        a, b, name = ln.split(None, 2)
        sections.append((int(a), int(b), name))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:klt_track"],
                                      capture_output=True, text=True).stdout.splitlines()))
inst = {}
cur, fname, iinst = None, "?", None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if len(r) > 4 and r[0] == "Line No":
        iinst = r.index("Instructions Executed")
        continue
    if iinst is None or len(r) <= iinst:
        continue
    if r[0].strip():
        cur = (fname, int(r[0]))
        continue
    addr = r[2].strip()
    if not addr.startswith("0x"):
        continue
    try:
        n = int(r[iinst] or 0)
    except ValueError:
        continue
    a = int(addr, 16)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
    op = m.group(2) if m else "?"
    # the same address can be listed under several inlined source lines: keep the klt.cu body line if there is one
    old = inst.get(a)
    if old is None or (cur[0] == "klt.cu" and old[1][0] != "klt.cu"):
        inst[a] = (n, cur, op)


def section_of(line):
    for a, b, name in sections:
        if a <= line <= b:
            return name
    return None


tot = sum(v[0] for v in inst.values())
sec_i = collections.Counter()
sec_ops = collections.defaultdict(collections.Counter)
ctx = "prologue / epilogue of the kernel"
for a in sorted(inst):
    n, (f, line), op = inst[a]
    s = section_of(line) if f == "klt.cu" else None
    if s is not None and not s.startswith("helper"):
        ctx = s
    sec_i[ctx] += n
    sec_ops[ctx][op.split(".")[0] + ("." + op.split(".")[1] if op.split(".")[0] in ("IDP", "REDUX", "LDS") and "." in op else "")] += n
print(f"total warp instructions: {tot}")
for s, n in sec_i.most_common():
    top = ", ".join(f"{o} {100.0 * c / n:.0f}%" for o, c in sec_ops[s].most_common(6))
    print(f"{n:12d} {100.0 * n / tot:6.2f} %  {s:55s} [{top}]")
