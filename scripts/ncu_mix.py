"""Instruction mix of one kernel from an ncu report (source page, cuda+sass view): per SASS opcode and per CUDA source line.
usage: ncu_mix.py <report.ncu-rep> [kernel-regex] [topN]"""
import collections
import csv
import os
import re
import subprocess
import sys

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kern:
    cmd += ["--kernel-name", "regex:" + kern]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
op_i, op_s, line_i, line_s, line_src = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), {}
cur, fname, iinst, isamp, tot = None, "?", None, None, 0
seen = set()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = os.path.basename(r[1])
        continue
    if len(r) > 4 and r[0] == "Line No":
        iinst, isamp = r.index("Instructions Executed"), r.index("# Samples")
        continue
    if iinst is None or len(r) <= iinst:
        continue
    if r[0].strip():  # CUDA source line row (aggregated over its SASS)
        cur = f"{fname}:{r[0].strip()}"
        line_src.setdefault(cur, r[1].strip())
        continue
    addr, sass = r[2].strip(), r[3]
    if not addr.startswith("0x") or addr in seen:   # inlined code appears under several files: count each address once
        continue
    seen.add(addr)
    try:
        n, s = int(r[iinst] or 0), int(r[isamp] or 0)
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    full = m.group(2) if m else "?"
    op = full.split(".")[0]
    if op in ("IDP", "IMAD", "SHF", "LDS", "REDUX", "I2F", "F2I", "LDG", "STG"):
        op = ".".join(full.split(".")[:2])
    op_i[op] += n
    op_s[op] += s
    if cur is not None:
        line_i[cur] += n
        line_s[cur] += s
    tot += n
print(f"total warp instructions executed: {tot}")
print("\n== by opcode ==")
for op, n in op_i.most_common(topn):
    print(f"{op:16s} {n:12d} {100.0 * n / tot:6.2f} %   stall samples {op_s[op]}")
print("\n== by CUDA source line ==")
for ln, n in line_i.most_common(topn):
    print(f"{ln:>18s} {n:12d} {100.0 * n / tot:6.2f} %  samples {line_s[ln]:6d}  {line_src[ln][:110]}")
