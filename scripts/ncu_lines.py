"""Aggregate an ncu source-page CSV by CUDA source line: stall samples + instructions.  usage: ncu_lines.py rep [topN] [kernel-regex]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3: cmd += ["--kernel-name", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
iline, isrc = 0, 1
isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed")
agg = collections.OrderedDict()
src = {}
cur = None
tot_s = tot_i = 0
for r in rows[hi + 1:]:
    if len(r) <= iinst: continue
    if r[iline].strip():
        cur = r[iline]; src.setdefault(cur, r[isrc])
    try:
        s_, i_ = int(r[isamp] or 0), int(r[iinst] or 0)
    except ValueError:
        continue
    if cur is None: continue
    a = agg.setdefault(cur, [0, 0]); a[0] += s_; a[1] += i_
    tot_s += s_; tot_i += i_
print("total samples", tot_s, "instructions", tot_i)
for k, (s_, i_) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k:>6s} samples {s_:7d} ({100*s_/max(1,tot_s):5.1f}%) inst {i_:10d}  {src[k].strip()[:110]}")
