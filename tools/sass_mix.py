"""Opcode mix of one kernel from `ncu --page source --csv --print-source sass` output.
usage: ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv; python tools/sass_mix.py x.csv [kernel-index]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
k = kernels[want]
h = {n: i for i, n in enumerate(k["hdr"])}
mix, samp = collections.Counter(), collections.Counter()
tot = 0
for r in k["rows"]:
    src = r[h["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    base = op.split(".")[0]
    if base in ("FFMA2", "FADD2", "FMUL2"): base = "F*2(packed)"
    elif base in ("FFMA", "FADD", "FMUL"): base = "F*(scalar)"
    n = int(r[h["Instructions Executed"]])
    mix[base] += n; tot += n
    samp[base] += int(r[h["# Samples"]])
print(k["name"][:90]); print("total warp instructions", tot)
ssum = sum(samp.values())
for op, n in mix.most_common(24):
    print(f"{op:14s} {n:12d} {100*n/tot:6.2f}%   samples {100*samp[op]/ssum:6.2f}%")
