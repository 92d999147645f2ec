"""SASS-level view of an .ncu-rep (dev tool): per-opcode totals of executed instructions, samples and the dominant stall
reasons; optionally the instruction listing of an address window.
usage: python tools/ncu_sass.py rep.ncu-rep [list lo hi]"""
import csv, subprocess, sys, re
from collections import defaultdict
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_smp = sum(int(r[ix["# Samples"]]) for r in body)
if len(sys.argv) > 2 and sys.argv[2] == "list":
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for k, r in enumerate(body[lo:hi]):
        st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:3]
        print(f"{lo+k:5d} {r[ix['Source']].strip()[:70]:70s} ex {int(r[ix['Instructions Executed']]):9d} thr {r[ix['Avg. Threads Executed']]:>5s} smp {int(r[ix['# Samples']]):5d} " +
              " ".join(f"{n}:{c}" for c, n in st if c))
    sys.exit()
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for r in body:
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)?)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("F2F", "MUFU", "LDS", "LDG", "STS", "LD") else "")
    a = agg[op]
    a[0] += int(r[ix["Instructions Executed"]]); a[1] += int(r[ix["# Samples"]])
    for s in stalls:
        a[2][s[6:]] += int(r[ix[s]])
print(f"total inst {tot_inst:.4g} samples {tot_smp}")
for op, (ins, smp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{op:12s} inst {100*ins/tot_inst:5.1f}%  samples {100*smp/tot_smp:5.1f}%  " + " ".join(f"{n}:{100*c/tot_smp:.1f}" for n, c in top if c))
