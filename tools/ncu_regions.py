"""Instruction / sample shares of source-line ranges from an .ncu-rep (dev tool).
usage: python tools/ncu_regions.py rep.ncu-rep file:lo-hi:name [...]"""
import csv, subprocess, sys
rep = sys.argv[1]
regions = []
for spec in sys.argv[2:]:
    f, lo, hi, name = spec.split(":")
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        num = lambda k: int(d[k]) if d.get(k, "").strip().lstrip("-").isdigit() else 0
        lines.append((cur_file, int(r[0]), num("Instructions Executed"), num("# Samples"), num("Thread Instructions Executed")))
ti = sum(l[2] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
acc = {name: [0, 0, 0] for *_, name in regions}
acc["other"] = [0, 0, 0]
for f, ln, ins, smp, thr in lines:
    for rf, lo, hi, name in regions:
        if rf == f and lo <= ln <= hi:
            break
    else:
        name = "other"
    a = acc[name]
    a[0] += ins; a[1] += smp; a[2] += thr
for name, (ins, smp, thr) in acc.items():
    print(f"{name:24s} inst {100*ins/ti:5.1f}%  samples {100*smp/ts:5.1f}%  lanes {thr/max(ins,1):4.1f}")
