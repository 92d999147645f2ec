"""Per-source-line instruction / stall-sample totals from an .ncu-rep (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, lines, hdr = None, [], None
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
        lines.append((cur_file, int(r[0]), r[1].strip(), num("Instructions Executed"), num("# Samples"),
                      num("Thread Instructions Executed")))
tot_i = sum(l[3] for l in lines) or 1
tot_s = sum(l[4] for l in lines) or 1
print(f"total warp instructions {tot_i:.4g}, samples {tot_s}")
print("--- by instructions")
for f, ln, src, ins, smp, thr in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{f}:{ln:4d} inst {100 * ins / tot_i:5.1f}%  samples {100 * smp / tot_s:5.1f}%  thr/inst {thr / max(ins, 1):4.1f} | {src[:110]}")
print("--- by samples")
for f, ln, src, ins, smp, thr in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{f}:{ln:4d} inst {100 * ins / tot_i:5.1f}%  samples {100 * smp / tot_s:5.1f}%  thr/inst {thr / max(ins, 1):4.1f} | {src[:110]}")
