"""Opcode histogram of one kernel's SASS in windows (dev tool): python tools/sass_hist.py <obj> <substring of the mangled name>"""
import re, subprocess, sys
from collections import Counter
obj, pat = sys.argv[1], sys.argv[2]
W = int(sys.argv[3]) if len(sys.argv) > 3 else 80
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, ops = None, []
for l in names.splitlines():
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(3)))
print(len(ops), "instructions")
tot = Counter(o for _, o in ops)
print(dict(tot.most_common(25)))
for i in range(0, len(ops), W):
    c = Counter(o for _, o in ops[i:i + W])
    print(i, hex(ops[i][0]), dict(c.most_common(7)))
