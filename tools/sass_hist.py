#!/usr/bin/env python
"""Per-function SASS instruction histogram of a .so/.cubin (cuobjdump -sass).  Usage: sass_hist.py lib.so [substr]"""
import collections, re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ""
cur = None; funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur is not None:
        funcs[cur].append(m.group(2).strip())
for name, ins in funcs.items():
    if want not in name: continue
    h = collections.Counter()
    for i in ins:
        op = i.split()[1] if i.startswith("@") else i.split()[0]
        h[op.split(".")[0]] += 1
    print(f"== {name}: {len(ins)} instr")
    print("   " + ", ".join(f"{k}:{v}" for k, v in h.most_common(30)))
    if "--dump" in sys.argv:
        for i in ins: print("     ", i)
