#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions with stall reasons, and the
executed-instruction mix weighted by execution count.  Usage: ncu_src.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, name):
    try: return float(r[col[name]])
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_samples = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
print("total samples %d, total warp-instr executed %.3e" % (tot_samples, tot_inst))
mix = collections.Counter(); smp = collections.Counter()
for r in data:
    src = r[col["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    mix[op] += f(r, "Instructions Executed"); smp[op] += f(r, "# Samples")
print("executed mix (share of warp-instr | share of stall samples):")
for op, n in mix.most_common(28):
    print("  %-10s %6.2f%%   %6.2f%%" % (op, 100 * n / tot_inst, 100 * smp[op] / max(tot_samples, 1)))
tot = collections.Counter()
for r in data:
    for s in stalls: tot[s] += f(r, s)
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100 * v / max(tot_samples, 1)) for k, v in tot.most_common(10)))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    ss = sorted(((f(r, s), s[6:]) for s in stalls), reverse=True)[:3]
    print("  %6d %5.2f%%  %-70s %s" % (f(r, "# Samples"), 100 * f(r, "# Samples") / tot_samples, r[col["Source"]].strip()[:70],
                                    " ".join("%s:%d" % (n, v) for v, n in ss if v > 0)))
