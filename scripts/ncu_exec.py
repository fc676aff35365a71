"""Executed-instruction totals of an ncu source csv grouped by CUDA source line, each SASS instruction counted ONCE
(the csv lists an inlined instruction under every line of its inline stack: keep the first = outermost... we keep the
line of its first occurrence).  usage: ncu_exec.py file.csv units [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1]))); units = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
hi = next(i for i, r in enumerate(rows[:20]) if 'Instructions Executed' in r)
hdr = rows[hi]; H = len(hdr); col = {n: i for i, n in enumerate(hdr)}
cur = None; seen = {}; text = {}
for r in rows[hi + 1:]:
    if not r or r[0] == "File Name": continue
    if r[0].strip().isdigit():
        cur = int(r[0]); text.setdefault(cur, ",".join(r[1:len(r) - (H - 2)])[:80]); continue
    if r[0] == "" and len(r) >= H:
        off = len(r) - H
        addr = r[col['Address'] + off]
        try: ie = int(r[col['Instructions Executed'] + off] or 0); ns = int(r[col['# Samples'] + off] or 0)
        except ValueError: continue
        seen[addr] = (cur, ie, ns)      # last occurrence = innermost line
by = collections.Counter(); sm = collections.Counter()
for a, (line, ie, ns) in seen.items(): by[line] += ie; sm[line] += ns
tot = sum(by.values()); ts = sum(sm.values())
print(f"total {tot} = {tot/units:.1f} per unit, {ts} samples")
for line, ie in by.most_common(top):
    print(f"L{line:<4d} {ie/units:7.2f}  {100*sm[line]/max(ts,1):5.1f}%  {text[line]}")
