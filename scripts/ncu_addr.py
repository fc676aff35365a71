"""Executed warp-instructions of an ncu source csv in SASS address order, in chunks (each address counted once).
usage: ncu_addr.py file.csv units [chunk]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); units = float(sys.argv[2]); chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 64
hi = next(i for i, r in enumerate(rows[:20]) if 'Instructions Executed' in r)
hdr = rows[hi]; H = len(hdr); col = {n: i for i, n in enumerate(hdr)}
cur = None; seen = {}
for r in rows[hi + 1:]:
    if not r or r[0] == "File Name": continue
    if r[0].strip().isdigit(): cur = int(r[0]); continue
    if r[0] == "" and len(r) >= H:
        off = len(r) - H
        try:
            addr = int(r[col['Address'] + off], 16); ie = int(r[col['Instructions Executed'] + off] or 0); ns = int(r[col['# Samples'] + off] or 0)
        except ValueError: continue
        sass = ",".join(r[3:4 + off])[:40]
        seen.setdefault(addr, [ie, ns, set(), sass])[2].add(cur)
addrs = sorted(seen)
tot = sum(seen[a][0] for a in addrs); ts = sum(seen[a][1] for a in addrs)
print(f"total {tot} = {tot/units:.2f} per unit; {len(addrs)} instructions; {ts} samples")
for i in range(0, len(addrs), chunk):
    c = addrs[i:i + chunk]
    ie = sum(seen[a][0] for a in c); ns = sum(seen[a][1] for a in c)
    lines = sorted(set().union(*[seen[a][2] for a in c]))
    print(f"[{i:5d}] {ie/units:7.2f} {100*ns/ts:5.1f}%  lines {lines[0]}..{lines[-1]}  {seen[c[0]][3]}")
