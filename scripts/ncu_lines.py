"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: python scripts/ncu_lines.py file.csv <units (e.g. nodes)> [min_per_unit]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
hi = next(i for i, r in enumerate(rows[:20]) if 'Instructions Executed' in r)
hdr = rows[hi]; H = len(hdr)
ie = H - hdr.index("Instructions Executed"); iss = H - hdr.index("# Samples")
cur = None; fname = ""
inst = collections.Counter(); samp = collections.Counter(); text = {}
for r in rows[hi + 1:]:
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0].strip().isdigit():
        cur = (fname, int(r[0])); text.setdefault(cur, ",".join(r[1:len(r) - (H - 2)])[:90]); continue
    if r[0] == "" and cur is not None and len(r) >= H:
        try:
            inst[cur] += int(r[len(r) - ie]); samp[cur] += int(r[len(r) - iss])
        except ValueError:
            pass
tot = sum(inst.values()); ts = sum(samp.values())
print(f"total {tot} inst = {tot/units:.1f} per unit; {ts} samples")
for k in sorted(inst, key=lambda k: (k[0], k[1])):
    if inst[k] / units >= thr or samp[k] / max(ts, 1) > 0.01:
        print(f"{k[0][:18]:18s} {k[1]:4d} {inst[k]/units:7.2f} {100*samp[k]/max(ts,1):5.1f}%  {text[k]}")
