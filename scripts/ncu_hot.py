"""Top sampled SASS instructions of an ncu `--page source --csv --print-source cuda,sass` dump, with the
dominant stall reasons and the CUDA source line they belong to.  usage: ncu_hot.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows[:20]) if 'Instructions Executed' in r)
hdr = rows[hi]; H = len(hdr)
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
cur = None; out = []
for r in rows[hi + 1:]:
    if not r: continue
    if r[0] == "File Name": continue
    if r[0].strip().isdigit():
        cur = int(r[0]); continue
    if r[0] == "" and len(r) >= H:
        off = len(r) - H
        try:
            ns = int(r[col['# Samples'] + off])
        except ValueError:
            continue
        st = {s: int(r[col[s] + off] or 0) for s in stalls}
        out.append((ns, cur, r[col['Address'] + off], ",".join(r[3:4 + off])[:60], st, int(r[col['Instructions Executed'] + off] or 0)))
tot = sum(o[0] for o in out)
print("total samples", tot)
for ns, line, addr, sass, st, ie in sorted(out, key=lambda o: -o[0])[:top]:
    ss = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*ns/tot:5.1f}% L{line:<4d} {sass:60s} x{ie:<9d} " + " ".join(f"{k[6:]}={v}" for k, v in ss if v))
