"""Developer check for the batched sequential mode (N2): write a synthetic MAT as parsimony.proto and a VCF of new
samples, run `usher -i tree.pb -v samples.vcf -d out` (default mode: every sample is grafted before the next one is
placed) and report the time per sample.  usage: seq_check.py [nodes] [samples] [mu] [shape]"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import build, capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
mu = float(sys.argv[3]) if len(sys.argv) > 3 else 1.2
shape = int(sys.argv[4]) if len(sys.argv) > 4 else 1
s = capi.Synth(n, mu, 29903, shape, 20260928)
parent, row_ptr, muts = s.arrays()
d = tempfile.mkdtemp()


def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80); v >>= 7
    out.append(v)
    return bytes(out)


def field(num, payload):
    return varint((num << 3) | 2) + varint(len(payload)) + payload


t = time.time()
kids = [[] for _ in range(n)]
for i in range(1, n):
    kids[parent[i]].append(i)
# newick, iteratively (leaves l<i>, branch lengths irrelevant)
out = []
st = [(0, 0)]
while st:
    u, k = st.pop()
    if not kids[u]:
        out.append(f"l{u}")
        continue
    if k == 0:
        out.append("(")
    if k < len(kids[u]):
        if k:
            out.append(",")
        st.append((u, k + 1))
        st.append((kids[u][k], 0))
    else:
        out.append(")")
nwk = "".join(out) + ";"
code = {1: 0, 2: 1, 4: 2, 8: 3}
msg = bytearray(field(1, nwk.encode()))
pos = muts["position"]; ref = muts["ref_nuc"]; par = muts["par_nuc"]; mt = muts["mut_nuc"]
for i in range(n):
    lst = bytearray()
    for k in range(int(row_ptr[i]), int(row_ptr[i + 1])):
        mm = bytearray()
        mm += varint(1 << 3) + varint(int(pos[k]))
        if code[int(ref[k])]: mm += varint(2 << 3) + varint(code[int(ref[k])])
        if code[int(par[k])]: mm += varint(3 << 3) + varint(code[int(par[k])])
        mm += field(4, varint(code[int(mt[k])]))
        lst += field(1, bytes(mm))
    msg += field(2, bytes(lst))
open(d + "/tree.pb", "wb").write(bytes(msg))
print(f"wrote {n}-node tree.pb ({len(msg)/1e6:.1f} MB) in {time.time()-t:.1f}s", flush=True)
sp, sc, _ = s.samples(B, 1, 99)
nuc = "NACMGRSVTWYHKDBN"
rows = {}
for k in range(B):
    for c in sc[int(sp[k]):int(sp[k + 1])]:
        rows.setdefault(int(c["position"]), {"ref": nuc[int(c["ref_nuc"])], "gt": {}})["gt"][k] = "N" if c["is_missing"] else nuc[int(c["mut_nuc"])]
with open(d + "/samples.vcf", "w") as f:
    f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"new{k}" for k in range(B)) + "\n")
    for p in sorted(rows):
        alts = sorted(set(rows[p]["gt"].values()))
        gts = [str(1 + alts.index(rows[p]["gt"][k])) if k in rows[p]["gt"] else "0" for k in range(B)]
        f.write("\t".join(["c", str(p), ".", rows[p]["ref"], ",".join(alts), ".", ".", ".", "GT"] + gts) + "\n")
build.build()
t = time.time()
r = subprocess.run([build.USHER, "-i", d + "/tree.pb", "-v", d + "/samples.vcf", "-d", d], capture_output=True, text=True)
wall = time.time() - t
assert r.returncode == 0, r.stderr[-3000:]
per = [int(l.split()[2]) for l in r.stderr.splitlines() if l.startswith("Completed in")]
lines = r.stderr.splitlines()
i0 = next(i for i, l in enumerate(lines) if l.startswith("Adding missing samples"))
place = [int(l.split()[2]) for l in lines[i0:] if l.startswith("Completed in")][:B]
print(f"usher default mode, {n} nodes, {B} samples: whole run {wall:.1f}s (load pb, flatten, graft, write outputs)")
print(f"per-sample placement+graft: first {place[0]} ms (includes the frozen-tree batch for all samples), "
      f"then mean {np.mean(place[1:]):.1f} ms, max {max(place[1:])} ms; sum {sum(place)/1000:.2f}s = {sum(place)/B:.1f} ms per sample")
print("\n".join(l for l in lines if "Loaded" in l or "Tree resident" in l)[:400])
