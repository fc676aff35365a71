"""Workload for ncu captures: one resident batch, a few scoring launches.
usage: prof_one.py <nodes> <family> <samples_per_launch> [mu] [shape] [genome_len] [n_samples]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usher_b200 import capi
a = sys.argv[1:]
n = int(a[0]) if len(a) > 0 else 500_000
fam = int(a[1]) if len(a) > 1 else 0
ps = int(a[2]) if len(a) > 2 else 32
mu = float(a[3]) if len(a) > 3 else 30.0
shape = int(a[4]) if len(a) > 4 else 0
L = int(a[5]) if len(a) > 5 else 30000
ns = int(a[6]) if len(a) > 6 else ps * 3
s = capi.Synth(n, mu, L, shape, 1)
m = capi.Mat.from_flat_struct(s.flat)
m.set_pass_samples(ps)
sp, sc, _ = s.samples(ns, fam, 3)
S = m.upload(sp, sc)
for _ in range(2):
    S.place()
t = m.timing()
print("nodes", n, "muts", s.m, "score_ms", t.score_ms, "launches", t.score_launches, "GB/s", t.score_bytes / t.score_ms / 1e6,
      "alg_bytes", m.info.algorithmic_bytes)
