"""Workload for ncu captures: one resident batch, a few scoring launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usher_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
fam = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ps = int(sys.argv[3]) if len(sys.argv) > 3 else 32
s = capi.Synth(n, 30.0, 30000, 0, 1)
m = capi.Mat.from_flat_struct(s.flat)
m.set_pass_samples(ps)
sp, sc, _ = s.samples(ps * 3, fam, 3)
S = m.upload(sp, sc)
for _ in range(2):
    S.place()
t = m.timing()
print("score_ms", t.score_ms, "launches", t.score_launches, "GB/s", t.score_bytes / t.score_ms / 1e6)
