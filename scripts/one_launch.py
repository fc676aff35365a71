"""Developer probe: a few scoring launches of one (tree shape, sample family) for ncu / timing.
usage: one_launch.py <c2|c3|mid|c4> <family 0|1|2> [n_samples] [pass_samples] [repeats] [groups_per_scan]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usher_b200 import capi

SHAPES = {"c2": (100_000, 30.0, 30000, 0, 20260927), "c3": (2_000_000, 1.2, 29903, 1, 20260928),
          "mid": (2_000_000, 30.0, 30000, 0, 20260930), "c4": (10_000_000, 30.0, 30000, 0, 20260929)}
name = sys.argv[1]; fam = int(sys.argv[2])
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ps = int(sys.argv[4]) if len(sys.argv) > 4 else 32
rep = int(sys.argv[5]) if len(sys.argv) > 5 else 3
nc = int(sys.argv[6]) if len(sys.argv) > 6 else 0
n, mu, L, shape, seed = SHAPES[name]
s = capi.Synth(n, mu, L, shape, seed)
m = capi.Mat.from_flat_struct(s.flat)
m.set_pass_samples(ps)
m.set_scan_sharing(nc)
sp, sc, _ = s.samples(ns, fam, 3)
S = m.upload(sp, sc)
for _ in range(rep):
    S.place()
    tm = m.timing()
    per = tm.score_ms / tm.score_launches
    print(f"{name} fam={fam} ns={ns} pass={ps} nc={nc}: launches={tm.score_launches} per_launch={per*1e3:.1f}us "
          f"{tm.score_bytes / tm.score_launches / per / 1e6:.0f} GB/s prep={tm.prep_ms:.3f} reduce={tm.reduce_ms:.3f}", flush=True)
