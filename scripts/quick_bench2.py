"""Developer probe (not the contract bench): per-launch timing of the scoring kernel per shape / family / pass
width / scan sharing.  usage: quick_bench2.py <shape>[,<shape>...] [fams] [configs pass:nc,...]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import capi

SHAPES = {"c2": (100_000, 30.0, 30000, 0, 20260927), "c3": (2_000_000, 1.2, 29903, 1, 20260928),
          "mid": (2_000_000, 30.0, 30000, 0, 20260930), "c4": (10_000_000, 30.0, 30000, 0, 20260929)}
shapes = sys.argv[1].split(",")
fams = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2]
cfgs = [tuple(int(y) for y in x.split(":")) for x in (sys.argv[3] if len(sys.argv) > 3 else "32:1,64:2,96:3,192:3").split(",")]
for name in shapes:
    n, mu, L, shape, seed = SHAPES[name]
    s = capi.Synth(n, mu, L, shape, seed)
    m = capi.Mat.from_flat_struct(s.flat)
    print(f"[{name}] nodes={n} muts={s.m} tiles={m.info.n_tiles} depth={m.info.max_level} alg={m.info.algorithmic_bytes/1e6:.1f}MB", flush=True)
    for fam in fams:
        ref = None
        for ps, nc in cfgs:
            ns = max(ps * 2, 192 if name == "c4" else 768)
            ns = (ns + ps - 1) // ps * ps
            sp, sc, _ = s.samples(ns, fam, 3)
            S = m.upload(sp, sc)
            m.set_pass_samples(ps); m.set_scan_sharing(nc)
            S.place(); S.place(); S.place()
            tm = m.timing()
            per = tm.score_ms / tm.score_launches
            gbs = tm.score_bytes / tm.score_launches / per / 1e6
            res = S.download()
            key = (res["score"][:64].tolist(), res["best_node"][:64].tolist())
            print(f"  fam={fam} pass={ps:3d} nc={nc}: launches={tm.score_launches} per_launch={per*1e3:8.1f}us {gbs:6.0f} GB/s "
                  f"prep={tm.prep_ms:.3f} reduce={tm.reduce_ms:.3f} -> {ns/(tm.score_ms+tm.prep_ms+tm.reduce_ms)*1e3:9.0f} placements/s", flush=True)
            S.close()
    m.close(); s.close()
