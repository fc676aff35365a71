"""Developer A/B: time kernel variants (usher_b200/build.py UB200_VARIANT builds) side by side on one tree, and check that
every variant returns the same placements as the default build.
usage: variants.py <c2|c3|mid|c4> <family> <n_samples:pass_samples:groups_per_scan>[,...] <variant>[ <variant> ...]
       ("base" = libusher_b200.so; "acc2_nofence" = libusher_b200_v_acc2_nofence.so)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import capi

SHAPES = {"c2": (100_000, 30.0, 30000, 0, 20260927), "c3": (2_000_000, 1.2, 29903, 1, 20260928),
          "mid": (2_000_000, 30.0, 30000, 0, 20260930), "c4": (10_000_000, 30.0, 30000, 0, 20260929)}
name, fam = sys.argv[1], int(sys.argv[2])
configs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[3].split(",")]
variants = sys.argv[4:]
n, mu, L, shape, seed = SHAPES[name]
s = capi.Synth(n, mu, L, shape, seed)
batches = {ns: s.samples(ns, fam, 3) for ns, _, _ in configs}
capi.lib()
capi._build.build = lambda *a, **k: None   # the variant libraries are prebuilt; never rebuild on the box
ref = {}
for v in variants:
    capi._lib = None
    capi._build.LIB = os.path.join(capi._build.PKG, "libusher_b200.so" if v == "base" else f"libusher_b200_v_{v}.so")
    m = capi.Mat.from_flat_struct(s.flat)
    for ns, ps, nc in configs:
        sp, sc, _ = batches[ns]
        m.set_pass_samples(ps)
        m.set_scan_sharing(nc)
        S = m.upload(sp, sc)
        per = []
        for _ in range(5):
            S.place()
            tm = m.timing()
            per.append(tm.score_ms / tm.score_launches * 1e3)
        key = np.asarray(S.download()).tobytes()
        ref.setdefault((ns, ps, nc), key)
        print(f"{name} fam={fam} ns={ns} pass={ps} nc={nc} {v:>16}: per launch {min(per[1:]):.1f} us (runs {', '.join('%.1f' % x for x in per)})"
              f" {'same results' if key == ref[(ns, ps, nc)] else 'RESULTS DIFFER'}", flush=True)
        S.close()
    m.close()
