import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import common
from oracle import port
from usher_b200 import capi
s = capi.Synth(30_000, 8.0, 30_000, capi.Synth.SC2, 20260929)
p, r, mu = s.arrays()
sp, sc, _ = s.samples(96, capi.Synth.AMBIG, 55)
pt = port.PortTree(p, r, mu)
o = pt.search(sp, sc)
for rep in range(4):
    m = capi.Mat.from_flat_struct(s.flat)
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    bad = []
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique"):
        if not np.array_equal(np.asarray(got[k]).astype(np.int64), o[k].astype(np.int64)):
            bad.append(k)
    gp = np.asarray(got["best_set_ptr"]).astype(np.int64); ep = o["best_set_ptr"].astype(np.int64)
    nb = 0
    for i in range(96):
        a = set(np.asarray(got["best_set"])[gp[i]:gp[i+1]].tolist()); b = set(o["best_set"][ep[i]:ep[i+1]].tolist())
        if a != b:
            nb += 1
            if nb <= 3:
                print(" sample", i, "num_best", int(o["num_best"][i]), "missing", sorted(b - a)[:6], "extra", sorted(a - b)[:6],
                      "levels", m.node_arrays()[2][sorted(b - a)[:6]] if b - a else "")
    print("rep", rep, "tiles", m.info.n_tiles, "max_level", m.info.max_level, "bad fields", bad, "bad sets", nb, flush=True)
    m.close()
