"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck) covering both kernels and all modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from usher_b200 import capi
import small_synth
for seed, n, L, mu, shape in ((1, 900, 120, 4.0, "uniform"), (2, 700, 300, 2.0, "chain")):
    parent, row_ptr, muts, refg = small_synth.random_mat(seed, n, L, mu, shape=shape)
    s_ptr, calls = small_synth.random_samples(seed + 50, parent, row_ptr, muts, refg, 70)
    m = capi.Mat(parent, row_ptr, muts)
    r = m.place_batch(s_ptr, calls, node_scores=True, best_set=True)
    print(seed, r["placements"]["score"][:8], int(r["best_set_ptr"][-1]))
    m.close()
s = capi.Synth(20000, 30.0, 30000, 0, 5)
m = capi.Mat.from_flat_struct(s.flat)
for ps, nc in ((32, 1), (64, 2), (96, 3)):       # single-group scans and shared scans (union bitmap, NC consumers)
    m.set_pass_samples(ps); m.set_scan_sharing(nc)
    sp, sc, _ = s.samples(100, 2, 1)
    print(ps, nc, m.place_batch(sp, sc, best_set=True)["placements"]["score"][:8])
