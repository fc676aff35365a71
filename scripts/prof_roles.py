"""Developer probe (UB200_PROFILE=1 build): cycles of the warps by phase for one shape / family.
usage: python scripts/prof_roles.py <shape> <fam> <pass> <nc> [kernel 4|5]"""
import os, sys, ctypes as C
os.environ["UB200_PROFILE"] = "1"
if len(sys.argv) > 5: os.environ["UB200_KERNEL"] = sys.argv[5]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import capi
SHAPES = {"c2": (100_000, 30.0, 30000, 0, 20260927), "c3": (2_000_000, 1.2, 29903, 1, 20260928),
          "mid": (2_000_000, 30.0, 30000, 0, 20260930), "c4": (10_000_000, 30.0, 30000, 0, 20260929)}
name, fam, ps, nc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
k5 = os.environ.get("UB200_KERNEL", "5") != "4"
n, mu, L, shape, seed = SHAPES[name]
s = capi.Synth(n, mu, L, shape, seed)
m = capi.Mat.from_flat_struct(s.flat)
m.set_pass_samples(ps); m.set_scan_sharing(nc)
sp, sc, _ = s.samples(ps, fam, 3)
S = m.upload(sp, sc)
lib = capi.lib()
lib.ub200_debug_prof.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
out = np.zeros(16, np.uint64)
for rep in range(3):
    S.place()
    tm = m.timing()
    lib.ub200_debug_prof(m.h, S.h, out.ctypes.data_as(C.c_void_p))
sm = 148
us = lambda cyc, k: cyc / k / 1965.0
if k5:
    W = int(os.environ.get("UB200_WORKERS", "26")); nw = sm * W
    print(f"{name} fam={fam} pass={ps} k_score5 ({W} workers): launch {tm.score_ms*1e3:.1f} us")
    print(f" worker (avg): total {us(out[0],nw):.1f} us | load+test {us(out[4],nw):.1f} over {out[5]/nw:.0f} steps | emit {us(out[6],nw):.1f} | "
          f"hit rows+apply {us(out[11],nw):.1f} | bound+eval+chain {us(out[14],nw):.1f} ({out[15]/nw:.1f} blocks evaluated)")
else:
    units = {1: 16, 2: 10, 3: 8}[nc]
    nscan = sm * units; ncons = sm * units * nc
    print(f"{name} fam={fam} pass={ps} nc={nc} k_score4: launch {tm.score_ms*1e3:.1f} us")
    print(f" scanner (avg per warp): total {us(out[0],nscan):.1f} us | slot waits {us(out[1],nscan):.1f} us ({out[2]/nscan:.0f} failed polls) | "
          f"load-wait {us(out[3],nscan):.1f} | load+test {us(out[4],nscan):.1f} | steps {out[5]/nscan:.0f} | emit {us(out[6],nscan):.1f}")
    print(f" consumer (avg per warp): total {us(out[8],ncons):.1f} us | msg waits {us(out[9],ncons):.1f} us ({out[10]/ncons:.0f} failed polls) | "
          f"process {us(out[11],ncons):.1f} us over {out[12]/ncons:.0f} msgs | tile waits {us(out[13],ncons):.1f} | bound+eval {us(out[14],ncons):.1f} "
          f"({out[15]/ncons:.1f} blocks evaluated)")
