"""Developer probe (not the contract bench): per-launch timing of the scoring kernel on a few shapes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import capi

def run(name, n, mu, L, shape, fams=(0, 1), nsamp=1024, passes=(32, 64, 128, 256), seed=1):
    t = time.time(); s = capi.Synth(n, mu, L, shape, seed); tg = time.time() - t
    t = time.time(); m = capi.Mat.from_flat_struct(s.flat); tc = time.time() - t
    print(f"[{name}] nodes={n} muts={s.m} gen={tg:.1f}s create={tc:.1f}s tiles={m.info.n_tiles} depth={m.info.max_level} algbytes={m.info.algorithmic_bytes/1e6:.1f}MB", flush=True)
    for fam in fams:
        sp, sc, _ = s.samples(nsamp, fam, 3)
        S = m.upload(sp, sc)
        for ps in passes:
            m.set_pass_samples(ps)
            S.place(); S.place()
            S.place()
            tm = m.timing()
            per = tm.score_ms / tm.score_launches
            gbs = tm.score_bytes / tm.score_launches / per / 1e6
            print(f"  fam={fam} pass={ps:3d}: launches={tm.score_launches} score={tm.score_ms:.3f}ms per_launch={per*1e3:.1f}us "
                  f"{gbs:.0f} GB/s  prep={tm.prep_ms:.3f} reduce={tm.reduce_ms:.3f}  -> {nsamp/(tm.score_ms+tm.prep_ms+tm.reduce_ms)*1e3:.0f} placements/s", flush=True)
        S.close()
    m.close(); s.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "mid"]
    if "c2" in which: run("c2", 100_000, 30.0, 30000, 0)
    if "c3" in which: run("c3", 2_000_000, 1.2, 29903, 1)
    if "mid" in which: run("mid", 2_000_000, 30.0, 30000, 0)
    if "c4" in which: run("c4", 10_000_000, 30.0, 30000, 0, nsamp=512, passes=(32, 64))
