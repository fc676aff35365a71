"""Developer check for several GPUs of one box through the C ABI (ub200_multi_*): results equal the single-device call,
and the batch throughput with every visible GPU.  usage: multi_check.py [nodes] [samples]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from usher_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
s = capi.Synth(n, 30.0, 30000, 0, 20260930)
ndev = capi.lib().ub200_device_count()
t = time.time(); mm = capi.MultiMat(s.flat); print(f"{ndev} GPUs: replicas of the {n}-node tree in {time.time()-t:.1f}s", flush=True)
mm.set_pass_samples(96)
one = capi.Mat.from_flat_struct(s.flat); one.set_pass_samples(96)
for fam in (0, 1, 2):
    sp, sc, _ = s.samples(B, fam, 5 + fam)
    mm.place_batch(sp, sc)
    t = time.time(); b = mm.place_batch(sp, sc); tm = time.time() - t
    t = time.time(); a = one.place_batch(sp, sc); t1 = time.time() - t
    ok = np.array_equal(a["placements"], b["placements"])
    print(f"family {fam}: equal={ok}  1 GPU {B/t1:.0f} placements/s, {ndev} GPUs {B/tm:.0f} placements/s (x{t1/tm:.2f}), host buffers in and out", flush=True)
    assert ok
