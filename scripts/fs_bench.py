"""Developer measurement of the GPU Fitch-Sankoff assignment (N3): sites/s through the C ABI on a random tree, next to
the plain restatement on a few sites.  usage: fs_bench.py [leaves] [sites]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from usher_b200 import capi
from oracle import fitch_sankoff
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(5)
# random tree in BFS order: node i attaches to a random earlier node of the previous "generation window"
n = 2 * nl
parent = np.zeros(n, np.int64); parent[0] = -1
parent[1:] = np.sort(rng.integers(0, np.maximum(1, np.arange(1, n) // 2), n - 1))   # non-decreasing parents = BFS order
nch = np.bincount(parent[1:], minlength=n)
leaves = np.flatnonzero(nch == 0)
ref_code = rng.integers(0, 4, ns).astype(np.uint8)
var_ptr = [0]; var_node = []; var_nuc = []
for s in range(ns):
    k = rng.integers(1, max(2, len(leaves) // 50))
    v = rng.choice(leaves, k, replace=False)
    var_node.append(np.sort(v)); var_nuc.append((1 << rng.integers(0, 4, k)).astype(np.uint8)); var_ptr.append(var_ptr[-1] + k)
var_node = np.concatenate(var_node).astype(np.uint32); var_nuc = np.concatenate(var_nuc)
capi.fitch_sankoff(parent, ref_code[:8], var_ptr[:9], var_node[:var_ptr[8]], var_nuc[:var_ptr[8]])
t = time.time(); got = capi.fitch_sankoff(parent, ref_code, var_ptr, var_node, var_nuc); dt = time.time() - t
print(f"GPU: {n} nodes x {ns} sites in {dt:.3f}s = {ns/dt:.0f} sites/s, {n*ns/dt/1e9:.2f} G node-sites/s, {len(got[0])} mutations (host buffers in/out)")
k = 3
t = time.time(); exp = fitch_sankoff.assign(parent, ref_code[:k], var_ptr[:k + 1], var_node[:var_ptr[k]], var_nuc[:var_ptr[k]]); dc = (time.time() - t) / k
sel = got[0] < k
assert all(np.array_equal(np.asarray(a)[sel].astype(np.int64), np.asarray(b).astype(np.int64)) for a, b in zip(got, exp))
print(f"restatement (python): {dc:.2f}s per site; first {k} sites identical")
# the reference's own mapper_body (oracle/_ref: tbb::flow shim on std::thread) building a MAT from files, beside the GPU
# on the same (smaller) input
from oracle import ref
if ref.available():
    import tempfile
    from test_oracle import _random_fs_case
    newick, vcf, pb, rc2, vp2, vn2, vc2 = _random_fs_case(3, 6_000, 300, p_amb=0.1)
    d = tempfile.mkdtemp(); open(d + "/t.nh", "w").write(newick); open(d + "/v.vcf", "w").write(vcf)
    thr = len(os.sched_getaffinity(0))
    t = time.time(); rt = ref.RefTree.from_newick_vcf(d + "/t.nh", d + "/v.vcf", False, thr); tr = time.time() - t
    rt.close()
    capi.fitch_sankoff(pb, rc2, vp2, vn2, vc2)
    t = time.time(); capi.fitch_sankoff(pb, rc2, vp2, vn2, vc2); tg = time.time() - t
    print(f"{len(pb)} nodes x {len(rc2)} sites: reference from_newick_vcf (parse + mapper_body, {thr} threads) {tr:.2f}s = {len(rc2)/tr:.0f} sites/s; "
          f"GPU assignment through the C ABI {tg*1e3:.1f} ms = {len(rc2)/tg:.0f} sites/s")
