"""TEST INFRASTRUCTURE ONLY — reference spot check of GPU results on trees where one full reference search is
minutes of CPU (10 M nodes).  Used by tests/ and, outside every timed region, by bench.py's cpu_baseline leg.

For a few samples the GPU's whole optimal set and per-node scores (`-p`) are compared with the reference's own
mapper2_body (oracle/_ref, `usher_ref_score_nodes`) at the GPU's optimal set plus `n_random` random nodes:
  (i)   the reference score (+1 on invalid nodes, usher_mapper.cpp:448-450,498-503) equals the GPU's at every checked node;
  (ii)  no VALID checked node scores below the GPU's best (usher_mapper.cpp:454-468);
  (iii) every node of the GPU's optimal set is valid, scores exactly the best, and carries the reference's has_unique;
  (iv)  every valid random node that scores the best is in the GPU's optimal set (completeness on the subset);
  (v)   num_best = |optimal set| and the reported best node is the set's maximum in the reference's tie-break
        order (more leaves first, then the larger BFS index; usher_mapper.cpp:476-493).
"""
import numpy as np


def spot_check(mat, rt, s_ptr, calls, sample_ids, n_random=100_000, seed=1, threads=1, label=""):
    """mat: usher_b200.capi.Mat, rt: oracle.ref.RefTree built from the same flat tree.  Raises AssertionError on
    any difference; returns a small summary dict."""
    sample_ids = [int(i) for i in sample_ids]
    lens = [int(s_ptr[i + 1]) - int(s_ptr[i]) for i in sample_ids]
    sp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    sc = np.concatenate([calls[int(s_ptr[i]):int(s_ptr[i + 1])] for i in sample_ids])
    res = mat.place_batch(sp, sc, node_scores=True, best_set=True)
    pl = res["placements"]
    bfs, num_leaves, _ = mat.node_arrays()
    rng = np.random.default_rng(seed)
    checked = 0
    out = []
    for k in range(len(sample_ids)):
        lo, hi = int(res["best_set_ptr"][k]), int(res["best_set_ptr"][k + 1])
        bset = res["best_set"][lo:hi].astype(np.uint32)
        bset_u = res["best_set_unique"][lo:hi]
        rnd = rng.integers(0, mat.n, size=n_random, dtype=np.uint32)
        nodes = np.concatenate([bset, rnd])
        ref_sc, ref_valid = rt.score_nodes(sc[int(sp[k]):int(sp[k + 1])], nodes, threads)
        gpu_sc = res["node_scores"][k][nodes]
        bad = np.flatnonzero(ref_sc != gpu_sc)
        assert bad.size == 0, f"{label} sample {sample_ids[k]}: per-node score differs at nodes {nodes[bad][:8]}: " \
                              f"ref {ref_sc[bad][:8]} gpu {gpu_sc[bad][:8]}"
        best = int(pl["score"][k])
        v = ref_valid != 0
        assert not np.any(ref_sc[v] < best), f"{label} sample {sample_ids[k]}: a valid node beats the GPU best {best}"
        nb = len(bset)
        assert nb == int(pl["num_best"][k]), f"{label}: num_best {pl['num_best'][k]} != |optimal set| {nb}"
        assert np.all(v[:nb]) and np.all(ref_sc[:nb] == best), f"{label}: optimal-set node invalid or not optimal in the reference"
        assert np.array_equal((ref_valid[:nb] >> 1) & 1, bset_u), f"{label}: has_unique of the optimal set differs"
        hit = rnd[v[nb:] & (ref_sc[nb:] == best)]
        assert np.all(np.isin(hit, bset)), f"{label}: optimal random node missing from the GPU's optimal set"
        order = np.lexsort((bfs[bset], num_leaves[bset]))
        assert int(bset[order[-1]]) == int(pl["best_node"][k]), f"{label}: tie-break differs"
        assert int(bfs[pl["best_node"][k]]) == int(pl["best_j"][k])
        checked += len(nodes)
        out.append({"sample": sample_ids[k], "score": best, "num_best": nb, "calls": lens[k]})
    return {"samples": out, "nodes_checked": int(checked)}
