"""CPU: the oracle itself.  (1) the plain-C port (oracle/usher_port.c) reproduces every golden vector minted
from the reference's own sources; (2) where oracle/_ref is present (this container; it also travels to the
GPU box as a prebuilt .so) the port and the reference agree on fresh randomized trees."""
import numpy as np
import pytest

import common
import small_synth
from oracle import port, ref


@pytest.mark.parametrize("path", common.golden_cases(), ids=lambda p: p.split("/")[-1])
def test_port_matches_golden(path):
    g = common.load(path)
    pt = port.PortTree(g["parent"], g["row_ptr"], g["muts"])
    got = pt.search(g["s_ptr"], g["calls"])
    got["node_scores"] = pt.search(g["s_ptr"], g["calls"], per_node=True)["node_scores"]
    common.assert_matches_expected(g, got, path)
    pt.close()


def test_config1_reference_values():
    """The frozen-tree placements of test/new_samples.vcf on the config-1 MAT (SURVEY.md §8c / BASELINE.md §2)."""
    g = common.load(common.GOLDEN + "/config1.npz")
    assert len(g["parent"]) == 474 and int(g["tree_parsimony"]) == 500 and int(g["final_parsimony"]) == 503
    assert g["exp_score"].tolist() == [1, 2, 2, 3, 3]
    assert g["exp_best_j"].tolist() == [38] * 5 and g["exp_num_best"].tolist() == [2] * 5
    names = g["names"].tolist()
    assert [names[i] for i in g["exp_best_dfs"]] == ["node_7"] * 5
    assert str(g["placement_stats"]) == "Sample1\t1\t2\t\nSample2\t1\t1\t\nSample3\t0\t1\t\nSample4\t1\t1\t\nSample5\t0\t1\t\n"


def test_branchlen2_known_answer():
    """scripts/testBranchLen2.*: the input branch lengths are the expected per-branch mutation counts."""
    g = common.load(common.GOLDEN + "/branchlen2.npz")
    assert int(g["tree_parsimony"]) == 17
    nm = dict(zip(g["names"].tolist(), np.diff(g["row_ptr"]).tolist()))
    assert nm["d"] == 1 and nm["node_4"] == 1 and nm["node_3"] == 2 and nm["node_6"] == 3
    assert nm["node_5"] == 4 and nm["node_2"] == 5 and nm["f"] == 1 and nm["h"] == 0


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(12))
def test_port_matches_reference_random(seed):
    n = [2, 9, 40, 150, 400, 700][seed % 6]
    L = [5, 25, 90, 300][seed % 4]
    mu = [0.3, 1.0, 3.0, 6.0][(seed // 3) % 4]
    shape = ["uniform", "chain", "star"][(seed // 2) % 3]
    parent, row_ptr, muts, refg = small_synth.random_mat(5000 + seed, n, L, mu, shape=shape)
    s_ptr, calls = small_synth.random_samples(6000 + seed, parent, row_ptr, muts, refg, 24)
    rt = ref.RefTree.from_flat(parent, row_ptr, muts)
    pt = port.PortTree(parent, row_ptr, muts)
    for threads in (1, 4):
        o = rt.search(s_ptr, calls, n, threads=threads)
        q = pt.search(s_ptr, calls)
        for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
            assert np.array_equal(o[k], q[k]), (seed, threads, k)
    assert np.array_equal(rt.search(s_ptr, calls, n, per_node=True)["node_scores"],
                          pt.search(s_ptr, calls, per_node=True)["node_scores"])
    rt.close()
    pt.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("path", common.golden_cases()[::4], ids=lambda p: p.split("/")[-1])
def test_reference_score_nodes_matches_golden(path):
    """usher_ref_score_nodes (the at-size spot checker, oracle/spotcheck.py): the reference's per-node score and
    validity at a node list equal the golden -p scores and the golden optimal sets."""
    g = common.load(path)
    n = len(g["parent"])
    rt = ref.RefTree.from_flat(g["parent"], g["row_ptr"], g["muts"])
    nodes = np.arange(n, dtype=np.uint32)[::-1].copy()
    for s in range(min(6, len(g["s_ptr"]) - 1)):
        c = g["calls"][int(g["s_ptr"][s]):int(g["s_ptr"][s + 1])]
        sc, valid = rt.score_nodes(c, nodes, threads=1 + s % 3)
        assert np.array_equal(sc, g["exp_node_scores"][s][nodes])
        opt = np.sort(nodes[(valid != 0) & (sc == g["exp_score"][s])])
        lo, hi = int(g["exp_best_set_ptr"][s]), int(g["exp_best_set_ptr"][s + 1])
        assert np.array_equal(opt, g["exp_best_set"][lo:hi])
        u = {int(a): int(b) for a, b in zip(nodes, (valid >> 1) & 1)}
        assert [u[int(a)] for a in g["exp_best_set"][lo:hi]] == g["exp_best_set_unique"][lo:hi].tolist()
    rt.close()
