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


def _random_fs_case(seed, n_leaves, n_sites, p_amb=0.1):
    """A random multifurcating tree (newick with named leaves) and per-site leaf genotypes; returns the tree in BFS order
    plus the VCF text the reference reads."""
    rng = np.random.default_rng(seed)
    kids = {0: []}
    nxt = 1
    frontier = [0]
    leaves = []
    while len(leaves) + len(frontier) < n_leaves:
        u = frontier.pop(int(rng.integers(len(frontier))))
        for _ in range(int(rng.integers(2, 5))):
            kids[u].append(nxt); kids[nxt] = []; frontier.append(nxt); nxt += 1
    leaves = sorted(v for v in kids if not kids[v])
    name = {v: f"s{v}" for v in leaves}

    def nwk(u):
        return name[u] if not kids[u] else "(" + ",".join(nwk(c) for c in kids[u]) + ")"
    newick = nwk(0) + ";"
    # BFS order
    order, q = [], [0]
    while q:
        u = q.pop(0); order.append(u); q.extend(kids[u])
    idx = {v: i for i, v in enumerate(order)}
    parent_bfs = np.array([-1] + [idx[next(p for p in kids if v in kids[p])] for v in order[1:]], np.int32)
    nuc = "NACMGRSVTWYHKDBN"
    ref_code, var_ptr, var_node, var_nuc, rows = [], [0], [], [], []
    for s in range(n_sites):
        ref = int(rng.integers(4))
        gts = {}
        for v in leaves:
            r = rng.random()
            if r < 0.25:
                a = 1 << int(rng.integers(4))
                if rng.random() < p_amb:
                    a |= 1 << int(rng.integers(4))
                if rng.random() < 0.05:
                    a = 15
                if a != (1 << ref):
                    gts[v] = a
        alts = sorted(set(gts.values()))
        if not alts:
            continue
        ref_code.append(ref)
        for v in leaves:
            if v in gts:
                var_node.append(idx[v]); var_nuc.append(gts[v])
        var_ptr.append(len(var_node))
        rows.append("\t".join(["c", str(10 * s + 1), ".", nuc[1 << ref], ",".join(nuc[a] for a in alts), ".", ".", ".", "GT"] +
                              [str(1 + alts.index(gts[v])) if v in gts else "0" for v in leaves]))
    vcf = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(name[v] for v in leaves) + "\n" + "\n".join(rows) + "\n"
    return newick, vcf, parent_bfs, np.array(ref_code, np.uint8), np.array(var_ptr, np.uint64), np.array(var_node, np.uint32), np.array(var_nuc, np.uint8)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", range(4))
def test_fitch_sankoff_oracle_matches_reference(seed, tmp_path):
    """oracle/fitch_sankoff.py against the reference's own mapper_body: a MAT built by oracle/_ref from a random newick +
    VCF (no condensing) carries exactly the mutations the restatement assigns, node by node."""
    from oracle import fitch_sankoff
    newick, vcf, parent_bfs, ref_code, var_ptr, var_node, var_nuc = _random_fs_case(700 + seed, [6, 20, 60, 150][seed], 40)
    (tmp_path / "t.nh").write_text(newick)
    (tmp_path / "v.vcf").write_text(vcf)
    rt = ref.RefTree.from_newick_vcf(str(tmp_path / "t.nh"), str(tmp_path / "v.vcf"), False, 1)
    parent, row_ptr, muts, names = rt.export()      # DFS order
    rt.close()
    site, node, par, st = fitch_sankoff.assign(parent_bfs, ref_code, var_ptr, var_node, var_nuc)
    # BFS index -> DFS index through the structure: rebuild the BFS order of the exported tree
    n = len(parent)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[parent[i]].append(i)
    bfs, q = [], [0]
    while q:
        u = q.pop(0); bfs.append(u); q.extend(kids[u])
    assert len(bfs) == len(parent_bfs)
    positions = sorted(set(int(p) for p in muts["position"]))
    got = sorted((int(muts[k]["position"]), i, int(muts[k]["par_nuc"]), int(muts[k]["mut_nuc"])) for i in range(n)
                 for k in range(int(row_ptr[i]), int(row_ptr[i + 1])))
    site_pos = [int(l.split("\t")[1]) for l in vcf.splitlines() if not l.startswith("#")]
    exp = sorted((site_pos[int(s)], bfs[int(v)], 1 << int(a), 1 << int(b)) for s, v, a, b in zip(site, node, par, st))
    assert got == exp
