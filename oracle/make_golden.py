"""Mint tests/golden/*.npz from the reference itself (oracle/_ref = the reference's own sources compiled
here).  Run in a container that mounts /root/reference:   python oracle/make_golden.py
The reference repository ships inputs but no expected outputs for this path, so these files ARE
the known-answer vectors; the script that made them is committed beside them."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
import small_synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def expected(rt, s_ptr, calls, n):
    o = rt.search(s_ptr, calls, n, threads=1)
    op = rt.search(s_ptr, calls, n, threads=1, per_node=True)
    return dict(exp_score=o["score"], exp_best_dfs=o["best_dfs"], exp_best_j=o["best_j"],
                exp_num_best=o["num_best"], exp_has_unique=o["has_unique"], exp_best_set=o["best_set"],
                exp_best_set_unique=o["best_set_unique"], exp_best_set_ptr=o["best_set_ptr"],
                exp_node_scores=op["node_scores"])


def config1():
    t = ref.RefTree.from_newick_vcf(f"{REF}/test/global_phylo.nh", f"{REF}/test/global_samples.vcf", True, 1)
    parent, row_ptr, muts, names = t.export()
    s_ptr, calls, snames = t.read_samples(f"{REF}/test/new_samples.vcf")
    e = expected(t, s_ptr, calls, len(parent))
    pars0 = t.parsimony()
    d = tempfile.mkdtemp()
    t.usher_common(d, threads=1)
    files = {k: open(os.path.join(d, k)).read() for k in ("placement_stats.tsv", "mutation-paths.txt", "final-tree.nh")}
    np.savez_compressed(
        os.path.join(OUT, "config1.npz"), parent=parent, row_ptr=row_ptr, muts=muts, names=np.array(names),
        s_ptr=s_ptr, calls=calls, snames=np.array(snames), tree_parsimony=pars0, final_parsimony=t.parsimony(),
        placement_stats=files["placement_stats.tsv"], mutation_paths=files["mutation-paths.txt"],
        final_tree=files["final-tree.nh"], **e)
    print("config1:", len(parent), "nodes;", e["exp_score"], e["exp_best_j"], e["exp_num_best"])


def branchlen2():
    t = ref.RefTree.from_newick_vcf(f"{REF}/scripts/testBranchLen2.nwk", f"{REF}/scripts/testBranchLen2.vcf", False, 1)
    parent, row_ptr, muts, names = t.export()
    np.savez_compressed(os.path.join(OUT, "branchlen2.npz"), parent=parent, row_ptr=row_ptr, muts=muts,
                        names=np.array(names), tree_parsimony=t.parsimony())
    print("branchlen2:", len(parent), "nodes, parsimony", t.parsimony())


def randoms():
    cases = []
    for seed in range(16):
        n = [1, 3, 17, 60, 200, 500, 900, 1500][seed % 8]
        L = [6, 30, 120, 400][seed % 4]
        mu = [0.4, 1.5, 4.0, 8.0][(seed // 2) % 4]
        shape = ["uniform", "chain", "star"][seed % 3]
        parent, row_ptr, muts, refg = small_synth.random_mat(100 + seed, n, L, mu, shape=shape)
        s_ptr, calls = small_synth.random_samples(200 + seed, parent, row_ptr, muts, refg, 40 if n < 600 else 70)
        rt = ref.RefTree.from_flat(parent, row_ptr, muts)
        e = expected(rt, s_ptr, calls, n)
        rt.close()
        np.savez_compressed(os.path.join(OUT, f"random_{seed:02d}.npz"), parent=parent, row_ptr=row_ptr, muts=muts,
                            s_ptr=s_ptr, calls=calls, **e)
        cases.append((seed, n, len(muts), len(calls)))
    print("random cases:", cases)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    config1()
    branchlen2()
    randoms()
