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
    cond = t.condensed()
    s_ptr, calls, snames = t.read_samples(f"{REF}/test/new_samples.vcf")
    e = expected(t, s_ptr, calls, len(parent))
    pars0 = t.parsimony()
    d = tempfile.mkdtemp()
    t.usher_common(d, threads=1)
    files = {k: open(os.path.join(d, k)).read() for k in ("placement_stats.tsv", "mutation-paths.txt", "final-tree.nh")}
    # -p (per-node scores, tree untouched) and -n (no-add) runs of the reference on fresh copies of the same tree
    t2 = ref.RefTree.from_newick_vcf(f"{REF}/test/global_phylo.nh", f"{REF}/test/global_samples.vcf", True, 1)
    t2.read_samples(f"{REF}/test/new_samples.vcf")
    d2 = tempfile.mkdtemp()
    t2.usher_common(d2, threads=1, print_parsimony_scores=True)
    files["parsimony-scores.tsv"] = open(os.path.join(d2, "parsimony-scores.tsv")).read()
    files["current-tree.nh"] = open(os.path.join(d2, "current-tree.nh")).read()
    t3 = ref.RefTree.from_newick_vcf(f"{REF}/test/global_phylo.nh", f"{REF}/test/global_samples.vcf", True, 1)
    t3.read_samples(f"{REF}/test/new_samples.vcf")
    d3 = tempfile.mkdtemp()
    t3.usher_common(d3, threads=1, no_add=True)
    files["noadd_placement_stats.tsv"] = open(os.path.join(d3, "placement_stats.tsv")).read()
    files["noadd_final-tree.nh"] = open(os.path.join(d3, "final-tree.nh")).read()
    np.savez_compressed(
        os.path.join(OUT, "config1.npz"), parent=parent, row_ptr=row_ptr, muts=muts, names=np.array(names),
        s_ptr=s_ptr, calls=calls, snames=np.array(snames), tree_parsimony=pars0, final_parsimony=t.parsimony(),
        placement_stats=files["placement_stats.tsv"], mutation_paths=files["mutation-paths.txt"],
        final_tree=files["final-tree.nh"], parsimony_scores=files["parsimony-scores.tsv"],
        current_tree=files["current-tree.nh"], noadd_placement_stats=files["noadd_placement_stats.tsv"],
        noadd_final_tree=files["noadd_final-tree.nh"], **e)
    write_pb(os.path.join(OUT, "config1.pb"), parent, row_ptr, muts, names, t.condensed() if False else cond)
    write_vcf(os.path.join(OUT, "config1_new_samples.vcf"), s_ptr, calls, snames)
    print("config1:", len(parent), "nodes;", e["exp_score"], e["exp_best_j"], e["exp_num_best"])


def newick_from_flat(parent, names, nmut):
    """What save_mutation_annotated_tree writes: no internal names, branch length = #mutations."""
    kids = [[] for _ in parent]
    for i, p in enumerate(parent):
        if p >= 0:
            kids[p].append(i)

    def rec(i):
        ln = ":%g" % float(nmut[i])
        if not kids[i]:
            return names[i] + ln
        return "(" + ",".join(rec(c) for c in kids[i]) + ")" + ln
    sys.setrecursionlimit(100000)
    return rec(0) + ";"


def write_pb(path, parent, row_ptr, muts, names, condensed):
    """Serialise with the reference's own parsimony_pb2.py (pure-python protobuf runtime)."""
    os.environ["PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION"] = "python"
    sys.path.insert(0, REF)
    import parsimony_pb2
    d = parsimony_pb2.data()
    nmut = np.diff(row_ptr.astype(np.int64))
    d.newick = newick_from_flat(parent.tolist(), names, nmut)
    code = {1: 0, 2: 1, 4: 2, 8: 3}
    for i in range(len(parent)):
        ml = d.node_mutations.add()
        d.metadata.add()
        for k in range(int(row_ptr[i]), int(row_ptr[i + 1])):
            m = ml.mutation.add()
            m.position = int(muts[k]["position"])
            if m.position < 0:
                m.ref_nuc = -1
                m.par_nuc = -1
            else:
                m.ref_nuc = code[int(muts[k]["ref_nuc"])]
                m.par_nuc = code[int(muts[k]["par_nuc"])]
                m.mut_nuc.append(code[int(muts[k]["mut_nuc"])])
    for name, members in condensed.items():
        c = d.condensed_nodes.add()
        c.node_name = name
        c.condensed_leaves.extend(members)
    open(path, "wb").write(d.SerializeToString())


def write_vcf(path, s_ptr, calls, snames):
    nuc = "NACMGRSVTWYHKDBN"
    rows = {}
    for s in range(len(snames)):
        for k in range(int(s_ptr[s]), int(s_ptr[s + 1])):
            c = calls[k]
            rows.setdefault(int(c["position"]), {"ref": nuc[int(c["ref_nuc"])], "gt": {}})["gt"][s] = (
                "N" if c["is_missing"] else nuc[int(c["mut_nuc"])])
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(snames) + "\n")
        for pos in sorted(rows):
            alts = sorted(set(rows[pos]["gt"].values()))
            gts = [str(1 + alts.index(rows[pos]["gt"][s])) if s in rows[pos]["gt"] else "0" for s in range(len(snames))]
            f.write("\t".join(["NC_045512v2", str(pos), ".", rows[pos]["ref"], ",".join(alts), ".", ".", ".", "GT"] + gts) + "\n")


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
