"""Mint tests/golden/hostgold.npz: the reference's own usher_common() (oracle/_ref) on the config-1 MAT with new
samples that carry IUPAC-ambiguous calls, N runs, duplicated genotypes (child vs sibling grafts, condensed leaves)
and private mutations, in the default sequential mode and with the sort pre-pass (-s, -S, -A, -r), the thresholds
(-e, -E) and --no-add.  The usher binary of this repo must reproduce placement_stats.tsv, mutation-paths.txt and
final-tree.nh byte for byte (tests/test_cli.py).  Run where /root/reference is mounted:
    python oracle/make_golden_host.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from oracle.make_golden import REF, OUT, write_vcf  # noqa: E402
import small_synth  # noqa: E402

MUT_DTYPE = small_synth.MUT_DTYPE
RUNS = {
    "default": {},
    "sort1": {"sort1": True},
    "sort2": {"sort2": True},
    "sort1_reverse": {"sort1": True, "reverse": True},
    "sort3": {"sort3": True},
    "max_uncertainty_2": {"max_uncertainty": 2},
    "max_parsimony_3": {"max_parsimony": 3},
    "no_add": {"no_add": True},
    "uncondensed": {"uncondensed": True},
}


def make_samples(parent, row_ptr, muts, seed=7, B=14):
    rng = np.random.default_rng(seed)
    n = len(parent)
    refs = {}
    for m in muts:
        if m["position"] >= 0:
            refs[int(m["position"])] = int(m["ref_nuc"])
    positions = sorted(refs)
    s_ptr, calls, names = [0], [], []
    base_nodes = [int(x) for x in rng.integers(1, n, B)]
    base_nodes[3] = base_nodes[2]          # two samples from the same node: the second is grafted next to the first
    base_nodes[5] = base_nodes[4]
    for s in range(B):
        g = small_synth.genotype(parent, row_ptr, muts, base_nodes[s])
        cur = {p: (nuc, 0) for p, nuc in g.items() if nuc != refs[p]}
        if s not in (3,):                   # sample 3 is an exact duplicate of sample 2's genotype
            for _ in range(int(rng.integers(0, 4))):       # private SNVs at positions the tree knows
                p = int(rng.choice(positions))
                cur.setdefault(p, (int(rng.choice([b for b in (1, 2, 4, 8) if b != refs[p]])), 0))
            for p in list(cur):
                r = rng.random()
                if r < 0.18:                # IUPAC-widened call, with or without the reference allele
                    nuc = cur[p][0]
                    for _ in range(int(rng.integers(1, 3))):
                        nuc |= 1 << int(rng.integers(0, 4))
                    cur[p] = (nuc, 0)
                elif r < 0.26:
                    cur[p] = (15, 1)        # N on a mutated path position
            if rng.random() < 0.6:          # an N run over consecutive known positions
                i0 = int(rng.integers(0, len(positions) - 12))
                for p in positions[i0:i0 + int(rng.integers(2, 12))]:
                    cur[p] = (15, 1)
        if s == 2:
            cur = {p: v for p, v in cur.items()}
        if s == 3:
            g2 = small_synth.genotype(parent, row_ptr, muts, base_nodes[2])
            cur = dict(sample2)
        if s == 2:
            sample2 = dict(cur)
        for p in sorted(cur):
            calls.append((p, refs[p], refs[p], cur[p][0], cur[p][1]))
        s_ptr.append(len(calls))
        names.append(f"New{s + 1}")
    c = np.zeros(len(calls), MUT_DTYPE)
    for i, t in enumerate(calls):
        c[i] = t
    return np.array(s_ptr, np.uint64), c, names


def main():
    t = ref.RefTree.from_newick_vcf(f"{REF}/test/global_phylo.nh", f"{REF}/test/global_samples.vcf", True, 1)
    parent, row_ptr, muts, names = t.export()
    t.close()
    s_ptr, calls, snames = make_samples(parent, row_ptr, muts)
    vcf = os.path.join(OUT, "hostgold_samples.vcf")
    write_vcf(vcf, s_ptr, calls, snames)
    out = {}
    for name, kw in RUNS.items():
        rt = ref.RefTree.from_newick_vcf(f"{REF}/test/global_phylo.nh", f"{REF}/test/global_samples.vcf", True, 1)
        rt.read_samples(vcf)
        d = tempfile.mkdtemp()
        rt.usher_common2(d, threads=1, **kw)
        for f in ("placement_stats.tsv", "mutation-paths.txt", "final-tree.nh", "uncondensed-final-tree.nh"):
            path = os.path.join(d, f)
            if os.path.exists(path):
                out[f"{name}__{f}"] = open(path).read()
        out[f"{name}__parsimony"] = rt.parsimony()
        rt.close()
        print(name, "->", [k for k in out if k.startswith(name + "__")], flush=True)
    np.savez_compressed(os.path.join(OUT, "hostgold.npz"), **out)
    print(out["default__placement_stats.tsv"])


if __name__ == "__main__":
    main()
