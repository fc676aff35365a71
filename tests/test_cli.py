"""The drop-in `usher` binary (usher_b200/csrc/host/*.cpp over the C ABI).
CPU: the host data layer — parsimony.proto reader/writer (byte-identical round trip of a message serialised by the
reference's own parsimony_pb2), newick parser naming/order, placement-mode VCF reader — against the config-1
golden minted from the reference.  GPU: `usher -i tree.pb -v new.vcf -d out` reproduces the reference's
placement_stats.tsv, mutation-paths.txt and final-tree.nh byte for byte (sequential graft = batch of 1 with a
re-flatten per sample), plus the --no-add and --write-parsimony-scores-per-node runs."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import common
from usher_b200 import build

PB = os.path.join(common.GOLDEN, "config1.pb")
VCF = os.path.join(common.GOLDEN, "config1_new_samples.vcf")


@pytest.fixture(scope="module")
def usher():
    build.build()
    assert os.path.exists(build.USHER)
    return build.USHER


def test_pb_newick_vcf_readers_match_reference(usher):
    g = common.load(os.path.join(common.GOLDEN, "config1.npz"))
    d = tempfile.mkdtemp()
    subprocess.check_call([usher, "-i", PB, "-v", VCF, "--dump-flat", d + "/flat.txt"], stderr=subprocess.DEVNULL)
    names = g["names"].tolist()
    parent = g["parent"]
    row_ptr = g["row_ptr"].astype(np.int64)
    muts = g["muts"]
    nodes = [l.rstrip("\n").split("\t") for l in open(d + "/flat.txt") if l.startswith("N\t")]
    assert [n[1] for n in nodes] == names
    assert [n[2] for n in nodes] == ["" if p < 0 else names[p] for p in parent]
    for i, n in enumerate(nodes):
        exp = "".join(f"{m['position']}:{m['ref_nuc']}:{m['par_nuc']}:{m['mut_nuc']}," for m in muts[row_ptr[i]:row_ptr[i + 1]])
        assert n[3] == exp, (i, n[3], exp)
    samples = [l.rstrip("\n").split("\t") for l in open(d + "/flat.txt") if l.startswith("S\t")]
    assert [s[1] for s in samples] == g["snames"].tolist()
    sp = g["s_ptr"].astype(np.int64)
    for i, s in enumerate(samples):
        exp = "".join(f"{c['position']}:{c['ref_nuc']}:{c['mut_nuc']}:{c['is_missing']}," for c in g["calls"][sp[i]:sp[i + 1]])
        assert s[2] == exp
    nwk = [l for l in open(d + "/flat.txt") if l.startswith("NEWICK\t")][0].split("\t")[1].strip()
    assert nwk == str(g["current_tree"]).strip()   # the reference's -p run writes the labelled input tree


def test_vcf_reader_ambiguity_counts_follow_the_reference_quirk(usher):
    """num_ambiguous (the -A sort key): the reference tests Mutation::mut_nuc for ambiguity for EVERY genotype, and for a
    reference call (GT 0) that field still holds the previous genotype's allele (src/mutation_annotated_tree.cpp:2246,
    2271).  A sequential restatement of exactly that gives the expected counts; along the sample order of the
    reference's own `-A` run (golden) they never decrease.  The reader splits the rows over host threads: same counts
    and sample lists for any split."""
    vcf = os.path.join(common.GOLDEN, "hostgold_samples.vcf")
    code = {"A": 1, "C": 2, "G": 4, "T": 8, "R": 5, "Y": 10, "S": 6, "W": 9, "K": 12, "M": 3, "B": 14, "D": 13, "H": 11,
            "V": 15, "N": 15}   # 'V' falls through to N in the reference's switch (missing break)
    names, exp, stale = [], [], 0
    for line in open(vcf):
        w = line.split()
        if len(w) > 1 and w[1] == "POS":
            names = w[9:]
            exp = [0] * len(names)
        elif names:
            alts = w[4].split(",")
            for k in range(len(names)):
                gt = w[9 + k]
                if gt[0].isdigit():
                    a = int(gt)
                    nuc = code.get(alts[a - 1][0], 15) if a > 0 else stale
                else:
                    nuc = 15
                exp[k] += 1 if nuc & (nuc - 1) else 0
                stale = nuc
    outs = []
    for nt in ("1", "2", "5"):
        d = tempfile.mkdtemp()
        subprocess.check_call([usher, "-i", PB, "-v", vcf, "--dump-flat", d + "/f.txt"], stderr=subprocess.DEVNULL,
                              env=dict(os.environ, UB200_HOST_THREADS=nt))
        outs.append([l for l in open(d + "/f.txt") if l[:2] in ("S\t", "A\t")])
    assert outs[0] == outs[1] == outs[2]
    got = {l.split("\t")[1]: int(l.split("\t")[2]) for l in outs[0] if l.startswith("A\t")}
    assert got == dict(zip(names, exp))
    g = common.load(os.path.join(common.GOLDEN, "hostgold.npz"))
    order = [l.split("\t")[0] for l in str(g["sort3__placement_stats.tsv"]).splitlines() if l]
    keys = [got[n] for n in order]
    assert keys == sorted(keys) and len(set(keys)) > 1


def test_pb_round_trip_is_byte_identical(usher):
    d = tempfile.mkdtemp()
    subprocess.check_call([usher, "-i", PB, "--resave", d + "/re.pb"], stderr=subprocess.DEVNULL)
    assert open(d + "/re.pb", "rb").read() == open(PB, "rb").read()
    subprocess.check_call([usher, "-i", PB, "--resave", d + "/re.pb.gz"], stderr=subprocess.DEVNULL)
    subprocess.check_call([usher, "-i", d + "/re.pb.gz", "--resave", d + "/re2.pb"], stderr=subprocess.DEVNULL)
    assert open(d + "/re2.pb", "rb").read() == open(PB, "rb").read()


@pytest.mark.skipif(not os.path.exists("/root/reference/scripts/testBranchLen2.nwk"), reason="reference not mounted")
def test_create_mat_mode_has_no_silent_cpu_path(usher, has_gpu):
    """Without a CUDA device `usher -t ... -v ...` stops with an error instead of falling back to a host assignment."""
    if has_gpu:
        pytest.skip("a GPU is visible")
    d = tempfile.mkdtemp()
    env = {k: v for k, v in os.environ.items() if k != "UB200_FS_HOST"}
    r = subprocess.run([usher, "-t", "/root/reference/scripts/testBranchLen2.nwk", "-v",
                        "/root/reference/scripts/testBranchLen2.vcf", "--dump-flat", d + "/f.txt"], capture_output=True, text=True, env=env)
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_cli_rejects_unsupported_and_missing_arguments(usher):
    assert subprocess.run([usher], capture_output=True).returncode != 0
    assert subprocess.run([usher, "-v", VCF], capture_output=True).returncode != 0
    r = subprocess.run([usher, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--load-mutation-annotated-tree" in r.stderr


@pytest.mark.gpu
def test_config1_sequential_placement_matches_reference(usher):
    g = common.load(os.path.join(common.GOLDEN, "config1.npz"))
    d = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", VCF, "-d", d], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(d + "/placement_stats.tsv").read() == str(g["placement_stats"])
    assert open(d + "/mutation-paths.txt").read() == str(g["mutation_paths"])
    assert open(d + "/final-tree.nh").read() == str(g["final_tree"])
    assert "The parsimony score for this tree is: 503" in r.stderr
    assert "Sample1" in r.stderr.split("multiple possibilities of parsimony-optimal placements:")[1]


@pytest.mark.gpu
def test_config1_no_add_and_per_node_scores_match_reference(usher):
    g = common.load(os.path.join(common.GOLDEN, "config1.npz"))
    d = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", VCF, "-d", d, "-n"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(d + "/placement_stats.tsv").read() == str(g["noadd_placement_stats"])
    assert open(d + "/final-tree.nh").read() == str(g["noadd_final_tree"])
    d2 = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", VCF, "-d", d2, "-p"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(d2 + "/current-tree.nh").read() == str(g["current_tree"])
    assert open(d2 + "/parsimony-scores.tsv").read() == str(g["parsimony_scores"])


@pytest.mark.gpu
def test_sorted_placement_and_saved_pb_reload(usher):
    d = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", VCF, "-d", d, "-s", "-o", d + "/out.pb.gz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    # the saved tree contains the five new samples and reloads cleanly
    d2 = tempfile.mkdtemp()
    subprocess.check_call([usher, "-i", d + "/out.pb.gz", "-v", VCF, "--dump-flat", d2 + "/flat.txt"], stderr=subprocess.DEVNULL)
    txt = open(d2 + "/flat.txt").read()
    assert all(f"Sample{i}" in txt for i in range(1, 6))   # as a node or as a member of a condensed node
    assert not [l for l in txt.splitlines() if l.startswith("S\t")]   # all five are already in the tree now


REF_TEST = "/root/reference/test"


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="reference test inputs are only mounted in the build container")
def test_build_mat_from_newick_and_vcf_matches_reference(usher):
    """`usher -t global_phylo.nh -v global_samples.vcf -o tree.pb` (Fitch-Sankoff per site + condense + save):
    same nodes, names, mutations, newick and condensed sets as the MAT the reference builds (config 1)."""
    d = tempfile.mkdtemp()
    # (no GPU in this test: the serial host restatement of the assignment is asked for by name; the GPU kernel has its
    # own parity tests in test_gpu_fitch_sankoff.py)
    r = subprocess.run([usher, "-t", REF_TEST + "/global_phylo.nh", "-v", REF_TEST + "/global_samples.vcf", "-o",
                        d + "/tree.pb", "-d", d], capture_output=True, text=True, env=dict(os.environ, UB200_FS_HOST="1"))
    assert r.returncode == 0 and "The parsimony score for this tree is: 500" in r.stderr
    subprocess.check_call([usher, "-i", d + "/tree.pb", "-v", VCF, "--dump-flat", d + "/a.txt"], stderr=subprocess.DEVNULL)
    subprocess.check_call([usher, "-i", PB, "-v", VCF, "--dump-flat", d + "/b.txt"], stderr=subprocess.DEVNULL)
    a, b = open(d + "/a.txt").read().splitlines(), open(d + "/b.txt").read().splitlines()
    assert [l for l in a if not l.startswith("C\t")] == [l for l in b if not l.startswith("C\t")]
    assert sorted(l for l in a if l.startswith("C\t")) == sorted(l for l in b if l.startswith("C\t"))


@pytest.mark.skipif(not os.path.exists("/root/reference/scripts/testBranchLen2.nwk"), reason="reference not mounted")
def test_fitch_sankoff_known_answer(usher):
    """scripts/testBranchLen2.*: the input branch lengths are the expected per-branch mutation counts."""
    d = tempfile.mkdtemp()
    subprocess.check_call([usher, "-t", "/root/reference/scripts/testBranchLen2.nwk", "-v",
                           "/root/reference/scripts/testBranchLen2.vcf", "--dump-flat", d + "/f.txt"], stderr=subprocess.DEVNULL,
                          env=dict(os.environ, UB200_FS_HOST="1"))
    nwk = [l for l in open(d + "/f.txt") if l.startswith("NEWICK\t")][0].split("\t")[1].strip()
    assert nwk == "((a:0,(b:0,(c:0,d:1)node_4:1)node_3:2,((e:0,f:1)node_6:3,g:0)node_5:4)node_2:5,h:0)node_1:0;"


HOST_RUNS = {
    "default": [], "sort1": ["-s"], "sort2": ["-S"], "sort1_reverse": ["-s", "-r"], "sort3": ["-A"],
    "max_uncertainty_2": ["-e", "2"], "max_parsimony_3": ["-E", "3"], "no_add": ["-n"], "uncondensed": ["-u"],
}


@pytest.mark.gpu
@pytest.mark.parametrize("run", sorted(HOST_RUNS))
def test_ambiguous_samples_host_outputs_match_reference(usher, run):
    """The host layer around the kernels (imputed mutations of ambiguous / N calls, excess lists, child vs sibling
    grafts, duplicated genotypes, sort pre-pass orders, -e / -E thresholds, --no-add, uncondensed output) against
    files the reference's own usher_common() wrote for 14 new samples with IUPAC codes and N runs on the config-1
    MAT (oracle/make_golden_host.py): byte for byte."""
    g = common.load(os.path.join(common.GOLDEN, "hostgold.npz"))
    d = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", os.path.join(common.GOLDEN, "hostgold_samples.vcf"), "-d", d] + HOST_RUNS[run],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    for f in ("placement_stats.tsv", "mutation-paths.txt", "final-tree.nh", "uncondensed-final-tree.nh"):
        key = f"{run}__{f}"
        if key in g.files:
            assert open(os.path.join(d, f)).read() == str(g[key]), (run, f)
    assert f"The parsimony score for this tree is: {int(g[run + '__parsimony'])}" in r.stderr or run == "no_add"


def test_flat_loader_round_trip_is_byte_identical(usher):
    """N1: parsimony.proto -> flat SoA -> parsimony.proto without Node objects reproduces the file byte for byte
    (plain and gzip-compressed), i.e. the flat loader sees the same nodes, names, rows and condensed sets."""
    d = tempfile.mkdtemp()
    subprocess.check_call([usher, "-i", PB, "--flat-resave", d + "/f.pb"], stderr=subprocess.DEVNULL)
    assert open(d + "/f.pb", "rb").read() == open(PB, "rb").read()
    subprocess.check_call([usher, "-i", PB, "--flat-resave", d + "/f.pb.gz"], stderr=subprocess.DEVNULL)
    subprocess.check_call([usher, "-i", d + "/f.pb.gz", "--flat-resave", d + "/g.pb"], stderr=subprocess.DEVNULL)
    assert open(d + "/g.pb", "rb").read() == open(PB, "rb").read()
    # the mutation lists are parsed by slices of the node range on host threads: same bytes for any split
    for nt in ("1", "3", "16"):
        subprocess.check_call([usher, "-i", PB, "--flat-resave", d + f"/t{nt}.pb"], stderr=subprocess.DEVNULL,
                              env=dict(os.environ, UB200_HOST_THREADS=nt))
        assert open(d + f"/t{nt}.pb", "rb").read() == open(PB, "rb").read()


@pytest.mark.gpu
@pytest.mark.parametrize("vcf,key", [(VCF, "noadd_placement_stats"), (os.path.join(common.GOLDEN, "hostgold_samples.vcf"), None)])
def test_flat_placement_equals_no_add_run(usher, vcf, key):
    """N1 end to end: pb -> flat SoA -> GPU -> records (no Node objects) gives the scores and numbers of optimal
    placements of the reference's --no-add run."""
    d = tempfile.mkdtemp()
    r = subprocess.run([usher, "-i", PB, "-v", vcf, "-d", d, "--place-flat"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = [l.split("\t")[:3] for l in open(d + "/flat-placements.tsv").read().splitlines() if not l.startswith("#")]
    exp_txt = str(common.load(os.path.join(common.GOLDEN, "config1.npz"))[key]) if key else \
        str(common.load(os.path.join(common.GOLDEN, "hostgold.npz"))["no_add__placement_stats.tsv"])
    exp = [l.split("\t")[:3] for l in exp_txt.splitlines() if l]
    assert got == exp


def _compat_binary():
    import shutil
    out = os.path.join(tempfile.mkdtemp(), "compat_main")
    host = os.path.join(build.CSRC, "host")
    subprocess.check_call([build.HOSTCXX, "-std=c++17", "-O1", "-I", build.INC, "-I", host,
                           os.path.join(common.HERE, "compat_main.cpp"), os.path.join(host, "mutation_annotated_tree.cpp"),
                           "-o", out, "-L", build.PKG, "-lusher_b200", f"-Wl,-rpath,{build.PKG}", "-lz"])
    return out


def test_compat_adapter_compiles_against_the_mat_api(usher):
    """include/usher_b200_compat.hpp (adapter for the matUtils / ripples callers of mapper2_body) builds against the MAT
    API and links the C ABI."""
    assert os.path.exists(_compat_binary())


@pytest.mark.gpu
def test_compat_adapter_with_dfs_tie_index(usher):
    """The other callers pass j = DFS index (src/matUtils/annotate.cpp:629): same score, num_best and optimal set as the
    reference's search, and the best node is the optimal node with the most leaves, then the LARGEST DFS index."""
    exe = _compat_binary()
    g = common.load(os.path.join(common.GOLDEN, "config1.npz"))
    r = subprocess.run([exe, PB, VCF], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1500:]
    rows = [l.split("\t") for l in r.stdout.splitlines()]
    parent = g["parent"]
    n = len(parent)
    nchild = np.bincount(parent[1:], minlength=n)
    leaves = (nchild == 0).astype(np.int64)
    for i in range(n - 1, 0, -1):
        leaves[parent[i]] += leaves[i]
    names = g["names"].tolist()
    assert [x[0] for x in rows] == g["snames"].tolist()
    for s, x in enumerate(rows):
        lo, hi = int(g["exp_best_set_ptr"][s]), int(g["exp_best_set_ptr"][s + 1])
        opt = g["exp_best_set"][lo:hi].astype(np.int64)
        assert int(x[2]) == int(g["exp_score"][s]) and int(x[3]) == int(g["exp_num_best"][s])
        assert [int(v) for v in x[6].split(",")] == opt.tolist()
        best = max(opt.tolist(), key=lambda v: (leaves[v], v))
        assert x[1] == names[best] and int(x[4]) == best


@pytest.mark.parametrize("seed,n_leaves,n_sites", [(77, 400, 120), (78, 60, 200)])
def test_create_mat_reader_and_host_assignment_vs_reference_built_mats(usher, seed, n_leaves, n_sites):
    """`usher -t tree.nh -v samples.vcf` on random trees with ambiguous / missing genotypes: the VCF site reader, the
    serial restatement of mapper_body (asked for by name: no GPU here) and the mutation lists they produce, against
    the MAT the reference itself builds from the same files (src/usher_mapper.cpp:6-161).  The GPU kernel goes through
    the same reader in tests/test_gpu_fitch_sankoff.py."""
    from oracle import ref
    from test_oracle import _random_fs_case
    if not ref.available():
        pytest.skip("oracle/_ref/libusher_ref.so missing")
    newick, vcf, *_ = _random_fs_case(seed, n_leaves, n_sites, p_amb=0.15)
    d = tempfile.mkdtemp()
    open(d + "/t.nh", "w").write(newick)
    open(d + "/v.vcf", "w").write(vcf)
    rt = ref.RefTree.from_newick_vcf(d + "/t.nh", d + "/v.vcf", False, 1)
    parent, row_ptr, muts, names = rt.export()
    rt.close()
    r = subprocess.run([usher, "-t", d + "/t.nh", "-v", d + "/v.vcf", "--dump-flat", d + "/flat.txt"], capture_output=True,
                       text=True, env=dict(os.environ, UB200_FS_HOST="1"))
    assert r.returncode == 0, r.stderr[-1500:]
    nodes = [l.rstrip("\n").split("\t") for l in open(d + "/flat.txt") if l.startswith("N\t")]
    assert [x[1] for x in nodes] == list(names)
    assert [x[2] for x in nodes] == ["" if p < 0 else names[p] for p in parent]
    rp = row_ptr.astype(np.int64)
    for i, x in enumerate(nodes):
        exp = "".join(f"{m['position']}:{m['ref_nuc']}:{m['par_nuc']}:{m['mut_nuc']}," for m in muts[rp[i]:rp[i + 1]])
        assert x[3] == exp, (i, x[3], exp)
