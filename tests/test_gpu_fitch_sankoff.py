"""GPU: the per-site Fitch-Sankoff assignment (ub200_fs_*, usher_b200/csrc/fs_kernel.cuh) that builds a MAT from a tree
and a VCF, against the restatement of the reference's mapper_body (oracle/fitch_sankoff.py, pinned against the reference
in tests/test_oracle.py) and, through the `usher -t ... -v ...` binary, against a MAT built by the reference itself."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import fitch_sankoff, ref
from test_oracle import _random_fs_case
from usher_b200 import build, capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n_leaves,n_sites", [(1, 5, 30), (2, 40, 60), (3, 300, 60), (4, 900, 40)])
def test_fitch_sankoff_vs_oracle(seed, n_leaves, n_sites):
    _, _, parent_bfs, ref_code, var_ptr, var_node, var_nuc = _random_fs_case(900 + seed, n_leaves, n_sites, p_amb=0.2)
    exp = fitch_sankoff.assign(parent_bfs, ref_code, var_ptr, var_node, var_nuc)
    got = capi.fitch_sankoff(parent_bfs, ref_code, var_ptr, var_node, var_nuc)
    for a, b in zip(got, exp):
        assert np.array_equal(np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64))


def test_single_node_and_empty_sites():
    got = capi.fitch_sankoff(np.array([-1], np.int32), np.array([2], np.uint8), np.array([0, 1], np.uint64),
                             np.array([0], np.uint32), np.array([8], np.uint8))
    assert [x.tolist() for x in got] == [[0], [0], [2], [3]]      # the root itself mutates G -> T
    got = capi.fitch_sankoff(np.array([-1, 0, 0], np.int32), np.zeros(0, np.uint8), np.array([0], np.uint64),
                             np.zeros(0, np.uint32), np.zeros(0, np.uint8))
    assert all(len(x) == 0 for x in got)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libusher_ref.so missing")
def test_cli_builds_the_reference_mat_on_the_gpu():
    """`usher -t tree.nh -v samples.vcf` (create-MAT mode) with the GPU assignment: same nodes, names and mutation lists
    as the MAT the reference builds from the same files with its own mapper_body."""
    build.build()
    newick, vcf, *_ = _random_fs_case(77, 400, 120, p_amb=0.15)
    d = tempfile.mkdtemp()
    open(d + "/t.nh", "w").write(newick)
    open(d + "/v.vcf", "w").write(vcf)
    rt = ref.RefTree.from_newick_vcf(d + "/t.nh", d + "/v.vcf", False, 1)
    parent, row_ptr, muts, names = rt.export()
    rt.close()
    r = subprocess.run([build.USHER, "-t", d + "/t.nh", "-v", d + "/v.vcf", "--dump-flat", d + "/flat.txt"], capture_output=True, text=True)
    assert r.returncode == 0 and "Fitch-Sankoff on the GPU" in r.stderr, r.stderr[-1500:]
    nodes = [l.rstrip("\n").split("\t") for l in open(d + "/flat.txt") if l.startswith("N\t")]
    assert [x[1] for x in nodes] == list(names)
    assert [x[2] for x in nodes] == ["" if p < 0 else names[p] for p in parent]
    rp = row_ptr.astype(np.int64)
    for i, x in enumerate(nodes):
        exp = "".join(f"{m['position']}:{m['ref_nuc']}:{m['par_nuc']}:{m['mut_nuc']}," for m in muts[rp[i]:rp[i + 1]])
        assert x[3] == exp, (i, x[3], exp)
