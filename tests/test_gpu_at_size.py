"""GPU: parity at the BASELINE configs' stated sizes (config 3: 2 M-node SARS-CoV-2-shaped MAT, 256 samples per
launch; configs 4/5: 10 M nodes, SNV / leaf-derived / ambiguous + N-run samples).  One full reference search of a
10 M-node tree is minutes of CPU, so at that size the reference's own mapper2_body (oracle/_ref, prebuilt .so that
travels to the GPU box) is run at the GPU's whole optimal set plus 10^5 random nodes per sample
(oracle/spotcheck.py): per-node scores, validity, optimal-set membership, num_best and the tie-break."""
import os

import numpy as np
import pytest

from oracle import ref, spotcheck
from usher_b200 import capi

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libusher_ref.so missing")]

THREADS = max(1, min(32, len(os.sched_getaffinity(0))))
FIELDS = ("score", "best_node", "best_j", "num_best", "has_unique", "best_num_leaves")


@pytest.fixture(scope="module")
def c3():
    s = capi.Synth(2_000_000, 1.2, 29_903, capi.Synth.SC2, 20260928)
    m = capi.Mat.from_flat_struct(s.flat)
    rt = ref.RefTree.from_flat(*s.arrays())
    yield s, m, rt
    rt.close(); m.close(); s.close()


@pytest.fixture(scope="module")
def c4():
    s = capi.Synth(10_000_000, 30.0, 30_000, capi.Synth.UNIFORM, 20260929)
    m = capi.Mat.from_flat_struct(s.flat)
    rt = ref.RefTree.from_flat(*s.arrays())
    yield s, m, rt
    rt.close(); m.close(); s.close()


def test_c3_256_per_launch_vs_reference(c3):
    """Config 3 at size: 512 leaf-derived samples, 256 per launch (8 sample groups per pass, 2 passes)."""
    s, m, rt = c3
    sp, sc, _ = s.samples(512, capi.Synth.LEAF, 41)
    m.set_pass_samples(256)
    wide = m.place_batch(sp, sc, best_set=True)
    t = m.timing()
    assert t.score_launches == 2
    m.set_pass_samples(32)
    narrow = m.place_batch(sp, sc, best_set=True)
    for k in FIELDS:
        assert np.array_equal(wide["placements"][k], narrow["placements"][k]), k
    assert np.array_equal(wide["best_set"], narrow["best_set"])
    # whole reference searches (two passes, optimal sets) of samples from different groups
    ids = [0, 37, 255, 256, 300, 511]
    for i in ids:
        o = rt.search(sp[i:i + 2] - sp[i], sc[int(sp[i]):int(sp[i + 1])], m.n, threads=THREADS)
        p = wide["placements"][i]
        assert (int(o["score"][0]), int(o["best_dfs"][0]), int(o["best_j"][0]), int(o["num_best"][0]), int(o["has_unique"][0])) == \
               (int(p["score"]), int(p["best_node"]), int(p["best_j"]), int(p["num_best"]), int(p["has_unique"])), i
        lo, hi = int(wide["best_set_ptr"][i]), int(wide["best_set_ptr"][i + 1])
        assert np.array_equal(o["best_set"], wide["best_set"][lo:hi]), i
        assert np.array_equal(o["best_set_unique"], wide["best_set_unique"][lo:hi]), i
    # per-node scores on a node subset for other samples (and the ambiguous family on this shape)
    m.set_pass_samples(256)
    spotcheck.spot_check(m, rt, sp, sc, [5, 100, 400], n_random=50_000, seed=3, threads=THREADS, label="c3 leaf")
    sp2, sc2, _ = s.samples(64, capi.Synth.AMBIG, 42)
    spotcheck.spot_check(m, rt, sp2, sc2, [1, 33], n_random=50_000, seed=4, threads=THREADS, label="c3 ambig")


@pytest.mark.parametrize("family,label", [(capi.Synth.AMBIG, "c5 ambig"), (capi.Synth.LEAF, "c4 leaf"),
                                          (capi.Synth.SNV40, "c4 snv40")])
def test_c4_size_vs_reference_spot_check(c4, family, label):
    """Configs 4/5 at size (10 M nodes): whole optimal sets, per-node scores at the sets + 10^5 random nodes, tie-breaks."""
    s, m, rt = c4
    sp, sc, _ = s.samples(64, family, 500 + family)
    m.set_pass_samples(32)
    r = spotcheck.spot_check(m, rt, sp, sc, [0, 40], n_random=100_000, seed=10 + family, threads=THREADS, label=label)
    assert r["nodes_checked"] >= 200_000


def test_c4_size_eight_groups_per_pass(c4):
    """256 ambiguous samples in ONE pass of 8 groups over the 10 M-node tree equal the 32-per-pass results."""
    s, m, rt = c4
    sp, sc, _ = s.samples(256, capi.Synth.AMBIG, 777)
    m.set_pass_samples(256)
    a = m.place_batch(sp, sc, best_set=True)
    assert m.timing().score_launches == 1
    m.set_pass_samples(32)
    b = m.place_batch(sp, sc, best_set=True)
    for k in FIELDS:
        assert np.array_equal(a["placements"][k], b["placements"][k]), k
    assert np.array_equal(a["best_set"], b["best_set"]) and np.array_equal(a["best_set_ptr"], b["best_set_ptr"])
    assert (a["placements"]["num_best"] > 1).any()
