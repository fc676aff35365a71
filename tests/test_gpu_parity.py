"""GPU: parity of the CUDA path (through the C ABI) with the oracle.  Bit-exact (integer work): best score,
best node, tie index, number of optimal placements, sibling/child flag, the full optimal set and every
per-node score."""
import os

import numpy as np
import pytest

import common
import small_synth
from oracle import port, ref
from usher_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    n = capi.lib().ub200_device_count()
    assert n > 0, "no CUDA device: the product has no CPU path, -m gpu tests must run on the B200 box"


@pytest.mark.parametrize("path", common.golden_cases(), ids=lambda p: p.split("/")[-1])
def test_golden_vectors(path):
    g = common.load(path)
    m = capi.Mat(g["parent"], g["row_ptr"], g["muts"])
    res = m.place_batch(g["s_ptr"], g["calls"], node_scores=True, best_set=True)
    common.assert_matches_expected(g, common.placements_to_dict(res), path)
    m.close()


@pytest.mark.parametrize("path", common.golden_cases()[::3], ids=lambda p: p.split("/")[-1])
def test_golden_vectors_many_tiles(path, monkeypatch):
    """Same vectors with the tree cut into many small tiles: every tile re-seeds its root path from seed segments."""
    monkeypatch.setenv("UB200_MIN_TILE", "96")
    g = common.load(path)
    m = capi.Mat(g["parent"], g["row_ptr"], g["muts"])
    assert m.info.n_tiles > 1 or len(g["parent"]) < 64
    res = m.place_batch(g["s_ptr"], g["calls"], best_set=True)
    common.assert_matches_expected(g, common.placements_to_dict(res), path)
    m.close()


@pytest.mark.parametrize("pass_samples", [32, 64, 256])
def test_pass_width_does_not_change_results(pass_samples):
    g = common.load(common.GOLDEN + "/random_07.npz")
    m = capi.Mat(g["parent"], g["row_ptr"], g["muts"])
    m.set_pass_samples(pass_samples)
    res = m.place_batch(g["s_ptr"], g["calls"], best_set=True)
    common.assert_matches_expected(g, common.placements_to_dict(res), f"pass={pass_samples}")
    m.close()


@pytest.mark.parametrize("seed", range(6))
def test_random_vs_port(seed):
    n = [3000, 8000, 20000][seed % 3]
    L = [60, 500, 4000][(seed + 1) % 3]
    mu = [1.0, 4.0, 9.0][seed % 3]
    shape = ["uniform", "chain", "uniform"][seed % 3]
    parent, row_ptr, muts, refg = small_synth.random_mat(900 + seed, n, L, mu, shape=shape)
    s_ptr, calls = small_synth.random_samples(950 + seed, parent, row_ptr, muts, refg, 70)
    pt = port.PortTree(parent, row_ptr, muts)
    q = pt.search(s_ptr, calls)
    m = capi.Mat(parent, row_ptr, muts)
    res = common.placements_to_dict(m.place_batch(s_ptr, calls, best_set=True))
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(res[k]).astype(np.int64), q[k].astype(np.int64)), (seed, k)
    m.close()
    pt.close()


def test_deep_tree_spills_the_stack():
    """A 300-level chain exceeds the in-shared-memory stack and exercises the HBM spill path."""
    parent, row_ptr, muts, refg = small_synth.random_mat(77, 2500, 300, 2.0, shape="chain")
    s_ptr, calls = small_synth.random_samples(78, parent, row_ptr, muts, refg, 40)
    pt = port.PortTree(parent, row_ptr, muts)
    q = pt.search(s_ptr, calls)
    m = capi.Mat(parent, row_ptr, muts)
    assert m.info.max_level > 64
    res = common.placements_to_dict(m.place_batch(s_ptr, calls, best_set=True, node_scores=True))
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(res[k]).astype(np.int64), q[k].astype(np.int64)), k
    assert np.array_equal(res["node_scores"], pt.search(s_ptr, calls, per_node=True)["node_scores"])
    m.close()
    pt.close()


def test_config2_shape_vs_reference_and_batch_invariance():
    """BASELINE config 2 (100k-node MAT, ~30 mutations/node, 30 kb genome).  (a) a few samples against the
    reference's own mapper2_body when oracle/_ref is present; (b) size-independent property: the placement of a
    sample does not depend on which other samples share its launch (different groups => different position
    bitmaps and tables)."""
    s = capi.Synth(100_000, 30.0, 30_000, capi.Synth.UNIFORM, 20260927)
    p, r, mu = s.arrays()
    m = capi.Mat.from_flat_struct(s.flat)
    out = {}
    for fam in (capi.Synth.SNV40, capi.Synth.LEAF, capi.Synth.AMBIG):
        sp, sc, so = s.samples(256, fam, 99 + fam)
        a = m.place_batch(sp, sc)["placements"]
        perm = np.random.default_rng(fam).permutation(256)
        lens = np.diff(sp.astype(np.int64))
        sp2 = np.concatenate([[0], np.cumsum(lens[perm])]).astype(np.uint64)
        sc2 = np.concatenate([sc[int(sp[i]):int(sp[i + 1])] for i in perm])
        b = m.place_batch(sp2, sc2)["placements"]
        for k in ("score", "best_node", "best_j", "num_best", "has_unique"):
            assert np.array_equal(a[k][perm], b[k]), (fam, k)
        assert np.all(a["score"] >= 0)
        out[fam] = (sp, sc, a)
    if ref.available():
        rt = ref.RefTree.from_flat(p, r, mu)
        for fam in out:
            sp, sc, a = out[fam]
            k = 3
            o = rt.search(sp[: k + 1], sc[: int(sp[k])], len(p), threads=os.cpu_count() or 1, want_set=False)
            for key, mine in (("score", "score"), ("best_dfs", "best_node"), ("best_j", "best_j"),
                              ("num_best", "num_best"), ("has_unique", "has_unique")):
                assert np.array_equal(o[key].astype(np.int64), a[mine][:k].astype(np.int64)), (fam, key)
        rt.close()
    m.close()
    s.close()


@pytest.mark.parametrize("seed,shape,fam", [(1, capi.Synth.UNIFORM, capi.Synth.SNV40), (2, capi.Synth.SC2, capi.Synth.LEAF),
                                            (3, capi.Synth.UNIFORM, capi.Synth.AMBIG), (4, capi.Synth.SC2, capi.Synth.SNV40)])
def test_optimal_sets_vs_port_many_tiles(seed, shape, fam, monkeypatch):
    """The second (collect) pass runs with the final best as its bound from the first block on, so any block whose
    exact lower bound is wrong loses optimal nodes: whole optimal sets against the port on trees cut into many
    tiles, shallow and deep optima, all three sample families."""
    monkeypatch.setenv("UB200_MIN_TILE", "512")
    s = capi.Synth(40_000, [6.0, 2.0, 12.0, 1.2][seed - 1], 4000, shape, 777 + seed)
    p, r, mu = s.arrays()
    sp, sc, _ = s.samples(160, fam, 31 + seed)
    m = capi.Mat.from_flat_struct(s.flat)
    assert m.info.n_tiles > 50
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    pt = port.PortTree(p, r, mu)
    q = pt.search(sp, sc)
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), q[k].astype(np.int64)), (seed, k)
    pt.close(); m.close(); s.close()


@pytest.mark.parametrize("genome_len", [200_000, 3_000_000])
def test_long_genomes_vs_port(genome_len):
    """Genomes whose position bitmap does not fit shared memory (global bitmap path) and, at 3 Mb, positions beyond
    2^21 (wide stream words)."""
    s = capi.Synth(6_000, 5.0, genome_len, capi.Synth.UNIFORM, 4242)
    p, r, mu = s.arrays()
    sp, sc, _ = s.samples(48, capi.Synth.AMBIG, 9)
    m = capi.Mat.from_flat_struct(s.flat)
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    pt = port.PortTree(p, r, mu)
    q = pt.search(sp, sc)
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), q[k].astype(np.int64)), (genome_len, k)
    pt.close(); m.close(); s.close()


def test_resident_api_matches_batch_api_and_times():
    g = common.load(common.GOLDEN + "/random_15.npz")
    m = capi.Mat(g["parent"], g["row_ptr"], g["muts"])
    S = m.upload(g["s_ptr"], g["calls"])
    S.place()
    a = S.download()
    t = m.timing()
    assert t.score_launches >= 1 and t.score_ms > 0 and t.score_bytes > 0
    b = m.place_batch(g["s_ptr"], g["calls"])["placements"]
    assert np.array_equal(a, b)
    S.close()
    m.close()


def test_bad_samples_are_rejected():
    g = common.load(common.GOLDEN + "/random_04.npz")
    m = capi.Mat(g["parent"], g["row_ptr"], g["muts"])
    calls = np.array([(9, 1, 1, 2, 0), (5, 1, 1, 2, 0)], capi.MUT_DTYPE)
    with pytest.raises(capi.UB200Error) as e:
        m.place_batch(np.array([0, 2], np.uint64), calls)
    assert e.value.code == -5
    m.close()


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_config5_ambiguity_and_n_runs_vs_reference():
    """BASELINE config 5 at reduced size: IUPAC-widened calls and N runs (mean length 200) over mutated path
    positions; score, best node, tie index, num_best and sibling/child flag bit-exact against the reference's own
    mapper2_body (many-way ties included)."""
    s = capi.Synth(30_000, 8.0, 30_000, capi.Synth.SC2, 20260929)
    p, r, mu = s.arrays()
    sp, sc, _ = s.samples(96, capi.Synth.AMBIG, 55)
    m = capi.Mat.from_flat_struct(s.flat)
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    rt = ref.RefTree.from_flat(p, r, mu)
    o = rt.search(sp, sc, len(p), threads=os.cpu_count() or 1)
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), o[k].astype(np.int64)), k
    assert (o["num_best"] > 1).any()
    rt.close(); m.close(); s.close()


@pytest.fixture(scope="module")
def sharing_case():
    s = capi.Synth(40_000, 6.0, 4000, capi.Synth.UNIFORM, 901)
    p, r, mu = s.arrays()
    sp, sc, _ = s.samples(200, capi.Synth.AMBIG, 17)
    pt = port.PortTree(p, r, mu)
    q = pt.search(sp, sc)
    pt.close()
    yield s, sp, sc, q
    s.close()


@pytest.mark.parametrize("nc,pass_samples", [(1, 32), (2, 64), (3, 96), (3, 256), (2, 96), (0, 256), (1, 256)])
def test_scan_sharing_does_not_change_results(sharing_case, nc, pass_samples, monkeypatch):
    """Several sample groups sharing one scan of the stream (union bitmap, NC consumers per scanner, foreign hits
    dropped by the empty sample mask) and the dense (transposed) form of the hit phase (N runs over a 4 kb genome:
    many samples call the same position) against the port: whole optimal sets, 7 groups with a short last one."""
    monkeypatch.setenv("UB200_MIN_TILE", "512")
    s, sp, sc, q = sharing_case
    m = capi.Mat.from_flat_struct(s.flat)
    m.set_pass_samples(pass_samples)
    m.set_scan_sharing(nc)
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), q[k].astype(np.int64)), (nc, pass_samples, k)
    m.close()


@pytest.mark.parametrize("nc", [1, 3])
def test_tiny_blocks_many_segments_per_step(nc, monkeypatch):
    """SARS-CoV-2-like rows (about one mutation per node, many empty rows): a scanner step of 512 words spans a dozen
    block segments, empty segments included."""
    monkeypatch.setenv("UB200_MIN_TILE", "256")
    s = capi.Synth(60_000, 0.7, 3000, capi.Synth.SC2, 902)
    p, r, mu = s.arrays()
    sp, sc, _ = s.samples(128, capi.Synth.LEAF, 19)
    m = capi.Mat.from_flat_struct(s.flat)
    m.set_pass_samples(96 if nc == 3 else 32)
    m.set_scan_sharing(nc)
    got = common.placements_to_dict(m.place_batch(sp, sc, best_set=True))
    pt = port.PortTree(p, r, mu)
    q = pt.search(sp, sc)
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), q[k].astype(np.int64)), (nc, k)
    pt.close(); m.close(); s.close()


@pytest.mark.parametrize("n_calls", [16_300, 17_500])
def test_many_calls_near_the_int16_stack_limit(n_calls):
    """A sample that carries every mutation of a deep chain: its path corrections fall by 2 per call and reach
    -2 * calls.  Up to 16383 calls they fit the streaming kernel's int16 stack rows; longer call lists must switch to
    the first-generation kernel — both against the port."""
    parent, row_ptr, muts, refg = small_synth.random_mat(4242, 500, 400_000, 300.0, p_masked=0.0, shape="chain")
    lvl = np.zeros(len(parent), np.int64)
    for i in range(1, len(parent)):
        lvl[i] = lvl[parent[i]] + 1
    deep = int(np.argmax(lvl))
    g = small_synth.genotype(parent, row_ptr, muts, deep)
    sites = sorted(p for p, nuc in g.items() if nuc != refg[p])
    assert len(sites) >= n_calls, len(sites)
    calls = np.zeros(n_calls + 3, capi.MUT_DTYPE)
    for i, p in enumerate(sites[:n_calls]):
        calls[i] = (p, int(refg[p]), int(refg[p]), g[p], 0)
    # a second, short sample in the same group
    short = [p for p in sites[n_calls - 50:n_calls - 47]]
    for i, p in enumerate(short):
        calls[n_calls + i] = (p, int(refg[p]), int(refg[p]), g[p], 0)
    s_ptr = np.array([0, n_calls, n_calls + 3], np.uint64)
    m = capi.Mat(parent, row_ptr, muts)
    got = common.placements_to_dict(m.place_batch(s_ptr, calls, best_set=True))
    pt = port.PortTree(parent, row_ptr, muts)
    q = pt.search(s_ptr, calls)
    for k in ("score", "best_dfs", "best_j", "num_best", "has_unique", "best_set", "best_set_unique"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), q[k].astype(np.int64)), (n_calls, k)
    pt.close(); m.close()


def test_multi_device_shards_equal_single_device():
    """ub200_multi_place_batch: the batch cut into contiguous shards over several replicas (here two replicas on the
    same GPU, so that the shard / gather code runs on a one-GPU box too; every visible GPU when there are more) gives
    the records and optimal sets of a single-device call."""
    s = capi.Synth(60_000, 8.0, 8000, capi.Synth.UNIFORM, 515)
    sp, sc, _ = s.samples(333, capi.Synth.AMBIG, 21)
    one = capi.Mat.from_flat_struct(s.flat)
    a = one.place_batch(sp, sc, best_set=True)
    ndev = capi.lib().ub200_device_count()
    for devices in ([0, 0], list(range(ndev)) + [0]):
        mm = capi.MultiMat(s.flat, devices=devices)
        assert mm.size == len(devices)
        mm.set_pass_samples(96)
        b = mm.place_batch(sp, sc, best_set=True)
        assert np.array_equal(a["placements"], b["placements"])
        assert np.array_equal(a["best_set_ptr"], b["best_set_ptr"]) and np.array_equal(a["best_set"], b["best_set"])
        assert np.array_equal(a["best_set_unique"], b["best_set_unique"])
        mm.close()
    one.close(); s.close()
