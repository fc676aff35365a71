"""CPU: host logic of the product.  The C++ derivation (usher_b200/csrc/derive.cpp) is run through the
host-only ub200_debug_derive hook and its arrays are pushed through a plain-Python model of the kernel's
arithmetic (tests/kernel_model.py); the result must reproduce the golden vectors.  Also: the C-ABI library
loads and exports every symbol include/usher_b200.h declares, input validation rejects malformed trees and
samples, and without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import re
import os

import numpy as np
import pytest

import common
import kernel_model
import small_synth
from usher_b200 import capi


@pytest.mark.parametrize("path", common.golden_cases(), ids=lambda p: p.split("/")[-1])
def test_derivation_and_closed_form_match_golden(path):
    g = common.load(path)
    d = capi.debug_derive(g["parent"], g["row_ptr"], g["muts"], target_tiles=7, min_tile_cost=96)
    n = len(g["parent"])
    if n > 600:  # keep the pure-Python model quick
        sel = np.arange(0, len(g["s_ptr"]) - 1)[:12]
    else:
        sel = np.arange(0, len(g["s_ptr"]) - 1)
    s_ptr = g["s_ptr"].astype(np.int64)
    sub_ptr = np.concatenate([[0], np.cumsum(np.diff(s_ptr)[sel])]).astype(np.uint64)
    sub_calls = np.concatenate([g["calls"][s_ptr[s]:s_ptr[s + 1]] for s in sel]) if len(sel) else g["calls"][:0]
    res, ns = kernel_model.place(d, sub_ptr, sub_calls, per_node=True)
    for i, s in enumerate(sel):
        r = res[i]
        assert (r["score"], r["best_node"], r["best_j"], r["num_best"], r["has_unique"]) == (
            int(g["exp_score"][s]), int(g["exp_best_dfs"][s]), int(g["exp_best_j"][s]), int(g["exp_num_best"][s]),
            int(g["exp_has_unique"][s])), (path, s)
        a, b = int(g["exp_best_set_ptr"][s]), int(g["exp_best_set_ptr"][s + 1])
        assert [x for x, _ in r["optimal"]] == g["exp_best_set"][a:b].tolist()
        assert [h for _, h in r["optimal"]] == g["exp_best_set_unique"][a:b].tolist()
        assert np.array_equal(ns[i], g["exp_node_scores"][s])
    # the segment layout of k_score3, interpreted the way that kernel does
    res3 = kernel_model.place3(d, sub_ptr, sub_calls)
    assert len(d["tile3_start"]) > 2 or n < 64
    for i, s in enumerate(sel):
        r = res3[i]
        assert (r["score"], r["best_node"], r["best_j"], r["num_best"], r["has_unique"]) == (
            int(g["exp_score"][s]), int(g["exp_best_dfs"][s]), int(g["exp_best_j"][s]), int(g["exp_num_best"][s]),
            int(g["exp_has_unique"][s])), (path, s, "k_score3 layout")
        a, b = int(g["exp_best_set_ptr"][s]), int(g["exp_best_set_ptr"][s + 1])
        assert [x for x, _ in r["optimal"]] == g["exp_best_set"][a:b].tolist()


def test_derive_structure():
    parent, row_ptr, muts, _ = small_synth.random_mat(1, 400, 100, 3.0)
    d = capi.debug_derive(parent, row_ptr, muts, target_tiles=16)
    n = d["n"]
    # tiles partition the DFS order; ancestor chains are root-first paths
    ts = d["tile_start"]
    assert ts[0] == 0 and ts[-1] == n and np.all(np.diff(ts.astype(np.int64)) > 0)
    for t in range(len(ts) - 1):
        chain = d["anc"][d["anc_ptr"][t]:d["anc_ptr"][t + 1]].tolist()
        exp = []
        a = parent[ts[t]]
        while a >= 0:
            exp.append(int(a))
            a = parent[a]
        assert chain == exp[::-1]
    # tiekey is a permutation; key order = (num_leaves desc, j desc)
    order = d["key_to_node"]
    assert sorted(order.tolist()) == list(range(n))
    k = [(int(d["num_leaves"][v]), int(d["tie_index"][v])) for v in order]
    assert k == sorted(k, reverse=True)
    # BFS index is a permutation starting at the root; levels consistent
    assert d["tie_index"][0] == 0 and sorted(d["tie_index"].tolist()) == list(range(n))
    assert all(d["level"][i] == d["level"][parent[i]] + 1 for i in range(1, n))


def test_header_declares_only_exported_symbols():
    hdr = open(os.path.join(os.path.dirname(common.HERE), "include", "usher_b200.h")).read()
    declared = set(re.findall(r"\b(ub200_[a-z_0-9]+)\s*\(", hdr))
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/usher_b200.h but not exported"
    assert declared >= set(capi.EXPORTS)
    assert L.ub200_abi_version() == 1
    S = capi.synth_lib()
    shdr = open(os.path.join(os.path.dirname(common.HERE), "include", "usher_b200_synth.h")).read()
    for name in set(re.findall(r"\b(ub200_synth_[a-z_0-9]+)\s*\(", shdr)):
        assert hasattr(S, name)


def _derive_err(parent, row_ptr, muts):
    with pytest.raises(capi.UB200Error) as e:
        capi.debug_derive(parent, row_ptr, muts)
    return e.value.code


def test_validation_rejects_bad_trees():
    M = capi.MUT_DTYPE
    ok_m = np.array([(5, 1, 1, 2, 0)], M)
    assert _derive_err([-1, 0, 0, 1], [0, 0, 0, 0, 1], ok_m) == -2          # node 3 under node 1 after node 2: not pre-order
    assert _derive_err([0, 0], [0, 0, 1], ok_m) == -2                        # root must have parent -1
    assert _derive_err([-1, 0], [0, 0, 1], np.array([(5, 1, 1, 3, 0)], M)) == -3   # multi-bit tree allele
    assert _derive_err([-1, 0], [0, 0, 2], np.array([(9, 1, 1, 2, 0), (5, 1, 1, 2, 0)], M)) == -4  # unsorted row
    assert _derive_err([-1, 0], [0, 0, 1], np.array([(1 << 27, 1, 1, 2, 0)], M)) == -4  # position too large
    assert _derive_err([-1, 0, 0], [0, 0, 1, 2], np.array([(5, 1, 1, 2, 0), (5, 2, 2, 4, 0)], M)) == -1  # ref disagreement


def test_no_device_fails_loudly(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    parent, row_ptr, muts, _ = small_synth.random_mat(3, 20, 30, 1.0)
    with pytest.raises(capi.UB200Error) as e:
        capi.Mat(parent, row_ptr, muts)
    assert e.value.code == capi.E_NO_DEVICE


def test_synth_generator_is_a_valid_mat():
    s = capi.Synth(3000, 4.0, 2000, capi.Synth.SC2, 11)
    p, r, m = s.arrays()
    d = capi.debug_derive(p, r, m)
    assert d["n"] == 3000 and d["m"] == len(m)
    # par_nuc recorded by the generator == the path state the derivation reconstructs
    prev = (d["mutw"] >> 2) & 3
    assert np.array_equal(1 << prev, m["par_nuc"])
    for fam in (0, 1, 2):
        sp, sc, so = s.samples(50, fam, 5)
        for i in range(50):
            pos = sc["position"][int(sp[i]):int(sp[i + 1])]
            assert np.all(np.diff(pos) > 0)
    s.close()


def test_segment_layout_structure():
    """k_score3 layout invariants, straight from the arrays: every mutation of a block appears exactly once in the
    block's segment with its node's lane; a tile's seed segments hold exactly the rows of the root path of its
    first node; segment sizes, tile starts, in-block ancestor masks and open flags are what the kernel assumes."""
    parent, row_ptr, muts, _ = small_synth.random_mat(5, 1500, 400, 3.0)
    d = capi.debug_derive(parent, row_ptr, muts, target_tiles=40, min_tile_cost=300)
    n, hdr, stream, nar = d["n"], d["hdr3"], d["stream"], d["narrow3"]
    pos_of = lambda w: ((int(w) >> (16 if nar else 14)) << 5) | (int(w) & 31)
    ts, w0, lvl, sseg, send, bw = (d[k] for k in ("tile3_start", "tile3_w0", "tile3_lvl", "tile3_sseg", "seed_end", "blk_words"))
    assert ts[0] == 0 and ts[-1] == n and all(int(t) % 32 == 0 for t in ts[:-1]) and len(ts) > 4
    level = d["level"]
    old = d["mutw"]          # node-major words of the first layout: pos<<6 | ref<<4 | prev<<2 | mut
    row32 = d["row32"]
    def row_words(i, lane):
        return sorted((int(w) >> 6, lane, (int(w) >> 2) & 3, int(w) & 3) for w in old[int(row32[i]):int(row32[i + 1])])
    def seg_words(o0, o1):
        out = []
        for w in stream[o0:o1]:
            w = int(w)
            if pos_of(w) == d["L"]:
                continue   # pad
            out.append((pos_of(w), (w >> 9) & 31, (w >> 7) & 3, (w >> 5) & 3))
        return sorted(out)
    for t in range(len(ts) - 1):
        n0, n1 = int(ts[t]), int(ts[t + 1])
        o0 = int(w0[t]) * 256
        assert int(lvl[t]) == int(level[n0])
        chain = []
        a = parent[n0]
        while a >= 0:
            chain.append(int(a)); a = parent[a]
        chain = chain[::-1]
        assert int(sseg[t + 1]) - int(sseg[t]) == (len(chain) + 31) // 32
        for g in range((len(chain) + 31) // 32):
            o1 = int(send[int(sseg[t]) + g]) * 4
            exp = sorted(x for l in range(32 * g, min(len(chain), 32 * g + 32)) for x in row_words(chain[l], l & 31))
            assert seg_words(o0, o1) == exp
            o0 = o1
        for blk in range(n0, n1, 32):
            o1 = o0 + int(bw[blk >> 5])
            assert o0 % 4 == 0 and o1 % 4 == 0
            exp = sorted(x for i in range(blk, min(blk + 32, n1)) for x in row_words(i, i & 31))
            assert seg_words(o0, o1) == exp
            o0 = o1
        assert o0 <= int(w0[t + 1]) * 256
    # in-block ancestor masks and open flags
    kids_beyond = np.zeros(n, bool)
    for i in range(1, n):
        a = parent[i]
        while a >= 0:
            if (a | 31) < i:
                kids_beyond[a] = True
            a = parent[a]
    for i in range(n):
        am, a = 0, parent[i]
        while a >= 0 and a >= (i & ~31):
            am |= 1 << (a & 31); a = parent[a]
        assert int(hdr[i]["tiekey"]) == am
        assert bool(int(hdr[i]["level_flags"]) & 32) == bool(kids_beyond[i])


@pytest.mark.parametrize("seed", range(8))
def test_random_trees_segment_model_vs_port(seed):
    """Fresh randomized trees (chains deeper than 64 levels, stars, masked mutations, tiny genomes with position
    collisions) cut into many tiles: the derivation + the k_score3 layout model against the oracle port."""
    from oracle import port
    n = [40, 150, 400, 700][seed % 4]
    L = [25, 90, 300][seed % 3]
    mu = [1.0, 3.0, 6.0][(seed // 2) % 3]
    shape = ["uniform", "chain", "star"][seed % 3]
    parent, row_ptr, muts, refg = small_synth.random_mat(7100 + seed, n, L, mu, shape=shape)
    s_ptr, calls = small_synth.random_samples(7200 + seed, parent, row_ptr, muts, refg, 6)
    d = capi.debug_derive(parent, row_ptr, muts, target_tiles=64, min_tile_cost=[64, 200][seed % 2])
    res3 = kernel_model.place3(d, s_ptr, calls)
    pt = port.PortTree(parent, row_ptr, muts)
    q = pt.search(s_ptr, calls)
    for i, r in enumerate(res3):
        assert (r["score"], r["best_node"], r["best_j"], r["num_best"], r["has_unique"]) == (
            int(q["score"][i]), int(q["best_dfs"][i]), int(q["best_j"][i]), int(q["num_best"][i]),
            int(q["has_unique"][i])), (seed, i)
        a, b = int(q["best_set_ptr"][i]), int(q["best_set_ptr"][i + 1])
        assert [x for x, _ in r["optimal"]] == q["best_set"][a:b].tolist()
    pt.close()


def test_derivation_does_not_depend_on_the_host_thread_count(monkeypatch):
    """The derivation runs on host threads (chunks of the node range: checks, tie-break ranks, path states, tile
    pieces): every derived array is identical for 1, 3 and 7 threads, on trees with masked rows and reversions."""
    import small_synth
    for seed, n, L, mu, shape in ((11, 900, 300, 4.0, "uniform"), (12, 2500, 60, 2.0, "chain"), (13, 1, 10, 1.0, "uniform")):
        parent, row_ptr, muts, _ = small_synth.random_mat(seed, n, L, mu, shape=shape)
        outs = []
        for nt in (1, 3, 7):
            monkeypatch.setenv("UB200_HOST_THREADS", str(nt))
            outs.append(capi.debug_derive(parent, row_ptr, muts, target_tiles=16, min_tile_cost=64))
        for o in outs[1:]:
            assert sorted(o) == sorted(outs[0])
            for k, v in outs[0].items():
                assert np.array_equal(np.asarray(v), np.asarray(o[k])), (seed, k)


def test_leaf_counts_bfs_index_and_tie_order_against_a_plain_restatement(monkeypatch):
    """num_leaves (mutation_annotated_tree.cpp:866-879), the breadth-first index (:1225-1251, children in DFS order) and
    the tie-break rank (usher_mapper.cpp:483-486) are computed by chunks of the node range: compare them with a
    sequential restatement, also for a caller's tie_index that repeats (equal keys keep node order)."""
    import small_synth
    for seed, n, shape in ((21, 1500, "uniform"), (22, 1200, "chain"), (23, 2, "uniform")):
        parent, row_ptr, muts, _ = small_synth.random_mat(seed, n, 80, 2.0, shape=shape)
        n = len(parent)
        leaves = np.zeros(n, np.int64)
        kids = [[] for _ in range(n)]
        for i in range(1, n):
            kids[parent[i]].append(i)
        for i in range(n - 1, -1, -1):
            if not kids[i]:
                leaves[i] = 1
            if i:
                leaves[parent[i]] += leaves[i]
        bfs, q = np.zeros(n, np.int64), [0]
        for h, u in enumerate(q):
            bfs[u] = h
            q.extend(kids[u])
        rep = np.random.default_rng(seed).integers(0, 5, n).astype(np.uint32)   # a tie_index with many repeats
        for nt in (1, 4):
            monkeypatch.setenv("UB200_HOST_THREADS", str(nt))
            d = capi.debug_derive(parent, row_ptr, muts, target_tiles=16, min_tile_cost=64)
            assert np.array_equal(d["num_leaves"], leaves) and np.array_equal(d["tie_index"], bfs)
            d = capi.debug_derive(parent, row_ptr, muts, tie_index=rep, target_tiles=16, min_tile_cost=64)
            exp = sorted(range(n), key=lambda v: (-int(leaves[v]), -int(rep[v]), v))
            assert d["key_to_node"].tolist() == exp
            assert np.array_equal(d["tiekey"][d["key_to_node"]], np.arange(n))


def test_tree_order_errors_do_not_depend_on_the_thread_split(monkeypatch):
    """The topology checks run on chunks of the node range: a broken parent entry anywhere in the array is reported with
    the same status and the same message (the FIRST offending node) as by a single pass."""
    import small_synth
    parent, row_ptr, muts, _ = small_synth.random_mat(31, 3000, 80, 2.0)
    n = len(parent)
    rng = np.random.default_rng(5)
    cases = []
    for _ in range(12):
        i = int(rng.integers(2, n))
        kind = int(rng.integers(0, 3))
        p = np.array(parent, dtype=np.int32).copy()
        if kind == 0:
            p[i] = i + int(rng.integers(0, 5))          # not an earlier node
        elif kind == 1:
            p[i] = -1                                    # a second root
        else:
            p[i] = int(rng.integers(0, i))               # an earlier node, most likely not on the path of i-1
        cases.append(p)
    two = np.array(parent, dtype=np.int32).copy()        # two errors: the first one in node order wins
    two[n - 5] = n + 3
    two[40] = 39 if parent[40] != 39 else 38
    cases.append(two)
    for p in cases:
        seen = []
        for nt in ("1", "5", "16"):
            monkeypatch.setenv("UB200_HOST_THREADS", nt)
            try:
                capi.debug_derive(p, row_ptr, muts, target_tiles=16, min_tile_cost=64)
                seen.append(("ok", ""))
            except capi.UB200Error as e:
                seen.append((e.code, str(e)))
        assert seen[0] == seen[1] == seen[2], seen
