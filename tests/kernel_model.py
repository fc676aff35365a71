"""Host model of the scoring kernel's arithmetic (usher_b200/csrc/score_kernel.cuh), in plain Python, run on
the arrays produced by the C++ derivation (usher_b200/csrc/derive.cpp through ub200_debug_derive).  Test
infrastructure: it lets the CPU suite check the derivation and the closed-form decomposition against the
oracle without a GPU.  It is NOT a fallback: nothing in usher_b200/ imports it."""
import numpy as np

F_LEAF, F_MASKED, F_ROOT, F_VALID0 = 1, 2, 4, 8


def place(d, s_ptr, calls, per_node=False):
    n, L = d["n"], d["L"]
    hdr, mutw, row32 = d["hdr"], d["mutw"], d["row32"]
    B = len(s_ptr) - 1
    res = []
    node_scores = np.zeros((B, n), np.int32) if per_node else None
    for s in range(B):
        tab = {}
        base = 0
        for k in range(int(s_ptr[s]), int(s_ptr[s + 1])):
            c = calls[k]
            st = int(c["mut_nuc"]) & 15
            if not c["is_missing"] and (st & int(c["ref_nuc"])) == 0:
                base += 1
            if int(c["position"]) < L:
                tab[int(c["position"])] = 0x10 | (0 if c["is_missing"] else (~st & 15))
        stack = {}
        best = None  # (sc, tiekey, hu)
        cnt = 0
        optimal = []
        for i in range(n):
            h = hdr[i]
            level, flags = int(h["level_flags"]) >> 14, int(h["level_flags"]) & 255
            nmut, c0 = int(h["nmut_c0"]) >> 16, int(h["nmut_c0"]) & 0xFFFF
            root = bool(flags & F_ROOT)
            cpar = 0 if root else stack[level - 1]
            dcorr = da = dcom = 0
            for k in range(int(row32[i]), int(row32[i + 1])):
                m = int(mutw[k])
                e = tab.get(m >> 6, 0)
                if e & 0x10:
                    refc, prevc, mutc = (m >> 4) & 3, (m >> 2) & 3, m & 3
                    rm, rp = int(mutc != refc), int(prevc != refc)
                    wm, wp = (e >> mutc) & 1, (e >> prevc) & 1
                    dcorr += (wm - wp) - (rm - rp)
                    tk, t0 = wm ^ 1, rm ^ 1
                    da += (tk & wp) - (t0 & rp)
                    dcom += tk - t0
            assert nmut == int(row32[i + 1]) - int(row32[i])
            stack[level] = cpar + dcorr
            masked = bool(flags & F_MASKED)
            if masked:
                da = dcom = 0
            sc = int(h["g"]) + dcorr if root else int(h["g"]) + cpar - da
            common = c0 + dcom
            hu = (not root) and (masked or nmut > common)
            valid = root or ((common > 0) if (flags & F_LEAF) else ((not hu) or common > 0))
            if per_node:
                node_scores[s, i] = sc + base + (0 if valid else 1)
            if valid:
                key = (sc, int(h["tiekey"]), int(hu))
                if best is None or sc < best[0]:
                    best, cnt, optimal = key, 1, [(i, int(hu))]
                elif sc == best[0]:
                    cnt += 1
                    optimal.append((i, int(hu)))
                    if key < best:
                        best = key
        node = int(d["key_to_node"][best[1]])
        res.append({"score": best[0] + base, "best_node": node, "best_j": int(d["tie_index"][node]),
                    "num_best": cnt, "has_unique": best[2], "optimal": optimal})
    return res, node_scores


def place3(d, s_ptr, calls):
    """Model of k_score3 (usher_b200/csrc/score_kernel3.cuh) on the segment layout: per tile seed segments ->
    stack rows, then block segments; the correction above a node inside its block is
    stack[level above the block] + the dnode sums of its in-block ancestors; stack rows of later blocks come
    from the open chain only.  No pruning (pruning is exact in the kernel, so results must agree)."""
    n, L = d["n"], d["L"]
    hdr, stream = d["hdr3"], d["stream"]
    ts, w0, lvl, sseg, send = d["tile3_start"], d["tile3_w0"], d["tile3_lvl"], d["tile3_sseg"], d["seed_end"]
    F_HU0, F_OPEN = 16, 32
    B = len(s_ptr) - 1
    res = []
    for s in range(B):
        tab, base = {}, 0
        for k in range(int(s_ptr[s]), int(s_ptr[s + 1])):
            c = calls[k]
            st = int(c["mut_nuc"]) & 15
            if not c["is_missing"] and (st & int(c["ref_nuc"])) == 0:
                base += 1
            if int(c["position"]) < L:
                refc = int(c["ref_nuc"]).bit_length() - 1
                tab[int(c["position"])] = (0 if c["is_missing"] else (~st & 15), refc)

        def seg_deltas(o0, o1):
            """hits of stream[o0:o1] -> per lane-field [dcorr, da, dcom]"""
            out = {}
            for k in range(o0, o1):
                w = int(stream[k])
                e = tab.get(((w >> (16 if d["narrow3"] else 14)) << 5) | (w & 31))
                if e is None:
                    continue
                e, refc = e
                nl, prevc, mutc = (w >> 9) & 31, (w >> 7) & 3, (w >> 5) & 3
                rm, rp = int(mutc != refc), int(prevc != refc)
                wm, wp = (e >> mutc) & 1, (e >> prevc) & 1
                tk, t0 = wm ^ 1, rm ^ 1
                a = out.setdefault(nl, [0, 0, 0])
                a[0] += (wm - wp) - (rm - rp)
                a[1] += (tk & wp) - (t0 & rp)
                a[2] += tk - t0
            return out

        best, cnt, optimal = None, 0, []
        for t in range(len(ts) - 1):
            n0, n1 = int(ts[t]), int(ts[t + 1])
            assert n0 % 32 == 0
            o0 = int(w0[t]) * 256
            stack = {}
            lvl0 = int(lvl[t])
            for g, l0 in enumerate(range(0, lvl0, 32)):
                o1 = int(send[int(sseg[t]) + g]) * 4
                dl = seg_deltas(o0, o1)
                o0 = o1
                v = stack[l0 - 1] if l0 else 0
                for j in range(min(32, lvl0 - l0)):
                    v += dl.get(j, [0, 0, 0])[0]
                    stack[l0 + j] = v
            assert int(sseg[t + 1]) - int(sseg[t]) == (lvl0 + 31) // 32
            for blk in range(n0, n1, 32):
                nodes = range(blk, min(blk + 32, n1))
                seglen = sum(int(hdr[i]["nmut_c0"]) >> 16 for i in nodes)
                o1 = o0 + ((seglen + 3) & ~3)
                assert int(d["blk_words"][blk >> 5]) == o1 - o0   # what the scanner warp reads instead of headers
                # block record (k_score4 consumer): min(G - nmut), open-chain mask, level of the first open node
                rec = d["blk_rec"][blk >> 5]
                assert int(np.int32(rec[0])) == min(int(hdr[i]["g"]) - (int(hdr[i]["nmut_c0"]) >> 16) for i in nodes)
                opens = [i for i in nodes if int(hdr[i]["level_flags"]) & F_OPEN]
                assert int(rec[1]) == sum(1 << (i & 31) for i in opens) and int(rec[3]) == o1 - o0
                assert not opens or int(rec[2]) == int(hdr[opens[0]]["level_flags"]) >> 14
                dl = seg_deltas(o0, o1)
                o0 = o1

                def above(i):
                    h = hdr[i]
                    am, level = int(h["tiekey"]), int(h["level_flags"]) >> 14
                    top = level - bin(am).count("1")
                    v = stack[top - 1] if top else 0
                    for a in range(32):
                        if (am >> a) & 1 and a in dl:
                            v += dl[a][0]
                    return v

                for i in nodes:
                    h = hdr[i]
                    flags = int(h["level_flags"]) & 0x3FFF
                    nmut, c0 = int(h["nmut_c0"]) >> 16, int(h["nmut_c0"]) & 0xFFFF
                    dcorr, da, dcom = dl.get(i & 31, [0, 0, 0])
                    root, masked = bool(flags & F_ROOT), bool(flags & F_MASKED)
                    if masked:
                        da = dcom = 0
                    sc = int(h["g"]) + dcorr if root else int(h["g"]) + above(i) - da
                    common = c0 + dcom
                    hu = (not root) and (masked or nmut > common)
                    valid = root or ((common > 0) if (flags & F_LEAF) else ((not hu) or common > 0))
                    if (i & 31) not in dl:   # the kernel's dense path uses the precomputed flags
                        assert valid == bool(flags & F_VALID0) and (root or hu == bool(flags & F_HU0))
                    if valid:
                        key = (sc, int(d["tiekey"][i]), int(hu))
                        if best is None or sc < best[0]:
                            best, cnt, optimal = key, 1, [(i, int(hu))]
                        elif sc == best[0]:
                            cnt += 1
                            optimal.append((i, int(hu)))
                            if key < best:
                                best = key
                # open chain -> stack rows
                chain = [i for i in nodes if int(hdr[i]["level_flags"]) & F_OPEN]
                if chain:
                    lv = int(hdr[chain[0]]["level_flags"]) >> 14
                    assert int(hdr[chain[0]]["tiekey"]) == 0
                    v = stack[lv - 1] if lv else 0
                    for i in chain:
                        assert int(hdr[i]["level_flags"]) >> 14 == lv
                        v += dl.get(i & 31, [0, 0, 0])[0]
                        stack[lv] = v
                        lv += 1
            assert o0 <= int(w0[t + 1]) * 256 and (int(w0[t + 1]) * 256 - o0) < 256
        node = int(d["key_to_node"][best[1]])
        res.append({"score": best[0] + base, "best_node": node, "best_j": int(d["tie_index"][node]),
                    "num_best": cnt, "has_unique": best[2], "optimal": sorted(optimal)})
    return res

