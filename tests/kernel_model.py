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
