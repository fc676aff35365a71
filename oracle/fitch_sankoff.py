"""TEST INFRASTRUCTURE ONLY — plain restatement of the reference's per-site Fitch-Sankoff assignment
(mapper_body::operator(), /root/reference/src/usher_mapper.cpp:6-161) on a tree given in BFS order: int score vectors,
BIG = number of nodes for impossible leaf states (:33-63), bottom-up min_k(score[c][k] + [k != j]) (:86-111), top-down
"parent's state unless another base is strictly cheaper, the first such base" (:114-156).  Pinned against the reference
itself (oracle/_ref builds a MAT from a newick + VCF with its own mapper_body) in tests/test_oracle.py."""
import numpy as np


def assign(parent_bfs, ref_code, var_ptr, var_node, var_nuc):
    """Returns (site, node, par_state, state) arrays sorted by (site, node): the mutations mapper_body would add."""
    n = len(parent_bfs)
    nch = np.zeros(n, np.int64)
    for i in range(1, n):
        nch[parent_bfs[i]] += 1
    out = []
    for s in range(len(ref_code)):
        ref = int(ref_code[s])
        score = np.zeros((n, 4), np.int64)
        leaves = nch == 0
        score[leaves] = n
        score[leaves, ref] = 0
        for k in range(int(var_ptr[s]), int(var_ptr[s + 1])):
            nuc = int(var_nuc[k])
            score[var_node[k]] = [0 if nuc & (1 << j) else n for j in range(4)]
        for i in range(n - 1, 0, -1):            # reverse BFS: children before parents
            p = parent_bfs[i]
            for j in range(4):
                score[p, j] += min(score[i, k] + (k != j) for k in range(4))
        state = np.zeros(n, np.int64)
        for i in range(n):
            par = ref if i == 0 else int(state[parent_bfs[i]])
            st, best = par, score[i, par]
            for j in range(4):
                if score[i, j] < best:
                    best, st = score[i, j], j
            state[i] = st
            if st != par:
                out.append((s, i, par, st))
    a = np.array(out, np.int64).reshape(-1, 4)
    return a[:, 0], a[:, 1], a[:, 2], a[:, 3]
