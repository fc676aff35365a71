"""Small randomized MATs and sample batches for parity tests, exercising every quirk of mapper2_body
(reference src/usher_mapper.cpp:167-504): masked mutations, reversions to the reference allele, IUPAC
ambiguity (with and without the reference allele), N calls on and off mutated positions, position collisions
between many branches (tiny genomes), empty branches, empty samples, single-node trees."""
import numpy as np

MUT_DTYPE = np.dtype(
    [("position", "<i4"), ("ref_nuc", "u1"), ("par_nuc", "u1"), ("mut_nuc", "u1"), ("is_missing", "u1")]
)


def random_mat(seed, n, L, mu, p_masked=0.04, shape="uniform"):
    rng = np.random.default_rng(seed)
    ref = np.zeros(L + 1, np.uint8)
    ref[1:] = 1 << rng.integers(0, 4, L)
    cpar = np.zeros(n, np.int64)
    for i in range(1, n):
        if shape == "chain":
            cpar[i] = i - 1 if rng.random() < 0.9 else rng.integers(0, i)
        elif shape == "star":
            cpar[i] = 0 if rng.random() < 0.7 else rng.integers(0, i)
        else:
            cpar[i] = rng.integers(0, i)
    kids = [[] for _ in range(n)]
    for i in range(1, n):
        kids[cpar[i]].append(i)
    order, st = [], [0]
    while st:
        u = st.pop()
        order.append(u)
        st.extend(reversed(kids[u]))
    newid = np.zeros(n, np.int64)
    newid[order] = np.arange(n)
    parent = np.array([-1 if d == 0 else newid[cpar[order[d]]] for d in range(n)], np.int32)
    # mutations by DFS with live state
    state = ref.copy()
    rows = [[] for _ in range(n)]
    undo = []  # (node, pos, old)
    path = []
    for d in range(n):
        while path and path[-1] != parent[d]:
            top = path.pop()
            while undo and undo[-1][0] == top:
                _, p, old = undo.pop()
                state[p] = old
        k = rng.poisson(mu) if (d != 0 or rng.random() < 0.3) else 0
        ps = sorted(set(int(x) for x in rng.integers(1, L + 1, k)))
        row = []
        if rng.random() < p_masked:
            for _ in range(int(rng.integers(1, 3))):
                row.append((-1, 0, 0, 0))
        for p in ps:
            cur = int(state[p])
            choices = [b for b in (1, 2, 4, 8) if b != cur]
            # bias toward reversions to ref so LOOP 1's "back to ref" branch is exercised
            if ref[p] != cur and rng.random() < 0.4:
                mut = int(ref[p])
            else:
                mut = int(rng.choice(choices))
            row.append((p, int(ref[p]), cur, mut))
            undo.append((d, p, cur))
            state[p] = mut
        rows[d] = row
        path.append(d)
    row_ptr = np.zeros(n + 1, np.uint64)
    muts = []
    for d in range(n):
        muts.extend(rows[d])
        row_ptr[d + 1] = len(muts)
    m = np.zeros(len(muts), MUT_DTYPE)
    for i, (p, r, pa, mu_) in enumerate(muts):
        m[i] = (p, r, pa, mu_, 0)
    return parent, row_ptr, m, ref


def genotype(parent, row_ptr, muts, v):
    g = {}
    n = v
    while n >= 0:
        for k in range(int(row_ptr[n]), int(row_ptr[n + 1])):
            p = int(muts[k]["position"])
            if p >= 0 and p not in g:
                g[p] = int(muts[k]["mut_nuc"])
        n = int(parent[n])
    return g


def random_samples(seed, parent, row_ptr, muts, ref, B, p_amb=0.15, p_n=0.2, extra=4):
    rng = np.random.default_rng(seed)
    n, L = len(parent), len(ref) - 1
    s_ptr = [0]
    calls = []
    for s in range(B):
        if rng.random() < 0.05:
            s_ptr.append(len(calls))  # empty sample
            continue
        v = int(rng.integers(0, n))
        g = genotype(parent, row_ptr, muts, v)
        cur = {}
        for p, nuc in g.items():
            if nuc != ref[p] and rng.random() < 0.85:
                cur[p] = (nuc, 0)
            elif nuc == ref[p] and rng.random() < 0.1:
                cur[p] = (int(rng.choice([1, 2, 4, 8])), 0)  # explicit call at a reverted position
        for _ in range(int(rng.integers(0, extra + 1))):
            p = int(rng.integers(1, L + 1))
            cur.setdefault(p, (int(rng.choice([b for b in (1, 2, 4, 8) if b != ref[p]])), 0))
        for p in list(cur):
            if rng.random() < p_amb:
                nuc = cur[p][0]
                for _ in range(int(rng.integers(1, 3))):
                    nuc |= 1 << int(rng.integers(0, 4))
                cur[p] = (nuc, 0)
        if rng.random() < p_n:
            # N run, often across mutated positions
            start = int(rng.integers(1, L + 1))
            for p in range(start, min(L, start + int(rng.integers(1, max(2, L // 6)))) + 1):
                cur[p] = (15, 1)
        if rng.random() < p_n and len(g):
            p = int(rng.choice(list(g)))
            cur[p] = (15, 1)
        for p in sorted(cur):
            calls.append((p, int(ref[p]), int(ref[p]), cur[p][0], cur[p][1]))
        s_ptr.append(len(calls))
    c = np.zeros(len(calls), MUT_DTYPE)
    for i, t in enumerate(calls):
        c[i] = t
    return np.array(s_ptr, np.uint64), c
