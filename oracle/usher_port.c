/* TEST INFRASTRUCTURE ONLY — never linked into or called from the product (usher_b200/).
 *
 * Plain-C restatement of the reference's sample-placement hot path, loop for loop:
 *   port_mapper2()  <- mapper2_body            /root/reference/src/usher_mapper.cpp:167-504
 *   port_search()   <- the per-sample search   /root/reference/src/usher_common.cpp:342-449
 *   port_bfs()      <- Tree::breadth_first_expansion   src/mutation_annotated_tree.cpp:1225-1251
 *   port_num_leaves <- Tree::get_num_leaves            src/mutation_annotated_tree.cpp:866-879
 * It deliberately keeps the reference's O(P^2) structure (ancestor gather with a linear "already seen"
 * search, linear scans in LOOP 2 / LOOP 3) so that it shares nothing with the closed form the CUDA path
 * uses.  Parity pinned: tests/test_oracle.py checks it against oracle/_ref (the reference's own sources
 * compiled here) on config 1 and on randomized trees, and against tests/golden/ minted from oracle/_ref.
 *
 * Tree input = the flat form of include/usher_b200.h: nodes in DFS pre-order, parent[] by DFS index,
 * CSR mutation rows (position-sorted, masked = position < 0 first).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t position;
    uint8_t ref_nuc, par_nuc, mut_nuc, is_missing;
} port_mut;

typedef struct {
    uint32_t n;
    const int32_t* parent;
    const uint64_t* row_ptr;
    const port_mut* muts;
    uint32_t* bfs;        /* bfs[j] = dfs index of j-th node in BFS order */
    uint32_t* num_leaves; /* per dfs index */
    uint8_t* is_leaf;
} port_tree;

/* shared "best" state that mapper2_input points at (src/usher_graph.hpp:73-101) */
typedef struct {
    int best_set_difference;
    size_t best_node_num_leaves;
    size_t best_j;
    size_t num_best;
    uint32_t best_node;
    int has_unique;
    size_t* best_j_vec;
    size_t best_j_len;
    uint8_t* node_has_unique; /* indexed by j */
} port_best;

typedef struct {
    port_mut* v;
    size_t len, cap;
} mvec;

static void mvec_push(mvec* a, port_mut m) {
    if (a->len == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 64;
        a->v = (port_mut*)realloc(a->v, a->cap * sizeof(port_mut));
    }
    a->v[a->len++] = m;
}

static int cmp_pos(const void* a, const void* b) {
    int32_t x = ((const port_mut*)a)->position, y = ((const port_mut*)b)->position;
    return (x > y) - (x < y);
}

/* usher_mapper.cpp:167-504.  set_difference_out may be NULL.  Returns nothing; folds into *B. */
static void port_mapper2(const port_tree* T, uint32_t node, size_t j, const port_mut* S, size_t nS, port_best* B,
                         int compute_parsimony_scores, int* set_difference_out, mvec* anc) {
    int set_difference = 0;
    int best_set_difference = B->best_set_difference; /* :176 snapshot */
    int has_unique = 0;
    int node_num_mut = 0;
    int num_common_mut = 0;
    const port_mut* row = T->muts + T->row_ptr[node];
    size_t nrow = (size_t)(T->row_ptr[node + 1] - T->row_ptr[node]);
    int is_root = T->parent[node] < 0;
    anc->len = 0;

    if (!is_root) { /* LOOP 1, :190-264 */
        size_t start_index = 0;
        for (size_t a = 0; a < nrow; a++) {
            port_mut m1 = row[a];
            node_num_mut++;
            uint8_t anc_nuc = m1.mut_nuc;
            if (m1.position < 0) { /* masked :197-200 */
                has_unique = 1;
                break;
            }
            int found = 0, found_pos = 0;
            for (size_t k = start_index; k < nS; k++) {
                port_mut m2 = S[k];
                start_index = k;
                if (m1.position == m2.position) {
                    found_pos = 1;
                    if (m2.is_missing) {
                        found = 1;
                        num_common_mut++;
                    } else {
                        uint8_t nuc = m2.mut_nuc;
                        if ((nuc & anc_nuc) != 0) {
                            port_mut m = m1;
                            m.mut_nuc = anc_nuc;
                            m.is_missing = 0;
                            mvec_push(anc, m);
                            found = 1;
                            num_common_mut++;
                            break;
                        }
                    }
                }
                if (m1.position < m2.position) break;
            }
            if (!found) {
                if (!found_pos && (anc_nuc == m1.ref_nuc)) { /* :244-259 */
                    port_mut m = m1;
                    m.mut_nuc = anc_nuc;
                    m.is_missing = 0;
                    mvec_push(anc, m);
                    num_common_mut++;
                } else {
                    has_unique = 1;
                }
            }
        }
    } else { /* :265-270 */
        for (size_t a = 0; a < nrow; a++) mvec_push(anc, row[a]);
    }

    { /* ancestor gather :275-286 (anc_positions == positions of anc so far) */
        int32_t n = (int32_t)node;
        while (T->parent[n] >= 0) {
            n = T->parent[n];
            const port_mut* r = T->muts + T->row_ptr[n];
            size_t nr = (size_t)(T->row_ptr[n + 1] - T->row_ptr[n]);
            for (size_t a = 0; a < nr; a++) {
                if (r[a].position < 0) continue;
                int seen = 0;
                for (size_t q = 0; q < anc->len; q++)
                    if (anc->v[q].position == r[a].position) { seen = 1; break; }
                if (!seen) mvec_push(anc, r[a]);
            }
        }
    }
    qsort(anc->v, anc->len, sizeof(port_mut), cmp_pos); /* :289 */

    for (size_t a = 0; a < nS; a++) { /* LOOP 2, :292-388 */
        port_mut m1 = S[a];
        if (m1.is_missing) continue;
        int found_pos = 0, found = 0, has_ref = 0;
        uint8_t anc_nuc = m1.ref_nuc;
        if ((m1.mut_nuc & m1.ref_nuc) != 0) has_ref = 1;
        for (size_t k = 0; k < anc->len; k++) {
            port_mut m2 = anc->v[k];
            if (m2.position < 0) continue;
            if (m1.position == m2.position) {
                found_pos = 1;
                anc_nuc = m2.mut_nuc;
                if ((m1.mut_nuc & anc_nuc) != 0) found = 1;
                break;
            }
        }
        if (found) {
        } else if (!found_pos && has_ref) {
        } else {
            uint8_t par = anc_nuc, mut = 0;
            if (has_ref) mut = m1.ref_nuc;
            else
                for (int b = 0; b < 4; b++)
                    if (((1 << b) & m1.mut_nuc) != 0) { mut = (uint8_t)(1 << b); break; }
            if (mut != par) {
                set_difference += 1;
                if (!compute_parsimony_scores && (set_difference > best_set_difference)) return; /* :383 */
            }
        }
    }

    for (size_t a = 0; a < anc->len; a++) { /* LOOP 3, :393-445 */
        port_mut m1 = anc->v[a];
        int found = 0, found_pos = 0;
        uint8_t anc_nuc = m1.mut_nuc;
        int masked = m1.position < 0;
        for (size_t k = 0; k < nS; k++) {
            if (masked) break;
            port_mut m2 = S[k];
            if (m1.position == m2.position) {
                found_pos = 1;
                if (m2.is_missing) { found = 1; break; }
                if ((m2.mut_nuc & anc_nuc) != 0) found = 1;
            }
        }
        if (found) {
        } else if (!found_pos && !masked && (anc_nuc == m1.ref_nuc)) {
        } else if (found_pos && !found) {
        } else {
            uint8_t par = anc_nuc, mut = m1.ref_nuc;
            if (mut != par) {
                set_difference += 1;
                if (!compute_parsimony_scores && (set_difference > best_set_difference)) return; /* :437 */
            }
        }
    }

    if (compute_parsimony_scores && set_difference_out) *set_difference_out = set_difference; /* :448-450 */

    int leaf = T->is_leaf[node];
    if (is_root || ((has_unique && !leaf && (num_common_mut > 0) && (node_num_mut != num_common_mut)) ||
                    (leaf && (num_common_mut > 0)) || (!has_unique && !leaf && (node_num_mut == num_common_mut)))) {
        if (set_difference > B->best_set_difference) return; /* :456-461 */
        size_t num_leaves = T->num_leaves[node];                /* :464 */
        if (set_difference < B->best_set_difference) {
            B->best_set_difference = set_difference;
            B->best_node = node;
            B->best_node_num_leaves = num_leaves;
            B->best_j = j;
            B->num_best = 1;
            B->has_unique = has_unique;
            B->node_has_unique[j] = (uint8_t)has_unique;
            B->best_j_len = 0;
            B->best_j_vec[B->best_j_len++] = j;
        } else if (set_difference == B->best_set_difference) {
            /* distance == best_distance == 0 always in usher_common (usher_graph.hpp:97-100) */
            if ((num_leaves > B->best_node_num_leaves) ||
                ((num_leaves == B->best_node_num_leaves) && (B->best_j < j))) {
                B->best_set_difference = set_difference;
                B->best_node = node;
                B->best_node_num_leaves = num_leaves;
                B->best_j = j;
                B->has_unique = has_unique;
            }
            B->num_best += 1;
            B->node_has_unique[j] = (uint8_t)has_unique;
            B->best_j_vec[B->best_j_len++] = j;
        }
    } else if (compute_parsimony_scores && set_difference_out) {
        *set_difference_out = set_difference + 1; /* :498-503 */
    }
}

/* ---- tree helpers ---- */
void* port_tree_create(uint32_t n, const int32_t* parent, const uint64_t* row_ptr, const port_mut* muts) {
    port_tree* T = (port_tree*)calloc(1, sizeof(port_tree));
    T->n = n;
    T->parent = parent;
    T->row_ptr = row_ptr;
    T->muts = muts;
    T->bfs = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    T->num_leaves = (uint32_t*)calloc(n ? n : 1, sizeof(uint32_t));
    T->is_leaf = (uint8_t*)malloc(n ? n : 1);
    memset(T->is_leaf, 1, n);
    /* children in stored order == increasing DFS index among nodes sharing a parent */
    uint32_t* cnt = (uint32_t*)calloc(n + 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < n; i++)
        if (parent[i] >= 0) { cnt[parent[i] + 1]++; T->is_leaf[parent[i]] = 0; }
    for (uint32_t i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    uint32_t* kids = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (n + 1));
    memcpy(fill, cnt, sizeof(uint32_t) * (n + 1));
    for (uint32_t i = 0; i < n; i++)
        if (parent[i] >= 0) kids[fill[parent[i]]++] = i;
    /* BFS (mutation_annotated_tree.cpp:1225-1251): queue from the root, children in order */
    uint32_t head = 0, tail = 0;
    if (n) T->bfs[tail++] = 0;
    while (head < tail) {
        uint32_t u = T->bfs[head++];
        for (uint32_t k = cnt[u]; k < cnt[u + 1]; k++) T->bfs[tail++] = kids[k];
    }
    /* get_num_leaves (:866-879): number of childless nodes in the subtree (a leaf counts itself) */
    for (uint32_t i = n; i-- > 0;) {
        if (T->is_leaf[i]) T->num_leaves[i] = 1;
        if (parent[i] >= 0) T->num_leaves[parent[i]] += T->num_leaves[i];
    }
    free(cnt); free(kids); free(fill);
    return T;
}

void port_tree_free(void* t) {
    port_tree* T = (port_tree*)t;
    if (!T) return;
    free(T->bfs); free(T->num_leaves); free(T->is_leaf); free(T);
}

void port_tree_bfs(void* t, uint32_t* bfs_out, uint32_t* num_leaves_out) {
    port_tree* T = (port_tree*)t;
    if (bfs_out) memcpy(bfs_out, T->bfs, sizeof(uint32_t) * T->n);
    if (num_leaves_out) memcpy(num_leaves_out, T->num_leaves, sizeof(uint32_t) * T->n);
}

/* usher_common.cpp:342-449 for each sample on the frozen tree.
 * mode 0: two-pass search; mode 1: -p (compute_parsimony_scores) single pass + node_scores[s*n + dfs].
 * best_set (optional): all optimal nodes as DFS indices, ascending, with node_has_unique flags. */
int port_search(void* t, uint32_t n_samples, const uint64_t* s_ptr, const port_mut* sm, int mode, int32_t* score,
                uint32_t* best_dfs, uint32_t* best_j_out, uint32_t* num_best_out, uint8_t* has_unique_out,
                int32_t* node_scores, uint32_t* best_set, uint8_t* best_set_unique, uint64_t* best_set_ptr,
                uint64_t best_set_cap) {
    port_tree* T = (port_tree*)t;
    const uint32_t n = T->n;
    mvec anc = {0, 0, 0};
    port_best B;
    B.best_j_vec = (size_t*)malloc(sizeof(size_t) * (n + 1));
    B.node_has_unique = (uint8_t*)malloc(n + 1);
    size_t* tmp_vec = (size_t*)malloc(sizeof(size_t) * (n + 1));
    uint64_t set_fill = 0;
    if (best_set_ptr) best_set_ptr[0] = 0;
    const int pps = (mode == 1);
    for (uint32_t s = 0; s < n_samples; s++) {
        const port_mut* S = sm + s_ptr[s];
        size_t nS = (size_t)(s_ptr[s + 1] - s_ptr[s]);
        /* :364-385 */
        B.best_node_num_leaves = 0;
        B.best_set_difference = (int)(nS + (size_t)(T->row_ptr[1] - T->row_ptr[0]) + 1);
        B.best_j = 0;
        B.has_unique = 0;
        memset(B.node_has_unique, 0, n + 1);
        B.best_j_len = 0;
        B.best_j_vec[B.best_j_len++] = 0;
        B.num_best = 1;
        B.best_node = 0;
        for (size_t k = 0; k < n; k++) { /* pass 1 :388-414 */
            int sd = 0;
            port_mapper2(T, T->bfs[k], k, S, nS, &B, pps, &sd, &anc);
            if (pps && node_scores) node_scores[(uint64_t)s * n + T->bfs[k]] = sd;
        }
        if (!pps) { /* pass 2 :416-449 */
            B.best_set_difference += 1;
            size_t nt = B.best_j_len;
            memcpy(tmp_vec, B.best_j_vec, sizeof(size_t) * nt);
            B.num_best = 0;
            B.best_j_len = 0;
            for (size_t l = 0; l < nt; l++) {
                size_t k = tmp_vec[l];
                port_mapper2(T, T->bfs[k], k, S, nS, &B, 0, NULL, &anc);
            }
        }
        score[s] = B.best_set_difference;
        best_dfs[s] = B.best_node;
        best_j_out[s] = (uint32_t)B.best_j;
        num_best_out[s] = (uint32_t)B.num_best;
        has_unique_out[s] = (uint8_t)B.has_unique;
        if (best_set && best_set_ptr) {
            /* emit in ascending DFS order */
            uint8_t* mark = (uint8_t*)calloc(n, 1);
            for (size_t q = 0; q < B.best_j_len; q++)
                mark[T->bfs[B.best_j_vec[q]]] = (uint8_t)(1 + B.node_has_unique[B.best_j_vec[q]]);
            for (uint32_t i = 0; i < n; i++)
                if (mark[i]) {
                    if (set_fill < best_set_cap) {
                        best_set[set_fill] = i;
                        if (best_set_unique) best_set_unique[set_fill] = (uint8_t)(mark[i] - 1);
                    }
                    set_fill++;
                }
            free(mark);
            best_set_ptr[s + 1] = set_fill;
        }
    }
    free(anc.v); free(B.best_j_vec); free(B.node_has_unique); free(tmp_vec);
    return (best_set && set_fill > best_set_cap) ? 1 : 0;
}
