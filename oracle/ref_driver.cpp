// TEST INFRASTRUCTURE ONLY — never linked into or called from the product (usher_b200/).
//
// C-ABI driver around the UNMODIFIED reference sources (compiled by path from /root/reference/src by
// oracle/Makefile into oracle/_ref/libusher_ref.so).  It lets tests/ and bench.py's cpu_baseline /
// `--impl reference` arm run the reference's own mapper2_body (src/usher_mapper.cpp:167-504) over a tree
// given as flat arrays, with exactly the search loop usher_common() wraps around it
// (src/usher_common.cpp:342-449).  Only the loop that fills mapper2_input is restated here (its locals
// best_node/best_j are not observable from outside usher_common); every line of arithmetic is the
// reference's.
//
// Also exposes the config-1 flow (newick + VCF -> MAT -> condense -> pb-style round trip -> place a second
// VCF) used to mint tests/golden/ (see oracle/make_golden.py).
#include "usher_common.hpp"   // reference header, found via -I/root/reference/src

#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace tbb { int oracle_threads = 1; }

// Definitions the reference expects from the translation units we do not compile (usher.cpp owns none of
// these; the pb load/save bodies are cut out of mutation_annotated_tree.cpp by the Makefile).
Mutation_Annotated_Tree::Tree Mutation_Annotated_Tree::load_mutation_annotated_tree(std::string) {
    fprintf(stderr, "oracle: protobuf load is not part of the oracle build\n");
    exit(1);
}
void Mutation_Annotated_Tree::save_mutation_annotated_tree(Mutation_Annotated_Tree::Tree, std::string) {
    fprintf(stderr, "oracle: protobuf save is not part of the oracle build\n");
    exit(1);
}

namespace {

struct RefTree {
    MAT::Tree T;
    std::vector<Missing_Sample> samples;
    std::vector<MAT::Node*> bfs_cache, dfs_cache;   // usher_ref_score_nodes: DFS expansion of a frozen from_flat tree
};

// What save_mutation_annotated_tree + load_mutation_annotated_tree do to a tree
// (src/mutation_annotated_tree.cpp:553-596, 614-659): the newick is written WITHOUT internal names, the
// tree is re-parsed (internal nodes renamed node_1.. in newick order), and the k-th node of a pre-order DFS
// gets the k-th mutation list back (entries with mut_nuc == par_nuc dropped unless masked).
MAT::Tree pb_round_trip(MAT::Tree& src) {
    std::string nwk = MAT::get_newick_string(src, false, true, true);
    auto dfs_src = src.depth_first_expansion();
    MAT::Tree dst = MAT::create_tree_from_newick_string(nwk);
    auto dfs_dst = dst.depth_first_expansion();
    if (dfs_src.size() != dfs_dst.size()) {
        fprintf(stderr, "oracle: round trip changed the node count\n");
        exit(1);
    }
    for (size_t i = 0; i < dfs_src.size(); i++) {
        for (auto& m : dfs_src[i]->mutations) {
            if (m.is_masked() || m.mut_nuc != m.par_nuc) dfs_dst[i]->add_mutation(m.copy());
        }
        if (!std::is_sorted(dfs_dst[i]->mutations.begin(), dfs_dst[i]->mutations.end()))
            std::sort(dfs_dst[i]->mutations.begin(), dfs_dst[i]->mutations.end());
    }
    // condensed nodes survive the pb (parsimony.proto: condensed_nodes)
    for (auto& kv : src.condensed_nodes) {
        dst.condensed_nodes[kv.first] = kv.second;
        for (auto& l : kv.second) dst.condensed_leaves.insert(l);
    }
    return dst;
}

}  // namespace

extern "C" {

struct ref_mut {          // one tree mutation or one sample call
    int32_t position;     // < 0 = masked (tree side only)
    uint8_t ref_nuc;      // one-hot A=1 C=2 G=4 T=8
    uint8_t par_nuc;
    uint8_t mut_nuc;      // tree: one-hot; sample: 4-bit IUPAC set
    uint8_t is_missing;   // sample side: N
};

void* usher_ref_tree_from_flat(uint32_t n_nodes, const int32_t* parent, const uint64_t* row_ptr,
                               const ref_mut* muts) {
    auto* h = new RefTree();
    std::vector<MAT::Node*> nodes(n_nodes);
    for (uint32_t i = 0; i < n_nodes; i++) {
        std::string id = "n" + std::to_string(i);
        if (parent[i] < 0) nodes[i] = h->T.create_node(id, -1.0f, 0);
        else nodes[i] = h->T.create_node(id, nodes[parent[i]], -1.0f);
        auto& v = nodes[i]->mutations;
        v.reserve((size_t)(row_ptr[i + 1] - row_ptr[i]));
        for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            MAT::Mutation m;
            m.position = muts[k].position;
            m.ref_nuc = (int8_t)muts[k].ref_nuc;
            m.par_nuc = (int8_t)muts[k].par_nuc;
            m.mut_nuc = (int8_t)muts[k].mut_nuc;
            m.is_missing = false;
            v.push_back(m);
        }
    }
    return h;
}

void usher_ref_tree_free(void* h) { delete (RefTree*)h; }

// config-1 style build: newick + VCF of the tree's own samples -> MAT (mapper_body Fitch-Sankoff through
// read_vcf(create_new_mat=true)), optional condense_leaves + pb round trip, as `usher -t -v -o` then
// `usher -i` would see it.
void* usher_ref_tree_from_newick_vcf(const char* newick_path, const char* vcf_path, int condense_and_round_trip,
                                     int threads) {
    tbb::oracle_threads = threads < 1 ? 1 : threads;
    auto* h = new RefTree();
    MAT::Tree T0 = MAT::create_tree_from_newick(newick_path);
    std::vector<Missing_Sample> ms0;
    std::string vcf(vcf_path);
    MAT::read_vcf(&T0, vcf, ms0, true);
    if (condense_and_round_trip) {
        T0.condense_leaves();
        h->T = pb_round_trip(T0);
    } else {
        h->T = T0;
    }
    return h;
}

uint64_t usher_ref_tree_parsimony(void* hv) { return ((RefTree*)hv)->T.get_parsimony_score(); }

// Export in DFS pre-order.  Call with NULL outputs to get sizes.
void usher_ref_tree_export(void* hv, uint32_t* n_nodes, uint64_t* n_muts, uint64_t* names_len, int32_t* parent,
                           uint64_t* row_ptr, ref_mut* muts, char* names /* '\n'-joined */) {
    auto* h = (RefTree*)hv;
    auto dfs = h->T.depth_first_expansion();
    std::unordered_map<MAT::Node*, int32_t> idx;
    for (size_t i = 0; i < dfs.size(); i++) idx[dfs[i]] = (int32_t)i;
    uint64_t nm = 0, nl = 0;
    for (auto* n : dfs) { nm += n->mutations.size(); nl += n->identifier.size() + 1; }
    *n_nodes = (uint32_t)dfs.size();
    *n_muts = nm;
    *names_len = nl;
    if (!parent) return;
    uint64_t k = 0, c = 0;
    for (size_t i = 0; i < dfs.size(); i++) {
        parent[i] = dfs[i]->parent ? idx[dfs[i]->parent] : -1;
        row_ptr[i] = k;
        for (auto& m : dfs[i]->mutations) {
            muts[k].position = m.position;
            muts[k].ref_nuc = (uint8_t)m.ref_nuc;
            muts[k].par_nuc = (uint8_t)m.par_nuc;
            muts[k].mut_nuc = (uint8_t)m.mut_nuc;
            muts[k].is_missing = 0;
            k++;
        }
        memcpy(names + c, dfs[i]->identifier.data(), dfs[i]->identifier.size());
        c += dfs[i]->identifier.size();
        names[c++] = '\n';
    }
    row_ptr[dfs.size()] = k;
}

// Placement-mode VCF read (src/mutation_annotated_tree.cpp:2180-2278) -> flat sample lists.
uint32_t usher_ref_read_samples(void* hv, const char* vcf_path) {
    auto* h = (RefTree*)hv;
    h->samples.clear();
    std::string vcf(vcf_path);
    MAT::read_vcf(&h->T, vcf, h->samples, false);
    return (uint32_t)h->samples.size();
}
void usher_ref_samples_export(void* hv, uint64_t* n_entries, uint64_t* names_len, uint64_t* s_ptr, ref_mut* out,
                              char* names) {
    auto* h = (RefTree*)hv;
    uint64_t ne = 0, nl = 0;
    for (auto& s : h->samples) { ne += s.mutations.size(); nl += s.name.size() + 1; }
    *n_entries = ne;
    *names_len = nl;
    if (!s_ptr) return;
    uint64_t k = 0, c = 0;
    for (size_t i = 0; i < h->samples.size(); i++) {
        s_ptr[i] = k;
        for (auto& m : h->samples[i].mutations) {
            out[k].position = m.position;
            out[k].ref_nuc = (uint8_t)m.ref_nuc;
            out[k].par_nuc = (uint8_t)m.par_nuc;
            out[k].mut_nuc = (uint8_t)m.mut_nuc;
            out[k].is_missing = m.is_missing ? 1 : 0;
            k++;
        }
        memcpy(names + c, h->samples[i].name.data(), h->samples[i].name.size());
        c += h->samples[i].name.size();
        names[c++] = '\n';
    }
    s_ptr[h->samples.size()] = k;
}

// The whole usher_common() flow (sequential graft) writing the reference's output files into outdir.
int usher_ref_usher_common(void* hv, const char* outdir, int threads, int print_parsimony_scores, int no_add) {
    auto* h = (RefTree*)hv;
    std::vector<std::string> low_conf;
    return usher_common("", outdir, (uint32_t)(threads < 1 ? 1 : threads), 1000000u, 1000000u,
                        /*sort_before_placement_1*/ false, /*_2*/ false, /*_3*/ false,
                        /*reverse_sort*/ false, /*collapse_tree*/ false, /*collapse_output_tree*/ false,
                        /*print_uncondensed_tree*/ false, /*print_parsimony_scores*/ print_parsimony_scores != 0,
                        /*retain_original_branch_len*/ false, /*no_add*/ no_add != 0, /*detailed_clades*/ false,
                        /*print_subtrees_size*/ 0, /*print_subtrees_single*/ 0, h->samples, low_conf, &h->T);
}

// Same with the sort pre-pass (-s / -S / -A, -r) and the thresholds (-e max_uncertainty, -E max_parsimony) exposed
// (src/usher.cpp:44-107 -> src/usher_common.cpp:7-21).
int usher_ref_usher_common2(void* hv, const char* outdir, int threads, int no_add, int sort1, int sort2, int sort3,
                            int reverse_sort, uint32_t max_uncertainty, uint32_t max_parsimony, int uncondensed) {
    auto* h = (RefTree*)hv;
    std::vector<std::string> low_conf;
    return usher_common("", outdir, (uint32_t)(threads < 1 ? 1 : threads), max_uncertainty, max_parsimony,
                        sort1 != 0, sort2 != 0, sort3 != 0, reverse_sort != 0, /*collapse_tree*/ false,
                        /*collapse_output_tree*/ false, /*print_uncondensed_tree*/ uncondensed != 0,
                        /*print_parsimony_scores*/ false, /*retain_original_branch_len*/ false, no_add != 0,
                        /*detailed_clades*/ false, 0, 0, h->samples, low_conf, &h->T);
}

// Frozen-tree search of every sample against every node, with the reference's loop
// (src/usher_common.cpp:342-449).  mode 0: two-pass search (pass 1 with early exit, pass 2 over the optimal
// set) exactly as the default CLI; mode 1: single pass with compute_parsimony_scores=true (-p semantics),
// filling node_scores[s*N + dfs_idx] with the reported score (score+1 on invalid nodes,
// src/usher_mapper.cpp:498-503).
// best_set: if non-NULL, receives for sample s the DFS indices of all optimal nodes at
// best_set[best_set_ptr[s] .. best_set_ptr[s+1]) (sorted ascending), best_set_unique the matching
// node_has_unique flags; capacity best_set_cap entries in total.
int usher_ref_search(void* hv, uint32_t n_samples, const uint64_t* s_ptr, const ref_mut* sm, int threads, int mode,
                     int32_t* score, uint32_t* best_dfs, uint32_t* best_j_out, uint32_t* num_best_out,
                     uint8_t* has_unique_out, int32_t* node_scores, uint32_t* best_set, uint8_t* best_set_unique,
                     uint64_t* best_set_ptr, uint64_t best_set_cap, double* seconds) {
    auto* h = (RefTree*)hv;
    MAT::Tree* T = &h->T;
    tbb::oracle_threads = threads < 1 ? 1 : threads;
    auto dfs = T->depth_first_expansion();
    std::unordered_map<MAT::Node*, uint32_t> dfs_idx;
    for (size_t i = 0; i < dfs.size(); i++) dfs_idx[dfs[i]] = (uint32_t)i;
    uint64_t set_fill = 0;
    if (best_set_ptr) best_set_ptr[0] = 0;
    for (uint32_t s = 0; s < n_samples; s++) {
        std::vector<MAT::Mutation> sample;
        for (uint64_t k = s_ptr[s]; k < s_ptr[s + 1]; k++) {
            MAT::Mutation m;
            m.chrom = "c";
            m.position = sm[k].position;
            m.ref_nuc = (int8_t)sm[k].ref_nuc;
            m.par_nuc = (int8_t)sm[k].par_nuc;
            m.mut_nuc = (int8_t)sm[k].mut_nuc;
            m.is_missing = sm[k].is_missing != 0;
            sample.push_back(m);
        }
        auto t0 = std::chrono::steady_clock::now();
        // ---- usher_common.cpp:342-379
        auto bfs = T->breadth_first_expansion();
        size_t total_nodes = bfs.size();
        std::vector<std::vector<MAT::Mutation>> node_excess_mutations(total_nodes);
        std::vector<std::vector<MAT::Mutation>> node_imputed_mutations(total_nodes);
        std::vector<int> node_set_difference;
        const bool pps = (mode == 1);
        if (pps) node_set_difference.resize(total_nodes);
        size_t best_node_num_leaves = 0;
        int best_set_difference = (int)(sample.size() + T->root->mutations.size() + 1);
        size_t best_j = 0;
        bool best_node_has_unique = false;
        std::vector<bool> node_has_unique(total_nodes, false);
        std::vector<size_t> best_j_vec;
        best_j_vec.emplace_back(0);
        size_t num_best = 1;
        MAT::Node* best_node = T->root;
        // ---- usher_common.cpp:388-414 (pass 1)
        tbb::parallel_for(tbb::blocked_range<size_t>(0, total_nodes), [&](tbb::blocked_range<size_t> r) {
            for (size_t k = r.begin(); k < r.end(); ++k) {
                mapper2_input inp;
                inp.T = T;
                inp.node = bfs[k];
                inp.missing_sample_mutations = &sample;
                inp.excess_mutations = &node_excess_mutations[k];
                inp.imputed_mutations = &node_imputed_mutations[k];
                inp.best_node_num_leaves = &best_node_num_leaves;
                inp.best_set_difference = &best_set_difference;
                inp.best_node = &best_node;
                inp.best_j = &best_j;
                inp.num_best = &num_best;
                inp.j = k;
                inp.has_unique = &best_node_has_unique;
                if (pps) inp.set_difference = &node_set_difference[k];
                inp.best_j_vec = &best_j_vec;
                inp.node_has_unique = &node_has_unique;
                mapper2_body(inp, pps, pps);
            }
        });
        // ---- usher_common.cpp:416-449 (pass 2)
        if (!pps) {
            best_set_difference += 1;
            auto tmp_vec = std::vector<size_t>(best_j_vec.begin(), best_j_vec.end());
            num_best = 0;
            best_j_vec.clear();
            tbb::parallel_for(tbb::blocked_range<size_t>(0, tmp_vec.size()), [&](tbb::blocked_range<size_t> r) {
                for (size_t l = r.begin(); l < r.end(); ++l) {
                    auto k = tmp_vec[l];
                    mapper2_input inp;
                    inp.T = T;
                    inp.node = bfs[k];
                    inp.missing_sample_mutations = &sample;
                    inp.excess_mutations = &node_excess_mutations[k];
                    inp.imputed_mutations = &node_imputed_mutations[k];
                    inp.best_node_num_leaves = &best_node_num_leaves;
                    inp.best_set_difference = &best_set_difference;
                    inp.best_node = &best_node;
                    inp.best_j = &best_j;
                    inp.num_best = &num_best;
                    inp.j = k;
                    inp.has_unique = &best_node_has_unique;
                    inp.best_j_vec = &best_j_vec;
                    inp.node_has_unique = &node_has_unique;
                    mapper2_body(inp, false);
                }
            });
        }
        auto t1 = std::chrono::steady_clock::now();
        if (seconds) seconds[s] = std::chrono::duration<double>(t1 - t0).count();
        score[s] = best_set_difference;
        best_dfs[s] = dfs_idx[best_node];
        best_j_out[s] = (uint32_t)best_j;
        num_best_out[s] = (uint32_t)num_best;
        has_unique_out[s] = best_node_has_unique ? 1 : 0;
        if (pps && node_scores) {
            for (size_t k = 0; k < total_nodes; k++)
                node_scores[(uint64_t)s * total_nodes + dfs_idx[bfs[k]]] = node_set_difference[k];
        }
        if (best_set && best_set_ptr) {
            std::vector<std::pair<uint32_t, uint8_t>> v;
            for (auto j : best_j_vec) v.emplace_back(dfs_idx[bfs[j]], node_has_unique[j] ? 1 : 0);
            std::sort(v.begin(), v.end());
            for (auto& p : v) {
                if (set_fill < best_set_cap) {
                    best_set[set_fill] = p.first;
                    if (best_set_unique) best_set_unique[set_fill] = p.second;
                }
                set_fill++;
            }
            best_set_ptr[s + 1] = set_fill;
        }
    }
    return set_fill > best_set_cap && best_set ? 1 : 0;
}

// Timing aid for trees where one full search is minutes of CPU: run the reference's two-pass search of ONE
// sample over every `stride`-th BFS node (k = offset, offset+stride, ...) with `threads` workers and return
// the seconds spent.  seconds*stride only ESTIMATES the full search: the ancestor gather of a node costs the same
// whichever nodes are visited, but the early exits depend on the best found so far (see _strided2).
double usher_ref_search_strided2(void* hv, uint64_t n_calls, const ref_mut* sm, uint32_t stride, uint32_t offset,
                                 int threads, int32_t seed_best, int32_t* best_score_seen);
double usher_ref_search_strided(void* hv, uint64_t n_calls, const ref_mut* sm, uint32_t stride, uint32_t offset,
                                int threads, int32_t* best_score_seen) {
    return usher_ref_search_strided2(hv, n_calls, sm, stride, offset, threads, -1, best_score_seen);
}
// seed_best >= 0: the running best starts from min(reference's initial bound, seed_best).  A strided visit finds
// its best later than the full search does, so its early exits (src/usher_mapper.cpp:383,437) fire less often and
// seconds*stride OVER-states the full search; seeded with the true best score they fire at least as often as in
// the full search and seconds*stride UNDER-states it.  The two bracket the stride-1 time.
double usher_ref_search_strided2(void* hv, uint64_t n_calls, const ref_mut* sm, uint32_t stride, uint32_t offset,
                                 int threads, int32_t seed_best, int32_t* best_score_seen) {
    auto* h = (RefTree*)hv;
    MAT::Tree* T = &h->T;
    tbb::oracle_threads = threads < 1 ? 1 : threads;
    std::vector<MAT::Mutation> sample;
    for (uint64_t k = 0; k < n_calls; k++) {
        MAT::Mutation m;
        m.position = sm[k].position;
        m.ref_nuc = (int8_t)sm[k].ref_nuc;
        m.par_nuc = (int8_t)sm[k].par_nuc;
        m.mut_nuc = (int8_t)sm[k].mut_nuc;
        m.is_missing = sm[k].is_missing != 0;
        sample.push_back(m);
    }
    // the O(N) expansion is done once per (frozen) tree, outside the timed part
    if (h->bfs_cache.empty()) h->bfs_cache = T->breadth_first_expansion();
    const std::vector<MAT::Node*>& bfs = h->bfs_cache;
    if (stride < 1) stride = 1;
    const size_t total_nodes = bfs.size();
    const size_t visits = (total_nodes > offset) ? (total_nodes - offset + stride - 1) / stride : 0;
    auto t0 = std::chrono::steady_clock::now();
    // one extra slot: pass 2 may revisit BFS node 0 (the initial best_j_vec entry) which is not on the stride
    std::vector<std::vector<MAT::Mutation>> node_excess_mutations(visits + 1);
    std::vector<std::vector<MAT::Mutation>> node_imputed_mutations(visits + 1);
    size_t best_node_num_leaves = 0;
    int best_set_difference = (int)(sample.size() + T->root->mutations.size() + 1);
    if (seed_best >= 0 && seed_best < best_set_difference) best_set_difference = seed_best;
    size_t best_j = 0;
    bool best_node_has_unique = false;
    std::vector<bool> node_has_unique(total_nodes, false);
    std::vector<size_t> best_j_vec;
    best_j_vec.emplace_back(0);
    size_t num_best = 1;
    MAT::Node* best_node = T->root;
    auto body = [&](bool second, const std::vector<size_t>* only) {
        const size_t cnt = only ? only->size() : visits;
        tbb::parallel_for(tbb::blocked_range<size_t>(0, cnt), [&](tbb::blocked_range<size_t> r) {
            for (size_t q = r.begin(); q < r.end(); ++q) {
                const size_t k = only ? (*only)[q] : offset + q * stride;
                const size_t slot = !only ? q : ((k >= offset && (k - offset) % stride == 0) ? (k - offset) / stride : visits);
                mapper2_input inp;
                inp.T = T;
                inp.node = bfs[k];
                inp.missing_sample_mutations = &sample;
                inp.excess_mutations = &node_excess_mutations[slot];
                inp.imputed_mutations = &node_imputed_mutations[slot];
                inp.best_node_num_leaves = &best_node_num_leaves;
                inp.best_set_difference = &best_set_difference;
                inp.best_node = &best_node;
                inp.best_j = &best_j;
                inp.num_best = &num_best;
                inp.j = k;
                inp.has_unique = &best_node_has_unique;
                inp.best_j_vec = &best_j_vec;
                inp.node_has_unique = &node_has_unique;
                if (second) mapper2_body(inp, false);
                else mapper2_body(inp, false, false);
            }
        });
    };
    body(false, nullptr);
    best_set_difference += 1;
    auto tmp_vec = std::vector<size_t>(best_j_vec.begin(), best_j_vec.end());
    num_best = 0;
    best_j_vec.clear();
    body(true, &tmp_vec);
    auto t1 = std::chrono::steady_clock::now();
    if (best_score_seen) *best_score_seen = best_set_difference;
    return std::chrono::duration<double>(t1 - t0).count();
}

// -p semantics at a LIST of nodes: mapper2_body(inp, true, false) of ONE sample at the nodes whose DFS indices
// are given (the tree must come from usher_ref_tree_from_flat, whose DFS order is the flat order).  scores_out[i]
// = the reported score of dfs_nodes[i] (score+1 on invalid nodes, src/usher_mapper.cpp:448-450,498-503).  Used by
// bench.py / tests to spot-check 10M-node results without a full O(N) reference search.
int usher_ref_score_nodes(void* hv, uint64_t n_calls, const ref_mut* sm, uint64_t n_list, const uint32_t* dfs_nodes,
                          int threads, int32_t* scores_out, uint8_t* valid_out) {
    auto* h = (RefTree*)hv;
    MAT::Tree* T = &h->T;
    tbb::oracle_threads = threads < 1 ? 1 : threads;
    if (h->dfs_cache.empty()) h->dfs_cache = T->depth_first_expansion();
    const std::vector<MAT::Node*>& dfs = h->dfs_cache;
    std::vector<MAT::Mutation> sample;
    for (uint64_t k = 0; k < n_calls; k++) {
        MAT::Mutation m;
        m.position = sm[k].position;
        m.ref_nuc = (int8_t)sm[k].ref_nuc;
        m.par_nuc = (int8_t)sm[k].par_nuc;
        m.mut_nuc = (int8_t)sm[k].mut_nuc;
        m.is_missing = sm[k].is_missing != 0;
        sample.push_back(m);
    }
    for (uint64_t i = 0; i < n_list; i++)
        if (dfs_nodes[i] >= dfs.size()) return -1;
    tbb::parallel_for(tbb::blocked_range<size_t>(0, (size_t)n_list), [&](tbb::blocked_range<size_t> r) {
        // a fresh best-state per node: the per-node score does not depend on it when compute_parsimony_scores
        // is true (no early exit, src/usher_mapper.cpp:383,437), and a node is a VALID placement exactly when the
        // call folds it into that fresh state (src/usher_mapper.cpp:454-468)
        for (size_t q = r.begin(); q < r.end(); ++q) {
            size_t best_node_num_leaves = 0, best_j = 0, num_best = 1;
            int best_set_difference = 1 << 30;
            bool best_node_has_unique = false;
            MAT::Node* best_node = T->root;
            std::vector<size_t> best_j_vec;
            std::vector<bool> node_has_unique(1, false);
            std::vector<MAT::Mutation> excess, imputed;
            int sd = 0;
            mapper2_input inp;
            inp.T = T;
            inp.node = dfs[dfs_nodes[q]];
            inp.missing_sample_mutations = &sample;
            inp.excess_mutations = &excess;
            inp.imputed_mutations = &imputed;
            inp.best_node_num_leaves = &best_node_num_leaves;
            inp.best_set_difference = &best_set_difference;
            inp.best_node = &best_node;
            inp.best_j = &best_j;
            inp.num_best = &num_best;
            inp.j = 0;
            inp.has_unique = &best_node_has_unique;
            inp.set_difference = &sd;
            inp.best_j_vec = &best_j_vec;
            inp.node_has_unique = &node_has_unique;
            mapper2_body(inp, true, false);
            scores_out[q] = sd;
            if (valid_out) valid_out[q] = best_set_difference != (1 << 30) ? (best_node_has_unique ? 3 : 1) : 0;
        }
    });
    return 0;
}

// condensed_nodes of the tree as text: one line per node, "name\tmember1,member2,...\n"
uint64_t usher_ref_condensed_export(void* hv, char* out, uint64_t cap) {
    auto* h = (RefTree*)hv;
    std::string s;
    for (auto* n : h->T.depth_first_expansion()) {
        auto it = h->T.condensed_nodes.find(n->identifier);
        if (it == h->T.condensed_nodes.end()) continue;
        s += n->identifier + "\t";
        for (size_t i = 0; i < it->second.size(); i++) { if (i) s += ","; s += it->second[i]; }
        s += "\n";
    }
    if (out && cap >= s.size()) memcpy(out, s.data(), s.size());
    return s.size();
}

}  // extern "C"
