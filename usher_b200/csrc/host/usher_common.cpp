// Placement driver of the drop-in `usher` binary: the flow of the reference's usher_common()
// (src/usher_common.cpp:7-1073) with the three tbb::parallel_for search loops (:252-273, :389-414, :426-449)
// replaced by calls into the C ABI of include/usher_b200.h.  Everything the kernel does not return — the
// excess / imputed mutation lists of the chosen node (mapper2_body's compute_vecs outputs), thresholds, clade
// lookup, grafting, output files — stays on the host, as in SURVEY.md §8(b).
#include "usher_common.hpp"

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <unordered_map>
#include <unordered_set>

#include "usher_b200.h"

namespace {

struct Flat {
    std::vector<int32_t> parent;
    std::vector<uint64_t> row_ptr;
    std::vector<ub200_mutation> muts;
    std::vector<MAT::Node*> dfs;
};

// Tree -> the flat DFS form of include/usher_b200.h
void flatten(const MAT::Tree& T, Flat& f) {
    f.dfs = T.depth_first_expansion();
    f.parent.resize(f.dfs.size());
    f.row_ptr.assign(1, 0);
    f.muts.clear();
    size_t total = 0;
    for (auto n : f.dfs) total += n->mutations.size();
    f.muts.reserve(total);
    f.row_ptr.reserve(f.dfs.size() + 1);
    // pre-order: the parent of a node is the nearest node still open on the walk's stack (no pointer -> index map)
    std::vector<std::pair<const MAT::Node*, int32_t>> open;
    for (size_t i = 0; i < f.dfs.size(); i++) {
        const MAT::Node* n = f.dfs[i];
        while (!open.empty() && open.back().first != n->parent) open.pop_back();
        f.parent[i] = open.empty() ? -1 : open.back().second;
        open.emplace_back(n, (int32_t)i);
        for (auto& m : n->mutations)
            f.muts.push_back({m.position, (uint8_t)m.ref_nuc, (uint8_t)m.par_nuc, (uint8_t)m.mut_nuc, 0});
        f.row_ptr.push_back(f.muts.size());
    }
}

struct DeviceTree {
    ub200_multi* multi = nullptr;   // one replica per GPU in use (frozen-tree batches use every visible GPU)
    ub200_mat* mat = nullptr;       // replica 0: the sequential search and the per-node dump
    Flat flat;
    ~DeviceTree() { if (multi) ub200_multi_destroy(multi); }
    // device >= 0: that GPU only; device < 0: every visible GPU when `all_devices`, else GPU 0
    void build(const MAT::Tree& T, int device, bool all_devices) {
        if (multi) { ub200_multi_destroy(multi); multi = nullptr; mat = nullptr; }
        flatten(T, flat);
        ub200_flat_mat v{(uint32_t)flat.parent.size(), flat.muts.size(), flat.parent.data(), flat.row_ptr.data(),
                         flat.muts.data(), nullptr};
        const int one = device < 0 ? 0 : device;
        const bool every = device < 0 && all_devices;
        if (ub200_multi_create(&v, every ? 0 : 1, every ? nullptr : &one, &multi) != UB200_OK) {
            fprintf(stderr, "ERROR: %s\n", ub200_last_error());
            exit(1);
        }
        mat = ub200_multi_mat(multi, 0);
        // wide passes share one scan of the tree between three sample groups
        for (int i = 0; i < ub200_multi_size(multi); i++) ub200_mat_set_pass_samples(ub200_multi_mat(multi, i), 96);
    }
};

void sample_calls(const Missing_Sample& s, std::vector<ub200_mutation>& out) {
    for (auto& m : s.mutations)
        out.push_back({m.position, (uint8_t)m.ref_nuc, (uint8_t)m.ref_nuc, (uint8_t)m.mut_nuc, (uint8_t)m.is_missing});
}

// What mapper2_body leaves in excess_mutations / imputed_mutations for ONE node with compute_vecs == true
// (reference src/usher_mapper.cpp:190-445), recomputed on the host for the chosen node(s) only.
// Also returns what the call leaves in set_difference / has_unique and whether it would fold the node into the best
// state (:448-455): the sequential search uses it to score the few nodes an earlier graft created or changed.
struct NodeEval { int set_difference = 0; bool has_unique = false; bool valid = false; };
NodeEval placement_vectors(MAT::Node* node, const std::vector<MAT::Mutation>& S, std::vector<MAT::Mutation>& excess,
                           std::vector<MAT::Mutation>& imputed) {
    excess.clear();
    imputed.clear();
    NodeEval ev;
    int node_num_mut = 0, num_common_mut = 0;
    std::vector<MAT::Mutation> anc;
    std::unordered_set<int> seen;
    auto take = [&](const MAT::Mutation& m1, bool to_excess) {
        MAT::Mutation m = m1;
        m.is_missing = false;
        anc.push_back(m);
        seen.insert(m.position);
        if (to_excess) excess.push_back(m);
    };
    if (!node->is_root()) {
        size_t start = 0;
        for (auto& m1 : node->mutations) {
            node_num_mut++;
            if (m1.is_masked()) { ev.has_unique = true; break; }
            bool found = false, found_pos = false;
            for (size_t k = start; k < S.size(); k++) {
                const auto& m2 = S[k];
                start = k;
                if (m1.position == m2.position) {
                    found_pos = true;
                    if (m2.is_missing) { found = true; num_common_mut++; }
                    else if (m2.mut_nuc & m1.mut_nuc) { take(m1, true); found = true; num_common_mut++; break; }
                }
                if (m1.position < m2.position) break;
            }
            if (!found) {
                if (!found_pos && m1.mut_nuc == m1.ref_nuc) { take(m1, true); num_common_mut++; }
                else ev.has_unique = true;
            }
        }
    } else {
        for (auto& m : node->mutations) { anc.push_back(m); seen.insert(m.position); }
    }
    for (MAT::Node* n = node->parent; n; n = n->parent)
        for (auto& m : n->mutations)
            if (!m.is_masked() && !seen.count(m.position)) { anc.push_back(m); seen.insert(m.position); }
    std::stable_sort(anc.begin(), anc.end());
    for (auto& m1 : S) {   // LOOP 2
        if (m1.is_missing) continue;
        bool found_pos = false, found = false;
        const bool has_ref = (m1.mut_nuc & m1.ref_nuc) != 0, amb = (m1.mut_nuc & (m1.mut_nuc - 1)) != 0;
        int8_t anc_nuc = m1.ref_nuc;
        for (auto& m2 : anc) {
            if (m2.is_masked()) continue;
            if (m2.position == m1.position) {
                found_pos = true;
                anc_nuc = m2.mut_nuc;
                found = (m1.mut_nuc & anc_nuc) != 0;
                break;
            }
        }
        MAT::Mutation m;
        m.chrom = m1.chrom; m.position = m1.position; m.ref_nuc = m1.ref_nuc; m.par_nuc = anc_nuc;
        if (found) {
            if (amb) { m.mut_nuc = anc_nuc; imputed.push_back(m); }
        } else if (!found_pos && has_ref) {
            if (amb) { m.mut_nuc = m1.ref_nuc; imputed.push_back(m); }
        } else {
            if (has_ref) m.mut_nuc = m1.ref_nuc;
            else for (int b = 0; b < 4; b++) if (m1.mut_nuc & (1 << b)) { m.mut_nuc = (int8_t)(1 << b); break; }
            if (amb) imputed.push_back(m);
            if (m.mut_nuc != m.par_nuc) { excess.push_back(m); ev.set_difference++; }
        }
    }
    for (auto& m1 : anc) {   // LOOP 3: back-mutations to the reference allele
        if (m1.is_masked()) continue;
        bool found = false, found_pos = false;
        for (auto& m2 : S) {
            if (m2.position != m1.position) continue;
            found_pos = true;
            if (m2.is_missing) { found = true; break; }
            if (m2.mut_nuc & m1.mut_nuc) found = true;
        }
        if (found || found_pos || m1.mut_nuc == m1.ref_nuc) continue;
        MAT::Mutation m;
        m.chrom = m1.chrom; m.position = m1.position; m.ref_nuc = m1.ref_nuc; m.par_nuc = m1.mut_nuc; m.mut_nuc = m1.ref_nuc;
        excess.push_back(m);
        ev.set_difference++;
    }
    const bool leaf = node->is_leaf();
    ev.valid = node->is_root() || (ev.has_unique && !leaf && num_common_mut > 0 && node_num_mut != num_common_mut) ||
               (leaf && num_common_mut > 0) || (!ev.has_unique && !leaf && node_num_mut == num_common_mut);
    return ev;
}

// a before b in Tree::breadth_first_expansion() (level order, children in stored order)
bool bfs_before(const MAT::Node* a, const MAT::Node* b) {
    if (a->level != b->level) return a->level < b->level;
    while (a->parent != b->parent) { a = a->parent; b = b->parent; }
    if (!a->parent) return false;
    for (auto c : a->parent->children) {
        if (c == a) return true;
        if (c == b) return false;
    }
    return false;
}

void die_cuda() {
    fprintf(stderr, "ERROR: %s\n", ub200_last_error());
    // (the device path keeps one reference allele per position: see `usher --help`, "Input restrictions")
    fprintf(stderr, "Samples are scored in batches on the GPU; a batch with a sample the device path cannot take is refused "
                    "as a whole.  Remove the sample named above from the VCF (or fix its REF column) and run again.\n");
    exit(1);
}

}  // namespace

int usher_common(std::string dout_filename, std::string outdir, uint32_t max_trees, uint32_t max_uncertainty,
                 uint32_t max_parsimony, bool sort1, bool sort2, bool sort3, bool reverse_sort, bool collapse_tree,
                 bool collapse_output_tree, bool print_uncondensed_tree, bool print_parsimony_scores,
                 bool retain_original_branch_len, bool no_add, bool detailed_clades, size_t print_subtrees_size,
                 size_t print_subtrees_single, std::vector<Missing_Sample>& missing_samples,
                 std::vector<std::string>& low_confidence_samples, MAT::Tree* loaded_MAT, int device) {
    // ---- flag validation (reference :14-71)
    if (print_subtrees_size == 1) { fprintf(stderr, "ERROR: print-subtrees-size should be larger than 1\n"); return 1; }
    if (print_subtrees_single == 1) { fprintf(stderr, "ERROR: print-subtrees-single should be larger than 1\n"); return 1; }
    if (sort1 && sort2) { fprintf(stderr, "ERROR: Can't use sort-before-placement-1 and sort-before-placement-2 simultaneously. Please specify only one.\n"); return 1; }
    if (sort1 && sort3) { fprintf(stderr, "ERROR: Can't use sort-before-placement-1 and sort-before-placement-3 simultaneously. Please specify only one.\n"); return 1; }
    if (sort2 && sort3) { fprintf(stderr, "ERROR: Can't use sort-before-placement-2 and sort-before-placement-3 simultaneously. Please specify only one.\n"); return 1; }
    if (reverse_sort && !sort1 && !sort2 && !sort3) { fprintf(stderr, "ERROR: Can't use reverse-sort without sorting options (sort-before-placement-1 or sort-before-placement-2 or sort-before-placement-3)\n"); return 1; }
    if (max_trees == 0) { fprintf(stderr, "ERROR: Number of trees specified by --multiple-placements should be >= 1\n"); return 1; }
    if (max_trees >= 256) { fprintf(stderr, "ERROR: Number of trees specified by --multiple-placements should be <= 255\n"); return 1; }
    if (max_trees > 1 || collapse_tree || collapse_output_tree || print_subtrees_size > 1 || print_subtrees_single > 1) {
        fprintf(stderr, "ERROR: --multiple-placements, --collapse-tree, --collapse-output-tree and the subtree writers are "
                        "outside this build's scope (DESIGN.md §7).\n");
        return 1;
    }
    if (max_parsimony == 0 && max_uncertainty == 0) no_add = false;   // nothing would be placed anyway
    mkdir(outdir.c_str(), 0755);

    MAT::Tree* T = loaded_MAT;
    Timer timer;
    fprintf(stderr, "Found %zu missing samples.\n\n", missing_samples.size());
    if (sort3) {
        std::stable_sort(missing_samples.begin(), missing_samples.end());
        if (reverse_sort) std::reverse(missing_samples.begin(), missing_samples.end());
    }
    FILE* parsimony_scores_file = nullptr;
    DeviceTree dev;
    const bool frozen = print_parsimony_scores || no_add;   // every sample sees the same tree version

    if (!missing_samples.empty()) {
        std::vector<size_t> indexes(missing_samples.size());
        std::iota(indexes.begin(), indexes.end(), 0);
        if (print_parsimony_scores) {
            const std::string fn = outdir + "/current-tree.nh";
            fprintf(stderr, "Writing current tree with internal nodes labelled to file %s \n", fn.c_str());
            FILE* f = fopen(fn.c_str(), "w");
            fprintf(f, "%s\n", MAT::get_newick_string(*T, true, true, retain_original_branch_len).c_str());
            fclose(f);
        }
        dev.build(*T, device, frozen || ((sort1 || sort2) && missing_samples.size() > 64));

        // one batched launch set scores every sample against the current tree (sort pre-pass :187-301, and the
        // whole search when the tree is frozen)
        std::vector<ub200_placement> batch(missing_samples.size());
        std::vector<uint64_t> set_ptr(missing_samples.size() + 1, 0);
        std::vector<uint32_t> best_set;
        auto place_all = [&](bool want_set) {
            std::vector<uint64_t> sp{0};
            std::vector<ub200_mutation> calls;
            for (auto& s : missing_samples) { sample_calls(s, calls); sp.push_back(calls.size()); }
            best_set.assign(std::max<size_t>(16, 4 * missing_samples.size()), 0);
            for (;;) {
                int rc = ub200_multi_place_batch(dev.multi, (uint32_t)missing_samples.size(), sp.data(), calls.data(),
                                           want_set ? UB200_WANT_BEST_SET : 0, batch.data(), nullptr,
                                           want_set ? best_set.data() : nullptr, want_set ? set_ptr.data() : nullptr,
                                           best_set.size());
                if (rc == UB200_E_CAPACITY) { best_set.assign(set_ptr[missing_samples.size()] + 16, 0); continue; }
                if (rc != UB200_OK) die_cuda();
                break;
            }
        };
        // ---- sequential mode, batched.  The reference re-searches the whole tree after every graft.
        // A graft creates at most two nodes (the new leaf, and a new internal node when the branch is split) and changes
        // the mutation list of at most one existing node (the split branch); the score and validity of every other node
        // are untouched (its root path carries the same mutations), only tie-break data moves.  So the GPU scores ALL
        // the remaining samples against one frozen tree version in one batched call (whole optimal sets), and per sample
        // the host re-scores the few nodes created or changed since the freeze (placement_vectors = mapper2_body for one
        // node) and redoes the tie-break on the live tree.  After kRefreeze grafts the tree is flattened again.
        const size_t kRefreeze = getenv("UB200_REFREEZE") ? (size_t)atoi(getenv("UB200_REFREEZE")) : 128;
        std::vector<size_t> frozen_of;            // frozen_of[k] = sample index of batch record k (current freeze)
        std::vector<ub200_placement> fbatch;
        std::vector<uint64_t> fset_ptr;
        std::vector<uint32_t> fbest_set;
        std::unordered_map<size_t, size_t> frec;  // sample index -> record of the current freeze
        std::unordered_set<MAT::Node*> changed;   // nodes of the frozen version whose mutation list a graft split
        std::vector<MAT::Node*> created;          // nodes created since the freeze
        bool have_freeze = false;
        bool tree_dirty = false;
        auto freeze_from = [&](size_t idx0) {
            if (have_freeze || tree_dirty) dev.build(*T, device, indexes.size() - idx0 > 64);   // else: the build above
            tree_dirty = false;
            frozen_of.clear(); frec.clear(); changed.clear(); created.clear();
            std::vector<uint64_t> sp{0};
            std::vector<ub200_mutation> calls;
            for (size_t k = idx0; k < indexes.size(); k++) {
                if (T->get_node(missing_samples[indexes[k]].name)) continue;
                frec[indexes[k]] = frozen_of.size();
                frozen_of.push_back(indexes[k]);
                sample_calls(missing_samples[indexes[k]], calls);
                sp.push_back(calls.size());
            }
            fbatch.assign(frozen_of.size(), ub200_placement{});
            fset_ptr.assign(frozen_of.size() + 1, 0);
            fbest_set.assign(std::max<size_t>(16, 4 * frozen_of.size()), 0);
            for (;;) {
                if (frozen_of.empty()) break;
                int rc = ub200_multi_place_batch(dev.multi, (uint32_t)frozen_of.size(), sp.data(), calls.data(), UB200_WANT_BEST_SET,
                                                 fbatch.data(), nullptr, fbest_set.data(), fset_ptr.data(), fbest_set.size());
                if (rc == UB200_E_CAPACITY) { fbest_set.assign(fset_ptr[frozen_of.size()] + 16, 0); continue; }
                if (rc != UB200_OK) die_cuda();
                break;
            }
            have_freeze = true;
        };
        if (!print_parsimony_scores && (sort1 || sort2) && missing_samples.size() > 1) {
            timer.Start();
            fprintf(stderr, "Computing parsimony scores and number of parsimony-optimal placements for new samples and using them to sort the samples.\n");
            for (auto& s : missing_samples) std::sort(s.mutations.begin(), s.mutations.end());   // :203
            place_all(false);
            auto key = [&](size_t i) { return std::make_pair(batch[i].score, batch[i].num_best); };
            if (sort1)
                std::stable_sort(indexes.begin(), indexes.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
            else
                std::stable_sort(indexes.begin(), indexes.end(), [&](size_t a, size_t b) {
                    return std::make_pair(batch[a].num_best, batch[a].score) < std::make_pair(batch[b].num_best, batch[b].score);
                });
            if (reverse_sort) std::reverse(indexes.begin(), indexes.end());
            fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
        }
        if (frozen) place_all(true);
        fprintf(stderr, "Adding missing samples to the tree.\n");

        const std::string stats_fn = outdir + "/placement_stats.tsv";
        FILE* stats = fopen(stats_fn.c_str(), "w");
        std::vector<int32_t> pps_scores;   // -p: per-node scores of samples indexes[pps_first .. pps_first + pps_count)
        size_t pps_first = 0, pps_count = 0;
        std::vector<MAT::Node*> pps_bfs;        // -p: rows go out in BFS order; the tree does not change in this mode, so
        std::vector<uint32_t> pps_bfs_to_dfs;   // the order and its DFS indices are computed once, not per sample
        for (size_t idx = 0; idx < indexes.size(); idx++) {
            timer.Start();
            const size_t s = indexes[idx];
            const std::string& sample = missing_samples[s].name;
            if (T->get_node(sample)) {
                fprintf(stderr, "WARNING: Sample %s already in the tree! Ignoring.\n\n", sample.c_str());
                continue;
            }
            if (print_parsimony_scores && s == 0) {
                const std::string fn = outdir + "/parsimony-scores.tsv";
                fprintf(stderr, "\nNow computing branch parsimony scores for adding the missing samples at each of the %zu nodes in the existing tree without modifying the tree.\n", dev.flat.dfs.size());
                fprintf(stderr, "The branch parsimony scores will be written to file %s\n\n", fn.c_str());
                parsimony_scores_file = fopen(fn.c_str(), "w");
                fprintf(parsimony_scores_file, "#Sample\tTree node\tParsimony score\tOptimal (y/n)\tParsimony-increasing mutations (for optimal nodes)\n");
            }
            // ---- the search: one sample against the current tree version
            ub200_placement res;
            std::vector<std::pair<MAT::Node*, bool>> optn;   // optimal nodes of the CURRENT tree, node_has_unique
            std::vector<int32_t> node_scores;
            size_t current_nodes = dev.flat.dfs.size();
            if (frozen) {
                res = batch[s];
                for (uint64_t k = set_ptr[s]; k < set_ptr[s + 1]; k++)
                    optn.emplace_back(dev.flat.dfs[best_set[k] & 0x7fffffffu], (best_set[k] >> 31) != 0);
            } else {
                if (!have_freeze || created.size() >= 2 * kRefreeze) freeze_from(idx);
                for (int attempt = 0;; attempt++) {
                    const size_t k = frec.at(s);
                    res = fbatch[k];
                    optn.clear();
                    for (uint64_t q = fset_ptr[k]; q < fset_ptr[k + 1]; q++) {
                        MAT::Node* n = dev.flat.dfs[fbest_set[q] & 0x7fffffffu];
                        if (!changed.count(n)) optn.emplace_back(n, (fbest_set[q] >> 31) != 0);
                    }
                    // every optimal node of the frozen version has been split since: the best score over the untouched
                    // nodes is unknown -> flatten the live tree again (rare)
                    if (optn.empty() && attempt == 0) { freeze_from(idx); continue; }
                    break;
                }
                int best_sd = optn.empty() ? INT32_MAX : res.score;
                std::vector<MAT::Mutation> ex, im;
                auto consider = [&](MAT::Node* n) {
                    const NodeEval ev = placement_vectors(n, missing_samples[s].mutations, ex, im);
                    if (!ev.valid || ev.set_difference > best_sd) return;
                    if (ev.set_difference < best_sd) { best_sd = ev.set_difference; optn.clear(); }
                    optn.emplace_back(n, ev.has_unique);
                };
                for (MAT::Node* n : changed) consider(n);
                for (MAT::Node* n : created) consider(n);
                // tie-break on the live tree (:476-493): more leaves first, then the larger BFS index
                size_t bi = 0, bl = 0;
                for (size_t q = 0; q < optn.size(); q++) {
                    const size_t nl = T->get_num_leaves(optn[q].first);
                    if (q == 0 || nl > bl || (nl == bl && bfs_before(optn[bi].first, optn[q].first))) { bi = q; bl = nl; }
                }
                res.score = best_sd;
                res.num_best = (uint32_t)optn.size();
                res.has_unique = optn[bi].second ? 1u : 0u;
                std::swap(optn[0], optn[bi]);               // the chosen node first
                current_nodes += created.size();
            }
            if (print_parsimony_scores) {
                // per-node scores of the next samples in ONE batched call (as many as fit ~1 GB of host memory),
                // sharded over the GPUs in use; the tree is frozen in this mode
                const size_t n_nodes = dev.flat.dfs.size();
                if (idx < pps_first || idx >= pps_first + pps_count) {
                    pps_first = idx;
                    pps_count = std::min<size_t>(indexes.size() - idx, std::max<size_t>(1, ((size_t)1 << 28) / std::max<size_t>(n_nodes, 1)));
                    std::vector<uint64_t> sp{0};
                    std::vector<ub200_mutation> calls;
                    for (size_t k = 0; k < pps_count; k++) { sample_calls(missing_samples[indexes[idx + k]], calls); sp.push_back(calls.size()); }
                    pps_scores.resize(pps_count * n_nodes);
                    std::vector<ub200_placement> tmp(pps_count);
                    if (ub200_multi_place_batch(dev.multi, (uint32_t)pps_count, sp.data(), calls.data(), UB200_WANT_NODE_SCORES,
                                                tmp.data(), pps_scores.data(), nullptr, nullptr, 0) != UB200_OK) die_cuda();
                }
                node_scores.assign(pps_scores.begin() + (idx - pps_first) * n_nodes, pps_scores.begin() + (idx - pps_first + 1) * n_nodes);
            }
            const int best_set_difference = res.score;
            size_t num_best = res.num_best;
            MAT::Node* best_node = frozen ? dev.flat.dfs[res.best_node] : optn[0].first;
            const bool best_node_has_unique = res.has_unique != 0;
            const size_t total_nodes = current_nodes;

            if (!print_parsimony_scores) {
                fprintf(stderr, "Current tree size (#nodes): %zu\tSample name: %s\tParsimony score: %d\tNumber of parsimony-optimal placements: %zu\n",
                        total_nodes, sample.c_str(), best_set_difference, num_best);
                fprintf(stats, "%s\t%d\t%zu\t", sample.c_str(), best_set_difference, num_best);
                if (num_best > 1) {
                    if (max_trees == 1) low_confidence_samples.emplace_back(sample);
                    if (num_best > max_uncertainty)
                        fprintf(stderr, "WARNING: Number of parsimony-optimal placements exceeds maximum allowed value (%u). Ignoring sample %s.\n", max_uncertainty, sample.c_str());
                    else if (best_set_difference <= (int)max_parsimony)
                        fprintf(stderr, "WARNING: Multiple parsimony-optimal placements found. Placement done without high confidence.\n");
                }
                if (best_set_difference > (int)max_parsimony)
                    fprintf(stderr, "WARNING: Parsimony score of the most parsimonious placement exceeds the maximum allowed value (%u). Ignoring sample %s.\n", max_parsimony, sample.c_str());
            } else {
                fprintf(stderr, "Missing sample: %s\t Best parsimony score: %d\tNumber of parsimony-optimal placements: %zu\n",
                        sample.c_str(), best_set_difference, num_best);
            }

            if (print_parsimony_scores) {   // :557-578, rows in BFS order
                if (pps_bfs.empty()) {
                    pps_bfs = T->breadth_first_expansion();
                    std::unordered_map<const MAT::Node*, uint32_t> didx;
                    didx.reserve(dev.flat.dfs.size() * 2);
                    for (size_t i = 0; i < dev.flat.dfs.size(); i++) didx[dev.flat.dfs[i]] = (uint32_t)i;
                    pps_bfs_to_dfs.resize(pps_bfs.size());
                    for (size_t q = 0; q < pps_bfs.size(); q++) pps_bfs_to_dfs[q] = didx.at(pps_bfs[q]);
                }
                std::vector<MAT::Mutation> ex, im;
                for (size_t q = 0; q < pps_bfs.size(); q++) {
                    MAT::Node* const n = pps_bfs[q];
                    const int sc = node_scores[pps_bfs_to_dfs[q]];
                    const bool optimal = sc == best_set_difference;
                    fprintf(parsimony_scores_file, "%s\t%s\t%d\t\t%c\t", sample.c_str(), n->identifier.c_str(), sc, optimal ? 'y' : 'n');
                    if (optimal) {
                        if (sc == 0) fprintf(parsimony_scores_file, "*");
                        placement_vectors(n, missing_samples[s].mutations, ex, im);
                        for (size_t i = 0; i < (size_t)sc && i < ex.size(); i++)
                            fprintf(parsimony_scores_file, "%s%s", ex[i].get_string().c_str(), i + 1 < (size_t)sc ? "," : "");
                    } else {
                        fprintf(parsimony_scores_file, "N/A");
                    }
                    fprintf(parsimony_scores_file, "\n");
                }
            } else if (num_best <= max_uncertainty && best_set_difference <= (int)max_parsimony) {
                // clade assignment over the optimal set (:601-619)
                const size_t na = T->get_num_annotations();
                missing_samples[s].clade_assignments.assign(na, {});
                missing_samples[s].best_clade_assignment.assign(na, "");
                for (size_t c = 0; c < na; c++) {
                    for (auto& on : optn) {
                        MAT::Node* n = on.first;
                        const bool include_self = !n->is_leaf() && !on.second;
                        auto ca = T->get_clade_assignment(n, (int)c, include_self);
                        missing_samples[s].clade_assignments[c].push_back(ca);
                        if (n == best_node) missing_samples[s].best_clade_assignment[c] = ca;
                    }
                    std::sort(missing_samples[s].clade_assignments[c].begin(), missing_samples[s].clade_assignments[c].end());
                }
                std::vector<MAT::Mutation> excess, imputed;
                placement_vectors(best_node, missing_samples[s].mutations, excess, imputed);
                if (!no_add && !T->get_node(sample)) {   // graft (:652-765)
                    auto same = [](const MAT::Mutation& a, const MAT::Mutation& b) {
                        return a.position == b.position && a.mut_nuc == b.mut_nuc;
                    };
                    const std::vector<MAT::Mutation> branch = best_node->mutations;
                    if (best_node->is_leaf() || best_node_has_unique) {   // sibling: split the branch
                        const std::string nid = T->new_internal_node_id();
                        T->create_node(nid, best_node->parent->identifier);
                        T->create_node(sample, nid);
                        T->move_node(best_node->identifier, nid);
                        std::vector<MAT::Mutation> common, l1, l2;
                        for (auto& m1 : branch) {
                            bool found = false;
                            if (!m1.is_masked()) for (auto& m2 : excess) if (same(m1, m2)) { found = true; break; }
                            if (!found) l1.push_back(m1);
                        }
                        for (auto& m1 : excess) {
                            bool found = false;
                            if (!m1.is_masked()) for (auto& m2 : branch) if (same(m1, m2)) { found = true; break; }
                            (found ? common : l2).push_back(m1);
                        }
                        best_node->clear_mutations();
                        for (auto& m : common) T->get_node(nid)->add_mutation(m);
                        for (auto& m : l1) best_node->add_mutation(m);
                        for (auto& m : l2) T->get_node(sample)->add_mutation(m);
                        // a node created since the freeze is re-scored anyway; a frozen-version node now differs from its
                        // device copy
                        if (std::find(created.begin(), created.end(), best_node) == created.end()) changed.insert(best_node);
                        created.push_back(T->get_node(nid));
                        created.push_back(T->get_node(sample));
                    } else {                                              // child
                        MAT::Node* node = T->create_node(sample, best_node->identifier);
                        for (auto& m1 : excess) {
                            bool found = false;
                            if (!m1.is_masked()) for (auto& m2 : branch) if (same(m1, m2)) { found = true; break; }
                            if (!found) node->add_mutation(m1);
                        }
                        created.push_back(node);
                    }
                    tree_dirty = true;
                }
                if (!imputed.empty()) {
                    fprintf(stderr, "Imputed mutations:\t");
                    for (size_t i = 0; i < imputed.size(); i++) {
                        const char* sep = i + 1 < imputed.size() ? ";" : "";
                        fprintf(stderr, "%i:%c%s", imputed[i].position, MAT::get_nuc(imputed[i].mut_nuc), sep);
                        fprintf(stats, "%i:%c%s", imputed[i].position, MAT::get_nuc(imputed[i].mut_nuc), sep);
                    }
                    fprintf(stderr, "\n");
                }
            }
            if (!print_parsimony_scores) fputc('\n', stats);
            fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
        }
        fclose(stats);
    }

    if (print_parsimony_scores) {
        if (parsimony_scores_file) fclose(parsimony_scores_file);
        return 0;
    }

    {   // final tree (:855-881)
        timer.Start();
        const std::string fn = outdir + (print_uncondensed_tree ? "/uncondensed-final-tree.nh" : "/final-tree.nh");
        fprintf(stderr, "Writing %sfinal tree to file %s \n", print_uncondensed_tree ? "uncondensed " : "", fn.c_str());
        fprintf(stderr, "The parsimony score for this tree is: %zu \n", T->get_parsimony_score());
        FILE* f = fopen(fn.c_str(), "w");
        fputs(MAT::get_newick_string(*T, T->root, true, true, retain_original_branch_len, print_uncondensed_tree).c_str(), f);
        fclose(f);
        fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
    }
    if (!missing_samples.empty()) {
        timer.Start();
        std::vector<std::string> targets;
        for (auto& s : missing_samples) targets.push_back(s.name);
        const std::string fn = outdir + "/mutation-paths.txt";
        fprintf(stderr, "Writing mutation paths to file %s \n", fn.c_str());
        MAT::get_sample_mutation_paths(T, targets, fn);
        fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
        const size_t na = T->get_num_annotations();
        if (na > 0) {   // clades.txt (:908-970)
            const std::string cfn = outdir + "/clades.txt";
            fprintf(stderr, "Writing clade annotations to file %s \n", cfn.c_str());
            FILE* f = fopen(cfn.c_str(), "w");
            for (auto& ms : missing_samples) {
                if (ms.best_clade_assignment.empty()) continue;
                fprintf(f, "%s\t", ms.name.c_str());
                for (size_t k = 0; k < na; k++) {
                    fprintf(f, "%s", ms.best_clade_assignment[k].c_str());
                    if (detailed_clades) {
                        fprintf(f, "*|");
                        std::string cur;
                        int cnt = 0;
                        bool first = true;
                        auto flush = [&](bool last) {
                            if (cnt > 0) fprintf(f, "%s(%i/%zu)%s", cur.c_str(), cnt, ms.clade_assignments[k].size(), last ? "" : ",");
                            (void)first;
                        };
                        for (auto& cl : ms.clade_assignments[k]) {
                            if (cl == cur) cnt++;
                            else { flush(false); cur = cl; cnt = 1; }
                        }
                        flush(true);
                    }
                    if (k + 1 < na) fprintf(f, "\t");
                }
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    if (!low_confidence_samples.empty()) {
        fprintf(stderr, "WARNING: Following samples had multiple possibilities of parsimony-optimal placements:\n");
        for (auto& l : low_confidence_samples) fprintf(stderr, "%s\n", l.c_str());
    }
    if (!dout_filename.empty()) {   // :1024-1044
        timer.Start();
        fprintf(stderr, "Saving mutation-annotated tree object to file (after condensing identical sequences) %s\n", dout_filename.c_str());
        if (!T->condensed_nodes.empty()) T->uncondense_leaves();
        T->condense_leaves();
        MAT::save_mutation_annotated_tree(*T, dout_filename);
        fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
    }
    return 0;
}
