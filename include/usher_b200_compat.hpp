// usher_b200_compat.hpp — header-only C++ adapter for the OTHER callers of mapper2_body:
//   src/matUtils/uncertainty.cpp:214-246, src/matUtils/annotate.cpp:613-637, src/matUtils/merge.cpp:255-300,
//   src/ripples/main.cpp:350-380.
// Each of them fills one mapper2_input per node of `dfs = T.depth_first_expansion()` with `inp.j = k` (the DFS index is
// their tie index), calls mapper2_body(inp, false) inside a tbb::parallel_for and reads best_node / best_set_difference /
// num_best / best_j_vec / node_has_unique back.  This adapter gives them the same results from one batched GPU call
// (include/usher_b200.h); it is written against the MAT API only (Tree::depth_first_expansion, Node::parent / mutations,
// Mutation::position / ref_nuc / par_nuc / mut_nuc / is_missing), so it compiles against the reference's
// mutation_annotated_tree.hpp and against usher_b200/csrc/host/mutation_annotated_tree.hpp alike.
//
//   ub200_compat::Searcher<MAT::Tree, MAT::Node, MAT::Mutation> search(T);            // flatten + stage once per tree
//   auto r = search.place(ancestral_mutations);                                        // one sample
//   auto rs = search.place_all(list_of_mutation_vectors);                              // or a batch, all GPUs
//   r.best_node, r.best_set_difference, r.num_best, r.best_j, r.best_node_has_unique, r.best_j_vec, r.node_has_unique
//
// Not covered: uncertainty.cpp's "skip the sample's own node" filter (:216) — a caller that needs it compares
// r.best_j_vec with the excluded index and falls back to its own loop when the excluded node is the only optimum.
#ifndef USHER_B200_COMPAT_HPP
#define USHER_B200_COMPAT_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "usher_b200.h"

namespace ub200_compat {

template <class Node>
struct Result {
    Node* best_node = nullptr;
    int best_set_difference = 0;
    size_t best_j = 0, num_best = 0;
    bool best_node_has_unique = false;
    std::vector<size_t> best_j_vec;      // every parsimony-optimal node, as indices into dfs (ascending)
    std::vector<bool> node_has_unique;   // parallel to best_j_vec
};

template <class Tree, class Node, class Mutation>
class Searcher {
  public:
    // tie index = DFS index, as the matUtils / ripples callers pass it; all visible GPUs
    explicit Searcher(const Tree& T) : dfs_(T.depth_first_expansion()) {
        std::unordered_map<const Node*, int32_t> idx;
        idx.reserve(dfs_.size() * 2);
        for (size_t i = 0; i < dfs_.size(); i++) idx[dfs_[i]] = (int32_t)i;
        std::vector<int32_t> parent(dfs_.size());
        std::vector<uint64_t> row_ptr(1, 0);
        std::vector<ub200_mutation> muts;
        std::vector<uint32_t> tie(dfs_.size());
        for (size_t i = 0; i < dfs_.size(); i++) {
            const Node* n = dfs_[i];
            parent[i] = n->parent ? idx[n->parent] : -1;
            tie[i] = (uint32_t)i;
            for (auto& m : n->mutations)
                muts.push_back({m.position, (uint8_t)m.ref_nuc, (uint8_t)m.par_nuc, (uint8_t)m.mut_nuc, 0});
            row_ptr.push_back(muts.size());
        }
        ub200_flat_mat v{(uint32_t)dfs_.size(), (uint64_t)muts.size(), parent.data(), row_ptr.data(), muts.data(), tie.data()};
        if (ub200_multi_create(&v, 0, nullptr, &multi_) != UB200_OK) throw std::runtime_error(ub200_last_error());
    }
    ~Searcher() { ub200_multi_destroy(multi_); }
    Searcher(const Searcher&) = delete;
    Searcher& operator=(const Searcher&) = delete;

    const std::vector<Node*>& dfs() const { return dfs_; }

    std::vector<Result<Node>> place_all(const std::vector<std::vector<Mutation>>& samples) const {
        std::vector<uint64_t> sp(1, 0);
        std::vector<ub200_mutation> calls;
        for (auto& s : samples) {
            for (auto& m : s) calls.push_back({m.position, (uint8_t)m.ref_nuc, (uint8_t)m.ref_nuc, (uint8_t)m.mut_nuc, (uint8_t)m.is_missing});
            sp.push_back(calls.size());
        }
        std::vector<ub200_placement> res(samples.size());
        std::vector<uint64_t> set_ptr(samples.size() + 1, 0);
        std::vector<uint32_t> set(std::max<size_t>(64, 8 * samples.size()));
        for (;;) {
            const int rc = ub200_multi_place_batch(multi_, (uint32_t)samples.size(), sp.data(), calls.data(), UB200_WANT_BEST_SET,
                                                   res.data(), nullptr, set.data(), set_ptr.data(), set.size());
            if (rc == UB200_E_CAPACITY) { set.resize(set_ptr[samples.size()] + 16); continue; }
            if (rc != UB200_OK) throw std::runtime_error(ub200_last_error());
            break;
        }
        std::vector<Result<Node>> out(samples.size());
        for (size_t i = 0; i < samples.size(); i++) {
            Result<Node>& r = out[i];
            r.best_node = dfs_[res[i].best_node];
            r.best_set_difference = res[i].score;
            r.best_j = res[i].best_j;
            r.num_best = res[i].num_best;
            r.best_node_has_unique = res[i].has_unique != 0;
            for (uint64_t k = set_ptr[i]; k < set_ptr[i + 1]; k++) {
                r.best_j_vec.push_back(set[k] & 0x7fffffffu);
                r.node_has_unique.push_back((set[k] >> 31) != 0);
            }
        }
        return out;
    }
    Result<Node> place(const std::vector<Mutation>& sample) const { return place_all({sample})[0]; }

  private:
    std::vector<Node*> dfs_;
    ub200_multi* multi_ = nullptr;
};

}  // namespace ub200_compat
#endif
