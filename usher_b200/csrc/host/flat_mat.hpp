// N1: parsimony.proto <-> the flat SoA of include/usher_b200.h WITHOUT building Node objects.
// A 10 M-node tree costs > 13 GB as MAT::Node / MAT::Mutation objects (a std::string per mutation) and the
// pb -> Node -> flatten() detour dominates the end-to-end time of a frozen-tree batch; here the newick is turned
// straight into parent[] (node order = creation order of the reference's parser = DFS pre-order,
// src/mutation_annotated_tree.cpp:215-356, 553-596) and the k-th mutation list into row k of the CSR.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "usher_b200.h"

namespace Mutation_Annotated_Tree {

struct FlatTree {
    std::vector<int32_t> parent;          // [n] DFS index of the parent, -1 for the root
    std::vector<uint64_t> row_ptr;        // [n+1]
    std::vector<ub200_mutation> muts;     // rows in stored (position-sorted) order; position < 0 = masked
    std::vector<std::string> names;       // [n] leaf names as written, internal nodes node_1.. in '(' order
    std::vector<uint32_t> n_children;     // [n]
    std::string chrom;                    // the one chromosome name of every mutation ("" if none is stored)
    std::vector<std::pair<std::string, std::vector<std::string>>> condensed;   // in file order
    std::vector<std::vector<std::string>> annotations;                         // [n] clade annotations (metadata)
    bool have_metadata = false;
    ub200_flat_mat view() const {
        return ub200_flat_mat{(uint32_t)parent.size(), (uint64_t)muts.size(), parent.data(), row_ptr.data(), muts.data(), nullptr};
    }
};

// Load a .pb / .pb.gz straight into the SoA.  Returns false (and fills err) on a malformed file or when the file
// uses more than one chromosome name (the flat mutation carries none).
bool load_flat_mutation_annotated_tree(const std::string& filename, FlatTree& out, std::string& err);
// Write it back: byte-identical to save_mutation_annotated_tree() of the same tree.
bool save_flat_mutation_annotated_tree(const FlatTree& t, const std::string& filename, std::string& err);

}  // namespace Mutation_Annotated_Tree
