// Host-side data model of the drop-in `usher` binary: the Mutation_Annotated_Tree API surface of the reference
// (names, fields and observable behaviour of src/mutation_annotated_tree.hpp:34-185) re-implemented on plain
// C++17 + zlib — no TBB, Boost or libprotobuf.  Only what the placement path of `usher` touches is provided
// (SURVEY.md §2 #4-#7, §9); matUtils/matOptimize-only helpers are out of scope.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

struct Missing_Sample;

namespace Mutation_Annotated_Tree {

// one-hot A=1 C=2 G=4 T=8, IUPAC codes as bit sets (reference src/mutation_annotated_tree.cpp:17-139)
int8_t get_nuc_id(char nuc);
int8_t get_nuc_id(const std::vector<int8_t>& nuc_vec);
char get_nuc(int8_t nuc_id);
int8_t get_nt(int8_t nuc_id);
std::vector<int8_t> get_nuc_vec_from_id(int8_t nuc_id);

struct Mutation {
    std::string chrom;
    int position = 0;
    int8_t ref_nuc = 0;
    int8_t par_nuc = 0;
    int8_t mut_nuc = 0;
    bool is_missing = false;
    bool operator<(const Mutation& m) const { return position < m.position; }
    Mutation copy() const { return *this; }
    bool is_masked() const { return position < 0; }
    std::string get_string() const {
        if (is_masked()) return "MASKED";
        return get_nuc(par_nuc) + std::to_string(position) + get_nuc(mut_nuc);
    }
};

class Node {
  public:
    size_t level = 0;
    float branch_length = -1.0f;
    std::string identifier;
    std::vector<std::string> clade_annotations;
    Node* parent = nullptr;
    std::vector<Node*> children;
    std::vector<Mutation> mutations;
    bool is_leaf() const { return children.empty(); }
    bool is_root() const { return parent == nullptr; }
    // keeps the row position-sorted; a second mutation at a position overwrites the allele or cancels the entry
    // when it reverts to the entry's par_nuc (reference :720-752)
    void add_mutation(const Mutation& mut);
    void clear_mutations() { mutations.clear(); }
};

class Tree {
  public:
    Tree() = default;
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;
    Tree(Tree&& o) noexcept { *this = std::move(o); }
    Tree& operator=(Tree&& o) noexcept;
    ~Tree();

    Node* root = nullptr;
    std::unordered_map<std::string, std::vector<std::string>> condensed_nodes;
    std::vector<std::string> condensed_order;   // insertion order of condensed_nodes (deterministic iteration)
    std::unordered_set<std::string> condensed_leaves;
    size_t curr_internal_node = 0;

    std::string new_internal_node_id() { return "node_" + std::to_string(++curr_internal_node); }
    size_t get_num_annotations() const { return root ? root->clade_annotations.size() : 0; }
    Node* create_node(const std::string& identifier, float branch_length = -1.0f, size_t num_annotations = 0);
    Node* create_node(const std::string& identifier, Node* par, float branch_length = -1.0f);
    Node* create_node(const std::string& identifier, const std::string& parent_id, float branch_length = -1.0f);
    Node* get_node(const std::string& identifier) const;
    void reserve_nodes(size_t n) { all_nodes.reserve(n); }   // (loader hint; not in the reference's API)
    std::vector<Node*> rsearch(const std::string& nid, bool include_self = false) const;
    std::string get_clade_assignment(const Node* n, int clade_id, bool include_self = true) const;
    size_t get_num_leaves(Node* node = nullptr) const;
    std::vector<Node*> get_leaves(const std::string& nid = "") const;
    void move_node(const std::string& source, const std::string& destination);
    void remove_node(const std::string& nid, bool move_level);
    std::vector<Node*> breadth_first_expansion(const std::string& nid = "") const;
    std::vector<Node*> depth_first_expansion(Node* node = nullptr) const;
    size_t get_parsimony_score() const;
    void condense_leaves(const std::vector<std::string>& missing_samples = {});
    void uncondense_leaves();

  private:
    std::unordered_map<std::string, Node*> all_nodes;
};

std::string get_newick_string(const Tree& T, bool print_internal, bool print_branch_len,
                              bool retain_original_branch_len = false, bool uncondense_leaves = false);
std::string get_newick_string(const Tree& T, Node* node, bool print_internal, bool print_branch_len,
                              bool retain_original_branch_len = false, bool uncondense_leaves = false);
Tree create_tree_from_newick(const std::string& filename);
Tree create_tree_from_newick_string(const std::string& newick_string);
void string_split(const std::string& s, char delim, std::vector<std::string>& words);
void string_split(const std::string& s, std::vector<std::string>& words);

// parsimony.proto (hand-rolled wire codec; ".gz" anywhere in the name = gzip, as the reference)
Tree load_mutation_annotated_tree(const std::string& filename);
void save_mutation_annotated_tree(const Tree& tree, const std::string& filename);

void get_sample_mutation_paths(Tree* T, const std::vector<std::string>& samples, const std::string& filename);
// placement mode only (create_new_mat == false): fills Missing_Sample::mutations in VCF row order
// placement-mode VCF reader on a membership predicate (the flat path has no Tree): samples already in the tree are skipped
void read_vcf_samples(const std::string& vcf_filename, const std::function<bool(const std::string&)>& in_tree,
                      std::vector<Missing_Sample>& missing_samples);
void read_vcf(Tree* T, const std::string& vcf_filename, std::vector<Missing_Sample>& missing_samples,
              bool create_new_mat);

}  // namespace Mutation_Annotated_Tree
