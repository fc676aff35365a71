// Test program for include/usher_b200_compat.hpp: what a matUtils-style caller of mapper2_body (tie index j = DFS
// index) gets through the adapter.  usage: compat_main tree.pb samples.vcf  -> one line per new sample:
//   name <tab> best node <tab> score <tab> num_best <tab> best_j <tab> has_unique <tab> best_j_vec (comma separated)
#include <cstdio>

#include "mutation_annotated_tree.hpp"
#include "usher_b200_compat.hpp"
#include "usher_graph.hpp"
namespace MAT = Mutation_Annotated_Tree;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    MAT::Tree T = MAT::load_mutation_annotated_tree(argv[1]);
    std::vector<Missing_Sample> missing;
    MAT::read_vcf(&T, argv[2], missing, false);
    ub200_compat::Searcher<MAT::Tree, MAT::Node, MAT::Mutation> search(T);
    std::vector<std::vector<MAT::Mutation>> samples;
    for (auto& s : missing) samples.push_back(s.mutations);
    auto res = search.place_all(samples);
    for (size_t i = 0; i < res.size(); i++) {
        printf("%s\t%s\t%d\t%zu\t%zu\t%d\t", missing[i].name.c_str(), res[i].best_node->identifier.c_str(), res[i].best_set_difference,
               res[i].num_best, res[i].best_j, (int)res[i].best_node_has_unique);
        for (size_t k = 0; k < res[i].best_j_vec.size(); k++) printf("%s%zu", k ? "," : "", res[i].best_j_vec[k]);
        printf("\n");
    }
    return 0;
}
