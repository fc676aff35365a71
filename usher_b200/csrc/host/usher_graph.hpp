// Placement-side structs of the drop-in `usher` binary (surface of reference src/usher_graph.hpp:15-53).
#pragma once
#include <sys/time.h>

#include <string>
#include <vector>

#include "mutation_annotated_tree.hpp"

namespace MAT = Mutation_Annotated_Tree;

class Timer {
    struct timeval a_, b_;
  public:
    void Start() { gettimeofday(&a_, nullptr); }
    long Stop() {
        gettimeofday(&b_, nullptr);
        return (long)((b_.tv_sec - a_.tv_sec) * 1000 + (b_.tv_usec - a_.tv_usec) / 1000.0 + 0.5);
    }
};

struct Missing_Sample {
    std::string name;
    std::vector<MAT::Mutation> mutations;
    size_t num_ambiguous = 0;
    std::vector<std::string> best_clade_assignment;
    std::vector<std::vector<std::string>> clade_assignments;
    explicit Missing_Sample(const std::string& n) : name(n) {}
    bool operator==(const Missing_Sample& o) const { return name == o.name; }
    bool operator<(const Missing_Sample& o) const { return num_ambiguous < o.num_ambiguous; }
};
