// `usher` — drop-in CLI for the placement path (flags of reference src/usher.cpp:44-107; own parser, no Boost).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "flat_mat.hpp"
#include "usher_b200.h"
#include "usher_common.hpp"

namespace {
struct Opt { const char* lng; char sht; bool arg; const char* help; };
const Opt kOpts[] = {
    {"vcf", 'v', true, "Input VCF file (in uncompressed or gzip-compressed .gz format) [REQUIRED]"},
    {"tree", 't', true, "Input tree file"},
    {"outdir", 'd', true, "Output directory to dump output and log files [DEFAULT uses current directory]"},
    {"load-mutation-annotated-tree", 'i', true, "Load mutation-annotated tree object"},
    {"save-mutation-annotated-tree", 'o', true, "Save output mutation-annotated tree object to the specified filename"},
    {"sort-before-placement-1", 's', false, "Sort new samples based on computed parsimony score and then number of optimal placements before the actual placement"},
    {"sort-before-placement-2", 'S', false, "Sort new samples based on the number of optimal placements and then the parsimony score before the actual placement"},
    {"sort-before-placement-3", 'A', false, "Sort new samples based on the number of ambiguous bases"},
    {"reverse-sort", 'r', false, "Reverse the sorting order of sorting options"},
    {"collapse-tree", 'c', false, "(not supported in this build)"},
    {"collapse-output-tree", 'C', false, "(not supported in this build)"},
    {"max-uncertainty-per-sample", 'e', true, "Maximum number of equally parsimonious placements allowed per sample beyond which the sample is ignored"},
    {"max-parsimony-per-sample", 'E', true, "Maximum parsimony score of the most parsimonious placement(s) allowed per sample beyond which the sample is ignored"},
    {"write-uncondensed-final-tree", 'u', false, "Write the final tree in uncondensed format and save to file uncondensed-final-tree.nh in outdir"},
    {"write-subtrees-size", 'k', true, "(not supported in this build)"},
    {"write-single-subtree", 'K', true, "(not supported in this build)"},
    {"write-parsimony-scores-per-node", 'p', false, "Write the parsimony scores for adding new samples at each existing node in the tree without modifying the tree in a file names parsimony-scores.tsv in outdir"},
    {"multiple-placements", 'M', true, "Create a new tree up to this limit for each possibility of parsimony-optimal placement (only 1 is supported in this build)"},
    {"retain-input-branch-lengths", 'l', false, "Retain the branch lengths from the input tree in out newick files instead of using number of mutations for the branch lengths."},
    {"no-add", 'n', false, "Do not add new samples to the tree"},
    {"detailed-clades", 'D', false, "In clades.txt, write a histogram of annotated clades and counts across all equally parsimonious placements"},
    {"threads", 'T', true, "Accepted for compatibility (the search runs on the GPU)"},
    {"flat-resave", 0, true, "Load the -i protobuf straight into the flat SoA (no Node objects) and write it back to this file"},
    {"place-flat", 0, false, "Frozen-tree placement of the VCF's new samples on the -i protobuf loaded straight into the flat SoA (no Node objects); writes flat-placements.tsv"},
    {"device", 0, true, "CUDA device ordinal [DEFAULT: every visible GPU for frozen-tree batches (-n, -p, the -s/-S pre-pass), GPU 0 otherwise]"},
    {"dump-flat", 0, true, "(diagnostic) write the loaded tree and VCF samples as text and exit; needs no GPU"},
    {"resave", 0, true, "(diagnostic) write the loaded tree back as protobuf and exit; needs no GPU"},
    {"version", 0, false, "Print version number"},
    {"help", 'h', false, "Print help messages"},
};
void usage() {
    fprintf(stderr, "UShER (usher_b200 placement build)\nOptions:\n");
    for (auto& o : kOpts) {
        if (o.sht) fprintf(stderr, "  -%c [ --%s ]%s  %s\n", o.sht, o.lng, o.arg ? " arg" : "", o.help);
        else fprintf(stderr, "  --%s%s  %s\n", o.lng, o.arg ? " arg" : "", o.help);
    }
    fprintf(stderr,
            "Input restrictions of this build (the reference scores such input on the CPU, src/usher_mapper.cpp):\n"
            "  * every VCF REF allele of a placed sample is ONE base A/C/G/T, and equals the reference allele of the\n"
            "    tree's mutations at that position (the device keeps one reference allele per position);\n"
            "  * a sample that breaks this makes the run stop with the sample and position named -- nothing is placed\n"
            "    on a CPU fallback.\n");
}
}  // namespace

int main(int argc, char** argv) {
    std::string vcf, tree_fn, outdir = ".", din, dout, dump_flat, resave, flat_resave;
    bool place_flat = false;
    bool s1 = false, s2 = false, s3 = false, rev = false, ct = false, cot = false, unc = false, pps = false, keep = false,
         no_add = false, detailed = false;
    uint32_t max_trees = 1, max_unc = 1000000, max_pars = 1000000;
    size_t sub_k = 0, sub_K = 0;
    int device = -1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i], val;
        const Opt* hit = nullptr;
        if (a.rfind("--", 0) == 0) {
            auto eq = a.find('=');
            std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            for (auto& o : kOpts) if (name == o.lng) hit = &o;
            if (hit && hit->arg) {
                if (eq != std::string::npos) val = a.substr(eq + 1);
                else if (i + 1 < argc) val = argv[++i];
                else { fprintf(stderr, "the required argument for option '--%s' is missing\n", hit->lng); return 1; }
            }
        } else if (a.size() >= 2 && a[0] == '-') {
            for (auto& o : kOpts) if (o.sht && a[1] == o.sht) hit = &o;
            if (hit && hit->arg) {
                if (a.size() > 2) val = a.substr(2);
                else if (i + 1 < argc) val = argv[++i];
                else { fprintf(stderr, "the required argument for option '-%c' is missing\n", hit->sht); return 1; }
            }
        }
        if (!hit) { fprintf(stderr, "unrecognised option '%s'\n", a.c_str()); usage(); return 1; }
        const std::string n = hit->lng;
        if (n == "vcf") vcf = val; else if (n == "tree") tree_fn = val; else if (n == "outdir") outdir = val;
        else if (n == "load-mutation-annotated-tree") din = val; else if (n == "save-mutation-annotated-tree") dout = val;
        else if (n == "sort-before-placement-1") s1 = true; else if (n == "sort-before-placement-2") s2 = true;
        else if (n == "sort-before-placement-3") s3 = true; else if (n == "reverse-sort") rev = true;
        else if (n == "collapse-tree") ct = true; else if (n == "collapse-output-tree") cot = true;
        else if (n == "max-uncertainty-per-sample") max_unc = (uint32_t)std::stoul(val);
        else if (n == "max-parsimony-per-sample") max_pars = (uint32_t)std::stoul(val);
        else if (n == "write-uncondensed-final-tree") unc = true;
        else if (n == "write-subtrees-size") sub_k = std::stoul(val); else if (n == "write-single-subtree") sub_K = std::stoul(val);
        else if (n == "write-parsimony-scores-per-node") pps = true;
        else if (n == "multiple-placements") max_trees = (uint32_t)std::stoul(val);
        else if (n == "retain-input-branch-lengths") keep = true; else if (n == "no-add") no_add = true;
        else if (n == "detailed-clades") detailed = true; else if (n == "threads") {}
        else if (n == "device") device = std::stoi(val);
        else if (n == "dump-flat") dump_flat = val; else if (n == "resave") resave = val;
        else if (n == "flat-resave") flat_resave = val; else if (n == "place-flat") place_flat = true;
        else if (n == "version") { printf("UShER usher_b200 (placement build)\n"); return 0; }
        else if (n == "help") { usage(); return 0; }
    }
    if (!flat_resave.empty() || place_flat) {
        // ---- N1: protobuf -> flat SoA -> GPU, no MAT::Node objects anywhere
        if (din.empty()) { fprintf(stderr, "--flat-resave / --place-flat need --load-mutation-annotated-tree\n"); return 1; }
        MAT::FlatTree ft;
        std::string err;
        Timer tm;
        tm.Start();
        if (!MAT::load_flat_mutation_annotated_tree(din, ft, err)) { fprintf(stderr, "ERROR: %s\n", err.c_str()); return 1; }
        fprintf(stderr, "Loaded %zu nodes, %zu mutations straight into the flat form in %ld msec\n", ft.parent.size(), ft.muts.size(), tm.Stop());
        if (!flat_resave.empty()) {
            if (!MAT::save_flat_mutation_annotated_tree(ft, flat_resave, err)) { fprintf(stderr, "ERROR: %s\n", err.c_str()); return 1; }
            if (!place_flat) return 0;
        }
        if (vcf.empty()) { fprintf(stderr, "the option '--vcf' is required but missing\n"); return 1; }
        std::unordered_set<std::string> known(ft.names.begin(), ft.names.end());
        for (auto& c : ft.condensed) for (auto& m : c.second) known.insert(m);
        std::vector<Missing_Sample> missing;
        MAT::read_vcf_samples(vcf, [&](const std::string& nm) { return known.count(nm) != 0; }, missing);
        fprintf(stderr, "Found %zu missing samples.\n", missing.size());
        std::vector<uint64_t> sp{0};
        std::vector<ub200_mutation> calls;
        for (auto& s : missing) {
            for (auto& m : s.mutations)
                calls.push_back({m.position, (uint8_t)m.ref_nuc, (uint8_t)m.ref_nuc, (uint8_t)m.mut_nuc, (uint8_t)m.is_missing});
            sp.push_back(calls.size());
        }
        ub200_flat_mat v = ft.view();
        ub200_multi* multi = nullptr;
        const int one = device < 0 ? 0 : device;
        tm.Start();
        if (ub200_multi_create(&v, device < 0 ? 0 : 1, device < 0 ? nullptr : &one, &multi) != UB200_OK) {
            fprintf(stderr, "ERROR: %s\n", ub200_last_error());
            return 1;
        }
        for (int i = 0; i < ub200_multi_size(multi); i++) ub200_mat_set_pass_samples(ub200_multi_mat(multi, i), 96);
        fprintf(stderr, "Tree resident on %d GPU(s) in %ld msec\n", ub200_multi_size(multi), tm.Stop());
        std::vector<ub200_placement> res(missing.size());
        tm.Start();
        if (!missing.empty() && ub200_multi_place_batch(multi, (uint32_t)missing.size(), sp.data(), calls.data(), 0, res.data(),
                                                        nullptr, nullptr, nullptr, 0) != UB200_OK) {
            fprintf(stderr, "ERROR: %s\n", ub200_last_error());
            return 1;
        }
        fprintf(stderr, "Placed %zu samples in %ld msec\n", missing.size(), tm.Stop());
        const std::string fn = outdir + "/flat-placements.tsv";
        FILE* f = fopen(fn.c_str(), "w");
        if (!f) { fprintf(stderr, "ERROR: cannot write %s\n", fn.c_str()); return 1; }
        fprintf(f, "#Sample\tParsimony score\tNumber of parsimony-optimal placements\tBest node\tSibling (1) or child (0)\n");
        for (size_t i = 0; i < missing.size(); i++)
            fprintf(f, "%s\t%d\t%u\t%s\t%u\n", missing[i].name.c_str(), res[i].score, res[i].num_best,
                    ft.names[res[i].best_node].c_str(), res[i].has_unique);
        fclose(f);
        ub200_multi_destroy(multi);
        return 0;
    }
    if (vcf.empty() && resave.empty()) { fprintf(stderr, "the option '--vcf' is required but missing\n"); usage(); return 1; }
    MAT::Tree T;
    Timer timer;
    bool from_newick = false;
    // the reference takes the tree + VCF build path when both -t and -i are given (src/usher.cpp:130)
    if (!tree_fn.empty()) {
        fprintf(stderr, "Loading input tree.\n");
        timer.Start();
        T = MAT::create_tree_from_newick(tree_fn);
        if (!T.root) { fprintf(stderr, "ERROR: Empty tree.\n"); return 1; }
        fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
        from_newick = true;
    } else if (!din.empty()) {
        timer.Start();
        fprintf(stderr, "Loading existing mutation-annotated tree object from file %s\n", din.c_str());
        T = MAT::load_mutation_annotated_tree(din);
        fprintf(stderr, "Completed in %ld msec \n\n", timer.Stop());
    } else {
        fprintf(stderr, "Error! No input tree or assignment file provided!\n");
        return 1;
    }
    if (!resave.empty()) { MAT::save_mutation_annotated_tree(T, resave); return 0; }
    std::vector<Missing_Sample> missing;
    MAT::read_vcf(&T, vcf, missing, from_newick);
    if (!dump_flat.empty()) {
        FILE* f = fopen(dump_flat.c_str(), "w");
        for (auto n : T.depth_first_expansion()) {
            fprintf(f, "N\t%s\t%s\t", n->identifier.c_str(), n->parent ? n->parent->identifier.c_str() : "");
            for (auto& m : n->mutations) fprintf(f, "%d:%d:%d:%d,", m.position, m.ref_nuc, m.par_nuc, m.mut_nuc);
            fprintf(f, "\n");
        }
        for (auto& s : missing) {
            fprintf(f, "S\t%s\t", s.name.c_str());
            for (auto& m : s.mutations) fprintf(f, "%d:%d:%d:%d,", m.position, m.ref_nuc, m.mut_nuc, (int)m.is_missing);
            fprintf(f, "\n");
        }
        for (auto& s : missing) fprintf(f, "A\t%s\t%zu\n", s.name.c_str(), s.num_ambiguous);   // the -A sort key
        for (auto& name : T.condensed_order) {
            fprintf(f, "C\t%s\t", name.c_str());
            for (auto& m : T.condensed_nodes.at(name)) fprintf(f, "%s,", m.c_str());
            fprintf(f, "\n");
        }
        fprintf(f, "NEWICK\t%s\n", MAT::get_newick_string(T, true, true).c_str());
        fclose(f);
        return 0;
    }
    std::vector<std::string> low_conf;
    return usher_common(dout, outdir, max_trees, max_unc, max_pars, s1, s2, s3, rev, ct, cot, unc, pps, keep, no_add,
                        detailed, sub_k, sub_K, missing, low_conf, &T, device);
}
