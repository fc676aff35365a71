/* usher_b200.h — C ABI of the B200-native sample-placement engine.
 *
 * This is the drop-in boundary for ONE path of yatisht/usher: the per-candidate-node parsimony scoring
 * `mapper2_body` (reference src/usher_mapper.cpp:167-504) together with the per-sample search/reduction
 * that usher_common() wraps around it (reference src/usher_common.cpp:342-453, the three
 * `tbb::parallel_for` bodies at :252-273, :389-414, :426-449).  The reference has no plugin API for this
 * path; the seam is "fill a mapper2_input per node, call mapper2_body, read best_* back"
 * (src/usher_graph.hpp:73-103).  The entry points below replace that seam with one batched call on a
 * frozen tree: all candidate nodes x a batch of samples -> per sample (best score, best node, number of
 * optimal placements, sibling/child flag), optionally every node's score (`-p`,
 * src/usher_common.cpp:557-578) and the full optimal set (best_j_vec / node_has_unique).
 *
 * Plain C types only; all pointers are caller-owned host memory unless the name says `_dev`.  The library
 * owns its device memory behind opaque handles.  No entry point calls exit(); every one returns an int
 * status (0 = ok, <0 = invalid input, >0 = CUDA runtime error code) and ub200_last_error() describes the
 * last failure of the calling thread.  There is NO CPU fallback: without a CUDA device every call that
 * would compute fails with UB200_E_NO_DEVICE.
 */
#ifndef USHER_B200_H
#define USHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB200_ABI_VERSION 1

/* status codes */
#define UB200_OK 0
#define UB200_E_ARG (-1)          /* NULL / inconsistent argument */
#define UB200_E_TREE_ORDER (-2)   /* nodes not in DFS pre-order / bad parent[] */
#define UB200_E_NOT_ONE_HOT (-3)  /* tree mutation with a multi-bit or zero nucleotide (outside the
                                     reference's asserted domain, src/usher_mapper.cpp:201) */
#define UB200_E_POSITION (-4)     /* position >= 2^26-1 or row not position-sorted */
#define UB200_E_SAMPLE_ORDER (-5) /* a sample's calls are not strictly position-increasing (the reference's
                                     LOOP 1 merge-scan assumes sorted calls, src/usher_mapper.cpp:191,239) */
#define UB200_E_CAPACITY (-6)     /* best_set capacity too small; required size is in best_set_ptr[n] */
#define UB200_E_NO_DEVICE (-7)    /* no CUDA device / extension not usable */
#define UB200_E_LIMIT (-8)        /* a size exceeds what this build supports (see message) */

/* One tree mutation or one sample call.  Field meaning = MAT::Mutation minus `chrom`
 * (reference src/mutation_annotated_tree.hpp:45-52).  Nucleotides are the reference's codes
 * (src/mutation_annotated_tree.cpp:17-74): one-hot A=1 C=2 G=4 T=8; a sample's mut_nuc may be any
 * non-zero 4-bit IUPAC set. */
typedef struct ub200_mutation {
    int32_t position;   /* 1-based genome coordinate; < 0 = masked (tree rows only) */
    uint8_t ref_nuc;    /* reference allele (one-hot; 0 for masked) */
    uint8_t par_nuc;    /* carried for round-tripping; scoring derives the parental state from the path */
    uint8_t mut_nuc;    /* tree: one-hot allele after the branch; sample: 4-bit allele set */
    uint8_t is_missing; /* sample calls only: 1 = N (Missing_Sample entry with is_missing) */
} ub200_mutation;

/* A mutation-annotated tree flattened by the caller (one DFS).
 * Node i is the i-th node of Tree::depth_first_expansion() (pre-order, children in stored order), so
 * parent[i] < i and parent[0] == -1.  Row i = Node::mutations in stored (position-sorted) order. */
typedef struct ub200_flat_mat {
    uint32_t n_nodes;
    uint64_t n_mutations;
    const int32_t* parent;           /* [n_nodes] DFS index of the parent, -1 for the root */
    const uint64_t* row_ptr;         /* [n_nodes+1] CSR offsets into mutations */
    const ub200_mutation* mutations; /* [n_mutations] */
    const uint32_t* tie_index;       /* [n_nodes] the index `j` the caller would pass as mapper2_input::j,
                                        or NULL for the BFS index of Tree::breadth_first_expansion() that
                                        usher_common uses (src/usher_common.cpp:342,402) */
} ub200_flat_mat;

typedef struct ub200_mat ub200_mat;         /* tree resident on one GPU */
typedef struct ub200_samples ub200_samples; /* a batch of samples resident on the same GPU */

/* Per-sample result = what the reference leaves in best_set_difference / best_node / best_j / num_best /
 * best_node_has_unique after pass 2 (src/usher_common.cpp:416-453). */
typedef struct ub200_placement {
    int32_t score;            /* parsimony score of the best placement */
    uint32_t best_node;       /* DFS index of best_node */
    uint32_t best_j;          /* its tie index (BFS j) */
    uint32_t num_best;        /* number of parsimony-optimal placements */
    uint32_t has_unique;      /* best_node_has_unique: 1 = graft as sibling, 0 = as child */
    uint32_t best_num_leaves; /* get_num_leaves(best_node) used in the tie-break */
    uint32_t reserved[2];
} ub200_placement;

typedef struct ub200_mat_info {
    uint32_t n_nodes;
    uint32_t max_level;          /* depth of the deepest node (root = 0) */
    uint64_t n_mutations;        /* unmasked tree mutations resident on the device */
    uint32_t genome_len;         /* 1 + largest tree position */
    uint32_t n_tiles;            /* DFS segments the scoring kernel schedules */
    uint64_t device_bytes;       /* HBM held by the handle */
    uint64_t algorithmic_bytes;  /* bytes one scoring launch must read from the flattened MAT:
                                    4*n_mutations + 16*n_nodes (DESIGN.md "roofline") */
    int32_t device;
    uint32_t reserved;
} ub200_mat_info;

typedef struct ub200_timing {
    float prep_ms;    /* sample-side table build kernels of the last launch set */
    float score_ms;   /* scoring kernel(s) */
    float reduce_ms;  /* final per-sample reduction kernel(s) */
    uint32_t score_launches;  /* scoring-kernel launches in the last call */
    uint32_t total_launches;  /* all kernels launched by the last call */
    uint64_t score_bytes;     /* algorithmic bytes those scoring launches read */
} ub200_timing;

/* flags for the place calls */
#define UB200_WANT_NODE_SCORES 1u /* fill node_scores (the `-p` output, score+1 on invalid nodes) */
#define UB200_WANT_BEST_SET 2u    /* fill best_set / best_set_ptr (best_j_vec + node_has_unique) */

const char* ub200_last_error(void);
int ub200_abi_version(void);
int ub200_device_count(void);

/* Validate + derive (levels, BFS index, leaf counts, path states) + upload.  One-time per tree version.
 * `device` = CUDA ordinal. */
int ub200_mat_create(const ub200_flat_mat* flat, int device, ub200_mat** out);
void ub200_mat_destroy(ub200_mat* mat);
int ub200_mat_info_get(const ub200_mat* mat, ub200_mat_info* out);
/* Host copies of the derived per-node arrays (any pointer may be NULL). */
int ub200_mat_node_arrays(const ub200_mat* mat, uint32_t* bfs_index, uint32_t* num_leaves, uint32_t* level);
/* Tunables: samples scored per pass (= launch) over the tree (multiple of 32, 32..768; 0 = default 32), and how
 * many groups of 32 samples share one scan of the mutation stream inside a pass (1..3; 0 = chosen from the pass
 * width: the scan is sample-independent, so wide passes share it three ways). */
int ub200_mat_set_pass_samples(ub200_mat* mat, uint32_t samples_per_pass);
int ub200_mat_set_scan_sharing(ub200_mat* mat, uint32_t groups_per_scan);

/* The batched replacement of the search loop.  Host buffers in, host buffers out; H2D/D2H inside.
 *   sample_ptr[n_samples+1], sample_calls[sample_ptr[n_samples]]: CSR of Missing_Sample::mutations,
 *     each sample strictly position-increasing (position, ref_nuc, mut_nuc set, is_missing).
 *   out[n_samples].
 *   node_scores: [n_samples * n_nodes] int32 indexed [s * n_nodes + dfs_index], iff UB200_WANT_NODE_SCORES.
 *   best_set / best_set_ptr[n_samples+1] / best_set_cap, iff UB200_WANT_BEST_SET: optimal nodes of sample s
 *     at best_set[best_set_ptr[s] .. best_set_ptr[s+1]), ascending DFS index, bit 31 = node_has_unique. */
int ub200_place_batch(ub200_mat* mat, uint32_t n_samples, const uint64_t* sample_ptr,
                      const ub200_mutation* sample_calls, uint32_t flags, ub200_placement* out,
                      int32_t* node_scores, uint32_t* best_set, uint64_t* best_set_ptr, uint64_t best_set_cap);

/* Split form of the same call for callers that keep a batch resident (bench "value", multi-launch use). */
int ub200_samples_upload(ub200_mat* mat, uint32_t n_samples, const uint64_t* sample_ptr,
                         const ub200_mutation* sample_calls, ub200_samples** out);
void ub200_samples_free(ub200_samples* s);
/* Score the resident batch; results stay on the device until downloaded.  Asynchronous on the handle's
 * stream unless `sync` != 0. */
int ub200_place_resident(ub200_mat* mat, ub200_samples* s, uint32_t flags, int sync);
int ub200_results_download(ub200_mat* mat, ub200_samples* s, ub200_placement* out);
/* Device pointer to the n_samples ub200_placement records of a resident batch (for a collective on them). */
int ub200_results_device_ptr(ub200_samples* s, void** dev_ptr, size_t* bytes);
/* Device-to-device copy of the n_samples result records into caller-owned device memory, ordered on the
 * handle's stream (e.g. straight into the send buffer of an allgather). */
int ub200_results_copy_device(ub200_mat* mat, ub200_samples* s, void* dst_dev);
int ub200_node_scores_download(ub200_mat* mat, ub200_samples* s, int32_t* node_scores);
int ub200_best_set_download(ub200_mat* mat, ub200_samples* s, uint32_t* best_set, uint64_t* best_set_ptr,
                            uint64_t best_set_cap);

/* Run the handle's work on a caller-provided CUDA stream (cudaStream_t as void*; NULL = library stream). */
int ub200_mat_set_stream(ub200_mat* mat, void* cuda_stream);
int ub200_mat_synchronize(ub200_mat* mat);
/* CUDA-event timings of the last place call on this handle (valid after it completed). */
int ub200_last_timing(ub200_mat* mat, ub200_timing* out);

/* ---- several GPUs of one box (replaces the one tbb::parallel_for over nodes of src/usher_common.cpp:389-414 by
 * "samples sharded over devices"; BASELINE configs 4/5).  The tree is derived once and replicated on `n_devices`
 * devices (0 = every visible device; `devices` = their ordinals or NULL for 0..n-1), one host thread per device.
 * ub200_multi_place_batch has the contract of ub200_place_batch: the batch is cut into contiguous shards of whole
 * 32-sample groups, every device scores its shard against its replica, and each writes its records straight into
 * the caller's host arrays (results are identical to a single-device call: placements do not depend on how samples
 * are grouped).  ub200_multi_mat() exposes replica i for the tunables above. */
typedef struct ub200_multi ub200_multi;
int ub200_multi_create(const ub200_flat_mat* flat, int n_devices, const int* devices, ub200_multi** out);
void ub200_multi_destroy(ub200_multi* multi);
int ub200_multi_size(const ub200_multi* multi);
ub200_mat* ub200_multi_mat(ub200_multi* multi, int i);
int ub200_multi_place_batch(ub200_multi* multi, uint32_t n_samples, const uint64_t* sample_ptr,
                            const ub200_mutation* sample_calls, uint32_t flags, ub200_placement* out,
                            int32_t* node_scores, uint32_t* best_set, uint64_t* best_set_ptr, uint64_t best_set_cap);

/* ---- Fitch-Sankoff per VCF site: the parsimony assignment that builds a MAT from a tree and a VCF (replaces
 * mapper_body::operator(), reference src/usher_mapper.cpp:6-161, driven by the tbb::flow graph of
 * src/mutation_annotated_tree.cpp:2099-2179).  Nodes are given in the order of Tree::breadth_first_expansion()
 * (parent_bfs[0] == -1).  A site = its reference base (0..3 = A,C,G,T) and the genotype overrides of tree nodes:
 * var_node (BFS index) / var_nuc (4-bit allele set, 15 = N); leaves without an override carry the reference allele.
 * Output: one record per (site, node) whose assigned state differs from its parent's (the root's parent state is the
 * reference allele): out_states = parent state << 4 | state, both 0..3.  Records come in no particular order.
 * Returns UB200_E_CAPACITY (with *out_count = the number needed) when out_cap is too small. */
typedef struct ub200_fs_tree ub200_fs_tree;
int ub200_fs_tree_create(uint32_t n_nodes, const int32_t* parent_bfs, int device, ub200_fs_tree** out);
void ub200_fs_tree_destroy(ub200_fs_tree* tree);
int ub200_fs_sites(ub200_fs_tree* tree, uint32_t n_sites, const uint8_t* ref_code, const uint64_t* var_ptr,
                   const uint32_t* var_node, const uint8_t* var_nuc, uint64_t out_cap, uint32_t* out_site,
                   uint32_t* out_node, uint8_t* out_states, uint64_t* out_count);

#ifdef __cplusplus
}
#endif
#endif /* USHER_B200_H */
