// Internal layout shared by the host-side derivation (derive.cpp), the kernels (score_kernel.cu) and the
// C-ABI glue (api.cu).  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "usher_b200.h"

#ifdef __CUDACC__
#define UB200_HD __host__ __device__
#else
#define UB200_HD
#endif

namespace ub200 {

// ---- packed tree mutation: pos:26 | ref:2 | prev:2 | mut:2 (nucleotide codes = log2 of the one-hot) ----
constexpr uint32_t kPosBits = 26;
constexpr uint32_t kMaxPos = (1u << kPosBits) - 2;  // positions are < 2^26-1
UB200_HD inline uint32_t pack_mut(uint32_t pos, uint32_t refc, uint32_t prevc, uint32_t mutc) {
    return (pos << 6) | (refc << 4) | (prevc << 2) | mutc;
}

// ---- per-node header (16 B) ----
//   x = G        : int32  Dref(parent) - A0(node)      (root: Dref(root))
//   y = tiekey   : uint32 N-1-rank of (num_leaves, tie_index): smaller = preferred on equal score
//   z = level:18 | plane:6 | flags:8   (plane = 1 + (parent & 31) when the parent lies in the node's own
//                                       aligned 32-node block, else 0)
//   w = nmut<<16 | c0   (unmasked row length, number of row mutations that are "common" with a sample
//                        lacking every position of the row; both < 65535)
constexpr uint32_t kFlagLeaf = 1u;
constexpr uint32_t kFlagMasked = 2u;   // row holds a masked mutation: LOOP 1 takes nothing (usher_mapper.cpp:197-200)
constexpr uint32_t kFlagRoot = 4u;
constexpr uint32_t kFlagValid0 = 8u;   // validity predicate for a sample that hits no position of the row
constexpr uint32_t kFlagHu0 = 16u;     // has_unique for such a sample
constexpr uint32_t kMaxRow = 65534;
constexpr uint32_t kLevelShift = 14;
constexpr uint32_t kMaxLevel = (1u << 18) - 1;
UB200_HD inline uint32_t hdr_level(uint32_t z) { return z >> kLevelShift; }
UB200_HD inline uint32_t hdr_plane(uint32_t z) { return (z >> 8) & 63u; }
UB200_HD inline uint32_t hdr_flags(uint32_t z) { return z & 255u; }

struct NodeHdr {
    int32_t g;
    uint32_t tiekey;
    uint32_t level_flags;
    uint32_t nmut_c0;
};
static_assert(sizeof(NodeHdr) == 16, "header must be 16 bytes");

// ring geometry of the scoring kernel (absolute-aligned chunks)
constexpr uint32_t kMutChunk = 256;   // mutation words per bulk copy (1 KB)
constexpr uint32_t kHdrChunk = 32;    // headers per bulk copy (512 B)

// ---- k_score3 layout (score_kernel3.cuh, DESIGN.md "Data layout") ----
// stream word, wide form  : pos>>5 :18 | lane:5 | prev:2 | mut:2 | pos&31 :5
//              narrow form: pos>>5 :16 | 00 | lane:5 | prev:2 | mut:2 | pos&31 :5      (genomes < 2^21 positions)
// lane = node & 31 in a block segment, level & 31 in a seed segment; the reference allele comes from the sample
// table row.  The split position gives the bit index for free (funnel shifts wrap at 32) and, in the narrow
// form, the byte offset of the bitmap word with one shift (w >> 14).  Pad words carry pos = L (never called).
constexpr uint32_t kMaxPos3 = (1u << 23) - 2;
constexpr uint32_t kMaxPos3Narrow = (1u << 21) - 2;
UB200_HD inline uint32_t pack_mut3(bool narrow, uint32_t pos, uint32_t lane, uint32_t prevc, uint32_t mutc) {
    return ((pos >> 5) << (narrow ? 16 : 14)) | (lane << 9) | (prevc << 7) | (mutc << 5) | (pos & 31u);
}
template <bool NARROW>
UB200_HD inline uint32_t mut3_pos(uint32_t w) { return ((w >> (NARROW ? 16 : 14)) << 5) | (w & 31u); }
// header of the k_score3 layout: x = G, y = bitmask of the node's ancestors inside its own aligned 32-node
// block, z = level:18 | flags:14, w = nmut<<16 | c0
constexpr uint32_t kFlagOpen = 32u;    // internal node with a descendant beyond its 32-node block
constexpr uint32_t kChunk3 = 256;      // stream words per bulk copy (1 KB); tiles start on chunk boundaries

// std::vector that leaves new elements uninitialised (resize() of the 1.3 GB stream must not write zeros first)
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
template <class T> using RawVec = std::vector<T, NoInitAlloc<T>>;
using StreamVec = RawVec<uint32_t>;

struct Derived {
    uint32_t n = 0;
    uint64_t m = 0;          // unmasked mutations kept on the device
    uint32_t L = 0;          // 1 + largest tree position
    uint32_t max_level = 0;
    uint32_t max_row = 0;    // longest unmasked mutation row
    std::vector<uint32_t> level, tie_index, num_leaves, tiekey, key_to_node;
    std::vector<uint32_t> row32;      // [n+1] offsets into mutw
    StreamVec mutw;                   // padded to kMutChunk
    RawVec<NodeHdr> hdr;              // padded to kHdrChunk
    std::vector<uint8_t> ref_of;      // [L] one-hot reference allele where the tree mutates, else 0
    // tiles = contiguous DFS ranges; anc = root..parent chain of each tile's first node
    std::vector<uint32_t> tile_start; // [T+1]
    std::vector<uint32_t> anc_ptr;    // [T+1]
    std::vector<uint32_t> anc;        // node ids, root first
    int32_t root_init_extra = 0;      // |root row| (incl. masked) for the reference's initial bound
    // ---- k_score3 layout: per tile one contiguous piece of `stream` = [seed segments][block segments];
    // a seed segment holds the rows of up to 32 consecutive levels of the root path of the tile's first node,
    // a block segment the rows of one aligned 32-node block; every segment starts on a 4-word boundary.
    bool have3 = false;               // false: genome too long for the 23-bit position field
    bool narrow3 = false;             // stream words in the narrow form
    StreamVec stream;                 // padded to kChunk3
    RawVec<NodeHdr> hdr3;             // padded like hdr
    std::vector<uint32_t> tile3_start;// [T3+1] first node of each tile (multiple of 32)
    std::vector<uint32_t> tile3_w0;   // [T3+1] first stream chunk of each tile (tile t ends where t+1 starts)
    std::vector<uint32_t> tile3_lvl;  // [T3]   level of the tile's first node = number of seeded levels
    std::vector<uint32_t> tile3_sseg; // [T3+1] offsets into seed_end
    std::vector<uint32_t> seed_end;   // end of each seed segment, in 4-word units
    std::vector<uint32_t> blk_words;  // [blocks] stream words of each block segment (row lengths, rounded up to 4)
    // per-block record the k_score4 consumer reads instead of the 32 headers (score_kernel4.cuh):
    //   x = min over the block's nodes of G - nmut (the block term of the exact lower bound)
    //   y = open mask: nodes with descendants beyond the block (the chain whose stack rows later blocks read)
    //   z = level of the first open node,  w = words of the block segment (= blk_words)
    std::vector<uint32_t> blk_rec;    // [blocks][4]
    uint64_t seed_words = 0;          // stream words spent on seed segments
};

// Validate the caller's flat tree and derive everything the kernels need.  Returns UB200_* status and
// fills err on failure.
// min_tile_cost: floor of a tile's cost (mutations + 4 per node); 0 = default (tests pass small values to
// cut small trees into many tiles).
int derive(const ub200_flat_mat& flat, uint32_t target_tiles, Derived& out, std::string& err,
           uint32_t min_tile_cost = 0);

inline int nuc_code(uint8_t one_hot) {
    switch (one_hot) {
        case 1: return 0;
        case 2: return 1;
        case 4: return 2;
        case 8: return 3;
        default: return -1;
    }
}

}  // namespace ub200
