// k_fitch_sankoff — the per-site parsimony assignment that builds a MAT from a tree and a VCF
// (reference mapper_body::operator(), src/usher_mapper.cpp:6-161).
//
// The reference runs Sankoff's algorithm with unit costs on int score vectors: bottom-up
//   score[p][j] += min_k(score[c][k] + [k != j])                                   (:86-111)
// and top-down "keep the parent's state unless another base is strictly cheaper, the lowest such base" (:114-156).
// With unit costs min_k(score[c][k] + [k != j]) = m_c + [j not in M0_c], where m_c is the child's minimum and M0_c the set
// of bases that reach it, so only the 4-bit sets travel:  score[p][j] - const = #children whose M0 lacks j, restricted to
// the bases the node itself allows (a leaf: its genotype, default the reference allele; N = all four, :33-63), and
//   M0_p = the allowed bases contained in the most children's sets;  state = parent's state if it is in M0, else the
//   lowest base of M0.
// One CTA per site: nodes in BFS order (a node's children are contiguous), one pass per level bottom-up, one per level
// top-down, a byte of scratch per node and pass.  Sites are independent, so a VCF of S sites keeps S CTAs busy.
#pragma once
#include <cstdint>

namespace ub200 {

struct FsParams {
    uint32_t n_nodes, n_levels, n_sites;
    const uint32_t* level_start;   // [n_levels + 1] BFS ranges of the levels
    const uint32_t* parent;        // [n] BFS index of the parent (root: itself)
    const uint32_t* child_start;   // [n] BFS index of the first child
    const uint32_t* n_children;    // [n]
    const uint8_t* ref_code;       // [sites] reference base 0..3
    const unsigned long long* var_ptr;   // [sites + 1] genotype overrides of the site
    const uint32_t* var_node;      // BFS index
    const uint8_t* var_nuc;        // allowed 4-bit set (15 = N)
    uint8_t* scratch;              // [gridDim.x][2 n]: M0 sets, states
    unsigned long long out_cap;
    uint32_t* out_site; uint32_t* out_node; uint8_t* out_states;   // par << 4 | state
    unsigned long long* out_count;
};

__global__ void __launch_bounds__(1024) k_fitch_sankoff(const FsParams p) {
    uint8_t* m0 = p.scratch + (size_t)blockIdx.x * 2u * p.n_nodes;
    uint8_t* st = m0 + p.n_nodes;
    const uint32_t n = p.n_nodes, T = blockDim.x, t = threadIdx.x;
    for (uint32_t site = blockIdx.x; site < p.n_sites; site += gridDim.x) {
        const uint32_t ref = p.ref_code[site];
        // allowed sets: leaves default to the reference allele, internal nodes to every base; genotypes override
        for (uint32_t i = t; i < n; i += T) m0[i] = p.n_children[i] ? 15u : (uint8_t)(1u << ref);
        __syncthreads();
        for (unsigned long long k = p.var_ptr[site] + t; k < p.var_ptr[site + 1]; k += T) m0[p.var_node[k]] = p.var_nuc[k];
        __syncthreads();
        // bottom-up, deepest level first (a level only reads the level below)
        for (uint32_t lv = p.n_levels; lv-- > 0;) {
            for (uint32_t i = p.level_start[lv] + t; i < p.level_start[lv + 1]; i += T) {
                const uint32_t nc = p.n_children[i];
                if (!nc) continue;
                const uint32_t allowed = m0[i];
                uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                const uint8_t* ch = m0 + p.child_start[i];
                for (uint32_t c = 0; c < nc; c++) {
                    const uint32_t m = ch[c];
                    c0 += m & 1u; c1 += (m >> 1) & 1u; c2 += (m >> 2) & 1u; c3 += (m >> 3) & 1u;
                }
                uint32_t best = 0;
                if (allowed & 1u) best = max(best, c0 + 1u);
                if (allowed & 2u) best = max(best, c1 + 1u);
                if (allowed & 4u) best = max(best, c2 + 1u);
                if (allowed & 8u) best = max(best, c3 + 1u);
                uint32_t s = 0;
                if ((allowed & 1u) && c0 + 1u == best) s |= 1u;
                if ((allowed & 2u) && c1 + 1u == best) s |= 2u;
                if ((allowed & 4u) && c2 + 1u == best) s |= 4u;
                if ((allowed & 8u) && c3 + 1u == best) s |= 8u;
                m0[i] = (uint8_t)s;
            }
            __syncthreads();
        }
        // top-down: the root's "parent state" is the reference allele (:123-125)
        for (uint32_t lv = 0; lv < p.n_levels; lv++) {
            for (uint32_t i = p.level_start[lv] + t; i < p.level_start[lv + 1]; i += T) {
                const uint32_t par = lv ? st[p.parent[i]] : ref;
                const uint32_t m = m0[i];
                const uint32_t s = ((m >> par) & 1u) ? par : (uint32_t)(__ffs((int)m) - 1);
                st[i] = (uint8_t)s;
                // one atomic per warp for the records of its lanes
                const unsigned am = __activemask();
                const unsigned em = __ballot_sync(am, s != par);
                if (s != par) {
                    const unsigned lane = threadIdx.x & 31u, leader = (unsigned)__ffs((int)em) - 1u;
                    unsigned long long o = 0;
                    if (lane == leader) o = atomicAdd(p.out_count, (unsigned long long)__popc(em));
                    o = __shfl_sync(em, o, (int)leader) + (unsigned long long)__popc(em & ((1u << lane) - 1u));
                    if (o < p.out_cap) { p.out_site[o] = site; p.out_node[o] = i; p.out_states[o] = (uint8_t)((par << 4) | s); }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace ub200
