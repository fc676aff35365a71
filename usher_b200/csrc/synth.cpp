// Seeded synthetic MAT / sample generator for the benchmark configs of BASELINE.json:
//   G(N, mu, L, shape, seed): RNG std::mt19937_64(seed); ref[p] = 1 << (rng()%4);
//   topology `uniform`: node i (creation order) attaches to rng()%i; `sc2`: with prob 0.6 to
//   i-1-rng()%min(i,64) (recent-node attachment, long backbones) else rng()%i;
//   mutations: one DFS with a live state array, each non-root node draws k~Poisson(mu) positions in [1,L],
//   dedupes, par = state[p], mut = uniform among the other three bases (reversions arise naturally).
// Output is the flat DFS-pre-order form of include/usher_b200.h.  Bench/test tooling, no device code.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "usher_b200.h"
#include "usher_b200_synth.h"

struct ub200_synth {
    uint32_t n = 0, L = 0;
    std::vector<int32_t> parent;      // DFS order
    std::vector<uint64_t> row_ptr;
    std::vector<ub200_mutation> muts;
    std::vector<uint8_t> ref;         // [L+1]
    // last generated sample batch
    std::vector<uint64_t> s_ptr;
    std::vector<ub200_mutation> s_calls;
    std::vector<uint32_t> s_origin;   // node each sample was derived from
};

namespace {

struct Site { int32_t pos; uint8_t nuc; };

// non-reference sites of the genotype at node v (most recent mutation per position on the root path)
void genotype(const ub200_synth* g, uint32_t v, std::vector<Site>& out, std::vector<int32_t>& seen_scratch) {
    out.clear();
    std::vector<Site> all;
    for (int32_t n = (int32_t)v; n >= 0; n = g->parent[n]) {
        for (uint64_t k = g->row_ptr[n]; k < g->row_ptr[n + 1]; k++) {
            const auto& m = g->muts[k];
            if (m.position < 0) continue;
            all.push_back({m.position, m.mut_nuc});  // child-first order: first seen wins
        }
    }
    std::stable_sort(all.begin(), all.end(), [](const Site& a, const Site& b) { return a.pos < b.pos; });
    for (size_t i = 0; i < all.size(); i++) {
        if (i && all[i].pos == all[i - 1].pos) continue;
        if (all[i].nuc != g->ref[all[i].pos]) out.push_back(all[i]);
    }
    (void)seen_scratch;
}

uint8_t other_base(std::mt19937_64& rng, uint8_t one_hot) {
    int cur = __builtin_ctz(one_hot);
    int k = (int)(rng() % 3);
    int b = (cur + 1 + k) % 4;
    return (uint8_t)(1u << b);
}

}  // namespace

extern "C" {

int ub200_synth_mat_create(uint32_t n_nodes, double mu, uint32_t genome_len, int shape, uint64_t seed,
                           ub200_synth** out) {
    if (!out || n_nodes == 0 || genome_len == 0) return UB200_E_ARG;
    auto* g = new ub200_synth();
    g->n = n_nodes;
    g->L = genome_len;
    std::mt19937_64 rng(seed);
    g->ref.assign(genome_len + 1, 1);
    for (uint32_t p = 1; p <= genome_len; p++) g->ref[p] = (uint8_t)(1u << (rng() % 4));
    // topology in creation order
    std::vector<uint32_t> cpar(n_nodes, 0);
    for (uint32_t i = 1; i < n_nodes; i++) {
        if (shape == UB200_SYNTH_SC2 && (rng() % 10) < 6) {
            uint32_t w = std::min<uint32_t>(i, 64);
            cpar[i] = i - 1 - (uint32_t)(rng() % w);
        } else {
            cpar[i] = (uint32_t)(rng() % i);
        }
    }
    // children CSR (creation order), then DFS pre-order relabel
    std::vector<uint32_t> cnt(n_nodes + 1, 0);
    for (uint32_t i = 1; i < n_nodes; i++) cnt[cpar[i] + 1]++;
    for (uint32_t i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
    std::vector<uint32_t> kids(n_nodes ? n_nodes - 1 : 0), fill(cnt.begin(), cnt.end() - 1);
    for (uint32_t i = 1; i < n_nodes; i++) kids[fill[cpar[i]]++] = i;
    std::vector<uint32_t> order;  // order[dfs] = creation id
    order.reserve(n_nodes);
    std::vector<uint32_t> newid(n_nodes);
    {
        std::vector<uint32_t> st;
        st.push_back(0);
        while (!st.empty()) {
            uint32_t u = st.back();
            st.pop_back();
            newid[u] = (uint32_t)order.size();
            order.push_back(u);
            for (uint32_t k = cnt[u + 1]; k-- > cnt[u];) st.push_back(kids[k]);
        }
    }
    g->parent.resize(n_nodes);
    for (uint32_t d = 0; d < n_nodes; d++) g->parent[d] = d == 0 ? -1 : (int32_t)newid[cpar[order[d]]];
    // mutations: DFS order with live state; undo log per depth
    g->row_ptr.assign((size_t)n_nodes + 1, 0);
    g->muts.reserve((size_t)(mu * n_nodes * 1.02) + 16);
    std::vector<uint8_t> state(g->ref);
    std::poisson_distribution<int> pois(mu > 0 ? mu : 1e-9);
    struct Undo { uint32_t node; int32_t pos; uint8_t old; };
    std::vector<Undo> undo;
    std::vector<uint32_t> path;  // current root path (dfs ids)
    std::vector<int32_t> ps;
    for (uint32_t d = 0; d < n_nodes; d++) {
        // pop to parent
        while (!path.empty() && (int32_t)path.back() != g->parent[d]) {
            uint32_t top = path.back();
            path.pop_back();
            while (!undo.empty() && undo.back().node == top) {
                state[undo.back().pos] = undo.back().old;
                undo.pop_back();
            }
        }
        g->row_ptr[d] = g->muts.size();
        if (d != 0 && mu > 0) {
            int k = pois(rng);
            ps.clear();
            for (int q = 0; q < k; q++) ps.push_back(1 + (int32_t)(rng() % genome_len));
            std::sort(ps.begin(), ps.end());
            ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
            for (int32_t p : ps) {
                ub200_mutation m;
                m.position = p;
                m.ref_nuc = g->ref[p];
                m.par_nuc = state[p];
                m.mut_nuc = other_base(rng, state[p]);
                m.is_missing = 0;
                undo.push_back({d, p, state[p]});
                state[p] = m.mut_nuc;
                g->muts.push_back(m);
            }
        }
        path.push_back(d);
    }
    g->row_ptr[n_nodes] = g->muts.size();
    *out = g;
    return UB200_OK;
}

void ub200_synth_free(ub200_synth* g) { delete g; }

int ub200_synth_flat(ub200_synth* g, ub200_flat_mat* out) {
    if (!g || !out) return UB200_E_ARG;
    out->n_nodes = g->n;
    out->n_mutations = g->muts.size();
    out->parent = g->parent.data();
    out->row_ptr = g->row_ptr.data();
    out->mutations = g->muts.data();
    out->tie_index = nullptr;
    return UB200_OK;
}

const uint8_t* ub200_synth_reference(ub200_synth* g) { return g ? g->ref.data() : nullptr; }

int ub200_synth_samples(ub200_synth* g, uint32_t n_samples, int family, uint64_t seed, const uint64_t** sample_ptr,
                        const ub200_mutation** calls, const uint32_t** origin) {
    if (!g || !sample_ptr || !calls) return UB200_E_ARG;
    std::mt19937_64 rng(seed);
    g->s_ptr.assign(1, 0);
    g->s_calls.clear();
    g->s_origin.clear();
    std::vector<Site> gv, keep;
    std::vector<int32_t> scratch;
    std::vector<ub200_mutation> cur;
    for (uint32_t s = 0; s < n_samples; s++) {
        uint32_t v = (uint32_t)(rng() % g->n);
        g->s_origin.push_back(v);
        genotype(g, v, gv, scratch);
        keep.clear();
        const bool snv40 = (family == UB200_FAMILY_SNV40 || family == UB200_FAMILY_AMBIG);
        if (snv40) {
            // random <=36-subset of G_v, then private SNVs until 40 calls
            std::vector<Site> sh(gv);
            std::shuffle(sh.begin(), sh.end(), rng);
            if (sh.size() > 36) sh.resize(36);
            keep = sh;
        } else {
            // leaf-derived: G_v minus <=2 sites (family LEAF) ...
            keep = gv;
            uint32_t drop = (uint32_t)(rng() % 3);
            for (uint32_t q = 0; q < drop && !keep.empty(); q++) keep.erase(keep.begin() + (rng() % keep.size()));
        }
        std::sort(keep.begin(), keep.end(), [](const Site& a, const Site& b) { return a.pos < b.pos; });
        // private SNVs at unused positions
        size_t target = snv40 ? 40 : keep.size() + (size_t)(rng() % 6);
        while (keep.size() < target) {
            int32_t p = 1 + (int32_t)(rng() % g->L);
            auto it = std::lower_bound(keep.begin(), keep.end(), p, [](const Site& a, int32_t b) { return a.pos < b; });
            if (it != keep.end() && it->pos == p) continue;
            keep.insert(it, Site{p, other_base(rng, g->ref[p])});
        }
        cur.clear();
        for (auto& st : keep) {
            ub200_mutation m;
            m.position = st.pos;
            m.ref_nuc = g->ref[st.pos];
            m.par_nuc = m.ref_nuc;
            m.mut_nuc = st.nuc;
            m.is_missing = 0;
            cur.push_back(m);
        }
        if (family == UB200_FAMILY_AMBIG) {
            // 10% of SNV calls widened to 2-3 bit IUPAC sets (may include ref)
            for (auto& m : cur) {
                if (rng() % 10 == 0) {
                    int extra = 1 + (int)(rng() % 2);
                    for (int q = 0; q < extra; q++) m.mut_nuc |= (uint8_t)(1u << (rng() % 4));
                }
            }
            // 1-4 N-runs of length Geometric(mean 200); the first is anchored on a mutated path position
            int runs = 1 + (int)(rng() % 4);
            std::geometric_distribution<int> geo(1.0 / 200.0);
            for (int r = 0; r < runs; r++) {
                int len = 1 + geo(rng);
                int32_t start;
                if (r == 0 && !gv.empty()) {
                    int32_t anchor = gv[rng() % gv.size()].pos;
                    start = std::max<int32_t>(1, anchor - (int32_t)(rng() % len));
                } else {
                    start = 1 + (int32_t)(rng() % g->L);
                }
                int32_t end = std::min<int64_t>((int64_t)start + len - 1, g->L);
                for (int32_t p = start; p <= end; p++) {
                    auto it = std::lower_bound(cur.begin(), cur.end(), p,
                                               [](const ub200_mutation& a, int32_t b) { return a.position < b; });
                    ub200_mutation m;
                    m.position = p;
                    m.ref_nuc = g->ref[p];
                    m.par_nuc = m.ref_nuc;
                    m.mut_nuc = 15;
                    m.is_missing = 1;
                    if (it != cur.end() && it->position == p) *it = m;
                    else cur.insert(it, m);
                }
            }
        }
        g->s_calls.insert(g->s_calls.end(), cur.begin(), cur.end());
        g->s_ptr.push_back(g->s_calls.size());
    }
    *sample_ptr = g->s_ptr.data();
    *calls = g->s_calls.data();
    if (origin) *origin = g->s_origin.data();
    return UB200_OK;
}

}  // extern "C"
