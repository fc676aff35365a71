// Host-side derivation: caller's flat MAT (include/usher_b200.h) -> the SoA the scoring kernel streams.
//
// What is precomputed here is exactly the sample-INDEPENDENT part of mapper2_body
// (reference src/usher_mapper.cpp:167-504), i.e. its result for a sample that carries no call at any
// position of the node's root path:
//   * prev(m): the true path state just above the branch at m's position (what the reference finds by
//     walking parents and keeping the most recent mutation per position, :275-286);
//   * Dref(n): number of path positions whose state differs from the reference allele = what LOOP 3
//     (:393-445) counts for such a sample;
//   * A0(n), c0(n): LOOP 1 (:190-264) outcome for such a sample: a branch mutation is "common" iff it
//     mutates back to the reference allele (:244-259); A0 = how many of those undo a non-reference state;
//   * valid0(n): the placement-validity predicate (:454-455) for such a sample;
//   * tiekey(n): total order of the tie-break (:483-486): more leaves first, then larger index j;
//   * num_leaves (mutation_annotated_tree.cpp:866-879), BFS index (:1225-1251), level.
// The kernel then only has to correct these for the few mutations that hit a position the sample calls.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <chrono>
#include <cstdio>
#include <thread>

#include "ub200_internal.h"

namespace ub200 {

namespace {
struct PhaseTimer {
    bool on = getenv("UB200_DERIVE_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[derive] %-28s %7.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// Host threads for the derivation: contiguous chunks of [0, n), one chunk per thread (results are independent of the
// thread count: every chunk writes its own slice, and the reductions are order-free min / max / sum / first error).
unsigned host_threads(uint64_t work) {
    unsigned t = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
    if (const char* e = getenv("UB200_HOST_THREADS")) return (unsigned)std::max(1, atoi(e));   // tests force the split
    return work < (1u << 18) ? 1u : t;
}
template <class F>
void parallel_chunks(uint32_t n, unsigned nthr, F body) {   // body(chunk, lo, hi)
    if (nthr <= 1) { body(0u, 0u, n); return; }
    std::vector<std::thread> pool;
    for (unsigned c = 0; c < nthr; c++) {
        const uint32_t lo = (uint32_t)((uint64_t)n * c / nthr), hi = (uint32_t)((uint64_t)n * (c + 1) / nthr);
        pool.emplace_back([=, &body]() { body(c, lo, hi); });
    }
    for (auto& th : pool) th.join();
}
}  // namespace

// The k_score3 layout (ub200_internal.h), built from the arrays derive() has already filled.
static void derive3(const ub200_flat_mat& f, uint32_t target_tiles, uint32_t min_tile_cost, Derived& d) {
    const uint32_t n = d.n;
    d.have3 = d.L <= kMaxPos3;
    if (!d.have3) return;
    d.narrow3 = d.L <= kMaxPos3Narrow && !getenv("UB200_WIDE_WORDS");   // env: test hook for the wide form
    const uint32_t nblk = (n + 31) / 32;
    PhaseTimer pt;
    // subtree ends (DFS pre-order: subtree of i = [i, send[i])), words on the root path above each node
    std::vector<uint32_t> send(n);
    for (uint32_t i = 0; i < n; i++) send[i] = i + 1;
    for (uint32_t i = n; i-- > 1;) send[f.parent[i]] = std::max(send[f.parent[i]], send[i]);
    std::vector<uint64_t> pathw(n, 0);
    for (uint32_t i = 1; i < n; i++) {
        const uint32_t p = (uint32_t)f.parent[i];
        pathw[i] = pathw[p] + (d.row32[p + 1] - d.row32[p]);
    }
    // headers: y = in-block ancestor mask, flags + open
    d.hdr3 = d.hdr;
    {
        std::vector<uint32_t> am(n, 0);
        for (uint32_t i = 0; i < n; i++) {
            if (i && (uint32_t)f.parent[i] >= (i & ~31u)) am[i] = am[f.parent[i]] | (1u << (f.parent[i] & 31));
            NodeHdr& h = d.hdr3[i];
            const uint32_t flags = hdr_flags(h.level_flags);
            const bool leaf = flags & kFlagLeaf;
            const bool open = !leaf && send[i] > (i | 31u) + 1u;
            h.tiekey = am[i];
            h.level_flags = (d.level[i] << kLevelShift) | flags | (open ? kFlagOpen : 0u);
        }
    }
    pt.lap("  3: subtree ends, headers");
    // tiles: whole blocks, roughly equal cost; the seed stream must stay a small fraction of the tree
    const uint64_t node_cost = 4;
    const uint64_t total = d.m + node_cost * n;
    uint64_t per = total / (target_tiles ? target_tiles : 1);
    per = std::min<uint64_t>(std::max<uint64_t>(per, min_tile_cost ? min_tile_cost : 4096), 1u << 16);
    for (;;) {
        d.tile3_start.assign(1, 0);
        uint64_t acc = 0, seed = 0, seen = 0;
        for (uint32_t b = 0; b < nblk; b++) {
            const uint32_t e = std::min(n, b * 32 + 32);
            const uint64_t cost = (d.row32[e] - d.row32[b * 32]) + node_cost * (e - b * 32);
            acc += cost;
            seen += cost;
            // the last tiles handed out are smaller, so that the workers finish closer together
            const uint64_t lim = seen * 10 > total * 9 ? per / 4 : (seen * 10 > total * 7 ? per / 2 : per);
            if (acc >= std::max<uint64_t>(lim, min_tile_cost ? min_tile_cost : 2048) && e < n) {
                d.tile3_start.push_back(e);
                seed += pathw[e];
                acc = 0;
            }
        }
        d.tile3_start.push_back(n);
        if (seed <= std::max<uint64_t>(d.m / 8, 1u << 16) || d.tile3_start.size() <= 2) break;
        per += per / 2;
    }
    pt.lap("  3: tile boundaries");
    const size_t T = d.tile3_start.size() - 1;
    d.tile3_w0.assign(T + 1, 0);
    d.tile3_lvl.assign(T, 0);
    d.tile3_sseg.assign(T + 1, 0);
    d.seed_end.clear();
    d.blk_words.assign(nblk, 0);
    const bool nw = d.narrow3;
    const uint32_t pad = pack_mut3(nw, d.L, 0, 0, 0);
    auto conv = [&](uint32_t w, uint32_t lane) { return pack_mut3(nw, w >> 6, lane, (w >> 2) & 3u, w & 3u); };
    const uint32_t pshift = nw ? 16 : 14;
    const bool pad_steps = !getenv("UB200_NO_STEP_PAD");   // developer switch: measure what the padding buys
    // Every tile's piece of the stream starts on a 1 KB boundary, so the pieces are built independently (one
    // host thread per slice of tiles) and concatenated afterwards.
    struct Piece { std::vector<uint32_t> words; std::vector<uint32_t> seed_end4; uint64_t seed_words = 0; };
    std::vector<Piece> pieces(T);
    auto build_tile = [&](size_t t, std::vector<uint32_t>& chain, std::vector<uint32_t>& seg) {
        Piece& pc = pieces[t];
        std::vector<uint32_t>& out = pc.words;
        // A segment's words go out sorted by position and transposed inside every 128-word row piece: the
        // scanner reads a row with one 16-byte load per lane, so component j of the 32 lanes should hold 32
        // CONSECUTIVE sorted words -> their bitmap words are consecutive too and the 32 bitmap reads fall into
        // distinct shared-memory banks (a random order costs ~3.5 wavefronts per read).  Which node (or level) a
        // word belongs to is in the word, so the order inside a segment is free.
        auto emit_segment = [&]() {
            std::sort(seg.begin(), seg.end(), [&](uint32_t x, uint32_t y) {
                const uint32_t px = ((x >> pshift) << 5) | (x & 31u), py = ((y >> pshift) << 5) | (y & 31u);
                return px != py ? px < py : x < y;
            });
            while (seg.size() % 4) seg.push_back(pad);
            // The kernel takes the tile's piece in steps of 512 words (4 rows) counted from the tile's start and
            // hands the hits of a step out segment by segment.  A segment that starts on a step boundary is padded
            // up to the next one when that costs at most an eighth of its length: it then spans the fewest possible
            // steps, every step it touches belongs to it alone, and the segments behind it stay aligned.
            if (pad_steps && out.size() % 512 == 0) {
                const size_t padw = (512 - seg.size() % 512) % 512;
                if (padw && padw * 8 <= seg.size()) seg.resize(seg.size() + padw, pad);
            }
            size_t s0 = 0;
            while (s0 < seg.size()) {
                const size_t off = out.size();
                const size_t m = std::min<size_t>(128 - off % 128, seg.size() - s0), nl = m / 4;
                out.resize(off + m);
                for (size_t i = 0; i < nl; i++)
                    for (size_t j = 0; j < 4; j++) out[off + 4 * i + j] = seg[s0 + j * nl + i];
                s0 += m;
            }
            seg.clear();
        };
        const uint32_t n0 = d.tile3_start[t], n1 = d.tile3_start[t + 1];
        const uint32_t lvl0 = d.level[n0];
        d.tile3_lvl[t] = lvl0;
        chain.assign(lvl0, 0);
        for (int32_t a = f.parent[n0]; a >= 0; a = f.parent[a]) chain[d.level[a]] = (uint32_t)a;
        out.reserve((size_t)(d.row32[n1] - d.row32[n0]) + (size_t)(pathw[n0]) + (n1 - n0) * 3 + 1024);
        for (uint32_t l0 = 0; l0 < lvl0; l0 += 32) {
            for (uint32_t l = l0; l < std::min(lvl0, l0 + 32); l++)
                for (uint32_t k = d.row32[chain[l]]; k < d.row32[chain[l] + 1]; k++)
                    seg.push_back(conv(d.mutw[k], l & 31u));
            emit_segment();
            pc.seed_end4.push_back((uint32_t)(out.size() / 4));
        }
        pc.seed_words = out.size();
        for (uint32_t b = n0; b < n1; b += 32) {
            const uint32_t e = std::min(n1, b + 32);
            const size_t s0 = out.size();
            for (uint32_t i = b; i < e; i++)
                for (uint32_t k = d.row32[i]; k < d.row32[i + 1]; k++) seg.push_back(conv(d.mutw[k], i & 31u));
            emit_segment();
            d.blk_words[b >> 5] = (uint32_t)(out.size() - s0);
        }
        while (out.size() % kChunk3) out.push_back(pad);
    };
    {
        unsigned nthr = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (d.m < (1u << 20)) nthr = 1;
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            std::vector<uint32_t> chain, seg;
            for (size_t t = next.fetch_add(1); t < T; t = next.fetch_add(1)) build_tile(t, chain, seg);
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < nthr; i++) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
    }
    pt.lap("  3: tile pieces (threads)");
    uint64_t total_words = 0;
    for (size_t t = 0; t < T; t++) total_words += pieces[t].words.size();
    d.stream.resize(total_words);
    d.seed_words = 0;
    uint64_t w = 0;
    std::vector<uint64_t> piece_off(T + 1, 0);
    for (size_t t = 0; t < T; t++) {
        Piece& pc = pieces[t];
        piece_off[t] = w;
        d.tile3_w0[t] = (uint32_t)(w / kChunk3);
        for (uint32_t e4 : pc.seed_end4) d.seed_end.push_back((uint32_t)(w / 4) + e4);
        d.tile3_sseg[t + 1] = (uint32_t)d.seed_end.size();
        d.seed_words += pc.seed_words;
        w += pc.words.size();
    }
    parallel_chunks((uint32_t)T, host_threads(total_words), [&](unsigned, uint32_t lo, uint32_t hi) {
        for (uint32_t t = lo; t < hi; t++) {
            Piece& pc = pieces[t];
            std::memcpy(d.stream.data() + piece_off[t], pc.words.data(), pc.words.size() * sizeof(uint32_t));
            std::vector<uint32_t>().swap(pc.words);
        }
    });
    d.tile3_w0[T] = (uint32_t)(w / kChunk3);
    d.seed_end.push_back(0);   // never empty
    pt.lap("  3: concatenate");
    // ---- block records: what the consumer needs of a block that cannot hold an optimum (score_kernel4.cuh)
    d.blk_rec.assign((size_t)nblk * 4, 0);
    for (uint32_t b = 0; b < n; b += 32) {
        const uint32_t e = std::min(n, b + 32);
        int32_t omin = INT32_MAX;
        uint32_t open = 0, lv0 = 0;
        for (uint32_t i = b; i < e; i++) {
            omin = std::min(omin, d.hdr3[i].g - (int32_t)(d.hdr3[i].nmut_c0 >> 16));
            if (d.hdr3[i].level_flags & kFlagOpen) {
                if (!open) lv0 = d.level[i];
                open |= 1u << (i & 31);
            }
        }
        uint32_t* r = &d.blk_rec[(size_t)(b >> 5) * 4];
        r[0] = (uint32_t)omin; r[1] = open; r[2] = lv0; r[3] = d.blk_words[b >> 5];
    }
}

int derive(const ub200_flat_mat& f, uint32_t target_tiles, Derived& d, std::string& err, uint32_t min_tile_cost) {
    const uint32_t n = f.n_nodes;
    if (n == 0 || !f.parent || !f.row_ptr || (f.n_mutations && !f.mutations)) {
        err = "flat MAT: empty tree or NULL array";
        return UB200_E_ARG;
    }
    if (n >= (1u << 31)) { err = "flat MAT: more than 2^31-1 nodes"; return UB200_E_LIMIT; }
    if (f.row_ptr[0] != 0 || f.row_ptr[n] != f.n_mutations) {
        err = "flat MAT: row_ptr does not span [0, n_mutations]";
        return UB200_E_ARG;
    }
    PhaseTimer pt;
    d.n = n;
    // ---- topology checks: DFS pre-order <=> parent[i] lies on the root path of node i-1
    d.level.assign(n, 0);
    if (f.parent[0] != -1) { err = "flat MAT: node 0 must be the root (parent -1)"; return UB200_E_TREE_ORDER; }
    {
        std::vector<uint32_t> path;
        path.push_back(0);
        for (uint32_t i = 1; i < n; i++) {
            int32_t p = f.parent[i];
            if (p < 0 || (uint32_t)p >= i) {
                err = "flat MAT: parent[" + std::to_string(i) + "] is not an earlier node";
                return UB200_E_TREE_ORDER;
            }
            while (!path.empty() && path.back() != (uint32_t)p) path.pop_back();
            if (path.empty()) {
                err = "flat MAT: nodes are not in DFS pre-order at node " + std::to_string(i);
                return UB200_E_TREE_ORDER;
            }
            d.level[i] = d.level[p] + 1;
            if (d.level[i] > kMaxLevel) { err = "flat MAT: tree deeper than 2^18-1"; return UB200_E_LIMIT; }
            path.push_back(i);
        }
    }
    d.max_level = *std::max_element(d.level.begin(), d.level.end());
    pt.lap("topology checks");
    // ---- leaves, leaf counts (reverse sweep), BFS index
    std::vector<uint32_t> nchild(n + 1, 0);
    for (uint32_t i = 1; i < n; i++) nchild[f.parent[i] + 1]++;
    d.num_leaves.assign(n, 0);
    for (uint32_t i = n; i-- > 0;) {
        if (nchild[i + 1] == 0) d.num_leaves[i] = 1;
        if (i) d.num_leaves[f.parent[i]] += d.num_leaves[i];
    }
    d.tie_index.resize(n);
    if (f.tie_index) {
        std::memcpy(d.tie_index.data(), f.tie_index, sizeof(uint32_t) * n);
    } else {
        std::vector<uint32_t> off(nchild);
        for (uint32_t i = 0; i < n; i++) off[i + 1] += off[i];
        std::vector<uint32_t> kids(n > 1 ? n - 1 : 1), fill(off.begin(), off.end() - 1);
        for (uint32_t i = 1; i < n; i++) kids[fill[f.parent[i]]++] = i;
        std::vector<uint32_t> q(n);
        uint32_t head = 0, tail = 0;
        q[tail++] = 0;
        while (head < tail) {
            uint32_t u = q[head];
            d.tie_index[u] = head++;
            for (uint32_t k = off[u]; k < off[u + 1]; k++) q[tail++] = kids[k];
        }
    }
    pt.lap("leaves + BFS index");
    // ---- tie-break order: preferred = more leaves, then larger j  -> tiekey 0 is the most preferred
    {
        std::vector<uint64_t> keys(n);
        for (uint32_t i = 0; i < n; i++) keys[i] = ((uint64_t)d.num_leaves[i] << 32) | d.tie_index[i];
        std::vector<uint32_t> ord(n);
        std::iota(ord.begin(), ord.end(), 0u);
        auto before = [&](uint32_t a, uint32_t b) { return keys[a] != keys[b] ? keys[a] > keys[b] : a < b; };
        // sorted runs by thread, then pairwise merges (a strict total order: the result does not depend on the split)
        const unsigned nt = host_threads(n);
        std::vector<uint32_t> cut(nt + 1);
        for (unsigned c = 0; c <= nt; c++) cut[c] = (uint32_t)((uint64_t)n * c / nt);
        parallel_chunks(n, nt, [&](unsigned, uint32_t lo, uint32_t hi) { std::sort(ord.begin() + lo, ord.begin() + hi, before); });
        for (unsigned w = 1; w < nt; w *= 2) {
            std::vector<std::thread> pool;
            for (unsigned c = 0; c + w < nt; c += 2 * w)
                pool.emplace_back([&, c]() {
                    std::inplace_merge(ord.begin() + cut[c], ord.begin() + cut[c + w], ord.begin() + cut[std::min(nt, c + 2 * w)], before);
                });
            for (auto& th : pool) th.join();
        }
        d.tiekey.resize(n);
        d.key_to_node = ord;
        for (uint32_t r = 0; r < n; r++) d.tiekey[ord[r]] = r;
    }
    pt.lap("tie-break rank (sort)");
    // ---- mutation checks, genome extent (chunks of nodes in parallel; the first error in node order is reported)
    int64_t maxpos = 0;
    uint64_t kept = 0;
    std::vector<uint32_t> row_kept_of(n);
    {
        const unsigned nt = host_threads(f.n_mutations + n);
        struct Part { int64_t maxpos = 0; uint64_t kept = 0; uint32_t max_row = 0; int rc = 0; std::string err; };
        std::vector<Part> part(nt);
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            Part& P = part[c];
            auto bad = [&](int rc, std::string msg) { P.rc = rc; P.err = std::move(msg); };
            for (uint32_t i = lo; i < hi && !P.rc; i++) {
                if (f.row_ptr[i + 1] < f.row_ptr[i]) { bad(UB200_E_ARG, "flat MAT: row_ptr not monotone"); break; }
                int32_t last = INT32_MIN;
                uint64_t row_kept = 0;
                for (uint64_t k = f.row_ptr[i]; k < f.row_ptr[i + 1]; k++) {
                    const ub200_mutation& m = f.mutations[k];
                    if (m.position < last) { bad(UB200_E_POSITION, "flat MAT: row " + std::to_string(i) + " is not position-sorted"); break; }
                    if (m.position >= 0 && m.position == last) {
                        bad(UB200_E_POSITION, "flat MAT: row " + std::to_string(i) + " repeats position " + std::to_string(last));
                        break;
                    }
                    last = m.position;
                    if (m.position < 0) continue;
                    if ((uint32_t)m.position > kMaxPos) { bad(UB200_E_POSITION, "flat MAT: position >= 2^26-1"); break; }
                    if (nuc_code(m.mut_nuc) < 0 || nuc_code(m.ref_nuc) < 0) {
                        bad(UB200_E_NOT_ONE_HOT, "flat MAT: node " + std::to_string(i) + " position " + std::to_string(m.position) +
                                                     " has a non-one-hot ref/mut nucleotide");
                        break;
                    }
                    P.maxpos = std::max<int64_t>(P.maxpos, m.position);
                    row_kept++;
                }
                if (P.rc) break;
                P.max_row = std::max<uint32_t>(P.max_row, (uint32_t)row_kept);
                if (row_kept > kMaxRow) { bad(UB200_E_LIMIT, "flat MAT: a branch with more than 65534 mutations"); break; }
                row_kept_of[i] = (uint32_t)row_kept;
                P.kept += row_kept;
            }
        });
        for (auto& P : part) {
            if (P.rc) { err = P.err; return P.rc; }
            maxpos = std::max(maxpos, P.maxpos);
            kept += P.kept;
            d.max_row = std::max(d.max_row, P.max_row);
        }
    }
    if (kept >= (1ull << 32) - kMutChunk) { err = "flat MAT: more than 2^32 mutations"; return UB200_E_LIMIT; }
    d.m = kept;
    d.L = (uint32_t)maxpos + 1;
    d.ref_of.assign(d.L, 0);
    d.root_init_extra = (int32_t)(f.row_ptr[1] - f.row_ptr[0]);
    // reference allele per position: any writer wins (relaxed byte stores; all writers agree on a consistent tree),
    // then every mutation is checked against what was kept, so an inconsistent tree fails whoever won
    parallel_chunks(n, host_threads(f.n_mutations), [&](unsigned, uint32_t lo, uint32_t hi) {
        for (uint64_t k = f.row_ptr[lo]; k < f.row_ptr[hi]; k++) {
            const ub200_mutation& m = f.mutations[k];
            if (m.position >= 0 && __atomic_load_n(&d.ref_of[m.position], __ATOMIC_RELAXED) == 0)
                __atomic_store_n(&d.ref_of[m.position], m.ref_nuc, __ATOMIC_RELAXED);
        }
    });
    {
        const unsigned nt = host_threads(f.n_mutations);
        std::vector<int64_t> bad_pos(nt, -1);
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            for (uint64_t k = f.row_ptr[lo]; k < f.row_ptr[hi]; k++) {
                const ub200_mutation& m = f.mutations[k];
                if (m.position >= 0 && d.ref_of[m.position] != m.ref_nuc) { bad_pos[c] = m.position; break; }
            }
        });
        for (int64_t bp : bad_pos)
            if (bp >= 0) {
                err = "flat MAT: tree mutations disagree on the reference allele at position " + std::to_string(bp);
                return UB200_E_ARG;
            }
    }

    pt.lap("mutation checks");
    // ---- path states: one DFS with a live state array per chunk of the node range.  A chunk starts from the state of
    // its first node's root path (rebuilt by walking that path root-first); every node writes only its own rows.
    d.row32.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; i++) d.row32[i + 1] = d.row32[i] + row_kept_of[i];
    std::vector<uint32_t>().swap(row_kept_of);
    d.mutw.assign(((kept + kMutChunk - 1) / kMutChunk + 1) * kMutChunk, 0);
    d.hdr.assign(((size_t)n + kHdrChunk - 1) / kHdrChunk * kHdrChunk + kHdrChunk, NodeHdr{0, 0, 0, 0});
    std::vector<int32_t> dref(n, 0);
    {
        const unsigned nt = host_threads(kept + n);
        // chunk boundaries by mutation count
        std::vector<uint32_t> cut(nt + 1, n);
        cut[0] = 0;
        for (unsigned c = 1; c < nt; c++)
            cut[c] = (uint32_t)(std::lower_bound(d.row32.begin(), d.row32.begin() + n, (uint32_t)(kept * c / nt)) - d.row32.begin());
        for (unsigned c = 1; c <= nt; c++) cut[c] = std::max(cut[c], cut[c - 1]);
        auto chunk = [&](unsigned c) {
            const uint32_t lo = cut[c], hi = cut[c + 1];
            if (lo >= hi) return;
            std::vector<uint8_t> state(d.L, 0);  // 0 = never mutated on the current path, else one-hot
            struct Undo { uint32_t node; uint32_t pos; uint8_t old; };
            std::vector<Undo> undo;
            std::vector<uint32_t> path;
            std::vector<int32_t> dref_anc;       // Dref of the ancestors of `lo`, by level
            // root path of the chunk's first node, root first
            for (int32_t a = lo ? f.parent[lo] : -1; a >= 0; a = f.parent[a]) path.push_back((uint32_t)a);
            std::reverse(path.begin(), path.end());
            for (uint32_t a : path) {
                int32_t dd = 0;
                for (uint64_t k = f.row_ptr[a]; k < f.row_ptr[a + 1]; k++) {
                    const ub200_mutation& m = f.mutations[k];
                    if (m.position < 0) continue;
                    const uint32_t pos = (uint32_t)m.position;
                    const uint8_t prev = state[pos] ? state[pos] : m.ref_nuc;
                    dd += (m.mut_nuc != m.ref_nuc) - (prev != m.ref_nuc);
                    undo.push_back({a, pos, state[pos]});
                    state[pos] = m.mut_nuc;
                }
                dref_anc.push_back((dref_anc.empty() ? 0 : dref_anc.back()) + dd);
            }
            for (uint32_t i = lo; i < hi; i++) {
                while (!path.empty() && (int32_t)path.back() != f.parent[i]) {
                    uint32_t top = path.back();
                    path.pop_back();
                    while (!undo.empty() && undo.back().node == top) {
                        state[undo.back().pos] = undo.back().old;
                        undo.pop_back();
                    }
                }
                const bool is_root = (i == 0);
                const bool leaf = nchild[i + 1] == 0;
                bool masked = false;
                int32_t dd = 0, a0 = 0;
                uint32_t c0 = 0, nm = 0;
                uint64_t w = d.row32[i];
                for (uint64_t k = f.row_ptr[i]; k < f.row_ptr[i + 1]; k++) {
                    const ub200_mutation& m = f.mutations[k];
                    if (m.position < 0) { masked = true; continue; }
                    const uint32_t pos = (uint32_t)m.position;
                    const uint8_t prev = state[pos] ? state[pos] : m.ref_nuc;
                    const int rp = prev != m.ref_nuc, rm = m.mut_nuc != m.ref_nuc;
                    dd += rm - rp;
                    if (!rm) { c0++; a0 += rp; }   // LOOP 1 for an absent position: common iff back to ref (:244-259)
                    d.mutw[w++] = pack_mut(pos, (uint32_t)nuc_code(m.ref_nuc), (uint32_t)nuc_code(prev),
                                           (uint32_t)nuc_code(m.mut_nuc));
                    undo.push_back({i, pos, state[pos]});
                    state[pos] = m.mut_nuc;
                    nm++;
                }
                const int32_t par = f.parent[i];
                const int32_t dpar = is_root ? 0 : ((uint32_t)par >= lo ? dref[par] : dref_anc[d.level[par]]);
                dref[i] = dpar + dd;
                if (masked || is_root) { a0 = 0; c0 = 0; }  // masked: LOOP 1 breaks before taking anything (:197-200)
                const bool hu0 = masked || (nm > c0);
                const bool valid0 = is_root || (leaf ? c0 > 0 : (!hu0 || c0 > 0));
                NodeHdr h;
                h.g = is_root ? dref[i] : dpar - a0;
                h.tiekey = d.tiekey[i];
                const uint32_t plane = (!is_root && (uint32_t)par >= (i & ~31u)) ? 1u + ((uint32_t)par & 31u) : 0u;
                h.level_flags = (d.level[i] << kLevelShift) | (plane << 8) | (leaf ? kFlagLeaf : 0) | (masked ? kFlagMasked : 0) |
                                (is_root ? kFlagRoot : 0) | (valid0 ? kFlagValid0 : 0) | ((hu0 && !is_root) ? kFlagHu0 : 0);
                h.nmut_c0 = (nm << 16) | c0;
                d.hdr[i] = h;
                path.push_back(i);
            }
        };
        if (nt <= 1) chunk(0);
        else {
            std::vector<std::thread> pool;
            for (unsigned c = 0; c < nt; c++) pool.emplace_back(chunk, c);
            for (auto& th : pool) th.join();
        }
    }

    pt.lap("path-state DFS");
    // ---- tiles: contiguous DFS ranges of roughly equal cost (mutations + per-node overhead)
    {
        const uint64_t node_cost = 4;
        const uint64_t total = kept + node_cost * n;
        uint64_t per = total / (target_tiles ? target_tiles : 1);
        per = std::min<uint64_t>(std::max<uint64_t>(per, min_tile_cost ? min_tile_cost : 6144), 1u << 16);   // >= ~180 nodes: keeps root-path seeding < 10%
        d.tile_start.clear();
        d.tile_start.push_back(0);
        uint64_t acc = 0;
        for (uint32_t i = 0; i < n; i++) {
            acc += (d.row32[i + 1] - d.row32[i]) + node_cost;
            if (acc >= per && i + 1 < n) {
                d.tile_start.push_back(i + 1);
                acc = 0;
            }
        }
        d.tile_start.push_back(n);
        const size_t T = d.tile_start.size() - 1;
        d.anc_ptr.assign(T + 1, 0);
        d.anc.clear();
        std::vector<uint32_t> chain;
        for (size_t t = 0; t < T; t++) {
            chain.clear();
            for (int32_t a = f.parent[d.tile_start[t]]; a >= 0; a = f.parent[a]) chain.push_back((uint32_t)a);
            d.anc.insert(d.anc.end(), chain.rbegin(), chain.rend());
            d.anc_ptr[t + 1] = (uint32_t)d.anc.size();
        }
    }
    pt.lap("k_score tiles");
    derive3(f, target_tiles, min_tile_cost, d);
    pt.lap("segment layout (derive3)");
    if (d.have3 && d.stream.size() / 4 >= (1ull << 32)) { err = "flat MAT: stream longer than 2^34 words"; return UB200_E_LIMIT; }
    return UB200_OK;
}

}  // namespace ub200

// ---- host-only inspection hooks (used by the CPU test-suite to check the derivation without a GPU) ----
extern "C" {

struct ub200_derived_view {
    uint32_t n_nodes, genome_len, max_level, n_tiles;
    uint64_t n_mutations;
    const uint32_t* level; const uint32_t* tie_index; const uint32_t* num_leaves; const uint32_t* tiekey;
    const uint32_t* key_to_node; const uint32_t* row32; const uint32_t* mutw; const void* hdr;
    const uint8_t* ref_of; const uint32_t* tile_start; const uint32_t* anc_ptr; const uint32_t* anc;
    // k_score3 layout
    uint32_t n_tiles3, n_seed_segs, narrow3, reserved3;
    uint64_t stream_words;
    const uint32_t* stream; const void* hdr3; const uint32_t* tile3_start; const uint32_t* tile3_w0;
    const uint32_t* tile3_lvl; const uint32_t* tile3_sseg; const uint32_t* seed_end; const uint32_t* blk_words;
    const uint32_t* blk_rec;
};

int ub200_debug_derive(const ub200_flat_mat* flat, uint32_t target_tiles, uint32_t min_tile_cost, void** handle,
                       ub200_derived_view* view, char* errbuf, size_t errlen) {
    auto* d = new ub200::Derived();
    std::string err;
    int rc = ub200::derive(*flat, target_tiles, *d, err, min_tile_cost);
    if (rc != UB200_OK) {
        if (errbuf && errlen) { std::strncpy(errbuf, err.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
        delete d;
        return rc;
    }
    view->n_nodes = d->n; view->genome_len = d->L; view->max_level = d->max_level;
    view->n_tiles = (uint32_t)d->tile_start.size() - 1; view->n_mutations = d->m;
    view->level = d->level.data(); view->tie_index = d->tie_index.data(); view->num_leaves = d->num_leaves.data();
    view->tiekey = d->tiekey.data(); view->key_to_node = d->key_to_node.data(); view->row32 = d->row32.data();
    view->mutw = d->mutw.data(); view->hdr = d->hdr.data(); view->ref_of = d->ref_of.data();
    view->tile_start = d->tile_start.data(); view->anc_ptr = d->anc_ptr.data(); view->anc = d->anc.data();
    view->n_tiles3 = d->have3 ? (uint32_t)d->tile3_start.size() - 1 : 0;
    view->n_seed_segs = (uint32_t)d->seed_end.size();
    view->narrow3 = d->narrow3 ? 1u : 0u; view->reserved3 = 0;
    view->stream_words = d->stream.size();
    view->stream = d->stream.data(); view->hdr3 = d->hdr3.data(); view->tile3_start = d->tile3_start.data();
    view->tile3_w0 = d->tile3_w0.data(); view->tile3_lvl = d->tile3_lvl.data();
    view->tile3_sseg = d->tile3_sseg.data(); view->seed_end = d->seed_end.data();
    view->blk_words = d->blk_words.data();
    view->blk_rec = d->blk_rec.data();
    *handle = d;
    return UB200_OK;
}
void ub200_debug_derive_free(void* handle) { delete (ub200::Derived*)handle; }
}
