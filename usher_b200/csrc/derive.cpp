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

// Sort the keys of one segment (bits [4, 4 + kbits) carry position and lane, unique inside a segment).  Segments are a
// few hundred to a thousand words: an LSD radix sort with 7..8-bit digits beats a comparison sort about three times.
static void sort_segment(std::vector<uint32_t>& seg, std::vector<uint32_t>& tmp, uint32_t kbits) {
    const size_t m = seg.size();
    if (m <= 48) { std::sort(seg.begin(), seg.end()); return; }
    const uint32_t passes = (kbits + 7) / 8, dbits = (kbits + passes - 1) / passes, mask = (1u << dbits) - 1u;
    uint32_t hist[4][256];
    for (uint32_t p = 0; p < passes; p++) std::memset(hist[p], 0, sizeof(uint32_t) << dbits);
    for (size_t i = 0; i < m; i++) {
        const uint32_t k = seg[i] >> 4;
        for (uint32_t p = 0; p < passes; p++) hist[p][(k >> (p * dbits)) & mask]++;
    }
    tmp.resize(m);
    uint32_t* src = seg.data();
    uint32_t* dst = tmp.data();
    for (uint32_t p = 0; p < passes; p++) {
        uint32_t* h = hist[p];
        uint32_t run = 0;
        for (uint32_t b = 0; b <= mask; b++) { const uint32_t c = h[b]; h[b] = run; run += c; }
        const uint32_t sh = 4 + p * dbits;
        for (size_t i = 0; i < m; i++) dst[h[(src[i] >> sh) & mask]++] = src[i];
        std::swap(src, dst);
    }
    if (src != seg.data()) seg.swap(tmp);
}

// The k_score3 layout (ub200_internal.h), built from the arrays derive() has already filled.
static void derive3(const ub200_flat_mat& f, uint32_t target_tiles, uint32_t min_tile_cost, Derived& d) {
    const uint32_t n = d.n;
    d.have3 = d.L <= kMaxPos3;
    if (!d.have3) return;
    d.narrow3 = d.L <= kMaxPos3Narrow && !getenv("UB200_WIDE_WORDS");   // env: test hook for the wide form
    const uint32_t nblk = (n + 31) / 32;
    PhaseTimer pt;
    // words on the root path above a node (the seed stream of a tile that starts there)
    auto path_words = [&](uint32_t node) {
        uint64_t w = 0;
        for (int32_t a = f.parent[node]; a >= 0; a = f.parent[a]) w += d.row32[a + 1] - d.row32[a];
        return w;
    };
    // headers: y = in-block ancestor mask, flags + open.  A node is open when its subtree runs past its block, i.e.
    // when it is an ancestor of the next block's first node (DFS pre-order: subtrees are contiguous).
    d.hdr3.resize(d.hdr.size());
    std::copy(d.hdr.begin() + n, d.hdr.end(), d.hdr3.begin() + n);   // the zero padding
    parallel_chunks(nblk, host_threads(n), [&](unsigned, uint32_t blo, uint32_t bhi) {
        for (uint32_t b = blo; b < bhi; b++) {
            const uint32_t n0 = b * 32, n1 = std::min(n, n0 + 32);
            uint32_t open = 0, am[32];
            if (n0 + 32 < n)
                for (int32_t a = f.parent[n0 + 32]; a >= (int32_t)n0; a = f.parent[a]) open |= 1u << (a & 31);
            for (uint32_t i = n0; i < n1; i++) {
                const uint32_t p = (uint32_t)f.parent[i];
                am[i & 31] = (i && p >= n0) ? (am[p & 31] | (1u << (p & 31))) : 0u;
                NodeHdr& h = d.hdr3[i];
                h = d.hdr[i];
                const uint32_t flags = hdr_flags(h.level_flags);
                h.tiekey = am[i & 31];
                h.level_flags = (d.level[i] << kLevelShift) | flags | (((open >> (i & 31)) & 1u) ? kFlagOpen : 0u);
            }
        }
    });
    pt.lap("  3: headers");
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
                seed += path_words(e);
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
    const bool pad_steps = !getenv("UB200_NO_STEP_PAD");   // developer switch: measure what the padding buys
    // Length a segment of `raw` words takes when it is appended at word `at` of its tile's piece: a multiple of 4;
    // the kernel takes the piece in steps of 512 words (4 rows) counted from the tile's start and hands the hits of a
    // step out segment by segment, so a segment that starts on a step boundary is padded up to the next one when that
    // costs at most an eighth of its length: it then spans the fewest possible steps, every step it touches belongs to
    // it alone, and the segments behind it stay aligned.
    auto seg_len = [&](size_t at, size_t raw) {
        size_t m = (raw + 3) & ~(size_t)3;
        if (pad_steps && at % 512 == 0) {
            const size_t padw = (512 - m % 512) % 512;
            if (padw && padw * 8 <= m) m += padw;
        }
        return m;
    };
    // Every tile's piece of the stream starts on a 1 KB boundary and its length follows from the row lengths alone:
    // pass 1 lays the pieces out (sizes, seed-segment ends, block words), pass 2 has one host thread per slice of tiles
    // sort and write its pieces straight into the stream.
    struct Piece { uint64_t words = 0, seed_words = 0; uint32_t nseed = 0; };
    std::vector<Piece> pieces(T);
    auto root_chain = [&](uint32_t n0, std::vector<uint32_t>& chain) {   // chain[level] = ancestor of n0 at that level
        chain.assign(d.level[n0], 0);
        for (int32_t a = f.parent[n0]; a >= 0; a = f.parent[a]) chain[d.level[a]] = (uint32_t)a;
    };
    parallel_chunks((uint32_t)T, host_threads(n), [&](unsigned, uint32_t tlo, uint32_t thi) {
        std::vector<uint32_t> chain;
        for (uint32_t t = tlo; t < thi; t++) {
            const uint32_t n0 = d.tile3_start[t], n1 = d.tile3_start[t + 1], lvl0 = d.level[n0];
            d.tile3_lvl[t] = lvl0;
            root_chain(n0, chain);
            size_t at = 0;
            for (uint32_t l0 = 0; l0 < lvl0; l0 += 32) {
                size_t raw = 0;
                for (uint32_t l = l0; l < std::min(lvl0, l0 + 32); l++) raw += d.row32[chain[l] + 1] - d.row32[chain[l]];
                at += seg_len(at, raw);
                pieces[t].nseed++;
            }
            pieces[t].seed_words = at;
            for (uint32_t b = n0; b < n1; b += 32) {
                const size_t m = seg_len(at, d.row32[std::min(n1, b + 32)] - d.row32[b]);
                d.blk_words[b >> 5] = (uint32_t)m;
                at += m;
            }
            pieces[t].words = (at + kChunk3 - 1) / kChunk3 * kChunk3;
        }
    });
    std::vector<uint64_t> piece_off(T + 1, 0);
    d.seed_words = 0;
    for (size_t t = 0; t < T; t++) {
        piece_off[t + 1] = piece_off[t] + pieces[t].words;
        d.tile3_w0[t] = (uint32_t)(piece_off[t] / kChunk3);
        d.tile3_sseg[t + 1] = d.tile3_sseg[t] + pieces[t].nseed;
        d.seed_words += pieces[t].seed_words;
    }
    d.tile3_w0[T] = (uint32_t)(piece_off[T] / kChunk3);
    d.seed_end.assign((size_t)d.tile3_sseg[T] + 1, 0);   // never empty (one spare entry)
    d.stream.resize(piece_off[T]);
    pt.lap("  3: piece layout");
    // sort key of a stream word: pos | lane:5 | prev:2 | mut:2 -- the order of (position, stream word)
    uint32_t pbits = 1;
    while ((1ull << pbits) <= d.L) pbits++;
    const uint32_t kbits = pbits + 5;
    const uint32_t padk = d.L << 9;
    auto build_tile = [&](size_t t, std::vector<uint32_t>& chain, std::vector<uint32_t>& seg, std::vector<uint32_t>& tmp) {
        uint32_t* const out = d.stream.data() + piece_off[t];
        size_t at = 0;
        // A segment's words go out sorted by position and transposed inside every 128-word row piece: the
        // scanner reads a row with one 16-byte load per lane, so component j of the 32 lanes should hold 32
        // CONSECUTIVE sorted words -> their bitmap words are consecutive too and the 32 bitmap reads fall into
        // distinct shared-memory banks (a random order costs ~3.5 wavefronts per read).  Which node (or level) a
        // word belongs to is in the word, so the order inside a segment is free.
        auto emit_segment = [&]() {
            sort_segment(seg, tmp, kbits);
            seg.resize(seg_len(at, seg.size()), padk);
            size_t s0 = 0;
            while (s0 < seg.size()) {
                const size_t m = std::min<size_t>(128 - at % 128, seg.size() - s0), nl = m / 4;
                for (size_t j = 0; j < 4; j++) {
                    const uint32_t* src = seg.data() + s0 + j * nl;
                    for (size_t i = 0; i < nl; i++) {
                        const uint32_t k = src[i];
                        out[at + 4 * i + j] = pack_mut3(nw, k >> 9, (k >> 4) & 31u, (k >> 2) & 3u, k & 3u);
                    }
                }
                s0 += m;
                at += m;
            }
            seg.clear();
        };
        auto take_row = [&](uint32_t node, uint32_t lane) {
            for (uint32_t k = d.row32[node]; k < d.row32[node + 1]; k++) {
                const uint32_t w = d.mutw[k];
                seg.push_back(((w >> 6) << 9) | (lane << 4) | (w & 15u));
            }
        };
        const uint32_t n0 = d.tile3_start[t], n1 = d.tile3_start[t + 1];
        const uint32_t lvl0 = d.level[n0];
        root_chain(n0, chain);
        uint32_t* se = d.seed_end.data() + d.tile3_sseg[t];
        for (uint32_t l0 = 0; l0 < lvl0; l0 += 32) {
            for (uint32_t l = l0; l < std::min(lvl0, l0 + 32); l++) take_row(chain[l], l & 31u);
            emit_segment();
            *se++ = (uint32_t)((piece_off[t] + at) / 4);
        }
        for (uint32_t b = n0; b < n1; b += 32) {
            for (uint32_t i = b; i < std::min(n1, b + 32); i++) take_row(i, i & 31u);
            emit_segment();
        }
        const uint32_t pad = pack_mut3(nw, d.L, 0, 0, 0);
        while (at < pieces[t].words) out[at++] = pad;
    };
    {
        const unsigned nthr = host_threads(d.m);   // (tiles are handed out one by one: uneven tiles balance themselves)
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            std::vector<uint32_t> chain, seg, tmp;
            for (size_t t = next.fetch_add(1); t < T; t = next.fetch_add(1)) build_tile(t, chain, seg, tmp);
        };
        std::vector<std::thread> pool;
        for (unsigned i = 1; i < nthr; i++) pool.emplace_back(worker);
        worker();
        for (auto& th : pool) th.join();
    }
    pt.lap("  3: tile pieces (threads)");
    // ---- block records: what the consumer needs of a block that cannot hold an optimum (score_kernel4.cuh)
    d.blk_rec.resize((size_t)nblk * 4);
    parallel_chunks(nblk, host_threads(n), [&](unsigned, uint32_t blo, uint32_t bhi) {
        for (uint32_t bi = blo; bi < bhi; bi++) {
            const uint32_t b = bi * 32, e = std::min(n, b + 32);
            int32_t omin = INT32_MAX;
            uint32_t open = 0, lv0 = 0;
            for (uint32_t i = b; i < e; i++) {
                omin = std::min(omin, d.hdr3[i].g - (int32_t)(d.hdr3[i].nmut_c0 >> 16));
                if (d.hdr3[i].level_flags & kFlagOpen) {
                    if (!open) lv0 = d.level[i];
                    open |= 1u << (i & 31);
                }
            }
            uint32_t* r = &d.blk_rec[(size_t)bi * 4];
            r[0] = (uint32_t)omin; r[1] = open; r[2] = lv0; r[3] = d.blk_words[bi];
        }
    });
    pt.lap("  3: block records");
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
    // ---- topology checks: DFS pre-order <=> parent[i] lies on the root path of node i-1.  Chunks of the node range
    // in parallel: a chunk starts from the root path of the node before it, rebuilt by walking the parents (safe once
    // every parent is known to be an earlier node); the first error in node order is the one reported.
    d.level.resize(n);
    d.level[0] = 0;
    if (f.parent[0] != -1) { err = "flat MAT: node 0 must be the root (parent -1)"; return UB200_E_TREE_ORDER; }
    {
        const unsigned nt = host_threads(n);
        std::vector<uint32_t> first_bad(nt, n);
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            for (uint32_t i = std::max(lo, 1u); i < hi; i++)
                if (f.parent[i] < 0 || (uint32_t)f.parent[i] >= i) { first_bad[c] = i; break; }
        });
        const uint32_t bad_parent = *std::min_element(first_bad.begin(), first_bad.end());
        struct Bad { uint32_t node; int rc; const char* what; };
        std::vector<Bad> bad(nt, Bad{n, 0, nullptr});
        parallel_chunks(bad_parent, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {   // nodes before the first bad parent
            std::vector<uint32_t> path;
            if (lo == 0) { path.push_back(0); lo = 1; }
            else {
                for (int32_t a = (int32_t)lo - 1; a >= 0; a = f.parent[a]) path.push_back((uint32_t)a);
                std::reverse(path.begin(), path.end());
            }
            for (uint32_t i = lo; i < hi; i++) {
                const uint32_t p = (uint32_t)f.parent[i];
                while (!path.empty() && path.back() != p) path.pop_back();
                if (path.empty()) { bad[c] = Bad{i, UB200_E_TREE_ORDER, "flat MAT: nodes are not in DFS pre-order at node "}; return; }
                d.level[i] = (uint32_t)path.size();   // = level of the parent + 1
                if (d.level[i] > kMaxLevel) { bad[c] = Bad{i, UB200_E_LIMIT, "flat MAT: tree deeper than 2^18-1 at node "}; return; }
                path.push_back(i);
            }
        });
        for (const Bad& x : bad)
            if (x.rc) { err = x.what + std::to_string(x.node); return x.rc; }
        if (bad_parent < n) {
            err = "flat MAT: parent[" + std::to_string(bad_parent) + "] is not an earlier node";
            return UB200_E_TREE_ORDER;
        }
    }
    d.max_level = *std::max_element(d.level.begin(), d.level.end());
    pt.lap("topology checks");
    // ---- leaves, leaf counts, BFS index (chunks of the node range in parallel)
    // DFS pre-order: a node is a leaf iff the next node is not its child.
    auto is_leaf = [&](uint32_t i) { return i + 1 == n || (uint32_t)f.parent[i + 1] != i; };
    {
        // leaf counts: a reverse sweep inside every chunk; what a chunk owes to nodes before it goes to ancestors of
        // the chunk's first node (subtrees are contiguous), collected per ancestor level and handed up afterwards
        const unsigned nt = host_threads(n);
        std::vector<uint32_t> cut(nt + 1);
        for (unsigned c = 0; c <= nt; c++) cut[c] = (uint32_t)((uint64_t)n * c / nt);
        d.num_leaves.assign(n, 0);
        std::vector<std::vector<std::pair<uint32_t, uint32_t>>> owed(nt);   // (level of the ancestor, leaves)
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            for (uint32_t i = hi; i-- > lo;) {
                if (is_leaf(i)) d.num_leaves[i] = 1;
                if (!i) continue;
                const uint32_t p = (uint32_t)f.parent[i];
                if (p >= lo) d.num_leaves[p] += d.num_leaves[i];
                else owed[c].push_back({d.level[p], d.num_leaves[i]});
            }
        });
        for (unsigned c = 1; c < nt; c++) {
            if (owed[c].empty()) continue;
            std::vector<uint32_t> by_level(d.level[cut[c]], 0);
            for (auto& o : owed[c]) by_level[o.first] += o.second;
            uint32_t run = 0;
            for (int32_t a = f.parent[cut[c]]; a >= 0; a = f.parent[a]) {   // deepest ancestor first
                run += by_level[d.level[a]];
                d.num_leaves[a] += run;
            }
        }
    }
    d.tie_index.resize(n);
    if (f.tie_index) {
        std::memcpy(d.tie_index.data(), f.tie_index, sizeof(uint32_t) * n);
    } else {
        // breadth-first index (children in DFS order): nodes of one level appear in DFS order, so the index is a
        // stable counting sort of the nodes by level
        const unsigned nt = host_threads(n);
        const uint32_t nl = d.max_level + 1;
        std::vector<std::vector<uint32_t>> cnt(nt, std::vector<uint32_t>(nl, 0));
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; i++) cnt[c][d.level[i]]++;
        });
        uint32_t run = 0;
        for (uint32_t l = 0; l < nl; l++)
            for (unsigned c = 0; c < nt; c++) { const uint32_t k = cnt[c][l]; cnt[c][l] = run; run += k; }
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; i++) d.tie_index[i] = cnt[c][d.level[i]]++;
        });
    }
    pt.lap("leaves + BFS index");
    // ---- tie-break order: preferred = more leaves, then larger j  -> tiekey 0 is the most preferred; equal keys (a
    // caller's tie_index may repeat) keep node order.  A stable LSD radix sort of (~key, node) in 11-bit digits, every
    // pass split over the host threads; digits on which all keys agree are skipped.
    {
        const unsigned nt = host_threads(n);
        std::vector<uint64_t, NoInitAlloc<uint64_t>> ka(n), kb(n);
        std::vector<uint32_t, NoInitAlloc<uint32_t>> ia(n), ib(n);
        std::vector<uint64_t> diff_part(nt, 0);
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            uint64_t df = 0;
            const uint64_t k0 = ~(((uint64_t)d.num_leaves[0] << 32) | d.tie_index[0]);
            for (uint32_t i = lo; i < hi; i++) {
                ka[i] = ~(((uint64_t)d.num_leaves[i] << 32) | d.tie_index[i]);
                ia[i] = i;
                df |= ka[i] ^ k0;
            }
            diff_part[c] = df;
        });
        uint64_t differ = 0;   // bits on which some keys differ
        for (uint64_t x : diff_part) differ |= x;
        constexpr uint32_t kDigit = 11, kBuckets = 1u << kDigit;
        std::vector<uint32_t> hist((size_t)nt * kBuckets);
        uint64_t* ks = ka.data(); uint64_t* kd = kb.data();
        uint32_t* is = ia.data(); uint32_t* id = ib.data();
        for (uint32_t sh = 0; sh < 64; sh += kDigit) {
            if (((differ >> sh) & (kBuckets - 1)) == 0) continue;
            std::fill(hist.begin(), hist.end(), 0u);
            parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
                uint32_t* h = &hist[(size_t)c * kBuckets];
                for (uint32_t i = lo; i < hi; i++) h[(ks[i] >> sh) & (kBuckets - 1)]++;
            });
            uint32_t run = 0;
            for (uint32_t b = 0; b < kBuckets; b++)
                for (unsigned c = 0; c < nt; c++) { uint32_t& h = hist[(size_t)c * kBuckets + b]; const uint32_t k = h; h = run; run += k; }
            parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
                uint32_t* h = &hist[(size_t)c * kBuckets];
                for (uint32_t i = lo; i < hi; i++) {
                    const uint32_t at = h[(ks[i] >> sh) & (kBuckets - 1)]++;
                    kd[at] = ks[i];
                    id[at] = is[i];
                }
            });
            std::swap(ks, kd);
            std::swap(is, id);
        }
        d.tiekey.resize(n);
        d.key_to_node.resize(n);
        parallel_chunks(n, nt, [&](unsigned, uint32_t lo, uint32_t hi) {
            for (uint32_t r = lo; r < hi; r++) { d.key_to_node[r] = is[r]; d.tiekey[is[r]] = r; }
        });
    }
    pt.lap("tie-break rank (sort)");
    // ---- mutation checks, genome extent (chunks of nodes in parallel; the first error in node order is reported)
    int64_t maxpos = 0;
    uint64_t kept = 0;
    std::vector<uint32_t> row_kept_of(n);
    {
        const unsigned nt = host_threads(f.n_mutations + n);
        // ref = the chunk's own view of the reference allele per position (grown on demand): one pass over the mutations
        // checks them, and the views are merged afterwards -- positions whose mutations disagree fail either way
        struct Part { int64_t maxpos = 0, bad_ref = -1; uint64_t kept = 0; uint32_t max_row = 0; int rc = 0; std::string err; std::vector<uint8_t> ref; };
        std::vector<Part> part(nt);
        parallel_chunks(n, nt, [&](unsigned c, uint32_t lo, uint32_t hi) {
            // running values live in locals (the byte stores into the reference view would otherwise force every field
            // of the shared record to be re-read per mutation); one-hot test: exactly one of the low four bits
            Part& P = part[c];
            auto bad = [&](int rc, std::string msg) { P.rc = rc; P.err = std::move(msg); };
            std::vector<uint8_t> ref;
            uint8_t* rp = nullptr;
            size_t rn = 0;
            int64_t maxp = 0, bad_ref = -1;
            uint64_t kept_c = 0;
            uint32_t max_row = 0;
            auto one_hot = [](uint8_t x) { return x && x < 16 && !(x & (x - 1)); };
            for (uint32_t i = lo; i < hi && !P.rc; i++) {
                const uint64_t k0 = f.row_ptr[i], k1 = f.row_ptr[i + 1];
                if (k1 < k0 || k1 > f.n_mutations) { bad(UB200_E_ARG, "flat MAT: row_ptr not monotone"); break; }
                int32_t last = INT32_MIN;
                uint32_t row_kept = 0;
                for (uint64_t k = k0; k < k1; k++) {
                    const ub200_mutation m = f.mutations[k];
                    const int32_t pos = m.position;
                    if (pos < last) { bad(UB200_E_POSITION, "flat MAT: row " + std::to_string(i) + " is not position-sorted"); break; }
                    if (pos >= 0 && pos == last) {
                        bad(UB200_E_POSITION, "flat MAT: row " + std::to_string(i) + " repeats position " + std::to_string(last));
                        break;
                    }
                    last = pos;
                    if (pos < 0) continue;
                    if ((uint32_t)pos > kMaxPos) { bad(UB200_E_POSITION, "flat MAT: position >= 2^26-1"); break; }
                    if (!one_hot(m.mut_nuc) || !one_hot(m.ref_nuc)) {
                        bad(UB200_E_NOT_ONE_HOT, "flat MAT: node " + std::to_string(i) + " position " + std::to_string(pos) +
                                                     " has a non-one-hot ref/mut nucleotide");
                        break;
                    }
                    if ((size_t)pos >= rn) {
                        ref.resize(std::max<size_t>((size_t)pos + 1, rn * 2), 0);
                        rp = ref.data(); rn = ref.size();
                    }
                    const uint8_t r = rp[pos];
                    if (!r) rp[pos] = m.ref_nuc;
                    else if (r != m.ref_nuc && bad_ref < 0) bad_ref = pos;
                    row_kept++;
                }
                if (P.rc) break;
                if (last > maxp) maxp = last;   // rows are position-sorted: the last one is the row's largest
                max_row = std::max(max_row, row_kept);
                if (row_kept > kMaxRow) { bad(UB200_E_LIMIT, "flat MAT: a branch with more than 65534 mutations"); break; }
                row_kept_of[i] = row_kept;
                kept_c += row_kept;
            }
            P.maxpos = maxp; P.bad_ref = bad_ref; P.kept = kept_c; P.max_row = max_row; P.ref.swap(ref);
        });
        for (auto& P : part) {
            if (P.rc) { err = P.err; return P.rc; }
            maxpos = std::max(maxpos, P.maxpos);
            kept += P.kept;
            d.max_row = std::max(d.max_row, P.max_row);
        }
        d.L = (uint32_t)maxpos + 1;
        d.ref_of.assign(d.L, 0);
        int64_t bad_ref = -1;
        for (auto& P : part)
            if (P.bad_ref >= 0 && bad_ref < 0) bad_ref = P.bad_ref;
        std::vector<int64_t> bad_merge(nt, -1);
        parallel_chunks(d.L, d.L < (1u << 16) ? 1u : nt, [&](unsigned c, uint32_t plo, uint32_t phi) {   // slices of the genome
            for (auto& P : part)
                for (size_t pos = plo, e = std::min<size_t>(P.ref.size(), phi); pos < e; pos++) {
                    const uint8_t v = P.ref[pos];
                    if (!v) continue;
                    if (!d.ref_of[pos]) d.ref_of[pos] = v;
                    else if (d.ref_of[pos] != v && bad_merge[c] < 0) bad_merge[c] = (int64_t)pos;
                }
        });
        for (int64_t b : bad_merge)
            if (b >= 0 && bad_ref < 0) bad_ref = b;
        if (bad_ref >= 0) {
            err = "flat MAT: tree mutations disagree on the reference allele at position " + std::to_string(bad_ref);
            return UB200_E_ARG;
        }
    }
    if (kept >= (1ull << 32) - kMutChunk) { err = "flat MAT: more than 2^32 mutations"; return UB200_E_LIMIT; }
    d.m = kept;
    d.root_init_extra = (int32_t)(f.row_ptr[1] - f.row_ptr[0]);

    pt.lap("mutation checks");
    // ---- path states: one DFS with a live state array per chunk of the node range.  A chunk starts from the state of
    // its first node's root path (rebuilt by walking that path root-first); every node writes only its own rows.
    d.row32.assign((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; i++) d.row32[i + 1] = d.row32[i] + row_kept_of[i];
    std::vector<uint32_t>().swap(row_kept_of);
    d.mutw.resize(((kept + kMutChunk - 1) / kMutChunk + 1) * kMutChunk);   // not zero-filled: every row is written below
    std::fill(d.mutw.begin() + kept, d.mutw.end(), 0u);
    d.hdr.resize(((size_t)n + kHdrChunk - 1) / kHdrChunk * kHdrChunk + kHdrChunk);   // rows [0, n) are written below
    std::fill(d.hdr.begin() + n, d.hdr.end(), NodeHdr{0, 0, 0, 0});
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
            std::vector<Undo> undo;              // rows of the ancestors above the chunk only
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
                    if (top >= lo) {   // a row this chunk wrote: put back the state above it (ref allele = never mutated)
                        for (uint64_t k = d.row32[top]; k < d.row32[top + 1]; k++)
                            state[d.mutw[k] >> 6] = (uint8_t)(1u << ((d.mutw[k] >> 2) & 3u));
                    } else {
                        while (!undo.empty() && undo.back().node == top) {
                            state[undo.back().pos] = undo.back().old;
                            undo.pop_back();
                        }
                    }
                }
                const bool is_root = (i == 0);
                const bool leaf = is_leaf(i);
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
