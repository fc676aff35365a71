// k_score5 — streaming best-placement kernel on the segment layout, symmetric workers (DESIGN.md "Kernels").
//
// Measured on the B200 with in-kernel cycle counters (UB200_PROFILE build of k_score4): a warp that scans the
// mutation stream runs at ~0.06 instructions per cycle — every step is a chain of dependent shared-memory reads,
// votes and stores — and the SM issues less than one instruction in five cycles per scheduler.  The scoring pass is
// bound by the latency of each warp's own dependent chain, not by issue slots, shared memory or HBM; what it needs
// is more independent instruction streams per SM.  k_score4's scanner/consumer pairs spend half the warps waiting
// and 6 KB of shared memory per pair on the hand-over (ring + message slots).  Here every warp is a WORKER that
// does the whole job for its own tile, with no hand-over at all:
//
//   * the stream goes global -> REGISTERS: a step = 4 coalesced 512 B rows = four LDG.128 per lane, issued one step
//     ahead (the next step's 16 words are in flight while the current 16 are tested), so there is no ring, no
//     cp.async bookkeeping and no re-read of hit words from shared memory: a hit word is stored from the register
//     it was tested in;
//   * per word: byte offset of the bitmap word, LDS, wrap shift, funnel shift collecting the hit bits (as before);
//     the hits of a step are compacted (two ballots) into the worker's own hit list, segment by segment;
//   * at the end of a segment the worker fetches the table rows of its hits (one 256-bit load per hit, two hits per
//     lane in flight) and applies them (sparse: lane = hit with shared-memory atomics; dense: 32 x 32 bit transpose,
//     lane = sample, no atomics), then runs the block phases of k_score4 unchanged: exact bound, E/F evaluation of
//     the blocks that can still hold an optimum, open-chain stack rows; seed segments initialise the stack.
//   The L2 latency of the table rows is hidden by the other 25 workers of the SM, not by a partner warp.
// Shared memory per worker: dnode 4 KB + stack 2 KB + neg + scratch + hit list 1 KB = 8 KB -> 26 workers per SM
// (k_score4: 16 scanners).  All pruning is exact, so results are schedule-independent.
#pragma once
#include "score_kernel4.cuh"

namespace ub200 {

constexpr int kWorkers5 = 26;
constexpr int kThreads5 = kWorkers5 * 32;
constexpr int kStack5 = 32;                                // stack levels kept in shared memory (deeper: HBM spill)
constexpr uint32_t kHitCap5 = 256;                         // hit words a worker collects before it must apply them
constexpr uint32_t kO5Dnode = 0;                           // i32[32][32] packed deltas
constexpr uint32_t kO5Stack = 4096;                        // i16[kStack5][32]
constexpr uint32_t kO5Neg = kO5Stack + kStack5 * 64;       // i32[32]
constexpr uint32_t kO5Area = kO5Neg + 128;                 // u32[224]: dense-form staging / header copies
constexpr uint32_t kO5Hits = kO5Area + kPairCap4 * 4;      // u32[kHitCap5]
constexpr uint32_t kWorker5 = kO5Hits + kHitCap5 * 4;      // 8192
constexpr uint32_t kFixed5 = kLut4Bytes + kWorkers5 * kWorker5;

struct Score5Params {
    const uint32_t* stream;
    const NodeHdr* hdr;           // hdr3
    const uint32_t* tiekey;
    const uint4* blk_rec;         // [blocks] x = min(G - nmut), y = open-chain mask, z = level of the first open node,
                                  //          w = stream words of the block's segment
    const uint32_t* tile_start;   // [T+1]
    const uint32_t* tile_w0;      // [T+1]
    const uint32_t* tile_lvl;     // [T]
    const uint32_t* tile_sseg;    // [T+1]
    const uint32_t* seed_end;
    uint32_t n_nodes, n_tiles, L, bitmap_words;
    const uint32_t* bitmap;       // [all groups][bitmap_words]
    const uint32_t* tab;          // [groups][L][8]: mask, ref<<4, nibbles[4], -, -
    int32_t* gbest;
    uint32_t n_samples, group0, ngroups;
    unsigned long long* part_key;
    uint32_t* part_cnt;
    int32_t* gstack;
    uint32_t gstack_levels;
    const int32_t* target_rel;
    uint32_t* set_out;
    const unsigned long long* set_ptr;
    uint32_t* set_fill;
    uint32_t* tile_counter;
    unsigned long long* prof;
};

// COLLECT = false: best placement per sample.  COLLECT = true: second pass that lists every optimal node of each
// sample (best_j_vec + node_has_unique); the final best score is the bound, so almost every block is pruned.
template <bool SMEM_BITMAP, bool COLLECT, bool NARROW>
__global__ void __launch_bounds__(kThreads5, 1) k_score5(const Score5Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x % p.ngroups;
    const uint32_t cta_in_group = blockIdx.x / p.ngroups;
    const uint32_t ctas_per_group = gridDim.x / p.ngroups;
    const uint32_t ggroup = p.group0 + group;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int BIG = 0x3fffffff;
#ifdef UB200_PROFILE
    unsigned long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_start = clock64();
#endif

    // ---- shared memory: [bitmap][lut][worker 0 .. worker kWorkers5-1]
    uint32_t* bm_s = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bm_bytes = SMEM_BITMAP ? ((p.bitmap_words * 4u + 127u) & ~127u) : 0u;
    int* lut = reinterpret_cast<int*>(smem + bm_bytes);
    uint8_t* wbase = smem + bm_bytes + kLut4Bytes + warp * kWorker5;
    int* dnode = reinterpret_cast<int*>(wbase + kO5Dnode);
    int16_t* stk = reinterpret_cast<int16_t*>(wbase + kO5Stack);
    int* neg = reinterpret_cast<int*>(wbase + kO5Neg);
    uint32_t* area = reinterpret_cast<uint32_t*>(wbase + kO5Area);
    uint32_t* hits = reinterpret_cast<uint32_t*>(wbase + kO5Hits);
    const uint32_t hits_a = smem_u32(hits), bm_a = smem_u32(bm_s);

    const uint32_t* bm_g = p.bitmap + (size_t)ggroup * p.bitmap_words;
    if (SMEM_BITMAP) {
        const uint4* src = reinterpret_cast<const uint4*>(bm_g);
        uint4* dst = reinterpret_cast<uint4*>(bm_s);
        for (uint32_t i = threadIdx.x; i < p.bitmap_words / 4; i += kThreads5) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = threadIdx.x; i < 1024; i += kThreads5) lut[i] = lut_delta4(i);
    __syncthreads();

    const uint32_t* tabg = p.tab + (size_t)ggroup * p.L * 8u;
    int32_t* gstk = p.gstack ? p.gstack + ((size_t)(blockIdx.x * kWorkers5 + warp) * p.gstack_levels) * 32u : nullptr;
    const uint32_t sample = ggroup * 32u + lane;
    const bool live = sample < p.n_samples;

    auto stack_read = [&](uint32_t level, uint32_t s) -> int {
        if (__builtin_expect(level >= (uint32_t)kStack5, 0)) return spill_read4(gstk, level - kStack5, s);
        return stk[level * 32u + s];
    };
    auto stack_write = [&](uint32_t level, uint32_t s, int v) {
        if (__builtin_expect(level >= (uint32_t)kStack5, 0)) spill_write4(gstk, level - kStack5, s, v);
        else stk[level * 32u + s] = (int16_t)v;
    };

    // per-lane (= sample) running best (COLLECT: the known final best, fixed)
    int bsc = COLLECT ? (live ? p.target_rel[sample] : (int)0x80000000) : 0x7fffffff;
    unsigned long long bkey = ~0ull;
    uint32_t cnt = 0;
    auto merge = [&](int sc, uint32_t hu, uint32_t node) {
        if (COLLECT) {
            if (sc == bsc) {
                const uint32_t k = atomicAdd(p.set_fill + sample, 1u);
                p.set_out[p.set_ptr[sample] + k] = node | (hu ? 0x80000000u : 0u);
            }
            return;
        }
        const uint32_t tiekey = __ldg(p.tiekey + node);
        const unsigned long long key =
            ((unsigned long long)(uint32_t)(sc + kScoreBias) << 33) | ((unsigned long long)tiekey << 1) | hu;
        if (sc < bsc) { bsc = sc; cnt = 1; bkey = key; }
        else if (sc == bsc) { cnt++; if (key < bkey) bkey = key; }
    };
    int negr = 0;   // this lane's (= sample's) share of neg accumulated by the dense form of the hit phase
    auto zero_dnode = [&]() {
#pragma unroll
        for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(dnode)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
        neg[lane] = 0;
        negr = 0;
    };
    // ---- hit phase (lane = hit while the table rows are fetched), see score_kernel4.cuh
    auto apply_hit = [&](uint32_t w, const uint4& r0, const uint4& r1) {
        const uint32_t nl = (w >> 9) & 31u;
        const uint32_t lo = r0.y | ((w >> 5) & 15u);
        uint32_t pm = r0.x;
        while (pm) {
            const uint32_t s = __ffs(pm) - 1;
            pm &= pm - 1;
            const uint32_t nw = (s & 16u) ? ((s & 8u) ? r1.y : r1.x) : ((s & 8u) ? r0.w : r0.z);
            const uint32_t e4 = (nw >> ((s & 7u) * 4u)) & 15u;
            const int d = lut[(e4 << 6) | lo];
            atomicAdd(&dnode[nl * 32u + s], d);
            const int dc = dc_of(d);
            if (dc < 0) atomicAdd(&neg[s], dc);
        }
    };
    auto dense32 = [&](uint32_t w, const uint4& r0, const uint4& r1, uint32_t t) {
        area[lane] = r0.z;               // word j of hit h at j * 32 + ((h + 8 j) & 31): conflict-free for the
        area[32u + ((lane + 8u) & 31u)] = r0.w;     // usual access patterns of the loop below
        area[64u + ((lane + 16u) & 31u)] = r1.x;
        area[96u + ((lane + 24u) & 31u)] = r1.y;
        area[128u + lane] = ((w >> 9) & 31u) | ((r0.y | ((w >> 5) & 15u)) << 5);
        __syncwarp();
        const uint32_t j = lane >> 3, sh = (lane & 7u) * 4u;
        while (t) {
            const uint32_t h = __ffs(t) - 1;
            t &= t - 1;
            const uint32_t info = area[128u + h];
            const uint32_t e4 = (area[j * 32u + ((h + 8u * j) & 31u)] >> sh) & 15u;
            const int d = lut[(e4 << 6) | (info >> 5)];
            dnode[(info & 31u) * 32u + lane] += d;
            negr += min(dc_of(d), 0);
        }
        __syncwarp();
    };
    // 32 hits: the sparse form costs ~20 issue slots per caller of the busiest hit, the dense one ~12 per hit of the
    // busiest sample (an N run makes ONE sample own most of 32 position-sorted hits: sparse wins there)
    auto half = [&](uint32_t w, const uint4& r0, const uint4& r1) {
        const uint32_t mp = __reduce_max_sync(FULL, (uint32_t)__popc(r0.x));
        if (mp > 2u) {
            const uint32_t t = transpose32(r0.x, lane);
            const uint32_t mt = __reduce_max_sync(FULL, (uint32_t)__popc(t));
            if (12u * mt + 10u < 20u * mp) {
                dense32(w, r0, r1, t);
                return;
            }
        }
        if (r0.x) apply_hit(w, r0, r1);
    };
    uint32_t hl_fill = 0;   // hit words collected for the current segment (warp-uniform)
    // fetch the table rows of the collected hits (64 at a time, two per lane in flight) and apply them
    auto process_hits = [&]() {
        PROF_T0(tp);
        __syncwarp();
        for (uint32_t h0 = 0; h0 < hl_fill; h0 += 64u) {
            const uint32_t n = min(64u, hl_fill - h0);
            const bool ha = lane < n, hb2 = lane + 32u < n;
            uint32_t w0 = 0, w1 = 0;
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
            if (ha) {
                w0 = hits[h0 + lane];
                ldg_row(tabg + (size_t)mut3_pos<NARROW>(w0) * 8u, a0, a1);
            }
            if (hb2) {
                w1 = hits[h0 + 32u + lane];
                ldg_row(tabg + (size_t)mut3_pos<NARROW>(w1) * 8u, b0, b1);
            }
            half(w0, a0, a1);
            if (n > 32u) half(w1, b0, b1);
        }
        __syncwarp();
        hl_fill = 0;
        PROF_ADD(11, tp);
    };

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(p.tile_counter + group, 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= p.n_tiles) break;
        const uint32_t n0 = p.tile_start[t], n1 = p.tile_start[t + 1];
        const uint32_t lvl0 = p.tile_lvl[t], sseg = p.tile_sseg[t];
        const uint32_t w0 = p.tile_w0[t], w1 = p.tile_w0[t + 1];
        const uint32_t rows_end = w1 * (kChunk3 / 128u);
        const uint32_t nseed = (lvl0 + 31u) >> 5;
        const uint32_t nb = (n1 - n0 + 31u) >> 5;
        const uint32_t nseg = nseed + nb;
        // block records: one 16-byte word per block (same address for every lane), fetched one block ahead
        const uint4* recp = p.blk_rec + (n0 >> 5);
        uint4 rnext = __ldg(recp);
        uint4 rec = rnext;

        // cross-worker bound of this lane's sample, and the tile-local floor of every stack value
        const int gb = COLLECT ? bsc : (live ? *(volatile int*)(p.gbest + sample) : 0x7fffffff);
        int gmin = 0;

        uint32_t off = w0 * kChunk3;     // current stream word (multiple of 4)
        uint32_t si = 0;                 // current segment: seeds first, then one per block
        uint32_t seg_end;
        auto next_block_rec = [&]() {    // record of block segment si (>= nseed); prefetch the one after
            rec = rnext;
            if (si - nseed + 1u < nb) rnext = __ldg(++recp);
        };
        if (nseed) seg_end = p.seed_end[sseg] * 4u;
        else { next_block_rec(); seg_end = off + rec.w; }
        zero_dnode();
        __syncwarp();

        // one step = 4 rows of 128 words: lane l holds words 4l..4l+3 of each row
        auto load_row = [&](uint32_t base, uint32_t k) -> uint4 {
            const uint32_t row = (base >> 7) + k;
            if (row < rows_end) return __ldg(reinterpret_cast<const uint4*>(p.stream + ((size_t)row << 7)) + lane);
            return make_uint4(0, 0, 0, 0);
        };
        uint4 a0 = load_row(off, 0), a1 = load_row(off, 1), a2 = load_row(off, 2), a3 = load_row(off, 3);

        for (uint32_t base = off;; base += 512u) {
            PROF_T0(tl);
            const uint4 q0 = a0, q1 = a1, q2 = a2, q3 = a3;
            // the next step's rows are in flight while this one is tested and its hits are applied
            a0 = load_row(base + 512u, 0); a1 = load_row(base + 512u, 1);
            a2 = load_row(base + 512u, 2); a3 = load_row(base + 512u, 3);
            uint32_t acc = 0;                                  // hit bits enter at bit 31, oldest ends lowest
            auto test4 = [&](const uint4& q) {
                acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.x), 0u, q.x), 1u);
                acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.y), 0u, q.y), 1u);
                acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.z), 0u, q.z), 1u);
                acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.w), 0u, q.w), 1u);
            };
            test4(q0); test4(q1); test4(q2); test4(q3);
            const uint32_t hb_step = acc >> 16;                // bit 4k+j = word j of quad k (row k of the step)
            PROF_ADD(4, tl); PROF_INC(5, 1);

            // append the hits `hbp` (bits of hb_step) to the hit list: lane l's hits follow those of lanes < l
            auto emit_part = [&](uint32_t hbp) {
                const uint32_t c = __popc(hbp);
                uint32_t excl, total;
                if (__ballot_sync(FULL, c > 3u) == 0u) {
                    const uint32_t v0 = __ballot_sync(FULL, c & 1u), v1 = __ballot_sync(FULL, c & 2u);
                    if ((v0 | v1) == 0u) return;
                    excl = __popc(v0 & lt_mask) + 2u * __popc(v1 & lt_mask);
                    total = __popc(v0) + 2u * __popc(v1);
                } else {
                    uint32_t incl = c;
#pragma unroll
                    for (int dlt = 1; dlt < 32; dlt <<= 1) {
                        const uint32_t v = __shfl_up_sync(FULL, incl, dlt);
                        if (lane >= (uint32_t)dlt) incl += v;
                    }
                    total = __shfl_sync(FULL, incl, 31);
                    excl = incl - c;
                }
                if (hl_fill + total > kHitCap5) process_hits();   // list full: apply what is there (same segment)
                uint32_t pa = hits_a + ((hl_fill + excl) << 2);
#define UB200_PUT(bit, val) if (hbp & (1u << (bit))) { sts32_4(pa, (val)); pa += 4u; }
                UB200_PUT(0, q0.x) UB200_PUT(1, q0.y) UB200_PUT(2, q0.z) UB200_PUT(3, q0.w)
                UB200_PUT(4, q1.x) UB200_PUT(5, q1.y) UB200_PUT(6, q1.z) UB200_PUT(7, q1.w)
                UB200_PUT(8, q2.x) UB200_PUT(9, q2.y) UB200_PUT(10, q2.z) UB200_PUT(11, q2.w)
                UB200_PUT(12, q3.x) UB200_PUT(13, q3.y) UB200_PUT(14, q3.z) UB200_PUT(15, q3.w)
#undef UB200_PUT
                hl_fill += total;
            };

            bool tile_done = false;
            for (;;) {
                const uint32_t lim = min(seg_end, base + 512u);
                if (lim > off) {
                    PROF_T0(te);
                    const uint32_t idx = base + 4u * lane;
                    uint32_t hb = hb_step;
                    if (off != base || lim != base + 512u) {          // quads outside [off, lim): other segments
                        const int lo = (int)(off - idx), hi = (int)(lim - idx);   // multiples of 4
                        const uint32_t k_lo = lo > 0 ? min((uint32_t)(lo + 127) >> 7, 4u) : 0u;
                        const uint32_t k_hi = hi > 0 ? min((uint32_t)(hi + 127) >> 7, 4u) : 0u;
                        hb &= ((1u << (4u * k_hi)) - 1u) & ~((1u << (4u * k_lo)) - 1u);
                    }
                    // a step can hold up to 512 hits, the list 256: very dense steps go in two halves
                    if (__any_sync(FULL, __popc(hb) > 8)) { emit_part(hb & 0x00ffu); emit_part(hb & 0xff00u); }
                    else emit_part(hb);
                    PROF_ADD(6, te);
                }
                off = lim;
                if (off == seg_end) {
                    // ================= end of a segment =================
                    process_hits();
                    if (si < nseed) {
                        // ---- seed: path corrections of levels 32 si .. of the tile's root path -> stack rows
                        const uint32_t l0 = si << 5;
                        const uint32_t cn = min(32u, lvl0 - l0);
                        int v = l0 ? stack_read(l0 - 1u, lane) : 0;
                        for (uint32_t j = 0; j < cn; j++) {
                            v += dc_of(dnode[j * 32u + lane]);
                            stack_write(l0 + j, lane, v);
                            gmin = min(gmin, v);
                        }
                    } else {
                        const uint32_t blk = n0 + ((si - nseed) << 5);
                        PROF_T0(tn);
                        // ---- bound: can any pair of this block still be optimal?
                        const int lbase = gmin + neg[lane] + negr;
                        const int bound = min(bsc, gb);
                        const uint32_t needs = __ballot_sync(FULL, live && (int)rec.x + lbase <= bound);
                        if (needs) {
                            PROF_INC(15, 1);
                            // ---- A: headers (lane = node), only for blocks that get here
                            const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.hdr) + blk + lane);
                            const bool act = blk + lane < n1;
                            const uint32_t level = h.z >> kLevelShift, flags = h.z & 0x3fffu;
                            const bool dense_ok = act && (flags & kFlagValid0);
                            const int min_g = __reduce_min_sync(FULL, dense_ok ? (int)h.x : BIG);   // signed: G can be < 0
                            // hit nodes of this lane's sample = non-zero column entries (a pair whose packed delta is
                            // zero scores exactly like a pair without a hit)
                            uint32_t hmv = 0;
#pragma unroll 8
                            for (uint32_t n = 0; n < 32u; n++) hmv |= (dnode[n * 32u + lane] != 0 ? 1u : 0u) << n;
                            area[kA4G + lane] = (uint32_t)h.x;
                            area[kA4Z + lane] = h.z;
                            area[kA4W + lane] = h.w;
                            area[kA4Am + lane] = h.y;
                            area[kA4Hm + lane] = hmv;
                            __syncwarp();
                            // correction of the path above node (level, am) for sample s
                            auto above = [&](uint32_t lvl, uint32_t am, uint32_t hmask, uint32_t s) -> int {
                                const uint32_t top = lvl - __popc(am);
                                int v = top ? stack_read(top - 1u, s) : 0;
                                uint32_t m = am & hmask;
                                while (m) {
                                    const uint32_t a = __ffs(m) - 1;
                                    m &= m - 1;
                                    v += dc_of(dnode[a * 32u + s]);
                                }
                                return v;
                            };
                            // ---- E: non-hit pairs (lane = node), one sample at a time
                            uint32_t need_e = __ballot_sync(FULL, live && min_g < BIG && min_g + lbase <= bound);
                            while (need_e) {
                                const uint32_t s = __ffs(need_e) - 1;
                                need_e &= need_e - 1;
                                const uint32_t hm_s = area[kA4Hm + s];
                                const int sc = (int)h.x + above(level, h.y, hm_s, s);
                                const int bs = __shfl_sync(FULL, bsc, s);
                                uint32_t cm = __ballot_sync(FULL, dense_ok && !((hm_s >> lane) & 1u) && sc <= bs);
                                while (cm) {
                                    const uint32_t j = __ffs(cm) - 1;
                                    cm &= cm - 1;
                                    const int scj = __shfl_sync(FULL, sc, j);
                                    const uint32_t huj = __shfl_sync(FULL, (flags & kFlagHu0) ? 1u : 0u, j);
                                    if (lane == s) merge(scj, huj, blk + j);
                                }
                            }
                            // ---- F: hit pairs, exact (lane = sample)
                            if ((needs >> lane) & 1u) {
                                uint32_t hmw = hmv;
                                while (hmw) {
                                    const uint32_t n = __ffs(hmw) - 1;
                                    hmw &= hmw - 1;
                                    int dcorr, da, dcom;
                                    unpack_delta4(dnode[n * 32u + lane], dcorr, da, dcom);
                                    const uint32_t z = area[kA4Z + n], w = area[kA4W + n];
                                    const uint32_t fl = z & 0x3fffu;
                                    const int g = (int)area[kA4G + n];
                                    int sc;
                                    bool valid;
                                    uint32_t hu;
                                    if (fl & kFlagRoot) {
                                        sc = g + dcorr; valid = true; hu = 0;
                                    } else {
                                        const bool masked = fl & kFlagMasked;
                                        if (masked) { da = 0; dcom = 0; }
                                        sc = g + above(z >> kLevelShift, area[kA4Am + n], hmv, lane) - da;
                                        const int common = (int)(w & 0xffffu) + dcom;
                                        hu = (masked || (int)(w >> 16) > common) ? 1u : 0u;
                                        valid = (fl & kFlagLeaf) ? common > 0 : (!hu || common > 0);
                                    }
                                    if (valid && sc <= bsc) merge(sc, hu, blk + n);
                                }
                            }
                            __syncwarp();
                        }
                        // ---- G: stack rows of the open chain (lane = sample)
                        uint32_t chain = rec.y;
                        if (chain) {
                            uint32_t lv = rec.z;
                            int v = lv ? stack_read(lv - 1u, lane) : 0;
                            while (chain) {
                                const uint32_t n = __ffs(chain) - 1;
                                chain &= chain - 1;
                                v += dc_of(dnode[n * 32u + lane]);
                                stack_write(lv, lane, v);
                                gmin = min(gmin, v);
                                lv++;
                            }
                        }
                        PROF_ADD(14, tn);
                    }
                    __syncwarp();
                    if (++si == nseg) { tile_done = true; break; }
                    zero_dnode();
                    __syncwarp();
                    if (si < nseed) seg_end = p.seed_end[sseg + si] * 4u;
                    else { next_block_rec(); seg_end = off + rec.w; }
                    if (seg_end == off) continue;              // empty segment
                }
                if (off == base + 512u) break;
            }
            if (tile_done) break;
        }
        // publish an improved bound for the other workers of this sample group
        if (!COLLECT && live && bsc < gb) atomicMin(p.gbest + sample, bsc);
        __syncwarp();
    }

#ifdef UB200_PROFILE
    pc[0] = (unsigned long long)(clock64() - prof_start);
    if (lane == 0) for (int i = 0; i < 16; i++) atomicAdd(p.prof + i, pc[i]);
#endif
    if (COLLECT) return;
    // park the worker's result in its own rows and fold the CTA's workers: one partial row per CTA
    reinterpret_cast<unsigned long long*>(dnode)[lane] = bkey;
    neg[lane] = (int)cnt;
    __syncthreads();
    if (warp == 0) {
        unsigned long long best = ~0ull;
        for (int w = 0; w < kWorkers5; w++) {
            const uint8_t* pb = smem + bm_bytes + kLut4Bytes + w * kWorker5;
            best = min(best, reinterpret_cast<const unsigned long long*>(pb + kO5Dnode)[lane]);
        }
        uint32_t c = 0;
        for (int w = 0; w < kWorkers5; w++) {
            const uint8_t* pb = smem + bm_bytes + kLut4Bytes + w * kWorker5;
            if ((reinterpret_cast<const unsigned long long*>(pb + kO5Dnode)[lane] >> 33) == (best >> 33))
                c += reinterpret_cast<const uint32_t*>(pb + kO5Neg)[lane];
        }
        const size_t o = ((size_t)group * ctas_per_group + cta_in_group) * 32u + lane;
        p.part_key[o] = best;
        p.part_cnt[o] = c;
    }
}

}  // namespace ub200
