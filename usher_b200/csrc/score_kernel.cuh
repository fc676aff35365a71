// Scoring kernels (sm_100a).  See DESIGN.md "Kernels".
//
// k_score<MODE>: one persistent launch scores every node of the tree against NG groups of 32 samples.
//   * lane  = one sample of the group (its running state lives in registers / a per-warp smem stack);
//   * warp  = an independent worker streaming contiguous DFS tiles of the flattened MAT through its own
//             shared-memory rings, filled by 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP) that
//             complete on per-stage mbarriers;
//   * CTA   = 8 warps bound to one sample group (the group's position bitmap sits in shared memory);
//             CTAs of different groups walk the tile list in the same order so the MAT is read from HBM
//             once per launch and from L2 for the other groups.
// Per node the warp (a) tests each branch mutation's position against the group's bitmap (lane = mutation,
// one ballot per 32 mutations), (b) for the few hits broadcasts the mutation and lets every lane correct
// its sample's running distance from a position-major byte table, (c) combines the precomputed
// sample-independent header terms with the per-lane corrections into the branch parsimony score
// (closed form of reference src/usher_mapper.cpp:167-504, DESIGN.md "Closed form") and folds it into the
// per-lane best key (score, tie-break rank, has_unique) and optimal-placement count.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "ub200_internal.h"

namespace ub200 {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMutStages = 4;
constexpr int kHdrStages = 4;
constexpr int kStackDepth = 32;             // levels kept in shared memory per warp (deeper ones spill to HBM)
constexpr uint32_t kMutRingWords = kMutChunk * kMutStages;   // 1024 words = 4 KB
constexpr uint32_t kHdrRingNodes = kHdrChunk * kHdrStages;   // 128 headers = 2 KB
constexpr uint32_t kWarpSmemBytes = kMutRingWords * 4 + kHdrRingNodes * 16 + kStackDepth * 32 * 4 + 128;
constexpr int32_t kScoreBias = 1 << 28;
constexpr uint32_t kMaxSmemBitmapBytes = 64 * 1024;

enum ScoreMode { kModeBest = 0, kModeNodeScores = 1, kModeCollect = 2 };

struct ScoreParams {
    const uint32_t* mutw;
    const NodeHdr* hdr;
    const uint32_t* row32;
    const uint32_t* tile_start;
    const uint32_t* anc_ptr;
    const uint32_t* anc;
    uint32_t n_nodes;
    uint32_t n_tiles;
    uint32_t L;
    uint32_t bitmap_words;        // per group (multiple of 4)
    const uint32_t* bitmap;       // [all groups][bitmap_words]
    const uint32_t* tab;          // [all groups][L][8]: w0 = lanes that call the position, w1 = ref code << 4,
                                  // w2..w5 = cost nibbles
    int32_t* gbest;               // [samples] running upper bound of the best relative score (pruning only)
    const int32_t* base;          // [samples] LOOP-2 count against the pure reference genome
    uint32_t n_samples;
    uint32_t group0;              // first group of this launch
    uint32_t ngroups;             // groups in this launch; gridDim.x % ngroups == 0
    unsigned long long* part_key; // [ngroups][ctas_per_group][32]
    uint32_t* part_cnt;
    int32_t* gstack;              // spill: [total warps][gstack_levels][32]
    uint32_t gstack_levels;
    int32_t* node_scores;         // MODE 1: [n_samples][n_nodes]
    const int32_t* target_rel;    // MODE 2: per sample best score minus base
    uint32_t* set_out;            // MODE 2
    const unsigned long long* set_ptr;
    uint32_t* set_fill;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must trap (-> CUDA error through the C ABI), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
// same, for waits on another warp (not on a copy in flight): the hardware may keep the warp suspended longer
// between polls, so that a waiting warp does not eat issue slots
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 22)) __trap();
    }
}
// waits of a warp that has slack (the consumer on its scanner).  A polling warp costs issue slots and, since the
// barrier lives in shared memory, shared-memory wavefronts; __nanosleep() does not hold it back here (measured:
// 14 ns per iteration whatever the argument), so every failed poll is followed by one volatile global load whose
// ~700-cycle latency paces the loop at two instructions per iteration.
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, const int* pace) {
    uint32_t spins = 0, acc = 0;
    while (!mbar_try_wait(bar, parity)) {
        uint32_t x;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(x) : "l"(pace) : "memory");
        acc ^= x;
        if (++spins > (1u << 22)) __trap();
    }
    if (acc == 0x9e3779b9u && spins == 0x7fffffffu) __trap();   // keeps the load's result live
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA unit, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of the (converged) warp: lets ptxas feed UBLKCP from uniform registers without a per-lane loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- per-hit correction
// e: table byte of this lane's sample at the mutation's position: bit4 = sample calls this position,
//    bits0-3 = cost of each path state there (0 for an N call).  m: packed tree mutation (warp-uniform).
__device__ __forceinline__ void apply_hit(uint32_t m, uint32_t e, int& dcorr, int& da, int& dcom) {
    const uint32_t refc = (m >> 4) & 3u, prevc = (m >> 2) & 3u, mutc = m & 3u;
    const int rm = (mutc != refc), rp = (prevc != refc);  // cost vs the bare reference (what the header assumed)
    if (e & 0x10u) {
        const int wm = (e >> mutc) & 1u, wp = (e >> prevc) & 1u;
        dcorr += (wm - wp) - (rm - rp);
        const int tk = wm ^ 1;          // LOOP 1: mutation is shared with the sample (usher_mapper.cpp:204-242)
        const int t0 = rm ^ 1;          // ... what the header assumed (:244-259)
        da += (tk & wp) - (t0 & rp);
        dcom += tk - t0;
    }
}

// this lane's table entry (present<<4 | cost nibble) from the position's 32-byte row
__device__ __forceinline__ uint32_t tab_entry(const uint32_t* tabg, uint32_t pos, uint32_t lane) {
    const uint32_t* row = tabg + (size_t)pos * 8u;
    const uint32_t pm = __ldg(row);
    const uint32_t nw = __ldg(row + 2 + (lane >> 3));
    return (((pm >> lane) & 1u) << 4) | ((nw >> ((lane & 7u) * 4u)) & 15u);
}

template <bool SMEM_BITMAP>
__device__ __forceinline__ bool bitmap_test(const uint32_t* bm_s, const uint32_t* bm_g, uint32_t pos) {
    const uint32_t w = SMEM_BITMAP ? bm_s[pos >> 5] : __ldg(bm_g + (pos >> 5));
    return (w >> (pos & 31u)) & 1u;
}

template <int MODE, bool SMEM_BITMAP>
__global__ void __launch_bounds__(kThreads, 2) k_score(const ScoreParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x % p.ngroups;
    const uint32_t cta_in_group = blockIdx.x / p.ngroups;
    const uint32_t ctas_per_group = gridDim.x / p.ngroups;
    const uint32_t ggroup = p.group0 + group;
    const uint32_t FULL = 0xffffffffu;

    // ---- shared memory carve-up
    uint32_t* bm_s = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bm_bytes = SMEM_BITMAP ? p.bitmap_words * 4u : 0u;
    uint8_t* wbase = smem + ((bm_bytes + 127u) & ~127u) + warp * kWarpSmemBytes;
    uint32_t* mring = reinterpret_cast<uint32_t*>(wbase);
    uint4* hring = reinterpret_cast<uint4*>(wbase + kMutRingWords * 4);
    int32_t* stk = reinterpret_cast<int32_t*>(wbase + kMutRingWords * 4 + kHdrRingNodes * 16);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + kMutRingWords * 4 + kHdrRingNodes * 16 + kStackDepth * 128);
    const uint32_t mring_a = smem_u32(mring), hring_a = smem_u32(hring), bars_a = smem_u32(bars);

    const uint32_t* bm_g = p.bitmap + (size_t)ggroup * p.bitmap_words;
    if (SMEM_BITMAP) {
        const uint4* src = reinterpret_cast<const uint4*>(bm_g);
        uint4* dst = reinterpret_cast<uint4*>(bm_s);
        for (uint32_t i = threadIdx.x; i < p.bitmap_words / 4; i += kThreads) dst[i] = __ldg(src + i);
    }
    if (lane == 0) {
        for (int i = 0; i < kMutStages + kHdrStages; i++) mbar_init(bars_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t* tabg = p.tab + (size_t)ggroup * p.L * 8u;
    const uint32_t wig = cta_in_group * kWarpsPerCta + warp;     // warp index within the group
    const uint32_t wpg = ctas_per_group * kWarpsPerCta;          // warps per group
    int32_t* gstk = p.gstack ? p.gstack + ((size_t)(blockIdx.x * kWarpsPerCta + warp) * p.gstack_levels) * 32u : nullptr;
    const uint32_t sample = ggroup * 32u + lane;

    auto stack_read = [&](uint32_t level) -> int {
        return level < (uint32_t)kStackDepth ? stk[level * 32u + lane] : gstk[(size_t)(level - kStackDepth) * 32u + lane];
    };
    auto stack_write = [&](uint32_t level, int v) {
        if (level < (uint32_t)kStackDepth) stk[level * 32u + lane] = v;
        else gstk[(size_t)(level - kStackDepth) * 32u + lane] = v;
    };

    // per-lane running best (MODE 0)
    int bsc = 0x7fffffff;
    unsigned long long bkey = ~0ull;
    uint32_t cnt = 0;
    int target = 0;
    if (MODE == kModeCollect) target = sample < p.n_samples ? p.target_rel[sample] : 0x7fffffff;
    int base = 0;
    if (MODE == kModeNodeScores) base = sample < p.n_samples ? p.base[sample] : 0;

    uint32_t mphase = 0, hphase = 0;  // parity bit per stage

    for (uint32_t t = wig; t < p.n_tiles; t += wpg) {
        const uint32_t n0 = p.tile_start[t], n1 = p.tile_start[t + 1];
        const uint32_t ms = p.row32[n0], me = p.row32[n1];
        // chunk ranges (absolute, aligned): mutation chunk c = words [256c, 256c+256), header chunk = 32 nodes
        uint32_t mc_issue = ms / kMutChunk;
        const uint32_t mc_end = (me > ms) ? (me - 1) / kMutChunk + 1 : mc_issue;  // exclusive
        uint32_t mc_wait = mc_issue;
        uint32_t hc_issue = n0 / kHdrChunk;
        const uint32_t hc_end = (n1 - 1) / kHdrChunk + 1;
        uint32_t hc_wait = hc_issue;
        if (lane == 0) {
            for (int i = 0; i < kMutStages && mc_issue + i < mc_end; i++) {
                const uint32_t c = mc_issue + i, s = c % kMutStages;
                mbar_expect_tx(bars_a + 8 * s, kMutChunk * 4);
                bulk_g2s(mring_a + s * kMutChunk * 4, p.mutw + (size_t)c * kMutChunk, kMutChunk * 4, bars_a + 8 * s);
            }
            for (int i = 0; i < kHdrStages && hc_issue + i < hc_end; i++) {
                const uint32_t c = hc_issue + i, s = c % kHdrStages;
                mbar_expect_tx(bars_a + 8 * (kMutStages + s), kHdrChunk * 16);
                bulk_g2s(hring_a + s * kHdrChunk * 16, p.hdr + (size_t)c * kHdrChunk, kHdrChunk * 16,
                         bars_a + 8 * (kMutStages + s));
            }
        }
        mc_issue = min(mc_issue + kMutStages, mc_end);
        hc_issue = min(hc_issue + kHdrStages, hc_end);

        // ---- seed the stack with the running corrections of the tile's root path (rows read straight from HBM/L2)
        for (uint32_t ai = p.anc_ptr[t]; ai < p.anc_ptr[t + 1]; ai++) {
            const uint32_t a = p.anc[ai];
            const uint32_t lvl = hdr_level(p.hdr[a].level_flags);
            const uint32_t r0 = p.row32[a], r1 = p.row32[a + 1];
            int dcorr = 0, da = 0, dcom = 0;
            for (uint32_t i = r0; i < r1; i += 32) {
                const bool in = (i + lane) < r1;
                const uint32_t m = in ? __ldg(p.mutw + i + lane) : 0u;
                uint32_t hm = __ballot_sync(FULL, in && bitmap_test<SMEM_BITMAP>(bm_s, bm_g, m >> 6));
                while (hm) {
                    const int j = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const uint32_t mm = __shfl_sync(FULL, m, j);
                    apply_hit(mm, tab_entry(tabg, mm >> 6, lane), dcorr, da, dcom);
                }
            }
            const int up = lvl ? stack_read(lvl - 1) : 0;
            stack_write(lvl, up + dcorr);
        }
        __syncwarp();

        uint32_t rs = ms;
        uint32_t cur_level = 0xffffffffu;  // level of the previous node in this tile
        int ccur = 0;                      // its running correction
        for (uint32_t n = n0; n < n1; n++) {
            // ---- header ring
            const uint32_t hc = n / kHdrChunk;
            if (hc >= hc_wait) {  // entering a new header chunk
                const uint32_t s = hc % kHdrStages;
                mbar_wait(bars_a + 8 * (kMutStages + s), (hphase >> s) & 1u);
                hphase ^= 1u << s;
                hc_wait = hc + 1;
                // the previous chunk's stage is now free: refill it
                if (hc_issue < hc_end && hc_issue < hc + kHdrStages) {
                    __syncwarp();   // every lane is done with the chunk whose stage is refilled
                    if (lane == 0) {
                        const uint32_t c = hc_issue, s2 = c % kHdrStages;
                        mbar_expect_tx(bars_a + 8 * (kMutStages + s2), kHdrChunk * 16);
                        bulk_g2s(hring_a + s2 * kHdrChunk * 16, p.hdr + (size_t)c * kHdrChunk, kHdrChunk * 16,
                                 bars_a + 8 * (kMutStages + s2));
                    }
                    hc_issue++;
                }
            }
            const uint4 h = hring[n % kHdrRingNodes];
            const uint32_t level = hdr_level(h.z), flags = hdr_flags(h.z);
            const uint32_t nmut = h.w >> 16, c0 = h.w & 0xffffu;
            const bool root = flags & kFlagRoot;

            int cpar;
            if (root) cpar = 0;
            else if (level == cur_level + 1) cpar = ccur;
            else cpar = stack_read(level - 1);

            // ---- scan the branch's mutations
            int dcorr = 0, da = 0, dcom = 0;
            const uint32_t re = rs + nmut;
            for (uint32_t i = rs; i < re; i += 32) {
                const uint32_t last = min(i + 32u, re) - 1u;
                while (mc_wait <= last / kMutChunk) {
                    const uint32_t s = mc_wait % kMutStages;
                    mbar_wait(bars_a + 8 * s, (mphase >> s) & 1u);
                    mphase ^= 1u << s;
                    mc_wait++;
                }
                // chunks below i/kMutChunk are dead: refill their stages
                while (mc_issue < mc_end && mc_issue < i / kMutChunk + kMutStages) {
                    __syncwarp();   // every lane's reads of the dead stage are ordered before its refill
                    if (lane == 0) {
                        const uint32_t c = mc_issue, s = c % kMutStages;
                        mbar_expect_tx(bars_a + 8 * s, kMutChunk * 4);
                        bulk_g2s(mring_a + s * kMutChunk * 4, p.mutw + (size_t)c * kMutChunk, kMutChunk * 4,
                                 bars_a + 8 * s);
                    }
                    mc_issue++;
                }
                const bool in = (i + lane) < re;
                const uint32_t m = in ? mring[(i + lane) % kMutRingWords] : 0u;   // lanes past the row may fall in a chunk still in flight
                uint32_t hm = __ballot_sync(FULL, in && bitmap_test<SMEM_BITMAP>(bm_s, bm_g, m >> 6));
                while (hm) {
                    const int j = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const uint32_t mm = __shfl_sync(FULL, m, j);
                    apply_hit(mm, tab_entry(tabg, mm >> 6, lane), dcorr, da, dcom);
                }
            }
            rs = re;

            ccur = cpar + dcorr;
            cur_level = level;
            if (!(flags & kFlagLeaf)) stack_write(level, ccur);

            // ---- branch parsimony score of placing each lane's sample at this node
            const bool masked = flags & kFlagMasked;
            if (masked) { da = 0; dcom = 0; }
            const int sc = root ? (h.x + dcorr) : (h.x + cpar - da);
            if (MODE == kModeBest) {
                if (!masked || root) {
                    if (__any_sync(FULL, sc <= bsc)) {
                        const int common = (int)c0 + dcom;
                        const bool hu = !root && (masked || (int)nmut > common);
                        const bool valid = root || ((flags & kFlagLeaf) ? common > 0 : (!hu || common > 0));
                        if (valid && sc <= bsc) {
                            const unsigned long long key = ((unsigned long long)(uint32_t)(sc + kScoreBias) << 33) |
                                                           ((unsigned long long)h.y << 1) | (hu ? 1ull : 0ull);
                            if (sc < bsc) { bsc = sc; cnt = 1; bkey = key; }
                            else { cnt++; if (key < bkey) bkey = key; }
                        }
                    }
                }
            } else {
                const int common = (int)c0 + dcom;
                const bool hu = !root && (masked || (int)nmut > common);
                const bool valid = root || ((flags & kFlagLeaf) ? common > 0 : (!hu || common > 0));
                if (MODE == kModeNodeScores) {
                    if (sample < p.n_samples)
                        p.node_scores[(size_t)sample * p.n_nodes + n] = sc + base + (valid ? 0 : 1);
                } else {
                    if (valid && sc == target && sample < p.n_samples) {
                        const uint32_t k = atomicAdd(p.set_fill + sample, 1u);
                        p.set_out[p.set_ptr[sample] + k] = n | (hu ? 0x80000000u : 0u);
                    }
                }
            }
        }
        __syncwarp();
    }

    if (MODE == kModeBest) {
        // fold the CTA's 8 warps in shared memory (the rings are dead now), one partial row per CTA
        __syncthreads();
        unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem);
        uint32_t* scnt = reinterpret_cast<uint32_t*>(smem + kWarpsPerCta * 32 * 8);
        skey[warp * 32 + lane] = bkey;
        scnt[warp * 32 + lane] = cnt;
        __syncthreads();
        if (warp == 0) {
            unsigned long long best = ~0ull;
            for (int w = 0; w < kWarpsPerCta; w++) best = min(best, skey[w * 32 + lane]);
            uint32_t c = 0;
            for (int w = 0; w < kWarpsPerCta; w++)
                if ((skey[w * 32 + lane] >> 33) == (best >> 33)) c += scnt[w * 32 + lane];
            const size_t o = ((size_t)group * ctas_per_group + cta_in_group) * 32u + lane;
            p.part_key[o] = best;
            p.part_cnt[o] = c;
        }
    }
}

// One block per sample group: fold the per-warp partial bests into the final placement records.
struct ReduceParams {
    const unsigned long long* part_key;
    const uint32_t* part_cnt;
    uint32_t wpg;          // partial rows per group
    uint32_t stride;       // rows reserved per group (>= wpg)
    uint32_t part_group0;  // first group's row block
    uint32_t group0;
    uint32_t n_samples;
    const int32_t* base;
    const uint32_t* key_to_node;
    const uint32_t* tie_index;
    const uint32_t* num_leaves;
    ub200_placement* out;
    int32_t* best_rel;   // per sample: best score minus base (input of the collect pass)
};

__global__ void __launch_bounds__(256) k_reduce(const ReduceParams p) {
    __shared__ unsigned long long skey[8][32];
    __shared__ uint32_t scnt[8][32];
    const uint32_t lane = threadIdx.x & 31u, slice = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x;
    const unsigned long long* pk = p.part_key + (size_t)(p.part_group0 + group) * p.stride * 32u;
    const uint32_t* pc = p.part_cnt + (size_t)(p.part_group0 + group) * p.stride * 32u;
    unsigned long long best = ~0ull;
    for (uint32_t w = slice; w < p.wpg; w += 8) best = min(best, pk[(size_t)w * 32u + lane]);
    skey[slice][lane] = best;
    __syncthreads();
    for (int i = 0; i < 8; i++) best = min(best, skey[i][lane]);
    uint32_t c = 0;
    for (uint32_t w = slice; w < p.wpg; w += 8) {
        const unsigned long long k = pk[(size_t)w * 32u + lane];
        if ((k >> 33) == (best >> 33)) c += pc[(size_t)w * 32u + lane];
    }
    scnt[slice][lane] = c;
    __syncthreads();
    if (slice == 0) {
        uint32_t tot = 0;
        for (int i = 0; i < 8; i++) tot += scnt[i][lane];
        const uint32_t s = (p.group0 + group) * 32u + lane;
        if (s < p.n_samples) {
            const int rel = (int)(uint32_t)(best >> 33) - kScoreBias;
            const uint32_t node = p.key_to_node[(uint32_t)((best >> 1) & 0xffffffffu)];
            ub200_placement r;
            r.score = rel + p.base[s];
            r.best_node = node;
            r.best_j = p.tie_index[node];
            r.num_best = tot;
            r.has_unique = (uint32_t)(best & 1ull);
            r.best_num_leaves = p.num_leaves[node];
            r.reserved[0] = 0;
            r.reserved[1] = 0;
            p.out[s] = r;
            p.best_rel[s] = rel;
        }
    }
}

// ---- sample-side tables -------------------------------------------------------------------------------
// One thread per sample call: set the group's position bit, write the lane's table byte, count the calls
// that already disagree with the bare reference (LOOP 2 on an unmutated path, usher_mapper.cpp:292-388).
struct PrepParams {
    const ub200_mutation* calls;
    const unsigned long long* sample_ptr;   // [n_samples+1]
    const uint32_t* call_sample;            // [n_calls] owning sample of each call
    unsigned long long n_calls;
    uint32_t L;
    uint32_t bitmap_words;
    uint32_t* bitmap;
    uint32_t* tab;
    int32_t* base;
    // union bitmaps: groups are scored in passes of `pass_groups`, `nc` consecutive groups of a pass share one scan of
    // the mutation stream and therefore one bitmap (nc = 1: one bitmap per group)
    uint32_t pass_groups, nc, nsg_per_pass;
};

__global__ void k_prep_scatter(const PrepParams p) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_calls) return;
    const ub200_mutation c = p.calls[i];
    const uint32_t s = p.call_sample[i];
    const uint32_t g = s >> 5, lane = s & 31u;
    const uint32_t set = c.mut_nuc & 15u;
    if (!c.is_missing && (set & c.ref_nuc) == 0) atomicAdd(p.base + s, 1);
    const uint32_t pos = (uint32_t)c.position;
    if (pos < p.L) {
        const uint32_t ub = (g / p.pass_groups) * p.nsg_per_pass + (g % p.pass_groups) / p.nc;
        atomicOr(p.bitmap + (size_t)ub * p.bitmap_words + (pos >> 5), 1u << (pos & 31u));
        const uint32_t cost = c.is_missing ? 0u : (~set & 15u);
        uint32_t* row = p.tab + ((size_t)g * p.L + pos) * 8u;
        atomicOr(row, 1u << lane);
        row[1] = (uint32_t)(31 - __clz((int)c.ref_nuc)) << 4;   // same value from every caller (checked on upload)
        if (cost) atomicOr(row + 2 + (lane >> 3), cost << ((lane & 7u) * 4u));
    }
}

// gbest[s] = the reference's initial bound (usher_common.cpp:374: |S| + |root row| + 1) relative to base[s]
__global__ void k_prep_bound(const unsigned long long* sample_ptr, const int32_t* base, uint32_t n_samples,
                             int32_t root_extra, int32_t* gbest) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_samples) gbest[s] = (int32_t)(sample_ptr[s + 1] - sample_ptr[s]) + root_extra + 1 - base[s];
}

}  // namespace ub200
