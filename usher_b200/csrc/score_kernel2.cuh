// k_score2 — block-parallel scoring kernel (best-placement mode).  See DESIGN.md "Kernels".
//
// Same contract and data as k_score<kModeBest> (score_kernel.cuh): a persistent launch scores every node of the
// tree against NG groups of 32 samples; a warp streams contiguous DFS tiles through per-warp shared-memory
// rings filled by cp.async.bulk (UBLKCP) + mbarriers.  What changes is WHO does the work inside a warp.
// k_score ran everything with lane = sample, so all sample-independent work (row scan, header decode, stack
// bookkeeping) and every hit (which concerns ~1 of the 32 samples) was executed 32-wide for nothing.  Here a
// warp takes an aligned block of 32 DFS nodes at a time and switches the lane role per phase:
//   A  lane = node      headers -> registers/smem, row offsets by a warp scan
//   B  lane = mutation  4 mutations per lane per step against the group's position bitmap; hits are compacted
//                       into a shared-memory list with ballot/popc
//   C  lane = hit       (node by binary search in the block's row offsets) x (samples calling the position, from
//                       the position's 32-byte table row) -> packed (dcorr, da, dcommon) from a 1024-entry LUT,
//                       accumulated with shared-memory atomics into dnode[node][sample]
//   D  lane = node      each node's "value source" (own materialised row / inherited row / stack level) by
//                       pointer jumping; then lane = sample materialises the running correction of the INTERNAL
//                       nodes that had hits (the only rows any child can inherit from)
//   E  lane = sample    exact lower bound of every non-hit pair of the block against the running best
//                       (local and cross-warp global); only samples that can still improve or tie are
//                       evaluated, lane = node
//   F  lane = sample    the hit pairs of each sample, exactly (score, validity, tie key)
//   G  lane = sample    stack levels the following blocks can inherit from
// All pruning is exact: a pair is skipped only when its score is provably greater than the final best.
#pragma once
#include "score_kernel.cuh"

namespace ub200 {

constexpr int kWarps2 = 16;
constexpr int kThreads2 = kWarps2 * 32;
constexpr uint32_t kHitCap = 192;
constexpr uint32_t kMutChunk2 = 256;                      // words per bulk copy in this kernel (1 KB)
constexpr int kMutStages2 = 2;
constexpr uint32_t kMutRing2 = kMutChunk2 * kMutStages2;  // 512 words = 2 KB
// per-warp shared memory (bytes)
constexpr uint32_t kOffMring = 0;       // u32[512]
constexpr uint32_t kOffHring = 2048;    // uint4[64]  (2 stages)
constexpr uint32_t kOffDnode = 3072;    // i32[32][32] packed deltas
constexpr uint32_t kOffVals = 7168;     // i16[66][32]: rows 0..31 materialised block rows, 32..63 stack levels
                                        //              0..31, 64 = the all-zero row
constexpr uint32_t kOffHit = 11392;     // uint2[192]
constexpr uint32_t kOffInfo = 12928;    // per-block node info (see kInfo* below)
constexpr uint32_t kOffBars = 13888;    // mbarriers
constexpr uint32_t kWarpSmem2 = 13952;
constexpr int kHdrStages2 = 2;
constexpr uint32_t kInfoG = 0, kInfoTie = 32, kInfoMisc = 64, kInfoNc0 = 96, kInfoPsrc = 128, kInfoRs = 160,
                   kInfoHm = 200, kInfoH = 232;
// value-source codes: < 65 = row of `vals`; >= kSrcSpill = stack level (code - kSrcSpill + 32) in HBM
constexpr uint32_t kRowStack = 32, kRowZero = 64, kSrcSpill = 128;
constexpr uint32_t kLutBytes = 4096;
constexpr uint32_t kMaxRowV2 = 500;     // packed 10-bit delta fields; longer rows take the k_score path
constexpr uint32_t kMaxCallsV2 = 32000; // |corr| <= calls per sample must fit int16

constexpr uint32_t kSrcPtr = 0x20000000u;   // unresolved: inherit from block lane (low 5 bits)
__device__ __forceinline__ uint32_t src_level(uint32_t level) {
    return level < (uint32_t)kStackDepth ? kRowStack + level : kSrcSpill + (level - kStackDepth);
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s16(uint32_t a) {
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
// stack levels beyond the 32 kept in shared memory live in HBM; out of line so the hot path stays short
__device__ __noinline__ int spill_read(const int32_t* gstk, uint32_t code, uint32_t s) {
    return gstk[(size_t)(code - kSrcSpill) * 32u + s];
}
__device__ __noinline__ void spill_write(int32_t* gstk, uint32_t level, uint32_t s, int v) {
    gstk[(size_t)(level - kStackDepth) * 32u + s] = v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

__device__ __forceinline__ int lut_delta(uint32_t i) {
    const uint32_t e = i >> 6, refc = (i >> 4) & 3u, prevc = (i >> 2) & 3u, mutc = i & 3u;
    const int rm = (mutc != refc), rp = (prevc != refc);
    const int wm = (e >> mutc) & 1u, wp = (e >> prevc) & 1u;
    const int dcorr = (wm - wp) - (rm - rp);
    const int tk = wm ^ 1, t0 = rm ^ 1;
    const int da = (tk & wp) - (t0 & rp);
    const int dcom = tk - t0;
    return dcorr * (1 << 20) + da * (1 << 10) + dcom;
}
__device__ __forceinline__ void unpack_delta(int v, int& dcorr, int& da, int& dcom) {
    dcom = (int)((uint32_t)v << 22) >> 22;
    const int v1 = (v - dcom) >> 10;
    da = (int)((uint32_t)v1 << 22) >> 22;
    dcorr = (v1 - da) >> 10;
}

// COLLECT = false: best placement per sample.  COLLECT = true: second pass that lists every optimal node of each
// sample (best_j_vec + node_has_unique); the final best score is the bound, so almost every block is pruned.
template <bool SMEM_BITMAP, bool COLLECT>
__global__ void __launch_bounds__(kThreads2, 1) k_score2(const ScoreParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.x % p.ngroups;
    const uint32_t cta_in_group = blockIdx.x / p.ngroups;
    const uint32_t ctas_per_group = gridDim.x / p.ngroups;
    const uint32_t ggroup = p.group0 + group;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- shared memory carve-up: [bitmap][lut][warp 0 .. warp 11]
    uint32_t* bm_s = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bm_bytes = SMEM_BITMAP ? ((p.bitmap_words * 4u + 127u) & ~127u) : 0u;
    int* lut = reinterpret_cast<int*>(smem + bm_bytes);
    uint8_t* wbase = smem + bm_bytes + kLutBytes + warp * kWarpSmem2;
    uint32_t* mring = reinterpret_cast<uint32_t*>(wbase + kOffMring);
    uint4* hring = reinterpret_cast<uint4*>(wbase + kOffHring);
    int* dnode = reinterpret_cast<int*>(wbase + kOffDnode);
    int16_t* vals = reinterpret_cast<int16_t*>(wbase + kOffVals);
    uint2* hitbuf = reinterpret_cast<uint2*>(wbase + kOffHit);
    uint32_t* info = reinterpret_cast<uint32_t*>(wbase + kOffInfo);
    const uint32_t mring_a = smem_u32(mring), hring_a = smem_u32(hring), bars_a = smem_u32(wbase + kOffBars);
    const uint32_t bm_a = smem_u32(bm_s);

    const uint32_t* bm_g = p.bitmap + (size_t)ggroup * p.bitmap_words;
    if (SMEM_BITMAP) {
        const uint4* src = reinterpret_cast<const uint4*>(bm_g);
        uint4* dst = reinterpret_cast<uint4*>(bm_s);
        for (uint32_t i = threadIdx.x; i < p.bitmap_words / 4; i += kThreads2) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = threadIdx.x; i < 1024; i += kThreads2) lut[i] = lut_delta(i);
    // the ring must only ever hold valid mutation words (lanes past a row's end still index the bitmap)
    for (uint32_t i = lane; i < kMutRing2; i += 32) mring[i] = 0u;
    vals[kRowZero * 32u + lane] = 0;
    if (lane == 0) {
        for (int i = 0; i < kMutStages2 + kHdrStages2; i++) mbar_init(bars_a + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t* tabg = p.tab + (size_t)ggroup * p.L * 8u;
    const uint32_t wig = cta_in_group * kWarps2 + warp;
    const uint32_t wpg = ctas_per_group * kWarps2;
    int32_t* gstk = p.gstack ? p.gstack + ((size_t)(blockIdx.x * kWarps2 + warp) * p.gstack_levels) * 32u : nullptr;
    const uint32_t sample = ggroup * 32u + lane;
    const bool live = sample < p.n_samples;

    const uint32_t vals_a = smem_u32(vals);
    auto value_of = [&](uint32_t code, uint32_t s) -> int {
        if (__builtin_expect(code >= kSrcSpill, 0)) return spill_read(gstk, code, s);
        return lds_s16(vals_a + ((code * 32u + s) << 1));
    };
    auto stack_read = [&](uint32_t level, uint32_t s) -> int { return value_of(src_level(level), s); };
    auto stack_write = [&](uint32_t level, uint32_t s, int v) {
        if (__builtin_expect(level >= (uint32_t)kStackDepth, 0)) spill_write(gstk, level, s, v);
        else vals[(kRowStack + level) * 32u + s] = (int16_t)v;
    };

    // per-lane (= sample) running best (COLLECT: the known final best, fixed)
    int bsc = COLLECT ? (live ? p.target_rel[sample] : (int)0x80000000) : 0x7fffffff;
    unsigned long long bkey = ~0ull;
    uint32_t cnt = 0;
    auto merge = [&](int sc, uint32_t tiekey, uint32_t hu, uint32_t node) {
        if (COLLECT) {
            if (sc == bsc) {
                const uint32_t k = atomicAdd(p.set_fill + sample, 1u);
                p.set_out[p.set_ptr[sample] + k] = node | (hu ? 0x80000000u : 0u);
            }
            return;
        }
        const unsigned long long key =
            ((unsigned long long)(uint32_t)(sc + kScoreBias) << 33) | ((unsigned long long)tiekey << 1) | hu;
        if (sc < bsc) { bsc = sc; cnt = 1; bkey = key; }
        else if (sc == bsc) { cnt++; if (key < bkey) bkey = key; }
    };

    // hits of the list [0, H) -> dnode / hm.  Entry = (mutation word, y): y bit31 set -> low bits are the block
    // lane of the node; else y is the absolute mutation index, resolved through the block's row offsets.
    auto process_hits = [&](uint32_t H) {
        for (uint32_t k0 = 0; k0 < H; k0 += 32) {
            const uint32_t k = k0 + lane;
            if (k < H) {
                const uint2 hv = hitbuf[k];
                uint32_t nl;
                if (hv.y & 0x80000000u) {
                    nl = hv.y & 31u;
                } else {
                    nl = 0;
#pragma unroll
                    for (uint32_t st = 16; st >= 1; st >>= 1)
                        if (info[kInfoRs + nl + st] <= hv.y) nl += st;
                }
                const uint32_t m = hv.x;
                const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(tabg + (size_t)(m >> 6) * 8u));
                const uint32_t r4 = __ldg(tabg + (size_t)(m >> 6) * 8u + 4);
                uint32_t pm = r0.x;
                while (pm) {
                    const uint32_t s = __ffs(pm) - 1;
                    pm &= pm - 1;
                    const uint32_t wsel = s >> 3;
                    const uint32_t nw = wsel == 0 ? r0.y : wsel == 1 ? r0.z : wsel == 2 ? r0.w : r4;
                    const uint32_t e = (nw >> ((s & 7u) * 4u)) & 15u;
                    atomicAdd(&dnode[nl * 32u + s], lut[(e << 6) | (m & 63u)]);
                    atomicOr(&info[kInfoHm + s], 1u << nl);
                }
            }
        }
    };

    uint32_t mphase = 0, hphase = 0;

    (void)wig; (void)wpg;
    for (;;) {
        // tiles are handed out in DFS order by a per-group counter: balances uneven tiles, and the CTAs of
        // different groups still walk the tree in the same order (one HBM read, the rest from L2)
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(p.tile_counter + group, 1u);
        t = __shfl_sync(FULL, t, 0);
        if (t >= p.n_tiles) break;
        const uint32_t n0 = p.tile_start[t], n1 = p.tile_start[t + 1];
        const uint32_t ms = p.row32[n0], me = p.row32[n1];
        uint32_t mc_issue = ms / kMutChunk2;
        const uint32_t mc_end = (me > ms) ? (me - 1) / kMutChunk2 + 1 : mc_issue;
        uint32_t mc_wait = mc_issue;
        uint32_t hc_issue = n0 / kHdrChunk;
        const uint32_t hc_end = (n1 - 1) / kHdrChunk + 1;
        if (lane == 0) {
            for (int i = 0; i < kMutStages2 && mc_issue + i < mc_end; i++) {
                const uint32_t c = mc_issue + i, s = c % kMutStages2;
                mbar_expect_tx(bars_a + 8 * s, kMutChunk2 * 4);
                bulk_g2s(mring_a + s * kMutChunk2 * 4, p.mutw + (size_t)c * kMutChunk2, kMutChunk2 * 4, bars_a + 8 * s);
            }
            for (int i = 0; i < kHdrStages2 && hc_issue + i < hc_end; i++) {
                const uint32_t c = hc_issue + i, s = c % kHdrStages2;
                mbar_expect_tx(bars_a + 8 * (kMutStages2 + s), kHdrChunk * 16);
                bulk_g2s(hring_a + s * kHdrChunk * 16, p.hdr + (size_t)c * kHdrChunk, kHdrChunk * 16,
                         bars_a + 8 * (kMutStages2 + s));
            }
        }
        mc_issue = min(mc_issue + kMutStages2, mc_end);
        hc_issue = min(hc_issue + kHdrStages2, hc_end);

        // cross-warp bound of this lane's sample, and the tile-local floor of every value the tile can reference
        int gb = COLLECT ? bsc : (live ? *(volatile int*)(p.gbest + sample) : 0x7fffffff);
        int gmin = 0;

        // ================= seed: running corrections of the tile's root path (levels 0 .. depth-1) =================
        {
            const uint32_t a0 = p.anc_ptr[t], a1 = p.anc_ptr[t + 1];
            for (uint32_t c0i = a0; c0i < a1; c0i += 32) {
                const uint32_t cn = min(32u, a1 - c0i);
#pragma unroll
                for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(dnode)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
                info[kInfoHm + lane] = 0;
                __syncwarp();
                uint32_t H = 0;
                for (uint32_t j = 0; j < cn; j++) {
                    const uint32_t a = p.anc[c0i + j];
                    const uint32_t r0 = p.row32[a], r1 = p.row32[a + 1];
                    for (uint32_t i = r0; i < r1; i += 32) {
                        const bool in = (i + lane) < r1;
                        const uint32_t m = in ? __ldg(p.mutw + i + lane) : 0u;
                        const bool hit = in && bitmap_test<SMEM_BITMAP>(bm_s, bm_g, m >> 6);
                        const uint32_t bal = __ballot_sync(FULL, hit);
                        if (hit) hitbuf[H + __popc(bal & lt_mask)] = make_uint2(m, 0x80000000u | j);
                        H += __popc(bal);
                        if (H > kHitCap - 32) { __syncwarp(); process_hits(H); H = 0; __syncwarp(); }
                    }
                }
                __syncwarp();
                if (H) process_hits(H);
                __syncwarp();
                for (uint32_t j = 0; j < cn; j++) {
                    const uint32_t lvl = (c0i - a0) + j;   // the chain is root-first: index == level
                    const int v = dnode[j * 32u + lane];
                    const int c = (lvl ? stack_read(lvl - 1, lane) : 0) + ((v + (1 << 19)) >> 20);
                    stack_write(lvl, lane, c);
                    gmin = min(gmin, c);
                }
                __syncwarp();
            }
        }

        uint32_t rs_run = ms;
        for (uint32_t blk = n0 & ~31u; blk < n1; blk += 32) {
            const uint32_t b0 = max(blk, n0), b1 = min(blk + 32u, n1);
            // ================= A: headers (lane = node) =================
            {
                const uint32_t hc = blk / kHdrChunk, s = hc % kHdrStages2;
                mbar_wait(bars_a + 8 * (kMutStages2 + s), (hphase >> s) & 1u);
                hphase ^= 1u << s;
                if (hc_issue < hc_end && hc_issue < hc + kHdrStages2) {
                    // the previous block's header stage was consumed before its __syncwarp()s
                    if (lane == 0) {
                        const uint32_t c = hc_issue, s2 = c % kHdrStages2;
                        mbar_expect_tx(bars_a + 8 * (kMutStages2 + s2), kHdrChunk * 16);
                        bulk_g2s(hring_a + s2 * kHdrChunk * 16, p.hdr + (size_t)c * kHdrChunk, kHdrChunk * 16,
                                 bars_a + 8 * (kMutStages2 + s2));
                    }
                    hc_issue++;
                }
            }
            const uint4 h = hring[(blk % (kHdrChunk * kHdrStages2)) + lane];
            const uint32_t node = blk + lane;
            const bool act = node >= b0 && node < b1;
            const uint32_t level = hdr_level(h.z), plane = hdr_plane(h.z), flags = hdr_flags(h.z);
            const uint32_t nmut = act ? (h.w >> 16) : 0u;
            const bool root = act && (flags & kFlagRoot);
            const bool leaf = flags & kFlagLeaf;
            uint32_t incl = nmut;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            const uint32_t reb = rs_run + __shfl_sync(FULL, incl, 31);
            info[kInfoRs + lane] = rs_run + incl - nmut;
            if (lane == 31) info[kInfoRs + 32] = reb;
            info[kInfoG + lane] = (uint32_t)h.x;
            info[kInfoTie + lane] = h.y;
            info[kInfoMisc + lane] = h.z;
            info[kInfoNc0 + lane] = h.w;
            info[kInfoHm + lane] = 0;
            if (lane == 0) info[kInfoH] = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(dnode)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
            __syncwarp();

            // ================= B + C: scan the block's mutations, accumulate the hits =================
            if (reb != rs_run) {
                // One step = one aligned 128-word chunk of the stream = one ring stage (4 consecutive words per lane,
                // one LDS.128).  Chunks cut by the block's ends are masked; a chunk shared by two blocks is scanned
                // by both.  Hits are appended to the list with a shared-memory atomic.
                const uint32_t c_first = rs_run / kMutChunk2, c_last = (reb - 1u) / kMutChunk2;
                for (uint32_t c = c_first; c <= c_last; c++) {
                    if (c >= mc_wait) {
                        const uint32_t s = c % kMutStages2;
                        mbar_wait(bars_a + 8 * s, (mphase >> s) & 1u);
                        mphase ^= 1u << s;
                        mc_wait = c + 1;
                    }
                    // chunks below c are dead (only this block needed them): refill their stages
                    while (mc_issue < mc_end && mc_issue < c + kMutStages2) {
                        if (lane == 0) {
                            const uint32_t cc = mc_issue, s = cc % kMutStages2;
                            mbar_expect_tx(bars_a + 8 * s, kMutChunk2 * 4);
                            bulk_g2s(mring_a + s * kMutChunk2 * 4, p.mutw + (size_t)cc * kMutChunk2, kMutChunk2 * 4,
                                     bars_a + 8 * s);
                        }
                        mc_issue++;
                    }
                    const bool edge = (c == c_first || c == c_last);   // warp-uniform
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const uint32_t base = c * kMutChunk2 + half * 128u + 4u * lane;
                        const uint4 q = lds128(mring_a + (((c % kMutStages2) * kMutChunk2 + half * 128u + 4u * lane) << 2));
                        const uint32_t w0 = SMEM_BITMAP ? lds32(bm_a + ((q.x >> 11) << 2)) : __ldg(bm_g + (q.x >> 11));
                        const uint32_t w1 = SMEM_BITMAP ? lds32(bm_a + ((q.y >> 11) << 2)) : __ldg(bm_g + (q.y >> 11));
                        const uint32_t w2 = SMEM_BITMAP ? lds32(bm_a + ((q.z >> 11) << 2)) : __ldg(bm_g + (q.z >> 11));
                        const uint32_t w3 = SMEM_BITMAP ? lds32(bm_a + ((q.w >> 11) << 2)) : __ldg(bm_g + (q.w >> 11));
                        uint32_t hb = ((w0 >> ((q.x >> 6) & 31u)) & 1u) | (((w1 >> ((q.y >> 6) & 31u)) & 1u) << 1) |
                                      (((w2 >> ((q.z >> 6) & 31u)) & 1u) << 2) | (((w3 >> ((q.w >> 6) & 31u)) & 1u) << 3);
                        if (edge) {   // mask the words outside [rs_run, reb)
                            const uint32_t span = reb - rs_run, o = base - rs_run;
                            hb &= (o < span ? 1u : 0u) | (o + 1u < span ? 2u : 0u) | (o + 2u < span ? 4u : 0u) |
                                  (o + 3u < span ? 8u : 0u);
                        }
                        if (hb) {
                            const uint32_t slot = atomicAdd(&info[kInfoH], (uint32_t)__popc(hb));
                            uint32_t k = slot;
                            if (hb & 1u) hitbuf[k++] = make_uint2(q.x, base);
                            if (hb & 2u) hitbuf[k++] = make_uint2(q.y, base + 1u);
                            if (hb & 4u) hitbuf[k++] = make_uint2(q.z, base + 2u);
                            if (hb & 8u) hitbuf[k++] = make_uint2(q.w, base + 3u);
                        }
                        if (half == 0) {
                            __syncwarp();
                            if (info[kInfoH] > kHitCap - 128) {
                                process_hits(info[kInfoH]);
                                __syncwarp();
                                if (lane == 0) info[kInfoH] = 0;
                                __syncwarp();
                            }
                        }
                    }
                    __syncwarp();
                    const uint32_t H = info[kInfoH];
                    if (H > kHitCap - 128) {
                        process_hits(H);
                        __syncwarp();
                        if (lane == 0) info[kInfoH] = 0;
                        __syncwarp();
                    }
                }
                __syncwarp();
                const uint32_t H = info[kInfoH];
                if (H) process_hits(H);
                __syncwarp();
            }
            rs_run = reb;

            // ================= D: value sources (lane = node), materialise internal hit rows (lane = sample) ======
            const uint32_t hmv = info[kInfoHm + lane];                   // lane = sample: its hit nodes
            const uint32_t hitnodes = __reduce_or_sync(FULL, hmv);
            const bool hitn = (hitnodes >> lane) & 1u;                    // lane = node
            const bool par_in = act && plane != 0 && (blk + plane - 1u) >= b0;
            const uint32_t pl = (plane - 1u) & 31u;
            uint32_t own;
            if (!act) own = kRowZero;
            else if (hitn && !leaf) own = lane;                          // own materialised row
            else if (root) own = kRowZero;
            else if (par_in) own = kSrcPtr | pl;
            else own = src_level(level - 1u);
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t o2 = __shfl_sync(FULL, own, (own & kSrcPtr) ? (own & 31u) : lane);
                if (own & kSrcPtr) own = o2;
            }
            const uint32_t own_pl = __shfl_sync(FULL, own, pl);
            const uint32_t psrc = (root || !act) ? kRowZero : (par_in ? own_pl : src_level(level - 1u));
            info[kInfoPsrc + lane] = psrc;
            {
                uint32_t mm = __ballot_sync(FULL, act && hitn && !leaf);
                while (mm) {
                    const uint32_t n = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const uint32_t ps = __shfl_sync(FULL, psrc, n);
                    const int v = dnode[n * 32u + lane];
                    const int c = value_of(ps, lane) + ((v + (1 << 19)) >> 20);
                    vals[n * 32u + lane] = (int16_t)c;
                    gmin = min(gmin, c);
                }
            }
            __syncwarp();

            // ================= E: non-hit pairs, pruned by an exact lower bound =================
            {
                const bool dense_ok = act && (flags & kFlagValid0);
                const int gv = dense_ok ? h.x : 0x3fffffff;
                const int min_g = __reduce_min_sync(FULL, gv);
                const int bound = min(bsc, gb);
                uint32_t need = __ballot_sync(FULL, live && min_g < 0x3fffffff && min_g + gmin <= bound);
                while (need) {
                    const uint32_t s = __ffs(need) - 1;
                    need &= need - 1;
                    const uint32_t hm_s = info[kInfoHm + s];
                    const int sc = h.x + value_of(psrc, s);
                    const int bs = __shfl_sync(FULL, bsc, s);
                    uint32_t cm = __ballot_sync(FULL, dense_ok && !((hm_s >> lane) & 1u) && sc <= bs);
                    while (cm) {
                        const uint32_t j = __ffs(cm) - 1;
                        cm &= cm - 1;
                        const int scj = __shfl_sync(FULL, sc, j);
                        const uint32_t tkj = __shfl_sync(FULL, h.y, j);
                        const uint32_t huj = __shfl_sync(FULL, (flags & kFlagHu0) ? 1u : 0u, j);
                        if (lane == s) merge(scj, tkj, huj, blk + j);
                    }
                }
            }

            // ================= F: hit pairs, exact (lane = sample) =================
            {
                uint32_t hmw = hmv;
                while (hmw) {
                    const uint32_t n = __ffs(hmw) - 1;
                    hmw &= hmw - 1;
                    int dcorr, da, dcom;
                    unpack_delta(dnode[n * 32u + lane], dcorr, da, dcom);
                    const uint32_t z = info[kInfoMisc + n], w = info[kInfoNc0 + n];
                    const uint32_t fl = hdr_flags(z);
                    const int g = (int)info[kInfoG + n];
                    int sc;
                    bool valid;
                    uint32_t hu;
                    if (fl & kFlagRoot) {
                        sc = g + dcorr; valid = true; hu = 0;
                    } else {
                        const bool masked = fl & kFlagMasked;
                        if (masked) { da = 0; dcom = 0; }
                        sc = g + value_of(info[kInfoPsrc + n], lane) - da;
                        const int common = (int)(w & 0xffffu) + dcom;
                        hu = (masked || (int)(w >> 16) > common) ? 1u : 0u;
                        valid = (fl & kFlagLeaf) ? common > 0 : (!hu || common > 0);
                    }
                    if (valid && sc <= bsc) merge(sc, info[kInfoTie + n], hu, blk + n);
                }
            }

            // ================= G: stack levels later blocks can inherit from (last internal node per level) ======
            __syncwarp();   // E/F read rows of `vals` with lane = node / per-lane loops; G overwrites stack rows
            {
                const bool isint = act && !leaf;
                const uint32_t peers = __match_any_sync(FULL, isint ? level : (0xffff0000u | lane));
                uint32_t wm = __ballot_sync(FULL, isint && lane == 31u - __clz(peers));
                while (wm) {
                    const uint32_t n = __ffs(wm) - 1;
                    wm &= wm - 1;
                    const uint32_t o = __shfl_sync(FULL, own, n);
                    const uint32_t lv = __shfl_sync(FULL, level, n);
                    const int v = value_of(o, lane);
                    stack_write(lv, lane, v);
                    gmin = min(gmin, v);
                }
            }
            __syncwarp();
        }
        // publish an improved bound for the other warps working on this sample group
        if (!COLLECT && live && bsc < gb) atomicMin(p.gbest + sample, bsc);
        __syncwarp();
    }

    if (COLLECT) return;
    // fold the CTA's warps in shared memory (the rings are dead now), one partial row per CTA
    __syncthreads();
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(smem);
    uint32_t* scnt = reinterpret_cast<uint32_t*>(smem + kWarps2 * 32 * 8);
    skey[warp * 32 + lane] = bkey;
    scnt[warp * 32 + lane] = cnt;
    __syncthreads();
    if (warp == 0) {
        unsigned long long best = ~0ull;
        for (int w = 0; w < kWarps2; w++) best = min(best, skey[w * 32 + lane]);
        uint32_t c = 0;
        for (int w = 0; w < kWarps2; w++)
            if ((skey[w * 32 + lane] >> 33) == (best >> 33)) c += scnt[w * 32 + lane];
        const size_t o = ((size_t)group * ctas_per_group + cta_in_group) * 32u + lane;
        p.part_key[o] = best;
        p.part_cnt[o] = c;
    }
}

}  // namespace ub200
