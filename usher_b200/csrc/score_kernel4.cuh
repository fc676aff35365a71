// k_score4 — streaming best-placement kernel on the segment layout (ub200_internal.h, DESIGN.md "Kernels").
//
// One persistent launch scores every node of the tree against `nsg` SCAN GROUPS; a scan group is NC consecutive
// groups of 32 samples that share one pass over the mutation stream.  The unit of work is a tile: a contiguous DFS
// range of whole 32-node blocks whose mutations are ONE contiguous piece of the stream, [seed segments][block
// segments].  A CTA holds kUnits UNITS of 1 + NC warps; a unit works on one tile at a time:
//
//   scanner warp   pulls the tile's stream through a 4 KB shared-memory ring of two 2 KB stages: a STEP = 4 rows of
//                  128 words = ONE bulk async copy (TMA unit, cp.async.bulk, completion on the stage's mbarrier)
//                  issued by an elected lane a whole step ahead.  Steps are aligned to the tile, not to its segments:
//                  every word is loaded and tested exactly once against the scan group's UNION bitmap (per word:
//                  byte offset of the bitmap word, LDS, a wrap shift and a funnel shift collecting the hit bits).
//                  The hit bits of a step are then handed out segment by segment: masked to the segment's quads,
//                  compacted with two ballots and copied into one of eight 64-word message slots that all NC
//                  consumers read (full barrier: 1 arrival, empty barrier: NC arrivals).  Sparse single-group steps
//                  re-read the few hit words from the ring; shared scans and dense steps store them from the
//                  registers they were tested in, into the slots used as one circular buffer.  A segment = one or
//                  more messages, the last one flagged.  The scanner needs no sample state.
//   consumer warp  owns the state of ONE group of 32 samples.  Per block:
//     rec            16-byte block record fetched a block ahead (min(G - nmut), open-chain mask and level): the 32
//                    node headers are only read for blocks that can still hold an optimum
//     C  lane = hit  table row of the position (32 B, L2; two rows per lane in flight).  Rows whose sample mask is
//                    empty belong to another group of the scan group and are dropped.  Per 32 hits, by cost: the
//                    SPARSE form (lane = hit walks the samples that call its position: packed (dcorr, da, dcommon)
//                    from a 1024-entry LUT, shared-memory atomics into dnode[node][sample], neg[sample] +=
//                    min(dcorr, 0)) or the DENSE form (32 x 32 bit transpose with 5 shuffles, then lane = SAMPLE
//                    walks its own hits, owns its dnode column: no atomics, no bank conflicts, neg in a register)
//     bound          exact lower bound of every pair of the block:  min(G - nmut) + gmin + neg  against the running
//                    best; blocks that cannot hold an optimum skip E and F entirely
//     E  lane = node     non-hit pairs of a sample that can still improve or tie
//     F  lane = sample   hit pairs, exactly (score, validity, tie key)
//     G  lane = sample   path corrections of the block's OPEN chain -> stack rows the following blocks read
//   Seed segments (the rows of the tile's root path, 32 levels per segment) go through the same scanner ->
//   consumer path and initialise the stack.
// The correction of the path above a node inside its own block is never materialised: it is
// stack[level above the block] + sum of dnode over (in-block ancestors & hit nodes of the sample).  All pruning is
// exact, so results are schedule-independent.
#pragma once
#include "score_kernel.cuh"

namespace ub200 {

constexpr uint32_t kRingWords4 = 1024;                     // two stages of one step (4 rows of 128 words) = 4 KB
constexpr uint32_t kSlots4 = 8, kSlotCap4 = 64;            // scanner -> consumer messages
constexpr uint32_t kAreaWords4 = 192;                      // per-consumer scratch (dense-form staging: 160, header copies: 160)
constexpr uint32_t kLut4Bytes = 4096;
constexpr uint32_t kMaxRowV4 = 500;      // packed 10-bit delta fields
constexpr uint32_t kMaxCallsV4 = 16383;  // path corrections are int16 and reach -2 per call
constexpr uint32_t kSmemLimit4 = 232448; // 227 KB
constexpr uint32_t kMsgLast4 = 1u << 16, kMsgTile4 = 2u << 16, kMsgEnd4 = 4u << 16;

// Geometry of a CTA for NC consumers per scanner.
template <int NC>
struct Cfg4 {
    static constexpr int kUnits = NC == 1 ? 16 : (NC == 2 ? 10 : 8);
    static constexpr int kWarps = kUnits * (1 + NC);
    static constexpr int kThreads = kWarps * 32;
    static constexpr int kStack = NC == 3 ? 32 : 40;        // stack levels kept in shared memory (deeper: HBM spill)
    // unit-shared part (bytes): ring, hit lists, message words, barriers
    static constexpr uint32_t kORing = 0;                   // u32[1024]
    static constexpr uint32_t kOList = 4096;                // u32[8][64]
    static constexpr uint32_t kOMsg = kOList + kSlots4 * kSlotCap4 * 4;   // uint2[8]
    static constexpr uint32_t kOBars = kOMsg + kSlots4 * 8;               // 8 full + 8 empty mbarriers + 2 ring stages
    static constexpr uint32_t kShared = kOBars + 2 * kSlots4 * 8 + 16;
    // per-consumer part
    static constexpr uint32_t kODnode = 0;                  // i32[32][32] packed deltas
    static constexpr uint32_t kOStack = 4096;               // i16[kStack][32]
    static constexpr uint32_t kONeg = kOStack + kStack * 64;   // i32[32]
    static constexpr uint32_t kOArea = kONeg + 128;         // u32[192]: dense-form staging while hits arrive, header copies
                                                            // (G, Z, W, Am, Hm) while a block is evaluated
    static constexpr uint32_t kCons = kOArea + kAreaWords4 * 4;
    static constexpr uint32_t kUnit = (kShared + NC * kCons + 127) & ~127u;
    static constexpr uint32_t kFixed = kLut4Bytes + kUnits * kUnit;
    // the 3.75 KB position bitmap of a 30 kb genome has to fit beside the units (else every bitmap test goes to L1/L2)
    static_assert(kFixed + 3840 <= kSmemLimit4, "a 30 kb bitmap no longer fits shared memory");
};
constexpr uint32_t kA4G = 0, kA4Z = 32, kA4W = 64, kA4Am = 96, kA4Hm = 128;

struct Score4Params {
    const uint32_t* stream;
    const NodeHdr* hdr;           // hdr3
    const uint32_t* tiekey;
    const uint32_t* blk_words;    // [blocks] stream words of each 32-node block segment (multiple of 4)
    const uint4* blk_rec;         // [blocks] x = min(G - nmut), y = open-chain mask, z = level of the first open node
    const uint32_t* tile_start;   // [T+1]
    const uint32_t* tile_w0;      // [T+1]
    const uint32_t* tile_lvl;     // [T]
    const uint32_t* tile_sseg;    // [T+1]
    const uint32_t* seed_end;
    uint32_t n_nodes, n_tiles, L, bitmap_words;
    const uint32_t* bitmap;       // union bitmaps of this pass's scan groups: [nsg][bitmap_words]
    const uint32_t* tab;          // [groups][L][8]: mask, ref<<4, nibbles[4], -, -
    int32_t* gbest;
    uint32_t n_samples, group0, ngroups, nsg;   // first group / number of groups / scan groups of this pass
    unsigned long long* part_key;
    uint32_t* part_cnt;
    uint32_t part_group0, part_stride;   // partial row of (group g of the pass, CTA c) = (part_group0 + g) * part_stride + c
    int32_t* gstack;
    uint32_t gstack_levels;
    const int32_t* target_rel;
    uint32_t* set_out;
    const unsigned long long* set_ptr;
    uint32_t* set_fill;
    int32_t* tile_min;            // [groups of the batch][T][32] or NULL.  MODE 0 leaves in it the smallest score a
                                  // sample's candidates reached inside a tile; MODE 1 then only scans the tiles whose
                                  // minimum equals the sample's final best (every optimal node was such a candidate)
    const int32_t* base;          // MODE 2: per sample, the calls that already disagree with the bare reference
    int32_t* node_scores;         // MODE 2: [n_samples][n_nodes]
    uint32_t* tile_counter;
    unsigned long long* prof;     // UB200_PROFILE builds only: cycle counters per role (api.cu ub200_debug_prof)
};

// Developer instrumentation (compiled in with -DUB200_PROFILE): where the scanner and consumer warps spend their cycles.
#ifdef UB200_PROFILE
#define PROF_T0(v) const long long v = clock64()
#define PROF_ADD(slot, v) pc[slot] += (unsigned long long)(clock64() - (v))
#define PROF_INC(slot, n) pc[slot] += (n)
#else
#define PROF_T0(v)
#define PROF_ADD(slot, v)
#define PROF_INC(slot, n)
#endif

__device__ __forceinline__ uint32_t lds32_4(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32_4(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128_4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Waits on another warp.  mbarrier.try_wait does not suspend on this part (measured 21 ns per failed poll with or
// without a time hint, scripts/ubench/sleep_probe.cu) while nanosleep sleeps what it is told, so a warp that finds its
// barrier closed backs off exponentially: a waiting warp then costs a handful of issue slots and shared-memory
// wavefronts per microsecond instead of fifty.
__device__ __forceinline__ uint32_t mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t cap_ns) {
    if (mbar_try_wait(bar, parity)) return 0u;
    uint32_t ns = 32, spins = 0;
    do {
        __nanosleep(ns);
        ns = min(ns * 2u, cap_ns);
        if (++spins > (1u << 21)) __trap();   // a protocol bug must trap, never hang the GPU
    } while (!mbar_try_wait(bar, parity));
    return spins;   // failed polls (instrumentation)
}
// one 32-byte table row per lane: a single 256-bit load (one L1 wavefront per row instead of two)
__device__ __forceinline__ void ldg_row(const uint32_t* row, uint4& r0, uint4& r1) {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0.x), "=r"(r0.y), "=r"(r0.z), "=r"(r0.w), "=r"(r1.x), "=r"(r1.y), "=r"(r1.z), "=r"(r1.w)
                 : "l"(row));
}
// row `lane` of a 32x32 bit matrix in, row `lane` of its transpose out (recursive block transpose, 5 shuffles)
__device__ __forceinline__ uint32_t transpose32(uint32_t x, uint32_t lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & (uint32_t)j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}
__device__ __noinline__ int spill_read4(const int32_t* gstk, uint32_t level, uint32_t s) {
    return gstk[(size_t)level * 32u + s];
}
__device__ __noinline__ void spill_write4(int32_t* gstk, uint32_t level, uint32_t s, int v) {
    gstk[(size_t)level * 32u + s] = v;
}
__device__ __forceinline__ int lut_delta4(uint32_t i) {
    const uint32_t e = i >> 6, refc = (i >> 4) & 3u, prevc = (i >> 2) & 3u, mutc = i & 3u;
    const int rm = (mutc != refc), rp = (prevc != refc);
    const int wm = (e >> mutc) & 1u, wp = (e >> prevc) & 1u;
    const int dcorr = (wm - wp) - (rm - rp);
    const int tk = wm ^ 1, t0 = rm ^ 1;
    const int da = (tk & wp) - (t0 & rp);
    const int dcom = tk - t0;
    return dcorr * (1 << 20) + da * (1 << 10) + dcom;
}
// bitmap word of a stream word's position: narrow words carry the byte offset at bit 14.  The shared-memory form
// is an explicit ld.shared on the 32-bit window address (a generic load here costs a trip through the L1TEX pipe).
template <bool SMEM_BITMAP, bool NARROW>
__device__ __forceinline__ uint32_t bitmap_word(uint32_t bm_a, const uint32_t* bm_g, uint32_t w) {
    if (NARROW) {
        const uint32_t off = w >> 14;
        return SMEM_BITMAP ? lds32_4(bm_a + off)
                           : __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(bm_g) + off));
    }
    return SMEM_BITMAP ? lds32_4(bm_a + ((w >> 14) << 2)) : __ldg(bm_g + (w >> 14));
}
__device__ __forceinline__ int dc_of(int v) { return (v + (1 << 19)) >> 20; }
__device__ __forceinline__ void unpack_delta4(int v, int& dcorr, int& da, int& dcom) {
    dcom = (int)((uint32_t)v << 22) >> 22;
    const int v1 = (v - dcom) >> 10;
    da = (int)((uint32_t)v1 << 22) >> 22;
    dcorr = (v1 - da) >> 10;
}

// MODE 0: best placement per sample.  MODE 1 (collect): second pass that lists every optimal node of each sample
// (best_j_vec + node_has_unique); the final best score is the bound, so almost every block is pruned.  MODE 2: the
// reported score of EVERY node (`-p`, src/usher_common.cpp:557-578: score + 1 on invalid nodes): every block is
// evaluated, non-hit pairs are written lane = node (coalesced), hit pairs one by one.
constexpr int kMode4Best = 0, kMode4Collect = 1, kMode4NodeScores = 2, kMode4BestNotes = 3;   // 3 = MODE 0 + tile_min notes
template <int NC, bool SMEM_BITMAP, int MODE, bool NARROW>
__global__ void __launch_bounds__(Cfg4<NC>::kThreads, 1) k_score4(const Score4Params p) {
    using C = Cfg4<NC>;
    constexpr bool COLLECT = MODE == kMode4Collect, SCORES = MODE == kMode4NodeScores;
    constexpr bool BEST = MODE == kMode4Best || MODE == kMode4BestNotes, NOTES = MODE == kMode4BestNotes;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // unit / role of this warp; the scanners are spread over the four SM sub-partitions (warp % 4)
    uint32_t unit, role;   // role 0 = scanner, 1..NC = consumer role-1
    if (NC == 1) {
        unit = warp >> 1;
        role = ((warp & 1u) == ((warp >> 2) & 1u)) ? 0u : 1u;
    } else if (NC == 2) {
        unit = warp / 3u;
        role = warp % 3u;
    } else {
        unit = warp >> 2;
        role = ((warp & 3u) - (unit & 3u)) & 3u;
    }
    const uint32_t sg = blockIdx.x % p.nsg;
    const uint32_t cta_in_sg = blockIdx.x / p.nsg;
    const uint32_t ctas_per_sg = gridDim.x / p.nsg;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int BIG = 0x3fffffff;
#ifdef UB200_PROFILE
    unsigned long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_start = clock64();
#endif

    // ---- shared memory: [bitmap][lut][unit 0 .. unit kUnits-1]
    uint32_t* bm_s = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bm_bytes = SMEM_BITMAP ? ((p.bitmap_words * 4u + 127u) & ~127u) : 0u;
    int* lut = reinterpret_cast<int*>(smem + bm_bytes);
    uint8_t* ubase = smem + bm_bytes + kLut4Bytes + unit * C::kUnit;
    uint32_t* mring = reinterpret_cast<uint32_t*>(ubase + C::kORing);
    uint32_t* list = reinterpret_cast<uint32_t*>(ubase + C::kOList);
    volatile uint2* msg = reinterpret_cast<volatile uint2*>(ubase + C::kOMsg);
    const uint32_t mring_a = smem_u32(mring), bars_a = smem_u32(ubase + C::kOBars), list_a = smem_u32(list);
    const uint32_t bm_a = smem_u32(bm_s);
    constexpr uint32_t kBarFull = 0, kBarEmpty = kSlots4, kBarStage = 2 * kSlots4;

    const uint32_t* bm_g = p.bitmap + (size_t)sg * p.bitmap_words;
    if (SMEM_BITMAP) {
        const uint4* src = reinterpret_cast<const uint4*>(bm_g);
        uint4* dst = reinterpret_cast<uint4*>(bm_s);
        for (uint32_t i = threadIdx.x; i < p.bitmap_words / 4; i += C::kThreads) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = threadIdx.x; i < 1024; i += C::kThreads) lut[i] = lut_delta4(i);
    if (role == 0) {
        // lanes past a segment's end still index the bitmap with what the ring holds: only ever valid words
        for (uint32_t i = lane; i < kRingWords4; i += 32) mring[i] = 0u;
        // the zeros above are the only generic-proxy WRITES the ring ever sees: order them before the bulk copies once.
        // No proxy fence per step: a stage is only re-filled after every lane has consumed what it read from it
        // (measured: -0.8 % per launch at one group per scan, -2 % at three)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) {
            for (uint32_t i = 0; i < kSlots4; i++) {
                mbar_init(bars_a + 8 * (kBarFull + i), 1);
                mbar_init(bars_a + 8 * (kBarEmpty + i), NC);
            }
            mbar_init(bars_a + 8 * (kBarStage + 0), 1);
            mbar_init(bars_a + 8 * (kBarStage + 1), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    if (role == 0) {
        // =====================================================================================================
        // scanner
        // =====================================================================================================
        uint32_t rows_end = 0;   // end of the tile's rows
        uint32_t nmsg = 0;        // messages sent so far; the open one lives in slot nmsg % kSlots4
        uint32_t fill = 0;        // hit words in the open message
        auto open_msg = [&]() {   // wait until every consumer has released the slot
            const uint32_t s = nmsg % kSlots4;
            PROF_T0(tw);
            const uint32_t polls = mbar_wait_sleep(bars_a + 8 * (kBarEmpty + s), ((nmsg / kSlots4) & 1u) ^ 1u, 512u);
            PROF_ADD(1, tw); PROF_INC(2, polls); (void)polls;
            fill = 0;
        };
        auto send_msg = [&](uint32_t flags, uint32_t payload) {
            const uint32_t s = nmsg % kSlots4;
            __syncwarp();
            if (elect_one()) {
                msg[s].x = fill | flags;
                msg[s].y = payload;
                mbar_arrive(bars_a + 8 * (kBarFull + s));
            }
            nmsg++;
        };
        // ring loader: the ring is two stages of one step (4 rows = 2 KB) each; a step is ONE bulk async copy (TMA
        // unit, cp.async.bulk) that completes on the stage's mbarrier, issued by one elected lane a whole step ahead.
        // gstep counts the steps this scanner has started: stage = gstep & 1, barrier phase = (gstep >> 1) & 1.
        // A copy is issued exactly for the steps that will be processed (first row < rows_end: the tile's piece is
        // padded by less than half a step), so every barrier phase is consumed in order.
        uint32_t gstep = 0;
        auto stage_load = [&](uint32_t step_id, uint32_t first_row) {   // rows [first_row, first_row + 4) of the stream
            const uint32_t nrows = min(4u, rows_end - min(rows_end, first_row));
            __syncwarp();                                               // every lane is done reading this stage
            if (nrows && elect_one()) {
                const uint32_t st = step_id & 1u;
                mbar_expect_tx(bars_a + 8 * (kBarStage + st), nrows * 512u);
                bulk_g2s(mring_a + st * 2048u, p.stream + ((size_t)first_row << 7), nrows * 512u, bars_a + 8 * (kBarStage + st));
            }
        };
        // hand the hits of the current step that lie in stream words [off, lim) (multiples of 4, inside the step
        // starting at word `base`) to the open message; full messages are sent (not flagged last) and reopened
        uint4 q0, q1, q2, q3;   // the current step's four rows (16 words per lane)
        uint32_t stage_a = mring_a;   // shared-memory address of the current step's stage
        auto emit = [&](uint32_t hb_step, uint32_t base, uint32_t off, uint32_t lim) {
            const uint32_t idx = base + 4u * lane;
            uint32_t hb = hb_step;                            // bit 4k+j = word j of quad k (row k of the step)
            if (off != base || lim != base + 512u) {          // quads outside [off, lim) belong to other segments
                const int lo = (int)(off - idx), hi = (int)(lim - idx);   // multiples of 4
                const uint32_t k_lo = lo > 0 ? min((uint32_t)(lo + 127) >> 7, 4u) : 0u;
                const uint32_t k_hi = hi > 0 ? min((uint32_t)(hi + 127) >> 7, 4u) : 0u;
                hb &= ((1u << (4u * k_hi)) - 1u) & ~((1u << (4u * k_lo)) - 1u);
            }
            // compact: lane l's hits follow those of lanes < l.  Exclusive prefix of the per-lane counts:
            // two ballots when no lane has more than 3 hits (the usual case), else a shuffle scan
            const uint32_t c = __popc(hb);
            const uint32_t big = __ballot_sync(FULL, c > 3u);
            uint32_t excl, remaining;
            if (big == 0u) {
                const uint32_t v0 = __ballot_sync(FULL, c & 1u), v1 = __ballot_sync(FULL, c & 2u);
                if ((v0 | v1) == 0u) return;
                excl = __popc(v0 & lt_mask) + 2u * __popc(v1 & lt_mask);
                remaining = __popc(v0) + 2u * __popc(v1);
            } else {
                uint32_t incl = c;
#pragma unroll
                for (int dlt = 1; dlt < 32; dlt <<= 1) {
                    const uint32_t v = __shfl_up_sync(FULL, incl, dlt);
                    if (lane >= (uint32_t)dlt) incl += v;
                }
                remaining = __shfl_sync(FULL, incl, 31);
                excl = incl - c;
            }
            if (NC > 1 || big != 0u) {
                // Shared scans see 2-3 times the hits, and so do single-group scans of dense sample families (some lane
                // holds more than three hits): the hit words are stored straight from the registers they were
                // tested in (16 predicated stores, no ring re-read, no per-hit loop).  The eight 64-word slots are one
                // circular buffer of 512 words: hit p of this run goes to word (64 * open slot + fill + p) mod 512.
                // Hits that spill over into further messages first reserve every slot they need (a step holds at most
                // 512 hits = 8 slots); the full messages are sent after the copy, the last one stays open.
                uint32_t nslots = 1;
                if (remaining > kSlotCap4 - fill) {
                    if (fill + remaining > kSlots4 * kSlotCap4) {   // would need a ninth slot: flush the open message
                        send_msg(0u, 0u);
                        open_msg();
                    }
                    nslots = (fill + remaining + kSlotCap4 - 1u) / kSlotCap4;
                    for (uint32_t i = 1; i < nslots; i++) {
                        const uint32_t mi = nmsg + i;
                        mbar_wait_sleep(bars_a + 8 * (kBarEmpty + mi % kSlots4), ((mi / kSlots4) & 1u) ^ 1u, 512u);
                    }
                }
                uint32_t pi = (nmsg % kSlots4) * kSlotCap4 + fill + excl;
#define UB200_PUT(bit, val) \
    if (hb & (1u << (bit))) { sts32_4(list_a + ((pi & (kSlots4 * kSlotCap4 - 1u)) << 2), (val)); pi++; }
                UB200_PUT(0, q0.x) UB200_PUT(1, q0.y) UB200_PUT(2, q0.z) UB200_PUT(3, q0.w)
                UB200_PUT(4, q1.x) UB200_PUT(5, q1.y) UB200_PUT(6, q1.z) UB200_PUT(7, q1.w)
                UB200_PUT(8, q2.x) UB200_PUT(9, q2.y) UB200_PUT(10, q2.z) UB200_PUT(11, q2.w)
                UB200_PUT(12, q3.x) UB200_PUT(13, q3.y) UB200_PUT(14, q3.z) UB200_PUT(15, q3.w)
#undef UB200_PUT
                const uint32_t total = fill + remaining;
                for (uint32_t i = 1; i < nslots; i++) {
                    fill = kSlotCap4;
                    send_msg(0u, 0u);
                }
                fill = total - (nslots - 1u) * kSlotCap4;
                return;
            }
            if (remaining <= kSlotCap4 - fill) {
                // common case: all of these hits fit the open message
                uint32_t pa = list_a + (((nmsg % kSlots4) * kSlotCap4 + fill + excl) << 2);
                while (hb) {
                    uint32_t bit;
                    asm("bfind.u32 %0, %1;" : "=r"(bit) : "r"(hb));   // highest set bit (FLO)
                    hb ^= 1u << bit;
                    const uint32_t wi = idx + ((bit & 12u) << 5) + (bit & 3u);
                    sts32_4(pa, lds32_4(stage_a + ((wi - base) << 2)));
                    pa += 4u;
                }
                fill += remaining;
                return;
            }
            // the hits spill over into further messages: reserve every slot they need first (a step holds at most 512
            // hits = 8 slots), let all lanes copy their hits in ONE pass (hit p of the run goes to slot p / 64, word
            // p % 64) and then send the full messages; the last one stays open
            if (fill + remaining > kSlots4 * kSlotCap4) {   // would need a ninth slot: flush the open message
                send_msg(0u, 0u);
                open_msg();
            }
            const uint32_t total = fill + remaining;
            const uint32_t nslots = (total + kSlotCap4 - 1u) / kSlotCap4;
            for (uint32_t i = 1; i < nslots; i++) {
                const uint32_t mi = nmsg + i;
                mbar_wait_sleep(bars_a + 8 * (kBarEmpty + mi % kSlots4), ((mi / kSlots4) & 1u) ^ 1u, 512u);
            }
            uint32_t pidx = fill + excl;
            while (hb) {
                const uint32_t bit = __ffs(hb) - 1;
                hb &= hb - 1;
                const uint32_t wi = idx + ((bit >> 2) << 7) + (bit & 3u);
                const uint32_t slot = (nmsg + (pidx >> 6)) % kSlots4;
                sts32_4(list_a + ((slot * kSlotCap4 + (pidx & 63u)) << 2), lds32_4(stage_a + ((wi - base) << 2)));
                pidx++;
            }
            for (uint32_t i = 1; i < nslots; i++) {
                fill = kSlotCap4;
                send_msg(0u, 0u);
            }
            fill = total - (nslots - 1u) * kSlotCap4;
        };

        // tiles are handed out in DFS order by a per-scan-group counter: balances uneven tiles, and the CTAs of
        // different scan groups still walk the tree in the same order (one HBM read, the rest from L2).  The next
        // tile and its metadata are fetched while the last blocks of the current one are scanned.
        struct TileMeta { uint32_t t, n0, n1, lvl0, sseg, w0, w1; };
        auto fetch_tile = [&]() -> TileMeta {
            TileMeta m;
            uint32_t t = 0;
            for (;;) {
                if (lane == 0) t = atomicAdd(p.tile_counter + sg, 1u);
                m.t = __shfl_sync(FULL, t, 0);
                if (!COLLECT) break;
                if (!p.tile_min || m.t >= p.n_tiles) break;
                // collect pass: does any sample of this scan's groups have a candidate at its final best in this tile?
                bool want = false;
                for (uint32_t c = 0; c < (uint32_t)NC; c++) {
                    const uint32_t lg = sg * NC + c;
                    if (lg >= p.ngroups) break;
                    const uint32_t smp = (p.group0 + lg) * 32u + lane;
                    if (smp < p.n_samples)
                        want |= p.tile_min[((size_t)(p.group0 + lg) * p.n_tiles + m.t) * 32u + lane] == p.target_rel[smp];
                }
                if (__any_sync(FULL, want)) break;
            }
            const uint32_t tt = min(m.t, p.n_tiles - 1u);
            m.n0 = p.tile_start[tt]; m.n1 = p.tile_start[tt + 1];
            m.lvl0 = p.tile_lvl[tt]; m.sseg = p.tile_sseg[tt];
            m.w0 = p.tile_w0[tt]; m.w1 = p.tile_w0[tt + 1];
            return m;
        };
        TileMeta cur = fetch_tile();
        for (;;) {
            open_msg();
            if (cur.t >= p.n_tiles) {
                send_msg(kMsgEnd4, 0u);
                break;
            }
            send_msg(kMsgTile4, cur.t);
            rows_end = cur.w1 * (kChunk3 / 128u);
            stage_load(gstep, cur.w0 * (kChunk3 / 128u));
            // segments of the tile in stream order: seed segments, then one segment per block
            const uint32_t nseed = (cur.lvl0 + 31u) >> 5;
            const uint32_t b0 = cur.n0 >> 5, nb = (cur.n1 - cur.n0 + 31u) >> 5;
            const uint32_t nseg = nseed + nb;
            const uint32_t b_fetch = nb > 3u ? nb - 3u : 0u;
            TileMeta nxt = cur;
            uint32_t bw = 0;
            uint32_t off = cur.w0 * kChunk3;
            uint32_t si = 0;                       // current segment
            uint32_t seg_end = off;
            auto next_segment = [&]() {            // end of segment si (stream word offset)
                if (si < nseed) {
                    seg_end = p.seed_end[cur.sseg + si] * 4u;
                } else {
                    const uint32_t b = si - nseed;
                    if ((b & 31u) == 0) bw = (b + lane < nb) ? __ldg(p.blk_words + b0 + b + lane) : 0u;
                    if (b == b_fetch) nxt = fetch_tile();
                    seg_end = off + __shfl_sync(FULL, bw, b & 31u);
                }
            };
            next_segment();
            open_msg();
            for (uint32_t base = off;; base += 512u) {
                // ---- one step: 4 rows = 16 words per lane, loaded and tested once
                // everything issued so far has to be there (this step's rows went out one step ago)
                PROF_T0(tl);
                // the next step's copy goes out first (into the stage read one step ago), then wait for this step's
                stage_load(gstep + 1u, (base >> 7) + 4u);
                stage_a = mring_a + (gstep & 1u) * 2048u;
                // (a tile without any mutation has no rows: its empty segments still take one pass through this loop)
                if ((base >> 7) < rows_end) {
                    mbar_wait(bars_a + 8 * (kBarStage + (gstep & 1u)), (gstep >> 1) & 1u);
                    gstep++;
                }
                PROF_ADD(3, tl);
                const uint32_t idx = base + 4u * lane;
                uint32_t acc = 0;                                  // hit bits enter at bit 31, oldest ends lowest
                auto row_of = [&](uint32_t k) { return lds128_4(stage_a + ((128u * k + 4u * lane) << 2)); };
                auto test4 = [&](const uint4& q) {
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.x), 0u, q.x), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.y), 0u, q.y), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.z), 0u, q.z), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_a, bm_g, q.w), 0u, q.w), 1u);
                };
                q0 = row_of(0); q1 = row_of(1); q2 = row_of(2); q3 = row_of(3);
                test4(q0); test4(q1); test4(q2); test4(q3);
                const uint32_t hb_step = acc >> 16;                // bit 4k+j = word j of quad k
                PROF_ADD(4, tl);                                   // load wait + test
                PROF_INC(5, 1);                                    // steps
                // ---- hand the step's hits out, segment by segment
                PROF_T0(te);
                bool tile_done = false;
                for (;;) {
                    const uint32_t lim = min(seg_end, base + 512u);
                    if (lim > off) emit(hb_step, base, off, lim);
                    off = lim;
                    if (off == seg_end) {
                        send_msg(kMsgLast4, 0u);
                        if (++si == nseg) { tile_done = true; break; }
                        next_segment();
                        open_msg();
                        if (seg_end == off) continue;              // empty segment
                    }
                    if (off == base + 512u) break;
                }
                PROF_ADD(6, te);                                   // emission incl. slot waits
                if (tile_done) break;
            }
            cur = nxt;
        }
#ifdef UB200_PROFILE
        pc[0] = (unsigned long long)(clock64() - prof_start);
        if (lane == 0) for (int i = 0; i < 7; i++) atomicAdd(p.prof + i, pc[i]);
#endif
    } else {
        // =====================================================================================================
        // consumer role-1 of the unit: one group of 32 samples
        // =====================================================================================================
        const uint32_t cons = role - 1u;
        const uint32_t lgroup = sg * NC + cons;                  // group within this pass
        const bool dead = lgroup >= p.ngroups;                   // the last scan group may be short
        const uint32_t ggroup = p.group0 + lgroup;
        uint8_t* cbase = ubase + C::kShared + cons * C::kCons;
        int* dnode = reinterpret_cast<int*>(cbase + C::kODnode);
        int16_t* stk = reinterpret_cast<int16_t*>(cbase + C::kOStack);
        int* neg = reinterpret_cast<int*>(cbase + C::kONeg);
        uint32_t* area = reinterpret_cast<uint32_t*>(cbase + C::kOArea);
        const uint32_t* tabg = p.tab + (size_t)ggroup * p.L * 8u;
        int32_t* gstk = p.gstack
                            ? p.gstack + ((size_t)((blockIdx.x * C::kUnits + unit) * NC + cons) * p.gstack_levels) * 32u
                            : nullptr;
        const uint32_t sample = ggroup * 32u + lane;
        const bool live = !dead && sample < p.n_samples;

        auto stack_read = [&](uint32_t level, uint32_t s) -> int {
            if (__builtin_expect(level >= (uint32_t)C::kStack, 0)) return spill_read4(gstk, level - C::kStack, s);
            return stk[level * 32u + s];
        };
        auto stack_write = [&](uint32_t level, uint32_t s, int v) {
            if (__builtin_expect(level >= (uint32_t)C::kStack, 0)) spill_write4(gstk, level - C::kStack, s, v);
            else stk[level * 32u + s] = (int16_t)v;
        };

        // per-lane (= sample) running best (COLLECT: the known final best, fixed)
        int bsc = COLLECT ? (live ? p.target_rel[sample] : (int)0x80000000) : 0x7fffffff;
        unsigned long long bkey = ~0ull;
        uint32_t cnt = 0;
        uint32_t cur_t = 0;            // the tile being scored (tile_min notes)
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
            if (NOTES) atomicMin(p.tile_min + ((size_t)ggroup * p.n_tiles + cur_t) * 32u + lane, sc);   // candidates are rare
        };
        int negr = 0;   // this lane's (= sample's) share of neg accumulated by the dense form of C
        auto zero_dnode = [&]() {
#pragma unroll
            for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(dnode)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
            neg[lane] = 0;
            negr = 0;
        };
        // ---- C: the hit words of one message, 32 at a time (lane = hit while the table rows are fetched)
        // sparse form: the lane walks the samples that call its hit's position
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
        // dense form (some position is called by more than two samples of the group): the 32 x 32 bit matrix
        // hit x sample is transposed with 5 shuffles, after which lane = SAMPLE walks its own hits.  The lane owns
        // column `lane` of dnode, so the updates need no atomics and hit no bank twice, and its share of neg stays
        // in a register.  The hits' cost nibbles and (node lane, ref/prev/mut) words are staged in `area`.
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
        // 32 hits: the sparse form costs ~20 issue slots per caller of the busiest hit, the dense one ~12 per hit of
        // the busiest sample (an N run makes ONE sample own most of 32 position-sorted hits: sparse wins there)
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
        auto process = [&](uint32_t slot, uint32_t n) {
            static_assert(kSlotCap4 == 64, "two hits per lane");
            const bool h0 = lane < n, h1 = lane + 32u < n;
            uint32_t w0 = 0, w1 = 0;
            uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
            if (h0) {
                w0 = list[slot * kSlotCap4 + lane];
                ldg_row(tabg + (size_t)mut3_pos<NARROW>(w0) * 8u, a0, a1);
            }
            if (h1) {
                w1 = list[slot * kSlotCap4 + 32u + lane];
                ldg_row(tabg + (size_t)mut3_pos<NARROW>(w1) * 8u, b0, b1);
            }
            // rows of absent hits are zero; rows of another group's hits have an empty sample mask
            half(w0, a0, a1);
            if (n > 32u) half(w1, b0, b1);
        };
        uint32_t nmsg = 0;
        // receive the messages of one segment and fold their hits into dnode / neg
        auto take_segment = [&]() {
            for (;;) {
                const uint32_t s = nmsg % kSlots4;
                PROF_T0(tw);
                const uint32_t polls = mbar_wait_sleep(bars_a + 8 * (kBarFull + s), (nmsg / kSlots4) & 1u, 512u);
                PROF_ADD(9, tw); PROF_INC(10, polls); (void)polls;
                const uint32_t m = msg[s].x;
                PROF_T0(tp);
                if (m & 0xffffu) process(s, m & 0xffffu);
                PROF_ADD(11, tp); PROF_INC(12, 1);
                __syncwarp();
                if (elect_one()) mbar_arrive(bars_a + 8 * (kBarEmpty + s));
                nmsg++;
                if (m & kMsgLast4) break;
            }
            __syncwarp();
        };

        for (;;) {
            uint32_t t;
            {
                const uint32_t s = nmsg % kSlots4;
                PROF_T0(tw);
                const uint32_t polls = mbar_wait_sleep(bars_a + 8 * (kBarFull + s), (nmsg / kSlots4) & 1u, 1024u);
                PROF_ADD(13, tw); PROF_INC(10, polls); (void)polls;
                const uint32_t m = msg[s].x;
                t = msg[s].y;
                __syncwarp();
                if (elect_one()) mbar_arrive(bars_a + 8 * (kBarEmpty + s));
                nmsg++;
                if (m & kMsgEnd4) break;
                if (dead || !(m & kMsgTile4)) continue;   // a consumer without a group only releases the slots
            }
            const uint32_t n0 = p.tile_start[t], n1 = p.tile_start[t + 1];
            const uint32_t lvl0 = p.tile_lvl[t];
            if (NOTES) cur_t = t;
            // block records: one 16-byte word per block (same address for every lane), fetched one block ahead
            const uint4* recp = p.blk_rec + (n0 >> 5);
            uint4 rnext = __ldg(recp);

            // cross-unit bound of this lane's sample, and the tile-local floor of every stack value
            const int gb = COLLECT ? bsc : (live ? *(volatile int*)(p.gbest + sample) : 0x7fffffff);
            int gmin = 0;

            // ================= seed: path corrections of levels 0 .. lvl0-1, 32 levels per segment =============
            for (uint32_t l0 = 0; l0 < lvl0; l0 += 32u) {
                zero_dnode();
                __syncwarp();
                take_segment();
                const uint32_t cn = min(32u, lvl0 - l0);
                int v = l0 ? stack_read(l0 - 1u, lane) : 0;
                for (uint32_t j = 0; j < cn; j++) {
                    v += dc_of(dnode[j * 32u + lane]);
                    stack_write(l0 + j, lane, v);
                    gmin = min(gmin, v);
                }
                __syncwarp();
            }

            for (uint32_t blk = n0; blk < n1; blk += 32u) {
                const uint4 rec = rnext;
                if (blk + 32u < n1) rnext = __ldg(++recp);
                zero_dnode();
                __syncwarp();

                // ================= C: the block's hits from the scanner =================
                take_segment();

                // ================= bound: can any pair of this block still be optimal? =================
                const int lbase = gmin + neg[lane] + negr;
                const int bound = min(bsc, gb);
                const uint32_t needs = __ballot_sync(FULL, live && (SCORES || (int)rec.x + lbase <= bound));
                PROF_T0(tn);
                if (needs) {
                    PROF_INC(15, 1);
                    // ---- A: headers (lane = node), only for blocks that get here
                    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.hdr) + blk + lane);
                    const bool act = blk + lane < n1;
                    const uint32_t level = h.z >> kLevelShift, flags = h.z & 0x3fffu;
                    const bool dense_ok = act && (flags & kFlagValid0);
                    const int min_g = __reduce_min_sync(FULL, dense_ok ? (int)h.x : BIG);            // signed: G can be < 0
                    // hit nodes of this lane's sample = non-zero column entries (a pair whose packed delta is zero
                    // scores exactly like a pair without a hit)
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
                    uint32_t need_e = __ballot_sync(FULL, live && (SCORES || (min_g < BIG && min_g + lbase <= bound)));
                    while (need_e) {
                        const uint32_t s = __ffs(need_e) - 1;
                        need_e &= need_e - 1;
                        const uint32_t hm_s = area[kA4Hm + s];
                        const int sc = (int)h.x + above(level, h.y, hm_s, s);
                        if (SCORES) {
                            const uint32_t gs = ggroup * 32u + s;
                            if (act && !((hm_s >> lane) & 1u))
                                p.node_scores[(size_t)gs * p.n_nodes + blk + lane] = p.base[gs] + sc + ((flags & kFlagValid0) ? 0 : 1);
                            continue;
                        }
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
                            if (SCORES) p.node_scores[(size_t)sample * p.n_nodes + blk + n] = p.base[sample] + sc + (valid ? 0 : 1);
                            else if (valid && sc <= bsc) merge(sc, hu, blk + n);
                        }
                    }
                }
                __syncwarp();
                PROF_ADD(14, tn);

                // ================= G: stack rows of the open chain (lane = sample) =================
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
                __syncwarp();
            }
            // publish an improved bound for the other units working on this sample group
            if (BEST && live && bsc < gb) atomicMin(p.gbest + sample, bsc);
            __syncwarp();
        }

#ifdef UB200_PROFILE
        pc[8] = (unsigned long long)(clock64() - prof_start);
        if (lane == 0 && !dead) for (int i = 8; i < 16; i++) atomicAdd(p.prof + i, pc[i]);
#endif
        if (BEST) {
            // park the consumer's result in its own rows for the fold below
            reinterpret_cast<unsigned long long*>(dnode)[lane] = bkey;
            neg[lane] = (int)cnt;
        }
    }

    if (!BEST) return;
    // fold the CTA's units, one partial row per CTA and group: warp c folds consumer c of every unit
    __syncthreads();
    if (warp < (uint32_t)NC && sg * NC + warp < p.ngroups) {
        unsigned long long best = ~0ull;
        for (int u = 0; u < C::kUnits; u++) {
            const uint8_t* pb = smem + bm_bytes + kLut4Bytes + u * C::kUnit + C::kShared + warp * C::kCons;
            best = min(best, reinterpret_cast<const unsigned long long*>(pb + C::kODnode)[lane]);
        }
        uint32_t c = 0;
        for (int u = 0; u < C::kUnits; u++) {
            const uint8_t* pb = smem + bm_bytes + kLut4Bytes + u * C::kUnit + C::kShared + warp * C::kCons;
            if ((reinterpret_cast<const unsigned long long*>(pb + C::kODnode)[lane] >> 33) == (best >> 33))
                c += reinterpret_cast<const uint32_t*>(pb + C::kONeg)[lane];
        }
        const size_t o = ((size_t)(p.part_group0 + sg * NC + warp) * p.part_stride + cta_in_sg) * 32u + lane;
        p.part_key[o] = best;
        p.part_cnt[o] = c;
    }
}

}  // namespace ub200
