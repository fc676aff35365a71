// k_score3 — streaming best-placement kernel on the segment layout (ub200_internal.h, DESIGN.md "Kernels").
//
// One persistent launch scores every node of the tree against NG groups of 32 samples.  The unit of work is a
// tile: a contiguous DFS range of whole 32-node blocks whose mutations are ONE contiguous piece of the stream,
// [seed segments][block segments].  A CTA holds 16 warp PAIRS; a pair works on one tile at a time:
//
//   scanner warp   pulls the tile's stream through a 4 KB shared-memory ring of 512 B rows, one 16-byte cp.async
//                  per lane per row (LDGSTS, cp.async groups; a one-warp producer pays ~3 issue slots per row
//                  this way, a bulk copy with its mbarrier ~12), and tests every word's position against the
//                  group's bitmap: a step = 4 rows = 16 words per lane: four LDS.128, then per word LEA.HI (byte
//                  offset of the bitmap word), LDS, a wrap shift (bit to bit 0) and a funnel shift that collects
//                  the hit bits.  Hit words are compacted (two ballots per step) into one of eight 64-word
//                  message slots and handed to the consumer through full/empty mbarriers; a segment is one or
//                  more messages, the last one flagged.  The scanner needs no sample state, only the bitmap.
//   consumer warp  owns the sample state.  Per block:
//     A  lane = node     header decode (one coalesced 512 B load per block, fetched a block ahead into registers)
//     C  lane = hit      hit word -> node lane (stored in the word), table row of the position (32 B, L2; two
//                        rows per lane in flight) ->
//                        for every sample calling the position: packed (dcorr, da, dcommon) from a 1024-entry
//                        LUT, shared-memory atomics into dnode[node][sample], hm[sample] |= node,
//                        neg[sample] += min(dcorr, 0)
//     bound              exact lower bound of every pair of the block:  min(G - nmut) + gmin + neg  against the
//                        running best; blocks that cannot hold an optimum skip E and F entirely
//     E  lane = node     non-hit pairs of a sample that can still improve or tie
//     F  lane = sample   hit pairs, exactly (score, validity, tie key)
//     G  lane = sample   path corrections of the block's OPEN chain (nodes with descendants in later blocks)
//                        -> stack rows the following blocks read
//   Seed segments (the rows of the tile's root path, 32 levels per segment) go through the same scanner ->
//   consumer path and initialise the stack.
// The correction of the path above a node inside its own block is never materialised: it is
// stack[level above the block] + sum of dnode over (in-block ancestors & hit nodes of the sample), with the
// in-block ancestor mask precomputed in the header.  All pruning is exact, so results are schedule-independent.
#pragma once
#include "score_kernel.cuh"

namespace ub200 {

constexpr int kPairs3 = 16;
constexpr int kThreads3 = kPairs3 * 64;
constexpr uint32_t kRingRows3 = 8;                         // rows of 128 words (512 B)
constexpr uint32_t kRingWords3 = kRingRows3 * 128;         // 1024 words = 4 KB
constexpr uint32_t kSlots3 = 8, kSlotCap3 = 64;            // scanner -> consumer messages
constexpr int kStack3 = 40;                                // levels kept in shared memory (deeper: HBM spill)
// per-pair shared memory (bytes)
constexpr uint32_t kO3Mring = 0;                           // u32[1024]            scanner
constexpr uint32_t kO3Dnode = 4096;                        // i32[32][32] packed deltas   consumer
constexpr uint32_t kO3Stack = 8192;                        // i16[40][32]
constexpr uint32_t kO3List = kO3Stack + kStack3 * 64;      // u32[8][64] hit words
constexpr uint32_t kO3Info = kO3List + kSlots3 * kSlotCap3 * 4;   // u32[6][32]: G, z, w, am, hm, neg
constexpr uint32_t kO3Msg = kO3Info + 6 * 128;             // uint2[8]: (count | flags << 16, payload)
constexpr uint32_t kO3Bars = kO3Msg + kSlots3 * 8;         // mbarriers: 8 full, 8 empty
constexpr uint32_t kWarpSmem3 = (kO3Bars + 2 * kSlots3 * 8 + 127) & ~127u;
constexpr uint32_t kI3G = 0, kI3Z = 32, kI3W = 64, kI3Am = 96, kI3Hm = 128, kI3Neg = 160;
constexpr uint32_t kBarFull = 0, kBarEmpty = kSlots3;
constexpr uint32_t kMsgLast = 1u << 16, kMsgTile = 2u << 16, kMsgEnd = 4u << 16;
constexpr uint32_t kLut3Bytes = 4096;
constexpr uint32_t kMaxRowV3 = 500;      // packed 10-bit delta fields
constexpr uint32_t kMaxCallsV3 = 32000;  // path corrections are int16
constexpr uint32_t kSmemLimit3 = 232448; // 227 KB

struct Score3Params {
    const uint32_t* stream;
    const NodeHdr* hdr;           // hdr3
    const uint32_t* tiekey;
    const uint32_t* blk_words;    // [blocks] stream words of each 32-node block segment (multiple of 4)
    const uint32_t* tile_start;   // [T+1]
    const uint32_t* tile_w0;      // [T+1]
    const uint32_t* tile_lvl;     // [T]
    const uint32_t* tile_sseg;    // [T+1]
    const uint32_t* seed_end;
    uint32_t n_nodes, n_tiles, L, bitmap_words;
    const uint32_t* bitmap;
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
};

__device__ __forceinline__ uint32_t lds32_3(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32_3(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128_3(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __noinline__ int spill_read3(const int32_t* gstk, uint32_t level, uint32_t s) {
    return gstk[(size_t)(level - kStack3) * 32u + s];
}
__device__ __noinline__ void spill_write3(int32_t* gstk, uint32_t level, uint32_t s, int v) {
    gstk[(size_t)(level - kStack3) * 32u + s] = v;
}
__device__ __forceinline__ int lut_delta3(uint32_t i) {
    const uint32_t e = i >> 6, refc = (i >> 4) & 3u, prevc = (i >> 2) & 3u, mutc = i & 3u;
    const int rm = (mutc != refc), rp = (prevc != refc);
    const int wm = (e >> mutc) & 1u, wp = (e >> prevc) & 1u;
    const int dcorr = (wm - wp) - (rm - rp);
    const int tk = wm ^ 1, t0 = rm ^ 1;
    const int da = (tk & wp) - (t0 & rp);
    const int dcom = tk - t0;
    return dcorr * (1 << 20) + da * (1 << 10) + dcom;
}
// bitmap word of a stream word's position: narrow words carry the byte offset at bit 14
template <bool SMEM_BITMAP, bool NARROW>
__device__ __forceinline__ uint32_t bitmap_word(const uint32_t* bm_s, const uint32_t* bm_g, uint32_t w) {
    if (NARROW) {
        const uint32_t off = w >> 14;
        return SMEM_BITMAP ? *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(bm_s) + off)
                           : __ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(bm_g) + off));
    }
    return SMEM_BITMAP ? bm_s[w >> 14] : __ldg(bm_g + (w >> 14));
}
__device__ __forceinline__ int dc_of(int v) { return (v + (1 << 19)) >> 20; }
__device__ __forceinline__ void unpack_delta3(int v, int& dcorr, int& da, int& dcom) {
    dcom = (int)((uint32_t)v << 22) >> 22;
    const int v1 = (v - dcom) >> 10;
    da = (int)((uint32_t)v1 << 22) >> 22;
    dcorr = (v1 - da) >> 10;
}

// COLLECT = false: best placement per sample.  COLLECT = true: second pass that lists every optimal node of each
// sample (best_j_vec + node_has_unique); the final best score is the bound, so almost every block is pruned.
template <bool SMEM_BITMAP, bool COLLECT, bool NARROW>
__global__ void __launch_bounds__(kThreads3, 1) k_score3(const Score3Params p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t pair = warp >> 1;
    // scanner / consumer alternate so that every SM sub-partition (warp % 4) gets both kinds
    const bool is_scanner = (warp & 1u) == ((warp >> 2) & 1u);
    const uint32_t group = blockIdx.x % p.ngroups;
    const uint32_t cta_in_group = blockIdx.x / p.ngroups;
    const uint32_t ctas_per_group = gridDim.x / p.ngroups;
    const uint32_t ggroup = p.group0 + group;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t lt_mask = (1u << lane) - 1u;
    constexpr int BIG = 0x3fffffff;

    // ---- shared memory: [bitmap][lut][pair 0 .. pair 15]
    uint32_t* bm_s = reinterpret_cast<uint32_t*>(smem);
    const uint32_t bm_bytes = SMEM_BITMAP ? ((p.bitmap_words * 4u + 127u) & ~127u) : 0u;
    int* lut = reinterpret_cast<int*>(smem + bm_bytes);
    uint8_t* wbase = smem + bm_bytes + kLut3Bytes + pair * kWarpSmem3;
    uint32_t* mring = reinterpret_cast<uint32_t*>(wbase + kO3Mring);
    int* dnode = reinterpret_cast<int*>(wbase + kO3Dnode);
    int16_t* stk = reinterpret_cast<int16_t*>(wbase + kO3Stack);
    uint32_t* list = reinterpret_cast<uint32_t*>(wbase + kO3List);
    uint32_t* info = reinterpret_cast<uint32_t*>(wbase + kO3Info);
    volatile uint2* msg = reinterpret_cast<volatile uint2*>(wbase + kO3Msg);
    const uint32_t mring_a = smem_u32(mring), bars_a = smem_u32(wbase + kO3Bars);
    const uint32_t list_a = smem_u32(list);

    const uint32_t* bm_g = p.bitmap + (size_t)ggroup * p.bitmap_words;
    if (SMEM_BITMAP) {
        const uint4* src = reinterpret_cast<const uint4*>(bm_g);
        uint4* dst = reinterpret_cast<uint4*>(bm_s);
        for (uint32_t i = threadIdx.x; i < p.bitmap_words / 4; i += kThreads3) dst[i] = __ldg(src + i);
    }
    for (uint32_t i = threadIdx.x; i < 1024; i += kThreads3) lut[i] = lut_delta3(i);
    if (is_scanner) {
        // lanes past a segment's end still index the bitmap with what the ring holds: only ever valid words
        for (uint32_t i = lane; i < kRingWords3; i += 32) mring[i] = 0u;
        if (lane == 0) {
            for (uint32_t i = 0; i < kBarEmpty + kSlots3; i++) mbar_init(bars_a + 8 * i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();

    if (is_scanner) {
        // =====================================================================================================
        // scanner
        // =====================================================================================================
        uint32_t lrow = 0, rows_end = 0;   // next ring row to load, end of the tile's rows
        uint32_t nmsg = 0;        // messages sent so far; the open one lives in slot nmsg % kSlots3
        uint32_t fill = 0;        // hit words in the open message
        auto open_msg = [&]() {   // wait until the consumer has released the slot
            const uint32_t s = nmsg % kSlots3;
            mbar_wait_long(bars_a + 8 * (kBarEmpty + s), ((nmsg / kSlots3) & 1u) ^ 1u);
            fill = 0;
        };
        auto send_msg = [&](uint32_t flags, uint32_t payload) {
            const uint32_t s = nmsg % kSlots3;
            __syncwarp();
            if (elect_one()) {
                msg[s].x = fill | flags;
                msg[s].y = payload;
                mbar_arrive(bars_a + 8 * (kBarFull + s));
            }
            nmsg++;
        };

        // scan stream words [o0, o1) (both multiples of 4) of the current tile, 512 words per step; the hit
        // words go to the open message, which is sent (not flagged last) and reopened whenever it is full
        // ring loader: rows of 128 words (512 B, absolute-aligned), one 16-byte cp.async per lane per row
        // (LDGSTS, L2 -> shared memory without registers).  The bookkeeping is two counters and cp.async groups;
        // a one-warp producer pays ~3 issue slots per 512 B this way, against ~12 for a bulk copy with its
        // mbarrier — and this kernel is bound by issue slots, not by copy bandwidth.
        auto ring_fill = [&](uint32_t cur_row) {
            const uint32_t lim_row = min(rows_end, cur_row + kRingRows3);
            while (lrow < lim_row) {
                cp_async16(mring_a + (((lrow & (kRingRows3 - 1u)) << 9) + (lane << 4)),
                           p.stream + ((size_t)lrow << 7) + (lane << 2));
                lrow++;
            }
            cp_async_commit();
        };
        auto scan = [&](uint32_t o0, uint32_t o1) {
            for (uint32_t off = o0; off < o1;) {
                // a step = the (up to) 4 rows from the one holding `off`, clipped to the segment
                const uint32_t r0 = off >> 7;
                const uint32_t lim = min((r0 + 4u) << 7, o1);
                // everything issued so far has to be there (this step's rows went out one step ago)
                cp_async_wait<0>();
                __syncwarp();
                const uint32_t base = r0 << 7;
                const uint32_t nq = (lim - base + 127u) >> 7;     // rows (LDS.128 per lane) in this step, 1..4
                const uint32_t idx = base + 4u * lane;
                uint32_t acc = 0;                                  // hit bits enter at bit 31, oldest ends lowest
                auto row_of = [&](uint32_t k) { return lds128_3(mring_a + (((idx + 128u * k) & (kRingWords3 - 1u)) << 2)); };
                auto test4 = [&](const uint4& q) {
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_s, bm_g, q.x), 0u, q.x), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_s, bm_g, q.y), 0u, q.y), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_s, bm_g, q.z), 0u, q.z), 1u);
                    acc = __funnelshift_r(acc, __funnelshift_r(bitmap_word<SMEM_BITMAP, NARROW>(bm_s, bm_g, q.w), 0u, q.w), 1u);
                };
                if (nq == 4u) {   // the usual step: one basic block, so that the loads of all four rows overlap
                    const uint4 q0 = row_of(0), q1 = row_of(1), q2 = row_of(2), q3 = row_of(3);
                    // rows below r0 are dead: top the ring up (needed one step from now) while the loads are in flight
                    ring_fill(r0);
                    test4(q0); test4(q1); test4(q2); test4(q3);
                } else {
                    ring_fill(r0);
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++) {
                        if (k < nq) test4(row_of(k));
                        else acc >>= 4;
                    }
                }
                uint32_t hb = acc >> 16;                           // bit 4k+j = word j of quad k
                if (off != base || lim != base + 512u) {          // quads outside [off, lim) belong to other segments
                    const int lo = (int)(off - idx), hi = (int)(lim - idx);   // multiples of 4
                    const uint32_t k_lo = lo > 0 ? min((uint32_t)(lo + 127) >> 7, 4u) : 0u;
                    const uint32_t k_hi = hi > 0 ? min((uint32_t)(hi + 127) >> 7, 4u) : 0u;
                    hb &= ((1u << (4u * k_hi)) - 1u) & ~((1u << (4u * k_lo)) - 1u);
                }
                off = lim;
                // compact: lane l's hits follow those of lanes < l.  Exclusive prefix of the per-lane counts:
                // two ballots when no lane has more than 3 hits (the usual case), else a shuffle scan
                const uint32_t c = __popc(hb);
                uint32_t excl, remaining;
                if (__ballot_sync(FULL, c > 3u) == 0u) {
                    const uint32_t v0 = __ballot_sync(FULL, c & 1u), v1 = __ballot_sync(FULL, c & 2u);
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
                if (remaining <= kSlotCap3 - fill) {
                    // common case: all of the step's hits fit the open message
                    uint32_t pa = list_a + (((nmsg % kSlots3) * kSlotCap3 + fill + excl) << 2);
                    while (hb) {
                        uint32_t bit;
                        asm("bfind.u32 %0, %1;" : "=r"(bit) : "r"(hb));   // highest set bit (FLO)
                        hb ^= 1u << bit;
                        const uint32_t wi = idx + ((bit & 12u) << 5) + (bit & 3u);
                        sts32_3(pa, lds32_3(mring_a + ((wi & (kRingWords3 - 1u)) << 2)));
                        pa += 4u;
                    }
                    fill += remaining;
                    continue;
                }
                uint32_t mine = excl;           // index of this lane's next hit among the step's hits
                uint32_t done = 0;              // hits of this step already placed in messages
                while (remaining) {
                    if (fill == kSlotCap3) {
                        send_msg(0u, 0u);
                        open_msg();
                    }
                    const uint32_t take = min(kSlotCap3 - fill, remaining);
                    const uint32_t slot_a = (nmsg % kSlots3) * kSlotCap3 + fill;
                    while (hb && mine < done + take) {
                        const uint32_t bit = __ffs(hb) - 1;
                        hb &= hb - 1;
                        const uint32_t wi = idx + ((bit >> 2) << 7) + (bit & 3u);
                        list[slot_a + (mine - done)] = lds32_3(mring_a + ((wi & (kRingWords3 - 1u)) << 2));
                        mine++;
                    }
                    fill += take;
                    remaining -= take;
                    done += take;
                }
            }
        };

        // tiles are handed out in DFS order by a per-group counter: balances uneven tiles, and the CTAs of
        // different groups still walk the tree in the same order (one HBM read, the rest from L2).  The next
        // tile and its metadata are fetched while the last blocks of the current one are scanned.
        struct TileMeta { uint32_t t, n0, n1, lvl0, sseg, w0, w1; };
        auto fetch_tile = [&]() -> TileMeta {
            TileMeta m;
            uint32_t t = 0;
            if (lane == 0) t = atomicAdd(p.tile_counter + group, 1u);
            m.t = __shfl_sync(FULL, t, 0);
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
                send_msg(kMsgEnd, 0u);
                break;
            }
            send_msg(kMsgTile, cur.t);
            lrow = cur.w0 * (kChunk3 / 128u);
            rows_end = cur.w1 * (kChunk3 / 128u);
            uint32_t o0 = cur.w0 * kChunk3;
            ring_fill(lrow);
            // seed segments, then one segment per block
            for (uint32_t l0 = 0; l0 < cur.lvl0; l0 += 32u) {
                const uint32_t o1 = p.seed_end[cur.sseg + (l0 >> 5)] * 4u;
                open_msg();
                scan(o0, o1);
                send_msg(kMsgLast, 0u);
                o0 = o1;
            }
            const uint32_t b0 = cur.n0 >> 5, nb = (cur.n1 - cur.n0 + 31u) >> 5;
            const uint32_t b_fetch = nb > 3u ? nb - 3u : 0u;
            TileMeta nxt = cur;
            uint32_t bw = 0;
            for (uint32_t b = 0; b < nb; b++) {
                if ((b & 31u) == 0) bw = (b + lane < nb) ? __ldg(p.blk_words + b0 + b + lane) : 0u;
                if (b == b_fetch) nxt = fetch_tile();
                const uint32_t o1 = o0 + __shfl_sync(FULL, bw, b & 31u);
                open_msg();
                scan(o0, o1);
                send_msg(kMsgLast, 0u);
                o0 = o1;
            }
            cur = nxt;
        }
    } else {
        // =====================================================================================================
        // consumer
        // =====================================================================================================
        const uint32_t* tabg = p.tab + (size_t)ggroup * p.L * 8u;
        int32_t* gstk = p.gstack ? p.gstack + ((size_t)(blockIdx.x * kPairs3 + pair) * p.gstack_levels) * 32u : nullptr;
        const uint32_t sample = ggroup * 32u + lane;
        const bool live = sample < p.n_samples;

        auto stack_read = [&](uint32_t level, uint32_t s) -> int {
            if (__builtin_expect(level >= (uint32_t)kStack3, 0)) return spill_read3(gstk, level, s);
            return stk[level * 32u + s];
        };
        auto stack_write = [&](uint32_t level, uint32_t s, int v) {
            if (__builtin_expect(level >= (uint32_t)kStack3, 0)) spill_write3(gstk, level, s, v);
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
        auto zero_dnode = [&]() {
#pragma unroll
            for (int k = 0; k < 8; k++) reinterpret_cast<uint4*>(dnode)[k * 32 + lane] = make_uint4(0, 0, 0, 0);
            info[kI3Hm + lane] = 0;
            info[kI3Neg + lane] = 0;
        };
        // C: the hit words of one message (lane = hit, two hits per lane so that all table rows are in flight
        // together)
        auto apply_hit = [&](uint32_t w, const uint4& r0, const uint2& r1) {
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
                atomicOr(&info[kI3Hm + s], 1u << nl);
                const int dc = dc_of(d);
                if (dc < 0) atomicAdd(reinterpret_cast<int*>(&info[kI3Neg + s]), dc);
            }
        };
        auto process = [&](uint32_t slot, uint32_t n) {
            static_assert(kSlotCap3 == 64, "two hits per lane");
            const bool h0 = lane < n, h1 = lane + 32u < n;
            uint32_t w0 = 0, w1 = 0;
            uint4 a0 = make_uint4(0, 0, 0, 0), b0 = a0;
            uint2 a1 = make_uint2(0, 0), b1 = a1;
            if (h0) {
                w0 = list[slot * kSlotCap3 + lane];
                const uint32_t* row = tabg + (size_t)mut3_pos<NARROW>(w0) * 8u;
                a0 = __ldg(reinterpret_cast<const uint4*>(row));
                a1 = __ldg(reinterpret_cast<const uint2*>(row + 4));
            }
            if (h1) {
                w1 = list[slot * kSlotCap3 + 32u + lane];
                const uint32_t* row = tabg + (size_t)mut3_pos<NARROW>(w1) * 8u;
                b0 = __ldg(reinterpret_cast<const uint4*>(row));
                b1 = __ldg(reinterpret_cast<const uint2*>(row + 4));
            }
            if (h0) apply_hit(w0, a0, a1);
            if (h1) apply_hit(w1, b0, b1);
        };
        uint32_t nmsg = 0;
        // receive the messages of one segment and fold their hits into dnode / hm / neg
        auto take_segment = [&]() {
            for (;;) {
                const uint32_t s = nmsg % kSlots3;
                mbar_wait_idle(bars_a + 8 * (kBarFull + s), (nmsg / kSlots3) & 1u, p.gbest);
                const uint32_t m = msg[s].x;
                process(s, m & 0xffffu);
                __syncwarp();
                if (elect_one()) mbar_arrive(bars_a + 8 * (kBarEmpty + s));
                nmsg++;
                if (m & kMsgLast) break;
            }
            __syncwarp();
        };

        for (;;) {
            uint32_t t;
            {
                const uint32_t s = nmsg % kSlots3;
                mbar_wait_long(bars_a + 8 * (kBarFull + s), (nmsg / kSlots3) & 1u);
                const uint32_t m = msg[s].x;
                t = msg[s].y;
                __syncwarp();
                if (elect_one()) mbar_arrive(bars_a + 8 * (kBarEmpty + s));
                nmsg++;
                if (m & kMsgEnd) break;
            }
            const uint32_t n0 = p.tile_start[t], n1 = p.tile_start[t + 1];
            const uint32_t lvl0 = p.tile_lvl[t];
            // headers: one coalesced 512 B load per block, fetched one block ahead into registers
            uint4 hnext = __ldg(reinterpret_cast<const uint4*>(p.hdr) + n0 + lane);

            // cross-pair bound of this lane's sample, and the tile-local floor of every stack value
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
                // ================= A: headers (lane = node) =================
                const uint4 h = hnext;
                if (blk + 32u < n1) hnext = __ldg(reinterpret_cast<const uint4*>(p.hdr) + blk + 32u + lane);
                const bool act = blk + lane < n1;
                const uint32_t level = h.z >> kLevelShift, flags = h.z & 0x3fffu;
                const uint32_t nmut = act ? (h.w >> 16) : 0u;
                const bool dense_ok = act && (flags & kFlagValid0);
                const int min_g = __reduce_min_sync(FULL, dense_ok ? (int)h.x : BIG);            // signed: G can be < 0
                const int min_gn = __reduce_min_sync(FULL, act ? (int)h.x - (int)nmut : BIG);
                zero_dnode();
                __syncwarp();

                // ================= C: the block's hits from the scanner =================
                take_segment();

                // ================= bound: can any pair of this block still be optimal? =================
                const uint32_t hmv = info[kI3Hm + lane];             // lane = sample: its hit nodes
                const int lbase = gmin + (int)info[kI3Neg + lane];
                const int bound = min(bsc, gb);
                const uint32_t needs = __ballot_sync(FULL, live && min_gn + lbase <= bound);
                if (needs) {
                    // node-indexed copies of the headers for the lane = sample phase (only blocks that get here pay)
                    info[kI3G + lane] = (uint32_t)h.x;
                    info[kI3Z + lane] = h.z;
                    info[kI3W + lane] = h.w;
                    info[kI3Am + lane] = h.y;
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
                        const uint32_t hm_s = info[kI3Hm + s];
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
                            unpack_delta3(dnode[n * 32u + lane], dcorr, da, dcom);
                            const uint32_t z = info[kI3Z + n], w = info[kI3W + n];
                            const uint32_t fl = z & 0x3fffu;
                            const int g = (int)info[kI3G + n];
                            int sc;
                            bool valid;
                            uint32_t hu;
                            if (fl & kFlagRoot) {
                                sc = g + dcorr; valid = true; hu = 0;
                            } else {
                                const bool masked = fl & kFlagMasked;
                                if (masked) { da = 0; dcom = 0; }
                                sc = g + above(z >> kLevelShift, info[kI3Am + n], hmv, lane) - da;
                                const int common = (int)(w & 0xffffu) + dcom;
                                hu = (masked || (int)(w >> 16) > common) ? 1u : 0u;
                                valid = (fl & kFlagLeaf) ? common > 0 : (!hu || common > 0);
                            }
                            if (valid && sc <= bsc) merge(sc, hu, blk + n);
                        }
                    }
                }
                __syncwarp();

                // ================= G: stack rows of the open chain (lane = sample) =================
                uint32_t chain = __ballot_sync(FULL, act && (flags & kFlagOpen));
                if (chain) {
                    uint32_t lv = __shfl_sync(FULL, level, __ffs(chain) - 1);
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
            // publish an improved bound for the other pairs working on this sample group
            if (!COLLECT && live && bsc < gb) atomicMin(p.gbest + sample, bsc);
            __syncwarp();
        }

        if (!COLLECT) {
            // park the pair's result in its own rows for the fold below
            reinterpret_cast<unsigned long long*>(dnode)[lane] = bkey;
            info[lane] = cnt;
        }
    }

    if (COLLECT) return;
    // fold the CTA's pairs, one partial row per CTA
    __syncthreads();
    if (warp == 0) {
        unsigned long long best = ~0ull;
        for (int w = 0; w < kPairs3; w++) {
            const uint8_t* pb = smem + bm_bytes + kLut3Bytes + w * kWarpSmem3;
            best = min(best, reinterpret_cast<const unsigned long long*>(pb + kO3Dnode)[lane]);
        }
        uint32_t c = 0;
        for (int w = 0; w < kPairs3; w++) {
            const uint8_t* pb = smem + bm_bytes + kLut3Bytes + w * kWarpSmem3;
            if ((reinterpret_cast<const unsigned long long*>(pb + kO3Dnode)[lane] >> 33) == (best >> 33))
                c += reinterpret_cast<const uint32_t*>(pb + kO3Info)[lane];
        }
        const size_t o = ((size_t)group * ctas_per_group + cta_in_group) * 32u + lane;
        p.part_key[o] = best;
        p.part_cnt[o] = c;
    }
}

}  // namespace ub200
