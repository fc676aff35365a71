// C-ABI of the placement engine (include/usher_b200.h): handles, device memory, launches, timing.
// No CPU fallback lives here: every compute entry point needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <cstdlib>

#include "score_kernel.cuh"
#include "fs_kernel.cuh"
#include "score_kernel4.cuh"
#include "ub200_internal.h"
#include "usher_b200.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return (int)e_ > 0 ? (int)e_ : 1;                                                      \
        }                                                                                          \
    } while (0)

template <class T>
int dev_upload(T** dst, const T* src, size_t count, cudaStream_t st) {
    *dst = nullptr;
    CU(cudaMalloc((void**)dst, std::max<size_t>(count, 1) * sizeof(T)));
    if (count) CU(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace

struct ub200_mat {
    // host copies of the derived arrays: the small ones stay, the streamed ones are freed after upload.  Replicas of
    // one tree on several devices (ub200_multi) share one derivation.
    std::shared_ptr<ub200::Derived> dp;
    ub200::Derived& d;
    explicit ub200_mat(std::shared_ptr<ub200::Derived> p) : dp(std::move(p)), d(*dp) {}
    int device = 0;
    uint32_t n = 0, n_tiles = 0, L = 0;
    uint64_t m = 0;
    // device: k_score3 layout (always resident when the genome fits its 23-bit position field)
    uint32_t* mstream = nullptr;
    ub200::NodeHdr* hdr3 = nullptr;
    uint32_t *blk_words = nullptr, *blk_rec = nullptr;
    uint32_t consumers = 0;     // sample groups per scanner (k_score4 NC); 0 = chosen per pass
    uint32_t *tiekey = nullptr, *tile3_start = nullptr, *tile3_w0 = nullptr, *tile3_lvl = nullptr, *tile3_sseg = nullptr,
             *seed_end = nullptr;
    int32_t* gstack3 = nullptr;
    uint32_t gstack3_levels = 0, n_tiles3 = 0;
    // device: k_score layout, uploaded on first use (per-node scores, oversized rows / call lists)
    bool v1_resident = false;
    uint32_t* mutw = nullptr;
    ub200::NodeHdr* hdr = nullptr;
    uint32_t *row32 = nullptr, *tile_start = nullptr, *anc_ptr = nullptr, *anc = nullptr;
    uint32_t *key_to_node = nullptr, *tie_index = nullptr, *num_leaves = nullptr;
    int32_t* gstack = nullptr;
    uint32_t gstack_levels = 0;
    uint64_t device_bytes = 0;
    int num_sms = 0;
    uint32_t grid = 0;          // CTAs per scoring launch
    uint32_t pass_groups = 1;   // sample groups (x32 samples) per pass over the tree
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::vector<cudaEvent_t> ev;
    // timing of the last place call
    ub200_timing last = {};
    struct Span { int a, b, kind; };
    std::vector<Span> spans;
    int ev_used = 0;
    struct ub200_samples* scratch = nullptr;   // reused by ub200_place_batch
};

struct ub200_samples {
    ub200_mat* mat = nullptr;
    uint32_t n_samples = 0, n_groups = 0;
    uint64_t n_calls = 0;
    ub200_mutation* calls = nullptr;
    unsigned long long* sample_ptr = nullptr;
    uint32_t* call_sample = nullptr;
    uint32_t* bitmap = nullptr;
    uint32_t bitmap_words = 0;
    uint32_t* tab = nullptr;
    int32_t* base = nullptr;
    int32_t* gbest = nullptr;
    uint32_t* tile_counter = nullptr;
    ub200_placement* results = nullptr;
    int32_t* best_rel = nullptr;
    unsigned long long* part_key = nullptr;
    uint32_t* part_cnt = nullptr;
    uint32_t part_groups = 0, part_wpg = 0;
    size_t part_rows = 0;          // rows of 32 lanes allocated in part_key / part_cnt
    uint32_t launch_idx = 0;       // scoring launches of the current place call (tile counter block)
    int32_t* node_scores = nullptr;
    uint32_t* set_out = nullptr;
    unsigned long long* set_ptr = nullptr;
    uint32_t* set_fill = nullptr;
    uint64_t set_total = 0, set_cap = 0;
    int32_t* tile_min = nullptr;   // [groups][tiles][32]: per tile the best candidate score of a sample (optimal-set pass)
    size_t tile_min_cap = 0;
    bool tile_min_on = false;
    bool have_results = false, have_node_scores = false, have_set = false;
    float prep_ms = 0.f;
    uint64_t max_calls = 0;    // longest call list in the batch
    uint32_t cap_groups = 0;   // allocation capacities (scratch batches are reused by ub200_place_batch)
    uint64_t cap_calls = 0;
    size_t node_scores_cap = 0;
};

namespace {

int ensure_events(ub200_mat* M, int need) {
    while ((int)M->ev.size() < need) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        M->ev.push_back(e);
    }
    return 0;
}

int span_begin(ub200_mat* M, int kind) {
    int rc = ensure_events(M, M->ev_used + 2);
    if (rc) return rc;
    CU(cudaEventRecord(M->ev[M->ev_used], M->stream));
    M->spans.push_back({M->ev_used, M->ev_used + 1, kind});
    M->ev_used += 2;
    return 0;
}
int span_end(ub200_mat* M) {
    CU(cudaEventRecord(M->ev[M->spans.back().b], M->stream));
    return 0;
}

// exclusive prefix of num_best over the batch (one block: per-thread chunks, then a scan of the chunk sums)
__global__ void k_prefix_numbest(const ub200_placement* r, uint32_t n, unsigned long long* ptr) {
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
    unsigned long long acc = 0;
    for (uint32_t i = lo; i < hi; i++) acc += r[i].num_best;
    part[t] = acc;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        const unsigned long long v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    acc = t ? part[t - 1] : 0;
    for (uint32_t i = lo; i < hi; i++) { ptr[i] = acc; acc += r[i].num_best; }
    if (t == blockDim.x - 1) ptr[n] = part[t];
}

template <int MODE>
int launch_score(ub200_mat* M, ub200_samples* S, uint32_t group0, uint32_t ngroups, bool smem_bitmap) {
    using namespace ub200;
    ScoreParams p;
    p.mutw = M->mutw; p.hdr = M->hdr; p.row32 = M->row32;
    p.tile_start = M->tile_start; p.anc_ptr = M->anc_ptr; p.anc = M->anc;
    p.n_nodes = M->n; p.n_tiles = M->n_tiles; p.L = M->L;
    p.bitmap_words = S->bitmap_words; p.bitmap = S->bitmap; p.tab = S->tab; p.base = S->base; p.gbest = S->gbest;
    p.n_samples = S->n_samples; p.group0 = group0; p.ngroups = ngroups;
    p.part_key = S->part_key; p.part_cnt = S->part_cnt;
    p.gstack = M->gstack; p.gstack_levels = M->gstack_levels;
    p.node_scores = S->node_scores; p.target_rel = S->best_rel;
    p.set_out = S->set_out; p.set_ptr = S->set_ptr; p.set_fill = S->set_fill;
    const uint32_t grid = std::max<uint32_t>(ngroups, (M->grid / ngroups) * ngroups);
    const uint32_t bm_bytes = smem_bitmap ? ((S->bitmap_words * 4u + 127u) & ~127u) : 0u;
    const size_t smem = bm_bytes + (size_t)kWarpsPerCta * kWarpSmemBytes;
    if (smem_bitmap) {
        auto k = k_score<MODE, true>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kThreads, smem, M->stream>>>(p);
    } else {
        auto k = k_score<MODE, false>;
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, kThreads, smem, M->stream>>>(p);
    }
    CU(cudaGetLastError());
    return 0;
}

// Upload the k_score layout the first time a launch needs it.
int ensure_v1(ub200_mat* M) {
    if (M->v1_resident) return 0;
    auto& d = M->d;
    int rc = 0;
    auto guard = [&](int r) { if (r && !rc) rc = r; };
    guard(dev_upload(&M->mutw, d.mutw.data(), d.mutw.size(), M->stream));
    guard(dev_upload(&M->hdr, d.hdr.data(), d.hdr.size(), M->stream));
    guard(dev_upload(&M->row32, d.row32.data(), d.row32.size(), M->stream));
    guard(dev_upload(&M->tile_start, d.tile_start.data(), d.tile_start.size(), M->stream));
    guard(dev_upload(&M->anc_ptr, d.anc_ptr.data(), d.anc_ptr.size(), M->stream));
    guard(dev_upload(&M->anc, d.anc.data(), d.anc.size(), M->stream));
    auto drop_v1 = [&]() {   // a failed upload must not leave half a layout behind (the next call starts over)
        cudaFree(M->mutw); cudaFree(M->hdr); cudaFree(M->row32); cudaFree(M->tile_start); cudaFree(M->anc_ptr);
        cudaFree(M->anc); cudaFree(M->gstack);
        M->mutw = nullptr; M->hdr = nullptr; M->row32 = M->tile_start = M->anc_ptr = M->anc = nullptr; M->gstack = nullptr;
        M->gstack_levels = 0;
    };
    if (rc) { drop_v1(); return rc; }
    uint64_t v1_bytes = d.mutw.size() * 4 + d.hdr.size() * 16 + d.row32.size() * 4 + d.tile_start.size() * 4 +
                        d.anc_ptr.size() * 4 + d.anc.size() * 4;
    if (d.max_level + 1 > (uint32_t)ub200::kStackDepth) {
        M->gstack_levels = d.max_level + 1 - ub200::kStackDepth;
        const size_t bytes = (size_t)M->grid * ub200::kWarpsPerCta * M->gstack_levels * 32 * sizeof(int32_t);
        cudaError_t e = cudaMalloc((void**)&M->gstack, bytes);
        if (e != cudaSuccess) { drop_v1(); return fail((int)e, std::string("cudaMalloc spill stack: ") + cudaGetErrorString(e)); }
        v1_bytes += bytes;
    }
    {
        cudaError_t e = cudaStreamSynchronize(M->stream);
        if (e != cudaSuccess) { drop_v1(); return fail((int)e, std::string("upload: ") + cudaGetErrorString(e)); }
    }
    M->device_bytes += v1_bytes;
    M->v1_resident = true;
    return 0;
}

// Consumers per scanner (NC) for a pass of `ng` sample groups: the scan of the mutation stream is shared by NC
// groups, so wide passes want NC = 3 and a single group NC = 1.
uint32_t pick_consumers(const ub200_mat* M, uint32_t ng) {
    uint32_t nc = M->consumers;
    if (const char* e = getenv("UB200_NC")) nc = (uint32_t)atoi(e);
    if (nc == 0) nc = ng >= 3 ? 3u : ng;
    return std::max(1u, std::min(nc, std::min(ng, 3u)));
}

// Streaming best-placement kernel (score_kernel4.cuh): one CTA of scanner + NC consumer warp units per SM.
template <int NC>
int launch_score4_nc(ub200_mat* M, ub200_samples* S, const ub200::Score4Params& p, uint32_t grid, int mode) {
    using namespace ub200;
    using C = Cfg4<NC>;
    const uint32_t bm_need = (S->bitmap_words * 4u + 127u) & ~127u;
    const bool smem_bitmap = C::kFixed + bm_need <= kSmemLimit4;
    const size_t smem = C::kFixed + (smem_bitmap ? bm_need : 0u);
    auto go = [&](auto k) -> int {
        CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, C::kThreads, smem, M->stream>>>(p);
        return 0;
    };
    int rc;
    if (mode == kMode4NodeScores) {
        if constexpr (NC == 1) {
            if (M->d.narrow3) rc = smem_bitmap ? go(k_score4<1, true, kMode4NodeScores, true>) : go(k_score4<1, false, kMode4NodeScores, true>);
            else rc = smem_bitmap ? go(k_score4<1, true, kMode4NodeScores, false>) : go(k_score4<1, false, kMode4NodeScores, false>);
        } else {
            return fail(UB200_E_ARG, "per-node scores run one sample group per scan");
        }
    } else {
        // best placement (with the tile notes when the optimal sets follow) or the collect pass
        const bool notes = mode == kMode4Best && p.tile_min != nullptr;
        auto pick = [&](auto best, auto best_notes, auto collect) { return mode == kMode4Collect ? go(collect) : (notes ? go(best_notes) : go(best)); };
        if (M->d.narrow3) {
            if (smem_bitmap) rc = pick(k_score4<NC, true, kMode4Best, true>, k_score4<NC, true, kMode4BestNotes, true>, k_score4<NC, true, kMode4Collect, true>);
            else rc = pick(k_score4<NC, false, kMode4Best, true>, k_score4<NC, false, kMode4BestNotes, true>, k_score4<NC, false, kMode4Collect, true>);
        } else {
            if (smem_bitmap) rc = pick(k_score4<NC, true, kMode4Best, false>, k_score4<NC, true, kMode4BestNotes, false>, k_score4<NC, true, kMode4Collect, false>);
            else rc = pick(k_score4<NC, false, kMode4Best, false>, k_score4<NC, false, kMode4BestNotes, false>, k_score4<NC, false, kMode4Collect, false>);
        }
    }
    if (rc) return rc;
    CU(cudaGetLastError());
    return 0;
}

// One pass: groups [group0, group0 + ngroups) of the batch, `nc` groups per scanner; bm0 = index of the pass's first
// union bitmap.  *wpg_out = partial rows per group (CTAs per scan group) for k_reduce.
int launch_score4(ub200_mat* M, ub200_samples* S, uint32_t group0, uint32_t ngroups, uint32_t nc, uint32_t bm0,
                  uint32_t* wpg_out, int mode = 0, uint32_t part_stride = 0) {
    using namespace ub200;
    Score4Params p;
    p.stream = M->mstream; p.hdr = M->hdr3; p.tiekey = M->tiekey;
    p.tile_start = M->tile3_start; p.tile_w0 = M->tile3_w0; p.tile_lvl = M->tile3_lvl; p.tile_sseg = M->tile3_sseg;
    p.seed_end = M->seed_end; p.blk_words = M->blk_words; p.blk_rec = reinterpret_cast<const uint4*>(M->blk_rec);
    p.n_nodes = M->n; p.n_tiles = M->n_tiles3; p.L = M->L;
    p.bitmap_words = S->bitmap_words; p.bitmap = S->bitmap + (size_t)bm0 * S->bitmap_words; p.tab = S->tab;
    p.gbest = S->gbest;
    p.n_samples = S->n_samples; p.group0 = group0; p.ngroups = ngroups; p.nsg = (ngroups + nc - 1) / nc;
    p.part_key = S->part_key; p.part_cnt = S->part_cnt;
    p.gstack = M->gstack3; p.gstack_levels = M->gstack3_levels;
    p.target_rel = S->best_rel; p.set_out = S->set_out; p.set_ptr = S->set_ptr; p.set_fill = S->set_fill;
    p.base = S->base; p.node_scores = S->node_scores; p.tile_min = S->tile_min_on ? S->tile_min : nullptr;
    // every launch of a place call takes its own block of 16 tile counters (one memset per 256 launches, not per launch)
    if (S->launch_idx % 256u == 0) CU(cudaMemsetAsync(S->tile_counter, 0, 4096 * 4, M->stream));
    p.tile_counter = S->tile_counter + 16u * (S->launch_idx % 256u);
    S->launch_idx++;
    p.prof = reinterpret_cast<unsigned long long*>(S->tile_counter + 4096);   // 16 counters behind the tile counters
    const uint32_t grid = std::max<uint32_t>(p.nsg, ((uint32_t)M->num_sms / p.nsg) * p.nsg);
    *wpg_out = grid / p.nsg;
    p.part_group0 = part_stride ? group0 : 0u;
    p.part_stride = part_stride ? part_stride : grid / p.nsg;
    if (nc == 1) return launch_score4_nc<1>(M, S, p, grid, mode);
    if (nc == 2) return launch_score4_nc<2>(M, S, p, grid, mode);
    return launch_score4_nc<3>(M, S, p, grid, mode);
}

// Build the per-group position bitmap + position-major cost table + per-sample base count on the device.
// Part of every place call (it is sample-side work of the hot path), timed as "prep".
int run_prep(ub200_mat* M, ub200_samples* S, uint32_t pass_groups, uint32_t nc) {
    cudaStream_t st = M->stream;
    int rc = span_begin(M, 0);
    if (rc) return rc;
    CU(cudaMemsetAsync(S->bitmap, 0, (size_t)S->n_groups * S->bitmap_words * 4, st));
    CU(cudaMemsetAsync(S->tab, 0, (size_t)S->n_groups * M->L * 32, st));
    CU(cudaMemsetAsync(S->base, 0, (size_t)S->n_groups * 32 * 4, st));
    if (S->n_calls) {
        ub200::PrepParams pp;
        pp.calls = S->calls; pp.sample_ptr = S->sample_ptr; pp.call_sample = S->call_sample;
        pp.n_calls = S->n_calls; pp.L = M->L; pp.bitmap_words = S->bitmap_words;
        pp.bitmap = S->bitmap; pp.tab = S->tab; pp.base = S->base;
        pp.pass_groups = pass_groups; pp.nc = nc; pp.nsg_per_pass = (pass_groups + nc - 1) / nc;
        const uint32_t blocks = (uint32_t)((S->n_calls + 255) / 256);
        ub200::k_prep_scatter<<<blocks, 256, 0, st>>>(pp);
        CU(cudaGetLastError());
    }
    ub200::k_prep_bound<<<(S->n_samples + 255) / 256, 256, 0, st>>>(S->sample_ptr, S->base, S->n_samples,
                                                                     M->d.root_init_extra, S->gbest);
    CU(cudaGetLastError());
    return span_end(M);
}

}  // namespace

extern "C" {

const char* ub200_last_error(void) { return g_err.c_str(); }
int ub200_abi_version(void) { return UB200_ABI_VERSION; }

int ub200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void ub200_mat_destroy(ub200_mat* M) {
    if (!M) return;
    cudaSetDevice(M->device);
    cudaFree(M->mutw); cudaFree(M->hdr); cudaFree(M->row32); cudaFree(M->tile_start); cudaFree(M->anc_ptr);
    cudaFree(M->anc); cudaFree(M->key_to_node); cudaFree(M->tie_index); cudaFree(M->num_leaves);
    cudaFree(M->gstack);
    cudaFree(M->mstream); cudaFree(M->hdr3); cudaFree(M->tiekey); cudaFree(M->tile3_start); cudaFree(M->tile3_w0);
    cudaFree(M->tile3_lvl); cudaFree(M->tile3_sseg); cudaFree(M->seed_end); cudaFree(M->blk_words); cudaFree(M->blk_rec); cudaFree(M->gstack3);
    if (M->scratch) ub200_samples_free(M->scratch);
    for (auto e : M->ev) cudaEventDestroy(e);
    if (M->own_stream) cudaStreamDestroy(M->own_stream);
    delete M;
}

// Tile workers of the scoring kernels on a device (k_score: 2 CTAs x 8 warps per SM, k_score4: 16 units per SM)
static int device_workers(int device, int* num_sms, uint32_t* workers) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    *num_sms = prop.multiProcessorCount;
    const uint32_t warps3 = (uint32_t)std::max(*num_sms, 8) * 16u;
    *workers = std::max<uint32_t>((uint32_t)*num_sms * 2u * ub200::kWarpsPerCta, warps3);
    return 0;
}

static int derive_for(const ub200_flat_mat* flat, uint32_t workers, std::shared_ptr<ub200::Derived>& out) {
    auto d = std::make_shared<ub200::Derived>();
    std::string err;
    const char* mt = getenv("UB200_MIN_TILE");   // test hook: cut small trees into many tiles
    const char* tw = getenv("UB200_TILES_PER_WORKER");
    int rc = ub200::derive(*flat, workers * (tw ? (uint32_t)atoi(tw) : 3u), *d, err, mt ? (uint32_t)atoi(mt) : 0u);
    if (rc != UB200_OK) return fail(rc, err);
    out = d;
    return UB200_OK;
}

// Stage a derived tree on one device.  The streamed host arrays (d.stream, d.hdr3) are left alone: the caller frees them
// once every replica is resident.
static int mat_upload(std::shared_ptr<ub200::Derived> dp, int device, ub200_mat** out) {
    *out = nullptr;
    CU(cudaSetDevice(device));
    int num_sms = 0;
    uint32_t workers = 0;
    { int rc = device_workers(device, &num_sms, &workers); if (rc) return rc; }
    auto* M = new ub200_mat(std::move(dp));
    M->device = device;
    M->num_sms = num_sms;
    M->grid = (uint32_t)M->num_sms * 2u;
    auto& d = M->d;
    M->n = d.n; M->m = d.m; M->L = d.L; M->n_tiles = (uint32_t)d.tile_start.size() - 1;
    int rc = 0;
    auto guard = [&](int r) { if (r && !rc) rc = r; };
    {
        cudaError_t e = cudaStreamCreateWithFlags(&M->own_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            ub200_mat_destroy(M);
            return fail((int)e, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e));
        }
    }
    M->stream = M->own_stream;
    guard(dev_upload(&M->key_to_node, d.key_to_node.data(), d.key_to_node.size(), M->stream));
    guard(dev_upload(&M->tie_index, d.tie_index.data(), d.tie_index.size(), M->stream));
    guard(dev_upload(&M->num_leaves, d.num_leaves.data(), d.num_leaves.size(), M->stream));
    M->device_bytes = (uint64_t)d.n * 12;
    if (d.have3) {
        M->n_tiles3 = (uint32_t)d.tile3_start.size() - 1;
        guard(dev_upload(&M->mstream, d.stream.data(), d.stream.size(), M->stream));
        guard(dev_upload(&M->hdr3, d.hdr3.data(), d.hdr3.size(), M->stream));
        guard(dev_upload(&M->tiekey, d.tiekey.data(), d.tiekey.size(), M->stream));
        guard(dev_upload(&M->tile3_start, d.tile3_start.data(), d.tile3_start.size(), M->stream));
        guard(dev_upload(&M->tile3_w0, d.tile3_w0.data(), d.tile3_w0.size(), M->stream));
        guard(dev_upload(&M->tile3_lvl, d.tile3_lvl.data(), d.tile3_lvl.size(), M->stream));
        guard(dev_upload(&M->tile3_sseg, d.tile3_sseg.data(), d.tile3_sseg.size(), M->stream));
        guard(dev_upload(&M->seed_end, d.seed_end.data(), d.seed_end.size(), M->stream));
        guard(dev_upload(&M->blk_words, d.blk_words.data(), d.blk_words.size(), M->stream));
        guard(dev_upload(&M->blk_rec, d.blk_rec.data(), d.blk_rec.size(), M->stream));
        M->device_bytes += d.stream.size() * 4 + d.hdr3.size() * 16 + d.tiekey.size() * 4 +
                           (d.tile3_start.size() * 4 + d.seed_end.size() + d.blk_words.size() + d.blk_rec.size()) * 4;
        // spill rows of the consumers' level stacks: at most 24 consumers per SM (NC = 3), 32 levels in shared memory
        constexpr uint32_t kMinStack = (uint32_t)ub200::Cfg4<3>::kStack;
        if (!rc && d.max_level + 1 > kMinStack) {
            M->gstack3_levels = d.max_level + 1 - kMinStack;
            const size_t bytes = (size_t)std::max(M->num_sms, 8) * 24u * M->gstack3_levels * 32 * sizeof(int32_t);
            if (bytes > (size_t)16 << 30) {
                rc = fail(UB200_E_LIMIT, "tree too deep for the spill stack (" + std::to_string(d.max_level) + " levels)");
            } else {
                cudaError_t e = cudaMalloc((void**)&M->gstack3, bytes);
                if (e != cudaSuccess) rc = fail((int)e, std::string("cudaMalloc spill stack: ") + cudaGetErrorString(e));
                M->device_bytes += bytes;
            }
        }
    }
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(M->stream);
        if (e != cudaSuccess) rc = fail((int)e, std::string("upload: ") + cudaGetErrorString(e));
    }
    if (rc) { ub200_mat_destroy(M); return rc; }
    *out = M;
    return UB200_OK;
}

static void drop_streamed_host_arrays(ub200::Derived& d) {
    ub200::StreamVec().swap(d.stream);
    ub200::RawVec<ub200::NodeHdr>().swap(d.hdr3);
}

int ub200_mat_create(const ub200_flat_mat* flat, int device, ub200_mat** out) {
    if (!flat || !out) return fail(UB200_E_ARG, "ub200_mat_create: NULL argument");
    *out = nullptr;
    int ndev = ub200_device_count();
    if (ndev <= 0) return fail(UB200_E_NO_DEVICE, "ub200_mat_create: no CUDA device (there is no CPU path)");
    if (device < 0 || device >= ndev) return fail(UB200_E_ARG, "ub200_mat_create: bad device ordinal");
    int num_sms = 0;
    uint32_t workers = 0;
    { int rc = device_workers(device, &num_sms, &workers); if (rc) return rc; }
    std::shared_ptr<ub200::Derived> dp;
    { int rc = derive_for(flat, workers, dp); if (rc) return rc; }
    int rc = mat_upload(dp, device, out);
    if (rc) return rc;
    // host copies of the streamed arrays are no longer needed (the k_score layout stays for ensure_v1)
    drop_streamed_host_arrays(*dp);
    return UB200_OK;
}

int ub200_mat_info_get(const ub200_mat* M, ub200_mat_info* o) {
    if (!M || !o) return fail(UB200_E_ARG, "ub200_mat_info_get: NULL argument");
    o->n_nodes = M->n; o->max_level = M->d.max_level; o->n_mutations = M->m; o->genome_len = M->L;
    o->n_tiles = M->d.have3 ? M->n_tiles3 : M->n_tiles; o->device_bytes = M->device_bytes;
    o->algorithmic_bytes = 4ull * M->m + 16ull * M->n;
    o->device = M->device; o->reserved = 0;
    return UB200_OK;
}

int ub200_mat_node_arrays(const ub200_mat* M, uint32_t* bfs_index, uint32_t* num_leaves, uint32_t* level) {
    if (!M) return fail(UB200_E_ARG, "ub200_mat_node_arrays: NULL handle");
    if (bfs_index) memcpy(bfs_index, M->d.tie_index.data(), sizeof(uint32_t) * M->n);
    if (num_leaves) memcpy(num_leaves, M->d.num_leaves.data(), sizeof(uint32_t) * M->n);
    if (level) memcpy(level, M->d.level.data(), sizeof(uint32_t) * M->n);
    return UB200_OK;
}

int ub200_mat_set_pass_samples(ub200_mat* M, uint32_t spp) {
    if (!M) return fail(UB200_E_ARG, "NULL handle");
    if (spp == 0) spp = 32;
    if (spp % 32 || spp > 768) return fail(UB200_E_ARG, "samples per pass must be a multiple of 32, at most 768");
    M->pass_groups = spp / 32;
    return UB200_OK;
}

int ub200_mat_set_scan_sharing(ub200_mat* M, uint32_t nc) {
    if (!M) return fail(UB200_E_ARG, "NULL handle");
    if (nc > 3) return fail(UB200_E_ARG, "groups per scan must be 0 (auto), 1, 2 or 3");
    M->consumers = nc;
    return UB200_OK;
}

int ub200_mat_set_stream(ub200_mat* M, void* s) {
    if (!M) return fail(UB200_E_ARG, "NULL handle");
    M->stream = s ? (cudaStream_t)s : M->own_stream;
    return UB200_OK;
}

int ub200_mat_synchronize(ub200_mat* M) {
    if (!M) return fail(UB200_E_ARG, "NULL handle");
    CU(cudaSetDevice(M->device));
    CU(cudaStreamSynchronize(M->stream));
    return UB200_OK;
}

void ub200_samples_free(ub200_samples* S) {
    if (!S) return;
    cudaSetDevice(S->mat->device);
    cudaFree(S->calls); cudaFree(S->sample_ptr); cudaFree(S->call_sample); cudaFree(S->bitmap); cudaFree(S->tab);
    cudaFree(S->base); cudaFree(S->gbest); cudaFree(S->tile_counter); cudaFree(S->results); cudaFree(S->best_rel); cudaFree(S->part_key); cudaFree(S->part_cnt);
    cudaFree(S->node_scores); cudaFree(S->set_out); cudaFree(S->set_ptr); cudaFree(S->set_fill); cudaFree(S->tile_min);
    delete S;
}

// Validate a host batch and copy it into `S` (allocating or growing its device buffers as needed).
static int samples_fill(ub200_mat* M, ub200_samples* S, uint32_t n_samples, const uint64_t* sample_ptr,
                        const ub200_mutation* calls) {
    const uint64_t n_calls = sample_ptr[n_samples];
    if (n_calls && !calls) return fail(UB200_E_ARG, "ub200_samples_upload: NULL calls");
    if (sample_ptr[0] != 0) return fail(UB200_E_ARG, "ub200_samples_upload: sample_ptr[0] != 0");
    // ---- host validation (the reference's merge-scan assumes sorted calls, usher_mapper.cpp:191,239)
    std::vector<uint32_t> owner(n_calls);
    uint64_t max_calls = 0;
    for (uint32_t s = 0; s < n_samples; s++) {
        if (sample_ptr[s + 1] < sample_ptr[s]) return fail(UB200_E_ARG, "sample_ptr not monotone");
        max_calls = std::max<uint64_t>(max_calls, sample_ptr[s + 1] - sample_ptr[s]);
        int64_t last = -1;
        for (uint64_t k = sample_ptr[s]; k < sample_ptr[s + 1]; k++) {
            const ub200_mutation& c = calls[k];
            if (c.position < 0 || (int64_t)c.position <= last)
                return fail(UB200_E_SAMPLE_ORDER, "sample " + std::to_string(s) +
                                                      ": calls must be strictly position-increasing and >= 0");
            last = c.position;
            if ((c.mut_nuc & 15u) == 0 || c.mut_nuc > 15u)
                return fail(UB200_E_ARG, "sample " + std::to_string(s) + ": mut_nuc must be a non-empty 4-bit set");
            if (ub200::nuc_code(c.ref_nuc) < 0)
                return fail(UB200_E_NOT_ONE_HOT, "sample " + std::to_string(s) + ": ref_nuc must be one-hot");
            if ((uint32_t)c.position < M->L && M->d.ref_of[c.position] && M->d.ref_of[c.position] != c.ref_nuc)
                return fail(UB200_E_ARG, "sample " + std::to_string(s) + " position " + std::to_string(c.position) +
                                             ": reference allele differs from the tree's");
            owner[k] = s;
        }
    }
    CU(cudaSetDevice(M->device));
    const uint32_t n_groups = (n_samples + 31) / 32;
    S->bitmap_words = ((M->L + 1 + 31) / 32 + 3) & ~3u;   // bit L exists and stays clear (stream pad words)
    if ((size_t)n_groups * M->L * 32 > ((size_t)24 << 30))
        return fail(UB200_E_LIMIT, "sample batch too large for one resident table; split the batch");
    auto alloc = [&](void** p, size_t bytes) -> int {
        cudaFree(*p);
        *p = nullptr;
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return fail((int)e, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        return 0;
    };
    int rc = 0;
    if (n_groups > S->cap_groups) {
        const uint32_t g = n_groups;
        if (!rc) rc = alloc((void**)&S->bitmap, (size_t)g * S->bitmap_words * 4);
        if (!rc) rc = alloc((void**)&S->tab, (size_t)g * M->L * 32);
        if (!rc) rc = alloc((void**)&S->base, (size_t)g * 32 * 4);
        if (!rc) rc = alloc((void**)&S->gbest, (size_t)g * 32 * 4);
        if (!rc && !S->tile_counter) {
            rc = alloc((void**)&S->tile_counter, (4096 + 32) * 4);   // 256 blocks of 16 counters + 16 profile counters
            if (!rc) CU(cudaMemset(S->tile_counter, 0, (4096 + 32) * 4));
        }
        if (!rc) rc = alloc((void**)&S->results, (size_t)g * 32 * sizeof(ub200_placement));
        if (!rc) rc = alloc((void**)&S->best_rel, (size_t)g * 32 * 4);
        if (!rc) rc = alloc((void**)&S->sample_ptr, ((size_t)g * 32 + 1) * 8);
        if (!rc) rc = alloc((void**)&S->set_ptr, ((size_t)g * 32 + 1) * 8);
        if (!rc) rc = alloc((void**)&S->set_fill, (size_t)g * 32 * 4);
        // partial bests: one row of 32 lanes per (group, CTA serving it); all groups of the batch keep their rows until
        // ONE reduction at the end of the call (up to 256 MB; larger batches reduce pass by pass)
        const size_t per_pass_rows = (size_t)std::max<uint32_t>(M->grid, 8u) * 3u + 32u;
        const size_t all_rows = (size_t)g * (size_t)std::max(M->num_sms, 8);
        const size_t part_rows = std::max(per_pass_rows, all_rows * 32 * 12 <= ((size_t)256 << 20) ? all_rows : 0);
        if (!rc) rc = alloc((void**)&S->part_key, part_rows * 32 * 8);
        if (!rc) rc = alloc((void**)&S->part_cnt, part_rows * 32 * 4);
        S->part_rows = part_rows;
        if (rc) { S->cap_groups = 0; return rc; }
        S->cap_groups = g;
    }
    if (n_calls > S->cap_calls) {
        const uint64_t c = n_calls + n_calls / 4;
        if (!rc) rc = alloc((void**)&S->calls, (size_t)c * sizeof(ub200_mutation));
        if (!rc) rc = alloc((void**)&S->call_sample, (size_t)c * 4);
        if (rc) { S->cap_calls = 0; return rc; }
        S->cap_calls = c;
    }
    S->n_samples = n_samples;
    S->n_groups = n_groups;
    S->n_calls = n_calls;
    S->max_calls = max_calls;
    S->have_results = S->have_node_scores = S->have_set = false;
    cudaStream_t st = M->stream;
    if (n_calls) {
        CU(cudaMemcpyAsync(S->calls, calls, n_calls * sizeof(ub200_mutation), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(S->call_sample, owner.data(), n_calls * 4, cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpyAsync(S->sample_ptr, sample_ptr, ((size_t)n_samples + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));   // `owner` is a local; the caller's buffers may be reused after return
    return UB200_OK;
}

int ub200_samples_upload(ub200_mat* M, uint32_t n_samples, const uint64_t* sample_ptr, const ub200_mutation* calls,
                         ub200_samples** out) {
    if (!M || !out || !sample_ptr || n_samples == 0) return fail(UB200_E_ARG, "ub200_samples_upload: bad argument");
    *out = nullptr;
    auto* S = new ub200_samples();
    S->mat = M;
    int rc = samples_fill(M, S, n_samples, sample_ptr, calls);
    if (rc) { ub200_samples_free(S); return rc; }
    *out = S;
    return UB200_OK;
}

int ub200_results_copy_device(ub200_mat* M, ub200_samples* S, void* dst_dev) {
    if (!M || !S || !dst_dev || !S->have_results) return fail(UB200_E_ARG, "ub200_results_copy_device: nothing to copy");
    CU(cudaSetDevice(M->device));
    CU(cudaMemcpyAsync(dst_dev, S->results, (size_t)S->n_samples * sizeof(ub200_placement), cudaMemcpyDeviceToDevice,
                       M->stream));
    return UB200_OK;
}

static int place_resident_impl(ub200_mat* M, ub200_samples* S, uint32_t flags, int sync);

int ub200_place_resident(ub200_mat* M, ub200_samples* S, uint32_t flags, int sync) {
    if (!M || !S || S->mat != M) return fail(UB200_E_ARG, "ub200_place_resident: bad handles");
    const int rc = place_resident_impl(M, S, flags, sync);
    if (rc) {   // a failed call leaves no half-recorded timing spans behind (ub200_last_timing would read them)
        M->spans.clear(); M->ev_used = 0; M->last = {};
        S->have_results = S->have_node_scores = S->have_set = false;
    }
    return rc;
}

static int place_resident_impl(ub200_mat* M, ub200_samples* S, uint32_t flags, int sync) {
    CU(cudaSetDevice(M->device));
    const bool smem_bitmap = S->bitmap_words * 4u <= ub200::kMaxSmemBitmapBytes;
    // the streaming kernel packs per-(node,sample) deltas in 10-bit fields and path corrections in int16
    const char* force = getenv("UB200_KERNEL");
    const bool use_v3 = !(force && force[0] == '1') && M->d.have3 && M->d.max_row <= ub200::kMaxRowV4 &&
                        S->max_calls <= ub200::kMaxCallsV4;
    if (!use_v3) { int rc = ensure_v1(M); if (rc) return rc; }   // the first-generation layout, only when it is needed
    const uint32_t NG = M->pass_groups;
    // optimal sets: pass 1 notes per tile the best candidate score of every sample, so that the collect pass only scans
    // the few tiles that hold an optimal node (up to 1 GB of notes; beyond that the collect pass scans every tile)
    S->tile_min_on = false;
    if (use_v3 && (flags & UB200_WANT_BEST_SET)) {
        const size_t bytes = (size_t)S->n_groups * M->n_tiles3 * 32 * sizeof(int32_t);
        if (bytes <= ((size_t)1 << 30)) {
            if (bytes > S->tile_min_cap) {
                cudaFree(S->tile_min); S->tile_min = nullptr; S->tile_min_cap = 0;
                CU(cudaMalloc((void**)&S->tile_min, bytes));
                S->tile_min_cap = bytes;
            }
            CU(cudaMemsetAsync(S->tile_min, 0x7f, bytes, M->stream));
            S->tile_min_on = true;
        }
    }
    M->spans.clear(); M->ev_used = 0;
    M->last = {};
    // k_score (per-node scores, fallback) indexes one bitmap per group: no shared scans then
    const uint32_t nc = (use_v3 && !(flags & UB200_WANT_NODE_SCORES)) ? pick_consumers(M, std::min(NG, S->n_groups)) : 1u;
    const uint32_t nsgpp = (NG + nc - 1) / nc;   // union bitmaps per pass
    { int rc = run_prep(M, S, NG, nc); if (rc) return rc; }
    M->last.total_launches += 5;
    // k_score4 keeps the partial rows of every group and reduces ONCE after the last pass (one launch per distinct
    // number of CTAs per group: the full passes, and a shorter last one)
    const uint32_t R = (uint32_t)std::max(M->num_sms, 8);
    const bool deferred = use_v3 && (size_t)S->n_groups * R <= S->part_rows;
    S->launch_idx = 0;
    auto reduce = [&](uint32_t g0, uint32_t ng, uint32_t wpg, uint32_t stride, uint32_t part_g0) -> int {
        ub200::ReduceParams rp;
        rp.part_key = S->part_key; rp.part_cnt = S->part_cnt; rp.wpg = wpg; rp.stride = stride; rp.part_group0 = part_g0;
        rp.group0 = g0; rp.n_samples = S->n_samples; rp.base = S->base; rp.key_to_node = M->key_to_node;
        rp.tie_index = M->tie_index; rp.num_leaves = M->num_leaves; rp.out = S->results; rp.best_rel = S->best_rel;
        int rc = span_begin(M, 2); if (rc) return rc;
        ub200::k_reduce<<<ng, 256, 0, M->stream>>>(rp);
        CU(cudaGetLastError());
        rc = span_end(M); if (rc) return rc;
        M->last.total_launches += 1;
        return 0;
    };
    uint32_t wpg_first = 0, wpg_last = 0, last_g0 = 0;
    for (uint32_t g0 = 0; g0 < S->n_groups; g0 += NG) {
        const uint32_t ng = std::min(NG, S->n_groups - g0);
        int rc = span_begin(M, 1); if (rc) return rc;
        uint32_t wpg = std::max<uint32_t>(ng, (M->grid / ng) * ng) / ng;
        if (use_v3) rc = launch_score4(M, S, g0, ng, nc, (g0 / NG) * nsgpp, &wpg, 0, deferred ? R : 0u);
        else rc = launch_score<ub200::kModeBest>(M, S, g0, ng, smem_bitmap);
        if (rc) return rc;
        rc = span_end(M); if (rc) return rc;
        if (!deferred) { rc = reduce(g0, ng, wpg, wpg, 0); if (rc) return rc; }
        if (g0 == 0) wpg_first = wpg;
        wpg_last = wpg; last_g0 = g0;
        M->last.score_launches++;
        M->last.total_launches += 1;
        M->last.score_bytes += 4ull * M->m + 16ull * M->n;
    }
    if (deferred) {
        int rc = 0;
        if (wpg_last == wpg_first) rc = reduce(0, S->n_groups, wpg_first, R, 0);
        else {
            rc = reduce(0, last_g0, wpg_first, R, 0);
            if (!rc) rc = reduce(last_g0, S->n_groups - last_g0, wpg_last, R, last_g0);
        }
        if (rc) return rc;
    }
    S->have_results = true;
    if (flags & UB200_WANT_NODE_SCORES) {
        const size_t bytes = (size_t)S->n_samples * M->n * sizeof(int32_t);
        if (bytes > S->node_scores_cap) {
            cudaFree(S->node_scores); S->node_scores = nullptr; S->node_scores_cap = 0;
            CU(cudaMalloc((void**)&S->node_scores, bytes));
            S->node_scores_cap = bytes;
        }
        for (uint32_t g0 = 0; g0 < S->n_groups; g0 += NG) {
            const uint32_t ng = std::min(NG, S->n_groups - g0);
            uint32_t wpg_unused = 0;
            int rc = use_v3 ? launch_score4(M, S, g0, ng, 1, (g0 / NG) * nsgpp, &wpg_unused, ub200::kMode4NodeScores)
                            : launch_score<ub200::kModeNodeScores>(M, S, g0, ng, smem_bitmap);
            if (rc) return rc;
            M->last.total_launches++;
        }
        S->have_node_scores = true;
    }
    if (flags & UB200_WANT_BEST_SET) {
        k_prefix_numbest<<<1, 1024, 0, M->stream>>>(S->results, S->n_samples, S->set_ptr);
        CU(cudaGetLastError());
        unsigned long long total = 0;
        CU(cudaMemcpyAsync(&total, S->set_ptr + S->n_samples, 8, cudaMemcpyDeviceToHost, M->stream));
        CU(cudaStreamSynchronize(M->stream));
        // the size of the optimal sets is only known now: this request synchronises the stream (documented in the
        // header); the buffer is kept and only ever grows
        if (total > S->set_cap || !S->set_out) {
            cudaFree(S->set_out); S->set_out = nullptr; S->set_cap = 0;
            const size_t cap = std::max<size_t>((size_t)total + (size_t)total / 4, 1024);
            CU(cudaMalloc((void**)&S->set_out, cap * 4));
            S->set_cap = cap;
        }
        S->set_total = total;
        CU(cudaMemsetAsync(S->set_fill, 0, (size_t)S->n_groups * 32 * 4, M->stream));
        for (uint32_t g0 = 0; g0 < S->n_groups; g0 += NG) {
            const uint32_t ng = std::min(NG, S->n_groups - g0);
            uint32_t wpg_unused = 0;
            int rc = use_v3 ? launch_score4(M, S, g0, ng, nc, (g0 / NG) * nsgpp, &wpg_unused, ub200::kMode4Collect)
                            : launch_score<ub200::kModeCollect>(M, S, g0, ng, smem_bitmap);
            if (rc) return rc;
            M->last.total_launches++;
        }
        M->last.total_launches++;
        S->have_set = true;
    }
    if (sync) CU(cudaStreamSynchronize(M->stream));
    return UB200_OK;
}

int ub200_last_timing(ub200_mat* M, ub200_timing* out) {
    if (!M || !out) return fail(UB200_E_ARG, "ub200_last_timing: NULL argument");
    CU(cudaSetDevice(M->device));
    CU(cudaStreamSynchronize(M->stream));
    if (M->spans.empty()) { *out = M->last; return UB200_OK; }   // already folded (ub200_place_batch)
    float sc = 0.f, rd = 0.f, pr = 0.f;
    for (auto& s : M->spans) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, M->ev[s.a], M->ev[s.b]));
        if (s.kind == 1) sc += ms; else if (s.kind == 2) rd += ms; else pr += ms;
    }
    M->last.prep_ms = pr;
    M->last.score_ms = sc;
    M->last.reduce_ms = rd;
    *out = M->last;
    return UB200_OK;
}

int ub200_results_download(ub200_mat* M, ub200_samples* S, ub200_placement* out) {
    if (!M || !S || !out || !S->have_results) return fail(UB200_E_ARG, "ub200_results_download: nothing to download");
    CU(cudaSetDevice(M->device));
    CU(cudaMemcpyAsync(out, S->results, (size_t)S->n_samples * sizeof(ub200_placement), cudaMemcpyDeviceToHost, M->stream));
    CU(cudaStreamSynchronize(M->stream));
    return UB200_OK;
}

// Developer hook (UB200_PROFILE builds): the 16 cycle counters the scoring kernels accumulated since the last call.
int ub200_debug_prof(ub200_mat* M, ub200_samples* S, unsigned long long* out16) {
    if (!M || !S || !out16) return fail(UB200_E_ARG, "ub200_debug_prof: NULL argument");
    CU(cudaSetDevice(M->device));
    CU(cudaStreamSynchronize(M->stream));
    CU(cudaMemcpy(out16, S->tile_counter + 4096, 128, cudaMemcpyDeviceToHost));
    CU(cudaMemset(S->tile_counter + 4096, 0, 128));
    return UB200_OK;
}

int ub200_results_device_ptr(ub200_samples* S, void** p, size_t* bytes) {
    if (!S || !p) return fail(UB200_E_ARG, "ub200_results_device_ptr: NULL argument");
    *p = S->results;
    if (bytes) *bytes = (size_t)S->n_samples * sizeof(ub200_placement);
    return UB200_OK;
}

int ub200_node_scores_download(ub200_mat* M, ub200_samples* S, int32_t* out) {
    if (!M || !S || !out || !S->have_node_scores) return fail(UB200_E_ARG, "node scores were not requested");
    CU(cudaSetDevice(M->device));
    CU(cudaMemcpyAsync(out, S->node_scores, (size_t)S->n_samples * M->n * 4, cudaMemcpyDeviceToHost, M->stream));
    CU(cudaStreamSynchronize(M->stream));
    return UB200_OK;
}

int ub200_best_set_download(ub200_mat* M, ub200_samples* S, uint32_t* best_set, uint64_t* best_set_ptr, uint64_t cap) {
    if (!M || !S || !best_set_ptr || !S->have_set) return fail(UB200_E_ARG, "best set was not requested");
    CU(cudaSetDevice(M->device));
    CU(cudaMemcpyAsync(best_set_ptr, S->set_ptr, ((size_t)S->n_samples + 1) * 8, cudaMemcpyDeviceToHost, M->stream));
    CU(cudaStreamSynchronize(M->stream));
    if (S->set_total > cap || !best_set) return fail(UB200_E_CAPACITY, "best_set capacity too small");
    CU(cudaMemcpyAsync(best_set, S->set_out, (size_t)S->set_total * 4, cudaMemcpyDeviceToHost, M->stream));
    CU(cudaStreamSynchronize(M->stream));
    // ascending DFS index within each sample (bit 31 carries node_has_unique)
    for (uint32_t s = 0; s < S->n_samples; s++)
        std::sort(best_set + best_set_ptr[s], best_set + best_set_ptr[s + 1],
                  [](uint32_t a, uint32_t b) { return (a & 0x7fffffffu) < (b & 0x7fffffffu); });
    return UB200_OK;
}

int ub200_place_batch(ub200_mat* M, uint32_t n_samples, const uint64_t* sample_ptr, const ub200_mutation* calls,
                      uint32_t flags, ub200_placement* out, int32_t* node_scores, uint32_t* best_set,
                      uint64_t* best_set_ptr, uint64_t best_set_cap) {
    if (!M || !out || !sample_ptr) return fail(UB200_E_ARG, "ub200_place_batch: NULL argument");
    if ((flags & UB200_WANT_NODE_SCORES) && !node_scores) return fail(UB200_E_ARG, "node_scores is NULL");
    if ((flags & UB200_WANT_BEST_SET) && !best_set_ptr) return fail(UB200_E_ARG, "best_set_ptr is NULL");
    if (n_samples == 0) { if (best_set_ptr) best_set_ptr[0] = 0; return UB200_OK; }
    // sub-batches bounded by the resident table size
    const size_t per_group = (size_t)M->L * 32;
    uint32_t max_groups = (uint32_t)std::min<size_t>(1u << 16, std::max<size_t>(1, ((size_t)4 << 30) / std::max<size_t>(per_group, 1)));
    uint32_t step = max_groups * 32;
    if (flags & UB200_WANT_NODE_SCORES) {
        // per-node scores are n_nodes * 4 bytes per sample on the device: keep a sub-batch's block under ~2 GB
        const uint64_t fit = ((uint64_t)2 << 30) / ((uint64_t)M->n * 4u);
        step = (uint32_t)std::max<uint64_t>(32, std::min<uint64_t>(step, fit / 32 * 32));
    }
    uint64_t set_off = 0;
    bool overflow = false;
    if (best_set_ptr) best_set_ptr[0] = 0;
    float prep = 0, sc = 0, rd = 0;
    ub200_timing acc = {};
    for (uint32_t s0 = 0; s0 < n_samples; s0 += step) {
        const uint32_t ns = std::min(step, n_samples - s0);
        std::vector<uint64_t> ptr(ns + 1);
        for (uint32_t i = 0; i <= ns; i++) ptr[i] = sample_ptr[s0 + i] - sample_ptr[s0];
        if (!M->scratch) { M->scratch = new ub200_samples(); M->scratch->mat = M; }
        ub200_samples* S = M->scratch;
        int rc = samples_fill(M, S, ns, ptr.data(), calls ? calls + sample_ptr[s0] : nullptr);
        if (rc) return rc;
        rc = ub200_place_resident(M, S, overflow ? (flags & ~UB200_WANT_BEST_SET) : flags, 0);
        if (!rc) rc = ub200_results_download(M, S, out + s0);
        if (!rc && (flags & UB200_WANT_NODE_SCORES)) rc = ub200_node_scores_download(M, S, node_scores + (size_t)s0 * M->n);
        if (!rc && (flags & UB200_WANT_BEST_SET) && !overflow) {
            std::vector<uint64_t> lp(ns + 1);
            rc = ub200_best_set_download(M, S, best_set ? best_set + set_off : nullptr, lp.data(),
                                         best_set_cap > set_off ? best_set_cap - set_off : 0);
            if (rc == UB200_E_CAPACITY) { overflow = true; rc = 0; }
            else if (!rc) {
                for (uint32_t i = 1; i <= ns; i++) best_set_ptr[s0 + i] = set_off + lp[i];
                set_off += lp[ns];
            }
        }
        ub200_timing t;
        if (!rc && ub200_last_timing(M, &t) == UB200_OK) {
            prep += t.prep_ms; sc += t.score_ms; rd += t.reduce_ms;
            acc.score_launches += t.score_launches; acc.total_launches += t.total_launches;
            acc.score_bytes += t.score_bytes;
        }
        if (rc) return rc;
    }
    if (overflow) {
        // every out[] record is valid; the required capacity is the sum of num_best
        uint64_t need = 0;
        for (uint32_t i = 0; i < n_samples; i++) need += out[i].num_best;
        best_set_ptr[n_samples] = need;
        return fail(UB200_E_CAPACITY, "best_set capacity too small: need " + std::to_string(need));
    }
    acc.prep_ms = prep; acc.score_ms = sc; acc.reduce_ms = rd;
    M->last = acc;
    M->spans.clear();
    return UB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Several GPUs of one box: one replica of the tree per device (derived once, uploaded by one host thread per device),
// the samples of a batch cut into contiguous shards, one host thread per device driving ub200_place_batch, and the
// results written by each device straight into the caller's host arrays (a device -> host copy per shard).  The
// caller lives on the host (the usher binary writes files), so there is no device-side gather: an NCCL allgather
// would only add a device copy before the same D2H.  Multi-PROCESS runs (bench.py under torchrun) gather the
// device-resident records with one ncclAllGather instead.
struct ub200_multi {
    std::vector<ub200_mat*> mats;
};

void ub200_multi_destroy(ub200_multi* X) {
    if (!X) return;
    for (auto* m : X->mats) ub200_mat_destroy(m);
    delete X;
}

int ub200_multi_create(const ub200_flat_mat* flat, int n_devices, const int* devices, ub200_multi** out) {
    if (!flat || !out) return fail(UB200_E_ARG, "ub200_multi_create: NULL argument");
    *out = nullptr;
    const int ndev = ub200_device_count();
    if (ndev <= 0) return fail(UB200_E_NO_DEVICE, "ub200_multi_create: no CUDA device (there is no CPU path)");
    if (n_devices <= 0) n_devices = ndev;
    std::vector<int> ids(n_devices);
    for (int i = 0; i < n_devices; i++) {
        ids[i] = devices ? devices[i] : i;
        if (ids[i] < 0 || ids[i] >= ndev) return fail(UB200_E_ARG, "ub200_multi_create: bad device ordinal");
    }
    int num_sms = 0;
    uint32_t workers = 0;
    { int rc = device_workers(ids[0], &num_sms, &workers); if (rc) return rc; }
    std::shared_ptr<ub200::Derived> dp;
    { int rc = derive_for(flat, workers, dp); if (rc) return rc; }
    auto* X = new ub200_multi();
    X->mats.assign(n_devices, nullptr);
    std::vector<int> rcs(n_devices, 0);
    std::vector<std::string> errs(n_devices);
    std::vector<std::thread> pool;
    for (int i = 0; i < n_devices; i++)
        pool.emplace_back([&, i]() {
            rcs[i] = mat_upload(dp, ids[i], &X->mats[i]);
            if (rcs[i]) errs[i] = g_err;   // the message lives in the worker thread
        });
    for (auto& t : pool) t.join();
    drop_streamed_host_arrays(*dp);
    for (int i = 0; i < n_devices; i++)
        if (rcs[i]) {
            const int rc = rcs[i];
            const std::string e = "device " + std::to_string(ids[i]) + ": " + errs[i];
            ub200_multi_destroy(X);
            return fail(rc, e);
        }
    *out = X;
    return UB200_OK;
}

int ub200_multi_size(const ub200_multi* X) { return X ? (int)X->mats.size() : 0; }
ub200_mat* ub200_multi_mat(ub200_multi* X, int i) { return (X && i >= 0 && i < (int)X->mats.size()) ? X->mats[i] : nullptr; }

int ub200_multi_place_batch(ub200_multi* X, uint32_t n_samples, const uint64_t* sample_ptr, const ub200_mutation* calls,
                            uint32_t flags, ub200_placement* out, int32_t* node_scores, uint32_t* best_set,
                            uint64_t* best_set_ptr, uint64_t best_set_cap) {
    if (!X || X->mats.empty() || !out || !sample_ptr) return fail(UB200_E_ARG, "ub200_multi_place_batch: NULL argument");
    if ((flags & UB200_WANT_NODE_SCORES) && !node_scores) return fail(UB200_E_ARG, "node_scores is NULL");
    if ((flags & UB200_WANT_BEST_SET) && !best_set_ptr) return fail(UB200_E_ARG, "best_set_ptr is NULL");
    const uint32_t nd = (uint32_t)X->mats.size();
    if (nd == 1 || n_samples < 64)
        return ub200_place_batch(X->mats[0], n_samples, sample_ptr, calls, flags, out, node_scores, best_set, best_set_ptr,
                                 best_set_cap);
    // contiguous shards of whole 32-sample groups
    const uint32_t per = ((n_samples + nd - 1) / nd + 31u) / 32u * 32u;
    struct Shard { uint32_t s0 = 0, ns = 0; int rc = 0; std::string err; std::vector<uint32_t> set; std::vector<uint64_t> ptr; };
    std::vector<Shard> sh(nd);
    std::vector<std::thread> pool;
    for (uint32_t i = 0; i < nd; i++) {
        sh[i].s0 = std::min(n_samples, i * per);
        sh[i].ns = std::min(per, n_samples - sh[i].s0);
        if (!sh[i].ns) continue;
        pool.emplace_back([&, i]() {
            Shard& h = sh[i];
            std::vector<uint64_t> ptr(h.ns + 1);
            for (uint32_t k = 0; k <= h.ns; k++) ptr[k] = sample_ptr[h.s0 + k] - sample_ptr[h.s0];
            const ub200_mutation* c = calls ? calls + sample_ptr[h.s0] : nullptr;
            int32_t* ns_out = node_scores ? node_scores + (size_t)h.s0 * X->mats[i]->n : nullptr;
            if (flags & UB200_WANT_BEST_SET) {
                h.ptr.assign(h.ns + 1, 0);
                h.set.assign(std::max<size_t>(1024, 4 * (size_t)h.ns), 0);
                for (;;) {
                    h.rc = ub200_place_batch(X->mats[i], h.ns, ptr.data(), c, flags, out + h.s0, ns_out, h.set.data(),
                                             h.ptr.data(), h.set.size());
                    if (h.rc == UB200_E_CAPACITY) { h.set.assign(h.ptr[h.ns] + 16, 0); continue; }
                    break;
                }
            } else {
                h.rc = ub200_place_batch(X->mats[i], h.ns, ptr.data(), c, flags, out + h.s0, ns_out, nullptr, nullptr, 0);
            }
            if (h.rc) h.err = g_err;
        });
    }
    for (auto& t : pool) t.join();
    for (uint32_t i = 0; i < nd; i++)
        if (sh[i].rc) return fail(sh[i].rc, "device " + std::to_string(X->mats[i]->device) + ": " + sh[i].err);
    if (flags & UB200_WANT_BEST_SET) {
        uint64_t off = 0;
        best_set_ptr[0] = 0;
        for (uint32_t i = 0; i < nd; i++) {
            if (!sh[i].ns) continue;
            for (uint32_t k = 1; k <= sh[i].ns; k++) best_set_ptr[sh[i].s0 + k] = off + sh[i].ptr[k];
            off += sh[i].ptr[sh[i].ns];
        }
        const uint64_t total = best_set_ptr[n_samples];
        if (total > best_set_cap || !best_set) return fail(UB200_E_CAPACITY, "best_set capacity too small: need " + std::to_string(total));
        for (uint32_t i = 0; i < nd; i++)
            if (sh[i].ns) memcpy(best_set + best_set_ptr[sh[i].s0], sh[i].set.data(), (size_t)sh[i].ptr[sh[i].ns] * 4);
    }
    return UB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fitch-Sankoff per VCF site (mapper_body, src/usher_mapper.cpp:6-161): building a MAT from a tree and a VCF.
struct ub200_fs_tree {
    int device = 0;
    uint32_t n = 0, n_levels = 0;
    uint32_t *level_start = nullptr, *parent = nullptr, *child_start = nullptr, *n_children = nullptr;
    uint8_t* scratch = nullptr;
    uint32_t grid = 0;
    cudaStream_t stream = nullptr;
};

void ub200_fs_tree_destroy(ub200_fs_tree* F) {
    if (!F) return;
    cudaSetDevice(F->device);
    cudaFree(F->level_start); cudaFree(F->parent); cudaFree(F->child_start); cudaFree(F->n_children); cudaFree(F->scratch);
    if (F->stream) cudaStreamDestroy(F->stream);
    delete F;
}

int ub200_fs_tree_create(uint32_t n_nodes, const int32_t* parent_bfs, int device, ub200_fs_tree** out) {
    if (!parent_bfs || !out || n_nodes == 0) return fail(UB200_E_ARG, "ub200_fs_tree_create: bad argument");
    *out = nullptr;
    const int ndev = ub200_device_count();
    if (ndev <= 0) return fail(UB200_E_NO_DEVICE, "ub200_fs_tree_create: no CUDA device");
    if (device < 0 || device >= ndev) return fail(UB200_E_ARG, "ub200_fs_tree_create: bad device ordinal");
    // BFS order: parents before children, levels non-decreasing, the children of a node contiguous
    std::vector<uint32_t> level(n_nodes, 0), par(n_nodes, 0), cstart(n_nodes, 0), nch(n_nodes, 0), lstart;
    if (parent_bfs[0] != -1) return fail(UB200_E_TREE_ORDER, "fs tree: node 0 must be the root");
    lstart.push_back(0);
    for (uint32_t i = 1; i < n_nodes; i++) {
        const int32_t p = parent_bfs[i];
        if (p < 0 || (uint32_t)p >= i) return fail(UB200_E_TREE_ORDER, "fs tree: parent is not an earlier node");
        par[i] = (uint32_t)p;
        level[i] = level[p] + 1;
        if (level[i] < level[i - 1] || p < parent_bfs[i - 1]) return fail(UB200_E_TREE_ORDER, "fs tree: nodes are not in BFS order");
        if (level[i] != level[i - 1]) lstart.push_back(i);
        if (nch[p]++ == 0) cstart[p] = i;
    }
    lstart.push_back(n_nodes);
    auto* F = new ub200_fs_tree();
    F->device = device; F->n = n_nodes; F->n_levels = (uint32_t)lstart.size() - 1;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete F;
        return fail(1, "ub200_fs_tree_create: cannot use the device");
    }
    // one CTA per site in flight; scratch = 2 bytes per node and CTA, at most ~2 GB
    const uint64_t per = 2ull * n_nodes;
    F->grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)prop.multiProcessorCount * 2, ((uint64_t)2 << 30) / per));
    int rc = 0;
    auto guard = [&](int r) { if (r && !rc) rc = r; };
    if (cudaStreamCreateWithFlags(&F->stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(1, "cudaStreamCreate failed");
    if (!rc) {
        guard(dev_upload(&F->level_start, lstart.data(), lstart.size(), F->stream));
        guard(dev_upload(&F->parent, par.data(), par.size(), F->stream));
        guard(dev_upload(&F->child_start, cstart.data(), cstart.size(), F->stream));
        guard(dev_upload(&F->n_children, nch.data(), nch.size(), F->stream));
        if (!rc && cudaMalloc((void**)&F->scratch, per * F->grid) != cudaSuccess) rc = fail(1, "cudaMalloc scratch failed");
        if (!rc && cudaStreamSynchronize(F->stream) != cudaSuccess) rc = fail(1, "fs tree upload failed");
    }
    if (rc) { ub200_fs_tree_destroy(F); return rc; }
    *out = F;
    return UB200_OK;
}

int ub200_fs_sites(ub200_fs_tree* F, uint32_t n_sites, const uint8_t* ref_code, const uint64_t* var_ptr,
                   const uint32_t* var_node, const uint8_t* var_nuc, uint64_t out_cap, uint32_t* out_site,
                   uint32_t* out_node, uint8_t* out_states, uint64_t* out_count) {
    if (!F || !ref_code || !var_ptr || !out_count) return fail(UB200_E_ARG, "ub200_fs_sites: NULL argument");
    *out_count = 0;
    if (n_sites == 0) return UB200_OK;
    const uint64_t nv = var_ptr[n_sites];
    if (nv && (!var_node || !var_nuc)) return fail(UB200_E_ARG, "ub200_fs_sites: NULL genotype arrays");
    for (uint32_t s = 0; s < n_sites; s++)
        if (ref_code[s] > 3 || var_ptr[s + 1] < var_ptr[s]) return fail(UB200_E_ARG, "ub200_fs_sites: bad site " + std::to_string(s));
    for (uint64_t k = 0; k < nv; k++)
        if (var_node[k] >= F->n || var_nuc[k] == 0 || var_nuc[k] > 15) return fail(UB200_E_ARG, "ub200_fs_sites: bad genotype entry");
    CU(cudaSetDevice(F->device));
    uint8_t *d_ref = nullptr, *d_nuc = nullptr, *d_ostates = nullptr;
    unsigned long long *d_ptr = nullptr, *d_count = nullptr;
    uint32_t *d_node = nullptr, *d_osite = nullptr, *d_onode = nullptr;
    int rc = 0;
    auto guard = [&](int r) { if (r && !rc) rc = r; };
    guard(dev_upload(&d_ref, ref_code, n_sites, F->stream));
    guard(dev_upload(&d_ptr, reinterpret_cast<const unsigned long long*>(var_ptr), (size_t)n_sites + 1, F->stream));
    guard(dev_upload(&d_node, var_node, nv, F->stream));
    guard(dev_upload(&d_nuc, var_nuc, nv, F->stream));
    auto dmalloc = [&](void** q, size_t bytes) { if (!rc && cudaMalloc(q, std::max<size_t>(bytes, 16)) != cudaSuccess) rc = fail(1, "cudaMalloc failed"); };
    dmalloc((void**)&d_osite, out_cap * 4); dmalloc((void**)&d_onode, out_cap * 4); dmalloc((void**)&d_ostates, out_cap);
    dmalloc((void**)&d_count, 8);
    if (!rc) {
        cudaMemsetAsync(d_count, 0, 8, F->stream);
        ub200::FsParams p;
        p.n_nodes = F->n; p.n_levels = F->n_levels; p.n_sites = n_sites;
        p.level_start = F->level_start; p.parent = F->parent; p.child_start = F->child_start; p.n_children = F->n_children;
        p.ref_code = d_ref; p.var_ptr = d_ptr; p.var_node = d_node; p.var_nuc = d_nuc; p.scratch = F->scratch;
        p.out_cap = out_cap; p.out_site = d_osite; p.out_node = d_onode; p.out_states = d_ostates; p.out_count = d_count;
        const uint32_t threads = F->n >= 4096 ? 1024u : (F->n >= 512 ? 256u : 64u);
        ub200::k_fitch_sankoff<<<std::min(F->grid, n_sites), threads, 0, F->stream>>>(p);
        unsigned long long cnt = 0;
        cudaError_t e = cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, F->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(F->stream);
        if (e != cudaSuccess) rc = fail((int)e, std::string("k_fitch_sankoff: ") + cudaGetErrorString(e));
        *out_count = cnt;
        if (!rc && cnt > out_cap) rc = fail(UB200_E_CAPACITY, "ub200_fs_sites: output capacity too small: need " + std::to_string(cnt));
        if (!rc && cnt) {
            cudaMemcpyAsync(out_site, d_osite, cnt * 4, cudaMemcpyDeviceToHost, F->stream);
            cudaMemcpyAsync(out_node, d_onode, cnt * 4, cudaMemcpyDeviceToHost, F->stream);
            cudaMemcpyAsync(out_states, d_ostates, cnt, cudaMemcpyDeviceToHost, F->stream);
            if (cudaStreamSynchronize(F->stream) != cudaSuccess) rc = fail(1, "ub200_fs_sites: download failed");
        }
    }
    cudaFree(d_ref); cudaFree(d_ptr); cudaFree(d_node); cudaFree(d_nuc); cudaFree(d_osite); cudaFree(d_onode);
    cudaFree(d_ostates); cudaFree(d_count);
    return rc;
}

}  // extern "C"
