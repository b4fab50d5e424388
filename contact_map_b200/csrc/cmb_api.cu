// cmb_api.cu -- C ABI of libcmb200.so (see include/cmb200.h for the contract and the
// reference interfaces each entry point replaces).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cmb200.h"
#include "cmb_aux_kernels.cuh"
#include "cmb_frames_large.cuh"
#include "cmb_frames_small.cuh"
#include "cmb_verlet.cuh"
#include "cmb_prune.cuh"

using namespace cmb;

struct FrameJob {
    const float* xyz; const float* box; long long n_frames; long long frame0;
    unsigned int* atom_table; unsigned int* res_table;
    unsigned long long* atom_keys; long long atom_cap;
    unsigned long long* res_keys; long long res_cap;
    unsigned long long* ext_counters;      // caller-provided device counters or nullptr
    float4* boundary; long long boundary_cap; int boundary_ulps;
    int shared_list;                       // 1: a chunk of a streamed trajectory: walk the context's shared pair
                                           // list (built from the first chunk, rebuilt when too many frames trip)
};

struct cmb_ctx {
    int device = 0;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::string err;

    // system
    bool have_system = false;
    int n = 0;
    int n_res_c = 0;
    float cutoff_f = 0.f;
    int n_ignored = 0;
    std::vector<int32_t> res_map;     // compact -> global residue index
    int2* d_meta = nullptr;           // {ekey, res_c | role << 30}   (global-memory cell path)
    unsigned int* d_meta32 = nullptr; // packed word of the fused kernel (cmb_frames_small.cuh)
    bool all_roles = true;            // every used atom is both query and haystack
    int meta_cap = 0;
    int path = 0;                     // 0 fused shared-memory kernel, 1 global-memory cell path

    // fused-kernel launch configuration
    int fs_maxc = 4096;
    int fs_bm_words = 0;
    int fs_bm_in_smem = 1;
    int fs_smem = 0;
    int fs_ctas_per_sm = 1;
    int fs_threads = 512;             // 512 (two CTAs per SM) or 1024 (one CTA per SM)

    // scratch, two slots so that two in-flight launches never share it
    unsigned int* d_bitmap[2] = {nullptr, nullptr};
    size_t bitmap_bytes[2] = {0, 0};
    unsigned long long* d_counters[2] = {nullptr, nullptr};   // 4 x u64 each
    int slot = 0;

    // large path scratch
    LargeScratch large[2];

    // frame-batch pair-list path (cmb_verlet.cuh), one scratch set per stream slot
    VlScratch vl[2];
    int* d_res_anchor = nullptr;              // [n_res_c] first used atom of each compact residue
    int verlet = 1;                           // 0 off, 1 when the batch is long enough, 2 always
    int verlet_graph = 1;                     // replay the list builder as one CUDA graph (0: plain launches)
    int verlet_audit = 0;                     // run the exact arithmetic on every listed pair and count disagreements
    int verlet_max_rlist_pct = 175;           // lists with r_list above this percentage of the cutoff are not built
    int verlet_tile_share_pct = 60;           // optimistic tile sizing: share of the shell atoms counted as slots (tests lower it
                                              // to provoke the overflow -> pessimistic rebuild path)
    unsigned long long* d_vl_audit = nullptr; // [2] audited pairs, disagreements
    int last_used_verlet = 0;
    int last_vl_slot = 0;
    int vl_conservative = 0;                  // the optimistic tile size overflowed a tile of this system once
    VlRun vl_run[2];                          // per stream slot: state of one walk over a list
    int vl_shared = -1;                       // scratch holding the list the chunks of a streamed trajectory share
    int vl_shared_stale = 0;                  // too many frames of a chunk tripped: the next chunk rebuilds the list
    cudaEvent_t vl_built[2] = {nullptr, nullptr};            // list of scratch X complete
    cudaEvent_t vl_walked[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // last walk of list X on stream slot s
    bool vl_walked_set[2][2] = {{false, false}, {false, false}};
    cudaEvent_t vl_info_ready[2] = {nullptr, nullptr};       // read-back of slot s's walk has landed
    // the large path's fall back of a streamed chunk is finished when its stream slot comes round again
    struct Pending { bool active = false; bool info = false; bool built = false; FrameJob job; cudaStream_t stream = nullptr; int list = 0; } pending[2];

    // host pipeline
    float* d_stage_xyz[2] = {nullptr, nullptr};
    float* d_stage_box[2] = {nullptr, nullptr};
    size_t stage_xyz_bytes = 0, stage_box_bytes = 0;
    float* h_pinned[2] = {nullptr, nullptr};      // gather staging (cmb_accumulate_host_gather)
    size_t pinned_bytes = 0;
    cudaEvent_t copy_done[2] = {nullptr, nullptr};
    cudaEvent_t box_ready = nullptr;              // the one copy of all unit cells has landed
    cudaEvent_t ext_ready = nullptr;              // cmb_wait_for_stream: what the caller queued before the next host call
    bool ext_ready_pending = false;
    cudaStream_t streams[2] = {nullptr, nullptr};

    // sort / min-dist scratch
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;

    // options / stats
    int count_tests = 0;
    int pair_tasks = 1;
    int stage_pageable = 1;                   // pageable host input goes through threaded pinned staging
    long long stage_chunk_frames = 0;         // frames per staging chunk of the host paths (0: automatic)
    int prune = 1;                            // NearestAtoms / MinimumDistanceCounter: cell-pruned search (cmb_prune.cuh)
    PgScratch pg;
    int large_batch = 1;                      // global-memory cell path: 0 one frame at a time, 1 auto, 2 batched whenever possible
    int atom_log2 = 0, res_log2 = 0;          // hashed count tables when > 0 (cmb_set_table_mode)
    unsigned long long* d_hstats = nullptr;   // [4] new atom keys, new residue keys, parked hits, failed re-inserts
    unsigned long long* d_ovf = nullptr;      // caller-owned overflow list (cmb_hash_overflow_buffer)
    long long ovf_cap = 0;
    int last_slot = 0;
    std::map<std::string, int64_t> stats;

    // measurement (option time_kernels): CUDA events around every launch of the pair-list frame kernel, on
    // the stream it is launched on; the stats vl_frames_ns / vl_frames_timed read and reset them
    int time_kernels = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kernel_events;
    size_t kernel_events_used = 0;
    // host pipeline trace (option stage_trace): events after the copy and after the kernels of every chunk
    int stage_trace = 0;
    cudaEvent_t trace_begin = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> trace_events;
    size_t trace_used = 0;
};

// a recycled pair of timing events
static int timing_pair(std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& pool, size_t& used, cudaEvent_t* a, cudaEvent_t* b)
{
    if (used == pool.size()) {
        cudaEvent_t x = nullptr, y = nullptr;
        if (cudaEventCreate(&x) != cudaSuccess || cudaEventCreate(&y) != cudaSuccess) return -1;
        pool.push_back({x, y});
    }
    *a = pool[used].first; *b = pool[used].second;
    used++;
    return 0;
}

static std::string g_create_err;

static int fail(cmb_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_create_err = buf;
    return code;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(ctx, CMB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                     \
    } while (0)

static int ensure_bytes(cmb_ctx* ctx, void** ptr, size_t* have, size_t need)
{
    if (*have >= need && *ptr) return CMB_OK;
    if (*ptr) CK(cudaFree(*ptr));
    *ptr = nullptr;
    *have = 0;
    size_t want = need + need / 4 + 256;
    CK(cudaMalloc(ptr, want));
    *have = want;
    return CMB_OK;
}

extern "C" int cmb_version(void) { return 100; }

extern "C" const char* cmb_last_error(const cmb_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int cmb_ctx_create(int device, cmb_ctx** out)
{
    cmb_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, CMB_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CMB_ERR_CUDA, "no CUDA device available (%s); libcmb200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, CMB_ERR_ARG, "device %d out of range", device);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CMB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, CMB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, CMB_ERR_CUDA, "device %d is sm_%d%d; libcmb200 is built for sm_100a only", device,
                    prop.major, prop.minor);
    ctx = new cmb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    for (int s = 0; s < 2; s++) {
        if (cudaMalloc(&ctx->d_counters[s], 4 * sizeof(unsigned long long)) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ctx->streams[s], cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return fail(nullptr, CMB_ERR_CUDA, "context allocation failed");
        }
    }
    *out = ctx;
    return CMB_OK;
}

extern "C" void cmb_ctx_destroy(cmb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_meta);
    cudaFree(ctx->d_meta32);
    for (int s = 0; s < 2; s++) {
        cudaFree(ctx->d_bitmap[s]);
        cudaFree(ctx->d_counters[s]);
        cudaFree(ctx->d_stage_xyz[s]);
        cudaFree(ctx->d_stage_box[s]);
        if (ctx->h_pinned[s]) cudaFreeHost(ctx->h_pinned[s]);
        if (ctx->copy_done[s]) cudaEventDestroy(ctx->copy_done[s]);
        if (ctx->streams[s]) cudaStreamDestroy(ctx->streams[s]);
        large_scratch_free(ctx->large[s]);
    }
    for (int s = 0; s < 2; s++) {
        vl_free(ctx->vl[s]);
        vl_run_free(ctx->vl_run[s]);
        if (ctx->vl_built[s]) cudaEventDestroy(ctx->vl_built[s]);
        if (ctx->vl_info_ready[s]) cudaEventDestroy(ctx->vl_info_ready[s]);
        for (int t = 0; t < 2; t++)
            if (ctx->vl_walked[s][t]) cudaEventDestroy(ctx->vl_walked[s][t]);
    }
    for (auto& e : ctx->kernel_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto& e : ctx->trace_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (ctx->trace_begin) cudaEventDestroy(ctx->trace_begin);
    if (ctx->box_ready) cudaEventDestroy(ctx->box_ready);
    if (ctx->ext_ready) cudaEventDestroy(ctx->ext_ready);
    pg_free(ctx->pg);
    cudaFree(ctx->d_res_anchor);
    cudaFree(ctx->d_vl_audit);
    cudaFree(ctx->d_tmp);
    cudaFree(ctx->d_hstats);
    delete ctx;
}

extern "C" int cmb_set_option(cmb_ctx* ctx, const char* name, int64_t value)
{
    if (!ctx || !name) return CMB_ERR_ARG;
    if (!strcmp(name, "count_tests")) { ctx->count_tests = value != 0; return CMB_OK; }
    if (!strcmp(name, "pair_tasks")) { ctx->pair_tasks = value != 0; return CMB_OK; }
    if (!strcmp(name, "stage_pageable")) { ctx->stage_pageable = value != 0; return CMB_OK; }
    if (!strcmp(name, "stage_chunk_frames")) {
        if (value < 0) return fail(ctx, CMB_ERR_ARG, "stage_chunk_frames must be >= 0");
        ctx->stage_chunk_frames = value;
        return CMB_OK;
    }
    if (!strcmp(name, "large_batch")) {
        if (value < 0 || value > 2) return fail(ctx, CMB_ERR_ARG, "large_batch must be 0, 1 or 2");
        ctx->large_batch = (int)value;
        return CMB_OK;
    }
    if (!strcmp(name, "max_cells")) {
        if (value < 8 || value > 16383) return fail(ctx, CMB_ERR_ARG, "max_cells out of range");
        ctx->fs_maxc = (int)value;
        return CMB_OK;
    }
    if (!strcmp(name, "force_path")) { ctx->stats["force_path"] = value; return CMB_OK; }
    if (!strcmp(name, "verlet")) {
        if (value < 0 || value > 2) return fail(ctx, CMB_ERR_ARG, "verlet must be 0 (off), 1 (auto) or 2 (always)");
        ctx->verlet = (int)value;
        return CMB_OK;
    }
    if (!strcmp(name, "verlet_audit")) { ctx->verlet_audit = value != 0; return CMB_OK; }
    if (!strcmp(name, "verlet_graph")) { ctx->verlet_graph = value != 0; return CMB_OK; }
    if (!strcmp(name, "verlet_tile_share_pct")) {
        if (value < 0 || value > 100) return fail(ctx, CMB_ERR_ARG, "verlet_tile_share_pct out of range");
        ctx->verlet_tile_share_pct = (int)value;
        return CMB_OK;
    }
    if (!strcmp(name, "time_kernels")) { ctx->time_kernels = value != 0; ctx->kernel_events_used = 0; return CMB_OK; }
    if (!strcmp(name, "stage_trace")) { ctx->stage_trace = value != 0; ctx->trace_used = 0; return CMB_OK; }
    if (!strcmp(name, "prune")) { ctx->prune = value != 0; return CMB_OK; }
    if (!strcmp(name, "verlet_conservative_tiles")) { ctx->vl_conservative = value != 0; return CMB_OK; }
    if (!strcmp(name, "verlet_max_rlist_pct")) {
        if (value < 101 || value > 400) return fail(ctx, CMB_ERR_ARG, "verlet_max_rlist_pct out of range");
        ctx->verlet_max_rlist_pct = (int)value;
        return CMB_OK;
    }
    return fail(ctx, CMB_ERR_ARG, "unknown option %s", name);
}

extern "C" int cmb_get_stat(cmb_ctx* ctx, const char* name, int64_t* value)
{
    if (!ctx || !name || !value) return CMB_ERR_ARG;
    if (!strcmp(name, "pair_tests")) {
        unsigned long long h[4];
        CK(cudaSetDevice(ctx->device));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, ctx->d_counters[ctx->last_slot], sizeof(h), cudaMemcpyDeviceToHost));
        *value = (int64_t)h[3];
        return CMB_OK;
    }
    if (!strcmp(name, "sm_count")) { *value = ctx->sm_count; return CMB_OK; }
    if (!strcmp(name, "ctas_per_sm")) { *value = ctx->fs_ctas_per_sm; return CMB_OK; }
    if (!strcmp(name, "smem_bytes")) { *value = ctx->fs_smem; return CMB_OK; }
    if (!strcmp(name, "max_cells")) { *value = ctx->fs_maxc; return CMB_OK; }
    if (!strcmp(name, "kernel_launches")) { *value = ctx->stats["kernel_launches"]; return CMB_OK; }
    if (!strcmp(name, "vl_used")) { *value = ctx->last_used_verlet; return CMB_OK; }
    if (!strcmp(name, "vl_graph_state")) { *value = ctx->vl[ctx->last_vl_slot].graph_state; return CMB_OK; }
    if (!strcmp(name, "vl_graph_captures")) { *value = ctx->vl[0].captures + ctx->vl[1].captures; return CMB_OK; }
    if (!strcmp(name, "vl_graph_capture_us")) { *value = ctx->vl[0].capture_us + ctx->vl[1].capture_us; return CMB_OK; }
    if (!strcmp(name, "vl_frames_timed")) { *value = (int64_t)ctx->kernel_events_used; return CMB_OK; }
    if (!strcmp(name, "vl_frames_ns")) {
        // total duration of the frame-kernel launches timed since the last read (synchronises, resets)
        CK(cudaSetDevice(ctx->device));
        CK(cudaDeviceSynchronize());
        double ms = 0.0;
        for (size_t k = 0; k < ctx->kernel_events_used; k++) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, ctx->kernel_events[k].first, ctx->kernel_events[k].second));
            ms += t;
        }
        ctx->kernel_events_used = 0;
        *value = (int64_t)(ms * 1.0e6);
        return CMB_OK;
    }
    if (!strncmp(name, "trace_", 6)) {
        // host pipeline trace of the last streamed call: trace_chunks, trace_copy_<c>, trace_done_<c>
        // (ns after the first copy was issued)
        if (!strcmp(name, "trace_chunks")) { *value = (int64_t)ctx->trace_used; return CMB_OK; }
        const bool copy = !strncmp(name, "trace_copy_", 11);
        const bool done = !strncmp(name, "trace_done_", 11);
        if (!copy && !done) return fail(ctx, CMB_ERR_ARG, "unknown stat %s", name);
        const size_t c = (size_t)atoll(name + 11);
        if (c >= ctx->trace_used || !ctx->trace_begin) return fail(ctx, CMB_ERR_ARG, "no trace for chunk %zu", c);
        CK(cudaSetDevice(ctx->device));
        CK(cudaDeviceSynchronize());
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, ctx->trace_begin, copy ? ctx->trace_events[c].first : ctx->trace_events[c].second));
        *value = (int64_t)((double)t * 1.0e6);
        return CMB_OK;
    }
    if (!strcmp(name, "pruned_frames")) { *value = ctx->stats["pruned_frames"]; return CMB_OK; }
    if (!strcmp(name, "brute_frames")) { *value = ctx->stats["brute_frames"]; return CMB_OK; }
    if (!strcmp(name, "vl_rebuilds")) { *value = ctx->stats["vl_rebuilds"]; return CMB_OK; }
    if (!strcmp(name, "vl_rebuild_code")) { *value = ctx->stats["vl_rebuild_code"]; return CMB_OK; }
    if (!strcmp(name, "vl_conservative")) { *value = ctx->vl_conservative; return CMB_OK; }
    if (!strncmp(name, "vl_", 3)) {
        // pair-list path diagnostics of the last launch (synchronises)
        const VlScratch& V = ctx->vl[ctx->last_vl_slot];
        if (!V.L.hdr) { *value = 0; return CMB_OK; }
        CK(cudaSetDevice(ctx->device));
        CK(cudaDeviceSynchronize());
        VlHeader h;
        CK(cudaMemcpy(&h, V.L.hdr, sizeof(h), cudaMemcpyDeviceToHost));
        if (!strcmp(name, "vl_valid")) { *value = h.valid; return CMB_OK; }
        if (!strcmp(name, "vl_fail")) { *value = h.fail_code; return CMB_OK; }
        if (!strcmp(name, "vl_tripped")) {
            int t = 0;
            if (ctx->vl_run[ctx->last_slot].state) CK(cudaMemcpy(&t, ctx->vl_run[ctx->last_slot].state, sizeof(int), cudaMemcpyDeviceToHost));
            *value = t;
            return CMB_OK;
        }
        if (!strcmp(name, "vl_entries")) { *value = h.n_entries; return CMB_OK; }
        if (!strcmp(name, "vl_tiles")) { *value = h.n_tiles; return CMB_OK; }
        if (!strcmp(name, "vl_tdim")) { *value = h.tdim; return CMB_OK; }
        if (!strcmp(name, "vl_rows")) { *value = h.n_rows; return CMB_OK; }
        if (!strcmp(name, "vl_slots")) { *value = h.n_slots; return CMB_OK; }
        if (!strcmp(name, "vl_segs")) { *value = h.n_segs; return CMB_OK; }
        if (!strcmp(name, "vl_bundles")) { *value = h.n_bundles; return CMB_OK; }
        if (!strcmp(name, "vl_rlist_um")) { *value = (int64_t)(h.rlist * 1.0e6f); return CMB_OK; }
        if (!strcmp(name, "vl_audited") || !strcmp(name, "vl_disagree")) {
            unsigned long long a[2] = {0, 0};
            if (ctx->d_vl_audit) CK(cudaMemcpy(a, ctx->d_vl_audit, sizeof(a), cudaMemcpyDeviceToHost));
            *value = (int64_t)a[!strcmp(name, "vl_disagree") ? 1 : 0];
            return CMB_OK;
        }
    }
    return fail(ctx, CMB_ERR_ARG, "unknown stat %s", name);
}

// ---------------------------------------------------------------------------------
// system set-up
// ---------------------------------------------------------------------------------
template <int THREADS>
static int fs_prepare(cmb_ctx* ctx, int smem_total, int* ctas_per_sm)
{
    CK(cudaFuncSetAttribute(k_frames_small<0, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total));
    CK(cudaFuncSetAttribute(k_frames_small<FS_ROLES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total));
    CK(cudaFuncSetAttribute(k_frames_small<FS_EMIT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total));
    CK(cudaFuncSetAttribute(k_frames_small<FS_EMIT | FS_ROLES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            smem_total));
    CK(cudaFuncSetAttribute(k_frames_small<FS_BOUNDARY | FS_ROLES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            smem_total));
    CK(cudaFuncSetAttribute(k_frames_small<FS_COUNT | FS_ROLES, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            smem_total));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, k_frames_small<0, THREADS>, THREADS,
                                                     (size_t)smem_total));
    return CMB_OK;
}

static int configure_small_path(cmb_ctx* ctx)
{
    const int n = ctx->n;
    const long long bits = (long long)ctx->n_res_c * ctx->n_res_c;
    ctx->fs_bm_words = (int)((bits + 31) / 32);
    // residue bitmap in shared memory when it leaves room for two CTAs per SM
    int maxc = ctx->fs_maxc;
    const int budget_two = (ctx->max_smem_optin + 1024) / 2 - 1024;
    // 512 threads and two CTAs per SM when two copies of the arrays fit, else one CTA of 1024 threads
    // (same number of warps per SM; its per-warp candidate lists and rings need 12 KB more)
    int warps = 16;
    {
        const FsLayout two = fs_layout(n, maxc, ctx->fs_bm_words, 16);
        const FsLayout two_nobm = fs_layout(n, maxc, ctx->fs_bm_words * 4 > 48 * 1024 ? 0 : ctx->fs_bm_words, 16);
        if (two.total > budget_two && two_nobm.total > budget_two &&
            fs_layout(n, maxc, 0, 32).total <= ctx->max_smem_optin)
            warps = 32;
    }
    FsLayout with_bm = fs_layout(n, maxc, ctx->fs_bm_words, warps);
    ctx->fs_bm_in_smem = 1;
    if (with_bm.total > ctx->max_smem_optin ||
        (warps == 16 && ctx->fs_bm_words * 4 > 48 * 1024 && with_bm.total > budget_two))
        ctx->fs_bm_in_smem = 0;
    FsLayout L = fs_layout(n, maxc, ctx->fs_bm_in_smem ? ctx->fs_bm_words : 0, warps);
    while (L.total > ctx->max_smem_optin && maxc > 64) {
        maxc /= 2;
        L = fs_layout(n, maxc, ctx->fs_bm_in_smem ? ctx->fs_bm_words : 0, warps);
    }
    if (L.total > ctx->max_smem_optin) return fail(ctx, CMB_ERR_LIMIT, "system does not fit the fused kernel");
    ctx->fs_maxc = maxc;
    ctx->fs_smem = L.total;
    ctx->fs_threads = warps * 32;
    int nb = 0;
    if (warps == 16) {
        int rc = fs_prepare<512>(ctx, L.total, &nb);
        if (rc != CMB_OK) return rc;
    } else {
        int rc = fs_prepare<1024>(ctx, L.total, &nb);
        if (rc != CMB_OK) return rc;
    }
    if (nb < 1) return fail(ctx, CMB_ERR_LIMIT, "fused kernel cannot be resident (smem %d)", L.total);
    ctx->fs_ctas_per_sm = nb;
    return CMB_OK;
}

extern "C" int cmb_set_system(cmb_ctx* ctx, int n_used, const int32_t* atom_res, const int32_t* atom_chain,
                              const uint8_t* role, double cutoff, int n_ignored)
{
    if (!ctx) return CMB_ERR_ARG;
    if (n_used <= 0 || !atom_res || !atom_chain || !role) return fail(ctx, CMB_ERR_ARG, "bad system arrays");
    if (n_ignored < 0) return fail(ctx, CMB_ERR_ARG, "n_ignored must be >= 0");
    if (!(cutoff > 0.0)) return fail(ctx, CMB_ERR_ARG, "cutoff must be positive");
    if (n_used >= (1 << 20)) return fail(ctx, CMB_ERR_LIMIT, "more than 2^20 used atoms");
    CK(cudaSetDevice(ctx->device));

    // compact residue ids (rank among the used residues)
    std::vector<int32_t> uniq(atom_res, atom_res + n_used);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    ctx->res_map = uniq;
    ctx->n_res_c = (int)uniq.size();

    // exclusion key: |ekey_a - ekey_b| <= n_ignored  <=>  same chain and |res_a - res_b| <= n_ignored.
    // Used residues are laid out on the key axis chain after chain; inside a chain consecutive used
    // residues advance the key by min(gap, n_ignored + 1), a new chain by n_ignored + 1.  Gaps larger
    // than n_ignored are thereby compressed without changing the predicate, which keeps the keys small.
    std::vector<std::pair<int32_t, int32_t>> chain_res;   // (chain, residue) of every used residue
    std::vector<int32_t> chain_of_uniq(uniq.size(), 0);
    std::vector<int> rc_of_atom(n_used);
    {
        std::vector<char> seen(uniq.size(), 0);
        for (int i = 0; i < n_used; i++) {
            const int rc = (int)(std::lower_bound(uniq.begin(), uniq.end(), atom_res[i]) - uniq.begin());
            rc_of_atom[i] = rc;
            if (!seen[rc]) { seen[rc] = 1; chain_of_uniq[rc] = atom_chain[i]; }
            else if (chain_of_uniq[rc] != atom_chain[i])
                return fail(ctx, CMB_ERR_ARG, "residue %d is assigned to two chains", atom_res[i]);
        }
        chain_res.reserve(uniq.size());
        for (size_t k = 0; k < uniq.size(); k++) chain_res.push_back({chain_of_uniq[k], uniq[k]});
        std::sort(chain_res.begin(), chain_res.end());
    }
    std::vector<long long> ekey_of_rc(uniq.size(), 0);
    long long cursor = 0;
    for (size_t k = 0; k < chain_res.size(); k++) {
        if (k > 0) {
            const bool same_chain = chain_res[k].first == chain_res[k - 1].first;
            const long long gap = (long long)chain_res[k].second - chain_res[k - 1].second;
            cursor += same_chain ? std::min<long long>(gap, (long long)n_ignored + 1) : (long long)n_ignored + 1;
        }
        const size_t rc = (size_t)(std::lower_bound(uniq.begin(), uniq.end(), chain_res[k].second) - uniq.begin());
        ekey_of_rc[rc] = cursor;
    }
    // keys travel as floats in the cell-sorted records: they must be exactly representable
    if (cursor >= (1ll << 24)) return fail(ctx, CMB_ERR_LIMIT, "exclusion key range too large (>= 2^24)");
    const long long ekey_max = cursor;
    std::vector<int2> meta(n_used);
    std::vector<unsigned int> meta32(n_used);
    bool all_roles = true;
    for (int i = 0; i < n_used; i++) {
        const int rc = rc_of_atom[i];
        const int ekey = (int)ekey_of_rc[rc];
        meta[i].x = ekey;
        meta[i].y = rc | ((int)(role[i] & 3u) << 30);
        meta32[i] = ((unsigned)ekey & FS_EKEY_MASK) | (((unsigned)rc & FS_RES_MASK) << FS_EKEY_BITS) |
                    ((unsigned)(role[i] & 3u) << 30);
        if ((role[i] & 3u) != 3u) all_roles = false;
    }
    ctx->all_roles = all_roles;
    // earlier launches (possibly on non-blocking streams) may still read the old tables
    CK(cudaDeviceSynchronize());
    if (ctx->meta_cap < n_used) {            // device buffers are reused across systems
        if (ctx->d_meta) CK(cudaFree(ctx->d_meta));
        if (ctx->d_meta32) CK(cudaFree(ctx->d_meta32));
        ctx->d_meta = nullptr; ctx->d_meta32 = nullptr; ctx->meta_cap = 0;
        CK(cudaMalloc(&ctx->d_meta, sizeof(int2) * (size_t)n_used));
        CK(cudaMalloc(&ctx->d_meta32, sizeof(unsigned int) * (size_t)n_used));
        ctx->meta_cap = n_used;
    }
    CK(cudaMemcpy(ctx->d_meta, meta.data(), sizeof(int2) * (size_t)n_used, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_meta32, meta32.data(), sizeof(unsigned int) * (size_t)n_used, cudaMemcpyHostToDevice));
    {
        // first used atom of every compact residue (tile assignment of the pair-list path)
        std::vector<int> anchor(uniq.size(), -1);
        for (int i = n_used - 1; i >= 0; i--) anchor[rc_of_atom[i]] = i;
        if (ctx->d_res_anchor) CK(cudaFree(ctx->d_res_anchor));
        ctx->d_res_anchor = nullptr;
        CK(cudaMalloc(&ctx->d_res_anchor, sizeof(int) * anchor.size()));
        CK(cudaMemcpy(ctx->d_res_anchor, anchor.data(), sizeof(int) * anchor.size(), cudaMemcpyHostToDevice));
    }

    ctx->n = n_used;
    ctx->vl_conservative = 0;
    ctx->vl_shared = -1;
    ctx->cutoff_f = (float)cutoff;
    ctx->n_ignored = n_ignored;
    ctx->have_system = false;
    int64_t force = ctx->stats.count("force_path") ? ctx->stats["force_path"] : -1;
    // the fused kernel packs ekey (FS_EKEY_BITS = 16 bits), the compact residue id (13 bits) and the role (2 bits) into one word
    ctx->path = (n_used <= FS_MAX_ATOMS && ekey_max <= (long long)FS_EKEY_MASK &&
                 (long long)ctx->n_res_c <= (long long)FS_RES_MASK) ? 0 : 1;
    if (force == 1) ctx->path = 1;
    if (ctx->path == 0) {
        int rc = configure_small_path(ctx);
        if (rc != CMB_OK) return rc;
    }
    ctx->have_system = true;
    return CMB_OK;
}

extern "C" int cmb_n_residues(const cmb_ctx* ctx) { return ctx && ctx->have_system ? ctx->n_res_c : CMB_ERR_STATE; }

extern "C" int cmb_residue_map(const cmb_ctx* ctx, int32_t* out)
{
    if (!ctx || !ctx->have_system || !out) return CMB_ERR_STATE;
    memcpy(out, ctx->res_map.data(), sizeof(int32_t) * ctx->res_map.size());
    return CMB_OK;
}

extern "C" int cmb_table_sizes(const cmb_ctx* ctx, int64_t* atom_elems, int64_t* res_elems)
{
    if (!ctx || !ctx->have_system) return CMB_ERR_STATE;
    // dense: uint32 entries; hashed (cmb_set_table_mode): uint64 slots
    if (atom_elems) *atom_elems = ctx->atom_log2 > 0 ? (1ll << ctx->atom_log2) : (int64_t)ctx->n * ctx->n;
    if (res_elems) *res_elems = ctx->res_log2 > 0 ? (1ll << ctx->res_log2) : (int64_t)ctx->n_res_c * ctx->n_res_c;
    return CMB_OK;
}

extern "C" int cmb_path_kind(const cmb_ctx* ctx) { return ctx && ctx->have_system ? ctx->path : CMB_ERR_STATE; }

// ---------------------------------------------------------------------------------
// frame kernels
// ---------------------------------------------------------------------------------
// the fused-kernel variant a job asks for
template <int THREADS>
static void fs_launch(const cmb_ctx* ctx, const FrameJob& job, const FrameParams& p, unsigned g, size_t sm,
                      cudaStream_t stream)
{
    const bool emit = job.atom_keys != nullptr || job.res_keys != nullptr;
    if (job.boundary_ulps != 0)
        k_frames_small<FS_BOUNDARY | FS_ROLES, THREADS><<<g, THREADS, sm, stream>>>(p);
    else if (ctx->count_tests && !emit)
        k_frames_small<FS_COUNT | FS_ROLES, THREADS><<<g, THREADS, sm, stream>>>(p);
    else if (emit && ctx->all_roles)
        k_frames_small<FS_EMIT, THREADS><<<g, THREADS, sm, stream>>>(p);
    else if (emit)
        k_frames_small<FS_EMIT | FS_ROLES, THREADS><<<g, THREADS, sm, stream>>>(p);
    else if (ctx->all_roles)
        k_frames_small<0, THREADS><<<g, THREADS, sm, stream>>>(p);
    else
        k_frames_small<FS_ROLES, THREADS><<<g, THREADS, sm, stream>>>(p);
}


static int large_cells(cmb_ctx* ctx, const FrameJob& job, int slot, cudaStream_t stream,
                       const std::vector<std::pair<long long, long long>>& runs);

// The frames of a walked pair list that the list could not take (guard tripped, list unusable),
// global-memory cell path: the count was requested from the device when the walk was enqueued;
// wait for it, fetch the frame numbers and run the cell kernels on them as runs of consecutive
// frames.  Small systems never come here: the fused kernel reads the device-side list itself.
static int finish_pending(cmb_ctx* ctx, int slot)
{
    cmb_ctx::Pending& P = ctx->pending[slot];
    if (!P.active) return CMB_OK;
    P.active = false;
    VlRun& R = ctx->vl_run[slot];
    CK(cudaEventSynchronize(ctx->vl_info_ready[slot]));
    const int n_trip = R.h_info[0];
    if (P.job.shared_list && n_trip * 8 > P.job.n_frames) ctx->vl_shared_stale = 1;
    if (P.built && !ctx->vl_conservative && R.h_info[1] == 0 && R.h_info[2] >= 9 && R.h_info[2] <= 13) {
        // the list this chunk built overflowed a tile with the optimistic tile size (its frames all
        // came back on the fall-back list): the next chunk rebuilds it with the pessimistic one
        ctx->vl_conservative = 1;
        ctx->vl_shared_stale = 1;
        ctx->stats["vl_rebuilds"] += 1;
        ctx->stats["vl_rebuild_code"] = R.h_info[2];
    }
    P.built = false;
    if (n_trip == 0) return CMB_OK;
    std::vector<int> trip((size_t)n_trip);
    CK(cudaMemcpy(trip.data(), R.tripped, sizeof(int) * (size_t)n_trip, cudaMemcpyDeviceToHost));
    std::sort(trip.begin(), trip.end());
    std::vector<std::pair<long long, long long>> runs;      // [first, last) frame ranges to process
    for (size_t k = 0; k < trip.size();) {
        size_t e = k + 1;
        while (e < trip.size() && trip[e] == trip[e - 1] + 1) e++;
        runs.push_back({trip[k], (long long)trip[e - 1] + 1});
        k = e;
    }
    return large_cells(ctx, P.job, slot, P.stream, runs);
}

// defer_fallback: the caller (a streamed host trajectory) calls finish_pending(slot) itself before
// it reuses the slot's coordinates; otherwise the fall back is finished before returning.
static int launch_frames(cmb_ctx* ctx, const FrameJob& job, int slot, cudaStream_t stream, bool defer_fallback = false)
{
    if (!ctx->have_system) return fail(ctx, CMB_ERR_STATE, "cmb_set_system has not been called");
    if (job.n_frames <= 0) return CMB_OK;
    if (!job.xyz) return fail(ctx, CMB_ERR_ARG, "xyz is NULL");
    if (job.frame0 + job.n_frames >= (1ll << 24) && (job.atom_keys || job.res_keys))
        return fail(ctx, CMB_ERR_LIMIT, "frame ids in emitted keys are limited to 2^24 per call");
    CK(cudaSetDevice(ctx->device));
    int rc = finish_pending(ctx, slot);                    // (a no-op unless a deferred fall back is still open)
    if (rc != CMB_OK) return rc;
    unsigned long long* counters = job.ext_counters ? job.ext_counters : ctx->d_counters[slot];
    if (!job.ext_counters) CK(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), stream));
    ctx->last_slot = slot;

    const float c2 = ctx->cutoff_f * ctx->cutoff_f;   // float multiply, as in MDTraj
    if ((ctx->atom_log2 > 0 || ctx->res_log2 > 0) && ctx->path == 0)
        return fail(ctx, CMB_ERR_STATE, "hashed count tables are only used by the global-memory cell path");

    // ---- frame-batch pair-list path (cmb_verlet.cuh): plain accumulation into dense tables over a
    //      batch long enough to pay for the list; the frames it cannot take (guard tripped, list
    //      unusable) come back on a device-side list and run through the cell kernels below --------
    const int* frame_list = nullptr;
    const int* frame_count = nullptr;
    bool built_unchecked = false;          // a list was built and its validity has not been looked at yet
    ctx->last_used_verlet = 0;
    {
        const bool plain = !job.atom_keys && !job.res_keys && job.boundary_ulps == 0 && !ctx->count_tests &&
                           (job.atom_table || job.res_table) && ctx->atom_log2 == 0 && ctx->res_log2 == 0;
        const long long min_frames = ctx->path == 0 ? 256 : 32;
        const bool have_shared = job.shared_list && ctx->vl_shared >= 0;
        if (plain && ctx->verlet != 0 && (ctx->verlet == 2 || job.n_frames >= min_frames || have_shared) &&
            job.n_frames >= 2 && job.n_frames < (1ll << 30)) {
            std::string err;
            int launches = 0;
            VlRun& R = ctx->vl_run[slot];
            rc = vl_run_ensure(R, job.n_frames, err);
            if (rc != 0) return fail(ctx, rc, "%s", err.c_str());
            for (int x = 0; x < 2; x++) {
                if (!ctx->vl_built[x]) CK(cudaEventCreateWithFlags(&ctx->vl_built[x], cudaEventDisableTiming));
                if (!ctx->vl_info_ready[x]) CK(cudaEventCreateWithFlags(&ctx->vl_info_ready[x], cudaEventDisableTiming));
                for (int t = 0; t < 2; t++)
                    if (!ctx->vl_walked[x][t]) CK(cudaEventCreateWithFlags(&ctx->vl_walked[x][t], cudaEventDisableTiming));
            }
            // which scratch holds the list this batch walks, and does it have to be built?
            //   ordinary call: the slot's own scratch, built from this batch;
            //   chunk of a streamed trajectory: the shared list if there is a fresh one, else a new
            //   one in the scratch no chunk in flight is walking.
            int lx = slot;
            bool build = true;
            if (job.shared_list) {
                if (have_shared && !ctx->vl_shared_stale) { lx = ctx->vl_shared; build = false; }
                else if (ctx->vl_shared >= 0) lx = ctx->vl_shared ^ 1;
            }
            VlScratch& V = ctx->vl[lx];
            rc = vl_ensure(V, ctx->n, ctx->n_res_c, err);
            if (rc != 0) return fail(ctx, rc, "%s", err.c_str());
            if (ctx->verlet_audit && !ctx->d_vl_audit) {
                CK(cudaMalloc(&ctx->d_vl_audit, 2 * sizeof(unsigned long long)));
                CK(cudaMemset(ctx->d_vl_audit, 0, 2 * sizeof(unsigned long long)));
            }
            VlJob vj;
            vj.xyz = job.xyz; vj.box = job.box; vj.n_frames = job.n_frames; vj.n = ctx->n; vj.n_res_c = ctx->n_res_c;
            vj.meta = ctx->d_meta; vj.res_anchor = ctx->d_res_anchor; vj.cutoff = ctx->cutoff_f; vj.c2 = c2;
            vj.n_ignored = ctx->n_ignored; vj.atom_table = job.atom_table; vj.res_table = job.res_table;
            vj.audit = ctx->verlet_audit; vj.d_audit = ctx->d_vl_audit;
            vj.max_rlist_factor = 0.01f * (float)ctx->verlet_max_rlist_pct;
            vj.tile_share = 0.01f * (float)ctx->verlet_tile_share_pct;
            if (ctx->time_kernels && timing_pair(ctx->kernel_events, ctx->kernel_events_used, &vj.t_begin, &vj.t_end) != 0)
                return fail(ctx, CMB_ERR_CUDA, "cudaEventCreate failed");
            if (build) {
                // earlier walks of the list this scratch held must be over before it is overwritten
                for (int t = 0; t < 2; t++)
                    if (ctx->vl_walked_set[lx][t]) CK(cudaStreamWaitEvent(stream, ctx->vl_walked[lx][t], 0));
                for (int attempt = 0; attempt < 2; attempt++) {
                    vj.conservative_tiles = ctx->vl_conservative;
                    rc = vl_build(vj, V, stream, err, launches, ctx->verlet_graph != 0);
                    if (rc != 0 || ctx->n <= VL_SINGLE_TILE_ATOMS || ctx->vl_conservative) break;
                    // lists of several tiles: a tile that overflowed the frame kernel's shared memory
                    // with the optimistic tile size gets the list rebuilt with the pessimistic one
                    if (job.shared_list && defer_fallback && ctx->path == 0) break;   // (streamed, fused-kernel sizes: no retry)
                    CK(cudaMemcpyAsync(R.h_info + 1, &V.L.hdr->valid, sizeof(int), cudaMemcpyDeviceToHost, stream));
                    CK(cudaMemcpyAsync(R.h_info + 2, &V.L.hdr->fail_code, sizeof(int), cudaMemcpyDeviceToHost, stream));
                    if (job.shared_list && defer_fallback) {
                        // streamed trajectory: never wait for the device here (the copy of the next
                        // chunk would be issued late); an overflowing list sends all its frames to the
                        // fall back, and finish_pending() switches the tile size for the next build
                        built_unchecked = true;
                        break;
                    }
                    CK(cudaStreamSynchronize(stream));
                    if (R.h_info[1] != 0 || R.h_info[2] < 9 || R.h_info[2] > 13) break;
                    ctx->vl_conservative = 1;
                    ctx->stats["vl_rebuilds"] += 1;
                    ctx->stats["vl_rebuild_code"] = R.h_info[2];
                }
                if (rc != 0) return fail(ctx, rc, "%s", err.c_str());
                CK(cudaEventRecord(ctx->vl_built[lx], stream));
                ctx->stats["vl_builds"] += 1;
                if (job.shared_list) { ctx->vl_shared = lx; ctx->vl_shared_stale = 0; }
            } else {
                CK(cudaStreamWaitEvent(stream, ctx->vl_built[lx], 0));
            }
            rc = vl_run_frames(vj, V.L, R, ctx->sm_count, stream, err, launches);
            ctx->stats["kernel_launches"] += launches;
            if (rc != 0) return fail(ctx, rc, "%s", err.c_str());
            CK(cudaEventRecord(ctx->vl_walked[lx][slot], stream));
            ctx->vl_walked_set[lx][slot] = true;
            ctx->last_used_verlet = 1;
            ctx->last_vl_slot = lx;
            frame_list = R.tripped;
            frame_count = R.state;
        }
    }
    if (ctx->path == 0) {
        FrameParams p;
        memset(&p, 0, sizeof(p));
        p.frame_list = frame_list; p.frame_count = frame_count;
        p.xyz = job.xyz; p.box = job.box; p.n_frames = job.n_frames; p.frame0 = job.frame0;
        p.n = ctx->n; p.meta32 = ctx->d_meta32;
        p.cutoff = ctx->cutoff_f; p.c2 = c2;
        p.boundary_ulps = job.boundary_ulps;
        p.n_ignored = ctx->n_ignored; p.n_res_c = ctx->n_res_c;
        p.maxc = ctx->fs_maxc; p.bm_words = ctx->fs_bm_words; p.bm_in_smem = ctx->fs_bm_in_smem;
        p.pair_tasks = ctx->pair_tasks;
        p.atom_table = job.atom_table; p.res_table = job.res_table;
        p.atom_keys = job.atom_keys; p.atom_cap = job.atom_cap;
        p.res_keys = job.res_keys; p.res_cap = job.res_cap;
        p.counters = counters;
        p.boundary = job.boundary; p.boundary_cap = job.boundary_cap;
        long long grid = (long long)ctx->sm_count * ctx->fs_ctas_per_sm;
        if (grid > job.n_frames) grid = job.n_frames;
        if (!p.bm_in_smem) {
            size_t need = (size_t)grid * (size_t)p.bm_words * 4;
            rc = ensure_bytes(ctx, (void**)&ctx->d_bitmap[slot], &ctx->bitmap_bytes[slot], need);
            if (rc) return rc;
            p.g_bitmap = ctx->d_bitmap[slot];
        }
        const unsigned g = (unsigned)grid;
        const size_t sm = (size_t)ctx->fs_smem;
        if (ctx->fs_threads == 1024) fs_launch<1024>(ctx, job, p, g, sm, stream);
        else fs_launch<512>(ctx, job, p, g, sm, stream);
        CK(cudaGetLastError());
        ctx->stats["kernel_launches"] += 1;
        if (frame_list != nullptr && job.shared_list) {
            // streamed trajectory: the tripped count of this chunk steers the rebuild of the shared
            // list; it is looked at when the slot comes round again (never waited for)
            cmb_ctx::Pending& P = ctx->pending[slot];
            if (P.info && cudaEventQuery(ctx->vl_info_ready[slot]) == cudaSuccess &&
                ctx->vl_run[slot].h_info[0] * 8 > ctx->vl_run[slot].h_info[3])
                ctx->vl_shared_stale = 1;
            ctx->vl_run[slot].h_info[3] = (int)job.n_frames;
            CK(cudaMemcpyAsync(ctx->vl_run[slot].h_info, frame_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
            CK(cudaEventRecord(ctx->vl_info_ready[slot], stream));
            P.info = true;
        }
        return CMB_OK;
    }
    // global-memory cell path
    if (frame_list != nullptr) {
        cmb_ctx::Pending& P = ctx->pending[slot];
        CK(cudaMemcpyAsync(ctx->vl_run[slot].h_info, frame_count, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaEventRecord(ctx->vl_info_ready[slot], stream));
        P.active = true; P.job = job; P.stream = stream; P.built = built_unchecked;
        return defer_fallback ? CMB_OK : finish_pending(ctx, slot);
    }
    std::vector<std::pair<long long, long long>> runs;
    runs.push_back({0, job.n_frames});
    return large_cells(ctx, job, slot, stream, runs);
}

// global-memory cell kernels over runs of consecutive frames of a job
static int large_cells(cmb_ctx* ctx, const FrameJob& job, int slot, cudaStream_t stream,
                       const std::vector<std::pair<long long, long long>>& runs)
{
    const float c2 = ctx->cutoff_f * ctx->cutoff_f;
    unsigned long long* counters = job.ext_counters ? job.ext_counters : ctx->d_counters[slot];
    LargeJob lj;
    lj.xyz = job.xyz; lj.box = job.box; lj.n_frames = job.n_frames; lj.frame0 = job.frame0;
    lj.n = ctx->n; lj.meta = ctx->d_meta; lj.cutoff = ctx->cutoff_f; lj.c2 = c2;
    lj.boundary_ulps = job.boundary_ulps;
    lj.n_ignored = ctx->n_ignored; lj.n_res_c = ctx->n_res_c;
    lj.atom_table = job.atom_table; lj.res_table = job.res_table;
    lj.atom_keys = job.atom_keys; lj.atom_cap = job.atom_cap;
    lj.res_keys = job.res_keys; lj.res_cap = job.res_cap;
    lj.counters = counters; lj.boundary = job.boundary; lj.boundary_cap = job.boundary_cap;
    lj.count_tests = ctx->count_tests;
    lj.all_roles = ctx->all_roles ? 1 : 0;
    lj.atom_log2 = ctx->atom_log2; lj.res_log2 = ctx->res_log2; lj.hstats = ctx->d_hstats;
    lj.ovf_list = ctx->d_ovf; lj.ovf_cap = ctx->ovf_cap;
    lj.batch_mode = ctx->large_batch;
    for (const auto& run : runs) {
        LargeJob part = lj;
        part.xyz = lj.xyz + (size_t)run.first * (size_t)ctx->n * 3;
        part.box = lj.box ? lj.box + (size_t)run.first * 9 : nullptr;
        part.n_frames = run.second - run.first;
        part.frame0 = lj.frame0 + run.first;
        std::string err;
        int launches = 0;
        int rc = large_run(part, ctx->large[slot], ctx->sm_count, stream, err, launches);
        ctx->stats["kernel_launches"] += launches;
        if (rc != CMB_OK) return fail(ctx, rc, "%s", err.c_str());
    }
    return CMB_OK;
}

extern "C" int cmb_accumulate(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                              uint32_t* d_atom_table, uint32_t* d_res_table, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    FrameJob job;
    memset(&job, 0, sizeof(job));
    job.xyz = d_xyz; job.box = d_box; job.n_frames = n_frames;
    job.atom_table = d_atom_table; job.res_table = d_res_table;
    ctx->slot ^= 1;
    return launch_frames(ctx, job, ctx->slot, (cudaStream_t)stream);
}

static int64_t pick_chunk(const cmb_ctx* ctx, int64_t n_frames, int64_t chunk_frames, bool* automatic = nullptr)
{
    const size_t frame_bytes = (size_t)ctx->n * 12;
    if (chunk_frames <= 0 && ctx->stage_chunk_frames > 0) chunk_frames = ctx->stage_chunk_frames;
    if (automatic) *automatic = chunk_frames <= 0;
    if (chunk_frames <= 0) {
        chunk_frames = (int64_t)((80ull << 20) / frame_bytes);
        // at least four chunks so copies and kernels overlap, but keep every SM busy
        const int64_t quarter = (n_frames + 3) / 4;
        if (chunk_frames > quarter) chunk_frames = quarter;
        const int64_t fill = (int64_t)ctx->sm_count * std::max(1, ctx->fs_ctas_per_sm) * 2;
        if (chunk_frames < fill) chunk_frames = fill;
    }
    if (chunk_frames > n_frames) chunk_frames = n_frames;
    if (chunk_frames < 1) chunk_frames = 1;
    return chunk_frames;
}

// Chunk lengths of a streamed trajectory.  The copies run back to back on the copy engine, which makes
// two places of the schedule matter when the chunk size is chosen automatically:
//   head: the pair list is built from the first chunk, by ~30 small dependent launches whose launch
//         latency grows several-fold while a copy saturates the link; the copy of chunk 2 has to wait
//         for the kernels of chunk 0 (same staging buffer), so the first chunk is a short one -- the
//         build then hides under the copy of chunk 1;
//   tail: the call ends one chunk's worth of kernels after the LAST copy, so the final chunk is split
//         and only a quarter of it (but enough frames to occupy every SM) is left for that tail.
static std::vector<int64_t> chunk_schedule(const cmb_ctx* ctx, int64_t n_frames, int64_t chunk_frames, bool automatic)
{
    std::vector<int64_t> sizes;
    // fewest frames worth a chunk of their own: the fused kernels parallelise over frames, the
    // global-memory sizes over tiles x frames (and their list build outlasts the copy of a small chunk)
    const int64_t fill = (int64_t)ctx->sm_count * std::max(1, ctx->fs_ctas_per_sm) * 2;
    const bool large = ctx->path == 1;
    int64_t f0 = 0;
    if (automatic && n_frames >= 3 * chunk_frames) {
        const int64_t head = large ? 32 : std::max<int64_t>(fill, chunk_frames / 4);
        if (chunk_frames >= 2 * head) { sizes.push_back(head); f0 = head; }
    }
    for (; f0 < n_frames; f0 += chunk_frames) sizes.push_back(std::min<int64_t>(chunk_frames, n_frames - f0));
    if (automatic && sizes.size() >= 2) {
        const int64_t last = sizes.back();
        const int64_t tail = std::max<int64_t>(large ? 32 : fill, last / 4);
        if (last >= 2 * tail) {
            sizes.back() = last - tail;
            sizes.push_back(tail);
        }
    }
    return sizes;
}

static int ensure_staging(cmb_ctx* ctx, int64_t chunk_frames, bool with_box, int64_t box_frames)
{
    const size_t need_xyz = (size_t)chunk_frames * (size_t)ctx->n * 12;
    const size_t need_box = (size_t)box_frames * 36;
    if (ctx->stage_xyz_bytes < need_xyz) {
        for (int s = 0; s < 2; s++) {
            if (ctx->d_stage_xyz[s]) CK(cudaFree(ctx->d_stage_xyz[s]));
            ctx->d_stage_xyz[s] = nullptr;
            CK(cudaMalloc(&ctx->d_stage_xyz[s], need_xyz + 256));
        }
        ctx->stage_xyz_bytes = need_xyz;
    }
    if (with_box && ctx->stage_box_bytes < need_box) {
        for (int s = 0; s < 2; s++) {
            if (ctx->d_stage_box[s]) CK(cudaFree(ctx->d_stage_box[s]));
            ctx->d_stage_box[s] = nullptr;
            CK(cudaMalloc(&ctx->d_stage_box[s], need_box + 256));
        }
        ctx->stage_box_bytes = need_box;
    }
    return CMB_OK;
}

// pack out[f][i][:] = full[f][used[i]][:] for frames [f0, f0 + nf), split over host threads
static void gather_chunk(const float* full, int n_total, const int32_t* used, int n_used, int64_t f0,
                         int64_t nf, float* out, int n_threads)
{
    auto work = [=](int64_t a, int64_t b) {
        if (used == nullptr) {                   // all atoms: one contiguous block
            memcpy(out + (size_t)a * (size_t)n_used * 3, full + (size_t)(f0 + a) * (size_t)n_total * 3,
                   (size_t)(b - a) * (size_t)n_used * 12);
            return;
        }
        for (int64_t f = a; f < b; f++) {
            const float* src = full + (size_t)(f0 + f) * (size_t)n_total * 3;
            float* dst = out + (size_t)f * (size_t)n_used * 3;
            int i = 0;
            while (i < n_used) {                 // copy maximal runs of consecutive atom indices
                int j = i + 1;
                while (j < n_used && used[j] == used[j - 1] + 1) j++;
                memcpy(dst + (size_t)i * 3, src + (size_t)used[i] * 3, (size_t)(j - i) * 12);
                i = j;
            }
        }
    };
    if (n_threads <= 1 || nf < 2 * n_threads) { work(0, nf); return; }
    std::vector<std::thread> pool;
    const int64_t per = (nf + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        const int64_t a = t * per, b = std::min<int64_t>(nf, a + per);
        if (a < b) pool.emplace_back(work, a, b);
    }
    for (auto& th : pool) th.join();
}

static bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes attr;
    const cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type != cudaMemoryTypeUnregistered;
}

// Stream a HOST trajectory through the frame kernels, chunk by chunk over two streams: the copy of
// chunk k + 1 overlaps the kernels of chunk k.  `proto` says what the kernels produce (count tables
// and / or per-frame keys appended to caller-owned device buffers through proto.ext_counters).
//   direct: the frames are read by cudaMemcpyAsync straight from the caller's (pinned) buffer;
//   staged: host threads pack chunk k + 1 into the context's pinned staging buffers -- gathering
//           the used atoms if h_used != nullptr -- while chunk k is copied and processed (pageable
//           memory, memory-mapped files, trajectories that still hold all atoms).
static int stream_host_frames(cmb_ctx* ctx, const float* h_xyz_full, int n_total, const int32_t* h_used,
                              const float* h_box, int64_t n_frames, int64_t chunk_frames, int n_threads,
                              const FrameJob& proto)
{
    CK(cudaSetDevice(ctx->device));
    const bool direct = h_used == nullptr && n_total == ctx->n &&
                        (!ctx->stage_pageable || host_pointer_is_pinned(h_xyz_full));
    if (n_threads <= 0) {
        n_threads = (int)std::thread::hardware_concurrency();
        if (n_threads > 16) n_threads = 16;
        if (n_threads < 1) n_threads = 1;
    }
    const size_t frame_bytes = (size_t)ctx->n * 12;
    bool automatic = false;
    chunk_frames = pick_chunk(ctx, n_frames, chunk_frames, &automatic);
    const std::vector<int64_t> schedule = chunk_schedule(ctx, n_frames, chunk_frames, automatic);
    // the unit cells of ALL frames go over in one copy ahead of the first chunk (a small copy per chunk
    // costs the copy engine ~20 us each while the link is saturated)
    const bool box_once = h_box != nullptr && (size_t)n_frames * 36 <= ((size_t)256 << 20);
    int rc0 = ensure_staging(ctx, chunk_frames, h_box != nullptr, box_once ? n_frames : chunk_frames);
    if (rc0 != CMB_OK) return rc0;
    // the chunks share ONE pair list, built from the first chunk (cmb_verlet.cuh: the guard decides
    // frame by frame whether the list still covers a frame, whichever batch it was built from)
    ctx->vl_shared = -1;
    ctx->vl_shared_stale = 0;
    for (int s = 0; s < 2; s++) ctx->pending[s].info = false;
    if (!direct) {
        const size_t need_pinned = (size_t)chunk_frames * frame_bytes;
        if (ctx->pinned_bytes < need_pinned) {
            for (int s = 0; s < 2; s++) {
                if (ctx->h_pinned[s]) CK(cudaFreeHost(ctx->h_pinned[s]));
                ctx->h_pinned[s] = nullptr;
                CK(cudaMallocHost(&ctx->h_pinned[s], need_pinned));
            }
            ctx->pinned_bytes = need_pinned;
        }
        for (int s = 0; s < 2; s++)
            if (!ctx->copy_done[s]) CK(cudaEventCreateWithFlags(&ctx->copy_done[s], cudaEventDisableTiming));
    }
    if (box_once) {
        if (!ctx->box_ready) CK(cudaEventCreateWithFlags(&ctx->box_ready, cudaEventDisableTiming));
        CK(cudaMemcpyAsync(ctx->d_stage_box[0], h_box, (size_t)n_frames * 36, cudaMemcpyHostToDevice, ctx->streams[0]));
        CK(cudaEventRecord(ctx->box_ready, ctx->streams[0]));
        CK(cudaStreamWaitEvent(ctx->streams[1], ctx->box_ready, 0));
    }
    int c = 0;
    bool ext_waited[2] = {false, false};
    if (ctx->stage_trace) {
        ctx->trace_used = 0;
        if (!ctx->trace_begin) CK(cudaEventCreate(&ctx->trace_begin));
        CK(cudaEventRecord(ctx->trace_begin, ctx->streams[0]));
    }
    for (int64_t f0 = 0; f0 < n_frames; f0 += schedule[c], c++) {
        const int s = c & 1;
        const int64_t nf = schedule[c];
        const float* src = h_xyz_full + (size_t)f0 * (size_t)ctx->n * 3;
        {   // the frames chunk c - 2 left to the cell kernels still sit in staging buffer s
            int rc = finish_pending(ctx, s);
            if (rc != CMB_OK) return rc;
        }
        if (!direct) {
            if (c >= 2) CK(cudaEventSynchronize(ctx->copy_done[s]));   // pinned buffer s is free again
            gather_chunk(h_xyz_full, n_total, h_used, ctx->n, f0, nf, ctx->h_pinned[s], n_threads);
            src = ctx->h_pinned[s];
        }
        // stream order on streams[s] serialises reuse of staging buffer s
        CK(cudaMemcpyAsync(ctx->d_stage_xyz[s], src, (size_t)nf * frame_bytes, cudaMemcpyHostToDevice, ctx->streams[s]));
        if (!direct) CK(cudaEventRecord(ctx->copy_done[s], ctx->streams[s]));
        if (h_box && !box_once)
            CK(cudaMemcpyAsync(ctx->d_stage_box[s], h_box + (size_t)f0 * 9, (size_t)nf * 36,
                               cudaMemcpyHostToDevice, ctx->streams[s]));
        if (ctx->ext_ready_pending && !ext_waited[s]) {
            // the caller's earlier work (zero-filling the tables): kernels wait for it, the copy above did not
            CK(cudaStreamWaitEvent(ctx->streams[s], ctx->ext_ready, 0));
            ext_waited[s] = true;
        }
        cudaEvent_t t_copy = nullptr, t_done = nullptr;
        if (ctx->stage_trace) {
            if (timing_pair(ctx->trace_events, ctx->trace_used, &t_copy, &t_done) != 0)
                return fail(ctx, CMB_ERR_CUDA, "cudaEventCreate failed");
            CK(cudaEventRecord(t_copy, ctx->streams[s]));
        }
        FrameJob job = proto;
        job.xyz = ctx->d_stage_xyz[s];
        job.box = !h_box ? nullptr : box_once ? ctx->d_stage_box[0] + (size_t)f0 * 9 : ctx->d_stage_box[s];
        job.n_frames = nf; job.frame0 = proto.frame0 + f0;
        job.shared_list = 1;
        int rc = launch_frames(ctx, job, s, ctx->streams[s], true);
        if (rc != CMB_OK) return rc;
        if (t_done) CK(cudaEventRecord(t_done, ctx->streams[s]));
    }
    for (int s = 0; s < 2; s++) {
        int rc = finish_pending(ctx, s);
        if (rc != CMB_OK) return rc;
    }
    CK(cudaStreamSynchronize(ctx->streams[0]));
    CK(cudaStreamSynchronize(ctx->streams[1]));
    ctx->ext_ready_pending = false;
    return CMB_OK;
}

extern "C" int cmb_wait_for_stream(cmb_ctx* ctx, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->ext_ready) CK(cudaEventCreateWithFlags(&ctx->ext_ready, cudaEventDisableTiming));
    CK(cudaEventRecord(ctx->ext_ready, (cudaStream_t)stream));
    ctx->ext_ready_pending = true;
    return CMB_OK;
}

static int check_used(cmb_ctx* ctx, const int32_t* h_used, int n_total)
{
    for (int i = 0; i < ctx->n; i++)
        if (h_used[i] < 0 || h_used[i] >= n_total || (i > 0 && h_used[i] <= h_used[i - 1]))
            return fail(ctx, CMB_ERR_ARG, "used atom indices must be ascending and inside [0, n_total)");
    return CMB_OK;
}

extern "C" int cmb_accumulate_host(cmb_ctx* ctx, const float* h_xyz, const float* h_box, int64_t n_frames,
                                   int64_t chunk_frames, uint32_t* d_atom_table, uint32_t* d_res_table)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!ctx->have_system) return fail(ctx, CMB_ERR_STATE, "cmb_set_system has not been called");
    if (n_frames <= 0) return CMB_OK;
    if (!h_xyz) return fail(ctx, CMB_ERR_ARG, "xyz is NULL");
    FrameJob proto;
    memset(&proto, 0, sizeof(proto));
    proto.atom_table = d_atom_table; proto.res_table = d_res_table;
    return stream_host_frames(ctx, h_xyz, ctx->n, nullptr, h_box, n_frames, chunk_frames, 0, proto);
}

extern "C" int cmb_accumulate_host_gather(cmb_ctx* ctx, const float* h_xyz_full, int n_total,
                                          const int32_t* h_used, const float* h_box, int64_t n_frames,
                                          int64_t chunk_frames, int n_threads, uint32_t* d_atom_table,
                                          uint32_t* d_res_table)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!ctx->have_system) return fail(ctx, CMB_ERR_STATE, "cmb_set_system has not been called");
    if (n_frames <= 0) return CMB_OK;
    if (!h_xyz_full || !h_used || n_total < ctx->n) return fail(ctx, CMB_ERR_ARG, "bad gather arguments");
    int rc = check_used(ctx, h_used, n_total);
    if (rc != CMB_OK) return rc;
    FrameJob proto;
    memset(&proto, 0, sizeof(proto));
    proto.atom_table = d_atom_table; proto.res_table = d_res_table;
    return stream_host_frames(ctx, h_xyz_full, n_total, h_used, h_box, n_frames, chunk_frames, n_threads, proto);
}

// Per-frame contact keys of a HOST trajectory (the host-side twin of cmb_frame_contacts): the
// frames stream through the key kernels exactly as in cmb_accumulate_host[_gather]; every chunk
// APPENDS its keys -- which carry trajectory frame numbers frame0 + f -- to the caller's device
// buffers through the shared counters d_counts, so no chunk waits for the host.  The caller
// sorts the buffers afterwards (cmb_sort_keys) and turns them into CSR arrays (cmb_keys_to_csr).
extern "C" int cmb_frame_contacts_host(cmb_ctx* ctx, const float* h_xyz_full, int n_total, const int32_t* h_used,
                                       const float* h_box, int64_t n_frames, int64_t frame0, int64_t chunk_frames,
                                       int n_threads, uint64_t* d_atom_keys, int64_t atom_cap,
                                       uint64_t* d_res_keys, int64_t res_cap, uint64_t* d_counts)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!ctx->have_system) return fail(ctx, CMB_ERR_STATE, "cmb_set_system has not been called");
    if (!d_counts) return fail(ctx, CMB_ERR_ARG, "d_counts is NULL");
    if (n_frames <= 0) return CMB_OK;
    if (!h_xyz_full || n_total < ctx->n || (h_used == nullptr && n_total != ctx->n))
        return fail(ctx, CMB_ERR_ARG, "bad trajectory arguments");
    if (h_used) {
        int rc = check_used(ctx, h_used, n_total);
        if (rc != CMB_OK) return rc;
    }
    if (frame0 < 0 || frame0 + n_frames >= (1ll << 24))
        return fail(ctx, CMB_ERR_LIMIT, "frame ids in emitted keys are limited to 2^24 per call");
    FrameJob proto;
    memset(&proto, 0, sizeof(proto));
    proto.frame0 = frame0;
    proto.atom_keys = (unsigned long long*)d_atom_keys; proto.atom_cap = d_atom_keys ? atom_cap : 0;
    proto.res_keys = (unsigned long long*)d_res_keys; proto.res_cap = d_res_keys ? res_cap : 0;
    proto.ext_counters = (unsigned long long*)d_counts;
    return stream_host_frames(ctx, h_xyz_full, n_total, h_used, h_box, n_frames, chunk_frames, n_threads, proto);
}

// ---------------------------------------------------------------------------------
// hashed count tables
// ---------------------------------------------------------------------------------
extern "C" int cmb_set_table_mode(cmb_ctx* ctx, int atom_log2, int res_log2)
{
    if (!ctx) return CMB_ERR_ARG;
    if (atom_log2 < 0 || atom_log2 > 31 || res_log2 < 0 || res_log2 > 31 || (atom_log2 > 0 && atom_log2 < 5) ||
        (res_log2 > 0 && res_log2 < 5))
        return fail(ctx, CMB_ERR_ARG, "table sizes must be 0 (dense) or 2^5 .. 2^31 slots");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->d_hstats) {
        CK(cudaMalloc(&ctx->d_hstats, 4 * sizeof(unsigned long long)));
        CK(cudaMemset(ctx->d_hstats, 0, 4 * sizeof(unsigned long long)));
    }
    ctx->atom_log2 = atom_log2;
    ctx->res_log2 = res_log2;
    return CMB_OK;
}

extern "C" int cmb_hash_overflow_buffer(cmb_ctx* ctx, uint64_t* d_list, int64_t cap)
{
    if (!ctx || cap < 0 || (cap > 0 && !d_list)) return CMB_ERR_ARG;
    ctx->d_ovf = reinterpret_cast<unsigned long long*>(d_list);
    ctx->ovf_cap = cap;
    return CMB_OK;
}

extern "C" int cmb_hash_add(cmb_ctx* ctx, uint64_t* d_slots, int log2_slots, const uint64_t* d_key1,
                            const uint64_t* d_count, int64_t n, int which, void* stream)
{
    if (!ctx || !d_slots || log2_slots < 5 || log2_slots > 31 || n < 0 || which < 0 || which > 1) return CMB_ERR_ARG;
    if (!ctx->d_hstats) return fail(ctx, CMB_ERR_STATE, "cmb_set_table_mode has not been called");
    if (n == 0) return CMB_OK;
    CK(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    k_hash_add<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(d_slots),
                                                        (1u << log2_slots) - 1u,
                                                        reinterpret_cast<const unsigned long long*>(d_key1),
                                                        reinterpret_cast<const unsigned long long*>(d_count), n,
                                                        ctx->d_hstats, which);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_hash_stats(cmb_ctx* ctx, int64_t* new_atom_keys, int64_t* new_res_keys, int64_t* overflow)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!ctx->d_hstats) return fail(ctx, CMB_ERR_STATE, "cmb_set_table_mode has not been called");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    unsigned long long h[4];
    CK(cudaMemcpy(h, ctx->d_hstats, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaMemset(ctx->d_hstats, 0, sizeof(h)));
    if (new_atom_keys) *new_atom_keys = (int64_t)h[0];
    if (new_res_keys) *new_res_keys = (int64_t)h[1];
    if (overflow) *overflow = (int64_t)h[2];
    if (h[3]) return fail(ctx, CMB_ERR_LIMIT, "%llu entries could not be re-inserted into a hashed table", h[3]);
    return CMB_OK;
}

extern "C" int cmb_export_hashed(cmb_ctx* ctx, const uint64_t* d_slots, int log2_slots, int64_t cap,
                                 int64_t* d_idx, uint32_t* d_cnt, uint64_t* d_n, void* stream)
{
    if (!ctx || !d_slots || !d_n || log2_slots < 5 || log2_slots > 31) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const long long n_slots = 1ll << log2_slots;
    const int grid = (int)std::min<long long>((n_slots + 255) / 256, (long long)ctx->sm_count * 16);
    k_export_hashed<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(d_slots), n_slots, cap, reinterpret_cast<long long*>(d_idx), d_cnt,
        reinterpret_cast<unsigned long long*>(d_n));
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_hash_rehash(cmb_ctx* ctx, const uint64_t* d_src, int src_log2, uint64_t* d_dst, int dst_log2,
                               void* stream)
{
    if (!ctx || !d_src || !d_dst || src_log2 < 5 || dst_log2 < src_log2 || dst_log2 > 31) return CMB_ERR_ARG;
    if (!ctx->d_hstats) return fail(ctx, CMB_ERR_STATE, "cmb_set_table_mode has not been called");
    CK(cudaSetDevice(ctx->device));
    const long long n_src = 1ll << src_log2;
    const int grid = (int)std::min<long long>((n_src + 255) / 256, (long long)ctx->sm_count * 16);
    k_rehash<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_src), n_src,
                                                      reinterpret_cast<unsigned long long*>(d_dst),
                                                      (1u << dst_log2) - 1u, ctx->d_hstats);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_frame_contacts(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                                  int64_t frame0, uint64_t* d_atom_keys, int64_t atom_cap,
                                  uint64_t* d_res_keys, int64_t res_cap, uint64_t* d_counts,
                                  uint32_t* d_atom_table, uint32_t* d_res_table, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_counts) return fail(ctx, CMB_ERR_ARG, "d_counts is NULL");
    FrameJob job;
    memset(&job, 0, sizeof(job));
    job.xyz = d_xyz; job.box = d_box; job.n_frames = n_frames; job.frame0 = frame0;
    job.atom_table = d_atom_table; job.res_table = d_res_table;
    job.atom_keys = (unsigned long long*)d_atom_keys; job.atom_cap = d_atom_keys ? atom_cap : 0;
    job.res_keys = (unsigned long long*)d_res_keys; job.res_cap = d_res_keys ? res_cap : 0;
    job.ext_counters = (unsigned long long*)d_counts;
    ctx->slot ^= 1;
    return launch_frames(ctx, job, ctx->slot, (cudaStream_t)stream);
}

extern "C" int cmb_keys_to_csr(cmb_ctx* ctx, const uint64_t* d_sorted_keys, int64_t n_keys, int64_t n_frames,
                               int64_t frame0, const int32_t* d_index_map, int64_t* d_ptr, int32_t* d_i,
                               int32_t* d_j, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (n_keys < 0 || n_frames < 1 || !d_ptr || (n_keys > 0 && (!d_sorted_keys || !d_i || !d_j)))
        return fail(ctx, CMB_ERR_ARG, "bad CSR arguments");
    CK(cudaSetDevice(ctx->device));
    const long long blocks = std::max<long long>(1, (n_keys + 255) / 256);
    k_keys_to_csr<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long*>(d_sorted_keys), n_keys, n_frames, frame0, d_index_map,
        reinterpret_cast<long long*>(d_ptr), d_i, d_j);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_boundary_pairs(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                                  int64_t frame0, int ulps, void* d_records, int64_t cap, uint64_t* d_n,
                                  void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (ulps < 1 || !d_records || !d_n) return fail(ctx, CMB_ERR_ARG, "bad boundary arguments");
    // the kernel writes the record count to counters[2]; give it a private 4-slot block
    CK(cudaSetDevice(ctx->device));
    ctx->slot ^= 1;
    FrameJob job;
    memset(&job, 0, sizeof(job));
    job.xyz = d_xyz; job.box = d_box; job.n_frames = n_frames; job.frame0 = frame0;
    job.boundary = (float4*)d_records; job.boundary_cap = cap; job.boundary_ulps = ulps;
    int rc = launch_frames(ctx, job, ctx->slot, (cudaStream_t)stream);
    if (rc != CMB_OK) return rc;
    CK(cudaMemcpyAsync(d_n, ctx->d_counters[ctx->slot] + 2, sizeof(unsigned long long),
                       cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CMB_OK;
}

// Audit of the screening shortcut of the cell kernels: every candidate pair of the
// stencil that is not residue-excluded is decided by the reference arithmetic (nl_d2) AND by the
// production thresholds of its frame; d_out[0] = pairs audited, d_out[1] = pairs on which the screen
// alone (d2 <= c2_lo: contact, d2 >= c2_hi: no contact) would have decided differently -- must be 0.
extern "C" int cmb_screen_audit(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                                uint64_t* d_out, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_out) return fail(ctx, CMB_ERR_ARG, "d_out is NULL");
    CK(cudaSetDevice(ctx->device));
    ctx->slot ^= 1;
    FrameJob job;
    memset(&job, 0, sizeof(job));
    job.xyz = d_xyz; job.box = d_box; job.n_frames = n_frames;
    job.boundary_ulps = -1;
    int rc = launch_frames(ctx, job, ctx->slot, (cudaStream_t)stream);
    if (rc != CMB_OK) return rc;
    CK(cudaMemcpyAsync(d_out, ctx->d_counters[ctx->slot] + 2, 2 * sizeof(unsigned long long),
                       cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CMB_OK;
}

// ---------------------------------------------------------------------------------
// sort / export / gather / synth
// ---------------------------------------------------------------------------------
// in-place radix sort of 64-bit keys on the bits [begin_bit, end_bit)
static int sort_key_bits(cmb_ctx* ctx, uint64_t* d_keys, int64_t n, int begin_bit, int end_bit, cudaStream_t stream)
{
    if (n <= 1) return CMB_OK;
    if (n > 0x7fffffffll) return fail(ctx, CMB_ERR_LIMIT, "too many keys for one sort");
    CK(cudaSetDevice(ctx->device));
    size_t tmp = 0;
    unsigned long long* keys = (unsigned long long*)d_keys;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys, keys, (int)n, begin_bit, end_bit, stream));
    const size_t alt = ((size_t)n * 8 + 255) & ~(size_t)255;
    int rc = ensure_bytes(ctx, &ctx->d_tmp, &ctx->tmp_bytes, alt + tmp);
    if (rc) return rc;
    unsigned long long* out = (unsigned long long*)ctx->d_tmp;
    CK(cub::DeviceRadixSort::SortKeys((char*)ctx->d_tmp + alt, tmp, keys, out, (int)n, begin_bit, end_bit, stream));
    CK(cudaMemcpyAsync(keys, out, (size_t)n * 8, cudaMemcpyDeviceToDevice, stream));
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_sort_keys(cmb_ctx* ctx, uint64_t* d_keys, int64_t n, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    return sort_key_bits(ctx, d_keys, n, 0, 64, (cudaStream_t)stream);
}

extern "C" int cmb_export_nonzero(cmb_ctx* ctx, const uint32_t* d_table, int64_t n_elems, int64_t cap,
                                  int64_t* d_idx, uint32_t* d_cnt, uint64_t* d_n, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_table || !d_idx || !d_cnt || !d_n || n_elems < 0) return fail(ctx, CMB_ERR_ARG, "bad export arguments");
    if (((uintptr_t)d_table & 15) != 0) return fail(ctx, CMB_ERR_ARG, "table must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    if (n_elems == 0) return CMB_OK;
    long long blocks = (n_elems / 4 + 255) / 256;
    const long long maxb = (long long)ctx->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    k_export_nonzero<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_table, n_elems, cap, (long long*)d_idx, d_cnt, (unsigned long long*)d_n);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

// Sorted sparse export: d_packed[cap] receives (flat index << 24 | count) of the non-zero entries
// in ascending index order, followed by 0xff..ff padding; d_n[0] = entries found (entries beyond
// cap are dropped: retry with a larger buffer), d_n[1] != 0 if a count did not fit 24 bits.
// log2_slots > 0: d_table is a hashed table of 2^log2_slots uint64 slots, else a dense uint32 table.
extern "C" int cmb_export_sorted(cmb_ctx* ctx, const void* d_table, int64_t n_elems, int log2_slots, int64_t cap,
                                 uint64_t* d_packed, uint64_t* d_n, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_table || !d_packed || !d_n || n_elems < 0 || cap < 1) return fail(ctx, CMB_ERR_ARG, "bad export arguments");
    if (((uintptr_t)d_table & 15) != 0) return fail(ctx, CMB_ERR_ARG, "table must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (log2_slots <= 0) {
        // dense table: ordered two-pass export, no sort
        CK(cudaMemsetAsync(d_n, 0, 16, st));
        const long long nvec = n_elems >> 2;
        long long blocks = std::min<long long>((nvec + 32 * 8 * 8 - 1) / (32 * 8 * 8), (long long)ctx->sm_count * 8);
        if (blocks < 1) blocks = 1;
        const long long n_warps = blocks * (EO_THREADS / 32);
        long long span = (nvec + n_warps - 1) / n_warps;
        span = (span + 1023) / 1024 * 1024;                // whole flag words (32 warp iterations of 32 vectors)
        if (span < 1024) span = 1024;
        const size_t flag_words = (size_t)n_warps * (size_t)(span / 1024);
        int rc = ensure_bytes(ctx, &ctx->d_tmp, &ctx->tmp_bytes, (size_t)n_warps * sizeof(int) + flag_words * sizeof(unsigned));
        if (rc) return rc;
        int* counts = reinterpret_cast<int*>(ctx->d_tmp);
        unsigned int* flags = reinterpret_cast<unsigned int*>(counts + n_warps);
        k_export_count<<<(unsigned)blocks, EO_THREADS, 0, st>>>(reinterpret_cast<const unsigned int*>(d_table), n_elems,
                                                                span, counts, flags);
        k_export_write<<<(unsigned)blocks, EO_THREADS, 0, st>>>(reinterpret_cast<const unsigned int*>(d_table), n_elems,
                                                                span, counts, flags, cap,
                                                                reinterpret_cast<unsigned long long*>(d_packed),
                                                                reinterpret_cast<unsigned long long*>(d_n));
        CK(cudaGetLastError());
        ctx->stats["kernel_launches"] += 2;
        return CMB_OK;
    }
    CK(cudaMemsetAsync(d_packed, 0xff, (size_t)cap * 8, st));
    CK(cudaMemsetAsync(d_n, 0, 16, st));
    {
        const long long n_slots = 1ll << log2_slots;
        const int grid = (int)std::min<long long>((n_slots + 255) / 256, (long long)ctx->sm_count * 16);
        k_export_hashed<true><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_table), n_slots, cap,
                                                    reinterpret_cast<long long*>(d_packed), nullptr,
                                                    reinterpret_cast<unsigned long long*>(d_n));
    }
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    // the indices are distinct: sort on the index bits only (the all-ones padding sorts last)
    return sort_key_bits(ctx, d_packed, cap, 24, 64, st);
}

__global__ void __launch_bounds__(256) k_unpack_sorted(const unsigned long long* __restrict__ packed,
                                                       const unsigned long long* __restrict__ d_n, long long cap,
                                                       long long idx_offset, long long* __restrict__ idx,
                                                       unsigned int* __restrict__ cnt)
{
    const long long n = min((long long)d_n[0], cap);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const unsigned long long w = packed[k];
        idx[k] = (long long)(w >> 24) + idx_offset;
        cnt[k] = (unsigned int)(w & 0xffffffull);
    }
}

extern "C" int cmb_unpack_sorted(cmb_ctx* ctx, const uint64_t* d_packed, const uint64_t* d_n, int64_t cap,
                                 int64_t idx_offset, int64_t* d_idx, uint32_t* d_cnt, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_packed || !d_n || !d_idx || !d_cnt || cap < 1) return fail(ctx, CMB_ERR_ARG, "bad unpack arguments");
    CK(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<long long>((cap + 255) / 256, (long long)ctx->sm_count * 8);
    k_unpack_sorted<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_packed),
                                                           reinterpret_cast<const unsigned long long*>(d_n), cap, idx_offset,
                                                           reinterpret_cast<long long*>(d_idx), d_cnt);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

__global__ void __launch_bounds__(256) k_unpack_pairs(const unsigned long long* __restrict__ packed,
                                                      const unsigned long long* __restrict__ d_n, long long cap,
                                                      long long n_cols0, long long off1, long long n_cols1,
                                                      int* __restrict__ lo, int* __restrict__ hi,
                                                      unsigned int* __restrict__ cnt, unsigned long long* __restrict__ d_split)
{
    const long long n = min((long long)d_n[0], cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // entries of the first table: lower bound of flat index off1 in the sorted words
        long long a = 0, b = n;
        while (a < b) {
            const long long mid = (a + b) >> 1;
            if ((long long)(packed[mid] >> 24) < off1) a = mid + 1; else b = mid;
        }
        d_split[0] = (unsigned long long)a;
    }
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const unsigned long long w = packed[k];
        long long idx = (long long)(w >> 24), cols = n_cols0;
        if (idx >= off1) { idx -= off1; cols = n_cols1; }
        const long long row = idx / cols;
        lo[k] = (int)row;
        hi[k] = (int)(idx - row * cols);
        cnt[k] = (unsigned int)(w & 0xffffffull);
    }
}

extern "C" int cmb_unpack_pairs(cmb_ctx* ctx, const uint64_t* d_packed, const uint64_t* d_n, int64_t cap,
                                int64_t n_cols0, int64_t off1, int64_t n_cols1, int32_t* d_lo, int32_t* d_hi,
                                uint32_t* d_cnt, uint64_t* d_split, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_packed || !d_n || !d_lo || !d_hi || !d_cnt || !d_split || cap < 1 || n_cols0 < 1 || n_cols1 < 1 ||
        off1 < n_cols0 * n_cols0 || n_cols0 > 0x7fffffffll || n_cols1 > 0x7fffffffll)
        return fail(ctx, CMB_ERR_ARG, "bad unpack arguments");
    CK(cudaSetDevice(ctx->device));
    const int grid = (int)std::min<long long>((cap + 255) / 256, (long long)ctx->sm_count * 8);
    k_unpack_pairs<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_packed),
                                                          reinterpret_cast<const unsigned long long*>(d_n), cap, n_cols0,
                                                          off1, n_cols1, d_lo, d_hi, d_cnt,
                                                          reinterpret_cast<unsigned long long*>(d_split));
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_group_contacts(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames, int n_atoms,
                                  const int32_t* d_atoms, int n_sel, const int32_t* d_pairs,
                                  const int64_t* d_group_ptr, int n_groups, int orthogonal, float cutoff,
                                  uint8_t* d_out, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_xyz || !d_atoms || !d_pairs || !d_group_ptr || !d_out || n_frames < 0 || n_atoms <= 0 || n_sel <= 0 ||
        n_groups < 0)
        return fail(ctx, CMB_ERR_ARG, "bad group-contact arguments");
    if (n_frames == 0 || n_groups == 0) return CMB_OK;
    if (n_frames > 0x7fffffffll) return fail(ctx, CMB_ERR_LIMIT, "too many frames for one call");
    CK(cudaSetDevice(ctx->device));
    const size_t smem = (size_t)n_sel * 12;
    if (smem > (size_t)ctx->max_smem_optin)
        return fail(ctx, CMB_ERR_LIMIT, "the contacts touch %d atoms; at most %d fit in shared memory", n_sel,
                    ctx->max_smem_optin / 12);
    CK(cudaFuncSetAttribute(k_group_contacts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_group_contacts<<<(unsigned)n_frames, CC_THREADS, smem, (cudaStream_t)stream>>>(
        d_xyz, d_box, n_atoms, d_atoms, n_sel, reinterpret_cast<const int2*>(d_pairs),
        reinterpret_cast<const long long*>(d_group_ptr), n_groups, orthogonal, cutoff, d_out);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_gather_atoms(cmb_ctx* ctx, const float* d_xyz_full, int64_t n_frames, int n_total,
                                const int32_t* d_used, int n_used, float* d_out, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_xyz_full || !d_used || !d_out || n_total <= 0 || n_used <= 0) return fail(ctx, CMB_ERR_ARG, "bad gather arguments");
    if (n_frames <= 0) return CMB_OK;
    CK(cudaSetDevice(ctx->device));
    long long total = n_frames * (long long)n_used;
    long long blocks = std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
    k_gather_atoms<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_xyz_full, n_frames, n_total, d_used,
                                                                        n_used, d_out);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

extern "C" int cmb_synth_frames(cmb_ctx* ctx, const float* d_base, const int32_t* d_atom_group, int n_atoms,
                                const float* d_box9, uint64_t seed, float amplitude, int image_shifts,
                                int64_t frame0, int64_t n_frames, float* d_out, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_base || !d_out || n_atoms <= 0) return fail(ctx, CMB_ERR_ARG, "bad synth arguments");
    if (image_shifts && (!d_atom_group || !d_box9)) return fail(ctx, CMB_ERR_ARG, "image shifts need groups and a box");
    if (n_frames <= 0) return CMB_OK;
    CK(cudaSetDevice(ctx->device));
    long long total = n_frames * (long long)n_atoms;
    long long blocks = std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 16);
    k_synth_frames<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_base, d_atom_group, n_atoms, d_box9, (unsigned long long)seed, amplitude, image_shifts, frame0,
        n_frames, d_out);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

// ---------------------------------------------------------------------------------
// min-distance analyses
// ---------------------------------------------------------------------------------
// brute force over the whole product for frames [f0, f0 + nf) (at most 32768 frames)
static int min_dist_brute(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t f0, int64_t nf, int n_atoms,
                          const int32_t* d_query, int n_q, const int32_t* d_haystack, int n_h, int orthogonal,
                          float* d_min, int64_t* d_arg, cudaStream_t st)
{
    const int n_chunks = (n_h + MD_HCHUNK - 1) / MD_HCHUNK;
    const size_t need = (size_t)nf * n_chunks * (sizeof(float) + sizeof(long long)) + 256;
    int rc = ensure_bytes(ctx, &ctx->d_tmp, &ctx->tmp_bytes, need);
    if (rc) return rc;
    long long* part_idx = (long long*)ctx->d_tmp;
    float* part_d = (float*)((char*)ctx->d_tmp + (size_t)nf * n_chunks * sizeof(long long));
    dim3 grid((unsigned)n_chunks, (unsigned)nf);
    k_min_dist_partial<<<grid, MD_THREADS, 0, st>>>(
        d_xyz + (size_t)f0 * n_atoms * 3, d_box ? d_box + (size_t)f0 * 9 : nullptr, n_atoms, d_query, n_q,
        d_haystack, n_h, orthogonal, part_d, part_idx);
    CK(cudaGetLastError());
    k_min_dist_final<<<(unsigned)((nf + 127) / 128), 128, 0, st>>>(part_d, part_idx, n_chunks, nf, d_min + f0,
                                                                   (long long*)d_arg + f0);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 2;
    ctx->stats["brute_frames"] += nf;
    return CMB_OK;
}

extern "C" int cmb_min_dist(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                            int n_atoms, const int32_t* d_query, int n_q, const int32_t* d_haystack, int n_h,
                            int orthogonal, float* d_min, int64_t* d_arg, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_xyz || !d_query || !d_haystack || !d_min || !d_arg || n_q <= 0 || n_h <= 0 || n_atoms <= 0)
        return fail(ctx, CMB_ERR_ARG, "bad min_dist arguments");
    if (n_frames <= 0) return CMB_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // Cell-pruned search (cmb_prune.cuh) when the product is large enough to pay for the binning:
    // the haystack atoms of a batch of frames are binned, every query atom searches its 27-cell
    // stencil, and the frames whose stencil minimum is not inside the guaranteed radius are re-done
    // by the brute-force kernel (synchronises once per batch to learn which).
    const bool pruned = ctx->prune && (long long)n_q * (long long)n_h >= (1ll << 18) && n_h >= 512;
    if (!pruned) {
        for (int64_t f0 = 0; f0 < n_frames; f0 += 32768) {
            int rc = min_dist_brute(ctx, d_xyz, d_box, f0, std::min<int64_t>(32768, n_frames - f0), n_atoms, d_query,
                                    n_q, d_haystack, n_h, orthogonal, d_min, d_arg, st);
            if (rc != CMB_OK) return rc;
        }
        return CMB_OK;
    }
    const int64_t batch = pg_batch_frames(n_h);
    const int qblocks = (n_q + PG_THREADS - 1) / PG_THREADS;
    PgScratch& S = ctx->pg;
    for (int64_t f0 = 0; f0 < n_frames; f0 += batch) {
        const int64_t nf = std::min<int64_t>(batch, n_frames - f0);
        std::string err;
        int launches = 0;
        int rc = pg_ensure(S, nf, n_h, err);
        if (rc != 0) return fail(ctx, CMB_ERR_CUDA, "%s", err.c_str());
        const float* x = d_xyz + (size_t)f0 * n_atoms * 3;
        const float* b = d_box ? d_box + (size_t)f0 * 9 : nullptr;
        rc = pg_bin_frames(S, x, b, nf, n_atoms, d_haystack, n_h, d_query, n_q, 0.f, st, err, launches);
        if (rc != 0) return fail(ctx, CMB_ERR_CUDA, "%s", err.c_str());
        const size_t need = (size_t)nf * qblocks * (sizeof(float) + sizeof(long long)) + 256;
        rc = ensure_bytes(ctx, &ctx->d_tmp, &ctx->tmp_bytes, need);
        if (rc) return rc;
        long long* part_idx = (long long*)ctx->d_tmp;
        float* part_d = (float*)((char*)ctx->d_tmp + (size_t)nf * qblocks * sizeof(long long));
        k_min_dist_cells<<<dim3((unsigned)qblocks, (unsigned)nf), PG_THREADS, 0, st>>>(
            S.frames, x, b, n_atoms, d_query, n_q, n_h, S.cap_sel, orthogonal, S.cellptr, S.sorted, part_d, part_idx);
        k_min_dist_final<<<(unsigned)((nf + 127) / 128), 128, 0, st>>>(part_d, part_idx, qblocks, nf, d_min + f0,
                                                                       (long long*)d_arg + f0);
        CK(cudaMemsetAsync(S.flags + S.cap_frames, 0, sizeof(int), st));
        pg_flag<<<(unsigned)((nf + 127) / 128), 128, 0, st>>>(S.frames, d_min + f0, nf, S.flags, S.flags + S.cap_frames);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(S.h_flags + S.cap_frames, S.flags + S.cap_frames, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats["kernel_launches"] += launches + 3;
        const int n_flagged = S.h_flags[S.cap_frames];
        ctx->stats["pruned_frames"] += nf - n_flagged;
        if (n_flagged == 0) continue;
        CK(cudaMemcpy(S.h_flags, S.flags, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost));
        for (int64_t k = 0; k < nf;) {
            if (!S.h_flags[k]) { k++; continue; }
            int64_t e = k + 1;
            while (e < nf && S.h_flags[e]) e++;
            rc = min_dist_brute(ctx, d_xyz, d_box, f0 + k, e - k, n_atoms, d_query, n_q, d_haystack, n_h, orthogonal,
                                d_min, d_arg, st);
            if (rc != CMB_OK) return rc;
            k = e;
        }
    }
    return CMB_OK;
}

extern "C" int cmb_min_dist_pairs(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                                  int n_atoms, const int32_t* d_pairs, int64_t n_pairs, int orthogonal,
                                  float* d_min, int64_t* d_arg, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_xyz || !d_pairs || !d_min || !d_arg || n_pairs <= 0 || n_atoms <= 0)
        return fail(ctx, CMB_ERR_ARG, "bad min_dist_pairs arguments");
    if (n_frames <= 0) return CMB_OK;
    CK(cudaSetDevice(ctx->device));
    const long long n_chunks_ll = (n_pairs + MD_PCHUNK - 1) / MD_PCHUNK;
    if (n_chunks_ll > 0x7fffffffll) return fail(ctx, CMB_ERR_LIMIT, "too many atom pairs");
    const int n_chunks = (int)n_chunks_ll;
    const int64_t fmax = 32768;   // gridDim.y limit is 65535
    for (int64_t f0 = 0; f0 < n_frames; f0 += fmax) {
        const int64_t nf = std::min<int64_t>(fmax, n_frames - f0);
        const size_t need = (size_t)nf * n_chunks * (sizeof(float) + sizeof(long long)) + 256;
        int rc = ensure_bytes(ctx, &ctx->d_tmp, &ctx->tmp_bytes, need);
        if (rc) return rc;
        long long* part_idx = (long long*)ctx->d_tmp;
        float* part_d = (float*)((char*)ctx->d_tmp + (size_t)nf * n_chunks * sizeof(long long));
        dim3 grid((unsigned)n_chunks, (unsigned)nf);
        k_min_dist_pairs_partial<<<grid, MD_THREADS, 0, (cudaStream_t)stream>>>(
            d_xyz + (size_t)f0 * n_atoms * 3, d_box ? d_box + (size_t)f0 * 9 : nullptr, n_atoms,
            reinterpret_cast<const int2*>(d_pairs), n_pairs, orthogonal, part_d, part_idx);
        CK(cudaGetLastError());
        k_min_dist_final<<<(unsigned)((nf + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            part_d, part_idx, n_chunks, nf, d_min + f0, (long long*)d_arg + f0);
        CK(cudaGetLastError());
        ctx->stats["kernel_launches"] += 2;
    }
    return CMB_OK;
}

extern "C" int cmb_nearest(cmb_ctx* ctx, const float* d_xyz_frame, const float* d_box9, int n_atoms,
                           double cutoff, int orthogonal, int excl_mode, const int32_t* d_atom_res,
                           const int64_t* d_excl_ptr, const int32_t* d_excl_idx, int32_t* d_nearest,
                           float* d_dist, void* stream)
{
    if (!ctx) return CMB_ERR_ARG;
    if (!d_xyz_frame || !d_nearest || !d_dist || n_atoms <= 0) return fail(ctx, CMB_ERR_ARG, "bad nearest arguments");
    if (excl_mode == 1 && !d_atom_res) return fail(ctx, CMB_ERR_ARG, "excl_mode 1 needs atom_res");
    if (excl_mode == 2 && (!d_excl_ptr || !d_excl_idx)) return fail(ctx, CMB_ERR_ARG, "excl_mode 2 needs CSR lists");
    if (excl_mode < 0 || excl_mode > 2) return fail(ctx, CMB_ERR_ARG, "bad excl_mode");
    CK(cudaSetDevice(ctx->device));
    const float cf = (float)cutoff;
    const float c2 = cf * cf;
    if (ctx->prune && n_atoms >= 4096 && cf > 0.f) {
        // cell-pruned search (cmb_prune.cuh): cells of the contact kernels' width, 27-cell stencil;
        // grids with fewer than 3 cells along a periodic dimension keep the brute-force kernel
        PgScratch& S = ctx->pg;
        std::string err;
        int launches = 0;
        cudaStream_t st = (cudaStream_t)stream;
        int rc = pg_ensure(S, 1, n_atoms, err);
        if (rc != 0) return fail(ctx, CMB_ERR_CUDA, "%s", err.c_str());
        rc = pg_bin_frames(S, d_xyz_frame, d_box9, 1, n_atoms, nullptr, n_atoms, nullptr, 0, cf, st, err, launches);
        if (rc != 0) return fail(ctx, CMB_ERR_CUDA, "%s", err.c_str());
        CK(cudaMemcpyAsync(S.h_flags, &S.frames[0].usable, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->stats["kernel_launches"] += launches;
        if (S.h_flags[0]) {
            k_nearest_cells<<<(unsigned)((n_atoms + PG_THREADS / 32 - 1) / (PG_THREADS / 32)), PG_THREADS, 0, st>>>(
                S.frames, d_xyz_frame, d_box9, n_atoms, c2, orthogonal, excl_mode, d_atom_res,
                (const long long*)d_excl_ptr, d_excl_idx, S.cellptr, S.sorted, d_nearest, d_dist);
            CK(cudaGetLastError());
            ctx->stats["kernel_launches"] += 1;
            ctx->stats["pruned_frames"] += 1;
            return CMB_OK;
        }
    }
    ctx->stats["brute_frames"] += 1;
    k_nearest<<<(unsigned)((n_atoms + NA_THREADS - 1) / NA_THREADS), NA_THREADS, 0, (cudaStream_t)stream>>>(
        d_xyz_frame, d_box9, n_atoms, c2, orthogonal, excl_mode, d_atom_res, (const long long*)d_excl_ptr,
        d_excl_idx, d_nearest, d_dist);
    CK(cudaGetLastError());
    ctx->stats["kernel_launches"] += 1;
    return CMB_OK;
}

// ---------------------------------------------------------------------------------
// device memory and the multi-GPU reduce for hosts that are not Python
// (the Python layer uses torch tensors / torch.distributed for the same things)
// ---------------------------------------------------------------------------------
extern "C" int cmb_device_alloc(cmb_ctx* ctx, uint64_t n_bytes, void** d_ptr)
{
    if (!ctx || !d_ptr) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *d_ptr = nullptr;
    CK(cudaMalloc(d_ptr, n_bytes ? (size_t)n_bytes : 16));
    CK(cudaMemset(*d_ptr, 0, (size_t)n_bytes));
    return CMB_OK;
}

extern "C" int cmb_device_free(cmb_ctx* ctx, void* d_ptr)
{
    if (!ctx) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(d_ptr));
    return CMB_OK;
}

extern "C" int cmb_device_zero(cmb_ctx* ctx, void* d_ptr, uint64_t n_bytes, void* stream)
{
    if (!ctx || !d_ptr) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(d_ptr, 0, (size_t)n_bytes, (cudaStream_t)stream));
    return CMB_OK;
}

extern "C" int cmb_copy_to_host(cmb_ctx* ctx, void* h_dst, const void* d_src, uint64_t n_bytes, void* stream)
{
    if (!ctx || (n_bytes && (!h_dst || !d_src))) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(h_dst, d_src, (size_t)n_bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return CMB_OK;
}

extern "C" int cmb_copy_to_device(cmb_ctx* ctx, void* d_dst, const void* h_src, uint64_t n_bytes, void* stream)
{
    if (!ctx || (n_bytes && (!d_dst || !h_src))) return CMB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(d_dst, h_src, (size_t)n_bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return CMB_OK;
}

// NCCL is bound at run time (dlopen): the library loads and every single-GPU entry point works on
// machines without NCCL; the prototypes below are NCCL's public, stable C API (nccl.h).
namespace {
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, cmb_nccl_id, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load(cmb_ctx* ctx)
{
    if (g_nccl.handle) return CMB_OK;
    const char* names[] = {getenv("CMB200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* name : names) {
        if (!name || !*name) continue;
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(ctx, CMB_ERR_STATE, "NCCL not found (libnccl.so.2; set CMB200_NCCL to its path): %s", dlerror());
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, cmb_nccl_id, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    a.Reduce = (int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t))dlsym(h, "ncclReduce");
    a.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.Reduce || !a.GetErrorString)
        return fail(ctx, CMB_ERR_STATE, "the NCCL library lacks an expected symbol");
    g_nccl = a;
    return CMB_OK;
}
#define NCCL_CK(call)                                                                              \
    do {                                                                                           \
        const int r_ = (call);                                                                     \
        if (r_ != 0) return fail(ctx, CMB_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)
}  // namespace

extern "C" int cmb_nccl_unique_id(cmb_ctx* ctx, cmb_nccl_id* out)
{
    if (!ctx || !out) return CMB_ERR_ARG;
    int rc = nccl_load(ctx);
    if (rc != CMB_OK) return rc;
    NCCL_CK(g_nccl.GetUniqueId(out));
    return CMB_OK;
}

extern "C" int cmb_nccl_comm_create(cmb_ctx* ctx, int n_ranks, const cmb_nccl_id* id, int rank, void** comm)
{
    if (!ctx || !id || !comm || n_ranks < 1 || rank < 0 || rank >= n_ranks) return CMB_ERR_ARG;
    int rc = nccl_load(ctx);
    if (rc != CMB_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    NCCL_CK(g_nccl.CommInitRank(comm, n_ranks, *id, rank));
    return CMB_OK;
}

extern "C" int cmb_nccl_comm_destroy(cmb_ctx* ctx, void* comm)
{
    if (!ctx || !comm) return CMB_ERR_ARG;
    int rc = nccl_load(ctx);
    if (rc != CMB_OK) return rc;
    NCCL_CK(g_nccl.CommDestroy(comm));
    return CMB_OK;
}

// reduce_all_results (frequency_task.py:125-141) across GPUs: the ranks' dense uint32 count tables
// -- identical layout on every rank -- are summed in place with ONE collective.  root < 0: every
// rank receives the sum (all-reduce); root >= 0: only that rank does (what the reference's
// client-side reduce delivers, at about half the traffic).  `comm` is an ncclComm_t (from
// cmb_nccl_comm_create, or any communicator of the same NCCL library the host already owns).
extern "C" int cmb_reduce(cmb_ctx* ctx, void* comm, uint32_t* d_table, int64_t n_elems, int root, void* stream)
{
    if (!ctx || !comm || !d_table || n_elems < 0) return CMB_ERR_ARG;
    int rc = nccl_load(ctx);
    if (rc != CMB_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const int nccl_uint32 = 3, nccl_sum = 0;
    if (root < 0)
        NCCL_CK(g_nccl.AllReduce(d_table, d_table, (size_t)n_elems, nccl_uint32, nccl_sum, comm, (cudaStream_t)stream));
    else
        NCCL_CK(g_nccl.Reduce(d_table, d_table, (size_t)n_elems, nccl_uint32, nccl_sum, root, comm, (cudaStream_t)stream));
    return CMB_OK;
}
