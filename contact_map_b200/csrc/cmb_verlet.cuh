// cmb_verlet.cuh -- frame-batch pair-list path ("Verlet batch"): the candidate pair list of a
// batch of frames is built ONCE, every frame of the batch then only walks that list.
//
// Replaces, per frame (reference file:line under /root/reference/contact_map/), the same lines as
// the cell kernels (cmb_frames_small.cuh / cmb_frames_large.cuh):
//   md.compute_neighborlist(used_traj, cutoff, frame)      contact_map.py:575
//   the Python set algebra of ContactObject._contact_map   contact_map.py:582-607
//   Counter.update / Counter += of ContactFrequency        contact_map.py:739-740
//
// Why: the cell kernels screen ~8 candidate pairs per contact and spend 2/3 of their issue slots
// on per-task set-up, ring traffic and per-hit atomics.  Frames of one trajectory
// are strongly correlated, so the classical MD answer applies: a skin-padded pair list.
//
//   builder (once per batch, all on the device, no host round trip):
//     1. REFERENCE positions: the mean of VL_SAMPLES frames spread over the batch (every sample
//        unwrapped to the periodic image next to the first one), wrapped into the primary cell;
//        skin = 1.3 x the largest displacement of any sampled atom from that mean (+ 0.002 nm).
//     2. cell grid with cells >= r_list = 1.002 c + 2 skin; every eligible pair (roles, residue
//        exclusion: contact_map.py:41-60,442-461 -- both are frame independent, so they are
//        applied HERE and never again) whose reference minimum-image distance is <= r_list is
//        recorded once, with the lattice shift of its image.
//     3. the records are sorted by (tile, residue pair) -- the atom pairs of one residue pair form
//        a SEGMENT with one "seen in this frame" flag and one counter -- and, stably, by (tile,
//        owner atom): the partners of one atom form an OWNER LIST.  Owner lists are sorted by
//        length and dealt 32 at a time to "bundles" (lane = owner); a bundle is stored as ROWS of
//        4 steps x 32 lanes.  A tile is a block of cells whose segments, atoms ("slots", periodic
//        images of an atom are separate slots that carry their lattice shift) and counters fit one
//        CTA's shared memory: small systems are one tile, the 100k-atom system a few hundred.
//        The slot chain (keys, sort, unique, reference positions, fixed-point grid) runs on a second
//        stream beside the segment / owner-list / bundle chain; vl_fill joins them and balances the
//        shared-memory banks of the gathers (order of a lane's four partners, deal of the owners
//        to the half warps).
//   frame kernel (persistent CTAs, work item = tile x range of frames):
//     a. every thread loads its slots' RAW coordinates (prefetched one frame ahead), moves them
//        to the periodic image next to the slot's reference position (one rounding: fp64, or an
//        fp32 fma for axis-aligned cells), checks the GUARD |x - ref| <= skin -- a frame that
//        violates it (or whose unit cell differs from the batch's, or whose coordinates exceed the
//        magnitude the error bound was derived for) is appended to the fall-back list and skipped
//        -- and stores the position as three 16-bit fixed-point numbers relative to the tile's
//        origin (8 bytes per slot);
//     b. warps walk rows; lane l of a row walks four partners of ITS owner atom: one LDS.64 for
//        the owner, one per partner, exact integer differences on the grid, 1 FMUL + 2 FFMA, two
//        compares against thresholds widened by the rigorous grid bound (vl_quant_setup);
//        d2 <= c2q_lo is a certain contact, c2q_lo < d2 < c2q_hi is AMBIGUOUS: parked and decided
//        one frame later by MDTraj's exact non-fused arithmetic on the raw coordinates
//        (cmb_math.cuh::nl_d2), exactly as in the cell kernels (same per-batch error bound,
//        frame_margins); contacts increment a per-entry byte counter in shared memory (the
//        thread owns the entry: no atomics) and set the byte flag of the entry's residue pair --
//        the per-frame set semantics of contact_map.py:601-607,740 without a bitmap;
//     c. after the last frame of the item the counters are flushed with one RED per non-zero
//        entry (a batch of 64 frames issues ~30x fewer atomics than one per hit).
//   fall back: the frames on the fall-back list are processed by the cell kernels (which accept
//   a device-side frame list), so the result never depends on the list being usable.
//
// Exactness.  A pair the reference arithmetic calls a contact in frame f has a true minimum-image
// distance < 1.002 c (DESIGN.md 4.0 "domain").  Both atoms are within `skin` of their reference
// positions (guard), so the same image of the pair is within 1.002 c + 2 skin = r_list at the
// reference positions and the pair is on the list (the builder tests with a relative slack of
// 1e-5 and bins in fp64).  For listed pairs the screening distance differs from the exact one by
// at most the frame_margins bound, evaluated for the largest coordinate magnitude the guard
// admits.  Unit cells need >= 3 list cells per periodic dimension (then the image of a pair
// inside r_list is unique) and the lower-triangular form MDTraj produces; otherwise the builder
// marks the list invalid and every frame falls back.
#pragma once
#include <cub/cub.cuh>

#include <chrono>
#include <cstring>
#include <string>

#include "cmb_cells.cuh"

namespace cmb {

constexpr int VL_THREADS = 1024;
constexpr int VL_SPT = 4;                                // slots per thread
constexpr int VL_MAX_SLOTS = VL_THREADS * VL_SPT - 1;    // + 1 "far" slot for padding entries
constexpr int VL_ROW = 128;                              // entries per row: 32 lanes x 4 steps
constexpr int VL_MAX_ROWS = 768;                         // per tile
constexpr int VL_MAX_ENTRIES = VL_MAX_ROWS * VL_ROW;     // 98 304 byte counters
constexpr int VL_MAX_SEGS = 24576;                       // residue pairs per tile
constexpr int VL_SEG_BITS = 15;                          // bits of a tile-relative segment id
constexpr int VL_SLOT_BYTES = 8;                         // a slot of the frame kernel: three 16-bit grid coordinates
constexpr int VL_MAX_TILES = 4096;
constexpr int VL_MAXC = 1 << 17;                         // cells of the list grid
constexpr int VL_SAMPLES = 16;
constexpr int VL_ITEM_FRAMES = 255;                      // byte counters: flush at least this often
constexpr int VL_MIN_ITEM_FRAMES = 12;                   // shorter items would not amortise set-up and flush
constexpr int VL_SINGLE_TILE_ATOMS = 3800;
constexpr int VL_LEN_BITS = 20;                          // list length field of the bundle sort key
constexpr unsigned long long VL_PAD = ~0ull;

struct VlHeader {
    int valid;                 // 0: the list cannot be used, every frame falls back
    int n_tiles, tdim, ntx, nty, ntz;
    int n_entries;             // list pairs found (may exceed the capacity -> invalid)
    int n_segs, n_lists, n_slots, n_bundles, n_rows;
    int rbits, abits;          // bits of a compact residue id / of a used-atom index in the sort keys
    unsigned d_max_bits, a_max_bits;       // sample maxima (non-negative float bits)
    unsigned bb_lo[3], bb_hi[3];           // bounding box of the reference positions (ordered uints)
    float skin2, rlist, a_assumed, c2_lo, c2_hi;
    // the frame kernel keeps the positions of a tile as 16-bit fixed-point numbers relative to the
    // tile's origin (8-byte shared-memory gathers): grid units per nm, thresholds in grid units^2
    float skin, qscale, c2q_lo, c2q_hi;
    int periodic;
    unsigned box_bits[9];      // the batch's unit cell (compared bit for bit per frame)
    double hinv[9], h[9];
    float4 shift[27];
    GridS grid;
    int fail_code;             // why valid == 0 (diagnostics)
};

struct VlTile {
    int slot0, n_slots;        // into slot_word / slot_ref
    int seg0, n_segs;          // into seg_pair: the residue pairs of the tile
    int list0, n_lists;        // owner lists of the tile (length-sorted order; builder only)
    int row0, n_rows;          // into entries (rows of VL_ROW words) and row_owner
    float ox, oy, oz;          // origin of the tile's fixed-point grid (below every slot position the guard admits)
};

// everything the builder writes and the frame kernel reads (device pointers)
struct VlList {
    VlHeader* hdr;
    VlTile* tiles;             // [VL_MAX_TILES]
    unsigned* slot_word;       // [slot_cap] atom | shift id << 24
    float4* slot_ref;          // [slot_cap] reference position of the slot's atom (primary cell)
    unsigned* tile_bb;         // [VL_MAX_TILES][6] bounding box of the tile's slot reference positions (ordered uints)
    unsigned short* row_owner; // [row_cap * 32] slot (byte offset into the position array) of the atom that owns
                               // lane l of the row: the lane walks (a part of) that atom's partner list
    unsigned* entries;         // [row_cap * VL_ROW] partner slot byte offset (bits 3-15) | direction (bit 0) |
                               // tile-relative segment (residue pair) << 16
    uint2* seg_pair;           // [entry_cap] {compact residue lo, hi} of every segment
    const float4* ref_w;       // [n] wrapped reference position of every atom (guard pass)
    long long entry_cap, slot_cap, row_cap;
};

// State of ONE walk over a list (a list outlives the batch it was built from: the chunks of a
// host trajectory all walk the list built from the first one, on two streams).
struct VlRun {
    int* state = nullptr;      // device [2]: frames on the fall-back list, work queue of the frame kernel
    int* tripped = nullptr;    // [frame_cap] frames of the current call that fall back
    int* frame_flag = nullptr; // [frame_cap] 1: the frame is on the fall-back list (multi-tile lists: set by vl_guard)
    long long frame_cap = 0;
    int* h_info = nullptr;     // pinned [4]: read-backs of the large path (tripped count, list validity, fail code)
};

// builder scratch
struct VlScratch {
    int cap_atoms = 0;
    long long entry_cap = 0;
    VlList L{};
    float4* ref_raw = nullptr;      // [n] mean of the unwrapped samples
    float4* ref_w = nullptr;        // [n] wrapped reference {x, y, z, cell}
    int* atomrank = nullptr;        // [n]
    int* cellcnt = nullptr;         // [VL_MAXC + 1]
    int* cellptr = nullptr;         // [VL_MAXC + 1]
    float4* sref = nullptr;         // [n] cell-sorted reference {x, y, z, atom}
    int2* smeta = nullptr;          // [n] cell-sorted {ekey, res_c | role << 30}
    int* scell = nullptr;           // [n] cell of each sorted position
    int* tile_of_res = nullptr;     // [n_res_c]
    unsigned long long* rec_key[2] = {nullptr, nullptr};   // [entry_cap]
    unsigned long long* rec_val[2] = {nullptr, nullptr};
    unsigned long long* seg_key = nullptr;                 // [entry_cap] unique keys
    int* seg_len = nullptr;                                // [entry_cap]
    int* seg_start = nullptr;                              // [entry_cap + 1]
    int* n_runs = nullptr;                                 // [3] segments, unique slots, owner lists
    unsigned long long* own_key = nullptr;                 // [entry_cap] records re-sorted by (tile, owner atom)
    unsigned long long* own_val = nullptr;
    unsigned long long* list_key = nullptr;                // [entry_cap] unique (tile, owner atom)
    int* list_len = nullptr;                               // [entry_cap]
    int* list_start = nullptr;                             // [entry_cap + 1]
    unsigned long long* skey[2] = {nullptr, nullptr};      // [entry_cap] tile << VL_LEN_BITS | ~len
    int* sval[2] = {nullptr, nullptr};                     // [entry_cap] owner list id
    unsigned long long* slot_key[2] = {nullptr, nullptr};  // [2 * entry_cap + n]
    unsigned long long* slot_uniq = nullptr;               // [2 * entry_cap + n]
    int* tile_seg0 = nullptr;       // [VL_MAX_TILES + 1] first segment of each tile
    int* tile_list0 = nullptr;      // [VL_MAX_TILES + 1] first owner list (length-sorted order) of each tile
    int* tile_bundle0 = nullptr;    // [VL_MAX_TILES + 1]
    int* tile_slot0 = nullptr;      // [VL_MAX_TILES + 1]
    int* bundle_row0 = nullptr;     // [bundle_cap + 1]
    long long bundle_cap = 0;
    void* cub_tmp = nullptr;
    void* cub_tmp2 = nullptr;       // temp storage of the slot chain, which runs beside the owner-list chain
    size_t cub_bytes = 0;
    cudaStream_t aux = nullptr;     // the builder's second stream (vl_build: fork after sort 1, join before vl_fill)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // the builder as ONE graph launch (vl_build): per-call arguments in device memory, the launch
    // sequence captured once per set of fixed arguments on the builder's own stream
    void* d_args = nullptr;         // VlBuildArgs
    cudaStream_t bs = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    cudaGraphExec_t gexec = nullptr;
    unsigned char gkey[64] = {0};   // VlGraphKey of gexec
    int graph_state = 0;            // 0 not captured yet, 1 usable, -1 capture failed once: plain launches from then on
    long long captures = 0, capture_us = 0;   // diagnostics: graphs captured, host time spent capturing + instantiating
};

inline void vl_free(VlScratch& S)
{
    cudaFree(S.L.hdr); cudaFree(S.L.tiles); cudaFree(S.L.slot_word); cudaFree(S.L.slot_ref);
    cudaFree(S.L.row_owner); cudaFree(S.L.entries); cudaFree(S.L.seg_pair); cudaFree(S.L.tile_bb);
    cudaFree(S.own_key); cudaFree(S.own_val); cudaFree(S.list_key); cudaFree(S.list_len); cudaFree(S.list_start);
    cudaFree(S.tile_list0);
    cudaFree(S.ref_raw); cudaFree(S.ref_w); cudaFree(S.atomrank); cudaFree(S.cellcnt); cudaFree(S.cellptr);
    cudaFree(S.sref); cudaFree(S.smeta); cudaFree(S.scell); cudaFree(S.tile_of_res);
    for (int k = 0; k < 2; k++) {
        cudaFree(S.rec_key[k]); cudaFree(S.rec_val[k]); cudaFree(S.skey[k]); cudaFree(S.sval[k]);
        cudaFree(S.slot_key[k]);
    }
    cudaFree(S.seg_key); cudaFree(S.seg_len); cudaFree(S.seg_start); cudaFree(S.n_runs); cudaFree(S.slot_uniq);
    cudaFree(S.tile_seg0); cudaFree(S.tile_bundle0); cudaFree(S.tile_slot0); cudaFree(S.bundle_row0);
    cudaFree(S.cub_tmp); cudaFree(S.cub_tmp2); cudaFree(S.d_args);
    if (S.gexec) cudaGraphExecDestroy(S.gexec);
    if (S.bs) cudaStreamDestroy(S.bs);
    if (S.ev_in) cudaEventDestroy(S.ev_in);
    if (S.ev_out) cudaEventDestroy(S.ev_out);
    if (S.aux) cudaStreamDestroy(S.aux);
    if (S.ev_fork) cudaEventDestroy(S.ev_fork);
    if (S.ev_join) cudaEventDestroy(S.ev_join);
    S = VlScratch();
}

inline void vl_run_free(VlRun& R)
{
    cudaFree(R.state); cudaFree(R.tripped); cudaFree(R.frame_flag);
    if (R.h_info) cudaFreeHost(R.h_info);
    R = VlRun();
}

// ----- small helpers --------------------------------------------------------------------
__device__ __forceinline__ unsigned vl_ordered(float v)      // monotone float -> uint map
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float vl_unordered(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// position of `p` moved to the periodic image next to `r` (fp64, rounded once to fp32)
__device__ __forceinline__ void vl_unwrap(const double* hinv, const double* h, int periodic, float px, float py,
                                          float pz, float rx, float ry, float rz, float& ox, float& oy, float& oz)
{
    if (!periodic) { ox = px; oy = py; oz = pz; return; }
    const double dx = (double)px - (double)rx, dy = (double)py - (double)ry, dz = (double)pz - (double)rz;
    const double n0 = rint(dx * hinv[0] + dy * hinv[3] + dz * hinv[6]);
    const double n1 = rint(dx * hinv[1] + dy * hinv[4] + dz * hinv[7]);
    const double n2 = rint(dx * hinv[2] + dy * hinv[5] + dz * hinv[8]);
    ox = (float)((double)px - (n0 * h[0] + n1 * h[3] + n2 * h[6]));
    oy = (float)((double)py - (n0 * h[1] + n1 * h[4] + n2 * h[7]));
    oz = (float)((double)pz - (n0 * h[2] + n1 * h[5] + n2 * h[8]));
}

// The same for a cell whose vectors lie along the axes, in fp32: n = rint((p - r) / L) needs no more
// than that -- whenever the result passes the guard, (p - r) / L lies within skin / L << 1/2 of the
// integer both precisions round it to -- and fma(-n, L, p) is the exact p - n L rounded once, which is
// what the fp64 version computes (L is a float, n L and the difference are exact in fp64).  A frame
// for which the two could differ fails the guard in both and is re-done by the fall back.
__device__ __forceinline__ void vl_unwrap_ortho(const float* hf, float px, float py, float pz, float rx, float ry,
                                                float rz, float& ox, float& oy, float& oz)
{
    const float n0 = rintf(__fmul_rn(__fsub_rn(px, rx), hf[0]));
    const float n1 = rintf(__fmul_rn(__fsub_rn(py, ry), hf[1]));
    const float n2 = rintf(__fmul_rn(__fsub_rn(pz, rz), hf[2]));
    ox = __fmaf_rn(-n0, hf[3], px);
    oy = __fmaf_rn(-n1, hf[4], py);
    oz = __fmaf_rn(-n2, hf[5], pz);
}

__device__ __forceinline__ void vl_box_inverse(const float* box9, double* h, double* m)
{
    for (int k = 0; k < 9; k++) h[k] = (double)box9[k];
    const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) +
                       h[2] * (h[3] * h[7] - h[4] * h[6]);
    const double id = 1.0 / det;
    m[0] = (h[4] * h[8] - h[5] * h[7]) * id; m[1] = (h[2] * h[7] - h[1] * h[8]) * id;
    m[2] = (h[1] * h[5] - h[2] * h[4]) * id; m[3] = (h[5] * h[6] - h[3] * h[8]) * id;
    m[4] = (h[0] * h[8] - h[2] * h[6]) * id; m[5] = (h[2] * h[3] - h[0] * h[5]) * id;
    m[6] = (h[3] * h[7] - h[4] * h[6]) * id; m[7] = (h[1] * h[6] - h[0] * h[7]) * id;
    m[8] = (h[0] * h[4] - h[1] * h[3]) * id;
}

// shared-memory atomics as single instructions (the compiler's warp-aggregated forms cost a dozen
// issue slots when one lane is active)
__device__ __forceinline__ int vl_atoms_add(int* addr, int v)
{
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                 : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void vl_reds_or(unsigned* addr, unsigned v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(addr)), "r"(v) : "memory");
}

// ----- builder kernels --------------------------------------------------------------------
// What changes from one build to the next.  It lives in device memory (copied ahead of the build) so
// that every launch of the builder has the SAME arguments and the whole builder can be replayed as
// one CUDA graph (vl_build).
struct VlBuildArgs {
    const float* xyz;          // [n_frames][n][3]
    const float* box;          // [n_frames][9] or nullptr
    long long n_frames;
};

__global__ void vl_init(VlHeader* hdr, const VlBuildArgs* __restrict__ args, int rbits, int abits)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float* box = args->box;
    const int periodic = box != nullptr;
    VlHeader& H = *hdr;
    H.valid = 1; H.fail_code = 0;
    H.n_tiles = H.n_entries = H.n_segs = H.n_lists = H.n_slots = H.n_bundles = H.n_rows = 0;
    H.rbits = rbits; H.abits = abits;
    H.d_max_bits = 0u; H.a_max_bits = 0u;
    for (int k = 0; k < 3; k++) { H.bb_lo[k] = 0xffffffffu; H.bb_hi[k] = 0u; }
    H.periodic = periodic;
    for (int k = 0; k < 9; k++) { H.hinv[k] = 0.0; H.h[k] = 0.0; H.box_bits[k] = 0u; }
    if (periodic) {
        for (int k = 0; k < 9; k++) H.box_bits[k] = __float_as_uint(box[k]);
        vl_box_inverse(box, H.h, H.hinv);
    }
}

// Reference positions: mean over the sample frames (each unwrapped next to the first sample),
// largest displacement of a sample from the mean, largest |raw coordinate|, bounding box.
// 16 lanes per atom (lane = sample), 8 atoms per block.
__global__ void __launch_bounds__(128) vl_sample(VlHeader* hdr, const VlBuildArgs* __restrict__ args, int n,
                                                 float4* __restrict__ ref_raw)
{
    __shared__ double hs[18];
    __shared__ int box_ok;
    const float* __restrict__ xyz = args->xyz;
    const float* __restrict__ box = args->box;
    const long long n_frames = args->n_frames;
    const int n_samp = (int)(n_frames < VL_SAMPLES ? n_frames : VL_SAMPLES);
    if (threadIdx.x < 18) hs[threadIdx.x] = threadIdx.x < 9 ? hdr->hinv[threadIdx.x] : hdr->h[threadIdx.x - 9];
    if (threadIdx.x == 0) {
        // the sampled frames must share the batch's unit cell (all frames are checked again by the
        // frame kernel / the guard pass)
        int ok = 1;
        if (box != nullptr)
            for (int k = 0; k < n_samp; k++) {
                const long long f = n_samp > 1 ? (long long)k * (n_frames - 1) / (n_samp - 1) : 0;
                for (int q = 0; q < 9; q++) ok = ok && (__float_as_uint(box[f * 9 + q]) == hdr->box_bits[q]);
            }
        box_ok = ok;
    }
    __syncthreads();
    if (!box_ok) {
        if (threadIdx.x == 0 && blockIdx.x == 0) { hdr->valid = 0; hdr->fail_code = 1; }
        return;
    }
    const int k = threadIdx.x & 15;                       // sample of this lane
    const int i = blockIdx.x * 8 + (threadIdx.x >> 4);    // atom of this half warp
    const int periodic = box != nullptr;
    const bool live = i < n && k < n_samp;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) {
        const long long f = n_samp > 1 ? (long long)k * (n_frames - 1) / (n_samp - 1) : 0;
        const float* p = xyz + ((size_t)f * (size_t)n + (size_t)i) * 3;
        px = __ldg(p); py = __ldg(p + 1); pz = __ldg(p + 2);
    }
    // anchor: the first sample of the atom
    const float ax = __shfl_sync(0xffffffffu, px, 0, 16), ay = __shfl_sync(0xffffffffu, py, 0, 16),
                az = __shfl_sync(0xffffffffu, pz, 0, 16);
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (live) vl_unwrap(hs, hs + 9, periodic, px, py, pz, ax, ay, az, ux, uy, uz);
    float sx = ux, sy = uy, sz = uz;
#pragma unroll
    for (int d = 8; d >= 1; d >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, d, 16);
        sy += __shfl_xor_sync(0xffffffffu, sy, d, 16);
        sz += __shfl_xor_sync(0xffffffffu, sz, d, 16);
    }
    const float inv = 1.0f / (float)n_samp;
    const float mx = sx * inv, my = sy * inv, mz = sz * inv;
    float dmax = 0.f, amax = 0.f;
    if (live) {
        const float ex = ux - mx, ey = uy - my, ez = uz - mz;
        dmax = sqrtf(ex * ex + ey * ey + ez * ez);
        amax = fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz)));
        // NaN / Inf coordinates: the bits of +Inf exceed every finite value, vl_setup rejects them
        if (!(dmax >= 0.f)) dmax = __int_as_float(0x7f800000);
        if (!(amax >= 0.f)) amax = __int_as_float(0x7f800000);
    }
    if (i < n && k == 0) {
        ref_raw[i] = make_float4(mx, my, mz, 0.f);
        if (!periodic) {
            atomicMin(&hdr->bb_lo[0], vl_ordered(mx)); atomicMax(&hdr->bb_hi[0], vl_ordered(mx));
            atomicMin(&hdr->bb_lo[1], vl_ordered(my)); atomicMax(&hdr->bb_hi[1], vl_ordered(my));
            atomicMin(&hdr->bb_lo[2], vl_ordered(mz)); atomicMax(&hdr->bb_hi[2], vl_ordered(mz));
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, d));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&hdr->d_max_bits, __float_as_uint(dmax));
        atomicMax(&hdr->a_max_bits, __float_as_uint(amax));
    }
}

// skin, list radius, error bound, list grid, tiling (one thread)
__global__ void vl_setup(VlHeader* hdr, const VlBuildArgs* __restrict__ args, int n, float cutoff, float c2,
                         float max_rlist_factor, int conservative_tiles)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float* box = args->box;
    VlHeader& H = *hdr;
    if (!H.valid) return;
    const float dmax = __uint_as_float(H.d_max_bits), amax = __uint_as_float(H.a_max_bits);
    if (!(dmax < 1.0e6f) || !(amax < 1.0e6f)) { H.valid = 0; H.fail_code = 2; return; }
    const float skin = 1.3f * dmax + 0.002f;
    const float rlist = cutoff * 1.002f + 2.0f * skin + 1.0e-6f;
    if (rlist > max_rlist_factor * cutoff) { H.valid = 0; H.fail_code = 3; return; }   // the list would not pay
    // the guard compares in fp32 against a slightly smaller radius than the list was built for
    H.skin2 = skin * skin * 0.9999f;
    H.skin = skin;
    H.rlist = rlist;
    H.a_assumed = 1.25f * amax + 0.5f;
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
    if (!H.periodic)
        for (int k = 0; k < 3; k++) { lo[k] = vl_unordered(H.bb_lo[k]); hi[k] = vl_unordered(H.bb_hi[k]); }
    make_grid(H.grid, H.periodic ? box : nullptr, (double)rlist, VL_MAXC, lo, hi);
    if (H.periodic && !H.grid.distinct) { H.valid = 0; H.fail_code = 4; return; }     // < 3 list cells per dimension
    // screening thresholds for every frame the guard admits
    frame_margins(H.periodic ? box : nullptr, 1, H.a_assumed + skin, c2, cutoff, 0, H.c2_lo, H.c2_hi);
    if (!(H.c2_lo > 0.f)) { H.valid = 0; H.fail_code = 5; return; }                   // no usable error bound
    for (int id = 0; id < 27; id++) H.shift[id] = lattice_shift(H.periodic ? box : nullptr, id);
    // tiles: blocks of tdim^3 cells; one tile for systems that fit a CTA
    const GridS& G = H.grid;
    int t;
    if (n <= VL_SINGLE_TILE_ATOMS) {
        t = max(G.nx, max(G.ny, G.nz));
    } else {
        // slots of a tile: its own atoms plus the partners of its pairs in the shell of cells around
        // it.  A pair belongs to the tile of its lower residue, so about half the shell shows up
        // (measured on the 100k-atom system: average tile = estimate x 0.94, fullest tile < x 1.48).
        // The allowance is deliberately tight -- bigger tiles mean fewer halo slots (tiles of 3^3
        // instead of 2^3 cells on that system: 27 % fewer slots, frame kernel 21 % faster); should a
        // tile overflow (fail codes 9-13), the host rebuilds the list with the whole shell counted.
        const float apc = (float)n / (float)G.ncell;
        t = 6;
        for (; t > 1; t--) {
            const float core = (float)(t * t * t), halo = (float)((t + 2) * (t + 2) * (t + 2));
            const float est = (conservative_tiles ? halo * 1.15f : 0.5f * (core + halo) * 1.4f) * apc;
            // (rows: ~ 0.19 per slot of such a tile: 128 entries per row at ~ 70 % fill)
            if (est <= (float)VL_MAX_SLOTS && (conservative_tiles || est * 0.19f <= (float)VL_MAX_ROWS)) break;
        }
    }
    H.tdim = t;
    H.ntx = (G.nx + t - 1) / t; H.nty = (G.ny + t - 1) / t; H.ntz = (G.nz + t - 1) / t;
    const long long nt = (long long)H.ntx * H.nty * H.ntz;
    if (nt > VL_MAX_TILES) { H.valid = 0; H.fail_code = 6; return; }
    H.n_tiles = (int)nt;
}

__global__ void __launch_bounds__(256) vl_zero_cells(const VlHeader* hdr, int* cellcnt)
{
    if (!hdr->valid) return;
    const int ncell = hdr->grid.ncell;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= ncell; c += gridDim.x * blockDim.x) cellcnt[c] = 0;
}

__global__ void __launch_bounds__(256) vl_bin(const VlHeader* hdr, int n, const float4* __restrict__ ref_raw,
                                              float4* __restrict__ ref_w, int* __restrict__ atomrank,
                                              int* __restrict__ cellcnt)
{
    if (!hdr->valid) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridS& G = hdr->grid;
    const float4 r = ref_raw[i];
    float x = r.x, y = r.y, z = r.z;
    const int c = cell_wrap(G, x, y, z);
    if (!G.periodic) { x = r.x; y = r.y; z = r.z; }
    ref_w[i] = make_float4(x, y, z, __int_as_float(c));
    atomrank[i] = atomicAdd(&cellcnt[c], 1);
}

// exclusive scan over the cells (one CTA)
__global__ void __launch_bounds__(1024) vl_scan_cells(const VlHeader* hdr, const int* __restrict__ cellcnt,
                                                      int* __restrict__ cellptr, int n)
{
    if (!hdr->valid) return;
    __shared__ int wsum[32];
    __shared__ int carry;
    const int ncell = hdr->grid.ncell;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int c = base + threadIdx.x;
        const int v = c < ncell ? cellcnt[c] : 0;
        int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = wsum[lane];
            int iw = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, iw, d);
                if (lane >= d) iw += u;
            }
            wsum[lane] = iw - w;
        }
        __syncthreads();
        const int excl = carry + wsum[warp] + incl - v;
        if (c < ncell) cellptr[c] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cellptr[ncell] = n;
}

// Tile edge from the ACTUAL occupancy of the list cells (lists of several tiles; vl_setup's choice
// from the average density stands for systems that fill their cell evenly and is replaced here).
// For every candidate edge t, largest first, every tile's slots are estimated as its own atoms plus a
// share of the atoms in the one-cell shell around it -- a pair belongs to the tile of its lower
// residue, so about half of the shell shows up as partners (0.6 with allowance; the pessimistic
// rebuild after an overflow counts the whole shell) -- and the largest t whose FULLEST tile fits the
// frame kernel's shared memory is taken.  A protein in vacuum or a membrane slab no longer gets the
// tile size of its average density.
__global__ void __launch_bounds__(1024) vl_pick_tiles(VlHeader* hdr, const int* __restrict__ cellptr, int n,
                                                      int conservative_tiles, float optimistic_share)
{
    if (!hdr->valid || n <= VL_SINGLE_TILE_ATOMS) return;
    __shared__ int s_max, s_pick;
    const GridS& G = hdr->grid;
    const int nx = G.nx, ny = G.ny, nz = G.nz, periodic = G.periodic;
    const float share = conservative_tiles ? 1.0f : optimistic_share;
    if (threadIdx.x == 0) s_pick = 0;
    for (int t = 6; t >= 1; t--) {
        const int ntx = (nx + t - 1) / t, nty = (ny + t - 1) / t, ntz = (nz + t - 1) / t;
        const long long nt = (long long)ntx * nty * ntz;
        const bool feasible = nt <= VL_MAX_TILES;
        if (threadIdx.x == 0) s_max = 0;
        __syncthreads();
        if (feasible) {
            float worst = 0.f;
            for (int tile = threadIdx.x; tile < (int)nt; tile += 1024) {
                const int tyz = tile / ntx, tx = tile - tyz * ntx, tz = tyz / nty, ty = tyz - tz * nty;
                const int x0 = tx * t, x1 = min(x0 + t, nx), y0 = ty * t, y1 = min(y0 + t, ny), z0 = tz * t,
                          z1 = min(z0 + t, nz);
                int own = 0, all = 0;
                for (int z = z0 - 1; z <= z1; z++) {
                    if (!periodic && (z < 0 || z >= nz)) continue;
                    const int cz = (z + nz) % nz;
                    for (int y = y0 - 1; y <= y1; y++) {
                        if (!periodic && (y < 0 || y >= ny)) continue;
                        const int cy = (y + ny) % ny;
                        for (int x = x0 - 1; x <= x1; x++) {
                            if (!periodic && (x < 0 || x >= nx)) continue;
                            const int c = (cz * ny + cy) * nx + (x + nx) % nx;
                            const int cnt = cellptr[c + 1] - cellptr[c];
                            all += cnt;
                            if (z >= z0 && z < z1 && y >= y0 && y < y1 && x >= x0 && x < x1) own += cnt;
                        }
                    }
                }
                worst = fmaxf(worst, (float)own + share * (float)(all - own));
            }
            atomicMax(&s_max, (int)(worst * 1.05f) + 1);
        }
        __syncthreads();
        if (threadIdx.x == 0 && feasible) {
            const float est = (float)s_max;
            // (rows: ~ 0.19 per slot of such a tile, as in vl_setup)
            if (est <= (float)VL_MAX_SLOTS && (conservative_tiles || est * 0.19f <= (float)VL_MAX_ROWS)) s_pick = t;
        }
        __syncthreads();
        if (s_pick) break;
    }
    if (threadIdx.x == 0) {
        VlHeader& H = *hdr;
        const int t = s_pick ? s_pick : 1;
        H.tdim = t;
        H.ntx = (nx + t - 1) / t; H.nty = (ny + t - 1) / t; H.ntz = (nz + t - 1) / t;
        const long long nt = (long long)H.ntx * H.nty * H.ntz;
        if (nt > VL_MAX_TILES) { H.valid = 0; H.fail_code = 6; return; }
        H.n_tiles = (int)nt;
    }
}

__global__ void __launch_bounds__(256) vl_scatter(const VlHeader* hdr, int n, const float4* __restrict__ ref_w,
                                                  const int* __restrict__ atomrank, const int* __restrict__ cellptr,
                                                  const int2* __restrict__ meta, float4* __restrict__ sref,
                                                  int2* __restrict__ smeta, int* __restrict__ scell)
{
    if (!hdr->valid) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 w = ref_w[i];
    const int c = __float_as_int(w.w);
    const int pos = cellptr[c] + atomrank[i];
    sref[pos] = make_float4(w.x, w.y, w.z, __int_as_float(i));
    smeta[pos] = __ldg(&meta[i]);
    scell[pos] = c;
}

__device__ __forceinline__ int vl_tile_of_cell(const VlHeader& H, int c)
{
    const GridS& G = H.grid;
    const int cyz = c / G.nx, cx = c - cyz * G.nx, cz = cyz / G.ny, cy = cyz - cz * G.ny;
    return ((cz / H.tdim) * H.nty + cy / H.tdim) * H.ntx + cx / H.tdim;
}

// tile of a residue = tile of the cell of its anchor (first used) atom: every atom pair of a
// residue pair is then assigned to the same tile (the tile of the lower residue)
__global__ void __launch_bounds__(256) vl_res_tiles(const VlHeader* hdr, int n_res_c, const int* __restrict__ res_anchor,
                                                    const float4* __restrict__ ref_w, int* __restrict__ tile_of_res)
{
    if (!hdr->valid) return;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_res_c) return;
    tile_of_res[r] = vl_tile_of_cell(*hdr, __float_as_int(ref_w[res_anchor[r]].w));
}

// Candidate pairs of the sorted position p (one WARP per position): q > p in the 27 neighbour
// cells, reference minimum-image distance <= r_list, roles and residue exclusion.  The warp
// counts its pairs, reserves room with one atomic and writes the records on a second sweep.
// record: key = tile << 2 rbits | res_lo << rbits | res_hi,
//         value = owner atom | partner atom << 20 | shift id of the partner << 40 | direction << 45
// The OWNER of a pair is its atom in the lower residue (the residue whose tile the pair belongs
// to): in the frame kernel a lane walks the partners of one owner.  The partner carries the lattice
// shift of its image relative to the owner (the negated shift when p and q swap roles: shift ids
// are 13 + (wx, wy, wz) . (1, 3, 9), so -shift is 26 - id and lattice_shift(26 - id) = -lattice_shift(id)
// bit for bit).  direction = 1: the exact arithmetic runs partner -> owner (always p -> q).
__global__ void __launch_bounds__(256) vl_pairs(VlHeader* hdr, int n, int n_ignored, const float4* __restrict__ sref,
                                                const int2* __restrict__ smeta, const int* __restrict__ scell,
                                                const int* __restrict__ cellptr, const int* __restrict__ tile_of_res,
                                                unsigned long long* __restrict__ rec_key,
                                                unsigned long long* __restrict__ rec_val, long long entry_cap)
{
    if (!hdr->valid) return;
    const int p = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (p >= n) return;
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    const VlHeader& H = *hdr;
    const GridS& G = H.grid;
    const float4 pi = sref[p];
    const int2 mi = smeta[p];
    const int c = scell[p];
    const int cyz = c / G.nx, cx = c - cyz * G.nx, cz = cyz / G.ny, cy = cyz - cz * G.ny;
    const float r2 = H.rlist * H.rlist * 1.00001f;
    const unsigned role_i = (unsigned)mi.y >> 30;
    const int res_i = mi.y & 0x3fffffff;
    const int rbits = H.rbits;
    // lane l < 27 owns stencil cell l
    int my_lo = 0, my_hi = 0, my_id = CT_NOSHIFT;
    if (lane < 27) {
        const int dz = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dx = lane % 3 - 1;
        int x = cx + dx, y = cy + dy, z = cz + dz, wx = 0, wy = 0, wz = 0;
        bool ok = true;
        if (x < 0) { x += G.nx; wx = -1; ok = G.periodic; } else if (x >= G.nx) { x -= G.nx; wx = 1; ok = G.periodic; }
        if (y < 0) { y += G.ny; wy = -1; ok = ok && G.periodic; } else if (y >= G.ny) { y -= G.ny; wy = 1; ok = ok && G.periodic; }
        if (z < 0) { z += G.nz; wz = -1; ok = ok && G.periodic; } else if (z >= G.nz) { z -= G.nz; wz = 1; ok = ok && G.periodic; }
        if (ok) {
            const int D = (z * G.ny + y) * G.nx + x;
            my_lo = max(cellptr[D], p + 1);
            my_hi = cellptr[D + 1];
            my_id = (wx + 1) + 3 * (wy + 1) + 9 * (wz + 1);
        }
    }
    long long base = 0;
    for (int pass = 0; pass < 2; pass++) {
        int found = 0;
        for (int l = 0; l < 27; l++) {
            const int lo = __shfl_sync(0xffffffffu, my_lo, l), hi = __shfl_sync(0xffffffffu, my_hi, l);
            if (lo >= hi) continue;
            const int id = __shfl_sync(0xffffffffu, my_id, l);
            const float4 sh = H.shift[id];
            for (int q0 = lo; q0 < hi; q0 += 32) {
                const int q = q0 + lane;
                bool take = false;
                float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                int2 mj = make_int2(0, 0);
                if (q < hi) {
                    pj = sref[q];
                    const float ex = (pj.x + sh.x) - pi.x, ey = (pj.y + sh.y) - pi.y, ez = (pj.z + sh.z) - pi.z;
                    if (ex * ex + ey * ey + ez * ez <= r2) {
                        mj = smeta[q];
                        const int dk = mi.x - mj.x;
                        const unsigned role_j = (unsigned)mj.y >> 30;
                        take = (dk < 0 ? -dk : dk) > n_ignored &&              // residue exclusion (also same residue)
                               ((((role_i & (role_j >> 1)) | (role_j & (role_i >> 1))) & 1u) != 0u);
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, take);
                if (pass == 1 && take) {
                    const long long out = base + found + __popc(m & ltmask);
                    if (out < entry_cap) {
                        const int res_j = mj.y & 0x3fffffff;
                        const int rlo = min(res_i, res_j), rhi = max(res_i, res_j);
                        const unsigned long long tile = (unsigned long long)tile_of_res[rlo];
                        rec_key[out] = (tile << (2 * rbits)) | ((unsigned long long)rlo << rbits) | (unsigned long long)rhi;
                        const unsigned long long ap = (unsigned long long)__float_as_int(pi.w),
                                                 aq = (unsigned long long)__float_as_int(pj.w);
                        rec_val[out] = res_i < res_j ? (ap | (aq << 20) | ((unsigned long long)id << 40))
                                                     : (aq | (ap << 20) | ((unsigned long long)(26 - id) << 40) | (1ull << 45));
                    }
                }
                found += __popc(m);
            }
        }
        if (pass == 0) {
            if (found == 0) return;
            int b = 0;
            if (lane == 0) b = atomicAdd(&hdr->n_entries, found);
            base = (long long)__shfl_sync(0xffffffffu, b, 0);
        }
    }
}

__global__ void vl_after_pairs(VlHeader* hdr, long long entry_cap)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || !hdr->valid) return;
    if ((long long)hdr->n_entries > entry_cap) { hdr->valid = 0; hdr->fail_code = 7; }
    if (hdr->n_entries <= 0) { hdr->valid = 0; hdr->fail_code = 8; }     // no pairs (or overflow): the fall back handles it
}

// bookkeeping after a run-length encode of sorted keys: which = 0 segments (residue pairs of the
// records sorted by tile and residue pair), 1 owner lists (records sorted by tile and owner atom)
__global__ void vl_after_rle(VlHeader* hdr, const unsigned long long* __restrict__ run_key, const int* __restrict__ n_runs,
                             int which)
{
    if (threadIdx.x != 0 || blockIdx.x != 0 || !hdr->valid) return;
    int runs = n_runs[0];
    if (runs > 0 && run_key[runs - 1] == VL_PAD) runs--;         // the padding forms one trailing run
    if (which == 0) hdr->n_segs = runs; else hdr->n_lists = runs;
}

// bundle sort key of every owner list: tile, then length descending
__global__ void __launch_bounds__(256) vl_list_sortkeys(const VlHeader* hdr, const unsigned long long* __restrict__ list_key,
                                                        const int* __restrict__ list_len,
                                                        unsigned long long* __restrict__ skey, int* __restrict__ sval,
                                                        long long cap)
{
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap) return;
    const bool live = hdr->valid && s < hdr->n_lists;
    const unsigned lmax = (1u << VL_LEN_BITS) - 1u;
    skey[s] = live ? (((list_key[s] >> hdr->abits) << VL_LEN_BITS) |
                      (unsigned long long)(lmax - min((unsigned)list_len[s], lmax)))
                   : VL_PAD;
    sval[s] = (int)s;
}

// lower bound of `key` in the sorted array a[0, n)
__device__ __forceinline__ int vl_lower_bound(const unsigned long long* __restrict__ a, int n, unsigned long long key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// last index t in [0, n) with a[t] <= v (a ascending, a[0] <= v)
__device__ __forceinline__ int vl_last_le(const int* __restrict__ a, int n, int v)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int vl_block_excl_scan(int v, int* wsum, int& total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += u;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = wsum[lane];
        int iw = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, iw, d);
            if (lane >= d) iw += u;
        }
        wsum[lane] = iw - w;
        if (lane == 31) wsum[32] = iw;
    }
    __syncthreads();
    total = wsum[32];
    return wsum[warp] + incl - v;
}

// Segments (residue pairs) of every tile: the unique record keys are sorted by tile.
__global__ void __launch_bounds__(256) vl_tile_segs(VlHeader* hdr, VlTile* __restrict__ tiles,
                                                    const unsigned long long* __restrict__ seg_key,
                                                    int* __restrict__ tile_seg0)
{
    if (!hdr->valid) return;
    const int n_tiles = hdr->n_tiles, n_segs = hdr->n_segs, sh = 2 * hdr->rbits;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    const int s0 = t == n_tiles ? n_segs : vl_lower_bound(seg_key, n_segs, (unsigned long long)t << sh);
    tile_seg0[t] = s0;
    if (t < n_tiles) {
        const int s1 = t + 1 == n_tiles ? n_segs : vl_lower_bound(seg_key, n_segs, (unsigned long long)(t + 1) << sh);
        tiles[t].seg0 = s0;
        tiles[t].n_segs = s1 - s0;
        if (s1 - s0 > VL_MAX_SEGS) { hdr->valid = 0; hdr->fail_code = 11; }
    }
}

// Second sort key of every record (in the order of the first sort: tile, residue pair): tile and
// owner atom.  The radix sort is stable, so the partners of one owner stay ordered by residue pair.
// value: partner atom | shift id << 20 | direction << 25 | tile-relative segment << 26
__global__ void __launch_bounds__(256) vl_owner_keys(const VlHeader* hdr, const unsigned long long* __restrict__ rec_key,
                                                     const unsigned long long* __restrict__ rec_val,
                                                     const int* __restrict__ seg_start, const int* __restrict__ tile_seg0,
                                                     unsigned long long* __restrict__ own_key,
                                                     unsigned long long* __restrict__ own_val, long long cap)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap) return;
    unsigned long long k = VL_PAD, v = 0ull;
    if (hdr->valid && e < hdr->n_entries) {
        const unsigned long long tile = rec_key[e] >> (2 * hdr->rbits), r = rec_val[e];
        const int seg = vl_last_le(seg_start, hdr->n_segs, (int)e) - tile_seg0[tile];
        k = (tile << hdr->abits) | (r & 0xfffffull);
        v = ((r >> 20) & 0xfffffull) | (((r >> 40) & 63ull) << 20) | ((unsigned long long)seg << 26);
    }
    own_key[e] = k;
    own_val[e] = v;
}

// Layout of tiles / bundles / rows (one CTA of 1024 threads).  A bundle = 32 owner lists of
// similar length (lane = list), stored as rows of 4 steps x 32 lanes.
__global__ void __launch_bounds__(1024) vl_layout(VlHeader* hdr, VlTile* __restrict__ tiles,
                                                  const unsigned long long* __restrict__ skey_sorted,
                                                  const int* __restrict__ sval_sorted, const int* __restrict__ seg_len,
                                                  int* __restrict__ tile_seg0, int* __restrict__ tile_bundle0,
                                                  int* __restrict__ bundle_row0, long long bundle_cap, long long row_cap)
{
    if (!hdr->valid) return;
    __shared__ int wsum[33];
    __shared__ int carry_a;
    const int n_tiles = hdr->n_tiles, n_segs = hdr->n_lists;     // ("seg" below: an owner list)
    for (int t = threadIdx.x; t <= n_tiles; t += 1024)
        tile_seg0[t] = t == n_tiles ? n_segs : vl_lower_bound(skey_sorted, n_segs, (unsigned long long)t << VL_LEN_BITS);
    __syncthreads();
    // bundles per tile -> tile_bundle0
    if (threadIdx.x == 0) carry_a = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const int nb = t < n_tiles ? (tile_seg0[t + 1] - tile_seg0[t] + 31) / 32 : 0;
        int total;
        const int excl = vl_block_excl_scan(nb, wsum, total);
        if (t < n_tiles) tile_bundle0[t] = carry_a + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_a += total;
        __syncthreads();
    }
    const int n_bundles = carry_a;
    if (threadIdx.x == 0) { tile_bundle0[n_tiles] = n_bundles; hdr->n_bundles = n_bundles; }
    if ((long long)n_bundles > bundle_cap) {
        if (threadIdx.x == 0) { hdr->valid = 0; hdr->fail_code = 9; }
        return;
    }
    __syncthreads();
    // rows per bundle (global running offsets; tile-relative ones are derived below)
    if (threadIdx.x == 0) carry_a = 0;
    __syncthreads();
    for (int base = 0; base < n_bundles; base += 1024) {
        const int g = base + threadIdx.x;
        int rows = 0;
        if (g < n_bundles) {
            const int t = vl_last_le(tile_bundle0, n_tiles, g);
            const int first = tile_seg0[t] + 32 * (g - tile_bundle0[t]);
            // segments are sorted by length (descending, lengths clipped to VL_LEN_BITS in the key):
            // the first one is the longest unless it was clipped, then look at all 32
            int maxlen = seg_len[sval_sorted[first]];
            if (maxlen >= (1 << VL_LEN_BITS) - 1) {
                const int last = min(first + 32, tile_seg0[t + 1]);
                for (int s = first; s < last; s++) maxlen = max(maxlen, seg_len[sval_sorted[s]]);
            }
            rows = (maxlen + 3) / 4;
        }
        int total_r;
        const int er = vl_block_excl_scan(rows, wsum, total_r);
        if (g < n_bundles) bundle_row0[g] = carry_a + er;
        __syncthreads();
        if (threadIdx.x == 0) carry_a += total_r;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bundle_row0[n_bundles] = carry_a;
        hdr->n_rows = carry_a;
        if ((long long)carry_a > row_cap) { hdr->valid = 0; hdr->fail_code = 10; }
    }
    __syncthreads();
    if (!hdr->valid) return;
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        const int b0 = tile_bundle0[t], b1 = tile_bundle0[t + 1];
        const int rows = bundle_row0[b1] - bundle_row0[b0];
        tiles[t].list0 = tile_seg0[t]; tiles[t].n_lists = tile_seg0[t + 1] - tile_seg0[t];
        tiles[t].row0 = bundle_row0[b0]; tiles[t].n_rows = rows;
        // (slot0 / n_slots belong to vl_slot_ranges, which runs beside this kernel on the builder's second stream)
        if (rows > VL_MAX_ROWS) { hdr->valid = 0; hdr->fail_code = 11; }
    }
}

// the two slots every list entry refers to: key = tile << (abits + 5) | atom << 5 | shift id
// (single-tile lists: every used atom becomes a slot as well -- keys [2 cap, 2 cap + n) -- so that
// the guard inside the frame kernel sees ALL atoms: an atom without list pairs that wanders off
// its reference position could otherwise come into contact unnoticed)
__global__ void __launch_bounds__(256) vl_slot_keys(const VlHeader* hdr, const unsigned long long* __restrict__ rec_key,
                                                    const unsigned long long* __restrict__ rec_val,
                                                    unsigned long long* __restrict__ slot_key, long long cap, int n)
{
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n)
        slot_key[2 * cap + e] = (hdr->valid && hdr->n_tiles == 1)
                                    ? (((unsigned long long)e << 5) | (unsigned long long)CT_NOSHIFT) : VL_PAD;
    if (e >= cap) return;
    unsigned long long k0 = VL_PAD, k1 = VL_PAD;
    if (hdr->valid && e < hdr->n_entries) {
        const unsigned long long tile = (rec_key[e] >> (2 * hdr->rbits)) << (hdr->abits + 5), v = rec_val[e];
        k0 = tile | ((v & 0xfffffull) << 5) | (unsigned long long)CT_NOSHIFT;
        k1 = tile | (((v >> 20) & 0xfffffull) << 5) | ((v >> 40) & 31ull);
    }
    slot_key[2 * e] = k0;
    slot_key[2 * e + 1] = k1;
}

__global__ void __launch_bounds__(256) vl_slot_ranges(VlHeader* hdr, VlTile* __restrict__ tiles,
                                                      const unsigned long long* __restrict__ slot_uniq,
                                                      const int* __restrict__ n_uniq_ptr, int* __restrict__ tile_slot0,
                                                      long long slot_cap)
{
    if (!hdr->valid) return;
    int n_uniq = n_uniq_ptr[1];
    if (n_uniq > 0 && slot_uniq[n_uniq - 1] == VL_PAD) n_uniq--;
    const int n_tiles = hdr->n_tiles, tb = hdr->abits + 5;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        hdr->n_slots = n_uniq;
        if ((long long)n_uniq > slot_cap) { hdr->valid = 0; hdr->fail_code = 12; }
    }
    if (t > n_tiles) return;
    const int s0 = t == n_tiles ? n_uniq : vl_lower_bound(slot_uniq, n_uniq, (unsigned long long)t << tb);
    tile_slot0[t] = s0;
    if (t < n_tiles) {
        const int s1 = t + 1 == n_tiles ? n_uniq : vl_lower_bound(slot_uniq, n_uniq, (unsigned long long)(t + 1) << tb);
        tiles[t].slot0 = s0;
        tiles[t].n_slots = s1 - s0;
        if (s1 - s0 > VL_MAX_SLOTS) { hdr->valid = 0; hdr->fail_code = 13; }
    }
}

__global__ void __launch_bounds__(256) vl_slot_fill(const VlHeader* hdr, const unsigned long long* __restrict__ slot_uniq,
                                                    const float4* __restrict__ ref_w, const int* __restrict__ tile_slot0,
                                                    unsigned* __restrict__ slot_word, float4* __restrict__ slot_ref,
                                                    unsigned* __restrict__ tile_bb, long long slot_cap)
{
    if (!hdr->valid) return;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hdr->n_slots || s >= slot_cap) return;
    const unsigned long long k = slot_uniq[s];
    const unsigned atom = (unsigned)((k >> 5) & ((1ull << hdr->abits) - 1ull)), id = (unsigned)(k & 31ull);
    slot_word[s] = atom | (id << 24);
    const float4 r = ref_w[atom];
    slot_ref[s] = make_float4(r.x, r.y, r.z, 0.f);
    // bounding box of the tile's slots at their reference positions (lattice shift included)
    const int t = vl_last_le(tile_slot0, hdr->n_tiles, (int)s);
    const float4 sh = hdr->shift[id];
    const float px = __fadd_rn(r.x, sh.x), py = __fadd_rn(r.y, sh.y), pz = __fadd_rn(r.z, sh.z);
    unsigned* bb = tile_bb + 6 * (size_t)t;
    atomicMin(&bb[0], vl_ordered(px)); atomicMax(&bb[3], vl_ordered(px));
    atomicMin(&bb[1], vl_ordered(py)); atomicMax(&bb[4], vl_ordered(py));
    atomicMin(&bb[2], vl_ordered(pz)); atomicMax(&bb[5], vl_ordered(pz));
}

__global__ void __launch_bounds__(256) vl_bb_init(unsigned* __restrict__ tile_bb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < VL_MAX_TILES * 6) tile_bb[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

// Fixed-point grid of the frame kernel (one CTA): every slot position the guard admits lies within
// `skin` of its reference position, so a tile's positions fit [bb_lo - skin, bb_hi + skin]; the
// common scale maps the widest tile onto 16 bits.  A position is rounded to the grid once, so a
// coordinate difference is off by at most one grid step h (plus the fp32 rounding of the scaling,
// < 0.01 h): |d_grid - d_screen| <= sqrt(3) * 1.02 h =: eq.  Thresholds on the grid distance that
// keep the screen's guarantees: sqrt(c2q_lo) = sqrt(c2_lo) - eq, sqrt(c2q_hi) =
// sqrt(c2_hi) + eq, expressed in grid units^2 with a relative allowance of 1e-6 for the fp32
// evaluation of the squared grid distance (integers below 2^16, three products: < 2e-7).
__global__ void __launch_bounds__(1024) vl_quant_setup(VlHeader* hdr, VlTile* __restrict__ tiles,
                                                       const unsigned* __restrict__ tile_bb, float cutoff)
{
    if (!hdr->valid) return;
    __shared__ float wmax[32];
    const int n_tiles = hdr->n_tiles;
    const float skin = hdr->skin, pad = skin + 1.0e-3f;
    float ext = 0.f;
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        if (tiles[t].n_slots <= 0) continue;
        const unsigned* bb = tile_bb + 6 * (size_t)t;
        for (int k = 0; k < 3; k++) {
            const float lo = vl_unordered(bb[k]), hi = vl_unordered(bb[3 + k]);
            ext = fmaxf(ext, hi - lo);
            (k == 0 ? tiles[t].ox : k == 1 ? tiles[t].oy : tiles[t].oz) = lo - pad;
        }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) ext = fmaxf(ext, __shfl_xor_sync(0xffffffffu, ext, d));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = ext;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; w++) ext = fmaxf(ext, wmax[w]);
        const float range = ext + 2.0f * pad;
        const float qscale = 65000.0f / range;               // grid units per nm (positions stay below 65 001)
        const float eq = 1.7320508f * 1.02f / qscale;
        const float lo = sqrtf(hdr->c2_lo) * 0.999999f - eq, hi = sqrtf(hdr->c2_hi) * 1.000001f + eq;
        hdr->qscale = qscale;
        hdr->c2q_lo = lo * lo * qscale * qscale * 0.999998f;
        hdr->c2q_hi = hi * hi * qscale * qscale * 1.000002f;
        // a grid too coarse to be useful (huge sparse tiles): every frame falls back
        if (!(eq < 0.02f * cutoff) || !(lo > 0.f) || !(range < 1.0e6f)) { hdr->valid = 0; hdr->fail_code = 14; }
    }
}

// residue pair of every segment
__global__ void __launch_bounds__(256) vl_seg_fill(const VlHeader* hdr, const unsigned long long* __restrict__ seg_key,
                                                   uint2* __restrict__ seg_pair, long long cap)
{
    if (!hdr->valid) return;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hdr->n_segs || s >= cap) return;
    const unsigned long long k = seg_key[s];
    const int rbits = hdr->rbits;
    const unsigned long long rmask = (1ull << rbits) - 1ull;
    seg_pair[s] = make_uint2((unsigned)((k >> rbits) & rmask), (unsigned)(k & rmask));
}

// Entry words and owner slots: one block of 128 threads per ROW (thread = lane * 4 + step).
// Lane l of the rows of bundle g walks the owner list 32 g + l of its tile (length-sorted order).
__global__ void __launch_bounds__(VL_ROW) vl_fill(const VlHeader* hdr, const VlTile* __restrict__ tiles,
                                                  const int* __restrict__ tile_bundle0, const int* __restrict__ bundle_row0,
                                                  const int* __restrict__ sval_sorted, const int* __restrict__ list_len,
                                                  const int* __restrict__ list_start,
                                                  const unsigned long long* __restrict__ list_key,
                                                  const unsigned long long* __restrict__ own_val,
                                                  const unsigned long long* __restrict__ slot_uniq,
                                                  const int* __restrict__ tile_slot0, unsigned short* __restrict__ row_owner,
                                                  unsigned* __restrict__ entries)
{
    if (!hdr->valid) return;
    const int R = blockIdx.x;
    if (R >= hdr->n_rows) return;
    __shared__ int sh[4];          // bundle, tile, first row of the bundle
    if (threadIdx.x == 0) {
        const int g = vl_last_le(bundle_row0, hdr->n_bundles, R);
        const int t = vl_last_le(tile_bundle0, hdr->n_tiles, g);
        sh[0] = g; sh[1] = t; sh[2] = bundle_row0[g];
    }
    __syncthreads();
    const int g = sh[0], t = sh[1];
    const VlTile T = tiles[t];
    const int lane = threadIdx.x >> 2, step = (R - sh[2]) * 4 + (threadIdx.x & 3);
    const int list_rel = 32 * (g - tile_bundle0[t]) + lane;
    const unsigned far_off = (unsigned)T.n_slots * VL_SLOT_BYTES;            // pads never pass the screen
    const int tb = hdr->abits + 5;
    const unsigned long long tkey = (unsigned long long)t << tb;
    const unsigned long long* uniq = slot_uniq + tile_slot0[t];
    unsigned word = far_off, owner = 0u;         // (a lane without a list: any finite slot against the far one)
    if (list_rel < T.n_lists) {
        const int l = sval_sorted[T.list0 + list_rel];
        const unsigned long long atom = list_key[l] & ((1ull << hdr->abits) - 1ull);
        owner = (unsigned)vl_lower_bound(uniq, T.n_slots, tkey | (atom << 5) | (unsigned long long)CT_NOSHIFT) * VL_SLOT_BYTES;
        if (step < list_len[l]) {
            const unsigned long long v = own_val[list_start[l] + step];
            const unsigned long long k1 = tkey | ((v & 0xfffffull) << 5) | ((v >> 20) & 31ull);
            const unsigned sj = (unsigned)vl_lower_bound(uniq, T.n_slots, k1);
            word = (sj * VL_SLOT_BYTES) | (unsigned)((v >> 25) & 1ull) | ((unsigned)(v >> 26) << 16);
        }
    }
    // Bank balancing.  In step q of the row the 32 lanes gather 32 unrelated 8-byte slots: a half
    // warp that draws its 16 bank pairs at random needs ~3.1 shared-memory wavefronts where 1 would
    // do.  The ORDER of the four partners inside a lane's quadruple is free, so lane after lane the
    // permutation is taken that puts its partners where the half warp has loaded the bank pairs of
    // the steps least so far (greedy, 24 candidates evaluated by 24 threads): ~2.1 wavefronts.
    // The lanes of a ROW are free as well (every row carries its own owner slots): the owners are
    // dealt to the two half warps so that each sees every bank pair as rarely as possible.
    __shared__ unsigned w_s[VL_ROW];
    __shared__ unsigned char occ[2][4][16];
    __shared__ unsigned short own_s[32];
    __shared__ unsigned char place[32];            // lane of the row -> lane it was computed as
    if ((threadIdx.x & 3) == 0) own_s[lane] = (unsigned short)owner;
    (&occ[0][0][0])[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned char seen[16];
        int fill[2] = {0, 0};
        for (int b = 0; b < 16; b++) seen[b] = 0;
        for (int l = 0; l < 32; l++) {
            const int b = (own_s[l] >> 3) & 15;
            int h = seen[b]++ & 1;                 // members of one bank pair alternate between the halves
            if (fill[h] == 16) h ^= 1;
            place[16 * h + fill[h]++] = (unsigned char)l;
        }
    }
    __syncthreads();
    w_s[threadIdx.x] = word;
    __syncthreads();
    word = w_s[4 * place[lane] + (threadIdx.x & 3)];
    owner = own_s[place[lane]];
    __syncthreads();
    w_s[threadIdx.x] = word;
    __syncthreads();
    const int hw = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (hw < 2) {
        // permutation ln of (0, 1, 2, 3) in the factorial number system
        int perm[4] = {0, 1, 2, 3};
        if (ln < 24) {
            int pool[4] = {0, 1, 2, 3};
            int code = ln;
            const int a = code / 6; code -= 6 * a;
            const int b = code / 2, c = code - 2 * b;
            perm[0] = pool[a];
            for (int k = a; k < 3; k++) pool[k] = pool[k + 1];
            perm[1] = pool[b];
            for (int k = b; k < 2; k++) pool[k] = pool[k + 1];
            perm[2] = pool[c];
            perm[3] = pool[1 - c];
        }
        for (int i = 0; i < 16; i++) {
            const int L = 16 * hw + i;
            unsigned e[4];
#pragma unroll
            for (int q = 0; q < 4; q++) e[q] = w_s[4 * L + q];
            unsigned cost = 0xffffu;
            if (ln < 24) {
                cost = 0u;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned off = e[perm[q]] & 0xfff8u;
                    if (off != far_off) cost += occ[hw][q][(off >> 3) & 15u];
                }
            }
            unsigned key = (cost << 5) | (unsigned)ln;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, d));
            __syncwarp();
            if (ln == (int)(key & 31u)) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned v = e[perm[q]];
                    w_s[4 * L + q] = v;
                    if ((v & 0xfff8u) != far_off) occ[hw][q][((v & 0xfff8u) >> 3) & 15u]++;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    entries[(size_t)R * VL_ROW + threadIdx.x] = w_s[threadIdx.x];
    if ((threadIdx.x & 3) == 0) row_owner[(size_t)R * 32 + lane] = (unsigned short)owner;
}

// ----- frame kernel --------------------------------------------------------------------------
struct VlFrameParams {
    const float* xyz;            // [n_frames][n][3]
    const float* box;            // [n_frames][9] or nullptr
    long long n_frames;
    int n;
    int n_res_c;
    float c2;
    int rounds;                  // work items per SM the frame kernel aims at (it sizes the items itself)
    unsigned int* atom_table;    // dense [n][n] or nullptr
    unsigned int* res_table;     // dense [n_res_c][n_res_c] or nullptr
    unsigned long long* audit;   // [2] pairs audited, disagreements (VL_AUDIT) or nullptr
};

constexpr int VL_AUDIT = 1;      // template flag: run the exact arithmetic on EVERY listed pair and
                                 // count decisions that differ from the screen's

// dynamic shared memory of the frame kernel (bytes); the position array sits at offset 0 so that
// the slot fields of the entry words are ready-made byte offsets
constexpr int VL_QCAP = 1024;                                          // ambiguous pairs parked per tile and frame
constexpr int VL_SM_POS = 0;                                           // uint2 [VL_MAX_SLOTS + 1]: x | y << 16, float bits of 2^23 + z
constexpr int VL_SM_CNT = VL_SM_POS + (VL_MAX_SLOTS + 1) * VL_SLOT_BYTES;   // u32 [VL_MAX_ENTRIES / 4]: byte counters
constexpr int VL_SM_RES = VL_SM_CNT + VL_MAX_ENTRIES;                  // u8 [VL_MAX_SEGS]: frames in which the pair was seen
constexpr int VL_SM_FLAG = VL_SM_RES + VL_MAX_SEGS;                    // u8 [2][VL_MAX_SEGS] "seen in this frame" (frame parity)
constexpr int VL_SM_QUEUE = VL_SM_FLAG + 2 * VL_MAX_SEGS;              // uint2 [VL_QCAP]
constexpr int VL_SM_HDR = VL_SM_QUEUE + VL_QCAP * 8;                   // double[18], float4[27], misc
constexpr int VL_SM_TOTAL = VL_SM_HDR + 18 * 8 + 27 * 16 + 64;
static_assert(VL_SM_POS == 0, "entry words hold byte offsets from the start of the position array");
static_assert(VL_SM_TOTAL <= 232448, "frame kernel exceeds the 227 KB of shared memory a CTA can have");
static_assert(VL_MAX_SEGS <= (1 << VL_SEG_BITS) && VL_MAX_ENTRIES <= (1 << 17) && VL_SEG_BITS + 17 <= 32,
              "parked-pair word layout");
static_assert((VL_MAX_SLOTS + 1) * VL_SLOT_BYTES <= 65536, "slot byte offsets are 16-bit fields");

// 32-bit shared-memory accesses by address (the addresses are kept in registers across the row loop)
__device__ __forceinline__ uint2 vl_lds64(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void vl_sts8(unsigned addr, unsigned v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// the two 16-bit halves of a word as the exact floats 2^23 + value (one PRMT each): differences of
// such floats are exact integers
__device__ __forceinline__ float vl_lo16f(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7610)); }
__device__ __forceinline__ float vl_hi16f(unsigned w) { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7632)); }
// (the z word is stored as those float bits already)
__device__ __forceinline__ float vl_z23f(unsigned w) { return __uint_as_float(w); }

// Exact decision of one parked (ambiguous) list entry and its accounting.
//   e.x = owner slot byte offset | (partner slot byte offset | direction) << 16
//   e.y = byte-counter index of the entry inside its tile (row * 128 + lane * 4 + step) | segment << 17
// MDTraj's arithmetic on the RAW coordinates of the frame, always in the direction p -> q of the
// builder's cell order (direction bit set: partner -> owner; DESIGN.md "direction").  The counter
// word of an entry may be shared with three other parked entries: shared-memory atomic.
__device__ __forceinline__ void vl_exact_account(uint2 e, const float* __restrict__ frame, const NlBox& B, float c2,
                                                 const unsigned* __restrict__ tile_slot_word, unsigned* cnt32,
                                                 unsigned char* segflag)
{
    const unsigned own = (e.x & 0xffffu) / VL_SLOT_BYTES, part = ((e.x >> 16) & 0xfff8u) / VL_SLOT_BYTES;
    const unsigned rev = (e.x >> 16) & 1u;
    const unsigned ia = __ldg(&tile_slot_word[own]) & 0xffffffu, ib = __ldg(&tile_slot_word[part]) & 0xffffffu;
    const float* pa = frame + 3 * (size_t)(rev ? ib : ia);
    const float* pb = frame + 3 * (size_t)(rev ? ia : ib);
    const float d2x = nl_d2_dyn(__ldg(pa), __ldg(pa + 1), __ldg(pa + 2), __ldg(pb), __ldg(pb + 1), __ldg(pb + 2), B);
    if (d2x <= c2) {
        const unsigned idx = e.y & 0x1ffffu;
        atomicAdd(&cnt32[idx >> 2], 1u << (8u * (idx & 3u)));
        segflag[e.y >> 17] = 1;
    }
}

template <int FLAGS>
__global__ void __launch_bounds__(VL_THREADS, 1) vl_frames(const VlFrameParams p, const VlList L, const VlRun R)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint2* pos = reinterpret_cast<uint2*>(smem + VL_SM_POS);
    unsigned* cnt32 = reinterpret_cast<unsigned*>(smem + VL_SM_CNT);
    unsigned char* rescnt = smem + VL_SM_RES;
    unsigned char* segflag2 = smem + VL_SM_FLAG;                               // [2][VL_MAX_SEGS]
    uint2* queue = reinterpret_cast<uint2*>(smem + VL_SM_QUEUE);
    double* hs = reinterpret_cast<double*>(smem + VL_SM_HDR);                  // hinv[9], h[9]
    float4* shift_s = reinterpret_cast<float4*>(smem + VL_SM_HDR + 18 * 8);
    // [0] item, [2 + parity] parked pairs of the frame with that parity
    int* misc = reinterpret_cast<int*>(smem + VL_SM_HDR + 18 * 8 + 27 * 16);
    float* hf = reinterpret_cast<float*>(smem + VL_SM_HDR + 18 * 8 + 27 * 16 + 32);   // 1 / L[3], L[3] of an axis-aligned cell
    const VlHeader& H = *L.hdr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!H.valid) return;                       // every frame is on the fall-back list (vl_all_tripped)
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    if (tid < 18) hs[tid] = tid < 9 ? H.hinv[tid] : H.h[tid - 9];
    if (tid >= 32 && tid < 32 + 27) shift_s[tid - 32] = H.shift[tid - 32];
    if (tid >= 64 && tid < 64 + 3) { hf[tid - 64] = (float)H.hinv[4 * (tid - 64)]; hf[tid - 61] = (float)H.h[4 * (tid - 64)]; }
    const int periodic = H.periodic;
    // axis-aligned periodic cell: the slots are unwrapped in fp32 (vl_unwrap_ortho)
    const bool ortho = periodic && H.h[1] == 0.0 && H.h[2] == 0.0 && H.h[3] == 0.0 && H.h[5] == 0.0 && H.h[6] == 0.0 &&
                       H.h[7] == 0.0;
    const float skin2 = H.skin2, a_assumed = H.a_assumed, c2q_lo = H.c2q_lo, c2q_hi = H.c2q_hi, qscale = H.qscale;
    const int n_tiles = H.n_tiles;
    const bool multi_tile = n_tiles > 1;
    // work item = tile x range of frames: about p.rounds items per SM, of equal length, long enough
    // to amortise the tile set-up / flush and short enough for the byte counters
    long long n_chunks = max(((long long)p.rounds * gridDim.x + n_tiles - 1) / n_tiles,
                             (p.n_frames + VL_ITEM_FRAMES - 1) / VL_ITEM_FRAMES);
    // (a batch too short to give every CTA an item of VL_MIN_ITEM_FRAMES frames -- the tail chunk of a
    // streamed trajectory -- is cut finer: its latency matters, not its efficiency)
    long long max_chunks = (p.n_frames + VL_MIN_ITEM_FRAMES - 1) / VL_MIN_ITEM_FRAMES;
    if (max_chunks * n_tiles < gridDim.x)
        max_chunks = max(max_chunks, min((long long)gridDim.x / n_tiles, (p.n_frames + 2) / 3));
    n_chunks = max(1ll, min(n_chunks, max_chunks));
    const long long item_frames = (p.n_frames + n_chunks - 1) / n_chunks;
    n_chunks = (p.n_frames + item_frames - 1) / item_frames;
    const long long n_items = n_chunks * n_tiles;
    const NlBox B = make_nlbox(p.box);          // the batch's unit cell (frames with another one trip)
    unsigned long long audited = 0, disagreements = 0;
    __syncthreads();

    for (;;) {
        if (tid == 0) misc[0] = atomicAdd(&R.state[1], 1);
        __syncthreads();
        const long long item = misc[0];
        __syncthreads();
        if (item >= n_items) break;
        const int t = (int)(item % n_tiles);
        const long long chunk = item / n_tiles;
        const long long fa = chunk * item_frames, fb = min(p.n_frames, fa + item_frames);
        const VlTile T = L.tiles[t];
        if (T.n_rows == 0) continue;
        const int seg_words = (T.n_segs + 3) / 4;
        // ---- tile set-up: counters, flags, this thread's slots ---------------------------------
        for (int w = tid; w < T.n_rows * (VL_ROW / 4); w += VL_THREADS) cnt32[w] = 0u;
        for (int w = tid; w < seg_words; w += VL_THREADS) {
            reinterpret_cast<unsigned*>(rescnt)[w] = 0u;
            reinterpret_cast<unsigned*>(segflag2)[w] = 0u;
            reinterpret_cast<unsigned*>(segflag2 + VL_MAX_SEGS)[w] = 0u;
        }
        // the "far" slot of padding entries: z = 2^23 - 1 grid units, beyond every threshold
        if (tid == 0) pos[T.n_slots] = make_uint2(0u, 0x4b7fffffu);
        // per slot only the word (atom | shift id << 24) and the prefetched raw coordinates stay in
        // registers across the pair phase; the reference position is re-read (L2) every frame
        unsigned s_word[VL_SPT];
        float raw_x[VL_SPT], raw_y[VL_SPT], raw_z[VL_SPT];
        const float4* slot_ref = L.slot_ref + T.slot0;
        const unsigned* tile_slot_word = L.slot_word + T.slot0;
#pragma unroll
        for (int k = 0; k < VL_SPT; k++) {
            const int s = tid + k * VL_THREADS;
            s_word[k] = 0xffffffffu;
            raw_x[k] = raw_y[k] = raw_z[k] = 0.f;
            if (s < T.n_slots) {
                s_word[k] = tile_slot_word[s];
                const float* q = p.xyz + ((size_t)fa * (size_t)p.n + (size_t)(s_word[k] & 0xffffffu)) * 3;
                raw_x[k] = __ldg(q); raw_y[k] = __ldg(q + 1); raw_z[k] = __ldg(q + 2);
            }
        }
        const uint4* rows4 = reinterpret_cast<const uint4*>(L.entries + (size_t)T.row0 * VL_ROW);
        const unsigned short* rown = L.row_owner + (size_t)T.row0 * 32;
        const int n_rows = T.n_rows;
        const unsigned far_off = (unsigned)T.n_slots * VL_SLOT_BYTES;
        // Per frame f: [phase 1] the pairs frame f-1 could not decide by screening (parked in
        // `queue`) are decided exactly, one per thread, while every thread brings its slots of frame
        // f into `pos`; barrier; [phase 2] the residue pairs seen in frame f-1 are folded into their
        // counters, the rows of frame f are walked; barrier.  The "seen" flags and the queue
        // counter alternate between two copies with the parity of the frame.
        for (long long f = fa; f <= fb; f++) {
            const int par = (int)((f - fa) & 1);
            unsigned char* flag_cur = segflag2 + par * VL_MAX_SEGS;
            unsigned char* flag_prev = segflag2 + (par ^ 1) * VL_MAX_SEGS;
            // ---- phase 1a: parked pairs of the previous frame -----------------------------------
            const int parked = f > fa ? min(misc[2 + (par ^ 1)], VL_QCAP) : 0;
            if (parked > 0) {
                const float* prev = p.xyz + (size_t)(f - 1) * (size_t)p.n * 3;
                for (int w = tid; w < parked; w += VL_THREADS)
                    vl_exact_account(queue[w], prev, B, p.c2, tile_slot_word, cnt32, flag_prev);
            }
            if (f == fb) {
                // after the last frame: fold its flags once the parked pairs are in
                __syncthreads();
                for (int w = tid; w < seg_words; w += VL_THREADS)
                    reinterpret_cast<unsigned*>(rescnt)[w] += reinterpret_cast<const unsigned*>(flag_prev)[w];
                break;
            }
            if (tid == 0) misc[2 + par] = 0;
            // ---- phase 1b: this frame's slots: unwrap next to the reference, guard, fixed point ---
            // (a frame is re-done by the fall back as a WHOLE: lists with several tiles take the
            // decision from the guard pass vl_guard, which looks at every atom of the frame; a
            // single tile holds every atom as a slot and decides here)
            int trip = 0;
            if (multi_tile) trip = R.frame_flag[f];
            else if (p.box != nullptr && tid < 9)
                trip = __float_as_uint(__ldg(&p.box[f * 9 + tid])) != H.box_bits[tid];
#pragma unroll
            for (int k = 0; k < VL_SPT; k++) {
                if (s_word[k] != 0xffffffffu) {
                    const int atom = (int)(s_word[k] & 0xffffffu);
                    const float4 ref = __ldg(&slot_ref[tid + k * VL_THREADS]);
                    const float4 sh = shift_s[s_word[k] >> 24];
                    float x, y, z;
                    if (ortho) vl_unwrap_ortho(hf, raw_x[k], raw_y[k], raw_z[k], ref.x, ref.y, ref.z, x, y, z);
                    else vl_unwrap(hs, hs + 9, periodic, raw_x[k], raw_y[k], raw_z[k], ref.x, ref.y, ref.z, x, y, z);
                    if (!multi_tile) {
                        const float ex = x - ref.x, ey = y - ref.y, ez = z - ref.z;
                        const float a = fmaxf(fabsf(raw_x[k]), fmaxf(fabsf(raw_y[k]), fabsf(raw_z[k])));
                        // (negated comparisons: NaN coordinates trip)
                        trip |= !(ex * ex + ey * ey + ez * ez <= skin2) || !(a <= a_assumed);
                    }
                    // grid coordinates relative to the tile's origin; in range for every frame the
                    // guard admits (vl_quant_setup), clamped for the frames it does not
                    const float gx = __fmul_rn(__fsub_rn(__fadd_rn(x, sh.x), T.ox), qscale),
                                gy = __fmul_rn(__fsub_rn(__fadd_rn(y, sh.y), T.oy), qscale),
                                gz = __fmul_rn(__fsub_rn(__fadd_rn(z, sh.z), T.oz), qscale);
                    // round to the grid by adding 2^23 (round to nearest even, as cvt.rni would, on the
                    // FP32 pipe): the sum's bits are 0x4b000000 | grid coordinate
                    const unsigned qx = __float_as_uint(__fadd_rn(fminf(fmaxf(gx, 0.f), 65535.f), 8388608.f)),
                                   qy = __float_as_uint(__fadd_rn(fminf(fmaxf(gy, 0.f), 65535.f), 8388608.f)),
                                   qz = __float_as_uint(__fadd_rn(fminf(fmaxf(gz, 0.f), 65535.f), 8388608.f));
                    pos[tid + k * VL_THREADS] = make_uint2(__byte_perm(qx, qy, 0x5410), qz);
                    if (f + 1 < fb) {           // prefetch the next frame under the pair phase
                        const float* q = p.xyz + ((size_t)(f + 1) * (size_t)p.n + (size_t)atom) * 3;
                        raw_x[k] = __ldg(q); raw_y[k] = __ldg(q + 1); raw_z[k] = __ldg(q + 2);
                    }
                }
            }
            trip = __syncthreads_or(trip);
            // ---- phase 2a: residue pairs seen in the previous frame (set semantics per frame:
            //      a flag byte is 0 or 1, the counters are bytes too: packed add, no carries) -------
            for (int w = tid; w < seg_words; w += VL_THREADS) {
                const unsigned m = reinterpret_cast<const unsigned*>(flag_prev)[w];
                if (m) {
                    reinterpret_cast<unsigned*>(flag_prev)[w] = 0u;
                    reinterpret_cast<unsigned*>(rescnt)[w] += m;
                }
            }
            if (trip) {
                if (tid == 0 && !multi_tile) {
                    const int w = atomicAdd(&R.state[0], 1);
                    R.tripped[w] = (int)f;
                }
                __syncthreads();
                continue;
            }
            // ---- phase 2b: warp w walks the rows w, w + 32, ...; lane l of a row walks four partners
            //      of ITS owner atom, whose position stays in registers; the entry words of the next
            //      row are requested from L2 before the current ones are worked on ------------------
            const float* frame = p.xyz + (size_t)f * (size_t)p.n * 3;
            const unsigned flag_addr = smem_base + VL_SM_FLAG + (unsigned)par * VL_MAX_SEGS;
            int r = warp;
            uint4 cur = make_uint4(far_off, far_off, far_off, far_off);
            unsigned own = 0u;
            if (r < n_rows) { cur = __ldg(&rows4[(size_t)r * 32 + lane]); own = __ldg(&rown[(size_t)r * 32 + lane]); }
            while (r < n_rows) {
                const int rn = r + VL_THREADS / 32;
                uint4 nxt = cur;
                unsigned nown = own;
                if (rn < n_rows) { nxt = __ldg(&rows4[(size_t)rn * 32 + lane]); nown = __ldg(&rown[(size_t)rn * 32 + lane]); }
                const unsigned ew[4] = {cur.x, cur.y, cur.z, cur.w};
                const uint2 a = vl_lds64(smem_base + own);
                uint2 b[4];
#pragma unroll
                for (int q = 0; q < 4; q++) b[q] = vl_lds64(smem_base + (ew[q] & 0xfff8u));
                const float ax = vl_lo16f(a.x), ay = vl_hi16f(a.x), az = vl_z23f(a.y);
                unsigned c = 0u, amb = 0u;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    // exact integer differences on the grid, squared distance in grid units^2
                    const float dx = __fsub_rn(vl_lo16f(b[q].x), ax), dy = __fsub_rn(vl_hi16f(b[q].x), ay),
                                dz = __fsub_rn(vl_z23f(b[q].y), az);
                    const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                    // d2 <= c2q_lo: certain contact; c2q_lo < d2 < c2q_hi: AMBIGUOUS, parked and decided
                    // by MDTraj's arithmetic in phase 1a of the next frame
                    const bool hit = d2 <= c2q_lo;
                    if (hit) {
                        vl_sts8(flag_addr + (ew[q] >> 16), 1u);      // this residue pair was seen in this frame
                        c += 1u << (8 * q);
                    }
                    amb |= (d2 < c2q_hi && !hit) ? (1u << q) : 0u;
                    if ((FLAGS & VL_AUDIT) && (ew[q] & 0xfff8u) != far_off) {
                        // audit: the exact arithmetic on EVERY listed pair; the screen's decision
                        // must agree wherever it takes one
                        const unsigned ia = __ldg(&tile_slot_word[own / VL_SLOT_BYTES]) & 0xffffffu,
                                       ib = __ldg(&tile_slot_word[(ew[q] & 0xfff8u) / VL_SLOT_BYTES]) & 0xffffffu;
                        const bool rev = (ew[q] & 1u) != 0u;
                        const float* pa = frame + 3 * (size_t)(rev ? ib : ia);
                        const float* pb = frame + 3 * (size_t)(rev ? ia : ib);
                        const bool exact = nl_d2_dyn(__ldg(pa), __ldg(pa + 1), __ldg(pa + 2), __ldg(pb),
                                                     __ldg(pb + 1), __ldg(pb + 2), B) <= p.c2;
                        audited++;
                        if ((hit && !exact) || (!(d2 < c2q_hi) && exact)) disagreements++;
                    }
                }
                if (c) cnt32[r * 32 + lane] += c;        // the entry's byte counters belong to this thread
                if (amb) {
                    // park (rare, divergent); a full queue decides on the spot
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (amb & (1u << q)) {
                            const uint2 e = make_uint2(own | (ew[q] << 16),
                                                       (unsigned)(r * VL_ROW + lane * 4 + q) | ((ew[q] >> 16) << 17));
                            const int w = vl_atoms_add(&misc[2 + par], 1);
                            if (w < VL_QCAP) queue[w] = e;
                            else vl_exact_account(e, frame, B, p.c2, tile_slot_word, cnt32, flag_cur);
                        }
                    }
                }
                r = rn;
                cur = nxt;
                own = nown;
            }
            __syncthreads();
        }
        __syncthreads();
        // ---- end of the item: flush the counters ------------------------------------------------
        if (p.atom_table != nullptr) {
            for (int w = tid; w < T.n_rows * (VL_ROW / 4); w += VL_THREADS) {
                const unsigned c = cnt32[w];
                if (c == 0u) continue;
                const uint4 e4 = rows4[w];
                const unsigned ew[4] = {e4.x, e4.y, e4.z, e4.w};
                const unsigned ia = tile_slot_word[(unsigned)rown[w] / VL_SLOT_BYTES] & 0xffffffu;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const unsigned cq = (c >> (8 * q)) & 0xffu;
                    if (cq) {
                        const unsigned ib = tile_slot_word[(ew[q] & 0xfff8u) / VL_SLOT_BYTES] & 0xffffffu;
                        const unsigned lo = min(ia, ib), hi = max(ia, ib);
                        atomicAdd(&p.atom_table[(size_t)lo * (size_t)p.n + (size_t)hi], cq);
                    }
                }
            }
        }
        if (p.res_table != nullptr) {
            for (int s = tid; s < T.n_segs; s += VL_THREADS) {
                const unsigned c = rescnt[s];
                if (c) {
                    const uint2 rp = L.seg_pair[T.seg0 + s];
                    atomicAdd(&p.res_table[(size_t)rp.x * (size_t)p.n_res_c + (size_t)rp.y], c);
                }
            }
        }
        __syncthreads();
    }
    if ((FLAGS & VL_AUDIT) && p.audit != nullptr) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            audited += __shfl_xor_sync(0xffffffffu, audited, d);
            disagreements += __shfl_xor_sync(0xffffffffu, disagreements, d);
        }
        if (lane == 0) {
            if (audited) atomicAdd(&p.audit[0], audited);
            if (disagreements) atomicAdd(&p.audit[1], disagreements);
        }
    }
}

// Guard pass for lists with more than one tile: a frame is re-done by the fall back as a WHOLE,
// so the trip decision is taken once per frame, over all atoms, before the frame kernel runs;
// tripped frames get their flag set and every tile skips them.  (Single-tile lists check inside
// the frame kernel: the tile sees every atom.)  grid = (atom blocks, frames).
__global__ void __launch_bounds__(256) vl_guard(const VlHeader* hdr, const float* __restrict__ xyz,
                                                const float* __restrict__ box, long long f0, int n,
                                                const float4* __restrict__ ref_w, int* __restrict__ frame_flag,
                                                int* __restrict__ tripped, int* __restrict__ state)
{
    __shared__ double hs[18];
    if (!hdr->valid || hdr->n_tiles <= 1) return;
    if (threadIdx.x < 18) hs[threadIdx.x] = threadIdx.x < 9 ? hdr->hinv[threadIdx.x] : hdr->h[threadIdx.x - 9];
    __syncthreads();
    const long long f = f0 + blockIdx.y;
    const int periodic = hdr->periodic;
    const float skin2 = hdr->skin2, a_assumed = hdr->a_assumed;
    int trip = 0;
    if (box != nullptr && blockIdx.x == 0 && threadIdx.x < 9)
        trip = __float_as_uint(__ldg(&box[f * 9 + threadIdx.x])) != hdr->box_bits[threadIdx.x];
    const float* frame = xyz + (size_t)f * (size_t)n * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float px = __ldg(&frame[3 * (size_t)i]), py = __ldg(&frame[3 * (size_t)i + 1]),
                    pz = __ldg(&frame[3 * (size_t)i + 2]);
        const float4 r = __ldg(&ref_w[i]);
        float x, y, z;
        vl_unwrap(hs, hs + 9, periodic, px, py, pz, r.x, r.y, r.z, x, y, z);
        const float ex = x - r.x, ey = y - r.y, ez = z - r.z;
        const float a = fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz)));
        trip |= !(ex * ex + ey * ey + ez * ez <= skin2) || !(a <= a_assumed);
    }
    trip = __syncthreads_or(trip);
    if (trip && threadIdx.x == 0 && atomicExch(&frame_flag[f], 1) == 0) {
        const int w = atomicAdd(&state[0], 1);
        tripped[w] = (int)f;
    }
}

// an invalid list puts every frame on the fall-back list
__global__ void __launch_bounds__(256) vl_all_tripped(const VlHeader* hdr, int* __restrict__ tripped,
                                                      int* __restrict__ state, long long n_frames)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = 0;          // work queue of the frame kernel
    if (hdr->valid) {
        if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = 0;
        return;
    }
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames;
         f += (long long)gridDim.x * blockDim.x)
        tripped[f] = (int)f;
    if (blockIdx.x == 0 && threadIdx.x == 0) state[0] = (int)n_frames;
}

// ----- host driver ----------------------------------------------------------------------------
struct VlJob {
    const float* xyz; const float* box; long long n_frames;
    int n; int n_res_c;
    const int2* meta;            // [n] {ekey, res_c | role << 30}
    const int* res_anchor;       // [n_res_c] first used atom of each compact residue
    float cutoff; float c2; int n_ignored;
    unsigned int* atom_table; unsigned int* res_table;
    int audit; unsigned long long* d_audit;
    float max_rlist_factor;      // lists with r_list > factor * cutoff are not built (all frames fall back)
    int conservative_tiles;      // 1: tile size from the pessimistic slot estimate (vl_setup, vl_pick_tiles)
    float tile_share = 0.6f;     // optimistic estimate: share of a tile's shell atoms that become its slots
    cudaEvent_t t_begin = nullptr, t_end = nullptr;   // recorded around the frame kernel when set (option time_kernels)
};

#define VL_CK(call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err = std::string(#call) + " failed: " + cudaGetErrorString(e_);         \
            return -2;                                                               \
        }                                                                            \
    } while (0)

inline int vl_bits(long long v)       // bits needed for values in [0, v)
{
    int b = 1;
    while ((1ll << b) < v) b++;
    return b;
}

inline int vl_run_ensure(VlRun& R, long long n_frames, std::string& err)
{
    if (R.frame_cap < n_frames || R.state == nullptr) {
        vl_run_free(R);
        const long long frames = std::max<long long>(n_frames, 1024);
        VL_CK(cudaMalloc(&R.state, sizeof(int) * 2));
        VL_CK(cudaMalloc(&R.tripped, sizeof(int) * (size_t)frames));
        VL_CK(cudaMalloc(&R.frame_flag, sizeof(int) * (size_t)frames));
        VL_CK(cudaMallocHost(&R.h_info, 4 * 4));
        R.frame_cap = frames;
    }
    return 0;
}

inline int vl_ensure(VlScratch& S, int n_atoms, int n_res_c, std::string& err)
{
    if (S.cap_atoms < n_atoms) {
        vl_free(S);
        // (sized with head room: re-allocating ~60 buffers, three streams and the builder graph for a
        // system that is a few atoms larger than the last one costs 30 - 150 ms)
        const int n = (int)std::min<long long>(std::max<long long>(4096, (long long)n_atoms + n_atoms / 4), 0x3fffffff);
        const long long E = std::max<long long>((long long)n * 80, 1 << 15);
        if (2 * E + n > 0x7fffffffll) { err = "system too large for the pair-list path"; return -4; }
        S.entry_cap = E;
        S.L.entry_cap = E;
        S.L.slot_cap = (long long)n * 12 + 8192;
        S.L.row_cap = E / 32 + 1024;
        S.bundle_cap = E / 32 + VL_MAX_TILES;
        VL_CK(cudaMalloc(&S.L.hdr, sizeof(VlHeader)));
        VL_CK(cudaMalloc(&S.L.tiles, sizeof(VlTile) * VL_MAX_TILES));
        VL_CK(cudaMalloc(&S.L.slot_word, sizeof(unsigned) * (size_t)S.L.slot_cap));
        VL_CK(cudaMalloc(&S.L.slot_ref, sizeof(float4) * (size_t)S.L.slot_cap));
        VL_CK(cudaMalloc(&S.L.row_owner, sizeof(unsigned short) * (size_t)S.L.row_cap * 32));
        VL_CK(cudaMalloc(&S.L.tile_bb, sizeof(unsigned) * VL_MAX_TILES * 6));
        VL_CK(cudaMalloc(&S.L.entries, sizeof(unsigned) * (size_t)S.L.row_cap * VL_ROW));
        VL_CK(cudaMalloc(&S.L.seg_pair, sizeof(uint2) * (size_t)E));
        VL_CK(cudaMalloc(&S.ref_raw, sizeof(float4) * (size_t)n));
        VL_CK(cudaMalloc(&S.ref_w, sizeof(float4) * (size_t)n));
        S.L.ref_w = S.ref_w;
        VL_CK(cudaMalloc(&S.atomrank, sizeof(int) * (size_t)n));
        VL_CK(cudaMalloc(&S.cellcnt, sizeof(int) * (size_t)(VL_MAXC + 1)));
        VL_CK(cudaMalloc(&S.cellptr, sizeof(int) * (size_t)(VL_MAXC + 1)));
        VL_CK(cudaMalloc(&S.sref, sizeof(float4) * (size_t)n));
        VL_CK(cudaMalloc(&S.smeta, sizeof(int2) * (size_t)n));
        VL_CK(cudaMalloc(&S.scell, sizeof(int) * (size_t)n));
        VL_CK(cudaMalloc(&S.tile_of_res, sizeof(int) * (size_t)std::max(n_res_c, n)));
        for (int k = 0; k < 2; k++) {
            VL_CK(cudaMalloc(&S.rec_key[k], 8 * (size_t)E));
            VL_CK(cudaMalloc(&S.rec_val[k], 8 * (size_t)E));
            VL_CK(cudaMalloc(&S.skey[k], 8 * (size_t)E));
            VL_CK(cudaMalloc(&S.sval[k], 4 * (size_t)E));
            VL_CK(cudaMalloc(&S.slot_key[k], 8 * (size_t)(2 * E + n)));
        }
        VL_CK(cudaMalloc(&S.seg_key, 8 * (size_t)E));
        VL_CK(cudaMalloc(&S.seg_len, 4 * (size_t)E));
        VL_CK(cudaMalloc(&S.seg_start, 4 * (size_t)(E + 1)));
        VL_CK(cudaMalloc(&S.n_runs, 4 * 3));
        VL_CK(cudaMalloc(&S.own_key, 8 * (size_t)E));
        VL_CK(cudaMalloc(&S.own_val, 8 * (size_t)E));
        VL_CK(cudaMalloc(&S.list_key, 8 * (size_t)E));
        VL_CK(cudaMalloc(&S.list_len, 4 * (size_t)E));
        VL_CK(cudaMalloc(&S.list_start, 4 * (size_t)(E + 1)));
        VL_CK(cudaMalloc(&S.tile_list0, 4 * (VL_MAX_TILES + 1)));
        VL_CK(cudaMalloc(&S.slot_uniq, 8 * (size_t)(2 * E + n)));
        VL_CK(cudaMalloc(&S.tile_seg0, 4 * (VL_MAX_TILES + 1)));
        VL_CK(cudaMalloc(&S.tile_bundle0, 4 * (VL_MAX_TILES + 1)));
        VL_CK(cudaMalloc(&S.tile_slot0, 4 * (VL_MAX_TILES + 1)));
        VL_CK(cudaMalloc(&S.bundle_row0, 4 * (size_t)(S.bundle_cap + 1)));
        size_t need = 0, t = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t, S.rec_key[0], S.rec_key[1], S.rec_val[0], S.rec_val[1], (int)E, 0, 64);
        need = std::max(need, t);
        cub::DeviceRadixSort::SortPairs(nullptr, t, S.skey[0], S.skey[1], S.sval[0], S.sval[1], (int)E, 0, 64);
        need = std::max(need, t);
        cub::DeviceRadixSort::SortKeys(nullptr, t, S.slot_key[0], S.slot_key[1], (int)(2 * E + n), 0, 64);
        need = std::max(need, t);
        cub::DeviceRunLengthEncode::Encode(nullptr, t, S.rec_key[1], S.seg_key, S.seg_len, S.n_runs, (int)E);
        need = std::max(need, t);
        cub::DeviceSelect::Unique(nullptr, t, S.slot_key[1], S.slot_uniq, S.n_runs + 1, (int)(2 * E + n));
        need = std::max(need, t);
        cub::DeviceScan::ExclusiveSum(nullptr, t, S.seg_len, S.seg_start, (int)E);
        need = std::max(need, t);
        VL_CK(cudaMalloc(&S.cub_tmp, need + 256));
        VL_CK(cudaMalloc(&S.cub_tmp2, need + 256));
        S.cub_bytes = need + 256;
        VL_CK(cudaMalloc(&S.d_args, 64));
        VL_CK(cudaStreamCreateWithFlags(&S.bs, cudaStreamNonBlocking));
        VL_CK(cudaEventCreateWithFlags(&S.ev_in, cudaEventDisableTiming));
        VL_CK(cudaEventCreateWithFlags(&S.ev_out, cudaEventDisableTiming));
        VL_CK(cudaStreamCreateWithFlags(&S.aux, cudaStreamNonBlocking));
        VL_CK(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming));
        VL_CK(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming));
        S.cap_atoms = n;
    }
    return 0;
}

// The launch sequence of the builder on stream `st` (per-call arguments: S.d_args).
inline int vl_build_body(const VlJob& job, VlScratch& S, cudaStream_t st, std::string& err, int& launches)
{
    const int n = job.n;
    // (the scratch is allocated with head room; the capacity-sized sorts and scans only cover what a
    // system of this size can need)
    const long long E = std::min<long long>(S.entry_cap, std::max<long long>((long long)n * 80, 1 << 15));
    const VlBuildArgs* args = reinterpret_cast<const VlBuildArgs*>(S.d_args);
    const bool single = n <= VL_SINGLE_TILE_ATOMS;
    // sort keys only carry the bits their fields need (fewer radix passes); the all-ones padding
    // still sorts last: no live key has all of its compared bits set
    const int rbits = vl_bits(job.n_res_c + 1), abits = vl_bits(n + 1), tbits = single ? 0 : 12;
    vl_init<<<1, 1, 0, st>>>(S.L.hdr, args, rbits, abits);
    vl_sample<<<(n + 7) / 8, 128, 0, st>>>(S.L.hdr, args, n, S.ref_raw);
    vl_setup<<<1, 1, 0, st>>>(S.L.hdr, args, n, job.cutoff, job.c2, job.max_rlist_factor, job.conservative_tiles);
    vl_zero_cells<<<64, 256, 0, st>>>(S.L.hdr, S.cellcnt);
    vl_bin<<<(n + 255) / 256, 256, 0, st>>>(S.L.hdr, n, S.ref_raw, S.ref_w, S.atomrank, S.cellcnt);
    vl_scan_cells<<<1, 1024, 0, st>>>(S.L.hdr, S.cellcnt, S.cellptr, n);
    vl_pick_tiles<<<1, 1024, 0, st>>>(S.L.hdr, S.cellptr, n, job.conservative_tiles, job.tile_share);
    vl_scatter<<<(n + 255) / 256, 256, 0, st>>>(S.L.hdr, n, S.ref_w, S.atomrank, S.cellptr, job.meta, S.sref, S.smeta,
                                                S.scell);
    vl_res_tiles<<<(job.n_res_c + 255) / 256, 256, 0, st>>>(S.L.hdr, job.n_res_c, job.res_anchor, S.ref_w, S.tile_of_res);
    VL_CK(cudaMemsetAsync(S.rec_key[0], 0xff, 8 * (size_t)E, st));
    vl_pairs<<<(n + 7) / 8, 256, 0, st>>>(S.L.hdr, n, job.n_ignored, S.sref, S.smeta, S.scell, S.cellptr, S.tile_of_res,
                                          S.rec_key[0], S.rec_val[0], E);
    vl_after_pairs<<<1, 1, 0, st>>>(S.L.hdr, E);
    VL_CK(cudaMemsetAsync(S.seg_len, 0, 4 * (size_t)E, st));
    VL_CK(cudaMemsetAsync(S.list_len, 0, 4 * (size_t)E, st));
    VL_CK(cudaMemsetAsync(S.n_runs, 0, 12, st));
    // sort 1: by tile and residue pair -> segments (the residue pairs of every tile)
    size_t tb = S.cub_bytes;
    VL_CK(cub::DeviceRadixSort::SortPairs(S.cub_tmp, tb, S.rec_key[0], S.rec_key[1], S.rec_val[0], S.rec_val[1], (int)E, 0,
                                          2 * rbits + tbits, st));
    // From the sorted records two independent chains of small dependent launches follow; they run
    // side by side (the builder is launch-latency bound):
    //   aux:  the SLOTS of every tile (atoms and their periodic images): keys, sort, unique, ranges,
    //         reference positions, fixed-point grid;
    //   main: segments, owner lists, bundles, row layout.
    VL_CK(cudaEventRecord(S.ev_fork, st));
    VL_CK(cudaStreamWaitEvent(S.aux, S.ev_fork, 0));
    {
        cudaStream_t sa = S.aux;
        const long long n_keys = 2 * E + n;
        vl_slot_keys<<<(unsigned)((std::max<long long>(E, n) + 255) / 256), 256, 0, sa>>>(S.L.hdr, S.rec_key[1], S.rec_val[1],
                                                                                          S.slot_key[0], E, n);
        size_t tb2 = S.cub_bytes;
        VL_CK(cub::DeviceRadixSort::SortKeys(S.cub_tmp2, tb2, S.slot_key[0], S.slot_key[1], (int)n_keys, 0, abits + 5 + tbits, sa));
        tb2 = S.cub_bytes;
        VL_CK(cub::DeviceSelect::Unique(S.cub_tmp2, tb2, S.slot_key[1], S.slot_uniq, S.n_runs + 1, (int)n_keys, sa));
        vl_slot_ranges<<<(VL_MAX_TILES + 1 + 255) / 256, 256, 0, sa>>>(S.L.hdr, S.L.tiles, S.slot_uniq, S.n_runs, S.tile_slot0,
                                                                       S.L.slot_cap);
        vl_bb_init<<<(VL_MAX_TILES * 6 + 255) / 256, 256, 0, sa>>>(S.L.tile_bb);
        vl_slot_fill<<<(unsigned)((S.L.slot_cap + 255) / 256), 256, 0, sa>>>(S.L.hdr, S.slot_uniq, S.ref_w, S.tile_slot0,
                                                                             S.L.slot_word, S.L.slot_ref, S.L.tile_bb,
                                                                             S.L.slot_cap);
        vl_quant_setup<<<1, 1024, 0, sa>>>(S.L.hdr, S.L.tiles, S.L.tile_bb, job.cutoff);
        VL_CK(cudaEventRecord(S.ev_join, sa));
    }
    tb = S.cub_bytes;
    VL_CK(cub::DeviceRunLengthEncode::Encode(S.cub_tmp, tb, S.rec_key[1], S.seg_key, S.seg_len, S.n_runs, (int)E, st));
    vl_after_rle<<<1, 1, 0, st>>>(S.L.hdr, S.seg_key, S.n_runs, 0);
    tb = S.cub_bytes;
    VL_CK(cub::DeviceScan::ExclusiveSum(S.cub_tmp, tb, S.seg_len, S.seg_start, (int)E, st));
    vl_tile_segs<<<(VL_MAX_TILES + 1 + 255) / 256, 256, 0, st>>>(S.L.hdr, S.L.tiles, S.seg_key, S.tile_seg0);
    vl_seg_fill<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(S.L.hdr, S.seg_key, S.L.seg_pair, E);
    // sort 2 (stable): by tile and owner atom -> owner lists, each still ordered by residue pair
    vl_owner_keys<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(S.L.hdr, S.rec_key[1], S.rec_val[1], S.seg_start, S.tile_seg0,
                                                               S.rec_key[0], S.rec_val[0], E);
    tb = S.cub_bytes;
    VL_CK(cub::DeviceRadixSort::SortPairs(S.cub_tmp, tb, S.rec_key[0], S.own_key, S.rec_val[0], S.own_val, (int)E, 0,
                                          abits + tbits, st));
    tb = S.cub_bytes;
    VL_CK(cub::DeviceRunLengthEncode::Encode(S.cub_tmp, tb, S.own_key, S.list_key, S.list_len, S.n_runs + 2, (int)E, st));
    vl_after_rle<<<1, 1, 0, st>>>(S.L.hdr, S.list_key, S.n_runs + 2, 1);
    tb = S.cub_bytes;
    VL_CK(cub::DeviceScan::ExclusiveSum(S.cub_tmp, tb, S.list_len, S.list_start, (int)E, st));
    // bundles: the lists of a tile sorted by length, 32 to a bundle
    vl_list_sortkeys<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(S.L.hdr, S.list_key, S.list_len, S.skey[0], S.sval[0], E);
    tb = S.cub_bytes;
    VL_CK(cub::DeviceRadixSort::SortPairs(S.cub_tmp, tb, S.skey[0], S.skey[1], S.sval[0], S.sval[1], (int)E, 0,
                                          VL_LEN_BITS + tbits, st));
    vl_layout<<<1, 1024, 0, st>>>(S.L.hdr, S.L.tiles, S.skey[1], S.sval[1], S.list_len, S.tile_list0, S.tile_bundle0,
                                  S.bundle_row0, S.bundle_cap, S.L.row_cap);
    VL_CK(cudaStreamWaitEvent(st, S.ev_join, 0));
    vl_fill<<<(unsigned)S.L.row_cap, VL_ROW, 0, st>>>(S.L.hdr, S.L.tiles, S.tile_bundle0, S.bundle_row0, S.sval[1],
                                                      S.list_len, S.list_start, S.list_key, S.own_val, S.slot_uniq,
                                                      S.tile_slot0, S.L.row_owner, S.L.entries);
    VL_CK(cudaGetLastError());
    launches += 33;
    return 0;
}

// everything a captured builder graph has baked in
struct VlGraphKey {
    int n, n_res_c, n_ignored, conservative_tiles;
    float cutoff, c2, max_rlist_factor, tile_share;
    const void* meta; const void* res_anchor;
    long long entry_cap;
};
static_assert(sizeof(VlGraphKey) <= 64, "VlScratch::gkey");

// Build the list of a batch (asynchronous, everything stays on the device).  The builder is ~35
// small dependent launches: issued one by one they cost the host 4-10 us each (more than the
// kernels take on a busy host, several times more while a host-to-device copy saturates the link),
// so the sequence is captured ONCE per system into a CUDA graph on the builder's own stream (the
// caller's may be the legacy default stream, which cannot be captured) and replayed with one launch;
// what differs between builds is read from S.d_args.  If capturing fails the launches are issued
// directly, as before.
inline int vl_build(const VlJob& job, VlScratch& S, cudaStream_t st, std::string& err, int& launches, bool use_graph = true)
{
    VlBuildArgs a;
    a.xyz = job.xyz; a.box = job.box; a.n_frames = job.n_frames;
    if (use_graph && S.graph_state >= 0) {
        VlGraphKey key;
        memset(&key, 0, sizeof(key));
        key.n = job.n; key.n_res_c = job.n_res_c; key.n_ignored = job.n_ignored; key.conservative_tiles = job.conservative_tiles;
        key.cutoff = job.cutoff; key.c2 = job.c2; key.max_rlist_factor = job.max_rlist_factor; key.tile_share = job.tile_share;
        key.meta = job.meta; key.res_anchor = job.res_anchor; key.entry_cap = S.entry_cap;
        VL_CK(cudaEventRecord(S.ev_in, st));
        VL_CK(cudaStreamWaitEvent(S.bs, S.ev_in, 0));
        // (pageable source: the driver has taken its copy of `a` when the call returns)
        VL_CK(cudaMemcpyAsync(S.d_args, &a, sizeof(a), cudaMemcpyHostToDevice, S.bs));
        if (S.gexec == nullptr || memcmp(&key, S.gkey, sizeof(key)) != 0) {
            const auto t_capture = std::chrono::steady_clock::now();
            if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
            cudaGraph_t graph = nullptr;
            int dummy = 0;
            std::string cerr;
            bool ok = cudaStreamBeginCapture(S.bs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = vl_build_body(job, S, S.bs, cerr, dummy) == 0;
                ok = (cudaStreamEndCapture(S.bs, &graph) == cudaSuccess) && ok && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&S.gexec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) {
                cudaGetLastError();
                if (S.gexec) { cudaGraphExecDestroy(S.gexec); S.gexec = nullptr; }
                S.graph_state = -1;
            } else {
                memcpy(S.gkey, &key, sizeof(key));
                S.graph_state = 1;
            }
            S.captures++;
            S.capture_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_capture).count();
        }
        if (S.graph_state == 1) {
            VL_CK(cudaGraphLaunch(S.gexec, S.bs));
            VL_CK(cudaEventRecord(S.ev_out, S.bs));
            VL_CK(cudaStreamWaitEvent(st, S.ev_out, 0));
            launches += 33;
            return 0;
        }
        // capture failed: the argument copy below is ordered behind the one queued on S.bs
        VL_CK(cudaEventRecord(S.ev_out, S.bs));
        VL_CK(cudaStreamWaitEvent(st, S.ev_out, 0));
    }
    VL_CK(cudaMemcpyAsync(S.d_args, &a, sizeof(a), cudaMemcpyHostToDevice, st));
    return vl_build_body(job, S, st, err, launches);
}

// Walk a list over the frames of the job; frames that cannot use it end up on R.tripped (count in
// R.state[0]) for the caller's fall back.  The list may have been built from another batch of the
// same trajectory: the guard decides frame by frame.
inline int vl_run_frames(const VlJob& job, const VlList& L, const VlRun& R, int sm_count, cudaStream_t st,
                         std::string& err, int& launches)
{
    static bool attr_set[2] = {false, false};
    if (!attr_set[0]) {
        VL_CK(cudaFuncSetAttribute(vl_frames<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, VL_SM_TOTAL));
        attr_set[0] = true;
    }
    if (job.audit && !attr_set[1]) {
        VL_CK(cudaFuncSetAttribute(vl_frames<VL_AUDIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, VL_SM_TOTAL));
        attr_set[1] = true;
    }
    vl_all_tripped<<<64, 256, 0, st>>>(L.hdr, R.tripped, R.state, job.n_frames);
    if (job.n > VL_SINGLE_TILE_ATOMS) {
        VL_CK(cudaMemsetAsync(R.frame_flag, 0, sizeof(int) * (size_t)job.n_frames, st));
        const int blocks = std::max(1, std::min((job.n + 255) / 256, 64));
        for (long long f0 = 0; f0 < job.n_frames; f0 += 32768) {
            const unsigned nf = (unsigned)std::min<long long>(32768, job.n_frames - f0);
            vl_guard<<<dim3(blocks, nf), 256, 0, st>>>(L.hdr, job.xyz, job.box, f0, job.n, L.ref_w, R.frame_flag,
                                                       R.tripped, R.state);
            launches++;
        }
    }
    VlFrameParams p;
    p.xyz = job.xyz; p.box = job.box; p.n_frames = job.n_frames; p.n = job.n; p.n_res_c = job.n_res_c; p.c2 = job.c2;
    // single tile: every item costs the same, so one item per SM (static split) wastes nothing on
    // set-up / flush; lists of many tiles have tiles of different weight and are dealt dynamically
    p.rounds = job.n <= VL_SINGLE_TILE_ATOMS ? 1 : 6;
    p.atom_table = job.atom_table; p.res_table = job.res_table; p.audit = job.d_audit;
    if (job.t_begin) VL_CK(cudaEventRecord(job.t_begin, st));
    if (job.audit) vl_frames<VL_AUDIT><<<sm_count, VL_THREADS, VL_SM_TOTAL, st>>>(p, L, R);
    else vl_frames<0><<<sm_count, VL_THREADS, VL_SM_TOTAL, st>>>(p, L, R);
    VL_CK(cudaGetLastError());
    if (job.t_end) VL_CK(cudaEventRecord(job.t_end, st));
    launches += 2;
    return 0;
}

}  // namespace cmb
