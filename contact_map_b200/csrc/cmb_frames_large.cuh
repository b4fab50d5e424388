// cmb_frames_large.cuh -- global-memory cell path for systems that do not fit the
// fused shared-memory kernel (n_used > 6144, up to the 100k-atom solvated system and beyond).
//
// Same pair-phase code as the fused kernel (cmb_cells.cuh::ct_pair_phase, LargeTraits); the
// cell arrays and the cell-sorted float4 copy of the frame live in HBM/L2 instead of
// shared memory, one kernel per stage, frame `blockIdx.y` of a batch in slice blockIdx.y of
// every scratch array (LargeBatch):
//   kl_prepare  grid of the frame (+ bounding box for open boundaries), zero counters
//   kl_bin      cell id + rank + wrapped position per atom (L2 atomics on the cell histogram),
//               max |coordinate| of the frame
//   kl_scan     exclusive scan over cells + occupied-cell list (one CTA per frame)
//   kl_scatter  cell-sorted float4 {wrapped x, y, z, atom index | role << 30}
//   kl_pairs    persistent CTAs, one warp per task (one task counter per frame)
// Two schedules (large_run): BATCHED -- many frames per launch, per-frame residue-pair bitmaps
// in global memory (BatchSink), binning of the next batch on an auxiliary stream; and ONE FRAME
// AT A TIME -- the whole GPU on a frame, per-frame residue de-duplication
// (contact_map.py:601-607,740) by a frame-stamp table: frames are processed in stream order, so
// atomicMax(stamp[slot], seq) < seq identifies the first hit of a residue pair in the current
// frame (StampSink; also the schedule of the hashed count tables).
#pragma once
#include <string>

#include "cmb_cells.cuh"

namespace cmb {

constexpr int LG_MAXC = 1 << 21;
#ifndef LG_PAIR_THREADS_CFG
#define LG_PAIR_THREADS_CFG 256
#endif
constexpr int LG_PAIR_THREADS = LG_PAIR_THREADS_CFG;
#ifndef LG_MIN_CTAS
#define LG_MIN_CTAS 2
#endif
constexpr int LG_BIN_THREADS = 256;      // single-CTA set-up / scan kernels: small enough to share an SM with kl_pairs
constexpr int LG_RCACHE = 128;          // per-warp cache of stamped residue pairs (entries)

// Per-frame binning buffers.  Two sets: the binning kernels of frame f + 1 run on a second stream
// while the pair kernel of frame f is still busy (it is latency-bound on the count tables and
// leaves most issue slots idle), see large_run.
struct LargeBins {
    GridS* d_grid = nullptr;
    int* d_cellcnt = nullptr;      // [LG_MAXC + 1]
    int* d_cellptr = nullptr;      // [LG_MAXC + 1]
    int* d_occ = nullptr;          // [LG_MAXC]
    int* d_misc = nullptr;         // [0] n_occ, [1] task counter, [2] max |coordinate| (float bits)
    float4* d_wrapped = nullptr;   // [n] {wrapped x, y, z, cell id}
    int* d_atomrank = nullptr;     // [n]
    float4* d_sorted = nullptr;    // [n]
    float2* d_sorted_kr = nullptr;   // [n] {exclusion key, compact residue}
    cudaEvent_t binned = nullptr;    // binning kernels of the set's current frame have finished
    cudaEvent_t paired = nullptr;    // pair kernel of the set's current frame has finished
    bool paired_recorded = false;
};

struct LargeScratch {
    int cap_atoms = 0;
    LargeBins bins[2];
    cudaStream_t aux = nullptr;      // stream of the binning kernels
    cudaEvent_t entry = nullptr;     // everything the caller queued before large_run
    // batched mode (large_run_batched): one slice per frame of a batch
    LargeBins batch[2];              // two sets: binning of batch k + 1 overlaps the pair kernel of batch k
    long long batch_frames = 0;
    int batch_atoms = 0;
    size_t batch_bm_words = 0;
    unsigned int* d_bitmaps[2] = {nullptr, nullptr};   // per set: [batch_frames][R_c^2 / 32] per-frame residue-pair bitmaps
    cudaStream_t aux_batch = nullptr;
    cudaEvent_t entry_batch = nullptr;
    unsigned int* d_stamp = nullptr;
    size_t stamp_elems = 0;
    unsigned int seq = 0;
};

inline void large_scratch_free(LargeScratch& S)
{
    for (LargeBins& B : S.bins) {
        cudaFree(B.d_grid); cudaFree(B.d_cellcnt); cudaFree(B.d_cellptr); cudaFree(B.d_occ);
        cudaFree(B.d_misc); cudaFree(B.d_wrapped); cudaFree(B.d_atomrank); cudaFree(B.d_sorted);
        cudaFree(B.d_sorted_kr);
        if (B.binned) cudaEventDestroy(B.binned);
        if (B.paired) cudaEventDestroy(B.paired);
    }
    if (S.aux) cudaStreamDestroy(S.aux);
    if (S.entry) cudaEventDestroy(S.entry);
    for (LargeBins& B : S.batch) {
        cudaFree(B.d_grid); cudaFree(B.d_cellcnt); cudaFree(B.d_cellptr); cudaFree(B.d_occ);
        cudaFree(B.d_misc); cudaFree(B.d_wrapped); cudaFree(B.d_atomrank); cudaFree(B.d_sorted);
        cudaFree(B.d_sorted_kr);
        if (B.binned) cudaEventDestroy(B.binned);
        if (B.paired) cudaEventDestroy(B.paired);
    }
    cudaFree(S.d_bitmaps[0]); cudaFree(S.d_bitmaps[1]);
    if (S.aux_batch) cudaStreamDestroy(S.aux_batch);
    if (S.entry_batch) cudaEventDestroy(S.entry_batch);
    cudaFree(S.d_stamp);
    S = LargeScratch();
}

struct LargeJob {
    const float* xyz; const float* box; long long n_frames; long long frame0;
    int n; const int2* meta; float cutoff; float c2; int boundary_ulps;
    int n_ignored; int n_res_c;
    unsigned int* atom_table; unsigned int* res_table;
    unsigned long long* atom_keys; long long atom_cap;
    unsigned long long* res_keys; long long res_cap;
    unsigned long long* counters; float4* boundary; long long boundary_cap; int count_tests;
    int all_roles;
    int atom_log2, res_log2;        // hashed tables when > 0 (table pointers are uint64 slot arrays)
    unsigned long long* hstats;
    unsigned long long* ovf_list; long long ovf_cap;
    int batch_mode;                 // 0: one frame at a time, 1: batched when it pays (default), 2: batched whenever possible
};

// ----- traits + hit sink of the global-memory path -------------------------------------
struct LargeTraits {
    typedef int pos_t;
    typedef unsigned int cand_t;                    // sorted position [0,20) | wrap flags [20,23)
    typedef uint4 hit_t;          // {own base, jpos [0,20) | shift id [20,25), own-atom mask, -}
    // task word: cell [0,28) | interior | pair
    static constexpr unsigned TASK_CMASK = 0x0fffffffu, TASK_INTERIOR = 0x10000000u, TASK_PAIR = 0x20000000u,
                              TASK_EMPTY = 0xffffffffu;
    static constexpr int TASK_BATCH = 1;          // the whole GPU shares one frame's task list
    static constexpr bool CELL_ROLES = false;
    static __device__ __forceinline__ cand_t cand(int pos, int wrap)
    {
        return (unsigned)pos | ((unsigned)wrap << 20);
    }
    static __device__ __forceinline__ int cand_pos(cand_t e) { return (int)(e & 0xfffffu); }
    static __device__ __forceinline__ int cand_wrap(cand_t e) { return (int)(e >> 20); }
    static __device__ __forceinline__ unsigned jp(int jpos, int id) { return (unsigned)jpos | ((unsigned)id << 20); }
    static __device__ __forceinline__ hit_t entry(int ibase, unsigned jp, unsigned mask)
    {
        return make_uint4((unsigned)ibase, jp, mask, 0u);
    }
    static __device__ __forceinline__ hit_t with_mask(hit_t e, unsigned mask) { return make_uint4(e.x, e.y, mask, 0u); }
    static __device__ __forceinline__ int e_ibase(hit_t e) { return (int)e.x; }
    static __device__ __forceinline__ int e_jpos(hit_t e) { return (int)(e.y & 0xfffffu); }
    static __device__ __forceinline__ int e_id(hit_t e) { return (int)(e.y >> 20); }
    static __device__ __forceinline__ unsigned e_mask(hit_t e) { return e.z; }
    static __device__ __forceinline__ int next_tasks(int* counter, int count)        // task counter in global memory
    {
        int old;
        asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(counter), "r"(count) : "memory");
        return old;
    }
    // sorted[].w bits: atom index [0,30) | role [30,32); exclusion keys in a separate array
    static __device__ __forceinline__ int atom(unsigned bits) { return (int)(bits & 0x3fffffffu); }
    static __device__ __forceinline__ unsigned role(unsigned bits) { return bits >> 30; }
    static __device__ __forceinline__ float ekey_f(const PairCtx<LargeTraits>& cx, int pos, unsigned)
    {
        return __ldg(&cx.kr[pos]).x;
    }
    static __device__ __forceinline__ void decode(int c, const GridS& G, int& cx, int& cy, int& cz)
    {
        const int cyz = c / G.nx;
        cx = c - cyz * G.nx;
        cz = cyz / G.ny;
        cy = cyz - cz * G.ny;
    }
};

// ----- hashed count tables (north_star: "dense or hashed count matrix") -------------------------
// Open addressing with linear probing over uint64 slots, slot = (key + 1) << 24 | count; an
// all-zero slot is empty, so a zero-filled buffer is an empty table.  key = flat index of the
// dense layout (lo * n + hi < 2^40), counts < 2^24 (checked on the host).  The pair universe of a
// trajectory is a few tens of partners per atom, so the table is a few tens of MB and stays in
// L2, where the dense [n][n] table of a 100k-atom system (40 GB) turns every hit into DRAM
// traffic.
constexpr int HT_COUNT_BITS = 24;
constexpr int HT_MAX_PROBES = 1 << 14;

__device__ __forceinline__ unsigned ht_hash(unsigned long long k)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (unsigned)k;
}

// slot of `key1` (= key + 1), claiming an empty one if the key is new; 0xffffffff: table full
// (the caller then parks the hit in the overflow list).  stats[which] counts the new keys.
// Probing is bucketed: the probe sequence starts at a 32-byte aligned group of four slots and one
// probe reads the whole group (one L2 sector, two 16-byte loads), so a lookup is one round trip
// for > 95 % of the hits at load <= 0.7 instead of one round trip per slot.
__device__ __forceinline__ ulonglong2 ht_load2(const unsigned long long* p)
{
    ulonglong2 v;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ht_find_or_insert(unsigned long long* slots, unsigned mask,
                                                      unsigned long long key1, unsigned long long* stats,
                                                      int which)
{
    unsigned h0 = (ht_hash(key1) & mask) & ~3u;
    for (int probe = 0; probe < HT_MAX_PROBES / 4; probe++) {
        const ulonglong2 lo = ht_load2(&slots[h0]), hi = ht_load2(&slots[h0 + 2]);
        const unsigned long long s[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((s[j] >> HT_COUNT_BITS) == key1) return h0 + j;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if ((s[j] >> HT_COUNT_BITS) == 0ull) {
                const unsigned long long old = atomicCAS(&slots[h0 + j], 0ull, key1 << HT_COUNT_BITS);
                if (old == 0ull) {
                    atomicAdd(&stats[which], 1ull);
                    return h0 + j;
                }
                if ((old >> HT_COUNT_BITS) == key1) return h0 + j;
            }
        }
        h0 = (h0 + 4u) & mask;
    }
    return 0xffffffffu;
}

// A hit that found no slot is parked in the overflow list (stats[2] = entries requested); the
// host grows the table and adds the parked hits afterwards, so nothing is ever dropped.  Atom
// entry: key1; residue entry: 1 << 63 | key1 << 24 | (frame sequence number & 0xffffff) -- the
// host removes duplicates of one residue pair within one frame before adding.
struct HashOverflow {
    unsigned long long* list;
    long long cap;
};
__device__ __forceinline__ void ht_park(const HashOverflow& o, unsigned long long* stats, unsigned long long entry)
{
    const unsigned long long w = atomicAdd(&stats[2], 1ull);
    if ((long long)w < o.cap) o.list[w] = entry;
}

// Residue exclusion + accounting of one hit: RED into the dense atom table; per-frame residue
// de-duplication through the frame-stamp table (frames run in stream order, so
// atomicMax(stamp, seq) < seq is the first hit of that residue pair in the current frame).  The
// atomicMax result is only consumed at the NEXT call (software pipelining): the L2 round trip of
// the stamp table overlaps the pair tests in between instead of stalling the warp.  Lanes of one
// drain that hit the same residue pair are merged first (__match_any_sync), which removes most of
// the atomics for 3-atom solvent residues.
struct StampSink {
    static constexpr int MIN_CTAS = LG_MIN_CTAS;    // resident CTAs per SM of kl_pairs with this sink
    const int2* meta;               // [n] {ekey, res_c | role << 30}
    unsigned int* atom_table;
    unsigned int* stamp;            // nullptr: residues not requested
    unsigned int* res_table;
    // hashed tables (mask != 0): the table pointers above are uint64 slot arrays
    unsigned int atom_mask, res_mask;
    unsigned long long* hstats;     // [0] new atom keys, [1] new residue keys, [2] parked hits
    HashOverflow ovf;
    unsigned int* rcache;           // this warp's residue pairs already stamped in this frame (shared memory)
    unsigned long long* res_keys;
    long long res_cap;
    unsigned long long* counters;
    unsigned int seq;
    unsigned int n;
    unsigned int n_res_c;
    long long frame_id;
    // pipeline state of this lane
    unsigned int pend_old;
    unsigned int pend_lo, pend_hi, pend_slot;
    int pending;
    int ra, rb;                     // compact residues of the pair handed to prepare()

    __device__ __forceinline__ void set_frame(int) {}

    __device__ __forceinline__ void resolve()
    {
        if (pending && pend_old < seq) {
            const size_t slot = (size_t)pend_lo * n_res_c + (size_t)pend_hi;
            if (res_table != nullptr) {
                if (res_mask) atomicAdd(reinterpret_cast<unsigned long long*>(res_table) + pend_slot, 1ull);
                else atomicAdd(&res_table[slot], 1u);
            }
            if (res_keys != nullptr) {
                const unsigned long long w = atomicAdd(&counters[1], 1ull);
                if ((long long)w < res_cap)
                    res_keys[w] = ((unsigned long long)frame_id << 40) | ((unsigned long long)pend_lo << 20) |
                                  (unsigned long long)pend_hi;
            }
        }
        pending = 0;
    }

    // residues from the per-position array: the positions of a task are neighbours in memory
    // (L1 hits), where meta[atom] would be two random L2 sectors per pair
    __device__ __forceinline__ void prepare(const PairCtx<LargeTraits>& cx, int ipos, int jpos, int, int)
    {
        ra = __float_as_int(__ldg(&cx.kr[ipos]).y);
        rb = __float_as_int(__ldg(&cx.kr[jpos]).y);
    }

    // called by the lanes of one drain whose pair is a hit
    __device__ __forceinline__ void account(int ia, int ib)
    {
#ifdef CMB_EXPERIMENT_NO_SINK
        if (ia == -12345) atomicAdd(&atom_table[0], (unsigned)ib);    // keep the call alive, touch nothing
        return;
#endif
        const int a = min(ia, ib), b = max(ia, ib);
        if (atom_table != nullptr) {
            const unsigned long long flat = (unsigned long long)a * (unsigned long long)n + (unsigned long long)b;
            if (atom_mask) {
                unsigned long long* slots = reinterpret_cast<unsigned long long*>(atom_table);
                const unsigned h = ht_find_or_insert(slots, atom_mask, flat + 1ull, hstats, 0);
                if (h != 0xffffffffu) atomicAdd(&slots[h], 1ull);
                else ht_park(ovf, hstats, flat + 1ull);
            } else {
                atomicAdd(&atom_table[flat], 1u);
            }
        }
        if (stamp == nullptr) return;
        resolve();                                   // consume the previous call's atomicMax
        const unsigned rlo = (unsigned)min(ra, rb), rhi = (unsigned)max(ra, rb);
        const unsigned slot32 = rlo * n_res_c + rhi;     // R_c^2 < 2^32 (checked on the host)
        const unsigned peers = __match_any_sync(__activemask(), slot32);
        if ((__ffs(peers) - 1) == (int)(threadIdx.x & 31)) {
            // the atom pairs of one residue pair come from the same few tasks: a small per-warp
            // cache of the pairs this warp has stamped in this frame (= this launch) filters most
            // repeats before they become atomics on the stamp table.  A tag is only ever written
            // by a lane that goes on to stamp the pair, so a matching tag is always safe to skip.
            unsigned int* tag = &rcache[slot32 & (LG_RCACHE - 1)];
            if (*(volatile unsigned int*)tag == slot32) return;
            *tag = slot32;
            unsigned where = slot32;                 // index into the stamp array
            if (res_mask) {
                where = ht_find_or_insert(reinterpret_cast<unsigned long long*>(res_table), res_mask,
                                          (unsigned long long)slot32 + 1ull, hstats, 1);
                if (where == 0xffffffffu) {
                    ht_park(ovf, hstats, (1ull << 63) | (((unsigned long long)slot32 + 1ull) << 24) |
                                             (unsigned long long)(seq & 0xffffffu));
                    return;
                }
            }
            pend_old = atomicMax(&stamp[(size_t)where], seq);
            pend_lo = rlo; pend_hi = rhi; pend_slot = where;
            pending = 1;
        }
    }
};

// The frames one launch works on: frame `fb = blockIdx.y` of the batch uses slice fb of every
// array (one frame at a time, batch of 1: the whole GPU on slice 0).
struct LargeBatch {
    const float* xyz;     // [frames][n][3], first frame of the batch
    const float* box;     // [frames][9] or nullptr
    int n;
    int maxc;             // cell capacity of one slice
    GridS* grid;          // [frames]
    int* cellcnt;         // [frames][maxc + 1]
    int* cellptr;         // [frames][maxc + 1]
    int* occ;             // [frames][maxc]
    int* misc;            // [frames][4]: n_occ, task counter, max |coordinate| (float bits)
    float4* wrapped;      // [frames][n]
    int* atomrank;        // [frames][n]
    float4* sorted;       // [frames][n]
    float2* sorted_kr;    // [frames][n]
};

// Sink of the BATCHED mode (mid-sized systems: many frames per launch, a few CTAs each).  Frames
// of a batch run concurrently, so the per-frame residue de-duplication cannot rely on stream
// order: every frame of the batch has its own residue-pair bitmap in global memory (R_c^2 bits,
// L2-resident for the systems this mode is used for) and the lane whose atomicOr sets a new bit
// counts / emits the residue pair -- the scheme of the fused kernel, in global memory.
struct BatchSink {
    // resident CTAs per SM of kl_pairs with this sink: the batched schedule is latency-bound and not
    // limited by table thrashing, more warps pay (measured 2 / 3 / 4 / 5 / 6: best at 4)
    static constexpr int MIN_CTAS = 4;
    unsigned int* atom_table;
    unsigned int* res_table;
    unsigned int* bitmap;           // [frames][bm_words], nullptr: residues not requested
    long long bm_words;
    unsigned long long* res_keys;
    long long res_cap;
    unsigned long long* counters;
    unsigned int n;
    unsigned int n_res_c;
    long long frame_id;             // id of frame 0 of the batch, then of this CTA's frame
    unsigned int* rcache;           // unused (interface of kl_pairs)
    int ra, rb;

    __device__ __forceinline__ void set_frame(int fb)
    {
        if (bitmap != nullptr) bitmap += (size_t)fb * (size_t)bm_words;
        frame_id += fb;
    }
    __device__ __forceinline__ void resolve() {}
    __device__ __forceinline__ void prepare(const PairCtx<LargeTraits>& cx, int ipos, int jpos, int, int)
    {
        ra = __float_as_int(__ldg(&cx.kr[ipos]).y);
        rb = __float_as_int(__ldg(&cx.kr[jpos]).y);
    }
    __device__ __forceinline__ void account(int ia, int ib)
    {
        const int a = min(ia, ib), b = max(ia, ib);
        if (atom_table != nullptr)
            atomicAdd(&atom_table[(unsigned long long)a * (unsigned long long)n + (unsigned long long)b], 1u);
        if (bitmap == nullptr) return;
        const unsigned rlo = (unsigned)min(ra, rb), rhi = (unsigned)max(ra, rb);
        const unsigned slot = rlo * n_res_c + rhi;              // R_c^2 < 2^32 (checked on the host)
        const unsigned bit = 1u << (slot & 31u);
        unsigned int* word = &bitmap[slot >> 5];
        if (*(volatile unsigned int*)word & bit) return;        // most repeats stop here
        if (atomicOr(word, bit) & bit) return;
        if (res_table != nullptr) atomicAdd(&res_table[slot], 1u);
        if (res_keys != nullptr) {
            const unsigned long long w = atomicAdd(&counters[1], 1ull);
            if ((long long)w < res_cap)
                res_keys[w] = ((unsigned long long)frame_id << 40) | ((unsigned long long)rlo << 20) |
                              (unsigned long long)rhi;
        }
    }
};

__global__ void __launch_bounds__(LG_BIN_THREADS) kl_prepare(LargeBatch lb, float cutoff)
{
    __shared__ float red[6 * 32];
    __shared__ GridS Gs;
    const int fb = blockIdx.y, n = lb.n;
    const float* __restrict__ x = lb.xyz + (size_t)fb * (size_t)n * 3;
    const float* __restrict__ box9 = lb.box ? lb.box + (size_t)fb * 9 : nullptr;
    int* __restrict__ cellcnt = lb.cellcnt + (size_t)fb * (size_t)(lb.maxc + 1);
    int* __restrict__ misc = lb.misc + (size_t)fb * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
    if (box9 == nullptr) {
        float l0 = 3.0e38f, l1 = 3.0e38f, l2 = 3.0e38f, h0 = -3.0e38f, h1 = -3.0e38f, h2 = -3.0e38f;
        for (int i = tid; i < n; i += LG_BIN_THREADS) {
            const float px = x[3 * (size_t)i], py = x[3 * (size_t)i + 1], pz = x[3 * (size_t)i + 2];
            l0 = fminf(l0, px); h0 = fmaxf(h0, px);
            l1 = fminf(l1, py); h1 = fmaxf(h1, py);
            l2 = fminf(l2, pz); h2 = fmaxf(h2, pz);
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            l0 = fminf(l0, __shfl_xor_sync(0xffffffffu, l0, d));
            l1 = fminf(l1, __shfl_xor_sync(0xffffffffu, l1, d));
            l2 = fminf(l2, __shfl_xor_sync(0xffffffffu, l2, d));
            h0 = fmaxf(h0, __shfl_xor_sync(0xffffffffu, h0, d));
            h1 = fmaxf(h1, __shfl_xor_sync(0xffffffffu, h1, d));
            h2 = fmaxf(h2, __shfl_xor_sync(0xffffffffu, h2, d));
        }
        if (lane == 0) {
            red[warp] = l0; red[32 + warp] = l1; red[64 + warp] = l2;
            red[96 + warp] = h0; red[128 + warp] = h1; red[160 + warp] = h2;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < LG_BIN_THREADS / 32; w++) {
                red[0] = fminf(red[0], red[w]); red[32] = fminf(red[32], red[32 + w]);
                red[64] = fminf(red[64], red[64 + w]); red[96] = fmaxf(red[96], red[96 + w]);
                red[128] = fmaxf(red[128], red[128 + w]); red[160] = fmaxf(red[160], red[160 + w]);
            }
            lo[0] = red[0]; lo[1] = red[32]; lo[2] = red[64];
            hi[0] = red[96]; hi[1] = red[128]; hi[2] = red[160];
        }
    }
    if (tid == 0) {
        make_grid(Gs, box9, (double)cutoff, lb.maxc, lo, hi);
        lb.grid[fb] = Gs;
        misc[0] = 0;
        misc[1] = 0;
        misc[2] = 0;
    }
    __syncthreads();
    const int ncell = Gs.ncell;
    for (int c = tid; c <= ncell; c += LG_BIN_THREADS) cellcnt[c] = 0;
}

__global__ void __launch_bounds__(256) kl_bin(LargeBatch lb)
{
    const int fb = blockIdx.y, n = lb.n;
    const float* __restrict__ x = lb.xyz + (size_t)fb * (size_t)n * 3;
    int* __restrict__ cellcnt = lb.cellcnt + (size_t)fb * (size_t)(lb.maxc + 1);
    float4* __restrict__ wrapped = lb.wrapped + (size_t)fb * (size_t)n;
    int* __restrict__ atomrank = lb.atomrank + (size_t)fb * (size_t)n;
    const GridS G = lb.grid[fb];
    float amax = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float px = __ldg(&x[3 * (size_t)i]), py = __ldg(&x[3 * (size_t)i + 1]), pz = __ldg(&x[3 * (size_t)i + 2]);
        amax = fmaxf(amax, fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz))));
        const int c = cell_wrap(G, px, py, pz);
        wrapped[i] = make_float4(px, py, pz, __int_as_float(c));
        atomrank[i] = atomicAdd(&cellcnt[c], 1);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
    // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(&lb.misc[(size_t)fb * 4 + 2], __float_as_int(amax));
}

__global__ void __launch_bounds__(LG_BIN_THREADS) kl_scan(LargeBatch lb)
{
    __shared__ int s_sum[32], s_occ[32];
    const int fb = blockIdx.y, n = lb.n;
    const GridS* __restrict__ grid = lb.grid + fb;
    const int* __restrict__ cellcnt = lb.cellcnt + (size_t)fb * (size_t)(lb.maxc + 1);
    int* __restrict__ cellptr = lb.cellptr + (size_t)fb * (size_t)(lb.maxc + 1);
    int* __restrict__ occ = lb.occ + (size_t)fb * (size_t)lb.maxc;
    int* __restrict__ misc = lb.misc + (size_t)fb * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncell = grid->ncell;
    const int cpt = (ncell + LG_BIN_THREADS - 1) / LG_BIN_THREADS;
    const int c0 = min(tid * cpt, ncell), c1 = min(c0 + cpt, ncell);
    int sum = 0, nocc = 0;
    for (int c = c0; c < c1; c++) {
        const int v = cellcnt[c];
        sum += v;
        nocc += (v != 0);
    }
    int isum = sum, iocc = nocc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, isum, d);
        const int b = __shfl_up_sync(0xffffffffu, iocc, d);
        if (lane >= d) { isum += a; iocc += b; }
    }
    if (lane == 31) { s_sum[warp] = isum; s_occ[warp] = iocc; }
    __syncthreads();
    if (warp == 0) {
        int a = lane < LG_BIN_THREADS / 32 ? s_sum[lane] : 0, b = lane < LG_BIN_THREADS / 32 ? s_occ[lane] : 0;
        int ia = a, ib = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ia, d);
            const int v = __shfl_up_sync(0xffffffffu, ib, d);
            if (lane >= d) { ia += u; ib += v; }
        }
        s_sum[lane] = ia - a;
        s_occ[lane] = ib - b;
        if (lane == 31) misc[0] = ib;      // lanes beyond the warp count add nothing
    }
    __syncthreads();
    int run = isum - sum + s_sum[warp];
    int ocur = iocc - nocc + s_occ[warp];
    const GridS& G = *grid;
    for (int c = c0; c < c1; c++) {
        const int v = cellcnt[c];
        cellptr[c] = run;
        if (v) {
            int cx, cy, cz;
            LargeTraits::decode(c, G, cx, cy, cz);
            occ[ocur++] = (int)((unsigned)c | (cell_interior(G, cx, cy, cz) ? LargeTraits::TASK_INTERIOR : 0u));
        }
        run += v;
    }
    if (tid == 0) cellptr[ncell] = n;
}

__global__ void __launch_bounds__(256) kl_scatter(LargeBatch lb, const int2* __restrict__ meta)
{
    const int fb = blockIdx.y, n = lb.n;
    const int* __restrict__ cellptr = lb.cellptr + (size_t)fb * (size_t)(lb.maxc + 1);
    const float4* __restrict__ wrapped = lb.wrapped + (size_t)fb * (size_t)n;
    const int* __restrict__ atomrank = lb.atomrank + (size_t)fb * (size_t)n;
    float4* __restrict__ sorted = lb.sorted + (size_t)fb * (size_t)n;
    float2* __restrict__ sorted_kr = lb.sorted_kr + (size_t)fb * (size_t)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 w = wrapped[i];
        const int pos = cellptr[__float_as_int(w.w)] + atomrank[i];
        const int2 mw = __ldg(&meta[i]);
        const unsigned bits = (unsigned)i | ((unsigned)mw.y & (3u << 30));
        sorted[pos] = make_float4(w.x, w.y, w.z, __uint_as_float(bits));
        // keys are < 2^24 (checked on the host): exact as floats
        sorted_kr[pos] = make_float2((float)mw.x, __int_as_float(mw.y & 0x3fffffff));
    }
}

// gridDim.x CTAs per frame, frame blockIdx.y of the batch; pp.xyz / pp.box point at the batch's
// first frame and pp.frame0 is its id
template <int FLAGS, typename Sink>
__global__ void __launch_bounds__(LG_PAIR_THREADS, Sink::MIN_CTAS) kl_pairs(PairParams pp, Sink sink_in, LargeBatch lb,
                                                                         float cutoff)
{
    __shared__ unsigned int cand_all[(LG_PAIR_THREADS / 32) * CT_CAND];
    __shared__ uint4 hitq_all[(LG_PAIR_THREADS / 32) * CT_QCAP];
    __shared__ float4 shift_s[27];
    __shared__ float thresh_s[4];
    __shared__ unsigned int rcache_all[(LG_PAIR_THREADS / 32) * LG_RCACHE];
    __shared__ GridS G;
    const int fb = blockIdx.y;
    const float* __restrict__ box9 = lb.box ? lb.box + (size_t)fb * 9 : nullptr;
    int* misc = lb.misc + (size_t)fb * 4;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 64) G = lb.grid[fb];
    __syncthreads();
    for (int t = threadIdx.x; t < (LG_PAIR_THREADS / 32) * LG_RCACHE; t += LG_PAIR_THREADS) rcache_all[t] = 0xffffffffu;
    if (threadIdx.x < 27) shift_s[threadIdx.x] = lattice_shift(box9, threadIdx.x);
    if (threadIdx.x == 32) {
        float lo, hi;
        frame_margins(box9, G.distinct, __int_as_float(misc[2]), pp.c2, cutoff, pp.boundary_ulps, lo, hi);
        thresh_s[2] = lo; thresh_s[3] = hi;
        if ((FLAGS & CT_BOUNDARY) && pp.boundary_ulps < 0) { lo = -1.0e30f; hi = 1.0e30f; }   // audit
        thresh_s[0] = lo; thresh_s[1] = hi;
    }
    __syncthreads();
    pp.c2_lo = thresh_s[0]; pp.c2_hi = thresh_s[1];
    pp.audit_lo = thresh_s[2]; pp.audit_hi = thresh_s[3];
    pp.f = fb;
    PairCtx<LargeTraits> cx;
    cx.sorted = lb.sorted + (size_t)fb * (size_t)lb.n;
    cx.kr = lb.sorted_kr + (size_t)fb * (size_t)lb.n;
    cx.cellptr = lb.cellptr + (size_t)fb * (size_t)(lb.maxc + 1);
    cx.cand = cand_all + warp * CT_CAND; cx.hitq = hitq_all + warp * CT_QCAP; cx.cellrole = nullptr;
    cx.shift = shift_s;
    unsigned long long n_tests = 0;
    Sink sink = sink_in;
    sink.set_frame(fb);
    sink.rcache = rcache_all + warp * LG_RCACHE;
    ct_pair_phase_dyn<LargeTraits, Sink, FLAGS, int>(pp, G, cx, sink, lb.occ + (size_t)fb * (size_t)lb.maxc, misc[0],
                                                     &misc[1], n_tests);
    sink.resolve();
    if (FLAGS & CT_COUNT) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_tests += __shfl_xor_sync(0xffffffffu, n_tests, d);
        if ((threadIdx.x & 31) == 0 && n_tests) atomicAdd(&pp.counters[3], n_tests);
    }
}

#define LG_CK(call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err = std::string(#call) + " failed: " + cudaGetErrorString(e_);         \
            return -2;                                                               \
        }                                                                            \
    } while (0)

// launch the pair kernel variant the job asks for
template <typename Sink>
inline void launch_pairs(const LargeJob& job, const PairParams& pp, const Sink& sink, const LargeBatch& lb,
                         dim3 grid, cudaStream_t stream)
{
    const bool emit = job.atom_keys != nullptr;
    if (job.boundary_ulps != 0)
        kl_pairs<CT_BOUNDARY | CT_ROLES, Sink><<<grid, LG_PAIR_THREADS, 0, stream>>>(pp, sink, lb, job.cutoff);
    else if (job.count_tests && !emit)
        kl_pairs<CT_COUNT | CT_ROLES, Sink><<<grid, LG_PAIR_THREADS, 0, stream>>>(pp, sink, lb, job.cutoff);
    else if (emit)
        kl_pairs<CT_EMIT | CT_ROLES, Sink><<<grid, LG_PAIR_THREADS, 0, stream>>>(pp, sink, lb, job.cutoff);
    else if (job.all_roles)
        kl_pairs<0, Sink><<<grid, LG_PAIR_THREADS, 0, stream>>>(pp, sink, lb, job.cutoff);
    else
        kl_pairs<CT_ROLES, Sink><<<grid, LG_PAIR_THREADS, 0, stream>>>(pp, sink, lb, job.cutoff);
}

inline PairParams large_pair_params(const LargeJob& job, const float* x, const float* box9, long long frame0)
{
    PairParams pp;
    pp.n = job.n; pp.n_ignored = job.n_ignored; pp.c2 = job.c2; pp.c2_lo = pp.c2_hi = job.c2;
    pp.boundary_ulps = job.boundary_ulps; pp.xyz = x; pp.box = box9; pp.f = 0; pp.frame0 = frame0;
    pp.atom_keys = job.atom_keys; pp.atom_cap = job.atom_cap; pp.counters = job.counters;
    pp.boundary = job.boundary; pp.boundary_cap = job.boundary_cap;
    return pp;
}

// BATCHED mode: mid-sized systems (a frame is far too little work for the whole GPU, and five
// launches per frame would dominate).  `frames` frames per launch, `ctas` CTAs of the pair kernel
// each; per-frame residue bitmaps instead of the stream-ordered frame stamps.  Returns 1 when
// the mode does not apply (hashed tables, bitmaps too large, a single frame).
#ifndef LG_BATCH_ATOMS_PER_CTA
#define LG_BATCH_ATOMS_PER_CTA 8192
#endif
constexpr int LG_BATCH_MAXC = 1 << 16;                  // cell capacity of one batch slice
constexpr size_t LG_BATCH_BYTES = (size_t)2048 << 20;   // scratch budget of the batched mode
constexpr int LG_BATCH_MAX_ATOMS = 1 << 18;             // mode 1: above this the one-frame-at-a-time mode is used

inline int large_run_batched(const LargeJob& job, LargeScratch& S, int sm_count, cudaStream_t stream,
                             std::string& err, int& launches)
{
    const int n = job.n;
    const bool want_res = (job.res_table != nullptr) || (job.res_keys != nullptr);
    if (job.batch_mode == 0 || job.atom_log2 > 0 || job.res_log2 > 0 || job.n_frames < 2) return 1;
    if (job.batch_mode == 1 && n > LG_BATCH_MAX_ATOMS) return 1;     // big frames fill the GPU on their own
    const size_t bm_words = want_res ? ((size_t)job.n_res_c * (size_t)job.n_res_c + 31) / 32 : 0;
    const size_t per_frame = 2 * ((size_t)n * 44 + (size_t)LG_BATCH_MAXC * 12 + 64 + sizeof(GridS) + bm_words * 4);
    const int total_ctas = sm_count * BatchSink::MIN_CTAS;
    // as many frames per launch as the scratch budget allows, at least one CTA (8 warps) per
    // LG_BATCH_ATOMS_PER_CTA atoms of a frame (measured: fewer, longer-lived CTAs per frame win),
    // and the CTAs the frame count leaves free are spread over the frames
    const int ctas_min = std::max(1, std::min(total_ctas, (n + LG_BATCH_ATOMS_PER_CTA - 1) / LG_BATCH_ATOMS_PER_CTA));
    long long frames = std::min<long long>(job.n_frames, std::max(1, total_ctas / ctas_min));
    frames = std::min<long long>(frames, (long long)(LG_BATCH_BYTES / per_frame));
    if (frames < 2) return 1;
    const int ctas = std::max<int>(ctas_min, (int)(total_ctas / frames));
    if (S.batch_frames < frames || S.batch_atoms < n || S.batch_bm_words < bm_words) {
        cudaFree(S.d_bitmaps[0]); cudaFree(S.d_bitmaps[1]);
        S.d_bitmaps[0] = S.d_bitmaps[1] = nullptr;
        S.batch_frames = 0;
        const size_t F = (size_t)frames;
        for (LargeBins& B : S.batch) {
            cudaFree(B.d_grid); cudaFree(B.d_cellcnt); cudaFree(B.d_cellptr); cudaFree(B.d_occ); cudaFree(B.d_misc);
            cudaFree(B.d_wrapped); cudaFree(B.d_atomrank); cudaFree(B.d_sorted); cudaFree(B.d_sorted_kr);
            B.d_grid = nullptr; B.d_cellcnt = B.d_cellptr = B.d_occ = B.d_misc = B.d_atomrank = nullptr;
            B.d_wrapped = B.d_sorted = nullptr; B.d_sorted_kr = nullptr;
            LG_CK(cudaMalloc(&B.d_grid, sizeof(GridS) * F));
            LG_CK(cudaMalloc(&B.d_cellcnt, sizeof(int) * F * (size_t)(LG_BATCH_MAXC + 1)));
            LG_CK(cudaMalloc(&B.d_cellptr, sizeof(int) * F * (size_t)(LG_BATCH_MAXC + 1)));
            LG_CK(cudaMalloc(&B.d_occ, sizeof(int) * F * (size_t)LG_BATCH_MAXC));
            LG_CK(cudaMalloc(&B.d_misc, sizeof(int) * F * 4));
            LG_CK(cudaMalloc(&B.d_wrapped, sizeof(float4) * F * (size_t)n));
            LG_CK(cudaMalloc(&B.d_atomrank, sizeof(int) * F * (size_t)n));
            LG_CK(cudaMalloc(&B.d_sorted, sizeof(float4) * F * (size_t)n));
            LG_CK(cudaMalloc(&B.d_sorted_kr, sizeof(float2) * F * (size_t)n));
            if (!B.binned) LG_CK(cudaEventCreateWithFlags(&B.binned, cudaEventDisableTiming));
            if (!B.paired) LG_CK(cudaEventCreateWithFlags(&B.paired, cudaEventDisableTiming));
            B.paired_recorded = false;
        }
        for (int k = 0; k < 2 && bm_words; k++) LG_CK(cudaMalloc(&S.d_bitmaps[k], sizeof(unsigned int) * F * bm_words));
        S.batch_frames = frames; S.batch_atoms = n; S.batch_bm_words = bm_words;
    }
    if (!S.aux_batch) {
        LG_CK(cudaEventCreateWithFlags(&S.entry_batch, cudaEventDisableTiming));
        LG_CK(cudaStreamCreateWithFlags(&S.aux_batch, cudaStreamNonBlocking));
    }
    const int atom_blocks = std::min((n + 255) / 256, sm_count * 8);
    // binning of batch k + 1 and the clearing of its bitmaps (aux stream, other buffer set) run
    // under the pair kernel of batch k
    LG_CK(cudaEventRecord(S.entry_batch, stream));         // the frames may still be in flight (H2D)
    LG_CK(cudaStreamWaitEvent(S.aux_batch, S.entry_batch, 0));
    int kb = 0;
    for (long long f0 = 0; f0 < job.n_frames; f0 += frames, kb++) {
        const unsigned nf = (unsigned)std::min<long long>(frames, job.n_frames - f0);
        LargeBins& B = S.batch[kb & 1];
        LargeBatch lb;
        lb.xyz = job.xyz + (size_t)f0 * (size_t)n * 3;
        lb.box = job.box ? job.box + (size_t)f0 * 9 : nullptr;
        lb.n = n; lb.maxc = LG_BATCH_MAXC;
        lb.grid = B.d_grid; lb.cellcnt = B.d_cellcnt; lb.cellptr = B.d_cellptr; lb.occ = B.d_occ; lb.misc = B.d_misc;
        lb.wrapped = B.d_wrapped; lb.atomrank = B.d_atomrank; lb.sorted = B.d_sorted; lb.sorted_kr = B.d_sorted_kr;
        if (B.paired_recorded) LG_CK(cudaStreamWaitEvent(S.aux_batch, B.paired, 0));   // the set is free again
        unsigned int* bitmaps = S.d_bitmaps[kb & 1];
        if (bm_words) LG_CK(cudaMemsetAsync(bitmaps, 0, sizeof(unsigned int) * (size_t)nf * bm_words, S.aux_batch));
        kl_prepare<<<dim3(1, nf), LG_BIN_THREADS, 0, S.aux_batch>>>(lb, job.cutoff);
        kl_bin<<<dim3(atom_blocks, nf), 256, 0, S.aux_batch>>>(lb);
        kl_scan<<<dim3(1, nf), LG_BIN_THREADS, 0, S.aux_batch>>>(lb);
        kl_scatter<<<dim3(atom_blocks, nf), 256, 0, S.aux_batch>>>(lb, job.meta);
        LG_CK(cudaEventRecord(B.binned, S.aux_batch));
        LG_CK(cudaStreamWaitEvent(stream, B.binned, 0));
        const PairParams pp = large_pair_params(job, lb.xyz, lb.box, job.frame0 + f0);
        BatchSink sink;
        sink.atom_table = job.atom_table; sink.res_table = job.res_table;
        sink.bitmap = want_res ? bitmaps : nullptr; sink.bm_words = (long long)bm_words;
        sink.res_keys = job.res_keys; sink.res_cap = job.res_cap; sink.counters = job.counters;
        sink.n = (unsigned)n; sink.n_res_c = (unsigned)job.n_res_c; sink.frame_id = job.frame0 + f0;
        sink.rcache = nullptr; sink.ra = sink.rb = 0;
        launch_pairs(job, pp, sink, lb, dim3(ctas, nf), stream);
        LG_CK(cudaEventRecord(B.paired, stream));
        B.paired_recorded = true;
        launches += 5;
    }
    LG_CK(cudaGetLastError());
    return 0;
}

inline int large_run(const LargeJob& job, LargeScratch& S, int sm_count, cudaStream_t stream,
                     std::string& err, int& launches)
{
    const int n = job.n;
    const bool want_res = (job.res_table != nullptr) || (job.res_keys != nullptr);
    if (want_res && (unsigned long long)job.n_res_c * (unsigned long long)job.n_res_c >= (1ull << 32)) {
        err = "more than 65535 used residues: residue table index exceeds 32 bits";
        return -4;
    }
    if (job.res_log2 > 0 && job.res_table == nullptr && job.res_keys != nullptr) {
        err = "per-frame residue keys need a residue table (dense or hashed) for the de-duplication";
        return -1;
    }
    {
        const int rc = large_run_batched(job, S, sm_count, stream, err, launches);
        if (rc <= 0) return rc;
    }
    // ONE FRAME AT A TIME: the whole GPU on a frame, residue de-duplication by frame stamps
    if (S.cap_atoms < n) {
        for (LargeBins& B : S.bins) {
            cudaFree(B.d_wrapped); cudaFree(B.d_atomrank); cudaFree(B.d_sorted); cudaFree(B.d_sorted_kr);
            B.d_atomrank = nullptr; B.d_sorted = nullptr; B.d_wrapped = nullptr; B.d_sorted_kr = nullptr;
            LG_CK(cudaMalloc(&B.d_sorted_kr, sizeof(float2) * (size_t)n));
            LG_CK(cudaMalloc(&B.d_wrapped, sizeof(float4) * (size_t)n));
            LG_CK(cudaMalloc(&B.d_atomrank, sizeof(int) * (size_t)n));
            LG_CK(cudaMalloc(&B.d_sorted, sizeof(float4) * (size_t)n));
        }
        S.cap_atoms = n;
    }
    if (!S.aux) {
        for (LargeBins& B : S.bins) {
            LG_CK(cudaMalloc(&B.d_grid, sizeof(GridS)));
            LG_CK(cudaMalloc(&B.d_cellcnt, sizeof(int) * (size_t)(LG_MAXC + 1)));
            LG_CK(cudaMalloc(&B.d_cellptr, sizeof(int) * (size_t)(LG_MAXC + 1)));
            LG_CK(cudaMalloc(&B.d_occ, sizeof(int) * (size_t)LG_MAXC));
            LG_CK(cudaMalloc(&B.d_misc, sizeof(int) * 4));
            LG_CK(cudaEventCreateWithFlags(&B.binned, cudaEventDisableTiming));
            LG_CK(cudaEventCreateWithFlags(&B.paired, cudaEventDisableTiming));
        }
        LG_CK(cudaEventCreateWithFlags(&S.entry, cudaEventDisableTiming));
        LG_CK(cudaStreamCreateWithFlags(&S.aux, cudaStreamNonBlocking));
    }
    const size_t need_stamp = job.res_log2 > 0 ? ((size_t)1 << job.res_log2)
                                               : (size_t)job.n_res_c * (size_t)job.n_res_c;
    if (want_res && S.stamp_elems < need_stamp) {
        cudaFree(S.d_stamp);
        S.d_stamp = nullptr;
        LG_CK(cudaMalloc(&S.d_stamp, sizeof(unsigned int) * need_stamp));
        LG_CK(cudaMemsetAsync(S.d_stamp, 0, sizeof(unsigned int) * need_stamp, stream));
        S.stamp_elems = need_stamp;
        S.seq = 0;
    }
    const int atom_blocks = std::min((n + 255) / 256, sm_count * 8);
    // Two-stage pipeline over the frames: bins[f & 1] is filled on the aux stream while kl_pairs
    // of the previous frame (other set) runs on the caller's stream.  kl_pairs launches stay in
    // stream order, which is what the frame-stamp de-duplication needs.
    LG_CK(cudaEventRecord(S.entry, stream));              // the frames may still be in flight (H2D)
    LG_CK(cudaStreamWaitEvent(S.aux, S.entry, 0));
    for (long long f = 0; f < job.n_frames; f++) {
        if (want_res) {
            if (S.seq == 0xffffffffu) {
                LG_CK(cudaMemsetAsync(S.d_stamp, 0, sizeof(unsigned int) * S.stamp_elems, stream));
                S.seq = 0;
            }
            S.seq++;
        }
        LargeBins& B = S.bins[f & 1];
        LargeBatch lb;
        lb.xyz = job.xyz + (size_t)f * (size_t)n * 3;
        lb.box = job.box ? job.box + (size_t)f * 9 : nullptr;
        lb.n = n; lb.maxc = LG_MAXC;
        lb.grid = B.d_grid; lb.cellcnt = B.d_cellcnt; lb.cellptr = B.d_cellptr; lb.occ = B.d_occ; lb.misc = B.d_misc;
        lb.wrapped = B.d_wrapped; lb.atomrank = B.d_atomrank; lb.sorted = B.d_sorted; lb.sorted_kr = B.d_sorted_kr;
        if (B.paired_recorded) LG_CK(cudaStreamWaitEvent(S.aux, B.paired, 0));   // the set is free again
        kl_prepare<<<1, LG_BIN_THREADS, 0, S.aux>>>(lb, job.cutoff);
        kl_bin<<<atom_blocks, 256, 0, S.aux>>>(lb);
        kl_scan<<<1, LG_BIN_THREADS, 0, S.aux>>>(lb);
        kl_scatter<<<atom_blocks, 256, 0, S.aux>>>(lb, job.meta);
        LG_CK(cudaEventRecord(B.binned, S.aux));
        LG_CK(cudaStreamWaitEvent(stream, B.binned, 0));
        const PairParams pp = large_pair_params(job, lb.xyz, lb.box, job.frame0 + f);
        StampSink sink;
        sink.meta = job.meta; sink.atom_table = job.atom_table;
        sink.stamp = want_res ? S.d_stamp : nullptr;
        sink.atom_mask = job.atom_log2 > 0 ? (1u << job.atom_log2) - 1u : 0u;
        sink.res_mask = job.res_log2 > 0 ? (1u << job.res_log2) - 1u : 0u;
        sink.hstats = job.hstats; sink.pend_slot = 0u;
        sink.ovf.list = job.ovf_list; sink.ovf.cap = job.ovf_cap;
        sink.res_table = job.res_table; sink.res_keys = job.res_keys; sink.res_cap = job.res_cap;
        sink.counters = job.counters; sink.seq = S.seq; sink.n = (unsigned)n;
        sink.n_res_c = (unsigned)job.n_res_c; sink.frame_id = job.frame0 + f;
        sink.pending = 0; sink.pend_old = 0u; sink.pend_lo = sink.pend_hi = 0u; sink.ra = sink.rb = 0;
        launch_pairs(job, pp, sink, lb, dim3(sm_count * LG_MIN_CTAS, 1), stream);
        LG_CK(cudaEventRecord(B.paired, stream));
        B.paired_recorded = true;
        launches += 5;
    }
    LG_CK(cudaGetLastError());
    return 0;
}

}  // namespace cmb
