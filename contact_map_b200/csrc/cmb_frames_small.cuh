// cmb_frames_small.cuh -- fused per-frame contact kernel for systems whose used
// atoms fit in one CTA's shared memory (n_used <= 6144).
//
// Replaces, per frame (reference file:line under /root/reference/contact_map/):
//   md.compute_neighborlist(used_traj, cutoff, frame)      contact_map.py:575
//   the Python set algebra of ContactObject._contact_map   contact_map.py:582-607
//   Counter.update / Counter += of ContactFrequency        contact_map.py:739-740
//   per-frame pair sets of _build_contacts                 contact_trajectory.py:32-44
//
// Persistent CTAs (grid = SMs x resident CTAs: two of 512 threads, or one of 1024 threads when
// the system's arrays do not leave room for two), frames dealt round-robin:
//   1. the frame's fp32 xyz (12*n bytes, contiguous in HBM) is pulled into shared
//      memory with ONE 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) signalled on an
//      mbarrier; the copy of the NEXT frame is issued as soon as the scatter pass
//      has consumed the buffer, so it overlaps the pair phase;
//   2. atoms are binned into >= cutoff-wide cells of the (orthorhombic / triclinic /
//      open) cell grid: u16 shared-memory histogram, block scan, scatter into a
//      cell-sorted float4 array {wrapped x, y, z, atom index | role | exclusion key};
//   3. one warp per task (an occupied cell, or two sparse x-adjacent cells): the candidate atoms of the half shell
//      (neighbour cells with id >= own) are held in registers (up to 4 per lane), the
//      cell's own atoms are broadcast from shared memory, every candidate pair is
//      SCREENED with an FMA distance on the wrapped coordinates, and the few pairs the
//      per-frame error bound cannot decide are adjudicated with MDTraj's exact float32
//      arithmetic on the RAW coordinates (cmb_cells.cuh, cmb_math.cuh);
//   4. hits that pass the role + residue-exclusion tests are (a) RED-added into the
//      dense uint32 atom-pair table, (b) OR'd into a per-frame residue-pair bitmap
//      in shared memory that is popcount-flushed into the residue table at frame
//      end (per-frame set semantics), and/or (c) warp-ballot compacted into 64-bit
//      (frame, a, b) keys for ContactTrajectory.
#pragma once
#include "cmb_cells.cuh"

namespace cmb {

#ifndef FS_THREADS_CFG
#define FS_THREADS_CFG 512
#endif
constexpr int FS_THREADS = FS_THREADS_CFG;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_APT = 6144 / FS_THREADS;           // atoms per thread in the binning passes
constexpr int FS_MAX_ATOMS = FS_THREADS * FS_APT;   // 6144
constexpr int FS_HEADER_BYTES = 2048;
constexpr int FS_CAND = CT_CAND;
constexpr int FS_QCAP = CT_QCAP;
// header layout (bytes): mbarrier 0, task counter 8, task count 12, thresholds 16/20, grid cache 24..68,
// per-warp max|coordinate| 96, reduction scratch 256, grid 1024, shifts 1536
constexpr int FS_H_AMAX = 96, FS_H_RED = 256, FS_H_GRID = 1024, FS_H_SHIFT = 1536;

// packed per-atom meta word: ekey [0,16) | compact residue [16,29) | role [30,32)
constexpr unsigned FS_EKEY_BITS = 16;
constexpr unsigned FS_EKEY_MASK = (1u << FS_EKEY_BITS) - 1u;
constexpr unsigned FS_RES_MASK = (1u << 13) - 1u;

// kernel variants (template FLAGS), see cmb_cells.cuh
constexpr int FS_EMIT = CT_EMIT;
constexpr int FS_BOUNDARY = CT_BOUNDARY;
constexpr int FS_COUNT = CT_COUNT;
constexpr int FS_ROLES = CT_ROLES;

struct FrameParams {
    const float* xyz;            // [n_frames][n][3] fp32, device
    const float* box;            // [n_frames][9] fp32 row vectors, device, or nullptr
    long long n_frames;
    long long frame0;            // id of frame 0 of this call (for emitted keys)
    int n;                       // used atoms
    const unsigned int* meta32;  // [n] packed meta words
    float cutoff;
    float c2;                    // cutoff_f * cutoff_f
    int boundary_ulps;
    int n_ignored;
    int n_res_c;                 // compact residue count (bitmap is n_res_c x n_res_c bits)
    int maxc;                    // cell capacity of the shared-memory arrays
    int bm_words;                // bitmap words
    int bm_in_smem;
    int pair_tasks;              // merge sparsely populated x-adjacent interior cells into one task
    unsigned int* g_bitmap;      // [gridDim.x][bm_words] scratch when the bitmap is not in smem
    unsigned int* atom_table;    // [n][n] uint32 (entry [lo][hi]) or nullptr
    unsigned int* res_table;     // [n_res_c][n_res_c] uint32 or nullptr
    unsigned long long* atom_keys;   // emitted (frame<<40 | lo<<20 | hi) or nullptr
    long long atom_cap;
    unsigned long long* res_keys;    // emitted (frame<<40 | rlo<<20 | rhi) or nullptr
    long long res_cap;
    unsigned long long* counters;    // [0] atom keys, [1] residue keys, [2] boundary records, [3] pair tests
    float4* boundary;            // records {frame, lo, hi, d2} (ints bit-cast) or nullptr
    long long boundary_cap;
    // fall back of the pair-list path (cmb_verlet.cuh): only the frames frame_list[0 .. *frame_count)
    // are processed (both device side: no host round trip); nullptr: all n_frames frames
    const int* frame_list;
    const int* frame_count;
};

// ----- mbarrier / TMA bulk helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----- shared-memory layout ------------------------------------------------------
struct FsLayout {
    int raw_off, sorted_off, cell_off, occ_off, role_off, bm_off, cand_off, hitq_off, total;
};

__host__ __device__ inline int fs_align(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline FsLayout fs_layout(int n, int maxc, int bm_words_smem, int warps = FS_WARPS)
{
    FsLayout L;
    int off = FS_HEADER_BYTES;
    L.raw_off = off;
    off += fs_align(12 * n + 32, 128);
    L.sorted_off = off;
    off += fs_align(16 * n, 128);
    L.cell_off = off;
    off += fs_align(2 * (maxc + 2), 128);
    L.occ_off = off;
    off += fs_align(2 * (maxc < n ? maxc : n) + 2, 128);
    L.role_off = off;
    off += fs_align(4 * ((maxc + 15) / 16), 128);      // 2 role bits per cell
    L.bm_off = off;
    off += fs_align(4 * bm_words_smem, 128);
    L.cand_off = off;
    off += warps * FS_CAND * 2;
    L.hitq_off = off;
    off += warps * FS_QCAP * 8;
    L.total = off;
    return L;
}

// ----- traits + hit sink of the fused path ------------------------------------------
struct SmallTraits {
    typedef unsigned short pos_t;
    typedef unsigned short cand_t;                  // sorted position [0,13) | wrap flags [13,16)
    typedef uint2 hit_t;          // {own base [0,13) | jpos [13,26) | shift id [26,31), own-atom mask}
    // task word: cell [0,14) | interior | pair
    static constexpr unsigned TASK_CMASK = 0x3fffu, TASK_INTERIOR = 0x4000u, TASK_PAIR = 0x8000u,
                              TASK_EMPTY = 0xffffu;
    static constexpr int TASK_BATCH = 4;          // ~550 list entries per frame for 16 warps
    static constexpr bool CELL_ROLES = true;      // per-cell role bits let role-restricted runs drop tasks early
    static __device__ __forceinline__ cand_t cand(int pos, int wrap)
    {
        return (cand_t)((unsigned)pos | ((unsigned)wrap << 13));
    }
    static __device__ __forceinline__ int cand_pos(cand_t e) { return (int)(e & 0x1fffu); }
    static __device__ __forceinline__ int cand_wrap(cand_t e) { return (int)(e >> 13); }
    static __device__ __forceinline__ unsigned jp(int jpos, int id)
    {
        return ((unsigned)jpos << 13) | ((unsigned)id << 26);
    }
    static __device__ __forceinline__ hit_t entry(int ibase, unsigned jp, unsigned mask)
    {
        return make_uint2((unsigned)ibase | jp, mask);
    }
    static __device__ __forceinline__ hit_t with_mask(hit_t e, unsigned mask) { return make_uint2(e.x, mask); }
    static __device__ __forceinline__ int e_ibase(hit_t e) { return (int)(e.x & 0x1fffu); }
    static __device__ __forceinline__ int e_jpos(hit_t e) { return (int)((e.x >> 13) & 0x1fffu); }
    static __device__ __forceinline__ int e_id(hit_t e) { return (int)(e.x >> 26); }
    static __device__ __forceinline__ unsigned e_mask(hit_t e) { return e.y; }
    // task counter in shared memory; plain ATOMS (the compiler's warp-aggregated form costs 16
    // issue slots for a single active lane)
    static __device__ __forceinline__ int next_tasks(int* counter, int count)
    {
        int old;
        asm volatile("atom.shared.add.u32 %0, [%1], %2;"
                     : "=r"(old)
                     : "r"((unsigned)__cvta_generic_to_shared(counter)), "r"(count)
                     : "memory");
        return old;
    }
    // sorted[].w bits: atom index [0,13) | role [14,16) | exclusion key [16,32)
    static __device__ __forceinline__ unsigned bits(int atom, unsigned role, unsigned ekey)
    {
        return (unsigned)atom | (role << 14) | (ekey << 16);
    }
    static __device__ __forceinline__ int atom(unsigned bits) { return (int)(bits & 0x1fffu); }
    static __device__ __forceinline__ unsigned role(unsigned bits) { return (bits >> 14) & 3u; }
    static __device__ __forceinline__ float ekey_f(const PairCtx<SmallTraits>&, int, unsigned bits)
    {
        return (float)(bits >> 16);
    }
    static __device__ __forceinline__ void decode(int c, const GridS& G, int& cx, int& cy, int& cz)
    {
        const int cyz = (int)__umulhi((unsigned)c << 6, G.magic_x);     // c < 2^16
        cx = c - cyz * G.nx;
        cz = (int)__umulhi((unsigned)cyz << 6, G.magic_y);
        cy = cyz - cz * G.ny;
    }
};

// Residue exclusion + accounting of one hit: RED into the atom-pair table, test-then-OR into the
// per-frame residue-pair bitmap (shared or global), flushed at frame end.
struct BitmapSink {
    const unsigned int* meta32;
    unsigned int* atom_table;
    unsigned int* bitmap;
    unsigned int* res_inline;  // residue table when a new bit is counted at once (nullptr: the
                               // end-of-frame sweep counts / emits the set bits)
    unsigned n;
    unsigned n_res_c;
    unsigned ra, rb;          // compact residues of the pair handed to prepare()
    __device__ __forceinline__ void prepare(const PairCtx<SmallTraits>&, int, int, int ia, int ib)
    {
        ra = (__ldg(&meta32[ia]) >> FS_EKEY_BITS) & FS_RES_MASK;
        rb = (__ldg(&meta32[ib]) >> FS_EKEY_BITS) & FS_RES_MASK;
    }
    __device__ __forceinline__ void account(int ia, int ib)
    {
        const int a = min(ia, ib), b = max(ia, ib);
        if (atom_table != nullptr) atomicAdd(&atom_table[(size_t)a * (size_t)n + (size_t)b], 1u);
        const unsigned bit = min(ra, rb) * n_res_c + max(ra, rb);
        const unsigned msk = 1u << (bit & 31);
        unsigned int* wp = &bitmap[bit >> 5];
        if (!(*(volatile unsigned int*)wp & msk)) {
            const unsigned old = atomicOr(wp, msk);
            // first hit of this residue pair in this frame (per-frame set semantics)
            if (res_inline != nullptr && !(old & msk)) atomicAdd(&res_inline[bit], 1u);
        }
    }
};

#ifndef FS_MIN_CTAS
#define FS_MIN_CTAS 2
#endif

// THREADS = 512: two CTAs per SM when the system's arrays leave room for that (<= ~2 800 atoms);
// THREADS = 1024: one CTA per SM with the same number of warps for the systems that do not
template <int FLAGS, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS > 512 ? 1 : FS_MIN_CTAS) k_frames_small(const FrameParams p)
{
    // shadow the namespace-level defaults with this instantiation's values
    constexpr int FS_THREADS = THREADS, FS_WARPS = THREADS / 32, FS_APT = FS_MAX_ATOMS / THREADS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const FsLayout L = fs_layout(n, p.maxc, p.bm_in_smem ? p.bm_words : 0, FS_WARPS);

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
    int* task_counter = reinterpret_cast<int*>(smem + 8);
    int* n_occ_s = reinterpret_cast<int*>(smem + 12);
    float* thresh_s = reinterpret_cast<float*>(smem + 16);                 // {c2_lo, c2_hi} of the frame
    unsigned int* box_valid_s = reinterpret_cast<unsigned int*>(smem + 24);   // grid cache: box it was built for
    unsigned int* box_saved_s = reinterpret_cast<unsigned int*>(smem + 32);   // [9]
    float* amax_s = reinterpret_cast<float*>(smem + FS_H_AMAX);
    float* red = reinterpret_cast<float*>(smem + FS_H_RED);                // 6 * 32 floats
    unsigned int* scan_s = reinterpret_cast<unsigned int*>(smem + FS_H_RED);   // overlays `red`
    GridS* Gs = reinterpret_cast<GridS*>(smem + FS_H_GRID);
    float4* shift_s = reinterpret_cast<float4*>(smem + FS_H_SHIFT);        // [27]
    static_assert(sizeof(GridS) <= FS_H_SHIFT - FS_H_GRID, "grid does not fit its header slot");
    unsigned char* raw = smem + L.raw_off;
    float4* sorted = reinterpret_cast<float4*>(smem + L.sorted_off);
    unsigned int* cell32 = reinterpret_cast<unsigned int*>(smem + L.cell_off);
    unsigned short* cell16 = reinterpret_cast<unsigned short*>(smem + L.cell_off);
    unsigned short* occ = reinterpret_cast<unsigned short*>(smem + L.occ_off);
    unsigned int* cellrole = reinterpret_cast<unsigned int*>(smem + L.role_off);
    unsigned int* bitmap = p.bm_in_smem ? reinterpret_cast<unsigned int*>(smem + L.bm_off)
                                        : p.g_bitmap + (size_t)blockIdx.x * (size_t)p.bm_words;
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + L.cand_off) + warp * FS_CAND;
    uint2* hitq = reinterpret_cast<uint2*>(smem + L.hitq_off) + warp * FS_QCAP;

    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *box_valid_s = 0u;
    }
    for (int w = tid; w < p.bm_words; w += FS_THREADS) bitmap[w] = 0u;
    __syncthreads();

    unsigned long long n_tests = 0;
    uint32_t parity = 0;
    const long long stride = gridDim.x;
    // the frames this CTA works on: li-th entries of the frame list (identity without a list)
    const long long n_list = p.frame_list != nullptr ? (long long)*p.frame_count : p.n_frames;
    long long li = blockIdx.x;
    auto frame_of = [&](long long k) { return p.frame_list != nullptr ? (long long)p.frame_list[k] : k; };
    long long f = li < n_list ? frame_of(li) : 0;

    auto issue_load = [&](long long frame) {
        const unsigned long long addr = (unsigned long long)(p.xyz + (size_t)frame * (size_t)n * 3);
        const unsigned long long a0 = addr & ~15ull;
        const uint32_t lead = (uint32_t)(addr - a0);
        const uint32_t bytes = (lead + 12u * (uint32_t)n + 15u) & ~15u;
        mbar_expect_tx(mbar, bytes);
        tma_bulk_g2s(raw, reinterpret_cast<const void*>(a0), bytes, mbar);
    };
    if (tid == 0 && li < n_list) issue_load(f);

    // without residue keys the residue table is updated at the first hit of a pair and the
    // bitmap is only cleared between frames (together with the histogram, before the grid barrier)
    const bool res_at_once = p.res_table != nullptr && !((FLAGS & FS_EMIT) && p.res_keys != nullptr);
    bool first_frame = true;
    for (; li < n_list; li += stride) {
        f = frame_of(li);
        const float* box9 = p.box ? p.box + (size_t)f * 9 : nullptr;
        const unsigned long long addr = (unsigned long long)(p.xyz + (size_t)f * (size_t)n * 3);
        float* xr = reinterpret_cast<float*>(raw + (addr & 15ull));

        mbar_wait(mbar, parity);
        parity ^= 1u;

        // ---- open boundaries: bounding box of the frame ---------------------------
        float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
        if (box9 == nullptr) {
            float l0 = 3.0e38f, l1 = 3.0e38f, l2 = 3.0e38f, h0 = -3.0e38f, h1 = -3.0e38f, h2 = -3.0e38f;
            for (int i = tid; i < n; i += FS_THREADS) {
                const float x = xr[3 * i], y = xr[3 * i + 1], z = xr[3 * i + 2];
                l0 = fminf(l0, x); h0 = fmaxf(h0, x);
                l1 = fminf(l1, y); h1 = fmaxf(h1, y);
                l2 = fminf(l2, z); h2 = fmaxf(h2, z);
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
                for (int w = 1; w < FS_WARPS; w++) {
                    red[0] = fminf(red[0], red[w]); red[32] = fminf(red[32], red[32 + w]);
                    red[64] = fminf(red[64], red[64 + w]); red[96] = fmaxf(red[96], red[96 + w]);
                    red[128] = fmaxf(red[128], red[128 + w]); red[160] = fmaxf(red[160], red[160 + w]);
                }
                lo[0] = red[0]; lo[1] = red[32]; lo[2] = red[64];
                hi[0] = red[96]; hi[1] = red[128]; hi[2] = red[160];
            }
        }
        // the cell arrays and the residue bitmap of the previous frame are dead: clear them here, so
        // that the barrier below is the only one before the histogram
        for (int w = tid; w < (p.maxc + 3) / 2; w += FS_THREADS) cell32[w] = 0u;
        if (FLAGS & FS_ROLES)
            for (int w = tid; w < (p.maxc + 15) / 16; w += FS_THREADS) cellrole[w] = 0u;
        if (res_at_once && !first_frame)
            for (int w = tid; w < p.bm_words; w += FS_THREADS) bitmap[w] = 0u;
        first_frame = false;
        if (tid == 0) {
            // the grid only depends on the box: constant-volume trajectories build it once
            bool same = box9 != nullptr && *box_valid_s != 0;
            if (same)
                for (int k = 0; k < 9; k++) same = same && (__float_as_uint(box9[k]) == box_saved_s[k]);
            if (!same) {
                make_grid(*Gs, box9, (double)p.cutoff, p.maxc, lo, hi);
                *box_valid_s = box9 != nullptr;
                if (box9 != nullptr)
                    for (int k = 0; k < 9; k++) box_saved_s[k] = __float_as_uint(box9[k]);
            }
            *task_counter = 0;
        }
        __syncthreads();
        const GridS& G = *Gs;      // constant until the next frame (shared memory)

        // ---- pass A: histogram (u16 counters packed in u32 words); on periodic grids the raw
        //      buffer is overwritten with the wrapped positions ----------------------------
        unsigned my_cr[FS_APT];              // cell | rank << 16 of this thread's atoms
        float amax = 0.f;
#pragma unroll
        for (int k = 0; k < FS_APT; k++) {
            const int i = tid + k * FS_THREADS;
            my_cr[k] = 0u;
            if (i < n) {
                float x = xr[3 * i], y = xr[3 * i + 1], z = xr[3 * i + 2];
                amax = fmaxf(amax, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
                const int c = cell_wrap(G, x, y, z);
                if (G.periodic) { xr[3 * i] = x; xr[3 * i + 1] = y; xr[3 * i + 2] = z; }
                const unsigned sh = (c & 1) * 16;
                const unsigned old = atomicAdd(&cell32[c >> 1], 1u << sh);
                my_cr[k] = (unsigned)c | (((old >> sh) & 0xffffu) << 16);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d));
        if (lane == 0) amax_s[warp] = amax;
        __syncthreads();

        // screening thresholds and lattice shifts of the frame (idle lanes of the scan below)
        if (tid == 32) {
            float A = 0.f;
            for (int w = 0; w < FS_WARPS; w++) A = fmaxf(A, amax_s[w]);
            float c2_lo, c2_hi;
            frame_margins(box9, G.distinct, A, p.c2, p.cutoff, p.boundary_ulps, c2_lo, c2_hi);
            if ((FLAGS & FS_BOUNDARY) && p.boundary_ulps < 0) { c2_lo = -1.0e30f; c2_hi = 1.0e30f; }   // audit
            thresh_s[0] = c2_lo; thresh_s[1] = c2_hi;
        }
        if (tid >= 64 && tid < 64 + 27) shift_s[tid - 64] = lattice_shift(box9, tid - 64);

        // ---- pass B: exclusive scan over cells (+ occupied-cell list) --------------
        {
            const int cpt = (G.ncell + FS_THREADS - 1) / FS_THREADS;
            const int c0 = tid * cpt;
            const int c1 = min(c0 + cpt, G.ncell);
            unsigned sum = 0, nocc = 0;
            for (int c = c0; c < c1; c++) {
                const unsigned v = cell16[c];
                sum += v;
                nocc += (v != 0);
            }
            unsigned packed = sum | (nocc << 16);
            unsigned incl = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            if (lane == 31) scan_s[warp] = incl;     // `red` was consumed before the barriers above
            __syncthreads();
            if (warp == 0) {
                unsigned v = lane < FS_WARPS ? scan_s[lane] : 0u;
                unsigned iv = v;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned u = __shfl_up_sync(0xffffffffu, iv, d);
                    if (lane >= d) iv += u;
                }
                if (lane < FS_WARPS) scan_s[lane] = iv - v;
                if (lane == FS_WARPS - 1) *n_occ_s = (int)(iv >> 16);
            }
            __syncthreads();
            const unsigned excl = incl - packed + scan_s[warp];
            unsigned run = excl & 0xffffu, ocur = excl >> 16;
            // task list: one task per occupied cell, or per two x-adjacent sparsely populated
            // interior cells of this thread's range.  (The pair merge makes the list shorter than
            // the occupied-cell count the scan reserved: unused entries are marked empty.)
            int cxx, cy, cz;
            SmallTraits::decode(c0, G, cxx, cy, cz);
            unsigned prev = 0u;           // count of the previous cell if it can still take a partner
            const unsigned oend = ocur + nocc;
            for (int c = c0; c < c1; c++) {
                const unsigned v = cell16[c];
                cell16[c] = (unsigned short)run;
                run += v;
                const bool inter = cell_interior(G, cxx, cy, cz);
                if (v) {
                    if (prev && inter && prev + v <= (unsigned)CT_PAIR_MAX_OWN && p.pair_tasks) {
                        occ[ocur - 1] |= (unsigned short)SmallTraits::TASK_PAIR;
                        prev = 0u;
                    } else {
                        occ[ocur++] = (unsigned short)((unsigned)c | (inter ? SmallTraits::TASK_INTERIOR : 0u));
                        prev = inter ? v : 0u;
                    }
                } else {
                    prev = 0u;
                }
                if (++cxx == G.nx) { cxx = 0; if (++cy == G.ny) { cy = 0; cz++; } }
            }
            while (ocur < oend) occ[ocur++] = (unsigned short)0xffffu;      // empty task
            if (tid == 0) cell16[G.ncell] = (unsigned short)n;
        }
        __syncthreads();

        // ---- pass C: scatter into the cell-sorted array -------------------------------
#pragma unroll
        for (int k = 0; k < FS_APT; k++) {
            const int i = tid + k * FS_THREADS;
            if (i < n) {
                const int pos = (int)cell16[my_cr[k] & 0xffffu] + (int)(my_cr[k] >> 16);
                const unsigned mw = __ldg(&p.meta32[i]);
                sorted[pos] = make_float4(xr[3 * i], xr[3 * i + 1], xr[3 * i + 2],
                                          __uint_as_float(SmallTraits::bits(i, mw >> 30, mw & FS_EKEY_MASK)));
                if (FLAGS & FS_ROLES) {
                    const unsigned c = my_cr[k] & 0xffffu;
                    atomicOr(&cellrole[c >> 4], (mw >> 30) << ((c & 15u) * 2u));
                }
            }
        }
        __syncthreads();
        // raw buffer is free: prefetch this CTA's next frame under the pair phase
        if (tid == 0 && li + stride < n_list) {
            fence_proxy_async();
            issue_load(frame_of(li + stride));
        }

        // ---- pass D: pair phase, one warp per occupied cell ---------------------------
        const long long frame_id = p.frame0 + f;
        PairParams pp;
        pp.n = n; pp.n_ignored = p.n_ignored; pp.c2 = p.c2; pp.c2_lo = thresh_s[0]; pp.c2_hi = thresh_s[1];
        pp.boundary_ulps = p.boundary_ulps; pp.xyz = p.xyz; pp.box = p.box; pp.f = (int)f; pp.frame0 = p.frame0;
        pp.audit_lo = pp.audit_hi = p.c2;
        if ((FLAGS & FS_BOUNDARY) && p.boundary_ulps < 0) {
            // audit: the thresholds the production variants would use for this frame
            float A = 0.f;
            for (int w = 0; w < FS_WARPS; w++) A = fmaxf(A, amax_s[w]);
            frame_margins(box9, G.distinct, A, p.c2, p.cutoff, 0, pp.audit_lo, pp.audit_hi);
        }
        pp.atom_keys = p.atom_keys; pp.atom_cap = p.atom_cap; pp.counters = p.counters;
        pp.boundary = p.boundary; pp.boundary_cap = p.boundary_cap;
        PairCtx<SmallTraits> cx;
        cx.sorted = sorted; cx.cellptr = cell16; cx.cand = cand; cx.hitq = hitq; cx.cellrole = cellrole;
        cx.shift = shift_s; cx.kr = nullptr;
        BitmapSink sink{p.meta32, p.atom_table, bitmap, res_at_once ? p.res_table : nullptr, (unsigned)n,
                        (unsigned)p.n_res_c, 0u, 0u};
        ct_pair_phase_dyn<SmallTraits, BitmapSink, FLAGS, unsigned short>(pp, G, cx, sink, occ, *n_occ_s,
                                                                          task_counter, n_tests);
        __syncthreads();

        // ---- pass E: sweep the per-frame residue bitmap (only when its bits were not counted at
        //      once: residue keys requested, or no residue table) ---------------------------
        if (!res_at_once)
        for (int w = tid; w < p.bm_words; w += FS_THREADS) {
            unsigned v = bitmap[w];
            if (v) {
                bitmap[w] = 0u;
                if (p.res_table != nullptr) {
                    unsigned t = v;
                    while (t) {
                        const int b = __ffs(t) - 1;
                        t &= t - 1;
                        atomicAdd(&p.res_table[(size_t)w * 32 + b], 1u);
                    }
                }
                if ((FLAGS & FS_EMIT) && p.res_keys != nullptr) {
                    const unsigned long long wbase = atomicAdd(&p.counters[1], (unsigned long long)__popc(v));
                    unsigned t = v;
                    unsigned long long o = 0;
                    while (t) {
                        const int b = __ffs(t) - 1;
                        t &= t - 1;
                        const unsigned bit = (unsigned)w * 32u + (unsigned)b;
                        const unsigned rlo = bit / (unsigned)p.n_res_c, rhi = bit % (unsigned)p.n_res_c;
                        if ((long long)(wbase + o) < p.res_cap)
                            p.res_keys[wbase + o] = ((unsigned long long)frame_id << 40) |
                                                    ((unsigned long long)rlo << 20) | (unsigned long long)rhi;
                        o++;
                    }
                }
            }
        }
        if (!res_at_once) __syncthreads();
    }
    if (FLAGS & FS_COUNT) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_tests += __shfl_xor_sync(0xffffffffu, n_tests, d);
        if (lane == 0 && n_tests) atomicAdd(&p.counters[3], n_tests);
    }
}

}  // namespace cmb
