// cmb_frames_small.cuh -- fused per-frame contact kernel for systems whose used
// atoms fit in one CTA's shared memory (n_used <= 6144).
//
// Replaces, per frame (reference file:line under /root/reference/contact_map/):
//   md.compute_neighborlist(used_traj, cutoff, frame)      contact_map.py:575
//   the Python set algebra of ContactObject._contact_map   contact_map.py:582-607
//   Counter.update / Counter += of ContactFrequency        contact_map.py:739-740
//   per-frame pair sets of _build_contacts                 contact_trajectory.py:32-44
//
// Persistent CTAs (grid = SMs x resident CTAs), frames dealt round-robin:
//   1. the frame's fp32 xyz (12*n bytes, contiguous in HBM) is pulled into shared
//      memory with ONE 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) signalled on an
//      mbarrier; the copy of the NEXT frame is issued as soon as the scatter pass
//      has consumed the buffer, so it overlaps the pair phase;
//   2. atoms are binned into >= cutoff-wide cells of the (orthorhombic / triclinic /
//      open) cell grid: u16 shared-memory histogram, block scan, scatter into a
//      cell-sorted float4 array {x, y, z, packed meta} (+ u16 atom index);
//   3. one warp per occupied cell: the candidate atoms of the half shell
//      (neighbour cells with id >= own) are held in registers (4 per lane), the
//      cell's own atoms are broadcast from shared memory, and every candidate pair
//      is adjudicated with MDTraj's exact float32 arithmetic on the RAW
//      coordinates (cmb_math.cuh);
//   4. hits that pass the role + residue-exclusion tests are (a) RED-added into the
//      dense uint32 atom-pair table, (b) OR'd into a per-frame residue-pair bitmap
//      in shared memory that is popcount-flushed into the residue table at frame
//      end (per-frame set semantics), and/or (c) warp-ballot compacted into 64-bit
//      (frame, a, b) keys for ContactTrajectory.
#pragma once
#include "cmb_cells.cuh"

namespace cmb {

constexpr int FS_THREADS = 512;
constexpr int FS_WARPS = FS_THREADS / 32;
constexpr int FS_APT = 12;                          // atoms per thread in the binning passes
constexpr int FS_MAX_ATOMS = FS_THREADS * FS_APT;   // 6144
constexpr int FS_HEADER_BYTES = 1024;
constexpr int FS_KSLOT = 4;                         // register-resident candidates per lane
constexpr int FS_CAND = 32 * FS_KSLOT;              // candidates per chunk
constexpr int FS_HITQ = 64;                         // per-warp queue of compacted hits

// packed per-atom meta word: ekey [0,17) | compact residue [17,30) | role [30,32)
constexpr unsigned FS_EKEY_BITS = 17;
constexpr unsigned FS_EKEY_MASK = (1u << FS_EKEY_BITS) - 1u;
constexpr unsigned FS_RES_MASK = (1u << 13) - 1u;

// kernel variants (template FLAGS)
constexpr int FS_EMIT = 1;        // emit (frame, a, b) keys
constexpr int FS_BOUNDARY = 2;    // report pairs within boundary_ulps of the cutoff
constexpr int FS_COUNT = 4;       // count executed pair tests
constexpr int FS_ROLES = 8;       // query / haystack are not "all atoms": test the roles

struct FrameParams {
    const float* xyz;            // [n_frames][n][3] fp32, device
    const float* box;            // [n_frames][9] fp32 row vectors, device, or nullptr
    long long n_frames;
    long long frame0;            // id of frame 0 of this call (for emitted keys)
    int n;                       // used atoms
    const unsigned int* meta32;  // [n] packed meta words
    float cutoff;
    float c2;                    // cutoff_f * cutoff_f
    float c2_hi;                 // c2 advanced by `boundary_ulps` float steps
    int boundary_ulps;
    int n_ignored;
    int n_res_c;                 // compact residue count (bitmap is n_res_c x n_res_c bits)
    int maxc;                    // cell capacity of the shared-memory arrays
    int bm_words;                // bitmap words
    int bm_in_smem;
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
    int raw_off, sorted_off, sidx_off, cell_off, occ_off, bm_off, cand_off, hitq_off, total;
};

__host__ __device__ inline int fs_align(int x, int a) { return (x + a - 1) / a * a; }

__host__ __device__ inline FsLayout fs_layout(int n, int maxc, int bm_words_smem)
{
    FsLayout L;
    int off = FS_HEADER_BYTES;
    L.raw_off = off;
    off += fs_align(12 * n + 32, 128);
    L.sorted_off = off;
    off += fs_align(16 * n, 128);
    L.sidx_off = off;
    off += fs_align(2 * n, 128);
    L.cell_off = off;
    off += fs_align(2 * (maxc + 2), 128);
    L.occ_off = off;
    off += fs_align(2 * (maxc < n ? maxc : n) + 2, 128);
    L.bm_off = off;
    off += fs_align(4 * bm_words_smem, 128);
    L.cand_off = off;
    off += FS_WARPS * FS_CAND * 2;
    L.hitq_off = off;
    off += FS_WARPS * FS_HITQ * 4;
    L.total = off;
    return L;
}

// everything the pair phase needs, in one place
struct FsPairCtx {
    const float4* sorted;            // [n] {x, y, z, meta32}
    const unsigned short* sidx;      // [n] atom index of each sorted position
    const unsigned short* cellptr;   // [ncell + 1]
    unsigned int* bitmap;            // residue-pair bitmap of this CTA
    unsigned int* hitq;              // this warp's queue of hits (ipos | jpos << 16)
    long long frame_id;
};

// Account one hit (sorted positions ipos < jpos): atom-pair table + residue-pair bitmap.
__device__ __forceinline__ void fs_on_hit(const FrameParams& p, const FsPairCtx& cx, int ipos, int jpos)
{
    const int ia = cx.sidx[ipos], ib = cx.sidx[jpos];
    const int a = min(ia, ib), b = max(ia, ib);
    if (p.atom_table != nullptr) atomicAdd(&p.atom_table[(size_t)a * (size_t)p.n + (size_t)b], 1u);
    const unsigned mi = __float_as_uint(cx.sorted[ipos].w), mj = __float_as_uint(cx.sorted[jpos].w);
    const unsigned ra = (mi >> FS_EKEY_BITS) & FS_RES_MASK, rb = (mj >> FS_EKEY_BITS) & FS_RES_MASK;
    const unsigned bit = min(ra, rb) * (unsigned)p.n_res_c + max(ra, rb);
    const unsigned msk = 1u << (bit & 31);
    unsigned int* wp = &cx.bitmap[bit >> 5];
    if (!(*(volatile unsigned int*)wp & msk)) atomicOr(wp, msk);
}

// Warp-ballot compaction: the hits of one slot iteration (about 5 % of the lanes) are appended
// to the warp's shared-memory queue and accounted 32 at a time with all lanes busy.
__device__ __forceinline__ void fs_queue_hits(const FrameParams& p, const FsPairCtx& cx, bool hit,
                                              int ipos, int jpos, int& qn)
{
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    if (hm == 0u) return;
    const int lane = threadIdx.x & 31;
    if (hit) cx.hitq[qn + __popc(hm & ((1u << lane) - 1u))] = (unsigned)ipos | ((unsigned)jpos << 16);
    qn += __popc(hm);
    if (qn >= 32) {
        __syncwarp();
        const unsigned e = cx.hitq[lane];
        const unsigned rest = (32 + lane < qn) ? cx.hitq[32 + lane] : 0u;
        __syncwarp();
        cx.hitq[lane] = rest;
        qn -= 32;
        fs_on_hit(p, cx, (int)(e & 0xffffu), (int)(e >> 16));
    }
}

__device__ __forceinline__ void fs_drain_hits(const FrameParams& p, const FsPairCtx& cx, int& qn)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if (lane < qn) {
        const unsigned e = cx.hitq[lane];
        fs_on_hit(p, cx, (int)(e & 0xffffu), (int)(e >> 16));
    }
    qn = 0;
    __syncwarp();
}

// all (own atom) x (NK register slots) tests of one candidate chunk
template <int MODE, int NK, int FLAGS>
__device__ __forceinline__ void fs_inner(const FrameParams& p, const NlBox& B, const FsPairCtx& cx,
                                         int cstart, int cend, const float (&xj)[FS_KSLOT],
                                         const float (&yj)[FS_KSLOT], const float (&zj)[FS_KSLOT],
                                         const unsigned (&mj)[FS_KSLOT], const int (&jpos)[FS_KSLOT], int& qn)
{
    const int lane = threadIdx.x & 31;
    for (int ipos = cstart; ipos < cend; ipos++) {
        const float4 pi = cx.sorted[ipos];
        const unsigned mi = __float_as_uint(pi.w);
#pragma unroll
        for (int k = 0; k < NK; k++) {
            // tested in the direction lower sorted position -> higher; the arithmetic is
            // antisymmetric for every pair that can be a hit ("direction")
            const float d2 = nl_d2<MODE>(pi.x, pi.y, pi.z, xj[k], yj[k], zj[k], B);
            int de = (int)(mi & FS_EKEY_MASK) - (int)(mj[k] & FS_EKEY_MASK);
            de = de < 0 ? -de : de;
            bool elig = (de > p.n_ignored) && (jpos[k] > ipos);
            if (FLAGS & FS_ROLES) {
                const unsigned ri = mi >> 30, rj = mj[k] >> 30;
                elig = elig && (((ri & (rj >> 1)) | (rj & (ri >> 1))) & 1u);
            }
            if (FLAGS & FS_BOUNDARY) {
                if (elig && d2 <= p.c2_hi) {
                    int diff = __float_as_int(d2) - __float_as_int(p.c2);
                    diff = diff < 0 ? -diff : diff;
                    if (diff <= p.boundary_ulps) {
                        const int ia = cx.sidx[ipos], ib = cx.sidx[jpos[k]];
                        const unsigned long long w = atomicAdd(&p.counters[2], 1ull);
                        if ((long long)w < p.boundary_cap)
                            p.boundary[w] = make_float4(__int_as_float((int)cx.frame_id),
                                                        __int_as_float(min(ia, ib)),
                                                        __int_as_float(max(ia, ib)), d2);
                    }
                }
            }
            const bool hit = elig && (d2 <= p.c2);
            if (FLAGS & FS_EMIT) {
                // warp-ballot compaction of this slot's hits into the key stream
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hm) {
                    const int leader = __ffs(hm) - 1;
                    unsigned long long wbase = 0;
                    if (lane == leader) wbase = atomicAdd(&p.counters[0], (unsigned long long)__popc(hm));
                    wbase = __shfl_sync(0xffffffffu, wbase, leader);
                    if (hit) {
                        const unsigned long long w = wbase + __popc(hm & ((1u << lane) - 1u));
                        const int ia = cx.sidx[ipos], ib = cx.sidx[jpos[k]];
                        if ((long long)w < p.atom_cap)
                            p.atom_keys[w] = ((unsigned long long)cx.frame_id << 40) |
                                             ((unsigned long long)min(ia, ib) << 20) |
                                             (unsigned long long)max(ia, ib);
                    }
                }
            }
            fs_queue_hits(p, cx, hit, ipos, jpos[k], qn);
        }
    }
}

// One warp adjudicates all candidate pairs of one occupied cell `c`.
//
// Candidates are the atoms of the stencil cells with id >= c (every unordered cell pair is
// visited once).  Cells are numbered x-fastest, so for an INTERIOR cell (no periodic wrap, no
// grid edge) they form five contiguous runs of the cell-sorted array:
//     (cx..cx+1, cy, cz), (cx-1..cx+1, cy+1, cz), (cx-1..cx+1, cy-1..cy+1, cz+1)
// and a candidate slot maps to its sorted position with four compares.  Edge cells take the
// generic route: stencil lanes, de-duplication, warp scan, per-warp candidate list.
template <int MODE, int FLAGS>
__device__ __forceinline__ void fs_cell_task(const FrameParams& p, const NlBox& B, const GridS& G,
                                             const FsPairCtx& cx, int c, unsigned short* cand,
                                             unsigned long long& n_tests, int& qn)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int cyz = (int)__umulhi((unsigned)c << 6, G.magic_x);
    const int cxx = c - cyz * G.nx;
    const int cz = (int)__umulhi((unsigned)cyz << 6, G.magic_y);
    const int cy = cyz - cz * G.ny;
    const int cstart = (int)cx.cellptr[c], cend = (int)cx.cellptr[c + 1];

    const bool interior = (cxx >= 1) && (cxx <= G.nx - 2) && (cy >= 1) && (cy <= G.ny - 2) &&
                          (cz >= 1) && (cz <= G.nz - 2);
    int M;
    // interior: run r covers slots [o_r, o_{r+1}) and sorted positions slot + a_r
    int o1 = 0, o2 = 0, o3 = 0, o4 = 0, a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    // generic: this lane's stencil cell
    int mystart = 0, mycnt = 0, off = 0;
    if (interior) {
        const int nxy = G.nx * G.ny;
        const int b1 = c + G.nx - 1, b2 = c + nxy - G.nx - 1, b3 = c + nxy - 1, b4 = c + nxy + G.nx - 1;
        const int s0 = cstart, e0 = (int)cx.cellptr[c + 2];
        const int s1 = (int)cx.cellptr[b1], e1 = (int)cx.cellptr[b1 + 3];
        const int s2 = (int)cx.cellptr[b2], e2 = (int)cx.cellptr[b2 + 3];
        const int s3 = (int)cx.cellptr[b3], e3 = (int)cx.cellptr[b3 + 3];
        const int s4 = (int)cx.cellptr[b4], e4 = (int)cx.cellptr[b4 + 3];
        o1 = e0 - s0; o2 = o1 + (e1 - s1); o3 = o2 + (e2 - s2); o4 = o3 + (e3 - s3);
        M = o4 + (e4 - s4);
        a0 = s0; a1 = s1 - o1; a2 = s2 - o2; a3 = s3 - o3; a4 = s4 - o4;
    } else {
        int myD = -1;
        if (lane < 27) {
            const int l9 = lane / 9, l3 = (lane - l9 * 9) / 3;
            int x = cxx + (lane - l9 * 9 - l3 * 3) - 1;
            int y = cy + l3 - 1;
            int z = cz + l9 - 1;
            if (G.periodic) {
                x = x < 0 ? x + G.nx : (x >= G.nx ? x - G.nx : x);
                y = y < 0 ? y + G.ny : (y >= G.ny ? y - G.ny : y);
                z = z < 0 ? z + G.nz : (z >= G.nz ? z - G.nz : z);
            }
            const bool ok = (x >= 0 && x < G.nx && y >= 0 && y < G.ny && z >= 0 && z < G.nz);
            if (ok) myD = (z * G.ny + y) * G.nx + x;
        }
        bool keep = (myD >= c);
        if (!G.distinct) {   // grids with < 3 cells along a dimension: drop duplicate stencil cells
            const unsigned same = __match_any_sync(FULL, myD);
            keep = keep && ((__ffs(same) - 1) == lane);
        }
        if (keep) {
            mystart = (int)cx.cellptr[myD];
            mycnt = (int)cx.cellptr[myD + 1] - mystart;
        }
        int incl = mycnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        M = __shfl_sync(FULL, incl, 31);
        off = incl - mycnt;
    }

    for (int base = 0; base < M; base += FS_CAND) {
        if (!interior) {
            __syncwarp();
            // expand this lane's stencil cell into the per-warp candidate list
            const int r0 = max(0, base - off), r1 = min(mycnt, base + FS_CAND - off);
            for (int r = r0; r < r1; r++) cand[off + r - base] = (unsigned short)(mystart + r);
            __syncwarp();
        }
        float xj[FS_KSLOT], yj[FS_KSLOT], zj[FS_KSLOT];
        unsigned mj[FS_KSLOT];
        int jpos[FS_KSLOT];
#pragma unroll
        for (int k = 0; k < FS_KSLOT; k++) {
            const int s = base + k * 32 + lane;
            jpos[k] = -1; mj[k] = 0u; xj[k] = yj[k] = zj[k] = 0.f;
            if (s < M) {
                int pos;
                if (interior) {
                    int adj = a0;
                    adj = s >= o1 ? a1 : adj;
                    adj = s >= o2 ? a2 : adj;
                    adj = s >= o3 ? a3 : adj;
                    adj = s >= o4 ? a4 : adj;
                    pos = s + adj;
                } else {
                    pos = cand[s - base];
                }
                const float4 q = cx.sorted[pos];
                xj[k] = q.x; yj[k] = q.y; zj[k] = q.z;
                mj[k] = __float_as_uint(q.w);
                jpos[k] = pos;
            }
        }
        const int left = M - base;
        if (FLAGS & FS_COUNT) {
            if (lane == 0) n_tests += (unsigned long long)(cend - cstart) * (unsigned long long)min(left, FS_CAND);
        }
        if (left > 96) fs_inner<MODE, 4, FLAGS>(p, B, cx, cstart, cend, xj, yj, zj, mj, jpos, qn);
        else if (left > 64) fs_inner<MODE, 3, FLAGS>(p, B, cx, cstart, cend, xj, yj, zj, mj, jpos, qn);
        else if (left > 32) fs_inner<MODE, 2, FLAGS>(p, B, cx, cstart, cend, xj, yj, zj, mj, jpos, qn);
        else fs_inner<MODE, 1, FLAGS>(p, B, cx, cstart, cend, xj, yj, zj, mj, jpos, qn);
    }
}

template <int MODE, int FLAGS>
__device__ __forceinline__ void fs_pair_phase(const FrameParams& p, const NlBox& B, const GridS& G,
                                           const FsPairCtx& cx, const unsigned short* occ, int n_occ,
                                           int* task_counter, unsigned short* cand,
                                           unsigned long long& n_tests)
{
    const int lane = threadIdx.x & 31;
    int qn = 0;      // hits waiting in this warp's queue
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(task_counter, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_occ) break;
        fs_cell_task<MODE, FLAGS>(p, B, G, cx, (int)occ[k], cand, n_tests, qn);
    }
    fs_drain_hits(p, cx, qn);
}

#ifndef FS_MIN_CTAS
#define FS_MIN_CTAS 2
#endif

template <int FLAGS>
__global__ void __launch_bounds__(FS_THREADS, FS_MIN_CTAS) k_frames_small(const FrameParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n = p.n;
    const FsLayout L = fs_layout(n, p.maxc, p.bm_in_smem ? p.bm_words : 0);

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
    int* task_counter = reinterpret_cast<int*>(smem + 8);
    int* n_occ_s = reinterpret_cast<int*>(smem + 12);
    float* red = reinterpret_cast<float*>(smem + 64);                    // 6 * 32 floats
    unsigned int* scan_s = reinterpret_cast<unsigned int*>(smem + 64);   // overlays `red`
    GridS* Gs = reinterpret_cast<GridS*>(smem + 64 + 6 * 32 * 4);
    unsigned char* raw = smem + L.raw_off;
    float4* sorted = reinterpret_cast<float4*>(smem + L.sorted_off);
    unsigned short* sidx = reinterpret_cast<unsigned short*>(smem + L.sidx_off);
    unsigned int* cell32 = reinterpret_cast<unsigned int*>(smem + L.cell_off);
    unsigned short* cell16 = reinterpret_cast<unsigned short*>(smem + L.cell_off);
    unsigned short* occ = reinterpret_cast<unsigned short*>(smem + L.occ_off);
    unsigned int* bitmap = p.bm_in_smem ? reinterpret_cast<unsigned int*>(smem + L.bm_off)
                                        : p.g_bitmap + (size_t)blockIdx.x * (size_t)p.bm_words;
    unsigned short* cand = reinterpret_cast<unsigned short*>(smem + L.cand_off) + warp * FS_CAND;
    unsigned int* hitq = reinterpret_cast<unsigned int*>(smem + L.hitq_off) + warp * FS_HITQ;

    if (tid == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int w = tid; w < p.bm_words; w += FS_THREADS) bitmap[w] = 0u;
    __syncthreads();

    unsigned long long n_tests = 0;
    uint32_t parity = 0;
    const long long stride = gridDim.x;
    long long f = blockIdx.x;

    auto issue_load = [&](long long frame) {
        const unsigned long long addr = (unsigned long long)(p.xyz + (size_t)frame * (size_t)n * 3);
        const unsigned long long a0 = addr & ~15ull;
        const uint32_t lead = (uint32_t)(addr - a0);
        const uint32_t bytes = (lead + 12u * (uint32_t)n + 15u) & ~15u;
        mbar_expect_tx(mbar, bytes);
        tma_bulk_g2s(raw, reinterpret_cast<const void*>(a0), bytes, mbar);
    };
    if (tid == 0 && f < p.n_frames) issue_load(f);

    for (; f < p.n_frames; f += stride) {
        const float* box9 = p.box ? p.box + (size_t)f * 9 : nullptr;
        const NlBox B = make_nlbox(box9);
        const unsigned long long addr = (unsigned long long)(p.xyz + (size_t)f * (size_t)n * 3);
        const float* xr = reinterpret_cast<const float*>(raw + (addr & 15ull));

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
        if (tid == 0) {
            make_grid(*Gs, box9, (double)p.cutoff, p.maxc, lo, hi);
            *task_counter = 0;
        }
        __syncthreads();
        const GridS G = *Gs;

        // ---- pass A: histogram (u16 counters packed in u32 words) -----------------
        for (int w = tid; w < (G.ncell + 3) / 2; w += FS_THREADS) cell32[w] = 0u;
        __syncthreads();
        int my_cell[FS_APT], my_rank[FS_APT];
#pragma unroll
        for (int k = 0; k < FS_APT; k++) {
            const int i = tid + k * FS_THREADS;
            my_cell[k] = -1; my_rank[k] = 0;
            if (i < n) {
                const int c = cell_of(G, xr[3 * i], xr[3 * i + 1], xr[3 * i + 2]);
                const unsigned sh = (c & 1) * 16;
                const unsigned old = atomicAdd(&cell32[c >> 1], 1u << sh);
                my_cell[k] = c;
                my_rank[k] = (int)((old >> sh) & 0xffffu);
            }
        }
        __syncthreads();

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
            for (int c = c0; c < c1; c++) {
                const unsigned v = cell16[c];
                cell16[c] = (unsigned short)run;
                if (v) occ[ocur++] = (unsigned short)c;
                run += v;
            }
            if (tid == 0) cell16[G.ncell] = (unsigned short)n;
        }
        __syncthreads();

        // ---- pass C: scatter into the cell-sorted arrays ------------------------------
#pragma unroll
        for (int k = 0; k < FS_APT; k++) {
            const int i = tid + k * FS_THREADS;
            if (i < n) {
                const int pos = (int)cell16[my_cell[k]] + my_rank[k];
                sorted[pos] = make_float4(xr[3 * i], xr[3 * i + 1], xr[3 * i + 2],
                                          __uint_as_float(__ldg(&p.meta32[i])));
                sidx[pos] = (unsigned short)i;
            }
        }
        __syncthreads();
        // raw buffer is free: prefetch this CTA's next frame under the pair phase
        if (tid == 0 && f + stride < p.n_frames) {
            fence_proxy_async();
            issue_load(f + stride);
        }

        // ---- pass D: pair phase, one warp per occupied cell ---------------------------
        FsPairCtx cx;
        cx.sorted = sorted; cx.sidx = sidx; cx.cellptr = cell16; cx.bitmap = bitmap; cx.hitq = hitq;
        cx.frame_id = p.frame0 + f;
        const int n_occ = *n_occ_s;
        if (B.mode == BOX_ORTHO)
            fs_pair_phase<BOX_ORTHO, FLAGS>(p, B, G, cx, occ, n_occ, task_counter, cand, n_tests);
        else if (B.mode == BOX_TRICLINIC)
            fs_pair_phase<BOX_TRICLINIC, FLAGS>(p, B, G, cx, occ, n_occ, task_counter, cand, n_tests);
        else
            fs_pair_phase<BOX_NONE, FLAGS>(p, B, G, cx, occ, n_occ, task_counter, cand, n_tests);
        __syncthreads();

        // ---- pass E: flush the per-frame residue bitmap ------------------------------
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
                            p.res_keys[wbase + o] = ((unsigned long long)cx.frame_id << 40) |
                                                    ((unsigned long long)rlo << 20) | (unsigned long long)rhi;
                        o++;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (FLAGS & FS_COUNT) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) n_tests += __shfl_xor_sync(0xffffffffu, n_tests, d);
        if (lane == 0 && n_tests) atomicAdd(&p.counters[3], n_tests);
    }
}

}  // namespace cmb
