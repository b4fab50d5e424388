// cmb_cells.cuh -- cell grid + per-cell pair adjudication shared by the fused
// shared-memory kernel (cmb_frames_small.cuh) and the global-memory cell path
// (cmb_frames_large.cuh).
//
// Replaces, per frame (reference file:line under /root/reference/contact_map/):
//   md.compute_neighborlist(used_traj, cutoff, frame)      contact_map.py:575
//   the Python set algebra of ContactObject._contact_map   contact_map.py:582-607
// Closed form of that set algebra (validated against the unmodified reference,
// tests/golden): unordered pair {a,b} is a contact of the frame iff
//     d2(a,b) <= c2  and  not excl(a,b)  and  ((a in Q and b in H) or (b in Q and a in H))
// with excl(a,b) <=> same chain and |res_a - res_b| <= n_ignored
// (residue_neighborhood, contact_map.py:41-60).
#pragma once
#include "cmb_math.cuh"

namespace cmb {

struct GridS {
    double m[9];     // fractional transform: s_k = sum_r (p_r - o_r) * m[r*3+k]
    double o[3];
    int nx, ny, nz, ncell;
    int periodic;
    // c / nx == __umulhi(c << 6, magic_x) for c < 2^16 (ceil(2^26 / nx)); same for ny
    unsigned magic_x, magic_y;
    int distinct;    // the 27 stencil cells of any cell are pairwise distinct (no dedup needed)
};

// ----- per-frame cell grid (one thread) ------------------------------------------
__device__ inline void fit_cells(int& nx, int& ny, int& nz, int maxc)
{
    if (nx < 1) nx = 1;
    if (ny < 1) ny = 1;
    if (nz < 1) nz = 1;
    if (nx > 1023) nx = 1023;
    if (ny > 1023) ny = 1023;
    if (nz > 1023) nz = 1023;
    // coarser cells stay conservative: scale the grid down until it fits the capacity
    for (int round = 0; round < 6 && (long long)nx * ny * nz > (long long)maxc; round++) {
        const int dims = (nx > 1) + (ny > 1) + (nz > 1);
        const double f = pow((double)maxc / ((double)nx * ny * nz), 1.0 / (double)(dims > 0 ? dims : 1));
        if (nx > 1) nx = max(1, (int)floor(nx * f));
        if (ny > 1) ny = max(1, (int)floor(ny * f));
        if (nz > 1) nz = max(1, (int)floor(nz * f));
    }
    while ((long long)nx * ny * nz > (long long)maxc) {
        if (nx >= ny && nx >= nz) nx--;
        else if (ny >= nz) ny--;
        else nz--;
    }
}

__device__ inline void make_grid(GridS& G, const float* box9, double cutoff, int maxc,
                                    const float lo[3], const float hi[3])
{
    const double cw = cutoff * 1.0005 + 1e-7;     // conservative minimum cell width
    if (box9 != nullptr) {
        double h[9];
        for (int k = 0; k < 9; k++) h[k] = (double)box9[k];
        const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) +
                           h[2] * (h[3] * h[7] - h[4] * h[6]);
        const double id = 1.0 / det;
        G.m[0] = (h[4] * h[8] - h[5] * h[7]) * id;
        G.m[1] = (h[2] * h[7] - h[1] * h[8]) * id;
        G.m[2] = (h[1] * h[5] - h[2] * h[4]) * id;
        G.m[3] = (h[5] * h[6] - h[3] * h[8]) * id;
        G.m[4] = (h[0] * h[8] - h[2] * h[6]) * id;
        G.m[5] = (h[2] * h[3] - h[0] * h[5]) * id;
        G.m[6] = (h[3] * h[7] - h[4] * h[6]) * id;
        G.m[7] = (h[1] * h[6] - h[0] * h[7]) * id;
        G.m[8] = (h[0] * h[4] - h[1] * h[3]) * id;
        G.o[0] = G.o[1] = G.o[2] = 0.0;
        int nc[3];
        for (int k = 0; k < 3; k++) {
            const double nrm = sqrt(G.m[k] * G.m[k] + G.m[3 + k] * G.m[3 + k] + G.m[6 + k] * G.m[6 + k]);
            const double width = 1.0 / nrm;       // perpendicular height of the cell along k
            double q = floor(width / cw);
            if (!(q >= 1.0)) q = 1.0;             // also catches NaN (singular box)
            if (q > 1023.0) q = 1023.0;
            nc[k] = (int)q;
        }
        G.nx = nc[0]; G.ny = nc[1]; G.nz = nc[2];
        G.periodic = 1;
    } else {
        int nc[3];
        for (int k = 0; k < 9; k++) G.m[k] = 0.0;
        for (int k = 0; k < 3; k++) {
            const double ext = (double)hi[k] - (double)lo[k];
            G.o[k] = (double)lo[k];
            double q = floor(ext / cw);
            if (!(q >= 1.0)) q = 1.0;
            if (q > 1023.0) q = 1023.0;
            nc[k] = (int)q;
            G.m[k * 3 + k] = ext > 0.0 ? 1.0 / ext : 0.0;
        }
        G.nx = nc[0]; G.ny = nc[1]; G.nz = nc[2];
        G.periodic = 0;
    }
    fit_cells(G.nx, G.ny, G.nz, maxc);
    G.ncell = G.nx * G.ny * G.nz;
    G.magic_x = ((1u << 26) + (unsigned)G.nx - 1u) / (unsigned)G.nx;
    G.magic_y = ((1u << 26) + (unsigned)G.ny - 1u) / (unsigned)G.ny;
    G.distinct = !G.periodic || (G.nx >= 3 && G.ny >= 3 && G.nz >= 3);
}

__device__ __forceinline__ int cell_of(const GridS& G, float x, float y, float z)
{
    const double px = (double)x - G.o[0], py = (double)y - G.o[1], pz = (double)z - G.o[2];
    double sx = px * G.m[0] + py * G.m[3] + pz * G.m[6];
    double sy = px * G.m[1] + py * G.m[4] + pz * G.m[7];
    double sz = px * G.m[2] + py * G.m[5] + pz * G.m[8];
    if (G.periodic) {
        sx -= floor(sx); sy -= floor(sy); sz -= floor(sz);
    }
    int cx = (int)(sx * (double)G.nx);
    int cy = (int)(sy * (double)G.ny);
    int cz = (int)(sz * (double)G.nz);
    cx = cx < 0 ? 0 : (cx >= G.nx ? G.nx - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= G.ny ? G.ny - 1 : cy);
    cz = cz < 0 ? 0 : (cz >= G.nz ? G.nz - 1 : cz);
    return (cz * G.ny + cy) * G.nx + cx;
}


// =================================================================================
// Pair phase shared by both paths.
//
// The frame's atoms are cell-sorted into `sorted[pos] = {x, y, z, ekey}` where `ekey` is the
// exclusion key AS A FLOAT (exact: keys are < 2^24), so the residue-exclusion test is one FADD and
// one FSETP on |.|; `sidx[pos]` is the atom index, through which the rare consumers (hits, role
// test) reach the per-atom meta words.  A `Traits` type fixes the integer widths of the two paths:
//   pos_t       element type of cellptr / sidx / the per-warp candidate list
//   role()      role bits (bit0 = query, bit1 = haystack) of an atom
//   decode()    cell id -> (cx, cy, cz)
//   hit_t       queue entry of one hit, with pack()/ipos()/jpos()
// and a `Sink` object accounts one hit (atom table, residue de-duplication).
// =================================================================================
constexpr int CT_KSLOT = 4;                 // register-resident candidates per lane
constexpr int CT_CAND = 32 * CT_KSLOT;      // candidates per chunk
constexpr int CT_HITQ = 64;                 // per-warp queue of compacted hits

// kernel variants (template FLAGS)
constexpr int CT_EMIT = 1;        // emit (frame, a, b) keys
constexpr int CT_BOUNDARY = 2;    // report pairs within boundary_ulps of the cutoff
constexpr int CT_COUNT = 4;       // count executed pair tests
constexpr int CT_ROLES = 8;       // query / haystack are not "all atoms": test the roles

// hot-loop parameters (everything a cell task needs besides the geometry)
struct PairParams {
    int n;                           // used atoms
    float n_ignored_f;               // n_ignored as a float
    const void* meta;                // per-atom meta words (layout known to Traits / Sink)
    float c2;                        // cutoff_f * cutoff_f
    float c2_hi;                     // c2 advanced by boundary_ulps float steps
    int boundary_ulps;
    unsigned int* atom_table;        // [n][n] uint32 (entry [lo][hi]) or nullptr
    unsigned long long* atom_keys;   // emitted (frame<<40 | lo<<20 | hi) or nullptr
    long long atom_cap;
    unsigned long long* counters;    // [0] atom keys, [1] residue keys, [2] boundary records, [3] pair tests
    float4* boundary;                // records {frame, lo, hi, d2} (ints bit-cast) or nullptr
    long long boundary_cap;
    long long frame_id;
};

template <typename Traits>
struct PairCtx {
    const float4* sorted;                       // [n] {x, y, z, word}
    const typename Traits::pos_t* sidx;         // [n] atom index of each sorted position
    const typename Traits::pos_t* cellptr;      // [ncell + 1]
    typename Traits::pos_t* cand;               // this warp's candidate list [CT_CAND]
    typename Traits::hit_t* hitq;               // this warp's hit queue [CT_HITQ]
};

// Warp-ballot compaction: the hits of one slot iteration (a few % of the lanes) are appended to
// the warp's queue and accounted 32 at a time with all lanes busy.
template <typename Traits, typename Sink>
__device__ __forceinline__ void ct_queue_hits(const PairCtx<Traits>& cx, Sink& sink, bool hit, int ipos,
                                              int jpos, int& qn)
{
    const unsigned hm = __ballot_sync(0xffffffffu, hit);
    if (hm == 0u) return;
    const int lane = threadIdx.x & 31;
    if (hit) cx.hitq[qn + __popc(hm & ((1u << lane) - 1u))] = Traits::pack(ipos, jpos);
    qn += __popc(hm);
    if (qn >= 32) {
        __syncwarp();
        const typename Traits::hit_t e = cx.hitq[lane];
        typename Traits::hit_t rest = Traits::pack(0, 0);
        if (32 + lane < qn) rest = cx.hitq[32 + lane];
        __syncwarp();
        cx.hitq[lane] = rest;
        qn -= 32;
        sink.account(Traits::ipos(e), Traits::jpos(e));
    }
}

template <typename Traits, typename Sink>
__device__ __forceinline__ void ct_drain_hits(const PairCtx<Traits>& cx, Sink& sink, int& qn)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    if (lane < qn) {
        const typename Traits::hit_t e = cx.hitq[lane];
        sink.account(Traits::ipos(e), Traits::jpos(e));
    }
    qn = 0;
    __syncwarp();
}

// all (own atom) x (NK register slots) tests of one candidate chunk
template <typename Traits, typename Sink, int MODE, int NK, int FLAGS>
__device__ __forceinline__ void ct_inner(const PairParams& p, const NlBox& B, const PairCtx<Traits>& cx,
                                         Sink& sink, int cstart, int cend, const float (&xj)[CT_KSLOT],
                                         const float (&yj)[CT_KSLOT], const float (&zj)[CT_KSLOT],
                                         const float (&ej)[CT_KSLOT], const unsigned (&rj)[CT_KSLOT],
                                         const int (&jpos)[CT_KSLOT], int& qn)
{
    const int lane = threadIdx.x & 31;
    for (int ipos = cstart; ipos < cend; ipos++) {
        const float4 pi = cx.sorted[ipos];
        unsigned ri = 3u;
        if (FLAGS & CT_ROLES) ri = Traits::role(p.meta, (int)cx.sidx[ipos]);
#pragma unroll
        for (int k = 0; k < NK; k++) {
            // tested in the direction lower sorted position -> higher; the arithmetic is
            // antisymmetric for every pair that can be a hit ("direction")
            const float d2 = nl_d2<MODE>(pi.x, pi.y, pi.z, xj[k], yj[k], zj[k], B);
            // |ekey_i - ekey_j| > n_ignored  (exact small integers held as floats)
            bool elig = (fabsf(__fsub_rn(pi.w, ej[k])) > p.n_ignored_f) && (jpos[k] > ipos);
            if (FLAGS & CT_ROLES)
                elig = elig && (((ri & (rj[k] >> 1)) | (rj[k] & (ri >> 1))) & 1u);
            if (FLAGS & CT_BOUNDARY) {
                if (elig && d2 <= p.c2_hi) {
                    int diff = __float_as_int(d2) - __float_as_int(p.c2);
                    diff = diff < 0 ? -diff : diff;
                    if (diff <= p.boundary_ulps) {
                        const int ia = (int)cx.sidx[ipos], ib = (int)cx.sidx[jpos[k]];
                        const unsigned long long w = atomicAdd(&p.counters[2], 1ull);
                        if ((long long)w < p.boundary_cap)
                            p.boundary[w] = make_float4(__int_as_float((int)p.frame_id),
                                                        __int_as_float(min(ia, ib)),
                                                        __int_as_float(max(ia, ib)), d2);
                    }
                }
            }
            const bool hit = elig && (d2 <= p.c2);
            if (FLAGS & CT_EMIT) {
                // warp-ballot compaction of this slot's hits into the key stream
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                if (hm) {
                    const int leader = __ffs(hm) - 1;
                    unsigned long long wbase = 0;
                    if (lane == leader) wbase = atomicAdd(&p.counters[0], (unsigned long long)__popc(hm));
                    wbase = __shfl_sync(0xffffffffu, wbase, leader);
                    if (hit) {
                        const unsigned long long w = wbase + __popc(hm & ((1u << lane) - 1u));
                        const int ia = (int)cx.sidx[ipos], ib = (int)cx.sidx[jpos[k]];
                        if ((long long)w < p.atom_cap)
                            p.atom_keys[w] = ((unsigned long long)p.frame_id << 40) |
                                             ((unsigned long long)min(ia, ib) << 20) |
                                             (unsigned long long)max(ia, ib);
                    }
                }
            }
            ct_queue_hits<Traits, Sink>(cx, sink, hit, ipos, jpos[k], qn);
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
// generic route: 27 stencil lanes, de-duplication (grids with < 3 cells along a dimension),
// warp scan, per-warp candidate list.  Sorted-position order (jpos > ipos) removes self pairs
// and double visits inside the own cell.
template <typename Traits, typename Sink, int MODE, int FLAGS>
__device__ __forceinline__ void ct_cell_task(const PairParams& p, const NlBox& B, const GridS& G,
                                             const PairCtx<Traits>& cx, Sink& sink, int c,
                                             unsigned long long& n_tests, int& qn)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    int cxx, cy, cz;
    Traits::decode(c, G, cxx, cy, cz);
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
        if (!G.distinct) {
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

    for (int base = 0; base < M; base += CT_CAND) {
        if (!interior) {
            __syncwarp();
            // expand this lane's stencil cell into the per-warp candidate list
            const int r0 = max(0, base - off), r1 = min(mycnt, base + CT_CAND - off);
            for (int r = r0; r < r1; r++) cx.cand[off + r - base] = (typename Traits::pos_t)(mystart + r);
            __syncwarp();
        }
        float xj[CT_KSLOT], yj[CT_KSLOT], zj[CT_KSLOT], ej[CT_KSLOT];
        unsigned rj[CT_KSLOT];
        int jpos[CT_KSLOT];
#pragma unroll
        for (int k = 0; k < CT_KSLOT; k++) {
            const int s = base + k * 32 + lane;
            jpos[k] = -1; rj[k] = 3u; xj[k] = yj[k] = zj[k] = ej[k] = 0.f;
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
                    pos = (int)cx.cand[s - base];
                }
                const float4 q = cx.sorted[pos];
                xj[k] = q.x; yj[k] = q.y; zj[k] = q.z; ej[k] = q.w;
                if (FLAGS & CT_ROLES) rj[k] = Traits::role(p.meta, (int)cx.sidx[pos]);
                jpos[k] = pos;
            }
        }
        const int left = M - base;
        if (FLAGS & CT_COUNT) {
            if (lane == 0) n_tests += (unsigned long long)(cend - cstart) * (unsigned long long)min(left, CT_CAND);
        }
        if (left > 96) ct_inner<Traits, Sink, MODE, 4, FLAGS>(p, B, cx, sink, cstart, cend, xj, yj, zj, ej, rj, jpos, qn);
        else if (left > 64) ct_inner<Traits, Sink, MODE, 3, FLAGS>(p, B, cx, sink, cstart, cend, xj, yj, zj, ej, rj, jpos, qn);
        else if (left > 32) ct_inner<Traits, Sink, MODE, 2, FLAGS>(p, B, cx, sink, cstart, cend, xj, yj, zj, ej, rj, jpos, qn);
        else ct_inner<Traits, Sink, MODE, 1, FLAGS>(p, B, cx, sink, cstart, cend, xj, yj, zj, ej, rj, jpos, qn);
    }
}

// Warps pull occupied cells from a task counter until the frame is exhausted.
template <typename Traits, typename Sink, int MODE, int FLAGS, typename OccT>
__device__ __forceinline__ void ct_pair_phase(const PairParams& p, const NlBox& B, const GridS& G,
                                              const PairCtx<Traits>& cx, Sink& sink, const OccT* occ,
                                              int n_occ, int* task_counter, unsigned long long& n_tests)
{
    const int lane = threadIdx.x & 31;
    int qn = 0;      // hits waiting in this warp's queue
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(task_counter, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_occ) break;
        ct_cell_task<Traits, Sink, MODE, FLAGS>(p, B, G, cx, sink, (int)occ[k], n_tests, qn);
    }
    ct_drain_hits<Traits, Sink>(cx, sink, qn);
}

template <typename Traits, typename Sink, int FLAGS, typename OccT>
__device__ __forceinline__ void ct_pair_phase_dyn(const PairParams& p, const NlBox& B, const GridS& G,
                                                  const PairCtx<Traits>& cx, Sink& sink, const OccT* occ,
                                                  int n_occ, int* task_counter, unsigned long long& n_tests)
{
    if (B.mode == BOX_ORTHO)
        ct_pair_phase<Traits, Sink, BOX_ORTHO, FLAGS, OccT>(p, B, G, cx, sink, occ, n_occ, task_counter, n_tests);
    else if (B.mode == BOX_TRICLINIC)
        ct_pair_phase<Traits, Sink, BOX_TRICLINIC, FLAGS, OccT>(p, B, G, cx, sink, occ, n_occ, task_counter, n_tests);
    else
        ct_pair_phase<Traits, Sink, BOX_NONE, FLAGS, OccT>(p, B, G, cx, sink, occ, n_occ, task_counter, n_tests);
}

}  // namespace cmb
