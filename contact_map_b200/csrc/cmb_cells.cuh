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
    double h[9];     // unit cell rows (periodic grids): wrapped position = p - floor(s) . h
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
    // Minimum cell width.  The slack over the cutoff (0.2 %) is what lets the stencil find every
    // pair the reference arithmetic can call a contact: its rounding error is <= ~16 ulp of the
    // largest coordinate, i.e. below the slack for |coordinates| < ~900 nm at a 0.45 nm cutoff
    // ("domain")
    const double cw = cutoff * 1.002 + 1e-7;
    if (box9 != nullptr) {
        double h[9];
        for (int k = 0; k < 9; k++) h[k] = (double)box9[k];
        const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) +
                           h[2] * (h[3] * h[7] - h[4] * h[6]);
        const double id = 1.0 / det;
        for (int k = 0; k < 9; k++) G.h[k] = h[k];
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
        for (int k = 0; k < 9; k++) { G.m[k] = 0.0; G.h[k] = 0.0; }
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

// Cell of an atom; on periodic grids (x, y, z) is replaced by the position wrapped into the
// primary cell image (fp64, rounded once to fp32).  The cell only selects candidates and the
// wrapped position only feeds the screening distance: neither ever decides an ambiguous pair.
__device__ __forceinline__ int cell_wrap(const GridS& G, float& x, float& y, float& z)
{
    const double px = (double)x - G.o[0], py = (double)y - G.o[1], pz = (double)z - G.o[2];
    double sx = px * G.m[0] + py * G.m[3] + pz * G.m[6];
    double sy = px * G.m[1] + py * G.m[4] + pz * G.m[7];
    double sz = px * G.m[2] + py * G.m[5] + pz * G.m[8];
    if (G.periodic) {
        const double fx = floor(sx), fy = floor(sy), fz = floor(sz);
        sx -= fx; sy -= fy; sz -= fz;
        x = (float)(px - (fx * G.h[0] + fy * G.h[3] + fz * G.h[6]));
        y = (float)(py - (fx * G.h[1] + fy * G.h[4] + fz * G.h[7]));
        z = (float)(pz - (fx * G.h[2] + fy * G.h[5] + fz * G.h[8]));
    }
    int cx = (int)(sx * (double)G.nx);
    int cy = (int)(sy * (double)G.ny);
    int cz = (int)(sz * (double)G.nz);
    cx = cx < 0 ? 0 : (cx >= G.nx ? G.nx - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= G.ny ? G.ny - 1 : cy);
    cz = cz < 0 ? 0 : (cz >= G.nz ? G.nz - 1 : cz);
    return (cz * G.ny + cy) * G.nx + cx;
}


// interior cell: no periodic wrap and no grid edge anywhere in its 27-cell stencil
__device__ __forceinline__ bool cell_interior(const GridS& G, int cx, int cy, int cz)
{
    return (cx >= 1) && (cx <= G.nx - 2) && (cy >= 1) && (cy <= G.ny - 2) && (cz >= 1) && (cz <= G.nz - 2);
}

// two x-adjacent occupied interior cells are merged into one task when they are sparsely
// populated (their candidate set, ~19 cells, then still fits the register-resident slots)
#ifndef CT_PAIR_MAX_OWN_CFG
#define CT_PAIR_MAX_OWN_CFG 18
#endif
constexpr int CT_PAIR_MAX_OWN = CT_PAIR_MAX_OWN_CFG;

// =================================================================================
// Pair phase shared by both paths.
//
// Screening and adjudication are split:
//   * the frame's atoms are cell-sorted into `sorted[pos] = {wx, wy, wz, bits}`: the position
//     WRAPPED into the primary cell image (fp64 in the binning pass, rounded to fp32) and
//     `bits` = atom index, role and (fused kernel) exclusion key, layout known to Traits;
//   * one warp per task (an occupied cell, or two sparse x-adjacent cells); every candidate pair
//     of the half-shell stencil costs 10 issue slots: t = dx*dx + dy*dy + dz*dz - c2_hi on the
//     wrapped coordinates (+ the lattice shift of the candidate's stencil cell) as 3 FADD +
//     3 FFMA, the residue exclusion q = n_ignored + 0.5 - |ekey_i - ekey_j| as 2 FADD, and the
//     sign of max(t, q) funnel-shifted into a per-lane bit mask (one bit per own atom of the
//     task, one mask per register-resident candidate);
//   * after the own-atom loop the non-zero masks are pushed to a per-warp shared-memory ring
//     and accounted 32 pairs at a time with all lanes busy: d2a <= c2_lo is a certain contact,
//     anything else that passed the screen is AMBIGUOUS and is decided by MDTraj's exact
//     non-fused fp32 arithmetic on the RAW coordinates (cmb_math.cuh::nl_d2).
//     [c2_lo, c2_hi] = c2 -/+ a rigorous per-frame bound on |d2_exact - d2_screen|
//     (frame_margins); when the bound cannot be established (fewer than 3 cells along a
//     periodic dimension, non-triangular box, huge coordinates) the band is everything and
//     every candidate is adjudicated exactly.
// A `Traits` type fixes the integer widths of the two paths (cell pointers, candidate list
// entries, ring entries); a `Sink` object owns the per-atom meta words and accounts one hit
// (atom table, per-frame residue de-duplication).
// =================================================================================
constexpr int CT_KSLOT = 4;                 // register-resident candidates per lane
constexpr int CT_CAND = 32 * CT_KSLOT;      // candidates per chunk
constexpr int CT_QCAP = 64;                 // per-warp ring of screened (candidate, own-atom mask) entries
constexpr int CT_NOSHIFT = 13;              // shift id of "no lattice shift"
constexpr float CT_FAR = 1.0e18f;           // coordinate of an empty candidate slot

// kernel variants (template FLAGS)
constexpr int CT_EMIT = 1;        // emit (frame, a, b) keys
constexpr int CT_BOUNDARY = 2;    // report pairs within boundary_ulps of the cutoff
constexpr int CT_COUNT = 4;       // count executed pair tests
constexpr int CT_ROLES = 8;       // query / haystack are not "all atoms": test the roles

// hot-loop parameters (everything a cell task needs besides the geometry)
struct PairParams {
    int n;                           // used atoms
    int n_ignored;
    float c2;                        // cutoff_f * cutoff_f
    float c2_lo, c2_hi;              // certain-contact / screening thresholds of this frame
    int boundary_ulps;               // CT_BOUNDARY: > 0 report pairs within that many float steps of the cutoff;
                                     // < 0 AUDIT: every candidate that is not residue-excluded is decided by the
                                     // reference arithmetic and compared with what the screen would have decided
    float audit_lo, audit_hi;        // audit: the thresholds the production kernels use for this frame
    // the frame: addressed through its index so that only one register stays live in the hot loop
    const float* xyz;                // RAW coordinates [frames][n][3] (global memory)
    const float* box;                // unit cell rows [frames][9] (global memory) or nullptr
    int f;                           // index of the frame in xyz / box
    long long frame0;                // id of frame 0 (emitted keys, boundary records)
    unsigned long long* atom_keys;   // emitted (frame<<40 | lo<<20 | hi) or nullptr
    long long atom_cap;
    unsigned long long* counters;    // [0] atom keys, [1] residue keys, [2] boundary records, [3] pair tests
    float4* boundary;                // records {frame, lo, hi, d2} (ints bit-cast) or nullptr
    long long boundary_cap;
};

template <typename Traits>
struct PairCtx {
    const float4* sorted;                       // [n] {wx, wy, wz, bits}: atom index, role (+ exclusion
                                                // key in the fused kernel), layout known to Traits
    const float2* kr;                           // [n] {exclusion key, compact residue (int bits)} per sorted
                                                // position (global-memory path; nullptr in the fused kernel)
    const typename Traits::pos_t* cellptr;      // [ncell + 1]
    typename Traits::cand_t* cand;              // this warp's candidate list [CT_CAND]
    typename Traits::hit_t* hitq;               // this warp's ring [CT_QCAP]
    const float4* shift;                        // [27] lattice shift of each wrap combination
    const unsigned* cellrole;                   // 2 role bits per cell (OR over its atoms), 16 cells per word,
                                                // or nullptr (Traits::CELL_ROLES)
};

// screening distance, as recomputed when a screened pair is accounted
__device__ __forceinline__ float d2_screen(float xi, float yi, float zi, float xj, float yj, float zj)
{
    const float dx = __fsub_rn(xj, xi), dy = __fsub_rn(yj, yi), dz = __fsub_rn(zj, zi);
    return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// lane id / lanes-below mask straight from the special registers: one S2R when the compiler
// re-materialises them, instead of S2R tid + mask + shift + subtract
__device__ __forceinline__ int ct_lane()
{
    unsigned v;
    asm("mov.u32 %0, %%laneid;" : "=r"(v));
    return (int)v;
}
__device__ __forceinline__ unsigned ct_ltmask()
{
    unsigned v;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(v));
    return v;
}

// Per-warp ring of screened entries: head / tail live in (warp-uniform) registers, pushes are
// warp-aggregated with one ballot (no atomics, no shared counter).
struct Ring {
    unsigned head, tail;
};

template <typename Traits>
__device__ __forceinline__ void ring_push(const PairCtx<Traits>& cx, Ring& ring, unsigned ltmask, bool want,
                                          const typename Traits::hit_t& e)
{
    const unsigned nz = __ballot_sync(0xffffffffu, want);
    if (want) cx.hitq[(ring.tail + __popc(nz & ltmask)) & (CT_QCAP - 1)] = e;
    ring.tail += __popc(nz);
}

// Per-frame thresholds.  A = max |raw coordinate| of the frame, box9 = unit cell rows or nullptr.
// Bound on |d2_exact - d2_screen| for pairs whose distance is near the cutoff:
// per-component error of the exact chain <= 8u * (largest intermediate magnitude), of the
// screening chain <= 10u * (sum of |box entries|) (wrapped position, shift, subtraction);
// both chains resolve the same periodic image when the stencil cells are distinct, the box is
// lower triangular and the image counts stay far below 2^23.  The rounding of either
// accumulation order of the squares is covered by the 24u c^2 term.  Evaluated in fp32 (one
// thread, on the critical path of the frame): every quantity is non-negative and only an upper
// bound is needed, so the fp32 rounding of this evaluation itself is absorbed by the 1.01 factor
// on the final margin, which is then applied with directed rounding.
__device__ inline void frame_margins(const float* box9, int distinct, float A, float c2, float cutoff,
                                     int extra_ulps, float& c2_lo, float& c2_hi)
{
    const float u = 5.9604644775390625e-8f;   // 2^-24
    bool ok = true;
    float e;
    if (box9 == nullptr) {
        e = 4.0f * u * A;
    } else {
        const float ax = box9[0], ay = box9[1], az = box9[2], bx = fabsf(box9[3]), by = box9[4],
                    bz = box9[5], cx = fabsf(box9[6]), cy = fabsf(box9[7]), cz = box9[8];
        ok = distinct && ay == 0.f && az == 0.f && bz == 0.f && ax > 0.f && by > 0.f && cz > 0.f;
        const float sc = 2.0f * A / cz + 2.0f;
        const float sb = (2.0f * A + sc * cy) / by + 2.0f;
        const float sa = (2.0f * A + sc * cx + sb * bx) / ax + 2.0f;
        ok = ok && sa < 1.0e5f && sb < 1.0e5f && sc < 1.0e5f;
        const float mx = 2.0f * A + sc * cx + sb * bx + sa * ax;
        const float my = 2.0f * A + sc * cy + sb * by;
        const float mz = 2.0f * A + sc * cz;
        const float mmax = fmaxf(mx, fmaxf(my, mz));
        const float S = ax + bx + by + cx + cy + cz;
        e = 8.0f * u * mmax + 10.0f * u * S + 1.0e-15f * mmax;
    }
    const float c = cutoff * 1.01f;
    const float bound = 1.7320508f * e * (2.0f * c + 2.0f * e) + 24.0f * u * c * c;
    const float M = 2.0f * 1.01f * bound;
    ok = ok && (M < 0.25f * c2);        // also false for NaN
    if (!ok) {
        // adjudicate everything exactly; still below CT_FAR^2, so that empty slots never pass
        c2_lo = -1.0e30f;
        c2_hi = 1.0e30f;
        return;
    }
    c2_lo = __fsub_rd(c2, M);
    c2_hi = __fadd_ru(c2, M);
    if (extra_ulps > 0) {
        c2_lo = fminf(c2_lo, __int_as_float(__float_as_int(c2) - extra_ulps - 1));
        c2_hi = fmaxf(c2_hi, __int_as_float(__float_as_int(c2) + extra_ulps + 1));
    }
}

// lattice shift of wrap combination id = (wx+1) + 3(wy+1) + 9(wz+1), w in {-1,0,1}: an atom of a
// stencil cell reached through a periodic wrap appears at (wrapped position) + shift
__device__ inline float4 lattice_shift(const float* box9, int id)
{
    if (box9 == nullptr) return make_float4(0.f, 0.f, 0.f, 0.f);
    const double wx = (double)(id % 3 - 1), wy = (double)((id / 3) % 3 - 1), wz = (double)(id / 9 - 1);
    return make_float4((float)(wx * box9[0] + wy * box9[3] + wz * box9[6]),
                       (float)(wx * box9[1] + wy * box9[4] + wz * box9[7]),
                       (float)(wx * box9[2] + wy * box9[5] + wz * box9[8]), 0.f);
}

__device__ __forceinline__ int box_mode(const float* __restrict__ box9)
{
    if (box9 == nullptr) return BOX_NONE;
    const bool tri = (box9[1] != 0.f) || (box9[2] != 0.f) || (box9[3] != 0.f) || (box9[5] != 0.f) ||
                     (box9[6] != 0.f) || (box9[7] != 0.f);
    return tri ? BOX_TRICLINIC : BOX_ORTHO;
}

// Account one screened pair (sorted positions ipos < jpos, shift id of jpos' stencil cell);
// `active` is false for the idle lanes of a partial drain.  Called by all 32 lanes.
template <typename Traits, typename Sink, int MODE, int FLAGS>
__device__ __forceinline__ void ct_account(const PairParams& p, const PairCtx<Traits>& cx, Sink& sink, int ipos,
                                           int jpos, int id, bool active)
{
    const int lane = ct_lane();
    const float4 wi = cx.sorted[ipos], wj = cx.sorted[jpos], sh = cx.shift[id];
    const float d2a = d2_screen(wi.x, wi.y, wi.z, __fadd_rn(wj.x, sh.x), __fadd_rn(wj.y, sh.y),
                                __fadd_rn(wj.z, sh.z));
    const unsigned bi = __float_as_uint(wi.w), bj = __float_as_uint(wj.w);
    const int ia = Traits::atom(bi), ib = Traits::atom(bj);
    sink.prepare(cx, ipos, jpos, ia, ib);      // residues of the pair (the exclusion was part of the screen)
    bool elig = active;
    if (FLAGS & CT_ROLES) {
        const unsigned ri = Traits::role(bi), rj = Traits::role(bj);
        elig = elig && (((ri & (rj >> 1)) | (rj & (ri >> 1))) & 1u);
    }
    bool hit = elig && (d2a <= p.c2_lo);
    if (elig && !hit) {
        // ambiguous: MDTraj's arithmetic on the raw coordinates, direction lower -> higher sorted
        // position (antisymmetric for every pair that can be a hit, DESIGN.md "direction")
        const NlBox B = make_nlbox(p.box ? p.box + (size_t)p.f * 9 : nullptr);
        const float* frame = p.xyz + (size_t)p.f * (size_t)p.n * 3;
        const float* pa = frame + 3 * (size_t)ia;
        const float* pb = frame + 3 * (size_t)ib;
        const float d2 = nl_d2<MODE>(__ldg(pa), __ldg(pa + 1), __ldg(pa + 2), __ldg(pb), __ldg(pb + 1),
                                     __ldg(pb + 2), B);
        hit = d2 <= p.c2;
        if ((FLAGS & CT_BOUNDARY) && p.boundary_ulps > 0) {
            int diff = __float_as_int(d2) - __float_as_int(p.c2);
            diff = diff < 0 ? -diff : diff;
            if (diff <= p.boundary_ulps) {
                const unsigned long long w = atomicAdd(&p.counters[2], 1ull);
                if ((long long)w < p.boundary_cap)
                    p.boundary[w] = make_float4(__int_as_float((int)(p.frame0 + p.f)), __int_as_float(min(ia, ib)),
                                                __int_as_float(max(ia, ib)), d2);
            }
        }
    }
    if ((FLAGS & CT_BOUNDARY) && p.boundary_ulps < 0) {
        // audit (counters[2] = pairs audited, counters[3] = disagreements): the production screen
        // calls d2a <= lo a contact without looking and drops d2a >= hi without looking
        const bool bad = elig && ((d2a <= p.audit_lo && !hit) || (d2a >= p.audit_hi && hit));
        const unsigned am = __ballot_sync(0xffffffffu, elig), bm = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (am) atomicAdd(&p.counters[2], (unsigned long long)__popc(am));
            if (bm) atomicAdd(&p.counters[3], (unsigned long long)__popc(bm));
        }
    }
    if (FLAGS & CT_EMIT) {
        if (p.atom_keys != nullptr) {
            // warp-ballot compaction of this batch's hits into the key stream
            const unsigned hm = __ballot_sync(0xffffffffu, hit);
            if (hm) {
                const int leader = __ffs(hm) - 1;
                unsigned long long wbase = 0;
                if (lane == leader) wbase = atomicAdd(&p.counters[0], (unsigned long long)__popc(hm));
                wbase = __shfl_sync(0xffffffffu, wbase, leader);
                if (hit) {
                    const unsigned long long w = wbase + __popc(hm & ((1u << lane) - 1u));
                    if ((long long)w < p.atom_cap)
                        p.atom_keys[w] = ((unsigned long long)(p.frame0 + p.f) << 40) |
                                         ((unsigned long long)min(ia, ib) << 20) | (unsigned long long)max(ia, ib);
                }
            }
        }
    }
    if (hit) sink.account(ia, ib);
}

// Pop up to 32 ring entries (one per lane); an entry is (own-atom base, candidate, mask of own
// atoms): the lane accounts the lowest own atom of its mask and pushes the rest back.
template <typename Traits, typename Sink, int MODE, int FLAGS>
__device__ __forceinline__ void ct_drain(const PairParams& p, const PairCtx<Traits>& cx, Sink& sink, Ring& ring,
                                         unsigned ltmask)
{
    const int lane = ct_lane();
    const unsigned left = ring.tail - ring.head;
    const bool active = (unsigned)lane < left;
    __syncwarp();
    const typename Traits::hit_t e = cx.hitq[(ring.head + (active ? lane : 0)) & (CT_QCAP - 1)];
    ring.head += left < 32u ? left : 32u;
    const unsigned m = Traits::e_mask(e);
    const int bit = __ffs(m) - 1;
    const unsigned rest = m & (m - 1u);
    __syncwarp();
    ring_push<Traits>(cx, ring, ltmask, active && rest != 0u, Traits::with_mask(e, rest));
    ct_account<Traits, Sink, MODE, FLAGS>(p, cx, sink, Traits::e_ibase(e) + bit, Traits::e_jpos(e),
                                          Traits::e_id(e), active);
}

// Screening loop: own atoms [g0, g1) (at most 32, highest first so that bit b of a mask is own
// atom g0 + b) against NK register-resident candidates per lane.  Per pair: 3 FADD + 3 FFMA for
// t = d2 - c2_hi, 2 FADD + 1 FMNMX for the residue exclusion (q = n_ignored + 0.5 - |ekey_i -
// ekey_j| is negative iff the pair is NOT excluded; keys are small integers, exact in fp32), and
// one SHF that shifts the sign of max(t, q) into the mask.
template <typename Traits, int NK, int FLAGS>
__device__ __forceinline__ void ct_screen(const PairCtx<Traits>& cx, int g0, int g1, float neg_c2_hi, float nq,
                                          const float (&xj)[CT_KSLOT], const float (&yj)[CT_KSLOT],
                                          const float (&zj)[CT_KSLOT], const float (&ej)[CT_KSLOT], bool any_q,
                                          bool any_h, unsigned (&m)[CT_KSLOT])
{
#pragma unroll 2
    for (int ipos = g1 - 1; ipos >= g0; ipos--) {
        const float4 pi = cx.sorted[ipos];
        const float ei = Traits::ekey_f(cx, ipos, __float_as_uint(pi.w));
        if (FLAGS & CT_ROLES) {
            const unsigned ri = Traits::role(__float_as_uint(pi.w));
            if (!(((ri & 1u) && any_h) || ((ri & 2u) && any_q))) {
#pragma unroll
                for (int k = 0; k < NK; k++) m[k] <<= 1;
                continue;
            }
        }
#pragma unroll
        for (int k = 0; k < NK; k++) {
            const float dx = __fsub_rn(xj[k], pi.x), dy = __fsub_rn(yj[k], pi.y), dz = __fsub_rn(zj[k], pi.z);
            const float t = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmaf_rn(dx, dx, neg_c2_hi)));
            const float q = __fsub_rn(nq, fabsf(__fsub_rn(ej[k], ei)));
            m[k] = __funnelshift_l(__float_as_uint(fmaxf(t, q)), m[k], 1);   // sign: d2 < c2_hi, not excluded
        }
    }
}

// One warp screens all candidate pairs of one task: an occupied cell `c`, or (TASK_PAIR) the two
// x-adjacent interior cells c, c + 1 of a sparsely populated region (halves the per-task set-up
// where a cell holds only a handful of atoms).
//
// Candidates are the atoms of the stencil cells with id >= c (every unordered cell pair is
// visited once), own cell(s) first.  Cells are numbered x-fastest, so for an INTERIOR task (no
// periodic wrap, no grid edge) they form five contiguous runs of the cell-sorted array
//     (cx..cx+L, cy, cz), (cx-1..cx+L, cy+1, cz), (cx-1..cx+L, cy-1..cy+1, cz+1)       L = 1 or 2
// and a candidate slot maps to its sorted position with four compares (for L = 2 the runs also
// cover cells two columns away from one of the own cells: wasted but harmless tests).  Edge
// cells take the generic route: 27 stencil lanes (lane 0 = own cell), de-duplication (grids with
// < 3 cells along a dimension), warp scan, per-warp candidate list whose entries carry the wrap
// flags of their stencil cell (-> lattice shift added when the candidate is loaded).
// Own-cell candidates occupy the first slots in sorted order, so "each unordered pair once, no
// self pair" is `slot > own index`, applied to the masks before they are pushed to the ring.
template <typename Traits, typename Sink, int MODE, int FLAGS>
__device__ __forceinline__ void ct_cell_task(const PairParams& p, const GridS& G, const PairCtx<Traits>& cx,
                                             Sink& sink, unsigned task, unsigned long long& n_tests, Ring& ring)
{
    const int lane = ct_lane();
    const unsigned FULL = 0xffffffffu;
    const unsigned ltmask = ct_ltmask();
    const int c = (int)(task & Traits::TASK_CMASK);
    const int L = (task & Traits::TASK_PAIR) ? 2 : 1;
    if ((FLAGS & CT_ROLES) && Traits::CELL_ROLES && (task & Traits::TASK_INTERIOR)) {
        // Query / haystack subsets (a ligand against a receptor): most tasks cannot produce a
        // query-haystack pair at all.  One lane per cell of the task's five runs reads the cell's
        // role bits; the task is dropped before any candidate is loaded unless the own cells hold
        // a query atom and the runs a haystack atom, or the other way round.
        const int nx = G.nx, nxy = G.nx * G.ny, w = 2 + L;
        const int t = lane - (1 + L);
        const int r = (t >= 0) + (t >= w) + (t >= 2 * w) + (t >= 3 * w);
        int cell = c + lane;                                    // run 0: own cells and the next one
        cell = r == 1 ? c + nx - 1 + t : cell;
        cell = r == 2 ? c + nxy - nx - 1 + (t - w) : cell;
        cell = r == 3 ? c + nxy - 1 + (t - 2 * w) : cell;
        cell = r == 4 ? c + nxy + nx - 1 + (t - 3 * w) : cell;
        unsigned bits = 0u;
        if (t < 4 * w) bits = (cx.cellrole[cell >> 4] >> ((cell & 15) * 2)) & 3u;
        const unsigned all = __reduce_or_sync(FULL, bits);
        const unsigned own = __reduce_or_sync(FULL, lane < L ? bits : 0u);
        if (!(((own & 1u) && (all & 2u)) || ((own & 2u) && (all & 1u)))) return;
    }
    const int cstart = (int)cx.cellptr[c], cend = (int)cx.cellptr[c + L];

    for (int base = 0;; base += CT_CAND) {
        // ---- candidates of this chunk (everything is re-derived per chunk so that nothing of it
        //      stays live across the screening loop; a second chunk is rare) ------------------
        float xj[CT_KSLOT], yj[CT_KSLOT], zj[CT_KSLOT], ej[CT_KSLOT];
        unsigned jp[CT_KSLOT];
        unsigned roles_or = 0u;
        int M;
        if (task & Traits::TASK_INTERIOR) {
            // run r covers slots [o_r, o_{r+1}) and sorted positions slot + a_r
            const int nx = G.nx, nxy = G.nx * G.ny;
            const int b1 = c + nx - 1, b2 = c + nxy - nx - 1, b3 = c + nxy - 1, b4 = c + nxy + nx - 1;
            const int s0 = cstart, e0 = (int)cx.cellptr[c + 1 + L];
            const int s1 = (int)cx.cellptr[b1], e1 = (int)cx.cellptr[b1 + 2 + L];
            const int s2 = (int)cx.cellptr[b2], e2 = (int)cx.cellptr[b2 + 2 + L];
            const int s3 = (int)cx.cellptr[b3], e3 = (int)cx.cellptr[b3 + 2 + L];
            const int s4 = (int)cx.cellptr[b4], e4 = (int)cx.cellptr[b4 + 2 + L];
            const int o1 = e0 - s0, o2 = o1 + (e1 - s1), o3 = o2 + (e2 - s2), o4 = o3 + (e3 - s3);
            M = o4 + (e4 - s4);
            const int a0 = s0, a1 = s1 - o1, a2 = s2 - o2, a3 = s3 - o3, a4 = s4 - o4;
#pragma unroll
            for (int k = 0; k < CT_KSLOT; k++) {
                xj[k] = CT_FAR; yj[k] = zj[k] = ej[k] = 0.f; jp[k] = 0u;
                if (k >= 2 && base + k * 32 >= M) continue;      // slot beyond the candidates: never screened
                const int s = base + k * 32 + lane;
                int adj = a0;
                adj = s >= o1 ? a1 : adj;
                adj = s >= o2 ? a2 : adj;
                adj = s >= o3 ? a3 : adj;
                adj = s >= o4 ? a4 : adj;
                const int pos = s < M ? s + adj : cstart;       // empty slots: any valid position
                const float4 q = cx.sorted[pos];
                const unsigned bits = __float_as_uint(q.w);
                xj[k] = s < M ? q.x : CT_FAR; yj[k] = q.y; zj[k] = q.z;
                ej[k] = Traits::ekey_f(cx, pos, bits);
                jp[k] = Traits::jp(pos, CT_NOSHIFT);
                if (FLAGS & CT_ROLES) roles_or |= s < M ? Traits::role(bits) : 0u;
            }
        } else {
            int cxx, cy, cz;
            Traits::decode(c, G, cxx, cy, cz);
            // a wrap along x happens at one side only (grids with >= 3 cells; coarser periodic
            // grids are adjudicated exactly and never use the shift)
            const int idx_x = (cxx == 0) ? -1 : 1, idx_y = (cy == 0) ? -3 : 3, idx_z = (cz == 0) ? -9 : 9;
            int myD = -1, mywrap = 0;
            if (lane < 27) {
                const int si = lane == 0 ? 13 : (lane == 13 ? 0 : lane);     // own cell first
                const int l9 = si / 9, l3 = (si - l9 * 9) / 3;
                int x = cxx + (si - l9 * 9 - l3 * 3) - 1;
                int y = cy + l3 - 1;
                int z = cz + l9 - 1;
                if (G.periodic) {
                    if (x < 0) { x += G.nx; mywrap |= 1; } else if (x >= G.nx) { x -= G.nx; mywrap |= 1; }
                    if (y < 0) { y += G.ny; mywrap |= 2; } else if (y >= G.ny) { y -= G.ny; mywrap |= 2; }
                    if (z < 0) { z += G.nz; mywrap |= 4; } else if (z >= G.nz) { z -= G.nz; mywrap |= 4; }
                }
                const bool ok = (x >= 0 && x < G.nx && y >= 0 && y < G.ny && z >= 0 && z < G.nz);
                if (ok) myD = (z * G.ny + y) * G.nx + x;
            }
            bool keep = (myD >= c);
            if (!G.distinct) {
                const unsigned same = __match_any_sync(FULL, myD);
                keep = keep && ((__ffs(same) - 1) == lane);
            }
            int mystart = 0, mycnt = 0;
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
            const int off = incl - mycnt;
            __syncwarp();
            // expand this lane's stencil cell into the per-warp candidate list
            const int r0 = max(0, base - off), r1 = min(mycnt, base + CT_CAND - off);
            for (int r = r0; r < r1; r++) cx.cand[off + r - base] = Traits::cand(mystart + r, mywrap);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < CT_KSLOT; k++) {
                const int s = base + k * 32 + lane;
                const typename Traits::cand_t ce = s < M ? cx.cand[s - base] : Traits::cand(cstart, 0);
                const int pos = Traits::cand_pos(ce);
                const int wb = Traits::cand_wrap(ce);
                const int id = CT_NOSHIFT + ((wb & 1) ? idx_x : 0) + ((wb & 2) ? idx_y : 0) +
                               ((wb & 4) ? idx_z : 0);
                const float4 q = cx.sorted[pos];
                const float4 sh = cx.shift[id];
                const unsigned bits = __float_as_uint(q.w);
                xj[k] = s < M ? __fadd_rn(q.x, sh.x) : CT_FAR;
                yj[k] = __fadd_rn(q.y, sh.y); zj[k] = __fadd_rn(q.z, sh.z);
                ej[k] = Traits::ekey_f(cx, pos, bits);
                jp[k] = Traits::jp(pos, id);
                if (FLAGS & CT_ROLES) roles_or |= s < M ? Traits::role(bits) : 0u;
            }
        }
        bool any_q = true, any_h = true;
        if (FLAGS & CT_ROLES) {
            any_q = __any_sync(FULL, roles_or & 1u);
            any_h = __any_sync(FULL, roles_or & 2u);
        }
        const int left = M - base;
        if (FLAGS & CT_COUNT) {
            if (lane == 0) n_tests += (unsigned long long)(cend - cstart) * (unsigned long long)min(left, CT_CAND);
        }
        const float neg_c2_hi = -p.c2_hi, nq = (float)p.n_ignored + 0.5f;
        for (int g0 = cstart; g0 < cend; g0 += 32) {
            const int g1 = min(g0 + 32, cend);
            unsigned m[CT_KSLOT];
#pragma unroll
            for (int k = 0; k < CT_KSLOT; k++) m[k] = 0u;
            if (left > 96) ct_screen<Traits, 4, FLAGS>(cx, g0, g1, neg_c2_hi, nq, xj, yj, zj, ej, any_q, any_h, m);
            else if (left > 64) ct_screen<Traits, 3, FLAGS>(cx, g0, g1, neg_c2_hi, nq, xj, yj, zj, ej, any_q, any_h, m);
            else if (left > 32) ct_screen<Traits, 2, FLAGS>(cx, g0, g1, neg_c2_hi, nq, xj, yj, zj, ej, any_q, any_h, m);
            else ct_screen<Traits, 1, FLAGS>(cx, g0, g1, neg_c2_hi, nq, xj, yj, zj, ej, any_q, any_h, m);
            // own-cell slots: own atoms g0 + b with b < v pair with slot s (each unordered pair once,
            // no self pair).  Empty slots never pass the screen (CT_FAR).
            if (base == 0 && g0 == cstart) {
                m[0] &= ltmask;                       // v = lane; slots >= 32 pair with every own atom
            } else {
#pragma unroll
                for (int k = 0; k < CT_KSLOT; k++) {
                    const int v = base + 32 * k + lane - (g0 - cstart);
                    if (v <= 0) m[k] = 0u;
                    else if (v < 32) m[k] &= (1u << v) - 1u;
                }
            }
            // push the non-empty masks to the ring: all at once when they fit (the usual case),
            // else slot by slot with full batches accounted in between.  ONE accounting site:
            // the code behind ct_drain is large.
            unsigned nz[CT_KSLOT], total = 0u;
#pragma unroll
            for (int k = 0; k < CT_KSLOT; k++) {
                nz[k] = 0u;
                if (k == 0 || left > 32 * k) {                   // slots beyond `left` are empty
                    nz[k] = __ballot_sync(FULL, m[k] != 0u);
                    total += (unsigned)__popc(nz[k]);
                }
            }
            int done = 0;
            if (ring.tail - ring.head + total <= (unsigned)CT_QCAP) {
                unsigned at = ring.tail;
#pragma unroll
                for (int k = 0; k < CT_KSLOT; k++) {
                    if (nz[k] != 0u) {
                        if (m[k] != 0u)
                            cx.hitq[(at + __popc(nz[k] & ltmask)) & (CT_QCAP - 1)] = Traits::entry(g0, jp[k], m[k]);
                        at += (unsigned)__popc(nz[k]);
                    }
                }
                ring.tail = at;
                done = CT_KSLOT;
            }
            for (;;) {
                while (ring.tail - ring.head > 31u) ct_drain<Traits, Sink, MODE, FLAGS>(p, cx, sink, ring, ltmask);
                if (done >= CT_KSLOT) break;
                // (rare) next slot on its own: at most 32 entries, the ring holds at most 31
                unsigned mk = m[0], jpk = jp[0];
#pragma unroll
                for (int k = 1; k < CT_KSLOT; k++) {
                    mk = done == k ? m[k] : mk;
                    jpk = done == k ? jp[k] : jpk;
                }
                ring_push<Traits>(cx, ring, ltmask, mk != 0u, Traits::entry(g0, jpk, mk));
                done++;
            }
        }
        if (left <= CT_CAND) break;
    }
}

// Warps pull tasks from a task counter until the frame is exhausted.
template <typename Traits, typename Sink, int MODE, int FLAGS, typename OccT>
__device__ __forceinline__ void ct_pair_phase(const PairParams& p, const GridS& G, const PairCtx<Traits>& cx,
                                              Sink& sink, const OccT* occ, int n_occ, int* task_counter,
                                              unsigned long long& n_tests)
{
    const int lane = ct_lane();
    Ring ring;
    ring.head = ring.tail = 0u;
    for (;;) {
        // TASK_BATCH consecutive list entries per fetch: the fetch (atomic + SHFL + bounds) and
        // the entries the pair merge left empty cost one trip per batch instead of one each
        int k = 0;
        if (lane == 0) k = Traits::next_tasks(task_counter, Traits::TASK_BATCH);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_occ) break;
        unsigned mine = Traits::TASK_EMPTY;
        if (lane < Traits::TASK_BATCH && k + lane < n_occ) mine = (unsigned)occ[k + lane];
#pragma unroll 1
        for (int t = 0; t < Traits::TASK_BATCH; t++) {
            const unsigned task = __shfl_sync(0xffffffffu, mine, t);
            if (task == Traits::TASK_EMPTY) continue;
            ct_cell_task<Traits, Sink, MODE, FLAGS>(p, G, cx, sink, task, n_tests, ring);
        }
    }
    const unsigned ltmask = ct_ltmask();
    while (ring.tail != ring.head) ct_drain<Traits, Sink, MODE, FLAGS>(p, cx, sink, ring, ltmask);
}

template <typename Traits, typename Sink, int FLAGS, typename OccT>
__device__ __forceinline__ void ct_pair_phase_dyn(const PairParams& p, const GridS& G, const PairCtx<Traits>& cx,
                                                  Sink& sink, const OccT* occ, int n_occ, int* task_counter,
                                                  unsigned long long& n_tests)
{
    const int mode = box_mode(p.box ? p.box + (size_t)p.f * 9 : nullptr);
    if (mode == BOX_ORTHO)
        ct_pair_phase<Traits, Sink, BOX_ORTHO, FLAGS, OccT>(p, G, cx, sink, occ, n_occ, task_counter, n_tests);
    else if (mode == BOX_TRICLINIC)
        ct_pair_phase<Traits, Sink, BOX_TRICLINIC, FLAGS, OccT>(p, G, cx, sink, occ, n_occ, task_counter, n_tests);
    else
        ct_pair_phase<Traits, Sink, BOX_NONE, FLAGS, OccT>(p, G, cx, sink, occ, n_occ, task_counter, n_tests);
}

}  // namespace cmb
