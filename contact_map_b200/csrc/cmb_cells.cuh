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


// ----- pair eligibility ----------------------------------------------------------
// meta = {ekey, res_c | role << 30}; role bit0 = query, bit1 = haystack
__device__ __forceinline__ bool pair_allowed(int2 mi, int2 mj, int n_ignored)
{
    const unsigned ri = (unsigned)mi.y >> 30, rj = (unsigned)mj.y >> 30;
    const unsigned role_ok = (ri & (rj >> 1)) | (rj & (ri >> 1));
    int de = mi.x - mj.x;
    de = de < 0 ? -de : de;
    return (role_ok & 1u) && (de > n_ignored);
}

// Everything a cell task needs besides the geometry.
struct TaskParams {
    const int2* meta;            // [n] {ekey, res_c | role << 30}
    int n;
    int n_ignored;
    float c2;                    // cutoff_f * cutoff_f
    float c2_hi;                 // c2 advanced by boundary_ulps float steps (== c2 when off)
    int boundary_ulps;
    unsigned int* atom_table;    // [n][n] uint32 (entry [lo][hi]) or nullptr
    unsigned long long* atom_keys;   // emitted (frame<<40 | lo<<20 | hi) or nullptr
    long long atom_cap;
    unsigned long long* counters;    // [0] atom keys, [1] residue keys, [2] boundary records, [3] pair tests
    float4* boundary;            // records {frame, lo, hi, d2} (ints bit-cast) or nullptr
    long long boundary_cap;
    long long frame_id;
    int count_tests;
};

constexpr int CT_KSLOT = 4;      // register-resident candidates per lane

// One warp adjudicates all candidate pairs of one occupied cell `c`:
//  * the <= 27 stencil cells are enumerated by lanes 0..26, de-duplicated (grids with
//    fewer than 3 cells along a dimension) and kept only if their id >= c, so every
//    unordered cell pair is visited once;
//  * candidates (own cell + kept neighbours) are loaded CT_KSLOT per lane into
//    registers, the cell's own atoms are broadcast one by one;
//  * sorted position order (jpos > ipos) removes self pairs and double visits.
// CellT: element type of the exclusive cell prefix (`cellptr[ncell+1]`).
// Sink: object with  __device__ void residue(int rlo, int rhi)  called for every hit.
template <int MODE, typename CellT, typename Sink>
__device__ __forceinline__ void cell_task(const TaskParams& p, const NlBox& B, const GridS& G,
                                          const float4* __restrict__ sorted,
                                          const CellT* __restrict__ cellptr, int c, int* wscr,
                                          Sink& sink, unsigned long long& n_tests)
{
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int cx = c % G.nx;
    const int cyz = c / G.nx;
    const int cy = cyz % G.ny;
    const int cz = cyz / G.ny;

    int myD = -1;
    if (lane < 27) {
        int x = cx + (lane % 3) - 1;
        int y = cy + ((lane / 3) % 3) - 1;
        int z = cz + (lane / 9) - 1;
        if (G.periodic) {
            x = x < 0 ? x + G.nx : (x >= G.nx ? x - G.nx : x);
            y = y < 0 ? y + G.ny : (y >= G.ny ? y - G.ny : y);
            z = z < 0 ? z + G.nz : (z >= G.nz ? z - G.nz : z);
        }
        const bool ok = (x >= 0 && x < G.nx && y >= 0 && y < G.ny && z >= 0 && z < G.nz);
        if (ok) myD = (z * G.ny + y) * G.nx + x;
    }
    const unsigned same = __match_any_sync(FULL, myD);
    const bool keep = (myD >= c) && ((__ffs(same) - 1) == lane);
    int mystart = 0, mycnt = 0;
    if (keep) {
        mystart = (int)cellptr[myD];
        mycnt = (int)cellptr[myD + 1] - mystart;
    }
    int incl = mycnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int M = __shfl_sync(FULL, incl, 31);
    const int off = incl - mycnt;
    __syncwarp();
    wscr[lane] = off;                  // candidate-slot offset of this neighbour cell
    wscr[32 + lane] = mystart - off;   // slot s -> sorted position s + wscr[32+t]
    __syncwarp();

    const int cstart = (int)cellptr[c], cend = (int)cellptr[c + 1];

    for (int base = 0; base < M; base += 32 * CT_KSLOT) {
        float xj[CT_KSLOT], yj[CT_KSLOT], zj[CT_KSLOT];
        int jpos[CT_KSLOT], jidx[CT_KSLOT];
        int2 mj[CT_KSLOT];
#pragma unroll
        for (int k = 0; k < CT_KSLOT; k++) {
            const int s = base + k * 32 + lane;
            jpos[k] = -1; jidx[k] = 0; xj[k] = yj[k] = zj[k] = 0.f; mj[k] = make_int2(0, 0);
            if (s < M) {
                int t = 0;
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1)
                    if (wscr[t + step] <= s) t += step;
                const int pos = s + wscr[32 + t];
                const float4 q = sorted[pos];
                xj[k] = q.x; yj[k] = q.y; zj[k] = q.z;
                jidx[k] = __float_as_int(q.w);
                jpos[k] = pos;
                mj[k] = __ldg(&p.meta[jidx[k]]);
            }
        }
        const int nk = min(CT_KSLOT, (M - base + 31) >> 5);
        if (p.count_tests && lane == 0)
            n_tests += (unsigned long long)(cend - cstart) * (unsigned long long)min(M - base, 32 * CT_KSLOT);

        for (int ipos = cstart; ipos < cend; ipos++) {
            const float4 pi = sorted[ipos];
            const int iidx = __float_as_int(pi.w);
            const int2 mi = __ldg(&p.meta[iidx]);
#pragma unroll
            for (int k = 0; k < CT_KSLOT; k++) {
                if (k < nk) {
                    // tested in the direction lower sorted position -> higher; the arithmetic is
                    // antisymmetric for every pair that can be a hit ("direction")
                    const float d2 = nl_d2<MODE>(pi.x, pi.y, pi.z, xj[k], yj[k], zj[k], B);
                    const bool near = (d2 <= p.c2_hi) && (jpos[k] > ipos) && pair_allowed(mi, mj[k], p.n_ignored);
                    const bool hit = near && (d2 <= p.c2);
                    const int a = min(iidx, jidx[k]), b = max(iidx, jidx[k]);
                    if (p.boundary_ulps > 0 && near) {
                        int diff = __float_as_int(d2) - __float_as_int(p.c2);
                        diff = diff < 0 ? -diff : diff;
                        if (diff <= p.boundary_ulps && p.boundary != nullptr) {
                            const unsigned long long w = atomicAdd(&p.counters[2], 1ull);
                            if ((long long)w < p.boundary_cap)
                                p.boundary[w] = make_float4(__int_as_float((int)p.frame_id), __int_as_float(a),
                                                            __int_as_float(b), d2);
                        }
                    }
                    if (p.atom_keys != nullptr) {
                        // warp-ballot compaction of this slot's hits into the key stream
                        const unsigned hm = __ballot_sync(FULL, hit);
                        if (hm) {
                            const int leader = __ffs(hm) - 1;
                            unsigned long long wbase = 0;
                            if (lane == leader) wbase = atomicAdd(&p.counters[0], (unsigned long long)__popc(hm));
                            wbase = __shfl_sync(FULL, wbase, leader);
                            if (hit) {
                                const unsigned long long w = wbase + __popc(hm & ((1u << lane) - 1u));
                                if ((long long)w < p.atom_cap)
                                    p.atom_keys[w] = ((unsigned long long)p.frame_id << 40) |
                                                     ((unsigned long long)a << 20) | (unsigned long long)b;
                            }
                        }
                    }
                    if (hit) {
                        if (p.atom_table != nullptr)
                            atomicAdd(&p.atom_table[(size_t)a * (size_t)p.n + (size_t)b], 1u);
                        const int ra = mi.y & 0x3fffffff, rb = mj[k].y & 0x3fffffff;
                        sink.residue(min(ra, rb), max(ra, rb));
                    }
                }
            }
        }
    }
}

}  // namespace cmb
