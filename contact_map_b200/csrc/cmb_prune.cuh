// cmb_prune.cuh -- cell-pruned NearestAtoms and MinimumDistanceCounter.
//
// Replaces (reference file:line under /root/reference/contact_map/):
//   NearestAtoms._calculate_nearest      min_dist.py:54-69   (md.compute_neighborlist + compute_distances + argmin)
//   MinimumDistanceCounter (Q x H)       min_dist.py:134-168 (md.compute_distances over the product, min / argmin)
//
// Both answers are decided by pairs that are CLOSE: a neighbour within the cutoff, the closest
// query / haystack pair of a frame.  The brute-force kernels (cmb_aux_kernels.cuh: k_nearest,
// k_min_dist_partial) look at all N^2 / |Q| x |H| pairs; here the atoms of every frame of a batch
// are binned into cells of a guaranteed width w (perpendicular width for triclinic cells, fp64
// fractional coordinates: two atoms whose true minimum-image distance is below w sit in adjacent
// cells) and only the 27-cell stencil of every query atom is searched -- with the SAME arithmetic
// as the brute-force kernels (nl_d2 / dist2_pair + __fsqrt_rn, same tie rules), so wherever the
// pruned search is valid its result is bit-identical:
//   nearest:  w = 1.002 cutoff (the candidate coverage argument of the contact kernels, DESIGN.md
//             4.0 "domain"); needs >= 3 cells per periodic dimension, else brute force;
//   min-dist: w from the haystack density; the stencil minimum M is accepted for a frame iff
//             M <= 0.9 w: every pair outside the stencil is > w apart, so (10 % slack against
//             float32 rounding) it can neither beat M nor tie with it.  Frames that fail the test
//             (query and haystack far apart, degenerate cells) are flagged and re-done by the
//             brute-force kernel.
#pragma once
#include "cmb_cells.cuh"

namespace cmb {

constexpr int PG_MAXC = 1 << 16;            // cells of one frame's grid (coarser cells stay valid)
constexpr int PG_THREADS = 128;

struct PgFrame {
    GridS grid;
    float width;                            // guaranteed search radius of the 27-cell stencil
    int usable;                             // 0: this frame cannot be pruned
    unsigned bb_lo[3], bb_hi[3];            // bounding box (ordered uints), frames without a unit cell
};

// per-batch device buffers (slices per frame)
struct PgScratch {
    long long cap_frames = 0;
    int cap_sel = 0;
    PgFrame* frames = nullptr;              // [cap_frames]
    int* cellcnt = nullptr;                 // [cap_frames][PG_MAXC + 1]
    int* cellptr = nullptr;                 // [cap_frames][PG_MAXC + 1]
    int* rank = nullptr;                    // [cap_frames][cap_sel]
    int* cell = nullptr;                    // [cap_frames][cap_sel]
    float4* sorted = nullptr;               // [cap_frames][cap_sel] {raw x, y, z, position in the selection}
    int* flags = nullptr;                   // [cap_frames + 1] min-dist: frame needs the brute-force kernel; [cap] = count
    int* h_flags = nullptr;                 // pinned copy
};

inline void pg_free(PgScratch& S)
{
    cudaFree(S.frames); cudaFree(S.cellcnt); cudaFree(S.cellptr); cudaFree(S.rank); cudaFree(S.cell);
    cudaFree(S.sorted); cudaFree(S.flags);
    if (S.h_flags) cudaFreeHost(S.h_flags);
    S = PgScratch();
}

__device__ __forceinline__ unsigned pg_ordered(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float pg_unordered(unsigned o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void __launch_bounds__(128) pg_init(PgFrame* frames, long long n_frames)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    for (int k = 0; k < 3; k++) { frames[f].bb_lo[k] = 0xffffffffu; frames[f].bb_hi[k] = 0u; }
    frames[f].usable = 0;
}

// bounding box of the listed atoms (two lists: both must lie inside the grid), frames without a cell
__global__ void __launch_bounds__(256) pg_bbox(PgFrame* frames, const float* __restrict__ xyz, int n_atoms,
                                               const int* __restrict__ sel_a, int n_a, const int* __restrict__ sel_b,
                                               int n_b)
{
    const long long f = blockIdx.y;
    const float* x = xyz + (size_t)f * (size_t)n_atoms * 3;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_a + n_b; t += gridDim.x * blockDim.x) {
        const int a = t < n_a ? (sel_a ? __ldg(&sel_a[t]) : t) : __ldg(&sel_b[t - n_a]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = __ldg(&x[(size_t)a * 3 + k]);
            lo[k] = fminf(lo[k], v);      // (NaN is ignored by fminf / fmaxf)
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&frames[f].bb_lo[k], pg_ordered(lo[k]));
            atomicMax(&frames[f].bb_hi[k], pg_ordered(hi[k]));
        }
    }
}

// grid of every frame (one thread per frame).  width > 0: fixed search radius (nearest: the
// cutoff); width <= 0: from the density of the n_sel binned atoms, 1.5 mean spacings.
__global__ void __launch_bounds__(64) pg_setup(PgFrame* frames, const float* __restrict__ box, long long n_frames,
                                               float width, int n_sel, int* __restrict__ cellcnt)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    PgFrame& P = frames[f];
    const float* b = box ? box + (size_t)f * 9 : nullptr;
    float lo[3] = {0.f, 0.f, 0.f}, hi[3] = {0.f, 0.f, 0.f};
    double volume;
    if (b == nullptr) {
        for (int k = 0; k < 3; k++) { lo[k] = pg_unordered(P.bb_lo[k]); hi[k] = pg_unordered(P.bb_hi[k]); }
        volume = fmax((double)hi[0] - lo[0], 0.05) * fmax((double)hi[1] - lo[1], 0.05) * fmax((double)hi[2] - lo[2], 0.05);
    } else {
        volume = fabs((double)b[0] * ((double)b[4] * b[8] - (double)b[5] * b[7]) -
                      (double)b[1] * ((double)b[3] * b[8] - (double)b[5] * b[6]) +
                      (double)b[2] * ((double)b[3] * b[7] - (double)b[4] * b[6]));
    }
    float w = width;
    if (!(w > 0.f)) {
        w = (float)(1.5 * cbrt(volume / (double)(n_sel > 0 ? n_sel : 1)));
        if (!(w >= 0.05f)) w = 0.05f;     // also NaN
        if (w > 1.0e6f) w = 1.0e6f;
    }
    bool ok = volume > 0.0 && volume < 1.0e30;
    if (b == nullptr) for (int k = 0; k < 3; k++) ok = ok && (hi[k] >= lo[k]) && (hi[k] - lo[k] < 1.0e9f);
    // make_grid makes the cells at least 1.002 w + 1e-7 wide; the guaranteed radius is w
    make_grid(P.grid, b, (double)w, PG_MAXC, lo, hi);
    P.width = w;
    P.usable = ok && P.grid.distinct && P.grid.ncell >= 1;
    (void)cellcnt;
}

__global__ void __launch_bounds__(256) pg_zero(const PgFrame* __restrict__ frames, int* __restrict__ cellcnt)
{
    const long long f = blockIdx.y;
    const int ncell = frames[f].grid.ncell;
    int* cnt = cellcnt + (size_t)f * (PG_MAXC + 1);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= ncell; c += gridDim.x * blockDim.x) cnt[c] = 0;
}

// sel == nullptr: every atom
__global__ void __launch_bounds__(256) pg_bin(const PgFrame* __restrict__ frames, const float* __restrict__ xyz,
                                              int n_atoms, const int* __restrict__ sel, int n_sel, int cap_sel,
                                              int* __restrict__ cellcnt, int* __restrict__ rank, int* __restrict__ cell)
{
    const long long f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sel) return;
    const int a = sel ? __ldg(&sel[s]) : s;
    const float* p = xyz + ((size_t)f * (size_t)n_atoms + (size_t)a) * 3;
    float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    const int c = cell_wrap(frames[f].grid, x, y, z);
    cell[(size_t)f * cap_sel + s] = c;
    rank[(size_t)f * cap_sel + s] = atomicAdd(&cellcnt[(size_t)f * (PG_MAXC + 1) + c], 1);
}

// exclusive scan over the cells of one frame per CTA
__global__ void __launch_bounds__(1024) pg_scan(const PgFrame* __restrict__ frames, const int* __restrict__ cellcnt,
                                                int* __restrict__ cellptr, int n_sel)
{
    __shared__ int wsum[32];
    __shared__ int carry;
    const long long f = blockIdx.x;
    const int ncell = frames[f].grid.ncell;
    const int* cnt = cellcnt + (size_t)f * (PG_MAXC + 1);
    int* ptr = cellptr + (size_t)f * (PG_MAXC + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int c = base + threadIdx.x;
        const int v = c < ncell ? cnt[c] : 0;
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
        if (c < ncell) ptr[c] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[ncell] = n_sel;
}

__global__ void __launch_bounds__(256) pg_scatter(const float* __restrict__ xyz, int n_atoms,
                                                  const int* __restrict__ sel, int n_sel, int cap_sel,
                                                  const int* __restrict__ cellptr, const int* __restrict__ rank,
                                                  const int* __restrict__ cell, float4* __restrict__ sorted)
{
    const long long f = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sel) return;
    const int a = sel ? __ldg(&sel[s]) : s;
    const float* p = xyz + ((size_t)f * (size_t)n_atoms + (size_t)a) * 3;
    const int pos = cellptr[(size_t)f * (PG_MAXC + 1) + cell[(size_t)f * cap_sel + s]] + rank[(size_t)f * cap_sel + s];
    sorted[(size_t)f * cap_sel + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __int_as_float(s));
}

// the cell ranges of the 27-cell stencil around cell (cx, cy, cz): calls visit(first, last)
template <class F>
__device__ __forceinline__ void pg_stencil(const GridS& G, const int* __restrict__ ptr, int cx, int cy, int cz, F visit)
{
    for (int dz = -1; dz <= 1; dz++) {
        int z = cz + dz;
        if (z < 0) { if (!G.periodic) continue; z += G.nz; } else if (z >= G.nz) { if (!G.periodic) continue; z -= G.nz; }
        for (int dy = -1; dy <= 1; dy++) {
            int y = cy + dy;
            if (y < 0) { if (!G.periodic) continue; y += G.ny; } else if (y >= G.ny) { if (!G.periodic) continue; y -= G.ny; }
            // the three x cells are contiguous unless the row wraps
            int x0 = cx - 1, x1 = cx + 1;
            const int row = (z * G.ny + y) * G.nx;
            if (x0 < 0 || x1 >= G.nx) {
                for (int dx = -1; dx <= 1; dx++) {
                    int x = cx + dx;
                    if (x < 0) { if (!G.periodic) continue; x += G.nx; } else if (x >= G.nx) { if (!G.periodic) continue; x -= G.nx; }
                    visit(ptr[row + x], ptr[row + x + 1]);
                }
            } else {
                visit(ptr[row + x0], ptr[row + x1 + 1]);
            }
        }
    }
}

// ---- NearestAtoms, one frame (frame slice 0 of the scratch): one WARP per atom, the lanes share
//      the candidates of every stencil range (coalesced 16-byte loads of the cell-sorted array) ----
__global__ void __launch_bounds__(PG_THREADS) k_nearest_cells(
    const PgFrame* __restrict__ frames, const float* __restrict__ x, const float* __restrict__ box9, int n_atoms,
    float c2, int orthogonal, int excl_mode, const int* __restrict__ atom_res,
    const long long* __restrict__ excl_ptr, const int* __restrict__ excl_idx, const int* __restrict__ cellptr,
    const float4* __restrict__ sorted, int* __restrict__ nearest, float* __restrict__ dist)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (PG_THREADS / 32) + (threadIdx.x >> 5);
    if (i >= n_atoms) return;
    const GridS& G = frames[0].grid;
    const NlBox B = make_nlbox(box9);
    const DistBox D = make_distbox(box9, orthogonal);
    const float xi = __ldg(&x[(size_t)i * 3]), yi = __ldg(&x[(size_t)i * 3 + 1]), zi = __ldg(&x[(size_t)i * 3 + 2]);
    const int ri = excl_mode == 1 ? __ldg(&atom_res[i]) : -1;
    long long e0 = 0, e1 = 0;
    if (excl_mode == 2) { e0 = __ldg(&excl_ptr[i]); e1 = __ldg(&excl_ptr[i + 1]); }
    float wx = xi, wy = yi, wz = zi;
    const int c = cell_wrap(G, wx, wy, wz);
    const int cyz = c / G.nx, cx = c - cyz * G.nx, cz = cyz / G.ny, cy = cyz - cz * G.ny;
    float best = __int_as_float(0x7f800000);
    int bj = 0x7fffffff;
    pg_stencil(G, cellptr, cx, cy, cz, [&](int first, int last) {
        for (int p = first + lane; p < last; p += 32) {
            const float4 q = __ldg(&sorted[p]);
            const int j = __float_as_int(q.w);
            // the neighbour-list predicate of the ordered pair i -> j, then the compute_distances value
            const float d2 = nl_d2_dyn(xi, yi, zi, q.x, q.y, q.z, B);
            if (d2 > c2 || j == i) continue;
            bool allowed = true;
            if (excl_mode == 1) allowed = (__ldg(&atom_res[j]) != ri);
            else if (excl_mode == 2) {
                for (long long e = e0; e < e1; e++)
                    if (__ldg(&excl_idx[e]) == j) { allowed = false; break; }
            }
            if (!allowed) continue;
            const float d = __fsqrt_rn(dist2_pair(xi, yi, zi, q.x, q.y, q.z, D));
            if (d < best || (d == best && j < bj)) { best = d; bj = j; }
        }
    });
    // smallest distance, ties -> lower atom index (the first entry of the ascending neighbour list)
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, s);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, s);
        if (od < best || (od == best && oj < bj)) { best = od; bj = oj; }
    }
    if (lane == 0) {
        nearest[i] = bj == 0x7fffffff ? -1 : bj;
        dist[i] = best;
    }
}

// ---- MinimumDistanceCounter over query x haystack: stencil search, one thread per query atom ----
// grid = (query blocks, frames); partial[(f * gridDim.x + block)] = best of the block, index =
// position in the query-major product (numpy argmin: first index among equal float32 distances)
__global__ void __launch_bounds__(PG_THREADS) k_min_dist_cells(
    const PgFrame* __restrict__ frames, const float* __restrict__ xyz, const float* __restrict__ box, int n_atoms,
    const int* __restrict__ query, int n_q, int n_h, int cap_sel, int orthogonal, const int* __restrict__ cellptr,
    const float4* __restrict__ sorted, float* __restrict__ part_d, long long* __restrict__ part_idx)
{
    __shared__ float red_d[PG_THREADS / 32];
    __shared__ long long red_i[PG_THREADS / 32];
    const long long f = blockIdx.y;
    const PgFrame& P = frames[f];
    float best = __int_as_float(0x7f800000);
    long long bidx = 0x7fffffffffffffffll;
    const int t = blockIdx.x * PG_THREADS + threadIdx.x;
    if (P.usable && t < n_q) {
        const GridS& G = P.grid;
        const DistBox D = make_distbox(box ? box + (size_t)f * 9 : nullptr, orthogonal);
        const float* p = xyz + ((size_t)f * (size_t)n_atoms + (size_t)__ldg(&query[t])) * 3;
        const float qx = __ldg(p), qy = __ldg(p + 1), qz = __ldg(p + 2);
        float wx = qx, wy = qy, wz = qz;
        const int c = cell_wrap(G, wx, wy, wz);
        const int cyz = c / G.nx, cx = c - cyz * G.nx, cz = cyz / G.ny, cy = cyz - cz * G.ny;
        const int* ptr = cellptr + (size_t)f * (PG_MAXC + 1);
        const float4* srt = sorted + (size_t)f * cap_sel;
        pg_stencil(G, ptr, cx, cy, cz, [&](int first, int last) {
            for (int s = first; s < last; s++) {
                const float4 h = __ldg(&srt[s]);
                // compute_distances(pair = (query, haystack)): r = p_haystack - p_query
                const float d = __fsqrt_rn(dist2_pair(qx, qy, qz, h.x, h.y, h.z, D));
                const long long idx = (long long)t * (long long)n_h + (long long)__float_as_int(h.w);
                if (md_better(d, idx, best, bidx)) { best = d; bidx = idx; }
            }
        });
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, s);
        const long long oi = __shfl_xor_sync(0xffffffffu, bidx, s);
        if (md_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { red_d[threadIdx.x >> 5] = best; red_i[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < PG_THREADS / 32; w++)
            if (md_better(red_d[w], red_i[w], best, bidx)) { best = red_d[w]; bidx = red_i[w]; }
        part_d[(size_t)f * gridDim.x + blockIdx.x] = best;
        part_idx[(size_t)f * gridDim.x + blockIdx.x] = bidx;
    }
}

// accept the stencil minimum of a frame iff it is inside the guaranteed radius (see the header)
__global__ void __launch_bounds__(128) pg_flag(const PgFrame* __restrict__ frames, const float* __restrict__ d_min,
                                               long long n_frames, int* __restrict__ flags, int* __restrict__ n_flagged)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const bool ok = frames[f].usable && d_min[f] <= 0.9f * frames[f].width;      // false for +inf / NaN
    flags[f] = ok ? 0 : 1;
    if (!ok) atomicAdd(n_flagged, 1);
}

#define PG_CK(call)                                                                  \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) {                                                     \
            err = std::string(#call) + " failed: " + cudaGetErrorString(e_);         \
            return -2;                                                               \
        }                                                                            \
    } while (0)

inline int pg_ensure(PgScratch& S, long long n_frames, int n_sel, std::string& err)
{
    if (S.cap_frames < n_frames || S.cap_sel < n_sel) {
        const long long frames = std::max(n_frames, S.cap_frames);
        const int sel = std::max(n_sel, S.cap_sel);
        pg_free(S);
        PG_CK(cudaMalloc(&S.frames, sizeof(PgFrame) * (size_t)frames));
        PG_CK(cudaMalloc(&S.cellcnt, sizeof(int) * (size_t)frames * (PG_MAXC + 1)));
        PG_CK(cudaMalloc(&S.cellptr, sizeof(int) * (size_t)frames * (PG_MAXC + 1)));
        PG_CK(cudaMalloc(&S.rank, sizeof(int) * (size_t)frames * (size_t)sel));
        PG_CK(cudaMalloc(&S.cell, sizeof(int) * (size_t)frames * (size_t)sel));
        PG_CK(cudaMalloc(&S.sorted, sizeof(float4) * (size_t)frames * (size_t)sel));
        PG_CK(cudaMalloc(&S.flags, sizeof(int) * (size_t)(frames + 1)));
        PG_CK(cudaMallocHost(&S.h_flags, sizeof(int) * (size_t)(frames + 1)));
        S.cap_frames = frames;
        S.cap_sel = sel;
    }
    return 0;
}

// frames a batch may hold: ~256 MB of scratch
inline long long pg_batch_frames(int n_sel)
{
    const size_t per_frame = 2 * sizeof(int) * (size_t)(PG_MAXC + 1) + (size_t)n_sel * 24 + sizeof(PgFrame);
    return std::max<long long>(1, std::min<long long>(32768, (long long)((256ull << 20) / per_frame)));
}

// bin the selected atoms of `n_frames` frames (asynchronous)
inline int pg_bin_frames(PgScratch& S, const float* xyz, const float* box, long long n_frames, int n_atoms,
                         const int* sel, int n_sel, const int* other, int n_other, float width, cudaStream_t st,
                         std::string& err, int& launches)
{
    const unsigned nf = (unsigned)n_frames;
    pg_init<<<(nf + 127) / 128, 128, 0, st>>>(S.frames, n_frames);
    if (box == nullptr) {
        const int blocks = std::max(1, std::min((n_sel + n_other + 255) / 256, 32));
        pg_bbox<<<dim3(blocks, nf), 256, 0, st>>>(S.frames, xyz, n_atoms, sel, n_sel, other, n_other);
        launches++;
    }
    pg_setup<<<(nf + 63) / 64, 64, 0, st>>>(S.frames, box, n_frames, width, n_sel, S.cellcnt);
    pg_zero<<<dim3(16, nf), 256, 0, st>>>(S.frames, S.cellcnt);
    pg_bin<<<dim3((n_sel + 255) / 256, nf), 256, 0, st>>>(S.frames, xyz, n_atoms, sel, n_sel, S.cap_sel, S.cellcnt,
                                                          S.rank, S.cell);
    pg_scan<<<nf, 1024, 0, st>>>(S.frames, S.cellcnt, S.cellptr, n_sel);
    pg_scatter<<<dim3((n_sel + 255) / 256, nf), 256, 0, st>>>(xyz, n_atoms, sel, n_sel, S.cap_sel, S.cellptr, S.rank,
                                                              S.cell, S.sorted);
    PG_CK(cudaGetLastError());
    launches += 6;
    return 0;
}

}  // namespace cmb
