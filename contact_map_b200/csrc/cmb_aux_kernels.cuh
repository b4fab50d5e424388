// cmb_aux_kernels.cuh -- export / gather / synthetic-generator / min-distance kernels.
#pragma once
#include "cmb_math.cuh"

namespace cmb {

// ---------------------------------------------------------------------------------
// hashed count tables (cmb_frames_large.cuh): sparse export and growth
// ---------------------------------------------------------------------------------
// PACKED: out_idx receives (flat index << 24 | count) and out_cnt is unused (cmb_export_sorted)
template <bool PACKED>
__global__ void __launch_bounds__(256) k_export_hashed(const unsigned long long* __restrict__ slots,
                                                       long long n_slots, long long cap,
                                                       long long* __restrict__ out_idx,
                                                       unsigned int* __restrict__ out_cnt,
                                                       unsigned long long* __restrict__ out_n)
{
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long warp_first = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane;
    for (long long v0 = warp_first; v0 < n_slots; v0 += stride) {
        const long long v = v0 + lane;
        const unsigned long long s = v < n_slots ? __ldg(&slots[v]) : 0ull;
        const bool have = (s & 0xffffffull) != 0ull;            // a claimed slot with count 0 is no contact
        const unsigned m = __ballot_sync(0xffffffffu, have);
        if (m == 0u) continue;
        unsigned long long base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(out_n, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (have) {
            const unsigned long long w = base + __popc(m & ((1u << lane) - 1u));
            if ((long long)w < cap) {
                if (PACKED) {
                    out_idx[w] = (long long)(s - (1ull << 24));
                } else {
                    out_idx[w] = (long long)((s >> 24) - 1ull);
                    out_cnt[w] = (unsigned int)(s & 0xffffffull);
                }
            }
        }
    }
}

// re-insert every entry of `src` into the (larger, zero-filled) table `dst`
__global__ void __launch_bounds__(256) k_rehash(const unsigned long long* __restrict__ src, long long n_src,
                                                unsigned long long* __restrict__ dst, unsigned dst_mask,
                                                unsigned long long* __restrict__ stats)
{
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n_src;
         v += (long long)gridDim.x * blockDim.x) {
        const unsigned long long s = __ldg(&src[v]);
        if (s == 0ull) continue;
        const unsigned long long key1 = s >> 24;
        unsigned long long k = key1;
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
        unsigned h = ((unsigned)k & dst_mask) & ~3u;         // bucketed probe sequence (ht_find_or_insert)
        bool placed = false;
        for (int probe = 0; probe < (1 << 14) && !placed; probe++) {
            const unsigned long long old = atomicCAS(&dst[h], 0ull, s);      // keys are distinct in src
            if (old == 0ull) placed = true;
            else h = (h + 1u) & dst_mask;
        }
        if (!placed) atomicAdd(&stats[3], 1ull);
    }
}

// slots[key1] += count for a list of distinct keys (parked hits after a table was grown)
__global__ void __launch_bounds__(256) k_hash_add(unsigned long long* __restrict__ slots, unsigned mask,
                                                  const unsigned long long* __restrict__ key1,
                                                  const unsigned long long* __restrict__ count, long long n,
                                                  unsigned long long* __restrict__ stats, int which)
{
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (long long)gridDim.x * blockDim.x) {
        const unsigned long long k1 = key1[v];
        unsigned long long k = k1;
        k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
        unsigned h = ((unsigned)k & mask) & ~3u;
        bool placed = false;
        for (int probe = 0; probe < (1 << 14) && !placed; probe++) {
            unsigned long long cur = *(volatile unsigned long long*)&slots[h];
            if (cur == 0ull) {
                const unsigned long long old = atomicCAS(&slots[h], 0ull, k1 << 24);
                if (old == 0ull) atomicAdd(&stats[which], 1ull);
                cur = old == 0ull ? (k1 << 24) : old;
            }
            if ((cur >> 24) == k1) {
                atomicAdd(&slots[h], count[v]);
                placed = true;
            } else {
                h = (h + 1u) & mask;
            }
        }
        if (!placed) atomicAdd(&stats[3], 1ull);
    }
}

// ---------------------------------------------------------------------------------
// sorted per-frame keys -> CSR arrays (the container of contact_trajectory.py:7-44 as arrays):
// ptr[f] = first key of frame f (ptr[n_frames] = n_keys), i / j = the pair in REAL indices
// (index_map[compact index], or the compact index itself).  One thread per key; the thread of
// the first key of a frame also fills the ptr entries of the empty frames before it.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_keys_to_csr(const unsigned long long* __restrict__ keys, long long n_keys,
                                                     long long n_frames, long long frame0,
                                                     const int* __restrict__ index_map, long long* __restrict__ ptr,
                                                     int* __restrict__ out_i, int* __restrict__ out_j)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0 && n_keys == 0)
        for (long long f = 0; f <= n_frames; f++) ptr[f] = 0;
    if (k >= n_keys) return;
    const unsigned long long key = keys[k];
    long long f = (long long)(key >> 40) - frame0;
    f = f < 0 ? 0 : (f >= n_frames ? n_frames - 1 : f);          // (keys of other frames cannot occur)
    long long fprev = -1;
    if (k > 0) {
        fprev = (long long)(keys[k - 1] >> 40) - frame0;
        fprev = fprev < 0 ? 0 : (fprev >= n_frames ? n_frames - 1 : fprev);
    }
    for (long long g = fprev + 1; g <= f; g++) ptr[g] = k;
    if (k == n_keys - 1)
        for (long long g = f + 1; g <= n_frames; g++) ptr[g] = n_keys;
    const int lo = (int)((key >> 20) & 0xfffffull), hi = (int)(key & 0xfffffull);
    int a = lo, b = hi;
    if (index_map != nullptr) { a = __ldg(&index_map[lo]); b = __ldg(&index_map[hi]); }
    out_i[k] = a < b ? a : b;
    out_j[k] = a < b ? b : a;
}

// ---------------------------------------------------------------------------------
// sparse export of a uint32 count table (result container of contact_map.py:724-744)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_export_nonzero(const unsigned int* __restrict__ table,
                                                        long long n_elems, long long cap,
                                                        long long* __restrict__ out_idx,
                                                        unsigned int* __restrict__ out_cnt,
                                                        unsigned long long* __restrict__ out_n)
{
    const int lane = threadIdx.x & 31;
    const long long nvec = n_elems >> 2;
    const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(table);
    const long long stride = (long long)gridDim.x * blockDim.x;
    // uniform trip count per warp so the ballots below are legal
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long warp_first = first - lane;
    for (long long v0 = warp_first; v0 < nvec; v0 += stride) {
        const long long v = v0 + lane;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (v < nvec) q = __ldg(&t4[v]);
        const unsigned vals[4] = {q.x, q.y, q.z, q.w};
        const int mine = (q.x != 0) + (q.y != 0) + (q.z != 0) + (q.w != 0);
        // warp-aggregated reservation
        int incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out_n, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        unsigned long long w = base + (unsigned long long)(incl - mine);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (vals[k]) {
                if ((long long)w < cap) {
                    out_idx[w] = v * 4 + k;
                    out_cnt[w] = vals[k];
                }
                w++;
            }
        }
    }
    // tail (n_elems % 4), handled by one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (long long e = nvec * 4; e < n_elems; e++) {
            const unsigned val = table[e];
            if (val) {
                const unsigned long long w = atomicAdd(out_n, 1ull);
                if ((long long)w < cap) { out_idx[w] = e; out_cnt[w] = val; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// ordered sparse export of a dense table in two passes (no atomics, no sort): every warp owns a
// contiguous span of the table; pass 1 counts its non-zero entries, pass 2 turns the counts into
// offsets (each block sums the counts of the warps before it) and writes the packed
// (flat index << 24 | count) words in index order.
// ---------------------------------------------------------------------------------
constexpr int EO_THREADS = 256;

__device__ __forceinline__ int eo_count4(const uint4& q) { return (q.x != 0) + (q.y != 0) + (q.z != 0) + (q.w != 0); }

// `flags`: one bit per warp iteration (32 vectors = 512 bytes of table), one word per 32
// iterations of a warp's span (span_vec is a multiple of 1024 vectors): the write pass only
// revisits the iterations that hold something -- a 40 GB table with 3 M entries is read ONCE.
__global__ void __launch_bounds__(EO_THREADS) k_export_count(const unsigned int* __restrict__ table, long long n_elems,
                                                             long long span_vec, int* __restrict__ counts,
                                                             unsigned int* __restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (EO_THREADS / 32) + (threadIdx.x >> 5);
    const long long nvec = n_elems >> 2;
    const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(table);
    const long long v0 = warp * span_vec, v1 = min(nvec, v0 + span_vec);
    const long long words_per_warp = span_vec / 1024;
    int mine = 0;
    for (long long g = 0; g < words_per_warp; g++) {
        const long long gb = v0 + g * 1024;
        if (gb >= v1) { if (lane == 0) flags[warp * words_per_warp + g] = 0u; continue; }
        unsigned word = 0u;
#pragma unroll 1
        for (int i0 = 0; i0 < 32; i0 += 8) {
            int c[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {               // eight independent 16-byte loads in flight per thread
                const long long v = gb + (long long)(i0 + k) * 32 + lane;
                c[k] = v < v1 ? eo_count4(__ldg(&t4[v])) : 0;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                mine += c[k];
                if (__any_sync(0xffffffffu, c[k] != 0)) word |= 1u << (i0 + k);
            }
        }
        if (lane == 0) flags[warp * words_per_warp + g] = word;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if (lane == 0) counts[warp] = mine;
}

__global__ void __launch_bounds__(EO_THREADS) k_export_write(const unsigned int* __restrict__ table, long long n_elems,
                                                             long long span_vec, const int* __restrict__ counts,
                                                             const unsigned int* __restrict__ flags,
                                                             long long cap, unsigned long long* __restrict__ out,
                                                             unsigned long long* __restrict__ out_n)
{
    __shared__ long long red[EO_THREADS / 32];
    __shared__ long long block_base;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long first_warp = (long long)blockIdx.x * (EO_THREADS / 32);
    const long long n_warps = (long long)gridDim.x * (EO_THREADS / 32);
    // offset of this block's first warp: sum of the counts of all earlier warps
    long long part = 0;
    for (long long w = threadIdx.x; w < first_warp; w += EO_THREADS) part += counts[w];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) red[wib] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long b = 0;
        for (int k = 0; k < EO_THREADS / 32; k++) b += red[k];
        block_base = b;
    }
    __syncthreads();
    long long base = block_base;
    for (int k = 0; k < wib; k++) base += counts[first_warp + k];

    const long long warp = first_warp + wib;
    const long long nvec = n_elems >> 2;
    const uint4* __restrict__ t4 = reinterpret_cast<const uint4*>(table);
    const long long v0 = warp * span_vec, v1 = min(nvec, v0 + span_vec);
    const long long words_per_warp = span_vec / 1024;
    bool wide = false;
    if (counts[warp] != 0) {
        for (long long g = 0; g < words_per_warp; g++) {
            unsigned word = __ldg(&flags[warp * words_per_warp + g]);      // uniform over the warp
            while (word) {
                const int it = __ffs(word) - 1;
                word &= word - 1;
                const long long v = v0 + g * 1024 + (long long)it * 32 + lane;
                uint4 q = make_uint4(0u, 0u, 0u, 0u);
                if (v < v1) q = __ldg(&t4[v]);
                const unsigned vals[4] = {q.x, q.y, q.z, q.w};
                const int mine = eo_count4(q);
                int incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += u;
                }
                long long w = base + (incl - mine);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (vals[k]) {
                        if (w < cap) out[w] = ((unsigned long long)(v * 4 + k) << 24) | (unsigned long long)(vals[k] & 0xffffffu);
                        wide = wide || (vals[k] >> 24);
                        w++;
                    }
                }
                base += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
    }
    if (warp == n_warps - 1 && lane == 0) {
        // tail (n_elems % 4) and the total
        for (long long e = nvec * 4; e < n_elems; e++) {
            const unsigned val = table[e];
            if (val) {
                if (base < cap) out[base] = ((unsigned long long)e << 24) | (unsigned long long)(val & 0xffffffu);
                wide = wide || (val >> 24);
                base++;
            }
        }
        out_n[0] = (unsigned long long)base;
    }
    if (wide) out_n[1] = 1ull;
}

// ---------------------------------------------------------------------------------
// _atom_slice (atom_indexer.py:5-22) on device
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_atoms(const float* __restrict__ full, long long n_frames,
                                                      int n_total, const int* __restrict__ used, int n_used,
                                                      float* __restrict__ out)
{
    const long long total = n_frames * (long long)n_used;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long f = t / n_used;
        const int i = (int)(t - f * n_used);
        const float* src = full + ((size_t)f * n_total + (size_t)__ldg(&used[i])) * 3;
        float* dst = out + (size_t)t * 3;
        dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);
    }
}

// ---------------------------------------------------------------------------------
// synthetic generator -- must match contact_map_b200/synth.py bit for bit
// ---------------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
#define CMB_P1 0xD6E8FEB86659FD93ull
#define CMB_P2 0xA24BAED4963EE407ull
#define CMB_P3 0x9FB21C651E98DF25ull

__global__ void __launch_bounds__(256) k_synth_frames(const float* __restrict__ base,
                                                      const int* __restrict__ group, int n_atoms,
                                                      const float* __restrict__ box9,
                                                      unsigned long long seed, float amplitude,
                                                      int image_shifts, long long frame0,
                                                      long long n_frames, float* __restrict__ out)
{
    const long long total = n_frames * (long long)n_atoms;
    float bv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (box9 != nullptr)
        for (int k = 0; k < 9; k++) bv[k] = __ldg(&box9[k]);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long fl = t / n_atoms;
        const int a = (int)(t - fl * n_atoms);
        const unsigned long long f = (unsigned long long)(frame0 + fl);
        float p[3];
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const unsigned long long h = splitmix64(seed ^ (f * CMB_P1) ^ ((unsigned long long)a * CMB_P2) ^
                                                    ((unsigned long long)ax * CMB_P3));
            const float u = (float)(h >> 40) * 5.9604644775390625e-08f;   // 2^-24, exact
            const float w = __fsub_rn(__fmul_rn(2.0f, u), 1.0f);
            p[ax] = __fadd_rn(__ldg(&base[(size_t)a * 3 + ax]), __fmul_rn(amplitude, w));
        }
        if (image_shifts && box9 != nullptr) {
            const unsigned long long g = (unsigned long long)__ldg(&group[a]);
#pragma unroll
            for (int v = 0; v < 3; v++) {
                const unsigned long long h = splitmix64(seed ^ (f * CMB_P1) ^ (g * CMB_P2) ^
                                                        ((unsigned long long)(v + 3) * CMB_P3));
                const int m = (int)(((h >> 40) * 3ull) >> 24) - 1;
                const float fm = (float)m;
#pragma unroll
                for (int ax = 0; ax < 3; ax++) p[ax] = __fadd_rn(p[ax], __fmul_rn(fm, bv[v * 3 + ax]));
            }
        }
        float* dst = out + (size_t)t * 3;
        dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
    }
}

// ---------------------------------------------------------------------------------
// MinimumDistanceCounter (min_dist.py:142-168)
// ---------------------------------------------------------------------------------
constexpr int MD_THREADS = 256;
constexpr int MD_HPT = 4;                          // haystack atoms per thread (registers)
constexpr int MD_HCHUNK = MD_THREADS * MD_HPT;     // haystack atoms per CTA
constexpr int MD_QTILE = 1024;                     // query atoms staged in shared memory

struct MinRec { float d; long long idx; };

__device__ __forceinline__ bool md_better(float d, long long idx, float bd, long long bidx)
{
    return (d < bd) || (d == bd && idx < bidx);
}

// grid = (n_hchunks, n_frames).  partial[(f * n_hchunks + chunk)] = best of that chunk.
__global__ void __launch_bounds__(MD_THREADS) k_min_dist_partial(
    const float* __restrict__ xyz, const float* __restrict__ box, int n_atoms,
    const int* __restrict__ query, int n_q, const int* __restrict__ hay, int n_h, int orthogonal,
    float* __restrict__ part_d, long long* __restrict__ part_idx)
{
    __shared__ float qs[MD_QTILE * 3];
    __shared__ float red_d[MD_THREADS / 32];
    __shared__ long long red_i[MD_THREADS / 32];
    const long long f = blockIdx.y;
    const int chunk = blockIdx.x;
    const float* x = xyz + (size_t)f * (size_t)n_atoms * 3;
    const DistBox D = make_distbox(box ? box + (size_t)f * 9 : nullptr, orthogonal);

    float hx[MD_HPT], hy[MD_HPT], hz[MD_HPT];
    int hpos[MD_HPT];
#pragma unroll
    for (int k = 0; k < MD_HPT; k++) {
        const int j = chunk * MD_HCHUNK + k * MD_THREADS + threadIdx.x;
        hpos[k] = j < n_h ? j : -1;
        hx[k] = hy[k] = hz[k] = 0.f;
        if (j < n_h) {
            const int a = __ldg(&hay[j]);
            hx[k] = __ldg(&x[(size_t)a * 3]); hy[k] = __ldg(&x[(size_t)a * 3 + 1]); hz[k] = __ldg(&x[(size_t)a * 3 + 2]);
        }
    }
    float best = __int_as_float(0x7f800000);   // +inf
    long long bidx = 0x7fffffffffffffffll;
    for (int q0 = 0; q0 < n_q; q0 += MD_QTILE) {
        const int qn = min(MD_QTILE, n_q - q0);
        __syncthreads();
        for (int t = threadIdx.x; t < qn; t += MD_THREADS) {
            const int a = __ldg(&query[q0 + t]);
            qs[3 * t] = __ldg(&x[(size_t)a * 3]); qs[3 * t + 1] = __ldg(&x[(size_t)a * 3 + 1]); qs[3 * t + 2] = __ldg(&x[(size_t)a * 3 + 2]);
        }
        __syncthreads();
        for (int t = 0; t < qn; t++) {
            const float qx = qs[3 * t], qy = qs[3 * t + 1], qz = qs[3 * t + 2];
#pragma unroll
            for (int k = 0; k < MD_HPT; k++) {
                // compute_distances(pair = (query, haystack)): r = p_haystack - p_query
                const float d = __fsqrt_rn(dist2_pair(qx, qy, qz, hx[k], hy[k], hz[k], D));
                const long long idx = (long long)(q0 + t) * (long long)n_h + (long long)hpos[k];
                if (hpos[k] >= 0 && md_better(d, idx, best, bidx)) { best = d; bidx = idx; }
            }
        }
    }
    // block reduction
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, s);
        const long long oi = __shfl_xor_sync(0xffffffffu, bidx, s);
        if (md_better(od, oi, best, bidx)) { best = od; bidx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { red_d[threadIdx.x >> 5] = best; red_i[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < MD_THREADS / 32; w++)
            if (md_better(red_d[w], red_i[w], best, bidx)) { best = red_d[w]; bidx = red_i[w]; }
        part_d[(size_t)f * gridDim.x + chunk] = best;
        part_idx[(size_t)f * gridDim.x + chunk] = bidx;
    }
}

// MinimumDistanceCounter._compute_from_atom_pairs (min_dist.py:142-168) for an ARBITRARY pair
// list: grid = (n_chunks, n_frames), chunk c owns pairs [c * MD_PCHUNK, (c + 1) * MD_PCHUNK);
// partial[(f * n_chunks + chunk)] = best of that chunk, index = position in the pair list
// (numpy argmin: first index among equal float32 distances).
constexpr int MD_PCHUNK = MD_THREADS * 16;

__global__ void __launch_bounds__(MD_THREADS) k_min_dist_pairs_partial(
    const float* __restrict__ xyz, const float* __restrict__ box, int n_atoms,
    const int2* __restrict__ pairs, long long n_pairs, int orthogonal,
    float* __restrict__ part_d, long long* __restrict__ part_idx)
{
    __shared__ float red_d[MD_THREADS / 32];
    __shared__ long long red_i[MD_THREADS / 32];
    const long long f = blockIdx.y;
    const float* x = xyz + (size_t)f * (size_t)n_atoms * 3;
    const DistBox D = make_distbox(box ? box + (size_t)f * 9 : nullptr, orthogonal);
    float best = __int_as_float(0x7f800000);
    long long bidx = 0x7fffffffffffffffll;
    const long long p0 = (long long)blockIdx.x * MD_PCHUNK;
    const long long p1 = min(n_pairs, p0 + (long long)MD_PCHUNK);
    for (long long p = p0 + threadIdx.x; p < p1; p += MD_THREADS) {
        const int2 pr = __ldg(&pairs[p]);
        const float* a = x + (size_t)pr.x * 3;
        const float* b = x + (size_t)pr.y * 3;
        // compute_distances(pair = (i, j)): r = p_j - p_i
        const float d = __fsqrt_rn(dist2_pair(__ldg(a), __ldg(a + 1), __ldg(a + 2), __ldg(b), __ldg(b + 1),
                                              __ldg(b + 2), D));
        if (md_better(d, p, best, bidx)) { best = d; bidx = p; }
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
        for (int w = 1; w < MD_THREADS / 32; w++)
            if (md_better(red_d[w], red_i[w], best, bidx)) { best = red_d[w]; bidx = red_i[w]; }
        part_d[(size_t)f * gridDim.x + blockIdx.x] = best;
        part_idx[(size_t)f * gridDim.x + blockIdx.x] = bidx;
    }
}

__global__ void __launch_bounds__(128) k_min_dist_final(const float* __restrict__ part_d,
                                                        const long long* __restrict__ part_idx,
                                                        int n_chunks, long long n_frames,
                                                        float* __restrict__ out_d,
                                                        long long* __restrict__ out_idx)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    float best = part_d[(size_t)f * n_chunks];
    long long bidx = part_idx[(size_t)f * n_chunks];
    for (int c = 1; c < n_chunks; c++) {
        const float d = part_d[(size_t)f * n_chunks + c];
        const long long i = part_idx[(size_t)f * n_chunks + c];
        if (md_better(d, i, best, bidx)) { best = d; bidx = i; }
    }
    out_d[f] = best;
    out_idx[f] = bidx;
}

// ---------------------------------------------------------------------------------
// NearestAtoms (min_dist.py:47-70), one frame, tiled brute force
// ---------------------------------------------------------------------------------
constexpr int NA_THREADS = 256;
constexpr int NA_TILE = 1024;

__global__ void __launch_bounds__(NA_THREADS) k_nearest(
    const float* __restrict__ x, const float* __restrict__ box9, int n_atoms, float c2,
    int orthogonal, int excl_mode, const int* __restrict__ atom_res,
    const long long* __restrict__ excl_ptr, const int* __restrict__ excl_idx,
    int* __restrict__ nearest, float* __restrict__ dist)
{
    __shared__ float4 tile[NA_TILE];
    const NlBox B = make_nlbox(box9);
    const DistBox D = make_distbox(box9, orthogonal);
    const int i = blockIdx.x * NA_THREADS + threadIdx.x;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    int ri = -1;
    long long e0 = 0, e1 = 0;
    if (i < n_atoms) {
        xi = __ldg(&x[(size_t)i * 3]); yi = __ldg(&x[(size_t)i * 3 + 1]); zi = __ldg(&x[(size_t)i * 3 + 2]);
        if (excl_mode == 1) ri = __ldg(&atom_res[i]);
        if (excl_mode == 2) { e0 = __ldg(&excl_ptr[i]); e1 = __ldg(&excl_ptr[i + 1]); }
    }
    float best = __int_as_float(0x7f800000);
    int bj = -1;
    for (int j0 = 0; j0 < n_atoms; j0 += NA_TILE) {
        const int jn = min(NA_TILE, n_atoms - j0);
        __syncthreads();
        for (int t = threadIdx.x; t < jn; t += NA_THREADS) {
            const int j = j0 + t;
            const int rj = (excl_mode == 1) ? __ldg(&atom_res[j]) : 0;
            tile[t] = make_float4(__ldg(&x[(size_t)j * 3]), __ldg(&x[(size_t)j * 3 + 1]),
                                  __ldg(&x[(size_t)j * 3 + 2]), __int_as_float(rj));
        }
        __syncthreads();
        if (i < n_atoms) {
            for (int t = 0; t < jn; t++) {
                const float4 q = tile[t];
                const int j = j0 + t;
                const float d2 = nl_d2_dyn(xi, yi, zi, q.x, q.y, q.z, B);
                if (d2 > c2 || j == i) continue;
                bool allowed = true;
                if (excl_mode == 1) allowed = (__float_as_int(q.w) != ri);
                else if (excl_mode == 2) {
                    // `n not in excluded[atom]`: note a list that omits the atom itself
                    // never matters because the neighbour list has no self entry
                    for (long long e = e0; e < e1; e++)
                        if (__ldg(&excl_idx[e]) == j) { allowed = false; break; }
                }
                if (!allowed) continue;
                const float d = __fsqrt_rn(dist2_pair(xi, yi, zi, q.x, q.y, q.z, D));
                if (d < best || (d == best && j < bj)) { best = d; bj = j; }
            }
        }
    }
    if (i < n_atoms) { nearest[i] = bj; dist[i] = best; }
}

// ---------------------------------------------------------------------------------
// contact concurrences (concurrence.py:100-162): for groups of atom pairs, is the minimum
// compute_distances() distance of the group below the cutoff in each frame?  One CTA per frame:
// the atoms the pairs touch are gathered into shared memory once, every thread walks whole
// groups.  out[f][g] (frame-major, coalesced); the host transposes.
// ---------------------------------------------------------------------------------
constexpr int CC_THREADS = 256;

__global__ void __launch_bounds__(CC_THREADS) k_group_contacts(
    const float* __restrict__ xyz, const float* __restrict__ box, int n_atoms,
    const int* __restrict__ atoms, int n_sel,          // atoms the pairs touch (gathered per frame)
    const int2* __restrict__ pairs,                    // indices INTO `atoms`
    const long long* __restrict__ group_ptr, int n_groups, int orthogonal, float cutoff_f,
    unsigned char* __restrict__ out)
{
    extern __shared__ float cc_xyz[];                  // [n_sel][3]
    const long long f = blockIdx.x;
    const float* x = xyz + (size_t)f * (size_t)n_atoms * 3;
    const DistBox D = make_distbox(box ? box + (size_t)f * 9 : nullptr, orthogonal);
    for (int t = threadIdx.x; t < n_sel; t += CC_THREADS) {
        const int a = __ldg(&atoms[t]);
        cc_xyz[3 * t] = __ldg(&x[(size_t)a * 3]);
        cc_xyz[3 * t + 1] = __ldg(&x[(size_t)a * 3 + 1]);
        cc_xyz[3 * t + 2] = __ldg(&x[(size_t)a * 3 + 2]);
    }
    __syncthreads();
    for (int g = threadIdx.x; g < n_groups; g += CC_THREADS) {
        float best = __int_as_float(0x7f800000);       // min() of an empty group: no contact
        for (long long k = __ldg(&group_ptr[g]); k < __ldg(&group_ptr[g + 1]); k++) {
            const int2 pr = __ldg(&pairs[k]);
            // compute_distances(pair = (i, j)): r = p_j - p_i, float32 sqrt
            const float d = __fsqrt_rn(dist2_pair(cc_xyz[3 * pr.x], cc_xyz[3 * pr.x + 1], cc_xyz[3 * pr.x + 2],
                                                  cc_xyz[3 * pr.y], cc_xyz[3 * pr.y + 1], cc_xyz[3 * pr.y + 2], D));
            best = fminf(best, d);
        }
        out[(size_t)f * (size_t)n_groups + g] = best < cutoff_f ? 1 : 0;
    }
}

}  // namespace cmb
