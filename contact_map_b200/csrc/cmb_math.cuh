// cmb_math.cuh -- bit-exact float32 distance arithmetic for sm_100a.
//
// Device restatement of the arithmetic the reference obtains from MDTraj at
//   md.compute_neighborlist   /root/reference/contact_map/contact_map.py:575, min_dist.py:54
//   md.compute_distances      /root/reference/contact_map/min_dist.py:63,165
// Every operation is a separately rounded IEEE-754 binary32 op (MDTraj evaluates
// them with SSE intrinsics, which never contract): __fmul_rn / __fadd_rn / __fsub_rn
// are used so nvcc cannot fuse them into FFMA regardless of -fmad.
//
// rint() and floor() compile to FRND (cvt.rni / cvt.rmi f32 -> f32) on the XU pipe: one issue
// slot each, exact for every input.  The kernels are issue-bound on the FMA/ALU pipes while the
// XU pipe is mostly idle, so this is cheaper than the two-FADD 1.5*2^23 trick (kept below for
// reference; it is bit-identical for |x| < 2^22).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmb {

enum BoxMode : int { BOX_NONE = 0, BOX_ORTHO = 1, BOX_TRICLINIC = 2 };

// Per-frame box in the form the neighbour-list arithmetic consumes.
struct NlBox {
    float ax, ay, az;   // row 0 of unitcell_vectors[frame]
    float bx, by, bz;   // row 1
    float cx, cy, cz;   // row 2
    float ix, iy, iz;   // 1.0f / (ax, by, cz)
    int mode;
};

__device__ __forceinline__ NlBox make_nlbox(const float* __restrict__ box9)
{
    NlBox B;
    if (box9 == nullptr) {
        B.ax = B.ay = B.az = B.bx = B.by = B.bz = B.cx = B.cy = B.cz = 0.f;
        B.ix = B.iy = B.iz = 0.f;
        B.mode = BOX_NONE;
        return B;
    }
    B.ax = box9[0]; B.ay = box9[1]; B.az = box9[2];
    B.bx = box9[3]; B.by = box9[4]; B.bz = box9[5];
    B.cx = box9[6]; B.cy = box9[7]; B.cz = box9[8];
    B.ix = __fdiv_rn(1.0f, B.ax);
    B.iy = __fdiv_rn(1.0f, B.by);
    B.iz = __fdiv_rn(1.0f, B.cz);
    const bool tri = (B.ay != 0.f) || (B.az != 0.f) || (B.bx != 0.f) ||
                     (B.bz != 0.f) || (B.cx != 0.f) || (B.cy != 0.f);
    B.mode = tri ? BOX_TRICLINIC : BOX_ORTHO;
    return B;
}

#define CMB_MAGIC 12582912.0f   /* 1.5 * 2^23 */

#ifdef CMB_USE_MAGIC_ROUND
// rintf(x), round-half-even, exact for |x| < 2^22
__device__ __forceinline__ float rint_magic(float x)
{
    return __fadd_rn(__fadd_rn(x, CMB_MAGIC), -CMB_MAGIC);
}

// floorf(x), exact for |x| < 2^22
__device__ __forceinline__ float floor_magic(float x)
{
    float r = rint_magic(x);
    return (r > x) ? __fadd_rn(r, -1.0f) : r;
}
#else
__device__ __forceinline__ float rint_magic(float x) { return rintf(x); }     // FRND.F32
__device__ __forceinline__ float floor_magic(float x) { return floorf(x); }   // FRND.F32.FLOOR
#endif

// (x*x + y*y) + z*z with every op rounded (dpps ordering)
__device__ __forceinline__ float dot3_exact(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// Squared neighbour-list distance of the ordered pair i -> j.
template <int MODE>
__device__ __forceinline__ float nl_d2(float xi, float yi, float zi,
                                       float xj, float yj, float zj, const NlBox& B)
{
    float dx = __fsub_rn(xj, xi);
    float dy = __fsub_rn(yj, yi);
    float dz = __fsub_rn(zj, zi);
    if (MODE == BOX_ORTHO) {
        const float bx = __fmul_rn(rint_magic(__fmul_rn(dx, B.ix)), B.ax);
        const float by = __fmul_rn(rint_magic(__fmul_rn(dy, B.iy)), B.by);
        const float bz = __fmul_rn(rint_magic(__fmul_rn(dz, B.iz)), B.cz);
        dx = __fsub_rn(dx, bx);
        dy = __fsub_rn(dy, by);
        dz = __fsub_rn(dz, bz);
    } else if (MODE == BOX_TRICLINIC) {
        float s = floor_magic(__fadd_rn(__fmul_rn(dz, B.iz), 0.5f));
        dx = __fsub_rn(dx, __fmul_rn(B.cx, s));
        dy = __fsub_rn(dy, __fmul_rn(B.cy, s));
        dz = __fsub_rn(dz, __fmul_rn(B.cz, s));
        s = floor_magic(__fadd_rn(__fmul_rn(dy, B.iy), 0.5f));
        dx = __fsub_rn(dx, __fmul_rn(B.bx, s));
        dy = __fsub_rn(dy, __fmul_rn(B.by, s));
        dz = __fsub_rn(dz, __fmul_rn(B.bz, s));
        s = floor_magic(__fadd_rn(__fmul_rn(dx, B.ix), 0.5f));
        dx = __fsub_rn(dx, __fmul_rn(B.ax, s));
        dy = __fsub_rn(dy, __fmul_rn(B.ay, s));
        dz = __fsub_rn(dz, __fmul_rn(B.az, s));
    }
    return dot3_exact(dx, dy, dz);
}

__device__ __forceinline__ float nl_d2_dyn(float xi, float yi, float zi,
                                           float xj, float yj, float zj, const NlBox& B)
{
    if (B.mode == BOX_ORTHO) return nl_d2<BOX_ORTHO>(xi, yi, zi, xj, yj, zj, B);
    if (B.mode == BOX_TRICLINIC) return nl_d2<BOX_TRICLINIC>(xi, yi, zi, xj, yj, zj, B);
    return nl_d2<BOX_NONE>(xi, yi, zi, xj, yj, zj, B);
}

// ---------------------------------------------------------------------------
// compute_distances arithmetic (dist / dist_mic / dist_mic_triclinic)
// ---------------------------------------------------------------------------
struct DistBox {
    float v1x, v1y, v1z, v2x, v2y, v2z, v3x, v3y, v3z;   // reduced box vectors
    float r1, r2, r3;                                     // 1/v1x, 1/v2y, 1/v3z
    int mode;   // 0 none, 1 orthogonal (dist_mic), 2 triclinic (27-image search)
};

__device__ __forceinline__ DistBox make_distbox(const float* __restrict__ box9, int orthogonal)
{
    DistBox D;
    if (box9 == nullptr) {
        D.v1x = D.v1y = D.v1z = D.v2x = D.v2y = D.v2z = D.v3x = D.v3y = D.v3z = 0.f;
        D.r1 = D.r2 = D.r3 = 0.f;
        D.mode = 0;
        return D;
    }
    D.v1x = box9[0]; D.v1y = box9[1]; D.v1z = box9[2];
    D.v2x = box9[3]; D.v2y = box9[4]; D.v2z = box9[5];
    D.v3x = box9[6]; D.v3y = box9[7]; D.v3z = box9[8];
    float s = roundf(__fdiv_rn(D.v3y, D.v2y));
    D.v3x = __fsub_rn(D.v3x, __fmul_rn(D.v2x, s));
    D.v3y = __fsub_rn(D.v3y, __fmul_rn(D.v2y, s));
    D.v3z = __fsub_rn(D.v3z, __fmul_rn(D.v2z, s));
    s = roundf(__fdiv_rn(D.v3x, D.v1x));
    D.v3x = __fsub_rn(D.v3x, __fmul_rn(D.v1x, s));
    D.v3y = __fsub_rn(D.v3y, __fmul_rn(D.v1y, s));
    D.v3z = __fsub_rn(D.v3z, __fmul_rn(D.v1z, s));
    s = roundf(__fdiv_rn(D.v2x, D.v1x));
    D.v2x = __fsub_rn(D.v2x, __fmul_rn(D.v1x, s));
    D.v2y = __fsub_rn(D.v2y, __fmul_rn(D.v1y, s));
    D.v2z = __fsub_rn(D.v2z, __fmul_rn(D.v1z, s));
    D.r1 = __fdiv_rn(1.0f, D.v1x);
    D.r2 = __fdiv_rn(1.0f, D.v2y);
    D.r3 = __fdiv_rn(1.0f, D.v3z);
    D.mode = orthogonal ? 1 : 2;
    return D;
}

// squared distance p1 -> p2 before the final sqrtf
__device__ __forceinline__ float dist2_pair(float x1, float y1, float z1,
                                            float x2, float y2, float z2, const DistBox& D)
{
    float rx = __fsub_rn(x2, x1), ry = __fsub_rn(y2, y1), rz = __fsub_rn(z2, z1);
    if (D.mode == 0) return dot3_exact(rx, ry, rz);
    float s = roundf(__fmul_rn(rz, D.r3));
    rx = __fsub_rn(rx, __fmul_rn(D.v3x, s));
    ry = __fsub_rn(ry, __fmul_rn(D.v3y, s));
    rz = __fsub_rn(rz, __fmul_rn(D.v3z, s));
    s = roundf(__fmul_rn(ry, D.r2));
    rx = __fsub_rn(rx, __fmul_rn(D.v2x, s));
    ry = __fsub_rn(ry, __fmul_rn(D.v2y, s));
    rz = __fsub_rn(rz, __fmul_rn(D.v2z, s));
    s = roundf(__fmul_rn(rx, D.r1));
    rx = __fsub_rn(rx, __fmul_rn(D.v1x, s));
    ry = __fsub_rn(ry, __fmul_rn(D.v1y, s));
    rz = __fsub_rn(rz, __fmul_rn(D.v1z, s));
    float best = dot3_exact(rx, ry, rz);
    if (D.mode == 1) return best;
#pragma unroll
    for (int x = -1; x < 2; x++) {
        const float fx = (float)x;
        const float ax = __fadd_rn(rx, __fmul_rn(D.v1x, fx));
        const float ay = __fadd_rn(ry, __fmul_rn(D.v1y, fx));
        const float az = __fadd_rn(rz, __fmul_rn(D.v1z, fx));
#pragma unroll
        for (int y = -1; y < 2; y++) {
            const float fy = (float)y;
            const float bx = __fadd_rn(ax, __fmul_rn(D.v2x, fy));
            const float by = __fadd_rn(ay, __fmul_rn(D.v2y, fy));
            const float bz = __fadd_rn(az, __fmul_rn(D.v2z, fy));
#pragma unroll
            for (int z = -1; z < 2; z++) {
                const float fz = (float)z;
                const float qx = __fadd_rn(bx, __fmul_rn(D.v3x, fz));
                const float qy = __fadd_rn(by, __fmul_rn(D.v3y, fz));
                const float qz = __fadd_rn(bz, __fmul_rn(D.v3z, fz));
                const float d2 = dot3_exact(qx, qy, qz);
                if (d2 <= best) best = d2;
            }
        }
    }
    return best;
}

}  // namespace cmb
