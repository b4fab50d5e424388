// Issue rate of scalar / packed FP32 and packed FP16 arithmetic on one SM (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_fp16.h>
typedef unsigned long long u64;
#define ITER 4096
template <int OP>
__global__ void k(float* out, long long* cyc, float seed)
{
    float a[8], b = seed, c = seed * 0.5f;
    u64 A[8], B, C;
    __half2 h[8], hb = __float2half2_rn(seed), hc = __float2half2_rn(seed * 0.5f);
    for (int i = 0; i < 8; i++) {
        a[i] = seed + i + threadIdx.x;
        asm("mov.b64 %0, {%1,%2};" : "=l"(A[i]) : "f"(a[i]), "f"(a[i] + 1.f));
        h[i] = __float2half2_rn(a[i] * 1e-3f);
    }
    asm("mov.b64 %0, {%1,%2};" : "=l"(B) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1,%2};" : "=l"(C) : "f"(c), "f"(c));
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __fmaf_rn(a[i], b, c);
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(B), "l"(C));
            if (OP == 2) h[i] = __hfma2(h[i], hb, hc);
            if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B));
            if (OP == 4) h[i] = __hadd2(h[i], hb);
            if (OP == 5) a[i] = __fadd_rn(a[i], b);
            if (OP == 6) h[i] = __hmax2(h[i], hb);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < 8; i++) {
        float lo, hi;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(A[i]));
        s += a[i] + lo + hi + __low2float(h[i]) + __high2float(h[i]);
    }
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int threads)
{
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<1, threads>>>(out, cyc, 1.0001f);
    k<OP><<<1, threads>>>(out, cyc, 1.0001f);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr = (double)ITER * 8 * (threads / 32);
    printf("%-10s threads %4d: %.3f cycles per warp instruction per SMSP\n", name, threads, (double)h / (warp_instr / 4));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int threads : {512, 1024}) {
        run<0>("FFMA", threads); run<1>("FFMA2", threads); run<2>("HFMA2", threads);
        run<5>("FADD", threads); run<3>("FADD2", threads); run<4>("HADD2", threads); run<6>("HMNMX2", threads);
    }
    return 0;
}
