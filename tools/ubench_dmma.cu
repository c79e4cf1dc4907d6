// Does DMMA (mma.sync m8n8k4 f64) overlap with packed fp32 work on B200? products/clk/SM for
//   0: residual trio only (FMUL2+FFMA2+FADD2), 1: DMMA only, 2: both, 3: trio + DFMA (for comparison)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(128, 3) k(float* out, int iters, float seed) {
    float2 es[16]; double acc[8][2]; double dacc[32];
    float f[8];
#pragma unroll
    for (int i = 0; i < 16; i++) es[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i][0] = acc[i][1] = 0.0; f[i] = seed + 0.001f * threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 32; i++) dacc[i] = 0.0;
    double ad = seed, bd = seed * 0.5;
    for (int it = 0; it < iters; it++) {
        const float sc = __int_as_float(0x3f800000 + it);
        if (MODE == 0 || MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float2 x = make_float2(f[i & 7], f[(i + 1) & 7]);
                const float2 b2 = make_float2(sc + (i >> 3), sc + (i >> 3));
                const float2 p = __fmul2_rn(x, b2);
                const float2 e = __ffma2_rn(x, b2, make_float2(-p.x, -p.y));
                es[i] = __fadd2_rn(es[i], e);
            }
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma(acc[i][0], acc[i][1], ad + i, bd + it);   // 4 DMMA = 32 FMAs per thread
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 32; i++) dacc[i] = fma(ad + (i & 3), bd, dacc[i]);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += es[i].x + es[i].y;
#pragma unroll
    for (int i = 0; i < 8; i++) s += (float)(acc[i][0] + acc[i][1]);
#pragma unroll
    for (int i = 0; i < 32; i++) s += (float)dacc[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
    const int blocks = 148 * 3, iters = 20000;
    float* out; cudaMalloc(&out, blocks * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 128>>>(out, 10, 1.0f); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(out, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double prod = (double)blocks * 128 * iters * 32;
    printf("%-40s %7.3f ms  %6.2f products/clk/SM  err=%s\n", name, best, prod / (best * 1e-3) / 148 / (clk_khz * 1e3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() { run<0>("trio only"); run<1>("DMMA only (32 FMA/thread/iter)"); run<2>("trio + DMMA"); run<3>("trio + DFMA"); return 0; }
