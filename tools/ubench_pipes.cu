// Pipe-rate microbenchmark for the exact-arithmetic SGM / cost-volume kernels (sm_100a).
// The reference carries its SGM state and its dot-product accumulator in fp64
// (SURVEY.md App. A1/A2), so the design needs the B200's measured rates for DADD,
// fp64 min, f32<->f64 conversion, DFMA, REDUX and SHFL. Prints ops/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NCH 8

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, float* outf, double seed, int iters) {
    double a[NCH];
    float f[NCH];
    int q[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { a[i] = seed + i * 0.37 + threadIdx.x * 1e-3; f[i] = (float)a[i]; q[i] = threadIdx.x * 7 + i; }
    double p = seed * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (OP == 0) a[i] = a[i] + p;                         // DADD
            if (OP == 1) a[i] = fmin(a[i], a[(i + 1) % NCH] + 0.0 * it) ; // approx DMIN mix (see OP 7)
            if (OP == 2) a[i] = fma(a[i], p, 0.25);               // DFMA
            if (OP == 3) a[i] += (double)f[i], f[i] += 1.0f;      // F2F.F64.F32 + DADD + FADD
            if (OP == 4) f[i] = (float)a[i] + f[i], a[i] += p;    // F2F.F32.F64 + FADD + DADD
            if (OP == 5) q[i] = __reduce_min_sync(0xffffffffu, q[i] + it);   // REDUX
            if (OP == 6) q[i] = __shfl_xor_sync(0xffffffffu, q[i], 1) + it;  // SHFL
            if (OP == 7) a[i] = (a[i] < a[(i + 3) % NCH]) ? a[i] : a[(i + 3) % NCH] ; // DSETP+SEL
            if (OP == 8) f[i] = fminf(f[i], f[(i + 1) % NCH]) + 1.0f;        // FMNMX + FADD
            if (OP == 9) f[i] = fmaf(f[i], 1.0001f, 0.5f);        // FFMA
            if (OP == 10 && (i & 1) == 0) {                       // FFMA2 (one packed op = 2 flops-pairs)
                float2 r = __ffma2_rn(make_float2(f[i], f[i + 1]), make_float2(1.0001f, 1.0002f), make_float2(0.5f, 0.25f));
                f[i] = r.x; f[i + 1] = r.y;
            }
            if (OP >= 11 && (i & 1) == 0) {                       // product-residual trio: FMUL2, FFMA2, FADD2 (+2 DFMA, + F2F)
                const float2 x = make_float2(f[(i + 2) % NCH], f[(i + 3) % NCH]);
                const float sc = __int_as_float(0x3f800000 + it);
                const float2 pr = __fmul2_rn(x, make_float2(sc, sc));
                const float2 e = __ffma2_rn(x, make_float2(sc, sc), make_float2(-pr.x, -pr.y));
                const float2 r = __fadd2_rn(make_float2(f[i], f[i + 1]), e);
                f[i] = r.x; f[i + 1] = r.y;
                if (OP >= 12) { a[i] = fma(a[i], p, 0.25); a[i + 1] = fma(a[i + 1], p, 0.25); }
                if (OP == 13 && (i & 2) == 0) { a[i] += (double)sc; }   // 1 widening per 4 products (kernel: 0.375)
            }
        }
    }
    double s = 0; float sf = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) { s += a[i]; sf += f[i] + q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = sf;
}

template <int OP>
void run(const char* name, double ops_per_inner) {
    int sms = 148, blocks = sms * 8, threads = 256;
    double* out; float* outf;
    cudaMalloc(&out, blocks * threads * sizeof(double));
    cudaMalloc(&outf, blocks * threads * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, outf, 1.0, 64);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, threads>>>(out, outf, 1.0, ITERS);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double total = (double)blocks * threads * ITERS * NCH * ops_per_inner;
    double per_s = total / (best * 1e-3);
    printf("%-28s %8.3f ms  %9.2f Gop/s  %7.2f op/clk/SM @%d MHz(max)  err=%s\n", name, best, per_s * 1e-9,
           per_s / sms / (clk_khz * 1e3), clk_khz / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(outf);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs=%d\n", p.name, p.major, p.minor, p.multiProcessorCount);
    run<0>("DADD", 1);
    run<2>("DFMA", 1);
    run<7>("DSETP+SEL (min by compare)", 1);
    run<1>("fmin(double)+DADD", 1);
    run<3>("cvt f32->f64 (+DADD+FADD)", 1);
    run<4>("cvt f64->f32 (+FADD+DADD)", 1);
    run<5>("REDUX.min s32 (+IADD)", 1);
    run<6>("SHFL.BFLY (+IADD)", 1);
    run<8>("FMNMX+FADD", 1);
    run<9>("FFMA", 1);
    run<10>("FFMA2 (flops pairs; x2 lanes)", 1);
    run<11>("FMUL2+FFMA2+FADD2 products", 1);
    run<12>(" + 1 DFMA per product", 1);
    run<13>(" + 0.25 F2F(+DADD) per product", 1);
    return 0;
}
