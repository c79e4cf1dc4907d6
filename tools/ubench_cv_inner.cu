// Inner-loop variants of the exact cost-volume dot product (sm_100a): products per clock per SM for one 8x4
// register tile per thread, operands read from shared memory exactly as cost_volume.cu does.
//   S1  FMUL + F2F.F64.F32 + DADD           (literal contract)
//   S2  FMUL + 3 ALU ops (bit widening) + DADD
//   S3  FMUL/FFMA/FADD residual chain (packed or scalar) + DFMA on pre-widened operands
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_cv_inner ubench_cv_inner.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NF = 64, T = 64;

__device__ __forceinline__ double widen3(float p) {  // |p| < 2, normal, non-zero
    const uint32_t x = __float_as_uint(p);
    const int32_t t = (int32_t)x >> 3;
    const uint32_t hi = ((uint32_t)t & 0x8fffffffu) | 0x38000000u;
    return __hiloint2double((int)hi, (int)__funnelshift_l(0u, x, 29));
}

__device__ __forceinline__ double widen3m(float p) {  // left shift on the FMA pipe (IMAD), one SHF left
    const uint32_t x = __float_as_uint(p);
    const int32_t t = (int32_t)x >> 3;
    const uint32_t hi = ((uint32_t)t & 0x8fffffffu) | 0x38000000u;
    uint32_t lo;
    asm("mad.lo.u32 %0, %1, 536870912, 0;" : "=r"(lo) : "r"(x));
    return __hiloint2double((int)hi, (int)lo);
}

// MODE: 0 S3 packed, 1 S3 scalar, 2 S1, 3 S1/S2 checkerboard, 4 S2, 5 S3 packed rows 0..5 + S1 rows 6,7, 6: S1 on (i+j)%3==0 else S2
template <int MODE>
__global__ void __launch_bounds__(128, 3) inner(float* out, int reps) {
    __shared__ __align__(16) float a[NF][T];
    __shared__ __align__(16) float b[NF][T];
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    for (int i = tid; i < NF * T; i += 128) { (&a[0][0])[i] = 0.01f + 1e-4f * (i % 97); (&b[0][0])[i] = 0.02f - 1e-4f * (i % 89); }
    __syncthreads();
    double acc[8][4]; float2 es[4][4]; float es1[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j] = 0.0; es1[i][j] = 0.f; }
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int j = 0; j < 4; j++) es[q][j] = make_float2(0.f, 0.f);
    for (int r = 0; r < reps; r++) {
#pragma unroll 2
        for (int k = 0; k < NF; k++) {
            const int key = (k >> 2) & 15;
            const float4 a0 = *reinterpret_cast<const float4*>(&a[k][((tx ^ key) & 15) << 2]);
            const float4 a1 = *reinterpret_cast<const float4*>(&a[k][(((tx + 8) ^ key) & 15) << 2]);
            const float4 bv = *reinterpret_cast<const float4*>(&b[k][((ty ^ key) & 15) << 2]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
            if (MODE == 0 || MODE == 5) {
                constexpr int NI = MODE == 0 ? 8 : 6;
                double ad[NI];
#pragma unroll
                for (int i = 0; i < NI; i++) ad[i] = (double)av[i];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 b2 = make_float2(bb[j], bb[j]);
                    const double bd = (double)bb[j];
#pragma unroll
                    for (int q = 0; q < NI / 2; q++) {
                        const float2 a2 = make_float2(av[2 * q], av[2 * q + 1]);
                        const float2 p = __fmul2_rn(a2, b2);
                        const float2 e = __ffma2_rn(a2, b2, make_float2(-p.x, -p.y));
                        es[q][j] = __fadd2_rn(es[q][j], e);
                    }
#pragma unroll
                    for (int i = 0; i < NI; i++) acc[i][j] = fma(ad[i], bd, acc[i][j]);
#pragma unroll
                    for (int i = NI; i < 8; i++) acc[i][j] += (double)__fmul_rn(av[i], bb[j]);
                }
            } else if (MODE == 1) {
                double ad[8];
#pragma unroll
                for (int i = 0; i < 8; i++) ad[i] = (double)av[i];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const double bd = (double)bb[j];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float p = __fmul_rn(av[i], bb[j]);
                        es1[i][j] += __fmaf_rn(av[i], bb[j], -p);
                        acc[i][j] = fma(ad[i], bd, acc[i][j]);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float p = __fmul_rn(av[i], bb[j]);
                        bool f2f = MODE == 2 || (MODE == 3 && ((i + j) & 1)) || (MODE == 6 && ((i + j) % 3 == 0)) || (MODE == 7 && ((i + 2 * j) % 5 < 2)) ||
                                   (MODE == 8 && ((i + j) & 1)) || (MODE == 9 && ((i + 3 * j) % 8 < 3)) || (MODE == 10 && ((i + j) % 3 == 0)) ||
                                   (MODE == 11 && ((i + j) % 4 == 0));
                        acc[i][j] += f2f ? (double)p : (MODE >= 8 ? widen3m(p) : widen3(p));
                    }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += (float)acc[i][j] + es1[i][j] + ((i & 1) ? es[i >> 1][j].y : es[i >> 1][j].x);
    out[blockIdx.x * 128 + tid] = s;
}

template <int MODE>
void run(const char* name) {
    const int blocks = 148 * 3, reps = 200;
    float* out; cudaMalloc(&out, blocks * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    inner<MODE><<<blocks, 128>>>(out, 4); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); inner<MODE><<<blocks, 128>>>(out, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double prod = (double)blocks * 128 * reps * NF * 32;
    printf("%-44s %7.3f ms  %6.2f products/clk/SM  err=%s\n", name, best, prod / (best * 1e-3) / 148 / (clk_khz * 1e3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    run<0>("S3 packed (FMUL2/FFMA2/FADD2 + DFMA)");
    run<1>("S3 scalar (FMUL/FFMA/FADD + DFMA)");
    run<2>("S1 all F2F");
    run<3>("S1/S2 checkerboard 1:1");
    run<6>("S1/S2 1:2");
    run<7>("S1/S2 2:3");
    run<4>("S2 all 3-op bit widening");
    run<5>("S3 packed rows 0-5 + S1 rows 6-7");
    run<8>("S1/S2m (IMAD left shift) 1:1");
    run<9>("S1/S2m 3:5");
    run<10>("S1/S2m 1:2");
    run<11>("S1/S2m 1:3");
    run<12>("S2m all");
    return 0;
}
