// Tensor-core variant of the exact cost volume (sm_100a): tcgen05.mma on 8-bit slices of the features + FP32 residuals.
//
// Same contract as cost_volume.cu (compute_cost_volume_kernel, process_functional.py:120-131; fill :1111-1114), same
// bits out. north_star asks for "a tensor-core row-banded variant kept only if ncu shows it wins".
//
// The reference value of one evaluation is RN32(-temp), temp = fp64 sequential sum of p_k = RN32(f_k g_k), k = 0..63.
//   sum_k p_k = sum_k f_k g_k - sum_k e_k,     e_k = fma(f_k, g_k, -p_k) = f_k g_k - p_k   (exact in fp32)
// * sum_k f_k g_k comes from the tensor cores, EXACTLY up to a known truncation: every feature vector is scaled by a
//   power of two 2^-E so that |m| <= 1/2 and cut into six signed 8-bit slices (m = sum_s q_s 2^-8(s+1), |q_s| <= 128,
//   stored as fp16 integers). A slice-pair product summed over the 64 features is an integer below 2^20, so fp32
//   accumulation of up to 6 pairs of equal weight is exact; the 21 pairs with s + t <= 5 go to six TMEM accumulators
//   (one per weight 2^-8(s+t+2)), 84 MMAs (M=128, N=64, K=16) per 128x64 tile. The epilogue adds the six integers in
//   fp64. Truncation (features below 2^-48 of the scale, the dropped pairs) is bounded by 2^(Ea+Eb) * 2^-41.
// * sum_k e_k runs on the CUDA cores as FMUL2 / FFMA2 / FADD2 (3 FP32 lane-slots per product instead of the
//   FMUL + widening + DADD of the literal loop), overlapped with the asynchronous MMAs of the same tile.
// * T = (tensor-core sum) - (residual sum) differs from the reference's temp by less than eps (derivation in DESIGN.md
//   4.2); if RN32(T - eps) == RN32(T + eps) that value is the reference's; otherwise (about 3 in 10^4, and for every
//   pixel whose features are not finite / not comfortably scaled) the evaluation is queued and redone with the literal
//   loop by a small fix-up kernel.
//
// Thread map: 16 compute warps (8x2 register tiles for the residuals; in the epilogue warp w owns TMEM lanes 32*(w%4)..,
// i.e. tile rows x, and the 16 tile columns u of quarter w/4) + 1 control warp (TMA, MMA issue).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;
constexpr int NS = 6;                     // slices per feature
constexpr int NACC = 6;                   // accumulators: weight classes s + t = 0..5
constexpr int TM = 128, TN = 64;          // tile: x pixels, u pixels
constexpr int A_SLICE = TM * 128, B_SLICE = TN * 128;  // bytes of one fp16 slice tile (128-byte rows)
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + NS * A_SLICE;            // 98304
constexpr int OFF_A32 = OFF_B + NS * B_SLICE;          // 147456: two k-halves of [128 px][32 floats]
constexpr int OFF_B32 = OFF_A32 + 2 * TM * 128;        // 180224: two k-halves of [64 px][32 floats]
constexpr int RES_PITCH = 66;
constexpr int OFF_RES = OFF_B32 + 2 * TN * 128;        // 196608
constexpr int OFF_SB = OFF_RES + TM * RES_PITCH * 4;   // 230400: scale, norm of the 64 u pixels
constexpr int OFF_BAR = OFF_SB + 2 * TN * 8;           // 231424 (scale, norm * 2^-41 as doubles)
constexpr int TC_SMEM = OFF_BAR + 64 + 960;            // barriers + alignment slack = 232448, the sm_100 maximum
constexpr int NCW = 16;            // compute warps
constexpr int CVT_THREADS = 32 * (NCW + 1);  // + 1 control warp
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);  // f16 x f16 -> f32
constexpr int PADPIX = 128;               // per-pixel arrays carry this many entries of padding on both ends
constexpr float kInfF = __builtin_huge_valf();

// ---------------------------------------------------------------- PTX wrappers (same forms as conv_tc.cu)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_addr(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- pre-pass: slices, scales, norms
// 8 threads per pixel (8 features each). slices: [NS][P][64] fp16; scale[pix] = 2^E (NaN when the pixel must take the
// literal loop); norm[pix] >= |f|_2.
__global__ void __launch_bounds__(256) cv_slice_kernel(const float* __restrict__ feat, __half* __restrict__ slices,
                                                      float* __restrict__ scale, float* __restrict__ norm, long long P) {
    const long long pix = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
    const int part = threadIdx.x & 7;
    float f[8];
    if (pix < P) {
        const float4 v0 = *reinterpret_cast<const float4*>(feat + pix * NF + 8 * part);
        const float4 v1 = *reinterpret_cast<const float4*>(feat + pix * NF + 8 * part + 4);
        f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = 0.f;
    }
    float amax = 0.f, ss = 0.f;
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        amax = fmaxf(amax, fabsf(f[i]));
        ss = fmaf(f[i], f[i], ss);
        fin = fin && (fabsf(f[i]) < kInfF);  // false for Inf and NaN
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        fin = __shfl_xor_sync(0xffffffffu, (int)fin, o) && fin;
    }
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu) - 127;
    const bool ok = fin && e >= -60 && e <= 60;  // also rejects the all-zero pixel (e = -127)
    const int E = e + 2;                         // |f| * 2^-E <= 1/2
    const float inv = __uint_as_float((uint32_t)(127 - E) << 23);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = ok ? f[i] * inv : 0.f;
    if (pix < P) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            __align__(16) __half q[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float r = m[i] * 256.0f;
                const float qi = (r + 12582912.0f) - 12582912.0f;  // rint, |r| <= 128
                m[i] = r - qi;                                      // exact: the fraction bits of r
                q[i] = __float2half_rn(qi);
            }
            *reinterpret_cast<uint4*>(slices + ((long long)s * P + pix) * NF + 8 * part) = *reinterpret_cast<const uint4*>(q);
        }
        if (part == 0) {
            scale[pix] = ok ? __uint_as_float((uint32_t)(127 + E) << 23) : __int_as_float(0x7fc00000);
            norm[pix] = ok ? sqrtf(ss) * 1.001f : kInfF;
        }
    }
}

// entries no evaluation writes: fill where the match falls outside the other image, +INF pads
__global__ void __launch_bounds__(256) cv_fill_kernel(float* __restrict__ CL, float* __restrict__ CR, int W, int D, int Dp,
                                                     long long P, float fill) {
    const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pix >= P) return;
    const int x = (int)(pix % W);
    float* rl = CL + pix * Dp;
    for (int d = x + 1 + lane; d < D; d += 32) rl[d] = fill;        // x - d < 0
    if (Dp > D && lane < Dp - D) rl[D + lane] = kInfF;
    if (CR != nullptr) {
        float* rr = CR + pix * Dp;
        for (int d = max(W - x, 0) + lane; d < D; d += 32) rr[d] = fill;  // x + d >= W
        if (Dp > D && lane < Dp - D) rr[D + lane] = kInfF;
    }
}

struct CvTcArgs {
    const float* fl;
    const float* fr;
    const float* scaleL;  // padded arrays: index PADPIX + pixel
    const float* normL;
    const float* scaleR;
    const float* normR;
    float* CL;
    float* CR;
    int H, W, D, Dp;
    long long P;
    int tiles_x, nitems;
    // evaluations whose rounding could not be proven: (pixel index of x) << 12 | d, redone by cv_fixup_kernel
    unsigned long long* queue;
    unsigned* queue_count;
    unsigned queue_cap;
};

// the literal reference loop for one evaluation, operands from global memory (only when the queue is full)
__device__ __noinline__ float exact_dot_global(const float* __restrict__ a, const float* __restrict__ b) {
    double t = 0.0;
    for (int k = 0; k < NF; k++) t += (double)__fmul_rn(a[k], b[k]);
    return (float)(-t);
}

// the literal reference loop for the queued evaluations: one thread each, both volumes patched
__global__ void __launch_bounds__(128) cv_fixup_kernel(const float* __restrict__ fl, const float* __restrict__ fr,
                                                      float* __restrict__ CL, float* __restrict__ CR,
                                                      const unsigned long long* __restrict__ queue,
                                                      const unsigned* __restrict__ queue_count, unsigned cap, int Dp) {
    const unsigned n = min(*queue_count, cap);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long e = queue[i];
        const long long px = (long long)(e >> 12);
        const int d = (int)(e & 0xfffu);
        const float4* pa = reinterpret_cast<const float4*>(fl + px * NF);
        const float4* pb = reinterpret_cast<const float4*>(fr + (px - d) * NF);
        double t = 0.0;
#pragma unroll 4
        for (int k4 = 0; k4 < NF / 4; k4++) {
            const float4 av = __ldg(pa + k4), bv = __ldg(pb + k4);
            t += (double)__fmul_rn(av.x, bv.x);
            t += (double)__fmul_rn(av.y, bv.y);
            t += (double)__fmul_rn(av.z, bv.z);
            t += (double)__fmul_rn(av.w, bv.w);
        }
        const float v = (float)(-t);
        CL[px * Dp + d] = v;
        if (CR != nullptr) CR[(px - d) * Dp + d] = v;
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NCW) : "memory"); }  // the compute warps

// Warp roles: warps 0..7 compute (residuals, epilogue, stores), warp 8 = control (TMA loads, MMA issue, TMEM allocation).
// Barriers (one phase per tile unless noted): b_full (TMA bytes), mma (tcgen05.commit), b_free (one arrival per compute warp: the compute
// warps have read the fp32 B rows), t_free (one per compute warp: TMEM has been read), a_full (TMA bytes, one phase per item).
__global__ void __launch_bounds__(CVT_THREADS, 1)
cost_volume_tc_kernel(const __grid_constant__ CUtensorMap tmSL, const __grid_constant__ CUtensorMap tmSR,
                      const __grid_constant__ CUtensorMap tmFL, const __grid_constant__ CUtensorMap tmFR, const CvTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bar_a_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_b_full = bar_a_full + 1;
    uint64_t* bar_mma = bar_a_full + 2;
    uint64_t* bar_b_free = bar_a_full + 3;
    uint64_t* bar_t_free = bar_a_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a_full + 5);
    float* res = reinterpret_cast<float*>(sm + OFF_RES);  // [128][RES_PITCH]: residual sums, then the results
    double* sbd = reinterpret_cast<double*>(sm + OFF_SB);  // [0..63] scale, [64..127] norm * 2^-41 of the u pixels

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_b_full, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_b_free, NCW);
        mbar_init(bar_t_free, NCW);
        mbar_fence_init();
    }
    if (warp == NCW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == NCW) {
        // ================================================================= control warp
        if (lane == 0) {
            uint32_t n_item = 0, n_tile = 0;  // items / tiles issued so far (barrier phases)
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, n_item++) {
                const int y = item / a.tiles_x, x0 = (item % a.tiles_x) * TM;
                const long long prow = (long long)y * a.W;
                const int u_lo = max(x0 - (a.D - 1), 0), u_hi = min(x0 + TM - 1, a.W - 1);
                bool a_pending = true;
                for (int ut = u_lo / TN; ut <= u_hi / TN; ut++, n_tile++) {
                    const int u0 = ut * TN;
                    if (n_tile > 0) {  // the previous tile has left the operand buffers
                        mbar_wait(bar_b_free, (n_tile - 1) & 1u);
                        mbar_wait(bar_mma, (n_tile - 1) & 1u);
                    }
                    if (a_pending) {
                        mbar_expect_tx(bar_a_full, NS * A_SLICE + 2 * TM * 128);
                        for (int s = 0; s < NS; s++)
                            tma_load_2d(base + OFF_A + s * A_SLICE, &tmSL, 0, (int)((long long)s * a.P + prow + x0), bar_a_full);
                        tma_load_2d(base + OFF_A32, &tmFL, 0, (int)(prow + x0), bar_a_full);
                        tma_load_2d(base + OFF_A32 + TM * 128, &tmFL, 32, (int)(prow + x0), bar_a_full);
                    }
                    mbar_expect_tx(bar_b_full, NS * B_SLICE + 2 * TN * 128);
                    for (int s = 0; s < NS; s++)
                        tma_load_2d(base + OFF_B + s * B_SLICE, &tmSR, 0, (int)((long long)s * a.P + prow + u0), bar_b_full);
                    tma_load_2d(base + OFF_B32, &tmFR, 0, (int)(prow + u0), bar_b_full);
                    tma_load_2d(base + OFF_B32 + TN * 128, &tmFR, 32, (int)(prow + u0), bar_b_full);
                    if (a_pending) {
                        mbar_wait(bar_a_full, n_item & 1u);
                        a_pending = false;
                    }
                    mbar_wait(bar_b_full, n_tile & 1u);
                    if (n_tile > 0) mbar_wait(bar_t_free, (n_tile - 1) & 1u);  // the accumulators have been read
                    tc_fence_after();
#pragma unroll 1
                    for (int v = 0; v < NACC; v++) {
                        for (int s = 0; s <= v; s++) {
                            const int t = v - s;
                            const uint64_t ad = sw128_desc(base + OFF_A + s * A_SLICE);
                            const uint64_t bd = sw128_desc(base + OFF_B + t * B_SLICE);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                umma_f16(tmem_base + (uint32_t)v * TN, ad + 2 * k, bd + 2 * k, (s | k) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(bar_mma);
                }
            }
        }
    } else {
        // ================================================================= compute warps
        // residual tile of a thread: x = 64*wx + tx + 8i (i < 8), u = 8*wy + tyl + 4j (j < 2)
        const int tx = lane & 7, tyl = lane >> 3, wx = warp & 1, wy = warp >> 1;
        const int ub = 8 * wy + tyl;
        // epilogue row / columns of a thread (TMEM lane = tile row): x = 32*q + lane, u = 16*part + 0..15
        const int q = warp & 3, part = warp >> 2;
        const int xl = 32 * q + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + 16u * part;
        uint32_t n_item = 0, n_tile = 0;

        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, n_item++) {
            const int y = item / a.tiles_x, x0 = (item % a.tiles_x) * TM;
            const long long prow = (long long)y * a.W;
            const int x = x0 + xl;
            const bool x_ok = x < a.W;
            const double sa = (double)a.scaleL[PADPIX + prow + min(x, a.W - 1)];
            const double na = (double)a.normL[PADPIX + prow + min(x, a.W - 1)];
            const int u_lo = max(x0 - (a.D - 1), 0), u_hi = min(x0 + TM - 1, a.W - 1);
            bool first = true;
            for (int ut = u_lo / TN; ut <= u_hi / TN; ut++, n_tile++) {
                const int u0 = ut * TN;
                if (tid < 2 * TN) {
                    const float vv = (tid < TN ? a.scaleR : a.normR)[PADPIX + prow + u0 + (tid & (TN - 1))];
                    sbd[tid] = tid < TN ? (double)vv : (double)vv * 0x1p-41;
                }
                if (first) {
                    mbar_wait(bar_a_full, n_item & 1u);
                    first = false;
                }
                mbar_wait(bar_b_full, n_tile & 1u);
                // ---- fp32 rounding residuals of the 64 products of each of the 8 x 2 evaluations of this thread
                float2 es[8][2];
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++) es[i][j] = make_float2(0.f, 0.f);
                // this warp's block is x in [64 wx, +64) x u in [8 wy, +8): skipped when it lies outside the band 0 <= d < D
                const int wd_max = (x0 + 64 * wx + 63) - (u0 + 8 * wy), wd_min = (x0 + 64 * wx) - (u0 + 8 * wy + 7);
                if (wd_max >= 0 && wd_min < a.D && x0 + 64 * wx < a.W && u0 + 8 * wy < a.W) {
                    const unsigned char* ap = sm + OFF_A32 + (64 * wx + tx) * 128;  // + i * 1024: rows tx + 8i keep (row & 7) = tx
                    const unsigned char* bp = sm + OFF_B32 + ub * 128;              // + j * 512: rows ub + 4j, (row & 7) = tyl + 4j
#pragma unroll 4
                    for (int kc = 0; kc < 16; kc++) {  // 4 features per step: 16-byte chunk kc & 7 of k half kc >> 3
                        const int c = kc & 7;
                        const unsigned char* apk = ap + (kc >> 3) * (TM * 128) + ((c ^ tx) << 4);
                        const unsigned char* bpk = bp + (kc >> 3) * (TN * 128);
                        float4 bv[2];
#pragma unroll
                        for (int j = 0; j < 2; j++) bv[j] = *reinterpret_cast<const float4*>(bpk + j * 512 + ((c ^ (tyl + 4 * j)) << 4));
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            const float4 av = *reinterpret_cast<const float4*>(apk + i * 1024);
                            const float2 a01 = make_float2(av.x, av.y), a23 = make_float2(av.z, av.w);
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                const float2 b01 = make_float2(bv[j].x, bv[j].y), b23 = make_float2(bv[j].z, bv[j].w);
                                const float2 p0 = __fmul2_rn(a01, b01), p1 = __fmul2_rn(a23, b23);
                                const float2 e0 = __ffma2_rn(a01, b01, make_float2(-p0.x, -p0.y));
                                const float2 e1 = __ffma2_rn(a23, b23, make_float2(-p1.x, -p1.y));
                                es[i][j] = __fadd2_rn(es[i][j], __fadd2_rn(e0, e1));
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_b_free);  // the fp32 rows of this tile are no longer read
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        res[(64 * wx + tx + 8 * i) * RES_PITCH + ub + 4 * j] = es[i][j].x + es[i][j].y;
                compute_sync();  // residual sums and sbd are in shared memory

                // ---- epilogue: integers from TMEM -> fp64 sum, minus the residuals, rounding-interval test
                mbar_wait(bar_mma, n_tile & 1u);
                tc_fence_after();
                float* myrow = res + xl * RES_PITCH + 16 * part;
                // evaluations jj of this thread with 0 <= d < D and u < W: jj in [j_first, j_last]; d = dbase - jj
                const int dbase = x - (u0 + 16 * part);
                const int j_first = x_ok ? max(0, dbase - (a.D - 1)) : 16;
                const int j_last = min(15, min(dbase, a.W - 1 - (u0 + 16 * part)));
                float* crp = a.CR ? a.CR + (prow + u0 + 16 * part) * a.Dp + dbase : nullptr;  // CR[y][u][d] of jj = 0; +Dp-1 per jj
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    // nothing of this chunk is inside the band for any lane (tcgen05.ld is warp-collective: uniform test)
                    if (__all_sync(0xffffffffu, j_last < 8 * c || j_first > 8 * c + 7)) continue;
                    float r[NACC][8];
#pragma unroll
                    for (int v = 0; v < NACC; v++) tmem_ld8(taddr + (uint32_t)v * TN + 8u * c, r[v]);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        // the six accumulators hold exact integers below 2^23: pairs are merged in int32 (hi * 256 + lo),
                        // int32 -> fp64 by bit pasting (2^52 + 2^31 + n), no conversion-pipe instruction
                        double dsum = 0.0;
#pragma unroll
                        for (int g = NACC / 2 - 1; g >= 0; g--) {  // smallest weight first
                            const float2 m2 = __fadd2_rn(make_float2(r[2 * g][j], r[2 * g + 1][j]), make_float2(12582912.0f, 12582912.0f));
                            const int n = (__float_as_int(m2.x) - 0x4B400000) * 256 + (__float_as_int(m2.y) - 0x4B400000);
                            const double nd = __hiloint2double(0x43300000, n ^ (int)0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
                            const double wg = __longlong_as_double((long long)(1023 - 8 * (2 * g + 3)) << 52);       // 2^-8(2g+3)
                            dsum = fma(nd, wg, dsum);
                        }
                        const int jj = 8 * c + j, ul = 16 * part + jj;
                        const double sc = sa * sbd[ul];  // 2^(Ea+Eb), NaN for a wild pixel
                        const double T = fma(dsum, sc, -(double)myrow[jj]);
                        const double eps = fma(na, sbd[TN + ul], fma(sc, 0x1p-40, 0x1p-140));
                        const float lo = __double2float_rn(T - eps), hi = __double2float_rn(T + eps);
                        float outv = -lo;
                        const bool valid = jj >= j_first && jj <= j_last;
                        if (valid && !(lo == hi)) {  // cannot be proven: queue it for the literal loop
                            const unsigned slot = atomicAdd(a.queue_count, 1u);
                            if (slot < a.queue_cap)
                                a.queue[slot] = ((unsigned long long)(prow + x) << 12) | (unsigned)(dbase - jj);
                            else
                                outv = exact_dot_global(a.fl + (prow + x) * NF, a.fr + (prow + x - (dbase - jj)) * NF);
                        }
                        myrow[jj] = outv;
                        // CR[y][u][d]: for a fixed u the lanes of a warp hold consecutive x = consecutive d
                        if (valid && crp != nullptr) crp[(long long)jj * (a.Dp - 1)] = outv;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_free);  // TMEM may take the next tile
                compute_sync();                           // results complete in shared memory
                // ---- CL[y][x][d], d = x - u: for a fixed x the tile's u range is a contiguous run of d
                {
                    const bool interior = x0 + TM <= a.W && u0 + TN <= a.W && x0 - (u0 + TN - 1) >= 0 && (x0 + TM - 1) - u0 < a.D;
                    const int ul0 = (TN - 1) - lane, ul1 = (TN - 1) - (lane + 32);
                    for (int r = warp; r < TM; r += NCW) {
                        const int xr = x0 + r;
                        if (xr >= a.W) break;
                        float* row = a.CL + (prow + xr) * a.Dp + (xr - u0);
                        const float v0 = res[r * RES_PITCH + ul0], v1 = res[r * RES_PITCH + ul1];
                        if (interior) {
                            row[-ul0] = v0;
                            row[-ul1] = v1;
                        } else {
                            const int d0 = xr - u0 - ul0, d1 = xr - u0 - ul1;
                            if (d0 >= 0 && d0 < a.D && u0 + ul0 < a.W) row[-ul0] = v0;
                            if (d1 >= 0 && d1 < a.D && u0 + ul1 < a.W) row[-ul1] = v1;
                        }
                    }
                }
                compute_sync();  // the result tile may be overwritten
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NCW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap* tm, CUtensorMapDataType dt, const void* ptr, size_t elem, int inner, size_t rows, int box_inner,
             int box_rows) {
    EncodeTiledFn enc = get_encode();
    MCCNN_REQUIRE(enc != nullptr, MCCNN_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)inner * elem};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCCNN_REQUIRE(r == CUDA_SUCCESS, MCCNN_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_cost_volume_tc_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    const size_t P = (size_t)H * W;
    return 2 * a256((size_t)NS * P * NF * sizeof(__half)) + 4 * a256((P + 2 * PADPIX) * sizeof(float)) +
           a256((8 * P + 65536) * sizeof(unsigned long long)) + 256;
}

extern "C" int mccnn_cost_volume_tc(const float* fl, const float* fr, float* CL, float* CR, void* workspace,
                                    size_t workspace_bytes, int H, int W, int D, float fill, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(fl && fr && CL && workspace, MCCNN_EINVAL, "mccnn_cost_volume_tc: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096, MCCNN_EINVAL, "mccnn_cost_volume_tc: bad shape H=%d W=%d D=%d", H, W, D);
    MCCNN_REQUIRE((long long)NS * H * W + TM < 0x7fffffffLL, MCCNN_EINVAL, "mccnn_cost_volume_tc: image too large for 32-bit tile rows");
    MCCNN_REQUIRE(aligned16(fl) && aligned16(fr) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN,
                  "mccnn_cost_volume_tc: features must be 16-byte, the workspace 256-byte aligned");
    MCCNN_REQUIRE(workspace_bytes >= mccnn_cost_volume_tc_workspace_bytes(H, W), MCCNN_EWORKSPACE,
                  "mccnn_cost_volume_tc: workspace too small");
    const long long P = (long long)H * W;
    const int Dp = disp_pitch(D);
    char* ws = reinterpret_cast<char*>(workspace);
    const size_t slice_bytes = a256((size_t)NS * P * NF * sizeof(__half)), arr_bytes = a256((P + 2 * PADPIX) * sizeof(float));
    __half* sl = reinterpret_cast<__half*>(ws);
    __half* sr = reinterpret_cast<__half*>(ws + slice_bytes);
    float* scaleL = reinterpret_cast<float*>(ws + 2 * slice_bytes);
    float* normL = reinterpret_cast<float*>(ws + 2 * slice_bytes + arr_bytes);
    float* scaleR = reinterpret_cast<float*>(ws + 2 * slice_bytes + 2 * arr_bytes);
    float* normR = reinterpret_cast<float*>(ws + 2 * slice_bytes + 3 * arr_bytes);
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(ws + 2 * slice_bytes + 4 * arr_bytes);
    const size_t queue_cap = 8 * (size_t)P + 65536;
    unsigned* queue_count = reinterpret_cast<unsigned*>(ws + 2 * slice_bytes + 4 * arr_bytes + a256(queue_cap * sizeof(unsigned long long)));
    MCCNN_CUDA(cudaMemsetAsync(scaleL, 0, 4 * arr_bytes, stream));  // the padding entries must be readable numbers
    MCCNN_CUDA(cudaMemsetAsync(queue_count, 0, 256, stream));

    const unsigned nb = (unsigned)((P + 31) / 32);
    cv_slice_kernel<<<nb, 256, 0, stream>>>(fl, sl, scaleL + PADPIX, normL + PADPIX, P);
    MCCNN_LAUNCH_CHECK("cv_slice_kernel");
    cv_slice_kernel<<<nb, 256, 0, stream>>>(fr, sr, scaleR + PADPIX, normR + PADPIX, P);
    MCCNN_LAUNCH_CHECK("cv_slice_kernel");
    cv_fill_kernel<<<(unsigned)((P + 7) / 8), 256, 0, stream>>>(CL, CR, W, D, Dp, P, fill);
    MCCNN_LAUNCH_CHECK("cv_fill_kernel");

    CUtensorMap tmSL, tmSR, tmFL, tmFR;
    if (int e = make_map(&tmSL, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, sl, 2, NF, (size_t)NS * P, NF, TM)) return e;
    if (int e = make_map(&tmSR, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, sr, 2, NF, (size_t)NS * P, NF, TN)) return e;
    if (int e = make_map(&tmFL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fl, 4, NF, (size_t)P, 32, TM)) return e;
    if (int e = make_map(&tmFR, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, fr, 4, NF, (size_t)P, 32, TN)) return e;
    CvTcArgs a{};
    a.fl = fl; a.fr = fr;
    a.scaleL = scaleL; a.normL = normL; a.scaleR = scaleR; a.normR = normR;
    a.CL = CL; a.CR = CR;
    a.H = H; a.W = W; a.D = D; a.Dp = Dp; a.P = P;
    a.tiles_x = ceil_div(W, TM);
    a.nitems = a.tiles_x * H;
    a.queue = queue;
    a.queue_count = queue_count;
    a.queue_cap = (unsigned)(queue_cap > 0xffffffffu ? 0xffffffffu : queue_cap);
    MCCNN_CUDA(cudaFuncSetAttribute(cost_volume_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    int grid = sm_count();
    if (grid > a.nitems) grid = a.nitems;
    cost_volume_tc_kernel<<<grid, CVT_THREADS, TC_SMEM, stream>>>(tmSL, tmSR, tmFL, tmFR, a);
    MCCNN_LAUNCH_CHECK("cost_volume_tc_kernel");
    cv_fixup_kernel<<<sm_count() * 8, 128, 0, stream>>>(fl, fr, CL, CR, queue, queue_count, a.queue_cap, Dp);
    MCCNN_LAUNCH_CHECK("cv_fixup_kernel");
    return 0;
}
