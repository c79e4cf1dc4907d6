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
constexpr int TM = 128, TN = 32;          // tile: x pixels, u pixels
constexpr int A_SLICE = TM * 128, B_SLICE = TN * 128;  // bytes of one fp16 slice tile (128-byte rows)
constexpr int OFF_A = 0;                               // six slices of the x block
constexpr int OFF_A32 = OFF_A + NS * A_SLICE;          // 98304: two k-halves of [128 px][32 floats]
constexpr int OFF_B = OFF_A32 + 2 * TM * 128;          // 131072: two stages of {six slices, two k-halves of [32 px][32 floats]}
constexpr int OFF_B32 = NS * B_SLICE;                  // fp32 rows inside a B stage
constexpr int B_STAGE = OFF_B32 + 2 * TN * 128;        // 32768
constexpr int OFF_RES = OFF_B + 2 * B_STAGE;           // 196608: two stages of [128][32] floats (column ^ (row & 31))
constexpr int RES_STAGE = TM * TN * 4;                 // 16384
constexpr int OFF_SB = OFF_RES + 2 * RES_STAGE;        // 229376: two stages of {scale, norm * 2^-41} of the 32 u pixels (doubles)
constexpr int OFF_BAR = OFF_SB + 2 * 2 * TN * 8;       // 230400
constexpr int TC_SMEM = OFF_BAR + 128 + 1024;          // barriers + alignment slack
#ifndef MCCNN_CVTC_NRW
#define MCCNN_CVTC_NRW 8
#endif
constexpr int NRW = MCCNN_CVTC_NRW;  // residual warps (8 or 16)
constexpr int NWX = NRW / 4;       // residual warps along x; a warp covers 128 / NWX tile rows x 8 tile columns
constexpr int XW = TM / NWX;       // tile rows per residual warp
constexpr int TI = XW / 8;         // evaluations along x per residual thread
constexpr int NEW = 8;             // epilogue warps
constexpr int CVT_THREADS = 32 * (NRW + NEW + 1);  // + 1 control warp
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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __uint_as_float(r[i]);
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
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * NEW) : "memory"); }  // the epilogue warps

// position in the tile sequence of one CTA: items (image row y, block of TM x pixels), each with its run of u tiles
struct TileCursor {
    int item, stride, nitems;
    int y, x0, ut, ut_last;
    bool first;  // first tile of its item (the A operand changes)
    __device__ void enter(const CvTcArgs& a) {
        if (item >= nitems) return;
        y = item / a.tiles_x;
        x0 = (item % a.tiles_x) * TM;
        ut = max(x0 - (a.D - 1), 0) / TN;
        ut_last = min(x0 + TM - 1, a.W - 1) / TN;
        first = true;
    }
    __device__ void start(const CvTcArgs& a, int first_item, int step) {
        item = first_item; stride = step; nitems = a.nitems;
        enter(a);
    }
    __device__ bool valid() const { return item < nitems; }
    __device__ void next(const CvTcArgs& a) {
        if (ut < ut_last) {
            ut++;
            first = false;
        } else {
            item += stride;
            enter(a);
        }
    }
};

// Warp roles: warps 0..15 residuals (FP32 pipe), warps 16..23 epilogue (TMEM -> fp64 -> rounding test -> stores; warp 16+e owns
// TMEM lanes 32(e&3).. = tile rows x and the 16 tile columns u of half e>>2), warp 24 = control (TMA loads, MMA issue, TMEM allocation). The B operands, the
// TMEM accumulators and the residual / result tile are double-buffered (stage = tile & 1), so the residual loop of tile
// n + 1 runs while the epilogue warps finish tile n: the kernel is bound by the FP32 pipe alone.
// Barriers, per stage unless noted: b_full (TMA bytes), mma (tcgen05.commit), b_free (16 arrivals: the residual warps have read
// the fp32 B rows), t_free (8: TMEM has been read), res_full (16: residual sums are in shared memory), res_free (8: the
// result tile has been stored), a_full (TMA bytes, one phase per item).
__global__ void __launch_bounds__(CVT_THREADS, 1)
cost_volume_tc_kernel(const __grid_constant__ CUtensorMap tmSL, const __grid_constant__ CUtensorMap tmSR,
                      const __grid_constant__ CUtensorMap tmFL, const __grid_constant__ CUtensorMap tmFR, const CvTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bar_a_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_b_full = bar_a_full + 1;    // [2]
    uint64_t* bar_mma = bar_a_full + 3;       // [2]
    uint64_t* bar_b_free = bar_a_full + 5;    // [2]
    uint64_t* bar_t_free = bar_a_full + 7;    // [2]
    uint64_t* bar_res_full = bar_a_full + 9;  // [2]
    uint64_t* bar_res_free = bar_a_full + 11; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a_full + 13);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_a_full, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_b_full + s, 1);
            mbar_init(bar_mma + s, 1);
            mbar_init(bar_b_free + s, NRW);
            mbar_init(bar_t_free + s, NEW);
            mbar_init(bar_res_full + s, NRW);
            mbar_init(bar_res_free + s, NEW);
        }
        mbar_fence_init();
    }
    if (warp == NRW + NEW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == NRW + NEW) {
        // ================================================================= control warp
        if (lane == 0) {
            TileCursor ld, mm;  // load cursor runs up to two tiles ahead of the MMA cursor
            ld.start(a, blockIdx.x, gridDim.x);
            mm.start(a, blockIdx.x, gridDim.x);
            uint32_t nl = 0, nm = 0, n_item = 0;
            // tile k has left its operand stage: its fp32 rows are read and its MMAs are complete
            auto wait_tile_done = [&](uint32_t k) {
                mbar_wait(bar_b_free + (k & 1u), (k >> 1) & 1u);
                mbar_wait(bar_mma + (k & 1u), (k >> 1) & 1u);
            };
            while (mm.valid()) {
                while (ld.valid() && nl < nm + 2) {
                    if (ld.first && nl != nm) break;  // the A operand changes: every MMA of the previous item must be issued
                    const uint32_t s = nl & 1u;
                    const long long prow = (long long)ld.y * a.W;
                    if (ld.first) {
                        if (nl >= 1) wait_tile_done(nl - 1);
                        if (nl >= 2) wait_tile_done(nl - 2);
                        mbar_expect_tx(bar_a_full, NS * A_SLICE + 2 * TM * 128);
                        for (int sl = 0; sl < NS; sl++)
                            tma_load_2d(base + OFF_A + sl * A_SLICE, &tmSL, 0, (int)((long long)sl * a.P + prow + ld.x0), bar_a_full);
                        tma_load_2d(base + OFF_A32, &tmFL, 0, (int)(prow + ld.x0), bar_a_full);
                        tma_load_2d(base + OFF_A32 + TM * 128, &tmFL, 32, (int)(prow + ld.x0), bar_a_full);
                    } else if (nl >= 2) {
                        wait_tile_done(nl - 2);
                    }
                    const int u0 = ld.ut * TN;
                    const uint32_t bst = base + OFF_B + s * B_STAGE;
                    mbar_expect_tx(bar_b_full + s, NS * B_SLICE + 2 * TN * 128);
                    for (int sl = 0; sl < NS; sl++)
                        tma_load_2d(bst + sl * B_SLICE, &tmSR, 0, (int)((long long)sl * a.P + prow + u0), bar_b_full + s);
                    tma_load_2d(bst + OFF_B32, &tmFR, 0, (int)(prow + u0), bar_b_full + s);
                    tma_load_2d(bst + OFF_B32 + TN * 128, &tmFR, 32, (int)(prow + u0), bar_b_full + s);
                    ld.next(a);
                    nl++;
                }
                const uint32_t s = nm & 1u;
                if (mm.first) {
                    mbar_wait(bar_a_full, n_item & 1u);
                    n_item++;
                }
                mbar_wait(bar_b_full + s, (nm >> 1) & 1u);
                if (nm >= 2) mbar_wait(bar_t_free + s, ((nm - 2) >> 1) & 1u);  // the accumulators of tile nm - 2 have been read
                tc_fence_after();
                const uint32_t bst = base + OFF_B + s * B_STAGE;
#pragma unroll 1
                for (int v = 0; v < NACC; v++) {
                    for (int sl = 0; sl <= v; sl++) {
                        const int t = v - sl;
                        const uint64_t ad = sw128_desc(base + OFF_A + sl * A_SLICE);
                        const uint64_t bd = sw128_desc(bst + t * B_SLICE);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            umma_f16(tmem_base + s * (NACC * TN) + (uint32_t)v * TN, ad + 2 * k, bd + 2 * k, (sl | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(bar_mma + s);
                mm.next(a);
                nm++;
            }
        }
    } else if (warp < NRW) {
        // ================================================================= residual warps
        // residual tile of a thread: x = XW*wx + tx + 8i (i < TI), u = 8*wy + tyl + 4j (j < 2)
        const int tx = lane & 7, tyl = lane >> 3, wx = warp % NWX, wy = warp / NWX;
        const int ub = 8 * wy + tyl;
        TileCursor tc;
        tc.start(a, blockIdx.x, gridDim.x);
        uint32_t n = 0, n_item = 0;
        for (; tc.valid(); tc.next(a), n++) {
            const uint32_t s = n & 1u;
            const int x0 = tc.x0, u0 = tc.ut * TN;
            if (tc.first) {
                mbar_wait(bar_a_full, n_item & 1u);
                n_item++;
            }
            mbar_wait(bar_b_full + s, (n >> 1) & 1u);
            // ---- fp32 rounding residuals of the 64 products of each of the TI x 2 evaluations of this thread
            float2 es[TI][2];
#pragma unroll
            for (int i = 0; i < TI; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) es[i][j] = make_float2(0.f, 0.f);
            // this warp's block is x in [XW wx, +XW) x u in [8 wy, +8): skipped when it lies outside the band 0 <= d < D
            const int wd_max = (x0 + XW * wx + XW - 1) - (u0 + 8 * wy), wd_min = (x0 + XW * wx) - (u0 + 8 * wy + 7);
            if (wd_max >= 0 && wd_min < a.D && x0 + XW * wx < a.W && u0 + 8 * wy < a.W) {
                const unsigned char* ap = sm + OFF_A32 + (XW * wx + tx) * 128;  // + i * 1024: rows tx + 8i keep (row & 7) = tx
                const unsigned char* bp = sm + OFF_B + s * B_STAGE + OFF_B32 + ub * 128;  // + j * 512: rows ub + 4j, (row & 7) = tyl + 4j
#pragma unroll 4
                for (int kc = 0; kc < 16; kc++) {  // 4 features per step: 16-byte chunk kc & 7 of k half kc >> 3
                    const int c = kc & 7;
                    const unsigned char* apk = ap + (kc >> 3) * (TM * 128) + ((c ^ tx) << 4);
                    const unsigned char* bpk = bp + (kc >> 3) * (TN * 128);
                    float4 bv[2];
#pragma unroll
                    for (int j = 0; j < 2; j++) bv[j] = *reinterpret_cast<const float4*>(bpk + j * 512 + ((c ^ (tyl + 4 * j)) << 4));
#pragma unroll
                    for (int i = 0; i < TI; i++) {
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
            if (n >= 2) mbar_wait(bar_res_free + s, ((n - 2) >> 1) & 1u);  // the results of tile n - 2 have left the buffer
            float* res = reinterpret_cast<float*>(sm + OFF_RES + s * RES_STAGE);
#pragma unroll
            for (int i = 0; i < TI; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int r = XW * wx + tx + 8 * i;
                    res[r * TN + ((ub + 4 * j) ^ (r & 31))] = es[i][j].x + es[i][j].y;
                }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_res_full + s);  // residual sums of this warp are in shared memory
                mbar_arrive(bar_b_free + s);    // the fp32 rows of this tile are no longer read
            }
        }
    } else {
        // ================================================================= epilogue warps
        // row / columns of a thread (TMEM lane = tile row): x = x0 + 32*q + lane, u = u0 + 16*half + 0..15
        const int ew = warp - NRW, q = ew & 3, half = ew >> 2, et = tid - 32 * NRW;
        const int xl = 32 * q + lane;
        TileCursor tc;
        tc.start(a, blockIdx.x, gridDim.x);
        uint32_t n = 0;
        double sa = 0.0, na = 0.0;
        for (; tc.valid(); tc.next(a), n++) {
            const uint32_t s = n & 1u;
            const int x0 = tc.x0, u0 = tc.ut * TN;
            const long long prow = (long long)tc.y * a.W;
            const int x = x0 + xl;
            const bool x_ok = x < a.W;
            if (tc.first) {
                sa = (double)a.scaleL[PADPIX + prow + min(x, a.W - 1)];
                na = (double)a.normL[PADPIX + prow + min(x, a.W - 1)];
            }
            double* sbd = reinterpret_cast<double*>(sm + OFF_SB) + s * (2 * TN);  // [0..31] scale, [32..63] norm * 2^-41 of the u pixels
            if (et < 2 * TN) {
                const float vv = (et < TN ? a.scaleR : a.normR)[PADPIX + prow + u0 + (et & (TN - 1))];
                sbd[et] = et < TN ? (double)vv : (double)vv * 0x1p-41;
            }
            mbar_wait(bar_res_full + s, (n >> 1) & 1u);
            mbar_wait(bar_mma + s, (n >> 1) & 1u);
            tc_fence_after();
            epi_sync();  // sbd is complete
            float* res = reinterpret_cast<float*>(sm + OFF_RES + s * RES_STAGE);
            float* myrow = res + xl * TN;  // entry jj lives at column jj ^ lane
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + s * (NACC * TN);
            // evaluations jj of this thread with 0 <= d < D and u < W: jj in [j_first, j_last]; d = dbase - jj
            const int dbase = x - u0;
            const int j_first = x_ok ? max(0, dbase - (a.D - 1)) : TN;
            const int j_last = min(TN - 1, min(dbase, a.W - 1 - u0));
            float* crp = a.CR ? a.CR + (prow + u0) * a.Dp + dbase : nullptr;  // CR[y][u][d] of jj = 0; +Dp-1 per jj
            unsigned unproven = 0;
#pragma unroll 1
            for (int c = 4 * half; c < 4 * half + 4; c++) {
                // nothing of this chunk is inside the band for any lane (tcgen05.ld is warp-collective: uniform test)
                if (__all_sync(0xffffffffu, j_last < 4 * c || j_first > 4 * c + 3)) continue;
                float r[NACC][4];
#pragma unroll
                for (int v = 0; v < NACC; v++) tmem_ld4(taddr + (uint32_t)v * TN + 4u * c, r[v]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // the six accumulators hold exact integers below 2^23: made int32 by a magic-number add, merged pairwise
                    // (hi * 256 + lo), int32 -> fp64 by bit pasting (2^52 + 2^31 + n): no conversion-pipe instruction
                    double dsum = 0.0;
#pragma unroll
                    for (int g = NACC / 2 - 1; g >= 0; g--) {  // smallest weight first
                        const float mh = r[2 * g][j] + 12582912.0f, ml = r[2 * g + 1][j] + 12582912.0f;
                        const int nn = (__float_as_int(mh) - 0x4B400000) * 256 + (__float_as_int(ml) - 0x4B400000);
                        const double nd = __hiloint2double(0x43300000, nn ^ (int)0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
                        const double wg = __longlong_as_double((long long)(1023 - 8 * (2 * g + 3)) << 52);       // 2^-8(2g+3)
                        dsum = fma(nd, wg, dsum);
                    }
                    const int jj = 4 * c + j;
                    const double sc = sa * sbd[jj];  // 2^(Ea+Eb), NaN for a wild pixel
                    const double T = fma(dsum, sc, -(double)myrow[jj ^ lane]);
                    const double eps = fma(na, sbd[TN + jj], fma(sc, 0x1p-40, 0x1p-140));
                    const float lo = __double2float_rn(T - eps), hi = __double2float_rn(T + eps);
                    const bool valid = jj >= j_first && jj <= j_last;
                    if (valid && !(lo == hi)) unproven |= 1u << jj;  // no branch here: the four fp64 chains of a chunk interleave
                    myrow[jj ^ lane] = -lo;
                    // CR[y][u][d]: for a fixed u the lanes of a warp hold consecutive x = consecutive d
                    if (valid && crp != nullptr) crp[(long long)jj * (a.Dp - 1)] = -lo;
                }
            }
            // roundings that could not be proven (about 3 in 10^4): queued for the literal loop of cv_fixup_kernel, which
            // patches both volumes after this kernel (the value stored above is then a placeholder)
            while (unproven) {
                const int jj = __ffs(unproven) - 1;
                unproven &= unproven - 1;
                const unsigned slot = atomicAdd(a.queue_count, 1u);
                if (slot < a.queue_cap) {
                    a.queue[slot] = ((unsigned long long)(prow + x) << 12) | (unsigned)(dbase - jj);
                } else {
                    const float outv = exact_dot_global(a.fl + (prow + x) * NF, a.fr + (prow + x - (dbase - jj)) * NF);
                    myrow[jj ^ lane] = outv;
                    if (crp != nullptr) crp[(long long)jj * (a.Dp - 1)] = outv;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t_free + s);  // TMEM stage may take tile n + 2
            epi_sync();                                   // results complete in shared memory
            // ---- CL[y][x][d], d = x - u: for a fixed x the tile's u range is a contiguous run of d
            {
                const bool interior = x0 + TM <= a.W && u0 + TN <= a.W && x0 - (u0 + TN - 1) >= 0 && (x0 + TM - 1) - u0 < a.D;
                const int ul = (TN - 1) - lane;
                if (interior) {
                    // row r of the tile: CL[y][x0 + r][x0 + r - u0 - ul]; r = ew + 8k, so (r & 31) = (ew + 8k) & 31
                    float* p = a.CL + (prow + x0 + ew) * a.Dp + (x0 + ew - u0 - ul);
                    const float* rp = res + ew * TN;
                    const size_t step = (size_t)NEW * (a.Dp + 1);
#pragma unroll
                    for (int k = 0; k < TM / NEW; k++) {
                        *p = rp[ul ^ ((ew + NEW * k) & 31)];
                        p += step;
                        rp += NEW * TN;
                    }
                } else {
                    for (int r = ew; r < TM; r += NEW) {
                        const int xr = x0 + r;
                        if (xr >= a.W) break;
                        const int d0 = xr - u0 - ul;
                        if (d0 >= 0 && d0 < a.D && u0 + ul < a.W) a.CL[(prow + xr) * a.Dp + d0] = res[r * TN + (ul ^ (r & 31))];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_res_free + s);  // the result tile may be overwritten
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NRW + NEW) {
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
    if (int e = kernel_setup<cost_volume_tc_kernel>(CVT_THREADS, TC_SMEM, nullptr)) return e;
    int grid = sm_count();
    if (grid > a.nitems) grid = a.nitems;
    cost_volume_tc_kernel<<<grid, CVT_THREADS, TC_SMEM, stream>>>(tmSL, tmSR, tmFL, tmFR, a);
    MCCNN_LAUNCH_CHECK("cost_volume_tc_kernel");
    cv_fixup_kernel<<<sm_count() * 8, 128, 0, stream>>>(fl, fr, CL, CR, queue, queue_count, a.queue_cap, Dp);
    MCCNN_LAUNCH_CHECK("cv_fixup_kernel");
    return 0;
}
