// Conv tower layers 2..n as an implicit GEMM on the 5th-generation tensor cores (sm_100a):
// tcgen05.mma with accumulators in TMEM, operands staged in shared memory by TMA.
//
// Replaces the tf.nn.conv2d + bias + relu / l2_normalize calls of conv() and Net.construct
// (mc_cnn_brunch.py:38-48, 70-92) that compute_feature evaluates with sess.run
// (process_functional.py:38-39).
//
// GEMM view of one 3x3 VALID layer, 64 -> 64 channels, activations NHWC:
//   M = 128 consecutive output pixels of one image row, N = 64 output channels,
//   K = 9 taps x 64 input channels. For tap (ky,kx) the A operand is the [128 pixels][64 channels]
//   block starting at input pixel (y+ky, x0+kx): contiguous in memory. One TMA 2-D tile load per ky
//   brings the 136-pixel window of that image row into the 128-byte-swizzled K-major layout tcgen05
//   wants; the three kx taps read it through descriptors whose start is shifted by kx rows (128 B).
//   The 128B swizzle is a function of the shared-memory ADDRESS bits (measured: the shifted start
//   needs base_offset 0; base_offset = kx gives wrong results), so the rows TMA wrote are found
//   where the MMA looks for them. Every activation byte crosses L2->smem 3x, not 9x.
//   B = the tap's [64 cout][64 cin] weights, resident in shared memory for the whole (persistent) CTA.
//
// Precision: the reference runs fp32 convolutions. Every operand is split into two fp16 numbers,
// x = hi + lo / 2048 (22-bit significand, lo scaled so that it cannot underflow), and each k-step
// issues three MMAs: hi*hi, hi*lo and lo*hi into three fp32 TMEM accumulators; the
// epilogue forms main + (corr1 + corr2) / 2048. The dropped lo*lo term is 2^-22 relative. Activations travel
// between layers as two fp16 NHWC tensors (same bytes as fp32).
//
// Warp roles (320 threads, one CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM
// allocator, warps 2..9 = epilogue (TMEM lane quarter x channel half: TMEM -> registers -> bias/ReLU/split or
// l2-normalise -> global; with four epilogue warps the epilogue, not the MMAs, set the pace of a tile).
// Two smem stages for A, two TMEM accumulator sets so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstring>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;
constexpr int TILE_M = 128;
constexpr int WIN_ROWS = 136;               // 128 + 2 pixels of kx shift, rounded up to the 8-row swizzle atom
constexpr int A_TILE_BYTES = WIN_ROWS * 128;  // one image-row window: 136 pixels x 64 fp16
constexpr int B_TILE_BYTES = NF * 128;      // 64 cout x 64 fp16
constexpr int W_TC_BYTES = 2 * 9 * B_TILE_BYTES;  // hi + lo, 9 taps
constexpr int TC_STAGES = 2;
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES;  // hi + lo
constexpr int SMEM_A_OFF = W_TC_BYTES;         // 147456, a multiple of 1024
constexpr int SMEM_BAR_OFF = SMEM_A_OFF + TC_STAGES * STAGE_BYTES;
constexpr int SMEM_SS_OFF = SMEM_BAR_OFF + 128;       // l2-normalise: partial sums of squares of the two channel halves, [2 parities][2][128]
constexpr int TC_SMEM_BYTES = SMEM_SS_OFF + 2 * 2 * TILE_M * 4 + 1024;  // + alignment slack
constexpr int TC_EPI_WARPS = 8;             // epilogue warps: TMEM lane quarter x channel half (the epilogue, not the MMAs, bounded the tile: ncu)
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);
constexpr uint32_t TMEM_COLS = 512;  // 2 accumulator sets x 3 accumulators x 64 columns = 384, rounded up to a power of two
constexpr uint32_t IDESC_F16_M128_N64 = (1u << 4) | ((uint32_t)(NF >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
constexpr uint32_t IDESC_F16_M128_N128 = (1u << 4) | ((uint32_t)(2 * NF >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
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
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
// base_offset stays 0 also for starts shifted by whole 128-byte rows (see the header comment).
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr, uint32_t base_offset = 0) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(base_offset & 7u) << 49) | ((uint64_t)2 << 61);
}
// Issued from WARP-CONVERGENT code: every lane runs the instruction stream, elect.sync picks the one lane that issues.
// (Inside an `if (lane == 0)` region the compiler wraps every uniform-datapath instruction in an ELECT / BRA.U.ANY loop:
// ~75 cycles per MMA, twice the 37 cycles an M128 N64 K16 MMA takes: the issue, not the pipe, set the pace.)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}\n" ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void split_store(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn((v - __half2float(h)) * 2048.0f);
}

// ---------------------------------------------------------------- layer 1 (1 -> 64) on CUDA cores, fp16 hi/lo output
__global__ void __launch_bounds__(256) conv1_split_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ b, __half* __restrict__ out_hi,
                                                         __half* __restrict__ out_lo, int Hin, int Win) {
    // thread = 8 output channels (weights in registers) x a strided set of pixels; 8 threads share a pixel
    const int q = threadIdx.x & 7;
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; t++)
#pragma unroll
        for (int c = 0; c < 8; c++) wr[t][c] = __ldg(&w[t * NF + 8 * q + c]);
#pragma unroll
    for (int c = 0; c < 8; c++) br[c] = __ldg(&b[8 * q + c]);
    const int Hout = Hin - 2, Wout = Win - 2;
    const size_t npix = (size_t)Hout * Wout;
    for (size_t pix = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3); pix < npix; pix += (size_t)gridDim.x * 32) {
        const int y = (int)(pix / Wout), x = (int)(pix % Wout);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; c++) acc[c] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) {
                const float v = __ldg(&in[(size_t)(y + ky) * Win + x + kx]);
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = fmaf(v, wr[ky * 3 + kx][c], acc[c]);
            }
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; c++) split_store(fmaxf(acc[c] + br[c], 0.f), hi[c], lo[c]);
        *reinterpret_cast<uint4*>(&out_hi[pix * NF + 8 * q]) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(&out_lo[pix * NF + 8 * q]) = *reinterpret_cast<const uint4*>(lo);
    }
}

// ---------------------------------------------------------------- layers 2..n on tcgen05
struct TcArgs {
    const unsigned char* w_tc;  // [hi: 9 x 8 KB][lo: 9 x 8 KB], rows pre-swizzled (128B pattern)
    const float* bias;
    __half* out_hi;
    __half* out_lo;
    float* out_f32;
    int Hin, Win;  // input activation size; output is (Hin-2) x (Win-2)
    int tiles_per_row, ntiles;
};

template <bool LAST>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SMEM_BAR_OFF);
    uint64_t* full = bars;                    // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;   // [2]
    uint64_t* tempty = bars + 2 * TC_STAGES + 2;  // [2]
    uint64_t* wbar = bars + 2 * TC_STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], TC_EPI_WARPS); }
        mbar_init(wbar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int Wout = a.Win - 2;

    if (warp == 0) {
        // ===== TMA producer
        if (lane == 0) {
            mbar_expect_tx(wbar, W_TC_BYTES);
            // per tap the hi tile and the lo tile side by side: [64 cout hi][64 cout lo] is ONE N = 128 B operand
            for (int tap = 0; tap < 9; tap++) {
                bulk_g2s_addr(base + tap * 2 * B_TILE_BYTES, a.w_tc + (size_t)tap * B_TILE_BYTES, B_TILE_BYTES, wbar);
                bulk_g2s_addr(base + tap * 2 * B_TILE_BYTES + B_TILE_BYTES, a.w_tc + (size_t)(9 + tap) * B_TILE_BYTES, B_TILE_BYTES, wbar);
            }
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                const int y = tile / a.tiles_per_row, x0 = (tile % a.tiles_per_row) * TILE_M;
                for (int ky = 0; ky < 3; ky++, it++) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    const int pix = (y + ky) * a.Win + x0;
                    const uint32_t dst = base + SMEM_A_OFF + s * STAGE_BYTES;
                    tma_load_2d(dst, &tm_hi, 0, pix, &full[s]);
                    tma_load_2d(dst + A_TILE_BYTES, &tm_lo, 0, pix, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp runs the loop (convergent), one elected lane issues
        {
            mbar_wait(wbar, 0);
            uint32_t it = 0, tc = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tc++) {
                const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
                mbar_wait(&tempty[acc], aph ^ 1u);
                tc_fence_after();
                // three accumulators per tile (hi*hi, hi*lo, lo*hi): consecutive MMAs never target the same TMEM columns, so the
                // tensor pipe does not wait for an accumulate to land (with hi*lo and lo*hi in ONE accumulator the pipe was busy
                // 49 % of the time although neither loads nor the epilogue held the MMA warp back: ncu)
                const uint32_t d_main = tmem_base + acc * 192u, d_c2 = d_main + 128u;
                for (int ky = 0; ky < 3; ky++, it++) {
                    const uint32_t s = it % TC_STAGES, ph = (it / TC_STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_addr = base + SMEM_A_OFF + s * STAGE_BYTES;
#pragma unroll
                    for (int kx = 0; kx < 3; kx++) {
                        const int tap = ky * 3 + kx;
                        // same window, start shifted by kx pixels (= kx 128-byte rows)
                        const uint64_t a_hi = sw128_desc(a_addr + kx * 128);
                        const uint64_t a_lo = sw128_desc(a_addr + A_TILE_BYTES + kx * 128);
                        const uint64_t b_hilo = sw128_desc(base + tap * 2 * B_TILE_BYTES);  // rows 0..63 hi, 64..127 lo
#pragma unroll
                        for (int k = 0; k < 4; k++) {  // UMMA_K = 16 fp16 = 32 bytes = 2 descriptor units
                            const uint32_t first = (tap | k) != 0 ? 1u : 0u;
                            // a_hi x [b_hi | b_lo] -> columns 0..63 = hi*hi (main), 64..127 = hi*lo; a_lo x b_hi -> third accumulator
                            umma_f16(d_main, a_hi + 2 * k, b_hilo + 2 * k, first, IDESC_F16_M128_N128);
                            umma_f16(d_c2, a_lo + 2 * k, b_hilo + 2 * k, first, IDESC_F16_M128_N64);
                        }
                    }
                    umma_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
                }
                umma_commit(&tfull[acc]);  // accumulators of this tile are complete
            }
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31; two warps per lane quarter, 32 output channels each
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        constexpr int CH = NF / 2;
        float bias[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) bias[j] = __ldg(&a.bias[CH * half + j]);
        float* ssbuf = reinterpret_cast<float*>(sm + SMEM_SS_OFF);
        uint32_t tc = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tc++) {
            const uint32_t acc = tc & 1u, aph = (tc >> 1) & 1u;
            const int y = tile / a.tiles_per_row, x0 = (tile % a.tiles_per_row) * TILE_M;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 192u + CH * half;
            float v[CH];
#pragma unroll
            for (int c = 0; c < CH / 16; c++) {
                float mn[16], c1[16], c2[16];
                tmem_ld16(taddr + c * 16, mn);
                tmem_ld16(taddr + 64 + c * 16, c1);
                tmem_ld16(taddr + 128 + c * 16, c2);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) v[c * 16 + j] = fmaf(c1[j] + c2[j], 1.0f / 2048.0f, mn[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);

            const int x = x0 + m;
            const size_t o = ((size_t)y * Wout + min(x, Wout - 1)) * NF + CH * half;
            if (!LAST) {
                if (x < Wout) {
#pragma unroll
                    for (int c8 = 0; c8 < CH / 8; c8++) {
                        __align__(16) __half hi[8], lo[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) split_store(fmaxf(v[c8 * 8 + j] + bias[c8 * 8 + j], 0.f), hi[j], lo[j]);
                        *reinterpret_cast<uint4*>(&a.out_hi[o + c8 * 8]) = *reinterpret_cast<const uint4*>(hi);
                        *reinterpret_cast<uint4*>(&a.out_lo[o + c8 * 8]) = *reinterpret_cast<const uint4*>(lo);
                    }
                }
            } else {
                // l2-normalise over all 64 channels: the two halves exchange their sums of squares (double-buffered by
                // tile parity: a warp is at most one barrier ahead of the slowest one)
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < CH; j++) {
                    v[j] += bias[j];
                    ss = fmaf(v[j], v[j], ss);
                }
                float* sb = ssbuf + (tc & 1u) * 2 * TILE_M;
                sb[half * TILE_M + m] = ss;
                asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
                const float tot = sb[m] + sb[TILE_M + m];  // same order in both halves: identical scale
                const float s = 1.0f / sqrtf(fmaxf(tot, 1e-12f));
                if (x < Wout) {
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; c4++)
                        *reinterpret_cast<float4*>(&a.out_f32[o + c4 * 4]) =
                            make_float4(v[c4 * 4] * s, v[c4 * 4 + 1] * s, v[c4 * 4 + 2] * s, v[c4 * 4 + 3] * s);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
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

int make_act_map(CUtensorMap* tm, const __half* ptr, size_t npix) {
    EncodeTiledFn enc = get_encode();
    MCCNN_REQUIRE(enc != nullptr, MCCNN_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)NF, (cuuint64_t)npix};
    cuuint64_t strides[1] = {(cuuint64_t)NF * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)NF, (cuuint32_t)WIN_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCCNN_REQUIRE(r == CUDA_SUCCESS, MCCNN_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

}  // namespace

// fp16 hi/lo images of one layer's weights in the smem layout the kernel copies verbatim: per tap a
// [64 cout][64 cin] K-major tile, 128-byte rows, 16-byte chunks XOR-swizzled by (row % 8).
void pack_tc_weights(const float* hwio, unsigned char* dst) {
    for (int part = 0; part < 2; part++)
        for (int tap = 0; tap < 9; tap++)
            for (int n = 0; n < NF; n++)
                for (int k = 0; k < NF; k++) {
                    const float w = hwio[((size_t)tap * NF + k) * NF + n];
                    const __half h = __float2half_rn(w);
                    const __half l = __float2half_rn((w - __half2float(h)) * 2048.0f);
                    const size_t off = (size_t)part * 9 * B_TILE_BYTES + (size_t)tap * B_TILE_BYTES + (size_t)(n / 8) * 1024 +
                                       (size_t)(n % 8) * 128 + (size_t)(((k / 8) ^ (n % 8)) * 16) + (size_t)(k % 8) * 2;
                    const __half val = part == 0 ? h : l;
                    memcpy(dst + off, &val, 2);
                }
}

size_t tc_weight_bytes_per_layer() { return W_TC_BYTES; }

int conv_tower_tc(const float* padded, const float* w_fp32, const size_t* layer_off_floats, const unsigned char* w_tc,
                  float* features, void* workspace, size_t half_bytes, int H, int W, int num_layers, cudaStream_t stream) {
    // workspace: two ping-pong activation sets, each = [hi fp16 | lo fp16] of the largest layer output
    __half* buf[2][2];
    for (int i = 0; i < 2; i++) {
        buf[i][0] = reinterpret_cast<__half*>(reinterpret_cast<char*>(workspace) + i * half_bytes);
        buf[i][1] = reinterpret_cast<__half*>(reinterpret_cast<char*>(workspace) + i * half_bytes + half_bytes / 2);
    }
    int Hin = H + 2 * num_layers, Win = W + 2 * num_layers;
    {
        const size_t npix = (size_t)(Hin - 2) * (Win - 2);
        const float* w = w_fp32 + layer_off_floats[0];
        size_t blocks = (npix + 31) / 32;
        if (blocks > (size_t)sm_count() * 16) blocks = (size_t)sm_count() * 16;
        conv1_split_kernel<<<(unsigned)blocks, 256, 0, stream>>>(padded, w, w + 9 * NF, buf[0][0], buf[0][1], Hin, Win);
        MCCNN_LAUNCH_CHECK("conv1_split_kernel");
        Hin -= 2; Win -= 2;
    }
    if (int e = kernel_setup<conv_tc_kernel<false>>(TC_THREADS, TC_SMEM_BYTES, nullptr)) return e;
    if (int e = kernel_setup<conv_tc_kernel<true>>(TC_THREADS, TC_SMEM_BYTES, nullptr)) return e;
    int cur = 0;
    for (int l = 1; l < num_layers; l++) {
        const bool last = (l == num_layers - 1);
        CUtensorMap tm_hi, tm_lo;
        const size_t npix_in = (size_t)Hin * Win;
        if (int e = make_act_map(&tm_hi, buf[cur][0], npix_in)) return e;
        if (int e = make_act_map(&tm_lo, buf[cur][1], npix_in)) return e;
        TcArgs a{};
        a.w_tc = w_tc + (size_t)(l - 1) * W_TC_BYTES;
        a.bias = w_fp32 + layer_off_floats[l] + (size_t)9 * NF * NF;
        a.out_hi = buf[cur ^ 1][0];
        a.out_lo = buf[cur ^ 1][1];
        a.out_f32 = features;
        a.Hin = Hin; a.Win = Win;
        a.tiles_per_row = ceil_div(Win - 2, TILE_M);
        a.ntiles = a.tiles_per_row * (Hin - 2);
        int grid = sm_count();
        if (grid > a.ntiles) grid = a.ntiles;
        if (last)
            conv_tc_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tm_hi, tm_lo, a);
        else
            conv_tc_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tm_hi, tm_lo, a);
        MCCNN_LAUNCH_CHECK("conv_tc_kernel");
        cur ^= 1;
        Hin -= 2; Win -= 2;
    }
    return 0;
}

}  // namespace mccnn
