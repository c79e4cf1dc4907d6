// Cost volume of the FUSED mode on the tensor cores (sm_100a): a plain banded GEMM, no rounding proof.
//
// Same layout, fills and pads as compute_cost_volume_kernel's replacement in cost_volume.cu (process_functional.py:120-131,
// fill :1111-1114): CL[y][x][d] = CR[y][x-d][d] = -<fl[y][x], fr[y][x-d]>. What differs is the arithmetic: the reference
// accumulates fp32 products in fp64; here every feature is split into two fp16 numbers, f = hi + lo (22 significand bits),
// and the dot product is hi.hi + hi.lo + lo.hi accumulated in fp32 in tensor memory: |error| <= ~1e-6 on unit-norm features,
// inside north_star's 1e-4 but not the reference's bits -- MCCNN_SGM_FUSED only (the exact mode keeps cost_volume.cu /
// cost_volume_tc.cu).
//
// With u = x - d the volume of an image row is the band 0 <= x - u < D of C'[x][u] = -<fl[x], fr[u]>. A tile is 128 x pixels
// (TMEM lanes) by 128 u pixels (TMEM columns), K = 64 features = one 128-byte swizzled row per pixel, so a pixel's hi (or lo)
// features are exactly one row of a SWIZZLE_128B K-major operand tile and TMA loads a tile of 128 pixels as one 16 KB box.
// Per tile 12 tcgen05.mma (M 128, N 128, K 16): 4 k-steps x {hi.hi, hi.lo, lo.hi}. The kernel is bound by the 2 x 64 KB it
// writes per tile, not by the MMAs (0.4 us of tensor time against ~3 us of HBM time per tile and SM).
// Warp roles: warps 0..7 epilogue (TMEM -> shared staging tile -> both volumes as runs along d), warp 8 control (TMA, MMA issue).
// The epilogue warps form two independent groups of four (one per TMEM lane quarter), each with its own staging tile for 64 of
// the tile's 128 columns: the groups drift apart, so one group's TMEM reads overlap the other's global stores.
// The B operands and the accumulators are double-buffered, so the loads and MMAs of tile n + 1 run under the stores of tile n.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;
constexpr int TM = 128, TN = 128;
constexpr int TILE_BYTES = 128 * 128;             // one fp16 operand tile: 128 pixels x 128 bytes
constexpr int OFF_A = 0;                          // hi, lo of the x block
constexpr int OFF_B = 2 * TILE_BYTES;             // two stages of {hi, lo} of a u block
constexpr int GC = 64;                            // tile columns per epilogue group
constexpr int RES_PITCH = GC + 1;
constexpr int RES_GROUP = TM * RES_PITCH;         // floats of one group's staging tile [128][65]
constexpr int OFF_RES = OFF_B + 4 * TILE_BYTES;   // two staging tiles
constexpr int OFF_BAR = OFF_RES + ((2 * RES_GROUP * 4 + 127) & ~127);
constexpr int FT_SMEM = OFF_BAR + 128 + 1024;     // barriers + alignment slack
constexpr int NEW = 8;                            // epilogue warps
constexpr int FT_THREADS = 32 * (NEW + 1);
constexpr uint32_t TMEM_COLS = 256;               // two accumulators of 128 columns
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);  // f16 x f16 -> f32
constexpr float kInfF = __builtin_huge_valf();

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the four warps of one epilogue group (named barriers 1 and 2)
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// ---------------------------------------------------------------- pre-pass: f = hi + lo, both fp16; planes [2][P][64]
__global__ void __launch_bounds__(256) cv_split_kernel(const float* __restrict__ feat, __half* __restrict__ planes, long long P) {
    const long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8;   // 8 features per thread
    if (i >= P * NF) return;
    const float4 v0 = *reinterpret_cast<const float4*>(feat + i);
    const float4 v1 = *reinterpret_cast<const float4*>(feat + i + 4);
    const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        hi[k] = __float2half_rn(f[k]);
        lo[k] = __float2half_rn(f[k] - __half2float(hi[k]));
    }
    *reinterpret_cast<uint4*>(planes + i) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(planes + P * NF + i) = *reinterpret_cast<const uint4*>(lo);
}

// entries no evaluation writes: fill where the match falls outside the other image, +INF pads
__global__ void __launch_bounds__(256) cvf_fill_kernel(float* __restrict__ CL, float* __restrict__ CR, int W, int D, int Dp,
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

struct FtArgs {
    float* CL;
    float* CR;
    int H, W, D, Dp;
    long long P;
    int tiles_x, nitems;
};

// position in the tile sequence of one CTA: items (image row y, block of TM x pixels), each with its run of u tiles
struct Cursor {
    int item, stride, nitems;
    int y, x0, ut, ut_last;
    bool first;
    __device__ void enter(const FtArgs& a) {
        if (item >= nitems) return;
        y = item / a.tiles_x;
        x0 = (item % a.tiles_x) * TM;
        ut = max(x0 - (a.D - 1), 0) / TN;
        ut_last = min(x0 + TM - 1, a.W - 1) / TN;
        first = true;
    }
    __device__ void start(const FtArgs& a, int first_item, int step) {
        item = first_item; stride = step; nitems = a.nitems;
        enter(a);
    }
    __device__ bool valid() const { return item < nitems; }
    __device__ void next(const FtArgs& a) {
        if (ut < ut_last) {
            ut++;
            first = false;
        } else {
            item += stride;
            enter(a);
        }
    }
};

// Barriers: a_full (TMA bytes, one phase per item), and per stage b_full (TMA bytes), mma (tcgen05.commit: the tile's MMAs are
// complete, its operands can be overwritten and its accumulator read), t_free (NEW arrivals: the accumulator has been read).
__global__ void __launch_bounds__(FT_THREADS, 1)
cost_volume_fast_tc_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR, const FtArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bar_a_full = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* bar_b_full = bar_a_full + 1;   // [2]
    uint64_t* bar_mma = bar_a_full + 3;      // [2]
    uint64_t* bar_t_free = bar_a_full + 5;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_a_full + 7);
    float* res = reinterpret_cast<float*>(sm + OFF_RES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        mbar_init(bar_a_full, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(bar_b_full + s, 1);
            mbar_init(bar_mma + s, 1);
            mbar_init(bar_t_free + s, NEW);
        }
        mbar_fence_init();
    }
    if (warp == NEW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == NEW) {
        // ================================================================= control warp: TMA loads + MMA issue
        if (lane == 0) {
            Cursor ld, mm;  // the load cursor runs up to two tiles ahead of the MMA cursor
            ld.start(a, blockIdx.x, gridDim.x);
            mm.start(a, blockIdx.x, gridDim.x);
            uint32_t nl = 0, nm = 0, n_item = 0;
            while (mm.valid()) {
                while (ld.valid() && nl < nm + 2) {
                    if (ld.first && nl != nm) break;  // the A operand changes: every MMA of the previous item must be issued ...
                    const uint32_t s = nl & 1u;
                    const long long prow = (long long)ld.y * a.W;
                    if (ld.first) {
                        // ... and complete (tiles nl - 1 and nl - 2 are the last two that read the old A tile)
                        if (nl >= 1) mbar_wait(bar_mma + ((nl - 1) & 1u), ((nl - 1) >> 1) & 1u);
                        if (nl >= 2) mbar_wait(bar_mma + ((nl - 2) & 1u), ((nl - 2) >> 1) & 1u);
                        mbar_expect_tx(bar_a_full, 2 * TILE_BYTES);
                        tma_load_2d(base + OFF_A, &tmL, 0, (int)(prow + ld.x0), bar_a_full);
                        tma_load_2d(base + OFF_A + TILE_BYTES, &tmL, 0, (int)(a.P + prow + ld.x0), bar_a_full);
                    } else if (nl >= 2) {
                        mbar_wait(bar_mma + s, ((nl - 2) >> 1) & 1u);   // the MMAs that read this B stage are complete
                    }
                    const int u0 = ld.ut * TN;
                    const uint32_t bst = base + OFF_B + s * 2 * TILE_BYTES;
                    mbar_expect_tx(bar_b_full + s, 2 * TILE_BYTES);
                    tma_load_2d(bst, &tmR, 0, (int)(prow + u0), bar_b_full + s);
                    tma_load_2d(bst + TILE_BYTES, &tmR, 0, (int)(a.P + prow + u0), bar_b_full + s);
                    ld.next(a);
                    nl++;
                }
                const uint32_t s = nm & 1u;
                if (mm.first) {
                    mbar_wait(bar_a_full, n_item & 1u);
                    n_item++;
                }
                mbar_wait(bar_b_full + s, (nm >> 1) & 1u);
                if (nm >= 2) mbar_wait(bar_t_free + s, ((nm - 2) >> 1) & 1u);  // the accumulator of tile nm - 2 has been read
                tc_fence_after();
                const uint32_t bst = base + OFF_B + s * 2 * TILE_BYTES;
                const uint64_t a_hi = sw128_desc(base + OFF_A), a_lo = sw128_desc(base + OFF_A + TILE_BYTES);
                const uint64_t b_hi = sw128_desc(bst), b_lo = sw128_desc(bst + TILE_BYTES);
                const uint32_t acc = tmem_base + s * TN;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    umma_f16(acc, a_hi + 2 * k, b_hi + 2 * k, k != 0 ? 1u : 0u);
                    umma_f16(acc, a_hi + 2 * k, b_lo + 2 * k, 1u);
                    umma_f16(acc, a_lo + 2 * k, b_hi + 2 * k, 1u);
                }
                umma_commit(bar_mma + s);
                mm.next(a);
                nm++;
            }
        }
    } else {
        // ================================================================= epilogue warps
        // TMEM lane = tile row: thread owns x = x0 + 32 q + lane and the 64 tile columns u of its group g
        const int q = warp & 3, g = warp >> 2;
        const int xl = 32 * q + lane;
        float* resg = res + g * RES_GROUP;
        Cursor tc;
        tc.start(a, blockIdx.x, gridDim.x);
        uint32_t n = 0;
        for (; tc.valid(); tc.next(a), n++) {
            const uint32_t s = n & 1u;
            const int x0 = tc.x0, u0 = tc.ut * TN + g * GC;   // first u of this group's columns
            const long long prow = (long long)tc.y * a.W;
            mbar_wait(bar_mma + s, (n >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; c++) {
                float v[32];
                tmem_ld32(tmem_base + s * TN + (uint32_t)(g * GC + c * 32) + ((uint32_t)(32 * q) << 16), v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) resg[xl * RES_PITCH + c * 32 + j] = -v[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t_free + s);   // this warp has read its part of the accumulator
            group_sync(g);                                // the group's staging tile is complete
            // CL[y][x][d], d = x - u: for a fixed x the group's u range is a contiguous run of d
            for (int r = q; r < TM; r += 4) {
                const int x = x0 + r;
                if (x >= a.W) break;
                float* row = a.CL + (prow + x) * a.Dp;
#pragma unroll
                for (int h = 0; h < GC / 32; h++) {
                    const int ul = (GC - 1) - (lane + 32 * h);  // descending u = ascending d
                    const int u = u0 + ul, d = x - u;
                    if (u >= 0 && d >= 0 && d < a.D) row[d] = resg[r * RES_PITCH + ul];
                }
            }
            // CR[y][u][d], d = x - u: for a fixed u its x range is a contiguous run of d
            if (a.CR != nullptr) {
                for (int ul = q; ul < GC; ul += 4) {
                    const int u = u0 + ul;
                    if (u < 0) continue;
                    if (u >= a.W) break;
                    float* row = a.CR + (prow + u) * a.Dp;
#pragma unroll
                    for (int h = 0; h < TM / 32; h++) {
                        const int r = lane + 32 * h;
                        const int x = x0 + r, d = x - u;
                        if (x < a.W && d >= 0 && d < a.D) row[d] = resg[r * RES_PITCH + ul];
                    }
                }
            }
            group_sync(g);                                // the staging tile may be overwritten
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEW) {
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

int make_map(CUtensorMap* tm, const void* ptr, size_t rows) {
    EncodeTiledFn enc = get_encode();
    MCCNN_REQUIRE(enc != nullptr, MCCNN_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)NF, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)NF * 2};
    cuuint32_t box[2] = {(cuuint32_t)NF, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCCNN_REQUIRE(r == CUDA_SUCCESS, MCCNN_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_cost_volume_fast_tc_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    return 2 * a256((size_t)2 * H * W * NF * sizeof(__half));
}

extern "C" int mccnn_cost_volume_fast_tc(const float* fl, const float* fr, float* CL, float* CR, void* workspace,
                                         size_t workspace_bytes, int H, int W, int D, float fill, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(fl && fr && CL && workspace, MCCNN_EINVAL, "mccnn_cost_volume_fast_tc: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096, MCCNN_EINVAL, "mccnn_cost_volume_fast_tc: bad shape H=%d W=%d D=%d", H, W, D);
    MCCNN_REQUIRE((long long)2 * H * W + TM < 0x7fffffffLL, MCCNN_EINVAL, "mccnn_cost_volume_fast_tc: image too large for 32-bit tile rows");
    MCCNN_REQUIRE(aligned16(fl) && aligned16(fr) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN,
                  "mccnn_cost_volume_fast_tc: features must be 16-byte, the workspace 256-byte aligned");
    MCCNN_REQUIRE(workspace_bytes >= mccnn_cost_volume_fast_tc_workspace_bytes(H, W), MCCNN_EWORKSPACE,
                  "mccnn_cost_volume_fast_tc: workspace too small");
    const long long P = (long long)H * W;
    const int Dp = disp_pitch(D);
    char* ws = reinterpret_cast<char*>(workspace);
    const size_t plane_bytes = a256((size_t)2 * P * NF * sizeof(__half));
    __half* pl = reinterpret_cast<__half*>(ws);
    __half* pr = reinterpret_cast<__half*>(ws + plane_bytes);
    const unsigned nb = (unsigned)((P * NF / 8 + 255) / 256);
    cv_split_kernel<<<nb, 256, 0, stream>>>(fl, pl, P);
    MCCNN_LAUNCH_CHECK("cv_split_kernel");
    cv_split_kernel<<<nb, 256, 0, stream>>>(fr, pr, P);
    MCCNN_LAUNCH_CHECK("cv_split_kernel");
    cvf_fill_kernel<<<(unsigned)((P + 7) / 8), 256, 0, stream>>>(CL, CR, W, D, Dp, P, fill);
    MCCNN_LAUNCH_CHECK("cvf_fill_kernel");
    CUtensorMap tmL, tmR;
    if (int e = make_map(&tmL, pl, (size_t)2 * P)) return e;
    if (int e = make_map(&tmR, pr, (size_t)2 * P)) return e;
    FtArgs a{};
    a.CL = CL; a.CR = CR;
    a.H = H; a.W = W; a.D = D; a.Dp = Dp; a.P = P;
    a.tiles_x = ceil_div(W, TM);
    a.nitems = a.tiles_x * H;
    if (int e = kernel_setup<cost_volume_fast_tc_kernel>(FT_THREADS, FT_SMEM, nullptr)) return e;
    int grid = sm_count();
    if (grid > a.nitems) grid = a.nitems;
    cost_volume_fast_tc_kernel<<<grid, FT_THREADS, FT_SMEM, stream>>>(tmL, tmR, a);
    MCCNN_LAUNCH_CHECK("cost_volume_fast_tc_kernel");
    return 0;
}
