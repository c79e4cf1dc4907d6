// Normalised dot-product cost volume over the disparity range (sm_100a).
//
// Replaces compute_cost_volume_kernel (process_functional.py:120-131) and the host-side
// np.ones fill of both volumes (:1111-1114).
//
// Arithmetic contract (pinned bit for bit by tests/golden): temp is fp64; every
// product fl*fr is an fp32 multiply that is widened and added in fp64, i = 0..63 in order; the store
// rounds -temp to fp32. CL[y][x][d] and CR[y][x-d][d] receive the same value; entries never written
// keep `fill`.
//
// Shape: with u = x - d (the right pixel) the volume of one image row is a BAND of the plain product
// matrix C'[x][u] = -<fl[x], fr[u]>, 0 <= x-u < D. A CTA computes one 64x64 (x,u) tile of the band as a
// small GEMM: both 64-pixel feature tiles are transposed into shared memory (k-major, 16-byte-chunk
// XOR swizzle), each thread owns a 4x4 register tile (two 128-bit shared loads per k for 16 products),
// and the finished tile goes through shared memory so that BOTH outputs are written as contiguous
// runs along d: for a fixed x the tile's u-range is a run of CL[y][x][.], for a fixed u its x-range
// is a run of CR[y][u][.]. Tiles past the band edges (u < 0, x >= W) only write `fill`; every element
// of both volumes is written exactly once, pad entries [D, Dp) as +INF (the SGM kernels rely on it).
//
// The exact contract costs one fp32->fp64 widening per product. F2F.F64.F32 issues at 16/clk/SM on
// B200 (profiles/r01_ubench_pipes.txt), which bounds a straightforward kernel far below HBM speed, so
// half of the products are widened on the integer pipe instead (3 ALU ops, bit-exact for normal
// numbers below 2), and the multiplies are issued two at a time (FMUL2). A CTA whose staged features are not all
// comfortably normal falls back to F2F for all. The kernel is issue-bound (ncu: 82 % issue slots busy); the
// alternatives measured in tools/ubench_cv_inner.cu (residual + DFMA formulation, DMMA) sit within 10 % of it
// because DFMA / DMMA share the FP32 FMA datapath on B200 (profiles/r01c_ubench_cv_inner.txt).
#include "common.cuh"
#include <type_traits>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;
constexpr int T = 64;  // tile edge in x and in u
constexpr int CV_THREADS = 256;
constexpr int RES_PITCH = T + 1;
constexpr float kInfF = __builtin_huge_valf();

struct CvSmem {
    union {
        struct {
            float a[NF][T];  // [k][x], chunk-swizzled
            float b[NF][T];  // [k][u]
        } op;
        float res[T][RES_PITCH];  // [x][u]
    };
};

// bit-exact fp32 -> fp64 widening of a NORMAL fp32 number with |p| < 2 on the integer pipe, 3 ALU operations:
// the biased exponent is <= 127, so its top bit is clear and the rebias 127 -> 1023 (+0x380) is a plain OR
__device__ __forceinline__ double widen_normal(float p) {
    const uint32_t x = __float_as_uint(p);
    const int32_t t = (int32_t)x >> 3;  // sign copies | exponent | mantissa >> 3
    const uint32_t hi = ((uint32_t)t & 0x8fffffffu) | 0x38000000u;
    return __hiloint2double((int)hi, (int)__funnelshift_l(0u, x, 29));
}

// features whose pairwise products are guaranteed normal, non-zero and below 1: 2^-60 <= |v| < 1
// (unit-norm features always are, unless one of them is exactly +-1 or 0)
__device__ __forceinline__ bool comfortably_normal(float v) {
    const uint32_t e = (__float_as_uint(v) >> 23) & 0xffu;
    return e >= 67u && e <= 126u;
}

// Stage 64 pixels x 64 features: global [pixel][k] -> shared [k][pixel], 4-pixel chunks XOR-swizzled by k/4.
// Out-of-image pixels get 1.0 (their results are never stored). Returns false if a value is not comfortably normal.
__device__ __forceinline__ bool stage_tile(float (*dst)[T], const float* __restrict__ feat, int y, int W, int p0, int tid) {
    bool ok = true;
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
        const int pl = (tid >> 4) + 16 * rep;  // pixel within tile
        const int k4 = tid & 15;               // features 4*k4 .. 4*k4+3
        const int p = p0 + pl;
        float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p >= 0 && p < W) v = *reinterpret_cast<const float4*>(&feat[((size_t)y * W + p) * NF + 4 * k4]);
        ok = ok && comfortably_normal(v.x) && comfortably_normal(v.y) && comfortably_normal(v.z) && comfortably_normal(v.w);
        const int col = ((((pl >> 2) ^ k4) & 15) << 2) | (pl & 3);  // swizzle key = (k / 4) & 15 = k4
        dst[4 * k4 + 0][col] = v.x;
        dst[4 * k4 + 1][col] = v.y;
        dst[4 * k4 + 2][col] = v.z;
        dst[4 * k4 + 3][col] = v.w;
    }
    return ok;
}

template <bool SPLIT>
__device__ __forceinline__ void tile_products(const CvSmem& sm, int tx, int ty, double (&acc)[4][4]) {
#pragma unroll 2
    for (int k = 0; k < NF; k++) {
        const int key = (k >> 2) & 15;
        const float4 av = *reinterpret_cast<const float4*>(&sm.op.a[k][((tx ^ key) & 15) << 2]);
        const float4 bv = *reinterpret_cast<const float4*>(&sm.op.b[k][((ty ^ key) & 15) << 2]);
        const float2 a01 = make_float2(av.x, av.y), a23 = make_float2(av.z, av.w);
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // two fp32 products per FMUL2 (each lane is an IEEE round-to-nearest multiply, like FMUL)
            const float2 p01 = __fmul2_rn(a01, make_float2(b[j], b[j]));
            const float2 p23 = __fmul2_rn(a23, make_float2(b[j], b[j]));
            const float p[4] = {p01.x, p01.y, p23.x, p23.y};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (SPLIT && ((i + j) & 1))
                    acc[i][j] += widen_normal(p[i]);  // integer pipe
                else
                    acc[i][j] += (double)p[i];  // F2F.F64.F32
            }
        }
    }
}

// MCCNN_SGM_FUSED's cost volume: the same band GEMM with plain fp32 FMA accumulation, k = 0..63 in order (one FFMA per
// product instead of FMUL + widening + DADD). |error| <= 64 * 2^-24 * sum|f g| <= 4e-6 for unit-norm features: inside
// north_star's 1e-4, but not the reference's bits.
__device__ __forceinline__ void tile_products_fast(const CvSmem& sm, int tx, int ty, float (&acc)[4][4]) {
#pragma unroll 4
    for (int k = 0; k < NF; k++) {
        const int key = (k >> 2) & 15;
        const float4 av = *reinterpret_cast<const float4*>(&sm.op.a[k][((tx ^ key) & 15) << 2]);
        const float4 bv = *reinterpret_cast<const float4*>(&sm.op.b[k][((ty ^ key) & 15) << 2]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(CV_THREADS) cost_volume_band_kernel(const float* __restrict__ fl,
                                                                     const float* __restrict__ fr,
                                                                     float* __restrict__ CL, float* __restrict__ CR, int H,
                                                                     int W, int D, int Dp, float fill, int ut_min) {
    __shared__ __align__(16) CvSmem sm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int y = blockIdx.z;
    const int x0 = blockIdx.y * T;
    // u tiles of this x tile hang off the diagonal: [x0 - T*t, +T), t = 0 .. ceil((D-1+T)/T)-1. Their union
    // [x0 - T*(nut-1), x0 + T) holds every u = x - d, 0 <= d < D, of the x tile with the fewest tiles (the staging
    // takes any start pixel), e.g. 14 instead of 15 tiles at D = 800 and 3 instead of 4 at D = 128.
    (void)ut_min;
    const int u0 = x0 - T * (int)blockIdx.x;
    if (CR == nullptr && x0 >= W) return;         // nothing to write
    const bool has_left = x0 < W, has_right = (u0 + T - 1 >= 0) && (u0 < W);
    const bool compute = has_left && has_right;

    const int tx = tid & 15, ty = tid >> 4;  // thread tile: x = x0 + 4tx + i, u = u0 + 4ty + j
    using acc_t = typename std::conditional<EXACT, double, float>::type;
    acc_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = (acc_t)0;

    if (compute) {
        bool ok = stage_tile(sm.op.a, fl, y, W, x0, tid);
        ok = stage_tile(sm.op.b, fr, y, W, u0, tid) && ok;
        if constexpr (EXACT) {
            const int all_ok = __syncthreads_and(ok ? 1 : 0);
            if (all_ok)
                tile_products<true>(sm, tx, ty, acc);
            else
                tile_products<false>(sm, tx, ty, acc);
        } else {
            __syncthreads();
            tile_products_fast(sm, tx, ty, acc);
        }
    }
    __syncthreads();  // operands are dead: the result tile aliases them
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int xl = 4 * tx + i, ul = 4 * ty + j;
            const int x = x0 + xl, u = u0 + ul;
            const bool valid = compute && x < W && u >= 0 && u < W;
            sm.res[xl][ul] = valid ? (float)(-acc[i][j]) : fill;
        }
    __syncthreads();

    // CL[y][x][d], d = x - u: for a fixed x the tile's u range is a contiguous run of d
    if (has_left) {
        for (int xl = warp; xl < T; xl += CV_THREADS / 32) {
            const int x = x0 + xl;
            if (x >= W) break;
            float* row = CL + ((size_t)y * W + x) * Dp;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int ul = (T - 1) - (lane + 32 * h);  // descending u = ascending d
                const int d = x - (u0 + ul);
                if (d >= 0 && d < D) row[d] = sm.res[xl][ul];
            }
            // the tile that holds d = D-1 for this x also writes the +INF pad
            const int dl = x - (u0 + T - 1), dh = x - u0;
            if (Dp > D && dl <= D - 1 && D - 1 <= dh && lane < Dp - D) row[D + lane] = kInfF;
        }
    }
    // CR[y][u][d], d = x - u: for a fixed u the tile's x range is a contiguous run of d
    if (CR != nullptr && has_right) {
        for (int ul = warp; ul < T; ul += CV_THREADS / 32) {
            const int u = u0 + ul;
            if (u < 0) continue;
            if (u >= W) break;
            float* row = CR + ((size_t)y * W + u) * Dp;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int xl = lane + 32 * h;
                const int d = (x0 + xl) - u;
                if (d >= 0 && d < D) row[d] = sm.res[xl][ul];
            }
            const int dl = x0 - u, dh = x0 + T - 1 - u;
            if (Dp > D && dl <= D - 1 && D - 1 <= dh && lane < Dp - D) row[D + lane] = kInfF;
        }
    }
}

__global__ void volume_to_dhw_kernel(const float* __restrict__ vol, float* __restrict__ out, int H, int W, int D, int Dp) {
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int xb = blockIdx.x * 32, db = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int x = xb + r, d = db + threadIdx.x;
        tile[r][threadIdx.x] = (x < W && d < D) ? vol[((size_t)y * W + x) * Dp + d] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int d = db + r, x = xb + threadIdx.x;
        if (x < W && d < D) out[((size_t)d * H + y) * W + x] = tile[threadIdx.x][r];
    }
}

}  // namespace
}  // namespace mccnn

using namespace mccnn;

static int cost_volume_launch(const float* fl, const float* fr, float* CL, float* CR, int H, int W, int D, float fill, bool exact,
                              void* stream_);

extern "C" int mccnn_cost_volume(const float* fl, const float* fr, float* CL, float* CR, int H, int W, int D, float fill,
                                 void* stream_) {
    return cost_volume_launch(fl, fr, CL, CR, H, W, D, fill, true, stream_);
}

extern "C" int mccnn_cost_volume_fast(const float* fl, const float* fr, float* CL, float* CR, int H, int W, int D, float fill,
                                      void* stream_) {
    return cost_volume_launch(fl, fr, CL, CR, H, W, D, fill, false, stream_);
}

static int cost_volume_launch(const float* fl, const float* fr, float* CL, float* CR, int H, int W, int D, float fill, bool exact,
                              void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(fl && fr && CL, MCCNN_EINVAL, "mccnn_cost_volume: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096, MCCNN_EINVAL, "mccnn_cost_volume: bad shape H=%d W=%d D=%d", H,
                  W, D);
    MCCNN_REQUIRE(H <= 65535, MCCNN_EINVAL, "mccnn_cost_volume: H=%d exceeds 65535", H);
    MCCNN_REQUIRE(aligned16(fl) && aligned16(fr), MCCNN_EALIGN, "mccnn_cost_volume: features must be 16-byte aligned");
    const int Dp = disp_pitch(D);
    // x tiles cover [0, W + D - 1) when CR is written (the tiles past the image edge write CR's fill entries);
    // each x tile meets at most ceil((D - 1 + 2T - 1) / T) u tiles of the band
    const int nxt = ceil_div(CR ? (W + D - 1) : W, T);
    const int nut = (D - 1 + T + T - 1) / T;
    MCCNN_REQUIRE(nxt <= 65535, MCCNN_EINVAL, "mccnn_cost_volume: image too wide");
    dim3 grid(nut, nxt, H);
    if (exact)
        cost_volume_band_kernel<true><<<grid, CV_THREADS, 0, stream>>>(fl, fr, CL, CR, H, W, D, Dp, fill, 0);
    else
        cost_volume_band_kernel<false><<<grid, CV_THREADS, 0, stream>>>(fl, fr, CL, CR, H, W, D, Dp, fill, 0);
    MCCNN_LAUNCH_CHECK("cost_volume_band_kernel");
    return 0;
}

extern "C" int mccnn_volume_to_dhw(const float* vol, float* out_dhw, int H, int W, int D, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(vol && out_dhw, MCCNN_EINVAL, "mccnn_volume_to_dhw: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && H <= 65535, MCCNN_EINVAL, "mccnn_volume_to_dhw: bad shape");
    dim3 grid(ceil_div(W, 32), ceil_div(D, 32), H);
    volume_to_dhw_kernel<<<grid, dim3(32, 8), 0, stream>>>(vol, out_dhw, H, W, D, disp_pitch(D));
    MCCNN_LAUNCH_CHECK("volume_to_dhw_kernel");
    return 0;
}
