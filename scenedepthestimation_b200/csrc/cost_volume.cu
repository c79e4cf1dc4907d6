// Normalised dot-product cost volume over the disparity range (sm_100a).
//
// Replaces compute_cost_volume_kernel (process_functional.py:120-131) and the host-side
// np.ones fill of both volumes (:1111-1114).
//
// Arithmetic contract: temp is fp64; every product fl*fr is an fp32 multiply
// that is widened and added in fp64, i = 0..63 in order; the store rounds -temp to fp32.
// CL[y][x][d] and CR[y][x-d][d] receive the same value; entries never written keep `fill`.
//
// Shape: a CTA owns (row y, TX left pixels, DB disparities). The 64-float feature rows of the TX
// left pixels and of the TX+DB-1 right pixels they can meet are staged in shared memory (right rows
// padded to 65 floats so that lanes walking x-d hit distinct banks); lanes run along d so the CL
// stores are 128-byte coalesced; the CR values of the tile are transposed through shared memory and
// written as runs along d as well. The x grid extends to W+D-1 so that the tiles past the right image
// edge write the `fill` entries of CR: every element of both volumes is written exactly once. The pad
// entries [D, Dp) of both volumes are written as +INF (the SGM kernels use them as "never the minimum").
#include "common.cuh"

namespace mccnn {
namespace {

constexpr int TX = 32;
constexpr int DB = 128;
constexpr int NF = MCCNN_FEATURES;
constexpr int FR_PITCH = NF + 1;
constexpr int WIN = TX + DB - 1;
constexpr int CV_THREADS = 256;

struct CvSmem {
    float fl[TX][NF];
    float fr[WIN][FR_PITCH];
    float res[TX][DB + 1];
};

__global__ void __launch_bounds__(CV_THREADS) cost_volume_exact_kernel(const float* __restrict__ fl,
                                                                      const float* __restrict__ fr,
                                                                      float* __restrict__ CL, float* __restrict__ CR,
                                                                      int H, int W, int D, int Dp, float fill) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CvSmem& sm = *reinterpret_cast<CvSmem*>(smem_raw);
    const int x0 = blockIdx.x * TX;
    const int y = blockIdx.y;
    const int dblk = blockIdx.z * DB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int xr0 = x0 - dblk - (DB - 1);  // first right pixel of the window
    const bool any_left = x0 < W;

    if (any_left) {
        // stage features (coalesced: 64 consecutive floats per pixel)
        for (int i = tid; i < TX * NF; i += CV_THREADS) {
            const int p = i / NF, f = i % NF;
            const int x = x0 + p;
            sm.fl[p][f] = (x < W) ? fl[((size_t)y * W + x) * NF + f] : 0.0f;
        }
        for (int i = tid; i < WIN * NF; i += CV_THREADS) {
            const int p = i / NF, f = i % NF;
            const int x = xr0 + p;
            sm.fr[p][f] = (x >= 0 && x < W) ? fr[((size_t)y * W + x) * NF + f] : 0.0f;
        }
    }
    __syncthreads();

    // warp w owns left pixels x0 + 4w + a (a = 0..3); lane owns d = dblk + lane + 32k (k = 0..3)
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[a][k] = 0.0;
    if (any_left) {
        const int base = 4 * warp - lane + (DB - 1);  // window index of (a = 0, k = 0)
#pragma unroll 4
        for (int i = 0; i < NF; i++) {
            float fa[4];
#pragma unroll
            for (int a = 0; a < 4; a++) fa[a] = sm.fl[4 * warp + a][i];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float fb = sm.fr[base + a - 32 * k][i];
                    acc[a][k] += (double)__fmul_rn(fa[a], fb);
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int xl = 4 * warp + a;
        const int x = x0 + xl;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int dl = lane + 32 * k;
            const int d = dblk + dl;
            const bool valid = (x < W) && (x - d >= 0);
            const float v = valid ? (float)(-acc[a][k]) : fill;
            sm.res[xl][dl] = v;
            if (x < W && d < Dp) CL[((size_t)y * W + x) * Dp + d] = (d < D) ? v : __int_as_float(0x7f800000);
        }
    }
    if (CR == nullptr) return;
    __syncthreads();
    // CR[y][xr][d] = value of left pixel x = xr + d: for a fixed xr the tile holds a run of <= TX disparities
    for (int p = warp; p < WIN; p += CV_THREADS / 32) {
        const int xr = xr0 + p;
        if (xr < 0 || xr >= W) continue;
        const int x = x0 + lane;  // lane runs along the left pixels of the tile == along d
        const int d = x - xr;
        if (d >= dblk && d < dblk + DB && d < D) CR[((size_t)y * W + xr) * Dp + d] = sm.res[lane][d - dblk];
    }
    // pad entries [D, Dp) of CR: +INF (the SGM kernels rely on it)
    if (Dp > D && dblk + DB >= D && dblk < D) {
        for (int i = tid; i < TX * (Dp - D); i += CV_THREADS) {
            const int x = x0 + i / (Dp - D), d = D + i % (Dp - D);
            if (x < W) CR[((size_t)y * W + x) * Dp + d] = __int_as_float(0x7f800000);
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

extern "C" int mccnn_cost_volume(const float* fl, const float* fr, float* CL, float* CR, int H, int W, int D, float fill,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(fl && fr && CL, MCCNN_EINVAL, "mccnn_cost_volume: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096, MCCNN_EINVAL, "mccnn_cost_volume: bad shape H=%d W=%d D=%d", H,
                  W, D);
    MCCNN_REQUIRE(H <= 65535, MCCNN_EINVAL, "mccnn_cost_volume: H=%d exceeds 65535", H);
    const int Dp = disp_pitch(D);
    const size_t smem = sizeof(CvSmem);
    MCCNN_CUDA(cudaFuncSetAttribute(cost_volume_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int xt = ceil_div(CR ? (W + D - 1) : W, TX);
    dim3 grid(xt, H, ceil_div(D, DB));
    cost_volume_exact_kernel<<<grid, CV_THREADS, smem, stream>>>(fl, fr, CL, CR, H, W, D, Dp, fill);
    MCCNN_LAUNCH_CHECK("cost_volume_exact_kernel");
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
