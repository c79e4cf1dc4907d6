// Cross-based cost aggregation (CBCA) via per-row and per-column prefix sums (sm_100a).
//
// north_star names this stage; the reference does NOT have it: only a timing-slot label (match.py:98) and the
// parameter name d_cost_volumel_after_aggr (process_functional.py:347) survive, the raw cost volume goes
// straight into SGM (:1166). Parity is therefore UNPINNED: the definition below follows the MC-CNN paper
// (Zbontar & LeCun, JMLR 2016, section 5.1, after Zhang et al. 2009 / Mei et al. 2011) and is checked against
// this repo's own CPU oracle (oracle/stereo.py: cross_arms, cbca_iteration). Default off (cbca_iters = 0).
//
// Definition. Arms: from pixel p walk in a direction; position q at distance k is part of the arm iff it is
// inside the image and (k == 1, or |I(q) - I(p)| < tau and k < L1); the arm stops at the first q that fails.
// len = distance to that first excluded position (1 <= len <= L1). For the volume of image A matched against
// image B (B-pixel of (x, d) is xo = x + dir * d, dir = -1 for the left volume) the support of (y, x, d) is
//   rows  yy in (y - min(upA(y,x), upB(y,xo)),  y + min(downA(y,x), downB(y,xo)))      [exclusive bounds]
//   cols  xx in (x - min(leftA(yy,x), leftB(yy,xo)), x + min(rightA(yy,x), rightB(yy,xo)))  per row yy
// and the aggregated cost is the mean of vol[yy][xx][d] over it. Entries whose B-pixel is outside the image
// are passed through unchanged.
//
// Two HBM-bound passes per iteration (16 B per evaluation): the row pass marches along x with a running fp64
// prefix sum held in a thread-private shared-memory ring (threads along d, coalesced) and writes the row sums
// H[yy][x][d] as fp32; the column pass marches along y over H with running prefixes of H and of the row
// counts and writes mean = (prefix difference) / count. Windows are at most 2*L1 - 1 <= 63 long.
#include "common.cuh"

namespace mccnn {
namespace {

constexpr float kInf = __builtin_huge_valf();
constexpr int CB_THREADS = 128;

__global__ void __launch_bounds__(256) cross_arms_kernel(const uint8_t* __restrict__ img, uchar4* __restrict__ arms, int H,
                                                        int W, int L1, int tau) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int c = img[(size_t)y * W + x];
    int len[4];
    const int dxs[4] = {-1, 1, 0, 0}, dys[4] = {0, 0, -1, 1};
#pragma unroll
    for (int dir = 0; dir < 4; dir++) {
        int k = 1;
        for (;; k++) {
            const int xx = x + dxs[dir] * k, yy = y + dys[dir] * k;
            if (xx < 0 || xx >= W || yy < 0 || yy >= H) break;
            if (k == 1) continue;
            const int q = img[(size_t)yy * W + xx];
            if (abs(q - c) >= tau) break;
            if (k >= L1) break;
        }
        len[dir] = k;
    }
    arms[(size_t)y * W + x] = make_uchar4((unsigned char)len[0], (unsigned char)len[1], (unsigned char)len[2],
                                          (unsigned char)len[3]);
}

// Row pass: H[y][x][d] = sum of vol[y][xx][d] over the column window of (y, x, d).
template <int RING>
__global__ void __launch_bounds__(CB_THREADS) cbca_row_kernel(const float* __restrict__ vol, float* __restrict__ Hs,
                                                             const uchar4* __restrict__ armsA,
                                                             const uchar4* __restrict__ armsB, int W, int D, int Dp,
                                                             int dir, int L1) {
    extern __shared__ double ring_d[];  // [RING][CB_THREADS]
    const int y = blockIdx.y;
    const int d = blockIdx.x * CB_THREADS + threadIdx.x;
    if (d >= Dp) return;
    double* ring = ring_d + threadIdx.x;
    const float* vrow = vol + (size_t)y * W * Dp + d;
    float* hrow = Hs + (size_t)y * W * Dp + d;
    const uchar4* aA = armsA + (size_t)y * W;
    const uchar4* aB = armsB + (size_t)y * W;
    if (d >= D) {  // pad entries stay +INF
        for (int x = 0; x < W; x++) hrow[(size_t)x * Dp] = kInf;
        return;
    }
    double run = 0.0;
    ring[0] = 0.0;  // prefix before element 0 sits in slot 0; prefix including element x in slot (x + 1) % RING
    const int la = L1 - 1;  // an arm reaches at most x + L1 - 1
    // software pipeline: the loads of chunk k+1 (8 cost rows, the arms of the 8 pixels it emits) are in flight
    // while chunk k runs through the dependent prefix / ring updates
    float v[8], vn[8];
    uchar4 ea[8], eb[8], ean[8], ebn[8];
    auto fetch = [&](int s0, float (&vv)[8], uchar4 (&aa)[8], uchar4 (&bb)[8]) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int s = s0 + u;
            vv[u] = (s < W) ? __ldg(vrow + (size_t)s * Dp) : 0.0f;
            const int x = min(max(s - la, 0), W - 1);
            const int xo = min(max(x + dir * d, 0), W - 1);
            aa[u] = aA[x];
            bb[u] = aB[xo];
        }
    };
    fetch(0, vn, ean, ebn);
    for (int s0 = 0; s0 < W + la; s0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) { v[u] = vn[u]; ea[u] = ean[u]; eb[u] = ebn[u]; }
        if (s0 + 8 < W + la) fetch(s0 + 8, vn, ean, ebn);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int s = s0 + u;
            if (s < W) {
                run += (double)v[u];
                ring[((s + 1) & (RING - 1)) * CB_THREADS] = run;
            }
            const int x = s - la;
            if (x >= 0 && x < W) {
                const int xo = x + dir * d;
                float out = 0.0f;  // no B-pixel: never used, the column pass passes the original entry through
                if (xo >= 0 && xo < W) {
                    const int lo = x - min((int)ea[u].x, (int)eb[u].x);      // exclusive
                    const int hi = x + min((int)ea[u].y, (int)eb[u].y) - 1;  // inclusive
                    out = (float)(ring[((hi + 1) & (RING - 1)) * CB_THREADS] - ring[((lo + 1) & (RING - 1)) * CB_THREADS]);
                }
                hrow[(size_t)x * Dp] = out;
            }
        }
    }
}

// Column pass: out[y][x][d] = sum of H[yy][x][d] over the row window / number of summed entries.
template <int RING>
__global__ void __launch_bounds__(CB_THREADS) cbca_col_kernel(const float* __restrict__ Hs, const float* __restrict__ vol,
                                                             float* __restrict__ out, const uchar4* __restrict__ armsA,
                                                             const uchar4* __restrict__ armsB, int H, int W, int D, int Dp,
                                                             int dir, int L1) {
    extern __shared__ double ring_d[];  // [RING][CB_THREADS] sums, then [RING][CB_THREADS] counts (int)
    const int x = blockIdx.y;
    const int d = blockIdx.x * CB_THREADS + threadIdx.x;
    if (d >= Dp) return;
    double* ring = ring_d + threadIdx.x;
    int* ringn = reinterpret_cast<int*>(ring_d + RING * CB_THREADS) + threadIdx.x;
    const size_t rstride = (size_t)W * Dp;
    const float* hcol = Hs + (size_t)x * Dp + d;
    const float* vcol = vol + (size_t)x * Dp + d;
    float* ocol = out + (size_t)x * Dp + d;
    if (d >= D) {
        for (int y = 0; y < H; y++) ocol[(size_t)y * rstride] = kInf;
        return;
    }
    const int xo = x + dir * d;
    if (xo < 0 || xo >= W) {  // no B-pixel: pass through
        for (int y = 0; y < H; y++) ocol[(size_t)y * rstride] = __ldg(vcol + (size_t)y * rstride);
        return;
    }
    const uchar4* aA = armsA + x;
    const uchar4* aB = armsB + xo;
    double run = 0.0;
    int runn = 0;
    ring[0] = 0.0;
    ringn[0] = 0;
    const int la = L1 - 1;
    float v[8], vn[8];
    uchar4 a[8], b[8], an[8], bn[8], ea[8], eb[8], ean[8], ebn[8];
    auto fetch = [&](int s0, float (&vv)[8], uchar4 (&aa)[8], uchar4 (&bb)[8], uchar4 (&eaa)[8], uchar4 (&ebb)[8]) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int s = s0 + u;
            const int sc = min(s, H - 1);
            vv[u] = (s < H) ? __ldg(hcol + (size_t)s * rstride) : 0.0f;
            aa[u] = aA[(size_t)sc * W];
            bb[u] = aB[(size_t)sc * W];
            const int y = min(max(s - la, 0), H - 1);
            eaa[u] = aA[(size_t)y * W];
            ebb[u] = aB[(size_t)y * W];
        }
    };
    fetch(0, vn, an, bn, ean, ebn);
    for (int s0 = 0; s0 < H + la; s0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) { v[u] = vn[u]; a[u] = an[u]; b[u] = bn[u]; ea[u] = ean[u]; eb[u] = ebn[u]; }
        if (s0 + 8 < H + la) fetch(s0 + 8, vn, an, bn, ean, ebn);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int s = s0 + u;
            if (s < H) {
                run += (double)v[u];
                runn += min((int)a[u].x, (int)b[u].x) + min((int)a[u].y, (int)b[u].y) - 1;
                ring[((s + 1) & (RING - 1)) * CB_THREADS] = run;
                ringn[((s + 1) & (RING - 1)) * CB_THREADS] = runn;
            }
            const int y = s - la;
            if (y >= 0 && y < H) {
                const int lo = y - min((int)ea[u].z, (int)eb[u].z);      // exclusive
                const int hi = y + min((int)ea[u].w, (int)eb[u].w) - 1;  // inclusive
                const int ih = ((hi + 1) & (RING - 1)) * CB_THREADS, il = ((lo + 1) & (RING - 1)) * CB_THREADS;
                const double sum = ring[ih] - ring[il];
                const int cnt = ringn[ih] - ringn[il];
                ocol[(size_t)y * rstride] = (float)(sum / (double)cnt);
            }
        }
    }
}

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" int mccnn_cross_arms(const uint8_t* image, uint8_t* arms4, int H, int W, int L1, int tau, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(image && arms4, MCCNN_EINVAL, "mccnn_cross_arms: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && H <= 65535, MCCNN_EINVAL, "mccnn_cross_arms: bad shape %dx%d", W, H);
    MCCNN_REQUIRE(L1 >= 1 && L1 <= 32 && tau >= 0, MCCNN_EINVAL, "mccnn_cross_arms: L1=%d outside 1..32 or tau=%d < 0", L1, tau);
    dim3 grid(ceil_div(W, 256), H);
    cross_arms_kernel<<<grid, 256, 0, stream>>>(image, reinterpret_cast<uchar4*>(arms4), H, W, L1, tau);
    MCCNN_LAUNCH_CHECK("cross_arms_kernel");
    return 0;
}

extern "C" int mccnn_cbca(const float* vol_in, float* vol_out, float* tmp, const uint8_t* arms_self, const uint8_t* arms_other,
                          int H, int W, int D, int direction, int L1, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(vol_in && vol_out && tmp && arms_self && arms_other, MCCNN_EINVAL, "mccnn_cbca: null argument");
    MCCNN_REQUIRE(vol_in != vol_out && tmp != vol_in && tmp != vol_out, MCCNN_EINVAL, "mccnn_cbca: the three volumes must be distinct");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096 && H <= 65535 && W <= 65535, MCCNN_EINVAL,
                  "mccnn_cbca: bad shape H=%d W=%d D=%d", H, W, D);
    MCCNN_REQUIRE(direction == -1 || direction == 1, MCCNN_EINVAL, "mccnn_cbca: direction must be -1 (left volume) or +1 (right)");
    MCCNN_REQUIRE(L1 >= 1 && L1 <= 32, MCCNN_EINVAL, "mccnn_cbca: L1=%d outside 1..32", L1);
    const int Dp = disp_pitch(D);
    const uchar4* aA = reinterpret_cast<const uchar4*>(arms_self);
    const uchar4* aB = reinterpret_cast<const uchar4*>(arms_other);
    const int nd = ceil_div(Dp, CB_THREADS);
    if (L1 <= 16) {
        const size_t smr = (size_t)32 * CB_THREADS * sizeof(double), smc = smr + (size_t)32 * CB_THREADS * sizeof(int);
        MCCNN_CUDA(cudaFuncSetAttribute(cbca_col_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
        cbca_row_kernel<32><<<dim3(nd, H), CB_THREADS, smr, stream>>>(vol_in, tmp, aA, aB, W, D, Dp, direction, L1);
        MCCNN_LAUNCH_CHECK("cbca_row_kernel");
        cbca_col_kernel<32><<<dim3(nd, W), CB_THREADS, smc, stream>>>(tmp, vol_in, vol_out, aA, aB, H, W, D, Dp, direction, L1);
        MCCNN_LAUNCH_CHECK("cbca_col_kernel");
    } else {
        const size_t smr = (size_t)64 * CB_THREADS * sizeof(double), smc = smr + (size_t)64 * CB_THREADS * sizeof(int);
        MCCNN_CUDA(cudaFuncSetAttribute(cbca_row_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smr));
        MCCNN_CUDA(cudaFuncSetAttribute(cbca_col_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
        cbca_row_kernel<64><<<dim3(nd, H), CB_THREADS, smr, stream>>>(vol_in, tmp, aA, aB, W, D, Dp, direction, L1);
        MCCNN_LAUNCH_CHECK("cbca_row_kernel");
        cbca_col_kernel<64><<<dim3(nd, W), CB_THREADS, smc, stream>>>(tmp, vol_in, vol_out, aA, aB, H, W, D, Dp, direction, L1);
        MCCNN_LAUNCH_CHECK("cbca_col_kernel");
    }
    return 0;
}
