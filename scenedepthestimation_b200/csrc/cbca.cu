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

// Both passes are bound by instruction issue (ncu: ~95 instructions per evaluation in the first version, 16 warps per SM
// allowed by the fp64 rings), so the per-step work is kept lean: a flat thread index over (row, d) (no idle lanes, no
// per-thread 64-bit index products), the arms of a pixel as two u16x2 words (one PRMT each from the uchar4) so that one
// VIMNMX.U16x2 takes both minima of an intersection, ring slots addressed by a masked position, and in the column
// pass the vertical arms of a row stored with that row's count prefix (no second pass over the arms).
constexpr int CB_PF = 8;  // steps whose loads are in flight while the previous chunk runs through the prefix chain

__device__ __forceinline__ unsigned arms_lr(unsigned w) { return __byte_perm(w, 0u, 0x4140); }  // left | right << 16
__device__ __forceinline__ unsigned arms_ud(unsigned w) { return __byte_perm(w, 0u, 0x4342); }  // up | down << 16

// Row pass: H[y][x][d] = sum of vol[y][xx][d] over the column window of (y, x, d).
// Ring: slot p % RING holds the fp64 prefix of elements < p, as [RING][CB_THREADS] doubles.
template <int RING, int DIR>
__global__ void __launch_bounds__(CB_THREADS) cbca_row_kernel(const float* __restrict__ vol, float* __restrict__ Hs,
                                                             const unsigned* __restrict__ armsA,
                                                             const unsigned* __restrict__ armsB, int H, int W, int D, int Dp,
                                                             int L1) {
    extern __shared__ double ring_d[];
    const long long gid = (long long)blockIdx.x * CB_THREADS + threadIdx.x;
    if (gid >= (long long)H * Dp) return;
    const int y = (int)(gid / Dp), d = (int)(gid - (long long)y * Dp);
    double* ring = ring_d + threadIdx.x;
    const float* vrow = vol + (size_t)y * W * Dp + d;
    float* hrow = Hs + (size_t)y * W * Dp + d;
    if (d >= D) {  // pad entries stay +INF
        for (int x = 0; x < W; x++) hrow[(size_t)x * Dp] = kInf;
        return;
    }
    const unsigned* aA = armsA + (size_t)y * W;
    const unsigned* aB = armsB + (size_t)y * W + DIR * d;  // B pixel of x: aB[x], valid for 0 <= x + DIR * d < W
    const int xo_lo = DIR < 0 ? d : 0, xo_hi = DIR < 0 ? W : W - d;  // x range with a B pixel
    const int la = L1 - 1;  // an arm reaches at most x + L1 - 1
    double run = 0.0;
    ring[0] = 0.0;
    // Addresses: one pointer per stream, advanced once per chunk of CB_PF steps; inside a chunk the offsets u * Dp are loop
    // invariants (a 64-bit index product per access was a third of all instructions). Chunks that lie entirely inside
    // [la, W) take a path without range tests on s and x.
    float v[CB_PF], vn[CB_PF];
    unsigned ea[CB_PF], ean[CB_PF], eb[CB_PF], ebn[CB_PF];
    const unsigned nb = (unsigned)(xo_hi - xo_lo);  // x has a B pixel iff (unsigned)(x - xo_lo) < nb
    auto fetch = [&](int s0, const float* vp, float (&vv)[CB_PF], unsigned (&aa)[CB_PF], unsigned (&bb)[CB_PF]) {
        if (s0 >= la && s0 + CB_PF <= W) {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                vv[u] = __ldg(vp + u * Dp);
                const int x = s0 - la + u;
                const bool has_b = (unsigned)(x - xo_lo) < nb;
                aa[u] = has_b ? __ldg(aA + x) : 0u;
                bb[u] = has_b ? __ldg(aB + x) : 0u;
            }
        } else {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                const int s = s0 + u;
                vv[u] = (s < W) ? __ldg(vp + u * Dp) : 0.0f;
                const int x = s - la;
                const bool has_b = x >= 0 && x < W && (unsigned)(x - xo_lo) < nb;
                aa[u] = has_b ? __ldg(aA + x) : 0u;
                bb[u] = has_b ? __ldg(aB + x) : 0u;
            }
        }
    };
    const float* vp = vrow;          // element s0 of the row
    float* hp = hrow - (size_t)la * Dp;  // element s0 - la (not dereferenced before s0 - la + u >= 0)
    const size_t chunk = (size_t)CB_PF * Dp;
    fetch(0, vp, vn, ean, ebn);
    for (int s0 = 0; s0 < W + la; s0 += CB_PF, vp += chunk, hp += chunk) {
#pragma unroll
        for (int u = 0; u < CB_PF; u++) { v[u] = vn[u]; ea[u] = ean[u]; eb[u] = ebn[u]; }
        if (s0 + CB_PF < W + la) fetch(s0 + CB_PF, vp + chunk, vn, ean, ebn);
        auto emit = [&](int u, int x) {
            float out = 0.0f;  // no B pixel: never used, the column pass passes the original entry through
            if ((unsigned)(x - xo_lo) < nb) {
                const unsigned m = __vminu2(arms_lr(ea[u]), arms_lr(eb[u]));
                const int lo1 = x + 1 - (int)(m & 0xffffu);  // first summed element
                const int hi1 = x + (int)(m >> 16);         // one past the last
                out = (float)(ring[(hi1 & (RING - 1)) * CB_THREADS] - ring[(lo1 & (RING - 1)) * CB_THREADS]);
            }
            hp[u * Dp] = out;
        };
        if (s0 >= la && s0 + CB_PF <= W) {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                run += (double)v[u];
                ring[((s0 + u + 1) & (RING - 1)) * CB_THREADS] = run;
                emit(u, s0 - la + u);
            }
        } else {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                const int s = s0 + u;
                if (s < W) {
                    run += (double)v[u];
                    ring[((s + 1) & (RING - 1)) * CB_THREADS] = run;
                }
                const int x = s - la;
                if (x >= 0 && x < W) emit(u, x);
            }
        }
    }
}

// Column pass: out[y][x][d] = sum of H[yy][x][d] over the row window / number of summed entries.
// Rings: [RING][CB_THREADS] fp64 prefix sums; [RING][CB_THREADS] words = prefix of the row counts << 10 (22 bits: H * 63 < 2^22)
// | (up - 1) << 5 | (down - 1) of the row the prefix ends with.
template <int RING, int DIR>
__global__ void __launch_bounds__(CB_THREADS) cbca_col_kernel(const float* __restrict__ Hs, const float* __restrict__ vol,
                                                             float* __restrict__ out, const unsigned* __restrict__ armsA,
                                                             const unsigned* __restrict__ armsB, int H, int W, int D, int Dp,
                                                             int L1) {
    extern __shared__ double ring_d[];
    const long long gid = (long long)blockIdx.x * CB_THREADS + threadIdx.x;
    if (gid >= (long long)W * Dp) return;
    const int x = (int)(gid / Dp), d = (int)(gid - (long long)x * Dp);
    double* ring = ring_d + threadIdx.x;
    unsigned* ringw = reinterpret_cast<unsigned*>(ring_d + RING * CB_THREADS) + threadIdx.x;
    const size_t rstride = (size_t)W * Dp;
    const float* hcol = Hs + (size_t)x * Dp + d;
    const float* vcol = vol + (size_t)x * Dp + d;
    float* ocol = out + (size_t)x * Dp + d;
    if (d >= D) {
        for (int yy = 0; yy < H; yy++) ocol[(size_t)yy * rstride] = kInf;
        return;
    }
    const int xo = x + DIR * d;
    if (xo < 0 || xo >= W) {  // no B-pixel: pass through
        for (int yy = 0; yy < H; yy++) ocol[(size_t)yy * rstride] = __ldg(vcol + (size_t)yy * rstride);
        return;
    }
    const unsigned* aA = armsA + x;
    const unsigned* aB = armsB + xo;
    double run = 0.0;
    unsigned runn = 0;
    ring[0] = 0.0;
    ringw[0] = 0u;
    const int la = L1 - 1;
    float v[CB_PF], vn[CB_PF];
    unsigned a[CB_PF], an[CB_PF], b[CB_PF], bn[CB_PF];
    auto fetch = [&](int s0, const float* hq, const unsigned* pa, const unsigned* pb, float (&vv)[CB_PF], unsigned (&aa)[CB_PF],
                     unsigned (&bb)[CB_PF]) {
        if (s0 + CB_PF <= H) {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                vv[u] = __ldg(hq + (size_t)u * rstride);
                aa[u] = __ldg(pa + u * W);
                bb[u] = __ldg(pb + u * W);
            }
        } else {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                const bool in = s0 + u < H;
                vv[u] = in ? __ldg(hq + (size_t)u * rstride) : 0.0f;
                aa[u] = in ? __ldg(pa + u * W) : 0x01010101u;
                bb[u] = in ? __ldg(pb + u * W) : 0x01010101u;
            }
        }
    };
    // one pointer per stream, advanced once per chunk (see the row pass)
    const float* hq = hcol;                       // row s0
    const unsigned *pa = aA, *pb = aB;            // row s0
    float* oq = ocol - (size_t)la * rstride;      // row s0 - la (not dereferenced before that row exists)
    const size_t chunk = (size_t)CB_PF * rstride, achunk = (size_t)CB_PF * W;
    fetch(0, hq, pa, pb, vn, an, bn);
    for (int s0 = 0; s0 < H + la; s0 += CB_PF, hq += chunk, pa += achunk, pb += achunk, oq += chunk) {
#pragma unroll
        for (int u = 0; u < CB_PF; u++) { v[u] = vn[u]; a[u] = an[u]; b[u] = bn[u]; }
        if (s0 + CB_PF < H + la) fetch(s0 + CB_PF, hq + chunk, pa + achunk, pb + achunk, vn, an, bn);
        auto accumulate = [&](int u, int s) {
            run += (double)v[u];
            const unsigned lr = __vminu2(arms_lr(a[u]), arms_lr(b[u]));
            const unsigned ud = __vminu2(arms_ud(a[u]), arms_ud(b[u]));
            runn += (lr & 0xffffu) + (lr >> 16) - 1u;  // entries of the row window of (s, x, d)
            const int sl = ((s + 1) & (RING - 1)) * CB_THREADS;
            ring[sl] = run;
            // the arms of row s travel with its prefix: (up - 1) << 5 | (down - 1), both fields from one subtraction
            const unsigned udm = ud - 0x00010001u;
            ringw[sl] = (runn << 10) | ((udm & 0xffffu) << 5) | (udm >> 16);
        };
        auto emit = [&](int u, int y) {
            const unsigned e = ringw[((y + 1) & (RING - 1)) * CB_THREADS];
            const int up = (int)((e >> 5) & 31u), dn = (int)(e & 31u);  // lengths - 1
            const int ih = ((y + dn + 1) & (RING - 1)) * CB_THREADS, il = ((y - up) & (RING - 1)) * CB_THREADS;
            const double sum = ring[ih] - ring[il];
            const unsigned cnt = (ringw[ih] >> 10) - (ringw[il] >> 10);
            oq[(size_t)u * rstride] = (float)sum / (float)cnt;
        };
        if (s0 >= la && s0 + CB_PF <= H) {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                accumulate(u, s0 + u);
                emit(u, s0 - la + u);
            }
        } else {
#pragma unroll
            for (int u = 0; u < CB_PF; u++) {
                const int s = s0 + u;
                if (s < H) accumulate(u, s);
                const int y = s - la;
                if (y >= 0 && y < H) emit(u, y);
            }
        }
    }
}

template <int RING, int DIR>
int launch_cbca(const float* vol_in, float* vol_out, float* tmp, const unsigned* aA, const unsigned* aB, int H, int W, int D,
                int Dp, int L1, cudaStream_t stream) {
    const size_t smr = (size_t)RING * CB_THREADS * sizeof(double);
    const size_t smc = smr + (size_t)RING * CB_THREADS * sizeof(unsigned);
    if (int e = kernel_setup<cbca_row_kernel<RING, DIR>>(CB_THREADS, smr, nullptr)) return e;
    if (int e = kernel_setup<cbca_col_kernel<RING, DIR>>(CB_THREADS, smc, nullptr)) return e;
    const long long nrow = (long long)H * Dp, ncol = (long long)W * Dp;
    cbca_row_kernel<RING, DIR><<<(unsigned)((nrow + CB_THREADS - 1) / CB_THREADS), CB_THREADS, smr, stream>>>(vol_in, tmp, aA, aB, H, W, D,
                                                                                                         Dp, L1);
    MCCNN_LAUNCH_CHECK("cbca_row_kernel");
    cbca_col_kernel<RING, DIR><<<(unsigned)((ncol + CB_THREADS - 1) / CB_THREADS), CB_THREADS, smc, stream>>>(tmp, vol_in, vol_out, aA, aB,
                                                                                                         H, W, D, Dp, L1);
    MCCNN_LAUNCH_CHECK("cbca_col_kernel");
    return 0;
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
    const unsigned* aA = reinterpret_cast<const unsigned*>(arms_self);  // uchar4 {left, right, up, down} as one word
    const unsigned* aB = reinterpret_cast<const unsigned*>(arms_other);
    // ring slots: a window reaches L1 - 1 elements ahead of and L1 behind the emission position
    if (L1 <= 16)
        return direction < 0 ? launch_cbca<32, -1>(vol_in, vol_out, tmp, aA, aB, H, W, D, Dp, L1, stream)
                             : launch_cbca<32, 1>(vol_in, vol_out, tmp, aA, aB, H, W, D, Dp, L1, stream);
    return direction < 0 ? launch_cbca<64, -1>(vol_in, vol_out, tmp, aA, aB, H, W, D, Dp, L1, stream)
                         : launch_cbca<64, 1>(vol_in, vol_out, tmp, aA, aB, H, W, D, Dp, L1, stream);
}
