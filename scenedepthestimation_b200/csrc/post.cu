// Winner-takes-all, left-right consistency check, occlusion fill, 5x5 median, 9x9 bilateral filter,
// u8 encode and the bad-pixel metric (sm_100a). All are byte/row-bound per-pixel kernels.
//
// Replaces, with the reference's exact arithmetic:
//   WTA_and_SupixelRefinement_kernel  process_functional.py:800-837   (and CPU WTA1, :96-113)
//   is_error_match_kernel             process_functional.py:977-1000
//   LRC_kernel                        process_functional.py:1003-1088
//   Median_Filter_kernel              process_functional.py:840-879, launched at :1250
//   Bilateral_Filter_kernel           process_functional.py:882-974  (launch commented out at :1260)
//   astype('uint8')[*2]               match_single.py:55, match.py:90
//   error_calculate.py:68-83
#include "common.cuh"

namespace mccnn {
namespace {

// ---- WTA over [H][W][Dp]: one warp per pixel, lanes along d (coalesced)
__global__ void __launch_bounds__(256) wta_kernel(const float* __restrict__ S, float* __restrict__ disp, int npix, int D,
                                                 int Dp) {
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const float* row = S + (size_t)pix * Dp;
    float best = __int_as_float(0x7f800000);
    int bd = 0x7fffffff;
    for (int d = lane; d < D; d += 32) {
        const float v = row[d] + 0.0f;
        if (v < best) { best = v; bd = d; }
    }
    int k = __float_as_int(best);
    k ^= (k >> 31) & 0x7fffffff;
    const int mk = __reduce_min_sync(0xffffffffu, k);
    const int md = __reduce_min_sync(0xffffffffu, k == mk ? bd : 0x7fffffff);
    if (lane == 0) disp[pix] = (float)md;
}

// ---- WTA + parabola refinement (:813-819), same typing as sgm.cu's subpixel_refine
__global__ void __launch_bounds__(256) wta_subpixel_kernel(const float* __restrict__ S, float* __restrict__ disp, int npix,
                                                          int D, int Dp) {
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const float* row = S + (size_t)pix * Dp;
    float best = __int_as_float(0x7f800000);
    int bd = 0x7fffffff;
    for (int d = lane; d < D; d += 32) {
        const float v = row[d] + 0.0f;
        if (v < best) { best = v; bd = d; }
    }
    int k = __float_as_int(best);
    k ^= (k >> 31) & 0x7fffffff;
    const int mk = __reduce_min_sync(0xffffffffu, k);
    const int md = __reduce_min_sync(0xffffffffu, k == mk ? bd : 0x7fffffff);
    if (lane == 0) {
        float out = (float)md;
        if (md > 0 && md < D - 1) {
            const float cm = row[md - 1], c = row[md], cp = row[md + 1];
            const float num = cp - cm;
            const double den = 2.0 * ((double)(cm + cp) - 2.0 * (double)c);
            if (den > 0.0) out = (float)((double)md - (double)num / den);
        }
        disp[pix] = out;
    }
}

// ---- WTA over dense [D][H][W] (CPU WTA1): thread per pixel, coalesced along x
__global__ void wta_dhw_kernel(const float* __restrict__ vol, float* __restrict__ disp, int npix, int D) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    float best = __int_as_float(0x7f800000);
    int bd = -1;
    for (int d = 0; d < D; d++) {
        const float v = vol[(size_t)d * npix + p];
        if (v < best) { best = v; bd = d; }
    }
    disp[p] = (float)bd;
}

// ---- left-right consistency flags (:977-1000). The reference indexes with uint8(x - ld); run on a
// B200 the index is NOT truncated (tests/golden/ref_wide_9x300.npz), so neither is it here.
__global__ void lr_flags_kernel(const float* __restrict__ dl, const float* __restrict__ dr, unsigned char* __restrict__ fL,
                                unsigned char* __restrict__ fR, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const size_t o = (size_t)y * W;
    const float ld = dl[o + x];
    const double rd = (double)x - (double)ld;
    unsigned char f = 0;
    if (rd >= 0) {
        const int idx = (int)rd;
        const float minus = ld - dr[o + idx];
        f = (minus > 1.0f || minus < -1.0f) ? 1 : 0;
    }
    fL[o + x] = f;
    if (fR) {
        const float rdv = dr[o + x];
        const double ldx = (double)x + (double)rdv;
        unsigned char g = 0;
        if (ldx < W) {
            const int idx = (int)ldx;
            const float minus = rdv - dl[o + idx];
            g = (minus > 1.0f || minus < -1.0f) ? 1 : 0;
        }
        fR[o + x] = g;
    }
}

// ---- occlusion fill (:1003-1088): mean of the nearest unflagged raw disparity up, down, right, left.
// The reference walks four data-dependent while-loops per flagged pixel; here the four "nearest unflagged
// value" maps are running carries: two column sweeps (thread per column, coalesced along x) and two row
// sweeps (warp per row, 32-pixel chunks, ballot + shuffle), O(H*W) work whatever the flag pattern.
// Disparities are >= 0, so -1 marks "none found".
__global__ void __launch_bounds__(128) lrc_vertical_kernel(const float* __restrict__ dl, const unsigned char* __restrict__ fl,
                                                          float* __restrict__ up, float* __restrict__ down, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    float carry = -1.0f;
#pragma unroll 8
    for (int y = 0; y < H; y++) {
        const size_t o = (size_t)y * W + x;
        const unsigned char f = fl[o];
        const float v = dl[o];
        if (f == 1) up[o] = carry; else carry = v;
    }
    carry = -1.0f;
#pragma unroll 8
    for (int y = H - 1; y >= 0; y--) {
        const size_t o = (size_t)y * W + x;
        const unsigned char f = fl[o];
        const float v = dl[o];
        if (f == 1) down[o] = carry; else carry = v;
    }
}

__global__ void __launch_bounds__(256) lrc_horizontal_kernel(const float* __restrict__ dl, const unsigned char* __restrict__ fl,
                                                            const float* __restrict__ up, const float* __restrict__ down,
                                                            float* __restrict__ out, int H, int W) {
    const int lane = threadIdx.x & 31;
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (y >= H) return;
    const size_t ro = (size_t)y * W;
    const int nchunks = (W + 31) / 32;
    // sweep 1, right to left: nearest unflagged value to the right, parked in `out` at flagged pixels
    float carry = -1.0f;
    for (int c = nchunks - 1; c >= 0; c--) {
        const int x = c * 32 + lane;
        const bool inb = x < W;
        const unsigned char f = inb ? fl[ro + x] : 1;
        const float v = inb ? dl[ro + x] : 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, inb && f != 1);
        const unsigned right = m & (0xfffffffeu << lane);
        const float rv = __shfl_sync(0xffffffffu, v, right ? __ffs(right) - 1 : 0);
        if (inb && f == 1) out[ro + x] = right ? rv : carry;
        if (m) carry = __shfl_sync(0xffffffffu, v, __ffs(m) - 1);
    }
    // sweep 2, left to right: nearest unflagged value to the left, then combine in the reference's order
    carry = -1.0f;
    for (int c = 0; c < nchunks; c++) {
        const int x = c * 32 + lane;
        const bool inb = x < W;
        const unsigned char f = inb ? fl[ro + x] : 1;
        const float v = inb ? dl[ro + x] : 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, inb && f != 1);
        const unsigned left = m & ((1u << lane) - 1u);
        const float lv = __shfl_sync(0xffffffffu, v, left ? 31 - __clz(left) : 0);
        if (inb) {
            float r = v;
            if (f == 1) {
                const float cand[4] = {up[ro + x], down[ro + x], out[ro + x], left ? lv : carry};
                int number = 0;
                double sum_d = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (cand[i] >= 0.0f) { number++; sum_d += (double)cand[i]; }
                if (number > 0) r = (float)(sum_d / (double)number);
            }
            out[ro + x] = r;
        }
        if (m) carry = __shfl_sync(0xffffffffu, v, 31 - __clz(m));
    }
}

// ---- 5x5 median (:840-879): the 13th smallest of 25; the 2-pixel border keeps the raw WTA map
__global__ void median5_kernel(const float* __restrict__ filled, const float* __restrict__ wta, float* __restrict__ out,
                               int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t o = (size_t)y * W + x;
    if (x < 2 || y < 2 || x + 2 >= W || y + 2 >= H) {
        if (out != wta) out[o] = wta[o];
        return;
    }
    float w[25];
#pragma unroll
    for (int i = -2; i <= 2; i++)
#pragma unroll
        for (int j = -2; j <= 2; j++) w[(i + 2) * 5 + j + 2] = filled[(size_t)(y + i) * W + x + j];
    // partial selection: after pass i, w[i] holds the (i+1)-th smallest (fully unrolled, registers only)
    float cur = 0.0f;
#pragma unroll
    for (int i = 0; i < 13; i++) {
        cur = w[i];
#pragma unroll
        for (int j = i + 1; j < 25; j++) {
            const float lo = fminf(cur, w[j]);
            w[j] = fmaxf(cur, w[j]);
            cur = lo;
        }
    }
    out[o] = cur;
}

// ---- 9x9 range-weighted mean (:882-974); uint8 differences wrap to uint64 (App. A7)
__constant__ float kBilateralW[10] = {0.167747f, 0.165145f, 0.157581f, 0.145735f, 0.130632f,
                                      0.113490f, 0.095563f, 0.077991f, 0.061692f, 0.047297f};

__global__ void bilateral9_kernel(const unsigned char* __restrict__ img, const float* __restrict__ disp,
                                  float* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int ic = img[(size_t)y * W + x];
    double wsum = 0.0, dsum = 0.0;
    for (int i = -4; i <= 4; i++)
        for (int j = -4; j <= 4; j++) {
            const int yy = y + i, xx = x + j;
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            const int in_i = in ? (int)img[(size_t)yy * W + xx] : 0;
            const float in_d = in ? disp[(size_t)yy * W + xx] : 0.0f;
            const int minus = ic - in_i;  // unsigned wrap: negative differences are huge, never < 5
            if (minus >= 0 && minus < 5) {
                const float wgt = kBilateralW[minus];
                wsum += (double)wgt;
                dsum += (double)__fmul_rn(wgt, in_d);
            }
        }
    out[(size_t)y * W + x] = (float)(dsum / wsum);
}

__global__ void encode_u8_kernel(const float* __restrict__ disp, unsigned char* __restrict__ out, size_t n, int scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned v = (unsigned)(long long)disp[i];  // astype('uint8'): truncate toward zero, modulo 256
    out[i] = (unsigned char)((v & 255u) * (unsigned)scale);
}

__global__ void encode_u16_kernel(const float* __restrict__ disp, unsigned short* __restrict__ out, size_t n, float scale) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = disp[i] * scale;
    out[i] = (unsigned short)(v <= 0.f ? 0.f : (v >= 65535.f ? 65535.f : v));
}

template <typename T>
__global__ void bad_pixels_kernel(const T* __restrict__ disp, const float* __restrict__ gt,
                                  unsigned long long* __restrict__ counts, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bad = 0, valid = 0;
    if (i < n) {
        const float t = gt[i];
        if (!(t == __int_as_float(0x7f800000) || t == 0.0f)) {
            valid = 1;
            bad = fabsf((float)disp[i] - t) > 1.0f ? 1 : 0;
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, bad), v = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0) {
        if (b) atomicAdd(&counts[0], (unsigned long long)__popc(b));
        if (v) atomicAdd(&counts[1], (unsigned long long)__popc(v));
    }
}

inline int check_hw(const char* who, int H, int W) {
    MCCNN_REQUIRE(H >= 1 && W >= 1 && H <= 65535, MCCNN_EINVAL, "%s: bad shape H=%d W=%d", who, H, W);
    return 0;
}

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" int mccnn_wta(const float* S, float* disp, int H, int W, int D, void* stream) {
    MCCNN_REQUIRE(S && disp, MCCNN_EINVAL, "mccnn_wta: null argument");
    if (int e = check_hw("mccnn_wta", H, W)) return e;
    MCCNN_REQUIRE(D >= 1, MCCNN_EINVAL, "mccnn_wta: D=%d", D);
    const int npix = H * W;
    wta_kernel<<<ceil_div(npix, 8), 256, 0, (cudaStream_t)stream>>>(S, disp, npix, D, disp_pitch(D));
    MCCNN_LAUNCH_CHECK("wta_kernel");
    return 0;
}

extern "C" int mccnn_wta_dhw(const float* vol, float* disp, int H, int W, int D, void* stream) {
    MCCNN_REQUIRE(vol && disp, MCCNN_EINVAL, "mccnn_wta_dhw: null argument");
    if (int e = check_hw("mccnn_wta_dhw", H, W)) return e;
    MCCNN_REQUIRE(D >= 1, MCCNN_EINVAL, "mccnn_wta_dhw: D=%d", D);
    const int npix = H * W;
    wta_dhw_kernel<<<ceil_div(npix, 256), 256, 0, (cudaStream_t)stream>>>(vol, disp, npix, D);
    MCCNN_LAUNCH_CHECK("wta_dhw_kernel");
    return 0;
}

extern "C" int mccnn_lr_flags(const float* dispL, const float* dispR, uint8_t* flagL, uint8_t* flagR, int H, int W,
                              void* stream) {
    MCCNN_REQUIRE(dispL && dispR && flagL, MCCNN_EINVAL, "mccnn_lr_flags: null argument");
    if (int e = check_hw("mccnn_lr_flags", H, W)) return e;
    lr_flags_kernel<<<dim3(ceil_div(W, 128), H), 128, 0, (cudaStream_t)stream>>>(dispL, dispR, flagL, flagR, H, W);
    MCCNN_LAUNCH_CHECK("lr_flags_kernel");
    return 0;
}

extern "C" size_t mccnn_lrc_fill_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    return 2 * (((size_t)H * W * sizeof(float) + 255) & ~(size_t)255);
}

extern "C" int mccnn_lrc_fill(const float* dispL, const uint8_t* flagL, float* filled, void* workspace, size_t workspace_bytes,
                              int H, int W, void* stream) {
    MCCNN_REQUIRE(dispL && flagL && filled && workspace, MCCNN_EINVAL, "mccnn_lrc_fill: null argument");
    MCCNN_REQUIRE(dispL != filled, MCCNN_EINVAL, "mccnn_lrc_fill: output must not alias the input map");
    if (int e = check_hw("mccnn_lrc_fill", H, W)) return e;
    MCCNN_REQUIRE(workspace_bytes >= mccnn_lrc_fill_workspace_bytes(H, W), MCCNN_EWORKSPACE, "mccnn_lrc_fill: workspace too small");
    float* up = reinterpret_cast<float*>(workspace);
    float* down = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + mccnn_lrc_fill_workspace_bytes(H, W) / 2);
    lrc_vertical_kernel<<<ceil_div(W, 128), 128, 0, (cudaStream_t)stream>>>(dispL, flagL, up, down, H, W);
    MCCNN_LAUNCH_CHECK("lrc_vertical_kernel");
    lrc_horizontal_kernel<<<ceil_div(H, 8), 256, 0, (cudaStream_t)stream>>>(dispL, flagL, up, down, filled, H, W);
    MCCNN_LAUNCH_CHECK("lrc_horizontal_kernel");
    return 0;
}

extern "C" int mccnn_median5(const float* filled, const float* wta, float* out, int H, int W, void* stream) {
    MCCNN_REQUIRE(filled && wta && out, MCCNN_EINVAL, "mccnn_median5: null argument");
    MCCNN_REQUIRE(filled != out, MCCNN_EINVAL, "mccnn_median5: out must not alias the filtered input");
    if (int e = check_hw("mccnn_median5", H, W)) return e;
    median5_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(filled, wta, out, H, W);
    MCCNN_LAUNCH_CHECK("median5_kernel");
    return 0;
}

extern "C" int mccnn_bilateral9(const uint8_t* image, const float* disp, float* out, int H, int W, void* stream) {
    MCCNN_REQUIRE(image && disp && out && disp != out, MCCNN_EINVAL, "mccnn_bilateral9: null or aliased argument");
    if (int e = check_hw("mccnn_bilateral9", H, W)) return e;
    bilateral9_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(image, disp, out, H, W);
    MCCNN_LAUNCH_CHECK("bilateral9_kernel");
    return 0;
}

extern "C" int mccnn_encode_u8(const float* disp, uint8_t* out, int H, int W, int scale, void* stream) {
    MCCNN_REQUIRE(disp && out, MCCNN_EINVAL, "mccnn_encode_u8: null argument");
    if (int e = check_hw("mccnn_encode_u8", H, W)) return e;
    const size_t n = (size_t)H * W;
    encode_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(disp, out, n, scale);
    MCCNN_LAUNCH_CHECK("encode_u8_kernel");
    return 0;
}

extern "C" int mccnn_encode_u16(const float* disp, uint16_t* out, int H, int W, int frac_bits, void* stream) {
    MCCNN_REQUIRE(disp && out, MCCNN_EINVAL, "mccnn_encode_u16: null argument");
    MCCNN_REQUIRE(frac_bits >= 0 && frac_bits <= 8, MCCNN_EINVAL, "mccnn_encode_u16: frac_bits=%d outside 0..8", frac_bits);
    if (int e = check_hw("mccnn_encode_u16", H, W)) return e;
    const size_t n = (size_t)H * W;
    encode_u16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(disp, out, n, (float)(1 << frac_bits));
    MCCNN_LAUNCH_CHECK("encode_u16_kernel");
    return 0;
}

extern "C" int mccnn_wta_subpixel(const float* S, float* disp, int H, int W, int D, void* stream) {
    MCCNN_REQUIRE(S && disp, MCCNN_EINVAL, "mccnn_wta_subpixel: null argument");
    if (int e = check_hw("mccnn_wta_subpixel", H, W)) return e;
    MCCNN_REQUIRE(D >= 1, MCCNN_EINVAL, "mccnn_wta_subpixel: D=%d", D);
    const int npix = H * W;
    wta_subpixel_kernel<<<ceil_div(npix, 8), 256, 0, (cudaStream_t)stream>>>(S, disp, npix, D, disp_pitch(D));
    MCCNN_LAUNCH_CHECK("wta_subpixel_kernel");
    return 0;
}

extern "C" int mccnn_bad_pixels(const uint8_t* disp_u8, const float* gt_half, unsigned long long* counts2, int H, int W,
                                void* stream) {
    MCCNN_REQUIRE(disp_u8 && gt_half && counts2, MCCNN_EINVAL, "mccnn_bad_pixels: null argument");
    if (int e = check_hw("mccnn_bad_pixels", H, W)) return e;
    MCCNN_CUDA(cudaMemsetAsync(counts2, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));
    const size_t n = (size_t)H * W;
    bad_pixels_kernel<unsigned char><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(disp_u8, gt_half, counts2, n);
    MCCNN_LAUNCH_CHECK("bad_pixels_kernel");
    return 0;
}

extern "C" int mccnn_bad_pixels_u16(const uint16_t* disp_u16, const float* gt_half, unsigned long long* counts2, int H, int W,
                                    void* stream) {
    MCCNN_REQUIRE(disp_u16 && gt_half && counts2, MCCNN_EINVAL, "mccnn_bad_pixels_u16: null argument");
    if (int e = check_hw("mccnn_bad_pixels_u16", H, W)) return e;
    MCCNN_CUDA(cudaMemsetAsync(counts2, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));
    const size_t n = (size_t)H * W;
    bad_pixels_kernel<unsigned short><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(disp_u16, gt_half, counts2, n);
    MCCNN_LAUNCH_CHECK("bad_pixels_kernel");
    return 0;
}
