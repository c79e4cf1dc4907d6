// Image standardisation + one-off zero padding (sm_100a).
//
// Replaces match_single.py:34-43 / match.py:51-57 ((I - mean) / std with the population std, on the
// host in NumPy) and process_functional.py:13-19 / match.py:61-67 (zero-pad by (patch-1)/2 once; the
// tower then runs VALID convolutions).
//
// The u8 sums are accumulated exactly in 64-bit integers, so mean and std are the correctly rounded
// values; NumPy's fp32 pairwise sums differ from them in the last bits only (documented tolerance of
// the conv-tower stage, not bit-exact by construction).
#include "common.cuh"

namespace mccnn {
namespace {

__global__ void image_sums_kernel(const unsigned char* __restrict__ img, size_t n, unsigned long long* __restrict__ sums) {
    unsigned long long s = 0, s2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned v = img[i];
        s += v;
        s2 += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sums[0], s);
        atomicAdd(&sums[1], s2);
    }
}

__global__ void standardize_pad_kernel(const unsigned char* __restrict__ img, float* __restrict__ out,
                                       const unsigned long long* __restrict__ sums, int H, int W, int pad) {
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int yp = blockIdx.y;
    if (xp >= Wp || yp >= Hp) return;
    const double n = (double)H * (double)W;
    const double mean = (double)sums[0] / n;
    double var = (double)sums[1] / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float meanf = (float)mean, stdf = (float)sqrt(var);
    const int x = xp - pad, y = yp - pad;
    float v = 0.0f;
    if (x >= 0 && x < W && y >= 0 && y < H) v = ((float)img[(size_t)y * W + x] - meanf) / stdf;
    out[(size_t)yp * Wp + xp] = v;
}

__global__ void pad_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W, int pad) {
    const int Wp = W + 2 * pad, Hp = H + 2 * pad;
    const int xp = blockIdx.x * blockDim.x + threadIdx.x;
    const int yp = blockIdx.y;
    if (xp >= Wp || yp >= Hp) return;
    const int x = xp - pad, y = yp - pad;
    out[(size_t)yp * Wp + xp] = (x >= 0 && x < W && y >= 0 && y < H) ? img[(size_t)y * W + x] : 0.0f;
}

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" int mccnn_standardize_pad(const uint8_t* image, float* out_padded, double* scratch4, int H, int W, int pad,
                                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(image && out_padded && scratch4, MCCNN_EINVAL, "mccnn_standardize_pad: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && pad >= 0 && H + 2 * pad <= 65535, MCCNN_EINVAL, "mccnn_standardize_pad: bad shape");
    unsigned long long* sums = reinterpret_cast<unsigned long long*>(scratch4);
    MCCNN_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), stream));
    const size_t n = (size_t)H * W;
    int blocks = (int)((n + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    image_sums_kernel<<<blocks, 256, 0, stream>>>(image, n, sums);
    MCCNN_LAUNCH_CHECK("image_sums_kernel");
    standardize_pad_kernel<<<dim3(ceil_div(W + 2 * pad, 128), H + 2 * pad), 128, 0, stream>>>(image, out_padded, sums, H, W, pad);
    MCCNN_LAUNCH_CHECK("standardize_pad_kernel");
    return 0;
}

extern "C" int mccnn_pad_f32(const float* image, float* out_padded, int H, int W, int pad, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(image && out_padded, MCCNN_EINVAL, "mccnn_pad_f32: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && pad >= 0 && H + 2 * pad <= 65535, MCCNN_EINVAL, "mccnn_pad_f32: bad shape");
    pad_f32_kernel<<<dim3(ceil_div(W + 2 * pad, 128), H + 2 * pad), 128, 0, stream>>>(image, out_padded, H, W, pad);
    MCCNN_LAUNCH_CHECK("pad_f32_kernel");
    return 0;
}
