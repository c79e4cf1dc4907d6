// Siamese branch conv tower (sm_100a).
//
// Replaces Net.construct / conv (mc_cnn_brunch.py:31-48, 70-92) as executed by compute_feature
// (process_functional.py:21-45) and match.py:72-73: `num_layers` 3x3 VALID stride-1 convolutions
// (1->64, then 64->64), bias, ReLU on all layers but the last, then tf.nn.l2_normalize over the 64
// channels (x * rsqrt(max(sum x^2, 1e-12))). The input was zero-padded ONCE by num_layers pixels.
//
// This file holds the fp32 CUDA-core implementation (exact fp32 products, fp32 accumulation): it is
// the numerically trusted path and the on-device reference for the tcgen05 implementation.
//
// Activations are NHWC fp32 [h][w][64]; weights are the reference's HWIO tensors, i.e. [tap][cin][cout].
#include "common.cuh"
#include <cstring>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;

// ---- layer 1: 1 -> 64 channels, 9 taps: thread = (pixel, 4 output channels)
__global__ void __launch_bounds__(256) conv1_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                   const float* __restrict__ b, float* __restrict__ out, int Hin, int Win,
                                                   int relu) {
    __shared__ float ws[9 * NF];
    __shared__ float bs[NF];
    for (int i = threadIdx.x; i < 9 * NF; i += 256) ws[i] = w[i];
    if (threadIdx.x < NF) bs[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const int Hout = Hin - 2, Wout = Win - 2;
    const size_t npix = (size_t)Hout * Wout;
    const size_t gid = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t pix = gid >> 4;
    const int q = (int)(gid & 15);
    if (pix >= npix) return;
    const int y = (int)(pix / Wout), x = (int)(pix % Wout);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
            const float v = in[(size_t)(y + ky) * Win + x + kx];
            const float4 wv = *reinterpret_cast<const float4*>(&ws[(ky * 3 + kx) * NF + 4 * q]);
            acc[0] = fmaf(v, wv.x, acc[0]);
            acc[1] = fmaf(v, wv.y, acc[1]);
            acc[2] = fmaf(v, wv.z, acc[2]);
            acc[3] = fmaf(v, wv.w, acc[3]);
        }
    float4 o;
    o.x = acc[0] + bs[4 * q];
    o.y = acc[1] + bs[4 * q + 1];
    o.z = acc[2] + bs[4 * q + 2];
    o.w = acc[3] + bs[4 * q + 3];
    if (relu) {
        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *reinterpret_cast<float4*>(&out[pix * NF + 4 * q]) = o;
}

// ---- layers 2..n: 64 -> 64 channels. CTA = 8 x 16 output pixels x 64 channels, 256 threads,
// thread = (row r, 8 consecutive columns) x 4 output channels, input channels in chunks of 16.
constexpr int TILE_H = 8, TILE_W = 16, CCH = 16;
constexpr int PATCH_H = TILE_H + 2, PATCH_W = TILE_W + 2;

struct ConvSmem {
    float in[PATCH_H * PATCH_W][CCH];
    float w[9][CCH][NF];
};

template <bool LAST>
__global__ void __launch_bounds__(256) conv64_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                    const float* __restrict__ b, float* __restrict__ out, int Hin, int Win) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ConvSmem& sm = *reinterpret_cast<ConvSmem*>(smem_raw);
    const int Hout = Hin - 2, Wout = Win - 2;
    const int tx0 = blockIdx.x * TILE_W, ty0 = blockIdx.y * TILE_H;
    const int tid = threadIdx.x;
    const int q = tid & 15;   // output channels 4q..4q+3
    const int g = tid >> 4;   // pixel group
    const int r = g >> 1, c0 = (g & 1) * 8;

    float acc[8][4];
#pragma unroll
    for (int p = 0; p < 8; p++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[p][c] = 0.f;

    for (int cb = 0; cb < NF; cb += CCH) {
        __syncthreads();
        for (int i = tid; i < PATCH_H * PATCH_W * (CCH / 4); i += 256) {
            const int pp = i / (CCH / 4), c4 = i % (CCH / 4);
            const int py = ty0 + pp / PATCH_W, px = tx0 + pp % PATCH_W;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (py < Hin && px < Win) v = *reinterpret_cast<const float4*>(&in[((size_t)py * Win + px) * NF + cb + 4 * c4]);
            *reinterpret_cast<float4*>(&sm.in[pp][4 * c4]) = v;
        }
        for (int i = tid; i < 9 * CCH * (NF / 4); i += 256) {
            const int c4 = i % (NF / 4), ci = (i / (NF / 4)) % CCH, tap = i / (NF / 4 * CCH);
            *reinterpret_cast<float4*>(&sm.w[tap][ci][4 * c4]) =
                *reinterpret_cast<const float4*>(&w[((size_t)tap * NF + cb + ci) * NF + 4 * c4]);
        }
        __syncthreads();
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
#pragma unroll
            for (int cq = 0; cq < CCH / 4; cq++) {
                float4 iv[10];
#pragma unroll
                for (int p = 0; p < 10; p++) iv[p] = *reinterpret_cast<const float4*>(&sm.in[(r + ky) * PATCH_W + c0 + p][4 * cq]);
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
#pragma unroll
                    for (int cc = 0; cc < 4; cc++) {
                        const float4 wv = *reinterpret_cast<const float4*>(&sm.w[ky * 3 + kx][4 * cq + cc][4 * q]);
#pragma unroll
                        for (int p = 0; p < 8; p++) {
                            const float4 t = iv[p + kx];
                            const float v = cc == 0 ? t.x : cc == 1 ? t.y : cc == 2 ? t.z : t.w;
                            acc[p][0] = fmaf(v, wv.x, acc[p][0]);
                            acc[p][1] = fmaf(v, wv.y, acc[p][1]);
                            acc[p][2] = fmaf(v, wv.z, acc[p][2]);
                            acc[p][3] = fmaf(v, wv.w, acc[p][3]);
                        }
                    }
                }
            }
        }
    }
    const float4 bv = *reinterpret_cast<const float4*>(&b[4 * q]);
    const int oy = ty0 + r;
#pragma unroll
    for (int p = 0; p < 8; p++) {
        float4 o = make_float4(acc[p][0] + bv.x, acc[p][1] + bv.y, acc[p][2] + bv.z, acc[p][3] + bv.w);
        if (!LAST) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        } else {
            // l2_normalize over the 64 channels of this pixel: 16 lanes (q = 0..15) hold 4 channels each
            float ss = o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            ss += __shfl_xor_sync(0xffffffffu, ss, 8);
            const float s = 1.0f / sqrtf(fmaxf(ss, 1e-12f));
            o.x *= s; o.y *= s; o.z *= s; o.w *= s;
        }
        const int ox = tx0 + c0 + p;
        if (oy < Hout && ox < Wout) *reinterpret_cast<float4*>(&out[((size_t)oy * Wout + ox) * NF + 4 * q]) = o;
    }
}

inline size_t layer_w_floats(int layer) { return (size_t)9 * (layer == 0 ? 1 : NF) * NF; }
inline size_t layer_offset_floats(int layer) {
    size_t o = 0;
    for (int i = 0; i < layer; i++) o += layer_w_floats(i) + NF;
    return o;
}
// packed blob = [fp32 section: per layer HWIO weights + bias][pad to 1 KB][tensor-core section: per layer >= 2
// the fp16 hi/lo tiles in their shared-memory image, conv_tc.cu]
inline size_t tc_section_offset(int num_layers) { return (layer_offset_floats(num_layers) * sizeof(float) + 1023) & ~(size_t)1023; }

}  // namespace

// conv_tc.cu
void pack_tc_weights(const float* hwio, unsigned char* dst);
size_t tc_weight_bytes_per_layer();
int conv_tower_tc(const float* padded, const float* w_fp32, const size_t* layer_off_floats, const unsigned char* w_tc,
                  float* features, void* workspace, size_t half_bytes, int H, int W, int num_layers, cudaStream_t stream);

}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_conv_packed_weight_bytes(int num_layers) {
    if (num_layers < 2 || num_layers > 16) return 0;
    return tc_section_offset(num_layers) + (size_t)(num_layers - 1) * tc_weight_bytes_per_layer();
}

extern "C" int mccnn_pack_weights_host(const float* const* hwio_host, const float* const* bias_host, int num_layers,
                                       void* packed_host) {
    MCCNN_REQUIRE(hwio_host && bias_host && packed_host, MCCNN_EINVAL, "mccnn_pack_weights_host: null argument");
    MCCNN_REQUIRE(num_layers >= 2 && num_layers <= 16, MCCNN_EINVAL, "mccnn_pack_weights_host: num_layers=%d outside 2..16",
                  num_layers);
    float* dst = reinterpret_cast<float*>(packed_host);
    for (int l = 0; l < num_layers; l++) {
        MCCNN_REQUIRE(hwio_host[l] && bias_host[l], MCCNN_EINVAL, "mccnn_pack_weights_host: layer %d missing", l + 1);
        // HWIO [3][3][Cin][64] is already [tap][cin][cout]
        memcpy(dst + layer_offset_floats(l), hwio_host[l], layer_w_floats(l) * sizeof(float));
        memcpy(dst + layer_offset_floats(l) + layer_w_floats(l), bias_host[l], NF * sizeof(float));
        if (l >= 1)
            pack_tc_weights(hwio_host[l], reinterpret_cast<unsigned char*>(packed_host) + tc_section_offset(num_layers) +
                                              (size_t)(l - 1) * tc_weight_bytes_per_layer());
    }
    return 0;
}

extern "C" size_t mccnn_conv_workspace_bytes(int H, int W, int num_layers) {
    if (H < 1 || W < 1 || num_layers < 2) return 0;
    // two ping-pong activation buffers sized for the largest (first) layer output
    const size_t a = (size_t)(H + 2 * num_layers - 2) * (W + 2 * num_layers - 2) * NF * sizeof(float);
    return 2 * ((a + 255) & ~(size_t)255);
}

extern "C" int mccnn_conv_tower(const float* padded, const void* packed_weights, float* features, void* workspace,
                                size_t workspace_bytes, int H, int W, int num_layers, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(padded && packed_weights && features && workspace, MCCNN_EINVAL, "mccnn_conv_tower: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && num_layers >= 2 && num_layers <= 16, MCCNN_EINVAL,
                  "mccnn_conv_tower: bad shape H=%d W=%d layers=%d", H, W, num_layers);
    MCCNN_REQUIRE(workspace_bytes >= mccnn_conv_workspace_bytes(H, W, num_layers), MCCNN_EWORKSPACE,
                  "mccnn_conv_tower: workspace too small");
    MCCNN_REQUIRE(aligned16(features) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0 && aligned16(packed_weights),
                  MCCNN_EALIGN, "mccnn_conv_tower: features/weights must be 16-byte and workspace 256-byte aligned");
    MCCNN_REQUIRE((size_t)(H + 2 * num_layers) * (W + 2 * num_layers) < 0x7fffffffu, MCCNN_EINVAL, "mccnn_conv_tower: image too large");
    size_t offs[16];
    for (int l = 0; l < num_layers; l++) offs[l] = layer_offset_floats(l);
    const unsigned char* blob = reinterpret_cast<const unsigned char*>(packed_weights);
    return conv_tower_tc(padded, reinterpret_cast<const float*>(packed_weights), offs, blob + tc_section_offset(num_layers),
                         features, workspace, mccnn_conv_workspace_bytes(H, W, num_layers) / 2, H, W, num_layers, stream);
}

extern "C" int mccnn_conv_tower_fp32(const float* padded, const void* packed_weights, float* features, void* workspace,
                                     size_t workspace_bytes, int H, int W, int num_layers, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(padded && packed_weights && features && workspace, MCCNN_EINVAL, "mccnn_conv_tower_fp32: null argument");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && num_layers >= 2 && num_layers <= 16, MCCNN_EINVAL,
                  "mccnn_conv_tower: bad shape H=%d W=%d layers=%d", H, W, num_layers);
    MCCNN_REQUIRE(workspace_bytes >= mccnn_conv_workspace_bytes(H, W, num_layers), MCCNN_EWORKSPACE,
                  "mccnn_conv_tower: workspace too small");
    MCCNN_REQUIRE(aligned16(features) && aligned16(workspace) && aligned16(packed_weights), MCCNN_EALIGN,
                  "mccnn_conv_tower: pointers must be 16-byte aligned");
    const float* wts = reinterpret_cast<const float*>(packed_weights);
    const size_t half = mccnn_conv_workspace_bytes(H, W, num_layers) / 2;
    float* buf[2] = {reinterpret_cast<float*>(workspace), reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + half)};

    int Hin = H + 2 * num_layers, Win = W + 2 * num_layers;
    {
        const size_t npix = (size_t)(Hin - 2) * (Win - 2);
        const float* w = wts + layer_offset_floats(0);
        conv1_kernel<<<(unsigned)((npix * 16 + 255) / 256), 256, 0, stream>>>(padded, w, w + layer_w_floats(0), buf[0], Hin, Win, 1);
        MCCNN_LAUNCH_CHECK("conv1_kernel");
        Hin -= 2; Win -= 2;
    }
    const size_t smem = sizeof(ConvSmem);
    if (int e = kernel_setup<conv64_kernel<false>>(256, smem, nullptr)) return e;
    if (int e = kernel_setup<conv64_kernel<true>>(256, smem, nullptr)) return e;
    int cur = 0;
    for (int l = 1; l < num_layers; l++) {
        const float* w = wts + layer_offset_floats(l);
        const float* b = w + layer_w_floats(l);
        const bool last = (l == num_layers - 1);
        float* dst = last ? features : buf[cur ^ 1];
        dim3 grid(ceil_div(Win - 2, TILE_W), ceil_div(Hin - 2, TILE_H));
        if (last)
            conv64_kernel<true><<<grid, 256, smem, stream>>>(buf[cur], w, b, dst, Hin, Win);
        else
            conv64_kernel<false><<<grid, 256, smem, stream>>>(buf[cur], w, b, dst, Hin, Win);
        MCCNN_LAUNCH_CHECK("conv64_kernel");
        cur ^= 1;
        Hin -= 2; Win -= 2;
    }
    return 0;
}
