// One training step of the siamese tower: forward, hinge loss, backward, momentum update (sm_100a).
//
// Replaces the graph train.py builds and runs per batch (train.py:71-99, executed by sess.run at :143-150):
//   three weight-sharing branches Net(..., num_of_conv_layers=5, num_of_conv_feature_maps=64) on [B,p,p,1] patches
//   (mc_cnn_brunch.py:31-48: 3x3 VALID conv + bias, ReLU on all but the last layer, l2_normalize), squeezed to [B,64];
//   cosine_pos = <fl, fr+>, cosine_neg = <fl, fr->, loss = mean(max(0, margin - cosine_pos + cosine_neg)) (:83-89);
//   tf.train.MomentumOptimizer(lr, beta): accum = beta * accum + grad; var -= lr * accum (:97-99).
// The arithmetic of that graph belongs to TensorFlow (not installed, version unpinned): parity is held to a tolerance
// against oracle/train_step.py (torch CPU fp64 autograd), not bit for bit.
//
// The workload is tiny (3 * B = 384 patches of 11x11, 9.4 MFLOP per patch and direction), launch-latency bound: plain
// CUDA-core kernels, one thread per output element, the three branches batched as one N = 3B tensor. fp32 throughout.
#include "common.cuh"

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;

// out[n][y][x][co] = bias[co] + sum_{ky,kx,ci} in[n][y+ky][x+kx][ci] * w[ky][kx][ci][co], optional ReLU
__global__ void __launch_bounds__(256) tr_conv_fwd(const float* __restrict__ in, const float* __restrict__ w,
                                                  const float* __restrict__ b, float* __restrict__ out, int N, int hin, int cin,
                                                  int relu) {
    const int hout = hin - 2;
    const long long total = (long long)N * hout * hout * NF;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int co = (int)(i % NF);
    long long r = i / NF;
    const int x = (int)(r % hout); r /= hout;
    const int y = (int)(r % hout);
    const int n = (int)(r / hout);
    float acc = b[co];
    for (int ky = 0; ky < 3; ky++)
        for (int kx = 0; kx < 3; kx++) {
            const float* ip = in + (((size_t)n * hin + y + ky) * hin + x + kx) * cin;
            const float* wp = w + (size_t)((ky * 3 + kx) * cin) * NF + co;
            for (int ci = 0; ci < cin; ci++) acc = fmaf(ip[ci], wp[(size_t)ci * NF], acc);
        }
    out[i] = relu ? fmaxf(acc, 0.f) : acc;
}

// feat[n][:] = x[n][:] * rsqrt(max(sum x^2, 1e-12)) (tf.nn.l2_normalize); inv[n] = that factor, clamped[n] = 1 if the max hit
__global__ void tr_l2norm_fwd(const float* __restrict__ x, float* __restrict__ feat, float* __restrict__ inv, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float ss = 0.f;
    for (int c = 0; c < NF; c++) ss = fmaf(x[(size_t)n * NF + c], x[(size_t)n * NF + c], ss);
    const float s = rsqrtf(fmaxf(ss, 1e-12f));
    for (int c = 0; c < NF; c++) feat[(size_t)n * NF + c] = x[(size_t)n * NF + c] * s;
    inv[n] = ss > 1e-12f ? s : -s;  // sign carries the clamp flag
}

// hinge loss and the gradient w.r.t. the three normalised features; one thread per batch element; loss via atomicAdd
__global__ void tr_loss(const float* __restrict__ feat, float* __restrict__ dfeat, float* __restrict__ loss, int B, float margin) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float *fl = feat + (size_t)b * NF, *fp = feat + (size_t)(B + b) * NF, *fn = feat + (size_t)(2 * B + b) * NF;
    float cp = 0.f, cn = 0.f;
    for (int c = 0; c < NF; c++) { cp = fmaf(fl[c], fp[c], cp); cn = fmaf(fl[c], fn[c], cn); }
    const float h = margin - cp + cn;
    const float g = h > 0.f ? 1.0f / (float)B : 0.f;  // d loss / d h
    if (h > 0.f) atomicAdd(loss, h / (float)B);
    for (int c = 0; c < NF; c++) {
        dfeat[(size_t)b * NF + c] = g * (fn[c] - fp[c]);
        dfeat[(size_t)(B + b) * NF + c] = -g * fl[c];
        dfeat[(size_t)(2 * B + b) * NF + c] = g * fl[c];
    }
}

// dx = s * (dy - y * <y, dy>) where y = x * s; when the norm was clamped the map is linear: dx = s * dy
__global__ void tr_l2norm_bwd(const float* __restrict__ feat, const float* __restrict__ dfeat, const float* __restrict__ inv,
                              float* __restrict__ dx, int N) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float s = fabsf(inv[n]);
    const bool clamped = inv[n] < 0.f;
    float dot = 0.f;
    for (int c = 0; c < NF; c++) dot = fmaf(feat[(size_t)n * NF + c], dfeat[(size_t)n * NF + c], dot);
    for (int c = 0; c < NF; c++) {
        const float dy = dfeat[(size_t)n * NF + c];
        dx[(size_t)n * NF + c] = clamped ? s * dy : s * (dy - feat[(size_t)n * NF + c] * dot);
    }
}

// gradient through the ReLU of a layer, in place: d *= (out > 0)
__global__ void tr_relu_bwd(float* __restrict__ d, const float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(out[i] > 0.f)) d[i] = 0.f;
}

// din[n][y][x][ci] = sum_{ky,kx,co} dout[n][y-ky][x-kx][co] * w[ky][kx][ci][co]
__global__ void __launch_bounds__(256) tr_conv_bwd_data(const float* __restrict__ dout, const float* __restrict__ w,
                                                       float* __restrict__ din, int N, int hin, int cin) {
    const int hout = hin - 2;
    const long long total = (long long)N * hin * hin * cin;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ci = (int)(i % cin);
    long long r = i / cin;
    const int x = (int)(r % hin); r /= hin;
    const int y = (int)(r % hin);
    const int n = (int)(r / hin);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ky++) {
        const int oy = y - ky;
        if (oy < 0 || oy >= hout) continue;
        for (int kx = 0; kx < 3; kx++) {
            const int ox = x - kx;
            if (ox < 0 || ox >= hout) continue;
            const float* dp = dout + (((size_t)n * hout + oy) * hout + ox) * NF;
            const float* wp = w + ((size_t)(ky * 3 + kx) * cin + ci) * NF;
            for (int co = 0; co < NF; co++) acc = fmaf(dp[co], wp[co], acc);
        }
    }
    din[i] = acc;
}

// dw[ky][kx][ci][co] = sum_{n,y,x} in[n][y+ky][x+kx][ci] * dout[n][y][x][co]; one block per (ky, kx, ci), thread = co,
// the (n, y, x) range split over gridDim.y blocks and combined with atomicAdd (dw zeroed by the caller)
__global__ void __launch_bounds__(NF) tr_conv_bwd_w(const float* __restrict__ in, const float* __restrict__ dout,
                                                   float* __restrict__ dw, int N, int hin, int cin) {
    const int hout = hin - 2;
    const int tap_ci = blockIdx.x;  // (ky * 3 + kx) * cin + ci
    const int ci = tap_ci % cin, tap = tap_ci / cin, ky = tap / 3, kx = tap % 3;
    const int co = threadIdx.x;
    const long long npos = (long long)N * hout * hout;
    const long long per = (npos + gridDim.y - 1) / gridDim.y;
    const long long p0 = (long long)blockIdx.y * per, p1 = min(npos, p0 + per);
    float acc = 0.f;
    for (long long p = p0; p < p1; p++) {
        const int x = (int)(p % hout);
        const long long r = p / hout;
        const int y = (int)(r % hout), n = (int)(r / hout);
        acc = fmaf(in[(((size_t)n * hin + y + ky) * hin + x + kx) * cin + ci], dout[(size_t)p * NF + co], acc);
    }
    atomicAdd(dw + (size_t)tap_ci * NF + co, acc);
}

// db[co] = sum over (n, y, x) of dout
__global__ void __launch_bounds__(NF) tr_conv_bwd_b(const float* __restrict__ dout, float* __restrict__ db, long long npos) {
    const int co = threadIdx.x;
    const long long per = (npos + gridDim.x - 1) / gridDim.x;
    const long long p0 = (long long)blockIdx.x * per, p1 = min(npos, p0 + per);
    float acc = 0.f;
    for (long long p = p0; p < p1; p++) acc += dout[(size_t)p * NF + co];
    atomicAdd(db + co, acc);
}

// tf.train.MomentumOptimizer: accum = beta * accum + grad; var -= lr * accum
__global__ void tr_momentum(float* __restrict__ var, float* __restrict__ accum, const float* __restrict__ grad, long long n, float lr,
                            float beta) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = fmaf(beta, accum[i], grad[i]);
    accum[i] = a;
    var[i] = fmaf(-lr, a, var[i]);
}

inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }
inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// offsets (in floats) of layer l's weights / biases in the flat parameter vector
inline size_t w_off(int l) { return l == 0 ? 0 : (size_t)(9 * NF + NF) + (size_t)(l - 1) * (9 * NF * NF + NF); }
inline size_t w_count(int l) { return (size_t)9 * (l == 0 ? 1 : NF) * NF; }

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_train_param_count(int num_layers) {
    if (num_layers < 1) return 0;
    return w_off(num_layers);
}

extern "C" size_t mccnn_train_workspace_bytes(int batch, int patch, int num_layers) {
    if (batch < 1 || num_layers < 1 || patch != 2 * num_layers + 1) return 0;
    const size_t N = 3 * (size_t)batch;
    size_t bytes = a256(N * patch * patch * sizeof(float));  // input patches of the three branches
    for (int l = 0; l < num_layers; l++) {
        const size_t h = patch - 2 * (l + 1);
        bytes += 2 * a256(N * h * h * NF * sizeof(float));  // activation + its gradient
    }
    bytes += a256(N * patch * patch * sizeof(float));        // gradient w.r.t. the input (discarded)
    bytes += 2 * a256(N * NF * sizeof(float)) + a256(N * sizeof(float));  // features, their gradient, 1 / norm
    bytes += a256(mccnn_train_param_count(num_layers) * sizeof(float));  // gradients
    return bytes + 256;
}

extern "C" int mccnn_train_step(const float* left, const float* right_pos, const float* right_neg, float* params, float* velocity,
                                float* grads_out, float* loss_out, void* workspace, size_t workspace_bytes, int batch, int patch,
                                int num_layers, float margin, float lr, float momentum, int apply_update, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(left && right_pos && right_neg && params && loss_out && workspace, MCCNN_EINVAL, "mccnn_train_step: null argument");
    MCCNN_REQUIRE(batch >= 1 && num_layers >= 1 && num_layers <= 16 && patch == 2 * num_layers + 1, MCCNN_EINVAL,
                  "mccnn_train_step: need patch = 2 * num_layers + 1 (the tower reduces a patch to one pixel), got patch=%d layers=%d",
                  patch, num_layers);
    MCCNN_REQUIRE(!apply_update || velocity, MCCNN_EINVAL, "mccnn_train_step: apply_update needs the momentum accumulators");
    MCCNN_REQUIRE(workspace_bytes >= mccnn_train_workspace_bytes(batch, patch, num_layers), MCCNN_EWORKSPACE,
                  "mccnn_train_step: workspace too small");
    MCCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN, "mccnn_train_step: workspace must be 256-byte aligned");
    const int B = batch, N = 3 * batch;
    char* ws = reinterpret_cast<char*>(workspace);
    size_t o = 0;
    auto take = [&](size_t bytes) { float* p = reinterpret_cast<float*>(ws + o); o += a256(bytes); return p; };
    float* in0 = take((size_t)N * patch * patch * sizeof(float));
    float* act[16];
    float* dact[16];
    for (int l = 0; l < num_layers; l++) {
        const size_t h = patch - 2 * (l + 1);
        act[l] = take((size_t)N * h * h * NF * sizeof(float));
        dact[l] = take((size_t)N * h * h * NF * sizeof(float));
    }
    float* din0 = take((size_t)N * patch * patch * sizeof(float));
    float* feat = take((size_t)N * NF * sizeof(float));
    float* dfeat = take((size_t)N * NF * sizeof(float));
    float* inv = take((size_t)N * sizeof(float));
    const size_t nparam = mccnn_train_param_count(num_layers);
    float* grads = take(nparam * sizeof(float));
    (void)din0;

    const size_t pbytes = (size_t)B * patch * patch * sizeof(float);
    MCCNN_CUDA(cudaMemcpyAsync(in0, left, pbytes, cudaMemcpyDeviceToDevice, stream));
    MCCNN_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(in0) + pbytes, right_pos, pbytes, cudaMemcpyDeviceToDevice, stream));
    MCCNN_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(in0) + 2 * pbytes, right_neg, pbytes, cudaMemcpyDeviceToDevice, stream));
    MCCNN_CUDA(cudaMemsetAsync(grads, 0, nparam * sizeof(float), stream));
    MCCNN_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), stream));

    // ---- forward
    for (int l = 0; l < num_layers; l++) {
        const int hin = patch - 2 * l, cin = l == 0 ? 1 : NF;
        const long long total = (long long)N * (hin - 2) * (hin - 2) * NF;
        tr_conv_fwd<<<blocks_for(total, 256), 256, 0, stream>>>(l == 0 ? in0 : act[l - 1], params + w_off(l), params + w_off(l) + w_count(l),
                                                              act[l], N, hin, cin, l + 1 < num_layers ? 1 : 0);
        MCCNN_LAUNCH_CHECK("tr_conv_fwd");
    }
    tr_l2norm_fwd<<<blocks_for(N, 128), 128, 0, stream>>>(act[num_layers - 1], feat, inv, N);
    MCCNN_LAUNCH_CHECK("tr_l2norm_fwd");
    tr_loss<<<blocks_for(B, 128), 128, 0, stream>>>(feat, dfeat, loss_out, B, margin);
    MCCNN_LAUNCH_CHECK("tr_loss");
    // ---- backward
    tr_l2norm_bwd<<<blocks_for(N, 128), 128, 0, stream>>>(feat, dfeat, inv, dact[num_layers - 1], N);
    MCCNN_LAUNCH_CHECK("tr_l2norm_bwd");
    for (int l = num_layers - 1; l >= 0; l--) {
        const int hin = patch - 2 * l, hout = hin - 2, cin = l == 0 ? 1 : NF;
        const long long nout = (long long)N * hout * hout * NF, npos = (long long)N * hout * hout;
        if (l + 1 < num_layers) {
            tr_relu_bwd<<<blocks_for(nout, 256), 256, 0, stream>>>(dact[l], act[l], nout);
            MCCNN_LAUNCH_CHECK("tr_relu_bwd");
        }
        const float* lin = l == 0 ? in0 : act[l - 1];
        const int split = (int)min((long long)64, max((long long)1, npos / 256));
        tr_conv_bwd_w<<<dim3(9 * cin, split), NF, 0, stream>>>(lin, dact[l], grads + w_off(l), N, hin, cin);
        MCCNN_LAUNCH_CHECK("tr_conv_bwd_w");
        tr_conv_bwd_b<<<split, NF, 0, stream>>>(dact[l], grads + w_off(l) + w_count(l), npos);
        MCCNN_LAUNCH_CHECK("tr_conv_bwd_b");
        if (l > 0) {
            const long long nin = (long long)N * hin * hin * cin;
            tr_conv_bwd_data<<<blocks_for(nin, 256), 256, 0, stream>>>(dact[l], params + w_off(l), dact[l - 1], N, hin, cin);
            MCCNN_LAUNCH_CHECK("tr_conv_bwd_data");
        }
    }
    if (grads_out) MCCNN_CUDA(cudaMemcpyAsync(grads_out, grads, nparam * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (apply_update) {
        tr_momentum<<<blocks_for((long long)nparam, 256), 256, 0, stream>>>(params, velocity, grads, (long long)nparam, lr, momentum);
        MCCNN_LAUNCH_CHECK("tr_momentum");
    }
    return 0;
}
