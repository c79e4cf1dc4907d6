// Whole-path orchestration behind two C entry points (sm_100a).
//
//   mccnn_disparity_pipeline  replaces disparity_compute_by_gpu (process_functional.py:1093-1267)
//   mccnn_match_pair          replaces the body of match_single.py:34-55 / the loop body of match.py:51-90
//
// All buffers live in one caller-provided workspace (no allocation here, unlike the reference's
// ~15 cuda.to_device calls including four host-initialised volumes, :1103-1124). Stages are enqueued
// back to back on the caller's stream; the 7-slot stage timing of the reference's detail_time
// (match.py:95-103) is taken with CUDA events when the caller asks for it.
#include "common.cuh"

namespace mccnn {
namespace {

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct PipeLayout {
    size_t CL, CR, SL, SR, dl_wta, filled, flagL, armsL, armsR, sgm_ws, lrc_ws, fc_ws, end;
};

PipeLayout pipe_layout(int H, int W, int D, bool accurate = false) {
    PipeLayout l{};
    const size_t vol = align_up((size_t)H * W * disp_pitch(D) * sizeof(float));
    const size_t map = align_up((size_t)H * W * sizeof(float));
    size_t o = 0;
    l.CL = o; o += vol;
    l.CR = o; o += vol;
    l.SL = o; o += vol;
    l.SR = o; o += vol;
    l.dl_wta = o; o += map;
    l.filled = o; o += map;
    l.flagL = o; o += align_up((size_t)H * W);
    l.armsL = o; o += align_up((size_t)H * W * 4);
    l.armsR = o; o += align_up((size_t)H * W * 4);
    l.sgm_ws = o; o += align_up(mccnn_sgm_workspace_bytes(H, W, D));
    l.lrc_ws = o; o += align_up(mccnn_lrc_fill_workspace_bytes(H, W));
    l.fc_ws = o; o += accurate ? align_up(mccnn_fc_head_workspace_bytes(H, W)) : 0;  // fc1 outputs of the MC-CNN-accurate head
    l.end = o;
    return l;
}

struct MatchLayout {
    size_t padL, padR, featL, featR, conv_ws, scratch, pipe, end;
};

MatchLayout match_layout(int H, int W, int D, int nl, bool accurate = false) {
    MatchLayout l{};
    const size_t pad = align_up((size_t)(H + 2 * nl) * (W + 2 * nl) * sizeof(float));
    const size_t feat = align_up((size_t)H * W * MCCNN_FEATURES * sizeof(float));
    size_t o = 0;
    l.padL = o; o += pad;
    l.padR = o; o += pad;
    l.featL = o; o += feat;
    l.featR = o; o += feat;
    l.conv_ws = o; o += align_up(mccnn_conv_workspace_bytes(H, W, nl));
    l.scratch = o; o += 256;
    l.pipe = o; o += pipe_layout(H, W, D, accurate).end;
    l.end = o;
    return l;
}

// Stage timing events: created once per host thread and device and kept (9 events; cudaEventCreate / Destroy on every timed
// call cost more than a small-D SGM pass, and a failure half-way used to leak the ones already created).
struct StageTimer {
    cudaEvent_t* ev = nullptr;
    int n = 0;
    bool on = false;
    cudaStream_t s{};
    int begin(cudaStream_t stream, bool enable) {
        on = enable;
        s = stream;
        n = 0;
        if (!on) return 0;
        struct PerDevice { cudaEvent_t ev[9]; bool made; };
        static thread_local PerDevice cache[64] = {};
        PerDevice& c = cache[current_device() & 63];
        if (!c.made) {
            int made = 0;
            cudaError_t err = cudaSuccess;
            for (; made < 9 && err == cudaSuccess; made++) err = cudaEventCreate(&c.ev[made]);
            if (err != cudaSuccess) {
                for (int i = 0; i < made - 1; i++) cudaEventDestroy(c.ev[i]);
                on = false;
                return cuda_fail(err, "cudaEventCreate");
            }
            c.made = true;
        }
        ev = c.ev;
        return mark();
    }
    int mark() {
        if (!on) return 0;
        MCCNN_CUDA(cudaEventRecord(ev[n++], s));
        return 0;
    }
    void destroy() { on = false; }   // the events stay cached for the next timed call of this thread
};

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_pipeline_workspace_bytes(int H, int W, int D) {
    if (H < 1 || W < 1 || D < 1) return 0;
    return pipe_layout(H, W, D).end;
}

extern "C" size_t mccnn_match_workspace_bytes(int H, int W, int D, int num_layers) {
    if (H < 1 || W < 1 || D < 1 || num_layers < 2) return 0;
    return match_layout(H, W, D, num_layers).end;
}

static int run_pipeline(const uint8_t* imageL, const uint8_t* imageR, const float* fl, const float* fr, float* dispL_out,
                        float* dispR_out, char* ws, int H, int W, int D, const mccnn_sgm_params* params, int mode,
                        StageTimer& tm, cudaStream_t stream, const mccnn_fc_weights* head = nullptr) {
    const PipeLayout l = pipe_layout(H, W, D, head != nullptr);
    float* CL = reinterpret_cast<float*>(ws + l.CL);
    float* CR = reinterpret_cast<float*>(ws + l.CR);
    float* SL = reinterpret_cast<float*>(ws + l.SL);
    float* SR = reinterpret_cast<float*>(ws + l.SR);
    float* dl_wta = reinterpret_cast<float*>(ws + l.dl_wta);
    float* filled = reinterpret_cast<float*>(ws + l.filled);
    uint8_t* flagL = reinterpret_cast<uint8_t*>(ws + l.flagL);

    // Exact cost volume, two bit-identical implementations (tests/test_gpu_parity.py): the band GEMM on the CUDA cores and
    // the tensor-core variant, which wins once the disparity band is wide enough to fill its 128 x 32 tiles (measured on
    // B200: c4, D = 800: 56.3 vs 59.9 ms; c3, D = 400: 9.2 vs 8.6 ms). Its workspace borrows the S volumes, idle until SGM.
    const size_t tc_ws = mccnn_cost_volume_tc_workspace_bytes(H, W);
    if (head != nullptr) {
        // MC-CNN-accurate: the matching cost is the fully-connected head on the two feature vectors (fc_head.cu)
        if (int e = mccnn_cost_volume_accurate(fl, fr, head, CL, CR, ws + l.fc_ws, l.end - l.fc_ws, H, W, D, 1.0f, stream)) return e;
    } else if (mode == MCCNN_SGM_FUSED) {
        // opt-in tolerance mode: fp32-accumulated cost volume (1e-4 contract of north_star, not the reference's bits)
        // (tensor-core band GEMM; its fp16 operand planes borrow the S volumes, idle until SGM)
        const size_t ft_ws = mccnn_cost_volume_fast_tc_workspace_bytes(H, W);
        if (ft_ws <= l.dl_wta - l.SL) {
            if (int e = mccnn_cost_volume_fast_tc(fl, fr, CL, CR, ws + l.SL, l.dl_wta - l.SL, H, W, D, 1.0f, stream)) return e;
        } else {
            if (int e = mccnn_cost_volume_fast(fl, fr, CL, CR, H, W, D, 1.0f, stream)) return e;
        }
    } else if (D >= 512 && tc_ws <= l.dl_wta - l.SL) {
        if (int e = mccnn_cost_volume_tc(fl, fr, CL, CR, ws + l.SL, l.dl_wta - l.SL, H, W, D, 1.0f, stream)) return e;
    } else {
        if (int e = mccnn_cost_volume(fl, fr, CL, CR, H, W, D, 1.0f, stream)) return e;
    }
    if (int e = tm.mark()) return e;  // [1] cost volume
    if (params->cbca_iters > 0) {
        // the stage behind the reference's unused "aggregation" slot (match.py:98); SL / SR are free until SGM starts
        uint8_t* armsL = reinterpret_cast<uint8_t*>(ws + l.armsL);
        uint8_t* armsR = reinterpret_cast<uint8_t*>(ws + l.armsR);
        if (int e = mccnn_cross_arms(imageL, armsL, H, W, params->cbca_L1, params->cbca_tau, stream)) return e;
        if (int e = mccnn_cross_arms(imageR, armsR, H, W, params->cbca_L1, params->cbca_tau, stream)) return e;
        float *curL = CL, *curR = CR, *altL = SL, *tmp = SR;
        for (int it = 0; it < params->cbca_iters; it++) {
            // left: cur -> alt (row sums in tmp); right: cur -> (old left buffer is free now)
            if (int e = mccnn_cbca(curL, altL, tmp, armsL, armsR, H, W, D, -1, params->cbca_L1, stream)) return e;
            float* freeL = curL;
            curL = altL;
            if (int e = mccnn_cbca(curR, freeL, tmp, armsR, armsL, H, W, D, 1, params->cbca_L1, stream)) return e;
            altL = curR;
            curR = freeL;
        }
        // the four buffers are interchangeable from here on: SGM reads (curL, curR) and writes the other two
        float* others[2];
        int n = 0;
        float* all4[4] = {CL, CR, SL, SR};
        for (int i = 0; i < 4; i++)
            if (all4[i] != curL && all4[i] != curR) others[n++] = all4[i];
        CL = curL; CR = curR; SL = others[0]; SR = others[1];
    }
    if (int e = tm.mark()) return e;  // [2] aggregation
    if (int e = mccnn_sgm(CL, CR, imageL, imageR, SL, SR, dl_wta, dispR_out, ws + l.sgm_ws,
                          mccnn_sgm_workspace_bytes(H, W, D), H, W, D, params, mode, 0, stream))
        return e;
    if (int e = tm.mark()) return e;  // [3] SGM (+ fused WTA)
    if (int e = mccnn_lr_flags(dl_wta, dispR_out, flagL, nullptr, H, W, stream)) return e;
    if (int e = mccnn_lrc_fill(dl_wta, flagL, filled, ws + l.lrc_ws, mccnn_lrc_fill_workspace_bytes(H, W), H, W, stream)) return e;
    if (int e = tm.mark()) return e;  // [5] L-R check + fill
    if (int e = mccnn_median5(filled, dl_wta, dispL_out, H, W, stream)) return e;
    // the reference's (commented-out) bilateral launch reads the FILLED map and overwrites the median output (:1260)
    if (params->bilateral)
        if (int e = mccnn_bilateral9(imageL, filled, dispL_out, H, W, stream)) return e;
    if (int e = tm.mark()) return e;  // [6] filter
    return 0;
}

extern "C" int mccnn_disparity_pipeline(const uint8_t* imageL, const uint8_t* imageR, const float* fl, const float* fr,
                                        float* dispL_out, float* dispR_out, void* workspace, size_t workspace_bytes, int H,
                                        int W, int D, const mccnn_sgm_params* params, int mode, float* stage_ms_host,
                                        void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(imageL && imageR && fl && fr && dispL_out && dispR_out && workspace && params, MCCNN_EINVAL,
                  "mccnn_disparity_pipeline: null argument");
    MCCNN_REQUIRE(H >= 3 && W >= 3 && D >= 1 && D <= 1024, MCCNN_EINVAL, "mccnn_disparity_pipeline: bad shape H=%d W=%d D=%d",
                  H, W, D);
    MCCNN_REQUIRE(workspace_bytes >= mccnn_pipeline_workspace_bytes(H, W, D), MCCNN_EWORKSPACE,
                  "mccnn_disparity_pipeline: workspace too small (%zu < %zu)", workspace_bytes,
                  mccnn_pipeline_workspace_bytes(H, W, D));
    MCCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN,
                  "mccnn_disparity_pipeline: workspace must be 256-byte aligned");
    StageTimer tm;
    if (int e = tm.begin(stream, stage_ms_host != nullptr)) return e;
    int rc = run_pipeline(imageL, imageR, fl, fr, dispL_out, dispR_out, reinterpret_cast<char*>(workspace), H, W, D, params,
                          mode, tm, stream);
    if (rc == 0 && stage_ms_host) {
        cudaError_t e = cudaEventSynchronize(tm.ev[tm.n - 1]);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "cudaEventSynchronize");
        } else {
            float ms[5] = {0, 0, 0, 0, 0};
            for (int i = 0; i < 5; i++) cudaEventElapsedTime(&ms[i], tm.ev[i], tm.ev[i + 1]);
            stage_ms_host[1] += ms[0];  // cost volume
            stage_ms_host[2] += ms[1];  // "*" aggregation: a label without a stage in the reference (match.py:98); CBCA here
            stage_ms_host[3] += ms[2];  // SGM
            stage_ms_host[5] += ms[3];  // L-R check; slot [4] (WTA) is fused into the last SGM pass
            stage_ms_host[6] += ms[4];  // filter
        }
    }
    tm.destroy();
    return rc;
}

static int match_pair_impl(const uint8_t* imageL, const uint8_t* imageR, const void* packed_weights, const mccnn_fc_weights* head,
                           float* dispL_out, float* dispR_out, void* workspace, size_t workspace_bytes, int H, int W, int D,
                           int num_layers, const mccnn_sgm_params* params, int mode, float* stage_ms_host, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(imageL && imageR && packed_weights && dispL_out && dispR_out && workspace && params, MCCNN_EINVAL,
                  "mccnn_match_pair: null argument");
    MCCNN_REQUIRE(H >= 3 && W >= 3 && D >= 1 && D <= 1024 && num_layers >= 2 && num_layers <= 16, MCCNN_EINVAL,
                  "mccnn_match_pair: bad shape H=%d W=%d D=%d layers=%d", H, W, D, num_layers);
    const size_t need = match_layout(H, W, D, num_layers, head != nullptr).end;
    MCCNN_REQUIRE(workspace_bytes >= need, MCCNN_EWORKSPACE, "mccnn_match_pair: workspace too small (%zu < %zu)", workspace_bytes, need);
    MCCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN,
                  "mccnn_match_pair: workspace must be 256-byte aligned");
    char* ws = reinterpret_cast<char*>(workspace);
    const MatchLayout l = match_layout(H, W, D, num_layers, head != nullptr);
    float* padL = reinterpret_cast<float*>(ws + l.padL);
    float* padR = reinterpret_cast<float*>(ws + l.padR);
    float* featL = reinterpret_cast<float*>(ws + l.featL);
    float* featR = reinterpret_cast<float*>(ws + l.featR);
    double* scratch = reinterpret_cast<double*>(ws + l.scratch);
    const size_t conv_ws = mccnn_conv_workspace_bytes(H, W, num_layers);

    StageTimer tm;
    if (int e = tm.begin(stream, stage_ms_host != nullptr)) return e;
    int rc = 0;
    do {
        if ((rc = mccnn_standardize_pad(imageL, padL, scratch, H, W, num_layers, stream))) break;
        if ((rc = mccnn_standardize_pad(imageR, padR, scratch + 4, H, W, num_layers, stream))) break;
        if ((rc = mccnn_conv_tower(padL, packed_weights, featL, ws + l.conv_ws, conv_ws, H, W, num_layers, stream))) break;
        if ((rc = mccnn_conv_tower(padR, packed_weights, featR, ws + l.conv_ws, conv_ws, H, W, num_layers, stream))) break;
        if ((rc = tm.mark())) break;  // [0] features
        rc = run_pipeline(imageL, imageR, featL, featR, dispL_out, dispR_out, ws + l.pipe, H, W, D, params, mode, tm, stream, head);
    } while (0);
    if (rc == 0 && stage_ms_host) {
        cudaError_t e = cudaEventSynchronize(tm.ev[tm.n - 1]);
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "cudaEventSynchronize");
        } else {
            float ms[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 6; i++) cudaEventElapsedTime(&ms[i], tm.ev[i], tm.ev[i + 1]);
            stage_ms_host[0] += ms[0];
            stage_ms_host[1] += ms[1];
            stage_ms_host[2] += ms[2];
            stage_ms_host[3] += ms[3];
            stage_ms_host[5] += ms[4];
            stage_ms_host[6] += ms[5];
        }
    }
    tm.destroy();
    return rc;
}

extern "C" int mccnn_match_pair(const uint8_t* imageL, const uint8_t* imageR, const void* packed_weights, float* dispL_out,
                                float* dispR_out, void* workspace, size_t workspace_bytes, int H, int W, int D,
                                int num_layers, const mccnn_sgm_params* params, int mode, float* stage_ms_host,
                                void* stream_) {
    return match_pair_impl(imageL, imageR, packed_weights, nullptr, dispL_out, dispR_out, workspace, workspace_bytes, H, W, D,
                           num_layers, params, mode, stage_ms_host, stream_);
}

extern "C" size_t mccnn_match_accurate_workspace_bytes(int H, int W, int D, int num_layers) {
    if (H < 1 || W < 1 || D < 1 || num_layers < 2) return 0;
    return match_layout(H, W, D, num_layers, true).end;
}

extern "C" int mccnn_match_pair_accurate(const uint8_t* imageL, const uint8_t* imageR, const void* packed_weights,
                                         const mccnn_fc_weights* head, float* dispL_out, float* dispR_out, void* workspace,
                                         size_t workspace_bytes, int H, int W, int D, int num_layers,
                                         const mccnn_sgm_params* params, int mode, float* stage_ms_host, void* stream_) {
    MCCNN_REQUIRE(head != nullptr, MCCNN_EINVAL, "mccnn_match_pair_accurate: null head weights");
    return match_pair_impl(imageL, imageR, packed_weights, head, dispL_out, dispR_out, workspace, workspace_bytes, H, W, D,
                           num_layers, params, mode, stage_ms_host, stream_);
}
