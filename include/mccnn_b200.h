/*
 * mccnn_b200 — C ABI of the B200-native MC-CNN stereo-matching hot path.
 *
 * Drop-in boundary for WHDY/SceneDepthEstimation's hot path. The reference has no FFI
 * layer: the path sits behind Python functions and Numba-CUDA launch objects, so every
 * entry point below names the reference function / kernel launch it replaces
 * (file:line relative to the reference root). The Python shim in
 * scenedepthestimation_b200/ binds these with ctypes and keeps the reference's
 * signatures.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / numba types.
 *  - Every pointer is a DEVICE pointer unless the name ends in _host. The caller owns all
 *    memory, including scratch: sizes come from the *_workspace_bytes queries. Nothing is
 *    allocated or freed behind the caller's back.
 *  - Every launch entry takes a cudaStream_t (passed as void*) and is asynchronous.
 *  - Return value: 0 on success, a positive cudaError_t, or a negative MCCNN_E* argument
 *    error. mccnn_last_error() returns a thread-local message. There is no CPU fallback.
 *  - Volumes are fp32 [H][W][Dp] with the disparity innermost and pitch
 *    Dp = mccnn_disp_pitch(D) (D rounded up to a multiple of 4). Pad entries [D, Dp) of a COST volume must
 *    be +INF (mccnn_cost_volume writes them so; the SGM kernels rely on it); those of S volumes are undefined.
 *    Feature maps are fp32 [H][W][64]. Images are u8 [H][W]. Disparity maps are fp32 [H][W]
 *    holding integer values, as in the reference.
 */
#ifndef MCCNN_B200_H
#define MCCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCCNN_ABI_VERSION 4
#define MCCNN_FEATURES 64 /* num_of_feature_maps: hard-coded 64 in the reference (process_functional.py:128) */

enum {
    MCCNN_OK = 0,
    MCCNN_EINVAL = -1,    /* bad size / null pointer / unsupported D */
    MCCNN_EALIGN = -2,    /* pointer not 16-byte aligned */
    MCCNN_EWORKSPACE = -3 /* workspace too small */
};

/* SGM arithmetic.
 * MCCNN_SGM_EXACT (default everywhere) reproduces the reference bit for bit: fp64 path state, fp32 S rounded once per path
 *   in launch order (process_functional.py:265-343, 1166-1202), fp64-accumulated cost volume (:120-131).
 * MCCNN_SGM_FUSED is this library's opt-in throughput mode, held to north_star's tolerance instead (costs and aggregated
 *   costs within 1e-4 relative; disparities equal except at near-ties, measured by tools/fused_census.py): the same
 *   recurrence, extents, wraps and penalties with fp32 path state, the 8 contributions added to S in the order down,
 *   down-right, up, left, down-left, right, up-right, up-left (4 sweeps over the volumes instead of 7, csrc/sgm_fused.cu),
 *   and, inside mccnn_disparity_pipeline / mccnn_match_pair, an fp32-accumulated cost volume. Needs all four penalties
 *   >= 0 (MCCNN_EINVAL otherwise; the exact mode only needs P1 >= 0). Environment: MCCNN_FUSED_STRICT=0 drops the gpu-scope
 *   release of the row hand-over between CTAs (csrc/sgm_fused.cu, "link warps"): faster, measured to lose rows when several
 *   pairs are in flight; for measurements only. */
enum { MCCNN_SGM_EXACT = 0, MCCNN_SGM_FUSED = 1 };

typedef struct {
    float P1;       /* 2.3   process_functional.py:1141 (stored as fp32, :149) */
    float P2;       /* 55.9  :1142 */
    float P1_red;   /* fp32(2.3 / 4)   :141 */
    float P2_red;   /* fp32(55.9 / 4)  :142 */
    int threshold;  /* 30    :1143 */
    /* Stages the reference holds but does not run (both 0 = the reference's behaviour; their parity is unpinned): */
    int subpixel;   /* 1: parabola refinement of the WTA index, the formula commented out at :813-819 */
    int bilateral;  /* 1: the 9x9 bilateral filter whose launch is commented out at :1260; as in that launch it reads
                       the filled map and overwrites the median output */
    /* Cross-based cost aggregation between the cost volume and SGM: the stage north_star names and the reference
     * only has a timing slot for (match.py:98, detail_time[2]). 0 iterations = the reference's behaviour. */
    int cbca_iters; /* aggregation passes over both volumes (0 = off; MC-CNN uses 2) */
    int cbca_L1;    /* maximum arm length, 1..32 (default 14) */
    int cbca_tau;   /* arm stops where |I(q) - I(p)| >= tau on u8 grey levels (default 6) */
} mccnn_sgm_params;

const char* mccnn_last_error(void);
int mccnn_abi_version(void);
/* Fills p with the reference's constants (process_functional.py:1141-1144). */
void mccnn_default_sgm_params(mccnn_sgm_params* p);
/* Pitch (in floats) of the disparity axis of a volume with D disparities. */
int mccnn_disp_pitch(int D);
/* 1 if the library was built with kernels for the device's architecture (sm_100a). */
int mccnn_device_supported(int device);

/* ---- pre-processing ------------------------------------------------------------------------
 * Replaces match_single.py:34-43 (standardise with population std) + process_functional.py:13-19
 * (zero-pad by `pad` = (patch-1)/2 on every side). out is fp32 [(H+2*pad)][(W+2*pad)].
 * scratch: 4 doubles. */
int mccnn_standardize_pad(const uint8_t* image, float* out_padded, double* scratch4,
                          int H, int W, int pad, void* stream);
/* Same padding for an already standardised fp32 [H][W] image (the compute_feature signature). */
int mccnn_pad_f32(const float* image, float* out_padded, int H, int W, int pad, void* stream);

/* ---- conv tower ----------------------------------------------------------------------------
 * Replaces Net.construct (mc_cnn_brunch.py:31-48) as run by compute_feature
 * (process_functional.py:21-45): num_layers 3x3 VALID convolutions, bias, ReLU on all but the
 * last, then l2_normalize over the 64 channels.
 *  padded   : fp32 [(H+2*num_layers)][(W+2*num_layers)] (one channel)
 *  weights  : packed by mccnn_pack_weights_host from the reference's HWIO tensors
 *  features : fp32 [H][W][64]
 *  workspace: mccnn_conv_workspace_bytes(H, W, num_layers) bytes. */
size_t mccnn_conv_packed_weight_bytes(int num_layers);
/* hwio_host[i] -> conv{i+1}/weights [3][3][Cin][64], bias_host[i] -> conv{i+1}/biases [64]
 * (mc_cnn_brunch.py:76-77); packed_host is host memory to be copied to the device verbatim. */
int mccnn_pack_weights_host(const float* const* hwio_host, const float* const* bias_host,
                            int num_layers, void* packed_host);
size_t mccnn_conv_workspace_bytes(int H, int W, int num_layers);
int mccnn_conv_tower(const float* padded, const void* packed_weights, float* features,
                     void* workspace, size_t workspace_bytes, int H, int W, int num_layers, void* stream);
/* Same contract on the CUDA cores in plain fp32 (exact fp32 products, fp32 accumulation): the numerically
 * trusted twin of the tensor-core path above (which splits every operand into two fp16 numbers and
 * accumulates in fp32 in TMEM); used to cross-check it on the device. */
int mccnn_conv_tower_fp32(const float* padded, const void* packed_weights, float* features,
                          void* workspace, size_t workspace_bytes, int H, int W, int num_layers, void* stream);

/* ---- cost volume ---------------------------------------------------------------------------
 * Replaces compute_cost_volume_kernel (process_functional.py:120-131) and the host np.ones
 * fill (:1111-1114): CL[y][x][d] = CR[y][x-d][d] = fp32(-sum_i fp32(fl*fr)) with an fp64
 * accumulator, entries never written = fill (1.0 in the reference). CR may be NULL. */
int mccnn_cost_volume(const float* fl, const float* fr, float* CL, float* CR,
                      int H, int W, int D, float fill, void* stream);
/* MCCNN_SGM_FUSED's cost volume: the same layout, fills and pads, fp32 FMA accumulation instead of the reference's fp64
 * accumulator (|difference| <= 4e-6 on unit-norm features; not bit-identical to the reference). */
int mccnn_cost_volume_fast(const float* fl, const float* fr, float* CL, float* CR,
                           int H, int W, int D, float fill, void* stream);
/* The same on the tensor cores: features split into two fp16 numbers each, hi.hi + hi.lo + lo.hi accumulated in fp32 by
 * tcgen05.mma (csrc/cost_volume_fast_tc.cu); |difference to the exact volume| ~1e-6. workspace:
 * mccnn_cost_volume_fast_tc_workspace_bytes(H, W) bytes, 256-byte aligned. */
size_t mccnn_cost_volume_fast_tc_workspace_bytes(int H, int W);
int mccnn_cost_volume_fast_tc(const float* fl, const float* fr, float* CL, float* CR, void* workspace, size_t workspace_bytes,
                              int H, int W, int D, float fill, void* stream);
/* Tensor-core variant, same contract and the same bits out: the exact sum of f*g is taken from tcgen05 MMAs on 8-bit
 * slices of the features, the fp32 rounding residuals of the reference's products from the CUDA cores, and every
 * evaluation whose rounding cannot be proven is redone with the literal loop (csrc/cost_volume_tc.cu).
 * workspace: mccnn_cost_volume_tc_workspace_bytes(H, W) bytes, 256-byte aligned. */
size_t mccnn_cost_volume_tc_workspace_bytes(int H, int W);
int mccnn_cost_volume_tc(const float* fl, const float* fr, float* CL, float* CR, void* workspace, size_t workspace_bytes,
                         int H, int W, int D, float fill, void* stream);

/* ---- MC-CNN-accurate decision head (BASELINE.json config 3; north_star names the net) ------------
 * The reference holds only the layer helper fc() (xw_plus_b + ReLU, weights [num_in][num_out], mc_cnn_brunch.py:95-106)
 * and never builds the head: parity is UNPINNED, the architecture follows the MC-CNN paper (3 hidden layers of
 * MCCNN_FC_UNITS, one sigmoid output) and results are checked against oracle/fc_head.py.
 *   h1 = relu([fl(y,x) ; fr(y,x-d)] W1 + b1), h2 = relu(h1 W2 + b2), h3 = relu(h2 W3 + b3),
 *   CL[y][x][d] = CR[y][x-d][d] = -sigmoid(h3 . w4 + b4); entries never written = fill, pads +INF (as mccnn_cost_volume).
 * fc2 / fc3 run on tcgen05 with fp16 operands and fp32 accumulation (csrc/fc_head.cu); |error| of a cost is a few 1e-4.
 * All pointers are device pointers:
 *   w1_left / w1_right: fp32 [64][384], rows 0..63 / 64..127 of fc1/weights (the layer is split per image)
 *   w2_blocks_f16 / w3_blocks_f16: fc2/weights, fc3/weights ([384 in][384 out]) packed by mccnn_pack_fc_matrix_host into the
 *                       12 pre-swizzled [192 out][64 in] fp16 blocks the kernel streams (mccnn_fc_matrix_blocks_bytes() bytes)
 *   b1, b2, b3, w4    : fp32 [384]; b4: the scalar fc4 bias
 * workspace: mccnn_fc_head_workspace_bytes(H, W) bytes, 256-byte aligned. */
#define MCCNN_FC_UNITS 384
typedef struct {
    const float* w1_left;
    const float* w1_right;
    const float* b1;
    const void* w2_blocks_f16;
    const float* b2;
    const void* w3_blocks_f16;
    const float* b3;
    const float* w4;
    float b4;
} mccnn_fc_weights;
size_t mccnn_fc_matrix_blocks_bytes(void);
int mccnn_pack_fc_matrix_host(const float* w_in_out_host, void* blocks_f16_host);
size_t mccnn_fc_head_workspace_bytes(int H, int W);
int mccnn_cost_volume_accurate(const float* fl, const float* fr, const mccnn_fc_weights* weights, float* CL, float* CR,
                               void* workspace, size_t workspace_bytes, int H, int W, int D, float fill, void* stream);
/* [H][W][Dp] -> dense [D][H][W] (layout of the reference's CPU compute_cost_volume, :48-73). */
int mccnn_volume_to_dhw(const float* vol, float* out_dhw, int H, int W, int D, void* stream);

/* ---- cross-based cost aggregation (north_star kernel 3; no reference implementation: parity unpinned) -----
 * The reference passes the raw volume into SGM under the name d_cost_volumel_after_aggr (process_functional.py:347,
 * :1166) and keeps an unused timing slot for the stage (match.py:98); the definition is the MC-CNN paper's (see
 * csrc/cbca.cu). arms4: u8 [H][W][4] = distance to the first excluded pixel to the left, right, up, down. */
int mccnn_cross_arms(const uint8_t* image, uint8_t* arms4, int H, int W, int L1, int tau, void* stream);
/* One aggregation pass: vol_out[y][x][d] = mean of vol_in over the cross-based support of (y, x, d), taken as the
 * intersection of the supports in the volume's own image (arms_self) and in the other image at x + direction*d
 * (arms_other); direction = -1 for the left volume, +1 for the right one. tmp is a third volume of the same size
 * (row sums). Pad entries [D, Dp) of vol_out are +INF. */
int mccnn_cbca(const float* vol_in, float* vol_out, float* tmp, const uint8_t* arms_self, const uint8_t* arms_other,
               int H, int W, int D, int direction, int L1, void* stream);

/* ---- semi-global matching ------------------------------------------------------------------
 * Replaces sgm_penelty_kernel (:134-262, penalties are recomputed from the images on the fly and
 * never stored), SGM_Interation (:265-343), the 8 path kernels (:346-797) launched at :1166-1202
 * and, fused into the last pass, WTA_and_SupixelRefinement_kernel (:800-837).
 *  CL, CR   : cost volumes (read only)
 *  SL, SR   : aggregated volumes (written; need no initialisation). With keep_volumes == 0 the
 *             final contents are unspecified (the last pass does not store S).
 *  dispL/R  : fp32 [H][W] raw winner-takes-all maps
 *  workspace: mccnn_sgm_workspace_bytes(H, W, D) bytes, 256-byte aligned (one size serves both modes). */
size_t mccnn_sgm_workspace_bytes(int H, int W, int D);
int mccnn_sgm(const float* CL, const float* CR, const uint8_t* imageL, const uint8_t* imageR,
              float* SL, float* SR, float* dispL, float* dispR,
              void* workspace, size_t workspace_bytes,
              int H, int W, int D, const mccnn_sgm_params* params, int mode, int keep_volumes, void* stream);
/* ---- one pair split over several GPUs of a node ------------------------------
 * Row-band sharding: rank r owns image rows [row0, row0 + rows) and holds only those rows of the cost / S volumes
 * and of the output maps ([rows][W][Dp] / [rows][W]); the u8 images are whole on every rank (penalties). Horizontal
 * paths are band-local. A vertical or diagonal scanline that crosses a band boundary is resumed by the next rank from
 * the fp64 path state the previous rank wrote into that rank's exchange buffer over NVLink peer memory (state + flag,
 * inside the scan kernel: no collective, no host round trip); the result is bit-identical to the unsharded run.
 *  xchg_local: this rank's exchange buffer, mccnn_sgm_shard_exchange_bytes(W) bytes, peer-mapped on its neighbours
 *  xchg_prev / xchg_next: peer-mapped device pointers to the neighbours' buffers (NULL at the ends)
 *  epoch: non-zero, the same on every rank, different for every pair (flags are compared with it, never reset);
 *         ranks must not start pair k+1 before their neighbours have finished pair k (one barrier per pair)
 *  pass_mask: bit p launches pass p of 7 (0x7f = all); partial masks let a test run several bands on one GPU in
 *         dependency order. */
typedef struct {
    int rank, world;
    int H_full;
    int row0, rows;
    void* xchg_local;
    void* xchg_prev;
    void* xchg_next;
    unsigned epoch;
    /* Robustness (a rank that dies, fails its argument check or never launches must not wedge the other GPUs):
     *  go_flag   : NULL, or a device int every rank agrees on before its launch (e.g. an all-reduced MIN of "my arguments are
     *              fine and I will launch"); the scan kernels return at once when it is 0;
     *  timeout_ms: a scanline waits at most this long for its neighbour's hand-over (0 = 2000 ms), then the launch gives
     *              up everywhere on this rank and mccnn_sgm_shard_status reports it. */
    const int* go_flag;
    unsigned timeout_ms;
} mccnn_shard;
size_t mccnn_sgm_shard_exchange_bytes(int W);
int mccnn_sgm_sharded(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR,
                      float* SLb, float* SRb, float* dispLb, float* dispRb, void* workspace, size_t workspace_bytes,
                      int W, int D, const mccnn_sgm_params* params, int mode, int keep_volumes,
                      const mccnn_shard* shard, int pass_mask, void* stream);

/* The same split for the opt-in fused mode (MCCNN_SGM_FUSED, csrc/sgm_fused.cu): band-local volumes, whole images, neighbours'
 * exchange buffers in peer memory. Sweeps 1 and 2 walk image rows, and a row needs only the PREVIOUS step of its neighbour row,
 * so the bands of all ranks advance in lock step: the last row of a band streams one state per step into the next rank's buffer
 * (bulk copies over NVLink, epoch-tagged counter), no pipeline fill. Sweep 0 (columns) and sweep 3 (diagonal scanlines) run
 * through the ranks one after the other, their states handed over per column / scanline as in mccnn_sgm_sharded. The result
 * equals the unsharded fused mode value for value.
 *  exchange buffers: mccnn_sgm_fused_shard_exchange_bytes(W, D) bytes each; epoch, go_flag, timeout_ms as above
 *  sweep_mask: bit s launches sweep s of 4 (15 = all); partial masks let a test run several bands on one GPU in dependency
 *              order (sweeps 0, 1: bands top-down; sweeps 2, 3: bottom-up). */
size_t mccnn_sgm_fused_shard_exchange_bytes(int W, int D);
int mccnn_sgm_fused_sharded(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR,
                            float* SLb, float* SRb, float* dispLb, float* dispRb, void* workspace, size_t workspace_bytes,
                            int W, int D, const mccnn_sgm_params* params, int keep_volumes,
                            const mccnn_shard* shard, int sweep_mask, void* stream);

/* After mccnn_sgm_sharded: synchronises the stream and returns the launch's status word in *status_host
 * (0 = every scanline got its hand-over; 1 = a wait hit the deadline, the outputs are invalid). */
int mccnn_sgm_shard_status(const void* workspace, int* status_host, void* stream);

/* One path kernel on one volume, S += path (launch-for-launch twin of :1166-1202; for tests).
 * path: 0..7 in the reference's launch order. */
int mccnn_sgm_single_path(const float* C, const uint8_t* image, float* S, void* workspace, size_t workspace_bytes,
                          int H, int W, int D, const mccnn_sgm_params* params, int path, void* stream);
/* Stand-alone WTA (:800-837): first strict minimum over d, stored as fp32. */
int mccnn_wta(const float* S, float* disp, int H, int W, int D, void* stream);
/* WTA over a dense [D][H][W] volume (CPU WTA1, :96-113). */
int mccnn_wta_dhw(const float* vol_dhw, float* disp, int H, int W, int D, void* stream);

/* ---- left-right check, fill, filters -------------------------------------------------------
 * is_error_match_kernel (:977-1000). flagR may be NULL. */
int mccnn_lr_flags(const float* dispL, const float* dispR, uint8_t* flagL, uint8_t* flagR,
                   int H, int W, void* stream);
/* LRC_kernel (:1003-1088), left map only (the reference never writes the right output).
 * workspace: mccnn_lrc_fill_workspace_bytes(H, W) bytes (two fp32 maps). filled must not alias dispL. */
size_t mccnn_lrc_fill_workspace_bytes(int H, int W);
int mccnn_lrc_fill(const float* dispL, const uint8_t* flagL, float* filled, void* workspace, size_t workspace_bytes,
                   int H, int W, void* stream);
/* Median_Filter_kernel (:840-879) launched as at :1250: interior <- 5x5 median of `filled`,
 * 2-pixel border <- `wta` (the raw map). out may alias wta. */
int mccnn_median5(const float* filled, const float* wta, float* out, int H, int W, void* stream);
/* Bilateral_Filter_kernel (:882-974); its launch is commented out in the reference (:1260). */
int mccnn_bilateral9(const uint8_t* image, const float* disp, float* out, int H, int W, void* stream);
/* astype('uint8') [* scale] of match_single.py:55 / match.py:90 (truncation toward zero, wrap mod 256). */
int mccnn_encode_u8(const float* disp, uint8_t* out, int H, int W, int scale, void* stream);
/* 16-bit variant for disparity ranges above 255 (the reference's uint8 PNG overflows there): fixed point with
 * `frac_bits` fractional bits, saturating. */
int mccnn_encode_u16(const float* disp, uint16_t* out, int H, int W, int frac_bits, void* stream);
/* Stand-alone WTA with the parabola refinement of :813-819 (own definition where the reference is silent:
 * the index is kept when it sits on the range border or the parabola is flat). */
int mccnn_wta_subpixel(const float* S, float* disp, int H, int W, int D, void* stream);
/* error_calculate.py:68-83 on the device: counts[0] = bad pixels, counts[1] = valid GT pixels.
 * gt_half is the ground truth already resized and halved (fp32 [H][W]). */
int mccnn_bad_pixels(const uint8_t* disp_u8, const float* gt_half, unsigned long long* counts2,
                     int H, int W, void* stream);
/* The same count for the 16-bit maps written when the disparity range exceeds what the reference's uint8 PNG holds
 * (match_single.py:55 wraps there). */
int mccnn_bad_pixels_u16(const uint16_t* disp_u16, const float* gt_half, unsigned long long* counts2,
                         int H, int W, void* stream);

/* ---- whole path ----------------------------------------------------------------------------
 * disparity_compute_by_gpu (:1093-1267) from device-resident inputs: cost volume -> SGM -> WTA ->
 * L-R flags -> fill -> median. Returns the filtered left map and the raw right WTA map (the
 * reference's right output is the median of an uninitialised buffer, so only its border, which
 * equals the raw map, is defined). stage_ms_host (may be NULL) receives 7 floats in the layout of
 * the reference's detail_time (match.py:95-103); filling it synchronises the stream.
 *  workspace: mccnn_pipeline_workspace_bytes(H, W, D) bytes. */
size_t mccnn_pipeline_workspace_bytes(int H, int W, int D);
int mccnn_disparity_pipeline(const uint8_t* imageL, const uint8_t* imageR, const float* fl, const float* fr,
                             float* dispL_out, float* dispR_out, void* workspace, size_t workspace_bytes,
                             int H, int W, int D, const mccnn_sgm_params* params, int mode,
                             float* stage_ms_host, void* stream);

/* match_single.py:34-55 between imread and imwrite, from device u8 images: standardise, pad, conv
 * tower (both images), then mccnn_disparity_pipeline.
 *  workspace: mccnn_match_workspace_bytes(H, W, D, num_layers) bytes. */
size_t mccnn_match_workspace_bytes(int H, int W, int D, int num_layers);
int mccnn_match_pair(const uint8_t* imageL, const uint8_t* imageR, const void* packed_weights,
                     float* dispL_out, float* dispR_out, void* workspace, size_t workspace_bytes,
                     int H, int W, int D, int num_layers, const mccnn_sgm_params* params, int mode,
                     float* stage_ms_host, void* stream);
/* The same path with the MC-CNN-accurate decision head as the matching cost (mccnn_cost_volume_accurate) in place of the
 * dot product; everything after the cost volume is unchanged. stage_ms_host[1] then holds the head's time. */
size_t mccnn_match_accurate_workspace_bytes(int H, int W, int D, int num_layers);
int mccnn_match_pair_accurate(const uint8_t* imageL, const uint8_t* imageR, const void* packed_weights,
                              const mccnn_fc_weights* head, float* dispL_out, float* dispR_out, void* workspace,
                              size_t workspace_bytes, int H, int W, int D, int num_layers, const mccnn_sgm_params* params,
                              int mode, float* stage_ms_host, void* stream);

/* ---- training step of the siamese tower ------------------------------------------------------------
 * Replaces the graph train.py builds (train.py:71-99) and runs once per batch: three weight-sharing branches
 * Net(num_of_conv_layers = L, 64 maps) on [B][p][p] patches with p = 2 L + 1 (mc_cnn_brunch.py:31-48), cosine similarities of
 * the left feature with the positive / negative right feature, loss = mean(max(0, margin - cos_pos + cos_neg)) (:83-89),
 * tf.train.MomentumOptimizer: accum = momentum * accum + grad; var -= lr * accum (:97-99). TensorFlow's arithmetic is not
 * pinned: results are held to a tolerance against oracle/train_step.py (fp64 autograd).
 *  params / velocity / grads_out: flat fp32 vectors of mccnn_train_param_count(L) floats, per layer i the HWIO weights
 *      [3][3][cin][64] (conv{i}/weights) followed by the 64 biases (conv{i}/biases); cin = 1 for the first layer
 *  left / right_pos / right_neg: fp32 [B][p][p] device patches; loss_out: one device float
 *  grads_out may be NULL; apply_update = 0 computes loss and gradients only (params, velocity untouched). */
size_t mccnn_train_param_count(int num_layers);
size_t mccnn_train_workspace_bytes(int batch, int patch, int num_layers);
int mccnn_train_step(const float* left, const float* right_pos, const float* right_neg, float* params, float* velocity,
                     float* grads_out, float* loss_out, void* workspace, size_t workspace_bytes, int batch, int patch,
                     int num_layers, float margin, float lr, float momentum, int apply_update, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCCNN_B200_H */
