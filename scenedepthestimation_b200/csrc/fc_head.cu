// MC-CNN-accurate decision head: the cost of an evaluation is a small fully-connected network on the two feature
// vectors instead of their dot product (BASELINE.json config 3).
//
// The reference holds only the layer helper (fc: xw_plus_b + ReLU, weights [num_in][num_out], mc_cnn_brunch.py:95-106)
// and never builds the head, so PARITY IS UNPINNED: the architecture follows the MC-CNN paper (Zbontar & LeCun, JMLR
// 2016, "accurate" Middlebury net: 3 hidden layers of 384 units, one sigmoid output) on top of the tower this repo
// already runs, and results are checked against this repo's own oracle (oracle/fc_head.py):
//   h1 = relu(W1^T [fl(y,x) ; fr(y,x-d)] + b1)        W1 [128][384]   (fc1)
//   h2 = relu(W2^T h1 + b2)                            W2 [384][384]   (fc2)
//   h3 = relu(W3^T h2 + b3)                            W3 [384][384]   (fc3)
//   CL[y][x][d] = CR[y][x-d][d] = -sigmoid(w4 . h3 + b4)               (fc4), `fill` where x - d < 0, +INF pads
//
// 0.59 MFLOP per evaluation (c3, 1440x994x400: 338 TFLOP): the one tensor-pipe-bound stage of the repo besides the tower.
//   * fc1 is linear in the concatenation, so it is split per image and done once per PIXEL on the CUDA cores:
//     A1 = fl W1[:64] + b1, B1 = fr W1[64:], stored as fp16 [P][384]; per evaluation h1 = relu(A1[x] + B1[x-d]).
//   * fc2 / fc3: tcgen05.mma kind::f16 (fp16 operands, fp32 accumulation in TMEM). A tile is 128 consecutive x of one
//     image row at one disparity: the A1 and B1 rows of the tile are two contiguous [128][384] blocks. A1 comes by TMA
//     in six K chunks of 64 straight into the 128-byte-swizzled K-major layout; the activation warps add the B1 rows
//     (held in registers, loaded one tile ahead) + ReLU in place. W2 / W3 stream through a five-stage ring of
//     [192 n][64 k] blocks (24 KB each, pre-swizzled on the host, one bulk copy from L2 per block), four K = 16 MMAs of N = 192 per block. The 128 x 384 fp32 accumulator
//     (384 TMEM columns) is drained by the activation warps: bias + ReLU + fp16 back into the same shared-memory tile as
//     the A operand of fc3; after fc3: bias + ReLU, dot with w4, sigmoid, store.
//   * Warp roles: 8 activation / epilogue warps (TMEM lane quarter x column half), 1 TMA warp for the A1 chunks, 1 TMA
//     warp for the weight ring, 1 MMA warp. Persistent, one CTA of 222 KB per SM.
// Precision: fp16 operands give |d cost| of a few 1e-4 against the fp32 network (tests: 2e-3 absolute).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace mccnn {
namespace {

constexpr int NF = MCCNN_FEATURES;
constexpr int FC = MCCNN_FC_UNITS;            // 384
constexpr int TM = 128;                        // evaluations (consecutive x) per tile
constexpr int KC = 64;                         // K chunk: 64 fp16 = one 128-byte swizzled row
constexpr int NKC = FC / KC;                   // 6
constexpr int NH = FC / 2;                     // 192: N of one MMA
constexpr int H_CHUNK = TM * 128;              // 16384 bytes: [128 rows][64 fp16]
constexpr int W_HALF = NH * 128;               // 24576 bytes: [192 n][64 fp16], one stage of the weight ring
constexpr int NWS = 4;                         // weight ring stages (96 KB in flight; 3 stages make the weight stream the bound, measured)
constexpr int NH1 = 1;                         // h1 chunks 0..2 have their own buffer (built for the NEXT tile while fc3 of this one runs);
                                               // chunks 3..5 borrow the first three chunks of the h2 tile, idle between fc3 and the fc2 epilogue
constexpr int OFF_H1 = 0;                      // h1 chunks 0..2
constexpr int OFF_H2 = OFF_H1 + NH1 * H_CHUNK; // 49152: the whole h2 tile (6 chunks)
constexpr int OFF_W = OFF_H2 + NKC * H_CHUNK;  // 147456
constexpr int OFF_BIAS = OFF_W + NWS * W_HALF; // 221184: b2, b3, w4 (3 x 384 floats)
constexpr int OFF_Z = OFF_BIAS + 3 * FC * 4;   // 225792: partial fc4 sums of the upper column half (128 floats)
constexpr int OFF_BAR = OFF_Z + TM * 4;        // 226304
constexpr int FC_SMEM = OFF_BAR + 256 + 1024;  // barriers + alignment slack = 227584
constexpr int NAW = 8;                         // activation / epilogue warps: warp w owns TMEM lanes 32 (w & 3).. and column half w >> 2
constexpr int FC_THREADS = 32 * (NAW + 3);
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NH >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);  // f16 x f16 -> f32, M128 N192
constexpr float kInfF = __builtin_huge_valf();

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- fc1, split per image (CUDA cores, fp32 -> fp16)
// out[p][j] = sum_i feat[p][i] * w[i][j] (+ bias[j]); block = 32 pixels x 384 outputs, thread = 3 outputs x 32 pixels.
__global__ void __launch_bounds__(128) fc1_half_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                      const float* __restrict__ bias, __half* __restrict__ out, long long P) {
    __shared__ float fs[32][NF + 1];
    const long long p0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * NF; i += 128) {
        const long long p = p0 + i / NF;
        fs[i / NF][i % NF] = p < P ? feat[p * NF + (i % NF)] : 0.f;
    }
    __syncthreads();
    float acc[3][32];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float b = bias ? __ldg(bias + threadIdx.x + 128 * c) : 0.f;
#pragma unroll
        for (int q = 0; q < 32; q++) acc[c][q] = b;
    }
    for (int i = 0; i < NF; i++) {
        float wv[3];
#pragma unroll
        for (int c = 0; c < 3; c++) wv[c] = __ldg(w + (size_t)i * FC + threadIdx.x + 128 * c);
#pragma unroll
        for (int q = 0; q < 32; q++) {
            const float f = fs[q][i];
#pragma unroll
            for (int c = 0; c < 3; c++) acc[c][q] = fmaf(f, wv[c], acc[c][q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const long long p = p0 + q;
        if (p < P)
#pragma unroll
            for (int c = 0; c < 3; c++) out[p * FC + threadIdx.x + 128 * c] = __float2half_rn(acc[c][q]);
    }
}

// entries no evaluation writes: fill where the match falls outside the other image, +INF pads
__global__ void __launch_bounds__(256) fc_fill_kernel(float* __restrict__ CL, float* __restrict__ CR, int W, int D, int Dp,
                                                     long long P, float fill) {
    const long long pix = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pix >= P) return;
    const int x = (int)(pix % W);
    float* rl = CL + pix * Dp;
    for (int d = x + 1 + lane; d < D; d += 32) rl[d] = fill;  // x - d < 0
    if (Dp > D && lane < Dp - D) rl[D + lane] = kInfF;
    if (CR != nullptr) {
        float* rr = CR + pix * Dp;
        for (int d = max(W - x, 0) + lane; d < D; d += 32) rr[d] = fill;  // x + d >= W
        if (Dp > D && lane < Dp - D) rr[D + lane] = kInfF;
    }
}

struct FcArgs {
    const __half* B1;  // fc1 of the right image, [P][384]
    const unsigned char* w2b;  // fc2 / fc3 weights as 12 pre-swizzled [192 n][64 k] fp16 blocks each (mccnn_pack_fc_matrix_host)
    const unsigned char* w3b;
    const float* b2;
    const float* b3;
    const float* w4;
    float b4;
    float* CL;
    float* CR;
    int H, W, D, Dp;
    int tiles_x;
    long long nitems;  // H * tiles_x * D
};

// work items of one CTA: (y, x block, d), d fastest so that concurrently running CTAs share the B1 rows in L2;
// items whose 128 pixels all have x < d hold no valid evaluation and are skipped by every role
struct ItemCursor {
    long long item, stride, nitems;
    int y, x0, d;
    __device__ void decode(const FcArgs& a) {
        for (; item < nitems; item += stride) {
            const long long per_row = (long long)a.tiles_x * a.D;
            y = (int)(item / per_row);
            const int rem = (int)(item - (long long)y * per_row);
            x0 = (rem / a.D) * TM;
            d = rem % a.D;
            if (x0 + TM - 1 >= d) return;
        }
    }
    __device__ void start(const FcArgs& a, int first, int step) {
        item = first; stride = step; nitems = a.nitems;
        decode(a);
    }
    __device__ bool valid() const { return item < nitems; }
    __device__ void next(const FcArgs& a) {
        item += stride;
        decode(a);
    }
};

// Barriers (one completion per tile unless noted): a_full[6] (TMA bytes of A1 chunk kc), h_full[6] (8 arrivals: chunk kc holds
// relu(A1 + B1)), h1_free[3] (tcgen05.commit: the fc2 MMAs reading h1 chunk kc < 3 are done), l3_done (commit: fc3 no longer
// reads the h2 tile), h2_full (8 arrivals: the fc2 epilogue has written the h2 tile and drained the accumulator), w_full[3]
// (TMA bytes) / w_free[3] (commit), per use of a ring stage; acc_full (commit; twice per tile: fc2, fc3), acc_free (8 arrivals:
// the accumulator has been drained after fc3).
// Timeline of tile t on the activation warps: [h1 chunks 3..5 of t in place in the idle h2 tile, under the first fc2 MMAs] ->
// wait fc2 -> fc2 epilogue (h2 tile) -> [B1 loads and h1 chunks 0..2 of t + 1, under the fc3 MMAs] -> wait fc3 -> fc3 epilogue +
// fc4 -> store. The A1 chunks 3..5 of t + 1 land in the h2 tile as soon as fc3 of t has finished, i.e. under the fc3 epilogue.
__global__ void __launch_bounds__(FC_THREADS, 1)
fc_head_kernel(const __grid_constant__ CUtensorMap tmA, const FcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint64_t* a_full = bars;                  // [NKC]
    uint64_t* h_full = bars + NKC;            // [NKC]
    uint64_t* h1_free = bars + 2 * NKC;       // [NH1]
    uint64_t* w_full = h1_free + NH1;         // [NWS]
    uint64_t* w_free = w_full + NWS;          // [NWS]
    uint64_t* h2_full = w_free + NWS;
    uint64_t* acc_full = h2_full + 1;
    uint64_t* acc_free = h2_full + 2;
    uint64_t* l3_done = h2_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h2_full + 4);
    float* sbias = reinterpret_cast<float*>(sm + OFF_BIAS);  // [0] b2, [1] b3, [2] w4
    float* zbuf = reinterpret_cast<float*>(sm + OFF_Z);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int k = 0; k < NKC; k++) {
            mbar_init(a_full + k, 1);
            mbar_init(h_full + k, NAW);
        }
        for (int k = 0; k < NH1; k++) mbar_init(h1_free + k, 1);
        mbar_init(l3_done, 1);
        for (int s = 0; s < NWS; s++) {
            mbar_init(w_full + s, 1);
            mbar_init(w_free + s, 1);
        }
        mbar_init(h2_full, NAW);
        mbar_init(acc_full, 1);
        mbar_init(acc_free, NAW);
        mbar_fence_init();
    }
    for (int i = tid; i < FC; i += FC_THREADS) {
        sbias[i] = a.b2[i];
        sbias[FC + i] = a.b3[i];
        sbias[2 * FC + i] = a.w4[i];
    }
    if (warp == NAW + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == NAW) {
        // ================================================================= TMA: A1 chunks 0..2 into the h1 buffer (free once fc2 of the previous
        // tile has consumed them), chunks 3..5 into the h2 tile (free once fc3 of the previous tile has finished)
        if (lane == 0) {
            ItemCursor it;
            it.start(a, blockIdx.x, gridDim.x);
            uint32_t t = 0;
            for (; it.valid(); it.next(a), t++) {
                const long long prow = (long long)it.y * a.W;
                for (int kc = 0; kc < NKC; kc++) {
                    if (t > 0 && kc < NH1) mbar_wait(h1_free + kc, (t - 1) & 1u);
                    if (t > 0 && kc == NH1) mbar_wait(l3_done, (t - 1) & 1u);
                    mbar_expect_tx(a_full + kc, H_CHUNK);
                    const uint32_t dst = kc < NH1 ? OFF_H1 + kc * H_CHUNK : OFF_H2 + (kc - NH1) * H_CHUNK;
                    tma_load_2d(base + dst, &tmA, kc * KC, (int)(prow + it.x0), a_full + kc);
                }
            }
        }
    } else if (warp == NAW + 1) {
        // ================================================================= TMA: weight ring; per tile W2 then W3, each as (kc, N half) blocks
        if (lane == 0) {
            ItemCursor it;
            it.start(a, blockIdx.x, gridDim.x);
            uint32_t g = 0;
            for (; it.valid(); it.next(a)) {
                for (int layer = 0; layer < 2; layer++)
                    for (int kc = 0; kc < NKC; kc++)
                        for (int nh = 0; nh < 2; nh++, g++) {
                            const uint32_t st = g % NWS, use = g / NWS;
                            if (use > 0) mbar_wait(w_free + st, (use - 1) & 1u);
                            mbar_expect_tx(w_full + st, W_HALF);
                            // one contiguous 24 KB block, already in the 128-byte-swizzled layout the MMA reads
                            bulk_g2s(sm + OFF_W + st * W_HALF, (layer ? a.w3b : a.w2b) + (size_t)(kc * 2 + nh) * W_HALF, W_HALF, w_full + st);
                        }
            }
        }
    } else if (warp == NAW + 2) {
        // ================================================================= MMA issue
        if (lane == 0) {
            ItemCursor it;
            it.start(a, blockIdx.x, gridDim.x);
            uint32_t t = 0, g = 0;
            for (; it.valid(); it.next(a), t++) {
                if (t > 0) mbar_wait(acc_free, (t - 1) & 1u);  // the previous tile's accumulator has been drained
                for (int layer = 0; layer < 2; layer++) {
                    if (layer) mbar_wait(h2_full, t & 1u);
                    for (int kc = 0; kc < NKC; kc++) {
                        uint32_t a_addr;
                        if (layer == 0) {
                            mbar_wait(h_full + kc, t & 1u);
                            a_addr = base + (kc < NH1 ? OFF_H1 + kc * H_CHUNK : OFF_H2 + (kc - NH1) * H_CHUNK);
                        } else {
                            a_addr = base + OFF_H2 + kc * H_CHUNK;
                        }
                        const uint64_t ad = sw128_desc(a_addr);
                        for (int nh = 0; nh < 2; nh++, g++) {
                            const uint32_t st = g % NWS, use = g / NWS;
                            mbar_wait(w_full + st, use & 1u);
                            tc_fence_after();
                            const uint64_t bd = sw128_desc(base + OFF_W + st * W_HALF);
#pragma unroll
                            for (int k = 0; k < KC / 16; k++) umma_f16(tmem_base + nh * NH, ad + 2 * k, bd + 2 * k, (kc | k) != 0 ? 1u : 0u);
                            umma_commit(w_free + st);
                        }
                        if (layer == 0 && kc < NH1) umma_commit(h1_free + kc);
                    }
                    umma_commit(acc_full);
                }
                umma_commit(l3_done);
            }
        }
    } else {
        // ================================================================= activation / epilogue warps
        // epilogues: thread = tile row r (TMEM lane) x column half (accumulator columns 192 half..)
        const int q = warp & 3, half = warp >> 2;
        const int r = 32 * q + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(NH * half);
        const uint32_t row_off = (uint32_t)r * 128u, rsw = (uint32_t)(r & 7);
        const long long P = (long long)a.H * a.W;
        // The B1 rows of a tile (one contiguous 96 KB block of global memory) are added from registers, loaded one tile ahead.
        // For this step a thread is not tied to its TMEM row: per K chunk it takes the 16-byte unit bu of rows br + 32 i, so that
        // 8 lanes read one full 128-byte line (a row-per-thread mapping costs 32 lines per load instruction and 4.8k of the
        // 26k cycles of a tile, measured).
        const int bu = tid & 7, br = tid >> 3;
        uint4 breg[NKC][4];
        auto load_b1 = [&](const ItemCursor& c) {
            const long long row0 = (long long)c.y * a.W + c.x0 - c.d + br;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const long long row = row0 + 32 * i;
                const bool ok = row >= 0 && row < P;  // outside: the evaluation is invalid anyway
                const uint4* src = reinterpret_cast<const uint4*>(a.B1 + (ok ? row : 0) * FC) + bu;
#pragma unroll
                for (int kc = 0; kc < NKC; kc++) breg[kc][i] = ok ? __ldg(src + kc * 8) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        // h1 chunk kc of tile number tt = relu(A1 + B1), in place where the TMA put the A1 chunk
        auto build_h1 = [&](int kc, uint32_t tt) {
            mbar_wait(a_full + kc, tt & 1u);
            unsigned char* hq = sm + (kc < NH1 ? OFF_H1 + kc * H_CHUNK : OFF_H2 + (kc - NH1) * H_CHUNK) + br * 128 +
                                ((uint32_t)(bu ^ (br & 7)) << 4);  // + 32 i rows: same swizzle phase
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4* hp = reinterpret_cast<uint4*>(hq + i * 32 * 128);
                uint4 x = *hp;
                const __half2 z = __float2half2_rn(0.f);
                __half2* xh = reinterpret_cast<__half2*>(&x);
                const __half2* yh = reinterpret_cast<const __half2*>(&breg[kc][i]);
#pragma unroll
                for (int e = 0; e < 4; e++) xh[e] = __hmax2(__hadd2(xh[e], yh[e]), z);
                *hp = x;
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(h_full + kc);
        };
        ItemCursor it;
        it.start(a, blockIdx.x, gridDim.x);
        if (it.valid()) {
            load_b1(it);
#pragma unroll
            for (int kc = 0; kc < NH1; kc++) build_h1(kc, 0u);
        }
        uint32_t t = 0;
        for (; it.valid(); t++) {
            // ---- h1 chunks 3..5 of this tile, in the idle h2 tile (their A1 rows arrived under the previous fc3 epilogue)
#pragma unroll
            for (int kc = NH1; kc < NKC; kc++) build_h1(kc, t);
            const int x = it.x0 + r, d = it.d;
            const long long prow = (long long)it.y * a.W;
            it.next(a);
            // ---- fc2 epilogue: h2 = relu(acc + b2) as fp16 into the h2 tile (fc3 of the previous tile has long finished)
            mbar_wait(acc_full, 0);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < NH / 16; c += 2) {
                float v[2][16];
                tmem_ld16(taddr + 16u * c, v[0]);
                tmem_ld16(taddr + 16u * (c + 1), v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const int col = NH * half + 16 * (c + p);
                    __align__(16) __half2 hv[8];
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        hv[e] = __floats2half2_rn(fmaxf(v[p][2 * e] + sbias[col + 2 * e], 0.f), fmaxf(v[p][2 * e + 1] + sbias[col + 2 * e + 1], 0.f));
                    // columns col..col+15 = K chunk col / 64, 16-byte units (col % 64) / 8 and the next one of row r
                    unsigned char* hrow = sm + OFF_H2 + (col >> 6) * H_CHUNK + row_off;
                    const uint32_t u0 = (uint32_t)(col & 63) >> 3;
                    *reinterpret_cast<uint4*>(hrow + ((u0 ^ rsw) << 4)) = *reinterpret_cast<const uint4*>(&hv[0]);
                    *reinterpret_cast<uint4*>(hrow + (((u0 + 1) ^ rsw) << 4)) = *reinterpret_cast<const uint4*>(&hv[4]);
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(h2_full);  // also: the accumulator may be overwritten by fc3
            // ---- under the fc3 MMAs: the B1 rows and the first h1 chunks of the next tile
            if (it.valid()) {
                load_b1(it);
#pragma unroll
                for (int kc = 0; kc < NH1; kc++) build_h1(kc, t + 1);
            }
            // ---- fc3 epilogue + fc4: z = w4 . relu(acc + b3) + b4, cost = -sigmoid(z)
            mbar_wait(acc_full, 1);
            tc_fence_after();
            float z = 0.f;
#pragma unroll 1
            for (int c = 0; c < NH / 16; c += 2) {
                float v[2][16];
                tmem_ld16(taddr + 16u * c, v[0]);
                tmem_ld16(taddr + 16u * (c + 1), v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const int col = NH * half + 16 * (c + p);
#pragma unroll
                    for (int e = 0; e < 16; e++) z = fmaf(fmaxf(v[p][e] + sbias[FC + col + e], 0.f), sbias[2 * FC + col + e], z);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_free);
            if (half) zbuf[r] = z;
            asm volatile("bar.sync 1, %0;" ::"n"(32 * NAW) : "memory");
            if (!half && x < a.W && x >= d) {
                const float cost = -1.0f / (1.0f + expf(-(z + zbuf[r] + a.b4)));
                a.CL[(prow + x) * a.Dp + d] = cost;
                if (a.CR != nullptr) a.CR[(prow + x - d) * a.Dp + d] = cost;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NAW + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 [rows][FC] row-major, box = 64 columns x box_rows rows, 128-byte swizzle
int make_map(CUtensorMap* tm, const void* ptr, size_t rows, int box_rows) {
    EncodeTiledFn enc = get_encode();
    MCCNN_REQUIRE(enc != nullptr, MCCNN_EINVAL, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {(cuuint64_t)FC, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)FC * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MCCNN_REQUIRE(r == CUDA_SUCCESS, MCCNN_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_fc_matrix_blocks_bytes(void) { return (size_t)FC * FC * sizeof(__half); }

// fc2 / fc3 weights [384 in][384 out] fp32 (the reference's fc() layout, mc_cnn_brunch.py:97) -> the 12 blocks the kernel
// streams: block (kc, nh) = outputs 192 nh.., inputs 64 kc.., as [192 n][64 k] fp16 K-major rows of 128 bytes whose 16-byte
// units sit at position u ^ (n & 7) (the 128-byte swizzle tcgen05 reads), blocks in (kc, nh) order.
extern "C" int mccnn_pack_fc_matrix_host(const float* w_in_out_host, void* blocks_f16_host) {
    MCCNN_REQUIRE(w_in_out_host && blocks_f16_host, MCCNN_EINVAL, "mccnn_pack_fc_matrix_host: null argument");
    __half* out = reinterpret_cast<__half*>(blocks_f16_host);
    for (int kc = 0; kc < NKC; kc++)
        for (int nh = 0; nh < 2; nh++) {
            __half* blk = out + (size_t)(kc * 2 + nh) * (W_HALF / sizeof(__half));
            for (int n = 0; n < NH; n++)
                for (int u = 0; u < 8; u++)
                    for (int e = 0; e < 8; e++) {
                        const int k = kc * KC + 8 * u + e, o = nh * NH + n;
                        blk[(size_t)n * 64 + ((u ^ (n & 7)) << 3) + e] = __float2half_rn(w_in_out_host[(size_t)k * FC + o]);
                    }
        }
    return 0;
}

extern "C" size_t mccnn_fc_head_workspace_bytes(int H, int W) {
    if (H < 1 || W < 1) return 0;
    return 2 * a256((size_t)H * W * FC * sizeof(__half));
}

extern "C" int mccnn_cost_volume_accurate(const float* fl, const float* fr, const mccnn_fc_weights* w, float* CL, float* CR,
                                          void* workspace, size_t workspace_bytes, int H, int W, int D, float fill,
                                          void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(fl && fr && w && CL && workspace, MCCNN_EINVAL, "mccnn_cost_volume_accurate: null argument");
    MCCNN_REQUIRE(w->w1_left && w->w1_right && w->b1 && w->w2_blocks_f16 && w->b2 && w->w3_blocks_f16 && w->b3 && w->w4, MCCNN_EINVAL,
                  "mccnn_cost_volume_accurate: null weight pointer");
    MCCNN_REQUIRE(H >= 1 && W >= 1 && D >= 1 && D <= 4096, MCCNN_EINVAL, "mccnn_cost_volume_accurate: bad shape H=%d W=%d D=%d", H, W, D);
    MCCNN_REQUIRE((long long)H * W + TM < 0x7fffffffLL, MCCNN_EINVAL, "mccnn_cost_volume_accurate: image too large for 32-bit tile rows");
    MCCNN_REQUIRE(aligned16(fl) && aligned16(fr) && aligned16(w->w2_blocks_f16) && aligned16(w->w3_blocks_f16) && aligned16(w->b2) &&
                      aligned16(w->b3) && aligned16(w->w4) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                  MCCNN_EALIGN, "mccnn_cost_volume_accurate: features / weights must be 16-byte, the workspace 256-byte aligned");
    MCCNN_REQUIRE(workspace_bytes >= mccnn_fc_head_workspace_bytes(H, W), MCCNN_EWORKSPACE,
                  "mccnn_cost_volume_accurate: workspace too small");
    const long long P = (long long)H * W;
    const int Dp = disp_pitch(D);
    char* ws = reinterpret_cast<char*>(workspace);
    __half* A1 = reinterpret_cast<__half*>(ws);
    __half* B1 = reinterpret_cast<__half*>(ws + a256((size_t)P * FC * sizeof(__half)));

    const unsigned nb = (unsigned)((P + 31) / 32);
    fc1_half_kernel<<<nb, 128, 0, stream>>>(fl, w->w1_left, w->b1, A1, P);
    MCCNN_LAUNCH_CHECK("fc1_half_kernel");
    fc1_half_kernel<<<nb, 128, 0, stream>>>(fr, w->w1_right, nullptr, B1, P);
    MCCNN_LAUNCH_CHECK("fc1_half_kernel");
    fc_fill_kernel<<<(unsigned)((P + 7) / 8), 256, 0, stream>>>(CL, CR, W, D, Dp, P, fill);
    MCCNN_LAUNCH_CHECK("fc_fill_kernel");

    CUtensorMap tmA;
    if (int e = make_map(&tmA, A1, (size_t)P, TM)) return e;
    FcArgs a{};
    a.B1 = B1;
    a.w2b = reinterpret_cast<const unsigned char*>(w->w2_blocks_f16);
    a.w3b = reinterpret_cast<const unsigned char*>(w->w3_blocks_f16);
    a.b2 = w->b2; a.b3 = w->b3; a.w4 = w->w4; a.b4 = w->b4;
    a.CL = CL; a.CR = CR;
    a.H = H; a.W = W; a.D = D; a.Dp = Dp;
    a.tiles_x = ceil_div(W, TM);
    a.nitems = (long long)H * a.tiles_x * D;
    if (int e = kernel_setup<fc_head_kernel>(FC_THREADS, FC_SMEM, nullptr)) return e;
    long long grid = sm_count();
    if (grid > a.nitems) grid = a.nitems;
    fc_head_kernel<<<(unsigned)grid, FC_THREADS, FC_SMEM, stream>>>(tmA, a);
    MCCNN_LAUNCH_CHECK("fc_head_kernel");
    return 0;
}
