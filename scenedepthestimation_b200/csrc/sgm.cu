// Semi-global matching, exact-arithmetic mode (sm_100a).
//
// Replaces, for any disparity count D <= 1024:
//   sgm_penelty_kernel        process_functional.py:134-262  (penalties recomputed on the fly)
//   SGM_Interation            process_functional.py:265-343
//   the 8 SGM_*_kernel paths  process_functional.py:346-797, launched at :1166-1202
//   WTA_and_SupixelRefinement process_functional.py:800-837  (fused into the last pass)
//
// Arithmetic contract (pinned by tests/golden/ref_*.npz):
//   * path state L, min and min+P2 are fp64; P1/P2 are fp32 constants widened on use;
//   * c[d] = fp64(C[d]) + (min(L[d-1]+P1, L[d], L[d+1]+P1, minL+P2) - minL), with the raw cost on the
//     first pixel of a scanline and after a diagonal column wrap;
//   * S = fp32(fp64(S) + c) once per path, paths in the launch order
//     down, up, right, left, down-right, up-right, down-left, up-left;
//   * a scanline visits N-1 of its N pixels; the "up" path has P1 = P2 = 0 (the reference never
//     writes penalty channels 0/1) and therefore adds the raw cost: it is fused into the first pass.
//   * penalty pair for a step prev -> cur: (P1, P2) iff 0 <= I[cur] - I[prev] <= threshold, else the
//     reduced pair (uint8 differences wrap to uint64 in the reference, App. A3).
//
// Kernel shape: one warp owns one scanline. Lane l holds NPL consecutive disparities of the fp64
// state in registers, so the d-1 / d+1 neighbours need two shuffles per step, and the minimum over
// all disparities is two integer REDUX operations on an order-preserving key. Cost / S rows
// (D * 4 bytes, contiguous) stream through shared memory with 1-D bulk async copies
// (cp.async.bulk + mbarrier, a ring of STAGES rows per warp) and the updated S row goes back with a
// bulk store. Warps pull scanlines from a global counter (persistent grid, multiple of the SM count).
#include "common.cuh"

namespace mccnn {
// MCCNN_SGM_FUSED: csrc/sgm_fused.cu
size_t sgm_fused_workspace_bytes(int H, int W, int D);
int run_sgm_fused(const float* CL, const float* CR, const uint8_t* imageL, const uint8_t* imageR, float* SL, float* SR, float* dispL,
                  float* dispR, void* workspace, int H, int W, int D, const mccnn_sgm_params* p, int keep_volumes, cudaStream_t stream);
size_t sgm_fused_xchg_bytes(int W, int D);
int run_sgm_fused_band(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR, float* SLb, float* SRb,
                       float* dispLb, float* dispRb, void* workspace, int W, int D, const mccnn_sgm_params* p, int keep_volumes,
                       const mccnn_shard* sh, int sweep_mask, cudaStream_t stream);
namespace {

enum SgmMode { SGM_MID = 0, SGM_FIRST_FUSED = 1, SGM_LAST_WTA = 2 };

struct SgmArgs {
    const float* C[2];
    float* S[2];
    const unsigned char* img[2];
    float* disp[2];
    int H, W, D, Dp;
    int dy, dx, horizontal;
    int nlines;        // scanlines per side
    int nsteps_dp;     // pixels visited by the dynamic program
    int nsteps_total;  // >= nsteps_dp: FIRST_FUSED / LAST_WTA also touch the pixel the path skips
    int nsides;
    int side0;         // first side handled by this launch (single-side launches of the sharded schedule)
    double P1, P2, P1r, P2r;
    int threshold;
    int subpixel;
    int store_s;
    unsigned* counter;
    // row-band sharding (one pair split over several GPUs): this launch owns image rows [row0, row0 + Hb); the
    // volumes / maps it is given hold only those rows, the u8 images are whole. A scanline that enters the band
    // from another rank resumes from the fp64 path state that rank left in hand_in (local memory, written by
    // the peer over NVLink) once its flag shows `epoch`; a scanline that leaves the band publishes its state
    // to hand_out (PEER memory). Unsharded: row0 = 0, Hb = H, both null.
    int row0, Hb;
    const double* hand_in;
    const unsigned* flag_in;
    double* hand_out;
    unsigned* flag_out;
    unsigned epoch;
    // sharded runs only: `go` (may be null) is a device word every rank agrees on before the launch (an all-reduced
    // "all ranks validated their arguments and will launch"); 0 = return at once. A scanline waits for its hand-over at
    // most `timeout_ns`; a warp that gives up sets *status (the workspace word the host reads back) and every warp that
    // sees it set leaves too, so a rank that died or never launched cannot wedge the other GPUs.
    const int* go;
    unsigned* status;
    unsigned long long timeout_ns;
};

constexpr int STATUS_WORD = 32;  // index of the status word inside the 64-word SGM workspace (counters use 0..13)

constexpr int HAND_STRIDE = 1024 + 8;  // doubles per (side, scanline) slot: state of up to 1024 disparities + min

constexpr int WARPS_PER_CTA = 4;

__device__ __forceinline__ void scan_pixel(const SgmArgs& a, int line, int t, int& row, int& col) {
    if (a.horizontal) {
        row = a.row0 + line;
        col = a.dx > 0 ? t : a.W - 1 - t;
    } else {
        row = a.dy > 0 ? t : a.H - 1 - t;
        if (a.dx == 0) {
            col = line;
        } else {
            int c = (line + a.dx * t) % a.W;
            col = c < 0 ? c + a.W : c;
        }
    }
}

// one step along the scanline (the incremental form of scan_pixel: no modulo in the inner loop)
__device__ __forceinline__ void scan_advance(const SgmArgs& a, int& row, int& col) {
    if (a.horizontal) {
        col += a.dx;
    } else {
        row += a.dy;
        col += a.dx;
        if (col >= a.W) col -= a.W;
        if (col < 0) col += a.W;
    }
}

// Parabola through (d-1, cm), (d, c), (d+1, cp), typed as the reference's commented-out expression would be by
// Numba (:818): fp32 differences, the factor 2 promotes the denominator to fp64. Border indices and a flat
// parabola keep the integer index (the reference is silent there).
__device__ __forceinline__ float subpixel_refine(int idx, int D, float cm, float c, float cp) {
    if (idx <= 0 || idx >= D - 1) return (float)idx;
    const float num = cp - cm;
    const double den = 2.0 * ((double)(cm + cp) - 2.0 * (double)c);
    if (!(den > 0.0)) return (float)idx;
    return (float)((double)idx - (double)num / den);
}

// order-preserving map double -> signed 64-bit integer
__device__ __forceinline__ long long dkey(double v) {
    long long b = __double_as_longlong(v);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}

// inputs are never NaN: a plain compare-select (DSETP + 2 SEL) instead of fmin()'s NaN-propagating sequence
__device__ __forceinline__ double dmin(double x, double y) { return x < y ? x : y; }

// minimum of N values as a balanced tree: the scan is bound by dependent-instruction latency (3 warps per
// scheduler), and a serial chain of N compare-selects is its longest one
template <int N>
__device__ __forceinline__ double tree_min(const double (&v)[N]) {
    double t[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = v[i];
#pragma unroll
    for (int stride = 1; stride < N; stride *= 2)
#pragma unroll
        for (int i = 0; i + stride < N; i += 2 * stride) t[i] = dmin(t[i], t[i + stride]);
    return t[0];
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ double warp_min_f64(double v) {
    long long k = dkey(v);
    int hi = (int)(k >> 32);
    unsigned lo = (unsigned)k;
    int mh = __reduce_min_sync(0xffffffffu, hi);
    unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    long long mk = ((long long)mh << 32) | (long long)ml;
    return __longlong_as_double(mk ^ ((mk >> 63) & 0x7fffffffffffffffLL));
}

template <int NPL>
__device__ __forceinline__ void load_chunk(const float* buf, int lane, float (&v)[NPL]) {
    const float* p = buf + lane * NPL;
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 4; j++) {
            float4 q = reinterpret_cast<const float4*>(p)[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else if constexpr (NPL % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 2; j++) {
            float2 q = reinterpret_cast<const float2*>(p)[j];
            v[2 * j] = q.x; v[2 * j + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPL; j++) v[j] = p[j];
    }
}

template <int NPL>
__device__ __forceinline__ void store_chunk(float* buf, int lane, const float (&v)[NPL]) {
    float* p = buf + lane * NPL;
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 4; j++)
            reinterpret_cast<float4*>(p)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else if constexpr (NPL % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 2; j++) reinterpret_cast<float2*>(p)[j] = make_float2(v[2 * j], v[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < NPL; j++) p[j] = v[j];
    }
}

// STORE = false (last pass when the caller does not keep S): no output staging rows, so more warps fit per SM;
// that pass is latency-bound (it carries the WTA reductions), not bandwidth-bound.
template <int NPL, int STAGES, int MODE, int OUTB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) sgm_scan_kernel(const SgmArgs a) {
    constexpr bool STORE = OUTB != 0;
    constexpr bool DIRECT = OUTB < 0;  // short rows: S leaves straight from the registers (vector stores), no staging
    constexpr bool kReadS = (MODE != SGM_FIRST_FUSED);
    constexpr int ROW = 32 * NPL;  // floats per row buffer
    constexpr int IN_BUFS = kReadS ? 2 : 1;
    constexpr int OUT_BUFS = OUTB > 0 ? OUTB : 0;  // 0: S is not staged; n: ring of n staging rows (a bulk store takes ~1 us to drain)
    extern __shared__ __align__(128) unsigned char smem_raw[];

    if (a.go != nullptr && *reinterpret_cast<const volatile int*>(a.go) == 0) return;  // some rank will not launch: nobody waits
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER_WARP_FLOATS = ROW * (STAGES * IN_BUFS + OUT_BUFS);
    float* wbase = reinterpret_cast<float*>(smem_raw) + (size_t)warp * PER_WARP_FLOATS;
    float* inbuf = wbase;                           // [STAGES][IN_BUFS][ROW]
    float* outbuf = wbase + ROW * STAGES * IN_BUFS;  // [2][ROW]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS_PER_CTA * PER_WARP_FLOATS * sizeof(float)) +
                     warp * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    // Disparities d >= D never need a validity test: the cost they read is +INF (the volume's pad entries
    // [D, Dp) are written as +INF by mccnn_cost_volume; the row-buffer tail [Dp, 32*NPL) is set here and never
    // overwritten by the bulk copies), so their state stays +INF and they never win a minimum.
    for (int i = lane; i < ROW * STAGES * IN_BUFS; i += 32) inbuf[i] = __int_as_float(0x7f800000);
    fence_proxy_async_smem();
    __syncwarp();

    const uint32_t copy_bytes = (uint32_t)a.Dp * 4u;
    const size_t pix_stride = (size_t)a.Dp;
    const int npix_line = a.horizontal ? a.W : a.H;
    const int d0 = lane * NPL;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    (void)d0;

    uint32_t gstep = 0;  // rows consumed by this warp so far (ring position / mbarrier phase)
    uint32_t ostep = 0;  // rows stored so far (output staging parity)

    for (;;) {
        unsigned q = 0;
        if (lane == 0) q = atomicAdd(a.counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= (unsigned)(a.nsides * a.nlines)) break;
        const int side = a.side0 + (int)q / a.nlines;
        const int line = (int)q - (side - a.side0) * a.nlines;
        const float* __restrict__ Cv = a.C[side];
        float* __restrict__ Sv = a.S[side];
        const unsigned char* __restrict__ img = a.img[side];

        // steps of this scanline that fall into the band [row0, row0 + Hb)
        int t_begin = 0, t_end = a.nsteps_total;
        if (!a.horizontal) {
            if (a.dy > 0) {
                t_begin = a.row0;
                t_end = min(a.nsteps_total, a.row0 + a.Hb);
            } else {
                t_begin = max(0, a.H - (a.row0 + a.Hb));
                t_end = min(a.nsteps_total, a.H - a.row0);
            }
        }
        if (t_begin >= t_end) continue;

        int lrow, lcol;  // pixel of the next row to prefetch (loads are issued in scanline order)
        scan_pixel(a, line, t_begin, lrow, lcol);
        auto issue_load = [&](uint32_t g) {
            const size_t off = ((size_t)(lrow - a.row0) * a.W + lcol) * pix_stride;
            const int st = g % STAGES;
            float* dst = inbuf + (size_t)st * IN_BUFS * ROW;
            // called by the whole warp (convergent); one elected lane issues (see common.cuh)
            mbar_expect_tx_elect(&bars[st], copy_bytes * IN_BUFS);
            bulk_g2s_elect(dst, Cv + off, copy_bytes, &bars[st]);
            if constexpr (kReadS) bulk_g2s_elect(dst + ROW, Sv + off, copy_bytes, &bars[st]);
        };
        auto image_at = [&](int t) -> int {
            int row, col;
            scan_pixel(a, line, max(0, min(t, npix_line - 1)), row, col);
            return (int)img[(size_t)row * a.W + col];
        };

        const int pre = min(STAGES, t_end - t_begin);
        for (int k = 0; k < pre; k++) {
            issue_load(gstep + k);
            scan_advance(a, lrow, lcol);
        }
        // image values: lane l of blk_cur holds I[pixel tb + 1 + l]
        int i_cur = image_at(t_begin);
        int blk_cur = image_at(t_begin + 1 + lane);
        int blk_next = image_at(t_begin + 33 + lane);

        double L[NPL];
#pragma unroll
        for (int j = 0; j < NPL; j++) L[j] = 1.0;
        double minL = 1.0, minLP2 = 1.0;
        bool edge_full = true;  // penalty class of the edge (t-1 -> t)

        if (t_begin > 0 && t_begin < a.nsteps_dp) {
            // resume a scanline started on another rank: wait for its state (fp64 L[], min over d)
            const size_t slot = (size_t)side * a.nlines + line;
            int gave_up = 0;
            if (lane == 0) {
                const unsigned long long t0 = global_timer_ns();
                unsigned spins = 0;
                while (ld_acquire_sys(a.flag_in + slot) != a.epoch) {
                    __nanosleep(64);
                    if ((++spins & 63u) == 0) {  // every ~4 us: has another warp given up, or is the deadline over?
                        if (*reinterpret_cast<volatile unsigned*>(a.status) != 0u) { gave_up = 1; break; }
                        if (global_timer_ns() - t0 > a.timeout_ns) {
                            atomicExch(a.status, 1u);
                            gave_up = 1;
                            break;
                        }
                    }
                }
            }
            gave_up = __shfl_sync(0xffffffffu, gave_up, 0);
            if (gave_up) {
                // the prefetched rows are already on their way into this warp's ring: let them land, then leave the kernel
                for (int k = 0; k < pre; k++) mbar_wait(&bars[(gstep + k) % STAGES], ((gstep + k) / STAGES) & 1u);
                break;
            }
            const double* src = a.hand_in + slot * HAND_STRIDE;
#pragma unroll
            for (int j = 0; j < NPL; j++) L[j] = __ldcv(src + lane * NPL + j);
            minL = __ldcv(src + 1024);
            const int dprev = i_cur - image_at(t_begin - 1);
            edge_full = (dprev >= 0) && (dprev <= a.threshold);
            minLP2 = minL + (edge_full ? a.P2 : a.P2r);
        }

        int row, col;
        scan_pixel(a, line, t_begin, row, col);
        for (int t = t_begin; t < t_end; t++, scan_advance(a, row, col)) {
            const int st = gstep % STAGES;
            mbar_wait(&bars[st], (gstep / STAGES) & 1u);
            float cf[NPL], sf[NPL];
            const float* ib = inbuf + (size_t)st * IN_BUFS * ROW;
            load_chunk<NPL>(ib, lane, cf);
            if constexpr (kReadS) load_chunk<NPL>(ib + ROW, lane, sf);
            // The stage is refilled at the END of the step, after every lane has consumed cf/sf in arithmetic:
            // a warp barrier alone does not wait for outstanding shared loads, and an early refill (async
            // proxy) could overwrite the row under a still-queued LDS.

            if (((t - t_begin) & 31) == 0 && t > t_begin) {
                blk_cur = blk_next;
                blk_next = image_at(t + 33 + lane);
            }
            const int i_next = __shfl_sync(0xffffffffu, blk_cur, (t - t_begin) & 31);
            const int dn = i_next - i_cur;
            const bool next_full = (dn >= 0) && (dn <= a.threshold);

            const bool dp_active = t < a.nsteps_dp;
            if (dp_active) {
                const bool wrapped = (!a.horizontal) && (a.dx != 0) && (t > 0) && (a.dx > 0 ? col == 0 : col == a.W - 1);
                if (t == 0 || wrapped) {
#pragma unroll
                    for (int j = 0; j < NPL; j++) L[j] = (double)cf[j];
                } else {
                    const double P1 = edge_full ? a.P1 : a.P1r;
                    double up = __shfl_up_sync(0xffffffffu, L[NPL - 1], 1);
                    double dn_ = __shfl_down_sync(0xffffffffu, L[0], 1);
                    if (lane == 0) up = INF;    // d-1 < 0: the reference clamps to L[0]+P1 >= L[0] (:300-301)
                    if (lane == 31) dn_ = INF;  // d+1 >= 32*NPL
                    double a_prev = up + P1;
                    double a_cur = L[0] + P1;
#pragma unroll
                    for (int j = 0; j < NPL; j++) {
                        const double a_next = ((j + 1 < NPL) ? L[j + 1 < NPL ? j + 1 : j] : dn_) + P1;
                        const double m = dmin(dmin(a_prev, a_next), dmin(L[j], minLP2));
                        double c = (double)cf[j];
                        c += (m - minL);
                        a_prev = a_cur;
                        a_cur = a_next;
                        L[j] = c;
                    }
                }
                // minimum over d, consumed at the next pixel together with this pixel's P2 (:332-341)
                minL = warp_min_f64(tree_min<NPL>(L));
                minLP2 = minL + (next_full ? a.P2 : a.P2r);
            }
            edge_full = next_full;
            i_cur = i_next;

            // ---- S update: fp32 += fp64, one rounding per path (:321-330)
            float so[NPL];
#pragma unroll
            for (int j = 0; j < NPL; j++) {
                if constexpr (MODE == SGM_FIRST_FUSED) {
                    // S starts at 0 (:1116-1117): down path, then the up path's raw-cost add on rows >= 1
                    float s = dp_active ? (float)L[j] : 0.0f;
                    if (row >= 1) s = (float)((double)s + (double)cf[j]);
                    so[j] = s;
                } else {
                    so[j] = dp_active ? (float)((double)sf[j] + L[j]) : sf[j];
                }
            }

            if constexpr (DIRECT) {
                float* dstp = Sv + ((size_t)(row - a.row0) * a.W + col) * pix_stride + d0;
                if constexpr (NPL % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < NPL; j += 4)
                        if (d0 + j < a.Dp) *reinterpret_cast<float4*>(dstp + j) = make_float4(so[j], so[j + 1], so[j + 2], so[j + 3]);
                } else if constexpr (NPL % 2 == 0) {
#pragma unroll
                    for (int j = 0; j < NPL; j += 2)
                        if (d0 + j < a.Dp) *reinterpret_cast<float2*>(dstp + j) = make_float2(so[j], so[j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < NPL; j++)
                        if (d0 + j < a.Dp) dstp[j] = so[j];
                }
            } else if constexpr (STORE) {
                float* ob = outbuf + (OUTB > 1 ? (ostep % OUTB) * ROW : 0);
                // bulk-group bookkeeping is per thread: elect.sync names the same lane every time (full mask), and the warp stays
                // convergent, so the copy is not wrapped in an ELECT / BRA.U.ANY loop (common.cuh)
                bulk_wait_read_elect<OUTB - 1>();  // the store issued OUTB steps ago has left its staging row
                __syncwarp();
                store_chunk<NPL>(ob, lane, so);
                fence_proxy_async_smem();
                __syncwarp();
                bulk_s2g_commit_elect(Sv + ((size_t)(row - a.row0) * a.W + col) * pix_stride, ob, copy_bytes);
                ostep++;
            }

            if constexpr (MODE == SGM_LAST_WTA) {
                // first strict minimum over d (:805-811); the neighbours of the running minimum are tracked for
                // the optional parabola refinement (:813-819)
                // (value tree first, then the lowest index holding it: no 32*NPL-long dependent chain)
                float best = __int_as_float(0x7f800000), bl = 0.f, br = 0.f;
                int bj = 0;
                {
                    float tm[NPL];  // entries d >= D are +INF or NaN here: fminf drops NaN, +INF never wins
#pragma unroll
                    for (int j = 0; j < NPL; j++) tm[j] = so[j];
#pragma unroll
                    for (int stride = 1; stride < NPL; stride *= 2)
#pragma unroll
                        for (int i = 0; i + stride < NPL; i += 2 * stride) tm[i] = fminf(tm[i], tm[i + stride]);
                    best = fminf(tm[0], best) + 0.0f;  // -0 and +0 are one value (and one key below)
                    unsigned hit[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int j = 0; j < NPL; j++) hit[j & 3] |= (so[j] == best) ? (1u << j) : 0u;
                    const unsigned any = (hit[0] | hit[1]) | (hit[2] | hit[3]);
                    bj = any ? __ffs(any) - 1 : 0;
                }
                if (a.subpixel) {
                    float left = __shfl_up_sync(0xffffffffu, so[NPL - 1], 1);
                    const float right_edge = __shfl_down_sync(0xffffffffu, so[0], 1);
#pragma unroll
                    for (int j = 0; j < NPL; j++) {
                        const float nxt = (j + 1 < NPL) ? so[j + 1 < NPL ? j + 1 : j] : right_edge;
                        if (j == bj) { bl = left; br = nxt; }
                        left = so[j];
                    }
                }
                int k = __float_as_int(best);
                k ^= (k >> 31) & 0x7fffffff;
                const int mk = __reduce_min_sync(0xffffffffu, k);
                const unsigned who = __ballot_sync(0xffffffffu, k == mk);
                const int src = __ffs(who) - 1;
                const int idx = __shfl_sync(0xffffffffu, d0 + bj, src);
                float outv = (float)idx;
                if (a.subpixel) {
                    const float c = __shfl_sync(0xffffffffu, best, src);
                    const float cm = __shfl_sync(0xffffffffu, bl, src), cp = __shfl_sync(0xffffffffu, br, src);
                    outv = subpixel_refine(idx, a.D, cm, c, cp);
                }
                if (lane == 0) a.disp[side][(size_t)(row - a.row0) * a.W + col] = outv;
            }
            // every lane's results (which depend on all of its cf/sf loads) are stored or reduced: refill the stage
            __syncwarp();
            if (t + STAGES < t_end) {
                issue_load(gstep + STAGES);
                scan_advance(a, lrow, lcol);
            }
            gstep++;
        }
        if (a.hand_out != nullptr && t_end < a.nsteps_dp) {
            // the path continues on the neighbouring rank: publish the state over NVLink, then raise its flag
            const size_t slot = (size_t)side * a.nlines + line;
            double* dst = a.hand_out + slot * HAND_STRIDE;
#pragma unroll
            for (int j = 0; j < NPL; j++) dst[lane * NPL + j] = L[j];
            if (lane == 0) dst[1024] = minL;
            __threadfence_system();
            __syncwarp();
            if (lane == 0) st_release_sys(a.flag_out + slot, a.epoch);
        }
    }
    // drained by the lane elect.sync names (the one that committed the groups)
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.wait_group 0;\n\t"
        "}" ::: "memory");
}

template <int NPL, int STAGES, int MODE, int OUTB>
int launch_scan(const SgmArgs& a, cudaStream_t stream) {
    constexpr int IN_BUFS = (MODE != SGM_FIRST_FUSED) ? 2 : 1;
    const size_t smem = (size_t)WARPS_PER_CTA * (32 * NPL * (STAGES * IN_BUFS + (OUTB > 0 ? OUTB : 0))) * sizeof(float) +
                        (size_t)WARPS_PER_CTA * STAGES * sizeof(uint64_t);
    auto kern = sgm_scan_kernel<NPL, STAGES, MODE, OUTB>;
    int per_sm = 0;
    if (int e = kernel_setup<sgm_scan_kernel<NPL, STAGES, MODE, OUTB>>(WARPS_PER_CTA * 32, smem, &per_sm)) return e;
    MCCNN_REQUIRE(per_sm >= 1, MCCNN_EINVAL, "sgm_scan_kernel<%d>: does not fit on an SM (smem %zu)", NPL, smem);
    const int total_lines = a.nsides * a.nlines;
    int grid = sm_count() * per_sm;
    const int need = ceil_div(total_lines, WARPS_PER_CTA);
    if (grid > need) grid = need;
    kern<<<grid, WARPS_PER_CTA * 32, smem, stream>>>(a);
    MCCNN_LAUNCH_CHECK("sgm_scan_kernel");
    return 0;
}

template <int MODE>
int dispatch_scan(const SgmArgs& a, cudaStream_t stream) {
    const int need = ceil_div(a.D, 32);
    // Large-D instantiations are shared-memory limited (a row is D*4 bytes): the read-modify-write passes and the
    // store-less last pass trade prefetch depth (2 stages, one staging row) for 12 instead of 8 warps per SM
    // (measured on c4: 20.0 -> 18.2..19.3 ms per RMW pass, 20.7 -> 17.5 ms for the last pass).
#define MCCNN_SGM_CASE(N, ST)                                                                       \
    if (need <= N) {                                                                                \
        if constexpr (MODE == SGM_LAST_WTA) {                                                       \
            if (!a.store_s) return launch_scan<N, (N >= 13 ? 2 : ST), MODE, 0>(a, stream);         \
        }                                                                                           \
        if constexpr (MODE == SGM_MID && N >= 13) return launch_scan<N, 2, MODE, 1>(a, stream);      \
        return launch_scan<N, ST, MODE, (N <= 8 ? -1 : 2)>(a, stream);                               \
    }
    MCCNN_SGM_CASE(1, 6)
    MCCNN_SGM_CASE(2, 6)
    MCCNN_SGM_CASE(3, 6)
    MCCNN_SGM_CASE(4, 6)
    MCCNN_SGM_CASE(5, 4)
    MCCNN_SGM_CASE(6, 4)
    MCCNN_SGM_CASE(7, 4)
    MCCNN_SGM_CASE(8, 4)
    MCCNN_SGM_CASE(10, 4)
    MCCNN_SGM_CASE(13, 3)
    MCCNN_SGM_CASE(16, 3)
    MCCNN_SGM_CASE(20, 3)
    MCCNN_SGM_CASE(25, 3)
    MCCNN_SGM_CASE(32, 3)
#undef MCCNN_SGM_CASE
    set_error("sgm: D=%d exceeds the supported maximum of 1024", a.D);
    return MCCNN_EINVAL;
}

// geometry of the reference's 8 path kernels in launch order (:1166-1202)
const int kPathDy[8] = {1, -1, 0, 0, 1, -1, 1, -1};
const int kPathDx[8] = {0, 0, 1, -1, 1, 1, -1, -1};

void set_path(SgmArgs& a, int path) {
    if (a.Hb == 0) { a.row0 = 0; a.Hb = a.H; }
    a.dy = kPathDy[path];
    a.dx = kPathDx[path];
    a.horizontal = (a.dy == 0);
    a.nlines = a.horizontal ? a.Hb : a.W;
    a.nsteps_dp = (a.horizontal ? a.W : a.H) - 1;
    a.nsteps_total = a.nsteps_dp;
}

void set_params(SgmArgs& a, const mccnn_sgm_params* p) {
    a.P1 = (double)p->P1;
    a.P2 = (double)p->P2;
    a.P1r = (double)p->P1_red;
    a.P2r = (double)p->P2_red;
    a.threshold = p->threshold;
    a.subpixel = p->subpixel;
}

int check_common(const void* C, int H, int W, int D) {
    MCCNN_REQUIRE(C != nullptr, MCCNN_EINVAL, "sgm: null volume");
    MCCNN_REQUIRE(H >= 3 && W >= 3, MCCNN_EINVAL, "sgm: image %dx%d too small (need >= 3x3)", W, H);
    MCCNN_REQUIRE(D >= 1 && D <= 1024, MCCNN_EINVAL, "sgm: D=%d outside 1..1024", D);
    return 0;
}

}  // namespace
}  // namespace mccnn

using namespace mccnn;

extern "C" size_t mccnn_sgm_workspace_bytes(int H, int W, int D) {
    if (H < 1 || W < 1 || D < 1) return 0;
    // exact mode: 64 words (scanline counters + the status word of sharded launches); the fused mode's hand-over rings and
    // flags follow (one size for both modes, so that a workspace serves either)
    return 256 + sgm_fused_workspace_bytes(H, W, D);
}

static int run_sgm(const float* CL, const float* CR, const uint8_t* imageL, const uint8_t* imageR, float* SL, float* SR,
                   float* dispL, float* dispR, void* workspace, int H, int W, int D, const mccnn_sgm_params* params,
                   int keep_volumes, const mccnn_shard* sh, int pass_mask, cudaStream_t stream) {
    unsigned* counters = reinterpret_cast<unsigned*>(workspace);
    MCCNN_CUDA(cudaMemsetAsync(counters, 0, 256, stream));

    SgmArgs a{};
    a.C[0] = CL; a.C[1] = CR;
    a.S[0] = SL; a.S[1] = SR;
    a.img[0] = imageL; a.img[1] = imageR;
    a.disp[0] = dispL; a.disp[1] = dispR;
    a.H = H; a.W = W; a.D = D; a.Dp = disp_pitch(D);
    a.nsides = 2;
    set_params(a, params);
    a.row0 = sh ? sh->row0 : 0;
    a.Hb = sh ? sh->rows : H;
    a.epoch = sh ? sh->epoch : 0;
    a.go = sh ? sh->go_flag : nullptr;
    a.status = counters + STATUS_WORD;
    a.timeout_ns = (unsigned long long)((sh && sh->timeout_ms) ? sh->timeout_ms : 2000u) * 1000000ull;

    // pass -> reference path: pass 0 = down (+ the up path's raw-cost add), 1..5 = right, left, down-right, up-right,
    // down-left, 6 = up-left + winner-takes-all (also visits row 0, which the path skips)
    static const int kPassPath[7] = {0, 2, 3, 4, 5, 6, 7};
    const size_t slot_doubles = (size_t)2 * W * HAND_STRIDE;
    // one launch: pass `pass` of side `side0` .. `side0 + nsides - 1`
    auto launch = [&](int pass, int side0, int nsides) -> int {
        set_path(a, kPassPath[pass]);
        if (pass == 0 || pass == 6) a.nsteps_total = H;
        a.store_s = (pass == 6 && !keep_volumes) ? 0 : 1;
        a.side0 = side0;
        a.nsides = nsides;
        a.counter = counters + 2 * pass + side0;
        a.hand_in = nullptr; a.flag_in = nullptr; a.hand_out = nullptr; a.flag_out = nullptr;
        if (sh && sh->world > 1 && !a.horizontal) {
            // a rank receives from the rank the path comes from and publishes to the rank it runs into
            const bool down = a.dy > 0;
            char* from_me = reinterpret_cast<char*>(sh->xchg_local);
            char* to_peer = reinterpret_cast<char*>(down ? sh->xchg_next : sh->xchg_prev);
            const bool has_in = down ? sh->rank > 0 : sh->rank < sh->world - 1;
            const bool has_out = down ? sh->rank < sh->world - 1 : sh->rank > 0;
            const size_t flags_off = (size_t)7 * slot_doubles * sizeof(double);
            if (has_in) {
                a.hand_in = reinterpret_cast<const double*>(from_me) + (size_t)pass * slot_doubles;
                a.flag_in = reinterpret_cast<const unsigned*>(from_me + flags_off) + (size_t)pass * 2 * W;
            }
            if (has_out) {
                MCCNN_REQUIRE(to_peer != nullptr, MCCNN_EINVAL, "mccnn_sgm_sharded: missing peer exchange pointer");
                a.hand_out = reinterpret_cast<double*>(to_peer) + (size_t)pass * slot_doubles;
                a.flag_out = reinterpret_cast<unsigned*>(to_peer + flags_off) + (size_t)pass * 2 * W;
            }
        }
        return pass == 0 ? dispatch_scan<SGM_FIRST_FUSED>(a, stream)
               : pass == 6 ? dispatch_scan<SGM_LAST_WTA>(a, stream)
                           : dispatch_scan<SGM_MID>(a, stream);
    };

    // (measured on c4: 2 ranks 113 -> 106 ms per pair. With more ranks the fill of a launch, world-1 waves, outweighs
    // a single-side launch of 1.6 waves and the sequential single-side launches do not win: 4 ranks 67 -> 73 ms. There
    // both sides stay in one launch until the two launches of a step can share the SMs.)
    if (!(sh && sh->world == 2)) {
        for (int pass = 0; pass < 7; pass++)
            if (pass_mask & (1 << pass))
                if (int e = launch(pass, 0, 2)) return e;
        return 0;
    }
    // Two ranks: a vertical / diagonal pass is a pipeline over the ranks (rank r resumes a scanline when rank
    // r-1, or r+1 for the upward passes, has finished its part), so with both sides in one launch every rank idles
    // for one wave of scanlines per rank boundary and pass. The two sides are independent chains: run the right
    // volume ONE PASS AHEAD of the left one. Consecutive passes either run in opposite directions (3..6: down, up,
    // down, up) or are band-local (1, 2), so in every step {left pass k-1, right pass k} each rank has a launch that
    // is fed early: it runs the one whose first scanlines arrive sooner first. Same order of steps on all ranks; within
    // a step the first launches form wait-free chains from the two ends of the rank line, so no cycle of waits.
    auto arrival = [&](int pass) -> int {  // in waves: how long until this rank's first scanline of `pass` can start
        const int path = kPassPath[pass];
        if (kPathDy[path] == 0) return 0;
        return kPathDy[path] > 0 ? sh->rank : sh->world - 1 - sh->rank;
    };
    for (int step = 0; step <= 7; step++) {
        const int pl = step - 1, pr = step;  // left pass, right pass of this step
        const bool do_l = pl >= 0 && pl < 7 && (pass_mask & (1 << pl));
        const bool do_r = pr < 7 && (pass_mask & (1 << pr));
        const bool left_first = do_l && (!do_r || arrival(pl) <= arrival(pr));
        if (do_l && left_first)
            if (int e = launch(pl, 0, 1)) return e;
        if (do_r)
            if (int e = launch(pr, 1, 1)) return e;
        if (do_l && !left_first)
            if (int e = launch(pl, 0, 1)) return e;
    }
    return 0;
}

extern "C" int mccnn_sgm(const float* CL, const float* CR, const uint8_t* imageL, const uint8_t* imageR, float* SL,
                         float* SR, float* dispL, float* dispR, void* workspace, size_t workspace_bytes, int H, int W,
                         int D, const mccnn_sgm_params* params, int mode, int keep_volumes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int e = check_common(CL, H, W, D)) return e;
    MCCNN_REQUIRE(CR && SL && SR && imageL && imageR && dispL && dispR && params && workspace, MCCNN_EINVAL,
                  "mccnn_sgm: null argument");
    MCCNN_REQUIRE(mode == MCCNN_SGM_EXACT || mode == MCCNN_SGM_FUSED, MCCNN_EINVAL, "mccnn_sgm: unknown mode %d", mode);
    MCCNN_REQUIRE(workspace_bytes >= mccnn_sgm_workspace_bytes(H, W, D), MCCNN_EWORKSPACE, "mccnn_sgm: workspace too small");
    MCCNN_REQUIRE(aligned16(CL) && aligned16(CR) && aligned16(SL) && aligned16(SR), MCCNN_EALIGN,
                  "mccnn_sgm: volumes must be 16-byte aligned");
    MCCNN_REQUIRE(params->P1 >= 0 && params->P1_red >= 0, MCCNN_EINVAL, "mccnn_sgm: negative P1");
    if (mode == MCCNN_SGM_FUSED) {
        MCCNN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, MCCNN_EALIGN, "mccnn_sgm: workspace must be 256-byte aligned");
        MCCNN_REQUIRE(params->P2 >= 0 && params->P2_red >= 0, MCCNN_EINVAL, "mccnn_sgm: the fused mode needs penalties >= 0");
        return run_sgm_fused(CL, CR, imageL, imageR, SL, SR, dispL, dispR, reinterpret_cast<char*>(workspace) + 256, H, W, D, params,
                             keep_volumes, stream);
    }
    return run_sgm(CL, CR, imageL, imageR, SL, SR, dispL, dispR, workspace, H, W, D, params, keep_volumes, nullptr, 0x7f, stream);
}

extern "C" size_t mccnn_sgm_shard_exchange_bytes(int W) {
    if (W < 1) return 0;
    const size_t payload = (size_t)7 * 2 * W * HAND_STRIDE * sizeof(double);
    const size_t flags = (size_t)7 * 2 * W * sizeof(unsigned);
    return (payload + flags + 255) & ~(size_t)255;
}

extern "C" int mccnn_sgm_sharded(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR, float* SLb,
                                 float* SRb, float* dispLb, float* dispRb, void* workspace, size_t workspace_bytes, int W,
                                 int D, const mccnn_sgm_params* params, int mode, int keep_volumes, const mccnn_shard* shard,
                                 int pass_mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(shard != nullptr, MCCNN_EINVAL, "mccnn_sgm_sharded: null shard");
    const int H = shard->H_full;
    if (int e = check_common(CLb, H, W, D)) return e;
    MCCNN_REQUIRE(CRb && SLb && SRb && imageL && imageR && dispLb && dispRb && params && workspace, MCCNN_EINVAL,
                  "mccnn_sgm_sharded: null argument");
    MCCNN_REQUIRE(mode == MCCNN_SGM_EXACT, MCCNN_EINVAL, "mccnn_sgm_sharded: mode %d is not supported (exact only)", mode);
    MCCNN_REQUIRE(workspace_bytes >= 256, MCCNN_EWORKSPACE, "mccnn_sgm_sharded: workspace too small");
    MCCNN_REQUIRE(aligned16(CLb) && aligned16(CRb) && aligned16(SLb) && aligned16(SRb), MCCNN_EALIGN,
                  "mccnn_sgm_sharded: volumes must be 16-byte aligned");
    MCCNN_REQUIRE(shard->world >= 1 && shard->rank >= 0 && shard->rank < shard->world && shard->rows >= 1 && shard->row0 >= 0 &&
                      shard->row0 + shard->rows <= H,
                  MCCNN_EINVAL, "mccnn_sgm_sharded: bad band rank=%d/%d rows [%d,+%d) of %d", shard->rank, shard->world,
                  shard->row0, shard->rows, H);
    MCCNN_REQUIRE(shard->world == 1 || (shard->xchg_local != nullptr && shard->epoch != 0), MCCNN_EINVAL,
                  "mccnn_sgm_sharded: exchange buffer and a non-zero epoch are required");
    MCCNN_REQUIRE(params->P1 >= 0 && params->P1_red >= 0, MCCNN_EINVAL, "mccnn_sgm_sharded: negative P1");
    return run_sgm(CLb, CRb, imageL, imageR, SLb, SRb, dispLb, dispRb, workspace, H, W, D, params, keep_volumes, shard,
                   pass_mask & 0x7f, stream);
}

extern "C" size_t mccnn_sgm_fused_shard_exchange_bytes(int W, int D) {
    if (W < 1 || D < 1 || D > 1024) return 0;
    return sgm_fused_xchg_bytes(W, D);
}

extern "C" int mccnn_sgm_fused_sharded(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR, float* SLb,
                                       float* SRb, float* dispLb, float* dispRb, void* workspace, size_t workspace_bytes, int W, int D,
                                       const mccnn_sgm_params* params, int keep_volumes, const mccnn_shard* shard, int sweep_mask,
                                       void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(shard != nullptr, MCCNN_EINVAL, "mccnn_sgm_fused_sharded: null shard");
    const int H = shard->H_full;
    if (int e = check_common(CLb, H, W, D)) return e;
    MCCNN_REQUIRE(CRb && SLb && SRb && imageL && imageR && dispLb && dispRb && params && workspace, MCCNN_EINVAL,
                  "mccnn_sgm_fused_sharded: null argument");
    MCCNN_REQUIRE(workspace_bytes >= mccnn_sgm_workspace_bytes(H, W, D), MCCNN_EWORKSPACE, "mccnn_sgm_fused_sharded: workspace too small");
    MCCNN_REQUIRE(aligned16(CLb) && aligned16(CRb) && aligned16(SLb) && aligned16(SRb) && (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                  MCCNN_EALIGN, "mccnn_sgm_fused_sharded: volumes must be 16-byte, the workspace 256-byte aligned");
    MCCNN_REQUIRE(shard->world >= 1 && shard->rank >= 0 && shard->rank < shard->world && shard->rows >= 1 && shard->row0 >= 0 &&
                      shard->row0 + shard->rows <= H,
                  MCCNN_EINVAL, "mccnn_sgm_fused_sharded: bad band rank=%d/%d rows [%d,+%d) of %d", shard->rank, shard->world,
                  shard->row0, shard->rows, H);
    MCCNN_REQUIRE((shard->rank == 0) == (shard->row0 == 0) && (shard->rank == shard->world - 1) == (shard->row0 + shard->rows == H),
                  MCCNN_EINVAL, "mccnn_sgm_fused_sharded: the bands must tile the image in rank order");
    MCCNN_REQUIRE(shard->world == 1 || (shard->xchg_local != nullptr && shard->epoch != 0), MCCNN_EINVAL,
                  "mccnn_sgm_fused_sharded: exchange buffer and a non-zero epoch are required");
    MCCNN_REQUIRE((shard->rank == 0 || shard->xchg_prev != nullptr) && (shard->rank == shard->world - 1 || shard->xchg_next != nullptr),
                  MCCNN_EINVAL, "mccnn_sgm_fused_sharded: missing peer exchange pointer");
    MCCNN_REQUIRE(params->P1 >= 0 && params->P1_red >= 0 && params->P2 >= 0 && params->P2_red >= 0, MCCNN_EINVAL,
                  "mccnn_sgm_fused_sharded: negative penalty");
    return run_sgm_fused_band(CLb, CRb, imageL, imageR, SLb, SRb, dispLb, dispRb, workspace, W, D, params, keep_volumes, shard,
                              sweep_mask & 15, stream);
}

extern "C" int mccnn_sgm_shard_status(const void* workspace, int* status_host, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MCCNN_REQUIRE(workspace && status_host, MCCNN_EINVAL, "mccnn_sgm_shard_status: null argument");
    unsigned v = 0;
    MCCNN_CUDA(cudaMemcpyAsync(&v, reinterpret_cast<const unsigned*>(workspace) + STATUS_WORD, sizeof(v), cudaMemcpyDeviceToHost, stream));
    MCCNN_CUDA(cudaStreamSynchronize(stream));
    *status_host = (int)v;
    if (v != 0) set_error("mccnn_sgm_sharded: a scanline waited longer than the hand-over deadline for its neighbour rank (status %u)", v);
    return 0;
}

extern "C" int mccnn_sgm_single_path(const float* C, const uint8_t* image, float* S, void* workspace,
                                     size_t workspace_bytes, int H, int W, int D, const mccnn_sgm_params* params,
                                     int path, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (int e = check_common(C, H, W, D)) return e;
    MCCNN_REQUIRE(S && image && params && workspace, MCCNN_EINVAL, "mccnn_sgm_single_path: null argument");
    MCCNN_REQUIRE(path >= 0 && path < 8, MCCNN_EINVAL, "mccnn_sgm_single_path: path %d outside 0..7", path);
    MCCNN_REQUIRE(workspace_bytes >= 8, MCCNN_EWORKSPACE, "mccnn_sgm_single_path: workspace too small");
    MCCNN_REQUIRE(aligned16(C) && aligned16(S), MCCNN_EALIGN, "mccnn_sgm_single_path: volumes must be 16-byte aligned");
    unsigned* counter = reinterpret_cast<unsigned*>(workspace);
    MCCNN_CUDA(cudaMemsetAsync(counter, 0, 4, stream));
    SgmArgs a{};
    a.C[0] = C; a.S[0] = S; a.img[0] = image; a.disp[0] = nullptr;
    a.C[1] = C; a.S[1] = S; a.img[1] = image; a.disp[1] = nullptr;
    a.H = H; a.W = W; a.D = D; a.Dp = disp_pitch(D);
    a.nsides = 1;
    a.side0 = 0;
    a.store_s = 1;
    a.status = counter + 1;   // unused without sharding
    set_params(a, params);
    if (path == 1) {  // penalty channels 0/1 are never written by the reference: P1 = P2 = 0
        a.P1 = a.P2 = a.P1r = a.P2r = 0.0;
    }
    set_path(a, path);
    a.counter = counter;
    return dispatch_scan<SGM_MID>(a, stream);
}
