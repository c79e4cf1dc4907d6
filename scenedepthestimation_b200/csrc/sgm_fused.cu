// Semi-global matching, FUSED mode (MCCNN_SGM_FUSED, sm_100a): the 8 paths in 4 sweeps instead of the exact mode's 7 passes.
//
// Same recurrence, traversal extents, column wraps, restarts and penalty rule as the reference (SGM_Interation,
// process_functional.py:265-343; the 8 path kernels :346-797; penalties :134-262; WTA :800-837) -- see sgm.cu for the
// contract. What this mode gives up, and why it is opt-in: the reference keeps the path state in fp64 and rounds S to fp32
// once per path IN LAUNCH ORDER (:1166-1202), which forces one read-modify-write pass of S per path (76 B per evaluation and
// side, sgm.cu). Here the path state is fp32 and the contributions are added into S in the order
//     down, down-right, up (raw cost), left, down-left, right, up-right, up-left
// so that two paths share every sweep: 8 + 12 + 12 + 8 = 40 B per evaluation and side. oracle/stereo.py::sgm_all_paths_fused
// restates exactly this arithmetic (the kernels are compared with it value for value); tools/fused_census.py measures the
// distance to the exact mode.
//
// Sweep structure. A sweep pairs a path that runs ALONG a unit (a column for sweep 0, an image row for sweeps 1 and 2; its
// state never leaves the owning warp's registers) with the diagonal path whose predecessor pixel is the previous step of
// the NEIGHBOURING unit:
//     sweep 0: unit = column x, steps down the rows:      down (own)  + down-right (from column x-1, row y-1) + raw cost
//     sweep 1: unit = row y,    steps from right to left: left (own)  + down-left  (from row y-1, column x+1)
//     sweep 2: unit = row y,    steps from left to right: right (own) + up-right   (from row y+1, column x-1)
//     sweep 3: unit = diagonal scanline (as in sgm.cu):   up-left + winner-takes-all, nothing handed over
// One warp owns a unit; lane l holds NPL consecutive disparities. Unit u at step t needs the diagonal state unit u-1 produced
// at step t-1, so the warps of a chain advance in lock step, each one step behind nothing: the hand-over only makes a warp
// wait when its neighbour is late. The state (one row of D floats + its minimum) travels through a 2-slot ring in shared
// memory between warps of a CTA, through an 8-slot ring in global memory (L2-resident) between the last warp of a CTA and
// the first warp of the next one, and through a T-slot global buffer from the last warp of the chain to the first one,
// which by then works on the next round of units (unit u -> warp u mod n). The global rings are served by two LINK WARPS
// per CTA (one per direction: prefetch of announced rows into a shared-memory ring for warp 0, forwarding of the last compute
// warp's ring), so every compute warp sees shared-memory rings on both sides and runs the same step. The grid is launched
// cooperatively: every warp of a chain must be resident.
// Cost and S rows stream through shared memory with 1-D bulk copies + mbarriers, as in sgm.cu.
// BAND kernels (mccnn_sgm_fused_sharded): the same sweeps on a band of image rows, the neighbour GPUs' bands reached through
// peer memory (entry / exit states per column in sweep 0, a row FIFO per step in sweeps 1 and 2, a state per scanline in
// sweep 3); value-identical to the unsharded sweeps.
#include "common.cuh"
#include <cstdlib>

namespace mccnn {
namespace {

constexpr int FW_MAX = 10;  // compute warps per CTA of the chain kernel (9 for rows of 1024 floats: shared memory); + 2 link warps
constexpr int RING = 2;     // hand-over slots between warps of one CTA (shared memory)
constexpr int GRING = 8;    // hand-over slots between neighbouring CTAs (global memory)
constexpr int XPUB = 4;    // a band's last unit publishes its rows to the neighbour rank every XPUB steps, XPUB steps late
constexpr int INR = 4;      // rows of the global input link prefetched into shared memory (per CTA)
constexpr int FSTAGES = 2;  // rows in flight per warp and stream
constexpr int FLAG_STRIDE = 32;  // unsigned words between two flags (one 128-byte line each)
constexpr int MAX_CHAIN_CTAS = 512;
constexpr float kInf = __builtin_huge_valf();

struct FusedArgs {
    int side_base;   // first side (0 = left volume, 1 = right) of a chain launch; a launch covers gridDim.x / ctas sides
    const float* C[2];
    float* S[2];
    const unsigned char* img[2];
    float* disp[2];
    int H, W, D, Dp;
    int sweep;        // 0, 1, 2: see above
    int U, T;         // units per side, steps per unit
    int ctas;         // CTAs per chain (one chain per side)
    float P1, P2, P1r, P2r;
    int threshold;
    int subpixel;
    int store_s;
    float* glink;     // global hand-over buffers of this launch: [side][ctas x GRING slots | wrap: T slots][slot_floats]
    unsigned* gflags; // [side][ctas][2][FLAG_STRIDE]: produced / consumed counters of the link leaving CTA c
    int slot_floats;  // 32 * NPL + 4
    unsigned* counter;  // sweep 3: scanline counter
    int strict;         // 1 (default): gpu-scope release on the producer side of the global links; MCCNN_FUSED_STRICT=0 drops it (measurement only)
    // ---- row-band sharding (BAND kernels; one pair split over several GPUs, mccnn_sgm_fused_sharded): this launch owns image rows
    // [row0, row0 + Hb) of H; the volumes and maps it is given hold only those rows, the u8 images are whole. What crosses a
    // band boundary travels through the neighbours' exchange buffers (peer memory): see the layout in fused_xchg_layout().
    int row0, Hb;
    unsigned epoch;
    const float* ent_in;  const unsigned* ent_flag_in;     // sweep 0: states of row row0 - 1 per column, written by the rank above
    float* ent_out;       unsigned* ent_flag_out;          //          ... of my last row, into the rank below
    const float* ext_in;  const unsigned long long* ext_prod_in;   // sweeps 1, 2: the row FIFO feeding my first unit
    float* ext_out;       unsigned long long* ext_prod_out;        //              ... fed by my last unit (neighbour's memory)
    const float* hand_in; const unsigned* hand_flag_in;    // sweep 3: state of a scanline entering from the rank below
    float* hand_out;      unsigned* hand_flag_out;         //          ... leaving into the rank above
    const int* go;                 // all ranks agreed to launch (0 = return at once); may be null
    unsigned* status;              // set to 1 when a cross-rank wait hit the deadline (results invalid)
    unsigned long long timeout_ns;
};

// ------------------------------------------------------------------------------------------------ small device helpers
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all_elect() {   // full completion (writes performed), not just the source reads
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.wait_group %0;\n\t"
        "}" ::"n"(N)
        : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wait (lane 0 polls, the warp follows) until a flag written by ANOTHER GPU equals `want`; gives up at the deadline, or at once
// when some warp has already given up: a rank that died or never launched cannot wedge this one (the results are then invalid
// and *status says so).
__device__ __forceinline__ void wait_peer_flag(const unsigned* flag, unsigned want, unsigned* status, unsigned long long timeout_ns, int lane) {
    if (lane == 0) {
        const unsigned long long t0 = timer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys_u32(flag) != want) {
            __nanosleep(64);
            if ((++spins & 63u) == 0) {
                if (*reinterpret_cast<volatile unsigned*>(status) != 0u) break;
                if (timer_ns() - t0 > timeout_ns) { atomicExch(status, 1u); break; }
            }
        }
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned ld_acquire_cta_smem(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_smem(unsigned* p, unsigned v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

__device__ __forceinline__ float warp_min_f32(float v) {
    int k = __float_as_int(v);
    k ^= (k >> 31) & 0x7fffffff;  // order-preserving float -> int
    k = __reduce_min_sync(0xffffffffu, k);
    k ^= (k >> 31) & 0x7fffffff;
    return __int_as_float(k) + 0.0f;  // -0 and +0 are one value
}

template <int N>
__device__ __forceinline__ float tree_min_f32(const float (&v)[N]) {
    float t[N];
#pragma unroll
    for (int i = 0; i < N; i++) t[i] = v[i];
#pragma unroll
    for (int stride = 1; stride < N; stride *= 2)
#pragma unroll
        for (int i = 0; i + stride < N; i += 2 * stride) t[i] = fminf(t[i], t[i + stride]);
    return t[0];
}

template <int NPL>
__device__ __forceinline__ void load_row(const float* buf, int lane, float (&v)[NPL]) {
    const float* p = buf + lane * NPL;
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 4; j++) {
            const float4 q = reinterpret_cast<const float4*>(p)[j];
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else if constexpr (NPL % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 2; j++) {
            const float2 q = reinterpret_cast<const float2*>(p)[j];
            v[2 * j] = q.x; v[2 * j + 1] = q.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPL; j++) v[j] = p[j];
    }
}

template <int NPL>
__device__ __forceinline__ void store_row(float* buf, int lane, const float (&v)[NPL]) {
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

// global-memory twins (L2 only: the other SM's data must never be served from this SM's L1)
template <int NPL>
__device__ __forceinline__ void load_row_cg(const float* buf, int lane, float (&v)[NPL]) {
    const float* p = buf + lane * NPL;
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 4; j++) {
            const float4 q = __ldcg(reinterpret_cast<const float4*>(p) + j);
            v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPL; j++) v[j] = __ldcg(p + j);
    }
}
template <int NPL>
__device__ __forceinline__ void store_row_cg(float* buf, int lane, const float (&v)[NPL]) {
    float* p = buf + lane * NPL;
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NPL / 4; j++)
            __stcg(reinterpret_cast<float4*>(p) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    } else {
#pragma unroll
        for (int j = 0; j < NPL; j++) __stcg(p + j, v[j]);
    }
}

// One step of the recurrence, in place: L <- c + (min(min(L[d-1], L[d+1]) + P1, L[d], minL + P2) - minL).
// min(a + P1, b + P1) == min(a, b) + P1 in floating point (rounding is monotonic). Out-of-range neighbours are +INF: the
// reference's clamps L[-1] -> L[0], L[D] -> L[D-1] (:300-303) never win against L[d] itself.
template <int NPL>
__device__ __forceinline__ void dp_step(float (&L)[NPL], const float (&c)[NPL], float minL, float P1, float P2, int lane) {
    float up = __shfl_up_sync(0xffffffffu, L[NPL - 1], 1);
    float dn = __shfl_down_sync(0xffffffffu, L[0], 1);
    if (lane == 0) up = kInf;
    if (lane == 31) dn = kInf;
    const float mp2 = minL + P2;
    float prev = up;
#pragma unroll
    for (int j = 0; j < NPL; j++) {
        const float next = (j + 1 < NPL) ? L[j + 1 < NPL ? j + 1 : j] : dn;
        const float m = fminf(fminf(prev, next) + P1, fminf(L[j], mp2));
        prev = L[j];
        L[j] = c[j] + (m - minL);
    }
}

// first strict minimum over d of this warp's row (:805-811), optional parabola refinement (:813-819), as in sgm.cu
__device__ __forceinline__ float subpixel_refine_f(int idx, int D, float cm, float c, float cp) {
    if (idx <= 0 || idx >= D - 1) return (float)idx;
    const float num = cp - cm;
    const double den = 2.0 * ((double)(cm + cp) - 2.0 * (double)c);
    if (!(den > 0.0)) return (float)idx;
    return (float)((double)idx - (double)num / den);
}

template <int NPL>
__device__ __forceinline__ float warp_wta(const float (&so)[NPL], int lane, int D, int subpixel) {
    const int d0 = lane * NPL;
    float best = kInf, bl = 0.f, br = 0.f;
    int bj = 0;
    {
        best = fminf(tree_min_f32<NPL>(so), best) + 0.0f;
        unsigned hit[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < NPL; j++) hit[j & 3] |= (so[j] == best) ? (1u << j) : 0u;
        const unsigned any = (hit[0] | hit[1]) | (hit[2] | hit[3]);
        bj = any ? __ffs(any) - 1 : 0;
    }
    if (subpixel) {
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
    if (subpixel) {
        const float c = __shfl_sync(0xffffffffu, best, src);
        const float cm = __shfl_sync(0xffffffffu, bl, src), cp = __shfl_sync(0xffffffffu, br, src);
        outv = subpixel_refine_f(idx, D, cm, c, cp);
    }
    return outv;
}

// ------------------------------------------------------------------------------------------------ sweeps 0, 1, 2
// Shared-memory layout of the chain kernel (floats unless noted):
//   per compute warp : inbuf [FSTAGES][NIN][ROWF] | outbuf [ROWF] | ring [RING][SLOTF]   (SLOTF = ROWF + 4: the row and, at [ROWF], its minimum)
//   per CTA  : inring [INR][SLOTF]  rows prefetched from the previous CTA's global ring (read by warp 0 only)
//              prodc [FW], consc [FW] (unsigned)   counters of the links leaving the compute warps
//              inpop, incons (unsigned)            warp 0 -> link warp: rows popped from inring (running total / link numbering)
//              xcons (unsigned)                    link warp -> the warp of a band's last unit: rows of that unit forwarded
//              mbarriers: FSTAGES per warp, then INR for inring
// Warps: fw <= FW compute warps (one unit each at a time) and ONE LINK WARP (the last warp of the CTA) that moves the rows of
// the two global links of the CTA: it prefetches the rows the previous CTA published into inring for warp 0 and forwards the
// rows warp fw - 1 leaves in its ring to the next CTA (or, BAND, to the neighbour GPU). Every compute warp therefore sees the
// same thing on both sides -- a shared-memory ring with a counter -- and runs the same step. (Before, warp 0 and warp fw - 1 did
// the global polling, fencing and copying themselves; ncu showed the interior warps polling their predecessor 27 times per
// step on average: the whole chain ran at the pace of those two warps.)
// Rows in flight per warp: a sweep that reads S as well has one (cost + S) row pair under way per warp, refilled right after its
// consumption; the first sweep (cost only) has two stages.
template <int NPL, bool READS>
struct ChainSmem {
    static constexpr int FW = NPL > 25 ? 9 : FW_MAX;   // compute warps (the block has two more warps)
    static constexpr int STG = READS ? 1 : 2;
    static constexpr int NIN = READS ? 2 : 1;
    static constexpr int ROWF = 32 * NPL;
    static constexpr int SLOTF = ROWF + 4;
    static constexpr int RG = RING;   // hand-over slots between the warps of a CTA (3 measured: no gain, c4 63.3 vs 64.0 ms)
    static constexpr int PER_WARP = ROWF * (STG * NIN + 1) + RG * SLOTF;
    static constexpr size_t BYTES = (size_t)FW * PER_WARP * 4 + (size_t)INR * SLOTF * 4 + (size_t)(2 * FW + 4) * 4 + (size_t)(FW * STG + INR) * 8;
};

template <int NPL, bool READS, bool BAND>
__global__ void __launch_bounds__((ChainSmem<NPL, READS>::FW + 2) * 32) sgm_chain_kernel(const FusedArgs a) {
    if constexpr (BAND) {
        if (a.go != nullptr && *reinterpret_cast<const volatile int*>(a.go) == 0) return;   // some rank will not launch: nobody waits
    }
    using L = ChainSmem<NPL, READS>;
    constexpr int FW = L::FW;
    constexpr int FSTAGES = L::STG;   // (shadows the file-level default)
    constexpr int RING = L::RG;       // (likewise)
    constexpr int ROWF = L::ROWF, SLOTF = L::SLOTF;
    constexpr int NIN = L::NIN;
    constexpr int PER_WARP = L::PER_WARP;
    constexpr uint32_t SLOT_BYTES = SLOTF * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fw = (int)(blockDim.x >> 5) - 2;   // compute warps of this launch (<= FW; the shared-memory layout is FW's); warps fw, fw + 1 = link warps
    float* smem_f = reinterpret_cast<float*>(smem_raw);
    float* wbase = smem_f + (size_t)(warp < fw ? warp : 0) * PER_WARP;
    float* inbuf = wbase;                            // [FSTAGES][NIN][ROWF]
    float* outbuf = wbase + ROWF * FSTAGES * NIN;    // [ROWF]
    float* ring = outbuf + ROWF;                     // [RING][SLOTF]: the link warp -> warp + 1 (the last compute warp's is drained by the link warp)
    float* inring = smem_f + (size_t)FW * PER_WARP;                              // [INR][SLOTF]
    unsigned* prodc = reinterpret_cast<unsigned*>(inring + INR * SLOTF);         // [FW] rows published on the link leaving warp w
    unsigned* consc = prodc + FW;                                                // [FW] rows the reader of that link is done with
    unsigned* inpop = consc + FW;                                                // rows warp 0 has popped from inring, running total
    unsigned* incons = inpop + 1;                                                // ... in the numbering of the chain's global link
    unsigned* xcons = inpop + 2;                                                 // BAND: rows of the band's last unit the link warp has taken
    uint64_t* bars_all = reinterpret_cast<uint64_t*>(inpop + 4);
    uint64_t* bars = bars_all + (warp < fw ? warp : 0) * FSTAGES;
    uint64_t* inbars = bars_all + FW * FSTAGES;                                  // [INR]

    if (lane == 0 && warp < fw) {
#pragma unroll
        for (int s = 0; s < FSTAGES; s++) mbar_init(&bars[s], 1);
        if (warp == 0) {
            for (int s = 0; s < INR; s++) mbar_init(&inbars[s], 1);
            *inpop = 0u;
            *incons = 0u;
            *xcons = 0u;
        }
        mbar_fence_init();
        prodc[warp] = 0u;
        consc[warp] = 0u;
    }
    // tails [Dp, 32 * NPL) of the row buffers stay +INF for ever (the bulk copies bring Dp floats): those disparities never win
    if (warp < fw)
        for (int i = lane; i < ROWF * FSTAGES * NIN; i += 32) inbuf[i] = kInf;
    fence_proxy_async_smem();
    __syncthreads();

    const int side = a.side_base + blockIdx.x / a.ctas, cta = blockIdx.x % a.ctas;
    const int n = a.ctas * fw;                 // compute warps of this chain
    const int W = a.W, H = a.H, T = a.T;

    // links. The one that leaves CTA c for CTA c + 1 is global link c; link ctas - 1 closes the ring (T slots).
    // A global link is a FIFO of rows: the producer side pushes one row per step of every unit it hands over (T per unit,
    // payload or not), the consumer pops T rows per unit; row index = round * T + step.
    const int lin = cta == 0 ? a.ctas - 1 : cta - 1, lout = cta;
    const size_t side_floats = ((size_t)a.ctas * GRING + (size_t)T) * SLOTF;
    auto link_base = [&](int l) { return a.glink + (size_t)side * side_floats + (size_t)l * GRING * SLOTF; };
    const unsigned depth_in = lin == a.ctas - 1 ? (unsigned)T : (unsigned)GRING;
    const unsigned depth_out = lout == a.ctas - 1 ? (unsigned)T : (unsigned)GRING;

    // BAND: local rows are [0, Hb) of the volumes, image rows [row0, row0 + Hb); ug = the unit's index in the WHOLE image's chain
    const int row0 = BAND ? a.row0 : 0;
    const int Hloc = BAND ? a.Hb : H;       // rows of the volumes this launch works on
    const int ubase = !BAND ? 0 : (a.sweep == 1 ? row0 : (a.sweep == 2 ? H - (row0 + Hloc) : 0));
    // which units take rows from / hand rows to a global link (both warps of a link and the link warps use these)
    auto unit_ext_in = [&](int u) -> bool { return BAND && a.sweep != 0 && a.ext_in != nullptr && u == 0; };
    auto unit_ext_out = [&](int u) -> bool { return BAND && a.sweep != 0 && a.ext_out != nullptr && u == a.U - 1; };
    auto unit_pops = [&](int u) -> bool { return ubase + u >= 1 && (u >= 1 || unit_ext_in(u)); };       // unit of a warp 0
    auto unit_pushes = [&](int u) -> bool { return a.sweep == 0 ? true : (ubase + u <= H - 2); };       // unit of a warp fw - 1

    // ================================================================================================ the link warp
    if (warp >= fw) {
        if (lane != 0) return;
        const bool do_out = warp == fw, do_in = warp == fw + 1;   // one warp per direction: a gpu-scope release / acquire per row each
        const float* gin = link_base(lin);
        float* gout = link_base(lout);
        unsigned* gprod_in = a.gflags + ((size_t)(side * a.ctas + lin) * 2 + 0) * FLAG_STRIDE;
        unsigned* gcons_in = a.gflags + ((size_t)(side * a.ctas + lin) * 2 + 1) * FLAG_STRIDE;
        unsigned* gprod_out = a.gflags + ((size_t)(side * a.ctas + lout) * 2 + 0) * FLAG_STRIDE;
        unsigned* gcons_out = a.gflags + ((size_t)(side * a.ctas + lout) * 2 + 1) * FLAG_STRIDE;
        // the rows to forward come from the ring of warp fw - 1; BAND: the band's last unit feeds the neighbour GPU from whatever
        // warp it runs on (it is the last unit of the chain, so it comes after everything else this warp forwards)
        int osrc = fw - 1;
        bool ext_tail = BAND && a.sweep != 0 && a.ext_out != nullptr && ((a.U - 1) % n) / fw == cta && ((a.U - 1) % n) % fw != fw - 1 &&
                        unit_pushes(a.U - 1);
        // Ordering on the global links. The rows reach L2 through the copy engine. Producer side: cp.async.bulk.wait_group (the
        // copy is complete), fence.proxy.async.global, counter store. Consumer side: counter load, fence.acq_rel.gpu,
        // fence.proxy.async.global, and only then the copy engine reads the rows.
        // The consumer's gpu-scope fence is NOT optional. Measured at c4 (tools/check_race_c4.py): announcing a row right after
        // wait_group and fetching it without that fence corrupts thousands of rows per sweep (the global slots are reused every
        // GRING rows, and the reading SM's copy engine can still be served the slot's previous contents); announcing it a step
        // later (what round 1 / 2 did from the compute warps, never observed to fail) only makes that window improbable. With
        // the fence: 0 differing rows in 24 / 24 repeated c4 sweeps, and the announcement no longer has to lag.
        // The gpu-scope RELEASE on the producer side (the counter store is st.release.gpu; completion of a bulk copy is defined
        // for the issuing thread only) is not optional either: without it a single pair at c4 came out right in 72 / 72 repeated
        // sweeps, but with several pairs in flight on their own streams (test_fused_pairs_in_flight_equal_sequential) two runs
        // in three lost a few rows; with it, none. It costs 62.5 -> 72.5 ms of SGM at c4 (a MEMBAR.GPU under 6 TB/s of traffic
        // takes microseconds). MCCNN_FUSED_STRICT=0 drops it, for measurements only.
        // None of this is in the step of a compute warp: one link warp per direction.
        // ---- input side: the units of warp 0
        int iu = cta * fw, iround = 0;
        unsigned it = 0;            // rows of the current unit requested so far
        unsigned in_req = 0;        // rows ever requested into inring: slot and mbarrier phase
        unsigned avail = 0;         // rows published on the current source
        unsigned cons_pub = 0;      // last value forwarded to the producer of the chain link
        unsigned long long wait_t0 = 0;
        while (iu < a.U && !unit_pops(iu)) { iu += n; iround++; }
        // ---- output side: the units of warp fw - 1
        int ou = cta * fw + fw - 1, oround = 0;
        unsigned ot = 0;            // rows of the current unit forwarded so far
        unsigned cons_seen = 0;     // rows the next CTA has popped from the chain link
        bool unpublished = false;   // copies have been committed whose rows are not announced yet
        unsigned idle = 0;          // polls without progress
        auto publish = [&](unsigned v) { if (a.strict) st_release_gpu(gprod_out, v); else st_relaxed_gpu(gprod_out, v); };
        bool in_closed = false;
        while (ou < a.U && !unit_pushes(ou)) { ou += n; oround++; }
        auto next_out_unit = [&]() {
            do { ou += n; oround++; } while (ou < a.U && !unit_pushes(ou));
            if (ou >= a.U && ext_tail) {
                ext_tail = false;
                ou = a.U - 1; oround = (a.U - 1) / n; osrc = ((a.U - 1) % n) % fw;
            }
        };
        if (ou >= a.U && ext_tail) { ou -= n; oround--; next_out_unit(); }

        if (!do_in) iu = a.U;
        if (!do_out) ou = a.U;
        while (iu < a.U || ou < a.U) {
            bool progress = false;
            if (ou < a.U) {
                const unsigned qo = (unsigned)oround * (unsigned)T + ot;        // the row's index on the link leaving warp fw - 1
                if (ld_acquire_cta_smem(&prodc[osrc]) >= qo + 1u) {
                    const bool ext = unit_ext_out(ou);
                    bool slot_free = ext || qo + 1u <= depth_out || cons_seen >= qo + 1u - depth_out;
                    if (!slot_free) {
                        cons_seen = ld_relaxed_gpu(gcons_out);
                        slot_free = cons_seen >= qo + 1u - depth_out;
                        if (slot_free && a.strict) fence_acq_rel_gpu();   // the consumer's reads of the slot happened before what we write now
                    }
                    if (slot_free) {
                        fence_proxy_async_smem();
                        float* dst = gout + (size_t)(qo % depth_out) * SLOTF;
                        if constexpr (BAND) {
                            if (ext) dst = a.ext_out + ((size_t)side * T + ot) * SLOTF;   // T slots: none is reused inside a pair
                        }
                        const float* src_ring = smem_f + (size_t)osrc * PER_WARP + ROWF * (FSTAGES * NIN + 1);
                        bulk_s2g(dst, src_ring + (size_t)(qo % RING) * SLOTF, SLOT_BYTES);
                        bulk_commit();
                        bulk_wait_read<0>();                                   // the ring slot has been read: its warp may reuse it
                        ot++;
                        if (osrc == fw - 1) st_release_cta_smem(&consc[fw - 1], qo + 1u);
                        else st_release_cta_smem(xcons, ot);
                        const bool unit_done = ot == (unsigned)T;
                        // Publication runs behind the copies: a row is announced once its copy is COMPLETE, and waiting for the
                        // copy just committed would put the write latency of L2 (or of NVLink) into every step.
                        if (!ext) {   // the next CTA: announce the row as soon as its copy is complete
                            bulk_wait_all<0>();
                            fence_proxy_async_global();
                            publish(qo + 1u);
                        } else if constexpr (BAND) {
                            unpublished = !unit_done;
                            // the neighbour GPU: the counter carries the pair's epoch; every XPUB rows, XPUB rows late
                            if (unit_done) {
                                bulk_wait_all<0>();
                                fence_proxy_async_global();
                                st_release_sys_u64(a.ext_prod_out + side * 16, ((unsigned long long)a.epoch << 32) | (unsigned long long)T);
                            } else if (ot >= 2u * XPUB && (ot % XPUB) == 0u) {
                                bulk_wait_all<XPUB>();
                                fence_proxy_async_global();
                                st_release_sys_u64(a.ext_prod_out + side * 16, ((unsigned long long)a.epoch << 32) | (unsigned long long)(ot - XPUB));
                            }
                        }
                        if (unit_done) {
                            ot = 0;
                            next_out_unit();
                        }
                        progress = true;
                    }
                }
            }
            if (iu < a.U) {
                const bool ext = unit_ext_in(iu);
                if (!ext) {   // tell the producer which slots of the chain link are free again
                    const unsigned cv = ld_acquire_cta_smem(incons);
                    if (cv - cons_pub >= (unsigned)(GRING / 2)) {   // (a release per row is not needed: the ring is GRING deep)
                        if (a.strict) st_release_gpu(gcons_in, cv); else st_relaxed_gpu(gcons_in, cv);
                        cons_pub = cv;
                    }
                }
                if (in_req < ld_acquire_cta_smem(inpop) + (unsigned)INR) {   // a free inring slot
                    const unsigned qbase = ext ? 0u : (unsigned)(cta == 0 ? iround - 1 : iround) * (unsigned)T;
                    const unsigned q = qbase + it;
                    if (avail <= q) {
                        if constexpr (BAND) {
                            if (ext) {   // epoch-tagged 64-bit counter in this rank's exchange buffer, written by the neighbour GPU
                                const unsigned long long v = ld_acquire_sys_u64(a.ext_prod_in + side * 16);
                                avail = (unsigned)(v >> 32) == a.epoch ? (unsigned)v : 0u;
                                if (avail <= q) {   // another GPU feeds this FIFO: the wait is bounded
                                    if (wait_t0 == 0) wait_t0 = timer_ns();
                                    if (*reinterpret_cast<volatile unsigned*>(a.status) != 0u) avail = (unsigned)T;
                                    else if (timer_ns() - wait_t0 > a.timeout_ns) { atomicExch(a.status, 1u); avail = (unsigned)T; }
                                } else {
                                    wait_t0 = 0;
                                }
                            } else {
                                avail = ld_relaxed_gpu(gprod_in);
                            }
                        } else {
                            avail = ld_relaxed_gpu(gprod_in);
                        }
                        if (avail > q) fence_acq_rel_gpu();   // acquire: the rows below `avail` are complete in L2
                    }
                    if (avail > q) {
                        const unsigned sl = in_req % INR;
                        const float* src = gin + (size_t)(q % depth_in) * SLOTF;
                        if constexpr (BAND) {
                            if (ext) src = a.ext_in + ((size_t)side * T + it) * SLOTF;
                        }
                        fence_proxy_async_global();
                        mbar_expect_tx(&inbars[sl], SLOT_BYTES);
                        bulk_g2s(inring + (size_t)sl * SLOTF, src, SLOT_BYTES, &inbars[sl]);
                        in_req++;
                        it++;
                        if (it == (unsigned)T) {
                            it = 0;
                            if (ext) avail = 0;   // back to the chain's own link, whose numbering starts at 0
                            do { iu += n; iround++; } while (iu < a.U && !unit_pops(iu));
                        }
                        progress = true;
                    }
                }
            }
            if (do_in && iu >= a.U && !in_closed) {
                // no more units for warp 0: whatever still arrives on the input link is not needed (its rows are all in shared
                // memory); the producer must not wait for free slots any more
                st_relaxed_gpu(gcons_in, 0xffffffffu);
                in_closed = true;
            }
            if (progress) {
                idle = 0;
            } else {
                if (BAND && unpublished && ou < a.U && ++idle >= 8u) {   // neighbour GPU: nothing else to do, announce the latest rows now
                    bulk_wait_all<0>();
                    fence_proxy_async_global();
                    st_release_sys_u64(a.ext_prod_out + side * 16, ((unsigned long long)a.epoch << 32) | (unsigned long long)ot);
                    unpublished = false;
                    idle = 0;
                } else {
                    __nanosleep(40);
                }
            }
        }
        if (do_in && !in_closed) st_relaxed_gpu(gcons_in, 0xffffffffu);
        bulk_wait_all<0>();
        return;
    }

    // ================================================================================================ the compute warps
    const int w = cta * fw + warp;             // position in the chain
    const float* __restrict__ Cv = a.C[side];
    float* __restrict__ Sv = a.S[side];
    const unsigned char* __restrict__ img = a.img[side];
    const uint32_t copy_bytes = (uint32_t)a.Dp * 4u;
    const size_t pitch = (size_t)a.Dp;
    const bool in_local = warp > 0, out_local = warp < fw - 1;
    const float* lring_in = ring - PER_WARP;   // the previous warp's ring (only dereferenced when in_local)

    // geometry of the sweep: pixel(u, t), the step along the unit and the offset to the neighbouring unit's pixel
    const int dpix = a.sweep == 0 ? W : (a.sweep == 1 ? -1 : 1);
    const int upoff = a.sweep == 0 ? -1 : (a.sweep == 1 ? -W : W);

    uint32_t gstep = 0;      // rows consumed by this warp so far (stage ring position / mbarrier phase)
    unsigned ring_popped = 0;   // warp 0: rows ever popped from inring (slot and mbarrier phase)

    int round = 0;
    for (int u = w; u < a.U; u += n, round++) {
        const int ug = ubase + u;
        // first pixel of the unit in the (band-local) volumes and in the (whole) image
        const long long pix0 = a.sweep == 0 ? (long long)u : (a.sweep == 1 ? (long long)u * W + (W - 1) : (long long)(Hloc - 1 - u) * W);
        const long long ipix0 = pix0 + (long long)row0 * W;
        const bool has_up = ug >= 1;
        const bool diag_unit = unit_pushes(u);
        // the producer of my input link works on unit u - 1: same round, or the previous one across the closing link
        const unsigned qbase_in = unit_ext_in(u) ? 0u : (unsigned)((w == 0 ? round - 1 : round)) * (unsigned)T;   // only used when has_up
        const unsigned qbase_out = (unsigned)round * (unsigned)T;
        const bool gin_active = !in_local && unit_pops(u);       // this unit pops T rows from inring (filled by the link warp)
        const bool ext_tail_unit = out_local && unit_ext_out(u);  // BAND: the band's last unit on a warp other than the CTA's last
        const bool gout_push = (!out_local || ext_tail_unit) && diag_unit;   // ... leaves T rows in its ring for the link warp
        const bool gin_counts = gin_active && !unit_ext_in(u);   // pops are reported in the chain link's numbering

        const long long dstep = (long long)dpix * (long long)pitch;   // floats from one step's row to the next
        long long loff = pix0 * (long long)pitch;                        // row to prefetch next
        long long soff = loff;                                          // row of the current step
        auto issue_load = [&](uint32_t g) {
            const long long off = loff;
            const int st = g % FSTAGES;
            float* dst = inbuf + (size_t)st * NIN * ROWF;
            mbar_expect_tx_elect(&bars[st], copy_bytes * NIN);
            bulk_g2s_elect(dst, Cv + off, copy_bytes, &bars[st]);
            if constexpr (READS) bulk_g2s_elect(dst + ROWF, Sv + off, copy_bytes, &bars[st]);
            loff += dstep;
        };
        {
            const int pre = min(FSTAGES, T);
            for (int k = 0; k < pre; k++) issue_load(gstep + k);
        }
        // image values of the next 32 steps of this unit and of the neighbouring unit (lane l: step tb + l)
        auto img_own = [&](int t) -> int { return (int)img[ipix0 + (long long)min(t, T - 1) * dpix]; };
        auto img_up = [&](int t) -> int { return has_up ? (int)img[ipix0 + upoff + (long long)min(t, T - 1) * dpix] : 0; };
        int blk_own = img_own(lane), blk_up = img_up(lane);
        int nblk_own = img_own(32 + lane), nblk_up = img_up(32 + lane);
        int i_prev_own = 0, i_prev_up = 0;
        // BAND, sweep 0, not the image's first rows: the paths continue from the rank above (entry states, written into this
        // rank's exchange buffer per column: row 0 = the down path's state, row 1 = the down-right path's state, each + minimum)
        const bool entry = BAND && a.sweep == 0 && a.ent_in != nullptr;
        if constexpr (BAND) {
            if (entry) {
                i_prev_own = (int)img[ipix0 - W];
                i_prev_up = has_up ? (int)img[ipix0 - W - 1] : 0;
            }
        }
        // pop the next row of inring (warp 0): wait for the link warp's copy to land
        auto pop_row = [&]() -> const float* {
            const unsigned k = ring_popped++;
            mbar_wait(&inbars[k % INR], (k / INR) & 1u);
            return inring + (size_t)(k % INR) * SLOTF;
        };
        // ... and give the slot back once the row has been read in arithmetic (rows_done = pops of this unit so far)
        auto popped_done = [&](unsigned rows_done) {
            __syncwarp();
            if (lane == 0) {
                if (gin_counts) st_release_cta_smem(incons, qbase_in + rows_done);
                st_release_cta_smem(inpop, ring_popped);
            }
        };

        float Lo[NPL];   // state of the path that runs along this unit
        float mo = 0.f;
#pragma unroll
        for (int j = 0; j < NPL; j++) Lo[j] = 0.f;

        for (int t = 0; t < T; t++, soff += dstep) {
            const int st = gstep % FSTAGES;
            mbar_wait(&bars[st], (gstep / FSTAGES) & 1u);
            float cf[NPL], so[NPL];
            const float* ib = inbuf + (size_t)st * NIN * ROWF;
            load_row<NPL>(ib, lane, cf);
            if constexpr (READS) load_row<NPL>(ib + ROWF, lane, so);

            if ((t & 31) == 0 && t > 0) {
                blk_own = nblk_own; blk_up = nblk_up;
                nblk_own = img_own(t + 32 + lane); nblk_up = img_up(t + 32 + lane);
            }
            const int i_cur = __shfl_sync(0xffffffffu, blk_own, t & 31);
            const int i_upcur = __shfl_sync(0xffffffffu, blk_up, t & 31);

            // sweep 0 steps down the image rows (global row yg); sweeps 1 and 2 step along a row (all rows are visited)
            const int yg = row0 + t;
            const bool own_active = a.sweep == 0 ? (yg <= H - 2) : (t <= T - 2);
            const bool diag_active = diag_unit && (a.sweep != 0 || yg <= H - 2);
            const bool diag_first = a.sweep == 0 ? (yg == 0 || u == 0) : (t == 0 || ug == 0);
            const bool from_entry = entry && t == 0;                                // continue the rank above's paths

            // ---- the path along the unit
            if constexpr (BAND) {
                if (from_entry) {
                    const size_t slot = (size_t)side * W + u;
                    wait_peer_flag(a.ent_flag_in + slot, a.epoch, a.status, a.timeout_ns, lane);
                    const float* src = a.ent_in + slot * (2 * SLOTF);
                    load_row_cg<NPL>(src, lane, Lo);
                    mo = __ldcg(src + ROWF);
                }
            }
            if (own_active) {
                // The first pixel of a path takes the raw cost. No branch for it: the state starts as Lo = 0, mo = 0, and with
                // penalties >= 0 (checked on the host) the step gives min(0 + P1, 0, 0 + P2) - 0 = 0, i.e. Lo = cf + 0.
                const int dn = i_cur - i_prev_own;
                const bool full = (dn >= 0) && (dn <= a.threshold);
                dp_step<NPL>(Lo, cf, mo, full ? a.P1 : a.P1r, full ? a.P2 : a.P2r, lane);
                mo = warp_min_f32(tree_min_f32<NPL>(Lo));
            }
            if constexpr (READS) {
                if (own_active) {
#pragma unroll
                    for (int j = 0; j < NPL; j++) so[j] = so[j] + Lo[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < NPL; j++) so[j] = own_active ? Lo[j] : 0.0f;
            }
            // Every lane has consumed this stage's cost (and S) row in arithmetic (the recurrence and the minimum read every
            // cf[j], the add every so[j]; refills only happen for t <= T - 3, where the own path is active), so all its shared
            // loads have returned: refill the stage NOW, most of a step earlier than at the end of the step.
            __syncwarp();
            if (t + FSTAGES < T) issue_load(gstep + FSTAGES);

            // ---- the diagonal path: predecessor = step t - 1 of unit u - 1
            const float* grow = nullptr;
            if (gin_active && t >= 1) grow = pop_row();   // popped whether it is needed or not (FIFO)
            // the row this step hands to unit u + 1 (row qo of my output link); the last compute warp leaves one per step for
            // the link warp, payload or not (FIFO). Its slot must have been read by the consumer (0xffffffff = reader gone).
            const bool push = gout_push || (out_local && diag_active);
            const unsigned qo = qbase_out + (unsigned)t;
            if (ext_tail_unit) {   // the link warp counts this unit's rows from 0
                if (push && t + 1 > RING) {
                    if (lane == 0)
                        while (ld_acquire_cta_smem(xcons) < (unsigned)(t + 1 - RING)) __nanosleep(20);
                    __syncwarp();
                }
            } else if (push && qo + 1u > RING) {
                if (lane == 0)
                    while (ld_acquire_cta_smem(&consc[warp]) < qo + 1u - RING) __nanosleep(20);
                __syncwarp();
            }
            float md = 0.f;
            if (diag_active) {
                float Ld[NPL];
                if (diag_first) {   // first pixel of the path: zero state, the step below then yields the raw cost (see the own path)
#pragma unroll
                    for (int j = 0; j < NPL; j++) Ld[j] = 0.f;
                } else {
                    if (BAND && from_entry) {
                        const size_t slot = (size_t)side * W + (u - 1);
                        wait_peer_flag(a.ent_flag_in + slot, a.epoch, a.status, a.timeout_ns, lane);
                        const float* src = a.ent_in + slot * (2 * SLOTF) + SLOTF;
                        load_row_cg<NPL>(src, lane, Ld);
                        md = __ldcg(src + ROWF);
                    } else if (in_local) {
                        const unsigned q = qbase_in + (unsigned)(t - 1);   // index of the row I need on my input link
                        if (lane == 0)
                            while (ld_acquire_cta_smem(&prodc[warp - 1]) < q + 1u) __nanosleep(20);
                        __syncwarp();
                        const float* src = lring_in + (size_t)(q % RING) * SLOTF;
                        load_row<NPL>(src, lane, Ld);
                        md = src[ROWF];
                    } else {
                        load_row<NPL>(grow, lane, Ld);
                        md = grow[ROWF];
                    }
                }
                {
                    const int dn = i_cur - i_prev_up;
                    const bool full = (dn >= 0) && (dn <= a.threshold);
                    dp_step<NPL>(Ld, cf, md, full ? a.P1 : a.P1r, full ? a.P2 : a.P2r, lane);
                }
                md = warp_min_f32(tree_min_f32<NPL>(Ld));
#pragma unroll
                for (int j = 0; j < NPL; j++) so[j] = so[j] + Ld[j];

                if (push) {
                    float* dst = ring + (size_t)(qo % RING) * SLOTF;
                    store_row<NPL>(dst, lane, Ld);
                    if (lane == 0) dst[ROWF] = md;
                    // the link warp hands this slot to the copy engine: the WRITING lanes order their generic-proxy stores
                    // before the async proxy's read (a fence on the reading side alone does not)
                    if (gout_push) fence_proxy_async_smem();
                }
                if constexpr (BAND) {
                    // sweep 0, last row of the band: the rank below continues both paths from here (peer memory)
                    if (a.sweep == 0 && a.ent_out != nullptr && t == T - 1) {
                        const size_t slot = (size_t)side * W + u;
                        float* dst = a.ent_out + slot * (2 * SLOTF);
                        store_row_cg<NPL>(dst, lane, Lo);
                        store_row_cg<NPL>(dst + SLOTF, lane, Ld);
                        if (lane == 0) {
                            __stcg(dst + ROWF, mo);
                            __stcg(dst + SLOTF + ROWF, md);
                        }
                        __threadfence_system();
                        __syncwarp();
                        if (lane == 0) st_release_sys_u32(a.ent_flag_out + slot, a.epoch);
                    }
                }
            }
            if (push) {
                __syncwarp();
                if (lane == 0) st_release_cta_smem(&prodc[warp], qo + 1u);
            }
            // my input link: everything up to the row of step t - 1 is consumed (whether it was needed or not)
            if (in_local && has_up && t >= 1) {
                __syncwarp();
                if (lane == 0) st_release_cta_smem(&consc[warp - 1], qbase_in + (unsigned)t);
            }
            if (gin_active && t >= 1) popped_done((unsigned)t);

            // ---- the "up" path adds the raw cost on rows >= 1 (its penalties are never written, sgm.cu); sweep 0 only
            if constexpr (!READS) {
                if (yg >= 1) {  // sweep 0: image row yg
#pragma unroll
                    for (int j = 0; j < NPL; j++) so[j] = so[j] + cf[j];
                }
            }

            // ---- S row out: staged, one bulk store per row. The store of the previous step has left the staging row.
            bulk_wait_read_elect<0>();
            __syncwarp();
            store_row<NPL>(outbuf, lane, so);
            fence_proxy_async_smem();
            __syncwarp();
            bulk_s2g_commit_elect(Sv + soff, outbuf, copy_bytes);

            i_prev_own = i_cur;
            i_prev_up = i_upcur;
            gstep++;
        }
        // the unit is done
        if (in_local && has_up) {   // its input link is consumed to the end of the producer's round
            __syncwarp();
            if (lane == 0) st_release_cta_smem(&consc[warp - 1], qbase_in + (unsigned)T);
        }
        if (gin_active) {           // FIFO: pop the row of the producer's last step too
            pop_row();
            popped_done((unsigned)T);
        }
    }
    // no more units for this warp: whatever still arrives on its input link is not needed
    __syncwarp();
    if (lane == 0 && in_local) st_release_cta_smem(&consc[warp - 1], 0xffffffffu);
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.wait_group 0;\n\t"
        "}" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ sweep 3
// up-left path + winner-takes-all: one warp per diagonal scanline (start column at the bottom row, column - 1 per step,
// wrapping modulo W with a restart), S is read, the sum goes to the WTA and, if asked for, back to S.
constexpr int LW = 4;  // warps per CTA

template <int NPL, bool STORE, bool BAND>
__global__ void __launch_bounds__(LW * 32) sgm_fused_last_kernel(const FusedArgs a) {
    if constexpr (BAND) {
        if (a.go != nullptr && *reinterpret_cast<const volatile int*>(a.go) == 0) return;
    }
    constexpr int ROWF = 32 * NPL;
    constexpr int SLOTF = ROWF + 4;
    constexpr int PER_WARP = ROWF * (FSTAGES * 2 + (STORE ? 1 : 0));
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* wbase = reinterpret_cast<float*>(smem_raw) + (size_t)warp * PER_WARP;
    float* inbuf = wbase;
    float* outbuf = wbase + ROWF * FSTAGES * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)LW * PER_WARP * sizeof(float)) + warp * FSTAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < FSTAGES; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    for (int i = lane; i < ROWF * FSTAGES * 2; i += 32) inbuf[i] = kInf;
    fence_proxy_async_smem();
    __syncwarp();

    const int W = a.W, H = a.H;
    const uint32_t copy_bytes = (uint32_t)a.Dp * 4u;
    const size_t pitch = (size_t)a.Dp;
    // BAND: this launch owns image rows [row0, row0 + Hb), i.e. the steps [tb, te) of every scanline (step t is row H - 1 - t)
    const int row0 = BAND ? a.row0 : 0;
    const int tb = BAND ? H - (a.row0 + a.Hb) : 0, te = BAND ? H - a.row0 : H;
    uint32_t gstep = 0;
    for (;;) {
        unsigned q = 0;
        if (lane == 0) q = atomicAdd(a.counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= 2u * (unsigned)W) break;
        const int side = (int)q / W, line = (int)q - side * W;
        const float* __restrict__ Cv = a.C[side];
        float* __restrict__ Sv = a.S[side];
        const unsigned char* __restrict__ img = a.img[side];
        auto col_at = [&](int t) -> int {
            int c = (line - t) % W;
            return c < 0 ? c + W : c;
        };
        int lrow = H - 1 - tb, lcol = col_at(tb);
        auto issue_load = [&](uint32_t g) {
            const size_t off = ((size_t)(lrow - row0) * W + lcol) * pitch;
            const int st = g % FSTAGES;
            float* dst = inbuf + (size_t)st * 2 * ROWF;
            mbar_expect_tx_elect(&bars[st], copy_bytes * 2);
            bulk_g2s_elect(dst, Cv + off, copy_bytes, &bars[st]);
            bulk_g2s_elect(dst + ROWF, Sv + off, copy_bytes, &bars[st]);
            lrow -= 1;
            lcol = lcol == 0 ? W - 1 : lcol - 1;
        };
        {
            const int pre = min(FSTAGES, te - tb);
            for (int k = 0; k < pre; k++) issue_load(gstep + k);
        }
        auto image_at = [&](int t) -> int {
            const int tt = min(t, H - 1);
            return (int)img[(size_t)(H - 1 - tt) * W + col_at(tt)];
        };
        int blk = image_at(tb + lane), nblk = image_at(tb + 32 + lane);
        int i_prev = 0;
        float L[NPL];
        float mL = 0.f;
#pragma unroll
        for (int j = 0; j < NPL; j++) L[j] = 0.f;
        if constexpr (BAND) {
            if (tb > 0 && tb <= H - 2) {   // the scanline enters from the rank below: continue from its state
                const size_t slot = (size_t)side * W + line;
                wait_peer_flag(a.hand_flag_in + slot, a.epoch, a.status, a.timeout_ns, lane);
                const float* src = a.hand_in + slot * SLOTF;
                load_row_cg<NPL>(src, lane, L);
                mL = __ldcg(src + ROWF);
                i_prev = image_at(tb - 1);
            }
        }
        int row = H - 1 - tb, col = col_at(tb);
        for (int t = tb; t < te; t++) {
            const int st = gstep % FSTAGES;
            mbar_wait(&bars[st], (gstep / FSTAGES) & 1u);
            float cf[NPL], so[NPL];
            const float* ib = inbuf + (size_t)st * 2 * ROWF;
            load_row<NPL>(ib, lane, cf);
            load_row<NPL>(ib + ROWF, lane, so);
            if (((t - tb) & 31) == 0 && t > tb) {
                blk = nblk;
                nblk = image_at(t + 32 + lane);
            }
            const int i_cur = __shfl_sync(0xffffffffu, blk, (t - tb) & 31);
            if (t <= H - 2) {
                if (t == 0 || col == W - 1) {   // first pixel, or the scanline has just wrapped around the image
#pragma unroll
                    for (int j = 0; j < NPL; j++) L[j] = cf[j];
                } else {
                    const int dn = i_cur - i_prev;
                    const bool full = (dn >= 0) && (dn <= a.threshold);
                    dp_step<NPL>(L, cf, mL, full ? a.P1 : a.P1r, full ? a.P2 : a.P2r, lane);
                }
                mL = warp_min_f32(tree_min_f32<NPL>(L));
#pragma unroll
                for (int j = 0; j < NPL; j++) so[j] = so[j] + L[j];
            }
            i_prev = i_cur;
            if constexpr (STORE) {
                bulk_wait_read_elect<0>();
                __syncwarp();
                store_row<NPL>(outbuf, lane, so);
                fence_proxy_async_smem();
                __syncwarp();
                bulk_s2g_commit_elect(Sv + ((size_t)(row - row0) * W + col) * pitch, outbuf, copy_bytes);
            }
            const float outv = warp_wta<NPL>(so, lane, a.D, a.subpixel);
            if (lane == 0) a.disp[side][(size_t)(row - row0) * W + col] = outv;
            __syncwarp();
            if (t + FSTAGES < te) issue_load(gstep + FSTAGES);
            gstep++;
            row -= 1;
            col = col == 0 ? W - 1 : col - 1;
        }
        if constexpr (BAND) {
            if (a.hand_out != nullptr && te <= H - 2) {   // the path goes on in the rank above (peer memory)
                const size_t slot = (size_t)side * W + line;
                float* dst = a.hand_out + slot * SLOTF;
                store_row_cg<NPL>(dst, lane, L);
                if (lane == 0) __stcg(dst + ROWF, mL);
                __threadfence_system();
                __syncwarp();
                if (lane == 0) st_release_sys_u32(a.hand_flag_out + slot, a.epoch);
            }
        }
    }
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.wait_group 0;\n\t"
        "}" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ host side
inline int npl_for(int D) {
    static const int kSizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 16, 20, 25, 32};
    const int need = ceil_div(D, 32);
    for (int s : kSizes)
        if (need <= s) return s;
    return 0;
}

struct FusedLayout {
    size_t flags, counter, links, end;
    int slot_floats;
};

FusedLayout fused_layout(int H, int W, int D) {
    FusedLayout l{};
    const int npl = npl_for(D);
    l.slot_floats = 32 * (npl ? npl : 32) + 4;
    size_t o = 0;
    l.flags = o; o += (size_t)3 * 2 * MAX_CHAIN_CTAS * 2 * FLAG_STRIDE * sizeof(unsigned);   // 3 sweeps x 2 sides x links x {prod, cons}
    l.counter = o; o += 256;
    l.links = o; o += (size_t)2 * ((size_t)MAX_CHAIN_CTAS * GRING + (size_t)(H > W ? H : W)) * l.slot_floats * sizeof(float);
    l.end = (o + 255) & ~(size_t)255;
    return l;
}

inline int fused_strict() {
    static const int v = [] { const char* e = getenv("MCCNN_FUSED_STRICT"); return e ? atoi(e) : 1; }();
    return v;
}

// rounds a chain of `units` needs when `sides` chains share the GPU
template <int NPL, bool READS, bool BAND>
int chain_rounds(int units, int sides, int* rounds) {
    constexpr int FW = ChainSmem<NPL, READS>::FW;
    int per_sm = 0;
    if (int e = kernel_setup<sgm_chain_kernel<NPL, READS, BAND>>((FW + 2) * 32, ChainSmem<NPL, READS>::BYTES, &per_sm)) return e;
    int ctas = (sm_count() * per_sm) / sides;
    if (ctas > MAX_CHAIN_CTAS) ctas = MAX_CHAIN_CTAS;
    *rounds = ctas >= 1 ? ceil_div(units, ctas * FW) : (1 << 20);
    return 0;
}

// sides = 2: both volumes in one launch (one chain each, half the GPU each); sides = 1: the chain of side a.side_base alone
template <int NPL, bool READS, bool BAND>
int launch_chain(FusedArgs a, cudaStream_t stream, int sides = 2) {
    constexpr int FW = ChainSmem<NPL, READS>::FW;
    const size_t smem = ChainSmem<NPL, READS>::BYTES;
    int per_sm = 0;
    if (int e = kernel_setup<sgm_chain_kernel<NPL, READS, BAND>>((FW + 2) * 32, smem, &per_sm)) return e;
    MCCNN_REQUIRE(per_sm >= 1, MCCNN_EINVAL, "sgm_chain_kernel<%d>: does not fit on an SM (smem %zu)", NPL, smem);
    int ctas = (sm_count() * per_sm) / sides;   // one chain per side; every CTA must be resident (cooperative launch)
    if (ctas > MAX_CHAIN_CTAS) ctas = MAX_CHAIN_CTAS;
    MCCNN_REQUIRE(ctas >= 1, MCCNN_EINVAL, "sgm_chain_kernel: no resident CTA available");
    // full rounds: the units of a side are dealt to ctas * fw warps; keep the last round as full as the others, and spread
    // its units over every SM (fewer warps per CTA) rather than filling some SMs with FW warps and leaving the others idle
    const int rounds = ceil_div(a.U, ctas * FW);
    const int per_round = ceil_div(a.U, rounds);
    int fw = ceil_div(per_round, ctas);
    static const int env_fw = [] { const char* e = getenv("MCCNN_FUSED_FW"); return e ? atoi(e) : 0; }();
    if (env_fw > 0 && env_fw <= FW && ceil_div(per_round, env_fw) <= ctas) fw = env_fw;
    if (fw > FW) fw = FW;
    ctas = ceil_div(per_round, fw);
    a.ctas = ctas;
    void* params[] = {&a};
    MCCNN_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(sgm_chain_kernel<NPL, READS, BAND>), dim3(sides * ctas), dim3((fw + 2) * 32), params,
                                           smem, stream));
    return 0;
}

template <int NPL, bool STORE, bool BAND>
int launch_last(const FusedArgs& a, cudaStream_t stream) {
    constexpr int ROWF = 32 * NPL;
    const size_t smem = (size_t)LW * ROWF * (FSTAGES * 2 + (STORE ? 1 : 0)) * sizeof(float) + (size_t)LW * FSTAGES * sizeof(uint64_t);
    int per_sm = 0;
    if (int e = kernel_setup<sgm_fused_last_kernel<NPL, STORE, BAND>>(LW * 32, smem, &per_sm)) return e;
    MCCNN_REQUIRE(per_sm >= 1, MCCNN_EINVAL, "sgm_fused_last_kernel<%d>: does not fit on an SM (smem %zu)", NPL, smem);
    int grid = sm_count() * per_sm;
    const int need = ceil_div(2 * a.W, LW);
    if (grid > need) grid = need;
    sgm_fused_last_kernel<NPL, STORE, BAND><<<grid, LW * 32, smem, stream>>>(a);
    MCCNN_LAUNCH_CHECK("sgm_fused_last_kernel");
    return 0;
}

template <int NPL>
int run_fused_npl(FusedArgs a, int keep_volumes, cudaStream_t stream, unsigned* flags) {
    const char* env_mask = getenv("MCCNN_FUSED_SWEEPS");
    const int mask = env_mask ? atoi(env_mask) : 15;
    a.strict = fused_strict();
    for (int sweep = 0; sweep < 3; sweep++) {
        if (!(mask & (1 << sweep))) continue;
        a.sweep = sweep;
        a.U = sweep == 0 ? a.W : a.H;
        a.T = sweep == 0 ? a.H : a.W;
        a.gflags = flags + (size_t)sweep * 2 * MAX_CHAIN_CTAS * 2 * FLAG_STRIDE;
        if (int e = (sweep == 0 ? launch_chain<NPL, false, false>(a, stream) : launch_chain<NPL, true, false>(a, stream))) return e;
    }
    a.store_s = keep_volumes;
    if (!(mask & 8)) return 0;
    return keep_volumes ? launch_last<NPL, true, false>(a, stream) : launch_last<NPL, false, false>(a, stream);
}

// ---- one pair split over several GPUs by image rows (mccnn_sgm_fused_sharded)
// Exchange buffer of a rank = everything its neighbours write INTO it (peer memory), one section per sweep:
//   ent   [2 sides][W][2 slots]   sweep 0: states of the down / down-right paths after the last row of the rank above (+ flags)
//   fifo1 [2][W steps][slot]      sweep 1: the down-left states of the rank above's last row, one per step (+ counter)
//   fifo2 [2][W steps][slot]      sweep 2: the up-right states of the rank below's first row (+ counter)
//   hand  [2][W][slot]            sweep 3: state of every up-left scanline leaving the rank below (+ flags)
// Flags carry the pair's epoch and are never reset; the counters are (epoch << 32 | rows pushed).
struct XchgLayout {
    size_t ent, ent_flag, fifo[2], prod, hand, hand_flag, end;
};

XchgLayout fused_xchg_layout(int W, int D) {
    XchgLayout l{};
    const int npl = npl_for(D);
    const size_t slot = (size_t)(32 * (npl ? npl : 32) + 4) * sizeof(float);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o = 0;
    l.ent = o; o += up((size_t)2 * W * 2 * slot);
    l.ent_flag = o; o += up((size_t)2 * W * sizeof(unsigned));
    for (int k = 0; k < 2; k++) { l.fifo[k] = o; o += up((size_t)2 * W * slot); }
    l.prod = o; o += up((size_t)2 * 2 * 16 * sizeof(unsigned long long));
    l.hand = o; o += up((size_t)2 * W * slot);
    l.hand_flag = o; o += up((size_t)2 * W * sizeof(unsigned));
    l.end = o;
    return l;
}

template <int NPL>
int run_fused_band_npl(FusedArgs a, int keep_volumes, cudaStream_t stream, unsigned* flags, const mccnn_shard* sh, int sweep_mask) {
    const XchgLayout x = fused_xchg_layout(a.W, a.D);
    char* mine = reinterpret_cast<char*>(sh->xchg_local);
    char* prev = reinterpret_cast<char*>(sh->xchg_prev);
    char* next = reinterpret_cast<char*>(sh->xchg_next);
    const bool has_prev = sh->rank > 0, has_next = sh->rank < sh->world - 1;
    for (int sweep = 0; sweep < 3; sweep++) {
        if (!(sweep_mask & (1 << sweep))) continue;
        a.sweep = sweep;
        a.U = sweep == 0 ? a.W : a.Hb;
        a.T = sweep == 0 ? a.Hb : a.W;
        a.gflags = flags + (size_t)sweep * 2 * MAX_CHAIN_CTAS * 2 * FLAG_STRIDE;
        a.ent_in = nullptr; a.ent_out = nullptr; a.ext_in = nullptr; a.ext_out = nullptr;
        if (sweep == 0) {
            if (has_prev) {
                a.ent_in = reinterpret_cast<const float*>(mine + x.ent);
                a.ent_flag_in = reinterpret_cast<const unsigned*>(mine + x.ent_flag);
            }
            if (has_next) {
                a.ent_out = reinterpret_cast<float*>(next + x.ent);
                a.ent_flag_out = reinterpret_cast<unsigned*>(next + x.ent_flag);
            }
        } else {
            // sweep 1 runs down the rows (fed by the rank above, feeds the rank below), sweep 2 up
            char* from = sweep == 1 ? (has_prev ? mine : nullptr) : (has_next ? mine : nullptr);
            char* to = sweep == 1 ? (has_next ? next : nullptr) : (has_prev ? prev : nullptr);
            const size_t prod_off = x.prod + (size_t)(sweep - 1) * 2 * 16 * sizeof(unsigned long long);
            if (from) {
                a.ext_in = reinterpret_cast<const float*>(from + x.fifo[sweep - 1]);
                a.ext_prod_in = reinterpret_cast<const unsigned long long*>(from + prod_off);
            }
            if (to) {
                a.ext_out = reinterpret_cast<float*>(to + x.fifo[sweep - 1]);
                a.ext_prod_out = reinterpret_cast<unsigned long long*>(to + prod_off);
            }
        }
        a.side_base = 0;
        if (sweep == 0) {
            if (int e = launch_chain<NPL, false, true>(a, stream)) return e;
            continue;
        }
        // Row sweeps: rank r can start when the rank it follows starts its LAST round, so with R rounds per rank the sweep
        // takes (world - 1)(R - 1) + R round times. When a band has more rows than half the GPU holds warps but not more than
        // the whole GPU does (c4 on 2 GPUs: 994 rows, 888 / 1776 warps), one side after the other on the whole GPU is faster:
        // 2 x 1 round in lock step with the neighbour instead of 3.
        int r2 = 0, r1 = 0;
        if (int e = chain_rounds<NPL, true, true>(a.U, 2, &r2)) return e;
        if (int e = chain_rounds<NPL, true, true>(a.U, 1, &r1)) return e;
        const int hops = sh->world - 1;
        static const int env_seq = [] { const char* e = getenv("MCCNN_FUSED_SIDES_SEQ"); return e ? atoi(e) : -1; }();
        const bool seq = env_seq >= 0 ? env_seq != 0 : 2 * (hops * (r1 - 1) + r1) < hops * (r2 - 1) + r2;
        if (!seq) {
            if (int e = launch_chain<NPL, true, true>(a, stream)) return e;
        } else {
            for (int side = 0; side < 2; side++) {
                a.side_base = side;
                if (int e = launch_chain<NPL, true, true>(a, stream, 1)) return e;
            }
        }
    }
    if (!(sweep_mask & 8)) return 0;
    a.store_s = keep_volumes;
    a.hand_in = nullptr; a.hand_out = nullptr;
    if (has_next) {
        a.hand_in = reinterpret_cast<const float*>(mine + x.hand);
        a.hand_flag_in = reinterpret_cast<const unsigned*>(mine + x.hand_flag);
    }
    if (has_prev) {
        a.hand_out = reinterpret_cast<float*>(prev + x.hand);
        a.hand_flag_out = reinterpret_cast<unsigned*>(prev + x.hand_flag);
    }
    return keep_volumes ? launch_last<NPL, true, true>(a, stream) : launch_last<NPL, false, true>(a, stream);
}

}  // namespace

size_t sgm_fused_xchg_bytes(int W, int D) { return fused_xchg_layout(W, D).end; }

// workspace: the caller's SGM workspace (first 256 bytes = the exact mode's words incl. the status word at index 32, then the fused
// mode's flags and rings)
int run_sgm_fused_band(const float* CLb, const float* CRb, const uint8_t* imageL, const uint8_t* imageR, float* SLb, float* SRb,
                       float* dispLb, float* dispRb, void* workspace, int W, int D, const mccnn_sgm_params* p, int keep_volumes,
                       const mccnn_shard* sh, int sweep_mask, cudaStream_t stream) {
    const int H = sh->H_full;
    const FusedLayout l = fused_layout(H, W, D);
    char* base = reinterpret_cast<char*>(workspace);
    char* ws = base + 256;
    if (sweep_mask & 1) MCCNN_CUDA(cudaMemsetAsync(base, 0, 256, stream));
    MCCNN_CUDA(cudaMemsetAsync(ws + l.flags, 0, l.links - l.flags, stream));
    FusedArgs a{};
    a.C[0] = CLb; a.C[1] = CRb;
    a.S[0] = SLb; a.S[1] = SRb;
    a.img[0] = imageL; a.img[1] = imageR;
    a.disp[0] = dispLb; a.disp[1] = dispRb;
    a.H = H; a.W = W; a.D = D; a.Dp = disp_pitch(D);
    a.P1 = p->P1; a.P2 = p->P2; a.P1r = p->P1_red; a.P2r = p->P2_red;
    a.threshold = p->threshold;
    a.subpixel = p->subpixel;
    a.glink = reinterpret_cast<float*>(ws + l.links);
    a.slot_floats = l.slot_floats;
    a.counter = reinterpret_cast<unsigned*>(ws + l.counter);
    a.row0 = sh->row0; a.Hb = sh->rows; a.epoch = sh->epoch;
    a.go = sh->go_flag;
    a.strict = fused_strict();
    a.status = reinterpret_cast<unsigned*>(base) + 32;
    a.timeout_ns = (unsigned long long)(sh->timeout_ms ? sh->timeout_ms : 2000u) * 1000000ull;
    unsigned* flags = reinterpret_cast<unsigned*>(ws + l.flags);
    switch (npl_for(D)) {
#define MCCNN_FUSED_BAND_CASE(N) \
    case N: return run_fused_band_npl<N>(a, keep_volumes, stream, flags, sh, sweep_mask);
        MCCNN_FUSED_BAND_CASE(1) MCCNN_FUSED_BAND_CASE(2) MCCNN_FUSED_BAND_CASE(3) MCCNN_FUSED_BAND_CASE(4) MCCNN_FUSED_BAND_CASE(5)
        MCCNN_FUSED_BAND_CASE(6) MCCNN_FUSED_BAND_CASE(7) MCCNN_FUSED_BAND_CASE(8) MCCNN_FUSED_BAND_CASE(10) MCCNN_FUSED_BAND_CASE(13)
        MCCNN_FUSED_BAND_CASE(16) MCCNN_FUSED_BAND_CASE(20) MCCNN_FUSED_BAND_CASE(25) MCCNN_FUSED_BAND_CASE(32)
#undef MCCNN_FUSED_BAND_CASE
    }
    set_error("sgm (fused, sharded): D=%d exceeds the supported maximum of 1024", D);
    return MCCNN_EINVAL;
}

size_t sgm_fused_workspace_bytes(int H, int W, int D) { return fused_layout(H, W, D).end; }

int run_sgm_fused(const float* CL, const float* CR, const uint8_t* imageL, const uint8_t* imageR, float* SL, float* SR, float* dispL,
                  float* dispR, void* workspace, int H, int W, int D, const mccnn_sgm_params* p, int keep_volumes,
                  cudaStream_t stream) {
    const FusedLayout l = fused_layout(H, W, D);
    char* ws = reinterpret_cast<char*>(workspace);
    MCCNN_CUDA(cudaMemsetAsync(ws + l.flags, 0, l.links - l.flags, stream));
    FusedArgs a{};
    a.C[0] = CL; a.C[1] = CR;
    a.S[0] = SL; a.S[1] = SR;
    a.img[0] = imageL; a.img[1] = imageR;
    a.disp[0] = dispL; a.disp[1] = dispR;
    a.H = H; a.W = W; a.D = D; a.Dp = disp_pitch(D);
    a.P1 = p->P1; a.P2 = p->P2; a.P1r = p->P1_red; a.P2r = p->P2_red;
    a.threshold = p->threshold;
    a.subpixel = p->subpixel;
    a.glink = reinterpret_cast<float*>(ws + l.links);
    a.slot_floats = l.slot_floats;
    a.counter = reinterpret_cast<unsigned*>(ws + l.counter);
    unsigned* flags = reinterpret_cast<unsigned*>(ws + l.flags);
    switch (npl_for(D)) {
#define MCCNN_FUSED_CASE(N) \
    case N: return run_fused_npl<N>(a, keep_volumes, stream, flags);
        MCCNN_FUSED_CASE(1) MCCNN_FUSED_CASE(2) MCCNN_FUSED_CASE(3) MCCNN_FUSED_CASE(4) MCCNN_FUSED_CASE(5) MCCNN_FUSED_CASE(6)
        MCCNN_FUSED_CASE(7) MCCNN_FUSED_CASE(8) MCCNN_FUSED_CASE(10) MCCNN_FUSED_CASE(13) MCCNN_FUSED_CASE(16) MCCNN_FUSED_CASE(20)
        MCCNN_FUSED_CASE(25) MCCNN_FUSED_CASE(32)
#undef MCCNN_FUSED_CASE
    }
    set_error("sgm (fused): D=%d exceeds the supported maximum of 1024", D);
    return MCCNN_EINVAL;
}

}  // namespace mccnn
