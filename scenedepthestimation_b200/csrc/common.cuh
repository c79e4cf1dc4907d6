// Shared helpers for the mccnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <mutex>
#include "../../include/mccnn_b200.h"

namespace mccnn {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MCCNN_REQUIRE(cond, code, ...)            \
    do {                                          \
        if (!(cond)) {                            \
            ::mccnn::set_error(__VA_ARGS__);      \
            return (code);                        \
        }                                         \
    } while (0)

#define MCCNN_CUDA(expr)                                              \
    do {                                                              \
        cudaError_t _e = (expr);                                      \
        if (_e != cudaSuccess) return ::mccnn::cuda_fail(_e, #expr);  \
    } while (0)

#define MCCNN_LAUNCH_CHECK(name)                                        \
    do {                                                                \
        cudaError_t _e = cudaGetLastError();                            \
        if (_e != cudaSuccess) return ::mccnn::cuda_fail(_e, name);     \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int disp_pitch(int D) { return (D + 3) & ~3; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int sm_count();
int current_device();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the occupancy query cost several microseconds each: done once per
// (kernel instantiation, device, shared-memory size) instead of on every launch. Kern is a template ARGUMENT so that every
// kernel owns its cache (kernels of one signature share a function-pointer type).
template <auto Kern>
inline int kernel_setup(int threads, size_t smem, int* blocks_per_sm) {
    struct Slot { size_t smem; int per_sm; bool set; };
    static Slot slots[64] = {};
    static std::mutex mu;
    const int dev = current_device();
    std::lock_guard<std::mutex> lock(mu);
    Slot& s = slots[dev & 63];
    if (!s.set || s.smem != smem) {
        MCCNN_CUDA(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        MCCNN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, Kern, threads, smem));
        s.smem = smem; s.per_sm = per_sm; s.set = true;
    }
    if (blocks_per_sm) *blocks_per_sm = s.per_sm;
    return 0;
}

// ------------------------------------------------------------------ PTX: mbarrier + bulk async copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map); bytes % 16 == 0, 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// The same two operations issued from WARP-CONVERGENT code: every lane executes the instruction, elect.sync picks the lane that
// issues it. Inside an `if (lane == 0)` region the compiler wraps each uniform-datapath instruction (UBLKCP, UTCHMMA, ...) in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop, about ten dependent instructions per copy on the issuing warp's critical path.
__device__ __forceinline__ void mbar_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_elect(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
        "}" ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g_commit_elect(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t"
        "@q cp.async.bulk.commit_group;\n\t"
        "}" ::"l"(gmem_dst),
        "r"(smem_u32(smem_src)), "r"(bytes)
        : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read_elect() {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.wait_group.read %0;\n\t"
        "}" ::"n"(N)
        : "memory");
}
// 1-D bulk copy shared -> global.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace mccnn
