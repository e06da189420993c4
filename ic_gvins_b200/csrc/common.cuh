// common.cuh -- shared helpers for libicgvins_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/icgvins_b200.h"

namespace icg {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t) n, std::memory_order_relaxed); }
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-function, per-device setting shared by every handle of the process: it may only GROW
// (a handle with smaller capacities must not lower the limit under a live handle with larger ones -- launches of the latter would fail with
// "invalid argument").  Records the largest request per (device, function) and raises the attribute when needed.
cudaError_t raise_dynamic_smem(const void *func, size_t bytes);

#define ICG_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            ::icg::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return ICG_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

#define ICG_CHECK_LAUNCH()                                                                          \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ != cudaSuccess) {                                                                    \
            ::icg::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return ICG_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

// ---------------------------------------------------------------- TMA / mbarrier PTX wrappers (sm_90+/sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled TMA load: box -> shared, completion on mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"((uint64_t) map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t) map) : "memory");
}

// host: encode a tiled tensor map through the driver entry point (no link-time dependency on libcuda)
int encode_tensor_map_u8_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                            uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);

}  // namespace icg
