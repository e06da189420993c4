// TMA probe 2: NVIDIA's documented libcu++ path (cuda::barrier + cde::cp_async_bulk_tensor_2d_global_to_shared), 2-D,
// descriptor in __grid_constant__ param, vs descriptor in global memory with my raw PTX.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda/barrier>
#include "../ic_gvins_b200/csrc/common.cuh"
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
using namespace icg;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_libcu(const __grid_constant__ CUtensorMap map, int c0, int c1, unsigned *out) {
    __shared__ alignas(128) uint8_t buf[32 * 32];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &map, c0, c1, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    unsigned s = 0;
    for (int i = threadIdx.x; i < 1024; i += 32) s += buf[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (threadIdx.x == 0) *out = s;
}
__global__ void k_raw2d(const CUtensorMap *map, int c0, int c1, unsigned *out) {
    __shared__ __align__(128) uint8_t buf[32 * 32];
    __shared__ __align__(8) uint64_t bar;
    int lane = threadIdx.x;
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, 1024);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(buf)), "l"((uint64_t) map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = lane; i < 1024; i += 32) s += buf[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) *out = s;
}
int main(int argc, char **argv) {
    int mode = atoi(argv[1]);
    int W = 160, H = 70, pitch = 160;
    uint8_t *d; cudaMalloc(&d, (size_t) pitch * H);
    std::vector<uint8_t> h((size_t) pitch * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t) (i % 251);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled) p;
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t) W, (cuuint64_t) H}; cuuint64_t strides[1] = {(cuuint64_t) pitch};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, mode >= 10 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d\n", (int) r, (int) q);
    unsigned *out; cudaMalloc(&out, 4);
    CUtensorMap *dmap; cudaMalloc(&dmap, sizeof(map)); cudaMemcpy(dmap, &map, sizeof(map), cudaMemcpyHostToDevice);
    if (mode % 10 == 0) k_libcu<<<1, 32>>>(map, 10, 10, out);
    else k_raw2d<<<1, 32>>>(dmap, 10, 10, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned res = 0; cudaMemcpy(&res, out, 4, cudaMemcpyDeviceToHost);
    unsigned ex = 0;
    for (int y = 0; y < 32; y++) for (int x = 0; x < 32; x++) ex += h[(size_t) (10 + y) * pitch + 10 + x];
    printf("mode %d: %s sum=%u expected=%u\n", mode, cudaGetErrorString(e), res, ex);
    return 0;
}
