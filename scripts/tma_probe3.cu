#include <cstdio>
#include <cstdlib>
#include <vector>
#include <dlfcn.h>
#include <cuda/barrier>
#include "../ic_gvins_b200/csrc/common.cuh"
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
using namespace icg;
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// (1) 1-D bulk copy, no descriptor
__global__ void k_bulk1d(const uint8_t *src, unsigned *out) {
    __shared__ __align__(128) uint8_t buf[1024];
    __shared__ __align__(8) uint64_t bar;
    int lane = threadIdx.x;
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(&bar, 1024);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)),
                     "l"(src), "r"(1024), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = lane; i < 1024; i += 32) s += buf[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) *out = s;
}
// (2) NVIDIA programming-guide sample: int32 64x64 box
constexpr int SM_H = 64, SM_W = 64;
__global__ void k_sample(const __grid_constant__ CUtensorMap tensor_map, int x, int y, unsigned *out) {
    __shared__ alignas(128) int smem_buffer[SM_H][SM_W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    unsigned s = 0;
    for (int i = threadIdx.x; i < SM_H * SM_W; i += blockDim.x) s += (unsigned) smem_buffer[i / SM_W][i % SM_W];
    atomicAdd(out, s);
}
int main(int argc, char **argv) {
    int mode = atoi(argv[1]);
    unsigned *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    if (mode == 1) {
        uint8_t *d; cudaMalloc(&d, 4096); cudaMemset(d, 3, 4096);
        k_bulk1d<<<1, 32>>>(d, out);
        cudaError_t e = cudaDeviceSynchronize(); unsigned r = 0; cudaMemcpy(&r, out, 4, cudaMemcpyDeviceToHost);
        printf("bulk1d: %s sum=%u expected=3072\n", cudaGetErrorString(e), r); return 0;
    }
    const int GW = 1024, GH = 1024;
    int *d; cudaMalloc(&d, sizeof(int) * GW * GH);
    std::vector<int> h((size_t) GW * GH); for (size_t i = 0; i < h.size(); i++) h[i] = (int) (i % 1000);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    PFN_encodeTiled enc = nullptr;
    if (mode == 2) { void *p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (PFN_encodeTiled) p; }
    else { void *lib = dlopen("libcuda.so.1", RTLD_NOW); enc = (PFN_encodeTiled) dlsym(lib, "cuTensorMapEncodeTiled"); }
    CUtensorMap map{};
    cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SM_W, SM_H}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d; desc:", (int) r);
    const uint32_t *w = (const uint32_t *) &map; for (int i = 0; i < 32; i++) printf(" %08x", w[i]); printf("\n  base=%p\n", (void *) d);
    k_sample<<<1, 128>>>(map, 64, 128, out);
    cudaError_t e = cudaDeviceSynchronize(); unsigned res = 0; cudaMemcpy(&res, out, 4, cudaMemcpyDeviceToHost);
    unsigned ex = 0; for (int y = 0; y < 64; y++) for (int x = 0; x < 64; x++) ex += (unsigned) h[(size_t) (128 + y) * GW + 64 + x];
    printf("sample mode %d: %s sum=%u expected=%u\n", mode, cudaGetErrorString(e), res, ex);
    return 0;
}
