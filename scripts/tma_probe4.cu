#include <cstdio>
#include <cstdlib>
#include <vector>
#include <dlfcn.h>
#include "../ic_gvins_b200/csrc/common.cuh"
using namespace icg;
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int bytes, unsigned *out) {
    extern __shared__ __align__(1024) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    int lane = threadIdx.x;
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, bytes);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(buf)), "l"((uint64_t) &map), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = lane; i < bytes; i += 32) s += buf[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) *out = s;
}
int main(int argc, char **argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), pitch = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]), es = atoi(argv[6]);
    int c0 = atoi(argv[7]), c1 = atoi(argv[8]);
    uint8_t *d; cudaMalloc(&d, (size_t) pitch * H);
    std::vector<uint8_t> h((size_t) pitch * H); for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t) (i % 251);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN_encodeTiled enc = (PFN_encodeTiled) p;
    CUtensorMap map{};
    cuuint64_t size[2] = {(cuuint64_t) W, (cuuint64_t) H}; cuuint64_t stride[1] = {(cuuint64_t) pitch};
    cuuint32_t box[2] = {(cuuint32_t) bw, (cuuint32_t) bh}, est[2] = {1, 1};
    CUresult r = enc(&map, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    unsigned *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    int bytes = bw * bh * es;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 32, bytes>>>(map, c0, c1, bytes, out);
    cudaError_t e = cudaDeviceSynchronize(); unsigned res = 0; cudaMemcpy(&res, out, 4, cudaMemcpyDeviceToHost);
    unsigned ex = 0;
    for (int y = 0; y < bh; y++) for (int x = 0; x < bw * es; x++) { int X = c0 * es + x, Y = c1 + y; if (X >= 0 && X < W * es && Y >= 0 && Y < H) ex += h[(size_t) Y * pitch + X]; }
    printf("W=%d H=%d pitch=%d box=%dx%d es=%d at(%d,%d) enc=%d: %s sum=%u expected=%u\n", W, H, pitch, bw, bh, es, c0, c1, (int) r, cudaGetErrorString(e), res, ex);
    return 0;
}
