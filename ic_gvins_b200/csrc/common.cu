// common.cu -- error reporting, launch counter, TMA descriptor encoding.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace icg {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

cudaError_t raise_dynamic_smem(const void *func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> seen;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = seen[std::make_pair(dev, func)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled) p;
    });
    return fn;
}

int encode_tensor_map_u8_3d(CUtensorMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                            uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available (driver too old / no device)");
        return ICG_ENODEVICE;
    }
    cuuint64_t dims[3]    = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3]     = {b0, b1, b2};
    cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: CUresult %d (dims %llu %llu %llu strides %llu %llu box %u %u %u)", (int) r,
                  (unsigned long long) d0, (unsigned long long) d1, (unsigned long long) d2, (unsigned long long) stride1_bytes,
                  (unsigned long long) stride2_bytes, b0, b1, b2);
        return ICG_ECUDA;
    }
    return ICG_OK;
}

}  // namespace icg

extern "C" {
const char *icg_last_error(void) { return icg::g_err; }
int icg_version(void) { return 100; }
uint64_t icg_launch_count(void) { return icg::g_launches.load(); }
void icg_launch_count_reset(void) { icg::g_launches.store(0); }
}
