// TMA probe: which tensor-map / box combinations does UTMALDG.3D accept on this part?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ic_gvins_b200/csrc/common.cuh"
using namespace icg;
struct Maps { CUtensorMap m[2]; };
__global__ void probe(const __grid_constant__ Maps maps, int which, int c0, int c1, int c2, int bytes, unsigned *out) {
    __shared__ __align__(128) uint8_t buf[48 * 32];
    __shared__ __align__(8) uint64_t bar;
    int lane = threadIdx.x;
    if (lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();
        mbar_expect_tx(&bar, bytes);
        tma_load_3d(buf, &maps.m[which], c0, c1, c2, &bar);
    }
    mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = lane; i < bytes; i += 32) s += buf[i];
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) *out = s;
}
int main(int argc, char **argv) {
    int W = atoi(argv[1]), H = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), c0 = atoi(argv[5]), c1 = atoi(argv[6]);
    int which = argc > 7 ? atoi(argv[7]) : 0;
    int pitch = (W + 15) & ~15, slots = 4;
    uint8_t *d; cudaMalloc(&d, (size_t) pitch * H * slots);
    std::vector<uint8_t> h((size_t) pitch * H * slots);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t) (i % 251);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    Maps maps;
    int rc = encode_tensor_map_u8_3d(&maps.m[0], d, W, H, slots, pitch, (uint64_t) pitch * H, bw, bh, 1);
    rc |= encode_tensor_map_u8_3d(&maps.m[1], d, W, H, slots, pitch, (uint64_t) pitch * H, bw, bh, 1);
    if (rc) { printf("encode failed: %s\n", icg_last_error()); return 1; }
    unsigned *out; cudaMalloc(&out, 4);
    probe<<<1, 32>>>(maps, which, c0, c1, 1, bw * bh, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned r = 0; cudaMemcpy(&r, out, 4, cudaMemcpyDeviceToHost);
    // expected
    unsigned ex = 0;
    for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) { int X = c0 + x, Y = c1 + y; if (X >= 0 && X < W && Y >= 0 && Y < H) ex += h[(size_t) pitch * H + (size_t) Y * pitch + X]; }
    printf("W=%d H=%d box=%dx%d at (%d,%d) map %d: %s  sum=%u expected=%u\n", W, H, bw, bh, c0, c1, which, cudaGetErrorString(e), r, ex);
    return 0;
}
