// fp64_probe.cu -- measured FP64 rates on this part: vector DFMA vs tensor DMMA (mma.sync m8n8k4 / m16n8k8 f64).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_probe scripts/fp64_probe.cu ; prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * b + c, a1 = a1 * b + c, a2 = a2 * b + c, a3 = a3 * b + c, a4 = a4 * b + c, a5 = a5 * b + c, a6 = a6 * b + c, a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void k_dmma884(double *out, int iters) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int i = 0; i < iters; i++) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
__global__ void k_dmma1688(double *out, int iters) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double a[4] = {threadIdx.x * 1e-3, 1, 2, 3}, b[2] = {1.0 + threadIdx.x * 1e-6, 0.5};
    for (int i = 0; i < iters; i++) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c0[0]), "+d"(c0[1]), "+d"(c0[2]), "+d"(c0[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c1[0]), "+d"(c1[1]), "+d"(c1[2]), "+d"(c1[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c0[2] + c0[3] + c1[0] + c1[1] + c1[2] + c1[3];
}
template <typename F>
static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main() {
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    const int iters = 20000, blocks = 148 * 4, threads = 512;
    float t = timeit([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("DFMA vector   : %.2f TFLOP/s\n", 2.0 * 8 * iters * (double) blocks * threads / t / 1e9);
    t = timeit([&] { k_dmma884<<<blocks, threads>>>(out, iters); });
    printf("DMMA m8n8k4   : %.2f TFLOP/s\n", 2.0 * 256 * 4 * iters * (double) blocks * (threads / 32) / t / 1e9);
    t = timeit([&] { k_dmma1688<<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k8  : %.2f TFLOP/s\n", 2.0 * 1024 * 2 * iters * (double) blocks * (threads / 32) / t / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
