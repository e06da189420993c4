// Dependent-issue latencies of the FP64 building blocks ba_solve / marg_jacobi chain on (SM cycles per dependent operation, one warp).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe scripts/lat_probe.cu ; run on the B200
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int N = 512;
__global__ void probe(double *out, long long *cyc, double seed) {
    __shared__ double sm[1024];
    const int t = threadIdx.x;
    sm[t] = seed + t, sm[t + 256] = seed * 0.5 + t, sm[t + 512] = 1.0, sm[t + 768] = 2.0;
    __syncthreads();
    double x = seed + 1e-3 * t, y = 1.000001, acc = 0;
    long long t0, t1;
    // 0: DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, y, 1e-9);
    t1 = clock64(); if (t == 0) cyc[0] = t1 - t0; acc += x;
    // 1: DMUL
    x = seed; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64(); if (t == 0) cyc[1] = t1 - t0; acc += x;
    // 2: rsqrt(double)
    x = seed + 2.0; t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) x = rsqrt(x) + 1.5;
    t1 = clock64(); if (t == 0) cyc[2] = t1 - t0; acc += x;   // includes one DADD per step
    // 3: sqrt(double)
    x = seed + 2.0; t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) x = sqrt(x) + 1.5;
    t1 = clock64(); if (t == 0) cyc[3] = t1 - t0; acc += x;
    // 4: 1.0 / x
    x = seed + 2.0; t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) x = 1.0 / x + 1.5;
    t1 = clock64(); if (t == 0) cyc[4] = t1 - t0; acc += x;
    // 5: a / b general
    x = seed + 2.0; t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) x = y / x + 1.5;
    t1 = clock64(); if (t == 0) cyc[5] = t1 - t0; acc += x;
    // 6: DMMA dependent (accumulator chain)
    double c0 = 0, c1 = 0; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) dmma884(c0, c1, x, y);
    t1 = clock64(); if (t == 0) cyc[6] = t1 - t0; acc += c0 + c1;
    // 7: shared-memory pointer chase (LDS.64 dependent)
    int idx = t & 255; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = ((int) sm[idx + 512] + idx) & 255;   // sm[512..767] = 1.0
    t1 = clock64(); if (t == 0) cyc[7] = t1 - t0; acc += idx;             // includes F2I + IADD + LOP
    // 8: shfl (double = 2 x SHFL) dependent
    x = seed; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    t1 = clock64(); if (t == 0) cyc[8] = t1 - t0; acc += x;               // includes one DADD
    // 9: __syncthreads (blockDim threads)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64(); if (t == 0) cyc[9] = t1 - t0;
    // 10: DADD
    x = seed; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + y;
    t1 = clock64(); if (t == 0) cyc[10] = t1 - t0; acc += x;
    // 11: FFMA (reference)
    float f = (float) seed, g = 1.000001f; t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, g, 1e-9f);
    t1 = clock64(); if (t == 0) cyc[11] = t1 - t0; acc += f;
    // 12: rsqrt via MUFU.RSQ (float) + 2 Newton steps in double
    x = seed + 2.0; t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) {
        double r = (double) rsqrtf((float) x);
        r = r * (1.5 - 0.5 * x * r * r);
        r = r * (1.5 - 0.5 * x * r * r);
        x = r + 1.5;
    }
    t1 = clock64(); if (t == 0) cyc[12] = t1 - t0; acc += x;
    // 13: the 8 x 8 in-register Cholesky of ba_solve's diagonal tile (every lane the same code on broadcast shared-memory loads), repeated
    {
        __syncthreads();
        if (t < 64) {
            for (int a = 0; a < 8; a++) sm[a * 8 + (t & 7)] = (a == (t & 7)) ? 40.0 + a : 1.0 / (1.0 + a + (t & 7));
        }
        __syncthreads();
        constexpr int NB = 8, REP = 32;
        double sink = 0;
        t0 = clock64();
        for (int rep = 0; rep < REP; rep++) {
            double Ld[NB][NB], dinv[NB];
#pragma unroll
            for (int a = 0; a < NB; a++)
#pragma unroll
                for (int b = 0; b < NB; b++) Ld[a][b] = b <= a ? sm[a * 8 + b] + sink * 1e-300 : 0.0;
#pragma unroll
            for (int j = 0; j < NB; j++) {
                double d = Ld[j][j];
#pragma unroll
                for (int k = 0; k < j; k++) d -= Ld[j][k] * Ld[j][k];
                const double di = rsqrt(d);
                dinv[j] = di;
                Ld[j][j] = d * di;
#pragma unroll
                for (int a = j + 1; a < NB; a++) {
                    double sum = Ld[a][j];
#pragma unroll
                    for (int k = 0; k < j; k++) sum -= Ld[a][k] * Ld[j][k];
                    Ld[a][j] = sum * di;
                }
            }
            sink += Ld[7][7] + dinv[3] + Ld[7][0];
        }
        t1 = clock64(); if (t == 0) cyc[13] = (t1 - t0) * (long long) N / REP; acc += sink;   // scaled so that the print divides by N
    }
    out[blockIdx.x * blockDim.x + t] = acc;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * 4096); cudaMallocManaged(&cyc, sizeof(long long) * 16);
    const char *nm[14] = {"DFMA", "DMUL", "rsqrt(double)+DADD", "sqrt(double)+DADD", "1.0/x+DADD", "y/x+DADD", "DMMA.884 (acc chain)", "LDS.64 chase (+F2I,IADD,LOP)",
                          "shfl double + DADD", "__syncthreads", "DADD", "FFMA", "rsqrtf + 2 Newton (double) + DADD", "8x8 register Cholesky (whole block)"};
    for (int threads : {32, 256}) {
        probe<<<1, threads>>>(out, cyc, 1.25);
        cudaDeviceSynchronize();
        printf("-- %d threads in the CTA (one CTA on the SM): cycles per dependent operation\n", threads);
        for (int k = 0; k < 14; k++) printf("  %-34s %8.1f\n", nm[k], (double) cyc[k] / N);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
