// fp64_rate.cu — development microbenchmark: throughput of scalar double-precision add / mul / fma on sm_100a (B200). The Brownian
// generator (AS241 in double, no FMA) and the double-then-round transcendentals are bound by this pipe, not by HBM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 2048;
template <int OP> __global__ void k(double* out, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
            else if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(b));
            else asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(a));
        }
    }
    double s = 0; for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    const int grid = sms * 8, thr = 256;
    const double ops = (double)grid * thr * 8.0 * ITER;
    const char* names[3] = {"DFMA", "DADD", "DMUL"};
    float t[3];
    t[0] = timeit([&] { k<0><<<grid, thr>>>(out, 1.0000001, 0.5); });
    t[1] = timeit([&] { k<1><<<grid, thr>>>(out, 1.0000001, 0.5); });
    t[2] = timeit([&] { k<2><<<grid, thr>>>(out, 1.0000001, 0.5); });
    for (int i = 0; i < 3; i++) printf("%s: %7.3f ms  %6.2f lane-ops/clk/SM (at 1.965 GHz), %6.2f T ops/s\n", names[i], t[i], ops / (t[i] * 1e-3) / 1.965e9 / sms, ops / (t[i] * 1e-3) / 1e12);
    return 0;
}
