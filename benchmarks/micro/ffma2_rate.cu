// ffma2_rate.cu — development microbenchmark: issue rate of packed fp32 (FFMA2 / FMUL2 / FADD2) against scalar FFMA on sm_100a,
// alone and mixed with integer instructions (does a packed instruction free an issue slot?).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

__global__ void k_scalar(float* out, float a, float b) {
    float x[16];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
    }
    float s = 0; for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
    unsigned long long x[8], pa, pb;
    asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
    for (int i = 0; i < 8; i++) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(x[i]) : "f"(v)); }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(pa), "l"(pb));
    }
    float s = 0;
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the same number of fp32 operations with 16 integer adds per iteration interleaved (the interpreter's overhead instructions)
__global__ void k_scalar_mix(float* out, float a, float b, int c) {
    float x[16]; int y[16];
    for (int i = 0; i < 16; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(y[i]) : "r"(c));
        }
    }
    float s = 0; for (int i = 0; i < 16; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed_mix(float* out, float a, float b, int c) {
    unsigned long long x[8], pa, pb; int y[16];
    asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
    for (int i = 0; i < 8; i++) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(x[i]) : "f"(v)); }
    for (int i = 0; i < 16; i++) y[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(pa), "l"(pb));
            asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(y[2 * i]) : "r"(c));
            asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(y[2 * i + 1]) : "r"(c));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    for (int i = 0; i < 16; i++) s += y[i];
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
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int grid = sms * 8, thr = 256;      // 64 warps per SM
    const double fmas = (double)grid * thr * 16.0 * ITER;
    float t1 = timeit([&] { k_scalar<<<grid, thr>>>(out, 1.0001f, 0.5f); });
    float t2 = timeit([&] { k_packed<<<grid, thr>>>(out, 1.0001f, 0.5f); });
    float t3 = timeit([&] { k_scalar_mix<<<grid, thr>>>(out, 1.0001f, 0.5f, 12345); });
    float t4 = timeit([&] { k_packed_mix<<<grid, thr>>>(out, 1.0001f, 0.5f, 12345); });
    printf("SMs %d\n", sms);
    printf("scalar FFMA        : %7.3f ms  %6.1f fma/clk/SM (at 1.965 GHz)\n", t1, fmas / (t1 * 1e-3) / 1.965e9 / sms);
    printf("packed FFMA2       : %7.3f ms  %6.1f fma/clk/SM\n", t2, fmas / (t2 * 1e-3) / 1.965e9 / sms);
    printf("scalar FFMA + LOP3 : %7.3f ms  %6.1f fma/clk/SM\n", t3, fmas / (t3 * 1e-3) / 1.965e9 / sms);
    printf("packed FFMA2 + LOP3: %7.3f ms  %6.1f fma/clk/SM\n", t4, fmas / (t4 * 1e-3) / 1.965e9 / sms);
    return 0;
}
