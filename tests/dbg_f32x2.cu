// microbenchmark: do packed fp32x2 FMUL2/FADD2 (sm_100) issue faster than scalar FMUL/FADD pairs?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ACC = 16;
__global__ void scalar_k(float *out, float w, int iters)
{
    float a[ACC], v = threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < ACC; i++) a[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) a[i] = __fadd_rn(a[i], __fmul_rn(w, v + i));
    }
    float s = 0; for (int i = 0; i < ACC; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void packed_k(float *out, float w, int iters)
{
    float2 a[ACC / 2]; float v = threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < ACC / 2; i++) a[i] = make_float2(2 * i, 2 * i + 1);
    float2 ww = make_float2(w, w);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ACC / 2; i++) a[i] = __fadd2_rn(a[i], __fmul2_rn(ww, make_float2(v + 2 * i, v + 2 * i + 1)));
    }
    float s = 0; for (int i = 0; i < ACC / 2; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); scalar_k<<<148 * 8, 256>>>(d, 1.0001f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = 148.0 * 8 * 256 * (double)iters * ACC * 2;
        printf("scalar FMUL+FADD : %.3f ms  %.1f Tflop-instr/s\n", ms, ops / ms / 1e9);
        cudaEventRecord(e0); packed_k<<<148 * 8, 256>>>(d, 1.0001f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FMUL2+FADD2: %.3f ms  %.1f Tflop-instr/s\n", ms, ops / ms / 1e9);
    }
    return 0;
}
