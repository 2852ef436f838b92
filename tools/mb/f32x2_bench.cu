// Microbenchmark (profiling tool, not product code): issue rate of scalar FMUL/FADD against packed
// FFMA2/FADD2 on sm_100a, alone and mixed with shared-memory loads.  Answers: does packed fp32x2 math raise
// the FP32 floor of the blur kernels, or only free issue slots?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_bench f32x2_bench.cu
// Note: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad false; the product is
// therefore written as fma.rn.f32x2(a, w, nz) with nz = -0.0f passed at run time (bit-identical to the rounded
// product, opaque to ptxas), the sum as add.rn.f32x2.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 pmul(float2 a, float w, float nz)
{
    float2 r;
    asm("{ .reg .b64 ra, rw, rz, rp; mov.b64 ra, {%2,%3}; mov.b64 rw, {%4,%4}; mov.b64 rz, {%5,%5};"
        " fma.rn.f32x2 rp, ra, rw, rz; mov.b64 {%0,%1}, rp; }" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(w), "f"(nz));
    return r;
}
constexpr int ACC = 16;   // floats of accumulator per thread

// mode 0: scalar FMUL + FADD ; 1: packed FFMA2 + FADD2 ; 2: scalar FADD only ; 3: FADD2 only ; 4: FFMA2 only ; 5: FMUL only
template <int MODE, int LDS_PER_ITER>
__global__ void __launch_bounds__(256) k(float *out, float w, float nz, int iters)
{
    __shared__ float4 sm[256];
    sm[threadIdx.x] = make_float4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    float a[ACC];
    float v[ACC];
#pragma unroll
    for (int i = 0; i < ACC; i++) { a[i] = i; v[i] = threadIdx.x * 1e-3f + i; }
    float4 ld = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    unsigned base = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x & 31]);
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int l = 0; l < LDS_PER_ITER; l++) {
            float4 u;
            asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(u.x), "=f"(u.y), "=f"(u.z), "=f"(u.w) : "r"(base + l * 512));
        }
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < ACC; i++) { v[i] = __fmul_rn(w, v[i]); a[i] = __fadd_rn(a[i], v[i]); }
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < ACC; i += 2) {
                float2 p = pmul(make_float2(v[i], v[i + 1]), w, nz);
                float2 s = __fadd2_rn(make_float2(a[i], a[i + 1]), p);
                a[i] = s.x; a[i + 1] = s.y; v[i] = p.x; v[i + 1] = p.y;
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = __fadd_rn(a[i], v[i]);
        } else if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < ACC; i += 2) {
                float2 s = __fadd2_rn(make_float2(a[i], a[i + 1]), make_float2(v[i], v[i + 1]));
                a[i] = s.x; a[i + 1] = s.y;
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < ACC; i += 2) {
                float2 s = pmul(make_float2(a[i], a[i + 1]), w, nz);
                a[i] = s.x; a[i + 1] = s.y;
            }
        } else {
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = __fmul_rn(a[i], w);
        }
    }
    float s = ld.x + idx;
    for (int i = 0; i < ACC; i++) s += a[i] + v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int L>
static void run(const char *name, float *d, int ctas_per_sm, double flops_per_iter_thread)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MODE, L><<<148 * ctas_per_sm, 256>>>(d, 1.0001f, -0.0f, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double threads = 148.0 * ctas_per_sm * 256;
    double lane_ops = threads * iters * flops_per_iter_thread;             // scalar-equivalent FP32 operations
    printf("%-34s ctas/SM %d  lds/iter %d : %8.3f ms  %7.2f Tlane-op/s  = %6.1f lane-ops/clk/SM @%d MHz nominal\n", name, ctas_per_sm, L, best,
           lane_ops / best / 1e9, lane_ops / (best * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
}

int main()
{
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    for (int c = 2; c <= 8; c *= 2) {
        run<0, 0>("scalar FMUL+FADD", d, c, 2.0 * ACC);
        run<1, 0>("packed FFMA2(+-0)+FADD2", d, c, 2.0 * ACC);
        run<2, 0>("scalar FADD", d, c, 1.0 * ACC);
        run<3, 0>("packed FADD2", d, c, 1.0 * ACC);
        run<4, 0>("packed FFMA2 (dependent chain x8)", d, c, 1.0 * ACC);
        run<5, 0>("scalar FMUL (dependent chain x16)", d, c, 1.0 * ACC);
    }
    run<0, 4>("scalar FMUL+FADD + 4 LDS.128", d, 8, 2.0 * ACC);
    run<1, 4>("packed + 4 LDS.128", d, 8, 2.0 * ACC);
    run<0, 8>("scalar FMUL+FADD + 8 LDS.128", d, 8, 2.0 * ACC);
    run<1, 8>("packed + 8 LDS.128", d, 8, 2.0 * ACC);
    return 0;
}
