// s3d_tiny.cuh -- the last octaves of the pyramid in ONE launch.
//
// At MNI size octaves 4 and 5 hold 11 x 13 x 11 and 5 x 6 x 5 voxels.  As five blur levels of two kernels each plus a
// subsample per octave they are 26 launches of 2-6 us that do nothing but wait for each other: 120 us on the dependency
// chain of every volume (profiles/r2_timeline_concurrent_b.txt) and a quarter of its launches.  Here one CTA walks
// every pass of every level of every such octave -- x, y, z (+ DoG) passes separated by block barriers, then the 2x
// subsample that feeds the next octave -- in 60 us, off the critical path (the volume's longest chain is octave 0).
//
// Arithmetic: the scalar form of the reference loop (GaussBlur3D.cpp:43-61, 329-479), the same expressions as the
// any-radius kernels of s3d_voxel.cuh: sum = 0.0f; sum = sum + w[j] * v for j = 0..2R, v = 0.0f outside the volume;
// DoG = prev + (-1) * g (fioMultSum); subsample = fioSubSampleInterpolate (FeatureIO.cpp:1474-1554).  Padding
// columns stay zero.
#pragma once
#include "s3d_voxel.cuh"

namespace s3d {

constexpr int kTinyThreads = 1024;
constexpr int kTinyMaxOct = 8;
constexpr long long kTinyMaxElems = 4096;      // pitch * Y * Z of an octave this kernel takes

struct TinyOct {
    int X, Y, Z, pitch;
    float *g[6];          // g[0] is complete when the kernel starts (first octave) or is produced by the subsample below
    float *d[5];
};
struct TinyDesc {
    int n_oct;
    TinyOct o[kTinyMaxOct];
    int ntaps[5];
    float taps[5][2 * kMaxFastR + 1];
};

// one separable pass over the whole octave, shared memory to shared memory; axis 0 / 1 / 2 = x / y / z; the z pass
// also writes the level and its DoG to global memory (g_out, g_dog; prev = the level this one was blurred from)
template <int AXIS>
__device__ __forceinline__ void tiny_pass(const float *in, float *out, const float *prev, float *g_out, float *g_dog,
                                          int X, int Y, int Z, int pitch, const float *w, int n)
{
    const int plane = pitch * Y, total = plane * Z, r = n / 2;
    for (int i = threadIdx.x; i < total; i += kTinyThreads) {
        const int x = i % pitch, y = (i / pitch) % Y, z = i / plane;
        float acc = 0.0f;
        if (AXIS == 0) {
            if (x < X) {
#pragma unroll 4
                for (int j = 0; j < n; j++) {
                    const int p = x + j - r;
                    const float v = (p >= 0 && p < X) ? in[i + j - r] : 0.0f;
                    acc = acc + w[j] * v;
                }
            }
        } else if (AXIS == 1) {
#pragma unroll 4
            for (int j = 0; j < n; j++) {
                const int p = y + j - r;
                const float v = (p >= 0 && p < Y) ? in[i + (j - r) * pitch] : 0.0f;
                acc = acc + w[j] * v;
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < n; j++) {
                const int p = z + j - r;
                const float v = (p >= 0 && p < Z) ? in[i + (j - r) * plane] : 0.0f;
                acc = acc + w[j] * v;
            }
        }
        out[i] = acc;
        if (AXIS == 2) {
            g_out[i] = acc;
            g_dog[i] = prev[i] + (-1.0f) * acc;
        }
    }
    __syncthreads();
}

// The octave lives in three shared-memory buffers of kTinyMaxElems floats that rotate through the roles
// (level j-1, scratch, level j): a pass reads shared memory only (the global buffers were just written by this CTA,
// every load from them would be an L2 round trip), the levels and DoGs go to global memory as plain stores.
__global__ void __launch_bounds__(kTinyThreads) tiny_octaves_kernel(const __grid_constant__ TinyDesc D)
{
    extern __shared__ __align__(16) float tiny_smem[];
    float *buf[3] = { tiny_smem, tiny_smem + kTinyMaxElems, tiny_smem + 2 * kTinyMaxElems };
    float *w = tiny_smem + 3 * kTinyMaxElems;                 // [5][2 * kMaxFastR + 1]
    constexpr int WN = 2 * kMaxFastR + 1;
    if (threadIdx.x < 5 * WN) w[threadIdx.x] = (&D.taps[0][0])[threadIdx.x];
    for (int oi = 0; oi < D.n_oct; oi++) {
        const TinyOct &o = D.o[oi];
        const int total = o.pitch * o.Y * o.Z;
        __syncthreads();
        for (int i = threadIdx.x; i < total; i += kTinyThreads) buf[0][i] = o.g[0][i];
        __syncthreads();
        int a = 0;                                            // buffer that holds level j-1
        for (int j = 1; j < 6; j++) {
            const int b = (a + 1) % 3, c = (a + 2) % 3;
            const int n = D.ntaps[j - 1];
            const float *wj = w + (j - 1) * WN;
            tiny_pass<0>(buf[a], buf[b], nullptr, nullptr, nullptr, o.X, o.Y, o.Z, o.pitch, wj, n);
            tiny_pass<1>(buf[b], buf[c], nullptr, nullptr, nullptr, o.X, o.Y, o.Z, o.pitch, wj, n);
            tiny_pass<2>(buf[c], buf[b], buf[a], o.g[j], o.d[j - 1], o.X, o.Y, o.Z, o.pitch, wj, n);
            a = b;
            if (j == 3 && oi + 1 < D.n_oct) {
                // level 0 of the next octave: 2x2x2 mean of level 3, the expression of subsample_kernel
                const TinyOct &nx = D.o[oi + 1];
                const float *lv = buf[a];
                const int oplane = nx.pitch * nx.Y, ototal = oplane * nx.Z;
                for (int i = threadIdx.x; i < ototal; i += kTinyThreads) {
                    const int x = i % nx.pitch, y = (i / nx.pitch) % nx.Y, z = i / oplane;
                    float rr = 0.0f;
                    if (x < nx.X) {
                        const float *p0 = lv + ((2 * z) * o.Y + 2 * y) * o.pitch + 2 * x;
                        const float *p1 = p0 + o.Y * o.pitch;
                        float s = 0.0f;
                        s = s + (((p0[0] + p0[o.pitch]) + p0[1]) + p0[o.pitch + 1]);
                        if (2 * z + 1 < o.Z) {
                            s = s + (((p1[0] + p1[o.pitch]) + p1[1]) + p1[o.pitch + 1]);
                            s = s * 0.125f;
                        } else {
                            s = s * 0.25f;
                        }
                        rr = s;
                    }
                    nx.g[0][i] = rr;
                }
            }
        }
    }
}
constexpr size_t kTinySmem = sizeof(float) * (3 * kTinyMaxElems + 5 * (2 * kMaxFastR + 1) + 3);

} // namespace s3d
