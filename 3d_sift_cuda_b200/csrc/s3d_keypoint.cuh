// s3d_keypoint.cuh -- per-keypoint stages for sm_100a: deferred validation + sub-voxel refinement,
// 11^3 patch gather, normalisation, structure-tensor eigen-orientation, canonical orientation
// histograms, SIFT-Rank / BRIEF-family descriptors, rank transform.
//
// These stages are CPU-only in the reference; here each keypoint (or feature row) is
// one CTA.  The arithmetic follows the reference operation for operation so the output is
// bit-identical: per-sample work (trilinear gathers, gradients, blurs, contributions) is done by all
// threads, while the few order-sensitive fp32 accumulations (patch mean / energy, structure tensor,
// histogram splats, descriptor bins) are walked in the reference's raster order by the thread that
// owns the accumulator.  Compiled with -fmad=false; divisions and square roots are IEEE (nvcc default).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "../../include/s3d.h"

namespace s3d {

constexpr int PD = 11;
constexpr int PV = 1331;
constexpr int kMaxOct = 12;
constexpr int kMaxRowsPerKp = 12;   // 1 non-reoriented + at most 11 orientations (reference MultiScale.cpp:2875, 2966)
constexpr int kMaxSphere = 640;

struct OctaveDesc {
    int X, Y, Z, pitch;
    int z_off, Zg, own0, own1;   // slab mode: buffer plane 0 = global plane z_off of an octave of depth Zg;
                                 // candidates only from local planes [own0, own1).  Whole volume: 0, Z, 0, Z.
    const float *g[6];
    const float *d[5];
    float sigma[6];
};
struct PyramidDesc {
    int n_oct;
    OctaveDesc oct[kMaxOct];
};

// Small tables shared by all keypoint kernels; filled once per context (s3d_engine.cu).
struct KpTables {
    int n_sphere;                         // voxels with dx^2+dy^2+dz^2 < 25, raster order
    unsigned short sphere[kMaxSphere];
    int n_hist_taps;                      // sigma 0.5 (fBlurGradOriHist, reference MultiScale.cpp:37)
    float hist_taps[9];
    int n_brief_taps;                     // sigma 0.95 (reference MultiScale.cpp:1032)
    float brief_taps[9];
    float desc_w[PD];                     // weight of the lower spatial bin per patch coordinate
    unsigned short brief_a[64], brief_b[64];   // linear patch offsets of the 64 BRIEF pairs
    float brief_dist[64];                 // (int)|a-b| as float, NRRIEF divisor
};
__constant__ KpTables c_tab;

enum { ERR_CAND_OVERFLOW = 1, ERR_KP_OVERFLOW = 2, ERR_ROW_OVERFLOW = 4 };

// Optional per-phase cycle counters (profiling builds only: -DS3D_PHASE_TIMERS); thread 0 of each CTA
// accumulates clock64() deltas so the split of a keypoint kernel's latency can be read back.
#ifdef S3D_PHASE_TIMERS
__device__ unsigned long long g_phase[32];
#define PHASE_INIT() long long ph_t_ = clock64()
#define PHASE(i)                                                                   \
    do {                                                                           \
        if (threadIdx.x == 0) {                                                    \
            long long n_ = clock64();                                              \
            atomicAdd(&g_phase[i], (unsigned long long)(n_ - ph_t_));              \
            ph_t_ = n_;                                                            \
        }                                                                          \
    } while (0)
#else
#define PHASE_INIT()
#define PHASE(i)
#endif

// ------------------------------------------------------------------------------------------------
// scalar helpers (reference MultiScale.cpp:1614-1697, 2531-2534; FeatureIO.cpp:757-850)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double finddet(double a1, double a2, double a3, double b1, double b2, double b3,
                                          double c1, double c2, double c3)
{
    return ((a1 * b2 * c3) - (a1 * b3 * c2) - (a2 * b1 * c3) + (a3 * b1 * c2) + (a2 * b3 * c1) - (a3 * b2 * c1));
}

__device__ double interp_quadratic(double x0, double x1, double x2, double fx0, double fx1, double fx2)
{
    if (!(fx1 < fx0 && fx1 < fx2) && !(fx1 > fx0 && fx1 > fx2)) return x1;
    double a1 = x0 * x0, b1 = x0, c1 = 1;
    double a2 = x1 * x1, b2 = x1, c2 = 1;
    double a3 = x2 * x2, b3 = x2, c3 = 1;
    double det = finddet(a1, a2, a3, b1, b2, b3, c1, c2, c3);
    double detx = finddet(fx0, fx1, fx2, b1, b2, b3, c1, c2, c3);
    double dety = finddet(a1, a2, a3, fx0, fx1, fx2, c1, c2, c3);
    if (fx0 == 0 && fx1 == 0 && fx2 == 0) return x1;
    if (det != 0) {
        if (detx != 0) return dety / (-2.0 * detx);
    }
    return x1;
}

__device__ __forceinline__ void interp_coord(float fX, float fMaxX, int &iX, float &fW)
{
    if (fX < 0.5f) {
        iX = 0;
        fW = 1.0f;
    } else if (fX >= fMaxX - 0.5f) {
        iX = (int)(fMaxX - 2);
        fW = 0.0f;
    } else {
        float fMinusHalf = fX - 0.5f;
        iX = (int)floorf(fMinusHalf);
        fW = 1.0f - (fMinusHalf - ((float)iX));
    }
}

__device__ __forceinline__ float trilinear_get(const float *__restrict__ img, int X, int Y, int Z, int pitch,
                                               float x, float y, float z)
{
    int iX, iY, iZ;
    float wx, wy, wz;
    interp_coord(x, (float)X, iX, wx);
    interp_coord(y, (float)Y, iY, wy);
    interp_coord(z, (float)Z, iZ, wz);
    const float *p = img + ((long long)iZ * Y + iY) * pitch + iX;
    long long plane = (long long)pitch * Y;
    float f000 = __ldg(p), f100 = __ldg(p + 1), f010 = __ldg(p + pitch), f110 = __ldg(p + pitch + 1);
    float f001 = __ldg(p + plane), f101 = __ldg(p + plane + 1), f011 = __ldg(p + plane + pitch), f111 = __ldg(p + plane + pitch + 1);
    float fn00 = wx * f000 + (1.0f - wx) * f100;
    float fn01 = wx * f001 + (1.0f - wx) * f101;
    float fn10 = wx * f010 + (1.0f - wx) * f110;
    float fn11 = wx * f011 + (1.0f - wx) * f111;
    float fnn0 = wy * fn00 + (1.0f - wy) * fn10;
    float fnn1 = wy * fn01 + (1.0f - wy) * fn11;
    return wz * fnn0 + (1.0f - wz) * fnn1;
}

// Isotropic resampling of an anisotropic input (section 8(f) N1; reference featExtract.cpp:118-204): output voxel
// (x,y,z) = fioGetPixelTrilinearInterp(in, x*rf0 + 0.5, y*rf1 + 0.5, z*rf2 + 0.5) with the reference's types:
// int * float is a float product, + 0.5 is a double sum, and the sum is narrowed to the float parameter.
__global__ void resample_iso_kernel(const float *__restrict__ in, int X, int Y, int Z, int pitch,
                                    float *__restrict__ out, int nX, int nY, int nZ, int opitch, float rf0, float rf1, float rf2)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, z = blockIdx.z;
    if (x >= opitch || y >= nY || z >= nZ) return;
    float v = 0.0f;
    if (x < nX)
        v = trilinear_get(in, X, Y, Z, pitch, (float)((double)((float)x * rf0) + 0.5), (float)((double)((float)y * rf1) + 0.5),
                          (float)((double)((float)z * rf2) + 0.5));
    out[((long long)z * nY + y) * opitch + x] = v;
}

// invert_3x3<float,double> (reference MultiScale.h:192-222)
__device__ void invert3(const float *m, float *o)
{
    float a11 = m[0], a12 = m[1], a13 = m[2];
    float a21 = m[3], a22 = m[4], a23 = m[5];
    float a31 = m[6], a32 = m[7], a33 = m[8];
    float det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
    double div = 1 / (double)det;
    o[0] = (float)((a33 * a22 - a32 * a23) * div);
    o[3] = (float)(-(a33 * a21 - a31 * a23) * div);
    o[6] = (float)((a32 * a21 - a31 * a22) * div);
    o[1] = (float)(-(a33 * a12 - a32 * a13) * div);
    o[4] = (float)((a33 * a11 - a31 * a13) * div);
    o[7] = (float)(-(a32 * a11 - a31 * a12) * div);
    o[2] = (float)((a23 * a12 - a22 * a13) * div);
    o[5] = (float)(-(a23 * a11 - a21 * a13) * div);
    o[8] = (float)((a22 * a11 - a21 * a12) * div);
}

__device__ __forceinline__ void vec_norm(float *p)
{
    float fSumSqr = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (fSumSqr > 0) {
        float fDiv = (float)(1.0 / (double)sqrtf(fSumSqr));
        p[0] *= fDiv; p[1] *= fDiv; p[2] *= fDiv;
    } else {
        p[0] = 1; p[1] = 0; p[2] = 0;
    }
}
__device__ __forceinline__ float vec_mag(const float *p)
{
    float fSumSqr = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    return fSumSqr > 0 ? sqrtf(fSumSqr) : 0.0f;
}
__device__ __forceinline__ float vec_dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ------------------------------------------------------------------------------------------------
// block-cooperative pieces (blockDim.x threads, all must call)
// ------------------------------------------------------------------------------------------------
constexpr int PVP = 1332;   // patch arrays are padded to a multiple of 4 floats so each starts 16-byte aligned

// sampleImage3D (reference MultiScale.cpp:2614-2714): inv = inverse orientation, already in smem.
// A thread takes U samples per round; the round is fully unrolled and branch-light so the 8 corner loads of all U
// samples are in flight together (the gather is pure latency otherwise).  U = 2: with 4 the gather of one keypoint is
// 6 % shorter but orient_patch_kernel needs 108 registers instead of 63, and a batch pays for the registers the tail
// CTAs hold (444.8 against 432.5 us per volume, same box; U = 1: 48 registers, 432.8 us).
template <int U = 2>
__device__ void gather_patch(const float *__restrict__ img, int X, int Y, int Zg, int z_off, int pitch,
                             float fx, float fy, float fz, float scale, const float *inv, float *patch)
{
    const float fImageRad = 2.0f * scale;
    const float fScale = fImageRad / (float)(PD / 2);
    const float m00 = inv[0], m01 = inv[1], m02 = inv[2], m10 = inv[3], m11 = inv[4], m12 = inv[5], m20 = inv[6], m21 = inv[7], m22 = inv[8];
    const long long plane = (long long)pitch * Y;
    for (int i0 = threadIdx.x; i0 < PV; i0 += U * blockDim.x) {
        float v[U][8], w[U][3];
        bool inside[U], live[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int i = i0 + u * blockDim.x;
            live[u] = i < PV;
            int ii = live[u] ? i : 0;
            float f0 = (float)(ii % PD - 5), f1 = (float)((ii / PD) % PD - 5), f2 = (float)(ii / (PD * PD) - 5);
            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
            p0 = p0 + m00 * f0; p0 = p0 + m01 * f1; p0 = p0 + m02 * f2;
            p1 = p1 + m10 * f0; p1 = p1 + m11 * f1; p1 = p1 + m12 * f2;
            p2 = p2 + m20 * f0; p2 = p2 + m21 * f1; p2 = p2 + m22 * f2;
            p0 = p0 * fScale; p1 = p1 * fScale; p2 = p2 * fScale;
            p0 = p0 + fx; p1 = p1 + fy; p2 = p2 + fz;
            inside[u] = !(p0 < 0 || p0 >= X);
            int iX, iY, iZ;
            interp_coord(p0, (float)X, iX, w[u][0]);
            interp_coord(p1, (float)Y, iY, w[u][1]);
            interp_coord(p2, (float)Zg, iZ, w[u][2]);      // global depth; the buffer starts at global plane z_off
            if (!inside[u]) iX = 0;   // value unused; keep the address legal
            const float *p = img + ((long long)(iZ - z_off) * Y + iY) * pitch + iX;
            v[u][0] = __ldg(p); v[u][1] = __ldg(p + 1); v[u][2] = __ldg(p + pitch); v[u][3] = __ldg(p + pitch + 1);
            v[u][4] = __ldg(p + plane); v[u][5] = __ldg(p + plane + 1); v[u][6] = __ldg(p + plane + pitch); v[u][7] = __ldg(p + plane + pitch + 1);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            float wx = w[u][0], wy = w[u][1], wz = w[u][2];
            float fn00 = wx * v[u][0] + (1.0f - wx) * v[u][1];
            float fn01 = wx * v[u][4] + (1.0f - wx) * v[u][5];
            float fn10 = wx * v[u][2] + (1.0f - wx) * v[u][3];
            float fn11 = wx * v[u][6] + (1.0f - wx) * v[u][7];
            float fnn0 = wy * fn00 + (1.0f - wy) * fn10;
            float fnn1 = wy * fn01 + (1.0f - wy) * fn11;
            float r = wz * fnn0 + (1.0f - wz) * fnn1;
            if (live[u]) patch[i0 + u * blockDim.x] = inside[u] ? r : 0.0f;
        }
    }
}

// left-to-right fp32 sum of n floats in shared memory (16-byte aligned, n padded reads allowed up to PVP)
__device__ __forceinline__ float seq_sum(const float *a, int n)
{
    float s = 0.0f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float4 u = *reinterpret_cast<const float4 *>(a + i), v = *reinterpret_cast<const float4 *>(a + i + 4);
        s = s + u.x; s = s + u.y; s = s + u.z; s = s + u.w;
        s = s + v.x; s = s + v.y; s = s + v.z; s = s + v.w;
    }
    for (; i < n; i++) s = s + a[i];
    return s;
}

// Feature3D::NormalizeData (reference MultiScale.cpp:127-205).  scratch: PVP floats, red: 2 floats.
__device__ void normalize_patch(float *patch, float *scratch, float *red)
{
    __syncthreads();
    if (threadIdx.x == 0) red[0] = seq_sum(patch, PV) / (float)(PD * PD * PD);
    __syncthreads();
    float mean = red[0];
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {
        float v = patch[i] - mean;
        patch[i] = v;
        scratch[i] = v * v;
    }
    __syncthreads();
    if (threadIdx.x == 0) red[1] = 1.0f / sqrtf(seq_sum(scratch, PV));
    __syncthreads();
    float fDiv = red[1];
    for (int i = threadIdx.x; i < PV; i += blockDim.x) patch[i] = patch[i] * fDiv;
    __syncthreads();
}

// fioGenerateEdgeImages3D on the patch (reference FeatureIO.cpp:2284-2326)
__device__ void patch_gradients(const float *p, float *dx, float *dy, float *dz)
{
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {
        int x = i % PD, y = (i / PD) % PD, z = i / (PD * PD);
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        if (x >= 1 && x < PD - 1 && y >= 1 && y < PD - 1 && z >= 1 && z < PD - 1) {
            gx = p[i + 1] - p[i - 1];
            gy = p[i + PD] - p[i - PD];
            gz = p[i + PD * PD] - p[i - PD * PD];
        }
        dx[i] = gx; dy[i] = gy; dz[i] = gz;
    }
    __syncthreads();
}

// blur_3d_simpleborders on an 11^3 image (reference GaussBlur3D.cpp:329-479): src -> dst, tmp scratch;
// src is left intact; src, tmp, dst must be distinct.  taps: shared or constant memory.
__device__ void blur_patch(const float *src, float *tmp, float *dst, const float *taps, int ntaps)
{
    int r = ntaps / 2;
    __syncthreads();
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {   // x: src -> dst
        int x = i % PD;
        float s = 0.0f;
        for (int j = 0; j < ntaps; j++) { int p = x + j - r; float v = (p >= 0 && p < PD) ? src[i + j - r] : 0.0f; s = s + taps[j] * v; }
        dst[i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {   // y: dst -> tmp
        int y = (i / PD) % PD;
        float s = 0.0f;
        for (int j = 0; j < ntaps; j++) { int p = y + j - r; float v = (p >= 0 && p < PD) ? dst[i + (j - r) * PD] : 0.0f; s = s + taps[j] * v; }
        tmp[i] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {   // z: tmp -> dst
        int z = i / (PD * PD);
        float s = 0.0f;
        for (int j = 0; j < ntaps; j++) { int p = z + j - r; float v = (p >= 0 && p < PD) ? tmp[i + (j - r) * PD * PD] : 0.0f; s = s + taps[j] * v; }
        dst[i] = s;
    }
    __syncthreads();
}

// The same blur for 3 taps (the orientation histograms: sigma 0.5) in ONE pass without scratch: an output voxel
// recomputes the x sums of its 9 (y,z) neighbours and the y sums of its 3 z neighbours in registers.  Every sum is the
// expression of the three-pass form (0.0f first, taps left to right, out-of-range inputs contribute taps[j] * 0.0f,
// which leaves a sum that started at +0.0 unchanged), so the bits are the same; two block barriers and two
// shared-memory round trips are gone.
__device__ void blur_patch3(const float *src, float *dst, const float *taps)
{
    const float t0 = taps[0], t1 = taps[1], t2 = taps[2];
    __syncthreads();
    for (int i = threadIdx.x; i < PV; i += blockDim.x) {
        const int x = i % PD, y = (i / PD) % PD, z = i / (PD * PD);
        float zs = 0.0f;
#pragma unroll
        for (int jz = 0; jz < 3; jz++) {
            const int pz = z + jz - 1;
            float vz = 0.0f;
            if (pz >= 0 && pz < PD) {
                float ys = 0.0f;
#pragma unroll
                for (int jy = 0; jy < 3; jy++) {
                    const int py = y + jy - 1;
                    float vy = 0.0f;
                    if (py >= 0 && py < PD) {
                        const float *row = src + (pz * PD + py) * PD;
                        float xs = 0.0f;
                        xs = xs + t0 * (x >= 1 ? row[x - 1] : 0.0f);
                        xs = xs + t1 * row[x];
                        xs = xs + t2 * (x + 1 < PD ? row[x + 1] : 0.0f);
                        vy = xs;
                    }
                    ys = ys + (jy == 0 ? t0 : jy == 1 ? t1 : t2) * vy;
                }
                vz = ys;
            }
            zs = zs + (jz == 0 ? t0 : jz == 1 ? t1 : t2) * vz;
        }
        dst[i] = zs;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// SingularValueDecomp<float,3,3> + SortEigenDecomp (reference SVD.h:15-228): float storage, double scalars.
// ------------------------------------------------------------------------------------------------
#define S3D_SIGN(a, b) ((b) >= 0.0 ? fabs(a) : -fabs(a))
#define S3D_PYTHAG(a, b) (sqrt((a) * (a) + (b) * (b)))

__device__ void svd3(float (*mat)[3], float *w, float (*v)[3], double *rv1)
{
    const int m = 3, n = 3;
    int flag, i, its, j, jj, k, l = 0, nm = 0;
    double anorm, c, f, g, h, s, scale, x, y, z;
    g = scale = anorm = 0.0;
    for (i = 1; i <= n; i++) {
        l = i + 1;
        rv1[i - 1] = scale * g;
        g = s = scale = 0.0;
        if (i <= m) {
            for (k = i; k <= m; k++) scale += fabs((double)mat[k - 1][i - 1]);
            if (scale) {
                for (k = i; k <= m; k++) {
                    mat[k - 1][i - 1] = (float)(mat[k - 1][i - 1] / scale);
                    s += (double)(mat[k - 1][i - 1] * mat[k - 1][i - 1]);
                }
                f = mat[i - 1][i - 1];
                g = -S3D_SIGN(sqrt(s), f);
                h = f * g - s;
                mat[i - 1][i - 1] = (float)(f - g);
                for (j = l; j <= n; j++) {
                    for (s = 0.0, k = i; k <= m; k++) s += (double)(mat[k - 1][i - 1] * mat[k - 1][j - 1]);
                    f = s / h;
                    for (k = i; k <= m; k++) mat[k - 1][j - 1] = (float)(mat[k - 1][j - 1] + f * mat[k - 1][i - 1]);
                }
                for (k = i; k <= m; k++) mat[k - 1][i - 1] = (float)(mat[k - 1][i - 1] * scale);
            }
        }
        w[i - 1] = (float)(scale * g);
        g = s = scale = 0.0;
        if (i <= m && i != n) {
            for (k = l; k <= n; k++) scale += fabs((double)mat[i - 1][k - 1]);
            if (scale) {
                for (k = l; k <= n; k++) {
                    mat[i - 1][k - 1] = (float)(mat[i - 1][k - 1] / scale);
                    s += (double)(mat[i - 1][k - 1] * mat[i - 1][k - 1]);
                }
                f = mat[i - 1][l - 1];
                g = -S3D_SIGN(sqrt(s), f);
                h = f * g - s;
                mat[i - 1][l - 1] = (float)(f - g);
                for (k = l; k <= n; k++) rv1[k - 1] = mat[i - 1][k - 1] / h;
                for (j = l; j <= m; j++) {
                    for (s = 0.0, k = l; k <= n; k++) s += (double)(mat[j - 1][k - 1] * mat[i - 1][k - 1]);
                    for (k = l; k <= n; k++) mat[j - 1][k - 1] = (float)(mat[j - 1][k - 1] + s * rv1[k - 1]);
                }
                for (k = l; k <= n; k++) mat[i - 1][k - 1] = (float)(mat[i - 1][k - 1] * scale);
            }
        }
        {
            double t = fabs((double)w[i - 1]) + fabs(rv1[i - 1]);
            anorm = (anorm > t ? anorm : t);
        }
    }
    for (i = n; i >= 1; i--) {
        if (i < n) {
            if (g) {
                for (j = l; j <= n; j++) v[j - 1][i - 1] = (float)((mat[i - 1][j - 1] / mat[i - 1][l - 1]) / g);
                for (j = l; j <= n; j++) {
                    for (s = 0.0, k = l; k <= n; k++) s += (double)(mat[i - 1][k - 1] * v[k - 1][j - 1]);
                    for (k = l; k <= n; k++) v[k - 1][j - 1] = (float)(v[k - 1][j - 1] + s * v[k - 1][i - 1]);
                }
            }
            for (j = l; j <= n; j++) v[i - 1][j - 1] = v[j - 1][i - 1] = 0.0f;
        }
        v[i - 1][i - 1] = 1.0f;
        g = rv1[i - 1];
        l = i;
    }
    for (i = n; i >= 1; i--) {
        l = i + 1;
        g = w[i - 1];
        for (j = l; j <= n; j++) mat[i - 1][j - 1] = 0.0f;
        if (g) {
            g = 1.0 / g;
            for (j = l; j <= n; j++) {
                for (s = 0.0, k = l; k <= m; k++) s += (double)(mat[k - 1][i - 1] * mat[k - 1][j - 1]);
                f = (s / mat[i - 1][i - 1]) * g;
                for (k = i; k <= m; k++) mat[k - 1][j - 1] = (float)(mat[k - 1][j - 1] + f * mat[k - 1][i - 1]);
            }
            for (j = i; j <= m; j++) mat[j - 1][i - 1] = (float)(mat[j - 1][i - 1] * g);
        } else {
            for (j = i; j <= m; j++) mat[j - 1][i - 1] = 0.0f;
        }
        mat[i - 1][i - 1] = mat[i - 1][i - 1] + 1;
    }
    for (k = n; k >= 1; k--) {
        for (its = 1; its <= 30; its++) {
            flag = 1;
            for (l = k; l >= 1; l--) {
                nm = l - 1;
                if ((double)(fabs(rv1[l - 1]) + anorm) == anorm) { flag = 0; break; }
                if ((double)(fabs((double)w[nm - 1]) + anorm) == anorm) break;
            }
            if (flag) {
                c = 0.0;
                s = 1.0;
                for (i = l; i <= k; i++) {
                    f = s * rv1[i - 1];
                    rv1[i - 1] = c * rv1[i - 1];
                    if ((double)(fabs(f) + anorm) == anorm) break;
                    g = w[i - 1];
                    h = S3D_PYTHAG(f, g);
                    w[i - 1] = (float)h;
                    h = 1.0 / h;
                    c = g * h;
                    s = -f * h;
                    for (j = 1; j <= m; j++) {
                        y = mat[j - 1][nm - 1];
                        z = mat[j - 1][i - 1];
                        mat[j - 1][nm - 1] = (float)(y * c + z * s);
                        mat[j - 1][i - 1] = (float)(z * c - y * s);
                    }
                }
            }
            z = w[k - 1];
            if (l == k) {
                if (z < 0.0) {
                    w[k - 1] = (float)(-z);
                    for (j = 1; j <= n; j++) v[j - 1][k - 1] = -v[j - 1][k - 1];
                }
                break;
            }
            x = w[l - 1];
            nm = k - 1;
            y = w[nm - 1];
            g = rv1[nm - 1];
            h = rv1[k - 1];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
            g = S3D_PYTHAG(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + S3D_SIGN(g, f))) - h)) / x;
            c = s = 1.0;
            for (j = l; j <= nm; j++) {
                i = j + 1;
                g = rv1[i - 1];
                y = w[i - 1];
                h = s * g;
                g = c * g;
                z = S3D_PYTHAG(f, h);
                rv1[j - 1] = z;
                c = f / z;
                s = h / z;
                f = x * c + g * s;
                g = g * c - x * s;
                h = y * s;
                y *= c;
                for (jj = 1; jj <= n; jj++) {
                    x = v[jj - 1][j - 1];
                    z = v[jj - 1][i - 1];
                    v[jj - 1][j - 1] = (float)(x * c + z * s);
                    v[jj - 1][i - 1] = (float)(z * c - x * s);
                }
                z = S3D_PYTHAG(f, h);
                w[j - 1] = (float)z;
                if (z) {
                    z = 1.0 / z;
                    c = f * z;
                    s = h * z;
                }
                f = c * g + s * y;
                x = c * y - s * g;
                for (jj = 1; jj <= m; jj++) {
                    y = mat[jj - 1][j - 1];
                    z = mat[jj - 1][i - 1];
                    mat[jj - 1][j - 1] = (float)(y * c + z * s);
                    mat[jj - 1][i - 1] = (float)(z * c - y * s);
                }
            }
            rv1[l - 1] = 0.0;
            rv1[k - 1] = f;
            w[k - 1] = (float)x;
        }
    }
}

__device__ void sort_eigen(float *w, float (*v)[3])
{
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (w[i] < w[j]) {
                float t = w[j]; w[j] = w[i]; w[i] = t;
                for (int k = 0; k < 3; k++) { t = v[k][j]; v[k][j] = v[k][i]; v[k][i] = t; }
            }
}


// ------------------------------------------------------------------------------------------------
// Candidate stage.  detect_kernel (s3d_voxel.cuh) appends candidates of every (octave, centre level,
// min/max) list in atomic order.  cand_refine_kernel then handles every candidate of every list in
// parallel: its raster rank inside its list (rank by counting; keys are unique voxel indices), the
// deferred validation against the coarser DoG, the parabola refinement and the support-box test
// (reference MultiScale.cpp:424-455, 1135-1318, 1372-1386, 2633-2643); survivors are staged at
// stage[list][rank].  compact_kernel makes the ordered keypoint array (octave, level, min then max,
// raster order = the reference's output order).
// ------------------------------------------------------------------------------------------------
struct ListDesc {            // list = (octave*3 + (c-1))*2 + is_max
    int n_lists, cap;
    const s3d_cand *raw;     // [n_lists][cap]
    const int *counts;       // [n_lists]
};

// One WARP per candidate: the lanes split the rank count over the list (coalesced 16-byte loads) and the 27
// neighbours of the deferred validation; lanes 0..3 then run the four parabola fits (x, y, z, scale) side by
// side.  (One thread per candidate made this kernel a 15-20 us latency chain: n sequential loads per thread.)
// Large volumes: rank by counting over the whole list is O(n^2) (1.0 s of a 1.2 s extraction at 1024^3 with 2e5
// candidates per list).  cand_bucket_kernel first groups a list by z plane -- one CTA per list: plane histogram in
// shared memory, exclusive scan, scatter (order inside a plane is arbitrary) -- and writes the plane offsets; the
// refinement kernel then counts only inside the candidate's plane: rank = plane_off[z] + #(same plane, smaller (y,x)).
__global__ void __launch_bounds__(1024) cand_bucket_kernel(ListDesc L, int list_begin, int Z, s3d_cand *__restrict__ sorted,
                                                           int *__restrict__ plane_off, int plane_stride)
{
    extern __shared__ int s_hist[];            // [Z + 1] counts -> exclusive offsets; then used as cursors
    __shared__ int s_warp[32];
    const int list = list_begin + blockIdx.x;
    const int n = min(L.counts[list], L.cap);
    const s3d_cand *raw = L.raw + (long long)list * L.cap;
    s3d_cand *dst = sorted + (long long)list * L.cap;
    int *poff = plane_off + (long long)list * plane_stride;
    for (int z = threadIdx.x; z <= Z; z += blockDim.x) s_hist[z] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) atomicAdd(&s_hist[raw[k].z], 1);
    __syncthreads();
    // exclusive scan over the Z + 1 plane counts, 1024 elements per round
    int carry = 0;
    for (int base = 0; base <= Z; base += blockDim.x) {
        const int z = base + threadIdx.x;
        const int v = z <= Z ? s_hist[z] : 0;
        int x = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < wid) woff += s_warp[w]; tot += s_warp[w]; }
        if (z <= Z) { const int ex = carry + woff + x - v; s_hist[z] = ex; poff[z] = ex; }
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) poff[Z + 1] = carry;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const s3d_cand c = raw[k];
        dst[atomicAdd(&s_hist[c.z], 1)] = c;
    }
}

__global__ void __launch_bounds__(256) cand_refine_kernel(const __grid_constant__ PyramidDesc pyr, ListDesc L, int list_begin,
                                                          s3d_keypoint *__restrict__ stage, unsigned char *__restrict__ flags,
                                                          int *err, const s3d_cand *__restrict__ sorted, const int *__restrict__ plane_off,
                                                          int plane_stride)
{
    const int list = list_begin + blockIdx.x;
    const int octave = list / 6, c = (list / 2) % 3 + 1, is_max = list & 1;
    const OctaveDesc &o = pyr.oct[octave];
    const int X = o.X, Y = o.Y, pitch = o.pitch;
    const long long plane = (long long)pitch * Y;
    int n = L.counts[list];
    if (n > L.cap) { if (threadIdx.x == 0 && blockIdx.y == 0) atomicOr(err, ERR_CAND_OVERFLOW); n = L.cap; }
    const s3d_cand *raw = L.raw + (long long)list * L.cap;
    const float *dH = o.d[c - 1], *dC = o.d[c], *dL = o.d[c + 1];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int k = blockIdx.y * warps_per_block + (threadIdx.x >> 5); k < n; k += gridDim.y * warps_per_block) {
        int rank;
        s3d_cand cd;
        if (sorted) {      // list grouped by plane: count inside the candidate's plane only
            const s3d_cand *srt = sorted + (long long)list * L.cap;
            const int *poff = plane_off + (long long)list * plane_stride;
            cd = srt[k];
            const int b0 = poff[cd.z], b1 = poff[cd.z + 1];
            const int key = cd.y * X + cd.x;
            int cnt = 0;
            for (int j = b0 + lane; j < b1; j += 32) {
                const s3d_cand q = srt[j];
                cnt += (q.y * X + q.x) < key;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            rank = b0 + cnt;
        } else {
            cd = raw[k];
            const long long key = ((long long)cd.z * Y + cd.y) * X + cd.x;
            int cnt = 0;
            for (int j = lane; j < n; j += 32) {
                const s3d_cand q = raw[j];
                cnt += (((long long)q.z * Y + q.y) * X + q.x) < key;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            rank = cnt;
        }
        const long long i = (long long)cd.z * plane + (long long)cd.y * pitch + cd.x;
        const float cv = cd.value;
        bool mine = true;
        if (lane < 27) {
            const int dz = lane / 9 - 1, dy = (lane / 3) % 3 - 1, dx = lane % 3 - 1;
            const float a = dL[i + dz * plane + dy * pitch + dx];
            mine = is_max ? (a < cv) : (a > cv);
        }
        const bool ok = __all_sync(0xffffffffu, mine);
        bool valid = false;
        s3d_keypoint kp;
        if (ok) {
            // geometry is computed in GLOBAL plane coordinates (slab mode: local z + z_off), because the
            // parabola arithmetic and later float sums depend on the magnitude of z
            const int gz = cd.z + o.z_off;
            float r = 0.0f;
            if (lane == 0) r = (float)interp_quadratic(cd.x - 1, cd.x, cd.x + 1, dC[i - 1], dC[i], dC[i + 1]);
            else if (lane == 1) r = (float)interp_quadratic(cd.y - 1, cd.y, cd.y + 1, dC[i - pitch], dC[i], dC[i + pitch]);
            else if (lane == 2) r = (float)interp_quadratic(gz - 1, gz, gz + 1, dC[i - plane], dC[i], dC[i + plane]);
            else if (lane == 3) r = (float)(2 * interp_quadratic(o.sigma[c - 1], o.sigma[c], o.sigma[c + 1], dH[i], dC[i], dL[i]));
            float fx = __shfl_sync(0xffffffffu, r, 0), fy = __shfl_sync(0xffffffffu, r, 1), fz = __shfl_sync(0xffffffffu, r, 2);
            const float scale = __shfl_sync(0xffffffffu, r, 3);
            fx += 0.5f; fy += 0.5f; fz += 0.5f;
            float fImageRad = 2.0f * scale;
            int iRadMax = (int)(fImageRad + 2);
            bool oob = (fx - iRadMax < 0 || fy - iRadMax < 0 || fz - iRadMax < 0 ||
                        fx + iRadMax >= X || fy + iRadMax >= Y || fz + iRadMax >= o.Zg);
            if (!oob) {
                valid = true;
                kp.octave = octave; kp.level = c; kp.is_max = is_max;
                kp.ix = cd.x; kp.iy = cd.y; kp.iz = gz;
                kp.x = fx; kp.y = fy; kp.z = fz; kp.scale = scale;
            }
        }
        if (lane == 0) {
            long long slot = (long long)list * L.cap + rank;
            flags[slot] = valid ? 1 : 0;
            if (valid) stage[slot] = kp;
        }
    }
}

// One CTA: ordered compaction of the staged keypoints over the concatenation of all lists.
__global__ void __launch_bounds__(1024) compact_kernel(ListDesc L, const s3d_keypoint *__restrict__ stage,
                                                       const unsigned char *__restrict__ flags,
                                                       s3d_keypoint *__restrict__ kps, int *kp_count, int kp_cap, int *err)
{
    __shared__ int s_pref[kMaxOct * 6 + 1];
    __shared__ int s_warp[32];
    __shared__ int s_run;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int l = 0; l < L.n_lists; l++) { s_pref[l] = acc; acc += min(L.counts[l], L.cap); }
        s_pref[L.n_lists] = acc;
        s_run = 0;
    }
    __syncthreads();
    const int total = s_pref[L.n_lists];
    for (int start = 0; start < total; start += blockDim.x) {
        int g = start + threadIdx.x;
        bool valid = false;
        long long slot = 0;
        if (g < total) {
            int lo = 0, hi = L.n_lists - 1;     // last list with s_pref[l] <= g
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_pref[mid] <= g) lo = mid; else hi = mid - 1; }
            slot = (long long)lo * L.cap + (g - s_pref[lo]);
            valid = flags[slot] != 0;
        }
        unsigned bal = __ballot_sync(0xffffffffu, valid);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < wid) woff += s_warp[w]; tot += s_warp[w]; }
        int pos = s_run + woff + __popc(bal & ((1u << lane) - 1));
        if (valid) {
            if (pos < kp_cap) kps[pos] = stage[slot];
            else atomicOr(err, ERR_KP_OVERFLOW);
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *kp_count = min(s_run, kp_cap);
}

// ------------------------------------------------------------------------------------------------
// Orientation assignment (reference generateFeature3D MultiScale.cpp:1705-1862, determineOrientation3D
// :2541-2607, determineCanonicalOrientation3D :2722-3037), split BY PHASE so that work which parity forces onto
// one thread never parks a whole CTA and its shared memory (round 1: one 256-thread CTA with 60 KB per keypoint,
// 100 K cycles of which 50 K waited for one warp's sequential histogram splat and one thread's SVD):
//   orient_patch_kernel  one CTA per keypoint: identity patch, NormalizeData, sphere gradients, structure tensor
//   orient_svd_kernel    one THREAD per keypoint: NR SVD in double + eigenvalue test -> kept keypoints
//   orient_hist_kernel   one CTA per kept keypoint: primary direction histogram, peaks -> up to 11 primary directions
//   orient_b_kernel      one CTA per (keypoint, primary direction): secondary histogram, peaks -> rotations
// The reference caps the total at 11 rotations per keypoint, taken in (primary, secondary) order; that
// truncation is applied when rows are counted (row_offsets_kernel).
// ------------------------------------------------------------------------------------------------
constexpr int kSph = 485;            // voxels of the 11^3 patch with dx^2+dy^2+dz^2 < 25 (the host checks c_tab.n_sphere)
constexpr int kSphP = 488;           // padded to a multiple of 4 and of the 8 splat warps' ranges
constexpr int kBinsP = 1332;         // 11^3 histogram bins, padded
constexpr int kHistThreads = 256, kHistWarps = kHistThreads / 32;
static_assert(kSphP <= kMaxSphere, "sphere table too small");

struct PatchSmem {                   // orient_patch_kernel
    float patch[PVP], scratch[PVP];
    float ex[kSphP], ey[kSphP], ez[kSphP];
    float inv[9];
    float red[2];
};

struct HistSmem {                    // orient_hist_kernel, orient_b_kernel
    union {                          // lifetimes: patch (until the contributions exist) -> sorted (splat) -> h1, h2 (blur, peaks)
        struct { float patch[PVP], h1[PVP], h2[PVP]; } p;
        float sorted[kSphP * 8];
    } a;
    union {                          // splat terms of a voxel (until the splat has sorted them) -> the histogram -> peak lists
        float4 cw[kSphP];            // value and the three lower-corner weights: its 8 contributions are formed when they are scattered
        float h0[PVP];
        struct { s3d_cand peaks[128], psort[128]; unsigned char pflag[736]; } k;
    } b;
    int cbase[kSphP];
    unsigned char cnt[kHistWarps][kBinsP];   // entries per (warp range, bin), then (low byte of) their exclusive prefix over the ranges
    unsigned char lrank[kSphP * 8];          // rank of an entry among the entries of its warp range that hit the same bin
    unsigned short start[kBinsP + 4];        // first sorted slot of every bin
    unsigned short hi[kBinsP + 4];           // bits 8-9 of the prefixes, 2 bits per warp range
    int warp_tot[kHistWarps];
    float taps[12];
    int np, nprim;
};

// Histogram splat (fioIncPixelTrilinearInterp, reference FeatureIO.cpp:853-889, called voxel after voxel in sphere
// order).  Every bin must receive its terms in voxel order (fp32 sums do not commute), but different bins are
// independent: the 8 x n_sphere (bin, value) entries are sorted by bin with a STABLE counting sort -- warp w ranks the
// entries of its contiguous voxel range with match.any, the per-range counts are prefix-summed over ranges and bins --
// and every bin is then summed left to right by one thread, 0.0f first like the zeroed histogram of the reference.
// Round 1 walked the voxels on one warp, one shared-memory read-modify-write round trip per voxel (20-45 K cycles per
// histogram); orientation histograms are peaked, so bin-owner schemes without the sort serialise on a few bins.
// All kHistThreads threads must call; hist is complete (and the block synchronised) on return.
#ifdef S3D_PHASE_TIMERS
#define SPLAT_ARGS , long long &ph_t_
#define SPLAT_PASS , ph_t_
#else
#define SPLAT_ARGS
#define SPLAT_PASS
#endif
// Splat terms of one voxel at histogram position (px,py,pz) with value v (fioIncPixelTrilinearInterp, reference
// FeatureIO.cpp:853-889): bin of the lower corner and (v, wx, wy, wz); its 8 contributions are the products below, formed
// when the splat scatters them (8 floats per voxel would be 15.6 KB of shared memory per CTA, and a batch pays for every
// kilobyte the tail CTAs hold: +16 KB here = -6 % throughput, profiles/README.md).
__device__ __forceinline__ void make_contrib(float px, float py, float pz, float v, float4 &cw, int &base)
{
    int iX, iY, iZ;
    float wx, wy, wz;
    interp_coord(px, (float)PD, iX, wx);
    interp_coord(py, (float)PD, iY, wy);
    interp_coord(pz, (float)PD, iZ, wz);
    base = (iZ * PD + iY) * PD + iX;
    cw = make_float4(v, wx, wy, wz);
}
__device__ __forceinline__ void contrib_values(const float4 cw, float *c8)
{
    const float v = cw.x, wx = cw.y, wy = cw.z, wz = cw.w;
    c8[0] = v * wx * wy * wz;
    c8[1] = v * (1.0f - wx) * wy * wz;
    c8[2] = v * wx * (1.0f - wy) * wz;
    c8[3] = v * (1.0f - wx) * (1.0f - wy) * wz;
    c8[4] = v * wx * wy * (1.0f - wz);
    c8[5] = v * (1.0f - wx) * wy * (1.0f - wz);
    c8[6] = v * wx * (1.0f - wy) * (1.0f - wz);
    c8[7] = v * (1.0f - wx) * (1.0f - wy) * (1.0f - wz);
}

__device__ __forceinline__ int splat_corner_offset(int c) { return (c & 1) + ((c >> 1) & 1) * PD + ((c >> 2) & 1) * PD * PD; }
// does the 2x2x2 footprint that starts at bin `base` cover bin `bin`?  (offsets 0, 1, 11, 12 and the same + 121)
__device__ __forceinline__ int splat_covers(int bin, int base)
{
    const int q = bin - base;
    const int r = q >= PD * PD ? q - PD * PD : q;
    return ((unsigned)r <= 12u) ? ((0x1803 >> r) & 1) : 0;
}

__device__ void splat_histogram_sorted(HistSmem &S, float *hist, int nsph SPLAT_ARGS)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kHistWarps * kBinsP / 4; i += kHistThreads) reinterpret_cast<unsigned int *>(&S.cnt[0][0])[i] = 0u;
    __syncthreads();
    // pass 1: 4 voxels x 8 corners per step, lanes in entry order.  A bin is hit at most once per voxel, so the rank
    // of an entry among the entries of its step is the number of EARLIER voxels of the step whose footprint covers
    // its bin (match.any would do, but its latency grows with the number of distinct values: 500+ cycles here).
    const int vr = (nsph + kHistWarps - 1) / kHistWarps;
    const int n0 = warp * vr, n1 = min(nsph, n0 + vr);
    const int j = lane >> 3;
    const int off = splat_corner_offset(lane & 7);
    unsigned char *cnt = S.cnt[warp];
    // (software pipelined: the footprint tests of the next step are done before the counter round trip of this one)
    int bin, rank, later;
    bool valid;
    auto prepare = [&](int nb) {
        const int n = nb + j;
        int base = n < n1 ? S.cbase[n] : -1;
        valid = base >= 0;
        if (!valid) base = -100000;                    // covers nothing
        bin = base + off;
        rank = 0; later = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int hit = splat_covers(bin, __shfl_sync(0xffffffffu, base, q * 8));
            if (q < j) rank += hit;
            if (q > j) later += hit;
        }
    };
    prepare(n0);
    for (int nb = n0; nb < n1; nb += 4) {
        const int c_bin = bin, c_rank = rank, c_later = later;
        const bool c_valid = valid;
        prepare(nb + 4);
        int old = 0;
        if (c_valid) old = cnt[c_bin];
        __syncwarp();
        if (c_valid) {
            S.lrank[(nb + j) * 8 + (lane & 7)] = (unsigned char)(old + c_rank);
            if (c_later == 0) cnt[c_bin] = (unsigned char)(old + c_rank + 1);     // last entry of its bin in this step
        }
        __syncwarp();
    }
    __syncthreads();
    PHASE(22);
    // per bin: exclusive prefix of the counts over the warp ranges (low byte in place, bits 8-9 packed in hi[]),
    // totals -> exclusive scan over the bins
    int tot[6], tsum = 0;
    const int b0 = t * 6;
    if (b0 < kBinsP) {
#pragma unroll
        for (int k = 0; k < 6; k++) {
            int run = 0, hi = 0;
#pragma unroll
            for (int w = 0; w < kHistWarps; w++) {
                const int c = S.cnt[w][b0 + k];
                S.cnt[w][b0 + k] = (unsigned char)run;
                hi |= (run >> 8) << (2 * w);
                run += c;
            }
            S.hi[b0 + k] = (unsigned short)hi;
            tot[k] = run;
            tsum += run;
        }
    }
    int inc = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
    if (lane == 31) S.warp_tot[warp] = inc;
    __syncthreads();
    int run = inc - tsum;
    for (int w = 0; w < warp; w++) run += S.warp_tot[w];
    if (b0 < kBinsP) {
#pragma unroll
        for (int k = 0; k < 6; k++) { S.start[b0 + k] = (unsigned short)run; run += tot[k]; }
    }
    __syncthreads();
    PHASE(23);
    // pass 2: a thread moves the 8 entries of a voxel to their slots = first slot of the bin + entries of the earlier
    // warp ranges + rank inside the range
    for (int n = t; n < nsph; n += kHistThreads) {
        const int base = S.cbase[n];
        if (base >= 0) {
            const int w = n / vr;
            float cv[8];
            contrib_values(S.b.cw[n], cv);
            const uint2 lr = *reinterpret_cast<const uint2 *>(&S.lrank[n * 8]);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const int bin = base + splat_corner_offset(c);
                const int rk = (int)(((c < 4 ? lr.x : lr.y) >> (8 * (c & 3))) & 0xffu);
                const int pre = (int)S.cnt[w][bin] | ((((int)S.hi[bin] >> (2 * w)) & 3) << 8);
                S.a.sorted[(int)S.start[bin] + pre + rk] = cv[c];
            }
        }
    }
    __syncthreads();
    PHASE(24);
    // every bin: left-to-right sum of its terms by one thread (bin PV of the padding has none: start[PV] is the total).
    // Orientation histograms are peaked -- a few bins receive a hundred terms and more -- so the walk is arranged to
    // run at the latency of the dependent adds: 128-bit loads from the first aligned slot on, issued one group ahead.
    for (int b = t; b < PV; b += kHistThreads) {
        int i = S.start[b];
        const int i1 = S.start[b + 1];
        float acc = 0.0f;
        for (; i < i1 && (i & 3); i++) acc = acc + S.a.sorted[i];
        if (i + 8 <= i1) {
            float4 u = *reinterpret_cast<const float4 *>(&S.a.sorted[i]), v = *reinterpret_cast<const float4 *>(&S.a.sorted[i + 4]);
            while (true) {
                const bool more = i + 16 <= i1;
                float4 nu = u, nv = v;
                if (more) { nu = *reinterpret_cast<const float4 *>(&S.a.sorted[i + 8]); nv = *reinterpret_cast<const float4 *>(&S.a.sorted[i + 12]); }
                acc = acc + u.x; acc = acc + u.y; acc = acc + u.z; acc = acc + u.w;
                acc = acc + v.x; acc = acc + v.y; acc = acc + v.z; acc = acc + v.w;
                i += 8;
                if (!more) break;
                u = nu; v = nv;
            }
        }
        for (; i < i1; i++) acc = acc + S.a.sorted[i];
        hist[b] = acc;
    }
    if (t == 0) hist[PV] = 0.0f;
    __syncthreads();
    PHASE(25);
}

// regFindFEATUREIOPeaks + lvSortHighLow on an 11^3 histogram (reference MultiScale.cpp:1987-2121,
// LocationValue.cpp:28-56).  All threads test the 729 interior bins; warp 0 compacts them in raster
// order; result in `sorted` (descending value, ties in raster order), count in *np.
__device__ void find_sort_peaks(const float *h, unsigned char *pflag, s3d_cand *raw, s3d_cand *sorted, int *np)
{
    for (int t = threadIdx.x; t < 729; t += blockDim.x) {
        int x = 1 + t % 9, y = 1 + (t / 9) % 9, z = 1 + t / 81;
        int i = (z * PD + y) * PD + x;
        float c = h[i];
        bool pk = true;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    if (dx == 0 && dy == 0 && dz == 0) continue;
                    pk = pk && (h[i + (dz * PD + dy) * PD + dx] < c);
                }
        pflag[t] = pk ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int lane = threadIdx.x, n = 0;
        for (int t0 = 0; t0 < 729; t0 += 32) {
            int t = t0 + lane;
            bool pk = (t < 729) && pflag[t];
            unsigned bal = __ballot_sync(0xffffffffu, pk);
            if (pk) {
                int pos = n + __popc(bal & ((1u << lane) - 1));
                int x = 1 + t % 9, y = 1 + (t / 9) % 9, z = 1 + t / 81;
                if (pos < 128) raw[pos] = s3d_cand{ x, y, z, h[(z * PD + y) * PD + x] };
            }
            n += __popc(bal);
        }
        if (lane == 0) *np = min(n, 128);
    }
    __syncthreads();
    int n = *np;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float v = raw[i].value;
        int rank = 0;
        for (int j = 0; j < n; j++) {
            float u = raw[j].value;
            rank += (u > v) || (u == v && j < i);
        }
        sorted[rank] = raw[i];
    }
    __syncthreads();
}

__device__ __forceinline__ void interp_point_patch(const float *h, int ix, int iy, int iz, float *o)
{
    int i = (iz * PD + iy) * PD + ix;
    o[0] = (float)interp_quadratic(ix - 1, ix, ix + 1, h[i - 1], h[i], h[i + 1]);
    o[1] = (float)interp_quadratic(iy - 1, iy, iy + 1, h[i - PD], h[i], h[i + PD]);
    o[2] = (float)interp_quadratic(iz - 1, iz, iz + 1, h[i - PD * PD], h[i], h[i + PD * PD]);
}

// central differences of a sphere voxel (fioGenerateEdgeImages3D, reference FeatureIO.cpp:2284-2326; sphere voxels
// are interior voxels of the patch: |d| <= 4)
__device__ __forceinline__ void sphere_gradient(const float *p, int i, float *e)
{
    e[0] = p[i + 1] - p[i - 1];
    e[1] = p[i + PD] - p[i - PD];
    e[2] = p[i + PD * PD] - p[i - PD * PD];
}

// identity patch, normalised (generateFeature3D :1721-1739) + structure tensor over the sphere voxels in raster order
// (determineOrientation3D :2575-2590)
__global__ void __launch_bounds__(256) orient_patch_kernel(const __grid_constant__ PyramidDesc pyr,
                                                           const s3d_keypoint *__restrict__ kps, const int *__restrict__ kp_count,
                                                           float *__restrict__ kp_patch0, float *__restrict__ kp_fmat)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PatchSmem &S = *reinterpret_cast<PatchSmem *>(smem_raw);
    const int nkp = *kp_count;
    const int nsph = c_tab.n_sphere;
    PHASE_INIT();
    for (int kpi = blockIdx.x; kpi < nkp; kpi += gridDim.x) {
        __syncthreads();
        const s3d_keypoint kp = kps[kpi];
        const OctaveDesc &o = pyr.oct[kp.octave];
        if (threadIdx.x == 0) { float id[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 }; invert3(id, S.inv); }
        __syncthreads();
        PHASE(0);
        gather_patch(o.g[kp.level], o.X, o.Y, o.Zg, o.z_off, o.pitch, kp.x, kp.y, kp.z, kp.scale, S.inv, S.patch);
        __syncthreads();
        PHASE(1);
        normalize_patch(S.patch, S.scratch, S.red);
        PHASE(2);
        for (int i = threadIdx.x; i < PV; i += blockDim.x) kp_patch0[(long long)kpi * PV + i] = S.patch[i];
        for (int n = threadIdx.x; n < kSphP; n += blockDim.x) {
            float e[3] = { 0.0f, 0.0f, 0.0f };
            if (n < nsph) sphere_gradient(S.patch, c_tab.sphere[n], e);
            S.ex[n] = e[0]; S.ey[n] = e[1]; S.ez[n] = e[2];
        }
        __syncthreads();
        // 9 lanes, one accumulator each, every sum in sphere order
        if (threadIdx.x < 9) {
            const float *ea = (threadIdx.x / 3 == 0) ? S.ex : (threadIdx.x / 3 == 1) ? S.ey : S.ez;
            const float *eb = (threadIdx.x % 3 == 0) ? S.ex : (threadIdx.x % 3 == 1) ? S.ey : S.ez;
            float m = 0.0f;
            int n = 0;
            for (; n + 4 <= nsph; n += 4) {
                float4 a = *reinterpret_cast<const float4 *>(ea + n), b = *reinterpret_cast<const float4 *>(eb + n);
                m = m + a.x * b.x; m = m + a.y * b.y; m = m + a.z * b.z; m = m + a.w * b.w;
            }
            for (; n < nsph; n++) m = m + ea[n] * eb[n];
            kp_fmat[(size_t)kpi * 9 + threadIdx.x] = m;
        }
        PHASE(3);
    }
}

// one thread per keypoint: SVD of the structure tensor, eigenvalue test (determineOrientation3D :2592-2607); kept
// keypoints are queued for the histogram kernel (any order: its outputs are indexed by keypoint)
__global__ void __launch_bounds__(32) orient_svd_kernel(const float *__restrict__ kp_fmat, const int *__restrict__ kp_count, float eig_thres,
                                                        float *__restrict__ kp_eigs, float *__restrict__ kp_ori0, int *__restrict__ kp_nprim,
                                                        int *__restrict__ work_a, int *work_a_count)
{
    const int nkp = *kp_count;
    PHASE_INIT();
    for (int kpi = blockIdx.x * blockDim.x + threadIdx.x; kpi < nkp; kpi += gridDim.x * blockDim.x) {
        float mat[3][3], v[3][3], w[4];
        double rv1[4];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) mat[a][b] = kp_fmat[(size_t)kpi * 9 + a * 3 + b];
        svd3(mat, w, v, rv1);
        sort_eigen(w, v);
        for (int a = 0; a < 3; a++) { kp_eigs[kpi * 3 + a] = w[a]; for (int b = 0; b < 3; b++) kp_ori0[(size_t)kpi * 9 + a * 3 + b] = v[a][b]; }
        float fEigSum = w[0] + w[1] + w[2];
        float fEigPrd = w[0] * w[1] * w[2];
        float fEigSumProd = fEigSum * fEigSum * fEigSum;
        if (fEigSumProd < eig_thres * fEigPrd || eig_thres < 0) work_a[atomicAdd(work_a_count, 1)] = kpi;
        else kp_nprim[kpi] = -1;
    }
    PHASE(4);
}

// one CTA per kept keypoint: primary direction histogram (determineCanonicalOrientation3D :2779-2885)
__global__ void __launch_bounds__(kHistThreads) orient_hist_kernel(const int *__restrict__ work_a, const int *__restrict__ work_a_count,
                                                                   const float *__restrict__ kp_patch0,
                                                                   int *__restrict__ kp_nprim, float *__restrict__ kp_p1,
                                                                   int *__restrict__ work_b, int *work_b_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HistSmem &S = *reinterpret_cast<HistSmem *>(smem_raw);
    const int n_work = *work_a_count;
    const int nsph = c_tab.n_sphere;
    const float fRadius = 5.0f;
    if (threadIdx.x < 12) S.taps[threadIdx.x] = threadIdx.x < 9 ? c_tab.hist_taps[threadIdx.x] : 0.0f;
    PHASE_INIT();
    for (int it = blockIdx.x; it < n_work; it += gridDim.x) {
        const int kpi = work_a[it];
        __syncthreads();
        for (int i = threadIdx.x; i < PV; i += blockDim.x) S.a.p.patch[i] = kp_patch0[(long long)kpi * PV + i];
        __syncthreads();
        for (int n = threadIdx.x; n < nsph; n += blockDim.x) {
            float e[3];
            sphere_gradient(S.a.p.patch, c_tab.sphere[n], e);
            float fEdgeMagSqr = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
            int base = -1;
            if (fEdgeMagSqr != 0) {
                float fEdgeMag = sqrtf(fEdgeMagSqr);
                float u[3];
                for (int k = 0; k < 3; k++) u[k] = e[k] * fRadius / fEdgeMag;
                for (int k = 0; k < 3; k++) u[k] = u[k] + fRadius;
                make_contrib((float)((double)u[0] + 0.5), (float)((double)u[1] + 0.5), (float)((double)u[2] + 0.5),
                             fEdgeMag, S.b.cw[n], base);
            }
            S.cbase[n] = base;
        }
        __syncthreads();
        PHASE(5);
        splat_histogram_sorted(S, S.b.h0, nsph SPLAT_PASS);
        PHASE(6);
        if (c_tab.n_hist_taps == 3) blur_patch3(S.b.h0, S.a.p.h2, S.taps);
        else blur_patch(S.b.h0, S.a.p.h1, S.a.p.h2, S.taps, c_tab.n_hist_taps);
        PHASE(7);
        find_sort_peaks(S.a.p.h2, S.b.k.pflag, S.b.k.psort, S.b.k.peaks, &S.np);
        PHASE(8);
        // primary directions: peaks >= 0.8 * strongest, at most 11 (:2853-2885)
        if (threadIdx.x == 0) {
            int np = S.np, cnt = 0;
            for (int pi = 0; pi < np && pi < PD; pi++) {
                if ((double)S.b.k.peaks[pi].value < 0.8 * (double)S.b.k.peaks[0].value) break;
                cnt++;
            }
            S.nprim = cnt;
        }
        __syncthreads();
        if ((int)threadIdx.x < S.nprim) {
            float o3[3];
            interp_point_patch(S.a.p.h2, S.b.k.peaks[threadIdx.x].x, S.b.k.peaks[threadIdx.x].y, S.b.k.peaks[threadIdx.x].z, o3);
            o3[0] -= fRadius; o3[1] -= fRadius; o3[2] -= fRadius;
            vec_norm(o3);
            float *dst = kp_p1 + ((long long)kpi * PD + threadIdx.x) * 3;
            dst[0] = o3[0]; dst[1] = o3[1]; dst[2] = o3[2];
        }
        if (threadIdx.x == 0) {
            kp_nprim[kpi] = S.nprim;
            // work items of orient_b: any order (its outputs are indexed by (keypoint, primary))
            int w0 = atomicAdd(work_b_count, S.nprim);
            for (int q = 0; q < S.nprim; q++) work_b[w0 + q] = kpi * PD + q;
        }
        PHASE(13);
    }
}

// one CTA per (keypoint, primary direction): secondary direction histogram (:2887-3033)
__global__ void __launch_bounds__(kHistThreads) orient_b_kernel(const int *__restrict__ work_b, const int *__restrict__ work_b_count,
                                                                const float *__restrict__ kp_p1, const float *__restrict__ kp_patch0,
                                                                int *__restrict__ kp_nsec, float *__restrict__ kp_rots)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HistSmem &S = *reinterpret_cast<HistSmem *>(smem_raw);
    const int n_work = *work_b_count;
    const int nsph = c_tab.n_sphere;
    const float fRadius = 5.0f;
    if (threadIdx.x < 12) S.taps[threadIdx.x] = threadIdx.x < 9 ? c_tab.hist_taps[threadIdx.x] : 0.0f;
    PHASE_INIT();
    for (int it = blockIdx.x; it < n_work; it += gridDim.x) {
        const long long wi = work_b[it];
        const int kpi = (int)(wi / PD), pi = (int)(wi % PD);
        __syncthreads();
        for (int i = threadIdx.x; i < PV; i += blockDim.x) S.a.p.patch[i] = kp_patch0[(long long)kpi * PV + i];
        __syncthreads();
        const float *pp = kp_p1 + ((long long)kpi * PD + pi) * 3;
        const float p1[3] = { pp[0], pp[1], pp[2] };
        for (int n = threadIdx.x; n < nsph; n += blockDim.x) {
            float e[3];
            sphere_gradient(S.a.p.patch, c_tab.sphere[n], e);
            float fEdgeMag = vec_mag(e);
            int base = -1;
            if (fEdgeMag != 0) {
                float u[3] = { e[0], e[1], e[2] };
                vec_norm(u);
                float fPar = vec_dot(p1, u);
                float perp[3];
                perp[0] = u[0] - fPar * p1[0];
                perp[1] = u[1] - fPar * p1[1];
                perp[2] = u[2] - fPar * p1[2];
                vec_norm(perp);
                for (int k = 0; k < 3; k++) { perp[k] = perp[k] * fRadius; perp[k] = perp[k] + fRadius; }
                make_contrib((float)((double)perp[0] + 0.5), (float)((double)perp[1] + 0.5), (float)((double)perp[2] + 0.5),
                             fEdgeMag, S.b.cw[n], base);
            }
            S.cbase[n] = base;
        }
        __syncthreads();
        PHASE(9);
        splat_histogram_sorted(S, S.b.h0, nsph SPLAT_PASS);
        PHASE(10);
        if (c_tab.n_hist_taps == 3) blur_patch3(S.b.h0, S.a.p.h2, S.taps);
        else blur_patch(S.b.h0, S.a.p.h1, S.a.p.h2, S.taps, c_tab.n_hist_taps);
        PHASE(11);
        find_sort_peaks(S.a.p.h2, S.b.k.pflag, S.b.k.psort, S.b.k.peaks, &S.np);
        PHASE(12);
        // secondary peaks >= 0.5 * strongest, at most 11 per primary (the global cap of 11 is applied later)
        const int np2 = S.np;
        if (threadIdx.x < PD) {
            const int j = threadIdx.x;
            bool take = j < np2;
            for (int q = 0; q <= j && take; q++)
                if (S.b.k.peaks[q].value < 0.5f * S.b.k.peaks[0].value) take = false;   // first failure breaks the loop
            if (take) {
                float p2[3], p3[3];
                interp_point_patch(S.a.p.h2, S.b.k.peaks[j].x, S.b.k.peaks[j].y, S.b.k.peaks[j].z, p2);
                p2[0] -= fRadius; p2[1] -= fRadius; p2[2] -= fRadius;
                vec_norm(p2);
                float fPar = vec_dot(p1, p2);
                p2[0] = p2[0] - fPar * p1[0];
                p2[1] = p2[1] - fPar * p1[1];
                p2[2] = p2[2] - fPar * p1[2];
                vec_norm(p2);
                p3[0] = p1[1] * p2[2] - p1[2] * p2[1];
                p3[1] = -p1[0] * p2[2] + p1[2] * p2[0];
                p3[2] = p1[0] * p2[1] - p1[1] * p2[0];
                float *m = kp_rots + ((size_t)wi * PD + j) * 9;
                for (int k = 0; k < 3; k++) { m[k] = p1[k]; m[3 + k] = p2[k]; m[6 + k] = p3[k]; }
            }
            unsigned bal = __ballot_sync(0x7ffu, take);
            if (j == 0) kp_nsec[wi] = __popc(bal);
        }
        PHASE(13);
    }
}

// Rows per keypoint (1 + min(11, sum of secondary counts), or 0 when the eigenvalue test failed),
// their exclusive prefix sum -> first feature row of each keypoint; total -> n_features.
__global__ void __launch_bounds__(1024) row_offsets_kernel(const int *__restrict__ kp_nprim, const int *__restrict__ kp_nsec,
                                                           const int *__restrict__ kp_count, int *__restrict__ nrows,
                                                           int *__restrict__ row_off, int *__restrict__ row_map,
                                                           int *n_features, int row_cap, int *err)
{
    __shared__ int s_warp[32];
    __shared__ int s_run;
    int n = *kp_count;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        int k = start + threadIdx.x;
        int v = 0;
        if (k < n) {
            int np = kp_nprim[k];
            if (np >= 0) {
                int tot = 0;
                for (int i = 0; i < np; i++) tot += kp_nsec[(long long)k * PD + i];
                v = 1 + min(tot, PD);
            }
            nrows[k] = v;
        }
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        int inc = v;
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; w++) { if (w < wid) woff += s_warp[w]; tot += s_warp[w]; }
        if (k < n) {
            int first = s_run + woff + inc - v;
            row_off[k] = first;
            for (int r = 0; r < v; r++)
                if (first + r < row_cap) row_map[first + r] = k * kMaxRowsPerKp + r;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_run > row_cap) atomicOr(err, ERR_ROW_OVERFLOW);
        *n_features = min(s_run, row_cap);
    }
}

// ------------------------------------------------------------------------------------------------
// Descriptor stage: one CTA per feature row (work item = keypoint * 12 + row): re-gather with the row's
// orientation, then the descriptor loop of featExtract's main() (reference featExtract.cpp:477-505):
// NormalizeData, descriptor, rank transform, size factor.
// ------------------------------------------------------------------------------------------------
// A CTA describes kDescGroup rows per round: their patches are gathered one after the other by all threads, then
// NORMALISED TOGETHER -- the two 1331-term sequential sums of Feature3D::NormalizeData must be walked by one thread
// each (parity), so one thread per row walks them side by side instead of one thread for one row while the CTA
// waits (normalize was 32 % of this kernel, profiles/r1_phases_e_current.txt) -- then described one after the other.
constexpr int kDescGroup = 4;
struct DescribeSmem {
    float patch[kDescGroup][PVP];
    float dx[PVP], dy[PVP], dz[PVP];   // SIFT: gradients -> (mag, bin); BRIEF: blur scratch
    float taps[12];
    float wlo[12];
    float inv[9];
    float ori[kDescGroup][9];
    float red[kDescGroup][2];
    float pc[64];
    float pc2[64];
};
__global__ void __launch_bounds__(128) describe_kernel(const __grid_constant__ PyramidDesc pyr,
                                                       const s3d_keypoint *__restrict__ kps, const int *__restrict__ n_features,
                                                       const int *__restrict__ row_map,
                                                       const int *__restrict__ kp_nsec,
                                                       const float *__restrict__ kp_eigs, const float *__restrict__ kp_ori0,
                                                       const float *__restrict__ kp_rots, const float *__restrict__ kp_patch0,
                                                       int descriptor, float size_factor, int octave_base, int row_cap,
                                                       s3d_feature *__restrict__ feats,
                                                       float *__restrict__ dbg_patches, float *__restrict__ dbg_prerank)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DescribeSmem &S = *reinterpret_cast<DescribeSmem *>(smem_raw);
    const int n_rows = min(*n_features, row_cap);
    if (threadIdx.x < 12) {
        S.taps[threadIdx.x] = threadIdx.x < 9 ? c_tab.brief_taps[threadIdx.x] : 0.0f;
        S.wlo[threadIdx.x] = threadIdx.x < PD ? c_tab.desc_w[threadIdx.x] : 0.0f;
    }
    PHASE_INIT();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    // group size: as many rows per CTA as it takes to cover the rows with this grid, at most kDescGroup -- with fewer rows
    // than CTAs every CTA takes one row (shortest latency: one volume alone), large row counts are walked in full groups
    const int G = max(1, min(kDescGroup, (n_rows + (int)gridDim.x - 1) / (int)gridDim.x));
    for (int row0 = blockIdx.x * G; row0 < n_rows; row0 += gridDim.x * G) {
      const int ng = min(G, n_rows - row0);
      // ---- phase A: orientation + patch of every row of the group
      for (int g = 0; g < ng; g++) {
        const int row = row0 + g;
        const int wi = row_map[row];
        const int kpi = wi / kMaxRowsPerKp, r = wi % kMaxRowsPerKp;
        __syncthreads();
        const s3d_keypoint kp = kps[kpi];
        const OctaveDesc &o = pyr.oct[kp.octave];
        // row r > 0 is the (r-1)-th rotation in (primary, secondary) order
        if (threadIdx.x == 0) {
            const float *src;
            if (r == 0) {
                src = kp_ori0 + (size_t)kpi * 9;
            } else {
                int left = r - 1, pi = 0;
                while (left >= kp_nsec[(long long)kpi * PD + pi]) { left -= kp_nsec[(long long)kpi * PD + pi]; pi++; }
                src = kp_rots + (((long long)kpi * PD + pi) * PD + left) * 9;
            }
            float m[9];
            for (int q = 0; q < 9; q++) { m[q] = src[q]; S.ori[g][q] = m[q]; }
            if (r > 0) invert3(m, S.inv);
        }
        PHASE(16);

        if (r == 0) {
            for (int i = threadIdx.x; i < PV; i += blockDim.x) S.patch[g][i] = kp_patch0[(long long)kpi * PV + i];
        } else {
            __syncthreads();
            gather_patch(o.g[kp.level], o.X, o.Y, o.Zg, o.z_off, o.pitch, kp.x, kp.y, kp.z, kp.scale, S.inv, S.patch[g]);
        }
        __syncthreads();
        if (dbg_patches) for (int i = threadIdx.x; i < PV; i += blockDim.x) dbg_patches[(long long)row * PV + i] = S.patch[g][i];
        PHASE(17);
      }
      // ---- phase B: Feature3D::NormalizeData (reference MultiScale.cpp:127-205) of the whole group: lane 0 of warp w
      //      walks the sequential sums of rows w, w + n_warps, ...; same operations in the same order as
      //      normalize_patch (mean, v = p - mean, sum of v*v, p = v * (1 / sqrt(sum)))
      __syncthreads();
      if (lane == 0)
          for (int g = warp; g < ng; g += n_warps) {
              const float *p = S.patch[g];
              const float mean = seq_sum(p, PV) / (float)(PD * PD * PD);
              float e = 0.0f;
              int i = 0;
              for (; i + 4 <= PV; i += 4) {
                  const float4 u = *reinterpret_cast<const float4 *>(p + i);
                  const float v0 = u.x - mean, v1 = u.y - mean, v2 = u.z - mean, v3 = u.w - mean;
                  e = e + v0 * v0; e = e + v1 * v1; e = e + v2 * v2; e = e + v3 * v3;
              }
              for (; i < PV; i++) { const float v = p[i] - mean; e = e + v * v; }
              S.red[g][0] = mean;
              S.red[g][1] = 1.0f / sqrtf(e);
          }
      __syncthreads();
      for (int g = 0; g < ng; g++) {
          const float mean = S.red[g][0], fDiv = S.red[g][1];
          for (int i = threadIdx.x; i < PV; i += blockDim.x) { const float v = S.patch[g][i] - mean; S.patch[g][i] = v * fDiv; }
      }
      __syncthreads();
      PHASE(18);
      // ---- phase C: descriptor of every row of the group
      for (int g = 0; g < ng; g++) {
        const int row = row0 + g;
        const int wi = row_map[row];
        const int kpi = wi / kMaxRowsPerKp, r = wi % kMaxRowsPerKp;
        const s3d_keypoint kp = kps[kpi];
        float *const patch = S.patch[g];
        __syncthreads();

        if (descriptor == S3D_DESC_SIFT) {
            // msResampleFeaturesGradientOrientationHistogram (reference MultiScale.cpp:583-710)
            patch_gradients(patch, S.dx, S.dy, S.dz);
            for (int i = threadIdx.x; i < PV; i += blockDim.x) {
                float e[3] = { S.dx[i], S.dy[i], S.dz[i] };
                float fEdgeMag = vec_mag(e);
                int bin = -1;
                if (fEdgeMag > 0) {
                    vec_norm(e);
                    // dot with (+-1,+-1,+-1) in the reference's bin order, first maximum wins
                    float best = 0.0f;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        float sx = (k & 4) ? -1.0f : 1.0f, sy = (k & 2) ? -1.0f : 1.0f, sz = (k & 1) ? -1.0f : 1.0f;
                        float fDot = sx * e[0] + sy * e[1] + sz * e[2];
                        if (k == 0 || fDot > best) { best = fDot; bin = k; }
                    }
                }
                S.dx[i] = fEdgeMag;
                S.dy[i] = __int_as_float(bin);
            }
            __syncthreads();
            PHASE(19);
            if (threadIdx.x < 32) {
                // thread owns the two bins PC[s*8 + k0], PC[s*8 + k0 + 1] of spatial cell s = (sz*2+sy)*2+sx and walks
                // the 6^3 voxels that can reach the cell in raster order, so each bin still receives its terms in
                // the reference's order; the weight product is formed once per voxel and goes to whichever of the
                // two bins the voxel's orientation selects.  (One warp, two accumulators per lane: 64 threads with
                // one bin each executed 2.2x the instructions -- describe is issue bound, profiles/README.md.)
                const int s = threadIdx.x >> 2, k0 = (threadIdx.x & 3) * 2, k1 = k0 + 1;
                const int sx = s & 1, sy = (s >> 1) & 1, sz = (s >> 2) & 1;
                const int x0 = sx ? 5 : 0, y0 = sy ? 5 : 0, z0 = sz ? 5 : 0;
                float wxs[6];
#pragma unroll
                for (int q = 0; q < 6; q++) wxs[q] = sx ? (1.0f - S.wlo[x0 + q]) : S.wlo[x0 + q];
                float acc0 = 0.0f, acc1 = 0.0f;
                for (int z = z0; z <= z0 + 5; z++) {
                    float wz = sz ? (1.0f - S.wlo[z]) : S.wlo[z];
                    for (int y = y0; y <= y0 + 5; y++) {
                        float wy = sy ? (1.0f - S.wlo[y]) : S.wlo[y];
                        const int i0 = (z * PD + y) * PD + x0;
                        int b[6];
                        float mg[6];
#pragma unroll
                        for (int q = 0; q < 6; q++) { b[q] = __float_as_int(S.dy[i0 + q]); mg[q] = S.dx[i0 + q]; }
#pragma unroll
                        for (int q = 0; q < 6; q++) {
                            const float v = mg[q] * wxs[q] * wy * wz;
                            if (b[q] == k0) acc0 = acc0 + v;
                            else if (b[q] == k1) acc1 = acc1 + v;
                        }
                    }
                }
                S.pc[s * 8 + k0] = acc0;
                S.pc[s * 8 + k1] = acc1;
            }
            __syncthreads();
            PHASE(20);
            // msNormalizeDataPositive (reference MultiScale.cpp:1580-1611)
            if (threadIdx.x == 0) {
                float fMin = 100000;
                for (int i = 0; i < 64; i++) if (S.pc[i] < fMin) fMin = S.pc[i];
                float fSumSqr = 0.0f;
                for (int i = 0; i < 64; i++) { float v = S.pc[i] - fMin; S.pc[i] = v; fSumSqr = fSumSqr + v * v; }
                float fDiv = 1.0f / sqrtf(fSumSqr);
                for (int i = 0; i < 64; i++) S.pc[i] = S.pc[i] * fDiv;
            }
            __syncthreads();
        } else {
            // msResampleFeaturesBRIEF (reference MultiScale.cpp:989-1049), blur with CPU semantics
            blur_patch(patch, S.dy, S.dx, S.taps, c_tab.n_brief_taps);
            if (threadIdx.x < 64) {
                float d = S.dx[c_tab.brief_a[threadIdx.x]] - S.dx[c_tab.brief_b[threadIdx.x]];
                float v;
                if (descriptor == S3D_DESC_BRIEF) v = (d < 0) ? 1.0f : 0.0f;
                else if (descriptor == S3D_DESC_RRIEF) v = d;
                else v = d / c_tab.brief_dist[threadIdx.x];
                S.pc[threadIdx.x] = v;
            }
            __syncthreads();
        }
        if (dbg_prerank && threadIdx.x < 64) dbg_prerank[(long long)row * 64 + threadIdx.x] = S.pc[threadIdx.x];

        // NormalizeDataRankedPCs (reference MultiScale.cpp:207-233, ties by index :3148-3176)
        if (threadIdx.x < 64) {
            float v = S.pc[threadIdx.x];
            int rank = 0;
            for (int j = 0; j < 64; j++) {
                float u = S.pc[j];
                rank += (u < v) || (u == v && j < (int)threadIdx.x);
            }
            S.pc2[threadIdx.x] = (float)rank;
        }
        __syncthreads();

        // geometry: octave rescale (reference MultiScale.cpp:531-543) then featExtract's size factor (:502-505)
        s3d_feature *f = feats + row;
        if (threadIdx.x == 0) {
            float fFactor = 1.0f;
            for (int q = 0; q < kp.octave + octave_base; q++) fFactor = fFactor * 2.0f;
            float sc = kp.scale * fFactor;
            float x = kp.x * fFactor + 0.0f, y = kp.y * fFactor + 0.0f, z = kp.z * fFactor + 0.0f;
            f->flag = (kp.is_max ? 0x10u : 0u) | (r > 0 ? 0x20u : 0u);
            f->x = x * size_factor; f->y = y * size_factor; f->z = z * size_factor; f->scale = sc * size_factor;
            for (int q = 0; q < 3; q++) f->eigs[q] = kp_eigs[kpi * 3 + q];
            for (int q = 0; q < 9; q++) f->ori[q] = S.ori[g][q];
        }
        if (threadIdx.x < 64) f->pc[threadIdx.x] = S.pc2[threadIdx.x];
        PHASE(21);
      }
    }
}

} // namespace s3d
