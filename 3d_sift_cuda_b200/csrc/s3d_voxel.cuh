// s3d_voxel.cuh -- voxel-parallel stages of the pyramid for sm_100a: separable Gaussian blur
// (x pass + strided "march" pass used for y and z, DoG fused into the z pass), 2x2x2 subsample,
// -2+/-2- resampling, 53-neighbour extrema detection, candidate ordering.
//
// Arithmetic contract (checked bit-for-bit by tests/): every product and every sum is rounded to
// fp32 separately and taps are accumulated left to right, exactly like the reference CPU loop
// filter_1d (reference GaussBlur3D.cpp:43-61).  The translation unit is compiled with -fmad=false
// so the compiler never contracts a*b+c.  Nothing here is a contraction over a long axis, so no
// tensor cores: these are streaming stencils bounded by HBM/L2 bandwidth and FP32 issue rate.
//
// Layout: element (x,y,z) at p[(z*Y+y)*pitch + x]; padding columns X..pitch-1 are kept at zero.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "../../include/s3d.h"

namespace s3d {

constexpr int kMaxFastR = 8;      // widest templated radius (17 taps, sigma 3.09 of the octave schedule)
constexpr int kMaxTaps = 129;

// z0 = -0.0f, handed to the kernels at run time: the packed products are issued as fma.rn.f32x2(v, w, z0), which is
// bit for bit the separately rounded product v*w (adding -0.0 changes nothing, not even the sign of a zero), while
// ptxas cannot see through the addend.  It has to be opaque: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even with --fmad false, which would break bit parity with the reference's unfused arithmetic.
struct TapsSmall { float w[2 * kMaxFastR + 1]; float z0; };
static inline TapsSmall make_taps_small(const float *taps, int n)
{
    TapsSmall t;
    memset(&t, 0, sizeof(t));
    for (int j = 0; j < n && j < 2 * kMaxFastR + 1; j++) t.w[j] = taps[j];
    t.z0 = -0.0f;
    return t;
}
struct TapsAny { int n; float w[kMaxTaps]; };

// Scalar x pass for any pitch / radius (slow path; identical results).
__global__ void blur_x_generic_kernel(const float *__restrict__ in, float *__restrict__ out,
                                      int pitch, int X, long long n_elems, const __grid_constant__ TapsAny taps)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_elems) return;
    int x = (int)(i % pitch);
    int r = taps.n / 2;
    float acc = 0.0f;
    if (x < X) {
        for (int j = 0; j < taps.n; j++) {
            int p = x + j - r;
            float v = (p >= 0 && p < X) ? in[i + j - r] : 0.0f;
            acc = acc + taps.w[j] * v;
        }
    }
    out[i] = acc;
}

// Generic march (any radius): gather form straight from global memory.
template <bool DOG>
__global__ void blur_march_generic_kernel(const float *__restrict__ in, float *__restrict__ out,
                                          const float *__restrict__ prev, float *__restrict__ dog,
                                          long long n_cols, long long w_inner, long long other_stride,
                                          long long stride, int len, const __grid_constant__ TapsAny taps)
{
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_cols) return;
    long long other = q / w_inner;
    long long col = other * other_stride + (q - other * w_inner);
    int r = taps.n / 2;
    for (int c = blockIdx.y; c < len; c += gridDim.y) {
        float acc = 0.0f;
        for (int j = 0; j < taps.n; j++) {
            int p = c + j - r;
            float v = (p >= 0 && p < len) ? in[col + (long long)p * stride] : 0.0f;
            acc = acc + taps.w[j] * v;
        }
        long long idx = col + (long long)c * stride;
        out[idx] = acc;
        if (DOG) dog[idx] = prev[idx] + (-1.0f) * acc;
    }
}

// counters are cleared by a kernel, not a memset node: inside a CUDA graph a memset node costs a hop to
// the copy engine (~60-80 us of idle time measured before the next kernel node)
__global__ void zero_ints_kernel(int *p, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = 0;
}

// out = a + (-1)*b   (fioMultSum, reference FeatureIO.cpp:1950-1987)
__global__ void dog_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + (-1.0f) * b[i];
}

// 2x2x2 mean (fioSubSampleInterpolate, reference FeatureIO.cpp:1474-1554); one thread per output voxel
// (including output padding columns, written as zero).
__global__ void subsample_kernel(const float *__restrict__ in, int X, int Y, int Z, int pitch,
                                 float *__restrict__ out, int ox, int oy, int oz, int opitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int z = blockIdx.z;
    if (x >= opitch || y >= oy) return;
    float r = 0.0f;
    if (x < ox) {
        const float *p0 = in + ((long long)(2 * z) * Y + 2 * y) * pitch + 2 * x;
        const float *p1 = p0 + (long long)Y * pitch;
        float s = 0.0f;
        s = s + (((p0[0] + p0[pitch]) + p0[1]) + p0[pitch + 1]);
        if (2 * z + 1 < Z) {
            s = s + (((p1[0] + p1[pitch]) + p1[1]) + p1[pitch + 1]);
            s = s * 0.125f;
        } else {
            s = s * 0.25f;
        }
        r = s;
    }
    out[((long long)z * oy + y) * opitch + x] = r;
}

// fioSubSample2DCenterPixel (-2-), reference FeatureIO.cpp:1670-1714.
__global__ void halve_kernel(const float *__restrict__ in, int X, int Y, int Z, int pitch,
                             float *__restrict__ out, int ox, int oy, int oz, int opitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int z = blockIdx.z;
    if (x >= opitch || y >= oy) return;
    float r = 0.0f;
    if (x < ox) {
        const float *p0 = in + ((long long)(2 * z) * Y + 2 * y) * pitch + 2 * x;
        const float *p1 = p0 + (long long)Y * pitch;
        float v = 0.0f;
        v = v + p0[0];
        v = v + p1[0];
        v = v + p0[pitch];
        v = v + p1[pitch];
        v = v + p0[1];
        v = v + p1[1];
        v = v + p0[pitch + 1];
        v = v + p1[pitch + 1];
        r = v / 8.0f;
    }
    out[((long long)z * oy + y) * opitch + x] = r;
}

// fioDoubleSize (-2+), reference FeatureIO.cpp:2452-2548; one thread per OUTPUT voxel.  The
// reference scatters 2x2x2 blocks per input voxel; every output voxel is written exactly once
// (dims are doubled exactly), by the input voxel (x/2,y/2,z/2) with sub-position (x&1,y&1,z&1).
__global__ void double_kernel(const float *__restrict__ in, int X, int Y, int Z, int pitch,
                              float *__restrict__ out, int opitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    int z = blockIdx.z;
    int DX = 2 * X, DY = 2 * Y;
    if (x >= opitch || y >= DY) return;
    float r = 0.0f;
    if (x < DX) {
        int sx = x >> 1, sy = y >> 1, sz = z >> 1;
        int bx = x & 1, by = y & 1, bz = z & 1;
        float lo[2][2][2];
#pragma unroll
        for (int zz = 0; zz < 2; zz++)
#pragma unroll
            for (int yy = 0; yy < 2; yy++)
#pragma unroll
                for (int xx = 0; xx < 2; xx++) {
                    int dz = (sz + zz >= Z) ? 0 : zz, dy = (sy + yy >= Y) ? 0 : yy, dx = (sx + xx >= X) ? 0 : xx;
                    lo[zz][yy][xx] = in[((long long)(sz + dz) * Y + (sy + dy)) * pitch + (sx + dx)];
                }
        int code = bz * 4 + by * 2 + bx;
        switch (code) {
        case 0: r = lo[0][0][0]; break;
        case 4: r = 0.5f * (lo[0][0][0] + lo[1][0][0]); break;
        case 2: r = 0.5f * (lo[0][0][0] + lo[0][1][0]); break;
        case 1: r = 0.5f * (lo[0][0][0] + lo[0][0][1]); break;
        case 6: r = 0.25f * (((lo[0][0][0] + lo[1][0][0]) + lo[0][1][0]) + lo[1][1][0]); break;
        case 3: r = 0.25f * (((lo[0][0][0] + lo[0][1][0]) + lo[0][0][1]) + lo[0][1][1]); break;
        case 5: r = 0.25f * (((lo[0][0][0] + lo[1][0][0]) + lo[0][0][1]) + lo[1][0][1]); break;
        default:
            r = 0.125f * (((((((lo[0][0][0] + lo[0][0][1]) + lo[0][1][0]) + lo[0][1][1]) + lo[1][0][0]) + lo[1][0][1]) + lo[1][1][0]) + lo[1][1][1]);
            break;
        }
    }
    out[((long long)z * DY + y) * opitch + x] = r;
}

// dense (pitch == X) <-> pitched copies with zeroed padding.  Block = (64, 4): four rows per block, a thread
// moves elements x = tx, tx + 64, ... of its row (coalesced 32-bit loads and stores, 32-bit index arithmetic
// inside the row); rows are walked with a grid-stride loop.
__global__ void pad_rows_kernel(const float *__restrict__ in, int X, long long rows, float *__restrict__ out, int pitch)
{
    for (long long row = blockIdx.y * blockDim.y + threadIdx.y; row < rows; row += (long long)gridDim.y * blockDim.y) {
        const float *src = in + row * X;
        float *dst = out + row * pitch;
        for (int x0 = blockIdx.x * 256 + threadIdx.x; x0 < pitch; x0 += gridDim.x * 256) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const int x = x0 + 64 * j; v[j] = (x < X) ? __ldg(src + x) : 0.0f; }
#pragma unroll
            for (int j = 0; j < 4; j++) { const int x = x0 + 64 * j; if (x < pitch) dst[x] = v[j]; }
        }
    }
}

// Typed input (NIfTI scalar datatypes): the reference converts every voxel to float with a plain C cast on the
// host before anything else (reg_changeDatatype1, R/featExtract/featExtract.cpp:18-77).  Here the raw voxels
// cross PCIe in their file datatype and the cast is fused into the re-pitch; int -> float and double -> float
// casts round to nearest even on both sides, so the result is the same bits.
template <typename T>
__global__ void convert_rows_kernel(const T *__restrict__ in, int X, long long rows, float *__restrict__ out, int pitch)
{
    for (long long row = blockIdx.y * blockDim.y + threadIdx.y; row < rows; row += (long long)gridDim.y * blockDim.y) {
        const T *src = in + row * X;
        float *dst = out + row * pitch;
        for (int x0 = blockIdx.x * 256 + threadIdx.x; x0 < pitch; x0 += gridDim.x * 256) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { const int x = x0 + 64 * j; v[j] = (x < X) ? (float)src[x] : 0.0f; }
#pragma unroll
            for (int j = 0; j < 4; j++) { const int x = x0 + 64 * j; if (x < pitch) dst[x] = v[j]; }
        }
    }
}

__global__ void unpad_rows_kernel(const float *__restrict__ in, int pitch, long long rows, float *__restrict__ out, int X)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * X) return;
    long long row = i / X;
    int x = (int)(i - row * X);
    out[i] = in[row * pitch + x];
}

// ---------------------------------------------------------------------------------------------
// Detection (reference MultiScale.cpp:2260-2524): a voxel of the centre DoG is a candidate when it
// is strictly above (below) its 26 neighbours and all 27 voxels of the finer DoG.  One thread per
// interior voxel; the x neighbours reject most voxels after two loads, the rest exit as soon as a
// comparison fails, so the pass costs little more than reading the centre volume once.
// Survivors are appended through an atomic counter; order_candidates_kernel restores raster order.
// ---------------------------------------------------------------------------------------------
struct CandList {
    s3d_cand *items;
    int *count;
};

constexpr int kDetectZ = 8;   // centre voxels per thread along z

// Pass 1: branch-free 6-face test of the centre DoG for every interior voxel.  All loads of a thread
// (its z column and the x/y neighbours of its 8 voxels) are issued up front, so the pass runs at memory
// speed; the few survivors (strict extrema along x, y and z) are appended to a list of linear voxel
// offsets with one aggregated atomic per warp.
__global__ void __launch_bounds__(256) detect_face_kernel(const float *__restrict__ finer, const float *__restrict__ centre,
                                                          int X, int Y, int Z, int pitch, int n_zblocks,
                                                          unsigned int *__restrict__ face, int *face_count, int face_cap)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int y = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const bool inside = (x <= X - 2 && y <= Y - 2);
    const long long plane = (long long)pitch * Y;
    const unsigned int base = (unsigned int)((inside ? y : 1) * pitch + (inside ? x : 1));
    const int lane = (threadIdx.y * blockDim.x + threadIdx.x) & 31;
    // gridDim.z may be smaller than the number of z blocks (contexts of a batch cap this kernel's footprint so that
    // other volumes' kernels stay resident next to it): a block then walks several z blocks
    for (int zb = blockIdx.z; zb < n_zblocks; zb += gridDim.z) {
        const int z0 = zb * kDetectZ + 1;
        // strict extremum over the 6 face neighbours  <=>  c > max(neighbours)  or  c < min(neighbours):
        // two 3-input min/max trees per voxel instead of twelve compares (the pass is issue bound, not
        // memory bound: ncu showed 58 instructions per voxel at 67 % issue utilisation before this form)
        float up, c, dn;                 // z-1, z, z+1 of the running column
        const float *p = centre + (long long)(z0 - 1) * plane + base;
        up = __ldg(p);
        p += plane;
        c = __ldg(p);
        unsigned livemask = 0;
        if (z0 + kDetectZ <= Z - 1) {    // every plane z0-1 .. z0+kDetectZ exists: no clamps (uniform per block)
#pragma unroll
            for (int k = 0; k < kDetectZ; k++) {
                const float xm = __ldg(p - 1), xp = __ldg(p + 1), ym = __ldg(p - pitch), yp = __ldg(p + pitch);
                dn = __ldg(p + plane);
                const float hi = fmaxf(fmaxf(fmaxf(xm, xp), fmaxf(ym, yp)), fmaxf(up, dn));
                const float lo = fminf(fminf(fminf(xm, xp), fminf(ym, yp)), fminf(up, dn));
                if (c > hi || c < lo) livemask |= 1u << k;
                up = c; c = dn; p += plane;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kDetectZ; k++) {
                const bool zin = (z0 + k <= Z - 2);
                const float *q = zin ? p : centre + base;       // keep the addresses legal past the last interior plane
                const float xm = __ldg(q - 1), xp = __ldg(q + 1), ym = __ldg(q - pitch), yp = __ldg(q + pitch);
                dn = __ldg(zin ? q + plane : q);
                const float hi = fmaxf(fmaxf(fmaxf(xm, xp), fmaxf(ym, yp)), fmaxf(up, dn));
                const float lo = fminf(fminf(fminf(xm, xp), fminf(ym, yp)), fminf(up, dn));
                if (zin && (c > hi || c < lo)) livemask |= 1u << k;
                up = c; c = dn; p += plane;
            }
        }
        if (!inside) livemask = 0;
        // one atomic per warp for all 8 z steps
        int mine = __popc(livemask), incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        int start = 0;
        if (lane == 31) start = atomicAdd(face_count, total);
        start = __shfl_sync(0xffffffffu, start, 31);
        int pos = start + incl - mine;
#pragma unroll
        for (int k = 0; k < kDetectZ; k++)
            if (livemask & (1u << k)) {
                if (pos < face_cap) face[pos] = (unsigned int)((long long)(z0 + k) * plane + base);
                pos++;
            }
    }
}

// Pass 2: the full 26 + 27 neighbour test (reference MultiScale.cpp:2260-2524) on the survivors only.
__global__ void __launch_bounds__(256) detect_full_kernel(const float *__restrict__ finer, const float *__restrict__ centre,
                                                          int X, int Y, int Z, int pitch,
                                                          const unsigned int *__restrict__ face, const int *__restrict__ face_count,
                                                          int face_cap, CandList mins, CandList maxs, int cap, int *err, int err_bit,
                                                          int own0, int own1)
{
    int n = *face_count;
    if (n > face_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(err, err_bit); n = face_cap; }
    const long long plane = (long long)pitch * Y;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const long long i = face[k];
        const float c = centre[i];
        bool mx = true, mn = true;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const float *row = centre + i + dz * plane + dy * pitch;
                float a = row[-1], b = row[0], d = row[1];
                if (dz == 0 && dy == 0) b = a; // skip self
                mx = mx && (a < c) && (b < c) && (d < c);
                mn = mn && (a > c) && (b > c) && (d > c);
                row = finer + i + dz * plane + dy * pitch;
                a = row[-1]; b = row[0]; d = row[1];
                mx = mx && (a < c) && (b < c) && (d < c);
                mn = mn && (a > c) && (b > c) && (d > c);
            }
        if (mx || mn) {
            int z = (int)(i / plane);
            if (z < own0 || z >= own1) continue;       // slab mode: halo planes belong to the neighbour
            int rem = (int)(i - (long long)z * plane);
            int y = rem / pitch, x = rem - y * pitch;
            if (mx) { int kk = atomicAdd(maxs.count, 1); if (kk < cap) maxs.items[kk] = s3d_cand{ x, y, z, c }; }
            if (mn) { int kk = atomicAdd(mins.count, 1); if (kk < cap) mins.items[kk] = s3d_cand{ x, y, z, c }; }
        }
    }
}

// single-kernel variant (kept for volumes too large for 32-bit voxel offsets)
// one block of the single-kernel detection: block (bx, by, bz) of a (32, 8) x kDetectZ tiling of the interior voxels
__device__ __forceinline__ void detect_block(const float *__restrict__ finer, const float *__restrict__ centre,
                                             int X, int Y, int Z, int pitch,
                                             CandList mins, CandList maxs, int cap, int own0, int own1, int bx, int by, int bz)
{
    int x = bx * blockDim.x + threadIdx.x + 1;
    int y = by * blockDim.y + threadIdx.y + 1;
    int z0 = bz * kDetectZ + 1;
    if (x > X - 2 || y > Y - 2 || z0 > Z - 2) return;
    long long plane = (long long)pitch * Y;
    long long base = (long long)y * pitch + x;
    float col[kDetectZ + 2];
#pragma unroll
    for (int k = 0; k < kDetectZ + 2; k++) {
        int z = z0 - 1 + k;
        col[k] = (z <= Z - 1) ? __ldg(centre + (long long)z * plane + base) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < kDetectZ; k++) {
        int z = z0 + k;
        if (z > Z - 2) break;
        if (z < own0 || z >= own1) continue;
        float c = col[k + 1];
        bool mx = (col[k] < c) && (col[k + 2] < c);
        bool mn = (col[k] > c) && (col[k + 2] > c);
        if (!(mx || mn)) continue;
        long long i = (long long)z * plane + base;
        {
            float a = centre[i - 1], b = centre[i + 1];
            mx = mx && (a < c) && (b < c);
            mn = mn && (a > c) && (b > c);
            if (!(mx || mn)) continue;
            a = centre[i - pitch]; b = centre[i + pitch];
            mx = mx && (a < c) && (b < c);
            mn = mn && (a > c) && (b > c);
            if (!(mx || mn)) continue;
            a = finer[i];
            mx = mx && (a < c);
            mn = mn && (a > c);
            if (!(mx || mn)) continue;
        }
#pragma unroll 1
        for (int dz = -1; dz <= 1 && (mx || mn); dz++)
#pragma unroll 1
            for (int dy = -1; dy <= 1 && (mx || mn); dy++) {
                const float *row = centre + i + dz * plane + dy * pitch;
                float a = row[-1], b = row[0], d = row[1];
                if (dz == 0 && dy == 0) b = a; // skip self
                mx = mx && (a < c) && (b < c) && (d < c);
                mn = mn && (a > c) && (b > c) && (d > c);
                row = finer + i + dz * plane + dy * pitch;
                a = row[-1]; b = row[0]; d = row[1];
                mx = mx && (a < c) && (b < c) && (d < c);
                mn = mn && (a > c) && (b > c) && (d > c);
            }
        if (mx) {
            int kk = atomicAdd(maxs.count, 1);
            if (kk < cap) maxs.items[kk] = s3d_cand{ x, y, z, c };
        }
        if (mn) {
            int kk = atomicAdd(mins.count, 1);
            if (kk < cap) mins.items[kk] = s3d_cand{ x, y, z, c };
        }
    }
}

__global__ void __launch_bounds__(256) detect_kernel(const float *__restrict__ finer, const float *__restrict__ centre,
                                                     int X, int Y, int Z, int pitch,
                                                     CandList mins, CandList maxs, int cap, int own0, int own1)
{
    detect_block(finer, centre, X, Y, Z, pitch, mins, maxs, cap, own0, own1, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Several detections in one launch: the three centre levels of a small octave (after its last level), or every
// (octave, level) pair of the tiny tail of the pyramid.  Block b belongs to the job whose block range holds it.
constexpr int kMaxDetectJobs = 12;
struct DetectJob {
    const float *finer, *centre;
    int X, Y, Z, pitch;
    CandList mins, maxs;
    int own0, own1;
    int nbx, nby, first_block;     // block grid (nbx, nby, nbz) of the job starts at linear block `first_block`
};
struct DetectJobs { int n; int cap; DetectJob job[kMaxDetectJobs]; };

__global__ void __launch_bounds__(256) detect_multi_kernel(const __grid_constant__ DetectJobs J)
{
    int k = 0;
    while (k + 1 < J.n && (int)blockIdx.x >= J.job[k + 1].first_block) k++;
    const DetectJob &q = J.job[k];
    const int b = blockIdx.x - q.first_block;
    const int bx = b % q.nbx, by = (b / q.nbx) % q.nby, bz = b / (q.nbx * q.nby);
    detect_block(q.finer, q.centre, q.X, q.Y, q.Z, q.pitch, q.mins, q.maxs, J.cap, q.own0, q.own1, bx, by, bz);
}

// Rank-by-counting sort of a candidate list into raster order (keys are unique voxel indices).
// When the list overflowed (count > cap) the retained subset depends on atomic order; the
// engine reports S3D_ERR_CAPACITY in that case.
__global__ void order_candidates_kernel(const s3d_cand *__restrict__ in, const int *__restrict__ count,
                                        s3d_cand *__restrict__ out, int X, int Y, int cap)
{
    int n = min(*count, cap);
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    s3d_cand me = in[k];
    long long key = ((long long)me.z * Y + me.y) * X + me.x;
    int rank = 0;
    for (int j = 0; j < n; j++) {
        s3d_cand o = in[j];
        long long kj = ((long long)o.z * Y + o.y) * X + o.x;
        rank += (kj < key);
    }
    out[rank] = me;
}

} // namespace s3d
