// s3d_blur_fused.cuh -- one-kernel 3D Gaussian blur level (+ fused DoG) for sm_100a.
//
// Reads the input level once and writes the blurred level (and the DoG) once: 12 B/voxel instead of the
// 32 B/voxel of the three separate passes (s3d_voxel.cuh).  "2.5-D" blocking:
//   * a CTA owns a 32x32 (x,y) tile and walks a z segment plane by plane;
//   * each input plane's tile plus its x/y halo is staged into shared memory by TMA
//     (cp.async.bulk.tensor.3d, mbarrier completion); out-of-volume elements are zero-filled by the
//     TMA unit, which is exactly the reference's zero padding (GaussBlur3D.cpp:329-479), so the kernel
//     has no border branches; tiles are double buffered, the copy of plane p+2 overlaps planes p, p+1;
//   * x pass: shared -> shared (rows incl. the y halo);  y pass: shared -> 4 registers per thread;
//   * z pass: scatter-form march, T = 2R+1 accumulators per column in registers (4 columns/thread);
//     the completed plane is stored with 128-byte coalesced rows, the DoG as in - out.
// Arithmetic is identical to the separate passes and to the reference CPU loop: per tap one FMUL and
// one FADD (no FMA, -fmad=false), taps left to right, axis order x, y, z with fp32 round trips.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "s3d_voxel.cuh"

namespace s3d {

constexpr int kFTX = 32, kFTY = 32, kFThreads = 256;

template <int R>
struct FusedCfg {
    static constexpr int T = 2 * R + 1;
    static constexpr int RP = (R + 3) & ~3;            // x halo rounded so the TMA box is a multiple of 16 bytes
    static constexpr int W = kFTX + 2 * RP;            // staged tile width (floats)
    static constexpr int ROWS = kFTY + 2 * R;          // staged tile height
    static constexpr int IN_FLOATS = (ROWS * W + 31) & ~31;   // each staged tile starts 128-byte aligned (TMA destination)
    static constexpr uint32_t TILE_BYTES = (uint32_t)(ROWS * W * sizeof(float));
    static constexpr int XB_FLOATS = ROWS * kFTX;
    static constexpr size_t SMEM = sizeof(float) * (2 * IN_FLOATS + 2 * XB_FLOATS) + 128 + 64;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// z-march update with static accumulator slot S (see blur_march_kernel): slot S starts a new output,
// the other slots receive their next tap, slot (S+1)%T is complete afterwards.
template <int R, int S>
__device__ __forceinline__ void z_update(float (&acc)[4][2 * R + 1], const float (&v)[4], const TapsSmall &taps, float (&done)[4])
{
    constexpr int T = 2 * R + 1;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        acc[c][S] = taps.w[0] * v[c];
#pragma unroll
        for (int j = 1; j <= 2 * R; j++) acc[c][(S - j + 2 * T) % T] = acc[c][(S - j + 2 * T) % T] + taps.w[j] * v[c];
        done[c] = acc[c][(S + 1) % T];
    }
}

template <int R, int S>
struct ZDispatch {
    static __device__ __forceinline__ void run(int s, float (&acc)[4][2 * R + 1], const float (&v)[4], const TapsSmall &taps, float (&done)[4])
    {
        if (s == S) z_update<R, S>(acc, v, taps, done);
        else ZDispatch<R, S - 1>::run(s, acc, v, taps, done);
    }
};
template <int R>
struct ZDispatch<R, -1> {
    static __device__ __forceinline__ void run(int, float (&)[4][2 * R + 1], const float (&)[4], const TapsSmall &, float (&)[4]) {}
};

template <int R, bool DOG>
__global__ void __launch_bounds__(kFThreads, 2)
blur_fused_kernel(const __grid_constant__ CUtensorMap in_map, const float *__restrict__ in,
                  float *__restrict__ out, float *__restrict__ dog,
                  int X, int Y, int Z, int pitch, int seg_len, const __grid_constant__ TapsSmall taps)
{
    using C = FusedCfg<R>;
    constexpr int T = C::T, RP = C::RP, W = C::W, ROWS = C::ROWS;
    extern __shared__ unsigned char fused_smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(fused_smem_raw) + 127) & ~(uintptr_t)127);
    float *IN = reinterpret_cast<float *>(base);                    // [2][ROWS][W]
    float *XB = IN + 2 * C::IN_FLOATS;                              // [2][ROWS][32]
    uint64_t *full = reinterpret_cast<uint64_t *>(XB + 2 * C::XB_FLOATS);   // [2]

    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kFTX, y0 = blockIdx.y * kFTY;
    const int a0 = blockIdx.z * seg_len, a1 = min(Z, a0 + seg_len);
    const int n_in = (a1 - a0) + 2 * R;
    const long long plane = (long long)pitch * Y;
    constexpr uint32_t kTileBytes = C::TILE_BYTES;

    if (t == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int p = 0; p < 2; p++)
            if (p < n_in) {
                mbar_expect_tx(&full[p], kTileBytes);
                tma_load_3d(IN + p * C::IN_FLOATS, &in_map, x0 - RP, y0 - R, a0 - R + p, &full[p]);
            }
    }

    // this thread's 4 output columns: x = x0 + (t & 31), y = y0 + (t >> 5) + 8k
    const int lx = t & 31, lyb = t >> 5;
    const int gx = x0 + lx;
    float acc[4][T];
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int s = 0; s < T; s++) acc[c][s] = 0.0f;

    int s_mod = 0;
    for (int p = 0; p < n_in; p++) {
        const int buf = p & 1;
        const float *in_t = IN + buf * C::IN_FLOATS;
        float *xb = XB + buf * C::XB_FLOATS;
        const int cz = a0 + p - 2 * R;           // output plane completed by this input plane
        const bool emit = (cz >= a0);            // cz < a1 always holds inside the loop

        // DoG minuend: the input volume at the output position (issued early, used at the end)
        float pv[4] = { 0.f, 0.f, 0.f, 0.f };
        if (DOG && emit && gx < pitch) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int gy = y0 + lyb + 8 * k;
                if (gy < Y) pv[k] = __ldg(in + (long long)cz * plane + (long long)gy * pitch + gx);
            }
        }

        mbar_wait(&full[buf], (p >> 1) & 1);

        // ---- x pass: staged tile -> xb (rows incl. y halo); one item = 2 adjacent outputs
        for (int item = t; item < ROWS * (kFTX / 2); item += kFThreads) {
            const int row = item >> 4, i2 = (item & 15) * 2;
            const float *src = in_t + row * W + i2 + (RP - R);
            float win[2 * R + 2];
            if ((RP - R) % 2 == 0) {
#pragma unroll
                for (int q = 0; q < R + 1; q++) {
                    float2 u = *reinterpret_cast<const float2 *>(src + 2 * q);
                    win[2 * q] = u.x; win[2 * q + 1] = u.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 2 * R + 2; q++) win[q] = src[q];
            }
            float o0 = 0.0f + taps.w[0] * win[0], o1 = 0.0f + taps.w[0] * win[1];      // reference: fSum = 0; fSum += ...
#pragma unroll
            for (int j = 1; j <= 2 * R; j++) { o0 = o0 + taps.w[j] * win[j]; o1 = o1 + taps.w[j] * win[j + 1]; }
            // padding columns (x >= X) must stay zero in every pass
            if (x0 + i2 >= X) o0 = 0.0f;
            if (x0 + i2 + 1 >= X) o1 = 0.0f;
            *reinterpret_cast<float2 *>(xb + row * kFTX + i2) = make_float2(o0, o1);
        }
        __syncthreads();      // xb complete; staged tile `buf` free again
        if (t == 0 && p + 2 < n_in) {
            mbar_expect_tx(&full[buf], kTileBytes);
            tma_load_3d(IN + buf * C::IN_FLOATS, &in_map, x0 - RP, y0 - R, a0 - R + p + 2, &full[buf]);
        }

        // ---- y pass: column lx of xb, rows lyb .. lyb+24+2R, four outputs 8 rows apart
        float v[4];
        {
            float ya[4] = { 0.f, 0.f, 0.f, 0.f };
            const float *colp = xb + lyb * kFTX + lx;
#pragma unroll
            for (int m = 0; m <= 24 + 2 * R; m++) {
                bool used = false;
#pragma unroll
                for (int k = 0; k < 4; k++) used = used || (m - 8 * k >= 0 && m - 8 * k <= 2 * R);
                if (!used) continue;
                float u = colp[m * kFTX];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int j = m - 8 * k;
                    if (j == 0) ya[k] = taps.w[0] * u;
                    else if (j > 0 && j <= 2 * R) ya[k] = ya[k] + taps.w[j] * u;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = ya[k];
        }

        // ---- z pass
        float done[4];
        ZDispatch<R, T - 1>::run(s_mod, acc, v, taps, done);
        s_mod = (s_mod + 1 == T) ? 0 : s_mod + 1;
        if (emit && gx < pitch) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int gy = y0 + lyb + 8 * k;
                if (gy < Y) {
                    long long idx = (long long)cz * plane + (long long)gy * pitch + gx;
                    out[idx] = done[k];
                    if (DOG) dog[idx] = pv[k] - done[k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// x+y pass in one kernel (default path): a CTA stages one 64x64 tile of one z plane plus its halo with a
// single TMA load (zero-filled outside the volume), runs the x pass shared -> shared and the y pass
// shared -> registers -> global.  Together with the z march (blur_march_kernel) a blur level moves
// 24 B/voxel instead of 32, and the two launches replace three.  Same arithmetic as blur_x_kernel /
// blur_march_kernel (taps left to right, FMUL + FADD).
// ---------------------------------------------------------------------------------------------------
constexpr int kXYT = 64;      // tile edge (x and y)

template <int R>
struct XYCfg {
    static constexpr int RP = (R + 3) & ~3;
    static constexpr int W = kXYT + 2 * RP;
    static constexpr int ROWS = kXYT + 2 * R;
    static constexpr int IN_FLOATS = (ROWS * W + 31) & ~31;
    static constexpr uint32_t TILE_BYTES = (uint32_t)(ROWS * W * sizeof(float));
    static constexpr size_t SMEM = sizeof(float) * (IN_FLOATS + ROWS * kXYT) + 128 + 16;
};

template <int R>
__global__ void __launch_bounds__(256) blur_xy_kernel(const __grid_constant__ CUtensorMap in_map, float *__restrict__ out,
                                                      int X, int Y, int pitch, const __grid_constant__ TapsSmall taps)
{
    using C = XYCfg<R>;
    constexpr int RP = C::RP, W = C::W, ROWS = C::ROWS;
    extern __shared__ unsigned char xy_smem_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(xy_smem_raw) + 127) & ~(uintptr_t)127);
    float *IN = reinterpret_cast<float *>(base);          // [ROWS][W]
    float *XB = IN + C::IN_FLOATS;                        // [ROWS][64]
    uint64_t *full = reinterpret_cast<uint64_t *>(XB + ROWS * kXYT);
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kXYT, y0 = blockIdx.y * kXYT, z = blockIdx.z;
    if (t == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(full, C::TILE_BYTES);
        tma_load_3d(IN, &in_map, x0 - RP, y0 - R, z, full);
    }
    __syncthreads();
    mbar_wait(full, 0);

    // ---- x pass: 4 outputs per item from an aligned register window
    for (int item = t; item < ROWS * (kXYT / 4); item += 256) {
        const int row = item >> 4, i4 = (item & 15) * 4;
        const float *src = IN + row * W + i4;
        float win[2 * RP + 4];
#pragma unroll
        for (int q = 0; q < (2 * RP + 4) / 4; q++) {
            float4 u = *reinterpret_cast<const float4 *>(src + 4 * q);
            win[4 * q] = u.x; win[4 * q + 1] = u.y; win[4 * q + 2] = u.z; win[4 * q + 3] = u.w;
        }
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float a = 0.0f + taps.w[0] * win[k + RP - R];      // reference: fSum = 0; fSum += ...
#pragma unroll
            for (int j = 1; j <= 2 * R; j++) a = a + taps.w[j] * win[k + j + RP - R];
            o[k] = (x0 + i4 + k < X) ? a : 0.0f;       // padding columns stay zero
        }
        *reinterpret_cast<float4 *>(XB + row * kXYT + i4) = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();

    // ---- y pass: thread = column lx, 16 consecutive outputs from a register window of 16 + 2R values
    const int lx = t & 63, yb = (t >> 6) * 16;
    const int gx = x0 + lx;
    float col[16 + 2 * R];
#pragma unroll
    for (int m = 0; m < 16 + 2 * R; m++) col[m] = XB[(yb + m) * kXYT + lx];
    if (gx < pitch) {
        float *dst = out + ((long long)z * Y + (y0 + yb)) * pitch + gx;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float a = taps.w[0] * col[k];
#pragma unroll
            for (int j = 1; j <= 2 * R; j++) a = a + taps.w[j] * col[k + j];
            if (y0 + yb + k < Y) dst[(long long)k * pitch] = a;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// tensor map of a pitched fp32 volume with a (W, ROWS, 1) box for radius R; false if unavailable
static bool make_volume_map(CUtensorMap *map, const float *vol, int Y, int Z, int pitch, int R)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    int RP = (R + 3) & ~3;
    cuuint64_t gdim[3] = { (cuuint64_t)pitch, (cuuint64_t)Y, (cuuint64_t)Z };
    cuuint64_t gstr[2] = { (cuuint64_t)pitch * 4, (cuuint64_t)pitch * Y * 4 };
    cuuint32_t box[3] = { (cuuint32_t)(kFTX + 2 * RP), (cuuint32_t)(kFTY + 2 * R), 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// tensor map with the (W, ROWS, 1) box of the x+y kernel
static bool make_volume_map_xy(CUtensorMap *map, const float *vol, int Y, int Z, int pitch, int R)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    int RP = (R + 3) & ~3;
    cuuint64_t gdim[3] = { (cuuint64_t)pitch, (cuuint64_t)Y, (cuuint64_t)Z };
    cuuint64_t gstr[2] = { (cuuint64_t)pitch * 4, (cuuint64_t)pitch * Y * 4 };
    cuuint32_t box[3] = { (cuuint32_t)(kXYT + 2 * RP), (cuuint32_t)(kXYT + 2 * R), 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int R>
static cudaError_t launch_blur_xy(cudaStream_t st, const CUtensorMap &map, float *out, int X, int Y, int Z, int pitch, const float *taps)
{
    using C = XYCfg<R>;
    static_assert(C::SMEM <= 48 * 1024, "x+y tile must fit the default dynamic shared memory limit");
    TapsSmall t;
    memset(&t, 0, sizeof(t));
    for (int j = 0; j < 2 * R + 1; j++) t.w[j] = taps[j];
    dim3 grid((pitch + kXYT - 1) / kXYT, (Y + kXYT - 1) / kXYT, Z);
    blur_xy_kernel<R><<<grid, 256, C::SMEM, st>>>(map, out, X, Y, pitch, t);
    return cudaGetLastError();
}

template <int R>
static cudaError_t launch_blur_fused(cudaStream_t st, const CUtensorMap &map, const float *in, float *out, float *dog,
                                     int X, int Y, int Z, int pitch, const float *taps, int sm_count, int target_ctas)
{
    using C = FusedCfg<R>;
    TapsSmall t;
    memset(&t, 0, sizeof(t));
    for (int j = 0; j < 2 * R + 1; j++) t.w[j] = taps[j];
    int tx = (pitch + kFTX - 1) / kFTX, ty = (Y + kFTY - 1) / kFTY;
    // z segments: enough CTAs to fill the GPU (2 per SM), but each segment re-does 2R planes of halo
    int want = target_ctas > 0 ? target_ctas : 2 * sm_count;
    int n_seg = (want + tx * ty - 1) / (tx * ty);
    int max_seg = (Z + 4 * R - 1) / (4 * R);      // keep the halo overhead <= 50 %
    if (n_seg > max_seg) n_seg = max_seg;
    if (n_seg < 1) n_seg = 1;
    int seg_len = (Z + n_seg - 1) / n_seg;
    n_seg = (Z + seg_len - 1) / seg_len;
    dim3 grid(tx, ty, n_seg);
    static_assert(C::SMEM <= 48 * 1024, "fused blur tile must fit the default dynamic shared memory limit");
    if (dog) blur_fused_kernel<R, true><<<grid, kFThreads, C::SMEM, st>>>(map, in, out, dog, X, Y, Z, pitch, seg_len, t);
    else blur_fused_kernel<R, false><<<grid, kFThreads, C::SMEM, st>>>(map, in, out, nullptr, X, Y, Z, pitch, seg_len, t);
    return cudaGetLastError();
}

} // namespace s3d
