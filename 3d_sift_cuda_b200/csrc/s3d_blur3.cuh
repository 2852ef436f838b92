// s3d_blur3.cuh -- one-kernel blur level (x, y, z passes + DoG) for sm_100a.
//
// Moves the algorithmic minimum through HBM: the input level is read once (plus tile halos), the blurred
// level and the DoG are written once -- 12 B/voxel instead of the 24 B/voxel of the two-kernel path
// (s3d_blur2.cuh), which at MNI size is what bounds the levels with few taps.
//
//   * a CTA (512 threads) owns a 64x32 (x,y) tile and a z segment and walks the segment plane by plane;
//   * every input plane's tile + halo is staged into a ring of shared-memory stages by TMA
//     (cp.async.bulk.tensor.3d + mbarrier), kPF planes ahead of the compute; elements outside the volume
//     (x, y and z) are zero-filled by the TMA unit = the reference's zero padding (GaussBlur3D.cpp:329-479);
//   * x pass: stage -> shared, static scatter segments of 8 outputs, 8 lanes = 8 rows, 128-bit accesses;
//     y pass: shared -> registers, each thread produces 4 consecutive rows of one column;
//     z pass: scatter march in registers -- the thread keeps the 2R+1 partial sums of its 4 columns;
//   * software pipeline with ONE __syncthreads per plane: in iteration u the warps that own x-pass items
//     first run the x pass of plane u+1 (into the other half of a double buffer), then every warp runs
//     the y and z passes of plane u, so the x pass (which only has work for ~60 % of the threads) never
//     leaves issue slots idle;
//   * the DoG minuend (the input level at the output position) is still in the ring R planes later, so
//     the DoG costs no extra global read: ring depth = R + 1 + kPF stages.
// Arithmetic is identical to the other paths and to the reference CPU loop: one FMUL and one FADD per
// tap (no FMA), taps left to right, symmetric products shared (see s3d_blur2.cuh).
#pragma once
#include "s3d_blur2.cuh"

namespace s3d {

constexpr int kF3TX = 64, kF3TY = 32, kF3Threads = 512, kF3KY = 4, kF3KX = 8, kF3PF = 3;

template <int R>
struct F3Cfg {
    static constexpr int T = 2 * R + 1;
    static constexpr int RP = (R + 3) & ~3;
    static constexpr int W0 = kF3TX + 2 * RP;
    static constexpr int W_in = ((W0 >> 2) & 1) ? W0 : W0 + 4;          // pitch / 4 odd
    static constexpr int W_xb = kF3TX + 4;                               // 68: 17 * 4
    static constexpr int ROWS = kF3TY + 2 * R;
    static constexpr int ROWS8 = (ROWS + 7) & ~7;
    static constexpr int NS = R + 1 + kF3PF;
    static constexpr int STAGE = ROWS8 * W_in;                           // floats, multiple of 32 (128 B)
    static constexpr int X_ITEMS = (ROWS8 / 8) * (kF3TX / kF3KX);        // groups of 8 lanes
    static constexpr uint32_t TILE_BYTES = (uint32_t)(ROWS * W_in * sizeof(float));
    static constexpr size_t SMEM = sizeof(float) * ((size_t)NS * STAGE + 2 * ROWS8 * W_xb) + NS * sizeof(uint64_t) + 16;
};

// z-march step on scalar columns (see z2_step in s3d_blur2.cuh); S is a literal after unrolling
template <int R>
__device__ __forceinline__ void z3_step(float (&acc)[2 * R + 1], const int S, const float v, const TapsSmall &taps)
{
    constexpr int T = 2 * R + 1;
    float p[R + 1];
#pragma unroll
    for (int k = 0; k <= R; k++) p[k] = taps.w[k] * v;
    acc[S] = p[0];
#pragma unroll
    for (int j = 1; j <= 2 * R; j++) {
        const int sl = (S - j + 2 * T) % T;
        acc[sl] = acc[sl] + p[j <= R ? j : 2 * R - j];
    }
}

template <int R, int S>
struct F3Dispatch {
    static __device__ __forceinline__ void run(int s, float (&acc)[kF3KY][2 * R + 1], const float (&v)[kF3KY], const TapsSmall &taps,
                                               float (&done)[kF3KY])
    {
        if (s == S) {
#pragma unroll
            for (int c = 0; c < kF3KY; c++) {
                z3_step<R>(acc[c], S, v[c], taps);
                done[c] = acc[c][(S + 1) % (2 * R + 1)];
            }
        } else {
            F3Dispatch<R, S - 1>::run(s, acc, v, taps, done);
        }
    }
};
template <int R>
struct F3Dispatch<R, -1> {
    static __device__ __forceinline__ void run(int, float (&)[kF3KY][2 * R + 1], const float (&)[kF3KY], const TapsSmall &, float (&)[kF3KY]) {}
};

// x pass of one staged plane: stage -> xb.  Only threads t < 8 * X_ITEMS have an item.
template <int R>
__device__ __forceinline__ void f3_x_pass(const float *in_t, float *xb, int t, int x0, int X, const TapsSmall &taps)
{
    using C = F3Cfg<R>;
    constexpr int RP = C::RP;
    const int it = t >> 3;
    if (it < C::X_ITEMS) {
        const int rg = it >> 3, xs = it & 7;            // kF3TX / kF3KX = 8 segments per row
        const int row = rg * 8 + (t & 7);
        const float *src = in_t + row * C::W_in + xs * kF3KX;
        float win[kF3KX + 2 * RP];
#pragma unroll
        for (int q = 0; q < (kF3KX + 2 * RP) / 4; q++) {
            float4 w4 = *reinterpret_cast<const float4 *>(src + 4 * q);
            win[4 * q] = w4.x; win[4 * q + 1] = w4.y; win[4 * q + 2] = w4.z; win[4 * q + 3] = w4.w;
        }
        float o[kF3KX];
        conv_segment<R, kF3KX, float, true>(win + (RP - R), o, taps);
        const int xg = x0 + xs * kF3KX;
        if (xg + kF3KX > X) {       // padding columns (x >= X) stay zero in every pass
#pragma unroll
            for (int k = 0; k < kF3KX; k++) if (xg + k >= X) o[k] = 0.0f;
        }
        float *dst = xb + row * C::W_xb + xs * kF3KX;
#pragma unroll
        for (int q = 0; q < kF3KX / 4; q++)
            *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
}

template <int R, bool DOG>
__global__ void __launch_bounds__(kF3Threads, 1)
blur_f3_kernel(const __grid_constant__ CUtensorMap in_map, float *__restrict__ out, float *__restrict__ dog,
               int X, int Y, int Z, int pitch, int seg_len, const __grid_constant__ TapsSmall taps)
{
    using C = F3Cfg<R>;
    constexpr int T = C::T, RP = C::RP, W_in = C::W_in, W_xb = C::W_xb, NS = C::NS, STAGE = C::STAGE;
    static_assert(8 * C::X_ITEMS <= kF3Threads, "x pass has more items than threads");
    extern __shared__ __align__(128) float f3_smem[];
    float *IN = f3_smem;                                          // [NS][ROWS8][W_in]
    float *XB = IN + NS * STAGE;                                  // [2][ROWS8][W_xb]
    uint64_t *full = reinterpret_cast<uint64_t *>(XB + 2 * C::ROWS8 * W_xb);   // [NS]

    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kF3TX, y0 = blockIdx.y * kF3TY;
    const int a0 = blockIdx.z * seg_len, a1 = min(Z, a0 + seg_len);
    const int n_in = (a1 - a0) + 2 * R;          // input steps u = 0..n_in-1 <-> input plane a0 - R + u

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int u = 0; u <= kF3PF; u++)
            if (u < n_in) {
                mbar_expect_tx(&full[u], C::TILE_BYTES);
                tma_load_3d(IN + u * STAGE, &in_map, x0 - RP, y0 - R, a0 - R + u, &full[u]);
            }
    }

    // this thread's columns: x = x0 + cx, rows y0 + 4*ys .. +3
    const int ys = t >> 6, cx = t & 63;
    const int gx = x0 + cx;
    float acc[kF3KY][T];
#pragma unroll
    for (int c = 0; c < kF3KY; c++)
#pragma unroll
        for (int s = 0; s < T; s++) acc[c][s] = 0.0f;

    // prologue: x pass of plane 0
    mbar_wait(&full[0], 0);
    f3_x_pass<R>(IN, XB, t, x0, X, taps);
    __syncthreads();

    int xstage = 1 % NS, xphase = (1 / NS) & 1;   // ring position / parity of plane u + 1
    int pstage = (kF3PF + 1) % NS;                // ring position of plane u + 1 + kPF (next TMA)
    int mstage = NS - R;                          // ring position of plane u - R (DoG minuend of the output completed at step u)
    int s_mod = 0;
    for (int u = 0; u < n_in; u++) {
        const float *xb = XB + (u & 1) * (C::ROWS8 * W_xb);
        // ---- x pass of the NEXT plane (threads with an item), into the other buffer
        if (u + 1 < n_in && t < 8 * C::X_ITEMS) {
            mbar_wait(&full[xstage], xphase);
            f3_x_pass<R>(IN + xstage * STAGE, XB + ((u + 1) & 1) * (C::ROWS8 * W_xb), t, x0, X, taps);
        }

        // ---- y pass of plane u: 4 consecutive rows of this thread's column
        float v[kF3KY];
        {
            float win[kF3KY + 2 * R];
            const float *col = xb + (ys * kF3KY) * W_xb + cx;
#pragma unroll
            for (int m = 0; m < kF3KY + 2 * R; m++) win[m] = col[m * W_xb];
            conv_segment<R, kF3KY, float>(win, v, taps);
        }

        // ---- z pass
        float done[kF3KY];
        F3Dispatch<R, T - 1>::run(s_mod, acc, v, taps, done);
        if (u >= 2 * R && gx < pitch) {
            const int zc = a0 + u - 2 * R;
            const float *mp = IN + mstage * STAGE + (R + ys * kF3KY) * W_in + RP + cx;
            float *po = out + ((long long)zc * Y + (y0 + ys * kF3KY)) * pitch + gx;
            float *pd = dog + ((long long)zc * Y + (y0 + ys * kF3KY)) * pitch + gx;
#pragma unroll
            for (int k = 0; k < kF3KY; k++) {
                if (y0 + ys * kF3KY + k < Y) {
                    po[(long long)k * pitch] = done[k];
                    if (DOG) pd[(long long)k * pitch] = mp[k * W_in] - done[k];   // prev + (-1)*g, fioMultSum
                }
            }
        }
        __syncthreads();      // XB[(u+1)&1] complete, XB[u&1] and stage of plane u - R free
        if (t == 0 && u + 1 + kF3PF < n_in) {
            mbar_expect_tx(&full[pstage], C::TILE_BYTES);
            tma_load_3d(IN + pstage * STAGE, &in_map, x0 - RP, y0 - R, a0 - R + u + 1 + kF3PF, &full[pstage]);
        }
        s_mod = (s_mod + 1 == T) ? 0 : s_mod + 1;
        if (++xstage == NS) { xstage = 0; xphase ^= 1; }
        if (++pstage == NS) pstage = 0;
        if (++mstage == NS) mstage = 0;
    }
}

// ---- host side -----------------------------------------------------------------------------------
template <int R>
static cudaError_t set_f3_attr_r()
{
    cudaError_t e = cudaFuncSetAttribute(blur_f3_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F3Cfg<R>::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(blur_f3_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F3Cfg<R>::SMEM);
}
static cudaError_t init_blur3_attrs()
{
    cudaError_t e;
    if ((e = set_f3_attr_r<1>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<2>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<3>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<4>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<5>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<6>()) != cudaSuccess) return e;
    if ((e = set_f3_attr_r<7>()) != cudaSuccess) return e;
    return set_f3_attr_r<8>();
}

// One-kernel level: in -> out (+ dog).  Returns false (nothing launched) when the tensor map cannot be encoded.
// z segments: as many as keep every SM busy with one CTA (`want_ctas`, 0 = one per SM), but no shorter
// than max(8, 2R) planes -- each segment re-does the x and y passes of 2R halo planes.
template <int R>
static bool launch_blur_f3(cudaStream_t st, const float *in, float *out, float *dog, int X, int Y, int Z, int pitch,
                           const float *taps, int sm_count, int want_ctas, cudaError_t *err)
{
    using C = F3Cfg<R>;
    CUtensorMap map;
    if (!make_volume_map_box(&map, in, Y, Z, pitch, C::W_in, C::ROWS)) return false;
    TapsSmall t;
    memset(&t, 0, sizeof(t));
    for (int j = 0; j < 2 * R + 1; j++) t.w[j] = taps[j];
    int tx = (pitch + kF3TX - 1) / kF3TX, ty = (Y + kF3TY - 1) / kF3TY;
    int want = want_ctas > 0 ? want_ctas : sm_count;
    int n_seg = want / (tx * ty);
    int min_len = 2 * R > 8 ? 2 * R : 8;
    int max_seg = Z / min_len;
    if (n_seg > max_seg) n_seg = max_seg;
    if (n_seg < 1) n_seg = 1;
    int seg_len = (Z + n_seg - 1) / n_seg;
    n_seg = (Z + seg_len - 1) / seg_len;
    if (n_seg > 65535) return false;
    dim3 grid(tx, ty, n_seg);
    if (dog) blur_f3_kernel<R, true><<<grid, kF3Threads, C::SMEM, st>>>(map, out, dog, X, Y, Z, pitch, seg_len, t);
    else blur_f3_kernel<R, false><<<grid, kF3Threads, C::SMEM, st>>>(map, out, nullptr, X, Y, Z, pitch, seg_len, t);
    *err = cudaGetLastError();
    return true;
}

} // namespace s3d
