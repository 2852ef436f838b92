// s3d_blur2.cuh -- second-generation blur level for sm_100a (default path): two kernels per level,
//   blur_xy2_kernel : x pass + y pass on a TMA-staged (x,y) tile of one z plane           (in  -> tmp)
//   blur_z2_kernel  : z pass as a register march over column pairs, DoG fused              (tmp -> out, dog)
// Both are bounded by FP32 issue (bit parity forbids FMA: every tap is one FMUL and one FADD, see
// s3d_voxel.cuh), so the design goal is the fewest instructions per voxel:
//   * symmetric taps: w[j] == w[2R-j] bit for bit (Gaussian taps are generated that way, the host checks
//     it), so the product w[j]*v feeds two outputs -- R+1 multiplies per input instead of 2R+1.  The sums
//     still receive their taps left to right, one rounding per product and per sum, exactly like
//     filter_1d of the reference (GaussBlur3D.cpp:43-61);
//   * "static scatter" segments: a work item is K consecutive outputs along the blur axis computed from a
//     register window of K+2R inputs with fully unrolled code -- no loop overhead, no predicates, products
//     shared inside the segment, 128-/64-bit shared memory accesses only;
//   * the z march keeps the 2R+1 partial sums of a column pair in registers (float2, 64-bit global
//     accesses), prefetches inputs a few planes ahead, peels the warm-up round (outputs in front of the
//     segment do not exist, their taps are never issued) and runs interior rounds without predicates.
// Shared-memory rows are padded so that 8 lanes reading 16 bytes from 8 consecutive rows hit 8 different
// bank groups (row pitch / 4 odd).
#pragma once
#include "s3d_voxel.cuh"
#include "s3d_tma.cuh"

namespace s3d {

// Packed fp32x2 primitives (sm_100a FFMA2 / FADD2): two separately rounded products / sums per issue slot.
__device__ __forceinline__ float2 pk_mul(float2 v, float w, float z0) { return __ffma2_rn(v, make_float2(w, w), make_float2(z0, z0)); }
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

__device__ __forceinline__ float mulw(const TapsSmall &tp, int k, float v) { return tp.w[k] * v; }
__device__ __forceinline__ float2 mulw(const TapsSmall &tp, int k, float2 v) { return pk_mul(v, tp.w[k], tp.z0); }
__device__ __forceinline__ float4 mulw(const TapsSmall &tp, int k, float4 v)
{
    const float2 a = pk_mul(make_float2(v.x, v.y), tp.w[k], tp.z0), b = pk_mul(make_float2(v.z, v.w), tp.w[k], tp.z0);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float addv(float a, float b) { return a + b; }
__device__ __forceinline__ float2 addv(float2 a, float2 b) { return pk_add(a, b); }
__device__ __forceinline__ float4 addv(float4 a, float4 b)
{
    const float2 lo = pk_add(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = pk_add(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float2 subv(float2 a, float2 b) { return pk_sub(a, b); }
__device__ __forceinline__ float4 subv(float4 a, float4 b)
{
    const float2 lo = pk_sub(make_float2(a.x, a.y), make_float2(b.x, b.y)), hi = pk_sub(make_float2(a.z, a.w), make_float2(b.z, b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// unpacked variants (one FMUL / FADD per lane) for the z march of wide radii: there the packed forms need aligned
// register pairs for 2R+1 partial sums, 46 more registers per thread at 17 taps, and the march is bound by the
// number of warps that cover its memory latency, not by issue slots
__device__ __forceinline__ float2 mulw_s(const TapsSmall &tp, int k, float2 v) { return make_float2(tp.w[k] * v.x, tp.w[k] * v.y); }
__device__ __forceinline__ float4 mulw_s(const TapsSmall &tp, int k, float4 v) { return make_float4(tp.w[k] * v.x, tp.w[k] * v.y, tp.w[k] * v.z, tp.w[k] * v.w); }
__device__ __forceinline__ float2 addv_s(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float4 addv_s(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float2 subv_s(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float4 subv_s(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
template <typename VT> __device__ __forceinline__ VT zerov();
template <> __device__ __forceinline__ float zerov<float>() { return 0.0f; }
template <> __device__ __forceinline__ float2 zerov<float2>() { return make_float2(0.f, 0.f); }
template <> __device__ __forceinline__ float4 zerov<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }
template <typename VT> __device__ __forceinline__ VT ldgv(const float *p) { return __ldg(reinterpret_cast<const VT *>(p)); }

// K outputs from a window of K+2R inputs (win[c + j] is tap j of output c).  Inputs are visited in
// increasing order, so every output accumulates its taps j = 0..2R in order.
// Every sum starts as 0.0f + (first product), like the reference's `fSum = 0; fSum += ...` (GaussBlur3D.cpp:54-58).
// It only matters for zeros: a window whose products are all -0.0 (masked images: negative value x 0; products of
// denormals that underflow) sums to +0.0 in the reference.  One extra FADD2 per output pair and pass.
template <int R, int K, typename F>
__device__ __forceinline__ void conv_segment(const F *win, F *acc, const TapsSmall &taps)
{
#pragma unroll
    for (int m = 0; m < K + 2 * R; m++) {
        F p[R + 1];
#pragma unroll
        for (int k = 0; k <= R; k++) p[k] = mulw(taps, k, win[m]);     // unused products are dead code
#pragma unroll
        for (int c = 0; c < K; c++) {
            const int j = m - c;
            if (j == 0) acc[c] = addv(zerov<F>(), p[0]);
            else if (j > 0 && j <= 2 * R) acc[c] = addv(acc[c], p[j <= R ? j : 2 * R - j]);
        }
    }
}

// x axis, packed: K outputs (K even) as K/2 pairs from a scalar register window; tap j of output c is
// win[OFF + c + j].  Output pairs start at even c, so a pair position q = c + j serves taps of one parity only:
// about (R+1)/2 products per position, (K+R)(R+1)/2 FFMA2 + K*R FADD2 per segment.  Positions are visited in
// increasing order: every output still accumulates its taps j = 0..2R left to right, starting from 0.0f.
template <int R, int K, int OFF>
__device__ __forceinline__ void seg_x_pk(const float *win, float2 *acc, const TapsSmall &tp)
{
#pragma unroll
    for (int q = 0; q <= K - 2 + 2 * R; q++) {
        const float2 v = make_float2(win[OFF + q], win[OFF + q + 1]);
        float2 p[R + 1];
#pragma unroll
        for (int k = 0; k <= R; k++) p[k] = pk_mul(v, tp.w[k], tp.z0);     // unused products are dead code
#pragma unroll
        for (int c2 = 0; c2 < K / 2; c2++) {
            const int j = q - 2 * c2;
            if (j == 0) acc[c2] = pk_add(make_float2(0.0f, 0.0f), p[0]);
            else if (j > 0 && j <= 2 * R) acc[c2] = pk_add(acc[c2], p[j <= R ? j : 2 * R - j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// x + y passes on one tile.  Tile size (TX x TY) is chosen by the host per volume shape.
// ---------------------------------------------------------------------------------------------------
constexpr int kKX = 16, kXY2Threads = 256;      // y segments are 8 or 16 outputs long (template parameter KY)

struct XY2Tile {
    int TX, TY;          // outputs per tile: TX % 32 == 0, TY % KY == 0
    int KY;              // y-segment length of the kernel instance that runs this tile (host side only)
    int W_in, W_xb;      // shared-memory row pitches (floats), pitch/4 odd
    int rows, rows8;     // TY + 2R, rounded up to 8
    unsigned tile_bytes; // bytes one TMA box brings in
    size_t smem;
};

// Persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (x fastest, then y, then z); the
// staged input is double buffered, so the TMA load of the CTA's next tile is in flight while it computes
// the current one (with one tile per CTA the TMA round trip was exposed: ncu showed the warps parked on
// the mbarrier 3.4 cycles per issued instruction).
template <int R, int KY>
__global__ void __launch_bounds__(kXY2Threads, 2)
blur_xy2_kernel(const __grid_constant__ CUtensorMap in_map, float *__restrict__ out, int X, int Y, int pitch,
                const __grid_constant__ XY2Tile tile, int n_tx, int n_ty, int n_tiles, const __grid_constant__ TapsSmall taps)
{
    constexpr int RP = (R + 3) & ~3;
    extern __shared__ __align__(128) float xy2_smem[];          // TMA destination: 128-byte aligned
    const int stage_floats = tile.rows8 * tile.W_in;
    float *IN = xy2_smem;                                        // [2][rows8][W_in]
    float *XB = IN + 2 * stage_floats;                           // [rows8][W_xb]
    uint64_t *full = reinterpret_cast<uint64_t *>(XB + tile.rows8 * tile.W_xb);   // [2]
    const int t = threadIdx.x;
    const int stride = gridDim.x;
    if (t == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int id = blockIdx.x + s * stride;
            if (id < n_tiles) {
                const int tx = id % n_tx, ty = (id / n_tx) % n_ty, tz = id / (n_tx * n_ty);
                mbar_expect_tx(&full[s], tile.tile_bytes);
                tma_load_3d(IN + s * stage_floats, &in_map, tx * tile.TX - RP, ty * tile.TY - R, tz, &full[s]);
            }
        }
    }
    __syncthreads();

    int k = 0;
    for (int id = blockIdx.x; id < n_tiles; id += stride, k++) {
        const int s = k & 1;
        const float *in_t = IN + s * stage_floats;
        const int x0 = (id % n_tx) * tile.TX, y0 = ((id / n_tx) % n_ty) * tile.TY, z = id / (n_tx * n_ty);
        mbar_wait(&full[s], (k >> 1) & 1);

        // ---- x pass: item = (group of 8 rows, segment of kKX outputs); lane & 7 = row inside the group
        {
            const int n_xseg = tile.TX / kKX;
            const int n_items = (tile.rows8 >> 3) * n_xseg;
            for (int it = t >> 3; it < n_items; it += (int)blockDim.x / 8) {
                const int rg = it / n_xseg, xs = it - rg * n_xseg;
                const int row = rg * 8 + (t & 7);
                const float *src = in_t + row * tile.W_in + xs * kKX;
                float win[kKX + 2 * RP];
#pragma unroll
                for (int q = 0; q < (kKX + 2 * RP) / 4; q++) {
                    float4 u = *reinterpret_cast<const float4 *>(src + 4 * q);
                    win[4 * q] = u.x; win[4 * q + 1] = u.y; win[4 * q + 2] = u.z; win[4 * q + 3] = u.w;
                }
                float2 acc[kKX / 2];
                seg_x_pk<R, kKX, RP - R>(win, acc, taps);
                const int xg = x0 + xs * kKX;
                if (xg + kKX > X) {       // padding columns (x >= X) stay zero in every pass
#pragma unroll
                    for (int q = 0; q < kKX / 2; q++) {
                        if (xg + 2 * q >= X) acc[q].x = 0.0f;
                        if (xg + 2 * q + 1 >= X) acc[q].y = 0.0f;
                    }
                }
                float *dst = XB + row * tile.W_xb + xs * kKX;
#pragma unroll
                for (int q = 0; q < kKX / 4; q++)
                    *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
            }
        }
        fence_proxy_async();  // the x pass read stage s with LDS; TMA refills it behind the barrier
        __syncthreads();      // XB complete, stage s free: refill it with the tile after next
        if (t == 0) {
            const int nid = id + 2 * stride;
            if (nid < n_tiles) {
                const int tx = nid % n_tx, ty = (nid / n_tx) % n_ty, tz = nid / (n_tx * n_ty);
                mbar_expect_tx(&full[s], tile.tile_bytes);
                tma_load_3d(IN + s * stage_floats, &in_map, tx * tile.TX - RP, ty * tile.TY - R, tz, &full[s]);
            }
        }

        // ---- y pass: item = (column pair, segment of KY outputs); consecutive lanes = consecutive pairs
        {
            const int n_pairs = tile.TX >> 1;
            const int n_items = n_pairs * (tile.TY / KY);
            for (int it = t; it < n_items; it += (int)blockDim.x) {
                const int ys = it / n_pairs, cp = it - ys * n_pairs;
                const float *col = XB + (ys * KY) * tile.W_xb + 2 * cp;
                float2 win[KY + 2 * R];
#pragma unroll
                for (int m = 0; m < KY + 2 * R; m++) win[m] = *reinterpret_cast<const float2 *>(col + m * tile.W_xb);
                float2 acc[KY];
                conv_segment<R, KY, float2>(win, acc, taps);
                const int gx = x0 + 2 * cp, gy = y0 + ys * KY;
                if (gx < pitch) {
                    float *dst = out + ((long long)z * Y + gy) * pitch + gx;
                    if (gy + KY <= Y) {
#pragma unroll
                        for (int q = 0; q < KY; q++) *reinterpret_cast<float2 *>(dst + (long long)q * pitch) = acc[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < KY; q++) if (gy + q < Y) *reinterpret_cast<float2 *>(dst + (long long)q * pitch) = acc[q];
                    }
                }
            }
        }
        __syncthreads();      // XB free for the next tile's x pass
    }
}

// ---------------------------------------------------------------------------------------------------
// z pass: register march over column pairs (see s3d_voxel.cuh blur_march_kernel for the scatter form).
// ---------------------------------------------------------------------------------------------------
// S (slot of the output started by this step) is a literal after the caller's loop is unrolled, so
// every accumulator index below is static and acc[] stays in registers.
template <int R, bool FIRST, typename VT, bool PK = true>
__device__ __forceinline__ void z2_step(VT (&acc)[2 * R + 1], const int S, const VT v, const TapsSmall &taps)
{
    constexpr int T = 2 * R + 1;
    VT p[R + 1];
#pragma unroll
    for (int k = 0; k <= R; k++) p[k] = PK ? mulw(taps, k, v) : mulw_s(taps, k, v);
    acc[S] = PK ? addv(zerov<VT>(), p[0]) : addv_s(zerov<VT>(), p[0]);
#pragma unroll
    for (int j = 1; j <= 2 * R; j++)
        if (!FIRST || j <= S) {      // warm-up round: output S - j of the segment does not exist for j > S
            const int sl = (S - j + 2 * T) % T;
            acc[sl] = PK ? addv(acc[sl], p[j <= R ? j : 2 * R - j]) : addv_s(acc[sl], p[j <= R ? j : 2 * R - j]);
        }
}

#ifndef S3D_Z2_PREFETCH_WIDE
#define S3D_Z2_PREFETCH_WIDE 10
#endif
// P: prefetch distance in steps (planes in flight per thread and stream); wide radii hold 148+ registers per thread
// (12 warps per SM), so they need more loads in flight per thread to cover the HBM latency
template <int R> struct Z2Cfg { static constexpr int T = 2 * R + 1; static constexpr int P = (T < 6) ? T - 1 : (R >= 5 ? S3D_Z2_PREFETCH_WIDE : 6); };

template <int R, bool DOG, bool FIRST, bool FAST, typename VT>
__device__ __forceinline__ void z2_round(VT (&acc)[2 * R + 1], VT (&vin)[2 * R + 1], VT (&pv)[2 * R + 1],
                                         const int ub, const int n_in, const int i_base, const int len,
                                         const float *&pin, const float *&ppv, float *&pout, float *&pdog,
                                         const long long plane, const TapsSmall &taps)
{
    constexpr int T = 2 * R + 1;
    constexpr int P = Z2Cfg<R>::P;
#pragma unroll
    for (int S = 0; S < T; S++) {
        const int u = ub + S;
        if (!FAST && u >= n_in) break;          // uniform over the block
        const VT v = vin[S];
        VT pvv = zerov<VT>();
        if (DOG) pvv = pv[S];
        {   // prefetch the input (and the DoG minuend) of step u + P
            const int up = u + P, sl = (S + P) % T;
            if (FAST) {
                vin[sl] = ldgv<VT>(pin);
                if (DOG) pv[sl] = ldgv<VT>(ppv);
            } else {
                const int i = i_base + up;
                vin[sl] = (i >= 0 && i < len && up < n_in) ? ldgv<VT>(pin) : zerov<VT>();
                if (DOG) pv[sl] = (up >= 2 * R && up < n_in) ? ldgv<VT>(ppv) : zerov<VT>();
            }
            pin += plane;
            if (DOG) ppv += plane;
        }
        z2_step<R, FIRST, VT, false>(acc, S, v, taps);
        if (!FIRST || S == 2 * R) {             // step u completes output u - 2R
            if (FAST || u >= 2 * R) {
                const VT g = acc[(S + 1) % T];
                *reinterpret_cast<VT *>(pout) = g;
                if (DOG) *reinterpret_cast<VT *>(pdog) = subv_s(pvv, g);   // prev + (-1)*g, fioMultSum
            }
        }
        pout += plane;
        if (DOG) pdog += plane;
    }
}

// VT = float2 / float4: a thread owns 2 / 4 adjacent columns (64- / 128-bit global accesses)
template <int R, bool DOG, typename VT>
__global__ void __launch_bounds__(128) blur_z2_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                      const float *__restrict__ prev, float *__restrict__ dog,
                                                      int n_vec, long long plane, int len, int seg_len,
                                                      const __grid_constant__ TapsSmall taps)
{
    constexpr int T = 2 * R + 1;
    constexpr int P = Z2Cfg<R>::P;
    constexpr int V = (int)(sizeof(VT) / sizeof(float));
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_vec) return;
    const int a0 = blockIdx.y * seg_len;
    const int a1 = min(len, a0 + seg_len);
    const int n_in = (a1 - a0) + 2 * R;          // input steps u = 0 .. n_in-1, input plane a0 - R + u
    const int i_base = a0 - R;
    VT acc[T], vin[T], pv[T];
#pragma unroll
    for (int s = 0; s < T; s++) { acc[s] = zerov<VT>(); vin[s] = zerov<VT>(); pv[s] = zerov<VT>(); }
    const float *pin = in + V * (long long)q + (long long)i_base * plane;            // input of step 0
    const float *ppv = prev + V * (long long)q + (long long)(a0 - 2 * R) * plane;     // minuend of the output completed at step 0
    float *pout = out + V * (long long)q + (long long)(a0 - 2 * R) * plane;
    float *pdog = dog + V * (long long)q + (long long)(a0 - 2 * R) * plane;
    // prologue: inputs of steps 0 .. P-1
#pragma unroll
    for (int s = 0; s < P; s++) {
        const int i = i_base + s;
        vin[s] = (i >= 0 && i < len && s < n_in) ? ldgv<VT>(pin) : zerov<VT>();
        if (DOG) pv[s] = (s >= 2 * R && s < n_in) ? ldgv<VT>(ppv) : zerov<VT>();
        pin += plane;
        if (DOG) ppv += plane;
    }
    z2_round<R, DOG, true, false, VT>(acc, vin, pv, 0, n_in, i_base, len, pin, ppv, pout, pdog, plane, taps);
    for (int ub = T; ub < n_in; ub += T) {
        // interior round: every prefetch [ub+P, ub+P+T) is inside the volume and the segment, every store is valid
        const bool fast = (i_base + ub + P >= 0) && (i_base + ub + P + T <= len) && (ub + P + T <= n_in) && (ub >= 2 * R);
        if (fast) z2_round<R, DOG, false, true, VT>(acc, vin, pv, ub, n_in, i_base, len, pin, ppv, pout, pdog, plane, taps);
        else z2_round<R, DOG, false, false, VT>(acc, vin, pv, ub, n_in, i_base, len, pin, ppv, pout, pdog, plane, taps);
    }
}

// ---- host side -----------------------------------------------------------------------------------
static inline bool taps_symmetric(const float *taps, int n)
{
    for (int j = 0; j < n / 2; j++)
        if (memcmp(&taps[j], &taps[n - 1 - j], sizeof(float)) != 0) return false;
    return true;
}

// shared-memory pitch >= w with pitch % 4 == 0 and (pitch / 4) odd
static inline int odd_pitch(int w) { int p = (w + 3) & ~3; if (((p >> 2) & 1) == 0) p += 4; return p; }

static inline XY2Tile make_xy2_tile(int TX, int TY, int R, int KY = 16)
{
    XY2Tile t;
    int RP = (R + 3) & ~3;
    t.TX = TX; t.TY = TY; t.KY = KY;
    t.W_in = odd_pitch(TX + 2 * RP);
    t.W_xb = odd_pitch(TX);
    t.rows = TY + 2 * R;
    t.rows8 = (t.rows + 7) & ~7;
    t.tile_bytes = (unsigned)((size_t)t.rows * t.W_in * sizeof(float));
    t.smem = sizeof(float) * (2 * (size_t)t.rows8 * t.W_in + (size_t)t.rows8 * t.W_xb) + 128 + 16;   // 2 input stages + x-pass buffer
    return t;
}

// experiment knobs of the x+y kernel (filled from the environment by the engine, see Tuning in s3d_engine.cu)
struct XY2Tune {
    int max_kb = 75;                 // shared memory per CTA the tile may use
    int force_tx = 0, force_ty = 0;  // forced tile (both or none)
    int ky = 16;                     // y-segment length (8 or 16)
    int threads = kXY2Threads;       // threads per CTA (128 or 256)
};

// Tile choice: minimise (rounds of CTAs per SM) x (work per CTA); work = x-pass rows + y-pass rows, both
// TX wide.  Tiles are limited to 2 resident CTAs per SM (<= ~110 KB of shared memory each).
static inline XY2Tile choose_xy2_tile(int pitch, int Y, int Z, int R, int sm_count, const XY2Tune &tn)
{
    const int max_kb = tn.max_kb, force_tx = tn.force_tx, force_ty = tn.force_ty, force_ky = tn.ky;
    if (force_tx >= 32 && force_tx % 32 == 0 && force_ty >= 8 && force_ty % 8 == 0) {
        int ky = (force_ty % force_ky == 0) ? force_ky : 8;
        XY2Tile t = make_xy2_tile(force_tx, force_ty, R, ky);
        if (t.smem <= 113 * 1024 && t.W_in <= 256 && t.rows <= 256) return t;
    }
    // Cost model in FP32 instructions per thread: the x pass runs ceil(items / 32) rounds of 16-output items on
    // 8-lane groups, the y pass ceil(items / 256) rounds of 2 x KY-output items; a round costs one item whether
    // or not every thread has one (the phases end at a block barrier), so shorter y segments (KY = 8: more,
    // smaller items) win whenever 16-output segments leave most of the threads without an item.  Two persistent
    // CTAs per SM share the tiles.
    XY2Tile best = make_xy2_tile(32, 16, R);
    double best_cost = 1e300;
    const double cx = (double)(kKX + R) * (R + 1) + 2.0 * R * kKX + 24.0;
    for (int KY = 8; KY <= 16; KY += 8) {
        // measured (profiles/README.md): 8-output y segments fill more threads but are never faster than 16-output
        // ones at MNI size (19.9 / 24.2 / 31.1 us against 18.5 / 23.2 / 30.1 us at 7 / 11 / 17 taps), so KY = 8 is
        // only used on request (S3D_XY2_KY=8)
        if (KY != force_ky) continue;
        const double cy = 2.0 * ((double)(KY + R) * (R + 1) + 2.0 * R * KY) + 2.0 * KY + 2.0 * R + 16.0;
        for (int TX = 32; TX <= 128; TX += 32)
            for (int TY = KY; TY <= 128; TY += KY) {
                XY2Tile t = make_xy2_tile(TX, TY, R, KY);
                if (t.smem > (size_t)max_kb * 1024 || t.W_in > 256 || t.rows > 256) continue;
                long long tiles = (long long)((pitch + TX - 1) / TX) * ((Y + TY - 1) / TY) * Z;
                long long rounds = (tiles + 2 * sm_count - 1) / (2 * sm_count);
                int x_items = (t.rows8 / 8) * (TX / kKX), y_items = (TX / 2) * (TY / KY);
                double work = cx * ((x_items + 31) / 32) + cy * ((y_items + kXY2Threads - 1) / kXY2Threads) + 150.0;
                double cost = (double)rounds * work;
                if (cost < best_cost) { best_cost = cost; best = t; }
            }
    }
    return best;
}

template <int R>
static cudaError_t set_xy2_attr_r()
{
    cudaError_t e = cudaFuncSetAttribute(blur_xy2_kernel<R, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(blur_xy2_kernel<R, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
}
// opt-in to > 48 KB of dynamic shared memory for every radius (once per device, from s3d_ctx_create)
static cudaError_t init_blur2_attrs()
{
    cudaError_t e;
    if ((e = set_xy2_attr_r<1>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<2>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<3>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<4>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<5>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<6>()) != cudaSuccess) return e;
    if ((e = set_xy2_attr_r<7>()) != cudaSuccess) return e;
    return set_xy2_attr_r<8>();
}

// x+y: in -> tmp.  Returns false (nothing launched) when the tensor map cannot be encoded.
template <int R>
static bool launch_blur_xy2(cudaStream_t st, const float *in, float *tmp, int X, int Y, int Z, int pitch, const float *taps,
                            int sm_count, int ctas_per_sm, const XY2Tune &tn, cudaError_t *err)
{
    XY2Tile tile = choose_xy2_tile(pitch, Y, Z, R, sm_count, tn);
    CUtensorMap map;
    if (!make_volume_map_box(&map, in, Y, Z, pitch, tile.W_in, tile.rows)) return false;
    TapsSmall t = make_taps_small(taps, 2 * R + 1);
    const int n_tx = (pitch + tile.TX - 1) / tile.TX, n_ty = (Y + tile.TY - 1) / tile.TY;
    const long long n_tiles = (long long)n_tx * n_ty * Z;
    if (n_tiles > 0x7fffffffll) return false;
    // persistent CTAs per SM: 2 when a volume has the GPU to itself; 1 inside a batch, where the other half of
    // every SM is better spent on the memory-bound kernels of the other volumes in flight
    const long long slots = (long long)(ctas_per_sm < 1 ? 1 : ctas_per_sm) * sm_count;
    const int grid = (int)(n_tiles < slots ? n_tiles : slots);
    const int threads = tn.threads;
    if (tile.KY == 8) blur_xy2_kernel<R, 8><<<grid, threads, tile.smem, st>>>(map, tmp, X, Y, pitch, tile, n_tx, n_ty, (int)n_tiles, t);
    else blur_xy2_kernel<R, 16><<<grid, threads, tile.smem, st>>>(map, tmp, X, Y, pitch, tile, n_tx, n_ty, (int)n_tiles, t);
    *err = cudaGetLastError();
    return true;
}

// z (+DoG): tmp -> out.  `target` = threads wanted in flight (segments along z are added to reach it).
// Radii up to kZ2Vec4MaxR march 4 columns per thread (128-bit accesses: twice the bytes in flight per
// load instruction), wider ones 2 columns (register budget: V*(2R+1) partial sums per thread).
constexpr int kZ2Vec4MaxR = 4;

template <int R>
static cudaError_t launch_blur_z2(cudaStream_t st, const float *tmp, float *out, const float *prev, float *dog,
                                  int Y, int Z, int pitch, const float *taps, int target, int force_v)
{
    TapsSmall t = make_taps_small(taps, 2 * R + 1);
    long long plane = (long long)pitch * Y;
    const int V = force_v ? force_v : (R <= kZ2Vec4MaxR ? 4 : 2);
    int n_vec = (int)(plane / V);
    int n_seg = (int)((target + n_vec - 1) / n_vec);
    int max_seg = (Z + 31) / 32;            // segments re-read 2R planes of halo: keep them >= 32 planes
    if (n_seg > max_seg) n_seg = max_seg;
    if (n_seg < 1) n_seg = 1;
    int seg_len = (Z + n_seg - 1) / n_seg;
    n_seg = (Z + seg_len - 1) / seg_len;
    dim3 grid((unsigned)((n_vec + 127) / 128), (unsigned)n_seg);
    if (V == 4) {
        if (dog) blur_z2_kernel<R, true, float4><<<grid, 128, 0, st>>>(tmp, out, prev, dog, n_vec, plane, Z, seg_len, t);
        else blur_z2_kernel<R, false, float4><<<grid, 128, 0, st>>>(tmp, out, tmp, out, n_vec, plane, Z, seg_len, t);
    } else {
        if (dog) blur_z2_kernel<R, true, float2><<<grid, 128, 0, st>>>(tmp, out, prev, dog, n_vec, plane, Z, seg_len, t);
        else blur_z2_kernel<R, false, float2><<<grid, 128, 0, st>>>(tmp, out, tmp, out, n_vec, plane, Z, seg_len, t);
    }
    return cudaGetLastError();
}

} // namespace s3d
