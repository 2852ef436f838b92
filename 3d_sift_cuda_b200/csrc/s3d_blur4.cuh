// s3d_blur4.cuh -- one-kernel blur level (x, y, z passes + DoG) for sm_100a: warp-specialised, packed fp32x2.
//
// Why one kernel: a level moves 12 B/voxel through HBM (read G_{j-1}, write G_j and the DoG).  The two-kernel
// level (s3d_blur2.cuh) moves 24 B/voxel and its z pass alone already runs at copy speed for those bytes, so
// the levels with few taps can only get faster by not making the round trip.
// Why packed: bit parity with the reference forbids FMA, every tap is a separately rounded product and sum
// (GaussBlur3D.cpp:43-61).  FFMA2/FADD2 process two lanes of fp32 per issue slot at the same lane throughput
// as FMUL/FADD (tools/mb/f32x2_bench.cu: 124 lane-ops/clk/SM either way), which halves the issue slots of the
// arithmetic and leaves room for the shared-memory traffic of a tiled kernel.
//
//   * a CTA owns a 64 x 16 (or 64 x 32) (x,y) tile and a z segment, and walks the segment plane by plane;
//   * TMA (cp.async.bulk.tensor.3d + mbarrier) stages every input plane's tile + halo into a ring, kF4Ahead planes
//     ahead; elements outside the volume are zero-filled = the reference's zero padding (GaussBlur3D.cpp:329-479);
//   * producer warps (x pass): ring stage -> shared buffer XB (a hand-over ring of kF4NXB slots), segments of KX outputs on packed
//     pairs (even/odd pair alignment costs one MOV per input), 8 lanes = 8 rows, 128-bit shared accesses, row
//     pitch / 4 odd -> conflict free;
//   * consumer warps (y and z passes): a thread owns VX adjacent columns (float2 / float4 = packed operands) and
//     4 rows.  y pass: XB -> registers.  z pass: scatter march in registers, 2R+1 packed partial sums per output,
//     rotating slot picked by a switch on (plane mod 2R+1) so every accumulator index is static.  The DoG minuend
//     (input level at the output position) is still in the ring R planes later (radii <= 3), or read back from L2;
//   * no CTA-wide barrier in the plane loop: producers and consumers hand XB over through two mbarrier pairs
//     (full / empty), and a ring stage is refilled as soon as the consumers' `empty` arrival proves that its
//     plane has been read as a minuend (one ncu capture of the barrier-per-plane predecessor: 40 % issue
//     utilisation, 0.7 barrier + 1.4 fixed-latency stall cycles per issue with 2 warps per scheduler).
#pragma once
#include "s3d_blur2.cuh"

namespace s3d {

// kF4NXB: depth of the x-pass -> y-pass hand-over ring.  The level runs the same with 2, 3 or 4 slots (257.9 / 259.0 / 259.0 us for the
// six levels of octave 0), but every slot is 6.5-8.7 KB of shared memory per CTA and a batch pays for the footprint of
// the level CTAs (other volumes' kernels share their SMs): 464 us per volume with 4 slots, 452 with 3, 449 with 2
constexpr int kF4TX = 64, kF4KY = 4, kF4Ahead = 3, kF4NXB = 3;

// R: radius, VX: columns per consumer thread (2 | 4), KX: outputs per x-pass segment (8 | 16), TY: tile rows, XW: producer warps,
// MRING: the DoG minuend is read from the ring (R + 1 more stages) instead of from global memory (L2)
template <int R, int VX, int KX, int TY, int XW, bool MRING>
struct F4Cfg {
    static constexpr int T = 2 * R + 1;
    static constexpr int YZ_WARPS = (kF4TX / VX) * (TY / kF4KY) / 32;
    static constexpr int THREADS = 32 * (YZ_WARPS + XW);
    static constexpr int RP = (R + 3) & ~3;
    static constexpr int W0 = kF4TX + 2 * RP;
    static constexpr int W_in = ((W0 >> 2) & 1) ? W0 : W0 + 4;          // pitch / 4 odd
    static constexpr int W_xb = kF4TX + 4;                               // 68 = 17 * 4
    static constexpr int ROWS = TY + 2 * R;
    static constexpr int ROWS8 = (ROWS + 7) & ~7;
    static constexpr int NS = (MRING ? R : 0) + kF4NXB + kF4Ahead;      // planes q-(NXB-1)(-R) .. q live, kF4Ahead in flight
    static constexpr int STAGE = ROWS8 * W_in;                           // floats, multiple of 32 (128 B)
    static constexpr int X_WARP_ITEMS = (ROWS8 / 8) * (kF4TX / (4 * KX));   // a warp item = 8 rows x 4 segments of KX outputs
    static constexpr uint32_t TILE_BYTES = (uint32_t)(ROWS * W_in * sizeof(float));
    static constexpr size_t SMEM = sizeof(float) * ((size_t)NS * STAGE + kF4NXB * ROWS8 * W_xb) + (NS + 2 * kF4NXB) * sizeof(uint64_t) + 16;
    // two full warpgroups (4 consumer + 4 producer warps) at two CTAs per SM: 128 registers per thread on average,
    // moved from the producers (the x pass needs ~60) to the consumers (2R+1 partial sums x 8 outputs) with setmaxnreg
    static constexpr bool REBALANCE = (R >= 5) && YZ_WARPS == 4 && XW == 4 && TY == 16;
    static constexpr int PRODUCER_REGS = 64, CONSUMER_REGS = 192;
    static constexpr int MIN_CTAS = (2 * (SMEM + 1024) <= 227 * 1024 && 2 * THREADS <= 512) ? 2 : 1;
    static_assert(SMEM + 1024 <= 227 * 1024, "one-kernel level: tile does not fit shared memory");   // resident CTAs per SM the register budget is sized for
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int VX> struct F4Vec;
template <> struct F4Vec<2> { typedef float2 type; };
template <> struct F4Vec<4> { typedef float4 type; };

// z step on the accumulators of KY outputs with static slot S; `done` = the outputs completed by this step
template <int R, int VX, int LO, int HI>
struct F4Dispatch {
    typedef typename F4Vec<VX>::type VT;
    static __device__ __forceinline__ void run(int s, VT (&acc)[kF4KY][2 * R + 1], const VT (&v)[kF4KY], const TapsSmall &taps, VT (&done)[kF4KY])
    {
        if constexpr (LO == HI) {
#pragma unroll
            for (int c = 0; c < kF4KY; c++) {
                z2_step<R, false, VT>(acc[c], LO, v[c], taps);
                done[c] = acc[c][(LO + 1) % (2 * R + 1)];
            }
        } else {
            constexpr int MID = (LO + HI) / 2;
            if (s <= MID) F4Dispatch<R, VX, LO, MID>::run(s, acc, v, taps, done);
            else F4Dispatch<R, VX, MID + 1, HI>::run(s, acc, v, taps, done);
        }
    }
};

template <int R, int VX, int KX, int TY, int XW, bool MRING, bool DOG>
__global__ void __launch_bounds__(F4Cfg<R, VX, KX, TY, XW, MRING>::THREADS, F4Cfg<R, VX, KX, TY, XW, MRING>::MIN_CTAS)
blur_f4_kernel(const __grid_constant__ CUtensorMap in_map, const float *__restrict__ in, float *__restrict__ out, float *__restrict__ dog,
               int X, int Y, int Z, int pitch, int seg_len, const __grid_constant__ TapsSmall taps)
{
    using C = F4Cfg<R, VX, KX, TY, XW, MRING>;
    typedef typename F4Vec<VX>::type VT;
    constexpr int T = C::T, RP = C::RP, W_in = C::W_in, W_xb = C::W_xb, NS = C::NS, STAGE = C::STAGE, KY = kF4KY;
    extern __shared__ __align__(128) float f4_smem[];
    float *IN = f4_smem;                                          // [NS][ROWS8][W_in]
    float *XB = IN + NS * STAGE;                                  // [NXB][ROWS8][W_xb]
    uint64_t *full_in = reinterpret_cast<uint64_t *>(XB + kF4NXB * C::ROWS8 * W_xb);   // [NS]  TMA landed
    uint64_t *xb_full = full_in + NS;                             // [NXB]  x pass of a plane complete (XW arrivals)
    uint64_t *xb_empty = xb_full + kF4NXB;                        // [NXB]  consumers have read the plane (YZ_WARPS arrivals)

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int x0 = blockIdx.x * kF4TX, y0 = blockIdx.y * TY;
    const int a0 = blockIdx.z * seg_len, a1 = min(Z, a0 + seg_len);
    const int n_in = (a1 - a0) + 2 * R;          // input steps u = 0..n_in-1 <-> input plane a0 - R + u

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) mbar_init(&full_in[s], 1);
#pragma unroll
        for (int s = 0; s < kF4NXB; s++) { mbar_init(&xb_full[s], XW); mbar_init(&xb_empty[s], C::YZ_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= C::YZ_WARPS) {
        // ======================= producer warps: TMA + x pass =======================
        if constexpr (C::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::PRODUCER_REGS));
        const int xw = warp - C::YZ_WARPS;
        if (xw == 0 && lane == 0) {
#pragma unroll
            for (int u = 0; u < kF4Ahead + kF4NXB; u++)
                if (u < n_in) {
                    mbar_expect_tx(&full_in[u], C::TILE_BYTES);
                    tma_load_3d(IN + u * STAGE, &in_map, x0 - RP, y0 - R, a0 - R + u, &full_in[u]);
                }
        }
        int stage = 0, sphase = 0;               // ring position / parity of plane q
        int fstage = (kF4Ahead + kF4NXB) % NS;   // ring position of plane q + kF4Ahead (refilled at q >= NXB)
        int xbs = 0, xphase = 1;                 // hand-over slot of plane q, parity of its previous use
        for (int q = 0; q < n_in; q++) {
            if (q >= kF4NXB) {
                mbar_wait(&xb_empty[xbs], xphase);     // consumers have read plane q-NXB (and, MRING, the minuend q-NXB-R)
                if (xw == 0 && lane == 0 && q + kF4Ahead < n_in) {
                    mbar_expect_tx(&full_in[fstage], C::TILE_BYTES);
                    tma_load_3d(IN + fstage * STAGE, &in_map, x0 - RP, y0 - R, a0 - R + q + kF4Ahead, &full_in[fstage]);
                }
                if (++fstage == NS) fstage = 0;
            }
            mbar_wait(&full_in[stage], sphase);
            const float *in_t = IN + stage * STAGE;
            float *xb = XB + xbs * (C::ROWS8 * W_xb);
            // warp items are dealt round-robin, rotated by plane so that no warp is always the one with an extra item
            for (int wi = (xw + q) % XW; wi < C::X_WARP_ITEMS; wi += XW) {
                constexpr int HALVES = kF4TX / (4 * KX);
                const int rg = wi / HALVES, half = wi - rg * HALVES;
                const int row = rg * 8 + (lane & 7);
                const int xs = half * 4 + (lane >> 3);
                const float *src = in_t + row * W_in + xs * KX;
                float win[KX + 2 * RP];
#pragma unroll
                for (int k = 0; k < (KX + 2 * RP) / 4; k++) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(src + 4 * k);
                    win[4 * k] = w4.x; win[4 * k + 1] = w4.y; win[4 * k + 2] = w4.z; win[4 * k + 3] = w4.w;
                }
                float2 o[KX / 2];
                seg_x_pk<R, KX, RP - R>(win, o, taps);
                const int xg = x0 + xs * KX;
                if (xg + KX > X) {       // padding columns (x >= X) stay zero in every pass
#pragma unroll
                    for (int k = 0; k < KX / 2; k++) {
                        if (xg + 2 * k >= X) o[k].x = 0.0f;
                        if (xg + 2 * k + 1 >= X) o[k].y = 0.0f;
                    }
                }
                float *dst = xb + row * W_xb + xs * KX;
#pragma unroll
                for (int k = 0; k < KX / 4; k++)
                    *reinterpret_cast<float4 *>(dst + 4 * k) = make_float4(o[2 * k].x, o[2 * k].y, o[2 * k + 1].x, o[2 * k + 1].y);
            }
            fence_proxy_async();     // this warp's reads of the ring stage come before its later TMA refill
            __syncwarp();
            if (lane == 0) mbar_arrive(&xb_full[xbs]);
            if (++stage == NS) { stage = 0; sphase ^= 1; }
            if (++xbs == kF4NXB) { xbs = 0; xphase ^= 1; }
        }
    } else {
        // ======================= consumer warps: y pass, z pass, DoG, stores =======================
        if constexpr (C::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::CONSUMER_REGS));
        constexpr int CG = kF4TX / VX;                       // column groups per tile row (32 or 16)
        const int cg = t % CG, rgy = t / CG;                 // this thread: columns x0 + VX*cg .. +VX-1, rows y0 + KY*rgy .. +KY-1
        const int gx = x0 + VX * cg;
        const int rows_ok = max(0, min(KY, Y - (y0 + rgy * KY)));
        const bool col_ok = gx < pitch;
        VT acc[KY][T];
#pragma unroll
        for (int c = 0; c < KY; c++)
#pragma unroll
            for (int s = 0; s < T; s++) acc[c][s] = zerov<VT>();
        // output position of step u is plane a0 + u - 2R
        long long off = ((long long)(a0 - 2 * R) * Y + (y0 + rgy * KY)) * pitch + gx;
        const long long plane = (long long)pitch * Y;
        const size_t pitch4 = (size_t)pitch * sizeof(float), plane4 = (size_t)plane * sizeof(float);
        // (byte pointers into the output volumes; they run ahead of the first plane by up to 2R planes and are not
        // dereferenced before step 2R)
        char *ob = reinterpret_cast<char *>(out) + off * (long long)sizeof(float), *db = reinterpret_cast<char *>(dog) + off * (long long)sizeof(float);
        int mstage = MRING ? (NS - R % NS) % NS : 0;   // ring position of plane u - R (DoG minuend of the output completed at step u)
        int s_mod = 0;
        int xbs = 0, xphase = 0;                 // hand-over slot of plane u and its parity
        for (int u = 0; u < n_in; u++) {
            const float *xb = XB + xbs * (C::ROWS8 * W_xb);
            mbar_wait(&xb_full[xbs], xphase);
            VT win[KY + 2 * R];
            {
                const float *col = xb + (rgy * KY) * W_xb + VX * cg;
#pragma unroll
                for (int m = 0; m < KY + 2 * R; m++) win[m] = *reinterpret_cast<const VT *>(col + m * W_xb);
            }
            VT mn[KY];
            if (DOG && u >= 2 * R) {
                if (MRING) {
                    const float *mp = IN + mstage * STAGE + (R + rgy * KY) * W_in + RP + VX * cg;
#pragma unroll
                    for (int k = 0; k < KY; k++) mn[k] = *reinterpret_cast<const VT *>(mp + k * W_in);
                } else if (col_ok) {      // the plane left the ring R steps ago: it is an L2 hit
#pragma unroll
                    for (int k = 0; k < KY; k++) if (k < rows_ok) mn[k] = ldgv<VT>(in + off + (long long)k * pitch);
                }
            }
            if (MRING) fence_proxy_async();     // the minuend was read from a ring stage that TMA refills after this arrival
            __syncwarp();
            if (lane == 0) mbar_arrive(&xb_empty[xbs]);
            if (++xbs == kF4NXB) { xbs = 0; xphase ^= 1; }

            VT v[KY];
            conv_segment<R, KY, VT>(win, v, taps);
            VT done[KY];
            // (unrolling the plane loop over the 2R+1 slots instead of dispatching on plane mod 2R+1 removes ~35 instructions per
            // plane but costs 30-40 registers: levels 1-2 us slower, batch 5 % slower -- measured, not shipped)
            F4Dispatch<R, VX, 0, T - 1>::run(s_mod, acc, v, taps, done);
            if (u >= 2 * R && col_ok) {
                // addresses as integers advanced plane by plane (a 64-bit multiply-add per row and stream was 9 % of
                // the kernel's instructions); whole row groups -- all but the last of the volume -- store unpredicated
                if (rows_ok == KY) {
#pragma unroll
                    for (int k = 0; k < KY; k++) {
                        *reinterpret_cast<VT *>(ob + k * pitch4) = done[k];
                        if (DOG) *reinterpret_cast<VT *>(db + k * pitch4) = subv(mn[k], done[k]);   // prev + (-1)*g, fioMultSum
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KY; k++) {
                        if (k < rows_ok) {
                            *reinterpret_cast<VT *>(ob + k * pitch4) = done[k];
                            if (DOG) *reinterpret_cast<VT *>(db + k * pitch4) = subv(mn[k], done[k]);
                        }
                    }
                }
            }
            off += plane;
            ob += plane4; db += plane4;
            s_mod = (s_mod + 1 == T) ? 0 : s_mod + 1;
            if (++mstage == NS) mstage = 0;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
// instances: 64 x TY tiles (TY = 16: two CTAs of 7-8 warps per SM; TY = 32: one CTA of 11-12 warps), 2 columns per
// consumer thread, 16-output x segments
template <int R, int TY> struct F4Pick {
    // producer warps: the three x-pass items of a 64 x 16 tile plane on TWO warps for radii <= 4 (a producer item is
    // ~130 instructions, a consumer warp issues ~250 per plane: with three producers they idled at xb_empty half of the time;
    // measured 39.2 -> 37.2 us at 9 taps, 33.2 -> 32.0 us at 7); radii 5, 6 need two full warpgroups for setmaxnreg
    static constexpr int VX = 2, KX = 16, XW = TY == 16 ? (R <= 4 ? 2 : 4) : 4;
    // The DoG minuend from the ring costs R + 1 more stages.  At 9 taps that is 115 KB per CTA instead of 82 KB: the level
    // (measured with a 4-slot hand-over ring; 68 KB now) itself runs the same (37.0 us either way) but a batch loses 2 % (475 -> 465 us per volume with the minuend read
    // from L2): what shares the SM with a level CTA there are other volumes' kernels and their shared memory.  At 7
    // taps the ring is kept (85 KB with the 3-slot hand-over ring; without it the warm level is 5 % slower and the batch the same)
    static constexpr bool MRING = R <= 3;
    using Cfg = F4Cfg<R, VX, KX, TY, XW, MRING>;
    template <bool DOG> static auto kernel() { return blur_f4_kernel<R, VX, KX, TY, XW, MRING, DOG>; }
};
constexpr int kF4MaxR = 6;

template <int R, int TY>
static cudaError_t set_f4_attr_r()
{
    using P = F4Pick<R, TY>;
    cudaError_t e = cudaFuncSetAttribute(P::template kernel<true>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::Cfg::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(P::template kernel<false>(), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::Cfg::SMEM);
}
template <int TY>
static cudaError_t init_blur4_attrs_ty()
{
    cudaError_t e;
    if ((e = set_f4_attr_r<1, TY>()) != cudaSuccess) return e;
    if ((e = set_f4_attr_r<2, TY>()) != cudaSuccess) return e;
    if ((e = set_f4_attr_r<3, TY>()) != cudaSuccess) return e;
    if ((e = set_f4_attr_r<4, TY>()) != cudaSuccess) return e;
    if ((e = set_f4_attr_r<5, TY>()) != cudaSuccess) return e;
    return set_f4_attr_r<6, TY>();
}
static cudaError_t init_blur4_attrs()
{
    cudaError_t e = init_blur4_attrs_ty<16>();
    if (e != cudaSuccess) return e;
    return init_blur4_attrs_ty<32>();
}

// One-kernel level: in -> out (+ dog).  Returns false (nothing launched) when the tensor map cannot be encoded.
// z segments: as many as fill every SM with its resident CTAs (`want_ctas`, 0 = occupancy x SMs), but no shorter
// than max(8, 2R) planes -- each segment re-does the x and y passes of 2R halo planes.
template <int R, int TY>
static bool launch_blur_f4(cudaStream_t st, const float *in, float *out, float *dog, int X, int Y, int Z, int pitch,
                           const float *taps, int sm_count, int want_ctas, cudaError_t *err)
{
    using P = F4Pick<R, TY>;
    using C = typename P::Cfg;
    CUtensorMap map;
    if (!make_volume_map_box(&map, in, Y, Z, pitch, C::W_in, C::ROWS)) return false;
    TapsSmall t = make_taps_small(taps, 2 * R + 1);
    int tx = (pitch + kF4TX - 1) / kF4TX, ty = (Y + TY - 1) / TY;
    int want = want_ctas > 0 ? want_ctas : sm_count * C::MIN_CTAS;
    int n_seg = want / (tx * ty);
    int min_len = 2 * R > 8 ? 2 * R : 8;
    int max_seg = Z / min_len;
    if (n_seg > max_seg) n_seg = max_seg;
    if (n_seg < 1) n_seg = 1;
    int seg_len = (Z + n_seg - 1) / n_seg;
    n_seg = (Z + seg_len - 1) / seg_len;
    if (n_seg > 65535) return false;
    dim3 grid(tx, ty, n_seg);
    if (dog) P::template kernel<true>()<<<grid, C::THREADS, C::SMEM, st>>>(map, in, out, dog, X, Y, Z, pitch, seg_len, t);
    else P::template kernel<false>()<<<grid, C::THREADS, C::SMEM, st>>>(map, in, out, nullptr, X, Y, Z, pitch, seg_len, t);
    *err = cudaGetLastError();
    return true;
}

} // namespace s3d
