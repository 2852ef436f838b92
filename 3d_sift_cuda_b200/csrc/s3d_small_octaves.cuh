// s3d_small_octaves.cuh -- all voxel stages of the small octaves in ONE kernel.
//
// An octave with a few thousand voxels costs microseconds of arithmetic but, as separate kernels, a
// dependent chain of ~13 launches (5 levels x (x+y pass, z pass) + subsample + detection) of ~3 us each,
// and octave o+1 cannot start before level 3 of octave o: for the MNI volume the last three octaves
// (22x27x22, 11x13x11, 5x6x5) kept the GPU almost idle for >100 us at the end of every extraction.
// Here one thread-block cluster (8 CTAs, hardware cluster barrier between passes) walks those octaves:
// per level the x, y and z(+DoG) passes, the 2x subsample after level 3, then the 53-neighbour detection
// into the same candidate lists the large-octave path fills.  Arithmetic and zero padding are those of
// the reference loop (filter_1d: fSum = 0; fSum += w[j]*v, GaussBlur3D.cpp:43-61).
//
// STATUS: optional (S3D_SMALL=1), bit-exact, but measured at ~120 us for the last three MNI octaves --
// the same as the launch chain it replaces: every pass is still one L2 round trip plus a cluster
// barrier (~2 us).  Kept as the starting point for a shared-memory-resident version.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include "s3d_voxel.cuh"
#include "s3d_keypoint.cuh"

namespace s3d {

namespace cg = cooperative_groups;

constexpr int kSmallClusterCtas = 8;
constexpr int kSmallThreads = 1024;
constexpr long long kSmallMaxVoxels = 20000;   // octaves up to this many (pitched) voxels go through this kernel

struct SmallOctArgs {
    int o_first, n_oct;
    int n_taps[5];
    float taps[5][2 * kMaxFastR + 1];
    float *tmp[kMaxOct];          // per-octave scratch volume
    s3d_cand *cand_raw;           // [list][cand_cap]
    int *counts;                  // [list]
    int cand_cap;
};

__global__ void __cluster_dims__(kSmallClusterCtas, 1, 1) __launch_bounds__(kSmallThreads)
small_octaves_kernel(const __grid_constant__ PyramidDesc pyr, const __grid_constant__ SmallOctArgs a)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int NT = kSmallClusterCtas * kSmallThreads;
    const int gt = (int)cluster.block_rank() * kSmallThreads + threadIdx.x;

    for (int o = a.o_first; o < a.n_oct; o++) {
        const OctaveDesc &od = pyr.oct[o];
        const int X = od.X, Y = od.Y, Z = od.Z, pitch = od.pitch;
        const int plane = pitch * Y, n = plane * Z;
        float *tmp = a.tmp[o];
        for (int j = 1; j < 6; j++) {
            const float *gin = od.g[j - 1];
            float *gout = const_cast<float *>(od.g[j]);
            float *dog = const_cast<float *>(od.d[j - 1]);
            const float *w = a.taps[j - 1];
            const int nt = a.n_taps[j - 1], r = nt / 2;
            for (int i = gt; i < n; i += NT) {          // x: gin -> gout
                int x = i % pitch;
                // all taps' loads first (independent, one L2 round trip), then the ordered accumulation
                float v[2 * kMaxFastR + 1];
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++) {
                    int p = x + t - r;
                    v[t] = (t < nt && x < X && p >= 0 && p < X) ? gin[i + t - r] : 0.0f;
                }
                float acc = 0.0f;
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++)
                    if (t < nt) acc = acc + w[t] * v[t];
                gout[i] = (x < X) ? acc : 0.0f;
            }
            cluster.sync();
            for (int i = gt; i < n; i += NT) {          // y: gout -> tmp
                int y = (i / pitch) % Y;
                float v[2 * kMaxFastR + 1];
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++) {
                    int p = y + t - r;
                    v[t] = (t < nt && p >= 0 && p < Y) ? gout[i + (t - r) * pitch] : 0.0f;
                }
                float acc = 0.0f;
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++)
                    if (t < nt) acc = acc + w[t] * v[t];
                tmp[i] = acc;
            }
            cluster.sync();
            for (int i = gt; i < n; i += NT) {          // z: tmp -> gout, DoG
                int z = i / plane;
                float v[2 * kMaxFastR + 1];
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++) {
                    int p = z + t - r;
                    v[t] = (t < nt && p >= 0 && p < Z) ? tmp[i + (t - r) * plane] : 0.0f;
                }
                const float g_in = gin[i];
                float acc = 0.0f;
#pragma unroll
                for (int t = 0; t < 2 * kMaxFastR + 1; t++)
                    if (t < nt) acc = acc + w[t] * v[t];
                gout[i] = acc;
                dog[i] = g_in + (-1.0f) * acc;
            }
            cluster.sync();
            if (j == 3 && o + 1 < a.n_oct) {            // level 3 -> level 0 of the next octave (2x2x2 mean)
                const OctaveDesc &nx = pyr.oct[o + 1];
                float *g0n = const_cast<float *>(nx.g[0]);
                const int on = nx.pitch * nx.Y * nx.Z;
                for (int i = gt; i < on; i += NT) {
                    int x = i % nx.pitch, y = (i / nx.pitch) % nx.Y, z = i / (nx.pitch * nx.Y);
                    float rr = 0.0f;
                    if (x < nx.X) {
                        const float *p0 = gout + ((2 * z) * Y + 2 * y) * pitch + 2 * x;
                        const float *p1 = p0 + plane;
                        float s = 0.0f;
                        s = s + (((p0[0] + p0[pitch]) + p0[1]) + p0[pitch + 1]);
                        if (2 * z + 1 < Z) {
                            s = s + (((p1[0] + p1[pitch]) + p1[1]) + p1[pitch + 1]);
                            s = s * 0.125f;
                        } else {
                            s = s * 0.25f;
                        }
                        rr = s;
                    }
                    g0n[i] = rr;
                }
                // visible to everyone after the cluster barriers of the following level
            }
        }
        // ---- detection on centre levels 1..3 (reference MultiScale.cpp:2260-2524)
        for (int c = 1; c <= 3; c++) {
            const float *finer = od.d[c - 1], *centre = od.d[c];
            const int l0 = (o * 3 + (c - 1)) * 2;
            const int nx_ = X - 2, ny_ = Y - 2, nz_ = od.own1 - od.own0;
            const int ni = nx_ * ny_ * Z;
            (void)nz_;
            for (int q = gt; q < ni; q += NT) {
                int x = 1 + q % nx_, y = 1 + (q / nx_) % ny_, z = q / (nx_ * ny_);
                if (z < 1 || z > Z - 2 || z < od.own0 || z >= od.own1) continue;
                int i = z * plane + y * pitch + x;
                float cv = centre[i];
                bool mx = true, mn = true;
                for (int dz = -1; dz <= 1 && (mx || mn); dz++)
                    for (int dy = -1; dy <= 1 && (mx || mn); dy++) {
                        const float *row = centre + i + dz * plane + dy * pitch;
                        float p = row[-1], qv = row[0], s = row[1];
                        if (dz == 0 && dy == 0) qv = p;
                        mx = mx && (p < cv) && (qv < cv) && (s < cv);
                        mn = mn && (p > cv) && (qv > cv) && (s > cv);
                        row = finer + i + dz * plane + dy * pitch;
                        p = row[-1]; qv = row[0]; s = row[1];
                        mx = mx && (p < cv) && (qv < cv) && (s < cv);
                        mn = mn && (p > cv) && (qv > cv) && (s > cv);
                    }
                if (mx) { int k = atomicAdd(a.counts + l0 + 1, 1); if (k < a.cand_cap) a.cand_raw[(size_t)(l0 + 1) * a.cand_cap + k] = s3d_cand{ x, y, z, cv }; }
                if (mn) { int k = atomicAdd(a.counts + l0, 1); if (k < a.cand_cap) a.cand_raw[(size_t)l0 * a.cand_cap + k] = s3d_cand{ x, y, z, cv }; }
            }
        }
        cluster.sync();
    }
}

} // namespace s3d
