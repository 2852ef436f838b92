// s3d_match.cuh -- exact k-nearest-neighbour search between two feature sets on the descriptor distance of the
// reference.
//
// The reference matches descriptors with Feature3DInfo::DistSqrPCs (R/src_common/MultiScale.h:60-73: sequential
// float sum of squared differences over the 64 descriptor entries) and finds neighbours with FLANN kd-trees
// (R/feat_common/featMatchUtilities.cpp:1449-1455, 1559, 1612: 8 trees, 64 checks, g_nn neighbours, sorted) -- an
// approximate search.  Here the search is exhaustive: every (query, database) distance is evaluated with exactly
// DistSqrPCs' arithmetic (difference, product and sum rounded separately, entries in order; the translation unit is
// built with -fmad=false), and the k smallest are kept in (distance, index) order, so the result is the exact
// answer FLANN approximates and is reproducible bit for bit.  Descriptors are ranks 0..63 stored as floats, so all
// sums are exact integers (max 64 * 63^2 = 254016 < 2^24) and ties are real: they are broken by the lower index.
//
//   * match_partial_kernel: a CTA owns 128 queries (one per thread, descriptor in registers) and one chunk of the
//     database; database descriptors are staged through shared memory 64 at a time (float4 broadcast reads, no
//     bank conflicts: every lane reads the same address); each thread keeps its k best of the chunk sorted in
//     registers (insertion with static indices);
//   * match_merge_kernel: one thread per query merges the per-chunk lists (they are sorted, chunk order = index
//     order, so a stable k-way pick keeps the (distance, index) order).
// HBM traffic is nA*256 + chunks*nB*256 bytes; the work is nA*nB*192 FP32 operations -- FP32-issue bound, no
// tensor cores: the distance is a sum of squares of differences with prescribed rounding, not a dot product.
#pragma once
#include <cuda_runtime.h>
#include "../../include/s3d.h"

namespace s3d {

constexpr int kMatchMaxK = 16, kMatchThreads = 128, kMatchTile = 64;

template <int K>
__device__ __forceinline__ void match_insert(float (&bd)[K], int (&bi)[K], float d, int j)
{
    // candidates arrive in increasing index order inside a chunk, so `<` keeps the lower index on ties
    if (d < bd[K - 1]) {
        bd[K - 1] = d; bi[K - 1] = j;
#pragma unroll
        for (int s = K - 1; s > 0; s--) {
            if (bd[s] < bd[s - 1]) {
                float td = bd[s]; bd[s] = bd[s - 1]; bd[s - 1] = td;
                int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
            }
        }
    }
}

// part_d / part_i: [n_chunks][nA][K]
template <int K>
__global__ void __launch_bounds__(kMatchThreads)
match_partial_kernel(const s3d_feature *__restrict__ fa, int nA, const s3d_feature *__restrict__ fb, int nB, int chunk,
                     float *__restrict__ part_d, int *__restrict__ part_i)
{
    __shared__ __align__(16) float tile[kMatchTile][64];
    const int q = blockIdx.x * kMatchThreads + threadIdx.x;
    const int b0 = blockIdx.y * chunk, b1 = min(nB, b0 + chunk);
    float a[64];
    const int qa = q < nA ? q : nA - 1;
#pragma unroll
    for (int i = 0; i < 64; i++) a[i] = fa[qa].pc[i];
    float bd[K]; int bi[K];
#pragma unroll
    for (int s = 0; s < K; s++) { bd[s] = __int_as_float(0x7f800000); bi[s] = -1; }
    for (int t0 = b0; t0 < b1; t0 += kMatchTile) {
        const int nt = min(kMatchTile, b1 - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * 64; e += kMatchThreads) tile[e >> 6][e & 63] = fb[t0 + (e >> 6)].pc[e & 63];
        __syncthreads();
        for (int j = 0; j < nt; j++) {
            float sum = 0.0f;                               // float fSumSqr = 0
#pragma unroll
            for (int i4 = 0; i4 < 16; i4++) {
                const float4 b = *reinterpret_cast<const float4 *>(&tile[j][4 * i4]);
                float df;
                df = a[4 * i4] - b.x;     sum = sum + df * df;      // fDiff = a - b; fSumSqr += fDiff*fDiff
                df = a[4 * i4 + 1] - b.y; sum = sum + df * df;
                df = a[4 * i4 + 2] - b.z; sum = sum + df * df;
                df = a[4 * i4 + 3] - b.w; sum = sum + df * df;
            }
            match_insert<K>(bd, bi, sum, t0 + j);
        }
    }
    if (q < nA) {
        const size_t o = ((size_t)blockIdx.y * nA + q) * K;
#pragma unroll
        for (int s = 0; s < K; s++) { part_d[o + s] = bd[s]; part_i[o + s] = bi[s]; }
    }
}

template <int K>
__global__ void match_merge_kernel(const float *__restrict__ part_d, const int *__restrict__ part_i, int nA, int n_chunks, int k_out,
                                   int *__restrict__ out_idx, float *__restrict__ out_dist)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nA) return;
    float bd[K]; int bi[K];
#pragma unroll
    for (int s = 0; s < K; s++) { bd[s] = __int_as_float(0x7f800000); bi[s] = -1; }
    for (int c = 0; c < n_chunks; c++) {
        const size_t o = ((size_t)c * nA + q) * K;
        for (int s = 0; s < K; s++) {
            const int j = part_i[o + s];
            if (j < 0) break;
            match_insert<K>(bd, bi, part_d[o + s], j);     // chunk lists are sorted and chunks ascend in index: ties keep the lower index
        }
    }
#pragma unroll
    for (int s = 0; s < K; s++)
        if (s < k_out) { out_idx[(size_t)q * k_out + s] = bi[s]; out_dist[(size_t)q * k_out + s] = bd[s]; }
}

// launches both kernels for k rounded up to 1, 2, 4, 8 or 16 list entries
static cudaError_t launch_match(cudaStream_t st, const s3d_feature *d_a, int nA, const s3d_feature *d_b, int nB, int k,
                                int n_chunks, int chunk, float *part_d, int *part_i, int *d_idx, float *d_dist)
{
    dim3 grid((unsigned)((nA + kMatchThreads - 1) / kMatchThreads), (unsigned)n_chunks);
    const int mblocks = (nA + 127) / 128;
#define S3D_MATCH_CASE(KK)                                                                                          \
    match_partial_kernel<KK><<<grid, kMatchThreads, 0, st>>>(d_a, nA, d_b, nB, chunk, part_d, part_i);              \
    match_merge_kernel<KK><<<mblocks, 128, 0, st>>>(part_d, part_i, nA, n_chunks, k, d_idx, d_dist)
    if (k <= 1) { S3D_MATCH_CASE(1); }
    else if (k <= 2) { S3D_MATCH_CASE(2); }
    else if (k <= 4) { S3D_MATCH_CASE(4); }
    else if (k <= 8) { S3D_MATCH_CASE(8); }
    else { S3D_MATCH_CASE(16); }
#undef S3D_MATCH_CASE
    return cudaGetLastError();
}
static inline int match_list_len(int k) { return k <= 1 ? 1 : k <= 2 ? 2 : k <= 4 ? 4 : k <= 8 ? 8 : 16; }

} // namespace s3d
