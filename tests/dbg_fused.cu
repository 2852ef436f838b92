#include <cstdio>
#include <cstring>
#include <vector>
#include <cmath>
#include "../3d_sift_cuda_b200/csrc/s3d_blur_fused.cuh"
using namespace s3d;
template <int R>
int run(int X, int Y, int Z, int pitch)
{
    std::vector<float> h((size_t)pitch * Y * Z, 0.f);
    for (int z = 0; z < Z; z++) for (int y = 0; y < Y; y++) for (int x = 0; x < X; x++) h[((size_t)z * Y + y) * pitch + x] = (float)((x * 7 + y * 13 + z * 29) % 101) * 0.37f + 1.f;
    float taps[2 * R + 1]; float s = 0; for (int j = 0; j <= 2 * R; j++) { taps[j] = expf(-0.5f * (j - R) * (j - R) / (R * R / 4.0f + 1)); s += taps[j]; } for (int j = 0; j <= 2 * R; j++) taps[j] /= s;
    size_t n = h.size();
    float *d, *o, *g; cudaMalloc(&d, n * 4); cudaMalloc(&o, n * 4); cudaMalloc(&g, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice); cudaMemset(o, 0, n * 4); cudaMemset(g, 0, n * 4);
    CUtensorMap map; if (!make_volume_map(&map, d, Y, Z, pitch, R)) { printf("map failed\n"); return 1; }
    cudaError_t e = launch_blur_fused<R>(0, map, d, o, g, X, Y, Z, pitch, taps, 148, 0);
    printf("R=%d launch: %s; ", R, cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    printf("sync: %s; ", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> got(n), t1(n, 0.f), t2(n, 0.f), t3(n, 0.f);
    cudaMemcpy(got.data(), o, n * 4, cudaMemcpyDeviceToHost);
    auto pass = [&](std::vector<float> &src, std::vector<float> &dst, int axis) {
        for (int z = 0; z < Z; z++) for (int y = 0; y < Y; y++) for (int x = 0; x < X; x++) {
            volatile float acc = 0; for (int j = 0; j <= 2 * R; j++) { int xx = x, yy = y, zz = z; int off = j - R; if (axis == 0) xx += off; else if (axis == 1) yy += off; else zz += off;
                float v = (xx >= 0 && xx < X && yy >= 0 && yy < Y && zz >= 0 && zz < Z) ? src[((size_t)zz * Y + yy) * pitch + xx] : 0.f; volatile float p = taps[j] * v; acc = acc + p; }
            dst[((size_t)z * Y + y) * pitch + x] = acc; } };
    pass(h, t1, 0); pass(t1, t2, 1); pass(t2, t3, 2);
    size_t bad = 0; for (size_t i = 0; i < n; i++) if (memcmp(&got[i], &t3[i], 4)) bad++;
    printf("mismatches %zu of %zu\n", bad, n);
    cudaFree(d); cudaFree(o); cudaFree(g);
    return 0;
}
int main()
{
    run<3>(48, 40, 36, 48); run<4>(48, 40, 36, 48); run<8>(48, 40, 36, 48); run<5>(37, 29, 23, 40); run<1>(9, 7, 5, 16);
    run<8>(182, 218, 182, 184);
    return 0;
}
