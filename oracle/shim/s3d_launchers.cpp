/*
 * s3d_launchers.cpp -- the reference's four CUDA launchers re-implemented on the stage level of include/s3d.h
 * (TEST INFRASTRUCTURE: proof of INTEGRATION.md level B, never part of the product library).
 *
 * The reference crosses from src_common into CUDA through exactly four C++ functions, declared in
 * R/cuda_common/SIFT_cuda_Tools.cuh and called from the four *_interleave dispatch sites when featExtract runs
 * with -dN:
 *   blur_3d_simpleborders_CUDA_Row_Col_Shared_mem  (.cuh:69-76,  called GaussBlur3D.cpp:1244)  -> s3d_blur3d
 *   fioCudaMultSum                                 (.cuh:213-217, called FeatureIO.cpp:1942)    -> s3d_dog
 *   SubSampleInterpolateCuda                       (.cuh:202-205, called FeatureIO.cpp:1559)    -> s3d_subsample2
 *   detectExtrema4D_test_cuda                      (.cuh:32-38,  called MultiScale.cpp:1531)    -> s3d_detect
 * This file defines those four symbols with their original signatures; oracle/Makefile links it with the
 * UNMODIFIED src_common + featExtract.cpp objects, the real cudart and lib3dsift_b200.so into
 * oracle/_ref/featExtract_ref_s3d.  tests/test_gpu_cli.py then checks that `featExtract_ref_s3d -d0 in.nii` writes
 * the same bytes as the reference's CPU path (`featExtract_ref in.nii`).
 *
 * Mirror contract: the reference keeps every volume valid in host memory (pfVectors) and mirrors launcher outputs
 * to it (SIFT_cuda_Tools.cu:216), while its device copies are not always current (-2+ never uploads the doubled
 * volume, the shipped blur clobbers its input's device copy: SURVEY.md section 0).  The shim therefore treats the
 * host copy as the source of truth: inputs are uploaded from pfVectors into pitched device buffers (the layout of
 * the fast kernels), outputs are mirrored to pfVectors and, when present, to the dense d_pfVectors.
 * All copies are enqueued on the context's own stream (s3d_stream), which is where the stage calls enqueue their
 * kernels: it is a non-blocking stream, so the legacy default stream would not be ordered against them.
 * Errors: like the reference's gpuErrchk (SIFT_cuda_Tools.cuh:13-21) -- message and exit -- because the original
 * signatures cannot report them.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "FeatureIO.h"
#include "LocationValue.h"
#include "PpImage.h"
#include "s3d.h"
#include "GaussBlur3D.h"
#include "MultiScale.h"

// debugging aid: S3D_SHIM_CPU="blur,dog,sub,det" routes the named stages back to the reference's CPU functions
static bool shim_cpu(const char *stage) { const char *e = getenv("S3D_SHIM_CPU"); return e && strstr(e, stage); }
int blur_3d_simpleborders(FEATUREIO &fio1, FEATUREIO &fioTemp, FEATUREIO &fio2, int iFeature, PpImage &ppImgFilter);

namespace {

s3d_ctx *g_ctx = nullptr;
int g_dev = -1;

void die(const char *what, const char *detail)
{
    fprintf(stderr, "s3d shim: %s failed: %s\n", what, detail ? detail : "");
    exit(2);
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) die(#call, cudaGetErrorString(e_)); } while (0)
#define S3(call) do { s3d_status s_ = (call); if (s_ != S3D_OK) die(#call, s3d_last_error(g_ctx)); } while (0)

s3d_ctx *ctx_for(int dev)
{
    if (dev < 0) dev = 0;
    if (g_ctx && g_dev != dev) { s3d_ctx_destroy(g_ctx); g_ctx = nullptr; }
    if (!g_ctx) {
        if (s3d_ctx_create(dev, &g_ctx) != S3D_OK) die("s3d_ctx_create", g_ctx ? s3d_last_error(g_ctx) : "no context");
        g_dev = dev;
    }
    CU(cudaSetDevice(dev));
    return g_ctx;
}

struct DevVol {
    float *p = nullptr;
    int X = 0, Y = 0, Z = 0, pitch = 0;
    size_t bytes() const { return sizeof(float) * (size_t)pitch * Y * Z; }
};

DevVol dev_alloc(int X, int Y, int Z)
{
    DevVol v;
    v.X = X; v.Y = Y; v.Z = Z; v.pitch = (X + 7) / 8 * 8;
    CU(cudaMalloc((void **)&v.p, v.bytes() + 256));
    // every copy / memset goes through the context's own (non-blocking) stream: the legacy default stream would
    // not be ordered against the kernels the stage calls enqueue there
    CU(cudaMemsetAsync(v.p, 0, v.bytes(), (cudaStream_t)s3d_stream(g_ctx)));         // padding columns are zero
    return v;
}

void check_scalar(const FEATUREIO &f, const char *who)
{
    if (f.t != 1 || f.iFeaturesPerVector != 1) die(who, "only scalar volumes (t == 1, one feature per vector) are supported");
    if (!f.pfVectors) die(who, "volume without host data");
}

DevVol upload(const FEATUREIO &f, const char *who)
{
    check_scalar(f, who);
    DevVol v = dev_alloc(f.x, f.y, f.z);
    CU(cudaMemcpy2DAsync(v.p, sizeof(float) * v.pitch, f.pfVectors, sizeof(float) * f.x, sizeof(float) * f.x, (size_t)f.y * f.z, cudaMemcpyHostToDevice,
                         (cudaStream_t)s3d_stream(g_ctx)));
    return v;
}

// mirror a pitched result to the volume's host copy and (when allocated) its dense device copy
void mirror(const DevVol &v, FEATUREIO &f)
{
    cudaStream_t st = (cudaStream_t)s3d_stream(g_ctx);
    CU(cudaMemcpy2DAsync(f.pfVectors, sizeof(float) * f.x, v.p, sizeof(float) * v.pitch, sizeof(float) * f.x, (size_t)f.y * f.z, cudaMemcpyDeviceToHost, st));
    if (f.d_pfVectors)
        CU(cudaMemcpy2DAsync(f.d_pfVectors, sizeof(float) * f.x, v.p, sizeof(float) * v.pitch, sizeof(float) * f.x, (size_t)f.y * f.z, cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
}

} // namespace

int blur_3d_simpleborders_CUDA_Row_Col_Shared_mem(FEATUREIO &fio1, FEATUREIO &fioTemp, FEATUREIO &fio2, int iFeature,
                                                  PpImage &ppImgFilter, int best_device_id)
{
    if (shim_cpu("blur")) return blur_3d_simpleborders(fio1, fioTemp, fio2, iFeature, ppImgFilter);
    s3d_ctx *ctx = ctx_for(best_device_id);
    if (fio1.x != fio2.x || fio1.y != fio2.y || fio1.z != fio2.z) return 0;      // the reference's "dimension mismatch"
    check_scalar(fio2, "blur");
    DevVol in = upload(fio1, "blur"), tmp = dev_alloc(fio1.x, fio1.y, fio1.z), out = dev_alloc(fio1.x, fio1.y, fio1.z);
    const float *taps = (const float *)ppImgFilter.ImageRow(0);
    S3(s3d_blur3d(ctx, in.p, tmp.p, out.p, in.X, in.Y, in.Z, in.pitch, taps, ppImgFilter.Cols(), nullptr));
    S3(s3d_sync(ctx));
    mirror(out, fio2);
    cudaFree(in.p); cudaFree(tmp.p); cudaFree(out.p);
    return 1;
}

int fioCudaMultSum(FEATUREIO &fioIn1, FEATUREIO &fioIn2, FEATUREIO &fioOut, const float &fMultIn2)
{
    if (shim_cpu("dog")) return fioMultSum(fioIn1, fioIn2, fioOut, fMultIn2);
    s3d_ctx *ctx = ctx_for(fioOut.device);
    if (fioIn1.x != fioOut.x || fioIn1.y != fioOut.y || fioIn1.z != fioOut.z ||
        fioIn2.x != fioOut.x || fioIn2.y != fioOut.y || fioIn2.z != fioOut.z) return 0;
    check_scalar(fioOut, "multsum");
    if (fMultIn2 != -1.0f) die("fioCudaMultSum", "the pyramid only ever calls this with a factor of -1 (DoG)");
    DevVol a = upload(fioIn1, "multsum"), b = upload(fioIn2, "multsum"), out = dev_alloc(fioOut.x, fioOut.y, fioOut.z);
    S3(s3d_dog(ctx, a.p, b.p, out.p, a.X, a.Y, a.Z, a.pitch));
    S3(s3d_sync(ctx));
    mirror(out, fioOut);
    cudaFree(a.p); cudaFree(b.p); cudaFree(out.p);
    return 1;
}

int SubSampleInterpolateCuda(FEATUREIO &fioIn, FEATUREIO &fioOut, int best_device_id)
{
    if (shim_cpu("sub")) return fioSubSampleInterpolate(fioIn, fioOut);
    s3d_ctx *ctx = ctx_for(best_device_id);
    if (fioOut.x != fioIn.x / 2 || fioOut.y != fioIn.y / 2 || fioOut.z != fioIn.z / 2) return 0;
    check_scalar(fioOut, "subsample");
    DevVol in = upload(fioIn, "subsample"), out = dev_alloc(fioOut.x, fioOut.y, fioOut.z);
    S3(s3d_subsample2(ctx, in.p, in.X, in.Y, in.Z, in.pitch, out.p, out.pitch));
    S3(s3d_sync(ctx));
    mirror(out, fioOut);
    cudaFree(in.p); cudaFree(out.p);
    return 1;
}

void detectExtrema4D_test_cuda(FEATUREIO &inputH, FEATUREIO &inputC, FEATUREIO &fioSumOfSign,
                               LOCATION_VALUE_XYZ_ARRAY &lvaMinima, LOCATION_VALUE_XYZ_ARRAY &lvaMaxima, int best_device_id)
{
    (void)fioSumOfSign;      // the reference's intermediate sign-sum volume: nobody reads it after the launcher
    if (shim_cpu("det")) { detectExtrema4D_test(&inputH, &inputC, 0, lvaMinima, lvaMaxima); return; }
    s3d_ctx *ctx = ctx_for(best_device_id);
    lvaMinima.iCount = 0;
    lvaMaxima.iCount = 0;
    if (inputC.x < 3 || inputC.y < 3 || inputC.z < 3) return;
    DevVol h = upload(inputH, "detect"), c = upload(inputC, "detect");
    long long cap64 = (long long)inputC.x * inputC.y * inputC.z;
    const int cap = (int)(cap64 > (1 << 22) ? (1 << 22) : cap64);
    s3d_cand *d_min = nullptr, *d_max = nullptr;
    int *d_n = nullptr;
    CU(cudaMalloc((void **)&d_min, sizeof(s3d_cand) * (size_t)cap));
    CU(cudaMalloc((void **)&d_max, sizeof(s3d_cand) * (size_t)cap));
    CU(cudaMalloc((void **)&d_n, 2 * sizeof(int)));
    S3(s3d_detect(ctx, h.p, c.p, c.X, c.Y, c.Z, c.pitch, d_min, d_n, d_max, d_n + 1, cap));
    S3(s3d_sync(ctx));
    int n[2] = { 0, 0 };
    cudaStream_t st = (cudaStream_t)s3d_stream(ctx);
    CU(cudaMemcpyAsync(n, d_n, sizeof(n), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n[0] > cap || n[1] > cap) die("s3d_detect", "candidate capacity exceeded");
    // s3d_cand and LOCATION_VALUE_XYZ have the same layout (x, y, z, value); lists arrive in raster order
    static_assert(sizeof(s3d_cand) == sizeof(LOCATION_VALUE_XYZ), "candidate record layouts differ");
    CU(cudaMemcpyAsync(lvaMinima.plvz, d_min, sizeof(s3d_cand) * (size_t)n[0], cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(lvaMaxima.plvz, d_max, sizeof(s3d_cand) * (size_t)n[1], cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    lvaMinima.iCount = n[0];
    lvaMaxima.iCount = n[1];
    cudaFree(h.p); cudaFree(c.p); cudaFree(d_min); cudaFree(d_max); cudaFree(d_n);
}
