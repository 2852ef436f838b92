/*
 * fake_cudart.cpp -- host-memory stand-ins for the CUDA runtime calls the reference's
 * *CPU* path makes (TEST INFRASTRUCTURE ONLY, linked into oracle/_ref only).
 *
 * Even without -d the reference allocates and mirrors every volume on "device 0"
 * (FeatureIO.cpp:384-387, 1549-1552, 1857-1860).  The oracle must
 * run on GPU-less hosts and must not touch the GPU the product is being measured on, so
 * these seven symbols are backed by malloc/memcpy.  The four CUDA launchers are
 * unreachable when best_device_id == -1; they abort if ever called.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "FeatureIO.h"
#include "LocationValue.h"
#include "PpImage.h"

extern "C" {
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, enum cudaMemcpyKind) { if (n) memcpy(dst, src, n); return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "fake cudart (oracle build)"; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(struct cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); return cudaSuccess; }
#ifdef cudaGetDeviceProperties
#undef cudaGetDeviceProperties
#endif
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp *p, int d) { return cudaGetDeviceProperties_v2(p, d); }
}

static void unreachable(const char *what)
{
    fprintf(stderr, "oracle/_ref: %s called on the CPU-only oracle build\n", what);
    abort();
}

int blur_3d_simpleborders_CUDA_Row_Col_Shared_mem(FEATUREIO &, FEATUREIO &, FEATUREIO &, int, PpImage &, int)
{ unreachable("blur_3d_simpleborders_CUDA_Row_Col_Shared_mem"); return 0; }
int fioCudaMultSum(FEATUREIO &, FEATUREIO &, FEATUREIO &, const float &)
{ unreachable("fioCudaMultSum"); return 0; }
int SubSampleInterpolateCuda(FEATUREIO &, FEATUREIO &, int)
{ unreachable("SubSampleInterpolateCuda"); return 0; }
void detectExtrema4D_test_cuda(FEATUREIO &, FEATUREIO &, FEATUREIO &, LOCATION_VALUE_XYZ_ARRAY &, LOCATION_VALUE_XYZ_ARRAY &, int)
{ unreachable("detectExtrema4D_test_cuda"); }
