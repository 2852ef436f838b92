// s3d_tma.cuh -- mbarrier / TMA (cp.async.bulk.tensor) helpers and tensor-map encoding for pitched fp32 volumes.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3d {

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
// Orders this thread's earlier generic-proxy accesses to shared memory (LDS / STS) before later async-proxy
// accesses (TMA) to the same bytes.  Needed by every thread that READ a staging buffer with LDS before it signals
// (mbarrier arrive, __syncthreads) that the buffer may be refilled by TMA: without it the refill can overtake
// loads that are still queued -- seen as sporadically wrong DoG minuends at 512^3+ when other kernels shared the SMs.
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one (box_w, box_h, 1) box of a 3-D tensor map into shared memory; elements outside the tensor are zero-filled
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled()
{
    // resolved once; the initialisation of a function-local static is thread safe (contexts live on several host threads)
    static PFN_encodeTiled fn = []() -> PFN_encodeTiled {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (PFN_encodeTiled)p;
        return nullptr;
    }();
    return fn;
}

// tensor map of a pitched fp32 volume (x fastest) with a (box_w, box_h, 1) box; false if the driver cannot encode it
static bool make_volume_map_box(CUtensorMap *map, const float *vol, int Y, int Z, int pitch, int box_w, int box_h)
{
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t gdim[3] = { (cuuint64_t)pitch, (cuuint64_t)Y, (cuuint64_t)Z };
    cuuint64_t gstr[2] = { (cuuint64_t)pitch * 4, (cuuint64_t)pitch * Y * 4 };
    cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)vol, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

} // namespace s3d
