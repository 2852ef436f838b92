// s3d_engine.cu -- context, pyramid plan, stream/graph orchestration and the C-ABI of include/s3d.h.
//
// One context = one device + one stream + one resident pyramid plan.  The whole extraction (pre-step,
// pyramid, DoG, detection, refinement, orientation, descriptors) is enqueued on the stream without any
// host round trip -- every data-dependent count (candidates, keypoints, feature rows) stays in device
// memory and the kernels size their own work from it -- and is replayed as one CUDA graph per
// (shape, options).  The reference instead crosses PCIe ~17 times per octave with blocking
// cudaMemcpy.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <string>
#include <vector>

#include "../../include/s3d.h"
#include "s3d_voxel.cuh"
#include "s3d_blur2.cuh"
#include "s3d_blur4.cuh"
#include "s3d_tiny.cuh"
#include "s3d_keypoint.cuh"
#include "s3d_match.cuh"

using namespace s3d;

// ---------------------------------------------------------------------------------------------------
// host-side parameter arithmetic (kept on the host, as in the reference)
// ---------------------------------------------------------------------------------------------------

// calculate_gaussian_filter_size (reference GaussianMask.cpp:12-57).  The reference is C++: exp() on a
// float argument is the float overload, i.e. expf.
static int filter_size(float fSigma, float fMinValue)
{
    float fPower = 0.0f;
    float fValue = expf(fPower);
    float fCurVolume = 1, fNewVolume = 1;
    int i = 0;
    if (fSigma == 0) return 1;
    do {
        i++;
        fCurVolume = fNewVolume;
        fPower = ((float)(i * i)) / ((float)-2.0 * fSigma * fSigma);
        fNewVolume = fCurVolume + 2 * expf(fPower);
    } while (fNewVolume - fCurVolume > 0.00001f);
    for (i = 1; fValue <= fCurVolume * (1.0f - fMinValue); i++) {
        fPower = ((float)(i * i)) / ((float)-2.0 * fSigma * fSigma);
        fValue += 2 * expf(fPower);
    }
    i--;
    return 2 * i + 1;
}

extern "C" int s3d_gaussian_taps(float sigma, float *taps, int cap)
{
    const double PI_ = 3.1415926535897932384626433832795;
    int n = filter_size(sigma, 0.01f);
    if (n > cap) return -n;
    if (sigma > 0.0f) {
        float fMeanCol = (float)(n / 2);
        float fSigmaColSqr = sigma * sigma;
        float fScale = (float)(1.0 / (sigma * sqrt(2.0 * PI_)));
        for (int j = 0; j < n; j++) {
            float fColPos = ((float)j - fMeanCol);
            float fPower = ((fColPos * fColPos) / fSigmaColSqr) / (float)(-2.0f);
            taps[j] = fScale * expf(fPower);
        }
    } else {
        taps[0] = 1;
    }
    float fSum = 0;
    for (int c = 0; c < n; c++) fSum += taps[c];
    for (int c = 0; c < n; c++) taps[c] /= fSum;
    return n;
}

// BRIEF pair table, method 2 (reference MultiScale.cpp:805-807), x,y,z triples.
static const unsigned char kBriefA[192] = { 5,4,4,4,4,2,6,5,5,4,4,4,3,8,5,5,6,3,5,5,5,5,6,5,4,6,6,6,3,4,4,4,5,3,4,5,4,5,5,4,2,7,7,5,3,5,4,5,3,5,7,3,5,5,2,3,5,5,6,6,4,6,5,4,4,6,5,3,5,6,4,3,6,4,4,5,3,3,3,6,6,5,2,4,4,6,3,6,3,2,3,5,4,5,3,4,3,6,5,4,3,6,4,5,2,4,3,7,2,3,6,5,2,6,3,3,5,6,3,6,3,5,3,6,5,7,4,2,5,5,5,2,5,7,4,2,5,3,4,3,3,7,4,4,7,6,4,4,2,8,7,6,5,4,7,3,6,6,5,2,4,5,3,2,5,5,1,6,3,6,3,6,2,5,4,4,7,2,6,3,2,2,4,3,3,2,3,4,2,5,6,7 };
static const unsigned char kBriefB[192] = { 6,5,3,4,5,3,7,4,6,4,3,2,4,7,5,3,5,1,5,4,7,6,8,4,4,5,6,5,2,5,4,6,4,0,4,3,3,4,4,2,1,7,8,6,4,4,1,6,1,3,7,2,3,3,1,3,6,1,6,6,4,7,6,4,3,5,4,2,3,6,4,5,6,3,3,5,1,3,1,6,7,4,1,4,3,5,2,4,2,1,2,5,4,5,2,3,3,3,3,4,2,6,3,4,3,3,3,6,1,2,5,4,2,4,1,4,6,7,3,6,2,4,3,6,5,6,4,0,6,6,5,1,4,7,2,1,5,3,4,2,2,7,3,3,6,4,2,4,1,9,7,7,5,2,7,1,7,5,5,1,5,4,1,3,3,4,0,5,1,6,3,5,3,2,3,3,7,2,5,1,1,0,4,1,3,1,0,3,1,6,5,9 };

static void build_tables(KpTables &t)
{
    memset(&t, 0, sizeof(t));
    // sphere voxels in raster order (reference MultiScale.cpp:2575-2583)
    float fRadiusSqr = (float)((PD / 2) * (PD / 2));
    int n = 0;
    for (int z = 0; z < PD; z++)
        for (int y = 0; y < PD; y++)
            for (int x = 0; x < PD; x++) {
                float dz = (float)(z - PD / 2), dy = (float)(y - PD / 2), dx = (float)(x - PD / 2);
                if (dz * dz + dy * dy + dx * dx < fRadiusSqr) t.sphere[n++] = (unsigned short)((z * PD + y) * PD + x);
            }
    t.n_sphere = n;
    t.n_hist_taps = s3d_gaussian_taps(0.5f, t.hist_taps, 9);    // fBlurGradOriHist
    t.n_brief_taps = s3d_gaussian_taps(0.95f, t.brief_taps, 9); // gb3d_blur3d(..., 0.95, 0.01, ...)
    // spatial-bin coordinate of each patch index (reference MultiScale.cpp:627-660) and its lower-bin
    // weight (_fioDetermineInterpCoord with dim 2, reference FeatureIO.cpp:757-782)
    float fBinSize = PD / (float)2;
    for (int q = 0; q < PD; q++) {
        float c = (int)(q / fBinSize) + 0.5f;
        if ((int)((q + 0) / fBinSize) != (int)((q + 1) / fBinSize)) {
            float fP0 = ((q + 0) / fBinSize);
            float fP1 = ((q + 1) / fBinSize);
            c = (fP0 + fP1) / 2.0f;
        }
        float w;
        if (c < 0.5f) w = 1.0f;
        else if (c >= 2.0f - 0.5f) w = 0.0f;
        else { float mh = c - 0.5f; int i = (int)floorf(mh); w = 1.0f - (mh - (float)i); }
        t.desc_w[q] = w;
    }
    for (int i = 0; i < 64; i++) {
        int ax = kBriefA[3 * i], ay = kBriefA[3 * i + 1], az = kBriefA[3 * i + 2];
        int bx = kBriefB[3 * i], by = kBriefB[3 * i + 1], bz = kBriefB[3 * i + 2];
        t.brief_a[i] = (unsigned short)(ax + ay * PD + az * PD * PD);
        t.brief_b[i] = (unsigned short)(bx + by * PD + bz * PD * PD);
        float fdx = (float)(ax - bx), fdy = (float)(ay - by), fdz = (float)(az - bz);
        t.brief_dist[i] = (float)(int)sqrtf(fdx * fdx + fdy * fdy + fdz * fdz);   // euclidean_distance_3d :1051-1056
    }
}

// One shared-memory carve-out for every kernel of the pipeline.  Kernels that prefer different L1 / shared
// splits cannot be resident on an SM together: the SM has to drain before it is reconfigured, so the
// graph's concurrent branches (and the contexts of a batch) serialise behind each other.  With one carve-out
// the shared-memory-heavy kernels (x+y tiles, orientation histograms) and the cache-reliant ones (detection,
// z march, refinement) overlap.  S3D_CARVEOUT=<percent of the maximum shared memory> overrides the default.
// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct Vol {
    float *p = nullptr;
    int X = 0, Y = 0, Z = 0, pitch = 0;
    size_t elems() const { return (size_t)pitch * Y * Z; }
};

struct Plan {
    int X = 0, Y = 0, Z = 0, double_mode = 0;      // input dims and pre-step
    int input_is_g0 = 0, octave_base = 0, max_octaves = 0, slab = 0, z_off = 0, z_global = 0, own_z0 = 0, own_z1 = 0, pre_step_done = 0;
    int kp_cap = 0, row_cap = 0, cand_cap = 0, keep_patches = 0;
    int n_oct = 0;
    Vol stage;                   // dense copy of the input (pitch == X)
    void *raw_stage = nullptr;   // typed input as it crossed PCIe (allocated on first use, raw_bytes large)
    size_t raw_bytes = 0;
    Vol img0;                    // pre-stepped, pitched input of the pyramid
    std::vector<Vol> g, d;       // [oct*6 + level], [oct*5 + level]
    float *tmp1 = nullptr;            // scratch of the initial blur (octave-0 size)
    float *oct_tmp[kMaxOct] = {};     // per-octave blur scratch (octaves run on concurrent branches)
    s3d_cand *cand_raw = nullptr;     // [list][cand_cap], list = (oct*3 + (c-1))*2 + is_max (atomic order)
    s3d_cand *cand_sorted = nullptr;  // large volumes: the lists grouped by z plane (cand_bucket_kernel), else null
    int *plane_off = nullptr;         // [list][plane_stride] plane offsets of the grouped lists
    int plane_stride = 0;
    unsigned int *face[kMaxOct * 3] = {};   // per (octave, centre level): voxels passing the face test
    int face_cap[kMaxOct * 3] = {};
    int *face_counts = nullptr;
    s3d_keypoint *kp_stage = nullptr; // [list][cand_cap] refined candidates at their raster rank
    unsigned char *stage_flags = nullptr;
    int *counts = nullptr;       // [n_lists] candidate counts, then kp_count, n_features, err
    int n_lists = 0;
    s3d_keypoint *kps = nullptr;
    int *nrows = nullptr, *row_off = nullptr;
    float *kp_eigs = nullptr, *kp_ori0 = nullptr, *kp_rots = nullptr, *kp_patch0 = nullptr, *kp_p1 = nullptr, *kp_fmat = nullptr;
    int *kp_nprim = nullptr, *kp_nsec = nullptr, *work_a = nullptr, *work_b = nullptr, *row_map = nullptr;
    s3d_feature *feats = nullptr;
    float *dbg_patches = nullptr, *dbg_prerank = nullptr;
    PyramidDesc pyr;
    float init_taps[kMaxTaps]; int n_init_taps = 0;
    float lvl_taps[5][kMaxTaps]; int n_lvl_taps[5];
    cudaGraphExec_t graph = nullptr;
    int graph_descriptor = -1; float graph_eig = 0;
    std::vector<void *> allocs;
    unsigned long long last_use = 0;
};

// Experiment / profiling knobs.  Every S3D_* environment variable the library looks at is listed here and is read
// exactly once per context (tuning_from_env, called by ctx_create); the defaults are what the tests and bench.py run.
struct Tuning {
    bool use_graph = true;       // S3D_NO_GRAPH=1: direct launches instead of one CUDA graph per (shape, options)
    bool serial = false;         // S3D_SERIAL=1: no octave / detection branches, every kernel on the main stream
    bool timing = false;         // S3D_STAGE_TIMING=1: no graph, events at stage boundaries of the main stream
    bool stamps = false;         // S3D_STAMPS=1: %globaltimer stamps around the graph (s3d_debug_stamps)
    bool tiny = true;            // S3D_TINY=0: the last octaves as ordinary blur / subsample launches instead of tiny_octaves_kernel
    int prof_sleep_us = 0;       // S3D_PROF_SLEEP_US (with S3D_PROF_SKIP & 1): a one-thread kernel of that duration replaces the keypoint tail
    int prof_skip = 0;           // S3D_PROF_SKIP (profiling only, results invalid): 1 no keypoint tail, 2 no detection/refinement, 4 no describe, 8 no orient_b, 16 no level-5 blur
    int f4_max_r = 6;            // S3D_F4_MAXR: one-kernel level (s3d_blur4.cuh) for radii up to this; wider levels use x+y / z kernels (s3d_blur2.cuh)
    long long f4_wide_min_voxels = 16000000;   // S3D_F4_WIDE_MIN_VOXELS: inside the pipeline radii 5 and 6 take the one-kernel level from this size on.  Their
                                 // CTAs (setmaxnreg 64/192, 8 warps, one per SM) keep the detection branch of the SAME volume from running beside
                                 // them: alone on the GPU an MNI volume takes 865 us with them and 726 us without (the level itself is faster:
                                 // 43.8 against 51.2 us at 11 taps).  Batch contexts set it to f4_min_voxels: other volumes fill the SMs there.
    bool f4_wide_forced = false;
    long long f4_min_voxels = 2000000;   // S3D_F4_MIN_VOXELS: smaller volumes (octaves >= 1 at MNI size) use the x+y / z kernels: the one-kernel level walks its
                                 // z segment plane by plane, a latency chain that a small volume cannot hide behind other CTAs (measured: 11-25 us against 4 + 6 us)
    long long bucket_min_voxels = 16000000;   // S3D_BUCKET_MIN_VOXELS: from this pyramid size on the candidate lists are grouped by plane before they are ranked
    long long detect2_min_voxels = 1500000;   // S3D_DETECT2_MIN_VOXELS: smaller volumes run the single-kernel extremum test instead of face test + full test
    int f4_ty = 16;              // S3D_F4_TY=16|32: tile rows of the one-kernel level (two / one resident CTAs per SM)
    bool f4_ctas_forced = false;
    int f4_ctas = 0;             // S3D_F4_CTAS: CTAs the one-kernel level aims for (0 = resident CTAs per SM x SMs)
    int xy2_ctas = 2;            // S3D_XY2_CTAS_PER_SM: persistent x+y CTAs per SM (contexts of an s3d_batch use 1)
    bool xy2_ctas_forced = false, xy2_threads_forced = false;
    XY2Tune xy2;                 // S3D_XY2_SMEM_KB, S3D_XY2_TX / S3D_XY2_TY, S3D_XY2_KY, S3D_XY2_THREADS: tile choice of the x+y kernel
    int z2_vec = 0;              // S3D_Z2_VEC=2|4: columns per thread of the z march (0 = by radius)
    int march_target = 0;        // S3D_MARCH_TARGET: threads wanted in flight in the z march (0 = 256 per SM)
    int detect_ctas = 0;         // S3D_DETECT_CTAS: resident detect_face blocks per SM (0 = whole grid; contexts of an s3d_batch: 4)
    bool detect_ctas_forced = false;
    bool tail_forced = false;
    int tail_a = 3, tail_b = 6, tail_d = 10;   // S3D_TAIL_BLOCKS=a,b,d: blocks per SM of orient_a / orient_b / describe
    int desc_threads = 128;      // S3D_DESC_THREADS=64|128: threads per describe block
    int max_plans = 6;           // S3D_PLAN_CACHE: resident plans (shapes) per context
};

static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

static Tuning tuning_from_env()
{
    Tuning t;
    t.use_graph = env_int("S3D_NO_GRAPH", 0) != 1;
    t.serial = env_int("S3D_SERIAL", 0) == 1;
    t.timing = env_int("S3D_STAGE_TIMING", 0) == 1;
    if (t.timing) t.use_graph = false;        // event nodes inside graphs carry no timestamps
    t.stamps = env_int("S3D_STAMPS", 0) == 1;
    t.prof_skip = env_int("S3D_PROF_SKIP", 0);
    t.tiny = env_int("S3D_TINY", 1) != 0;
    t.prof_sleep_us = env_int("S3D_PROF_SLEEP_US", 0);
    int v = env_int("S3D_F4_MAXR", t.f4_max_r);
    if (v >= 0) t.f4_max_r = v < kF4MaxR ? v : kF4MaxR;
    if (env_int("S3D_F4_TY", 16) == 32) t.f4_ty = 32;
    { const char *mv = getenv("S3D_F4_MIN_VOXELS"); if (mv && mv[0]) t.f4_min_voxels = atoll(mv); }
    { const char *mv = getenv("S3D_F4_WIDE_MIN_VOXELS"); if (mv && mv[0]) { t.f4_wide_min_voxels = atoll(mv); t.f4_wide_forced = true; } }
    { const char *mv = getenv("S3D_BUCKET_MIN_VOXELS"); if (mv && mv[0]) t.bucket_min_voxels = atoll(mv); }
    { const char *mv = getenv("S3D_DETECT2_MIN_VOXELS"); if (mv && mv[0]) t.detect2_min_voxels = atoll(mv); }
    t.f4_ctas = env_int("S3D_F4_CTAS", 0);
    t.f4_ctas_forced = t.f4_ctas > 0;
    v = env_int("S3D_XY2_CTAS_PER_SM", 0);
    if (v >= 1 && v <= 4) { t.xy2_ctas = v; t.xy2_ctas_forced = true; }
    v = env_int("S3D_XY2_SMEM_KB", t.xy2.max_kb);
    if (v >= 16 && v <= 112) t.xy2.max_kb = v;
    t.xy2.force_tx = env_int("S3D_XY2_TX", 0);
    t.xy2.force_ty = env_int("S3D_XY2_TY", 0);
    v = env_int("S3D_XY2_KY", 0);
    if (v == 8 || v == 16) t.xy2.ky = v;
    { const int v2 = env_int("S3D_XY2_THREADS", 0); if (v2 == 128 || v2 == 256) { t.xy2.threads = v2; t.xy2_threads_forced = true; } }
    v = env_int("S3D_Z2_VEC", 0);
    if (v == 2 || v == 4) t.z2_vec = v;
    t.march_target = env_int("S3D_MARCH_TARGET", 0);
    v = env_int("S3D_DETECT_CTAS", -1);
    if (v >= 0) { t.detect_ctas = v; t.detect_ctas_forced = true; }
    {
        const char *tb = getenv("S3D_TAIL_BLOCKS");
        int a = 0, b = 0, d = 0;
        if (tb && sscanf(tb, "%d,%d,%d", &a, &b, &d) == 3 && a > 0 && b > 0 && d > 0) { t.tail_a = a; t.tail_b = b; t.tail_d = d; t.tail_forced = true; }
    }
    v = env_int("S3D_DESC_THREADS", 0);
    if (v == 64 || v == 128) t.desc_threads = v;
    v = env_int("S3D_PLAN_CACHE", 0);
    if (v >= 1) t.max_plans = v;
    return t;
}

struct s3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    Plan *plan = nullptr;               // active plan (one of `plans`)
    std::vector<Plan *> plans;          // resident plans, least recently used evicted beyond tune.max_plans: a caller that
    unsigned long long use_clock = 0;   // alternates between a few shapes (slab mode: one per octave) keeps its buffers and graphs
    int launches = 0;
    int last_launches = 0;
    int sm_count = 148;
    bool has_result = false;
    int *h_counts = nullptr;     // pinned: kp_count, n_features, err
    // Speculative result copy: when rows will be fetched by the host (s3d_extract, s3d_batch_extract*), the counts
    // and the first spec_rows feature rows are copied to pinned memory by the stream right behind the graph, so
    // that s3d_fetch_features is one stream synchronisation and a memcpy instead of two round trips and a
    // pageable-memory DMA (measured: ~95 us per volume of the host's time in batch mode).
    bool spec_enable = false;
    s3d_feature *h_rows = nullptr;   // pinned
    int h_rows_cap = 0;              // rows h_rows can hold
    int spec_rows = 0;               // rows copied behind the current result (0 = none)
    int spec_guess = 2048;           // rows to copy next time (tracks 1.25 x the last count)
    bool spec_valid = false;         // h_counts / h_rows belong to the current result
    cudaStream_t cur = nullptr;  // stream the stage launchers enqueue on (main stream or an octave branch)
    bool in_pipeline = false;    // the stage launchers are called by enqueue_pipeline (concurrent branches), not through the stage-level API
    Tuning tune;                 // S3D_* environment knobs, read once in ctx_create
    unsigned long long *d_stamps = nullptr;   // S3D_STAMPS=1
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    cudaStream_t side[kMaxOct] = {}, det[kMaxOct] = {};     // octave branch, detection/refinement branch
    cudaEvent_t ev_fork[kMaxOct] = {}, ev_done[kMaxOct] = {}, ev_lvl[kMaxOct][4] = {};
};

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            char b_[512];                                                                            \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                           \
            return S3D_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

static s3d_status fail(s3d_ctx *ctx, s3d_status st, const char *msg)
{
    if (ctx) ctx->err = msg;
    return st;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" int s3d_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static s3d_status ctx_create(int device, void *stream, bool borrow, s3d_ctx **out)
{
    if (!out) return S3D_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        fprintf(stderr, "s3d: no usable CUDA device (%s); this engine has no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return S3D_ERR_CUDA;
    }
    if (device < 0 || device >= n) return S3D_ERR_INVALID;
    s3d_ctx *ctx = new s3d_ctx();
    ctx->device = device;
    *out = ctx;   // handed back even on failure so the caller can read the error text
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (borrow) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    KpTables t;
    build_tables(t);
    if (t.n_sphere != kSph) return fail(ctx, S3D_ERR_INVALID, "sphere table size");   // the keypoint kernels size their shared memory for it
    CK(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
    CK(cudaFuncSetAttribute(tiny_octaves_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTinySmem));
    CK(cudaFuncSetAttribute(orient_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PatchSmem)));
    CK(cudaFuncSetAttribute(orient_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HistSmem)));
    CK(cudaFuncSetAttribute(orient_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HistSmem)));
    CK(cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DescribeSmem)));
    CK(cudaMallocHost((void **)&ctx->h_counts, 4 * sizeof(int)));
    ctx->cur = ctx->stream;
    // The branches of the smaller octaves and the detection/refinement branches are chains of short,
    // latency-bound kernels: give them priority over octave 0's long blur kernels so they are not queued
    // behind them (stream priorities are kept by the captured graph nodes).
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int o = 1; o < kMaxOct; o++) {
        CK(cudaStreamCreateWithPriority(&ctx->side[o], cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&ctx->ev_done[o], cudaEventDisableTiming));
    }
    for (int o = 0; o < kMaxOct; o++) {
        CK(cudaEventCreateWithFlags(&ctx->ev_fork[o], cudaEventDisableTiming));
        CK(cudaStreamCreateWithPriority(&ctx->det[o], cudaStreamNonBlocking, prio_hi));
        if (!ctx->ev_done[o]) CK(cudaEventCreateWithFlags(&ctx->ev_done[o], cudaEventDisableTiming));
        for (int k = 0; k < 4; k++) CK(cudaEventCreateWithFlags(&ctx->ev_lvl[o][k], cudaEventDisableTiming));
    }
    ctx->tune = tuning_from_env();
    CK(init_blur2_attrs());
    CK(init_blur4_attrs());
    if (ctx->tune.stamps) CK(cudaMalloc((void **)&ctx->d_stamps, 4 * sizeof(unsigned long long)));
    return S3D_OK;
}

extern "C" s3d_status s3d_ctx_create(int device, s3d_ctx **ctx) { return ctx_create(device, nullptr, false, ctx); }
extern "C" s3d_status s3d_ctx_create_on_stream(int device, void *stream, s3d_ctx **ctx) { return ctx_create(device, stream, true, ctx); }

static void plan_destroy(s3d_ctx *ctx, Plan *p)
{
    if (!p) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (p->graph) cudaGraphExecDestroy(p->graph);
    for (void *a : p->allocs) cudaFree(a);
    for (size_t i = 0; i < ctx->plans.size(); i++)
        if (ctx->plans[i] == p) { ctx->plans.erase(ctx->plans.begin() + i); break; }
    if (ctx->plan == p) { ctx->plan = nullptr; ctx->has_result = false; }
    delete p;
}

static void plan_free(s3d_ctx *ctx)     // every resident plan
{
    while (!ctx->plans.empty()) plan_destroy(ctx, ctx->plans.back());
    ctx->plan = nullptr;
    ctx->has_result = false;
}

extern "C" void s3d_ctx_destroy(s3d_ctx *ctx)
{
    if (!ctx) return;
    plan_free(ctx);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->h_rows) cudaFreeHost(ctx->h_rows);
    for (int o = 0; o < kMaxOct; o++) {
        if (ctx->side[o]) cudaStreamDestroy(ctx->side[o]);
        if (ctx->det[o]) cudaStreamDestroy(ctx->det[o]);
        for (int k = 0; k < 4; k++) if (ctx->ev_lvl[o][k]) cudaEventDestroy(ctx->ev_lvl[o][k]);
        if (ctx->ev_fork[o]) cudaEventDestroy(ctx->ev_fork[o]);
        if (ctx->ev_done[o]) cudaEventDestroy(ctx->ev_done[o]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *s3d_last_error(const s3d_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void *s3d_stream(s3d_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" s3d_status s3d_sync(s3d_ctx *ctx)
{
    if (!ctx) return S3D_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return S3D_OK;
}
extern "C" void s3d_free(void *p) { free(p); }
extern "C" int s3d_last_launch_count(s3d_ctx *ctx) { return ctx ? ctx->last_launches : 0; }

#ifdef S3D_PHASE_TIMERS
// profiling builds only: read (and clear) the per-phase cycle counters of the keypoint kernels
extern "C" int s3d_debug_phase_cycles(unsigned long long *out32)
{
    unsigned long long z[32] = { 0 };
    if (cudaMemcpyFromSymbol(out32, g_phase, sizeof(z)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(g_phase, z, sizeof(z)) != cudaSuccess) return -1;
    return 0;
}
#endif

// S3D_STAMPS=1 (profiling): %globaltimer at four points of an extraction -- before / after the input
// re-pitch (outside the graph), first and last node of the graph -- to see launch gaps without a profiler
// profiling only (S3D_PROF_SLEEP_US): a one-thread kernel that lasts `us` microseconds and uses no shared resource --
// put in place of the keypoint tail it tells whether the tail costs the batch its latency or its resources
__global__ void sleep_kernel(int us)
{
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)us * 1900) __nanosleep(200);
}
__global__ void stamp_kernel(unsigned long long *slot)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
extern "C" int s3d_debug_stamps(s3d_ctx *ctx, unsigned long long *out4)
{
    if (!ctx || !ctx->d_stamps) return -1;
    if (cudaMemcpy(out4, ctx->d_stamps, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// stage launchers
// ---------------------------------------------------------------------------------------------------
static bool fast_layout(const void *a, const void *b, const void *c, int pitch)
{
    return pitch % 8 == 0 && ((uintptr_t)a % 32) == 0 && ((uintptr_t)b % 32) == 0 && ((uintptr_t)c % 32) == 0;
}

// One blur level on the fast path (pitched layout, symmetric taps of radius R <= kMaxFastR): the one-kernel level
// for radii up to tune.f4_max_r, else the x+y kernel followed by the z march.  Returns false when nothing was
// launched (the driver could not encode a tensor map): the caller falls back to the generic kernels.
template <int R>
static bool launch_blur_fast(s3d_ctx *ctx, const float *in, float *tmp, float *out, int X, int Y, int Z, int pitch,
                             const float *taps, float *dog, cudaError_t *err)
{
    const long long plane = (long long)pitch * Y;
    if (plane * Z >= (1ll << 31)) return false;
    const Tuning &tn = ctx->tune;
    if constexpr (R <= kF4MaxR) {
        const long long min_vox = (R >= 5 && ctx->in_pipeline && tn.f4_wide_min_voxels > tn.f4_min_voxels) ? tn.f4_wide_min_voxels : tn.f4_min_voxels;
        if (R <= tn.f4_max_r && plane * Z >= min_vox &&
            (tn.f4_ty == 32 ? launch_blur_f4<R, 32>(ctx->cur, in, out, dog, X, Y, Z, pitch, taps, ctx->sm_count, tn.f4_ctas, err)
                            : launch_blur_f4<R, 16>(ctx->cur, in, out, dog, X, Y, Z, pitch, taps, ctx->sm_count, tn.f4_ctas, err))) {
            ctx->launches += 1;
            return true;
        }
    }
    if (!launch_blur_xy2<R>(ctx->cur, in, tmp, X, Y, Z, pitch, taps, ctx->sm_count, tn.xy2_ctas, tn.xy2, err)) return false;
    ctx->launches += 1;
    if (*err != cudaSuccess) return true;
    const int target = tn.march_target > 0 ? tn.march_target : ctx->sm_count * 256;
    *err = launch_blur_z2<R>(ctx->cur, tmp, out, in, dog, Y, Z, pitch, taps, target, tn.z2_vec);
    ctx->launches += 1;
    return true;
}

static void launch_blur_generic(s3d_ctx *ctx, const float *in, float *tmp, float *out, int X, int Y, int Z, int pitch,
                                const float *taps, int ntaps, float *dog)
{
    TapsAny t;
    memset(&t, 0, sizeof(t));
    t.n = ntaps;
    for (int j = 0; j < ntaps; j++) t.w[j] = taps[j];
    long long plane = (long long)pitch * Y;
    long long n = plane * Z;
    blur_x_generic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->cur>>>(in, out, pitch, X, n, t);
    {
        long long cols = (long long)pitch * Z;
        dim3 grid((unsigned)((cols + 127) / 128), (unsigned)(Y < 32 ? Y : 32));
        blur_march_generic_kernel<false><<<grid, 128, 0, ctx->cur>>>(out, tmp, nullptr, nullptr, cols, pitch, plane, pitch, Y, t);
    }
    {
        dim3 grid((unsigned)((plane + 127) / 128), (unsigned)(Z < 32 ? Z : 32));
        if (dog) blur_march_generic_kernel<true><<<grid, 128, 0, ctx->cur>>>(tmp, out, in, dog, plane, plane, 0, plane, Z, t);
        else blur_march_generic_kernel<false><<<grid, 128, 0, ctx->cur>>>(tmp, out, nullptr, nullptr, plane, plane, 0, plane, Z, t);
    }
    ctx->launches += 3;
}

static s3d_status blur3d(s3d_ctx *ctx, const float *in, float *tmp, float *out, int X, int Y, int Z, int pitch,
                         const float *taps, int ntaps, float *dog)
{
    if (!in || !tmp || !out || !taps || X <= 0 || Y <= 0 || Z <= 0 || pitch < X || ntaps < 1 || (ntaps & 1) == 0 || ntaps > kMaxTaps)
        return fail(ctx, S3D_ERR_INVALID, "s3d_blur3d: bad argument");
    if (in == out || in == tmp || tmp == out) return fail(ctx, S3D_ERR_INVALID, "s3d_blur3d: in/tmp/out must be distinct");
    int R = ntaps / 2;
    bool fast = fast_layout(in, tmp, out, pitch) && (!dog || ((uintptr_t)dog % 32) == 0) && R >= 1 && R <= kMaxFastR &&
                taps_symmetric(taps, ntaps);
    if (fast) {
        cudaError_t e = cudaSuccess;
        bool launched = false;
        switch (R) {
        case 1: launched = launch_blur_fast<1>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 2: launched = launch_blur_fast<2>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 3: launched = launch_blur_fast<3>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 4: launched = launch_blur_fast<4>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 5: launched = launch_blur_fast<5>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 6: launched = launch_blur_fast<6>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        case 7: launched = launch_blur_fast<7>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        default: launched = launch_blur_fast<8>(ctx, in, tmp, out, X, Y, Z, pitch, taps, dog, &e); break;
        }
        if (launched) {
            if (e != cudaSuccess) {
                char b_[256];
                snprintf(b_, sizeof(b_), "blur level (radius %d) launch failed: %s", R, cudaGetErrorString(e));
                ctx->err = b_;
                return S3D_ERR_CUDA;
            }
            return S3D_OK;
        }
    }
    launch_blur_generic(ctx, in, tmp, out, X, Y, Z, pitch, taps, ntaps, dog);
    CK(cudaGetLastError());
    return S3D_OK;
}

extern "C" s3d_status s3d_blur3d(s3d_ctx *ctx, const float *d_in, float *d_tmp, float *d_out,
                                 int X, int Y, int Z, int pitch, const float *h_taps, int ntaps, float *d_dog)
{
    if (!ctx) return S3D_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    return blur3d(ctx, d_in, d_tmp, d_out, X, Y, Z, pitch, h_taps, ntaps, d_dog);
}

extern "C" s3d_status s3d_dog(s3d_ctx *ctx, const float *d_a, const float *d_b, float *d_out, int X, int Y, int Z, int pitch)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (!d_a || !d_b || !d_out || X <= 0 || Y <= 0 || Z <= 0 || pitch < X) return fail(ctx, S3D_ERR_INVALID, "s3d_dog: bad argument");
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    long long n = (long long)pitch * Y * Z;
    dog_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->cur>>>(d_a, d_b, d_out, n);
    ctx->launches++;
    CK(cudaGetLastError());
    return S3D_OK;
}

static s3d_status resize_launch(s3d_ctx *ctx, int kind, const float *in, int X, int Y, int Z, int pitch, float *out, int opitch)
{
    if (!in || !out || X <= 1 || Y <= 1 || Z <= 1 || pitch < X) return fail(ctx, S3D_ERR_INVALID, "resize: bad argument");
    int ox = kind == 2 ? 2 * X : X / 2, oy = kind == 2 ? 2 * Y : Y / 2, oz = kind == 2 ? 2 * Z : Z / 2;
    if (ox < 1 || oy < 1 || oz < 1 || opitch < ox) return fail(ctx, S3D_ERR_INVALID, "resize: bad output shape");
    dim3 block(32, 8), grid((opitch + 31) / 32, (oy + 7) / 8, oz);
    if (kind == 0) subsample_kernel<<<grid, block, 0, ctx->cur>>>(in, X, Y, Z, pitch, out, ox, oy, oz, opitch);
    else if (kind == 1) halve_kernel<<<grid, block, 0, ctx->cur>>>(in, X, Y, Z, pitch, out, ox, oy, oz, opitch);
    else double_kernel<<<grid, block, 0, ctx->cur>>>(in, X, Y, Z, pitch, out, opitch);
    ctx->launches++;
    CK(cudaGetLastError());
    return S3D_OK;
}

// Isotropic resampling on the device (section 8(f) N1; reference featExtract.cpp:118-204, a host triple loop there)
extern "C" s3d_status s3d_resample_iso(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch,
                                       float *d_out, int nX, int nY, int nZ, int out_pitch, float rf_x, float rf_y, float rf_z)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (!d_in || !d_out || X < 2 || Y < 2 || Z < 2 || pitch < X || nX < 1 || nY < 1 || nZ < 1 || out_pitch < nX)
        return fail(ctx, S3D_ERR_INVALID, "s3d_resample_iso: bad argument");
    CK(cudaSetDevice(ctx->device));
    dim3 block(32, 8), grid((out_pitch + 31) / 32, (nY + 7) / 8, nZ);
    resample_iso_kernel<<<grid, block, 0, ctx->stream>>>(d_in, X, Y, Z, pitch, d_out, nX, nY, nZ, out_pitch, rf_x, rf_y, rf_z);
    ctx->launches++;
    CK(cudaGetLastError());
    return S3D_OK;
}

// host-array form of s3d_resample_iso: upload, resample, download, synchronise (what the CLI's loader calls)
extern "C" s3d_status s3d_resample_iso_host(s3d_ctx *ctx, const float *h_in, int X, int Y, int Z,
                                            float *h_out, int nX, int nY, int nZ, float rf_x, float rf_y, float rf_z)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (!h_in || !h_out) return fail(ctx, S3D_ERR_INVALID, "s3d_resample_iso_host: null array");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t n_in = (size_t)X * Y * Z, n_out = (size_t)nX * nY * nZ;
    float *d_in = nullptr, *d_out = nullptr;
    CK(cudaMallocAsync((void **)&d_in, n_in * sizeof(float), st));
    CK(cudaMallocAsync((void **)&d_out, n_out * sizeof(float), st));
    CK(cudaMemcpyAsync(d_in, h_in, n_in * sizeof(float), cudaMemcpyHostToDevice, st));
    s3d_status s = s3d_resample_iso(ctx, d_in, X, Y, Z, X, d_out, nX, nY, nZ, nX, rf_x, rf_y, rf_z);
    if (s == S3D_OK) CK(cudaMemcpyAsync(h_out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, st));
    cudaFreeAsync(d_in, st);
    cudaFreeAsync(d_out, st);
    CK(cudaStreamSynchronize(st));
    return s;
}

extern "C" s3d_status s3d_subsample2(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch, float *d_out, int out_pitch)
{
    if (!ctx) return S3D_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    return resize_launch(ctx, 0, d_in, X, Y, Z, pitch, d_out, out_pitch);
}
extern "C" s3d_status s3d_halve_size(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch, float *d_out, int out_pitch)
{
    if (!ctx) return S3D_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    return resize_launch(ctx, 1, d_in, X, Y, Z, pitch, d_out, out_pitch);
}
extern "C" s3d_status s3d_double_size(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch, float *d_out, int out_pitch)
{
    if (!ctx) return S3D_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    return resize_launch(ctx, 2, d_in, X, Y, Z, pitch, d_out, out_pitch);
}

// detection into raw (atomic-order) lists
static s3d_status detect_raw_launch(s3d_ctx *ctx, const float *finer, const float *centre, int X, int Y, int Z, int pitch,
                                    s3d_cand *raw_min, int *n_min, s3d_cand *raw_max, int *n_max, int cap,
                                    int own0 = 0, int own1 = 0x7fffffff)
{
    if (X < 3 || Y < 3 || Z < 3) return S3D_OK;   // no interior voxel
    dim3 block(32, 8), grid((X - 2 + 31) / 32, (Y - 2 + 7) / 8, (Z - 2 + kDetectZ - 1) / kDetectZ);
    CandList lmin{ raw_min, n_min }, lmax{ raw_max, n_max };
    detect_kernel<<<grid, block, 0, ctx->cur>>>(finer, centre, X, Y, Z, pitch, lmin, lmax, cap, own0, own1);
    ctx->launches += 1;
    CK(cudaGetLastError());
    return S3D_OK;
}

// two-pass detection used by the pipeline: face test over the volume, full test on the survivors
static s3d_status detect_two_pass(s3d_ctx *ctx, const float *finer, const float *centre, int X, int Y, int Z, int pitch,
                                  unsigned int *face, int *face_count, int face_cap,
                                  s3d_cand *raw_min, int *n_min, s3d_cand *raw_max, int *n_max, int cap, int *err,
                                  int own0, int own1)
{
    if (X < 3 || Y < 3 || Z < 3) return S3D_OK;
    // 32-bit voxel offsets in the face list; small volumes: one launch instead of two latency-bound ones
    if ((long long)pitch * Y * Z >= (1ll << 32) || (long long)pitch * Y * Z < ctx->tune.detect2_min_voxels)
        return detect_raw_launch(ctx, finer, centre, X, Y, Z, pitch, raw_min, n_min, raw_max, n_max, cap, own0, own1);
    const int n_zblocks = (Z - 2 + kDetectZ - 1) / kDetectZ;
    dim3 block(32, 8), grid((X - 2 + 31) / 32, (Y - 2 + 7) / 8, n_zblocks);
    if (ctx->tune.detect_ctas > 0) {        // batch contexts: cap the resident blocks (see detect_face_kernel)
        long long per_layer = (long long)grid.x * grid.y;
        long long gz = ((long long)ctx->tune.detect_ctas * ctx->sm_count + per_layer - 1) / per_layer;
        if (gz < 1) gz = 1;
        if (gz < n_zblocks) grid.z = (unsigned)gz;
    }
    detect_face_kernel<<<grid, block, 0, ctx->cur>>>(finer, centre, X, Y, Z, pitch, n_zblocks, face, face_count, face_cap);
    CandList lmin{ raw_min, n_min }, lmax{ raw_max, n_max };
    int blocks = (int)(((long long)X * Y * Z / 64 + 255) / 256);
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    if (blocks < 1) blocks = 1;
    detect_full_kernel<<<blocks, 256, 0, ctx->cur>>>(finer, centre, X, Y, Z, pitch, face, face_count, face_cap, lmin, lmax, cap, err, ERR_CAND_OVERFLOW, own0, own1);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return S3D_OK;
}

// detection + raster ordering (stage-level API)
static s3d_status detect_launch(s3d_ctx *ctx, const float *finer, const float *centre, int X, int Y, int Z, int pitch,
                                s3d_cand *raw_min, s3d_cand *raw_max, s3d_cand *out_min, int *n_min,
                                s3d_cand *out_max, int *n_max, int cap)
{
    if (X < 3 || Y < 3 || Z < 3) return S3D_OK;
    s3d_status st = detect_raw_launch(ctx, finer, centre, X, Y, Z, pitch, raw_min, n_min, raw_max, n_max, cap);
    if (st != S3D_OK) return st;
    int blocks = (cap + 255) / 256;
    order_candidates_kernel<<<blocks, 256, 0, ctx->cur>>>(raw_min, n_min, out_min, X, Y, cap);
    order_candidates_kernel<<<blocks, 256, 0, ctx->cur>>>(raw_max, n_max, out_max, X, Y, cap);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return S3D_OK;
}

extern "C" s3d_status s3d_detect(s3d_ctx *ctx, const float *d_finer, const float *d_centre, int X, int Y, int Z, int pitch,
                                 s3d_cand *d_min, int *d_n_min, s3d_cand *d_max, int *d_n_max, int cap)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (!d_finer || !d_centre || !d_min || !d_max || !d_n_min || !d_n_max || cap <= 0 || pitch < X)
        return fail(ctx, S3D_ERR_INVALID, "s3d_detect: bad argument");
    CK(cudaSetDevice(ctx->device));
    ctx->cur = ctx->stream;
    s3d_cand *raw = nullptr;
    CK(cudaMallocAsync((void **)&raw, sizeof(s3d_cand) * 2 * (size_t)cap, ctx->stream));
    CK(cudaMemsetAsync(d_n_min, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(d_n_max, 0, sizeof(int), ctx->stream));
    s3d_status st = detect_launch(ctx, d_finer, d_centre, X, Y, Z, pitch, raw, raw + cap, d_min, d_n_min, d_max, d_n_max, cap);
    CK(cudaFreeAsync(raw, ctx->stream));
    return st;
}

// ---------------------------------------------------------------------------------------------------
// plan: every buffer of the pyramid for one input shape
// ---------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t plan_alloc(Plan *p, T **ptr, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
    if (e != cudaSuccess) return e;
    p->allocs.push_back(q);
    *ptr = (T *)q;
    return cudaSuccess;
}

static bool plan_matches(const Plan *old, int X, int Y, int Z, const s3d_params *prm, int kp_cap, int row_cap)
{
    return old && old->X == X && old->Y == Y && old->Z == Z && old->double_mode == prm->double_mode &&
           old->kp_cap == kp_cap && old->row_cap == row_cap && old->keep_patches == (prm->keep_patches ? 1 : 0) &&
           old->input_is_g0 == prm->input_is_g0 && old->octave_base == prm->octave_base && old->max_octaves == prm->max_octaves &&
           old->slab == prm->slab && old->z_off == prm->z_off && old->z_global == prm->z_global &&
           old->own_z0 == prm->own_z0 && old->own_z1 == prm->own_z1 && old->pre_step_done == prm->pre_step_done;
}

static s3d_status plan_fill(s3d_ctx *ctx, Plan *p, int X, int Y, int Z, const s3d_params *prm, int kp_cap, int row_cap);

// Make the plan for (shape, options) the active one: the active plan, a resident one, or a new one (the least
// recently used resident plan is released first when max_plans are resident).
static s3d_status plan_build(s3d_ctx *ctx, int X, int Y, int Z, const s3d_params *prm)
{
    int kp_cap = prm->max_keypoints > 0 ? prm->max_keypoints : 16384;
    int row_cap = prm->max_features > 0 ? prm->max_features : 8 * kp_cap;
    ctx->use_clock++;
    if (plan_matches(ctx->plan, X, Y, Z, prm, kp_cap, row_cap)) { ctx->plan->last_use = ctx->use_clock; return S3D_OK; }
    for (Plan *q : ctx->plans)
        if (plan_matches(q, X, Y, Z, prm, kp_cap, row_cap)) {
            ctx->plan = q;
            ctx->has_result = false;
            q->last_use = ctx->use_clock;
            return S3D_OK;
        }
    while ((int)ctx->plans.size() >= (ctx->tune.max_plans > 1 ? ctx->tune.max_plans : 1)) {
        Plan *lru = ctx->plans[0];
        for (Plan *q : ctx->plans) if (q->last_use < lru->last_use) lru = q;
        plan_destroy(ctx, lru);
    }
    s3d_status st = S3D_OK;
    for (int attempt = 0; attempt < 2; attempt++) {
        Plan *p = new Plan();
        p->last_use = ctx->use_clock;
        ctx->plans.push_back(p);
        ctx->plan = p;
        ctx->has_result = false;
        st = plan_fill(ctx, p, X, Y, Z, prm, kp_cap, row_cap);
        if (st == S3D_OK) break;
        // never leave a half-built plan behind (its key would match the next call)
        std::string msg = ctx->err;
        plan_destroy(ctx, p);
        ctx->err = msg;
        // out of device memory with other plans resident: release them all and try once more
        if (st != S3D_ERR_NOMEM || ctx->plans.empty() || attempt == 1) break;
        while (!ctx->plans.empty()) plan_destroy(ctx, ctx->plans.back());
    }
    return st;
}

static s3d_status plan_fill(s3d_ctx *ctx, Plan *p, int X, int Y, int Z, const s3d_params *prm, int kp_cap, int row_cap)
{
    p->X = X; p->Y = Y; p->Z = Z; p->double_mode = prm->double_mode;
    p->kp_cap = kp_cap; p->row_cap = row_cap; p->keep_patches = prm->keep_patches ? 1 : 0;
    p->input_is_g0 = prm->input_is_g0; p->octave_base = prm->octave_base; p->max_octaves = prm->max_octaves;
    p->slab = prm->slab; p->z_off = prm->z_off; p->z_global = prm->z_global; p->own_z0 = prm->own_z0; p->own_z1 = prm->own_z1; p->pre_step_done = prm->pre_step_done;

    int X0 = X, Y0 = Y, Z0 = Z;
    if (prm->double_mode == 1) { X0 *= 2; Y0 *= 2; Z0 *= 2; }
    else if (prm->double_mode == -1) { X0 /= 2; Y0 /= 2; Z0 /= 2; }
    if (X0 < 1 || Y0 < 1 || Z0 < 1) return fail(ctx, S3D_ERR_INVALID, "volume too small");

    cudaError_t e;
#define PA(ptr, count)                                                          \
    do {                                                                        \
        e = plan_alloc(p, ptr, (size_t)(count));                                \
        if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e); cudaGetLastError(); return S3D_ERR_NOMEM; } \
    } while (0)

    p->stage.X = X; p->stage.Y = Y; p->stage.Z = Z; p->stage.pitch = X;
    PA(&p->stage.p, p->stage.elems());
    p->img0.X = X0; p->img0.Y = Y0; p->img0.Z = Z0; p->img0.pitch = round_up(X0, 8);
    PA(&p->img0.p, p->img0.elems());
    PA(&p->tmp1, p->img0.elems());

    // octaves (reference MultiScale.cpp:359: stop when any dimension <= 2)
    int ox = X0, oy = Y0, oz = Z0, n_oct = 0;
    memset(&p->pyr, 0, sizeof(p->pyr));
    while (!(ox <= 2 || oy <= 2 || oz <= 2) && n_oct < kMaxOct && (prm->max_octaves <= 0 || n_oct < prm->max_octaves)) {
        OctaveDesc &od = p->pyr.oct[n_oct];
        od.X = ox; od.Y = oy; od.Z = oz; od.pitch = round_up(ox, 8);
        od.z_off = 0; od.Zg = oz; od.own0 = 0; od.own1 = oz;
        if (prm->slab) {      // one octave per call in slab mode (halos are refreshed between octaves)
            od.z_off = prm->z_off; od.Zg = prm->z_global;
            od.own0 = prm->own_z0 - prm->z_off; od.own1 = prm->own_z1 - prm->z_off;
        }
        for (int j = 0; j < 6; j++) {
            Vol v; v.X = ox; v.Y = oy; v.Z = oz; v.pitch = od.pitch;
            PA(&v.p, v.elems());
            p->g.push_back(v);
            od.g[j] = v.p;
        }
        for (int j = 0; j < 5; j++) {
            Vol v; v.X = ox; v.Y = oy; v.Z = oz; v.pitch = od.pitch;
            PA(&v.p, v.elems());
            p->d.push_back(v);
            od.d[j] = v.p;
        }
        PA(&p->oct_tmp[n_oct], (size_t)od.pitch * oy * oz);
        n_oct++;
        ox /= 2; oy /= 2; oz /= 2;
    }
    p->n_oct = n_oct;
    p->pyr.n_oct = n_oct;

    // schedule (reference MultiScale.cpp:288-294, 337-371, 526-527)
    {
        float fInitialImageScale = (prm->double_mode == 1 || prm->pre_step_done == 1) ? 0.5f : 1.0f;
        float fSigmaInit = 0.5f;
        if (fInitialImageScale > 0) fSigmaInit /= fInitialImageScale;
        float fSigma = 1.6f;
        float fSigmaExtra = sqrtf(fSigma * fSigma - fSigmaInit * fSigmaInit);
        p->n_init_taps = s3d_gaussian_taps(fSigmaExtra, p->init_taps, kMaxTaps);
        float fSigmaFactor = (float)pow(2.0, 1.0 / (double)3);
        float sig[6];
        sig[0] = fSigma;
        for (int j = 1; j < 6; j++) {
            float ex = fSigma * sqrtf(fSigmaFactor * fSigmaFactor - 1.0f);
            p->n_lvl_taps[j - 1] = s3d_gaussian_taps(ex, p->lvl_taps[j - 1], kMaxTaps);
            fSigma *= fSigmaFactor;
            sig[j] = fSigma;
        }
        for (int o = 0; o < n_oct; o++) for (int j = 0; j < 6; j++) p->pyr.oct[o].sigma[j] = sig[j];
        if (p->n_init_taps < 0) return fail(ctx, S3D_ERR_UNSUPPORTED, "initial blur too wide");
    }

    // candidate lists: capacity X0*Y0 per list like the reference (MultiScale.cpp:255-267), bounded
    long long cc = (long long)X0 * Y0;
    if (cc > (1 << 20)) cc = 1 << 20;
    if (cc < 1024) cc = 1024;
    p->cand_cap = (int)cc;
    p->n_lists = n_oct * 3 * 2;
    if (p->n_lists > 0) {
        PA(&p->cand_raw, (size_t)p->n_lists * p->cand_cap);
        PA(&p->kp_stage, (size_t)p->n_lists * p->cand_cap);
        // volumes whose candidate lists can get long: group the lists by plane before ranking (12000 planes = 48 KB of counters)
        if ((long long)X0 * Y0 * Z0 >= ctx->tune.bucket_min_voxels && Z0 + 2 <= 12000) {
            p->plane_stride = Z0 + 2;
            PA(&p->cand_sorted, (size_t)p->n_lists * p->cand_cap);
            PA(&p->plane_off, (size_t)p->n_lists * p->plane_stride);
        }
        PA(&p->stage_flags, (size_t)p->n_lists * p->cand_cap);
    }
    PA(&p->counts, p->n_lists + 8 + kMaxOct * 3);
    p->face_counts = p->counts + p->n_lists + 8;
    for (int o = 0; o < n_oct; o++) {
        const OctaveDesc &od = p->pyr.oct[o];
        long long nv = (long long)od.pitch * od.Y * od.Z;
        for (int c = 0; c < 3; c++) {
            long long fc = nv / 8 + 1024;
            p->face_cap[o * 3 + c] = (int)(fc > (1 << 26) ? (1 << 26) : fc);
            PA(&p->face[o * 3 + c], (size_t)p->face_cap[o * 3 + c]);
        }
    }
    PA(&p->kps, kp_cap);
    PA(&p->nrows, kp_cap);
    PA(&p->row_off, kp_cap);
    PA(&p->kp_eigs, (size_t)kp_cap * 3);
    PA(&p->kp_ori0, (size_t)kp_cap * 9);
    PA(&p->kp_fmat, (size_t)kp_cap * 9);
    PA(&p->work_a, kp_cap);
    PA(&p->kp_rots, (size_t)kp_cap * PD * PD * 9);
    PA(&p->kp_p1, (size_t)kp_cap * PD * 3);
    PA(&p->kp_nprim, kp_cap);
    PA(&p->kp_nsec, (size_t)kp_cap * PD);
    PA(&p->work_b, (size_t)kp_cap * PD);
    PA(&p->row_map, row_cap);
    PA(&p->kp_patch0, (size_t)kp_cap * PV);
    PA(&p->feats, row_cap);
    if (p->keep_patches) {
        PA(&p->dbg_patches, (size_t)row_cap * PV);
        PA(&p->dbg_prerank, (size_t)row_cap * 64);
    }
#undef PA
    // padding columns must be zero everywhere once; stages keep them zero afterwards
    CK(cudaMemsetAsync(p->img0.p, 0, p->img0.elems() * sizeof(float), ctx->stream));
    CK(cudaMemsetAsync(p->tmp1, 0, p->img0.elems() * sizeof(float), ctx->stream));
    for (auto &v : p->g) CK(cudaMemsetAsync(v.p, 0, v.elems() * sizeof(float), ctx->stream));
    for (auto &v : p->d) CK(cudaMemsetAsync(v.p, 0, v.elems() * sizeof(float), ctx->stream));
    for (int o = 0; o < n_oct; o++)
        CK(cudaMemsetAsync(p->oct_tmp[o], 0, (size_t)p->pyr.oct[o].pitch * p->pyr.oct[o].Y * p->pyr.oct[o].Z * sizeof(float), ctx->stream));
    return S3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// the whole path, enqueued on the stream (input already in plan->stage)
// ---------------------------------------------------------------------------------------------------
static void mark(s3d_ctx *ctx, const char *name)
{
    if (!ctx->tune.timing) return;
    if (!strcmp(name, "start")) {     // keep only the latest extraction
        for (auto &m : ctx->marks) cudaEventDestroy(m.second);
        ctx->marks.clear();
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    ctx->marks.emplace_back(name, e);
}

static void report_marks(s3d_ctx *ctx)
{
    if (!ctx->tune.timing || ctx->marks.empty()) return;
    fprintf(stderr, "s3d stage timing (main stream, us since previous mark):\n");
    for (size_t i = 1; i < ctx->marks.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->marks[i - 1].second, ctx->marks[i].second);
        fprintf(stderr, "  %-28s %9.1f\n", ctx->marks[i].first.c_str(), ms * 1e3f);
    }
    float tot = 0;
    cudaEventElapsedTime(&tot, ctx->marks.front().second, ctx->marks.back().second);
    fprintf(stderr, "  %-28s %9.1f\n", "total", tot * 1e3f);
    if (!ctx->tune.use_graph) {
        for (auto &m : ctx->marks) cudaEventDestroy(m.second);
        ctx->marks.clear();
    }
}

// Detection branch of one level.  Level j completes DoG j-1, so centre level j-1 can be detected (it needs DoG j-2 and
// j-1) and centre level j-2 can be validated / refined (it needs DoG j-1).  Enqueued on `sd`.
static s3d_status enqueue_detection(s3d_ctx *ctx, Plan *p, const ListDesc &L, int o, int j, cudaStream_t sd, int *err, int refine_octaves = 1)
{
    const OctaveDesc &od = p->pyr.oct[o];
    ctx->cur = sd;
    int c_det = j - 1, c_ref = j - 2;
    // Octaves >= 1 of volumes without bucketed ranking: their chains are launch-latency bound, so the three centre levels
    // are detected by ONE launch (detect_multi_kernel; small octaves only: larger ones keep the two-pass detection per
    // level) and validated / refined by ONE launch (lists 6o .. 6o+5), both after level 5.  `refine_octaves` < 0: that
    // many consecutive octaves ENDING with this one (the tiny tail of the pyramid) share the two launches.
    if (o >= 1 && !p->cand_sorted) {
        if (ctx->tune.prof_skip & 2) return S3D_OK;
        const bool small = (long long)od.pitch * od.Y * od.Z < ctx->tune.detect2_min_voxels;
        if (!small && c_det <= 3) {
            int l0 = (o * 3 + (c_det - 1)) * 2;
            s3d_status s = detect_two_pass(ctx, od.d[c_det - 1], od.d[c_det], od.X, od.Y, od.Z, od.pitch,
                                           p->face[o * 3 + c_det - 1], p->face_counts + o * 3 + c_det - 1, p->face_cap[o * 3 + c_det - 1],
                                           p->cand_raw + (size_t)l0 * p->cand_cap, p->counts + l0,
                                           p->cand_raw + (size_t)(l0 + 1) * p->cand_cap, p->counts + l0 + 1, p->cand_cap, err, od.own0, od.own1);
            if (s != S3D_OK) return s;
        }
        if (j == 5 && refine_octaves != 0) {
            const int n_ref = refine_octaves > 0 ? refine_octaves : -refine_octaves;
            const int o_first = refine_octaves > 0 ? o : o - n_ref + 1;
            if (small) {
                for (int q0 = o_first; q0 <= o; q0 += kMaxDetectJobs / 3) {
                    DetectJobs J;
                    memset(&J, 0, sizeof(J));
                    J.cap = p->cand_cap;
                    int blocks = 0;
                    for (int q = q0; q <= o && q < q0 + kMaxDetectJobs / 3; q++) {
                        const OctaveDesc &oq = p->pyr.oct[q];
                        if (oq.X < 3 || oq.Y < 3 || oq.Z < 3) continue;      // no interior voxel
                        for (int c = 1; c <= 3; c++) {
                            DetectJob &t = J.job[J.n++];
                            const int l0 = (q * 3 + (c - 1)) * 2;
                            t.finer = oq.d[c - 1]; t.centre = oq.d[c];
                            t.X = oq.X; t.Y = oq.Y; t.Z = oq.Z; t.pitch = oq.pitch;
                            t.mins = CandList{ p->cand_raw + (size_t)l0 * p->cand_cap, p->counts + l0 };
                            t.maxs = CandList{ p->cand_raw + (size_t)(l0 + 1) * p->cand_cap, p->counts + l0 + 1 };
                            t.own0 = oq.own0; t.own1 = oq.own1;
                            t.nbx = (oq.X - 2 + 31) / 32; t.nby = (oq.Y - 2 + 7) / 8;
                            t.first_block = blocks;
                            blocks += t.nbx * t.nby * ((oq.Z - 2 + kDetectZ - 1) / kDetectZ);
                        }
                    }
                    if (J.n > 0) {
                        detect_multi_kernel<<<blocks, dim3(32, 8), 0, sd>>>(J);
                        ctx->launches++;
                    }
                }
            }
            cand_refine_kernel<<<dim3(6 * n_ref, 32), 256, 0, sd>>>(p->pyr, L, o_first * 6, p->kp_stage, p->stage_flags, err, nullptr, nullptr, 0);
            ctx->launches++;
            CK(cudaGetLastError());
        }
        return S3D_OK;
    }
    if (c_det <= 3 && !(ctx->tune.prof_skip & 2)) {
        int l0 = (o * 3 + (c_det - 1)) * 2;
        s3d_status s = detect_two_pass(ctx, od.d[c_det - 1], od.d[c_det], od.X, od.Y, od.Z, od.pitch,
                                       p->face[o * 3 + c_det - 1], p->face_counts + o * 3 + c_det - 1, p->face_cap[o * 3 + c_det - 1],
                                       p->cand_raw + (size_t)l0 * p->cand_cap, p->counts + l0,
                                       p->cand_raw + (size_t)(l0 + 1) * p->cand_cap, p->counts + l0 + 1, p->cand_cap, err, od.own0, od.own1);
        if (s != S3D_OK) return s;
    }
    if (c_ref >= 1 && !(ctx->tune.prof_skip & 2)) {
        int l0 = (o * 3 + (c_ref - 1)) * 2;
        if (p->cand_sorted) {
            cand_bucket_kernel<<<2, 1024, (od.Z + 2) * sizeof(int), sd>>>(L, l0, od.Z, p->cand_sorted, p->plane_off, p->plane_stride);
            ctx->launches++;
        }
        cand_refine_kernel<<<dim3(2, p->cand_sorted ? 128 : 32), 256, 0, sd>>>(p->pyr, L, l0, p->kp_stage, p->stage_flags, err,
                                                                              p->cand_sorted, p->plane_off, p->plane_stride);
        ctx->launches++;
    }
    return S3D_OK;
}

static s3d_status enqueue_pipeline(s3d_ctx *ctx, const s3d_params *prm)
{
    Plan *p = ctx->plan;
    cudaStream_t st = ctx->stream;
    int *kp_count = p->counts + p->n_lists, *n_features = kp_count + 1, *err = kp_count + 2;
    ctx->cur = st;
    mark(ctx, "start");
    if (ctx->d_stamps) stamp_kernel<<<1, 1, 0, st>>>(ctx->d_stamps + 2);
    zero_ints_kernel<<<1, 256, 0, st>>>(p->counts, p->n_lists + 8 + kMaxOct * 3);
    ctx->launches++;

    // pre-step: -2+ / -2- / plain copy into the pitched pyramid input (or straight into level 0)
    // (the plain / level-0 input was already laid out in its pitched buffer by stage_input())
    if (p->input_is_g0) {
        if (p->n_oct == 0) { ctx->has_result = true; return S3D_OK; }
    } else {
        if (p->double_mode == 1) {
            s3d_status s = resize_launch(ctx, 2, p->stage.p, p->X, p->Y, p->Z, p->X, p->img0.p, p->img0.pitch);
            if (s != S3D_OK) return s;
        } else if (p->double_mode == -1) {
            s3d_status s = resize_launch(ctx, 1, p->stage.p, p->X, p->Y, p->Z, p->X, p->img0.p, p->img0.pitch);
            if (s != S3D_OK) return s;
        }
        if (p->n_oct == 0) {
            ctx->has_result = true;
            return S3D_OK;
        }
        mark(ctx, "pre-step/pad");
        // initial blur (MultiScale.cpp:298)
        Vol &g0 = p->g[0];
        s3d_status s = blur3d(ctx, p->img0.p, p->tmp1, g0.p, g0.X, g0.Y, g0.Z, g0.pitch, p->init_taps, p->n_init_taps, nullptr);
        if (s != S3D_OK) return s;
        mark(ctx, "initial blur");
    }
    // Octave o+1 only needs level 3 of octave o, so every octave runs on its own branch (stream / graph
    // branch): the tail of an octave overlaps the whole chain of smaller octaves, which is launch-latency bound.
    ListDesc L{ p->n_lists, p->cand_cap, p->cand_raw, p->counts };
    // The last octaves are a few thousand voxels each: from the first octave on from which all of them fit, one launch
    // of tiny_octaves_kernel replaces their blur levels and subsamples (s3d_tiny.cuh); their detection branches follow it.
    int tiny_first = p->n_oct;
    if (ctx->tune.tiny) {
        bool taps_ok = true;
        for (int j = 0; j < 5; j++) taps_ok = taps_ok && p->n_lvl_taps[j] <= 2 * kMaxFastR + 1;
        for (int o = p->n_oct - 1; o >= 0 && taps_ok && p->n_oct - o <= kTinyMaxOct; o--) {
            const OctaveDesc &q = p->pyr.oct[o];
            const bool whole = q.z_off == 0 && q.Zg == q.Z && q.own0 == 0 && q.own1 == q.Z;       // not a z slab
            if (!whole || (long long)q.pitch * q.Y * q.Z > kTinyMaxElems) break;
            tiny_first = o;
        }
    }
    for (int o = 0; o < p->n_oct; o++) {
        const OctaveDesc &od = p->pyr.oct[o];
        cudaStream_t so = (o == 0 || ctx->tune.serial) ? st : ctx->side[o];
        ctx->cur = so;
        if (o > 0) CK(cudaStreamWaitEvent(so, ctx->ev_fork[o - 1], 0));
        if (o == tiny_first) {
            TinyDesc td;
            memset(&td, 0, sizeof(td));
            td.n_oct = p->n_oct - o;
            for (int q = o; q < p->n_oct; q++) {
                const OctaveDesc &oq = p->pyr.oct[q];
                TinyOct &t = td.o[q - o];
                t.X = oq.X; t.Y = oq.Y; t.Z = oq.Z; t.pitch = oq.pitch;
                for (int j = 0; j < 6; j++) t.g[j] = p->g[q * 6 + j].p;
                for (int j = 0; j < 5; j++) t.d[j] = p->d[q * 5 + j].p;
            }
            for (int j = 0; j < 5; j++) {
                td.ntaps[j] = p->n_lvl_taps[j];
                for (int k = 0; k < p->n_lvl_taps[j]; k++) td.taps[j][k] = p->lvl_taps[j][k];
            }
            tiny_octaves_kernel<<<1, kTinyThreads, kTinySmem, so>>>(td);
            ctx->launches++;
            CK(cudaEventRecord(ctx->ev_lvl[o][0], so));
            if (o >= 1 && !p->cand_sorted) {
                // every detection of every tiny octave on one stream, then one refinement launch for all their lists
                cudaStream_t sd = ctx->tune.serial ? st : ctx->det[o];
                CK(cudaStreamWaitEvent(sd, ctx->ev_lvl[o][0], 0));
                for (int q = o; q < p->n_oct; q++)
                    for (int j = 2; j < 6; j++) {
                        s3d_status s = enqueue_detection(ctx, p, L, q, j, sd, err, (q == p->n_oct - 1 && j == 5) ? -(p->n_oct - o) : 0);
                        if (s != S3D_OK) { ctx->cur = st; return s; }
                    }
                for (int q = o; q < p->n_oct; q++) CK(cudaEventRecord(ctx->ev_done[q], sd));
            } else {
                for (int q = o; q < p->n_oct; q++) {
                    cudaStream_t sd = ctx->tune.serial ? st : ctx->det[q];
                    CK(cudaStreamWaitEvent(sd, ctx->ev_lvl[o][0], 0));
                    for (int j = 2; j < 6; j++) {
                        s3d_status s = enqueue_detection(ctx, p, L, q, j, sd, err);
                        if (s != S3D_OK) { ctx->cur = st; return s; }
                    }
                    CK(cudaEventRecord(ctx->ev_done[q], sd));
                }
            }
            ctx->cur = so;
            break;
        }
        for (int j = 1; j < 6; j++) {
            Vol &a = p->g[o * 6 + j - 1], &b = p->g[o * 6 + j], &dd = p->d[o * 5 + j - 1];
            s3d_status s = S3D_OK;
            if (!(j == 5 && (ctx->tune.prof_skip & 16)))      // profiling only: what the top level of every octave costs
                s = blur3d(ctx, a.p, p->oct_tmp[o], b.p, od.X, od.Y, od.Z, od.pitch, p->lvl_taps[j - 1], p->n_lvl_taps[j - 1], dd.p);
            if (s != S3D_OK) { ctx->cur = st; return s; }
            if (j == 3 && o + 1 < p->n_oct) {
                const OctaveDesc &nx = p->pyr.oct[o + 1];
                s = resize_launch(ctx, 0, b.p, od.X, od.Y, od.Z, od.pitch, p->g[(o + 1) * 6].p, nx.pitch);
                if (s != S3D_OK) { ctx->cur = st; return s; }
                CK(cudaEventRecord(ctx->ev_fork[o], so));
            }
            if (o == 0) { char nm[32]; snprintf(nm, sizeof(nm), "oct0 level %d", j); mark(ctx, nm); }
            if (j >= 2) {
                // Detection branch.  Level j completes DoG j-1, so centre level j-1 can be detected now (it
                // needs DoG j-2 and j-1) and centre level j-2 can be validated / refined (it needs DoG j-1).
                // The branch runs beside the blur of the next level; only the refinement of centre level 3
                // (after level 5) is exposed.
                CK(cudaEventRecord(ctx->ev_lvl[o][j - 2], so));
                cudaStream_t sd = ctx->tune.serial ? st : ctx->det[o];
                CK(cudaStreamWaitEvent(sd, ctx->ev_lvl[o][j - 2], 0));
                s = enqueue_detection(ctx, p, L, o, j, sd, err);
                if (s != S3D_OK) { ctx->cur = st; return s; }
                if (j == 5) CK(cudaEventRecord(ctx->ev_done[o], sd));
                ctx->cur = so;
            }
        }
    }
    ctx->cur = st;
    for (int o = 0; o < p->n_oct; o++) CK(cudaStreamWaitEvent(st, ctx->ev_done[o], 0));
    mark(ctx, "join (detect+refine tails)");
    if (ctx->tune.prof_skip & 1) {
        if (ctx->tune.prof_sleep_us > 0) sleep_kernel<<<1, 1, 0, st>>>(ctx->tune.prof_sleep_us);
        CK(cudaGetLastError());
        return S3D_OK;
    }
    compact_kernel<<<1, 1024, 0, st>>>(L, p->kp_stage, p->stage_flags, p->kps, kp_count, p->kp_cap, err);
    mark(ctx, "compact");
    // orientation: per keypoint, then per (keypoint, primary direction)
    float eig = prm->eig_thres;
    int *work_b_count = kp_count + 3, *work_a_count = kp_count + 4;
    orient_patch_kernel<<<ctx->sm_count * ctx->tune.tail_a, 256, sizeof(PatchSmem), st>>>(p->pyr, p->kps, kp_count, p->kp_patch0, p->kp_fmat);
    int grid_svd = (p->kp_cap + 31) / 32;
    if (grid_svd > ctx->sm_count * 32) grid_svd = ctx->sm_count * 32;
    orient_svd_kernel<<<grid_svd, 32, 0, st>>>(p->kp_fmat, kp_count, eig, p->kp_eigs, p->kp_ori0, p->kp_nprim, p->work_a, work_a_count);
    orient_hist_kernel<<<ctx->sm_count * ctx->tune.tail_a, kHistThreads, sizeof(HistSmem), st>>>(p->work_a, work_a_count, p->kp_patch0, p->kp_nprim,
                                                                                                  p->kp_p1, p->work_b, work_b_count);
    ctx->launches += 2;
    mark(ctx, "orient_patch+svd+hist");
    if (ctx->tune.prof_skip & 8) { CK(cudaGetLastError()); return S3D_OK; }
    orient_b_kernel<<<ctx->sm_count * ctx->tune.tail_b, kHistThreads, sizeof(HistSmem), st>>>(p->work_b, work_b_count, p->kp_p1, p->kp_patch0, p->kp_nsec, p->kp_rots);
    mark(ctx, "orient_b");
    row_offsets_kernel<<<1, 1024, 0, st>>>(p->kp_nprim, p->kp_nsec, kp_count, p->nrows, p->row_off, p->row_map, n_features, p->row_cap, err);
    float size_factor = 1.0f;
    if (p->double_mode > 0 || p->pre_step_done > 0) size_factor /= 2; else if (p->double_mode < 0 || p->pre_step_done < 0) size_factor *= 2;
    if (ctx->tune.prof_skip & 4) { CK(cudaGetLastError()); return S3D_OK; }
    int grid_d = ctx->sm_count * ctx->tune.tail_d;
    describe_kernel<<<grid_d, ctx->tune.desc_threads, sizeof(DescribeSmem), st>>>(p->pyr, p->kps, n_features, p->row_map, p->kp_nsec, p->kp_eigs,
                                                              p->kp_ori0, p->kp_rots, p->kp_patch0, prm->descriptor, size_factor, p->octave_base,
                                                              p->row_cap, p->feats, p->dbg_patches, p->dbg_prerank);
    mark(ctx, "row_offsets+describe");
    if (ctx->d_stamps) stamp_kernel<<<1, 1, 0, st>>>(ctx->d_stamps + 3);
    ctx->launches += 5;
    CK(cudaGetLastError());
    return S3D_OK;
}

static s3d_status check_params(s3d_ctx *ctx, const float *vol, int X, int Y, int Z, const s3d_params *prm)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (!vol || !prm) return fail(ctx, S3D_ERR_INVALID, "null volume or params");
    if (X < 1 || Y < 1 || Z < 1) return fail(ctx, S3D_ERR_INVALID, "bad dimensions");
    if (prm->double_mode < -1 || prm->double_mode > 1) return fail(ctx, S3D_ERR_INVALID, "double_mode must be -1, 0 or +1");
    if (prm->descriptor < 0 || prm->descriptor > 3) return fail(ctx, S3D_ERR_INVALID, "unknown descriptor");
    // capacities bound the per-keypoint work arrays (121 rotations x 9 floats and an 11^3 patch per keypoint)
    if (prm->max_keypoints < 0 || prm->max_keypoints > (1 << 21) || prm->max_features < 0 || prm->max_features > (1 << 25))
        return fail(ctx, S3D_ERR_INVALID, "max_keypoints must be <= 2^21 and max_features <= 2^25");
    if (prm->double_mode != 0 && (X < 2 || Y < 2 || Z < 2)) return fail(ctx, S3D_ERR_INVALID, "volume too small for -2+/-2-");
    if (prm->input_is_g0 && prm->double_mode != 0) return fail(ctx, S3D_ERR_INVALID, "input_is_g0 excludes -2+/-2-");
    if (prm->pre_step_done < -1 || prm->pre_step_done > 1 || (prm->pre_step_done != 0 && prm->double_mode != 0))
        return fail(ctx, S3D_ERR_INVALID, "pre_step_done must be -1, 0 or +1 and excludes double_mode");
    if (prm->octave_base < 0 || prm->octave_base > 20 || prm->max_octaves < 0) return fail(ctx, S3D_ERR_INVALID, "bad octave_base / max_octaves");
    if (prm->slab) {
        if (prm->max_octaves != 1 || prm->double_mode != 0) return fail(ctx, S3D_ERR_INVALID, "slab mode runs one octave per call, without pre-step");
        if (prm->z_off < 0 || prm->z_off + Z > prm->z_global || prm->own_z0 < prm->z_off || prm->own_z1 > prm->z_off + Z || prm->own_z0 > prm->own_z1)
            return fail(ctx, S3D_ERR_INVALID, "slab ranges are inconsistent");
    }
    return S3D_OK;
}

// run the pipeline on plan->stage (already filled, stream-ordered): direct launches or graph replay
static s3d_status run_pipeline(s3d_ctx *ctx, const s3d_params *prm)
{
    Plan *p = ctx->plan;
    ctx->has_result = false;
    if (!ctx->tune.use_graph) {
        ctx->launches = 0;
        ctx->in_pipeline = true;
        s3d_status s = enqueue_pipeline(ctx, prm);
        ctx->in_pipeline = false;
        ctx->last_launches = ctx->launches;
        if (s == S3D_OK) ctx->has_result = true;
        return s;
    }
    if (!p->graph || p->graph_descriptor != prm->descriptor || p->graph_eig != prm->eig_thres) {
        if (p->graph) { cudaGraphExecDestroy(p->graph); p->graph = nullptr; }
        cudaGraph_t g = nullptr;
        for (auto &m : ctx->marks) cudaEventDestroy(m.second);
        ctx->marks.clear();
        CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        ctx->launches = 0;
        ctx->in_pipeline = true;
        s3d_status s = enqueue_pipeline(ctx, prm);
        ctx->in_pipeline = false;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (s != S3D_OK) { if (g) cudaGraphDestroy(g); return s; }
        if (e != cudaSuccess) { ctx->err = std::string("graph capture failed: ") + cudaGetErrorString(e); return S3D_ERR_CUDA; }
        e = cudaGraphInstantiate(&p->graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { ctx->err = std::string("graph instantiate failed: ") + cudaGetErrorString(e); p->graph = nullptr; return S3D_ERR_CUDA; }
        p->graph_descriptor = prm->descriptor;
        p->graph_eig = prm->eig_thres;
        ctx->last_launches = ctx->launches;
    }
    CK(cudaGraphLaunch(p->graph, ctx->stream));
    ctx->has_result = true;
    return S3D_OK;
}

// Bring the caller's dense volume into the plan: straight into the pitched pyramid input (or level 0)
// with a strided copy when no resize pre-step is needed, else into the dense staging buffer.
// enqueue the speculative result copy behind the pipeline (see s3d_ctx::spec_enable)
static s3d_status enqueue_result_copy(s3d_ctx *ctx)
{
    ctx->spec_valid = false;
    ctx->spec_rows = 0;
    if (!ctx->spec_enable || !ctx->plan || !ctx->has_result) return S3D_OK;
    Plan *p = ctx->plan;
    int want = ctx->spec_guess < p->row_cap ? ctx->spec_guess : p->row_cap;
    if (want > ctx->h_rows_cap) {
        int cap = want + want / 2;
        if (cap > p->row_cap) cap = p->row_cap;
        s3d_feature *q = nullptr;
        if (cudaHostAlloc((void **)&q, sizeof(s3d_feature) * (size_t)cap, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return S3D_OK;          // no pinned memory: the fetch falls back to the two-step path
        }
        if (ctx->h_rows) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->h_rows); }
        ctx->h_rows = q;
        ctx->h_rows_cap = cap;
    }
    CK(cudaMemcpyAsync(ctx->h_counts, p->counts + p->n_lists, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (want > 0) CK(cudaMemcpyAsync(ctx->h_rows, p->feats, sizeof(s3d_feature) * (size_t)want, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->spec_rows = want;
    ctx->spec_valid = true;
    return S3D_OK;
}

static int dtype_bytes(int dtype)
{
    switch (dtype) {
    case S3D_DT_UINT8: case S3D_DT_INT8: return 1;
    case S3D_DT_INT16: case S3D_DT_UINT16: return 2;
    case S3D_DT_INT32: case S3D_DT_UINT32: case S3D_DT_FLOAT32: return 4;
    case S3D_DT_FLOAT64: return 8;
    default: return 0;
    }
}

template <typename T>
static void launch_convert(s3d_ctx *ctx, const void *src, int X, size_t rows, float *dst, int pitch)
{
    const size_t row_blocks = (rows + 3) / 4, max_blocks = (size_t)ctx->sm_count * 16;
    dim3 block(64, 4), grid((pitch + 255) / 256, (unsigned)(row_blocks < max_blocks ? row_blocks : max_blocks));
    convert_rows_kernel<T><<<grid, block, 0, ctx->stream>>>((const T *)src, X, (long long)rows, dst, pitch);
}

static s3d_status stage_input(s3d_ctx *ctx, const void *src_any, bool from_host, int dtype = S3D_DT_FLOAT32)
{
    Plan *p = ctx->plan;
    size_t rows = (size_t)p->Y * p->Z;
    if (dtype != S3D_DT_FLOAT32) {
        // typed host input: raw bytes over PCIe, cast + re-pitch on the device (dense when a resize pre-step follows)
        const int bpv = dtype_bytes(dtype);
        if (bpv == 0 || !from_host) return fail(ctx, S3D_ERR_INVALID, "unsupported input datatype");
        const size_t bytes = (size_t)bpv * p->X * rows;
        if (p->raw_bytes < bytes) {
            void *q = nullptr;
            cudaError_t e = cudaMalloc(&q, bytes + 256);
            if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e); cudaGetLastError(); return S3D_ERR_NOMEM; }
            p->allocs.push_back(q);
            p->raw_stage = q;
            p->raw_bytes = bytes;
        }
        CK(cudaMemcpyAsync(p->raw_stage, src_any, bytes, cudaMemcpyHostToDevice, ctx->stream));
        Vol &dst = (p->double_mode != 0) ? p->stage : (p->input_is_g0 ? (p->n_oct > 0 ? p->g[0] : p->img0) : p->img0);
        switch (dtype) {
        case S3D_DT_UINT8: launch_convert<unsigned char>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        case S3D_DT_INT8: launch_convert<signed char>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        case S3D_DT_INT16: launch_convert<short>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        case S3D_DT_UINT16: launch_convert<unsigned short>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        case S3D_DT_INT32: launch_convert<int>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        case S3D_DT_UINT32: launch_convert<unsigned int>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        default: launch_convert<double>(ctx, p->raw_stage, p->X, rows, dst.p, dst.pitch); break;
        }
        CK(cudaGetLastError());
        return S3D_OK;
    }
    const float *src = (const float *)src_any;
    size_t row = sizeof(float) * (size_t)p->X;
    if (p->double_mode != 0) {
        CK(cudaMemcpyAsync(p->stage.p, src, row * rows, from_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
        return S3D_OK;
    }
    Vol &dst = p->input_is_g0 ? (p->n_oct > 0 ? p->g[0] : p->img0) : p->img0;
    if (from_host) {   // one contiguous PCIe transfer, then the re-pitch on the device
        CK(cudaMemcpyAsync(p->stage.p, src, row * rows, cudaMemcpyHostToDevice, ctx->stream));
        src = p->stage.p;
    }
    // dense -> pitched in one pass straight from the caller's buffer (outside the graph: the source
    // pointer changes from call to call)
    const size_t row_blocks = (rows + 3) / 4, max_blocks = (size_t)ctx->sm_count * 16;     // grid-stride over rows: blocks live long enough to overlap their loads
    dim3 block(64, 4), grid((dst.pitch + 255) / 256, (unsigned)(row_blocks < max_blocks ? row_blocks : max_blocks));
    if (ctx->d_stamps) stamp_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_stamps + 0);
    pad_rows_kernel<<<grid, block, 0, ctx->stream>>>(src, p->X, (long long)rows, dst.p, dst.pitch);
    if (ctx->d_stamps) stamp_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_stamps + 1);
    CK(cudaGetLastError());
    return S3D_OK;
}

extern "C" s3d_status s3d_extract_device(s3d_ctx *ctx, const float *d_volume, int X, int Y, int Z, const s3d_params *prm)
{
    s3d_status s = check_params(ctx, d_volume, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    CK(cudaSetDevice(ctx->device));
    s = plan_build(ctx, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    s = stage_input(ctx, d_volume, false);
    if (s != S3D_OK) return s;
    s = run_pipeline(ctx, prm);
    if (s != S3D_OK) return s;
    return enqueue_result_copy(ctx);
}

extern "C" s3d_status s3d_extract_host_async(s3d_ctx *ctx, const float *h_volume, int X, int Y, int Z, const s3d_params *prm)
{
    s3d_status s = check_params(ctx, h_volume, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    CK(cudaSetDevice(ctx->device));
    s = plan_build(ctx, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    s = stage_input(ctx, h_volume, true);
    if (s != S3D_OK) return s;
    s = run_pipeline(ctx, prm);
    if (s != S3D_OK) return s;
    return enqueue_result_copy(ctx);
}

extern "C" s3d_status s3d_extract_typed_async(s3d_ctx *ctx, const void *h_volume, int dtype, int X, int Y, int Z, const s3d_params *prm)
{
    if (dtype == S3D_DT_FLOAT32) return s3d_extract_host_async(ctx, (const float *)h_volume, X, Y, Z, prm);
    s3d_status s = check_params(ctx, (const float *)h_volume, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    if (dtype_bytes(dtype) == 0) return fail(ctx, S3D_ERR_INVALID, "unsupported input datatype");
    CK(cudaSetDevice(ctx->device));
    s = plan_build(ctx, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    s = stage_input(ctx, h_volume, true, dtype);
    if (s != S3D_OK) return s;
    s = run_pipeline(ctx, prm);
    if (s != S3D_OK) return s;
    return enqueue_result_copy(ctx);
}

static s3d_status fetch_counts(s3d_ctx *ctx)
{
    if (!ctx->plan || !ctx->has_result) return fail(ctx, S3D_ERR_INVALID, "no extraction has been run");
    Plan *p = ctx->plan;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->spec_valid)
        CK(cudaMemcpyAsync(ctx->h_counts, p->counts + p->n_lists, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    report_marks(ctx);
    int err = ctx->h_counts[2];
    if (err) {
        char b[256];
        snprintf(b, sizeof(b), "capacity exceeded:%s%s%s (raise s3d_params.max_keypoints / max_features)",
                 (err & ERR_CAND_OVERFLOW) ? " candidate list" : "", (err & ERR_KP_OVERFLOW) ? " keypoints" : "",
                 (err & ERR_ROW_OVERFLOW) ? " feature rows" : "");
        ctx->err = b;
        return S3D_ERR_CAPACITY;
    }
    return S3D_OK;
}

extern "C" s3d_status s3d_fetch_counts(s3d_ctx *ctx, int *n_keypoints, int *n_features)
{
    if (!ctx) return S3D_ERR_INVALID;
    s3d_status s = fetch_counts(ctx);
    if (s != S3D_OK) return s;
    if (n_keypoints) *n_keypoints = ctx->h_counts[0];
    if (n_features) *n_features = ctx->h_counts[1];
    return S3D_OK;
}

extern "C" s3d_status s3d_fetch_features(s3d_ctx *ctx, s3d_feature **out, int *n_out)
{
    if (!ctx || !out || !n_out) return S3D_ERR_INVALID;
    *out = nullptr; *n_out = 0;
    s3d_status s = fetch_counts(ctx);
    if (s != S3D_OK) return s;
    int n = ctx->h_counts[1];
    s3d_feature *h = (s3d_feature *)malloc(sizeof(s3d_feature) * (size_t)(n > 0 ? n : 1));
    if (!h) return fail(ctx, S3D_ERR_NOMEM, "host allocation failed");
    int have = 0;
    if (ctx->spec_valid && ctx->spec_rows > 0) {      // rows already in pinned memory
        have = n < ctx->spec_rows ? n : ctx->spec_rows;
        memcpy(h, ctx->h_rows, sizeof(s3d_feature) * (size_t)have);
    }
    if (n > have) {
        cudaError_t e = cudaMemcpyAsync(h + have, ctx->plan->feats + have, sizeof(s3d_feature) * (size_t)(n - have),
                                        cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { free(h); ctx->err = cudaGetErrorString(e); return S3D_ERR_CUDA; }
    }
    if (ctx->spec_enable) {                           // next guess: 1.25 x this count, at least 1024 rows
        int g = n + n / 4;
        ctx->spec_guess = g < 1024 ? 1024 : g;
    }
    *out = h; *n_out = n;
    return S3D_OK;
}

extern "C" s3d_status s3d_result_device(s3d_ctx *ctx, const s3d_feature **d_features, const int **d_n_features)
{
    if (!ctx || !ctx->plan || !ctx->has_result) return S3D_ERR_INVALID;
    if (d_features) *d_features = ctx->plan->feats;
    if (d_n_features) *d_n_features = ctx->plan->counts + ctx->plan->n_lists + 1;
    return S3D_OK;
}

extern "C" s3d_status s3d_extract(s3d_ctx *ctx, const float *h_volume, int X, int Y, int Z, const s3d_params *prm,
                                  s3d_feature **out, int *n_out)
{
    if (!out || !n_out) return S3D_ERR_INVALID;
    if (ctx) ctx->spec_enable = true;
    s3d_status s = s3d_extract_host_async(ctx, h_volume, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    return s3d_fetch_features(ctx, out, n_out);
}

extern "C" s3d_status s3d_extract_typed(s3d_ctx *ctx, const void *h_volume, int dtype, int X, int Y, int Z, const s3d_params *prm,
                                        s3d_feature **out, int *n_out)
{
    if (!out || !n_out) return S3D_ERR_INVALID;
    if (ctx) ctx->spec_enable = true;
    s3d_status s = s3d_extract_typed_async(ctx, h_volume, dtype, X, Y, Z, prm);
    if (s != S3D_OK) return s;
    return s3d_fetch_features(ctx, out, n_out);
}

// ---------------------------------------------------------------------------------------------------
// batch level: n_contexts extraction contexts on one device, fed round-robin
// ---------------------------------------------------------------------------------------------------
struct s3d_batch {
    int device = 0;
    std::vector<s3d_ctx *> ctx;
    std::string err;
    int launches = 0;
};

extern "C" s3d_status s3d_batch_create(int device, int n_contexts, s3d_batch **out)
{
    if (!out || n_contexts < 1 || n_contexts > 64) return S3D_ERR_INVALID;
    *out = nullptr;
    s3d_batch *b = new s3d_batch();
    b->device = device;
    for (int i = 0; i < n_contexts; i++) {
        s3d_ctx *c = nullptr;
        s3d_status st = s3d_ctx_create(device, &c);
        if (st != S3D_OK) {
            if (c) s3d_ctx_destroy(c);
            for (s3d_ctx *q : b->ctx) s3d_ctx_destroy(q);
            delete b;
            return st;
        }
        if (!c->tune.xy2_ctas_forced && n_contexts > 1) c->tune.xy2_ctas = 1;      // throughput mode, see launch_blur_xy2
        // ... of 128 threads: half the registers and warps of the x+y CTA for the kernels of the other volumes on the SM
        // (432.6 -> 426.4 us per volume; alone on the GPU 256 threads are faster)
        if (!c->tune.xy2_threads_forced && n_contexts > 1) c->tune.xy2.threads = 128;
        // ... and the z march runs as one segment: the threads it lacks to cover the memory latency alone are
        // provided by the other volumes in flight, and no halo planes are read twice (S3D_MARCH_TARGET overrides)
        if (c->tune.march_target == 0 && n_contexts > 1) c->tune.march_target = 1;
        if (!c->tune.detect_ctas_forced && n_contexts > 1) c->tune.detect_ctas = 4;
        // the one-kernel blur level: one resident CTA per SM instead of two (measured 529 -> 509 us per volume with 6 contexts)
        if (!c->tune.f4_ctas_forced && n_contexts > 1) c->tune.f4_ctas = c->sm_count;
        if (!c->tune.f4_wide_forced && n_contexts > 1) c->tune.f4_wide_min_voxels = 0;
        // describe: few CTAs, so that each walks full groups of rows (their NormalizeData sums side by side); a context
        // alone on the GPU spreads one row per CTA instead (shortest chain)
        if (!c->tune.tail_forced && n_contexts > 1) c->tune.tail_d = 2;
        b->ctx.push_back(c);
    }
    *out = b;
    return S3D_OK;
}

extern "C" void s3d_batch_destroy(s3d_batch *b)
{
    if (!b) return;
    for (s3d_ctx *c : b->ctx) s3d_ctx_destroy(c);
    delete b;
}

extern "C" const char *s3d_batch_last_error(const s3d_batch *b) { return b ? b->err.c_str() : "null batch"; }
extern "C" int s3d_batch_launches_per_volume(s3d_batch *b) { return b ? b->launches : 0; }

// volume i runs on context i % n; before a context is reused its previous volume is collected
static s3d_status batch_run(s3d_batch *b, const void *const *vols, int n, int X, int Y, int Z, const s3d_params *prm,
                            bool from_host, int dtype, s3d_feature **rows, int *n_rows, int *n_kps)
{
    if (!b || !vols || n < 0 || !prm) return S3D_ERR_INVALID;
    const int nc = (int)b->ctx.size();
    for (s3d_ctx *c : b->ctx) c->spec_enable = (rows != nullptr);     // rows wanted: copy them behind each graph
    std::vector<int> pending(nc, -1);
    auto collect = [&](int c) -> s3d_status {
        const int j = pending[c];
        pending[c] = -1;
        if (j < 0) return S3D_OK;
        s3d_status st;
        if (rows) {
            int nr = 0;
            st = s3d_fetch_features(b->ctx[c], &rows[j], &nr);
            if (n_rows) n_rows[j] = nr;
            if (n_kps) n_kps[j] = b->ctx[c]->h_counts[0];
        } else {
            int nk = 0, nr = 0;
            st = s3d_fetch_counts(b->ctx[c], &nk, &nr);
            if (n_rows) n_rows[j] = nr;
            if (n_kps) n_kps[j] = nk;
        }
        if (st != S3D_OK) b->err = b->ctx[c]->err;
        return st;
    };
    s3d_status first_err = S3D_OK;
    for (int i = 0; i < n && first_err == S3D_OK; i++) {
        const int c = i % nc;
        s3d_status st = collect(c);
        if (st == S3D_OK)
            st = from_host ? s3d_extract_typed_async(b->ctx[c], vols[i], dtype, X, Y, Z, prm)
                           : s3d_extract_device(b->ctx[c], (const float *)vols[i], X, Y, Z, prm);
        if (st != S3D_OK) { if (b->err.empty() || st != S3D_OK) b->err = b->ctx[c]->err; first_err = st; break; }
        pending[c] = i;
    }
    for (int k = 0; k < nc; k++) {      // drain in submission order
        const int c = (n + k) % nc;
        s3d_status st = collect(c);
        if (st != S3D_OK && first_err == S3D_OK) first_err = st;
    }
    if (nc > 0) b->launches = b->ctx[0]->last_launches;
    return first_err;
}

extern "C" s3d_status s3d_batch_extract(s3d_batch *b, const float *const *h_volumes, int n_volumes, int X, int Y, int Z,
                                        const s3d_params *prm, s3d_feature **rows, int *n_rows)
{
    if (!rows || !n_rows) return S3D_ERR_INVALID;
    for (int i = 0; i < n_volumes; i++) { rows[i] = nullptr; n_rows[i] = 0; }
    return batch_run(b, (const void *const *)h_volumes, n_volumes, X, Y, Z, prm, true, S3D_DT_FLOAT32, rows, n_rows, nullptr);
}

extern "C" s3d_status s3d_batch_extract_typed(s3d_batch *b, const void *const *h_volumes, int dtype, int n_volumes, int X, int Y, int Z,
                                              const s3d_params *prm, s3d_feature **rows, int *n_rows)
{
    if (!rows || !n_rows) return S3D_ERR_INVALID;
    for (int i = 0; i < n_volumes; i++) { rows[i] = nullptr; n_rows[i] = 0; }
    return batch_run(b, h_volumes, n_volumes, X, Y, Z, prm, true, dtype, rows, n_rows, nullptr);
}

extern "C" s3d_status s3d_batch_extract_device(s3d_batch *b, const float *const *d_volumes, int n_volumes, int X, int Y, int Z,
                                               const s3d_params *prm, int *n_keypoints, int *n_rows)
{
    return batch_run(b, (const void *const *)d_volumes, n_volumes, X, Y, Z, prm, false, S3D_DT_FLOAT32, nullptr, n_rows, n_keypoints);
}

extern "C" void *s3d_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void s3d_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------------
extern "C" int s3d_num_octaves(s3d_ctx *ctx) { return (ctx && ctx->plan) ? ctx->plan->n_oct : 0; }

extern "C" s3d_status s3d_get_level(s3d_ctx *ctx, int octave, int is_dog, int level, float *h_out, int dims[3])
{
    if (!ctx || !ctx->plan || !ctx->has_result) return S3D_ERR_INVALID;
    Plan *p = ctx->plan;
    if (octave < 0 || octave >= p->n_oct || level < 0 || level >= (is_dog ? 5 : 6)) return fail(ctx, S3D_ERR_INVALID, "no such level");
    CK(cudaSetDevice(ctx->device));
    Vol &v = is_dog ? p->d[octave * 5 + level] : p->g[octave * 6 + level];
    if (dims) { dims[0] = v.X; dims[1] = v.Y; dims[2] = v.Z; }
    if (h_out) {
        CK(cudaMemcpy2DAsync(h_out, sizeof(float) * v.X, v.p, sizeof(float) * v.pitch, sizeof(float) * v.X, (size_t)v.Y * v.Z,
                             cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return S3D_OK;
}

// Device pointer and row pitch (floats) of a pyramid level of the last extraction (valid until the next one on this
// context): lets a caller feed planes of a level straight into a stage call (the slab orchestration subsamples its
// own planes of level 3 in place instead of copying them out first).
extern "C" s3d_status s3d_level_device_ptr(s3d_ctx *ctx, int octave, int is_dog, int level, const float **d_ptr, int *pitch, int dims[3])
{
    if (!ctx || !d_ptr) return S3D_ERR_INVALID;
    if (!ctx->plan || !ctx->has_result) return fail(ctx, S3D_ERR_INVALID, "no extraction result");
    Plan *p = ctx->plan;
    if (octave < 0 || octave >= p->n_oct || level < 0 || level >= (is_dog ? 5 : 6)) return fail(ctx, S3D_ERR_INVALID, "bad octave / level");
    const Vol &v = is_dog ? p->d[octave * 5 + level] : p->g[octave * 6 + level];
    *d_ptr = v.p;
    if (pitch) *pitch = v.pitch;
    if (dims) { dims[0] = v.X; dims[1] = v.Y; dims[2] = v.Z; }
    return S3D_OK;
}

extern "C" s3d_status s3d_copy_level_device(s3d_ctx *ctx, int octave, int is_dog, int level, int z0, int z1, float *d_dst)
{
    if (!ctx || !ctx->plan || !ctx->has_result || !d_dst) return S3D_ERR_INVALID;
    Plan *p = ctx->plan;
    if (octave < 0 || octave >= p->n_oct || level < 0 || level >= (is_dog ? 5 : 6)) return fail(ctx, S3D_ERR_INVALID, "no such level");
    Vol &v = is_dog ? p->d[octave * 5 + level] : p->g[octave * 6 + level];
    if (z0 < 0 || z1 > v.Z || z0 >= z1) return fail(ctx, S3D_ERR_INVALID, "bad plane range");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy2DAsync(d_dst, sizeof(float) * v.X, v.p + (size_t)z0 * v.Y * v.pitch, sizeof(float) * v.pitch, sizeof(float) * v.X,
                         (size_t)v.Y * (z1 - z0), cudaMemcpyDeviceToDevice, ctx->stream));
    return S3D_OK;
}

extern "C" s3d_status s3d_get_row_keypoints(s3d_ctx *ctx, int **row_kp, int *n_out)
{
    if (!ctx || !row_kp || !n_out) return S3D_ERR_INVALID;
    s3d_status s = fetch_counts(ctx);
    if (s != S3D_OK) return s;
    int n = ctx->h_counts[1];
    int *h = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(h, ctx->plan->row_map, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < n; i++) h[i] /= kMaxRowsPerKp;
    }
    *row_kp = h; *n_out = n;
    return S3D_OK;
}

extern "C" s3d_status s3d_get_keypoints(s3d_ctx *ctx, s3d_keypoint **out, int *n_out)
{
    if (!ctx || !out || !n_out) return S3D_ERR_INVALID;
    s3d_status s = fetch_counts(ctx);
    if (s != S3D_OK) return s;
    int n = ctx->h_counts[0];
    s3d_keypoint *h = (s3d_keypoint *)malloc(sizeof(s3d_keypoint) * (size_t)(n > 0 ? n : 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(h, ctx->plan->kps, sizeof(s3d_keypoint) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *out = h; *n_out = n;
    return S3D_OK;
}

extern "C" s3d_status s3d_get_patches(s3d_ctx *ctx, float **patches, float **prerank, int *n_out)
{
    if (!ctx || !n_out) return S3D_ERR_INVALID;
    s3d_status s = fetch_counts(ctx);
    if (s != S3D_OK) return s;
    if (!ctx->plan->keep_patches) return fail(ctx, S3D_ERR_INVALID, "extraction was run without keep_patches");
    int n = ctx->h_counts[1];
    *n_out = n;
    if (patches) {
        *patches = (float *)malloc(sizeof(float) * PV * (size_t)(n > 0 ? n : 1));
        if (n > 0) CK(cudaMemcpyAsync(*patches, ctx->plan->dbg_patches, sizeof(float) * PV * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (prerank) {
        *prerank = (float *)malloc(sizeof(float) * 64 * (size_t)(n > 0 ? n : 1));
        if (n > 0) CK(cudaMemcpyAsync(*prerank, ctx->plan->dbg_prerank, sizeof(float) * 64 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return S3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// feature file writer (kept host code): msFeature3DVectorOutputText, reference MultiScale.h:386-474
// ---------------------------------------------------------------------------------------------------
extern "C" s3d_status s3d_write_features_text(const char *path, const s3d_feature *feats, int n, float fEigThres,
                                              int n_comments, const char *const *comments)
{
    if (!path || (n > 0 && !feats)) return S3D_ERR_INVALID;
    FILE *f = fopen(path, "wt");
    if (!f) return S3D_ERR_INVALID;
    auto keep = [&](const s3d_feature &ft) {
        float fEigSum = ft.eigs[0] + ft.eigs[1] + ft.eigs[2];
        float fEigPrd = ft.eigs[0] * ft.eigs[1] * ft.eigs[2];
        float fEigSumProd = fEigSum * fEigSum * fEigSum;
        return (fEigSumProd < fEigThres * fEigPrd || fEigThres < 0);
    };
    int cnt = 0;
    for (int i = 0; i < n; i++) if (keep(feats[i])) cnt++;
    fprintf(f, "# featExtract %s\n", "1.1");
    for (int i = 0; i < n_comments; i++) fprintf(f, "# %s\n", comments[i]);
    fprintf(f, "Features: %d\n", cnt);
    fprintf(f, "Scale-space location[x y z scale] orientation[o11 o12 o13 o21 o22 o23 o31 o32 o32] 2nd moment eigenvalues[e1 e2 e3] info flag[i1] descriptor[d1 .. d64]\n");
    for (int i = 0; i < n; i++) {
        const s3d_feature &ft = feats[i];
        if (!keep(ft)) continue;
        fprintf(f, "%f\t%f\t%f\t%f\t", ft.x, ft.y, ft.z, ft.scale);
        for (int j = 0; j < 9; j++) fprintf(f, "%f\t", ft.ori[j]);
        for (int j = 0; j < 3; j++) fprintf(f, "%f\t", ft.eigs[j]);
        fprintf(f, "%d\t", ft.flag);
        for (int j = 0; j < 64; j++) fprintf(f, "%i\t", (char)(ft.pc[j]));
        fprintf(f, "\n");
    }
    fclose(f);
    return S3D_OK;
}

// msFeature3DVectorOutputBin, reference MultiScale.h:228-303
extern "C" s3d_status s3d_write_features_bin(const char *path, const s3d_feature *feats, int n, float fEigThres)
{
    if (!path || (n > 0 && !feats)) return S3D_ERR_INVALID;
    FILE *f = fopen(path, "wb");
    if (!f) return S3D_ERR_INVALID;
    auto keep = [&](const s3d_feature &ft) {
        float fEigSum = ft.eigs[0] + ft.eigs[1] + ft.eigs[2];
        float fEigPrd = ft.eigs[0] * ft.eigs[1] * ft.eigs[2];
        float fEigSumProd = fEigSum * fEigSum * fEigSum;
        return (fEigSumProd < fEigThres * fEigPrd || fEigThres < 0);
    };
    int cnt = 0;
    for (int i = 0; i < n; i++) if (keep(feats[i])) cnt++;
    fprintf(f, "# featExtract %s\n", "1.1");
    fprintf(f, "Features: %d\n", cnt);
    for (int i = 0; i < n; i++) {
        const s3d_feature &ft = feats[i];
        if (!keep(ft)) continue;
        fwrite(&ft.x, sizeof(float), 1, f);
        fwrite(&ft.y, sizeof(float), 1, f);
        fwrite(&ft.z, sizeof(float), 1, f);
        fwrite(&ft.scale, sizeof(float), 1, f);
        fwrite(ft.ori, sizeof(float), 9, f);
        fwrite(ft.eigs, sizeof(float), 3, f);
        fwrite(&ft.flag, sizeof(unsigned int), 1, f);
        unsigned char pc[64];
        for (int j = 0; j < 64; j++) pc[j] = (unsigned char)(ft.pc[j]);
        fwrite(pc, sizeof(unsigned char), 64, f);
    }
    fclose(f);
    return S3D_OK;
}

// msFeature3DVectorInputText, reference MultiScale.h:305-384
extern "C" s3d_status s3d_read_features_text(const char *path, s3d_feature **out, int *n_out)
{
    if (!path || !out || !n_out) return S3D_ERR_INVALID;
    *out = nullptr; *n_out = 0;
    FILE *f = fopen(path, "rt");
    if (!f) return S3D_ERR_INVALID;
    char buff[400];
    buff[0] = '#';
    while (buff[0] == '#')       // read past comments
        if (!fgets(buff, sizeof(buff), f)) { fclose(f); return S3D_ERR_INVALID; }
    int n = 0;
    if (sscanf(buff, "Features: %d\n", &n) <= 0 || n <= 0) { fclose(f); return S3D_ERR_INVALID; }
    if (!fgets(buff, sizeof(buff), f) || !strstr(buff, "Scale-space location[x y z scale]")) { fclose(f); return S3D_ERR_INVALID; }
    s3d_feature *h = (s3d_feature *)malloc(sizeof(s3d_feature) * (size_t)n);
    if (!h) { fclose(f); return S3D_ERR_NOMEM; }
    for (int i = 0; i < n; i++) {
        s3d_feature &ft = h[i];
        bool ok = fscanf(f, "%f\t%f\t%f\t%f\t", &ft.x, &ft.y, &ft.z, &ft.scale) == 4;
        for (int j = 0; j < 9 && ok; j++) ok = fscanf(f, "%f\t", &ft.ori[j]) == 1;
        for (int j = 0; j < 3 && ok; j++) ok = fscanf(f, "%f\t", &ft.eigs[j]) == 1;
        int flag = 0;
        if (ok) ok = fscanf(f, "%d\t", &flag) == 1;
        ft.flag = (unsigned int)flag;
        for (int j = 0; j < 64 && ok; j++) ok = fscanf(f, "%f\t", &ft.pc[j]) == 1;
        if (!ok) { free(h); fclose(f); return S3D_ERR_INVALID; }     // the reference asserts here
    }
    fclose(f);
    *out = h; *n_out = n;
    return S3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// descriptor matching: exact k nearest neighbours on Feature3DInfo::DistSqrPCs
// ---------------------------------------------------------------------------------------------------
extern "C" s3d_status s3d_match_device(s3d_ctx *ctx, const s3d_feature *d_a, int nA, const s3d_feature *d_b, int nB, int k,
                                       int *d_idx, float *d_dist)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (nA < 0 || nB < 0 || k < 1 || k > kMatchMaxK || (nA > 0 && (!d_a || !d_idx || !d_dist)) || (nB > 0 && !d_b))
        return fail(ctx, S3D_ERR_INVALID, "s3d_match: bad argument (1 <= k <= 16)");
    if (nA == 0) return S3D_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    // database chunks: enough CTAs for ~2 per SM, chunks of at least 256 descriptors
    const int qblocks = (nA + kMatchThreads - 1) / kMatchThreads;
    int n_chunks = (2 * ctx->sm_count + qblocks - 1) / qblocks;
    if (n_chunks > (nB + 255) / 256) n_chunks = (nB + 255) / 256;
    if (n_chunks < 1) n_chunks = 1;
    const int chunk = nB > 0 ? (nB + n_chunks - 1) / n_chunks : 1;
    n_chunks = nB > 0 ? (nB + chunk - 1) / chunk : 1;
    const int K = match_list_len(k);
    float *part_d = nullptr; int *part_i = nullptr;
    CK(cudaMallocAsync((void **)&part_d, sizeof(float) * (size_t)n_chunks * nA * K, st));
    CK(cudaMallocAsync((void **)&part_i, sizeof(int) * (size_t)n_chunks * nA * K, st));
    cudaError_t e = launch_match(st, d_a, nA, d_b, nB, k, n_chunks, chunk, part_d, part_i, d_idx, d_dist);
    ctx->launches += 2;
    cudaFreeAsync(part_d, st);
    cudaFreeAsync(part_i, st);
    CK(e);
    return S3D_OK;
}

extern "C" s3d_status s3d_match(s3d_ctx *ctx, const s3d_feature *h_a, int nA, const s3d_feature *h_b, int nB, int k,
                                int *h_idx, float *h_dist)
{
    if (!ctx) return S3D_ERR_INVALID;
    if (nA < 0 || nB < 0 || k < 1 || k > kMatchMaxK || (nA > 0 && (!h_a || !h_idx || !h_dist)) || (nB > 0 && !h_b))
        return fail(ctx, S3D_ERR_INVALID, "s3d_match: bad argument (1 <= k <= 16)");
    if (nA == 0) return S3D_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    s3d_feature *d_a = nullptr, *d_b = nullptr; int *d_idx = nullptr; float *d_dist = nullptr;
    CK(cudaMallocAsync((void **)&d_a, sizeof(s3d_feature) * (size_t)nA, st));
    CK(cudaMallocAsync((void **)&d_b, sizeof(s3d_feature) * (size_t)(nB > 0 ? nB : 1), st));
    CK(cudaMallocAsync((void **)&d_idx, sizeof(int) * (size_t)nA * k, st));
    CK(cudaMallocAsync((void **)&d_dist, sizeof(float) * (size_t)nA * k, st));
    CK(cudaMemcpyAsync(d_a, h_a, sizeof(s3d_feature) * (size_t)nA, cudaMemcpyHostToDevice, st));
    if (nB > 0) CK(cudaMemcpyAsync(d_b, h_b, sizeof(s3d_feature) * (size_t)nB, cudaMemcpyHostToDevice, st));
    s3d_status s = s3d_match_device(ctx, d_a, nA, d_b, nB, k, d_idx, d_dist);
    if (s == S3D_OK) {
        CK(cudaMemcpyAsync(h_idx, d_idx, sizeof(int) * (size_t)nA * k, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(h_dist, d_dist, sizeof(float) * (size_t)nA * k, cudaMemcpyDeviceToHost, st));
    }
    cudaFreeAsync(d_a, st); cudaFreeAsync(d_b, st); cudaFreeAsync(d_idx, st); cudaFreeAsync(d_dist, st);
    CK(cudaStreamSynchronize(st));
    return s;
}
