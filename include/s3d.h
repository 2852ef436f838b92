/*
 * s3d.h -- C-ABI of the B200-native 3D SIFT engine (lib3dsift_b200.so).
 *
 * This is the drop-in boundary for the featExtract hot path of CarluerJB/3D_SIFT_CUDA.  It replaces
 * the reference's cuda_common layer (four C++ launchers taking FEATUREIO& + raw cudart calls scattered
 * through src_common) with plain pointers and sizes; each entry point cites what it replaces.
 *   R/ = 3dsift_cleanup-softVote_App_Weight_SoftMax/ in the reference tree.
 *
 * Conventions
 *   - Volumes are fp32, x fastest.  A volume is described by (X, Y, Z, pitch): element (x,y,z) lives at
 *     p[(z*Y + y)*pitch + x], pitch >= X (pitch == X is the reference's dense FEATUREIO layout,
 *     R/src_common/FeatureIO.cpp:739).  Columns X..pitch-1 are padding and are kept at zero by every
 *     stage.  Fast paths need pitch % 8 == 0 and 32-byte aligned bases; anything else takes a
 *     scalar path with identical results.
 *   - All device work is stream-ordered on the context's stream (or the stream given to
 *     s3d_ctx_create_on_stream, e.g. torch's current stream); stage-level calls never synchronise.
 *   - Every call returns an s3d_status; nothing ever calls exit() (the reference's gpuErrchk does,
 *     R/cuda_common/SIFT_cuda_Tools.cuh:13-21).  s3d_last_error() gives the text.
 *   - There is no CPU fallback: without a usable CUDA device s3d_ctx_create fails.
 *   - Arithmetic is the reference CPU path's, operation for operation (no FMA contraction, same
 *     summation order), so results are bit-identical to featExtract without -d.
 */
#ifndef S3D_H
#define S3D_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_VERSION 0x000100

typedef enum s3d_status {
    S3D_OK = 0,
    S3D_ERR_INVALID = 1,   /* bad argument (null pointer, dimension mismatch: the reference returns 0) */
    S3D_ERR_CUDA = 2,      /* CUDA runtime error (the reference prints GPUassert and exits) */
    S3D_ERR_NOMEM = 3,     /* allocation failed (the reference: "insufficient memory") */
    S3D_ERR_CAPACITY = 4,  /* candidate / keypoint / feature capacity exceeded (the reference overruns
                              its X0*Y0 arrays silently, R/src_common/MultiScale.cpp:255-267) */
    S3D_ERR_UNSUPPORTED = 5
} s3d_status;

typedef struct s3d_ctx s3d_ctx;

/* Feature record, layout-compatible with Feature3DInfo (R/src_common/MultiScale.h:111-129). */
typedef struct s3d_feature {
    unsigned int flag;   /* 0x10 = maximum (INFO_FLAG_MIN0MAX1), 0x20 = reoriented (INFO_FLAG_REORIENT) */
    float x, y, z, scale;
    float ori[9];        /* row-major 3x3 */
    float eigs[3];
    float pc[64];        /* descriptor: ranks 0..63 stored as float, as the reference does */
} s3d_feature;

/* Candidate record = LOCATION_VALUE_XYZ (R/src_common/LocationValue.h:41-47). */
typedef struct s3d_cand {
    int x, y, z;
    float value;
} s3d_cand;

/* A refined keypoint before orientation assignment (octave coordinates, +0.5 applied;
 * R/src_common/MultiScale.cpp:1372-1386). */
typedef struct s3d_keypoint {
    int octave, level, is_max;
    int ix, iy, iz;
    float x, y, z, scale;
} s3d_keypoint;

enum { S3D_DESC_SIFT = 0, S3D_DESC_BRIEF = 1, S3D_DESC_RRIEF = 2, S3D_DESC_NRRIEF = 3 };

/* Options of one extraction = featExtract's command line (R/featExtract/featExtract.cpp:299-350,
 * README -b/-br/-bn). */
typedef struct s3d_params {
    int double_mode;     /* 0, +1 (-2+ : double the input, features scaled by 1/2), -1 (-2- : halve) */
    int descriptor;      /* S3D_DESC_* ; featExtract default is S3D_DESC_SIFT (brief=0, featExtract.cpp:474) */
    float eig_thres;     /* structure-tensor test, 140 in featExtract (featExtract.cpp:297); < 0 disables */
    int max_keypoints;   /* 0 = default (16384) */
    int max_features;    /* 0 = default (8 * max_keypoints) */
    int keep_patches;    /* debug: keep the 11^3 patch of every feature row (s3d_get_patches) */
    /* --- octave-run / z-slab mode (multi-GPU single-volume decomposition; all zero = whole-volume mode) ---
     * input_is_g0 : the volume passed in already IS Gaussian level 0 of an octave (sigma 1.6): no pre-step,
     *               no initial blur.  octave_base = index of that octave in the full pyramid (features are
     *               scaled by 2^(octave_base + o)); max_octaves > 0 limits how many octaves are built.
     * slab        : the buffer holds global planes [z_off, z_off + Z) of an octave whose true depth is
     *               z_global; candidates are only taken from global planes [own_z0, own_z1).  Zero padding,
     *               the support-box test and the trilinear clamp use the GLOBAL depth, so with deep enough
     *               halos every owned keypoint is bit-identical to the whole-volume run. */
    int input_is_g0;
    int octave_base;
    int max_octaves;
    int slab;
    int z_off, z_global, own_z0, own_z1;
    int pre_step_done;   /* +1 / -1: the caller already applied -2+ / -2- to the volume (slab mode resizes each
                            slab itself); only the initial-blur sigma and the final size factor follow it */
} s3d_params;

/* ---- context ------------------------------------------------------------------------------------
 * Replaces cudaSetDevice/cudaMalloc/cudaFree scattered through src_common
 * (R/src_common/FeatureIO.cpp:384-421, 527-530) and check_best_device (featExtract.cpp:237-263).
 * A context owns one device, one stream, and every buffer of the pyramid; it is not thread-safe,
 * use one context per host thread (contexts are independent). */
s3d_status s3d_ctx_create(int device, s3d_ctx **ctx);
s3d_status s3d_ctx_create_on_stream(int device, void *cuda_stream, s3d_ctx **ctx);
void s3d_ctx_destroy(s3d_ctx *ctx);
const char *s3d_last_error(const s3d_ctx *ctx);
int s3d_device_count(void);
s3d_status s3d_sync(s3d_ctx *ctx);
void *s3d_stream(s3d_ctx *ctx);

/* ---- Gaussian taps (host arithmetic, as in the reference) --------------------------------------
 * calculate_gaussian_filter_size + generate_gaussian_filter1d + normalisation
 * (R/src_common/GaussianMask.cpp:12-57, 241-265; R/src_common/GaussBlur3D.cpp:1174-1206).
 * Returns the tap count (odd), or -needed if cap is too small. */
int s3d_gaussian_taps(float sigma, float *taps, int cap);

/* ---- stage level: drop-ins for the four dispatch sites ------------------------------------------
 * All pointers are DEVICE pointers. */

/* blur_3d_simpleborders_CUDA_Row_Col_Shared_mem (R/cuda_common/SIFT_cuda_Tools.cuh:69-76,
 * called from R/src_common/GaussBlur3D.cpp:1244).  Separable correlation, zero padding, x then y
 * then z.  d_in is NOT modified (the reference clobbers it).  d_tmp: scratch of the same size.
 * If d_dog != NULL, also writes d_dog = d_in - d_out (fioCudaMultSum fused into the last pass). */
s3d_status s3d_blur3d(s3d_ctx *ctx, const float *d_in, float *d_tmp, float *d_out,
                      int X, int Y, int Z, int pitch, const float *h_taps, int ntaps, float *d_dog);

/* fioCudaMultSum with m = -1 (R/cuda_common/SIFT_cuda_Tools.cuh:213-217, called from
 * R/src_common/FeatureIO.cpp:1942): d_out = d_a - d_b over Z*Y*pitch elements. */
s3d_status s3d_dog(s3d_ctx *ctx, const float *d_a, const float *d_b, float *d_out,
                   int X, int Y, int Z, int pitch);

/* SubSampleInterpolateCuda (SIFT_cuda_Tools.cuh:202-205, called from FeatureIO.cpp:1559):
 * 2x2x2 mean, output dims floor(X/2) etc. with its own pitch. */
s3d_status s3d_subsample2(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch,
                          float *d_out, int out_pitch);

/* detectExtrema4D_test_cuda (SIFT_cuda_Tools.cuh:32-38, called from R/src_common/MultiScale.cpp:1531).
 * Writes candidates in raster order (z, y, x), minima and maxima separately; counts are device ints.
 * Lists are truncated at cap (the counts still report the true totals). */
s3d_status s3d_detect(s3d_ctx *ctx, const float *d_finer, const float *d_centre,
                      int X, int Y, int Z, int pitch,
                      s3d_cand *d_min, int *d_n_min, s3d_cand *d_max, int *d_n_max, int cap);

/* fioDoubleSize / fioSubSample2DCenterPixel, the -2+ / -2- pre-steps
 * (R/src_common/FeatureIO.cpp:2452-2548, 1670-1714; host loops in the reference). */
s3d_status s3d_double_size(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch,
                           float *d_out, int out_pitch);
s3d_status s3d_halve_size(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch,
                          float *d_out, int out_pitch);

/* ---- pipeline level: what featExtract -dN calls --------------------------------------------------
 * Replaces msGeneratePyramidDOG3D_efficient (R/featExtract/featExtract.cpp:409,
 * R/src_common/MultiScale.cpp:236-570) plus the descriptor loop (featExtract.cpp:474-505).
 *
 * s3d_extract        : volume is a dense HOST array (X*Y*Z floats); H2D, compute and D2H of the
 *                      result are all inside the call.  *out is malloc'ed; release with s3d_free.
 * s3d_extract_device : volume is a dense DEVICE array; enqueues the whole path on the stream and
 *                      returns without synchronising; results stay on the device until
 *                      s3d_fetch_features.  This is the form the throughput benchmark times.
 */
s3d_status s3d_extract(s3d_ctx *ctx, const float *h_volume, int X, int Y, int Z,
                       const s3d_params *params, s3d_feature **out, int *n_out);
s3d_status s3d_extract_device(s3d_ctx *ctx, const float *d_volume, int X, int Y, int Z,
                              const s3d_params *params);
/* Stage the next volume from host memory (pinned or pageable) into the context's input slot
 * asynchronously, then run s3d_extract_device on it. */
s3d_status s3d_extract_host_async(s3d_ctx *ctx, const float *h_volume, int X, int Y, int Z,
                                  const s3d_params *params);
/* Typed input (sect. 8(f) N1): h_volume holds X*Y*Z voxels of a NIfTI scalar datatype (the codes of nifti1.h).
 * The reference converts them to float on the host with a plain cast (reg_changeDatatype1,
 * R/featExtract/featExtract.cpp:18-77) before fioReadNifti returns; here the raw voxels cross PCIe as they are
 * (1, 2, 4 or 8 bytes per voxel) and the cast is fused into the device-side re-pitch -- same bits, less traffic. */
typedef enum {
    S3D_DT_UINT8 = 2, S3D_DT_INT16 = 4, S3D_DT_INT32 = 8, S3D_DT_FLOAT32 = 16, S3D_DT_FLOAT64 = 64,
    S3D_DT_INT8 = 256, S3D_DT_UINT16 = 512, S3D_DT_UINT32 = 768
} s3d_dtype;
s3d_status s3d_extract_typed(s3d_ctx *ctx, const void *h_volume, int dtype, int X, int Y, int Z,
                             const s3d_params *params, s3d_feature **out, int *n_out);
s3d_status s3d_extract_typed_async(s3d_ctx *ctx, const void *h_volume, int dtype, int X, int Y, int Z,
                                   const s3d_params *params);
/* Synchronise, then copy the last extraction's feature rows to the host (malloc'ed). */
s3d_status s3d_fetch_features(s3d_ctx *ctx, s3d_feature **out, int *n_out);
/* Synchronise and return only the counts of the last extraction. */
s3d_status s3d_fetch_counts(s3d_ctx *ctx, int *n_keypoints, int *n_features);
/* Device-resident results of the last extraction (valid until the next one). */
s3d_status s3d_result_device(s3d_ctx *ctx, const s3d_feature **d_features, const int **d_n_features);
void s3d_free(void *p);

/* ---- batch level: many volumes on one GPU (BASELINE.json config 4, per GPU) -----------------------
 * The reference's featExtract handles one volume per process run (R/featExtract/featExtract.cpp:206-590);
 * a study of N volumes is N runs.  s3d_batch keeps `n_contexts` extraction contexts (streams + resident
 * pyramid plans) on one device and feeds them round-robin, so the PCIe transfer, the bandwidth-bound
 * pyramid kernels and the latency-bound keypoint kernels of different volumes overlap.  Volumes are
 * independent: results are identical to s3d_extract on each volume, in input order.  Across GPUs the batch
 * is sharded by the caller, one process (or thread) per GPU, with no communication.
 *
 * s3d_batch_extract        : volumes[i] is a dense HOST array (pinned memory from s3d_host_alloc lets the
 *                            H2D copy run asynchronously); rows[i] is malloc'ed (s3d_free), n_rows[i] its count.
 * s3d_batch_extract_device : volumes[i] is a dense DEVICE array; only the counts are brought back
 *                            (n_keypoints / n_rows may be NULL); this is the form bench.py's `value` times. */
typedef struct s3d_batch s3d_batch;
s3d_status s3d_batch_create(int device, int n_contexts, s3d_batch **batch);
void s3d_batch_destroy(s3d_batch *batch);
const char *s3d_batch_last_error(const s3d_batch *batch);
s3d_status s3d_batch_extract(s3d_batch *batch, const float *const *h_volumes, int n_volumes, int X, int Y, int Z,
                             const s3d_params *params, s3d_feature **rows, int *n_rows);
s3d_status s3d_batch_extract_typed(s3d_batch *batch, const void *const *h_volumes, int dtype, int n_volumes, int X, int Y, int Z,
                                   const s3d_params *params, s3d_feature **rows, int *n_rows);
s3d_status s3d_batch_extract_device(s3d_batch *batch, const float *const *d_volumes, int n_volumes, int X, int Y, int Z,
                                    const s3d_params *params, int *n_keypoints, int *n_rows);
/* kernel launches (graph nodes included) per volume of the last batch call */
int s3d_batch_launches_per_volume(s3d_batch *batch);
/* page-locked host memory for volumes (cudaHostAlloc / cudaFreeHost) */
void *s3d_host_alloc(size_t bytes);
void s3d_host_free(void *p);

/* ---- introspection for parity tests --------------------------------------------------------------
 * Pyramid of the last extraction: Gaussian level g (0..5) or DoG level (0..4) of an octave, copied
 * densely (X*Y*Z floats) to host memory.  dims receives X, Y, Z of that octave. */
int s3d_num_octaves(s3d_ctx *ctx);
s3d_status s3d_get_level(s3d_ctx *ctx, int octave, int is_dog, int level, float *h_out, int dims[3]);
/* Refined keypoints (after validation and the bounds test) in output order. */
s3d_status s3d_get_keypoints(s3d_ctx *ctx, s3d_keypoint **out, int *n_out);
/* With params.keep_patches: n_features * 1331 floats, the 11^3 patch of each row as the pyramid
 * stage leaves it (before the descriptor loop's NormalizeData), and the 64 pre-rank values. */
s3d_status s3d_get_patches(s3d_ctx *ctx, float **patches, float **prerank, int *n_out);
/* Copy planes [z0, z1) of a pyramid level of the last extraction into a dense DEVICE buffer
 * (X*Y*(z1-z0) floats, the reference's layout).  Used by the slab orchestration to hand level 3 to the
 * next octave. */
s3d_status s3d_copy_level_device(s3d_ctx *ctx, int octave, int is_dog, int level, int z0, int z1, float *d_dst);
/* Device pointer, row pitch (floats) and dimensions of a pyramid level of the last extraction (valid until the next
 * extraction on this context); dims = X, Y, Z of the buffer (slab mode: the local planes). */
s3d_status s3d_level_device_ptr(s3d_ctx *ctx, int octave, int is_dog, int level, const float **d_ptr, int *pitch, int dims[3]);
/* Per feature row of the last extraction: index of its keypoint (s3d_get_keypoints order). */
s3d_status s3d_get_row_keypoints(s3d_ctx *ctx, int **row_kp, int *n_out);
/* Number of kernel launches (graph nodes included) issued by the last extraction. */
int s3d_last_launch_count(s3d_ctx *ctx);

/* ---- feature file (kept host code) ----------------------------------------------------------------
 * msFeature3DVectorOutputText (R/src_common/MultiScale.h:386-474) with the three comment lines
 * featExtract writes (R/featExtract/featExtract.cpp:542-575). */
s3d_status s3d_write_features_text(const char *path, const s3d_feature *feats, int n, float eig_thres,
                                   int n_comments, const char *const *comments);

/* Binary feature file: msFeature3DVectorOutputBin (R/src_common/MultiScale.h:228-303) -- two text
 * header lines, then per kept feature x, y, z, scale, ori[9], eigs[3] (float32), the info flag (uint32) and the
 * 64 descriptor values as unsigned char.  eig_thres < 0 keeps every row, like the reference's default. */
s3d_status s3d_write_features_bin(const char *path, const s3d_feature *feats, int n, float eig_thres);
/* Text feature file reader: msFeature3DVectorInputText (R/src_common/MultiScale.h:305-384), the reader
 * featMatchMultiple uses (R/featMatchMultiple/featMatchMultiple.cpp:596).  *out is malloc'ed (s3d_free). */
s3d_status s3d_read_features_text(const char *path, s3d_feature **out, int *n_out);

/* ---- input path on the device ----------------------------------------------------
 * Isotropic resampling of an anisotropic volume: out(x,y,z) = trilinear(in, x*rf_x + 0.5, y*rf_y + 0.5, z*rf_z + 0.5),
 * rf = min voxel size / voxel size per axis, output dims = (int)(n * d / d_min).  Replaces the host triple loop of
 * fioReadNifti (R/featExtract/featExtract.cpp:183-199, fioGetPixelTrilinearInterp R/src_common/FeatureIO.cpp:757-850);
 * same bits.  Device arrays, stream-ordered; the matrix bookkeeping (:146-176) stays with the caller. */
s3d_status s3d_resample_iso(s3d_ctx *ctx, const float *d_in, int X, int Y, int Z, int pitch,
                            float *d_out, int nX, int nY, int nZ, int out_pitch, float rf_x, float rf_y, float rf_z);
/* the same on dense HOST arrays: upload, resample, download, synchronise */
s3d_status s3d_resample_iso_host(s3d_ctx *ctx, const float *h_in, int X, int Y, int Z,
                                 float *h_out, int nX, int nY, int nZ, float rf_x, float rf_y, float rf_z);

/* ---- multi-GPU (SURVEY.md section 8(b) last row, 8(e); BASELINE.json configs 4 and 5) ---------------------------
 * One host thread per GPU above the single-GPU entry points, for hosts that are one process -- like the
 * reference's featExtract (R/featExtract/featExtract.cpp:273-585, one volume, one device chosen by
 * check_best_device :237-263).  The reference has no multi-GPU mode and its launchers overflow at 1024^3
 * (R/cuda_common/SIFT_cuda_Tools.cu:187 computes byte counts in int), so there is nothing to match but the rows.
 *   s3d_multi_create        : devices[r] = CUDA device of GPU r (NULL = 0..n_gpus-1; a device may be listed twice,
 *                             which runs two slabs / shards on it); contexts_per_gpu is for batch mode (0 = 4).
 *   s3d_multi_batch_extract : volume i runs on GPU i mod n through an s3d_batch; no communication; rows[i]
 *                             (malloc'ed, s3d_free) and n_rows[i] in input order -- identical to s3d_extract.
 *   s3d_multi_extract_slab  : ONE volume, split into contiguous z slabs (planes aligned to 2^K).  Per octave every
 *                             GPU runs [halo | own | halo] (48 planes of halo), subsamples its own part of level 3
 *                             and pulls the next octave's halos from its two neighbours with cudaMemcpyPeerAsync
 *                             (NVLink P2P); octaves too thin to split collapse onto GPU 0; rows are merged by
 *                             offsets computed from per-(octave, level, min/max) counts.  *out is malloc'ed
 *                             (s3d_free), rows in the reference's order, bit-identical to the whole-volume run.
 *                             The octave-run / slab fields of params must be zero; max_keypoints = 0 sizes the
 *                             capacity from the slab and grows it on S3D_ERR_CAPACITY. */
typedef struct s3d_multi s3d_multi;
s3d_status s3d_multi_create(int n_gpus, const int *devices, int contexts_per_gpu, s3d_multi **multi);
void s3d_multi_destroy(s3d_multi *multi);
const char *s3d_multi_last_error(const s3d_multi *multi);
int s3d_multi_gpu_count(const s3d_multi *multi);
s3d_status s3d_multi_batch_extract(s3d_multi *multi, const float *const *h_volumes, int n_volumes, int X, int Y, int Z,
                                   const s3d_params *params, s3d_feature **rows, int *n_rows);
s3d_status s3d_multi_extract_slab(s3d_multi *multi, const float *h_volume, int X, int Y, int Z, const s3d_params *params,
                                  s3d_feature **out, int *n_out);

/* ---- descriptor matching -----------------------------------------------------------
 * Exact k nearest neighbours (1 <= k <= 16) of every feature of set A among the features of set B on the reference's
 * descriptor distance Feature3DInfo::DistSqrPCs (R/src_common/MultiScale.h:60-73: sequential float sum of squared
 * differences over the 64 descriptor entries).  Replaces the FLANN kd-tree search of the matcher
 * (R/feat_common/featMatchUtilities.cpp:1449-1455 build parameters, :1559 flann_build_index, :1612
 * flann_find_nearest_neighbors_index with g_nn neighbours, sorted) by an exhaustive one: out_idx / out_dist are
 * [nA][k], neighbours in (distance, index) order -- ties go to the lower index -- and the distances are bit for
 * bit DistSqrPCs.  Entries beyond nB neighbours are index -1, distance +inf.
 * s3d_match takes host arrays and synchronises; s3d_match_device takes device arrays and is stream-ordered. */
s3d_status s3d_match(s3d_ctx *ctx, const s3d_feature *h_feats_a, int n_a, const s3d_feature *h_feats_b, int n_b, int k,
                     int *h_out_idx, float *h_out_dist);
s3d_status s3d_match_device(s3d_ctx *ctx, const s3d_feature *d_feats_a, int n_a, const s3d_feature *d_feats_b, int n_b, int k,
                            int *d_out_idx, float *d_out_dist);

#ifdef __cplusplus
}
#endif
#endif /* S3D_H */
