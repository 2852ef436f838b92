/*
 * sift3d_oracle.h -- plain-C restatement of the reference featExtract hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (3d_sift_cuda_b200/, include/) may
 * include, link or call this.  It exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have a checker that travels to the GPU box.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py checks every entry point below
 * bit-for-bit against the reference's own sources compiled into oracle/_ref
 * (oracle/Makefile), and tests/golden/ holds vectors minted from that reference build.
 *
 * R/ = /root/reference/3dsift_cleanup-softVote_App_Weight_SoftMax/
 * Volumes: dense fp32, x fastest, idx = (z*Y + y)*X + x   (R/src_common/FeatureIO.cpp:739)
 */
#ifndef SIFT3D_ORACLE_H
#define SIFT3D_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define S3O_PATCH_DIM 11
#define S3O_PATCH_VOX 1331
#define S3O_NPC 64
#define S3O_MAX_TAPS 129
#define S3O_FLAG_MAX 0x10u      /* INFO_FLAG_MIN0MAX1, R/src_common/MultiScale.h:27-30 */
#define S3O_FLAG_REORIENT 0x20u /* INFO_FLAG_REORIENT */

/* Layout-compatible with Feature3DInfo (R/src_common/MultiScale.h:111-129). */
typedef struct s3o_feature {
    unsigned int flag;
    float x, y, z, scale;
    float ori[9];
    float eigs[3];
    float pc[S3O_NPC];
} s3o_feature;

/* Candidate record: LOCATION_VALUE_XYZ (R/src_common/LocationValue.h:41-47). */
typedef struct s3o_cand {
    int x, y, z;
    float value;
} s3o_cand;

/* One refined keypoint before orientation assignment (for stage-level checks). */
typedef struct s3o_keypoint {
    int octave, level, is_max;
    int ix, iy, iz;        /* voxel in octave coordinates */
    float x, y, z, scale;  /* refined, octave coordinates, +0.5 applied */
} s3o_keypoint;

enum { S3O_DESC_SIFT = 0, S3O_DESC_BRIEF = 1, S3O_DESC_RRIEF = 2, S3O_DESC_NRRIEF = 3 };

int  s3o_gaussian_taps(float sigma, float *taps, int cap);
void s3o_blur3d_taps(const float *in, float *out, int X, int Y, int Z, const float *taps, int ntaps);
int  s3o_blur3d(const float *in, float *out, int X, int Y, int Z, float sigma);
void s3o_dog(const float *a, const float *b, float *out, long n);
void s3o_subsample(const float *in, float *out, int X, int Y, int Z);
void s3o_double_size(const float *in, float *out, int X, int Y, int Z);
void s3o_halve_size(const float *in, float *out, int X, int Y, int Z);
void s3o_detect(const float *finer, const float *centre, int X, int Y, int Z,
                s3o_cand *mins, int *n_min, s3o_cand *maxs, int *n_max, int cap);

/* The six Gaussian levels and five DoG levels of one octave from its level-0 image.
 * g[6], d[5]: caller-allocated X*Y*Z floats each; sigmas[6] receives 1.6*k^j. */
void s3o_octave_levels(const float *g0, int X, int Y, int Z, float **g, float **d, float *sigmas);

/* Patch-level stages (11^3 patches, zyx order). */
int  s3o_sample_patch(const float *img, int X, int Y, int Z, float fx, float fy, float fz, float scale,
                      const float ori[9], float *patch);
void s3o_normalize_patch(float *patch);
void s3o_eigen_orientation(const float *patch, float eigs[3], float ori[9]);
int  s3o_canonical_orientations(const float *patch, float *rots /* 30*9 */, int max_ori);
void s3o_descriptor_sift(const float *patch, float pc[S3O_NPC]);
void s3o_descriptor_brief(const float *patch, int mode, float pc[S3O_NPC]);
void s3o_rank(float pc[S3O_NPC]);

/* Whole path: pre-step (-2+ => double_mode 1, -2- => -1), pyramid, detection, refinement,
 * orientation, descriptor.  Outputs are malloc'ed (free with s3o_free); any may be NULL.
 * Returns the number of feature rows (>= 0) or < 0 on error. */
int  s3o_extract(const float *vol, int X, int Y, int Z, int double_mode, int descriptor,
                 s3o_feature **feats, float **patches, float **prerank,
                 s3o_keypoint **keypoints, int *n_keypoints);
void s3o_free(void *p);

/* Descriptor matching: exact k nearest neighbours of every descriptor of set A among
 * set B on Feature3DInfo::DistSqrPCs (MultiScale.h:60-73), neighbours in (distance, index) order; entries beyond
 * nB neighbours are index -1 / distance +inf.  a, b: [n][S3O_NPC] descriptors. */
float s3o_dist_sqr_pcs(const float *a, const float *b);
void s3o_knn(const float *a, int nA, const float *b, int nB, int k, int *idx, float *dist);

#ifdef __cplusplus
}
#endif
#endif
