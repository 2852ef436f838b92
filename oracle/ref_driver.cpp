/*
 * ref_driver.cpp -- C-ABI driver around the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY).
 *
 * Built by oracle/Makefile together with the reference's own src_common/*.cpp (compiled
 * where they lie under /root/reference, never copied) into oracle/_ref/libref3dsift.so.
 * Every entry point calls the reference's own functions; nothing is re-implemented here
 * except the glue the reference keeps inside featExtract.cpp's main() (the per-feature
 * descriptor loop, featExtract.cpp:474-505), which is restated call-for-call.
 *
 * Used only by tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline
 * legs, as the checker and the timed CPU baseline -- never by the product path.
 *
 * Volumes are dense fp32, x fastest: idx = (z*Y + y)*X + x  (FeatureIO.cpp:739).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <utility>
#include <cmath>
#include <iostream>
#include <chrono>
#include <unistd.h>
#include <fcntl.h>

#include "FeatureIO.h"
#include "GaussBlur3D.h"
#include "GaussianMask.h"
#include "LocationValue.h"
#include "MultiScale.h"
#include "PpImage.h"

namespace {

/* The reference chats on stdout ("#<us>" stage timers, "done.") and writes ./image.pgm on
 * every pyramid run (MultiScale.cpp:374-384).  Silence fd 1 and run inside a scratch
 * directory for the duration of a call. */
struct QuietScope {
    int saved_fd;
    char cwd[4096];
    bool moved;
    QuietScope() : saved_fd(-1), moved(false)
    {
        fflush(stdout);
        saved_fd = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
        const char *scratch = getenv("S3D_REF_SCRATCH");
        if (!scratch) scratch = "/tmp";
        if (getcwd(cwd, sizeof(cwd)) && chdir(scratch) == 0) moved = true;
    }
    ~QuietScope()
    {
        fflush(stdout);
        std::cout.flush();
        if (saved_fd >= 0) { dup2(saved_fd, 1); close(saved_fd); }
        if (moved) { if (chdir(cwd) != 0) { /* nothing sensible to do */ } }
    }
};

FEATUREIO make_fio(int x, int y, int z)
{
    FEATUREIO f;
    memset(&f, 0, sizeof(f));
    f.x = x; f.y = y; f.z = z; f.t = 1; f.iFeaturesPerVector = 1;
    f.device = 0; /* what the shipped CPU path ends up with (featExtract.cpp:108-109) */
    fioAllocate(f);
    return f;
}

FEATUREIO make_fio_from(const float *src, int x, int y, int z)
{
    FEATUREIO f = make_fio(x, y, z);
    memcpy(f.pfVectors, src, sizeof(float) * (size_t)x * y * z);
    if (f.d_pfVectors) memcpy(f.d_pfVectors, src, sizeof(float) * (size_t)x * y * z);
    return f;
}

} // namespace

extern "C" {

/* Record layout shared with tests (ctypes): mirrors Feature3DInfo (MultiScale.h:111-129). */
struct ref_feature {
    unsigned int flag;
    float x, y, z, scale;
    float ori[9];
    float eigs[3];
    float pc[64];
};

/* GaussianMask.cpp:12-57 + 241-265 + the normalisation in GaussBlur3D.cpp:1190-1201. */
int ref_gaussian_taps(float sigma, float *taps, int cap)
{
    float fMin = 0.01f;
    int n = calculate_gaussian_filter_size(sigma, fMin);
    if (n > cap) return -n;
    PpImage pp;
    pp.Initialize(1, n, n * sizeof(float), sizeof(float) * 8);
    if (sigma > 0.0f) generate_gaussian_filter1d(pp, sigma, n / 2);
    else *((float *)pp.ImageRow(0)) = 1;
    float *pf = (float *)pp.ImageRow(0);
    float fSum = 0;
    for (int c = 0; c < n; c++) fSum += pf[c];
    for (int c = 0; c < n; c++) pf[c] /= fSum;
    memcpy(taps, pf, n * sizeof(float));
    return n;
}

/* gb3d_blur3d (GaussBlur3D.cpp:1261-1272) on the CPU branch (best_device_id = -1). */
int ref_blur3d(const float *in, float *out, int x, int y, int z, float sigma)
{
    QuietScope q;
    FEATUREIO a = make_fio_from(in, x, y, z), t = make_fio(x, y, z), b = make_fio(x, y, z);
    int r = gb3d_blur3d(a, t, b, sigma, 0.01f, -1);
    memcpy(out, b.pfVectors, sizeof(float) * (size_t)x * y * z);
    fioDelete(a); fioDelete(t); fioDelete(b);
    return r;
}

/* fioMultSum (FeatureIO.cpp:1950-1987) with m = -1: out = a - b. */
int ref_dog(const float *a_, const float *b_, float *out, int x, int y, int z)
{
    FEATUREIO a = make_fio_from(a_, x, y, z), b = make_fio_from(b_, x, y, z), o = make_fio(x, y, z);
    int r = fioMultSum(a, b, o, -1.0f);
    memcpy(out, o.pfVectors, sizeof(float) * (size_t)x * y * z);
    fioDelete(a); fioDelete(b); fioDelete(o);
    return r;
}

/* fioSubSampleInterpolate (FeatureIO.cpp:1474-1554): out dims = floor(in/2). */
int ref_subsample(const float *in, float *out, int x, int y, int z)
{
    FEATUREIO a = make_fio_from(in, x, y, z), o = make_fio(x / 2, y / 2, z / 2);
    int r = fioSubSampleInterpolate(a, o);
    memcpy(out, o.pfVectors, sizeof(float) * (size_t)(x / 2) * (y / 2) * (z / 2));
    fioDelete(a); fioDelete(o);
    return r;
}

/* fioDoubleSize (FeatureIO.cpp:2452-2548): out dims = 2*in. */
int ref_double_size(const float *in, float *out, int x, int y, int z)
{
    FEATUREIO a = make_fio_from(in, x, y, z);
    int r = fioDoubleSize(a);
    memcpy(out, a.pfVectors, sizeof(float) * (size_t)a.x * a.y * a.z);
    fioDelete(a);
    return r;
}

/* fioSubSample2DCenterPixel as called for -2- (featExtract.cpp:377-386): out dims = in/2. */
int ref_halve_size(const float *in, float *out, int x, int y, int z)
{
    FEATUREIO a = make_fio_from(in, x, y, z), o = make_fio(x / 2, y / 2, z / 2);
    int r = fioSubSample2DCenterPixel(a, o);
    memcpy(out, o.pfVectors, sizeof(float) * (size_t)(x / 2) * (y / 2) * (z / 2));
    fioDelete(a); fioDelete(o);
    return r;
}

/* detectExtrema4D_test (MultiScale.cpp:1548-1569) on the CPU branch.
 * xyzv arrays hold {x,y,z} ints and the centre value; returns counts through n_min/n_max. */
int ref_detect(const float *finer, const float *centre, int x, int y, int z,
               int *min_xyz, float *min_val, int *n_min,
               int *max_xyz, float *max_val, int *n_max, int cap)
{
    FEATUREIO h = make_fio_from(finer, x, y, z), c = make_fio_from(centre, x, y, z);
    std::vector<LOCATION_VALUE_XYZ> vmin((size_t)x * y * z / 8 + 64), vmax((size_t)x * y * z / 8 + 64);
    LOCATION_VALUE_XYZ_ARRAY amin, amax;
    memset(&amin, 0, sizeof(amin)); memset(&amax, 0, sizeof(amax));
    amin.plvz = vmin.data(); amax.plvz = vmax.data();
    detectExtrema4D_test(&h, &c, 0, amin, amax);
    *n_min = amin.iCount; *n_max = amax.iCount;
    for (int i = 0; i < amin.iCount && i < cap; i++) {
        min_xyz[3 * i] = vmin[i].x; min_xyz[3 * i + 1] = vmin[i].y; min_xyz[3 * i + 2] = vmin[i].z; min_val[i] = vmin[i].fValue;
    }
    for (int i = 0; i < amax.iCount && i < cap; i++) {
        max_xyz[3 * i] = vmax[i].x; max_xyz[3 * i + 1] = vmax[i].y; max_xyz[3 * i + 2] = vmax[i].z; max_val[i] = vmax[i].fValue;
    }
    fioDelete(h); fioDelete(c);
    return 1;
}

/* Whole path.  Restates featExtract.cpp main(): -2+/-2- pre-step (:366-388), pyramid (:409),
 * descriptor loop (:474-505).  descriptor: 0 = SIFT-Rank (the shipped default, brief=0),
 * 1 = BRIEF, 2 = RRIEF, 3 = NRRIEF -- the three variants restate MultiScale.cpp:1032-1045
 * with the blur forced onto the CPU branch (the shipped line passes device 0 and cannot run).
 *
 * Outputs (all optional, malloc'ed, release with ref_free):
 *   *feats      n ref_feature records in output-file order, geometry scaled as main() does;
 *   *patches    n*1331 floats: data_zyx as returned by the pyramid (before main's NormalizeData);
 *   *prerank    n*64 floats: m_pfPC before NormalizeDataRankedPCs.
 * Returns n >= 0, or < 0 on failure.  seconds (optional) = wall time of pyramid + descriptor loop. */
int ref_extract(const float *vol, int x, int y, int z, int double_mode, int descriptor,
                ref_feature **feats, float **patches, float **prerank, double *seconds)
{
    QuietScope q;
    FEATUREIO fioIn = make_fio_from(vol, x, y, z);
    float fInitialBlurScale = 1.0f;
    if (double_mode == 1) {
        fioDoubleSize(fioIn);
        fInitialBlurScale *= 0.5;
    } else if (double_mode == -1) {
        FEATUREIO fioTmp = fioIn;
        fioIn.x /= 2; fioIn.y /= 2; fioIn.z /= 2;
        fioAllocate(fioIn);
        fioSubSample2DCenterPixel(fioTmp, fioIn);
        fioDelete(fioTmp);
    }
    float fEigThres = 140;
    std::vector<Feature3D> vec;
    auto t0 = std::chrono::high_resolution_clock::now();
    int rc = msGeneratePyramidDOG3D_efficient(fioIn, vec, -1, fInitialBlurScale, 0, 0, fEigThres);
    if (rc != 1) { fioDelete(fioIn); return -1; }

    int n = (int)vec.size();
    if (patches) {
        *patches = (float *)malloc(sizeof(float) * 1331 * (size_t)(n ? n : 1));
        for (int i = 0; i < n; i++) memcpy(*patches + (size_t)i * 1331, &vec[i].data_zyx[0][0][0], sizeof(float) * 1331);
    }
    if (prerank) *prerank = (float *)malloc(sizeof(float) * 64 * (size_t)(n ? n : 1));

    float fSizeFactor = 1;
    if (double_mode > 0) fSizeFactor /= 2; else if (double_mode < 0) fSizeFactor *= 2;

    std::vector<LOCATION_VALUE_XYZ> ix, iy;
    msGenerateBRIEFindex(ix, iy, 64, fioIn);
    for (int i = 0; i < n; i++) {
        vec[i].NormalizeData();
        if (descriptor == 0) {
            msResampleFeaturesGradientOrientationHistogram(vec[i]);
        } else {
            /* MultiScale.cpp:1032-1045 with the CPU blur */
            FEATUREIO a, b, c;
            memset(&a, 0, sizeof(a));
            a.x = a.y = a.z = Feature3D::FEATURE_3D_DIM; a.t = 1; a.iFeaturesPerVector = 1; a.device = -1;
            b = a; c = a;
            Feature3D fT0, fT1;
            a.pfVectors = &(vec[i].data_zyx[0][0][0]);
            b.pfVectors = &(fT0.data_zyx[0][0][0]);
            c.pfVectors = &(fT1.data_zyx[0][0][0]);
            gb3d_blur3d(a, c, b, 0.95, 0.01, -1);
            for (int k = 0; k < 64; k++) {
                float d = b.pfVectors[ix[k].x + ix[k].y * a.x + ix[k].z * a.x * a.y]
                        - b.pfVectors[iy[k].x + iy[k].y * a.x + iy[k].z * a.x * a.y];
                if (descriptor == 1) vec[i].m_pfPC[k] = d < 0;
                else if (descriptor == 2) vec[i].m_pfPC[k] = d;
                else {
                    int e = euclidean_distance_3d(ix[k].x, iy[k].x, ix[k].y, iy[k].y, ix[k].z, iy[k].z);
                    vec[i].m_pfPC[k] = d / e;
                }
            }
        }
        if (prerank) memcpy(*prerank + (size_t)i * 64, vec[i].m_pfPC, sizeof(float) * 64);
        vec[i].NormalizeDataRankedPCs();
        vec[i].x *= fSizeFactor; vec[i].y *= fSizeFactor; vec[i].z *= fSizeFactor; vec[i].scale *= fSizeFactor;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    if (feats) {
        *feats = (ref_feature *)malloc(sizeof(ref_feature) * (size_t)(n ? n : 1));
        for (int i = 0; i < n; i++) {
            ref_feature &o = (*feats)[i];
            o.flag = vec[i].m_uiInfo;
            o.x = vec[i].x; o.y = vec[i].y; o.z = vec[i].z; o.scale = vec[i].scale;
            memcpy(o.ori, &vec[i].ori[0][0], sizeof(o.ori));
            memcpy(o.eigs, vec[i].eigs, sizeof(o.eigs));
            memcpy(o.pc, vec[i].m_pfPC, sizeof(o.pc));
        }
    }
    fioDelete(fioIn);
    return n;
}

/* msFeature3DVectorOutputText (MultiScale.h:386-474) on records produced by anyone. */
int ref_write_text(const ref_feature *feats, int n, const char *path, int x, int y, int z)
{
    std::vector<Feature3D> vec(n);
    for (int i = 0; i < n; i++) {
        vec[i].m_uiInfo = feats[i].flag;
        vec[i].x = feats[i].x; vec[i].y = feats[i].y; vec[i].z = feats[i].z; vec[i].scale = feats[i].scale;
        memcpy(&vec[i].ori[0][0], feats[i].ori, sizeof(feats[i].ori));
        memcpy(vec[i].eigs, feats[i].eigs, sizeof(feats[i].eigs));
        memcpy(vec[i].m_pfPC, feats[i].pc, sizeof(feats[i].pc));
    }
    /* the three comment lines as main() formats them for voxel coordinates (featExtract.cpp:542-571) */
    char c1[200], c2[200], c3[400];
    sprintf(c1, "Extraction Voxel Resolution (ijk) : %d %d %d", x, y, z);
    sprintf(c2, "Extraction Voxel Size (mm)  (ijk) : %f %f %f", 1.0f, 1.0f, 1.0f);
    sprintf(c3, "Feature Coordinate Space: voxels: 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0");
    char *cc[3] = { c1, c2, c3 };
    return msFeature3DVectorOutputText(vec, (char *)path, 140.0f, 3, cc);
}

/* msFeature3DVectorOutputBin (MultiScale.h:228-303) on records produced by anyone. */
int ref_write_bin(const ref_feature *feats, int n, const char *path, float eig_thres)
{
    std::vector<Feature3D> vec(n);
    for (int i = 0; i < n; i++) {
        vec[i].m_uiInfo = feats[i].flag;
        vec[i].x = feats[i].x; vec[i].y = feats[i].y; vec[i].z = feats[i].z; vec[i].scale = feats[i].scale;
        memcpy(&vec[i].ori[0][0], feats[i].ori, sizeof(feats[i].ori));
        memcpy(vec[i].eigs, feats[i].eigs, sizeof(feats[i].eigs));
        memcpy(vec[i].m_pfPC, feats[i].pc, sizeof(feats[i].pc));
    }
    return msFeature3DVectorOutputBin(vec, (char *)path, eig_thres);
}

/* msFeature3DVectorInputText (MultiScale.h:305-384): the reader featMatchMultiple uses. */
int ref_read_text(const char *path, ref_feature **feats)
{
    std::vector<Feature3D> vec;
    if (msFeature3DVectorInputText(vec, (char *)path, 140.0f) < 0) return -1;
    int n = (int)vec.size();
    *feats = (ref_feature *)malloc(sizeof(ref_feature) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        ref_feature &o = (*feats)[i];
        o.flag = vec[i].m_uiInfo;
        o.x = vec[i].x; o.y = vec[i].y; o.z = vec[i].z; o.scale = vec[i].scale;
        memcpy(o.ori, &vec[i].ori[0][0], sizeof(o.ori));
        memcpy(o.eigs, vec[i].eigs, sizeof(o.eigs));
        memcpy(o.pc, vec[i].m_pfPC, sizeof(o.pc));
    }
    return n;
}

/* Exhaustive k nearest neighbours on the reference's own distance function Feature3DInfo::DistSqrPCs
 * (MultiScale.h:60-73), the metric the reference hands to FLANN (featMatchUtilities.cpp:1612).  Neighbours in
 * (distance, index) order; a, b: ref_feature records. */
int ref_knn(const ref_feature *a, int nA, const ref_feature *b, int nB, int k, int *idx, float *dist)
{
    std::vector<Feature3DInfo> fa((size_t)nA), fb((size_t)nB);
    for (int i = 0; i < nA; i++) memcpy(fa[i].m_pfPC, a[i].pc, sizeof(float) * 64);
    for (int i = 0; i < nB; i++) memcpy(fb[i].m_pfPC, b[i].pc, sizeof(float) * 64);
    for (int q = 0; q < nA; q++) {
        std::vector<std::pair<float, int> > all((size_t)nB);
        for (int j = 0; j < nB; j++) all[j] = std::make_pair(fa[q].DistSqrPCs(fb[j], 64), j);
        std::stable_sort(all.begin(), all.end(), [](const std::pair<float, int> &x, const std::pair<float, int> &y) { return x.first < y.first; });
        for (int s = 0; s < k; s++) {
            idx[(size_t)q * k + s] = s < nB ? all[s].second : -1;
            dist[(size_t)q * k + s] = s < nB ? all[s].first : INFINITY;
        }
    }
    return 0;
}

void ref_free(void *p) { free(p); }

} // extern "C"
