/*
 * sift3d_oracle.c -- plain-C restatement of the reference featExtract hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sift3d_oracle.h).  Never linked into the product.
 * Parity status: PINNED against the reference's own sources compiled into oracle/_ref
 * (tests/test_oracle_vs_ref.py, bit-for-bit) and against tests/golden/.
 *
 * Every function cites the reference lines it restates.
 * R/ = /root/reference/3dsift_cleanup-softVote_App_Weight_SoftMax/
 *
 * Arithmetic notes that matter for bit parity (x86-64, SSE, no FMA contraction):
 *  - float*float and float+float round to float after every operation;
 *  - where the reference mixes a double literal or a double variable into an expression
 *    the operation is done in double and rounded once on assignment -- kept as written;
 *  - sums run in the reference's order (left to right, raster order);
 *  - the reference is C++: exp()/sqrt() on a float argument resolve to the float overloads
 *    (expf/sqrtf), which is what is called here.
 * The structure is NOT the reference's: the pyramid is kept as whole octaves (six Gaussian
 * and five DoG levels) instead of five rotating buffers, which changes no value.
 */
#include "sift3d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PD S3O_PATCH_DIM
#define PV S3O_PATCH_VOX

/* ------------------------------------------------------------------------------------------
 * Gaussian taps: R/src_common/GaussianMask.cpp:12-57 (size), :241-265 (values),
 * R/src_common/GaussBlur3D.cpp:1190-1201 (normalisation).  fMinValue is always 0.01.
 * ---------------------------------------------------------------------------------------- */
static int filter_size(float fSigma, float fMinValue)
{
    float fPower = 0.0f;
    float fValue = expf(fPower); /* C++ resolves exp(float) to the float overload */
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

int s3o_gaussian_taps(float sigma, float *taps, int cap)
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

/* ------------------------------------------------------------------------------------------
 * Separable blur: R/src_common/GaussBlur3D.cpp:43-61 (filter_1d), :329-479
 * (blur_3d_simpleborders).  Zero padding, no border renormalisation, taps accumulated left to
 * right from fSum = 0, axis order x, y, z with a float round trip between passes.
 * ---------------------------------------------------------------------------------------- */
static void blur_axis(const float *in, float *out, long n_lines_a, long stride_a, long n_lines_b, long stride_b,
                      int len, long stride, const float *taps, int ntaps)
{
    int r = ntaps / 2;
    for (long a = 0; a < n_lines_a; a++) {
        for (long b = 0; b < n_lines_b; b++) {
            const float *src = in + a * stride_a + b * stride_b;
            float *dst = out + a * stride_a + b * stride_b;
            for (int c = 0; c < len; c++) {
                float fSum = 0;
                for (int j = 0; j < ntaps; j++) {
                    int p = c + j - r;
                    float v = (p >= 0 && p < len) ? src[p * stride] : 0.0f;
                    fSum += taps[j] * v;
                }
                dst[c * stride] = fSum;
            }
        }
    }
}

void s3o_blur3d_taps(const float *in, float *out, int X, int Y, int Z, const float *taps, int ntaps)
{
    long n = (long)X * Y * Z;
    float *t1 = (float *)malloc(sizeof(float) * n);
    float *t2 = (float *)malloc(sizeof(float) * n);
    blur_axis(in, t1, Z, (long)X * Y, Y, X, X, 1, taps, ntaps);         /* x */
    blur_axis(t1, t2, Z, (long)X * Y, X, 1, Y, X, taps, ntaps);         /* y */
    blur_axis(t2, out, Y, X, X, 1, Z, (long)X * Y, taps, ntaps);        /* z */
    free(t1);
    free(t2);
}

/* gb3d_blur3d_interleave, R/src_common/GaussBlur3D.cpp:1159-1258 (CPU branch). */
int s3o_blur3d(const float *in, float *out, int X, int Y, int Z, float sigma)
{
    float taps[S3O_MAX_TAPS];
    int n = s3o_gaussian_taps(sigma, taps, S3O_MAX_TAPS);
    if (n < 0) return 0;
    s3o_blur3d_taps(in, out, X, Y, Z, taps, n);
    return 1;
}

/* fioMultSum with fMultIn2 = -1.0f, R/src_common/FeatureIO.cpp:1950-1987. */
void s3o_dog(const float *a, const float *b, float *out, long n)
{
    const float m = -1.0f;
    for (long i = 0; i < n; i++) out[i] = a[i] + m * b[i];
}

/* fioSubSampleInterpolate, R/src_common/FeatureIO.cpp:1474-1554: 2x2x2 mean, dims floor(/2). */
void s3o_subsample(const float *in, float *out, int X, int Y, int Z)
{
    int ox = X / 2, oy = Y / 2, oz = Z / 2;
#define IN_(x, y, z) in[((long)(z) * Y + (y)) * X + (x)]
    for (int z = 0; z < oz; z++)
        for (int y = 0; y < oy; y++)
            for (int x = 0; x < ox; x++) {
                float fSum = 0;
                fSum += IN_(2 * x, 2 * y, 2 * z) + IN_(2 * x, 2 * y + 1, 2 * z) + IN_(2 * x + 1, 2 * y, 2 * z) + IN_(2 * x + 1, 2 * y + 1, 2 * z);
                if (2 * z + 1 < Z) {
                    fSum += IN_(2 * x, 2 * y, 2 * z + 1) + IN_(2 * x, 2 * y + 1, 2 * z + 1) + IN_(2 * x + 1, 2 * y, 2 * z + 1) + IN_(2 * x + 1, 2 * y + 1, 2 * z + 1);
                    fSum *= 0.125;
                } else {
                    fSum *= 0.25;
                }
                out[((long)z * oy + y) * ox + x] = fSum;
            }
#undef IN_
}

/* fioDoubleSize, R/src_common/FeatureIO.cpp:2452-2548 (-2+): out dims 2X,2Y,2Z. */
void s3o_double_size(const float *in, float *out, int X, int Y, int Z)
{
    int DX = X > 1 ? 2 * X : X, DY = Y > 1 ? 2 * Y : Y, DZ = Z > 1 ? 2 * Z : Z;
    for (int z = 0; z < Z; z++)
        for (int y = 0; y < Y; y++)
            for (int x = 0; x < X; x++) {
                float lo[2][2][2], hi[2][2][2];
                for (int zz = 0; zz <= 1; zz++) {
                    int dz = zz; if (z + zz >= Z) dz = 0;
                    for (int yy = 0; yy <= 1; yy++) {
                        int dy = yy; if (y + yy >= Y) dy = 0;
                        for (int xx = 0; xx <= 1; xx++) {
                            int dx = xx; if (x + xx >= X) dx = 0;
                            lo[zz][yy][xx] = in[((long)(z + dz) * Y + (y + dy)) * X + (x + dx)];
                        }
                    }
                }
                hi[0][0][0] = lo[0][0][0];
                hi[1][0][0] = 0.5f * (lo[0][0][0] + lo[1][0][0]);
                hi[0][1][0] = 0.5f * (lo[0][0][0] + lo[0][1][0]);
                hi[0][0][1] = 0.5f * (lo[0][0][0] + lo[0][0][1]);
                hi[1][1][0] = 0.25f * (lo[0][0][0] + lo[1][0][0] + lo[0][1][0] + lo[1][1][0]);
                hi[0][1][1] = 0.25f * (lo[0][0][0] + lo[0][1][0] + lo[0][0][1] + lo[0][1][1]);
                hi[1][0][1] = 0.25f * (lo[0][0][0] + lo[1][0][0] + lo[0][0][1] + lo[1][0][1]);
                hi[1][1][1] = 0.125f * (lo[0][0][0] + lo[0][0][1] + lo[0][1][0] + lo[0][1][1]
                                        + lo[1][0][0] + lo[1][0][1] + lo[1][1][0] + lo[1][1][1]);
                for (int zz = 0; zz <= 1; zz++) {
                    int dz = zz; if (2 * z + zz >= DZ) dz = 0;
                    for (int yy = 0; yy <= 1; yy++) {
                        int dy = yy; if (2 * y + yy >= DY) dy = 0;
                        for (int xx = 0; xx <= 1; xx++) {
                            int dx = xx; if (2 * x + xx >= DX) dx = 0;
                            out[((long)(2 * z + dz) * DY + (2 * y + dy)) * DX + (2 * x + dx)] = hi[dz][dy][dx];
                        }
                    }
                }
            }
}

/* fioSubSample2DCenterPixel as used by -2-, R/src_common/FeatureIO.cpp:1670-1714. */
void s3o_halve_size(const float *in, float *out, int X, int Y, int Z)
{
    int ox = X / 2, oy = Y / 2, oz = Z / 2;
#define IN_(x, y, z) in[((long)(z) * Y + (y)) * X + (x)]
    for (int z = 0; z < oz; z++)
        for (int y = 0; y < oy; y++)
            for (int x = 0; x < ox; x++) {
                float v = 0;
                v += IN_(2 * x + 0, 2 * y + 0, 2 * z + 0);
                v += IN_(2 * x + 0, 2 * y + 0, 2 * z + 1);
                v += IN_(2 * x + 0, 2 * y + 1, 2 * z + 0);
                v += IN_(2 * x + 0, 2 * y + 1, 2 * z + 1);
                v += IN_(2 * x + 1, 2 * y + 0, 2 * z + 0);
                v += IN_(2 * x + 1, 2 * y + 0, 2 * z + 1);
                v += IN_(2 * x + 1, 2 * y + 1, 2 * z + 0);
                v += IN_(2 * x + 1, 2 * y + 1, 2 * z + 1);
                out[((long)z * oy + y) * ox + x] = v / 8.0f;
            }
#undef IN_
}

/* ------------------------------------------------------------------------------------------
 * Detection: R/src_common/MultiScale.cpp:1548-1569 (detectExtrema4D_test), :2260-2391
 * (regFindFEATUREIO), :2400-2524 (valley/peakFunction4D).  A voxel of the centre DoG is kept
 * when it is strictly above (below) its 26 neighbours and all 27 voxels of the finer DoG.
 * Interior voxels only, raster order, minima and maxima in separate lists.
 * ---------------------------------------------------------------------------------------- */
static void neighbour_offsets(int X, int Y, long off[26])
{
    int k = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (dx == 0 && dy == 0 && dz == 0) continue;
                off[k++] = (long)dz * X * Y + (long)dy * X + dx;
            }
}

void s3o_detect(const float *finer, const float *centre, int X, int Y, int Z,
                s3o_cand *mins, int *n_min, s3o_cand *maxs, int *n_max, int cap)
{
    long off[26];
    neighbour_offsets(X, Y, off);
    int nmin = 0, nmax = 0;
    for (int z = 1; z < Z - 1; z++)
        for (int y = 1; y < Y - 1; y++)
            for (int x = 1; x < X - 1; x++) {
                long i = ((long)z * Y + y) * X + x;
                float c = centre[i];
                int peak = 1, valley = 1;
                for (int n = 0; n < 26 && (peak || valley); n++) {
                    float v = centre[i + off[n]];
                    peak &= (v < c);
                    valley &= (v > c);
                }
                if (peak) {
                    peak &= (finer[i] < c);
                    for (int n = 0; n < 26 && peak; n++) peak &= (finer[i + off[n]] < c);
                    if (peak) {
                        if (nmax < cap) { maxs[nmax].x = x; maxs[nmax].y = y; maxs[nmax].z = z; maxs[nmax].value = c; }
                        nmax++;
                    }
                }
                if (valley) {
                    valley &= (finer[i] > c);
                    for (int n = 0; n < 26 && valley; n++) valley &= (finer[i + off[n]] > c);
                    if (valley) {
                        if (nmin < cap) { mins[nmin].x = x; mins[nmin].y = y; mins[nmin].z = z; mins[nmin].value = c; }
                        nmin++;
                    }
                }
            }
    *n_min = nmin;
    *n_max = nmax;
}

/* Deferred validation against the coarser DoG, R/src_common/MultiScale.cpp:1135-1318:
 * the candidate value must strictly beat all 27 voxels of L = G_a - G_b (float subtraction). */
static int validate_cand(const s3o_cand *c, const float *coarser, int X, int Y, int is_max)
{
    long off[26];
    neighbour_offsets(X, Y, off);
    long i = ((long)c->z * Y + c->y) * X + c->x;
    float v = coarser[i];
    int ok = is_max ? (v < c->value) : (v > c->value);
    for (int n = 0; n < 26 && ok; n++) {
        v = coarser[i + off[n]];
        ok &= is_max ? (v < c->value) : (v > c->value);
    }
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Octave schedule: R/src_common/MultiScale.cpp:337-371, 526-527.
 * ---------------------------------------------------------------------------------------- */
void s3o_octave_levels(const float *g0, int X, int Y, int Z, float **g, float **d, float *sigmas)
{
    long n = (long)X * Y * Z;
    float fSigma = 1.6f;
    float fSigmaFactor = (float)pow(2.0, 1.0 / (double)3);
    memcpy(g[0], g0, sizeof(float) * n);
    sigmas[0] = fSigma;
    for (int j = 1; j < 6; j++) {
        float fSigmaExtra = fSigma * sqrtf(fSigmaFactor * fSigmaFactor - 1.0f);
        s3o_blur3d(g[j - 1], g[j], X, Y, Z, fSigmaExtra);
        s3o_dog(g[j - 1], g[j], d[j - 1], n);
        fSigma *= fSigmaFactor;
        sigmas[j] = fSigma;
    }
}

/* ------------------------------------------------------------------------------------------
 * Sub-voxel refinement: R/src_common/MultiScale.cpp:1614-1697, 2531-2534.
 * ---------------------------------------------------------------------------------------- */
static double finddet(double a1, double a2, double a3, double b1, double b2, double b3, double c1, double c2, double c3)
{
    return ((a1 * b2 * c3) - (a1 * b3 * c2) - (a2 * b1 * c3) + (a3 * b1 * c2) + (a2 * b3 * c1) - (a3 * b2 * c1));
}

static double interp_quadratic(double x0, double x1, double x2, double fx0, double fx1, double fx2)
{
    if (!(fx1 < fx0 && fx1 < fx2) && !(fx1 > fx0 && fx1 > fx2)) return x1;
    double a1 = x0 * x0, b1 = x0, c1 = 1;
    double a2 = x1 * x1, b2 = x1, c2 = 1;
    double a3 = x2 * x2, b3 = x2, c3 = 1;
    double d1 = fx0, d2 = fx1, d3 = fx2;
    double det = finddet(a1, a2, a3, b1, b2, b3, c1, c2, c3);
    double detx = finddet(d1, d2, d3, b1, b2, b3, c1, c2, c3);
    double dety = finddet(a1, a2, a3, d1, d2, d3, c1, c2, c3);
    if (d1 == 0 && d2 == 0 && d3 == 0) return x1; /* both "all zero" branches fall through to x1 */
    if (det != 0) {
        if (detx != 0) return dety / (-2.0 * detx);
    }
    return x1;
}

static void interp_point(const float *img, int X, int Y, int ix, int iy, int iz, float *fx, float *fy, float *fz)
{
#define P_(x, y, z) img[((long)(z) * Y + (y)) * X + (x)]
    *fx = (float)interp_quadratic(ix - 1, ix, ix + 1, P_(ix - 1, iy, iz), P_(ix, iy, iz), P_(ix + 1, iy, iz));
    *fy = (float)interp_quadratic(iy - 1, iy, iy + 1, P_(ix, iy - 1, iz), P_(ix, iy, iz), P_(ix, iy + 1, iz));
    *fz = (float)interp_quadratic(iz - 1, iz, iz + 1, P_(ix, iy, iz - 1), P_(ix, iy, iz), P_(ix, iy, iz + 1));
#undef P_
}

/* ------------------------------------------------------------------------------------------
 * Trilinear read / splat: R/src_common/FeatureIO.cpp:757-782, 811-850, 853-889.
 * Pixel centres sit at +0.5; weight is that of the LOWER sample.
 * ---------------------------------------------------------------------------------------- */
static void interp_coord(float fX, float fMinX, float fMaxX, int *iX, float *fW)
{
    if (fX < fMinX + 0.5f) {
        *iX = (int)fMinX;
        *fW = 1.0f;
    } else if (fX >= fMaxX - 0.5f) {
        *iX = (int)(fMaxX - 2);
        *fW = 0.0f;
    } else {
        float fMinusHalf = fX - 0.5f;
        *iX = (int)floor(fMinusHalf);
        *fW = 1.0f - (fMinusHalf - ((float)*iX));
    }
}

static float trilinear_get(const float *img, int X, int Y, int Z, float x, float y, float z)
{
    int iX, iY, iZ;
    float wx, wy, wz;
    interp_coord(x, 0, (float)X, &iX, &wx);
    interp_coord(y, 0, (float)Y, &iY, &wy);
    interp_coord(z, 0, (float)Z, &iZ, &wz);
#define P_(a, b, c) img[((long)(iZ + (c)) * Y + (iY + (b))) * X + (iX + (a))]
    float f000 = P_(0, 0, 0), f100 = P_(1, 0, 0), f010 = P_(0, 1, 0), f110 = P_(1, 1, 0);
    float f001 = P_(0, 0, 1), f101 = P_(1, 0, 1), f011 = P_(0, 1, 1), f111 = P_(1, 1, 1);
#undef P_
    float fn00 = wx * f000 + (1.0f - wx) * f100;
    float fn01 = wx * f001 + (1.0f - wx) * f101;
    float fn10 = wx * f010 + (1.0f - wx) * f110;
    float fn11 = wx * f011 + (1.0f - wx) * f111;
    float fnn0 = wy * fn00 + (1.0f - wy) * fn10;
    float fnn1 = wy * fn01 + (1.0f - wy) * fn11;
    return wz * fnn0 + (1.0f - wz) * fnn1;
}

/* fioIncPixelTrilinearInterp on a dim^3 image with nfeat interleaved features. */
static void trilinear_inc(float *img, int dim, int nfeat, float x, float y, float z, int feat, float v)
{
    int iX, iY, iZ;
    float wx, wy, wz;
    interp_coord(x, 0, (float)dim, &iX, &wx);
    interp_coord(y, 0, (float)dim, &iY, &wy);
    interp_coord(z, 0, (float)dim, &iZ, &wz);
#define Q_(a, b, c) img[((((long)(iZ + (c)) * dim + (iY + (b))) * dim + (iX + (a))) * nfeat) + feat]
    Q_(0, 0, 0) += v * wx * wy * wz;
    Q_(1, 0, 0) += v * (1.0f - wx) * wy * wz;
    Q_(0, 1, 0) += v * wx * (1.0f - wy) * wz;
    Q_(1, 1, 0) += v * (1.0f - wx) * (1.0f - wy) * wz;
    Q_(0, 0, 1) += v * wx * wy * (1.0f - wz);
    Q_(1, 0, 1) += v * (1.0f - wx) * wy * (1.0f - wz);
    Q_(0, 1, 1) += v * wx * (1.0f - wy) * (1.0f - wz);
    Q_(1, 1, 1) += v * (1.0f - wx) * (1.0f - wy) * (1.0f - wz);
#undef Q_
}

/* ------------------------------------------------------------------------------------------
 * Patch gather: R/src_common/MultiScale.cpp:2614-2714 (sampleImage3D),
 * R/src_common/MultiScale.h:192-222 (invert_3x3<float,double>), :494-510 (mult_3x3).
 * Returns 0 on success, -1 when the support box leaves the volume.
 * ---------------------------------------------------------------------------------------- */
static void invert3(const float m[9], float o[9])
{
    float a11 = m[0], a12 = m[1], a13 = m[2];
    float a21 = m[3], a22 = m[4], a23 = m[5];
    float a31 = m[6], a32 = m[7], a33 = m[8];
    float det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
    double div = 1 / (double)det;
    o[0] = (float)((a33 * a22 - a32 * a23) * div);
    o[3] = (float)(-(a33 * a21 - a31 * a23) * div);
    o[6] = (float)((a32 * a21 - a31 * a22) * div);
    o[1] = (float)(-(a33 * a12 - a32 * a13) * div);
    o[4] = (float)((a33 * a11 - a31 * a13) * div);
    o[7] = (float)(-(a32 * a11 - a31 * a12) * div);
    o[2] = (float)((a23 * a12 - a22 * a13) * div);
    o[5] = (float)(-(a23 * a11 - a21 * a13) * div);
    o[8] = (float)((a22 * a11 - a21 * a12) * div);
}

int s3o_sample_patch(const float *img, int X, int Y, int Z, float fx, float fy, float fz, float scale,
                     const float ori[9], float *patch)
{
    float fImageRad = 2.0f * scale;
    int iRadMax = (int)(fImageRad + 2);
    if (fx - iRadMax < 0 || fy - iRadMax < 0 || fz - iRadMax < 0 ||
        fx + iRadMax >= X || fy + iRadMax >= Y || fz + iRadMax >= Z)
        return -1;
    float inv[9];
    invert3(ori, inv);
    int rad = PD / 2;
    for (int z = -rad; z <= rad; z++)
        for (int y = -rad; y <= rad; y++)
            for (int x = -rad; x <= rad; x++) {
                float f[3] = { (float)x, (float)y, (float)z }, p[3];
                for (int i = 0; i < 3; i++) {
                    p[i] = 0;
                    for (int j = 0; j < 3; j++) p[i] += inv[i * 3 + j] * f[j];
                }
                float fScale = fImageRad / (float)(rad);
                p[0] *= fScale; p[1] *= fScale; p[2] *= fScale;
                p[0] += fx; p[1] += fy; p[2] += fz;
                float pix;
                if (p[0] < 0 || p[0] >= X) pix = 0; /* only x is tested, MultiScale.cpp:2687-2689 */
                else pix = trilinear_get(img, X, Y, Z, p[0], p[1], p[2]);
                patch[((z + rad) * PD + (y + rad)) * PD + (x + rad)] = pix;
            }
    return 0;
}

/* Feature3D::NormalizeData, R/src_common/MultiScale.cpp:127-205. */
void s3o_normalize_patch(float *p)
{
    float fSum = 0;
    for (int i = 0; i < PV; i++) fSum += p[i];
    float fMean = fSum / (PD * PD * PD);
    float fSumSqr = 0;
    for (int i = 0; i < PV; i++) {
        p[i] -= fMean;
        fSumSqr += p[i] * p[i];
    }
    float fDiv = 1.0f / sqrtf(fSumSqr);
    for (int i = 0; i < PV; i++) p[i] *= fDiv;
}

/* fioGenerateEdgeImages3D on an 11^3 patch, R/src_common/FeatureIO.cpp:2284-2326. */
static void patch_gradients(const float *p, float *dx, float *dy, float *dz)
{
    memset(dx, 0, sizeof(float) * PV);
    memset(dy, 0, sizeof(float) * PV);
    memset(dz, 0, sizeof(float) * PV);
    for (int z = 1; z < PD - 1; z++)
        for (int y = 1; y < PD - 1; y++)
            for (int x = 1; x < PD - 1; x++) {
                int i = (z * PD + y) * PD + x;
                dx[i] = p[i + 1] - p[i - 1];
                dy[i] = p[i + PD] - p[i - PD];
                dz[i] = p[i + PD * PD] - p[i - PD * PD];
            }
}

/* ------------------------------------------------------------------------------------------
 * 3x3 SVD + sort: R/src_common/SVD.h:15-31 (SortEigenDecomp), :44-228
 * (SingularValueDecomp<float,3,3>, the Numerical Recipes svdcmp with float storage and double
 * scalars).  Expression types follow the C++ template instantiation with T = float.
 * ---------------------------------------------------------------------------------------- */
#define SIGN_(a, b) ((b) >= 0.0 ? fabs(a) : -fabs(a))
#define PYTHAG_(a, b) (sqrt((a) * (a) + (b) * (b)))

static void svd3(float mat[3][3], float w[3], float v[3][3])
{
    const int m = 3, n = 3;
    int flag, i, its, j, jj, k, l = 0, nm = 0;
    double anorm, c, f, g, h, s, scale, x, y, z;
    double rv1[3];
    g = scale = anorm = 0.0;
    for (i = 1; i <= n; i++) {
        l = i + 1;
        rv1[i - 1] = scale * g;
        g = s = scale = 0.0;
        if (i <= m) {
            for (k = i; k <= m; k++) scale += fabs(mat[k - 1][i - 1]);
            if (scale) {
                for (k = i; k <= m; k++) {
                    mat[k - 1][i - 1] = (float)(mat[k - 1][i - 1] / scale);
                    s += (float)(mat[k - 1][i - 1] * mat[k - 1][i - 1]);
                }
                f = mat[i - 1][i - 1];
                g = -SIGN_(sqrt(s), f);
                h = f * g - s;
                mat[i - 1][i - 1] = (float)(f - g);
                for (j = l; j <= n; j++) {
                    for (s = 0.0, k = i; k <= m; k++) s += (float)(mat[k - 1][i - 1] * mat[k - 1][j - 1]);
                    f = s / h;
                    for (k = i; k <= m; k++) mat[k - 1][j - 1] = (float)(mat[k - 1][j - 1] + f * mat[k - 1][i - 1]);
                }
                for (k = i; k <= m; k++) mat[k - 1][i - 1] = (float)(mat[k - 1][i - 1] * scale);
            }
        }
        w[i - 1] = (float)(scale * g);
        g = s = scale = 0.0;
        if (i <= m && i != n) {
            for (k = l; k <= n; k++) scale += fabs(mat[i - 1][k - 1]);
            if (scale) {
                for (k = l; k <= n; k++) {
                    mat[i - 1][k - 1] = (float)(mat[i - 1][k - 1] / scale);
                    s += (float)(mat[i - 1][k - 1] * mat[i - 1][k - 1]);
                }
                f = mat[i - 1][l - 1];
                g = -SIGN_(sqrt(s), f);
                h = f * g - s;
                mat[i - 1][l - 1] = (float)(f - g);
                for (k = l; k <= n; k++) rv1[k - 1] = mat[i - 1][k - 1] / h;
                for (j = l; j <= m; j++) {
                    for (s = 0.0, k = l; k <= n; k++) s += (float)(mat[j - 1][k - 1] * mat[i - 1][k - 1]);
                    for (k = l; k <= n; k++) mat[j - 1][k - 1] = (float)(mat[j - 1][k - 1] + s * rv1[k - 1]);
                }
                for (k = l; k <= n; k++) mat[i - 1][k - 1] = (float)(mat[i - 1][k - 1] * scale);
            }
        }
        {
            double t = fabs(w[i - 1]) + fabs(rv1[i - 1]);
            anorm = (anorm > t ? anorm : t);
        }
    }
    for (i = n; i >= 1; i--) {
        if (i < n) {
            if (g) {
                for (j = l; j <= n; j++) v[j - 1][i - 1] = (float)((mat[i - 1][j - 1] / mat[i - 1][l - 1]) / g);
                for (j = l; j <= n; j++) {
                    for (s = 0.0, k = l; k <= n; k++) s += (float)(mat[i - 1][k - 1] * v[k - 1][j - 1]);
                    for (k = l; k <= n; k++) v[k - 1][j - 1] = (float)(v[k - 1][j - 1] + s * v[k - 1][i - 1]);
                }
            }
            for (j = l; j <= n; j++) v[i - 1][j - 1] = v[j - 1][i - 1] = 0.0f;
        }
        v[i - 1][i - 1] = 1.0f;
        g = rv1[i - 1];
        l = i;
    }
    for (i = (m < n ? m : n); i >= 1; i--) {
        l = i + 1;
        g = w[i - 1];
        for (j = l; j <= n; j++) mat[i - 1][j - 1] = 0.0f;
        if (g) {
            g = 1.0 / g;
            for (j = l; j <= n; j++) {
                for (s = 0.0, k = l; k <= m; k++) s += (float)(mat[k - 1][i - 1] * mat[k - 1][j - 1]);
                f = (s / mat[i - 1][i - 1]) * g;
                for (k = i; k <= m; k++) mat[k - 1][j - 1] = (float)(mat[k - 1][j - 1] + f * mat[k - 1][i - 1]);
            }
            for (j = i; j <= m; j++) mat[j - 1][i - 1] = (float)(mat[j - 1][i - 1] * g);
        } else {
            for (j = i; j <= m; j++) mat[j - 1][i - 1] = 0.0f;
        }
        mat[i - 1][i - 1] = mat[i - 1][i - 1] + 1;
    }
    for (k = n; k >= 1; k--) {
        for (its = 1; its <= 30; its++) {
            flag = 1;
            for (l = k; l >= 1; l--) {
                nm = l - 1;
                if ((double)(fabs(rv1[l - 1]) + anorm) == anorm) { flag = 0; break; }
                if ((double)(fabs(w[nm - 1]) + anorm) == anorm) break;
            }
            if (flag) {
                c = 0.0;
                s = 1.0;
                for (i = l; i <= k; i++) {
                    f = s * rv1[i - 1];
                    rv1[i - 1] = c * rv1[i - 1];
                    if ((double)(fabs(f) + anorm) == anorm) break;
                    g = w[i - 1];
                    h = PYTHAG_(f, g);
                    w[i - 1] = (float)h;
                    h = 1.0 / h;
                    c = g * h;
                    s = -f * h;
                    for (j = 1; j <= m; j++) {
                        y = mat[j - 1][nm - 1];
                        z = mat[j - 1][i - 1];
                        mat[j - 1][nm - 1] = (float)(y * c + z * s);
                        mat[j - 1][i - 1] = (float)(z * c - y * s);
                    }
                }
            }
            z = w[k - 1];
            if (l == k) {
                if (z < 0.0) {
                    w[k - 1] = (float)(-z);
                    for (j = 1; j <= n; j++) v[j - 1][k - 1] = -v[j - 1][k - 1];
                }
                break;
            }
            x = w[l - 1];
            nm = k - 1;
            y = w[nm - 1];
            g = rv1[nm - 1];
            h = rv1[k - 1];
            f = ((y - z) * (y + z) + (g - h) * (g + h)) / (2.0 * h * y);
            g = PYTHAG_(f, 1.0);
            f = ((x - z) * (x + z) + h * ((y / (f + SIGN_(g, f))) - h)) / x;
            c = s = 1.0;
            for (j = l; j <= nm; j++) {
                i = j + 1;
                g = rv1[i - 1];
                y = w[i - 1];
                h = s * g;
                g = c * g;
                z = PYTHAG_(f, h);
                rv1[j - 1] = z;
                c = f / z;
                s = h / z;
                f = x * c + g * s;
                g = g * c - x * s;
                h = y * s;
                y *= c;
                for (jj = 1; jj <= n; jj++) {
                    x = v[jj - 1][j - 1];
                    z = v[jj - 1][i - 1];
                    v[jj - 1][j - 1] = (float)(x * c + z * s);
                    v[jj - 1][i - 1] = (float)(z * c - x * s);
                }
                z = PYTHAG_(f, h);
                w[j - 1] = (float)z;
                if (z) {
                    z = 1.0 / z;
                    c = f * z;
                    s = h * z;
                }
                f = c * g + s * y;
                x = c * y - s * g;
                for (jj = 1; jj <= m; jj++) {
                    y = mat[jj - 1][j - 1];
                    z = mat[jj - 1][i - 1];
                    mat[jj - 1][j - 1] = (float)(y * c + z * s);
                    mat[jj - 1][i - 1] = (float)(z * c - y * s);
                }
            }
            rv1[l - 1] = 0.0;
            rv1[k - 1] = f;
            w[k - 1] = (float)x;
        }
    }
}

static void sort_eigen(float w[3], float v[3][3])
{
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (w[i] < w[j]) {
                float t = w[j]; w[j] = w[i]; w[i] = t;
                for (int k = 0; k < 3; k++) { t = v[k][j]; v[k][j] = v[k][i]; v[k][i] = t; }
            }
}

/* determineOrientation3D, R/src_common/MultiScale.cpp:2541-2607. */
void s3o_eigen_orientation(const float *patch, float eigs[3], float ori[9])
{
    float dx[PV], dy[PV], dz[PV];
    patch_gradients(patch, dx, dy, dz);
    float fMat[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    float fRadiusSqr = (float)((PD / 2) * (PD / 2));
    for (int zz = 0; zz < PD; zz++)
        for (int yy = 0; yy < PD; yy++)
            for (int xx = 0; xx < PD; xx++) {
                float ddz = (float)(zz - PD / 2), ddy = (float)(yy - PD / 2), ddx = (float)(xx - PD / 2);
                if (ddz * ddz + ddy * ddy + ddx * ddx < fRadiusSqr) {
                    int i = (zz * PD + yy) * PD + xx;
                    float e[3] = { dx[i], dy[i], dz[i] };
                    for (int a = 0; a < 3; a++)
                        for (int b = 0; b < 3; b++) fMat[a][b] += e[a] * e[b];
                }
            }
    float v[3][3];
    svd3(fMat, eigs, v);
    sort_eigen(eigs, v);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) ori[a * 3 + b] = v[a][b];
}

/* ------------------------------------------------------------------------------------------
 * Canonical orientations: R/src_common/MultiScale.cpp:2722-3037 (+ vec3D_* :1066-1105,
 * 3040-3049; regFindFEATUREIOPeaks :1987-2121; lvSortHighLow R/src_common/LocationValue.cpp:28-56).
 * ---------------------------------------------------------------------------------------- */
static void vec_norm(float *p)
{
    float fSumSqr = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (fSumSqr > 0) {
        float fDiv = (float)(1.0 / sqrtf(fSumSqr));
        p[0] *= fDiv; p[1] *= fDiv; p[2] *= fDiv;
    } else {
        p[0] = 1; p[1] = 0; p[2] = 0;
    }
}

static float vec_mag(const float *p)
{
    float fSumSqr = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    return fSumSqr > 0 ? sqrtf(fSumSqr) : 0;
}

static float vec_dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static int cmp_high_low(const void *a, const void *b)
{
    const s3o_cand *p = (const s3o_cand *)a, *q = (const s3o_cand *)b;
    if (p->value < q->value) return 1;
    if (p->value > q->value) return -1;
    return 0;
}

static int patch_peaks(const float *h, s3o_cand *peaks)
{
    long off[26];
    neighbour_offsets(PD, PD, off);
    int n = 0;
    for (int z = 1; z < PD - 1; z++)
        for (int y = 1; y < PD - 1; y++)
            for (int x = 1; x < PD - 1; x++) {
                int i = (z * PD + y) * PD + x;
                float c = h[i];
                int peak = 1;
                for (int k = 0; k < 26 && peak; k++) peak &= (h[i + off[k]] < c);
                if (peak) { peaks[n].x = x; peaks[n].y = y; peaks[n].z = z; peaks[n].value = c; n++; }
            }
    qsort(peaks, n, sizeof(s3o_cand), cmp_high_low);
    return n;
}

int s3o_canonical_orientations(const float *patch, float *rots, int max_ori)
{
    float dx[PV], dy[PV], dz[PV], t0[PV], t2[PV];
    s3o_cand peaks[128], peaks2[128];
    float oriData[128 * 3];
    float fRadius = (float)(PD / 2);
    float fRadiusSqr = (float)((PD / 2) * (PD / 2));
    float p1[3], p2[3], p3[3];

    patch_gradients(patch, dx, dy, dz);
    memset(t0, 0, sizeof(t0));
    for (int zz = 0; zz < PD; zz++)
        for (int yy = 0; yy < PD; yy++)
            for (int xx = 0; xx < PD; xx++) {
                float ddz = (float)(zz - PD / 2), ddy = (float)(yy - PD / 2), ddx = (float)(xx - PD / 2);
                if (ddz * ddz + ddy * ddy + ddx * ddx < fRadiusSqr) {
                    int i = (zz * PD + yy) * PD + xx;
                    float e[3] = { dx[i], dy[i], dz[i] };
                    float fEdgeMagSqr = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
                    if (fEdgeMagSqr == 0) continue;
                    float fEdgeMag = sqrtf(fEdgeMagSqr);
                    float u[3];
                    for (int k = 0; k < 3; k++) u[k] = e[k] * fRadius / fEdgeMag;
                    for (int k = 0; k < 3; k++) u[k] += fRadius;
                    trilinear_inc(t0, PD, 1, (float)(u[0] + 0.5), (float)(u[1] + 0.5), (float)(u[2] + 0.5), 0, fEdgeMag);
                }
            }
    s3o_blur3d(t0, t2, PD, PD, PD, 0.5f);
    int np = patch_peaks(t2, peaks);

    for (int i = 0; i < np && i < PD && i < max_ori; i++) {
        float *o = &oriData[i * 3];
        interp_point(t2, PD, PD, peaks[i].x, peaks[i].y, peaks[i].z, &o[0], &o[1], &o[2]);
        o[0] -= fRadius; o[1] -= fRadius; o[2] -= fRadius;
        vec_norm(o);
    }

    int nret = 0;
    for (int i = 0; i < np && i < PD && nret < max_ori; i++) {
        if (peaks[i].value < 0.8 * peaks[0].value) break;
        p1[0] = oriData[i * 3]; p1[1] = oriData[i * 3 + 1]; p1[2] = oriData[i * 3 + 2];

        memset(t0, 0, sizeof(t0));
        for (int zz = 0; zz < PD; zz++)
            for (int yy = 0; yy < PD; yy++)
                for (int xx = 0; xx < PD; xx++) {
                    float ddx = (float)(xx - PD / 2), ddy = (float)(yy - PD / 2), ddz = (float)(zz - PD / 2);
                    if (ddz * ddz + ddy * ddy + ddx * ddx < fRadiusSqr) {
                        int ii = (zz * PD + yy) * PD + xx;
                        float e[3] = { dx[ii], dy[ii], dz[ii] };
                        float fEdgeMag = vec_mag(e);
                        if (fEdgeMag == 0) continue;
                        float u[3] = { e[0], e[1], e[2] };
                        vec_norm(u);
                        float perp[3];
                        float fPar = vec_dot(p1, u);
                        perp[0] = u[0] - fPar * p1[0];
                        perp[1] = u[1] - fPar * p1[1];
                        perp[2] = u[2] - fPar * p1[2];
                        vec_norm(perp);
                        for (int k = 0; k < 3; k++) { perp[k] *= fRadius; perp[k] += fRadius; }
                        trilinear_inc(t0, PD, 1, (float)(perp[0] + 0.5), (float)(perp[1] + 0.5), (float)(perp[2] + 0.5), 0, fEdgeMag);
                    }
                }
        s3o_blur3d(t0, t2, PD, PD, PD, 0.5f);
        int np2 = patch_peaks(t2, peaks2);
        for (int j = 0; j < np2 && nret < PD && nret < max_ori; j++) {
            if (peaks2[j].value < 0.5f * peaks2[0].value) break;
            interp_point(t2, PD, PD, peaks2[j].x, peaks2[j].y, peaks2[j].z, &p2[0], &p2[1], &p2[2]);
            p2[0] -= fRadius; p2[1] -= fRadius; p2[2] -= fRadius;
            vec_norm(p2);
            float fPar = vec_dot(p1, p2);
            p2[0] = p2[0] - fPar * p1[0];
            p2[1] = p2[1] - fPar * p1[1];
            p2[2] = p2[2] - fPar * p1[2];
            vec_norm(p2);
            p3[0] = p1[1] * p2[2] - p1[2] * p2[1];
            p3[1] = -p1[0] * p2[2] + p1[2] * p2[0];
            p3[2] = p1[0] * p2[1] - p1[1] * p2[0];
            float *m = rots + 9 * nret;
            for (int k = 0; k < 3; k++) { m[k] = p1[k]; m[3 + k] = p2[k]; m[6 + k] = p3[k]; }
            nret++;
        }
    }
    return nret;
}

/* ------------------------------------------------------------------------------------------
 * Descriptors.
 * ---------------------------------------------------------------------------------------- */
/* msNormalizeDataPositive, R/src_common/MultiScale.cpp:1580-1611. */
static void normalize_positive(float *v, int n)
{
    float fMin = 100000;
    for (int i = 0; i < n; i++) if (v[i] < fMin) fMin = v[i];
    float fSumSqr = 0;
    for (int i = 0; i < n; i++) { v[i] -= fMin; fSumSqr += v[i] * v[i]; }
    float fDiv = 1.0f / sqrtf(fSumSqr);
    for (int i = 0; i < n; i++) v[i] *= fDiv;
}

/* msResampleFeaturesGradientOrientationHistogram, R/src_common/MultiScale.cpp:583-710. */
void s3o_descriptor_sift(const float *patch, float pc[S3O_NPC])
{
    static const float oriAngles[8][3] = {
        { 1, 1, 1 }, { 1, 1, -1 }, { 1, -1, 1 }, { 1, -1, -1 }, { -1, 1, 1 }, { -1, 1, -1 }, { -1, -1, 1 }, { -1, -1, -1 },
    };
    float dx[PV], dy[PV], dz[PV];
    patch_gradients(patch, dx, dy, dz);
    float fBinSize = PD / (float)2;
    float coord[PD];
    for (int q = 0; q < PD; q++) {
        float c = (int)(q / fBinSize) + 0.5f;
        if ((int)((q + 0) / fBinSize) != (int)((q + 1) / fBinSize)) {
            float fP0 = ((q + 0) / fBinSize);
            float fP1 = ((q + 1) / fBinSize);
            c = (fP0 + fP1) / 2.0f;
        }
        coord[q] = c;
    }
    memset(pc, 0, sizeof(float) * S3O_NPC);
    for (int zz = 0; zz < PD; zz++)
        for (int yy = 0; yy < PD; yy++)
            for (int xx = 0; xx < PD; xx++) {
                int i = (zz * PD + yy) * PD + xx;
                float e[3] = { dx[i], dy[i], dz[i] };
                float fEdgeMag = vec_mag(e);
                if (fEdgeMag > 0) {
                    vec_norm(e);
                    int iMax = 0;
                    float fMaxDot = vec_dot(oriAngles[0], e);
                    for (int k = 1; k < 8; k++) {
                        float fDot = vec_dot(oriAngles[k], e);
                        if (fDot > fMaxDot) { fMaxDot = fDot; iMax = k; }
                    }
                    trilinear_inc(pc, 2, 8, coord[xx], coord[yy], coord[zz], iMax, fEdgeMag);
                }
            }
    normalize_positive(pc, S3O_NPC);
}

/* BRIEF pair table, method 2: R/src_common/MultiScale.cpp:805-807 (x,y,z triples). */
static const unsigned char brief_a[192] = { 5,4,4,4,4,2,6,5,5,4,4,4,3,8,5,5,6,3,5,5,5,5,6,5,4,6,6,6,3,4,4,4,5,3,4,5,4,5,5,4,2,7,7,5,3,5,4,5,3,5,7,3,5,5,2,3,5,5,6,6,4,6,5,4,4,6,5,3,5,6,4,3,6,4,4,5,3,3,3,6,6,5,2,4,4,6,3,6,3,2,3,5,4,5,3,4,3,6,5,4,3,6,4,5,2,4,3,7,2,3,6,5,2,6,3,3,5,6,3,6,3,5,3,6,5,7,4,2,5,5,5,2,5,7,4,2,5,3,4,3,3,7,4,4,7,6,4,4,2,8,7,6,5,4,7,3,6,6,5,2,4,5,3,2,5,5,1,6,3,6,3,6,2,5,4,4,7,2,6,3,2,2,4,3,3,2,3,4,2,5,6,7 };
static const unsigned char brief_b[192] = { 6,5,3,4,5,3,7,4,6,4,3,2,4,7,5,3,5,1,5,4,7,6,8,4,4,5,6,5,2,5,4,6,4,0,4,3,3,4,4,2,1,7,8,6,4,4,1,6,1,3,7,2,3,3,1,3,6,1,6,6,4,7,6,4,3,5,4,2,3,6,4,5,6,3,3,5,1,3,1,6,7,4,1,4,3,5,2,4,2,1,2,5,4,5,2,3,3,3,3,4,2,6,3,4,3,3,3,6,1,2,5,4,2,4,1,4,6,7,3,6,2,4,3,6,5,6,4,0,6,6,5,1,4,7,2,1,5,3,4,2,2,7,3,3,6,4,2,4,1,9,7,7,5,2,7,1,7,5,5,1,5,4,1,3,3,4,0,5,1,6,3,5,3,2,3,3,7,2,5,1,1,0,4,1,3,1,0,3,1,6,5,9 };

/* msResampleFeaturesBRIEF, R/src_common/MultiScale.cpp:989-1049 (RRIEF is the live line; BRIEF and
 * NRRIEF are the commented alternatives :1037-1045); blur run with CPU semantics. */
void s3o_descriptor_brief(const float *patch, int mode, float pc[S3O_NPC])
{
    float blurred[PV];
    s3o_blur3d(patch, blurred, PD, PD, PD, 0.95f);
    for (int i = 0; i < S3O_NPC; i++) {
        int ax = brief_a[3 * i], ay = brief_a[3 * i + 1], az = brief_a[3 * i + 2];
        int bx = brief_b[3 * i], by = brief_b[3 * i + 1], bz = brief_b[3 * i + 2];
        float d = blurred[ax + ay * PD + az * PD * PD] - blurred[bx + by * PD + bz * PD * PD];
        if (mode == S3O_DESC_BRIEF) pc[i] = d < 0;
        else if (mode == S3O_DESC_RRIEF) pc[i] = d;
        else {
            float fdx = (float)(ax - bx), fdy = (float)(ay - by), fdz = (float)(az - bz);
            int e = (int)sqrtf(fdx * fdx + fdy * fdy + fdz * fdz); /* euclidean_distance_3d :1051-1056 */
            pc[i] = d / e;
        }
    }
}

/* NormalizeDataRankedPCs, R/src_common/MultiScale.cpp:207-233, ties by index :3148-3176. */
void s3o_rank(float pc[S3O_NPC])
{
    float r[S3O_NPC];
    for (int i = 0; i < S3O_NPC; i++) {
        int k = 0;
        for (int j = 0; j < S3O_NPC; j++)
            if (pc[j] < pc[i] || (pc[j] == pc[i] && j < i)) k++;
        r[i] = (float)k;
    }
    memcpy(pc, r, sizeof(r));
}

/* ------------------------------------------------------------------------------------------
 * Whole path: R/featExtract/featExtract.cpp:366-505 + R/src_common/MultiScale.cpp:236-570
 * (msGeneratePyramidDOG3D_efficient), :1326-1424 (generateFeatures3D_efficient),
 * :1705-1862 (generateFeature3D).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    s3o_feature *f;
    float *patch;
    int n, cap;
} featvec;

static void fv_push(featvec *v, const s3o_feature *f, const float *patch)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? 2 * v->cap : 256;
        v->f = (s3o_feature *)realloc(v->f, sizeof(s3o_feature) * v->cap);
        v->patch = (float *)realloc(v->patch, sizeof(float) * PV * v->cap);
    }
    v->f[v->n] = *f;
    memcpy(v->patch + (size_t)PV * v->n, patch, sizeof(float) * PV);
    v->n++;
}

static void generate_feature(featvec *out, float fx, float fy, float fz, float scale, int is_max,
                             const float *img, int X, int Y, int Z, float fEigThres)
{
    s3o_feature f;
    float patch[PV];
    memset(&f, 0, sizeof(f));
    f.x = fx; f.y = fy; f.z = fz; f.scale = scale;
    f.flag = is_max ? S3O_FLAG_MAX : 0;
    f.ori[0] = 1; f.ori[4] = 1; f.ori[8] = 1;
    if (s3o_sample_patch(img, X, Y, Z, fx, fy, fz, scale, f.ori, patch) != 0) return;
    s3o_normalize_patch(patch);
    s3o_eigen_orientation(patch, f.eigs, f.ori);
    float fEigSum = f.eigs[0] + f.eigs[1] + f.eigs[2];
    float fEigPrd = f.eigs[0] * f.eigs[1] * f.eigs[2];
    float fEigSumProd = fEigSum * fEigSum * fEigSum;
    if (!(fEigSumProd < fEigThres * fEigPrd || fEigThres < 0)) return;
    fv_push(out, &f, patch);

    float rots[30 * 9];
    int nori = s3o_canonical_orientations(patch, rots, 30);
    for (int o = 0; o < nori; o++) {
        memcpy(f.ori, rots + 9 * o, sizeof(float) * 9);
        if (s3o_sample_patch(img, X, Y, Z, fx, fy, fz, scale, f.ori, patch) != 0) continue;
        f.flag |= S3O_FLAG_REORIENT;
        fv_push(out, &f, patch);
    }
}

int s3o_extract(const float *vol, int X, int Y, int Z, int double_mode, int descriptor,
                s3o_feature **feats, float **patches, float **prerank,
                s3o_keypoint **keypoints, int *n_keypoints)
{
    float fInitialImageScale = 1.0f;
    float *img = NULL;
    if (double_mode == 1) {
        img = (float *)malloc(sizeof(float) * 8 * (size_t)X * Y * Z);
        s3o_double_size(vol, img, X, Y, Z);
        X *= 2; Y *= 2; Z *= 2;
        fInitialImageScale *= 0.5;
    } else if (double_mode == -1) {
        img = (float *)malloc(sizeof(float) * (size_t)(X / 2) * (Y / 2) * (Z / 2) + 16);
        s3o_halve_size(vol, img, X, Y, Z);
        X /= 2; Y /= 2; Z /= 2;
    } else {
        img = (float *)malloc(sizeof(float) * (size_t)X * Y * Z);
        memcpy(img, vol, sizeof(float) * (size_t)X * Y * Z);
    }
    const int X0 = X, Y0 = Y;
    const float fEigThres = 140;
    size_t n0 = (size_t)X * Y * Z;
    float *g[6], *d[5], sig[6];
    for (int j = 0; j < 6; j++) g[j] = (float *)malloc(sizeof(float) * n0);
    for (int j = 0; j < 5; j++) d[j] = (float *)malloc(sizeof(float) * n0);
    float *g0 = (float *)malloc(sizeof(float) * n0);
    int cap = X0 * Y0;
    s3o_cand *mins = (s3o_cand *)malloc(sizeof(s3o_cand) * cap), *maxs = (s3o_cand *)malloc(sizeof(s3o_cand) * cap);

    featvec fv = { 0, 0, 0, 0 };
    s3o_keypoint *kps = NULL;
    int nkp = 0, kpcap = 0;

    /* initial blur, MultiScale.cpp:288-298 */
    float fSigmaInit = 0.5f;
    if (fInitialImageScale > 0) fSigmaInit /= fInitialImageScale;
    float fSigma = 1.6f;
    float fSigmaExtra = sqrtf(fSigma * fSigma - fSigmaInit * fSigmaInit);
    s3o_blur3d(img, g0, X, Y, Z, fSigmaExtra);

    float fScale = 1;
    for (int oct = 0;; oct++) {
        if (X <= 2 || Y <= 2 || Z <= 2) break;
        int first = fv.n;
        s3o_octave_levels(g0, X, Y, Z, g, d, sig);
        for (int c = 1; c <= 3; c++) {
            int nmin, nmax;
            s3o_detect(d[c - 1], d[c], X, Y, Z, mins, &nmin, maxs, &nmax, cap);
            if (nmin > cap) nmin = cap;
            if (nmax > cap) nmax = cap;
            for (int pass = 0; pass < 2; pass++) {
                s3o_cand *lst = pass ? maxs : mins;
                int cnt = pass ? nmax : nmin;
                for (int k = 0; k < cnt; k++) {
                    if (!validate_cand(&lst[k], d[c + 1], X, Y, pass)) continue;
                    long vi = ((long)lst[k].z * Y + lst[k].y) * X + lst[k].x;
                    float fx, fy, fz;
                    interp_point(d[c], X, Y, lst[k].x, lst[k].y, lst[k].z, &fx, &fy, &fz);
                    float scale = (float)(2 * interp_quadratic(sig[c - 1], sig[c], sig[c + 1], d[c - 1][vi], d[c][vi], d[c + 1][vi]));
                    fx += 0.5f; fy += 0.5f; fz += 0.5f;
                    /* a keypoint whose support box leaves the volume is dropped (sampleImage3D returns -1,
                     * MultiScale.cpp:2633-2643); only the survivors are reported as keypoints */
                    int iRadMax = (int)(2.0f * scale + 2);
                    int inside = !(fx - iRadMax < 0 || fy - iRadMax < 0 || fz - iRadMax < 0 ||
                                   fx + iRadMax >= X || fy + iRadMax >= Y || fz + iRadMax >= Z);
                    if (keypoints && inside) {
                        if (nkp == kpcap) { kpcap = kpcap ? 2 * kpcap : 256; kps = (s3o_keypoint *)realloc(kps, sizeof(s3o_keypoint) * kpcap); }
                        s3o_keypoint kp = { oct, c, pass, lst[k].x, lst[k].y, lst[k].z, fx, fy, fz, scale };
                        kps[nkp++] = kp;
                    }
                    generate_feature(&fv, fx, fy, fz, scale, pass, g[c], X, Y, Z, fEigThres);
                }
            }
        }
        /* rescale to input-image units, MultiScale.cpp:531-543 */
        for (int i = first; i < fv.n; i++) {
            fv.f[i].scale *= fScale;
            fv.f[i].x = fv.f[i].x * fScale + 0;
            fv.f[i].y = fv.f[i].y * fScale + 0;
            fv.f[i].z = fv.f[i].z * fScale + 0;
        }
        fScale *= 2.0f;
        /* next octave from level 3 (sigma 3.2), MultiScale.cpp:409-417, 546-556 */
        s3o_subsample(g[3], g0, X, Y, Z);
        X /= 2; Y /= 2; Z /= 2;
    }

    int n = fv.n;
    if (patches) {
        *patches = (float *)malloc(sizeof(float) * PV * (size_t)(n ? n : 1));
        memcpy(*patches, fv.patch, sizeof(float) * PV * (size_t)n);
    }
    if (prerank) *prerank = (float *)malloc(sizeof(float) * S3O_NPC * (size_t)(n ? n : 1));

    /* descriptor loop, featExtract.cpp:474-505 */
    float fSizeFactor = 1;
    if (double_mode > 0) fSizeFactor /= 2; else if (double_mode < 0) fSizeFactor *= 2;
    for (int i = 0; i < n; i++) {
        float *p = fv.patch + (size_t)PV * i;
        s3o_normalize_patch(p);
        if (descriptor == S3O_DESC_SIFT) s3o_descriptor_sift(p, fv.f[i].pc);
        else s3o_descriptor_brief(p, descriptor, fv.f[i].pc);
        if (prerank) memcpy(*prerank + (size_t)S3O_NPC * i, fv.f[i].pc, sizeof(float) * S3O_NPC);
        s3o_rank(fv.f[i].pc);
        fv.f[i].x *= fSizeFactor; fv.f[i].y *= fSizeFactor; fv.f[i].z *= fSizeFactor; fv.f[i].scale *= fSizeFactor;
    }
    if (feats) {
        *feats = (s3o_feature *)malloc(sizeof(s3o_feature) * (size_t)(n ? n : 1));
        memcpy(*feats, fv.f, sizeof(s3o_feature) * (size_t)n);
    }
    if (keypoints) { *keypoints = kps; if (n_keypoints) *n_keypoints = nkp; } else free(kps);

    free(fv.f); free(fv.patch);
    for (int j = 0; j < 6; j++) free(g[j]);
    for (int j = 0; j < 5; j++) free(d[j]);
    free(g0); free(img); free(mins); free(maxs);
    return n;
}

void s3o_free(void *p) { free(p); }

/* ---------------------------------------------------------------------------------------------------
 * Descriptor matching (section 8(f) N2).
 * Feature3DInfo::DistSqrPCs (MultiScale.h:60-73): float fSumSqr = 0; for i: fDiff = a[i] - b[i]; fSumSqr += fDiff*fDiff.
 * The reference feeds this metric to FLANN kd-trees (featMatchUtilities.cpp:1449-1455, 1559, 1612: approximate);
 * the restatement is the exhaustive search FLANN approximates, neighbours sorted by (distance, index).
 * --------------------------------------------------------------------------------------------------- */
float s3o_dist_sqr_pcs(const float *a, const float *b)
{
    float fSumSqr = 0;
    for (int i = 0; i < S3O_NPC; i++) {
        float fDiff = a[i] - b[i];
        fSumSqr += fDiff * fDiff;
    }
    return fSumSqr;
}

void s3o_knn(const float *a, int nA, const float *b, int nB, int k, int *idx, float *dist)
{
    for (int q = 0; q < nA; q++) {
        int *bi = idx + (size_t)q * k;
        float *bd = dist + (size_t)q * k;
        for (int s = 0; s < k; s++) { bi[s] = -1; bd[s] = INFINITY; }
        for (int j = 0; j < nB; j++) {
            float d = s3o_dist_sqr_pcs(a + (size_t)q * S3O_NPC, b + (size_t)j * S3O_NPC);
            if (!(d < bd[k - 1])) continue;          /* ties keep the lower index */
            int s = k - 1;
            while (s > 0 && d < bd[s - 1]) { bd[s] = bd[s - 1]; bi[s] = bi[s - 1]; s--; }
            bd[s] = d; bi[s] = j;
        }
    }
}
