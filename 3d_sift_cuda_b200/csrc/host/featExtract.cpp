// featExtract -- drop-in replacement for the reference CLI (reference featExtract/featExtract.cpp), with
// the whole extraction running on a B200 through the C-ABI of include/s3d.h.
//
//   featExtract [options] <input image> <output features>
//     -w / -ws   output feature geometry in world coordinates (NIfTI qto_xyz / sto_xyz); implies isotropic
//                processing (reference :327-339, 436-473, 507-538)
//     -2+ / -2-  double / halve the input image (reference :368-388)
//     -d[0-9]    CUDA device (reference :312-325; here every run is a GPU run, default device 0)
//     -b -br -bn BRIEF / RRIEF / NRRIEF descriptor instead of SIFT-Rank (README; dead code in the reference)
//     -r X Y Z   input is raw IEEE float32, little endian, x fastest, of these dimensions
//
// Kept host code: argument parsing, NIfTI/raw loading, isotropic resampling, the world-coordinate
// transform and the text feature file.  Output is byte-identical to the reference's CPU path.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "nifti_min.h"
#include "s3d.h"

using niftimin::Mat44;

static int print_options()
{
    printf("Volumetric local feature extraction v1.1 (B200 engine)\n");
    printf("Usage: %s [options] <input image> <output features>\n", "featExtract");
    printf("  <input image>: nifti (.nii,.hdr,.nii.gz) or, with -r, raw IEEE 32-bit float little endian.\n");
    printf("  <output features>: output file with features.\n");
    printf(" [options]\n");
    printf("  -w         : output feature geometry in world coordinates, NIFTI qto_xyz matrix (default is voxel units).\n");
    printf("  -2+        : double input image size.\n");
    printf("  -2-        : halve input image size.\n");
    printf("  -d[0-9]    : set device id to be used.\n");
    printf("  -b -br -bn : BRIEF / RRIEF / NRRIEF descriptor (default: gradient orientation histogram).\n");
    printf("  -r X Y Z   : raw float32 input of the given dimensions.\n");
    return 0;
}

// ---- small float helpers with the reference's arithmetic (MultiScale.cpp:1058-1105, MultiScale.h:192-222, 512-533)
static float vec3_mag(const float *p)
{
    float s = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    return s > 0 ? sqrtf(s) : 0;
}
static void vec3_norm(float *p)
{
    float s = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    if (s > 0) { float d = (float)(1.0 / sqrtf(s)); p[0] *= d; p[1] *= d; p[2] *= d; }
    else { p[0] = 1; p[1] = 0; p[2] = 0; }
}
static void invert3f(const float in[9], float out[9])   // invert_3x3<float,float>
{
    float a11 = in[0], a12 = in[1], a13 = in[2], a21 = in[3], a22 = in[4], a23 = in[5], a31 = in[6], a32 = in[7], a33 = in[8];
    float det = a11 * (a33 * a22 - a32 * a23) - a21 * (a33 * a12 - a32 * a13) + a31 * (a23 * a12 - a22 * a13);
    float div = 1 / (float)det;
    out[0] = (a33 * a22 - a32 * a23) * div;
    out[3] = -(a33 * a21 - a31 * a23) * div;
    out[6] = (a32 * a21 - a31 * a22) * div;
    out[1] = -(a33 * a12 - a32 * a13) * div;
    out[4] = (a33 * a11 - a31 * a13) * div;
    out[7] = -(a32 * a11 - a31 * a12) * div;
    out[2] = (a23 * a12 - a22 * a13) * div;
    out[5] = -(a23 * a11 - a21 * a13) * div;
    out[8] = (a22 * a11 - a21 * a12) * div;
}
static void mult3f(const float a[9], const float b[9], float o[9])
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            o[i * 3 + j] = 0;
            for (int k = 0; k < 3; k++) o[i * 3 + j] += a[i * 3 + k] * b[k * 3 + j];
        }
}

// _fioDetermineInterpCoord / fioGetPixelTrilinearInterp (reference FeatureIO.cpp:757-850), host copy used
// only for the isotropic resampling of anisotropic inputs (reference featExtract.cpp:118-204).
static void interp_coord(float fX, float fMaxX, int &iX, float &fW)
{
    if (fX < 0.5f) { iX = 0; fW = 1.0f; }
    else if (fX >= fMaxX - 0.5f) { iX = (int)(fMaxX - 2); fW = 0.0f; }
    else { float mh = fX - 0.5f; iX = (int)floor(mh); fW = 1.0f - (mh - ((float)iX)); }
}
static float trilinear(const float *img, int X, int Y, int Z, float x, float y, float z)
{
    int iX, iY, iZ; float wx, wy, wz;
    interp_coord(x, (float)X, iX, wx); interp_coord(y, (float)Y, iY, wy); interp_coord(z, (float)Z, iZ, wz);
    const float *p = img + ((size_t)iZ * Y + iY) * X + iX;
    size_t pl = (size_t)X * Y;
    float fn00 = wx * p[0] + (1.0f - wx) * p[1];
    float fn01 = wx * p[pl] + (1.0f - wx) * p[pl + 1];
    float fn10 = wx * p[X] + (1.0f - wx) * p[X + 1];
    float fn11 = wx * p[pl + X] + (1.0f - wx) * p[pl + X + 1];
    float fnn0 = wy * fn00 + (1.0f - wy) * fn10;
    float fnn1 = wy * fn01 + (1.0f - wy) * fn11;
    return wz * fnn0 + (1.0f - wz) * fnn1;
}

int main(int argc, char **argv)
{
    if (argc < 3) { print_options(); return -1; }
    int device = 0, iArg = 1, bDouble = 0, bWorld = 0, bIso = 0, descriptor = S3D_DESC_SIFT;
    int rawX = 0, rawY = 0, rawZ = 0;
    float fEigThres = 140;
    while (iArg < argc && argv[iArg][0] == '-') {
        switch (argv[iArg][1]) {
        case '2':
            bDouble = 1;
            if (argv[iArg][2] == '-') bDouble = -1;
            iArg++;
            break;
        case 'd': {
            int n = s3d_device_count();
            int d = argv[iArg][2] - '0';
            if (argv[iArg][2] == 0) d = 0;
            if (d < 0 || d > 9 || d >= (n > 0 ? n : 1)) {
                printf("Error: unknown device: %d\n", d);
                print_options();
                return -1;
            }
            device = d;
            iArg++;
            break;
        }
        case 'w': case 'W':
            bWorld = 1; bIso = 1;
            if (argv[iArg][2] == 's' || argv[iArg][2] == 'S') bWorld = 2;
            iArg++;
            break;
        case 'b':
            descriptor = argv[iArg][2] == 'r' ? S3D_DESC_RRIEF : argv[iArg][2] == 'n' ? S3D_DESC_NRRIEF : S3D_DESC_BRIEF;
            iArg++;
            break;
        case 'r':
            if (iArg + 3 >= argc) { print_options(); return -1; }
            rawX = atoi(argv[iArg + 1]); rawY = atoi(argv[iArg + 2]); rawZ = atoi(argv[iArg + 3]);
            iArg += 4;
            break;
        default:
            printf("Error: unknown command line argument: %s\n", argv[iArg]);
            print_options();
            return -1;
        }
    }
    if (argc - iArg < 2) { print_options(); return -1; }
    const char *inPath = argv[iArg], *outPath = argv[iArg + 1];
    printf("Extracting features: %s\n", inPath);

    niftimin::Image im;
    if (rawX > 0) {
        if (rawY <= 0 || rawZ <= 0) { printf("Error: bad raw dimensions\n"); return -1; }
        im.nx = rawX; im.ny = rawY; im.nz = rawZ;
        im.data.resize((size_t)rawX * rawY * rawZ);
        FILE *f = fopen(inPath, "rb");
        if (!f || fread(im.data.data(), sizeof(float), im.data.size(), f) != im.data.size()) {
            printf("Error: could not read input file: %s\n", inPath);
            return -1;
        }
        fclose(f);
        memset(&im.qto_xyz, 0, sizeof(Mat44));
        im.qto_xyz.m[0][0] = im.qto_xyz.m[1][1] = im.qto_xyz.m[2][2] = im.qto_xyz.m[3][3] = 1.0f;
        im.sto_xyz = im.qto_xyz;
    } else if (niftimin::read(inPath, im, /*keep_raw=*/true) < 0) {
        printf("Error: could not read input file: %s\n", inPath);
        return -1;
    }
    // NIfTI voxels stay in their file datatype when they can go to the device as they are (the cast to float
    // then runs there, s3d_extract_typed); the host-side isotropic resampling needs floats
    const bool need_iso = bIso && (im.dx != im.dy || im.dy != im.dz || im.dx != im.dz);
    const bool typed = !im.raw.empty() && !need_iso && im.datatype != 16;
    if (!im.raw.empty() && !typed) {
        im.data.resize((size_t)im.nx * im.ny * im.nz * im.nt);
        niftimin::cast_to_float(im.raw.data(), im.datatype, im.data.size(), im.data.data());
        im.raw.clear();
    }

    int X = im.nx, Y = im.ny, Z = im.nz;
    std::vector<float> vol;
    // isotropic resampling (reference featExtract.cpp:118-204)
    if (bIso && (im.dx != im.dy || im.dy != im.dz || im.dx != im.dz)) {
        float fMin = im.dx;
        if (im.dy < fMin) fMin = im.dy;
        if (im.dz < fMin) fMin = im.dz;
        int nX = (int)(im.nx * im.dx / fMin), nY = (int)(im.ny * im.dy / fMin), nZ = (int)(im.nz * im.dz / fMin);
        float rf[3] = { fMin / im.dx, fMin / im.dy, fMin / im.dz };
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                im.qto_xyz.m[i][j] *= rf[j];
                if (im.sform_code > 0) im.sto_xyz.m[i][j] *= rf[j];
            }
        vol.resize((size_t)nX * nY * nZ);
        for (int z = 0; z < nZ; z++)
            for (int y = 0; y < nY; y++)
                for (int x = 0; x < nX; x++)
                    vol[((size_t)z * nY + y) * nX + x] = trilinear(im.data.data(), X, Y, Z, (float)(x * rf[0] + 0.5), (float)(y * rf[1] + 0.5), (float)(z * rf[2] + 0.5));
        X = nX; Y = nY; Z = nZ;
        im.dx = im.dy = im.dz = fMin;
    } else if (!typed) {
        vol.assign(im.data.begin(), im.data.begin() + (size_t)X * Y * Z);
    }
    int eX = X, eY = Y, eZ = Z;   // extraction resolution (after -2+/-2-)
    if (bDouble == 1) { eX *= 2; eY *= 2; eZ *= 2; }
    else if (bDouble == -1) { eX /= 2; eY /= 2; eZ /= 2; }
    if (eZ <= 1) { printf("Could not read volume: %s\n", inPath); return -1; }
    printf("Input image: i=%d j=%d k=%d\n", eX, eY, eZ);

    s3d_ctx *ctx = nullptr;
    if (s3d_ctx_create(device, &ctx) != S3D_OK) {
        printf("Error: could not initialise CUDA device %d: %s\n", device, ctx ? s3d_last_error(ctx) : "no device");
        return -1;
    }
    s3d_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.double_mode = bDouble; prm.descriptor = descriptor; prm.eig_thres = fEigThres;
    s3d_feature *feats = nullptr;
    int n = 0;
    auto run = [&]() {
        return typed ? s3d_extract_typed(ctx, im.raw.data(), im.datatype, X, Y, Z, &prm, &feats, &n)
                     : s3d_extract(ctx, vol.data(), X, Y, Z, &prm, &feats, &n);
    };
    s3d_status st = run();
    if (st == S3D_ERR_CAPACITY) {   // retry once with room for a very dense volume
        prm.max_keypoints = 1 << 18;
        prm.max_features = 1 << 21;
        st = run();
    }
    if (st != S3D_OK) {
        printf("Error: could not extract features, %s.\n", st == S3D_ERR_NOMEM ? "insufficient memory" : s3d_last_error(ctx));
        s3d_ctx_destroy(ctx);
        return -1;
    }

    // world coordinates (reference featExtract.cpp:436-473, 507-538)
    Mat44 m_current;
    memset(&m_current, 0, sizeof(m_current));
    if (bWorld) {
        const Mat44 *pm = &im.qto_xyz;
        if (bWorld == 2) {
            if (im.sform_code > 0) pm = &im.sto_xyz;
            else printf("Error: sform_code <= 0, output to qto_xyz instead of sto_xyz");
        }
        m_current = *pm;
        float pfScale[3], fScaleSum = 0, rot[9];
        for (int i = 0; i < 3; i++) {
            pfScale[i] = vec3_mag(&pm->m[i][0]);
            fScaleSum += pfScale[i];
            memcpy(&rot[i * 3], &pm->m[i][0], 3 * sizeof(float));
            vec3_norm(&rot[i * 3]);
        }
        fScaleSum /= 3;
        for (int k = 0; k < n; k++) {
            s3d_feature &f = feats[k];
            float in[4] = { f.x, f.y, f.z, 1 }, out[4];
            for (int i = 0; i < 4; i++) {
                out[i] = 0;
                for (int j = 0; j < 4; j++) out[i] += m_current.m[i][j] * in[j];
            }
            f.x = out[0]; f.y = out[1]; f.z = out[2];
            f.scale *= fScaleSum;
            float oi[9], oo[9];
            invert3f(f.ori, oi);
            mult3f(rot, oi, oo);
            invert3f(oo, f.ori);
        }
    }

    char c1[200], c2[200], c3[400];
    snprintf(c1, sizeof(c1), "Extraction Voxel Resolution (ijk) : %d %d %d", eX, eY, eZ);
    snprintf(c2, sizeof(c2), "Extraction Voxel Size (mm)  (ijk) : %f %f %f", 1.0f * im.dx, 1.0f * im.dy, 1.0f * im.dz);
    if (bWorld) {
        snprintf(c3, sizeof(c3), "Feature Coordinate Space: millimeters (%s) : %f %f %f %f %f %f %f %f %f %f %f %f 0.0 0.0 0.0 1.0",
                 bWorld == 1 ? "qto_xyz" : "sto_xyz",
                 1.0f * m_current.m[0][0], 1.0f * m_current.m[0][1], 1.0f * m_current.m[0][2], 1.0f * m_current.m[0][3],
                 1.0f * m_current.m[1][0], 1.0f * m_current.m[1][1], 1.0f * m_current.m[1][2], 1.0f * m_current.m[1][3],
                 1.0f * m_current.m[2][0], 1.0f * m_current.m[2][1], 1.0f * m_current.m[2][2], 1.0f * m_current.m[2][3]);
    } else {
        snprintf(c3, sizeof(c3), "Feature Coordinate Space: voxels: 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0");
    }
    const char *comments[3] = { c1, c2, c3 };
    if (s3d_write_features_text(outPath, feats, n, fEigThres, 3, comments) != S3D_OK) {
        printf("Error: could not write %s\n", outPath);
        return -1;
    }
    s3d_free(feats);
    s3d_ctx_destroy(ctx);
    printf("\nDone.\n");
    return 0;
}
