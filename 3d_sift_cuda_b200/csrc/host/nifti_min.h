// nifti_min.h -- minimal NIfTI-1 reader for the featExtract CLI (host I/O, kept on the CPU).
// Covers what the reference's fioReadNifti needs (reference featExtract.cpp:84-220): single-file .nii,
// .nii.gz and .hdr/.img pairs, the eight scalar datatypes the reference converts
// (reg_changeDatatype, featExtract.cpp:18-77; raw values, no scl_slope), either byte order, and the
// qform/sform matrices of the NIfTI-1 standard.  Written from the published NIfTI-1 header layout
// (offsets below); gz transparently via zlib's gzread.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

namespace niftimin {

struct Mat44 { float m[4][4]; };

struct Image {
    int nx = 0, ny = 0, nz = 0, nt = 1;
    float dx = 1, dy = 1, dz = 1;
    int qform_code = 0, sform_code = 0;
    Mat44 qto_xyz, sto_xyz;
    std::vector<float> data;   // x fastest (filled by to_float(); read() fills it unless keep_raw)
    std::vector<unsigned char> raw;   // voxels in the file's datatype, host byte order (keep_raw only)
    int datatype = 16;         // NIfTI datatype code of `raw`
};

static inline void swap_bytes(void *p, size_t size, size_t count)
{
    unsigned char *b = (unsigned char *)p;
    for (size_t i = 0; i < count; i++, b += size)
        for (size_t j = 0; j < size / 2; j++) { unsigned char t = b[j]; b[j] = b[size - 1 - j]; b[size - 1 - j] = t; }
}

template <typename T>
static inline T rd(const unsigned char *h, size_t off, bool swp)
{
    T v;
    memcpy(&v, h + off, sizeof(T));
    if (swp) swap_bytes(&v, sizeof(T), 1);
    return v;
}

// quaternion -> matrix, NIfTI-1 standard (nifti1.h documentation, "METHOD 2"); long double like nifticlib
static inline Mat44 quatern_to_mat44(float qb, float qc, float qd, float qx, float qy, float qz,
                                     float dx, float dy, float dz, float qfac)
{
    Mat44 R;
    long double a, b = qb, c = qc, d = qd, xd, yd, zd;
    R.m[3][0] = R.m[3][1] = R.m[3][2] = 0.0f; R.m[3][3] = 1.0f;
    a = 1.0l - (b * b + c * c + d * d);
    if (a < 1.e-7l) {
        a = 1.0l / sqrtl(b * b + c * c + d * d);
        b *= a; c *= a; d *= a;
        a = 0.0l;
    } else {
        a = sqrtl(a);
    }
    xd = (dx > 0.0) ? dx : 1.0l;
    yd = (dy > 0.0) ? dy : 1.0l;
    zd = (dz > 0.0) ? dz : 1.0l;
    if (qfac < 0.0) zd = -zd;
    R.m[0][0] = (float)((a * a + b * b - c * c - d * d) * xd);
    R.m[0][1] = (float)(2.0l * (b * c - a * d) * yd);
    R.m[0][2] = (float)(2.0l * (b * d + a * c) * zd);
    R.m[1][0] = (float)(2.0l * (b * c + a * d) * xd);
    R.m[1][1] = (float)((a * a + c * c - b * b - d * d) * yd);
    R.m[1][2] = (float)(2.0l * (c * d - a * b) * zd);
    R.m[2][0] = (float)(2.0l * (b * d - a * c) * xd);
    R.m[2][1] = (float)(2.0l * (c * d + a * b) * yd);
    R.m[2][2] = (float)((a * a + d * d - c * c - b * b) * zd);
    R.m[0][3] = qx; R.m[1][3] = qy; R.m[2][3] = qz;
    return R;
}

static inline bool ends_with(const std::string &s, const char *suf)
{
    size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// returns 0 on success, <0 on failure (like fioReadNifti's negative codes)
// plain casts, as reg_changeDatatype1 does (reference featExtract.cpp:18-77)
static inline void cast_to_float(const void *raw, int datatype, size_t nvox, float *o)
{
    switch (datatype) {
    case 2:   for (size_t i = 0; i < nvox; i++) o[i] = (float)((const unsigned char *)raw)[i]; break;
    case 256: for (size_t i = 0; i < nvox; i++) o[i] = (float)((const signed char *)raw)[i]; break;
    case 4:   for (size_t i = 0; i < nvox; i++) o[i] = (float)((const short *)raw)[i]; break;
    case 512: for (size_t i = 0; i < nvox; i++) o[i] = (float)((const unsigned short *)raw)[i]; break;
    case 8:   for (size_t i = 0; i < nvox; i++) o[i] = (float)((const int *)raw)[i]; break;
    case 768: for (size_t i = 0; i < nvox; i++) o[i] = (float)((const unsigned int *)raw)[i]; break;
    case 16:  memcpy(o, raw, nvox * 4); break;
    case 64:  for (size_t i = 0; i < nvox; i++) o[i] = (float)((const double *)raw)[i]; break;
    }
}

// keep_raw: leave the voxels in their file datatype in img.raw (for the typed C-ABI entry point, which casts
// on the device) instead of converting to float here
static inline int read(const std::string &path, Image &img, bool keep_raw = false)
{
    std::string hdr_path = path, img_path = path;
    if (ends_with(path, ".img")) hdr_path = path.substr(0, path.size() - 4) + ".hdr";
    if (ends_with(path, ".img.gz")) hdr_path = path.substr(0, path.size() - 7) + ".hdr.gz";
    gzFile f = gzopen(hdr_path.c_str(), "rb");
    if (!f) return -1;
    unsigned char h[348];
    if (gzread(f, h, 348) != 348) { gzclose(f); return -1; }
    int sizeof_hdr;
    memcpy(&sizeof_hdr, h, 4);
    bool swp = false;
    if (sizeof_hdr != 348) {
        swap_bytes(&sizeof_hdr, 4, 1);
        if (sizeof_hdr != 348) { gzclose(f); return -1; }
        swp = true;
    }
    short dim[8];
    for (int i = 0; i < 8; i++) dim[i] = rd<short>(h, 40 + 2 * i, swp);
    short datatype = rd<short>(h, 70, swp);
    float pixdim[8];
    for (int i = 0; i < 8; i++) pixdim[i] = rd<float>(h, 76 + 4 * i, swp);
    float vox_offset = rd<float>(h, 108, swp);
    img.qform_code = rd<short>(h, 252, swp);
    img.sform_code = rd<short>(h, 254, swp);
    if (dim[0] < 1 || dim[0] > 7) { gzclose(f); return -1; }
    img.nx = dim[1]; img.ny = dim[0] >= 2 ? dim[2] : 1; img.nz = dim[0] >= 3 ? dim[3] : 1; img.nt = dim[0] >= 4 ? dim[4] : 1;
    if (img.nt < 1) img.nt = 1;
    // trust nothing in the header (nifti_image_read rejects such files too): every used dimension must be positive
    // and the voxel count must stay addressable
    if (img.nx < 1 || img.ny < 1 || img.nz < 1 || (double)img.nx * img.ny * img.nz * img.nt > 1.0e10) { gzclose(f); return -1; }
    img.dx = pixdim[1]; img.dy = pixdim[2]; img.dz = pixdim[3];
    float qfac = (pixdim[0] < 0.0f) ? -1.0f : 1.0f;
    if (img.qform_code > 0) {
        img.qto_xyz = quatern_to_mat44(rd<float>(h, 256, swp), rd<float>(h, 260, swp), rd<float>(h, 264, swp),
                                       rd<float>(h, 268, swp), rd<float>(h, 272, swp), rd<float>(h, 276, swp),
                                       img.dx, img.dy, img.dz, qfac);
    } else {   // no qform: grid spacings only (NIfTI-1 "METHOD 1")
        memset(&img.qto_xyz, 0, sizeof(Mat44));
        img.qto_xyz.m[0][0] = img.dx; img.qto_xyz.m[1][1] = img.dy; img.qto_xyz.m[2][2] = img.dz; img.qto_xyz.m[3][3] = 1.0f;
    }
    memset(&img.sto_xyz, 0, sizeof(Mat44));
    if (img.sform_code > 0) {
        for (int j = 0; j < 4; j++) {
            img.sto_xyz.m[0][j] = rd<float>(h, 280 + 4 * j, swp);
            img.sto_xyz.m[1][j] = rd<float>(h, 296 + 4 * j, swp);
            img.sto_xyz.m[2][j] = rd<float>(h, 312 + 4 * j, swp);
        }
        img.sto_xyz.m[3][3] = 1.0f;
    }
    bool single = (h[344] == 'n' && h[345] == '+');
    size_t nvox = (size_t)img.nx * img.ny * img.nz * img.nt;
    size_t bpv;
    switch (datatype) {
    case 2: case 256: bpv = 1; break;
    case 4: case 512: bpv = 2; break;
    case 8: case 768: case 16: bpv = 4; break;
    case 64: bpv = 8; break;
    default: gzclose(f); return -2;
    }
    if (single && !(vox_offset >= 352.0f)) { gzclose(f); return -1; }     // single-file images keep their data behind the 348 + 4 byte header
    if (single) {
        long skip = (long)vox_offset - 348;
        std::vector<unsigned char> junk(4096);
        while (skip > 0) { int k = gzread(f, junk.data(), (unsigned)(skip > 4096 ? 4096 : skip)); if (k <= 0) break; skip -= k; }
    } else {
        gzclose(f);
        if (ends_with(hdr_path, ".hdr")) img_path = hdr_path.substr(0, hdr_path.size() - 4) + ".img";
        else if (ends_with(hdr_path, ".hdr.gz")) img_path = hdr_path.substr(0, hdr_path.size() - 7) + ".img.gz";
        f = gzopen(img_path.c_str(), "rb");
        if (!f) return -2;
    }
    std::vector<unsigned char> raw(nvox * bpv);
    size_t got = 0;
    while (got < raw.size()) {
        size_t want = raw.size() - got;
        int k = gzread(f, raw.data() + got, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
        if (k <= 0) break;
        got += (size_t)k;
    }
    gzclose(f);
    if (got != raw.size()) return -2;
    if (swp && bpv > 1) swap_bytes(raw.data(), bpv, nvox);
    img.datatype = datatype;
    if (keep_raw) {
        img.raw.swap(raw);
        return 0;
    }
    img.data.resize(nvox);
    cast_to_float(raw.data(), datatype, nvox, img.data.data());
    return 0;
}

// general 4x4 inverse of an affine (last row 0 0 0 1), NIfTI convention
static inline Mat44 inverse(const Mat44 &R)
{
    double r11 = R.m[0][0], r12 = R.m[0][1], r13 = R.m[0][2];
    double r21 = R.m[1][0], r22 = R.m[1][1], r23 = R.m[1][2];
    double r31 = R.m[2][0], r32 = R.m[2][1], r33 = R.m[2][2];
    double v1 = R.m[0][3], v2 = R.m[1][3], v3 = R.m[2][3];
    double deti = r11 * r22 * r33 - r11 * r32 * r23 - r21 * r12 * r33 + r21 * r32 * r13 + r31 * r12 * r23 - r31 * r22 * r13;
    if (deti != 0.0) deti = 1.0 / deti;
    Mat44 Q;
    Q.m[0][0] = (float)(deti * (r22 * r33 - r32 * r23));
    Q.m[0][1] = (float)(deti * (-r12 * r33 + r32 * r13));
    Q.m[0][2] = (float)(deti * (r12 * r23 - r22 * r13));
    Q.m[0][3] = (float)(deti * (-r12 * r23 * v3 + r12 * v2 * r33 + r22 * r13 * v3 - r22 * v1 * r33 - r32 * r13 * v2 + r32 * v1 * r23));
    Q.m[1][0] = (float)(deti * (-r21 * r33 + r31 * r23));
    Q.m[1][1] = (float)(deti * (r11 * r33 - r31 * r13));
    Q.m[1][2] = (float)(deti * (-r11 * r23 + r21 * r13));
    Q.m[1][3] = (float)(deti * (r11 * r23 * v3 - r11 * v2 * r33 - r21 * r13 * v3 + r21 * v1 * r33 + r31 * r13 * v2 - r31 * v1 * r23));
    Q.m[2][0] = (float)(deti * (r21 * r32 - r31 * r22));
    Q.m[2][1] = (float)(deti * (-r11 * r32 + r31 * r12));
    Q.m[2][2] = (float)(deti * (r11 * r22 - r21 * r12));
    Q.m[2][3] = (float)(deti * (-r11 * r22 * v3 + r11 * r32 * v2 + r21 * r12 * v3 - r21 * r32 * v1 - r31 * r12 * v2 + r31 * r22 * v1));
    Q.m[3][0] = Q.m[3][1] = Q.m[3][2] = 0.0f;
    Q.m[3][3] = (deti == 0.0) ? 0.0f : 1.0f;
    return Q;
}

} // namespace niftimin
