// featExtract -- drop-in replacement for the reference CLI (reference featExtract/featExtract.cpp), with
// the whole extraction running on a B200 through the C-ABI of include/s3d.h.
//
//   featExtract [options] <input image> <output features>
//     -w / -ws   output feature geometry in world coordinates (NIfTI qto_xyz / sto_xyz); implies isotropic
//                processing (reference :327-339, 436-473, 507-538)
//     -2+ / -2-  double / halve the input image (reference :368-388)
//     -d[0-9]    CUDA device (reference :312-325; here every run is a GPU run, default device 0)
//     -dA,B,...  several devices: one volume is split into z slabs over them (s3d_multi_extract_slab), a list of
//                volumes (-l) is sharded over them (s3d_multi_batch_extract)
//     -l <file>  batch mode: every line of <file> is "<input image> <output features>"
//     -b -br -bn BRIEF / RRIEF / NRRIEF descriptor instead of SIFT-Rank (README; dead code in the reference)
//     -r X Y Z   input is raw IEEE float32, little endian, x fastest, of these dimensions
//     -r         ... dimensions guessed from the signal (the idea of the reference's unused fioReadRaw,
//                R/src_common/FeatureIO.cpp:3006-3227, made deterministic)
//
// Kept host code: argument parsing, NIfTI/raw loading, isotropic resampling, the world-coordinate
// transform and the text feature file.  Output is byte-identical to the reference's CPU path.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <functional>
#include <mutex>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "nifti_min.h"
#include "s3d.h"

using niftimin::Mat44;

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

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
    printf("  -d[0-9]    : set device id to be used; -dA,B,... : several devices (z slabs of one volume, shards of a list).\n");
    printf("  -l <file>  : batch mode, one \"<input image> <output features>\" pair per line of <file>.\n");
    printf("  -b -br -bn : BRIEF / RRIEF / NRRIEF descriptor (default: gradient orientation histogram).\n");
    printf("  -r X Y Z   : raw float32 input of the given dimensions (-r alone: guess them).\n");
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

// Dimensions of a raw float32 volume from the signal itself: a row length X makes s[i] and s[i + X] neighbours
// (small mean absolute difference), likewise a plane size X*Y.  Deterministic restatement of the idea behind the
// reference's get_image_dimensions (R/src_common/FeatureIO.cpp:3006-3175, random templates, never called).
static bool guess_raw_dims(const float *s, size_t n, int &X, int &Y, int &Z)
{
    auto cost = [&](size_t lag) {
        const size_t samples = 40000, span = n - lag;
        const size_t step = span / samples > 0 ? span / samples : 1;
        double acc = 0; size_t cnt = 0;
        for (size_t i = 0; i < span; i += step) { acc += fabs((double)s[i] - (double)s[i + lag]); cnt++; }
        return cnt ? acc / (double)cnt : 1e300;
    };
    size_t bestX = 0; double bc = 1e300;
    for (size_t L = 8; L <= 8192 && L * 4 <= n; L++)
        if (n % L == 0) { double c = cost(L); if (c < bc) { bc = c; bestX = L; } }
    if (!bestX) return false;
    const size_t rows = n / bestX;
    size_t bestY = 0; bc = 1e300;
    for (size_t k = 2; k <= 8192 && k * 2 <= rows; k++)
        if (rows % k == 0) { double c = cost(bestX * k); if (c < bc) { bc = c; bestY = k; } }
    if (!bestY) return false;
    X = (int)bestX; Y = (int)bestY; Z = (int)(rows / bestY);
    return Z >= 2;
}

// The float volume a job hands to the engine.  In list mode the buffers are page-locked (s3d_host_alloc: the batch entry
// point then copies them at PCIe speed instead of through the driver's pageable staging, 0.5 against 3.6 ms for an MNI
// volume) and recycled through a pool, because page-locking itself costs milliseconds per buffer.
class HostBuf {
public:
    static bool &pinned() { static bool v = false; return v; }
    HostBuf() = default;
    HostBuf(const HostBuf &) = delete;
    HostBuf &operator=(const HostBuf &) = delete;
    HostBuf(HostBuf &&o) noexcept { *this = std::move(o); }
    HostBuf &operator=(HostBuf &&o) noexcept
    {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; cap_ = o.cap_; pin_ = o.pin_; o.p_ = nullptr; o.n_ = o.cap_ = 0; }
        return *this;
    }
    ~HostBuf() { release(); }
    float *data() { return p_; }
    size_t size() const { return n_; }
    void resize(size_t n)          // contents are not kept
    {
        if (n <= cap_ && p_) { n_ = n; return; }
        release();
        if (pinned()) {
            {
                std::lock_guard<std::mutex> g(mu());
                auto &pl = pool();
                for (size_t i = 0; i < pl.size(); i++)
                    if (pl[i].second >= n) { p_ = pl[i].first; cap_ = pl[i].second; pin_ = true; pl.erase(pl.begin() + i); break; }
            }
            if (!p_) { p_ = (float *)s3d_host_alloc(n * sizeof(float)); cap_ = n; pin_ = p_ != nullptr; }
        }
        if (!p_) { p_ = (float *)malloc(n * sizeof(float)); cap_ = n; pin_ = false; }
        n_ = p_ ? n : 0;
    }
    void release()
    {
        if (!p_) return;
        if (pin_) { std::lock_guard<std::mutex> g(mu()); pool().push_back({ p_, cap_ }); }
        else free(p_);
        p_ = nullptr; n_ = cap_ = 0;
    }
    static void drain_pool()
    {
        std::lock_guard<std::mutex> g(mu());
        for (auto &e : pool()) s3d_host_free(e.first);
        pool().clear();
    }
private:
    static std::mutex &mu() { static std::mutex m; return m; }
    static std::vector<std::pair<float *, size_t>> &pool() { static std::vector<std::pair<float *, size_t>> v; return v; }
    float *p_ = nullptr;
    size_t n_ = 0, cap_ = 0;
    bool pin_ = false;
};

struct Options {
    std::vector<int> devices;
    int bDouble = 0, bWorld = 0, bIso = 0, descriptor = S3D_DESC_SIFT;
    int rawX = 0, rawY = 0, rawZ = 0;
    bool raw = false;
    float fEigThres = 140;
};

struct Job {
    std::string in, out;
    niftimin::Image im;
    HostBuf vol;                     // float volume at extraction input resolution (empty while `typed`)
    bool typed = false;              // voxels go to the device in their file datatype (single-device path only)
    bool need_iso = false;           // anisotropic voxels under -w: resampled on the device before the extraction
    std::string log;                 // what the loader has to say (printed in job order: lists are loaded in parallel)
    int X = 0, Y = 0, Z = 0;         // input of the extraction (after isotropic resampling)
    int eX = 0, eY = 0, eZ = 0;      // extraction resolution (after -2+/-2-)
};

static void logf(Job &j, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
static void logf(Job &j, const char *fmt, ...)
{
    char buf[4400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    j.log += buf;
}

// Host half of the loader: read the file (zlib inflate for .gz), datatype conversion.  Touches nothing but the job, so
// the jobs of a list are loaded on several host threads (section 8(f) N1: file decoding off the critical path).
// allow_typed: keep integer voxels for s3d_extract_typed.
static int load_job_host(Job &j, const Options &o, bool allow_typed)
{
    niftimin::Image &im = j.im;
    const char *inPath = j.in.c_str();
    if (o.raw) {
        FILE *f = fopen(inPath, "rb");
        if (!f) { logf(j, "Error: could not read input file: %s\n", inPath); return -1; }
        fseek(f, 0, SEEK_END);
        const long bytes = ftell(f);
        fseek(f, 0, SEEK_SET);
        im.data.resize((size_t)(bytes > 0 ? bytes : 0) / sizeof(float));
        const bool ok = !im.data.empty() && fread(im.data.data(), sizeof(float), im.data.size(), f) == im.data.size();
        fclose(f);
        if (!ok) { logf(j, "Error: could not read input file: %s\n", inPath); return -1; }
        int rx = o.rawX, ry = o.rawY, rz = o.rawZ;
        if (rx <= 0) {
            if (!guess_raw_dims(im.data.data(), im.data.size(), rx, ry, rz)) { logf(j, "Error: could not determine raw dimensions: %s\n", inPath); return -1; }
            logf(j, "Raw dimensions (guessed): %d %d %d\n", rx, ry, rz);
        }
        if (ry <= 0 || rz <= 0 || (size_t)rx * ry * rz > im.data.size()) { logf(j, "Error: bad raw dimensions\n"); return -1; }
        im.nx = rx; im.ny = ry; im.nz = rz;
        memset(&im.qto_xyz, 0, sizeof(Mat44));
        im.qto_xyz.m[0][0] = im.qto_xyz.m[1][1] = im.qto_xyz.m[2][2] = im.qto_xyz.m[3][3] = 1.0f;
        im.sto_xyz = im.qto_xyz;
    } else if (niftimin::read(inPath, im, /*keep_raw=*/true) < 0) {
        logf(j, "Error: could not read input file: %s\n", inPath);
        return -1;
    }
    // NIfTI voxels stay in their file datatype when they can go to the device as they are (the cast to float
    // then runs there, s3d_extract_typed); the isotropic resampling needs floats
    j.need_iso = o.bIso && (im.dx != im.dy || im.dy != im.dz || im.dx != im.dz);
    j.typed = allow_typed && !im.raw.empty() && !j.need_iso && im.datatype != 16;
    j.X = im.nx; j.Y = im.ny; j.Z = im.nz;
    const size_t n3 = (size_t)j.X * j.Y * j.Z;
    if (!im.raw.empty() && !j.typed) {
        if (!j.need_iso) {       // straight into the buffer the engine reads (first volume of a 4-D file)
            j.vol.resize(n3);
            if (j.vol.size() != n3) { logf(j, "Error: out of memory: %s\n", inPath); return -1; }
            niftimin::cast_to_float(im.raw.data(), im.datatype, n3, j.vol.data());
        } else {
            im.data.resize((size_t)im.nx * im.ny * im.nz * im.nt);
            niftimin::cast_to_float(im.raw.data(), im.datatype, im.data.size(), im.data.data());
        }
        im.raw.clear(); im.raw.shrink_to_fit();
    } else if (!j.need_iso && !j.typed) {
        j.vol.resize(n3);
        if (j.vol.size() != n3) { logf(j, "Error: out of memory: %s\n", inPath); return -1; }
        memcpy(j.vol.data(), im.data.data(), n3 * sizeof(float));
        im.data.clear(); im.data.shrink_to_fit();
    }
    return 0;
}

// Device half: isotropic resampling when needed, extraction resolution
static int load_job_device(Job &j, const Options &o, s3d_ctx *ctx)
{
    niftimin::Image &im = j.im;
    const char *inPath = j.in.c_str();
    if (j.need_iso) {
        // isotropic resampling (reference featExtract.cpp:118-204): matrices on the host, voxels on the device
        float fMin = im.dx;
        if (im.dy < fMin) fMin = im.dy;
        if (im.dz < fMin) fMin = im.dz;
        const int nX = (int)(im.nx * im.dx / fMin), nY = (int)(im.ny * im.dy / fMin), nZ = (int)(im.nz * im.dz / fMin);
        const float rf[3] = { fMin / im.dx, fMin / im.dy, fMin / im.dz };
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                im.qto_xyz.m[a][b] *= rf[b];
                if (im.sform_code > 0) im.sto_xyz.m[a][b] *= rf[b];
            }
        if (nX < 1 || nY < 1 || nZ < 1 || j.X < 2 || j.Y < 2 || j.Z < 2) { printf("Could not read volume: %s\n", inPath); return -1; }
        j.vol.resize((size_t)nX * nY * nZ);
        if (s3d_resample_iso_host(ctx, im.data.data(), j.X, j.Y, j.Z, j.vol.data(), nX, nY, nZ, rf[0], rf[1], rf[2]) != S3D_OK) {
            printf("Error: could not resample %s: %s\n", inPath, s3d_last_error(ctx));
            return -1;
        }
        j.X = nX; j.Y = nY; j.Z = nZ;
        im.dx = im.dy = im.dz = fMin;
        im.data.clear(); im.data.shrink_to_fit();
    }
    j.eX = j.X; j.eY = j.Y; j.eZ = j.Z;
    if (o.bDouble == 1) { j.eX *= 2; j.eY *= 2; j.eZ *= 2; }
    else if (o.bDouble == -1) { j.eX /= 2; j.eY /= 2; j.eZ /= 2; }
    if (j.eZ <= 1) { printf("Could not read volume: %s\n", inPath); return -1; }
    printf("Input image: i=%d j=%d k=%d\n", j.eX, j.eY, j.eZ);
    return 0;
}

static int load_job(Job &j, const Options &o, s3d_ctx *ctx, bool allow_typed)
{
    const int rc = load_job_host(j, o, allow_typed);
    fputs(j.log.c_str(), stdout);
    j.log.clear();
    return rc < 0 ? rc : load_job_device(j, o, ctx);
}

// fn(i) for i in [0, n) on up to `threads` host threads
static void parallel_for(size_t n, int threads, const std::function<void(size_t)> &fn)
{
    if (threads > (int)n) threads = (int)n;
    if (threads <= 1) { for (size_t i = 0; i < n; i++) fn(i); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() { for (size_t i = next++; i < n; i = next++) fn(i); });
    for (auto &th : pool) th.join();
}

// world coordinates (reference featExtract.cpp:436-473, 507-538) + the text feature file
static int finish_job(Job &j, const Options &o, s3d_feature *feats, int n)
{
    niftimin::Image &im = j.im;
    Mat44 m_current;
    memset(&m_current, 0, sizeof(m_current));
    if (o.bWorld) {
        const Mat44 *pm = &im.qto_xyz;
        if (o.bWorld == 2) {
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
                for (int q = 0; q < 4; q++) out[i] += m_current.m[i][q] * in[q];
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
    snprintf(c1, sizeof(c1), "Extraction Voxel Resolution (ijk) : %d %d %d", j.eX, j.eY, j.eZ);
    snprintf(c2, sizeof(c2), "Extraction Voxel Size (mm)  (ijk) : %f %f %f", 1.0f * im.dx, 1.0f * im.dy, 1.0f * im.dz);
    if (o.bWorld) {
        snprintf(c3, sizeof(c3), "Feature Coordinate Space: millimeters (%s) : %f %f %f %f %f %f %f %f %f %f %f %f 0.0 0.0 0.0 1.0",
                 o.bWorld == 1 ? "qto_xyz" : "sto_xyz",
                 1.0f * m_current.m[0][0], 1.0f * m_current.m[0][1], 1.0f * m_current.m[0][2], 1.0f * m_current.m[0][3],
                 1.0f * m_current.m[1][0], 1.0f * m_current.m[1][1], 1.0f * m_current.m[1][2], 1.0f * m_current.m[1][3],
                 1.0f * m_current.m[2][0], 1.0f * m_current.m[2][1], 1.0f * m_current.m[2][2], 1.0f * m_current.m[2][3]);
    } else {
        snprintf(c3, sizeof(c3), "Feature Coordinate Space: voxels: 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0 0.0 0.0 0.0 0.0 1.0");
    }
    const char *comments[3] = { c1, c2, c3 };
    if (s3d_write_features_text(j.out.c_str(), feats, n, o.fEigThres, 3, comments) != S3D_OK) {
        printf("Error: could not write %s\n", j.out.c_str());
        return -1;
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) { print_options(); return -1; }
    Options o;
    int iArg = 1;
    const char *listPath = nullptr;
    while (iArg < argc && argv[iArg][0] == '-') {
        switch (argv[iArg][1]) {
        case '2':
            o.bDouble = 1;
            if (argv[iArg][2] == '-') o.bDouble = -1;
            iArg++;
            break;
        case 'd': {
            const int n = s3d_device_count();
            o.devices.clear();
            const char *p = argv[iArg] + 2;
            if (*p == 0) o.devices.push_back(0);
            while (*p) {
                int d = *p - '0';
                if (d < 0 || d > 9 || d >= (n > 0 ? n : 1)) {
                    printf("Error: unknown device: %d\n", d);
                    print_options();
                    return -1;
                }
                o.devices.push_back(d);
                p++;
                if (*p == ',') p++;
                else if (*p) { printf("Error: unknown device: %s\n", argv[iArg]); print_options(); return -1; }
            }
            iArg++;
            break;
        }
        case 'l':
            if (iArg + 1 >= argc) { print_options(); return -1; }
            listPath = argv[iArg + 1];
            iArg += 2;
            break;
        case 'w': case 'W':
            o.bWorld = 1; o.bIso = 1;
            if (argv[iArg][2] == 's' || argv[iArg][2] == 'S') o.bWorld = 2;
            iArg++;
            break;
        case 'b':
            o.descriptor = argv[iArg][2] == 'r' ? S3D_DESC_RRIEF : argv[iArg][2] == 'n' ? S3D_DESC_NRRIEF : S3D_DESC_BRIEF;
            iArg++;
            break;
        case 'r':
            o.raw = true;
            if (iArg + 3 < argc && atoi(argv[iArg + 1]) > 0 && atoi(argv[iArg + 2]) > 0 && atoi(argv[iArg + 3]) > 0 && argc - (iArg + 4) >= 2) {
                o.rawX = atoi(argv[iArg + 1]); o.rawY = atoi(argv[iArg + 2]); o.rawZ = atoi(argv[iArg + 3]);
                iArg += 4;
            } else {
                iArg += 1;          // dimensions are guessed from the signal
            }
            break;
        default:
            printf("Error: unknown command line argument: %s\n", argv[iArg]);
            print_options();
            return -1;
        }
    }
    if (o.devices.empty()) o.devices.push_back(0);
    std::vector<Job> jobs;
    if (listPath) {
        FILE *f = fopen(listPath, "r");
        if (!f) { printf("Error: could not read list file: %s\n", listPath); return -1; }
        char a[4096], b[4096];
        while (fscanf(f, "%4095s %4095s", a, b) == 2) { Job j; j.in = a; j.out = b; jobs.push_back(std::move(j)); }
        fclose(f);
        if (jobs.empty()) { printf("Error: empty list file: %s\n", listPath); return -1; }
    } else {
        if (argc - iArg < 2) { print_options(); return -1; }
        Job j; j.in = argv[iArg]; j.out = argv[iArg + 1];
        jobs.push_back(std::move(j));
    }
    const bool multi = o.devices.size() > 1 || listPath != nullptr;

    s3d_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.double_mode = o.bDouble; prm.descriptor = o.descriptor; prm.eig_thres = o.fEigThres;

    if (!multi) {
        // ---- one volume, one device: the reference's mode of operation
        Job &j = jobs[0];
        printf("Extracting features: %s\n", j.in.c_str());
        s3d_ctx *ctx = nullptr;
        if (s3d_ctx_create(o.devices[0], &ctx) != S3D_OK) {
            printf("Error: could not initialise CUDA device %d: %s\n", o.devices[0], ctx ? s3d_last_error(ctx) : "no device");
            return -1;
        }
        if (load_job(j, o, ctx, /*allow_typed=*/true) < 0) { s3d_ctx_destroy(ctx); return -1; }
        s3d_feature *feats = nullptr;
        int n = 0;
        auto run = [&]() {
            return j.typed ? s3d_extract_typed(ctx, j.im.raw.data(), j.im.datatype, j.X, j.Y, j.Z, &prm, &feats, &n)
                           : s3d_extract(ctx, j.vol.data(), j.X, j.Y, j.Z, &prm, &feats, &n);
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
        const int rc = finish_job(j, o, feats, n);
        s3d_free(feats);
        s3d_ctx_destroy(ctx);
        if (rc < 0) return -1;
        printf("\nDone.\n");
        return 0;
    }

    // ---- several devices and / or a list of volumes
    const bool timing = getenv("S3D_CLI_TIMING") != nullptr;     // phase times of the list mode on stderr
    const double t_start = now_s();
    double t_load = 0, t_dev = 0, t_extract = 0, t_write = 0;
    s3d_multi *m = nullptr;
    if (s3d_multi_create((int)o.devices.size(), o.devices.data(), 4, &m) != S3D_OK) {
        printf("Error: could not initialise CUDA devices: %s\n", m ? s3d_multi_last_error(m) : "no device");
        return -1;
    }
    s3d_ctx *ctx0 = nullptr;       // for the device-side resampling of the loader
    if (s3d_ctx_create(o.devices[0], &ctx0) != S3D_OK) { printf("Error: could not initialise CUDA device %d\n", o.devices[0]); return -1; }
    int rc = 0;
    const double t_created_global = now_s();
    if (!listPath) {
        // one volume over several devices: z slabs
        Job &j = jobs[0];
        printf("Extracting features: %s\n", j.in.c_str());
        if (load_job(j, o, ctx0, /*allow_typed=*/false) < 0) return -1;
        s3d_feature *feats = nullptr;
        int n = 0;
        s3d_status st = s3d_multi_extract_slab(m, j.vol.data(), j.X, j.Y, j.Z, &prm, &feats, &n);
        if (st != S3D_OK) { printf("Error: could not extract features, %s.\n", s3d_multi_last_error(m)); s3d_multi_destroy(m); return -1; }
        rc = finish_job(j, o, feats, n);
        s3d_free(feats);
    } else {
        // A list, in windows of kWindow jobs: the files of window k+1 are read, inflated and converted on the host
        // threads while window k is extracted (runs of equal-shaped volumes are sharded over the devices) and its
        // feature files are written, also on the host threads.  File decoding and text formatting cost two orders of
        // magnitude more host time per volume than the extraction takes on the GPU.
        const size_t kWindow = 32;
        HostBuf::pinned() = true;
        int host_threads = (int)std::thread::hardware_concurrency();
        if (host_threads < 1) host_threads = 1;
        if (host_threads > 32) host_threads = 32;
        std::vector<int> load_rc(jobs.size(), 0);
        auto load_window = [&](size_t a, size_t b) {
            parallel_for(b - a, host_threads, [&](size_t k) { load_rc[a + k] = load_job_host(jobs[a + k], o, /*allow_typed=*/false); });
        };
        const double t_created = now_s();
        load_window(0, std::min(kWindow, jobs.size()));
        t_load += now_s() - t_created;
        for (size_t a = 0; a < jobs.size() && rc == 0; a += kWindow) {
            const size_t b = std::min(a + kWindow, jobs.size());
            std::future<void> next;
            if (b < jobs.size()) next = std::async(std::launch::async, load_window, b, std::min(b + kWindow, jobs.size()));
            double t0 = now_s();
            for (size_t k = a; k < b && rc == 0; k++) {
                printf("Extracting features: %s\n", jobs[k].in.c_str());
                fputs(jobs[k].log.c_str(), stdout);
                if (load_rc[k] < 0 || load_job_device(jobs[k], o, ctx0) < 0) rc = -1;
            }
            t_dev += now_s() - t0; t0 = now_s();
            std::vector<s3d_feature *> rows(b - a, nullptr);
            std::vector<int> n_rows(b - a, 0);
            size_t i0 = a;
            while (i0 < b && rc == 0) {
                size_t i1 = i0 + 1;
                while (i1 < b && jobs[i1].X == jobs[i0].X && jobs[i1].Y == jobs[i0].Y && jobs[i1].Z == jobs[i0].Z) i1++;
                const int nb = (int)(i1 - i0);
                std::vector<const float *> vols(nb);
                for (int k = 0; k < nb; k++) vols[k] = jobs[i0 + k].vol.data();
                s3d_status st = s3d_multi_batch_extract(m, vols.data(), nb, jobs[i0].X, jobs[i0].Y, jobs[i0].Z, &prm, rows.data() + (i0 - a), n_rows.data() + (i0 - a));
                if (st != S3D_OK) { printf("Error: could not extract features, %s.\n", s3d_multi_last_error(m)); rc = -1; }
                i0 = i1;
            }
            t_extract += now_s() - t0; t0 = now_s();
            std::atomic<int> write_rc(0);
            if (rc == 0)
                parallel_for(b - a, host_threads, [&](size_t k) {
                    if (finish_job(jobs[a + k], o, rows[k], n_rows[k]) < 0) write_rc = -1;
                });
            if (write_rc < 0) rc = -1;
            for (size_t k = a; k < b; k++) {
                if (rows[k - a]) s3d_free(rows[k - a]);
                jobs[k].vol.release();
            }
            t_write += now_s() - t0; t0 = now_s();
            if (next.valid()) next.get();
            t_load += now_s() - t0;
        }
    }
    for (Job &j : jobs) j.vol.release();
    HostBuf::drain_pool();
    s3d_ctx_destroy(ctx0);
    if (timing && listPath)
        fprintf(stderr, "featExtract -l: %zu volumes, %d host threads; startup %.3f s, waiting for file decoding %.3f s, device-side loader %.3f s, "
                "extraction %.3f s, world transform + feature files %.3f s, total %.3f s\n", jobs.size(), (int)std::thread::hardware_concurrency(),
                t_created_global - t_start, t_load, t_dev, t_extract, t_write, now_s() - t_start);
    s3d_multi_destroy(m);
    if (rc < 0) return -1;
    printf("\nDone.\n");
    return 0;
}
