// s3d_multi.cu -- multi-GPU entry points of include/s3d.h (SURVEY.md section 8(b) last row, 8(e)): one host thread
// per GPU above the single-GPU C-ABI, so that a C++ host -- the reference's host language, featExtract.cpp is one
// process -- reaches both multi-GPU modes without torch.distributed:
//   * batch (BASELINE config 4): volume i goes to GPU i mod n, no communication, rows returned in input order;
//   * z-slab (BASELINE config 5): ONE volume split into contiguous z slabs.  Per octave a GPU runs the engine on
//     [halo | own | halo] (s3d_params.slab: candidates only from owned planes; zero padding, support-box test and
//     trilinear clamp use the global depth), subsamples its own part of level 3 and pulls the halos of the next
//     octave from its two neighbours with cudaMemcpyPeerAsync (NVLink P2P when the devices allow peer access) --
//     one nearest-neighbour exchange per octave, no collective in the voxel stages.  When slabs get thinner than
//     the halo, the remaining small octaves collapse onto GPU 0.  Rows are merged without a host-side pass: the
//     per-(octave, level, min/max) counts of every slab give each row group its offset in the output array, and
//     every GPU's thread copies its groups there itself.
// The reference has nothing to match here: its launchers overflow at 1024^3 (R/cuda_common/SIFT_cuda_Tools.cu:187
// computes byte counts in int) and it runs one volume per process.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/s3d.h"

namespace {

constexpr int kSlabHalo = 48;        // planes of level 0 kept valid around the owned range, per octave: blur radii
                                     // 3+4+5+6+8 = 26, +1 detection/validation, + the 11^3 patch reach on level 3
constexpr int kInitBlurRadius = 4;   // initial blur: 9 taps (7 after -2+), reference MultiScale.cpp:288-298

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Barrier {                     // reusable host barrier for the per-GPU threads
    std::mutex m; std::condition_variable cv; int n = 1, waiting = 0; unsigned long gen = 0;
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

} // namespace

struct s3d_multi {
    std::vector<int> dev;
    std::vector<s3d_ctx *> ctx;          // slab mode: one context per GPU
    std::vector<s3d_batch *> batch;      // batch mode: created on first use
    int contexts_per_gpu = 4;
    std::string err;
    double last_ms[8] = { 0 };           // slab mode timing of the last call (s3d_multi_last_timing)
};

// Plane ownership for n slabs of an octave-0 volume of depth z0: K = octaves run in slab mode, bounds[r] ..
// bounds[r+1] = planes of octave 0 owned by slab r (multiples of 2^K so every 2x subsample stays inside a slab).
// K is the largest count for which every slab still owns >= halo planes at octave K-1.
static int slab_plan(int z0, int n, int halo, std::vector<int> &bounds)
{
    int bestK = 0;
    bounds.assign(n + 1, z0); bounds[0] = 0;
    for (int K = 1; K <= 12; K++) {
        const int step = 1 << K;
        std::vector<int> b(n + 1);
        b[0] = 0; b[n] = z0;
        for (int r = 1; r < n; r++) b[r] = (int)((double)r * z0 / n / step + 0.5) * step;
        bool ok = true;
        int own_min = 1 << 30;
        for (int r = 0; r < n; r++) {
            if (b[r + 1] <= b[r]) ok = false;
            own_min = std::min(own_min, (b[r + 1] >> (K - 1)) - (b[r] >> (K - 1)));
        }
        if (!ok || own_min < halo || (z0 >> (K - 1)) <= 2) break;
        bestK = K; bounds = b;
    }
    return bestK;
}

extern "C" s3d_status s3d_multi_create(int n_gpus, const int *devices, int contexts_per_gpu, s3d_multi **out)
{
    if (!out || n_gpus < 1 || n_gpus > 64) return S3D_ERR_INVALID;
    *out = nullptr;
    s3d_multi *m = new s3d_multi();
    *out = m;
    m->contexts_per_gpu = contexts_per_gpu > 0 ? contexts_per_gpu : 4;
    for (int r = 0; r < n_gpus; r++) m->dev.push_back(devices ? devices[r] : r);
    m->ctx.assign(n_gpus, nullptr);
    m->batch.assign(n_gpus, nullptr);
    for (int r = 0; r < n_gpus; r++) {
        s3d_status st = s3d_ctx_create(m->dev[r], &m->ctx[r]);
        if (st != S3D_OK) { m->err = m->ctx[r] ? s3d_last_error(m->ctx[r]) : "s3d_ctx_create failed"; return st; }
    }
    // the slab path allocates its per-octave staging buffers stream-ordered (cudaMallocAsync); keep what it frees in the
    // pool instead of handing gigabytes back to the driver at every synchronisation (measured: 125-480 ms per call)
    for (int r = 0; r < n_gpus; r++) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, m->dev[r]) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    // direct peer copies between neighbouring slabs when the hardware allows it (cudaMemcpyPeerAsync works either way)
    for (int r = 0; r + 1 < n_gpus; r++) {
        const int a = m->dev[r], b = m->dev[r + 1];
        if (a == b) continue;
        int ab = 0, ba = 0;
        cudaDeviceCanAccessPeer(&ab, a, b);
        cudaDeviceCanAccessPeer(&ba, b, a);
        if (ab) { cudaSetDevice(a); cudaDeviceEnablePeerAccess(b, 0); }
        if (ba) { cudaSetDevice(b); cudaDeviceEnablePeerAccess(a, 0); }
        cudaGetLastError();      // "already enabled" is not an error
    }
    {   // GPU 0 gathers every slab's part of octave K for the collapsed tail
        for (int r = 1; r < n_gpus; r++) {
            if (m->dev[r] == m->dev[0]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, m->dev[0], m->dev[r]);
            if (can) { cudaSetDevice(m->dev[0]); cudaDeviceEnablePeerAccess(m->dev[r], 0); }
            cudaGetLastError();
        }
    }
    return S3D_OK;
}

extern "C" void s3d_multi_destroy(s3d_multi *m)
{
    if (!m) return;
    for (s3d_batch *b : m->batch) if (b) s3d_batch_destroy(b);
    for (s3d_ctx *c : m->ctx) if (c) s3d_ctx_destroy(c);
    delete m;
}

extern "C" const char *s3d_multi_last_error(const s3d_multi *m) { return m ? m->err.c_str() : "null handle"; }
extern "C" int s3d_multi_gpu_count(const s3d_multi *m) { return m ? (int)m->dev.size() : 0; }

// ---------------------------------------------------------------------------------------------------
// batch mode: volume i -> GPU i mod n
// ---------------------------------------------------------------------------------------------------
extern "C" s3d_status s3d_multi_batch_extract(s3d_multi *m, const float *const *h_volumes, int n_volumes, int X, int Y, int Z,
                                              const s3d_params *prm, s3d_feature **rows, int *n_rows)
{
    if (!m || !h_volumes || n_volumes < 0 || !prm || !rows || !n_rows) return S3D_ERR_INVALID;
    const int n = (int)m->dev.size();
    std::vector<s3d_status> st(n, S3D_OK);
    std::vector<std::string> msg(n);
    std::vector<std::thread> th;
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            std::vector<const float *> mine;
            for (int i = r; i < n_volumes; i += n) mine.push_back(h_volumes[i]);
            if (mine.empty()) return;
            if (!m->batch[r]) {
                st[r] = s3d_batch_create(m->dev[r], m->contexts_per_gpu, &m->batch[r]);
                if (st[r] != S3D_OK) { msg[r] = m->batch[r] ? s3d_batch_last_error(m->batch[r]) : "s3d_batch_create failed"; return; }
            }
            std::vector<s3d_feature *> rr(mine.size(), nullptr);
            std::vector<int> nn(mine.size(), 0);
            st[r] = s3d_batch_extract(m->batch[r], mine.data(), (int)mine.size(), X, Y, Z, prm, rr.data(), nn.data());
            if (st[r] != S3D_OK) msg[r] = s3d_batch_last_error(m->batch[r]);
            for (size_t k = 0; k < mine.size(); k++) { rows[r + (int)k * n] = rr[k]; n_rows[r + (int)k * n] = nn[k]; }
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < n; r++)
        if (st[r] != S3D_OK) { m->err = "GPU " + std::to_string(m->dev[r]) + ": " + msg[r]; return st[r]; }
    return S3D_OK;
}

// ---------------------------------------------------------------------------------------------------
// z-slab mode
// ---------------------------------------------------------------------------------------------------
namespace {

struct SlabShared {
    s3d_multi *m;
    const float *vol; int X, Y, Z;        // input volume (host, dense)
    s3d_params prm;
    int n, K; std::vector<int> bounds;
    int X0, Y0, Z0;                       // pre-stepped (octave 0) dimensions
    Barrier bar;
    std::mutex err_m; s3d_status status = S3D_OK; std::string err;
    // per slab: own part of the current octave's level 0 (device, dense) and its plane count
    std::vector<float *> own_g0; std::vector<int> own_n;
    // per slab, per octave: fetched rows and their group counts (group = (level-1)*2 + is_max)
    std::vector<std::vector<s3d_feature *>> rows; std::vector<std::vector<std::vector<long long>>> cnt;
    s3d_feature *tail = nullptr; int n_tail = 0;
    s3d_feature *out = nullptr; long long n_out = 0;
    std::vector<std::vector<std::vector<long long>>> off;    // [octave][group][slab] offset in out

    bool failed() { std::lock_guard<std::mutex> lk(err_m); return status != S3D_OK; }
    void fail(s3d_status st, const std::string &what)
    {
        std::lock_guard<std::mutex> lk(err_m);
        if (status == S3D_OK) { status = st; err = what; }
    }
};

#define SLAB_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { S.fail(S3D_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); return; } } while (0)
#define SLAB_S3(call) do { s3d_status s_ = (call); if (s_ != S3D_OK) { S.fail(s_, std::string(#call) + ": " + s3d_last_error(ctx)); return; } } while (0)

// one octave of one slab: extraction (with a capacity retry), row fetch, group counts
static void slab_run_octave(SlabShared &S, int r, int o, s3d_ctx *ctx, const float *d_buf, int Xo, int Yo, int nz, int z_off, int Zo,
                            int own0, int own1, int &kp_cap)
{
    for (int attempt = 0; attempt < 6; attempt++) {
        s3d_params p = S.prm;
        p.double_mode = 0;
        p.max_keypoints = kp_cap; p.max_features = 0;
        p.input_is_g0 = o > 0; p.octave_base = o; p.max_octaves = 1;
        p.slab = 1; p.z_off = z_off; p.z_global = Zo; p.own_z0 = own0; p.own_z1 = own1;
        p.pre_step_done = S.prm.double_mode;
        SLAB_S3(s3d_extract_device(ctx, d_buf, Xo, Yo, nz, &p));
        s3d_feature *rows = nullptr; int n_rows = 0;
        s3d_status st = s3d_fetch_features(ctx, &rows, &n_rows);
        if (st == S3D_ERR_CAPACITY && attempt < 5) { kp_cap *= 2; continue; }     // keypoints are data dependent: grow and rerun
        if (st != S3D_OK) { S.fail(st, std::string("s3d_fetch_features: ") + s3d_last_error(ctx)); return; }
        s3d_keypoint *kps = nullptr; int n_kp = 0; int *row_kp = nullptr; int n_rk = 0;
        SLAB_S3(s3d_get_keypoints(ctx, &kps, &n_kp));
        SLAB_S3(s3d_get_row_keypoints(ctx, &row_kp, &n_rk));
        std::vector<long long> c(6, 0);
        int prev = -1;
        for (int i = 0; i < n_rk; i++) {
            const s3d_keypoint &k = kps[row_kp[i]];
            const int g = (k.level - 1) * 2 + (k.is_max ? 1 : 0);
            if (g < prev || g < 0 || g > 5) { S.fail(S3D_ERR_INVALID, "slab rows are not in (level, min/max) order"); break; }
            prev = g; c[g]++;
        }
        s3d_free(kps); s3d_free(row_kp);
        S.rows[r][o] = rows; S.cnt[r][o] = c;
        return;
    }
}

static void slab_thread(SlabShared &S, int r)
{
    s3d_ctx *ctx = S.m->ctx[r];
    const int dev = S.m->dev[r], n = S.n, K = S.K, halo = kSlabHalo;
    cudaSetDevice(dev);
    cudaStream_t st = (cudaStream_t)s3d_stream(ctx);
    int Xo = S.X0, Yo = S.Y0, Zo = S.Z0;
    // keypoint capacity: data dependent; start from the slab's share of one keypoint per 2048 voxels and grow on demand
    long long vox = (long long)S.X0 * S.Y0 * (S.Z0 / n + 2 * halo);
    int kp_cap = S.prm.max_keypoints > 0 ? S.prm.max_keypoints : (int)std::min<long long>(1 << 20, std::max<long long>(16384, vox / 2048));
    const bool timing = getenv("S3D_SLAB_TIMING") != nullptr;      // per-phase host wall clock of every slab on stderr
    double t_asm = 0, t_run = 0, t_next = 0, t_bar = 0, t0 = now_ms();
    auto lap = [&](double &acc) { if (timing) { cudaStreamSynchronize(st); const double t = now_ms(); acc += t - t0; t0 = t; } };
    std::vector<void *> to_free;
    // Every octave has exactly two barrier waits per thread, reached whether or not a phase failed (a failing phase
    // records the error and returns; the other threads see S.failed() and skip their phases too).
    for (int o = 0; o < K; o++) {
        const int own0 = S.bounds[r] >> o, own1 = (r + 1 < n) ? (S.bounds[r + 1] >> o) : Zo;
        float *d_buf = nullptr; int z_off = 0, nz = 0;
        const size_t plane = (size_t)Xo * Yo;
        auto assemble = [&]() {       // [halo | own | halo] of this octave's input
            if (o == 0) {
                const int hi_ = halo + kInitBlurRadius;
                const int lo = std::max(0, own0 - hi_), hi = std::min(Zo, own1 + hi_);
                z_off = lo; nz = hi - lo;
                const size_t in_plane = (size_t)S.X * S.Y;
                if (S.prm.double_mode == 0) {
                    SLAB_CU(cudaMallocAsync((void **)&d_buf, sizeof(float) * plane * nz, st));
                    to_free.push_back(d_buf);
                    SLAB_CU(cudaMemcpyAsync(d_buf, S.vol + in_plane * lo, sizeof(float) * plane * nz, cudaMemcpyHostToDevice, st));
                } else if (S.prm.double_mode == 1) {      // fioDoubleSize: doubled plane 2z+dz needs original planes z, z+1 (clamped)
                    const int o0 = lo / 2, o1 = std::min(S.Z, (hi - 1) / 2 + 2);
                    float *d_src = nullptr, *d_dst = nullptr;
                    SLAB_CU(cudaMallocAsync((void **)&d_src, sizeof(float) * in_plane * (o1 - o0), st));
                    to_free.push_back(d_src);
                    SLAB_CU(cudaMallocAsync((void **)&d_dst, sizeof(float) * plane * 2 * (o1 - o0), st));
                    to_free.push_back(d_dst);
                    SLAB_CU(cudaMemcpyAsync(d_src, S.vol + in_plane * o0, sizeof(float) * in_plane * (o1 - o0), cudaMemcpyHostToDevice, st));
                    SLAB_S3(s3d_double_size(ctx, d_src, S.X, S.Y, o1 - o0, S.X, d_dst, 2 * S.X));
                    d_buf = d_dst + plane * (lo - 2 * o0);
                } else {                                   // fioSubSample2DCenterPixel
                    const int o0 = 2 * lo, o1 = 2 * hi;
                    float *d_src = nullptr;
                    SLAB_CU(cudaMallocAsync((void **)&d_src, sizeof(float) * in_plane * (o1 - o0), st));
                    to_free.push_back(d_src);
                    SLAB_CU(cudaMallocAsync((void **)&d_buf, sizeof(float) * plane * nz, st));
                    to_free.push_back(d_buf);
                    SLAB_CU(cudaMemcpyAsync(d_src, S.vol + in_plane * o0, sizeof(float) * in_plane * (o1 - o0), cudaMemcpyHostToDevice, st));
                    SLAB_S3(s3d_halve_size(ctx, d_src, S.X, S.Y, o1 - o0, S.X, d_buf, Xo));
                }
            } else {
                const int lo_h = r > 0 ? halo : 0, hi_h = (r + 1 < n) ? halo : 0;
                z_off = own0 - lo_h; nz = lo_h + (own1 - own0) + hi_h;
                SLAB_CU(cudaMallocAsync((void **)&d_buf, sizeof(float) * plane * nz, st));
                to_free.push_back(d_buf);
                if (lo_h) SLAB_CU(cudaMemcpyPeerAsync(d_buf, dev, S.own_g0[r - 1] + plane * (S.own_n[r - 1] - halo), S.m->dev[r - 1], sizeof(float) * plane * halo, st));
                SLAB_CU(cudaMemcpyAsync(d_buf + plane * lo_h, S.own_g0[r], sizeof(float) * plane * (own1 - own0), cudaMemcpyDeviceToDevice, st));
                if (hi_h) SLAB_CU(cudaMemcpyPeerAsync(d_buf + plane * (lo_h + own1 - own0), dev, S.own_g0[r + 1], S.m->dev[r + 1], sizeof(float) * plane * halo, st));
                SLAB_CU(cudaStreamSynchronize(st));
            }
        };
        auto next_level0 = [&]() {    // own part of the next octave's level 0: 2x2x2 mean of the own planes of level 3
            const int n_next = (own1 >> 1) - (own0 >> 1);
            float *d_next = nullptr;
            const float *g3 = nullptr; int g3_pitch = 0;
            SLAB_CU(cudaMallocAsync((void **)&d_next, sizeof(float) * (size_t)(Xo / 2) * (Yo / 2) * n_next, st));
            S.own_g0[r] = d_next; S.own_n[r] = n_next;
            SLAB_S3(s3d_level_device_ptr(ctx, 0, 0, 3, &g3, &g3_pitch, nullptr));
            SLAB_S3(s3d_subsample2(ctx, g3 + (size_t)(own0 - z_off) * Yo * g3_pitch, Xo, Yo, 2 * n_next, g3_pitch, d_next, Xo / 2));
            SLAB_CU(cudaStreamSynchronize(st));
        };
        if (!S.failed()) assemble();
        lap(t_asm);
        S.bar.wait();            // every slab holds its halos: the previous octave's own parts can go
        lap(t_bar);
        if (o > 0 && S.own_g0[r]) { cudaFreeAsync(S.own_g0[r], st); S.own_g0[r] = nullptr; }
        if (!S.failed()) slab_run_octave(S, r, o, ctx, d_buf, Xo, Yo, nz, z_off, Zo, own0, own1, kp_cap);
        lap(t_run);
        if (!S.failed()) next_level0();
        lap(t_next);
        for (void *p : to_free) cudaFreeAsync(p, st);
        to_free.clear();
        S.bar.wait();            // every slab's next level 0 is complete
        lap(t_bar);
        Xo /= 2; Yo /= 2; Zo /= 2;
    }
    cudaStreamSynchronize(st);
    if (timing) fprintf(stderr, "s3d slab %d: assemble %.1f ms, octave run + fetch %.1f ms, next level 0 %.1f ms, barriers %.1f ms (K = %d, keypoint capacity %d)\n", r, t_asm, t_run, t_next, t_bar, K, kp_cap);
}

} // namespace

extern "C" s3d_status s3d_multi_extract_slab(s3d_multi *m, const float *h_volume, int X, int Y, int Z, const s3d_params *prm,
                                             s3d_feature **out, int *n_out)
{
    if (!m || !h_volume || !prm || !out || !n_out || X < 1 || Y < 1 || Z < 1) return S3D_ERR_INVALID;
    *out = nullptr; *n_out = 0;
    if (prm->slab || prm->input_is_g0 || prm->octave_base || prm->max_octaves || prm->pre_step_done) {
        m->err = "s3d_multi_extract_slab: octave-run / slab fields of s3d_params must be zero";
        return S3D_ERR_INVALID;
    }
    const int n = (int)m->dev.size();
    SlabShared S;
    S.m = m; S.vol = h_volume; S.X = X; S.Y = Y; S.Z = Z; S.prm = *prm; S.n = n;
    S.X0 = X; S.Y0 = Y; S.Z0 = Z;
    if (prm->double_mode == 1) { S.X0 *= 2; S.Y0 *= 2; S.Z0 *= 2; }
    else if (prm->double_mode == -1) { S.X0 /= 2; S.Y0 /= 2; S.Z0 /= 2; }
    S.K = n > 1 ? slab_plan(S.Z0, n, kSlabHalo, S.bounds) : 0;
    if (S.K == 0) {      // too thin to split (or one GPU): whole volume on GPU 0
        s3d_status st = s3d_extract(m->ctx[0], h_volume, X, Y, Z, prm, out, n_out);
        if (st != S3D_OK) m->err = s3d_last_error(m->ctx[0]);
        return st;
    }
    const int K = S.K;
    const bool timing = getenv("S3D_SLAB_TIMING") != nullptr;
    double tt0 = now_ms();
    S.bar.n = n;
    S.own_g0.assign(n, nullptr); S.own_n.assign(n, 0);
    S.rows.assign(n, std::vector<s3d_feature *>(K, nullptr));
    S.cnt.assign(n, std::vector<std::vector<long long>>(K, std::vector<long long>(6, 0)));
    {
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++) th.emplace_back([&S, r] { slab_thread(S, r); });
        for (auto &t : th) t.join();
    }
    auto cleanup = [&]() {
        for (int r = 0; r < n; r++) {
            for (int o = 0; o < K; o++) if (S.rows[r][o]) s3d_free(S.rows[r][o]);
            if (S.own_g0[r]) { cudaSetDevice(m->dev[r]); cudaFree(S.own_g0[r]); }
        }
        if (S.tail) s3d_free(S.tail);
    };
    if (S.status != S3D_OK) { m->err = S.err; cleanup(); return S.status; }
    const double t_slabs = now_ms() - tt0; tt0 = now_ms();

    // ---- collapse: the remaining octaves run on GPU 0 from the gathered level 0 of octave K
    const int Xk = S.X0 >> K, Yk = S.Y0 >> K, Zk = S.Z0 >> K;
    if (std::min(Xk, std::min(Yk, Zk)) > 2) {
        s3d_ctx *ctx = m->ctx[0];
        cudaSetDevice(m->dev[0]);
        cudaStream_t st = (cudaStream_t)s3d_stream(ctx);
        float *full = nullptr;
        const size_t plane = (size_t)Xk * Yk;
        cudaError_t e = cudaMallocAsync((void **)&full, sizeof(float) * plane * Zk, st);
        size_t pos = 0;
        for (int r = 0; r < n && e == cudaSuccess; r++) {
            e = cudaMemcpyPeerAsync(full + plane * pos, m->dev[0], S.own_g0[r], m->dev[r], sizeof(float) * plane * S.own_n[r], st);
            pos += S.own_n[r];
        }
        if (e != cudaSuccess || (int)pos != Zk) {
            m->err = e != cudaSuccess ? std::string("slab collapse: ") + cudaGetErrorString(e) : "slab collapse: plane count mismatch";
            cleanup();
            return e != cudaSuccess ? S3D_ERR_CUDA : S3D_ERR_INVALID;
        }
        s3d_params p = *prm;
        p.double_mode = 0; p.input_is_g0 = 1; p.octave_base = K; p.pre_step_done = prm->double_mode;
        s3d_status s = s3d_extract_device(ctx, full, Xk, Yk, Zk, &p);
        if (s == S3D_OK) s = s3d_fetch_features(ctx, &S.tail, &S.n_tail);
        cudaFreeAsync(full, st);
        if (s != S3D_OK) { m->err = s3d_last_error(ctx); cleanup(); return s; }
    }

    const double t_tail = now_ms() - tt0; tt0 = now_ms();
    // ---- merge: octave, level, minima then maxima, then slabs in z order (= raster order); offsets from the counts
    long long total = 0;
    S.off.assign(K, std::vector<std::vector<long long>>(6, std::vector<long long>(n, 0)));
    for (int o = 0; o < K; o++)
        for (int g = 0; g < 6; g++)
            for (int r = 0; r < n; r++) { S.off[o][g][r] = total; total += S.cnt[r][o][g]; }
    const long long tail_off = total;
    total += S.n_tail;
    if (total > 0x7fffffffll) { m->err = "too many feature rows"; cleanup(); return S3D_ERR_CAPACITY; }
    s3d_feature *res = (s3d_feature *)malloc(sizeof(s3d_feature) * (size_t)(total > 0 ? total : 1));
    if (!res) { m->err = "out of host memory"; cleanup(); return S3D_ERR_NOMEM; }
    {
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++)
            th.emplace_back([&, r] {
                for (int o = 0; o < K; o++) {
                    long long src = 0;
                    for (int g = 0; g < 6; g++) {
                        const long long c = S.cnt[r][o][g];
                        if (c) memcpy(res + S.off[o][g][r], S.rows[r][o] + src, sizeof(s3d_feature) * (size_t)c);
                        src += c;
                    }
                }
                if (r == 0 && S.n_tail) memcpy(res + tail_off, S.tail, sizeof(s3d_feature) * (size_t)S.n_tail);
            });
        for (auto &t : th) t.join();
    }
    cleanup();
    if (timing) fprintf(stderr, "s3d slab mode: slab octaves %.1f ms, collapsed tail %.1f ms, merge + cleanup %.1f ms, %lld rows\n", t_slabs, t_tail, now_ms() - tt0, total);
    *out = res;
    *n_out = (int)total;
    return S3D_OK;
}
