// rtb200_api.cu — the C ABI of include/rtb200.h: scene staging into HBM, batch scheduling of the trace /
// resolve kernels, device<->host copies and error reporting. No CPU render path exists in this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rtb200_kernels.cuh"

using namespace rtk;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
int fail_cuda(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RT_ERR_NO_DEVICE
           : (e == cudaErrorMemoryAllocation ? RT_ERR_OOM : RT_ERR_CUDA);
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

struct GrowBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); if (e != cudaSuccess) return e; p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) return e;
        cap = want;
        return cudaSuccess;
    }
};

// Per-device execution context: one stream, grow-only work buffers, timing events. One render at a time.
struct DeviceCtx {
    bool init = false;
    int device = -1;
    int sm_count = 0;
    size_t max_smem = 0;
    cudaStream_t stream = nullptr;
    // Two sets of per-frame work buffers: a frame loop that alternates two streams lets frame k+1 start tracing while frame k
    // drains its last paths and resolves (rtb200_render_device_async); blocking calls use set 0 only.
    struct WorkSet { GrowBuf samplebuf, accum, stack, small, frames, lterm; } ws[2];
    GrowBuf out_rgb8, out_lin, probe;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
};
DeviceCtx g_ctx[64];

int get_ctx(int device, DeviceCtx** out) {
    if (device < 0) {
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
    }
    if (device >= 64) return fail(RT_ERR_INVALID, "device ordinal out of range");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (device >= count) return fail(RT_ERR_NO_DEVICE, "no such CUDA device");
    CU(cudaSetDevice(device));
    DeviceCtx& c = g_ctx[device];
    if (!c.init) {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            char buf[160];
            snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
            return fail(RT_ERR_NO_DEVICE, buf);
        }
        c.device = device;
        c.sm_count = prop.multiProcessorCount;
        c.max_smem = prop.sharedMemPerBlockOptin;
        CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&c.ev_begin));
        CU(cudaEventCreate(&c.ev_end));
        c.init = true;
    }
    *out = &c;
    return RT_OK;
}

struct V3 { double x, y, z; };
inline V3 v3(const rt_vec3& a) { return V3{a.x, a.y, a.z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline double vlen(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
inline V3 vunit(V3 a) { double l = vlen(a); return V3{a.x / l, a.y / l, a.z / l}; }
inline V3 vcross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline rt_vec3 rv(V3 a) { return rt_vec3{a.x, a.y, a.z}; }

float f32_up(double x) {   // smallest float >= x
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

}  // namespace

struct rtb200_scene_t {
    int device = -1;
    DeviceCtx* ctx = nullptr;
    TraceParams tp{};
    rt_options opts{};
    bool exact = false;
    bool lanes = false;
    int block = 256;
    int minb = 4;
    int grid = 0;
    size_t smem = 0;
    uint32_t spp_batch = 0;
    std::vector<void*> owned;   // device allocations owned by the handle
    struct Upload { const void* src; size_t bytes; void** field; };
    std::vector<Upload> uploads;         // pending scene arrays (commit_uploads)
    std::vector<uint8_t> staging;        // host image of the device arena
    cudaStream_t last_stream = nullptr;   // stream, work set, batch and launch count of the most recently enqueued frame
    int last_set = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};   // distinct streams used by the pending frames
    int n_streams = 0;
    uint32_t frame_counter = 0;
    uint32_t last_batches = 0, last_launches = 0;
    uint32_t pending_frames = 0;          // frames enqueued since the last wait (their events sit in the context's event ring)
    uint64_t h2d_bytes = 0;
};

// Host-side construction of every array the closest-hit stage reads (no CUDA calls): recentring offset, first-level filter
// records (cluster bounds, or the spheres themselves), second-level records + slot map + |c| of the bounds, exact geometry
// and materials. Kept separate from the upload so that CPU tests can check the soundness of the records
// (rtb200_debug_filter_records, tests/test_filter_records_cpu.py).
struct FilterRecords {
    double g[3] = {0, 0, 0};
    bool two_level = false;
    uint32_t n_pairs = 0, n_clusters = 0;
    std::vector<float> first;     // n_pairs * 8 floats, pair-packed {cx0,cx1,cy0,cy1},{cz0,cz1,nk0,nk1}
    std::vector<float> sfilt;     // n_clusters * kClusterK * 4 floats
    std::vector<uint16_t> orig;   // n_clusters * kClusterK
    std::vector<float> cmeta;     // n_clusters
    std::vector<double> geo;      // n * 4
    std::vector<DevMat> mat;      // n
};

static void build_filter_records(const rt_scene* s, bool allow_two_level, FilterRecords& R) {
    const uint32_t n = (uint32_t)s->n_spheres;
    double g[3] = {0, 0, 0};
    // ---- recentring offset of the f32 filter frame: component-wise median of the centres ----
    if (n) {
        std::vector<double> tmp(n);
        for (int c = 0; c < 3; ++c) {
            for (uint32_t i = 0; i < n; ++i) tmp[i] = c == 0 ? s->spheres[i].center.x : (c == 1 ? s->spheres[i].center.y : s->spheres[i].center.z);
            std::nth_element(tmp.begin(), tmp.begin() + n / 2, tmp.end());
            g[c] = tmp[n / 2];
            if (!std::isfinite(g[c])) g[c] = 0.0;
        }
    }

    // ---- device records ----
    const double U = 5.9604644775390625e-8;   // 2^-24
    uint32_t n_pairs = ((n + 1) / 2 + 7) / 8 * 8;   // the scan loop consumes 2 blocks of 4 pairs per trip; padding records never hit
    if (n_pairs == 0) n_pairs = 8;
    std::vector<float> filt((size_t)n_pairs * 8);
    std::vector<double> geo((size_t)std::max<uint32_t>(n, 1) * 4, 0.0);
    std::vector<DevMat> mat(std::max<uint32_t>(n, 1));
    memset(mat.data(), 0, mat.size() * sizeof(DevMat));
    for (uint32_t pp = 0; pp < n_pairs; ++pp) {
        for (int k = 0; k < 2; ++k) {
            uint32_t i = 2 * pp + k;
            float cx = 0.f, cy = 0.f, cz = 0.f, nk = -INFINITY;
            if (i < n) {
                const rt_sphere& sp = s->spheres[i];
                double x = sp.center.x - g[0], y = sp.center.y - g[1], z = sp.center.z - g[2];
                double c2 = x * x + y * y + z * z, r2 = sp.radius * sp.radius;
                // candidate iff  b^2 + 2c.o - K - |o|^2 >= -(Es + Er):  nk = -K + Es rounded up
                double Es = 96.0 * U * c2 + 16.0 * U * r2 + 1e-30;
                double nkd = -(c2 - r2) + Es;
                cx = (float)x; cy = (float)y; cz = (float)z;
                nk = std::isfinite(nkd) ? f32_up(nkd) : INFINITY;
                if (!(std::isfinite(cx) && std::isfinite(cy) && std::isfinite(cz)) || !(c2 < 1e30)) { cx = cy = cz = 0.f; nk = INFINITY; }   // always a candidate
                geo[4 * (size_t)i + 0] = sp.center.x; geo[4 * (size_t)i + 1] = sp.center.y; geo[4 * (size_t)i + 2] = sp.center.z;
                geo[4 * (size_t)i + 3] = sp.radius;
                DevMat& m = mat[i];
                m.kind = sp.kind; m.param = sp.param; m.tex = sp.texture; m.pad = 0;
                if (sp.kind == RT_LAMBERTIAN || sp.kind == RT_METAL) { m.r = sp.albedo[0]; m.g = sp.albedo[1]; m.b = sp.albedo[2]; }
                else { m.r = m.g = m.b = 1.0f; }   // Glass/Light attenuation is (1,1,1) (materials.rs:67,179); Texture uses texels
            }
            // layout: A = {cx0,cx1,cy0,cy1}, B = {cz0,cz1,nk0,nk1}
            float* A = &filt[(size_t)pp * 8];
            A[0 + k] = cx; A[2 + k] = cy; A[4 + k] = cz; A[6 + k] = nk;
        }
    }

    // ---- two-level culling (N4): spheres grouped into clusters of 8 with bounding spheres ----------------------------
    // The cluster bound is tested with the SAME conservative 7-FMA filter as a sphere (a bounding sphere is a sphere),
    // so culling stays a superset of what the exact f64 test can accept; results are identical by construction.
    // Spheres are grouped by radius octave, then by Morton order of their centres; classes with <= 2 members and
    // non-finite spheres become singleton clusters. Cluster k owns sphere-record slots [8k, 8k+8) (padded with
    // never-hit records); orig[] maps a slot back to the sphere index (ties still go to the lowest ORIGINAL index).
    auto make_record = [&](double x, double y, double z, double r2, float rec[4]) {
        double c2 = x * x + y * y + z * z;
        double Es = 96.0 * U * c2 + 16.0 * U * r2 + 1e-30;
        double nkd = -(c2 - r2) + Es;
        rec[0] = (float)x; rec[1] = (float)y; rec[2] = (float)z;
        rec[3] = std::isfinite(nkd) ? f32_up(nkd) : INFINITY;
        if (!(std::isfinite(rec[0]) && std::isfinite(rec[1]) && std::isfinite(rec[2])) || !(c2 < 1e30)) { rec[0] = rec[1] = rec[2] = 0.f; rec[3] = INFINITY; }
    };
    bool two_level = n > 32 && allow_two_level;
    if (const char* e2 = getenv("RTB200_TWO_LEVEL")) two_level = two_level && atoi(e2) != 0;
    std::vector<float> cfilt, sfilt;
    std::vector<uint16_t> orig;
    std::vector<float> cmeta;
    uint32_t n_clusters = 0, n_cpairs = 0;
    if (two_level) {
        constexpr int K = kClusterK;
        std::vector<std::vector<uint32_t>> clusters;
        std::map<int, std::vector<uint32_t>> classes;
        for (uint32_t i = 0; i < n; ++i) {
            const rt_sphere& sp = s->spheres[i];
            bool fin = std::isfinite(sp.center.x) && std::isfinite(sp.center.y) && std::isfinite(sp.center.z) && std::isfinite(sp.radius);
            if (!fin || sp.radius == 0.0) { clusters.push_back({i}); continue; }
            classes[std::ilogb(std::fabs(sp.radius))].push_back(i);
        }
        for (auto& kv : classes) {
            std::vector<uint32_t>& mem = kv.second;
            if (mem.size() <= 2) { for (uint32_t i : mem) clusters.push_back({i}); continue; }
            // top-down median split on the widest axis until a leaf holds <= K spheres: compact leaves of K/2..K members
            std::vector<std::pair<size_t, size_t>> work{{0, mem.size()}};
            auto coord = [&](uint32_t i, int a) { return a == 0 ? s->spheres[i].center.x : a == 1 ? s->spheres[i].center.y : s->spheres[i].center.z; };
            while (!work.empty()) {
                auto [b0, e0] = work.back(); work.pop_back();
                if (e0 - b0 <= (size_t)K) { clusters.emplace_back(mem.begin() + b0, mem.begin() + e0); continue; }
                double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
                for (size_t t = b0; t < e0; ++t) for (int a = 0; a < 3; ++a) { double v = coord(mem[t], a); lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); }
                int ax = 0;
                for (int a = 1; a < 3; ++a) if (hi[a] - lo[a] > hi[ax] - lo[ax]) ax = a;
                // split so that both halves are multiples of K where possible (fewer padded slots)
                size_t cnt = e0 - b0, half = ((cnt / 2 + K - 1) / K) * K;
                if (half >= cnt) half = cnt / 2;
                std::nth_element(mem.begin() + b0, mem.begin() + b0 + half, mem.begin() + e0, [&](uint32_t x, uint32_t y) { double cx = coord(x, ax), cy = coord(y, ax); return cx < cy || (cx == cy && x < y); });
                work.emplace_back(b0, b0 + half); work.emplace_back(b0 + half, e0);
            }
        }
        while (clusters.size() % 4 != 0) clusters.push_back({});   // keep every per-cluster array a multiple of 16 bytes (TMA bulk copies)
        n_clusters = (uint32_t)clusters.size();
        n_cpairs = ((n_clusters + 1) / 2 + 7) / 8 * 8;
        cfilt.assign((size_t)n_cpairs * 8, 0.f);
        sfilt.assign((size_t)n_clusters * K * 4, 0.f);    // K records = K/2 pairs = K float4 per cluster
        orig.assign((size_t)n_clusters * K, 0xffff);
        cmeta.assign(n_clusters, 0.f);
        for (uint32_t pp = 0; pp < n_cpairs; ++pp) { float* A = &cfilt[(size_t)pp * 8]; A[6] = A[7] = -INFINITY; }
        for (uint32_t k = 0; k < n_clusters; ++k) {
            const std::vector<uint32_t>& cl = clusters[k];
            if (cl.empty()) {   // padding cluster: never hit
                float* A = &cfilt[(size_t)(k / 2) * 8]; int kk = k & 1; A[0 + kk] = A[2 + kk] = A[4 + kk] = 0.f; A[6 + kk] = -INFINITY;
                for (int j = 0; j < K; ++j) { float* B = &sfilt[(size_t)k * K * 4 + (size_t)(j / 2) * 8]; B[6 + (j & 1)] = -INFINITY; }
                continue;
            }
            // bounding sphere: centroid + max(|ci - c| + |ri|), inflated; non-finite members make the cluster "always hit"
            double c[3] = {0, 0, 0};
            bool fin = true;
            for (uint32_t i : cl) { c[0] += s->spheres[i].center.x; c[1] += s->spheres[i].center.y; c[2] += s->spheres[i].center.z; }
            for (int a = 0; a < 3; ++a) c[a] /= (double)cl.size();
            double R = 0;
            for (uint32_t i : cl) {
                const rt_sphere& sp = s->spheres[i];
                double dx = sp.center.x - c[0], dy = sp.center.y - c[1], dz = sp.center.z - c[2];
                double dist = std::sqrt(dx * dx + dy * dy + dz * dz) + std::fabs(sp.radius);
                if (!std::isfinite(dist)) fin = false;
                R = std::max(R, dist);
            }
            R = R * (1.0 + 1e-9) + 1e-12;
            float rec[4];
            if (fin) make_record(c[0] - g[0], c[1] - g[1], c[2] - g[2], R * R, rec);
            else { rec[0] = rec[1] = rec[2] = 0.f; rec[3] = INFINITY; }
            { float* A = &cfilt[(size_t)(k / 2) * 8]; int kk = k & 1; A[0 + kk] = rec[0]; A[2 + kk] = rec[1]; A[4 + kk] = rec[2]; A[6 + kk] = rec[3]; }
            {
                double cx = c[0] - g[0], cy = c[1] - g[1], cz = c[2] - g[2];
                double cn = std::sqrt(cx * cx + cy * cy + cz * cz);
                cmeta[k] = (fin && std::isfinite(cn) && cn < 1e15) ? f32_up(cn * (1.0 + 1e-6)) : INFINITY;   // INFINITY disables the behind-origin cull
            }
            for (int j = 0; j < K; ++j) {
                float r4[4] = {0.f, 0.f, 0.f, -INFINITY};
                if (j < (int)cl.size()) {
                    const rt_sphere& sp = s->spheres[cl[j]];
                    make_record(sp.center.x - g[0], sp.center.y - g[1], sp.center.z - g[2], sp.radius * sp.radius, r4);
                    orig[(size_t)k * K + j] = (uint16_t)cl[j];
                }
                float* A = &sfilt[(size_t)k * K * 4 + (size_t)(j / 2) * 8]; int kk = j & 1;
                A[0 + kk] = r4[0]; A[2 + kk] = r4[1]; A[4 + kk] = r4[2]; A[6 + kk] = r4[3];
            }
        }
    }

    R.g[0] = g[0]; R.g[1] = g[1]; R.g[2] = g[2];
    R.two_level = two_level;
    R.n_clusters = n_clusters;
    R.n_pairs = two_level ? n_cpairs : n_pairs;
    R.first = two_level ? std::move(cfilt) : std::move(filt);
    R.sfilt = std::move(sfilt); R.orig = std::move(orig); R.cmeta = std::move(cmeta);
    R.geo = std::move(geo); R.mat = std::move(mat);
}

extern "C" {

int rtb200_abi_version(void) { return RTB200_ABI_VERSION; }
const char* rtb200_last_error(void) { return g_last_error.c_str(); }

// Camera::new — camera.rs:45-77. Host, once per frame, f64, same operation order as the reference.
// (Compiled with -fmad=false / no host contraction: see the Makefile.)
int rtb200_camera_from_params(const rt_camera_params* p, rt_camera* out) {
    if (!p || !out) return fail(RT_ERR_INVALID, "null argument");
    const double PI = 3.14159265358979323846264338327950288;
    double theta = p->vfov_deg * (PI / 180.0);
    double half_height = std::tan(theta / 2.0);
    double half_width = p->aspect * half_height;
    V3 look_from = v3(p->look_from), look_at = v3(p->look_at), vup = v3(p->vup);
    V3 w = vunit(look_from - look_at);
    V3 u = vunit(vcross(vup, w));
    V3 v = vcross(w, u);
    V3 origin = look_from;
    V3 llc = origin - (u * half_width) - (v * half_height) - w;
    V3 horizontal = u * 2.0 * half_width;
    V3 vertical = v * 2.0 * half_height;
    out->origin = rv(origin); out->lower_left_corner = rv(llc); out->horizontal = rv(horizontal); out->vertical = rv(vertical);
    return RT_OK;
}

uint32_t rtb200_shard_rows(uint32_t height, int32_t rank, int32_t world, uint32_t band_rows) {
    if (world <= 1) return height;
    if (band_rows == 0) band_rows = 1;
    uint32_t rows = 0;
    for (uint32_t y = 0; y < height; ++y)
        if ((int32_t)((y / band_rows) % (uint32_t)world) == rank) ++rows;
    return rows;
}

static int render_collect(rtb200_scene_handle h, rt_stats* stats);

// Diagnostic (host only, no GPU needed): the filter records rtb200_scene_upload would stage for `scene`.
// info = {two_level, n_first_level_pairs, n_clusters, cluster_size}; arrays are filled up to their capacities (in elements).
int rtb200_debug_filter_records(const rt_scene* s, uint32_t variant, double recentre[3], uint32_t info[4],
                                float* first, uint64_t cap_first, float* second, uint64_t cap_second,
                                uint16_t* slot_to_sphere, uint64_t cap_slots, float* cluster_abs, uint64_t cap_clusters) {
    if (!s || !info) return fail(RT_ERR_INVALID, "null argument");
    if (s->n_spheres > 65534) return fail(RT_ERR_UNSUPPORTED, "more than 65534 spheres (16-bit candidate indices)");
    FilterRecords R;
    build_filter_records(s, variant != RT_VARIANT_BRUTE_FORCE && variant != RT_VARIANT_LANES && variant != RT_VARIANT_EXACT_F64, R);
    if (recentre) { recentre[0] = R.g[0]; recentre[1] = R.g[1]; recentre[2] = R.g[2]; }
    info[0] = R.two_level ? 1u : 0u; info[1] = R.n_pairs; info[2] = R.n_clusters; info[3] = (uint32_t)kClusterK;
    if (first) memcpy(first, R.first.data(), std::min<uint64_t>(cap_first, R.first.size()) * 4);
    if (second) memcpy(second, R.sfilt.data(), std::min<uint64_t>(cap_second, R.sfilt.size()) * 4);
    if (slot_to_sphere) memcpy(slot_to_sphere, R.orig.data(), std::min<uint64_t>(cap_slots, R.orig.size()) * 2);
    if (cluster_abs) memcpy(cluster_abs, R.cmeta.data(), std::min<uint64_t>(cap_clusters, R.cmeta.size()) * 4);
    return RT_OK;
}

int rtb200_scene_release(rtb200_scene_handle h) {
    if (!h) return RT_OK;
    if (h->device >= 0) cudaSetDevice(h->device);
    if (h->pending_frames) render_collect(h, nullptr);   // frames still in flight read the scene arrays
    for (void* p : h->owned) cudaFree(p);
    delete h;
    return RT_OK;
}

// Scene arrays are collected first and then placed in ONE device arena filled by ONE host->device copy
// (a per-frame upload costs one cudaMalloc, one copy, one cudaFree). `field` is patched with the device address.
static int upload_array(rtb200_scene_t* h, const void* src, size_t bytes, void** field) {
    *field = nullptr;
    if (bytes == 0) bytes = 16;
    h->uploads.push_back(rtb200_scene_t::Upload{src, bytes, field});
    return RT_OK;
}

static int commit_uploads(rtb200_scene_t* h) {
    size_t total = 0;
    for (auto& u : h->uploads) total += (u.bytes + 255) & ~(size_t)255;
    void* base = nullptr;
    CU(cudaMalloc(&base, total ? total : 256));
    h->owned.push_back(base);
    size_t off = 0;
    for (auto& u : h->uploads) { *u.field = (char*)base + off; off += (u.bytes + 255) & ~(size_t)255; }   // addresses first: tables may hold them
    h->staging.assign(total, 0);
    off = 0;
    for (auto& u : h->uploads) {
        if (u.src) { memcpy(h->staging.data() + off, u.src, u.bytes); h->h2d_bytes += u.bytes; }
        off += (u.bytes + 255) & ~(size_t)255;
    }
    if (total) CU(cudaMemcpyAsync(base, h->staging.data(), total, cudaMemcpyHostToDevice, h->ctx->stream));
    h->uploads.clear();
    return RT_OK;
}

int rtb200_scene_upload(const rt_scene* s, const rt_options* opts_in, rtb200_scene_handle* out) {
    if (!s || !out) return fail(RT_ERR_INVALID, "null argument");
    *out = nullptr;
    rt_options opts{};
    opts.device = -1; opts.rank = 0; opts.world = 1; opts.band_rows = 1; opts.variant = RT_VARIANT_AUTO;
    if (opts_in) opts = *opts_in;
    if (opts.world <= 0) opts.world = 1;
    if (opts.band_rows == 0) opts.band_rows = 1;
    if (opts.rank < 0 || opts.rank >= opts.world) return fail(RT_ERR_INVALID, "rank outside [0, world)");
    if (opts.flags != 0) return fail(RT_ERR_INVALID, "flags must be 0");
    if (s->width < 2 || s->height < 2) return fail(RT_ERR_INVALID, "width and height must be >= 2 (u,v divide by w-1, h-1: raytracer.rs:199-200)");
    if (s->samples_per_pixel == 0) return fail(RT_ERR_INVALID, "samples_per_pixel must be > 0");
    if ((uint64_t)s->width * s->height >= (1ull << 31)) return fail(RT_ERR_INVALID, "image too large");
    if (s->n_spheres > 65534) return fail(RT_ERR_UNSUPPORTED, "more than 65534 spheres (16-bit candidate indices)");
    if (s->n_spheres && !s->spheres) return fail(RT_ERR_INVALID, "spheres is null");

    uint32_t n = (uint32_t)s->n_spheres;
    uint32_t n_lights = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const rt_sphere& sp = s->spheres[i];
        if (sp.kind > RT_LIGHT) return fail(RT_ERR_INVALID, "unknown material kind");
        if (sp.kind == RT_LIGHT) ++n_lights;
        if (sp.kind == RT_TEXTURE) {
            if (sp.texture < 0 || (uint64_t)sp.texture >= s->n_textures) return fail(RT_ERR_INVALID, "texture index out of range");
            const rt_image& im = s->textures[sp.texture];
            if (!im.rgb8 || im.width == 0 || im.height == 0) return fail(RT_ERR_INVALID, "empty texture image");
        }
    }
    if (n_lights >= 10) return fail(RT_ERR_UNSUPPORTED, "10 or more lights: the reference's light recursion (raytracer.rs:99-114) does not terminate when n_lights * 0.1 >= 1");
    if (n_lights > 0 && opts.variant == RT_VARIANT_LANES) return fail(RT_ERR_UNSUPPORTED, "RT_VARIANT_LANES has no light support");
    if (s->sky.mode > RT_SKY_TEXTURE) return fail(RT_ERR_INVALID, "unknown sky mode");
    if (s->sky.mode == RT_SKY_TEXTURE && (!s->sky.tex.rgb8 || s->sky.tex.width == 0 || s->sky.tex.height == 0))
        return fail(RT_ERR_INVALID, "sky texture is empty");

    DeviceCtx* ctx = nullptr;
    int rc = get_ctx(opts.device, &ctx);
    if (rc != RT_OK) return rc;

    rtb200_scene_t* h = new rtb200_scene_t();
    h->device = ctx->device; h->ctx = ctx; h->opts = opts;
    h->exact = (opts.variant == RT_VARIANT_EXACT_F64);
    h->lanes = (opts.variant == RT_VARIANT_LANES);
    if (opts.variant > RT_VARIANT_BRUTE_FORCE) return fail(RT_ERR_INVALID, "unknown variant");
    struct Guard { rtb200_scene_t* h; bool ok = false; ~Guard() { if (!ok) rtb200_scene_release(h); } } guard{h};

    FilterRecords R;
    build_filter_records(s, !h->lanes && !h->exact && opts.variant != RT_VARIANT_BRUTE_FORCE, R);
    const bool two_level = R.two_level;
    const uint32_t n_clusters = R.n_clusters;
    uint32_t n_pairs = R.n_pairs;
    const double* g = R.g;
    std::vector<float>& filt = R.first; std::vector<float>& cfilt = R.first;
    std::vector<float>& sfilt = R.sfilt; std::vector<uint16_t>& orig = R.orig; std::vector<float>& cmeta = R.cmeta;
    std::vector<double>& geo = R.geo; std::vector<DevMat>& mat = R.mat;
    const double U = 5.9604644775390625e-8;   // 2^-24
    (void)n_clusters;

    TraceParams& tp = h->tp;
    void* d = nullptr;
    tp.two_level = two_level ? 1u : 0u;
    tp.n_clusters = n_clusters;
    if (two_level) {
        if ((rc = upload_array(h, cfilt.data(), cfilt.size() * 4, (void**)&tp.filt)) != RT_OK) return rc;
        if ((rc = upload_array(h, sfilt.data(), sfilt.size() * 4, (void**)&tp.sfilt)) != RT_OK) return rc;
        if ((rc = upload_array(h, orig.data(), orig.size() * 2, (void**)&tp.orig)) != RT_OK) return rc;
        if ((rc = upload_array(h, cmeta.data(), cmeta.size() * 4, (void**)&tp.cmeta)) != RT_OK) return rc;
    } else {
        if ((rc = upload_array(h, filt.data(), filt.size() * 4, (void**)&tp.filt)) != RT_OK) return rc;
        tp.sfilt = nullptr; tp.orig = nullptr; tp.cmeta = nullptr;
    }
    if ((rc = upload_array(h, geo.data(), geo.size() * 8, (void**)&tp.geo)) != RT_OK) return rc;
    if ((rc = upload_array(h, mat.data(), mat.size() * sizeof(DevMat), (void**)&tp.mat)) != RT_OK) return rc;

    std::vector<rtd::DevTex> texs(std::max<uint64_t>(s->n_textures, 1));
    for (uint64_t t = 0; t < s->n_textures; ++t) {
        const rt_image& im = s->textures[t];
        texs[t].rgb8 = nullptr; texs[t].width = im.width; texs[t].height = im.height;
        if (im.rgb8 && im.width && im.height) {
            if ((rc = upload_array(h, im.rgb8, im.width * im.height * 3, (void**)&texs[t].rgb8)) != RT_OK) return rc;
        }
    }
    if ((rc = upload_array(h, texs.data(), texs.size() * sizeof(rtd::DevTex), (void**)&tp.tex)) != RT_OK) return rc;
    tp.sky_mode = s->sky.mode;
    tp.sky.rgb8 = nullptr; tp.sky.width = 0; tp.sky.height = 0;
    if (s->sky.mode == RT_SKY_TEXTURE) {
        if ((rc = upload_array(h, s->sky.tex.rgb8, s->sky.tex.width * s->sky.tex.height * 3, (void**)&tp.sky.rgb8)) != RT_OK) return rc;
        tp.sky.width = s->sky.tex.width; tp.sky.height = s->sky.tex.height;
    }

    std::vector<uint32_t> lights;
    for (uint32_t i = 0; i < n; ++i) if (s->spheres[i].kind == RT_LIGHT) lights.push_back(i);
    lights.push_back(0);
    if ((rc = upload_array(h, lights.data(), lights.size() * 4, (void**)&tp.lights)) != RT_OK) return rc;
    tp.n = n; tp.n_pairs = n_pairs; tp.n_lights = n_lights;
    tp.gx = g[0]; tp.gy = g[1]; tp.gz = g[2];
    tp.er_coef = 1.0f - (float)(96.0 * U);
    tp.cam = s->camera;
    tp.width = s->width; tp.height = s->height; tp.spp = s->samples_per_pixel; tp.max_depth = s->max_depth;
    tp.key0 = (uint32_t)s->seed; tp.key1 = (uint32_t)(s->seed >> 32);
    tp.rank = opts.rank; tp.world = opts.world; tp.band_rows = opts.band_rows;
    tp.rows_local = rtb200_shard_rows(s->height, opts.rank, opts.world, opts.band_rows);
    tp.npix_local = tp.rows_local * s->width;

    // ---- launch geometry: persistent grid = SMs x resident CTAs; scene fully in shared memory when it fits ----
    if (h->lanes) {
        size_t per_cta_budget = ctx->max_smem;   // opt-in max per block (227 KB)
        size_t half_budget = (228 * 1024 - 2 * 1024 * kCtasPerSm) / kCtasPerSm;
        size_t full_smem = trace_smem_bytes(n, n_pairs, true);
        size_t filt_smem = trace_smem_bytes(n, n_pairs, false);
        int ctas_per_sm = kCtasPerSm;
        if (full_smem <= half_budget) { tp.scene_in_smem = 1; h->smem = full_smem; }
        else if (filt_smem <= half_budget) { tp.scene_in_smem = 0; h->smem = filt_smem; }
        else if (full_smem <= per_cta_budget) { tp.scene_in_smem = 1; h->smem = full_smem; ctas_per_sm = 1; }
        else if (filt_smem <= per_cta_budget) { tp.scene_in_smem = 0; h->smem = filt_smem; ctas_per_sm = 1; }
        else return fail(RT_ERR_UNSUPPORTED, "RT_VARIANT_LANES: the sphere filter records exceed shared memory (use the default variant, which culls through cluster bounds)");
        h->grid = ctx->sm_count * ctas_per_sm;
    } else {
        // What goes to shared memory besides the first-level filter records and the ray pool, in order of value:
        // second-level sphere records (two-level mode), exact geometry (read by every f64 confirmation), materials
        // (read once per hit). Pick the richest set that still leaves the targeted number of CTAs resident per SM.
        // RTB200_WF_SMEM=<mask> overrides (bit0 sfilt, bit1 geo, bit2 mat) for tuning experiments.
        const char* es = getenv("RTB200_WF_SMEM");
        const uint32_t masks[] = {7u, 3u, 1u, 0u};
        bool found = false;
        // first choice: four resident CTAs per SM (kernel built for 64 registers); else the 128-register build at whatever fits
        for (int minb : {4, 2}) {
            for (int need : {minb, 1}) {
                for (uint32_t mask : masks) {
                    if (found) break;
                    if (es && (uint32_t)atoi(es) != mask) continue;
                    if (!two_level && (mask & 1u) && mask != 7u) continue;            // bit0 is meaningless without a second level
                    size_t sm = wavefront_smem_bytes(n, n_pairs, n_clusters, two_level, mask, 256);
                    if (sm > ctx->max_smem) continue;
                    int occ = wavefront_max_ctas_per_sm(sm, minb);
                    if (occ < need || occ <= 0) continue;
                    h->block = 256; h->minb = minb; h->smem = sm; tp.scene_in_smem = mask; h->grid = ctx->sm_count * occ;
                    found = true;
                }
                if (minb == 4) break;   // the 64-register build is only worth it at full residency
            }
            if (found) break;
        }
        if (!found) return fail(RT_ERR_UNSUPPORTED, "first-level filter records exceed shared memory (needs a third level / streaming tiles)");
    }

    // ---- per-sample staging: samples per batch bounded by the buffer cap ----
    uint64_t cap = opts.sample_buffer_bytes ? opts.sample_buffer_bytes : (1ull << 30);
    uint64_t per_spp = (uint64_t)std::max<uint32_t>(tp.npix_local, 1) * 16ull;
    uint64_t spb = std::max<uint64_t>(1, cap / per_spp);
    spb = std::min<uint64_t>(spb, s->samples_per_pixel);
    while (spb > 1 && spb * tp.npix_local >= (1ull << 31)) spb /= 2;
    h->spp_batch = (uint32_t)spb;

    if ((rc = commit_uploads(h)) != RT_OK) return rc;
    CU(cudaStreamSynchronize(ctx->stream));   // host staging vectors go out of scope
    guard.ok = true;
    *out = h;
    return RT_OK;
}

// Enqueue one frame on `stream_in` (or the context's stream) without waiting for it.
static int render_enqueue(rtb200_scene_handle h, void* dev_rgb8, void* dev_linear_f32, void* stream_in, int set) {
    if (!h) return fail(RT_ERR_INVALID, "null scene handle");
    DeviceCtx* ctx = h->ctx;
    DeviceCtx::WorkSet& W = ctx->ws[set & 1];
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream_in ? (cudaStream_t)stream_in : ctx->stream;
    TraceParams tp = h->tp;
    h->last_stream = st; h->last_set = set & 1; h->last_batches = 0; h->last_launches = 0;
    if (h->n_streams < 2 && (h->n_streams == 0 || h->streams[0] != st)) h->streams[h->n_streams++] = st;
    if (tp.npix_local == 0) return RT_OK;

    const uint32_t spp = tp.spp, spb = h->spp_batch;
    const uint32_t n_batches = (spp + spb - 1) / spb;
    const uint32_t threads_total = (uint32_t)h->grid * (uint32_t)(h->lanes ? kBlock : h->block);

    CU(W.samplebuf.ensure((size_t)spb * tp.npix_local * 16));
    CU(W.accum.ensure((size_t)tp.npix_local * 12));
    CU(W.stack.ensure((size_t)std::max<uint32_t>(tp.max_depth, 1) * threads_total * 4));
    CU(W.small.ensure(256 + (size_t)n_batches * 4));
    if (tp.n_lights > 0) {
        // Nested light tests form a branching process: a vertex nests with probability 0.1 n and then spawns n shadow rays, so
        // depth d is reached with probability ~(0.1 n^2 P_hit)^d: harmless for 1-2 lights, near-critical for 3 (the reference
        // itself recurses hundreds of frames deep there) and super-critical beyond. Size the per-path frame stack accordingly;
        // an overflow is reported as an error, never rendered wrongly.
        tp.max_shadow = tp.n_lights == 1 ? 32u : tp.n_lights == 2 ? 96u : 384u;
        CU(W.frames.ensure((size_t)tp.max_shadow * threads_total * sizeof(ShadowFrame)));
        CU(W.lterm.ensure((size_t)6 * threads_total * 4));
    }
    tp.frames = (ShadowFrame*)W.frames.p;
    tp.lterm = (float*)W.lterm.p;
    // event ring: every pending frame owns 2 + 2*n_batches events (begin, end, and a pair around each trace launch)
    const uint32_t kRing = 64, per_frame = 2 + 2 * n_batches;
    if (h->pending_frames >= kRing) return fail(RT_ERR_INVALID, "more than 64 frames enqueued without rtb200_render_device_wait");
    while (ctx->ev.size() < (size_t)kRing * per_frame) {
        cudaEvent_t e; CU(cudaEventCreate(&e)); ctx->ev.push_back(e);
    }
    cudaEvent_t* fev = ctx->ev.data() + (size_t)h->pending_frames * per_frame;
    unsigned long long* stat = (unsigned long long*)W.small.p;
    unsigned int* counters = (unsigned int*)((char*)W.small.p + 256);
    CU(cudaMemsetAsync(W.small.p, 0, 256 + (size_t)n_batches * 4, st));

    tp.samplebuf = (float4*)W.samplebuf.p;
    tp.stack = (uint32_t*)W.stack.p;
    tp.stack_stride = threads_total;
    tp.stat = stat;

    CU(cudaEventRecord(fev[0], st));
    uint32_t launches = 0;
    for (uint32_t b = 0; b < n_batches; ++b) {
        tp.s0 = b * spb;
        tp.s_count = std::min(spb, spp - tp.s0);
        tp.total_work = tp.s_count * tp.npix_local;
        tp.work_counter = counters + b;
        CU(cudaEventRecord(fev[2 + 2 * b], st));
        if (tp.max_depth == 0) {
            CU(cudaMemsetAsync(tp.samplebuf, 0, (size_t)tp.total_work * 16, st));   // ray_color(depth 0) = black, no ray (raytracer.rs:80-82)
        } else if (h->lanes) {
            CU(launch_trace(tp, h->grid, h->smem, false, st));
        } else {
            CU(launch_wavefront(tp, h->grid, h->smem, h->minb, h->exact, st));
        }
        CU(cudaEventRecord(fev[3 + 2 * b], st));
        ResolveParams q{};
        q.samplebuf = tp.samplebuf; q.accum = (float*)W.accum.p; q.npix_local = tp.npix_local; q.s_count = tp.s_count;
        q.first = b == 0; q.last = b + 1 == n_batches; q.spp = spp;
        q.out_linear = (float*)dev_linear_f32; q.out_rgb8 = (uint8_t*)dev_rgb8;
        CU(launch_resolve(q, st));
        launches += 2;
    }
    CU(cudaEventRecord(fev[1], st));
    h->last_batches = n_batches; h->last_launches = launches;
    ++h->pending_frames;
    return RT_OK;
}

// Wait for the most recently enqueued frame of `h` and fetch its statistics.
static int render_collect(rtb200_scene_handle h, rt_stats* stats) {
    if (!h) return fail(RT_ERR_INVALID, "null scene handle");
    DeviceCtx* ctx = h->ctx;
    CU(cudaSetDevice(h->device));
    if (stats) memset(stats, 0, sizeof *stats);
    if (h->tp.npix_local == 0 || h->last_batches == 0 || h->pending_frames == 0) return RT_OK;
    cudaStream_t st = h->last_stream;
    DeviceCtx::WorkSet& W = ctx->ws[h->last_set];
    unsigned long long hstat[16] = {0};
    CU(cudaMemcpyAsync(hstat, W.small.p, sizeof hstat, cudaMemcpyDeviceToHost, st));
    for (int i = 0; i < h->n_streams; ++i) CU(cudaStreamSynchronize(h->streams[i]));
    h->n_streams = 0;
    if (hstat[5] != 0) { h->pending_frames = 0; }
    if (hstat[5] != 0) return fail(RT_ERR_UNSUPPORTED, "light-test recursion deeper than the shadow-frame stack occurred; the frame is not exact (the reference recursion is near-critical for this many lights)");
    if (stats) {
        // device_ms / trace_ms: summed over every frame enqueued since the previous wait; the counters are the last frame's
        float ms = 0.f;
        double dv = 0.0, tr = 0.0;
        const uint32_t per_frame = 2 + 2 * h->last_batches;
        for (uint32_t f = 0; f < h->pending_frames; ++f) {
            cudaEvent_t* fev = ctx->ev.data() + (size_t)f * per_frame;
            CU(cudaEventElapsedTime(&ms, fev[0], fev[1])); dv += ms;
            for (uint32_t b = 0; b < h->last_batches; ++b) { CU(cudaEventElapsedTime(&ms, fev[2 + 2 * b], fev[3 + 2 * b])); tr += ms; }
        }
        stats->device_ms = dv; stats->trace_ms = tr; stats->frames = h->pending_frames;
        stats->rays = hstat[0]; stats->candidates = hstat[1]; stats->samples = hstat[3]; stats->clusters = hstat[4];
        if (getenv("RTB200_PRINT_PHASES")) {
            fprintf(stderr, "[rtb200] stage_mismatch=%llu ovf=%llu phases(warp-cycles): scan=%llu confirm=%llu waitA=%llu sort=%llu shade=%llu waitC=%llu warp_iters=%llu\n",
                    hstat[6], hstat[2], hstat[8], hstat[9], hstat[10], hstat[11], hstat[12], hstat[13], hstat[14]);
        }
        if (h->tp.max_depth == 0) stats->samples = (uint64_t)h->tp.npix_local * h->tp.spp;   // no kernel ran: every sample is black
        stats->kernel_launches = h->last_launches * h->pending_frames; stats->batches = h->last_batches;
    }
    h->pending_frames = 0;
    return RT_OK;
}

int rtb200_render_device(rtb200_scene_handle h, void* dev_rgb8, void* dev_linear_f32, void* stream_in, rt_stats* stats) {
    auto wall0 = std::chrono::steady_clock::now();
    if (h && h->pending_frames) { int rcw = render_collect(h, nullptr); if (rcw != RT_OK) return rcw; }   // drain frames enqueued earlier
    int rc = render_enqueue(h, dev_rgb8, dev_linear_f32, stream_in, 0);
    if (rc != RT_OK) return rc;
    rc = render_collect(h, stats);
    if (rc == RT_OK && stats) stats->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return rc;
}

int rtb200_render_device_async(rtb200_scene_handle h, void* dev_rgb8, void* dev_linear_f32, void* stream_in) {
    if (!h) return fail(RT_ERR_INVALID, "null scene handle");
    return render_enqueue(h, dev_rgb8, dev_linear_f32, stream_in, (int)(h->frame_counter++ & 1u));
}

int rtb200_render_device_wait(rtb200_scene_handle h, rt_stats* stats) { return render_collect(h, stats); }

static int render_host(const rt_scene* s, const rt_options* opts, uint8_t* out_rgb8, float* out_lin, rt_stats* stats) {
    auto wall0 = std::chrono::steady_clock::now();
    rtb200_scene_handle h = nullptr;
    int rc = rtb200_scene_upload(s, opts, &h);
    if (rc != RT_OK) return rc;
    DeviceCtx* ctx = h->ctx;
    size_t npl = h->tp.npix_local;
    void *d8 = nullptr, *dl = nullptr;
    cudaError_t e = cudaSuccess;
    if (out_rgb8) { e = ctx->out_rgb8.ensure(npl * 3 + 16); d8 = ctx->out_rgb8.p; }
    if (e == cudaSuccess && out_lin) { e = ctx->out_lin.ensure(npl * 12 + 16); dl = ctx->out_lin.p; }
    if (e != cudaSuccess) { rtb200_scene_release(h); return fail_cuda(e, "output buffer allocation"); }
    rt_stats st{};
    rc = rtb200_render_device(h, d8, dl, nullptr, &st);
    if (rc == RT_OK && npl) {
        if (out_rgb8) e = cudaMemcpyAsync(out_rgb8, d8, npl * 3, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && out_lin) e = cudaMemcpyAsync(out_lin, dl, npl * 12, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail_cuda(e, "device->host copy of the frame");
    }
    st.h2d_bytes = h->h2d_bytes;
    st.d2h_bytes = (out_rgb8 ? npl * 3 : 0) + (out_lin ? npl * 12 : 0) + 32;
    rtb200_scene_release(h);
    st.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    if (stats) *stats = st;
    return rc;
}

int rtb200_render_rgb8(const rt_scene* scene, const rt_options* opts, uint8_t* out_rgb8, rt_stats* stats) {
    if (!scene || !out_rgb8) return fail(RT_ERR_INVALID, "null argument");
    return render_host(scene, opts, out_rgb8, nullptr, stats);
}
int rtb200_render_linear_f32(const rt_scene* scene, const rt_options* opts, float* out_rgb, rt_stats* stats) {
    if (!scene || !out_rgb) return fail(RT_ERR_INVALID, "null argument");
    return render_host(scene, opts, nullptr, out_rgb, stats);
}

// ---- probes ------------------------------------------------------------------------------------------
static int probe_io(const void* in, size_t in_bytes, size_t out_bytes, DeviceCtx** pctx, void** din, void** dout) {
    int rc = get_ctx(-1, pctx);
    if (rc != RT_OK) return rc;
    DeviceCtx* c = *pctx;
    CU(c->probe.ensure(in_bytes + out_bytes + 512));
    *din = c->probe.p;
    *dout = (char*)c->probe.p + ((in_bytes + 255) / 256) * 256;
    CU(cudaMemsetAsync(*dout, 0, out_bytes, c->stream));
    if (in_bytes) CU(cudaMemcpyAsync(*din, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    return RT_OK;
}
static int probe_finish(DeviceCtx* c, void* host_out, const void* dout, size_t out_bytes) {
    CU(cudaMemcpyAsync(host_out, dout, out_bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return RT_OK;
}

int rtb200_probe_sphere_hit(const rt_vec3* center, double radius, const rt_vec3* origin, const rt_vec3* dir, double t_min,
                            double t_max, int32_t* hit, double* t, rt_vec3* point, rt_vec3* normal, int32_t* front_face) {
    double in[12] = {center->x, center->y, center->z, radius, origin->x, origin->y, origin->z, dir->x, dir->y, dir->z, t_min, t_max};
    double out[9];
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(in, sizeof in, sizeof out, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_sphere_hit((const double*)din, (double*)dout, c->stream));
    if ((rc = probe_finish(c, out, dout, sizeof out)) != RT_OK) return rc;
    *hit = out[0] != 0.0;
    if (*hit) {
        *t = out[1]; *point = rt_vec3{out[2], out[3], out[4]}; *normal = rt_vec3{out[5], out[6], out[7]};
        *front_face = out[8] != 0.0;
    }
    return RT_OK;
}
int rtb200_probe_refract(const rt_vec3* uv, const rt_vec3* n, double eta, rt_vec3* o) {
    double in[7] = {uv->x, uv->y, uv->z, n->x, n->y, n->z, eta}, out[3];
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(in, sizeof in, sizeof out, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_refract((const double*)din, (double*)dout, c->stream));
    if ((rc = probe_finish(c, out, dout, sizeof out)) != RT_OK) return rc;
    *o = rt_vec3{out[0], out[1], out[2]};
    return RT_OK;
}
int rtb200_probe_reflectance(double cosine, double ref_idx, double* o) {
    double in[2] = {cosine, ref_idx};
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(in, sizeof in, 8, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_reflectance((const double*)din, (double*)dout, c->stream));
    return probe_finish(c, o, dout, 8);
}
int rtb200_probe_sky(const rt_vec3* dir, uint32_t sky_mode, float out_rgb[3]) {
    if (sky_mode == RT_SKY_TEXTURE) return fail(RT_ERR_INVALID, "probe_sky supports none/gradient only");
    double in[3] = {dir->x, dir->y, dir->z};
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(in, sizeof in, 12, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_sky((const double*)din, sky_mode, (float*)dout, c->stream));
    return probe_finish(c, out_rgb, dout, 12);
}
int rtb200_probe_get_ray(const rt_camera* cam, double u, double v, rt_vec3* origin, rt_vec3* dir) {
    struct { rt_camera cam; double uv[2]; } in;
    in.cam = *cam; in.uv[0] = u; in.uv[1] = v;
    double out[6];
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(&in, sizeof in, sizeof out, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_get_ray((const rt_camera*)din, (const double*)((char*)din + sizeof(rt_camera)), (double*)dout, c->stream));
    if ((rc = probe_finish(c, out, dout, sizeof out)) != RT_OK) return rc;
    *origin = rt_vec3{out[0], out[1], out[2]}; *dir = rt_vec3{out[3], out[4], out[5]};
    return RT_OK;
}
int rtb200_probe_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t kind, uint32_t n, double* o) {
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(nullptr, 0, (size_t)n * 8, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_rng(seed, pixel, sample, kind, n, (double*)dout, c->stream));
    return probe_finish(c, o, dout, (size_t)n * 8);
}
int rtb200_probe_sphere_uv(const double* hp_xyz, uint32_t n, double* out_uv) {
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(hp_xyz, (size_t)n * 24, (size_t)n * 16, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_sphere_uv((const double*)din, n, (double*)dout, c->stream));
    return probe_finish(c, out_uv, dout, (size_t)n * 16);
}
int rtb200_probe_quantise(const float* mean_linear, uint32_t n, uint8_t* o) {
    DeviceCtx* c; void *din, *dout;
    int rc = probe_io(mean_linear, (size_t)n * 4, n, &c, &din, &dout);
    if (rc != RT_OK) return rc;
    CU(probe_quantise((const float*)din, n, (uint8_t*)dout, c->stream));
    return probe_finish(c, o, dout, n);
}

}  // extern "C"
