// rt_oracle_capi.cpp — C entry points of the CPU oracle (test infrastructure; see rt_oracle.hpp header).
// Row-parallel driver = the reference's rayon loop: one task per image row, dynamic scheduling
// (raytracer.rs:254-262: chunks_mut(width*3).enumerate() -> into_par_iter().for_each(render_line)).
#include "rt_oracle.hpp"

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace rto;

extern "C" {

typedef struct {
    uint64_t rays, samples, draws;
    uint64_t hits[5];
    uint64_t term_sky, term_absorbed, term_depth, term_light;
    uint64_t texture_oob;
    uint64_t path_len_hist[64];
    double   render_ms;
    int32_t  threads;
    int32_t  reserved;
} oracle_stats;

static void export_stats(const Stats& st, double ms, int threads, oracle_stats* out) {
    if (!out) return;
    out->rays = st.rays; out->samples = st.samples; out->draws = st.draws;
    for (int i = 0; i < 5; ++i) out->hits[i] = st.hits[i];
    out->term_sky = st.term_sky; out->term_absorbed = st.term_absorbed; out->term_depth = st.term_depth;
    out->term_light = st.term_light; out->texture_oob = st.texture_oob;
    for (int i = 0; i < 64; ++i) out->path_len_hist[i] = st.path_len_hist[i];
    out->render_ms = ms; out->threads = threads; out->reserved = 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Renders rows [y0, y1) (y1 == 0 means height). Outputs are FULL-frame sized buffers (w*h*3); either may be NULL.
// threads <= 0: all host threads.
int oracle_render(const rt_scene* s, float* out_linear, uint8_t* out_rgb8, uint32_t y0, uint32_t y1, int threads,
                  oracle_stats* stats) {
    if (!s || s->width < 2 || s->height < 2 || s->samples_per_pixel == 0) return -1;
    if (y1 == 0 || y1 > s->height) y1 = s->height;
    Scene sc(s);
    // the reference recursion (raytracer.rs:99-114) does not terminate when n_lights * prob >= 1
    if (sc.lights.size() >= 10) return -4;
#ifdef _OPENMP
    int nt = threads > 0 ? threads : omp_get_max_threads();
#else
    int nt = 1;
#endif
    Stats total;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(nt)
    {
        Stats local;
#pragma omp for schedule(dynamic, 1)
        for (int64_t y = (int64_t)y0; y < (int64_t)y1; ++y) {
            for (uint32_t x = 0; x < s->width; ++x) {
                size_t o = ((size_t)y * s->width + x) * 3;
                render_pixel(sc, x, (uint32_t)y, out_linear ? out_linear + o : nullptr, out_rgb8 ? out_rgb8 + o : nullptr, local);
            }
        }
#pragma omp critical
        total.merge(local);
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    export_stats(total, ms, nt, stats);
    return 0;
}

// One sample's radiance (for per-sample debugging of GPU mismatches).
int oracle_sample(const rt_scene* s, uint32_t x, uint32_t y, uint32_t smp, float out_rgb[3], uint64_t* rays, uint64_t* draws) {
    Scene sc(s);
    Stats st;
    SampleRng rng(s->seed, y * s->width + x, smp);
    double u = ((double)x + rng.gen_f64()) / ((double)s->width - 1.0);
    double v = ((double)s->height - ((double)y + rng.gen_f64())) / ((double)s->height - 1.0);
    Ray r = get_ray(s->camera, u, v);
    Rgb c = ray_color(sc, r, s->max_depth, s->max_depth, rng, st, nullptr, nullptr);
    out_rgb[0] = c.r; out_rgb[1] = c.g; out_rgb[2] = c.b;
    if (rays) *rays = st.rays;
    if (draws) *draws = rng.draws;
    return 0;
}

// ---- known-answer probes (mirror the reference's #[test] functions) -------------------------------
int oracle_camera_new(const rt_camera_params* p, rt_camera* out) { camera_new(*p, out); return 0; }

int oracle_get_ray(const rt_camera* c, double u, double v, rt_vec3* origin, rt_vec3* dir) {
    Ray r = get_ray(*c, u, v);
    *origin = rt_vec3{r.origin.x, r.origin.y, r.origin.z};
    *dir = rt_vec3{r.direction.x, r.direction.y, r.direction.z};
    return 0;
}

int oracle_ray_at(const rt_vec3* o, const rt_vec3* d, double t, rt_vec3* out) {
    P3 p = ray_at(Ray{p3(*o), p3(*d)}, t);
    *out = rt_vec3{p.x, p.y, p.z};
    return 0;
}

int oracle_sphere_hit(const rt_vec3* center, double radius, const rt_vec3* origin, const rt_vec3* dir, double t_min,
                      double t_max, int32_t* hit, double* t, rt_vec3* point, rt_vec3* normal, int32_t* front_face,
                      double* u, double* v) {
    rt_sphere s{};
    s.center = *center; s.radius = radius; s.kind = RT_GLASS; s.param = 1.5; s.texture = -1;
    Hit h{};
    bool ok = sphere_hit(s, 0, Ray{p3(*origin), p3(*dir)}, t_min, t_max, &h);
    *hit = ok ? 1 : 0;
    if (ok) {
        *t = h.t; *point = rt_vec3{h.point.x, h.point.y, h.point.z};
        *normal = rt_vec3{h.normal.x, h.normal.y, h.normal.z}; *front_face = h.front_face ? 1 : 0;
        if (u) *u = h.u;
        if (v) *v = h.v;
    }
    return 0;
}

// atan2 as the oracle evaluates it (mode 1, default: the restated algorithm shared with the kernel; mode 0: host libm).
int oracle_set_atan2_mode(int mode) { int old = g_atan2_mode; g_atan2_mode = mode ? 1 : 0; return old; }
int oracle_atan2(const double* y, const double* x, uint32_t n, int mode, double* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = mode ? rt_atan2(y[i], x[i]) : std::atan2(y[i], x[i]);
    return 0;
}
// u_v_from_sphere_hit_point (sphere.rs:35-43) for n points hp = hit point - centre; out = {u, v} pairs.
int oracle_sphere_uv(const double* hp_xyz, uint32_t n, double* out_uv) {
    for (uint32_t i = 0; i < n; ++i) u_v_from_sphere_hit_point(P3{hp_xyz[3 * i], hp_xyz[3 * i + 1], hp_xyz[3 * i + 2]}, &out_uv[2 * i], &out_uv[2 * i + 1]);
    return 0;
}

int oracle_refract(const rt_vec3* uv, const rt_vec3* n, double eta, rt_vec3* out) {
    P3 r = refract(p3(*uv), p3(*n), eta);
    *out = rt_vec3{r.x, r.y, r.z};
    return 0;
}
int oracle_reflect(const rt_vec3* v, const rt_vec3* n, rt_vec3* out) {
    P3 r = reflect(p3(*v), p3(*n));
    *out = rt_vec3{r.x, r.y, r.z};
    return 0;
}
int oracle_reflectance(double cosine, double ref_idx, double* out) { *out = reflectance(cosine, ref_idx); return 0; }

// ray_color on an arbitrary ray with a fresh (seed, pixel=0, sample=0) stream — raytracer.rs:167-189 uses it on an empty world
int oracle_ray_color(const rt_scene* s, const rt_vec3* o, const rt_vec3* d, uint64_t max_depth, uint64_t depth, float out_rgb[3]) {
    Scene sc(s);
    Stats st;
    SampleRng rng(s->seed, 0, 0);
    Rgb c = ray_color(sc, Ray{p3(*o), p3(*d)}, max_depth, depth, rng, st, nullptr, nullptr);
    out_rgb[0] = c.r; out_rgb[1] = c.g; out_rgb[2] = c.b;
    return 0;
}

int oracle_find_lights(const rt_scene* s, int32_t* out_indices, uint32_t cap) {
    Scene sc(s);
    for (size_t i = 0; i < sc.lights.size() && i < cap; ++i) out_indices[i] = sc.lights[i];
    return (int)sc.lights.size();
}

int oracle_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t kind, uint32_t n, double* out) {
    SampleRng rng(seed, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) out[i] = kind == 0 ? rng.gen_f64() : rng.gen_range_m1_1();
    return 0;
}
int oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { Philox::block(ctr, key, out); return 0; }

int oracle_quantise(const float* mean_linear, uint32_t n, uint8_t* out) {
    for (uint32_t i = 0; i < n; ++i) out[i] = quantise_u8(std::sqrt(mean_linear[i]));
    return 0;
}

int oracle_texture_albedo(const rt_image* tex, double h_offset, double u, double v, float out_rgb[3]) {
    Stats st;
    Rgb c = texture_get_albedo(*tex, h_offset, u, v, st);
    out_rgb[0] = c.r; out_rgb[1] = c.g; out_rgb[2] = c.b;
    return (int)st.texture_oob;
}

// Point3D probes (point3d.rs:196-272)
int oracle_p3_ops(const rt_vec3* a, const rt_vec3* b, double s, double out[24]) {
    P3 A = p3(*a), B = p3(*b);
    P3 r;
    r = A + B; out[0] = r.x; out[1] = r.y; out[2] = r.z;
    r = A - B; out[3] = r.x; out[4] = r.y; out[5] = r.z;
    r = -A; out[6] = r.x; out[7] = r.y; out[8] = r.z;
    r = A * s; out[9] = r.x; out[10] = r.y; out[11] = r.z;
    r = A / s; out[12] = r.x; out[13] = r.y; out[14] = r.z;
    out[15] = dot(A, B); out[16] = length_squared(A); out[17] = length(A);
    r = cross(A, B); out[18] = r.x; out[19] = r.y; out[20] = r.z;
    out[21] = near_zero(A) ? 1.0 : 0.0;
    r = unit_vector(A); out[22] = r.x; out[23] = r.y;
    return 0;
}

}  // extern "C"
