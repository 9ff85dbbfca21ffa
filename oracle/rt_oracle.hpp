// rt_oracle.hpp — CPU restatement of the dps/rust-raytracer per-pixel render loop.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / `--impl reference` legs of bench.py may build, load or call anything under oracle/.
// The product library (rust-raytracer_b200/csrc) never includes or links this file.
//
// What it restates (file:line under /root/reference/raytracer/src/), operation for operation,
// f64 geometry / f32 colour exactly as the Rust code evaluates them (compile with -ffp-contract=off:
// rustc never contracts a*b+c into an FMA):
//   point3d.rs:22-38,52-86,89-171   Point3D arithmetic, random(), random_in_unit_sphere()
//   ray.rs:18-20                    Ray::at
//   camera.rs:45-84                 Camera::new, Camera::get_ray
//   sphere.rs:35-43,46-78           u_v_from_sphere_hit_point, Sphere::hit
//   materials.rs:65-69,84-95,111-129,144-155,176-199,236-267   scatter for all five materials
//   raytracer.rs:44-59,61-69,71-165,191-218,220-229            hit_world, clamp, ray_color, render_line, find_lights
//
// Parity status ("pinning"): the reference cannot be built here (no cargo/rustc, crates not vendored)
// and ships no golden image. The oracle is pinned against every known-answer test the reference's own
// test-suite holds for this path (tests/test_oracle_kat.py): sphere.rs:81-88, materials.rs:157-174,
// raytracer.rs:167-189, camera.rs:87-122, ray.rs:52-63, point3d.rs:196-272, raytracer.rs:231-248.
// A second restatement written separately (tests/py_restatement.py, pure Python) must produce the same bits
// (tests/test_oracle_vs_python_restatement.py).
// PARITY UNPINNED at three third-party boundaries whose crates are absent from /root/reference:
//   rand 0.8.x      (the reference draws from an OS-seeded ThreadRng, never reproducible) — replaced by the
//                   counter-based Philox4x32-10 stream defined below, drawn in the reference's draw ORDER;
//   palette 0.6.0   f32 -> u8 `into_format` (restated from its published source: min(x*255,255) + 2^23 trick,
//                   i.e. round-half-even; no reference test covers it);
//   jpeg-decoder    texture decode (host side, not in this file).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/rtb200.h"

namespace rto {

// ------------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11). One stream per (seed, pixel, sample).
// counter = (block, sample, pixel, 0), key = (seed_lo, seed_hi). Block b yields u64 draws 2b, 2b+1,
// each assembled low-word-first like rand_core's BlockRng::next_u64.
// ------------------------------------------------------------------------------------------------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    static void block(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
        uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
        uint32_t k0 = key_in[0], k1 = key_in[1];
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
            uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
            uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += W0; k1 += W1;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

struct SampleRng {
    uint32_t key[2];
    uint32_t pixel, sample, blk;
    uint64_t cached;
    bool has_cached;
    uint64_t draws;
    SampleRng(uint64_t seed, uint32_t pixel_, uint32_t sample_)
        : pixel(pixel_), sample(sample_), blk(0), cached(0), has_cached(false), draws(0) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    }
    uint64_t next_u64() {
        ++draws;
        if (has_cached) { has_cached = false; return cached; }
        uint32_t ctr[4] = {blk, sample, pixel, 0u}, w[4];
        Philox::block(ctr, key, w);
        ++blk;
        cached = ((uint64_t)w[3] << 32) | w[2];
        has_cached = true;
        return ((uint64_t)w[1] << 32) | w[0];
    }
    // rand 0.8 `Standard` for f64: 53 random bits * 2^-53 in [0,1)   (call sites raytracer.rs:100,199,200; materials.rs:189)
    double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rand 0.8 UniformFloat::sample_single(-1.0, 1.0): 52 mantissa bits -> [1,2), minus 1, times (high-low), plus low
    // (call site point3d.rs:25-27)
    double gen_range_m1_1() {
        uint64_t bits = (next_u64() >> 12) | 0x3FF0000000000000ull;
        double v12; std::memcpy(&v12, &bits, 8);
        double v01 = v12 - 1.0;
        return v01 * 2.0 + (-1.0);
    }
};

// ------------------------------------------------------------------------------------------------
// Point3D — point3d.rs
// ------------------------------------------------------------------------------------------------
struct P3 { double x, y, z; };
static inline P3 p3(const rt_vec3& v) { return P3{v.x, v.y, v.z}; }
static inline P3 operator+(P3 a, P3 b) { return P3{a.x + b.x, a.y + b.y, a.z + b.z}; }          // :89-99
static inline P3 operator-(P3 a, P3 b) { return P3{a.x - b.x, a.y - b.y, a.z - b.z}; }          // :101-111
static inline P3 operator-(P3 a) { return P3{-a.x, -a.y, -a.z}; }                                // :113-123
static inline P3 operator*(P3 a, double s) { return P3{a.x * s, a.y * s, a.z * s}; }             // :137-147
static inline P3 operator/(P3 a, double s) { return P3{a.x / s, a.y / s, a.z / s}; }             // :161-171
static inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }               // :72-74
static inline double length_squared(P3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }          // :59-61
static inline double length(P3 a) {                                                              // :63-65 via distance :52-57
    double dx = a.x - 0.0, dy = a.y - 0.0, dz = a.z - 0.0;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}
static inline P3 unit_vector(P3 a) { double l = length(a); return P3{a.x / l, a.y / l, a.z / l}; } // :67-70 (three divisions)
static inline P3 cross(P3 a, P3 b) {                                                             // :76-82
    return P3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline bool near_zero(P3 a) {                                                             // :84-86
    const double e = std::numeric_limits<double>::epsilon();
    return std::fabs(a.x) < e && std::fabs(a.y) < e && std::fabs(a.z) < e;
}
static inline P3 random_in_unit_sphere(SampleRng& rng) {                                         // :22-38
    for (;;) {
        double x = rng.gen_range_m1_1(), y = rng.gen_range_m1_1(), z = rng.gen_range_m1_1();
        P3 p{x, y, z};
        if (length_squared(p) < 1.0) return p;
    }
}

struct Ray { P3 origin, direction; };
static inline P3 ray_at(const Ray& r, double t) { return r.origin + r.direction * t; }          // ray.rs:18-20

// ------------------------------------------------------------------------------------------------
// Camera — camera.rs
// ------------------------------------------------------------------------------------------------
static inline void camera_new(const rt_camera_params& p, rt_camera* out) {                       // camera.rs:45-77
    const double PI = 3.14159265358979323846264338327950288;
    double theta = p.vfov_deg * (PI / 180.0);          // f64::to_radians
    double half_height = std::tan(theta / 2.0);
    double half_width = p.aspect * half_height;
    P3 look_from = p3(p.look_from), look_at = p3(p.look_at), vup = p3(p.vup);
    P3 w = unit_vector(look_from - look_at);
    P3 u = unit_vector(cross(vup, w));
    P3 v = cross(w, u);
    P3 origin = look_from;
    P3 llc = origin - (u * half_width) - (v * half_height) - w;
    P3 horizontal = u * 2.0 * half_width;
    P3 vertical = v * 2.0 * half_height;
    out->origin = rt_vec3{origin.x, origin.y, origin.z};
    out->lower_left_corner = rt_vec3{llc.x, llc.y, llc.z};
    out->horizontal = rt_vec3{horizontal.x, horizontal.y, horizontal.z};
    out->vertical = rt_vec3{vertical.x, vertical.y, vertical.z};
}
static inline Ray get_ray(const rt_camera& c, double u, double v) {                              // camera.rs:79-84
    P3 o = p3(c.origin);
    return Ray{o, p3(c.lower_left_corner) + (p3(c.horizontal) * u) + (p3(c.vertical) * v) - o};
}

// ------------------------------------------------------------------------------------------------
// Sphere::hit — sphere.rs:35-78
// ------------------------------------------------------------------------------------------------
struct Hit { double t; P3 point, normal; bool front_face; int sphere; double u, v; };

// f64::atan2 (sphere.rs:38). Rust forwards to the platform libm, so the reference's last bit is platform-defined.
// To make texel addresses reproducible on every host AND on the GPU, oracle and kernel share one explicit algorithm:
// the classic table-free atan (argument reduction at 7/16, 11/16, 19/16, 39/16 + an odd minimax polynomial of degree
// 23, < 1 ulp) and the usual quadrant logic of atan2, evaluated in plain IEEE f64 without contraction. The device
// copy is rtd::rt_atan2 (csrc/rtb200_device.cuh); tests compare the two bit for bit, and compare this one with the
// host libm (tests/test_oracle_semantics.py: <= 1 ulp apart; identical texels on the C1 frame). g_atan2_mode = 0
// switches the oracle back to the host libm for that comparison.
static int g_atan2_mode = 1;
static inline uint32_t hi_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (uint32_t)(b >> 32); }
static inline uint32_t lo_word(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (uint32_t)b; }
static inline double rt_atan(double x) {
    static const double hi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    static const double lo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    static const double T[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                                 9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                                 4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    const uint32_t hx = hi_word(x), ix = hx & 0x7fffffffu;
    const bool negative = (hx >> 31) != 0;
    if (ix >= 0x44100000u) {                       // |x| >= 2^66, inf or NaN
        if (x != x) return x + x;
        const double r = hi[3] + lo[3];
        return negative ? -r : r;
    }
    int id;
    if (ix < 0x3fdc0000u) {                        // |x| < 7/16
        if (ix < 0x3e400000u) return x;            // |x| < 2^-27
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3ff30000u) {                    // |x| < 19/16
            if (ix < 0x3fe60000u) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }
            else { id = 1; x = (x - 1.0) / (x + 1.0); }
        } else if (ix < 0x40038000u) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }   // |x| < 39/16
        else { id = 3; x = -1.0 / x; }
    }
    const double z = x * x, w = z * z;
    const double s1 = z * (T[0] + w * (T[2] + w * (T[4] + w * (T[6] + w * (T[8] + w * T[10])))));
    const double s2 = w * (T[1] + w * (T[3] + w * (T[5] + w * (T[7] + w * T[9]))));
    if (id < 0) return x - x * (s1 + s2);
    const double r = hi[id] - ((x * (s1 + s2) - lo[id]) - x);
    return negative ? -r : r;
}
static inline double rt_atan2(double y, double x) {
    const double PI_HI = 3.1415926535897931160e+00, PI_LO = 1.2246467991473531772e-16, TINY = 1.0e-300;
    if (x != x || y != y) return x + y;
    const uint32_t hx = hi_word(x), hy = hi_word(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (x == 1.0) return rt_atan(y);
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);   // 2*sign(x) + sign(y)
    if ((iy | lo_word(y)) == 0u) {                 // y = +-0
        switch (m) { case 0: case 1: return y; case 2: return PI_HI + TINY; default: return -PI_HI - TINY; }
    }
    if ((ix | lo_word(x)) == 0u) return (hy >> 31) ? -PI_HI / 2.0 - TINY : PI_HI / 2.0 + TINY;
    if (ix == 0x7ff00000u) {                       // x = +-inf
        if (iy == 0x7ff00000u) {
            switch (m) { case 0: return PI_HI / 4.0 + TINY; case 1: return -PI_HI / 4.0 - TINY;
                         case 2: return 3.0 * (PI_HI / 4.0) + TINY; default: return -3.0 * (PI_HI / 4.0) - TINY; }
        }
        switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return PI_HI + TINY; default: return -PI_HI - TINY; }
    }
    if (iy == 0x7ff00000u) return (hy >> 31) ? -PI_HI / 2.0 - TINY : PI_HI / 2.0 + TINY;
    const int k = ((int)iy - (int)ix) >> 20;       // exponent difference
    double z;
    int mm = m;
    if (k > 60) { z = PI_HI / 2.0 + 0.5 * PI_LO; mm &= 1; }
    else if ((hx >> 31) && k < -60) z = 0.0;
    else z = rt_atan(std::fabs(y / x));
    switch (mm) {
        case 0: return z;
        case 1: return -z;
        case 2: return PI_HI - (z - PI_LO);
        default: return (z - PI_LO) - PI_HI;
    }
}

static inline void u_v_from_sphere_hit_point(P3 hp, double* u, double* v) {                      // sphere.rs:35-43
    const double PI = 3.14159265358979323846264338327950288;
    P3 n = unit_vector(hp);
    const double at = g_atan2_mode ? rt_atan2(n.x, n.z) : std::atan2(n.x, n.z);
    *u = (at / (2.0 * PI)) + 0.5;
    *v = n.y * 0.5 + 0.5;
}
static inline bool sphere_hit(const rt_sphere& s, int index, const Ray& ray, double t_min, double t_max, Hit* h) {
    P3 c = p3(s.center);
    P3 oc = ray.origin - c;
    double a = length_squared(ray.direction);
    double half_b = dot(oc, ray.direction);
    double cc = length_squared(oc) - s.radius * s.radius;
    double discriminant = (half_b * half_b) - (a * cc);
    if (discriminant >= 0.0) {
        double sqrtd = std::sqrt(discriminant);
        double roots[2] = {((-half_b) - sqrtd) / a, ((-half_b) + sqrtd) / a};
        for (double root : roots) {
            if (root < t_max && root > t_min) {
                P3 p = ray_at(ray, root);
                P3 normal = (p - c) / s.radius;
                bool front_face = dot(ray.direction, normal) < 0.0;
                double u, v;
                u_v_from_sphere_hit_point(p - c, &u, &v);
                h->t = root; h->point = p; h->normal = front_face ? normal : -normal;
                h->front_face = front_face; h->sphere = index; h->u = u; h->v = v;
                return true;
            }
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Materials — materials.rs
// ------------------------------------------------------------------------------------------------
struct Rgb { float r, g, b; };
static inline P3 reflect(P3 v, P3 n) { return v - n * (2.0 * dot(v, n)); }                       // :111-113
static inline P3 refract(P3 uv, P3 n, double etai_over_etat) {                                   // :144-149
    double cos_theta = std::fmin(dot(-uv, n), 1.0);
    P3 r_out_perp = (uv + n * cos_theta) * etai_over_etat;
    P3 r_out_parallel = n * (-1.0 * std::sqrt(std::fabs(1.0 - length_squared(r_out_perp))));
    return r_out_perp + r_out_parallel;
}
static inline double powi5(double x) { double x2 = x * x; double x4 = x2 * x2; return x * x4; } // f64::powi(5): x * (x^2)^2 (compiler-rt __powidf2)
static inline double reflectance(double cosine, double ref_idx) {                                // :151-155
    double r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0 - r0) * powi5(1.0 - cosine);
}

struct Stats {
    uint64_t rays = 0, samples = 0, draws = 0;
    uint64_t hits[5] = {0, 0, 0, 0, 0};
    uint64_t term_sky = 0, term_absorbed = 0, term_depth = 0, term_light = 0;
    uint64_t texture_oob = 0;
    uint64_t path_len_hist[64] = {0};
    void merge(const Stats& o) {
        rays += o.rays; samples += o.samples; draws += o.draws;
        for (int i = 0; i < 5; ++i) hits[i] += o.hits[i];
        term_sky += o.term_sky; term_absorbed += o.term_absorbed; term_depth += o.term_depth; term_light += o.term_light;
        texture_oob += o.texture_oob;
        for (int i = 0; i < 64; ++i) path_len_hist[i] += o.path_len_hist[i];
    }
};

struct Scene {
    const rt_scene* s;
    std::vector<int> lights;   // find_lights: indices of spheres whose material is Light, list order  (raytracer.rs:220-229)
    explicit Scene(const rt_scene* sc) : s(sc) {
        for (uint64_t i = 0; i < sc->n_spheres; ++i)
            if (sc->spheres[i].kind == RT_LIGHT) lights.push_back((int)i);
    }
};

static inline uint64_t f64_as_u64_sat(double x) {  // Rust `as u64`: saturating, NaN -> 0
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)x;
}
static inline uint64_t f32_as_usize_sat(float x) {
    if (!(x > 0.0f)) return 0;
    if (x >= 18446744073709551616.0f) return ~0ull;
    return (uint64_t)x;
}

static inline Rgb texture_get_albedo(const rt_image& tex, double h_offset, double u, double v, Stats& st) {  // :236-253
    double rot = u + h_offset;
    if (rot > 1.0) rot = rot - 1.0;
    double uu = rot * (double)tex.width;
    double vv = (1.0 - v) * (double)(tex.height - 1);
    uint64_t base = 3ull * (f64_as_u64_sat(std::floor(vv)) * tex.width + f64_as_u64_sat(std::floor(uu)));
    uint64_t limit = tex.width * tex.height * 3ull;
    if (base + 2 >= limit) {      // the reference would panic (index out of bounds); clamp and count
        ++st.texture_oob;
        base = limit - 3;
    }
    return Rgb{(float)tex.rgb8[base] / 255.0f, (float)tex.rgb8[base + 1] / 255.0f, (float)tex.rgb8[base + 2] / 255.0f};
}

// Material::scatter — materials.rs:44-54. Returns: 0 = None (absorbed), 1 = Some((None, albedo)) (Light),
// 2 = Some((Some(ray), albedo)).
static inline int scatter(const Scene& sc, const rt_sphere& sph, const Ray& ray, const Hit& h, SampleRng& rng,
                          Ray* out, Rgb* albedo, Stats& st) {
    switch (sph.kind) {
    case RT_LAMBERTIAN:    // :84-95
    case RT_TEXTURE: {     // :256-267
        P3 dir = h.normal + random_in_unit_sphere(rng);
        if (near_zero(dir)) dir = h.normal;
        P3 target = h.point + dir;
        *out = Ray{h.point, target - h.point};
        if (sph.kind == RT_LAMBERTIAN) *albedo = Rgb{sph.albedo[0], sph.albedo[1], sph.albedo[2]};
        else *albedo = texture_get_albedo(sc.s->textures[sph.texture], sph.param, h.u, h.v, st);
        return 2;
    }
    case RT_METAL: {       // :115-129
        P3 reflected = reflect(ray.direction, h.normal);
        Ray scattered{h.point, reflected + random_in_unit_sphere(rng) * sph.param};
        *albedo = Rgb{sph.albedo[0], sph.albedo[1], sph.albedo[2]};
        if (dot(scattered.direction, h.normal) > 0.0) { *out = scattered; return 2; }
        return 0;
    }
    case RT_GLASS: {       // :176-199
        *albedo = Rgb{1.0f, 1.0f, 1.0f};
        double ratio = h.front_face ? 1.0 / sph.param : sph.param;
        P3 ud = unit_vector(ray.direction);
        double cos_theta = std::fmin(dot(-ud, h.normal), 1.0);
        double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0;
        // short-circuit: the uniform is drawn only when refraction is possible (:189)
        if (cannot_refract || reflectance(cos_theta, ratio) > rng.gen_f64()) {
            *out = Ray{h.point, reflect(ud, h.normal)};
        } else {
            *out = Ray{h.point, refract(ud, h.normal, ratio)};
        }
        return 2;
    }
    case RT_LIGHT:         // :65-69
    default:
        *albedo = Rgb{1.0f, 1.0f, 1.0f};
        return 1;
    }
}

static inline float clampf(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }         // raytracer.rs:61-69

static inline bool hit_world(const Scene& sc, const Ray& r, double t_min, double t_max, Hit* out, Stats& st) {  // :44-59
    ++st.rays;
    double closest = t_max;
    bool any = false;
    Hit h{};
    for (uint64_t i = 0; i < sc.s->n_spheres; ++i) {
        if (sphere_hit(sc.s->spheres[i], (int)i, r, t_min, closest, &h)) {
            closest = h.t;
            *out = h;
            any = true;
        }
    }
    return any;
}

static inline Rgb sky_color(const Scene& sc, const Ray& ray) {                                   // raytracer.rs:134-163
    P3 ud = unit_vector(ray.direction);
    float t = clampf(0.5f * ((float)ud.y + 1.0f));
    float u = clampf(0.5f * ((float)ud.x + 1.0f));
    const rt_sky& sky = sc.s->sky;
    if (sky.mode == RT_SKY_NONE) return Rgb{0.0f, 0.0f, 0.0f};
    if (sky.mode == RT_SKY_GRADIENT)
        return Rgb{(1.0f - t) * 1.0f + t * 0.5f, (1.0f - t) * 1.0f + t * 0.7f, (1.0f - t) * 1.0f + t * 1.0f};
    uint64_t W = sky.tex.width, H = sky.tex.height;
    uint64_t x = f32_as_usize_sat(u * (float)(W - 1));
    uint64_t y = f32_as_usize_sat((1.0f - t) * (float)(H - 1));
    const uint8_t* px = sky.tex.rgb8 + (y * W + x) * 3;
    return Rgb{0.7f * (float)px[0] / 255.0f, 0.7f * (float)px[1] / 255.0f, 0.7f * (float)px[2] / 255.0f};
}

// ray_color — raytracer.rs:71-165 (recursive, like the reference). `max_depth`/`depth` are usize in the
// reference; `depth > max_depth - 2` wraps for max_depth < 2 in a release build, i.e. is false.
static Rgb ray_color(const Scene& sc, const Ray& ray, uint64_t max_depth, uint64_t depth, SampleRng& rng,
                     Stats& st, int* path_len, int* term) {
    if (depth <= 0) { if (term) *term = 2; return Rgb{0.0f, 0.0f, 0.0f}; }
    Hit h{};
    if (path_len) ++*path_len;
    if (hit_world(sc, ray, 0.001, std::numeric_limits<double>::max(), &h, st)) {
        const rt_sphere& sph = sc.s->spheres[h.sphere];
        ++st.hits[sph.kind < 5 ? sph.kind : 4];
        Ray sr; Rgb albedo;
        int k = scatter(sc, sph, ray, h, rng, &sr, &albedo, st);
        if (k == 0) { if (term) *term = 1; return Rgb{0.0f, 0.0f, 0.0f}; }
        float light_red = 0.0f, light_green = 0.0f, light_blue = 0.0f;
        double prob = 0.1;
        if (sph.kind == RT_GLASS) prob = 0.05;
        size_t nl = sc.lights.size();
        uint64_t md2 = max_depth - 2;   // wrapping, as in a release build
        if (nl > 0 && rng.gen_f64() > (1.0 - (double)nl * prob) && depth > md2) {
            for (int li : sc.lights) {
                const rt_sphere& L = sc.s->spheres[li];
                Ray light_ray{h.point, p3(L.center) - h.point};
                Rgb tc = ray_color(sc, light_ray, 2, 1, rng, st, nullptr, nullptr);
                light_red += albedo.r * tc.r;
                light_green += albedo.g * tc.g;
                light_blue += albedo.b * tc.b;
            }
            light_red /= (float)nl; light_green /= (float)nl; light_blue /= (float)nl;
        }
        if (k == 2) {
            Rgb tc = ray_color(sc, sr, max_depth, depth - 1, rng, st, path_len, term);
            return Rgb{clampf(light_red + albedo.r * tc.r), clampf(light_green + albedo.g * tc.g),
                       clampf(light_blue + albedo.b * tc.b)};
        }
        if (term) *term = 3;
        return albedo;
    }
    if (term) *term = 0;
    return sky_color(sc, ray);
}

// palette 0.6 `FromComponent<f32> for u8` (into_format): min(x*255,255) + 2^23, low mantissa bits = round-half-even.
static inline uint8_t quantise_u8(float x) {
    float scaled = std::fmin(x * 255.0f, 255.0f);
    float f = scaled + 8388608.0f;
    uint32_t bits; std::memcpy(&bits, &f, 4);
    const uint32_t C23 = 0x4B000000u;
    uint32_t d = bits >= C23 ? bits - C23 : 0u;
    return (uint8_t)d;
}

// render_line for one pixel — raytracer.rs:196-217. linear_out = scale*sum (mean radiance) per channel.
static inline void render_pixel(const Scene& sc, uint32_t x, uint32_t y, float linear_out[3], uint8_t rgb8_out[3], Stats& st) {
    const rt_scene& s = *sc.s;
    float acc[3] = {0.0f, 0.0f, 0.0f};
    uint32_t pixel = y * s.width + x;
    for (uint32_t smp = 0; smp < s.samples_per_pixel; ++smp) {
        SampleRng rng(s.seed, pixel, smp);
        double u = ((double)x + rng.gen_f64()) / ((double)s.width - 1.0);
        double v = ((double)s.height - ((double)y + rng.gen_f64())) / ((double)s.height - 1.0);
        Ray r = get_ray(s.camera, u, v);
        int plen = 0, term = 0;
        Rgb c = ray_color(sc, r, s.max_depth, s.max_depth, rng, st, &plen, &term);
        acc[0] += c.r; acc[1] += c.g; acc[2] += c.b;
        ++st.samples; st.draws += rng.draws;
        st.path_len_hist[plen < 63 ? plen : 63]++;
        if (term == 0) ++st.term_sky; else if (term == 1) ++st.term_absorbed; else if (term == 2) ++st.term_depth; else ++st.term_light;
    }
    float scale = 1.0f / (float)s.samples_per_pixel;
    for (int c = 0; c < 3; ++c) {
        float mean = scale * acc[c];
        if (linear_out) linear_out[c] = mean;
        if (rgb8_out) rgb8_out[c] = quantise_u8(std::sqrt(mean));
    }
}

}  // namespace rto
