// rtb200_kernels.cu — the resolve kernel, the device-routine probes, and the earlier lane-autonomous trace kernel.
//
// rt_resolve_kernel: adds each pixel's samples in sample order (raytracer.rs:203-205), then scale*sum, sqrt and the u8
// quantisation (raytracer.rs:207-216). Used after every trace launch of the production kernel (rtb200_wavefront.cu).
//
// rt_trace_kernel (RT_VARIANT_LANES, kept for comparison and as a second implementation in the parity tests): every lane
// owns one path and refills itself from the global (pixel,sample) queue; same scene staging by TMA bulk copies, same
// single-level f32 filter + exact f64 confirmation, same albedo stack, no CTA-level sorting, no lights. It was the
// first measured kernel of round 1; the production kernel is rt_wavefront_kernel.
#include "rtb200_kernels.cuh"

using namespace rtd;

namespace rtk {

struct SmemLayout {
    uint32_t filt_off, geo_off, mat_off, cand_off, total;
};
__host__ __device__ inline SmemLayout smem_layout(uint32_t n, uint32_t n_pairs, bool scene_in_smem) {
    SmemLayout L;
    uint32_t off = 16;  // mbarrier
    L.filt_off = off; off += n_pairs * 32u;
    L.geo_off = off;  if (scene_in_smem) off += n * 32u;
    L.mat_off = off;  if (scene_in_smem) off += n * 32u;
    L.cand_off = off; off += (uint32_t)kMaxCand * kBlock * 2u;
    L.total = off;
    return L;
}
size_t trace_smem_bytes(uint32_t n, uint32_t n_pairs, bool scene_in_smem) { return smem_layout(n, n_pairs, scene_in_smem).total; }

RT_DEV void bulk_stage(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    // chunked so a single copy never exceeds 32 KB
    const uint32_t CH = 32768u;
    for (uint32_t o = 0; o < bytes; o += CH) {
        uint32_t nb = bytes - o < CH ? bytes - o : CH;
        tma_bulk_g2s((char*)dst + o, (const char*)src + o, nb, bar);
    }
}

RT_DEV void albedo_of(uint32_t code, const DevMat* mat, float& r, float& g, float& b) {
    if (code & 0x80000000u) {   // packed texel (materials.rs:248-252: pixel as f32 / 255.0)
        r = __fdiv_rn((float)(code & 0xffu), 255.0f);
        g = __fdiv_rn((float)((code >> 8) & 0xffu), 255.0f);
        b = __fdiv_rn((float)((code >> 16) & 0xffu), 255.0f);
    } else {
        const DevMat& m = mat[code];
        r = m.r; g = m.g; b = m.b;
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(kBlock, kCtasPerSm) rt_trace_kernel(const __grid_constant__ TraceParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(p.n, p.n_pairs, p.scene_in_smem != 0);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const float4* s_filt = reinterpret_cast<const float4*>(smem_raw + L.filt_off);
    uint16_t* s_cand = reinterpret_cast<uint16_t*>(smem_raw + L.cand_off);
    const double4* geo = p.scene_in_smem ? reinterpret_cast<const double4*>(smem_raw + L.geo_off) : p.geo;
    const DevMat* mat = p.scene_in_smem ? reinterpret_cast<const DevMat*>(smem_raw + L.mat_off) : p.mat;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned FULL = 0xffffffffu;

    // ---- stage the scene into shared memory (TMA bulk copies, one mbarrier) ----
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        uint32_t bytes = p.n_pairs * 32u + (p.scene_in_smem ? p.n * 64u : 0u);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_stage(smem_raw + L.filt_off, p.filt, p.n_pairs * 32u, bar);
        if (p.scene_in_smem) {
            bulk_stage(smem_raw + L.geo_off, p.geo, p.n * 32u, bar);
            bulk_stage(smem_raw + L.mat_off, p.mat, p.n * 32u, bar);
        }
    }
    mbar_wait(bar, 0);

    const uint32_t gtid = blockIdx.x * kBlock + tid;
    const uint32_t k0 = p.key0, k1 = p.key1;

    bool alive = false, exhausted = false;
    D3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    uint32_t depth_left = 0, level = 0, slot = 0, rays_sample = 0;
    Rng rng; rng_init(rng, 0, 0);
    unsigned long long st_rays = 0, st_cand = 0, st_ovf = 0, st_samples = 0;

    for (;;) {
        // =========================== ray-gen: refill dead lanes ===========================
        unsigned need = __ballot_sync(FULL, !alive && !exhausted);
        if (need) {
            int leader = __ffs(need) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(p.work_counter, (unsigned)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!alive && !exhausted) {
                unsigned my = base + __popc(need & ((1u << lane) - 1u));
                if (my < p.total_work) {
                    uint32_t s_local = my / p.npix_local;
                    uint32_t lp = my - s_local * p.npix_local;
                    uint32_t y_local = lp / p.width, x = lp - y_local * p.width;
                    uint32_t band = y_local / p.band_rows;
                    uint32_t y = (band * (uint32_t)p.world + (uint32_t)p.rank) * p.band_rows + (y_local - band * p.band_rows);
                    slot = my;
                    rng_init(rng, y * p.width + x, p.s0 + s_local);
                    // raytracer.rs:199-200
                    double xi1 = rng_f64(rng, k0, k1);
                    double u = __ddiv_rn(__dadd_rn((double)x, xi1), __dsub_rn((double)p.width, 1.0));
                    double xi2 = rng_f64(rng, k0, k1);
                    double v = __ddiv_rn(__dsub_rn((double)p.height, __dadd_rn((double)y, xi2)), __dsub_rn((double)p.height, 1.0));
                    get_ray(p.cam, u, v, o, d);
                    depth_left = p.max_depth; level = 0; rays_sample = 0;
                    alive = true;
                    ++st_samples;
                    if (depth_left == 0) {   // ray_color(depth = 0) is black without tracing (raytracer.rs:80-82)
                        p.samplebuf[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                        alive = false;
                    }
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(FULL, alive)) break;

        // =========================== closest-hit ===========================
        int nc = 0;
        bool ovf = false;
        const double a = length_squared(d);
        if (!EXACT) {
            // per-ray filter constants in the recentred f32 frame (DESIGN.md "filter soundness")
            float ofx = __double2float_rn(__dsub_rn(o.x, p.gx)), ofy = __double2float_rn(__dsub_rn(o.y, p.gy)),
                  ofz = __double2float_rn(__dsub_rn(o.z, p.gz));
            float dfx = __double2float_rn(d.x), dfy = __double2float_rn(d.y), dfz = __double2float_rn(d.z);
            float s = fmaf(dfx, dfx, fmaf(dfy, dfy, dfz * dfz));
            float oo = fmaf(ofx, ofx, fmaf(ofy, ofy, ofz * ofz));
            bool ok = (s > 1e-30f) && (s < 1e30f) && (oo < 1e30f);
            float inv = rsqrtf(s);
            float dnx = dfx * inv, dny = dfy * inv, dnz = dfz * inv;
            float nod = -fmaf(ofx, dnx, fmaf(ofy, dny, ofz * dnz));
            float thr = __fmul_rd(oo, p.er_coef);
            if (!alive) thr = __int_as_float(0x7fc00000);   // NaN: every comparison is false for dead lanes
            if (alive && !ok) { ovf = true; thr = __int_as_float(0x7fc00000); }
            const float2 dx2 = make_float2(dnx, dnx), dy2 = make_float2(dny, dny), dz2 = make_float2(dnz, dnz);
            const float2 ox2 = make_float2(2.f * ofx, 2.f * ofx), oy2 = make_float2(2.f * ofy, 2.f * ofy),
                         oz2 = make_float2(2.f * ofz, 2.f * ofz);
            const float2 nod2 = make_float2(nod, nod);
            // Blocks of 4 pairs (8 spheres): 8 LDS.128 + 28 FFMA2, branch-free; one rarely-taken branch per block
            // appends the block's candidates (ascending index order is preserved).
            const uint32_t np = p.n_pairs;   // host pads to a multiple of 4 with never-hit records
#pragma unroll 2
            for (uint32_t pp = 0; pp < np; pp += 4) {
                float2 Dv[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 A = s_filt[2 * (pp + q)], B = s_filt[2 * (pp + q) + 1];
                    float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
                    float2 bb = __ffma2_rn(cz, dz2, nod2);
                    float2 tt = __ffma2_rn(cz, oz2, nk);
                    bb = __ffma2_rn(cy, dy2, bb);
                    tt = __ffma2_rn(cy, oy2, tt);
                    bb = __ffma2_rn(cx, dx2, bb);
                    tt = __ffma2_rn(cx, ox2, tt);
                    Dv[q] = __ffma2_rn(bb, bb, tt);
                }
                float m = fmaxf(fmaxf(fmaxf(Dv[0].x, Dv[0].y), fmaxf(Dv[1].x, Dv[1].y)), fmaxf(fmaxf(Dv[2].x, Dv[2].y), fmaxf(Dv[3].x, Dv[3].y)));
                if (m >= thr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (Dv[q].x >= thr) { if (nc < kMaxCand) s_cand[nc * kBlock + tid] = (uint16_t)(2 * (pp + q)); ++nc; }
                        if (Dv[q].y >= thr) { if (nc < kMaxCand) s_cand[nc * kBlock + tid] = (uint16_t)(2 * (pp + q) + 1); ++nc; }
                    }
                }
            }
            if (nc > kMaxCand) ovf = true;
        } else {
            ovf = alive;
        }

        // exact f64 confirmation, ascending sphere index => first index wins ties like raytracer.rs:52-56
        double best_t = DBL_MAX;
        int best = -1;
        if (alive) {
            const int cnt = ovf ? (int)p.n : nc;
            if (ovf) ++st_ovf;
            st_cand += (unsigned)cnt;
            for (int k = 0; k < cnt; ++k) {
                int j = ovf ? k : (int)s_cand[k * kBlock + tid];
                if (j >= (int)p.n) continue;   // padding record of an odd sphere count
                double4 gq = geo[j];
                double root;
                if (sphere_root(mk(gq.x, gq.y, gq.z), gq.w, o, d, a, 0.001, best_t, root)) { best_t = root; best = j; }
            }
        }

        // =========================== shade / scatter ===========================
        // Staged so that lanes of different materials share the expensive pieces: one hit-record stage for all hit
        // lanes, ONE rejection-sampling stage for Lambertian+Texture+Metal lanes, one vector-length stage for
        // Glass+miss lanes.
        if (alive) {
            ++st_rays; ++rays_sample;
            float cr = 0.f, cg = 0.f, cb = 0.f;
            bool done = false;
            const bool hit = best >= 0;
            uint32_t kind = 0xffffffffu;
            DevMat m; m.r = m.g = m.b = 0.f; m.kind = 0; m.param = 0.0; m.tex = -1; m.pad = 0;
            HitRec h; h.point = o; h.normal = d; h.front_face = true;
            D3 center = mk(0, 0, 0);
            if (hit) {
                double4 gq = geo[best];
                center = mk(gq.x, gq.y, gq.z);
                h = hit_record(center, gq.w, o, d, best_t);
                m = mat[best];
                kind = m.kind;
            }
            const bool is_diffuse = kind == RT_LAMBERTIAN || kind == RT_TEXTURE;
            const bool is_metal = kind == RT_METAL, is_glass = kind == RT_GLASS;
            uint32_t code = (uint32_t)best;
            D3 nd = d;
            bool scattered = hit;
            if (is_diffuse || is_metal) {
                D3 rs = random_in_unit_sphere(rng, k0, k1);
                if (is_diffuse) {                                               // materials.rs:84-95, 256-267
                    D3 sd = add(h.normal, rs);
                    if (near_zero(sd)) sd = h.normal;
                    D3 target = add(h.point, sd);
                    nd = sub(target, h.point);
                } else {                                                        // materials.rs:115-129
                    D3 refl = reflect(d, h.normal);
                    nd = add(refl, mul(rs, m.param));
                    if (!(dot(nd, h.normal) > 0.0)) { scattered = false; done = true; }   // absorbed -> black
                }
            }
            if (kind == RT_TEXTURE) {
                double tu, tv;
                sphere_uv(sub(h.point, center), tu, tv);
                code = 0x80000000u | texture_texel(p.tex[m.tex], m.param, tu, tv);
            }
            if (is_glass || !hit) {
                const double len = length(d);                                   // unit_vector: point3d.rs:63-70
                if (is_glass) {                                                 // materials.rs:176-199
                    double ratio = h.front_face ? __ddiv_rn(1.0, m.param) : m.param;
                    D3 ud = divs(d, len);
                    double cos_theta = fmin(dot(neg(ud), h.normal), 1.0);
                    double sin_theta = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(cos_theta, cos_theta)));
                    bool refl = __dmul_rn(ratio, sin_theta) > 1.0;              // cannot_refract
                    if (!refl) refl = reflectance(cos_theta, ratio) > rng_f64(rng, k0, k1);   // drawn only if refraction is possible
                    nd = refl ? reflect(ud, h.normal) : refract(ud, h.normal, ratio);
                } else {                                                        // miss: raytracer.rs:134-163
                    sky_color(d, len, p.sky_mode, p.sky, cr, cg, cb);
                    done = true;
                }
            }
            if (kind == RT_LIGHT) {                                             // materials.rs:65-69
                cr = 1.f; cg = 1.f; cb = 1.f;
                scattered = false; done = true;
            }
            if (scattered) {
                p.stack[(size_t)level * p.stack_stride + gtid] = code;
                ++level;
                --depth_left;
                o = h.point; d = nd;
                if (depth_left == 0) done = true;   // the next ray_color call returns black (raytracer.rs:80-82)
            }
            if (done) {
                // unwind the recursion: c = clamp(light + albedo * c) per level, innermost first (raytracer.rs:117-122)
                if (cr != 0.f || cg != 0.f || cb != 0.f) {
                    for (int l = (int)level - 1; l >= 0; --l) {
                        float ar, ag, ab;
                        albedo_of(p.stack[(size_t)l * p.stack_stride + gtid], mat, ar, ag, ab);
                        cr = clampf(__fadd_rn(0.0f, __fmul_rn(ar, cr)));
                        cg = clampf(__fadd_rn(0.0f, __fmul_rn(ag, cg)));
                        cb = clampf(__fadd_rn(0.0f, __fmul_rn(ab, cb)));
                    }
                }
                p.samplebuf[slot] = make_float4(cr, cg, cb, __uint_as_float(rays_sample));
                alive = false;
            }
        }
    }

    // ---- statistics: one atomic per warp ----
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        st_rays += __shfl_down_sync(FULL, st_rays, off);
        st_cand += __shfl_down_sync(FULL, st_cand, off);
        st_ovf += __shfl_down_sync(FULL, st_ovf, off);
        st_samples += __shfl_down_sync(FULL, st_samples, off);
    }
    if (lane == 0) {
        atomicAdd(&p.stat[0], st_rays);
        atomicAdd(&p.stat[1], st_cand);
        atomicAdd(&p.stat[2], st_ovf);
        atomicAdd(&p.stat[3], st_samples);
    }
}

// Per pixel: add the batch's samples in sample order, f32, exactly like raytracer.rs:197-206; on the last batch
// produce mean = scale*sum (raytracer.rs:207), the linear output and the quantised RGB8 pixel (raytracer.rs:208-216).
__global__ void __launch_bounds__(256) rt_resolve_kernel(const ResolveParams q) {
    uint32_t lp = blockIdx.x * blockDim.x + threadIdx.x;
    if (lp >= q.npix_local) return;
    float r = 0.f, g = 0.f, b = 0.f;
    if (!q.first) { r = q.accum[3 * (size_t)lp]; g = q.accum[3 * (size_t)lp + 1]; b = q.accum[3 * (size_t)lp + 2]; }
    for (uint32_t s = 0; s < q.s_count; ++s) {
        float4 c = q.samplebuf[(size_t)s * q.npix_local + lp];
        r = __fadd_rn(r, c.x); g = __fadd_rn(g, c.y); b = __fadd_rn(b, c.z);
    }
    if (!q.last) {
        q.accum[3 * (size_t)lp] = r; q.accum[3 * (size_t)lp + 1] = g; q.accum[3 * (size_t)lp + 2] = b;
        return;
    }
    float scale = __fdiv_rn(1.0f, (float)q.spp);
    float mr = __fmul_rn(scale, r), mg = __fmul_rn(scale, g), mb = __fmul_rn(scale, b);
    if (q.out_linear) { q.out_linear[3 * (size_t)lp] = mr; q.out_linear[3 * (size_t)lp + 1] = mg; q.out_linear[3 * (size_t)lp + 2] = mb; }
    if (q.out_rgb8) {
        q.out_rgb8[3 * (size_t)lp] = quantise_u8(mr);
        q.out_rgb8[3 * (size_t)lp + 1] = quantise_u8(mg);
        q.out_rgb8[3 * (size_t)lp + 2] = quantise_u8(mb);
    }
}

cudaError_t trace_configure(int device, int* sm_count, size_t* max_smem_optin) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return e;
    *sm_count = prop.multiProcessorCount;
    *max_smem_optin = prop.sharedMemPerBlockOptin;
    return cudaSuccess;
}

cudaError_t launch_trace(const TraceParams& p, int grid, size_t smem, bool exact, cudaStream_t st) {
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(rt_trace_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rt_trace_kernel<true><<<grid, kBlock, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(rt_trace_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rt_trace_kernel<false><<<grid, kBlock, smem, st>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_resolve(const ResolveParams& q, cudaStream_t st) {
    int grid = (int)((q.npix_local + 255u) / 256u);
    rt_resolve_kernel<<<grid, 256, 0, st>>>(q);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// probes: the same device routines on one thread (known-answer tests of the reference, see include/rtb200.h)
// ---------------------------------------------------------------------------------------------------
__global__ void k_probe_sphere_hit(const double* in, double* out) {
    D3 c = mk(in[0], in[1], in[2]); double r = in[3];
    D3 o = mk(in[4], in[5], in[6]), d = mk(in[7], in[8], in[9]);
    double root = 0.0;
    bool ok = sphere_root(c, r, o, d, length_squared(d), in[10], in[11], root);
    out[0] = ok ? 1.0 : 0.0;
    if (ok) {
        HitRec h = hit_record(c, r, o, d, root);
        out[1] = root; out[2] = h.point.x; out[3] = h.point.y; out[4] = h.point.z;
        out[5] = h.normal.x; out[6] = h.normal.y; out[7] = h.normal.z; out[8] = h.front_face ? 1.0 : 0.0;
    }
}
__global__ void k_probe_refract(const double* in, double* out) {
    D3 r = refract(mk(in[0], in[1], in[2]), mk(in[3], in[4], in[5]), in[6]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
__global__ void k_probe_reflectance(const double* in, double* out) { out[0] = reflectance(in[0], in[1]); }
__global__ void k_probe_sky(const double* in, uint32_t mode, float* out) {
    DevTex none; none.rgb8 = nullptr; none.width = 0; none.height = 0;
    D3 dd = mk(in[0], in[1], in[2]);
    sky_color(dd, length(dd), mode, none, out[0], out[1], out[2]);
}
__global__ void k_probe_get_ray(const rt_camera* cam, const double* in, double* out) {
    D3 o, d;
    get_ray(*cam, in[0], in[1], o, d);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
}
__global__ void k_probe_rng(uint32_t k0, uint32_t k1, uint32_t pixel, uint32_t sample, uint32_t kind, uint32_t n, double* out) {
    Rng g; rng_init(g, pixel, sample);
    for (uint32_t i = 0; i < n; ++i) out[i] = kind == 0 ? rng_f64(g, k0, k1) : rng_m1_1(g, k0, k1);
}
__global__ void k_probe_quantise(const float* in, uint32_t n, uint8_t* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = quantise_u8(in[i]);
}

__global__ void k_probe_sphere_uv(const double* in, uint32_t n, double* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sphere_uv(mk(in[3 * i], in[3 * i + 1], in[3 * i + 2]), out[2 * i], out[2 * i + 1]);
}
cudaError_t probe_sphere_uv(const double* in, uint32_t n, double* out, cudaStream_t st) {
    k_probe_sphere_uv<<<(n + 255) / 256, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

cudaError_t probe_sphere_hit(const double* in, double* out, cudaStream_t st) { k_probe_sphere_hit<<<1, 1, 0, st>>>(in, out); return cudaGetLastError(); }
cudaError_t probe_refract(const double* in, double* out, cudaStream_t st) { k_probe_refract<<<1, 1, 0, st>>>(in, out); return cudaGetLastError(); }
cudaError_t probe_reflectance(const double* in, double* out, cudaStream_t st) { k_probe_reflectance<<<1, 1, 0, st>>>(in, out); return cudaGetLastError(); }
cudaError_t probe_sky(const double* in, uint32_t mode, float* out, cudaStream_t st) { k_probe_sky<<<1, 1, 0, st>>>(in, mode, out); return cudaGetLastError(); }
cudaError_t probe_get_ray(const rt_camera* cam, const double* in, double* out, cudaStream_t st) { k_probe_get_ray<<<1, 1, 0, st>>>(cam, in, out); return cudaGetLastError(); }
cudaError_t probe_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t kind, uint32_t n, double* out, cudaStream_t st) {
    k_probe_rng<<<1, 1, 0, st>>>((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, kind, n, out);
    return cudaGetLastError();
}
cudaError_t probe_quantise(const float* in, uint32_t n, uint8_t* out, cudaStream_t st) {
    k_probe_quantise<<<(n + 255) / 256, 256, 0, st>>>(in, n, out);
    return cudaGetLastError();
}

}  // namespace rtk
