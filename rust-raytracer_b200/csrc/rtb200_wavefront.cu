// rtb200_wavefront.cu — the production trace kernel: a persistent-threads WAVEFRONT tracer at CTA scope.
//
// One persistent CTA (256 threads) per resident slot of every SM owns a pool of 256 ray slots in shared memory,
// laid out SoA, next to the scene's sphere/material records (staged once by 1-D TMA bulk copies). Until the global
// (pixel,sample) queue is drained and the pool is empty, the CTA repeats three barrier-separated stages:
//
//   closest-hit   thread t <-> slot t. hit_world (raytracer.rs:44-59) over ALL spheres as a conservative f32 filter
//                 (7 FMA per sphere, two spheres per packed FFMA2, blocks of 8 spheres without a branch) that appends
//                 candidates to a per-thread list, then the reference-exact f64 Sphere::hit (sphere.rs:46-78) on the
//                 candidates only (ascending index: ties go to the first sphere like raytracer.rs:52-56).
//   sort          rays are classified {miss, diffuse, metal, glass, light} and compacted class by class with warp
//                 ballots + one shared-memory atomic per (warp, class): perm[] lists the live slots sorted by class.
//   shade+ray-gen thread i <-> slot perm[i], so a warp shades ONE material: Material::scatter (materials.rs:44-54) or
//                 the sky (raytracer.rs:134-163), iteratively (no recursion): albedo codes go to a per-slot stack that
//                 is unwound backwards on termination so the f32 products associate exactly like the reference's
//                 recursion (raytracer.rs:117-122). A terminated path writes its sample and the same thread
//                 immediately regenerates the slot from the queue (render_line's jitter + Camera::get_ray,
//                 raytracer.rs:199-201): one warp-aggregated atomic pops the work items.
//
// FMA-pipe work (closest-hit) of one CTA overlaps FP64-pipe work (shade) of the other CTAs resident on the SM.
#include "rtb200_kernels.cuh"

using namespace rtd;

#ifndef RT_L1_GUARD
#define RT_L1_GUARD 0
#endif
#ifndef RT_SAMPLE_ILP
#define RT_SAMPLE_ILP 1   // two rejection trials per trip with their Philox blocks computed together (bit-identical stream)
#endif
#ifndef RT_CONFIRM_ILP
#define RT_CONFIRM_ILP 0   // two candidates per trip with interleaved f64 chains: bit-identical, but measured 5 % SLOWER (wasted sqrt/div on misses)
#endif
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 3   // albedo-stack levels kept in shared memory per slot (deeper levels live in global memory)
#endif

namespace rtk {

namespace {

enum : uint32_t { CLS_MISS = 0, CLS_DIFFUSE = 1, CLS_METAL = 2, CLS_GLASS = 3, CLS_LIGHT = 4, CLS_DEAD = 5, N_CLS = 6 };
constexpr uint32_t kDeadLevel = 0xffffffffu;

struct WfSmem {
    uint32_t filt_off, sfilt_off, orig_off, cmeta_off, geo_off, mat_off, l1_off, l2_off;
    uint32_t ox, oy, oz, dx, dy, dz, bt;            // double[kBlock] each
    uint32_t bi, work, pix, smp, blk, clo, chi, lvl, shd;   // uint32[kBlock] each
    uint32_t perm;                                  // uint16[kBlock]
    uint32_t stk;                                   // uint32[RT_SMEM_STACK][kBlock]: first levels of the albedo stack
    uint32_t cnt;                                   // uint32[2][8]
    uint32_t flags;                                 // uint32[4]
    uint32_t total;
};

// mask: bit0 second-level sphere records (+slot map), bit1 exact geometry, bit2 materials in shared memory
__host__ __device__ inline WfSmem wf_layout(uint32_t n, uint32_t n_pairs, uint32_t n_clusters, bool two_level, uint32_t mask, uint32_t kBlock) {
    WfSmem L;
    uint32_t off = 16;   // mbarrier
    L.filt_off = off; off += n_pairs * 32u;
    L.sfilt_off = off; if (two_level && (mask & 1u)) off += n_clusters * (uint32_t)(kClusterK * 16);
    L.orig_off = off; if (two_level && (mask & 1u)) off += n_clusters * (uint32_t)(kClusterK * 2);
    L.cmeta_off = off; if (two_level && (mask & 1u)) off += n_clusters * 4u;
    off = (off + 15u) & ~15u;
    L.geo_off = off; if (mask & 2u) off += n * 32u;
    L.mat_off = off; if (mask & 4u) off += n * 32u;
    L.l1_off = off; off += (uint32_t)kWfMaxClus * kBlock * 2u;
    L.l2_off = off; if (two_level) off += (uint32_t)kWfMaxCand * kBlock * 2u;
    off = (off + 15u) & ~15u;
    L.ox = off; off += kBlock * 8u; L.oy = off; off += kBlock * 8u; L.oz = off; off += kBlock * 8u;
    L.dx = off; off += kBlock * 8u; L.dy = off; off += kBlock * 8u; L.dz = off; off += kBlock * 8u;
    L.bt = off; off += kBlock * 8u;
    L.bi = off; off += kBlock * 4u; L.work = off; off += kBlock * 4u; L.pix = off; off += kBlock * 4u; L.smp = off; off += kBlock * 4u;
    L.blk = off; off += kBlock * 4u; L.clo = off; off += kBlock * 4u; L.chi = off; off += kBlock * 4u; L.lvl = off; off += kBlock * 4u; L.shd = off; off += kBlock * 4u;
    L.perm = off; off += kBlock * 2u;
    L.stk = off; off += (uint32_t)RT_SMEM_STACK * kBlock * 4u;
    L.cnt = off; off += 2u * 8u * 4u;
    L.flags = off; off += 16u;
    L.total = off;
    return L;
}

RT_DEV void bulk_stage(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    const uint32_t CH = 32768u;
    for (uint32_t o = 0; o < bytes; o += CH) {
        uint32_t nb = bytes - o < CH ? bytes - o : CH;
        tma_bulk_g2s((char*)dst + o, (const char*)src + o, nb, bar);
    }
}

RT_DEV void albedo_of(uint32_t code, const DevMat* mat, float& r, float& g, float& b) {
    if (code == 0xffffffffu) { r = g = b = 1.0f; return; }   // Light: Srgb(1,1,1) (materials.rs:67)
    if (code & 0x80000000u) {   // packed texel (materials.rs:248-252: pixel as f32 / 255.0)
        r = __fdiv_rn((float)(code & 0xffu), 255.0f);
        g = __fdiv_rn((float)((code >> 8) & 0xffu), 255.0f);
        b = __fdiv_rn((float)((code >> 16) & 0xffu), 255.0f);
    } else {
        const DevMat& m = mat[code];
        r = m.r; g = m.g; b = m.b;
    }
}

}  // namespace

size_t wavefront_smem_bytes(uint32_t n, uint32_t n_pairs, uint32_t n_clusters, bool two_level, uint32_t smem_mask, int block) {
    return wf_layout(n, n_pairs, n_clusters, two_level, smem_mask, (uint32_t)block).total;
}

template <int kBlock, int MINB, bool EXACT, bool LIGHTS, bool TWO>
__global__ void __launch_bounds__(kBlock, MINB) rt_wavefront_kernel(const __grid_constant__ TraceParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WfSmem L = wf_layout(p.n, p.n_pairs, p.n_clusters, TWO, p.scene_in_smem, kBlock);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const float4* s_filt = reinterpret_cast<const float4*>(smem_raw + L.filt_off);
    uint16_t* s_l1 = reinterpret_cast<uint16_t*>(smem_raw + L.l1_off);     // first-level candidates (clusters, or spheres without clustering)
    uint16_t* s_l2 = reinterpret_cast<uint16_t*>(smem_raw + L.l2_off);     // second-level candidates (spheres), two-level mode
    const float4* sfilt = (TWO && (p.scene_in_smem & 1u)) ? reinterpret_cast<const float4*>(smem_raw + L.sfilt_off) : p.sfilt;
    const uint16_t* orig = (TWO && (p.scene_in_smem & 1u)) ? reinterpret_cast<const uint16_t*>(smem_raw + L.orig_off) : p.orig;
    const float* cmeta = (TWO && (p.scene_in_smem & 1u)) ? reinterpret_cast<const float*>(smem_raw + L.cmeta_off) : p.cmeta;   // |c| of each cluster bound
    const double4* geo = (p.scene_in_smem & 2u) ? reinterpret_cast<const double4*>(smem_raw + L.geo_off) : p.geo;
    const DevMat* mat = (p.scene_in_smem & 4u) ? reinterpret_cast<const DevMat*>(smem_raw + L.mat_off) : p.mat;
    double* s_ox = reinterpret_cast<double*>(smem_raw + L.ox); double* s_oy = reinterpret_cast<double*>(smem_raw + L.oy);
    double* s_oz = reinterpret_cast<double*>(smem_raw + L.oz); double* s_dx = reinterpret_cast<double*>(smem_raw + L.dx);
    double* s_dy = reinterpret_cast<double*>(smem_raw + L.dy); double* s_dz = reinterpret_cast<double*>(smem_raw + L.dz);
    double* s_bt = reinterpret_cast<double*>(smem_raw + L.bt);
    uint32_t* s_bi = reinterpret_cast<uint32_t*>(smem_raw + L.bi); uint32_t* s_work = reinterpret_cast<uint32_t*>(smem_raw + L.work);
    uint32_t* s_pix = reinterpret_cast<uint32_t*>(smem_raw + L.pix); uint32_t* s_smp = reinterpret_cast<uint32_t*>(smem_raw + L.smp);
    uint32_t* s_blk = reinterpret_cast<uint32_t*>(smem_raw + L.blk); uint32_t* s_clo = reinterpret_cast<uint32_t*>(smem_raw + L.clo);
    uint32_t* s_chi = reinterpret_cast<uint32_t*>(smem_raw + L.chi); uint32_t* s_lvl = reinterpret_cast<uint32_t*>(smem_raw + L.lvl);
    uint32_t* s_shd = reinterpret_cast<uint32_t*>(smem_raw + L.shd);   // depth of the shadow-frame stack (0 = main path)
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(smem_raw + L.perm);
    uint32_t* s_stk = reinterpret_cast<uint32_t*>(smem_raw + L.stk);
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw + L.cnt);
    volatile uint32_t* s_flags = reinterpret_cast<volatile uint32_t*>(smem_raw + L.flags);   // [0] = queue exhausted

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned FULL = 0xffffffffu;
    const uint32_t k0 = p.key0, k1 = p.key1;
    const uint32_t stack_col = blockIdx.x * kBlock;   // this CTA's columns of the albedo stack
    auto stack_push = [&](uint32_t lvl, uint32_t s, uint32_t code) {
        if (lvl < (uint32_t)RT_SMEM_STACK) s_stk[lvl * kBlock + s] = code; else p.stack[(size_t)lvl * p.stack_stride + stack_col + s] = code;
    };
    auto stack_get = [&](uint32_t lvl, uint32_t s) -> uint32_t {
        return lvl < (uint32_t)RT_SMEM_STACK ? s_stk[lvl * kBlock + s] : p.stack[(size_t)lvl * p.stack_stride + stack_col + s];
    };

    // ---- stage the scene into shared memory (TMA bulk copies, one mbarrier) ----
    if (tid == 0) mbar_init(bar, 1);
    if (tid < 16) s_cnt[tid] = 0;
    if (tid < 4) s_flags[tid] = 0;
    s_lvl[tid] = kDeadLevel;
    __syncthreads();
    if (tid == 0) {
        const bool sf = TWO && (p.scene_in_smem & 1u);
        uint32_t bytes = p.n_pairs * 32u + (sf ? p.n_clusters * (uint32_t)(kClusterK * 18 + 4) : 0u) + ((p.scene_in_smem & 2u) ? p.n * 32u : 0u) + ((p.scene_in_smem & 4u) ? p.n * 32u : 0u);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_stage(smem_raw + L.filt_off, p.filt, p.n_pairs * 32u, bar);
        if (sf) {
            bulk_stage(smem_raw + L.sfilt_off, p.sfilt, p.n_clusters * (uint32_t)(kClusterK * 16), bar);
            bulk_stage(smem_raw + L.orig_off, p.orig, p.n_clusters * (uint32_t)(kClusterK * 2), bar);
            bulk_stage(smem_raw + L.cmeta_off, p.cmeta, p.n_clusters * 4u, bar);
        }
        if (p.scene_in_smem & 2u) bulk_stage(smem_raw + L.geo_off, p.geo, p.n * 32u, bar);
        if (p.scene_in_smem & 4u) bulk_stage(smem_raw + L.mat_off, p.mat, p.n * 32u, bar);
    }
    mbar_wait(bar, 0);
#ifdef RT_DEBUG_STAGE
    {   // verify the TMA staging of the first-level records against global memory
        unsigned long long bad = 0;
        const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(p.filt);
        const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(smem_raw + L.filt_off);
        for (uint32_t i = tid; i < p.n_pairs * 8u; i += kBlock) if (gsrc[i] != ssrc[i]) ++bad;
        if (bad) atomicAdd(&p.stat[6], bad);
    }
#endif

    unsigned long long st_rays = 0, st_cand = 0, st_ovf = 0, st_samples = 0, st_clus = 0;
#ifdef RT_PROFILE_PHASES
    unsigned long long pf_scan = 0, pf_confirm = 0, pf_waitA = 0, pf_sort = 0, pf_shade = 0, pf_waitC = 0, pf_iters = 0, pf_t = clock64();
#define PF_MARK(acc) { unsigned long long now_ = clock64(); acc += now_ - pf_t; pf_t = now_; }
#else
#define PF_MARK(acc)
#endif

    // Regenerate slot `s` from the global (pixel,sample) queue. Warp-synchronous: every lane of the warp calls it,
    // `want` says whether this lane's slot needs a new path. raytracer.rs:199-201 + camera.rs:79-84.
    auto regenerate = [&](bool want, uint32_t s) {
        want = want && (s_flags[0] == 0u);
        unsigned need = __ballot_sync(FULL, want);
        if (!need) return;
        int leader = __ffs(need) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(p.work_counter, (unsigned)__popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (!want) return;
        unsigned my = base + __popc(need & ((1u << lane) - 1u));
        if (my >= p.total_work) { s_flags[0] = 1u; return; }
        uint32_t s_local = my / p.npix_local;
        uint32_t lp = my - s_local * p.npix_local;
        uint32_t y_local = lp / p.width, x = lp - y_local * p.width;
        uint32_t band = y_local / p.band_rows;
        uint32_t y = (band * (uint32_t)p.world + (uint32_t)p.rank) * p.band_rows + (y_local - band * p.band_rows);
        Rng rng; rng_init(rng, y * p.width + x, p.s0 + s_local);
        double xi1 = rng_f64(rng, k0, k1);
        double u = __ddiv_rn(__dadd_rn((double)x, xi1), __dsub_rn((double)p.width, 1.0));
        double xi2 = rng_f64(rng, k0, k1);
        double v = __ddiv_rn(__dsub_rn((double)p.height, __dadd_rn((double)y, xi2)), __dsub_rn((double)p.height, 1.0));
        D3 o, d;
        get_ray(p.cam, u, v, o, d);
        s_ox[s] = o.x; s_oy[s] = o.y; s_oz[s] = o.z; s_dx[s] = d.x; s_dy[s] = d.y; s_dz[s] = d.z;
        s_work[s] = my; s_pix[s] = rng.pixel; s_smp[s] = rng.sample;
        s_blk[s] = (rng.blk << 1) | rng.has; s_clo[s] = rng.c_lo; s_chi[s] = rng.c_hi;
        s_lvl[s] = 0u;
        if (LIGHTS) {
            s_shd[s] = 0u;
            for (int q = 0; q < 6; ++q) p.lterm[(size_t)q * p.stack_stride + stack_col + s] = 0.0f;
        }
        ++st_samples;
    };

    regenerate(true, (uint32_t)tid);   // initial fill of the pool
    __syncthreads();

    uint32_t it = 0;
    for (;; ++it) {
        uint32_t* cnt = s_cnt + (it & 1u) * 8u;
        // =========================== closest-hit: thread t <-> slot t ===========================
        const bool alive = s_lvl[tid] != kDeadLevel;
        uint32_t cls = CLS_DEAD;
        {
            const D3 o = mk(s_ox[tid], s_oy[tid], s_oz[tid]), d = mk(s_dx[tid], s_dy[tid], s_dz[tid]);
            bool ovf = false;
            const double a = length_squared(d);
            // exact f64 confirmation. hit_world (raytracer.rs:44-59) keeps the closest root and, on equal t, the first sphere in
            // list order; because Sphere::hit(t_max) accepts exactly r < t_max with r the first root beyond t_min, that fold equals
            // the lexicographic minimum of (r, index) over all spheres - so candidates may be confirmed in any order.
            double best_t = DBL_MAX;
            int best = -1;
            auto confirm = [&](int j) {
                if (j >= (int)p.n) return;   // padding record
                double4 gq = geo[j];
                double root;
                if (sphere_root(mk(gq.x, gq.y, gq.z), gq.w, o, d, a, 0.001, DBL_MAX, root)) {
                    if (best < 0 || root < best_t || (root == best_t && j < best)) { best_t = root; best = j; }
                }
                ++st_cand;
            };
            auto consider = [&](int j, bool hit, double root) {
                if (hit && (best < 0 || root < best_t || (root == best_t && j < best))) { best_t = root; best = j; }
            };
            // the candidates of a list, two at a time (interleaved f64 chains), then the odd one
            auto confirm_list = [&](const uint16_t* cl, int cnt) {
                int j = 0;
                if (RT_CONFIRM_ILP) {
                    for (; j + 1 < cnt; j += 2) {
                        const int j0 = (int)cl[j * kBlock + tid], j1 = (int)cl[(j + 1) * kBlock + tid];
                        if (j0 >= (int)p.n || j1 >= (int)p.n) { confirm(j0); confirm(j1); continue; }   // padding record in the pair
                        const double4 g0 = geo[j0], g1 = geo[j1];
                        bool h0, h1; double r0 = 0.0, r1 = 0.0;
                        sphere_root2(mk(g0.x, g0.y, g0.z), g0.w, mk(g1.x, g1.y, g1.z), g1.w, o, d, a, 0.001, h0, r0, h1, r1);
                        consider(j0, h0, r0); consider(j1, h1, r1);
                        st_cand += 2;
                    }
                }
                for (; j < cnt; ++j) confirm((int)cl[j * kBlock + tid]);
            };
            const bool warp_has_ray = __ballot_sync(FULL, alive) != 0u;   // a warp whose 32 slots are all empty skips the scan (frame tail)
            if (!EXACT && warp_has_ray) {
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
                if (!alive) thr = __int_as_float(0x7fc00000);   // NaN: every comparison is false, so empty slots (whose o,d are garbage, possibly +-inf) never produce candidates
                if (alive && !ok) { ovf = true; thr = __int_as_float(0x7fc00000); }
                const float2 dx2 = make_float2(dnx, dnx), dy2 = make_float2(dny, dny), dz2 = make_float2(dnz, dnz);
                const float2 ox2 = make_float2(2.f * ofx, 2.f * ofx), oy2 = make_float2(2.f * ofy, 2.f * ofy),
                             oz2 = make_float2(2.f * ofz, 2.f * ofz);
                const float2 nod2 = make_float2(nod, nod);
                const uint32_t np = p.n_pairs;   // multiple of 8; padding records never hit
                // Per-thread candidate lists live in shared memory as columns (entry k of thread t at [k*kBlock + t]); appends are
                // predicated stores through a 32-bit shared address. A list that is about to fill up is drained on the spot, so
                // no ray ever falls back to brute force because of list capacity.
                const uint32_t l1_base = smem_u32(s_l1 + tid);
                const uint32_t l1_full = l1_base + (uint32_t)(kWfMaxClus - 8) * kBlock * 2u;
                uint32_t l1_addr = l1_base;
                const uint32_t l2_base = smem_u32(s_l2 + tid);
                const uint32_t l2_full = l2_base + (uint32_t)(kWfMaxCand - kClusterK) * kBlock * 2u;
                uint32_t l2_addr = l2_base;
#define RT_FILTER_PAIRS(REC, DV, NP)                                                                                                  \
    _Pragma("unroll") for (int q = 0; q < (NP); ++q) {                                                                              \
        float4 A = (REC)[2 * q], B = (REC)[2 * q + 1];                                                                               \
        float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);       \
        float2 bb = __ffma2_rn(cz, dz2, nod2);                                                                                       \
        float2 tt = __ffma2_rn(cz, oz2, nk);                                                                                         \
        bb = __ffma2_rn(cy, dy2, bb);                                                                                                \
        tt = __ffma2_rn(cy, oy2, tt);                                                                                                \
        bb = __ffma2_rn(cx, dx2, bb);                                                                                                \
        tt = __ffma2_rn(cx, ox2, tt);                                                                                                \
        (DV)[q] = __ffma2_rn(bb, bb, tt);                                                                                            \
    }
#define RT_APPEND_IF(ADDR, VAL, ID)                                                                                                   \
    asm volatile("{\n.reg .pred p;\n.reg .b16 h;\nsetp.ge.f32 p, %1, %2;\ncvt.u16.u32 h, %3;\n@p st.shared.u16 [%0], h;\n@p add.u32 %0, %0, %4;\n}" \
                 : "+r"(ADDR) : "f"(VAL), "f"(thr), "r"(ID), "n"(kBlock * 2) : "memory")
                // Loop structure: each level fills its list until it is nearly full (or its input ends), the next level drains
                // it, and the outer loop resumes. Every drain therefore exists exactly once in the code (small I-cache footprint)
                // and no ray ever falls back to brute force because of list capacity.
                uint32_t pp = 0;
                for (;;) {
                    // ---- level 1: records of `filt` (cluster bounds, or the spheres themselves without clustering) ----
#pragma unroll 1
                    for (; pp < np && l1_addr <= l1_full; pp += 4) {
                        float2 Dv[4];
                        const float4* rec = s_filt + 2 * pp;
                        RT_FILTER_PAIRS(rec, Dv, 4)
                        // Without clustering a block rarely holds a candidate, so one max-reduction + branch guards the appends; with
                        // clustering some lane of the warp hits nearly every block of cluster bounds, so the guard is dropped.
                        bool any_hit = true;
                        if (!TWO || RT_L1_GUARD) {
                            float m = fmaxf(fmaxf(fmaxf(Dv[0].x, Dv[0].y), fmaxf(Dv[1].x, Dv[1].y)), fmaxf(fmaxf(Dv[2].x, Dv[2].y), fmaxf(Dv[3].x, Dv[3].y)));
                            any_hit = m >= thr;
                        }
                        if (any_hit) {
                            const uint32_t j0 = 2u * pp;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                RT_APPEND_IF(l1_addr, Dv[q].x, j0 + 2u * q);
                                RT_APPEND_IF(l1_addr, Dv[q].y, j0 + 2u * q + 1u);
                            }
                        }
                    }
                    if (TWO) {
                        // ---- level 2: the kClusterK member spheres of every listed cluster (per-thread addresses) ----
                        const int n1 = (int)((l1_addr - l1_base) / (kBlock * 2u));
                        const float beta_o = 16.0f * 5.9604645e-8f * sqrtf(oo);      // rounding bound of b = (c-o).d on the ray's side
                        int k = 0;
                        for (;;) {
#pragma unroll 1
                            for (; k < n1 && l2_addr <= l2_full; ++k) {
                                const uint32_t cb = s_l1[k * kBlock + tid];
                                if (cb >= p.n_clusters) continue;                   // padding record
                                // Behind-the-origin cull: with b = (c-o).d^ and q = |c-o|^2 - R^2, b < 0 and q > 0 put both roots of
                                // the bounding sphere at negative t, hence every member root too. Both signs are required beyond
                                // their rounding margins, otherwise the cluster is processed.
                                const uint32_t pr = cb >> 1, hi = cb & 1u;
                                const float4 CA = s_filt[2 * pr], CB = s_filt[2 * pr + 1];
                                const float ccx = hi ? CA.y : CA.x, ccy = hi ? CA.w : CA.z, ccz = hi ? CB.y : CB.x, cnk = hi ? CB.w : CB.z;
                                const float cbb = fmaf(ccx, dnx, fmaf(ccy, dny, fmaf(ccz, dnz, nod)));
                                const float ctt = fmaf(ccx, 2.f * ofx, fmaf(ccy, 2.f * ofy, fmaf(ccz, 2.f * ofz, cnk)));
                                const float cD = fmaf(cbb, cbb, ctt);               // ~ b^2 - q + Es + |o|^2
                                const float cabs = cmeta[cb];                       // |c| of the bound (f32, rounded up)
                                const float qlow = fmaf(cbb, cbb, oo) - cD;         // ~ q - Es  (<= q up to rounding)
                                const float eq = 2.0e-5f * fmaf(cabs, cabs, oo);    // >= 3x the rounding bound 96u(|c|^2+|o|^2)
                                if (cbb < -(16.0f * 5.9604645e-8f * cabs + beta_o) * 1.5f - 1e-30f && qlow > eq) continue;
                                float2 Dv[kClusterK / 2];
                                const float4* rec = sfilt + (uint32_t)kClusterK * cb;
                                RT_FILTER_PAIRS(rec, Dv, kClusterK / 2)
                                float m = fmaxf(Dv[0].x, Dv[0].y);
#pragma unroll
                                for (int q = 1; q < kClusterK / 2; ++q) m = fmaxf(m, fmaxf(Dv[q].x, Dv[q].y));
                                if (m >= thr) {
                                    const uint16_t* og = orig + (uint32_t)kClusterK * cb;
#pragma unroll
                                    for (int q = 0; q < kClusterK / 2; ++q) {
                                        const uint32_t w = *reinterpret_cast<const uint32_t*>(og + 2 * q);   // 2 x u16 slot -> sphere index
                                        RT_APPEND_IF(l2_addr, Dv[q].x, w & 0xffffu);
                                        RT_APPEND_IF(l2_addr, Dv[q].y, w >> 16);
                                    }
                                }
                                ++st_clus;
                            }
                            // ---- exact f64 confirmation of the listed spheres ----
                            const int n2 = (int)((l2_addr - l2_base) / (kBlock * 2u));
                            confirm_list(s_l2, n2);
                            l2_addr = l2_base;
                            if (k >= n1) break;
                        }
                        l1_addr = l1_base;
                    } else {
                        const int n1 = (int)((l1_addr - l1_base) / (kBlock * 2u));
                        confirm_list(s_l1, n1);
                        l1_addr = l1_base;
                    }
                    if (pp >= np) break;
                }
                PF_MARK(pf_scan)
#undef RT_FILTER_PAIRS
#undef RT_APPEND_IF
            } else if (EXACT) {
                ovf = alive;
            }
            if (alive) {
                if (ovf) {   // EXACT variant, or a ray outside the f32 filter's safe range: every sphere in f64
                    ++st_ovf;
                    for (int k = 0; k < (int)p.n; ++k) confirm(k);
                }
                s_bt[tid] = best_t;
                s_bi[tid] = (uint32_t)best;
                cls = CLS_MISS;
                if (best >= 0) {
                    uint32_t kind = mat[best].kind;
                    cls = (kind == RT_METAL) ? CLS_METAL : (kind == RT_GLASS) ? CLS_GLASS : (kind == RT_LIGHT) ? CLS_LIGHT : CLS_DIFFUSE;
                }
                ++st_rays;
            }
        }
        PF_MARK(pf_confirm)

        // =========================== sort: compact the live slots class by class ===========================
        uint32_t wbase = 0, rank = 0;
#pragma unroll
        for (uint32_t c = 0; c < CLS_DEAD; ++c) {
            unsigned b = __ballot_sync(FULL, cls == c);
            if (b) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt[c], (uint32_t)__popc(b));
                base = __shfl_sync(FULL, base, 0);
                if (cls == c) { wbase = base; rank = __popc(b & ((1u << lane) - 1u)); }
            }
        }
        PF_MARK(pf_sort)
        __syncthreads();   // A: class counts complete
        PF_MARK(pf_waitA)
        uint32_t c0 = cnt[0], c1 = cnt[1], c2 = cnt[2], c3 = cnt[3], c4 = cnt[4];
        const uint32_t e0 = c0, e1 = e0 + c1, e2 = e1 + c2, e3 = e2 + c3, n_live = e3 + c4;   // class end offsets
        if (cls != CLS_DEAD) {
            uint32_t start = cls == 0 ? 0u : cls == 1 ? e0 : cls == 2 ? e1 : cls == 3 ? e2 : e3;
            s_perm[start + wbase + rank] = (uint16_t)tid;
        }
        if (tid < 8) s_cnt[((it + 1u) & 1u) * 8u + tid] = 0u;   // reset the other counter set for the next iteration
        __syncthreads();   // B: perm complete
        PF_MARK(pf_sort)

        // =========================== shade + regenerate: thread i <-> slot perm[i] ===========================
        bool still_alive = false;
        {
            const bool active = (uint32_t)tid < n_live;
            const uint32_t s = active ? (uint32_t)s_perm[tid] : 0u;
            const uint32_t c = !active ? CLS_DEAD : ((uint32_t)tid < e0 ? CLS_MISS : (uint32_t)tid < e1 ? CLS_DIFFUSE : (uint32_t)tid < e2 ? CLS_METAL : (uint32_t)tid < e3 ? CLS_GLASS : CLS_LIGHT);
            bool done = false;
            if (active) {
                D3 o = mk(s_ox[s], s_oy[s], s_oz[s]), d = mk(s_dx[s], s_dy[s], s_dz[s]);
                uint32_t level = s_lvl[s];
                uint32_t shd = LIGHTS ? s_shd[s] : 0u;       // > 0: this ray is a shadow ray of the light test (raytracer.rs:103-106)
                const uint32_t rays_sample = level + 1u;   // main-path hit_world calls so far, this one included
                float cr = 0.f, cg = 0.f, cb = 0.f;
                bool have_tc = false;                      // a shadow ray's ray_color(.., 2, 1) value is ready
                float tr = 0.f, tg = 0.f, tb = 0.f;
                bool state_dirty = false;                  // o/d/rng/level must be written back to the pool
                Rng rng; rng.pixel = s_pix[s]; rng.sample = s_smp[s];
                { uint32_t bh = s_blk[s]; rng.blk = bh >> 1; rng.has = bh & 1u; }
                rng.c_lo = s_clo[s]; rng.c_hi = s_chi[s];
                if (c == CLS_MISS) {                                                // raytracer.rs:134-163
                    float x, y, z;
                    sky_color(d, length(d), p.sky_mode, p.sky, x, y, z);
                    if (shd == 0u) { cr = x; cg = y; cb = z; done = true; }
                    else { tr = x; tg = y; tb = z; have_tc = true; }
                } else {
                    const uint32_t best = s_bi[s];
                    const double best_t = s_bt[s];
                    double4 gq = geo[best];
                    const D3 center = mk(gq.x, gq.y, gq.z);
                    HitRec h = hit_record(center, gq.w, o, d, best_t);
                    const DevMat m = mat[best];
                    uint32_t code = best;
                    D3 nd = d;
                    bool absorbed = false;
                    const bool is_light = (c == CLS_LIGHT);                         // materials.rs:65-69: Some((None, white))
                    D3 rs = mk(0, 0, 0);
                    if (c == CLS_DIFFUSE || c == CLS_METAL) rs = RT_SAMPLE_ILP ? random_in_unit_sphere_ilp(rng, k0, k1) : random_in_unit_sphere(rng, k0, k1);   // one rejection loop for a warp that straddles both classes
                    if (c == CLS_DIFFUSE) {                                         // materials.rs:84-95, 256-267
                        D3 sd = add(h.normal, rs);
                        if (near_zero(sd)) sd = h.normal;
                        D3 target = add(h.point, sd);
                        nd = sub(target, h.point);
                        if (m.kind == RT_TEXTURE) {
                            double tu, tv;
                            sphere_uv(sub(h.point, center), tu, tv);
                            code = 0x80000000u | texture_texel(p.tex[m.tex], m.param, tu, tv);
                        }
                    } else if (c == CLS_METAL) {                                    // materials.rs:115-129
                        D3 refl = reflect(d, h.normal);
                        nd = add(refl, mul(rs, m.param));
                        if (!(dot(nd, h.normal) > 0.0)) absorbed = true;            // None -> black, no light test (raytracer.rs:127-131)
                    } else if (c == CLS_GLASS) {                                    // materials.rs:176-199
                        double ratio = h.front_face ? __ddiv_rn(1.0, m.param) : m.param;
                        D3 ud = unit_vector(d);
                        double cos_theta = fmin(dot(neg(ud), h.normal), 1.0);
                        double sin_theta = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(cos_theta, cos_theta)));
                        bool refl = __dmul_rn(ratio, sin_theta) > 1.0;              // cannot_refract
                        if (!refl) refl = reflectance(cos_theta, ratio) > rng_f64(rng, k0, k1);   // drawn only if refraction is possible
                        nd = refl ? reflect(ud, h.normal) : refract(ud, h.normal, ratio);
                    }
                    state_dirty = true;
                    if (absorbed) {
                        if (shd == 0u) done = true;            // main path ends black
                        else have_tc = true;                   // the shadow ray returns black
                    } else {
                        // ---- light test, raytracer.rs:89-101 (the uniform is drawn whenever the scene has lights) ----
                        bool pass = false;
                        if (LIGHTS) {
                            const double prob = (c == CLS_GLASS) ? 0.05 : 0.1;
                            const double xi = rng_f64(rng, k0, k1);
                            const unsigned long long depth_now = (unsigned long long)p.max_depth - level;
                            const bool depth_ok = (shd > 0u) ? true : (depth_now > (unsigned long long)p.max_depth - 2ull);   // usize wrap like a release build
                            pass = (xi > __dsub_rn(1.0, __dmul_rn((double)p.n_lights, prob))) && depth_ok;
                            if (pass && shd >= p.max_shadow) { pass = false; atomicAdd(&p.stat[5], 1ull); }   // reported as an error by the host
                        }
                        if (pass) {
                            float ar, ag, ab;
                            albedo_of(is_light ? 0xffffffffu : code, mat, ar, ag, ab);
                            ShadowFrame f;
                            f.px = h.point.x; f.py = h.point.y; f.pz = h.point.z; f.ndx = nd.x; f.ndy = nd.y; f.ndz = nd.z;
                            f.ar = ar; f.ag = ag; f.ab = ab; f.sr = 0.f; f.sg = 0.f; f.sb = 0.f; f.li = 0u; f.code = code; f.is_light = is_light ? 1u : 0u; f.pad = 0u;
                            p.frames[(size_t)shd * p.stack_stride + stack_col + s] = f;
                            ++shd;
                            double4 lq = geo[p.lights[0]];
                            o = h.point; d = sub(mk(lq.x, lq.y, lq.z), h.point);   // Ray::new(point, light.center - point)
                        } else if (shd == 0u) {
                            if (is_light) { cr = 1.f; cg = 1.f; cb = 1.f; done = true; }   // `None => albedo` (raytracer.rs:124)
                            else {
                                stack_push(level, s, code);
                                ++level;
                                o = h.point; d = nd;
                                if (level == p.max_depth) done = true;   // the next ray_color call returns black (raytracer.rs:80-82)
                            }
                        } else {
                            // nested vertex without light contribution: clamp(0 + albedo * black), or white for a Light
                            tr = tg = tb = is_light ? 1.f : 0.f;
                            have_tc = true;
                        }
                    }
                }
                if (LIGHTS) {
                    // return values travel up the shadow-frame stack without tracing (raytracer.rs:103-114)
                    while (have_tc) {
                        ShadowFrame f = p.frames[(size_t)(shd - 1u) * p.stack_stride + stack_col + s];
                        f.sr = __fadd_rn(f.sr, __fmul_rn(f.ar, tr)); f.sg = __fadd_rn(f.sg, __fmul_rn(f.ag, tg)); f.sb = __fadd_rn(f.sb, __fmul_rn(f.ab, tb));
                        ++f.li;
                        state_dirty = true;
                        if (f.li < p.n_lights) {                                   // next light of the same vertex
                            p.frames[(size_t)(shd - 1u) * p.stack_stride + stack_col + s] = f;
                            double4 lq = geo[p.lights[f.li]];
                            o = mk(f.px, f.py, f.pz); d = sub(mk(lq.x, lq.y, lq.z), o);
                            have_tc = false;
                        } else {
                            const float nl = (float)p.n_lights;
                            const float Lr = __fdiv_rn(f.sr, nl), Lg = __fdiv_rn(f.sg, nl), Lb = __fdiv_rn(f.sb, nl);
                            --shd;
                            if (shd == 0u) {                                        // back on the main path
                                have_tc = false;
                                if (f.is_light) { cr = 1.f; cg = 1.f; cb = 1.f; done = true; }
                                else {
                                    p.lterm[(size_t)(level * 3u + 0u) * p.stack_stride + stack_col + s] = Lr;   // level is 0 or 1 here
                                    p.lterm[(size_t)(level * 3u + 1u) * p.stack_stride + stack_col + s] = Lg;
                                    p.lterm[(size_t)(level * 3u + 2u) * p.stack_stride + stack_col + s] = Lb;
                                    stack_push(level, s, f.code);
                                    ++level;
                                    o = mk(f.px, f.py, f.pz); d = mk(f.ndx, f.ndy, f.ndz);
                                    if (level == p.max_depth) done = true;
                                }
                            } else if (f.is_light) { tr = tg = tb = 1.f; }
                            else {                                                  // clamp(light + albedo * ray_color(depth 0) = black)
                                tr = clampf(__fadd_rn(Lr, __fmul_rn(f.ar, 0.0f))); tg = clampf(__fadd_rn(Lg, __fmul_rn(f.ag, 0.0f))); tb = clampf(__fadd_rn(Lb, __fmul_rn(f.ab, 0.0f)));
                            }
                        }
                    }
                }
                if (state_dirty && !done) {
                    s_ox[s] = o.x; s_oy[s] = o.y; s_oz[s] = o.z; s_dx[s] = d.x; s_dy[s] = d.y; s_dz[s] = d.z;
                    s_blk[s] = (rng.blk << 1) | rng.has; s_clo[s] = rng.c_lo; s_chi[s] = rng.c_hi;
                    s_lvl[s] = level;
                    if (LIGHTS) s_shd[s] = shd;
                }
                if (done) {
                    // unwind the recursion: c = clamp(light + albedo * c) per level, innermost first (raytracer.rs:117-122)
                    if (LIGHTS || cr != 0.f || cg != 0.f || cb != 0.f) {
                        for (int l = (int)level - 1; l >= 0; --l) {
                            float ar, ag, ab;
                            albedo_of(stack_get((uint32_t)l, s), mat, ar, ag, ab);
                            float Lr = 0.f, Lg = 0.f, Lb = 0.f;
                            if (LIGHTS && l < 2) {
                                Lr = p.lterm[(size_t)(l * 3 + 0) * p.stack_stride + stack_col + s];
                                Lg = p.lterm[(size_t)(l * 3 + 1) * p.stack_stride + stack_col + s];
                                Lb = p.lterm[(size_t)(l * 3 + 2) * p.stack_stride + stack_col + s];
                            }
                            cr = clampf(__fadd_rn(Lr, __fmul_rn(ar, cr)));
                            cg = clampf(__fadd_rn(Lg, __fmul_rn(ag, cg)));
                            cb = clampf(__fadd_rn(Lb, __fmul_rn(ab, cb)));
                        }
                    }
                    p.samplebuf[s_work[s]] = make_float4(cr, cg, cb, __uint_as_float(rays_sample));
                    s_lvl[s] = kDeadLevel;
                }
            }
            regenerate(active && done, s);
            still_alive = active && (s_lvl[s] != kDeadLevel);
        }
        PF_MARK(pf_shade)
        bool any_alive = __syncthreads_or(still_alive ? 1 : 0);
        PF_MARK(pf_waitC)
#ifdef RT_PROFILE_PHASES
        ++pf_iters;
#endif
        if (!any_alive) break;   // C: pool written back; exit when the CTA has no ray left
    }

#ifdef RT_PROFILE_PHASES
    if (lane == 0) {   // per-warp cycle totals of each phase
        atomicAdd(&p.stat[8], pf_scan); atomicAdd(&p.stat[9], pf_confirm); atomicAdd(&p.stat[10], pf_waitA); atomicAdd(&p.stat[11], pf_sort);
        atomicAdd(&p.stat[12], pf_shade); atomicAdd(&p.stat[13], pf_waitC); atomicAdd(&p.stat[14], pf_iters);
    }
#endif
    // ---- statistics: one atomic per warp ----
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        st_rays += __shfl_down_sync(FULL, st_rays, off);
        st_cand += __shfl_down_sync(FULL, st_cand, off);
        st_ovf += __shfl_down_sync(FULL, st_ovf, off);
        st_samples += __shfl_down_sync(FULL, st_samples, off);
        st_clus += __shfl_down_sync(FULL, st_clus, off);
    }
    if (lane == 0) {
        atomicAdd(&p.stat[4], st_clus);
        atomicAdd(&p.stat[0], st_rays);
        atomicAdd(&p.stat[1], st_cand);
        atomicAdd(&p.stat[2], st_ovf);
        atomicAdd(&p.stat[3], st_samples);
    }
}

// MINB = CTAs per SM the register allocation targets: 4 (64 registers, small spills) when shared memory lets four pools be
// resident, else 2 (up to 128 registers, no spills) - e.g. when 10 k spheres' first-level records take 40 KB per CTA.
template <int MB, bool E, bool LI, bool TW>
static cudaError_t launch_wf(const TraceParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rt_wavefront_kernel<256, MB, E, LI, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rt_wavefront_kernel<256, MB, E, LI, TW><<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

template <int MB>
static cudaError_t launch_wf_minb(const TraceParams& p, int grid, size_t smem, bool exact, cudaStream_t st) {
    const bool li = p.n_lights > 0, tw = p.two_level != 0;
    if (exact) {   // every sphere in f64: the filter levels are not used at all
        return li ? launch_wf<MB, true, true, false>(p, grid, smem, st) : launch_wf<MB, true, false, false>(p, grid, smem, st);
    }
    if (tw) return li ? launch_wf<MB, false, true, true>(p, grid, smem, st) : launch_wf<MB, false, false, true>(p, grid, smem, st);
    return li ? launch_wf<MB, false, true, false>(p, grid, smem, st) : launch_wf<MB, false, false, false>(p, grid, smem, st);
}

cudaError_t launch_wavefront(const TraceParams& p, int grid, size_t smem, int minb, bool exact, cudaStream_t st) {
    return minb >= 4 ? launch_wf_minb<4>(p, grid, smem, exact, st) : launch_wf_minb<2>(p, grid, smem, exact, st);
}

int wavefront_max_ctas_per_sm(size_t smem, int minb) {
    int nb = 0;
    cudaError_t e;
    if (minb >= 4) {
        cudaFuncSetAttribute(rt_wavefront_kernel<256, 4, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rt_wavefront_kernel<256, 4, false, true, true>, 256, smem);
    } else {
        cudaFuncSetAttribute(rt_wavefront_kernel<256, 2, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rt_wavefront_kernel<256, 2, false, true, true>, 256, smem);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
}

}  // namespace rtk
