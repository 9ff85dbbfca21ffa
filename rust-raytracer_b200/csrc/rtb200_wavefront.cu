// rtb200_wavefront.cu — the production trace kernel: a persistent-threads WAVEFRONT tracer at CTA scope.
//
// One persistent CTA (256 threads) per resident slot of every SM owns a pool of 256 ray slots in shared memory, laid
// out SoA (the scene's hierarchy / sphere / material records are staged next to it by 1-D TMA bulk copies when they
// fit). Until the global (pixel,sample) queue is drained and the pool is empty, the CTA repeats three stages:
//
//   closest-hit   hit_world (raytracer.rs:44-59) as a WARP-COOPERATIVE traversal of an 8-wide BVH. Each warp owns the
//                 32 rays of its slots and three work lists in shared memory: (ray, node) pairs, (ray, leaf) pairs and
//                 (ray, sphere) candidates. A step pops up to 32 pairs, ONE PER LANE, whatever ray they belong to:
//                   node step   8 conservative f32 slab tests (FFMA2 + FMNMX3), children pushed by a warp prefix sum;
//                   leaf step   the 7-FMA conservative sphere test on the leaf's 4 spheres (two per packed FFMA2);
//                   exact step  the reference-exact f64 Sphere::hit (sphere.rs:46-78) on a candidate; the per-ray result
//                               is the lexicographic minimum of (root, ORIGINAL sphere index) - equal to hit_world's
//                               fold with its strict '<' (first sphere wins ties) whatever the visiting order.
//                 So 32 lanes do 32 tests at every level, however unevenly the work is spread over the rays.
//   sort          rays are classified {miss, diffuse, metal, glass, light} and compacted class by class with warp
//                 ballots + one shared-memory atomic per (warp, class): perm[] lists the live slots sorted by class.
//   shade+ray-gen thread i <-> slot perm[i], so a warp shades ONE material: Material::scatter (materials.rs:44-54) or
//                 the sky (raytracer.rs:134-163), iteratively (no recursion): albedo codes go to a per-slot stack that
//                 is unwound backwards on termination so the f32 products associate exactly like the reference's
//                 recursion (raytracer.rs:117-122). A terminated path writes its sample and the same thread
//                 immediately regenerates the slot from the queue (render_line's jitter + Camera::get_ray,
//                 raytracer.rs:199-201): one warp-aggregated atomic pops the work items.
#include <cstdio>

#include "rtb200_kernels.cuh"

using namespace rtd;

#ifndef RT_SAMPLE_ILP
#define RT_SAMPLE_ILP 1   // two rejection trials per trip with their Philox blocks computed together (bit-identical stream)
#endif
#ifndef RT_SMEM_STACK
#define RT_SMEM_STACK 3   // albedo-stack levels kept in shared memory per slot (deeper levels live in global memory)
#endif

namespace rtk {

namespace {

enum : uint32_t { CLS_MISS = 0, CLS_DIFFUSE = 1, CLS_METAL = 2, CLS_GLASS = 3, CLS_LIGHT = 4, CLS_DEAD = 5, N_CLS = 6 };
constexpr uint32_t kDeadLevel = 0xffffffffu;
constexpr uint32_t kLeafBit = 0x80000000u;
constexpr unsigned long long kNoHitBits = 0x7ff0000000000000ull;   // +inf as the "no root yet" key (roots are > t_min > 0)

struct WfSmem {
    uint32_t nodes_off, leafrec_off, leafid_off, filt_off, geo_off, mat_off;
    uint32_t cA, cB, cC;                            // float4[kBlock] each: per-ray f32 constants of the conservative tests
    uint32_t lists;                                 // uint32[warps][kCapIn + kCapLf + kCapCd]
    uint32_t ox, oy, oz, dx, dy, dz, bt;            // double[kBlock] each (bt: best root, updated as u64 bits)
    uint32_t bi, work, pix, smp, blk, clo, chi, lvl, shd;   // uint32[kBlock] each
    uint32_t perm;                                  // uint16[kBlock]
    uint32_t stk;                                   // uint32[RT_SMEM_STACK][kBlock]: first levels of the albedo stack
    uint32_t cnt;                                   // uint32[2][8]
    uint32_t total;
};

// mask: bit0 hierarchy (MODE_TREE) / flat records (MODE_BRUTE), bit1 exact geometry, bit2 materials in shared memory
__host__ __device__ inline WfSmem wf_layout(uint32_t n, uint32_t n_pairs, uint32_t n_nodes, uint32_t n_leaves, uint32_t mode, uint32_t mask) {
    WfSmem L;
    uint32_t off = 16;   // mbarrier
    L.nodes_off = off;   if (mode == MODE_TREE && (mask & 1u)) off += n_nodes * (uint32_t)(kNodeVec * 16);
    L.leafrec_off = off; if (mode == MODE_TREE && (mask & 1u)) off += n_leaves * (uint32_t)(kLeafK * 16);
    L.leafid_off = off;  if (mode == MODE_TREE && (mask & 1u)) off += n_leaves * (uint32_t)(kLeafK * 4);
    L.filt_off = off;    if (mode == MODE_BRUTE && (mask & 1u)) off += n_pairs * 32u;
    L.geo_off = off; if (mask & 2u) off += n * 32u;
    L.mat_off = off; if (mask & 4u) off += n * 32u;
    off = (off + 15u) & ~15u;
    L.cA = off; L.cB = off; L.cC = off; L.lists = off;
    if (mode == MODE_TREE) {
        L.cA = off; off += kBlock * 16u; L.cB = off; off += kBlock * 16u; L.cC = off; off += kBlock * 16u;
        L.lists = off; off += (uint32_t)(kBlock / 32) * (uint32_t)(kCapIn + kCapLf + kCapCd) * 4u;
    }
    L.ox = off; off += kBlock * 8u; L.oy = off; off += kBlock * 8u; L.oz = off; off += kBlock * 8u;
    L.dx = off; off += kBlock * 8u; L.dy = off; off += kBlock * 8u; L.dz = off; off += kBlock * 8u;
    L.bt = off; off += kBlock * 8u;
    L.bi = off; off += kBlock * 4u; L.work = off; off += kBlock * 4u; L.pix = off; off += kBlock * 4u; L.smp = off; off += kBlock * 4u;
    L.blk = off; off += kBlock * 4u; L.clo = off; off += kBlock * 4u; L.chi = off; off += kBlock * 4u; L.lvl = off; off += kBlock * 4u; L.shd = off; off += kBlock * 4u;
    L.perm = off; off += kBlock * 2u;
    L.stk = off; off += (uint32_t)RT_SMEM_STACK * kBlock * 4u;
    L.cnt = off; off += 2u * 8u * 4u;
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

RT_DEV float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
RT_DEV float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// inclusive warp prefix sum of a packed pair of 16-bit counters
RT_DEV uint32_t warp_scan_incl(uint32_t v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += t;
    }
    return v;
}

// the 7-FMA conservative sphere test on NP pair-packed records: D = (c.d^ - o.d^)^2 + 2 c.o + nk   (candidate iff D >= thr)
#define RT_FILTER_PAIRS(REC, DV, NP)                                                                                                  \
    _Pragma("unroll") for (int q = 0; q < (NP); ++q) {                                                                              \
        float4 A_ = (REC)[2 * q], B_ = (REC)[2 * q + 1];                                                                             \
        float2 cx = make_float2(A_.x, A_.y), cy = make_float2(A_.z, A_.w), cz = make_float2(B_.x, B_.y), nk = make_float2(B_.z, B_.w); \
        float2 bb = __ffma2_rn(cz, dz2, nod2);                                                                                       \
        float2 tt = __ffma2_rn(cz, oz2, nk);                                                                                         \
        bb = __ffma2_rn(cy, dy2, bb);                                                                                                \
        tt = __ffma2_rn(cy, oy2, tt);                                                                                                \
        bb = __ffma2_rn(cx, dx2, bb);                                                                                                \
        tt = __ffma2_rn(cx, ox2, tt);                                                                                                \
        (DV)[q] = __ffma2_rn(bb, bb, tt);                                                                                            \
    }

}  // namespace

size_t wavefront_smem_bytes(const TraceParams& p, uint32_t mode, uint32_t smem_mask) {
    return wf_layout(p.n, p.n_pairs, p.n_nodes, p.n_leaves, mode, smem_mask).total;
}

template <int MINB, uint32_t MODE, bool LIGHTS>
__global__ void __launch_bounds__(kBlock, MINB) rt_wavefront_kernel(const __grid_constant__ TraceParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WfSmem L = wf_layout(p.n, p.n_pairs, p.n_nodes, p.n_leaves, MODE, p.scene_in_smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    const bool tree_smem = (p.scene_in_smem & 1u) != 0u;
    const float4* nodes = (MODE == MODE_TREE && tree_smem) ? reinterpret_cast<const float4*>(smem_raw + L.nodes_off) : p.nodes;
    const float4* leaf_rec = (MODE == MODE_TREE && tree_smem) ? reinterpret_cast<const float4*>(smem_raw + L.leafrec_off) : p.leaf_rec;
    const uint32_t* leaf_id = (MODE == MODE_TREE && tree_smem) ? reinterpret_cast<const uint32_t*>(smem_raw + L.leafid_off) : p.leaf_id;
    const float4* s_filt = (MODE == MODE_BRUTE && tree_smem) ? reinterpret_cast<const float4*>(smem_raw + L.filt_off) : p.filt;
    const double4* geo = (p.scene_in_smem & 2u) ? reinterpret_cast<const double4*>(smem_raw + L.geo_off) : p.geo;
    const DevMat* mat = (p.scene_in_smem & 4u) ? reinterpret_cast<const DevMat*>(smem_raw + L.mat_off) : p.mat;
    float4* s_cA = reinterpret_cast<float4*>(smem_raw + L.cA);   // {o.x, o.y, o.z, m_ray}   recentred f32 origin, slab margin
    float4* s_cB = reinterpret_cast<float4*>(smem_raw + L.cB);   // {1/d^.x, 1/d^.y, 1/d^.z, thr}
    float4* s_cC = reinterpret_cast<float4*>(smem_raw + L.cC);   // {d^.x, d^.y, d^.z, -o.d^}
    uint32_t* s_lists = reinterpret_cast<uint32_t*>(smem_raw + L.lists);
    double* s_ox = reinterpret_cast<double*>(smem_raw + L.ox); double* s_oy = reinterpret_cast<double*>(smem_raw + L.oy);
    double* s_oz = reinterpret_cast<double*>(smem_raw + L.oz); double* s_dx = reinterpret_cast<double*>(smem_raw + L.dx);
    double* s_dy = reinterpret_cast<double*>(smem_raw + L.dy); double* s_dz = reinterpret_cast<double*>(smem_raw + L.dz);
    double* s_bt = reinterpret_cast<double*>(smem_raw + L.bt);
    unsigned long long* s_btu = reinterpret_cast<unsigned long long*>(smem_raw + L.bt);
    uint32_t* s_bi = reinterpret_cast<uint32_t*>(smem_raw + L.bi); uint32_t* s_work = reinterpret_cast<uint32_t*>(smem_raw + L.work);
    uint32_t* s_pix = reinterpret_cast<uint32_t*>(smem_raw + L.pix); uint32_t* s_smp = reinterpret_cast<uint32_t*>(smem_raw + L.smp);
    uint32_t* s_blk = reinterpret_cast<uint32_t*>(smem_raw + L.blk); uint32_t* s_clo = reinterpret_cast<uint32_t*>(smem_raw + L.clo);
    uint32_t* s_chi = reinterpret_cast<uint32_t*>(smem_raw + L.chi); uint32_t* s_lvl = reinterpret_cast<uint32_t*>(smem_raw + L.lvl);
    uint32_t* s_shd = reinterpret_cast<uint32_t*>(smem_raw + L.shd);   // depth of the shadow-frame stack (0 = main path)
    uint16_t* s_perm = reinterpret_cast<uint16_t*>(smem_raw + L.perm);
    uint32_t* s_stk = reinterpret_cast<uint32_t*>(smem_raw + L.stk);
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw + L.cnt);

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
    s_lvl[tid] = kDeadLevel;
    __syncthreads();
    if (tid == 0) {
        const uint32_t b_nodes = (MODE == MODE_TREE && tree_smem) ? p.n_nodes * (uint32_t)(kNodeVec * 16) : 0u;
        const uint32_t b_lrec = (MODE == MODE_TREE && tree_smem) ? p.n_leaves * (uint32_t)(kLeafK * 16) : 0u;
        const uint32_t b_lid = (MODE == MODE_TREE && tree_smem) ? p.n_leaves * (uint32_t)(kLeafK * 4) : 0u;
        const uint32_t b_filt = (MODE == MODE_BRUTE && tree_smem) ? p.n_pairs * 32u : 0u;
        const uint32_t b_geo = (p.scene_in_smem & 2u) ? p.n * 32u : 0u, b_mat = (p.scene_in_smem & 4u) ? p.n * 32u : 0u;
        mbar_arrive_expect_tx(bar, b_nodes + b_lrec + b_lid + b_filt + b_geo + b_mat);
        if (b_nodes) bulk_stage(smem_raw + L.nodes_off, p.nodes, b_nodes, bar);
        if (b_lrec) bulk_stage(smem_raw + L.leafrec_off, p.leaf_rec, b_lrec, bar);
        if (b_lid) bulk_stage(smem_raw + L.leafid_off, p.leaf_id, b_lid, bar);
        if (b_filt) bulk_stage(smem_raw + L.filt_off, p.filt, b_filt, bar);
        if (b_geo) bulk_stage(smem_raw + L.geo_off, p.geo, b_geo, bar);
        if (b_mat) bulk_stage(smem_raw + L.mat_off, p.mat, b_mat, bar);
    }
    mbar_wait(bar, 0);

    unsigned long long st_rays = 0, st_cand = 0, st_ovf = 0, st_samples = 0, st_leaves = 0, st_nodes = 0;
#ifdef RT_PROFILE_PHASES
    unsigned long long pf_trav = 0, pf_exact = 0, pf_waitA = 0, pf_sort = 0, pf_shade = 0, pf_waitC = 0, pf_iters = 0, pf_t = clock64();
#define PF_MARK(acc) { unsigned long long now_ = clock64(); acc += now_ - pf_t; pf_t = now_; }
#else
#define PF_MARK(acc)
#endif

    // Regenerate slot `s` from the global (pixel,sample) queue. Warp-synchronous: every lane of the warp calls it,
    // `want` says whether this lane's slot needs a new path. raytracer.rs:199-201 + camera.rs:79-84.
    bool exhausted = false;   // warp-uniform: this warp has seen the end of the queue
    auto regenerate = [&](bool want, uint32_t s) {
        want = want && !exhausted;
        unsigned need = __ballot_sync(FULL, want);
        if (!need) return;
        int leader = __ffs(need) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(p.work_counter, (unsigned)__popc(need));
        base = __shfl_sync(FULL, base, leader);
        if (base + (unsigned)__popc(need) >= p.total_work) exhausted = true;
        if (!want) return;
        unsigned my = base + __popc(need & ((1u << lane) - 1u));
        if (my >= p.total_work) return;
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

    const uint32_t wbase = (uint32_t)(tid & ~31);                      // first slot of this warp
    uint32_t* wl_in = s_lists + (uint32_t)(tid >> 5) * (uint32_t)(kCapIn + kCapLf + kCapCd);
    uint32_t* wl_lf = wl_in + kCapIn;
    uint32_t* wl_cd = wl_lf + kCapLf;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // LIFO reserve: single-entry descents grow the node stack by at most 7 per level, so multi-entry steps may fill it only
    // up to fat_in; above that the stack is popped one entry at a time and can never overflow.
    const uint32_t fat_in = (uint32_t)kCapIn - 7u * p.depth - 8u;

    uint32_t it = 0;
    for (;; ++it) {
        uint32_t* cnt = s_cnt + (it & 1u) * 8u;
        // =========================== closest-hit ===========================
        const bool alive = s_lvl[tid] != kDeadLevel;
        uint32_t cls = CLS_DEAD;
        if (__ballot_sync(FULL, alive) != 0u) {   // a warp whose 32 slots are all empty skips the stage (frame tail)
            const D3 o = mk(s_ox[tid], s_oy[tid], s_oz[tid]), d = mk(s_dx[tid], s_dy[tid], s_dz[tid]);
            const double a = length_squared(d);
            bool ovf = false;
            // thread-private exact f64 confirmation (fallback paths: every sphere / the always-list / MODE_BRUTE candidates).
            // hit_world (raytracer.rs:44-59) keeps the closest root and, on equal t, the first sphere in list order; because
            // Sphere::hit(t_max) accepts exactly r < t_max with r the first root beyond t_min, that fold equals the lexicographic
            // minimum of (r, index) over all spheres - so candidates may be confirmed in any order.
            double best_t = DBL_MAX;
            int best = -1;
            auto confirm = [&](int j) {
                double4 gq = geo[j];
                double root;
                if (sphere_root(mk(gq.x, gq.y, gq.z), gq.w, o, d, a, 0.001, DBL_MAX, root)) {
                    if (best < 0 || root < best_t || (root == best_t && j < best)) { best_t = root; best = j; }
                }
                ++st_cand;
            };
            s_btu[tid] = kNoHitBits;
            s_bi[tid] = 0xffffffffu;
            if (MODE != MODE_EXACT) {
                // per-ray constants in the recentred f32 frame (DESIGN.md "filter soundness")
                const float ofx = __double2float_rn(__dsub_rn(o.x, p.gx)), ofy = __double2float_rn(__dsub_rn(o.y, p.gy)),
                            ofz = __double2float_rn(__dsub_rn(o.z, p.gz));
                const float dfx = __double2float_rn(d.x), dfy = __double2float_rn(d.y), dfz = __double2float_rn(d.z);
                const float s = fmaf(dfx, dfx, fmaf(dfy, dfy, dfz * dfz));
                const float oo = fmaf(ofx, ofx, fmaf(ofy, ofy, ofz * ofz));
                const bool ok = (s > 1e-30f) && (s < 1e30f) && (oo < 1e30f);
                const float inv = rsqrtf(s);
                const float dnx = dfx * inv, dny = dfy * inv, dnz = dfz * inv;
                const float nod = -fmaf(ofx, dnx, fmaf(ofy, dny, ofz * dnz));
                const float thr = __fmul_rd(oo, p.er_coef);
                if (alive && !ok) ovf = true;
                if (MODE == MODE_TREE) {
                    // slab constants: 1/d^ with |d^| clamped away from zero (keeps every product finite), margin 32u|o| rounded up
                    const float ax = fabsf(dnx) < 1e-20f ? copysignf(1e-20f, dnx) : dnx;
                    const float ay = fabsf(dny) < 1e-20f ? copysignf(1e-20f, dny) : dny;
                    const float az = fabsf(dnz) < 1e-20f ? copysignf(1e-20f, dnz) : dnz;
                    const float mray = __fmul_ru(1.9073486328125e-6f, __fsqrt_ru(oo));
                    s_cA[tid] = make_float4(ofx, ofy, ofz, mray);
                    s_cB[tid] = make_float4(__frcp_rn(ax), __frcp_rn(ay), __frcp_rn(az), thr);
                    s_cC[tid] = make_float4(dnx, dny, dnz, nod);
                    // ---- warp-cooperative traversal ----
                    const bool enter = alive && ok && p.n_nodes != 0u;
                    const unsigned em = __ballot_sync(FULL, enter);
                    uint32_t n_in = (uint32_t)__popc(em), n_lf = 0u, n_cd = 0u;
                    if (enter) wl_in[__popc(em & lt_mask)] = (uint32_t)lane;   // (root node 0) << 5 | ray
                    __syncwarp();
                    uint32_t guard = 0;
                    for (;;) {
                        if (n_in != 0u && n_lf <= (uint32_t)(kCapLf - 64)) {
                            // ---------------- node step: lane <-> one (ray, node) pair from the top of the stack ----------------
                            const uint32_t m = n_in < 32u ? n_in : 32u;
                            const bool act = (uint32_t)lane < m;
                            const uint32_t e = act ? wl_in[n_in - 1u - (uint32_t)lane] : 0u;
                            const uint32_t ray = e & 31u, node = e >> 5;
                            uint32_t hit = 0u, leafbits = 0u;
                            const float4* N = nodes + (size_t)node * kNodeVec;
                            if (act) {
                                const float4 A = s_cA[wbase + ray], B = s_cB[wbase + ray];
                                // near/far plane of each axis by the sign of d^; planes shifted outwards by the per-ray margin
                                const uint32_t sx = __float_as_uint(B.x) >> 31, sy = __float_as_uint(B.y) >> 31, sz = __float_as_uint(B.z) >> 31;
                                const float mx = copysignf(A.w, B.x), my = copysignf(A.w, B.y), mz = copysignf(A.w, B.z);
                                const float cnx = __fmul_rn(__fadd_rn(A.x, mx), -B.x), cfx = __fmul_rn(__fsub_rn(A.x, mx), -B.x);
                                const float cny = __fmul_rn(__fadd_rn(A.y, my), -B.y), cfy = __fmul_rn(__fsub_rn(A.y, my), -B.y);
                                const float cnz = __fmul_rn(__fadd_rn(A.z, mz), -B.z), cfz = __fmul_rn(__fsub_rn(A.z, mz), -B.z);
                                const float2 ix2 = make_float2(B.x, B.x), iy2 = make_float2(B.y, B.y), iz2 = make_float2(B.z, B.z);
                                const float2 cnx2 = make_float2(cnx, cnx), cny2 = make_float2(cny, cny), cnz2 = make_float2(cnz, cnz);
                                const float2 cfx2 = make_float2(cfx, cfx), cfy2 = make_float2(cfy, cfy), cfz2 = make_float2(cfz, cfz);
                                const float4* Nnx = N + (sx ? 6 : 0); const float4* Nfx = N + (sx ? 0 : 6);
                                const float4* Nny = N + (sy ? 8 : 2); const float4* Nfy = N + (sy ? 2 : 8);
                                const float4* Nnz = N + (sz ? 10 : 4); const float4* Nfz = N + (sz ? 4 : 10);
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const float4 NX = Nnx[h], NY = Nny[h], NZ = Nnz[h], FX = Nfx[h], FY = Nfy[h], FZ = Nfz[h];
                                    const float2 tnx0 = __ffma2_rn(make_float2(NX.x, NX.y), ix2, cnx2), tnx1 = __ffma2_rn(make_float2(NX.z, NX.w), ix2, cnx2);
                                    const float2 tny0 = __ffma2_rn(make_float2(NY.x, NY.y), iy2, cny2), tny1 = __ffma2_rn(make_float2(NY.z, NY.w), iy2, cny2);
                                    const float2 tnz0 = __ffma2_rn(make_float2(NZ.x, NZ.y), iz2, cnz2), tnz1 = __ffma2_rn(make_float2(NZ.z, NZ.w), iz2, cnz2);
                                    const float2 tfx0 = __ffma2_rn(make_float2(FX.x, FX.y), ix2, cfx2), tfx1 = __ffma2_rn(make_float2(FX.z, FX.w), ix2, cfx2);
                                    const float2 tfy0 = __ffma2_rn(make_float2(FY.x, FY.y), iy2, cfy2), tfy1 = __ffma2_rn(make_float2(FY.z, FY.w), iy2, cfy2);
                                    const float2 tfz0 = __ffma2_rn(make_float2(FZ.x, FZ.y), iz2, cfz2), tfz1 = __ffma2_rn(make_float2(FZ.z, FZ.w), iz2, cfz2);
                                    // hit iff max(t_near, 0) <= t_far
                                    hit |= (fmaxf(fmax3(tnx0.x, tny0.x, tnz0.x), 0.f) <= fmin3(tfx0.x, tfy0.x, tfz0.x) ? 1u : 0u) << (4 * h + 0);
                                    hit |= (fmaxf(fmax3(tnx0.y, tny0.y, tnz0.y), 0.f) <= fmin3(tfx0.y, tfy0.y, tfz0.y) ? 1u : 0u) << (4 * h + 1);
                                    hit |= (fmaxf(fmax3(tnx1.x, tny1.x, tnz1.x), 0.f) <= fmin3(tfx1.x, tfy1.x, tfz1.x) ? 1u : 0u) << (4 * h + 2);
                                    hit |= (fmaxf(fmax3(tnx1.y, tny1.y, tnz1.y), 0.f) <= fmin3(tfx1.y, tfy1.y, tfz1.y) ? 1u : 0u) << (4 * h + 3);
                                }
                                const uint4 R0 = *reinterpret_cast<const uint4*>(N + 12), R1 = *reinterpret_cast<const uint4*>(N + 13);
                                leafbits = (R0.x >> 31) | ((R0.y >> 31) << 1) | ((R0.z >> 31) << 2) | ((R0.w >> 31) << 3) |
                                           ((R1.x >> 31) << 4) | ((R1.y >> 31) << 5) | ((R1.z >> 31) << 6) | ((R1.w >> 31) << 7);
                            }
                            const uint32_t packed = (uint32_t)__popc(hit & ~leafbits) | ((uint32_t)__popc(hit & leafbits) << 16);
                            const uint32_t inc = warp_scan_incl(packed, lane);
                            // commit the longest prefix of lanes (top of the stack first) whose pushes fit
                            const uint32_t new_in = n_in - ((uint32_t)lane + 1u) + (inc & 0xffffu);
                            const bool fits = new_in <= (lane == 0 ? (uint32_t)kCapIn : fat_in) && n_lf + (inc >> 16) <= (uint32_t)kCapLf;
                            const unsigned okm = __ballot_sync(FULL, fits || !act);
                            uint32_t k = okm == FULL ? 32u : (uint32_t)(__ffs(~okm) - 1);
                            k = k < m ? k : m;
                            if (k == 0u) { if (lane == 0) atomicAdd(&p.err[1], 1ull); break; }   // cannot happen (reserve argument); never spin
                            const uint32_t tot = __shfl_sync(FULL, inc, (int)k - 1);
                            __syncwarp();   // every lane has read its entry before the stack is overwritten
                            if (act && (uint32_t)lane < k) {
                                const uint32_t exc = inc - packed;
                                uint32_t pi = (n_in - k) + (exc & 0xffffu), pl = n_lf + (exc >> 16);
                                const uint32_t* refs = reinterpret_cast<const uint32_t*>(N + 12);
                                uint32_t mm = hit;
                                while (mm) {
                                    const int c = __ffs(mm) - 1;
                                    mm &= mm - 1u;
                                    const uint32_t ref = refs[c];
                                    if (ref & kLeafBit) wl_lf[pl++] = (ref << 5) | ray;   // the shift drops the leaf bit
                                    else wl_in[pi++] = (ref << 5) | ray;
                                }
                                ++st_nodes;
                            }
                            n_in = n_in - k + (tot & 0xffffu);
                            n_lf += tot >> 16;
                            __syncwarp();
                        } else if (n_lf != 0u && n_cd <= (uint32_t)(kCapCd - 64)) {
                            // ---------------- leaf step: lane <-> one (ray, leaf) pair: conservative sphere test on its spheres ----------------
                            const uint32_t m = n_lf < 32u ? n_lf : 32u;
                            const bool act = (uint32_t)lane < m;
                            const uint32_t e = act ? wl_lf[n_lf - 1u - (uint32_t)lane] : 0u;
                            const uint32_t ray = e & 31u, leaf = e >> 5;
                            uint32_t hit = 0u;
                            if (act) {
                                const float4 A = s_cA[wbase + ray], C = s_cC[wbase + ray];
                                const float th = s_cB[wbase + ray].w;
                                const float2 dx2 = make_float2(C.x, C.x), dy2 = make_float2(C.y, C.y), dz2 = make_float2(C.z, C.z);
                                const float2 ox2 = make_float2(2.f * A.x, 2.f * A.x), oy2 = make_float2(2.f * A.y, 2.f * A.y), oz2 = make_float2(2.f * A.z, 2.f * A.z);
                                const float2 nod2 = make_float2(C.w, C.w);
                                float2 Dv[kLeafK / 2];
                                const float4* rec = leaf_rec + (size_t)leaf * kLeafK;
                                RT_FILTER_PAIRS(rec, Dv, kLeafK / 2)
#pragma unroll
                                for (int q = 0; q < kLeafK / 2; ++q) hit |= (Dv[q].x >= th ? 1u : 0u) << (2 * q) | (Dv[q].y >= th ? 1u : 0u) << (2 * q + 1);
                            }
                            const uint32_t cntc = (uint32_t)__popc(hit);
                            const uint32_t inc = warp_scan_incl(cntc, lane);
                            const bool fits = n_cd + inc <= (uint32_t)kCapCd;
                            const unsigned okm = __ballot_sync(FULL, fits || !act);
                            uint32_t k = okm == FULL ? 32u : (uint32_t)(__ffs(~okm) - 1);
                            k = k < m ? k : m;
                            if (k == 0u) { if (lane == 0) atomicAdd(&p.err[1], 1ull); break; }
                            const uint32_t tot = __shfl_sync(FULL, inc, (int)k - 1);
                            if (act && (uint32_t)lane < k) {
                                uint32_t pc = n_cd + inc - cntc;
                                const uint32_t* ids = leaf_id + (size_t)leaf * kLeafK;
                                uint32_t mm = hit;
                                while (mm) {
                                    const int c = __ffs(mm) - 1;
                                    mm &= mm - 1u;
                                    wl_cd[pc++] = (ids[c] << 5) | ray;
                                }
                                ++st_leaves;
                            }
                            n_lf -= k;
                            n_cd += tot;
                            __syncwarp();
                        } else if (n_cd != 0u) {
                            // ---------------- exact step: lane <-> one (ray, sphere) candidate, reference-exact f64 Sphere::hit ----------------
                            const uint32_t m = n_cd < 32u ? n_cd : 32u;
                            const bool act = (uint32_t)lane < m;
                            const uint32_t e = act ? wl_cd[n_cd - 1u - (uint32_t)lane] : 0u;
                            const uint32_t slot = wbase + (e & 31u), sph = e >> 5;
                            unsigned long long key = ~0ull;
                            if (act) {
                                const D3 ro = mk(s_ox[slot], s_oy[slot], s_oz[slot]), rd = mk(s_dx[slot], s_dy[slot], s_dz[slot]);
                                const double4 gq = geo[sph];
                                double root;
                                if (sphere_root(mk(gq.x, gq.y, gq.z), gq.w, ro, rd, length_squared(rd), 0.001, DBL_MAX, root)) key = (unsigned long long)__double_as_longlong(root);
                                ++st_cand;
                            }
                            // per-ray lexicographic minimum of (root, sphere index): roots are positive, so their bit patterns order like the values
                            const bool h = key != ~0ull;
                            const unsigned long long before = h ? s_btu[slot] : 0ull;
                            __syncwarp();
                            if (h && key < before) atomicMin(&s_btu[slot], key);
                            __syncwarp();
                            const bool mine = h && key == s_btu[slot];
                            if (mine && key < before) atomicMax(&s_bi[slot], 0xffffffffu);   // the root got smaller in this step: forget the old index
                            __syncwarp();
                            if (mine) atomicMin(&s_bi[slot], sph);
                            n_cd -= m;
                            __syncwarp();
                        } else {
                            break;
                        }
                        if (++guard > (1u << 22)) { if (lane == 0) atomicAdd(&p.err[1], 1ull); break; }
                    }
                    PF_MARK(pf_trav)
                } else {   // MODE_BRUTE: hit_world's linear scan with the conservative sphere test in front of the exact one
                    if (alive && ok) {
                        const float2 dx2 = make_float2(dnx, dnx), dy2 = make_float2(dny, dny), dz2 = make_float2(dnz, dnz);
                        const float2 ox2 = make_float2(2.f * ofx, 2.f * ofx), oy2 = make_float2(2.f * ofy, 2.f * ofy), oz2 = make_float2(2.f * ofz, 2.f * ofz);
                        const float2 nod2 = make_float2(nod, nod);
#pragma unroll 1
                        for (uint32_t pp = 0; pp < p.n_pairs; pp += 4) {
                            float2 Dv[4];
                            const float4* rec = s_filt + 2 * pp;
                            RT_FILTER_PAIRS(rec, Dv, 4)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t j = 2u * (pp + (uint32_t)q);
                                if (Dv[q].x >= thr && j < p.n) confirm((int)j);
                                if (Dv[q].y >= thr && j + 1u < p.n) confirm((int)j + 1);
                            }
                        }
                    }
                }
            } else {
                ovf = alive;
            }
            if (alive) {
                if (ovf) {   // MODE_EXACT, or a ray outside the f32 frame's safe range: every sphere in f64
                    ++st_ovf;
                    for (int k = 0; k < (int)p.n; ++k) confirm(k);
                } else if (MODE == MODE_TREE) {
                    for (uint32_t k = 0; k < p.n_always; ++k) confirm((int)p.always[k]);
                }
                // merge the thread-private result with the traversal's (this slot is only touched by its own thread now)
                const unsigned long long tb = s_btu[tid];
                const uint32_t ti = s_bi[tid];
                if (best >= 0) {
                    const unsigned long long kb = (unsigned long long)__double_as_longlong(best_t);
                    if (ti == 0xffffffffu || kb < tb || (kb == tb && (uint32_t)best < ti)) { s_bt[tid] = best_t; s_bi[tid] = (uint32_t)best; }
                }
                const uint32_t fin = s_bi[tid];
                cls = CLS_MISS;
                if (fin != 0xffffffffu) {
                    uint32_t kind = mat[fin].kind;
                    cls = (kind == RT_METAL) ? CLS_METAL : (kind == RT_GLASS) ? CLS_GLASS : (kind == RT_LIGHT) ? CLS_LIGHT : CLS_DIFFUSE;
                }
                ++st_rays;
            }
        }
        PF_MARK(pf_exact)

        // =========================== sort: compact the live slots class by class ===========================
        uint32_t wbase_c = 0, rank = 0;
#pragma unroll
        for (uint32_t c = 0; c < CLS_DEAD; ++c) {
            unsigned b = __ballot_sync(FULL, cls == c);
            if (b) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt[c], (uint32_t)__popc(b));
                base = __shfl_sync(FULL, base, 0);
                if (cls == c) { wbase_c = base; rank = __popc(b & lt_mask); }
            }
        }
        PF_MARK(pf_sort)
        __syncthreads();   // A: class counts complete (and every warp's closest-hit results are in the pool)
        PF_MARK(pf_waitA)
        uint32_t c0 = cnt[0], c1 = cnt[1], c2 = cnt[2], c3 = cnt[3], c4 = cnt[4];
        const uint32_t e0 = c0, e1 = e0 + c1, e2 = e1 + c2, e3 = e2 + c3, n_live = e3 + c4;   // class end offsets
        if (cls != CLS_DEAD) {
            uint32_t start = cls == 0 ? 0u : cls == 1 ? e0 : cls == 2 ? e1 : cls == 3 ? e2 : e3;
            s_perm[start + wbase_c + rank] = (uint16_t)tid;
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
                            if (pass && shd >= p.max_shadow) { pass = false; atomicAdd(&p.err[0], 1ull); }   // reported as an error by the host
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
        atomicAdd(&p.stat[8], pf_trav); atomicAdd(&p.stat[9], pf_exact); atomicAdd(&p.stat[10], pf_waitA); atomicAdd(&p.stat[11], pf_sort);
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
        st_leaves += __shfl_down_sync(FULL, st_leaves, off);
        st_nodes += __shfl_down_sync(FULL, st_nodes, off);
    }
    if (lane == 0) {
        atomicAdd(&p.stat[0], st_rays);
        atomicAdd(&p.stat[1], st_cand);
        atomicAdd(&p.stat[2], st_ovf);
        atomicAdd(&p.stat[3], st_samples);
        atomicAdd(&p.stat[4], st_leaves);
        atomicAdd(&p.stat[6], st_nodes);
    }
}

// MINB = CTAs per SM the register allocation targets (3: up to 80 registers; 2: up to 128 when shared memory only lets
// two pools be resident).
template <int MB, uint32_t MODE, bool LI>
static cudaError_t launch_wf(const TraceParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(rt_wavefront_kernel<MB, MODE, LI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rt_wavefront_kernel<MB, MODE, LI><<<grid, kBlock, smem, st>>>(p);
    return cudaGetLastError();
}

template <typename F>
static auto dispatch(uint32_t mode, bool lights, int minb, F&& f) {
    // the validation modes exist in one register budget only
    if (mode == MODE_EXACT) return lights ? f(rt_wavefront_kernel<2, MODE_EXACT, true>) : f(rt_wavefront_kernel<2, MODE_EXACT, false>);
    if (mode == MODE_BRUTE) return lights ? f(rt_wavefront_kernel<2, MODE_BRUTE, true>) : f(rt_wavefront_kernel<2, MODE_BRUTE, false>);
    if (minb >= 3) return lights ? f(rt_wavefront_kernel<3, MODE_TREE, true>) : f(rt_wavefront_kernel<3, MODE_TREE, false>);
    return lights ? f(rt_wavefront_kernel<2, MODE_TREE, true>) : f(rt_wavefront_kernel<2, MODE_TREE, false>);
}

cudaError_t launch_wavefront(const TraceParams& p, uint32_t mode, int grid, size_t smem, int minb, cudaStream_t st) {
    return dispatch(mode, p.n_lights > 0, minb, [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, kBlock, smem, st>>>(p);
        return cudaGetLastError();
    });
}

int wavefront_max_ctas_per_sm(uint32_t mode, bool lights, size_t smem, int minb) {
    return dispatch(mode, lights, minb, [&](auto kern) -> int {
        int nb = 0;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kBlock, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        return nb;
    });
}

cudaError_t wavefront_info(uint32_t mode, bool lights, int minb, KernelInfo* out) {
    return dispatch(mode, lights, minb, [&](auto kern) -> cudaError_t {
        cudaFuncAttributes a;
        cudaError_t e = cudaFuncGetAttributes(&a, kern);
        if (e != cudaSuccess) return e;
        out->registers = a.numRegs; out->max_threads = a.maxThreadsPerBlock; out->const_bytes = (int)a.constSizeBytes; out->local_bytes = (int)a.localSizeBytes;
        snprintf(out->name, sizeof out->name, "rt_wavefront_kernel<%d,%s,%s>", mode == MODE_TREE ? (minb >= 3 ? 3 : 2) : 2,
                 mode == MODE_TREE ? "MODE_TREE" : mode == MODE_BRUTE ? "MODE_BRUTE" : "MODE_EXACT", lights ? "LIGHTS" : "NO_LIGHTS");
        return cudaSuccess;
    });
}

}  // namespace rtk
