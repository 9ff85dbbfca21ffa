// rtb200_kernels.cuh — parameter blocks and launch wrappers shared by rtb200_kernels.cu and rtb200_api.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtb200.h"
#include "rtb200_device.cuh"

namespace rtk {

constexpr int kBlock = 256;        // threads per CTA of the trace kernel
constexpr int kCtasPerSm = 2;        // lane-autonomous kernel (rt_trace_kernel)
constexpr int kMaxCand = 24;       // per-lane candidate slots in shared memory (lanes kernel)
#ifndef RT_CLUSTER_K
#define RT_CLUSTER_K 4
#endif
constexpr int kClusterK = RT_CLUSTER_K;       // spheres per second-level cluster (slots, padded)
#ifndef RT_WF_MAXCLUS
#define RT_WF_MAXCLUS 32
#endif
#ifndef RT_WF_MAXCAND
#define RT_WF_MAXCAND 16
#endif
constexpr int kWfMaxClus = RT_WF_MAXCLUS;     // wavefront kernel: per-thread list of candidate clusters / first-level candidates
constexpr int kWfMaxCand = RT_WF_MAXCAND;     // wavefront kernel, two-level mode: per-thread list of second-level (sphere) candidates

// 32-byte material record (device copy of the material half of rt_sphere)
struct DevMat { float r, g, b; uint32_t kind; double param; int32_t tex; int32_t pad; };
static_assert(sizeof(DevMat) == 32, "DevMat must be 32 bytes");

// One pending light test (raytracer.rs:99-114): the vertex it belongs to and the partial sum over the lights.
struct ShadowFrame {
    double px, py, pz;      // hit point = origin of the shadow rays and of the scattered ray
    double ndx, ndy, ndz;   // scattered direction (continuation of the main path)
    float ar, ag, ab;       // albedo of the vertex
    float sr, sg, sb;       // sum over lights of albedo * ray_color(light_ray, 2, 1)
    uint32_t li, code, is_light, pad;
};
static_assert(sizeof(ShadowFrame) == 88, "ShadowFrame layout");

struct TraceParams {
    // scene, resident in HBM
    const float4*       filt;      // n_pairs*2 float4: {cx0,cx1,cy0,cy1},{cz0,cz1,nk0,nk1}, recentred f32 filter records
    const double4*      geo;       // n: {cx,cy,cz,radius} exact f64
    const DevMat*       mat;       // n
    const rtd::DevTex*  tex;       // n_tex
    uint32_t n, n_pairs;
    uint32_t n_lights;
    uint32_t scene_in_smem;        // wavefront kernel: bit0 sfilt, bit1 geo, bit2 mat staged into shared memory; lanes kernel: 0/1 = geo+mat
    uint32_t two_level;            // 1: `filt` holds CLUSTER bounding-sphere records, `sfilt`/`orig` the member spheres (8 slots per cluster)
    uint32_t n_clusters;
    const float4*   sfilt;         // n_clusters*kClusterK float4 (kClusterK/2 pairs per cluster), filter records of the member spheres
    const uint16_t* orig;          // n_clusters*kClusterK: slot -> sphere index (0xffff = padding)
    const float*    cmeta;         // n_clusters: |c| of the cluster bound in the recentred frame, rounded up
    double gx, gy, gz;             // recentring offset of the filter frame
    float  er_coef;                // per-ray error coefficient (see DESIGN.md "filter soundness")
    rt_camera cam;
    uint32_t width, height, spp, max_depth;
    uint32_t sky_mode;
    rtd::DevTex sky;
    uint32_t key0, key1;           // Philox key = seed
    // work of this launch: samples [s0, s0+s_count) of every pixel of the shard
    uint32_t s0, s_count;
    uint32_t npix_local, rows_local;
    int32_t  rank, world;
    uint32_t band_rows;
    uint32_t total_work;           // npix_local * s_count
    unsigned int* work_counter;
    float4*  samplebuf;            // [s_count][npix_local] per-sample radiance (w = rays of the sample)
    uint32_t* stack;               // [max_depth][stack_stride] per-lane albedo codes
    uint32_t stack_stride;
    const uint32_t* lights;        // sphere indices of the Light spheres in list order (find_lights, raytracer.rs:220-229)
    ShadowFrame* frames;           // [max_shadow][stack_stride], only when n_lights > 0
    uint32_t max_shadow;           // nested light-test frames per path (a level nests with probability <= n_lights*0.1)
    float* lterm;                  // [2 levels][3][stack_stride] light terms of the first two path levels
    unsigned long long* stat;      // [0]=rays [1]=candidates [2]=overflows [3]=samples [5]=shadow-stack overflows
};

struct ResolveParams {
    const float4* samplebuf;
    float*   accum;        // [npix_local][3] running f32 sums in sample order
    uint32_t npix_local, s_count;
    uint32_t first, last;  // first batch zeroes accum, last batch writes outputs
    uint32_t spp;
    float*   out_linear;   // [npix_local][3] or null
    uint8_t* out_rgb8;     // [npix_local][3] or null
};

size_t trace_smem_bytes(uint32_t n, uint32_t n_pairs, bool scene_in_smem);
cudaError_t launch_trace(const TraceParams& p, int grid, size_t smem, bool exact, cudaStream_t st);
size_t wavefront_smem_bytes(uint32_t n, uint32_t n_pairs, uint32_t n_clusters, bool two_level, uint32_t smem_mask, int block);
cudaError_t launch_wavefront(const TraceParams& p, int grid, size_t smem, int minb, bool exact, cudaStream_t st);
int wavefront_max_ctas_per_sm(size_t smem, int minb);
cudaError_t launch_resolve(const ResolveParams& p, cudaStream_t st);
cudaError_t trace_configure(int device, int* sm_count, size_t* max_smem_optin);

// single-thread probes of the device routines (known-answer tests)
cudaError_t probe_sphere_hit(const double* in /*12*/, double* out /*9*/, cudaStream_t st);
cudaError_t probe_refract(const double* in /*7*/, double* out /*3*/, cudaStream_t st);
cudaError_t probe_reflectance(const double* in /*2*/, double* out /*1*/, cudaStream_t st);
cudaError_t probe_sky(const double* in /*3*/, uint32_t mode, float* out /*3*/, cudaStream_t st);
cudaError_t probe_get_ray(const rt_camera* cam_dev, const double* in /*2*/, double* out /*6*/, cudaStream_t st);
cudaError_t probe_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t kind, uint32_t n, double* out, cudaStream_t st);
cudaError_t probe_quantise(const float* in, uint32_t n, uint8_t* out, cudaStream_t st);
cudaError_t probe_sphere_uv(const double* in /*3n*/, uint32_t n, double* out /*2n*/, cudaStream_t st);

}  // namespace rtk
