// Micro-benchmarks that settle pipe-rate questions on the B200 (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    float2 A = make_float2(a, a), B = make_float2(b, b);
    float2 x0 = make_float2(threadIdx.x, 1), x1 = make_float2(2, 3), x2 = make_float2(4, 5), x3 = make_float2(6, 7);
    float2 x4 = make_float2(8, 9), x5 = make_float2(10, 11), x6 = make_float2(12, 13), x7 = make_float2(14, 15);
    for (int i = 0; i < iters; ++i) {
        x0 = __ffma2_rn(x0, A, B); x1 = __ffma2_rn(x1, A, B); x2 = __ffma2_rn(x2, A, B); x3 = __ffma2_rn(x3, A, B);
        x4 = __ffma2_rn(x4, A, B); x5 = __ffma2_rn(x5, A, B); x6 = __ffma2_rn(x6, A, B); x7 = __ffma2_rn(x7, A, B);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x1.x + x2.x + x3.x + x4.x + x5.x + x6.x + x7.x + x0.y + x1.y + x2.y + x3.y + x4.y + x5.y + x6.y + x7.y;
}
// the production scan block: 8 LDS.128 + 28 FFMA2 + max-reduce + compare, never-taken rare branch
__global__ void k_scan(float* out, const float4* __restrict__ g, int n_pairs, int reps, float thr) {
    extern __shared__ float4 s_filt[];
    for (int i = threadIdx.x; i < 2 * n_pairs; i += blockDim.x) s_filt[i] = g[i];
    __syncthreads();
    float t = threadIdx.x * 1e-3f;
    const float2 dx2 = make_float2(0.3f + t, 0.3f + t), dy2 = make_float2(0.4f, 0.4f), dz2 = make_float2(0.5f, 0.5f);
    const float2 ox2 = make_float2(1.f + t, 1.f + t), oy2 = make_float2(2.f, 2.f), oz2 = make_float2(3.f, 3.f), nod2 = make_float2(0.1f, 0.1f);
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 2
        for (int pp = 0; pp < n_pairs; pp += 4) {
            float2 Dv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 A = s_filt[2 * (pp + q)], B = s_filt[2 * (pp + q) + 1];
                float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w), cz = make_float2(B.x, B.y), nk = make_float2(B.z, B.w);
                float2 bb = __ffma2_rn(cz, dz2, nod2);
                float2 tt = __ffma2_rn(cz, oz2, nk);
                bb = __ffma2_rn(cy, dy2, bb); tt = __ffma2_rn(cy, oy2, tt);
                bb = __ffma2_rn(cx, dx2, bb); tt = __ffma2_rn(cx, ox2, tt);
                Dv[q] = __ffma2_rn(bb, bb, tt);
            }
            float m = fmaxf(fmaxf(fmaxf(Dv[0].x, Dv[0].y), fmaxf(Dv[1].x, Dv[1].y)), fmaxf(fmaxf(Dv[2].x, Dv[2].y), fmaxf(Dv[3].x, Dv[3].y)));
            if (m >= thr) hits += pp;
        }
        thr += 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hits;
}
// scalar-FFMA variant of the same block (14 FFMA per pair instead of 7 FFMA2)
__global__ void k_scan_scalar(float* out, const float4* __restrict__ g, int n_pairs, int reps, float thr) {
    extern __shared__ float4 s_filt[];
    for (int i = threadIdx.x; i < 2 * n_pairs; i += blockDim.x) s_filt[i] = g[i];
    __syncthreads();
    float t = threadIdx.x * 1e-3f;
    const float dx = 0.3f + t, dy = 0.4f, dz = 0.5f, ox = 1.f + t, oy = 2.f, oz = 3.f, nod = 0.1f;
    int hits = 0;
    for (int r = 0; r < reps; ++r) {
#pragma unroll 2
        for (int pp = 0; pp < n_pairs; pp += 4) {
            float m = -1e30f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 A = s_filt[2 * (pp + q)], B = s_filt[2 * (pp + q) + 1];
                float b0 = fmaf(B.x, dz, nod), b1 = fmaf(B.y, dz, nod), t0 = fmaf(B.x, oz, B.z), t1 = fmaf(B.y, oz, B.w);
                b0 = fmaf(A.z, dy, b0); b1 = fmaf(A.w, dy, b1); t0 = fmaf(A.z, oy, t0); t1 = fmaf(A.w, oy, t1);
                b0 = fmaf(A.x, dx, b0); b1 = fmaf(A.y, dx, b1); t0 = fmaf(A.x, ox, t0); t1 = fmaf(A.y, ox, t1);
                m = fmaxf(m, fmaxf(fmaf(b0, b0, t0), fmaf(b1, b1, t1)));
            }
            if (m >= thr) hits += pp;
        }
        thr += 1.0f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)hits;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    float* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, nominal clock %d MHz\n", prop.name, sms, clk_khz / 1000);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        int blocks = sms * warps_per_sm / 8, iters = 20000;
        for (int which = 0; which < 2; ++which) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_ffma<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f); else k_ffma2<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            }
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)blocks * 256 * iters * 8 * (which ? 2 : 1);
            printf("%-6s warps/SM=%2d: %.3f ms  -> %.1f FMA/clk/SM (at %d MHz)   %.2f TFLOP/s\n", which ? "FFMA2" : "FFMA", warps_per_sm, ms,
                   fma / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000, 2 * fma / (ms * 1e-3) / 1e12);
        }
    }
    int n_pairs = 248, reps = 400;
    float4* g; CK(cudaMalloc(&g, n_pairs * 32)); CK(cudaMemset(g, 0, n_pairs * 32));
    for (int warps_per_sm : {8, 16, 24, 32}) {
        int blocks = sms * warps_per_sm / 8;
        for (int which = 0; which < 2; ++which) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_scan<<<blocks, 256, n_pairs * 32>>>(out, g, n_pairs, reps, 1e30f);
                else k_scan_scalar<<<blocks, 256, n_pairs * 32>>>(out, g, n_pairs, reps, 1e30f);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            }
            cudaEventElapsedTime(&ms, e0, e1);
            double tests = (double)blocks * 256 * reps * n_pairs * 2;
            double cyc_per_pair_warp_smsp = (ms * 1e-3) * (clk_khz * 1e3) / ((double)reps * n_pairs * (warps_per_sm / 4.0));
            printf("%-12s warps/SM=%2d: %.3f ms  %.2f Gtests/s  %.2f cycles per (pair, warp) per SMSP  => %.2f Grays/s at 484 spheres\n",
                   which ? "scan(FFMA)" : "scan(FFMA2)", warps_per_sm, ms, tests / (ms * 1e-3) / 1e9, cyc_per_pair_warp_smsp, tests / (ms * 1e-3) / 484 / 1e9);
        }
    }
    return 0;
}
