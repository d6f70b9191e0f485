// Microbenchmark: per-SM throughput of tcgen05.ld / tcgen05.st (32x32b.x8) used as per-thread scratch, and of
// LDS.128 broadcast patterns.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vrvq_b200/csrc/common.cuh"
using namespace vrvq;
namespace vrvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int check_device() { return 0; } }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k_tmem(long long *out, int iters) {
    __shared__ uint32_t slot;
    __shared__ __align__(16) float sm[8192];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += 512) sm[i] = (float)i;
    if (w == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const uint32_t base = slot;
    const uint32_t tacc = base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * 64);
    uint32_t a[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    for (int g = 0; g < 8; ++g) tmem_st8(tacc + 8 * g, a);
    tmem_wait_st();
    __syncthreads();
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // ld only: 64 columns per thread per iteration
#pragma unroll
            for (int g = 0; g < 8; ++g) { uint32_t v[8]; tmem_ld8(tacc + 8 * g, v); tmem_wait_ld(); acc += __uint_as_float(v[0]) + __uint_as_float(v[7]); }
        } else if (MODE == 1) {  // ld + st
#pragma unroll
            for (int g = 0; g < 8; ++g) { uint32_t v[8]; tmem_ld8(tacc + 8 * g, v); tmem_wait_ld(); v[0] += 1; tmem_st8(tacc + 8 * g, v); }
            tmem_wait_st();
        } else if (MODE == 2) {  // ld pipelined: issue 8 loads, one wait
            uint32_t v[8][8];
#pragma unroll
            for (int g = 0; g < 8; ++g) tmem_ld8(tacc + 8 * g, v[g]);
            tmem_wait_ld();
#pragma unroll
            for (int g = 0; g < 8; ++g) acc += __uint_as_float(v[g][0]) + __uint_as_float(v[g][7]);
        } else if (MODE == 3) {  // LDS.128 warp-uniform broadcast
#pragma unroll
            for (int g = 0; g < 16; ++g) { float4 x = *reinterpret_cast<const float4 *>(&sm[((it + g) & 255) * 4 + w * 1024 / 4 * 0]); acc += x.x + x.w; }
        } else if (MODE == 4) {  // LDS.128 half-warp uniform (2 distinct)
#pragma unroll
            for (int g = 0; g < 16; ++g) { float4 x = *reinterpret_cast<const float4 *>(&sm[(((it + g) & 127) * 2 + (lane >> 4)) * 8]); acc += x.x + x.w; }
        } else if (MODE == 5) {  // LDS.128 quarter-warp uniform (4 distinct, consecutive 32 B rows)
#pragma unroll
            for (int g = 0; g < 16; ++g) { float4 x = *reinterpret_cast<const float4 *>(&sm[(((it + g) & 63) * 4 + (lane >> 3)) * 8]); acc += x.x + x.w; }
        } else if (MODE == 6) {  // LDS.128 all lanes distinct (512 B)
#pragma unroll
            for (int g = 0; g < 16; ++g) { float4 x = *reinterpret_cast<const float4 *>(&sm[(((it + g) & 15) * 32 + lane) * 4]); acc += x.x + x.w; }
        } else if (MODE == 7) {  // LDS.64 half-warp uniform
#pragma unroll
            for (int g = 0; g < 16; ++g) { float2 x = *reinterpret_cast<const float2 *>(&sm[(((it + g) & 127) * 2 + (lane >> 4)) * 2]); acc += x.x + x.y; }
        } else if (MODE == 8) {  // LDS.32 warp-uniform
#pragma unroll
            for (int g = 0; g < 16; ++g) { acc += sm[(it + g) & 1023]; }
        } else if (MODE == 9) {  // SHFL
#pragma unroll
            for (int g = 0; g < 16; ++g) { acc += __shfl_sync(0xffffffffu, acc, (g + it) & 31); }
        }
    }
    long long t1 = clock64();
    if (acc == 12345.678f) out[1] = 1;
    __syncthreads();
    if (tid == 0) out[0] = t1 - t0;
    tmem_fence_before_sync(); __syncthreads();
    if (w == 0) { tmem_fence_after_sync(); tmem_dealloc(base, 256); }
}

template <int MODE> void run(const char *name, double bytes_per_iter_per_cta, double instr_per_iter_per_warp) {
    long long *d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    const int iters = 2000;
    k_tmem<MODE><<<148, 512>>>(d, iters);
    k_tmem<MODE><<<148, 512>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0] / iters;
    printf("%-44s %8.1f cyc/iter  %7.1f B/clk/SM  %6.2f cyc per warp-instr (16 warps)  [%s]\n", name, cyc, bytes_per_iter_per_cta / cyc,
           cyc / (instr_per_iter_per_warp * 16), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<0>("tcgen05.ld x8 + wait each (8/iter/thread)", 512 * 64 * 4.0, 8);
    run<1>("tcgen05.ld x8 + st x8 (8+8/iter/thread)", 2 * 512 * 64 * 4.0, 16);
    run<2>("tcgen05.ld x8 pipelined (8/iter/thread)", 512 * 64 * 4.0, 8);
    run<3>("LDS.128 warp-uniform", 16 * 16 * 16.0, 16);
    run<4>("LDS.128 half-warp uniform (2 addr)", 16 * 16 * 32.0, 16);
    run<5>("LDS.128 quarter-warp uniform (4 addr)", 16 * 16 * 64.0, 16);
    run<6>("LDS.128 all lanes distinct", 16 * 16 * 512.0, 16);
    run<7>("LDS.64 half-warp uniform (2 addr)", 16 * 16 * 16.0, 16);
    run<8>("LDS.32 warp-uniform", 16 * 16 * 4.0, 16);
    run<9>("SHFL", 16 * 16 * 128.0, 16);
    return 0;
}
