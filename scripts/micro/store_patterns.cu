// Microbenchmark: HBM write ceiling for the z_q_is output layout [B][Nq][1024][T] under different work assignments.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_lin(float4 *out, size_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}

// lane = frame.  CTA c owns flattened frames [c*fpc, (c+1)*fpc); warp w: frame group w % FG, channel slice w / FG.
// order 0: stage-major (all channels of stage s, then s+1); VEC = floats per thread per store (1: 32 frames per warp-store,
// 4: 128 frames per warp-store, needs fpc % 128 == 0 alignment friendly T)
template <int VEC>
__global__ void __launch_bounds__(512, 1) k_store(float *out, int T, int total, int fpc, int NS, int nwarps, int interleave) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const int FPW = 32 * VEC;
    const int FG = (fpc + FPW - 1) / FPW;
    const int fg = w % FG, sl = w / FG, nsl = nwarps / FG;
    if (sl >= nsl) return;
    const int g = blockIdx.x * fpc + fg * FPW + lane * VEC;
    const bool ok = (fg * FPW + lane * VEC) < fpc && g < total;
    const int b = g / T, tt = g % T;
    const int chn = 1024 / nsl;
    for (int s = 0; s < NS; ++s) {
        // interleave: slice sl takes channels sl, sl+nsl, ... (adjacent rows written by different warps at the same time)
        float *o = out + (((long long)b * NS + s) * 1024 + (interleave ? sl : sl * chn)) * T + tt;
        const long long step = (long long)(interleave ? nsl : 1) * T;
        if (ok) {
#pragma unroll 16
            for (int ch = 0; ch < chn; ++ch) {
                if (VEC == 1) __stcs(o + ch * step, (float)s);
                else if (VEC == 2) __stcs(reinterpret_cast<float2 *>(o + ch * step), make_float2(1.f, 2.f));
                else __stcs(reinterpret_cast<float4 *>(o + ch * step), make_float4(1.f, 2.f, 3.f, 4.f));
            }
        }
    }
}

int main() {
    const int NS = 8, Bn = 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float *big;
    const size_t bytes = (size_t)Bn * NS * 1024 * 864 * 4;
    CK(cudaMalloc(&big, bytes));
    auto timeit = [&](auto fn) { float best = 1e9; for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); } return best; };
    {
        float ms = timeit([&] { cudaMemsetAsync(big, 0, bytes); });
        printf("cudaMemset %zu MB: %.1f us %.0f GB/s\n", bytes >> 20, ms * 1e3, bytes / ms / 1e6);
        for (int g : {148, 296, 592, 1184}) {
            ms = timeit([&] { k_lin<<<g, 512>>>((float4 *)big, bytes / 16); });
            printf("linear float4 grid %d x512: %.1f us %.0f GB/s\n", g, ms * 1e3, bytes / ms / 1e6);
        }
    }
    struct Cfg { int T, fpc, grid, nw, vec, il; };
    const Cfg cfgs[] = {
        {864, 96, 144, 3, 1, 0}, {864, 96, 144, 6, 1, 0}, {864, 96, 144, 12, 1, 0}, {864, 96, 144, 6, 1, 1}, {864, 96, 144, 12, 1, 1},
        {864, 128, 108, 4, 1, 0}, {864, 128, 108, 8, 1, 0}, {864, 128, 108, 16, 1, 0}, {864, 128, 108, 8, 1, 1}, {864, 128, 108, 16, 1, 1},
        {864, 128, 108, 1, 4, 0}, {864, 128, 108, 2, 4, 0}, {864, 128, 108, 4, 4, 0}, {864, 128, 108, 8, 4, 0}, {864, 128, 108, 16, 4, 0}, {864, 128, 108, 4, 4, 1}, {864, 128, 108, 16, 4, 1},
        {864, 64, 216, 2, 1, 0}, {864, 64, 216, 4, 1, 0}, {864, 64, 216, 8, 1, 0}, {864, 64, 216, 8, 1, 1},
        {864, 32, 432, 1, 1, 0}, {864, 32, 432, 4, 1, 0}, {864, 32, 432, 8, 1, 0}, {864, 32, 432, 8, 1, 1},
        {862, 94, 148, 3, 1, 0}, {862, 94, 148, 6, 1, 1}, {862, 94, 148, 12, 1, 1}, {862, 128, 108, 4, 1, 0}, {862, 128, 108, 8, 1, 1}, {862, 128, 108, 16, 1, 1},
        {862, 128, 108, 2, 2, 0}, {862, 128, 108, 4, 2, 0}, {862, 128, 108, 8, 2, 1},
    };
    for (const Cfg &c : cfgs) {
        const int total = Bn * c.T;
        const float ms = timeit([&] {
            if (c.vec == 1) k_store<1><<<c.grid, 512>>>(big, c.T, total, c.fpc, NS, c.nw, c.il);
            else if (c.vec == 2) k_store<2><<<c.grid, 512>>>(big, c.T, total, c.fpc, NS, c.nw, c.il);
            else k_store<4><<<c.grid, 512>>>(big, c.T, total, c.fpc, NS, c.nw, c.il);
        });
        printf("T=%d fpc=%3d grid=%3d warps=%2d vec=%d interleave=%d: %.1f us  %.0f GB/s\n", c.T, c.fpc, c.grid, c.nw, c.vec, c.il, ms * 1e3,
               (double)total * NS * 4096 / ms / 1e6);
    }
    return 0;
}
