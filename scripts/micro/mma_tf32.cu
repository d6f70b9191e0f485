// Microbenchmark: throughput of legacy warp-level mma.sync.m16n8k8 TF32 (and m16n8k8 via 3xTF32) on sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(512, 1) k(long long *out, float *sink, int iters) {
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    float d[NACC][4];
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NACC; ++j) mma_tf32(d[j], a, b);
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += d[j][i];
    if (s == 123.456f) sink[0] = s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

template <int NACC> void run(int warps) {
    long long *d; float *sink; cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 4);
    const int iters = 4000;
    k<NACC><<<148, warps * 32>>>(d, sink, iters);
    k<NACC><<<148, warps * 32>>>(d, sink, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h[0];
    double mmas = (double)iters * NACC * warps;  // per SM
    printf("warps=%2d acc=%d: %.2f cycles per MMA per SM  -> %.0f FMA/clk/SM  [%s]\n", warps, NACC, cyc / mmas, 16.0 * 8 * 8 * mmas / cyc, cudaGetErrorString(e));
    cudaFree(d); cudaFree(sink);
}

int main() {
    run<1>(16); run<2>(16); run<4>(16); run<8>(16);
    run<4>(4); run<8>(8);
    return 0;
}
