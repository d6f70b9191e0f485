// Microbenchmark: issue->completion latency of cp.async.bulk global->shared (one 36 KB piece, as the encode kernel's weight
// ring uses) when 148 CTAs fetch the SAME piece at the same time, split into 1/2/4/8 concurrent copies, vs. per-CTA pieces.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vrvq_b200/csrc/common.cuh"
using namespace vrvq;
namespace vrvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int check_device() { return 0; } }

template <int NSPLIT, bool SAME>
__global__ void __launch_bounds__(512, 1) k(const float *src, long long *out, int iters, int piece_floats) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    long long tot = 0;
    for (int it = 0; it < iters; ++it) {
        const float *s = src + (size_t)((SAME ? 0 : blockIdx.x) * 8 + (it & 7)) * piece_floats;
        __syncthreads();
        long long t0 = clock64();
        if (tid == 0) mbar_arrive_expect_tx(&bar, piece_floats * 4);
        __syncthreads();
        if (tid < NSPLIT) {
            const int chunk = piece_floats / NSPLIT;
            bulk_g2s(sm + tid * chunk, s + tid * chunk, chunk * 4, &bar);
        }
        mbar_wait(&bar, it & 1);
        long long t1 = clock64();
        tot += t1 - t0;
    }
    if (tid == 0) out[blockIdx.x] = tot;
}

template <int NSPLIT, bool SAME> void run(const float *src, const char *name) {
    long long *d; cudaMalloc(&d, 148 * 8);
    const int iters = 200, piece = 9216;
    cudaFuncSetAttribute(k<NSPLIT, SAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k<NSPLIT, SAME><<<148, 512, 200 * 1024>>>(src, d, iters, piece);
    k<NSPLIT, SAME><<<148, 512, 200 * 1024>>>(src, d, iters, piece);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0; for (int i = 0; i < 148; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
    printf("%-46s mean %.0f cycles  max %.0f cycles per 36 KB piece  (%.1f B/clk/SM) [%s]\n", name, avg / 148 / iters, (double)mx / iters,
           36864.0 / (avg / 148 / iters), cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    float *src; cudaMalloc(&src, (size_t)148 * 8 * 9216 * 4 + 4096); cudaMemset(src, 0, (size_t)148 * 8 * 9216 * 4);
    run<1, true>(src, "same piece, 1 bulk copy");
    run<2, true>(src, "same piece, 2 bulk copies");
    run<4, true>(src, "same piece, 4 bulk copies");
    run<8, true>(src, "same piece, 8 bulk copies");
    run<16, true>(src, "same piece, 16 bulk copies");
    run<1, false>(src, "per-CTA piece, 1 bulk copy");
    run<4, false>(src, "per-CTA piece, 4 bulk copies");
    return 0;
}
