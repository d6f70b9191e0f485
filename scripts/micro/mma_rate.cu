// Microbenchmark: issue/execution rate of tcgen05.mma kind::tf32 (SS, K-major no-swizzle tiles) from one thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vrvq_b200/csrc/common.cuh"
using namespace vrvq;
namespace vrvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int check_device() { return 0; } }

constexpr uint64_t DESC_128 = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(2048 >> 4) << 16);

template <int N, int COMMIT_EVERY, int FENCE = 0>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, w = t >> 5;
    for (int i = t; i < 16384; i += 128) reinterpret_cast<float *>(sm)[i] = 1.0f;
    if (t == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_mbar_init(); }
    if (w == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    fence_proxy_async();
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const uint32_t tb = slot;
    if (t == 0) {
        const uint64_t a = DESC_128 | (smem_u32(sm) >> 4), b = DESC_128 | ((smem_u32(sm) + 16384) >> 4);
        const uint32_t id = umma_idesc_tf32(128, N);
        uint32_t ph = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (FENCE & 1) fence_proxy_async();
            if (FENCE & 2) tmem_fence_after_sync();
            if (FENCE & 4) { while (!mbar_try_wait(&bar2, 1)) {} }
            if (FENCE & 8) {
                uint32_t ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(1u) : "memory");
                } while (!ok);
            }
            if (FENCE & 16) { volatile uint64_t *vb = &bar2; (void)*vb; }
            if (FENCE & 32) continue;
            umma_tf32(tb + (i & 1) * 256, a + ((i & 3) * 256), b + ((i & 3) * 256), id, true);
            if (COMMIT_EVERY > 0 && (i % COMMIT_EVERY) == COMMIT_EVERY - 1) umma_commit(&bar2);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        for (int spin = 0; spin < (1 << 24) && !mbar_try_wait(&bar, ph); ++spin) {}
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tmem_fence_before_sync(); __syncthreads();
    if (w == 0) { tmem_fence_after_sync(); tmem_dealloc(tb, 512); }
}

template <int N, int CE, int F = 0> void run(const char *name) {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k<N, CE, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    const int iters = 512;
    k<N, CE, F><<<148, 128, 131072>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-40s issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
    cudaFree(d);
}
int main() {
    run<64, 0>("N=64  no intermediate commits");
    run<128, 0>("N=128 no intermediate commits");
    run<256, 0>("N=256 no intermediate commits");
    run<128, 4>("N=128 commit every 4");
    run<128, 1>("N=128 commit every MMA");
    run<64, 1>("N=64 commit every MMA");
    run<16, 0>("N=16 no intermediate commits");
    run<16, 0, 1>("N=16 + fence.proxy.async each");
    run<16, 0, 2>("N=16 + tcgen05.fence::after each");
    run<16, 0, 4>("N=16 + mbarrier try_wait (true) each");
    run<128, 0, 1>("N=128 + fence.proxy.async each");
    run<16, 0, 8>("N=16 + mbarrier test_wait each");
    run<16, 0, 16>("N=16 + volatile smem load each");
    run<16, 0, 4 + 32>("try_wait only (no MMA)");
    run<16, 0, 8 + 32>("test_wait only (no MMA)");
    run<128, 0, 8>("N=128 + mbarrier test_wait each");
    return 0;
}
