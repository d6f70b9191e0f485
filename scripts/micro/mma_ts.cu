// Micro test: tcgen05.mma kind::tf32 with the A operand in tensor memory (lane = row, one 32-bit column per k), B in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ts mma_ts.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../vrvq_b200/csrc/common.cuh"
using namespace vrvq;
namespace vrvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int check_device() { return 0; } }

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

// A [128][16], B [64][16] -> out [128][64]; K = 16 = two k-steps; A columns at TMEM col 256.., D at col 0
__global__ void __launch_bounds__(128, 1) k(const float *A, const float *B, float *out, long long *cyc) {
    __shared__ __align__(128) float Bs[4 * 64 * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, w = t >> 5;
    if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (w == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const uint32_t tb = slot;
    if (t < 64)
        for (int k4 = 0; k4 < 4; ++k4)
            *reinterpret_cast<float4 *>(&Bs[(k4 * 64 + t) * 4]) = *reinterpret_cast<const float4 *>(&B[t * 16 + k4 * 4]);
    const uint32_t ta = tb + ((uint32_t)(32 * w) << 16) + 256;
    for (int g = 0; g < 2; ++g) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(A[t * 16 + g * 8 + i]);
        tmem_st8(ta + 8 * g, v);
    }
    tmem_wait_st();
    fence_proxy_async();
    tmem_fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tmem_fence_after_sync();
        const uint64_t bd = umma_desc(smem_u32(Bs), 1024, 128);
        const uint32_t id = umma_idesc_tf32(128, 64);
        long long t0 = clock64();
        for (int rep = 0; rep < 64; ++rep) {
            umma_tf32_ts(tb, tb + 256, bd, id, false);
            umma_tf32_ts(tb, tb + 256 + 8, bd + (2048 >> 4), id, true);
        }
        umma_commit(&bar);
        for (int spin = 0; spin < (1 << 24) && !mbar_try_wait(&bar, 0); ++spin) {}
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    for (int spin = 0; spin < (1 << 24) && !mbar_try_wait(&bar, 0); ++spin) {}
    tmem_fence_after_sync();
    const uint32_t td = tb + ((uint32_t)(32 * w) << 16);
    for (int g = 0; g < 8; ++g) {
        uint32_t v[8];
        tmem_ld8(td + 8 * g, v);
        tmem_wait_ld();
        for (int i = 0; i < 8; ++i) out[t * 64 + 8 * g + i] = __uint_as_float(v[i]);
    }
    tmem_fence_before_sync(); __syncthreads();
    if (w == 0) { tmem_fence_after_sync(); tmem_dealloc(tb, 512); }
}

int main() {
    std::vector<float> A(128 * 16), B(64 * 16), out(128 * 64);
    srand(3);
    for (auto &x : A) x = (float)(rand() % 17 - 8);  // small integers: exact in TF32
    for (auto &x : B) x = (float)(rand() % 13 - 6);
    float *dA, *dB, *dO; long long *dc;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, out.size() * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    k<<<1, 128>>>(dA, dB, dO, dc);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 64; ++n) {
            float r = 0; for (int kk = 0; kk < 16; ++kk) r += A[i * 16 + kk] * B[n * 16 + kk];
            if (out[i * 64 + n] != r) { if (bad < 5) printf("mismatch row %d col %d: got %g want %g\n", i, n, out[i * 64 + n], r); ++bad; }
        }
    printf("TS MMA (A in TMEM): %s, %d mismatches of %d, %.1f cycles per N=64 MMA (%s)\n", bad ? "FAIL" : "OK", bad, 128 * 64, (double)c / 128, cudaGetErrorString(e));
    return 0;
}
