// Microbenchmark: store throughput of ONE SM for the epilogue's pattern (lane = frame, one st.global.cs.f32 per lane per channel row,
// 4 warps = 128 consecutive frames) versus 16-byte stores, as a function of how many SMs store at the same time.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_per_sm store_per_sm.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int VEC>
__global__ void __launch_bounds__(128, 1) k(float *out, int T, int rows_per_cta, int reps) {
    const int lane_frame = threadIdx.x * VEC;  // 128 threads: 128 (VEC = 1) or 512 (VEC = 4) consecutive frames
    for (int r = 0; r < reps; ++r) {
        float *o = out + (size_t)blockIdx.x * rows_per_cta * T + (size_t)(r % 4) * 512 + lane_frame;
#pragma unroll 8
        for (int ch = 0; ch < rows_per_cta; ++ch) {
            if (VEC == 1) __stcs(o + (size_t)ch * T, 1.0f);
            else __stcs(reinterpret_cast<float4 *>(o + (size_t)ch * T), make_float4(1.f, 2.f, 3.f, 4.f));
        }
    }
}

int main() {
    const int T = 5168, rows = 1024, reps = 64;
    float *buf;
    CK(cudaMalloc(&buf, (size_t)148 * rows * T * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int vec : {1, 4})
        for (int n : {1, 8, 37, 74, 148}) {
            float best = 1e9;
            for (int it = 0; it < 4; ++it) {
                cudaEventRecord(e0);
                if (vec == 1) k<1><<<n, 128>>>(buf, T, rows, reps); else k<4><<<n, 128>>>(buf, T, rows, reps);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
            }
            const double bytes = (double)n * rows * reps * 128 * vec * 4;
            printf("vec %d, %3d SMs storing: %.1f us, %.0f GB/s total, %.1f B/clk/SM (at %d MHz)\n", vec, n, best * 1e3, bytes / best / 1e6,
                   bytes / n / (best * 1e-3) / (clk_khz * 1e3), clk_khz / 1000);
        }
    return 0;
}
