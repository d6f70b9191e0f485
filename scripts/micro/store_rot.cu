// Microbenchmark: write ceiling of the tensor-core kernel's z_q_is store pattern (lane = frame, one 128-channel unit at a
// time, channel 128j + 4i + p, sector-aligned row starts) as a function of the ORDER in which the CTAs walk the units.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_rot store_rot.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// rot: 0 = every CTA walks j = 0..7; 1 = CTA c starts at unit c % 8; 2 = starts at unit (tile-in-item) % 8; 3 = (c / 9) % 8 (per item)
// nw: 4 or 8 store warps (8: warps w and w+4 share a lane quarter and alternate units)
template <int OP>
__global__ void __launch_bounds__(512, 1) k_tile(float *out, int T, int adv, int tiles_per_item, int NS, int rot, int nw, int skew, int amask) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= nw) return;
    const int b = blockIdx.x / tiles_per_item, ti = blockIdx.x % tiles_per_item;
    const int t0 = ti * adv;
    const bool last = ti == tiles_per_item - 1;
    const int r = 32 * (w & 3) + lane;
    const int r0 = rot == 1 ? blockIdx.x % 8 : rot == 2 ? ti % 8 : rot == 3 ? b % 8 : 0;
    if (skew) {  // desynchronise the CTAs a little
        const long long t_end = clock64() + (long long)(blockIdx.x % 8) * skew;
        while (clock64() < t_end) {}
    }
    for (int s = 0; s < NS; ++s) {
        for (int jj = 0; jj < 8; ++jj) {
            if (nw == 8 && ((jj & 1) != (w >> 2))) continue;
            const int j = (jj + r0) & 7;
            float *row0 = out + (((long long)b * NS + s) * 1024 + 128 * j) * T;
            for (int p = 0; p < 4; ++p) {
                // shift so that the first lane of the class starts on a 32-byte sector
                const long long base = (((long long)b * NS + s) * 1024 + 128 * j + p) * (long long)T + t0;
                const int dl = (int)(base & amask);          // frames past a sector boundary
                const int frame = t0 - dl + r;
                const bool ok = frame >= 0 && (last ? frame < T : r < adv);
                if (ok) {
                    float *o = row0 + (long long)p * T + frame;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        float *q = o + (long long)(4 * i) * T;
                        if (OP == 0) __stcs(q, (float)s);
                        else if (OP == 1) *q = (float)s;
                        else if (OP == 2) __stcg(q, (float)s);
                        else if (OP == 3) __stwt(q, (float)s);
                        else if (OP == 4) asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(q), "f"((float)s) : "memory");
                        else asm volatile("st.global.L1::evict_last.f32 [%0], %1;" ::"l"(q), "f"((float)s) : "memory");
                    }
                }
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
    auto timeit = [&](auto fn) { float best = 1e9; for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms); } return best; };
    for (int T : {862, 864}) {
        for (int adv : {96, 120}) {
            const int tpi = (T + adv - 1) / adv;
            for (int nw : {4, 8}) {
                for (int rot = 0; rot < 4; ++rot) {
                    for (int op = 0; op < 6; ++op) { const int skew = 0, amask = 7;
                        if (rot != 0 || nw != 4) continue;
                        const float ms = timeit([&] {
                            switch (op) {
                                case 0: k_tile<0><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                                case 1: k_tile<1><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                                case 2: k_tile<2><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                                case 3: k_tile<3><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                                case 4: k_tile<4><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                                default: k_tile<5><<<Bn * tpi, 512>>>(big, T, adv, tpi, NS, rot, nw, skew, amask); break;
                            }
                        });
                        CK(cudaGetLastError());
                        printf("T=%d adv=%3d grid=%3d warps=%d rot=%d op=%d: %.1f us  %.0f GB/s\n", T, adv, Bn * tpi, nw, rot, op, ms * 1e3,
                               (double)Bn * T * NS * 4096 / ms / 1e6);
                    }
                }
            }
        }
    }
    return 0;
}
