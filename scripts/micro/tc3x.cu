// Microbenchmarks for the tensor-core (tcgen05, kind::tf32) formulation of the RVQ hot path.
//   A  accuracy of a 3xTF32 in_proj  z_e[128 frames x 64] = z[128 x 1024] W^T  against binary64, for several
//      accumulator-drain intervals (the tensor core rounds its fp32 accumulator toward zero on every k-step)
//   B  out_proj shape M=128 N=256 K=8 (3 k-steps hi/lo), checks the TMEM lane/column mapping
//   C  store ceiling of the epilogue pattern (lane = frame, st.global.b32, rows [ch][T])
//   D  load ceiling of the lane = frame ld.global.b32 pattern
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc3x tc3x.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../vrvq_b200/csrc/common.cuh"
using namespace vrvq;
namespace vrvq { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } int check_device() { return 0; } }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
    return d;                // layout type 0 = no swizzle, base offset 0
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (int i = 0; i < (1 << 24); ++i)
        if (mbar_try_wait(bar, parity)) return;
    printf("mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
}

// ---------------------------------------------------------------- A: in_proj accuracy
// z [1024][128] (channel-major), W [64][1024]; out [128][64].
__global__ void __launch_bounds__(128, 1) k_inproj(const float *z, const float *W, float *out, int drain_every, int split_lo, int rn, int terms) {
    extern __shared__ __align__(128) float dsm[];
    float *Ahi = dsm, *Alo = dsm + 8 * 128 * 4, *Bhi = dsm + 2 * 8 * 128 * 4, *Blo = Bhi + 8 * 64 * 4;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, w = t >> 5;
    if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (w == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const uint32_t tb = slot;
    float acc[64];
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    uint32_t phase = 0;
    bool fresh = true;
    for (int c = 0; c < 32; ++c) {
        for (int kg = 0; kg < 8; ++kg) {
            float h[4], l[4];
            for (int i = 0; i < 4; ++i) {
                const float x = z[(c * 32 + kg * 4 + i) * 128 + t];
                uint32_t hb;
                if (rn) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x)); else hb = __float_as_uint(x) & 0xFFFFE000u;
                h[i] = __uint_as_float(hb);
                l[i] = x - h[i];
            }
            *reinterpret_cast<float4 *>(&Ahi[(kg * 128 + t) * 4]) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(&Alo[(kg * 128 + t) * 4]) = make_float4(l[0], l[1], l[2], l[3]);
        }
        {
            const int n = t & 63;
            for (int kk = 0; kk < 4; ++kk) {
                const int kg = (t >> 6) * 4 + kk;
                float h[4], l[4];
                for (int i = 0; i < 4; ++i) {
                    const float x = W[n * 1024 + c * 32 + kg * 4 + i];
                    uint32_t hb;
                    if (rn) asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x)); else hb = __float_as_uint(x) & 0xFFFFE000u;
                    h[i] = __uint_as_float(hb);
                    l[i] = x - h[i];
                }
                *reinterpret_cast<float4 *>(&Bhi[(kg * 64 + n) * 4]) = make_float4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<float4 *>(&Blo[(kg * 64 + n) * 4]) = make_float4(l[0], l[1], l[2], l[3]);
            }
        }
        fence_proxy_async();
        tmem_fence_before_sync();
        __syncthreads();
        if (t == 0) {
            tmem_fence_after_sync();
            const uint32_t id = idesc_tf32(128, 64);
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ah = umma_desc(smem_u32(Ahi) + ks * 2 * 2048, 2048, 128);
                const uint64_t al = umma_desc(smem_u32(Alo) + ks * 2 * 2048, 2048, 128);
                const uint64_t bh = umma_desc(smem_u32(Bhi) + ks * 2 * 1024, 1024, 128);
                const uint64_t bl = umma_desc(smem_u32(Blo) + ks * 2 * 1024, 1024, 128);
                const bool first = fresh && ks == 0;
                const uint32_t dlo = split_lo ? tb + 64 : tb;
                umma_tf32(tb, ah, bh, id, !first);
                if (terms == 3) {
                    umma_tf32(dlo, ah, bl, id, split_lo ? !first : true);
                    umma_tf32(dlo, al, bh, id, true);
                }
            }
            umma_commit(&bar);
        }
        fresh = false;
        mbar_wait_bounded(&bar, phase); phase ^= 1;  // this chunk's MMAs are complete: smem reusable, accumulator readable
        tmem_fence_after_sync();
        if ((c + 1) % drain_every == 0 || c == 31) {
            const uint32_t ta = tb + ((uint32_t)(32 * w) << 16);
            for (int g = 0; g < 8; ++g) {
                uint32_t v[8], u[8];
                tmem_ld8(ta + 8 * g, v);
                if (split_lo && terms == 3) tmem_ld8(ta + 64 + 8 * g, u);
                tmem_wait_ld();
                for (int i = 0; i < 8; ++i) {
                    float x = __uint_as_float(v[i]);
                    if (split_lo && terms == 3) x += __uint_as_float(u[i]);
                    acc[8 * g + i] += x;
                }
            }
            fresh = true;
        }
    }
    for (int i = 0; i < 64; ++i) out[t * 64 + i] = acc[i];
    tmem_fence_before_sync(); __syncthreads();
    if (w == 0) { tmem_fence_after_sync(); tmem_dealloc(tb, 128); }
}

// ---------------------------------------------------------------- B: out_proj shape
// q [128][8], Wo [256][8]; out [128][256] = q Wo^T
__global__ void __launch_bounds__(128, 1) k_outproj(const float *q, const float *Wo, float *out) {
    __shared__ __align__(128) float Ahi[2 * 128 * 4], Alo[2 * 128 * 4], Bhi[2 * 256 * 4], Blo[2 * 256 * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, w = t >> 5;
    if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (w == 0) { tmem_alloc(&slot, 256); tmem_relinquish(); }
    tmem_fence_before_sync(); __syncthreads(); tmem_fence_after_sync();
    const uint32_t tb = slot;
    for (int kg = 0; kg < 2; ++kg) {
        float h[4], l[4];
        for (int i = 0; i < 4; ++i) { const float x = q[t * 8 + kg * 4 + i]; h[i] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); l[i] = x - h[i]; }
        *reinterpret_cast<float4 *>(&Ahi[(kg * 128 + t) * 4]) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4 *>(&Alo[(kg * 128 + t) * 4]) = make_float4(l[0], l[1], l[2], l[3]);
        for (int nn = 0; nn < 2; ++nn) {
            const int n = t + 128 * nn;
            for (int i = 0; i < 4; ++i) { const float x = Wo[n * 8 + kg * 4 + i]; h[i] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); l[i] = x - h[i]; }
            *reinterpret_cast<float4 *>(&Bhi[(kg * 256 + n) * 4]) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4 *>(&Blo[(kg * 256 + n) * 4]) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
    fence_proxy_async();
    tmem_fence_before_sync();
    __syncthreads();
    if (t == 0) {
        tmem_fence_after_sync();
        const uint32_t id = idesc_tf32(128, 256);
        const uint64_t ah = umma_desc(smem_u32(Ahi), 2048, 128), al = umma_desc(smem_u32(Alo), 2048, 128);
        const uint64_t bh = umma_desc(smem_u32(Bhi), 4096, 128), bl = umma_desc(smem_u32(Blo), 4096, 128);
        umma_tf32(tb, al, bh, id, false);
        umma_tf32(tb, ah, bl, id, true);
        umma_tf32(tb, ah, bh, id, true);
        umma_commit(&bar);
    }
    mbar_wait_bounded(&bar, 0);
    tmem_fence_after_sync();
    const uint32_t ta = tb + ((uint32_t)(32 * w) << 16);
    for (int g = 0; g < 32; ++g) {
        uint32_t v[8];
        tmem_ld8(ta + 8 * g, v);
        tmem_wait_ld();
        for (int i = 0; i < 8; ++i) out[t * 256 + 8 * g + i] = __uint_as_float(v[i]);
    }
    tmem_fence_before_sync(); __syncthreads();
    if (w == 0) { tmem_fence_after_sync(); tmem_dealloc(tb, 256); }
}

// ---------------------------------------------------------------- C: store ceiling (lane = frame)
// out [B][NS][1024][T]; CTA c owns flattened frames [c*fpc, (c+1)*fpc); warp w: frame group w % FG, channel slice w / FG
template <int CS>
__global__ void __launch_bounds__(512, 1) k_store(float *out, int T, int total, int fpc, int NS, int nwarps) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const int FG = (fpc + 31) / 32;
    const int fg = w % FG, sl = w / FG, nsl = nwarps / FG;
    if (sl >= nsl) return;
    const int g = blockIdx.x * fpc + fg * 32 + lane;
    const bool ok = (fg * 32 + lane) < fpc && g < total;
    const int b = g / T, tt = g % T;
    const int chn = 1024 / nsl;
    for (int s = 0; s < NS; ++s) {
        float *o = out + (((long long)b * NS + s) * 1024 + sl * chn) * T + tt;
        const float v = (float)(s + lane);
        if (ok) {
#pragma unroll 16
            for (int ch = 0; ch < chn; ++ch) {
                if (CS) __stcs(o + (long long)ch * T, v); else o[(long long)ch * T] = v;
            }
        }
    }
}
// ---------------------------------------------------------------- D: load ceiling (lane = frame)
__global__ void __launch_bounds__(512, 1) k_load(const float *z, float *sink, int T, int total, int fpc, int nwarps) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const int FG = (fpc + 31) / 32;
    const int fg = w % FG, sl = w / FG, nsl = nwarps / FG;
    if (sl >= nsl) return;
    const int g = blockIdx.x * fpc + fg * 32 + lane;
    const bool ok = (fg * 32 + lane) < fpc && g < total;
    const int b = g / T, tt = g % T;
    const int chn = 1024 / nsl;
    const float *o = z + ((long long)b * 1024 + sl * chn) * T + tt;
    float acc = 0.f;
    if (ok) {
#pragma unroll 16
        for (int ch = 0; ch < chn; ++ch) acc += __ldcs(o + (long long)ch * T);
    }
    if (acc == 1234.5f) sink[0] = acc;
}

static double rms(const std::vector<double> &v) { double s = 0; for (double x : v) s += x * x; return sqrt(s / v.size()); }

int main() {
    // ---- A
    std::vector<float> z(1024 * 128), W(64 * 1024);
    srand(1);
    auto rnd = []() { float u = 0; for (int i = 0; i < 12; ++i) u += (float)rand() / RAND_MAX; return u - 6.0f; };
    for (auto &x : z) x = rnd();
    for (auto &x : W) x = rnd() * 0.03f;
    std::vector<double> ref(128 * 64);
    std::vector<float> chain(128 * 64);
    for (int f = 0; f < 128; ++f)
        for (int n = 0; n < 64; ++n) {
            double s = 0; float c = 0.f;
            for (int k = 0; k < 1024; ++k) { s += (double)z[k * 128 + f] * (double)W[n * 1024 + k]; c = fmaf(z[k * 128 + f], W[n * 1024 + k], c); }
            ref[f * 64 + n] = s; chain[f * 64 + n] = c;
        }
    const double r = rms(ref);
    {
        double me = 0, se = 0;
        for (int i = 0; i < 128 * 64; ++i) { double e = fabs(chain[i] - ref[i]); me = fmax(me, e); se += e * e; }
        printf("A: fp32 FMA chain (host)                      max err / rms %.3e   rms err / rms %.3e\n", me / r, sqrt(se / (128 * 64)) / r);
    }
    float *dz, *dW, *dout;
    CK(cudaMalloc(&dz, z.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dout, 128 * 256 * 4));
    CK(cudaMemcpy(dz, z.data(), z.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> out(128 * 256);
    CK(cudaFuncSetAttribute(k_inproj, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152));
    const int drains[] = {32, 8, 4, 2, 1};
    for (int terms : {1, 3})
        for (int rn : {0, 1})
            for (int split : {0, 1})
                for (int de : drains) {
                    if (terms == 1 && (split || de != 32)) continue;
                    k_inproj<<<1, 128, 49152>>>(dz, dW, dout, de, split, rn, terms);
                    CK(cudaDeviceSynchronize());
                    CK(cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost));
                    double me = 0, se = 0, bias = 0;
                    for (int i = 0; i < 128 * 64; ++i) { double e = out[i] - ref[i]; me = fmax(me, fabs(e)); se += e * e; bias += e * (ref[i] > 0 ? 1 : -1); }
                    printf("A: terms %d rn %d split_lo %d drain every %2d chunks(x32 k): max err / rms %.3e   rms err / rms %.3e   signed-toward-zero bias %.3e\n",
                           terms, rn, split, de, me / r, sqrt(se / (128 * 64)) / r, -bias / (128 * 64) / r);
                }
    // ---- B
    {
        std::vector<float> q(128 * 8), Wo(256 * 8);
        for (auto &x : q) x = rnd();
        for (auto &x : Wo) x = rnd() * 0.3f;
        float *dq, *dWo;
        CK(cudaMalloc(&dq, q.size() * 4)); CK(cudaMalloc(&dWo, Wo.size() * 4));
        CK(cudaMemcpy(dq, q.data(), q.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dWo, Wo.data(), Wo.size() * 4, cudaMemcpyHostToDevice));
        k_outproj<<<1, 128>>>(dq, dWo, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), dout, 128 * 256 * 4, cudaMemcpyDeviceToHost));
        double me = 0, mc = 0, rr = 0;
        for (int f = 0; f < 128; ++f)
            for (int n = 0; n < 256; ++n) {
                double s = 0; float c = 0.f;
                for (int k = 0; k < 8; ++k) { s += (double)q[f * 8 + k] * (double)Wo[n * 8 + k]; c = fmaf(Wo[n * 8 + k], q[f * 8 + k], c); }
                me = fmax(me, fabs(out[f * 256 + n] - s)); mc = fmax(mc, fabs(c - s)); rr += s * s;
            }
        rr = sqrt(rr / (128 * 256));
        printf("B: out_proj 128x256x8 3xTF32: max err / rms %.3e (fp32 chain: %.3e)\n", me / rr, mc / rr);
    }
    // ---- C / D
    {
        const int NS = 8, Bn = 16;
        for (int T : {862, 864}) {
            const int total = Bn * T, fpc = (total + 147) / 148;
            float *big;
            CK(cudaMalloc(&big, (size_t)Bn * NS * 1024 * T * 4));
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int nw : {3, 6, 12, 15}) {
                for (int cs = 0; cs < 2; ++cs) {
                    float best = 1e9;
                    for (int rep = 0; rep < 4; ++rep) {
                        cudaEventRecord(e0);
                        if (cs) k_store<1><<<148, 512>>>(big, T, total, fpc, NS, nw); else k_store<0><<<148, 512>>>(big, T, total, fpc, NS, nw);
                        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                        float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
                    }
                    printf("C: T=%d fpc=%d warps=%2d cs=%d: %.1f us  %.0f GB/s\n", T, fpc, nw, cs, best * 1e3, (double)total * NS * 4096 / best / 1e6);
                }
            }
            for (int nw : {3, 6, 12, 15}) {
                float best = 1e9;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    for (int s = 0; s < NS; ++s) k_load<<<148, 512>>>(big + (size_t)s * Bn * 1024 * T, dout, T, total, fpc, nw);
                    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                    float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
                }
                printf("D: T=%d load warps=%2d: %.1f us per 8 launches  %.0f GB/s\n", T, nw, best * 1e3, (double)total * NS * 4096 / best / 1e6);
            }
            cudaFree(big);
        }
    }
    return 0;
}
