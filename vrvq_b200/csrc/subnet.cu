// Importance subnet: one block of models/importance_subnet.py:18-34, i.e.
//   y[b,co,t] = bias[co] + sum_{ci,k} W[co,ci,k] * snake(x[b,ci,t+k-1])        (Snake1d -> WNConv1d(k=3, padding=1))
// with snake(v) = v + 1/(alpha+1e-9) * sin(alpha v)^2 (models/layers.py:25-31) and, for the last block, the
// sigmoid of models/importance_subnet.py:43 fused into the epilogue.
//
// Formulation: an implicit GEMM per batch item, M = output channels, N = frames, K = 3*Cin, fp32 on the CUDA cores
// (first correct version of this row; the two big layers, 1024->1024 and 1024->512, are the tensor-core candidates).
//   * CTA tile 128 output channels x 32 NJ frames (NJ = 1..4 picked per launch to fill whole waves), 256 threads,
//     8 x 2 NJ accumulators per thread (8 channels x NJ frame pairs, so that the 4-value activation window of a pair
//     feeds all three taps: 14 shared-memory loads per 192 FMAs at NJ = 4);
//   * K runs in chunks of 8 input channels (24 K-rows): the weight chunk [24][128] comes from the host-packed
//     [Cin*3][Cout_padded] layout with coalesced 16-byte loads; the activation chunk [8][32 NJ + 2] (one halo frame each
//     side) is loaded once, passed through Snake, and serves all three taps from shared memory;
//   * global loads of chunk c+1 are issued before the FMAs of chunk c (register prefetch, two smem buffers,
//     one barrier per chunk);
//   * zero padding at the sequence ends is applied after Snake, as the reference's conv does (snake(0) = 0).
#include <cstdlib>

#include "common.cuh"

namespace vrvq {

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);

namespace {

constexpr int SN_TM = 128;  // output channels per CTA (= VRVQ_CONV3_COUT_ALIGN)
constexpr int SN_KC = 8;    // input channels per chunk
// frames per CTA = 32 NJ (NJ frame pairs per thread), chosen per launch so that the grid fills whole waves (pick_nj)
static_assert(SN_TM == VRVQ_CONV3_COUT_ALIGN, "packed weight rows are padded to the CTA tile");

__device__ __forceinline__ float snake_f32(float v, float a, float inv_a) {
    const float s = sinf(a * v);
    return v + inv_a * (s * s);
}

template <bool SIGMOID, int NJ>
__global__ void __launch_bounds__(256, 2)
snake_conv3_kernel(const float *__restrict__ x, long long x_sb, long long x_sc, const float *__restrict__ alpha,
                   const float *__restrict__ wp, int cout_pad, const float *__restrict__ bias, float *__restrict__ y, long long y_sb,
                   long long y_sc, int Cin, int Cout, int T) {
    __shared__ __align__(16) float Ws[2][SN_KC * 3][SN_TM];
    constexpr int SN_TN = 32 * NJ, SN_XW = SN_TN + 2, SN_XLD = SN_TN + 8;
    __shared__ __align__(16) float Xs[2][SN_KC][SN_XLD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int t0 = blockIdx.x * SN_TN;
    const int co0 = blockIdx.y * SN_TM;
    const int b = blockIdx.z;
    const float *xb = x + (long long)b * x_sb;

    // loader roles: activations -- channel lci, columns lane + 32 m (m < NJ) and 32 NJ + lane (lane < 2);
    //               weights     -- three float4 at flat index tid + 256 m of the [24][128] chunk
    const int lci = tid >> 5, lane = tid & 31;
    float xr[NJ + 1];
    float4 wr[3];

    // global -> registers (raw values: nothing here waits on the loads, so they fly under the FMAs of the current chunk)
    float sn_a = 0.0f, sn_inv = 0.0f;
    auto load_chunk = [&](int c) {
        const int ci = c * SN_KC + lci;
        sn_a = __ldg(alpha + ci);
        const float *xrow = xb + (long long)ci * x_sc;
#pragma unroll
        for (int m = 0; m < NJ + 1; ++m) {
            const int i = lane + 32 * m;
            const int t = t0 - 1 + i;
            xr[m] = (i < SN_XW && t >= 0 && t < T) ? __ldg(xrow + t) : 0.0f;
        }
        const float *wrow = wp + (long long)c * (SN_KC * 3) * cout_pad + co0;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int idx = tid + 256 * m;
            wr[m] = __ldg(reinterpret_cast<const float4 *>(wrow + (long long)(idx >> 5) * cout_pad) + (idx & 31));
        }
    };
    // registers -> shared memory, Snake applied on the way (snake(0) = 0 keeps the zero padding)
    auto store_chunk = [&](int buf) {
        sn_inv = 1.0f / (sn_a + 1e-9f);
#pragma unroll
        for (int m = 0; m < NJ + 1; ++m) {
            const int i = lane + 32 * m;
            if (i < SN_XW) Xs[buf][lci][i] = snake_f32(xr[m], sn_a, sn_inv);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const int idx = tid + 256 * m;
            *reinterpret_cast<float4 *>(&Ws[buf][idx >> 5][(idx & 31) * 4]) = wr[m];
        }
    };

    float acc[8][2 * NJ];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2 * NJ; ++j) acc[i][j] = 0.0f;

    const int nchunks = Cin / SN_KC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) load_chunk(c + 1);
#pragma unroll
        for (int ci = 0; ci < SN_KC; ++ci) {
            // frames 32 j + 2 tx + {0,1}: the four window values per j serve all three taps (two 8-byte loads)
            float xv[NJ][4];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float2 lo = *reinterpret_cast<const float2 *>(&Xs[buf][ci][32 * j + 2 * tx]);
                const float2 hi = *reinterpret_cast<const float2 *>(&Xs[buf][ci][32 * j + 2 * tx + 2]);
                xv[j][0] = lo.x, xv[j][1] = lo.y, xv[j][2] = hi.x, xv[j][3] = hi.y;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 a0 = *reinterpret_cast<const float4 *>(&Ws[buf][ci * 3 + k][ty * 8]);
                const float4 a1 = *reinterpret_cast<const float4 *>(&Ws[buf][ci * 3 + k][ty * 8 + 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        acc[i][2 * j] = fmaf(a[i], xv[j][k], acc[i][2 * j]);
                        acc[i][2 * j + 1] = fmaf(a[i], xv[j][k + 1], acc[i][2 * j + 1]);
                    }
            }
        }
        if (c + 1 < nchunks) store_chunk(buf ^ 1);
        __syncthreads();
    }

    float *yb = y + (long long)b * y_sb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int co = co0 + ty * 8 + i;
        if (co >= Cout) break;
        const float bi = __ldg(bias + co);
        float *yrow = yb + (long long)co * y_sc;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int t = t0 + 32 * j + 2 * tx + e;
                if (t < T) {
                    float v = acc[i][2 * j + e] + bi;
                    if (SIGMOID) v = 1.0f / (1.0f + expf(-v));
                    yrow[t] = v;
                }
            }
        }
    }
}

// Frames per CTA (32 nj), picked per launch from a cost model fitted to the config-2 sweep (profiles/r1q_subnet_tile_sweep.txt):
//   time ~ f(CTAs on the busiest SM) x (nj + 1/2)
// nj + 1/2: FMAs scale with nj, the weight loads of a tile do not; f(n) = 1.6 (n / 2) + (n % 2): two CTAs are resident per SM
// and share its FMA pipes (a pair takes 1.6x a lone CTA).  128-frame tiles at config 2 (B=16, T=862, 1024 output channels) are
// 896 CTAs -> 7 on the busiest SM; 96-frame tiles are 1152 -> 8, each 3/4 the length.  Ties go to the larger tile.  Results do
// not depend on the choice: every output sums its terms in the same order.
int pick_nj(int T, long long ctas_per_frame_tile, long long sms) {
    int best = 4;
    long long best_cost = -1;
    for (int nj = 4; nj >= 1; --nj) {
        const long long tiles = (T + 32 * nj - 1) / (32 * nj);
        const long long per_sm = (tiles * ctas_per_frame_tile + sms - 1) / sms;
        const long long cost = (16 * (per_sm / 2) + 10 * (per_sm % 2)) * (2 * nj + 1);
        if (best_cost < 0 || cost < best_cost) best = nj, best_cost = cost;
    }
    return best;
}

}  // namespace

int launch_snake_conv3(const float *x, long long x_sb, long long x_sc, const float *alpha, const float *wp, int cout_pad, const float *bias,
                       int B, int Cin, int Cout, int T, int sigmoid, float *y, long long y_sb, long long y_sc, cudaStream_t st) {
    if ((long long)B * T == 0) return VRVQ_OK;
    if (B > 65535) {
        set_error("vrvq_snake_conv3_f32: B must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int co_tiles = (Cout + SN_TM - 1) / SN_TM;
    int nj = pick_nj(T, (long long)co_tiles * B, sms);
    if (const char *e = getenv("VRVQ_SUBNET_NJ")) {  // tests force every instantiation; results are identical by construction
        const int v = atoi(e);
        if (v >= 1 && v <= 4) nj = v;
    }
    dim3 grid((T + 32 * nj - 1) / (32 * nj), co_tiles, B);
#define VRVQ_SN_LAUNCH(SG, NJ_) \
    snake_conv3_kernel<SG, NJ_><<<grid, 256, 0, st>>>(x, x_sb, x_sc, alpha, wp, cout_pad, bias, y, y_sb, y_sc, Cin, Cout, T)
#define VRVQ_SN_CASE(NJ_)               \
    case NJ_:                           \
        if (sigmoid)                    \
            VRVQ_SN_LAUNCH(true, NJ_);  \
        else                            \
            VRVQ_SN_LAUNCH(false, NJ_); \
        break;
    switch (nj) {
        VRVQ_SN_CASE(1)
        VRVQ_SN_CASE(2)
        VRVQ_SN_CASE(3)
        default:
            VRVQ_SN_CASE(4)
    }
#undef VRVQ_SN_CASE
#undef VRVQ_SN_LAUNCH
    return check_cuda(cudaGetLastError(), "snake_conv3_kernel launch");
}

}  // namespace vrvq
