// Importance subnet, the two wide blocks (1024 -> 1024 and 1024 -> 512: 95 % of the 9.85 MFLOP per frame of
// models/importance_subnet.py:18-45) on the 5th-generation tensor cores.
//
//   y[b, co, t] = bias[co] + sum_{tap < 3} sum_{ci} W[co, ci, tap] * snake(x[b, ci, t + tap - 1]),   snake(v) = v + sin(alpha v)^2 / (alpha + 1e-9)
//
// as an implicit GEMM per batch item with the FRAMES as M (tile row = TMEM lane = frame), the output channels as N and
// K = 3 Cin ordered (32-channel chunk, tap, channel): the three taps of a chunk are the same staged activation rows read one frame
// apart.  fp32 in, fp32-grade out: 3xTF32 (hi*hi + hi*lo + lo*hi), the accumulator drained into running fp32 sums every 96 k
// (the tensor core rounds its accumulator toward zero on every k-step; profiles/r1_micro_tc3x.txt) -- cuDNN's default TF32
// convolution moves imp_map by 1.9e-3, i.e. mask edges.
//
// One CTA of 16 warps per SM walks tiles of 128 frames x 128 output channels (all Cin):
//   warp 14 (lane 0)   TMA: the activation chunk [32 channels][136 frames from t0 - 4] through the channel-class tensor maps
//                      (tmaps.cuh: any row pitch, e.g. T = 862) into a 3-slot ring
//   warps 0-7          Snake in place in shared memory (once per element; frames outside [0, T) become the conv's zero padding),
//                      then per tap: lane = frame reads its 16 channels one frame further right, splits them into TF32 head +
//                      exact remainder and writes them into tensor memory as the A operand (tcgen05.st, 2-slot ring)
//   warp 13 (lane 0)   cp.async.bulk of the weight chunk [hi 128 x 32 | lo 128 x 32] (32 KB, canonical K-major UMMA layout, packed by
//                      vrvq_pack_conv3_tc_weights) into a 4-slot ring
//   warp 12 (lane 0)   tcgen05.mma kind::tf32 M = 128 x N = 128 x K = 8, A in TMEM: 12 per chunk-tap (4 k-steps x 3 products) into one of
//                      two accumulator sets; after the 3 taps of a chunk the set goes to the drain
//   warps 8-11         drain: running sums (128 TMEM columns) += accumulator set; after the last one: + bias, store (lane = frame:
//                      every warp-level store writes 128 contiguous bytes of one channel row)
// Rings run across tiles (positions are running totals); the only cross-role ordering per tile is through the mbarriers.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tmaps.cuh"

namespace vrvq {

namespace {

constexpr int ST_NTH = 512;
constexpr int ST_XW = 136;                      // staged frames per row: t0 - 4 .. t0 + 131 (needs t0 - 1 .. t0 + 128, + <= 3 of class shift)
constexpr int ST_XSLOT = 32 * ST_XW * 4, ST_XSLOTS = 3;
constexpr int ST_WSLOT = 32768, ST_WSLOTS = 4;  // [hi | lo] x [8 kg][128 rows][4]
constexpr int ST_SM_W = 0, ST_SM_X = ST_WSLOTS * ST_WSLOT, ST_SM_BAR = ST_SM_X + ST_XSLOTS * ST_XSLOT, ST_SM_TMEM = ST_SM_BAR + 256;
constexpr int ST_SMEM = ST_SM_TMEM + 16;
static_assert(ST_SMEM <= 232448 && ST_XSLOT % 128 == 0, "shared memory map");
enum { SB_X_FULL = 0, SB_X_EMPTY = 3, SB_A_FULL = 6, SB_A_EMPTY = 8, SB_W_FULL = 10, SB_W_EMPTY = 14, SB_SET_FULL = 18, SB_SET_EMPTY = 20, SB_COUNT = 22 };
// tensor memory: running sums | two accumulator sets | A ring (2 slots x (32 heads | 32 remainders))
constexpr uint32_t ST_TM_RUN = 0, ST_TM_ACC = 128, ST_TM_A = 384;
constexpr int ST_DRAIN = 3;  // chunk-taps per accumulator set (one 32-channel chunk x 3 taps = 96 k: 36 accumulations between drains)

struct SnTcParams {
    const float *alpha, *wtc, *bias;  // alpha == NULL: x is already Snake-activated (a previous launch stored it that way)
    const float *post_alpha;          // != NULL: store snake(y[co], post_alpha[co]) -- the NEXT block's activation, fused into this epilogue
    float *y;
    long long y_sb, y_sc;
    int B, Cin, Cout, T;
    int n_ft, n_ct, n_tiles;  // frame tiles per item, output-channel tiles, tiles in total (tile = (b * n_ft + ft) * n_ct + ct)
    int nc_log2, shift[4];    // channel classes of the activation tensor maps
    int debug;                // profiling knob VRVQ_SUBNET_DEBUG: 1 = no Snake, 2 = no MMAs, 4 = no split / TMEM stores, 8 = no drains, 16 = no weight loads, 32 = no activation loads
};

constexpr unsigned long long ST_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __noinline__ unsigned long long st_wait_check(unsigned long long t0, int bar, uint32_t parity) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (t0 == 0) return now;
    if (now - t0 > ST_TIMEOUT_NS) {
        printf("[vrvq subnet tc] mbarrier %d wait timed out (parity %u, block %d, thread %d)\n", bar, parity, (int)blockIdx.x, (int)threadIdx.x);
        __trap();
    }
    return t0;
}
#define ST_WAIT(idx_, parity_)                                                                    \
    do {                                                                                          \
        const int i__ = (idx_);                                                                   \
        const uint32_t p__ = (parity_);                                                           \
        uint32_t spins__ = 0;                                                                     \
        unsigned long long t0__ = 0;                                                              \
        while (!mbar_try_wait(&bars[i__], p__)) {                                                 \
            if ((++spins__ & 0x3fffu) == 0) t0__ = st_wait_check(t0__, i__, p__);                 \
        }                                                                                         \
    } while (0)

constexpr uint64_t ST_DESC = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(2048 >> 4) << 16);  // 128-row K-major tile

__device__ __forceinline__ float snake_tc(float v, float a, float inv_a) {
    const float s = sinf(a * v);
    return v + inv_a * (s * s);
}

__global__ void __launch_bounds__(ST_NTH, 1) snake_conv3_tc_kernel(const SnTcParams P, const __grid_constant__ ZMaps xmaps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ST_SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + ST_SM_TMEM);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NCC = P.Cin / 32;     // 32-channel chunks
    const int NCT = 3 * NCC;        // chunk-taps (A chunks) per tile
    const int NGRP = NCT / ST_DRAIN;  // accumulator sets per tile (Cin % 32 == 0: checked on the host)
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (n_my <= 0) return;

    if (tid == 0) {
        for (int i = 0; i < ST_XSLOTS; ++i) { mbar_init(&bars[SB_X_FULL + i], 1); mbar_init(&bars[SB_X_EMPTY + i], 8); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[SB_A_FULL + i], 8); mbar_init(&bars[SB_A_EMPTY + i], 1);
            mbar_init(&bars[SB_SET_FULL + i], 1); mbar_init(&bars[SB_SET_EMPTY + i], 4);
        }
        for (int i = 0; i < ST_WSLOTS; ++i) { mbar_init(&bars[SB_W_FULL + i], 1); mbar_init(&bars[SB_W_EMPTY + i], 1); }
        fence_mbar_init();
    }
    if (w == 12) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t smem_base = smem_u32(smem);
    const int lg = P.nc_log2, nc = 1 << lg;

    for (int it = 0; it < n_my; ++it) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int ct = tile % P.n_ct, ft = (tile / P.n_ct) % P.n_ft, b = tile / (P.n_ct * P.n_ft);
        const int t0 = ft * 128, co0 = ct * 128;
        const uint32_t xbase = (uint32_t)it * (uint32_t)NCC, nbase = (uint32_t)it * (uint32_t)NCT, gbase = (uint32_t)it * (uint32_t)NGRP;

        if (w < 8) {
            // ================= loaders: Snake in place, then the three taps of the chunk into the A ring =================
            const int f = tid & 127, q = tid >> 7;
            const uint32_t tq = tmem + ((uint32_t)(32 * (w & 3)) << 16);
            const int prow = tid >> 3, pcol0 = (tid & 7) * 17;  // Snake pass: row of the slot, 17 of its 136 columns
            // the slot stores channel class k = channel % nc class-major: slot row = k * (32 / nc) + (channel within chunk) / nc
            const int pk = prow >> (5 - lg), pch = ((prow & ((32 >> lg) - 1)) << lg) + pk;  // class and channel (within the chunk) of the Snake row
            const int psh = P.shift[pk];
            for (int cc = 0; cc < NCC; ++cc) {
                const uint32_t xn = xbase + (uint32_t)cc, xs = xn % ST_XSLOTS;
                ST_WAIT(SB_X_FULL + xs, (xn / ST_XSLOTS) & 1u);
                float *slot = reinterpret_cast<float *>(smem + ST_SM_X + xs * ST_XSLOT);
                const bool preact = P.alpha == nullptr;
                if (!preact) {
                    const float a = __ldg(P.alpha + 32 * cc + pch), inv_a = 1.0f / (a + 1e-9f);
                    float *rowp = slot + prow * ST_XW;
#pragma unroll
                    for (int i = 0; i < 17; ++i) {
                        const int col = pcol0 + i, fr = t0 - 4 + col - psh;  // frame held by this column of the row
                        const float v = rowp[col];
                        rowp[col] = (fr >= 0 && fr < P.T) ? ((P.debug & 1) ? v : snake_tc(v, a, inv_a)) : 0.0f;  // (zero padding of the convolution)
                    }
                }
                if (!preact) named_bar_sync(2, 256);
                // columns of frame t0 + f + tap - 1: f + tap + 3 (+ the class shift of the row)
                for (int tap = 0; tap < 3; ++tap) {
                    const bool tapok = (t0 + f + tap - 1) >= 0 && (t0 + f + tap - 1) < P.T;
                    float h[16], l[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (P.debug & 4) { h[i] = 0.f; l[i] = 0.f; continue; }
                        const int ch = 16 * q + i, k = ch & (nc - 1), row = (k << (5 - lg)) + (ch >> lg);
                        float x = slot[row * ST_XW + f + tap + 3 + P.shift[k]];
                        if (preact && !tapok) x = 0.0f;  // zero padding (positions next to a row's ends hold the neighbouring rows' elements)
                        h[i] = tf32_hi(x);  // round-to-nearest head: the remainder (and with it the dropped lo*lo term) is half as large
                        l[i] = __fsub_rn(x, h[i]);
                    }
                    const uint32_t n = nbase + (uint32_t)(3 * cc + tap), sl = n & 1u, use = n >> 1;
                    if (use >= 1) {
                        ST_WAIT(SB_A_EMPTY + sl, (use - 1) & 1u);
                        tmem_fence_after_sync();
                    }
                    {
                        uint32_t hv[16], lv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { hv[i] = __float_as_uint(h[i]); lv[i] = __float_as_uint(l[i]); }
                        const uint32_t ta = tq + ST_TM_A + 64u * sl + 16u * (uint32_t)q;
                        tmem_st16(ta, hv);
                        tmem_st16(ta + 32, lv);
                    }
                    tmem_wait_st();
                    tmem_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[SB_A_FULL + sl]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SB_X_EMPTY + xs]);  // every lane of the warp has read its three taps
            }
        } else if (w < 12) {
            // ================= drains and the final store =================
            const int q4 = w - 8, r = 32 * q4 + lane;
            const uint32_t tq = tmem + ((uint32_t)(32 * q4) << 16);
            for (int g = 0; g < NGRP; ++g) {
                const uint32_t gg = gbase + (uint32_t)g, set = gg & 1u;
                ST_WAIT(SB_SET_FULL + set, (gg >> 1) & 1u);
                tmem_fence_after_sync();
                const uint32_t ts = tq + ST_TM_ACC + 128u * set;
#pragma unroll 2
                for (int c8 = 0; c8 < ((P.debug & 8) ? 0 : 16); ++c8) {
                    uint32_t acc[8], run[8];
                    tmem_ld8(ts + 8 * c8, acc);
                    if (g > 0) tmem_ld8(tq + ST_TM_RUN + 8 * c8, run);
                    tmem_wait_ld(acc);
                    if (g > 0) {
                        tmem_wait_ld(run);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = __float_as_uint(__fadd_rn(__uint_as_float(run[i]), __uint_as_float(acc[i])));
                    }
                    tmem_st8(tq + ST_TM_RUN + 8 * c8, acc);
                }
                tmem_wait_st();
                tmem_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SB_SET_EMPTY + set]);
            }
            // y[b][co0 + c][t0 + r] = running sum + bias
            const int t = t0 + r;
            float *yb = P.y + (long long)b * P.y_sb + t;
            for (int c8 = 0; c8 < 16; ++c8) {
                uint32_t v[8];
                tmem_ld8(tq + ST_TM_RUN + 8 * c8, v);
                tmem_wait_ld(v);
                if (t < P.T) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int co = co0 + 8 * c8 + i;
                        float o = __fadd_rn(__uint_as_float(v[i]), __ldg(P.bias + co));
                        if (P.post_alpha != nullptr) {
                            const float a = __ldg(P.post_alpha + co);
                            o = snake_tc(o, a, 1.0f / (a + 1e-9f));
                        }
                        yb[(long long)co * P.y_sc] = o;
                    }
                }
            }
            tmem_fence_before_sync();
        } else if (w == 12) {
            // ================= MMA issuer =================
            if (lane == 0) {
                constexpr uint32_t ID_128 = umma_idesc_tf32(128, 128);
                for (int c = 0; c < NCT; ++c) {
                    const uint32_t n = nbase + (uint32_t)c, sl = n & 1u, wsl = n % ST_WSLOTS;
                    const uint32_t gg = gbase + (uint32_t)(c / ST_DRAIN), set = gg & 1u;
                    const bool first = (c % ST_DRAIN) == 0;
                    ST_WAIT(SB_A_FULL + sl, (n >> 1) & 1u);
                    ST_WAIT(SB_W_FULL + wsl, (n / ST_WSLOTS) & 1u);
                    if (first && gg >= 2) ST_WAIT(SB_SET_EMPTY + set, ((gg >> 1) - 1) & 1u);
                    tmem_fence_after_sync();
                    const uint32_t a_hi = tmem + ST_TM_A + 64u * sl, a_lo = a_hi + 32;
                    const uint64_t bh = ST_DESC | (uint64_t)((smem_base + ST_SM_W + wsl * ST_WSLOT) >> 4), bl = bh + (16384 >> 4);
                    const uint32_t d = tmem + ST_TM_ACC + 128u * set;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (P.debug & 2) break;  // (profiling knob: no MMAs)
                        umma_tf32_ts(d, a_lo + 8 * ks, bh + ks * (4096 >> 4), ID_128, !(first && ks == 0));  // smallest terms first
                        umma_tf32_ts(d, a_hi + 8 * ks, bl + ks * (4096 >> 4), ID_128, true);
                        umma_tf32_ts(d, a_hi + 8 * ks, bh + ks * (4096 >> 4), ID_128, true);
                    }
                    umma_commit(&bars[SB_A_EMPTY + sl]);
                    umma_commit(&bars[SB_W_EMPTY + wsl]);
                    if ((c % ST_DRAIN) == ST_DRAIN - 1) umma_commit(&bars[SB_SET_FULL + set]);
                }
            }
            __syncwarp();
        } else if (w == 13) {
            // ================= weight producer =================
            if (lane == 0) {
                const float *wt = P.wtc + (size_t)ct * (size_t)NCT * 8192;
                for (int c = 0; c < NCT; ++c) {
                    const uint32_t m = nbase + (uint32_t)c, slot = m % ST_WSLOTS, use = m / ST_WSLOTS;
                    if (use >= 1) ST_WAIT(SB_W_EMPTY + slot, (use - 1) & 1u);
                    if (P.debug & 16) { mbar_arrive(&bars[SB_W_FULL + slot]); continue; }
                    mbar_arrive_expect_tx(&bars[SB_W_FULL + slot], ST_WSLOT);
                    bulk_g2s(smem + ST_SM_W + slot * ST_WSLOT, wt + (size_t)c * 8192, ST_WSLOT, &bars[SB_W_FULL + slot]);
                }
            }
            __syncwarp();
        } else if (w == 14) {
            // ================= activation producer (TMA) =================
            if (lane == 0) {
                const int rows = 32 >> lg;
                for (int cc = 0; cc < NCC; ++cc) {
                    const uint32_t m = xbase + (uint32_t)cc, slot = m % ST_XSLOTS, use = m / ST_XSLOTS;
                    if (use >= 1) ST_WAIT(SB_X_EMPTY + slot, (use - 1) & 1u);
                    unsigned char *dst = smem + ST_SM_X + slot * ST_XSLOT;
                    if (P.debug & 32) { mbar_arrive(&bars[SB_X_FULL + slot]); continue; }
                    mbar_arrive_expect_tx(&bars[SB_X_FULL + slot], ST_XSLOT);
                    for (int k = 0; k < nc; ++k) tma_load_3d(dst + k * rows * (ST_XW * 4), &xmaps.m[k], t0 - 4, rows * cc, b, &bars[SB_X_FULL + slot]);
                }
            }
            __syncwarp();
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    if (w == 12) tmem_dealloc(tmem, 512);
}

// y[b][c][t] = snake(x[b][c][t], alpha[c]): the activation of the FIRST tensor-core block's input, computed once instead of once per
// 128-channel output tile (8 times for 1024 -> 1024); the later blocks get theirs from the previous epilogue (post_alpha)
__global__ void snake_kernel(const float *__restrict__ x, long long x_sb, long long x_sc, const float *__restrict__ alpha, float *__restrict__ y,
                             long long y_sb, long long y_sc, int C, int T) {
    const int c = blockIdx.y, b = blockIdx.z;
    const float a = __ldg(alpha + c), inv_a = 1.0f / (a + 1e-9f);
    const float *xr = x + (long long)b * x_sb + (long long)c * x_sc;
    float *yr = y + (long long)b * y_sb + (long long)c * y_sc;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) yr[t] = snake_tc(xr[t], a, inv_a);
}

}  // namespace

int launch_snake(const float *x, long long x_sb, long long x_sc, const float *alpha, int B, int C, int T, float *y, long long y_sb, long long y_sc,
                 cudaStream_t st) {
    if ((long long)B * C * T == 0) return VRVQ_OK;
    if (B > 65535 || C > 65535) {
        set_error("vrvq_snake_f32: B and C must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    snake_kernel<<<dim3((unsigned)((T + 255) / 256 < 8 ? (T + 255) / 256 : 8), (unsigned)C, (unsigned)B), 256, 0, st>>>(x, x_sb, x_sc, alpha, y, y_sb, y_sc, C, T);
    return check_cuda(cudaGetLastError(), "snake_kernel launch");
}

// ---- host side -----------------------------------------------------------------------------------
// shapes the tensor-core block serves: whole 128-channel output tiles, whole 32-channel input chunks (a drain covers the 3 taps of one)
int snake_conv3_tc_usable(int Cin, int Cout) { return Cin >= 64 && Cin % 32 == 0 && Cout >= 128 && Cout % 128 == 0; }

size_t conv3_tc_packed_floats(int Cout, int Cin) {
    if (!snake_conv3_tc_usable(Cin, Cout)) return 0;
    return (size_t)(Cout / 128) * (size_t)(3 * Cin / 32) * 8192;
}

// w [Cout][Cin][3] (weight-norm folded) -> [Cout / 128][3 Cin / 32 chunk-taps][hi | lo][8 kg][128 rows][4]:
// chunk-tap n = 3 * (ci / 32) + tap, row = co % 128, k = ci % 32; head = TF32 round-to-nearest, remainder exact
int pack_conv3_tc_weights(int Cout, int Cin, const float *w, float *out) {
    const int nct = 3 * Cin / 32;
    for (int ct = 0; ct < Cout / 128; ++ct)
        for (int n = 0; n < nct; ++n) {
            float *hi = out + ((size_t)ct * nct + n) * 8192, *lo = hi + 4096;
            const int cc = n / 3, tap = n % 3;
            for (int r = 0; r < 128; ++r)
                for (int k = 0; k < 32; ++k) {
                    const float x = w[((size_t)(128 * ct + r) * Cin + 32 * cc + k) * 3 + tap];
                    uint32_t u;
                    memcpy(&u, &x, 4);
                    float h = x;
                    if ((u & 0x7f800000u) != 0x7f800000u) {
                        u = (u + 0x1000u) & 0xffffe000u;
                        memcpy(&h, &u, 4);
                    }
                    const size_t idx = ((size_t)(k / 4) * 128 + r) * 4 + (k % 4);
                    hi[idx] = h;
                    lo[idx] = x - h;
                }
        }
    return VRVQ_OK;
}

int launch_snake_conv3_tc(const float *x, long long x_sb, long long x_sc, const float *alpha, const float *wtc, const float *bias,
                          const float *post_alpha, int B, int Cin, int Cout, int T, float *y, long long y_sb, long long y_sc, cudaStream_t st) {
    if ((long long)B * T == 0) return VRVQ_OK;
    if (!snake_conv3_tc_usable(Cin, Cout)) {
        set_error("vrvq_snake_conv3_tc_f32: Cin must be a multiple of 64 and Cout a multiple of 128 (got %d -> %d)", Cin, Cout);
        return VRVQ_EUNSUPPORTED;
    }
    SnTcParams P{};
    ZMaps maps;
    if (!build_class_maps(x, B, Cin, T, x_sc, x_sb, ST_XW, 32, &maps, &P.nc_log2, P.shift)) {
        set_error("vrvq_snake_conv3_tc_f32: the activation layout allows no tensor map (row pitch %lld, item pitch %lld)", x_sc, x_sb);
        return VRVQ_EUNSUPPORTED;
    }
    P.alpha = alpha; P.wtc = wtc; P.bias = bias; P.post_alpha = post_alpha; P.y = y; P.y_sb = y_sb; P.y_sc = y_sc;
    P.B = B; P.Cin = Cin; P.Cout = Cout; P.T = T;
    P.debug = getenv("VRVQ_SUBNET_DEBUG") ? atoi(getenv("VRVQ_SUBNET_DEBUG")) : 0;
    P.n_ft = (T + 127) / 128; P.n_ct = Cout / 128;
    const long long tiles = (long long)B * P.n_ft * P.n_ct;
    if (tiles > 0x7fffffffLL) {
        set_error("vrvq_snake_conv3_tc_f32: too many tiles");
        return VRVQ_EUNSUPPORTED;
    }
    P.n_tiles = (int)tiles;
    const int sms = current_sm_count();
    if (sms <= 0) {
        set_error("cannot query the SM count of the current device");
        return VRVQ_ECUDA;
    }
    int rc = ensure_dynamic_smem<snake_conv3_tc_kernel>(ST_SMEM, "cudaFuncSetAttribute(snake_conv3_tc_kernel)");
    if (rc) return rc;
    const int grid = P.n_tiles < sms ? P.n_tiles : sms;
    snake_conv3_tc_kernel<<<grid, ST_NTH, ST_SMEM, st>>>(P, maps);
    return check_cuda(cudaGetLastError(), "snake_conv3_tc_kernel launch");
}

}  // namespace vrvq
