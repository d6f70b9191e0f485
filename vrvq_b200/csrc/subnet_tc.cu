// Importance subnet, the two wide blocks (1024 -> 1024 and 1024 -> 512: 95 % of the 9.85 MFLOP per frame of
// models/importance_subnet.py:18-45) on the 5th-generation tensor cores.
//
//   y[b, co, t] = bias[co] + sum_{tap < 3} sum_{ci} W[co, ci, tap] * snake(x[b, ci, t + tap - 1]),   snake(v) = v + sin(alpha v)^2 / (alpha + 1e-9)
//
// as an implicit GEMM per batch item with the FRAMES as M (tile row = TMEM lane = frame), the output channels as N and
// K = 3 Cin ordered (32-channel chunk, tap, channel): the three taps of a chunk are the same staged activation rows read one frame
// apart.  fp32 in, fp32-grade out: 3xTF32 (hi*hi + hi*lo + lo*hi), the accumulators drained into running fp32 sums after 24 accumulations
// (the tensor core rounds its accumulator toward zero on every k-step; profiles/r1_micro_tc3x.txt) -- cuDNN's default TF32
// convolution moves imp_map by 1.9e-3, i.e. mask edges.
//
// One CTA of 20 warps per SM walks tiles of 128 frames x 128 output channels (all Cin):
//   warp 19 (lane 0)   TMA: the activation chunk [32 channels][136 frames from t0 - 4] through the channel-class tensor maps
//                      (tmaps.cuh: any row pitch, e.g. T = 862) into a 3-slot ring
//   warps 0-7          (Snake in place in shared memory when the input is not activated yet,) then per tap: lane = frame reads its 16
//                      channels one frame further right, splits them into TF32 head + exact remainder and writes them into tensor
//                      memory as the A operand (tcgen05.st) -- a 4-slot ring: the hand-over chain loader -> issuer -> commit -> loader
//                      is ~2000 cycles long, an MMA batch 768, so with 2 slots the tensor pipe idled 60 % of the time
//                      (profiles/r2c_subnet_tc_v1_debug_sweep.txt)
//   warp 18 (lane 0)   cp.async.bulk of the weight chunk [hi 128 x 32 | lo 128 x 32] (32 KB, canonical K-major UMMA layout, packed by
//                      vrvq_pack_conv3_tc_weights) into a 4-slot ring
//   warps 16, 17       tcgen05.mma kind::tf32 M = 128 x N = 128 x K = 8, A in TMEM: 12 per chunk-tap (4 k-steps x 3 products).  TWO issuing
//                      warps, even / odd chunk-taps, each into its OWN accumulator set: issuing an MMA costs its thread ~46 cycles and
//                      an mbarrier wait ~200 (profiles/r1_micro_mma_rate.txt), so one thread needs ~1000 cycles per chunk-tap of 768
//                      tensor-pipe cycles (v1/v2 traces: profiles/r2*_subnet_tc_*_debug_sweep.txt); two keep the pipe fed.  The loops
//                      run on warp-uniform values (tcgen05.mma then issues straight from uniform registers; under `if (lane == 0)`
//                      every MMA sat in a vote / elect / R2UR loop).  A set is handed to the drain after 2 chunk-taps (24
//                      accumulations: the tensor core truncates its accumulator on every step), the two issuers' hand-overs alternate
//                      (after chunk-tap n = 2 mod 4 and n = 3 mod 4), so one keeps issuing while the other's set is drained
//   warps 8-15         drain: running sums in REGISTERS (warp = lane quadrant x column half: 64 sums per thread) += accumulator set;
//                      after the last one: + bias, store (lane = frame: every warp-level store writes 128 contiguous bytes of a row)
// Rings run across tiles (positions are running totals); the only cross-role ordering per tile is through the mbarriers.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tmaps.cuh"

namespace vrvq {

namespace {

constexpr int ST_NTH = 640;  // 8 loader warps, 8 drain warps, 2 issuers, weight producer, activation producer
constexpr int ST_XW = 136;                      // staged frames per row: t0 - 4 .. t0 + 131 (needs t0 - 1 .. t0 + 128, + <= 3 of class shift)
constexpr int ST_XSLOT = 32 * ST_XW * 4, ST_XSLOTS = 2;
constexpr int ST_WSLOT = 32768, ST_WSLOTS = 4;  // [hi | lo] x [8 kg][128 rows][4]
constexpr int ST_ASLOTS = 4;                    // A ring in tensor memory: 4 x (32 heads | 32 remainders)
constexpr int ST_SM_W = 0, ST_SM_X = ST_WSLOTS * ST_WSLOT, ST_SM_STAGE = ST_SM_X + ST_XSLOTS * ST_XSLOT;  // STAGE: the output tile [128 channels][128 frames]
constexpr int ST_SM_ALPHA = ST_SM_STAGE + 65536, ST_SM_BAR = ST_SM_ALPHA + 512, ST_SM_TMEM = ST_SM_BAR + 256;     // ALPHA: post_alpha of the tile's channels
constexpr int ST_SMEM = ST_SM_TMEM + 16;
static_assert(ST_SMEM <= 232448 && ST_XSLOT % 128 == 0, "shared memory map");
// FULL / EMPTY: one barrier pair per ring position for BOTH operands of a chunk-tap (A slot sl in tensor memory and weight slot sl in
// shared memory are filled by 4 loader warps + 1 bulk copy and released by one tcgen05.commit): a successful mbarrier wait costs the
// waiting thread ~170-250 cycles (profiles/r1_micro_mma_rate.txt), so the consumer side waits once per chunk-tap, not twice
enum { SB_X_FULL = 0, SB_X_EMPTY = 3, SB_FULL = 6, SB_EMPTY = 10, SB_SET_FULL = 14, SB_SET_EMPTY = 16, SB_COUNT = 18 };
static_assert(ST_ASLOTS == ST_WSLOTS, "A and weight rings share their barriers");
static_assert(SB_COUNT * 8 <= 256, "barrier block");
// tensor memory: two accumulator sets | A ring
constexpr uint32_t ST_TM_ACC = 0, ST_TM_A = 256;

struct SnTcParams {
    const float *alpha, *wtc, *bias;  // alpha == NULL: x is already Snake-activated (a previous launch stored it that way)
    const float *post_alpha;          // != NULL: store snake(y[co], post_alpha[co]) -- the NEXT block's activation, fused into this epilogue
    float *y;
    long long y_sb, y_sc;
    int B, Cin, Cout, T;
    int n_ft, n_ct, n_tiles;  // frame tiles per item, output-channel tiles, tiles in total (tile = (b * n_ft + ft) * n_ct + ct)
    int nc_log2, shift[4];    // channel classes of the activation tensor maps
    int y_tma;                // the output allows a tensor map (16-byte aligned rows): tiles leave through shared memory and one TMA store
    long long *trace;         // VRVQ_SUBNET_TRACE=1: block 0 writes per-role cycle counters (total, waits per barrier class)
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
// the same, with the wait's cycles added to a per-role trace counter when tracing
#define ST_WAIT_TR(idx_, parity_, ctr_)                       \
    do {                                                      \
        const unsigned c0__ = tr ? clock() : 0u;              \
        ST_WAIT(idx_, parity_);                               \
        if (tr) ctr_ += clock() - c0__;                       \
    } while (0)

constexpr uint64_t ST_DESC = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(2048 >> 4) << 16);  // 128-row K-major tile

// sin(t)^2 without the library's branchy sinf: sin^2 has period pi, so t is reduced to r = t - j pi in [-pi/2, pi/2] (two-term Cody-Waite
// with FMAs: exact to ~1e-7 for |t| up to 1e5) and sin(r) is a degree-13 odd polynomial (truncation error 7e-10 at pi/2).  Branch-free:
// the 64 activations of a drain thread interleave, where 64 sinf calls ran one after the other (35k cycles per tile).  Absolute error of
// sin(t)^2 <= 3e-7, the size of one fp32 rounding of the activation; NaN / Inf propagate as in sinf.
__device__ __forceinline__ float sin_sq(float t) {
    const float j = rintf(__fmul_rn(t, 0.318309886183790672f));
    float r = fmaf(j, -3.14159274101257324f, t);  // float(pi)
    r = fmaf(j, 8.74227765734758577e-8f, r);      // float(pi) - pi
    const float s2 = __fmul_rn(r, r);
    float p = fmaf(s2, 1.60590438368216146e-10f, -2.50521083854417188e-8f);
    p = fmaf(s2, p, 2.75573192239858907e-6f);
    p = fmaf(s2, p, -1.98412698412698413e-4f);
    p = fmaf(s2, p, 8.33333333333333333e-3f);
    p = fmaf(s2, p, -1.66666666666666667e-1f);
    const float s = fmaf(__fmul_rn(r, s2), p, r);
    return __fmul_rn(s, s);
}
// 1 / (alpha + 1e-9) as one MUFU.RCP (<= 1 ulp from the rounded quotient: 1e-7 relative on the sin^2 term): the IEEE division is a
// call-guarded sequence per element, which serialised the 64 channels of a drain thread
__device__ __forceinline__ float snake_inv(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(a, 1e-9f)));
    return r;
}
__device__ __forceinline__ float snake_tc(float v, float a, float inv_a) { return __fadd_rn(v, __fmul_rn(inv_a, sin_sq(__fmul_rn(a, v)))); }

template <int LG>
__global__ void __launch_bounds__(ST_NTH, 1) snake_conv3_tc_kernel(const SnTcParams P, const __grid_constant__ ZMaps xmaps, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ST_SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + ST_SM_TMEM);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NCC = P.Cin / 32;     // 32-channel chunks
    const int NCT = 3 * NCC;        // chunk-taps (A chunks) per tile
    // hand-overs per tile: set 0 after chunk-taps 2, 6, 10, .. and its last one, set 1 after 3, 7, 11, .. and its last one: they alternate, set 0 first
    const int NGRP = ((NCT + 1) / 2 + 1) / 2 + (NCT / 2 + 1) / 2;
    const int n_my = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (n_my <= 0) return;

    if (tid == 0) {
        for (int i = 0; i < ST_XSLOTS; ++i) { mbar_init(&bars[SB_X_FULL + i], 1); mbar_init(&bars[SB_X_EMPTY + i], 8); }
        for (int i = 0; i < ST_ASLOTS; ++i) { mbar_init(&bars[SB_FULL + i], 9); mbar_init(&bars[SB_EMPTY + i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars[SB_SET_FULL + i], 1); mbar_init(&bars[SB_SET_EMPTY + i], 8); }
        fence_mbar_init();
    }
    if (w == 16) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem);  // the same value, provably warp-uniform (issuer)
    uint32_t setcnt0 = 0, setcnt1 = 0;  // hand-overs so far (running totals) of accumulator set 0 / 1: issuers and drains count alike
    const uint32_t smem_base = smem_u32(smem);
    constexpr int lg = LG, nc = 1 << LG;
    const bool tr = P.trace != nullptr && blockIdx.x == 0;
    const long long tr_total = tr ? clock64() : 0;
    unsigned tr_a = 0, tr_b = 0, tr_c = 0, tr_d = 0;
    const int sh0 = P.shift[0], sh1 = P.shift[1], sh2 = P.shift[2], sh3 = P.shift[3];
    auto shift_of = [&](int k) { return k == 0 ? sh0 : k == 1 ? sh1 : k == 2 ? sh2 : sh3; };

    for (int it = 0; it < n_my; ++it) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        const int ct = tile % P.n_ct, ft = (tile / P.n_ct) % P.n_ft, b = tile / (P.n_ct * P.n_ft);
        const int t0 = ft * 128, co0 = ct * 128;
        const uint32_t xbase = (uint32_t)it * (uint32_t)NCC, nbase = (uint32_t)it * (uint32_t)NCT;

        if (w < 8) {
            // ================= loaders: (Snake in place,) then the three taps of the chunk into the A ring =================
            const int f = tid & 127, q = tid >> 7;  // frame row of the tile = TMEM lane; channel half of the chunk
            const uint32_t tq = tmem + ((uint32_t)(32 * (w & 3)) << 16);
            const bool preact = P.alpha == nullptr;
            for (int cc = 0; cc < NCC; ++cc) {
                const uint32_t xn = xbase + (uint32_t)cc, xs = xn % ST_XSLOTS;
                ST_WAIT_TR(SB_X_FULL + xs, (xn / ST_XSLOTS) & 1u, tr_a);
                float *slot = reinterpret_cast<float *>(smem + ST_SM_X + xs * ST_XSLOT);
                if (!preact) {
                    // the slot stores channel class k = channel % nc class-major: slot row = k * (32 / nc) + (channel within chunk) / nc
                    const int prow = tid >> 3, pcol0 = (tid & 7) * 17;  // Snake pass: row of the slot, 17 of its 136 columns
                    const int pk = prow >> (5 - lg), pch = ((prow & ((32 >> lg) - 1)) << lg) + pk;  // class and channel (within the chunk) of the row
                    const int psh = shift_of(pk);
                    const float a = __ldg(P.alpha + 32 * cc + pch), inv_a = snake_inv(a);
                    float *rowp = slot + prow * ST_XW;
                    for (int i = 0; i < 17; ++i) {
                        const int col = pcol0 + i, fr = t0 - 4 + col - psh;  // frame held by this column of the row
                        const float v = rowp[col];
                        rowp[col] = (fr >= 0 && fr < P.T) ? ((P.debug & 1) ? v : snake_tc(v, a, inv_a)) : 0.0f;  // (zero padding of the convolution)
                    }
                    named_bar_sync(2, 256);
                }
                // columns of frame t0 + f + tap - 1: f + tap + 3 (+ the class shift of the row).  The loads and the split of tap + 1 run
                // while the tensor-memory stores of tap complete (tcgen05.st reads its registers at issue; ~240 cycles to wait::st)
                uint32_t hv[16], lv[16];
                auto load_split = [&](int tap) {
                    const bool tapok = (t0 + f + tap - 1) >= 0 && (t0 + f + tap - 1) < P.T;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (P.debug & 4) { hv[i] = 0u; lv[i] = 0u; continue; }
                        const int k = i & (nc - 1);  // (16 * q + i) % nc: class of the channel; its slot row: k * (32 / nc) + (16 q + i) / nc
                        float x = slot[((k << (5 - lg)) + (i >> lg) + (q << (4 - lg))) * ST_XW + f + tap + 3 + shift_of(k)];
                        if (preact && !tapok) x = 0.0f;  // zero padding (positions next to a row's ends hold the neighbouring rows' elements)
                        // TF32 head, round to nearest (ties away, like cvt.rna.tf32.f32 -- which issues at a quarter of the integer rate): the
                        // remainder (and with it the dropped lo*lo term) is half as large as after truncation; x - head is exact
                        hv[i] = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
                        lv[i] = __float_as_uint(__fsub_rn(x, __uint_as_float(hv[i])));
                    }
                };
                load_split(0);
#pragma unroll
                for (int tap = 0; tap < 3; ++tap) {
                    const uint32_t n = nbase + (uint32_t)(3 * cc + tap), sl = n % ST_ASLOTS, use = n / ST_ASLOTS;
                    const uint32_t ta = tq + ST_TM_A + 64u * sl + 16u * (uint32_t)q;
                    if (use >= 1) {
                        ST_WAIT_TR(SB_EMPTY + sl, (use - 1) & 1u, tr_b);
                        tmem_fence_after_sync();
                    }
                    const unsigned c_tap = tr ? clock() : 0u;
                    tmem_st16(ta, hv);
                    tmem_st16(ta + 32, lv);
                    if (tap < 2) load_split(tap + 1);
                    tmem_wait_st();
                    tmem_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[SB_FULL + sl]);
                    if (tr) tr_c += clock() - c_tap;  // TMEM stores, next tap's loads and split, arrive
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SB_X_EMPTY + xs]);  // every lane of the warp has read its three taps
            }
        } else if (w < 16) {
            // ================= drains and the final store =================
            const int q4 = w & 3, half = (w - 8) >> 2, r = 32 * q4 + lane;
            const uint32_t tq = tmem + ((uint32_t)(32 * q4) << 16) + 64u * (uint32_t)half;
            float run[64];  // starts from the bias: nothing but the (optional) Snake and the stores are left after the last drain
#pragma unroll
            for (int i = 0; i < 64; ++i) run[i] = __ldg(P.bias + co0 + 64 * half + i);
            if (P.post_alpha != nullptr) {  // this tile's Snake parameters: read back as broadcasts after the last drain
                named_bar_sync(4, 256);     // (the previous tile's epilogue has read its own)
                if (tid < 256 + 128) reinterpret_cast<float *>(smem + ST_SM_ALPHA)[tid - 256] = __ldg(P.post_alpha + co0 + tid - 256);
                named_bar_sync(4, 256);
            }
            for (int g = 0; g < NGRP; ++g) {
                const uint32_t set = (uint32_t)(g & 1);  // the hand-overs of a tile alternate, set 0 first
                const uint32_t k = set ? setcnt1 : setcnt0;
                if (set) ++setcnt1; else ++setcnt0;
                ST_WAIT_TR(SB_SET_FULL + set, k & 1u, tr_a);
                tmem_fence_after_sync();
                const uint32_t ts = tq + ST_TM_ACC + 128u * set;
                if (!(P.debug & 8)) {
#pragma unroll
                    for (int c16 = 0; c16 < 4; ++c16) {  // (16 at a time: 640 threads leave 96 registers per thread, 64 of them hold the sums)
                        uint32_t acc[16];
                        tmem_ld16(ts + 16 * c16, acc);
                        tmem_wait_ld16(acc);
#pragma unroll
                        for (int i = 0; i < 16; ++i) run[16 * c16 + i] = __fadd_rn(run[16 * c16 + i], __uint_as_float(acc[i]));
                    }
                }
                tmem_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[SB_SET_EMPTY + set]);
            }
            // y[b][co0 + c][t0 + r] = running sum (bias included), through the next block's Snake when asked for
            const int t = t0 + r;
            const unsigned c_fin = tr ? clock() : 0u;
            const float *al = reinterpret_cast<const float *>(smem + ST_SM_ALPHA) + 64 * half;  // (written before the first drain of the tile)
            if (P.y_tma) {
                // through shared memory and ONE tensor store (frames >= T are clipped by the map): per-lane 4-byte stores kept the drain warps
                // ~13k cycles per tile in the store queue, with both accumulator sets waiting for them
                float *stage = reinterpret_cast<float *>(smem + ST_SM_STAGE) + (64 * half) * 128 + r;
                if (tid == 256) bulk_wait_group_read0();  // the previous tile's store has read the staging tile
                named_bar_sync(3, 256);
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    float o = run[i];
                    if (P.post_alpha != nullptr) {
                        const float a = al[i];
                        o = snake_tc(o, a, snake_inv(a));
                    }
                    stage[i * 128] = o;
                }
                fence_proxy_async();
                named_bar_sync(3, 256);
                if (tid == 256) {
                    tma_store_3d(&ymap, smem + ST_SM_STAGE, t0, co0, b);
                    bulk_commit_group();
                }
            } else if (t < P.T) {
                float *yb = P.y + (long long)b * P.y_sb + t;
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const int co = co0 + 64 * half + i;
                    float o = run[i];
                    if (P.post_alpha != nullptr) {
                        const float a = al[i];
                        o = snake_tc(o, a, snake_inv(a));
                    }
                    yb[(long long)co * P.y_sc] = o;
                }
            }
            if (tr) tr_b += clock() - c_fin;
        } else if (w < 18) {
            // ================= MMA issuers: X = 0 the even chunk-taps into set 0, X = 1 the odd ones into set 1 =================
            constexpr uint32_t ID_128 = umma_idesc_tf32(128, 128);
            const uint32_t X = (uint32_t)(w - 16);
            bool fresh = true;  // the set holds nothing yet: the first MMA overwrites, and the previous hand-over must have been drained
            for (int c = (int)X; c < NCT; c += 2) {
                const uint32_t n = nbase + (uint32_t)c, sl = n % ST_ASLOTS;
                ST_WAIT_TR(SB_FULL + sl, (n / ST_ASLOTS) & 1u, tr_a);
                uint32_t &setcnt = X ? setcnt1 : setcnt0;
                if (fresh && setcnt >= 1) ST_WAIT_TR(SB_SET_EMPTY + X, (setcnt - 1) & 1u, tr_c);
                tmem_fence_after_sync();
                const unsigned c_iss = tr ? clock() : 0u;
                const uint32_t a_hi = tmem_u + ST_TM_A + 64u * sl, a_lo = a_hi + 32;
                const uint64_t bh = ST_DESC | (uint64_t)((smem_base + ST_SM_W + sl * ST_WSLOT) >> 4), bl = bh + (16384 >> 4);
                const uint32_t d = tmem_u + ST_TM_ACC + 128u * X;
                const bool handover = ((c - (int)X) % 4) == 2 || c + 2 >= NCT;  // after two chunk-taps of this issuer (24 accumulations), and after its last
                if (elect_one()) {
                    if (!(P.debug & 2)) {  // (profiling knob: no MMAs)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            umma_tf32_ts(d, a_lo + 8 * ks, bh + ks * (4096 >> 4), ID_128, !(fresh && ks == 0));  // smallest terms first
                            umma_tf32_ts(d, a_hi + 8 * ks, bl + ks * (4096 >> 4), ID_128, true);
                            umma_tf32_ts(d, a_hi + 8 * ks, bh + ks * (4096 >> 4), ID_128, true);
                        }
                    }
                    umma_commit(&bars[SB_EMPTY + sl]);
                    if (handover) umma_commit(&bars[SB_SET_FULL + X]);
                }
                fresh = handover;
                if (handover) ++setcnt;
                if (tr) tr_d += clock() - c_iss;
            }
        } else if (w == 18) {
            // ================= weight producer =================
            if (lane == 0) {
                const float *wt = P.wtc + (size_t)ct * (size_t)NCT * 8192;
                for (int c = 0; c < NCT; ++c) {
                    const uint32_t m = nbase + (uint32_t)c, slot = m % ST_WSLOTS, use = m / ST_WSLOTS;
                    if (use >= 1) ST_WAIT_TR(SB_EMPTY + slot, (use - 1) & 1u, tr_a);
                    if (P.debug & 16) { mbar_arrive(&bars[SB_FULL + slot]); continue; }
                    mbar_arrive_expect_tx(&bars[SB_FULL + slot], ST_WSLOT);
                    bulk_g2s(smem + ST_SM_W + slot * ST_WSLOT, wt + (size_t)c * 8192, ST_WSLOT, &bars[SB_FULL + slot]);
                }
            }
            __syncwarp();
        } else if (w == 19) {
            // ================= activation producer (TMA) =================
            if (lane == 0) {
                const int rows = 32 >> lg;
                for (int cc = 0; cc < NCC; ++cc) {
                    const uint32_t m = xbase + (uint32_t)cc, slot = m % ST_XSLOTS, use = m / ST_XSLOTS;
                    if (use >= 1) ST_WAIT_TR(SB_X_EMPTY + slot, (use - 1) & 1u, tr_a);
                    unsigned char *dst = smem + ST_SM_X + slot * ST_XSLOT;
                    if (P.debug & 32) { mbar_arrive(&bars[SB_X_FULL + slot]); continue; }
                    mbar_arrive_expect_tx(&bars[SB_X_FULL + slot], ST_XSLOT);
                    for (int k = 0; k < nc; ++k) tma_load_3d(dst + k * rows * (ST_XW * 4), &xmaps.m[k], t0 - 4, rows * cc, b, &bars[SB_X_FULL + slot]);
                }
            }
            __syncwarp();
        }
    }
    if (tid == 256 && P.y_tma) bulk_wait_group0();  // the last tile's store has completed
    if (tr && lane == 0 && (w == 0 || w == 8 || w >= 16)) {  // one warp per role
        long long *o = P.trace + 8 * (w == 0 ? 0 : w == 8 ? 1 : w - 14);
        o[0] = clock64() - tr_total; o[1] = (long long)tr_a; o[2] = (long long)tr_b; o[3] = (long long)tr_c; o[4] = (long long)tr_d; o[5] = n_my;
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    if (w == 16) tmem_dealloc(tmem, 512);
}

// y[b][c][t] = snake(x[b][c][t], alpha[c]): the activation of the FIRST tensor-core block's input, computed once instead of once per
// 128-channel output tile (8 times for 1024 -> 1024); the later blocks get theirs from the previous epilogue (post_alpha)
__global__ void __launch_bounds__(256) snake_kernel(const float *__restrict__ x, long long x_sb, long long x_sc, const float *__restrict__ alpha,
                                                    float *__restrict__ y, long long y_sb, long long y_sc, int C, int T) {
    // one CTA = 8 channel rows of one item: warp = row, 4 independent loads in flight per lane
    const int c = 8 * blockIdx.x + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
    if (c >= C) return;
    const float a = __ldg(alpha + c), inv_a = snake_inv(a);
    const float *xr = x + (long long)b * x_sb + (long long)c * x_sc;
    float *yr = y + (long long)b * y_sb + (long long)c * y_sc;
    int t = lane;
    for (; t + 96 < T; t += 128) {
        const float v0 = __ldcs(xr + t), v1 = __ldcs(xr + t + 32), v2 = __ldcs(xr + t + 64), v3 = __ldcs(xr + t + 96);
        yr[t] = snake_tc(v0, a, inv_a); yr[t + 32] = snake_tc(v1, a, inv_a); yr[t + 64] = snake_tc(v2, a, inv_a); yr[t + 96] = snake_tc(v3, a, inv_a);
    }
    for (; t < T; t += 32) yr[t] = snake_tc(__ldcs(xr + t), a, inv_a);
}

// ---- the narrow tail of the subnet in ONE launch ----------------------------------------------------------------------------------
// models/importance_subnet.py's last three blocks (128 -> 32 -> 8 -> 1, then the sigmoid of line 43) are 13 kMAC per frame -- three
// launches of the generic block kernel spent ~100 us at config-2 size on launch latency and half-empty grids.  Here one CTA of 256
// threads takes 60 output frames through all three: it stages the 128-channel input for frames t0 - 3 .. t0 + 62 and block A's weights
// with cp.async (everything in flight at once; Snake applied in place unless the producer already stored the input activated),
// computes block A on 64 positions (warp = 8 of its 32 channels x 32 positions, lane = position), block B on 62, block C + sigmoid on
// 60, each block reading its predecessor's Snake-activated output from shared memory.  Positions outside [0, T) are the convolutions'
// zero padding at every level.  Weights: the generic kernel's packed layout [(ci * 3 + tap)][128 columns].
constexpr int TL_C0 = 128, TL_C1 = 32, TL_C2 = 8, TL_F = 60, TL_NA = TL_F + 6, TL_AP = 68, TL_HP = 65, TL_NTH = 256;
constexpr int TL_SM_A = 0, TL_SM_W3 = TL_SM_A + TL_C0 * TL_AP * 4, TL_SM_H3 = TL_SM_W3 + TL_C0 * 3 * TL_C1 * 4, TL_SM_W4 = TL_SM_H3 + TL_C1 * TL_HP * 4;
constexpr int TL_SM_H4 = TL_SM_W4 + TL_C1 * 3 * TL_C2 * 4, TL_SM_W5 = TL_SM_H4 + TL_C2 * 64 * 4, TL_SMEM = TL_SM_W5 + TL_C2 * 3 * 4;
static_assert(TL_SM_W3 % 16 == 0 && TL_SM_W4 % 16 == 0, "16-byte cp.async / vector loads");

struct TailParams {
    const float *x;
    long long x_sb, x_sc;
    const float *alpha0, *w0, *bias0, *alpha1, *w1, *bias1, *alpha2, *w2, *bias2;
    float *y;
    long long y_sb;
    int B, T, pre_activated;
};

__global__ void __launch_bounds__(TL_NTH) subnet_tail_kernel(const TailParams P) {
    extern __shared__ __align__(16) unsigned char tsm[];
    float *A = reinterpret_cast<float *>(tsm + TL_SM_A), *W3 = reinterpret_cast<float *>(tsm + TL_SM_W3), *H3 = reinterpret_cast<float *>(tsm + TL_SM_H3);
    float *W4 = reinterpret_cast<float *>(tsm + TL_SM_W4), *H4 = reinterpret_cast<float *>(tsm + TL_SM_H4), *W5 = reinterpret_cast<float *>(tsm + TL_SM_W5);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, cg = w & 3, pos = 32 * (w >> 2) + lane;
    const int b = blockIdx.y, t0 = blockIdx.x * TL_F;
    // weights: the first 32 / 8 / 1 columns of the packed rows
    for (int i = tid; i < TL_C0 * 3 * TL_C1 / 4; i += TL_NTH) cp_async<16>(W3 + 4 * i, P.w0 + (size_t)(i >> 3) * VRVQ_CONV3_COUT_ALIGN + 4 * (i & 7));
    for (int i = tid; i < TL_C1 * 3 * TL_C2 / 4; i += TL_NTH) cp_async<16>(W4 + 4 * i, P.w1 + (size_t)(i >> 1) * VRVQ_CONV3_COUT_ALIGN + 4 * (i & 1));
    if (tid < TL_C2 * 3) W5[tid] = __ldg(P.w2 + (size_t)tid * VRVQ_CONV3_COUT_ALIGN);
    // input: A[ci][q] = x[b][ci][t0 - 3 + q], q < 66; zeros outside [0, T) (cp.async with a source size of 0 writes zeros)
    {
        const float *xb = P.x + (long long)b * P.x_sb;
        for (int i = tid; i < TL_C0 * TL_NA; i += TL_NTH) {
            const int ci = i / TL_NA, q = i - ci * TL_NA, t = t0 - 3 + q;
            const bool ok = t >= 0 && t < P.T;
            const float *src = xb + (long long)ci * P.x_sc + (ok ? t : 0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(A + ci * TL_AP + q)), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    if (!P.pre_activated) {
        for (int i = tid; i < TL_C0 * TL_NA; i += TL_NTH) {
            const int ci = i / TL_NA, q = i - ci * TL_NA;
            const float a = __ldg(P.alpha0 + ci);
            A[ci * TL_AP + q] = snake_tc(A[ci * TL_AP + q], a, snake_inv(a));  // (snake(0) = 0: the padding stays zero)
        }
        __syncthreads();
    }
    // block A: position pos (frame t0 - 2 + pos), channels 8 cg .. 8 cg + 7
    {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = __ldg(P.bias0 + 8 * cg + j);
#pragma unroll 4
        for (int ci = 0; ci < TL_C0; ++ci) {
            const float *ar = A + ci * TL_AP + pos;
            const float a3[3] = {ar[0], ar[1], ar[2]};
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
                const float4 w0 = *reinterpret_cast<const float4 *>(W3 + (ci * 3 + tap) * TL_C1 + 8 * cg);
                const float4 w1 = *reinterpret_cast<const float4 *>(W3 + (ci * 3 + tap) * TL_C1 + 8 * cg + 4);
                acc[0] = fmaf(w0.x, a3[tap], acc[0]); acc[1] = fmaf(w0.y, a3[tap], acc[1]); acc[2] = fmaf(w0.z, a3[tap], acc[2]); acc[3] = fmaf(w0.w, a3[tap], acc[3]);
                acc[4] = fmaf(w1.x, a3[tap], acc[4]); acc[5] = fmaf(w1.y, a3[tap], acc[5]); acc[6] = fmaf(w1.z, a3[tap], acc[6]); acc[7] = fmaf(w1.w, a3[tap], acc[7]);
            }
        }
        const int t = t0 - 2 + pos;
        const bool ok = t >= 0 && t < P.T;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = __ldg(P.alpha1 + 8 * cg + j);
            H3[(8 * cg + j) * TL_HP + pos] = ok ? snake_tc(acc[j], a, snake_inv(a)) : 0.0f;
        }
    }
    __syncthreads();
    // block B: position pos < 62 (frame t0 - 1 + pos), channels 2 cg, 2 cg + 1
    if (pos < TL_F + 2) {
        float acc0 = __ldg(P.bias1 + 2 * cg), acc1 = __ldg(P.bias1 + 2 * cg + 1);
#pragma unroll 4
        for (int ci = 0; ci < TL_C1; ++ci) {
            const float *hr = H3 + ci * TL_HP + pos;
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
                const float2 wv = *reinterpret_cast<const float2 *>(W4 + (ci * 3 + tap) * TL_C2 + 2 * cg);
                acc0 = fmaf(wv.x, hr[tap], acc0);
                acc1 = fmaf(wv.y, hr[tap], acc1);
            }
        }
        const int t = t0 - 1 + pos;
        const bool ok = t >= 0 && t < P.T;
        const float a0 = __ldg(P.alpha2 + 2 * cg), a1 = __ldg(P.alpha2 + 2 * cg + 1);
        H4[(2 * cg) * 64 + pos] = ok ? snake_tc(acc0, a0, snake_inv(a0)) : 0.0f;
        H4[(2 * cg + 1) * 64 + pos] = ok ? snake_tc(acc1, a1, snake_inv(a1)) : 0.0f;
    }
    __syncthreads();
    // block C + sigmoid: frame t0 + tid, tid < 60
    if (tid < TL_F && t0 + tid < P.T) {
        float acc = __ldg(P.bias2);
#pragma unroll
        for (int ci = 0; ci < TL_C2; ++ci)
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) acc = fmaf(W5[ci * 3 + tap], H4[ci * 64 + tid + tap], acc);
        P.y[(long long)b * P.y_sb + t0 + tid] = 1.0f / (1.0f + expf(-acc));
    }
}

}  // namespace

int launch_snake(const float *x, long long x_sb, long long x_sc, const float *alpha, int B, int C, int T, float *y, long long y_sb, long long y_sc,
                 cudaStream_t st) {
    if ((long long)B * C * T == 0) return VRVQ_OK;
    if (B > 65535) {
        set_error("vrvq_snake_f32: B must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    snake_kernel<<<dim3((unsigned)((C + 7) / 8), (unsigned)B), 256, 0, st>>>(x, x_sb, x_sc, alpha, y, y_sb, y_sc, C, T);
    return check_cuda(cudaGetLastError(), "snake_kernel launch");
}

int subnet_tail_usable(int C0, int C1, int C2) { return C0 == TL_C0 && C1 == TL_C1 && C2 == TL_C2; }

int launch_subnet_tail(const float *x, long long x_sb, long long x_sc, int pre_activated, const float *alpha0, const float *w0, const float *bias0,
                       const float *alpha1, const float *w1, const float *bias1, const float *alpha2, const float *w2, const float *bias2, int B, int T,
                       float *y, long long y_sb, cudaStream_t st) {
    if ((long long)B * T == 0) return VRVQ_OK;
    if (B > 65535) {
        set_error("vrvq_subnet_tail_f32: B must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    TailParams P{x, x_sb, x_sc, alpha0, w0, bias0, alpha1, w1, bias1, alpha2, w2, bias2, y, y_sb, B, T, pre_activated};
    int rc = ensure_dynamic_smem<subnet_tail_kernel>(TL_SMEM, "cudaFuncSetAttribute(subnet_tail_kernel)");
    if (rc) return rc;
    subnet_tail_kernel<<<dim3((unsigned)((T + TL_F - 1) / TL_F), (unsigned)B), TL_NTH, TL_SMEM, st>>>(P);
    return check_cuda(cudaGetLastError(), "subnet_tail_kernel launch");
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

template <int LG>
static int launch_lg(const SnTcParams &P, const ZMaps &maps, const CUtensorMap &ymap, int grid, cudaStream_t st) {
    int rc = ensure_dynamic_smem<snake_conv3_tc_kernel<LG>>(ST_SMEM, "cudaFuncSetAttribute(snake_conv3_tc_kernel)");
    if (rc) return rc;
    snake_conv3_tc_kernel<LG><<<grid, ST_NTH, ST_SMEM, st>>>(P, maps, ymap);
    return check_cuda(cudaGetLastError(), "snake_conv3_tc_kernel launch");
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
    CUtensorMap ymap;
    P.y_tma = build_store_map(y, B, Cout, T, y_sc, y_sb, 128, 128, &ymap) ? 1 : 0;  // else: per-lane stores
    if (getenv("VRVQ_SUBNET_NO_TMA_STORE")) P.y_tma = 0;
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
    const int grid = P.n_tiles < sms ? P.n_tiles : sms;
    const bool trace = getenv("VRVQ_SUBNET_TRACE") != nullptr;  // profiling only: synchronous, allocates
    if (trace && (cudaMalloc(&P.trace, 48 * sizeof(long long)) != cudaSuccess || cudaMemsetAsync(P.trace, 0, 48 * sizeof(long long), st) != cudaSuccess)) P.trace = nullptr;
    int rc = P.nc_log2 == 0 ? launch_lg<0>(P, maps, ymap, grid, st) : P.nc_log2 == 1 ? launch_lg<1>(P, maps, ymap, grid, st) : launch_lg<2>(P, maps, ymap, grid, st);
    if (P.trace != nullptr) {
        long long h[48];
        if (rc == 0 && cudaStreamSynchronize(st) == cudaSuccess && cudaMemcpy(h, P.trace, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
            static const char *names[6] = {"loader   [x_full empty st+next_load+arrive -]", "drain    [set_full final_store - -]", "issuer 0 [full - set_empty issue]", "issuer 1 [full - set_empty issue]",
                                           "w_prod   [w_empty - - -]", "x_prod   [x_empty - - -]"};
            fprintf(stderr, "[vrvq subnet trace] %d -> %d, B=%d T=%d, block 0: %lld tiles\n", Cin, Cout, B, T, h[5]);
            for (int r = 0; r < 6; ++r)
                fprintf(stderr, "  %-42s total %9lld | %9lld %9lld %9lld %9lld\n", names[r], h[8 * r], h[8 * r + 1], h[8 * r + 2], h[8 * r + 3], h[8 * r + 4]);
        }
        cudaFree(P.trace);
    }
    return rc;
}

}  // namespace vrvq
