// Fused residual-vector-quantisation encode for B200 (sm_100a).
//
// One persistent CTA per SM walks tiles of TF=32 consecutive latent frames of one batch item.
// Per tile the [D x 32] residual lives in shared memory for all stages, the z_q accumulators live
// in tensor memory (each thread parks its 2 x D/32 partial sums in its own TMEM lane and streams them
// through tcgen05.ld/st once per stage, which keeps the register file free for pipelined operands),
// and the per-stage weights (in_proj / codebook / out_proj pieces, ~36 KB each) are
// streamed L2 -> shared memory through a two-slot ring filled by cp.async.bulk (TMA unit) and
// tracked with mbarriers, one piece ahead of the math.  The latent is read from HBM exactly once;
// codes, latents, mask, z_q and (optionally) z_q_is are written exactly once.
//
// Stage i, per frame (reference: models/quantize.py:42-103, loop :182-202 / :353-365):
//   in_proj   z_e = W_in r + b_in                      (quantize.py:66)     CUDA cores, fp32 FMA
//   normalise e = z_e / max(||z_e||, 1e-12)            (quantize.py:92)     exact op order of torch
//   search    argmin_j fl(fl(e2 - (2e).c_j) + c2_j)    (quantize.py:96-101) first index on ties
//   gather    c = codebook_raw[idx]; q = z_e + (c - z_e)  (quantize.py:73-75,102)
//   out_proj  z_q_i = W_out q + b_out                  (quantize.py:77)
//   update    r -= z_q_i ; z_q += mask_i * z_q_i       (quantize.py:194-195,360,421)
// The hard importance mask (models/utils.py:55-61), the masked loss sums (quantize.py:422-423) and
// the per-stage kept-frame counts (numerator of models/utils.py:64-73) are produced in the same pass.
//
// Arithmetic that must be reproduced bit-for-bit uses explicit __f*_rn intrinsics (packed
// fma.rn.f32x2 where two independent IEEE FMAs share an instruction); the file is compiled with
// --fmad=false so nothing else is contracted behind our back.
//
// Shared-memory cost model used for the thread mappings (measured with ncu on B200, profiles/r1a_*):
// an LDS.128 whose lanes read distinct 16-byte chunks costs 4 wavefronts (512 B), identical quarter-warps
// are NOT merged; a broadcast LDS.128 costs one wavefront per distinct address; an LDS.64 with 4 distinct
// 8-byte addresses costs 1.  So every quarter-warp reads a different residual row, and broadcast operands
// are fetched with as few distinct addresses per instruction as the tiling allows.
#include "common.cuh"

namespace vrvq {

constexpr int TF = 32;   // frames per tile
constexpr int NT = 512;  // threads per CTA
constexpr int NW = NT / 32;

struct EncodeParams {
    const float *blob;
    const float *z;
    long long z_sb, z_sd;
    const float *imp;
    long long imp_sb;
    const float *level_dev;
    long long level_stride;
    float level_host;
    long long *codes;
    long long codes_sb, codes_sq;
    float *z_q;
    long long zq_sb, zq_sd;
    float *z_q_is;
    long long zqis_sb, zqis_sq, zqis_sd;
    float *latents;
    long long lat_sb, lat_sc;
    float *mask;
    long long mask_sb, mask_sq;
    float *loss_pf;
    long long loss_sb, loss_sq;
    double *loss_sum;
    unsigned long long *kept;
    int B, T, Nq, n_run, tiles_per_b, n_tiles;
    int vec_ld;  // 4 / 2 / 1 floats per global load of z
    int vec_st;  // 2 / 1 floats per global store of z_q, z_q_is
};

template <int D, int K>
struct EncodeSmem {
    static constexpr BlobLayout L = BlobLayout(D, K);
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int R_FLOATS = D * TF;
    static constexpr int WB_FLOATS = (cmax(cmax(L.p0_floats(), L.p1_floats()), L.p2_floats()) + 3) / 4 * 4;
    static constexpr int PART_FLOATS = NW * CD * TF;
    static constexpr int OFF_R = 0;
    static constexpr int OFF_WB0 = OFF_R + R_FLOATS;
    static constexpr int OFF_WB1 = OFF_WB0 + WB_FLOATS;
    static constexpr int OFF_PART = OFF_WB1 + WB_FLOATS;
    static constexpr int OFF_ZE = OFF_PART + PART_FLOATS;
    static constexpr int OFF_ES = OFF_ZE + CD * TF;
    static constexpr int OFF_E2 = OFF_ES + CD * TF;
    static constexpr int OFF_NKEEP = OFF_E2 + TF;
    static constexpr int OFF_BARS = OFF_NKEEP + TF;  // 2 x uint64
    static constexpr int OFF_TMEM = OFF_BARS + 4;    // TMEM base address written by tcgen05.alloc
    static constexpr int TOTAL_FLOATS = OFF_TMEM + 4;
    static constexpr int BYTES = TOTAL_FLOATS * 4;
    static_assert(OFF_BARS % 2 == 0, "mbarriers need 8-byte alignment");
    static_assert(OFF_WB0 % 4 == 0 && OFF_WB1 % 4 == 0, "bulk copy destinations need 16-byte alignment");
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

// global -> shared load of one [D x TF] latent tile; frames >= fv are zero-filled.
template <int D, int VEC>
__device__ __forceinline__ void load_tile(float *R, const float *zb, long long z_sd, int fv, int tid) {
    constexpr int CPR = TF / VEC;  // chunks per row
#pragma unroll 4
    for (int c = tid; c < D * CPR; c += NT) {
        const int d = c / CPR, q = c % CPR;
        float *dst = R + d * TF + q * VEC;
        if (q * VEC < fv) {
            cp_async<VEC * 4>(dst, zb + (long long)d * z_sd + q * VEC);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = 0.0f;
        }
    }
}

__device__ __forceinline__ float2 dup2(float x) { return make_float2(x, x); }

// out_proj + residual update + masked accumulate for one thread: channels dbase + 2i (i < NI), frames f0, f0+1.
// FIRST / LAST and the store shapes are compile-time, so the unrolled body has no branches and the scheduler can
// overlap the shared-memory loads of the next channels with the FMA chains of the current ones.
// The z_q accumulators stream through TMEM in groups of 8 (4 channels x 2 frames): load (unless FIRST), fma with the
// 0/1 mask, store back (unless LAST, where the finished sums go straight to global memory).
template <int NI, bool FIRST, bool LAST, bool ZQIS, int VEC_ST>
__device__ __forceinline__ void out_proj_thread(const float *__restrict__ wp, const float *__restrict__ bp, float *__restrict__ rp,
                                                const float (&q0)[CD], const float (&q1)[CD], float m0, float m1, uint32_t tacc,
                                                float *zo, long long zstep, float *zq, long long zqstep, bool v0ok, bool v1ok) {
    static_assert(NI % 4 == 0, "channels per thread must be a multiple of 4");
#pragma unroll
    for (int g = 0; g < NI / 4; ++g) {
        uint32_t a8[8];
        if (!FIRST) tmem_ld8(tacc + 8 * g, a8);
        float v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * g + u;
            const float4 wa = *reinterpret_cast<const float4 *>(wp + i * 2 * CD);
            const float4 wb = *reinterpret_cast<const float4 *>(wp + i * 2 * CD + 4);
            float v0 = bp[2 * i], v1 = v0;
            v0 = __fmaf_rn(wa.x, q0[0], v0); v1 = __fmaf_rn(wa.x, q1[0], v1);
            v0 = __fmaf_rn(wa.y, q0[1], v0); v1 = __fmaf_rn(wa.y, q1[1], v1);
            v0 = __fmaf_rn(wa.z, q0[2], v0); v1 = __fmaf_rn(wa.z, q1[2], v1);
            v0 = __fmaf_rn(wa.w, q0[3], v0); v1 = __fmaf_rn(wa.w, q1[3], v1);
            v0 = __fmaf_rn(wb.x, q0[4], v0); v1 = __fmaf_rn(wb.x, q1[4], v1);
            v0 = __fmaf_rn(wb.y, q0[5], v0); v1 = __fmaf_rn(wb.y, q1[5], v1);
            v0 = __fmaf_rn(wb.z, q0[6], v0); v1 = __fmaf_rn(wb.z, q1[6], v1);
            v0 = __fmaf_rn(wb.w, q0[7], v0); v1 = __fmaf_rn(wb.w, q1[7], v1);
            if (!LAST) {
                float2 *rr = reinterpret_cast<float2 *>(rp + i * 2 * TF);
                float2 r = *rr;
                r.x = __fsub_rn(r.x, v0);
                r.y = __fsub_rn(r.y, v1);
                *rr = r;
            }
            if (ZQIS) {
                if (VEC_ST == 2) {
                    if (v0ok) st_cs2(zo, v0, v1);  // fv is even whenever VEC_ST == 2
                } else {
                    if (v0ok) st_cs(zo, v0);
                    if (v1ok) st_cs(zo + 1, v1);
                }
                zo += zstep;
            }
            v[2 * u] = v0;
            v[2 * u + 1] = v1;
        }
        if (!FIRST) tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // z_q += mask * z_q_i  (quantize.py:194 / :421), ascending stage order from 0
            const float a0 = FIRST ? 0.0f : __uint_as_float(a8[2 * u]);
            const float a1 = FIRST ? 0.0f : __uint_as_float(a8[2 * u + 1]);
            a8[2 * u] = __float_as_uint(__fmaf_rn(m0, v[2 * u], a0));
            a8[2 * u + 1] = __float_as_uint(__fmaf_rn(m1, v[2 * u + 1], a1));
        }
        if (!LAST) {
            tmem_st8(tacc + 8 * g, a8);
        } else if (zq != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (VEC_ST == 2) {
                    if (v0ok) st_cs2(zq, __uint_as_float(a8[2 * u]), __uint_as_float(a8[2 * u + 1]));
                } else {
                    if (v0ok) st_cs(zq, __uint_as_float(a8[2 * u]));
                    if (v1ok) st_cs(zq + 1, __uint_as_float(a8[2 * u + 1]));
                }
                zq += zqstep;
            }
        }
    }
    if (!LAST) tmem_wait_st();
}

template <int D, int K, int VEC_ST, bool ZQIS>
__global__ void __launch_bounds__(NT, 1) rvq_encode_kernel(const EncodeParams p) {
    using S = EncodeSmem<D, K>;
    constexpr BlobLayout L = BlobLayout(D, K);
    constexpr int DW = D / NW;  // channels per warp
    static_assert(D % (NW * 4) == 0, "D must be a multiple of 64");
    static_assert(K % 64 == 0, "codebook size must be a multiple of 64");

    extern __shared__ __align__(128) float smem[];
    float *R = smem + S::OFF_R;
    float *wb0 = smem + S::OFF_WB0;
    float *wb1 = smem + S::OFF_WB1;
    float *part = smem + S::OFF_PART;  // [NW][CD][TF]; re-used as sbest/sidx after the reduce
    float *sbest = part;               // [NW][TF]
    int *sidx = reinterpret_cast<int *>(part + NW * TF);
    float *ze = smem + S::OFF_ZE;  // [CD][TF] pre-normalisation latents
    float *es = smem + S::OFF_ES;  // [CD][TF] 2*e
    float *e2s = smem + S::OFF_E2;
    int *nkeep = reinterpret_cast<int *>(smem + S::OFF_NKEEP);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::OFF_TMEM);
    constexpr uint32_t TMEM_COLS = (NW / 4) * DW;  // accumulator columns: DW per thread, 4 lane quadrants
    static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM allocation must be a power of two in [32, 512]");

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_run = p.n_run;
    const int n_my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (n_my_tiles <= 0) return;
    const uint32_t n_pieces = (uint32_t)n_my_tiles * (uint32_t)n_run * 3u;
    const float *stages = p.blob + BLOB_HDR_FLOATS;

    // ---- weight ring ----------------------------------------------------------------------
    uint32_t piece = 0;  // uniform across the CTA
    auto issue_piece = [&](uint32_t n) {  // thread 0 only
        const uint32_t ph = n % 3u, s = (n / 3u) % (uint32_t)n_run;
        const float *src = stages + (size_t)s * L.stage_floats();
        uint32_t bytes;
        if (ph == 0) {
            src += L.off_p0();
            bytes = L.p0_floats() * 4;
        } else if (ph == 1) {
            src += L.off_p1();
            bytes = L.p1_floats() * 4;
        } else {
            src += L.off_p2();
            bytes = L.p2_floats() * 4;
        }
        uint64_t *bar = &bars[n & 1u];
        fence_proxy_async();  // generic-proxy reads of the slot are ordered before the async-proxy overwrite
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s((n & 1u) ? wb1 : wb0, src, bytes, bar);
    };
    auto acquire = [&]() -> const float * {  // all threads, at the start of every phase
        if (tid == 0 && piece + 1 < n_pieces) issue_piece(piece + 1);
        mbar_wait(&bars[piece & 1u], (piece >> 1) & 1u);
        const float *buf = (piece & 1u) ? wb1 : wb0;
        ++piece;
        return buf;
    };

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    if (w == 0) {  // one warp owns the TMEM allocation for the CTA's lifetime
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // this thread's accumulator strip: its own lane of quadrant w%4, DW columns starting at (w/4)*DW
    const uint32_t tacc = tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * DW);
    if (tid == 0) issue_piece(0);

    auto tile_coords = [&](int it, int &b, int &t0, int &fv) {
        const int tile = (int)blockIdx.x + it * (int)gridDim.x;
        b = tile / p.tiles_per_b;
        t0 = (tile % p.tiles_per_b) * TF;
        fv = min(TF, p.T - t0);
    };
    auto start_tile_load = [&](int it) {
        int b, t0, fv;
        tile_coords(it, b, t0, fv);
        const float *zb = p.z + (long long)b * p.z_sb + t0;
        if (p.vec_ld == 4)
            load_tile<D, 4>(R, zb, p.z_sd, fv, tid);
        else if (p.vec_ld == 2)
            load_tile<D, 2>(R, zb, p.z_sd, fv, tid);
        else
            load_tile<D, 1>(R, zb, p.z_sd, fv, tid);
        cp_async_commit();
    };
    start_tile_load(0);

    // thread coordinates of the phase mappings
    const int l4 = lane & 7, g4 = lane >> 3;   // in_proj : frames 4*l4..+3, quarter-warp g4 = K sub-slice
    const int l2 = lane & 15, g2 = lane >> 4;  // search / out_proj : frames 2*l2, 2*l2+1; half-warp g2
    const int f0 = 2 * l2;

    double loss_acc = 0.0;               // lane 0 of warp 0
    unsigned long long kept_acc = 0ull;  // lane k of warp 0 counts stage k

    for (int it = 0; it < n_my_tiles; ++it) {
        int b, t0, fv;
        tile_coords(it, b, t0, fv);

        // ---- per-frame keep counts: mask[k] = (imp*level*Nq - k >= 0)   (quantize.py:389, utils.py:59-60)
        if (w == 0) {
            int nk = 0;
            if (lane < fv) {
                if (p.imp != nullptr) {
                    const float lv = p.level_dev ? p.level_dev[(long long)b * p.level_stride] : p.level_host;
                    const float x = __fmul_rn(__fmul_rn(p.imp[(long long)b * p.imp_sb + t0 + lane], lv), (float)p.Nq);
                    for (int k = 0; k < n_run; ++k) nk += (__fsub_rn(x, (float)k) >= 0.0f) ? 1 : 0;
                } else {
                    nk = n_run;
                }
            }
            nkeep[lane] = nk;
            for (int k = 0; k < n_run; ++k) {
                const bool on = nk > k;  // the mask is a prefix of ones: x - k is decreasing in k
                const unsigned bal = __ballot_sync(0xffffffffu, on);
                if (lane == k) kept_acc += (unsigned long long)__popc(bal);
                if (p.mask != nullptr && lane < fv)
                    p.mask[(long long)b * p.mask_sb + (long long)k * p.mask_sq + t0 + lane] = on ? 1.0f : 0.0f;
            }
        }
        cp_async_wait_all();
        __syncthreads();  // residual tile and nkeep visible

        for (int s = 0; s < n_run; ++s) {
            const bool last = (s == n_run - 1);
            // ================= in_proj: z_e[c][f] = sum_d W_in[c][d] r[d][f] =================
            // Thread tile 8 channels x 4 frames; the four quarter-warps take four consecutive residual rows per step
            // (so one LDS.128 delivers 512 useful bytes), i.e. K is split 4 ways inside the warp and NW ways across warps.
            {
                const float *W = acquire();
                float a[CD][4];
#pragma unroll
                for (int c = 0; c < CD; ++c) a[c][0] = a[c][1] = a[c][2] = a[c][3] = 0.0f;
                const float *Wp = W + (w * DW + g4) * CD;
                const float *Rp = R + (w * DW + g4) * TF + 4 * l4;
#pragma unroll 4
                for (int i = 0; i < DW / 4; ++i) {
                    const float4 rv = *reinterpret_cast<const float4 *>(Rp + i * 4 * TF);
                    const float4 wa = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD);
                    const float4 wb = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD + 4);
                    const float wv[CD] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int c = 0; c < CD; ++c) {
                        a[c][0] = fmaf(wv[c], rv.x, a[c][0]);
                        a[c][1] = fmaf(wv[c], rv.y, a[c][1]);
                        a[c][2] = fmaf(wv[c], rv.z, a[c][2]);
                        a[c][3] = fmaf(wv[c], rv.w, a[c][3]);
                    }
                }
                // reduce-scatter over the four K sub-slices: after lane^16 a thread keeps 4 channels, after lane^8 two.
                float h[4][4];
                const bool up16 = (g4 & 2) != 0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const float keep = up16 ? a[c + 4][f] : a[c][f];
                        const float send = up16 ? a[c][f] : a[c + 4][f];
                        h[c][f] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 16));
                    }
                float q2[2][4];
                const bool up8 = (g4 & 1) != 0;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        const float keep = up8 ? h[c + 2][f] : h[c][f];
                        const float send = up8 ? h[c][f] : h[c + 2][f];
                        q2[c][f] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, 8));
                    }
                const int c0 = 4 * (g4 >> 1) + 2 * (g4 & 1);
                *reinterpret_cast<float4 *>(&part[(w * CD + c0) * TF + 4 * l4]) = make_float4(q2[0][0], q2[0][1], q2[0][2], q2[0][3]);
                *reinterpret_cast<float4 *>(&part[(w * CD + c0 + 1) * TF + 4 * l4]) = make_float4(q2[1][0], q2[1][1], q2[1][2], q2[1][3]);
                __syncthreads();
                // the residual is dead after the last stage's in_proj: start fetching the next tile
                if (last && it + 1 < n_my_tiles) start_tile_load(it + 1);
                if (tid < CD * TF) {
                    const int c = w;  // tid >> 5
                    float sacc = part[c * TF + lane];
#pragma unroll
                    for (int ww = 1; ww < NW; ++ww) sacc = __fadd_rn(sacc, part[(ww * CD + c) * TF + lane]);
                    ze[c * TF + lane] = __fadd_rn(sacc, W[D * CD + c]);
                }
                __syncthreads();
                // normalise (quantize.py:92): torch's exact op order, SURVEY.md A.3
                if (tid < CD * TF) {
                    const int c = w;
                    float x[CD];
#pragma unroll
                    for (int k = 0; k < CD; ++k) x[k] = ze[k * TF + lane];
                    float ss = __fmul_rn(x[0], x[0]);
#pragma unroll
                    for (int k = 1; k < CD; ++k) ss = __fadd_rn(ss, __fmul_rn(x[k], x[k]));
                    const float den = fmaxf(__fsqrt_rn(ss), 1e-12f);
                    float xc = x[0];
#pragma unroll
                    for (int k = 1; k < CD; ++k) xc = (c == k) ? x[k] : xc;
                    es[c * TF + lane] = __fmul_rn(2.0f, __fdiv_rn(xc, den));
                    if (c == 0) {
                        float e = __fdiv_rn(x[0], den);
                        float e2 = __fmul_rn(e, e);
#pragma unroll
                        for (int k = 1; k < CD; ++k) {
                            e = __fdiv_rn(x[k], den);
                            e2 = __fadd_rn(e2, __fmul_rn(e, e));
                        }
                        e2s[lane] = e2;
                    }
                    if (p.latents != nullptr && lane < fv)
                        p.latents[(long long)b * p.lat_sb + (long long)(s * CD + c) * p.lat_sc + t0 + lane] = xc;
                }
                __syncthreads();
            }
            // ================= search over the normalised codebook =================
            // Lane = frame; warp w scans code pairs [w*K/32, (w+1)*K/32) in ascending order with warp-uniform
            // (broadcast) codebook reads.  The codebook piece is pair-interleaved ([pair][k][2]) so that one
            // fma.rn.f32x2 advances the dot products of two adjacent codes.
            {
                const float *CB = acquire();
                const float *c2 = CB + K * CD;
                float2 ea[CD];  // (2e_k, 2e_k) of frame `lane`
#pragma unroll
                for (int k = 0; k < CD; ++k) ea[k] = dup2(es[k * TF + lane]);
                const float2 e2a = dup2(e2s[lane]);
                float best = __int_as_float(0x7f800000);
                int bi = 0;
                constexpr int PPW = K / 2 / NW;  // code pairs per warp
                const float4 *cp = reinterpret_cast<const float4 *>(CB + (w * PPW) * 2 * CD);
                const float2 *ccp = reinterpret_cast<const float2 *>(c2 + 2 * w * PPW);
#pragma unroll 4
                for (int i = 0; i < PPW; ++i) {
                    const float4 c01 = cp[4 * i], c23 = cp[4 * i + 1], c45 = cp[4 * i + 2], c67 = cp[4 * i + 3];
                    const float2 cc = ccp[i];
                    float2 da = __fmul2_rn(ea[0], make_float2(c01.x, c01.y));
                    da = __ffma2_rn(ea[1], make_float2(c01.z, c01.w), da);
                    da = __ffma2_rn(ea[2], make_float2(c23.x, c23.y), da);
                    da = __ffma2_rn(ea[3], make_float2(c23.z, c23.w), da);
                    da = __ffma2_rn(ea[4], make_float2(c45.x, c45.y), da);
                    da = __ffma2_rn(ea[5], make_float2(c45.z, c45.w), da);
                    da = __ffma2_rn(ea[6], make_float2(c67.x, c67.y), da);
                    da = __ffma2_rn(ea[7], make_float2(c67.z, c67.w), da);
                    // dist = fl(fl(e2 - dot) + c2)
                    const float2 ta = __fadd2_rn(__fadd2_rn(e2a, make_float2(-da.x, -da.y)), cc);
                    const int j = 2 * (w * PPW + i);
                    if (ta.x < best) { best = ta.x; bi = j; }
                    if (ta.y < best) { best = ta.y; bi = j + 1; }
                }
                sbest[w * TF + lane] = best;
                sidx[w * TF + lane] = bi;
                __syncthreads();
            }
            // ===== argmin merge, gather, loss, straight-through: every warp redundantly, lane = frame (no extra barrier) =====
            float qv[CD];
            const float *WO;
            {
                float best = sbest[lane];
                int bi = sidx[lane];
#pragma unroll
                for (int ww = 1; ww < NW; ++ww) {
                    const float ob = sbest[ww * TF + lane];
                    const int oi = sidx[ww * TF + lane];
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                const float *raw = stages + (size_t)s * L.stage_floats() + L.off_raw() + (size_t)bi * CD;
                const float4 ra = __ldg(reinterpret_cast<const float4 *>(raw));
                const float4 rb = __ldg(reinterpret_cast<const float4 *>(raw + 4));
                WO = acquire();  // out_proj weights: the mbarrier wait overlaps the L2 gather latency
                const float cr[CD] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                float ls = 0.0f;
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float x = ze[k * TF + lane];
                    const float diff = __fsub_rn(x, cr[k]);  // quantize.py:69-71
                    const float sq = __fmul_rn(diff, diff);
                    ls = (k == 0) ? sq : __fadd_rn(ls, sq);
                    qv[k] = __fadd_rn(x, __fsub_rn(cr[k], x));  // quantize.py:73-75
                }
                if (w == 0) {
                    const float loss = __fdiv_rn(ls, (float)CD);
                    const bool valid = lane < fv;
                    if (valid) {
                        p.codes[(long long)b * p.codes_sb + (long long)s * p.codes_sq + t0 + lane] = (long long)bi;
                        if (p.loss_pf != nullptr)
                            p.loss_pf[(long long)b * p.loss_sb + (long long)s * p.loss_sq + t0 + lane] = loss;
                    }
                    double ml = (valid && nkeep[lane] > s) ? (double)loss : 0.0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) ml += __shfl_xor_sync(0xffffffffu, ml, off);
                    loss_acc += ml;
                }
            }
            // ================= out_proj + residual update + masked accumulate =================
            {
                const float *bo = WO + D * CD;
                float q0[CD], q1[CD];
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    q0[k] = __shfl_sync(0xffffffffu, qv[k], f0);
                    q1[k] = __shfl_sync(0xffffffffu, qv[k], f0 + 1);
                }
                const int2 nk = *reinterpret_cast<const int2 *>(&nkeep[f0]);
                const float m0 = nk.x > s ? 1.0f : 0.0f, m1 = nk.y > s ? 1.0f : 0.0f;
                const bool v0ok = f0 < fv, v1ok = f0 + 1 < fv;
                const int dbase = w * DW + g2;
                float *zo = nullptr;
                if (ZQIS) zo = p.z_q_is + (long long)b * p.zqis_sb + (long long)s * p.zqis_sq + (long long)dbase * p.zqis_sd + t0 + f0;
                const long long zstep = 2 * p.zqis_sd;
                const float *wp = WO + dbase * CD;
                const float *bp = bo + dbase;
                float *rp = R + dbase * TF + f0;
                float *zq = nullptr;
                if (p.z_q != nullptr) zq = p.z_q + (long long)b * p.zq_sb + (long long)dbase * p.zq_sd + t0 + f0;
                const long long zqstep = 2 * p.zq_sd;
                const bool first = (s == 0);
                if (first && last)
                    out_proj_thread<DW / 2, true, true, ZQIS, VEC_ST>(wp, bp, rp, q0, q1, m0, m1, tacc, zo, zstep, zq, zqstep, v0ok, v1ok);
                else if (first)
                    out_proj_thread<DW / 2, true, false, ZQIS, VEC_ST>(wp, bp, rp, q0, q1, m0, m1, tacc, zo, zstep, zq, zqstep, v0ok, v1ok);
                else if (last)
                    out_proj_thread<DW / 2, false, true, ZQIS, VEC_ST>(wp, bp, rp, q0, q1, m0, m1, tacc, zo, zstep, zq, zqstep, v0ok, v1ok);
                else
                    out_proj_thread<DW / 2, false, false, ZQIS, VEC_ST>(wp, bp, rp, q0, q1, m0, m1, tacc, zo, zstep, zq, zqstep, v0ok, v1ok);
                __syncthreads();  // residual updated; sbest/ze and the weight slot are free again
            }
        }  // stages

    }  // tiles

    tmem_fence_before_sync();
    __syncthreads();
    if (w == 0) {
        tmem_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
        if (lane == 0 && p.loss_sum != nullptr) atomicAdd(p.loss_sum, loss_acc);
        if (lane < n_run && p.kept != nullptr && kept_acc != 0ull) atomicAdd(&p.kept[lane], kept_acc);
    }
}

// ---- host launcher ---------------------------------------------------------------------------
template <int D, int K, int VEC_ST, bool ZQIS>
static int launch_one(const EncodeParams &p, int grid, cudaStream_t stream) {
    using S = EncodeSmem<D, K>;
    static bool attr_done = false;  // benign race: the attribute is idempotent
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(rvq_encode_kernel<D, K, VEC_ST, ZQIS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES),
                            "cudaFuncSetAttribute(rvq_encode_kernel)");
        if (rc) return rc;
        attr_done = true;
    }
    rvq_encode_kernel<D, K, VEC_ST, ZQIS><<<grid, NT, S::BYTES, stream>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_encode_kernel launch");
}

template <int D, int K>
static int launch_encode(const EncodeParams &p, int grid, cudaStream_t stream) {
    const bool zqis = p.z_q_is != nullptr;
    if (p.vec_st == 2) return zqis ? launch_one<D, K, 2, true>(p, grid, stream) : launch_one<D, K, 2, false>(p, grid, stream);
    return zqis ? launch_one<D, K, 1, true>(p, grid, stream) : launch_one<D, K, 1, false>(p, grid, stream);
}

template <int D, int K>
static int smem_bytes_of() {
    return EncodeSmem<D, K>::BYTES;
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int encode_supported(int D, int K, int cd) { return cd == CD && K == 1024 && (D == 1024 || D == 512 || D == 256); }

static int fill_params(const vrvq_encode_args *a, EncodeParams &p) {
    if (a == nullptr || a->struct_size != sizeof(vrvq_encode_args)) {
        set_error("vrvq_rvq_encode_f32: args is NULL or struct_size mismatch (ABI %d expects %zu bytes)", VRVQ_ABI_VERSION,
                  sizeof(vrvq_encode_args));
        return VRVQ_EINVAL;
    }
    if (a->B < 0 || a->T < 0 || a->n_run < 1 || a->n_run > a->n_codebooks || a->n_run > 32) {
        set_error("vrvq_rvq_encode_f32: bad sizes B=%d T=%d n_run=%d n_codebooks=%d (need 1 <= n_run <= min(n_codebooks, 32))", a->B,
                  a->T, a->n_run, a->n_codebooks);
        return a->n_run > 32 ? VRVQ_EUNSUPPORTED : VRVQ_EINVAL;
    }
    if (!encode_supported(a->input_dim, a->codebook_size, CD)) {
        set_error("vrvq_rvq_encode_f32: no kernel for input_dim=%d codebook_size=%d (built: D in {256,512,1024}, K=1024, codebook_dim=8)",
                  a->input_dim, a->codebook_size);
        return VRVQ_EUNSUPPORTED;
    }
    if (a->blob == nullptr || a->z == nullptr || a->codes == nullptr) {
        set_error("vrvq_rvq_encode_f32: blob, z and codes must be non-NULL");
        return VRVQ_EINVAL;
    }
    if (!aligned(a->blob, 16)) {
        set_error("vrvq_rvq_encode_f32: blob must be 16-byte aligned");
        return VRVQ_EINVAL;
    }
    if (a->imp_map != nullptr && a->level_dev == nullptr && !(a->level_host == a->level_host)) {
        set_error("vrvq_rvq_encode_f32: level is NaN");
        return VRVQ_EINVAL;
    }
    p.blob = static_cast<const float *>(a->blob);
    p.z = a->z; p.z_sb = a->z_stride_b; p.z_sd = a->z_stride_d;
    p.imp = a->imp_map; p.imp_sb = a->imp_stride_b;
    p.level_dev = a->level_dev; p.level_stride = a->level_stride; p.level_host = a->level_host;
    p.codes = reinterpret_cast<long long *>(a->codes); p.codes_sb = a->codes_stride_b; p.codes_sq = a->codes_stride_q;
    p.z_q = a->z_q; p.zq_sb = a->z_q_stride_b; p.zq_sd = a->z_q_stride_d;
    p.z_q_is = a->z_q_is; p.zqis_sb = a->z_q_is_stride_b; p.zqis_sq = a->z_q_is_stride_q; p.zqis_sd = a->z_q_is_stride_d;
    p.latents = a->latents; p.lat_sb = a->latents_stride_b; p.lat_sc = a->latents_stride_c;
    p.mask = a->mask; p.mask_sb = a->mask_stride_b; p.mask_sq = a->mask_stride_q;
    p.loss_pf = a->loss_pf; p.loss_sb = a->loss_pf_stride_b; p.loss_sq = a->loss_pf_stride_q;
    p.loss_sum = a->loss_masked_sum; p.kept = a->kept;
    p.B = a->B; p.T = a->T; p.Nq = a->n_codebooks; p.n_run = a->n_run;
    p.tiles_per_b = (a->T + TF - 1) / TF;
    p.n_tiles = p.tiles_per_b * a->B;
    // widest global access every row start allows (tile starts are multiples of 32 frames)
    p.vec_ld = 1;
    if (aligned(a->z, 16) && a->z_stride_b % 4 == 0 && a->z_stride_d % 4 == 0 && a->T % 4 == 0) p.vec_ld = 4;
    else if (aligned(a->z, 8) && a->z_stride_b % 2 == 0 && a->z_stride_d % 2 == 0 && a->T % 2 == 0) p.vec_ld = 2;
    bool st2 = a->T % 2 == 0;
    if (a->z_q) st2 = st2 && aligned(a->z_q, 8) && a->z_q_stride_b % 2 == 0 && a->z_q_stride_d % 2 == 0;
    if (a->z_q_is) st2 = st2 && aligned(a->z_q_is, 8) && a->z_q_is_stride_b % 2 == 0 && a->z_q_is_stride_q % 2 == 0 && a->z_q_is_stride_d % 2 == 0;
    p.vec_st = st2 ? 2 : 1;
    return VRVQ_OK;
}

static int pick_grid(const EncodeParams &p, int *grid) {
    int dev = 0, sms = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute");
    if (rc) return rc;
    *grid = p.n_tiles < sms ? p.n_tiles : sms;  // persistent: one CTA per SM
    return VRVQ_OK;
}

int encode_launch_info(const vrvq_encode_args *a, int *grid, int *block, int *smem) {
    EncodeParams p{};
    int rc = fill_params(a, p);
    if (rc) return rc;
    int g = 0;
    rc = pick_grid(p, &g);
    if (rc) return rc;
    if (grid) *grid = g;
    if (block) *block = NT;
    if (smem) *smem = a->input_dim == 1024 ? smem_bytes_of<1024, 1024>() : a->input_dim == 512 ? smem_bytes_of<512, 1024>() : smem_bytes_of<256, 1024>();
    return VRVQ_OK;
}

int encode(const vrvq_encode_args *a, void *stream) {
    EncodeParams p{};
    int rc = fill_params(a, p);
    if (rc) return rc;
    if (p.n_tiles == 0) return VRVQ_OK;
    rc = check_device();
    if (rc) return rc;
    int grid = 0;
    rc = pick_grid(p, &grid);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (a->input_dim) {
        case 1024: return launch_encode<1024, 1024>(p, grid, st);
        case 512: return launch_encode<512, 1024>(p, grid, st);
        case 256: return launch_encode<256, 1024>(p, grid, st);
    }
    return VRVQ_EUNSUPPORTED;
}

}  // namespace vrvq
