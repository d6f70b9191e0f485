// Fused residual-vector-quantisation encode for B200 (sm_100a).
//
// One persistent CTA per SM walks tiles of TF=32 consecutive latent frames of one batch item.
// Per tile the [D x 32] residual lives in shared memory for all stages, the z_q accumulators live
// in registers, and the per-stage weights (in_proj / codebook / out_proj pieces, ~36 KB each) are
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
// Arithmetic that must be reproduced bit-for-bit uses explicit __f*_rn intrinsics; the file is
// compiled with --fmad=false so nothing else is contracted behind our back.
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
    static constexpr int OFF_QS = OFF_ES + CD * TF;
    static constexpr int OFF_E2 = OFF_QS + CD * TF;
    static constexpr int OFF_NKEEP = OFF_E2 + TF;
    static constexpr int OFF_BARS = OFF_NKEEP + TF;  // 2 x uint64
    static constexpr int TOTAL_FLOATS = OFF_BARS + 4;
    static constexpr int BYTES = TOTAL_FLOATS * 4;
    static_assert(OFF_BARS % 2 == 0, "mbarriers need 8-byte alignment");
    static_assert(OFF_WB0 % 4 == 0 && OFF_WB1 % 4 == 0, "bulk copy destinations need 16-byte alignment");
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

// global -> shared load of one [D x TF] latent tile; frames >= fv are zero-filled.
template <int D, int VEC>
__device__ __forceinline__ void load_tile(float *R, const float *zb, long long z_sd, int fv, int tid) {
    constexpr int CPR = TF / VEC;  // chunks per row
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

template <int D, int K>
__global__ void __launch_bounds__(NT, 1) rvq_encode_kernel(const EncodeParams p) {
    using S = EncodeSmem<D, K>;
    constexpr BlobLayout L = BlobLayout(D, K);
    constexpr int DW = D / NW;  // channels per warp
    static_assert(D % (NW * 2) == 0, "D must be a multiple of 32");
    static_assert(K % 32 == 0, "codebook size must be a multiple of 32");

    extern __shared__ __align__(128) float smem[];
    float *R = smem + S::OFF_R;
    float *wb0 = smem + S::OFF_WB0;
    float *wb1 = smem + S::OFF_WB1;
    float *part = smem + S::OFF_PART;  // [NW][CD][TF]; re-used as sbest/sidx after the reduce
    float *sbest = part;               // [NW][TF]
    int *sidx = reinterpret_cast<int *>(part + NW * TF);
    float *ze = smem + S::OFF_ZE;  // [CD][TF] pre-normalisation latents
    float *es = smem + S::OFF_ES;  // [CD][TF] 2*e
    float *qs = smem + S::OFF_QS;  // [CD][TF] straight-through values
    float *e2s = smem + S::OFF_E2;
    int *nkeep = reinterpret_cast<int *>(smem + S::OFF_NKEEP);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::OFF_BARS);

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
    __syncthreads();
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

    // thread coordinates of the three phase mappings
    const int l4 = lane & 7, g4 = lane >> 3;    // in_proj : 4 frames x channel pair g4
    const int l2 = lane & 15, g2 = lane >> 4;   // search / out_proj : 2 frames x (codes | channels) g2
    const int f0 = 2 * l2;

    float acc[DW / 2][2];  // z_q accumulators: channels w*DW + 2i + g2, frames f0, f0+1
    double loss_acc = 0.0; // lane 0 of warp 0
    unsigned long long kept_acc = 0ull;  // lane k of warp 0 counts stage k

    for (int it = 0; it < n_my_tiles; ++it) {
        int b, t0, fv;
        tile_coords(it, b, t0, fv);

#pragma unroll
        for (int i = 0; i < DW / 2; ++i) acc[i][0] = acc[i][1] = 0.0f;

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
            {
                const float *W = acquire();
                float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
                const float *Wp = W + (w * DW) * CD + 2 * g4;
                const float *Rp = R + (w * DW) * TF + 4 * l4;
#pragma unroll 8
                for (int i = 0; i < DW; ++i) {
                    const float2 wv = *reinterpret_cast<const float2 *>(Wp + i * CD);
                    const float4 rv = *reinterpret_cast<const float4 *>(Rp + i * TF);
                    a0[0] = fmaf(wv.x, rv.x, a0[0]);
                    a0[1] = fmaf(wv.x, rv.y, a0[1]);
                    a0[2] = fmaf(wv.x, rv.z, a0[2]);
                    a0[3] = fmaf(wv.x, rv.w, a0[3]);
                    a1[0] = fmaf(wv.y, rv.x, a1[0]);
                    a1[1] = fmaf(wv.y, rv.y, a1[1]);
                    a1[2] = fmaf(wv.y, rv.z, a1[2]);
                    a1[3] = fmaf(wv.y, rv.w, a1[3]);
                }
                *reinterpret_cast<float4 *>(&part[(w * CD + 2 * g4) * TF + 4 * l4]) = make_float4(a0[0], a0[1], a0[2], a0[3]);
                *reinterpret_cast<float4 *>(&part[(w * CD + 2 * g4 + 1) * TF + 4 * l4]) = make_float4(a1[0], a1[1], a1[2], a1[3]);
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
            {
                const float *CB = acquire();
                const float *c2 = CB + K * CD;
                float ex0[CD], ex1[CD];
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float2 t = *reinterpret_cast<const float2 *>(&es[k * TF + f0]);
                    ex0[k] = t.x;
                    ex1[k] = t.y;
                }
                const float2 e2v = *reinterpret_cast<const float2 *>(&e2s[f0]);
                float best0 = __int_as_float(0x7f800000), best1 = best0;
                int bi0 = 0, bi1 = 0;
#pragma unroll 4
                for (int i = 0; i < K / 32; ++i) {
                    const int j = 32 * i + 2 * w + g2;  // ascending per thread: strict '<' keeps the first minimum
                    const float4 ca = *reinterpret_cast<const float4 *>(CB + j * CD);
                    const float4 cb = *reinterpret_cast<const float4 *>(CB + j * CD + 4);
                    const float cc = c2[j];
                    float d0 = __fmul_rn(ex0[0], ca.x), d1 = __fmul_rn(ex1[0], ca.x);
                    d0 = __fmaf_rn(ex0[1], ca.y, d0); d1 = __fmaf_rn(ex1[1], ca.y, d1);
                    d0 = __fmaf_rn(ex0[2], ca.z, d0); d1 = __fmaf_rn(ex1[2], ca.z, d1);
                    d0 = __fmaf_rn(ex0[3], ca.w, d0); d1 = __fmaf_rn(ex1[3], ca.w, d1);
                    d0 = __fmaf_rn(ex0[4], cb.x, d0); d1 = __fmaf_rn(ex1[4], cb.x, d1);
                    d0 = __fmaf_rn(ex0[5], cb.y, d0); d1 = __fmaf_rn(ex1[5], cb.y, d1);
                    d0 = __fmaf_rn(ex0[6], cb.z, d0); d1 = __fmaf_rn(ex1[6], cb.z, d1);
                    d0 = __fmaf_rn(ex0[7], cb.w, d0); d1 = __fmaf_rn(ex1[7], cb.w, d1);
                    const float dist0 = __fadd_rn(__fsub_rn(e2v.x, d0), cc);
                    const float dist1 = __fadd_rn(__fsub_rn(e2v.y, d1), cc);
                    if (dist0 < best0) { best0 = dist0; bi0 = j; }
                    if (dist1 < best1) { best1 = dist1; bi1 = j; }
                }
                {   // merge the two code groups of the warp (lane ^ 16)
                    const float ob0 = __shfl_xor_sync(0xffffffffu, best0, 16), ob1 = __shfl_xor_sync(0xffffffffu, best1, 16);
                    const int oi0 = __shfl_xor_sync(0xffffffffu, bi0, 16), oi1 = __shfl_xor_sync(0xffffffffu, bi1, 16);
                    if (ob0 < best0 || (ob0 == best0 && oi0 < bi0)) { best0 = ob0; bi0 = oi0; }
                    if (ob1 < best1 || (ob1 == best1 && oi1 < bi1)) { best1 = ob1; bi1 = oi1; }
                }
                if (g2 == 0) {
                    *reinterpret_cast<float2 *>(&sbest[w * TF + f0]) = make_float2(best0, best1);
                    *reinterpret_cast<int2 *>(&sidx[w * TF + f0]) = make_int2(bi0, bi1);
                }
                __syncthreads();
            }
            // ================= argmin merge, gather, loss, straight-through (one warp) =================
            if (w == 0) {
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
                const float cr[CD] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                float ls = 0.0f;
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float x = ze[k * TF + lane];
                    const float diff = __fsub_rn(x, cr[k]);  // quantize.py:69-71
                    const float sq = __fmul_rn(diff, diff);
                    ls = (k == 0) ? sq : __fadd_rn(ls, sq);
                    qs[k * TF + lane] = __fadd_rn(x, __fsub_rn(cr[k], x));  // quantize.py:73-75
                }
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
            __syncthreads();
            // ================= out_proj + residual update + masked accumulate =================
            {
                const float *WO = acquire();
                const float *bo = WO + D * CD;
                float q0[CD], q1[CD];
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float2 t = *reinterpret_cast<const float2 *>(&qs[k * TF + f0]);
                    q0[k] = t.x;
                    q1[k] = t.y;
                }
                const int2 nk = *reinterpret_cast<const int2 *>(&nkeep[f0]);
                const float m0 = nk.x > s ? 1.0f : 0.0f, m1 = nk.y > s ? 1.0f : 0.0f;
                const bool v0ok = f0 < fv, v1ok = f0 + 1 < fv;
                const int dbase = w * DW + g2;
                float *zis = nullptr;
                if (p.z_q_is != nullptr)
                    zis = p.z_q_is + (long long)b * p.zqis_sb + (long long)s * p.zqis_sq + (long long)dbase * p.zqis_sd + t0 + f0;
#pragma unroll
                for (int i = 0; i < DW / 2; ++i) {
                    const int d = dbase + 2 * i;
                    const float4 wa = *reinterpret_cast<const float4 *>(WO + d * CD);
                    const float4 wb = *reinterpret_cast<const float4 *>(WO + d * CD + 4);
                    float v0 = bo[d], v1 = v0;
                    v0 = __fmaf_rn(wa.x, q0[0], v0); v1 = __fmaf_rn(wa.x, q1[0], v1);
                    v0 = __fmaf_rn(wa.y, q0[1], v0); v1 = __fmaf_rn(wa.y, q1[1], v1);
                    v0 = __fmaf_rn(wa.z, q0[2], v0); v1 = __fmaf_rn(wa.z, q1[2], v1);
                    v0 = __fmaf_rn(wa.w, q0[3], v0); v1 = __fmaf_rn(wa.w, q1[3], v1);
                    v0 = __fmaf_rn(wb.x, q0[4], v0); v1 = __fmaf_rn(wb.x, q1[4], v1);
                    v0 = __fmaf_rn(wb.y, q0[5], v0); v1 = __fmaf_rn(wb.y, q1[5], v1);
                    v0 = __fmaf_rn(wb.z, q0[6], v0); v1 = __fmaf_rn(wb.z, q1[6], v1);
                    v0 = __fmaf_rn(wb.w, q0[7], v0); v1 = __fmaf_rn(wb.w, q1[7], v1);
                    if (!last) {
                        float2 *rp = reinterpret_cast<float2 *>(&R[d * TF + f0]);
                        float2 r = *rp;
                        r.x = __fsub_rn(r.x, v0);
                        r.y = __fsub_rn(r.y, v1);
                        *rp = r;
                    }
                    acc[i][0] = __fmaf_rn(m0, v0, acc[i][0]);
                    acc[i][1] = __fmaf_rn(m1, v1, acc[i][1]);
                    if (zis != nullptr) {
                        float *o = zis + (long long)(2 * i) * p.zqis_sd;
                        if (p.vec_st == 2) {
                            if (v0ok) st_cs2(o, v0, v1);  // fv is even whenever vec_st == 2
                        } else {
                            if (v0ok) st_cs(o, v0);
                            if (v1ok) st_cs(o + 1, v1);
                        }
                    }
                }
                __syncthreads();  // residual updated; qs and the weight slot are free again
            }
        }  // stages

        if (p.z_q != nullptr) {
            const bool v0ok = f0 < fv, v1ok = f0 + 1 < fv;
            float *zo = p.z_q + (long long)b * p.zq_sb + (long long)(w * DW + g2) * p.zq_sd + t0 + f0;
#pragma unroll
            for (int i = 0; i < DW / 2; ++i) {
                float *o = zo + (long long)(2 * i) * p.zq_sd;
                if (p.vec_st == 2) {
                    if (v0ok) st_cs2(o, acc[i][0], acc[i][1]);
                } else {
                    if (v0ok) st_cs(o, acc[i][0]);
                    if (v1ok) st_cs(o + 1, acc[i][1]);
                }
            }
        }
    }  // tiles

    if (w == 0) {
        if (lane == 0 && p.loss_sum != nullptr) atomicAdd(p.loss_sum, loss_acc);
        if (lane < n_run && p.kept != nullptr && kept_acc != 0ull) atomicAdd(&p.kept[lane], kept_acc);
    }
}

// ---- host launcher ---------------------------------------------------------------------------
template <int D, int K>
static int launch_encode(const EncodeParams &p, int grid, cudaStream_t stream) {
    using S = EncodeSmem<D, K>;
    static bool attr_done = false;  // benign race: the attribute is idempotent
    if (!attr_done) {
        int rc = check_cuda(cudaFuncSetAttribute(rvq_encode_kernel<D, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES),
                            "cudaFuncSetAttribute(rvq_encode_kernel)");
        if (rc) return rc;
        attr_done = true;
    }
    rvq_encode_kernel<D, K><<<grid, NT, S::BYTES, stream>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_encode_kernel launch");
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
