// Fused residual-vector-quantisation encode for B200 (sm_100a).
//
// One persistent CTA per SM walks tiles of TF=32 consecutive latent frames of one batch item.
// Thread (warp w, quarter-warp g4, l4) owns the 4 frames 4*l4..4*l4+3 of the D/64 channels
// d = w*D/16 + 4i + g4 for the whole tile.  Its slice of the residual and of the z_q accumulators is
// thread-private, so both live in TENSOR MEMORY: 2 x D/16 columns of the thread's own TMEM lane, moved
// with tcgen05.ld / tcgen05.st (~800 B/clk/SM measured) instead of through shared memory (64 B/clk for
// lane-distinct LDS.128, profiles/r1_micro_tmem_lds.txt) or the register file (which stays free for
// pipelined operands).  Shared memory holds only: the staging buffer the next tile's latent is
// prefetched into (cp.async), the two-slot ring of per-stage weight pieces (cp.async.bulk + mbarrier,
// one piece ahead of the math) and small per-stage exchange buffers.  The latent is read from HBM
// exactly once; codes, latents, mask, z_q and (optionally) z_q_is are written exactly once.
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
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "encode_params.cuh"

namespace vrvq {

constexpr int TF = 32;   // frames per tile
constexpr int NT = 512;  // threads per CTA
constexpr int NW = NT / 32;


template <int D, int K>
struct EncodeSmem {
    static constexpr BlobLayout L = BlobLayout(D, K);
    static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    static constexpr int R_FLOATS = D * TF;  // staging buffer for the next tile's latent
    static constexpr int WB_FLOATS = (cmax(cmax(L.p0_floats(), L.p1_floats()), L.p2_floats()) + 3) / 4 * 4;
    static constexpr int PART_ROW = 10;  // 8 partial sums per (warp, frame), padded to 10 floats: 8-byte aligned pairs, <=2-way banks
    static constexpr int PART_FLOATS = NW * TF * PART_ROW;
    static constexpr int OFF_R = 0;
    static constexpr int OFF_WB0 = OFF_R + R_FLOATS;
    static constexpr int OFF_WB1 = OFF_WB0 + WB_FLOATS;
    static constexpr int OFF_PART = OFF_WB1 + WB_FLOATS;
    static constexpr int OFF_ZE = OFF_PART + PART_FLOATS;
    static constexpr int OFF_ES = OFF_ZE + CD * TF;
    static constexpr int OFF_E2 = OFF_ES + CD * TF;
    static constexpr int OFF_QS = OFF_E2 + TF;  // [CD][TF] straight-through vectors
    static constexpr int OFF_NKEEP = OFF_QS + CD * TF;
    static constexpr int OFF_BARS = OFF_NKEEP + TF;  // 2 x uint64
    static constexpr int OFF_TMEM = OFF_BARS + 4;    // TMEM base address written by tcgen05.alloc
    static constexpr int TOTAL_FLOATS = OFF_TMEM + 4;
    static constexpr int BYTES = TOTAL_FLOATS * 4;
    static_assert(OFF_BARS % 2 == 0, "mbarriers need 8-byte alignment");
    static_assert(OFF_WB0 % 4 == 0 && OFF_WB1 % 4 == 0, "bulk copy destinations need 16-byte alignment");
    static_assert(BYTES <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
};

// global -> shared prefetch of one [D x TF] latent tile; frames >= fv are zero-filled.
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

// Which 4 of the tile's 32 frames a thread owns depends on the widest store every output row allows, so that each
// warp-level store instruction writes contiguous bytes (full 32-byte sectors) whatever the row alignment is:
//   V = 4 (rows 16-byte aligned): frames 4*l4 + j                   -> one STG.128, 8 lanes x 16 B contiguous
//   V = 2 (rows  8-byte aligned): frames 2*l4 + (j&1) + 16*(j>>1)   -> two STG.64, each 8 lanes x 8 B contiguous
//   V = 1 (rows  4-byte aligned): frames l4 + 8*j                   -> four STG.32, each 8 lanes x 4 B contiguous
// (Measured on B200: with frames 4*l4+j and split stores, T=862 ran 25% slower than T=864; profiles/r1d_*.)
template <int V>
__device__ __forceinline__ int frame_of(int l4, int j) {
    return V == 4 ? 4 * l4 + j : V == 2 ? 2 * l4 + (j & 1) + 16 * (j >> 1) : l4 + 8 * j;
}

// The thread's 4 frames of one 32-frame row in shared memory (row = pointer to frame 0).
template <int V>
__device__ __forceinline__ void row_load4(const float *row, int l4, float (&x)[4]) {
    if (V == 4) {
        const float4 t = *reinterpret_cast<const float4 *>(row + 4 * l4);
        x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
    } else if (V == 2) {
        const float2 a = *reinterpret_cast<const float2 *>(row + 2 * l4);
        const float2 b = *reinterpret_cast<const float2 *>(row + 16 + 2 * l4);
        x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = row[l4 + 8 * j];
    }
}
template <int V>
__device__ __forceinline__ void row_store4(float *row, int l4, float a, float b, float c, float d) {
    if (V == 4) {
        *reinterpret_cast<float4 *>(row + 4 * l4) = make_float4(a, b, c, d);
    } else if (V == 2) {
        *reinterpret_cast<float2 *>(row + 2 * l4) = make_float2(a, b);
        *reinterpret_cast<float2 *>(row + 16 + 2 * l4) = make_float2(c, d);
    } else {
        row[l4] = a; row[l4 + 8] = b; row[l4 + 16] = c; row[l4 + 24] = d;
    }
}
// The same 4 frames to global memory (write-once, streaming); o points at frame 0 of the row, fv = valid frames.
template <int V>
__device__ __forceinline__ void store4(float *o, int l4, float a, float b, float c, float d, int fv) {
    if (V == 4) {
        if (4 * l4 < fv) st_cs4(o + 4 * l4, make_float4(a, b, c, d));  // fv is a multiple of 4 whenever V == 4
    } else if (V == 2) {
        if (2 * l4 < fv) st_cs2(o + 2 * l4, a, b);  // fv is even whenever V == 2
        if (16 + 2 * l4 < fv) st_cs2(o + 16 + 2 * l4, c, d);
    } else {
        if (l4 < fv) st_cs(o + l4, a);
        if (l4 + 8 < fv) st_cs(o + l4 + 8, b);
        if (l4 + 16 < fv) st_cs(o + l4 + 16, c);
        if (l4 + 24 < fv) st_cs(o + l4 + 24, d);
    }
}

// One 1x8 by 8x4 FMA block: a[c][f] += w[c] * r[f], as 16 packed fma.rn.f32x2 over adjacent out-channel pairs
// (the weight pairs come straight from the LDS.128; the residual value is duplicated).  a2[cp][f] = (a[2cp][f], a[2cp+1][f]).
__device__ __forceinline__ void fma_8x4(float2 (&a2)[CD / 2][4], const float4 &wa, const float4 &wb, float r0, float r1, float r2, float r3) {
    const float2 w2[CD / 2] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
    const float2 rr[4] = {dup2(r0), dup2(r1), dup2(r2), dup2(r3)};
#pragma unroll
    for (int cp = 0; cp < CD / 2; ++cp)
#pragma unroll
        for (int f = 0; f < 4; ++f) a2[cp][f] = __ffma2_rn(w2[cp], rr[f], a2[cp][f]);
}

// out_proj + residual update + masked accumulate for one thread: NCH channels (stride 4 in d) x 4 frames.
// LAST and the store shapes are compile-time, so the body is branch-free.  Residual and z_q accumulators stream through
// TMEM in groups of 8 columns (2 channels x 4 frames); the accumulators were zeroed when the tile was parked in TMEM.
template <int NCH, bool LAST, bool ZQIS, int VEC_ST>
__device__ __forceinline__ void out_proj_thread(const float *__restrict__ wp, const float *__restrict__ bp, const float (&q)[CD][4],
                                                const float (&m)[4], uint32_t tR, uint32_t tA, float *zo, long long zstep, float *zq,
                                                long long zqstep, int l4, int fv) {
    static_assert(NCH % 2 == 0, "channels per thread must be even");
#pragma unroll 2
    for (int g = 0; g < NCH / 2; ++g) {
        uint32_t r8[8], a8[8];
        if (!LAST) tmem_ld8(tR + 8 * g, r8);
        tmem_ld8(tA + 8 * g, a8);
        float v[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = 2 * g + u;
            const float4 wa = *reinterpret_cast<const float4 *>(wp + i * 4 * CD);
            const float4 wb = *reinterpret_cast<const float4 *>(wp + i * 4 * CD + 4);
            const float wv[CD] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            const float bias = bp[4 * i];
            float x[4] = {bias, bias, bias, bias};
#pragma unroll
            for (int k = 0; k < CD; ++k) {  // bias-initialised ascending-k chain, 4 independent frames
                x[0] = __fmaf_rn(wv[k], q[k][0], x[0]);
                x[1] = __fmaf_rn(wv[k], q[k][1], x[1]);
                x[2] = __fmaf_rn(wv[k], q[k][2], x[2]);
                x[3] = __fmaf_rn(wv[k], q[k][3], x[3]);
            }
            if (ZQIS) {
                store4<VEC_ST>(zo, l4, x[0], x[1], x[2], x[3], fv);
                zo += zstep;
            }
            v[4 * u] = x[0]; v[4 * u + 1] = x[1]; v[4 * u + 2] = x[2]; v[4 * u + 3] = x[3];
        }
        if (!LAST) tmem_wait_ld(r8);
        tmem_wait_ld(a8);
        if (!LAST) {  // r <- r - z_q_i  (quantize.py:195 / :360)
#pragma unroll
            for (int e = 0; e < 8; ++e) r8[e] = __float_as_uint(__fsub_rn(__uint_as_float(r8[e]), v[e]));
            tmem_st8(tR + 8 * g, r8);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {  // z_q += mask * z_q_i  (quantize.py:194 / :421), ascending stage order from 0
            a8[e] = __float_as_uint(__fmaf_rn(m[e & 3], v[e], __uint_as_float(a8[e])));
        }
        if (!LAST) {
            tmem_st8(tA + 8 * g, a8);
        } else if (zq != nullptr) {
            store4<VEC_ST>(zq, l4, __uint_as_float(a8[0]), __uint_as_float(a8[1]), __uint_as_float(a8[2]), __uint_as_float(a8[3]), fv);
            store4<VEC_ST>(zq + zqstep, l4, __uint_as_float(a8[4]), __uint_as_float(a8[5]), __uint_as_float(a8[6]), __uint_as_float(a8[7]), fv);
            zq += 2 * zqstep;
        }
    }
    if (!LAST) tmem_wait_st();
}

template <int D, int K, int VEC_ST, bool ZQIS>
__global__ void __launch_bounds__(NT, 1) rvq_encode_kernel(const EncodeParams p) {
    using S = EncodeSmem<D, K>;
    constexpr BlobLayout L = BlobLayout(D, K);
    constexpr int DW = D / NW;   // channels per warp
    constexpr int NCH = DW / 4;  // channels per thread
    static_assert(D % (NW * 8) == 0, "D must be a multiple of 128");
    static_assert(K % 64 == 0, "codebook size must be a multiple of 64");

    extern __shared__ __align__(128) float smem[];
    float *R = smem + S::OFF_R;  // staging: the tile being prefetched
    float *wb0 = smem + S::OFF_WB0;
    float *wb1 = smem + S::OFF_WB1;
    constexpr int PR = S::PART_ROW;
    float *part = smem + S::OFF_PART;  // in_proj partial sums [NW][TF][PR]; re-used as sbest/sidx after the reduce
    float *sbest = part;               // [NW][TF]
    int *sidx = reinterpret_cast<int *>(part + NW * TF);
    float *ze = smem + S::OFF_ZE;  // [CD][TF] pre-normalisation latents
    float *es = smem + S::OFF_ES;  // [CD][TF] 2*e
    float *e2s = smem + S::OFF_E2;
    float *qs = smem + S::OFF_QS;
    int *nkeep = reinterpret_cast<int *>(smem + S::OFF_NKEEP);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::OFF_BARS);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + S::OFF_TMEM);
    // per thread: NCH*4 residual columns + NCH*4 accumulator columns; the 4 warps of a lane quadrant sit side by side
    constexpr uint32_t TMEM_COLS = (NW / 4) * 2 * DW;
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
        // No proxy fence here: the slot was only READ through the generic proxy, and every reader has passed a __syncthreads
        // before this point (the usual consumer-release -> TMA-refill hand-over).  A fence.proxy.async would also wait for
        // thread 0's outstanding global loads (measured: ~600 cycles per stage behind the raw-row gather).
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s((n & 1u) ? wb1 : wb0, src, bytes, bar);
    };
    // all threads, at the start of every phase: wait for this phase's piece; `prefetch` also issues the next piece into the
    // other slot, which requires that every thread is done with that slot (i.e. has passed a barrier since its last read).
    auto acquire = [&](bool prefetch) -> const float * {
        if (prefetch && tid == 0 && piece + 1 < n_pieces) issue_piece(piece + 1);
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
    // this thread's strip: its own lane of quadrant w%4; [residual DW cols | accumulators DW cols] at (w/4)*2*DW
    const uint32_t tR = tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * 2 * DW);
    const uint32_t tA = tR + DW;
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

    // thread coordinates
    const int l4 = lane & 7, g4 = lane >> 3;   // frames frame_of<VEC_ST>(l4, 0..3); quarter-warp g4 -> channels w*DW + 4i + g4
    const int l2 = lane & 15, g2 = lane >> 4;  // search: frames 2*l2, 2*l2+1; half-warp g2 -> code pairs
    const int dbase = w * DW + g4;

    long long ph_last = 0, ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // thread 0 only, when p.phase_cycles != NULL
    const bool ph_on = (p.phase_cycles != nullptr) && tid == 0;
    if (ph_on) ph_last = clock64();
    auto ph_mark = [&](int k) {
        if (ph_on) {
            const long long t = clock64();
            ph_acc[k] += t - ph_last;
            ph_last = t;
        }
    };
    double loss_acc = 0.0;               // lanes 0 and 16 of every warp (the frames they finalise)
    unsigned long long kept_acc = 0ull;  // lane k of warp 0 counts stage k

    // reduce-scatter of the in_proj partial sums over the four K sub-slices (quarter-warps) of a warp, then to `part`
    auto scatter_partials = [&](float (&a)[CD][4]) {
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
#pragma unroll
        for (int f = 0; f < 4; ++f)
            *reinterpret_cast<float2 *>(&part[(w * TF + frame_of<VEC_ST>(l4, f)) * PR + c0]) = make_float2(q2[0][f], q2[1][f]);
    };

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
        __syncthreads();  // staged latent tile and nkeep visible
        ph_mark(0);  // tile setup + wait for the prefetched latent

        for (int s = 0; s < n_run; ++s) {
            const bool last = (s == n_run - 1);
            // ================= in_proj: z_e[c][f] = sum_d W_in[c][d] r[d][f] =================
            // Thread tile 8 channels-out x 4 frames over its own NCH residual channels (K split 4 ways inside the warp,
            // NW ways across warps).  Stage 0 reads the staged latent and moves it into TMEM on the way.
            {
                const float *W = acquire(false);  // the other slot (previous out_proj weights) may still be in use: no barrier since
                float2 a2[CD / 2][4];
#pragma unroll
                for (int cp = 0; cp < CD / 2; ++cp) a2[cp][0] = a2[cp][1] = a2[cp][2] = a2[cp][3] = make_float2(0.0f, 0.0f);
                const float *Wp = W + dbase * CD;
                if (s == 0) {
                    const float *Rp = R + dbase * TF;
                    const uint32_t zero8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll 2
                    for (int g = 0; g < NCH / 2; ++g) {
                        uint32_t r8[8];
                        tmem_st8(tA + 8 * g, zero8);  // z_q accumulators start at +0 (quantize.py:165,335)
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = 2 * g + u;
                            float rv[4];
                            row_load4<VEC_ST>(Rp + i * 4 * TF, l4, rv);
                            const float4 wa = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD);
                            const float4 wb = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD + 4);
                            fma_8x4(a2, wa, wb, rv[0], rv[1], rv[2], rv[3]);
                            r8[4 * u] = __float_as_uint(rv[0]); r8[4 * u + 1] = __float_as_uint(rv[1]);
                            r8[4 * u + 2] = __float_as_uint(rv[2]); r8[4 * u + 3] = __float_as_uint(rv[3]);
                        }
                        if (n_run > 1) tmem_st8(tR + 8 * g, r8);
                    }
                    tmem_wait_st();
                } else {
                    uint32_t r8[2][8];
                    tmem_ld8(tR, r8[0]);
#pragma unroll 2
                    for (int g = 0; g < NCH / 2; ++g) {
                        tmem_wait_ld(r8[g & 1]);
                        if (g + 1 < NCH / 2) tmem_ld8(tR + 8 * (g + 1), r8[(g + 1) & 1]);
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int i = 2 * g + u;
                            const float4 wa = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD);
                            const float4 wb = *reinterpret_cast<const float4 *>(Wp + i * 4 * CD + 4);
                            fma_8x4(a2, wa, wb, __uint_as_float(r8[g & 1][4 * u]), __uint_as_float(r8[g & 1][4 * u + 1]),
                                    __uint_as_float(r8[g & 1][4 * u + 2]), __uint_as_float(r8[g & 1][4 * u + 3]));
                        }
                    }
                }
                float a[CD][4];
#pragma unroll
                for (int cp = 0; cp < CD / 2; ++cp)
#pragma unroll
                    for (int f = 0; f < 4; ++f) {
                        a[2 * cp][f] = a2[cp][f].x;
                        a[2 * cp + 1][f] = a2[cp][f].y;
                    }
                scatter_partials(a);
                __syncthreads();
                if (tid == 0 && piece < n_pieces) issue_piece(piece);  // search piece; every thread has left the previous out_proj
                ph_mark(1);  // in_proj (incl. weight wait)
                // the staging buffer is free once stage 0 has consumed it: prefetch the next tile for the rest of this one
                if (s == 0 && it + 1 < n_my_tiles) start_tile_load(it + 1);
                // cross-warp reduce + normalise in one phase: thread = (frame tid>>3, channel tid&7); the 8 channels of a frame sit
                // in 8 adjacent lanes and are exchanged with shuffles (quantize.py:92 in torch's exact op order)
                if (tid < CD * TF) {
                    const int f = tid >> 3, c = tid & 7, lb = lane & ~7;
                    float sacc = part[f * PR + c];
#pragma unroll
                    for (int ww = 1; ww < NW; ++ww) sacc = __fadd_rn(sacc, part[(ww * TF + f) * PR + c]);
                    const float xc = __fadd_rn(sacc, W[D * CD + c]);
                    float ss = 0.0f;
#pragma unroll
                    for (int k = 0; k < CD; ++k) {
                        const float xk = __shfl_sync(0xffffffffu, xc, lb + k);
                        const float sq = __fmul_rn(xk, xk);
                        ss = (k == 0) ? sq : __fadd_rn(ss, sq);
                    }
                    const float den = fmaxf(__fsqrt_rn(ss), 1e-12f);
                    const float ec = __fdiv_rn(xc, den);
                    const float sqe = __fmul_rn(ec, ec);
                    float e2 = __shfl_sync(0xffffffffu, sqe, lb);
#pragma unroll
                    for (int k = 1; k < CD; ++k) e2 = __fadd_rn(e2, __shfl_sync(0xffffffffu, sqe, lb + k));
                    ze[c * TF + f] = xc;
                    es[c * TF + f] = __fmul_rn(2.0f, ec);
                    if (c == 0) e2s[f] = e2;
                }
                __syncthreads();
                // latents of this stage (pre-normalisation z_e), coalesced: warp c writes channel c
                if (w < CD && p.latents != nullptr && lane < fv)
                    p.latents[(long long)b * p.lat_sb + (long long)(s * CD + w) * p.lat_sc + t0 + lane] = ze[w * TF + lane];
                ph_mark(2);  // cross-warp reduce + normalise
            }
            // ================= search over the normalised codebook =================
            // Thread: 2 frames x K/32 code pairs (ascending); the two half-warps take adjacent pairs, so every codebook
            // LDS.128 is a 2-address broadcast.  The piece is pair-interleaved ([pair][k][2]): one fma.rn.f32x2 advances
            // the dot products of two adjacent codes.
            {
                const float *CB = acquire(true);
                const float *c2 = CB + K * CD;
                const int f0 = 2 * l2;
                float2 ea[CD], eb[CD];  // (2e_k, 2e_k) of frame f0 / f0+1
#pragma unroll
                for (int k = 0; k < CD; ++k) {
                    const float2 t = *reinterpret_cast<const float2 *>(&es[k * TF + f0]);
                    ea[k] = dup2(t.x);
                    eb[k] = dup2(t.y);
                }
                const float2 e2v = *reinterpret_cast<const float2 *>(&e2s[f0]);
                const float2 e2a = dup2(e2v.x), e2b = dup2(e2v.y);
                float best0 = __int_as_float(0x7f800000), best1 = best0;
                int bi0 = 0, bi1 = 0;
#pragma unroll 2
                for (int i = 0; i < K / 64; ++i) {
                    const int pr = 32 * i + 2 * w + g2;  // code pair; codes 2pr, 2pr+1 ascend with i
                    const float4 *cp = reinterpret_cast<const float4 *>(CB + pr * 2 * CD);
                    const float4 c01 = cp[0], c23 = cp[1], c45 = cp[2], c67 = cp[3];  // (c_j,k c_j+1,k) for k = 0..7
                    const float2 cc = *reinterpret_cast<const float2 *>(c2 + 2 * pr);
                    float2 da = __fmul2_rn(ea[0], make_float2(c01.x, c01.y));
                    float2 db = __fmul2_rn(eb[0], make_float2(c01.x, c01.y));
                    da = __ffma2_rn(ea[1], make_float2(c01.z, c01.w), da); db = __ffma2_rn(eb[1], make_float2(c01.z, c01.w), db);
                    da = __ffma2_rn(ea[2], make_float2(c23.x, c23.y), da); db = __ffma2_rn(eb[2], make_float2(c23.x, c23.y), db);
                    da = __ffma2_rn(ea[3], make_float2(c23.z, c23.w), da); db = __ffma2_rn(eb[3], make_float2(c23.z, c23.w), db);
                    da = __ffma2_rn(ea[4], make_float2(c45.x, c45.y), da); db = __ffma2_rn(eb[4], make_float2(c45.x, c45.y), db);
                    da = __ffma2_rn(ea[5], make_float2(c45.z, c45.w), da); db = __ffma2_rn(eb[5], make_float2(c45.z, c45.w), db);
                    da = __ffma2_rn(ea[6], make_float2(c67.x, c67.y), da); db = __ffma2_rn(eb[6], make_float2(c67.x, c67.y), db);
                    da = __ffma2_rn(ea[7], make_float2(c67.z, c67.w), da); db = __ffma2_rn(eb[7], make_float2(c67.z, c67.w), db);
                    // dist = fl(fl(e2 - dot) + c2)
                    const float2 ta = __fadd2_rn(__fadd2_rn(e2a, make_float2(-da.x, -da.y)), cc);
                    const float2 tb = __fadd2_rn(__fadd2_rn(e2b, make_float2(-db.x, -db.y)), cc);
                    const int j = 2 * pr;
                    if (ta.x < best0) { best0 = ta.x; bi0 = j; }
                    if (ta.y < best0) { best0 = ta.y; bi0 = j + 1; }
                    if (tb.x < best1) { best1 = tb.x; bi1 = j; }
                    if (tb.y < best1) { best1 = tb.y; bi1 = j + 1; }
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
                ph_mark(3);  // search
            }
            // ===== argmin merge, gather, loss, straight-through: warp w finalises frames 2w and 2w+1 =====
            // Half-warp h handles frame f = 2w+h: lane i of the half holds warp i's minimum, a 4-step xor butterfly with the
            // first-index tie rule leaves the winner in every lane, lanes 0..7 fetch the 8 floats of the raw codebook row
            // (one 32-byte sector per frame; quantize.py:81-85,102) and form the straight-through value (quantize.py:73-75).
            const float *WO = acquire(true);  // out_proj weights (and kicks off the next stage's in_proj piece)
            {
                const int h = lane >> 4, f = 2 * w + h;
                float best = sbest[(lane & 15) * TF + f];
                int bi = sidx[(lane & 15) * TF + f];
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                const int k = lane & 7;
                const float cr = __ldg(stages + (size_t)s * L.stage_floats() + L.off_raw() + (size_t)bi * CD + k);
                const float x = ze[k * TF + f];
                const float diff = __fsub_rn(x, cr);  // quantize.py:69-71
                const float sq = __fmul_rn(diff, diff);
                if ((lane & 8) == 0) qs[k * TF + f] = __fadd_rn(x, __fsub_rn(cr, x));
                float ls = __shfl_sync(0xffffffffu, sq, lane & 16);
#pragma unroll
                for (int kk = 1; kk < CD; ++kk) ls = __fadd_rn(ls, __shfl_sync(0xffffffffu, sq, (lane & 16) + kk));
                if ((lane & 15) == 0 && f < fv) {
                    const float loss = __fdiv_rn(ls, (float)CD);
                    p.codes[(long long)b * p.codes_sb + (long long)s * p.codes_sq + t0 + f] = (long long)bi;
                    if (p.loss_pf != nullptr) p.loss_pf[(long long)b * p.loss_sb + (long long)s * p.loss_sq + t0 + f] = loss;
                    if (nkeep[f] > s) loss_acc += (double)loss;
                }
            }
            __syncthreads();
            ph_mark(4);  // argmin merge + gather + straight-through
            // ================= out_proj + residual update + masked accumulate =================
            {
                const float *bo = WO + D * CD;
                float q[CD][4], m[4];
#pragma unroll
                for (int k = 0; k < CD; ++k) {  // lane = frame read (conflict-free), then distribute with shuffles
                    const float qk = qs[k * TF + lane];
#pragma unroll
                    for (int j = 0; j < 4; ++j) q[k][j] = __shfl_sync(0xffffffffu, qk, frame_of<VEC_ST>(l4, j));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = nkeep[frame_of<VEC_ST>(l4, j)] > s ? 1.0f : 0.0f;
                float *zo = nullptr;
                if (ZQIS) zo = p.z_q_is + (long long)b * p.zqis_sb + (long long)s * p.zqis_sq + (long long)dbase * p.zqis_sd + t0;
                const long long zstep = 4 * p.zqis_sd;
                float *zq = nullptr;
                if (p.z_q != nullptr) zq = p.z_q + (long long)b * p.zq_sb + (long long)dbase * p.zq_sd + t0;
                const long long zqstep = 4 * p.zq_sd;
                const float *wp = WO + dbase * CD;
                const float *bp = bo + dbase;
                if (last)
                    out_proj_thread<NCH, true, ZQIS, VEC_ST>(wp, bp, q, m, tR, tA, zo, zstep, zq, zqstep, l4, fv);
                else
                    out_proj_thread<NCH, false, ZQIS, VEC_ST>(wp, bp, q, m, tR, tA, zo, zstep, zq, zqstep, l4, fv);
                ph_mark(6);  // out_proj, thread 0's own work
                // No barrier here: the next stage's in_proj only touches thread-private TMEM columns, its own weight slot and (after
                // its own barrier) `part`; the out_proj weight slot is refilled only after that barrier (see in_proj above).
                ph_mark(5);  // out_proj
            }
        }  // stages
        __syncthreads();  // end of tile: nkeep / qs / the exchange buffers are rewritten by the next tile's setup
    }  // tiles

    if (ph_on)
        for (int k = 0; k < 8; ++k) p.phase_cycles[(size_t)blockIdx.x * 8 + k] = ph_acc[k];
    if ((lane & 15) == 0 && p.loss_sum != nullptr && loss_acc != 0.0) atomicAdd(p.loss_sum, loss_acc);
    tmem_fence_before_sync();
    __syncthreads();
    if (w == 0) {
        tmem_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);

        if (lane < n_run && p.kept != nullptr && kept_acc != 0ull) atomicAdd(&p.kept[lane], kept_acc);
    }
}

// ---- host launcher ---------------------------------------------------------------------------
template <int D, int K, int VEC_ST, bool ZQIS>
static int launch_one(const EncodeParams &p, int grid, cudaStream_t stream) {
    using S = EncodeSmem<D, K>;
    int rc = ensure_dynamic_smem<rvq_encode_kernel<D, K, VEC_ST, ZQIS>>(S::BYTES, "cudaFuncSetAttribute(rvq_encode_kernel)");
    if (rc) return rc;
    rvq_encode_kernel<D, K, VEC_ST, ZQIS><<<grid, NT, S::BYTES, stream>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_encode_kernel launch");
}

template <int D, int K>
static int launch_encode(const EncodeParams &p, int grid, cudaStream_t stream) {
    const bool zqis = p.z_q_is != nullptr;
    if (p.vec_st == 4) return zqis ? launch_one<D, K, 4, true>(p, grid, stream) : launch_one<D, K, 4, false>(p, grid, stream);
    if (p.vec_st == 2) return zqis ? launch_one<D, K, 2, true>(p, grid, stream) : launch_one<D, K, 2, false>(p, grid, stream);
    return zqis ? launch_one<D, K, 1, true>(p, grid, stream) : launch_one<D, K, 1, false>(p, grid, stream);
}

template <int D, int K>
static int smem_bytes_of() {
    return EncodeSmem<D, K>::BYTES;
}

static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int encode_supported(int D, int K, int cd) { return cd == CD && K == 1024 && (D == 1024 || D == 512 || D == 256); }

int fill_encode_params(const vrvq_encode_args *a, EncodeParams &p) {
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
    auto st_ok = [&](int v) {  // every row start of every output is v*4-byte aligned
        bool ok = a->T % v == 0;
        if (a->z_q) ok = ok && aligned(a->z_q, 4 * v) && a->z_q_stride_b % v == 0 && a->z_q_stride_d % v == 0;
        if (a->z_q_is)
            ok = ok && aligned(a->z_q_is, 4 * v) && a->z_q_is_stride_b % v == 0 && a->z_q_is_stride_q % v == 0 && a->z_q_is_stride_d % v == 0;
        return ok;
    };
    p.vec_st = st_ok(4) ? 4 : st_ok(2) ? 2 : 1;
    if (const char *dbg = getenv("VRVQ_DEBUG_MAX_VEC_LD")) {  // profiling knob: cap the load width (never widens it)
        const int cap = atoi(dbg);
        if (cap == 1 || cap == 2) p.vec_ld = p.vec_ld < cap ? p.vec_ld : cap;
    }
    if (const char *dbg = getenv("VRVQ_DEBUG_MAX_VEC_ST")) {  // profiling knob: cap the store width (never widens it)
        const int cap = atoi(dbg);
        if (cap == 1 || cap == 2) p.vec_st = p.vec_st < cap ? p.vec_st : cap;
    }
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

// Kernel choice.  The tensor-core kernel has a fixed cost per tile of ~130 us (in_proj stream, eight serial stages) whatever the
// number of frames in it; the CUDA-core kernel finishes one wave of 32-frame tiles in ~95 us.  So calls that fit one such wave
// (B * ceil(T / 32) <= #SMs, e.g. a single 1 s .. 60 s clip) stay on the CUDA-core kernel, everything larger goes to the tensor
// cores.  VRVQ_ENCODE_IMPL=cuda / =tc force one or the other (A/B comparisons, tests).
bool prefer_tc_for_size(int B, int T) {
    const int sms = current_sm_count();  // of the current device, per call
    if (sms == 0) return true;
    const char *impl = getenv("VRVQ_ENCODE_IMPL");
    if (impl != nullptr && impl[0] == 't') return true;
    return (long long)B * ((T + 31) / 32) > sms;
}

static bool use_tc(const vrvq_encode_args *a) {
    const char *impl = getenv("VRVQ_ENCODE_IMPL");
    if (impl != nullptr && impl[0] == 'c') return false;
    return encode_tc_usable(a) != 0 && prefer_tc_for_size(a->B, a->T);
}

const char *encode_kernel_name(const vrvq_encode_args *a) {
    EncodeParams p{};
    if (fill_encode_params(a, p)) return nullptr;
    return use_tc(a) ? "tc" : "cuda";
}

int encode_launch_info(const vrvq_encode_args *a, int *grid, int *block, int *smem) {
    EncodeParams p{};
    int rc = fill_encode_params(a, p);
    if (rc) return rc;
    if (use_tc(a)) return encode_tc_launch_info(a, grid, block, smem);
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
    int rc = fill_encode_params(a, p);
    if (rc) return rc;
    if (p.n_tiles == 0) return VRVQ_OK;
    rc = check_device();
    if (rc) return rc;
    if (use_tc(a)) return encode_tc(a, p, stream);
    int grid = 0;
    rc = pick_grid(p, &grid);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool dbg = getenv("VRVQ_DEBUG_PHASES") != nullptr;  // profiling only: synchronises and prints per-phase cycles
    if (dbg) {
        if (cudaMalloc(&p.phase_cycles, sizeof(long long) * 8 * (size_t)grid) != cudaSuccess) p.phase_cycles = nullptr;
    }
    switch (a->input_dim) {
        case 1024: rc = launch_encode<1024, 1024>(p, grid, st); break;
        case 512: rc = launch_encode<512, 1024>(p, grid, st); break;
        case 256: rc = launch_encode<256, 1024>(p, grid, st); break;
        default: rc = VRVQ_EUNSUPPORTED;
    }
    if (dbg && p.phase_cycles != nullptr) {
        static const char *names[8] = {"tile_setup", "in_proj", "reduce_norm", "search", "merge_gather", "out_proj_barrier", "out_proj_work", "-"};
        long long *h = static_cast<long long *>(malloc(sizeof(long long) * 8 * (size_t)grid));
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.phase_cycles, sizeof(long long) * 8 * (size_t)grid, cudaMemcpyDeviceToHost);
        double tot = 0, acc[8] = {0};
        for (int g = 0; g < grid; ++g)
            for (int k = 0; k < 8; ++k) { acc[k] += (double)h[g * 8 + k] / grid; tot += (double)h[g * 8 + k] / grid; }
        fprintf(stderr, "[vrvq phases] mean cycles per CTA: total %.0f |", tot);
        for (int k = 0; k < 7; ++k) fprintf(stderr, " %s %.0f (%.1f%%)", names[k], acc[k], 100.0 * acc[k] / tot);
        fprintf(stderr, "\n");
        free(h);
        cudaFree(p.phase_cycles);
    }
    return rc;
}

}  // namespace vrvq
