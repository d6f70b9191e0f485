// Companion kernels of the fused encode: decode-from-codes, mask utilities, level-sweep re-mask.
// All are streaming, HBM-bound kernels: coalesced along T, grid sized in multiples of the SM count.
#include "common.cuh"
#include "encode_params.cuh"

namespace vrvq {

// ---------------------------------------------------------------------------------------------
// generate_mask_hard (models/utils.py:55-61): mask[b,k,t] = (x[b,t] - k >= 0)
// ---------------------------------------------------------------------------------------------
__global__ void mask_hard_kernel(const float *__restrict__ x, long long x_sb, int B, int T, int nq, float *__restrict__ mask,
                                 long long m_sb, long long m_sq) {
    const long long total = (long long)B * nq * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % T);
        const long long r = i / T;
        const int k = (int)(r % nq);
        const int b = (int)(r / nq);
        const float xm = __fsub_rn(x[(long long)b * x_sb + t], (float)k);
        mask[(long long)b * m_sb + (long long)k * m_sq + t] = (xm >= 0.0f) ? 1.0f : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// numerator of cal_bpf_from_mask (models/utils.py:64-73): sums[k] += sum_{b,t} mask[b,k,t]  (binary64)
// grid = (blocks_per_row, nq, B)
// ---------------------------------------------------------------------------------------------
__global__ void mask_sum_kernel(const float *__restrict__ mask, long long m_sb, long long m_sq, int T, double *__restrict__ sums) {
    const int k = blockIdx.y, b = blockIdx.z;
    const float *row = mask + (long long)b * m_sb + (long long)k * m_sq;
    double acc = 0.0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) acc += (double)row[t];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    __shared__ double wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = (lane < (int)(blockDim.x >> 5)) ? wsum[lane] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0 && acc != 0.0) atomicAdd(&sums[k], acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Level-sweep re-mask (scripts/inference.py:95-100): z_q[b,d,t] = sum_k z_q_is[b,k,d,t] * mask[b,k,t]
// with mask from imp_map * level_scaled.  One thread per (b,d,t); the Nq reads per output are each
// coalesced along t.  grid.x covers t in chunks, grid.y = d-blocks, grid.z = b.
// ---------------------------------------------------------------------------------------------
constexpr int RM_DPB = 8;  // channels per CTA row-group (one warp each)
// VEC consecutive frames per lane (widest access every row start allows), one warp per channel row, Nq independent
// streaming loads in flight per lane; the mask / kept outputs are produced by the warps of the blockIdx.y == 0 row-group.
template <int VEC>
__global__ void __launch_bounds__(256) remask_kernel(const float *__restrict__ zis, long long s_b, long long s_q, long long s_d,
                                                     const float *__restrict__ imp, long long imp_sb, float level_scaled, int D, int T,
                                                     int nq, float *__restrict__ zq, long long zq_sb, long long zq_sd,
                                                     float *__restrict__ mask, long long m_sb, long long m_sq,
                                                     unsigned long long *__restrict__ kept) {
    const int b = blockIdx.z;
    const int lane = threadIdx.x & 31, dr = threadIdx.x >> 5;
    const int t = (blockIdx.x * 32 + lane) * VEC;
    int nk[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        nk[e] = 0;
        if (t + e < T) {
            const float x = __fmul_rn(imp[(long long)b * imp_sb + t + e], level_scaled);
            for (int k = 0; k < nq; ++k) nk[e] += (__fsub_rn(x, (float)k) >= 0.0f) ? 1 : 0;
        }
    }
    if (blockIdx.y == 0 && dr == 0) {  // one warp per (b, t-chunk) owns the mask / kept outputs
        for (int k = 0; k < nq; ++k) {
            int cnt = 0;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const bool on = nk[e] > k;
                cnt += on ? 1 : 0;
                if (mask != nullptr && t + e < T) mask[(long long)b * m_sb + (long long)k * m_sq + t + e] = on ? 1.0f : 0.0f;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && kept != nullptr && cnt) atomicAdd(&kept[k], (unsigned long long)cnt);
        }
    }
    if (t >= T) return;  // T % VEC == 0 whenever VEC > 1, so a lane is entirely valid or entirely out of range
    for (int d = blockIdx.y * RM_DPB + dr; d < D; d += gridDim.y * RM_DPB) {
        const float *src = zis + (long long)b * s_b + (long long)d * s_d + t;
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.0f;
#pragma unroll 4
        for (int k = 0; k < nq; ++k) {
            float v[VEC];
            if (VEC == 4) {
                const float4 q = __ldcs(reinterpret_cast<const float4 *>(src + (long long)k * s_q));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else if (VEC == 2) {
                const float2 q = __ldcs(reinterpret_cast<const float2 *>(src + (long long)k * s_q));
                v[0] = q.x; v[1] = q.y;
            } else {
                v[0] = __ldcs(src + (long long)k * s_q);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = __fmaf_rn(k < nk[e] ? 1.0f : 0.0f, v[e], acc[e]);  // mask multiply, ascending-k sum
        }
        float *o = zq + (long long)b * zq_sb + (long long)d * zq_sd + t;
        if (VEC == 4) __stcs(reinterpret_cast<float4 *>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
        else if (VEC == 2) __stcs(reinterpret_cast<float2 *>(o), make_float2(acc[0], acc[1]));
        else __stcs(o, acc[0]);
    }
}

// ---------------------------------------------------------------------------------------------
// from_codes (models/quantize.py:217-249): z_q = sum_i (W_out,i codebook_i[codes_i] + b_out,i)
// Tile = 32 frames of one batch item; 256 threads; thread (lane = frame, warp = channel slice).
// The gathered rows go to shared memory once per stage; W_out rows are read straight from L2/L1
// (8 floats per channel, warp-uniform address = one broadcast transaction).
// ---------------------------------------------------------------------------------------------
struct FromCodesParams {
    const float *blob;
    const long long *codes;
    long long c_sb, c_sq;
    const float *mask;
    long long m_sb, m_sq;
    float *z_q;
    long long zq_sb, zq_sd;
    float *z_p;
    long long zp_sb, zp_sc;
    float *z_q_is;
    long long zqis_sb, zqis_sq, zqis_sd;
    int *error_flag;
    int B, T, D, K, n_run, tiles_per_b, stage_floats, off_p2, off_raw;
};

constexpr int FC_NT = 256;
constexpr int FC_DCH = 128;  // channels handled per CTA (grid.y covers D / FC_DCH)
constexpr int FC_SG = 8;     // stages whose W_out chunk is staged in shared memory at a time
// Shared memory: gathered rows qs[n_run<=32][8][32] (padded), masks ms[32][32], W_out/b_out chunk for FC_SG stages of
// this CTA's FC_DCH channels.  Thread = 4 consecutive frames x 4 channels (dd = 16i + 4w' + g4); per stage it loads its
// 32 q values once (8 LDS.128) and then runs 4 x 32 FMAs on quarter-uniform broadcast reads of W_out.
__global__ void __launch_bounds__(FC_NT) from_codes_kernel(const FromCodesParams p) {
    extern __shared__ __align__(16) float fc_smem[];  // sized by the launcher: n_run*(256+32) + FC_SG*FC_DCH*(CD+1) floats
    float(*qs)[CD][32] = reinterpret_cast<float(*)[CD][32]>(fc_smem);
    float(*ms)[32] = reinterpret_cast<float(*)[32]>(fc_smem + p.n_run * CD * 32);
    float(*wsm)[FC_DCH][CD] = reinterpret_cast<float(*)[FC_DCH][CD]>(fc_smem + p.n_run * (CD * 32 + 32));
    float(*bsm)[FC_DCH] = reinterpret_cast<float(*)[FC_DCH]>(fc_smem + p.n_run * (CD * 32 + 32) + FC_SG * FC_DCH * CD);
    const int tile = blockIdx.x;
    const int b = tile / p.tiles_per_b, t0 = (tile % p.tiles_per_b) * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = t0 + lane;
    const bool valid = t < p.T;
    const float *stages = p.blob + BLOB_HDR_FLOATS;
    // gather: warp w handles stages w, w+8, ...
    for (int s = w; s < p.n_run; s += FC_NT / 32) {
        long long idx = valid ? p.codes[(long long)b * p.c_sb + (long long)s * p.c_sq + t] : 0;
        if (idx < 0 || idx >= p.K) {
            if (p.error_flag) atomicOr(p.error_flag, 1);
            idx = 0;
        }
        const float *raw = stages + (size_t)s * p.stage_floats + p.off_raw + (size_t)idx * CD;
        const float4 ra = __ldg(reinterpret_cast<const float4 *>(raw));
        const float4 rb = __ldg(reinterpret_cast<const float4 *>(raw + 4));
        const float cr[CD] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int k = 0; k < CD; ++k) {
            qs[s][k][lane] = cr[k];
            if (blockIdx.y == 0 && p.z_p != nullptr && valid)
                p.z_p[(long long)b * p.zp_sb + (long long)(s * CD + k) * p.zp_sc + t] = cr[k];
        }
        ms[s][lane] = (p.mask != nullptr && valid) ? p.mask[(long long)b * p.m_sb + (long long)s * p.m_sq + t] : 1.0f;
    }
    const int l4 = lane & 7, g4 = lane >> 3;
    const int tq = t0 + 4 * l4;
    const int nvalid = max(0, min(4, p.T - tq));
    const bool vec4 = (nvalid == 4) && ((p.zq_sd & 3) == 0) && ((p.zq_sb & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.z_q) & 15) == 0) &&
                      (p.z_q_is == nullptr || (((p.zqis_sd | p.zqis_sq | p.zqis_sb) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.z_q_is) & 15) == 0));
    const int d_begin = blockIdx.y * FC_DCH;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
    for (int s0 = 0; s0 < p.n_run; s0 += FC_SG) {
        const int ns = min(FC_SG, p.n_run - s0);
        __syncthreads();  // gather done (first pass) / previous chunk consumed
        for (int i = threadIdx.x; i < ns * FC_DCH * CD / 4; i += FC_NT) {  // W_out rows of this chunk, 16 bytes per thread
            const int sl = i / (FC_DCH * CD / 4), r = i % (FC_DCH * CD / 4);
            const int d = d_begin + (r * 4) / CD;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d < p.D) v = __ldg(reinterpret_cast<const float4 *>(stages + (size_t)(s0 + sl) * p.stage_floats + p.off_p2 + (size_t)d_begin * CD) + r);
            reinterpret_cast<float4 *>(&wsm[sl][0][0])[r] = v;
        }
        for (int i = threadIdx.x; i < ns * FC_DCH; i += FC_NT) {
            const int sl = i / FC_DCH, dd = i % FC_DCH;
            bsm[sl][dd] = (d_begin + dd < p.D) ? __ldg(stages + (size_t)(s0 + sl) * p.stage_floats + p.off_p2 + (size_t)p.D * CD + d_begin + dd) : 0.0f;
        }
        __syncthreads();
        for (int sl = 0; sl < ns; ++sl) {
            const int s = s0 + sl;
            float q[CD][4];
#pragma unroll
            for (int k = 0; k < CD; ++k) {
                const float4 qv = *reinterpret_cast<const float4 *>(&qs[s][k][4 * l4]);
                q[k][0] = qv.x; q[k][1] = qv.y; q[k][2] = qv.z; q[k][3] = qv.w;
            }
            const float4 mv = *reinterpret_cast<const float4 *>(&ms[s][4 * l4]);
            const float m[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int dd = 32 * i + 4 * w + g4;
                const float4 wa = *reinterpret_cast<const float4 *>(&wsm[sl][dd][0]);
                const float4 wb = *reinterpret_cast<const float4 *>(&wsm[sl][dd][4]);
                const float wv[CD] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                const float bias = bsm[sl][dd];
                float v[4] = {bias, bias, bias, bias};
#pragma unroll
                for (int k = 0; k < CD; ++k) {  // bias-initialised ascending-k chain, as in the encode kernel and the oracle
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __fmaf_rn(wv[k], q[k][j], v[j]);
                }
                const int d = d_begin + dd;
                if (p.z_q_is != nullptr && d < p.D) {
                    float *o = p.z_q_is + (long long)b * p.zqis_sb + (long long)s * p.zqis_sq + (long long)d * p.zqis_sd + tq;
                    if (vec4) {
                        __stcs(reinterpret_cast<float4 *>(o), make_float4(v[0], v[1], v[2], v[3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nvalid) __stcs(o + j, v[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(m[j], v[j], acc[i][j]);  // ascending stage order from 0
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int d = d_begin + 32 * i + 4 * w + g4;
        if (d >= p.D) continue;
        float *o = p.z_q + (long long)b * p.zq_sb + (long long)d * p.zq_sd + tq;
        if (vec4) {
            __stcs(reinterpret_cast<float4 *>(o), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nvalid) __stcs(o + j, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// search_latents (models/quantize.py:87-101 on given latents): grid = (frame tiles, stages); 256 threads;
// lane = frame, warp w scans codes [w*K/8, (w+1)*K/8) in ascending order (warp-uniform codebook reads),
// then the 8 partial minima are merged with the first-index tie rule.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) search_latents_kernel(const float *__restrict__ blob, int stage_floats, int off_p1, int K,
                                                             const float *__restrict__ lat, long long l_sb, long long l_sc, int T,
                                                             int tiles_per_b, long long *__restrict__ codes, long long c_sb,
                                                             long long c_sq) {
    __shared__ float sb[8][32];
    __shared__ int si[8][32];
    const int s = blockIdx.y;
    const int b = blockIdx.x / tiles_per_b, t0 = (blockIdx.x % tiles_per_b) * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = t0 + lane;
    const bool valid = t < T;
    float x[CD];
#pragma unroll
    for (int k = 0; k < CD; ++k) x[k] = valid ? lat[(long long)b * l_sb + (long long)(s * CD + k) * l_sc + t] : 0.0f;
    float ss = __fmul_rn(x[0], x[0]);
#pragma unroll
    for (int k = 1; k < CD; ++k) ss = __fadd_rn(ss, __fmul_rn(x[k], x[k]));
    const float den = fmaxf(__fsqrt_rn(ss), 1e-12f);
    float e2x[CD], e2 = 0.0f;
#pragma unroll
    for (int k = 0; k < CD; ++k) {
        const float e = __fdiv_rn(x[k], den);
        const float sq = __fmul_rn(e, e);
        e2 = (k == 0) ? sq : __fadd_rn(e2, sq);
        e2x[k] = __fmul_rn(2.0f, e);
    }
    const float *cbn = blob + BLOB_HDR_FLOATS + (size_t)s * stage_floats + off_p1;
    const float *c2 = cbn + (size_t)K * CD;
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    const int per = K / 8;  // codes per warp (even); the blob stores code pairs interleaved: [(j/2)][k][2]
    for (int j = w * per; j < (w + 1) * per; j += 2) {
        const float4 *cp = reinterpret_cast<const float4 *>(cbn + (size_t)(j >> 1) * 2 * CD);
        const float4 c01 = __ldg(cp), c23 = __ldg(cp + 1), c45 = __ldg(cp + 2), c67 = __ldg(cp + 3);
        float d0 = __fmul_rn(e2x[0], c01.x), d1 = __fmul_rn(e2x[0], c01.y);
        d0 = __fmaf_rn(e2x[1], c01.z, d0); d1 = __fmaf_rn(e2x[1], c01.w, d1);
        d0 = __fmaf_rn(e2x[2], c23.x, d0); d1 = __fmaf_rn(e2x[2], c23.y, d1);
        d0 = __fmaf_rn(e2x[3], c23.z, d0); d1 = __fmaf_rn(e2x[3], c23.w, d1);
        d0 = __fmaf_rn(e2x[4], c45.x, d0); d1 = __fmaf_rn(e2x[4], c45.y, d1);
        d0 = __fmaf_rn(e2x[5], c45.z, d0); d1 = __fmaf_rn(e2x[5], c45.w, d1);
        d0 = __fmaf_rn(e2x[6], c67.x, d0); d1 = __fmaf_rn(e2x[6], c67.y, d1);
        d0 = __fmaf_rn(e2x[7], c67.z, d0); d1 = __fmaf_rn(e2x[7], c67.w, d1);
        const float dist0 = __fadd_rn(__fsub_rn(e2, d0), __ldg(c2 + j));
        const float dist1 = __fadd_rn(__fsub_rn(e2, d1), __ldg(c2 + j + 1));
        if (dist0 < best) {
            best = dist0;
            bi = j;
        }
        if (dist1 < best) {
            best = dist1;
            bi = j + 1;
        }
    }
    sb[w][lane] = best;
    si[w][lane] = bi;
    __syncthreads();
    if (w == 0 && valid) {
        for (int ww = 1; ww < 8; ++ww) {
            const float ob = sb[ww][lane];
            const int oi = si[ww][lane];
            if (ob < best || (ob == best && oi < bi)) {
                best = ob;
                bi = oi;
            }
        }
        codes[(long long)b * c_sb + (long long)s * c_sq + t] = (long long)bi;
    }
}

// ---- host launchers ----------------------------------------------------------------------------
static int sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

int launch_mask_hard(const float *x, long long x_sb, int B, int T, int nq, float *mask, long long m_sb, long long m_sq, cudaStream_t st) {
    const long long total = (long long)B * nq * T;
    if (total == 0) return VRVQ_OK;
    const int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    mask_hard_kernel<<<(int)blocks, threads, 0, st>>>(x, x_sb, B, T, nq, mask, m_sb, m_sq);
    return check_cuda(cudaGetLastError(), "mask_hard_kernel launch");
}

int launch_mask_sum(const float *mask, long long m_sb, long long m_sq, int B, int T, int nq, double *sums, cudaStream_t st) {
    if ((long long)B * nq * T == 0) return VRVQ_OK;
    const int threads = 256;
    int bx = (T + threads * 4 - 1) / (threads * 4);
    if (bx < 1) bx = 1;
    if (nq > 65535 || B > 65535) {
        set_error("vrvq_mask_sum_f32: nq and B must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    mask_sum_kernel<<<dim3(bx, nq, B), threads, 0, st>>>(mask, m_sb, m_sq, T, sums);
    return check_cuda(cudaGetLastError(), "mask_sum_kernel launch");
}

int launch_remask(const float *zis, long long s_b, long long s_q, long long s_d, const float *imp, long long imp_sb, float level_scaled,
                  int B, int D, int T, int nq, float *zq, long long zq_sb, long long zq_sd, float *mask, long long m_sb, long long m_sq,
                  unsigned long long *kept, cudaStream_t st) {
    if ((long long)B * D * T == 0) return VRVQ_OK;
    if (B > 65535) {
        set_error("vrvq_remask_f32: B must be <= 65535");
        return VRVQ_EUNSUPPORTED;
    }
    auto ok = [&](int v) {
        return T % v == 0 && reinterpret_cast<uintptr_t>(zis) % (4 * v) == 0 && reinterpret_cast<uintptr_t>(zq) % (4 * v) == 0 && s_b % v == 0 &&
               s_q % v == 0 && s_d % v == 0 && zq_sb % v == 0 && zq_sd % v == 0;
    };
    const int vec = ok(4) ? 4 : ok(2) ? 2 : 1;
    const int tx = (T + 32 * vec - 1) / (32 * vec);
    int dy = (D + RM_DPB - 1) / RM_DPB;
    if (dy > 64) dy = 64;  // each warp then walks D / (8 * 64) channel rows; plenty of CTAs (tx * 64 * B) for 148 SMs
    const dim3 grid(tx, dy, B);
    if (vec == 4)
        remask_kernel<4><<<grid, 256, 0, st>>>(zis, s_b, s_q, s_d, imp, imp_sb, level_scaled, D, T, nq, zq, zq_sb, zq_sd, mask, m_sb, m_sq, kept);
    else if (vec == 2)
        remask_kernel<2><<<grid, 256, 0, st>>>(zis, s_b, s_q, s_d, imp, imp_sb, level_scaled, D, T, nq, zq, zq_sb, zq_sd, mask, m_sb, m_sq, kept);
    else
        remask_kernel<1><<<grid, 256, 0, st>>>(zis, s_b, s_q, s_d, imp, imp_sb, level_scaled, D, T, nq, zq, zq_sb, zq_sd, mask, m_sb, m_sq, kept);
    return check_cuda(cudaGetLastError(), "remask_kernel launch");
}

int launch_search_latents(const float *blob, int D, int K, const float *lat, long long l_sb, long long l_sc, int B, int T, int n_run,
                          long long *codes, long long c_sb, long long c_sq, cudaStream_t st) {
    if ((long long)B * T == 0 || n_run == 0) return VRVQ_OK;
    if (K % 16 != 0 || n_run > 65535) {
        set_error("vrvq_search_latents_f32: codebook_size must be a multiple of 16");
        return VRVQ_EUNSUPPORTED;
    }
    const BlobLayout L(D, K);
    const int tpb = (T + 31) / 32;
    const long long tiles = (long long)tpb * B;
    if (tiles > 0x7fffffffLL) {
        set_error("vrvq_search_latents_f32: too many tiles");
        return VRVQ_EUNSUPPORTED;
    }
    search_latents_kernel<<<dim3((unsigned)tiles, n_run), 256, 0, st>>>(blob, L.stage_floats(), L.off_p1(), K, lat, l_sb, l_sc, T, tpb,
                                                                        codes, c_sb, c_sq);
    return check_cuda(cudaGetLastError(), "search_latents_kernel launch");
}

int launch_from_codes(const vrvq_from_codes_args *a, cudaStream_t st) {
    if (a == nullptr || a->struct_size != sizeof(vrvq_from_codes_args)) {
        set_error("vrvq_from_codes_f32: args is NULL or struct_size mismatch");
        return VRVQ_EINVAL;
    }
    if (a->B < 0 || a->T < 0 || a->n_run < 1 || a->n_run > a->n_codebooks || a->n_run > 32) {
        set_error("vrvq_from_codes_f32: bad sizes B=%d T=%d n_run=%d n_codebooks=%d", a->B, a->T, a->n_run, a->n_codebooks);
        return a->n_run > 32 ? VRVQ_EUNSUPPORTED : VRVQ_EINVAL;
    }
    if (a->blob == nullptr || a->codes == nullptr || a->z_q == nullptr) {
        set_error("vrvq_from_codes_f32: blob, codes and z_q must be non-NULL");
        return VRVQ_EINVAL;
    }
    if (a->input_dim <= 0 || a->codebook_size <= 0) {
        set_error("vrvq_from_codes_f32: bad input_dim/codebook_size");
        return VRVQ_EINVAL;
    }
    if ((long long)a->B * a->T == 0) return VRVQ_OK;
    if (from_codes_tc_usable(a)) return from_codes_tc(a, st);  // <= 8 codebooks: the encode kernel's gather + out_proj GEMMs
    const BlobLayout L(a->input_dim, a->codebook_size);
    FromCodesParams p{};
    p.blob = static_cast<const float *>(a->blob);
    p.codes = reinterpret_cast<const long long *>(a->codes); p.c_sb = a->codes_stride_b; p.c_sq = a->codes_stride_q;
    p.mask = a->mask; p.m_sb = a->mask_stride_b; p.m_sq = a->mask_stride_q;
    p.z_q = a->z_q; p.zq_sb = a->z_q_stride_b; p.zq_sd = a->z_q_stride_d;
    p.z_p = a->z_p; p.zp_sb = a->z_p_stride_b; p.zp_sc = a->z_p_stride_c;
    p.z_q_is = a->z_q_is; p.zqis_sb = a->z_q_is_stride_b; p.zqis_sq = a->z_q_is_stride_q; p.zqis_sd = a->z_q_is_stride_d;
    p.error_flag = a->error_flag;
    p.B = a->B; p.T = a->T; p.D = a->input_dim; p.K = a->codebook_size; p.n_run = a->n_run;
    p.tiles_per_b = (a->T + 31) / 32;
    p.stage_floats = L.stage_floats(); p.off_p2 = L.off_p2(); p.off_raw = L.off_raw();
    const long long tiles = (long long)p.tiles_per_b * a->B;
    if (tiles > 0x7fffffffLL) {
        set_error("vrvq_from_codes_f32: too many tiles");
        return VRVQ_EUNSUPPORTED;
    }
    const int gy = (a->input_dim + FC_DCH - 1) / FC_DCH;
    const int smem = (int)sizeof(float) * (a->n_run * (CD * 32 + 32) + FC_SG * FC_DCH * (CD + 1));
    if (smem > 48 * 1024) {
        int rc = ensure_dynamic_smem<from_codes_kernel>(96 * 1024, "cudaFuncSetAttribute(from_codes_kernel)");
        if (rc) return rc;
    }
    from_codes_kernel<<<dim3((unsigned)tiles, gy), FC_NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "from_codes_kernel launch");
}

}  // namespace vrvq
