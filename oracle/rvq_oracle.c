/*
 * rvq_oracle.c -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * A plain-C, CPU restatement of the residual-vector-quantisation hot path of
 * lixinghe1999/VRVQ (eager PyTorch in the reference), used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg as the *checker* for the
 * sm_100a kernel in vrvq_b200/csrc/.  Nothing under vrvq_b200/ may import, link or
 * call this file.
 *
 * Parity pin: the reference ships no golden vectors or tests for this path
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the reference
 * itself, executed in the build container (tests/golden/make_golden.py writes the
 * fixtures; tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py check them).
 *
 * What each function follows (paths relative to the reference checkout):
 *   vrvq_oracle_prepare_codebook  models/quantize.py:93,99   F.normalize(codebook), codebook.pow(2).sum(1)
 *   stage_quantize (static)       models/quantize.py:42-79   VectorQuantize.forward
 *                                 models/quantize.py:87-103  VectorQuantize.decode_latents
 *   vrvq_oracle_encode            models/quantize.py:136-214 ResidualVectorQuantize.forward (eval)
 *                                 models/quantize.py:328-443 VBRResidualVectorQuantize.forward (eval)
 *                                 models/utils.py:55-61      generate_mask_hard (== forward value of generate_mask_ste)
 *   vrvq_oracle_from_codes        models/quantize.py:217-249 ResidualVectorQuantize.from_codes
 *   vrvq_oracle_from_latents      models/quantize.py:251-285 ResidualVectorQuantize.from_latents
 *   vrvq_oracle_mask_hard         models/utils.py:55-61
 *   vrvq_oracle_mask_sum          models/utils.py:64-73      cal_bpf_from_mask (numerator, per codebook)
 *
 * Arithmetic contract (each item was checked bit-for-bit against
 * torch 2.11 CPU in the build container):
 *   - weights arrive already folded: W = torch._weight_norm(v, g, 0) is computed by the
 *     caller with torch itself, so the fold is the reference's own arithmetic.
 *   - in_proj: the reference's K=D reduction order is oneDNN's and is not reproducible
 *     (it changes with batch shape and thread count); here the sum is accumulated in
 *     binary64 and rounded once to binary32, i.e. the value every fp32 order approximates.
 *   - F.normalize: n = sqrt(sequential sum of individually rounded squares), x / max(n, 1e-12).
 *   - dist_j = fl(fl(e2 - dot2_j) + c2_j), dot2_j = ascending-k FMA chain of (2 e_k) * c_jk from 0,
 *     index = first j with the smallest dist (torch: (-dist).max(1)[1]).
 *   - straight-through value q_k = fl(z_e_k + fl(c_k - z_e_k)) (not a bit-exact no-op).
 *   - out_proj: order unspecified by the reference (oneDNN); fixed here as a bias-initialised
 *     ascending-k FMA chain.  The CUDA kernel uses the same chain, so everything downstream of
 *     z_e is bit-reproducible between this file and the kernel.
 *   - residual r <- fl(r - z_q_i); z_q <- fl(z_q + m * z_q_i), ascending stage order from 0.
 *
 * Build: see oracle/build_oracle.py (gcc -O2 -ffp-contract=off -fopenmp -mfma -shared).
 * -ffp-contract=off matters: every rounding above is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VRVQ_CD 8 /* codebook_dim; the reference configs all use 8 (conf/base.yml:11) */

typedef struct {
    int n_codebooks;     /* Nq */
    int input_dim;       /* D  */
    int codebook_size;   /* K  */
    const float *w_in;   /* [Nq][8][D]   folded in_proj weight  */
    const float *b_in;   /* [Nq][8]                              */
    const float *w_out;  /* [Nq][D][8]   folded out_proj weight */
    const float *b_out;  /* [Nq][D]                              */
    const float *cb_raw; /* [Nq][K][8]   codebook.weight         */
    const float *cb_nrm; /* [Nq][K][8]   F.normalize(codebook)   */
    const float *c2;     /* [Nq][K]      sum_k cb_nrm^2          */
} vrvq_oracle_weights;

int vrvq_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void vrvq_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* models/quantize.py:92-93 -- torch.nn.functional.normalize(x, p=2, dim=1, eps=1e-12) on 8-vectors */
static inline float norm8(const float *x) {
    float s = 0.0f;
    for (int k = 0; k < VRVQ_CD; ++k) {
        float sq = x[k] * x[k];
        s = s + sq;
    }
    return sqrtf(s);
}

static inline void normalize8(const float *x, float *e) {
    float n = norm8(x);
    float den = n > 1e-12f ? n : 1e-12f;
    for (int k = 0; k < VRVQ_CD; ++k) e[k] = x[k] / den;
}

static inline float sumsq8(const float *x) {
    float s = 0.0f;
    for (int k = 0; k < VRVQ_CD; ++k) {
        float sq = x[k] * x[k];
        s = s + sq;
    }
    return s;
}

/* models/quantize.py:93,99 : codebook = F.normalize(codebook); codebook.pow(2).sum(1) */
int vrvq_oracle_prepare_codebook(const float *cb_raw, int K, float *cb_nrm, float *c2) {
    for (int j = 0; j < K; ++j) {
        normalize8(cb_raw + (size_t)j * VRVQ_CD, cb_nrm + (size_t)j * VRVQ_CD);
        c2[j] = sumsq8(cb_nrm + (size_t)j * VRVQ_CD);
    }
    return 0;
}

/* models/quantize.py:87-101 : nearest normalised codebook row of one 8-vector.
 * Returns the index; *gap receives (second smallest dist - smallest dist), used by the
 * near-tie audit in tests. */
static inline int search8(const float *z_e, const float *cb_nrm, const float *c2, int K, float *gap) {
    float e[VRVQ_CD], e2x[VRVQ_CD];
    normalize8(z_e, e);
    float e2 = sumsq8(e);
    for (int k = 0; k < VRVQ_CD; ++k) e2x[k] = 2.0f * e[k];
    float best = INFINITY, second = INFINITY;
    int idx = 0;
    for (int j = 0; j < K; ++j) {
        const float *c = cb_nrm + (size_t)j * VRVQ_CD;
        float dot = 0.0f;
        for (int k = 0; k < VRVQ_CD; ++k) dot = fmaf(e2x[k], c[k], dot);
        float d1 = e2 - dot;
        float dist = d1 + c2[j];
        if (dist < best) { /* strict: first index wins ties, like (-dist).max(1)[1] */
            second = best;
            best = dist;
            idx = j;
        } else if (dist < second) {
            second = dist;
        }
    }
    if (gap) *gap = second - best;
    return idx;
}

/* One VectorQuantize.forward on one frame (models/quantize.py:42-79).
 * r[D] is the stage input; writes z_e[8], zq[D]; returns index; *loss = mean_k (z_e - c)^2. */
static inline int stage_quantize(const vrvq_oracle_weights *w, int s, const float *r, float *z_e, float *zq,
                                 float *loss, float *gap) {
    const int D = w->input_dim, K = w->codebook_size;
    const float *w_in = w->w_in + (size_t)s * VRVQ_CD * D;
    const float *b_in = w->b_in + (size_t)s * VRVQ_CD;
    const float *w_out = w->w_out + (size_t)s * D * VRVQ_CD;
    const float *b_out = w->b_out + (size_t)s * D;
    const float *cb_raw = w->cb_raw + (size_t)s * K * VRVQ_CD;
    const float *cb_nrm = w->cb_nrm + (size_t)s * K * VRVQ_CD;
    const float *c2 = w->c2 + (size_t)s * K;

    /* in_proj (quantize.py:66): binary64 accumulation, one rounding */
    {
        double acc[VRVQ_CD];
        for (int c = 0; c < VRVQ_CD; ++c) acc[c] = 0.0;
        for (int d = 0; d < D; ++d) { /* ascending d for every c; 8 independent chains */
            const double rd = (double)r[d];
            for (int c = 0; c < VRVQ_CD; ++c) acc[c] += (double)w_in[(size_t)c * D + d] * rd;
        }
        for (int c = 0; c < VRVQ_CD; ++c) z_e[c] = (float)(acc[c] + (double)b_in[c]);
    }
    int idx = search8(z_e, cb_nrm, c2, K, gap);
    const float *cr = cb_raw + (size_t)idx * VRVQ_CD; /* quantize.py:81-85,102: raw row */
    float q[VRVQ_CD];
    float ls = 0.0f;
    for (int k = 0; k < VRVQ_CD; ++k) {
        float diff = z_e[k] - cr[k]; /* quantize.py:69-71 */
        float sq = diff * diff;
        ls = ls + sq;
        float t = cr[k] - z_e[k]; /* quantize.py:73-75 */
        q[k] = z_e[k] + t;
    }
    if (loss) *loss = ls / (float)VRVQ_CD;
    for (int d = 0; d < D; ++d) { /* quantize.py:77 */
        const float *wo = w_out + (size_t)d * VRVQ_CD;
        float acc = b_out[d];
        for (int k = 0; k < VRVQ_CD; ++k) acc = fmaf(wo[k], q[k], acc);
        zq[d] = acc;
    }
    return idx;
}

/*
 * Full eval forward of ResidualVectorQuantize / VBRResidualVectorQuantize.
 *
 *  z        [B][D][T]
 *  n_run    stages executed: Nq in VBR mode, n_quantizers in CBR mode (quantize.py:183-184,354-356)
 *  imp_map  [B][T] or NULL.  NULL selects CBR masking (mask = 1 for every executed stage).
 *  level    level[b * level_stride]; level_stride 0 broadcasts one scalar (quantize.py:389)
 *  outputs (any may be NULL except codes):
 *   codes   [B][n_run][T] int64         z_q     [B][D][T]
 *   z_q_is  [B][n_run][D][T]            latents [B][8*n_run][T]
 *   mask    [B][n_run][T]  (hard mask, utils.py:55-61; ones in CBR mode)
 *   loss_pf [B][n_run][T]  per-frame MSE (quantize.py:69-71 with loss_per_frame=True)
 *   loss_masked_sum  scalar double: sum_{b,k,t} mask * loss_pf
 *   kept    [n_run] int64: number of (b,t) with mask == 1 per stage (numerator of utils.py:64-73)
 *   gap     [B][n_run][T]  second-best minus best distance (audit aid, not a reference output)
 */
int vrvq_oracle_encode(const vrvq_oracle_weights *w, const float *z, int B, int T, int n_run,
                       const float *imp_map, const float *level, int level_stride, int64_t *codes, float *z_q,
                       float *z_q_is, float *latents, float *mask, float *loss_pf, double *loss_masked_sum,
                       int64_t *kept, float *gap) {
    const int D = w->input_dim, Nq = w->n_codebooks;
    if (n_run < 0 || n_run > Nq || B < 0 || T < 0) return -1;
    if (imp_map && !level) return -1;
    const long nframes = (long)B * T;
    double loss_total = 0.0;
    int64_t *kept_local = (int64_t *)calloc((size_t)(n_run > 0 ? n_run : 1), sizeof(int64_t));
    if (!kept_local) return -2;
    int fail = 0;

#pragma omp parallel
    {
        float *r = (float *)malloc(sizeof(float) * (size_t)D);
        float *zq = (float *)malloc(sizeof(float) * (size_t)D);
        float *acc = (float *)malloc(sizeof(float) * (size_t)D);
        int64_t *kept_t = (int64_t *)calloc((size_t)(n_run > 0 ? n_run : 1), sizeof(int64_t));
        double loss_t = 0.0;
        if (!r || !zq || !acc || !kept_t) {
#pragma omp atomic write
            fail = 1;
        } else {
#pragma omp for schedule(static)
            for (long f = 0; f < nframes; ++f) {
                const int b = (int)(f / T), t = (int)(f % T);
                for (int d = 0; d < D; ++d) {
                    r[d] = z[((size_t)b * D + d) * T + t];
                    acc[d] = 0.0f;
                }
                /* quantize.py:389 : imp_map * level * n_codebooks, two fp32 multiplies */
                float x = 0.0f;
                if (imp_map) {
                    float lv = level[(size_t)b * level_stride];
                    float t1 = imp_map[(size_t)b * T + t] * lv;
                    x = t1 * (float)Nq;
                }
                for (int s = 0; s < n_run; ++s) {
                    float z_e[VRVQ_CD], loss, g;
                    int idx = stage_quantize(w, s, r, z_e, zq, &loss, &g);
                    float m = 1.0f;
                    if (imp_map) { /* utils.py:59-60 */
                        float xm = x - (float)s;
                        m = (xm >= 0.0f) ? 1.0f : 0.0f;
                    }
                    const size_t o = ((size_t)b * n_run + s) * T + t;
                    codes[o] = idx;
                    if (mask) mask[o] = m;
                    if (loss_pf) loss_pf[o] = loss;
                    if (gap) gap[o] = g;
                    loss_t += (double)(m * loss);
                    if (m != 0.0f) kept_t[s] += 1;
                    if (latents)
                        for (int k = 0; k < VRVQ_CD; ++k)
                            latents[((size_t)b * n_run * VRVQ_CD + (size_t)s * VRVQ_CD + k) * T + t] = z_e[k];
                    for (int d = 0; d < D; ++d) {
                        float v = zq[d];
                        if (z_q_is) z_q_is[(((size_t)b * n_run + s) * D + d) * T + t] = v;
                        float mv = m * v; /* quantize.py:421 (VBR) / :194 (CBR) */
                        acc[d] = acc[d] + mv;
                        r[d] = r[d] - v; /* quantize.py:195,360 */
                    }
                }
                if (z_q)
                    for (int d = 0; d < D; ++d) z_q[((size_t)b * D + d) * T + t] = acc[d];
            }
#pragma omp critical
            {
                loss_total += loss_t;
                for (int s = 0; s < n_run; ++s) kept_local[s] += kept_t[s];
            }
        }
        free(r);
        free(zq);
        free(acc);
        free(kept_t);
    }
    if (loss_masked_sum) *loss_masked_sum = loss_total;
    if (kept)
        for (int s = 0; s < n_run; ++s) kept[s] = kept_local[s];
    free(kept_local);
    return fail ? -2 : 0;
}

/* models/quantize.py:217-249 : codes [B][n][T] -> z_q [B][D][T], z_p [B][8n][T], optional z_q_is */
int vrvq_oracle_from_codes(const vrvq_oracle_weights *w, const int64_t *codes, int B, int T, int n, float *z_q,
                           float *z_p, float *z_q_is) {
    const int D = w->input_dim, K = w->codebook_size;
    if (n < 0 || n > w->n_codebooks) return -1;
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (long f = 0; f < (long)B * T; ++f) {
        const int b = (int)(f / T), t = (int)(f % T);
        for (int d = 0; d < D; ++d) z_q[((size_t)b * D + d) * T + t] = 0.0f;
        for (int s = 0; s < n; ++s) {
            int64_t idx = codes[((size_t)b * n + s) * T + t];
            if (idx < 0 || idx >= K) {
#pragma omp atomic write
                bad = 1;
                continue;
            }
            const float *cr = w->cb_raw + ((size_t)s * K + (size_t)idx) * VRVQ_CD;
            if (z_p)
                for (int k = 0; k < VRVQ_CD; ++k)
                    z_p[((size_t)b * n * VRVQ_CD + (size_t)s * VRVQ_CD + k) * T + t] = cr[k];
            const float *w_out = w->w_out + (size_t)s * D * VRVQ_CD;
            const float *b_out = w->b_out + (size_t)s * D;
            for (int d = 0; d < D; ++d) {
                float acc = b_out[d];
                for (int k = 0; k < VRVQ_CD; ++k) acc = fmaf(w_out[(size_t)d * VRVQ_CD + k], cr[k], acc);
                if (z_q_is) z_q_is[(((size_t)b * n + s) * D + d) * T + t] = acc;
                float *o = &z_q[((size_t)b * D + d) * T + t];
                *o = *o + acc;
            }
        }
    }
    return bad ? -3 : 0;
}

/* models/quantize.py:251-285 : latents [B][8n][T] -> z_q, z_p (raw rows), codes */
int vrvq_oracle_from_latents(const vrvq_oracle_weights *w, const float *latents, int B, int T, int n, float *z_q,
                             float *z_p, int64_t *codes) {
    const int D = w->input_dim, K = w->codebook_size;
    if (n < 0 || n > w->n_codebooks) return -1;
#pragma omp parallel for schedule(static)
    for (long f = 0; f < (long)B * T; ++f) {
        const int b = (int)(f / T), t = (int)(f % T);
        for (int d = 0; d < D; ++d) z_q[((size_t)b * D + d) * T + t] = 0.0f;
        for (int s = 0; s < n; ++s) {
            float z_e[VRVQ_CD];
            for (int k = 0; k < VRVQ_CD; ++k)
                z_e[k] = latents[((size_t)b * n * VRVQ_CD + (size_t)s * VRVQ_CD + k) * T + t];
            int idx = search8(z_e, w->cb_nrm + (size_t)s * K * VRVQ_CD, w->c2 + (size_t)s * K, K, NULL);
            codes[((size_t)b * n + s) * T + t] = idx;
            const float *cr = w->cb_raw + ((size_t)s * K + (size_t)idx) * VRVQ_CD;
            if (z_p)
                for (int k = 0; k < VRVQ_CD; ++k)
                    z_p[((size_t)b * n * VRVQ_CD + (size_t)s * VRVQ_CD + k) * T + t] = cr[k];
            const float *w_out = w->w_out + (size_t)s * D * VRVQ_CD;
            const float *b_out = w->b_out + (size_t)s * D;
            for (int d = 0; d < D; ++d) {
                float acc = b_out[d];
                for (int k = 0; k < VRVQ_CD; ++k) acc = fmaf(w_out[(size_t)d * VRVQ_CD + k], cr[k], acc);
                float *o = &z_q[((size_t)b * D + d) * T + t];
                *o = *o + acc;
            }
        }
    }
    return 0;
}

/* models/utils.py:55-61 : x [B][T] -> mask [B][nq][T] */
int vrvq_oracle_mask_hard(const float *x, int B, int T, int nq, float *mask) {
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < nq; ++k)
            for (int t = 0; t < T; ++t) {
                float xm = x[(size_t)b * T + t] - (float)k;
                mask[((size_t)b * nq + k) * T + t] = (xm >= 0.0f) ? 1.0f : 0.0f;
            }
    return 0;
}

/* models/utils.py:64-73 : per-codebook sum of mask over (b,t), binary64.  The caller forms
 * sum_k bits[k] * out[k] / (B*T).  (The reference sums mask*bits in fp32, which is inexact
 * above 2^24; see DESIGN.md.) */
int vrvq_oracle_mask_sum(const float *mask, int B, int T, int nq, double *out) {
    for (int k = 0; k < nq; ++k) out[k] = 0.0;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < nq; ++k) {
            double s = 0.0;
            for (int t = 0; t < T; ++t) s += (double)mask[((size_t)b * nq + k) * T + t];
            out[k] += s;
        }
    return 0;
}
