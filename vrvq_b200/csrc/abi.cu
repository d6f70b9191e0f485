// extern "C" surface of libvrvq.so (see include/vrvq.h).  Host-side logic only: argument checks,
// weight packing, error strings.  The kernels live in rvq_encode.cu and rvq_aux.cu.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cstdint>

#include "common.cuh"

namespace vrvq {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return VRVQ_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VRVQ_ECUDA;
}

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("no CUDA device: %s (libvrvq has no CPU fallback)", cudaGetErrorString(e));
        return VRVQ_ENODEVICE;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess || major != 10) {
        cudaGetLastError();
        set_error("device %d has compute capability %d.x; libvrvq is built for sm_100a only", dev, major);
        return VRVQ_ENODEVICE;
    }
    return VRVQ_OK;
}

// implemented in rvq_encode.cu / rvq_aux.cu
int encode_supported(int D, int K, int cd);
int encode(const vrvq_encode_args *a, void *stream);
int encode_launch_info(const vrvq_encode_args *a, int *grid, int *block, int *smem);
const char *encode_kernel_name(const vrvq_encode_args *a);
int launch_mask_hard(const float *x, long long x_sb, int B, int T, int nq, float *mask, long long m_sb, long long m_sq, cudaStream_t st);
int launch_mask_sum(const float *mask, long long m_sb, long long m_sq, int B, int T, int nq, double *sums, cudaStream_t st);
int launch_remask(const float *zis, long long s_b, long long s_q, long long s_d, const float *imp, long long imp_sb, float level_scaled,
                  int B, int D, int T, int nq, float *zq, long long zq_sb, long long zq_sd, float *mask, long long m_sb, long long m_sq,
                  unsigned long long *kept, cudaStream_t st);
int launch_from_codes(const vrvq_from_codes_args *a, cudaStream_t st);
int launch_pack_codes(const long long *codes, long long c_sb, long long c_sq, const float *mask, long long m_sb, long long m_sq, int B, int T,
                      int nq, unsigned short *out, unsigned char *counts, int *error_flag, cudaStream_t st);
int launch_unpack_codes(const unsigned short *in, const unsigned char *counts, int B, int T, int nq, long long *codes, long long c_sb,
                        long long c_sq, float *mask, long long m_sb, long long m_sq, int *error_flag, cudaStream_t st);
int launch_snake_conv3(const float *x, long long x_sb, long long x_sc, const float *alpha, const float *wp, int cout_pad, const float *bias,
                       int B, int Cin, int Cout, int T, int sigmoid, float *y, long long y_sb, long long y_sc, cudaStream_t st);
int snake_conv3_tc_usable(int Cin, int Cout);
size_t conv3_tc_packed_floats(int Cout, int Cin);
int pack_conv3_tc_weights(int Cout, int Cin, const float *w, float *out);
int launch_snake_conv3_tc(const float *x, long long x_sb, long long x_sc, const float *alpha, const float *wtc, const float *bias,
                          const float *post_alpha, int B, int Cin, int Cout, int T, float *y, long long y_sb, long long y_sc, cudaStream_t st);
int launch_snake(const float *x, long long x_sb, long long x_sc, const float *alpha, int B, int C, int T, float *y, long long y_sb, long long y_sc,
                 cudaStream_t st);
int subnet_tail_usable(int C0, int C1, int C2);
int launch_subnet_tail(const float *x, long long x_sb, long long x_sc, int pre_activated, const float *alpha0, const float *w0, const float *bias0,
                       const float *alpha1, const float *w1, const float *bias1, const float *alpha2, const float *w2, const float *bias2, int B, int T,
                       float *y, long long y_sb, cudaStream_t st);
int launch_search_latents(const float *blob, int D, int K, const float *lat, long long l_sb, long long l_sc, int B, int T, int n_run,
                          long long *codes, long long c_sb, long long c_sq, cudaStream_t st);

// F.normalize(codebook) and codebook.pow(2).sum(1) with torch's op order (models/quantize.py:93,99;
// SURVEY.md A.3/A.4).  This translation unit is compiled with -ffp-contract=off on the host side.
static void normalize_row(const float *x, float *e, float *c2) {
    float s = 0.0f;
    for (int k = 0; k < CD; ++k) {
        volatile float sq = x[k] * x[k];
        s = s + sq;
    }
    const float n = sqrtf(s);
    const float den = n > 1e-12f ? n : 1e-12f;
    float s2 = 0.0f;
    for (int k = 0; k < CD; ++k) {
        e[k] = x[k] / den;
        volatile float sq = e[k] * e[k];
        s2 = s2 + sq;
    }
    *c2 = s2;
}


// round-to-nearest (ties away from zero, like cvt.rna.tf32.f32) TF32 head of x; x - head is exact in fp32
static float tf32_head(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;  // inf / nan
    u = (u + 0x1000u) & 0xffffe000u;
    float h;
    memcpy(&h, &u, 4);
    return h;
}

// element (row r, k) of a canonical K-major UMMA tile with R rows (common.cuh)
static inline size_t umma_idx(int R, int r, int k) { return ((size_t)(k / 4) * R + r) * 4 + (k % 4); }

// Tensor-core section of the blob (layout: common.cuh TcLayout).  w_in [Nq][8][D], w_out [Nq][D][8], b_out [Nq][D].
static void pack_tc_section(int Nq, int D, int K, const float *w_in, const float *b_in, const float *w_out, const float *b_out,
                            const float *codebook, float *tc) {
    const TcLayout T(D, Nq);
    memset(tc, 0, sizeof(float) * (size_t)T.total());
    // WIN: group g, chunk c = channels 32c..32c+31 (k = channel - 32c); rows 0..63 heads (row = 8*(stage - 8g) + out-channel), rows 64..127 remainders
    for (int g = 0; g < T.ngrp(); ++g)
        for (int c = 0; c < T.nch(); ++c) {
            float *tile = tc + T.off_win() + ((size_t)g * T.nch() + c) * 4096;
            for (int s = 8 * g; s < Nq && s < 8 * g + 8; ++s)
                for (int oc = 0; oc < CD; ++oc)
                    for (int k = 0; k < 32; ++k) {
                        const float x = w_in[((size_t)s * CD + oc) * D + 32 * c + k];
                        const float h = tf32_head(x);
                        tile[umma_idx(128, 8 * (s - 8 * g) + oc, k)] = h;
                        tile[umma_idx(128, 64 + 8 * (s - 8 * g) + oc, k)] = x - h;
                    }
        }
    // G[s][j] = W_in[s] W_out[j] (8x8) and gv[s][j] = W_in[s] b_out[j] for all j < s, in binary64
    std::vector<double> Gd((size_t)Nq * Nq * 72, 0.0);
    for (int j = 0; j < Nq; ++j)
        for (int s = j + 1; s < Nq; ++s) {
            double *G = Gd.data() + ((size_t)s * Nq + j) * 72;
            for (int c = 0; c < CD; ++c) {
                const float *wi = w_in + ((size_t)s * CD + c) * D;
                for (int k = 0; k < CD; ++k) {
                    double acc = 0.0;
                    for (int d = 0; d < D; ++d) acc += (double)wi[d] * (double)w_out[((size_t)j * D + d) * CD + k];
                    G[c * 8 + k] = acc;
                }
                double acc = 0.0;
                for (int d = 0; d < D; ++d) acc += (double)wi[d] * (double)b_out[(size_t)j * D + d];
                G[64 + c] = acc;
            }
        }
    // GX: virtual-channel chunks of group g >= 1 (cross-group corrections as extra K of the in_proj GEMM): B = -G[s][j]
    for (int g = 1; g < T.ngrp(); ++g)
        for (int v = 0; v < TcLayout::vp(g); ++v) {
            float *tile = tc + T.off_gx() + (size_t)(TcLayout::gx_base(g) + v) * 4096;
            for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * v + jj;
                if (j >= 8 * g) continue;  // padding chunk / stage of this or a later group: zero
                for (int s = 8 * g; s < Nq && s < 8 * g + 8; ++s)
                    for (int c = 0; c < CD; ++c)
                        for (int kk = 0; kk < CD; ++kk) {
                            const float x = (float)(-Gd[((size_t)s * Nq + j) * 72 + c * 8 + kk]);
                            const float h = tf32_head(x);
                            tile[umma_idx(128, 8 * (s - 8 * g) + c, 8 * jj + kk)] = h;
                            tile[umma_idx(128, 64 + 8 * (s - 8 * g) + c, 8 * jj + kk)] = x - h;
                        }
            }
        }
    // WOUT / BOUT: row i of chunk j <-> channel 128j + 4(i%32) + i/32
    for (int j = 0; j < T.nj(); ++j)
        for (int i = 0; i < 128; ++i) {
            const int ch = 128 * j + 4 * (i % 32) + i / 32;
            for (int s = 0; s < Nq; ++s) {
                float *hi = tc + T.off_wout() + ((size_t)s * T.nj() + j) * 3072, *lo = hi + 1024, *bt = hi + 2048;
                for (int k = 0; k < CD; ++k) {
                    const float x = w_out[((size_t)s * D + ch) * CD + k];
                    const float h = tf32_head(x);
                    hi[umma_idx(128, i, k)] = h;
                    lo[umma_idx(128, i, k)] = x - h;
                }
                const float bv = b_out[(size_t)s * D + ch];
                const float bh = tf32_head(bv);
                bt[umma_idx(128, i, 0)] = bh;
                bt[umma_idx(128, i, 1)] = bv - bh;
                float *bhi = tc + T.off_bout() + ((size_t)(s / 8) * T.nj() + j) * 2048, *blo = bhi + 1024;
                bhi[umma_idx(128, i, s % 8)] = bh;
                blo[umma_idx(128, i, s % 8)] = bv - bh;
            }
        }
    // GG: for j < s: G and g rounded once
    for (int j = 0; j < Nq; ++j)
        for (int s = j + 1; s < Nq; ++s) {
            float *G = tc + T.off_gg() + (size_t)TcLayout::pair_index(Nq, j, s) * 72;
            for (int i = 0; i < 72; ++i) G[i] = (float)Gd[((size_t)s * Nq + j) * 72 + i];
        }
    memcpy(tc + T.off_bin(), b_in, sizeof(float) * (size_t)Nq * CD);
    // b_in' = b_in[s] - sum over the stages j of EARLIER groups of g[s][j] (the in-group terms stay in GG)
    for (int s = 0; s < Nq; ++s)
        for (int c = 0; c < CD; ++c) {
            double acc = (double)b_in[(size_t)s * CD + c];
            for (int j = 0; j < 8 * (s / 8); ++j) acc -= Gd[((size_t)s * Nq + j) * 72 + 64 + c];
            tc[T.off_bin() + (size_t)Nq * CD + (size_t)s * CD + c] = (float)acc;
        }
    // CBK: F.normalize(codebook) (same arithmetic as the P1 section) as a [2 kg][K][4] tile, then c2[K]; SPC: the non-unit rows
    for (int s = 0; s < Nq; ++s) {
        float *cbk = tc + T.off_cbk() + (size_t)s * 9216;
        int32_t *spc = reinterpret_cast<int32_t *>(tc + T.off_spc() + (size_t)s * 16);
        int nsp = 0;
        for (int j = 0; j < K; ++j) {
            float e[CD];
            normalize_row(codebook + ((size_t)s * K + j) * CD, e, cbk + 8192 + j);
            for (int k = 0; k < CD; ++k) cbk[(size_t)(k / 4) * 4096 + (size_t)j * 4 + (k % 4)] = e[k];
            const float c2 = cbk[8192 + j];
            if (!(fabsf(c2 - 1.0f) <= 1e-3f)) {  // (also catches NaN rows)
                if (nsp < 15) spc[1 + nsp] = j;
                ++nsp;
            }
        }
        spc[0] = nsp > 15 ? 255 : nsp;
    }
}

}  // namespace vrvq

using namespace vrvq;

extern "C" {

int vrvq_abi_version(void) { return VRVQ_ABI_VERSION; }
int vrvq_flat_tile_order(int B, int T, int q) {
    if (B < 1 || T < 128 || q < 0 || (long long)q >= ((long long)B * T + 127) / 128) return VRVQ_EINVAL;
    return vrvq::flat_tile_at(q, B, T);
}

const char *vrvq_last_error(void) { return g_err; }

int vrvq_supported(int input_dim, int codebook_size, int codebook_dim) { return encode_supported(input_dim, codebook_size, codebook_dim); }

size_t vrvq_blob_bytes(int n_codebooks, int input_dim, int codebook_size, int codebook_dim) {
    if (n_codebooks <= 0 || input_dim <= 0 || codebook_size <= 0 || codebook_dim != CD) return 0;
    const BlobLayout L(input_dim, codebook_size);
    size_t floats = (size_t)BLOB_HDR_FLOATS + (size_t)n_codebooks * (size_t)L.stage_floats();
    if (tc_shape_ok(input_dim, codebook_size, n_codebooks)) floats += (size_t)TcLayout(input_dim, n_codebooks).total();
    return sizeof(float) * floats;
}

int vrvq_pack_weights(int n_codebooks, int input_dim, int codebook_size, int codebook_dim, const float *w_in, const float *b_in,
                      const float *w_out, const float *b_out, const float *codebook, void *blob_host, size_t blob_bytes) {
    if (codebook_dim != CD) {
        set_error("vrvq_pack_weights: codebook_dim=%d unsupported (only %d)", codebook_dim, CD);
        return VRVQ_EUNSUPPORTED;
    }
    if (n_codebooks <= 0 || input_dim <= 0 || codebook_size <= 0 || (input_dim % 4) != 0 || (codebook_size % 8) != 0) {
        set_error("vrvq_pack_weights: bad sizes Nq=%d D=%d K=%d (D and K must be positive multiples of 4)", n_codebooks, input_dim,
                  codebook_size);
        return VRVQ_EINVAL;
    }
    if (!w_in || !b_in || !w_out || !b_out || !codebook || !blob_host) {
        set_error("vrvq_pack_weights: NULL pointer");
        return VRVQ_EINVAL;
    }
    const size_t need = vrvq_blob_bytes(n_codebooks, input_dim, codebook_size, codebook_dim);
    if (blob_bytes < need) {
        set_error("vrvq_pack_weights: blob buffer too small (%zu < %zu)", blob_bytes, need);
        return VRVQ_EINVAL;
    }
    const int D = input_dim, K = codebook_size;
    const BlobLayout L(D, K);
    float *blob = static_cast<float *>(blob_host);
    BlobHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = BLOB_MAGIC;
    h.version = VRVQ_ABI_VERSION;
    h.n_codebooks = n_codebooks;
    h.input_dim = D;
    h.codebook_size = K;
    h.stage_floats = L.stage_floats();
    const bool tc = tc_shape_ok(D, K, n_codebooks);
    const size_t tc_off = (size_t)BLOB_HDR_FLOATS + (size_t)n_codebooks * (size_t)L.stage_floats();
    h.pad[0] = tc ? (int32_t)tc_off : 0;  // tensor-core section (rvq_encode_tc.cu), 0 = absent
    h.pad[1] = tc ? TcLayout(D, n_codebooks).total() : 0;
    memcpy(blob, &h, sizeof(h));
    for (int s = 0; s < n_codebooks; ++s) {
        float *st = blob + BLOB_HDR_FLOATS + (size_t)s * L.stage_floats();
        float *p0 = st + L.off_p0(), *p1 = st + L.off_p1(), *p2 = st + L.off_p2(), *raw = st + L.off_raw();
        const float *wi = w_in + (size_t)s * CD * D;  // [8][D] -> win_t[D][8]
        for (int d = 0; d < D; ++d)
            for (int c = 0; c < CD; ++c) p0[d * CD + c] = wi[(size_t)c * D + d];
        memcpy(p0 + (size_t)D * CD, b_in + (size_t)s * CD, sizeof(float) * CD);
        const float *cb = codebook + (size_t)s * K * CD;
        for (int j = 0; j < K; ++j) {  // P1 is pair-interleaved: element k of code j lives at [(j/2)*16 + 2k + (j&1)]
            float e[CD];
            normalize_row(cb + (size_t)j * CD, e, p1 + (size_t)K * CD + j);
            for (int k = 0; k < CD; ++k) p1[(size_t)(j >> 1) * 2 * CD + 2 * k + (j & 1)] = e[k];
        }
        memcpy(p2, w_out + (size_t)s * D * CD, sizeof(float) * (size_t)D * CD);
        memcpy(p2 + (size_t)D * CD, b_out + (size_t)s * D, sizeof(float) * D);
        memcpy(raw, cb, sizeof(float) * (size_t)K * CD);
    }
    if (tc) pack_tc_section(n_codebooks, D, K, w_in, b_in, w_out, b_out, codebook, blob + tc_off);
    return VRVQ_OK;
}

int vrvq_blob_codebook(const void *blob_host, size_t blob_bytes, int stage, float *cb_norm_out, float *c2_out) {
    if (!blob_host || blob_bytes < sizeof(BlobHeader)) {
        set_error("vrvq_blob_codebook: bad blob");
        return VRVQ_EBLOB;
    }
    BlobHeader h;
    memcpy(&h, blob_host, sizeof(h));
    if (h.magic != BLOB_MAGIC || stage < 0 || stage >= h.n_codebooks ||
        blob_bytes < vrvq_blob_bytes(h.n_codebooks, h.input_dim, h.codebook_size, CD)) {
        set_error("vrvq_blob_codebook: bad blob header or stage");
        return VRVQ_EBLOB;
    }
    const BlobLayout L(h.input_dim, h.codebook_size);
    const float *p1 = static_cast<const float *>(blob_host) + BLOB_HDR_FLOATS + (size_t)stage * L.stage_floats() + L.off_p1();
    if (cb_norm_out)
        for (int j = 0; j < h.codebook_size; ++j)
            for (int k = 0; k < CD; ++k) cb_norm_out[(size_t)j * CD + k] = p1[(size_t)(j >> 1) * 2 * CD + 2 * k + (j & 1)];
    if (c2_out) memcpy(c2_out, p1 + (size_t)h.codebook_size * CD, sizeof(float) * h.codebook_size);
    return VRVQ_OK;
}

int vrvq_rvq_encode_f32(const vrvq_encode_args *args, void *stream) { return encode(args, stream); }

int vrvq_rvq_encode_launch_info(const vrvq_encode_args *args, int *grid, int *block, int *smem_bytes) {
    return encode_launch_info(args, grid, block, smem_bytes);
}

const char *vrvq_rvq_encode_kernel_name(const vrvq_encode_args *args) { return encode_kernel_name(args); }

int vrvq_from_codes_f32(const vrvq_from_codes_args *args, void *stream) {
    int rc = check_device();
    if (rc) return rc;
    return launch_from_codes(args, static_cast<cudaStream_t>(stream));
}

int vrvq_search_latents_f32(const void *blob, int n_codebooks, int input_dim, int codebook_size, const float *latents,
                            int64_t lat_stride_b, int64_t lat_stride_c, int B, int T, int n_run, int64_t *codes, int64_t codes_stride_b,
                            int64_t codes_stride_q, void *stream) {
    if (B < 0 || T < 0 || n_run < 0 || n_run > n_codebooks || input_dim <= 0 || codebook_size <= 0 ||
        ((long long)B * T * n_run > 0 && (!blob || !latents || !codes))) {
        set_error("vrvq_search_latents_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_search_latents(static_cast<const float *>(blob), input_dim, codebook_size, latents, lat_stride_b, lat_stride_c, B, T, n_run,
                                 reinterpret_cast<long long *>(codes), codes_stride_b, codes_stride_q, static_cast<cudaStream_t>(stream));
}

int vrvq_generate_mask_hard_f32(const float *x, int64_t x_stride_b, int B, int T, int nq, float *mask, int64_t mask_stride_b,
                                int64_t mask_stride_q, void *stream) {
    if (B < 0 || T < 0 || nq < 0 || ((long long)B * T * nq > 0 && (!x || !mask))) {
        set_error("vrvq_generate_mask_hard_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_mask_hard(x, x_stride_b, B, T, nq, mask, mask_stride_b, mask_stride_q, static_cast<cudaStream_t>(stream));
}

int vrvq_mask_sum_f32(const float *mask, int64_t mask_stride_b, int64_t mask_stride_q, int B, int T, int nq, double *sums, void *stream) {
    if (B < 0 || T < 0 || nq < 0 || ((long long)B * T * nq > 0 && (!mask || !sums))) {
        set_error("vrvq_mask_sum_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_mask_sum(mask, mask_stride_b, mask_stride_q, B, T, nq, sums, static_cast<cudaStream_t>(stream));
}

int vrvq_remask_f32(const float *z_q_is, int64_t s_b, int64_t s_q, int64_t s_d, const float *imp_map, int64_t imp_stride_b,
                    float level_times_nq, int B, int D, int T, int nq, float *z_q, int64_t zq_stride_b, int64_t zq_stride_d, float *mask,
                    int64_t mask_stride_b, int64_t mask_stride_q, unsigned long long *kept, void *stream) {
    if (B < 0 || D < 0 || T < 0 || nq < 1 || ((long long)B * D * T > 0 && (!z_q_is || !imp_map || !z_q))) {
        set_error("vrvq_remask_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_remask(z_q_is, s_b, s_q, s_d, imp_map, imp_stride_b, level_times_nq, B, D, T, nq, z_q, zq_stride_b, zq_stride_d, mask,
                         mask_stride_b, mask_stride_q, kept, static_cast<cudaStream_t>(stream));
}

int vrvq_pack_codes_u16(const int64_t *codes, int64_t codes_stride_b, int64_t codes_stride_q, const float *mask, int64_t mask_stride_b,
                        int64_t mask_stride_q, int B, int T, int nq, uint16_t *codes_u16, uint8_t *counts, int32_t *error_flag, void *stream) {
    const bool work = (long long)B * T * nq > 0;
    if (B < 0 || T < 0 || nq < 0 || nq > 255 || (work && (!codes || !codes_u16 || (mask && !counts)))) {
        set_error("vrvq_pack_codes_u16: bad arguments (nq <= 255; counts is required with a mask)");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_pack_codes(reinterpret_cast<const long long *>(codes), codes_stride_b, codes_stride_q, mask, mask_stride_b, mask_stride_q, B,
                             T, nq, codes_u16, counts, error_flag, static_cast<cudaStream_t>(stream));
}

int vrvq_unpack_codes_u16(const uint16_t *codes_u16, const uint8_t *counts, int B, int T, int nq, int64_t *codes, int64_t codes_stride_b,
                          int64_t codes_stride_q, float *mask, int64_t mask_stride_b, int64_t mask_stride_q, int32_t *error_flag,
                          void *stream) {
    const bool work = (long long)B * T * nq > 0;
    if (B < 0 || T < 0 || nq < 0 || nq > 255 || (work && (!codes_u16 || !codes))) {
        set_error("vrvq_unpack_codes_u16: bad arguments (nq <= 255)");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_unpack_codes(codes_u16, counts, B, T, nq, reinterpret_cast<long long *>(codes), codes_stride_b, codes_stride_q, mask,
                               mask_stride_b, mask_stride_q, error_flag, static_cast<cudaStream_t>(stream));
}

size_t vrvq_conv3_packed_floats(int Cout, int Cin) {
    if (Cout <= 0 || Cin <= 0) return 0;
    const size_t cp = (size_t)((Cout + VRVQ_CONV3_COUT_ALIGN - 1) / VRVQ_CONV3_COUT_ALIGN) * VRVQ_CONV3_COUT_ALIGN;
    return (size_t)Cin * 3 * cp;
}

int vrvq_pack_conv3_weights(int Cout, int Cin, const float *w, float *packed, size_t packed_floats) {
    const size_t need = vrvq_conv3_packed_floats(Cout, Cin);
    if (need == 0 || !w || !packed || packed_floats < need) {
        set_error("vrvq_pack_conv3_weights: bad arguments (Cout=%d Cin=%d, %zu floats given, %zu needed)", Cout, Cin, packed_floats, need);
        return VRVQ_EINVAL;
    }
    const size_t cp = need / ((size_t)Cin * 3);
    memset(packed, 0, need * sizeof(float));
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int k = 0; k < 3; ++k) packed[((size_t)ci * 3 + k) * cp + co] = w[((size_t)co * Cin + ci) * 3 + k];
    return VRVQ_OK;
}

int vrvq_snake_conv3_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, const float *packed, const float *bias,
                         int B, int Cin, int Cout, int T, int apply_sigmoid, float *y, int64_t y_stride_b, int64_t y_stride_c, void *stream) {
    if (B < 0 || T < 0 || Cin < 1 || Cout < 1 || !alpha || !packed || !bias || ((long long)B * T > 0 && (!x || !y))) {
        set_error("vrvq_snake_conv3_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    if (Cin % 8 != 0) {
        set_error("vrvq_snake_conv3_f32: Cin=%d must be a multiple of 8 (all reference subnet widths are)", Cin);
        return VRVQ_EUNSUPPORTED;
    }
    int rc = check_device();
    if (rc) return rc;
    const int cp = (int)(vrvq_conv3_packed_floats(Cout, Cin) / ((size_t)Cin * 3));
    return launch_snake_conv3(x, x_stride_b, x_stride_c, alpha, packed, cp, bias, B, Cin, Cout, T, apply_sigmoid, y, y_stride_b, y_stride_c,
                              static_cast<cudaStream_t>(stream));
}

size_t vrvq_conv3_tc_packed_floats(int Cout, int Cin) { return conv3_tc_packed_floats(Cout, Cin); }

int vrvq_pack_conv3_tc_weights(int Cout, int Cin, const float *w, float *packed, size_t packed_floats) {
    const size_t need = conv3_tc_packed_floats(Cout, Cin);
    if (need == 0 || !w || !packed || packed_floats < need) {
        set_error("vrvq_pack_conv3_tc_weights: bad arguments (Cout=%d Cin=%d, %zu floats given, %zu needed; 0 = shape not served)", Cout, Cin,
                  packed_floats, need);
        return need == 0 ? VRVQ_EUNSUPPORTED : VRVQ_EINVAL;
    }
    return pack_conv3_tc_weights(Cout, Cin, w, packed);
}

int vrvq_snake_conv3_tc_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, const float *packed_tc, const float *bias,
                            const float *post_alpha, int B, int Cin, int Cout, int T, float *y, int64_t y_stride_b, int64_t y_stride_c,
                            void *stream) {
    if (!x || !packed_tc || !bias || !y || B < 0 || T < 0 || Cin <= 0 || Cout <= 0) {
        set_error("vrvq_snake_conv3_tc_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_snake_conv3_tc(x, x_stride_b, x_stride_c, alpha, packed_tc, bias, post_alpha, B, Cin, Cout, T, y, y_stride_b, y_stride_c,
                                 static_cast<cudaStream_t>(stream));
}

int vrvq_snake_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, int B, int C, int T, float *y, int64_t y_stride_b,
                   int64_t y_stride_c, void *stream) {
    if (!x || !alpha || !y || B < 0 || C < 0 || T < 0) {
        set_error("vrvq_snake_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_snake(x, x_stride_b, x_stride_c, alpha, B, C, T, y, y_stride_b, y_stride_c, static_cast<cudaStream_t>(stream));
}

int vrvq_subnet_tail_usable(int C0, int C1, int C2) { return subnet_tail_usable(C0, C1, C2); }

int vrvq_subnet_tail_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, int pre_activated, int C0, int C1, int C2, const float *alpha0,
                         const float *packed0, const float *bias0, const float *alpha1, const float *packed1, const float *bias1,
                         const float *alpha2, const float *packed2, const float *bias2, int B, int T, float *y, int64_t y_stride_b, void *stream) {
    if (!subnet_tail_usable(C0, C1, C2)) {
        set_error("vrvq_subnet_tail_f32: serves the widths %d -> %d -> %d -> 1 only (got %d -> %d -> %d)", 128, 32, 8, C0, C1, C2);
        return VRVQ_EUNSUPPORTED;
    }
    if (B < 0 || T < 0 || !alpha0 || !packed0 || !bias0 || !alpha1 || !packed1 || !bias1 || !alpha2 || !packed2 || !bias2 || ((long long)B * T > 0 && (!x || !y))) {
        set_error("vrvq_subnet_tail_f32: bad arguments");
        return VRVQ_EINVAL;
    }
    int rc = check_device();
    if (rc) return rc;
    return launch_subnet_tail(x, x_stride_b, x_stride_c, pre_activated, alpha0, packed0, bias0, alpha1, packed1, bias1, alpha2, packed2, bias2, B, T, y,
                              y_stride_b, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
