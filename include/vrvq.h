/*
 * vrvq.h -- C ABI of libvrvq.so: the B200 (sm_100a) residual-vector-quantisation hot path.
 *
 * The reference (lixinghe1999/VRVQ) has no FFI: its boundary for this path is a Python class API
 * (SURVEY.md section 8(b)).  Each entry point below names the reference interface it replaces;
 * vrvq_b200/quantize.py, vrvq_b200/utils.py bind these with ctypes and expose the reference's own
 * names and signatures on top.
 *
 * Conventions
 *  - extern "C", plain pointers and integer sizes only.  No C++ types, no exceptions cross the boundary.
 *  - every function returns 0 on success or a negative VRVQ_E* code; vrvq_last_error() returns a
 *    thread-local human-readable message for the last failure on the calling thread.
 *  - all device buffers are allocated and owned by the caller and must stay alive until the stream
 *    has passed the launch.  The library allocates nothing persistent and holds no global mutable state.
 *  - launches go to the `stream` argument (a cudaStream_t / CUstream passed as void*; NULL = legacy
 *    default stream).  No entry point synchronises the device.
 *  - tensors are float32 (codes int64) laid out like the reference's: [B, channels, T] with unit
 *    stride along T; batch/row strides are given in ELEMENTS so that a frame range [t0, t0+T) of a
 *    larger tensor (the multi-GPU / chunked shard of SURVEY.md section 8(e)) is passed as a view.
 *  - there is no CPU fallback: without a CUDA device of compute capability 10.x the compute entry
 *    points return VRVQ_ENODEVICE.
 */
#ifndef VRVQ_H_
#define VRVQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRVQ_ABI_VERSION 1
#define VRVQ_CODEBOOK_DIM 8 /* conf/base.yml:11 -- the only codebook_dim the reference configs use */

enum {
    VRVQ_OK = 0,
    VRVQ_EINVAL = -1,       /* bad argument (NULL pointer, negative size, n_run out of range ...) */
    VRVQ_EUNSUPPORTED = -2, /* shape outside what the kernels are instantiated for (see vrvq_supported) */
    VRVQ_ECUDA = -3,        /* a CUDA runtime call failed; message has cudaGetErrorString */
    VRVQ_ENODEVICE = -4,    /* no sm_100 device is current */
    VRVQ_EBLOB = -5         /* weight blob header does not match the call */
};

int vrvq_abi_version(void);
const char *vrvq_last_error(void);

/* 1 if (input_dim, codebook_size, codebook_dim) has a kernel instantiation, else 0. */
int vrvq_supported(int input_dim, int codebook_size, int codebook_dim);

/* ---------------------------------------------------------------------------------------------
 * Weights.  Replaces the per-forward weight_norm hook + F.normalize(codebook) of
 *   models/layers.py:17-18 (WNConv1d), models/quantize.py:38-40,89-93,99.
 * The caller folds W = v * (g / ||v||) itself (torch._weight_norm on the CPU in the Python mirror,
 * so the fold is bit-identical to the reference's) and passes, per stage i < n_codebooks:
 *   w_in  [Nq][8][D]  b_in [Nq][8]  w_out [Nq][D][8]  b_out [Nq][D]  codebook [Nq][K][8]   (host, float32)
 * vrvq_pack_weights normalises the codebook rows and forms c2 = sum(c_hat^2) on the host with the
 * reference's exact arithmetic and writes the kernel's blob layout into
 * `blob_host` (vrvq_blob_bytes(...) bytes).  The caller copies the blob to the device once
 * (cudaMemcpy / tensor.cuda()); the device copy must be 16-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t vrvq_blob_bytes(int n_codebooks, int input_dim, int codebook_size, int codebook_dim);
int vrvq_pack_weights(int n_codebooks, int input_dim, int codebook_size, int codebook_dim, const float *w_in,
                      const float *b_in, const float *w_out, const float *b_out, const float *codebook,
                      void *blob_host, size_t blob_bytes);
/* Host-side read-back of the normalised codebook / c2 stored in a packed blob (tests, debugging). */
int vrvq_blob_codebook(const void *blob_host, size_t blob_bytes, int stage, float *cb_norm_out /*[K][8]*/,
                       float *c2_out /*[K]*/);

/* ---------------------------------------------------------------------------------------------
 * Fused RVQ encode.  Replaces, in one launch,
 *   ResidualVectorQuantize.forward (eval)      models/quantize.py:136-214
 *   VBRResidualVectorQuantize.forward (eval)   models/quantize.py:328-443
 *   VectorQuantize.forward / decode_latents    models/quantize.py:42-103
 *   generate_mask_ste (forward value) / generate_mask_hard    models/utils.py:45-61
 *   the numerator of cal_bpf_from_mask         models/utils.py:64-73
 * The importance map is an INPUT of this call; its producer (ImportanceSubnet, models/importance_subnet.py) has
 * its own entry point below, vrvq_snake_conv3_f32 (one fused Snake + k=3 conv block per launch).
 * ------------------------------------------------------------------------------------------- */
typedef struct vrvq_encode_args {
    uint32_t struct_size; /* = sizeof(vrvq_encode_args); lets the ABI grow */
    int32_t B, T;         /* batch items, latent frames per item */
    int32_t input_dim;    /* D */
    int32_t n_codebooks;  /* Nq of the packed model */
    int32_t codebook_size;
    int32_t n_run; /* stages to execute: Nq (VBR) or n_quantizers (CBR early exit, quantize.py:183-184) */

    const void *blob; /* device, from vrvq_pack_weights */

    const float *z; /* device [B][D][T] latent */
    int64_t z_stride_b, z_stride_d;

    /* VBR masking (quantize.py:389-395): mask[b,k,t] = (imp*level*Nq - k >= 0).  imp_map == NULL selects
     * CBR masking: every executed stage is kept (quantize.py:191-194). */
    const float *imp_map; /* device [B][T] or NULL */
    int64_t imp_stride_b;
    const float *level_dev; /* device level[b*level_stride] (stride 0 = one scalar) or NULL */
    int64_t level_stride;
    float level_host; /* used when level_dev == NULL */

    /* outputs; any pointer may be NULL (= not wanted) except codes */
    int64_t *codes; /* [B][n_run][T] */
    int64_t codes_stride_b, codes_stride_q;
    float *z_q; /* [B][D][T] = sum_k mask_k * z_q_k */
    int64_t z_q_stride_b, z_q_stride_d;
    float *z_q_is; /* [B][n_run][D][T] per-stage out_proj outputs (quantize.py:420) */
    int64_t z_q_is_stride_b, z_q_is_stride_q, z_q_is_stride_d;
    float *latents; /* [B][8*n_run][T] pre-normalisation z_e (quantize.py:205,426) */
    int64_t latents_stride_b, latents_stride_c;
    float *mask; /* [B][n_run][T] hard mask as float 0/1 */
    int64_t mask_stride_b, mask_stride_q;
    float *loss_pf; /* [B][n_run][T] per-frame MSE(z_e, c) (quantize.py:69-71, loss_per_frame=True) */
    int64_t loss_pf_stride_b, loss_pf_stride_q;

    /* accumulators the caller zeroes before the launch (device): */
    double *loss_masked_sum;  /* 1 double: += sum_{b,k,t} mask*loss_pf; both losses = this / (B*T) */
    unsigned long long *kept; /* [n_run]: += #(b,t) with mask==1 per stage; bpf = sum_k bits_k*kept_k/(B*T) */
} vrvq_encode_args;

int vrvq_rvq_encode_f32(const vrvq_encode_args *args, void *stream);

/* Number of CTAs / dynamic shared memory bytes the encode launch would use (bench/roofline reporting). */
int vrvq_rvq_encode_launch_info(const vrvq_encode_args *args, int *grid, int *block, int *smem_bytes);
/* Diagnostic: the tile visited at position q of the order in which the kernel's CTAs walk the "flat" tiling of calls without z_q_is
 * (tiles of 128 consecutive frames of the flattened (item, frame) sequence; the B - 1 tiles that span two items come first).  A
 * permutation of [0, ceil(B T / 128)); VRVQ_EINVAL outside (T < 128, q out of range).  No reference counterpart: scheduling detail. */
int vrvq_flat_tile_order(int B, int T, int q);

/* Which kernel vrvq_rvq_encode_f32 would launch for these arguments: "tc" = rvq_encode_tc_kernel (tcgen05 tensor cores),
 * "cuda" = rvq_encode_kernel (CUDA cores); NULL (and an error message) for invalid arguments.  Static strings. */
const char *vrvq_rvq_encode_kernel_name(const vrvq_encode_args *args);

/* ---------------------------------------------------------------------------------------------
 * Decode side.  Replaces ResidualVectorQuantize.from_codes, models/quantize.py:217-249
 * (z_q = sum_i out_proj_i(codebook_i[codes_i]); z_p = the gathered raw rows).
 * ------------------------------------------------------------------------------------------- */
typedef struct vrvq_from_codes_args {
    uint32_t struct_size;
    int32_t B, T, input_dim, n_codebooks, codebook_size;
    int32_t n_run; /* codes.shape[1] */
    const void *blob;
    const int64_t *codes; /* device [B][n_run][T] */
    int64_t codes_stride_b, codes_stride_q;
    const float *mask; /* optional device [B][n_run][T]: z_q = sum_i mask_i*z_q_i (VBR-aware decode); NULL = all ones */
    int64_t mask_stride_b, mask_stride_q;
    float *z_q; /* [B][D][T] */
    int64_t z_q_stride_b, z_q_stride_d;
    float *z_p; /* [B][8*n_run][T] or NULL */
    int64_t z_p_stride_b, z_p_stride_c;
    float *z_q_is; /* [B][n_run][D][T] or NULL */
    int64_t z_q_is_stride_b, z_q_is_stride_q, z_q_is_stride_d;
    int32_t *error_flag; /* optional device int (caller-zeroed): bit 0 is OR-ed in if any code is outside [0, K) */
} vrvq_from_codes_args;

int vrvq_from_codes_f32(const vrvq_from_codes_args *args, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Search only.  Replaces VectorQuantize.decode_latents as used by ResidualVectorQuantize.from_latents,
 * models/quantize.py:87-101,251-285: for every stage i < n_run, codes[b,i,t] = nearest normalised codebook
 * row of latents[b, 8i:8i+8, t].  (from_latents = this + vrvq_from_codes_f32.)
 * ------------------------------------------------------------------------------------------- */
int vrvq_search_latents_f32(const void *blob, int n_codebooks, int input_dim, int codebook_size, const float *latents,
                            int64_t lat_stride_b, int64_t lat_stride_c, int B, int T, int n_run, int64_t *codes,
                            int64_t codes_stride_b, int64_t codes_stride_q, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Mask utilities.  Replace models/utils.py:55-61 (generate_mask_hard) and :64-73 (cal_bpf_from_mask).
 * ------------------------------------------------------------------------------------------- */
/* mask[b,k,t] = (x[b,t] - k >= 0) ? 1 : 0 for k < nq.  x is [B][T] (the reference's [B,1,T]). */
int vrvq_generate_mask_hard_f32(const float *x, int64_t x_stride_b, int B, int T, int nq, float *mask,
                                int64_t mask_stride_b, int64_t mask_stride_q, void *stream);
/* sums[k] += sum_{b,t} mask[b,k,t] in binary64 (caller zeroes sums[nq], device).  The host then forms
 * bpf = sum_k bits[k]*sums[k] / (B*T). */
int vrvq_mask_sum_f32(const float *mask, int64_t mask_stride_b, int64_t mask_stride_q, int B, int T, int nq,
                      double *sums, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Level sweep re-mask.  Replaces the per-level recipe of scripts/inference.py:95-100 / README.md:75-79:
 *   mask = generate_mask_hard(imp_map * level * Nq, Nq); z_q = sum_k z_q_is[:,k] * mask[:,k]
 * without re-running the search.  kept[k] as in vrvq_rvq_encode_f32.
 * ------------------------------------------------------------------------------------------- */
int vrvq_remask_f32(const float *z_q_is, int64_t s_b, int64_t s_q, int64_t s_d, const float *imp_map,
                    int64_t imp_stride_b, float level_times_nq, int B, int D, int T, int nq, float *z_q,
                    int64_t zq_stride_b, int64_t zq_stride_d, float *mask, int64_t mask_stride_b,
                    int64_t mask_stride_q, unsigned long long *kept, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Compact code / mask wire format.  codes_u16 is what the reference's DACFile.save writes
 * (models/dac_base.py:34: codes.numpy().astype(np.uint16); read back with .astype(int), :52); counts replaces the Nq
 * floats per frame of the hard importance mask (models/utils.py:55-61: a prefix of ones) by one byte per frame.
 *   pack:   codes_u16[b,k,t] = k < counts[b,t] ? (uint16)codes[b,k,t] : 0;  counts[b,t] = number of leading ones of
 *           mask[b,:,t] (mask == NULL: every stage kept, counts may be NULL).  codes_u16 is contiguous [B][nq][T],
 *           counts contiguous [B][T].
 *   unpack: codes[b,k,t] = codes_u16[b,k,t];  mask[b,k,t] = k < counts[b,t] ? 1 : 0  (counts == NULL: all ones).
 * error_flag (device int32, caller-zeroed, may be NULL) gets bit 0 when a code is outside [0, 65535] and bit 1 when the mask is
 * not a 0/1 prefix mask (pack) or a count exceeds nq (unpack).
 * ------------------------------------------------------------------------------------------- */
int vrvq_pack_codes_u16(const int64_t *codes, int64_t codes_stride_b, int64_t codes_stride_q, const float *mask,
                        int64_t mask_stride_b, int64_t mask_stride_q, int B, int T, int nq, uint16_t *codes_u16,
                        uint8_t *counts, int32_t *error_flag, void *stream);
int vrvq_unpack_codes_u16(const uint16_t *codes_u16, const uint8_t *counts, int B, int T, int nq, int64_t *codes,
                          int64_t codes_stride_b, int64_t codes_stride_q, float *mask, int64_t mask_stride_b,
                          int64_t mask_stride_q, int32_t *error_flag, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Importance subnet.  Replaces one `nn.Sequential(Snake1d(Cin), WNConv1d(Cin, Cout,
 * kernel_size=3, padding=1))` block of models/importance_subnet.py:18-34 (forward :38-44); the caller chains the six
 * blocks (1024 -> 1024 -> 512 -> 128 -> 32 -> 8 -> 1) and sets apply_sigmoid on the last one (:43).
 *   y[b,co,t] = bias[co] + sum_{ci<Cin, k<3} w[co,ci,k] * snake(x[b,ci,t+k-1]),  zero outside [0,T)
 *   snake(v)  = v + 1/(alpha[ci] + 1e-9) * sin(alpha[ci] * v)^2                   (models/layers.py:25-31)
 * w is the weight-norm-folded conv weight (torch._weight_norm(v, g, 0), models/layers.py:17-18), re-laid out by
 * vrvq_pack_conv3_weights (host buffers) as packed[(ci*3+k) * Cout_padded + co], Cout_padded = Cout rounded up to
 * VRVQ_CONV3_COUT_ALIGN, zero filled.  alpha[Cin], bias[Cout], packed: device fp32.  Cin must be a multiple of 8.
 * ------------------------------------------------------------------------------------------- */
#define VRVQ_CONV3_COUT_ALIGN 128
size_t vrvq_conv3_packed_floats(int Cout, int Cin);
int vrvq_pack_conv3_weights(int Cout, int Cin, const float *w, float *packed, size_t packed_floats);
int vrvq_snake_conv3_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, const float *packed,
                         const float *bias, int B, int Cin, int Cout, int T, int apply_sigmoid, float *y, int64_t y_stride_b,
                         int64_t y_stride_c, void *stream);

/* The same block on the tensor cores (csrc/subnet_tc.cu: tcgen05 3xTF32 implicit GEMM, activations through TMA) for the wide
 * layers of the subnet: Cin a multiple of 64 and Cout a multiple of 128 (models/importance_subnet.py: 1024 -> 1024 and
 * 1024 -> 512, 95 % of its FLOPs); no sigmoid variant (only the 1-channel last block has one).  Weights are packed once on the
 * host by vrvq_pack_conv3_tc_weights into vrvq_conv3_tc_packed_floats(Cout, Cin) floats (0 = shape not served: use
 * vrvq_snake_conv3_f32).  x needs unit stride along T and an item pitch that is a multiple of 4 elements (or B = 1). */
size_t vrvq_conv3_tc_packed_floats(int Cout, int Cin);
int vrvq_pack_conv3_tc_weights(int Cout, int Cin, const float *w, float *packed, size_t packed_floats);
/* alpha == NULL: x is already Snake-activated (by vrvq_snake_f32 or by the previous block's post_alpha); post_alpha != NULL: the
 * output is stored as snake(y[co], post_alpha[co]), i.e. already activated for the NEXT block -- so a chain of tensor-core blocks
 * evaluates every Snake once instead of once per 128-channel output tile of its consumer. */
int vrvq_snake_conv3_tc_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, const float *packed_tc,
                            const float *bias, const float *post_alpha, int B, int Cin, int Cout, int T, float *y, int64_t y_stride_b,
                            int64_t y_stride_c, void *stream);
/* y = snake(x, alpha[c]) elementwise (models/layers.py:25-31): the activation of the first tensor-core block's input. */
int vrvq_snake_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, const float *alpha, int B, int C, int T, float *y,
                   int64_t y_stride_b, int64_t y_stride_c, void *stream);

/* The last three blocks of the subnet (C0 -> C1 -> C2 -> 1, models/importance_subnet.py:36-43 with the sigmoid of line 43) in one
 * launch (csrc/subnet_tc.cu: subnet_tail_kernel); serves the shipped widths 128 -> 32 -> 8 -> 1 (vrvq_subnet_tail_usable; others: chain
 * vrvq_snake_conv3_f32).  packed0..2 / alpha0..2 / bias0..2: the three blocks' vrvq_pack_conv3_weights output, Snake alphas and biases;
 * pre_activated != 0: x already went through block 0's Snake (a tensor-core producer's post_alpha).  y [B][1][T], unit stride along T. */
int vrvq_subnet_tail_usable(int C0, int C1, int C2);
int vrvq_subnet_tail_f32(const float *x, int64_t x_stride_b, int64_t x_stride_c, int pre_activated, int C0, int C1, int C2, const float *alpha0,
                         const float *packed0, const float *bias0, const float *alpha1, const float *packed1, const float *bias1,
                         const float *alpha2, const float *packed2, const float *bias2, int B, int T, float *y, int64_t y_stride_b, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VRVQ_H_ */
