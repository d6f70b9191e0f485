// Shared device helpers and the weight-blob layout for libvrvq.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vrvq.h"

namespace vrvq {

constexpr int CD = VRVQ_CODEBOOK_DIM;  // codebook_dim
constexpr uint32_t BLOB_MAGIC = 0x56525651u;  // "VRVQ"
constexpr int BLOB_HDR_FLOATS = 16;

// ---- weight blob ---------------------------------------------------------------------------
// header (16 x 4 B): magic, version, Nq, D, K, stage_stride_floats, 0...
// per stage, four 16-byte aligned sections (float counts):
//   P0 in_proj : win_t[D][8] (d-major, the 8 output channels contiguous), b_in[8]
//   P1 search  : cbn[K/2][8][2] normalised codebook rows, interleaved by pairs of adjacent codes, c2[K]
//   P2 out_proj: wout[D][8], bout[D]
//   RAW        : cbraw[K][8] un-normalised codebook rows (gathered per frame)
// P0/P1/P2 are the three pieces the encode kernel streams into shared memory with cp.async.bulk.
struct BlobLayout {
    int D, K;
    __host__ __device__ constexpr BlobLayout(int d, int k) : D(d), K(k) {}
    __host__ __device__ constexpr int p0_floats() const { return D * CD + 8; }
    __host__ __device__ constexpr int p1_floats() const { return K * CD + K; }
    __host__ __device__ constexpr int p2_floats() const { return D * CD + D; }
    __host__ __device__ constexpr int raw_floats() const { return K * CD; }
    __host__ __device__ constexpr int off_p0() const { return 0; }
    __host__ __device__ constexpr int off_p1() const { return p0_floats(); }
    __host__ __device__ constexpr int off_p2() const { return p0_floats() + p1_floats(); }
    __host__ __device__ constexpr int off_raw() const { return p0_floats() + p1_floats() + p2_floats(); }
    __host__ __device__ constexpr int stage_floats() const { return off_raw() + raw_floats(); }
};

struct BlobHeader {
    uint32_t magic, version;
    int32_t n_codebooks, input_dim, codebook_size, stage_floats;
    int32_t pad[10];
};
static_assert(sizeof(BlobHeader) == BLOB_HDR_FLOATS * 4, "blob header is 64 bytes");

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the warp until the phase completes): the result can be consumed many instructions later
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void *dst_smem, const void *src_gmem) {
    if constexpr (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "n"(BYTES)
                     : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- tensor memory (TMEM) as per-thread scratch ----------------------------------------------------
// The 256 KB of TMEM per SM are otherwise unused by this kernel (no tcgen05.mma); each thread parks its
// z_q accumulators in its own TMEM lane (32x32b shape: thread i of warp w owns lane 32*(w%4)+i) and moves
// them with tcgen05.ld / tcgen05.st (SASS: LDTM / STTM), which frees half of the register file.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // the allocating warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05 ld/st/wait only move data between registers and TMEM: no "memory" clobber, so shared/global accesses may be
// scheduled across them; `volatile` keeps their order relative to each other.
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;"); }
// wait::ld that also ties the destination registers, so no consumer can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[8]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]));
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;"); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}

// ---- tensor-core path (rvq_encode_tc.cu): extra blob section + tcgen05.mma helpers -------------------------------
// The TC section is appended to the blob after the per-stage sections (header.pad[0] = offset in floats, 0 = absent).
// All operand tiles are stored in the canonical K-major no-swizzle UMMA layout: element (row r, k) of a tile with R rows
// lives at float index ((k / 4) * R + r) * 4 + (k % 4), i.e. 8 rows x 16 bytes core matrices, SBO = 128 B, LBO = R * 16 B.
// The stages are processed in groups of 8 (group g = stages 8g .. 8g+7; one group for models with <= 8 codebooks,
// up to four: conf/base_24kbps.yml has 28):
//   WIN  [groups][D/32 chunks][8 kg][128 rows][4]   in_proj B operand of one 32-channel chunk: rows 0..63 = TF32 heads of W_in
//        (row n = 8*(stage - 8g) + out-channel, zero rows for absent stages), rows 64..127 = the remainders (W - head)
//   GX   [sum_g vp(g)][8 kg][128 rows][4]   "virtual channel" chunks of group g >= 1: the cross-group corrections
//        z_e[s] -= G[s][j] q_j (j < 8g) as extra K of the in_proj GEMM: chunk v covers stages j = 4v .. 4v+3 (k = 8 (j - 4v) + kk),
//        B = -G[s][j][c][kk] (heads / remainders as in WIN); vp(g) = 2g rounded up to a multiple of 4 (zero chunks pad)
//   WOUT [Nq][D/128 chunks][3072]: [hi | lo | bias] tiles of [2 kg][128 rows][4]; row i of chunk j = channel 128j + 4(i%32) + i/32;
//        bias tile: k = 0 -> head of b_out, k = 1 -> remainder, other k zero (multiplied by a constant tile of ones)
//   BOUT [groups][D/128 chunks][hi,lo][2 kg][128 rows][4]  bias as a B operand for the final GEMM: element (row, k) = b_out[8g + k][channel(row)]
//   GG   [Nq(Nq-1)/2][72]  for j < s: G = W_in[s] W_out[j] (8x8, row-major [c][k]) followed by g = W_in[s] b_out[j] (8)
//   BIN  [Nq][8] b_in, then [Nq][8] b_in' = b_in[s] - sum_{j < 8 (s / 8)} g[s][j]  (the cross-group bias terms folded in)
//   SPC  [Nq][16] (int32 bits): codes whose normalised row is not a unit vector (|c2 - 1| > 1e-3: rows with norm < 1e-12, which
//        F.normalize leaves at ~0).  The TF32 score filter ranks codes by 2 e.c only, which equals -distance up to a constant
//        only for unit rows; these codes are always re-scored exactly.  [0] = count (255: more than 15, the stage falls back to
//        the exact scan), [1..15] = indices
//   CBK  [Nq][9216]: normalised codebook as a K-major B operand [2 kg][1024 codes][4] (also read row-wise for the exact re-scoring),
//        then c2[1024]
struct TcLayout {
    int D, Nq;
    __host__ __device__ constexpr TcLayout(int d, int nq) : D(d), Nq(nq) {}
    __host__ __device__ constexpr int nch() const { return D / 32; }
    __host__ __device__ constexpr int nj() const { return D / 128; }
    __host__ __device__ constexpr int ngrp() const { return (Nq + 7) / 8; }
    __host__ __device__ static constexpr int vp(int g) { return (2 * g + 3) / 4 * 4; }         // virtual chunks of group g
    __host__ __device__ static constexpr int gx_base(int g) { return g <= 1 ? 0 : g == 2 ? 4 : g == 3 ? 8 : 16; }  // sum_{h<g} vp(h)
    __host__ __device__ constexpr int gx_chunks() const { return gx_base(ngrp()); }
    __host__ __device__ constexpr int off_win() const { return 0; }
    __host__ __device__ constexpr int off_gx() const { return ngrp() * nch() * 4096; }
    __host__ __device__ constexpr int off_wout() const { return off_gx() + gx_chunks() * 4096; }
    __host__ __device__ constexpr int off_bout() const { return off_wout() + Nq * nj() * 3072; }
    __host__ __device__ constexpr int off_gg() const { return off_bout() + ngrp() * nj() * 2048; }
    __host__ __device__ constexpr int gg_floats() const { return (Nq * (Nq - 1) / 2 * 72 + 3) / 4 * 4; }
    __host__ __device__ constexpr int off_bin() const { return off_gg() + gg_floats(); }
    __host__ __device__ constexpr int off_spc() const { return off_bin() + 2 * Nq * 8; }
    __host__ __device__ constexpr int off_cbk() const { return off_spc() + Nq * 16; }
    __host__ __device__ constexpr int total() const { return off_cbk() + Nq * 9216; }
    __host__ __device__ static constexpr int pair_index(int nq, int j, int s) { return j * nq - j * (j + 1) / 2 + (s - j - 1); }
};
// Flat tiling of the encode without z_q_is (rvq_encode_tc.cu, FLAT): tiles are 128 consecutive frames of the flattened (item, frame)
// sequence; the one holding flat frame j T (j = 1 .. B-1) spans two items and costs ~20 % more, so CTAs walk a permuted order: position
// q < B-1 is the q-th spanning tile (they go first, one per CTA), position q >= B-1 the (q - (B-1))-th of the others -- the least fixpoint of
// t = m + #{j : floor(j T / 128) <= t}.  A permutation of [0, ceil(B T / 128)) for T >= 128 (tests/test_abi_cpu.py through vrvq_flat_tile_order).
__host__ __device__ inline int flat_tile_at(int q, int B, int T) {
    const int S = B - 1;
    if (q < S) return (int)(((long long)(q + 1) * T) / 128);
    const int m = q - S;
    int t = m;
    for (int guard = 0; guard < 64; ++guard) {
        const long long cl = (128ll * (t + 1) - 1) / T;
        const int c = cl < S ? (int)cl : S;
        if (m + c == t) break;
        t = m + c;
    }
    return t;
}
constexpr int TC_MAX_NQ = 32;       // stages per model on the tensor-core path (4 groups of 8)
constexpr int TC_MAX_NQ_ZQIS = 8;   // with the per-stage outputs z_q_is: one group (the per-stage out_proj ring shares the TMEM)
__host__ __device__ constexpr bool tc_shape_ok(int D, int K, int Nq) { return K == 1024 && (D == 1024 || D == 512 || D == 256) && Nq >= 1 && Nq <= TC_MAX_NQ; }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
// instruction descriptor for kind::tf32 with fp32 accumulation, A and B K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// same with the A operand in tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}
// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// arrives on the mbarrier once every tcgen05 operation this thread issued so far has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
__device__ __forceinline__ void tmem_wait_ld32(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}
// round-to-nearest TF32 head of an fp32 value (low 13 mantissa bits zero); x - head is exact in fp32
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    return __uint_as_float(h);
}

// streaming (evict-first) global stores for write-once outputs
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs2(float *p, float a, float b) { __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
__device__ __forceinline__ void st_cs4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// ---- error plumbing (host) -----------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int check_device();

// cudaFuncAttributeMaxDynamicSharedMemorySize is set per (kernel, device): remember it per device ordinal, so that the first
// launch on a second GPU of the same process raises the limit there too.  One table per kernel instantiation (`auto` template
// parameter = the kernel's address); the race between host threads is benign (the attribute is idempotent).
template <auto Kernel>
int ensure_dynamic_smem(int bytes, const char *what) {
    static int granted[64];
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    const bool tracked = dev >= 0 && dev < 64;
    if (tracked && granted[dev] >= bytes) return 0;
    rc = check_cuda(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes), what);
    if (rc) return rc;
    if (tracked) granted[dev] = bytes;
    return 0;
}
// number of SMs of the current device (queried per call: a process may drive several devices)
inline int current_sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

}  // namespace vrvq
