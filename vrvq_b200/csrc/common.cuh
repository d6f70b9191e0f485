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

// streaming (evict-first) global stores for write-once outputs
__device__ __forceinline__ void st_cs(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs2(float *p, float a, float b) { __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b)); }
__device__ __forceinline__ void st_cs4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// ---- error plumbing (host) -----------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int check_device();

}  // namespace vrvq
