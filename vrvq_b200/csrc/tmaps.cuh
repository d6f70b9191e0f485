// TMA tensor maps of a [B, C, T] fp32 activation tensor with an arbitrary (element) row pitch -- shared by the fused RVQ encode
// (rvq_encode_tc.cu: the latent) and the importance-subnet convolution (subnet_tc.cu: the feature map).
//
// A tensor map needs a 16-byte aligned base and 16-byte multiples for its pitches, which a [B, 1024, 862] tensor does not offer
// (rows are 8-byte aligned).  But the channels of one class k = channel % nc (nc = 1, 2 or 4) do: their rows are nc * pitch apart
// (a 16-byte multiple as soon as nc * pitch % 4 == 0), and the 16-byte aligned address at or below the first row of the class is
// a legal base -- the row then starts s_k = 0..3 elements into the map's row.  So the tensor is seen through nc rank-3 maps
// (x = frame + s_k, row = channel / nc, item); boxes start at 16-byte multiples of x (the TMA unit of B200 faults on others, and
// on an x extent that is not a 16-byte multiple: neither is rejected by cuTensorMapEncodeTiled), the consumer adds s_k to its
// column.  Frames outside [0, T) come back as zeros -- or, for the <= 3 positions next to a row's ends, as the neighbouring row's
// elements: consumers must not use those positions as frames (the encode kernel never does; the convolution zeroes them itself).
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the driver entry point is fetched through the runtime, no -lcuda)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vrvq {

struct ZMaps {
    CUtensorMap m[4];  // one per channel class
};

// one TMA box of a rank-3 tensor map into shared memory, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}

// one TMA box from shared memory into a rank-3 tensor (SASS: UTMASTG), tracked by the thread's bulk async-group; elements outside the
// tensor are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src_smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
                 "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources may be overwritten
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }            // writes are complete

// cuTensorMapEncodeTiled through the runtime's driver entry point (libvrvq.so does not link libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline encode_tiled_fn get_encode_tiled() {
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(ptr);
        else
            cudaGetLastError();
        tried = true;
    }
    return fn;
}

// Builds the class maps of x [B][C][T] (element strides stride_c, stride_b; unit stride along T) with boxes of box_w frames x
// (box_c / nc) channels of one class.  Returns false when the layout allows none (caller falls back); *lg_out = log2(nc),
// shift[k] = s_k.
inline bool build_class_maps(const float *x, int B, int C, int T, long long stride_c, long long stride_b, int box_w, int box_c, ZMaps *maps, int *lg_out,
                             int shift[4]) {
    memset(maps, 0, sizeof(*maps));
    *lg_out = 0;
    for (int k = 0; k < 4; ++k) shift[k] = 0;
    if (stride_c <= 0 || stride_b < 0 || T < 1 || C % 4 != 0 || box_c % 4 != 0) return false;
    if (!(stride_b % 4 == 0 || B == 1)) return false;
    encode_tiled_fn enc = get_encode_tiled();
    if (enc == nullptr) return false;
    const int al = (int)((reinterpret_cast<uintptr_t>(x) >> 2) & 3);  // base misalignment in floats
    int lg = (stride_c % 4 == 0 && al == 0) ? 0 : (stride_c % 2 == 0) ? 1 : 2;
    if (const char *env = getenv("VRVQ_DEBUG_ZNC_LOG2")) lg = atoi(env) > lg && atoi(env) <= 2 ? atoi(env) : lg;  // more classes than needed is always legal
    const int nc = 1 << lg;
    for (int k = 0; k < nc; ++k) {
        const int sk = (int)((al + (long long)k * stride_c) & 3);
        const float *base = x + (long long)k * stride_c - sk;
        const cuuint64_t dims[3] = {((cuuint64_t)T + (cuuint64_t)sk + 3ull) & ~3ull, (cuuint64_t)(C / nc), (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)nc * (cuuint64_t)stride_c * 4ull,
                                       B == 1 ? (cuuint64_t)nc * (cuuint64_t)stride_c * 4ull * (cuuint64_t)(C / nc) : (cuuint64_t)stride_b * 4ull};
        const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)(box_c / nc), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const bool ok = enc(&maps->m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        shift[k] = sk;
        if (getenv("VRVQ_DEBUG_ZMAP"))
            fprintf(stderr, "[vrvq tmap] class %d/%d: ok=%d shift=%d base%%16=%d dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)\n", k, nc, (int)ok, sk,
                    (int)(reinterpret_cast<uintptr_t>(base) % 16), (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                    (unsigned long long)strides[0], (unsigned long long)strides[1], box[0], box[1], box[2]);
        if (!ok) {
            for (int j = 0; j < 4; ++j) shift[j] = 0;
            return false;
        }
    }
    *lg_out = lg;
    return true;
}

// Tensor map for STORING boxes of box_w frames x box_c channels into y [B][C][T] (element strides): needs what the hardware needs --
// a 16-byte aligned base and 16-byte multiples for both pitches; returns false otherwise (the caller stores per lane).
inline bool build_store_map(float *y, int B, int C, int T, long long stride_c, long long stride_b, int box_w, int box_c, CUtensorMap *map) {
    memset(map, 0, sizeof(*map));
    if (stride_c <= 0 || stride_b < 0 || T < 1 || (reinterpret_cast<uintptr_t>(y) & 15) != 0 || stride_c % 4 != 0 || !(stride_b % 4 == 0 || B == 1)) return false;
    encode_tiled_fn enc = get_encode_tiled();
    if (enc == nullptr) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)T, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)stride_c * 4ull, B == 1 ? (cuuint64_t)stride_c * 4ull * (cuuint64_t)C : (cuuint64_t)stride_b * 4ull};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_c, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace vrvq
