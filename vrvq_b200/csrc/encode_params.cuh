// Kernel-side view of vrvq_encode_args, shared by the CUDA-core kernel (rvq_encode.cu) and the tensor-core kernel (rvq_encode_tc.cu).
#pragma once
#include "common.cuh"

namespace vrvq {

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
    int vec_st;  // 4 / 2 / 1 floats per global store of z_q, z_q_is
    long long *phase_cycles;  // profiling only (VRVQ_DEBUG_PHASES=1): [gridDim.x][8] clock64 totals per phase, else NULL
};

// validates the ABI struct and fills the kernel parameters (rvq_encode.cu)
int fill_encode_params(const vrvq_encode_args *a, EncodeParams &p);
// true if a call of this size should go to the tensor-core kernel (rvq_encode.cu)
bool prefer_tc_for_size(int B, int T);
// tensor-core path (rvq_encode_tc.cu): 1 if this call can run on it
int encode_tc_usable(const vrvq_encode_args *a);
int encode_tc(const vrvq_encode_args *a, const EncodeParams &p, void *stream);
int encode_tc_launch_info(const vrvq_encode_args *a, int *grid, int *block, int *smem);
// from_codes on the same kernel (FC instantiation)
int from_codes_tc_usable(const vrvq_from_codes_args *a);
int from_codes_tc(const vrvq_from_codes_args *a, void *stream);

}  // namespace vrvq
