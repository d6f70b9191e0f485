// Compact code / mask wire format: codes as uint16 the way the reference's DACFile stores them
// (models/dac_base.py:34, `codes.numpy().astype(np.uint16)`; read back with `.astype(int)`, :52), and the hard importance mask
// (a prefix of ones per frame, models/utils.py:55-61) as one uint8 count per frame instead of Nq floats.
// Streaming kernels: one thread per frame, every access coalesced along T, grid-stride over B*T.
#include "common.cuh"

namespace vrvq {

__global__ void pack_codes_kernel(const long long *__restrict__ codes, long long c_sb, long long c_sq, const float *__restrict__ mask,
                                  long long m_sb, long long m_sq, int B, int T, int nq, unsigned short *__restrict__ out,
                                  unsigned char *__restrict__ counts, int *__restrict__ error_flag) {
    const long long total = (long long)B * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / T), t = (int)(i % T);
        int n = nq, bad = 0;
        if (mask != nullptr) {
            // count of leading ones; a one after the first zero cannot be expressed by a count
            n = 0;
            bool open = true;
            for (int k = 0; k < nq; ++k) {
                const float m = mask[(long long)b * m_sb + (long long)k * m_sq + t];
                const bool one = m != 0.0f;
                if (one && m != 1.0f) bad = 2;
                if (one && !open) bad = 2;
                if (one && open) ++n;
                if (!one) open = false;
            }
        }
        for (int k = 0; k < nq; ++k) {
            const long long c = codes[(long long)b * c_sb + (long long)k * c_sq + t];
            if (c < 0 || c > 65535) bad |= 1;
            // stages past the count are not part of the payload: written as 0 so that the packed tensor is canonical
            out[((long long)b * nq + k) * T + t] = k < n ? (unsigned short)c : (unsigned short)0;
        }
        if (counts != nullptr) counts[i] = (unsigned char)n;
        if (bad && error_flag != nullptr) atomicOr(error_flag, bad);
    }
}

__global__ void unpack_codes_kernel(const unsigned short *__restrict__ in, const unsigned char *__restrict__ counts, int B, int T, int nq,
                                    long long *__restrict__ codes, long long c_sb, long long c_sq, float *__restrict__ mask, long long m_sb,
                                    long long m_sq, int *__restrict__ error_flag) {
    const long long total = (long long)B * T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / T), t = (int)(i % T);
        int n = nq;
        if (counts != nullptr) {
            n = counts[i];
            if (n > nq) {
                if (error_flag != nullptr) atomicOr(error_flag, 2);
                n = nq;
            }
        }
        for (int k = 0; k < nq; ++k) {
            codes[(long long)b * c_sb + (long long)k * c_sq + t] = (long long)in[((long long)b * nq + k) * T + t];
            if (mask != nullptr) mask[(long long)b * m_sb + (long long)k * m_sq + t] = k < n ? 1.0f : 0.0f;
        }
    }
}

static int wire_grid(long long total, int threads) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)(sms > 0 ? sms : 148) * 8;
    return (int)(blocks > cap ? cap : blocks);
}

int launch_pack_codes(const long long *codes, long long c_sb, long long c_sq, const float *mask, long long m_sb, long long m_sq, int B, int T,
                      int nq, unsigned short *out, unsigned char *counts, int *error_flag, cudaStream_t st) {
    const long long total = (long long)B * T;
    if (total == 0 || nq == 0) return VRVQ_OK;
    pack_codes_kernel<<<wire_grid(total, 256), 256, 0, st>>>(codes, c_sb, c_sq, mask, m_sb, m_sq, B, T, nq, out, counts, error_flag);
    return check_cuda(cudaGetLastError(), "pack_codes_kernel launch");
}

int launch_unpack_codes(const unsigned short *in, const unsigned char *counts, int B, int T, int nq, long long *codes, long long c_sb,
                        long long c_sq, float *mask, long long m_sb, long long m_sq, int *error_flag, cudaStream_t st) {
    const long long total = (long long)B * T;
    if (total == 0 || nq == 0) return VRVQ_OK;
    unpack_codes_kernel<<<wire_grid(total, 256), 256, 0, st>>>(in, counts, B, T, nq, codes, c_sb, c_sq, mask, m_sb, m_sq, error_flag);
    return check_cuda(cudaGetLastError(), "unpack_codes_kernel launch");
}

}  // namespace vrvq
