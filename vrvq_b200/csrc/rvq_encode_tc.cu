// Fused residual-vector-quantisation encode for B200 (sm_100a) -- tensor-core formulation.
//
// The reference loop (models/quantize.py:182-202 / :353-365) keeps a [D x T] residual and runs, per stage,
// in_proj (D -> 8), a nearest-neighbour search and out_proj (8 -> D).  Because both projections are linear, the
// pre-normalisation latents of ALL stages are one skinny GEMM of the input plus 8x8 corrections:
//     z_e[s] = W_in[s] (z - sum_{j<s} (W_out[j] q_j + b_out[j])) + b_in[s]
//            = (W_in[s] z + b_in[s]) - sum_{j<s} (G[s][j] q_j + g[s][j]),   G[s][j] = W_in[s] W_out[j],  g = W_in[s] b_out[j]
// so the residual never exists.  One CTA owns a tile of up to 120 consecutive frames of one batch item (128 without z_q_is, which needs no halo) plus the 8 frames
// before it (halo); tile row r = TMEM lane r = MMA row r <-> frame t0 - 8 + r.
//   phase L  warps 0-7 load the tile (lane = frame), split every value into a TF32 head and an exact fp32 remainder and write
//            both straight into tensor memory as the A operand; tcgen05.mma kind::tf32 (A in TMEM, W_in chunks in shared
//            memory by cp.async.bulk) with the 3xTF32 split hi*hi + hi*lo + lo*hi, M = 128 frames, N = 64 = 8 x Nq, K = D.
//            The tensor core rounds its fp32 accumulator toward zero on every k-step, so the accumulators are drained into
//            running fp32 sums (also in TMEM) every 128 channels; measured error 3.5e-7 rms / 1.3e-6 max relative, better
//            than an fp32 FMA chain (profiles/r1_micro_tc3x.txt).
//   phase S  per stage: bias + L2-normalise (torch's op order); search = TF32 score MMAs (128 frames x 64 codes per MMA) into
//            rotating TMEM buffers, a branch-free scan for the 8-code groups within a rigorous margin of the running
//            maximum, and exact fp32 re-scoring of those candidates with the reference's expression (first index on ties):
//            codes are decided by exactly the arithmetic of the CUDA-core kernel and the oracle.  Then raw-row gather,
//            straight-through q, per-frame loss, codes, the 8x8 corrections of the later stages through TMEM, and q (head,
//            remainder) as the A operand of the stage's out_proj: tcgen05.mma M = 128 frames x 128 channels x K = 8 (x3 split,
//            + a ones-tile x bias-tile MMA) into a two-deep TMEM ring that dedicated epilogue warps drain to global memory
//            (lane = frame: every warp-level store writes 128 contiguous bytes of one channel row), overlapping the next
//            stage's search.  Per channel class (channel % 4) the A descriptor is shifted by 1..8 rows so that every store
//            starts on a 32-byte sector whatever the row alignment is (that is what the halo is for).
//   final    z_q = sum_s mask_s (W_out[s] q_s + b_out[s]) as one GEMM over K = 8 Nq (+ the mask-weighted bias rows) from
//            the masked A tiles kept in shared memory.
// Warp roles: 0-7 load/split (phase L) and search (phase S), warps 0-3 additionally own one tile row per thread (normalise,
// merge, gather); 8-11 epilogue (TMEM -> global) and the phase-L accumulator drains; lane 0 of warp 12 issues the in_proj /
// out_proj / final MMAs, lane 0 of warp 13 every cp.async.bulk, lane 0 of warp 14 the search-score MMAs.  Everything is
// synchronised with mbarriers (bounded waits: a protocol bug traps instead of hanging); two CTA-wide barriers per tile.
//
// Exactness: only z_e differs from the reference's conv1d by rounding (as any two conv implementations do), so codes can
// differ only at fp32 near-ties (audited in tests); z_q / z_q_is carry the tensor core's accumulation rounding (<= 2e-6).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "encode_params.cuh"
#include "tmaps.cuh"

// Experiment kept compiled out: the epilogue warps as a third scan group when there is no z_q_is to store --
// slower in both forms tried (three groups on three 128-code buffers: +4-8 %; 64-code chunks in six buffers: +20 %), because the
// stage chain is bound by the frame threads' serial part and by hand-over latency, not by scan throughput; not maintained for GRP.
#ifndef VRVQ_THIRD_SCAN_GROUP
#define VRVQ_THIRD_SCAN_GROUP 0
#endif
#ifndef VRVQ_ROW_PREFETCH
#define VRVQ_ROW_PREFETCH 0
#ifndef VRVQ_SCAN_PROBE
#define VRVQ_SCAN_PROBE 0     // probe the full barrier of the scan group's next score chunk one chunk ahead (test_wait): no gain, measured
#endif
#ifndef VRVQ_SCAN_PREFETCH_GRP
#define VRVQ_SCAN_PREFETCH_GRP 0
#endif
#ifndef VRVQ_SCAN_PREFETCH
#define VRVQ_SCAN_PREFETCH 1  // without z_q_is: issue the tcgen05.ld of the next 64-code piece before the candidate appends of the current one
#endif
#endif

namespace vrvq {

constexpr int TCK = 1024;      // codebook size
constexpr int TC_NTH = 512;    // 16 warps: 8 search/load, 4 epilogue, MMA issuer, copy producer, search-MMA issuer / latent producer, second phase-L MMA issuer
constexpr int TC_NSEARCH = 256;

// shared memory map (bytes).  Phase L uses [0, 49152); phase S re-uses that region.
constexpr int SM_LR = 0;           // phase L W_in ring, 4 slots x 16 KB: [8 kg][128 rows][4] (rows 0-63 heads, 64-127 remainders)
constexpr int L_SLOT = 16384, WL_SLOTS = 4;  // (4 chunks of prefetch: a chunk lands ~1.8k cycles after its request while z streams in)
constexpr int L_SLOTS = 3;         // the matching A operand (the split latent chunk) lives in tensor memory: 3 slots
// phase L latent staging ring (free in phase L: the per-stage A tiles, the W_out ring and codebook buffer 1 live there later):
// 5 slots x [32 channels][132 floats], one slot per 32-channel chunk.  A slot is filled by the TMA unit -- cp.async.bulk.tensor
// boxes through the channel-class tensor maps (pick_zmode); or, for A/B runs, one cp.async.bulk per channel row from the 16-byte
// aligned address at or below the row's first frame -- so the 256 loader threads read the latent with conflict-free LDS
// (lane = frame) instead of issuing 128 misaligned LDGs per chunk and waiting on HBM.
constexpr int SM_ZR = WL_SLOTS * L_SLOT, Z_CH = 32, Z_PITCH = 136, Z_SLOT = Z_CH * Z_PITCH * 4, Z_SLOTS = 5;
constexpr int SM_AT = 0;           // phase S: per-stage A tiles, 8 x (hi 4 KB | lo 4 KB): [2 kg][128 frames][4]
constexpr int SM_AM = 65536;       // phase S: mask tile (A operand of the bias rows of the final GEMM) [2 kg][128][4]
constexpr int SM_WO = 69632;       // phase S: W_out ring, 4 slots x 12 KB (hi | lo | bias tile)
constexpr int SM_RAW = SM_WO + 16384;  // without z_q_is (W_out ring idle; candidate lists in its first 16 KB): un-normalised codebook of the stage, 32 KB
constexpr int SM_CB1 = 118784;     // search codebook buffer 1 (36864 B)
constexpr int SM_CB0 = 155648;     // search codebook buffer 0 (outside the phase-L region: prefetched during phase L)
constexpr int SM_ES = 192512;      // A operand of the search MMA: 2*e as a [2 kg][128 frames][4] tile
constexpr int SM_E2 = SM_ES + 4096;   // [128]
constexpr int SM_SB = SM_E2 + 512;    // candidate lists: 256 threads x 16 x u32 (16 KB); then [2][128] best distance / index
constexpr int SM_GG = SM_SB + 16384;   // correction matrices, <= 28 x 72 floats
constexpr int SM_BIN = SM_GG + 8192;  // b_in [8][8]
constexpr int SM_NK = SM_BIN + 256;   // keep counts [128]
constexpr int SM_ONES = SM_NK + 512;  // constant A tile [2 kg][128][4]: k = 0, 1 -> 1.0 (multiplies the bias tile)
constexpr int SM_BAR = SM_ONES + 4096;  // mbarriers
constexpr int SM_TMEM = SM_BAR + 1024;
constexpr int SM_TOTAL = SM_TMEM + 16;
// grouped instantiation: codes of all stages as u16 [32 stages][128 rows] -- stages 0-15 over the (unused) ones tile, 16-31 appended
constexpr int SM_XC = (SM_TOTAL + 127) / 128 * 128, SM_TOTAL_G = SM_XC + 4096;
constexpr int G_FI = 4, G_ASLOT = G_FI * 8192;  // grouped final GEMM: stages per ring step; A staging slot (2 slots at SM_AT)
static_assert(SM_TOTAL_G <= 232448 && 2 * G_ASLOT <= SM_AM, "grouped shared memory map");
constexpr int W_SLOT = 12288, W_SLOTS = 4;
constexpr int F_SLOT = 40960, F_SLOTS = 3, F_ITEMS = 5;  // final-GEMM ring (<= 5 chunks of 8 KB per step) over the W_out ring and both codebook buffers
static_assert(SM_TOTAL <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
static_assert(SM_WO + W_SLOTS * W_SLOT == SM_CB1 && SM_CB1 + 36864 == SM_CB0 && SM_CB0 + 36864 == SM_ES, "shared memory map");
static_assert(SM_WO + F_SLOTS * F_SLOT == SM_ES, "final ring covers [SM_WO, SM_ES)");
static_assert(SM_RAW + 32768 <= SM_CB1 && !VRVQ_THIRD_SCAN_GROUP, "un-normalised codebook sits behind the candidate lists of two scan groups in the idle W_out ring");
static_assert(SM_ZR + Z_SLOTS * Z_SLOT <= SM_CB0, "phase-L rings must not reach codebook buffer 0");
static_assert(SM_ZR % 128 == 0 && Z_SLOT % 128 == 0, "TMA destinations are 128-byte aligned");

enum {
    B_L_FULL = 0, B_L_EMPTY = 4, B_SET_FULL = 24, B_SET_EMPTY = 26, B_W_FULL = 28, B_W_EMPTY = 32,
    B_D_FULL = 36, B_D_EMPTY = 38, B_CB_FULL = 40, B_A_READY = 42, B_ZQ_READY = 50, B_MMA_DONE = 51, B_F_FULL = 52, B_F_EMPTY = 56,
    B_E_READY = 60, B_L_FULL2 = 73, B_WL_FULL = 76, B_WL_EMPTY = 80, B_Z_FULL = 84, B_Z_EMPTY = 89, B_SB_FULL = 94, B_SB_EMPTY = 100, B_RAW_FULL = 106, B_D4_FULL = 107, B_D4_EMPTY = 111, B_COUNT = 115
};
enum { ZMODE_LDG = 0, ZMODE_BULK = 1, ZMODE_TMA = 2 };  // how phase L fetches the latent

// TMEM columns: [0,64) running z_e sums of all stages; phase L accumulator sets at 64 + 128*set (hi*hi | lo terms);
// phase S out_proj ring at 64 + 128*buf.
constexpr uint32_t TM_COLS = 512;
constexpr uint32_t TM_RUN = 0, TM_SET = 64, TM_SCORE = 320;  // search scores: 3 x 64 columns at TM_SCORE (chunk g -> buffer g % 3)
constexpr uint32_t TM_AL = 320;  // phase L: A-operand ring, 3 slots x (32 columns heads | 32 columns remainders) of one 32-channel chunk
constexpr float SEARCH_MARGIN = 0.012f;  // > 2 x the TF32 score error bound 2^-9 * sum|2 e_k c_k| <= 2^-8 (unit vectors)

// Every wait in this kernel is bounded in WALL-CLOCK time (%globaltimer, 20 s -- far beyond any slow-down a debugger, a
// sanitizer, MPS time-slicing or memory contention can cause; a spin count would fire on correct runs there): a protocol bug
// reports which barrier and traps instead of hanging the GPU until the watchdog.
constexpr unsigned long long TC_WAIT_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __noinline__ void tc_wait_timeout(const uint64_t *bar, const uint64_t *bars, uint32_t parity) {
    printf("[vrvq tc] mbarrier %d wait timed out (parity %u, block %d, thread %d)\n", (int)(bar - bars), parity, (int)blockIdx.x, (int)threadIdx.x);
    __trap();
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// called every 2^14 failed polls with the time of the first such call (0 = this is the first): returns that time, or reports
// and traps once it is older than the timeout.  By value: nothing of the wait loop lives in local memory.
__device__ __noinline__ unsigned long long tc_wait_check(unsigned long long t0, const uint64_t *bar, const uint64_t *bars, uint32_t parity) {
    const unsigned long long now = global_timer_ns();
    if (t0 == 0) return now;
    if (now - t0 > TC_WAIT_TIMEOUT_NS) tc_wait_timeout(bar, bars, parity);
    return t0;
}
#define TC_WAIT(bar_, parity_)                                                   \
    do {                                                                         \
        uint64_t *b__ = (bar_);                                                  \
        const uint32_t p__ = (parity_);                                          \
        uint32_t spins__ = 0;                                                    \
        unsigned long long t0__ = 0;                                             \
        while (!mbar_try_wait(b__, p__)) {                                       \
            if ((++spins__ & 0x3fffu) == 0) t0__ = tc_wait_check(t0__, b__, bars, p__); \
        }                                                                        \
    } while (0)

struct TcParams {
    EncodeParams e;
    const float *tc;  // TC section of the blob
    int adv;          // frames per tile (multiple of 8, <= 120)
    int tiles_per_b, n_tiles;
    int flat;         // without z_q_is, T >= 128, TMA latent path: tiles are 128 consecutive frames of the FLATTENED (item, frame) sequence
                      // and may span two items (config 3: 431 tiles = 3 waves of 148 instead of 64 x 7 = 448 = 4 waves)
    int zmode;        // ZMODE_*
    int znc_log2;     // ZMODE_TMA: log2 of the number of channel classes (1, 2 or 4 tensor maps; see pick_zmode)
    int zshift[4];    // ZMODE_TMA: x offset of frame 0 in the rows of class k
    int single_issuer;  // debugging knob
    int stagger;      // cycles of start delay per (blockIdx % 8): de-synchronises the CTAs' W_in chunk requests (L2 hot spot)
    int trace;        // VRVQ_DEBUG_PHASES=3 (profiling instantiation only): block 0 records per-chunk timestamps of its first tile's phase L
    // from_codes mode (FC): codes [B][n_run][T] are an input, mask_in an optional 0/1 mask [B][n_run][T], error_flag is set when
    // a code lies outside [0, K)
    const long long *codes_in;
    long long cin_sb, cin_sq;
    const float *mask_in;
    long long min_sb, min_sq;
    int *error_flag;
};

// all K-major no-swizzle tiles of this kernel have 128 rows: LBO = 2048 B (next 4-wide k group), SBO = 128 B (next 8 rows)
constexpr uint64_t DESC_128 = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(2048 >> 4) << 16);
__device__ __forceinline__ uint64_t desc128(uint32_t saddr) { return DESC_128 | (uint64_t)(saddr >> 4); }

// Sector-aligned stores.  TMEM lane r of an out_proj accumulator holds frame t0 - 8 + delta + r, where delta in [1, 8] is the
// A-operand row shift of the MMA that produced it (tile row r' <-> frame t0 - 8 + r': rows 0-7 are the previous tile's last
// frames, re-computed here as a halo).  delta is chosen per channel class p = channel % 4 (column 32p + i of a 128-channel unit
// holds channel 128j + 4i + p) so that each warp-level store of 32 consecutive frames of one channel row starts on a 32-byte
// sector: with an even row pitch the start alignment of a row only depends on p.  Odd pitches get delta = 8 (no shift).
// Returns delta_p for p = 0..3 packed in 4 bytes; base8 = float index of (row 128j, frame t0 - 8) modulo 8, same for every j.
__device__ __forceinline__ uint32_t class_shifts(uint32_t base8, long long rstride) {
    uint32_t packed = 0;
#pragma unroll
    for (int pc = 0; pc < 4; ++pc) {
        uint32_t c = (base8 + (uint32_t)pc * (uint32_t)(rstride & 7)) & 7u;
        uint32_t dl = (rstride & 1) ? 8u : (c == 0u ? 8u : 8u - c);
        packed |= dl << (8 * pc);
    }
    return packed;
}
__device__ __forceinline__ uint32_t base_mod8(const float *ptr, long long off) {
    return (uint32_t)(((reinterpret_cast<uintptr_t>(ptr) >> 2) + (unsigned long long)off) & 7ull);
}

// FC = from_codes mode (models/quantize.py:217-249): no latent, no in_proj, no search -- the codes are an input; the gather, the
// out_proj units, the masking and the final GEMM are the encode path's own.
// GRP = models with more than 8 codebooks (conf/base_24kbps.yml: 28), without z_q_is: the stages run in groups of 8 on the same
// tile, one pass of the tile loop per (tile, group).  Group g re-streams the latent through the in_proj GEMM with ITS W_in rows and
// appends "virtual channel" chunks -- the raw codebook rows of the stages of the earlier groups, gathered by their codes, against
// -G[s][j] = -W_in[s] W_out[j] -- so the cross-group corrections are extra K of the same GEMM (the straight-through vector q_j is
// taken as the code's row there: it differs from it by <= 1 ulp, below the rounding of z_e itself); codes stay in shared memory,
// and the final z_q GEMM over K = 8 n_run regenerates its A tiles from them, four stages per ring step.
template <int D, bool ZQIS, bool PROFILE, bool FC, bool GRP = false, bool FLAT = false>
__global__ void __launch_bounds__(TC_NTH, 1) rvq_encode_tc_kernel(const TcParams P, const __grid_constant__ ZMaps zmaps) {
    static_assert(!GRP || (!ZQIS && !FC && !PROFILE), "the grouped instantiation has no z_q_is, from_codes or profiling variant");
    static_assert(!FLAT || (!ZQIS && !FC && !PROFILE), "flat tiling (tiles that may span two items) exists for the encode without z_q_is only");
    constexpr int NCH = D / 32, NG = NCH / 4, NJ = D / 128;
    // rows of a tile in front of its own frames: the 8-frame halo exists for the shifted, sector-aligned stores of z_q_is; without them a
    // tile is 128 own frames (config-4 shard: 41 instead of 44 tiles per item = 9 waves instead of 10)
    constexpr int HALO = ZQIS ? 8 : 0;  // 32-channel chunks, accumulator drained every 4 chunks
    // search-score chunks: 64 codes per MMA into 3 x 64 TMEM columns; without z_q_is the out_proj ring is idle during the
    // searches, so the scores take 3 x 128 columns from TM_SET on and half as many (170-cycle) barrier hand-overs
    // (third scan group: 64-code chunks in six buffers, so that each of the three groups has two of its own in flight)
    constexpr int SCW = (ZQIS || VRVQ_THIRD_SCAN_GROUP) ? 64 : 128, NSC = TCK / SCW;
    constexpr uint32_t NSB = (!ZQIS && VRVQ_THIRD_SCAN_GROUP) ? 6u : 3u;  // score buffers
    constexpr uint32_t TM_SC = ZQIS ? TM_SCORE : TM_SET;
    static_assert(NCH % 4 == 0, "D must be a multiple of 128");
    const EncodeParams &p = P.e;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM_BAR);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SM_TMEM);
    float *es = reinterpret_cast<float *>(smem + SM_ES);
    float *e2s = reinterpret_cast<float *>(smem + SM_E2);
    float *ggs = reinterpret_cast<float *>(smem + SM_GG);
    float *bins = reinterpret_cast<float *>(smem + SM_BIN);
    int *nkeep = reinterpret_cast<int *>(smem + SM_NK);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_run = p.n_run, Nq = p.Nq;
    const TcLayout TL(D, Nq);
    constexpr BlobLayout L = BlobLayout(D, TCK);
    const float *stages = p.blob + BLOB_HDR_FLOATS;
    const int n_my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (n_my_tiles <= 0) return;

    // ---- one-time setup -------------------------------------------------------------------------------------
    if (tid == 0) {
        // phase-L A slot (tensor memory): the 8 loader warps arrive; released by one tcgen05.commit
        // (a slot has two full barriers, for its even and its odd uses: see l_full)
        for (int i = 0; i < L_SLOTS; ++i) { mbar_init(&bars[B_L_FULL + i], 8); mbar_init(&bars[B_L_FULL2 + i], 8); mbar_init(&bars[B_L_EMPTY + i], 1); }
        // W_in ring slot: the producer's expect_tx + bytes; released by the commit of the chunk's MMAs (uses of a slot are 4 chunks
        // apart: always the same issuing thread, so one barrier per slot is enough here)
        for (int i = 0; i < WL_SLOTS; ++i) { mbar_init(&bars[B_WL_FULL + i], 1); mbar_init(&bars[B_WL_EMPTY + i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars[B_SET_FULL + i], 2); mbar_init(&bars[B_SET_EMPTY + i], 4);  // full: one commit per phase-L issuer
            mbar_init(&bars[B_D_FULL + i], 1); mbar_init(&bars[B_D_EMPTY + i], 4);
            mbar_init(&bars[B_CB_FULL + i], 1);
        }
        // (grouped: slots 0-1 are the A staging ring of the final GEMM, filled by the 8 search warps)
        for (int i = 0; i < W_SLOTS; ++i) { mbar_init(&bars[B_W_FULL + i], GRP ? 8 : 1); mbar_init(&bars[B_W_EMPTY + i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&bars[B_A_READY + i], 4);
        mbar_init(&bars[B_ZQ_READY], 4);
        mbar_init(&bars[B_MMA_DONE], 1);
        for (int i = 0; i < F_SLOTS; ++i) { mbar_init(&bars[B_F_FULL + i], 1); mbar_init(&bars[B_F_EMPTY + i], 1); }
        mbar_init(&bars[B_E_READY], 4);
        for (int i = 0; i < 4; ++i) { mbar_init(&bars[B_D4_FULL + i], 1); mbar_init(&bars[B_D4_EMPTY + i], 4); }  // grouped final GEMM: four accumulators
        mbar_init(&bars[B_RAW_FULL], 1);  // un-normalised codebook of the stage (without z_q_is): the producer's expect_tx + 32 KB
        for (int i = 0; i < 6; ++i) { mbar_init(&bars[B_SB_FULL + i], 1); mbar_init(&bars[B_SB_EMPTY + i], 4); }
        // latent staging slot: filled by the producer's expect_tx + the TMA bytes, released by the 8 loader warps
        for (int i = 0; i < Z_SLOTS; ++i) { mbar_init(&bars[B_Z_FULL + i], 1); mbar_init(&bars[B_Z_EMPTY + i], 8); }
        fence_mbar_init();
    }
    if (w == 12) {
        tmem_alloc(tmem_slot, TM_COLS);
        tmem_relinquish();
    }
    if constexpr (!GRP) {  // (grouped: the in-group correction matrices and biases are loaded per group)
        for (int i = tid; i < TL.gg_floats(); i += TC_NTH) ggs[i] = P.tc[TL.off_gg() + i];
        for (int i = tid; i < Nq * 8; i += TC_NTH) bins[i] = P.tc[TL.off_bin() + i];
    }
    for (int i = tid; i < 256; i += TC_NTH)  // ones tile: k group 0 = (1, 1, 0, 0) for every row, k group 1 = 0
        reinterpret_cast<float4 *>(smem + SM_ONES)[i] = i < 128 ? make_float4(1.f, 1.f, 0.f, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem);  // the same value, provably warp-uniform: MMAs issued by a whole warp (elect_one) take
                                                                   // their operands straight from uniform registers (no vote / elect / R2UR loop per MMA)
    const uint32_t smem_base = smem_u32(smem);
    if (P.stagger > 0) {
        const long long until = clock64() + (long long)P.stagger * (long long)(blockIdx.x & 7);
        while (clock64() < until) {
        }
    }

    // profiling only (VRVQ_DEBUG_PHASES=1): clock64 totals per phase for one thread of each role
    const int ph_role = tid == 0 ? 0 : tid == 256 ? 1 : tid == 384 ? 2 : tid == 416 ? 3 : -1;
    uint32_t gstage = 0;  // stages processed so far by this CTA (search-score ring / E_READY phase bookkeeping)
    const bool ph_on = PROFILE && p.phase_cycles != nullptr && ph_role >= 0;
    const bool lite = !PROFILE && p.phase_cycles != nullptr && tid == 0;  // VRVQ_DEBUG_PHASES=2: three timestamps, production code
    long long lite_t0 = 0, lite_l = 0;
    if (lite) lite_t0 = clock64();
    long long ph_last = 0, ph_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (ph_on) ph_last = clock64();
    auto ph_mark = [&](int k) {
        if (PROFILE && ph_on) {
            const long long t = clock64();
            ph_acc[k] += t - ph_last;
            ph_last = t;
        }
    };

    // VRVQ_DEBUG_PHASES=3: phase_cycles[1024 + 32 * event + chunk] = clock64 of block 0's first tile (events: 0 latent slot full,
    // 1 split done, 2 A slot empty, 3 A slot handed over, 4 issuer saw the slot, 5 MMAs issued, 6 producer saw the W_in slot empty,
    // 7 drains: [4g] accumulator set full, [4g+1] drained)
    auto trace = [&](int ev, int idx) {
        if (PROFILE && P.trace && blockIdx.x == 0 && p.phase_cycles != nullptr) p.phase_cycles[1024 + 32 * ev + idx] = clock64();
    };

    // VRVQ_DEBUG_PHASES=3, stage time line of block 0's first pass: phase_cycles[1024 + 32 * event + stage], events 8-15 thread 0 (frame
    // thread, scan group 0), 16-20 thread 128 (scan group 1), 21-23 the search-MMA issuer (see the table printed by the launcher)
    auto strace = [&](int ev, int stg, bool first_pass) {
        if (PROFILE && P.trace && blockIdx.x == 0 && first_pass && p.phase_cycles != nullptr) p.phase_cycles[1024 + 32 * ev + stg] = clock64();
    };
    // Full barrier of phase-L chunk n (slot n % 3, use n / 3).  Consecutive uses of a slot belong to different issuing threads
    // (3 is odd), and a parity wait can only tell the current phase from the previous one: an issuer that ran one use ahead of
    // the other would see "complete" for a slot that is still being filled.  So even and odd uses get their own barrier, each
    // waited on by one thread, in order.  Returns the barrier; *parity = the phase parity of use n / 3 on it.
    auto l_full = [&](uint32_t n, uint32_t *parity) -> uint64_t * {
        const uint32_t sl = n % L_SLOTS, use = n / L_SLOTS;
        *parity = (use >> 1) & 1u;
        return &bars[((use & 1u) ? B_L_FULL2 : B_L_FULL) + sl];
    };

    double loss_acc = 0.0;               // frame threads
    unsigned long long kept_acc = 0ull;  // lane k of warps 8-11 counts stage k
    uint32_t wn = 0, fn = 0, dn = 0;     // W_out ring / final ring step counters, out_proj unit counter
    uint32_t cbu0 = 0, cbu1 = 0;         // uses of the two codebook buffers

    // One pass of this loop = one group of <= 8 stages of one tile (one pass per tile unless GRP).  Ring positions are running
    // totals, since the passes of different groups have different chunk counts.
    const int n_grp = GRP ? (n_run + 7) / 8 : 1;
    const int n_my_passes = n_my_tiles * n_grp;
    uint32_t lbase = 0, gbase = 0, zbase = 0;  // phase-L chunks / accumulator groups / latent chunks consumed so far
    uint32_t astep = 0;                        // grouped final GEMM: A staging steps so far
    unsigned short *codes_lo = reinterpret_cast<unsigned short *>(smem + SM_ONES), *codes_hi = reinterpret_cast<unsigned short *>(smem + SM_XC);
    auto code_slot = [&](int s) -> unsigned short * { return (s < 16 ? codes_lo : codes_hi) + (s & 15) * 128; };  // [128 rows] of stage s (GRP)

    for (int it = 0; it < n_my_passes; ++it) {
        const int tile_i = GRP ? it / n_grp : it, grp = GRP ? it - tile_i * n_grp : 0;
        const int s0 = 8 * grp;                                   // first stage of this pass
        const int nl = GRP ? min(8, n_run - s0) : n_run;          // stages of this pass
        const bool first_grp = grp == 0, last_grp = grp == n_grp - 1;
        const int NV = GRP ? TcLayout::vp(grp) : 0, NCT = NCH + NV;  // virtual chunks (cross-group corrections); chunks of this pass
        int tile = (int)blockIdx.x + tile_i * (int)gridDim.x;
        if constexpr (FLAT) tile = flat_tile_at(tile, p.B, p.T);  // permuted order: the spanning tiles first, one per CTA (common.cuh)
        // Flat tiling (P.flat; never with z_q_is or from_codes): row r of the tile is flat frame 128 tile + r of the (item, frame) sequence;
        // rows [0, ra) belong to item b (frames t0 ..), rows [ra, ra + nb2) to item b + 1 (frames 0 ..) -- T >= 128, so at most two items.
        constexpr bool flat = FLAT;  // (its own instantiation: the extra per-row bookkeeping costs the other shapes 4-8 % through register pressure)
        const long long g0 = 128ll * (long long)tile;
        const int b = flat ? (int)(g0 / p.T) : tile / P.tiles_per_b;
        const int t0 = flat ? (int)(g0 - (long long)b * p.T) : (tile % P.tiles_per_b) * P.adv;
        const int ra = flat ? min(128, p.T - t0) : 128;
        const int nb2 = (flat && b + 1 < p.B) ? 128 - ra : 0;
        const int fv = flat ? ra : min(P.adv, p.T - t0);  // frames of item b this tile owns: tile rows HALO .. HALO+fv-1 (rows 0-7 with z_q_is: halo = the previous 8 frames)
        const bool last_tile = (tile % P.tiles_per_b) == P.tiles_per_b - 1;
        // The rest of the pass is instantiated twice in the flat kernel: for tiles inside one item (everything per-item stays warp-uniform, the
        // code of the per-item tiling) and for the ~B spanning tiles (SPAN: per-row item, two latent boxes per chunk) -- with the per-row
        // bookkeeping compiled into every pass the flat kernel cost 8-10 % per tile.
        auto pass_body = [&](auto span_c) {
        constexpr bool SPAN = FLAT && decltype(span_c)::value;
        constexpr uint32_t zper = SPAN ? 2u : 1u;  // latent ring slots per chunk: a second box for the rows of item b + 1
        // ... which starts at a NEGATIVE x, so that only the nb2 (+ class shift) frames it needs lie inside the tensor (what is outside arrives
        // as zeros without being fetched, like the frames past T of the first box): frame fr of item b + 1 sits in column x2off + fr + shift
        const int x2off = Z_PITCH - ((nb2 + 6) & ~3);
        // latent staging geometry of this tile (ZMODE_BULK / ZMODE_TMA): smem column of frame fr in a staged row = fr - tstart + shift(row)
        // (flat: t0 is any frame of the item, and a TMA box must start at a multiple of 4 elements)
        const int tstart = P.zmode == ZMODE_TMA ? (flat ? (t0 & ~3) : t0 - HALO) : max(t0 - HALO, 0);
        const uint32_t tpar = (uint32_t)it & 1u;  // phase parity of the once-per-pass barriers
        // A_READY[s] completes one phase per pass that HAS a stage s (the last group of a tile may be shorter): its phase index
        auto apar = [&](int s) -> uint32_t {
            return GRP ? (((uint32_t)tile_i * (uint32_t)((n_run - s + 7) / 8) + (uint32_t)grp) & 1u) : tpar;
        };
        const int n_stage_steps = ZQIS ? n_run * NJ : 0;
        const int G_NST = (n_run + G_FI - 1) / G_FI + 1;  // grouped final GEMM: ring steps per 128-channel chunk (stages in fours, then the bias)
        // grouped final GEMM: the A tiles (regenerated from the codes) do not depend on the channel chunk, so they are staged once per
        // QD = 4 chunks, whose products accumulate in four TMEM buffers (columns 0-511: nothing else is live by then) -- staging them
        // per chunk (8 x 8 steps of L2 gathers per tile) made the final phase 107k cycles per tile for 44k of tensor work
        constexpr int QD = NJ < 4 ? NJ : 4;
        constexpr bool D4 = !ZQIS;  // the final GEMM's accumulators: QD buffers from column 0 (no per-stage units share the tensor memory)
        const int n_astage_steps = (p.z_q == nullptr || !last_grp || !GRP) ? 0 : (NJ / QD) * G_NST;
        const int n_final_steps = (p.z_q == nullptr || !last_grp) ? 0 : GRP ? NJ * G_NST : NJ * ((n_run + F_ITEMS) / F_ITEMS);  // !GRP: ceil((n_run + 1) / F_ITEMS) per chunk
        if constexpr (GRP) {
            // this group's correction matrices (pairs j < s inside the group, local pair order) and biases (cross-group terms folded in)
            for (int i = tid; i < 28 * 72; i += TC_NTH) {
                const int pr = i / 72, e = i - pr * 72;
                int j = 0, rem = pr;  // local pair index -> (j, s2)
                while (rem >= 7 - j) { rem -= 7 - j; ++j; }
                const int s2 = j + 1 + rem;
                ggs[i] = (s2 < nl) ? P.tc[TL.off_gg() + TcLayout::pair_index(Nq, s0 + j, s0 + s2) * 72 + e] : 0.0f;
            }
            for (int i = tid; i < 64; i += TC_NTH) bins[i] = (i < nl * 8) ? P.tc[TL.off_bin() + Nq * 8 + s0 * 8 + i] : 0.0f;
            // (visible to the frame threads after the L -> S barrier)
        }

// Fold the phase-L accumulator set of channel group g into the running fp32 sums (also in TMEM); run by the epilogue warps,
// which have nothing to store yet (tq = the warp's TMEM lane quarter)
auto drain = [&](int g, uint32_t tq) {
            const uint32_t gg = gbase + (uint32_t)g, set = gg & 1u;
            TC_WAIT(&bars[B_SET_FULL + set], (gg >> 1) & 1u);
            tmem_fence_after_sync();
            if (PROFILE && tid == 256 && it == 0) trace(7, 4 * g);
            const uint32_t ts = tq + TM_SET + 128u * set;
#pragma unroll 2
            for (int c8 = 0; c8 < 8; ++c8) {
                uint32_t hh[8], lo[8], run[8];
                tmem_ld8(ts + 8 * c8, hh);
                tmem_ld8(ts + 64 + 8 * c8, lo);
                if (g > 0) tmem_ld8(tq + TM_RUN + 8 * c8, run);
                tmem_wait_ld(hh);
                tmem_wait_ld(lo);
                if (g > 0) tmem_wait_ld(run);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float v = __fadd_rn(__uint_as_float(hh[i]), __uint_as_float(lo[i]));
                    if (g > 0) v = __fadd_rn(__uint_as_float(run[i]), v);
                    run[i] = __float_as_uint(v);
                }
                tmem_st8(tq + TM_RUN + 8 * c8, run);
            }
            tmem_wait_st();
            tmem_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_SET_EMPTY + set]);
            if (PROFILE && tid == 256 && it == 0) trace(7, 4 * g + 1);
        };
        // Phase-L MMA issue, split over two issuing warps (12 and 15; the loop runs warp-wide on uniform values, one elected lane issues): issuer X takes the chunks c % 2 == X and
        // accumulates all three 3xTF32 products of its chunks -- hi*W_hi + hi*W_lo + lo*W_hi, three N = 64 MMAs per k-step -- into its
        // own 64 columns of the current accumulator set (columns [64 X, 64 X + 64) of TM_SET + 128 set; the drain adds the two
        // halves).  A tcgen05.mma blocks its issuing thread while the tensor pipe is busy and a completed mbarrier wait costs
        // ~170 cycles, so one thread alone alternates between the two (trace: 370 + 730 cycles per chunk); two threads overlap them.
        auto issue_phase_l = [&](int X0) {
            constexpr uint32_t ID_64 = umma_idesc_tf32(128, 64);
            const bool solo = P.single_issuer != 0;  // debugging knob (VRVQ_DEBUG_SINGLE_ISSUER=1): issuer 0 takes every chunk
            if (solo && X0 == 1) {
                return;
            }
            for (int c = X0; c < NCT; c += solo ? 1 : 2) {
                const int X = c & 1;
                const uint32_t n = lbase + (uint32_t)c, sl = n % L_SLOTS;
                const uint32_t wsl = n % WL_SLOTS;
                {
                    uint32_t par;
                    uint64_t *fb = l_full(n, &par);
                    TC_WAIT(fb, par);                                           // A operand (loaders)
                    TC_WAIT(&bars[B_WL_FULL + wsl], (n / WL_SLOTS) & 1u);      // B operand (W_in chunk)
                }
                if (PROFILE && it == 0 && X == 0 && lane == 0) trace(4, c);
                const uint32_t gg = gbase + (uint32_t)(c >> 2), set = gg & 1u;
                if ((c & 3) == X && gg >= 2) TC_WAIT(&bars[B_SET_EMPTY + set], ((gg >> 1) - 1) & 1u);
                tmem_fence_after_sync();
                const uint32_t a_hi = tmem_u + TM_AL + 64u * sl, a_lo = a_hi + 32;  // A in tensor memory: 8 columns per k-step
                const uint64_t bh = desc128(smem_base + SM_LR + wsl * L_SLOT), bl = bh + (1024 >> 4);  // B rows 0-63 heads, 64-127 remainders
                const uint32_t d = tmem_u + TM_SET + 128u * set + 64u * (uint32_t)X;
                if (elect_one()) {  // (the whole warp runs this loop on warp-uniform values; one lane issues)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_tf32_ts(d, a_lo + 8 * ks, bh + ks * (4096 >> 4), ID_64, (c & 3) != X || ks != 0);  // smallest terms first
                        umma_tf32_ts(d, a_hi + 8 * ks, bl + ks * (4096 >> 4), ID_64, true);
                        umma_tf32_ts(d, a_hi + 8 * ks, bh + ks * (4096 >> 4), ID_64, true);
                    }
                    umma_commit(&bars[B_L_EMPTY + sl]);
                    umma_commit(&bars[B_WL_EMPTY + wsl]);
                    if ((c & 3) == 2 + X) umma_commit(&bars[B_SET_FULL + set]);
                }
                if (PROFILE && it == 0 && X == 0 && lane == 0) trace(5, c);
            }
        };
        // ---- search of one stage by one scan group (quantize.py:96-101): the tensor core scores all 1024 codes per frame (TF32,
        // |error| <= 2^-8); scan group h (warps 4h .. 4h+3, thread = frame f) takes the score chunks ck = h, h + NSG, ..., keeps the
        // 8-code groups within SEARCH_MARGIN of its running maximum and re-scores those with the reference's exact fp32 arithmetic
        // (first index wins ties).  Groups: warps 0-3, 4-7 and -- when there is no z_q_is to store -- the epilogue warps 8-11.
        constexpr int NSG = (ZQIS || !VRVQ_THIRD_SCAN_GROUP) ? 2 : 3, NSCAN = 128 * NSG;
        constexpr bool PIPE = ZQIS;  // stage-ahead normalise (phase S below)
        unsigned char *list_base = smem + (ZQIS ? SM_SB : SM_WO);  // candidate lists [16][NSCAN] u32 (without z_q_is the W_out ring is idle)
        auto scan_stage = [&](int s, int h, int f, uint32_t tq, float &bd_out, int &bi_out) {
                    const uint32_t cbuse = (s & 1) ? (cbu1 + (uint32_t)(s >> 1)) : (cbu0 + (uint32_t)(s >> 1));
                    const float *CB = reinterpret_cast<const float *>(smem + ((s & 1) ? SM_CB1 : SM_CB0));
                    uint32_t *list = reinterpret_cast<uint32_t *>(list_base) + tid;  // (group maximum, group id) entries, [entry][thread]: no bank conflicts
                    float bd = __int_as_float(0x7f800000);
                    int bidx = 0x7fffffff;
                    float runmax = __int_as_float(0xff800000);
                    int cnt = 0;
                    // (no wait for the codebook buffer here: the scores waited for below come from MMAs that the issuer released only after
                    // it had seen B_CB_FULL of this stage, so the copy is complete and visible by then; a completed wait would still cost
                    // its ~200 cycles on the path to the first scores)
                    (void)cbuse;
                    // codes whose normalised row is not a unit vector (blob section SPC; none in a trained or random model): the
                    // score filter cannot rank them, so they are always re-scored exactly -- or, if there are many, everything is
                    const int *spc = reinterpret_cast<const int *>(P.tc + TL.off_spc()) + (s0 + s) * 16;
                    const int nsp = __ldg(spc);
                    constexpr int PPC = SCW / 64;                                // 64-code pieces per score chunk
                    constexpr int NPIECE = (NSC + NSG - 1) / NSG * PPC;          // this group's score chunks (SCW codes each), in pieces
                    // Software pipeline (without z_q_is, where the scans are the critical path: config-4 shape 760 -> 745 us; with z_q_is it
                    // costs 1.5 %): the tcgen05.ld of piece pc + 1 are issued as soon as the group maxima of piece pc are in registers, so
                    // their ~200 cycles pass under the candidate appends.  (Probing the full barrier of the next chunk one chunk ahead with a
                    // non-blocking test_wait, VRVQ_SCAN_PROBE, changed nothing.)
                    constexpr bool PREFETCH = VRVQ_SCAN_PREFETCH && !ZQIS && (!GRP || VRVQ_SCAN_PREFETCH_GRP);  // (the grouped instantiation spills with it: config 3 1165 -> 1216 us)
                    uint32_t va[32], vb[32];
                    bool nxt_ok = false;
                    auto piece_addr = [&](int pc_, uint32_t &sbuf_, uint32_t &par_, bool &in_range) -> uint32_t {
                        const int ck_ = NSG * (pc_ / PPC) + h, sub_ = pc_ % PPC;
                        in_range = pc_ < NPIECE && !(NSC % NSG != 0 && ck_ >= NSC);
                        const uint32_t gc_ = (gstage + (uint32_t)s) * (uint32_t)NSC + (uint32_t)ck_;
                        sbuf_ = gc_ % NSB;
                        par_ = (gc_ / NSB) & 1u;
                        return tq + TM_SC + (uint32_t)SCW * sbuf_ + 64u * (uint32_t)sub_;
                    };
                    // acquire the chunk that piece pc_ starts (pc_ % PPC == 0), probe the one after it, then issue the piece's loads
                    auto start_piece = [&](int pc_) {
                        uint32_t sbuf_, par_;
                        bool ok_;
                        const uint32_t tsc_ = piece_addr(pc_, sbuf_, par_, ok_);
                        if (!ok_) return;
                        if (pc_ % PPC == 0) {
                            if (!nxt_ok) TC_WAIT(&bars[B_SB_FULL + sbuf_], par_);
                            tmem_fence_after_sync();
                            uint32_t sbuf2, par2;
                            bool ok2;
                            piece_addr(pc_ + PPC, sbuf2, par2, ok2);
                            if (VRVQ_SCAN_PROBE) nxt_ok = ok2 && mbar_test_wait(&bars[B_SB_FULL + sbuf2], par2);
                        }
                        tmem_ld32(tsc_, va);
                        tmem_ld32(tsc_ + 32, vb);
                    };
                    start_piece(0);
                    if (PROFILE && (tid & 127) == 0) strace(tid == 0 ? 9 : 18, s, it == 0);
                    for (int pc = 0; pc < NPIECE; ++pc) {
                        const int ck = NSG * (pc / PPC) + h, sub = pc % PPC;
                        if (NSC % NSG != 0 && ck >= NSC) break;
                        const uint32_t gc = (gstage + (uint32_t)s) * (uint32_t)NSC + (uint32_t)ck, sbuf = gc % NSB;
                        ph_mark(8);
                        if (!PREFETCH && pc > 0) start_piece(pc);
                        tmem_wait_ld32(va);
                        tmem_wait_ld32(vb);
                        if (sub == PPC - 1) {
                            tmem_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bars[B_SB_EMPTY + sbuf]);
                        }
                        ph_mark(9);
                        // maxima of the eight 8-code groups of this chunk (FMNMX3)
                        float g[8];
#define VRVQ_GROUP_MAX(dst, v, o)                                                                                                       \
    {                                                                                                                                   \
        float m_;                                                                                                                       \
        asm("max.f32 %0, %1, %2, %3;" : "=f"(m_) : "f"(__uint_as_float(v[o])), "f"(__uint_as_float(v[o + 1])), "f"(__uint_as_float(v[o + 2]))); \
        asm("max.f32 %0, %1, %2, %3;" : "=f"(m_) : "f"(m_), "f"(__uint_as_float(v[o + 3])), "f"(__uint_as_float(v[o + 4])));               \
        asm("max.f32 %0, %1, %2, %3;" : "=f"(m_) : "f"(m_), "f"(__uint_as_float(v[o + 5])), "f"(__uint_as_float(v[o + 6])));               \
        dst = fmaxf(m_, __uint_as_float(v[o + 7]));                                                                                     \
    }
                        VRVQ_GROUP_MAX(g[0], va, 0) VRVQ_GROUP_MAX(g[1], va, 8) VRVQ_GROUP_MAX(g[2], va, 16) VRVQ_GROUP_MAX(g[3], va, 24)
                        VRVQ_GROUP_MAX(g[4], vb, 0) VRVQ_GROUP_MAX(g[5], vb, 8) VRVQ_GROUP_MAX(g[6], vb, 16) VRVQ_GROUP_MAX(g[7], vb, 24)
#undef VRVQ_GROUP_MAX
                        float cm;
                        asm("max.f32 %0, %1, %2, %3;" : "=f"(cm) : "f"(g[0]), "f"(g[1]), "f"(g[2]));
                        asm("max.f32 %0, %1, %2, %3;" : "=f"(cm) : "f"(cm), "f"(g[3]), "f"(g[4]));
                        asm("max.f32 %0, %1, %2, %3;" : "=f"(cm) : "f"(cm), "f"(g[5]), "f"(g[6]));
                        cm = fmaxf(cm, g[7]);
                        if (PREFETCH) start_piece(pc + 1);  // (va / vb are dead from here on: g[] holds all that the appends need)
                        runmax = fmaxf(runmax, cm);
                        const float thr = runmax - SEARCH_MARGIN;
                        const uint32_t gid0 = (uint32_t)(ck * (SCW / 8) + sub * 8);  // group id = code / 8
                        // branch-free append of every group whose maximum is within the margin: entry = (max & ~0xff) | group id
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t entry = (__float_as_uint(g[i]) & 0xffffff00u) | (gid0 + (uint32_t)i);
                            const uint32_t addr = smem_u32(list) + (4u * NSCAN) * (uint32_t)min(cnt, 15);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p st.shared.u32 [%3], %4;\n\t@p add.s32 %0, %0, 1;\n\t}"
                                : "+r"(cnt)
                                : "f"(g[i]), "f"(thr), "r"(addr), "r"(entry)
                                : "memory");
                        }
                        ph_mark(10);
                    }
                    if (PROFILE && (tid & 127) == 0) strace(tid == 0 ? 10 : 19, s, it == 0);
                    // 2e / e2 of the frame (written by its frame thread before it released the stage to the search-MMA issuer: the
                    // scores waited for above are causally after those writes)
                    const float4 ea = *reinterpret_cast<const float4 *>(&es[f * 4]), eb = *reinterpret_cast<const float4 *>(&es[512 + f * 4]);
                    const float e2 = e2s[f];
                    {
                        // Work list of 8-code groups to re-score exactly.  Normal case: the (<= 3) groups still within the margin of
                        // the final maximum (the stored maximum lost its low 8 bits: 1e-4 of slack), compacted first so that a warp
                        // runs the expensive body only max-over-lanes times.  Fallbacks: every listed group (more than 3 hits), or
                        // every group of this thread's share of the codebook (list overflow: degenerate frames, e.g. an all-zero latent).
                        const float thr = runmax - SEARCH_MARGIN - 1e-4f;
                        int c0 = -1, c1 = -1, c2 = -1, nc = 0;
                        const int nlist = cnt <= 16 ? cnt : 0;
                        for (int i = 0; i < nlist; ++i) {
                            const uint32_t entry = list[i * NSCAN];
                            const bool hit = __uint_as_float(entry & 0xffffff00u) >= thr;
                            const int gidx = (int)(entry & 0xffu);
                            c2 = (hit && nc == 2) ? gidx : c2;
                            c1 = (hit && nc == 1) ? gidx : c1;
                            c0 = (hit && nc == 0) ? gidx : c0;
                            nc += hit ? 1 : 0;
                        }
                        const int mode = (cnt > 16 || nsp > 15) ? 2 : nc > 3 ? 1 : 0;
                        const int total = mode == 2 ? ((NSC - h + NSG - 1) / NSG) * (SCW / 8) : mode == 1 ? cnt : nc;
                        for (int wi = 0; wi < total; ++wi) {
                            int gid;
                            if (mode == 0) {
                                gid = wi == 0 ? c0 : wi == 1 ? c1 : c2;
                            } else if (mode == 1) {
                                const uint32_t entry = list[wi * NSCAN];
                                gid = (__uint_as_float(entry & 0xffffff00u) >= thr) ? (int)(entry & 0xffu) : -1;
                            } else {
                                gid = (NSG * (wi / (SCW / 8)) + h) * (SCW / 8) + wi % (SCW / 8);  // wi-th group of this thread's share
                            }
                            if (gid >= 0) {
                                const float4 *r0 = reinterpret_cast<const float4 *>(CB + gid * 32), *r1 = reinterpret_cast<const float4 *>(CB + 4096 + gid * 32);
                                const float *c2p = CB + 8192 + gid * 8;
#pragma unroll
                                for (int j8 = 0; j8 < 8; ++j8) {  // exact distance of code 8*gid + jj; lexicographic (distance, index) minimum
                                    const int jj = (j8 + lane) & 7;  // lane-staggered: the rows of different groups share their banks
                                    const float4 ca = r0[jj], cb4 = r1[jj];
                                    float d = __fmul_rn(ea.x, ca.x);
                                    d = __fmaf_rn(ea.y, ca.y, d); d = __fmaf_rn(ea.z, ca.z, d); d = __fmaf_rn(ea.w, ca.w, d);
                                    d = __fmaf_rn(eb.x, cb4.x, d); d = __fmaf_rn(eb.y, cb4.y, d); d = __fmaf_rn(eb.z, cb4.z, d); d = __fmaf_rn(eb.w, cb4.w, d);
                                    const float t = __fadd_rn(__fadd_rn(e2, -d), c2p[jj]);  // dist = fl(fl(e2 - dot) + c2)
                                    const int jx = gid * 8 + jj;
                                    if (t < bd || (t == bd && jx < bidx)) { bd = t; bidx = jx; }
                                }
                            }
                        }
                        for (int i = 0; i < nsp && nsp <= 15; ++i) {  // (every scan group re-scores them: the merge is idempotent)
                            const int jx = __ldg(spc + 1 + i);
                            const float4 ca = *reinterpret_cast<const float4 *>(CB + jx * 4), cb4 = *reinterpret_cast<const float4 *>(CB + 4096 + jx * 4);
                            float d = __fmul_rn(ea.x, ca.x);
                            d = __fmaf_rn(ea.y, ca.y, d); d = __fmaf_rn(ea.z, ca.z, d); d = __fmaf_rn(ea.w, ca.w, d);
                            d = __fmaf_rn(eb.x, cb4.x, d); d = __fmaf_rn(eb.y, cb4.y, d); d = __fmaf_rn(eb.z, cb4.z, d); d = __fmaf_rn(eb.w, cb4.w, d);
                            const float t = __fadd_rn(__fadd_rn(e2, -d), CB[8192 + jx]);
                            if (t < bd || (t == bd && jx < bidx)) { bd = t; bidx = jx; }
                        }
                    }
                    ph_mark(11);
                    if (PROFILE && (tid & 127) == 0) strace(tid == 0 ? 11 : 20, s, it == 0);
                    bd_out = bd;
                    bi_out = bidx;
        };
        // A scan group's result for the merge by the frame threads: best (distance, index) per frame, and -- without z_q_is, where the
        // candidate lists live in the idle W_out ring and this area is free -- the un-normalised codebook row of that index, fetched
        // here so that its L2 latency passes during the hand-over instead of on the frame thread's critical path.
        auto publish_best = [&](int s, int h, int f, float bd, int bidx) {
            float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
            if constexpr (!ZQIS && VRVQ_ROW_PREFETCH) {
                if (bidx < TCK) {
                    const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)(s0 + s) * L.stage_floats() + L.off_raw() + (size_t)bidx * 8);
                    ra = __ldg(rawp);
                    rb = __ldg(rawp + 1);
                }
            }
            if constexpr (ZQIS) named_bar_sync(1, NSCAN);  // (with z_q_is the lists share this area: every group must be done with them)
            reinterpret_cast<float *>(smem + SM_SB)[h * 128 + f] = bd;
            reinterpret_cast<int *>(smem + SM_SB + 2048)[h * 128 + f] = bidx;
            if constexpr (!ZQIS && VRVQ_ROW_PREFETCH) {
                float4 *rowp = reinterpret_cast<float4 *>(smem + SM_SB + 4096) + (h * 128 + f) * 2;
                rowp[0] = ra;
                rowp[1] = rb;
            }
        };
        if (w < 8) {
            // =====================================================================================================
            // Search group.  Frame threads (warps 0-3): f = tid = TMEM lane.
            // =====================================================================================================
            const int f = tid & 127;            // tile row = TMEM lane
            const bool row2 = SPAN && f >= ra;  // (spanning tile) the row belongs to item b + 1
            const int bq = b + (row2 ? 1 : 0);  // its item
            const int fr = row2 ? f - ra : t0 - HALO + f;  // its frame
            const bool own = flat ? (f < ra || (SPAN && f - ra < nb2)) : (f >= HALO && f - HALO < fv);  // frames whose per-frame outputs this tile writes
            const bool inb = flat ? own : (fr >= 0 && (f < HALO || f - HALO < fv));  // rows that hold a real frame (halo rows of the first tile do not)
            const uint32_t tq = tmem + ((uint32_t)(32 * (w & 3)) << 16);  // this warp's lane quarter
            if constexpr (!FC) {
            ph_mark(0);
            // ---- phase L: load, split, stage (256 threads: frame f, half q of each 32-channel chunk) ----
            if (P.zmode != ZMODE_LDG) {
                // staged path: the chunk's rows are in shared memory (TMA); lane = frame, so every LDS is conflict-free
                const int q = tid >> 7;
                const bool valid = inb;
                const int col = row2 ? x2off + f - ra : min(max(fr - tstart, 0), flat ? 131 : 127);
                // row shift: (alignment of the row's first frame) mod 4 floats; rows 4 apart share it (32 and 16 rows apart too)
                uint32_t shw[4];
                if (P.zmode == ZMODE_TMA) {
                    // boxes start at x = tstart (the TMA unit faults on an x that is not a 16-byte multiple): a frame of class
                    // k = channel % nc sits zshift[k] elements further right in its staged row
#pragma unroll
                    for (int k = 0; k < 4; ++k) shw[k] = (uint32_t)P.zshift[k & ((1 << P.znc_log2) - 1)];
                } else {
                    const uint32_t sh0 = (uint32_t)(((reinterpret_cast<uintptr_t>(p.z) >> 2) + (unsigned long long)((long long)b * p.z_sb + tstart)) & 3ull);
                    const uint32_t zs4 = (uint32_t)(p.z_sd & 3);
#pragma unroll
                    for (int k = 0; k < 4; ++k) shw[k] = (sh0 + (uint32_t)k * zs4) & 3u;
                }
                // slot row of the i-th channel of this thread's 16: channel classes (channel % NC) are stored class-major; NC is a
                // compile-time constant per instance of the loop so that every LDS has an immediate offset
                auto stage = [&](auto nc_tag) {
                    constexpr int NC = decltype(nc_tag)::value;
                    for (int c = 0; c < NCH; ++c) {
                        const uint32_t zn = zbase + (uint32_t)c * zper, zsl = zn % Z_SLOTS;
                        const uint32_t zn2 = zn + 1u, zsl2 = zn2 % Z_SLOTS;  // (flat tiling, tile spans two items: the box of item b + 1)
                        TC_WAIT(&bars[B_Z_FULL + zsl], (zn / Z_SLOTS) & 1u);
                        if (zper == 2u) TC_WAIT(&bars[B_Z_FULL + zsl2], (zn2 / Z_SLOTS) & 1u);
                        if (PROFILE && tid == 0 && it == 0) trace(0, c);
                        const float *zr = reinterpret_cast<const float *>(smem + SM_ZR + (row2 ? zsl2 : zsl) * Z_SLOT) + ((16 * q) / NC) * Z_PITCH + col;
                        float h[16], l[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x = valid ? zr[((i % NC) * (Z_CH / NC) + i / NC) * Z_PITCH + shw[i & 3]] : 0.0f;
                            h[i] = __uint_as_float(__float_as_uint(x) & 0xffffe000u);  // TF32 head by truncation; x - head is exact
                            l[i] = __fsub_rn(x, h[i]);
                        }
                        __syncwarp();  // the slot is in registers: hand it back to the producer
                        if (lane == 0) {
                            mbar_arrive(&bars[B_Z_EMPTY + zsl]);
                            if (zper == 2u) mbar_arrive(&bars[B_Z_EMPTY + zsl2]);
                        }
                        if (PROFILE && tid == 0 && it == 0) trace(1, c);
                        const uint32_t n = lbase + (uint32_t)c, sl = n % L_SLOTS, use = n / L_SLOTS;
                        if (use >= 1) {
                            TC_WAIT(&bars[B_L_EMPTY + sl], (use - 1) & 1u);
                            tmem_fence_after_sync();
                        }
                        if (PROFILE && tid == 0 && it == 0) trace(2, c);
                        {
                            uint32_t hv[16], lv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) { hv[i] = __float_as_uint(h[i]); lv[i] = __float_as_uint(l[i]); }
                            const uint32_t ta = tq + TM_AL + 64u * sl + 16u * (uint32_t)q;
                            tmem_st16(ta, hv);
                            tmem_st16(ta + 32, lv);
                        }
                        tmem_wait_st();
                        tmem_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) { uint32_t par_; mbar_arrive(l_full(n, &par_)); }
                        if (PROFILE && tid == 0 && it == 0) trace(3, c);
                    }
                };
                const int lg = P.zmode == ZMODE_TMA ? P.znc_log2 : 0;
                if (lg == 0) stage(std::integral_constant<int, 1>{});
                else if (lg == 1) stage(std::integral_constant<int, 2>{});
                else stage(std::integral_constant<int, 4>{});
            } else
            {
                const int q = tid >> 7;
                const bool valid = inb;
                const float *zp = p.z + (long long)b * p.z_sb + fr + (long long)(16 * q) * p.z_sd;
                const long long zstep = 32 * p.z_sd;
                constexpr int PF = 2;  // chunks of latent in flight per thread (2 x 16 loads; 3 spills and is slower)
                float x[PF][16];
                auto ldchunk = [&](const float *src, float (&v)[16]) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = valid ? __ldcs(src + (long long)i * p.z_sd) : 0.0f;
                };
                const float *zc = zp;
#pragma unroll
                for (int u = 0; u < PF; ++u) { ldchunk(zc, x[u]); zc += zstep; }
                for (int c0 = 0; c0 < NCH; c0 += PF) {
#pragma unroll
                    for (int u = 0; u < PF; ++u) {
                        const int c = c0 + u;
                        float h[16], l[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            // TF32 head by truncation (one LOP3; the remainder x - head stays exact); W_in heads are rounded to nearest
                            h[i] = __uint_as_float(__float_as_uint(x[u][i]) & 0xffffe000u);
                            l[i] = __fsub_rn(x[u][i], h[i]);
                        }
                        if (c + PF < NCH) { ldchunk(zc, x[u]); zc += zstep; }
                        const uint32_t n = lbase + (uint32_t)c, sl = n % L_SLOTS, use = n / L_SLOTS;
                        if (use >= 1) {
                            TC_WAIT(&bars[B_L_EMPTY + sl], (use - 1) & 1u);
                            tmem_fence_after_sync();
                        }
                        // A operand straight into tensor memory: this thread's row (= lane), columns = its 16 channels of the chunk
                        {
                            uint32_t hv[16], lv[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) { hv[i] = __float_as_uint(h[i]); lv[i] = __float_as_uint(l[i]); }
                            const uint32_t ta = tq + TM_AL + 64u * sl + 16u * (uint32_t)q;
                            tmem_st16(ta, hv);
                            tmem_st16(ta + 32, lv);
                        }
                        tmem_wait_st();
                        tmem_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) { uint32_t par_; mbar_arrive(l_full(n, &par_)); }
                    }
                }
            }
            if constexpr (GRP) {
                // ---- virtual chunks: the code rows of the stages of the earlier groups as extra "channels" (B = -G, blob section GX) ----
                const int q = tid >> 7;
                for (int v = 0; v < NV; ++v) {
                    float h[16], l[16];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const int j = 4 * v + 2 * q + jj;  // stage whose row fills this thread's columns 8 jj .. 8 jj + 7 of the chunk half
                        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
                        if (j < s0) {
                            const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)j * L.stage_floats() + L.off_raw() + (size_t)code_slot(j)[f] * 8);
                            ra = __ldg(rawp);
                            rb = __ldg(rawp + 1);
                        }
                        const float x[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            h[8 * jj + k] = __uint_as_float(__float_as_uint(x[k]) & 0xffffe000u);
                            l[8 * jj + k] = __fsub_rn(x[k], h[8 * jj + k]);
                        }
                    }
                    const uint32_t n = lbase + (uint32_t)(NCH + v), sl = n % L_SLOTS, use = n / L_SLOTS;
                    if (use >= 1) {
                        TC_WAIT(&bars[B_L_EMPTY + sl], (use - 1) & 1u);
                        tmem_fence_after_sync();
                    }
                    {
                        uint32_t hv[16], lv[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) { hv[i] = __float_as_uint(h[i]); lv[i] = __float_as_uint(l[i]); }
                        const uint32_t ta = tq + TM_AL + 64u * sl + 16u * (uint32_t)q;
                        tmem_st16(ta, hv);
                        tmem_st16(ta + 32, lv);
                    }
                    tmem_wait_st();
                    tmem_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) { uint32_t par_; mbar_arrive(l_full(n, &par_)); }
                }
            }
            }  // !FC
            ph_mark(1);
            tmem_fence_before_sync();
            __syncthreads();  // L -> S: every phase-L MMA has completed (the last drain waited for them); region re-usable
            tmem_fence_after_sync();
            ph_mark(2);
            if (lite) lite_l += clock64() - lite_t0;

            if constexpr (FC) {
                if (w < 4) {
                    const bool rowok = fr >= 0 && fr < p.T && f < HALO + fv;  // the row holds a real frame (own or halo)
                    for (int s = 0; s < n_run; ++s) {
                        long long code = rowok ? P.codes_in[(long long)b * P.cin_sb + (long long)s * P.cin_sq + fr] : 0;
                        if (code < 0 || code >= TCK) {  // F.embedding raises on such an index (quantize.py:82): report it
                            if (P.error_flag != nullptr) atomicOr(P.error_flag, 1);
                            code = 0;
                        }
                        const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)s * L.stage_floats() + L.off_raw() + (size_t)code * 8);
                        const float4 ra = __ldg(rawp), rb = __ldg(rawp + 1);
                        const float qv[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                        if (p.latents != nullptr && own) {  // z_p: the gathered un-normalised rows
#pragma unroll
                            for (int k = 0; k < 8; ++k) p.latents[(long long)b * p.lat_sb + (long long)(s * 8 + k) * p.lat_sc + fr] = qv[k];
                        }
                        float h[8], l[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) { h[k] = tf32_hi(qv[k]); l[k] = __fsub_rn(qv[k], h[k]); }
                        unsigned char *at = smem + SM_AT + s * 8192 + f * 16;
                        *reinterpret_cast<float4 *>(at) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4 *>(at + 2048) = make_float4(h[4], h[5], h[6], h[7]);
                        *reinterpret_cast<float4 *>(at + 4096) = make_float4(l[0], l[1], l[2], l[3]);
                        *reinterpret_cast<float4 *>(at + 4096 + 2048) = make_float4(l[4], l[5], l[6], l[7]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars[B_A_READY + s]);
                    }
                    // masked A tiles for the final GEMM: row <- m * q (exact for 0/1 masks), mask tile <- m
                    TC_WAIT(&bars[B_MMA_DONE], tpar);
                    float mv[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s)
                        mv[s] = s >= n_run ? 0.0f : (P.mask_in != nullptr && rowok) ? P.mask_in[(long long)b * P.min_sb + (long long)s * P.min_sq + fr] : 1.0f;
                    for (int s = 0; s < n_run; ++s) {
                        if (mv[s] != 1.0f) {
                            unsigned char *at = smem + SM_AT + s * 8192 + f * 16;
#pragma unroll
                            for (int kg = 0; kg < 2; ++kg) {
                                const float4 hv = *reinterpret_cast<const float4 *>(at + kg * 2048), lv = *reinterpret_cast<const float4 *>(at + 4096 + kg * 2048);
                                const float qm[4] = {__fmul_rn(mv[s], __fadd_rn(hv.x, lv.x)), __fmul_rn(mv[s], __fadd_rn(hv.y, lv.y)),
                                                     __fmul_rn(mv[s], __fadd_rn(hv.z, lv.z)), __fmul_rn(mv[s], __fadd_rn(hv.w, lv.w))};
                                const float hh[4] = {tf32_hi(qm[0]), tf32_hi(qm[1]), tf32_hi(qm[2]), tf32_hi(qm[3])};
                                *reinterpret_cast<float4 *>(at + kg * 2048) = make_float4(hh[0], hh[1], hh[2], hh[3]);
                                *reinterpret_cast<float4 *>(at + 4096 + kg * 2048) =
                                    make_float4(__fsub_rn(qm[0], hh[0]), __fsub_rn(qm[1], hh[1]), __fsub_rn(qm[2], hh[2]), __fsub_rn(qm[3], hh[3]));
                            }
                        }
                    }
                    unsigned char *am = smem + SM_AM + f * 16;
                    *reinterpret_cast<float4 *>(am) = make_float4(mv[0], mv[1], mv[2], mv[3]);
                    *reinterpret_cast<float4 *>(am + 2048) = make_float4(mv[4], mv[5], mv[6], mv[7]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[B_ZQ_READY]);
                }
            } else {
            // ---- phase S ----
            float zev[8];  // frame threads: z_e of the current stage
            float4 *zsh = reinterpret_cast<float4 *>(smem + SM_SB + 8192);  // [2][128]: its home during the scan (without z_q_is)
            // bias, latents, normalise (quantize.py:66,92 in torch's op order) of stage s: frame threads.  Runs one stage ahead of the
            // searches: right after stage s - 1 has corrected the running sum of stage s (below), so the score MMAs of stage s
            // start while the corrections of the later stages are still being applied.
            auto prep_from = [&](int s, const uint32_t (&r8)[8]) {  // r8: the (corrected) running sum W_in[s] z of the frame
                float ss = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    zev[k] = __fadd_rn(__uint_as_float(r8[k]), bins[s * 8 + k]);
                    const float sq = __fmul_rn(zev[k], zev[k]);
                    ss = (k == 0) ? sq : __fadd_rn(ss, sq);
                }
                const float den = fmaxf(__fsqrt_rn(ss), 1e-12f);
                float e2 = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float ec = __fdiv_rn(zev[k], den);
                    const float sqe = __fmul_rn(ec, ec);
                    e2 = (k == 0) ? sqe : __fadd_rn(e2, sqe);
                    es[(k >> 2) * 512 + f * 4 + (k & 3)] = __fmul_rn(2.0f, ec);
                }
                e2s[f] = e2;
                if constexpr (!ZQIS) {  // z_e waits for the merge in shared memory: 8 registers fewer across the scan
                    zsh[f] = make_float4(zev[0], zev[1], zev[2], zev[3]);
                    zsh[128 + f] = make_float4(zev[4], zev[5], zev[6], zev[7]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_E_READY]);  // the search-MMA issuer may start on this stage
                if (p.latents != nullptr && own) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) p.latents[(long long)bq * p.lat_sb + (long long)((s0 + s) * 8 + k) * p.lat_sc + fr] = zev[k];
                }
            };
            auto prep = [&](int s) {
                uint32_t r8[8];
                tmem_ld8(tq + TM_RUN + 8 * s, r8);
                tmem_wait_ld(r8);
                prep_from(s, r8);
            };
            // correction of a later stage of the pass by the straight-through vector of stage s: z_e[s2] -= G[s2][s] q + g[s2][s];
            // corrected(): the new running sum in registers (not stored)
            auto corrected = [&](int s, int s2, const float (&qv)[8], uint32_t (&r8)[8]) {
                const float *G = ggs + TcLayout::pair_index(GRP ? 8 : Nq, s, s2) * 72;
                tmem_ld8(tq + TM_RUN + 8 * s2, r8);
                float a[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 g0 = *reinterpret_cast<const float4 *>(G + c * 8), g1 = *reinterpret_cast<const float4 *>(G + c * 8 + 4);
                    float acc = G[64 + c];
                    acc = __fmaf_rn(g0.x, qv[0], acc); acc = __fmaf_rn(g0.y, qv[1], acc);
                    acc = __fmaf_rn(g0.z, qv[2], acc); acc = __fmaf_rn(g0.w, qv[3], acc);
                    acc = __fmaf_rn(g1.x, qv[4], acc); acc = __fmaf_rn(g1.y, qv[5], acc);
                    acc = __fmaf_rn(g1.z, qv[6], acc); acc = __fmaf_rn(g1.w, qv[7], acc);
                    a[c] = acc;
                }
                tmem_wait_ld(r8);
#pragma unroll
                for (int c = 0; c < 8; ++c) r8[c] = __float_as_uint(__fsub_rn(__uint_as_float(r8[c]), a[c]));
            };
            auto correct = [&](int s, int s2, const float (&qv)[8]) {
                uint32_t r8[8];
                corrected(s, s2, qv, r8);
                tmem_st8(tq + TM_RUN + 8 * s2, r8);
            };
            // (the pipelined order pays where the stores are the critical path: config 2 with z_q_is 148.0 -> 145.8 us; without z_q_is
            // the frame threads are the critical path and it costs 2 %: 934 -> 955 us on the config-4 shape, so that variant keeps the
            // classic order: prep, barrier, scan, merge, all corrections)
            // Without z_q_is (SPLIT) the frame threads are the critical path, so after the merge they only correct and normalise the NEXT
            // stage (its score MMAs start right away) and hand the straight-through vector to warps 4-7 through shared memory, which
            // apply the corrections of the later stages while the tensor core scores (they have nothing else to do until the first
            // scores arrive).  Running sum x is touched by warps 4-7 in the stages <= x - 2 and by the frame threads in stage x - 1:
            // ordered by the barrier after the scans (tcgen05 fences on both sides).
            constexpr bool SPLIT = !ZQIS && !VRVQ_THIRD_SCAN_GROUP;
            float4 *qsh = reinterpret_cast<float4 *>(smem + SM_SB + 4096);  // [2][128]: q of the stage just merged (SPLIT)
            if ((PIPE || SPLIT) && w < 4) prep(0);
            if constexpr (SPLIT) named_bar_sync(1, NSCAN);  // 2e / e2 of the first stage visible to every scan group
            for (int s = 0; s < nl; ++s) {  // s: stage within this pass, sg = s0 + s: stage of the model
                const int sg = s0 + s;
                if constexpr (!PIPE && !SPLIT) {
                    if (w < 4) prep(s);
                    named_bar_sync(1, NSCAN);  // 2e / e2 of the stage visible to every scan group
                }
                ph_mark(3);
                if (PROFILE && tid == 0) strace(8, s, it == 0);
                float bd_pub;
                int bi_pub;
                scan_stage(s, w >> 2, f, tq, bd_pub, bi_pub);
                publish_best(s, w >> 2, f, bd_pub, bi_pub);
                if constexpr (!ZQIS) {  // un-normalised codebook of the stage in shared memory (in flight since the previous merge); a
                                        // completed wait still costs ~200 cycles: here, where the frame warps usually wait for warps 4-7
                    if (w < 4) TC_WAIT(&bars[B_RAW_FULL], (gstage + (uint32_t)s) & 1u);
                }
                if constexpr (SPLIT) tmem_fence_before_sync();
                named_bar_sync(1, NSCAN);
                if constexpr (SPLIT) tmem_fence_after_sync();
                ph_mark(4);
                if (PROFILE && tid == 0) strace(12, s, it == 0);
                if constexpr (SPLIT) {
                    if (w >= 4) {
                        named_bar_sync(2, NSCAN);  // q of stage s is in shared memory (2e / e2 of stage s + 1 reach these warps through the score barriers)
                        if (PROFILE && tid == 128) strace(16, s, it == 0);
                        if (s + 2 < nl) {
                            const float4 qa = qsh[f], qb = qsh[128 + f];
                            const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
                            for (int s2 = s + 2; s2 < nl; ++s2) correct(s, s2, qv);
                            tmem_wait_st();
                        }
                        if (PROFILE && tid == 128) strace(17, s, it == 0);
                    }
                }
                if (w < 4) {
                    // ---- merge of the groups' results (first index on ties), gather, loss, straight-through (quantize.py:69-75,81-85,102) ----
                    const float *sbd = reinterpret_cast<const float *>(smem + SM_SB);
                    const int *sbi = reinterpret_cast<const int *>(smem + SM_SB + 2048);
                    float best = sbd[f];
                    if constexpr (!ZQIS) {
                        const float4 za = zsh[f], zb = zsh[128 + f];
                        zev[0] = za.x; zev[1] = za.y; zev[2] = za.z; zev[3] = za.w;
                        zev[4] = zb.x; zev[5] = zb.y; zev[6] = zb.z; zev[7] = zb.w;
                    }
                    int bi = sbi[f], bh = 0;
#pragma unroll
                    for (int h2 = 1; h2 < NSG; ++h2) {
                        const float ob = sbd[h2 * 128 + f];
                        const int oi = sbi[h2 * 128 + f];
                        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; bh = h2; }
                    }
                    float4 ra, rb;
                    if (bi >= TCK) {  // no candidate at all (NaN latent): code 0, like an argmin over NaNs that never updates
                        bi = 0;
                        const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)sg * L.stage_floats() + L.off_raw());
                        ra = __ldg(rawp);
                        rb = __ldg(rawp + 1);
                    } else if constexpr (!ZQIS && VRVQ_ROW_PREFETCH) {  // the winning group already fetched the un-normalised row (publish_best)
                        const float4 *rowp = reinterpret_cast<const float4 *>(smem + SM_SB + 4096) + (bh * 128 + f) * 2;
                        ra = rowp[0];
                        rb = rowp[1];
                    } else if constexpr (!ZQIS) {  // (an L2 round trip per stage on the critical path otherwise)
                        const float4 *rawp = reinterpret_cast<const float4 *>(smem + SM_RAW) + (size_t)bi * 2;
                        ra = rawp[0];
                        rb = rawp[1];
                    } else {
                        const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)sg * L.stage_floats() + L.off_raw() + (size_t)bi * 8);
                        ra = __ldg(rawp);
                        rb = __ldg(rawp + 1);
                    }
                    const float cr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                    float qv[8], ls = 0.0f;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float diff = __fsub_rn(zev[k], cr[k]);
                        const float sq = __fmul_rn(diff, diff);
                        ls = (k == 0) ? sq : __fadd_rn(ls, sq);
                        qv[k] = __fadd_rn(zev[k], __fsub_rn(cr[k], zev[k]));
                    }
                    if (PROFILE && tid == 0) strace(13, s, it == 0);
                    if constexpr (SPLIT) {  // warps 4-7 start on the corrections of the stages after the next one
                        qsh[f] = make_float4(qv[0], qv[1], qv[2], qv[3]);
                        qsh[128 + f] = make_float4(qv[4], qv[5], qv[6], qv[7]);
                        named_bar_arrive(2, NSCAN);
                    }
                    // the next stage first: correct its running sum, normalise, hand 2e to the search-MMA issuer
                    if (SPLIT && s + 1 < nl) {  // (nobody reads that running sum again: it goes from registers straight into the normalise)
                        uint32_t r8[8];
                        corrected(s, s + 1, qv, r8);
                        prep_from(s + 1, r8);  // (overwrites zev: everything of stage s that needs z_e is done)
                    }
                    if (PIPE && s + 1 < nl) {
                        correct(s, s + 1, qv);
                        tmem_wait_st();
                        prep(s + 1);
                    }
                    if (PROFILE && tid == 0) strace(14, s, it == 0);
                    if (own) {
                        const float loss = __fdiv_rn(ls, 8.0f);
                        p.codes[(long long)bq * p.codes_sb + (long long)sg * p.codes_sq + fr] = (long long)bi;
                        if (p.loss_pf != nullptr) p.loss_pf[(long long)bq * p.loss_sb + (long long)sg * p.loss_sq + fr] = loss;
                        if (nkeep[f] > sg) loss_acc += (double)loss;
                    }
                    if constexpr (GRP) code_slot(sg)[f] = (unsigned short)bi;  // for the later groups' virtual chunks and the final GEMM
                    // A operand of this stage's out_proj: q split into TF32 head and remainder
                    if constexpr (!GRP) {
                        float h[8], l[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) { h[k] = tf32_hi(qv[k]); l[k] = __fsub_rn(qv[k], h[k]); }
                        unsigned char *at = smem + SM_AT + s * 8192 + f * 16;
                        *reinterpret_cast<float4 *>(at) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4 *>(at + 2048) = make_float4(h[4], h[5], h[6], h[7]);
                        *reinterpret_cast<float4 *>(at + 4096) = make_float4(l[0], l[1], l[2], l[3]);
                        *reinterpret_cast<float4 *>(at + 4096 + 2048) = make_float4(l[4], l[5], l[6], l[7]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[B_A_READY + s]);
                    // corrections of the stages after the next one (in the shadow of the next stage's score MMAs)
                    if constexpr (!SPLIT) {
                        for (int s2 = s + (PIPE ? 2 : 1); s2 < nl; ++s2) correct(s, s2, qv);
                        tmem_wait_st();
                    }
                }
                ph_mark(5);
                if (PROFILE && tid == 0) strace(15, s, it == 0);
            }
            if (!GRP && w < 4) {
                // ---- mask the A tiles for the final z_q GEMM (quantize.py:194 / :421) once every per-stage MMA has read them ----
                TC_WAIT(&bars[B_MMA_DONE], tpar);
                const int nk = nkeep[f];
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int s = nk; s < n_run; ++s) {
                    unsigned char *at = smem + SM_AT + s * 8192 + f * 16;
                    *reinterpret_cast<float4 *>(at) = zero;
                    *reinterpret_cast<float4 *>(at + 2048) = zero;
                    *reinterpret_cast<float4 *>(at + 4096) = zero;
                    *reinterpret_cast<float4 *>(at + 4096 + 2048) = zero;
                }
                unsigned char *am = smem + SM_AM + f * 16;
                *reinterpret_cast<float4 *>(am) = make_float4(nk > 0 ? 1.f : 0.f, nk > 1 ? 1.f : 0.f, nk > 2 ? 1.f : 0.f, nk > 3 ? 1.f : 0.f);
                *reinterpret_cast<float4 *>(am + 2048) = make_float4(nk > 4 ? 1.f : 0.f, nk > 5 ? 1.f : 0.f, nk > 6 ? 1.f : 0.f, nk > 7 ? 1.f : 0.f);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_ZQ_READY]);
            }
            if constexpr (GRP) {
                if (last_grp && p.z_q != nullptr) {
                    // ---- A tiles of the final z_q GEMM (quantize.py:194 / :421), regenerated from the codes: per 128-channel chunk j the
                    // stages go through a two-slot staging ring four at a time (rows of masked stages zero), then one step with the
                    // mask tiles that multiply the bias rows.  Warps 0-3 build items 0-1 of a step, warps 4-7 items 2-3; row = f.
                    const int hsel = w >> 2;
                    named_bar_sync(1, NSCAN);  // the code of the last stage (written by the frame threads in its merge) is visible to warps 4-7
                    uint32_t m = astep;
                    for (int jh = 0; jh < NJ / QD; ++jh)
                        for (int st = 0; st < G_NST; ++st, ++m) {
                            const uint32_t sa = m & 1u, use = m >> 1;
                            if (use >= 1) TC_WAIT(&bars[B_W_EMPTY + sa], (use - 1) & 1u);
                            unsigned char *base = smem + SM_AT + sa * G_ASLOT;
                            const int nk = nkeep[f];
#pragma unroll
                            for (int ii = 0; ii < 2; ++ii) {
                                const int i = 2 * hsel + ii;
                                if (st < G_NST - 1) {
                                    const int s = G_FI * st + i;
                                    if (s < n_run) {
                                        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
                                        if (nk > s) {
                                            const float4 *rawp = reinterpret_cast<const float4 *>(stages + (size_t)s * L.stage_floats() + L.off_raw() + (size_t)code_slot(s)[f] * 8);
                                            ra = __ldg(rawp);
                                            rb = __ldg(rawp + 1);
                                        }
                                        const float qv[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
                                        float h[8], l[8];
#pragma unroll
                                        for (int k = 0; k < 8; ++k) { h[k] = tf32_hi(qv[k]); l[k] = __fsub_rn(qv[k], h[k]); }
                                        unsigned char *at = base + i * 8192 + f * 16;
                                        *reinterpret_cast<float4 *>(at) = make_float4(h[0], h[1], h[2], h[3]);
                                        *reinterpret_cast<float4 *>(at + 2048) = make_float4(h[4], h[5], h[6], h[7]);
                                        *reinterpret_cast<float4 *>(at + 4096) = make_float4(l[0], l[1], l[2], l[3]);
                                        *reinterpret_cast<float4 *>(at + 4096 + 2048) = make_float4(l[4], l[5], l[6], l[7]);
                                    }
                                } else if (i < n_grp) {  // mask tile of stage group i: k = stage - 8 i
                                    const int r = nk - 8 * i;
                                    unsigned char *am = base + i * 4096 + f * 16;
                                    *reinterpret_cast<float4 *>(am) = make_float4(r > 0 ? 1.f : 0.f, r > 1 ? 1.f : 0.f, r > 2 ? 1.f : 0.f, r > 3 ? 1.f : 0.f);
                                    *reinterpret_cast<float4 *>(am + 2048) = make_float4(r > 4 ? 1.f : 0.f, r > 5 ? 1.f : 0.f, r > 6 ? 1.f : 0.f, r > 7 ? 1.f : 0.f);
                                }
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&bars[B_W_FULL + sa]);
                        }
                }
            }
            }  // !FC
            ph_mark(6);
        } else if (w < 12) {
            // =====================================================================================================
            // Epilogue warps: TMEM -> global.  Lane quarter q4 = w - 8, frame f = 32*q4 + lane.
            // =====================================================================================================
            if constexpr (!FC) {
                if (first_grp) {
                    // ---- keep counts, mask, kept-frame counts (quantize.py:389, utils.py:59-60): by these warps, which only have a drain
                    // every fourth chunk to do in phase L -- on the loader warps it delayed the first chunk of every tile by ~4k cycles ----
                    const int f = tid & 127;  // tile row
                    const bool row2 = SPAN && f >= ra;
                    const int bq = b + (row2 ? 1 : 0), fr = row2 ? f - ra : t0 - HALO + f;  // its item, its frame
                    const bool own = flat ? (f < ra || (SPAN && f - ra < nb2)) : (f >= HALO && f - HALO < fv);
                    int nk = 0;
                    if (own) {
                        if (p.imp != nullptr) {
                            const float lv = p.level_dev ? p.level_dev[(long long)bq * p.level_stride] : p.level_host;
                            const float x = __fmul_rn(__fmul_rn(p.imp[(long long)bq * p.imp_sb + fr], lv), (float)Nq);
                            for (int k = 0; k < n_run; ++k) nk += (__fsub_rn(x, (float)k) >= 0.0f) ? 1 : 0;
                        } else {
                            nk = n_run;
                        }
                    }
                    nkeep[f] = nk;  // (read by the frame threads after the L -> S barrier)
                    for (int k = 0; k < n_run; ++k) {
                        const bool on = nk > k;
                        const unsigned bal = __ballot_sync(0xffffffffu, on);
                        if (lane == k) kept_acc += (unsigned long long)__popc(bal);
                        if (p.mask != nullptr && own) p.mask[(long long)bq * p.mask_sb + (long long)k * p.mask_sq + fr] = on ? 1.0f : 0.0f;
                    }
                }
            }
            if constexpr (!FC)
                for (int g = 0; g < NCT / 4; ++g) drain(g, tmem + ((uint32_t)(32 * (w - 8)) << 16));
            tmem_fence_before_sync();
            __syncthreads();  // L -> S
            tmem_fence_after_sync();
            ph_mark(0);
            const int q4 = w - 8, r = 32 * q4 + lane;  // TMEM lane
            const uint32_t tq = tmem + ((uint32_t)(32 * q4) << 16);
            if constexpr (!ZQIS && !FC && VRVQ_THIRD_SCAN_GROUP) {
                // no per-stage outputs to store: these warps are the third scan group of every stage (lane r = frame row r)
                for (int s = 0; s < nl; ++s) {
                    if constexpr (!PIPE) named_bar_sync(1, NSCAN);  // 2e / e2 of the stage are visible
                    float bd;
                    int bidx;
                    scan_stage(s, 2, r, tq, bd, bidx);
                    publish_best(s, 2, r, bd, bidx);
                    named_bar_sync(1, NSCAN);
                }
            }
            // One unit = 128 channels x 128 lanes: column 32p + i of the TMEM buffer <-> channel 128j + 4i + p, lane r <-> frame
            // t0 - 8 + delta_p + r.  The tile owns, per class, the frames [t0 - 8 + delta_p, t0 + adv - 8 + delta_p) (the last tile of
            // an item up to T), so consecutive tiles cover every row exactly once and every 32-lane store starts on a sector.
            auto unit = [&](float *row0, long long rstride, uint32_t shifts) {  // row0 = (row 128j, frame 0) of the output
                // (without z_q_is: QD accumulators at columns 128 k -- all of the tensor memory is free by the final GEMM; grouped: filled
                // together, see the issuer; otherwise one after the other, so that the MMAs run up to QD units ahead of the stores)
                const uint32_t buf = D4 ? dn % QD : (dn & 1u);
                TC_WAIT(&bars[(D4 ? B_D4_FULL : B_D_FULL) + buf], (D4 ? dn / QD : (dn >> 1)) & 1u);
                tmem_fence_after_sync();
                ph_mark(1);
                const uint32_t tcol = tq + (D4 ? 0u : TM_SET) + 128u * buf;
                const uint32_t stepb = (uint32_t)(16 * rstride);  // bytes between channels 4 apart (row pitch < 2^28 floats: checked on the host)
                uint32_t va[32], vb[32];
                auto put = [&](const uint32_t (&v)[32], int piece) {
                    const int dl = (int)((shifts >> (8 * piece)) & 0xffu);
                    const bool r2 = SPAN && r >= ra;  // (spanning tile: the lane's row belongs to item b + 1; only z_q is stored then)
                    const int frame = r2 ? r - ra : t0 - HALO + dl + r;
                    const bool ok = flat ? (r < ra || (SPAN && r - ra < nb2)) : (frame >= 0 && (last_tile ? frame < p.T : r < P.adv));
                    if (ok) {
                        const unsigned long long o = reinterpret_cast<unsigned long long>(row0 + (r2 ? p.zq_sb : 0ll) + (long long)piece * rstride + frame);
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            // one independent IMAD.WIDE per address (a running 64-bit pointer costs a 4-instruction dependent chain per store)
                            unsigned long long a;
                            asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(a) : "r"(stepb), "r"((uint32_t)i), "l"(o));
                            asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(a), "f"(__uint_as_float(v[i])) : "memory");
                        }
                    }
                };
                tmem_ld32(tcol, va);
                tmem_wait_ld32(va);
                tmem_ld32(tcol + 32, vb);
                put(va, 0);
                tmem_wait_ld32(vb);
                tmem_ld32(tcol + 64, va);
                put(vb, 1);
                tmem_wait_ld32(va);
                tmem_ld32(tcol + 96, vb);
                put(va, 2);
                tmem_wait_ld32(vb);
                tmem_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[(D4 ? B_D4_EMPTY : B_D_EMPTY) + buf]);  // the MMA thread may refill this buffer
                put(vb, 3);
                ++dn;
                ph_mark(2);
            };
            if (ZQIS) {
                for (int s = 0; s < n_run; ++s) {
                    const long long off = (long long)b * p.zqis_sb + (long long)s * p.zqis_sq;
                    const uint32_t shifts = class_shifts(base_mod8(p.z_q_is, off + t0), p.zqis_sd);
                    for (int j = 0; j < NJ; ++j) unit(p.z_q_is + off + (long long)(128 * j) * p.zqis_sd, p.zqis_sd, shifts);
                }
            }
            if (p.z_q != nullptr && last_grp) {  // the final GEMM uses delta = 8 for every class: lane r <-> frame t0 + r
                float *outp = p.z_q + (long long)b * p.zq_sb;
                for (int j = 0; j < NJ; ++j) unit(outp + (long long)(128 * j) * p.zq_sd, p.zq_sd, (uint32_t)HALO * 0x01010101u);
            }
        } else if (w == 12) {
            // =====================================================================================================
            // MMA issuer: lane 0 of warp 12.
            // =====================================================================================================
            constexpr uint32_t ID_128 = umma_idesc_tf32(128, 128), ID_64 = umma_idesc_tf32(128, 64), ID_32 = umma_idesc_tf32(128, 32);
            if (!FC) issue_phase_l(0);
            ph_mark(0);
            __syncwarp();
            tmem_fence_before_sync();
            __syncthreads();  // L -> S
            tmem_fence_after_sync();
            ph_mark(1);
            {  // the whole warp runs this on warp-uniform values; one elected lane issues the MMAs and commits (see tmem_u)
                const uint64_t ones = desc128(smem_base + SM_ONES);
                uint32_t step = wn;
                auto take_w = [&]() -> uint64_t {
                    const uint32_t slot = step % W_SLOTS;
                    TC_WAIT(&bars[B_W_FULL + slot], (step / W_SLOTS) & 1u);
                    return desc128(smem_base + SM_WO + slot * W_SLOT);
                };
                auto release_w = [&]() {
                    if (elect_one()) umma_commit(&bars[B_W_EMPTY + step % W_SLOTS]);
                    ++step;
                };
                auto wait_dbuf = [&]() -> uint32_t {
                    const uint32_t buf = dn & 1u, use = dn >> 1;
                    if (use >= 1) TC_WAIT(&bars[B_D_EMPTY + buf], (use - 1) & 1u);
                    return buf;
                };
                for (int s = 0; s < nl; ++s) {
                    TC_WAIT(&bars[B_A_READY + s], apar(s));
                    fence_proxy_async();
                    tmem_fence_after_sync();
                    ph_mark(2);
                    if (ZQIS) {
                        const uint64_t ah = desc128(smem_base + SM_AT + s * 8192), al = ah + (4096 >> 4);
                        // A-operand row shift per channel class (one tile row = 16 bytes = one unit of the descriptor address)
                        const uint32_t shifts = class_shifts(base_mod8(p.z_q_is, (long long)b * p.zqis_sb + (long long)s * p.zqis_sq + t0), p.zqis_sd);
                        const bool uniform = shifts == (shifts & 0xffu) * 0x01010101u;
                        for (int j = 0; j < NJ; ++j) {
                            const uint64_t wb = take_w();  // hi tile; lo at +4096 B, bias tile at +8192 B
                            const uint32_t buf = wait_dbuf();
                            tmem_fence_after_sync();
                            const uint32_t d = tmem_u + TM_SET + 128u * buf;
                            if (elect_one()) {
                                if (uniform) {
                                    const uint64_t sh = shifts & 0xffu;
                                    umma_tf32(d, al + sh, wb, ID_128, false);
                                    umma_tf32(d, ah + sh, wb + (4096 >> 4), ID_128, true);
                                    umma_tf32(d, ones, wb + (8192 >> 4), ID_128, true);
                                    umma_tf32(d, ah + sh, wb, ID_128, true);
                                } else {
#pragma unroll
                                    for (int pc = 0; pc < 4; ++pc) {  // class pc: B rows / D columns 32pc .. 32pc+31
                                        const uint64_t sh = (shifts >> (8 * pc)) & 0xffu, wp = wb + 32 * pc;
                                        umma_tf32(d + 32 * pc, al + sh, wp, ID_32, false);
                                        umma_tf32(d + 32 * pc, ah + sh, wp + (4096 >> 4), ID_32, true);
                                        umma_tf32(d + 32 * pc, ones, wp + (8192 >> 4), ID_32, true);
                                        umma_tf32(d + 32 * pc, ah + sh, wp, ID_32, true);
                                    }
                                }
                            }
                            release_w();
                            if (elect_one()) umma_commit(&bars[B_D_FULL + buf]);
                            ++dn;
                        }
                        ph_mark(3);
                    }
                }
                if (elect_one()) umma_commit(&bars[B_MMA_DONE]);
                if (GRP && p.z_q != nullptr && last_grp) {
                    // grouped final GEMM: A tiles from the staging ring (search warps; one staging step serves QD channel chunks), W_out /
                    // bias tiles from the final ring (one ring step per (staging step, chunk)), QD accumulators at TMEM columns 128 k
                    uint32_t fstep = fn, am_ = astep;
                    for (int jh = 0; jh < NJ / QD; ++jh) {
                        for (int k = 0; k < QD; ++k) {  // the accumulators of the previous QD chunks are in the epilogue's registers
                            const uint32_t u = dn + (uint32_t)k, use = u / QD;
                            if (use >= 1) TC_WAIT(&bars[B_D4_EMPTY + u % QD], (use - 1) & 1u);
                        }
                        tmem_fence_after_sync();
                        for (int st = 0; st < G_NST; ++st, ++am_) {
                            const uint32_t sa = am_ & 1u;
                            TC_WAIT(&bars[B_W_FULL + sa], (am_ >> 1) & 1u);
                            fence_proxy_async();
                            const uint32_t abase = smem_base + SM_AT + sa * G_ASLOT;
                            for (int k = 0; k < QD; ++k, ++fstep) {
                                const uint32_t slot = fstep % F_SLOTS;
                                TC_WAIT(&bars[B_F_FULL + slot], (fstep / F_SLOTS) & 1u);
                                tmem_fence_after_sync();
                                const uint32_t wbase = smem_base + SM_WO + slot * F_SLOT;
                                const uint32_t d = tmem_u + 128u * ((dn + (uint32_t)k) % QD);
                                if (elect_one()) {
                                    if (st < G_NST - 1) {
                                        for (int i = 0; i < G_FI && G_FI * st + i < n_run; ++i) {
                                            const uint64_t ah = desc128(abase + i * 8192) + HALO, al = ah + (4096 >> 4);  // + HALO rows: skip the halo
                                            const uint64_t wb = desc128(wbase + i * 8192);
                                            umma_tf32(d, al, wb, ID_128, st > 0 || i > 0);
                                            umma_tf32(d, ah, wb + (4096 >> 4), ID_128, true);
                                            umma_tf32(d, ah, wb, ID_128, true);
                                        }
                                    } else {
                                        for (int i = 0; i < n_grp; ++i) {  // + sum_s mask_s b_out[s], eight stages per mask tile
                                            const uint64_t am = desc128(abase + i * 4096) + HALO, wb = desc128(wbase + i * 8192);
                                            umma_tf32(d, am, wb + (4096 >> 4), ID_128, true);
                                            umma_tf32(d, am, wb, ID_128, true);
                                        }
                                    }
                                    umma_commit(&bars[B_F_EMPTY + slot]);
                                    if (k == QD - 1) umma_commit(&bars[B_W_EMPTY + sa]);
                                    if (st == G_NST - 1) umma_commit(&bars[B_D4_FULL + (dn + (uint32_t)k) % QD]);
                                }
                            }
                        }
                        dn += QD;
                    }
                }
                if (!GRP && p.z_q != nullptr) {
                    TC_WAIT(&bars[B_ZQ_READY], tpar);
                    fence_proxy_async();
                    tmem_fence_after_sync();
                    ph_mark(4);
                    const uint64_t am = desc128(smem_base + SM_AM) + HALO;
                    uint32_t fstep = fn;
                    for (int j = 0; j < NJ; ++j) {
                        uint32_t buf;
                        if constexpr (D4) {
                            buf = dn % QD;
                            if (dn / QD >= 1) TC_WAIT(&bars[B_D4_EMPTY + buf], (dn / QD - 1) & 1u);
                        } else {
                            buf = wait_dbuf();
                        }
                        const uint32_t d = tmem_u + (D4 ? 0u : TM_SET) + 128u * buf;
                        for (int s0 = 0; s0 <= n_run; s0 += F_ITEMS) {  // one ring step = up to F_ITEMS chunks (stages s0.., then the bias)
                            const uint32_t slot = fstep % F_SLOTS;
                            TC_WAIT(&bars[B_F_FULL + slot], (fstep / F_SLOTS) & 1u);
                            tmem_fence_after_sync();
                            const int s1 = min(s0 + F_ITEMS, n_run + 1);
                            if (elect_one()) {
                            for (int s = s0; s < s1; ++s) {
                                const uint64_t wb = desc128(smem_base + SM_WO + slot * F_SLOT + (s - s0) * 8192);
                                if (s < n_run) {
                                    const uint64_t ah = desc128(smem_base + SM_AT + s * 8192) + HALO, al = ah + (4096 >> 4);  // + HALO rows: skip the halo
                                    umma_tf32(d, al, wb, ID_128, s > 0);
                                    umma_tf32(d, ah, wb + (4096 >> 4), ID_128, true);
                                    umma_tf32(d, ah, wb, ID_128, true);
                                } else {  // + sum_s mask_s b_out[s]
                                    umma_tf32(d, am, wb + (4096 >> 4), ID_128, true);
                                    umma_tf32(d, am, wb, ID_128, true);
                                }
                            }
                            umma_commit(&bars[B_F_EMPTY + slot]);
                            }
                            ++fstep;
                        }
                        if (elect_one()) umma_commit(&bars[(D4 ? B_D4_FULL : B_D_FULL) + buf]);
                        ++dn;
                    }
                    ph_mark(5);
                }
            }
            __syncwarp();
        } else if (w == 15) {
            // =====================================================================================================
            // Second phase-L MMA issuer: lane 0 of warp 15 (odd chunks); idle afterwards.
            // =====================================================================================================
            if (!FC) issue_phase_l(1);
            __syncwarp();
            tmem_fence_before_sync();
            __syncthreads();  // L -> S
            tmem_fence_after_sync();
        } else if (w == 14) {
            // =====================================================================================================
            // Search-MMA issuer: lane 0 of warp 14.  Per stage, NSC MMAs of M = 128 frames x N = SCW codes x K = 8 (plain TF32)
            // into three rotating SCW-column score buffers; chunk c is scanned by warps 4(c%2) .. 4(c%2)+3.
            // =====================================================================================================
            if (!FC && P.zmode != ZMODE_LDG) {
                // ---- phase L: producer of the latent staging ring (this warp has nothing else to do before the searches) ----
                const int fvz = t0 + fv - tstart;  // frames to stage per row (<= 128)
                for (int zq = 0; zq < NCH * (int)zper; ++zq) {
                    const int zc = zq / (int)zper, second = zq % (int)zper;  // (flat tiling, tile spans two items: second box = item b + 1 from its frame 0)
                    const uint32_t m = zbase + (uint32_t)zq, slot = m % Z_SLOTS, use = m / Z_SLOTS;
                    if (use >= 1) TC_WAIT(&bars[B_Z_EMPTY + slot], (use - 1) & 1u);
                    unsigned char *dst = smem + SM_ZR + slot * Z_SLOT;
                    if (P.zmode == ZMODE_TMA) {
                        if (lane == 0) {  // nc boxes [1 item][64 / nc channels of one class][136 frames]; x outside the map arrives as zeros
                            const int lg = P.znc_log2, rows = Z_CH >> lg;
                            mbar_arrive_expect_tx(&bars[B_Z_FULL + slot], Z_SLOT);
                            for (int k = 0; k < (1 << lg); ++k)
                                tma_load_3d(dst + k * rows * (Z_PITCH * 4), &zmaps.m[k], second ? -x2off : tstart, rows * zc, b + second, &bars[B_Z_FULL + slot]);
                        }
                    } else {
                        // one bulk copy per channel row, from the 16-byte aligned address at or below its first frame to the
                        // 16-byte boundary at or above its last one (<= 12 bytes of over-read on either side, inside the same
                        // 16-byte granule as a valid element); lane r copies row r of the slot
                        const float *row0 = p.z + (long long)b * p.z_sb + (long long)(Z_CH * zc) * p.z_sd + tstart;
                        const unsigned long long a = reinterpret_cast<unsigned long long>(row0 + (long long)lane * p.z_sd);
                        const unsigned long long a0 = a & ~15ull;
                        const uint32_t nb = (uint32_t)(((a + 4ull * (unsigned long long)fvz + 15ull) & ~15ull) - a0);
                        const uint32_t total = __reduce_add_sync(0xffffffffu, nb);
                        if (lane == 0) mbar_arrive_expect_tx(&bars[B_Z_FULL + slot], total);
                        __syncwarp();
                        bulk_g2s(dst + lane * (Z_PITCH * 4), reinterpret_cast<const void *>(a0), nb, &bars[B_Z_FULL + slot]);
                    }
                }
            }
            __syncwarp();
            tmem_fence_before_sync();
            __syncthreads();  // L -> S
            tmem_fence_after_sync();
            if (!FC) {  // (whole warp on warp-uniform values, one elected lane issues: see tmem_u)
                constexpr uint32_t ID_S = umma_idesc_tf32(128, SCW);
                // codebook tile: 1024 rows -> LBO = 16384 B, SBO = 128 B
                constexpr uint64_t DESC_CB = ((uint64_t)1 << 46) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)(16384 >> 4) << 16);
                const uint64_t ae = desc128(smem_base + SM_ES);
                for (int s = 0; s < nl; ++s) {
                    const uint32_t gs = gstage + (uint32_t)s;
                    // (the codebook and the first score buffer are free long before the stage's 2e tile is: those waits -- ~170 cycles
                    // each even when complete -- come first, so that only the fence stands between E_READY and the first MMA)
                    const uint32_t cbuse = (s & 1) ? (cbu1 + (uint32_t)(s >> 1)) : (cbu0 + (uint32_t)(s >> 1));
                    TC_WAIT(&bars[B_CB_FULL + (s & 1)], cbuse & 1u);
                    {
                        const uint32_t gc0 = gs * (uint32_t)NSC, use0 = gc0 / NSB;
                        if (use0 >= 1) TC_WAIT(&bars[B_SB_EMPTY + gc0 % NSB], (use0 - 1) & 1u);
                    }
                    TC_WAIT(&bars[B_E_READY], gs & 1u);
                    if (PROFILE && lane == 0) strace(21, s, it == 0);
                    fence_proxy_async();
                    const uint64_t cb = DESC_CB | (uint64_t)((smem_base + ((s & 1) ? SM_CB1 : SM_CB0)) >> 4);
                    for (int c = 0; c < NSC; ++c) {
                        const uint32_t gc = gs * (uint32_t)NSC + (uint32_t)c, sbuf = gc % NSB, use = gc / NSB;
                        if (c > 0 && use >= 1) TC_WAIT(&bars[B_SB_EMPTY + sbuf], (use - 1) & 1u);
                        tmem_fence_after_sync();
                        if (elect_one()) {
                            umma_tf32(tmem_u + TM_SC + (uint32_t)SCW * sbuf, ae, cb + (uint64_t)(c * (SCW * 16 >> 4)), ID_S, false);
                            umma_commit(&bars[B_SB_FULL + sbuf]);
                        }
                        if (PROFILE && lane == 0 && (c == 0 || c == NSC - 1)) strace(c == 0 ? 22 : 23, s, it == 0);
                    }
                }
            }
            __syncwarp();
        } else {
            // =====================================================================================================
            // Copy producer: lane 0 of warp 13 issues every cp.async.bulk (W_in ring, W_out ring, search codebooks).
            // =====================================================================================================
            if (!FC && lane == 0) {
                const float *win = P.tc + TL.off_win() + (size_t)grp * NCH * 4096;  // this group's W_in rows
                const float *gx = P.tc + TL.off_gx() + (size_t)TcLayout::gx_base(grp) * 4096 - (size_t)NCH * 4096;  // virtual chunks follow (c >= NCH)
                // codebook of the pass's first stage (its buffer is outside the phase-L region)
                mbar_arrive_expect_tx(&bars[B_CB_FULL + 0], 36864);
                bulk_g2s(smem + SM_CB0, P.tc + TL.off_cbk() + (size_t)s0 * 9216, 36864, &bars[B_CB_FULL + 0]);
                for (int c = 0; c < NCT; ++c) {
                    const uint32_t m = lbase + (uint32_t)c, slot = m % WL_SLOTS, use = m / WL_SLOTS;
                    if (use >= 1) TC_WAIT(&bars[B_WL_EMPTY + slot], (use - 1) & 1u);
                    if (PROFILE && it == 0) trace(6, c);
                    mbar_arrive_expect_tx(&bars[B_WL_FULL + slot], 16384);
                    bulk_g2s(smem + SM_LR + slot * L_SLOT, (c < NCH ? win : gx) + (size_t)c * 4096, 16384, &bars[B_WL_FULL + slot]);
                }
            }
            ph_mark(0);
            __syncwarp();
            tmem_fence_before_sync();
            __syncthreads();  // L -> S
            tmem_fence_after_sync();
            if (lane == 0) {
                if (!FC && nl > 1) {  // codebook of the pass's second stage (its buffer is inside the phase-L region)
                    mbar_arrive_expect_tx(&bars[B_CB_FULL + 1], 36864);
                    bulk_g2s(smem + SM_CB1, P.tc + TL.off_cbk() + (size_t)(s0 + 1) * 9216, 36864, &bars[B_CB_FULL + 1]);
                }
                if (!FC && !ZQIS) {  // un-normalised codebook of the pass's first stage (inside the phase-L region too)
                    mbar_arrive_expect_tx(&bars[B_RAW_FULL], 32768);
                    bulk_g2s(smem + SM_RAW, stages + (size_t)s0 * L.stage_floats() + L.off_raw(), 32768, &bars[B_RAW_FULL]);
                }
                // W_out ring of the per-stage out_proj: chunk (s, j), row-major; refills of the search codebooks interleaved
                const float *wout = P.tc + TL.off_wout();
                const float *bout = P.tc + TL.off_bout();
                int wi = 0, cs = FC ? nl : 0;  // (from_codes: no search codebooks to refill)
                uint32_t spins = 0;
                unsigned long long spin_t0 = 0;
                while (wi < n_stage_steps || cs < nl) {
                    bool progressed = false;
                    if (wi < n_stage_steps) {
                        const uint32_t m = wn + (uint32_t)wi, slot = m % W_SLOTS, use = m / W_SLOTS;
                        if (use == 0 || mbar_try_wait(&bars[B_W_EMPTY + slot], (use - 1) & 1u)) {
                            mbar_arrive_expect_tx(&bars[B_W_FULL + slot], 12288);
                            bulk_g2s(smem + SM_WO + slot * W_SLOT, wout + (size_t)wi * 3072, 12288, &bars[B_W_FULL + slot]);
                            ++wi;
                            progressed = true;
                        }
                    }
                    if (cs < nl && mbar_try_wait(&bars[B_A_READY + cs], apar(cs))) {
                        // the search group is done with the codebook of stage cs: refill its buffer
                        if (cs + 2 < nl) {
                            mbar_arrive_expect_tx(&bars[B_CB_FULL + (cs & 1)], 36864);
                            bulk_g2s(smem + ((cs & 1) ? SM_CB1 : SM_CB0), P.tc + TL.off_cbk() + (size_t)(s0 + cs + 2) * 9216, 36864,
                                     &bars[B_CB_FULL + (cs & 1)]);
                        }
                        if (!ZQIS && !FC && cs + 1 < nl) {  // ... and with the un-normalised rows of stage cs (gathered in its merge)
                            mbar_arrive_expect_tx(&bars[B_RAW_FULL], 32768);
                            bulk_g2s(smem + SM_RAW, stages + (size_t)(s0 + cs + 1) * L.stage_floats() + L.off_raw(), 32768, &bars[B_RAW_FULL]);
                        }
                        ++cs;
                        progressed = true;
                    }
                    if (progressed) { spins = 0; spin_t0 = 0; }
                    else if ((++spins & 0x3fffu) == 0) spin_t0 = tc_wait_check(spin_t0, &bars[B_W_EMPTY], bars, 99);
                }
                // final GEMM ring: for j: W_out (0..n_run-1, j) then the bias chunk j.  Its slots overlay the W_out ring and the
                // codebook buffers, so it starts once every per-stage MMA has completed (and the last search is over).
                if (GRP && n_final_steps > 0) {  // grouped: four stages per step (W_out hi | lo tiles), last step = the groups' bias tiles
                    TC_WAIT(&bars[B_MMA_DONE], tpar);
                    uint32_t m = fn;
                    for (int jq = 0; jq < (NJ / QD) * G_NST * QD; ++jq, ++m) {  // order of the issuer: (chunk quad, staging step, chunk of the quad)
                            const int st = (jq / QD) % G_NST, j = (jq / (QD * G_NST)) * QD + jq % QD;
                            const uint32_t slot = m % F_SLOTS, use = m / F_SLOTS;
                            if (use >= 1) TC_WAIT(&bars[B_F_EMPTY + slot], (use - 1) & 1u);
                            const int cnt = st < G_NST - 1 ? min(G_FI, n_run - G_FI * st) : n_grp;
                            mbar_arrive_expect_tx(&bars[B_F_FULL + slot], (uint32_t)cnt * 8192u);
                            for (int i = 0; i < cnt; ++i) {
                                const float *src = st < G_NST - 1 ? wout + ((size_t)(G_FI * st + i) * NJ + j) * 3072 : bout + ((size_t)i * NJ + j) * 2048;
                                bulk_g2s(smem + SM_WO + slot * F_SLOT + i * 8192, src, 8192, &bars[B_F_FULL + slot]);
                            }
                        }
                }
                if (!GRP && n_final_steps > 0) {
                    TC_WAIT(&bars[B_MMA_DONE], tpar);
                    uint32_t m = fn;
                    for (int j = 0; j < NJ; ++j)
                        for (int s0 = 0; s0 <= n_run; s0 += F_ITEMS, ++m) {
                            const uint32_t slot = m % F_SLOTS, use = m / F_SLOTS;
                            if (use >= 1) TC_WAIT(&bars[B_F_EMPTY + slot], (use - 1) & 1u);
                            const int s1 = min(s0 + F_ITEMS, n_run + 1);
                            mbar_arrive_expect_tx(&bars[B_F_FULL + slot], (uint32_t)(s1 - s0) * 8192u);
                            for (int sx = s0; sx < s1; ++sx) {
                                const float *src = sx < n_run ? wout + ((size_t)sx * NJ + j) * 3072 : bout + (size_t)j * 2048;
                                bulk_g2s(smem + SM_WO + slot * F_SLOT + (sx - s0) * 8192, src, 8192, &bars[B_F_FULL + slot]);
                            }
                        }
                }
            }
            ph_mark(1);
            __syncwarp();
        }
        wn += (uint32_t)n_stage_steps;
        fn += (uint32_t)n_final_steps;
        if (GRP) astep += (uint32_t)n_astage_steps;
        gstage += (uint32_t)nl;
        cbu0 += (uint32_t)((nl + 1) >> 1);  // buffer 0 serves the even stages, buffer 1 the odd ones
        cbu1 += (uint32_t)(nl >> 1);
        lbase += (uint32_t)NCT;
        gbase += (uint32_t)(NCT / 4);
        zbase += (uint32_t)NCH * zper;
        tmem_fence_before_sync();
        __syncthreads();  // end of tile: every MMA of the tile has completed (the epilogue waited for the last one)
        tmem_fence_after_sync();
        ph_mark(7);
        if (lite) {  // [0] += time to the L->S barrier, [1] += whole tile
            const long long t = clock64();
            p.phase_cycles[(size_t)blockIdx.x * 64 + 0] += lite_l;
            p.phase_cycles[(size_t)blockIdx.x * 64 + 1] += t - lite_t0;
            lite_t0 = t;
            lite_l = 0;
        }
        };  // pass_body
        if (FLAT && nb2 > 0) pass_body(std::true_type{});
        else pass_body(std::false_type{});
    }
    if (PROFILE && ph_on && !P.trace)
        for (int k = 0; k < 16; ++k) p.phase_cycles[((size_t)blockIdx.x * 4 + ph_role) * 16 + k] = ph_acc[k];

    // ---- teardown ----
    if (w < 4) {
        // loss: one binary64 atomic per warp; kept counts: one atomic per stage per warp
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
        if (lane == 0 && p.loss_sum != nullptr && loss_acc != 0.0) atomicAdd(p.loss_sum, loss_acc);
    }
    if (w >= 8 && w < 12 && lane < n_run && p.kept != nullptr && kept_acc != 0ull) atomicAdd(&p.kept[lane], kept_acc);
    if (w == 12) tmem_dealloc(tmem, TM_COLS);
}

// ---- host side -----------------------------------------------------------------------------------
static int pick_tiling(int B, int T, int sms, bool zqis, int *adv, int *tiles_per_b) {
    // tiles of `adv` frames (multiple of 8, <= 120: 8 of the 128 rows are the halo) per batch item; minimise waves x time per tile.
    // With z_q_is a tile's time is its stores, proportional to its frames (+ a fixed part: pipeline fill, first search); without,
    // it is the in_proj stream and the serial stage chain, the same for any number of frames (measured: 85.9 us per wave at 104,
    // 112 and 120 frames on the config-4 shape), so the fewest waves win and the frame count only breaks ties.
    const int amax = zqis ? 120 : 128;  // own frames per tile: 128 rows minus the 8-row halo that only the z_q_is stores need
    const int nt_min = (T + amax - 1) / amax;
    const long fixed = zqis ? 24 : 1000;
    long best_cost = -1;
    int best_nt = nt_min, best_adv = amax;
    for (int nt = nt_min; nt <= nt_min * 4 + 4; ++nt) {
        int a = ((T + nt - 1) / nt + 7) / 8 * 8;
        if (a > amax) continue;
        if (a < 8) a = 8;
        const int nt_eff = (T + a - 1) / a;
        const long tiles = (long)B * nt_eff;
        const long waves = (tiles + sms - 1) / sms;
        const long cost = waves * (a + fixed);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nt = nt_eff; best_adv = a; }
    }
    *adv = best_adv;
    *tiles_per_b = best_nt;
    return 0;
}

int encode_tc_usable(const vrvq_encode_args *a) {
    if (!tc_shape_ok(a->input_dim, a->codebook_size, a->n_codebooks)) return 0;
    if (a->n_codebooks > TC_MAX_NQ_ZQIS && a->z_q_is != nullptr) return 0;  // more than one stage group: no per-stage outputs on this path
    // the epilogue forms store addresses as base + i * (16 * row pitch) with a 32-bit step
    const long long lim = 1ll << 27;
    if (a->z_q_is != nullptr && (a->z_q_is_stride_d < 0 || a->z_q_is_stride_d >= lim)) return 0;
    if (a->z_q != nullptr && (a->z_q_stride_d < 0 || a->z_q_stride_d >= lim)) return 0;
    return 1;
}

// How phase L fetches the latent.  Default: TMA tensor maps for EVERY layout (channel-class maps, tmaps.cuh).
// VRVQ_LATENT_LOAD=ldg|bulk|tma overrides (ldg = per-thread loads, the round-1 path; bulk = one cp.async.bulk per channel row; both
// kept for A/B runs).
static int pick_zmode(const vrvq_encode_args *a, ZMaps *maps, TcParams &P) {
    memset(maps, 0, sizeof(*maps));
    P.znc_log2 = 0;
    for (int k = 0; k < 4; ++k) P.zshift[k] = 0;
    int want = ZMODE_TMA;
    if (const char *env = getenv("VRVQ_LATENT_LOAD")) {
        if (env[0] == 'l') return ZMODE_LDG;
        if (env[0] == 'b') want = ZMODE_BULK;
    }
    if (a->z_stride_d <= 0 || a->z_stride_b < 0) return ZMODE_LDG;
    if (want == ZMODE_TMA && build_class_maps(a->z, a->B, a->input_dim, a->T, a->z_stride_d, a->z_stride_b, Z_PITCH, Z_CH, maps, &P.znc_log2, P.zshift))
        return ZMODE_TMA;
    return ZMODE_BULK;
}

static int make_params(const vrvq_encode_args *a, const EncodeParams &e, TcParams &P, int *grid, ZMaps *zmap) {
    P.e = e;
    const BlobLayout L(a->input_dim, a->codebook_size);
    const size_t tc_off = (size_t)BLOB_HDR_FLOATS + (size_t)a->n_codebooks * (size_t)L.stage_floats();
    P.tc = static_cast<const float *>(a->blob) + tc_off;
    const int sms = current_sm_count();
    if (sms <= 0) {
        set_error("cannot query the SM count of the current device");
        return VRVQ_ECUDA;
    }
    pick_tiling(a->B, a->T, sms, a->z_q_is != nullptr, &P.adv, &P.tiles_per_b);
    if (const char *dbg = getenv("VRVQ_DEBUG_TILE_FRAMES")) {  // profiling knob
        const int v = atoi(dbg);
        if (v >= 8 && v <= (a->z_q_is != nullptr ? 120 : 128) && v % 8 == 0) { P.adv = v; P.tiles_per_b = (a->T + v - 1) / v; }
    }
    P.n_tiles = P.tiles_per_b * a->B;
    P.zmode = pick_zmode(a, zmap, P);
    // Flat tiling: without z_q_is a tile costs the same whatever it holds, so only the number of waves counts, and items whose length is not
    // a multiple of 128 waste the rest of their last tile (config 3: 64 x 7 = 448 tiles = 3.03 waves of 148).  Tiles of 128 consecutive frames
    // of the flattened (item, frame) sequence -- a tile may then span two items, hence T >= 128 -- need ceil(B T / 128) (431 = 3 waves).
    // Taken when it saves a wave; the latent of a spanning tile comes through two TMA boxes, so only on the TMA path.
    P.flat = 0;
    const char *dbgp = getenv("VRVQ_DEBUG_PHASES");
    if (a->z_q_is == nullptr && a->T >= 128 && P.zmode == ZMODE_TMA && !(dbgp != nullptr && dbgp[0] != '2')) {
        const long long flat_tiles = ((long long)a->B * a->T + 127) / 128;
        const long long waves_now = ((long long)P.n_tiles + sms - 1) / sms, waves_flat = (flat_tiles + sms - 1) / sms;
        const char *env = getenv("VRVQ_FLAT_TILES");
        // (a pass of the flat instantiation costs ~10 % more -- per-row items, the permuted order, two boxes for the spanning tiles -- so it
        // has to save more than that: config 3 4 -> 3 waves; not all of config 4 on one GPU, 71 -> 70 waves: 5230 vs 5817 us)
        if (env ? env[0] == '1' : (waves_flat * 112 < waves_now * 100 && P.n_tiles > sms)) {
            P.flat = 1;
            P.adv = 128;
            P.tiles_per_b = (a->T + 127) / 128;
            P.n_tiles = (int)flat_tiles;
        }
    }
    P.single_issuer = getenv("VRVQ_DEBUG_SINGLE_ISSUER") != nullptr;
    P.stagger = getenv("VRVQ_DEBUG_STAGGER") ? atoi(getenv("VRVQ_DEBUG_STAGGER")) : 0;
    *grid = P.n_tiles < sms ? P.n_tiles : sms;
    return VRVQ_OK;
}

template <int D, bool ZQIS, bool PROFILE, bool FC = false, bool GRP = false, bool FLAT = false>
static int launch_tc_one(const TcParams &P, const ZMaps &zmap, int grid, cudaStream_t st) {
    constexpr int SMEM = GRP ? SM_TOTAL_G : SM_TOTAL;
    if constexpr (!FLAT && !ZQIS && !FC && !PROFILE) {  // flat tiling is its own instantiation
        if (P.flat) return launch_tc_one<D, ZQIS, PROFILE, FC, GRP, true>(P, zmap, grid, st);
    }
    int rc = ensure_dynamic_smem<rvq_encode_tc_kernel<D, ZQIS, PROFILE, FC, GRP, FLAT>>(SMEM, "cudaFuncSetAttribute(rvq_encode_tc_kernel)");
    if (rc) return rc;
    rvq_encode_tc_kernel<D, ZQIS, PROFILE, FC, GRP, FLAT><<<grid, TC_NTH, SMEM, st>>>(P, zmap);
    return check_cuda(cudaGetLastError(), "rvq_encode_tc_kernel launch");
}
template <int D, bool ZQIS>
static int launch_tc(const TcParams &P, const ZMaps &zmap, int grid, cudaStream_t st) {
    // the phase-counter build (VRVQ_DEBUG_PHASES=1) is a separate instantiation: the production kernel carries no counters
    const char *dbg = getenv("VRVQ_DEBUG_PHASES");
    const bool full = P.e.phase_cycles != nullptr && !(dbg != nullptr && dbg[0] == '2');  // "2": production code + 3 timestamps per CTA; "3": per-chunk trace
    return full ? launch_tc_one<D, ZQIS, true>(P, zmap, grid, st) : launch_tc_one<D, ZQIS, false>(P, zmap, grid, st);
}

int encode_tc_launch_info(const vrvq_encode_args *a, int *grid, int *block, int *smem) {
    EncodeParams e{};
    int rc = fill_encode_params(a, e);
    if (rc) return rc;
    TcParams P{};
    ZMaps zmap;
    int g = 0;
    rc = make_params(a, e, P, &g, &zmap);
    if (rc) return rc;
    if (grid) *grid = g;
    if (block) *block = TC_NTH;
    if (smem) *smem = a->n_codebooks > TC_MAX_NQ_ZQIS ? SM_TOTAL_G : SM_TOTAL;
    return VRVQ_OK;
}

int encode_tc(const vrvq_encode_args *a, const EncodeParams &e, void *stream) {
    TcParams P{};
    ZMaps zmap;
    int grid = 0;
    int rc = make_params(a, e, P, &grid, &zmap);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool zqis = a->z_q_is != nullptr;
    const bool dbg = getenv("VRVQ_DEBUG_PHASES") != nullptr;  // profiling only: synchronises and prints per-phase cycles
    P.e.phase_cycles = nullptr;
    if (dbg && cudaMalloc(&P.e.phase_cycles, sizeof(long long) * 64 * (size_t)grid) != cudaSuccess) P.e.phase_cycles = nullptr;
    if (P.e.phase_cycles != nullptr) cudaMemsetAsync(P.e.phase_cycles, 0, sizeof(long long) * 64 * (size_t)grid, st);
    P.trace = (dbg && getenv("VRVQ_DEBUG_PHASES")[0] == '3' && grid >= 32) ? 1 : 0;
    if (a->n_codebooks > TC_MAX_NQ_ZQIS) {  // stage groups (no profiling instantiation)
        switch (a->input_dim) {
            case 1024: return launch_tc_one<1024, false, false, false, true>(P, zmap, grid, st);
            case 512: return launch_tc_one<512, false, false, false, true>(P, zmap, grid, st);
            case 256: return launch_tc_one<256, false, false, false, true>(P, zmap, grid, st);
            default: return VRVQ_EUNSUPPORTED;
        }
    }
    switch (a->input_dim) {
        case 1024: rc = zqis ? launch_tc<1024, true>(P, zmap, grid, st) : launch_tc<1024, false>(P, zmap, grid, st); break;
        case 512: rc = zqis ? launch_tc<512, true>(P, zmap, grid, st) : launch_tc<512, false>(P, zmap, grid, st); break;
        case 256: rc = zqis ? launch_tc<256, true>(P, zmap, grid, st) : launch_tc<256, false>(P, zmap, grid, st); break;
        default: rc = VRVQ_EUNSUPPORTED;
    }
    if (dbg && P.e.phase_cycles != nullptr) {
        static const char *names[4][16] = {
            {"setup", "phaseL", "L2S_wait", "prep", "search_tail", "merge_corr", "mask_pass", "end_wait", "scan:wait_full", "scan:ld", "scan:filter", "scan:rescore", "#cnt", "#nc", "#warpmaxcnt", "-"},
            {"L_idle", "wait_full", "store", "-", "-", "-", "-", "end_wait", "-", "-", "-", "-", "-", "-", "-", "-"},
            {"phaseL", "L2S_wait", "wait_ready", "stage_units", "wait_zq", "final", "-", "end_wait", "-", "-", "-", "-", "-", "-", "-", "-"},
            {"phaseL", "phaseS", "-", "-", "-", "-", "-", "end_wait", "-", "-", "-", "-", "-", "-", "-", "-"}};
        static const char *roles[4] = {"search", "epilogue", "issuer", "producer"};
        long long *h = static_cast<long long *>(malloc(sizeof(long long) * 64 * (size_t)grid));
        cudaStreamSynchronize(st);
        cudaMemcpy(h, P.e.phase_cycles, sizeof(long long) * 64 * (size_t)grid, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[vrvq tc phases] grid %d, tile %d frames, %d tiles; mean cycles per CTA\n", grid, P.adv, P.n_tiles);
        if (P.trace) {
            static const char *ev[8] = {"z_full", "split", "a_empty", "a_handed", "iss_saw", "iss_done", "w_empty", "drain"};
            long long t0 = 0;
            for (int e = 0; e < 7; ++e)
                for (int c = 0; c < 32; ++c) { const long long v = h[1024 + 32 * e + c]; if (v != 0 && (t0 == 0 || v < t0)) t0 = v; }
            fprintf(stderr, "[vrvq tc trace] block 0, first tile, cycles since the first event; chunk:");
            for (int e = 0; e < 8; ++e) fprintf(stderr, " %s", ev[e]);
            fprintf(stderr, "\n");
            for (int c = 0; c < 32; ++c) {
                fprintf(stderr, "  %2d:", c);
                for (int e = 0; e < 8; ++e) { const long long v = h[1024 + 32 * e + c]; fprintf(stderr, " %7lld", v ? v - t0 : -1); }
                fprintf(stderr, "\n");
            }
            {
                static const char *sev[16] = {"top", "g0_scores", "g0_scanned", "g0_rescored", "g0_postbar", "g0_q", "g0_prepped", "g0_end",
                                              "g1_q_seen", "g1_corrected", "g1_scores", "g1_scanned", "g1_rescored", "iss_e_ready", "iss_first", "iss_last"};
                long long s0t = 0;
                for (int e = 8; e < 24; ++e)
                    for (int c = 0; c < 8; ++c) { const long long v = h[1024 + 32 * e + c]; if (v != 0 && (s0t == 0 || v < s0t)) s0t = v; }
                fprintf(stderr, "[vrvq tc trace] block 0, first pass, stage time line (cycles since its first event); stage:");
                for (int e = 0; e < 16; ++e) fprintf(stderr, " %s", sev[e]);
                fprintf(stderr, "\n");
                for (int c = 0; c < 8; ++c) {
                    fprintf(stderr, "  %2d:", c);
                    for (int e = 8; e < 24; ++e) { const long long v = h[1024 + 32 * e + c]; fprintf(stderr, " %7lld", v ? v - s0t : -1); }
                    fprintf(stderr, "\n");
                }
            }
            free(h);
            cudaFree(P.e.phase_cycles);
            return rc;
        }
        if (getenv("VRVQ_DEBUG_PHASES")[0] == '2') {
            double a0 = 0, a1 = 0;
            for (int g = 0; g < grid; ++g) { a0 += (double)h[g * 64 + 0] / grid; a1 += (double)h[g * 64 + 1] / grid; }
            fprintf(stderr, "  production instantiation: to the L->S barrier %.0f, whole tiles %.0f\n", a0, a1);
            free(h);
            cudaFree(P.e.phase_cycles);
            return rc;
        }
        for (int r = 0; r < 4; ++r) {
            double acc[16] = {0}, tot = 0;
            for (int g = 0; g < grid; ++g)
                for (int k = 0; k < 16; ++k) { acc[k] += (double)h[(g * 4 + r) * 16 + k] / grid; if (k < 12) tot += (double)h[(g * 4 + r) * 16 + k] / grid; }
            fprintf(stderr, "  %-8s total %.0f |", roles[r], tot);
            for (int k = 0; k < 16; ++k)
                if (names[r][k][0] != '-') fprintf(stderr, " %s %.0f", names[r][k], acc[k]);
            fprintf(stderr, "\n");
        }
        free(h);
        cudaFree(P.e.phase_cycles);
    }
    return rc;
}


// ---- from_codes on the tensor-core path (vrvq_from_codes_f32 for models with <= 8 codebooks) ------------------------
int from_codes_tc_usable(const vrvq_from_codes_args *a) {
    const long long lim = 1ll << 27;
    if (!tc_shape_ok(a->input_dim, a->codebook_size, a->n_codebooks) || a->n_codebooks > TC_MAX_NQ_ZQIS) return 0;
    if (a->z_q_stride_d < 0 || a->z_q_stride_d >= lim) return 0;
    if (a->z_q_is != nullptr && (a->z_q_is_stride_d < 0 || a->z_q_is_stride_d >= lim)) return 0;
    const char *impl = getenv("VRVQ_ENCODE_IMPL");
    if (impl != nullptr && impl[0] == 'c') return 0;
    return prefer_tc_for_size(a->B, a->T) ? 1 : 0;
}

int from_codes_tc(const vrvq_from_codes_args *a, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TcParams P{};
    EncodeParams &e = P.e;
    e.blob = static_cast<const float *>(a->blob);
    e.B = a->B; e.T = a->T; e.Nq = a->n_codebooks; e.n_run = a->n_run;
    e.z_q = a->z_q; e.zq_sb = a->z_q_stride_b; e.zq_sd = a->z_q_stride_d;
    e.z_q_is = a->z_q_is; e.zqis_sb = a->z_q_is_stride_b; e.zqis_sq = a->z_q_is_stride_q; e.zqis_sd = a->z_q_is_stride_d;
    e.latents = a->z_p; e.lat_sb = a->z_p_stride_b; e.lat_sc = a->z_p_stride_c;  // z_p = the gathered rows, latents layout
    const BlobLayout L(a->input_dim, a->codebook_size);
    P.tc = e.blob + (size_t)BLOB_HDR_FLOATS + (size_t)a->n_codebooks * (size_t)L.stage_floats();
    P.codes_in = reinterpret_cast<const long long *>(a->codes); P.cin_sb = a->codes_stride_b; P.cin_sq = a->codes_stride_q;
    P.mask_in = a->mask; P.min_sb = a->mask_stride_b; P.min_sq = a->mask_stride_q;
    P.error_flag = a->error_flag;
    const int sms = current_sm_count();
    if (sms <= 0) {
        set_error("cannot query the SM count of the current device");
        return VRVQ_ECUDA;
    }
    pick_tiling(a->B, a->T, sms, a->z_q_is != nullptr, &P.adv, &P.tiles_per_b);
    P.n_tiles = P.tiles_per_b * a->B;
    P.zmode = ZMODE_LDG;  // no latent on this path
    ZMaps zmap;
    memset(&zmap, 0, sizeof(zmap));
    const int grid = P.n_tiles < sms ? P.n_tiles : sms;
    const bool zqis = a->z_q_is != nullptr;
    switch (a->input_dim) {
        case 1024: return zqis ? launch_tc_one<1024, true, false, true>(P, zmap, grid, st) : launch_tc_one<1024, false, false, true>(P, zmap, grid, st);
        case 512: return zqis ? launch_tc_one<512, true, false, true>(P, zmap, grid, st) : launch_tc_one<512, false, false, true>(P, zmap, grid, st);
        case 256: return zqis ? launch_tc_one<256, true, false, true>(P, zmap, grid, st) : launch_tc_one<256, false, false, true>(P, zmap, grid, st);
        default: return VRVQ_EUNSUPPORTED;
    }
}

}  // namespace vrvq
