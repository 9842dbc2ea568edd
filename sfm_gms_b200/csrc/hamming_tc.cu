// hamming_tc.cu — brute-force Hamming nearest neighbour on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Replaces cv::BFMatcher(NORM_HAMMING)::match (reference call site FeatureMatchUtil.cpp:68).
//
// Idea: unpack each 256-bit descriptor to 256 int8 values in {-1,+1}.  Then for query a and train b
//     <a,b> = 256 - 2*hamming(a,b)      (exact in the s32 accumulator)
// so the nearest neighbour is the arg-MAX of an int8 GEMM row, lowest column on ties.  The GEMM runs as
// tcgen05.mma.kind::i8 (M=128, N=128, K=32 per instruction) with the accumulators in TMEM; the epilogue
// never materialises the N1 x N2 distance matrix: it reduces each accumulator row to (max, first argmax)
// straight out of TMEM and merges across CTAs with one atomicMin on the packed key (dist<<18 | trainIdx),
// which is order-independent => deterministic and bit-identical to OpenCV's strict-'<' scan.
//
// Operand placement (this is what the roofline hinges on): an SS-mode 128x128x32 MMA reads 4 KB of A and
// 4 KB of B from shared memory every 64 clk = 128 B/clk, the whole shared-memory bandwidth of an SM, and
// measured at only ~60 % of the tensor peak.  So the QUERY tile is stationary in TENSOR MEMORY (TS-mode MMA:
// A from TMEM), written there once per work unit by tcgen05.st straight from the packed 32-byte descriptors,
// and only the TRAIN tile streams through shared memory (TMA, 128B swizzle): 64 B/clk of MMA reads.
//
// Kernel structure (persistent, warp-specialised, 1 CTA/SM, 512 threads):
//   warp 0      TMA producer: train tiles [128 rows x 256 B] through a 6-stage ring (cp.async.bulk.tensor.2d)
//   warp 1      MMA issuer: one lane issues tcgen05.mma; tcgen05.commit signals the mbarriers
//   warp 2      TMEM allocator (512 columns = MSUB x 64 query-operand columns + accumulator slots of 128)
//   warps 4-11  epilogue: tcgen05.ld 32 lanes x 32 columns at a time; a chunk max (cheap) guards the
//               rare "new row maximum" path that extracts the first column attaining it
//   warps 12-15 query loaders: packed descriptor row -> +-1 bytes in registers -> tcgen05.st into TMEM
// Work unit = (pair, MSUB*128 query rows, contiguous range of 128-row train tiles).
//
// Roofline: tensor pipe.  Algorithmic work = 2*256 int8 OPs per distance.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "hamming_tc.cuh"
#include "tc_ptx.cuh"

namespace sfmgms {

namespace {

constexpr int BM = 128;            // UMMA M: query rows per sub-tile (TMEM lanes)
#ifndef SFMGMS_TC_MSUB
#define SFMGMS_TC_MSUB 2
#endif
constexpr int MSUB = SFMGMS_TC_MSUB;   // query sub-tiles per work unit (stationary in TMEM)
constexpr int BN = 128;            // UMMA N: train rows per B tile (accumulator columns)
constexpr int KBYTES = 256;        // unpacked descriptor: 256 x int8
constexpr int KCH = 128;           // bytes per 128B-swizzle chunk
constexpr int NKCH = KBYTES / KCH; // 2
constexpr int UMMA_K = 32;         // K per tcgen05.mma for 8-bit operands
constexpr int STAGES = 6;          // B ring (32 KB per stage)
constexpr int A_COLS = MSUB * (KBYTES / 4);      // TMEM columns of the query operand (64 per sub-tile)
constexpr int ACC_SLOTS = (512 - A_COLS) / BN;   // accumulator slots of 128 columns
constexpr int TILE_BYTES = BN * KCH;             // 16 KB: one [128 rows x 128 B] swizzled chunk
constexpr int B_STAGE_BYTES = NKCH * TILE_BYTES; // 32 KB
constexpr int SMEM_DATA = STAGES * B_STAGE_BYTES;  // 192 KB
constexpr int SMEM_BYTES = SMEM_DATA + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 512;
constexpr int kEpiWarp0 = 4;
constexpr int kEpiWarps = 8;
constexpr int kLoadWarp0 = 12;
constexpr int kLoadWarps = 4;
static_assert(ACC_SLOTS >= 2, "need at least two accumulator slots");

struct alignas(16) WorkUnit {
    const uint8_t* a_packed;  // packed (32 B/row) descriptors of this unit's first query row
    uint32_t* key;            // &key[first query row of the unit]
    uint32_t b_row0;          // row of the train image's first descriptor in the unpacked operand array
    int32_t n_rows;           // valid query rows in this unit (<= MSUB*128)
    int32_t t_begin;          // first train row (multiple of BN)
    int32_t t_end;            // one past the last train row of this unit (<= n2)
};

using namespace tcptx;

// TS-mode MMA: D[tmem] (+)= A[tmem] * B[smem descriptor].  The B descriptor is passed as (lo, hi) halves:
// hi is one constant for every tile of this kernel, lo = (address >> 4) | LBO field, so stepping K or
// switching stage is a 32-bit add.
template <int kAccumulate>
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t desc_hi,
                                             uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .b64 db;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], db, %4, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(kAccumulate) : "memory");
}
// K-major, 128B-swizzled operand tile [rows x 128 B], rows contiguous (8-row groups 1024 B apart).
// Descriptor fields (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=SWIZZLE_128B(2) [61,64).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4)      // SBO: 8 rows x 128 B
                             | (1u << 14)               // descriptor version (Blackwell)
                             | (2u << 29);              // SWIZZLE_128B
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) {
    return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);  // start address | LBO (ignored for swizzled K-major)
}
// Instruction descriptor (InstrDescriptor): c_format=S32(2)@4, a_format=INT8(1)@7, b_format=INT8(1)@10,
// a_major=K(0)@15, b_major=K(0)@16, n_dim=N>>3 @17, m_dim=M>>4 @24.
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// 4 descriptor bits -> 4 bytes of +-1 (bit b of the nibble -> byte b): 0x01 where set, 0xFF (-1) where clear
__device__ __forceinline__ uint32_t pm1_from_nibble(uint32_t nib) {
    const uint32_t m = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
    return ~(m * 0xFFu) | m;
}

// ---- operand unpack for the TRAIN side: 256 bits -> 256 x int8 in {-1,+1} (K index = bit index) -----------
// thread = one 32-bit word of a descriptor -> 32 output bytes (2 x 128-bit stores)
__global__ void __launch_bounds__(256) unpack_pm1_kernel(const uint32_t* __restrict__ desc, long long n_words,
                                                         uint4* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n_words; i += step) {
        const uint32_t w = __ldg(desc + i);
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = pm1_from_nibble((w >> (4 * k)) & 0xFu);
        out[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
        out[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

// Pulls a small table out of pinned (UVA-mapped) host memory with SM loads instead of a copy-engine H2D copy:
// the copy engine serves requests in submission order, so a table upload issued while a large image-set
// transfer is queued (sfmgms_match_image_set) would wait behind all of it and stall the compute stream.
__global__ void __launch_bounds__(256) pull_table_kernel(const uint4* __restrict__ host_src, uint4* __restrict__ dst, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = host_src[i];
}

// ---- the tcgen05 kernel -------------------------------------------------------------------------------
template <int dbg>   // dbg != 0: timing experiments only (results invalid); production is dbg = 0
__global__ void __launch_bounds__(kThreads, 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap tmap, const WorkUnit* __restrict__ units, int n_units) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // 128B swizzle needs 1024 B alignment
    const uint32_t b_smem = smem_base;
    const uint32_t bar_base = smem_base + SMEM_DATA;
    const uint32_t full_bar = bar_base;                       // [STAGES]  TMA -> MMA
    const uint32_t empty_bar = full_bar + 8 * STAGES;          // [STAGES]  MMA -> TMA
    const uint32_t a_full_bar = empty_bar + 8 * STAGES;        // loaders -> MMA (query operand in TMEM)
    const uint32_t a_empty_bar = a_full_bar + 8;               // MMA -> loaders
    const uint32_t tfull_bar = a_empty_bar + 8;                // [ACC_SLOTS] MMA -> epilogue
    const uint32_t tempty_bar = tfull_bar + 8 * ACC_SLOTS;     // [ACC_SLOTS] epilogue -> MMA
    const uint32_t tmem_ptr_smem = tempty_bar + 8 * ACC_SLOTS; // 4 B: TMEM base address
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(a_full_bar, kLoadWarps);
        mbar_init(a_empty_bar, 1);
        for (int s = 0; s < ACC_SLOTS; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;
    const uint32_t tmem_acc = tmem_base + A_COLS;             // accumulator slots follow the query operand

    if (warp == 0) {
        // ================= TMA producer: train tiles (warp converged, one elected lane issues) =================
        {
            if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            uint32_t stage = 0, phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const WorkUnit wu = units[u];
                for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)B_STAGE_BYTES);
                        for (int kc = 0; kc < NKCH; ++kc)
                            tma_load_2d(b_smem + stage * B_STAGE_BYTES + kc * TILE_BYTES, &tmap, kc * KCH,
                                        (int)(wu.b_row0 + t), full_bar + 8 * stage);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (warp converged, one elected lane issues) =================
        {
            uint32_t stage = 0, phase = 0, a_phase = 0, slot = 0, slot_phase = 0;
            const uint32_t b_lo0 = sdesc_lo(b_smem);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const WorkUnit wu = units[u];
                const int nsub = (wu.n_rows + BM - 1) / BM;
                const int ntiles = (wu.t_end - wu.t_begin + BN - 1) / BN;
                mbar_wait(a_full_bar, a_phase);                             // query operand is in TMEM
                a_phase ^= 1;
                tc_fence_after();
                for (int ti = 0; ti < ntiles; ++ti) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    const uint32_t b_lo = b_lo0 + stage * (B_STAGE_BYTES >> 4);
                    for (int s = 0; s < nsub; ++s) {
                        mbar_wait(tempty_bar + 8 * slot, slot_phase ^ 1);   // epilogue drained this accumulator
                        tc_fence_after();
                        const uint32_t d = tmem_acc + slot * BN;
                        const uint32_t a = tmem_base + s * (KBYTES / 4);    // 8 TMEM columns per K step of 32 bytes
                        // K = 256 = 2 swizzle chunks x 4 steps of 32 (one 128x128x32 int8 MMA each)
                        if (elect_one()) {
                        if (dbg != 3) {   // DEBUG 3: no tensor work
                        tc_mma_i8_ts<0>(d, a + 0, b_lo + 0, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 8, b_lo + 2, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 16, b_lo + 4, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 24, b_lo + 6, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 32, b_lo + (TILE_BYTES >> 4) + 0, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 40, b_lo + (TILE_BYTES >> 4) + 2, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 48, b_lo + (TILE_BYTES >> 4) + 4, kDescHi, kIdesc);
                        tc_mma_i8_ts<1>(d, a + 56, b_lo + (TILE_BYTES >> 4) + 6, kDescHi, kIdesc);
                        }
                        tc_commit(tfull_bar + 8 * slot);                   // accumulator ready for the epilogue
                        }
                        if (++slot == ACC_SLOTS) { slot = 0; slot_phase ^= 1; }
                    }
                    if (elect_one()) tc_commit(empty_bar + 8 * stage);     // B stage reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) tc_commit(a_empty_bar);                    // query operand may be overwritten
            }
        }
    } else if (warp >= kLoadWarp0) {
        // ================= query loaders: packed bits -> +-1 bytes -> TMEM (thread = query row = TMEM lane) ====
        const int quad = warp & 3;
        uint32_t a_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = units[u];
            const int nsub = (wu.n_rows + BM - 1) / BM;
            uint4 p[MSUB][2];
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
                int row = s * BM + quad * 32 + lane;
                row = row < wu.n_rows ? row : wu.n_rows - 1;     // rows past the unit re-read its last row; never stored
                const uint4* src = reinterpret_cast<const uint4*>(wu.a_packed) + (size_t)row * 2;
                p[s][0] = __ldg(src);
                p[s][1] = __ldg(src + 1);
            }
            mbar_wait(a_empty_bar, a_phase ^ 1);                  // MMAs of the previous unit no longer read A
            a_phase ^= 1;
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
                if (s < nsub) {
                    const uint32_t w[8] = {p[s][0].x, p[s][0].y, p[s][0].z, p[s][0].w, p[s][1].x, p[s][1].y, p[s][1].z, p[s][1].w};
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + s * (KBYTES / 4);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {                 // 16 columns = 64 K-bytes = 2 packed words
                        uint32_t o[16];
#pragma unroll
                        for (int k = 0; k < 16; ++k) o[k] = pm1_from_nibble((w[2 * c + (k >> 3)] >> (4 * (k & 7))) & 0xFu);
                        tc_st16(taddr + c * 16, o);
                    }
                }
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full_bar);
        }
    } else if (warp >= kEpiWarp0) {
        // ================= epilogue: row-wise (max, first argmax) straight from TMEM =================
        const int e = warp - kEpiWarp0;
        const int quad = warp & 3;              // TMEM lane quadrant this warp may access
        const int half = e >> 2;                // which 64 of the 128 accumulator columns
        uint32_t slot = 0, slot_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = units[u];
            const int nsub = (wu.n_rows + BM - 1) / BM;
            int best_val[MSUB], best_idx[MSUB];
#pragma unroll
            for (int s = 0; s < MSUB; ++s) { best_val[s] = -0x7fffffff; best_idx[s] = 0; }
            for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                const int valid = wu.t_end - t;       // columns >= valid are outside the train image
#pragma unroll
                for (int s = 0; s < MSUB; ++s) {
                    if (s < nsub) {
                        mbar_wait(tfull_bar + 8 * slot, slot_phase);
                        tc_fence_after();
                        const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + slot * BN + half * 64;
                        int v0[32], v1[32];
                        if (dbg == 2 || dbg == 4) {       // DEBUG (timing experiments only): no TMEM read
#pragma unroll
                            for (int j = 0; j < 32; ++j) { v0[j] = j; v1[j] = j; }
                        } else {
                            tc_ld32(taddr, v0);
                            tc_ld32(taddr + 32, v1);
                            tc_wait_ld();
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar + 8 * slot);   // values are in registers now
                        if (++slot == ACC_SLOTS) { slot = 0; slot_phase ^= 1; }
                        const int c0 = half * 64;
                        if (dbg == 1 || dbg == 4) { best_val[s] = max(best_val[s], v0[0] + v1[31]); continue; }  // DEBUG: no ALU
                        if (c0 + 64 > valid) {                               // tail tile: mask outside columns
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (c0 + j >= valid) v0[j] = -0x7fffffff;
                                if (c0 + 32 + j >= valid) v1[j] = -0x7fffffff;
                            }
                        }
                        int m0 = v0[0], m1 = v1[0];
#pragma unroll
                        for (int j = 1; j < 32; ++j) { m0 = max(m0, v0[j]); m1 = max(m1, v1[j]); }
                        if (m0 > best_val[s]) {          // rare after the first tiles: strict '>' keeps the earliest
                            int idx = 31;
#pragma unroll
                            for (int j = 30; j >= 0; --j) idx = (v0[j] == m0) ? j : idx;
                            best_val[s] = m0; best_idx[s] = t + c0 + idx;
                        }
                        if (m1 > best_val[s]) {
                            int idx = 31;
#pragma unroll
                            for (int j = 30; j >= 0; --j) idx = (v1[j] == m1) ? j : idx;
                            best_val[s] = m1; best_idx[s] = t + c0 + 32 + idx;
                        }
                    }
                }
            }
            // merge: two column halves (and other train-range splits) meet in global memory
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
                const int row = s * BM + quad * 32 + lane;
                if (s < nsub && row < wu.n_rows && best_val[s] > -0x7fffffff) {
                    const uint32_t dist = (uint32_t)(KBYTES - best_val[s]) >> 1;   // <a,b> = 256 - 2*hamming
                    atomicMin(wu.key + row, (dist << kTrainIdxBits) | (uint32_t)best_idx[s]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// ---- host side ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

thread_local char g_tc_err[256] = "";   // one context per host thread: messages never interleave
EncodeTiledFn g_encode = nullptr;

bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
        return false;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return true;
}

bool ensure_dev(void*& p, size_t& cap, size_t bytes) {
    if (bytes <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 4096;
    if (cudaMalloc(&p, want) != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "cudaMalloc(%zu) failed", want); return false; }
    cap = want;
    return true;
}

}  // namespace

bool tc_available() { return true; }
const char* tc_last_error() { return g_tc_err; }
void tc_invalidate(TcState& s) { s.set_valid = false; }
void tc_reset_arena(TcState& s) { s.work_used = 0; }
void tc_release(TcState& s) {
    if (s.d_ops) cudaFree(s.d_ops);
    if (s.d_work) cudaFree(s.d_work);
    if (s.d_cnt) cudaFree(s.d_cnt);
    if (s.h_work) cudaFreeHost(s.h_work);
    s = TcState();
}

// Unpacks descriptor rows [0, n_rows) at `desc` (device, 32 B each) into s.d_ops (256 B each).
int tc_unpack(TcState& s, const uint8_t* desc, long long n_rows, cudaStream_t st) {
    if (!ensure_dev(s.d_ops, s.ops_cap, (size_t)(n_rows + BM) * KBYTES)) return -1;
    if (n_rows == 0) return 0;
    const long long n_words = n_rows * kDescWords;
    long long blocks = (n_words + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    unpack_pm1_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(desc), n_words,
                                                        static_cast<uint4*>(s.d_ops));
    kmark("unpack_pm1", st);
    s.ops_rows = n_rows;
    return 1;
}

// Precondition (as for the popc kernel): keys initialised to kKeyInit.  The operands of all pairs must lie
// in ONE contiguous descriptor array starting at `desc_base` (the image set, or the ad-hoc q|t buffers):
// row index of a pair's descriptors = (ptr - desc_base) / 32.
int launch_hamming_tc(TcState& s, const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                      cudaStream_t st) {
    (void)d_pairs;
    if (n_pairs <= 0) return 0;
    if (!load_encode()) return -1;
    int launches = 0;
    // operand array = span of descriptor rows referenced by this batch
    const uint8_t* lo = nullptr;
    const uint8_t* hi = nullptr;
    for (int p = 0; p < n_pairs; ++p) {
        const PairDesc& pd = h_pairs[p];
        if (pd.n1 <= 0 || pd.n2 <= 0) continue;
        const uint8_t* ends[2][2] = {{pd.desc1, pd.desc1 + (size_t)pd.n1 * 32}, {pd.desc2, pd.desc2 + (size_t)pd.n2 * 32}};
        for (auto& e : ends) {
            if (!lo || e[0] < lo) lo = e[0];
            if (!hi || e[1] > hi) hi = e[1];
        }
    }
    if (!lo) return 0;
    const bool pinned = s.cache_enabled && s.span_lo && lo >= s.span_lo && hi <= s.span_lo + (size_t)s.span_rows * 32;
    tc_apply_span(s, lo, hi);
    const long long n_rows = (hi - lo) / 32;
    long long ref_rows = 0;
    for (int p = 0; p < n_pairs; ++p) ref_rows += (long long)h_pairs[p].n1 + h_pairs[p].n2;
    if (!pinned && n_rows > 64 * ref_rows + 4096) {   // the span must be one descriptor array, not two unrelated allocations
        snprintf(g_tc_err, sizeof g_tc_err, "descriptor operands are not in one contiguous array");
        return -1;
    }
    if (n_rows >= (1ll << 31)) { snprintf(g_tc_err, sizeof g_tc_err, "operand span too large"); return -1; }
    if (!(s.set_valid && s.ops_src == lo && s.ops_rows == n_rows && s.ops_row_bytes == KBYTES)) {
        const int l = tc_unpack(s, lo, n_rows, st);
        if (l < 0) return -1;
        launches += l;
        s.ops_src = lo;
        s.ops_row_bytes = KBYTES;
        s.set_valid = s.cache_enabled;
    }
    // tensor map over the unpacked array: dim0 = K bytes (256), dim1 = rows; box 128 B x 128 rows, 128B swizzle
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)KBYTES, (cuuint64_t)(n_rows + BM)};
        const cuuint64_t gstride[1] = {(cuuint64_t)KBYTES};
        const cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)BM};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, s.d_ops, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_tc_err, sizeof g_tc_err, "cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    }
    // work units: (pair, 256-query block, train-tile range).  Small batches split the train range so every SM works.
    long long qblocks = 0;
    int max_tiles = 1;
    for (int p = 0; p < n_pairs; ++p) {
        if (h_pairs[p].n1 <= 0 || h_pairs[p].n2 <= 0) continue;
        qblocks += (h_pairs[p].n1 + BM * MSUB - 1) / (BM * MSUB);
        const int tiles = (h_pairs[p].n2 + BN - 1) / BN;
        if (tiles > max_tiles) max_tiles = tiles;
    }
    int tsplit = 1;
    if (qblocks < 2LL * sm_count) {
        long long t = (2LL * sm_count + qblocks - 1) / qblocks;
        if (t > max_tiles) t = max_tiles;
        tsplit = (int)(t < 1 ? 1 : t);
    }
    const size_t max_units = (size_t)qblocks * tsplit;
    const size_t wbytes = (max_units * sizeof(WorkUnit) + 255) & ~(size_t)255;
    if (s.work_used + wbytes > s.h_work_cap || s.work_used + wbytes > s.work_cap) {
        // arena exhausted: drain the stream (earlier tables are then dead), then grow both sides
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "stream sync failed"); return -1; }
        s.work_used = 0;
        if (16 * wbytes > s.h_work_cap) {   // room for many launches between synchronisations (pipelined sub-batches)
            if (s.h_work) cudaFreeHost(s.h_work);
            s.h_work = nullptr; s.h_work_cap = 0;
            const size_t want = 16 * wbytes + (8u << 20);
            if (cudaMallocHost(&s.h_work, want) != cudaSuccess) { snprintf(g_tc_err, sizeof g_tc_err, "cudaMallocHost failed"); return -1; }
            s.h_work_cap = want;
        }
        if (!ensure_dev(s.d_work, s.work_cap, s.h_work_cap)) return -1;
    }
    WorkUnit* wu = reinterpret_cast<WorkUnit*>(static_cast<char*>(s.h_work) + s.work_used);
    WorkUnit* d_wu = reinterpret_cast<WorkUnit*>(static_cast<char*>(s.d_work) + s.work_used);
    size_t n_units = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const PairDesc& pd = h_pairs[p];
        if (pd.n1 <= 0 || pd.n2 <= 0) continue;
        const uint32_t brow = (uint32_t)((pd.desc2 - lo) / 32);
        const int tiles = (pd.n2 + BN - 1) / BN;
        const int tper = (tiles + tsplit - 1) / tsplit;
        for (int q0 = 0; q0 < pd.n1; q0 += BM * MSUB) {
            for (int ts = 0; ts * tper < tiles; ++ts) {
                WorkUnit& w = wu[n_units++];
                w.a_packed = pd.desc1 + (size_t)q0 * 32;
                w.b_row0 = brow;
                w.n_rows = (pd.n1 - q0 < BM * MSUB) ? pd.n1 - q0 : BM * MSUB;
                w.t_begin = ts * tper * BN;
                const int te = (ts + 1) * tper * BN;
                w.t_end = te < pd.n2 ? te : pd.n2;
                w.key = pd.key + q0;
            }
        }
    }
    if (n_units == 0) return launches;
    {
        const int n16 = (int)(n_units * sizeof(WorkUnit) / 16);
        int blocks = (n16 + 255) / 256;
        if (blocks > 64) blocks = 64;
        pull_table_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(wu), reinterpret_cast<uint4*>(d_wu), n16);
        ++launches;
    }
    static const int dbg = getenv("SFMGMS_TC_DEBUG") ? atoi(getenv("SFMGMS_TC_DEBUG")) : 0;   // timing experiments only
    auto kern = dbg == 1 ? hamming_tc_kernel<1> : dbg == 2 ? hamming_tc_kernel<2> : dbg == 3 ? hamming_tc_kernel<3>
              : dbg == 4 ? hamming_tc_kernel<4> : hamming_tc_kernel<0>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
        snprintf(g_tc_err, sizeof g_tc_err, "cudaFuncSetAttribute(smem=%d) failed", SMEM_BYTES);
        return -1;
    }
    const int grid = (int)(n_units < (size_t)sm_count ? n_units : (size_t)sm_count);
    kern<<<grid, kThreads, SMEM_BYTES, st>>>(tmap, d_wu, (int)n_units);
    kmark("hamming_tc", st);
    s.work_used += wbytes;
    return launches + 1;
}

}  // namespace sfmgms
