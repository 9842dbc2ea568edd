// l2_tc.cu — brute-force L2 nearest neighbour for integer-valued 128-d descriptors (OpenCV SIFT) on tcgen05.
//
// the reference's literal main path is SIFT + cv::BFMatcher(NORM_L2)::match
// (FeatureMatchUtil.cpp:10, 66-68).  With descriptors that are integers in [0,255] (what OpenCV's SIFT emits)
//     d2(a,b) = |a|^2 + |b|^2 - 2<a,b>
// is exact integer arithmetic and OpenCV's float result is sqrtf((float)d2) bit for bit (see l2_dp4a.cu).
// <a,b> is a u8 x u8 GEMM: tcgen05.mma.kind::i8 (unsigned operands, s32 accumulators in TMEM), one descriptor =
// 128 bytes = one 128B-swizzle row = the whole K (4 instructions of K=32 per 128x240 accumulator).
// The epilogue never writes the distance matrix: per accumulator element one IMAD builds
//     key_j = (|b_j|^2 - 2<a,b_j>) * 128 + (j mod 80)
// from a per-column constant (|b_j|^2*128 + j mod 80, staged in shared memory by a helper warp), a VIMNMX3 tree
// takes the minimum: smallest squared distance, lowest column on ties; |a|^2 is added once per row at the end.
// OpenCV compares FLOAT distances: above d2 = 2^22 neighbouring integers can share a float, so rows whose minimum
// lands there (never for real SIFT, whose d2 <= ~1.05e6) are re-scanned exactly by l2_fixup_kernel.
//
// Same persistent warp-specialised structure as hamming_fp4.cu: warps 0-11 epilogue (TMEM lane quadrant x
// 80-column part), warp 12 TMA producer, warp 13 MMA issuer, warp 14 TMEM allocator, warp 15 column-constant stager.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace sfmgms {

namespace {

using namespace tcptx;

constexpr int kDim = 128;
constexpr int BM = 128;
constexpr int MSUB = 3;                   // query sub-tiles per work unit (384 rows)
constexpr int BN = 240;
constexpr int ROWB = kDim;                // u8
constexpr int A_BUFS = 2;
constexpr int STAGES = 3;
constexpr int NRING = 8;                  // column-constant ring (tiles)
constexpr int ACC_SLOTS = 2;
constexpr int TILE_BYTES = BM * ROWB;     // 16 KB
constexpr int BTILE_BYTES = BN * ROWB;    // 30 KB
constexpr int A_BUF_BYTES = MSUB * TILE_BYTES;
constexpr int NORM_BYTES = NRING * BN * 4;
constexpr int SMEM_DATA = A_BUFS * A_BUF_BYTES + STAGES * BTILE_BYTES + NORM_BYTES;
constexpr int SMEM_BYTES = SMEM_DATA + 1024 + 512;
constexpr int kEpiWarps = 12;
constexpr int kEpiCols = BN / 3;          // 80
constexpr int kProdWarp = 12, kMmaWarp = 13, kAllocWarp = 14, kNormWarp = 15;
constexpr int kThreads = 512;
constexpr int kIntMax = 0x7fffffff;

constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
// InstrDescriptor: c_format=S32(2)@4, a_format=b_format=UINT8(0), K-major both, n_dim=N>>3 @17, m_dim=M>>4 @24
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

template <int kAccumulate>
__device__ __forceinline__ void tc_mma_u8(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo) {
    asm volatile(
        "{\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(kIdesc), "n"(kAccumulate) : "memory");
}

struct L2Params {
    const uint32_t* qn;      // |a|^2 per query row
    const uint32_t* tn;      // |b|^2 per train row
    unsigned long long* key; // per query row: (d2 << 18) | trainIdx, pre-set to ~0
    int nq, nt, tsplit, nts, tper, n_units;
    uint32_t a_row0, b_row0; // operand rows of query row 0 / train row 0
};

struct Unit { int q0, n_rows, t_begin, t_end; };
__device__ __forceinline__ Unit make_unit(const L2Params& P, int u) {
    const int qb = u / P.nts, ts = u - qb * P.nts;
    Unit w;
    w.q0 = qb * (BM * MSUB);
    w.n_rows = min(P.nq - w.q0, BM * MSUB);
    w.t_begin = ts * P.tper * BN;
    w.t_end = min((ts + 1) * P.tper * BN, P.nt);
    return w;
}

__global__ void __launch_bounds__(kThreads, 1)
l2_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
             const __grid_constant__ L2Params P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = smem_base;
    const uint32_t b_smem = a_smem + A_BUFS * A_BUF_BYTES;
    const uint32_t n_smem = b_smem + STAGES * BTILE_BYTES;
    const uint32_t bar_base = smem_base + SMEM_DATA;
    const uint32_t full_bar = bar_base;
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t a_full_bar = empty_bar + 8 * STAGES;
    const uint32_t a_empty_bar = a_full_bar + 8 * A_BUFS;
    const uint32_t tfull_bar = a_empty_bar + 8 * A_BUFS;
    const uint32_t tempty_bar = tfull_bar + 8 * ACC_SLOTS;
    const uint32_t nfull_bar = tempty_bar + 8 * ACC_SLOTS;
    const uint32_t nempty_bar = nfull_bar + 8 * NRING;
    const uint32_t tmem_ptr_smem = nempty_bar + 8 * NRING;
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));
    int* norm_generic = reinterpret_cast<int*>(smem_raw + (n_smem - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < A_BUFS; ++s) { mbar_init(a_full_bar + 8 * s, 1); mbar_init(a_empty_bar + 8 * s, 1); }
        for (int s = 0; s < ACC_SLOTS; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, kEpiWarps); }
        for (int s = 0; s < NRING; ++s) { mbar_init(nfull_bar + 8 * s, 1); mbar_init(nempty_bar + 8 * s, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == kAllocWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;
    const int n_units = P.n_units;

    // single-issuer roles: the whole warp runs the loop converged and elects one lane per issue (see hamming_fp4.cu)
    if (warp == kProdWarp) {
        {
            if (elect_one()) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
            }
            uint32_t stage = 0, phase = 0, abuf = 0, a_phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit wu = make_unit(P, u);
                const int nsub = (wu.n_rows + BM - 1) / BM;
                mbar_wait(a_empty_bar + 8 * abuf, a_phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(a_full_bar + 8 * abuf, (uint32_t)(nsub * TILE_BYTES));
                    for (int s = 0; s < nsub; ++s)
                        tma_load_2d(a_smem + abuf * A_BUF_BYTES + s * TILE_BYTES, &tmapA, 0, (int)(P.a_row0 + wu.q0 + s * BM),
                                    a_full_bar + 8 * abuf);
                }
                if (++abuf == A_BUFS) { abuf = 0; a_phase ^= 1; }
                for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                    mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)BTILE_BYTES);
                        tma_load_2d(b_smem + stage * BTILE_BYTES, &tmapB, 0, (int)(P.b_row0 + t), full_bar + 8 * stage);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        {
            uint32_t stage = 0, phase = 0, abuf = 0, a_phase = 0, slot = 0, slot_phase = 0;
            const uint32_t a_lo0 = sdesc_lo(a_smem), b_lo0 = sdesc_lo(b_smem);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const Unit wu = make_unit(P, u);
                const int nsub = (wu.n_rows + BM - 1) / BM;
                const int ntiles = (wu.t_end - wu.t_begin + BN - 1) / BN;
                mbar_wait(a_full_bar + 8 * abuf, a_phase);
                const uint32_t a_lo_buf = a_lo0 + abuf * (A_BUF_BYTES >> 4);
                for (int ti = 0; ti < ntiles; ++ti) {
                    mbar_wait(full_bar + 8 * stage, phase);
                    const uint32_t b_lo = b_lo0 + stage * (BTILE_BYTES >> 4);
                    for (int s = 0; s < nsub; ++s) {
                        mbar_wait(tempty_bar + 8 * slot, slot_phase ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + slot * BN;
                        const uint32_t a_lo = a_lo_buf + s * (TILE_BYTES >> 4);
                        if (elect_one()) {
                            tc_mma_u8<0>(d, a_lo + 0, b_lo + 0);      // K = 128 = 4 steps of 32 bytes inside the swizzle row
                            tc_mma_u8<1>(d, a_lo + 2, b_lo + 2);
                            tc_mma_u8<1>(d, a_lo + 4, b_lo + 4);
                            tc_mma_u8<1>(d, a_lo + 6, b_lo + 6);
                            tc_commit(tfull_bar + 8 * slot);
                        }
                        if (++slot == ACC_SLOTS) { slot = 0; slot_phase ^= 1; }
                    }
                    if (elect_one()) tc_commit(empty_bar + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) tc_commit(a_empty_bar + 8 * abuf);
                if (++abuf == A_BUFS) { abuf = 0; a_phase ^= 1; }
            }
        }
    } else if (warp == kNormWarp) {
        // ===== column constants of each train tile: |b_j|^2 * 128 + (j mod 80), ring of NRING tiles =====
        uint32_t ns = 0, n_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Unit wu = make_unit(P, u);
            for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                mbar_wait(nempty_bar + 8 * ns, n_phase ^ 1);
                for (int j = lane; j < BN; j += 32) {
                    const int row = t + j;
                    const uint32_t nb = row < P.nt ? __ldg(P.tn + row) : 0u;
                    norm_generic[ns * BN + j] = (int)(nb * 128u + (uint32_t)(j % kEpiCols));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(nfull_bar + 8 * ns);      // release: the stores above are visible to waiters
                if (++ns == NRING) { ns = 0; n_phase ^= 1; }
            }
        }
    } else if (warp < kEpiWarps) {
        const int quad = warp & 3;
        const int c0 = (warp >> 2) * kEpiCols;
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
        uint32_t slot = 0, slot_phase = 0, ns = 0, n_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const Unit wu = make_unit(P, u);
            const int nsub = (wu.n_rows + BM - 1) / BM;
            int best[MSUB], beat[MSUB], base[MSUB];
#pragma unroll
            for (int s = 0; s < MSUB; ++s) { best[s] = kIntMax; beat[s] = kIntMax & ~127; base[s] = -1; }
            for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                const int valid = wu.t_end - t;
                mbar_wait(nfull_bar + 8 * ns, n_phase);
                const int4* cn = reinterpret_cast<const int4*>(norm_generic + ns * BN + c0);
#pragma unroll
                for (int s = 0; s < MSUB; ++s) {
                    if (s < nsub) {
                        mbar_wait(tfull_bar + 8 * slot, slot_phase);
                        tc_fence_after();
                        int r[kEpiCols];
                        tc_ld64(tbase + slot * BN, reinterpret_cast<int(&)[64]>(r[0]));
                        tc_ld16(tbase + slot * BN + 64, reinterpret_cast<int(&)[16]>(r[64]));
                        tc_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar + 8 * slot);
                        if (++slot == ACC_SLOTS) { slot = 0; slot_phase ^= 1; }
                        // key_j = (|b_j|^2 - 2<a,b_j>) * 128 + (j mod 80)
#pragma unroll
                        for (int k = 0; k < kEpiCols / 4; ++k) {
                            const int4 c = cn[k];                      // same address in every lane: broadcast
                            r[4 * k + 0] = c.x - 256 * r[4 * k + 0];
                            r[4 * k + 1] = c.y - 256 * r[4 * k + 1];
                            r[4 * k + 2] = c.z - 256 * r[4 * k + 2];
                            r[4 * k + 3] = c.w - 256 * r[4 * k + 3];
                        }
                        if (c0 + kEpiCols > valid) {
#pragma unroll
                            for (int j = 0; j < kEpiCols; ++j)
                                if (c0 + j >= valid) r[j] = kIntMax;
                        }
                        int m = r[0];
#pragma unroll
                        for (int j = 1; j < kEpiCols; ++j) m = min(m, r[j]);
                        if (m < beat[s]) {                             // strictly smaller squared distance
                            best[s] = m;
                            beat[s] = m & ~127;
                            base[s] = t + c0;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(nempty_bar + 8 * ns);       // this warp is done with the tile's constants
                if (++ns == NRING) { ns = 0; n_phase ^= 1; }
            }
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
                const int row = s * BM + quad * 32 + lane;
                if (s < nsub && row < wu.n_rows && base[s] >= 0) {
                    const int grow = wu.q0 + row;
                    const uint32_t d2 = (uint32_t)((best[s] >> 7) + (int)__ldg(P.qn + grow));
                    const uint32_t idx = (uint32_t)(base[s] + (best[s] & 127));
                    atomicMin(P.key + grow, ((unsigned long long)d2 << kTrainIdxBits) | idx);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// float [n][128] -> u8 [n][128] + squared norm; flags any value that is not an integer in [0,255]
__global__ void __launch_bounds__(128) l2_narrow_kernel(const float* __restrict__ src, int n, uint32_t* __restrict__ dst,
                                                        uint32_t* __restrict__ norm2, int* __restrict__ bad) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float4 v = reinterpret_cast<const float4*>(src + (size_t)row * kDim)[lane];
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t w = 0, s = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float r = rintf(f[k]);
        ok &= (r == f[k]) && (r >= 0.f) && (r <= 255.f);
        const uint32_t u = ok ? (uint32_t)r : 0u;
        w |= u << (8 * k);
        s += u * u;
    }
    if (!ok) atomicExch(bad, 1);
    dst[(size_t)row * (kDim / 4) + lane] = w;
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) norm2[row] = s;
}

// Rows whose minimum squared distance is >= 2^22: OpenCV orders by the FLOAT sqrt, where neighbouring integers can
// coincide and the lower index wins; re-scan such a row exactly (one warp per row, DP4A).  Never taken for SIFT.
__global__ void __launch_bounds__(256) l2_fixup_kernel(const uint32_t* __restrict__ q8, const uint32_t* __restrict__ qn, int nq,
                                                       const uint32_t* __restrict__ t8, const uint32_t* __restrict__ tn,
                                                       int nt, unsigned long long* __restrict__ key) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= nq) return;
    if ((uint32_t)(key[row] >> kTrainIdxBits) < (1u << 22)) return;       // warp-uniform
    uint32_t a[kDim / 4];
#pragma unroll
    for (int w = 0; w < kDim / 4; ++w) a[w] = __ldg(q8 + (size_t)row * (kDim / 4) + w);
    const uint32_t an = qn[row];
    unsigned long long best = ~0ull;
    for (int j = lane; j < nt; j += 32) {
        int dot = 0;
#pragma unroll
        for (int w = 0; w < kDim / 4; ++w) dot = __dp4a(a[w], __ldg(t8 + (size_t)j * (kDim / 4) + w), (unsigned)dot);
        uint32_t d2 = an + tn[j] - 2u * (uint32_t)dot;
        if (d2 >= (1u << 22)) d2 = (1u << 22) + (__float_as_uint(__fsqrt_rn((float)d2)) - 0x45000000u);
        const unsigned long long kk = ((unsigned long long)d2 << kTrainIdxBits) | (unsigned)j;
        best = kk < best ? kk : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) key[row] = best;
}

__global__ void l2_decode_kernel(const unsigned long long* __restrict__ key, int n, int32_t* __restrict__ train_idx,
                                 float* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    train_idx[i] = (int32_t)(k & kTrainIdxMask);
    const uint32_t g = (uint32_t)(k >> kTrainIdxBits);
    dist[i] = g < (1u << 22) ? __fsqrt_rn((float)g) : __uint_as_float(g - (1u << 22) + 0x45000000u);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return false;
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return true;
}

}  // namespace

size_t l2_tc_scratch_bytes(int nq, int nt) {
    return (size_t)nq * 8 + (size_t)(nq + nt) * 4 + 64 + (size_t)(nq + nt + 256) * kDim + 1024;
}

// d_q/d_t: device float descriptors [n][128].  Returns the number of kernel launches, -1 on a setup error.
int launch_l2_tc(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, int32_t* d_train_idx,
                 float* d_dist, int* d_bad, int sm_count, cudaStream_t st) {
    if (!load_encode()) return -1;
    uint8_t* p = static_cast<uint8_t*>(d_scratch);
    unsigned long long* key = reinterpret_cast<unsigned long long*>(p); p += (size_t)nq * 8;
    uint32_t* qn = reinterpret_cast<uint32_t*>(p); p += (size_t)nq * 4;
    uint32_t* tn = reinterpret_cast<uint32_t*>(p); p += (size_t)nt * 4;
    p = reinterpret_cast<uint8_t*>(((uintptr_t)p + 1023) & ~(uintptr_t)1023);
    uint32_t* ops = reinterpret_cast<uint32_t*>(p);              // [(nq + nt + 256)][128] u8: queries, then train rows
    uint32_t* q8 = ops;
    uint32_t* t8 = ops + (size_t)nq * (kDim / 4);
    cudaMemsetAsync(key, 0xFF, (size_t)nq * 8, st);
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    l2_narrow_kernel<<<(nq + 3) / 4, 128, 0, st>>>(d_q, nq, q8, qn, d_bad);
    l2_narrow_kernel<<<(nt + 3) / 4, 128, 0, st>>>(d_t, nt, t8, tn, d_bad);

    CUtensorMap tmA, tmB;
    const cuuint64_t gdim[2] = {(cuuint64_t)ROWB, (cuuint64_t)(nq + nt + 256)};
    const cuuint64_t gstride[1] = {(cuuint64_t)ROWB};
    const cuuint32_t estr[2] = {1, 1};
    const cuuint32_t boxA[2] = {(cuuint32_t)ROWB, (cuuint32_t)BM}, boxB[2] = {(cuuint32_t)ROWB, (cuuint32_t)BN};
    if (g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ops, gdim, gstride, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        g_encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ops, gdim, gstride, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1;
    L2Params P;
    P.qn = qn; P.tn = tn; P.key = key; P.nq = nq; P.nt = nt;
    P.a_row0 = 0; P.b_row0 = (uint32_t)nq;
    const int qblocks = (nq + BM * MSUB - 1) / (BM * MSUB);
    const int tiles = (nt + BN - 1) / BN;
    int tsplit = 1;
    {
        double best = 1e30;
        for (int ts = 1; ts <= 64 && ts <= tiles; ++ts) {
            const long long units = (long long)qblocks * ts;
            const double makespan = (double)((units + sm_count - 1) / sm_count) / ts * (1.0 + 0.01 * (ts - 1));
            if (makespan < best - 1e-9) { best = makespan; tsplit = ts; }
        }
    }
    P.tsplit = tsplit;
    P.tper = (tiles + tsplit - 1) / tsplit;
    P.nts = (tiles + P.tper - 1) / P.tper;
    P.n_units = qblocks * P.nts;
    if (cudaFuncSetAttribute(l2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) return -1;
    const int grid = P.n_units < sm_count ? P.n_units : sm_count;
    l2_tc_kernel<<<grid, kThreads, SMEM_BYTES, st>>>(tmA, tmB, P);
    l2_fixup_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(q8, qn, nq, t8, tn, nt, key);
    l2_decode_kernel<<<(nq + 255) / 256, 256, 0, st>>>(key, nq, d_train_idx, d_dist);
    return 5;
}

}  // namespace sfmgms
