// hamming_fp4.cu — brute-force Hamming nearest neighbour as a block-scaled FP4 GEMM (tcgen05 kind::mxf4).
//
// Same idea as hamming_tc.cu (replaces cv::BFMatcher(NORM_HAMMING)::match, FeatureMatchUtil.cpp:68): a descriptor
// bit becomes +1 / -1, <a,b> = 256 - 2*hamming(a,b), the nearest train row is the row arg-max.  Here the +-1
// values are E2M1 (4-bit) numbers, two per byte, with all UE8M0 block scales = 2^0: the products and the fp32
// accumulation are exact (|sum| <= 256), the operand rows shrink to 128 bytes (one 128B-swizzle row = the whole
// K), and the tensor pipe runs the mxf4 rate (K = 64 per instruction, 4 instructions per 128x128 accumulator).
// Block scales: every scale byte this kernel could ever address is 0x7F, so the scale-factor layout in TMEM does
// not matter: 32 TMEM columns are filled with 0x7F7F7F7F once per CTA.
//
// Kernel structure (persistent, warp-specialised, 1 CTA/SM, 512 threads): warps 0-11 epilogue (warp = TMEM lane
// quadrant x 80-column part: tcgen05.ld -> index-packed keys -> max tree -> atomicMin(dist<<18|trainIdx)),
// warp 12 TMA producer (query sub-tiles double-buffered per work unit, 240-row train tiles through a ring),
// warp 13 MMA issuer (4 x M128 N240 K64 per accumulator, 2 accumulator slots in TMEM), warp 14 TMEM allocator;
// warps 14-15 then settle the tie rule of every finished unit on the packed descriptors (fused tie resolution, below)
// while the other warps are already on the next unit.
// The single-lane roles sit in the highest warp ids because the warp scheduler favours those.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <type_traits>
#include <utility>
#include <vector>

#include "hamming_tc.cuh"
#include "tc_ptx.cuh"

namespace sfmgms {

namespace {

using namespace tcptx;

constexpr int BM = 128;            // UMMA M
// SFMGMS_FP4_V3: three TMEM accumulator slots of 160 columns, three query sub-tiles per unit, and every epilogue warp OWNS one
// (lane quadrant, sub-tile = slot): it reads all 160 columns of every third accumulator instead of 80 columns of every
// accumulator.  The warps of a scheduler are then in different phases (one waits for tensor memory while another runs its max
// tree) instead of in lockstep on the same accumulator, and a slot has two accumulator times to drain instead of one.
// MEASURED (bit-exact, 177 parity tests + stress): 1.795 vs 1.760 ms per 256 cfg2 pairs -- 2 % SLOWER than the two-slot layout,
// so it is not the default.  The lockstep was not the limit: a warp alone in its max tree exposes the FMNMX3 dependency
// latency that three interleaved warps hide (profiles/r2_notes.md).
#ifndef SFMGMS_FP4_V3
#define SFMGMS_FP4_V3 0
#endif
constexpr bool kV3 = SFMGMS_FP4_V3 != 0;
#ifndef SFMGMS_FP4_MSUB
#define SFMGMS_FP4_MSUB (SFMGMS_FP4_V3 ? 3 : 4)
#endif
constexpr int MSUB = SFMGMS_FP4_MSUB;  // query sub-tiles per work unit
constexpr int BN = kV3 ? 160 : 240;    // UMMA N: accumulator slots x BN columns + 32 scale columns = 512 TMEM columns
constexpr int ROWB = 128;          // bytes per unpacked row: 256 x e2m1
constexpr int UMMA_KB = 32;        // bytes of K per mxf4 MMA (64 elements)
constexpr int NKSTEP = ROWB / UMMA_KB;   // 4
constexpr int A_BUFS = 2;
constexpr int TILE_BYTES = BM * ROWB;    // 16 KB: one query sub-tile
constexpr int BTILE_BYTES = BN * ROWB;   // 30 KB: one train tile (multiple of 1024: stays swizzle-aligned)
constexpr int A_BUF_BYTES = MSUB * TILE_BYTES;
constexpr int STAGES = (222 * 1024 - A_BUFS * A_BUF_BYTES) / BTILE_BYTES;   // 3 (MSUB=4) / 5 (MSUB=2)
constexpr int SMEM_DATA = A_BUFS * A_BUF_BYTES + STAGES * BTILE_BYTES;

constexpr int ACC_SLOTS = kV3 ? 3 : 2;
static_assert(!kV3 || MSUB == ACC_SLOTS, "V3: sub-tile s lives in slot s");
constexpr int SF_COL = ACC_SLOTS * BN;   // TMEM columns [480, 512): block scales (all 1.0)
constexpr int SF_COLS = 32;
constexpr int kEpiWarp0 = 0;            // epilogue warps first: the scheduler favours HIGH warp ids, which must be the
#ifndef SFMGMS_FP4_EPIW
#define SFMGMS_FP4_EPIW 12
#endif
constexpr int kEpiWarps = SFMGMS_FP4_EPIW;        // 12: 3 per TMEM lane quadrant x 80 columns; 16: 4 x 60 columns
constexpr int kEpiCols = kV3 ? 80 : BN / (kEpiWarps / 4);    // columns per tcgen05.ld batch: 80 = x64 + x16; 60 = x32 + x16 + x8 + x4
#ifndef SFMGMS_FP4_PRODW
#define SFMGMS_FP4_PRODW 0
#define SFMGMS_FP4_MMAW 1
#define SFMGMS_FP4_ALLOCW 2
#endif
constexpr int kProdWarp = kEpiWarps + SFMGMS_FP4_PRODW;     // latency-critical single-lane roles (TMA producer, MMA issuer)
constexpr int kMmaWarp = kEpiWarps + SFMGMS_FP4_MMAW;
constexpr int kAllocWarp = kEpiWarps + SFMGMS_FP4_ALLOCW;
constexpr int kThreads = 32 * (kEpiWarps + 4);   // 512
constexpr int kResWarp0 = kAllocWarp;            // fused tie resolution: the allocator warp and the spare one behind it
constexpr int kResThreads = 64;
constexpr int kKeyParts = kV3 ? 1 : kEpiWarps / 4;   // epilogue warps that hold a candidate for the same query row (column parts)
static_assert(!kV3 || kEpiWarps == 4 * ACC_SLOTS, "V3: one epilogue warp per (lane quadrant, slot)");
constexpr int SMEM_KEYS = kKeyParts * MSUB * BM * 4;   // one candidate key per (part, row) of a unit
constexpr int SMEM_LUT = 256 * 4;                      // packed query tiles: descriptor byte -> 8 e2m1 values
constexpr int SMEM_BYTES = SMEM_DATA + 1024 + 256 + SMEM_KEYS + SMEM_LUT;
static_assert(SMEM_BYTES <= 227 * 1024 && kAllocWarp == kEpiWarps + 2, "shared memory / warp roles");
static_assert(STAGES >= 3, "B ring too shallow");
static_assert(BTILE_BYTES % 1024 == 0 && (kEpiCols == 80 || kEpiCols == 60) && SF_COL + SF_COLS <= 512, "layout");

// Epilogue reduction.  A warp's 80 accumulator columns are 9 GROUPS of 9 consecutive columns (the last one has 8).  Per
// group the plain maximum of its dot products, then key_g = groupmax_g + (31 - g)/32 (one packed FADD2 per two groups, FMA
// pipe) and the maximum of the 9 keys: its integer part is the best dot product of the part, its fraction names the FIRST
// group that attains it.  |dot| <= 256 and the fraction has 5 bits: keys are exact in fp32.  All fractions lie within a span
// below 1/2, so "this key's dot product is larger than the best so far" is simply key > best + 1/2 (no floor in the loop).
// Which of the group's train rows is the lowest-index minimum is settled afterwards on the original descriptors (9
// candidates per query row instead of n2: warps 14-15, or hamming_resolve_kernel for small launches).
// Why 9: a 3-input maximum (FMNMX3) folds 2 values per instruction, so groups of ODD size cost exactly (size - 1) / 2
// instructions -- 9 values = 4, and the 9 keys = 4 more: 36 + 5 (tags) + 4 = 45 ALU instructions per 80 columns.  Groups of 8
// (round 2 until here) need 4 per group as well (3 FMNMX3 + 1 FMNMX) but 10 of them: 40 + 5 + 5 = 50.  The epilogue is what
// the kernel waits for (profiles/r2_notes.md), so its instruction count is the kernel's time.  -DSFMGMS_FP4_GROUP=8 builds
// the old grouping.
#ifndef SFMGMS_FP4_GROUP
#define SFMGMS_FP4_GROUP 9
#endif
constexpr int kGroup = SFMGMS_FP4_GROUP;                           // columns per tagged group (the last group may be shorter)
constexpr int kGroups = (kEpiCols + kGroup - 1) / kGroup;          // 9 (10 with groups of 8)
constexpr int kLastGroup = kEpiCols - (kGroups - 1) * kGroup;      // 8
constexpr int kCandBits = 4;                                       // candidate number within a group: 4 bits
constexpr int kTagDen = 32;                                        // tag of group g = (kTagDen - 1 - g) / kTagDen: all tags within a span < 1/2
static_assert(kEpiCols == 80 && kGroup >= 5 && kGroup <= 16 && kGroups <= 16 && kLastGroup >= 1, "group tags: at most 16 groups, candidates in 4 bits");
__constant__ float2 c_grouptag[8] = {{31.f / 32.f, 30.f / 32.f}, {29.f / 32.f, 28.f / 32.f}, {27.f / 32.f, 26.f / 32.f}, {25.f / 32.f, 24.f / 32.f},
                                     {23.f / 32.f, 22.f / 32.f}, {21.f / 32.f, 20.f / 32.f}, {19.f / 32.f, 18.f / 32.f}, {17.f / 32.f, 16.f / 32.f}};

// maximum of N values as a tree of 3-input maxima (ptxas: FMNMX3)
template <int N>
__device__ __forceinline__ float max_tree(const float* x) {
    if constexpr (N == 1) return x[0];
    else if constexpr (N == 2) return fmaxf(x[0], x[1]);
    else {
        constexpr int n0 = (N + 2) / 3, n1 = (N + 1) / 3, n2 = N / 3;
        return fmaxf(fmaxf(max_tree<n0>(x), max_tree<n1>(x + n0)), max_tree<n2>(x + n0 + n1));
    }
}

// (lo, hi) + (c.x, c.y) with one packed fp32x2 add (sm_100: add.rn.f32x2 -> FADD2)
__device__ __forceinline__ void add2(float lo, float hi, float2 c, float& out_lo, float& out_hi) {
    unsigned long long x, cc, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(lo), "f"(hi));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c.x), "f"(c.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(out_lo), "=f"(out_hi) : "l"(r));
}

// one warp's part of an accumulator row-quadrant: kEpiCols consecutive 32-bit TMEM columns -> registers
__device__ __forceinline__ void ld_part(uint32_t taddr, int (&r)[kEpiCols]) {
    if constexpr (kEpiCols == 80) {
        tc_ld64(taddr, reinterpret_cast<int(&)[64]>(r[0]));
        tc_ld16(taddr + 64, reinterpret_cast<int(&)[16]>(r[64]));
    } else {
        tc_ld32(taddr, reinterpret_cast<int(&)[32]>(r[0]));
        tc_ld16(taddr + 32, reinterpret_cast<int(&)[16]>(r[32]));
        tc_ld8(taddr + 48, reinterpret_cast<int(&)[8]>(r[48]));
        tc_ld4(taddr + 56, reinterpret_cast<int(&)[4]>(r[56]));
    }
}

struct WorkUnit {
    uint32_t a_row0, b_row0;   // operand rows (in the unpacked array) of the unit's first query / the train image's row 0
    int32_t n_rows, t_begin, t_end;
    uint32_t* key;
    // fused tie resolution: the packed descriptors behind the operands, and the units that share this unit's query rows
    const uint8_t *q_desc, *t_desc;   // first query row of the unit / train row 0
    int32_t n2, nts, group;           // train rows; train-range splits of the pair; index of the group's first unit
};

// Work decomposition travels as a kernel PARAMETER (no table upload: the copy engine may be busy with a queued
// image-set transfer, and a PCIe read from the SMs would contend with it): prefix[p] = first unit of pair p.
// unit = (pair, block of MSUB*128 query rows, one of `tsplit` contiguous ranges of train tiles).
constexpr int kMaxPairsPerLaunch = 768;
struct LaunchMap {
    const uint8_t* lo;         // first descriptor row of the operand span (row index = (ptr - lo) / 32)
    int n_pairs, tsplit, n_units;
    int fused;                 // 0: keys leave as (dist, group base), hamming_resolve_kernel follows; 1: warps 14-15 resolve
    int a_packed;              // 1: query tiles are expanded from the packed descriptors by warps 14-15 (no unpacked copy of query images)
    int vec16;                 // every descriptor row is 16-byte aligned (128-bit loads in warps 14-15)
    int* group_cnt;            // fused, tsplit > 1: arrival counter per group of units sharing query rows (zero before and after the launch)
    long long* trace;          // DEBUG 7 only: per-accumulator clock64() stamps of CTA 0 (see kTraceAccs), else NULL
    int prefix[kMaxPairsPerLaunch + 1];
};

constexpr int kTraceAccs = 512;      // DEBUG 7: stamps[acc][warp 0..15][4]
__device__ __forceinline__ void trace_stamp(const LaunchMap& lm, int acc, int warp, int k) {
    if (blockIdx.x == 0 && acc < kTraceAccs && (threadIdx.x & 31) == 0) lm.trace[(acc * 16 + warp) * 4 + k] = clock64();
}

__host__ __device__ inline int units_of_pair(int n1, int n2, int tsplit, int* nts_out) {
    if (n1 <= 0 || n2 <= 0) { if (nts_out) *nts_out = 0; return 0; }
    const int tiles = (n2 + BN - 1) / BN;
    const int tper = (tiles + tsplit - 1) / tsplit;
    const int nts = (tiles + tper - 1) / tper;
    if (nts_out) *nts_out = nts;
    return ((n1 + BM * MSUB - 1) / (BM * MSUB)) * nts;
}

__device__ __forceinline__ WorkUnit make_unit(const LaunchMap& lm, const PairDesc* __restrict__ pairs, int u) {
    int lo = 0, hi = lm.n_pairs;                 // largest p with prefix[p] <= u
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (lm.prefix[mid] <= u) lo = mid; else hi = mid;
    }
    const PairDesc* pd = pairs + lo;
    const int n1 = pd->n1, n2 = pd->n2;
    int nts;
    units_of_pair(n1, n2, lm.tsplit, &nts);
    const int local = u - lm.prefix[lo];
    const int qb = local / nts, ts = local - qb * nts;
    const int tiles = (n2 + BN - 1) / BN;
    const int tper = (tiles + lm.tsplit - 1) / lm.tsplit;
    const int q0 = qb * (BM * MSUB);
    WorkUnit w;
    w.a_row0 = (uint32_t)((pd->desc1 - lm.lo) >> 5) + q0;
    w.b_row0 = (uint32_t)((pd->desc2 - lm.lo) >> 5);
    w.n_rows = min(n1 - q0, BM * MSUB);
    w.t_begin = ts * tper * BN;
    w.t_end = min((ts + 1) * tper * BN, n2);
    w.key = pd->key + q0;
    w.q_desc = pd->desc1 + (size_t)q0 * 32;
    w.t_desc = pd->desc2;
    w.n2 = n2; w.nts = nts; w.group = u - ts;
    return w;
}

constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
// InstrDescriptorBlockScaled: a_format=E2M1(1)@7, b_format=E2M1(1)@10, K-major both, n_dim=N>>3 @17,
// scale_format=UE8M0(1)@23, m_dim=M>>4 @24, a_sf_id=b_sf_id=0, k_size=0 (K64 dense)
constexpr uint32_t kIdesc = (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24);

template <int kAccumulate>
__device__ __forceinline__ void tc_mma_mxf4(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t sfa, uint32_t sfb) {
    asm volatile(
        "{\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %7, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], da, db, %4, [%5], [%6], p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(sfa), "r"(sfb), "n"(kAccumulate) : "memory");
}

// Operand encodings.  QUERY rows: bit -> +1 / -1.  TRAIN rows: bit -> +1 / 0 (SFMGMS_FP4_TRAIN01, default) -- then
// <a,b> = 2 pop(a & b) - pop(b) = pop(a) - hamming(a,b): for a fixed query row still "nearest = arg-max of the dot product", and
// the tensor pipe multiplies by zero half of the time: the MMA loop draws ~10 % less power than with +-1 x +-1 operands
// (scripts/power_probe.cu, profiles/r2_power_probe.txt), and sustained runs of this kernel sit AT the 1,000 W cap, where power
// is clock.  The exact distance of the winner is recomputed on the packed descriptors by the tie resolution either way, so
// the epilogue's value only has to order the candidates.  Query tiles are then always expanded in the kernel (the unpacked
// array in global memory holds the train encoding of every image).
#ifndef SFMGMS_FP4_TRAIN01
#define SFMGMS_FP4_TRAIN01 1
#endif
constexpr bool kTrain01 = SFMGMS_FP4_TRAIN01 != 0;

// 32 descriptor bits -> 32 e2m1 values (+1.0 = 0x2, -1.0 = 0xA), element k in nibble k (low nibble first)
__device__ __forceinline__ uint4 fp4_from_word(uint32_t w) {
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t b = (w >> (8 * q)) & 0xFFu;       // 8 bits -> 8 nibbles
        // spread bit i to bit 4*i+3 (the e2m1 sign) : set bit = +1 => sign 0; clear bit = -1 => sign 1
        uint32_t s = (b | (b << 12)) & 0x000F000Fu;       // nibbles: bits 0-3 at [0,4), bits 4-7 at [16,20)
        s = (s | (s << 6)) & 0x03030303u;                 // pairs of bits per byte
        s = (s | (s << 3)) & 0x11111111u;                 // one bit per nibble (at nibble bit 0)
        o[q] = 0x22222222u | ((~s & 0x11111111u) << 3);   // 0x2 magnitude (1.0) | sign where the bit is clear
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// train rows: set bit -> +1.0 (0x2), clear bit -> 0 (0x0) with SFMGMS_FP4_TRAIN01, else the +-1 encoding
__device__ __forceinline__ uint4 fp4_train_from_word(uint32_t w) {
    if constexpr (!kTrain01) return fp4_from_word(w);
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t b = (w >> (8 * q)) & 0xFFu;
        uint32_t s = (b | (b << 12)) & 0x000F000Fu;
        s = (s | (s << 6)) & 0x03030303u;
        s = (s | (s << 3)) & 0x11111111u;                 // one bit per nibble (at nibble bit 0)
        o[q] = s << 1;                                    // 0x2 where the bit is set
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) unpack_fp4_kernel(const uint32_t* __restrict__ desc, long long n_words,
                                                         uint4* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n_words; i += step) out[i] = fp4_train_from_word(__ldg(desc + i));
}

// Operands derived per launch: only TRAIN images need the unpacked copy in global memory (the kernel expands its query tiles
// itself).  One grid row per pair; `first` masks pairs whose train image an earlier pair of the launch already covers.
struct FirstUse { uint32_t bits[kMaxPairsPerLaunch / 32]; };
__global__ void __launch_bounds__(256) unpack_fp4_train_kernel(const PairDesc* __restrict__ pairs, const uint8_t* lo,
                                                               const __grid_constant__ FirstUse first, uint4* __restrict__ out) {
    const int p = blockIdx.y;
    if (!((first.bits[p >> 5] >> (p & 31)) & 1u)) return;
    const PairDesc& pd = pairs[p];
    if (pd.n1 <= 0 || pd.n2 <= 0) return;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(pd.desc2);
    uint4* dst = out + ((size_t)(pd.desc2 - lo) >> 5) * kDescWords;
    const int n_words = pd.n2 * kDescWords;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += 4 * stride) {   // four words in flight per thread
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = i + k * stride < n_words ? __ldg(src + i + k * stride) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k * stride < n_words) dst[i + k * stride] = fp4_train_from_word(w[k]);
    }
}

template <int dbg>   // dbg != 0: timing experiments only (results invalid); production is dbg = 0
__global__ void __launch_bounds__(kThreads, 1)
hamming_fp4_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapB,
                   const __grid_constant__ LaunchMap lm, const PairDesc* __restrict__ pairs) {
    const int n_units = lm.n_units;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = smem_base;
    const uint32_t b_smem = smem_base + A_BUFS * A_BUF_BYTES;
    const uint32_t bar_base = smem_base + SMEM_DATA;
    const uint32_t full_bar = bar_base;
    const uint32_t empty_bar = full_bar + 8 * STAGES;
    const uint32_t a_full_bar = empty_bar + 8 * STAGES;
    const uint32_t a_empty_bar = a_full_bar + 8 * A_BUFS;
    const uint32_t tfull_bar = a_empty_bar + 8 * A_BUFS;
    const uint32_t tempty_bar = tfull_bar + 8 * ACC_SLOTS;
    const uint32_t tmem_ptr_smem = tempty_bar + 8 * ACC_SLOTS;
    const uint32_t done_bar = tmem_ptr_smem + 8;      // fused tie resolution: epilogue warps -> resolver warps, a unit's candidates are in skeys
    const uint32_t free_bar = done_bar + 8;           // resolver warps -> epilogue warps: skeys may be overwritten
    uint32_t* const skeys = reinterpret_cast<uint32_t*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));   // [part][row of the unit]
    volatile uint32_t* tmem_ptr_generic =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < A_BUFS; ++s) { mbar_init(a_full_bar + 8 * s, 1); mbar_init(a_empty_bar + 8 * s, 1); }
        for (int s = 0; s < ACC_SLOTS; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, kV3 ? 4 : kEpiWarps); }
        mbar_init(done_bar, kEpiWarps); mbar_init(free_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 32) tmem_ptr_generic[1] = tfull_bar;   // read back below as an opaque value (see the epilogue)
    if (warp == kAllocWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_generic;
    // block scales: 0x7F = 2^0 everywhere (warps 4-7 cover the four lane quadrants)
    if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        uint32_t ones[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) ones[k] = 0x7F7F7F7Fu;
        const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + SF_COL;
        tc_st16(t, ones);
        tc_st16(t + 16, ones);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // The two single-issuer roles run with the WHOLE warp converged and elect one lane per issue (elect.sync): every
    // operand of the TMA / MMA instructions is then warp-uniform and ptxas keeps it in uniform registers.  (With the role
    // loop inside `if (lane == 0)` it had to wrap every UTCOMMA / UTMALDG in an R2UR + ELECT + BRA.U.ANY serialisation loop;
    // ncu showed the MMA thread spending ~75 % of its samples in that issue sequence and only ~12 % waiting on barriers.)
    if (warp == kProdWarp) {
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
        }
        uint32_t stage = 0, phase = 0, abuf = 0, a_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = make_unit(lm, pairs, u);
            const int nsub = (wu.n_rows + BM - 1) / BM;
            if (!lm.a_packed) {   // (packed: warps 14-15 fill the query tiles)
                mbar_wait(a_empty_bar + 8 * abuf, a_phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(a_full_bar + 8 * abuf, (uint32_t)(nsub * TILE_BYTES));
                    for (int s = 0; s < nsub; ++s)
                        tma_load_2d(a_smem + abuf * A_BUF_BYTES + s * TILE_BYTES, &tmap, 0, (int)(wu.a_row0 + s * BM),
                                    a_full_bar + 8 * abuf);
                }
                if (++abuf == A_BUFS) { abuf = 0; a_phase ^= 1; }
            }
            for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    if (dbg >= 5) {   // DEBUG 5/6: train tiles are not re-read (whatever the stage holds is multiplied)
                        mbar_arrive(full_bar + 8 * stage);
                    } else {
                        mbar_expect_tx(full_bar + 8 * stage, (uint32_t)BTILE_BYTES);
                        tma_load_2d(b_smem + stage * BTILE_BYTES, &tmapB, 0, (int)(wu.b_row0 + t), full_bar + 8 * stage);
                    }
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kMmaWarp) {
        // accumulator slot = sub-tile index & 1 (compile-time in the epilogue's unrolled loop); one phase bit per slot
        uint32_t stage = 0, phase = 0, abuf = 0, a_phase = 0, slot_phase[ACC_SLOTS] = {};
        int tr_acc = 0;
        const uint32_t a_lo0 = sdesc_lo(a_smem), b_lo0 = sdesc_lo(b_smem);
        const uint32_t sf = tmem_base + SF_COL;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = make_unit(lm, pairs, u);
            const int nsub = (wu.n_rows + BM - 1) / BM;
            const int ntiles = (wu.t_end - wu.t_begin + BN - 1) / BN;
            mbar_wait(a_full_bar + 8 * abuf, a_phase);
            const uint32_t a_lo_buf = a_lo0 + abuf * (A_BUF_BYTES >> 4);
            for (int ti = 0; ti < ntiles; ++ti) {
                mbar_wait(full_bar + 8 * stage, phase);
                const uint32_t b_lo = b_lo0 + stage * (BTILE_BYTES >> 4);
#pragma unroll
                for (int s = 0; s < MSUB; ++s) {
                    if (s >= nsub) break;
                    const int slot = kV3 ? s : (s & 1);
                    if (dbg == 7) trace_stamp(lm, tr_acc, warp, 0);
                    mbar_wait(tempty_bar + 8 * slot, slot_phase[slot] ^ 1);
                    slot_phase[slot] ^= 1;
                    tc_fence_after();
                    if (dbg == 7) trace_stamp(lm, tr_acc, warp, 1);
                    const uint32_t d = tmem_base + slot * BN;
                    const uint32_t a_lo = a_lo_buf + s * (TILE_BYTES >> 4);
                    if (elect_one()) {
                        if (dbg != 3) {   // DEBUG 3: no tensor work
                            tc_mma_mxf4<0>(d, a_lo + 0, b_lo + 0, kDescHi, kIdesc, sf, sf);
                            tc_mma_mxf4<1>(d, a_lo + 2, b_lo + 2, kDescHi, kIdesc, sf, sf);
                            tc_mma_mxf4<1>(d, a_lo + 4, b_lo + 4, kDescHi, kIdesc, sf, sf);
                            tc_mma_mxf4<1>(d, a_lo + 6, b_lo + 6, kDescHi, kIdesc, sf, sf);
                        }
                        tc_commit(tfull_bar + 8 * slot);
                    }
                    if (dbg == 7) trace_stamp(lm, tr_acc++, warp, 2);
                }
                if (elect_one()) tc_commit(empty_bar + 8 * stage);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(a_empty_bar + 8 * abuf);
            if (++abuf == A_BUFS) { abuf = 0; a_phase ^= 1; }
        }
    } else if (kV3 && warp < kEpiWarps) {
        // ================= epilogue, V3: warp = (lane quadrant, sub-tile = TMEM slot); all BN columns of its accumulators ====
        const int quad = warp & 3;                        // TMEM lanes [32*quad, 32*quad+32)
        const int sub = (warp - kEpiWarp0) >> 2;          // query sub-tile of the unit = accumulator slot
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * BN);
        const uint32_t tfull_e = tmem_ptr_generic[1] + 8 * sub, tempty_e = tmem_ptr_generic[1] + 8 * ACC_SLOTS + 8 * sub;
        uint32_t ph = 0, free_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = make_unit(lm, pairs, u);
            const int nsub = (wu.n_rows + BM - 1) / BM;
            float best_key = -1.0e30f, beat = -1.0e29f;   // see the two-slot epilogue below for the key / tag scheme
            int best_base = -1;
            // kEpiCols columns starting at accumulator column col0 of the tile at train row t: group-tagged maximum, one-compare update
            auto reduce = [&](const int (&r)[kEpiCols], int col0, int t, int valid) {
                float v[kEpiCols];
#pragma unroll
                for (int j = 0; j < kEpiCols; ++j) v[j] = __int_as_float(r[j]);
                if (col0 + kEpiCols > valid) {                       // tail tile: mask columns outside the image
#pragma unroll
                    for (int j = 0; j < kEpiCols; ++j)
                        if (col0 + j >= valid) v[j] = -1.0e30f;
                }
                float gm[kGroups + 1];
#pragma unroll
                for (int g = 0; g + 1 < kGroups; ++g) gm[g] = max_tree<kGroup>(v + kGroup * g);
                gm[kGroups - 1] = max_tree<kLastGroup>(v + kGroup * (kGroups - 1));
                gm[kGroups] = 0.f;
#pragma unroll
                for (int k = 0; k < kGroups / 2; ++k) add2(gm[2 * k], gm[2 * k + 1], c_grouptag[k], gm[2 * k], gm[2 * k + 1]);
                if constexpr (kGroups & 1) gm[kGroups - 1] += c_grouptag[kGroups / 2].x;
                const float m = max_tree<kGroups>(gm);
                if (m > beat) { best_key = m; beat = m + 0.5f; best_base = t + col0; }
            };
            if (sub < nsub) {
                for (int t = wu.t_begin; t < wu.t_end; t += BN) {
                    const int valid = wu.t_end - t;
                    mbar_wait(tfull_e, ph);
                    ph ^= 1;
                    tc_fence_after();
                    int r[kEpiCols];
#pragma unroll
                    for (int c = 0; c < BN / kEpiCols; ++c) {
                        ld_part(tbase + c * kEpiCols, r);
                        tc_wait_ld();
                        if (c == BN / kEpiCols - 1) {                // everything is in registers: hand the slot back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(tempty_e);
                        }
                        reduce(r, c * kEpiCols, t, valid);
                    }
                }
            }
            const bool fused = dbg == 0 && lm.fused != 0;
            if (fused) {
                mbar_wait(free_bar, free_phase ^ 1);            // the helper warps have taken the previous unit's candidates
                free_phase ^= 1;
            }
            {
                const int row = sub * BM + quad * 32 + lane;
                const bool have = sub < nsub && row < wu.n_rows && (best_base >= 0 || dbg != 0);
                uint32_t packed = kKeyInit;
                if (have) {
                    const float fl = floorf(best_key);
                    const int g = kTagDen - 1 - (int)((best_key - fl) * (float)kTagDen);
                    constexpr bool kTiming = dbg != 0;
                    const uint32_t dist = kTiming ? ((uint32_t)(256 - (int)fl) >> 1) & 7u : kTrain01 ? (uint32_t)(256 - (int)fl) : (uint32_t)(256 - (int)fl) >> 1;
                    const uint32_t idx = kTiming ? (uint32_t)(best_base + kGroup * g) & 7u : (uint32_t)(best_base + kGroup * g);
                    packed = (dist << kTrainIdxBits) | idx;
                }
                if (fused) skeys[row] = packed;
                else if (have) atomicMin(wu.key + row, packed);
            }
            if (fused) {
                __syncwarp();
                if (lane == 0) mbar_arrive(done_bar);
            }
        }
    } else if (!kV3 && warp < kEpiWarps) {
        // ================= epilogue: 12 warps; warp = (lane quadrant, 80-column part) of every accumulator ==========
        const int quad = warp & 3;                        // TMEM lanes [32*quad, 32*quad+32)
        const int c0 = ((warp - kEpiWarp0) >> 2) * kEpiCols; // accumulator columns [c0, c0+80)
        const uint32_t tbase = tmem_base + ((uint32_t)(quad * 32) << 16) + c0;
        // the accumulator barriers' address as a LOADED value: at the 128-register cap ptxas otherwise re-derives the
        // shared-memory base (6 instructions) in front of every wait of the loop
        const uint32_t tfull_e = tmem_ptr_generic[1], tempty_e = tfull_e + 8 * ACC_SLOTS;
        uint32_t slot_phase[ACC_SLOTS] = {0, 0};
        uint32_t free_phase = 0;
        int tr_acc = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const WorkUnit wu = make_unit(lm, pairs, u);
            const int nsub = (wu.n_rows + BM - 1) / BM;
            // per row: best key so far (dot + index fraction), the threshold a key must exceed to beat it (best + 1/2:
            // strict '>' on the dot product keeps the earliest tile / part) and the column base it came from
            float best_key[MSUB], beat[MSUB];
            int best_base[MSUB];
#pragma unroll
            for (int s = 0; s < MSUB; ++s) { best_key[s] = -1.0e30f; beat[s] = -1.0e29f; best_base[s] = -1; }
            // one accumulator (sub-tile S of the tile at train row t): wait, TMEM -> registers, hand the slot back, reduce.
            // kTail: the tile may reach past the image (columns >= valid are masked).
            auto pass = [&](auto s_c, auto tail_c, uint32_t parity, int t, int valid) {
                constexpr int S = decltype(s_c)::value;
                constexpr bool kTail = decltype(tail_c)::value;
                constexpr int slot = S & 1;
                if (dbg == 7) trace_stamp(lm, tr_acc, warp, 0);
                mbar_wait(tfull_e + 8 * slot, parity);
                tc_fence_after();
                if (dbg == 7) trace_stamp(lm, tr_acc, warp, 1);
                int r[kEpiCols];
                if (dbg == 4 || dbg == 6) {    // DEBUG 4/6: no TMEM read, no ALU
#pragma unroll
                    for (int j = 0; j < kEpiCols; ++j) r[j] = 0;
                } else {
                    ld_part(tbase + slot * BN, r);
                    tc_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_e + 8 * slot);   // values are in registers now
                if (dbg == 7) trace_stamp(lm, tr_acc, warp, 2);
                if (dbg == 4 || dbg == 6) return;
                float v[kEpiCols];
#pragma unroll
                for (int j = 0; j < kEpiCols; ++j) v[j] = __int_as_float(r[j]);
                if (dbg == 1) { best_key[S] = fmaxf(best_key[S], v[0] + v[kEpiCols - 1]); return; }   // DEBUG 1: no max tree
                if (kTail && c0 + kEpiCols > valid) {                // tail tile: mask columns outside the image
#pragma unroll
                    for (int j = 0; j < kEpiCols; ++j)
                        if (c0 + j >= valid) v[j] = -1.0e30f;
                }
                float gm[kGroups + 1];
#pragma unroll
                for (int g = 0; g + 1 < kGroups; ++g) gm[g] = max_tree<kGroup>(v + kGroup * g);
                gm[kGroups - 1] = max_tree<kLastGroup>(v + kGroup * (kGroups - 1));
                gm[kGroups] = 0.f;
#pragma unroll
                for (int k = 0; k < kGroups / 2; ++k) add2(gm[2 * k], gm[2 * k + 1], c_grouptag[k], gm[2 * k], gm[2 * k + 1]);
                if constexpr (kGroups & 1) gm[kGroups - 1] += c_grouptag[kGroups / 2].x;
                const float m = max_tree<kGroups>(gm);
                if (m > beat[S]) {                                   // integer part exceeds the best so far (see kTagDen)
                    best_key[S] = m;
                    beat[S] = m + 0.5f;
                    best_base[S] = t + c0;
                }
                if (dbg == 7) { if (best_key[S] == 12345.f) tr_acc += 1 << 20; trace_stamp(lm, tr_acc++, warp, 3); }
            };
            using std::integral_constant;
            int t = wu.t_begin;
            if (MSUB == 4 && nsub == MSUB) {
                // full unit, whole tiles: each slot is used twice per tile, so the barrier parities repeat tile after tile
                // (no phase bookkeeping, no sub-tile or tail tests in the loop)
                const uint32_t p0 = slot_phase[0], p1 = slot_phase[1];
                // (the always-true row tests keep every pass in its own basic block: in straight-line code ptxas hoists
                // the next pass's try_wait above the current max tree, the warp is then SUSPENDED with that ALU work
                // undone, and the slot comes back late -- measured 3 % slower)
                for (; t + BN <= wu.t_end; t += BN) {
                    pass(integral_constant<int, 0>{}, integral_constant<bool, false>{}, p0, t, BN);
                    if (wu.n_rows > 1 * BM) pass(integral_constant<int, 1>{}, integral_constant<bool, false>{}, p1, t, BN);
                    if (wu.n_rows > 2 * BM) pass(integral_constant<int, 2 % MSUB>{}, integral_constant<bool, false>{}, p0 ^ 1, t, BN);
                    if (wu.n_rows > 3 * BM) pass(integral_constant<int, 3 % MSUB>{}, integral_constant<bool, false>{}, p1 ^ 1, t, BN);
                }
            }
            for (; t < wu.t_end; t += BN) {
                const int valid = wu.t_end - t;
#pragma unroll
                for (int s = 0; s < MSUB; ++s) {
                    if (s < nsub) {
                        const int slot = s & 1;                              // compile-time after unrolling
                        const uint32_t parity = slot_phase[slot];
                        slot_phase[slot] ^= 1;
                        if (s == 0) pass(integral_constant<int, 0>{}, integral_constant<bool, true>{}, parity, t, valid);
                        if (s == 1) pass(integral_constant<int, 1 % MSUB>{}, integral_constant<bool, true>{}, parity, t, valid);
                        if (s == 2) pass(integral_constant<int, 2 % MSUB>{}, integral_constant<bool, true>{}, parity, t, valid);
                        if (s == 3) pass(integral_constant<int, 3 % MSUB>{}, integral_constant<bool, true>{}, parity, t, valid);
                    }
                }
            }
            // the unit's result per row: key = dot + (kTagDen - 1 - g)/kTagDen : <a,b> = 256 - 2*hamming; idx = first train row of
            // the winning group.  Fused: the three column parts leave their candidates in shared memory for the resolver
            // warps (no atomics, no fence on this path); otherwise atomicMin into the global key array.
            const bool fused = dbg == 0 && lm.fused != 0;
            if (fused) {
                mbar_wait(free_bar, free_phase ^ 1);            // the resolver warps have taken the previous unit's candidates
                free_phase ^= 1;
            }
#pragma unroll
            for (int s = 0; s < MSUB; ++s) {
                const int row = s * BM + quad * 32 + lane;
                const bool have = s < nsub && row < wu.n_rows && (best_base[s] >= 0 || (dbg != 0 && dbg != 7));   // (dbg: the values are junk, the work is real)
                uint32_t packed = kKeyInit;
                if (have) {
                    const float fl = floorf(best_key[s]);
                    const int g = kTagDen - 1 - (int)((best_key[s] - fl) * (float)kTagDen);
                    constexpr bool kTiming = dbg != 0 && dbg != 7;   // timing builds: keep the work alive, keep the key in range
                    constexpr bool kNoData = dbg == 4 || dbg == 6;
                    // (+-1 x +-1: dot = 256 - 2 d; +-1 x {0,1}: dot = pop(a) - d, and 256 - dot orders the candidates of a row like d does --
                    // the resolution pass replaces it by the recomputed distance)
                    const uint32_t dist = kNoData ? 7u : kTiming ? ((uint32_t)(256 - (int)fl) >> 1) & 7u
                                        : kTrain01 ? (uint32_t)(256 - (int)fl) : (uint32_t)(256 - (int)fl) >> 1;
                    const uint32_t idx = kNoData ? 0u : kTiming ? (uint32_t)(best_base[s] + kGroup * g) & 7u : (uint32_t)(best_base[s] + kGroup * g);
                    packed = (dist << kTrainIdxBits) | idx;
                }
                if (fused) skeys[((warp - kEpiWarp0) >> 2) * (MSUB * BM) + row] = packed;
                else if (have) atomicMin(wu.key + row, packed);
            }
            if (fused) {
                __syncwarp();
                if (lane == 0) mbar_arrive(done_bar);
            }
        }
    } else if (dbg == 0 && (lm.fused != 0 || lm.a_packed != 0)) {
        // ================= fused tie resolution: warps 14-15, one unit behind the epilogue ==========================
        // Per unit: merge the column parts' candidates (min of dist << 18 | group base) into the global key array -- a plain
        // store when the unit covers the whole train image, atomicMin + an arrival counter when several units (train-range
        // splits, other CTAs) share the query rows: the LAST unit of the group to arrive resolves.  Resolving = BFMatcher's
        // tie rule on the packed 256-bit descriptors: kGroup lanes per query row re-score the kGroup train rows of the winning
        // group (XOR / POPC) and the key becomes (distance << 18 | lowest train row attaining it).
        // One lane per query row (8 rows per lane and unit).  The candidates' cache lines are PREFETCHED to L2 for all of a
        // lane's rows before the first one is scored: with one row in flight per lane the two warps could not keep up with the
        // epilogue whenever the descriptors come from DRAM (every image new, as in the streaming batches) -- measured 20 %
        // slower than the separate kernel; the unit is then bound by 64 x 8 dependent DRAM round trips.
        const int rt = (int)threadIdx.x - kResWarp0 * 32;
        const bool vec = lm.vec16 != 0;
        volatile uint32_t* s_last = tmem_ptr_generic + 8;
        constexpr int kRowsPerLane = MSUB * BM / kResThreads;   // 8
        uint32_t done_phase = 0;
        auto resolve_unit = [&](const WorkUnit& wu) {
            mbar_wait(done_bar, done_phase);
            done_phase ^= 1;
            uint32_t mk[kRowsPerLane];
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const int row = rt + i * kResThreads;
                uint32_t k = kKeyInit;
                if (row < wu.n_rows) {
                    k = skeys[row];
#pragma unroll
                    for (int part = 1; part < kKeyParts; ++part) k = min(k, skeys[part * (MSUB * BM) + row]);
                    if (wu.nts > 1 && k != kKeyInit) atomicMin(wu.key + row, k);
                }
                mk[i] = k;
            }
            if (wu.nts > 1) __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(kResThreads) : "memory");
            if (rt == 0) mbar_arrive(free_bar);
            if (wu.nts > 1) {
                if (rt == 0) {
                    const bool last = atomicAdd(lm.group_cnt + wu.group, 1) == wu.nts - 1;
                    if (last) lm.group_cnt[wu.group] = 0;       // leave the counters as they were found
                    *s_last = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kResThreads) : "memory");
                if (*s_last == 0u) return;                     // (the next write of s_last comes after the next bar.sync)
                __threadfence();
#pragma unroll
                for (int i = 0; i < kRowsPerLane; ++i) {
                    const int row = rt + i * kResThreads;
                    mk[i] = row < wu.n_rows ? __ldcg(wu.key + row) : kKeyInit;
                }
            }
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                if (mk[i] == kKeyInit) continue;
                const int row = rt + i * kResThreads;
                const uint32_t c0 = mk[i] & kTrainIdxMask;
                const uint8_t* t = wu.t_desc + (size_t)c0 * 32;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(wu.q_desc + (size_t)row * 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(t));
                if ((int)c0 + 4 < wu.n2) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + 128));
                if ((int)c0 + kGroup - 1 < wu.n2) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + (kGroup - 1) * 32));   // (unaligned groups span 3 lines)
                if (kGroup > 9 && (int)c0 + 8 < wu.n2) asm volatile("prefetch.global.L2 [%0];" ::"l"(t + 256));
            }
#pragma unroll
            for (int i = 0; i < kRowsPerLane; ++i) {
                const uint32_t key = mk[i];
                if (key == kKeyInit) continue;
                const int row = rt + i * kResThreads;
                const int c0 = (int)(key & kTrainIdxMask);
                const uint8_t* q = wu.q_desc + (size_t)row * 32;
                const uint8_t* t = wu.t_desc + (size_t)c0 * 32;
                uint32_t a[8];
                uint32_t best = 0xFFFFFFFFu;
                if (vec) {
                    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(q)), a1 = __ldg(reinterpret_cast<const uint4*>(q) + 1);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    // candidates in two batches of loads (the lane keeps at most 5 rows = 40 registers in flight)
                    constexpr int kHalf = (kGroup + 1) / 2;
#pragma unroll
                    for (int cb = 0; cb < kGroup; cb += kHalf) {
                        uint4 b0[kHalf], b1[kHalf];
#pragma unroll
                        for (int c = 0; c < kHalf; ++c) {
                            b0[c] = b1[c] = make_uint4(0u, 0u, 0u, 0u);
                            if (cb + c < kGroup && c0 + cb + c < wu.n2) {
                                b0[c] = __ldg(reinterpret_cast<const uint4*>(t) + 2 * (cb + c));
                                b1[c] = __ldg(reinterpret_cast<const uint4*>(t) + 2 * (cb + c) + 1);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < kHalf; ++c) {
                            const uint32_t d = __popc(a[0] ^ b0[c].x) + __popc(a[1] ^ b0[c].y) + __popc(a[2] ^ b0[c].z) + __popc(a[3] ^ b0[c].w) +
                                               __popc(a[4] ^ b1[c].x) + __popc(a[5] ^ b1[c].y) + __popc(a[6] ^ b1[c].z) + __popc(a[7] ^ b1[c].w);
                            if (cb + c < kGroup && c0 + cb + c < wu.n2) best = min(best, (d << kCandBits) | (uint32_t)(cb + c));
                        }
                    }
                } else {
                    const uint32_t* qw = reinterpret_cast<const uint32_t*>(q);
#pragma unroll
                    for (int w = 0; w < 8; ++w) a[w] = __ldg(qw + w);
#pragma unroll
                    for (int c = 0; c < kGroup; ++c) {
                        if (c0 + c < wu.n2) {
                            const uint32_t* tw = reinterpret_cast<const uint32_t*>(t) + 8 * c;
                            uint32_t d = 0;
#pragma unroll
                            for (int w = 0; w < 8; ++w) d += __popc(a[w] ^ __ldg(tw + w));
                            best = min(best, (d << kCandBits) | (uint32_t)c);
                        }
                    }
                }
                wu.key[row] = ((best >> kCandBits) << kTrainIdxBits) | ((uint32_t)c0 + (best & ((1u << kCandBits) - 1u)));
            }
        };
        // ---- query tiles from the packed descriptors (a_packed): the four 128-row sub-tiles of a unit, written in the layout a
        // 128B-swizzled TMA box load would leave (row r of a tile at r * 128, its 16-byte chunk c at (c ^ (r & 7)) * 16; one
        // chunk = one descriptor word = 32 e2m1 values), then handed to the MMA warp through the same a_full barrier.  Double
        // buffered like the TMA path: the tiles of unit k+2 are written while unit k+1 runs.
        uint32_t abuf = 0, a_phase = 0;
        uint8_t* const a_generic = smem_raw + (a_smem - smem_u32(smem_raw));
        // byte -> 8 e2m1 values through a 1 KB table (4 shared-memory loads per descriptor word instead of ~24 ALU instructions:
        // these two warps share their schedulers' issue slots with six epilogue warps)
        const uint32_t* const lut = skeys + SMEM_KEYS / 4;
        if (lm.a_packed) {
            for (int b = rt; b < 256; b += kResThreads) const_cast<uint32_t*>(lut)[b] = fp4_from_word((uint32_t)b).x;
            asm volatile("bar.sync 1, %0;" ::"n"(kResThreads) : "memory");
        }
        auto expand = [&](uint32_t w) {
            return make_uint4(lut[w & 0xFFu], lut[(w >> 8) & 0xFFu], lut[(w >> 16) & 0xFFu], lut[w >> 24]);
        };
        auto fill_unit = [&](const WorkUnit& wu) {
            mbar_wait(a_empty_bar + 8 * abuf, a_phase ^ 1);
            uint8_t* const dst = a_generic + abuf * A_BUF_BYTES;
            if (vec) {
                const int n_half = wu.n_rows * 2;                            // i = (row, half): 16 packed bytes -> 64 unpacked
                for (int i0 = rt; i0 < n_half; i0 += 4 * kResThreads) {        // four loads in flight per lane
                    uint4 w[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        w[k] = i0 + k * kResThreads < n_half ? __ldg(reinterpret_cast<const uint4*>(wu.q_desc) + i0 + k * kResThreads)
                                                             : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int i = i0 + k * kResThreads;
                        if (i >= n_half) break;
                        const int row = i >> 1, h = i & 1;
                        uint8_t* const rp = dst + row * ROWB;
                        const int sw = row & 7;
                        *reinterpret_cast<uint4*>(rp + (((4 * h + 0) ^ sw) << 4)) = expand(w[k].x);
                        *reinterpret_cast<uint4*>(rp + (((4 * h + 1) ^ sw) << 4)) = expand(w[k].y);
                        *reinterpret_cast<uint4*>(rp + (((4 * h + 2) ^ sw) << 4)) = expand(w[k].z);
                        *reinterpret_cast<uint4*>(rp + (((4 * h + 3) ^ sw) << 4)) = expand(w[k].w);
                    }
                }
            } else {
                for (int i = rt; i < wu.n_rows * kDescWords; i += kResThreads) {
                    const int row = i >> 3, c = i & 7;
                    *reinterpret_cast<uint4*>(dst + row * ROWB + ((c ^ (row & 7)) << 4)) =
                        expand(__ldg(reinterpret_cast<const uint32_t*>(wu.q_desc) + i));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
            asm volatile("bar.sync 1, %0;" ::"n"(kResThreads) : "memory");
            if (rt == 0) mbar_arrive(a_full_bar + 8 * abuf);
            if (++abuf == A_BUFS) { abuf = 0; a_phase ^= 1; }
        };
        const int ustep = (int)gridDim.x;
        if (lm.a_packed) {
            for (int k = 0; k < A_BUFS; ++k)
                if ((int)blockIdx.x + k * ustep < n_units) fill_unit(make_unit(lm, pairs, (int)blockIdx.x + k * ustep));
        }
        for (int u = blockIdx.x; u < n_units; u += ustep) {
            if (lm.fused) resolve_unit(make_unit(lm, pairs, u));
            if (lm.a_packed && u + A_BUFS * ustep < n_units) fill_unit(make_unit(lm, pairs, u + A_BUFS * ustep));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

// Small launches (a pair or two) keep the tie resolution as a kernel of its own: after the tensor-core pass every key holds
// (distance << 18 | first train row of the winning group).  Sixteen lanes per query row (kGroup of them busy) recompute the
// candidates' Hamming distances on the original 256-bit descriptors and the key becomes (distance << 18 | lowest train row
// attaining it) -- BFMatcher's tie rule.
constexpr int kResolveThreads = 128;
constexpr int kResolveLanes = 16;
template <bool kVec>   // kVec: every descriptor row is 16-byte aligned (two 128-bit loads per row)
__global__ void __launch_bounds__(kResolveThreads) hamming_resolve_kernel(const PairDesc* __restrict__ pairs) {
    const PairDesc& pd = pairs[blockIdx.y];
    const int row = blockIdx.x * (kResolveThreads / kResolveLanes) + (threadIdx.x / kResolveLanes);
    const int l = threadIdx.x & (kResolveLanes - 1);
    if (pd.n1 <= 0 || pd.n2 <= 0 || row >= pd.n1) return;      // whole 16-lane groups leave together
    const uint32_t key = pd.key[row];
    const int cand = (int)(key & kTrainIdxMask) + l;
    uint32_t d = 0x1FFFu;
    if (l < kGroup && cand < pd.n2) {
        d = 0;
        if constexpr (kVec) {
            const uint4* q = reinterpret_cast<const uint4*>(pd.desc1) + (size_t)row * 2;
            const uint4* t = reinterpret_cast<const uint4*>(pd.desc2) + (size_t)cand * 2;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 a = __ldg(q + h), b = __ldg(t + h);
                d += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
            }
        } else {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(pd.desc1) + (size_t)row * kDescWords;
            const uint32_t* t = reinterpret_cast<const uint32_t*>(pd.desc2) + (size_t)cand * kDescWords;
#pragma unroll
            for (int w = 0; w < kDescWords; ++w) d += __popc(__ldg(q + w) ^ __ldg(t + w));
        }
    }
    uint32_t k = (d << kCandBits) | (uint32_t)l;
    const uint32_t mask = 0xFFFFu << (threadIdx.x & 16);
#pragma unroll
    for (int o = 1; o < kResolveLanes; o <<= 1) k = min(k, __shfl_xor_sync(mask, k, o, kResolveLanes));
    if (l == 0) pd.key[row] = ((k >> kCandBits) << kTrainIdxBits) | ((key & kTrainIdxMask) + (k & ((1u << kCandBits) - 1u)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
thread_local char g_err[256] = "";   // one context per host thread: messages never interleave
EncodeTiledFn g_encode = nullptr;

bool load_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
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
    if (cudaMalloc(&p, want) != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaMalloc(%zu) failed", want); return false; }
    cap = want;
    return true;
}

}  // namespace

const char* fp4_last_error() { return g_err; }

// SFMGMS_FP4_PACKED_QUERIES=0: launches that derive their operands unpack query images too and load them with TMA
static bool fp4_packed_queries() {
    static const bool on = !(getenv("SFMGMS_FP4_PACKED_QUERIES") && atoi(getenv("SFMGMS_FP4_PACKED_QUERIES")) == 0);
    return on;
}

bool fp4_fused_resolve() {
    static const bool on = !(getenv("SFMGMS_FP4_FUSED_RESOLVE") && atoi(getenv("SFMGMS_FP4_FUSED_RESOLVE")) == 0);
    return on;
}

int launch_hamming_fp4(TcState& s, const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                       cudaStream_t st, cudaStream_t resolve_st, cudaEvent_t resolve_ev) {
    if (n_pairs <= 0) return 0;
    if (!load_encode()) return -1;
    int launches = 0;
    const uint8_t* lo = nullptr;
    const uint8_t* hi = nullptr;
    long long ref_rows = 0;
    bool aligned16 = true;   // descriptor rows are 32 bytes: a 16-byte aligned base keeps every row aligned
    for (int p = 0; p < n_pairs; ++p) {
        const PairDesc& pd = h_pairs[p];
        ref_rows += (long long)pd.n1 + pd.n2;
        if (pd.n1 <= 0 || pd.n2 <= 0) continue;
        if ((reinterpret_cast<uintptr_t>(pd.desc1) | reinterpret_cast<uintptr_t>(pd.desc2)) & 15) aligned16 = false;
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
    if (n_rows >= (1ll << 31) || (!pinned && n_rows > 64 * ref_rows + 4096)) {
        snprintf(g_err, sizeof g_err, "descriptor operands are not in one contiguous array");
        return -1;
    }
    // work decomposition: pick the train-range split that minimises the static-schedule makespan
    // ceil(units / #SM) / tsplit (small sub-batches of the pipelined host path otherwise lose a whole round)
    long long qblocks = 0;
    int min_tiles = 1 << 30;
    for (int p = 0; p < n_pairs; ++p) {
        if (h_pairs[p].n1 <= 0 || h_pairs[p].n2 <= 0) continue;
        qblocks += (h_pairs[p].n1 + BM * MSUB - 1) / (BM * MSUB);
        const int tiles = (h_pairs[p].n2 + BN - 1) / BN;
        if (tiles < min_tiles) min_tiles = tiles;
    }
    // Small launches (a pair or two: every CTA gets about one short unit) keep the separate unpack / resolve kernels -- there the
    // helper warps' work would sit on the critical path of a 20-us kernel; batches let warps 14-15 do both jobs in the shadow
    // of the tensor work.
    const bool batch = qblocks >= 2LL * sm_count;
    static const int dbg = getenv("SFMGMS_TC_DEBUG") ? atoi(getenv("SFMGMS_TC_DEBUG")) : 0;   // timing experiments only
    if (n_pairs > kMaxPairsPerLaunch) { snprintf(g_err, sizeof g_err, "too many pairs per launch"); return -1; }
    // train encoding {0,1}: the unpacked array never holds query operands, every production launch expands its query tiles
    bool a_packed = kTrain01 && dbg == 0;
    if (!(s.set_valid && s.ops_src == lo && s.ops_rows == n_rows && s.ops_row_bytes == ROWB)) {
        if (!ensure_dev(s.d_ops, s.ops_cap, (size_t)(n_rows + 256) * ROWB)) return -1;
        if (pinned || dbg != 0 || (!kTrain01 && (!batch || !fp4_packed_queries()))) {
            // a registered image set with the operand cache on: every image once, valid for all later launches
            const long long n_words = n_rows * kDescWords;
            long long blocks = (n_words + 255) / 256;
            if (blocks > 148 * 32) blocks = 148 * 32;
            unpack_fp4_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(lo), n_words,
                                                                static_cast<uint4*>(s.d_ops));
            ++launches; kmark("unpack_fp4", st);
            s.ops_src = lo; s.ops_rows = n_rows; s.ops_row_bytes = ROWB;
            s.set_valid = s.cache_enabled;
        } else {
            // operands derived for this launch only: the train images (each once), the kernel expands its query tiles itself
            FirstUse fu;
            memset(&fu, 0, sizeof fu);
            std::vector<std::pair<const uint8_t*, int>> seen;
            seen.reserve((size_t)n_pairs);
            for (int p = 0; p < n_pairs; ++p)
                if (h_pairs[p].n1 > 0 && h_pairs[p].n2 > 0) seen.push_back({h_pairs[p].desc2, p});
            std::sort(seen.begin(), seen.end());
            int max_n2 = 0;
            for (size_t k = 0; k < seen.size(); ++k) {
                if (k > 0 && seen[k].first == seen[k - 1].first && h_pairs[seen[k].second].n2 <= h_pairs[seen[k - 1].second].n2) {
                    seen[k].second = seen[k - 1].second;   // same image (rows of the earlier entry cover it)
                    continue;
                }
                const int p = seen[k].second;
                fu.bits[p >> 5] |= 1u << (p & 31);
                if (h_pairs[p].n2 > max_n2) max_n2 = h_pairs[p].n2;
            }
            int bx = (max_n2 * kDescWords + 256 * 16 - 1) / (256 * 16);   // ~16 words per thread
            if (bx < 1) bx = 1;
            unpack_fp4_train_kernel<<<dim3((unsigned)bx, (unsigned)n_pairs), 256, 0, st>>>(d_pairs, lo, fu, static_cast<uint4*>(s.d_ops));
            ++launches; kmark("unpack_fp4", st);
            s.set_valid = false;
            a_packed = true;
        }
    }
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)ROWB, (cuuint64_t)(n_rows + 256)};
        const cuuint64_t gstride[1] = {(cuuint64_t)ROWB};
        const cuuint32_t box[2] = {(cuuint32_t)ROWB, (cuuint32_t)BM};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, s.d_ops, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    }
    CUtensorMap tmapB;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)ROWB, (cuuint64_t)(n_rows + 256)};
        const cuuint64_t gstride[1] = {(cuuint64_t)ROWB};
        const cuuint32_t box[2] = {(cuuint32_t)ROWB, (cuuint32_t)BN};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = g_encode(&tmapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, s.d_ops, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled(B) failed (%d)", (int)r); return -1; }
    }
    int tsplit = 1;
    {
        double best = 1e30;
        const int ts_max = qblocks < 2LL * sm_count ? 64 : 4;
        for (int ts = 1; ts <= ts_max && ts <= (min_tiles > 0 ? min_tiles : 1); ++ts) {
            const long long units = qblocks * ts;
            const double makespan = (double)((units + sm_count - 1) / sm_count) / ts * (1.0 + 0.01 * (ts - 1));
            if (makespan < best - 1e-9) { best = makespan; tsplit = ts; }
        }
    }
    if (n_pairs > kMaxPairsPerLaunch) { snprintf(g_err, sizeof g_err, "too many pairs per launch"); return -1; }
    LaunchMap lm;
    lm.lo = lo; lm.n_pairs = n_pairs; lm.tsplit = tsplit; lm.fused = 0; lm.a_packed = 0; lm.vec16 = aligned16 ? 1 : 0; lm.group_cnt = nullptr; lm.trace = nullptr;
    int n_units = 0;
    for (int p = 0; p < n_pairs; ++p) {
        lm.prefix[p] = n_units;
        n_units += units_of_pair(h_pairs[p].n1, h_pairs[p].n2, tsplit, nullptr);
    }
    lm.prefix[n_pairs] = n_units;
    lm.n_units = n_units;
    if (n_units == 0) return launches;
    lm.a_packed = a_packed ? 1 : 0;
    auto kern = dbg == 1 ? hamming_fp4_kernel<1> : dbg == 3 ? hamming_fp4_kernel<3> : dbg == 4 ? hamming_fp4_kernel<4>
              : dbg == 5 ? hamming_fp4_kernel<5> : dbg == 6 ? hamming_fp4_kernel<6> : dbg == 7 ? hamming_fp4_kernel<7>
                                                                                         : hamming_fp4_kernel<0>;
    if (dbg == 7) {   // DEBUG 7: clock stamps of CTA 0's first kTraceAccs accumulators -> $SFMGMS_TC_TRACE (raw int64)
        static long long* d_trace = nullptr;
        const size_t tb = sizeof(long long) * kTraceAccs * 16 * 4;
        if (!d_trace) cudaMalloc(&d_trace, tb);
        cudaMemsetAsync(d_trace, 0, tb, st);
        lm.trace = d_trace;
    }
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute(smem=%d) failed", SMEM_BYTES);
        return -1;
    }
    const bool fused = dbg == 0 && batch && fp4_fused_resolve();
    if (fused) {
        lm.fused = 1;
        if (tsplit > 1) {   // units that share query rows count their arrivals; the last one resolves and re-zeroes its counter
            const size_t need = (size_t)n_units * sizeof(int);
            if (need > s.cnt_cap) {
                if (!ensure_dev(s.d_cnt, s.cnt_cap, need)) return -1;
                if (cudaMemsetAsync(s.d_cnt, 0, s.cnt_cap, st) != cudaSuccess) { snprintf(g_err, sizeof g_err, "cudaMemsetAsync(counters) failed"); return -1; }
            }
            lm.group_cnt = static_cast<int*>(s.d_cnt);
        }
    }
    const int grid = n_units < sm_count ? n_units : sm_count;
    kern<<<grid, kThreads, SMEM_BYTES, st>>>(tmap, tmapB, lm, d_pairs);
    kmark("hamming_fp4", st);
    if (dbg == 7 && getenv("SFMGMS_TC_TRACE")) {
        const size_t tb = sizeof(long long) * kTraceAccs * 16 * 4;
        std::vector<long long> h(tb / sizeof(long long));
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), lm.trace, tb, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(getenv("SFMGMS_TC_TRACE"), "wb")) { fwrite(h.data(), 1, tb, f); fclose(f); }
    }
    if ((dbg && dbg != 7) || fused) return launches + 1;
    int max_n1 = 0;
    for (int p = 0; p < n_pairs; ++p)
        if (h_pairs[p].n2 > 0 && h_pairs[p].n1 > max_n1) max_n1 = h_pairs[p].n1;
    // The exact tie resolution is a small L2-bound kernel.  Given a second stream it runs BESIDE the next tensor-core launch
    // (its 128-thread, 32-register CTAs fit into the registers the persistent kernel leaves free on every SM).
    cudaStream_t rst = st;
    if (resolve_st && resolve_ev && !tl_marks) {
        if (cudaEventRecord(resolve_ev, st) != cudaSuccess || cudaStreamWaitEvent(resolve_st, resolve_ev, 0) != cudaSuccess) {
            snprintf(g_err, sizeof g_err, "event hand-over to the resolve stream failed");
            return -1;
        }
        rst = resolve_st;
    }
    const int rows_per_block = kResolveThreads / kResolveLanes;
    const dim3 rgrid((unsigned)((max_n1 + rows_per_block - 1) / rows_per_block), (unsigned)n_pairs);
    if (aligned16) hamming_resolve_kernel<true><<<rgrid, kResolveThreads, 0, rst>>>(d_pairs);
    else hamming_resolve_kernel<false><<<rgrid, kResolveThreads, 0, rst>>>(d_pairs);
    kmark("hamming_resolve", st);
    return launches + 2;
}

}  // namespace sfmgms
