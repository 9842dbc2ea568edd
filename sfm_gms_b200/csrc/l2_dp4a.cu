// l2_dp4a.cu — brute-force L2 nearest neighbour for integer-valued descriptors (OpenCV SIFT), SURVEY §8f-3.
//
// The reference's literal main path is SIFT + cv::BFMatcher(NORM_L2)::match (FeatureMatchUtil.cpp:10, 66-68).
// OpenCV's SIFT descriptors are float32 but integer-valued in [0,255]; for such data every partial sum of
// sum((a-b)^2) is an integer below 2^24, so OpenCV's float accumulation is exact whatever its SIMD order and
//     distance = sqrtf((float)d2),  d2 = |a|^2 + |b|^2 - 2<a,b>  (exact integer),
// which this kernel reproduces bit for bit: descriptors are narrowed to u8 (validated: anything that is not an
// integer in [0,255] is rejected), <a,b> runs on DP4A (4 u8 MACs per lane-instruction), the per-row result is the
// 64-bit key (d2 << 18 | trainIdx) minimised with atomicMin (lowest trainIdx on ties, order-free).
// First cut on the CUDA cores; the tcgen05 (kind::i8, unsigned) variant with the |b|^2 bias folded into the
// epilogue is the planned follow-up.
#include "common.cuh"

namespace sfmgms {

namespace {

constexpr int kDim = 128;                 // SIFT
constexpr int kWords = kDim / 4;          // 32 x (4 x u8)
constexpr int kThreads = 128;
constexpr int kQPT = 2;                   // queries per thread (2 x 32 registers)
constexpr int kQTile = kThreads * kQPT;
constexpr int kTTile = 64;                // train rows per shared-memory stage (8 KB + norms)

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
    dst[(size_t)row * kWords + lane] = w;
    s = __reduce_add_sync(0xffffffffu, s);
    if (lane == 0) norm2[row] = s;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

__global__ void __launch_bounds__(kThreads) l2_dp4a_kernel(const uint32_t* __restrict__ q, const uint32_t* __restrict__ qn,
                                                           int nq, const uint32_t* __restrict__ t,
                                                           const uint32_t* __restrict__ tn, int nt, int tsplit,
                                                           unsigned long long* __restrict__ key) {
    const int q0 = blockIdx.x * kQTile;
    if (q0 >= nq) return;
    const int tiles_total = (nt + kTTile - 1) / kTTile;
    const int tiles_per = (tiles_total + tsplit - 1) / tsplit;
    const int tile_lo = blockIdx.y * tiles_per, tile_hi = min(tiles_total, tile_lo + tiles_per);
    if (tile_lo >= tile_hi) return;
    __shared__ __align__(16) uint4 stage[2][kTTile * kWords / 4];
    __shared__ uint32_t snorm[2][kTTile];

    uint32_t a[kQPT][kWords];
    uint32_t an[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int row = min(q0 + k * kThreads + (int)threadIdx.x, nq - 1);
        const uint4* src = reinterpret_cast<const uint4*>(q + (size_t)row * kWords);
#pragma unroll
        for (int w = 0; w < kWords / 4; ++w) {
            const uint4 v = __ldg(src + w);
            a[k][4 * w] = v.x; a[k][4 * w + 1] = v.y; a[k][4 * w + 2] = v.z; a[k][4 * w + 3] = v.w;
        }
        an[k] = __ldg(qn + row);
    }
    unsigned long long best[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) best[k] = ~0ull;

    auto fill = [&](int buf, int tile) {
        const int base_row = tile * kTTile;
        const uint4* src = reinterpret_cast<const uint4*>(t + (size_t)base_row * kWords);
        const int limit = (nt - base_row) * (kWords / 4);      // uint4 units available
#pragma unroll
        for (int i = 0; i < (kTTile * kWords / 4) / kThreads; ++i) {
            const int e = i * kThreads + threadIdx.x;
            if (e < limit) cp_async16(&stage[buf][e], src + e);
        }
        if (threadIdx.x < kTTile && base_row + (int)threadIdx.x < nt) snorm[buf][threadIdx.x] = __ldg(tn + base_row + threadIdx.x);
        asm volatile("cp.async.commit_group;\n" ::);
    };

    fill(0, tile_lo);
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int buf = (tile - tile_lo) & 1;
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        if (tile + 1 < tile_hi) fill(buf ^ 1, tile + 1);
        const int jbase = tile * kTTile;
        const int jn = min(kTTile, nt - jbase);
        for (int j = 0; j < jn; ++j) {
            const uint4* row = &stage[buf][j * (kWords / 4)];
            int dot[kQPT];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) dot[k] = 0;
#pragma unroll
            for (int w = 0; w < kWords / 4; ++w) {
                const uint4 b = row[w];                              // warp-wide broadcast
#pragma unroll
                for (int k = 0; k < kQPT; ++k) {
                    dot[k] = __dp4a(a[k][4 * w], b.x, (unsigned)dot[k]);
                    dot[k] = __dp4a(a[k][4 * w + 1], b.y, (unsigned)dot[k]);
                    dot[k] = __dp4a(a[k][4 * w + 2], b.z, (unsigned)dot[k]);
                    dot[k] = __dp4a(a[k][4 * w + 3], b.w, (unsigned)dot[k]);
                }
            }
            const uint32_t bn = snorm[buf][j];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) {
                uint32_t d2 = an[k] + bn - 2u * (uint32_t)dot[k];
                // OpenCV compares the FLOAT distances sqrtf(d2) with strict '<'.  sqrtf is injective on integers
                // below 2^22 (spacing 1/(2*sqrt(n)) exceeds an ulp), so there the exact d2 orders identically; above,
                // neighbouring integers can share a float, and the lower index must win among them: order by the
                // float's bit pattern instead (monotone; consecutive floats from 2048.0f = 0x45000000 on).
                if (d2 >= (1u << 22)) d2 = (1u << 22) + (__float_as_uint(__fsqrt_rn((float)d2)) - 0x45000000u);
                const unsigned long long kk = ((unsigned long long)d2 << kTrainIdxBits) | (unsigned)(jbase + j);
                best[k] = kk < best[k] ? kk : best[k];
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        const int row = q0 + k * kThreads + threadIdx.x;
        if (row < nq) atomicMin(&key[row], best[k]);
    }
}

__global__ void l2_decode_kernel(const unsigned long long* __restrict__ key, int n, int32_t* __restrict__ train_idx,
                                 float* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = key[i];
    train_idx[i] = (int32_t)(k & kTrainIdxMask);
    const uint32_t g = (uint32_t)(k >> kTrainIdxBits);
    // g < 2^22: the exact squared distance -> OpenCV's std::sqrt of the exact float sum; else: the float's bit rank
    dist[i] = g < (1u << 22) ? __fsqrt_rn((float)g) : __uint_as_float(g - (1u << 22) + 0x45000000u);
}

}  // namespace

// d_q/d_t: device float descriptors [n][128]; scratch: u8 rows + norms + keys (sized by l2_scratch_bytes).
size_t l2_scratch_bytes(int nq, int nt) {
    return (size_t)(nq + nt) * (kDim + 4) + (size_t)nq * 8 + 256 + 64;
}

int launch_l2_dp4a(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, int32_t* d_train_idx,
                   float* d_dist, int* d_bad, int sm_count, cudaStream_t st) {
    uint8_t* p = static_cast<uint8_t*>(d_scratch);
    unsigned long long* key = reinterpret_cast<unsigned long long*>(p); p += (size_t)nq * 8;
    uint32_t* qn = reinterpret_cast<uint32_t*>(p); p += (size_t)nq * 4;
    uint32_t* tn = reinterpret_cast<uint32_t*>(p); p += (size_t)nt * 4;
    p = reinterpret_cast<uint8_t*>(((uintptr_t)p + 15) & ~(uintptr_t)15);
    uint32_t* q8 = reinterpret_cast<uint32_t*>(p); p += (size_t)nq * kDim;
    uint32_t* t8 = reinterpret_cast<uint32_t*>(p);
    cudaMemsetAsync(key, 0xFF, (size_t)nq * 8, st);
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    l2_narrow_kernel<<<(nq + 3) / 4, 128, 0, st>>>(d_q, nq, q8, qn, d_bad);
    l2_narrow_kernel<<<(nt + 3) / 4, 128, 0, st>>>(d_t, nt, t8, tn, d_bad);
    const int qtiles = (nq + kQTile - 1) / kQTile;
    const int tiles = (nt + kTTile - 1) / kTTile;
    int tsplit = (2 * sm_count + qtiles - 1) / qtiles;
    if (tsplit > tiles) tsplit = tiles;
    if (tsplit < 1) tsplit = 1;
    l2_dp4a_kernel<<<dim3(qtiles, tsplit), kThreads, 0, st>>>(q8, qn, nq, t8, tn, nt, tsplit, key);
    l2_decode_kernel<<<(nq + 255) / 256, 256, 0, st>>>(key, nq, d_train_idx, d_dist);
    return 4;
}

}  // namespace sfmgms
