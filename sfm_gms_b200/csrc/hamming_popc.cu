// hamming_popc.cu — brute-force Hamming nearest neighbour on the CUDA cores (XOR + POPC).
//
// Replaces cv::BFMatcher(NORM_HAMMING)::match (reference call site FeatureMatchUtil.cpp:68).
// Result per query row: key = (distance << 18) | trainIdx, minimised over all train rows.  Because the
// train index sits in the low bits, an unsigned min IS OpenCV's strict-'<' scan: the lowest trainIdx wins
// among equal distances, independent of tiling / split order (atomicMin merges are order-free => deterministic).
//
// Layout: descriptors are N x 32 B row-major (CV_8U), read as two 128-bit words per row.
//   * each thread keeps QPT query descriptors in registers (8 x u32 each),
//   * a CTA streams the train descriptors through a double-buffered shared-memory tile filled with
//     cp.async (LDGSTS); all lanes of a warp read the same train row => shared-memory broadcast,
//   * inner loop per distance: 8 LOP3(xor) + 8 POPC + adds, then one IMAD (key) + one IMNMX.
// Roofline: POPC issues at 16 lanes/clk/SM => 2 distances/clk/SM -- the INT/popc pipe,
// not HBM, bounds this kernel (compulsory traffic is (N1+N2)*32 B).
#include "common.cuh"

namespace sfmgms {

namespace {

constexpr int kThreads = 128;
constexpr int kQPT = 4;                       // queries per thread
constexpr int kQTile = kThreads * kQPT;       // 512 queries per CTA
constexpr int kTTile = 256;                   // train rows per shared-memory stage (8 KB)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__global__ void __launch_bounds__(kThreads) hamming_popc_kernel(const PairDesc* __restrict__ pairs, int tsplit) {
    const PairDesc pd = pairs[blockIdx.z];
    const int q0 = blockIdx.x * kQTile;
    if (q0 >= pd.n1 || pd.n2 <= 0) return;

    // train range of this split, aligned to the tile size
    const int tiles_total = (pd.n2 + kTTile - 1) / kTTile;
    const int tiles_per = (tiles_total + tsplit - 1) / tsplit;
    const int tile_lo = blockIdx.y * tiles_per;
    const int tile_hi = min(tiles_total, tile_lo + tiles_per);
    if (tile_lo >= tile_hi) return;

    __shared__ __align__(16) uint4 stage[2][kTTile * 2];

    // queries -> registers (coalesced 2 x 128-bit per row; rows past n1 re-read the last row, never stored)
    uint32_t q[kQPT][kDescWords];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        int row = min(q0 + k * kThreads + (int)threadIdx.x, pd.n1 - 1);
        const uint4* src = reinterpret_cast<const uint4*>(pd.desc1) + (size_t)row * 2;
        uint4 a = __ldg(src), b = __ldg(src + 1);
        q[k][0] = a.x; q[k][1] = a.y; q[k][2] = a.z; q[k][3] = a.w;
        q[k][4] = b.x; q[k][5] = b.y; q[k][6] = b.z; q[k][7] = b.w;
    }
    uint32_t best[kQPT];
#pragma unroll
    for (int k = 0; k < kQPT; ++k) best[k] = kKeyInit;

    const uint4* tsrc = reinterpret_cast<const uint4*>(pd.desc2);
    auto fill = [&](int buf, int tile) {
        const int base = tile * kTTile * 2;                      // in uint4 units
        const int limit = pd.n2 * 2;
#pragma unroll
        for (int i = 0; i < (kTTile * 2) / kThreads; ++i) {
            int e = i * kThreads + threadIdx.x;
            int g = base + e;
            if (g < limit) cp_async16(&stage[buf][e], tsrc + g);
        }
        cp_async_commit();
    };

    fill(0, tile_lo);
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int buf = (tile - tile_lo) & 1;
        cp_async_wait_all();
        __syncthreads();                                  // stage[buf] landed; stage[buf^1] free (read 2 iters ago)
        if (tile + 1 < tile_hi) fill(buf ^ 1, tile + 1);
        const int jbase = tile * kTTile;
        const int jn = min(kTTile, pd.n2 - jbase);
#pragma unroll 2
        for (int j = 0; j < jn; ++j) {
            const uint4 t0 = stage[buf][2 * j], t1 = stage[buf][2 * j + 1];
#pragma unroll
            for (int k = 0; k < kQPT; ++k) {
                int d = __popc(q[k][0] ^ t0.x) + __popc(q[k][1] ^ t0.y) + __popc(q[k][2] ^ t0.z) +
                        __popc(q[k][3] ^ t0.w) + __popc(q[k][4] ^ t1.x) + __popc(q[k][5] ^ t1.y) +
                        __popc(q[k][6] ^ t1.z) + __popc(q[k][7] ^ t1.w);
                uint32_t key = ((uint32_t)d << kTrainIdxBits) + (uint32_t)(jbase + j);
                best[k] = min(best[k], key);
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int k = 0; k < kQPT; ++k) {
        int row = q0 + k * kThreads + threadIdx.x;
        if (row < pd.n1) {
            if (tsplit == 1) pd.key[row] = best[k];
            else atomicMin(&pd.key[row], best[k]);
        }
    }
}

}  // namespace

// Precondition: every pair's key[] is initialised to kKeyInit (0xFFFFFFFF) by the caller, so that split
// launches can merge with atomicMin and pairs with an empty train set keep "no match".
int launch_hamming_popc(const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                        cudaStream_t st) {
    if (n_pairs <= 0) return 0;
    int max_n1 = 0, max_n2 = 0;
    for (int p = 0; p < n_pairs; ++p) {
        if (h_pairs[p].n1 > max_n1) max_n1 = h_pairs[p].n1;
        if (h_pairs[p].n2 > max_n2) max_n2 = h_pairs[p].n2;
    }
    if (max_n1 == 0 || max_n2 == 0) return 0;
    const int qtiles = (max_n1 + kQTile - 1) / kQTile;
    // small batches: split the train range so that at least ~2 CTAs per SM exist
    const long long ctas = (long long)qtiles * n_pairs;
    const long long want = 2LL * sm_count;
    int tsplit = 1;
    if (ctas < want) {
        const int max_tiles = (max_n2 + kTTile - 1) / kTTile;
        long long t = (want + ctas - 1) / ctas;
        if (t > max_tiles) t = max_tiles;
        if (t < 1) t = 1;
        tsplit = (int)t;
    }
    dim3 grid(qtiles, tsplit, n_pairs);
    hamming_popc_kernel<<<grid, kThreads, 0, st>>>(d_pairs, tsplit);
    kmark("hamming_popc", st);
    return 1;
}

}  // namespace sfmgms
