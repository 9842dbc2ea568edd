// l2_f32.cu — brute-force L2 nearest neighbour for GENERAL float descriptors, SURVEY §8f-3.
//
// cv::BFMatcher(NORM_L2)::match (FeatureMatchUtil.cpp:22-23, 40-41, 66-68) on CV_32F rows evaluates, per
// (query, train) pair, sqrt(hal::normL2Sqr_(a, b, n)) and keeps the lowest train index among equal FLOAT
// distances.  For integer-valued data (OpenCV SIFT) any summation order is exact and the tensor-core kernel
// (l2_tc.cu) is used; for everything else the float result depends on the ORDER of the additions, so this kernel
// restates that order operation by operation (OpenCV core, norm.cpp `normL2Sqr_`, universal-intrinsics build with
// 4 float lanes = the SSE baseline of the stock x86-64 packages):
//     four vector accumulators d0..d3 of 4 lanes; per 16 elements  t = a - b;  d_k += t * t   (mul and add rounded
//     separately, no FMA);  s = ((d0 + d1) + d2) + d3 lane-wise;  d = (s[0] + s[2]) + (s[1] + s[3]);
//     then the n % 16 tail elements one by one: d += t * t;   distance = sqrtf(d).
// Pinned bit-exact against cv2 4.13 in this image (tests/golden/l2_float.npz and the live-cv2 CPU test of the
// oracle); an OpenCV build with wider baseline vectors (AVX2/AVX-512 `normL2Sqr_`) may differ in the last ulp.
// CUDA cores; subtract and accumulate are packed fp32x2 (sub/add.rn.f32x2: two train columns per instruction, each
// lane rounded exactly like the scalar op), the square is a scalar FMUL so that it cannot be contracted into an FMA.  64 x 64 output tile per CTA, 4 x 4 per thread, operands transposed in
// shared memory.  Result per query: atomicMin of (float bits of distance << 32 | trainIdx).
#include "common.cuh"

namespace sfmgms {

namespace {

constexpr int TM = 64, TN = 64;          // queries x train rows per CTA
constexpr int kPad = TM + 4;             // row pitch of the transposed tiles (floats): keeps float4 loads aligned
constexpr int kThreads = 256;            // 16 x 16 threads, 4 x 4 outputs each
static_assert(TM == TN, "one pitch for both tiles");

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// t*t per lane with SCALAR multiplies: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under
// --fmad=false (seen in SASS), which would round once instead of twice; scalar FMUL + packed FADD2 stays unfused.
__device__ __forceinline__ u64 sqr2(u64 a) {
    float lo, hi;
    unpack2(a, lo, hi);
    return pack2(__fmul_rn(lo, lo), __fmul_rn(hi, hi));
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct Acc {           // 4 query rows x 2 packed column pairs
    u64 v[4][2];
};

// one SIMD lane c of OpenCV's loop: sum over k of (a[c+16k] - b[c+16k])^2, accumulated in k order from 0
__device__ __forceinline__ void lane_sum(const float* __restrict__ As, const float* __restrict__ Bs, int c, int nblk,
                                         int ty, int tx, Acc& acc) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc.v[i][0] = acc.v[i][1] = 0ull;       // (+0.f, +0.f)
    for (int k = 0; k < nblk; ++k) {
        const int e = c + 16 * k;
        const float4 a = *reinterpret_cast<const float4*>(As + e * kPad + 4 * ty);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(Bs + e * kPad + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 aa = pack2(av[i], av[i]);
            const u64 t0 = sub2(aa, b.x), t1 = sub2(aa, b.y);
            acc.v[i][0] = add2(acc.v[i][0], sqr2(t0));
            acc.v[i][1] = add2(acc.v[i][1], sqr2(t1));
        }
    }
}

// s_l = ((d0[l] + d1[l]) + d2[l]) + d3[l]: element lanes l, 4+l, 8+l, 12+l of each 16-element block
__device__ __forceinline__ void vec_lane(const float* As, const float* Bs, int l, int nblk, int ty, int tx, Acc& s) {
    Acc t;
    lane_sum(As, Bs, l, nblk, ty, tx, s);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        lane_sum(As, Bs, l + 4 * k, nblk, ty, tx, t);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s.v[i][0] = add2(s.v[i][0], t.v[i][0]); s.v[i][1] = add2(s.v[i][1], t.v[i][1]); }
    }
}

__global__ void __launch_bounds__(kThreads) l2_f32_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t,
                                                          int nt, int dim, u64* __restrict__ key) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                         // [dim][kPad]: As[e * kPad + r] = q[(q0 + r) * dim + e]
    float* Bs = smem + (size_t)dim * kPad;
    const int q0 = blockIdx.y * TM, t0 = blockIdx.x * TN;
    for (int i = threadIdx.x; i < TM * dim; i += kThreads) {
        const int r = i / dim, e = i - r * dim;
        As[e * kPad + r] = (q0 + r < nq) ? __ldg(q + (size_t)(q0 + r) * dim + e) : 0.f;
        Bs[e * kPad + r] = (t0 + r < nt) ? __ldg(t + (size_t)(t0 + r) * dim + e) : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int nblk = dim / 16;

    // d = (s0 + s2) + (s1 + s3)
    Acc d, u;
    vec_lane(As, Bs, 0, nblk, ty, tx, d);
    vec_lane(As, Bs, 2, nblk, ty, tx, u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { d.v[i][0] = add2(d.v[i][0], u.v[i][0]); d.v[i][1] = add2(d.v[i][1], u.v[i][1]); }
    Acc w;
    vec_lane(As, Bs, 1, nblk, ty, tx, w);
    vec_lane(As, Bs, 3, nblk, ty, tx, u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        d.v[i][0] = add2(d.v[i][0], add2(w.v[i][0], u.v[i][0]));
        d.v[i][1] = add2(d.v[i][1], add2(w.v[i][1], u.v[i][1]));
    }
    // scalar tail, element by element
    for (int e = nblk * 16; e < dim; ++e) {
        const float4 a = *reinterpret_cast<const float4*>(As + e * kPad + 4 * ty);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(Bs + e * kPad + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u64 aa = pack2(av[i], av[i]);
            const u64 x0 = sub2(aa, b.x), x1 = sub2(aa, b.y);
            d.v[i][0] = add2(d.v[i][0], sqr2(x0));
            d.v[i][1] = add2(d.v[i][1], sqr2(x1));
        }
    }
    // per query row: lowest (float distance, train index) over this tile's 64 columns
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float dd[4];
        unpack2(d.v[i][0], dd[0], dd[1]);
        unpack2(d.v[i][1], dd[2], dd[3]);
        u64 best = ~0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = t0 + 4 * tx + j;
            const u64 k = ((u64)__float_as_uint(__fsqrt_rn(dd[j])) << 32) | (uint32_t)col;
            if (col < nt && k < best) best = k;
        }
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const u64 other = __shfl_xor_sync(0xffffffffu, best, o, 16);
            best = other < best ? other : best;
        }
        const int row = q0 + 4 * ty + i;
        if (tx == 0 && row < nq && best != ~0ull) atomicMin(key + row, best);
    }
}

__global__ void l2_f32_decode_kernel(const u64* __restrict__ key, int nq, int32_t* __restrict__ idx, float* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const u64 k = key[i];
    idx[i] = (int32_t)(k & 0xffffffffu);
    dist[i] = __uint_as_float((uint32_t)(k >> 32));
}

}  // namespace

size_t l2_f32_scratch_bytes(int nq) { return (size_t)nq * 8 + 64; }

int l2_f32_max_dim() { return 256; }

int launch_l2_f32(const float* d_q, int nq, const float* d_t, int nt, int dim, void* d_scratch, int32_t* d_train_idx,
                  float* d_dist, cudaStream_t st) {
    if (dim < 1 || dim > l2_f32_max_dim()) return -1;
    u64* key = static_cast<u64*>(d_scratch);
    const size_t smem = 2 * (size_t)dim * kPad * sizeof(float);
    if (cudaFuncSetAttribute(l2_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    cudaMemsetAsync(key, 0xFF, (size_t)nq * 8, st);
    const dim3 grid((unsigned)((nt + TN - 1) / TN), (unsigned)((nq + TM - 1) / TM));
    l2_f32_kernel<<<grid, kThreads, smem, st>>>(d_q, nq, d_t, nt, dim, key);
    l2_f32_decode_kernel<<<(nq + 255) / 256, 256, 0, st>>>(key, nq, d_train_idx, d_dist);
    return 2;
}

}  // namespace sfmgms
