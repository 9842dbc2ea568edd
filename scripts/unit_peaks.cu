// unit_peaks.cu — measured ceilings of the units the hot-path kernels are bound by ("must be
// microbenchmarked on the box").  Built by scripts/unit_peaks.py (nvcc, sm_100a) and run on the B200; its JSON line
// becomes profiles/r2_unit_peaks.json, which bench.py uses as roofline denominators.
//
//   mxf4_tops         tcgen05.mma.cta_group::1.kind::mxf4.block_scale, M128 N240 K64, both operands from shared memory,
//                     two TMEM accumulators alternating, 4 K-steps per accumulator (exactly the instruction stream of
//                     hamming_fp4_kernel) with NO epilogue and NO TMA: one elected thread issues, one commit per
//                     accumulator.  2*128*240*64 OPs per instruction, all 148 SMs.
//   i8_tops           the same with kind::i8, M128 N128 K32 (SS mode).
//   smem_atomic_gops  shared-memory atomicAdd on random 32-bit words of a 100 KB array, 512 threads per CTA, 2 CTAs per SM.
//   smem_vote_gvotes  the vote pattern of gms_vote2_kernel: atomicAdd (returning) on a random histogram word + atomicMax on
//                     one of 140 row slots: two dependent atomics per vote.
//   popc_gpopc_s      POPC.b32 issue loop (8 independent chains per thread), all SMs.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../sfm_gms_b200/csrc/tc_ptx.cuh"

using namespace sfmgms::tcptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

template <int kAcc>
__device__ __forceinline__ void mma_mxf4(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t sf) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], da, db, %4, [%5], [%5], p;\n\t}\n"
        ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(sf), "n"(kAcc) : "memory");
}
template <int kAcc>
__device__ __forceinline__ void mma_i8(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, p;\n\t}\n"
        ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "n"(kAcc) : "memory");
}

// kind 0: mxf4 M128 N240 K64;  kind 1: i8 M128 N128 K32
template <int kKind>
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + 16384, bar = base + 16384 + 32768, tptr = bar + 16;
    volatile uint32_t* tptr_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
    uint32_t* data = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) data[i] = kKind == 0 ? 0x2A2A2A2Au : 0x01FF01FFu;   // +-1 values
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr_g;
    if (kKind == 0) {   // unit block scales in TMEM columns [480, 512)
        uint32_t ones[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) ones[k] = 0x7F7F7F7Fu;
        const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + 480;
        tc_st16(t, ones);
        tc_st16(t + 16, ones);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t a_lo = sdesc_lo(a_smem), b_lo = sdesc_lo(b_smem);
        if (kKind == 0) {
            constexpr uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(240 >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t sf = tmem + 480;
            for (int it = 0; it < iters; ++it) {
                const uint32_t d = tmem + (it & 1) * 240;
                mma_mxf4<0>(d, a_lo + 0, b_lo + 0, idesc, sf);
                mma_mxf4<1>(d, a_lo + 2, b_lo + 2, idesc, sf);
                mma_mxf4<1>(d, a_lo + 4, b_lo + 4, idesc, sf);
                mma_mxf4<1>(d, a_lo + 6, b_lo + 6, idesc, sf);
            }
        } else {
            constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int it = 0; it < iters; ++it) {
                const uint32_t d = tmem + (it & 1) * 128;
                mma_i8<0>(d, a_lo + 0, b_lo + 0, idesc);
                mma_i8<1>(d, a_lo + 2, b_lo + 2, idesc);
                mma_i8<1>(d, a_lo + 4, b_lo + 4, idesc);
                mma_i8<1>(d, a_lo + 6, b_lo + 6, idesc);
            }
        }
        tc_commit(bar);
        mbar_wait(bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

__device__ __forceinline__ uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int kMode>   // 0: plain atomicAdd on random words; 1: the vote pattern (returning add + atomicMax on a row slot)
__global__ void __launch_bounds__(512) smem_atomic_kernel(int iters, unsigned* sink) {
    extern __shared__ uint32_t sm[];
    constexpr int kWords = 25600;          // 100 KB
    uint32_t* best = sm + kWords;          // 140 row slots
    for (int i = threadIdx.x; i < kWords + 160; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t r = lcg(s);
            const uint32_t w = r % kWords;
            if (kMode == 0) {
                atomicAdd(&sm[w], 1u);
            } else {
                const uint32_t sh = (r >> 20 & 1u) * 16;
                const uint32_t old = atomicAdd(&sm[w], 1u << sh);
                const uint32_t c = ((old >> sh) & 0xFFFFu) + 1u;
                atomicMax(&best[w % 140], (c << 11) | (r & 1023u));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = sm[7] + best[3];
}

__global__ void __launch_bounds__(256) popc_kernel(int iters, unsigned* sink) {
    uint32_t v[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = threadIdx.x * 2654435761u + k * 97u + blockIdx.x; acc[k] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += __popc(v[k] ^ acc[k]); }
    }
    uint32_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += acc[k];
    if (t == 0x12345678u) sink[0] = t;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned* sink;
    CK(cudaMalloc(&sink, 4096 * 4));
    const int smem_mma = 16384 + 32768 + 1024 + 64;
    CK(cudaFuncSetAttribute(mma_peak_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // 1 CTA per SM
    CK(cudaFuncSetAttribute(mma_peak_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int it_mma = 20000;
    const double ms_fp4 = time_ms([&] { mma_peak_kernel<0><<<sms, 128, 200 * 1024>>>(it_mma); }, 5);
    const double ms_i8 = time_ms([&] { mma_peak_kernel<1><<<sms, 128, 200 * 1024>>>(it_mma); }, 5);
    (void)smem_mma;
    const double mxf4_tops = (double)sms * it_mma * 4 * (2.0 * 128 * 240 * 64) / (ms_fp4 * 1e-3) / 1e12;
    const double i8_tops = (double)sms * it_mma * 4 * (2.0 * 128 * 128 * 32) / (ms_i8 * 1e-3) / 1e12;
    // sustained variant: 10x longer (power / clock droop shows here)
    const double ms_fp4_long = time_ms([&] { mma_peak_kernel<0><<<sms, 128, 200 * 1024>>>(it_mma * 10); }, 2);
    const double mxf4_tops_long = (double)sms * it_mma * 10 * 4 * (2.0 * 128 * 240 * 64) / (ms_fp4_long * 1e-3) / 1e12;

    const int smem_at = (25600 + 160) * 4;
    CK(cudaFuncSetAttribute(smem_atomic_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_at));
    CK(cudaFuncSetAttribute(smem_atomic_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_at));
    const int it_at = 2000, grid_at = sms * 2;
    const double ms_a0 = time_ms([&] { smem_atomic_kernel<0><<<grid_at, 512, smem_at>>>(it_at, sink); }, 5);
    const double ms_a1 = time_ms([&] { smem_atomic_kernel<1><<<grid_at, 512, smem_at>>>(it_at, sink); }, 5);
    const double atom_gops = (double)grid_at * 512 * it_at * 8 / (ms_a0 * 1e-3) / 1e9;
    const double vote_gops = (double)grid_at * 512 * it_at * 8 / (ms_a1 * 1e-3) / 1e9;

    const int it_p = 20000, grid_p = sms * 8;
    const double ms_p = time_ms([&] { popc_kernel<<<grid_p, 256>>>(it_p, sink); }, 5);
    const double popc_g = (double)grid_p * 256 * it_p * 8 / (ms_p * 1e-3) / 1e9;

    printf("{\"gpu\": \"%s\", \"sms\": %d, \"mxf4_tops\": %.1f, \"mxf4_tops_sustained_10x\": %.1f, \"i8_tops\": %.1f, "
           "\"smem_atomic_gops\": %.1f, \"smem_vote_gvotes\": %.1f, \"popc_gpopc_s\": %.1f, "
           "\"ms\": {\"mxf4\": %.3f, \"mxf4_long\": %.3f, \"i8\": %.3f, \"atomic\": %.3f, \"vote\": %.3f, \"popc\": %.3f}, "
           "\"how\": \"scripts/unit_peaks.cu: best of 5 launches, CUDA events; mxf4 = tcgen05.mma kind::mxf4.block_scale M128 N240 K64 "
           "SS-mode issue loop (4 K-steps per accumulator, 2 TMEM accumulators, no epilogue, no TMA) on every SM; i8 = kind::i8 M128 N128 "
           "K32; smem atomics = random 32-bit words of a 100 KB array, 2 CTAs x 512 threads per SM; vote = returning atomicAdd + "
           "atomicMax (the gms_vote2 pattern); popc = 8 independent POPC chains per thread\"}\n",
           prop.name, sms, mxf4_tops, mxf4_tops_long, i8_tops, atom_gops, vote_gops, popc_g, ms_fp4, ms_fp4_long, ms_i8, ms_a0, ms_a1, ms_p);
    return 0;
}
