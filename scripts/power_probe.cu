// power_probe.cu — does the tensor pipe's sustained mxf4 rate depend on the operand DATA?  (B200 runs the Hamming kernel
// at its power cap.)  Same instruction stream as hamming_fp4_kernel's MMA loop (M128 N240 K64, 4 K-steps per accumulator,
// 4 query tiles x 3 train tiles rotating, no epilogue, no TMA); the shared-memory operands are filled per mode:
//   0  constant +1/-1 pattern (what unit_peaks.cu measures)      1  random +-1 x random +-1 (the production encoding)
//   2  random +-1 x random {0,+1}                                 3  random {0,+1} x random {0,+1}
// Prints TOP/s of a short launch (cold, boost clock) and of a ~1.5 s launch (power-limited clock) per mode.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../sfm_gms_b200/csrc/tc_ptx.cuh"

using namespace sfmgms::tcptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

template <int kAcc>
__device__ __forceinline__ void mma_mxf4(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t sf) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], da, db, %4, [%5], [%5], p;\n\t}\n"
        ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(sf), "n"(kAcc) : "memory");
}

constexpr int A_TILE = 128 * 128, B_TILE = 240 * 128, NA = 4, NB = 3;

__device__ __forceinline__ uint32_t nibbles(uint32_t& s, int zero_one) {
    s = s * 1664525u + 1013904223u;
    const uint32_t bits = (s >> 8) & 0xFFu;   // 8 random bits -> 8 nibbles
    uint32_t w = 0;
    for (int k = 0; k < 8; ++k) {
        const uint32_t b = (bits >> k) & 1u;
        const uint32_t nib = zero_one ? (b ? 0x2u : 0x0u) : (b ? 0x2u : 0xAu);
        w |= nib << (4 * k);
    }
    return w;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int iters, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + NA * A_TILE, bar = b_smem + NB * B_TILE, tptr = bar + 16;
    volatile uint32_t* tptr_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
    uint32_t* data = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    uint32_t seed = 12345u + 7919u * threadIdx.x + 104729u * blockIdx.x;
    for (int i = threadIdx.x; i < (NA * A_TILE + NB * B_TILE) / 4; i += blockDim.x) {
        const bool is_b = i >= NA * A_TILE / 4;
        data[i] = mode == 0 ? 0x2A2A2A2Au : nibbles(seed, mode == 3 || (mode == 2 && is_b));
    }
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    {
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
        constexpr uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(240 >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sf = tmem + 480;
        const uint32_t a_lo0 = sdesc_lo(a_smem), b_lo0 = sdesc_lo(b_smem);
        int bt = 0;
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem + (it & 1) * 240;
            const uint32_t a_lo = a_lo0 + (it & 3) * (A_TILE >> 4);
            const uint32_t b_lo = b_lo0 + bt * (B_TILE >> 4);
            mma_mxf4<0>(d, a_lo + 0, b_lo + 0, idesc, sf);
            mma_mxf4<1>(d, a_lo + 2, b_lo + 2, idesc, sf);
            mma_mxf4<1>(d, a_lo + 4, b_lo + 4, idesc, sf);
            mma_mxf4<1>(d, a_lo + 6, b_lo + 6, idesc, sf);
            if ((it & 3) == 3 && ++bt == NB) bt = 0;
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

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int smem = NA * A_TILE + NB * B_TILE + 1024 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));   // 1 CTA per SM
    (void)smem;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int it_short = 20000, it_long = argc > 1 ? atoi(argv[1]) : 5000000;
    const char* names[4] = {"constant +-1 pattern", "random +-1 x random +-1", "random +-1 x random {0,1}", "random {0,1} x random {0,1}"};
    printf("# %s, %d SMs; tcgen05.mma kind::mxf4 M128 N240 K64 issue loop, 4 A x 3 B tiles rotating\n", prop.name, sms);
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 4; ++mode) {
            float ms_s = 0, ms_l = 0;
            probe_kernel<<<sms, 128, 200 * 1024>>>(2000, mode);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0)); probe_kernel<<<sms, 128, 200 * 1024>>>(it_short, mode); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_s, e0, e1));
            CK(cudaEventRecord(e0)); probe_kernel<<<sms, 128, 200 * 1024>>>(it_long, mode); CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_l, e0, e1));
            const double ops = (double)sms * 4 * (2.0 * 128 * 240 * 64);
            printf("mode %d (%-28s) rep %d: short %.1f TOP/s (%.2f ms) | sustained %.1f TOP/s (%.0f ms)\n", mode, names[mode], rep,
                   ops * it_short / (ms_s * 1e-3) / 1e12, ms_s, ops * it_long / (ms_l * 1e-3) / 1e12, ms_l);
            fflush(stdout);
        }
    return 0;
}
