// epi_probe.cu — what bounds hamming_fp4_kernel's epilogue next to its MMA stream?  Unit rates on one B200, every SM busy:
//   mma N      tcgen05.mma kind::mxf4 M128 N{240,160,128,120} K64 issue loop (SS mode): does a smaller N (more TMEM
//              accumulator slots) keep the tensor rate, or does the shared-memory operand read rate cap it?
//   ldtm W     W warps per SM (W/4 per TMEM lane quadrant) reading 80 fp32 columns per lane per iteration
//              (tcgen05.ld 32x32b x64 + x16, the kernel's epilogue load): TMEM read bytes per clock per SM.
//   mma+ldtm, mma+max, mma+ldtm+max: the MMA loop with 12 reader / ALU warps beside it (no barriers between them): the
//              tensor rate each mix leaves, and the reader / ALU rate.
// Built and run by scripts/epi_probe.py; output -> profiles/r2_epi_probe.txt.
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

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// modes (bit mask): 1 = MMA loop (warp 15), 2 = TMEM reads (warps 0..nw-1), 4 = max tree (warps 0..nw-1)
template <int N, int mode>
__global__ void __launch_bounds__(512, 1) probe_kernel(int nw, int mma_iters, int epi_iters, float* sink, long long* clk_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + 65536, bar = base + 65536 + 3 * 30720, tptr = bar + 16;   // bar+32, bar+40: dummy barriers
    volatile uint32_t* tptr_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (tptr - smem_u32(smem_raw)));
    uint32_t* data = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < (65536 + 3 * 30720) / 4; i += blockDim.x) data[i] = 0x2A2A2A2Au;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 32, 1);
        mbar_init(bar + 40, 1);
        for (int k = 0; k < 4; ++k) { mbar_init(bar + 64 + 8 * k, 1); mbar_init(bar + 96 + 8 * k, nw > 0 ? nw : 1); }   // tfull[4], tempty[4]
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 14) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr_g;
    if (warp < 4) {
        uint32_t ones[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) ones[k] = 0x7F7F7F7Fu;
        const uint32_t t = tmem + ((uint32_t)(warp * 32) << 16) + 480;
        tc_st16(t, ones);
        tc_st16(t + 16, ones);
        // defined accumulator contents for the readers
#pragma unroll 1
        for (int c = 0; c < 480; c += 16) tc_st16(tmem + ((uint32_t)(warp * 32) << 16) + c, ones);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const long long t0 = clock64();
    if (warp == 15) {
        if (mode & 1) {
            constexpr uint32_t idesc = (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(128 >> 4) << 24);
            constexpr int kSlots = N >= 120 ? 480 / N : 4;
            const uint32_t a_lo0 = sdesc_lo(a_smem), b_lo0 = sdesc_lo(b_smem);
            const uint32_t sf = tmem + 480;
            int slot = 0;
            uint32_t ph = 0;          // bit k: phase of slot k
            for (int it = 0; it < mma_iters; ++it) {
                const uint32_t d = tmem + slot * N;
                // mode 64: operands rotate like the kernel's (4 query sub-tiles per train tile, 3 train-tile stages)
                const uint32_t a_lo = (mode & 64) ? a_lo0 + (uint32_t)(it & 3) * (16384 >> 4) : a_lo0;
                const uint32_t b_lo = (mode & 64) ? b_lo0 + (uint32_t)((it >> 2) % 3) * (30720 >> 4) : b_lo0;
                if (mode & 32) {
                    mbar_wait(bar + 96 + 8 * slot, ((ph >> slot) & 1) ^ 1);
                    ph ^= 1u << slot;
                    tc_fence_after();
                }
                if (elect_one()) {
                    mma_mxf4<0>(d, a_lo + 0, b_lo + 0, idesc, sf);
                    mma_mxf4<1>(d, a_lo + 2, b_lo + 2, idesc, sf);
                    mma_mxf4<1>(d, a_lo + 4, b_lo + 4, idesc, sf);
                    mma_mxf4<1>(d, a_lo + 6, b_lo + 6, idesc, sf);
                    if (mode & 8) tc_commit(bar + 32);          // a barrier nobody waits on: cost of one commit per accumulator
                    if (mode & 16) tc_commit(bar + 40);         // ... and a second one
                    if (mode & 32) tc_commit(bar + 64 + 8 * slot);
                }
                if (++slot == kSlots) slot = 0;
            }
            if (elect_one()) tc_commit(bar);
            mbar_wait(bar, 0);
            if (threadIdx.x == 15 * 32) clk_out[blockIdx.x * 2 + 0] = clock64() - t0;
        }
    } else if (warp < nw && (mode & (6 | 32))) {
        const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 80;
        float acc = 0.f;
        float v[80];
#pragma unroll
        for (int j = 0; j < 80; ++j) v[j] = (float)(threadIdx.x * 80 + j);
        constexpr int kSlotsE = N >= 120 ? 480 / N : 4;
        int eslot = 0;
        uint32_t eph = 0;
        for (int it = 0; it < epi_iters; ++it) {
            if (mode & 32) {
                mbar_wait(bar + 64 + 8 * eslot, (eph >> eslot) & 1);
                eph ^= 1u << eslot;
                tc_fence_after();
            }
            if (mode & 2) {
                int r[80];
                const uint32_t ta = (mode & 32) ? tmem + ((uint32_t)((warp & 3) * 32) << 16) + eslot * N + (warp >> 2) * (N / 3) : tb + (it & 1) * 240;
                tc_ld64(ta, reinterpret_cast<int(&)[64]>(r[0]));
                tc_ld16(ta + 64, reinterpret_cast<int(&)[16]>(r[64]));
                tc_wait_ld();
                if (mode & 4) {
#pragma unroll
                    for (int j = 0; j < 80; ++j) v[j] = __int_as_float(r[j]);
                } else {
                    acc += __int_as_float(r[0]) + __int_as_float(r[79]);
                }
            }
            if (mode & 32) {
                tc_fence_before();
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(bar + 96 + 8 * eslot);
                if (++eslot == kSlotsE) eslot = 0;
            }
            if (mode & 4) {
                float gm[10];
#pragma unroll
                for (int g = 0; g < 10; ++g) {
                    if (!(mode & 2)) {          // keep the tree alive without loads: the values are opaque every iteration
                        asm volatile("" : "+f"(v[8 * g]), "+f"(v[8 * g + 1]), "+f"(v[8 * g + 2]), "+f"(v[8 * g + 3]), "+f"(v[8 * g + 4]),
                                     "+f"(v[8 * g + 5]), "+f"(v[8 * g + 6]), "+f"(v[8 * g + 7]));
                    }
                    gm[g] = fmaxf(max3(v[8 * g + 6], v[8 * g + 7], max3(v[8 * g], v[8 * g + 1], v[8 * g + 2])),
                                  max3(v[8 * g + 3], v[8 * g + 4], v[8 * g + 5]));
                }
                float m = gm[0];
#pragma unroll
                for (int g = 1; g < 10; ++g) m = fmaxf(m, gm[g] + (float)g * 0.0625f);
                acc = fmaxf(acc, m);
            }
        }
        if (acc == 123456.f) sink[threadIdx.x] = acc;
        if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&clk_out[blockIdx.x * 2 + 1], (unsigned long long)(clock64() - t0));   // slowest warp
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 14) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

template <int N, int mode>
static void run(const char* name, int nw, int mma_iters, int epi_iters, int sms, float* sink, long long* d_clk) {
    CK(cudaFuncSetAttribute(probe_kernel<N, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long* h = (long long*)malloc(sizeof(long long) * 2 * sms);
    double best_mma = 1e30, best_epi = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaMemset(d_clk, 0, sizeof(long long) * 2 * sms));
        probe_kernel<N, mode><<<sms, 512, 200 * 1024>>>(nw, mma_iters, epi_iters, sink, d_clk);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_clk, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
        double a = 0, b = 0;
        for (int i = 0; i < sms; ++i) { a += (double)h[2 * i]; b += (double)h[2 * i + 1]; }
        a /= sms; b /= sms;
        if (rep && a < best_mma) best_mma = a;
        if (rep && b < best_epi) best_epi = b;
    }
    printf("%-34s", name);
    if (mode & 1) printf(" mma: %7.1f clk per N=%d accumulator (4 MMAs; ideal %5.1f)", best_mma / mma_iters, N, 128.0 * N / 240 * 4);
    if (mode & (6 | 32)) {
        printf(" | epi warps=%2d: %7.1f clk per 80-column pass per warp", nw, best_epi / epi_iters);
        if (mode & 2) printf(", TMEM read %6.1f B/clk/SM", (double)nw * 32 * 80 * 4 * epi_iters / best_epi);
        if (mode & 4) printf(", %5.2f max-tree values/clk/SMSP", (double)nw * 32 * 80 * epi_iters / best_epi / 4);
    }
    printf("\n");
    free(h);
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    float* sink;
    long long* d_clk;
    CK(cudaMalloc(&sink, 4096 * 4));
    CK(cudaMalloc(&d_clk, sizeof(long long) * 2 * sms));
    printf("# %s, %d SMs; clocks from clock64() inside the kernel, mean over SMs, best of 3\n", prop.name, sms);
    const int MI = 4000, EI = 4000;
    run<240, 1>("mma N=240", 0, MI, 0, sms, sink, d_clk);
    run<160, 1>("mma N=160", 0, MI, 0, sms, sink, d_clk);
    run<128, 1>("mma N=128", 0, MI, 0, sms, sink, d_clk);
    run<120, 1>("mma N=120", 0, MI, 0, sms, sink, d_clk);
    run<240, 9>("mma N=240, commit per accumulator", 0, MI, 0, sms, sink, d_clk);
    run<240, 25>("mma N=240, 2 commits per accumulator", 0, MI, 0, sms, sink, d_clk);
    run<120, 9>("mma N=120, commit per accumulator", 0, MI, 0, sms, sink, d_clk);
    run<240, 65>("mma N=240, rotating operands", 0, MI, 0, sms, sink, d_clk);
    run<240, 97>("handshake N=240x2, null epilogue, rotating operands", 12, MI, MI, sms, sink, d_clk);
    run<240, 103>("handshake N=240x2, ldtm + max tree, rotating operands", 12, MI, MI, sms, sink, d_clk);
    run<240, 33>("handshake N=240x2, null epilogue", 12, MI, MI, sms, sink, d_clk);
    run<240, 35>("handshake N=240x2, ldtm", 12, MI, MI, sms, sink, d_clk);
    run<240, 39>("handshake N=240x2, ldtm + max tree", 12, MI, MI, sms, sink, d_clk);
    run<160, 33>("handshake N=160x3, null epilogue", 12, MI, MI, sms, sink, d_clk);
    run<120, 33>("handshake N=120x4, null epilogue", 12, MI, MI, sms, sink, d_clk);
    run<64, 1>("mma N=64 (issue cost)", 0, MI, 0, sms, sink, d_clk);
    run<32, 1>("mma N=32 (issue cost)", 0, MI, 0, sms, sink, d_clk);
    run<16, 1>("mma N=16 (issue cost)", 0, MI, 0, sms, sink, d_clk);
    run<8, 1>("mma N=8 (issue cost)", 0, MI, 0, sms, sink, d_clk);
    for (int nw : {4, 8, 12}) run<240, 2>("ldtm", nw, 0, EI, sms, sink, d_clk);
    for (int nw : {4, 8, 12}) run<240, 6>("ldtm + max tree", nw, 0, EI, sms, sink, d_clk);
    run<240, 3>("mma N=240 + ldtm", 12, MI, EI * 2, sms, sink, d_clk);
    run<240, 7>("mma N=240 + ldtm + max tree", 12, MI, EI, sms, sink, d_clk);
    run<160, 7>("mma N=160 + ldtm + max tree", 12, MI * 3 / 2, EI, sms, sink, d_clk);
    return 0;
}
