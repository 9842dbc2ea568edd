// gms.cu — Grid-based Motion Statistics inlier voting (general path: dense histograms in global/L2).
//
// Replaces cv::xfeatures2d::matchGMS (reference call sites FeatureMatchUtil.cpp:69, DisparityUtil.cpp:149,299;
// algorithm = OpenCV-contrib 4.5.2 GMSMatcher, restated from SURVEY.md Appendix A).  Rows of SURVEY §8(a):
//   a4 normalizePoints + a5 getGridIndexLeft/Right + a7 assignMatchPairs  -> gms_assign_kernel
//   a8 verifyCellPairs (argmax, 3x3 support, f64 threshold)               -> gms_verify_kernel
//   a9 run (mark + count) / a10 getInlierMask (best hypothesis)          -> gms_count / gms_select / gms_mask
//
// Work sharing the reference does not do: the 4 left indices of a match do not depend on scale/rotation,
// the right index only on scale, rotation only on verify => 20 histograms serve all 40 hypotheses
// (the reference rebuilds 160).
//
// Floating point: the three decisions that hinge on FP are evaluated with explicitly rounded intrinsics
// (__fdiv_rn, __fmul_rn, __dadd_rn, __ddiv_rn, __dsqrt_rn, __dmul_rn): no FMA contraction, no fast math —
// bit-identical to the divss/mulss/addsd/divsd/sqrtsd/mulsd sequence of the reference binary.
// Integer histogram atomics are order-independent => results are deterministic.
//
// Roofline: HBM/L2 bandwidth.  Algorithmic bytes per pair (SURVEY §8d figure ii):
//   sum over used (scale,shift) of [4*400*G_r zero + 4*400*G_r scan + 12*N RMW] + 8*N*H mark + N.
#include <cstdlib>

#include "common.cuh"

namespace sfmgms {

namespace {

// ROT[8][9], 1-based (DLL .rdata 0x18012f520)
__constant__ int8_t c_rot[kNumRot][9] = {
    {1, 2, 3, 4, 5, 6, 7, 8, 9}, {4, 1, 2, 7, 5, 3, 8, 9, 6}, {7, 4, 1, 8, 5, 2, 9, 6, 3},
    {8, 7, 4, 9, 5, 1, 6, 3, 2}, {9, 8, 7, 6, 5, 4, 3, 2, 1}, {6, 9, 8, 3, 5, 7, 2, 1, 4},
    {3, 6, 9, 2, 5, 8, 1, 4, 7}, {2, 3, 6, 1, 5, 9, 4, 7, 8}};

// cvFloor: i = trunc(v); i - (i > v)   (SURVEY Appendix A helpers)
__device__ __forceinline__ int cv_floor_f(float v) { int i = (int)v; return i - ((float)i > v); }
__device__ __forceinline__ int cv_floor_d(double v) { int i = (int)v; return i - ((double)i > v); }

// neighbors(W,H) slot k of cell idx, -1 outside (DLL @VA 0x180048030) — computed, not tabulated
__device__ __forceinline__ int nb9(int idx, int k, int w, int h) {
    int x = idx % w + (k % 3) - 1;
    int y = idx / w + (k / 3) - 1;
    return (x < 0 || x >= w || y < 0 || y >= h) ? -1 : x + y * w;
}

// ---- per-pair scratch layout (int32 units unless noted) ---------------------------------------------
struct Layout {
    int n_scales, n_rot;
    size_t cnt_off;       // [4][400] int32
    size_t counts_off;    // [40] int32
    size_t hist_off[kNumScales];  // [4][400][G_r] int32
    size_t cp_off;        // [n_scales][n_rot][4][400] int16 (stored in int32 units, rounded up)
    size_t zero_words;    // words [0, zero_words) must be zero before assign (cnt, counts, hists)
    size_t total_words;
};

// dense = true: histograms live in global memory (general path); dense = false: they live in shared memory
// (gms_vote_smem_kernel) and the per-pair scratch shrinks to the hypothesis counters and the cell-pair table.
__host__ __device__ inline Layout make_layout(int n_scales, int n_rot, bool dense = true) {
    Layout L;
    L.n_scales = n_scales; L.n_rot = n_rot;
    size_t o = 0;
    L.cnt_off = o; o += 4 * kCellsL;   // both paths: cell totals per shift
    L.counts_off = o; o += 64;
    for (int s = 0; s < kNumScales; ++s) {
        L.hist_off[s] = o;
        if (dense && s < n_scales) { int w = right_grid_w(s); o += (size_t)4 * kCellsL * w * w; }
    }
    L.zero_words = o;
    L.cp_off = o; o += ((size_t)n_scales * n_rot * 4 * kCellsL + 1) / 2;
    L.total_words = (o + 31) & ~(size_t)31;
    return L;
}

// ---- a4 + a5 + a7: normalise, cell indices for all shifts/scales, histogram RMW ----------------------
template <bool kHist>   // kHist = false: only the per-match cell indices (the histograms are built in shared memory later)
__global__ void __launch_bounds__(256) gms_assign_kernel(const PairDesc* __restrict__ pairs, PairResult* results,
                                                         int32_t* scratch, Layout L, uint16_t* lidx,
                                                         uint16_t* ridx, long long chunk_match_base,
                                                         long long chunk_matches) {
    const PairDesc pd = pairs[blockIdx.z];
    int32_t* sp = scratch + (size_t)blockIdx.z * L.total_words;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pd.n_matches; i += gridDim.x * blockDim.x) {
        const long long mi = pd.match_base - chunk_match_base + i;
        int qi = pd.mq ? pd.mq[i] : i;
        int ti = pd.mt ? pd.mt[i] : (int)(pd.key[i] & kTrainIdxMask);
        bool ok = true;
        if (qi < 0 || qi >= pd.n1 || ti < 0 || ti >= pd.n2) { atomicMax(&results[blockIdx.z].status, 4); ok = false; }
        float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
        if (ok) {
            float2 a = reinterpret_cast<const float2*>(pd.kp1)[qi];
            float2 b = reinterpret_cast<const float2*>(pd.kp2)[ti];
            x1 = a.x; y1 = a.y; x2 = b.x; y2 = b.y;
            // supported domain 0 <= x < w, 0 <= y < h (outside is UB in the reference; NaN fails too)
            if (!(x1 >= 0.f && x1 < (float)pd.w1 && y1 >= 0.f && y1 < (float)pd.h1 && x2 >= 0.f &&
                  x2 < (float)pd.w2 && y2 >= 0.f && y2 < (float)pd.h2)) {
                atomicMax(&results[blockIdx.z].status, 3);
                ok = false;
            }
        }
        if (!ok) {
#pragma unroll
            for (int t = 0; t < 4; ++t) lidx[(size_t)t * chunk_matches + mi] = kNoCell;
            for (int s = 0; s < L.n_scales; ++s) ridx[(size_t)s * chunk_matches + mi] = kNoCell;
            continue;
        }
        // normalizePoints (f32 divide), DLL @VA 0x180048420
        const float nx1 = __fdiv_rn(x1, (float)pd.w1), ny1 = __fdiv_rn(y1, (float)pd.h1);
        const float nx2 = __fdiv_rn(x2, (float)pd.w2), ny2 = __fdiv_rn(y2, (float)pd.h2);
        // getGridIndexLeft, DLL @VA 0x180047bc0: f32 multiply, f64 +0.5 on the shifted axis only
        const float fx = __fmul_rn((float)kGridL, nx1), fy = __fmul_rn((float)kGridL, ny1);
        const int xa = cv_floor_f(fx), ya = cv_floor_f(fy);
        const int xb = cv_floor_d(__dadd_rn((double)fx, 0.5)), yb = cv_floor_d(__dadd_rn((double)fy, 0.5));
        int l[4];
        l[0] = (xa >= kGridL || ya >= kGridL) ? -1 : xa + ya * kGridL;
        l[1] = (xb >= kGridL || ya >= kGridL) ? -1 : xb + ya * kGridL;
        l[2] = (xa >= kGridL || yb >= kGridL) ? -1 : xa + yb * kGridL;
        l[3] = (xb >= kGridL || yb >= kGridL) ? -1 : xb + yb * kGridL;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            lidx[(size_t)t * chunk_matches + mi] = l[t] < 0 ? kNoCell : (uint16_t)l[t];
            if (kHist && l[t] >= 0) atomicAdd(&sp[L.cnt_off + t * kCellsL + l[t]], 1);
        }
        for (int s = 0; s < L.n_scales; ++s) {
            // getGridIndexRight, DLL @VA 0x180047d60
            const int w = right_grid_w(s);
            const int rx = cv_floor_f(__fmul_rn((float)w, nx2)), ry = cv_floor_f(__fmul_rn((float)w, ny2));
            const int r = rx + ry * w;   // in [0, w*w) for in-domain points
            ridx[(size_t)s * chunk_matches + mi] = (uint16_t)r;
            if (kHist) {
                int32_t* h = sp + L.hist_off[s];
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (l[t] >= 0) atomicAdd(&h[((size_t)t * kCellsL + l[t]) * (w * w) + r], 1);
            }
        }
    }
}

// ---- a8: verifyCellPairs, one warp per (pair, scale, shift, left cell) ------------------------------
__global__ void __launch_bounds__(256) gms_verify_kernel(int32_t* scratch, Layout L, double factor,
                                                         int n_pairs) {
    const int warps_per_pair = L.n_scales * 4 * kCellsL;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= (long long)warps_per_pair * n_pairs) return;
    const int lane = threadIdx.x & 31;
    const int p = (int)(gw / warps_per_pair);
    int rem = (int)(gw % warps_per_pair);
    const int s = rem / (4 * kCellsL); rem %= 4 * kCellsL;
    const int t = rem / kCellsL;
    const int cell = rem % kCellsL;
    int32_t* sp = scratch + (size_t)p * L.total_words;
    const int w = right_grid_w(s), gr = w * w;
    const int32_t* hist = sp + L.hist_off[s] + (size_t)t * kCellsL * gr;
    const int32_t* cnt = sp + L.cnt_off + t * kCellsL;
    int16_t* cpv = reinterpret_cast<int16_t*>(sp + L.cp_off);

    const int ccount = cnt[cell];
    int cp = -1;
    if (ccount > 0) {
        // argmax with strict '>' from maxv = 0 => lowest j among the maxima, count > 0
        // 64-bit key: the dense path serves pairs with >= 65536 matches, so a count may exceed 21 bits
        const int32_t* row = hist + (size_t)cell * gr;
        unsigned long long bestk = 0;
        for (int j = lane; j < gr; j += 32) {
            const unsigned long long c = (unsigned long long)(uint32_t)row[j];
            const unsigned long long k = (c << 11) | (unsigned long long)(2047 - j);
            bestk = (c > 0 && k > bestk) ? k : bestk;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestk, o);
            bestk = other > bestk ? other : bestk;
        }
        cp = 2047 - (int)(bestk & 2047ull);
    }
    for (int r = 0; r < L.n_rot; ++r) {
        int out = cp;
        if (cp >= 0) {
            int v = 0, c = 0;
            bool valid = false;
            if (lane < 9) {
                const int ll = nb9(cell, lane, kGridL, kGridL);
                const int rr = nb9(cp, c_rot[r][lane] - 1, w, w);
                valid = (ll != -1 && rr != -1);
                if (valid) { v = hist[(size_t)ll * gr + rr]; c = cnt[ll]; }
            }
            const int score = __reduce_add_sync(0xffffffffu, v);
            const int tsum = __reduce_add_sync(0xffffffffu, c);       // exact: integers
            const int num = __popc(__ballot_sync(0xffffffffu, valid));
            // thresh = factor * sqrt(T / n): divsd, sqrtsd, mulsd — three separately rounded f64 ops
            const double thresh = __dmul_rn(factor, __dsqrt_rn(__ddiv_rn((double)tsum, (double)num)));
            if ((double)score < thresh) out = -2;
        }
        if (lane == 0) cpv[(((size_t)s * L.n_rot + r) * 4 + t) * kCellsL + cell] = (int16_t)out;
    }
}

// ---- a7 + a8 in shared memory: histogram of one (pair, scale, shift) for a BAND of left-grid rows (+1 halo row
// on each side, which the 3x3 support needs), built with shared-memory atomics on 16-bit counters (two per word;
// valid while a pair has < 65536 matches, the launcher checks), then per-cell argmax and verification exactly as
// gms_verify_kernel.  Nothing but the 400-entry cell-pair table goes back to global memory.
__global__ void __launch_bounds__(512) gms_vote_smem_kernel(const PairDesc* __restrict__ pairs, int32_t* scratch, Layout L,
                                                            double factor, const uint16_t* __restrict__ lidx,
                                                            const uint16_t* __restrict__ ridx, long long chunk_match_base,
                                                            long long chunk_matches, int s, int band_rows) {
    extern __shared__ uint32_t sm_u32[];
    const PairDesc pd = pairs[blockIdx.z];
    const int t = blockIdx.y;
    const int w = right_grid_w(s), gr = w * w;
    const int y0 = blockIdx.x * band_rows, y1 = min(kGridL, y0 + band_rows);
    const int ys0 = max(0, y0 - 1), ys1 = min(kGridL, y1 + 1);
    const int cells_sm = (ys1 - ys0) * kGridL;                 // left cells held in shared memory
    const int hist_words = cells_sm * gr / 2;                   // gr is even for every scale
    uint32_t* h32 = sm_u32;
    int* cnt = reinterpret_cast<int*>(sm_u32 + hist_words);
    {
        uint4* z = reinterpret_cast<uint4*>(sm_u32);                   // (hist_words + cells_sm) is a multiple of 4
        for (int i = threadIdx.x; i < (hist_words + cells_sm) / 4; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const uint16_t* lt = lidx + (size_t)t * chunk_matches + (pd.match_base - chunk_match_base);
    const uint16_t* rs = ridx + (size_t)s * chunk_matches + (pd.match_base - chunk_match_base);
    const int lo = ys0 * kGridL, hi = ys1 * kGridL;
    // 4 independent loads in flight per thread: this loop is latency-bound, not bandwidth-bound
    for (int i0 = threadIdx.x; i0 < pd.n_matches; i0 += 4 * blockDim.x) {
        int l[4], r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * blockDim.x;
            l[k] = i < pd.n_matches ? (int)__ldg(lt + i) : 0xFFFF;
            r[k] = i < pd.n_matches ? (int)__ldg(rs + i) : 0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (l[k] >= lo && l[k] < hi) {                      // kNoCell (0xFFFF) fails this test too
                const int idx = (l[k] - lo) * gr + r[k];
                atomicAdd(&cnt[l[k] - lo], 1);
                atomicAdd(&h32[idx >> 1], 1u << ((idx & 1) * 16));
            }
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int16_t* cpv = reinterpret_cast<int16_t*>(scratch + (size_t)blockIdx.z * L.total_words + L.cp_off);
    const uint16_t* h16 = reinterpret_cast<const uint16_t*>(h32);
    for (int cell = y0 * kGridL + warp; cell < y1 * kGridL; cell += nwarps) {
        const int local = cell - lo;
        int cp = -1;
        if (cnt[local] > 0) {
            const uint32_t* row = h32 + (size_t)local * (gr / 2);
            uint32_t bestk = 0;
            for (int wd = lane; wd < gr / 2; wd += 32) {
                const uint32_t u = row[wd];
                const uint32_t c0 = u & 0xFFFFu, c1 = u >> 16;
                const uint32_t k0 = (c0 << 11) | (uint32_t)(2047 - 2 * wd), k1 = (c1 << 11) | (uint32_t)(2046 - 2 * wd);
                bestk = (c0 > 0 && k0 > bestk) ? k0 : bestk;
                bestk = (c1 > 0 && k1 > bestk) ? k1 : bestk;
            }
            bestk = __reduce_max_sync(0xffffffffu, bestk);
            cp = 2047 - (int)(bestk & 2047u);
        }
        for (int r = 0; r < L.n_rot; ++r) {
            int out = cp;
            if (cp >= 0) {
                int v = 0, c = 0;
                bool valid = false;
                if (lane < 9) {
                    const int ll = nb9(cell, lane, kGridL, kGridL);
                    const int rr = nb9(cp, c_rot[r][lane] - 1, w, w);
                    valid = (ll != -1 && rr != -1);
                    if (valid) { v = h16[(size_t)(ll - lo) * gr + rr]; c = cnt[ll - lo]; }
                }
                const int score = __reduce_add_sync(0xffffffffu, v);
                const int tsum = __reduce_add_sync(0xffffffffu, c);
                const int num = __popc(__ballot_sync(0xffffffffu, valid));
                const double thresh = __dmul_rn(factor, __dsqrt_rn(__ddiv_rn((double)tsum, (double)num)));
                if ((double)score < thresh) out = -2;
            }
            if (lane == 0) cpv[(((size_t)s * L.n_rot + r) * 4 + t) * kCellsL + cell] = (int16_t)out;
        }
    }
}

// ---- a4 + a5 (+ the cell totals of a7) for the shared-memory path: one CTA per (pair, slice of <= kAssignSlice matches).
// Per match: normalise, 4 left cells, S right cells -> lidx / ridx (as gms_assign_kernel<false>); the number of matches per
// (shift, left cell) is accumulated with shared-memory atomics and flushed once per CTA (non-zero cells only), so the
// vote kernels need neither a second atomic per match nor a row-sum scan.
constexpr int kAssignSlice = 16384;
__global__ void __launch_bounds__(1024) gms_assign_cnt_kernel(const PairDesc* __restrict__ pairs, PairResult* results,
                                                               int32_t* scratch, Layout L, uint16_t* lidx, uint16_t* ridx,
                                                               long long chunk_match_base, long long chunk_matches) {
    __shared__ int cnt_s[4 * kCellsL];
    const PairDesc pd = pairs[blockIdx.z];
    const int i_begin = blockIdx.x * kAssignSlice;
    if (i_begin >= pd.n_matches) return;                        // whole CTA leaves together
    const int i_end = min(pd.n_matches, i_begin + kAssignSlice);
    for (int k = threadIdx.x; k < 4 * kCellsL; k += blockDim.x) cnt_s[k] = 0;
    __syncthreads();
    int bad = 0;
    // the per-match chain is two dependent gathers (match index -> keypoint) in front of a little arithmetic: four matches
    // per thread are kept in flight through both of them
    constexpr int U = 4;
    for (int i0 = i_begin + threadIdx.x; i0 < i_end; i0 += U * blockDim.x) {
        int qi[U], ti[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k * blockDim.x;
            qi[k] = ti[k] = -1;
            if (i < i_end) {
                qi[k] = pd.mq ? pd.mq[i] : i;
                ti[k] = pd.mt ? pd.mt[i] : (int)(pd.key[i] & kTrainIdxMask);
            }
        }
        float2 pa[U], pb[U];
        bool inr[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            inr[k] = qi[k] >= 0 && qi[k] < pd.n1 && ti[k] >= 0 && ti[k] < pd.n2;
            pa[k] = pb[k] = make_float2(0.f, 0.f);
            if (inr[k]) {
                pa[k] = reinterpret_cast<const float2*>(pd.kp1)[qi[k]];
                pb[k] = reinterpret_cast<const float2*>(pd.kp2)[ti[k]];
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = i0 + k * blockDim.x;
            if (i >= i_end) continue;
            const long long mi = pd.match_base - chunk_match_base + i;
            bool ok = inr[k];
            if (!ok) bad = max(bad, 4);
            const float x1 = pa[k].x, y1 = pa[k].y, x2 = pb[k].x, y2 = pb[k].y;
            // supported domain 0 <= x < w, 0 <= y < h (outside is UB in the reference; NaN fails too)
            if (ok && !(x1 >= 0.f && x1 < (float)pd.w1 && y1 >= 0.f && y1 < (float)pd.h1 && x2 >= 0.f && x2 < (float)pd.w2 &&
                        y2 >= 0.f && y2 < (float)pd.h2)) {
                bad = max(bad, 3);
                ok = false;
            }
            if (!ok) {
#pragma unroll
                for (int t = 0; t < 4; ++t) lidx[(size_t)t * chunk_matches + mi] = kNoCell;
                for (int s = 0; s < L.n_scales; ++s) ridx[(size_t)s * chunk_matches + mi] = kNoCell;
                continue;
            }
            const float nx1 = __fdiv_rn(x1, (float)pd.w1), ny1 = __fdiv_rn(y1, (float)pd.h1);     // normalizePoints @VA 0x180048420
            const float nx2 = __fdiv_rn(x2, (float)pd.w2), ny2 = __fdiv_rn(y2, (float)pd.h2);
            const float fx = __fmul_rn((float)kGridL, nx1), fy = __fmul_rn((float)kGridL, ny1);   // getGridIndexLeft @VA 0x180047bc0
            const int xa = cv_floor_f(fx), ya = cv_floor_f(fy);
            const int xb = cv_floor_d(__dadd_rn((double)fx, 0.5)), yb = cv_floor_d(__dadd_rn((double)fy, 0.5));
            int l[4];
            l[0] = (xa >= kGridL || ya >= kGridL) ? -1 : xa + ya * kGridL;
            l[1] = (xb >= kGridL || ya >= kGridL) ? -1 : xb + ya * kGridL;
            l[2] = (xa >= kGridL || yb >= kGridL) ? -1 : xa + yb * kGridL;
            l[3] = (xb >= kGridL || yb >= kGridL) ? -1 : xb + yb * kGridL;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                lidx[(size_t)t * chunk_matches + mi] = l[t] < 0 ? kNoCell : (uint16_t)l[t];
                if (l[t] >= 0) atomicAdd(&cnt_s[t * kCellsL + l[t]], 1);
            }
            for (int s = 0; s < L.n_scales; ++s) {                                                // getGridIndexRight @VA 0x180047d60
                const int w = right_grid_w(s);
                const int rx = cv_floor_f(__fmul_rn((float)w, nx2)), ry = cv_floor_f(__fmul_rn((float)w, ny2));
                ridx[(size_t)s * chunk_matches + mi] = (uint16_t)(rx + ry * w);
            }
        }
    }
    if (bad) atomicMax(&results[blockIdx.z].status, bad);
    __syncthreads();
    int32_t* cnt_g = scratch + (size_t)blockIdx.z * L.total_words + L.cnt_off;
    const bool single = pd.n_matches <= kAssignSlice;          // one CTA owns the pair: plain stores into the zeroed scratch
    for (int k = threadIdx.x; k < 4 * kCellsL; k += blockDim.x) {
        const int c = cnt_s[k];
        if (c) { if (single) cnt_g[k] = c; else atomicAdd(&cnt_g[k], c); }
    }
}

// ---- a7 + a8 in shared memory, second generation.  One CTA per (pair, shift, band of left-grid rows + 1 halo row each
// side) for scale S.  The band's left-cell x right-cell histogram lives in shared memory as 16-bit counters (two per word;
// a pair has < 65536 matches on this path).  ONE atomic builds it and, from the value the atomic returns, a second one
// maintains the row arg-max directly: key = (new count << 11 | 2047 - r) only grows while a counter grows, so after the last
// vote best[l] holds (max count, lowest right cell attaining it) -- verifyCellPairs' strict-'>' scan (DLL @VA 0x180048d10)
// without scanning any row.  Cell totals come from gms_assign_cnt_kernel.  The kernel is instruction-issue bound (ncu:
// profiles/r2_notes.md), so: cell indices arrive as 128-bit loads of 8 matches, the right-grid width is a template constant,
// and verification runs one THREAD per (left cell, rotation) -- 9-slot support gather, the three separately rounded f64
// operations -- instead of one warp per cell.
// (ROT[r][k] - 1) = v as (dx, dy) = (v mod 3 - 1, v div 3 - 1) in the 3x3 neighbourhood of the right cell
__constant__ int8_t c_rdx[kNumRot][9] = {{-1, 0, 1, -1, 0, 1, -1, 0, 1}, {-1, -1, 0, -1, 0, 1, 0, 1, 1}, {-1, -1, -1, 0, 0, 0, 1, 1, 1}, {0, -1, -1, 1, 0, -1, 1, 1, 0}, {1, 0, -1, 1, 0, -1, 1, 0, -1}, {1, 1, 0, 1, 0, -1, 0, -1, -1}, {1, 1, 1, 0, 0, 0, -1, -1, -1}, {0, 1, 1, -1, 0, 1, -1, -1, 0}};
__constant__ int8_t c_rdy[kNumRot][9] = {{-1, -1, -1, 0, 0, 0, 1, 1, 1}, {0, -1, -1, 1, 0, -1, 1, 1, 0}, {1, 0, -1, 1, 0, -1, 1, 0, -1}, {1, 1, 0, 1, 0, -1, 0, -1, -1}, {1, 1, 1, 0, 0, 0, -1, -1, -1}, {0, 1, 1, -1, 0, 1, -1, -1, 0}, {-1, 0, 1, -1, 0, 1, -1, 0, 1}, {-1, -1, 0, -1, 0, 1, 0, 1, 1}};

template <int S>
__global__ void __launch_bounds__(1024) gms_vote2_kernel(const PairDesc* __restrict__ pairs, int32_t* scratch, Layout L,
                                                        double factor, const uint16_t* __restrict__ lidx,
                                                        const uint16_t* __restrict__ ridx, long long chunk_match_base,
                                                        long long chunk_matches, int band_rows) {
    extern __shared__ uint32_t sm_u32[];
    constexpr int w = S == 0 ? 20 : S == 1 ? 10 : S == 2 ? 14 : S == 3 ? 28 : 40, gr = w * w;
    const PairDesc pd = pairs[blockIdx.z];
    const int t = blockIdx.y;
    const int y0 = blockIdx.x * band_rows, y1 = min(kGridL, y0 + band_rows);
    const int ys0 = max(0, y0 - 1), ys1 = min(kGridL, y1 + 1);
    const int cells_sm = (ys1 - ys0) * kGridL;
    const int hist_words = cells_sm * gr / 2;                   // gr is even for every scale
    uint32_t* h32 = sm_u32;
    uint32_t* best = sm_u32 + hist_words;
    int* cnt_s = reinterpret_cast<int*>(best + cells_sm);
    const int lo = ys0 * kGridL, hi = ys1 * kGridL;
    int32_t* sp = scratch + (size_t)blockIdx.z * L.total_words;
    {
        uint4* z = reinterpret_cast<uint4*>(sm_u32);            // (hist_words + cells_sm) is a multiple of 4
        for (int i = threadIdx.x; i < (hist_words + cells_sm) / 4; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        const int32_t* cnt_g = sp + L.cnt_off + t * kCellsL + lo;
        for (int i = threadIdx.x; i < cells_sm; i += blockDim.x) cnt_s[i] = __ldg(cnt_g + i);
    }
    __syncthreads();
    // 8 matches per 128-bit load; the pair's rows start `shift` elements into an aligned group (rows of other pairs in the
    // first / last group are masked out)
    const long long e0 = pd.match_base - chunk_match_base;
    const int shift = (int)(e0 & 7);
    const uint4* l4 = reinterpret_cast<const uint4*>(lidx + (size_t)t * chunk_matches + (e0 - shift));
    const uint4* r4 = reinterpret_cast<const uint4*>(ridx + (size_t)S * chunk_matches + (e0 - shift));
    const int n = pd.n_matches;
    const int ngroups = n > 0 ? (shift + n + 7) >> 3 : 0;
    const unsigned span = (unsigned)(hi - lo);
    auto vote8 = [&](uint4 lv, uint4 rv, int g) {
        uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
        if (g == 0 || g == ngroups - 1) {                       // edge groups: drop elements outside [0, n)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = g * 8 + j - shift;
                if (i < 0 || i >= n) lw[j >> 1] |= 0xFFFFu << (16 * (j & 1));   // kNoCell fails the band test
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned rel = ((lw[j >> 1] >> (16 * (j & 1))) & 0xFFFFu) - (unsigned)lo;
            if (rel < span) {
                const unsigned r = (rw[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
                const unsigned idx = rel * gr + r;
                const unsigned sh = (idx & 1u) * 16u;
                const uint32_t old = atomicAdd(&h32[idx >> 1], 1u << sh);
                const uint32_t c = ((old >> sh) & 0xFFFFu) + 1u;
                atomicMax(&best[rel], (c << 11) | (2047u - r));
            }
        }
    };
    for (int g = threadIdx.x; g < ngroups; g += 2 * blockDim.x) {   // two groups (16 matches) in flight per thread
        const int g2 = g + blockDim.x;
        const uint4 la = __ldg(l4 + g), ra = __ldg(r4 + g);
        uint4 lb = make_uint4(0, 0, 0, 0), rb = lb;
        if (g2 < ngroups) { lb = __ldg(l4 + g2); rb = __ldg(r4 + g2); }
        vote8(la, ra, g);
        if (g2 < ngroups) vote8(lb, rb, g2);
    }
    __syncthreads();
    int16_t* cpv = reinterpret_cast<int16_t*>(sp + L.cp_off);
    const uint16_t* h16 = reinterpret_cast<const uint16_t*>(h32);
    const int n_own = (y1 - y0) * kGridL;
    for (int task = threadIdx.x; task < n_own * L.n_rot; task += blockDim.x) {
        const int r = task / n_own, cell = y0 * kGridL + (task - r * n_own);
        const uint32_t bk = best[cell - lo];
        int out = -1;                                           // no vote in the row <=> its total is 0
        if (bk) {
            const int cp = 2047 - (int)(bk & 2047u);
            const int x = cell % kGridL, y = cell / kGridL, cx = cp % w, cy = cp / w;
            int score = 0, tsum = 0, num = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int xx = x + (k % 3) - 1, yy = y + (k / 3) - 1;
                const int rx = cx + c_rdx[r][k], ry = cy + c_rdy[r][k];
                if ((unsigned)xx < (unsigned)kGridL && (unsigned)yy < (unsigned)kGridL && (unsigned)rx < (unsigned)w &&
                    (unsigned)ry < (unsigned)w) {
                    const int ll = xx + yy * kGridL - lo;
                    score += h16[ll * gr + rx + ry * w];
                    tsum += cnt_s[ll];
                    ++num;
                }
            }
            // thresh = factor * sqrt(T / n): divsd, sqrtsd, mulsd -- three separately rounded f64 operations
            const double thresh = __dmul_rn(factor, __dsqrt_rn(__ddiv_rn((double)tsum, (double)num)));
            out = ((double)score < thresh) ? -2 : cp;
        }
        cpv[(((size_t)S * L.n_rot + r) * 4 + t) * kCellsL + cell] = (int16_t)out;
    }
}

__device__ __forceinline__ bool is_inlier(const int16_t* cpv_h, const uint16_t* lidx, uint16_t r,
                                          long long chunk_matches, long long mi) {
    bool inl = false;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const uint16_t l = lidx[(size_t)t * chunk_matches + mi];
        if (l != kNoCell) inl |= (cpv_h[t * kCellsL + l] == (int16_t)r);
    }
    return inl;
}

// ---- a9: mark + count, for every hypothesis (blockIdx.y).  With write_mask (no-flag call) also the mask.
__global__ void __launch_bounds__(256) gms_count_kernel(const PairDesc* __restrict__ pairs, int32_t* scratch,
                                                        Layout L, const uint16_t* __restrict__ lidx,
                                                        const uint16_t* __restrict__ ridx,
                                                        long long chunk_match_base, long long chunk_matches,
                                                        int write_mask) {
    const PairDesc pd = pairs[blockIdx.z];
    int32_t* sp = scratch + (size_t)blockIdx.z * L.total_words;
    const int hyp = blockIdx.y;                    // s * n_rot + r
    const int s = hyp / L.n_rot;
    const int16_t* cpv_h = reinterpret_cast<const int16_t*>(sp + L.cp_off) + (size_t)hyp * 4 * kCellsL;
    int local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pd.n_matches; i += gridDim.x * blockDim.x) {
        const long long mi = pd.match_base - chunk_match_base + i;
        const uint16_t r = ridx[(size_t)s * chunk_matches + mi];
        const bool inl = (r != kNoCell) && is_inlier(cpv_h, lidx, r, chunk_matches, mi);
        local += inl;
        if (write_mask) pd.mask[i] = inl;
    }
    local = __reduce_add_sync(0xffffffffu, local);
    __shared__ int wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += wsum[k];
        if (tot) atomicAdd(&sp[L.counts_off + hyp], tot);
    }
}

// ---- a9 for the rotation / scale search: one CTA per (pair, scale, slice of matches) counts the inliers of ALL rotations
// of that scale at once.  The scale's cell-pair tables sit in shared memory TRANSPOSED to [shift][left cell][rotation]
// (8 x int16 = one 128-bit word per (shift, cell)): a match reads its 4 left cells and its right cell once and does 4
// shared-memory loads, each compared against the right cell for all rotations with packed 16-bit compares (the
// per-hypothesis kernel re-read the indices and gathered from global memory once per hypothesis).
__global__ void __launch_bounds__(256) gms_count_scale_kernel(const PairDesc* __restrict__ pairs, int32_t* scratch, Layout L,
                                                              const uint16_t* __restrict__ lidx, const uint16_t* __restrict__ ridx,
                                                              long long chunk_match_base, long long chunk_matches) {
    extern __shared__ uint32_t cs_u32[];                        // [4][400][8] int16
    __shared__ int cnt_s[kNumRot];
    const PairDesc pd = pairs[blockIdx.z];
    const int s = blockIdx.y;
    int32_t* sp = scratch + (size_t)blockIdx.z * L.total_words;
    const int16_t* src = reinterpret_cast<const int16_t*>(sp + L.cp_off) + (size_t)s * L.n_rot * 4 * kCellsL;   // [rot][shift][cell]
    int16_t* cp_t = reinterpret_cast<int16_t*>(cs_u32);
    for (int i = threadIdx.x; i < 4 * kCellsL * kNumRot; i += blockDim.x) {
        const int rot = i & 7, tc = i >> 3;                     // tc = shift * 400 + cell
        cp_t[i] = rot < L.n_rot ? src[(size_t)rot * 4 * kCellsL + tc] : (int16_t)-3;   // -3 never equals a right cell
    }
    if (threadIdx.x < kNumRot) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    const uint4* cp4 = reinterpret_cast<const uint4*>(cs_u32);
    const long long mb = pd.match_base - chunk_match_base;
    const int lane = threadIdx.x & 31;
    for (int base = blockIdx.x * blockDim.x; base < pd.n_matches; base += gridDim.x * blockDim.x) {   // warp-uniform trip count
        const int i = base + threadIdx.x;
        unsigned m = 0;                                         // bit rot: inlier under rotation rot
        if (i < pd.n_matches) {
            const uint32_t r = ridx[(size_t)s * chunk_matches + mb + i];
            if (r != kNoCell) {
                const uint32_t rr = r | (r << 16);
                uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t l = lidx[(size_t)t * chunk_matches + mb + i];
                    if (l != kNoCell) {
                        const uint4 c = cp4[t * kCellsL + l];
                        acc.x |= __vcmpeq2(c.x, rr); acc.y |= __vcmpeq2(c.y, rr);
                        acc.z |= __vcmpeq2(c.z, rr); acc.w |= __vcmpeq2(c.w, rr);
                    }
                }
                // halves are 0xFFFF / 0x0000: bit 0 and bit 16 of each word -> rotation bits 0..7
                m = (acc.x & 1u) | ((acc.x >> 15) & 2u) | ((acc.y & 1u) << 2) | ((acc.y >> 13) & 8u) | ((acc.z & 1u) << 4) |
                    ((acc.z >> 11) & 32u) | ((acc.w & 1u) << 6) | ((acc.w >> 9) & 128u);
            }
        }
        for (int rot = 0; rot < L.n_rot; ++rot) {
            const int c = __popc(__ballot_sync(0xffffffffu, (m >> rot) & 1u));
            if (lane == 0 && c) atomicAdd(&cnt_s[rot], c);
        }
    }
    __syncthreads();
    if (threadIdx.x < L.n_rot && cnt_s[threadIdx.x]) atomicAdd(&sp[L.counts_off + s * L.n_rot + threadIdx.x], cnt_s[threadIdx.x]);
}

// ---- a10: getInlierMask — best hypothesis, scale-major / rotation-minor, strict '>' (first wins) -----
__global__ void gms_select_kernel(const PairDesc* __restrict__ pairs, PairResult* results,
                                  const int32_t* scratch, Layout L, int n_pairs, int flags_on) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int32_t* counts = scratch + (size_t)p * L.total_words + L.counts_off;
    PairResult r = results[p];
    if (!flags_on) {
        r.n_inliers = counts[0]; r.best_hyp = 0; r.mask_len = pairs[p].n_matches;
    } else {
        int best = 0, bh = -1;
        for (int s = 0; s < L.n_scales; ++s)
            for (int k = 0; k < L.n_rot; ++k) {
                int c = counts[s * L.n_rot + k];
                if (c > best) { best = c; bh = s * kNumRot + k; }
            }
        r.n_inliers = best; r.best_hyp = bh; r.mask_len = bh < 0 ? 0 : pairs[p].n_matches;
    }
    results[p] = r;
}

// ---- mask of the winning hypothesis (flag calls only) ----------------------------------------------
__global__ void __launch_bounds__(256) gms_mask_kernel(const PairDesc* __restrict__ pairs,
                                                       const PairResult* __restrict__ results,
                                                       const int32_t* scratch, Layout L,
                                                       const uint16_t* __restrict__ lidx,
                                                       const uint16_t* __restrict__ ridx,
                                                       long long chunk_match_base, long long chunk_matches) {
    const PairDesc pd = pairs[blockIdx.z];
    const int bh = results[blockIdx.z].best_hyp;
    const int32_t* sp = scratch + (size_t)blockIdx.z * L.total_words;
    const int s = bh < 0 ? 0 : bh / kNumRot, r8 = bh < 0 ? 0 : bh % kNumRot;
    const int hyp = s * L.n_rot + r8;
    const int16_t* cpv_h = reinterpret_cast<const int16_t*>(sp + L.cp_off) + (size_t)hyp * 4 * kCellsL;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pd.n_matches; i += gridDim.x * blockDim.x) {
        const long long mi = pd.match_base - chunk_match_base + i;
        bool inl = false;
        if (bh >= 0) {
            const uint16_t r = ridx[(size_t)s * chunk_matches + mi];
            inl = (r != kNoCell) && is_inlier(cpv_h, lidx, r, chunk_matches, mi);
        }
        pd.mask[i] = inl;
    }
}

// ---- compacted outputs (§8f-1): offsets = exclusive scan of the inlier counts, then an ORDERED per-pair compaction
__global__ void __launch_bounds__(1024) gms_compact_offsets_kernel(const PairResult* __restrict__ results, int n_pairs,
                                                                    long long* base_io, long long* __restrict__ offsets) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = *base_io;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int p0 = 0; p0 < n_pairs; p0 += 1024) {
        const int p = p0 + threadIdx.x;
        const long long v = (p < n_pairs && results[p].mask_len > 0) ? results[p].n_inliers : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        long long woff = 0, tot = 0;
        for (int k = 0; k < 32; ++k) { const long long w = wsum[k]; if (k < warp) woff += w; tot += w; }
        if (p < n_pairs) offsets[p] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) { offsets[n_pairs] = carry; *base_io = carry; }
}

__global__ void __launch_bounds__(1024) gms_compact_kernel(const PairDesc* __restrict__ pairs, const PairResult* __restrict__ results,
                                                            const long long* __restrict__ offsets, long long capacity,
                                                            DMatchRec* __restrict__ matches, float2* __restrict__ pts1,
                                                            float2* __restrict__ pts2, int index_pairs) {
    const PairDesc pd = pairs[blockIdx.x];
    if (results[blockIdx.x].mask_len <= 0 || results[blockIdx.x].n_inliers <= 0) return;
    const long long out0 = offsets[blockIdx.x];
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < pd.n_matches; base += 1024) {
        const int i = base + threadIdx.x;
        const int m = (i < pd.n_matches) ? (pd.mask[i] != 0) : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int v = wsum[k]; if (k < warp) woff += v; tot += v; }
        const long long pos = out0 + carry + woff + __popc(bal & ((1u << lane) - 1u));
        if (m && pos < capacity) {
            const uint32_t key = pd.key ? pd.key[i] : 0u;
            const int qi = pd.mq ? pd.mq[i] : i;
            const int ti = pd.mt ? pd.mt[i] : (int)(key & kTrainIdxMask);
            if (matches && index_pairs) {
                reinterpret_cast<int2*>(matches)[pos] = make_int2(qi, ti);
            } else if (matches) {
                DMatchRec r;
                r.queryIdx = qi; r.trainIdx = ti; r.imgIdx = 0; r.distance = pd.key ? (float)(key >> kTrainIdxBits) : 0.f;
                *reinterpret_cast<int4*>(matches + pos) = *reinterpret_cast<const int4*>(&r);
            }
            if (pts1) pts1[pos] = reinterpret_cast<const float2*>(pd.kp1)[qi];
            if (pts2) pts2[pos] = reinterpret_cast<const float2*>(pd.kp2)[ti];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
}

}  // namespace

int launch_gms_compact(const PairDesc* d_pairs, const PairResult* d_results, int n_pairs, long long* d_base,
                       long long* d_offsets, long long capacity, DMatchRec* d_matches, float* d_pts1, float* d_pts2,
                       cudaStream_t st, bool index_pairs) {
    if (n_pairs <= 0) return 0;
    gms_compact_offsets_kernel<<<1, 1024, 0, st>>>(d_results, n_pairs, d_base, d_offsets);
    kmark("gms_compact_offsets", st);
    gms_compact_kernel<<<n_pairs, 1024, 0, st>>>(d_pairs, d_results, d_offsets, capacity, d_matches,
                                                 reinterpret_cast<float2*>(d_pts1), reinterpret_cast<float2*>(d_pts2), index_pairs ? 1 : 0);
    kmark("gms_compact", st);
    return 2;
}

static bool gms_dense_path(int max_matches, int force_dense) {
    static const bool env_dense = getenv("SFMGMS_GMS_DENSE") != nullptr;
    return force_dense || env_dense || max_matches >= 65536;
}

size_t gms_counts_offset_words(int n_scales, int n_rot, int max_matches, int force_dense) {
    return make_layout(n_scales, n_rot, gms_dense_path(max_matches, force_dense)).counts_off;
}

long long gms_match_rows(const PairDesc* h_pairs, int n) {
    if (n <= 0) return 0;
    const PairDesc& last = h_pairs[n - 1];
    const long long rows_last = last.mq ? last.n_matches : (last.n1 > last.n_matches ? last.n1 : last.n_matches);
    return last.match_base + rows_last - h_pairs[0].match_base;
}
size_t gms_scratch_bytes_per_pair(int n_scales) { return make_layout(n_scales, kNumRot, true).total_words * 4; }
size_t gms_match_scratch_bytes(long long n_matches_total, int n_scales) {
    return (size_t)(n_matches_total + 8) * 2 * (4 + n_scales) + 1024;
}

// Left-grid rows per shared-memory band for scale s: (rows + 2 halo) * 20 cells * (G_r u16 + one int) must fit.
static int smem_band_rows(int s) {
    static const int b0 = [] {   // tuning experiments only; clamped to a valid band height
        const int v = getenv("SFMGMS_GMS_BAND0") ? atoi(getenv("SFMGMS_GMS_BAND0")) : 10;
        return v < 1 ? 1 : (v > 10 ? 10 : v);
    }();
    return s == 0 ? b0 : s == 1 ? 20 : s == 2 ? 10 : s == 3 ? 4 : 1;
}
static size_t smem_band_bytes(int s, int extra_per_cell = 0) {
    const int w = right_grid_w(s), gr = w * w, b = smem_band_rows(s);
    const int rows = b >= kGridL ? kGridL : b + 2;
    return (size_t)rows * kGridL * ((size_t)gr * 2 + 4 + extra_per_cell);
}

// Runs GMS for n_pairs pairs.  Default: histograms in shared memory (one launch per scale for the whole batch).
// Fallback (a pair with >= 65536 matches, or SFMGMS_GMS_DENSE=1): dense histograms in global memory, in chunks
// sized so that a chunk's histograms fit the scratch budget (kept L2-resident by the caller's choice).
int launch_gms(const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int with_rotation, int with_scale,
               double factor, PairResult* d_results, void* d_hist_scratch, size_t hist_scratch_bytes,
               void* d_match_scratch, cudaStream_t st, int force_dense) {
    if (n_pairs <= 0) return 0;
    const int n_scales = with_scale ? kNumScales : 1;
    const int n_rot = with_rotation ? kNumRot : 1;
    const int flags_on = (with_rotation || with_scale) ? 1 : 0;
    int max_all = 0;
    for (int p = 0; p < n_pairs; ++p) if (h_pairs[p].n_matches > max_all) max_all = h_pairs[p].n_matches;
    const bool dense = gms_dense_path(max_all, force_dense);
    const Layout L = make_layout(n_scales, n_rot, dense);
    const size_t per_pair = L.total_words * 4;
    int chunk_cap = (int)(hist_scratch_bytes / per_pair);
    if (chunk_cap < 1) return -1;
    if (chunk_cap > 32768) chunk_cap = 32768;
    int launches = 0;
    int32_t* scratch = static_cast<int32_t*>(d_hist_scratch);
    static const bool use_v1 = getenv("SFMGMS_GMS_V1") != nullptr;   // previous-generation vote kernel (A/B measurements only)
    if (!dense) {
        if (cudaFuncSetAttribute(gms_vote_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
        static bool ready_dev[64] = {false};     // function attributes and __constant__ copies are PER DEVICE
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
        bool& v2_ready = ready_dev[dev];
        if (!v2_ready) {
            const void* fns[5] = {(const void*)gms_vote2_kernel<0>, (const void*)gms_vote2_kernel<1>, (const void*)gms_vote2_kernel<2>,
                                  (const void*)gms_vote2_kernel<3>, (const void*)gms_vote2_kernel<4>};
            for (const void* f : fns)
                if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return -1;
            v2_ready = true;
        }
    }
    for (int c0 = 0; c0 < n_pairs; c0 += chunk_cap) {
        const int cn = (n_pairs - c0 < chunk_cap) ? n_pairs - c0 : chunk_cap;
        int max_m = 0;
        for (int p = c0; p < c0 + cn; ++p) if (h_pairs[p].n_matches > max_m) max_m = h_pairs[p].n_matches;
        // rows of the per-match arrays spanned by this chunk: match_base advances by a pair's ROW count (n1 for
        // fused pairs, even when an empty train image leaves it with 0 matches), not by n_matches
        const long long cbase = h_pairs[c0].match_base;
        const long long cm = (gms_match_rows(h_pairs + c0, cn) + 7) & ~7LL;   // rows of lidx/ridx stay 16-byte aligned
        // per-match cell indices for this chunk: lidx[4][cm], ridx[n_scales][cm] (uint16)
        uint16_t* lidx = static_cast<uint16_t*>(d_match_scratch);
        uint16_t* ridx = lidx + (size_t)4 * cm;
        // zero cnt/counts/hist of every pair in the chunk (cp is fully rewritten by verify)
        cudaMemsetAsync(scratch, 0, (size_t)cn * per_pair, st);
        kmark("gms_memset", st);
        const int bx = max_m > 0 ? (max_m + 255) / 256 : 1;
        if (dense) {
            if (max_m > 0) {
                gms_assign_kernel<true><<<dim3(bx, 1, cn), 256, 0, st>>>(d_pairs + c0, d_results + c0, scratch, L, lidx, ridx,
                                                                       cbase, cm);
                ++launches; kmark("gms_assign_dense", st);
            }
            const long long warps = (long long)cn * n_scales * 4 * kCellsL;
            gms_verify_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(scratch, L, factor, cn);
            ++launches; kmark("gms_verify_dense", st);
        } else if (use_v1) {
            if (max_m > 0) {
                gms_assign_kernel<false><<<dim3(bx, 1, cn), 256, 0, st>>>(d_pairs + c0, d_results + c0, scratch, L, lidx, ridx,
                                                                        cbase, cm);
                ++launches; kmark("gms_assign", st);
            }
            for (int s = 0; s < n_scales; ++s) {
                const int b = smem_band_rows(s);
                const int bands = (kGridL + b - 1) / b;
                gms_vote_smem_kernel<<<dim3(bands, 4, cn), 512, smem_band_bytes(s), st>>>(d_pairs + c0, scratch, L, factor, lidx,
                                                                                        ridx, cbase, cm, s, b);
                ++launches; kmark("gms_vote_smem", st);
            }
        } else {
            if (max_m > 0) {
                const int nsplit = (max_m + kAssignSlice - 1) / kAssignSlice;
                gms_assign_cnt_kernel<<<dim3(nsplit, 1, cn), 1024, 0, st>>>(d_pairs + c0, d_results + c0, scratch, L, lidx, ridx,
                                                                          cbase, cm);
                ++launches; kmark("gms_assign_cnt", st);
            }
            for (int s = 0; s < n_scales; ++s) {
                const int b = smem_band_rows(s);
                const int bands = (kGridL + b - 1) / b;
                const dim3 grid(bands, 4, cn);
                const size_t sm = smem_band_bytes(s, 4);
                // one CTA per SM (band needs > half of the shared memory): 1024 threads hide the index-load latency better;
                // two CTAs per SM: 512 threads each
                static const int vthreads = getenv("SFMGMS_GMS_THREADS") ? atoi(getenv("SFMGMS_GMS_THREADS")) : 0;   // tuning only
                const int nthreads = vthreads ? (vthreads == 1024 ? 1024 : 512) : (sm > 113 * 1024 ? 1024 : 512);
#define SFMGMS_VOTE2(S) gms_vote2_kernel<S><<<grid, nthreads, sm, st>>>(d_pairs + c0, scratch, L, factor, lidx, ridx, cbase, cm, b)
                if (s == 0) SFMGMS_VOTE2(0); else if (s == 1) SFMGMS_VOTE2(1); else if (s == 2) SFMGMS_VOTE2(2);
                else if (s == 3) SFMGMS_VOTE2(3); else SFMGMS_VOTE2(4);
#undef SFMGMS_VOTE2
                ++launches; kmark("gms_vote2", st);
            }
        }
        if (max_m > 0 && flags_on) {
            int cbx = (max_m + 8191) / 8192;             // a CTA first stages 25.6 KB of tables: give it >= 8192 matches
            cbx = cbx > 32 ? 32 : cbx;
            gms_count_scale_kernel<<<dim3(cbx, n_scales, cn), 256, (size_t)kNumRot * 4 * kCellsL * 2, st>>>(d_pairs + c0, scratch, L, lidx,
                                                                                                      ridx, cbase, cm);
            ++launches; kmark("gms_count_scale", st);
        } else if (max_m > 0) {
            gms_count_kernel<<<dim3(bx, 1, cn), 256, 0, st>>>(d_pairs + c0, scratch, L, lidx, ridx, cbase, cm, 1);
            ++launches; kmark("gms_count", st);
        }
        gms_select_kernel<<<(cn + 127) / 128, 128, 0, st>>>(d_pairs + c0, d_results + c0, scratch, L, cn, flags_on);
        ++launches; kmark("gms_select", st);
        if (flags_on && max_m > 0) {
            gms_mask_kernel<<<dim3(bx, 1, cn), 256, 0, st>>>(d_pairs + c0, d_results + c0, scratch, L, lidx, ridx, cbase, cm);
            ++launches; kmark("gms_mask", st);
        }
    }
    return launches;
}

}  // namespace sfmgms
