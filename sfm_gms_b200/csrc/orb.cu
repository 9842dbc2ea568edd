// orb.cu — cv::ORB: `ORB::create()` + detectAndCompute / compute as DisparityUtil.cpp:107, 127-140 use
// them.  Restates OpenCV features2d orb.cpp / fast.cpp / keypoint.cpp and imgproc resize (un-vendored dependency of the
// reference; every step is pinned bit-exactly against cv2 4.13 by the CPU restatement in the test tree and tests/golden/orb_*.npz):
//   gray     = (B*3735 + G*19235 + R*9798 + 2^14) >> 15                          cvtColor(BGR2GRAY), 8-bit
//   pyramid  : level l has scale s_l = (float)pow((double)1.2f, l), size cvRound(w / s_l) x cvRound(h / s_l), and is
//              resize(level l-1, INTER_LINEAR_EXACT): 8.8 fixed-point weights round((frac)*256) from
//              f = (1/(dst/src))*(d + 0.5) - 0.5 in double, result (h0*(256-cy) + h1*cy + 2^15) >> 16
//   FAST     : FAST-9/16, threshold t, score = largest t' keeping the pixel a corner, 3x3 strict non-max suppression,
//              row-major order; KeyPointsFilter::runByImageBorder(31); retainBest(2 n_l) by FAST score
//   Harris   : 7x7 block of 3x3 Sobel-like sums (int), response = ((float)a*b - (float)c*c - 0.04f*(a+b)^2) * scale^4
//              with every float op rounded separately; retainBest(n_l) per level
//   retainBest = std::nth_element + std::partition on the host, exactly as OpenCV does it: the ORDER of the returned
//              keypoints is the order those two library calls leave (libstdc++ here and in the stock cv2 packages)
//   angle    : intensity centroid over the radius-15 disc (integer moments), cv::fastAtan2's float polynomial
//   blur     : 7x7 sigma-2 Gaussian in float32 with OpenCV's own tap order and FMA placement (the pyramid SUB-matrix
//              takes the float separable filter, not the 8-bit fixed-point kernel; pinned on 36 M descriptor bits)
//   BRIEF    : a = (float)cos(angle), b = (float)sin(angle); 512 pattern points rotated in float with separately
//              rounded products, cvRound, 256 comparisons -> 32 bytes
// HBM/L2-bound byte and integer work: one coalesced pass over the pyramid per stage; the per-keypoint stages gather.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace sfmgms {

namespace {

constexpr int kMaxLevels = 16;
constexpr int kMaxHalfPatch = 31;    // patchSize <= 63 (ORB::create default 31)

struct Level {
    long long off;   // byte offset of the level image in the pyramid buffers (dense rows, stride = w)
    int w, h;
    float scale;     // layerScale
    int row0;        // first row of this level in the concatenated row index
};
struct LevelTable {
    Level l[kMaxLevels];
    int n;
    int total_rows;
    int max_w;
    int edge;        // edgeThreshold (ORB::create default 31)
};

// bit_pattern_31_: the learned test pairs for patchSize 31 (host copy; the device reads the pattern of the current
// configuration from the workspace)
const signed char h_pattern31[512][2] = {
#include "orb_pattern.inc"
};

struct OrbCfg {      // per-call configuration of the keypoint kernels
    int patch, half;                  // patchSize, patchSize / 2
    int umax[kMaxHalfPatch + 2];      // end of each row of the radius-`half` disc (orb.cpp: umax)
};

__device__ __forceinline__ int reflect101(int i, int n) {   // BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}
// pixel of a level image; outside the image: what OpenCV's pyramid buffer holds there (reflect-101 border)
template <bool kInside>
__device__ __forceinline__ int pix(const uint8_t* img, int w, int h, int x, int y) {
    if (kInside) return img[(size_t)y * w + x];
    return img[(size_t)reflect101(y, h) * w + reflect101(x, w)];
}

__device__ __forceinline__ int level_of_row(const LevelTable& T, int row) {
    int l = 0;
    while (l + 1 < T.n && row >= T.l[l + 1].row0) ++l;
    return l;
}

__global__ void orb_gray_kernel(const uint8_t* __restrict__ bgr, int stride, int w, int h, uint8_t* __restrict__ gray) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = bgr + (size_t)y * stride + 3 * x;
    gray[(size_t)y * w + x] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
}

__global__ void orb_copy_kernel(const uint8_t* __restrict__ src, int stride, int w, int h, uint8_t* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w) dst[(size_t)y * w + x] = src[(size_t)y * stride + x];
}

// resize(INTER_LINEAR_EXACT), 8-bit: xo/xc/yo/yc = source offset and 8-bit weight of the second tap per column / row
__global__ void orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh,
                                  const int* __restrict__ xo, const int* __restrict__ xc, const int* __restrict__ yo,
                                  const int* __restrict__ yc) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    const int x0 = xo[x], x1 = min(x0 + 1, sw - 1), cx = xc[x];
    const int y0 = yo[y], y1 = min(y0 + 1, sh - 1), cy = yc[y];
    const uint8_t* r0 = src + (size_t)y0 * sw;
    const uint8_t* r1 = src + (size_t)y1 * sw;
    const int h0 = r0[x0] * (256 - cx) + r0[x1] * cx;          // 8.8 fixed point
    const int h1 = r1[x0] * (256 - cx) + r1[x1] * cx;
    dst[(size_t)y * dw + x] = (uint8_t)((h0 * (256 - cy) + h1 * cy + (1 << 15)) >> 16);
}

// FAST-9/16 score of every pixel of every level (0 = not a corner).  One thread per pixel, rows of all levels
// concatenated in blockIdx.y.
__global__ void __launch_bounds__(128) orb_fast_score_kernel(LevelTable T, const uint8_t* __restrict__ pyr,
                                                             uint8_t* __restrict__ score, int thr) {
    const int row = blockIdx.y;
    const int l = level_of_row(T, row);
    const int w = T.l[l].w, h = T.l[l].h;
    const int y = row - T.l[l].row0, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    uint8_t* out = score + T.l[l].off + (size_t)y * w + x;
    if (x < 3 || x >= w - 3 || y < 3 || y >= h - 3) { *out = 0; return; }
    const uint8_t* p = pyr + T.l[l].off + (size_t)y * w + x;
    const int v = p[0];
    // the 16-pixel Bresenham circle of radius 3, clockwise from (0, 3) as fast.cpp's makeOffsets lists it
    const int d[16] = {v - p[3 * w],      v - p[3 * w + 1],  v - p[2 * w + 2],  v - p[w + 3],      v - p[3],          v - p[-w + 3],
                       v - p[-2 * w + 2], v - p[-3 * w + 1], v - p[-3 * w],     v - p[-3 * w - 1], v - p[-2 * w - 2], v - p[-w - 3],
                       v - p[-3],         v - p[w - 3],      v - p[2 * w - 2],  v - p[3 * w - 1]};
    int amin = -1000, amax = 1000;       // max over the 16 arcs of 9 of min(d);  min over arcs of max(d)
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
#pragma unroll
        for (int j = 1; j < 9; ++j) { mn = min(mn, d[(s + j) & 15]); mx = max(mx, d[(s + j) & 15]); }
        amin = max(amin, mn);
        amax = min(amax, mx);
    }
    const bool corner = amin > thr || amax < -thr;
    *out = corner ? (uint8_t)(max(max(amin, thr), -min(amax, -thr)) - 1) : (uint8_t)0;
}

__device__ __forceinline__ bool nms_keep(const uint8_t* s, int w, int h, int x, int y, int edge) {
    if (x < edge || x >= w - edge || y < edge || y >= h - edge) return false;      // runByImageBorder
    const uint8_t* p = s + (size_t)y * w + x;
    const int v = p[0];
    return v > 0 && v > p[-1] && v > p[1] && v > p[-w - 1] && v > p[-w] && v > p[-w + 1] && v > p[w - 1] && v > p[w] &&
           v > p[w + 1];
}

__global__ void __launch_bounds__(256) orb_nms_count_kernel(LevelTable T, const uint8_t* __restrict__ score,
                                                            int* __restrict__ rowcount) {
    const int row = blockIdx.x;
    const int l = level_of_row(T, row);
    const int w = T.l[l].w, h = T.l[l].h, y = row - T.l[l].row0;
    const uint8_t* s = score + T.l[l].off;
    int c = 0;
    for (int x = threadIdx.x; x < w; x += 256) c += nms_keep(s, w, h, x, y, T.edge) ? 1 : 0;
    __shared__ int tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if (c) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) rowcount[row] = tot;
}

// exclusive scan of rowcount[n] -> rowoff[n], total -> rowoff[n]; single CTA
__global__ void __launch_bounds__(1024) orb_scan_kernel(const int* __restrict__ cnt, int n, int* __restrict__ off) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int s = wsum[lane], si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
            wsum[lane] = si - s;
        }
        __syncthreads();
        const int c0 = carry;
        if (i < n) off[i] = c0 + wsum[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + wsum[wid] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = carry;
}

struct Cand { int x, y, score, level; };

// corners that survive NMS + border filter, in row-major order per level (levels in order): FAST's output order
__global__ void __launch_bounds__(256) orb_nms_write_kernel(LevelTable T, const uint8_t* __restrict__ score,
                                                            const int* __restrict__ rowoff, Cand* __restrict__ out) {
    const int row = blockIdx.x;
    if (rowoff[row + 1] == rowoff[row]) return;
    const int l = level_of_row(T, row);
    const int w = T.l[l].w, h = T.l[l].h, y = row - T.l[l].row0;
    const uint8_t* s = score + T.l[l].off;
    __shared__ int wcnt[8];
    __shared__ int base;
    if (threadIdx.x == 0) base = rowoff[row];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int x0 = 0; x0 < w; x0 += 256) {
        const int x = x0 + threadIdx.x;
        const bool k = x < w && nms_keep(s, w, h, x, y, T.edge);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wcnt[wid] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { if (j < wid) before += wcnt[j]; total += wcnt[j]; }
        if (k) out[base + before + __popc(m & ((1u << lane) - 1))] = Cand{x, y, (int)s[(size_t)y * w + x], l};
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
}

struct Pt { int x, y, level; };

// HarrisResponses (orb.cpp), blockSize 7, k = 0.04f: one warp per keypoint
template <bool kInside>
__device__ __forceinline__ void harris_sums(const uint8_t* img, int w, int h, int px0, int py0, int lane, int& a, int& b, int& c) {
    for (int k = lane; k < 49; k += 32) {
        const int x = px0 - 3 + k % 7, y = py0 - 3 + k / 7;
        const int nw = pix<kInside>(img, w, h, x - 1, y - 1), nn = pix<kInside>(img, w, h, x, y - 1), ne = pix<kInside>(img, w, h, x + 1, y - 1);
        const int ww = pix<kInside>(img, w, h, x - 1, y), ee = pix<kInside>(img, w, h, x + 1, y);
        const int sw = pix<kInside>(img, w, h, x - 1, y + 1), ss = pix<kInside>(img, w, h, x, y + 1), se = pix<kInside>(img, w, h, x + 1, y + 1);
        const int Ix = (ee - ww) * 2 + (ne - nw) + (se - sw);
        const int Iy = (ss - nn) * 2 + (sw - nw) + (se - ne);
        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
}

__global__ void __launch_bounds__(256) orb_harris_kernel(LevelTable T, const uint8_t* __restrict__ pyr, const Pt* __restrict__ pts,
                                                         int n, float* __restrict__ resp) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const Pt p = pts[i];
    const int w = T.l[p.level].w, h = T.l[p.level].h;
    const uint8_t* img = pyr + T.l[p.level].off;
    int a = 0, b = 0, c = 0;
    if (p.x >= 4 && p.x < w - 4 && p.y >= 4 && p.y < h - 4) harris_sums<true>(img, w, h, p.x, p.y, lane, a, b, c);
    else harris_sums<false>(img, w, h, p.x, p.y, lane, a, b, c);      // edgeThreshold < 4: the block leaves the level
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
        const float scale = __fdiv_rn(1.f, __fmul_rn((float)(4 * 7), 255.f));
        const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
        const float fa = (float)a, fb = (float)b, fc = (float)c;
        const float t = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
        const float u = __fadd_rn(fa, fb);
        resp[i] = __fmul_rn(__fsub_rn(t, __fmul_rn(__fmul_rn(0.04f, u), u)), s4);
    }
}

// cv::fastAtan2 (mathfuncs_core): degrees in [0, 360)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float p1 = 57.283626556396484f, p3 = -18.66744613647461f, p5 = 8.914000511169434f, p7 = -2.539724588394165f;
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    float a;
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

struct KpOut {      // cv::KeyPoint layout (28 bytes)
    float x, y, size, angle, response;
    int octave, class_id;
};
struct OrbKp {      // prepared keypoint for the descriptor kernel
    int cx, cy, level;
    float a, b;
};

// ICAngles + the final scaling of computeKeyPoints; also prepares the descriptor centre exactly as
// computeOrbDescriptors recomputes it from the scaled point.  One warp per keypoint, lane v = disc rows +-v.
template <bool kInside>
__device__ __forceinline__ void ic_moments(const uint8_t* img, int w, int h, int x0, int y0, int v, const OrbCfg& cfg, int& m10, int& m01) {
    if (v == 0) {
        for (int u = -cfg.half; u <= cfg.half; ++u) m10 += u * pix<kInside>(img, w, h, x0 + u, y0);
    } else if (v <= cfg.half) {
        const int d = cfg.umax[v];
        int vs = 0;
        for (int u = -d; u <= d; ++u) {
            const int plus = pix<kInside>(img, w, h, x0 + u, y0 + v), minus = pix<kInside>(img, w, h, x0 + u, y0 - v);
            vs += plus - minus;
            m10 += u * (plus + minus);
        }
        m01 = v * vs;
    }
}

__global__ void __launch_bounds__(256) orb_angle_kernel(LevelTable T, OrbCfg cfg, const uint8_t* __restrict__ pyr,
                                                        const Pt* __restrict__ pts, const float* __restrict__ resp, int n,
                                                        KpOut* __restrict__ out, OrbKp* __restrict__ prep) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, v = threadIdx.x & 31;
    if (i >= n) return;                        // whole warps leave together
    const Pt p = pts[i];
    const int w = T.l[p.level].w, h = T.l[p.level].h;
    const uint8_t* img = pyr + T.l[p.level].off;
    int m10 = 0, m01 = 0;
    if (p.x >= cfg.half && p.x < w - cfg.half && p.y >= cfg.half && p.y < h - cfg.half) ic_moments<true>(img, w, h, p.x, p.y, v, cfg, m10, m01);
    else ic_moments<false>(img, w, h, p.x, p.y, v, cfg, m10, m01);      // small edgeThreshold: the disc leaves the level
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    if (v == 0) {
        const float sc = T.l[p.level].scale;
        KpOut k;
        k.x = __fmul_rn((float)p.x, sc);
        k.y = __fmul_rn((float)p.y, sc);
        k.size = __fmul_rn((float)cfg.patch, sc);
        k.angle = fast_atan2_deg((float)m01, (float)m10);
        k.response = resp[i];
        k.octave = p.level;
        k.class_id = -1;
        out[i] = k;
        const float inv = __fdiv_rn(1.f, sc);
        const float rad = __fmul_rn(k.angle, (float)(3.1415926535897932384626433832795 / 180.f));
        OrbKp q;
        q.cx = __float2int_rn(__fmul_rn(k.x, inv));
        q.cy = __float2int_rn(__fmul_rn(k.y, inv));
        q.level = p.level;
        q.a = (float)cos((double)rad);
        q.b = (float)sin((double)rad);
        prep[i] = q;
    }
}

// provided keypoints (x, y, angle_deg, octave) -> descriptor centre and rotation
__global__ void orb_prepare_kernel(LevelTable T, const float* __restrict__ xyao, int n, OrbKp* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int level = __float_as_int(xyao[4 * i + 3]);
    const float inv = __fdiv_rn(1.f, T.l[level].scale);
    const float rad = __fmul_rn(xyao[4 * i + 2], (float)(3.1415926535897932384626433832795 / 180.f));
    OrbKp k;
    k.cx = __float2int_rn(__fmul_rn(xyao[4 * i], inv));
    k.cy = __float2int_rn(__fmul_rn(xyao[4 * i + 1], inv));
    k.level = level;
    k.a = (float)cos((double)rad);
    k.b = (float)sin((double)rad);
    out[i] = k;
}

constexpr int BW = 32, BH = 16, R = 3;

// 7x7 sigma-2 Gaussian of one level, exactly as OpenCV's float separable filter evaluates it for ORB (the pyramid
// sub-matrix does not take the 8-bit fixed-point path): float taps (float)getGaussianKernel(7, 2); row pass
// s = x0*k0, s = fma(x_i, k_i, s) in the first 32*floor(w/32) columns and s = s + x_i*k_i (two roundings) in the
// rest; column pass s = r3*k3, s = fma(r[3-d] + r[3+d], k[3-d], s) (two roundings in the last w mod 4 columns);
// round half to even.
__global__ void __launch_bounds__(BW * BH) orb_blur_kernel(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst) {
    __shared__ uint8_t tile[BH + 2 * R][BW + 2 * R];
    __shared__ float rows[BH + 2 * R][BW];
    const float k[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f, 0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    const int row_vec_end = w & ~31;
    const int tid = threadIdx.y * BW + threadIdx.x;
    for (int i = tid; i < (BH + 2 * R) * (BW + 2 * R); i += BW * BH) {
        const int ty = i / (BW + 2 * R), tx = i - ty * (BW + 2 * R);
        tile[ty][tx] = src[(size_t)reflect101(y0 + ty - R, h) * w + reflect101(x0 + tx - R, w)];
    }
    __syncthreads();
    for (int i = tid; i < (BH + 2 * R) * BW; i += BW * BH) {
        const int ty = i / BW, tx = i - ty * BW;
        float s = __fmul_rn((float)tile[ty][tx], k[0]);
        if (x0 + tx < row_vec_end) {     // OpenCV's 32-pixel vector loop of the uchar->float row filter: fused
#pragma unroll
            for (int j = 1; j < 7; ++j) s = __fmaf_rn((float)tile[ty][tx + j], k[j], s);
        } else {                         // its scalar remainder: product and sum rounded separately
#pragma unroll
            for (int j = 1; j < 7; ++j) s = __fadd_rn(s, __fmul_rn((float)tile[ty][tx + j], k[j]));
        }
        rows[ty][tx] = s;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < w && y < h) {
        float s = __fmul_rn(rows[threadIdx.y + 3][threadIdx.x], k[3]);
        if (x < (w & ~3)) {              // the vector loops of the column filter (8 and 4 lanes): fused
#pragma unroll
            for (int d = 1; d <= 3; ++d)
                s = __fmaf_rn(__fadd_rn(rows[threadIdx.y + 3 - d][threadIdx.x], rows[threadIdx.y + 3 + d][threadIdx.x]), k[3 - d], s);
        } else {                         // its scalar tail (last w mod 4 columns): two roundings
#pragma unroll
            for (int d = 1; d <= 3; ++d)
                s = __fadd_rn(s, __fmul_rn(__fadd_rn(rows[threadIdx.y + 3 - d][threadIdx.x], rows[threadIdx.y + 3 + d][threadIdx.x]), k[3 - d]));
        }
        const float r = rintf(s);                              // half to even, as saturate_cast<uchar>(float)
        dst[(size_t)y * w + x] = (uint8_t)(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
    }
}

// one warp per keypoint (grid-stride); lane = descriptor byte, its pattern points (16 for WTA_K 2 and 4, 12 for
// WTA_K 3) stay in registers.  A sample outside the level image (small edgeThreshold, or caller-provided keypoints on
// coarse levels) reads what OpenCV's pyramid holds there: the reflect-101 border of the UNBLURRED level.
template <int kWta>
__global__ void __launch_bounds__(256) orb_desc_kernel(LevelTable T, const signed char* __restrict__ pattern, const uint8_t* __restrict__ blur,
                                                       const uint8_t* __restrict__ raw, const OrbKp* __restrict__ kps, int n,
                                                       uint8_t* __restrict__ desc) {
    constexpr int kPts = kWta == 3 ? 12 : 16;
    const int lane = threadIdx.x & 31;
    float px[kPts], py[kPts];
#pragma unroll
    for (int j = 0; j < kPts; ++j) { px[j] = (float)pattern[2 * (kPts * lane + j)]; py[j] = (float)pattern[2 * (kPts * lane + j) + 1]; }
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const OrbKp k = kps[i];
        const int w = T.l[k.level].w, h = T.l[k.level].h;
        const uint8_t* img = blur + T.l[k.level].off;
        int v[kPts];
#pragma unroll
        for (int j = 0; j < kPts; ++j) {
            const float x = __fsub_rn(__fmul_rn(px[j], k.a), __fmul_rn(py[j], k.b));
            const float y = __fadd_rn(__fmul_rn(px[j], k.b), __fmul_rn(py[j], k.a));
            const int xx = k.cx + __float2int_rn(x), yy = k.cy + __float2int_rn(y);
            if (xx >= 0 && xx < w && yy >= 0 && yy < h) v[j] = __ldg(img + (size_t)yy * w + xx);
            else v[j] = raw[T.l[k.level].off + (size_t)reflect101(yy, h) * w + reflect101(xx, w)];
        }
        unsigned val = 0;
        if (kWta == 2) {
#pragma unroll
            for (int b = 0; b < 8; ++b) val |= (unsigned)(v[2 * b] < v[2 * b + 1]) << b;
        } else if (kWta == 3) {              // index of the maximum of 3, ties as orb.cpp resolves them
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int t0 = v[3 * g], t1 = v[3 * g + 1], t2 = v[3 * g + 2];
                val |= (unsigned)(t2 > t1 ? (t2 > t0 ? 2 : 0) : (t1 > t0 ? 1 : 0)) << (2 * g);
            }
        } else {                             // index of the maximum of 4
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                int t0 = v[4 * g], t2 = v[4 * g + 2];
                const int t1 = v[4 * g + 1], t3 = v[4 * g + 3];
                int u = 0, vv = 2;
                if (t1 > t0) { t0 = t1; u = 1; }
                if (t3 > t2) { t2 = t3; vv = 3; }
                val |= (unsigned)(t0 > t2 ? u : vv) << (2 * g);
            }
        }
        desc[(size_t)i * 32 + lane] = (uint8_t)val;
    }
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) return false;
        cap = want;
        return true;
    }
    ~Buf() { if (p) cudaFree(p); }
};

// KeyPointsFilter::retainBest (features2d keypoint.cpp): nth_element + partition; the resulting order is the output order
struct Rec { float response; int id; };
void retain_best(std::vector<Rec>& v, int n_points) {
    if (n_points < 0 || v.size() <= (size_t)n_points) return;
    if (n_points == 0) { v.clear(); return; }
    std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const Rec& a, const Rec& b) { return a.response > b.response; });
    const float ambiguous = v[(size_t)n_points - 1].response;
    auto new_end = std::partition(v.begin() + n_points, v.end(), [ambiguous](const Rec& a) { return a.response >= ambiguous; });
    v.resize((size_t)(new_end - v.begin()));
}

// interpolationLinear<uint8_t>::getCoeffs (imgproc resize.cpp, bit-exact path): offset and 8-bit weight per output index
void linear_exact_coeffs(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& c1) {
    ofs.assign((size_t)dsize, 0); c1.assign((size_t)dsize, 0);
    const double inv_scale = (double)dsize / (double)ssize;
    const double scale = 1.0 / inv_scale;
    for (int v = 0; v < dsize; ++v) {
        const double f = scale * ((double)v + 0.5) - 0.5;
        const int i = (int)std::floor(f);
        if (i >= 0 && ssize > 1) {
            if (i < ssize - 1) { ofs[(size_t)v] = i; c1[(size_t)v] = (int)std::nearbyint((f - i) * 256.0); }
            else ofs[(size_t)v] = ssize - 1;      // right border: the last pixel with full weight
        }                                         // left border: pixel 0 with full weight (ofs 0, c1 0)
    }
}

}  // namespace

// SFMGMS_ORB_TRACE=1: wall time between the host synchronisation points of detectAndCompute (stderr)
struct StageTimer {
    bool on = getenv("SFMGMS_ORB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[orb] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

struct OrbWorkspace {
    Buf img, pyr, blur, score, rowcnt, rowoff, cand, pts, resp, kpout, prep, desc, coef, xyao, pattern;
    LevelTable T;
    int pattern_patch = 0, pattern_wta = 0;     // configuration the device pattern was built for
    char err[256] = "";
};

OrbWorkspace* orb_ws_create() { return new OrbWorkspace(); }
void orb_ws_destroy(OrbWorkspace* w) { delete w; }
const char* orb_ws_error(const OrbWorkspace* w) { return w->err; }

namespace {

bool fail_ws(OrbWorkspace* ws, const char* msg) { snprintf(ws->err, sizeof ws->err, "%s", msg); return false; }

// gray level 0 + the scale pyramid (levels 1..n-1) on the device; fills ws->T.  Host image in, any stride.
bool build_pyramid(OrbWorkspace* ws, const uint8_t* h_image, int w, int h, int channels, int stride, int nlevels,
                   double scale_factor, int edge, cudaStream_t st, int* launches) {
    LevelTable& T = ws->T;
    T.n = nlevels;
    T.edge = edge;
    long long off = 0;
    int row = 0, max_w = 0;
    for (int l = 0; l < nlevels; ++l) {
        const float scale = (float)std::pow(scale_factor, (double)l);     // getScale(level, firstLevel = 0, scaleFactor)
        const float inv = 1.0f / scale;
        Level& L = T.l[l];
        L.scale = scale;
        L.w = (int)lrintf((float)w * inv);
        L.h = (int)lrintf((float)h * inv);
        if (L.w < 1 || L.h < 1) return fail_ws(ws, "image too small for the requested pyramid level");
        L.off = off; L.row0 = row;
        off += (long long)L.w * L.h;
        off = (off + 255) & ~255ll;
        row += L.h;
        max_w = std::max(max_w, L.w);
    }
    T.total_rows = row; T.max_w = max_w;
    if (row > 65535) return fail_ws(ws, "image too tall: the pyramid has more than 65535 rows in total");
    const size_t img_bytes = (size_t)stride * h;
    if (!ws->img.ensure(img_bytes) || !ws->pyr.ensure((size_t)off) || !ws->blur.ensure((size_t)off) || !ws->score.ensure((size_t)off))
        return fail_ws(ws, "cudaMalloc failed (pyramid)");
    if (cudaMemcpyAsync(ws->img.p, h_image, img_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail_ws(ws, "H2D failed");
    uint8_t* pyr = (uint8_t*)ws->pyr.p;
    const dim3 g0((unsigned)((w + 255) / 256), (unsigned)h);
    if (channels == 3) orb_gray_kernel<<<g0, 256, 0, st>>>((const uint8_t*)ws->img.p, stride, w, h, pyr);
    else orb_copy_kernel<<<g0, 256, 0, st>>>((const uint8_t*)ws->img.p, stride, w, h, pyr);
    ++*launches;
    if (nlevels > 1) {
        // resize coefficients of every level, one upload
        std::vector<int> all, xo, xc, yo, yc;
        std::vector<size_t> pos((size_t)nlevels * 4, 0);
        for (int l = 1; l < nlevels; ++l) {
            linear_exact_coeffs(T.l[l - 1].w, T.l[l].w, xo, xc);
            linear_exact_coeffs(T.l[l - 1].h, T.l[l].h, yo, yc);
            const std::vector<int>* v[4] = {&xo, &xc, &yo, &yc};
            for (int k = 0; k < 4; ++k) { pos[(size_t)l * 4 + k] = all.size(); all.insert(all.end(), v[k]->begin(), v[k]->end()); }
        }
        if (!ws->coef.ensure(all.size() * 4)) return fail_ws(ws, "cudaMalloc failed (coefficients)");
        // pageable source: the copy is staged before the call returns, so `all` may go out of scope afterwards
        if (cudaMemcpyAsync(ws->coef.p, all.data(), all.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail_ws(ws, "H2D failed");
        const int* c = (const int*)ws->coef.p;
        for (int l = 1; l < nlevels; ++l) {
            const Level& S = T.l[l - 1];
            const Level& D = T.l[l];
            orb_resize_kernel<<<dim3((unsigned)((D.w + 255) / 256), (unsigned)D.h), 256, 0, st>>>(
                pyr + S.off, S.w, S.h, pyr + D.off, D.w, D.h, c + pos[(size_t)l * 4], c + pos[(size_t)l * 4 + 1], c + pos[(size_t)l * 4 + 2],
                c + pos[(size_t)l * 4 + 3]);
            ++*launches;
        }
    }
    return true;
}

void blur_levels(OrbWorkspace* ws, cudaStream_t st, int* launches) {
    for (int l = 0; l < ws->T.n; ++l) {
        const Level& L = ws->T.l[l];
        orb_blur_kernel<<<dim3((unsigned)((L.w + BW - 1) / BW), (unsigned)((L.h + BH - 1) / BH)), dim3(BW, BH), 0, st>>>(
            (const uint8_t*)ws->pyr.p + L.off, L.w, L.h, (uint8_t*)ws->blur.p + L.off);
        ++*launches;
    }
}

// cv::RNG (core/operations.hpp): multiply-with-carry generator; uniform(a, b) = a + next() % (b - a)
struct CvRng {
    unsigned long long state;
    explicit CvRng(unsigned long long s) : state(s ? s : 0xffffffffull) {}
    unsigned next() { state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32); return (unsigned)state; }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// the sample pattern of a configuration, as orb.cpp builds it: bit_pattern_31_ for patchSize 31, else
// makeRandomPattern (RNG 0x34985739); for WTA_K 3 / 4 initializeOrbPattern draws 128 tuples of distinct points from
// that pool (RNG 0x12345678)
bool upload_pattern(OrbWorkspace* ws, int patch, int wta, cudaStream_t st) {
    if (ws->pattern.p && ws->pattern_patch == patch && ws->pattern_wta == wta) return true;
    signed char pool[512][2];
    if (patch == 31) {
        memcpy(pool, h_pattern31, sizeof pool);
    } else {
        CvRng rng(0x34985739ull);
        for (int i = 0; i < 512; ++i) {
            pool[i][0] = (signed char)rng.uniform(-patch / 2, patch / 2 + 1);
            pool[i][1] = (signed char)rng.uniform(-patch / 2, patch / 2 + 1);
        }
    }
    signed char pat[512][2];
    int npts = 512;
    if (wta == 2) {
        memcpy(pat, pool, sizeof pat);
    } else {
        CvRng rng(0x12345678ull);
        const int ntuples = 32 * 4;
        npts = ntuples * wta;
        for (int i = 0; i < ntuples; ++i)
            for (int k = 0; k < wta; ++k)
                for (;;) {
                    const int idx = rng.uniform(0, 512);
                    int k1 = 0;
                    for (; k1 < k; ++k1)
                        if (pat[wta * i + k1][0] == pool[idx][0] && pat[wta * i + k1][1] == pool[idx][1]) break;
                    if (k1 == k) { pat[wta * i + k][0] = pool[idx][0]; pat[wta * i + k][1] = pool[idx][1]; break; }
                }
    }
    if (!ws->pattern.ensure(sizeof pat)) return fail_ws(ws, "cudaMalloc failed (pattern)");
    if (cudaMemcpyAsync(ws->pattern.p, pat, (size_t)npts * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) return fail_ws(ws, "H2D failed");
    ws->pattern_patch = patch; ws->pattern_wta = wta;
    return true;
}

// umax (orb.cpp computeKeyPoints): row ends of the radius-`half` disc, made symmetric
void make_cfg(int patch, OrbCfg& cfg) {
    cfg.patch = patch;
    const int half = patch / 2;
    cfg.half = half;
    for (int& u : cfg.umax) u = 0;
    const int vmax = (int)std::floor(half * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(half * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) cfg.umax[v] = (int)std::nearbyint(std::sqrt((double)half * half - (double)v * v));
    for (int v = half, v0 = 0; v >= vmin; --v) {
        while (cfg.umax[v0] == cfg.umax[v0 + 1]) ++v0;
        cfg.umax[v] = v0;
        ++v0;
    }
}

bool describe(OrbWorkspace* ws, int n, int patch, int wta, uint8_t* h_desc, int sm_count, cudaStream_t st, int* launches) {
    if (!ws->desc.ensure((size_t)n * 32)) return fail_ws(ws, "cudaMalloc failed (descriptors)");
    if (!upload_pattern(ws, patch, wta, st)) return false;
    long long blocks = ((long long)n + 7) / 8;
    if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
    const signed char* pat = (const signed char*)ws->pattern.p;
    const uint8_t* blur = (const uint8_t*)ws->blur.p;
    const uint8_t* raw = (const uint8_t*)ws->pyr.p;
    const OrbKp* prep = (const OrbKp*)ws->prep.p;
    uint8_t* d = (uint8_t*)ws->desc.p;
    if (wta == 2) orb_desc_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(ws->T, pat, blur, raw, prep, n, d);
    else if (wta == 3) orb_desc_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(ws->T, pat, blur, raw, prep, n, d);
    else orb_desc_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(ws->T, pat, blur, raw, prep, n, d);
    ++*launches;
    if (h_desc && cudaMemcpyAsync(h_desc, ws->desc.p, (size_t)n * 32, cudaMemcpyDeviceToHost, st) != cudaSuccess) return fail_ws(ws, "D2H failed");
    return true;
}

}  // namespace

// cv::ORB::compute on provided keypoints.  h_xyao: n records (x, y, angle_deg, octave as int bits), already
// border-filtered and grouped by level by the caller; nlevels = max octave + 1.
int orb_compute_provided(OrbWorkspace* ws, const uint8_t* h_image, int w, int h, int channels, int stride, const float* h_xyao,
                         int n, int nlevels, uint8_t* h_desc, int sm_count, cudaStream_t st, int* launches) {
    if (nlevels < 1 || nlevels > kMaxLevels) { fail_ws(ws, "octave out of range (0..15)"); return -1; }
    if (!build_pyramid(ws, h_image, w, h, channels, stride, nlevels, (double)1.2f, 31, st, launches)) return -1;
    blur_levels(ws, st, launches);
    if (!ws->xyao.ensure((size_t)n * 16) || !ws->prep.ensure((size_t)n * sizeof(OrbKp))) { fail_ws(ws, "cudaMalloc failed"); return -1; }
    if (cudaMemcpyAsync(ws->xyao.p, h_xyao, (size_t)n * 16, cudaMemcpyHostToDevice, st) != cudaSuccess) { fail_ws(ws, "H2D failed"); return -1; }
    orb_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws->T, (const float*)ws->xyao.p, n, (OrbKp*)ws->prep.p);
    ++*launches;
    if (!describe(ws, n, 31, 2, h_desc, sm_count, st, launches)) return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { fail_ws(ws, "CUDA error in ORB compute"); return -1; }
    return n;
}

// cv::ORB::detectAndCompute with ORB::create() defaults except nfeatures / fastThreshold.
// h_kp: capacity records in cv::KeyPoint layout (28 B); h_desc: capacity x 32.  Returns the number of keypoints,
// -1 on error, -2 when capacity is too small (*needed is set).
int orb_detect_and_compute(OrbWorkspace* ws, const uint8_t* h_image, int w, int h, int channels, int stride, int nfeatures,
                           int fast_threshold, int nlevels, float scale_factor_f, int edge, int score_type, int wta_k, int patch,
                           void* h_kp, uint8_t* h_desc, int capacity, int* needed, int sm_count, cudaStream_t st, int* launches) {
    if (patch < 2 || patch > 2 * kMaxHalfPatch + 1 || wta_k < 2 || wta_k > 4) { fail_ws(ws, "patchSize (2..63) / WTA_K (2..4) out of range"); return -1; }
    OrbCfg cfg;
    make_cfg(patch, cfg);
    const double scale_factor = (double)scale_factor_f;     // ORB::create takes a float, the class keeps a double
    if (nlevels < 1 || nlevels > kMaxLevels) { fail_ws(ws, "nlevels out of range (1..16)"); return -1; }
    StageTimer tm;
    if (!build_pyramid(ws, h_image, w, h, channels, stride, nlevels, scale_factor, edge, st, launches)) return -1;
    tm.mark("enqueue H2D + pyramid");
    const LevelTable& T = ws->T;
    // features per level (computeKeyPoints): geometric split in float, remainder to the last level
    std::vector<int> n_per((size_t)nlevels);
    {
        const float factor = (float)(1.0 / scale_factor);
        float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; ++l) {
            n_per[(size_t)l] = (int)lrintf(nd);
            sum += n_per[(size_t)l];
            nd *= factor;
        }
        n_per[(size_t)nlevels - 1] = std::max(nfeatures - sum, 0);
    }
    // FAST scores, NMS, ordered compaction
    const int rows = T.total_rows;
    if (!ws->rowcnt.ensure((size_t)rows * 4) || !ws->rowoff.ensure((size_t)(rows + 1) * 4)) { fail_ws(ws, "cudaMalloc failed"); return -1; }
    orb_fast_score_kernel<<<dim3((unsigned)((T.max_w + 127) / 128), (unsigned)rows), 128, 0, st>>>(T, (const uint8_t*)ws->pyr.p,
                                                                                                  (uint8_t*)ws->score.p, fast_threshold);
    orb_nms_count_kernel<<<rows, 256, 0, st>>>(T, (const uint8_t*)ws->score.p, (int*)ws->rowcnt.p);
    orb_scan_kernel<<<1, 1024, 0, st>>>((const int*)ws->rowcnt.p, rows, (int*)ws->rowoff.p);
    *launches += 3;
    int total = 0;
    if (cudaMemcpyAsync(&total, (const int*)ws->rowoff.p + rows, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { fail_ws(ws, "CUDA error in FAST"); return -1; }
    tm.mark("pyramid + FAST + count (sync)");
    std::vector<Cand> cand((size_t)total);
    if (total > 0) {
        if (!ws->cand.ensure((size_t)total * sizeof(Cand))) { fail_ws(ws, "cudaMalloc failed"); return -1; }
        orb_nms_write_kernel<<<rows, 256, 0, st>>>(T, (const uint8_t*)ws->score.p, (const int*)ws->rowoff.p, (Cand*)ws->cand.p);
        ++*launches;
        if (cudaMemcpyAsync(cand.data(), ws->cand.p, (size_t)total * sizeof(Cand), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { fail_ws(ws, "CUDA error in NMS"); return -1; }
    }
    tm.mark("NMS write + D2H (sync)");
    // per level: retainBest(2 * n_level) by FAST score (HARRIS_SCORE keeps twice as many for the second ranking)
    std::vector<Pt> sel;
    std::vector<float> sel_resp;      // FAST scores of the selection (the final response under FAST_SCORE)
    std::vector<int> counters((size_t)nlevels, 0);
    {
        size_t pos = 0;
        std::vector<Rec> v;
        v.reserve(cand.size());
        sel.reserve(cand.size() < (size_t)4 * (size_t)nfeatures + 64 ? cand.size() : (size_t)4 * (size_t)nfeatures + 64);
        for (int l = 0; l < nlevels; ++l) {
            v.clear();
            const size_t first = pos;
            while (pos < cand.size() && cand[pos].level == l) { v.push_back(Rec{(float)cand[pos].score, (int)(pos - first)}); ++pos; }
            retain_best(v, (score_type == 0 ? 2 : 1) * n_per[(size_t)l]);   // HARRIS_SCORE keeps twice as many for its own ranking
            counters[(size_t)l] = (int)v.size();
            for (const Rec& r : v) { sel.push_back(Pt{cand[first + (size_t)r.id].x, cand[first + (size_t)r.id].y, l}); sel_resp.push_back(r.response); }
        }
    }
    const int nsel = (int)sel.size();
    tm.mark("host retainBest (FAST)");
    if (nsel == 0) { if (needed) *needed = 0; return 0; }
    std::vector<Pt> fin;
    std::vector<float> fin_resp;
    if (!ws->pts.ensure((size_t)nsel * sizeof(Pt)) || !ws->resp.ensure((size_t)nsel * 4)) { fail_ws(ws, "cudaMalloc failed"); return -1; }
    if (score_type == 0) {
        // Harris responses of the selected corners, then retainBest(n_level) per level
        std::vector<float> resp((size_t)nsel);
        if (cudaMemcpyAsync(ws->pts.p, sel.data(), (size_t)nsel * sizeof(Pt), cudaMemcpyHostToDevice, st) != cudaSuccess) { fail_ws(ws, "H2D failed"); return -1; }
        orb_harris_kernel<<<(nsel + 7) / 8, 256, 0, st>>>(T, (const uint8_t*)ws->pyr.p, (const Pt*)ws->pts.p, nsel, (float*)ws->resp.p);
        ++*launches;
        if (cudaMemcpyAsync(resp.data(), ws->resp.p, (size_t)nsel * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { fail_ws(ws, "CUDA error in Harris"); return -1; }
        tm.mark("Harris + D2H (sync)");
        size_t offset = 0;
        std::vector<Rec> v;
        for (int l = 0; l < nlevels; ++l) {
            v.clear();
            for (int i = 0; i < counters[(size_t)l]; ++i) v.push_back(Rec{resp[offset + (size_t)i], i});
            retain_best(v, n_per[(size_t)l]);
            for (const Rec& r : v) { fin.push_back(sel[offset + (size_t)r.id]); fin_resp.push_back(r.response); }
            offset += (size_t)counters[(size_t)l];
        }
    } else {                                   // FAST_SCORE: the first selection is final
        fin = sel;
        fin_resp = sel_resp;
    }
    const int n = (int)fin.size();
    tm.mark("host retainBest (Harris)");
    if (needed) *needed = n;
    if (n > capacity) return -2;
    if (n == 0) return 0;
    // orientation, final coordinates, descriptors
    if (!ws->kpout.ensure((size_t)n * sizeof(KpOut)) || !ws->prep.ensure((size_t)n * sizeof(OrbKp))) { fail_ws(ws, "cudaMalloc failed"); return -1; }
    if (cudaMemcpyAsync(ws->pts.p, fin.data(), (size_t)n * sizeof(Pt), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(ws->resp.p, fin_resp.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { fail_ws(ws, "H2D failed"); return -1; }
    orb_angle_kernel<<<(n + 7) / 8, 256, 0, st>>>(T, cfg, (const uint8_t*)ws->pyr.p, (const Pt*)ws->pts.p, (const float*)ws->resp.p, n,
                                                    (KpOut*)ws->kpout.p, (OrbKp*)ws->prep.p);
    ++*launches;
    if (h_kp && cudaMemcpyAsync(h_kp, ws->kpout.p, (size_t)n * sizeof(KpOut), cudaMemcpyDeviceToHost, st) != cudaSuccess) { fail_ws(ws, "D2H failed"); return -1; }
    if (h_desc) {
        blur_levels(ws, st, launches);
        if (!describe(ws, n, patch, wta_k, h_desc, sm_count, st, launches)) return -1;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { fail_ws(ws, "CUDA error in ORB"); return -1; }
    tm.mark("angle + blur + BRIEF + D2H");
    return n;
}

}  // namespace sfmgms
