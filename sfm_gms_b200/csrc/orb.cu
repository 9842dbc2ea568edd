// orb.cu — cv::ORB::compute on provided level-0 keypoints (first half): the descriptor side of ORB,
// as DisparityUtil.cpp:127-134 runs it at every pixel (`ORB::create()` defaults, DisparityUtil.cpp:107).
//
// Restates OpenCV features2d orb.cpp (detectAndCompute with useProvidedKeypoints = true; un-vendored dependency):
//   gray = (B*3735 + G*19235 + R*9798 + 2^14) >> 15                      (cvtColor BGR2GRAY, 8-bit)
//   blurred = GaussianBlur(gray, 7x7, sigma 2, reflect-101): Gaussian-weighted sum rounded to nearest (the pyramid
//             sub-matrix takes OpenCV's floating-point path, not its 8-bit fixed-point one; oracle/orb.py has the pin)
//   per keypoint: a = (float)cos(angle_rad), b = (float)sin(angle_rad); 512 pattern points rotated in float with
//             separately rounded products, cvRound (half to even), 256 pixel comparisons -> 32 bytes.
// The border filter (KeyPointsFilter::runByImageBorder, edgeThreshold 31) is the caller's host loop in capi.cu, as it
// is a host loop in OpenCV.  HBM-bound byte work: one pass over the image per stage, descriptors gather from L1/L2.
#include "common.cuh"

namespace sfmgms {

namespace {

__constant__ signed char c_pattern[512][2] = {
#include "orb_pattern.inc"
};

__global__ void orb_gray_kernel(const uint8_t* __restrict__ bgr, int stride, int w, int h, uint8_t* __restrict__ gray) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = bgr + (size_t)y * stride + 3 * x;
    gray[(size_t)y * w + x] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + (1 << 14)) >> 15);
}

__device__ __forceinline__ int reflect101(int i, int n) {   // BORDER_REFLECT_101, |offset| <= 3
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

constexpr int BW = 32, BH = 16, R = 3;

// exp(-x^2 / 8) / sum, x = -3..3 (cv::getGaussianKernel(7, 2.0)), summed in the oracle's order
__global__ void __launch_bounds__(BW * BH) orb_blur_kernel(const uint8_t* __restrict__ src, int stride, int w, int h,
                                                           uint8_t* __restrict__ dst, double k0, double k1, double k2,
                                                           double k3) {
    __shared__ uint8_t tile[BH + 2 * R][BW + 2 * R];
    __shared__ double rows[BH + 2 * R][BW];
    const double k[7] = {k0, k1, k2, k3, k2, k1, k0};
    const int x0 = blockIdx.x * BW, y0 = blockIdx.y * BH;
    const int tid = threadIdx.y * BW + threadIdx.x;
    for (int i = tid; i < (BH + 2 * R) * (BW + 2 * R); i += BW * BH) {
        const int ty = i / (BW + 2 * R), tx = i - ty * (BW + 2 * R);
        tile[ty][tx] = src[(size_t)reflect101(y0 + ty - R, h) * stride + reflect101(x0 + tx - R, w)];
    }
    __syncthreads();
    for (int i = tid; i < (BH + 2 * R) * BW; i += BW * BH) {
        const int ty = i / BW, tx = i - ty * BW;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) s = __dadd_rn(s, __dmul_rn((double)tile[ty][tx + j], k[j]));
        rows[ty][tx] = s;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x < w && y < h) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) s = __dadd_rn(s, __dmul_rn(rows[threadIdx.y + j][threadIdx.x], k[j]));
        const double r = rint(s);                              // half to even, as cvRound / saturate_cast<uchar>
        dst[(size_t)y * w + x] = (uint8_t)(r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r));
    }
}

struct OrbKp {    // prepared keypoint: patch centre and rotation
    int cx, cy;
    float a, b;
};

__global__ void orb_prepare_kernel(const float* __restrict__ xya, int n, OrbKp* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = xya[3 * i], y = xya[3 * i + 1];
    const float rad = __fmul_rn(xya[3 * i + 2], (float)(3.1415926535897932384626433832795 / 180.f));
    OrbKp k;
    k.cx = __float2int_rn(x);
    k.cy = __float2int_rn(y);
    k.a = (float)cos((double)rad);
    k.b = (float)sin((double)rad);
    out[i] = k;
}

// one warp per keypoint (grid-stride); lane = descriptor byte, its 16 pattern points stay in registers
__global__ void __launch_bounds__(256) orb_desc_kernel(const uint8_t* __restrict__ img, int w, const OrbKp* __restrict__ kps,
                                                       int n, uint8_t* __restrict__ desc) {
    const int lane = threadIdx.x & 31;
    float px[16], py[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { px[j] = (float)c_pattern[16 * lane + j][0]; py[j] = (float)c_pattern[16 * lane + j][1]; }
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const OrbKp k = kps[i];
        const uint8_t* centre = img + (size_t)k.cy * w + k.cx;
        int v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float x = __fsub_rn(__fmul_rn(px[j], k.a), __fmul_rn(py[j], k.b));
            const float y = __fadd_rn(__fmul_rn(px[j], k.b), __fmul_rn(py[j], k.a));
            v[j] = __ldg(centre + __float2int_rn(y) * w + __float2int_rn(x));
        }
        unsigned val = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) val |= (unsigned)(v[2 * b] < v[2 * b + 1]) << b;
        desc[(size_t)i * 32 + lane] = (uint8_t)val;
    }
}

}  // namespace

size_t orb_kp_bytes() { return sizeof(OrbKp); }

// image on the device (channels 1 or 3, row stride `stride`); d_xya: n x (x, y, angle_deg) already border-filtered;
// d_gray / d_blur: w*h bytes each; d_prep: n * orb_kp_bytes().  Returns the number of kernel launches.
int launch_orb_compute(const uint8_t* d_image, int w, int h, int channels, int stride, const float* d_xya, int n,
                       uint8_t* d_gray, uint8_t* d_blur, void* d_prep, uint8_t* d_desc, int sm_count, cudaStream_t st) {
    int launches = 0;
    const uint8_t* gray = d_image;
    int gstride = stride;
    if (channels == 3) {
        orb_gray_kernel<<<dim3((unsigned)((w + 255) / 256), (unsigned)h), 256, 0, st>>>(d_image, stride, w, h, d_gray);
        gray = d_gray; gstride = w; ++launches;
    }
    double k[7], sum = 0.0;
    for (int i = 0; i < 7; ++i) { const double x = i - 3.0; k[i] = exp(-(x * x) / 8.0); sum += k[i]; }
    for (int i = 0; i < 7; ++i) k[i] /= sum;
    orb_blur_kernel<<<dim3((unsigned)((w + BW - 1) / BW), (unsigned)((h + BH - 1) / BH)), dim3(BW, BH), 0, st>>>(
        gray, gstride, w, h, d_blur, k[0], k[1], k[2], k[3]);
    ++launches;
    if (n > 0) {
        OrbKp* prep = static_cast<OrbKp*>(d_prep);
        orb_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_xya, n, prep);
        long long blocks = ((long long)n + 7) / 8;
        if (blocks > (long long)sm_count * 16) blocks = (long long)sm_count * 16;
        orb_desc_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_blur, w, prep, n, d_desc);
        launches += 2;
    }
    return launches;
}

}  // namespace sfmgms
