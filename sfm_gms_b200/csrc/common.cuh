// common.cuh — shared device/host definitions for the BF-Hamming + GMS path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfmgms {

constexpr int kDescBytes = 32;          // 256-bit ORB descriptor
constexpr int kDescWords = 8;           // as uint32
constexpr int kTrainIdxBits = 18;       // OpenCV: train rows < IMGIDX_ONE = 2^18
constexpr uint32_t kTrainIdxMask = (1u << kTrainIdxBits) - 1u;
constexpr uint32_t kKeyInit = 0xFFFFFFFFu;  // > any (dist<<18 | idx) with dist <= 256

constexpr int kGridL = 20;              // left grid 20x20 (DLL @VA 0x180046ac6)
constexpr int kCellsL = kGridL * kGridL;
constexpr int kNumScales = 5;
constexpr int kNumRot = 8;
constexpr int kMaxHyp = kNumScales * kNumRot;
constexpr uint16_t kNoCell = 0xFFFFu;

// right grid width per scale index: cvRound(20 * {1, .5, 1/sqrt2, sqrt2, 2}) (setScale, DLL @VA 0x180048c10)
__host__ __device__ inline int right_grid_w(int s) {
    return s == 0 ? 20 : s == 1 ? 10 : s == 2 ? 14 : s == 3 ? 28 : 40;
}

// One image pair of a batch, as the kernels see it (device pointers).
struct PairDesc {
    const uint8_t* desc1;   // n1 x 32 (may be null for a GMS-only call)
    const uint8_t* desc2;   // n2 x 32
    const float* kp1;       // n1 x 2 pixel coordinates
    const float* kp2;       // n2 x 2
    const int32_t* mq;      // n_matches queryIdx, or null => queryIdx = i (BFMatcher::match order)
    const int32_t* mt;      // n_matches trainIdx, or null => taken from key[i] & kTrainIdxMask
    uint32_t* key;          // n1 packed (dist << 18 | trainIdx) written by the Hamming kernels
    uint8_t* mask;          // n_matches inlier mask (output)
    int n1, n2, n_matches;
    int w1, h1, w2, h2;
    long long match_base;   // offset of this pair's rows in the batch-concatenated per-match arrays
    int img1, img2;         // image-set indices (-1 for ad-hoc single-pair calls)
    int pair_index;
};

// Per-batch GMS result (one per pair).
struct PairResult {
    int n_inliers;
    int best_hyp;   // scale*8 + rot-1, -1 if none
    int mask_len;   // n_matches or 0
    int status;     // 0 ok, SFMGMS_ERR_DOMAIN, SFMGMS_ERR_INDEX
};

// Per-kernel timing (SFMGMS_OPT_TIMING = 2, measurement runs only): launchers drop a named CUDA event after each
// kernel; the C ABI layer turns consecutive events into per-kernel milliseconds (sfmgms_kernel_times).
struct KernelMarks {
    cudaEvent_t ev[64];
    const char* name[64];
    int created = 0, used = 0;
    void mark(const char* nm, cudaStream_t st) {
        if (used >= 64) return;
        if (used >= created) { if (cudaEventCreate(&ev[created]) != cudaSuccess) return; ++created; }
        cudaEventRecord(ev[used], st);
        name[used++] = nm;
    }
};
extern thread_local KernelMarks* tl_marks;   // set by the C ABI layer around a batch (one context per host thread)
inline void kmark(const char* nm, cudaStream_t st) { if (tl_marks) tl_marks->mark(nm, st); }

// launchers (each returns the number of kernel launches it issued; errors via cudaGetLastError)
int launch_hamming_popc(const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int sm_count,
                        cudaStream_t st);

struct GmsScratch {
    // per pair in chunk, sized by gms_scratch_bytes()
    void* base;
    size_t bytes;
};
size_t gms_scratch_bytes_per_pair(int n_scales);
long long gms_match_rows(const PairDesc* h_pairs, int n);   // rows of the per-match arrays spanned by n pairs
size_t gms_match_scratch_bytes(long long n_matches_total, int n_scales);
int launch_gms(const PairDesc* d_pairs, const PairDesc* h_pairs, int n_pairs, int with_rotation, int with_scale,
               double factor, PairResult* d_results, void* d_hist_scratch, size_t hist_scratch_bytes,
               void* d_match_scratch, cudaStream_t st, int force_dense = 0);

// Number of int32 words from a pair's scratch base to its 40 per-hypothesis inlier counters, for the layout
// launch_gms picks for (n_scales, n_rot, max matches per pair) — sfmgms_gms_hypotheses reads them back.
size_t gms_counts_offset_words(int n_scales, int n_rot, int max_matches, int force_dense);

// (§8f-1 fused) compacted output of a batch: the reference's matchesGMS vectors (cv::DMatch records of the inliers,
// in match order) of all pairs back to back, plus the gathered coordinates SfMUtil.cpp:25-35 builds from them.
// offsets[p] = *base + sum of n_inliers of earlier pairs (exclusive scan, written by the offsets kernel, n_pairs + 1
// entries); *base += total afterwards.  Rows at or beyond `capacity` are not written.  2 launches.
struct DMatchRec { int32_t queryIdx, trainIdx, imgIdx; float distance; };
// index_pairs: d_matches receives 8-byte {queryIdx, trainIdx} records instead of DMatchRec (SFMGMS_OPT_COMPACT_RECORD)
int launch_gms_compact(const PairDesc* d_pairs, const PairResult* d_results, int n_pairs, long long* d_base,
                       long long* d_offsets, long long capacity, DMatchRec* d_matches, float* d_pts1, float* d_pts2,
                       cudaStream_t st, bool index_pairs = false);

// L2 brute force for integer-valued float descriptors (OpenCV SIFT), l2_dp4a.cu
size_t l2_scratch_bytes(int nq, int nt);
int launch_l2_dp4a(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, int32_t* d_train_idx,
                   float* d_dist, int* d_bad, int sm_count, cudaStream_t st);
// cv::ORB (detectAndCompute / compute), orb.cu.  The workspace owns the pyramid and every scratch buffer.
struct OrbWorkspace;
OrbWorkspace* orb_ws_create();
void orb_ws_destroy(OrbWorkspace* w);
const char* orb_ws_error(const OrbWorkspace* w);
int orb_compute_provided(OrbWorkspace* ws, const uint8_t* h_image, int w, int h, int channels, int stride, const float* h_xyao,
                         int n, int nlevels, uint8_t* h_desc, int sm_count, cudaStream_t st, int* launches);
int orb_detect_and_compute(OrbWorkspace* ws, const uint8_t* h_image, int w, int h, int channels, int stride, int nfeatures,
                           int fast_threshold, int nlevels, float scale_factor, int edge, int score_type, int wta_k, int patch,
                           void* h_kp, uint8_t* h_desc, int capacity, int* needed, int sm_count, cudaStream_t st, int* launches);
// order-exact fp32 kernel for general float descriptors, l2_f32.cu (dim in [1, l2_f32_max_dim()])
size_t l2_f32_scratch_bytes(int nq);
int l2_f32_max_dim();
int launch_l2_f32(const float* d_q, int nq, const float* d_t, int nt, int dim, void* d_scratch, int32_t* d_train_idx,
                  float* d_dist, cudaStream_t st);
// tcgen05 (kind::i8, unsigned) variant, l2_tc.cu — same contract; returns -1 on a setup error
size_t l2_tc_scratch_bytes(int nq, int nt);
int launch_l2_tc(const float* d_q, int nq, const float* d_t, int nt, void* d_scratch, int32_t* d_train_idx,
                 float* d_dist, int* d_bad, int sm_count, cudaStream_t st);

}  // namespace sfmgms
