/*
 * sfmgms.h — C ABI of the B200-native BF-Hamming + GMS matching stage.
 *
 * Drop-in boundary for the two calls the reference makes back to back
 *     matcherBF->match(desc1, desc2, matches);                         FeatureMatchUtil.cpp:68
 *     cv::xfeatures2d::matchGMS(size1, size2, kp1, kp2, matches, out,  FeatureMatchUtil.cpp:69
 *                               withRotation, withScale[, thresholdFactor]);
 * (also DisparityUtil.cpp:143+149 and :296+299; bruteForceMatch at FeatureMatchUtil.cpp:20-31 for the
 * cross-check variant).  The reference itself has no FFI: both callees are OpenCV C++ methods
 * (SURVEY.md §8b).  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; never throws; returns SFMGMS_OK (0) or an
 *    SFMGMS_ERR_* code, with a human-readable message available from sfmgms_last_error().
 *  - the caller owns every buffer it passes in (as with OpenCV's const& inputs / cleared outputs);
 *    the context owns all device memory it allocates.
 *  - one context per host thread / per GPU; calls on a context are synchronous on return.
 *  - descriptors are 256-bit ORB descriptors: CV_8U rows of desc_bytes == 32, C-contiguous.
 *  - keypoints are read as two consecutive floats (pt.x, pt.y) every `stride_bytes`
 *    (8 for packed xy, 28 for an array of cv::KeyPoint); matches as int32 queryIdx/trainIdx every
 *    `stride_bytes` (4 for plain arrays, 16 for an array of cv::DMatch).
 *  - there is no CPU fallback: without a CUDA device sfmgms_create fails.
 */
#ifndef SFMGMS_H_
#define SFMGMS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFMGMS_OK 0
#define SFMGMS_ERR_ARG 1        /* null/negative/inconsistent argument, desc_bytes != 32 */
#define SFMGMS_ERR_TRAIN_ROWS 2 /* train rows >= 2^18: OpenCV's CV_Assert(rows < IMGIDX_ONE) in BFMatcher */
#define SFMGMS_ERR_DOMAIN 3     /* a matched keypoint lies outside [0,w)x[0,h): undefined behaviour in OpenCV GMS */
#define SFMGMS_ERR_INDEX 4      /* queryIdx/trainIdx out of range: undefined behaviour in OpenCV GMS */
#define SFMGMS_ERR_CUDA 5       /* CUDA runtime error (message in sfmgms_last_error) */
#define SFMGMS_ERR_STATE 6      /* call order violated (e.g. match_pairs before set_images) */
#define SFMGMS_ERR_CAPACITY 7   /* a caller buffer of stated capacity is too small (needed size reported) */
#define SFMGMS_ERR_NCCL 8       /* multi-GPU: NCCL could not be loaded or a collective failed */

#define SFMGMS_MAX_TRAIN_ROWS (1 << 18)

/* where a buffer lives */
#define SFMGMS_HOST 0
#define SFMGMS_DEVICE 1

/* Hamming kernel selection (sfmgms_set_option key SFMGMS_OPT_HAMMING_KERNEL) */
#define SFMGMS_OPT_HAMMING_KERNEL 1
#define SFMGMS_HAMMING_AUTO 0
#define SFMGMS_HAMMING_POPC 1   /* XOR/popc CUDA-core kernel */
#define SFMGMS_HAMMING_TC 2     /* tcgen05 int8 tensor-core kernel (unpacked +-1 bits, TMEM accumulators) */
#define SFMGMS_HAMMING_FP4 3    /* tcgen05 block-scaled FP4 (kind::mxf4) kernel: +-1 as E2M1, unit scales */
#define SFMGMS_OPT_GMS_CHUNK_BYTES 2 /* scratch budget per GMS chunk (bytes), default 64 MiB */
#define SFMGMS_OPT_TC_OPERAND_CACHE 4 /* 1 (default): keep the unpacked tensor-core operands of the image set across
                                         sfmgms_match_pairs calls; 0: unpack again on every call */
#define SFMGMS_OPT_L2_KERNEL 5        /* sfmgms_bf_l2: 0 auto, 1 DP4A CUDA-core kernel, 2 tcgen05 kind::i8 kernel (1, 2: integer-valued
                                         data, else the fp32 kernel takes over), 3 always the order-exact fp32 kernel */
#define SFMGMS_OPT_TIMING 3          /* 1: record CUDA events around the Hamming and GMS stages of each batch;
                                        2: additionally one event after every kernel (sfmgms_kernel_times) */
#define SFMGMS_OPT_CHUNK_ROWS 6      /* sfmgms_match_pairs[_compact] walk a pair list in chunks of at most this many match
                                        rows (default 4 Mi): per-match device scratch is O(chunk), never O(list) */
#define SFMGMS_OPT_OVERLAP 8         /* only with the separate tie-resolution kernel (SFMGMS_FP4_FUSED_RESOLVE=0 in the environment; by
                                        default the fp4 kernel resolves its own ties): 1 (default) splits a batch's tensor-core work
                                        into up to three launches with the resolution of each on a second stream; 0: one stream */
#define SFMGMS_OPT_COMPACT_RECORD 9  /* records sfmgms_match_pairs_compact writes to `matches`: 0 (default) cv::DMatch, 16 bytes
                                        {queryIdx, trainIdx, imgIdx = 0, float distance}; 1: index pairs, 8 bytes {int32 queryIdx,
                                        int32 trainIdx} -- all that the reference's consumer of matchesGMS reads (SfMUtil.cpp:25-35
                                        gathers kpts1[queryIdx].pt / kpts2[trainIdx].pt); halves the result volume of long pair lists
                                        (the 8-GPU all-pairs run is bound by the host's ingest of the 16-byte records) */
#define SFMGMS_OPT_GMS_DENSE 7       /* 1: force the global-memory histogram path of GMS (otherwise only pairs with
                                        >= 65536 matches take it) -- for tests and measurements */

typedef struct sfmgms_ctx sfmgms_ctx;

/* ---- life cycle ------------------------------------------------------------------------------ */
int sfmgms_create(sfmgms_ctx** out, int device);
void sfmgms_destroy(sfmgms_ctx* ctx);
const char* sfmgms_last_error(const sfmgms_ctx* ctx); /* ctx may be NULL: message of the failed create */
int sfmgms_version(void);
int sfmgms_set_option(sfmgms_ctx* ctx, int key, int64_t value);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t sfmgms_kernel_launches(const sfmgms_ctx* ctx);
/* With SFMGMS_OPT_TIMING on: device milliseconds (CUDA events on the context stream) of the last batch:
 * out_ms[0] = Hamming kernel(s), out_ms[1] = GMS kernels, out_ms[2] = number of Hamming kernel launches. */
int sfmgms_last_timing(sfmgms_ctx* ctx, double* out_ms /*3*/);
/* With SFMGMS_OPT_TIMING = 2 (measurement runs: an event after every kernel): writes "name:total_ms:launches;..." for
 * every kernel launched by multi-pair / single-pair calls since the last read, then clears the accumulators. */
int sfmgms_kernel_times(sfmgms_ctx* ctx, char* buf, int buf_len);
/* the CUDA stream (cudaStream_t) the context launches on, for event timing by the caller */
void* sfmgms_stream(sfmgms_ctx* ctx);

/* ---- stage 1: replaces cv::BFMatcher(NORM_HAMMING, crossCheck=false)::match -------------------
 * (FeatureMatchUtil.cpp:66-68).  For every query row i: train_idx[i] = lowest j minimising
 * popcount(q_i ^ t_j), dist[i] = that popcount; matches are in query order (queryIdx = i, imgIdx = 0).
 * *n_matches = nq, or 0 when nt == 0 (OpenCV returns an empty vector).  nt >= 2^18 -> SFMGMS_ERR_TRAIN_ROWS. */
int sfmgms_bf_hamming(sfmgms_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                      int desc_bytes, int32_t* train_idx, int32_t* dist, int* n_matches);

/* Cross-check variant: cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match (bruteForceMatch,
 * FeatureMatchUtil.cpp:22-23).  keep[i] = 1 iff query i is also the nearest query of its own nearest
 * train row (lowest-index tie-breaks on both sides); OpenCV emits exactly the kept (i, train_idx[i]). */
int sfmgms_bf_hamming_crosscheck(sfmgms_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                                 int desc_bytes, int32_t* train_idx, int32_t* dist, uint8_t* keep);

/*  cv::BFMatcher(NORM_L2, crossCheck=false)::match on CV_32F descriptors — what the reference
 * literally runs (SIFT, FeatureMatchUtil.cpp:10, 66-68).  train_idx[i] = lowest j minimising the FLOAT distance
 * sqrtf(normL2Sqr(q_i, t_j)), dist[i] = that float.  Two kernels behind one call:
 *  - dim == 128 and every value an integer in [0,255] (OpenCV SIFT): OpenCV's float accumulation is exact in any
 *    order; tcgen05 u8 x u8 GEMM with a fused row-argmin epilogue (l2_tc.cu; SFMGMS_OPT_L2_KERNEL = 1: DP4A).
 *    Bit-identical to cv2, including OpenCV's float-tie rule for d2 >= 2^22.
 *  - anything else (RootSIFT, normalised or learned descriptors, other widths; 1 <= dim <= 256): the fp32 kernel
 *    l2_f32.cu restates OpenCV's hal::normL2Sqr_ addition order (4 accumulators x 4 lanes, no FMA, scalar tail) and
 *    is bit-identical to cv2 built with the 4-lane (SSE) baseline — the stock x86-64 packages; a build whose
 *    normL2Sqr_ uses wider vectors may differ in the last ulp of a distance. */
int sfmgms_bf_l2(sfmgms_ctx* ctx, const float* query, int nq, const float* train, int nt, int dim,
                 int32_t* train_idx, float* dist, int* n_matches);

/*  cv::BFMatcher(NORM_L2, crossCheck=true)::match — the matcher bruteForceMatch constructs
 * (FeatureMatchUtil.cpp:22-23).  Same data contract as sfmgms_bf_l2; keep[i] as in the Hamming cross-check
 * (two tensor-core passes, query->train and train->query, joined on the device).  nq must be < 2^18 too. */
int sfmgms_bf_l2_crosscheck(sfmgms_ctx* ctx, const float* query, int nq, const float* train, int nt, int dim,
                            int32_t* train_idx, float* dist, uint8_t* keep);

/*  The whole of bruteForceMatch (FeatureMatchUtil.cpp:20-31; cross_check = 1, the reference's
 * constants are distance_coef = 4.0 = kDistanceCoef and max_matching_size = 500 = kMaxMatchingSize,
 * FeatureMatchUtil.h:17-18) and of the inline match() (FeatureMatchUtil.cpp:38-50; cross_check = 0):
 *   BFMatcher(norm_type, cross_check).match -> sort by distance -> while (front*coef < back) pop_back ->
 *   while (size > max_matching_size) pop_back.
 * norm_type: SFMGMS_NORM_L2 (query/train = const float*, width = 128, integer-valued as for sfmgms_bf_l2) or
 * SFMGMS_NORM_HAMMING (const uint8_t*, width = 32).  Outputs: the surviving matches, ascending distance;
 * *n_out = their number (<= capacity, else SFMGMS_ERR_ARG).  std::sort leaves the order of EQUAL distances
 * unspecified; this library orders them by queryIdx (one of the orders std::sort may produce), so the output
 * is deterministic.  An empty match list gives *n_out = 0 (the reference would dereference front() of an
 * empty vector). */
#define SFMGMS_NORM_L2 4       /* cv::NORM_L2 */
#define SFMGMS_NORM_HAMMING 6  /* cv::NORM_HAMMING */
int sfmgms_brute_force_match(sfmgms_ctx* ctx, int norm_type, int cross_check, const void* query, int nq,
                             const void* train, int nt, int width, double distance_coef, int max_matching_size,
                             int32_t* query_idx, int32_t* train_idx, float* dist, int capacity, int* n_out);

/*  cv::ORB — `f2d = ORB::create()` (DisparityUtil.cpp:107; defaults scaleFactor 1.2f, 8 levels,
 * edgeThreshold 31, firstLevel 0, WTA_K 2, HARRIS_SCORE, patchSize 31).
 *
 * sfmgms_orb_compute = f2d->compute(img, keypoints, descriptors) (DisparityUtil.cpp:127-134, a KeyPoint at every
 * pixel): BGR -> gray, scale pyramid up to the highest octave present, drop keypoints whose rounded position is
 * within 31 pixels of the image border, regroup by octave if the input is not sorted by octave (input order kept
 * inside an octave), 7x7 sigma-2 Gaussian per level, rotated-BRIEF bits with each keypoint's OWN angle (degrees;
 * OpenCV does not re-estimate it on this path; a default KeyPoint carries -1).
 * image: 8-bit, channels 1 (gray) or 3 (BGR), rows stride_bytes apart.  keypoints: (x, y) floats at offset 0 of every
 * kp_stride_bytes; angle float at angle_offset_bytes (12 in cv::KeyPoint; -1 = every angle is -1); octave int32 at
 * octave_offset_bytes (20 in cv::KeyPoint; -1 = every octave is 0); octaves 0..15.
 * Out: kept_index[r] = input index of output row r, descriptors[r*32 .. r*32+31], *n_kept = rows (<= n_keypoints;
 * both arrays must have room for n_keypoints rows). */
int sfmgms_orb_compute(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                       const void* keypoints, int n_keypoints, int kp_stride_bytes, int angle_offset_bytes,
                       int octave_offset_bytes, int32_t* kept_index, uint8_t* descriptors, int* n_kept);

/* sfmgms_orb_detect_and_compute = ORB::create(nfeatures) [+ setFastThreshold(fast_threshold)] ->
 * detectAndCompute(img, noArray(), keypoints, descriptors) (DisparityUtil.cpp:139-140 with the defaults 500 / 20;
 * BASELINE config 1 uses 10000 / 0).  FAST-9/16 + non-max suppression per pyramid level, border filter, retainBest by
 * FAST score then by Harris response, intensity-centroid angle, descriptors.  keypoints: `capacity` records in
 * cv::KeyPoint layout (28 bytes: pt.x, pt.y, size, angle, response, octave, class_id = -1), in OpenCV's output
 * order; descriptors: capacity x 32 bytes (may be NULL: detect only).  *n_keypoints = count; it can exceed nfeatures
 * by ties at a level's cut (OpenCV keeps them) -- capacity too small -> SFMGMS_ERR_ARG with *n_keypoints = needed. */
int sfmgms_orb_detect_and_compute(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                                  int nfeatures, int fast_threshold, void* keypoints, uint8_t* descriptors, int capacity,
                                  int* n_keypoints);

/* The arguments of ORB::create(nfeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize,
 * fastThreshold), in that order.  Implemented: any nfeatures, scaleFactor > 1, nlevels 1..16, edgeThreshold >= 0,
 * WTA_K 2 / 3 / 4, scoreType 0 (HARRIS_SCORE) or 1 (FAST_SCORE), patchSize 2..63 (31: the learned pattern, else
 * OpenCV's seeded random pattern), any fastThreshold; firstLevel must be 0 -- anything else -> SFMGMS_ERR_ARG. */
typedef struct sfmgms_orb_params {
    int nfeatures;        /* 500 */
    float scale_factor;   /* 1.2f */
    int nlevels;          /* 8 */
    int edge_threshold;   /* 31 */
    int first_level;      /* 0 */
    int wta_k;            /* 2 */
    int score_type;       /* 0 = ORB::HARRIS_SCORE */
    int patch_size;       /* 31 */
    int fast_threshold;   /* 20 */
} sfmgms_orb_params;
int sfmgms_orb_detect_and_compute_ex(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                                     const sfmgms_orb_params* params, void* keypoints, uint8_t* descriptors, int capacity,
                                     int* n_keypoints);

/* Pixels in, image set out: runs ORB (the parameters above) on every image and makes the result the image set of
 * sfmgms_match_pairs / sfmgms_match_offsets / sfmgms_inlier_points -- the whole per-pair flow of
 * DisparityUtil.cpp:139-149 (detectAndCompute on both images, match, matchGMS) for a sequence of images.
 * images[i]: 8-bit, channels[i] in {1, 3}, rows strides[i] bytes apart (strides == NULL: packed).
 * kp_offsets_out (n_images + 1, may be NULL): keypoint range of every image.  The keypoints themselves (cv::KeyPoint
 * records, OpenCV's order) are read back with sfmgms_get_image_keypoints. */
int sfmgms_set_images_from_pixels(sfmgms_ctx* ctx, int n_images, const uint8_t* const* images, const int32_t* widths,
                                  const int32_t* heights, const int32_t* channels, const int32_t* strides,
                                  const sfmgms_orb_params* params, int64_t* kp_offsets_out);
int sfmgms_get_image_keypoints(sfmgms_ctx* ctx, int image, void* keypoints /* capacity x 28 bytes */, int capacity, int* n_out);

/* ---- stage 2: replaces cv::xfeatures2d::matchGMS ----------------------------------------------
 * (FeatureMatchUtil.cpp:69; DisparityUtil.cpp:149,299).  mask[i] (0/1) for each of the n_matches input
 * matches; *mask_len = n_matches, or 0 if rotation/scale search was requested and every hypothesis had
 * zero inliers (the reference then leaves its vector<bool> untouched = empty); *n_inliers = the count.
 * best_hyp (may be NULL): scaleIdx*8 + (rotationType-1) of the winning hypothesis, -1 if none.
 * NOTE argument order: with_rotation BEFORE with_scale, as in OpenCV's matchGMS. */
int sfmgms_gms(sfmgms_ctx* ctx, int w1, int h1, int w2, int h2, const void* kp1, int n1, int kp1_stride_bytes,
               const void* kp2, int n2, int kp2_stride_bytes, const int32_t* query_idx,
               const int32_t* train_idx, int idx_stride_bytes, int n_matches, int with_rotation,
               int with_scale, double threshold_factor, uint8_t* mask, int* mask_len, int* n_inliers,
               int* best_hyp);

/* ---- fused stage 1+2 for one pair (the whole FeatureMatchUtil.cpp:66-69 block) -----------------
 * Matches never leave the device between the stages.  Outputs as above; any of train_idx/dist/mask may
 * be NULL.  n_matches = n1 (0 when n2 == 0). */
int sfmgms_match_pair(sfmgms_ctx* ctx, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2,
                      int desc_bytes, const void* kp1, int kp1_stride_bytes, const void* kp2,
                      int kp2_stride_bytes, int w1, int h1, int w2, int h2, int with_rotation, int with_scale,
                      double threshold_factor, int32_t* train_idx, int32_t* dist, uint8_t* mask, int* mask_len,
                      int* n_inliers, int* best_hyp);

/* ---- multi-pair (all-pairs / sliding-window SfM matching over an image set) --------------------
 * sfmgms_set_images: register n_images images.  kp_offsets[n_images+1] (host array) gives each
 * image's first keypoint row in desc (rows of 32 bytes) and kp_xy (rows of 2 floats); sizes_wh is a HOST
 * array [n_images*2] (width, height).  location == SFMGMS_HOST: desc/kp_xy are host buffers, copied to
 * the device (pinned staging, async H2D).  location == SFMGMS_DEVICE: desc/kp_xy are device pointers on the
 * context's device (e.g. the target of an NCCL broadcast) and are adopted without a copy; they must stay
 * alive until the next set_images/destroy. */
int sfmgms_set_images(sfmgms_ctx* ctx, int n_images, const int64_t* kp_offsets, const uint8_t* desc,
                      const float* kp_xy, const int32_t* sizes_wh, int location);

/* sfmgms_match_pairs: for each pair p = (pairs[2p], pairs[2p+1]) = (query image, train image) run
 * BFMatcher::match + matchGMS.  Per-match outputs are concatenated in pair order; pair p's rows start at
 * match_offsets[p] (= sum of n1 over earlier pairs; sfmgms_match_offsets computes them).
 * out_location selects host or device pointers for ALL outputs; each output may be NULL.
 *   n_inliers[n_pairs], best_hyp[n_pairs], mask_len[n_pairs]  (int32)
 *   train_idx[total], dist[total] (int32), mask[total] (uint8, all-zero for a pair whose mask_len is 0) */
int sfmgms_match_pairs(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                       double threshold_factor, int out_location, int32_t* n_inliers, int32_t* best_hyp,
                       int32_t* mask_len, int32_t* train_idx, int32_t* dist, uint8_t* mask);
int sfmgms_match_offsets(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int64_t* match_offsets /*n_pairs+1*/);

/* sfmgms_match_pairs_compact: the same computation, returning what the reference's caller actually keeps -- the
 * `matchesGMS` vector of every pair (FeatureMatchUtil.cpp:69: matchGMS clears it and pushes matches1to2[i] for every set
 * mask bit, in input order) and the coordinate lists SfMUtil.cpp:25-35 / DisparityUtil.cpp:179-185 gather from it --
 * for the whole pair list, back to back:
 *   n_inliers[n_pairs], best_hyp[n_pairs]      int32 per pair
 *   inlier_offsets[n_pairs + 1]                int64: pair p owns rows [inlier_offsets[p], inlier_offsets[p+1])
 *   matches[capacity]                          16-byte cv::DMatch records {queryIdx, trainIdx, imgIdx = 0, distance (float)}
 *                                              (SFMGMS_OPT_COMPACT_RECORD = 1: 8-byte {int32 queryIdx, int32 trainIdx} rows)
 *   pts1[capacity * 2], pts2[capacity * 2]     kp1[queryIdx].pt / kp2[trainIdx].pt of the same rows
 * Any output may be NULL.  *n_total = total inliers; if it exceeds `capacity` the first `capacity` rows are valid and
 * the call returns SFMGMS_ERR_CAPACITY.  out_location selects host or device pointers for all outputs.  A pair costs
 * 16-32 bytes per INLIER on the way back to the host instead of 9 bytes per MATCH plus a per-pair gather call. */
int sfmgms_match_pairs_compact(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                               double threshold_factor, int out_location, int32_t* n_inliers, int32_t* best_hyp,
                               int64_t* inlier_offsets, void* matches, float* pts1, float* pts2, int64_t capacity,
                               int64_t* n_total);

/* Asynchronous variants (SURVEY §8b "synchronous by default with an async stream variant"): DEVICE outputs only.  The whole
 * pair list is enqueued on the context's stream (sfmgms_stream) and the call returns; outputs are valid after sfmgms_wait
 * (or after the caller's own event / stream synchronisation on that stream).  Conditions the device detects (keypoint
 * outside the image, index out of range, compact capacity exceeded) are reported by sfmgms_wait, which also returns the total
 * number of inliers.  No other call on the context between the two.  pairs must stay valid until the call returns only. */
int sfmgms_match_pairs_async(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                             double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int32_t* mask_len,
                             int32_t* train_idx, int32_t* dist, uint8_t* mask);
int sfmgms_match_pairs_compact_async(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                                     double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_offsets,
                                     void* matches, float* pts1, float* pts2, int64_t capacity);
int sfmgms_wait(sfmgms_ctx* ctx, int64_t* n_total /* may be NULL */);

/* sfmgms_match_image_set: sfmgms_set_images(SFMGMS_HOST) + sfmgms_match_pairs(SFMGMS_HOST outputs) in ONE
 * synchronous call that pipelines internally: the image set goes to the device in chunks on a copy stream while
 * earlier pairs already compute, and finished pairs' results return on a third stream.  Pairs are processed in
 * list order; a pair starts as soon as both of its images have arrived, so order the list by image index for the
 * best overlap (consecutive / sliding-window pairs do this naturally).  All pointers are HOST pointers (pinned
 * memory gives true overlap; pageable memory still works).  Outputs as for sfmgms_match_pairs.  On return the
 * image set stays registered (as after sfmgms_set_images). */
int sfmgms_match_image_set(sfmgms_ctx* ctx, int n_images, const int64_t* kp_offsets, const uint8_t* desc,
                           const float* kp_xy, const int32_t* sizes_wh, const int32_t* pairs, int n_pairs,
                           int with_rotation, int with_scale, double threshold_factor, int32_t* n_inliers,
                           int32_t* best_hyp, int32_t* mask_len, int32_t* train_idx, int32_t* dist, uint8_t* mask);

/* ----  inlier coordinate compaction: replaces the gather loop SfMUtil.cpp:25-35 ------
 * After sfmgms_match_pairs, emit for pair `pair_index` of the LAST batch the inlier coordinates
 * pts1[k] = kp1[queryIdx].pt, pts2[k] = kp2[trainIdx].pt in match order (k < n_inliers), ready for
 * findEssentialMat (SfMUtil.cpp:39).  pts1/pts2: host float arrays of capacity*2; *n_out = n_inliers. */
int sfmgms_inlier_points(sfmgms_ctx* ctx, int pair_index, float* pts1, float* pts2, int capacity, int* n_out);
/* (a one-pair view of sfmgms_match_pairs_compact; valid after a single-pair call or a pair list that ran as ONE chunk,
 * see SFMGMS_OPT_CHUNK_ROWS -- for whole lists use the compact call.) */

/* Inlier counts of all 40 (scale, rotation) hypotheses of one pair, scale-major: what GMSMatcher::run returns for each
 * (rotation 1..8 inside scale 0..4) and getInlierMask compares (DLL @VA 0x180047dc0).  Inputs as for sfmgms_gms. */
int sfmgms_gms_hypotheses(sfmgms_ctx* ctx, int w1, int h1, int w2, int h2, const void* kp1, int n1, int kp1_stride_bytes,
                          const void* kp2, int n2, int kp2_stride_bytes, const int32_t* query_idx,
                          const int32_t* train_idx, int idx_stride_bytes, int n_matches, double threshold_factor,
                          int32_t* counts /*40*/);

/* device memory currently owned by the context (bytes) */
int64_t sfmgms_device_bytes(const sfmgms_ctx* ctx);

/* Page-locked host memory for image sets and result tables handed to the calls above with SFMGMS_HOST (the reference
 * keeps descriptors / key points in cv::Mat / std::vector storage, FeatureMatchUtil.cpp:49-56; pageable memory works
 * too but is staged through a bounce buffer).  cudaHostAlloc(portable): visible to every GPU of a sfmgms_multi.
 * Measured on this pool: 55 GB/s H2D from these buffers (profiles/r2_h2d_probe.txt).  No context needed; returns
 * SFMGMS_ERR_CUDA when the allocation fails, SFMGMS_ERR_ARG on NULL. */
int sfmgms_host_alloc(size_t bytes, void** out);
int sfmgms_host_free(void* p);

/* ---- multi-GPU: one process, one host thread per GPU, pairs sharded, ONE broadcast of the set ----
 * The reference matches one pair per call (FeatureMatchUtil.cpp:66-69) inside loops over an image sequence
 * (main.cpp:32,39,47; SfMUtil.cpp:16-18).  A sfmgms_multi owns one sfmgms_ctx per GPU.  sfmgms_multi_set_images sends the
 * HOST image set to the first GPU once and from there to all others with one ncclBroadcast per array (ncclCommInitAll over
 * the listed GPUs; NCCL is bound with dlopen("libnccl.so.2") when n_devices > 1 -> SFMGMS_ERR_NCCL if absent).  The match
 * calls cut the pair list into contiguous shards of equal work (one per GPU, in list order), run them on one host thread
 * per GPU and write every GPU's results straight into the caller's HOST buffers; there is no collective on the result
 * path.  Results are bit-identical to the single-GPU calls.
 *   devices == NULL: GPUs 0 .. n_devices-1 (n_devices <= 0: all visible GPUs).
 *   sfmgms_multi_match_pairs          outputs exactly as sfmgms_match_pairs(SFMGMS_HOST).
 *   sfmgms_multi_match_pairs_compact  as sfmgms_match_pairs_compact(SFMGMS_HOST), except that the GPUs append to the one
 *       matches / pts1 / pts2 buffer as their chunks finish: rows of a pair are contiguous and in match order, pairs are
 *       NOT in list order; inlier_begin[p] (n_pairs entries) is the first row of pair p, n_inliers[p] its row count. */
typedef struct sfmgms_multi sfmgms_multi;
int sfmgms_multi_create(sfmgms_multi** out, const int* devices, int n_devices);
void sfmgms_multi_destroy(sfmgms_multi* m);
const char* sfmgms_multi_last_error(const sfmgms_multi* m); /* m may be NULL: message of the failed create */
int sfmgms_multi_device_count(const sfmgms_multi* m);
sfmgms_ctx* sfmgms_multi_context(sfmgms_multi* m, int i);  /* the i-th GPU's context (options, timing, launches) */
double sfmgms_multi_last_broadcast_ms(const sfmgms_multi* m); /* device time of the last set's broadcast on GPU 0 */
int sfmgms_multi_set_images(sfmgms_multi* m, int n_images, const int64_t* kp_offsets, const uint8_t* desc,
                            const float* kp_xy, const int32_t* sizes_wh);
int sfmgms_multi_match_pairs(sfmgms_multi* m, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                             double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int32_t* mask_len,
                             int32_t* train_idx, int32_t* dist, uint8_t* mask);
int sfmgms_multi_match_pairs_compact(sfmgms_multi* m, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                                     double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_begin,
                                     void* matches, float* pts1, float* pts2, int64_t capacity, int64_t* n_total);

#ifdef __cplusplus
}
#endif
#endif /* SFMGMS_H_ */
