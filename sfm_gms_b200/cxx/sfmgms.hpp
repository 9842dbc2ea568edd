// sfmgms.hpp — header-only C++ mirror of the reference's matching interface over the C ABI (include/sfmgms.h).
//
// The reference calls, back to back (SfM-GMS/SfM-GMS/FeatureMatchUtil.cpp:66-69):
//     Ptr<BFMatcher> matcherBF = BFMatcher::create();  matcherBF->match(desc1, desc2, matches);
//     cv::xfeatures2d::matchGMS(img1.size(), img2.size(), kpts1, kpts2, matches, matchesGMS, true, true);
// This header offers the same names, argument order and error behaviour on layout-compatible PODs
// (OpenCV's C++ headers are not needed):
//     sfmgms::cv::BFMatcher(sfmgms::cv::NORM_HAMMING[, crossCheck]).match(desc1, n1, desc2, n2, matches)
//     sfmgms::cv::xfeatures2d::matchGMS(size1, size2, kp1, kp2, matches1to2, matchesGMS,
//                                       withRotation=false, withScale=false, thresholdFactor=6.0)
//     sfmgms::gms_matcher(vkp1, size1, vkp2, size2, vDMatches).GetInlierMask(vbInliers, WithScale, WithRotation)
// and, with -DSFMGMS_WITH_OPENCV (OpenCV headers present), overloads on the real cv:: types so that
// FeatureMatchUtil.cpp only has to switch the two call lines.
//
// NOTE the two upstream conventions: matchGMS(..., withRotation, withScale, ...) versus
// GetInlierMask(mask, WithScale, WithRotation).  Both are reproduced as they are; nothing is transposed.
// Errors: where OpenCV throws cv::Exception (CV_Assert) this throws sfmgms::Error (std::runtime_error).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/sfmgms.h"

namespace sfmgms {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("sfmgms error " + std::to_string(c) + ": " + m), code(c) {}
};

// One context per host thread / per GPU (SURVEY §8b threading).
class Context {
   public:
    explicit Context(int device = 0) {
        if (int rc = sfmgms_create(&ctx_, device)) throw Error(rc, sfmgms_last_error(nullptr));
    }
    ~Context() { sfmgms_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    sfmgms_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc) throw Error(rc, sfmgms_last_error(ctx_));
    }
    static Context& thread_default() {
        static thread_local Context c(0);
        return c;
    }

   private:
    sfmgms_ctx* ctx_ = nullptr;
};

namespace cv {

enum { NORM_L2 = 4, NORM_HAMMING = 6 };

// Layout-compatible with cv::Point2f / cv::Size / cv::KeyPoint (28 B) / cv::DMatch (16 B); strides verified
// against the reference binary (0x1c and 0x10).
struct Point2f { float x, y; };
struct Size {
    int width, height;
    Size(int w = 0, int h = 0) : width(w), height(h) {}
};
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");
static_assert(sizeof(DMatch) == 16, "cv::DMatch layout");

// cv::BFMatcher(NORM_HAMMING[, crossCheck]) — descriptors are CV_8U rows of 32 bytes.
class BFMatcher {
   public:
    explicit BFMatcher(int normType = NORM_HAMMING, bool crossCheck = false, Context* ctx = nullptr)
        : cross_(crossCheck), ctx_(ctx) {
        if (normType != NORM_HAMMING && normType != NORM_L2)
            throw Error(SFMGMS_ERR_ARG, "implemented: NORM_HAMMING and NORM_L2, with/without crossCheck");
        norm_ = normType;
    }
    // NORM_L2 on integer-valued float descriptors with 128 columns (OpenCV SIFT; FeatureMatchUtil.cpp:10, 66-68):
    // distances are bit-identical to OpenCV's float results.
    void match(const float* query, int nq, const float* train, int nt, std::vector<DMatch>& matches, int dim = 128) const {
        if (norm_ != NORM_L2) throw Error(SFMGMS_ERR_ARG, "float descriptors need NORM_L2");
        Context& c = ctx_ ? *ctx_ : Context::thread_default();
        matches.clear();
        std::vector<int32_t> idx((size_t)(nq > 0 ? nq : 0));
        std::vector<float> dist(idx.size());
        if (cross_) {   // BFMatcher(NORM_L2, true): the matcher of bruteForceMatch (FeatureMatchUtil.cpp:22)
            std::vector<uint8_t> keep(idx.size());
            c.check(sfmgms_bf_l2_crosscheck(c.get(), query, nq, train, nt, dim, idx.data(), dist.data(), keep.data()));
            if (nt == 0) return;
            for (int i = 0; i < nq; ++i)
                if (keep[i]) matches.push_back(DMatch{i, idx[i], 0, dist[i]});
            return;
        }
        int n = 0;
        c.check(sfmgms_bf_l2(c.get(), query, nq, train, nt, dim, idx.data(), dist.data(), &n));
        matches.reserve((size_t)n);
        for (int i = 0; i < n; ++i) matches.push_back(DMatch{i, idx[i], 0, dist[i]});
    }
    // matches.clear(); then one DMatch per query row in query order (cross-check: only the mutual ones)
    void match(const uint8_t* query, int nq, const uint8_t* train, int nt, std::vector<DMatch>& matches,
               int desc_bytes = 32) const {
        if (norm_ != NORM_HAMMING) throw Error(SFMGMS_ERR_ARG, "uint8 descriptors need NORM_HAMMING");
        Context& c = ctx_ ? *ctx_ : Context::thread_default();
        matches.clear();
        std::vector<int32_t> idx((size_t)(nq > 0 ? nq : 0)), dist(idx.size());
        if (cross_) {
            std::vector<uint8_t> keep(idx.size());
            c.check(sfmgms_bf_hamming_crosscheck(c.get(), query, nq, train, nt, desc_bytes, idx.data(), dist.data(), keep.data()));
            if (nt == 0) return;
            for (int i = 0; i < nq; ++i)
                if (keep[i]) matches.push_back(DMatch{i, idx[i], 0, (float)dist[i]});
            return;
        }
        int n = 0;
        c.check(sfmgms_bf_hamming(c.get(), query, nq, train, nt, desc_bytes, idx.data(), dist.data(), &n));
        matches.reserve((size_t)n);
        for (int i = 0; i < n; ++i) matches.push_back(DMatch{i, idx[i], 0, (float)dist[i]});
    }

   private:
    bool cross_;
    Context* ctx_;
    int norm_ = NORM_HAMMING;
};

// cv::ORB: ORB::create(nfeatures) with OpenCV's other defaults, as DisparityUtil.cpp:107 builds it.
// detectAndCompute / detect / compute return what OpenCV returns, keypoint order included.  Like OpenCV, compute()
// REMOVES the keypoints it cannot describe (within 31 pixels of the border) from `keypoints`.
class ORB {
   public:
    explicit ORB(int nfeatures = 500, Context* ctx = nullptr) : nfeatures_(nfeatures), ctx_(ctx) {}
    static ORB create(int nfeatures = 500) { return ORB(nfeatures); }
    void setFastThreshold(int t) { fast_ = t; }
    int getFastThreshold() const { return fast_; }
    int descriptorSize() const { return 32; }
    // image: 8-bit, `channels` 1 (gray) or 3 (BGR), rows `stride_bytes` apart (0 = packed)
    void detectAndCompute(const uint8_t* image, Size size, int channels, std::vector<KeyPoint>& keypoints,
                          std::vector<uint8_t>& descriptors, int stride_bytes = 0) const {
        run(image, size, channels, keypoints, &descriptors, stride_bytes);
    }
    void detect(const uint8_t* image, Size size, int channels, std::vector<KeyPoint>& keypoints, int stride_bytes = 0) const {
        run(image, size, channels, keypoints, nullptr, stride_bytes);
    }
    void compute(const uint8_t* image, Size size, int channels, std::vector<KeyPoint>& keypoints,
                 std::vector<uint8_t>& descriptors, int stride_bytes = 0) const {
        Context& c = ctx_ ? *ctx_ : Context::thread_default();
        const int n = (int)keypoints.size();
        std::vector<int32_t> kept((size_t)n);
        descriptors.assign((size_t)n * 32, 0);
        int nk = 0;
        c.check(sfmgms_orb_compute(c.get(), image, size.width, size.height, channels,
                                   stride_bytes ? stride_bytes : size.width * channels, n ? &keypoints[0].pt.x : nullptr, n,
                                   (int)sizeof(KeyPoint), 12, 20, kept.data(), descriptors.data(), &nk));
        std::vector<KeyPoint> out((size_t)nk);
        for (int i = 0; i < nk; ++i) out[(size_t)i] = keypoints[(size_t)kept[(size_t)i]];
        keypoints.swap(out);
        descriptors.resize((size_t)nk * 32);
    }

   private:
    void run(const uint8_t* image, Size size, int channels, std::vector<KeyPoint>& keypoints, std::vector<uint8_t>* descriptors,
             int stride_bytes) const {
        Context& c = ctx_ ? *ctx_ : Context::thread_default();
        int cap = 2 * nfeatures_ + 64, n = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            keypoints.assign((size_t)cap, KeyPoint{});
            if (descriptors) descriptors->assign((size_t)cap * 32, 0);
            const int rc = sfmgms_orb_detect_and_compute(c.get(), image, size.width, size.height, channels,
                                                         stride_bytes ? stride_bytes : size.width * channels, nfeatures_, fast_,
                                                         keypoints.data(), descriptors ? descriptors->data() : nullptr, cap, &n);
            if (rc == SFMGMS_ERR_ARG && n > cap) { cap = n; continue; }    // ties at a level's cut: retry once
            c.check(rc);
            break;
        }
        keypoints.resize((size_t)n);
        if (descriptors) descriptors->resize((size_t)n * 32);
    }
    int nfeatures_, fast_ = 20;
    Context* ctx_;
};

namespace detail {
inline void brute_force(int norm, bool cross, const void* d1, int n1, const void* d2, int n2, int width,
                        std::vector<DMatch>& matches, double coef, int max_size, Context* ctx) {
    Context& c = ctx ? *ctx : Context::thread_default();
    matches.clear();
    const size_t cap = (size_t)(n1 > 0 ? n1 : 0);
    std::vector<int32_t> q(cap), t(cap);
    std::vector<float> d(cap);
    int n = 0;
    c.check(sfmgms_brute_force_match(c.get(), norm, cross ? 1 : 0, d1, n1, d2, n2, width, coef, max_size, q.data(),
                                     t.data(), d.data(), (int)cap, &n));
    matches.reserve((size_t)n);
    for (int i = 0; i < n; ++i) matches.push_back(DMatch{q[i], t[i], 0, d[i]});
}
}  // namespace detail

}  // namespace cv

// The reference's own helpers (FeatureMatchUtil.h:17-18, FeatureMatchUtil.cpp:20-31 and :38-50), same names:
constexpr double kDistanceCoef = 4.0;
constexpr int kMaxMatchingSize = 500;
// bruteForceMatch(desc1, desc2, matches): BFMatcher(NORM_L2, true).match, sort, ratio prune, cap 500
inline void bruteForceMatch(const float* desc1, int n1, const float* desc2, int n2, std::vector<cv::DMatch>& matches,
                            Context* ctx = nullptr) {
    cv::detail::brute_force(SFMGMS_NORM_L2, true, desc1, n1, desc2, n2, 128, matches, kDistanceCoef, kMaxMatchingSize, ctx);
}
inline void bruteForceMatch(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, std::vector<cv::DMatch>& matches,
                            Context* ctx = nullptr) {   // ORB descriptors: the same flow under NORM_HAMMING
    cv::detail::brute_force(SFMGMS_NORM_HAMMING, true, desc1, n1, desc2, n2, 32, matches, kDistanceCoef, kMaxMatchingSize, ctx);
}
// match(desc1, desc2, matches, kDistanceCoef, kMaxMatchingSize): BFMatcher::create() (NORM_L2, no cross-check)
inline void match(const float* desc1, int n1, const float* desc2, int n2, std::vector<cv::DMatch>& matches,
                  double distanceCoef, int maxMatchingSize, Context* ctx = nullptr) {
    cv::detail::brute_force(SFMGMS_NORM_L2, false, desc1, n1, desc2, n2, 128, matches, distanceCoef, maxMatchingSize, ctx);
}

namespace cv {
namespace xfeatures2d {

// cv::xfeatures2d::matchGMS — exact OpenCV parameter order (…, withRotation, withScale, thresholdFactor).
inline void matchGMS(const Size& size1, const Size& size2, const std::vector<KeyPoint>& keypoints1,
                     const std::vector<KeyPoint>& keypoints2, const std::vector<DMatch>& matches1to2,
                     std::vector<DMatch>& matchesGMS, bool withRotation = false, bool withScale = false,
                     double thresholdFactor = 6.0, Context* ctx = nullptr) {
    Context& c = ctx ? *ctx : Context::thread_default();
    std::vector<uint8_t> mask(matches1to2.size() + 1);
    int mask_len = 0, n_inliers = 0;
    c.check(sfmgms_gms(c.get(), size1.width, size1.height, size2.width, size2.height,
                       keypoints1.empty() ? nullptr : &keypoints1[0].pt.x, (int)keypoints1.size(), (int)sizeof(KeyPoint),
                       keypoints2.empty() ? nullptr : &keypoints2[0].pt.x, (int)keypoints2.size(), (int)sizeof(KeyPoint),
                       matches1to2.empty() ? nullptr : &matches1to2[0].queryIdx,
                       matches1to2.empty() ? nullptr : &matches1to2[0].trainIdx, (int)sizeof(DMatch),
                       (int)matches1to2.size(), withRotation ? 1 : 0, withScale ? 1 : 0, thresholdFactor, mask.data(),
                       &mask_len, &n_inliers, nullptr));
    matchesGMS.clear();                                   // as the reference does (DLL @VA 0x18004831e)
    for (int i = 0; i < mask_len; ++i)
        if (mask[i]) matchesGMS.push_back(matches1to2[i]);  // input order preserved
}

}  // namespace xfeatures2d
}  // namespace cv

// Upstream header-only API (JiaWang-Bian gms_matcher.h), of which OpenCV's GMSMatcher is the port.
class gms_matcher {
   public:
    gms_matcher(const std::vector<cv::KeyPoint>& vkp1, const cv::Size size1, const std::vector<cv::KeyPoint>& vkp2,
                const cv::Size size2, const std::vector<cv::DMatch>& vDMatches, Context* ctx = nullptr)
        : kp1_(vkp1), kp2_(vkp2), m_(vDMatches), s1_(size1), s2_(size2), ctx_(ctx) {}

    // returns the number of inliers; vbInliers gets one flag per input match.
    // Upstream order: WithScale FIRST, WithRotation second; threshold factor fixed at 6.
    int GetInlierMask(std::vector<bool>& vbInliers, bool WithScale = false, bool WithRotation = false) {
        Context& c = ctx_ ? *ctx_ : Context::thread_default();
        std::vector<uint8_t> mask(m_.size() + 1);
        int mask_len = 0, n_inliers = 0;
        c.check(sfmgms_gms(c.get(), s1_.width, s1_.height, s2_.width, s2_.height, kp1_.empty() ? nullptr : &kp1_[0].pt.x,
                           (int)kp1_.size(), (int)sizeof(cv::KeyPoint), kp2_.empty() ? nullptr : &kp2_[0].pt.x,
                           (int)kp2_.size(), (int)sizeof(cv::KeyPoint), m_.empty() ? nullptr : &m_[0].queryIdx,
                           m_.empty() ? nullptr : &m_[0].trainIdx, (int)sizeof(cv::DMatch), (int)m_.size(),
                           WithRotation ? 1 : 0, WithScale ? 1 : 0, 6.0, mask.data(), &mask_len, &n_inliers, nullptr));
        if (mask_len > 0 || (!WithScale && !WithRotation)) {   // best == 0 with a search on: vector left untouched
            vbInliers.assign((size_t)mask_len, false);
            for (int i = 0; i < mask_len; ++i) vbInliers[i] = mask[i] != 0;
        }
        return n_inliers;
    }

   private:
    const std::vector<cv::KeyPoint>& kp1_;
    const std::vector<cv::KeyPoint>& kp2_;
    const std::vector<cv::DMatch>& m_;
    cv::Size s1_, s2_;
    Context* ctx_;
};

// The whole FeatureMatchUtil.cpp:66-69 block in one call: matches never leave the device between stages.
inline void matchBFHammingGMS(const uint8_t* desc1, const uint8_t* desc2, const std::vector<cv::KeyPoint>& kpts1,
                              const std::vector<cv::KeyPoint>& kpts2, const cv::Size& size1, const cv::Size& size2,
                              std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& matchesGMS,
                              bool withRotation = false, bool withScale = false, double thresholdFactor = 6.0,
                              Context* ctx = nullptr) {
    Context& c = ctx ? *ctx : Context::thread_default();
    const int n1 = (int)kpts1.size(), n2 = (int)kpts2.size();
    std::vector<int32_t> idx((size_t)n1 + 1), dist((size_t)n1 + 1);
    std::vector<uint8_t> mask((size_t)n1 + 1);
    int mask_len = 0, n_inl = 0;
    c.check(sfmgms_match_pair(c.get(), desc1, n1, desc2, n2, 32, n1 ? &kpts1[0].pt.x : nullptr, (int)sizeof(cv::KeyPoint),
                              n2 ? &kpts2[0].pt.x : nullptr, (int)sizeof(cv::KeyPoint), size1.width, size1.height,
                              size2.width, size2.height, withRotation ? 1 : 0, withScale ? 1 : 0, thresholdFactor,
                              idx.data(), dist.data(), mask.data(), &mask_len, &n_inl, nullptr));
    matches.clear();
    matchesGMS.clear();
    const int nm = n2 == 0 ? 0 : n1;
    for (int i = 0; i < nm; ++i) matches.push_back(cv::DMatch{i, idx[i], 0, (float)dist[i]});
    for (int i = 0; i < mask_len; ++i)
        if (mask[i]) matchesGMS.push_back(matches[i]);
}

// ---- image sequences: the FeatureMatchUtil.cpp:66-69 block applied to a list of image pairs (all-pairs / sliding window
// SfM matching; the reference drives it pair by pair from main.cpp:32,39,47 and SfMUtil.cpp:16-18) ---------------------------
struct ImageSet {
    std::vector<int64_t> offsets{0};
    std::vector<uint8_t> desc;      // 32 bytes per keypoint
    std::vector<float> kp_xy;       // pt.x, pt.y per keypoint
    std::vector<int32_t> sizes_wh;  // width, height per image
    int size() const { return (int)offsets.size() - 1; }
    // one image: its keypoints, its ORB descriptors (keypoints.size() rows of 32 bytes), its size
    void add(const std::vector<cv::KeyPoint>& keypoints, const uint8_t* descriptors, cv::Size size) {
        for (const cv::KeyPoint& k : keypoints) { kp_xy.push_back(k.pt.x); kp_xy.push_back(k.pt.y); }
        desc.insert(desc.end(), descriptors, descriptors + keypoints.size() * 32);
        offsets.push_back(offsets.back() + (int64_t)keypoints.size());
        sizes_wh.push_back(size.width); sizes_wh.push_back(size.height);
    }
};

// Every pair's `matchesGMS` (what matchGMS leaves in its output vector) and the coordinate lists SfMUtil.cpp:25-35 builds
// from it, for a whole pair list: pair p owns rows [begin[p], begin[p] + n_inliers[p]) of matches / pts1 / pts2.
struct PairMatches {
    std::vector<int32_t> n_inliers, best_hyp;
    std::vector<int64_t> begin;
    std::vector<cv::DMatch> matches;
    std::vector<cv::Point2f> pts1, pts2;
    std::vector<cv::DMatch> matchesGMS(int p) const {
        return std::vector<cv::DMatch>(matches.begin() + begin[(size_t)p], matches.begin() + begin[(size_t)p] + n_inliers[(size_t)p]);
    }
};

namespace detail {
inline std::vector<int32_t> flat_pairs(const std::vector<std::pair<int, int>>& pairs) {
    std::vector<int32_t> f;
    f.reserve(pairs.size() * 2);
    for (const auto& p : pairs) { f.push_back(p.first); f.push_back(p.second); }
    return f;
}
inline int64_t pair_rows(const ImageSet& s, const std::vector<std::pair<int, int>>& pairs) {
    int64_t rows = 0;
    for (const auto& p : pairs)
        if (p.first >= 0 && p.first < s.size()) rows += s.offsets[(size_t)p.first + 1] - s.offsets[(size_t)p.first];
    return rows;
}
}  // namespace detail

// one GPU: Context + image set + pair list
inline void matchPairs(Context& c, const ImageSet& set, const std::vector<std::pair<int, int>>& pairs, PairMatches& out,
                       bool withRotation = false, bool withScale = false, double thresholdFactor = 6.0, bool wantPoints = true) {
    static const uint8_t zero_d[32] = {0};
    static const float zero_k[2] = {0.f, 0.f};
    c.check(sfmgms_set_images(c.get(), set.size(), set.offsets.data(), set.desc.empty() ? zero_d : set.desc.data(),
                              set.kp_xy.empty() ? zero_k : set.kp_xy.data(), set.sizes_wh.data(), SFMGMS_HOST));
    const std::vector<int32_t> fp = detail::flat_pairs(pairs);
    const int n = (int)pairs.size();
    const int64_t cap = detail::pair_rows(set, pairs);
    out.n_inliers.assign((size_t)n, 0); out.best_hyp.assign((size_t)n, -1);
    std::vector<int64_t> off((size_t)n + 1, 0);
    out.matches.resize((size_t)cap);
    out.pts1.resize(wantPoints ? (size_t)cap : 0); out.pts2.resize(wantPoints ? (size_t)cap : 0);
    int64_t total = 0;
    c.check(sfmgms_match_pairs_compact(c.get(), fp.data(), n, withRotation ? 1 : 0, withScale ? 1 : 0, thresholdFactor, SFMGMS_HOST,
                                       out.n_inliers.data(), out.best_hyp.data(), off.data(), out.matches.data(),
                                       wantPoints && cap ? &out.pts1[0].x : nullptr, wantPoints && cap ? &out.pts2[0].x : nullptr, cap,
                                       &total));
    out.begin.assign(off.begin(), off.begin() + n);
    out.matches.resize((size_t)total);
    if (wantPoints) { out.pts1.resize((size_t)total); out.pts2.resize((size_t)total); }
}

// several GPUs of one box: one host thread per GPU inside the library, the set is broadcast once (NCCL), pairs are sharded
class MultiGpuMatcher {
   public:
    explicit MultiGpuMatcher(const std::vector<int>& devices = {}) {   // empty: all visible GPUs
        if (int rc = sfmgms_multi_create(&m_, devices.empty() ? nullptr : devices.data(), (int)devices.size()))
            throw Error(rc, sfmgms_multi_last_error(nullptr));
    }
    ~MultiGpuMatcher() { sfmgms_multi_destroy(m_); }
    MultiGpuMatcher(const MultiGpuMatcher&) = delete;
    MultiGpuMatcher& operator=(const MultiGpuMatcher&) = delete;
    int deviceCount() const { return sfmgms_multi_device_count(m_); }
    double lastBroadcastMs() const { return sfmgms_multi_last_broadcast_ms(m_); }
    void setImages(const ImageSet& set) {
        static const uint8_t zero_d[32] = {0};
        static const float zero_k[2] = {0.f, 0.f};
        check(sfmgms_multi_set_images(m_, set.size(), set.offsets.data(), set.desc.empty() ? zero_d : set.desc.data(),
                                      set.kp_xy.empty() ? zero_k : set.kp_xy.data(), set.sizes_wh.data()));
        set_ = &set;
    }
    void matchPairs(const std::vector<std::pair<int, int>>& pairs, PairMatches& out, bool withRotation = false, bool withScale = false,
                    double thresholdFactor = 6.0, bool wantPoints = true) {
        if (!set_) throw Error(SFMGMS_ERR_STATE, "setImages has not been called");
        const std::vector<int32_t> fp = detail::flat_pairs(pairs);
        const int n = (int)pairs.size();
        const int64_t cap = detail::pair_rows(*set_, pairs);
        out.n_inliers.assign((size_t)n, 0); out.best_hyp.assign((size_t)n, -1); out.begin.assign((size_t)n, 0);
        out.matches.resize((size_t)cap);
        out.pts1.resize(wantPoints ? (size_t)cap : 0); out.pts2.resize(wantPoints ? (size_t)cap : 0);
        int64_t total = 0;
        check(sfmgms_multi_match_pairs_compact(m_, fp.data(), n, withRotation ? 1 : 0, withScale ? 1 : 0, thresholdFactor,
                                               out.n_inliers.data(), out.best_hyp.data(), out.begin.data(), out.matches.data(),
                                               wantPoints && cap ? &out.pts1[0].x : nullptr, wantPoints && cap ? &out.pts2[0].x : nullptr,
                                               cap, &total));
        out.matches.resize((size_t)total);
        if (wantPoints) { out.pts1.resize((size_t)total); out.pts2.resize((size_t)total); }
    }

   private:
    void check(int rc) const { if (rc) throw Error(rc, sfmgms_multi_last_error(m_)); }
    sfmgms_multi* m_ = nullptr;
    const ImageSet* set_ = nullptr;
};

}  // namespace sfmgms

#ifdef SFMGMS_WITH_OPENCV
// Adaptor on the real OpenCV types: cv::KeyPoint / cv::DMatch are layout-identical to the PODs above.
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
namespace sfmgms {
inline void match(const ::cv::Mat& desc1, const ::cv::Mat& desc2, std::vector<::cv::DMatch>& matches) {
    CV_Assert(desc1.type() == desc2.type() && desc1.isContinuous() && desc2.isContinuous());
    std::vector<cv::DMatch> m;
    if (desc1.type() == CV_32F) {   // SIFT: BFMatcher::create() = NORM_L2, exactly the reference's FeatureMatchUtil.cpp:66-68
        CV_Assert(desc1.cols == desc2.cols && desc1.cols >= 1 && desc1.cols <= 256);
        cv::BFMatcher(cv::NORM_L2).match(desc1.ptr<float>(), desc1.rows, desc2.ptr<float>(), desc2.rows, m, desc1.cols);
    } else {
        CV_Assert(desc1.type() == CV_8U && desc1.cols == 32 && desc2.cols == 32);
        cv::BFMatcher().match(desc1.ptr<uint8_t>(), desc1.rows, desc2.ptr<uint8_t>(), desc2.rows, m);
    }
    matches.resize(m.size());
    if (!m.empty()) std::memcpy((void*)matches.data(), m.data(), m.size() * sizeof(cv::DMatch));
}
inline void matchGMS(const ::cv::Size& size1, const ::cv::Size& size2, const std::vector<::cv::KeyPoint>& kp1,
                     const std::vector<::cv::KeyPoint>& kp2, const std::vector<::cv::DMatch>& matches1to2,
                     std::vector<::cv::DMatch>& matchesGMS, bool withRotation = false, bool withScale = false,
                     double thresholdFactor = 6.0) {
    static_assert(sizeof(::cv::KeyPoint) == sizeof(cv::KeyPoint) && sizeof(::cv::DMatch) == sizeof(cv::DMatch), "layout");
    Context& c = Context::thread_default();
    std::vector<uint8_t> mask(matches1to2.size() + 1);
    int mask_len = 0, n_inl = 0;
    c.check(sfmgms_gms(c.get(), size1.width, size1.height, size2.width, size2.height, kp1.empty() ? nullptr : &kp1[0].pt.x,
                       (int)kp1.size(), (int)sizeof(::cv::KeyPoint), kp2.empty() ? nullptr : &kp2[0].pt.x, (int)kp2.size(),
                       (int)sizeof(::cv::KeyPoint), matches1to2.empty() ? nullptr : &matches1to2[0].queryIdx,
                       matches1to2.empty() ? nullptr : &matches1to2[0].trainIdx, (int)sizeof(::cv::DMatch),
                       (int)matches1to2.size(), withRotation, withScale, thresholdFactor, mask.data(), &mask_len, &n_inl,
                       nullptr));
    matchesGMS.clear();
    for (int i = 0; i < mask_len; ++i)
        if (mask[i]) matchesGMS.push_back(matches1to2[i]);
}
}  // namespace sfmgms
#endif
