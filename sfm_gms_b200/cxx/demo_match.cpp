// demo_match.cpp — the reference's matching stage (FeatureMatchUtil.cpp:52-84, SIFT_matchGMS) written against
// sfmgms.hpp: BFMatcher::match + matchGMS on one image pair read from a small binary fixture.
//
// usage: demo_match <pair.bin> <withRotation 0/1> <withScale 0/1> [out.bin]
// pair.bin : int32 w1,h1,w2,h2,n1,n2 | float kp1[n1*2] | float kp2[n2*2] | uint8 desc1[n1*32] | uint8 desc2[n2*32]
// out.bin  : int32 n_matches | int32 trainIdx[n] | int32 dist[n] | int32 n_gms | int32 gms_queryIdx[n_gms]
//            | int32 n_gms_class | uint8 mask_class[n_matches] (gms_matcher::GetInlierMask, same flags)
//            | int32 n_fused (matchBFHammingGMS, must equal n_gms)
//            | int32 n_bf | int32 bf_queryIdx[n_bf] | int32 bf_trainIdx[n_bf]   (bruteForceMatch, FeatureMatchUtil.cpp:20-31)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sfmgms.hpp"

using namespace sfmgms;

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s pair.bin withRotation withScale [out.bin]\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("open"); return 2; }
    int32_t hdr[6];
    if (std::fread(hdr, 4, 6, f) != 6) return 2;
    const int w1 = hdr[0], h1 = hdr[1], w2 = hdr[2], h2 = hdr[3], n1 = hdr[4], n2 = hdr[5];
    std::vector<float> xy1((size_t)n1 * 2), xy2((size_t)n2 * 2);
    std::vector<uint8_t> d1((size_t)n1 * 32), d2((size_t)n2 * 32);
    if (std::fread(xy1.data(), 4, xy1.size(), f) != xy1.size() || std::fread(xy2.data(), 4, xy2.size(), f) != xy2.size() ||
        std::fread(d1.data(), 1, d1.size(), f) != d1.size() || std::fread(d2.data(), 1, d2.size(), f) != d2.size()) return 2;
    std::fclose(f);
    const bool rot = std::atoi(argv[2]) != 0, sc = std::atoi(argv[3]) != 0;

    std::vector<cv::KeyPoint> kpts1((size_t)n1), kpts2((size_t)n2);
    for (int i = 0; i < n1; ++i) kpts1[i] = cv::KeyPoint{{xy1[2 * i], xy1[2 * i + 1]}, 31.f, -1.f, 0.f, 0, -1};
    for (int i = 0; i < n2; ++i) kpts2[i] = cv::KeyPoint{{xy2[2 * i], xy2[2 * i + 1]}, 31.f, -1.f, 0.f, 0, -1};

    try {
        // ---- the two lines of FeatureMatchUtil.cpp:66-69 ----
        auto t0 = std::chrono::steady_clock::now();
        cv::BFMatcher matcherBF(cv::NORM_HAMMING);
        std::vector<cv::DMatch> matches, matchesGMS;
        matcherBF.match(d1.data(), n1, d2.data(), n2, matches);
        cv::xfeatures2d::matchGMS(cv::Size(w1, h1), cv::Size(w2, h2), kpts1, kpts2, matches, matchesGMS, rot, sc);
        auto t1 = std::chrono::steady_clock::now();
        // upstream class API (scale first, rotation second)
        std::vector<bool> vbInliers;
        gms_matcher gms(kpts1, cv::Size(w1, h1), kpts2, cv::Size(w2, h2), matches);
        const int n_class = gms.GetInlierMask(vbInliers, sc, rot);
        // fused one-call form
        std::vector<cv::DMatch> m2, g2;
        matchBFHammingGMS(d1.data(), d2.data(), kpts1, kpts2, cv::Size(w1, h1), cv::Size(w2, h2), m2, g2, rot, sc);
        // the reference's brute-force helper: cross-checked match, sort, ratio prune, cap 500
        std::vector<cv::DMatch> mbf;
        bruteForceMatch(d1.data(), n1, d2.data(), n2, mbf);
        std::printf("matches %zu  GMS %zu  class %d  fused %zu  (%.3f s for BF+GMS incl. first-call setup)\n", matches.size(),
                    matchesGMS.size(), n_class, g2.size(), std::chrono::duration<double>(t1 - t0).count());
        if (argc > 4) {
            FILE* o = std::fopen(argv[4], "wb");
            int32_t n = (int32_t)matches.size();
            std::fwrite(&n, 4, 1, o);
            for (auto& m : matches) std::fwrite(&m.trainIdx, 4, 1, o);
            for (auto& m : matches) { int32_t d = (int32_t)m.distance; std::fwrite(&d, 4, 1, o); }
            int32_t g = (int32_t)matchesGMS.size();
            std::fwrite(&g, 4, 1, o);
            for (auto& m : matchesGMS) std::fwrite(&m.queryIdx, 4, 1, o);
            int32_t nc = n_class;
            std::fwrite(&nc, 4, 1, o);
            for (int i = 0; i < n; ++i) { uint8_t b = (i < (int)vbInliers.size() && vbInliers[i]) ? 1 : 0; std::fwrite(&b, 1, 1, o); }
            int32_t nf = (int32_t)g2.size();
            std::fwrite(&nf, 4, 1, o);
            int32_t nb = (int32_t)mbf.size();
            std::fwrite(&nb, 4, 1, o);
            for (auto& m : mbf) std::fwrite(&m.queryIdx, 4, 1, o);
            for (auto& m : mbf) std::fwrite(&m.trainIdx, 4, 1, o);
            std::fclose(o);
        }
    } catch (const Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
