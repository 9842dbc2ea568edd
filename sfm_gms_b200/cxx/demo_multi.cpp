// demo_multi.cpp — an image sequence matched on several GPUs of one box from ONE C++ process (sfmgms::MultiGpuMatcher over
// sfmgms_multi_*: ncclCommInitAll, one broadcast of the set, one host thread per GPU, pair-sharded) and, for comparison, on
// one GPU (sfmgms::matchPairs).  The two must agree bit for bit; exit code 0 says they did.
//
// usage: demo_multi [n_gpus (0 = all)] [n_images] [keypoints per image] [withRotation] [withScale]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sfmgms.hpp"

using namespace sfmgms;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static float rndf() { return (float)((rnd() >> 40) * (1.0 / 16777216.0)); }

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 0;
    const int n_images = argc > 2 ? std::atoi(argv[2]) : 12;
    const int n_kp = argc > 3 ? std::atoi(argv[3]) : 4000;
    const bool rot = argc > 4 && std::atoi(argv[4]), sc = argc > 5 && std::atoi(argv[5]);
    const int W = 640, H = 480;
    // a landmark pool seen by every image through a sliding window (shifted positions, noisy descriptors); ragged sizes
    const int pool = n_kp * 2;
    std::vector<uint8_t> pool_desc((size_t)pool * 32);
    std::vector<float> pool_x((size_t)pool), pool_y((size_t)pool);
    for (auto& b : pool_desc) b = (uint8_t)rnd();
    for (int i = 0; i < pool; ++i) { pool_x[(size_t)i] = rndf() * 2.f * W; pool_y[(size_t)i] = rndf() * (H - 1); }
    ImageSet set;
    for (int k = 0; k < n_images; ++k) {
        const int n = (k % 5 == 3) ? n_kp / 3 : n_kp - 17 * k;
        const float x0 = (float)k * 0.02f * 2.f * W;
        std::vector<cv::KeyPoint> kps;
        std::vector<uint8_t> desc;
        for (int i = 0; i < pool && (int)kps.size() < n; ++i) {
            const float x = pool_x[(size_t)i] - x0;
            if (x < 0.f || x >= (float)W - 1.f) continue;
            kps.push_back(cv::KeyPoint{{x + rndf() * 0.5f, pool_y[(size_t)i]}, 31.f, -1.f, 0.f, 0, -1});
            for (int b = 0; b < 32; ++b) {
                uint8_t v = pool_desc[(size_t)i * 32 + b];
                if ((rnd() & 7) == 0) v ^= (uint8_t)(1u << (rnd() & 7));
                desc.push_back(v);
            }
        }
        set.add(kps, desc.data(), cv::Size(W, H));
    }
    std::vector<std::pair<int, int>> pairs;
    for (int i = 0; i < n_images; ++i)
        for (int j = i + 1; j < n_images; ++j) pairs.push_back({i, j});
    try {
        Context one(0);
        PairMatches a, b;
        auto t0 = std::chrono::steady_clock::now();
        matchPairs(one, set, pairs, a, rot, sc);
        auto t1 = std::chrono::steady_clock::now();
        std::vector<int> devs;
        for (int g = 0; g < n_gpus; ++g) devs.push_back(g);
        MultiGpuMatcher multi(devs);
        multi.setImages(set);
        multi.matchPairs(pairs, b, rot, sc);                    // first call: warms NCCL / allocations
        auto t2 = std::chrono::steady_clock::now();
        multi.matchPairs(pairs, b, rot, sc);
        auto t3 = std::chrono::steady_clock::now();
        long long total = 0;
        int bad = 0;
        for (size_t p = 0; p < pairs.size(); ++p) {
            total += a.n_inliers[p];
            if (a.n_inliers[p] != b.n_inliers[p] || a.best_hyp[p] != b.best_hyp[p]) { ++bad; continue; }
            const size_t n = (size_t)a.n_inliers[p];
            if (n && (std::memcmp(&a.matches[(size_t)a.begin[p]], &b.matches[(size_t)b.begin[p]], n * sizeof(cv::DMatch)) ||
                      std::memcmp(&a.pts1[(size_t)a.begin[p]], &b.pts1[(size_t)b.begin[p]], n * sizeof(cv::Point2f)) ||
                      std::memcmp(&a.pts2[(size_t)a.begin[p]], &b.pts2[(size_t)b.begin[p]], n * sizeof(cv::Point2f))))
                ++bad;
        }
        auto ms = [](auto x, auto y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
        std::printf("{\"gpus\": %d, \"images\": %d, \"pairs\": %zu, \"inliers\": %lld, \"mismatching_pairs\": %d, \"one_gpu_ms\": %.2f, "
                    "\"multi_gpu_ms\": %.2f, \"broadcast_ms\": %.3f, \"rows_multi\": %zu, \"rows_one\": %zu}\n",
                    multi.deviceCount(), n_images, pairs.size(), total, bad, ms(t0, t1), ms(t2, t3), multi.lastBroadcastMs(),
                    b.matches.size(), a.matches.size());
        return (bad == 0 && a.matches.size() == b.matches.size()) ? 0 : 1;
    } catch (const Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
