// TEST INFRASTRUCTURE ONLY (oracle/).  KeyPointsFilter::retainBest (OpenCV features2d keypoint.cpp, called from
// orb.cpp computeKeyPoints) on (response, id) records: std::nth_element + std::partition.  The ORDER those two calls
// leave is the order OpenCV returns keypoints in, so the oracle runs the same library calls (libstdc++).
#include <algorithm>
#include <cstdint>
#include <vector>
struct Rec { float response; int32_t id; };
extern "C" int retain_best(float* response, int32_t* id, int n, int n_points) {
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    std::vector<Rec> v((size_t)n);
    for (int i = 0; i < n; ++i) v[i] = Rec{response[i], id[i]};
    std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const Rec& a, const Rec& b) { return a.response > b.response; });
    const float amb = v[n_points - 1].response;
    auto e = std::partition(v.begin() + n_points, v.end(), [amb](const Rec& a) { return a.response >= amb; });
    const int m = (int)(e - v.begin());
    for (int i = 0; i < m; ++i) { response[i] = v[i].response; id[i] = v[i].id; }
    return m;
}
