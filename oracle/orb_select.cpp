// TEST INFRASTRUCTURE (CPU oracle) - never on the product path.
// KeyPointsFilter::retainBest of OpenCV (third party, opencv-python 4.13, features2d keypoint.cpp) restated on
// (response, index) pairs: std::nth_element on "response greater", boundary response, std::partition of the rest on
// "response >= boundary".  ORB's output ORDER is whatever permutation these two libstdc++ algorithms leave behind,
// so the oracle calls the same two algorithms.  Pinned against cv2.ORB_create(500).detectAndCompute
// (tests/test_oracle_orb.py).
#include <algorithm>
#include <vector>

namespace {
struct R { float r; int i; };
}

extern "C" int orb_retain_best(const float* resp, int n, int n_points, int* out) {
    std::vector<R> v(n);
    for (int i = 0; i < n; i++) v[i] = R{resp[i], i};
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const R& a, const R& b) { return a.r > b.r; });
        const float amb = v[n_points - 1].r;
        auto e = std::partition(v.begin() + n_points, v.end(), [amb](const R& a) { return a.r >= amb; });
        v.resize(e - v.begin());
    }
    for (size_t k = 0; k < v.size(); k++) out[k] = v[k].i;
    return (int)v.size();
}
