#include "orb_select.cuh"
#include <algorithm>
#include <vector>
#include <random>
#include <cstdio>
static int ref_retain(std::vector<RespIdx>& v, int n_points) {
    if (n_points >= 0 && v.size() > (size_t)n_points) {
        if (n_points == 0) { v.clear(); return 0; }
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const RespIdx& a, const RespIdx& b) { return a.r > b.r; });
        float amb = v[n_points - 1].r;
        auto e = std::partition(v.begin() + n_points, v.end(), [amb](const RespIdx& a) { return a.r >= amb; });
        v.resize(e - v.begin());
    }
    return (int)v.size();
}
int main() {
    std::mt19937 rng(1);
    long bad = 0, fallbacks = 0, cases = 0;
    for (int t = 0; t < 200000; ++t) {
        int n = 1 + rng() % (t % 50 == 0 ? 20000 : 300);
        int kinds = rng() % 4;
        std::vector<RespIdx> a(n);
        for (int i = 0; i < n; ++i) {
            float r;
            if (kinds == 0) r = (float)(rng() % 8);                 // heavy ties
            else if (kinds == 1) r = (float)(rng() % 235 + 20);     // FAST-like scores
            else if (kinds == 2) r = (float)rng() / 4294967296.f;   // distinct
            else r = (float)(i % 7 == 0 ? rng() % 3 : i);          // sorted-ish
            a[i] = RespIdx{r, i};
        }
        int np = rng() % (n + 3);
        std::vector<RespIdx> b = a, c = a;
        int m_ref = ref_retain(a, np);
        int m = orbsel::retain_best(b.data(), n, np);
        std::vector<unsigned short> Ls(n + 1), Rs(n + 1);
        int m2 = orbsel::retain_best_lists(c.data(), n, np, Ls.data(), Rs.data());     // the form the GPU warps run
        ++cases;
        if ((m < 0) != (m2 < 0)) { ++bad; continue; }
        if (m < 0) { ++fallbacks; continue; }
        if (m != m_ref || m2 != m_ref) { ++bad; continue; }
        for (int i = 0; i < m; ++i) if (a[i].i != b[i].i || a[i].i != c[i].i) { ++bad; break; }
    }
    printf("cases %ld bad %ld fallbacks %ld\n", cases, bad, fallbacks);
    return bad != 0;
}
