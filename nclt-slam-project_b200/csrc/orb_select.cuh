// KeyPointsFilter::retainBest (OpenCV features2d) as straight-line code for one thread: the libstdc++ algorithms it
// calls - std::nth_element (introselect: median-of-three to the front, unguarded Hoare partition, insertion sort below
// four elements) and std::partition (bidirectional form) - restated statement for statement, because the permutation
// they leave behind IS ORB's keypoint order.  tests/test_orb_select.py replays random and tie-heavy inputs through
// this code (compiled for the host) and through the real std:: calls.  The heap-select fallback of introselect
// (recursion budget 2*log2(n) exhausted) is not restated: the function reports it and the caller falls back to the host.
#pragma once
#ifdef __CUDACC__
#define ORB_HD __host__ __device__ __forceinline__
#else
#define ORB_HD inline
#endif

struct RespIdx { float r; int i; };

namespace orbsel {

ORB_HD bool gt(const RespIdx& a, const RespIdx& b) { return a.r > b.r; }
ORB_HD void swp(RespIdx* a, RespIdx* b) { RespIdx t = *a; *a = *b; *b = t; }

ORB_HD void move_median_to_first(RespIdx* result, RespIdx* a, RespIdx* b, RespIdx* c) {
    if (gt(*a, *b)) {
        if (gt(*b, *c)) swp(result, b);
        else if (gt(*a, *c)) swp(result, c);
        else swp(result, a);
    } else if (gt(*a, *c)) swp(result, a);
    else if (gt(*b, *c)) swp(result, c);
    else swp(result, b);
}

ORB_HD RespIdx* unguarded_partition(RespIdx* first, RespIdx* last, RespIdx* pivot) {
    while (true) {
        while (gt(*first, *pivot)) ++first;
        --last;
        while (gt(*pivot, *last)) --last;
        if (!(first < last)) return first;
        swp(first, last);
        ++first;
    }
}

ORB_HD void insertion_sort(RespIdx* first, RespIdx* last) {
    if (first == last) return;
    for (RespIdx* i = first + 1; i != last; ++i) {
        if (gt(*i, *first)) {
            RespIdx val = *i;
            for (RespIdx* p = i; p != first; --p) *p = *(p - 1);
            *first = val;
        } else {
            RespIdx val = *i;
            RespIdx* l = i;
            RespIdx* next = i - 1;
            while (gt(val, *next)) {
                *l = *next;
                l = next;
                --next;
            }
            *l = val;
        }
    }
}

// std::nth_element(first, nth, last, greater-by-response); false = heap-select fallback needed (nothing usable done)
ORB_HD bool nth_element_desc(RespIdx* first, RespIdx* nth, RespIdx* last) {
    if (first == last || nth == last) return true;
    long n = last - first;
    int lg = 0;
    while (n > 1) { n >>= 1; ++lg; }
    int depth = 2 * lg;
    while (last - first > 3) {
        if (depth == 0) return false;
        --depth;
        RespIdx* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1);
        RespIdx* cut = unguarded_partition(first + 1, last, first);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(first, last);
    return true;
}

// std::partition(first, last, r >= thr), bidirectional form
ORB_HD RespIdx* partition_ge(RespIdx* first, RespIdx* last, float thr) {
    while (true) {
        while (true) {
            if (first == last) return first;
            else if (first->r >= thr) ++first;
            else break;
        }
        --last;
        while (true) {
            if (first == last) return first;
            else if (!(last->r >= thr)) --last;
            else break;
        }
        swp(first, last);
        ++first;
    }
}

// returns the number kept (the kept ones are v[0 .. ret)), or -1 when the fallback is needed
ORB_HD int retain_best(RespIdx* v, int n, int n_points) {
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        if (!nth_element_desc(v, v + n_points - 1, v + n)) return -1;
        const float ambiguous = v[n_points - 1].r;
        return (int)(partition_ge(v + n_points, v + n, ambiguous) - v);
    }
    return n;
}

}  // namespace orbsel
