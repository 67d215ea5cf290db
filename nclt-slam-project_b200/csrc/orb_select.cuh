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


// ---- the same two partitions without the sequential pointer walk ----
// Both loops are "find the next stopper from the left, the next from the right, swap, repeat".  With
//   L = ascending positions in [lo, hi) that stop the left pointer, R = descending positions that stop the right one,
// the sequential walk swaps exactly the pairs (L[k], R[k]) with L[k] < R[k] (a prefix k < K, the pairs are disjoint),
// because each pointer only ever crosses elements the other has not touched yet.  What the walk returns:
//   Hoare (unguarded_partition, stoppers: r <= pivot from the left, r >= pivot from the right):
//       K == 0: L[0];   else L[K] if it exists and lies before R[K-1], else R[K-1] (which now holds a left stopper)
//   std::partition (stoppers: !pred from the left, pred from the right): lo + number of elements satisfying pred.
// `lists` versions below are the plain C++ statement of that (tests/orb_select_check.cpp replays them against std::);
// the warp versions build L and R with ballots, 32 positions per step, and swap the K pairs in parallel.
template <typename StopL, typename StopR>
ORB_HD int stopper_lists(const RespIdx* v, int lo, int hi, StopL sl, StopR sr, unsigned short* Ls, unsigned short* Rs, int& nL, int& nR) {
    nL = nR = 0;
    for (int i = lo; i < hi; ++i)
        if (sl(v[i].r)) Ls[nL++] = (unsigned short)i;
    for (int j = hi - 1; j >= lo; --j)
        if (sr(v[j].r)) Rs[nR++] = (unsigned short)j;
    int K = 0;
    while (K < nL && K < nR && Ls[K] < Rs[K]) ++K;
    return K;
}

// unguarded Hoare partition of v[lo, hi) around pivot value P by lists; returns the cut index
ORB_HD int hoare_by_lists(RespIdx* v, int lo, int hi, float P, unsigned short* Ls, unsigned short* Rs) {
    int nL, nR;
    const int K = stopper_lists(v, lo, hi, [P](float r) { return !(r > P); }, [P](float r) { return !(P > r); }, Ls, Rs, nL, nR);
    for (int k = 0; k < K; ++k) swp(v + Ls[k], v + Rs[k]);
    if (K == 0) return Ls[0];
    return (nL > K && Ls[K] < Rs[K - 1]) ? Ls[K] : Rs[K - 1];
}

ORB_HD int partition_ge_by_lists(RespIdx* v, int lo, int hi, float thr, unsigned short* Ls, unsigned short* Rs) {
    int nL, nR;
    const int K = stopper_lists(v, lo, hi, [thr](float r) { return !(r >= thr); }, [thr](float r) { return r >= thr; }, Ls, Rs, nL, nR);
    for (int k = 0; k < K; ++k) swp(v + Ls[k], v + Rs[k]);
    return lo + nR;
}

// retain_best with the list-based partitions (same results as retain_best above; n <= 65535)
ORB_HD int retain_best_lists(RespIdx* v, int n, int n_points, unsigned short* Ls, unsigned short* Rs) {
    if (!(n_points >= 0 && n > n_points)) return n;
    if (n_points == 0) return 0;
    int first = 0, last = n;
    const int nth = n_points - 1;
    int lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg;
    int depth = 2 * lg;
    while (last - first > 3) {
        if (depth == 0) return -1;
        --depth;
        move_median_to_first(v + first, v + first + 1, v + first + (last - first) / 2, v + last - 1);
        const int cut = hoare_by_lists(v, first + 1, last, v[first].r, Ls, Rs);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    insertion_sort(v + first, v + last);
    return partition_ge_by_lists(v, n_points, n, v[n_points - 1].r, Ls, Rs);
}

#ifdef __CUDACC__
// Warp-cooperative forms: every lane calls with the same arguments; v, Ls, Rs in shared or global memory.
template <typename StopL, typename StopR>
__device__ __forceinline__ int warp_stopper_lists(const RespIdx* v, int lo, int hi, StopL sl, StopR sr, unsigned short* Ls,
                                                  unsigned short* Rs, int& nL, int& nR) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    nL = nR = 0;
    for (int b = lo; b < hi; b += 32) {
        const int i = b + lane;
        const bool f = i < hi && sl(v[i].r);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
        if (f) Ls[nL + __popc(m & lt)] = (unsigned short)i;
        nL += __popc(m);
    }
    for (int b = hi - 1; b >= lo; b -= 32) {
        const int j = b - lane;
        const bool f = j >= lo && sr(v[j].r);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, f);
        if (f) Rs[nR + __popc(m & lt)] = (unsigned short)j;
        nR += __popc(m);
    }
    __syncwarp();
    const int lim = nL < nR ? nL : nR;
    int K = 0;
    for (int b = 0; b < lim; b += 32) {
        const int k = b + lane;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, k < lim && Ls[k] < Rs[k]);
        K += __popc(m);                      // the predicate holds on a prefix (L ascends, R descends)
        if (m != 0xFFFFFFFFu) break;
    }
    return K;
}

__device__ __forceinline__ void warp_swap_pairs(RespIdx* v, const unsigned short* Ls, const unsigned short* Rs, int K) {
    for (int k = threadIdx.x & 31; k < K; k += 32) swp(v + Ls[k], v + Rs[k]);
    __syncwarp();
}

__device__ __forceinline__ int warp_retain_best(RespIdx* v, int n, int n_points, unsigned short* Ls, unsigned short* Rs) {
    if (!(n_points >= 0 && n > n_points)) return n;
    if (n_points == 0) return 0;
    const int lane = threadIdx.x & 31;
    int first = 0, last = n;
    const int nth = n_points - 1;
    int lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg;
    int depth = 2 * lg;
    while (last - first > 3) {
        if (depth == 0) return -1;
        --depth;
        if (lane == 0) move_median_to_first(v + first, v + first + 1, v + first + (last - first) / 2, v + last - 1);
        __syncwarp();
        const float P = v[first].r;
        int nL, nR;
        const int K = warp_stopper_lists(v, first + 1, last, [P](float r) { return !(r > P); }, [P](float r) { return !(P > r); }, Ls, Rs, nL, nR);
        const int cut = K == 0 ? Ls[0] : ((nL > K && Ls[K] < Rs[K - 1]) ? Ls[K] : Rs[K - 1]);
        __syncwarp();
        warp_swap_pairs(v, Ls, Rs, K);
        if (cut <= nth) first = cut;
        else last = cut;
    }
    if (lane == 0) insertion_sort(v + first, v + last);
    __syncwarp();
    const float thr = v[n_points - 1].r;
    int nL, nR;
    const int K = warp_stopper_lists(v, n_points, n, [thr](float r) { return !(r >= thr); }, [thr](float r) { return r >= thr; }, Ls, Rs, nL, nR);
    warp_swap_pairs(v, Ls, Rs, K);
    return n_points + nR;
}
#endif

}  // namespace orbsel
