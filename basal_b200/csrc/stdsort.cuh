// stdsort.cuh — std::sort as libstdc++ (GCC 13, bits/stl_algo.h / stl_heap.h) performs it, restated over an index array.
//
// SortHits4PE (align.cpp:412-416) sorts every level list with std::sort and a comparator on (chr, loc) only. A gapped and an
// ungapped hit at the same place compare equal, and with more than 16 elements introsort does not keep equal elements in
// input order (SURVEY trap 10) — GetPairs then enumerates the pairs in whatever order came out, which decides the -S pick and
// the -r 2 listing. To be bit-exact the same permutation has to come out here: this file follows the library's control flow
// step by step (median-of-three to the front, unguarded Hoare partition, recursion on the right part, heap sort when the
// depth limit 2 * floor(log2 n) runs out, insertion sort over the 16-element pieces at the end). tests/test_stdsort.py
// compares it with the real std::sort on the host for arrays full of equal keys.
#pragma once
#include <stdint.h>

#ifndef HD
#define HD __host__ __device__ __forceinline__
#endif

// p[0..n) = indices to sort; less(a, b) compares the elements behind two indices. stack: 64 ints of scratch.
template <typename Less>
HD void stdsort_adjust_heap(uint16_t *p, int first, int hole, int len, uint16_t value, Less less) {
    const int top = hole; int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(p[first + child], p[first + child - 1])) child--;
        p[first + hole] = p[first + child]; hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); p[first + hole] = p[first + child - 1]; hole = child - 1; }
    int parent = (hole - 1) / 2;                                            // __push_heap
    while (hole > top && less(p[first + parent], value)) { p[first + hole] = p[first + parent]; hole = parent; parent = (hole - 1) / 2; }
    p[first + hole] = value;
}
template <typename Less>
HD void stdsort_heapsort(uint16_t *p, int first, int last, Less less) {      // __partial_sort(first, last, last)
    const int len = last - first;
    if (len >= 2) for (int parent = (len - 2) / 2;; parent--) { stdsort_adjust_heap(p, first, parent, len, p[first + parent], less); if (parent == 0) break; }
    for (int l = last; l - first > 1;) { --l; const uint16_t v = p[l]; p[l] = p[first]; stdsort_adjust_heap(p, first, 0, l - first, v, less); }
}
template <typename Less>
HD void stdsort_linear_insert(uint16_t *p, int last, Less less) {             // __unguarded_linear_insert
    const uint16_t v = p[last]; int next = last - 1;
    while (less(v, p[next])) { p[last] = p[next]; last = next; --next; }
    p[last] = v;
}
template <typename Less>
HD void stdsort_insertion(uint16_t *p, int first, int last, Less less) {      // __insertion_sort
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(p[i], p[first])) { const uint16_t v = p[i]; for (int k = i; k > first; k--) p[k] = p[k - 1]; p[first] = v; }
        else stdsort_linear_insert(p, i, less);
    }
}
template <typename Less>
HD void stdsort(uint16_t *p, int n, Less less) {
    if (n <= 1) return;
    int depth = 0; for (int m = n; m > 1; m >>= 1) depth++; depth *= 2;      // 2 * __lg(n)
    // __introsort_loop: loop on the left part, "recursion" on the right part through an explicit stack
    int sf[64], sl[64], sd[64], sp = 0;
    int first = 0, last = n, d = depth;
    for (;;) {
        while (last - first > 16) {
            if (d == 0) { stdsort_heapsort(p, first, last, less); break; }
            --d;
            // __unguarded_partition_pivot
            const int mid = first + (last - first) / 2, a = first + 1, b = mid, c = last - 1;
            int med;                                                        // __move_median_to_first(first, a, b, c)
            if (less(p[a], p[b])) { if (less(p[b], p[c])) med = b; else if (less(p[a], p[c])) med = c; else med = a; }
            else if (less(p[a], p[c])) med = a; else if (less(p[b], p[c])) med = c; else med = b;
            { const uint16_t t = p[first]; p[first] = p[med]; p[med] = t; }
            int lo = first + 1, hi = last;                                  // __unguarded_partition(first + 1, last, first)
            for (;;) {
                while (less(p[lo], p[first])) ++lo;
                --hi;
                while (less(p[first], p[hi])) --hi;
                if (!(lo < hi)) break;
                const uint16_t t = p[lo]; p[lo] = p[hi]; p[hi] = t;
                ++lo;
            }
            sf[sp] = lo; sl[sp] = last; sd[sp] = d; sp++;                   // __introsort_loop(cut, last, depth_limit) comes first ...
            // ... but runs to completion before the left part continues; emulate the order: process the right part now
            // (order of the two parts does not matter: they are disjoint ranges)
            last = lo;
        }
        if (sp == 0) break;
        sp--; first = sf[sp]; last = sl[sp]; d = sd[sp];
    }
    // __final_insertion_sort
    if (n > 16) { stdsort_insertion(p, 0, 16, less); for (int i = 16; i != n; ++i) stdsort_linear_insert(p, i, less); }
    else stdsort_insertion(p, 0, n, less);
}
