// stdsort_port.h — control-flow-exact restatement of libstdc++ 13's std::sort (introsort) for arrays of
// packed 64-bit elements, usable from host C++ and CUDA device code.
//
// Why: the reference's quadtree sorts its expandable nodes with
//   sort(v.begin(), v.end(), compareNodes)            (/root/reference/src/ORBextractor.cc:700)
// where compareNodes (:538-553) orders by (point count, UL.x) only.  Ties are frequent and std::sort is
// unstable, so the processing order — and, at the nFeatures cut-off, the selected keypoint SET —
// depends on the exact permutation libstdc++ produces.  The oracle uses the
// real std::sort; the device must reproduce the same permutation from the same comparison results.
//
// Element layout: bits [63:24] = sort key (here size<<16 | ULx), bits [23:0] = payload (node index),
// less(a,b) := (a >> 24) < (b >> 24).  Payload never takes part in comparisons.
//
// Follows /usr/include/c++/13/bits/stl_algo.h: __sort :1939, __introsort_loop :1918,
// __unguarded_partition_pivot :1893, __move_median_to_first :85, __unguarded_partition :1871,
// __final_insertion_sort :1854, __insertion_sort :1812, __unguarded_linear_insert :1792,
// __partial_sort :1905 (heap fallback; stl_heap.h __make_heap :340, __adjust_heap :224,
// __push_heap :135, __pop_heap :254, __sort_heap :419).
#ifndef ORBX_STDSORT_PORT_H
#define ORBX_STDSORT_PORT_H
#include <stdint.h>

#if defined(__CUDACC__)
#define ORBX_SORT_HD __host__ __device__ inline
#else
#define ORBX_SORT_HD static inline
#endif

namespace orbx_sort {

typedef unsigned long long elem_t;
enum { kPayloadBits = 24, kThreshold = 16 };

ORBX_SORT_HD bool less(elem_t a, elem_t b) { return (a >> kPayloadBits) < (b >> kPayloadBits); }
ORBX_SORT_HD void swp(elem_t *a, int i, int j) { elem_t t = a[i]; a[i] = a[j]; a[j] = t; }

ORBX_SORT_HD void unguarded_linear_insert(elem_t *a, int last) {
    const elem_t val = a[last];
    int next = last - 1;
    while (less(val, a[next])) {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = val;
}

ORBX_SORT_HD void insertion_sort(elem_t *a, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(a[i], a[first])) {
            const elem_t val = a[i];
            for (int k = i; k > first; --k) a[k] = a[k - 1];  // move_backward(first, i, i+1)
            a[first] = val;
        } else {
            unguarded_linear_insert(a, i);
        }
    }
}

// heap helpers operate on the sub-array starting at a[base]
ORBX_SORT_HD void push_heap(elem_t *a, int base, int hole, int top, elem_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && less(a[base + parent], value)) {
        a[base + hole] = a[base + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[base + hole] = value;
}

ORBX_SORT_HD void adjust_heap(elem_t *a, int base, int hole, int len, elem_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(a[base + child], a[base + child - 1])) --child;
        a[base + hole] = a[base + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[base + hole] = a[base + child - 1];
        hole = child - 1;
    }
    push_heap(a, base, hole, top, value);
}

// __partial_sort(first, last, last) == make_heap + sort_heap (the heap_select loop is empty)
ORBX_SORT_HD void heap_sort(elem_t *a, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const elem_t value = a[first + parent];
            adjust_heap(a, first, parent, len, value);
            if (parent == 0) break;
            --parent;
        }
    }
    int end = last;
    while (end - first > 1) {
        --end;
        const elem_t value = a[end];  // __pop_heap(first, end, end)
        a[end] = a[first];
        adjust_heap(a, first, 0, end - first, value);
    }
}

ORBX_SORT_HD void move_median_to_first(elem_t *a, int result, int x, int y, int z) {
    if (less(a[x], a[y])) {
        if (less(a[y], a[z])) swp(a, result, y);
        else if (less(a[x], a[z])) swp(a, result, z);
        else swp(a, result, x);
    } else if (less(a[x], a[z])) swp(a, result, x);
    else if (less(a[y], a[z])) swp(a, result, z);
    else swp(a, result, y);
}

ORBX_SORT_HD int unguarded_partition(elem_t *a, int first, int last, int pivot) {
    const elem_t pv = a[pivot];  // the pivot slot lies outside [first, last) and is never written here
    while (true) {
        elem_t x = a[first];
        while (less(x, pv)) x = a[++first];
        --last;
        elem_t y = a[last];
        while (less(pv, y)) y = a[--last];
        if (!(first < last)) return first;
        a[first] = y;
        a[last] = x;
        ++first;
    }
}

// __introsort_loop only: the partitioning phase (explicit stack replaces the recursion on the right partition).
// What std::sort does afterwards, __final_insertion_sort, is a STABLE sort of whatever order this phase leaves
// (insertion with strict '<' never reorders equal keys), so callers may finish with any stable sort by key —
// the quadtree kernel does that part in parallel.
ORBX_SORT_HD void introsort_loop_only(elem_t *a, int n) {
    if (n <= kThreshold) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;  // std::__lg(n)
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > kThreshold) {
            if (depth == 0) {
                heap_sort(a, first, last);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            move_median_to_first(a, first, first + 1, mid, last - 1);
            const int cut = unguarded_partition(a, first + 1, last, first);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
            last = cut;
        }
    }
}

// std::sort(a, a+n) under less(); explicit stack replaces the recursion on the right partition
ORBX_SORT_HD void sort(elem_t *a, int n) {
    if (n <= 0) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;  // std::__lg(n)
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > kThreshold) {
            if (depth == 0) {
                heap_sort(a, first, last);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            move_median_to_first(a, first, first + 1, mid, last - 1);
            const int cut = unguarded_partition(a, first + 1, last, first);
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > kThreshold) {
        insertion_sort(a, 0, kThreshold);
        for (int i = kThreshold; i != n; ++i) unguarded_linear_insert(a, i);
    } else {
        insertion_sort(a, 0, n);
    }
}

}  // namespace orbx_sort
#endif
