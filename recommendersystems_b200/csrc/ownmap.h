// ownmap.h -- which rank holds the raw links of a source node in a partitioned build (graph.cu, synth.cu).  Plain C++, no
// CUDA headers: tests/test_ownmap.py compiles it with g++ and checks that the pieces cover every node exactly once.
#pragma once

#if defined(__CUDACC__)
#define RWR_HD __host__ __device__
#else
#define RWR_HD
#endif

// Partitioned build: which rank holds the raw links of a source node.  The node range is cut into up to three segments
// (the synthetic generator: users, items, third-party users -- their degrees differ by an order of magnitude, and ids
// inside a class are scrambled) and every segment is dealt evenly over the ranks in contiguous pieces.
struct OwnMap {
    int parts = 1, rank = 0, n_segs = 1;
    long long seg[5] = {0, 0, 0, 0, 0};       // segment k = [seg[k], seg[k + 1]), none empty
};
RWR_HD inline int own_rank(const OwnMap& m, long long i) {
    int k = 0;
    while (k + 1 < m.n_segs && i >= m.seg[k + 1]) k++;
    return (int)(((i - m.seg[k]) * m.parts) / (m.seg[k + 1] - m.seg[k]));
}

