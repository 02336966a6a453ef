// issl_triple_tables.h -- the fixed tables of ISSL_LAYOUT_TRIPLE, shared by the host (issl_triple_visits) and the
// device code (issl_triple.cuh).
//
// kTripleLayout[t] = {L, M, H, p, q}: triple t is the slice set {L, M, H}; its buckets are keyed by
// slice L | slice M << 8 | slice H << 16, and an entry keeps slices p < q (the complement) as its residual.
// The byte order is chosen for locality: each triple is responsible for
//   * the hits that match the guide exactly on the pair {M, H} and on nothing else -- their buckets differ from the
//     guide's only in the LOW key byte, i.e. they lie among 256 consecutive buckets (32 KB of the blocked copy);
//   * (five of the triples) the hits that match exactly on slice H alone -- their buckets differ in the two low
//     bytes, an 8 MB window;
// so the ~1 400 reads of a guide fall into a few dozen DRAM pages instead of one page each.  Every pair of slices is
// {M, H} of exactly one triple (a perfect matching between the 10 pairs and the 10 triples).
#ifndef ISSL_TRIPLE_TABLES_H
#define ISSL_TRIPLE_TABLES_H

#include <stdint.h>

#ifdef __CUDACC__
#define ISSL_HD __host__ __device__
#else
#define ISSL_HD
#endif

#define ISSL_TRIPLE_LAYOUT_INIT                                                                                       \
    {{2, 1, 0, 3, 4}, {1, 0, 3, 2, 4}, {1, 0, 4, 2, 3}, {3, 0, 2, 1, 4}, {0, 2, 4, 1, 3},                              \
     {0, 3, 4, 1, 2}, {3, 2, 1, 0, 4}, {2, 1, 4, 0, 3}, {4, 1, 3, 0, 2}, {4, 2, 3, 0, 1}}

// resp(E): the triple responsible for a site whose set of exactly matching slices is E (bit s = slice s, E != 0):
//   |E| == 1: the triple whose H is that slice;  |E| == 2: the triple whose {M, H} is that pair;
//   |E| >= 3: the triple made of the three lowest slices of E.
ISSL_HD constexpr uint32_t issl_triple_resp(uint32_t E)
{
    const uint8_t lay[10][5] = ISSL_TRIPLE_LAYOUT_INIT;
    const uint8_t single[5] = {0, 6, 3, 1, 2};   // triple whose H is slice e (those whose pair {M, H} contains e as H)
    int n = 0, s0 = -1, s1 = -1, s2 = -1;
    for (int s = 0; s < 5; s++)
        if (E & (1u << s)) {
            if (n == 0) s0 = s; else if (n == 1) s1 = s; else if (n == 2) s2 = s;
            n++;
        }
    if (n == 0) return 15u;
    if (n == 1) return single[s0];
    for (uint32_t t = 0; t < 10; t++) {
        if (n == 2) {
            if ((lay[t][1] == s0 && lay[t][2] == s1) || (lay[t][1] == s1 && lay[t][2] == s0)) return t;
        } else {
            const uint32_t set = (1u << lay[t][0]) | (1u << lay[t][1]) | (1u << lay[t][2]);
            if (set == ((1u << s0) | (1u << s1) | (1u << s2))) return t;
        }
    }
    return 15u;
}

#endif
