// Host side of the all-vs-all stage: pair order, cost estimate and cost-balanced sharding of the pair list over GPUs.
// Reference contract: src/Utils/mlocarna:3547-3643 (compute_all_pairwise_alignments: pair (a, b) for a in 0..n-1, b in 0..a-1,
// A = the later sequence) and :2321-2344 (--compute-pairwise-scores k/N: process k of N computes its share of the pair list and
// writes a partial score list). The reference deals pairs round robin; here shares are balanced by estimated arc-match cost.
#include <algorithm>
#include <cstdint>
#include <functional>
#include <numeric>
#include <queue>
#include <vector>

#include "../../include/locarna_b200.h"

extern "C" {

int64_t lb200_all_vs_all(int n_seqs, int *seqA, int *seqB) {
    if (n_seqs < 0) return LB200_ERR_ARG;
    int64_t k = 0;
    for (int a = 0; a < n_seqs; a++)
        for (int b = 0; b < a; b++, k++)
            if (seqA && seqB) { seqA[k] = a; seqB[k] = b; }
    return k;
}

// candidate arc matches x band length: the D-fill work of a pair grows with the number of left-end pairs that carry arc matches
// and with the box each of them sweeps
double lb200_pair_cost(int n_arcsA, int n_arcsB, int lenA, int lenB) { return (double)n_arcsA * (double)n_arcsB * (double)(lenA + lenB); }

// Longest-processing-time-first: pairs by descending cost (ties by index) go to the rank with the smallest load (ties by rank).
// rank_of[k] = rank of pair k; order[] (optional) = pair indices grouped by rank, inside a rank by descending cost, with
// rank_begin[r] .. rank_begin[r+1] delimiting rank r (world + 1 entries).
int lb200_shard_pairs(int64_t n_pairs, const double *cost, int world, int *rank_of, int64_t *order, int64_t *rank_begin) {
    if (n_pairs < 0 || world < 1 || (n_pairs > 0 && (!cost || !rank_of))) return LB200_ERR_ARG;
    std::vector<int64_t> idx((size_t)n_pairs);
    std::iota(idx.begin(), idx.end(), (int64_t)0);
    std::sort(idx.begin(), idx.end(), [&](int64_t x, int64_t y) { return cost[x] != cost[y] ? cost[x] > cost[y] : x < y; });
    typedef std::pair<double, int> Load;
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int r = 0; r < world; r++) heap.push(Load(0.0, r));
    std::vector<int64_t> count((size_t)world + 1, 0);
    for (int64_t k : idx) {
        Load l = heap.top();
        heap.pop();
        rank_of[k] = l.second;
        count[l.second + 1]++;
        l.first += cost[k];
        heap.push(l);
    }
    for (int r = 0; r < world; r++) count[r + 1] += count[r];
    if (rank_begin) std::copy(count.begin(), count.end(), rank_begin);
    if (order) {
        std::vector<int64_t> fill(count.begin(), count.end() - 1);
        for (int64_t k : idx) order[fill[rank_of[k]]++] = k;
    }
    return LB200_OK;
}

}  // extern "C"
