// Device builder interface (builder.cu): arc matches, S-order, tasks of a batch of pairs.
#ifndef LB200_BUILDER_H
#define LB200_BUILDER_H
#include <cuda_runtime.h>
#include <stddef.h>
#include "dev_types.h"

namespace lb200 {

struct BuildCtx {
    // inputs
    DevPair *pairs;
    const uint8_t *codes;
    const uint8_t *acodes;       // anchor rank per position (same offsets as codes); read only for pairs with DevPair::anchored
    const int *band_lo, *band_hi, *cell_rev;   // cell_rev[al] = number of band cells in rows > al
    const int *arc_left, *arc_right, *arc_weight, *lptr, *lcount;
    const int *arc_sdelta;                      // per arc: stack weight - weight, LB_NOSTACK if the arc is not stackable
    const int *am_seq;                          // 256: (tau * ribosum arc-match score) / 100
    const int *ps_am;                           // profile pairs (DevPair::ps_am >= 0): sequence term per (arc of A, arc of B)
    int sigma8[64];
    int tau, use_ribosum, no_lonely_pairs, struct_local, max_diff_am, max_diff_at_am;
    // outputs / work arrays
    int *cell_start;                            // total_cells + 1: counts, then exclusive offsets (global L-order index)
    DevArcMatch *am;
    DevEntry *ent;
    uint2 *ent8;                                // packed copy of the S-order (dev_types.h LB_PACK_*), nullptr for long sequences
    int *sptr;
    unsigned long long *skeys, *skeys_sorted;
    unsigned *svals, *svals_sorted;
    DevTask *tasks_unsorted, *tasks;
    unsigned *tkeys, *tkeys_sorted, *tvals, *tvals_sorted;
    unsigned *n_tasks;
    int *qstart;                                // 4098 entries
    DevPairStats *stats;
    // dependency-driven D-fill schedule (nullptr: not built). After builder_sort_tasks: tvals_sorted = task order grouped by blocks of
    // sb_pairs pairs, levcnt[pair * n_groups + g] = number of tasks of the pair in level groups > g
    int *levcnt;
    int n_groups, sb_pairs;
};

// Row groups of the D fill (dfill_rows.cu): the level-sorted tasks are ordered by (pair, al descending, bl descending), every run of
// up to LB_GV tasks of one row becomes a group; claim order = row descending, larger boxes first; levcnt[pair * n_levels + al] =
// number of groups of the pair in rows > al.
struct GroupBuild {
    const DevTask *tasks;
    unsigned n_tasks;
    int n_pairs;
    unsigned long long *keys, *keys_sorted;   // n_tasks
    unsigned *vals, *vals_sorted;             // n_tasks
    int *gid;                                 // n_tasks + 1: leader flags, then their exclusive scan
    DevGroup *groups;                         // n_groups (set for the fill stage)
    unsigned *gkeys, *gkeys_sorted, *gvals, *order;   // n_groups
    int *levcnt;
    int n_levels;
    int *n_groups;                            // device counter
};
size_t builder_groups_tmp_bytes(unsigned n_tasks, int n_pairs);
// stage 1: sort + leader scan; *n_groups (device) holds the group count afterwards
cudaError_t builder_groups_scan(const GroupBuild &g, void *tmp, size_t tmp_bytes, cudaStream_t st);
// stage 2 (groups allocated for the count read back): group records, claim order, per-row dependency counts
cudaError_t builder_groups_fill(const GroupBuild &g, unsigned n_groups, void *tmp, size_t tmp_bytes, cudaStream_t st);

cudaError_t builder_count(const BuildCtx &b, int n_pairs, long long total_cells, void *tmp, size_t tmp_bytes, size_t *tmp_need, cudaStream_t st);
size_t builder_sort_tmp_bytes(long long total_am, int n_pairs);
cudaError_t builder_fill(const BuildCtx &b, int n_pairs, long long total_am, long long sptr_total, void *tmp, size_t tmp_bytes, cudaStream_t st);
cudaError_t builder_sort_tasks(const BuildCtx &b, int n_pairs, unsigned n_tasks, void *tmp, size_t tmp_bytes, cudaStream_t st);

}  // namespace lb200
#endif
