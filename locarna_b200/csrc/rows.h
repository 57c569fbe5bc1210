// Row-grouped D fill: kernel argument block and launch interface (dfill_rows.cu), group construction (builder.cu).
#ifndef LB200_ROWS_H
#define LB200_ROWS_H
#include <cuda_runtime.h>
#include "dev_ctx.h"
#include "dev_types.h"

namespace lb200 {

#define LB_ROWS_TG 264
#define LB_ROWS_RING 6   // anti-diagonals of arc-match accumulators per layer (dfill_rows.cu)   // groups of four target anti-diagonals of a box (sequences <= 510 nt)

struct RowsCtx {
    const DevGroup *groups;
    const unsigned *order;       // claim order: row al descending, larger boxes first
    const int *n_groups;         // device counter written by the builder
    const int *col_first;        // per pair at DevPair::sptr, entries 0..lenB: first / last row of the band that holds column j
    const int *col_last;         //   (first > last: empty column)
    const int *dep_need;         // [pair * n_levels + al]: groups of the pair in rows > al
    int *dep_done;               // per pair: completed groups
    int *row_built;              // per pair: groups that have finished their share of their row's entry list
    int n_levels;
    uint2 *clist;                // per pair clist_cap entries: the filtered entry list of the row in flight (idx | slot << 16, D)
    int *cnblk;                  // per pair LB_ROWS_TG ints: blocks of 32 entries per group of four target anti-diagonals
    long long clist_cap;
    int *scratch;                // per-CTA box area
    long long scratch_words;
    // shared memory layout of a one-warp CTA (bytes): [sigma 256][acc][gstart LB_ROWS_TG ints][colw][rowcode]
    int nc_max;                  // 1 or 2 columns per lane
    int force_nc;                // test knob (LB200_ROWS_FORCE_NC): use this many columns per lane wherever a box supports it
    int acc_words;               // RING * 32 * nc_max * LB_GV
    int colw_words;              // >= max lenB + 1 + 64 * nc_max
    int row_pad;                 // rowcode index of local row 0
    int rowcode_bytes;
};

int rows_smem_bytes(const RowsCtx &r);
cudaError_t rows_configure(int smem_bytes, int *ctas_per_sm);
void launch_dfill_rows(const DevCtx &c, const RowsCtx &r, int grid, int smem_bytes, int *cursor, cudaStream_t st);

}  // namespace lb200
#endif
