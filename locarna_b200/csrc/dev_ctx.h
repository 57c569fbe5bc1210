// Kernel argument block: raw device pointers into the batch arrays (see dev_types.h).
#ifndef LB200_DEV_CTX_H
#define LB200_DEV_CTX_H
#include "dev_types.h"

#define LB_MAX_NC 16      // diagonal pairs per lane: one warp covers up to 64*16 = 1024 band diagonals

struct DevCtx {
    DevParams params;
    const DevPair *pairs;
    const uint8_t *codes;
    const int *ps_sig;       // position-specific base match scores of profile pairs (DevPair::ps_sig), nullptr: none
    const int *band_lo, *band_hi;
    const int *sptr;
    DevEntry *ent;           // S-order; ent[k].d = D(arcA,arcB)
    uint2 *ent8;             // packed copy streamed by the single-state sweep (nullptr: stream ent); D is kept in both
    const DevArcMatch *am;   // L-order
    const DevTask *tasks;
    const int *qstart;       // task range of level group q = 4095 - ((al+bl)>>1): tasks[qstart[q] .. qstart[q+1])
    int *cursor;             // per level group work cursor (4097 entries) + 1 for the top level
    DevTopResult *top;
    int *scratch;            // per-CTA M box spill area (L2 resident)
    int scratch_words;
    int max_rows;            // largest lenA + 1 of the batch (rowrange holds max_rows + 2 words)
    int rowcode_bytes;       // >= max_rows + 2, multiple of 4
    int colcode_bytes;       // >= largest lenB + 2, multiple of 4
    int arcbuf_words;        // RING * 32 * NCmax
    // padded row / column tables of the single-state sweep (kernels.cu setup_box2): no index clamps in the cell loop
    int row_words, row_pad;  // roww[row_pad + 1 + ip], ip = -1..Rn+1; everything else holds the invalid-row sentinel
    int col_bytes, col_pad;  // colc[col_pad + jp]
    int region_bytes;        // max of the two layouts (they share the same shared-memory region)
    // traceback
    const unsigned *lpos;    // S-order position -> L-order index (relative to the pair)
    int *trace_edges;        // per pair n+m+3 slots (same offsets as sptr): edge at slot i+j = i << 2 | kind
    char *trace_str;         // per pair: structure string of A (n+1 bytes) then of B (m+1 bytes), at the sptr offset
    TraceJob *trace_stack;   // per pair trace_stack_cap pending boxes
    int trace_stack_cap;
    int *error_flag;
    // dependency-driven D fill (dfill_dep_kernel): one persistent launch per batch instead of one launch per level group
    const unsigned *dep_order;   // task indices grouped by pair blocks, inside a block level descending (builder.cu dep_prepare_kernel)
    const int *dep_need;         // [pair * n_groups + level group]: completed tasks of the pair a task of that group waits for
    int *dep_done;               // per pair: completed tasks
    int n_groups, n_tasks;
    // normalized / penalized alignment (aligner.cc:1522-1622): the TOP LEVEL box is filled and traced with the scoring modified by
    // lambda (sigma - 2 lambda, gap - lambda, D - lambda * arc lengths: aligner_impl.hh:190-275, scoring.cc:77-90); the boxes of the arc
    // matches on the path keep the unmodified scoring
    // AlignerRestriction (aligner_restriction.hh:26-130; k-best alignment, aligner.cc:1383-1514): the top level covers only
    // rows r_sa..r_ea and columns r_sb..r_eb (box origin (r_sa - 1, r_sb - 1)); set for single-pair launches only
    int r_on, r_sa, r_sb, r_ea, r_eb;
    int use_tl;                  // 1: the traceback's top level uses params_tl / ent_tl / ent8_tl
    DevParams params_tl;
    const DevEntry *ent_tl;
    const uint2 *ent8_tl;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, per-function MAXIMUM; contexts on several host threads configure the
// same kernels with different sizes, so a family only ever raises its value (never below what another context is about to launch with).
#include <map>
#include <mutex>
#include <utility>
inline int lb200_sticky_smem(int family, int smem_bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, int> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    int &m = seen[std::make_pair(dev, family)];
    if (smem_bytes > m) m = smem_bytes;
    return m;
}

#endif
