// Device-side data layout shared by the host runtime (ctx.cu) and the kernels (kernels.cu).
// All per-pair arrays live in a few large HBM allocations; DevPair holds the offsets.
#ifndef LB200_DEV_TYPES_H
#define LB200_DEV_TYPES_H
#include <stdint.h>

// -inf encoding for 32-bit scores. The reference uses 64-bit scores with -inf = LONG_MIN/5*2 and
// only normalises on store (infty_int.hh:39-60, :374-388); only finiteness and finite values are
// observable.  Here: finite values satisfy |v| < 2^27, every "-inf-like" value stays below
// LB_NEG_LIMIT, and the sum of two of them does not overflow int32 (see DESIGN.md "-inf in 32 bits").
#define LB_NEG (-0x20000000)        // -2^29
#define LB_NEG_LIMIT (-0x10000000)  // anything below is -inf

#define LB_MAXLEN 4095              // positions are packed into 12 bits
#define LB_CODE_N 4                 // symbol codes: A C G U = 0..3, N = 4, up to three further symbols 5..7 (per process)
#define LB_NCODES 8

// scoring parameters that the kernels need (single sequences: position independent gap cost)
struct DevParams {
    int gap;          // gapA(i) = gapB(j) = indel - unpaired_penalty     (scoring.cc:272-311, :64-74)
    int gap_open;     // gap + indel_opening
    int open;         // indel_opening
    int exclusion;
    int no_lonely_pairs, struct_local, sequ_local;
    int stacking;     // Scoring::stacking(): stacked arc matches (scoring.hh:652-655)
    int fe_left1, fe_right1, fe_left2, fe_right2;  // free_endgaps.hh:42-70
    int sigma8[64];   // base match score by symbol codes (8x8), unpaired penalty applied (scoring.cc:141-198)
};

// one sequence-structure alignment problem
struct DevPair {
    int lenA, lenB;
    int codesA, codesB;   // offsets into codes[] (1-based position p at codes[off + p])
    int band;             // offset into band_lo[] / band_hi[] / cell_rev[] (entries 0..lenA)
    int sptr;             // offset into sptr[] (entries 0..lenA+lenB+2): S-order start per anti-diagonal ar+br
    int arcsA, arcsB;     // offsets into arc_left[] / arc_right[] / arc_weight[] (arcs in BasePairs index order)
    int lptrA, lptrB;     // offsets into lptr[] / lcount[] (entries 0..len+1): arcs with a given left end
    int n_cells;          // band cells (al, bl) with al >= 1, bl >= 1
    int K;                // number of arc matches                                       (filled by the device builder)
    long long cell_base;  // offset of this pair's cells in the cell arrays (cells ranked al desc, bl desc)
    long long am_base;    // offset of this pair's arc matches in the L-order / S-order arrays  (device builder)
    int anchored;         // 1: arc matches must join positions of equal anchor rank (acodes[], same offsets as codes[])
    int n_arcsB;          // arcs of B (row length of the profile arc-match table)
    // profile (multi-row) inputs: position-specific scoring (scoring.cc:141-198, :272-311, :369-438) in the gap-free frame, see
    // host_model.h ProfileTables. -1: single sequences, score tables by symbol code.
    long long ps_sig;     // offset into DevCtx::ps_sig: (lenA+1) x (lenB+1) base match scores sigma'(i, j)
    long long ps_am;      // offset into DevCtx::ps_am: arcs(A) x arcs(B) sequence terms of the arc-match scores
};

// per-pair counters produced by the device builder
struct DevPairStats {
    long long n_tasks;
    long long cells;      // DP cell updates: sum over D-fill tasks of |box and band| + top level band area
    long long terms;      // arc-match entries streamed by all boxes
    long long pad;
};

// S-order entry (16 bytes, one LDG.128): one valid arc match, sorted by (ar+br ascending, (al-1)+(bl-1) descending)
//   x = (al-1) | (bl-1) << 16         source cell of the recurrence M(al-1, bl-1) + D
//   y = ar | br << 16                 target cell
//   d = D(arcA, arcB)                 written by the D-fill task of the left ends (al, bl); -inf until then
//   s = (al-1) + (bl-1)               anti-diagonal of the source cell
struct DevEntry { uint32_t x, y; int d; int s; };
// Packed S-order entry (8 bytes, one LDG.64) for batches whose sequences are at most LB_PACK_MAXLEN long: the copy of the S-order
// the single-state sweep streams (half the bytes per anti-diagonal list). 9-bit positions, 28-bit D:
//   w0 = (al-1) | (bl-1) << 9 | ar << 18 | (br & 31) << 27        w1 = br >> 5 | D28 << 4
//   D28 = D for finite D (|D| < 2^27), LB_PACK_NEG for -inf.  s = (al-1) + (bl-1) is recomputed.
#define LB_PACK_MAXLEN 510
#define LB_PACK_NEG (-(1 << 27))
#define LB_PACK_W1(br, d) ((uint32_t)((br) >> 5) | ((uint32_t)(((d) < LB_NEG_LIMIT) ? LB_PACK_NEG : (d)) << 4))
#define LB_ENT_LO(v) ((int)((v) & 0xffffu))
#define LB_ENT_HI(v) ((int)((v) >> 16))

// L-order record: arc matches grouped by common left ends (al desc, bl desc, ar asc, br asc)
struct DevArcMatch {
    uint32_t ends_a;   // al | ar << 12
    uint32_t ends_b;   // bl | br << 12
    int score;         // Scoring::arcmatch(am) (scoring.cc:441-485)
    int spos;          // position in S-order (relative to am_base)
    int inner;         // L-order index of the inner arc match (al+1,ar-1,bl+1,br-1) or -1
    int score_st;      // Scoring::arcmatch(am, true) if both arcs are stackable (scoring.cc:556-569), else LB_NOSTACK
};
#define LB_NOSTACK (-0x7fffffff - 1)

// one D-fill task: all arc matches with common left ends (arc_matches.cc:313-355, aligner.cc:660-732)
struct DevTask {
    int pair;
    short al, bl;      // origin of the M box (no-lonely-pairs mode: inner left ends)
    short R, C;        // last row / column of the box (max right end - 1)
    int run_start;     // L-order range of the arc matches whose D entries this task defines
    int run_count;
};

// Row-grouped D fill (dfill_rows.cu): the D-fill tasks of one origin row al of a pair form a ROW; up to LB_GV of them (bl descending)
// are swept together as the layers of one box (one GROUP = one warp's work item). All groups of a row use the geometry of the row's
// union box (bl0r, Rr, Cr), which lets them share one filtered arc-match entry list. task[] indexes DevCtx::tasks.
#define LB_GV 4
struct DevGroup {
    int pair;
    short al, nmem;      // origin row, number of member tasks
    short gi, G;         // index of the group inside its row, number of groups of the row
    short bl0r, Rr, Cr;  // row union box: smallest origin column, last row, last column
    short pad;
    int task[LB_GV];
};

// a box whose optimal path still has to be traced (traceback kernel)
struct TraceJob { short al, bl, R, C; int am; };  // am: L-order index of the arc match the box belongs to (struct-local)

struct DevTopResult {
    int score;         // LB_NEG.. if -inf
    int max_i, max_j;
    int min_ij;        // traceback of a sequence-local top level: the cell it stopped at, min_i | min_j << 16 (aligner.cc:1251-1255)
};

#endif
