/* locarna_b200 -- C ABI of the B200-native pairwise sequence-structure alignment path.
 *
 * This is the drop-in boundary for the hot path of LocARNA 2.0.1 (paths below are relative to the
 * reference tree, /root/reference):
 *
 *   lb200_seq_add_pp / lb200_seq_add   replace  RnaData(file, min_prob, ...)       src/LocARNA/rna_data.cc:45-68, :984-1103
 *                                      and      BasePairs(rna_data, min_prob)      src/LocARNA/basepairs.cc:68-77, :155-205
 *   lb200_pair_add (band == NULL)      replaces TraceController(seqA, seqB, ..)    src/LocARNA/trace_controller.cc:406-539
 *                                      and      restrict_trace_by_probabilities    src/LocARNA/main_helper.icc:408-426
 *   lb200_run                          replaces ArcMatches(...)                    src/LocARNA/arc_matches.cc:130-188
 *                                               Scoring(...)/Scoring::arcmatch     src/LocARNA/scoring.cc:28-52, :441-554
 *                                               Aligner::align()                   src/LocARNA/aligner.cc:924-962
 *                                               Aligner::trace()                   src/LocARNA/aligner.cc:1345-1363
 *   lb200_pair_score                   replaces the infty_score_t returned by Aligner::align()   src/LocARNA/aligner.hh:120
 *   lb200_pair_alignment               replaces Aligner::get_alignment()           src/LocARNA/aligner.hh:105, alignment.hh:217-249
 *   lb200_params                       carries what AlignerParams / ScoringParams / the locarna CLI pass down
 *                                               src/LocARNA/aligner_params.hh:49-116, scoring.hh:59-166, src/locarna.cc:83-272
 *
 * All functions return LB200_OK (0) or a negative error code; lb200_last_error() gives the message.
 * Nothing throws across this boundary. There is no CPU fallback: without a CUDA device
 * lb200_ctx_create fails.
 *
 * Threading: a context is not thread-safe; use one context per host thread / GPU.
 */
#ifndef LOCARNA_B200_H
#define LOCARNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB200_OK 0
#define LB200_ERR_ARG (-1)
#define LB200_ERR_CUDA (-2)
#define LB200_ERR_UNSUPPORTED (-3)
#define LB200_ERR_IO (-4)
#define LB200_ERR_STATE (-5)

/* score value used for "-inf" (InftyInt::neg_infty, src/LocARNA/infty_int.cc:7-23; printed "-inf") */
#define LB200_SCORE_NEG_INF INT64_MIN

/* lb200_run flags */
#define LB200_RUN_SCORE_ONLY 0 /* D fill + top level score (what mlocarna's guide tree stage consumes, mlocarna:3516-3527) */
#define LB200_RUN_TRACE 1      /* additionally trace back the alignment */
#define LB200_RUN_KEEP_D 2     /* keep the D table for lb200_pair_arcmatches (parity tests) */

typedef struct lb200_ctx lb200_ctx;

typedef struct lb200_params {
    /* heuristics (src/locarna.cc:183-200) */
    double min_prob;              /* --min-prob / -p            0.001 */
    int max_diff_am;              /* --max-diff-am / -D         -1 = off */
    int max_diff_at_am;           /* --max-diff-at-am           -1 = off */
    int max_diff;                 /* --max-diff / -d            -1 = off */
    double min_trace_probability; /* --min-trace-probability    1e-4, 0 = off */
    /* scoring (src/locarna.cc:99-148) */
    int struct_weight;            /* --struct-weight / -s       200 */
    int indel;                    /* --indel / -i               -150 */
    int indel_opening;            /* --indel-opening            -750 */
    int tau;                      /* --tau / -t                 50 */
    int exclusion;                /* --exclusion / -E           0 */
    int match, mismatch;          /* --match 50, --mismatch 0 (without ribosum, or symbols outside ACGU) */
    int use_ribosum;              /* --use-ribosum              1 (built-in RIBOSUM85_60) */
    int unpaired_penalty;         /* --unpaired-penalty         0 */
    int temperature_alipf;        /* --temperature-alipf        300 (envelope partition function) */
    /* locality / constraints (src/locarna.cc:150-181, :251-256) */
    int no_lonely_pairs;          /* --noLP */
    int struct_local;             /* --struct-local */
    int sequ_local;               /* --sequ-local */
    char free_endgaps[8];         /* --free-endgaps "----": left1 right1 left2 right2 */
    int pf_double;                /* 1: envelope in double precision (locarna_p default) instead of long double */
    double exp_prob;              /* --exp-prob / -e            < 0 = not given: background probability 1/(2 len) per sequence (src/locarna.cc:662-663) */
    double max_bps_length_ratio;  /* --max-bps-length-ratio     0 = off; keep only the ratio*length most probable base pairs per sequence (rna_data.cc:64-67) */
    int max_bp_span;              /* --maxBPspan                -1 = unrestricted; base pairs with j-i+1 > span are dropped on input (rna_data.cc:1078) */
    int stacking;                 /* --stacking                 stacked arc matches scored by P(pair | inner pair) (scoring.cc:201-248, aligner.cc:600-607);
                                                                needs the joint probabilities of the PP input (fourth column, #STACK) */
    int new_stacking;             /* --new-stacking             stack weight = weight + weight of the joint probability */
} lb200_params;

typedef struct lb200_pair_info {
    int lenA, lenB;
    int n_arcsA, n_arcsB;
    int64_t n_arcmatches;
    int64_t n_tasks;
    int64_t cells;                /* DP cell updates: sum over D-fill tasks of |box ∩ band| + top level band area */
    int64_t terms;                /* arc-match entries streamed by the D-fill tasks and the top level (each costs an add and a max) */
    int64_t n_edges;              /* alignment edges after lb200_run(LB200_RUN_TRACE) */
} lb200_pair_info;

void lb200_default_params(lb200_params *p);

/* device >= 0: CUDA device ordinal. LB200_DEVICE_NONE creates a host-only context that can read inputs,
 * derive bands and arc matches (lb200_prepare) and be inspected, but whose lb200_run fails: there is no CPU fallback. */
#define LB200_DEVICE_NONE (-1)
int lb200_ctx_create(int device, lb200_ctx **ctx);
void lb200_ctx_destroy(lb200_ctx *ctx);
const char *lb200_last_error(const lb200_ctx *ctx);
/* Device memory of destroyed contexts is kept in a process-wide cache and reused by later contexts (allocation and release of the
 * multi-GB tables otherwise dominate a job); this returns it to the driver. */
void lb200_release_device_cache(void);

/* Parameters must be set before sequences are added (min_prob filters the base pairs). */
int lb200_set_params(lb200_ctx *ctx, const lb200_params *p);

/* Add one RNA from a PP 2.0 file; returns the sequence id (>= 0) or an error code. */
int lb200_seq_add_pp(lb200_ctx *ctx, const char *path);
/* Add n RNAs from PP 2.0 files (parsed on the host's cores in parallel; each file once, where mlocarna's N(N-1)/2 locarna
 * processes parse every file N-1 times); returns the id of the first one, the others follow consecutively. */
int lb200_seqs_add_pp(lb200_ctx *ctx, int n, const char *const *paths);
/* Add one RNA from memory: sequence (ACGU..., T is read as U) and base pairs (1-based i<j, probability). */
int lb200_seq_add(lb200_ctx *ctx, const char *name, const char *seq, const int *pair_i, const int *pair_j, const double *pair_p,
                  int n_pairs);
int lb200_seq_length(const lb200_ctx *ctx, int seq);
/* name (at most name_cap-1 characters) and normalised sequence (length+1 bytes incl. NUL; LB200_ERR_ARG if sequence_cap is
 * smaller) of a sequence; either buffer may be NULL */
int lb200_seq_get(const lb200_ctx *ctx, int seq, char *name, int name_cap, char *sequence, int sequence_cap);
/* Profile input (the step after the guide tree, src/Utils/mlocarna:3660-3716): a PP 2.0 file whose sequence block holds an ALIGNMENT
 * (several rows, gap symbol '-') with its consensus dot plot (MultipleAlignment / Sequence with several rows, sequence.hh:24-80).
 * lb200_seq_add_pp reads it as one "sequence" whose length is the number of alignment columns. Pairs of a context that holds such
 * an input are scored position by position as Scoring does for alignment columns (scoring.cc:141-198 averaged sigma, :272-311
 * gap costs scaled by the gap frequency, :369-438 averaged ribosum arc-match score; stral_score.cc:29-44 for the envelope).
 * Global alignment without free end gaps only; LocARNA-P, anchors, normalized / penalized / k-best are refused for such input.
 * lb200_seq_num_rows: rows of the input (1 for a single sequence). lb200_seq_get_row: name and aligned string of one row (same buffer
 * rules as lb200_seq_get). */
int lb200_seq_num_rows(const lb200_ctx *ctx, int seq);
int lb200_seq_get_row(const lb200_ctx *ctx, int seq, int row, char *name, int name_cap, char *sequence, int sequence_cap);

/* Add one alignment problem (A = seqA, B = seqB). min_col/max_col (lenA+1 entries each) give the band
 * [min_col(i), max_col(i)] per row; pass NULL for both to have it derived like the reference does
 * (--max-diff, then the probability envelope). Returns the pair id (>= 0). */
int lb200_pair_add(lb200_ctx *ctx, int seqA, int seqB, const int *min_col, const int *max_col);
/* A pair whose band is given as TraceController has it BEFORE restrict_by_trace_probabilities (main_helper.icc:371-426): built from a
 * reference alignment (trace_controller.cc:406-539) or restricted by anchors; the probability envelope (min_trace_probability > 0) is
 * then applied inside this range, exactly as for the --max-diff band of lb200_pair_add(.., NULL, NULL). */
int lb200_pair_add_restricted(lb200_ctx *ctx, int seqA, int seqB, const int *min_col, const int *max_col);
/* Rows of TraceController(seqA, seqB, reference alignment, delta) for single sequences (--max-diff-aln / --max-diff-pw-aln with
 * --max-diff delta; TraceRange, trace_controller.cc:44-215, merge_in_trace_range :606-622). aliA / aliB: the rows of the two sequences in
 * the reference alignment (gap symbols "-_~."); relaxed != 0: --max-diff-relax (relaxed merging, trace_controller.cc:485-511); min_col /
 * max_col: lenA + 1 entries. Host only. */
int lb200_band_from_alignment(int lenA, int lenB, const char *aliA, const char *aliB, int delta, int relaxed, int *min_col, int *max_col);
/* Add n alignment problems at once (bands derived like the reference does); returns the id of the first one. This is what the
 * all-vs-all stage of mlocarna hands over (src/Utils/mlocarna:3577-3604: the list of (a, b) index pairs). */
int lb200_pairs_add(lb200_ctx *ctx, int n, const int *seqA, const int *seqB);
int lb200_num_pairs(const lb200_ctx *ctx);
int lb200_clear_pairs(lb200_ctx *ctx);

/* Host-side preparation of all pairs added so far (bands, arc matches, tasks); implied by lb200_run. */
int lb200_prepare(lb200_ctx *ctx);
/* Build the batch of all pairs added so far and copy it to the GPU (HBM resident until pairs change); implied by lb200_run. */
int lb200_upload(lb200_ctx *ctx);
/* Align all pairs added so far on the GPU. */
int lb200_run(lb200_ctx *ctx, int flags);
/* Normalized local alignment (Aligner::normalized_align(L), src/LocARNA/aligner.cc:1522-1597; `locarna --normalized L`, needs sequ_local) and
 * penalized alignment (Aligner::penalized_align(position_penalty), aligner.cc:1599-1622; `locarna --penalized PP`) of all pairs, with
 * traceback. Afterwards lb200_pair_score gives the normalized score (the final lambda) resp. the penalized score, lb200_pair_alignment
 * the alignment. Not available with struct_local. */
int lb200_run_normalized(lb200_ctx *ctx, int64_t L);
int lb200_run_penalized(lb200_ctx *ctx, int64_t position_penalty);
/* Restriction of a pair's top level to rows startA..endA and columns startB..endB (AlignerRestriction, aligner_restriction.hh:26-130;
 * Aligner::set_restriction, aligner.cc:1368-1376); (1, 1, lenA, lenB) lifts it. It applies to lb200_run_pair_toplevel. */
int lb200_pair_set_restriction(lb200_ctx *ctx, int pair, int startA, int startB, int endA, int endB);
/* Aligner::align (+ trace with LB200_RUN_TRACE) of ONE pair on the D table that is already filled - the D fill runs only if no table is
 * resident (AlignerImpl::D_created_, aligner.cc:924-962) - under the pair's restriction. mode: LB200_TOP_PLAIN, LB200_TOP_NORMALIZED
 * (arg = L) or LB200_TOP_PENALIZED (arg = position penalty). This is what k-best alignment by interval splitting (Aligner::suboptimal,
 * aligner.cc:1383-1514; include/locarna_b200.hh) calls per task. */
#define LB200_TOP_PLAIN 0
#define LB200_TOP_NORMALIZED 1
#define LB200_TOP_PENALIZED 2
int lb200_run_pair_toplevel(lb200_ctx *ctx, int pair, int mode, int64_t arg, int flags);
/* device time of the last lb200_run's kernels (CUDA events on the launching stream), milliseconds */
double lb200_last_kernel_ms(const lb200_ctx *ctx);
/* host->device bytes of the last lb200_upload (0 if lb200_run found the batch resident) and device->host bytes of the last lb200_run */
int64_t lb200_last_h2d_bytes(const lb200_ctx *ctx);
int64_t lb200_last_d2h_bytes(const lb200_ctx *ctx);
/* device time and launch count of the D-fill kernel (the dominant kernel) within the last lb200_run */
double lb200_last_dfill_ms(const lb200_ctx *ctx);
int64_t lb200_last_dfill_launches(const lb200_ctx *ctx);
/* which D-fill kernel the last lb200_run used: 0 = one launch per level group, 1 = dependency-driven persistent launch of single
 * boxes, 2 = row-grouped persistent launch (dfill_rows.cu); and how many chunks so far were re-run box by box because the row-grouped
 * kernel met a box it does not support */
int lb200_last_dfill_kind(const lb200_ctx *ctx);
int64_t lb200_rows_fallbacks(const lb200_ctx *ctx);
/* number of kernel launches of the last lb200_run */
int64_t lb200_last_launches(const lb200_ctx *ctx);

/* Band derivation of the last lb200_prepare / lb200_upload: pairs whose probability envelope was decided on the GPU (FP64
 * screening) and pairs that were (re)computed on the host in long double because a cell was too close to the threshold. */
int lb200_envelope_stats(const lb200_ctx *ctx, int64_t *device_pairs, int64_t *host_pairs);

int lb200_pair_score(const lb200_ctx *ctx, int pair, int64_t *score);
int lb200_get_scores(const lb200_ctx *ctx, int64_t *scores, int n);
int lb200_pair_get_info(const lb200_ctx *ctx, int pair, lb200_pair_info *info);
int lb200_pair_band(const lb200_ctx *ctx, int pair, int *min_col, int *max_col);
/* Arc matches in the reference's index order (arc_matches.cc:161-183) with Scoring::arcmatch and, after
 * lb200_run(.. | LB200_RUN_KEEP_D), the D entries (LB200_SCORE_NEG_INF for -inf). Arrays hold n_arcmatches items;
 * any may be NULL. */
int lb200_pair_arcmatches(const lb200_ctx *ctx, int pair, int *al, int *ar, int *bl, int *br, int *score, int64_t *D);
/* Alignment edges in order (position or -1 for a gap), and per-position structure strings
 * (lenA+1 / lenB+1 bytes incl. NUL) as in Alignment (alignment.cc:81-118). edges arrays hold n_edges items. */
int lb200_pair_alignment(const lb200_ctx *ctx, int pair, int *edges_a, int *edges_b, char *str_a, char *str_b);

/* LocARNA-P inside pass (src/locarna_p.cc:441-486; AlignerP<double>::align_inside, src/LocARNA/aligner_p.icc:148-438) for all pairs
 * of the resident batch, in FP64 on the GPU: the partition function Z = M(lenA, lenB) and the inside value D(a,b) of every arc
 * match. Uses the scoring parameters of the context, temperature_alipf as the Boltzmann temperature (scoring.hh:853-856) and
 * pf_scale as in locarna_p --pf-scale. locarna_p derives its band with min_trace_probability 1e-5 and the envelope in double
 * (pf_double = 1); set those in lb200_params for drop-in results. Not available with no_lonely_pairs / struct_local / sequ_local
 * (AlignerP has no such modes). */
int lb200_run_pf(lb200_ctx *ctx, double pf_scale);
int lb200_pair_partition_function(const lb200_ctx *ctx, int pair, double *Z);
/* inside values in the reference's arc-match index order (as lb200_pair_arcmatches); D holds n_arcmatches doubles */
int lb200_pair_arcmatch_pf(const lb200_ctx *ctx, int pair, double *D);
/* LocARNA-P complete (src/locarna_p.cc:483-526): inside pass as lb200_run_pf, then the reverse / outside passes
 * (AlignerP::align_outside, aligner_p.icc:440-1139), the arc-match probabilities (compute_arcmatch_probabilities, :1150-1195) and the
 * base-match probabilities (compute_basematch_probabilities(false), :1201-1399; arc matches with probability > sqrt(min_am_prob)
 * contribute the enclosed cases, as in the reference). */
int lb200_run_pf_probs(lb200_ctx *ctx, double pf_scale, double min_am_prob);
/* probability of every arc match, reference arc-match index order (locarna_p --write-arcmatch-probs lists those >= min_am_prob) */
int lb200_pair_arcmatch_probs(const lb200_ctx *ctx, int pair, double *prob);
/* base-match probabilities, dense (lenA+1) x (lenB+1) row major, entry [i][j] for positions i of A and j of B (1-based; 0 outside the
 * band). locarna_p --write-basematch-probs lists those >= min_bm_prob */
int lb200_pair_basematch_probs(const lb200_ctx *ctx, int pair, double *bm);

/* All-vs-all stage (host): mlocarna's pair order (src/Utils/mlocarna:3577-3604; returns n(n-1)/2, arrays may be NULL), the cost
 * estimate of a pair and the cost-balanced split of a pair list over `world` GPUs / processes - the counterpart of mlocarna's
 * --compute-pairwise-scores k/N (src/Utils/mlocarna:2321-2344). Longest-processing-time-first, deterministic. rank_of[k] = share of
 * pair k; order (optional, n_pairs entries) lists the pair indices share by share, each by descending cost, with rank_begin
 * (optional, world + 1 entries) delimiting the shares. */
int64_t lb200_all_vs_all(int n_seqs, int *seqA, int *seqB);
double lb200_pair_cost(int n_arcsA, int n_arcsB, int lenA, int lenB);
int lb200_shard_pairs(int64_t n_pairs, const double *cost, int world, int *rank_of, int64_t *order, int64_t *rank_begin);
/* Anchor constraints: the PP reader keeps the "#A<k>" annotation rows of a sequence (multiple_alignment.cc:324-346). Pairs whose two
 * sequences both carry anchor names (the same names, strictly increasing: strict semantics of AnchorConstraints, anchor_constraints.cc)
 * get their band restricted as TraceController::restrict_by_anchors does (trace_controller.cc:541-563) before the probability envelope,
 * and only arc matches between positions of equal names (or two unnamed positions) are built (arc_matches.cc:27-28). Global alignment
 * without free end gaps only; names that occur in one sequence only, relaxed anchors and LocARNA-P are refused.
 * lb200_seq_anchors: the annotation as SequenceAnnotation::single_string gives it (rows joined by '#', "" = none); returns its length. */
int lb200_seq_anchors(const lb200_ctx *ctx, int seq, char *out, int cap);
/* --ribosum-file (src/locarna.cc:86-88, main_helper.icc:311-350, RibosumFreq(filename) ribosum.cc:40-200): base-match and arc-match score
 * tables from a matrix file in the reference's extended ribosum format; NULL or "RIBOSUM85_60" selects the built-in matrix. Call it
 * before pairs are added (the tables of resident batches are not rebuilt). */
int lb200_set_ribosum_file(lb200_ctx *ctx, const char *path);
/* Append the parsed sequences of another context (same min_prob / maxBPspan / max-bps-length-ratio / stacking); returns the index of the
 * first one. For contexts that work off one job side by side (one per stream or device): every PP file is parsed once. */
int lb200_seqs_copy(lb200_ctx *ctx, const lb200_ctx *src);
/* The base pairs the PP reader kept for a sequence (RnaData::arc_prob > 0, rna_data.cc:1057-1093), sorted by (i, j): positions,
 * probability, joint probability with the inner pair (0 = none). Pass NULL arrays to get the count (the return value). cutoff_out:
 * RnaData::arc_cutoff_prob(), stacking_out: RnaData::has_stacking(). The input of a consensus dot plot (locarna --pp, rna_data.cc:1474-1548). */
int64_t lb200_seq_pairs(const lb200_ctx *ctx, int seq, int *i, int *j, double *p, double *p2, double *cutoff_out, int *stacking_out);
/* number of base pairs (arcs with probability >= min_prob) of a sequence: the input of lb200_pair_cost */
int lb200_seq_num_arcs(const lb200_ctx *ctx, int seq);
/* lb200_pair_cost + lb200_shard_pairs for pairs (seqA[k], seqB[k]) of the context's sequences in one call; rank_of may be NULL */
int lb200_shard_job(const lb200_ctx *ctx, int64_t n_pairs, const int *seqA, const int *seqB, int world, int *rank_of, int64_t *order,
                    int64_t *rank_begin);

/* Guide tree of the all-vs-all stage (host): UPGMA over the symmetric score matrix (n x n, row major, diagonal 0) with the tie
 * rules of lib/perl/MLocarna/Tree.pm:181-262; writes the newick string (without the trailing ';') that mlocarna stores in
 * results/result.tree (src/Utils/mlocarna:2381-2386). */
int lb200_upgma_newick(int n, const char *const *names, const int64_t *scores, char *out, size_t out_cap);

#ifdef __cplusplus
}
#endif
#endif
