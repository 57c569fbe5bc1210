/* C interface of the CPU restatement (oracle/port/locarna_port.cc). TEST INFRASTRUCTURE ONLY. */
#ifndef LOCARNA_PORT_H
#define LOCARNA_PORT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LOCARNA_PORT_NEG_INF (-4611686018427387904L) /* marker for -inf in D[] */

typedef struct LocarnaPortParams {
    double min_prob;              /* --min-prob            (locarna.cc:183) */
    int max_diff_am;              /* --max-diff-am, -1=off (locarna.cc:190) */
    int max_diff_at_am;           /* --max-diff-at-am                        */
    int max_diff;                 /* --max-diff                              */
    double min_trace_probability; /* --min-trace-probability, 0=off          */
    int noLP, struct_local, sequ_local;
    char free_endgaps[8];         /* "----" / '+' per end: left1 right1 left2 right2 (free_endgaps.hh:27-70) */
    int struct_weight, indel, indel_opening, tau, exclusion, match, mismatch, use_ribosum;
    int temperature_alipf, unpaired_penalty;
    int pf_double;                /* envelope in double instead of long double */
    int do_trace;
    int setup_only;               /* stop after band / arc matches / scores   */
} LocarnaPortParams;

typedef struct LocarnaPortResult {
    int lenA, lenB;
    char *seqA, *seqB;
    int n_arcsA, n_arcsB;
    int *arcsA, *arcsB;           /* (left,right) per arc in index order */
    long *weightsA, *weightsB;
    long *min_col, *max_col;      /* lenA+1 each */
    long n_am;
    int *am;                      /* 5 ints per arc match: al ar bl br inner_idx(-1) */
    long *am_score;               /* Scoring::arcmatch */
    long *D;                      /* per arc match, LOCARNA_PORT_NEG_INF for -inf */
    long score; int score_is_neg_inf;
    int max_i, max_j;
    uint64_t cells, terms, tasks; /* align_noex calls, arc-pair loop iterations, D-fill tasks */
    int n_edges; int *edgesA, *edgesB; /* raw trace edges, -1 = gap */
    char *strA, *strB;            /* per-position brackets of each sequence */
    char *rowA, *rowB, *structA, *structB; /* gapped rows incl. locality gaps */
} LocarnaPortResult;

void locarna_port_default_params(LocarnaPortParams *p);
int locarna_port_align(const char *ppA, const char *ppB, const LocarnaPortParams *p, LocarnaPortResult *res, char *err, int errlen);
void locarna_port_free(LocarnaPortResult *r);
/* LocARNA-P inside (aligner_p.icc:148-438, T = double): Z and one inside value per arc match (malloc'ed, free()) */
int locarna_port_inside_p(const char *ppA, const char *ppB, const LocarnaPortParams *p, double pf_scale, double *Z, double **D, long *n_am,
                          char *err, int errlen);

/* LocARNA-P complete (aligner_p.icc:440-1399, T = double): outside table, arc-match and base-match probabilities */
int locarna_port_probs_p(const char *ppA, const char *ppB, const LocarnaPortParams *p, double pf_scale, double min_am_prob, double *Z,
                         double **D, double **Dprime, double **am_prob, long *n_am, double **bm_prob, int *lenA, int *lenB, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
