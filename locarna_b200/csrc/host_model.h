// Host-side problem model of the B200 aligner: what the reference keeps in RnaData / BasePairs /
// TraceController / Scoring, reduced to flat arrays that can be uploaded once per sequence / pair.
// (Reference files cited per function in host_model.cc.)
#ifndef LB200_HOST_MODEL_H
#define LB200_HOST_MODEL_H
#include <stdint.h>
#include <string>
#include <vector>
#include "dev_types.h"

namespace lb200 {

// Score tables derived from a RIBOSUM matrix (ribosum.hh:34-439): base similarity as stored (used by the envelope's STRAL-like score,
// stral_score.cc:29-60), round2score(100 log2(P_basematch / (P_nonstruct P_nonstruct))) (scoring.cc:141-198, ribosum.cc:324-331) and
// round2score(100 log2(P_arcmatch / (P_basepair P_basepair))) (scoring.cc:369-438). Nucleotide order A C G U, pair index 4 left + right.
struct RibosumTables {
    double bm[16];
    int sigma4[16];
    int am16[256];
    double amlog2[256];   // log2(P_arcmatch / (P_basepair P_basepair)) unrounded: profile inputs average it over row pairs (scoring.cc:369-438)
    RibosumTables();   // the built-in RIBOSUM85_60
};
// --ribosum-file: a matrix in the reference's extended ribosum format (RibosumFreq(filename), ribosum.cc:40-200)
bool read_ribosum_file(const std::string &path, RibosumTables &out, std::string &err);

struct Params {  // mirrors the `locarna` CLI options that reach the path (locarna.cc:83-272)
    double min_prob = 0.001;
    int max_diff_am = -1, max_diff_at_am = -1, max_diff = -1;
    double min_trace_probability = 1e-4;
    bool no_lonely_pairs = false, struct_local = false, sequ_local = false;
    bool fe_left1 = false, fe_right1 = false, fe_left2 = false, fe_right2 = false;
    int struct_weight = 200, indel = -150, indel_opening = -750, tau = 50, exclusion = 0;
    int match = 50, mismatch = 0, unpaired_penalty = 0, temperature_alipf = 300;
    bool use_ribosum = true;
    double exp_prob = -1.0;  // --exp-prob (locarna.cc:115, :662-663); < 0: not given, background probability 1/(2 len)
    double max_bps_length_ratio = 0.0;  // --max-bps-length-ratio (locarna.cc:185, rna_data.cc:64-67); 0: keep all base pairs
    int max_bp_span = -1;    // --maxBPspan (locarna.cc:253, rna_data.cc:1078); -1: unrestricted
    bool pf_double = false;  // envelope in double (locarna_p default) instead of long double (locarna)
    bool stacking = false, new_stacking = false;  // --stacking / --new-stacking (locarna.cc:120-123, scoring.cc:201-248)
    RibosumTables ribosum;   // --ribosum-file (default: the built-in RIBOSUM85_60)
};

struct Arc { int left, right; };

// One RNA: sequence + sparse base pair probabilities, already filtered like RnaData/BasePairs do.
struct Sequence {
    std::string name, seq;             // seq[0] is position 1
    int len = 0;
    std::vector<uint8_t> codes;        // 1-based symbol codes (A C G U = 0..3, N = 4, other symbols 5..7)
    // all pairs kept by the PP reader (p > cutoff), for the envelope's paired-up/down sums
    std::vector<int> pp_i, pp_j;
    std::vector<double> pp_p;
    std::vector<double> pp_p2;         // joint probability of (i, j) and (i+1, j-1) (PP fourth column, rna_data.cc:1085-1093); 0: none
    double cutoff = 0;
    bool has_stacking = false;         // RnaData::has_stacking(): the file carries joint probabilities and they were read
    // arcs with p >= min_prob in the reference's index order (left descending, right ascending)
    std::vector<Arc> arcs;
    std::vector<double> arc_prob;
    std::vector<double> arc_joint;     // RnaData::joint_arc_prob of the arc
    std::vector<double> arc_inner;     // RnaData::arc_prob(left + 1, right - 1)
    std::vector<int> lptr;             // arcs with left end l are arcs[lptr[l] .. lptr[l]+lcount[l])
    std::vector<int> lcount;
    std::vector<double> p_up, p_down;  // rna_data.cc:713-732
    // anchor constraints (PP annotation lines "#A<k> <string>", multiple_alignment.cc:324-346): name of every position ("" = none) and
    // its rank among the names of this sequence (1-based, 0 = unnamed); names must increase strictly (anchor_constraints.cc:100-132)
    std::vector<std::string> anchor_rows;    // the annotation rows as read (each of length len); empty: no annotation
    std::vector<std::string> anchor_names;   // entries 0..len (0 unused); empty vector: no annotation
    std::vector<uint8_t> anchor_rank;        // entries 0..len
    // profile input (an alignment with its consensus dot plot, as mlocarna's progressive stage passes it on): the rows of the
    // alignment, each `len` columns long, gap symbol '-'. Single sequences have one row (rows may be empty: then it is `seq`).
    std::vector<std::string> row_names, rows;
    int num_rows() const { return rows.empty() ? 1 : (int)rows.size(); }
    const std::string &row(int k) const { return rows.empty() ? seq : rows[(size_t)k]; }
};

// Position-specific score tables of a pair with profile input (Scoring::precompute_sigma / precompute_gapcost / riboX_arcmatch_score,
// scoring.cc:141-198, :272-311, :369-438; unpaired penalty applied as scoring.cc:64-74).
// The kernels run such a pair in the GAP-FREE FRAME: every alignment of the cells (al, bl)..(i, j) deletes or matches each row
// al+1..i and inserts or matches each column bl+1..j exactly once, so subtracting gapA(i) and gapB(j) from every operation that
// consumes row i / column j shifts all alignments of the same box by the same amount:
//     sigma'(i, j) = sigma(i, j) - gapA(i) - gapB(j),   gap' = 0,   opening unchanged,
//     arcmatch'(a, b) = arcmatch(a, b) - gapA(al) - gapA(ar) - gapB(bl) - gapB(br),
//     M'(i, j) = M(i, j) - (PA[i] - PA[al]) - (PB[j] - PB[bl]),   D'(a, b) = D(a, b) - (PA[ar] - PA[al-1]) - (PB[br] - PB[bl-1])
// with the prefix sums PA, PB of the gap costs. Maxima, ties and therefore scores and tracebacks are those of the original recurrences
// (global alignment; a clamp at 0 or free end gaps would not commute with the shift and are refused for profile input).
struct ProfileTables {
    int n = 0, m = 0, arcsA = 0, arcsB = 0;
    std::vector<int> sigma;          // (n+1) x (m+1), Scoring::basematch(i, j)
    std::vector<int> gapA, gapB;     // Scoring::gapA(i), gapB(j), entries 1..len
    std::vector<long> PA, PB;        // prefix sums of the gap costs, entries 0..len
    std::vector<int> am_seq;         // arcsA x arcsB: (tau * sequence contribution) / 100 of Scoring::arcmatch
    int sigma_shifted(int i, int j) const { return sigma[(size_t)i * (m + 1) + j] - (i >= 1 ? gapA[i] : 0) - (j >= 1 ? gapB[j] : 0); }
};
void make_profile_tables(const Sequence &A, const Sequence &B, const Params &p, ProfileTables &out);

struct Band {
    int lenA = 0, lenB = 0;
    std::vector<int> lo, hi;  // min_col / max_col, entries 0..lenA
};

struct ScoreTables {
    int am_seq[256];           // (tau * ribosum arc match score) / 100 per (pair type A, pair type B), ACGU only
    DevParams dev;
};

// PP 2.0 reader
bool read_pp(const std::string &path, double p_bpcut, Sequence &out, std::string &err, int max_bp_span = -1, double max_bps_length_ratio = 0.0,
             bool stacking = false);
// sequence + explicit pair list (i, j, p): same filtering as the PP reader with #BPCUT = cutoff
bool make_sequence(const std::string &name, const std::string &seq, const int *pi, const int *pj, const double *pp, int npairs,
                   double p_bpcut, Sequence &out, std::string &err, int max_bp_span = -1, double max_bps_length_ratio = 0.0, const double *pp2 = nullptr);
void finish_sequence(Sequence &s, double min_prob);
bool set_anchors(Sequence &s, const std::vector<std::string> &anchor_rows, std::string &err);

std::vector<int> arc_weights(const Sequence &s, const Params &p);
// per arc: stack weight minus weight (scoring.cc:201-248), LB_NOSTACK for arcs that are not stackable (scoring.cc:556-564)
std::vector<int> arc_stack_deltas(const Sequence &s, const Params &p);
void make_score_tables(const Params &p, ScoreTables &t);
int base_match_score(const ScoreTables &t, uint8_t a, uint8_t b);
int arcmatch_score(const ScoreTables &t, const Params &p, const Sequence &A, const Sequence &B, int a, int b, const std::vector<int> &wA,
                   const std::vector<int> &wB);

Band make_band(int lenA, int lenB, int max_diff);
// Anchor constraints of a pair whose two sequences carry the same anchor names (strict semantics, anchor_constraints.cc:164-330): rows
// [min_col, max_col] of `band` restricted like TraceController::restrict_by_anchors (trace_controller.cc:541-563). Returns 0 if the pair
// has no constraints (one sequence without names), 1 if the band was restricted, -1 (err set) if the names differ between the sequences.
int restrict_band_by_anchors(Band &band, const Sequence &A, const Sequence &B, std::string &err);
// band around a pairwise reference alignment (trace_controller.cc:44-215, :606-622)
bool band_from_alignment(int lenA, int lenB, const std::string &aliA, const std::string &aliB, int delta, Band &b, std::string &err, bool relaxed = false);
// probability envelope (PFGotoh in 80-bit or 64-bit floating point on the host)
void restrict_band_by_envelope(Band &band, const Sequence &A, const Sequence &B, const Params &p);
// score parameters of the envelope partition function as the reference passes them (main_helper.icc:389-400)
void envelope_score_params(const Params &p, double bm[16], double *sw, double *open, double *ext, double *temp);

// Per-pair problem in device layout, built on the host (L-order / S-order, see dev_types.h)
struct PairProblem {
    std::vector<DevArcMatch> am;      // L-order
    std::vector<int> am_a, am_b;      // arc indices per L-order arc match
    std::vector<DevEntry> ent;        // S-order
    std::vector<int> sptr;            // lenA+lenB+2
    std::vector<DevTask> tasks;       // pair field left 0
    int wd_bound = 1;                 // upper bound on the number of band diagonals of any box
    int max_box_words = 0;
    uint64_t cells = 0;               // DP cell updates of all D-fill tasks + top level (reference count)
    uint64_t terms = 0;               // arc-match entries streamed by all boxes (S-order range of the box anti-diagonals)
};
void build_pair_problem(const Sequence &A, const Sequence &B, const Band &band, const Params &p, const ScoreTables &t, PairProblem &out, bool anchored = false,
                        const ProfileTables *profile = nullptr);

}  // namespace lb200
#endif
