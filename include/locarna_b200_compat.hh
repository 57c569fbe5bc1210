// locarna_b200_compat.hh -- source compatibility with the reference's pipeline code.
//
// include/locarna_b200.hh mirrors the reference's classes with plain C++ (chained setters, one GPU context per object). This header
// adds the reference's own SPELLING on top of it, so that code written against LocARNA 2.0.1 compiles unchanged after
//
//     #include <locarna_b200_compat.hh>
//     namespace LocARNA = LocARNA_B200::compat;
//
// The proof is in the build: the body of run_and_report() of the reference's src/locarna.cc (from "bool skip_aligning" down to
// "arc_matches.reset()", i.e. ribosum set-up, RnaData, TraceController, ArcMatches, ScoringParams / Scoring, AlignerParams / Aligner,
// align / trace / get_alignment) is extracted from /root/reference at build time and compiled, unmodified, inside
// locarna_b200/csrc/cli/refmain_driver.cc (Makefile target `refmain`; tests/test_boundary.py diffs the program's output with the
// reference binary's). Mirrored here, with the reference declaration each follows:
//
//   named arguments               src/LocARNA/named_arguments.hh   (Class::name(value) objects, any order, defaults)
//   PFoldParams                   src/LocARNA/pfold_params.hh:31-135
//   Sequence / SeqEntry           src/LocARNA/sequence.hh, multiple_alignment.hh:95-170
//   RnaData                       src/LocARNA/rna_data.hh:60-140
//   AnchorConstraints             src/LocARNA/anchor_constraints.hh:40-110
//   TraceController               src/LocARNA/trace_controller.hh:200-330
//   ArcMatches                    src/LocARNA/arc_matches.hh:300-330
//   ScoringParams / Scoring       src/LocARNA/scoring.hh:59-166, :271-330
//   FreeEndgaps                   src/LocARNA/free_endgaps.hh:17-80
//   AlignerParams / Aligner       src/LocARNA/aligner_params.hh:49-116, aligner.hh:67-189
//   MainHelper::*                 src/LocARNA/main_helper.icc (init_ribo_matrix, restrict_trace_by_probabilities, report_input, ...)
//
// The objects of the reference compute on construction (RnaData parses, ArcMatches enumerates, Scoring precomputes). Here they record
// their arguments; the device builds bands, arc matches and scores when Aligner runs. What the B200 path does not implement
// (relaxed anchors, MEA, explicit arc-match scores) throws LocARNA::failure from the object
// that would need it, so the caller's existing error handling applies.
#ifndef LOCARNA_B200_COMPAT_HH
#define LOCARNA_B200_COMPAT_HH

#include <cmath>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "locarna_b200.hh"

namespace LocARNA_B200 {
namespace compat {

typedef size_t size_type;
typedef long score_t;
using LocARNA_B200::Alignment;
using LocARNA_B200::failure;
using LocARNA_B200::infty_score_t;

#define LB200_NAMED_ARG(Name, Type, Member)             \
    struct Name {                                        \
        Type v;                                          \
        explicit Name(Type v_) : v(v_) {}                \
    };                                                   \
    void set(const Name &a) { Member = a.v; }

inline double prob_exp_f(int seqlen) { return 1.0 / (2.0 * seqlen); }   // aux.hh:216-220

inline void split_at_separator(const std::string &s, char c, std::vector<std::string> &v) {   // aux.cc:84-95
    std::string seg;
    std::istringstream in(s);
    v.clear();
    while (std::getline(in, seg, c)) v.push_back(seg);
}

class Ribosum { public: virtual ~Ribosum() {} };
class RibosumFreq : public Ribosum {   // ribosum.hh: the built-in RIBOSUM85_60 tables live in the library; a file is read by it, too
    std::string file_;
public:
    RibosumFreq() {}
    explicit RibosumFreq(const std::string &file) : file_(file) {   // parse now, so that a bad file is reported where the reference reports it
        lb200_ctx *c = nullptr;
        if (lb200_ctx_create(LB200_DEVICE_NONE, &c) != LB200_OK) throw failure("locarna_b200: cannot create a host context");
        const int rc = lb200_set_ribosum_file(c, file.c_str());
        const std::string msg = rc < 0 ? lb200_last_error(c) : "";
        lb200_ctx_destroy(c);
        if (rc < 0) throw failure(msg);
    }
    const std::string &file() const { return file_; }
};
class Ribofit {};
class MatchProbs {};

class PFoldParams {   // pfold_params.hh
    bool noLP_ = false, stacking_ = false;
    int max_bp_span_ = -1;
public:
    struct args {
        struct noLP { bool v; explicit noLP(bool v_) : v(v_) {} };
        struct stacking { bool v; explicit stacking(bool v_) : v(v_) {} };
        struct max_bp_span { int v; explicit max_bp_span(int v_) : v(v_) {} };
    };
    template <class... Args> explicit PFoldParams(Args... a) { int unused[] = {0, (set(a), 0)...}; (void)unused; }
    void set(const args::noLP &a) { noLP_ = a.v; }
    void set(const args::stacking &a) { stacking_ = a.v; }
    void set(const args::max_bp_span &a) { max_bp_span_ = a.v; }
    bool noLP() const { return noLP_; }
    bool stacking() const { return stacking_; }
    int max_bp_span() const { return max_bp_span_; }
};

class SequenceAnnotation {   // sequence_annotation.hh: here only the anchor rows of a PP file, joined by '#'
    std::string str_;
public:
    SequenceAnnotation() {}
    explicit SequenceAnnotation(const std::string &s) : str_(s) {}
    bool empty() const { return str_.empty(); }
    std::string single_string() const { return str_; }
};

class MultipleAlignment : public LocARNA_B200::MultipleAlignment {   // multiple_alignment.hh
public:
    enum class AnnoType { consensus_structure, structure, fixed_structure, anchors };
    MultipleAlignment(const Alignment &a, bool only_local = false, bool = false) : LocARNA_B200::MultipleAlignment(a, only_local) {}
    //! reference alignments for --max-diff-aln (CLUSTAL file) / --max-diff-pw-aln (two rows)
    explicit MultipleAlignment(const std::string &file) : LocARNA_B200::MultipleAlignment(file) {}
    MultipleAlignment(const std::string &nameA, const std::string &nameB, const std::string &aliA, const std::string &aliB)
        : LocARNA_B200::MultipleAlignment(nameA, nameB, aliA, aliB) {}
};

class Sequence {   // one row: the B200 path aligns single sequences (profile inputs: SURVEY 8f N3)
public:
    class SeqEntry {
        std::string name_, seq_;
    public:
        SeqEntry(const std::string &n, const std::string &s) : name_(n), seq_(s) {}
        const std::string &name() const { return name_; }
        const std::string &seq() const { return seq_; }
    };
    Sequence(const std::string &name, const std::string &seq, const std::string &anchors = "") : entry_(name, seq), anno_(anchors) {}
    size_type length() const { return entry_.seq().size(); }
    size_type num_of_rows() const { return 1; }
    const SeqEntry &seqentry(size_type) const { return entry_; }
    const SequenceAnnotation &annotation(MultipleAlignment::AnnoType) const { return anno_; }
private:
    SeqEntry entry_;
    SequenceAnnotation anno_;
};

class RnaData {   // rna_data.hh:102-106: RnaData(filename, p_bpcut, max_bps_length_ratio, pfoldparams); PP 2.0 input
    LocARNA_B200::RnaData data_;
    std::unique_ptr<Sequence> seq_;
public:
    RnaData(const std::string &file, double p_bpcut, double max_bps_length_ratio, const PFoldParams &pf)
        : data_(file, p_bpcut, max_bps_length_ratio, pf.max_bp_span()) {
        // parse once on the host to get name and sequence (and to report unreadable input where the reference does)
        lb200_ctx *c = nullptr;
        if (lb200_ctx_create(LB200_DEVICE_NONE, &c) != LB200_OK) throw failure("locarna_b200: cannot create a host context");
        lb200_params p;
        lb200_default_params(&p);
        p.min_prob = p_bpcut; p.max_bp_span = pf.max_bp_span(); p.max_bps_length_ratio = max_bps_length_ratio;
        int id = lb200_set_params(c, &p);
        if (id >= 0) id = lb200_seq_add_pp(c, file.c_str());
        if (id < 0) { const std::string msg = lb200_last_error(c); lb200_ctx_destroy(c); throw failure(msg); }
        const int len = lb200_seq_length(c, id);
        std::vector<char> name(512), seq((size_t)len + 1);
        lb200_seq_get(c, id, name.data(), (int)name.size(), seq.data(), (int)seq.size());
        std::vector<char> anchors((size_t)lb200_seq_anchors(c, id, nullptr, 0) + 1);
        lb200_seq_anchors(c, id, anchors.data(), (int)anchors.size());
        seq_.reset(new Sequence(name.data(), seq.data(), anchors.data()));
        lb200_ctx_destroy(c);
    }
    const Sequence &sequence() const { return *seq_; }
    size_type length() const { return seq_->length(); }
    double arc_cutoff_prob() const { return data_.arc_cutoff_prob(); }
    const LocARNA_B200::RnaData &data() const { return data_; }
};

class AnchorConstraints {   // anchor_constraints.hh: the library applies the "#A" annotation of the PP inputs itself (strict semantics)
    bool empty_;
public:
    AnchorConstraints(size_type, const std::string &anchorsA, size_type, const std::string &anchorsB, bool strict)
        : empty_(anchorsA.empty() || anchorsB.empty()) {   // one spec empty: no anchors at all (anchor_constraints.cc:27-31)
        if (!empty_ && !strict) throw failure("locarna_b200: relaxed anchor constraints are not supported");
    }
    bool empty() const { return empty_; }
};

class TraceController {   // trace_controller.hh:200: the band; its rows are derived on the device when the aligner runs
    int max_diff_;
    double min_trace_probability_ = 0.0;
    const MultipleAlignment *ref_aln_ = nullptr;
public:
    bool relax_ = false;
    TraceController(const Sequence &, const Sequence &, const MultipleAlignment *ma, int max_diff, bool relax = false) : max_diff_(max_diff), ref_aln_(ma), relax_(relax) {}
    const MultipleAlignment *reference_alignment() const { return ref_aln_; }
    bool relaxed_merging() const { return relax_; }
    void restrict_by_anchors(const AnchorConstraints &) {}   // done by the library when the pair is added (lb200_seq_anchors)
    //! what MainHelper::restrict_trace_by_probabilities records (main_helper.icc:408-426)
    void set_min_trace_probability(double p) { min_trace_probability_ = p; }
    int max_diff() const { return max_diff_; }
    double min_trace_probability() const { return min_trace_probability_; }
};

class Scoring;

class ArcMatches {   // arc_matches.hh:307-313
    const RnaData *rnaA_, *rnaB_;
    double min_prob_;
    size_type max_diff_am_, max_diff_at_am_;
    const TraceController *tc_;
public:
    ArcMatches(const RnaData &a, const RnaData &b, double min_prob, size_type max_length_diff, size_type max_diff_at_am, const TraceController &tc,
               const AnchorConstraints &constraints)
        : rnaA_(&a), rnaB_(&b), min_prob_(min_prob), max_diff_am_(max_length_diff), max_diff_at_am_(max_diff_at_am), tc_(&tc) {
        (void)constraints;   // applied by the library (arc matches join positions of equal anchor names only)
    }
    //! explicit arc-match scores (--read-arcmatch-scores / --read-arcmatch-probs, arc_matches.cc:190-282)
    ArcMatches(const Sequence &, const Sequence &, const std::string &file, int, size_type, size_type, const TraceController &, const AnchorConstraints &) {
        throw failure("locarna_b200: explicit arc-match scores (" + file + ") are not supported");
    }
    const RnaData &rnaA() const { return *rnaA_; }
    const RnaData &rnaB() const { return *rnaB_; }
    double min_prob() const { return min_prob_; }
    size_type max_length_diff() const { return max_diff_am_; }
    size_type max_diff_at_am() const { return max_diff_at_am_; }
    const TraceController &trace_controller() const { return *tc_; }
    inline void write_arcmatch_scores(const std::string &file, const Scoring &scoring) const;   // arc_matches.cc:285-311
};

class ScoringParams {   // scoring.hh:59-166 (same names, same defaults)
public:
    int match_ = 50, mismatch_ = 0, indel_ = -150, indel_opening_ = -750, unpaired_penalty_ = 0, struct_weight_ = 200, tau_factor_ = 50, exclusion_ = 0;
    const Ribosum *ribosum_ = nullptr;
    const Ribofit *ribofit_ = nullptr;
    double exp_probA_ = -1, exp_probB_ = -1;
    int temperature_alipf_ = 300;
    bool stacking_ = false, new_stacking_ = false, mea_scoring_ = false;
    int mea_alpha_ = 0, mea_beta_ = 200, mea_gamma_ = 100, probability_scale_ = 10000;
    LB200_NAMED_ARG(match, int, match_)
    LB200_NAMED_ARG(mismatch, int, mismatch_)
    LB200_NAMED_ARG(indel, int, indel_)
    LB200_NAMED_ARG(indel_opening, int, indel_opening_)
    LB200_NAMED_ARG(ribosum, const Ribosum *, ribosum_)
    LB200_NAMED_ARG(ribofit, const Ribofit *, ribofit_)
    LB200_NAMED_ARG(unpaired_penalty, int, unpaired_penalty_)
    LB200_NAMED_ARG(struct_weight, int, struct_weight_)
    LB200_NAMED_ARG(tau_factor, int, tau_factor_)
    LB200_NAMED_ARG(exclusion, int, exclusion_)
    LB200_NAMED_ARG(exp_probA, double, exp_probA_)
    LB200_NAMED_ARG(exp_probB, double, exp_probB_)
    LB200_NAMED_ARG(temperature_alipf, int, temperature_alipf_)
    LB200_NAMED_ARG(stacking, bool, stacking_)
    LB200_NAMED_ARG(new_stacking, bool, new_stacking_)
    LB200_NAMED_ARG(mea_scoring, bool, mea_scoring_)
    LB200_NAMED_ARG(mea_alpha, int, mea_alpha_)
    LB200_NAMED_ARG(mea_beta, int, mea_beta_)
    LB200_NAMED_ARG(mea_gamma, int, mea_gamma_)
    LB200_NAMED_ARG(probability_scale, int, probability_scale_)
    template <class... Args> explicit ScoringParams(Args... a) { int unused[] = {0, (set(a), 0)...}; (void)unused; }
};

class Scoring {   // scoring.hh:302-330: borrowed references, as in the reference
    const RnaData *rnaA_, *rnaB_;
    const ArcMatches *arc_matches_;
    ScoringParams params_;
public:
    Scoring(const Sequence &, const Sequence &, const RnaData &rnaA, const RnaData &rnaB, const ArcMatches &arc_matches, const MatchProbs *match_probs,
            const ScoringParams &params)
        : rnaA_(&rnaA), rnaB_(&rnaB), arc_matches_(&arc_matches), params_(params) {
        if (match_probs != nullptr || params.mea_scoring_) throw failure("locarna_b200: MEA scoring is not supported");
        if (params.ribofit_ != nullptr) throw failure("locarna_b200: ribofit is not supported");
    }
    const ScoringParams &params() const { return params_; }
    const ArcMatches &arc_matches() const { return *arc_matches_; }
    const RnaData &rnaA() const { return *rnaA_; }
    const RnaData &rnaB() const { return *rnaB_; }
    score_t indel() const { return params_.indel_; }
    score_t indel_opening() const { return params_.indel_opening_; }
    score_t exclusion() const { return params_.exclusion_; }
    bool stacking() const { return params_.stacking_ || params_.new_stacking_; }
};

class FreeEndgaps {   // free_endgaps.hh:17-80
    std::string d_;
public:
    FreeEndgaps() : d_("----") {}
    explicit FreeEndgaps(const std::string &d) : d_(d.size() >= 4 ? d : "----") {}
    bool allow_left_1() const { return d_[0] == '+'; }
    bool allow_right_1() const { return d_[1] == '+'; }
    bool allow_left_2() const { return d_[2] == '+'; }
    bool allow_right_2() const { return d_[3] == '+'; }
    const std::string &str() const { return d_; }
};

class AlignerParams {   // aligner_params.hh:51-115
public:
    const Sequence *seqA_ = nullptr, *seqB_ = nullptr;
    const Scoring *scoring_ = nullptr;
    bool no_lonely_pairs_ = false, struct_local_ = false, sequ_local_ = false, stacking_ = false;
    FreeEndgaps free_endgaps_;
    int max_diff_am_ = -1, max_diff_at_am_ = -1;
    const TraceController *trace_controller_ = nullptr;
    const AnchorConstraints *constraints_ = nullptr;
    LB200_NAMED_ARG(seqA, const Sequence *, seqA_)
    LB200_NAMED_ARG(seqB, const Sequence *, seqB_)
    LB200_NAMED_ARG(scoring, const Scoring *, scoring_)
    LB200_NAMED_ARG(no_lonely_pairs, bool, no_lonely_pairs_)
    LB200_NAMED_ARG(struct_local, bool, struct_local_)
    LB200_NAMED_ARG(sequ_local, bool, sequ_local_)
    LB200_NAMED_ARG(free_endgaps, FreeEndgaps, free_endgaps_)
    LB200_NAMED_ARG(max_diff_am, int, max_diff_am_)
    LB200_NAMED_ARG(max_diff_at_am, int, max_diff_at_am_)
    LB200_NAMED_ARG(trace_controller, const TraceController *, trace_controller_)
    LB200_NAMED_ARG(stacking, bool, stacking_)
    LB200_NAMED_ARG(constraints, const AnchorConstraints *, constraints_)
    template <class... Args> explicit AlignerParams(Args... a) { int unused[] = {0, (set(a), 0)...}; (void)unused; }
};

//! what the recorded objects amount to for the library (one conversion used by Aligner and by write_arcmatch_scores)
inline LocARNA_B200::AlignerParams to_b200_params(const Scoring &s, const TraceController &tc, bool noLP, bool struct_local, bool sequ_local,
                                                  const std::string &free_endgaps, int max_diff_am, int max_diff_at_am) {
    const ScoringParams &p = s.params();
    LocARNA_B200::ScoringParams sp;
    sp.match = p.match_; sp.mismatch = p.mismatch_; sp.indel = p.indel_; sp.indel_opening = p.indel_opening_; sp.unpaired_penalty = p.unpaired_penalty_;
    sp.struct_weight = p.struct_weight_; sp.tau_factor = p.tau_factor_; sp.exclusion = p.exclusion_; sp.temperature_alipf = p.temperature_alipf_;
    sp.use_ribosum = p.ribosum_ != nullptr;
    if (const RibosumFreq *rf = dynamic_cast<const RibosumFreq *>(p.ribosum_)) sp.ribosum_file = rf->file();
    sp.stacking = p.stacking_; sp.new_stacking = p.new_stacking_;
    // background probabilities: the library takes one value for both sequences, or derives 1 / (2 len) per sequence (locarna.cc:662-663)
    const double defA = prob_exp_f((int)s.rnaA().length()), defB = prob_exp_f((int)s.rnaB().length());
    if (p.exp_probA_ == defA && p.exp_probB_ == defB) sp.exp_prob = -1.0;
    else if (p.exp_probA_ == p.exp_probB_) sp.exp_prob = p.exp_probA_;
    else throw failure("locarna_b200: different background probabilities for the two sequences are not supported");
    LocARNA_B200::AlignerParams ap;
    ap.seqA(&s.rnaA().data()).seqB(&s.rnaB().data()).scoring(sp).min_prob(s.arc_matches().min_prob());
    ap.no_lonely_pairs(noLP).struct_local(struct_local).sequ_local(sequ_local).free_endgaps(free_endgaps);
    ap.max_diff_am(max_diff_am).max_diff_at_am(max_diff_at_am).max_diff(tc.max_diff()).min_trace_probability(tc.min_trace_probability());
    ap.reference_alignment(tc.reference_alignment(), tc.relaxed_merging());
    return ap;
}

class Aligner {   // aligner.hh:67-189
    std::unique_ptr<LocARNA_B200::Aligner> impl_;
public:
    explicit Aligner(const AlignerParams &ap) {
        if (!ap.seqA_ || !ap.seqB_ || !ap.scoring_ || !ap.trace_controller_) throw failure("AlignerParams: seqA, seqB, scoring and trace_controller are mandatory");
        impl_.reset(new LocARNA_B200::Aligner(to_b200_params(*ap.scoring_, *ap.trace_controller_, ap.no_lonely_pairs_, ap.struct_local_, ap.sequ_local_,
                                                               ap.free_endgaps_.str(), ap.max_diff_am_, ap.max_diff_at_am_)));
    }
    infty_score_t align() { return impl_->align(); }
    void trace() { impl_->trace(); }
    const Alignment &get_alignment() const { return impl_->get_alignment(); }
    infty_score_t normalized_align(score_t L, bool verbose) { return impl_->normalized_align(L, verbose); }
    infty_score_t penalized_align(score_t position_penalty) { return impl_->penalized_align(position_penalty); }
    void set_restriction(const LocARNA_B200::AlignerRestriction &r) { impl_->set_restriction(r); }
    const LocARNA_B200::AlignerRestriction &get_restriction() const { return impl_->get_restriction(); }
    void suboptimal(int k, score_t threshold, bool normalized, score_t normalized_L, size_t output_width, bool verbose, bool opt_local_output,
                    bool opt_pos_output, bool opt_write_structure) {
        impl_->suboptimal(k, threshold, normalized, normalized_L, output_width, verbose, opt_local_output, opt_pos_output, opt_write_structure);
    }
};

inline void ArcMatches::write_arcmatch_scores(const std::string &file, const Scoring &scoring) const {
    // the limits the caller passed are lengths ("no limit" = max(lenA, lenB), locarna.cc:617-626); the library takes -1 for "off"
    const size_type no_limit = std::max(rnaA_->length(), rnaB_->length());
    LocARNA_B200::Aligner a(to_b200_params(scoring, *tc_, false, false, false, "----", max_diff_am_ >= no_limit ? -1 : (int)max_diff_am_,
                                           max_diff_at_am_ >= no_limit ? -1 : (int)max_diff_at_am_));
    a.arc_matches().write_arcmatch_scores(file);
}

namespace MainHelper {   // main_helper.icc
template <class CLP>
void init_ribo_matrix(const CLP &clp, std::unique_ptr<RibosumFreq> &ribosum, std::unique_ptr<Ribofit> &ribofit) {   // :311-350
    ribofit.reset();
    ribosum.reset();
    if (clp.ribofit) throw failure("locarna_b200: ribofit is not supported");
    if (clp.use_ribosum) {
        if (clp.ribosum_file == "RIBOSUM85_60") ribosum.reset(new RibosumFreq());   // main_helper.icc:326-333
        else ribosum.reset(new RibosumFreq(clp.ribosum_file));
    }
}
template <class CLP, class PF>
void write_trace_probs(const CLP &, const RnaData *, const RnaData *, const Ribosum *, const Ribofit *, TraceController *, const PF &) {
    throw failure("locarna_b200: --write-trace-probs is not supported");
}
template <class CLP, class PF>
std::unique_ptr<MatchProbs> init_match_probs(CLP &, const RnaData *, const RnaData *, const TraceController *, const Ribosum *, const Ribofit *, const PF &) {
    throw failure("locarna_b200: match probabilities (MEA alignment) are not supported");
}
template <class CLP>
void write_match_probs(const CLP &, const MatchProbs *) { throw failure("locarna_b200: --write-match-probs is not supported"); }
inline void report_input(const Sequence &seqA, const Sequence &seqB, const ArcMatches &) {   // :465-487 (arc-match counts are known once the device has built them)
    std::cout << "Sequence A: " << seqA.seqentry(0).name() << " (Length:" << seqA.length() << ")" << std::endl;
    std::cout << "Sequence B: " << seqB.seqentry(0).name() << " (Length:" << seqB.length() << ")" << std::endl << std::endl;
}
}  // namespace MainHelper

//! main_helper.icc:408-426: the envelope itself is computed on the device when the aligner runs; the controller records the threshold
template <class CLP, class PF>
void restrict_trace_by_probabilities(CLP &clp, const RnaData *, const RnaData *, const Ribosum *, const Ribofit *, TraceController *tc, const PF &) {
    if (clp.min_trace_probability > 0.0) tc->set_min_trace_probability(clp.min_trace_probability);
}

}  // namespace compat
}  // namespace LocARNA_B200
#endif
