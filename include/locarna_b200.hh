// locarna_b200.hh -- C++ mirror of the reference's operator API for the pairwise alignment path, header-only over
// the C ABI (locarna_b200.h). Same class names, method names, argument meaning and error behaviour as LocARNA 2.0.1:
//
//   RnaData            src/LocARNA/rna_data.hh:60           (PP 2.0 input only)
//   ScoringParams      src/LocARNA/scoring.hh:59-166        (named arguments -> plain members with the same names)
//   AlignerParams      src/LocARNA/aligner_params.hh:49-116 (named arguments -> chained setters with the same names)
//   Aligner            src/LocARNA/aligner.hh:67-189        construct, align(), trace(), get_alignment() (aligner.hh:56)
//   Alignment          src/LocARNA/alignment.hh:84-281
//   MultipleAlignment  src/LocARNA/multiple_alignment.hh    (from an Alignment; CLUSTAL writer)
//   infty_score_t      src/LocARNA/infty_int.hh             (finite value or -inf; prints "-inf")
//   failure            src/LocARNA/aux.hh:160-209
//
// A maintainer switching a translation unit over writes `namespace LocARNA = LocARNA_B200;`.
#ifndef LOCARNA_B200_HH
#define LOCARNA_B200_HH

#include <algorithm>
#include <cmath>
#include <cstring>
#include <exception>
#include <fstream>
#include <memory>
#include <iostream>
#include <ostream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <queue>
#include <vector>

#include "locarna_b200.h"

namespace LocARNA_B200 {

class failure : public std::exception {
    std::string msg_;
public:
    explicit failure(const std::string &msg) : msg_(msg) {}
    const char *what() const noexcept override { return msg_.c_str(); }
};

class infty_score_t {
    long val_ = 0;
    bool neg_inf_ = false;
public:
    infty_score_t() {}
    explicit infty_score_t(long v) : val_(v) {}
    static infty_score_t neg_infty_value() { infty_score_t s; s.neg_inf_ = true; return s; }
    bool is_neg_infty() const { return neg_inf_; }
    bool is_finite() const { return !neg_inf_; }
    long finite_value() const { return val_; }
    //! order of the reference's InftyInt (infty_int.hh:419-447): -inf below every finite value
    bool operator<(const infty_score_t &o) const { return neg_inf_ ? !o.neg_inf_ : (!o.neg_inf_ && val_ < o.val_); }
};
inline std::ostream &operator<<(std::ostream &out, const infty_score_t &s) {  // infty_int.cc:33-43
    if (s.is_neg_infty()) return out << "-inf";
    return out << s.finite_value();
}

// one GPU context shared by the objects of a program (the reference has no such object: state lives in the classes)
class Context {
    lb200_ctx *ctx_ = nullptr;
public:
    explicit Context(int device = 0) {
        if (lb200_ctx_create(device, &ctx_) != LB200_OK) throw failure("locarna_b200: cannot create a CUDA context (no CPU fallback)");
    }
    ~Context() { lb200_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    lb200_ctx *get() const { return ctx_; }
    void check(int rc) const { if (rc < 0) throw failure(lb200_last_error(ctx_)); }
};

struct ScoringParams {  // scoring.hh:59-166 (defaults of the locarna CLI)
    int match = 50, mismatch = 0, indel = -150, indel_opening = -750, unpaired_penalty = 0;
    int struct_weight = 200, tau_factor = 50, exclusion = 0, temperature_alipf = 300;
    bool use_ribosum = true, stacking = false, new_stacking = false, mea_scoring = false;
    std::string ribosum_file;   // --ribosum-file: a matrix in extended ribosum format; empty or "RIBOSUM85_60": the built-in one
    //! background probability of a base pair (ScoringParams::exp_probA / exp_probB, scoring.hh:123-127); < 0: the CLI's default
    //! 1/(2 len) per sequence (locarna.cc:662-663). One value serves both sequences, as with `locarna --exp-prob`.
    double exp_prob = -1.0;
};

class RnaData {  // PP 2.0 input (rna_data.cc:984-1103); p_bpcut as in RnaData(file, p_bpcut, ...)
    std::string file_;
    double p_bpcut_;
    double max_bps_length_ratio_;
    int max_bp_span_;
public:
    //! as RnaData(filename, p_bpcut, max_bps_length_ratio, pfoldparams) (rna_data.hh:102-106); max_bp_span stands for
    //! PFoldParams::max_bp_span, -1 = unrestricted
    RnaData(const std::string &file, double p_bpcut, double max_bps_length_ratio = 0.0, int max_bp_span = -1)
        : file_(file), p_bpcut_(p_bpcut), max_bps_length_ratio_(max_bps_length_ratio), max_bp_span_(max_bp_span) {}
    double max_bps_length_ratio() const { return max_bps_length_ratio_; }
    const std::string &filename() const { return file_; }
    double arc_cutoff_prob() const { return p_bpcut_; }
    int max_bp_span() const { return max_bp_span_; }
};

class Alignment {  // alignment.hh:84-281
    friend class Aligner;
    std::string nameA_, nameB_, seqA_, seqB_;
    std::vector<std::pair<int, int>> edges_;  // position or -1 (gap), in order
    std::string strA_, strB_;                 // per position, '.', '(' or ')'
    std::vector<std::string> anchorsA_, anchorsB_;   // "#A<k>" annotation rows of the inputs (empty: none)
    // all rows (name, aligned string) of the two inputs: one each for single sequences, several for profile input (Sequence with
    // several rows, sequence.hh:24-80); nameA_ / seqA_ are row 0
    std::vector<std::pair<std::string, std::string>> rowsA_, rowsB_;
public:
    typedef std::vector<std::pair<int, int>> edges_t;
    size_t num_rowsA() const { return rowsA_.empty() ? 1 : rowsA_.size(); }
    size_t num_rowsB() const { return rowsB_.empty() ? 1 : rowsB_.size(); }
    std::string row_nameA(size_t k) const { return rowsA_.empty() ? nameA_ : rowsA_[k].first; }
    std::string row_nameB(size_t k) const { return rowsB_.empty() ? nameB_ : rowsB_[k].first; }
    std::string rowA(size_t k, bool only_local) const { return project(rowsA_.empty() ? seqA_ : rowsA_[k].second, true, only_local); }
    std::string rowB(size_t k, bool only_local) const { return project(rowsB_.empty() ? seqB_ : rowsB_[k].second, false, only_local); }
    const std::vector<std::string> &anchorsA() const { return anchorsA_; }
    const std::vector<std::string> &anchorsB() const { return anchorsB_; }
    // alignment_edges(only_local) (alignment.cc:120-167): locality gaps are reported as -3
    edges_t alignment_edges(bool only_local) const {
        edges_t res;
        int lastA = 1, lastB = 1;
        for (const auto &e : edges_) {
            if (e.first > 0) for (; lastA < e.first; lastA++) if (!only_local) res.emplace_back(lastA, -3);
            if (e.second > 0) for (; lastB < e.second; lastB++) if (!only_local) res.emplace_back(-3, lastB);
            if (e.first > 0) lastA++;
            if (e.second > 0) lastB++;
            res.push_back(e);
        }
        if (!only_local) {
            for (; lastA <= (int)seqA_.size(); lastA++) res.emplace_back(lastA, -3);
            for (; lastB <= (int)seqB_.size(); lastB++) res.emplace_back(-3, lastB);
        }
        return res;
    }
    std::string dot_bracket_structureA(bool only_local) const { return project(strA_, true, only_local); }
    std::string dot_bracket_structureB(bool only_local) const { return project(strB_, false, only_local); }
    std::string rowA(bool only_local) const { return project(seqA_, true, only_local); }
    std::string rowB(bool only_local) const { return project(seqB_, false, only_local); }
    const std::string &nameA() const { return nameA_; }
    const std::string &nameB() const { return nameB_; }
    bool empty() const { return edges_.empty(); }
    //! first / last aligned position of each sequence (alignment.cc:169-195); (lenA, lenB) resp. (0, 0) if there is none
    std::pair<size_t, size_t> start_positions() const {
        size_t a = seqA_.size(), b = seqB_.size();
        for (const auto &e : edges_) if (e.first > 0) { a = (size_t)e.first; break; }
        for (const auto &e : edges_) if (e.second > 0) { b = (size_t)e.second; break; }
        return std::make_pair(a, b);
    }
    std::pair<size_t, size_t> end_positions() const {
        size_t a = 0, b = 0;
        for (auto it = edges_.rbegin(); it != edges_.rend(); ++it) if (it->first > 0) { a = (size_t)it->first; break; }
        for (auto it = edges_.rbegin(); it != edges_.rend(); ++it) if (it->second > 0) { b = (size_t)it->second; break; }
        return std::make_pair(a, b);
    }
private:
    std::string project(const std::string &s, bool first, bool only_local) const {  // alignment.cc:215-230, aux.cc:23-39
        std::string out;
        for (const auto &e : alignment_edges(only_local)) {
            const int p = first ? e.first : e.second;
            out += p > 0 ? s[p - 1] : '-';
        }
        return out;
    }
};

class MultipleAlignment {  // rows of a pairwise Alignment + CLUSTAL writer (multiple_alignment.cc:153-249, :1007-1086)
public:
    struct SeqEntry { std::string name, seq; SeqEntry(const std::string &n, const std::string &s) : name(n), seq(s) {} };
    enum class FormatType { CLUSTAL, STOCKHOLM };
    MultipleAlignment() {}
    //! CLUSTAL W file (multiple_alignment.cc:253-340): header line, blocks of "name row" lines; annotation ('#...'), conservation
    //! (leading blank) and empty lines are skipped
    explicit MultipleAlignment(const std::string &file) {
        std::ifstream in(file.c_str());
        if (!in) throw failure("Cannot read " + file);
        std::string line;
        bool header = false;
        while (std::getline(in, line)) {
            if (!header) { if (line.compare(0, 7, "CLUSTAL") == 0) header = true; else if (!line.empty()) throw failure("missing CLUSTAL header in " + file); continue; }
            if (line.empty() || line[0] == ' ' || line[0] == '#' || line[0] == '/') continue;
            std::istringstream is(line);
            std::string name, row;
            if (!(is >> name >> row)) continue;
            bool found = false;
            for (auto &r : rows_) if (r.name == name) { r.seq += row; found = true; }
            if (!found) rows_.emplace_back(name, row);
        }
    }
    //! pairwise alignment from two rows (multiple_alignment.cc:132-151)
    MultipleAlignment(const std::string &nameA, const std::string &nameB, const std::string &aliA, const std::string &aliB) : pairwise_(true) {
        rows_.emplace_back(nameA, aliA);
        rows_.emplace_back(nameB, aliB);
    }
    //! built from two rows: row 0 belongs to the first, row 1 to the second sequence whatever their names
    bool pairwise() const { return pairwise_; }
    const SeqEntry &seqentry(size_t k) const { return rows_.at(k); }
    bool contains(const std::string &name) const { for (const auto &r : rows_) if (r.name == name) return true; return false; }
    const SeqEntry &seqentry(const std::string &name) const {
        for (const auto &r : rows_) if (r.name == name) return r;
        throw failure("MultipleAlignment: no sequence " + name);
    }
    MultipleAlignment(const Alignment &a, bool only_local = false) {
        // all rows of A, then all rows of B; "A." / "B." prefixes if a row name of A occurs in B (multiple_alignment.cc:184-247)
        bool clash = false;
        for (size_t k = 0; k < a.num_rowsA(); k++) for (size_t l = 0; l < a.num_rowsB(); l++) clash |= a.row_nameA(k) == a.row_nameB(l);
        for (size_t k = 0; k < a.num_rowsA(); k++) rows_.emplace_back(clash ? "A." + a.row_nameA(k) : a.row_nameA(k), a.rowA(k, only_local));
        for (size_t l = 0; l < a.num_rowsB(); l++) rows_.emplace_back(clash ? "B." + a.row_nameB(l) : a.row_nameB(l), a.rowB(l, only_local));
        // consensus anchor annotation (multiple_alignment.cc:157-165, sequence_annotation.cc:12-50): per column the name of the
        // non-gap side, of the named side, or the smaller of two names; dropped if a name would occur twice
        const auto &A = a.anchorsA(), &B = a.anchorsB();
        if (!A.empty() && !B.empty() && A.size() == B.size()) {
            std::vector<std::string> cons(A.size());
            auto name = [](const std::vector<std::string> &rows, int pos) { std::string n; for (const auto &r : rows) n += r[(size_t)pos - 1]; return n; };
            auto neutral = [](const std::string &n) { for (char c : n) if (!(c == ' ' || c == '.')) return false; return true; };
            std::vector<std::string> names;
            for (const auto &e : a.alignment_edges(only_local)) {
                std::string n;
                if (e.first <= 0) n = name(B, e.second);
                else if (e.second <= 0) n = name(A, e.first);
                else {
                    const std::string na = name(A, e.first), nb = name(B, e.second);
                    n = neutral(na) ? nb : neutral(nb) ? na : std::min(na, nb);
                }
                names.push_back(n);
                for (size_t k = 0; k < cons.size(); k++) cons[k] += n[k];
            }
            bool dup = false;
            for (size_t x = 0; x < names.size() && !dup; x++)
                if (!neutral(names[x])) for (size_t y = x + 1; y < names.size(); y++) if (!neutral(names[y]) && names[x] == names[y]) { dup = true; break; }
            if (!dup) anchors_ = cons;
        }
    }
    void prepend(const SeqEntry &e) { rows_.insert(rows_.begin(), e); }
    void append(const SeqEntry &e) { rows_.push_back(e); }
    size_t length() const { return rows_.empty() ? 0 : rows_[0].seq.size(); }
    size_t num_of_rows() const { return rows_.size(); }
    std::ostream &write(std::ostream &out, size_t width, FormatType format = FormatType::CLUSTAL) const {
        size_t namewidth = 18;
        for (const auto &r : rows_) namewidth = std::max(namewidth, r.name.size());
        size_t start = 0;
        do {
            const size_t end = std::min(length(), start + width);
            for (const auto &r : rows_) {
                std::string name = r.name;
                name.resize(namewidth, ' ');
                out << name << " " << r.seq.substr(start, end - start) << std::endl;
            }
            for (size_t k = 0; k < anchors_.size(); k++) {   // multi-line anchor annotation (multiple_alignment.cc:1036-1050)
                std::string name = (format == FormatType::STOCKHOLM ? "#=GC cA" : "#A") + std::to_string(k + 1);
                if (name.size() < namewidth) name.resize(namewidth, ' ');
                out << name << " " << anchors_[k].substr(start, end - start) << std::endl;
            }
            start = end;
        } while (start < length() && out << std::endl);
        if (format == FormatType::STOCKHOLM) out << "//" << std::endl;   // end marker (multiple_alignment.cc:1079-1082)
        return out;
    }
private:
    std::vector<SeqEntry> rows_;
    std::vector<std::string> anchors_;   // consensus anchor annotation rows
    bool pairwise_ = false;
};

//! One arc (basepairs.hh:33-89): positions only; the reference's arc index is internal to the device tables.
class Arc {
    int left_, right_;
public:
    Arc(int l, int r) : left_(l), right_(r) {}
    int left() const { return left_; }
    int right() const { return right_; }
};

//! arc_matches.hh:37-102
class ArcMatch {
    Arc a_, b_;
    size_t idx_;
public:
    ArcMatch(const Arc &a, const Arc &b, size_t idx) : a_(a), b_(b), idx_(idx) {}
    const Arc &arcA() const { return a_; }
    const Arc &arcB() const { return b_; }
    size_t idx() const { return idx_; }
};

//! Read-only view of the arc matches the device builder enumerated for one pair, in the reference's index order
//! (ArcMatches, arc_matches.hh:114-504; construction arc_matches.cc:130-188), with Scoring::arcmatch of each (scoring.cc:540-554).
//! Obtained from Aligner::arc_matches(); the reference builds it on the host and hands it to Scoring / Aligner instead.
class ArcMatches {
    friend class Aligner;
    std::vector<ArcMatch> ams_;
    std::vector<int> scores_;
public:
    size_t num_arc_matches() const { return ams_.size(); }
    const ArcMatch &arcmatch(size_t idx) const { return ams_.at(idx); }
    //! Scoring::arcmatch(am) (non-stacked)
    long get_score(const ArcMatch &am) const { return scores_.at(am.idx()); }
    bool explicit_scores() const { return false; }
    std::vector<ArcMatch>::const_iterator begin() const { return ams_.begin(); }
    std::vector<ArcMatch>::const_iterator end() const { return ams_.end(); }
    //! arc_matches.cc:285-311: lines "al ar bl br score"
    void write_arcmatch_scores(const std::string &arcmatch_scores_file) const {
        std::ofstream out(arcmatch_scores_file.c_str());
        if (!out.is_open()) throw failure("Cannot open file " + arcmatch_scores_file + " for writing arcmatch-scores.");
        for (const ArcMatch &am : ams_)
            out << am.arcA().left() << " " << am.arcA().right() << " " << am.arcB().left() << " " << am.arcB().right() << " " << scores_[am.idx()] << "\n";
    }
};

//! Restriction of the top level to subsequences (aligner_restriction.hh:26-130)
class AlignerRestriction {
    int startA_, startB_, endA_, endB_;
public:
    AlignerRestriction(int startA, int startB, int endA, int endB) : startA_(startA), startB_(startB), endA_(endA), endB_(endB) {}
    size_t startA() const { return startA_; }
    size_t endA() const { return endA_; }
    size_t startB() const { return startB_; }
    size_t endB() const { return endB_; }
    void set_startA(size_t p) { startA_ = (int)p; }
    void set_endA(size_t p) { endA_ = (int)p; }
    void set_startB(size_t p) { startB_ = (int)p; }
    void set_endB(size_t p) { endB_ = (int)p; }
};
inline std::ostream &operator<<(std::ostream &out, const AlignerRestriction &r) {   // aligner_restriction.hh:137-141
    return out << r.startA() << " " << r.startB() << " " << r.endA() << " " << r.endB();
}

class AlignerParams {  // aligner_params.hh:51-115: same argument names, chained setters instead of named-argument objects
    friend class Aligner;
    const RnaData *rnaA_ = nullptr, *rnaB_ = nullptr;
    ScoringParams scoring_;
    bool no_lonely_pairs_ = false, struct_local_ = false, sequ_local_ = false, stacking_ = false;
    std::string free_endgaps_ = "----";
    int max_diff_am_ = -1, max_diff_at_am_ = -1, max_diff_ = -1;
    double min_prob_ = 0.001, min_trace_probability_ = 1e-4;
    std::vector<int> min_col_, max_col_;
    const MultipleAlignment *ref_aln_ = nullptr;
    bool ref_relaxed_ = false;
public:
    AlignerParams &seqA(const RnaData *r) { rnaA_ = r; return *this; }
    AlignerParams &seqB(const RnaData *r) { rnaB_ = r; return *this; }
    AlignerParams &scoring(const ScoringParams &s) { scoring_ = s; return *this; }
    AlignerParams &no_lonely_pairs(bool b) { no_lonely_pairs_ = b; return *this; }
    AlignerParams &struct_local(bool b) { struct_local_ = b; return *this; }
    AlignerParams &sequ_local(bool b) { sequ_local_ = b; return *this; }
    AlignerParams &free_endgaps(const std::string &d) { free_endgaps_ = d; return *this; }
    AlignerParams &max_diff_am(int d) { max_diff_am_ = d; return *this; }
    AlignerParams &max_diff_at_am(int d) { max_diff_at_am_ = d; return *this; }
    AlignerParams &stacking(bool b) { stacking_ = b; return *this; }
    // the reference passes a TraceController; here either its rows (min_col/max_col) or the two numbers it is built from
    AlignerParams &trace_controller(const std::vector<int> &min_col, const std::vector<int> &max_col) { min_col_ = min_col; max_col_ = max_col; return *this; }
    //! reference alignment of TraceController(seqA, seqB, ma, max_diff) (trace_controller.cc:406-539): the band lies within max_diff of it
    AlignerParams &reference_alignment(const MultipleAlignment *ma, bool relaxed_merging = false) { ref_aln_ = ma; ref_relaxed_ = relaxed_merging; return *this; }
    AlignerParams &max_diff(int d) { max_diff_ = d; return *this; }
    AlignerParams &min_trace_probability(double p) { min_trace_probability_ = p; return *this; }
    AlignerParams &min_prob(double p) { min_prob_ = p; return *this; }
};

class Aligner {  // aligner.hh:67-189
    std::shared_ptr<Context> ctx_;
    int pair_ = -1;
    bool traced_ = false, have_ams_ = false, restricted_ = false;
    int seq_a_ = -1, seq_b_ = -1;
    AlignerRestriction r_{1, 1, 0, 0};
    Alignment alignment_;
    ArcMatches ams_;
public:
    explicit Aligner(const AlignerParams &ap, int device = 0) : ctx_(std::make_shared<Context>(device)) {
        if (!ap.rnaA_ || !ap.rnaB_) throw failure("AlignerParams: seqA and seqB are mandatory");
        if (ap.scoring_.mea_scoring) throw failure("locarna_b200: MEA scoring is not supported");
        lb200_params p;
        lb200_default_params(&p);
        const ScoringParams &s = ap.scoring_;
        p.min_prob = ap.min_prob_; p.max_diff_am = ap.max_diff_am_; p.max_diff_at_am = ap.max_diff_at_am_; p.max_diff = ap.max_diff_;
        p.min_trace_probability = ap.min_trace_probability_;
        p.struct_weight = s.struct_weight; p.indel = s.indel; p.indel_opening = s.indel_opening; p.tau = s.tau_factor;
        p.exclusion = s.exclusion; p.match = s.match; p.mismatch = s.mismatch; p.use_ribosum = s.use_ribosum;
        p.unpaired_penalty = s.unpaired_penalty; p.temperature_alipf = s.temperature_alipf;
        p.stacking = s.stacking; p.new_stacking = s.new_stacking;   // Scoring::stacking() decides (scoring.hh:652-655); AlignerParams::stacking is unused by the reference's aligner
        if (ap.rnaA_->max_bp_span() != ap.rnaB_->max_bp_span() || ap.rnaA_->max_bps_length_ratio() != ap.rnaB_->max_bps_length_ratio())
            throw failure("locarna_b200: both RnaData objects must use the same max_bp_span and max_bps_length_ratio");
        p.exp_prob = s.exp_prob; p.max_bp_span = ap.rnaA_->max_bp_span(); p.max_bps_length_ratio = ap.rnaA_->max_bps_length_ratio();
        p.no_lonely_pairs = ap.no_lonely_pairs_; p.struct_local = ap.struct_local_; p.sequ_local = ap.sequ_local_;
        strncpy(p.free_endgaps, ap.free_endgaps_.c_str(), sizeof(p.free_endgaps) - 1);
        ctx_->check(lb200_set_ribosum_file(ctx_->get(), s.ribosum_file.empty() ? nullptr : s.ribosum_file.c_str()));
        ctx_->check(lb200_set_params(ctx_->get(), &p));
        const int a = lb200_seq_add_pp(ctx_->get(), ap.rnaA_->filename().c_str());
        ctx_->check(a);
        const int b = lb200_seq_add_pp(ctx_->get(), ap.rnaB_->filename().c_str());
        ctx_->check(b);
        seq_a_ = a; seq_b_ = b;
        char name[256], *seq;
        const int la = lb200_seq_length(ctx_->get(), a), lb = lb200_seq_length(ctx_->get(), b);
        const int seq_cap = std::max(la, lb) + 1;
        seq = new char[seq_cap];
        ctx_->check(lb200_seq_get(ctx_->get(), a, name, sizeof name, seq, seq_cap)); alignment_.nameA_ = name; alignment_.seqA_ = seq;
        ctx_->check(lb200_seq_get(ctx_->get(), b, name, sizeof name, seq, seq_cap)); alignment_.nameB_ = name; alignment_.seqB_ = seq;
        for (int which = 0; which < 2; which++) {   // rows of profile inputs (one row: a single sequence)
            const int id = which ? b : a;
            auto &rows = which ? alignment_.rowsB_ : alignment_.rowsA_;
            const int nr = lb200_seq_num_rows(ctx_->get(), id);
            ctx_->check(nr);
            for (int k = 0; k < nr; k++) {
                ctx_->check(lb200_seq_get_row(ctx_->get(), id, k, name, sizeof name, seq, seq_cap));
                rows.emplace_back(std::string(name), std::string(seq));
            }
        }
        delete[] seq;
        for (int which = 0; which < 2; which++) {   // "#A<k>" rows of the inputs, for the consensus annotation of the output
            const int id = which ? b : a;
            std::vector<char> buf((size_t)lb200_seq_anchors(ctx_->get(), id, nullptr, 0) + 1);
            lb200_seq_anchors(ctx_->get(), id, buf.data(), (int)buf.size());
            std::vector<std::string> &rows = which ? alignment_.anchorsB_ : alignment_.anchorsA_;
            std::string cur;
            for (const char *p = buf.data(); ; p++) { if (*p == '#' || *p == 0) { if (!cur.empty()) rows.push_back(cur); cur.clear(); if (*p == 0) break; } else cur += *p; }
        }
        if (ap.ref_aln_ != nullptr && ap.max_diff_ != -1) {
            // (for profile input the reference merges the trace ranges of all row pairs, trace_controller.cc:426-511: not built)
            if (alignment_.num_rowsA() > 1 || alignment_.num_rowsB() > 1)
                throw failure("a band around a reference alignment (--max-diff-aln / --max-diff-pw-aln) is not supported for profile (multi-row) input");
            // TraceController(seqA, seqB, ma, delta): rows within delta of the reference alignment (trace_controller.cc:431-483); the
            // probability envelope is applied inside this range
            std::vector<int> lo((size_t)la + 1), hi((size_t)la + 1);
            const MultipleAlignment &ma = *ap.ref_aln_;
            const std::string &rowA = ma.pairwise() ? ma.seqentry((size_t)0).seq : ma.seqentry(alignment_.nameA_).seq;
            const std::string &rowB = ma.pairwise() ? ma.seqentry((size_t)1).seq : ma.seqentry(alignment_.nameB_).seq;
            if (lb200_band_from_alignment(la, lb, rowA.c_str(), rowB.c_str(),
                                          ap.max_diff_, ap.ref_relaxed_ ? 1 : 0, lo.data(), hi.data()) != LB200_OK)
                throw failure("Inconsistent trace range due to max-diff heuristic");
            pair_ = lb200_pair_add_restricted(ctx_->get(), a, b, lo.data(), hi.data());
        } else {
            pair_ = lb200_pair_add(ctx_->get(), a, b, ap.min_col_.empty() ? nullptr : ap.min_col_.data(), ap.max_col_.empty() ? nullptr : ap.max_col_.data());
        }
        ctx_->check(pair_);
        r_ = AlignerRestriction(1, 1, la, lb);
    }
    //! compute the alignment score (aligner.cc:924-962); under a restriction only the top level is redone on the filled D table
    infty_score_t align() {
        if (restricted_) ctx_->check(lb200_run_pair_toplevel(ctx_->get(), pair_, LB200_TOP_PLAIN, 0, LB200_RUN_TRACE));
        else ctx_->check(lb200_run(ctx_->get(), LB200_RUN_TRACE));
        traced_ = true;
        return result_score();
    }
    //! normalized local alignment by Dinkelbach's algorithm (aligner.cc:1522-1597); returns the normalized score, the alignment is traced
    infty_score_t normalized_align(long L, bool /*verbose*/ = false) {
        if (restricted_) ctx_->check(lb200_run_pair_toplevel(ctx_->get(), pair_, LB200_TOP_NORMALIZED, (int64_t)L, LB200_RUN_TRACE));
        else ctx_->check(lb200_run_normalized(ctx_->get(), (int64_t)L));
        return modified_result();
    }
    //! alignment with every aligned position penalized (aligner.cc:1599-1622); returns the penalized score, the alignment is traced
    infty_score_t penalized_align(long position_penalty) {
        if (restricted_) ctx_->check(lb200_run_pair_toplevel(ctx_->get(), pair_, LB200_TOP_PENALIZED, (int64_t)position_penalty, LB200_RUN_TRACE));
        else ctx_->check(lb200_run_penalized(ctx_->get(), (int64_t)position_penalty));
        return modified_result();
    }
    //! restrict the top level to subsequences (aligner.cc:1368-1376); the D table is not affected
    void set_restriction(const AlignerRestriction &r) {
        ctx_->check(lb200_pair_set_restriction(ctx_->get(), pair_, (int)r.startA(), (int)r.startB(), (int)r.endA(), (int)r.endB()));
        r_ = r;
        restricted_ = true;
    }
    const AlignerRestriction &get_restriction() const { return r_; }
    //! k-best alignments by interval splitting (aligner.cc:1383-1514): same task queue, same splits, same output
    void suboptimal(int k, long threshold, bool normalized, long normalized_L, size_t /*output_width*/, bool verbose, bool /*opt_local_output*/,
                    bool opt_pos_output, bool /*opt_write_structure*/) {
        typedef std::pair<AlignerRestriction, infty_score_t> task_t;
        struct greater_second { bool operator()(const task_t &a, const task_t &b) const { return a.second < b.second; } };   // aligner.hh:34-45
        Aligner &a = *this;
        infty_score_t a_score = !normalized ? a.align() : a.normalized_align(normalized_L, false);
        std::priority_queue<task_t, std::vector<task_t>, greater_second> tasks;
        tasks.push(task_t(a.get_restriction(), a_score));
        size_t i = 1;
        while (k < 0 || i <= (size_t)k) {
            task_t task = tasks.top();
            tasks.pop();
            AlignerRestriction &task_r = task.first;
            const infty_score_t task_score = task.second;
            if (task_score < infty_score_t(threshold + 1)) break;
            a.set_restriction(task_r);
            if (!normalized) { a.align(); a.trace(); }
            else a.normalized_align(normalized_L, verbose);
            Alignment alignment = a.get_alignment();
            if (alignment.empty()) continue;
            if (opt_pos_output) {
                std::cout << "HIT " << task_score << " " << alignment.start_positions().first << " " << alignment.start_positions().first << " "
                          << alignment.end_positions().second << " " << alignment.end_positions().second << " " << std::endl;
            } else {
                MultipleAlignment ma(alignment, true);
                std::cout << "Score: " << task_score << std::endl;
                ma.write(std::cout, 120, MultipleAlignment::FormatType::CLUSTAL);
            }
            if (!opt_pos_output) std::cout << std::endl << std::endl;
            if (k >= 0 && i == (size_t)k) break;
            const size_t lenA = task_r.endA() - task_r.startA(), lenB = task_r.endB() - task_r.startB();
            AlignerRestriction r1(task_r), r2(task_r);
            if (lenA > lenB) {
                const int splitA = (int)((alignment.start_positions().first + alignment.end_positions().first) / 2);
                if (verbose) std::cout << "Split A at " << splitA << std::endl;
                r1.set_endA(splitA); r2.set_startA(splitA);
            } else {
                const int splitB = (int)((alignment.start_positions().second + alignment.end_positions().second) / 2);
                if (verbose) std::cout << "Split B at " << splitB << std::endl;
                r1.set_endB(splitB); r2.set_startB(splitB);
            }
            a.set_restriction(r1);
            const infty_score_t a1_score = !normalized ? a.align() : a.normalized_align(normalized_L, false);
            a.set_restriction(r2);
            const infty_score_t a2_score = !normalized ? a.align() : a.normalized_align(normalized_L, false);
            tasks.push(task_t(r1, a1_score));
            tasks.push(task_t(r2, a2_score));
            ++i;
        }
    }
    //! trace back (aligner.cc:1345-1363); the device already traced during align()
    void trace() {
        if (!traced_) align();
        lb200_pair_info inf;
        ctx_->check(lb200_pair_get_info(ctx_->get(), pair_, &inf));
        std::vector<int> ea(inf.n_edges + 1), eb(inf.n_edges + 1);
        std::string sa(inf.lenA + 1, '\0'), sb(inf.lenB + 1, '\0');
        ctx_->check(lb200_pair_alignment(ctx_->get(), pair_, ea.data(), eb.data(), &sa[0], &sb[0]));
        alignment_.edges_.clear();
        for (int64_t k = 0; k < inf.n_edges; k++) alignment_.edges_.emplace_back(ea[k], eb[k]);
        alignment_.strA_ = sa.substr(0, inf.lenA); alignment_.strB_ = sb.substr(0, inf.lenB);
    }
    const Alignment &get_alignment() const { return alignment_; }
    //! Alignment with its consensus dot plot in PP 2.0 format (`locarna --pp`; MainHelper::consensus main_helper.icc:472-530, the consensus
    //! constructor of RnaData rna_data.cc:104-126 / :1474-1548, RnaData::write_pp :1242-1350). exp_prob < 0: 1 / (2 len) per sequence
    //! (locarna.cc:662-663). The base pairs are written in the iteration order of the reference's hash map (sparse_vector_base.hh:28,
    //! aux.hh:24-34): the same container type with the same hash, filled in the same order.
    void write_pp(std::ostream &out, bool only_local, double exp_prob) const {
        struct PairHash { size_t operator()(const std::pair<size_t, size_t> &p) const noexcept { return std::hash<size_t>()(p.first) ^ (std::hash<size_t>()(p.second) << 1); } };
        typedef std::unordered_map<std::pair<size_t, size_t>, double, PairHash> sparse_t;
        struct Probs { std::vector<int> i, j; std::vector<double> p, p2; double cutoff = 0; int stacking = 0;
            double get(const std::vector<double> &v, int a, int b) const {
                size_t lo = 0, hi = i.size();
                while (lo < hi) { const size_t mid = (lo + hi) / 2; if (i[mid] < a || (i[mid] == a && j[mid] < b)) lo = mid + 1; else hi = mid; }
                return (lo < i.size() && i[lo] == a && j[lo] == b) ? v[lo] : 0.0;
            } } P[2];
        for (int w = 0; w < 2; w++) {
            const int id = w ? seq_b_ : seq_a_;
            const int64_t n = lb200_seq_pairs(ctx_->get(), id, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            ctx_->check((int)std::min<int64_t>(n, 0));
            P[w].i.resize((size_t)n + 1); P[w].j.resize((size_t)n + 1); P[w].p.resize((size_t)n + 1); P[w].p2.resize((size_t)n + 1);
            lb200_seq_pairs(ctx_->get(), id, P[w].i.data(), P[w].j.data(), P[w].p.data(), P[w].p2.data(), &P[w].cutoff, &P[w].stacking);
            P[w].i.resize((size_t)n); P[w].j.resize((size_t)n); P[w].p.resize((size_t)n); P[w].p2.resize((size_t)n);
        }
        const double p_expA = exp_prob < 0 ? 1.0 / (2.0 * alignment_.seqA_.size()) : exp_prob, p_expB = exp_prob < 0 ? 1.0 / (2.0 * alignment_.seqB_.size()) : exp_prob;
        const bool stacking = P[0].stacking && P[1].stacking;
        const size_t rowsA = alignment_.num_rowsA(), rowsB = alignment_.num_rowsB();   // weights of the two inputs (rna_data.cc:1482-1489)
        const double p_minMean = std::exp((std::log(P[0].cutoff) * rowsA + std::log(P[1].cutoff) * rowsB) / (rowsA + rowsB));
        const double p_penalty = p_minMean * 0.1;
        auto consensus_probability = [&](double pA, double pB) {   // rna_data.cc:1550-1578: weighted geometric mean
            pA = std::max(std::min(p_expA, p_penalty), pA);
            pB = std::max(std::min(p_expB, p_penalty), pB);
            return std::exp((std::log(pA) * rowsA + std::log(pB) * rowsB) / (rowsA + rowsB));
        };
        const Alignment::edges_t edges = alignment_.alignment_edges(only_local);
        sparse_t arc_probs;
        for (size_t i = 0; i < edges.size(); i++)
            for (size_t j = i + 1; j < edges.size(); j++) {
                const auto &ei = edges[i], &ej = edges[j];
                const bool gapA = ei.first <= 0 || ej.first <= 0, gapB = ei.second <= 0 || ej.second <= 0;
                const double p = consensus_probability(gapA ? 0 : P[0].get(P[0].p, ei.first, ej.first), gapB ? 0 : P[1].get(P[1].p, ei.second, ej.second));
                if (stacking) {
                    const double st_p = consensus_probability(gapA ? 0 : P[0].get(P[0].p2, ei.first, ej.first), gapB ? 0 : P[1].get(P[1].p2, ei.second, ej.second));
                    // with joint probabilities on both sides a pair is also kept for its stacked probability; the consensus object itself
                    // never reports has_stacking (rna_data.cc:117), so neither "#STACK" nor the fourth column is written
                    if ((p > p_minMean || st_p > p_minMean) && p != 0.0) arc_probs[std::make_pair(i + 1, j + 1)] = p;
                } else if (p > p_minMean) arc_probs.insert(std::make_pair(std::make_pair(i + 1, j + 1), p));
            }
        auto format_prob = [](double prob) {   // rna_data.cc:1279-1304
            std::ostringstream outd;
            outd.precision(4);
            outd << prob;
            std::string t = outd.str();
            if (t.length() > 4 + 4) { std::ostringstream outs; outs.setf(std::ios::scientific, std::ios::floatfield); outs.precision(3); outs << prob; t = outs.str(); }
            const size_t pos = t.find("e-0");
            if (pos != std::string::npos) t.replace(pos, 3, "e-");
            return t;
        };
        out << "#PP 2.0" << std::endl << std::endl;
        MultipleAlignment ma(alignment_, only_local);
        ma.write(out, (size_t)-1, MultipleAlignment::FormatType::CLUSTAL);
        out << std::endl << "#END" << std::endl;
        out << std::endl << "#SECTION BASEPAIRS" << std::endl << std::endl << "#BPCUT " << format_prob(std::max(p_minMean, 0.0)) << std::endl;
        out << std::endl;
        for (const auto &x : arc_probs) {
            if (!(x.second > 0.0)) continue;
            out << x.first.first << " " << x.first.second << " " << format_prob(x.second);
            out << std::endl;
        }
        out << std::endl << "#END" << std::endl;
    }
private:
    infty_score_t result_score() {
        int64_t sc = 0;
        ctx_->check(lb200_pair_score(ctx_->get(), pair_, &sc));
        return sc == LB200_SCORE_NEG_INF ? infty_score_t::neg_infty_value() : infty_score_t((long)sc);
    }
    infty_score_t modified_result() {
        traced_ = true;
        trace();
        return result_score();
    }
public:
    //! arc matches and their scores as built on the device (no alignment is computed: lb200_upload only)
    const ArcMatches &arc_matches() {
        if (!have_ams_) {
            ctx_->check(lb200_upload(ctx_->get()));
            lb200_pair_info inf;
            ctx_->check(lb200_pair_get_info(ctx_->get(), pair_, &inf));
            const size_t K = (size_t)inf.n_arcmatches;
            std::vector<int> al(K + 1), ar(K + 1), bl(K + 1), br(K + 1), sc(K + 1);
            ctx_->check(lb200_pair_arcmatches(ctx_->get(), pair_, al.data(), ar.data(), bl.data(), br.data(), sc.data(), nullptr));
            ams_.ams_.clear(); ams_.scores_.assign(sc.begin(), sc.begin() + K);
            for (size_t k = 0; k < K; k++) ams_.ams_.emplace_back(Arc(al[k], ar[k]), Arc(bl[k], br[k]), k);
            have_ams_ = true;
        }
        return ams_;
    }
    void band(std::vector<int> &min_col, std::vector<int> &max_col) const {
        lb200_pair_info inf;
        ctx_->check(lb200_pair_get_info(ctx_->get(), pair_, &inf));
        min_col.resize(inf.lenA + 1); max_col.resize(inf.lenA + 1);
        ctx_->check(lb200_pair_band(ctx_->get(), pair_, min_col.data(), max_col.data()));
    }
};

class AlignerPParams {  // aligner_params.hh:122-156 on top of AlignerParams; defaults of the locarna_p CLI (locarna_p.cc:95-175)
    friend class AlignerP;
    const RnaData *rnaA_ = nullptr, *rnaB_ = nullptr;
    ScoringParams scoring_;
    int max_diff_am_ = -1, max_diff_at_am_ = -1, max_diff_ = -1;
    double min_prob_ = 0.001, min_trace_probability_ = 1e-5, min_am_prob_ = 0.001, min_bm_prob_ = 0.001, pf_scale_ = 1.0;
    std::vector<int> min_col_, max_col_;
public:
    AlignerPParams &seqA(const RnaData *r) { rnaA_ = r; return *this; }
    AlignerPParams &seqB(const RnaData *r) { rnaB_ = r; return *this; }
    AlignerPParams &scoring(const ScoringParams &s) { scoring_ = s; return *this; }
    AlignerPParams &max_diff_am(int d) { max_diff_am_ = d; return *this; }
    AlignerPParams &max_diff_at_am(int d) { max_diff_at_am_ = d; return *this; }
    AlignerPParams &trace_controller(const std::vector<int> &min_col, const std::vector<int> &max_col) { min_col_ = min_col; max_col_ = max_col; return *this; }
    AlignerPParams &max_diff(int d) { max_diff_ = d; return *this; }
    AlignerPParams &min_trace_probability(double p) { min_trace_probability_ = p; return *this; }
    AlignerPParams &min_prob(double p) { min_prob_ = p; return *this; }
    AlignerPParams &min_am_prob(double p) { min_am_prob_ = p; return *this; }
    AlignerPParams &min_bm_prob(double p) { min_bm_prob_ = p; return *this; }
    AlignerPParams &pf_scale(double s) { pf_scale_ = s; return *this; }
};

class AlignerP {  // aligner_p.hh:55-629, T = double
    std::shared_ptr<Context> ctx_;
    int pair_ = -1;
    AlignerPParams params_;
    bool probs_ = false;
    std::vector<int> al_, ar_, bl_, br_;
    std::vector<double> am_prob_, bm_prob_;
    int lenA_ = 0, lenB_ = 0;
    void fetch() {
        if (probs_) return;
        ctx_->check(lb200_run_pf_probs(ctx_->get(), params_.pf_scale_, params_.min_am_prob_));
        lb200_pair_info inf;
        ctx_->check(lb200_pair_get_info(ctx_->get(), pair_, &inf));
        const size_t K = (size_t)inf.n_arcmatches;
        al_.resize(K + 1); ar_.resize(K + 1); bl_.resize(K + 1); br_.resize(K + 1); am_prob_.resize(K + 1);
        ctx_->check(lb200_pair_arcmatches(ctx_->get(), pair_, al_.data(), ar_.data(), bl_.data(), br_.data(), nullptr, nullptr));
        ctx_->check(lb200_pair_arcmatch_probs(ctx_->get(), pair_, am_prob_.data()));
        al_.resize(K); am_prob_.resize(K);
        lenA_ = inf.lenA; lenB_ = inf.lenB;
        bm_prob_.resize((size_t)(lenA_ + 1) * (lenB_ + 1));
        ctx_->check(lb200_pair_basematch_probs(ctx_->get(), pair_, bm_prob_.data()));
        probs_ = true;
    }
public:
    explicit AlignerP(const AlignerPParams &ap, int device = 0) : ctx_(std::make_shared<Context>(device)), params_(ap) {
        if (!ap.rnaA_ || !ap.rnaB_) throw failure("AlignerPParams: seqA and seqB are mandatory");
        lb200_params p;
        lb200_default_params(&p);
        const ScoringParams &s = ap.scoring_;
        p.min_prob = ap.min_prob_; p.max_diff_am = ap.max_diff_am_; p.max_diff_at_am = ap.max_diff_at_am_; p.max_diff = ap.max_diff_;
        p.min_trace_probability = ap.min_trace_probability_;
        p.pf_double = 1;                                      // locarna_p.cc:285-294: the envelope is computed in double as well
        p.struct_weight = s.struct_weight; p.indel = s.indel; p.indel_opening = s.indel_opening; p.tau = s.tau_factor;
        p.match = s.match; p.mismatch = s.mismatch; p.use_ribosum = s.use_ribosum; p.temperature_alipf = s.temperature_alipf;
        if (ap.rnaA_->max_bp_span() != ap.rnaB_->max_bp_span() || ap.rnaA_->max_bps_length_ratio() != ap.rnaB_->max_bps_length_ratio())
            throw failure("locarna_b200: both RnaData objects must use the same max_bp_span and max_bps_length_ratio");
        p.exp_prob = s.exp_prob; p.max_bp_span = ap.rnaA_->max_bp_span(); p.max_bps_length_ratio = ap.rnaA_->max_bps_length_ratio();
        ctx_->check(lb200_set_ribosum_file(ctx_->get(), s.ribosum_file.empty() ? nullptr : s.ribosum_file.c_str()));
        ctx_->check(lb200_set_params(ctx_->get(), &p));
        const int a = lb200_seq_add_pp(ctx_->get(), ap.rnaA_->filename().c_str());
        ctx_->check(a);
        const int b = lb200_seq_add_pp(ctx_->get(), ap.rnaB_->filename().c_str());
        ctx_->check(b);
        pair_ = lb200_pair_add(ctx_->get(), a, b, ap.min_col_.empty() ? nullptr : ap.min_col_.data(), ap.max_col_.empty() ? nullptr : ap.max_col_.data());
        ctx_->check(pair_);
    }
    //! inside algorithm; returns the partition function (aligner_p.icc:413-438)
    double align_inside() {
        ctx_->check(lb200_run_pf(ctx_->get(), params_.pf_scale_));
        double z = 0;
        ctx_->check(lb200_pair_partition_function(ctx_->get(), pair_, &z));
        return z;
    }
    //! outside algorithm (aligner_p.icc:1113-1139); the device computes outside and probabilities in one call
    void align_outside() { fetch(); }
    void compute_arcmatch_probabilities() { fetch(); }
    void compute_basematch_probabilities(bool basematch_probs_include_arcmatch) {
        fetch();
        if (basematch_probs_include_arcmatch)                 // aligner_p.icc:1382-1398
            for (size_t k = 0; k < am_prob_.size(); k++) {
                bm_prob_[(size_t)al_[k] * (lenB_ + 1) + bl_[k]] += am_prob_[k];
                bm_prob_[(size_t)ar_[k] * (lenB_ + 1) + br_[k]] += am_prob_[k];
            }
    }
    //! aligner_p.icc:1404-1415
    void write_basematch_probabilities(std::ostream &out) {
        fetch();
        for (int i = 1; i <= lenA_; i++)
            for (int j = 1; j <= lenB_; j++)
                if (bm_prob_[(size_t)i * (lenB_ + 1) + j] >= params_.min_bm_prob_) out << i << " " << j << " " << bm_prob_[(size_t)i * (lenB_ + 1) + j] << std::endl;
    }
    //! aligner_p.icc:1420-1435
    void write_arcmatch_probabilities(std::ostream &out) {
        fetch();
        for (size_t k = 0; k < am_prob_.size(); k++)
            if (am_prob_[k] >= params_.min_am_prob_) out << al_[k] << " " << ar_[k] << " " << bl_[k] << " " << br_[k] << " " << am_prob_[k] << std::endl;
    }
};

}  // namespace LocARNA_B200
#endif
