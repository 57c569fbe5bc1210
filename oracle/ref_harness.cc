// ref_harness -- TEST INFRASTRUCTURE ONLY.
//
// A driver written for this repo around the *reference's own, unmodified* library classes
// (compiled from /root/reference/src by oracle/Makefile; nothing of the reference is copied here).
// It runs the same object pipeline as the reference CLI (locarna.cc:452-784:
// RnaData x2 -> AnchorConstraints -> TraceController (+ probability envelope, main_helper.icc:371-426)
// -> ArcMatches -> Scoring -> Aligner::align() -> trace()) and dumps the intermediates that the
// parity tests compare against, plus per-phase wall times for the CPU baseline.
//
// Usage:   ref_harness [flags] A.pp B.pp
//          ref_harness [flags] --pairs LIST      (LIST: one "A.pp B.pp" per line)
// Output:  line-oriented text, one record per pair (see emit_* below).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "LocARNA/sequence.hh"
#include "LocARNA/basepairs.hh"
#include "LocARNA/alignment.hh"
#include "LocARNA/aligner.hh"
#include "LocARNA/aligner_impl.hh"
#include "LocARNA/aligner_p.hh"
#include "LocARNA/rna_data.hh"
#include "LocARNA/arc_matches.hh"
#include "LocARNA/edge_probs.hh"
#include "LocARNA/ribosum.hh"
#include "LocARNA/ribofit.hh"
#include "LocARNA/anchor_constraints.hh"
#include "LocARNA/trace_controller.hh"
#include "LocARNA/multiple_alignment.hh"
#include "LocARNA/pfold_params.hh"
#include "LocARNA/scoring.hh"
#include "LocARNA/free_endgaps.hh"
#include "LocARNA/ribosum85_60.icc"

using namespace LocARNA;

// read access to the protected per-arc weight tables (scoring.hh:363-364)
struct ScoringPeek : public Scoring {
    using Scoring::Scoring;
    score_t wA(size_t k) const { return weightsA[k]; }
    score_t wB(size_t k) const { return weightsB[k]; }
};

// read access to AlignerP's protected inside table (aligner_p.hh:68ff)
struct AlignerPPeek : public AlignerP<double> {
    using AlignerP<double>::AlignerP;
    double Dval(const ArcMatch &x) const { return Dmat(x.arcA().idx(), x.arcB().idx()); }
};

struct Opts {
    double min_prob = 0.001;
    int max_diff_am = -1, max_diff_at_am = -1, max_diff = -1;
    double min_trace_probability = 1e-4;
    bool noLP = false, struct_local = false, sequ_local = false, use_ribosum = true;
    std::string free_endgaps = "----";
    int struct_weight = 200, indel = -150, indel_opening = -750, tau = 50, exclusion = 0;
    int match = 50, mismatch = 0, temperature_alipf = 300, unpaired_penalty = 0;
    bool do_trace = true, timing = false, pf_double = false;
    bool stacking = false, new_stacking = false;   // locarna --stacking / --new-stacking (locarna.cc:120-123)
    double exp_prob = -1.0;                        // --exp-prob; < 0: prob_exp_f(len) per sequence (locarna.cc:662-663)
    bool max_diff_relax = false;                   // --max-diff-relax: relaxed merging of the trace ranges
    std::string max_diff_pw_aln;                   // --max-diff-pw-aln "rowA&rowB" (locarna.cc:519-548): band around a reference alignment
    bool pf = false, pf_probs = false;  // LocARNA-P: inside partition function (+ D), outside + probabilities
    double pf_scale = 1.0, min_am_prob = 0.001, min_bm_prob = 0.001;
    std::string dump;  // comma list: arcs,band,am,D,aln,tables
    std::string pairs_file;
    std::vector<std::string> files;
    bool want(const char *k) const { return ("," + dump + ",").find(std::string(",") + k + ",") != std::string::npos; }
};

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static void print_score(std::ostream &o, infty_score_t s) {
    if (s.is_neg_infty()) o << "-inf"; else if (s.is_pos_infty()) o << "inf"; else o << s.finite_value();
}

template <class PF>
static void envelope(const Opts &o, const RnaData &A, const RnaData &B, const RibosumFreq *ribosum, TraceController &tc) {
    if (!(o.min_trace_probability > 0.0)) return;
    // same arguments as MainHelper::make_trace_probs (main_helper.icc:371-405)
    Alphabet<char, 4> alphabet;
    Matrix<double> bm;
    if (ribosum) { alphabet = ribosum->alphabet(); bm = ribosum->get_basematch_scores(); }
    else {
        alphabet = Alphabet<char, 4>("ACGU");
        bm.resize(4, 4); bm.fill(o.mismatch);
        for (size_t i = 0; i < 4; ++i) bm(i, i) = o.match;
    }
    PFTraceProbs<PF> tp(A, B, tc, bm, alphabet, o.indel_opening / 100.0, o.indel / 100.0, o.struct_weight / 100.0,
                        o.temperature_alipf / 100.0, FreeEndgaps(o.free_endgaps), o.sequ_local);
    if (tp.fail()) std::cerr << "WARNING: trace probabilities failed" << std::endl;
    tc.restrict_by_trace_probabilities(tp, o.min_trace_probability);
}

static std::map<std::string, std::shared_ptr<RnaData>> g_cache;

static std::shared_ptr<RnaData> load(const Opts &o, const std::string &f, const PFoldParams &pf) {
    auto it = g_cache.find(f);
    if (it != g_cache.end()) return it->second;
    auto r = std::make_shared<RnaData>(f, o.min_prob, 0.0, pf);
    g_cache[f] = r;
    return r;
}

static int run_pair(const Opts &o, const std::string &fA, const std::string &fB, const RibosumFreq *ribosum, long idx) {
    PFoldParams pfoldparams(PFoldParams::args::noLP(o.noLP), PFoldParams::args::stacking(o.stacking || o.new_stacking),
                            PFoldParams::args::max_bp_span(-1));
    double t0 = now_ms();
    std::shared_ptr<RnaData> rA, rB;
    try { rA = load(o, fA, pfoldparams); rB = load(o, fB, pfoldparams); }
    catch (failure &f) { std::cerr << "ERROR reading input: " << f.what() << std::endl; return -1; }
    const Sequence &seqA = rA->sequence();
    const Sequence &seqB = rB->sequence();
    size_t lenA = seqA.length(), lenB = seqB.length();
    double t1 = now_ms();

    AnchorConstraints constraints(lenA, seqA.annotation(MultipleAlignment::AnnoType::anchors).single_string(), lenB,
                                  seqB.annotation(MultipleAlignment::AnnoType::anchors).single_string(), true);
    std::unique_ptr<MultipleAlignment> ref_aln;
    if (!o.max_diff_pw_aln.empty()) {
        const size_t amp = o.max_diff_pw_aln.find('&');
        ref_aln = std::make_unique<MultipleAlignment>(seqA.seqentry(0).name(), seqB.seqentry(0).name(), o.max_diff_pw_aln.substr(0, amp),
                                                      o.max_diff_pw_aln.substr(amp + 1));
    }
    TraceController tc(seqA, seqB, ref_aln.get(), o.max_diff, o.max_diff_relax);
    tc.restrict_by_anchors(constraints);
    if (o.pf_double) envelope<double>(o, *rA, *rB, ribosum, tc);
    else envelope<long double>(o, *rA, *rB, ribosum, tc);  // locarna.cc:384-391 forces extended pf
    double t2 = now_ms();

    ArcMatches am(*rA, *rB, o.min_prob, o.max_diff_am != -1 ? (size_t)o.max_diff_am : std::max(lenA, lenB),
                  o.max_diff_at_am != -1 ? (size_t)o.max_diff_at_am : std::max(lenA, lenB), tc, constraints);
    double t3 = now_ms();

    auto sp = ScoringParams(ScoringParams::match(o.match), ScoringParams::mismatch(o.mismatch),
                            ScoringParams::indel(o.indel), ScoringParams::indel_opening(o.indel_opening),
                            ScoringParams::ribosum(ribosum), ScoringParams::ribofit(nullptr),
                            ScoringParams::unpaired_penalty(o.unpaired_penalty),
                            ScoringParams::struct_weight(o.struct_weight), ScoringParams::tau_factor(o.tau),
                            ScoringParams::exclusion(o.exclusion), ScoringParams::exp_probA(o.exp_prob >= 0 ? o.exp_prob : prob_exp_f(lenA)),
                            ScoringParams::exp_probB(o.exp_prob >= 0 ? o.exp_prob : prob_exp_f(lenB)),
                            ScoringParams::temperature_alipf(o.temperature_alipf), ScoringParams::stacking(o.stacking),
                            ScoringParams::new_stacking(o.new_stacking));
    ScoringPeek scoring(seqA, seqB, *rA, *rB, am, nullptr, sp);
    double t4 = now_ms();

    AlignerParams ap(AlignerParams::seqA(&seqA), AlignerParams::seqB(&seqB), AlignerParams::scoring(&scoring),
                     AlignerParams::no_lonely_pairs(o.noLP), AlignerParams::struct_local(o.struct_local),
                     AlignerParams::sequ_local(o.sequ_local), AlignerParams::free_endgaps(FreeEndgaps(o.free_endgaps)),
                     AlignerParams::max_diff_am(o.max_diff_am), AlignerParams::max_diff_at_am(o.max_diff_at_am),
                     AlignerParams::trace_controller(&tc), AlignerParams::stacking(o.stacking || o.new_stacking),
                     AlignerParams::constraints(&constraints));
    Aligner aligner(ap);
    // Aligner's only data member is its (private) pimpl pointer (aligner.hh:68); the harness reads D
    // and the raw alignment through it.
    AlignerImpl &impl = **reinterpret_cast<std::unique_ptr<AlignerImpl> *>(&aligner);
    infty_score_t score = aligner.align();
    double t5 = now_ms();
    if (o.do_trace) aligner.trace();
    double t6 = now_ms();

    std::ostream &out = std::cout;
    out << "PAIR " << idx << " " << fA << " " << fB << "\n";
    out << "LEN " << lenA << " " << lenB << "\n";
    out << "SEQA " << seqA.seqentry(0).seq().str() << "\n";
    out << "SEQB " << seqB.seqentry(0).seq().str() << "\n";
    const BasePairs &bA = am.get_base_pairsA(), &bB = am.get_base_pairsB();
    out << "NARCS " << bA.num_bps() << " " << bB.num_bps() << "\n";
    out << "NAM " << am.num_arc_matches() << "\n";
    if (o.want("arcs")) {
        for (int s = 0; s < 2; ++s) {
            const BasePairs &b = s ? bB : bA;
            const RnaData &r = s ? *rB : *rA;
            out << (s ? "ARCSB " : "ARCSA ") << b.num_bps() << "\n";
            for (size_t k = 0; k < b.num_bps(); ++k) {
                const auto &a = b.arc(k);
                char buf[64];
                snprintf(buf, sizeof buf, "%.17g", r.arc_prob(a.left(), a.right()));
                out << a.left() << " " << a.right() << " " << buf << " "
                    << (s ? scoring.wB(k) : scoring.wA(k)) << "\n";
            }
        }
    }
    if (o.want("band")) {
        out << "BANDMIN";
        for (size_t i = 0; i <= lenA; ++i) out << " " << tc.min_col(i);
        out << "\nBANDMAX";
        for (size_t i = 0; i <= lenA; ++i) out << " " << tc.max_col(i);
        out << "\n";
    }
    if (o.want("tables")) {
        out << "SIGMA";
        const char nt[] = "ACGU";
        (void)nt;
        for (size_t i = 1; i <= lenA; ++i)
            for (size_t j = 1; j <= lenB; ++j) out << " " << scoring.basematch(i, j);
        out << "\nGAPA";
        for (size_t i = 1; i <= lenA; ++i) out << " " << scoring.gapA(i);
        out << "\nGAPB";
        for (size_t j = 1; j <= lenB; ++j) out << " " << scoring.gapB(j);
        out << "\n";
    }
    if (o.want("am") || o.want("D")) {
        out << "AM " << am.num_arc_matches() << "\n";
        for (size_t k = 0; k < am.num_arc_matches(); ++k) {
            const ArcMatch &x = am.arcmatch(k);
            out << x.arcA().left() << " " << x.arcA().right() << " " << x.arcB().left() << " " << x.arcB().right()
                << " " << scoring.arcmatch(x) << " ";
            if (am.exists_inner_arc_match(x)) out << am.inner_arc_match(x).idx(); else out << -1;
            if (o.want("D")) { out << " "; print_score(out, impl.Dmat_(x.arcA().idx(), x.arcB().idx())); }
            out << "\n";
        }
    }
    out << "SCORE "; print_score(out, score); out << "\n";
    if (o.pf) {
        // same object pipeline as locarna_p.cc:441-526 (PFScoring<double>, AlignerP<double>)
        double tp0 = now_ms();
        PFScoring<double> pfs(seqA, seqB, *rA, *rB, am, nullptr, sp);
        using app_t = AlignerPParams<double>;
        AlignerPPeek ap2(app_t(AlignerParams::seqA(&seqA), AlignerParams::seqB(&seqB), AlignerParams::scoring(&pfs),
                               AlignerParams::max_diff_am(o.max_diff_am), AlignerParams::max_diff_at_am(o.max_diff_at_am),
                               AlignerParams::trace_controller(&tc), AlignerParams::constraints(&constraints),
                               app_t::min_am_prob(o.min_am_prob), app_t::min_bm_prob(o.min_bm_prob), app_t::pf_scale(o.pf_scale)));
        const double Z = ap2.align_inside();
        double tp1 = now_ms();
        char buf[64];
        snprintf(buf, sizeof buf, "%.17g", Z);
        out << "PF " << buf << "\n";
        if (o.want("D")) {
            out << "PFD " << am.num_arc_matches() << "\n";
            for (size_t k = 0; k < am.num_arc_matches(); ++k) { snprintf(buf, sizeof buf, "%.17g", ap2.Dval(am.arcmatch(k))); out << buf << "\n"; }
        }
        if (o.pf_probs) {
            ap2.align_outside();
            ap2.compute_arcmatch_probabilities();
            ap2.compute_basematch_probabilities(false);
            double tp2 = now_ms();
            std::ostringstream sa, sb;
            sa.precision(17); sb.precision(17);
            ap2.write_arcmatch_probabilities(sa);
            ap2.write_basematch_probabilities(sb);
            auto count = [](const std::string &t) { size_t n = 0; for (char ch : t) n += ch == '\n'; return n; };
            out << "AMPROBS " << count(sa.str()) << "\n" << sa.str();
            out << "BMPROBS " << count(sb.str()) << "\n" << sb.str();
            if (o.timing) { snprintf(buf, sizeof buf, "%.3f", tp2 - tp1); out << "TIMEPF_OUTSIDE " << buf << "\n"; }
        }
        if (o.timing) { snprintf(buf, sizeof buf, "%.3f", tp1 - tp0); out << "TIMEPF_INSIDE " << buf << "\n"; }
    }
    if (o.do_trace && o.want("aln")) {
        const Alignment &al = impl.alignment_;
        auto edges = al.alignment_edges(false);
        out << "EDGES " << edges.size() << "\n";
        for (const auto &e : edges) {
            // positions 1..len; gaps as -1 (regular), -2 (loop), -3 (locality), -4 (other)
            long a = e.first.is_pos() ? (long)(size_t)e.first : -1 - (long)(int)e.first.gap();
            long b = e.second.is_pos() ? (long)(size_t)e.second : -1 - (long)(int)e.second.gap();
            out << a << " " << b << "\n";
        }
        out << "STRA " << al.dot_bracket_structureA(false) << "\n";
        out << "STRB " << al.dot_bracket_structureB(false) << "\n";
        MultipleAlignment ma(al, false, false);
        out << "ROWA " << ma.seqentry(0).seq().str() << "\n";
        out << "ROWB " << ma.seqentry(1).seq().str() << "\n";
    }
    if (o.timing) {
        char buf[256];
        snprintf(buf, sizeof buf, "TIME read %.3f band %.3f am %.3f scoring %.3f align %.3f trace %.3f", t1 - t0, t2 - t1,
                 t3 - t2, t4 - t3, t5 - t4, t6 - t5);
        out << buf << "\n";
    }
    out << "END\n";
    return 0;
}

int main(int argc, char **argv) {
    Opts o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto nxt = [&]() -> const char * { if (i + 1 >= argc) { std::cerr << "missing value for " << a << "\n"; exit(2); } return argv[++i]; };
        if (a == "--min-prob") o.min_prob = atof(nxt());
        else if (a == "--max-diff-am") o.max_diff_am = atoi(nxt());
        else if (a == "--max-diff-at-am") o.max_diff_at_am = atoi(nxt());
        else if (a == "--max-diff") o.max_diff = atoi(nxt());
        else if (a == "--max-diff-pw-aln") o.max_diff_pw_aln = nxt();
        else if (a == "--max-diff-relax") o.max_diff_relax = true;
        else if (a == "--min-trace-probability") o.min_trace_probability = atof(nxt());
        else if (a == "--noLP") o.noLP = true;
        else if (a == "--struct-local") o.struct_local = true;
        else if (a == "--sequ-local") o.sequ_local = true;
        else if (a == "--free-endgaps") o.free_endgaps = nxt();
        else if (a == "--struct-weight") o.struct_weight = atoi(nxt());
        else if (a == "--indel") o.indel = atoi(nxt());
        else if (a == "--indel-opening") o.indel_opening = atoi(nxt());
        else if (a == "--tau") o.tau = atoi(nxt());
        else if (a == "--exclusion") o.exclusion = atoi(nxt());
        else if (a == "--match") o.match = atoi(nxt());
        else if (a == "--mismatch") o.mismatch = atoi(nxt());
        else if (a == "--unpaired-penalty") o.unpaired_penalty = atoi(nxt());
        else if (a == "--temperature-alipf") o.temperature_alipf = atoi(nxt());
        else if (a == "--no-ribosum") o.use_ribosum = false;
        else if (a == "--stacking") o.stacking = true;
        else if (a == "--new-stacking") o.new_stacking = true;
        else if (a == "--exp-prob") o.exp_prob = atof(nxt());
        else if (a == "--pf-double") o.pf_double = true;
        else if (a == "--no-trace") o.do_trace = false;
        else if (a == "--pf") o.pf = true;
        else if (a == "--pf-probs") { o.pf = true; o.pf_probs = true; }
        else if (a == "--pf-scale") o.pf_scale = atof(nxt());
        else if (a == "--min-am-prob") o.min_am_prob = atof(nxt());
        else if (a == "--min-bm-prob") o.min_bm_prob = atof(nxt());
        else if (a == "--time") o.timing = true;
        else if (a == "--dump") o.dump = nxt();
        else if (a == "--pairs") o.pairs_file = nxt();
        else if (a.size() > 1 && a[0] == '-' && a[1] == '-') { std::cerr << "unknown flag " << a << "\n"; return 2; }
        else o.files.push_back(a);
    }
    std::unique_ptr<RibosumFreq> ribosum;
    if (o.use_ribosum) ribosum = std::make_unique<Ribosum85_60>();
    if (o.want("ribosum")) {
        // dump the derived integer tables of the built-in matrix (scoring.cc:141-198, :369-438)
        const char nt[] = "ACGU";
        std::cout << "RIBOSUM_BM";
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
            char buf[64]; snprintf(buf, sizeof buf, " %.17g", ribosum->basematch_score(nt[a], nt[b])); std::cout << buf; }
        std::cout << "\nRIBOSUM_BM_CORRECTED";
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
            char buf[64]; snprintf(buf, sizeof buf, " %.17g", ribosum->basematch_score_corrected(nt[a], nt[b])); std::cout << buf; }
        std::cout << "\nRIBOSUM_AMLOG2";
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) for (int c = 0; c < 4; ++c) for (int d = 0; d < 4; ++d) {
            double v = log(ribosum->arcmatch_prob(nt[a], nt[b], nt[c], nt[d]) /
                           (ribosum->basepair_prob(nt[a], nt[b]) * ribosum->basepair_prob(nt[c], nt[d]))) / log(2);
            char buf[64]; snprintf(buf, sizeof buf, " %.17g", v); std::cout << buf; }
        std::cout << "\n";
    }
    std::vector<std::pair<std::string, std::string>> pairs;
    if (!o.pairs_file.empty()) {
        std::ifstream in(o.pairs_file);
        std::string a, b;
        while (in >> a >> b) pairs.emplace_back(a, b);
    } else if (o.files.size() == 2) pairs.emplace_back(o.files[0], o.files[1]);
    else if (!o.want("ribosum")) { std::cerr << "usage: ref_harness [flags] A.pp B.pp | --pairs LIST\n"; return 2; }
    long idx = 0;
    for (auto &p : pairs) {
        try { if (run_pair(o, p.first, p.second, ribosum.get(), idx++) != 0) return 1; }
        catch (failure &f) { std::cerr << "ERROR: " << f.what() << std::endl; return 1; }
    }
    return 0;
}
