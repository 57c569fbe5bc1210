// Host-side problem model (see host_model.h). Each function names the reference code whose observable
// behaviour it reproduces (paths relative to /root/reference/src/LocARNA).
#include "host_model.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <mutex>
#include <sstream>
#include <unordered_map>

#include "ribosum85_60_tables.h"

namespace lb200 {

static inline long round2score(double d) { return (long)((d < 0) ? (d - 0.5) : (d + 0.5)); }  // scoring.hh:384-387

// symbol codes: A C G U -> 0..3, N -> 4; any other symbol gets one of the codes 5..7, assigned per process in
// order of first appearance (the base match score of such symbols only depends on identity, scoring.cc:186-191)
static std::mutex g_symbol_mutex;
static char g_other_symbols[3] = {0, 0, 0};
static int symbol_code(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'U': return 3;
        case 'N': return LB_CODE_N;
        default: break;
    }
    std::lock_guard<std::mutex> lock(g_symbol_mutex);
    for (int k = 0; k < 3; k++) {
        if (g_other_symbols[k] == c) return 5 + k;
        if (g_other_symbols[k] == 0) { g_other_symbols[k] = c; return 5 + k; }
    }
    return -1;
}
static const uint8_t CODE_N = LB_CODE_N;

// ---------------------------------------------------------------------------------- input
// Line discipline of the reference reader: lines that are empty or start with white space are
// comments, trailing blanks are dropped (aux.cc:122-145).
static bool next_line(std::istream &in, std::string &line) {
    while (std::getline(in, line)) {
        if (line.empty() || isspace((unsigned char)line[0])) continue;
        size_t e = line.find_last_not_of(" \t\r");
        line.erase(e + 1);
        return true;
    }
    return false;
}

static bool begins(const std::string &s, const char *p) { return s.compare(0, strlen(p), p) == 0; }

// rna_data.cc:984-1103 (PP 2.0), multiple_alignment.cc:279-401 (sequence block), aux.cc:65-70
bool read_pp(const std::string &path, double p_bpcut, Sequence &out, std::string &err, int max_bp_span, double max_bps_length_ratio, bool stacking) {
    std::ifstream in(path.c_str());
    if (!in) { err = "cannot open " + path; return false; }
    std::string line;
    std::getline(in, line);
    if (!begins(line, "#PP 2")) { err = path + ": not in PP 2.0 format"; return false; }
    std::string name, seq;
    std::vector<std::string> anchor_rows;
    std::vector<std::string> row_names, rows;   // all rows of the sequence block, in order of first appearance (multiple_alignment.cc:279-401)
    while (next_line(in, line)) {
        if (line[0] == '#') {
            if (begins(line, "#END")) break;
            if (begins(line, "#A")) {   // anchor annotation "#A<k> <string>", possibly in several blocks (multiple_alignment.cc:324-346)
                std::istringstream ls(line.substr(2));
                int idx = 0; std::string astr;
                ls >> idx >> astr;
                if (idx < 1) { err = "Invalid index in anchor specification."; return false; }
                if ((size_t)idx > anchor_rows.size() + 1) { err = "Non-contiguous anchor specification."; return false; }
                if ((size_t)idx > anchor_rows.size()) anchor_rows.resize(idx);
                anchor_rows[idx - 1] += astr;
            }
            continue;
        }
        std::istringstream ls(line);
        std::string n, s;
        ls >> n >> s;
        size_t k = 0;
        while (k < row_names.size() && row_names[k] != n) k++;
        if (k == row_names.size()) { row_names.push_back(n); rows.push_back(std::string()); }
        rows[k] += s;
    }
    if (rows.empty()) { err = path + ": no sequence in the PP input"; return false; }
    for (const std::string &r : rows) if (r.size() != rows[0].size()) { err = "Rows of unequal length in the alignment of " + path; return false; }
    name = row_names[0]; seq = rows[0];
    if (!next_line(in, line) || line != "#SECTION BASEPAIRS") { err = path + ": Expected base pair section header."; return false; }
    std::vector<int> pi, pj;
    std::vector<double> pp, pp2;
    bool stack_keyword = false, any_p2 = false;
    double cut = p_bpcut;
    // #BPCUT may raise the cutoff at any point of the section; it applies to the lines after it (rna_data.cc:1047-1078)
    std::vector<double> cut_at;
    while (next_line(in, line)) {
        if (line[0] == '#') {
            if (begins(line, "#END")) break;
            if (begins(line, "#BPCUT")) {
                std::istringstream ls(line);
                std::string d; double p;
                ls >> d >> p;
                if (ls.fail()) { err = "Cannot parse line \"" + line + "\" in base pairs section."; return false; }
                cut = std::max(p, cut);
            } else if (begins(line, "#STACK")) stack_keyword = true;
            continue;
        }
        // "i j p [p2]" (the all-vs-all stage parses hundreds of files per job: no stream object per line)
        const char *cur = line.c_str();
        char *end = nullptr;
        const long i = strtol(cur, &end, 10);
        bool bad = end == cur;
        cur = end;
        const long j = strtol(cur, &end, 10);
        bad |= end == cur;
        cur = end;
        const double p = strtod(cur, &end);
        bad |= end == cur;
        cur = end;
        if (bad) { err = "Cannot parse line \"" + line + "\" in base pairs section."; return false; }
        if (!(1 <= i && i < j && j <= (long)seq.size())) { err = "Invalid indices in PP input line \"" + line + "\"."; return false; }
        if (p <= cut) continue;
        if (max_bp_span >= 0 && j - i + 1 > max_bp_span) continue;
        double p2 = 0.0;
        if (stacking) {   // joint probability of (i, j) and (i+1, j-1), kept above the cutoff (rna_data.cc:1083-1093)
            const double v = strtod(cur, &end);
            if (end != cur && v > cut) { p2 = v; any_p2 = true; }
        }
        pi.push_back((int)i); pj.push_back((int)j); pp.push_back(p); pp2.push_back(p2);
    }
    if (!stack_keyword && any_p2) { err = "Stacking probabilties found but stack keyword is missing."; return false; }   // rna_data.cc:1097-1100
    // the pairs were already filtered line by line; pass a cutoff that keeps them all
    // (profile input: the symbol codes are not used, and gap symbols must not take one of the three free codes)
    if (!make_sequence(name, rows.size() > 1 ? std::string(seq.size(), 'N') : seq, pi.data(), pj.data(), pp.data(), (int)pi.size(), -1.0, out, err, max_bp_span,
                       max_bps_length_ratio, pp2.data())) return false;
    if (rows.size() > 1) {            // profile input: keep all rows (normalised like the first one)
        for (std::string &r : rows) for (auto &ch : r) { ch = (char)toupper((unsigned char)ch); if (ch == 'T') ch = 'U'; }
        out.row_names = row_names; out.rows = rows; out.seq = rows[0];
        out.codes.assign(out.len + 1, CODE_N);   // never scored by symbol code: profile pairs use position-specific tables
    }
    out.cutoff = cut;                 // RnaData::arc_cutoff_prob(): the largest of the given cutoff and the file's #BPCUT lines
    out.has_stacking = stack_keyword && stacking;
    return set_anchors(out, anchor_rows, err);
}

// names per position from the anchor rows (AnchorConstraints::transform_input, anchor_constraints.cc:93-135, strict semantics)
bool set_anchors(Sequence &s, const std::vector<std::string> &rows, std::string &err) {
    s.anchor_names.clear(); s.anchor_rows.clear(); s.anchor_rank.assign(s.len + 1, 0);
    if (rows.empty()) return true;
    s.anchor_rows = rows;
    s.anchor_names.assign(s.len + 1, "");
    for (const std::string &x : rows) {
        if ((int)x.size() != s.len) { err = "Error during parsing of anchor constraints. Anchor specification strings must have exactly the same length as the corresponding sequences."; return false; }
        for (int i = 0; i < s.len; i++) s.anchor_names[i + 1].push_back(x[i]);
    }
    std::string last;
    int rank = 0;
    for (int i = 1; i <= s.len; i++) {
        std::string &x = s.anchor_names[i];
        bool dont_care = true;
        for (char c : x) if (!(c == ' ' || c == '.' || c == '-')) dont_care = false;
        if (dont_care) { x.clear(); continue; }
        if (x <= last) { err = "Error during parsing of constraints. Anchor names not in strict lexicographic order at name \"" + x + "\"."; return false; }
        last = x;
        if (++rank > 255) { err = "more than 255 anchors per sequence are not supported"; return false; }
        s.anchor_rank[i] = (uint8_t)rank;
    }
    if (rank == 0) s.anchor_names.clear();
    return true;
}

bool make_sequence(const std::string &name, const std::string &seq, const int *pi, const int *pj, const double *pp, int npairs,
                   double p_bpcut, Sequence &out, std::string &err, int max_bp_span, double max_bps_length_ratio, const double *pp2) {
    out = Sequence();
    out.name = name;
    out.seq = seq;
    for (auto &c : out.seq) { c = (char)toupper((unsigned char)c); if (c == 'T') c = 'U'; }
    out.len = (int)out.seq.size();
    if (out.len > LB_MAXLEN) { err = "sequence longer than 4095 positions"; return false; }
    out.codes.assign(out.len + 1, 0);
    for (int i = 1; i <= out.len; i++) {
        const int code = symbol_code(out.seq[i - 1]);
        if (code < 0) { err = "more than three distinct sequence symbols besides A C G U N are not supported"; return false; }
        out.codes[i] = (uint8_t)code;
    }
    out.cutoff = p_bpcut;
    // a repeated pair overwrites the earlier value (sparse matrix assignment): stable sort by (i, j), keep the last of every run
    struct Rec { int i, j; double p, p2; };
    std::vector<Rec> uniq;
    {
        std::vector<Rec> all;
        all.reserve((size_t)npairs);
        for (int k = 0; k < npairs; k++) {
            if (!(1 <= pi[k] && pi[k] < pj[k] && pj[k] <= out.len)) { err = "invalid base pair indices"; return false; }
            if (pp[k] <= p_bpcut) continue;
            if (max_bp_span >= 0 && pj[k] - pi[k] + 1 > max_bp_span) continue;  // rna_data.cc:1078, bp_span = j-i+1 (aux.hh:333)
            all.push_back(Rec{pi[k], pj[k], pp[k], (pp2 != nullptr && pp2[k] > 0) ? pp2[k] : -1.0});
        }
        std::stable_sort(all.begin(), all.end(), [](const Rec &x, const Rec &y) { return x.i != y.i ? x.i < y.i : x.j < y.j; });
        for (size_t k = 0; k < all.size(); k++) {
            if (!uniq.empty() && uniq.back().i == all[k].i && uniq.back().j == all[k].j) {
                uniq.back().p = all[k].p;
                if (all[k].p2 >= 0) uniq.back().p2 = all[k].p2;     // the joint value of an earlier line stays unless this line gives one
            } else uniq.push_back(all[k]);
        }
    }
    if (max_bps_length_ratio > 0.0) {
        // rna_data.cc:64-67, :1580-1601 (drop_worst_bps): only the `keep` most probable base pairs survive. The reference pops a
        // min-heap filled in hash-table order, so WHICH of several equally probable pairs at the cut goes is not defined by its
        // source; such an input is refused instead of guessed.
        const size_t keep = (size_t)(max_bps_length_ratio * out.len);
        if (uniq.size() > keep) {
            std::vector<double> ps;
            for (auto &kv : uniq) ps.push_back(kv.p);
            std::sort(ps.begin(), ps.end(), [](double x, double y) { return x > y; });
            if (keep > 0 && ps[keep - 1] == ps[keep]) { err = "--max-bps-length-ratio: equally probable base pairs at the cut (the reference's choice among them is unspecified)"; return false; }
            const double thr = keep > 0 ? ps[keep - 1] : 2.0;
            std::vector<Rec> kept;
            for (auto &kv : uniq) if (!(kv.p < thr)) kept.push_back(kv);
            uniq.swap(kept);
        }
    }
    out.pp_i.reserve(uniq.size()); out.pp_j.reserve(uniq.size()); out.pp_p.reserve(uniq.size()); out.pp_p2.reserve(uniq.size());
    for (auto &kv : uniq) {
        out.pp_i.push_back(kv.i); out.pp_j.push_back(kv.j); out.pp_p.push_back(kv.p);
        out.pp_p2.push_back(kv.p2 > 0 ? kv.p2 : 0.0);
    }
    return true;
}

// basepairs.cc:155-205: arcs (i,j), j >= i+3, p >= min_prob, indexed in loop order i = len-3..1, j = i+3..len;
// rna_data.cc:713-732: paired upstream/downstream mass, summed over ascending partner position
void finish_sequence(Sequence &s, double min_prob) {
    const int n = s.len;
    std::vector<int> order;
    for (size_t k = 0; k < s.pp_i.size(); k++)
        if (s.pp_j[k] >= s.pp_i[k] + 3 && s.pp_p[k] >= min_prob) order.push_back((int)k);
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        if (s.pp_i[x] != s.pp_i[y]) return s.pp_i[x] > s.pp_i[y];
        return s.pp_j[x] < s.pp_j[y];
    });
    s.arcs.clear(); s.arc_prob.clear(); s.arc_joint.clear(); s.arc_inner.clear();
    s.lptr.assign(n + 2, 0); s.lcount.assign(n + 2, 0);
    // RnaData::arc_prob of every pair the reader kept: pp_* is sorted by (i, j) ascending, so a binary search finds a pair
    auto arc_prob_of = [&](int i, int j) -> double {
        size_t lo = 0, hi = s.pp_i.size();
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (s.pp_i[mid] < i || (s.pp_i[mid] == i && s.pp_j[mid] < j)) lo = mid + 1; else hi = mid;
        }
        return (lo < s.pp_i.size() && s.pp_i[lo] == i && s.pp_j[lo] == j) ? s.pp_p[lo] : 0.0;
    };
    for (int k : order) {
        int l = s.pp_i[k];
        if (s.lcount[l] == 0) s.lptr[l] = (int)s.arcs.size();
        s.lcount[l]++;
        s.arcs.push_back(Arc{s.pp_i[k], s.pp_j[k]});
        s.arc_prob.push_back(s.pp_p[k]);
        s.arc_joint.push_back(k < (int)s.pp_p2.size() ? s.pp_p2[k] : 0.0);
        s.arc_inner.push_back(arc_prob_of(s.pp_i[k] + 1, s.pp_j[k] - 1));
    }
    // pp_* is sorted by (i, j) ascending (std::map order)
    s.p_up.assign(n + 1, 0.0); s.p_down.assign(n + 1, 0.0);
    for (size_t k = 0; k < s.pp_i.size(); k++) s.p_up[s.pp_i[k]] += s.pp_p[k];  // fixed i: j ascending
    {
        std::vector<int> idx(s.pp_i.size());
        for (size_t k = 0; k < idx.size(); k++) idx[k] = (int)k;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) {
            if (s.pp_j[x] != s.pp_j[y]) return s.pp_j[x] < s.pp_j[y];
            return s.pp_i[x] < s.pp_i[y];
        });
        for (int k : idx) s.p_down[s.pp_j[k]] += s.pp_p[k];  // fixed right end: left partner ascending
    }
}

// scoring.cc:201-265 (probToWeight with p_exp = 1/(2 len), aux.hh:216-220)
std::vector<int> arc_weights(const Sequence &s, const Params &p) {
    std::vector<int> w(s.arcs.size());
    const double pe = p.exp_prob >= 0 ? p.exp_prob : 1.0 / (2.0 * s.len);  // locarna.cc:662-663
    for (size_t k = 0; k < w.size(); k++) w[k] = (int)round2score(round(p.struct_weight * (1 - log(s.arc_prob[k]) / log(pe))));
    return w;
}

// scoring.cc:201-248 (stack_weights) relative to the weight, scoring.cc:556-564 (is_stackable_arc: joint probability > 0)
std::vector<int> arc_stack_deltas(const Sequence &s, const Params &p) {
    std::vector<int> d(s.arcs.size(), LB_NOSTACK);
    const double pe = p.exp_prob >= 0 ? p.exp_prob : 1.0 / (2.0 * s.len);
    auto weight = [&](double prob) { return round2score(round(p.struct_weight * (1 - log(prob) / log(pe)))); };
    for (size_t k = 0; k < d.size(); k++) {
        if (!(s.arc_joint[k] > 0)) continue;   // not stackable: the stacked score is never used
        const long w = weight(s.arc_prob[k]);
        long sw = 0;
        if (p.stacking && s.arc_inner[k] > 0) sw = weight(s.arc_joint[k] / s.arc_inner[k]);   // RnaData::stacked_arc_prob (rna_data.cc:705-710)
        if (p.new_stacking) {
            if (!p.stacking) sw = w;
            if (s.arc_inner[k] > 0) sw += weight(s.arc_joint[k]);
        }
        d[k] = (int)(sw - w);
    }
    return d;
}

// scoring.cc:141-198 (sigma_ for single sequences), :64-74 (unpaired penalty), :369-485 (arc match sequence term)
void make_score_tables(const Params &p, ScoreTables &t) {
    memset(&t, 0, sizeof t);
    for (int a = 0; a < LB_NCODES; a++)
        for (int b = 0; b < LB_NCODES; b++) {
            long s;
            if (a < 4 && b < 4) s = p.use_ribosum ? p.ribosum.sigma4[a * 4 + b] : (a == b ? p.match : p.mismatch);
            else if (a == CODE_N || b == CODE_N) s = 0;
            else s = (a == b) ? p.match : p.mismatch;
            t.dev.sigma8[a * LB_NCODES + b] = (int)s - 2 * p.unpaired_penalty;
        }
    for (int x = 0; x < 16; x++)
        for (int y = 0; y < 16; y++) t.am_seq[x * 16 + y] = (int)(((long)p.tau * p.ribosum.am16[x * 16 + y]) / 100);
    DevParams &d = t.dev;
    d.gap = p.indel - p.unpaired_penalty;
    d.open = p.indel_opening;
    d.gap_open = d.gap + d.open;
    d.exclusion = p.exclusion;
    d.no_lonely_pairs = p.no_lonely_pairs; d.struct_local = p.struct_local; d.sequ_local = p.sequ_local;
    d.stacking = p.stacking || p.new_stacking;
    d.fe_left1 = p.fe_left1; d.fe_right1 = p.fe_right1; d.fe_left2 = p.fe_left2; d.fe_right2 = p.fe_right2;
}

int base_match_score(const ScoreTables &t, uint8_t a, uint8_t b) { return t.dev.sigma8[a * LB_NCODES + b]; }

int arcmatch_score(const ScoreTables &t, const Params &p, const Sequence &A, const Sequence &B, int a, int b, const std::vector<int> &wA,
                   const std::vector<int> &wB) {
    const Arc &x = A.arcs[a], &y = B.arcs[b];
    long seqc = 0;
    if (p.tau != 0) {
        const uint8_t c1 = A.codes[x.left], c2 = A.codes[x.right], c3 = B.codes[y.left], c4 = B.codes[y.right];
        if (p.use_ribosum) {
            if (c1 < 4 && c2 < 4 && c3 < 4 && c4 < 4) return t.am_seq[(c1 * 4 + c2) * 16 + c3 * 4 + c4] + wA[a] + wB[b];
            seqc = 0;
        } else seqc = (long)base_match_score(t, c1, c3) + base_match_score(t, c2, c4);
    }
    return (int)(((long)p.tau * seqc) / 100) + wA[a] + wB[b];
}

// ---------------------------------------------------------------------------------- profile (multi-row) scoring
static inline int acgu_index(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'U' ? 3 : -1; }   // the ribosum alphabet

void make_profile_tables(const Sequence &A, const Sequence &B, const Params &p, ProfileTables &out) {
    const int n = A.len, m = B.len, ra = A.num_rows(), rb = B.num_rows();
    out = ProfileTables();
    out.n = n; out.m = m; out.arcsA = (int)A.arcs.size(); out.arcsB = (int)B.arcs.size();
    // Scoring::sigma_ (scoring.cc:141-198): sum over all row pairs of the rounded per-pair score, integer division by the number of
    // row pairs ("-" against "-" counts as a match, N never counts); unpaired penalty as apply_unpaired_penalty (:64-74)
    out.sigma.assign((size_t)(n + 1) * (m + 1), 0);
    std::vector<int8_t> ia((size_t)ra * (n + 1)), ib((size_t)rb * (m + 1));
    for (int k = 0; k < ra; k++) for (int i = 1; i <= n; i++) ia[(size_t)k * (n + 1) + i] = (int8_t)acgu_index(A.row(k)[i - 1]);
    for (int l = 0; l < rb; l++) for (int j = 1; j <= m; j++) ib[(size_t)l * (m + 1) + j] = (int8_t)acgu_index(B.row(l)[j - 1]);
    for (int i = 1; i <= n; i++)
        for (int j = 1; j <= m; j++) {
            long score = 0;
            for (int k = 0; k < ra; k++) {
                const char ca = A.row(k)[i - 1];
                const int xa = ia[(size_t)k * (n + 1) + i];
                for (int l = 0; l < rb; l++) {
                    const char cb = B.row(l)[j - 1];
                    const int xb = ib[(size_t)l * (m + 1) + j];
                    if (p.use_ribosum && xa >= 0 && xb >= 0) score += p.ribosum.sigma4[xa * 4 + xb];
                    else if (ca != 'N' && cb != 'N') score += (ca == cb) ? p.match : p.mismatch;
                }
            }
            out.sigma[(size_t)i * (m + 1) + j] = (int)round2score((double)(score / (long)(ra * rb))) - 2 * p.unpaired_penalty;
        }
    // Scoring::precompute_gapcost (scoring.cc:272-311): gap frequencies in float, as there
    auto gaps = [&](const Sequence &S, int len, int rows, std::vector<int> &g, std::vector<long> &P) {
        g.assign(len + 1, 0); P.assign(len + 1, 0);
        for (int i = 1; i <= len; i++) {
            float f = 0;
            for (int k = 0; k < rows; k++) f += (S.row(k)[i - 1] == '-') ? 1 : 0;
            f /= rows;
            g[i] = (int)round2score((1 - f) * p.indel) - p.unpaired_penalty;
            P[i] = P[i - 1] + g[i];
        }
    };
    gaps(A, n, ra, out.gapA, out.PA);
    gaps(B, m, rb, out.gapB, out.PB);
    // sequence contribution of Scoring::arcmatch (scoring.cc:441-485): riboX_arcmatch_score (:369-438) with a ribosum matrix,
    // sigma of the two ends without one
    out.am_seq.assign((size_t)out.arcsA * out.arcsB, 0);
    if (p.tau != 0) {
        for (int a = 0; a < out.arcsA; a++) {
            const int al = A.arcs[a].left, ar = A.arcs[a].right;
            for (int b = 0; b < out.arcsB; b++) {
                const int bl = B.arcs[b].left, br = B.arcs[b].right;
                long seqc;
                if (p.use_ribosum) {
                    double score = 0;
                    int considered = 0;
                    for (int k = 0; k < ra; k++) {
                        const char a1 = A.row(k)[al - 1], a2 = A.row(k)[ar - 1];
                        if (a1 == '-' || a2 == '-') continue;
                        const int x1 = ia[(size_t)k * (n + 1) + al], x2 = ia[(size_t)k * (n + 1) + ar];
                        for (int l = 0; l < rb; l++) {
                            const char b1 = B.row(l)[bl - 1], b2 = B.row(l)[br - 1];
                            if (b1 == '-' || b2 == '-') continue;
                            const int y1 = ib[(size_t)l * (m + 1) + bl], y2 = ib[(size_t)l * (m + 1) + br];
                            if (x1 < 0 || x2 < 0 || y1 < 0 || y2 < 0) continue;
                            considered++;
                            score += p.ribosum.amlog2[(x1 * 4 + x2) * 16 + y1 * 4 + y2];
                        }
                    }
                    seqc = considered == 0 ? 0 : round2score(100.0 * score / considered);
                } else {
                    // sigma_tab of both ends; the tables carry the unpaired penalty like the reference's
                    seqc = (long)out.sigma[(size_t)al * (m + 1) + bl] + out.sigma[(size_t)ar * (m + 1) + br];
                }
                out.am_seq[(size_t)a * out.arcsB + b] = (int)(((long)p.tau * seqc) / 100);
            }
        }
    }
}

// ---------------------------------------------------------------------------------- ribosum
RibosumTables::RibosumTables() {
    memcpy(bm, RIBOSUM85_60_BM, sizeof bm); memcpy(sigma4, RIBOSUM85_60_SIGMA4, sizeof sigma4); memcpy(am16, RIBOSUM85_60_AM16, sizeof am16);
    memcpy(amlog2, RIBOSUM85_60_AMLOG2, sizeof amlog2);
}

namespace {
long round2score_d(double d) { return (long)(d < 0 ? d - 0.5 : d + 0.5); }   // aux.hh round2score

// lower-triangular score matrix with a header line of names and a name in front of every row (Ribosum::read_matrix, ribosum.hh:383-424)
bool read_tri_matrix(std::istream &in, const std::vector<std::string> &names, std::vector<double> &mat, std::string &err) {
    const size_t n = names.size();
    std::string line;
    while (std::getline(in, line) && line == "") {}
    {
        std::istringstream ls(line);
        for (size_t i = 0; i < n; i++) { std::string nm; ls >> nm; if (nm != names[i]) { err = "Expecting correct table header. Found: " + line; return false; } }
    }
    mat.assign(n * n, 0.0);
    for (size_t i = 0; i < n; i++) {
        if (!std::getline(in, line)) { err = "unexpected end of the matrix"; return false; }
        std::istringstream ls(line);
        std::string base;
        ls >> base;
        if (base != names[i]) { err = "Expecting base name " + names[i] + " as row header"; return false; }
        for (size_t j = 0; j <= i; j++) { double v; if (!(ls >> v)) { err = "missing matrix entry"; return false; } mat[i * n + j] = mat[j * n + i] = v; }
    }
    return true;
}
// "HEADER" line followed by xdim * ydim numbers (RibosumFreq::read_matrix, ribosum.cc:169-197)
bool read_freq_matrix(std::istream &in, const std::string &header, size_t count, std::vector<double> &mat, std::string &err) {
    std::string line;
    auto blank = [](const std::string &s) { for (char c : s) if (!isspace((unsigned char)c)) return false; return true; };
    while (std::getline(in, line) && blank(line)) {}
    if (line != header) { err = "Expected header " + header + ". Read instead '" + line + "'."; return false; }
    mat.assign(count, 0.0);
    for (size_t k = 0; k < count; k++) if (!(in >> mat[k])) { err = "missing entries under " + header; return false; }
    return true;
}
}  // namespace

bool read_ribosum_file(const std::string &path, RibosumTables &out, std::string &err) {
    std::ifstream in(path.c_str());
    if (!in) { err = "Cannot open file " + path + " for reading ribosum data."; return false; }
    std::string line, e;
    if (!std::getline(in, line)) { err = "Cannot parse ribosum input. Expecting name."; return false; }
    const std::vector<std::string> bases = {"A", "C", "G", "U"};
    std::vector<std::string> arcs;
    for (const auto &x : bases) for (const auto &y : bases) arcs.push_back(x + y);
    std::vector<double> bm, am, f_base, f_nonstruct, f_pair, f_match, f_arcmatch;
    bool ok = read_tri_matrix(in, bases, bm, e);
    if (ok) { std::getline(in, line); std::getline(in, line); ok = read_tri_matrix(in, arcs, am, e); }   // "ignore two lines" (ribosum.cc:59-62)
    if (!ok) { err = "Cannot parse ribosum input. " + e + ": iostream error\nFile not in ribosum format."; return false; }   // text of the reference incl. ifstream::failure::what()
    std::getline(in, line); std::getline(in, line);
    ok = read_freq_matrix(in, "BASE FREQUENCIES", 4, f_base, e) && read_freq_matrix(in, "BASE NONSTRUCTURAL FREQUENCIES", 4, f_nonstruct, e) &&
         read_freq_matrix(in, "BASE PAIR FREQUENCIES", 16, f_pair, e) && read_freq_matrix(in, "BASE MATCH FREQUENCIES", 16, f_match, e) &&
         read_freq_matrix(in, "BASE PAIR MATCH FREQUENCIES", 256, f_arcmatch, e);
    if (!ok) { err = "Cannot parse ribosum frequency input. " + e + ": iostream error\nFile not in extended ribosum format."; return false; }
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            out.bm[a * 4 + b] = bm[a * 4 + b];
            // basematch_score_corrected (ribosum.cc:324-331), scoring.cc:174-177
            out.sigma4[a * 4 + b] = (int)round2score_d(100.0 * (log(f_match[a * 4 + b] / (f_nonstruct[a] * f_nonstruct[b])) / log(2)));
        }
    for (int x = 0; x < 16; x++)
        for (int y = 0; y < 16; y++) {   // scoring.cc:411-427 for one row per sequence
            out.amlog2[x * 16 + y] = log(f_arcmatch[x * 16 + y] / (f_pair[x] * f_pair[y])) / log(2);
            out.am16[x * 16 + y] = (int)round2score_d(100.0 * (log(f_arcmatch[x * 16 + y] / (f_pair[x] * f_pair[y])) / log(2)) / 1);
        }
    return true;
}

// ---------------------------------------------------------------------------------- band
// trace_controller.cc:406-424 and constrain_wo_ref :360-390 (unsigned 64-bit arithmetic as in the reference)
Band make_band(int lenA, int lenB, int max_diff) {
    Band b;
    b.lenA = lenA; b.lenB = lenB;
    b.lo.assign(lenA + 1, 0); b.hi.assign(lenA + 1, lenB);
    if (max_diff == -1 || lenA == 0 || lenB == 0) return b;
    const uint64_t A = lenA, B = lenB, d = (uint64_t)max_diff;
    for (uint64_t i = 0; i <= A; i++) {
        uint64_t x = i * B * (A + B), y = 2 * d * A * B, z = A * (A + B);
        if (A > B) y = std::max(y, (A + B) * A / 2);
        else if (B > A) y = std::max(y, (A + B) * B / 2);
        b.lo[i] = x > y ? (int)((x - y + z - 1) / z) : 0;
        b.hi[i] = (int)std::min((x + y) / z, B);
    }
    return b;
}

int restrict_band_by_anchors(Band &band, const Sequence &A, const Sequence &B, std::string &err) {
    if (A.anchor_names.empty() || B.anchor_names.empty()) return 0;   // one anchor spec empty: no anchors at all (anchor_constraints.cc:27-31)
    const int n = A.len, m = B.len;
    std::vector<int> ia, jb;   // positions of the names, in order
    for (int i = 1; i <= n; i++) if (A.anchor_rank[i]) ia.push_back(i);
    for (int j = 1; j <= m; j++) if (B.anchor_rank[j]) jb.push_back(j);
    bool same = ia.size() == jb.size();
    for (size_t k = 0; same && k < ia.size(); k++) same = A.anchor_names[ia[k]] == B.anchor_names[jb[k]];
    if (!same) { err = "anchor names that occur in only one of the two sequences are not supported"; return -1; }
    // allowed match / deletion / insertion ranges for names that all occur in both sequences (init_tables, anchor_constraints.cc:164-330,
    // strict): entries 0 keep the constructor's defaults (1, len)
    typedef std::pair<int, int> range;
    std::vector<range> ar(n + 1, range(1, m)), adr(n + 1, range(1, m)), air(m + 1, range(1, n));
    {
        int last = 0;
        for (int i = 1; i <= n; i++) { if (A.anchor_rank[i]) { last = jb[A.anchor_rank[i] - 1]; ar[i].first = last; } else ar[i].first = last + 1; }
        last = m + 1;
        for (int i = n; i >= 1; i--) { if (A.anchor_rank[i]) { last = jb[A.anchor_rank[i] - 1]; ar[i].second = last; } else ar[i].second = last - 1; }
        int prev = 0;   // partner of the last name at or before i
        std::vector<int> nextj(n + 2, m + 1);
        for (int i = n; i >= 1; i--) nextj[i] = A.anchor_rank[i] ? jb[A.anchor_rank[i] - 1] : nextj[i + 1];
        for (int i = 1; i <= n; i++) {
            if (A.anchor_rank[i]) { prev = jb[A.anchor_rank[i] - 1]; adr[i] = range(m + 1, 0); }
            else adr[i] = range(prev, nextj[i] - 1);
        }
        prev = 0;
        std::vector<int> nexti(m + 2, n + 1);
        for (int j = m; j >= 1; j--) nexti[j] = B.anchor_rank[j] ? ia[B.anchor_rank[j] - 1] : nexti[j + 1];
        for (int j = 1; j <= m; j++) {
            if (B.anchor_rank[j]) { prev = ia[B.anchor_rank[j] - 1]; air[j] = range(n + 1, 0); }
            else air[j] = range(prev, nexti[j] - 1);
        }
    }
    auto allowed_match = [&](int i, int j) { return ar[i].first <= j && j <= ar[i].second; };
    auto allowed_del = [&](int i, int j) { return adr[i].first <= j && j <= adr[i].second; };
    auto allowed_ins = [&](int i, int j) { return air[j].first <= i && i <= air[j].second; };
    for (int i = 1; i <= n; i++) {   // TraceController::restrict_by_anchors (trace_controller.cc:541-563); row 0 is not touched
        int cmin = band.hi[i], cmax = band.lo[i];
        for (int j = band.lo[i]; j <= band.hi[i]; j++)
            if (allowed_match(i, j) || (j > 0 && allowed_ins(i, j)) || (i > 0 && allowed_del(i, j))) { cmin = std::min(cmin, j); cmax = std::max(cmax, j); }
        band.lo[i] = std::max(band.lo[i], cmin);
        band.hi[i] = std::min(band.hi[i], cmax);
    }
    // The reference leaves row 0 as it was, i.e. possibly wider than row 1. Border cells (0, j) beyond row 1's last column are read by no
    // cell of the band and carry no trace probability, so cutting row 0 back changes no score, alignment or envelope - and keeps the
    // rows monotone, which the kernels rely on. (Without the envelope the reference's own row 0 stays wide: the only visible difference.)
    if (n >= 1) band.hi[0] = std::min(band.hi[0], band.hi[1]);
    return 1;
}

// Band around a pairwise reference alignment (--max-diff-aln / --max-diff-pw-aln with --max-diff delta): TraceRange(pseqA, pseqB,
// paliA, paliB, delta), trace_controller.cc:44-215 (position cut distance), for ungapped sequences A and B (their positions are their
// columns), intersected with the unconstrained range as merge_in_trace_range does (:606-622). aliA / aliB: the two rows of the
// reference alignment, gap symbols "-_~." (aux.cc:24). Returns false (err set) if the rows do not spell out sequences of these
// lengths or the range is inconsistent.
bool band_from_alignment(int lenA, int lenB, const std::string &aliA_in, const std::string &aliB_in, int delta_in, Band &b, std::string &err, bool relaxed) {
    const int delta = relaxed ? 0 : delta_in;   // relaxed merging: the trace of the reference alignment itself, widened afterwards (trace_controller.cc:462-468)
    auto gap = [](char c) { return c == '-' || c == '_' || c == '~' || c == '.'; };
    if (aliA_in.size() != aliB_in.size()) { err = "reference alignment rows have unequal lengths"; return false; }
    std::string aliA, aliB;   // remove_common_gaps (:25-42)
    for (size_t k = 0; k < aliA_in.size(); k++)
        if (!(gap(aliA_in[k]) && gap(aliB_in[k]))) { aliA += aliA_in[k]; aliB += aliB_in[k]; }
    size_t na = 0, nb = 0;
    for (char c : aliA) na += gap(c) ? 0 : 1;
    for (char c : aliB) nb += gap(c) ? 0 : 1;
    if ((int)na != lenA || (int)nb != lenB) { err = "reference alignment does not match the sequence lengths"; return false; }
    b.lenA = lenA; b.lenB = lenB;
    b.lo.assign(lenA + 1, lenB); b.hi.assign(lenA + 1, 0);
    const size_t d = (size_t)std::max(delta, 0), A = (size_t)lenA, B = (size_t)lenB;
    size_t i = 0, j = 0;
    for (size_t c = 0; c <= aliA.size(); c++) {   // cut of the reference alignment after column c
        if (c > 0) { i += gap(aliA[c - 1]) ? 0 : 1; j += gap(aliB[c - 1]) ? 0 : 1; }
        const size_t i_minus = std::max(d, i) - d, i_plus = std::min(A, i + d), j_minus = std::max(d, j) - d, j_plus = std::min(B, j + d);
        b.lo[i] = std::min(b.lo[i], (int)j_minus); b.hi[i] = std::max(b.hi[i], (int)j_plus);
        for (size_t pi = i_minus; pi < i; pi++) b.hi[pi] = std::max(b.hi[pi], (int)j);
        for (size_t pi = i + 1; pi <= i_plus; pi++) b.lo[pi] = std::min(b.lo[pi], (int)j);
    }
    if (relaxed) {
        // --max-diff-relax (trace_controller.cc:485-511): consensus trace through the trace ranges (one range here: single sequences),
        // TraceRange(lenA, lenB, trs, delta) :246-311 with consensus_cost :214-243 (its second branch takes the minimum with the running
        // value, as written there), then the range is blown up by delta in both directions
        const Band tr0 = b;
        const int n = lenA, m = lenB;
        auto cost = [&](long i, long j) -> size_t {
            size_t dprime = std::numeric_limits<size_t>::max();
            for (long i2 = 0; i2 <= n; i2++) {
                size_t d2;
                if (j < tr0.lo[i2]) d2 = (size_t)(labs(i - i2) + (tr0.lo[i2] - j));
                else if (j > tr0.hi[i2]) d2 = std::min(dprime, (size_t)(labs(i - i2) + (j - tr0.hi[i2])));
                else d2 = (size_t)labs(i - i2);
                dprime = std::min(dprime, d2);
            }
            return dprime;
        };
        std::vector<size_t> C((size_t)(n + 1) * (m + 1));
        std::vector<uint8_t> T((size_t)(n + 1) * (m + 1));
        auto at = [&](int i, int j) { return (size_t)i * (m + 1) + j; };
        T[at(0, 0)] = 3; C[at(0, 0)] = cost(0, 0);
        for (int i = 1; i <= n; i++) { T[at(i, 0)] = 1; C[at(i, 0)] = cost(i, 0) + C[at(i - 1, 0)]; }
        for (int j = 1; j <= m; j++) { T[at(0, j)] = 2; C[at(0, j)] = cost(0, j) + C[at(0, j - 1)]; }
        for (int i = 1; i <= n; i++)
            for (int j = 1; j <= m; j++) {
                size_t c = cost(i, j);
                if (C[at(i - 1, j - 1)] < C[at(i - 1, j)] && C[at(i - 1, j - 1)] < C[at(i, j - 1)]) { T[at(i, j)] = 0; c += C[at(i - 1, j - 1)]; }
                else if (C[at(i - 1, j)] < C[at(i, j - 1)]) { T[at(i, j)] = 1; c += C[at(i - 1, j)]; }
                else { T[at(i, j)] = 2; c += C[at(i, j - 1)]; }
                C[at(i, j)] = c;
            }
        Band tr;
        tr.lo.assign(n + 1, m); tr.hi.assign(n + 1, 0);
        for (int i = n, j = m;;) {
            tr.lo[i] = std::min(tr.lo[i], j); tr.hi[i] = std::max(tr.hi[i], j);
            if (T[at(i, j)] == 3) break;
            if (T[at(i, j)] == 0) { i--; j--; } else if (T[at(i, j)] == 1) i--; else j--;
        }
        const int dl = std::max(delta_in, 0);
        for (int i = 0; i <= n; i++) {
            int lo = std::min(std::max(tr.lo[i], dl) - dl, m), hi = std::max(std::min(tr.hi[i] + dl, m), 0);
            const int i_minus = std::max(dl, i) - dl, i_plus = std::min(i + dl, n);
            lo = std::min(tr.lo[i_minus], lo); hi = std::max(tr.hi[i_plus], hi);
            b.lo[i] = lo; b.hi[i] = hi;
        }
        return true;
    }
    for (int r = 0; r <= lenA; r++)
        if (b.lo[r] > b.hi[r] || (r > 0 && b.hi[r - 1] + 1 < b.lo[r])) { err = "Inconsistent trace range due to max-diff heuristic"; return false; }
    return true;
}

namespace {

// STRAL-like position score of the sequence-only partition function (stral_score.cc:29-60);
// sim = ribosum base match scores (main_helper.icc:291-293) or match/mismatch (:300-306)
struct EnvScore {
    const Sequence *A, *B;
    bool rev = false;
    double sw, match, mismatch;
    bool ribo;
    const double *ribo_bm;
    double up(const Sequence &s, int i) const { return rev ? s.p_down[s.len + 1 - i] : s.p_up[i]; }
    double down(const Sequence &s, int i) const { return rev ? s.p_up[s.len + 1 - i] : s.p_down[i]; }
    uint8_t code(const Sequence &s, int i) const { return s.codes[rev ? s.len + 1 - i : i]; }
    double sigma(int i, int j) const {
        double seq_score = 0;
        if (A->num_rows() > 1 || B->num_rows() > 1) {   // alignment columns: average over the row pairs of alphabet symbols (stral_score.cc:29-44)
            const int pi = rev ? A->len + 1 - i : i, pj = rev ? B->len + 1 - j : j;
            int pairs = 0;
            for (int k = 0; k < A->num_rows(); k++) {
                const int a = acgu_index(A->row(k)[pi - 1]);
                for (int l = 0; l < B->num_rows(); l++) {
                    const int b = acgu_index(B->row(l)[pj - 1]);
                    if (a >= 0 && b >= 0) { seq_score += ribo ? ribo_bm[a * 4 + b] : (a == b ? match : mismatch); pairs++; }
                }
            }
            if (pairs != 0) seq_score /= pairs;
            return sw * (sqrt(down(*A, i) * down(*B, j)) + sqrt(up(*A, i) * up(*B, j))) + seq_score;
        }
        const uint8_t a = code(*A, i), b = code(*B, j);
        if (a < 4 && b < 4) { seq_score += ribo ? ribo_bm[a * 4 + b] : (a == b ? match : mismatch); seq_score /= 1; }
        double res = sw * (sqrt(down(*A, i) * down(*B, j)) + sqrt(up(*A, i) * up(*B, j))) + seq_score;
        return res;
    }
};

template <class T> struct Grid {
    size_t cols = 0;
    std::vector<T> v;
    void reset(size_t r, size_t c) { cols = c; v.assign(r * c, T(0)); }
    T &operator()(size_t i, size_t j) { return v[i * cols + j]; }
};

// edge_probs.icc:76-162, evaluated in the same order and with the same operand types
template <class T>
void gotoh_pf(Grid<T> &zM, Grid<T> &zA, Grid<T> &zB, const std::vector<int> &lo, const std::vector<int> &hi, size_t lenA, size_t lenB,
              const EnvScore &sc, double open, double ext, double temp, bool free_left1, bool free_left2, bool local) {
    const double g_open = exp(open / temp);
    const double g_ext = exp(ext / temp);
    zM.reset(lenA + 1, lenB + 1); zA.reset(lenA + 1, lenB + 1); zB.reset(lenA + 1, lenB + 1);
    auto valid = [&](size_t i, size_t j) { return (size_t)lo[i] <= j && j <= (size_t)hi[i]; };
    if (valid(0, 0)) zM(0, 0) = local ? 0 : 1;
    if (lenA > 0 && valid(1, 0)) zA(1, 0) = g_open * g_ext;
    if (lenB > 0 && valid(0, 1)) zB(0, 1) = g_open * g_ext;
    for (size_t i = 2; i <= lenA; i++) {
        if (lo[i] > 0) break;
        zA(i, 0) = zA(i - 1, 0) * g_ext;
    }
    for (size_t j = std::max((size_t)lo[0], (size_t)2); j <= std::min((size_t)hi[0], lenB); j++) zB(0, j) = zB(0, j - 1) * g_ext;
    if (free_left2) for (size_t i = 1; i <= lenA; i++) zA(i, 0) += 1;
    if (free_left1) for (size_t j = 1; j <= lenB; j++) zB(0, j) += 1;
    for (size_t i = 1; i <= lenA; i++) {
        for (size_t j = std::max((size_t)lo[i], (size_t)1); j <= std::min((size_t)hi[i], lenB); j++) {
            double match_score_ij = sc.sigma((int)i, (int)j);
            double match_ij = exp(match_score_ij / temp);
            zM(i, j) = zM(i - 1, j - 1) * match_ij + zA(i - 1, j - 1) * match_ij + zB(i - 1, j - 1) * match_ij + (local ? match_ij : 0);
            zA(i, j) = zA(i - 1, j) * g_ext + zM(i - 1, j) * g_open * g_ext + zB(i - 1, j) * g_open * g_ext;
            zB(i, j) = zB(i, j - 1) * g_ext + zM(i, j - 1) * g_open * g_ext + zA(i, j - 1) * g_open * g_ext;
        }
    }
}

// edge_probs.icc:7-60 + :212-268 (trace probabilities), trace_controller.cc:565-597 (threshold + monotone closure)
template <class T>
void envelope_impl(Band &band, const Sequence &A, const Sequence &B, const Params &p) {
    const size_t lenA = A.len, lenB = B.len;
    EnvScore sc;
    sc.A = &A; sc.B = &B; sc.sw = p.struct_weight / 100.0; sc.match = p.match; sc.mismatch = p.mismatch; sc.ribo = p.use_ribosum; sc.ribo_bm = p.ribosum.bm;
    const double open = p.indel_opening / 100.0, ext = p.indel / 100.0, temp = p.temperature_alipf / 100.0;
    const bool local = p.sequ_local;
    Grid<T> zM, zA, zB, zMr, zAr, zBr;
    gotoh_pf<T>(zM, zA, zB, band.lo, band.hi, lenA, lenB, sc, open, ext, temp, p.fe_left1, p.fe_left2, local);
    // reversed problem: reversed sequences, up/down swapped, band mirrored, free end gaps swapped
    std::vector<int> rlo(lenA + 1), rhi(lenA + 1);
    for (size_t i = 0; i <= lenA; i++) { rhi[lenA - i] = (int)lenB - band.lo[i]; rlo[lenA - i] = (int)lenB - band.hi[i]; }
    sc.rev = true;
    gotoh_pf<T>(zMr, zAr, zBr, rlo, rhi, lenA, lenB, sc, open, ext, temp, p.fe_right1, p.fe_right2, local);
    T z;
    if (local) {
        z = 1;
        for (size_t i = 0; i <= lenA; i++)
            for (size_t j = 0; j <= lenB; j++) z += zM(i, j);
    } else {
        z = zM(lenA, lenB) + zA(lenA, lenB) + zB(lenA, lenB);
        if (p.fe_left2) for (size_t i = 0; i <= lenA; i++) z += zA(i, lenB);
        if (p.fe_left1) for (size_t j = 0; j <= lenB; j++) z += zB(lenA, j);
    }
    const double locality_add = local ? 1 : 0;
    const double g_open = exp(open / temp);
    for (size_t i = 0; i <= lenA; i++) {
        int new_min = band.hi[i], new_max = band.lo[i];
        for (size_t j = (size_t)std::max(band.lo[i], 0); j <= std::min((size_t)band.hi[i], lenB); j++) {
            const size_t ri = lenA - i, rj = lenB - j;
            T z_ij = zM(i, j) * (zMr(ri, rj) + zAr(ri, rj) + zBr(ri, rj) + locality_add) +
                     zA(i, j) * (zMr(ri, rj) + zAr(ri, rj) / g_open + zBr(ri, rj)) +
                     zB(i, j) * (zMr(ri, rj) + zAr(ri, rj) + zBr(ri, rj) / g_open);
            const double pr = (double)(z_ij / z);
            if (pr >= p.min_trace_probability) { new_min = std::min(new_min, (int)j); new_max = std::max(new_max, (int)j); }
        }
        band.lo[i] = std::max(band.lo[i], new_min);
        band.hi[i] = std::min(band.hi[i], new_max);
    }
    int run = 0;
    for (size_t i = 0; i <= lenA; i++) { band.hi[i] = std::max(band.hi[i], run); run = band.hi[i]; }
    run = band.hi[lenA];
    for (size_t i = lenA + 1; i-- > 0;) { band.lo[i] = std::min(band.lo[i], run); run = band.lo[i]; }
}

}  // namespace

void envelope_score_params(const Params &p, double bm[16], double *sw, double *open, double *ext, double *temp) {
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) bm[a * 4 + b] = p.use_ribosum ? p.ribosum.bm[a * 4 + b] : (a == b ? (double)p.match : (double)p.mismatch);
    *sw = p.struct_weight / 100.0; *open = p.indel_opening / 100.0; *ext = p.indel / 100.0; *temp = p.temperature_alipf / 100.0;
}

void restrict_band_by_envelope(Band &band, const Sequence &A, const Sequence &B, const Params &p) {
    if (!(p.min_trace_probability > 0.0)) return;  // main_helper.icc:416
    if (p.pf_double) envelope_impl<double>(band, A, B, p);
    else envelope_impl<long double>(band, A, B, p);  // locarna.cc:384-391: extended precision is forced
}

// ---------------------------------------------------------------------------------- arc matches, tasks
// arc_matches.cc:19-48 (validity), :130-188 (enumeration), :50-74 (inner arc matches), :313-355 (max right ends);
// aligner.cc:660-732 (task = left end pair)
void build_pair_problem(const Sequence &A, const Sequence &B, const Band &band, const Params &p, const ScoreTables &t, PairProblem &out, bool anchored,
                        const ProfileTables *profile) {
    out = PairProblem();
    const int n = A.len, m = B.len;
    const std::vector<int> &lo = band.lo, &hi = band.hi;
    auto valid = [&](int i, int j) { return lo[i] <= j && j <= hi[i]; };
    auto valid_match = [&](int i, int j) { return i >= 1 && j >= 1 && valid(i, j) && valid(i - 1, j - 1); };
    const long mdam = p.max_diff_am != -1 ? p.max_diff_am : std::max(n, m);
    const long mdat = p.max_diff_at_am != -1 ? p.max_diff_at_am : std::max(n, m);
    const std::vector<int> wA = arc_weights(A, p), wB = arc_weights(B, p);

    struct Run { int start, count; };
    std::unordered_map<int, Run> runs;  // key al * 4096 + bl
    std::vector<int> run_keys;
    for (int al = n; al >= 1; al--) {
        if (A.lcount[al] == 0) continue;
        for (int bl = std::min(hi[al], m); bl >= std::max(lo[al], 1); bl--) {
            if (B.lcount[bl] == 0 || !valid_match(al, bl)) continue;
            if (anchored && A.anchor_rank[al] != B.anchor_rank[bl]) continue;   // constraints.allowed_match of the left ends (arc_matches.cc:27)
            if (std::abs(al - bl) > mdat) continue;
            const int start = (int)out.am.size();
            for (int a = A.lptr[al]; a < A.lptr[al] + A.lcount[al]; a++) {
                const int ar = A.arcs[a].right;
                for (int b = B.lptr[bl]; b < B.lptr[bl] + B.lcount[bl]; b++) {
                    const int br = B.arcs[b].right;
                    if (!valid_match(ar, br)) continue;
                    if (anchored && A.anchor_rank[ar] != B.anchor_rank[br]) continue;   // ... and of the right ends (arc_matches.cc:28)
                    if (std::labs((long)(ar - al) - (long)(br - bl)) > mdam) continue;
                    if (std::abs(ar - br) > mdat) continue;
                    DevArcMatch x;
                    x.ends_a = (uint32_t)al | ((uint32_t)ar << 12);
                    x.ends_b = (uint32_t)bl | ((uint32_t)br << 12);
                    x.score = profile != nullptr ? profile->am_seq[(size_t)a * profile->arcsB + b] + wA[a] + wB[b] : arcmatch_score(t, p, A, B, a, b, wA, wB);
                    x.spos = -1; x.inner = -1; x.score_st = LB_NOSTACK;   // host-only inspection: no stacked scores
                    out.am.push_back(x); out.am_a.push_back(a); out.am_b.push_back(b);
                }
            }
            const int cnt = (int)out.am.size() - start;
            if (cnt > 0) { runs[al * 4096 + bl] = Run{start, cnt}; run_keys.push_back(al * 4096 + bl); }
        }
    }
    const int K = (int)out.am.size();
    // inner arc match: (al+1, ar-1, bl+1, br-1)
    for (int k = 0; k < K; k++) {
        const int al = out.am[k].ends_a & 0xfff, ar = out.am[k].ends_a >> 12, bl = out.am[k].ends_b & 0xfff, br = out.am[k].ends_b >> 12;
        auto it = runs.find((al + 1) * 4096 + bl + 1);
        if (it == runs.end()) continue;
        for (int x = it->second.start; x < it->second.start + it->second.count; x++)
            if ((int)(out.am[x].ends_a >> 12) == ar - 1 && (int)(out.am[x].ends_b >> 12) == br - 1) { out.am[k].inner = x; break; }
    }
    // S-order: stable sort of the L-order by (ar+br, ar); within equal keys the L-order (al desc, bl desc) is what
    // common_right_end_list uses (arc_matches.hh:188-220)
    std::vector<int> perm(K);
    for (int k = 0; k < K; k++) perm[k] = k;
    std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) {
        const int ax = out.am[x].ends_a >> 12, ay = out.am[y].ends_a >> 12;
        const int sx = ax + (int)(out.am[x].ends_b >> 12), sy = ay + (int)(out.am[y].ends_b >> 12);
        if (sx != sy) return sx < sy;
        const int px = (int)(out.am[x].ends_a & 0xfff) + (int)(out.am[x].ends_b & 0xfff), py = (int)(out.am[y].ends_a & 0xfff) + (int)(out.am[y].ends_b & 0xfff);
        return px > py;   // source anti-diagonal descending (kernels.cu Stream3)
    });
    out.ent.resize(K);
    out.sptr.assign(n + m + 3, 0);
    for (int s = 0; s < K; s++) {
        DevArcMatch &x = out.am[perm[s]];
        x.spos = s;
        const int al = x.ends_a & 0xfff, ar = x.ends_a >> 12, bl = x.ends_b & 0xfff, br = x.ends_b >> 12;
        out.ent[s].x = (uint32_t)(al - 1) | ((uint32_t)(bl - 1) << 16);
        out.ent[s].y = (uint32_t)ar | ((uint32_t)br << 16);
        out.ent[s].d = LB_NEG;
        out.ent[s].s = al + bl - 2;
        out.sptr[ar + br + 1]++;
    }
    for (int s = 0; s + 1 < (int)out.sptr.size(); s++) out.sptr[s + 1] += out.sptr[s];
    // tasks
    const bool nolp = p.no_lonely_pairs;
    for (int key : run_keys) {
        const Run r = runs[key];
        const int al = key / 4096, bl = key % 4096;
        int max_ar = 0, max_br = 0;
        for (int x = r.start; x < r.start + r.count; x++) {
            if (nolp && out.am[x].inner < 0) continue;
            max_ar = std::max(max_ar, (int)(out.am[x].ends_a >> 12));
            max_br = std::max(max_br, (int)(out.am[x].ends_b >> 12));
        }
        if (max_ar == 0) continue;
        DevTask tk;
        tk.pair = 0;
        tk.al = (short)(nolp ? al + 1 : al); tk.bl = (short)(nolp ? bl + 1 : bl);
        tk.R = (short)(nolp ? max_ar - 2 : max_ar - 1); tk.C = (short)(nolp ? max_br - 2 : max_br - 1);
        tk.run_start = r.start; tk.run_count = r.count;
        out.tasks.push_back(tk);
        int umax = 0;
        for (int i = tk.al + 1; i <= tk.R; i++) {
            const int jl = std::max((int)tk.bl + 1, lo[i]), jh = std::min((int)tk.C, hi[i]);
            if (jh >= jl) { out.cells += (uint64_t)(jh - jl + 1) * (p.struct_local ? 4 : 1); umax = (i - tk.al) + (jh - tk.bl); }
        }
        {   // entries streamed: right ends on local anti-diagonals 8..umax
            const int s0 = std::min(tk.al + tk.bl + 8, n + m + 1), s1 = std::min(tk.al + tk.bl + umax + 1, n + m + 1);
            if (s1 > s0) out.terms += (uint64_t)(out.sptr[s1] - out.sptr[s0]);
        }
    }
    // bound on the diagonals (j - i) touched by any box: band cells, plus row 0 of the top level box, which is
    // initialised from column 0 even where the band starts further right (aligner.cc:345-357)
    int dmin = 0, dmax = 0;
    for (int i = 0; i <= n; i++) {
        dmin = std::min(dmin, lo[i] - i); dmax = std::max(dmax, hi[i] - i);
        if (i >= 1) { const int jl = std::max(1, lo[i]), jh = std::min(m, hi[i]); if (jh >= jl) out.cells += jh - jl + 1; }
    }
    out.terms += (uint64_t)out.sptr[n + m + 1] - out.sptr[std::min(8, n + m + 1)];  // top level box
    out.wd_bound = dmax - dmin + 1;
    out.max_box_words = (n + m + 1) * ((out.wd_bound + 1) / 2);  // anti-diagonal major box: (umax + 1) * diagonal pairs
}

}  // namespace lb200
