// locarna_b200 -- command line front end with the flag surface of the reference's `locarna` (src/locarna.cc:83-272)
// for the modes the B200 path implements. Output follows src/locarna.cc:766-948: "Score: N", the alignment in
// CLUSTAL-like blocks of --width columns, and optionally --clustal <file> with the header mlocarna parses
// ("CLUSTAL W --- LocARNA 2.0.1 --- Score: N", main_helper.icc:556-566, mlocarna:3516-3527).
// Options of modes that are not implemented are recognised and rejected with an error (exit code 255 like the
// reference's `return -1`), never ignored.
#include <getopt.h>

#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "locarna_b200.hh"

using namespace LocARNA_B200;

namespace {
struct Opt { const char *name; int has_arg; int id; };
enum {
    O_INDEL = 1000, O_INDEL_OPENING, O_RIBOSUM_FILE, O_USE_RIBOSUM, O_MATCH, O_MISMATCH, O_UNPAIRED_PENALTY, O_STRUCT_WEIGHT, O_EXP_PROB,
    O_TAU, O_EXCLUSION, O_STACKING, O_NEW_STACKING, O_STRUCT_LOCAL, O_SEQU_LOCAL, O_FREE_ENDGAPS, O_NORMALIZED, O_PENALIZED, O_WIDTH,
    O_CLUSTAL, O_STOCKHOLM, O_PP, O_UNUSED_PP_PLACEHOLDER, O_LOCAL_OUTPUT, O_LOCAL_FILE_OUTPUT, O_POS_OUTPUT, O_WRITE_STRUCTURE, O_MIN_PROB, O_MAX_BPS_LENGTH_RATIO,
    O_MAX_DIFF_AM, O_MAX_DIFF, O_MAX_DIFF_AT_AM, O_MIN_TRACE_PROB, O_NOLP, O_MAXBPSPAN, O_TEMPERATURE_ALIPF, O_CONSENSUS_STRUCTURE,
    O_WRITE_ARCMATCH_SCORES, O_KBEST, O_BETTER, O_MAX_DIFF_ALN, O_MAX_DIFF_PW_ALN, O_MAX_DIFF_RELAX, O_UNSUPPORTED, O_MATCHPROB_PARAM, O_DEVICE, O_VERSION, O_QUIET, O_VERBOSE, O_HELP
};
bool parse_bool(const char *s) {
    const std::string v = s ? s : "";  // options.cc:867-880
    if (v == "t" || v == "true" || v == "on" || v == "1") return true;
    if (v == "f" || v == "false" || v == "off" || v == "0") return false;
    std::cerr << "ERROR: cannot parse boolean value \"" << v << "\"" << std::endl;
    exit(255);
}
}  // namespace

int main(int argc, char **argv) {
    static const struct option longopts[] = {
        {"indel", required_argument, 0, O_INDEL}, {"indel-opening", required_argument, 0, O_INDEL_OPENING},
        {"ribosum-file", required_argument, 0, O_RIBOSUM_FILE}, {"use-ribosum", required_argument, 0, O_USE_RIBOSUM},
        {"match", required_argument, 0, O_MATCH}, {"mismatch", required_argument, 0, O_MISMATCH},
        {"unpaired-penalty", required_argument, 0, O_UNPAIRED_PENALTY}, {"struct-weight", required_argument, 0, O_STRUCT_WEIGHT},
        {"exp-prob", required_argument, 0, O_EXP_PROB}, {"tau", required_argument, 0, O_TAU}, {"exclusion", required_argument, 0, O_EXCLUSION},
        {"stacking", no_argument, 0, O_STACKING}, {"new-stacking", no_argument, 0, O_NEW_STACKING},
        {"struct-local", required_argument, 0, O_STRUCT_LOCAL}, {"sequ-local", required_argument, 0, O_SEQU_LOCAL},
        {"free-endgaps", required_argument, 0, O_FREE_ENDGAPS}, {"normalized", required_argument, 0, O_NORMALIZED},
        {"penalized", required_argument, 0, O_PENALIZED}, {"width", required_argument, 0, O_WIDTH}, {"clustal", required_argument, 0, O_CLUSTAL},
        {"stockholm", required_argument, 0, O_STOCKHOLM}, {"pp", required_argument, 0, O_PP}, {"local-output", no_argument, 0, O_LOCAL_OUTPUT},
        {"local-file-output", no_argument, 0, O_LOCAL_FILE_OUTPUT}, {"pos-output", no_argument, 0, O_POS_OUTPUT},
        {"write-structure", no_argument, 0, O_WRITE_STRUCTURE}, {"min-prob", required_argument, 0, O_MIN_PROB},
        {"max-bps-length-ratio", required_argument, 0, O_MAX_BPS_LENGTH_RATIO}, {"max-diff-am", required_argument, 0, O_MAX_DIFF_AM},
        {"max-diff", required_argument, 0, O_MAX_DIFF}, {"max-diff-at-am", required_argument, 0, O_MAX_DIFF_AT_AM},
        {"min-trace-probability", required_argument, 0, O_MIN_TRACE_PROB}, {"noLP", no_argument, 0, O_NOLP},
        {"maxBPspan", required_argument, 0, O_MAXBPSPAN}, {"temperature-alipf", required_argument, 0, O_TEMPERATURE_ALIPF},
        {"consensus-structure", required_argument, 0, O_CONSENSUS_STRUCTURE}, {"write-arcmatch-scores", required_argument, 0, O_WRITE_ARCMATCH_SCORES},
        // recognised, not implemented on the B200 path
        {"max-diff-aln", required_argument, 0, O_MAX_DIFF_ALN}, {"max-diff-pw-aln", required_argument, 0, O_MAX_DIFF_PW_ALN},
        {"max-diff-relax", no_argument, 0, O_MAX_DIFF_RELAX}, {"kbest", required_argument, 0, O_KBEST}, {"better", required_argument, 0, O_BETTER},
        {"mea-alignment", no_argument, 0, O_UNSUPPORTED}, {"match-prob-method", required_argument, 0, O_UNSUPPORTED},
        {"read-match-probs", required_argument, 0, O_UNSUPPORTED}, {"write-match-probs", required_argument, 0, O_UNSUPPORTED},
        {"read-arcmatch-scores", required_argument, 0, O_UNSUPPORTED}, {"read-arcmatch-probs", required_argument, 0, O_UNSUPPORTED},
        {"write-trace-probs", required_argument, 0, O_UNSUPPORTED},
        // parameters of the base-match probability computation (locarna.cc:236-246): they only act in MEA mode / --write-match-probs, which
        // are refused above; mlocarna passes them on when the user sets them (mlocarna:1181-1218). Listed so that "--temperature" is not
        // taken for an abbreviation of --temperature-alipf.
        {"temperature", required_argument, 0, O_MATCHPROB_PARAM}, {"pf-struct-weight", required_argument, 0, O_MATCHPROB_PARAM},
        {"probcons-file", required_argument, 0, O_MATCHPROB_PARAM},
        {"alifold-consensus-dp", no_argument, 0, O_UNSUPPORTED}, {"ribofit", required_argument, 0, O_UNSUPPORTED},
        {"relaxed-anchors", no_argument, 0, O_UNSUPPORTED}, {"score-components", no_argument, 0, O_UNSUPPORTED},
        {"extended-pf", no_argument, 0, O_UNSUPPORTED}, {"quad-pf", no_argument, 0, O_UNSUPPORTED},
        {"device", required_argument, 0, O_DEVICE}, {"version", no_argument, 0, O_VERSION}, {"quiet", no_argument, 0, O_QUIET},
        {"verbose", no_argument, 0, O_VERBOSE}, {"help", no_argument, 0, O_HELP}, {"stopwatch", no_argument, 0, O_VERBOSE},
        {0, 0, 0, 0}};
    ScoringParams sp;
    AlignerParams ap;
    int width = 120, device = 0;
    double min_prob = 0.001;
    bool quiet = false, local_output = false, local_file_output = false, write_structure = false, pos_output = false;
    std::string clustal, stockholm, arcmatch_scores_file;
    int max_bp_span = -1;
    double max_bps_length_ratio = 0.0;
    bool verbose = false;
    bool struct_local = false, struct_local_given = false, sequ_local = false, sequ_local_given = false, normalized = false, penalized = false;
    long normalized_L = 0, position_penalty = 0, subopt_threshold = -1000000;
    bool subopt = false, max_diff_relax = false;
    std::string max_diff_alignment_file, max_diff_pw_alignment, pp_file;
    int kbest_k = -1;
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "i:m:M:s:e:t:E:w:Lp:D:d:PqvVh", longopts, &idx)) != -1) {
        switch (c) {
            case 'i': case O_INDEL: sp.indel = atoi(optarg); break;
            case O_INDEL_OPENING: sp.indel_opening = atoi(optarg); break;
            case O_RIBOSUM_FILE: sp.ribosum_file = optarg; break;   // read when the aligner is set up (main_helper.icc:311-350)
            case O_USE_RIBOSUM: sp.use_ribosum = parse_bool(optarg); break;
            case 'm': case O_MATCH: sp.match = atoi(optarg); break;
            case 'M': case O_MISMATCH: sp.mismatch = atoi(optarg); break;
            case O_UNPAIRED_PENALTY: sp.unpaired_penalty = atoi(optarg); break;
            case 's': case O_STRUCT_WEIGHT: sp.struct_weight = atoi(optarg); break;
            case 't': case O_TAU: sp.tau_factor = atoi(optarg); break;
            case 'E': case O_EXCLUSION: sp.exclusion = atoi(optarg); break;
            case O_STRUCT_LOCAL: struct_local = parse_bool(optarg); struct_local_given = true; break;
            case O_SEQU_LOCAL: sequ_local = parse_bool(optarg); sequ_local_given = true; break;
            case O_MAX_DIFF_RELAX: max_diff_relax = true; break;
            case O_MAX_DIFF_ALN: max_diff_alignment_file = optarg; break;
            case O_MAX_DIFF_PW_ALN: max_diff_pw_alignment = optarg; break;
            case O_KBEST: subopt = true; kbest_k = atoi(optarg); break;          // locarna.cc:203-207
            case O_BETTER: subopt = true; subopt_threshold = atol(optarg); break;
            case O_NORMALIZED: normalized = true; normalized_L = atol(optarg); break;
            case O_PENALIZED: penalized = true; position_penalty = atol(optarg); break;
            case O_FREE_ENDGAPS: ap.free_endgaps(optarg); break;
            case 'w': case O_WIDTH: width = atoi(optarg); break;
            case O_CLUSTAL: clustal = optarg; break;
            case O_STOCKHOLM: stockholm = optarg; break;
            case 'L': case O_LOCAL_OUTPUT: local_output = true; break;
            case O_LOCAL_FILE_OUTPUT: local_file_output = true; break;
            case 'P': case O_POS_OUTPUT: pos_output = true; break;
            case O_WRITE_STRUCTURE: write_structure = true; break;
            case 'p': case O_MIN_PROB: min_prob = atof(optarg); break;
            case 'D': case O_MAX_DIFF_AM: ap.max_diff_am(atoi(optarg)); break;
            case 'd': case O_MAX_DIFF: ap.max_diff(atoi(optarg)); break;
            case O_MAX_DIFF_AT_AM: ap.max_diff_at_am(atoi(optarg)); break;
            case O_MIN_TRACE_PROB: ap.min_trace_probability(atof(optarg)); break;
            case O_NOLP: ap.no_lonely_pairs(true); break;
            case O_TEMPERATURE_ALIPF: sp.temperature_alipf = atoi(optarg); break;
            case O_CONSENSUS_STRUCTURE:
                if (std::string(optarg) != "none") { std::cerr << "ERROR: --consensus-structure " << optarg << " needs ViennaRNA; only \"none\" is supported." << std::endl; return 255; }
                break;
            case O_MAX_BPS_LENGTH_RATIO: max_bps_length_ratio = atof(optarg); break;
            case O_MAXBPSPAN: max_bp_span = atoi(optarg); break;
            case 'e': case O_EXP_PROB: sp.exp_prob = atof(optarg); break;
            case O_WRITE_ARCMATCH_SCORES: arcmatch_scores_file = optarg; break;
            case O_STACKING: sp.stacking = true; break;
            case O_NEW_STACKING: sp.new_stacking = true; break;
            case O_PP: pp_file = optarg; break;
            case O_MATCHPROB_PARAM: break;
            case O_UNUSED_PP_PLACEHOLDER:
            case O_UNSUPPORTED:
                std::cerr << "ERROR: option --" << (idx >= 0 && longopts[idx].name ? longopts[idx].name : "?")
                          << " selects a mode that locarna_b200 does not implement." << std::endl;
                return 255;
            case O_DEVICE: device = atoi(optarg); break;
            case 'q': case O_QUIET: quiet = true; break;
            case 'v': case O_VERBOSE: verbose = true; break;
            case 'V': case O_VERSION: std::cout << "locarna_b200 (LocARNA 2.0.1 pairwise path, B200)" << std::endl; return 0;
            case 'h': case O_HELP: std::cout << "usage: locarna_b200 [options as locarna] <fileA.pp> <fileB.pp>" << std::endl; return 0;
            default: return 255;
        }
    }
    if (argc - optind != 2) { std::cerr << "ERROR: expected two input files (PP 2.0)." << std::endl; return 255; }
    if (normalized && penalized) {   // locarna.cc:369-373
        std::cerr << "One cannot specify penalized and normalized " << "simultaneously." << std::endl;
        return 255;
    }
    if (penalized && !sequ_local_given) sequ_local = true;   // locarna.cc:417-424
    if (normalized) {                                        // locarna.cc:426-447
        if (!sequ_local_given) sequ_local = true;
        else if (!sequ_local) { std::cerr << "ERROR: Cannot run normalized alignment " << "without --sequ_local on." << std::endl; return 255; }
        if (struct_local_given && struct_local) { std::cerr << "ERROR: Normalized structure local alignment " << "not supported." << std::endl; return 255; }
        struct_local = false;
    }
    ap.struct_local(struct_local).sequ_local(sequ_local);
    if (max_diff_pw_alignment != "" && max_diff_alignment_file != "") {   // locarna.cc:504-508
        std::cerr << "Cannot simultaneously use options --max-diff-pw-alignment" << " and --max-diff-alignment-file." << std::endl;
        return 255;
    }
    if (sp.stacking && sp.exp_prob < 0) {   // locarna.cc:406-414
        std::cerr << "WARNING: stacking turned off. "
                  << "Stacking requires setting a background probability "
                  << "explicitely (option --exp-prob)." << std::endl;
        sp.stacking = false;
    }
    try {
        RnaData rnaA(argv[optind], min_prob, max_bps_length_ratio, max_bp_span), rnaB(argv[optind + 1], min_prob, max_bps_length_ratio, max_bp_span);
        ap.seqA(&rnaA).seqB(&rnaB).scoring(sp).min_prob(min_prob);
        std::unique_ptr<MultipleAlignment> multiple_ref_alignment;                   // locarna.cc:514-548
        if (max_diff_alignment_file != "") {
            multiple_ref_alignment.reset(new MultipleAlignment(max_diff_alignment_file));
        } else if (max_diff_pw_alignment != "") {
            const size_t amp = max_diff_pw_alignment.find('&');
            if (amp == std::string::npos || max_diff_pw_alignment.find('&', amp + 1) != std::string::npos) {
                std::cerr << "Invalid argument to --max-diff-pw-alignemnt; require " << "exactly one '&' separating the alignment strings." << std::endl;
                return 255;
            }
            const std::string rowA = max_diff_pw_alignment.substr(0, amp), rowB = max_diff_pw_alignment.substr(amp + 1);
            if (rowA.length() != rowB.length()) {
                std::cerr << "Invalid argument to --max-diff-pw-alignemnt;" << " alignment strings have unequal lengths." << std::endl;
                return 255;
            }
            multiple_ref_alignment.reset(new MultipleAlignment("A", "B", rowA, rowB));
        }
        ap.reference_alignment(multiple_ref_alignment.get(), max_diff_relax);
        Aligner aligner(ap, device);
        if (!arcmatch_scores_file.empty()) {                                        // locarna.cc:705-720: write and return without aligning
            if (verbose) std::cout << "Write arcmatch scores to file " << arcmatch_scores_file << " and exit." << std::endl;
            aligner.arc_matches().write_arcmatch_scores(arcmatch_scores_file);
            return 0;
        }
        if (subopt) {                                                                // locarna.cc:740-747
            aligner.suboptimal(kbest_k, subopt_threshold, normalized, normalized_L, (size_t)width, verbose, local_output, pos_output, write_structure);
            return 0;
        }
        const infty_score_t score = normalized ? aligner.normalized_align(normalized_L, verbose)     // locarna.cc:751-763
                                    : penalized ? aligner.penalized_align(position_penalty) : aligner.align();
        if (!quiet) std::cout << "Score: " << score << std::endl << std::endl;     // locarna.cc:769-771
        aligner.trace();
        const Alignment &alignment = aligner.get_alignment();
        int rc = 0;
        if (!clustal.empty()) {                                                     // main_helper.icc:556-584
            std::ofstream out(clustal.c_str());
            if (out.good()) {
                MultipleAlignment ma(alignment, local_file_output);
                out << "CLUSTAL W --- LocARNA 2.0.1";
                // "for legacy, clustal files of pairwise alignments contain the score" (main_helper.icc:562-568): not for profile input
                if (alignment.num_rowsA() == 1 && alignment.num_rowsB() == 1) out << " --- Score: " << score;
                out << std::endl << std::endl;
                if (write_structure) {
                    ma.prepend(MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureA(local_file_output)));
                    ma.append(MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureB(local_file_output)));
                }
                ma.write(out, width);
            } else { std::cerr << "ERROR: Cannot write to " << clustal << "." << std::endl; rc = 255; }
        }
        if (!stockholm.empty()) {                                                   // main_helper.icc:590-615 (no consensus structure: --consensus-structure none)
            std::ofstream out(stockholm.c_str());
            if (out.good()) {
                MultipleAlignment ma(alignment, local_file_output);
                out << "# STOCKHOLM 1.0" << std::endl << "#=GF CC Generated by LocARNA 2.0.1" << std::endl << "#=GF SQ " << ma.num_of_rows() << std::endl << std::endl;
                ma.write(out, width, MultipleAlignment::FormatType::STOCKHOLM);
            } else { std::cerr << "ERROR: Cannot write to " << stockholm << "." << std::endl; rc = 255; }
        }
        if (!pp_file.empty()) {                                                      // main_helper.icc:617-634: alignment + consensus dot plot
            std::ofstream out(pp_file.c_str());
            if (out.good()) aligner.write_pp(out, local_file_output, sp.exp_prob);
            else { std::cerr << "ERROR: Cannot write to " << pp_file << std::endl; rc = 255; }
        }
        if (pos_output) {                                                           // locarna.cc:879-888
            const auto start = alignment.start_positions(), end = alignment.end_positions();
            std::cout << "HIT " << score << " " << start.first << " " << start.second << " " << end.first << " " << end.second << " " << std::endl;
            std::cout << std::endl;
        }
        if ((!pos_output && !quiet) || local_output) {                              // locarna.cc:890-944
            MultipleAlignment ma(alignment, local_output);
            if (write_structure) {
                ma.prepend(MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureA(local_output)));
                ma.append(MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureB(local_output)));
            }
            if (pos_output)
                std::cout << "\t+" << alignment.start_positions().first << std::endl << "\t+" << alignment.start_positions().second << std::endl
                          << std::endl << std::endl;
            ma.write(std::cout, width);
            if (pos_output)
                std::cout << std::endl << "\t+" << alignment.end_positions().first << std::endl << "\t+" << alignment.end_positions().second << std::endl
                          << std::endl;
        }
        if (!quiet) std::cout << std::endl;
        return rc;
    } catch (failure &f) {
        std::cerr << "ERROR: " << f.what() << std::endl;
        return 255;
    }
}
