// locarna_refmain_b200: the reference's own alignment pipeline - the body of run_and_report() of src/locarna.cc, cut from the reference
// tree at build time and compiled here UNMODIFIED - running on the B200 library through include/locarna_b200_compat.hh.
// It exists to prove the drop-in boundary by compilation (tests/test_boundary.py); the everyday front end is locarna_b200.
//
// What this file supplies around the block: the command line parameter object `clp` (the reference defines it with its option table,
// src/locarna.cc:57-272), DO_TRACE (:52), pf_score_t (:384-394) and the output stage after the block (src/locarna.cc:790-944,
// main_helper.icc:535-637), which is the same code as in locarna_main.cc.
#include <getopt.h>

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "locarna_b200_compat.hh"

namespace LocARNA = LocARNA_B200::compat;
using namespace LocARNA;

const bool DO_TRACE = true;

struct command_line_parameters {   // the fields the pipeline reads (src/locarna.cc:57-82, main_helper.icc std_command_line_parameters)
    std::string fileA, fileB, ribosum_file = "RIBOSUM85_60", free_endgaps = "----", max_diff_pw_alignment, max_diff_alignment_file;
    std::string arcmatch_scores_infile, arcmatch_scores_outfile, matchprobs_outfile, clustal_out;
    int match = 50, mismatch = 0, indel = -150, indel_opening = -750, unpaired_penalty = 0, struct_weight = 200, tau = 50, exclusion = 0;
    int temperature_alipf = 300, max_diff_am = -1, max_diff_at_am = -1, max_diff = -1, max_bp_span = -1, width = 120;
    int mea_alpha = 0, mea_beta = 200, mea_gamma = 100, probability_scale = 10000, kbest_k = -1, subopt_threshold = -1000000;
    long normalized_L = 0, position_penalty = 0;
    double min_prob = 0.001, max_bps_length_ratio = 0.0, exp_prob = 0.0, min_trace_probability = 1e-4;
    bool use_ribosum = true, ribofit = false, exp_prob_given = false, stacking = false, new_stacking = false, no_lonely_pairs = false;
    bool struct_local = false, struct_local_given = false, sequ_local = false, sequ_local_given = false, relaxed_anchors = false, max_diff_relax = false;
    bool normalized = false, penalized = false, subopt = false, mea_alignment = false, mea_gapcost = false;
    bool read_matchprobs = false, write_matchprobs = false, read_arcmatch_scores = false, read_arcmatch_probs = false, write_arcmatch_scores = false;
    bool write_traceprobs = false, verbose = false, quiet = false, local_output = false, local_file_output = false, pos_output = false, write_structure = false;
};
static command_line_parameters clp;

static bool parse_bool(const char *s) {
    const std::string v = s ? s : "";
    if (v == "t" || v == "true" || v == "on" || v == "1") return true;
    if (v == "f" || v == "false" || v == "off" || v == "0") return false;
    std::cerr << "ERROR: cannot parse boolean value \"" << v << "\"" << std::endl;
    exit(255);
}

typedef long double pf_score_t;

static int run_and_report() {
    typedef std::vector<int>::size_type size_type;
    infty_score_t score_out;
    std::unique_ptr<Alignment> alignment_out;
    {
        // ---- the reference's run_and_report() body (cut from /root/reference/src/locarna.cc by tools/extract_ref_block.py), verbatim
#include "refmain_block.inc"
        // ----
        score_out = score;
        alignment_out = std::move(alignment);
    }
    // ---- output stage (src/locarna.cc:790-944, main_helper.icc:556-584), as in locarna_main.cc
    const infty_score_t score = score_out;
    const Alignment &alignment = *alignment_out;
    int rc = 0;
    if (!clp.clustal_out.empty()) {
        std::ofstream out(clp.clustal_out.c_str());
        if (out.good()) {
            LocARNA_B200::MultipleAlignment ma(alignment, clp.local_file_output);
            out << "CLUSTAL W --- LocARNA 2.0.1";
            if (alignment.num_rowsA() == 1 && alignment.num_rowsB() == 1) out << " --- Score: " << score;   // main_helper.icc:562-568
            out << std::endl << std::endl;
            if (clp.write_structure) {
                ma.prepend(LocARNA_B200::MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureA(clp.local_file_output)));
                ma.append(LocARNA_B200::MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureB(clp.local_file_output)));
            }
            ma.write(out, clp.width);
        } else { std::cerr << "ERROR: Cannot write to " << clp.clustal_out << "." << std::endl; rc = -1; }
    }
    if (clp.pos_output) {
        const auto start = alignment.start_positions(), end = alignment.end_positions();
        std::cout << "HIT " << score << " " << start.first << " " << start.second << " " << end.first << " " << end.second << " " << std::endl;
        std::cout << std::endl;
    }
    if ((!clp.pos_output && !clp.quiet) || clp.local_output) {
        LocARNA_B200::MultipleAlignment ma(alignment, clp.local_output);
        if (clp.write_structure) {
            ma.prepend(LocARNA_B200::MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureA(clp.local_output)));
            ma.append(LocARNA_B200::MultipleAlignment::SeqEntry("", alignment.dot_bracket_structureB(clp.local_output)));
        }
        if (clp.pos_output)
            std::cout << "\t+" << alignment.start_positions().first << std::endl << "\t+" << alignment.start_positions().second << std::endl
                      << std::endl << std::endl;
        ma.write(std::cout, clp.width);
        if (clp.pos_output)
            std::cout << std::endl << "\t+" << alignment.end_positions().first << std::endl << "\t+" << alignment.end_positions().second << std::endl
                      << std::endl;
    }
    if (!clp.quiet) std::cout << std::endl;
    return rc;
}

int main(int argc, char **argv) {
    enum { O_INDEL_OPENING = 1000, O_RIBOSUM_FILE, O_USE_RIBOSUM, O_UNPAIRED_PENALTY, O_STRUCT_LOCAL, O_SEQU_LOCAL, O_FREE_ENDGAPS, O_MAX_DIFF_AT_AM, O_MIN_TRACE_PROB, O_NOLP,
           O_MAXBPSPAN, O_MAX_BPS_LENGTH_RATIO, O_TEMPERATURE_ALIPF, O_CLUSTAL, O_LOCAL_FILE_OUTPUT, O_WRITE_STRUCTURE, O_WRITE_AMS, O_STACKING, O_NORMALIZED,
           O_PENALIZED, O_KBEST, O_BETTER, O_MAX_DIFF_ALN, O_MAX_DIFF_PW_ALN, O_MAX_DIFF_RELAX };
    static const struct option longopts[] = {
        {"indel", required_argument, 0, 'i'}, {"indel-opening", required_argument, 0, O_INDEL_OPENING}, {"use-ribosum", required_argument, 0, O_USE_RIBOSUM}, {"ribosum-file", required_argument, 0, O_RIBOSUM_FILE},
        {"match", required_argument, 0, 'm'}, {"mismatch", required_argument, 0, 'M'}, {"unpaired-penalty", required_argument, 0, O_UNPAIRED_PENALTY},
        {"struct-weight", required_argument, 0, 's'}, {"exp-prob", required_argument, 0, 'e'}, {"tau", required_argument, 0, 't'},
        {"exclusion", required_argument, 0, 'E'}, {"struct-local", required_argument, 0, O_STRUCT_LOCAL}, {"sequ-local", required_argument, 0, O_SEQU_LOCAL},
        {"free-endgaps", required_argument, 0, O_FREE_ENDGAPS}, {"min-prob", required_argument, 0, 'p'}, {"max-diff-am", required_argument, 0, 'D'},
        {"max-diff", required_argument, 0, 'd'}, {"max-diff-at-am", required_argument, 0, O_MAX_DIFF_AT_AM}, {"min-trace-probability", required_argument, 0, O_MIN_TRACE_PROB},
        {"noLP", no_argument, 0, O_NOLP}, {"maxBPspan", required_argument, 0, O_MAXBPSPAN}, {"max-bps-length-ratio", required_argument, 0, O_MAX_BPS_LENGTH_RATIO},
        {"temperature-alipf", required_argument, 0, O_TEMPERATURE_ALIPF}, {"width", required_argument, 0, 'w'}, {"clustal", required_argument, 0, O_CLUSTAL},
        {"local-output", no_argument, 0, 'L'}, {"local-file-output", no_argument, 0, O_LOCAL_FILE_OUTPUT}, {"pos-output", no_argument, 0, 'P'},
        {"write-structure", no_argument, 0, O_WRITE_STRUCTURE}, {"write-arcmatch-scores", required_argument, 0, O_WRITE_AMS}, {"stacking", no_argument, 0, O_STACKING},
        {"normalized", required_argument, 0, O_NORMALIZED}, {"penalized", required_argument, 0, O_PENALIZED}, {"kbest", required_argument, 0, O_KBEST}, {"max-diff-aln", required_argument, 0, O_MAX_DIFF_ALN}, {"max-diff-relax", no_argument, 0, O_MAX_DIFF_RELAX}, {"max-diff-pw-aln", required_argument, 0, O_MAX_DIFF_PW_ALN}, {"better", required_argument, 0, O_BETTER},
        {"quiet", no_argument, 0, 'q'}, {"verbose", no_argument, 0, 'v'}, {0, 0, 0, 0}};
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "i:m:M:s:e:t:E:w:Lp:D:d:Pqv", longopts, &idx)) != -1) {
        switch (c) {
            case 'i': clp.indel = atoi(optarg); break;
            case O_INDEL_OPENING: clp.indel_opening = atoi(optarg); break;
            case O_USE_RIBOSUM: clp.use_ribosum = parse_bool(optarg); break;
            case O_RIBOSUM_FILE: clp.ribosum_file = optarg; break;
            case 'm': clp.match = atoi(optarg); break;
            case 'M': clp.mismatch = atoi(optarg); break;
            case O_UNPAIRED_PENALTY: clp.unpaired_penalty = atoi(optarg); break;
            case 's': clp.struct_weight = atoi(optarg); break;
            case 'e': clp.exp_prob = atof(optarg); clp.exp_prob_given = true; break;
            case 't': clp.tau = atoi(optarg); break;
            case 'E': clp.exclusion = atoi(optarg); break;
            case O_STRUCT_LOCAL: clp.struct_local = parse_bool(optarg); clp.struct_local_given = true; break;
            case O_SEQU_LOCAL: clp.sequ_local = parse_bool(optarg); clp.sequ_local_given = true; break;
            case O_FREE_ENDGAPS: clp.free_endgaps = optarg; break;
            case 'p': clp.min_prob = atof(optarg); break;
            case 'D': clp.max_diff_am = atoi(optarg); break;
            case 'd': clp.max_diff = atoi(optarg); break;
            case O_MAX_DIFF_AT_AM: clp.max_diff_at_am = atoi(optarg); break;
            case O_MIN_TRACE_PROB: clp.min_trace_probability = atof(optarg); break;
            case O_NOLP: clp.no_lonely_pairs = true; break;
            case O_MAXBPSPAN: clp.max_bp_span = atoi(optarg); break;
            case O_MAX_BPS_LENGTH_RATIO: clp.max_bps_length_ratio = atof(optarg); break;
            case O_TEMPERATURE_ALIPF: clp.temperature_alipf = atoi(optarg); break;
            case 'w': clp.width = atoi(optarg); break;
            case O_CLUSTAL: clp.clustal_out = optarg; break;
            case 'L': clp.local_output = true; break;
            case O_LOCAL_FILE_OUTPUT: clp.local_file_output = true; break;
            case 'P': clp.pos_output = true; break;
            case O_WRITE_STRUCTURE: clp.write_structure = true; break;
            case O_WRITE_AMS: clp.write_arcmatch_scores = true; clp.arcmatch_scores_outfile = optarg; break;
            case O_STACKING: clp.stacking = true; break;
            case O_NORMALIZED: clp.normalized = true; clp.normalized_L = atol(optarg); break;
            case O_PENALIZED: clp.penalized = true; clp.position_penalty = atol(optarg); break;
            case O_MAX_DIFF_RELAX: clp.max_diff_relax = true; break;
            case O_MAX_DIFF_ALN: clp.max_diff_alignment_file = optarg; break;
            case O_MAX_DIFF_PW_ALN: clp.max_diff_pw_alignment = optarg; break;
            case O_KBEST: clp.subopt = true; clp.kbest_k = atoi(optarg); break;
            case O_BETTER: clp.subopt = true; clp.subopt_threshold = atoi(optarg); break;
            case 'q': clp.quiet = true; break;
            case 'v': clp.verbose = true; break;
            default: return 255;
        }
    }
    if (argc - optind != 2) { std::cerr << "ERROR: expected two input files (PP 2.0)." << std::endl; return 255; }
    clp.fileA = argv[optind]; clp.fileB = argv[optind + 1];
    try {
        return run_and_report() == 0 ? 0 : 255;
    } catch (failure &f) {   // the reference lets failures of the aligner objects reach main's caller; report them as the B200 front end does
        std::cerr << "ERROR: " << f.what() << std::endl;
        return 255;
    }
}
