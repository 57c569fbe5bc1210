// locarna_p_b200 -- command line front end with the flag surface of the reference's `locarna_p` (src/locarna_p.cc:95-200) over the
// B200 LocARNA-P path: "Partition function: Z" on stdout (locarna_p.cc:483-487), --write-arcmatch-probs / --write-basematch-probs
// files in the reference's line format (aligner_p.icc:1404-1435). Options of modes that are not implemented are rejected with an
// error (exit code 255 like the reference's `return -1`), never ignored.
#include <getopt.h>

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>

#include "locarna_b200.hh"

using namespace LocARNA_B200;

namespace {
enum { O_INDEL_OPENING = 1000, O_RIBOSUM_FILE, O_USE_RIBOSUM, O_TEMPERATURE, O_PF_SCALE, O_WRITE_AM, O_WRITE_BM, O_MAX_DIFF_AT_AM, O_MIN_TRACE_PROB,
       O_INCLUDE_AM_IN_BM, O_MAXBPSPAN, O_MAX_BPS_LENGTH_RATIO, O_UNSUPPORTED, O_DEVICE };
bool parse_bool(const char *s) {
    const std::string v = s ? s : "";
    if (v == "t" || v == "true" || v == "on" || v == "1") return true;
    if (v == "f" || v == "false" || v == "off" || v == "0") return false;
    std::cerr << "ERROR: cannot parse boolean value \"" << v << "\"" << std::endl;
    exit(255);
}
}  // namespace

int main(int argc, char **argv) {
    static const struct option longopts[] = {
        {"indel", required_argument, 0, 'i'}, {"indel-opening", required_argument, 0, O_INDEL_OPENING},
        {"ribosum-file", required_argument, 0, O_RIBOSUM_FILE}, {"use-ribosum", required_argument, 0, O_USE_RIBOSUM},
        {"match", required_argument, 0, 'm'}, {"mismatch", required_argument, 0, 'M'}, {"struct-weight", required_argument, 0, 's'},
        {"tau", required_argument, 0, 't'}, {"temperature-alipf", required_argument, 0, O_TEMPERATURE}, {"pf-scale", required_argument, 0, O_PF_SCALE},
        {"min-am-prob", required_argument, 0, 'a'}, {"min-bm-prob", required_argument, 0, 'b'},
        {"write-arcmatch-probs", required_argument, 0, O_WRITE_AM}, {"write-basematch-probs", required_argument, 0, O_WRITE_BM},
        {"include-am-in-bm", no_argument, 0, O_INCLUDE_AM_IN_BM},
        {"min-prob", required_argument, 0, 'p'}, {"max-diff-am", required_argument, 0, 'D'}, {"max-diff", required_argument, 0, 'd'},
        {"max-diff-at-am", required_argument, 0, O_MAX_DIFF_AT_AM}, {"min-trace-probability", required_argument, 0, O_MIN_TRACE_PROB},
        // recognised, not implemented on the B200 path
        {"extended-pf", no_argument, 0, O_UNSUPPORTED}, {"quad-pf", no_argument, 0, O_UNSUPPORTED}, {"exp-prob", required_argument, 0, 'e'},
        {"max-diff-aln", required_argument, 0, O_UNSUPPORTED}, {"max-diff-pw-aln", required_argument, 0, O_UNSUPPORTED},
        {"max-diff-relax", no_argument, 0, O_UNSUPPORTED}, {"fragment-match-probs", required_argument, 0, O_UNSUPPORTED},
        {"relaxed-anchors", no_argument, 0, O_UNSUPPORTED}, {"maxBPspan", required_argument, 0, O_MAXBPSPAN}, {"ribofit", required_argument, 0, O_UNSUPPORTED},
        {"max-bps-length-ratio", required_argument, 0, O_MAX_BPS_LENGTH_RATIO},
        {"device", required_argument, 0, O_DEVICE}, {"quiet", no_argument, 0, 'q'}, {"verbose", no_argument, 0, 'v'}, {"stopwatch", no_argument, 0, 'v'},
        {"version", no_argument, 0, 'V'}, {"help", no_argument, 0, 'h'}, {0, 0, 0, 0}};
    ScoringParams sp;
    AlignerPParams ap;
    int device = 0;
    double min_prob = 0.001;
    int max_bp_span = -1;
    double max_bps_length_ratio = 0.0;
    bool quiet = false, verbose = false, include_am_in_bm = false;
    std::string am_file, bm_file;
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "i:m:M:s:e:t:a:b:p:D:d:qvVh", longopts, &idx)) != -1) {
        switch (c) {
            case 'i': sp.indel = atoi(optarg); break;
            case O_INDEL_OPENING: sp.indel_opening = atoi(optarg); break;
            case O_RIBOSUM_FILE: sp.ribosum_file = optarg; break;   // read when the aligner is set up (main_helper.icc:311-350)
            case O_USE_RIBOSUM: sp.use_ribosum = parse_bool(optarg); break;
            case 'm': sp.match = atoi(optarg); break;
            case 'M': sp.mismatch = atoi(optarg); break;
            case 's': sp.struct_weight = atoi(optarg); break;
            case 't': sp.tau_factor = atoi(optarg); break;
            case 'e': sp.exp_prob = atof(optarg); break;
            case O_MAXBPSPAN: max_bp_span = atoi(optarg); break;
            case O_MAX_BPS_LENGTH_RATIO: max_bps_length_ratio = atof(optarg); break;
            case O_TEMPERATURE: sp.temperature_alipf = atoi(optarg); break;
            case O_PF_SCALE: ap.pf_scale(atof(optarg)); break;
            case 'a': ap.min_am_prob(atof(optarg)); break;
            case 'b': ap.min_bm_prob(atof(optarg)); break;
            case O_WRITE_AM: am_file = optarg; break;
            case O_WRITE_BM: bm_file = optarg; break;
            case O_INCLUDE_AM_IN_BM: include_am_in_bm = true; break;
            case 'p': min_prob = atof(optarg); break;
            case 'D': ap.max_diff_am(atoi(optarg)); break;
            case 'd': ap.max_diff(atoi(optarg)); break;
            case O_MAX_DIFF_AT_AM: ap.max_diff_at_am(atoi(optarg)); break;
            case O_MIN_TRACE_PROB: ap.min_trace_probability(atof(optarg)); break;
            case O_UNSUPPORTED:
                std::cerr << "ERROR: option --" << (idx >= 0 && longopts[idx].name ? longopts[idx].name : "?")
                          << " selects a mode that locarna_p_b200 does not implement." << std::endl;
                return 255;
            case O_DEVICE: device = atoi(optarg); break;
            case 'q': quiet = true; break;
            case 'v': verbose = true; break;
            case 'V': std::cout << "locarna_p_b200 (LocARNA 2.0.1 LocARNA-P path, B200)" << std::endl; return 0;
            case 'h': std::cout << "usage: locarna_p_b200 [options as locarna_p] <fileA.pp> <fileB.pp>" << std::endl; return 0;
            default: return 255;
        }
    }
    if (argc - optind != 2) { std::cerr << "ERROR: expected two input files (PP 2.0)." << std::endl; return 255; }
    try {
        RnaData rnaA(argv[optind], min_prob, max_bps_length_ratio, max_bp_span), rnaB(argv[optind + 1], min_prob, max_bps_length_ratio, max_bp_span);
        ap.seqA(&rnaA).seqB(&rnaB).scoring(sp).min_prob(min_prob);
        AlignerP aligner(ap, device);
        if (verbose) std::cout << "Run inside algorithm." << std::endl;
        const double pf = aligner.align_inside();
        if (!quiet) std::cout << "Partition function: " << pf << std::endl;            // locarna_p.cc:483-487
        if (verbose) std::cout << "Run outside algorithm." << std::endl;
        aligner.align_outside();
        if (verbose) std::cout << "Compute probabilities." << std::endl;
        aligner.compute_arcmatch_probabilities();
        if (!am_file.empty()) {                                                          // locarna_p.cc:501-514
            if (verbose) std::cout << "Write Arc-match probabilities to file " << am_file << "." << std::endl;
            std::ofstream out(am_file.c_str());
            if (out.good()) aligner.write_arcmatch_probabilities(out);
            else { std::cerr << "Cannot write to " << am_file << "! Exit." << std::endl; return 255; }
        }
        aligner.compute_basematch_probabilities(include_am_in_bm);
        if (!bm_file.empty()) {                                                          // locarna_p.cc:518-531
            if (verbose) std::cout << "Write Base-match probabilities to file " << bm_file << "." << std::endl;
            std::ofstream out(bm_file.c_str());
            if (out.good()) aligner.write_basematch_probabilities(out);
            else { std::cerr << "Cannot write to " << bm_file << "! Exit." << std::endl; return 255; }
        }
        return 0;
    } catch (failure &f) {
        std::cerr << "ERROR: " << f.what() << std::endl;
        return 255;
    }
}
