// mlocarna_tree_b200: the all-vs-all guide-tree stage of mlocarna as ONE process on one B200 (host C++ over the C ABI).
//
// Replaces the N(N-1)/2 `locarna inA inB @locarna_params_tree --clustal tmp -q` process calls of
// src/Utils/mlocarna:3547-3643 (compute_all_pairwise_alignments; pair (a, b) for a in 0..n-1, b in 0..a-1, A = the later
// sequence), the score extraction (:3516-3527, "-inf" -> -1e8), the symmetric score matrix with zero diagonal written as
// results/result.matrix ("%6d" columns, :2353-2373, lib/perl/MLocarna.pm:2034-2052) and the UPGMA guide tree written as
// results/result.tree (:2381-2388, lib/perl/MLocarna/Tree.pm:181-322). Inputs are the PP 2.0 files mlocarna keeps in
// <tgtdir>/input/ (one per sequence, given in the order of the input sequences); each is parsed once instead of 2(N-1) times.
//
// Hand-over to stock mlocarna (no Perl changed): `mlocarna --similarity-matrix <matrix file>` (:2232-2234) or
// `mlocarna --treefile <tree file>` (:2223-2231), or the score list (`--score-list`, lines "<a> <b> <score>",
// lib/perl/MLocarna/SparseMatrix.pm:264-276) dropped into <tgtdir>/scores/ for `mlocarna --score-lists`.
//
// Several GPUs: `--gpus N` splits the pair list over the devices 0..N-1 of this box by estimated arc-match cost (lb200_shard_pairs,
// one host thread and one context per device, every PP file parsed once per device) and merges the scores on the host.
// `--compute-pairwise-scores k/N` is mlocarna's own way to distribute the stage over processes (src/Utils/mlocarna:2321-2344): this
// process computes share k of N (1-based) of the pair list, by the same cost-balanced split, and writes only the partial score list
// (`--score-list FILE`, or <tgtdir>/scores/scores-k); `mlocarna --score-lists` merges the lists.
//
// Defaults are the flags mlocarna passes in this stage (@locarna_params_tree: --struct-weight 200 --max-diff-am 30 --noLP
// --min-prob 0.001, mlocarna:1239-1242, :1433-1439, :1519-1594); every `locarna` scoring / heuristic flag of the path can be given
// to override them (--LP switches --noLP off, as mlocarna's --LP does).
#include <getopt.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "locarna_b200.h"

namespace {
enum {
    O_INDEL_OPENING = 1000, O_USE_RIBOSUM, O_RIBOSUM_FILE, O_UNPAIRED_PENALTY, O_STRUCT_LOCAL, O_SEQU_LOCAL, O_FREE_ENDGAPS, O_MAX_DIFF_AT_AM,
    O_MIN_TRACE_PROB, O_NOLP, O_LP, O_MAXBPSPAN, O_MAX_BPS_LENGTH_RATIO, O_TEMPERATURE_ALIPF, O_MATRIX, O_SCORE_LIST, O_TREE, O_TGTDIR, O_DEVICE, O_GPUS, O_SHARE
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
        {"indel", required_argument, 0, 'i'}, {"indel-opening", required_argument, 0, O_INDEL_OPENING}, {"use-ribosum", required_argument, 0, O_USE_RIBOSUM}, {"ribosum-file", required_argument, 0, O_RIBOSUM_FILE},
        {"match", required_argument, 0, 'm'}, {"mismatch", required_argument, 0, 'M'}, {"unpaired-penalty", required_argument, 0, O_UNPAIRED_PENALTY},
        {"struct-weight", required_argument, 0, 's'}, {"exp-prob", required_argument, 0, 'e'}, {"tau", required_argument, 0, 't'},
        {"exclusion", required_argument, 0, 'E'}, {"struct-local", required_argument, 0, O_STRUCT_LOCAL}, {"sequ-local", required_argument, 0, O_SEQU_LOCAL},
        {"free-endgaps", required_argument, 0, O_FREE_ENDGAPS}, {"min-prob", required_argument, 0, 'p'}, {"max-diff-am", required_argument, 0, 'D'},
        {"max-diff", required_argument, 0, 'd'}, {"max-diff-at-am", required_argument, 0, O_MAX_DIFF_AT_AM},
        {"min-trace-probability", required_argument, 0, O_MIN_TRACE_PROB}, {"noLP", no_argument, 0, O_NOLP}, {"LP", no_argument, 0, O_LP},
        {"maxBPspan", required_argument, 0, O_MAXBPSPAN}, {"max-bps-length-ratio", required_argument, 0, O_MAX_BPS_LENGTH_RATIO}, {"temperature-alipf", required_argument, 0, O_TEMPERATURE_ALIPF},
        {"matrix", required_argument, 0, O_MATRIX}, {"score-list", required_argument, 0, O_SCORE_LIST}, {"tree", required_argument, 0, O_TREE},
        {"tgtdir", required_argument, 0, O_TGTDIR}, {"device", required_argument, 0, O_DEVICE}, {"gpus", required_argument, 0, O_GPUS},
        {"compute-pairwise-scores", required_argument, 0, O_SHARE}, {"quiet", no_argument, 0, 'q'},
        {"verbose", no_argument, 0, 'v'}, {"help", no_argument, 0, 'h'}, {0, 0, 0, 0}};
    lb200_params p;
    lb200_default_params(&p);
    p.struct_weight = 200; p.max_diff_am = 30; p.no_lonely_pairs = 1; p.min_prob = 0.001;   // @locarna_params_tree
    std::string matrix_file, list_file, tree_file, tgtdir, ribosum_file;
    int device = 0, gpus = 1, share_k = 0, share_n = 0;
    bool quiet = false, verbose = false;
    int c, idx = 0;
    while ((c = getopt_long(argc, argv, "i:m:M:s:e:t:E:p:D:d:qvh", longopts, &idx)) != -1) {
        switch (c) {
            case 'i': p.indel = atoi(optarg); break;
            case O_INDEL_OPENING: p.indel_opening = atoi(optarg); break;
            case O_USE_RIBOSUM: p.use_ribosum = parse_bool(optarg); break;
            case O_RIBOSUM_FILE: ribosum_file = optarg; break;
            case 'm': p.match = atoi(optarg); break;
            case 'M': p.mismatch = atoi(optarg); break;
            case O_UNPAIRED_PENALTY: p.unpaired_penalty = atoi(optarg); break;
            case 's': p.struct_weight = atoi(optarg); break;
            case 'e': p.exp_prob = atof(optarg); break;
            case 't': p.tau = atoi(optarg); break;
            case 'E': p.exclusion = atoi(optarg); break;
            case O_STRUCT_LOCAL: p.struct_local = parse_bool(optarg); break;
            case O_SEQU_LOCAL: p.sequ_local = parse_bool(optarg); break;
            case O_FREE_ENDGAPS: strncpy(p.free_endgaps, optarg, sizeof(p.free_endgaps) - 1); break;
            case 'p': p.min_prob = atof(optarg); break;
            case 'D': p.max_diff_am = atoi(optarg); break;
            case 'd': p.max_diff = atoi(optarg); break;
            case O_MAX_DIFF_AT_AM: p.max_diff_at_am = atoi(optarg); break;
            case O_MIN_TRACE_PROB: p.min_trace_probability = atof(optarg); break;
            case O_NOLP: p.no_lonely_pairs = 1; break;
            case O_LP: p.no_lonely_pairs = 0; break;
            case O_MAXBPSPAN: p.max_bp_span = atoi(optarg); break;
            case O_MAX_BPS_LENGTH_RATIO: p.max_bps_length_ratio = atof(optarg); break;
            case O_TEMPERATURE_ALIPF: p.temperature_alipf = atoi(optarg); break;
            case O_MATRIX: matrix_file = optarg; break;
            case O_SCORE_LIST: list_file = optarg; break;
            case O_TREE: tree_file = optarg; break;
            case O_TGTDIR: tgtdir = optarg; break;
            case O_DEVICE: device = atoi(optarg); break;
            case O_GPUS: gpus = atoi(optarg); break;
            case O_SHARE:
                if (sscanf(optarg, "%d/%d", &share_k, &share_n) != 2 || share_n < 1 || share_k < 1 || share_k > share_n) {
                    std::cerr << "ERROR: --compute-pairwise-scores expects k/N with 1 <= k <= N" << std::endl;
                    return 255;
                }
                break;
            case 'q': quiet = true; break;
            case 'v': verbose = true; break;
            case 'h':
                std::cout << "usage: mlocarna_tree_b200 [locarna scoring/heuristic options] [--gpus N] [--compute-pairwise-scores k/N] [--tgtdir DIR | --matrix FILE --tree FILE --score-list FILE]"
                             " <seq0.pp> <seq1.pp> ...   (PP 2.0 files in input-sequence order)" << std::endl;
                return 0;
            default: return 255;
        }
    }
    const int n = argc - optind;
    if (n < 2) { std::cerr << "ERROR: expected at least two input files (PP 2.0)." << std::endl; return 255; }
    if (!tgtdir.empty()) {  // the files mlocarna itself writes / reads under its target directory (directories must exist, as after mlocarna's set-up)
        if (matrix_file.empty()) matrix_file = tgtdir + "/results/result.matrix";
        if (tree_file.empty()) tree_file = tgtdir + "/results/result.tree";
        if (list_file.empty()) list_file = tgtdir + "/scores/scores-0";
    }

    if (gpus < 1) { std::cerr << "ERROR: --gpus expects a positive number" << std::endl; return 255; }
    if (!tgtdir.empty() && share_n > 0 && list_file == tgtdir + "/scores/scores-0") list_file = tgtdir + "/scores/scores-" + std::to_string(share_k);

    // pair list in mlocarna's order (mlocarna:3577-3604: A = the later sequence)
    const int64_t n_pairs = lb200_all_vs_all(n, nullptr, nullptr);
    std::vector<int> pa((size_t)n_pairs), pb((size_t)n_pairs);
    lb200_all_vs_all(n, pa.data(), pb.data());
    std::vector<std::pair<int, int>> pairs((size_t)n_pairs);
    for (int64_t k = 0; k < n_pairs; k++) pairs[k] = std::make_pair(pa[k], pb[k]);
    std::vector<const char *> files(n);
    for (int k = 0; k < n; k++) files[k] = argv[optind + k];

    // one context per device; the first one also yields the names and the cost estimate of every pair
    const int n_ctx = gpus;
    std::vector<lb200_ctx *> ctxs(n_ctx, nullptr);
    std::vector<std::string> names(n);
    std::vector<std::string> errors(n_ctx);
    auto open_ctx = [&](int g) -> bool {
        if (lb200_ctx_create(device + g, &ctxs[g]) < 0) { errors[g] = "cannot create the device context"; return false; }
        // every .pp file is parsed once (by the first context, on all host cores); the other devices' contexts copy the result
        if ((!ribosum_file.empty() && lb200_set_ribosum_file(ctxs[g], ribosum_file.c_str()) < 0) || lb200_set_params(ctxs[g], &p) < 0 || (g == 0 ? lb200_seqs_add_pp(ctxs[g], n, files.data()) : lb200_seqs_copy(ctxs[g], ctxs[0])) < 0) {
            errors[g] = lb200_last_error(ctxs[g]);
            return false;
        }
        return true;
    };
    if (!open_ctx(0)) { std::cerr << "ERROR: " << errors[0] << std::endl; return 255; }
    std::vector<double> cost((size_t)n_pairs);
    for (int k = 0; k < n; k++) {
        char name[256];
        lb200_seq_get(ctxs[0], k, name, sizeof name, nullptr, 0);
        names[k] = name;
    }
    for (int64_t k = 0; k < n_pairs; k++)
        cost[k] = lb200_pair_cost(lb200_seq_num_arcs(ctxs[0], pa[k]), lb200_seq_num_arcs(ctxs[0], pb[k]), lb200_seq_length(ctxs[0], pa[k]), lb200_seq_length(ctxs[0], pb[k]));

    // this process's share of the pair list (--compute-pairwise-scores k/N), then its split over the devices
    std::vector<int64_t> mine;
    if (share_n > 0) {
        std::vector<int> rank_of((size_t)n_pairs);
        if (lb200_shard_pairs(n_pairs, cost.data(), share_n, rank_of.data(), nullptr, nullptr) < 0) { std::cerr << "ERROR: sharding failed" << std::endl; return 255; }
        for (int64_t k = 0; k < n_pairs; k++) if (rank_of[k] == share_k - 1) mine.push_back(k);
    } else {
        mine.resize((size_t)n_pairs);
        for (int64_t k = 0; k < n_pairs; k++) mine[k] = k;
    }
    std::vector<double> my_cost(mine.size());
    for (size_t k = 0; k < mine.size(); k++) my_cost[k] = cost[mine[k]];
    std::vector<int> dev_of(mine.size());
    std::vector<int64_t> order(mine.size()), begin((size_t)n_ctx + 1);
    if (lb200_shard_pairs((int64_t)mine.size(), my_cost.data(), n_ctx, dev_of.data(), order.data(), begin.data()) < 0) { std::cerr << "ERROR: sharding failed" << std::endl; return 255; }
    if (verbose) std::cerr << "aligning " << mine.size() << " of " << n_pairs << " pairs of " << n << " sequences on " << n_ctx << " GPU(s)" << std::endl;

    std::vector<int64_t> scores((size_t)n_pairs, 0);
    std::vector<char> have((size_t)n_pairs, 0);
    auto work = [&](int g) {
        if (g > 0 && !open_ctx(g)) return;
        lb200_ctx *ctx = ctxs[g];
        const int64_t cnt = begin[g + 1] - begin[g];
        std::vector<int> a((size_t)cnt), b((size_t)cnt);
        for (int64_t k = 0; k < cnt; k++) { const int64_t gp = mine[order[begin[g] + k]]; a[k] = pa[gp]; b[k] = pb[gp]; }
        std::vector<int64_t> sc((size_t)cnt);
        if (lb200_pairs_add(ctx, (int)cnt, a.data(), b.data()) < 0 || lb200_run(ctx, LB200_RUN_SCORE_ONLY) < 0 ||
            (cnt > 0 && lb200_get_scores(ctx, sc.data(), (int)cnt) < 0)) { errors[g] = lb200_last_error(ctx); return; }
        for (int64_t k = 0; k < cnt; k++) { const int64_t gp = mine[order[begin[g] + k]]; scores[gp] = sc[k]; have[gp] = 1; }
        if (verbose) std::cerr << "device " << device + g << ": " << cnt << " pairs, device time " << lb200_last_kernel_ms(ctx) << " ms, " << lb200_last_launches(ctx) << " kernel launches" << std::endl;
    };
    {
        std::vector<std::thread> pool;
        for (int g = 1; g < n_ctx; g++) pool.emplace_back(work, g);
        work(0);
        for (auto &t : pool) t.join();
    }
    bool failed = false;
    for (int g = 0; g < n_ctx; g++) {
        if (!errors[g].empty()) { std::cerr << "ERROR: device " << device + g << ": " << errors[g] << std::endl; failed = true; }
        if (ctxs[g]) lb200_ctx_destroy(ctxs[g]);
    }
    if (failed) return 255;

    if (share_n > 0) {   // partial score list only (merged by `mlocarna --score-lists`)
        std::ostream *out = &std::cout;
        std::ofstream f;
        if (!list_file.empty()) { f.open(list_file.c_str()); if (!f.good()) { std::cerr << "ERROR: Cannot write to " << list_file << "." << std::endl; return 255; } out = &f; }
        for (int64_t k : mine) {
            *out << pairs[k].first << " " << pairs[k].second << " ";
            if (scores[k] == LB200_SCORE_NEG_INF) *out << "-inf"; else *out << (long long)scores[k];
            *out << "\n";
        }
        return 0;
    }

    std::vector<int64_t> matrix((size_t)n * n, 0);
    for (size_t k = 0; k < pairs.size(); k++) {
        const int64_t v = scores[k] == LB200_SCORE_NEG_INF ? -100000000LL : scores[k];   // mlocarna:3523
        matrix[(size_t)pairs[k].first * n + pairs[k].second] = v;
        matrix[(size_t)pairs[k].second * n + pairs[k].first] = v;
    }
    std::vector<const char *> cnames(n);
    for (int k = 0; k < n; k++) cnames[k] = names[k].c_str();
    std::vector<char> newick((size_t)n * 320 + 64);
    if (lb200_upgma_newick(n, cnames.data(), matrix.data(), newick.data(), newick.size()) < 0) {
        std::cerr << "ERROR: guide tree construction failed" << std::endl;
        return 255;
    }

    int rc = 0;
    auto write_matrix = [&](std::ostream &out) {
        char buf[32];
        for (int a = 0; a < n; a++) {
            for (int b = 0; b < n; b++) { snprintf(buf, sizeof buf, "%6lld", (long long)matrix[(size_t)a * n + b]); out << (b ? " " : "") << buf; }
            out << "\n";
        }
    };
    if (!matrix_file.empty()) {
        std::ofstream out(matrix_file.c_str());
        if (out.good()) write_matrix(out); else { std::cerr << "ERROR: Cannot write to " << matrix_file << "." << std::endl; rc = 255; }
    }
    if (!list_file.empty()) {
        std::ofstream out(list_file.c_str());
        if (out.good()) {
            for (size_t k = 0; k < pairs.size(); k++) {
                out << pairs[k].first << " " << pairs[k].second << " ";
                if (scores[k] == LB200_SCORE_NEG_INF) out << "-inf"; else out << (long long)scores[k];
                out << "\n";
            }
        } else { std::cerr << "ERROR: Cannot write to " << list_file << "." << std::endl; rc = 255; }
    }
    if (!tree_file.empty()) {
        std::ofstream out(tree_file.c_str());
        if (out.good()) out << newick.data() << ";\n"; else { std::cerr << "ERROR: Cannot write to " << tree_file << "." << std::endl; rc = 255; }
    }
    if (!quiet && matrix_file.empty() && tree_file.empty() && list_file.empty()) { write_matrix(std::cout); std::cout << newick.data() << ";" << std::endl; }
    return rc;
}
