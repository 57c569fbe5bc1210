// Context, device memory management, scheduling and the C ABI (include/locarna_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <map>
#include <memory>
#include <vector>

#include "../../include/locarna_b200.h"
#include "builder.h"
#include "envelope.h"
#include "dev_ctx.h"
#include "host_model.h"
#include "rows.h"

namespace lb200 {
void launch_dfill(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int q, cudaStream_t st);
void launch_toplevel(const DevCtx &c, int ncmax, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st);
void launch_trace(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st);
cudaError_t configure_kernels(int ncmax, int smem_bytes, int *dfill_ctas_per_sm);
void launch_dfill_dep(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int *cursor, cudaStream_t st);
void launch_dfill_sl(const DevCtx &c, int grid, int smem_bytes, int q, cudaStream_t st);
cudaError_t configure_sl(int smem_bytes, int *ctas_per_sm);
struct PfCtx {   // must match pf_inside.cuh
    const double *esig; const double *bpow;
    double g, open, inv_scale, pf_scale, temp;
    double *dpf; double *ztop; double *scratch;
    long long scratch_dwords;
    int acc_doubles;
};
void launch_pfill(const DevCtx &c, const PfCtx &pc, int ncmax, int grid, int smem_bytes, int q, cudaStream_t st);
void launch_ptop(const DevCtx &c, const PfCtx &pc, int ncmax, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st);
cudaError_t configure_pf(int ncmax, int smem_bytes, int *ctas_per_sm);
struct PfoCtx {   // must match pf_outside.cuh
    PfCtx pc;
    const int *cell_start; const int *cell_rev;
    const int *arc_left, *arc_right;
    const int *lptr, *lcount;
    double *dpp; double *amp; double *mats; double *cta;
    long long mat_doubles;
    int acc_cols;
    double am_threshold;
};
void launch_pfo_prepare(const DevCtx &c, const PfoCtx &o, int n_pairs, int grid, cudaStream_t st);
void launch_pfo_outside(const DevCtx &c, const PfoCtx &o, int q, int grid, int *cursor, cudaStream_t st);
void launch_pfo_amprob(const DevCtx &c, const PfoCtx &o, int n_pairs, cudaStream_t st);
void launch_pfo_bm(const DevCtx &c, const PfoCtx &o, int n_tasks, int n_pairs, int grid, int *cursor, cudaStream_t st);
void launch_trace_sl(const DevCtx &c, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st);
}  // namespace lb200

using namespace lb200;

cudaError_t lb200_reset_d(DevEntry *ent, uint2 *ent8, size_t n, cudaStream_t st);

namespace {

// Device allocations are recycled across contexts of the process: cudaMalloc / cudaFree of the multi-GB tables cost more than a
// second per all-vs-all job (measured: context teardown 0.5-2.3 s, first build 0.33 s vs 1.0-1.3 s with fresh allocations), and a caller
// that aligns job after job creates a context per job. Blocks are handed out best-fit (at most twice the request), the cache is
// dropped when an allocation fails and by lb200_release_device_cache().
struct DevCache {
    struct Block { void *p; size_t bytes; int device; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t cached_bytes = 0;
    void *take(int device, size_t bytes, size_t *got) {
        std::lock_guard<std::mutex> lock(mu);
        int best = -1;
        for (int k = 0; k < (int)free_blocks.size(); k++) {
            const Block &b = free_blocks[k];
            if (b.device != device || b.bytes < bytes || b.bytes > 2 * bytes + (1 << 20)) continue;
            if (best < 0 || b.bytes < free_blocks[best].bytes) best = k;
        }
        if (best < 0) return nullptr;
        void *p = free_blocks[best].p;
        *got = free_blocks[best].bytes;
        cached_bytes -= free_blocks[best].bytes;
        free_blocks.erase(free_blocks.begin() + best);
        return p;
    }
    void give(int device, void *p, size_t bytes) {
        std::lock_guard<std::mutex> lock(mu);
        free_blocks.push_back(Block{p, bytes, device});
        cached_bytes += bytes;
    }
    void drop(int device) {   // device < 0: all devices
        std::lock_guard<std::mutex> lock(mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (size_t k = 0; k < free_blocks.size();) {
            if (device >= 0 && free_blocks[k].device != device) { k++; continue; }
            cudaSetDevice(free_blocks[k].device);
            cudaFree(free_blocks[k].p);
            cached_bytes -= free_blocks[k].bytes;
            free_blocks.erase(free_blocks.begin() + k);
        }
        cudaSetDevice(cur);
    }
};
DevCache g_dev_cache;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        int device = 0;
        cudaGetDevice(&device);
        // growth slack only for small buffers: a quarter more of a multi-GB scratch is what once exhausted the HBM
        const size_t want = bytes + (bytes < ((size_t)64 << 20) ? bytes / 4 : 0) + 256;
        size_t got = 0;
        if (void *q = g_dev_cache.take(device, want, &got)) { p = q; cap = got; dev = device; return cudaSuccess; }
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {   // give the cached blocks back to the driver and try once more
            cudaGetLastError();
            g_dev_cache.drop(device);
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) { cap = want; dev = device; } else p = nullptr;
        return e;
    }
    void release() { if (p) g_dev_cache.give(dev, p, cap); p = nullptr; cap = 0; }
    int dev = 0;
};

struct PairRec {
    int seqA, seqB;
    Band band;
    PairProblem prob;
    bool built = false;      // host mirror of the arc-match tables (host-only contexts / inspection without a GPU)
    bool banded = false;     // band available
    // device builder results
    long long am_base = 0; int K = 0;
    DevPairStats stats = {0, 0, 0, 0};
    // results
    int64_t score = 0; bool neg_inf = true; int max_i = 0, max_j = 0;
    std::vector<int> edges_a, edges_b; std::string str_a, str_b; bool traced = false;
    double pf_Z = 0; bool pf_done = false;   // LocARNA-P inside
    bool anchored = false;       // both sequences carry anchor names: band restricted and arc matches filtered by them
    bool band_initial = false;   // band given by the caller is the range BEFORE the probability envelope (reference alignment, anchors)
    bool restricted = false; int r_sa = 1, r_sb = 1, r_ea = 0, r_eb = 0;   // AlignerRestriction of the top level (k-best)
    std::shared_ptr<ProfileTables> prof;   // position-specific score tables (contexts with profile input)
};

}  // namespace

struct lb200_ctx {
    int device = 0;
    cudaDeviceProp prop;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
    std::string err;
    Params params;
    ScoreTables tables;
    bool params_locked = false;
    std::vector<Sequence> seqs;
    std::vector<PairRec> pairs;
    double last_kernel_ms = 0;
    int64_t last_launches = 0;
    int64_t last_h2d_bytes = 0, last_d2h_bytes = 0;
    double last_dfill_ms = 0;
    int64_t last_dfill_launches = 0;
    // batch resident in HBM (lb200_upload)
    struct Resident {
        bool valid = false;
        DevCtx dc;
        int p0 = 0, p1 = 0;  // pair range of the resident chunk
        int nc_inst = 1, smem_bytes = 0, grid_cap = 1, q_lo = 0, q_hi = 0;
        size_t total_am = 0;
        long long sptr_total = 0;
        int stack_cap = 0;
        std::vector<int> sptr_off;
        bool rows = false;       // D fill by the row-grouped kernel (dfill_rows.cu)
        RowsCtx rc;
        int rows_grid = 1, rows_smem = 0;
        unsigned n_groups = 0;
    } res;
    int host_threads = 0;
    bool d_filled = false;   // the resident chunk's D table is complete (Aligner's D_created_)
    RibosumTables ribosum;   // --ribosum-file (lb200_set_ribosum_file); default: the built-in matrix
    size_t seqs_uploaded = 0;  // sequences whose arrays are on the device
    std::vector<int> seq_codes_off, seq_arcs_off, seq_lptr_off;
    DevBuf d_arc_left, d_arc_right, d_arc_weight, d_arc_sdelta, d_lptr, d_lcount, d_am_seq, d_cell_rev, d_cell_start, d_skeys, d_skeys2, d_svals, d_svals2,
        d_tasks_unsorted, d_tkeys, d_tkeys2, d_tvals, d_tvals2, d_ntasks, d_qstart, d_stats, d_tmp, d_tr_edges, d_tr_str, d_tr_stack,
        d_pup, d_pdown, d_env_pairs, d_env_lo, d_env_hi, d_env_olo, d_env_ohi, d_env_flag, d_env_scratch;
    std::vector<int> seq_prob_off;
    int64_t env_device_pairs = 0, env_host_pairs = 0;  // statistics of the last band derivation
    int env_mode = 1;  // 1: device screening + host re-check, 0: host only (LB200_ENVELOPE=host)
    DevBuf d_pf_esig, d_pf_bpow, d_pf_d, d_pf_z, d_pf_scratch, d_pf_dp, d_pf_amp, d_pf_mats, d_pf_cta;
    PfCtx pf_last; bool pf_have = false; bool pf_probs_done = false; long long pf_mat_doubles = 0;
    int max_box_words = 1, max_len = 1;
    DevBuf d_acodes, d_pairs, d_codes, d_band_lo, d_band_hi, d_sptr, d_ent, d_am, d_tasks, d_top, d_scratch, d_cursor, d_flag, d_levcnt, d_done, d_ent8;
    DevBuf d_ent_mod, d_ent8_mod;   // D view of the modified scoring (normalized / penalized alignment)
    DevBuf d_ps_sig, d_ps_am;       // profile pairs: sigma'(i, j) tables and arc-match sequence terms (host_model.h ProfileTables)
    // profile mode: some sequence of the context is an alignment (several rows). All pairs are then scored position by position
    // (which for two single sequences gives the same numbers as the tables by symbol code).
    bool profile_mode() const { for (const Sequence &s : seqs) if (s.num_rows() > 1) return true; return false; }
    DevBuf d_col_first, d_col_last, d_groups, d_gorder, d_ngroups, d_rows_scratch, d_row_built, d_clist, d_cnblk;   // row-grouped D fill
    int pack_entries = 1;  // 8-byte packed S-order copy for the single-state sweep when every sequence is <= LB_PACK_MAXLEN (LB200_PACK=0 disables)
    int dfill_mode = 2;   // 2: automatic (row-grouped kernel where it applies), 3: row-grouped or fail (LB200_DFILL=rows), 1: dependency-driven persistent launch of single boxes (LB200_DFILL=dep), 0: one launch per level group (LB200_DFILL=levels)
    int last_dfill_kind = 0;      // D-fill kernel of the last run: 0 one launch per level group, 1 dependency-driven boxes, 2 row-grouped
    int64_t rows_fallbacks = 0;   // chunks that were re-run box by box because the row-grouped kernel met an unsupported box
    int sb_pairs = 1 << 30;   // pair block of the dependency-driven order (LB200_SB_PAIRS; default: one block)
    ~lb200_ctx() {
        DevBuf *all[] = {&d_ps_sig, &d_ps_am, &d_acodes, &d_pairs, &d_codes, &d_band_lo, &d_band_hi, &d_sptr, &d_ent, &d_am, &d_tasks, &d_top, &d_scratch, &d_cursor, &d_flag, &d_levcnt, &d_done, &d_ent8,
                         &d_ent_mod, &d_ent8_mod, &d_col_first, &d_col_last, &d_groups, &d_gorder, &d_ngroups, &d_rows_scratch, &d_row_built, &d_clist, &d_cnblk,
                         &d_arc_left, &d_arc_right, &d_arc_weight, &d_arc_sdelta, &d_lptr, &d_lcount, &d_am_seq, &d_cell_rev, &d_cell_start, &d_skeys, &d_skeys2,
                         &d_svals, &d_svals2, &d_tasks_unsorted, &d_tkeys, &d_tkeys2, &d_tvals, &d_tvals2, &d_ntasks, &d_qstart, &d_stats, &d_tmp, &d_tr_edges, &d_tr_str, &d_tr_stack,
                         &d_pf_esig, &d_pf_bpow, &d_pf_d, &d_pf_z, &d_pf_scratch, &d_pf_dp, &d_pf_amp, &d_pf_mats, &d_pf_cta,
                         &d_pup, &d_pdown, &d_env_pairs, &d_env_lo, &d_env_hi, &d_env_olo, &d_env_ohi, &d_env_flag, &d_env_scratch};
        for (auto *b : all) b->release();
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_mid) cudaEventDestroy(ev_mid);
        if (stream) cudaStreamDestroy(stream);
    }
};

static int fail(lb200_ctx *c, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CUDA_TRY(c, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t e__ = (call);                                                                           \
        if (e__ != cudaSuccess) return fail(c, LB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

static Params to_params(const lb200_params &p) {
    Params q;
    q.min_prob = p.min_prob; q.max_diff_am = p.max_diff_am; q.max_diff_at_am = p.max_diff_at_am; q.max_diff = p.max_diff;
    q.min_trace_probability = p.min_trace_probability;
    q.no_lonely_pairs = p.no_lonely_pairs != 0; q.struct_local = p.struct_local != 0; q.sequ_local = p.sequ_local != 0;
    const size_t fl = strnlen(p.free_endgaps, sizeof p.free_endgaps);
    if (fl >= 4) {  // free_endgaps.hh:27-32: fewer than 4 characters means no free end gaps
        q.fe_left1 = p.free_endgaps[0] == '+'; q.fe_right1 = p.free_endgaps[1] == '+';
        q.fe_left2 = p.free_endgaps[2] == '+'; q.fe_right2 = p.free_endgaps[3] == '+';
    }
    q.struct_weight = p.struct_weight; q.indel = p.indel; q.indel_opening = p.indel_opening; q.tau = p.tau; q.exclusion = p.exclusion;
    q.match = p.match; q.mismatch = p.mismatch; q.unpaired_penalty = p.unpaired_penalty; q.temperature_alipf = p.temperature_alipf;
    q.use_ribosum = p.use_ribosum != 0; q.pf_double = p.pf_double != 0;
    q.exp_prob = p.exp_prob; q.max_bp_span = p.max_bp_span; q.max_bps_length_ratio = p.max_bps_length_ratio;
    q.stacking = p.stacking != 0; q.new_stacking = p.new_stacking != 0;
    return q;
}

extern "C" {

void lb200_default_params(lb200_params *p) {
    memset(p, 0, sizeof *p);
    p->min_prob = 0.001; p->max_diff_am = -1; p->max_diff_at_am = -1; p->max_diff = -1; p->min_trace_probability = 1e-4;
    p->struct_weight = 200; p->indel = -150; p->indel_opening = -750; p->tau = 50; p->exclusion = 0;
    p->match = 50; p->mismatch = 0; p->use_ribosum = 1; p->unpaired_penalty = 0; p->temperature_alipf = 300;
    strcpy(p->free_endgaps, "----");
    p->exp_prob = -1.0; p->max_bp_span = -1;
}

// cudaGetDeviceProperties takes milliseconds; a job creates a context per run (all-vs-all stage), so ask once per device and process
static cudaError_t cached_device_properties(cudaDeviceProp *out, int device) {
    static std::mutex mu;
    static std::map<int, cudaDeviceProp> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(device);
    if (it == cache.end()) {
        cudaDeviceProp p;
        const cudaError_t e = cudaGetDeviceProperties(&p, device);
        if (e != cudaSuccess) return e;
        it = cache.emplace(device, p).first;
    }
    *out = it->second;
    return cudaSuccess;
}

int lb200_ctx_create(int device, lb200_ctx **out) {
    if (!out) return LB200_ERR_ARG;
    *out = nullptr;
    if (device == LB200_DEVICE_NONE) {  // host-only context: lb200_prepare + inspection, lb200_run refuses
        lb200_ctx *c = new lb200_ctx();
        c->device = LB200_DEVICE_NONE;
        lb200_params dp;
        lb200_default_params(&dp);
        c->params = to_params(dp);
        make_score_tables(c->params, c->tables);
        if (const char *s = getenv("LB200_HOST_THREADS")) c->host_threads = atoi(s);
        *out = c;
        return LB200_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        fprintf(stderr, "locarna_b200: no CUDA device available (%s); this library has no CPU fallback\n",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return LB200_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return LB200_ERR_ARG;
    lb200_ctx *c = new lb200_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cached_device_properties(&c->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess ||
        cudaEventCreate(&c->ev1) != cudaSuccess || cudaEventCreate(&c->ev_mid) != cudaSuccess) {
        fprintf(stderr, "locarna_b200: cannot initialise device %d: %s\n", device, cudaGetErrorString(cudaGetLastError()));
        delete c;
        return LB200_ERR_CUDA;
    }
    lb200_params dp;
    lb200_default_params(&dp);
    c->params = to_params(dp);
    make_score_tables(c->params, c->tables);
    if (const char *s = getenv("LB200_HOST_THREADS")) c->host_threads = atoi(s);
    if (const char *s = getenv("LB200_ENVELOPE")) c->env_mode = strcmp(s, "host") == 0 ? 0 : 1;
    if (const char *s = getenv("LB200_DFILL")) c->dfill_mode = strcmp(s, "levels") == 0 ? 0 : strcmp(s, "dep") == 0 ? 1 : strcmp(s, "rows") == 0 ? 3 : 2;
    if (const char *s = getenv("LB200_SB_PAIRS")) c->sb_pairs = std::max(1, atoi(s));
    if (const char *s = getenv("LB200_PACK")) c->pack_entries = atoi(s) != 0;
    *out = c;
    return LB200_OK;
}

void lb200_ctx_destroy(lb200_ctx *c) {
    if (c && c->stream) cudaStreamSynchronize(c->stream);   // its buffers go back to the process-wide cache
    delete c;
}
void lb200_release_device_cache(void) { g_dev_cache.drop(-1); }
const char *lb200_last_error(const lb200_ctx *c) { return c ? c->err.c_str() : "null context"; }

int lb200_set_params(lb200_ctx *c, const lb200_params *p) {
    if (!c || !p) return LB200_ERR_ARG;
    if (!c->seqs.empty()) return fail(c, LB200_ERR_STATE, "parameters must be set before sequences are added");
    c->params = to_params(*p);
    c->params.ribosum = c->ribosum;
    make_score_tables(c->params, c->tables);
    return LB200_OK;
}

int lb200_seq_add_pp(lb200_ctx *c, const char *path) {
    if (!c || !path) return LB200_ERR_ARG;
    Sequence s;
    std::string err;
    if (!read_pp(path, c->params.min_prob, s, err, c->params.max_bp_span, c->params.max_bps_length_ratio, c->params.stacking || c->params.new_stacking)) return fail(c, LB200_ERR_IO, "%s", err.c_str());
    finish_sequence(s, c->params.min_prob);
    c->seqs.push_back(std::move(s));
    return (int)c->seqs.size() - 1;
}

// parallel_for / parallel_blocks are defined below
static void parallel_for(int n, int threads, const std::function<void(int)> &fn);
static void parallel_blocks(int n, int threads, const std::function<void(int, int)> &fn);

int lb200_seqs_add_pp(lb200_ctx *c, int n, const char *const *paths) {
    if (!c || n < 0 || (n > 0 && !paths)) return LB200_ERR_ARG;
    for (int k = 0; k < n; k++) if (!paths[k]) return LB200_ERR_ARG;
    std::vector<Sequence> seqs((size_t)n);
    std::vector<std::string> errs((size_t)n);
    std::vector<char> ok((size_t)n, 0);
    parallel_for(n, c->host_threads, [&](int k) {
        if (!read_pp(paths[k], c->params.min_prob, seqs[k], errs[k], c->params.max_bp_span, c->params.max_bps_length_ratio, c->params.stacking || c->params.new_stacking)) return;
        finish_sequence(seqs[k], c->params.min_prob);
        ok[k] = 1;
    });
    for (int k = 0; k < n; k++) if (!ok[k]) return fail(c, LB200_ERR_IO, "%s", errs[k].c_str());
    const int first = (int)c->seqs.size();
    for (int k = 0; k < n; k++) c->seqs.push_back(std::move(seqs[k]));
    return first;
}

// the parsed sequences of another context (same input filters): contexts that work off one job side by side parse every file once
int lb200_seqs_copy(lb200_ctx *c, const lb200_ctx *src) {
    if (!c || !src || c == src) return LB200_ERR_ARG;
    if (c->params.min_prob != src->params.min_prob || c->params.max_bp_span != src->params.max_bp_span ||
        c->params.max_bps_length_ratio != src->params.max_bps_length_ratio || (c->params.stacking || c->params.new_stacking) != (src->params.stacking || src->params.new_stacking))
        return fail(c, LB200_ERR_ARG, "lb200_seqs_copy: the two contexts filter their inputs differently");
    const int first = (int)c->seqs.size();
    c->seqs.insert(c->seqs.end(), src->seqs.begin(), src->seqs.end());
    return first;
}

// --ribosum-file (locarna.cc:86-88, main_helper.icc:311-350): score tables from a matrix file in the reference's extended ribosum format;
// NULL or "RIBOSUM85_60" selects the built-in matrix. Takes effect for tables built afterwards (call it before adding pairs).
int lb200_set_ribosum_file(lb200_ctx *c, const char *path) {
    if (!c) return LB200_ERR_ARG;
    RibosumTables t;
    if (path != nullptr && strcmp(path, "RIBOSUM85_60") != 0) {
        std::string err;
        if (!read_ribosum_file(path, t, err)) return fail(c, LB200_ERR_IO, "%s", err.c_str());
    }
    c->ribosum = t;
    c->params.ribosum = t;
    make_score_tables(c->params, c->tables);
    c->res.valid = false;
    return LB200_OK;
}

int lb200_seq_add(lb200_ctx *c, const char *name, const char *seq, const int *pi, const int *pj, const double *pp, int n) {
    if (!c || !seq || n < 0 || (n > 0 && (!pi || !pj || !pp))) return LB200_ERR_ARG;
    Sequence s;
    std::string err;
    if (!make_sequence(name ? name : "seq", seq, pi, pj, pp, n, c->params.min_prob, s, err, c->params.max_bp_span, c->params.max_bps_length_ratio)) return fail(c, LB200_ERR_ARG, "%s", err.c_str());
    finish_sequence(s, c->params.min_prob);
    c->seqs.push_back(std::move(s));
    return (int)c->seqs.size() - 1;
}

int lb200_seq_length(const lb200_ctx *c, int seq) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    return c->seqs[seq].len;
}

int lb200_seq_num_arcs(const lb200_ctx *c, int seq) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    return (int)c->seqs[seq].arcs.size();
}

// cost-balanced split of a pair list over `world` shares from the context's own sequences (lb200_pair_cost + lb200_shard_pairs in one call)
int lb200_shard_job(const lb200_ctx *c, int64_t n_pairs, const int *seqA, const int *seqB, int world, int *rank_of, int64_t *order, int64_t *rank_begin) {
    if (!c || n_pairs < 0 || world < 1 || (n_pairs > 0 && (!seqA || !seqB))) return LB200_ERR_ARG;
    const int ns = (int)c->seqs.size();
    std::vector<double> cost((size_t)n_pairs);
    for (int64_t k = 0; k < n_pairs; k++) {
        const int a = seqA[k], b = seqB[k];
        if (a < 0 || a >= ns || b < 0 || b >= ns) return LB200_ERR_ARG;
        cost[(size_t)k] = lb200_pair_cost((int)c->seqs[a].arcs.size(), (int)c->seqs[b].arcs.size(), c->seqs[a].len, c->seqs[b].len);
    }
    std::vector<int> tmp_rank;
    if (!rank_of) { tmp_rank.resize((size_t)std::max<int64_t>(n_pairs, 1)); rank_of = tmp_rank.data(); }
    return lb200_shard_pairs(n_pairs, cost.data(), world, rank_of, order, rank_begin);
}

// The base pairs the reader kept (RnaData::arc_prob > 0, rna_data.cc:1057-1093), sorted by (i, j): positions, probability, joint
// probability with the inner pair (0 = none). Arrays of lb200_seq_pairs(ctx, seq, NULL, ..) entries; returns the count. cutoff_out:
// RnaData::arc_cutoff_prob(); stacking_out: RnaData::has_stacking(). What a consensus dot plot (locarna --pp) is computed from.
int64_t lb200_seq_pairs(const lb200_ctx *c, int seq, int *pi, int *pj, double *pp, double *pp2, double *cutoff_out, int *stacking_out) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    const Sequence &s = c->seqs[seq];
    const size_t n = s.pp_i.size();
    for (size_t k = 0; k < n; k++) {
        if (pi) pi[k] = s.pp_i[k];
        if (pj) pj[k] = s.pp_j[k];
        if (pp) pp[k] = s.pp_p[k];
        if (pp2) pp2[k] = k < s.pp_p2.size() ? s.pp_p2[k] : 0.0;
    }
    if (cutoff_out) *cutoff_out = s.cutoff;
    if (stacking_out) *stacking_out = s.has_stacking ? 1 : 0;
    return (int64_t)n;
}

// anchor annotation of a sequence as SequenceAnnotation::single_string gives it (rows joined by '#'; "" = none); returns the length needed
int lb200_seq_anchors(const lb200_ctx *c, int seq, char *out, int cap) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    const Sequence &s = c->seqs[seq];
    std::string str;
    for (size_t k = 0; k < s.anchor_rows.size(); k++) { if (k) str += '#'; str += s.anchor_rows[k]; }
    if (out && cap > 0) { strncpy(out, str.c_str(), cap - 1); out[cap - 1] = 0; }
    return (int)str.size();
}

int lb200_seq_num_rows(const lb200_ctx *c, int seq) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    return c->seqs[seq].num_rows();
}

int lb200_seq_get_row(const lb200_ctx *c, int seq, int row, char *name, int name_cap, char *sequence, int sequence_cap) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    const Sequence &s = c->seqs[seq];
    if (row < 0 || row >= s.num_rows()) return LB200_ERR_ARG;
    const std::string &nm = s.rows.empty() ? s.name : s.row_names[(size_t)row];
    const std::string &str = s.row(row);
    if (name && name_cap > 0) { strncpy(name, nm.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (sequence) {
        if (sequence_cap < (int)str.size() + 1) return LB200_ERR_ARG;
        memcpy(sequence, str.c_str(), str.size() + 1);
    }
    return LB200_OK;
}

int lb200_seq_get(const lb200_ctx *c, int seq, char *name, int name_cap, char *sequence, int sequence_cap) {
    if (!c || seq < 0 || seq >= (int)c->seqs.size()) return LB200_ERR_ARG;
    if (name && name_cap > 0) { strncpy(name, c->seqs[seq].name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (sequence) {
        if (sequence_cap < (int)c->seqs[seq].seq.size() + 1) return LB200_ERR_ARG;
        memcpy(sequence, c->seqs[seq].seq.c_str(), c->seqs[seq].seq.size() + 1);
    }
    return LB200_OK;
}

static int pair_add(lb200_ctx *c, int seqA, int seqB, const int *min_col, const int *max_col, bool initial);
int lb200_pair_add(lb200_ctx *c, int seqA, int seqB, const int *min_col, const int *max_col) { return pair_add(c, seqA, seqB, min_col, max_col, false); }
int lb200_pair_add_restricted(lb200_ctx *c, int seqA, int seqB, const int *min_col, const int *max_col) {
    if (!min_col || !max_col) return LB200_ERR_ARG;
    return pair_add(c, seqA, seqB, min_col, max_col, true);
}
int lb200_band_from_alignment(int lenA, int lenB, const char *aliA, const char *aliB, int delta, int relaxed, int *min_col, int *max_col) {
    if (lenA < 0 || lenB < 0 || !aliA || !aliB || !min_col || !max_col) return LB200_ERR_ARG;
    Band b;
    std::string err;
    if (!band_from_alignment(lenA, lenB, aliA, aliB, delta, b, err, relaxed != 0)) { fprintf(stderr, "locarna_b200: %s\n", err.c_str()); return LB200_ERR_ARG; }
    std::copy(b.lo.begin(), b.lo.end(), min_col); std::copy(b.hi.begin(), b.hi.end(), max_col);
    return LB200_OK;
}
static int pair_add(lb200_ctx *c, int seqA, int seqB, const int *min_col, const int *max_col, bool initial) {
    if (!c || seqA < 0 || seqB < 0 || seqA >= (int)c->seqs.size() || seqB >= (int)c->seqs.size()) return LB200_ERR_ARG;
    if ((min_col == nullptr) != (max_col == nullptr)) return LB200_ERR_ARG;
    PairRec r;
    r.seqA = seqA; r.seqB = seqB;
    const int n = c->seqs[seqA].len, m = c->seqs[seqB].len;
    if (n < 1 || m < 1) return fail(c, LB200_ERR_UNSUPPORTED, "empty sequences are not supported");
    if (min_col) {
        r.band.lenA = n; r.band.lenB = m;
        r.band.lo.assign(min_col, min_col + n + 1); r.band.hi.assign(max_col, max_col + n + 1);
        // a band as TraceController leaves it (trace_controller.cc:522-531, :585-596): rows inside [0, m], non-empty, both
        // borders non-decreasing - the kernels rely on it
        for (int i = 0; i <= n; i++) {
            if (r.band.lo[i] < 0 || r.band.hi[i] > m) return fail(c, LB200_ERR_ARG, "band out of range in row %d", i);
            if (r.band.lo[i] > r.band.hi[i]) return fail(c, LB200_ERR_ARG, "empty band row %d", i);
            if (i > 0 && (r.band.lo[i] < r.band.lo[i - 1] || r.band.hi[i] < r.band.hi[i - 1])) return fail(c, LB200_ERR_ARG, "band is not monotone in row %d", i);
        }
        r.band_initial = initial;
    }
    c->pairs.push_back(std::move(r));
    c->res.valid = false;
    return (int)c->pairs.size() - 1;
}

int lb200_pairs_add(lb200_ctx *c, int n, const int *seqA, const int *seqB) {
    if (!c || n < 0 || (n > 0 && (!seqA || !seqB))) return LB200_ERR_ARG;
    for (int k = 0; k < n; k++) {
        if (seqA[k] < 0 || seqB[k] < 0 || seqA[k] >= (int)c->seqs.size() || seqB[k] >= (int)c->seqs.size()) return LB200_ERR_ARG;
        if (c->seqs[seqA[k]].len < 1 || c->seqs[seqB[k]].len < 1) return fail(c, LB200_ERR_UNSUPPORTED, "empty sequences are not supported");
    }
    const int first = (int)c->pairs.size();
    c->pairs.resize(c->pairs.size() + (size_t)n);
    for (int k = 0; k < n; k++) { c->pairs[first + k].seqA = seqA[k]; c->pairs[first + k].seqB = seqB[k]; }
    c->res.valid = false;
    return first;
}

int lb200_num_pairs(const lb200_ctx *c) { return c ? (int)c->pairs.size() : LB200_ERR_ARG; }
int lb200_clear_pairs(lb200_ctx *c) {
    if (!c) return LB200_ERR_ARG;
    c->pairs.clear();
    c->res.valid = false;
    return LB200_OK;
}

}  // extern "C"

static void parallel_for(int n, int threads, const std::function<void(int)> &fn) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, n));
    if (threads == 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() { for (int i = next++; i < n; i = next++) fn(i); });
    for (auto &t : pool) t.join();
}

// contiguous index blocks [begin, end) on the host threads: for per-item work of a few microseconds
static void parallel_blocks(int n, int threads, const std::function<void(int, int)> &fn) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, (n + 63) / 64));
    if (threads == 1) { if (n > 0) fn(0, n); return; }
    const int block = std::max(16, (n + threads * 4 - 1) / (threads * 4));
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() { for (int b = next.fetch_add(block); b < n; b = next.fetch_add(block)) fn(b, std::min(n, b + block)); });
    for (auto &t : pool) t.join();
}

template <class T>
static cudaError_t upload(DevBuf &b, const std::vector<T> &v, cudaStream_t st) {
    cudaError_t e = b.ensure(std::max<size_t>(v.size() * sizeof(T), 16));
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

extern "C" {

// sequences -> device arrays (codes, arcs in index order, arc weights, left-end index)
static int upload_sequences(lb200_ctx *c) {
    if (c->seqs_uploaded == c->seqs.size()) return LB200_OK;
    std::vector<uint8_t> codes, acodes;
    std::vector<int> al, ar, aw, asd, lptr, lcount;
    std::vector<double> pup, pdown;
    c->seq_codes_off.clear(); c->seq_arcs_off.clear(); c->seq_lptr_off.clear(); c->seq_prob_off.clear();
    for (const Sequence &s : c->seqs) {
        c->seq_codes_off.push_back((int)codes.size());
        codes.insert(codes.end(), s.codes.begin(), s.codes.end());
        acodes.insert(acodes.end(), s.anchor_rank.begin(), s.anchor_rank.end());
        acodes.resize(codes.size(), 0);
        c->seq_arcs_off.push_back((int)al.size());
        const std::vector<int> w = arc_weights(s, c->params);
        const std::vector<int> sd = (c->params.stacking || c->params.new_stacking) ? arc_stack_deltas(s, c->params) : std::vector<int>(s.arcs.size(), LB_NOSTACK);
        for (size_t k = 0; k < s.arcs.size(); k++) { al.push_back(s.arcs[k].left); ar.push_back(s.arcs[k].right); aw.push_back(w[k]); asd.push_back(sd[k]); }
        c->seq_lptr_off.push_back((int)lptr.size());
        lptr.insert(lptr.end(), s.lptr.begin(), s.lptr.end());
        lcount.insert(lcount.end(), s.lcount.begin(), s.lcount.end());
        c->seq_prob_off.push_back((int)pup.size());
        pup.insert(pup.end(), s.p_up.begin(), s.p_up.end());
        pdown.insert(pdown.end(), s.p_down.begin(), s.p_down.end());
    }
    cudaStream_t st = c->stream;
    std::vector<int> amseq(c->tables.am_seq, c->tables.am_seq + 256);
    CUDA_TRY(c, upload(c->d_codes, codes, st));
    CUDA_TRY(c, upload(c->d_acodes, acodes, st));
    CUDA_TRY(c, upload(c->d_arc_left, al, st));
    CUDA_TRY(c, upload(c->d_arc_right, ar, st));
    CUDA_TRY(c, upload(c->d_arc_weight, aw, st));
    CUDA_TRY(c, upload(c->d_arc_sdelta, asd, st));
    CUDA_TRY(c, upload(c->d_lptr, lptr, st));
    CUDA_TRY(c, upload(c->d_lcount, lcount, st));
    CUDA_TRY(c, upload(c->d_am_seq, amseq, st));
    CUDA_TRY(c, upload(c->d_pup, pup, st));
    CUDA_TRY(c, upload(c->d_pdown, pdown, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    c->last_h2d_bytes += (int64_t)(codes.size() + (al.size() * 3 + lptr.size() * 2 + 256) * 4);
    c->seqs_uploaded = c->seqs.size();
    return LB200_OK;
}

// Bands of all pairs that do not have one yet. Device contexts screen the probability envelope on the GPU in FP64
// (envelope.cu) and recompute only the flagged pairs on the host in long double; host-only contexts (and
// LB200_ENVELOPE=host) compute everything on the host. Either way the bands equal the reference's.
static int derive_bands(lb200_ctx *c) {
    const int P = (int)c->pairs.size();
    // anchor constraints (AnchorConstraints + TraceController::restrict_by_anchors, locarna.cc:551-576): pairs of sequences that both
    // carry anchor names get their band restricted before the probability envelope, and their arc matches filtered
    bool any_names = false;
    for (const Sequence &s : c->seqs) any_names |= !s.anchor_names.empty();
    if (any_names) {
        for (int k = 0; k < P; k++) {
            PairRec &r = c->pairs[k];
            const Sequence &A = c->seqs[r.seqA], &B = c->seqs[r.seqB];
            if (r.banded || A.anchor_names.empty() || B.anchor_names.empty()) continue;
            if (c->params.sequ_local || c->params.struct_local || c->params.fe_left1 || c->params.fe_left2 || c->params.fe_right1 || c->params.fe_right2)
                return fail(c, LB200_ERR_UNSUPPORTED, "anchor constraints are supported for global alignment without free end gaps only");
            const bool given = !r.band.lo.empty();
            Band b = given ? r.band : make_band(A.len, B.len, c->params.max_diff);
            std::string err;
            const int rc = restrict_band_by_anchors(b, A, B, err);
            if (rc < 0) return fail(c, LB200_ERR_UNSUPPORTED, "%s", err.c_str());
            for (int i = 0; i <= A.len; i++)
                if (b.lo[i] > b.hi[i] || (i > 0 && (b.lo[i] < b.lo[i - 1] || b.hi[i] < b.hi[i - 1])))
                    return fail(c, LB200_ERR_ARG, "anchor constraints leave no consistent band in row %d", i);
            r.anchored = true;
            if (!given || r.band_initial) { r.band = b; r.band_initial = true; }   // a final band given by the caller is kept as it is
        }
    }
    const bool ps = c->profile_mode();
    if (ps) {
        const Params &q = c->params;
        if (q.sequ_local || q.struct_local || q.fe_left1 || q.fe_left2 || q.fe_right1 || q.fe_right2)
            return fail(c, LB200_ERR_UNSUPPORTED, "profile (multi-row) input is supported for global alignment without free end gaps only");
        if (any_names) return fail(c, LB200_ERR_UNSUPPORTED, "anchor constraints with profile (multi-row) input are not supported");
        parallel_for(P, c->host_threads, [&](int k) {
            PairRec &r = c->pairs[k];
            if (r.prof) return;
            r.prof = std::make_shared<ProfileTables>();
            make_profile_tables(c->seqs[r.seqA], c->seqs[r.seqB], c->params, *r.prof);
        });
    }
    std::vector<int> todo;
    for (int k = 0; k < P; k++) {
        PairRec &r = c->pairs[k];
        if (r.banded) continue;
        if (!r.band.lo.empty() && !r.band_initial) { r.banded = true; continue; }
        todo.push_back(k);
    }
    c->env_device_pairs = 0; c->env_host_pairs = 0;
    if (todo.empty()) return LB200_OK;
    const bool envelope = c->params.min_trace_probability > 0.0;
    std::vector<char> on_host(todo.size(), 1);
    if (envelope && c->device != LB200_DEVICE_NONE && c->env_mode == 1 && !ps) {   // (profile columns: host envelope, stral_score.cc:29-44)
        { const int rc = upload_sequences(c); if (rc != LB200_OK) return rc; }
        cudaStream_t st = c->stream;
        std::vector<EnvPair> ep(todo.size());
        std::vector<int> lo, hi;
        size_t max_cells = 1;
        int env_rows = 1, env_cols = 1;
        {   // band offsets first (serial, trivial), then the per-pair work on the host threads
            size_t off = 0;
            for (size_t t = 0; t < todo.size(); t++) {
                const PairRec &r = c->pairs[todo[t]];
                const Sequence &A = c->seqs[r.seqA], &B = c->seqs[r.seqB];
                ep[t].band = (int)off; off += (size_t)A.len + 1;
                max_cells = std::max(max_cells, (size_t)(A.len + 1) * (B.len + 1));
                env_rows = std::max(env_rows, A.len + 1); env_cols = std::max(env_cols, B.len + 1);
            }
            lo.resize(off); hi.resize(off);
            parallel_blocks((int)todo.size(), c->host_threads, [&](int t0, int t1) {
                for (int t = t0; t < t1; t++) {
                    PairRec &r = c->pairs[todo[t]];
                    const Sequence &A = c->seqs[r.seqA], &B = c->seqs[r.seqB];
                    if (!r.band_initial) r.band = make_band(A.len, B.len, c->params.max_diff);
                    EnvPair &e = ep[t];
                    e.lenA = A.len; e.lenB = B.len; e.codesA = c->seq_codes_off[r.seqA]; e.codesB = c->seq_codes_off[r.seqB];
                    e.probA = c->seq_prob_off[r.seqA]; e.probB = c->seq_prob_off[r.seqB]; e.pad = 0;
                    std::copy(r.band.lo.begin(), r.band.lo.end(), lo.begin() + e.band);
                    std::copy(r.band.hi.begin(), r.band.hi.end(), hi.begin() + e.band);
                }
            });
        }
        // eight resident CTAs per SM (envelope.cu); the partition-function scratch takes at most 4 GB and at most a quarter of
        // the memory that is free right now
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(c, cudaMemGetInfo(&free_b, &total_b));
        const size_t scratch_cap = std::min<size_t>((size_t)4 << 30, (free_b + c->d_env_scratch.cap) / 4);
        const int grid = (int)std::min<size_t>(std::min<size_t>(todo.size(), (size_t)c->prop.multiProcessorCount * 8),
                                                 std::max<size_t>(1, scratch_cap / (envelope_scratch_doubles(max_cells, env_rows, env_cols) * sizeof(double))));
        EnvCtx e;
        memset(&e, 0, sizeof e);
        CUDA_TRY(c, upload(c->d_env_pairs, ep, st));
        CUDA_TRY(c, upload(c->d_env_lo, lo, st));
        CUDA_TRY(c, upload(c->d_env_hi, hi, st));
        CUDA_TRY(c, c->d_env_olo.ensure(lo.size() * 4));
        CUDA_TRY(c, c->d_env_ohi.ensure(lo.size() * 4));
        CUDA_TRY(c, c->d_env_flag.ensure(todo.size() * 4 + 16));
        CUDA_TRY(c, c->d_env_scratch.ensure((size_t)grid * envelope_scratch_doubles(max_cells, env_rows, env_cols) * sizeof(double)));
        CUDA_TRY(c, c->d_cursor.ensure(4100 * 4));
        e.pairs = (const EnvPair *)c->d_env_pairs.p; e.codes = (const uint8_t *)c->d_codes.p;
        e.p_up = (const double *)c->d_pup.p; e.p_down = (const double *)c->d_pdown.p;
        e.band_lo = (int *)c->d_env_lo.p; e.band_hi = (int *)c->d_env_hi.p; e.out_lo = (int *)c->d_env_olo.p; e.out_hi = (int *)c->d_env_ohi.p;
        e.out_flag = (int *)c->d_env_flag.p; e.scratch = (double *)c->d_env_scratch.p; e.scratch_doubles = envelope_scratch_doubles(max_cells, env_rows, env_cols);
        envelope_score_params(c->params, e.bm, &e.sw, &e.open, &e.ext, &e.temp);
        e.min_prob = c->params.min_trace_probability; e.local = c->params.sequ_local;
        e.fe_left1 = c->params.fe_left1; e.fe_right1 = c->params.fe_right1; e.fe_left2 = c->params.fe_left2; e.fe_right2 = c->params.fe_right2;
        CUDA_TRY(c, launch_envelope(e, (int)todo.size(), grid, (int *)c->d_cursor.p, env_rows, env_cols, st));
        std::vector<int> olo(lo.size()), ohi(lo.size()), flag(todo.size());
        CUDA_TRY(c, cudaMemcpyAsync(olo.data(), c->d_env_olo.p, lo.size() * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(ohi.data(), c->d_env_ohi.p, lo.size() * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(flag.data(), c->d_env_flag.p, todo.size() * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        parallel_blocks((int)todo.size(), c->host_threads, [&](int t0, int t1) {
            for (int t = t0; t < t1; t++) {
                if (flag[t] != 0) continue;  // uncertain: exact host computation below
                PairRec &r = c->pairs[todo[t]];
                std::copy(olo.begin() + ep[t].band, olo.begin() + ep[t].band + ep[t].lenA + 1, r.band.lo.begin());
                std::copy(ohi.begin() + ep[t].band, ohi.begin() + ep[t].band + ep[t].lenA + 1, r.band.hi.begin());
                r.banded = true;
                on_host[t] = 0;
            }
        });
        for (size_t t = 0; t < todo.size(); t++) c->env_device_pairs += on_host[t] ? 0 : 1;
    }
    parallel_for((int)todo.size(), c->host_threads, [&](int t) {
        if (!on_host[t]) return;
        PairRec &r = c->pairs[todo[t]];
        const Sequence &A = c->seqs[r.seqA], &B = c->seqs[r.seqB];
        if (!r.band_initial) r.band = make_band(A.len, B.len, c->params.max_diff);
        restrict_band_by_envelope(r.band, A, B, c->params);
        r.banded = true;
    });
    for (size_t t = 0; t < todo.size(); t++) c->env_host_pairs += on_host[t] ? 1 : 0;
    return LB200_OK;
}

int lb200_prepare(lb200_ctx *c) {
    if (!c) return LB200_ERR_ARG;
    if (c->device != LB200_DEVICE_NONE) CUDA_TRY(c, cudaSetDevice(c->device));
    { const int rc = derive_bands(c); if (rc != LB200_OK) return rc; }
    if (c->device != LB200_DEVICE_NONE) return LB200_OK;
    // host-only context: mirror of the device builder, for inspection without a GPU
    const int P = (int)c->pairs.size();
    parallel_for(P, c->host_threads, [&](int k) {
        PairRec &r = c->pairs[k];
        if (r.built) return;
        build_pair_problem(c->seqs[r.seqA], c->seqs[r.seqB], r.band, c->params, c->tables, r.prob, r.anchored, r.prof.get());
        r.K = (int)r.prob.am.size();
        r.stats.n_tasks = (long long)r.prob.tasks.size(); r.stats.cells = (long long)r.prob.cells; r.stats.terms = (long long)r.prob.terms;
        r.built = true;
    });
    return LB200_OK;
}

// Build the arc-match tables of the pairs [p0, p1) on the device and make them resident.
static int upload_chunk(lb200_ctx *c, int p0, int p1) {
    const int P = p1 - p0;
    c->res.valid = false;
    c->d_filled = false;
    if (P <= 0) return LB200_OK;
    { const int rc = upload_sequences(c); if (rc != LB200_OK) return rc; }
    cudaStream_t st = c->stream;

    // ---- per-pair inputs: band, cell ranks, offsets
    std::vector<DevPair> h_pairs(P);
    std::vector<int> h_lo, h_hi, h_rev, h_cfirst, h_clast;
    long long total_cells = 0, sptr_total = 0;
    int wd_bound = 1, max_rows = 1, max_cols = 1, max_box_words = 1, max_active = 0;
    {   // offsets first (serial, trivial), the O(n + m) work per pair on the host threads, then the reductions
        size_t band_total = 0;
        for (int k = 0; k < P; k++) {
            const PairRec &r = c->pairs[p0 + k];
            const int n = c->seqs[r.seqA].len, m = c->seqs[r.seqB].len;
            DevPair &d = h_pairs[k];
            memset(&d, 0, sizeof d);
            d.band = (int)band_total; band_total += (size_t)n + 1;
            d.sptr = (int)sptr_total; sptr_total += n + m + 3;
        }
        h_lo.resize(band_total); h_hi.resize(band_total); h_rev.resize(band_total);
        h_cfirst.assign((size_t)sptr_total, 1 << 20); h_clast.assign((size_t)sptr_total, -1);
        std::vector<int> p_wd(P), p_active(P);
        parallel_blocks(P, c->host_threads, [&](int k0, int k1) {
            for (int k = k0; k < k1; k++) {
                const PairRec &r = c->pairs[p0 + k];
                DevPair &d = h_pairs[k];
                const int n = c->seqs[r.seqA].len, m = c->seqs[r.seqB].len;
                d.lenA = n; d.lenB = m; d.anchored = r.anchored ? 1 : 0;
                d.codesA = c->seq_codes_off[r.seqA]; d.codesB = c->seq_codes_off[r.seqB];
                d.arcsA = c->seq_arcs_off[r.seqA]; d.arcsB = c->seq_arcs_off[r.seqB];
                d.lptrA = c->seq_lptr_off[r.seqA]; d.lptrB = c->seq_lptr_off[r.seqB];
                d.n_arcsB = (int)c->seqs[r.seqB].arcs.size(); d.ps_sig = -1; d.ps_am = -1;
                std::copy(r.band.lo.begin(), r.band.lo.end(), h_lo.begin() + d.band);
                std::copy(r.band.hi.begin(), r.band.hi.end(), h_hi.begin() + d.band);
                // cells (al, bl), al >= 1, bl >= 1, ranked al descending / bl descending; diagonal bound of any box
                int cells = 0, dmin = 0, dmax = 0;
                for (int i = n; i >= 0; i--) {
                    h_rev[d.band + i] = cells;
                    if (i >= 1) cells += std::max(0, std::min(r.band.hi[i], m) - std::max(r.band.lo[i], 1) + 1);
                    dmin = std::min(dmin, r.band.lo[i] - i); dmax = std::max(dmax, r.band.hi[i] - i);
                }
                d.n_cells = cells;
                p_wd[k] = dmax - dmin + 1;
                // column view of the band for the row-grouped sweep: rows [first, last] of every column (empty: first > last), and
                // the largest number of columns an anti-diagonal of the band crosses
                int *cf = h_cfirst.data() + d.sptr, *cl = h_clast.data() + d.sptr;
                // the band is monotone: row i is the first row of the columns beyond hi[i-1] and the last row of the columns before lo[i+1]
                int prev_hi = -1, next_lo = m + 1;
                for (int i = 0; i <= n; i++) {
                    const int h = std::min(r.band.hi[i], m);
                    for (int j = std::max(std::max(r.band.lo[i], 0), prev_hi + 1); j <= h; j++) cf[j] = i;
                    prev_hi = std::max(prev_hi, h);
                }
                for (int i = n; i >= 0; i--) {
                    const int l = std::max(r.band.lo[i], 0);
                    for (int j = l; j <= std::min(std::min(r.band.hi[i], m), next_lo - 1); j++) cl[j] = i;
                    next_lo = std::min(next_lo, l);
                }
                int j2 = 0, active = 0;
                for (int j = 0; j <= m; j++) {
                    if (cl[j] < cf[j]) continue;
                    if (j2 < j) j2 = j;
                    while (j2 + 1 <= m && cl[j2 + 1] >= cf[j2 + 1] && cf[j2 + 1] + j2 + 1 <= cl[j] + j) j2++;
                    active = std::max(active, j2 - j + 1);
                }
                p_active[k] = active;
            }
        });
        for (int k = 0; k < P; k++) {
            DevPair &d = h_pairs[k];
            d.cell_base = total_cells; total_cells += d.n_cells;
            max_active = std::max(max_active, p_active[k]);
            wd_bound = std::max(wd_bound, p_wd[k]);
            max_box_words = std::max(max_box_words, (d.lenA + d.lenB + 1) * ((((p_wd[k] + 1) / 2) + 3) & ~3));  // anti-diagonal major box, row stride: diagonal pairs rounded up to 4
            max_rows = std::max(max_rows, d.lenA + 1); max_cols = std::max(max_cols, d.lenB + 1);
        }
    }
    if (total_cells >= (1LL << 31) - 2) return fail(c, LB200_ERR_UNSUPPORTED, "batch too large (%lld band cells); split it", total_cells);

    // ---- kernel configuration
    const int nslots_bound = (wd_bound + 1) / 2;
    const int ncmax = (nslots_bound + 31) / 32;
    if (ncmax > LB_MAX_NC) return fail(c, LB200_ERR_UNSUPPORTED, "band of %d diagonals is wider than the supported %d", wd_bound, 64 * LB_MAX_NC);
    const int nc_inst = ncmax <= 1 ? 1 : ncmax <= 2 ? 2 : ncmax <= 4 ? 4 : ncmax <= 8 ? 8 : 16;
    DevCtx dc;
    memset(&dc, 0, sizeof dc);
    dc.params = c->tables.dev;
    const bool ps = c->profile_mode();
    if (ps) {
        // profile pairs run in the gap-free frame (host_model.h ProfileTables): sigma'(i, j) and the arc-match sequence terms from
        // tables, gap extension 0
        dc.params.gap = 0; dc.params.gap_open = dc.params.open;
        std::vector<int> h_sig, h_am;
        for (int k = 0; k < P; k++) {
            const PairRec &r = c->pairs[p0 + k];
            if (!r.prof) return fail(c, LB200_ERR_STATE, "profile tables missing (lb200_prepare not run for pair %d)", p0 + k);
            const ProfileTables &T = *r.prof;
            const Sequence &A = c->seqs[r.seqA], &B = c->seqs[r.seqB];
            DevPair &d = h_pairs[k];
            d.ps_sig = (long long)h_sig.size(); d.ps_am = (long long)h_am.size();
            for (int i = 0; i <= T.n; i++) for (int j = 0; j <= T.m; j++) h_sig.push_back(T.sigma_shifted(i, j));
            for (int a = 0; a < T.arcsA; a++)
                for (int bb = 0; bb < T.arcsB; bb++)
                    h_am.push_back(T.am_seq[(size_t)a * T.arcsB + bb] - T.gapA[A.arcs[a].left] - T.gapA[A.arcs[a].right] - T.gapB[B.arcs[bb].left] - T.gapB[B.arcs[bb].right]);
        }
        if (h_am.empty()) h_am.push_back(0);
        CUDA_TRY(c, upload(c->d_ps_sig, h_sig, st));
        CUDA_TRY(c, upload(c->d_ps_am, h_am, st));
        dc.ps_sig = (const int *)c->d_ps_sig.p;
        c->last_h2d_bytes += (int64_t)((h_sig.size() + h_am.size()) * 4);
    }
    dc.max_rows = max_rows;
    dc.rowcode_bytes = (max_rows + 2 + 3) & ~3;
    dc.colcode_bytes = (max_cols + 1 + 3) & ~3;
    const bool sl = c->params.struct_local;
    if (sl && ncmax > 4) return fail(c, LB200_ERR_UNSUPPORTED, "--struct-local supports bands of up to 256 diagonals (this batch: %d)", wd_bound);
    dc.arcbuf_words = 8 * 32 * nc_inst * (sl ? 4 : 1);  // structure local: one accumulator ring per closed state
    // padded tables of the single-state sweep: row index U2 + 1 - gidx spans [-(32 nc - 2), Rn + Cn/2 + 1], column index
    // J2 + gidx spans [-(Rn/2), Rn/2 + Cn + 32 nc]
    dc.row_pad = 32 * nc_inst;
    dc.row_words = dc.row_pad + max_rows + max_cols / 2 + 8;
    dc.col_pad = max_rows / 2 + 4;
    dc.col_bytes = (dc.col_pad + max_rows / 2 + max_cols + 32 * nc_inst + 8 + 3) & ~3;
    const int old_region = (dc.max_rows + 2) * 4 + dc.rowcode_bytes + dc.colcode_bytes;
    dc.region_bytes = (std::max(old_region, dc.row_words * 4 + dc.col_bytes) + 15) & ~15;  // arcbuf stays 16-byte aligned
    const int smem_bytes = (64 + dc.arcbuf_words + 32) * 4 + dc.region_bytes;   // + 32 junk slots behind the ring (lanes without an arc-match hit)
    if (smem_bytes > (int)c->prop.sharedMemPerBlockOptin) return fail(c, LB200_ERR_UNSUPPORTED, "problem needs %d bytes of shared memory per warp", smem_bytes);
    int ctas_per_sm = 1;
    CUDA_TRY(c, configure_kernels(nc_inst, smem_bytes, &ctas_per_sm));
    if (sl) CUDA_TRY(c, configure_sl(smem_bytes, &ctas_per_sm));
    if (const char *s = getenv("LB200_CTAS_PER_SM")) ctas_per_sm = std::max(1, std::min(ctas_per_sm, atoi(s)));   // experiment knob: fewer boxes in flight
    const int grid_cap = std::max(1, ctas_per_sm) * c->prop.multiProcessorCount;
    dc.scratch_words = max_box_words * (sl ? 8 : 1);  // structure local: eight matrices per box
    c->max_box_words = max_box_words; c->max_len = std::max(max_rows, max_cols);

    // ---- device builder
    CUDA_TRY(c, upload(c->d_pairs, h_pairs, st));
    CUDA_TRY(c, upload(c->d_band_lo, h_lo, st));
    CUDA_TRY(c, upload(c->d_band_hi, h_hi, st));
    CUDA_TRY(c, upload(c->d_cell_rev, h_rev, st));
    CUDA_TRY(c, c->d_cell_start.ensure((size_t)(total_cells + 1) * 4));
    CUDA_TRY(c, c->d_sptr.ensure((size_t)sptr_total * 4));
    CUDA_TRY(c, c->d_stats.ensure((size_t)P * sizeof(DevPairStats)));
    CUDA_TRY(c, c->d_ntasks.ensure(16));
    CUDA_TRY(c, c->d_qstart.ensure(4098 * 4));
    BuildCtx b;
    memset(&b, 0, sizeof b);
    b.pairs = (DevPair *)c->d_pairs.p; b.codes = (const uint8_t *)c->d_codes.p; b.acodes = (const uint8_t *)c->d_acodes.p;
    b.band_lo = (const int *)c->d_band_lo.p; b.band_hi = (const int *)c->d_band_hi.p; b.cell_rev = (const int *)c->d_cell_rev.p;
    b.arc_left = (const int *)c->d_arc_left.p; b.arc_right = (const int *)c->d_arc_right.p; b.arc_weight = (const int *)c->d_arc_weight.p; b.arc_sdelta = (const int *)c->d_arc_sdelta.p;
    b.lptr = (const int *)c->d_lptr.p; b.lcount = (const int *)c->d_lcount.p; b.am_seq = (const int *)c->d_am_seq.p;
    b.ps_am = ps ? (const int *)c->d_ps_am.p : nullptr;
    memcpy(b.sigma8, c->tables.dev.sigma8, sizeof b.sigma8);
    b.tau = c->params.tau; b.use_ribosum = c->params.use_ribosum; b.no_lonely_pairs = c->params.no_lonely_pairs; b.struct_local = c->params.struct_local;
    b.max_diff_am = c->params.max_diff_am; b.max_diff_at_am = c->params.max_diff_at_am;
    b.cell_start = (int *)c->d_cell_start.p; b.sptr = (int *)c->d_sptr.p; b.stats = (DevPairStats *)c->d_stats.p;
    b.n_tasks = (unsigned *)c->d_ntasks.p; b.qstart = (int *)c->d_qstart.p;
    size_t tmp_need = 0;
    CUDA_TRY(c, builder_count(b, P, total_cells, nullptr, 0, &tmp_need, st));
    CUDA_TRY(c, c->d_tmp.ensure(tmp_need));
    CUDA_TRY(c, builder_count(b, P, total_cells, c->d_tmp.p, c->d_tmp.cap, &tmp_need, st));
    int total_am_i = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&total_am_i, (int *)c->d_cell_start.p + total_cells, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    const size_t total_am = (size_t)total_am_i;
    const size_t task_cap = std::min<size_t>(total_am, (size_t)total_cells) + 1;
    CUDA_TRY(c, c->d_am.ensure(std::max<size_t>(total_am, 1) * sizeof(DevArcMatch)));
    CUDA_TRY(c, c->d_ent.ensure(std::max<size_t>(total_am, 1) * sizeof(DevEntry)));
    CUDA_TRY(c, c->d_skeys.ensure(std::max<size_t>(total_am, 1) * 8));
    CUDA_TRY(c, c->d_skeys2.ensure(std::max<size_t>(total_am, 1) * 8));
    CUDA_TRY(c, c->d_svals.ensure(std::max<size_t>(total_am, 1) * 4));
    CUDA_TRY(c, c->d_svals2.ensure(std::max<size_t>(total_am, 1) * 4));
    CUDA_TRY(c, c->d_tasks_unsorted.ensure(task_cap * sizeof(DevTask)));
    CUDA_TRY(c, c->d_tasks.ensure(task_cap * sizeof(DevTask)));
    CUDA_TRY(c, c->d_tkeys.ensure(task_cap * 4));
    CUDA_TRY(c, c->d_tkeys2.ensure(task_cap * 4));
    CUDA_TRY(c, c->d_tvals.ensure(task_cap * 4));
    CUDA_TRY(c, c->d_tvals2.ensure(task_cap * 4));
    CUDA_TRY(c, c->d_tmp.ensure(builder_sort_tmp_bytes((long long)std::max<size_t>(total_am, 1), P)));
    b.am = (DevArcMatch *)c->d_am.p; b.ent = (DevEntry *)c->d_ent.p;
    const bool pack = c->pack_entries && !sl && max_rows - 1 <= LB_PACK_MAXLEN && max_cols - 1 <= LB_PACK_MAXLEN;
    if (pack) CUDA_TRY(c, c->d_ent8.ensure(std::max<size_t>(total_am, 1) * sizeof(uint2)));
    b.ent8 = pack ? (uint2 *)c->d_ent8.p : nullptr;
    b.skeys = (unsigned long long *)c->d_skeys.p; b.skeys_sorted = (unsigned long long *)c->d_skeys2.p;
    b.svals = (unsigned *)c->d_svals.p; b.svals_sorted = (unsigned *)c->d_svals2.p;
    b.tasks_unsorted = (DevTask *)c->d_tasks_unsorted.p; b.tasks = (DevTask *)c->d_tasks.p;
    b.tkeys = (unsigned *)c->d_tkeys.p; b.tkeys_sorted = (unsigned *)c->d_tkeys2.p;
    b.tvals = (unsigned *)c->d_tvals.p; b.tvals_sorted = (unsigned *)c->d_tvals2.p;
    CUDA_TRY(c, builder_fill(b, P, (long long)total_am, sptr_total, c->d_tmp.p, c->d_tmp.cap, st));
    unsigned n_tasks = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n_tasks, c->d_ntasks.p, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    // automatic: the persistent launch pays off once an average level group holds at least half a wave of tasks (measured on B200:
    // +9..14 % at 256-512 pairs of 300 nt, equal at 1024, slower below ~100 pairs where most claimed tasks would only wait)
    const int n_levels = std::max(1, (max_rows + max_cols) >> 1);
    // row-grouped kernel (dfill_rows.cu): single-state boxes of packed batches whose borders fall out of the recurrence and whose
    // anti-diagonals cross at most ~60 band columns; everything else runs box by box (kernels.cu)
    const bool rows = !sl && !ps && pack && c->params.indel_opening <= 0 && n_tasks > 0 && max_active <= 58 && (c->dfill_mode == 2 || c->dfill_mode == 3);
    if (c->dfill_mode == 3 && !rows) return fail(c, LB200_ERR_UNSUPPORTED, "LB200_DFILL=rows: this batch is not eligible for the row-grouped kernel (max active columns %d)", max_active);
    const bool dep = !sl && !rows && (c->dfill_mode == 1 || (c->dfill_mode >= 2 && (long long)n_tasks * 2 >= (long long)grid_cap * n_levels));
    b.levcnt = nullptr; b.n_groups = ((max_rows + max_cols) >> 1) + 2; b.sb_pairs = c->sb_pairs;
    if (dep) {
        CUDA_TRY(c, c->d_levcnt.ensure((size_t)P * b.n_groups * sizeof(int)));
        CUDA_TRY(c, c->d_done.ensure((size_t)P * sizeof(int)));
        b.levcnt = (int *)c->d_levcnt.p;
    }
    CUDA_TRY(c, builder_sort_tasks(b, P, n_tasks, c->d_tmp.p, c->d_tmp.cap, st));
    dc.dep_order = dep ? (const unsigned *)c->d_tvals2.p : nullptr; dc.dep_need = (const int *)c->d_levcnt.p; dc.dep_done = (int *)c->d_done.p;
    dc.n_groups = b.n_groups; dc.n_tasks = (int)n_tasks;
    lb200_ctx::Resident &RR = c->res;
    RR.rows = false;
    if (rows) {
        GroupBuild gb;
        memset(&gb, 0, sizeof gb);
        gb.tasks = (const DevTask *)c->d_tasks.p; gb.n_tasks = n_tasks; gb.n_pairs = P;
        gb.keys = (unsigned long long *)c->d_skeys.p; gb.keys_sorted = (unsigned long long *)c->d_skeys2.p;
        gb.vals = (unsigned *)c->d_tvals.p; gb.vals_sorted = (unsigned *)c->d_tvals2.p;
        gb.gid = (int *)c->d_tkeys.p;
        gb.n_levels = max_rows + 2;
        CUDA_TRY(c, c->d_levcnt.ensure((size_t)P * gb.n_levels * sizeof(int)));
        CUDA_TRY(c, c->d_done.ensure((size_t)P * sizeof(int)));
        CUDA_TRY(c, c->d_ngroups.ensure(16));
        CUDA_TRY(c, c->d_tmp.ensure(builder_groups_tmp_bytes(n_tasks, P)));
        gb.levcnt = (int *)c->d_levcnt.p; gb.n_groups = (int *)c->d_ngroups.p;
        CUDA_TRY(c, builder_groups_scan(gb, c->d_tmp.p, c->d_tmp.cap, st));
        int n_groups = 0;
        CUDA_TRY(c, cudaMemcpyAsync(&n_groups, c->d_ngroups.p, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        CUDA_TRY(c, c->d_groups.ensure(std::max<size_t>(n_groups, 1) * sizeof(DevGroup)));
        CUDA_TRY(c, c->d_gorder.ensure(std::max<size_t>(n_groups, 1) * 4));
        gb.groups = (DevGroup *)c->d_groups.p; gb.gkeys = (unsigned *)c->d_tkeys2.p; gb.gkeys_sorted = (unsigned *)c->d_tkeys.p;
        gb.gvals = (unsigned *)c->d_svals.p; gb.order = (unsigned *)c->d_gorder.p;
        CUDA_TRY(c, builder_groups_fill(gb, (unsigned)n_groups, c->d_tmp.p, c->d_tmp.cap, st));
        CUDA_TRY(c, upload(c->d_col_first, h_cfirst, st));
        CUDA_TRY(c, upload(c->d_col_last, h_clast, st));
        RowsCtx &rc = RR.rc;
        memset(&rc, 0, sizeof rc);
        rc.groups = gb.groups; rc.order = gb.order; rc.n_groups = gb.n_groups;
        rc.col_first = (const int *)c->d_col_first.p; rc.col_last = (const int *)c->d_col_last.p;
        rc.dep_need = gb.levcnt; rc.dep_done = (int *)c->d_done.p; rc.n_levels = gb.n_levels;
        rc.nc_max = max_active <= 26 ? 1 : 2;
        if (const char *s = getenv("LB200_ROWS_NC")) rc.nc_max = std::max(1, std::min(2, atoi(s)));
        if (const char *s = getenv("LB200_ROWS_FORCE_NC")) { rc.force_nc = atoi(s); if (rc.force_nc == 2) rc.nc_max = 2; }
        rc.acc_words = LB_ROWS_RING * 32 * rc.nc_max * LB_GV;
        // the row in flight of every pair keeps its filtered entry list here: at most the pair's arc matches + a block of padding per
        // group of four target anti-diagonals
        int max_K = 1;
        {
            std::vector<DevPair> hp(P);
            CUDA_TRY(c, cudaMemcpyAsync(hp.data(), c->d_pairs.p, (size_t)P * sizeof(DevPair), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaStreamSynchronize(st));
            for (int k = 0; k < P; k++) max_K = std::max(max_K, hp[k].K);
        }
        rc.clist_cap = ((long long)max_K + 32LL * LB_ROWS_TG + 31) & ~31LL;
        CUDA_TRY(c, c->d_clist.ensure((size_t)P * rc.clist_cap * sizeof(uint2)));
        CUDA_TRY(c, c->d_cnblk.ensure((size_t)P * LB_ROWS_TG * sizeof(int)));
        CUDA_TRY(c, c->d_row_built.ensure((size_t)P * sizeof(int)));
        rc.clist = (uint2 *)c->d_clist.p; rc.cnblk = (int *)c->d_cnblk.p; rc.row_built = (int *)c->d_row_built.p;
        rc.colw_words = (max_cols + 1 + 80 + 3) & ~3;
        rc.row_pad = 32 * rc.nc_max + 8;
        rc.rowcode_bytes = (rc.row_pad + max_rows + max_cols + 8 + 3) & ~3;
        rc.scratch_words = (long long)(max_rows + max_cols) * 32 * rc.nc_max * LB_GV;
        RR.rows_smem = rows_smem_bytes(rc);
        int per_sm = 1;
        CUDA_TRY(c, rows_configure(RR.rows_smem, &per_sm));
        if (const char *s = getenv("LB200_ROWS_CTAS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(s)));
        RR.rows_grid = std::max(1, per_sm) * c->prop.multiProcessorCount;
        CUDA_TRY(c, c->d_rows_scratch.ensure((size_t)RR.rows_grid * rc.scratch_words * 4 + 16));
        rc.scratch = (int *)c->d_rows_scratch.p;
        RR.n_groups = (unsigned)n_groups;
        RR.rows = true;
        c->last_h2d_bytes += (int64_t)(h_cfirst.size() + h_clast.size()) * 4;
    }
    std::vector<DevPairStats> h_stats(P);
    CUDA_TRY(c, cudaMemcpyAsync(h_stats.data(), c->d_stats.p, (size_t)P * sizeof(DevPairStats), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(h_pairs.data(), c->d_pairs.p, (size_t)P * sizeof(DevPair), cudaMemcpyDeviceToHost, st));

    CUDA_TRY(c, c->d_top.ensure((size_t)P * sizeof(DevTopResult)));
    CUDA_TRY(c, c->d_scratch.ensure((size_t)grid_cap * dc.scratch_words * 4 + 16));
    CUDA_TRY(c, c->d_cursor.ensure(4100 * 4));
    CUDA_TRY(c, c->d_flag.ensure(16));
    dc.pairs = (const DevPair *)c->d_pairs.p; dc.codes = (const uint8_t *)c->d_codes.p;
    dc.band_lo = (const int *)c->d_band_lo.p; dc.band_hi = (const int *)c->d_band_hi.p; dc.sptr = (const int *)c->d_sptr.p;
    dc.ent = (DevEntry *)c->d_ent.p; dc.ent8 = b.ent8; dc.am = (const DevArcMatch *)c->d_am.p;
    dc.tasks = (const DevTask *)c->d_tasks.p; dc.top = (DevTopResult *)c->d_top.p; dc.scratch = (int *)c->d_scratch.p;
    dc.qstart = (const int *)c->d_qstart.p; dc.cursor = (int *)c->d_cursor.p;
    dc.error_flag = (int *)c->d_flag.p;
    CUDA_TRY(c, cudaStreamSynchronize(st));
    for (int k = 0; k < P; k++) {
        PairRec &r = c->pairs[p0 + k];
        r.am_base = h_pairs[k].am_base; r.K = h_pairs[k].K; r.stats = h_stats[k];
    }
    c->last_h2d_bytes += (int64_t)(h_pairs.size() * sizeof(DevPair) + (h_lo.size() + h_hi.size() + h_rev.size()) * 4);
    lb200_ctx::Resident &R = c->res;
    R.dc = dc; R.nc_inst = nc_inst; R.smem_bytes = smem_bytes; R.grid_cap = grid_cap; R.total_am = total_am;
    R.sptr_total = sptr_total; R.stack_cap = std::min(max_rows, max_cols) / 2 + 8;
    R.sptr_off.resize(P);
    for (int k = 0; k < P; k++) R.sptr_off[k] = h_pairs[k].sptr;
    R.dc.lpos = (const unsigned *)c->d_svals2.p;
    R.q_lo = 4095 - ((max_rows + max_cols) >> 1); R.q_hi = 4095;
    if (R.q_lo < 0) R.q_lo = 0;
    R.p0 = p0; R.p1 = p1;
    R.valid = true;
    return LB200_OK;
}

// Chunks of the pair list that are built and aligned together: bounded by a pair count (LB200_CHUNK_PAIRS, default
// 4096) and by the number of band cells (32-bit cell offsets on the device).
static std::vector<int> chunk_plan(lb200_ctx *c) {
    std::vector<int> cuts(1, 0);
    const int P = (int)c->pairs.size();
    int max_pairs = 4096;
    if (const char *s = getenv("LB200_CHUNK_PAIRS")) max_pairs = std::max(1, atoi(s));
    long long cells = 0;
    int count = 0;
    for (int k = 0; k < P; k++) {
        const PairRec &r = c->pairs[k];
        long long pc = 0;
        for (size_t i = 1; i < r.band.lo.size(); i++) pc += std::max(0, r.band.hi[i] - std::max(r.band.lo[i], 1) + 1);
        if (count > 0 && (count >= max_pairs || cells + pc > 1500000000LL)) { cuts.push_back(k); cells = 0; count = 0; }
        cells += pc; count++;
    }
    cuts.push_back(P);
    return cuts;
}

int lb200_upload(lb200_ctx *c) {
    if (!c) return LB200_ERR_ARG;
    if (c->device == LB200_DEVICE_NONE) return fail(c, LB200_ERR_CUDA, "host-only context: lb200_upload needs a CUDA device (no CPU fallback)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->res.valid = false;
    c->last_h2d_bytes = 0;
    if (c->pairs.empty()) return LB200_OK;
    { const int rc = derive_bands(c); if (rc != LB200_OK) return rc; }
    const std::vector<int> cuts = chunk_plan(c);
    if (cuts.size() > 2) return LB200_OK;  // several chunks: lb200_run streams them, nothing stays resident
    const int rc = upload_chunk(c, 0, (int)c->pairs.size());
    // the batch is resident now: the band-derivation scratch and the builder's sort temporaries are not needed to run it
    // (lb200_run without a prior lb200_upload keeps them for the next batch of a streaming caller)
    DevBuf *transient[] = {&c->d_env_scratch, &c->d_env_pairs, &c->d_env_lo, &c->d_env_hi, &c->d_env_olo, &c->d_env_ohi, &c->d_env_flag,
                           &c->d_skeys, &c->d_skeys2, &c->d_svals, &c->d_tasks_unsorted, &c->d_tkeys, &c->d_tkeys2, &c->d_tvals, &c->d_tmp};
    for (DevBuf *b : transient) b->release();
    return rc;
}

static int run_chunk(lb200_ctx *c, int flags);

static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int lb200_run(lb200_ctx *c, int flags) {
    if (!c) return LB200_ERR_ARG;
    const bool timing = getenv("LB200_TIMING") != nullptr;
    double t_bands = 0, t_upload = 0, t_run = 0, t0 = now_s();
    if (c->device == LB200_DEVICE_NONE) return fail(c, LB200_ERR_CUDA, "host-only context: lb200_run needs a CUDA device (no CPU fallback)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int P = (int)c->pairs.size();
    c->last_kernel_ms = 0; c->last_launches = 0; c->last_d2h_bytes = 0; c->last_dfill_ms = 0; c->last_dfill_launches = 0;
    if (P == 0) return LB200_OK;
    if (c->res.valid && c->res.p0 == 0 && c->res.p1 == P) { c->last_h2d_bytes = 0; return run_chunk(c, flags); }
    c->last_h2d_bytes = 0;
    { const int rc = derive_bands(c); if (rc != LB200_OK) return rc; }
    t_bands = now_s() - t0;
    const std::vector<int> cuts = chunk_plan(c);
    for (size_t k = 0; k + 1 < cuts.size(); k++) {
        double t1 = now_s();
        int rc = upload_chunk(c, cuts[k], cuts[k + 1]);
        t_upload += now_s() - t1; t1 = now_s();
        if (rc == LB200_OK) rc = run_chunk(c, flags);
        t_run += now_s() - t1;
        if (rc != LB200_OK) return rc;
    }
    if (timing) fprintf(stderr, "lb200_run: %d pairs, %zu chunks: bands %.3f s, build+upload %.3f s, run %.3f s (kernels %.1f ms)\n", P, cuts.size() - 1, t_bands, t_upload, t_run, c->last_kernel_ms);
    return LB200_OK;
}

// D fill, top level and (optionally) traceback of the resident chunk
static int run_chunk(lb200_ctx *c, int flags) {
    const int P = c->res.p1 - c->res.p0;
    const int p0 = c->res.p0;
    lb200_ctx::Resident &R = c->res;
    cudaStream_t st = c->stream;
    const bool do_trace = (flags & LB200_RUN_TRACE) != 0;
    if (do_trace) {
        CUDA_TRY(c, c->d_tr_edges.ensure((size_t)R.sptr_total * 4));
        CUDA_TRY(c, c->d_tr_str.ensure((size_t)R.sptr_total));
        CUDA_TRY(c, c->d_tr_stack.ensure((size_t)P * R.stack_cap * sizeof(TraceJob)));
        R.dc.trace_edges = (int *)c->d_tr_edges.p; R.dc.trace_str = (char *)c->d_tr_str.p;
        R.dc.trace_stack = (TraceJob *)c->d_tr_stack.p; R.dc.trace_stack_cap = R.stack_cap;
    }
    const DevCtx &dc = R.dc;

    // ---- run: D entries start as -inf (aligner.cc:122-123)
    CUDA_TRY(c, cudaEventRecord(c->ev0, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_cursor.p, 0, 4100 * 4, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_flag.p, 0, 16, st));
    CUDA_TRY(c, lb200_reset_d((DevEntry *)c->d_ent.p, dc.ent8, R.total_am, st));
    int64_t launches = 1;
    int dfill_launches = 0;
    c->last_dfill_kind = R.rows ? 2 : (dc.dep_order != nullptr ? 1 : 0);
    if (R.rows) {   // row-grouped sweep: one persistent launch, groups ordered by origin row
        CUDA_TRY(c, cudaMemsetAsync(c->d_done.p, 0, (size_t)P * sizeof(int), st));
        CUDA_TRY(c, cudaMemsetAsync(c->d_row_built.p, 0, (size_t)P * sizeof(int), st));
        if (R.n_groups > 0) {
            launch_dfill_rows(dc, R.rc, (int)std::min<unsigned>((unsigned)R.rows_grid, R.n_groups), R.rows_smem, (int *)c->d_cursor.p + 4097, st);
            dfill_launches = 1;
        }
    } else if (dc.dep_order != nullptr) {   // one persistent launch, tasks ordered by their own dependencies
        CUDA_TRY(c, cudaMemsetAsync(c->d_done.p, 0, (size_t)P * sizeof(int), st));
        if (dc.n_tasks > 0) {
            launch_dfill_dep(dc, R.nc_inst, c->params.indel_opening > 0 || dc.ps_sig != nullptr, std::min(R.grid_cap, dc.n_tasks), R.smem_bytes, (int *)c->d_cursor.p + 4097, st);
            dfill_launches = 1;
        }
    } else {
        for (int q = R.q_lo; q <= R.q_hi; q++) {
            if (c->params.struct_local) launch_dfill_sl(dc, R.grid_cap, R.smem_bytes, q, st);
            else launch_dfill(dc, R.nc_inst, c->params.indel_opening > 0 || dc.ps_sig != nullptr, R.grid_cap, R.smem_bytes, q, st);
            dfill_launches++;
        }
    }
    launches += dfill_launches;
    CUDA_TRY(c, cudaEventRecord(c->ev_mid, st));
    launch_toplevel(dc, R.nc_inst, std::min(R.grid_cap, P), R.smem_bytes, 0, P, (int *)c->d_cursor.p + 4098, st);
    launches++;
    if (do_trace) {
        if (c->params.struct_local) launch_trace_sl(dc, std::min(R.grid_cap, P), R.smem_bytes, 0, P, (int *)c->d_cursor.p + 4099, st);
        else launch_trace(dc, R.nc_inst, c->params.indel_opening > 0 || dc.ps_sig != nullptr, std::min(R.grid_cap, P), R.smem_bytes, 0, P, (int *)c->d_cursor.p + 4099, st);
        launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->ev1, st));

    // ---- results
    std::vector<DevTopResult> h_top(P);
    int h_flag[4] = {0, 0, 0, 0};
    CUDA_TRY(c, cudaMemcpyAsync(h_top.data(), c->d_top.p, (size_t)P * sizeof(DevTopResult), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(h_flag, c->d_flag.p, 16, cudaMemcpyDeviceToHost, st));
    std::vector<int> h_edges;
    std::vector<char> h_str;
    if (do_trace) {
        h_edges.resize(R.sptr_total); h_str.resize(R.sptr_total);
        CUDA_TRY(c, cudaMemcpyAsync(h_edges.data(), c->d_tr_edges.p, (size_t)R.sptr_total * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(h_str.data(), c->d_tr_str.p, (size_t)R.sptr_total, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(c, cudaStreamSynchronize(st));
    float ms = 0, ms_dfill = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    CUDA_TRY(c, cudaEventElapsedTime(&ms_dfill, c->ev0, c->ev_mid));
    c->last_kernel_ms += ms; c->last_launches += launches; c->last_dfill_ms += ms_dfill; c->last_dfill_launches += dfill_launches;
    c->last_d2h_bytes += (int64_t)((size_t)P * sizeof(DevTopResult) + 16 + h_edges.size() * 4 + h_str.size());
    if (h_flag[0] == 5 && R.rows) {   // a box the row-grouped kernel does not handle: the whole chunk again, box by box
        R.rows = false;
        c->rows_fallbacks++;
        return run_chunk(c, flags);
    }
    if (h_flag[0] != 0)
        return fail(c, LB200_ERR_UNSUPPORTED, "kernel reported error %d (1: band too wide, 2: box exceeds scratch, 3: trace box failed, 4: traceback dead end)", h_flag[0]);
    for (int k = 0; k < P; k++) {
        PairRec &r = c->pairs[p0 + k];
        r.neg_inf = h_top[k].score < LB_NEG_LIMIT;
        r.score = r.neg_inf ? 0 : h_top[k].score;
        if (r.prof && !r.neg_inf) r.score += r.prof->PA[r.prof->n] + r.prof->PB[r.prof->m];   // back from the gap-free frame: M(n, m) = M'(n, m) + PA[n] + PB[m]
        r.max_i = h_top[k].max_i; r.max_j = h_top[k].max_j;
        r.edges_a.clear(); r.edges_b.clear(); r.str_a.clear(); r.str_b.clear();
        if (do_trace) {
            // compact the edge slots (slot index = i + j increases along the alignment)
            const int n = c->seqs[r.seqA].len, m = c->seqs[r.seqB].len, off = R.sptr_off[k];
            for (int idx = 0; idx <= n + m; idx++) {
                const int v = h_edges[off + idx];
                if (v == 0) continue;
                const int i = v >> 2, j = idx - i, kind = v & 3;
                r.edges_a.push_back(kind == 3 ? -1 : i);
                r.edges_b.push_back(kind == 2 ? -1 : j);
            }
            r.str_a.assign(h_str.begin() + off + 1, h_str.begin() + off + n + 1);
            r.str_b.assign(h_str.begin() + off + n + 2, h_str.begin() + off + n + 2 + m);
            r.traced = true;
        }
    }
    c->d_filled = true;
    return LB200_OK;
}

double lb200_last_kernel_ms(const lb200_ctx *c) { return c ? c->last_kernel_ms : 0; }
double lb200_last_dfill_ms(const lb200_ctx *c) { return c ? c->last_dfill_ms : 0; }
int64_t lb200_last_dfill_launches(const lb200_ctx *c) { return c ? c->last_dfill_launches : 0; }
int64_t lb200_last_h2d_bytes(const lb200_ctx *c) { return c ? c->last_h2d_bytes : 0; }
int64_t lb200_last_d2h_bytes(const lb200_ctx *c) { return c ? c->last_d2h_bytes : 0; }
int64_t lb200_last_launches(const lb200_ctx *c) { return c ? c->last_launches : 0; }
int lb200_last_dfill_kind(const lb200_ctx *c) { return c ? c->last_dfill_kind : LB200_ERR_ARG; }
int64_t lb200_rows_fallbacks(const lb200_ctx *c) { return c ? c->rows_fallbacks : 0; }
int lb200_envelope_stats(const lb200_ctx *c, int64_t *device_pairs, int64_t *host_pairs) {
    if (!c) return LB200_ERR_ARG;
    if (device_pairs) *device_pairs = c->env_device_pairs;
    if (host_pairs) *host_pairs = c->env_host_pairs;
    return LB200_OK;
}

int lb200_pair_score(const lb200_ctx *c, int pair, int64_t *score) {
    if (!c || !score || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    *score = r.neg_inf ? LB200_SCORE_NEG_INF : r.score;
    return LB200_OK;
}

int lb200_get_scores(const lb200_ctx *c, int64_t *scores, int n) {
    if (!c || !scores || n != (int)c->pairs.size()) return LB200_ERR_ARG;
    for (int k = 0; k < n; k++) scores[k] = c->pairs[k].neg_inf ? LB200_SCORE_NEG_INF : c->pairs[k].score;
    return LB200_OK;
}

int lb200_pair_get_info(const lb200_ctx *c, int pair, lb200_pair_info *info) {
    if (!c || !info || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    memset(info, 0, sizeof *info);
    info->lenA = c->seqs[r.seqA].len; info->lenB = c->seqs[r.seqB].len;
    info->n_arcsA = (int)c->seqs[r.seqA].arcs.size(); info->n_arcsB = (int)c->seqs[r.seqB].arcs.size();
    info->n_arcmatches = r.K; info->n_tasks = r.stats.n_tasks; info->cells = r.stats.cells; info->terms = r.stats.terms;
    info->n_edges = (int64_t)r.edges_a.size();
    return LB200_OK;
}

int lb200_pair_band(const lb200_ctx *c, int pair, int *min_col, int *max_col) {
    if (!c || !min_col || !max_col || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    if (r.band.lo.empty()) return LB200_ERR_STATE;
    std::copy(r.band.lo.begin(), r.band.lo.end(), min_col);
    std::copy(r.band.hi.begin(), r.band.hi.end(), max_col);
    return LB200_OK;
}

// ------------------------------------------------------------------------------------------------ normalized / penalized alignment
// aligner.cc:1522-1622. The D table is filled as for an ordinary alignment; then the TOP LEVEL is aligned and traced with the scoring
// modified by lambda (ModifiedScoringView, aligner_impl.hh:190-275; Scoring::modify_by_parameter, scoring.cc:77-90): sigma - 2 lambda,
// gap - lambda, D(a, b) - lambda * (length of arc a + length of arc b). Normalized alignment iterates lambda := score / (length + L)
// (Dinkelbach) until it does not change; penalized alignment is one pass with lambda = the position penalty. One pair per launch:
// lambda differs from pair to pair after the first iteration.
cudaError_t lb200_adjust_d(const DevEntry *ent, DevEntry *ent_mod, uint2 *ent8_mod, size_t n, int lambda, cudaStream_t st);

// mode 0: plain top level (Aligner::align + trace under a restriction), 1: normalized, 2: penalized; pair < 0: all pairs
static int run_modified(lb200_ctx *c, int mode, int64_t arg, int pair, bool do_trace) {
    const bool normalized = mode == 1;
    if (c->device == LB200_DEVICE_NONE) return fail(c, LB200_ERR_CUDA, "host-only context: needs a CUDA device (no CPU fallback)");
    if (c->profile_mode()) return fail(c, LB200_ERR_UNSUPPORTED, "normalized, penalized and restricted (k-best) alignment are not supported for profile (multi-row) input");
    if (c->params.struct_local) return fail(c, LB200_ERR_UNSUPPORTED, normalized ? "Normalized structure local alignment not supported." : mode == 2 ? "penalized structure local alignment is not supported" : "restricted structure local alignment is not supported");
    if (normalized && !c->params.sequ_local) return fail(c, LB200_ERR_ARG, "Cannot run normalized alignment without --sequ_local on.");   // locarna.cc:431-436
    const int P = (int)c->pairs.size();
    if (pair >= P) return fail(c, LB200_ERR_ARG, "no such pair");
    if (P == 0) return LB200_OK;
    lb200_ctx::Resident &R = c->res;
    // D fill (aligner.cc:1529-1530, :1607-1608); the table of an earlier run is kept (D_created_)
    if (!(R.valid && R.p0 == 0 && R.p1 == P && c->d_filled)) {
        const int rc = lb200_run(c, LB200_RUN_SCORE_ONLY);
        if (rc != LB200_OK) return rc;
    }
    if (!(R.valid && R.p0 == 0 && R.p1 == P)) return fail(c, LB200_ERR_UNSUPPORTED, "normalized / penalized / restricted alignment needs the whole batch resident; split the pair list");
    cudaStream_t st = c->stream;
    const bool packed = R.dc.ent8 != nullptr;
    if (mode != 0) {
        CUDA_TRY(c, c->d_ent_mod.ensure(std::max<size_t>(R.total_am, 1) * sizeof(DevEntry)));
        if (packed) CUDA_TRY(c, c->d_ent8_mod.ensure(std::max<size_t>(R.total_am, 1) * sizeof(uint2)));
    }
    CUDA_TRY(c, c->d_tr_edges.ensure((size_t)R.sptr_total * 4));
    CUDA_TRY(c, c->d_tr_str.ensure((size_t)R.sptr_total));
    CUDA_TRY(c, c->d_tr_stack.ensure((size_t)P * R.stack_cap * sizeof(TraceJob)));
    R.dc.trace_edges = (int *)c->d_tr_edges.p; R.dc.trace_str = (char *)c->d_tr_str.p;
    R.dc.trace_stack = (TraceJob *)c->d_tr_stack.p; R.dc.trace_stack_cap = R.stack_cap;
    int64_t launches = 0;
    CUDA_TRY(c, cudaEventRecord(c->ev0, st));
    for (int k = (pair < 0 ? 0 : pair); k < (pair < 0 ? P : pair + 1); k++) {
        PairRec &r = c->pairs[k];
        const int n = c->seqs[r.seqA].len, m = c->seqs[r.seqB].len, off = R.sptr_off[k];
        long new_lambda = normalized ? 0 : (long)arg, lambda = normalized ? -1 : (long)arg - 1;
        DevTopResult top;
        std::vector<int> h_edges((size_t)n + m + 3);
        std::vector<char> h_str((size_t)n + m + 3);
        int iterations = 0;
        while (lambda != new_lambda) {
            lambda = new_lambda;
            if (++iterations > 10000) return fail(c, LB200_ERR_STATE, "normalized alignment does not converge");
            if ((lambda < 0 ? -lambda : lambda) * (long)(n + m + 2) > 50000000L) return fail(c, LB200_ERR_UNSUPPORTED, "modification parameter %ld out of the supported range", lambda);
            // scoring view of this pair's top level
            DevCtx ct = R.dc;
            if (r.restricted) { ct.r_on = 1; ct.r_sa = r.r_sa; ct.r_sb = r.r_sb; ct.r_ea = r.r_ea; ct.r_eb = r.r_eb; }
            if (mode != 0) {
                for (int x = 0; x < 64; x++) ct.params.sigma8[x] -= 2 * (int)lambda;
                ct.params.gap -= (int)lambda;
                ct.params.gap_open = ct.params.gap + ct.params.open;
                CUDA_TRY(c, lb200_adjust_d((const DevEntry *)c->d_ent.p + r.am_base, (DevEntry *)c->d_ent_mod.p + r.am_base,
                                           packed ? (uint2 *)c->d_ent8_mod.p + r.am_base : nullptr, (size_t)r.K, (int)lambda, st));
                ct.ent = (DevEntry *)c->d_ent_mod.p; ct.ent8 = packed ? (uint2 *)c->d_ent8_mod.p : nullptr;
                launches++;
            }
            CUDA_TRY(c, cudaMemsetAsync(c->d_cursor.p, 0, 4100 * 4, st));
            CUDA_TRY(c, cudaMemsetAsync(c->d_flag.p, 0, 16, st));
            launch_toplevel(ct, R.nc_inst, 1, R.smem_bytes, k, k + 1, (int *)c->d_cursor.p + 4098, st);
            launches++;
            if (do_trace) {
                DevCtx dt = R.dc;
                dt.r_on = ct.r_on; dt.r_sa = ct.r_sa; dt.r_sb = ct.r_sb; dt.r_ea = ct.r_ea; dt.r_eb = ct.r_eb;
                if (mode != 0) { dt.use_tl = 1; dt.params_tl = ct.params; dt.ent_tl = ct.ent; dt.ent8_tl = ct.ent8; }
                launch_trace(dt, R.nc_inst, c->params.indel_opening > 0, 1, R.smem_bytes, k, k + 1, (int *)c->d_cursor.p + 4099, st);
                launches++;
            }
            int h_flag[4] = {0, 0, 0, 0};
            CUDA_TRY(c, cudaMemcpyAsync(&top, (DevTopResult *)c->d_top.p + k, sizeof top, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaMemcpyAsync(h_flag, c->d_flag.p, 16, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(c, cudaStreamSynchronize(st));
            if (h_flag[0] != 0) return fail(c, LB200_ERR_UNSUPPORTED, "kernel reported error %d in the top level of pair %d", h_flag[0], k);
            if (!normalized) break;
            if (top.score < LB_NEG_LIMIT) return fail(c, LB200_ERR_STATE, "normalized alignment: the local score is -inf");
            // aligner.cc:1566-1575: length of the aligned subsequences from the trace; the modified score plus length * lambda is the
            // unmodified score of this alignment (every position lost lambda); the arithmetic is the reference's (size_t + long, unsigned)
            const int min_i = top.min_ij & 0xffff, min_j = (top.min_ij >> 16) & 0xffff;
            const unsigned long length = (unsigned long)top.max_i - (unsigned long)min_i + 1 + (unsigned long)top.max_j - (unsigned long)min_j + 1;
            const long score = (long)top.score + (long)(length * (unsigned long)lambda);
            new_lambda = (long)((unsigned long)score / (length + (unsigned long)arg));
        }
        if (normalized) { r.neg_inf = false; r.score = new_lambda; }                       // aligner.cc:1596
        else { r.neg_inf = top.score < LB_NEG_LIMIT; r.score = r.neg_inf ? 0 : top.score; }
        r.max_i = top.max_i; r.max_j = top.max_j;
        r.edges_a.clear(); r.edges_b.clear(); r.str_a.clear(); r.str_b.clear(); r.traced = false;
        if (!do_trace) continue;
        CUDA_TRY(c, cudaMemcpyAsync(h_edges.data(), (int *)c->d_tr_edges.p + off, ((size_t)n + m + 3) * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaMemcpyAsync(h_str.data(), (char *)c->d_tr_str.p + off, (size_t)n + m + 3, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(c, cudaStreamSynchronize(st));
        for (int idx = 0; idx <= n + m; idx++) {
            const int v = h_edges[idx];
            if (v == 0) continue;
            const int i = v >> 2, j = idx - i, kind = v & 3;
            r.edges_a.push_back(kind == 3 ? -1 : i);
            r.edges_b.push_back(kind == 2 ? -1 : j);
        }
        r.str_a.assign(h_str.begin() + 1, h_str.begin() + n + 1);
        r.str_b.assign(h_str.begin() + n + 2, h_str.begin() + n + 2 + m);
        r.traced = true;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev1, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_kernel_ms += ms; c->last_launches += launches;
    return LB200_OK;
}

int lb200_run_normalized(lb200_ctx *c, int64_t L) { return c ? run_modified(c, 1, L, -1, true) : LB200_ERR_ARG; }
int lb200_run_penalized(lb200_ctx *c, int64_t position_penalty) { return c ? run_modified(c, 2, position_penalty, -1, true) : LB200_ERR_ARG; }

// AlignerRestriction of a pair's top level (Aligner::set_restriction, aligner.cc:1368-1376); (1, 1, lenA, lenB) lifts it
int lb200_pair_set_restriction(lb200_ctx *c, int pair, int startA, int startB, int endA, int endB) {
    if (!c || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    PairRec &r = c->pairs[pair];
    const int n = c->seqs[r.seqA].len, m = c->seqs[r.seqB].len;
    if (startA < 1 || startB < 1 || endA > n || endB > m || startA > endA + 1 || startB > endB + 1) return fail(c, LB200_ERR_ARG, "restriction (%d, %d, %d, %d) outside the sequences", startA, startB, endA, endB);
    r.restricted = !(startA == 1 && startB == 1 && endA == n && endB == m);
    r.r_sa = startA; r.r_sb = startB; r.r_ea = endA; r.r_eb = endB;
    return LB200_OK;
}

// Aligner::align (+ trace) of ONE pair on the D table that is already filled (aligner.cc:924-962: "alignment needs to be recomputed
// (this is fast)!", :1426-1431), under the pair's restriction; mode LB200_TOP_PLAIN / _NORMALIZED / _PENALIZED with its parameter
int lb200_run_pair_toplevel(lb200_ctx *c, int pair, int mode, int64_t arg, int flags) {
    if (!c || pair < 0 || mode < 0 || mode > 2) return LB200_ERR_ARG;
    return run_modified(c, mode, arg, pair, (flags & LB200_RUN_TRACE) != 0 || mode == 1);
}

// ------------------------------------------------------------------------------------------------ LocARNA-P inside
int lb200_run_pf(lb200_ctx *c, double pf_scale) {
    if (!c) return LB200_ERR_ARG;
    if (c->device == LB200_DEVICE_NONE) return fail(c, LB200_ERR_CUDA, "host-only context: lb200_run_pf needs a CUDA device (no CPU fallback)");
    if (c->params.no_lonely_pairs || c->params.struct_local || c->params.sequ_local)
        return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P (AlignerP) has no noLP / struct-local / sequ-local mode");
    if (!(pf_scale > 0)) return fail(c, LB200_ERR_ARG, "pf_scale must be positive");
    if (c->profile_mode()) return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P with profile (multi-row) input is not supported");
    for (const Sequence &s : c->seqs) if (!s.anchor_names.empty()) return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P with anchor constraints is not supported");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int P = (int)c->pairs.size();
    if (P == 0) return LB200_OK;
    if (!(c->res.valid && c->res.p0 == 0 && c->res.p1 == P)) {
        const int rc = lb200_upload(c);
        if (rc != LB200_OK) return rc;
        if (!(c->res.valid && c->res.p0 == 0 && c->res.p1 == P))
            return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P needs the whole batch resident; split the pair list (LB200_CHUNK_PAIRS)");
    }
    lb200_ctx::Resident &R = c->res;
    cudaStream_t st = c->stream;
    if (R.nc_inst > 8) return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P supports bands of up to 512 diagonals");
    // Boltzmann weights (scoring.hh:853-931), computed with the host's libm like the reference
    const double temp = (double)c->params.temperature_alipf;
    std::vector<double> esig(64), bpow((size_t)c->max_len + 2);
    for (int k = 0; k < 64; k++) esig[k] = std::exp((long)c->tables.dev.sigma8[k] / temp);
    const double g = std::exp((long)c->tables.dev.gap / temp), open = std::exp((long)c->tables.dev.open / temp);
    double v = open / pf_scale;                       // aligner_p.icc:156, :170
    bpow[0] = v;
    for (size_t k = 1; k < bpow.size(); k++) { v *= g; bpow[k] = v; }
    CUDA_TRY(c, upload(c->d_pf_esig, esig, st));
    CUDA_TRY(c, upload(c->d_pf_bpow, bpow, st));
    const int smem_bytes = ((512 + (R.dc.max_rows + 2) * 4 + R.dc.rowcode_bytes + R.dc.colcode_bytes + 7) & ~7) + 32 * R.nc_inst * 8;
    if (smem_bytes > (int)c->prop.sharedMemPerBlockOptin) return fail(c, LB200_ERR_UNSUPPORTED, "problem needs %d bytes of shared memory per warp", smem_bytes);
    int ctas_per_sm = 1;
    CUDA_TRY(c, configure_pf(R.nc_inst, smem_bytes, &ctas_per_sm));
    const int grid_cap = std::max(1, ctas_per_sm) * c->prop.multiProcessorCount;
    CUDA_TRY(c, c->d_pf_d.ensure(std::max<size_t>(R.total_am, 1) * 8));
    CUDA_TRY(c, c->d_pf_z.ensure((size_t)P * 8));
    CUDA_TRY(c, c->d_pf_scratch.ensure((size_t)grid_cap * c->max_box_words * 8 + 16));
    PfCtx pc;
    pc.esig = (const double *)c->d_pf_esig.p; pc.bpow = (const double *)c->d_pf_bpow.p;
    pc.g = g; pc.open = open; pc.inv_scale = 1.0 / pf_scale; pc.pf_scale = pf_scale; pc.temp = temp;
    pc.dpf = (double *)c->d_pf_d.p; pc.ztop = (double *)c->d_pf_z.p; pc.scratch = (double *)c->d_pf_scratch.p;
    pc.scratch_dwords = c->max_box_words; pc.acc_doubles = 32 * R.nc_inst;
    CUDA_TRY(c, cudaEventRecord(c->ev0, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_cursor.p, 0, 4100 * 4, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_flag.p, 0, 16, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_pf_d.p, 0, std::max<size_t>(R.total_am, 1) * 8, st));   // Dmat.fill(0), aligner_p.icc:24-26
    int64_t launches = 0;
    for (int q = R.q_lo; q <= R.q_hi; q++) { launch_pfill(R.dc, pc, R.nc_inst, grid_cap, smem_bytes, q, st); launches++; }
    launch_ptop(R.dc, pc, R.nc_inst, std::min(grid_cap, P), smem_bytes, 0, P, (int *)c->d_cursor.p + 4098, st);
    launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->ev1, st));
    std::vector<double> z(P);
    int h_flag[4] = {0, 0, 0, 0};
    CUDA_TRY(c, cudaMemcpyAsync(z.data(), c->d_pf_z.p, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaMemcpyAsync(h_flag, c->d_flag.p, 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_kernel_ms = ms; c->last_launches = launches; c->last_d2h_bytes = (int64_t)P * 8 + 16;
    if (h_flag[0] != 0) return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P kernel reported error %d (1: band too wide, 2: box exceeds scratch)", h_flag[0]);
    for (int k = 0; k < P; k++) { c->pairs[k].pf_Z = z[k]; c->pairs[k].pf_done = true; }
    c->pf_last = pc; c->pf_have = true; c->pf_probs_done = false;
    return LB200_OK;
}

static void reference_am_order(const std::vector<DevArcMatch> &am, std::vector<int> &order);

// LocARNA-P outside pass and probabilities (aligner_p.icc:440-1399) after the inside pass
int lb200_run_pf_probs(lb200_ctx *c, double pf_scale, double min_am_prob) {
    { const int rc = lb200_run_pf(c, pf_scale); if (rc != LB200_OK) return rc; }
    const int P = (int)c->pairs.size();
    if (P == 0) return LB200_OK;
    if (!(min_am_prob >= 0)) return fail(c, LB200_ERR_ARG, "min_am_prob must not be negative");
    lb200_ctx::Resident &R = c->res;
    cudaStream_t st = c->stream;
    const double kernel_ms_inside = c->last_kernel_ms;
    int64_t launches = c->last_launches;
    long long mat = 1;
    for (int k = 0; k < P; k++) mat = std::max(mat, (long long)(c->seqs[c->pairs[k].seqA].len + 1) * (c->seqs[c->pairs[k].seqB].len + 1));
    mat = (mat + 1) & ~1LL;
    const int grid = (int)std::max<long long>(1, std::min<long long>(4LL * c->prop.multiProcessorCount, (8LL << 30) / (4 * mat * 8)));   // <= 8 GB of CTA scratch
    CUDA_TRY(c, c->d_pf_dp.ensure(std::max<size_t>(R.total_am, 1) * 8));
    CUDA_TRY(c, c->d_pf_amp.ensure(std::max<size_t>(R.total_am, 1) * 8));
    CUDA_TRY(c, c->d_pf_mats.ensure((size_t)P * 5 * mat * 8));
    CUDA_TRY(c, c->d_pf_cta.ensure((size_t)grid * 4 * mat * 8));
    if (!getenv("LB200_PF_NOCLEAR")) {   // the dense tables start zeroed like the reference's freshly constructed matrices (aligner_p.icc:24-47)
        CUDA_TRY(c, cudaMemsetAsync(c->d_pf_mats.p, 0, (size_t)P * 5 * mat * 8, st));
        CUDA_TRY(c, cudaMemsetAsync(c->d_pf_cta.p, 0, (size_t)grid * 4 * mat * 8, st));
    }
    PfoCtx o;
    o.pc = c->pf_last;
    o.cell_start = (const int *)c->d_cell_start.p; o.cell_rev = (const int *)c->d_cell_rev.p;
    o.arc_left = (const int *)c->d_arc_left.p; o.arc_right = (const int *)c->d_arc_right.p;
    o.lptr = (const int *)c->d_lptr.p; o.lcount = (const int *)c->d_lcount.p;
    o.dpp = (double *)c->d_pf_dp.p; o.amp = (double *)c->d_pf_amp.p; o.mats = (double *)c->d_pf_mats.p; o.cta = (double *)c->d_pf_cta.p;
    o.mat_doubles = mat; o.am_threshold = std::sqrt(min_am_prob);
    o.acc_cols = (c->max_len + 4) & ~1;
    if (2 * (size_t)o.acc_cols * 8 > 40000) return fail(c, LB200_ERR_UNSUPPORTED, "LocARNA-P probabilities: sequence B is too long (%d) for the shared-memory accumulators", c->max_len);
    CUDA_TRY(c, cudaEventRecord(c->ev0, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_cursor.p, 0, 4100 * 4, st));
    CUDA_TRY(c, cudaMemsetAsync(c->d_pf_dp.p, 0, std::max<size_t>(R.total_am, 1) * 8, st));   // Dmatprime.fill(0), aligner_p.icc:46-47
    launch_pfo_prepare(R.dc, o, P, std::min(grid, P), st); launches++;
    for (int q = R.q_hi; q >= R.q_lo; q--) { launch_pfo_outside(R.dc, o, q, grid, (int *)c->d_cursor.p, st); launches++; }   // left ends ascending
    launch_pfo_amprob(R.dc, o, P, st); launches++;
    int n_tasks = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&n_tasks, (const int *)c->d_qstart.p + 4096, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    launch_pfo_bm(R.dc, o, n_tasks, P, grid, (int *)c->d_cursor.p + 4098, st); launches += 2;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->ev1, st));
    CUDA_TRY(c, cudaStreamSynchronize(st));
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_kernel_ms = kernel_ms_inside + ms; c->last_launches = launches;
    c->pf_mat_doubles = mat; c->pf_probs_done = true;
    return LB200_OK;
}

int lb200_pair_arcmatch_probs(const lb200_ctx *cc, int pair, double *prob) {
    lb200_ctx *c = const_cast<lb200_ctx *>(cc);
    if (!c || pair < 0 || pair >= (int)c->pairs.size() || !prob) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    if (!c->pf_probs_done || !c->res.valid || pair < c->res.p0 || pair >= c->res.p1) return fail(c, LB200_ERR_STATE, "no LocARNA-P probabilities for this pair (call lb200_run_pf_probs)");
    std::vector<DevArcMatch> am(r.K);
    std::vector<double> dv(r.K);
    if (r.K) {
        CUDA_TRY(c, cudaMemcpy(am.data(), (const DevArcMatch *)c->d_am.p + r.am_base, (size_t)r.K * sizeof(DevArcMatch), cudaMemcpyDeviceToHost));
        CUDA_TRY(c, cudaMemcpy(dv.data(), (const double *)c->d_pf_amp.p + r.am_base, (size_t)r.K * 8, cudaMemcpyDeviceToHost));
    }
    std::vector<int> order;
    reference_am_order(am, order);
    for (size_t k = 0; k < am.size(); k++) prob[k] = dv[am[order[k]].spos];
    return LB200_OK;
}

int lb200_pair_basematch_probs(const lb200_ctx *cc, int pair, double *bm) {
    lb200_ctx *c = const_cast<lb200_ctx *>(cc);
    if (!c || pair < 0 || pair >= (int)c->pairs.size() || !bm) return LB200_ERR_ARG;
    if (!c->pf_probs_done || !c->res.valid || pair < c->res.p0 || pair >= c->res.p1) return fail(c, LB200_ERR_STATE, "no LocARNA-P probabilities for this pair (call lb200_run_pf_probs)");
    const PairRec &r = c->pairs[pair];
    const size_t cells = (size_t)(c->seqs[r.seqA].len + 1) * (c->seqs[r.seqB].len + 1);
    CUDA_TRY(c, cudaMemcpy(bm, (const double *)c->d_pf_mats.p + ((size_t)(pair - c->res.p0) * 5 + 4) * c->pf_mat_doubles, cells * 8, cudaMemcpyDeviceToHost));
    return LB200_OK;
}

int lb200_pair_partition_function(const lb200_ctx *c, int pair, double *Z) {
    if (!c || pair < 0 || pair >= (int)c->pairs.size() || !Z) return LB200_ERR_ARG;
    if (!c->pairs[pair].pf_done) return LB200_ERR_STATE;
    *Z = c->pairs[pair].pf_Z;
    return LB200_OK;
}

int lb200_pair_arcmatch_pf(const lb200_ctx *cc, int pair, double *D) {
    lb200_ctx *c = const_cast<lb200_ctx *>(cc);
    if (!c || pair < 0 || pair >= (int)c->pairs.size() || !D) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    if (!r.pf_done || !c->res.valid || pair < c->res.p0 || pair >= c->res.p1) return fail(c, LB200_ERR_STATE, "no LocARNA-P inside table for this pair (call lb200_run_pf)");
    std::vector<DevArcMatch> am(r.K);
    std::vector<double> dv(r.K);
    if (r.K) {
        CUDA_TRY(c, cudaMemcpy(am.data(), (const DevArcMatch *)c->d_am.p + r.am_base, (size_t)r.K * sizeof(DevArcMatch), cudaMemcpyDeviceToHost));
        CUDA_TRY(c, cudaMemcpy(dv.data(), (const double *)c->d_pf_d.p + r.am_base, (size_t)r.K * 8, cudaMemcpyDeviceToHost));
    }
    std::vector<int> order;
    reference_am_order(am, order);
    for (size_t k = 0; k < am.size(); k++) D[k] = dv[am[order[k]].spos];
    return LB200_OK;
}

int lb200_pair_arcmatches(const lb200_ctx *cc, int pair, int *al, int *ar, int *bl, int *br, int *score, int64_t *D) {
    lb200_ctx *c = const_cast<lb200_ctx *>(cc);
    if (!c || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    std::vector<DevArcMatch> am;
    std::vector<DevEntry> dvals;
    if (c->device == LB200_DEVICE_NONE) {
        if (!r.built) return LB200_ERR_STATE;
        if (D) return fail(c, LB200_ERR_STATE, "no D table on a host-only context");
        am = r.prob.am;
    } else {
        if (!c->res.valid || pair < c->res.p0 || pair >= c->res.p1)
            return fail(c, LB200_ERR_STATE, "the arc-match tables of this pair are not resident (call lb200_upload / lb200_run; large batches are streamed in chunks)");
        am.resize(r.K); dvals.resize(r.K);
        if (r.K) {
            CUDA_TRY(c, cudaMemcpy(am.data(), (const DevArcMatch *)c->d_am.p + r.am_base, (size_t)r.K * sizeof(DevArcMatch), cudaMemcpyDeviceToHost));
            if (D) CUDA_TRY(c, cudaMemcpy(dvals.data(), (const DevEntry *)c->d_ent.p + r.am_base, (size_t)r.K * sizeof(DevEntry), cudaMemcpyDeviceToHost));
        }
    }
    const size_t K = am.size();
    // reference index order = (arc A index, arc B index) ascending (arc_matches.cc:161-183); arcs are indexed by
    // left end descending, right end ascending (basepairs.cc:162-163)
    std::vector<int> order;
    reference_am_order(am, order);
    for (size_t k = 0; k < K; k++) {
        const DevArcMatch &x = am[order[k]];
        if (al) al[k] = x.ends_a & 0xfff;
        if (ar) ar[k] = x.ends_a >> 12;
        if (bl) bl[k] = x.ends_b & 0xfff;
        if (br) br[k] = x.ends_b >> 12;
        // profile pairs on the device hold scores in the gap-free frame (host_model.h ProfileTables): shift back
        long shift_s = 0, shift_d = 0;
        if (r.prof && c->device != LB200_DEVICE_NONE) {
            const ProfileTables &T = *r.prof;
            const int xal = x.ends_a & 0xfff, xar = x.ends_a >> 12, xbl = x.ends_b & 0xfff, xbr = x.ends_b >> 12;
            shift_s = (long)T.gapA[xal] + T.gapA[xar] + T.gapB[xbl] + T.gapB[xbr];
            shift_d = (T.PA[xar] - T.PA[xal - 1]) + (T.PB[xbr] - T.PB[xbl - 1]);
        }
        if (score) score[k] = x.score + (int)shift_s;
        if (D) { const int d = dvals[x.spos].d; D[k] = d < LB_NEG_LIMIT ? LB200_SCORE_NEG_INF : d + shift_d; }
    }
    return LB200_OK;
}

// reference index order = (arc A index, arc B index) ascending (arc_matches.cc:161-183); arcs are indexed by left end
// descending, right end ascending (basepairs.cc:162-163)
static void reference_am_order(const std::vector<DevArcMatch> &am, std::vector<int> &order) {
    const size_t K = am.size();
    order.resize(K);
    for (size_t k = 0; k < K; k++) order[k] = (int)k;
    std::sort(order.begin(), order.end(), [&](int x, int y) {
        const int xal = am[x].ends_a & 0xfff, yal = am[y].ends_a & 0xfff, xar = am[x].ends_a >> 12, yar = am[y].ends_a >> 12;
        if (xal != yal) return xal > yal;
        if (xar != yar) return xar < yar;
        const int xbl = am[x].ends_b & 0xfff, ybl = am[y].ends_b & 0xfff, xbr = am[x].ends_b >> 12, ybr = am[y].ends_b >> 12;
        if (xbl != ybl) return xbl > ybl;
        return xbr < ybr;
    });
}

int lb200_pair_alignment(const lb200_ctx *c, int pair, int *ea, int *eb, char *sa, char *sb) {
    if (!c || pair < 0 || pair >= (int)c->pairs.size()) return LB200_ERR_ARG;
    const PairRec &r = c->pairs[pair];
    if (!r.traced) return LB200_ERR_STATE;
    if (ea) std::copy(r.edges_a.begin(), r.edges_a.end(), ea);
    if (eb) std::copy(r.edges_b.begin(), r.edges_b.end(), eb);
    if (sa) strcpy(sa, r.str_a.c_str());
    if (sb) strcpy(sb, r.str_b.c_str());
    return LB200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
__global__ void reset_d_kernel(DevEntry *ent, uint2 *ent8, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ent[i].d = LB_NEG;
        if (ent8 != nullptr) ent8[i].y = (ent8[i].y & 15u) | ((uint32_t)LB_PACK_NEG << 4);
    }
}
// D view of the modified scoring: D(a, b) - lambda * (arc_length(a) + arc_length(b)) (aligner_impl.hh:235-253); -inf stays -inf
__global__ void adjust_d_kernel(const DevEntry *ent, DevEntry *ent_mod, uint2 *ent8_mod, size_t n, int lambda) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        DevEntry e = ent[i];
        const int al1 = LB_ENT_LO(e.x), bl1 = LB_ENT_HI(e.x), ar = LB_ENT_LO(e.y), br = LB_ENT_HI(e.y);
        if (e.d >= LB_NEG_LIMIT) e.d -= lambda * ((ar - al1) + (br - bl1));
        else e.d = LB_NEG;
        ent_mod[i] = e;
        if (ent8_mod != nullptr) ent8_mod[i] = make_uint2((uint32_t)al1 | ((uint32_t)bl1 << 9) | ((uint32_t)ar << 18) | (((uint32_t)br & 31u) << 27), LB_PACK_W1(br, e.d));
    }
}
cudaError_t lb200_adjust_d(const DevEntry *ent, DevEntry *ent_mod, uint2 *ent8_mod, size_t n, int lambda, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    adjust_d_kernel<<<grid, 256, 0, st>>>(ent, ent_mod, ent8_mod, n, lambda);
    return cudaGetLastError();
}
cudaError_t lb200_reset_d(DevEntry *ent, uint2 *ent8, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    reset_d_kernel<<<grid, 256, 0, st>>>(ent, ent8, n);
    return cudaGetLastError();
}
