// Probability envelope of the band on the device.
//
// Reference: PFGotoh / PFTraceProbs (src/LocARNA/edge_probs.icc:7-268), StralScore::sigma (stral_score.cc:29-60),
// TraceController::restrict_by_trace_probabilities (trace_controller.cc:565-597), call site main_helper.icc:371-426.
// `locarna` evaluates this sequence-only partition function in 80-bit long double (locarna.cc:384-391); the GPU has no
// such type. The kernel evaluates the same recursions, in the same operation order, in FP64 and *screens*: rows get
// their new [min_col, max_col] from the FP64 probabilities, and a pair is flagged "uncertain" when any cell lies within a
// relative margin of 1e-9 of the threshold (FP64 error of these all-positive sums is < 1e-12) or when the partition
// function leaves the FP64 range. Flagged pairs are recomputed on the host in long double (host_model.cc), so the bands
// are always exactly the reference's; nothing is approximated silently.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <stdlib.h>

#include "dev_ctx.h"
#include "envelope.h"

namespace lb200 {

// two warps per pass, eight CTAs per SM: the sweep is bound by the latency of its global loads (profiles/r1d_envelope_ncu_summary.txt),
// so pairs in flight matter more than threads per pair
#define ENV_THREADS 128

struct EnvSeq {
    const uint8_t *code; const double *up, *down; int len; bool rev;
    __device__ int c(int i) const { return code[rev ? len + 1 - i : i]; }
    __device__ double u(int i) const { return rev ? down[len + 1 - i] : up[i]; }   // reversed: up/down swap (stral_score.cc:82-98)
    __device__ double d(int i) const { return rev ? up[len + 1 - i] : down[i]; }
};

// Layout of the (n+1) x (m+1) partition-function matrices: anti-diagonal major. Cell (i, j) lives at off(i+j) + i - max(0, i+j-m),
// off(d) = number of cells on the anti-diagonals before d. The sweep handles one anti-diagonal per step and reads the two before
// it, so every access of a step is to consecutive doubles (row-major storage put each cell of a step into its own 32-byte sector).
struct DiagIdx {
    int n, m, a, b;
    size_t total;
    __device__ DiagIdx(int n_, int m_) : n(n_), m(m_), a(min(n_, m_)), b(max(n_, m_)), total((size_t)(n_ + 1) * (m_ + 1)) {}
    __device__ size_t off(int d) const {
        if (d <= a) return (size_t)d * (d + 1) / 2;
        if (d <= b) return (size_t)a * (a + 1) / 2 + (size_t)(d - a) * (a + 1);
        const size_t r = (size_t)(n + m - d + 1);
        return total - r * (r + 1) / 2;
    }
    __device__ size_t operator()(int i, int j) const { const int d = i + j; return off(d) + (size_t)(i - max(0, d - m)); }
};

// one pass of pf_gotoh (edge_probs.icc:76-162) over the band [lo, hi] by anti-diagonals; matrices are (n+1) x (m+1) row major.
// The forward and the reversed pass are independent and have the same anti-diagonal count, so each half of the CTA (nt threads,
// thread index tid) runs one of them through this one call site; the block-wide barriers are shared.
__device__ void gotoh_pass(double *zM, double *zA, double *zB, const int *lo, const int *hi, bool band_rev, int n, int m, const EnvSeq &A,
                           const EnvSeq &B, const EnvCtx &e, bool free_left1, bool free_left2, int tid, int nt) {
    const DiagIdx at(n, m);
    const int total = (n + 1) * (m + 1);
    const double g_open = exp(e.open / e.temp), g_ext = exp(e.ext / e.temp);
    auto LO = [&](int i) { return band_rev ? m - hi[n - i] : lo[i]; };   // trace_controller.cc:319-338
    auto HI = [&](int i) { return band_rev ? m - lo[n - i] : hi[i]; };
    auto valid = [&](int i, int j) { return LO(i) <= j && j <= HI(i); };
    for (int k = tid; k < total; k += nt) { zM[k] = 0; zA[k] = 0; zB[k] = 0; }
    __syncthreads();
    if (tid == 0) {
        if (valid(0, 0)) zM[0] = e.local ? 0 : 1;
        if (n > 0 && valid(1, 0)) zA[at(1, 0)] = g_open * g_ext;
        if (m > 0 && valid(0, 1)) zB[at(0, 1)] = g_open * g_ext;
        for (int i = 2; i <= n; i++) { if (LO(i) > 0) break; zA[at(i, 0)] = zA[at(i - 1, 0)] * g_ext; }
        for (int j = max(LO(0), 2); j <= min(HI(0), m); j++) zB[at(0, j)] = zB[at(0, j - 1)] * g_ext;
        if (free_left2) for (int i = 1; i <= n; i++) zA[at(i, 0)] += 1;
        if (free_left1) for (int j = 1; j <= m; j++) zB[at(0, j)] += 1;
    }
    __syncthreads();
    // rows of the band on anti-diagonal d: i + LO(i) and i + HI(i) grow strictly with i (make_band's rows are monotone,
    // trace_controller.cc:522-531), so they form one range [imin, imax] whose ends only move forward with d
    int imin = 0, imax = 0;
    for (int d = 2; d <= n + m; d++) {
        while (imin <= n && imin + HI(imin) < d) imin++;
        while (imax < n && imax + 1 + LO(imax + 1) <= d) imax++;
        // diagonal d, d-1, d-2: cell (i, j) / (i-1, j) and (i, j-1) / (i-1, j-1)
        const size_t o0 = at.off(d) - max(0, d - m), o1 = at.off(d - 1) - max(0, d - 1 - m), o2 = at.off(d - 2) - max(0, d - 2 - m);
        for (int i = max(max(1, d - m), imin) + tid; i <= min(min(n, d - 1), imax); i += nt) {
            const int j = d - i;
            if (j >= max(LO(i), 1) && j <= min(HI(i), m)) {
                // StralScore::sigma
                double seq_score = 0;
                const int a = A.c(i), b = B.c(j);
                if (a < 4 && b < 4) seq_score = e.bm[a * 4 + b];
                const double s = e.sw * (sqrt(A.d(i) * B.d(j)) + sqrt(A.u(i) * B.u(j))) + seq_score;
                const double mt = exp(s / e.temp);
                const size_t p = o0 + i, pd = o2 + (i - 1), pu = o1 + (i - 1), pl = o1 + i;
                zM[p] = zM[pd] * mt + zA[pd] * mt + zB[pd] * mt + (e.local ? mt : 0);
                zA[p] = zA[pu] * g_ext + zM[pu] * g_open * g_ext + zB[pu] * g_open * g_ext;
                zB[p] = zB[pl] * g_ext + zM[pl] * g_open * g_ext + zA[pl] * g_open * g_ext;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(ENV_THREADS, 8) envelope_kernel(EnvCtx e, int n_pairs, int *cursor) {
    __shared__ int s_pair;
    __shared__ double s_red[ENV_THREADS];
    __shared__ int s_flag;
    double *base = e.scratch + (size_t)blockIdx.x * e.scratch_doubles;
    for (;;) {
        if (threadIdx.x == 0) s_pair = atomicAdd(cursor, 1);
        __syncthreads();
        const int pk = s_pair;
        __syncthreads();
        if (pk >= n_pairs) break;
        const EnvPair pr = e.pairs[pk];
        const int n = pr.lenA, m = pr.lenB;
        const size_t sz = (size_t)(n + 1) * (m + 1);
        const DiagIdx at(n, m);
        double *zM = base, *zA = base + sz, *zB = base + 2 * sz, *zMr = base + 3 * sz, *zAr = base + 4 * sz, *zBr = base + 5 * sz;
        int *lo = e.band_lo + pr.band, *hi = e.band_hi + pr.band;
        EnvSeq A, B;
        A.code = e.codes + pr.codesA; A.up = e.p_up + pr.probA; A.down = e.p_down + pr.probA; A.len = n; A.rev = false;
        B.code = e.codes + pr.codesB; B.up = e.p_up + pr.probB; B.down = e.p_down + pr.probB; B.len = m; B.rev = false;
        const bool rev = threadIdx.x >= ENV_THREADS / 2;   // first half: forward pass, second half: pass over the reversed sequences
        A.rev = rev; B.rev = rev;
        gotoh_pass(rev ? zMr : zM, rev ? zAr : zA, rev ? zBr : zB, lo, hi, rev, n, m, A, B, e, rev ? e.fe_right1 : e.fe_left1,
                   rev ? e.fe_right2 : e.fe_left2, (int)(threadIdx.x & (ENV_THREADS / 2 - 1)), ENV_THREADS / 2);   // FreeEndgaps::reverse (free_endgaps.hh:74-79)
        // partition function z (edge_probs.icc:34-59)
        double z;
        if (e.local) {
            double acc = 0;  // the reference sums row major starting from 1; the order only matters below the screening margin
            for (size_t k = threadIdx.x; k < sz; k += blockDim.x) acc += zM[k];
            s_red[threadIdx.x] = acc;
            __syncthreads();
            for (int o = ENV_THREADS / 2; o; o >>= 1) { if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o]; __syncthreads(); }
            z = 1 + s_red[0];
        } else {
            z = zM[sz - 1] + zA[sz - 1] + zB[sz - 1];
            if (e.fe_left2) for (int i = 0; i <= n; i++) z += zA[at(i, m)];
            if (e.fe_left1) for (int j = 0; j <= m; j++) z += zB[at(n, j)];
        }
        if (threadIdx.x == 0) s_flag = (!isfinite(z) || z <= 0) ? 1 : 0;
        __syncthreads();
        // trace probabilities and the new row ranges (edge_probs.icc:243-264, trace_controller.cc:567-583)
        const double locality_add = e.local ? 1 : 0, g_open = exp(e.open / e.temp);
        const double thr = e.min_prob, margin = 1e-9 * e.min_prob;
        int *nlo = e.out_lo + pr.band, *nhi = e.out_hi + pr.band;
        for (int i = threadIdx.x; i <= n; i += blockDim.x) {
            int new_min = hi[i], new_max = lo[i];
            bool unsure = false;
            for (int j = max(lo[i], 0); j <= min(hi[i], m); j++) {
                const size_t p = at(i, j), r = at(n - i, m - j);
                const double zij = zM[p] * (zMr[r] + zAr[r] + zBr[r] + locality_add) + zA[p] * (zMr[r] + zAr[r] / g_open + zBr[r]) +
                                   zB[p] * (zMr[r] + zAr[r] + zBr[r] / g_open);
                const double prob = zij / z;
                if (!(fabs(prob - thr) > margin)) unsure = true;  // also catches NaN
                if (prob >= thr) { new_min = min(new_min, j); new_max = max(new_max, j); }
            }
            nlo[i] = max(lo[i], new_min);
            nhi[i] = min(hi[i], new_max);
            if (unsure) s_flag = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            // monotone closure (trace_controller.cc:585-596)
            int run = 0;
            for (int i = 0; i <= n; i++) { nhi[i] = max(nhi[i], run); run = nhi[i]; }
            run = nhi[n];
            for (int i = n; i >= 0; i--) { nlo[i] = min(nlo[i], run); run = nlo[i]; }
            e.out_flag[pk] = s_flag;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Second version (round 2): the sweep keeps the last two anti-diagonals of its three matrices in SHARED memory (rotating sets indexed
// by the row), so a step reads no global memory; the forward matrices are written once (anti-diagonal major, every cell of a
// diagonal, zeros outside the band: no zero fill), and the pass over the reversed sequences is never stored at all: cell (i', j') of
// the reversed problem is the suffix term of cell (n - i', m - j'), so its trace probability (edge_probs.icc:243-264) is formed the
// moment the reverse sweep computes it, from the stored forward cell, and folded into the row's new [min_col, max_col] with
// shared-memory atomics. Per pair this moves 3 (n+1)(m+1) doubles out and in once (the first version moved about 5 times as much: zero
// fill, six matrices, neighbour reads from L2 / DRAM, final pass) - see profiles/r1d_envelope_ncu_summary.txt vs r2 in DESIGN.md.
struct EnvSmem {
    double *zb;        // [rotating set][M, A, B][row], rows = stride
    int stride;
    int *rmin, *rmax;
    double *red;
    double *sad, *sau, *sbd, *sbu;   // square roots of the pairing probabilities per position of A / B as the current pass sees them
};

template <bool REV>
__device__ void env_sweep(const EnvCtx &e, const EnvPair &pr, const EnvSmem &sm, double *fM, double *fA, double *fB, double z_total, double *local_sum,
                          int *s_flag) {
    const int n = pr.lenA, m = pr.lenB, tid = threadIdx.x, nt = blockDim.x;
    const DiagIdx at(n, m);
    const int *lo = e.band_lo + pr.band, *hi = e.band_hi + pr.band;
    EnvSeq A, B;
    A.code = e.codes + pr.codesA; A.up = e.p_up + pr.probA; A.down = e.p_down + pr.probA; A.len = n; A.rev = REV;
    B.code = e.codes + pr.codesB; B.up = e.p_up + pr.probB; B.down = e.p_down + pr.probB; B.len = m; B.rev = REV;
    const bool free_left1 = REV ? e.fe_right1 : e.fe_left1, free_left2 = REV ? e.fe_right2 : e.fe_left2;   // FreeEndgaps::reverse (free_endgaps.hh:74-79)
    const double g_open = exp(e.open / e.temp), g_ext = exp(e.ext / e.temp), log_ext = e.ext / e.temp;
    const double locality_add = e.local ? 1 : 0, thr = e.min_prob, margin = 1e-9 * e.min_prob;
    auto LO = [&](int i) { return REV ? m - hi[n - i] : lo[i]; };   // trace_controller.cc:319-338
    auto HI = [&](int i) { return REV ? m - lo[n - i] : hi[i]; };
    // borders (edge_probs.icc:88-113): column 0 carries g_open g_ext^i while the band contains it, row 0 likewise inside [LO(0), HI(0)]
    int ia_max = 0;   // last row whose band starts at column 0, contiguous from row 1
    if (n > 0 && LO(1) <= 0) { ia_max = 1; while (ia_max < n && LO(ia_max + 1) <= 0) ia_max++; }
    const int jb_max = min(HI(0), m);
    const bool v00 = LO(0) <= 0 && 0 <= HI(0);
    // sqrt(pA pB) = sqrt(pA) sqrt(pB): the roots are taken once per position and pass instead of twice per cell (the last bits differ
    // from the reference's expression; the screening margin covers that, and flagged pairs are redone on the host anyway)
    for (int i = 1 + tid; i <= n; i += nt) { sm.sad[i] = sqrt(A.d(i)); sm.sau[i] = sqrt(A.u(i)); }
    for (int j = 1 + tid; j <= m; j += nt) { sm.sbd[j] = sqrt(B.d(j)); sm.sbu[j] = sqrt(B.u(j)); }
    __syncthreads();
    double lsum = 0;
    bool unsure = false;
    auto cell_out = [&](int i, int j, double zm, double za, double zb) {
        if (!REV) {
            const size_t p = at(i, j);
            fM[p] = zm; fA[p] = za; fB[p] = zb;
            lsum += zm;
        } else {
            // (i, j) of the reversed problem = suffix of cell (n - i, m - j); band test in forward coordinates
            const int fi = n - i, fj = m - j;
            if (fj < max(lo[fi], 0) || fj > min(hi[fi], m)) return;
            const size_t p = at(fi, fj);
            const double a = fM[p], b = fA[p], c = fB[p];
            const double zij = a * (zm + za + zb + locality_add) + b * (zm + za / g_open + zb) + c * (zm + za + zb / g_open);
            const double prob = zij / z_total;
            if (!(fabs(prob - thr) > margin)) unsure = true;  // also catches NaN
            if (prob >= thr) { atomicMin(sm.rmin + fi, fj); atomicMax(sm.rmax + fi, fj); }
        }
    };
    for (int d = 0; d <= n + m; d++) {
        const int st = sm.stride;
        double *cM = sm.zb + (size_t)(d % 3) * 3 * st, *cA = cM + st, *cB = cA + st;
        const double *pM = sm.zb + (size_t)((d + 2) % 3) * 3 * st, *pA = pM + st, *pB = pA + st;     // d - 1
        const double *qM = sm.zb + (size_t)((d + 1) % 3) * 3 * st, *qA = qM + st, *qB = qA + st;     // d - 2
        for (int i = max(0, d - m) + tid; i <= min(n, d); i += nt) {
            const int j = d - i;
            double zm = 0, za = 0, zb = 0;
            if (i == 0 && j == 0) { zm = (v00 && !e.local) ? 1 : 0; }
            else if (j == 0) { za = (i <= ia_max ? g_open * exp(log_ext * i) : 0) + (free_left2 ? 1 : 0); }
            else if (i == 0) {
                // row 0: zB(0, 1) = g_open g_ext if valid, then products along [max(LO(0), 2), jb_max] from whatever (0, j - 1) holds
                const bool v01 = LO(0) <= 1 && 1 <= HI(0);
                const int jstart = max(LO(0), 2);
                double v = 0;
                if (j == 1) v = v01 ? g_open * g_ext : 0;
                else if (j >= jstart && j <= jb_max) v = (v01 && jstart == 2) ? g_open * exp(log_ext * j) : 0;
                zb = v + (free_left1 ? 1 : 0);
            } else if (j >= max(LO(i), 1) && j <= min(HI(i), m)) {
                double seq_score = 0;   // StralScore::sigma (stral_score.cc:29-60)
                const int a = A.c(i), b = B.c(j);
                if (a < 4 && b < 4) seq_score = e.bm[a * 4 + b];
                const double s = e.sw * (sm.sad[i] * sm.sbd[j] + sm.sau[i] * sm.sbu[j]) + seq_score;
                const double mt = exp(s / e.temp);
                zm = qM[i - 1] * mt + qA[i - 1] * mt + qB[i - 1] * mt + (e.local ? mt : 0);
                za = pA[i - 1] * g_ext + pM[i - 1] * g_open * g_ext + pB[i - 1] * g_open * g_ext;
                zb = pB[i] * g_ext + pM[i] * g_open * g_ext + pA[i] * g_open * g_ext;
            }
            cM[i] = zm; cA[i] = za; cB[i] = zb;
            cell_out(i, j, zm, za, zb);
        }
        __syncthreads();
    }
    if (!REV) *local_sum = lsum;
    if (REV && unsure) *s_flag = 1;
}

__global__ void __launch_bounds__(ENV_THREADS) envelope_kernel_v2(EnvCtx e, int n_pairs, int *cursor, int max_rows, int max_cols) {
    extern __shared__ __align__(16) unsigned char env_smem[];
    __shared__ int s_pair, s_flag;
    EnvSmem sm;
    {
        double *p = (double *)env_smem;
        sm.zb = p; sm.stride = max_rows; p += 9 * (size_t)max_rows;
        sm.red = p; p += ENV_THREADS;
        sm.sad = p; p += max_rows; sm.sau = p; p += max_rows; sm.sbd = p; p += max_cols; sm.sbu = p; p += max_cols;
        sm.rmin = (int *)p; sm.rmax = sm.rmin + max_rows;
    }
    double *base = e.scratch + (size_t)blockIdx.x * e.scratch_doubles;
    for (;;) {
        if (threadIdx.x == 0) { s_pair = atomicAdd(cursor, 1); s_flag = 0; }
        __syncthreads();
        const int pk = s_pair;
        __syncthreads();
        if (pk >= n_pairs) break;
        const EnvPair pr = e.pairs[pk];
        const int n = pr.lenA, m = pr.lenB;
        const size_t sz = (size_t)(n + 1) * (m + 1);
        const DiagIdx at(n, m);
        double *fM = base, *fA = base + sz, *fB = base + 2 * sz;
        const int *lo = e.band_lo + pr.band, *hi = e.band_hi + pr.band;
        for (int i = threadIdx.x; i <= n; i += blockDim.x) { sm.rmin[i] = hi[i]; sm.rmax[i] = lo[i]; }
        double lsum = 0;
        env_sweep<false>(e, pr, sm, fM, fA, fB, 1.0, &lsum, &s_flag);
        // partition function z (edge_probs.icc:34-59)
        double z;
        if (e.local) {
            sm.red[threadIdx.x] = lsum;
            __syncthreads();
            for (int o = ENV_THREADS / 2; o; o >>= 1) { if ((int)threadIdx.x < o) sm.red[threadIdx.x] += sm.red[threadIdx.x + o]; __syncthreads(); }
            z = 1 + sm.red[0];
            __syncthreads();
        } else {
            __threadfence_block();
            z = fM[sz - 1] + fA[sz - 1] + fB[sz - 1];
            if (e.fe_left2) for (int i = 0; i <= n; i++) z += fA[at(i, m)];
            if (e.fe_left1) for (int j = 0; j <= m; j++) z += fB[at(n, j)];
        }
        if (threadIdx.x == 0 && (!isfinite(z) || z <= 0)) s_flag = 1;
        env_sweep<true>(e, pr, sm, fM, fA, fB, z, nullptr, &s_flag);
        __syncthreads();
        int *nlo = e.out_lo + pr.band, *nhi = e.out_hi + pr.band;
        for (int i = threadIdx.x; i <= n; i += blockDim.x) { nlo[i] = max(lo[i], sm.rmin[i]); nhi[i] = min(hi[i], sm.rmax[i]); }
        __syncthreads();
        if (threadIdx.x == 0) {
            // monotone closure (trace_controller.cc:585-596)
            int run = 0;
            for (int i = 0; i <= n; i++) { nhi[i] = max(nhi[i], run); run = nhi[i]; }
            run = nhi[n];
            for (int i = n; i >= 0; i--) { nlo[i] = min(nlo[i], run); run = nlo[i]; }
            e.out_flag[pk] = s_flag;
        }
        __syncthreads();
    }
}

static bool envelope_use_v2(int max_rows, int max_cols) { return envelope_smem_bytes(max_rows, max_cols) <= 200 * 1024 && getenv("LB200_ENVELOPE_V1") == nullptr; }
size_t envelope_scratch_doubles(size_t max_cells, int max_rows, int max_cols) { return (envelope_use_v2(max_rows, max_cols) ? 3 : 6) * max_cells; }
int envelope_smem_bytes(int max_rows, int max_cols) { return (11 * max_rows + 2 * max_cols + ENV_THREADS) * 8 + 2 * max_rows * 4; }

cudaError_t launch_envelope(const EnvCtx &e, int n_pairs, int grid, int *cursor, int max_rows, int max_cols, cudaStream_t st) {
    cudaError_t err = cudaMemsetAsync(cursor, 0, sizeof(int), st);
    if (err != cudaSuccess) return err;
    const int smem = envelope_smem_bytes(max_rows, max_cols);
    if (envelope_use_v2(max_rows, max_cols)) {
        err = cudaFuncSetAttribute(envelope_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, lb200_sticky_smem(400, smem));
        if (err != cudaSuccess) return err;
        envelope_kernel_v2<<<grid, ENV_THREADS, smem, st>>>(e, n_pairs, cursor, max_rows, max_cols);
    } else {
        envelope_kernel<<<grid, ENV_THREADS, 0, st>>>(e, n_pairs, cursor);   // very long sequences: matrices in global memory only
    }
    return cudaGetLastError();
}

}  // namespace lb200
