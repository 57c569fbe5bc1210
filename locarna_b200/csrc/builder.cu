// Device construction of the arc-match tables of a batch of pairs.
//
// Replaces (paths relative to /root/reference/src/LocARNA):
//   ArcMatches::ArcMatches          arc_matches.cc:130-188   enumeration of valid arc matches
//   ArcMatches::is_valid_arcmatch   arc_matches.cc:19-48     band / length-difference filters
//   init_inner_arc_matchs           arc_matches.cc:50-74     inner arc match (al+1, ar-1, bl+1, br-1)
//   sort_right_adjacency_lists      arc_matches.cc:76-110    lists by common right ends, (al, bl) descending
//   Scoring::arcmatch               scoring.cc:441-485       arc-match score table
//   get_max_right_ends + align_D    arc_matches.cc:313-355, aligner.cc:660-732   one task per left-end pair
//
// Pipeline (one stream, two host syncs for allocation sizes):
//   count   one CTA per pair, one warp per row al, lanes over the band cells (al, bl): number of valid arc matches
//   scan    exclusive prefix sum over all cells of the batch (cub::DeviceScan) -> L-order offsets
//   fill    same enumeration, writes the L-order records (ends, score), the S-order sort key and the per-anti-diagonal histogram
//   inner   inner arc match of every arc match by lookup in the cell (al+1, bl+1)
//   sort    stable radix sort by (pair << 25 | (ar+br) << 12 | ar) (cub::DeviceRadixSort); ties keep the L-order = (al, bl) descending,
//           which is the order of common_right_end_list (arc_matches.hh:188-220)
//   scatter S-order entries + back pointers; sptr = prefix sums of the histogram
//   tasks   one task per cell with arc matches (no-lonely-pairs: with an inner arc match), sorted by level / size
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdint.h>

#include <algorithm>

#include "builder.h"

namespace lb200 {

__device__ __forceinline__ bool valid_cell(const int *lo, const int *hi, int i, int j) { return lo[i] <= j && j <= hi[i]; }
// trace_controller.hh:333-336
__device__ __forceinline__ bool valid_match(const int *lo, const int *hi, int i, int j) {
    return i >= 1 && j >= 1 && valid_cell(lo, hi, i, j) && valid_cell(lo, hi, i - 1, j - 1);
}

struct PairView {
    const int *lo, *hi, *rev;
    const int *alA, *arA, *wA, *lpA, *lcA, *sdA;
    const int *alB, *arB, *wB, *lpB, *lcB, *sdB;
    const uint8_t *cA, *cB;
    const uint8_t *acA, *acB;   // anchor ranks, nullptr without anchor constraints
    const int *psam;            // profile pair: sequence term of the arc-match score per (a, b), row length nB; nullptr: by symbol codes
    int nB;
    int n, m;
    long mdam, mdat;
};

__device__ __forceinline__ PairView view(const BuildCtx &b, const DevPair &p) {
    PairView v;
    v.lo = b.band_lo + p.band; v.hi = b.band_hi + p.band; v.rev = b.cell_rev + p.band;
    v.alA = b.arc_left + p.arcsA; v.arA = b.arc_right + p.arcsA; v.wA = b.arc_weight + p.arcsA;
    v.alB = b.arc_left + p.arcsB; v.arB = b.arc_right + p.arcsB; v.wB = b.arc_weight + p.arcsB;
    v.sdA = b.arc_sdelta + p.arcsA; v.sdB = b.arc_sdelta + p.arcsB;
    v.lpA = b.lptr + p.lptrA; v.lcA = b.lcount + p.lptrA; v.lpB = b.lptr + p.lptrB; v.lcB = b.lcount + p.lptrB;
    v.cA = b.codes + p.codesA; v.cB = b.codes + p.codesB;
    v.acA = p.anchored ? b.acodes + p.codesA : nullptr; v.acB = p.anchored ? b.acodes + p.codesB : nullptr;
    v.n = p.lenA; v.m = p.lenB;
    v.psam = (p.ps_am >= 0 && b.ps_am != nullptr) ? b.ps_am + p.ps_am : nullptr; v.nB = p.n_arcsB;
    const int mx = max(p.lenA, p.lenB);
    v.mdam = b.max_diff_am >= 0 ? b.max_diff_am : mx;      // locarna.cc:617-626
    v.mdat = b.max_diff_at_am >= 0 ? b.max_diff_at_am : mx;
    return v;
}

// rank of the band cell (al, bl) among the cells of its pair, ordered al descending, bl descending
__device__ __forceinline__ int cell_rank(const PairView &v, int al, int bl) { return v.rev[al] + (min(v.hi[al], v.m) - bl); }
__device__ __forceinline__ bool cell_exists(const PairView &v, int al, int bl) {
    return al >= 1 && al <= v.n && bl >= max(v.lo[al], 1) && bl <= min(v.hi[al], v.m);
}

// arc_matches.cc:19-48 for the right ends and the length difference (left ends are tested per cell)
// constraints.allowed_match (arc_matches.cc:27-28) for anchor names that occur in both sequences: equal ranks (0 = unnamed); everything
// else the constraints forbid lies outside the anchor-restricted band
__device__ __forceinline__ bool anchors_allow(const PairView &v, int i, int j) { return v.acA == nullptr || v.acA[i] == v.acB[j]; }
__device__ __forceinline__ bool valid_arcmatch_right(const PairView &v, int al, int ar, int bl, int br) {
    return valid_match(v.lo, v.hi, ar, br) && anchors_allow(v, ar, br) && labs((long)(ar - al) - (long)(br - bl)) <= v.mdam && labs((long)(ar - br)) <= v.mdat;
}

// scoring.cc:441-485 for single sequences
__device__ __forceinline__ int arcmatch_score(const BuildCtx &b, const PairView &v, int a, int bb, int al, int ar, int bl, int br) {
    if (v.psam != nullptr) return v.psam[(size_t)a * v.nB + bb] + v.wA[a] + v.wB[bb];   // profiles: scoring.cc:369-438, averaged on the host
    long seqc = 0;
    if (b.tau != 0) {
        const int c1 = v.cA[al], c2 = v.cA[ar], c3 = v.cB[bl], c4 = v.cB[br];
        if (b.use_ribosum) {
            if ((c1 | c2 | c3 | c4) < 4) return b.am_seq[(c1 * 4 + c2) * 16 + c3 * 4 + c4] + v.wA[a] + v.wB[bb];
        } else seqc = (long)b.sigma8[c1 * LB_NCODES + c3] + b.sigma8[c2 * LB_NCODES + c4];
    }
    return (int)(((long)b.tau * seqc) / 100) + v.wA[a] + v.wB[bb];
}

template <bool FILL>
__global__ void __launch_bounds__(128) enumerate_kernel(BuildCtx b) {
    const DevPair p = b.pairs[blockIdx.x];
    const PairView v = view(b, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int *cell = b.cell_start + p.cell_base;  // COUNT: counts; FILL: exclusive offsets
    for (int al = v.n - warp; al >= 1; al -= 4) {
        const int na = v.lcA[al];
        if (na == 0) continue;
        const int a0 = v.lpA[al];
        const int hi_eff = min(v.hi[al], v.m), lo_eff = max(v.lo[al], 1);
        for (int bl = hi_eff - lane; bl >= lo_eff; bl -= 32) {
            const int nb = v.lcB[bl];
            if (nb == 0 || !valid_match(v.lo, v.hi, al, bl) || !anchors_allow(v, al, bl) || labs((long)(al - bl)) > v.mdat) continue;
            const int b0 = v.lpB[bl];
            const int r = v.rev[al] + (hi_eff - bl);
            int k = 0;
            long long g = 0;
            if (FILL) g = (long long)cell[r];
            for (int a = a0; a < a0 + na; a++) {
                const int ar = v.arA[a];
                for (int bb = b0; bb < b0 + nb; bb++) {
                    const int br = v.arB[bb];
                    if (!valid_arcmatch_right(v, al, ar, bl, br)) continue;
                    if (FILL) {
                        DevArcMatch x;
                        x.ends_a = (uint32_t)al | ((uint32_t)ar << 12);
                        x.ends_b = (uint32_t)bl | ((uint32_t)br << 12);
                        x.score = arcmatch_score(b, v, a, bb, al, ar, bl, br);
                        x.spos = -1; x.inner = -1;
                        // Scoring::arcmatch(am, true): the same sequence term with the stack weights (scoring.cc:478-483)
                        const int da = v.sdA[a], db = v.sdB[bb];
                        x.score_st = (da == LB_NOSTACK || db == LB_NOSTACK) ? LB_NOSTACK : x.score + da + db;
                        b.am[g + k] = x;
                        // S-order: target anti-diagonal ascending, source anti-diagonal (al-1)+(bl-1) descending (see kernels.cu Stream3)
                        b.skeys[g + k] = ((unsigned long long)blockIdx.x << 26) | ((unsigned long long)(ar + br) << 13) | (unsigned long long)(8191 - (al + bl - 2));
                        b.svals[g + k] = (unsigned)(g + k - p.am_base);
                        atomicAdd(&b.sptr[p.sptr + ar + br + 1], 1);
                    }
                    k++;
                }
            }
            if (!FILL) cell[r] = k;
        }
    }
}

// am_base / K of every pair from the scanned cell offsets
__global__ void pair_offsets_kernel(BuildCtx b, int n_pairs) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    DevPair &p = b.pairs[k];
    p.am_base = b.cell_start[p.cell_base];
    p.K = (int)(b.cell_start[p.cell_base + p.n_cells] - b.cell_start[p.cell_base]);
}

__global__ void __launch_bounds__(128) inner_kernel(BuildCtx b) {
    const DevPair p = b.pairs[blockIdx.x];
    const PairView v = view(b, p);
    const int *cell = b.cell_start + p.cell_base;
    DevArcMatch *am = b.am + p.am_base;
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
        const DevArcMatch x = am[k];
        const int al = x.ends_a & 0xfff, ar = x.ends_a >> 12, bl = x.ends_b & 0xfff, br = x.ends_b >> 12;
        int inner = -1;
        if (cell_exists(v, al + 1, bl + 1)) {
            const int r = cell_rank(v, al + 1, bl + 1);
            const int s0 = cell[r] - (int)p.am_base, s1 = cell[r + 1] - (int)p.am_base;
            for (int t = s0; t < s1; t++) {
                const DevArcMatch y = am[t];
                if ((int)(y.ends_a >> 12) == ar - 1 && (int)(y.ends_b >> 12) == br - 1) { inner = t; break; }
            }
        }
        am[k].inner = inner;
    }
}

// sptr[s] = number of arc matches of the pair with ar + br < s (histogram was accumulated at index s + 1)
__global__ void sptr_scan_kernel(BuildCtx b, int n_pairs) {
    const int pk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pk >= n_pairs) return;
    const DevPair p = b.pairs[pk];
    int *sp = b.sptr + p.sptr;
    const int len = p.lenA + p.lenB + 3;
    int carry = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        int x = (i < len) ? sp[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (i < len) sp[i] = x + carry;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
}

// S-order entries and back pointers from the sorted (key, L-order index) pairs
__global__ void scatter_kernel(BuildCtx b, long long total) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int pk = (int)(b.skeys_sorted[t] >> 26);
        const DevPair &p = b.pairs[pk];
        const long long g = p.am_base + b.svals_sorted[t];
        const DevArcMatch x = b.am[g];
        DevEntry e;
        e.x = ((x.ends_a & 0xfff) - 1) | (((x.ends_b & 0xfff) - 1) << 16);
        e.y = (x.ends_a >> 12) | ((x.ends_b >> 12) << 16);
        e.d = LB_NEG;
        e.s = (int)(x.ends_a & 0xfff) + (int)(x.ends_b & 0xfff) - 2;
        b.ent[t] = e;
        if (b.ent8 != nullptr) {
            const uint32_t al1 = (x.ends_a & 0xfff) - 1, bl1 = (x.ends_b & 0xfff) - 1, ar = x.ends_a >> 12, br = x.ends_b >> 12;
            b.ent8[t] = make_uint2(al1 | (bl1 << 9) | (ar << 18) | ((br & 31u) << 27), LB_PACK_W1(br, LB_NEG));
        }
        b.am[g].spos = (int)(t - p.am_base);
    }
}

// One task per cell with arc matches. Sort key: level group (al+bl)>>1 descending, box area descending.
__global__ void __launch_bounds__(128) task_kernel(BuildCtx b) {
    const DevPair p = b.pairs[blockIdx.x];
    const PairView v = view(b, p);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int *cell = b.cell_start + p.cell_base;
    const DevArcMatch *am = b.am;
    const bool nolp = b.no_lonely_pairs != 0;
    unsigned long long n_tasks = 0, cells = 0, terms = 0;
    const int *sp = b.sptr + p.sptr;
    const int s_last = v.n + v.m + 1;
    for (int al = v.n - warp; al >= 1; al -= 4) {
        if (v.lcA[al] == 0) continue;
        const int hi_eff = min(v.hi[al], v.m), lo_eff = max(v.lo[al], 1);
        for (int bl = hi_eff - lane; bl >= lo_eff; bl -= 32) {
            const int r = v.rev[al] + (hi_eff - bl);
            const int s0 = cell[r], s1 = cell[r + 1];
            if (s1 == s0) continue;
            int max_ar = 0, max_br = 0;
            for (int t = s0; t < s1; t++) {
                const DevArcMatch x = am[t];
                if (nolp && x.inner < 0) continue;
                max_ar = max(max_ar, (int)(x.ends_a >> 12)); max_br = max(max_br, (int)(x.ends_b >> 12));
            }
            if (max_ar == 0) continue;
            DevTask tk;
            tk.pair = blockIdx.x;
            tk.al = (short)(nolp ? al + 1 : al); tk.bl = (short)(nolp ? bl + 1 : bl);
            tk.R = (short)(nolp ? max_ar - 2 : max_ar - 1); tk.C = (short)(nolp ? max_br - 2 : max_br - 1);
            tk.run_start = s0 - (int)p.am_base; tk.run_count = s1 - s0;
            const unsigned slot = atomicAdd(b.n_tasks, 1u);
            b.tasks_unsorted[slot] = tk;
            const int group = ((int)tk.al + (int)tk.bl) >> 1;
            const int area = ((int)tk.R - tk.al + 1) * ((int)tk.C - tk.bl + 1);
            b.tkeys[slot] = ((unsigned)(4095 - group) << 20) | (0xfffffu - (unsigned)min(area >> 4, 0xfffff));
            b.tvals[slot] = slot;
            // statistics: cell updates and streamed entries of this task
            int umax = 0;
            for (int i = tk.al + 1; i <= tk.R; i++) {
                const int jl = max((int)tk.bl + 1, v.lo[i]), jh = min((int)tk.C, v.hi[i]);
                if (jh >= jl) { cells += (unsigned long long)(jh - jl + 1) * (b.struct_local ? 4 : 1); umax = (i - tk.al) + (jh - tk.bl); }  // 4 align_noex states
            }
            const int t0 = min(tk.al + tk.bl + 8, s_last), t1 = min(tk.al + tk.bl + umax + 1, s_last);
            if (t1 > t0) terms += sp[t1] - sp[t0];
            n_tasks++;
        }
    }
    // top level box
    if (threadIdx.x == 0) {
        for (int i = 1; i <= v.n; i++) {
            const int jl = max(1, v.lo[i]), jh = min(v.m, v.hi[i]);
            if (jh >= jl) cells += jh - jl + 1;
        }
        terms += sp[s_last] - sp[min(8, s_last)];
    }
    DevPairStats *st = b.stats + blockIdx.x;
    if (n_tasks) atomicAdd((unsigned long long *)&st->n_tasks, n_tasks);
    if (cells) atomicAdd((unsigned long long *)&st->cells, cells);
    if (terms) atomicAdd((unsigned long long *)&st->terms, terms);
}

__global__ void gather_tasks_kernel(BuildCtx b, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) b.tasks[t] = b.tasks_unsorted[b.tvals_sorted[t]];
}

// qstart[q] = first sorted task whose key has level field >= q (q = 4095 - group), q = 0..4096
__global__ void group_bounds_kernel(BuildCtx b, unsigned n) {
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > 4096) return;
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if ((b.tkeys_sorted[mid] >> 20) < q) lo = mid + 1; else hi = mid;
    }
    b.qstart[q] = (int)lo;
}

// Dependency-driven schedule of the D fill (kernels.cu dfill_dep_kernel): the level-sorted tasks are regrouped by blocks of
// sb_pairs pairs (stable, so every block keeps the order level descending, area descending) and every (pair, level group) learns how
// many tasks of its pair lie in higher level groups - the number of completions a task has to wait for.
__global__ void dep_prepare_kernel(BuildCtx b, unsigned n) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const DevTask tk = b.tasks[t];
    b.tkeys[t] = (unsigned)tk.pair / (unsigned)b.sb_pairs;
    b.tvals[t] = t;
    atomicAdd(b.levcnt + (size_t)tk.pair * b.n_groups + (((int)tk.al + (int)tk.bl) >> 1), 1);
}
// levcnt[pair][g] := number of tasks of the pair in level groups > g
__global__ void dep_need_kernel(BuildCtx b, int n_pairs) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int *row = b.levcnt + (size_t)p * b.n_groups;
    int run = 0;
    for (int g = b.n_groups - 1; g >= 0; g--) { const int c = row[g]; row[g] = run; run += c; }
}

// ------------------------------------------------------------------------------------------------ row groups (dfill_rows.cu)
__global__ void group_keys_kernel(GroupBuild g) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n_tasks) return;
    const DevTask tk = g.tasks[t];
    g.keys[t] = ((unsigned long long)tk.pair << 24) | ((unsigned long long)(4095 - tk.al) << 12) | (unsigned long long)(4095 - tk.bl);
    g.vals[t] = t;
}
// a task leads a group if its rank inside its row (pair, al) is a multiple of LB_GV
__global__ void group_leader_kernel(GroupBuild g) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > g.n_tasks) return;
    if (s == g.n_tasks) { g.gid[s] = 0; return; }
    const unsigned long long row = g.keys_sorted[s] >> 12;
    unsigned lo = 0, hi = s;   // first task of the row: lower bound of row << 12
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if ((g.keys_sorted[mid] >> 12) < row) lo = mid + 1; else hi = mid;
    }
    g.gid[s] = ((s - lo) % LB_GV) == 0 ? 1 : 0;
}
__global__ void group_count_kernel(GroupBuild g) { *g.n_groups = g.gid[g.n_tasks]; }
__global__ void group_fill_kernel(GroupBuild g) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.n_tasks || g.gid[s + 1] == g.gid[s]) return;
    const unsigned long long row = g.keys_sorted[s] >> 12;
    // the row: tasks [r0, r1) of the sorted order
    unsigned lo = 0, hi = s;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if ((g.keys_sorted[mid] >> 12) < row) lo = mid + 1; else hi = mid; }
    const unsigned r0 = lo;
    lo = s; hi = g.n_tasks;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if ((g.keys_sorted[mid] >> 12) <= row) lo = mid + 1; else hi = mid; }
    const unsigned r1 = lo;
    DevGroup grp;
    int min_bl = 1 << 20, max_r = 0, max_c = 0;
    for (unsigned k = r0; k < r1; k++) {   // union box of the row (all its groups sweep this geometry)
        const DevTask tk = g.tasks[g.vals_sorted[k]];
        if (k == r0) { grp.pair = tk.pair; grp.al = tk.al; }
        min_bl = min(min_bl, (int)tk.bl); max_r = max(max_r, (int)tk.R); max_c = max(max_c, (int)tk.C);
    }
    int n = 0;
    for (; n < LB_GV && s + n < r1; n++) grp.task[n] = (int)g.vals_sorted[s + n];
    for (int k = n; k < LB_GV; k++) grp.task[k] = -1;
    grp.nmem = (short)n;
    grp.gi = (short)((s - r0) / LB_GV); grp.G = (short)((r1 - r0 + LB_GV - 1) / LB_GV);
    grp.bl0r = (short)min_bl; grp.Rr = (short)max_r; grp.Cr = (short)max_c; grp.pad = 0;
    const int gi = g.gid[s];
    g.groups[gi] = grp;
    // claim order: row descending, the groups of a row consecutively
    g.gkeys[gi] = ((unsigned)(4095 - grp.al) << 20) | (((unsigned)grp.pair & 0xfffu) << 8) | ((unsigned)grp.gi & 0xffu);
    g.gvals[gi] = (unsigned)gi;
    atomicAdd(g.levcnt + (size_t)grp.pair * g.n_levels + grp.al, 1);
}
// levcnt[pair][al] := number of groups of the pair in rows > al
__global__ void group_need_kernel(GroupBuild g) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.n_pairs) return;
    int *row = g.levcnt + (size_t)p * g.n_levels;
    int run = 0;
    for (int a = g.n_levels - 1; a >= 0; a--) { const int c = row[a]; row[a] = run; run += c; }
}

// ------------------------------------------------------------------------------------------------ host entry points
#define TRY(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return e__; } while (0)

cudaError_t builder_count(const BuildCtx &b, int n_pairs, long long total_cells, void *tmp, size_t tmp_bytes, size_t *tmp_need, cudaStream_t st) {
    size_t need = 0;
    TRY(cub::DeviceScan::ExclusiveSum(nullptr, need, b.cell_start, b.cell_start, total_cells + 1, st));
    *tmp_need = need;
    if (tmp == nullptr || tmp_bytes < need) return cudaSuccess;
    TRY(cudaMemsetAsync(b.cell_start, 0, (size_t)(total_cells + 1) * sizeof(int), st));
    enumerate_kernel<false><<<n_pairs, 128, 0, st>>>(b);
    TRY(cub::DeviceScan::ExclusiveSum(tmp, need, b.cell_start, b.cell_start, total_cells + 1, st));
    pair_offsets_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(b, n_pairs);
    return cudaGetLastError();
}

size_t builder_sort_tmp_bytes(long long total_am, int n_pairs) {
    size_t a = 0, c = 0;
    int pair_bits = 1;
    while ((1 << pair_bits) < n_pairs) pair_bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, total_am, 0, 26 + pair_bits);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr, (unsigned *)nullptr,
                                    total_am, 0, 32);
    return a > c ? a : c;
}

cudaError_t builder_fill(const BuildCtx &b, int n_pairs, long long total_am, long long sptr_total, void *tmp, size_t tmp_bytes, cudaStream_t st) {
    TRY(cudaMemsetAsync(b.sptr, 0, (size_t)sptr_total * sizeof(int), st));
    TRY(cudaMemsetAsync(b.stats, 0, (size_t)n_pairs * sizeof(DevPairStats), st));
    TRY(cudaMemsetAsync(b.n_tasks, 0, sizeof(unsigned), st));
    enumerate_kernel<true><<<n_pairs, 128, 0, st>>>(b);
    inner_kernel<<<n_pairs, 128, 0, st>>>(b);
    sptr_scan_kernel<<<(n_pairs + 3) / 4, 128, 0, st>>>(b, n_pairs);
    if (total_am > 0) {
        int pair_bits = 1;
        while ((1 << pair_bits) < n_pairs) pair_bits++;
        size_t need = tmp_bytes;
        TRY(cub::DeviceRadixSort::SortPairs(tmp, need, b.skeys, b.skeys_sorted, b.svals, b.svals_sorted, total_am, 0, 26 + pair_bits, st));
        const int grid = (int)((total_am + 255) / 256 < 148 * 16 ? (total_am + 255) / 256 : 148 * 16);
        scatter_kernel<<<grid, 256, 0, st>>>(b, total_am);
    }
    task_kernel<<<n_pairs, 128, 0, st>>>(b);
    return cudaGetLastError();
}

cudaError_t builder_sort_tasks(const BuildCtx &b, int n_pairs, unsigned n_tasks, void *tmp, size_t tmp_bytes, cudaStream_t st) {
    if (n_tasks > 0) {
        size_t need = tmp_bytes;
        TRY(cub::DeviceRadixSort::SortPairs(tmp, need, b.tkeys, b.tkeys_sorted, b.tvals, b.tvals_sorted, (long long)n_tasks, 0, 32, st));
        gather_tasks_kernel<<<(n_tasks + 255) / 256, 256, 0, st>>>(b, n_tasks);
    }
    group_bounds_kernel<<<(4097 + 255) / 256, 256, 0, st>>>(b, n_tasks);
    if (b.levcnt != nullptr) {
        TRY(cudaMemsetAsync(b.levcnt, 0, (size_t)n_pairs * b.n_groups * sizeof(int), st));
        if (n_tasks > 0) {
            dep_prepare_kernel<<<(n_tasks + 255) / 256, 256, 0, st>>>(b, n_tasks);
            if (b.sb_pairs >= n_pairs) {   // one block: the level order is the claim order
                TRY(cudaMemcpyAsync(b.tvals_sorted, b.tvals, (size_t)n_tasks * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
            } else {
                int sb_bits = 1;
                while ((1 << sb_bits) < (n_pairs + b.sb_pairs - 1) / b.sb_pairs) sb_bits++;
                size_t need = tmp_bytes;
                TRY(cub::DeviceRadixSort::SortPairs(tmp, need, b.tkeys, b.tkeys_sorted, b.tvals, b.tvals_sorted, (long long)n_tasks, 0, sb_bits, st));
            }
            dep_need_kernel<<<(n_pairs + 127) / 128, 128, 0, st>>>(b, n_pairs);
        }
    }
    return cudaGetLastError();
}

size_t builder_groups_tmp_bytes(unsigned n_tasks, int n_pairs) {
    size_t a = 0, b = 0, c = 0;
    int pair_bits = 1;
    while ((1 << pair_bits) < n_pairs) pair_bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, (long long)n_tasks, 0, 24 + pair_bits);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (int *)nullptr, (int *)nullptr, (long long)n_tasks + 1);
    cub::DeviceRadixSort::SortPairs(nullptr, c, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr, (unsigned *)nullptr,
                                    (long long)n_tasks, 0, 32);
    return std::max(a, std::max(b, c));
}

cudaError_t builder_groups_scan(const GroupBuild &g, void *tmp, size_t tmp_bytes, cudaStream_t st) {
    if (g.n_tasks == 0) return cudaMemsetAsync(g.n_groups, 0, sizeof(int), st);
    int pair_bits = 1;
    while ((1 << pair_bits) < g.n_pairs) pair_bits++;
    const unsigned blocks = (g.n_tasks + 1 + 255) / 256;
    group_keys_kernel<<<blocks, 256, 0, st>>>(g);
    size_t need = tmp_bytes;
    TRY(cub::DeviceRadixSort::SortPairs(tmp, need, g.keys, g.keys_sorted, g.vals, g.vals_sorted, (long long)g.n_tasks, 0, 24 + pair_bits, st));
    group_leader_kernel<<<blocks, 256, 0, st>>>(g);
    need = tmp_bytes;
    TRY(cub::DeviceScan::ExclusiveSum(tmp, need, g.gid, g.gid, (long long)g.n_tasks + 1, st));
    group_count_kernel<<<1, 1, 0, st>>>(g);
    return cudaGetLastError();
}

cudaError_t builder_groups_fill(const GroupBuild &g, unsigned n_groups, void *tmp, size_t tmp_bytes, cudaStream_t st) {
    TRY(cudaMemsetAsync(g.levcnt, 0, (size_t)g.n_pairs * g.n_levels * sizeof(int), st));
    if (g.n_tasks == 0 || n_groups == 0) return cudaSuccess;
    group_fill_kernel<<<(g.n_tasks + 255) / 256, 256, 0, st>>>(g);
    size_t need = tmp_bytes;
    TRY(cub::DeviceRadixSort::SortPairs(tmp, need, g.gkeys, g.gkeys_sorted, g.gvals, g.order, (long long)n_groups, 0, 32, st));
    group_need_kernel<<<(g.n_pairs + 127) / 128, 128, 0, st>>>(g);
    return cudaGetLastError();
}

}  // namespace lb200
