// Hand-written sm_100a kernels for the LocARNA pairwise alignment hot path.
//
// Reference path (file:line relative to /root/reference/src/LocARNA):
//   align_noex            aligner.cc:153-233   cell recurrence (M/E/F + arc-match term)
//   init_state            aligner.cc:264-368   box borders and -inf guards around the band
//   align_in_arcmatch     aligner.cc:373-565   fill of one M box
//   fill_D_entries[_noLP] aligner.cc:574-657   D(arcA,arcB) from the box
//   align_D               aligner.cc:660-732   schedule over left-end pairs
//   align_top_level_*     aligner.cc:736-880   top level box + score
//
// Design: one warp sweeps one M box by anti-diagonals u = i'+j'. Lanes own pairs of
// adjacent diagonals v = j'-i' (only diagonals of the parity of u are active in a step), so every
// lane has work in every step of a diagonal band, M/E/F of the three neighbours stay in registers and
// only the two cells at a lane boundary travel by warp shuffle. The recurrence uses the DPX
// instructions (VIADDMNMX / VIMNMX3). The box itself is written once per cell (shared memory when it
// fits, an L2-resident scratch otherwise) because arc-match terms read M(al'-1, bl'-1) at arbitrary
// earlier cells. Arc-match terms are not pulled per cell: the valid arc matches of a pair are kept
// sorted by the anti-diagonal of their right ends (S-order), the warp streams the entries four
// anti-diagonals ahead with coalesced loads and folds M(src)+D into a small ring of per-diagonal
// accumulators with shared-memory atomicMax.
#include <cuda_runtime.h>
#include <stdint.h>
#include "dev_types.h"
#include "dev_ctx.h"

namespace lb200 {

__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }  // max(a+b, c)
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// how the first row / column of a box are initialised (aligner.cc:295-301, :338-343)
struct BoxInit {
    int col_base, col_step;  // M(al+i', bl) = col_base + i'*col_step   (deleting a prefix of A)
    int row_base, row_step;  // M(al, bl+j') = row_base + j'*row_step   (inserting a prefix of B)
    int clamp0;              // sequence-local top level: M = max(M, 0)  (aligner.cc:866-868)
};

struct BoxGeom {
    int al, bl, Rn, Cn;      // origin and local extent (rows 0..Rn, cols 0..Cn; row/col 0 are borders)
    int vmin, strideD, umax;
    int nslots;              // number of diagonal pairs
};

constexpr int RING = 8;      // arc-term accumulators for 8 consecutive anti-diagonals
constexpr int LOOKAHEAD = 4; // entries of anti-diagonal u+4 are folded in step u; their sources are final:
                             // an arc spans >= 3 positions, so src = (al'-1, bl'-1) lies >= 8 anti-diagonals back

// per-warp shared memory carve-up
struct WarpSmem {
    uint32_t *rowinfo;  // Rn+1 words: jl | jh<<12 | codeA<<24
    uint8_t *colinfo;   // Cn+1 bytes: codeB
    int *arcbuf;        // RING * 32*NC
    int *box;           // smem box (or nullptr when the box lives in scratch)
};

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Stage the per-row band/sequence info of a box into shared memory and derive its geometry.
__device__ void setup_box(const DevCtx &c, const DevPair &pr, int al, int bl, int R, int C, BoxGeom &g, uint32_t *rowinfo,
                          uint8_t *colinfo) {
    const int lane = threadIdx.x & 31;
    g.al = al; g.bl = bl; g.Rn = R - al; g.Cn = C - bl;
    int vmin = 1 << 20, vmax = -(1 << 20), umax = 0;
    const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
    const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
    for (int ip = lane; ip <= g.Rn; ip += 32) {
        int i = al + ip;
        int l = lo[i], h = hi[i];
        int jl = (ip == 0 || l <= bl) ? 0 : (l - bl);
        int jh = min(g.Cn, h - bl);
        if (jh < jl) { jl = 1; jh = 0; }
        else { vmin = min(vmin, jl - ip); vmax = max(vmax, jh - ip); umax = max(umax, ip + jh); }
        uint32_t code = (i >= 1) ? ca[i] : 0;
        rowinfo[ip] = (uint32_t)jl | ((uint32_t)jh << 12) | (code << 24);
    }
    for (int jp = lane; jp <= g.Cn; jp += 32) colinfo[jp] = (bl + jp >= 1) ? cb[bl + jp] : 0;
    g.vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    g.umax = warp_max(umax);
    int wd = vmax - g.vmin + 1;
    g.nslots = (wd + 1) >> 1;
    g.strideD = wd | 1;
    __syncwarp();
}

// Fill one M box. NC = diagonal pairs per lane (the warp covers 64*NC diagonals).
template <int NC>
__device__ void fill_box(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const BoxInit &init, const WarpSmem &ws,
                         int *box, const int *sig) {
    const int lane = threadIdx.x & 31;
    constexpr int NW = 32 * NC;
    const int gap = c.params.gap, gap_open = c.params.gap_open;
    const DevEntry *ent = c.ent + pr.am_base;
    const int *dval = c.dval + pr.am_base;
    const int *sptr = c.sptr + pr.sptr;

    for (int k = lane; k < RING * NW; k += 32) ws.arcbuf[k] = LB_NEG;
    __syncwarp();

    // register state per diagonal pair: the cell last computed on the even / odd diagonal of the pair
    int mE[NC], eE[NC], fE[NC], mO[NC], eO[NC], fO[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) { mE[k] = eE[k] = fE[k] = mO[k] = eO[k] = fO[k] = LB_NEG; }

    // u runs over anti-diagonals; the parity of (u - vmin) selects which diagonal of each pair is active.
    // par0 = parity of u=0: diagonals c = 2g + par
    const int s_base = g.al + g.bl;
    for (int u = 0; u <= g.umax; u++) {
        const int par = (u - g.vmin) & 1;
        // ---- neighbour exchange across lane boundaries
        int xm, xo;  // par==0: left neighbour's (m,f) from lane-1; par==1: up neighbour's (m,e) from lane+1
        if (par == 0) {
            xm = __shfl_up_sync(0xffffffffu, mO[NC - 1], 1);
            xo = __shfl_up_sync(0xffffffffu, fO[NC - 1], 1);
            if (lane == 0) { xm = LB_NEG; xo = LB_NEG; }
        } else {
            xm = __shfl_down_sync(0xffffffffu, mE[0], 1);
            xo = __shfl_down_sync(0xffffffffu, eE[0], 1);
            if (lane == 31) { xm = LB_NEG; xo = LB_NEG; }
        }
        const int ring = (u & (RING - 1)) * NW;
        int nm[NC], ne[NC], nf[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const int gidx = lane * NC + k;
            const int cc = 2 * gidx + par;
            const int v = g.vmin + cc;
            const int ip = (u - v) >> 1, jp = (u + v) >> 1;
            int m_up, e_up, m_left, f_left, m_diag;
            if (par == 0) {
                m_up = mO[k]; e_up = eO[k]; m_diag = mE[k];
                if (k == 0) { m_left = xm; f_left = xo; } else { m_left = mO[k > 0 ? k - 1 : 0]; f_left = fO[k > 0 ? k - 1 : 0]; }
            } else {
                m_left = mE[k]; f_left = fE[k]; m_diag = mO[k];
                if (k == NC - 1) { m_up = xm; e_up = xo; } else { m_up = mE[k < NC - 1 ? k + 1 : k]; e_up = eE[k < NC - 1 ? k + 1 : k]; }
            }
            const bool inr = (ip >= 0) & (ip <= g.Rn) & (jp >= 0) & (jp <= g.Cn);
            const uint32_t ri = ws.rowinfo[inr ? ip : 0];
            const int jl = ri & 0xfff, jh = (ri >> 12) & 0xfff;
            const bool ok = inr & (jp >= jl) & (jp <= jh);
            const uint32_t a = ri >> 24, b = ws.colinfo[inr ? jp : 0];
            int sg;
            if ((a | b) < 4) sg = sig[a * 4 + b];
            else sg = (a == LB_CODE_N || b == LB_CODE_N) ? c.params.n_ext : (a == b ? c.params.match_ext : c.params.mismatch_ext);
            int e = addmax(e_up, gap, m_up + gap_open);
            int f = addmax(f_left, gap, m_left + gap_open);
            int arc = LB_NEG;
            if (gidx < g.nslots) { arc = ws.arcbuf[ring + gidx]; ws.arcbuf[ring + gidx] = LB_NEG; }
            int m = max3(addmax(m_diag, sg, e), f, arc);
            if (init.clamp0) m = max(m, 0);
            if (ip == 0) { m = (jp == 0) ? 0 : init.row_base + jp * init.row_step; e = LB_NEG; f = LB_NEG; }
            else if (jp == 0) { m = init.col_base + ip * init.col_step; e = LB_NEG; f = LB_NEG; }
            if (!ok) { m = LB_NEG; e = LB_NEG; f = LB_NEG; }
            else box[ip * g.strideD + cc] = m;
            nm[k] = m; ne[k] = e; nf[k] = f;
        }
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (par == 0) { mE[k] = nm[k]; eE[k] = ne[k]; fE[k] = nf[k]; }
            else { mO[k] = nm[k]; eO[k] = ne[k]; fO[k] = nf[k]; }
        }
        __syncwarp();
        // ---- fold the arc-match terms whose right ends lie on anti-diagonal u + LOOKAHEAD
        {
            const int ut = u + LOOKAHEAD;
            const int s = s_base + ut;
            if (ut <= g.umax && s <= pr.lenA + pr.lenB) {
                const int e0 = sptr[s], e1 = sptr[s + 1];
                const int tring = (ut & (RING - 1)) * NW;
                for (int e = e0 + lane; e < e1; e += 32) {
                    const DevEntry en = ent[e];
                    const int p = (en.x & 0xfff) - g.al, q = ((en.x >> 12) & 0xfff) - g.bl;
                    const int ar = (en.y & 0xfff) - g.al, br = ((en.y >> 12) & 0xfff) - g.bl;
                    if (p >= 0 && q >= 0 && ar <= g.Rn && br <= g.Cn) {
                        const int val = box[p * g.strideD + (q - p - g.vmin)] + dval[e];
                        const int ct = br - ar - g.vmin;
                        atomicMax(&ws.arcbuf[tring + (ct >> 1)], val);
                    }
                }
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ int box_get(const int *box, const BoxGeom &g, int ip, int jp) { return box[ip * g.strideD + (jp - ip - g.vmin)]; }

// Dispatch on the number of diagonal pairs per lane (NCMAX bounds the instantiated variants and thereby
// the register footprint of the kernel). Returns false if the band is wider than supported.
template <int NCMAX>
__device__ bool run_box(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const BoxInit &init, const WarpSmem &ws, int *box,
                        const int *sig) {
    const int nc = (g.nslots + 31) >> 5;
    if (nc <= 1) fill_box<1>(c, pr, g, init, ws, box, sig);
    else if (NCMAX >= 2 && nc <= 2) fill_box<(NCMAX >= 2 ? 2 : 1)>(c, pr, g, init, ws, box, sig);
    else if (NCMAX >= 3 && nc <= 3) fill_box<(NCMAX >= 3 ? 3 : 1)>(c, pr, g, init, ws, box, sig);
    else if (NCMAX >= 4 && nc <= 4) fill_box<(NCMAX >= 4 ? 4 : 1)>(c, pr, g, init, ws, box, sig);
    else if (NCMAX >= 8 && nc <= 8) fill_box<(NCMAX >= 8 ? 8 : 1)>(c, pr, g, init, ws, box, sig);
    else if (NCMAX >= 16 && nc <= 16) fill_box<(NCMAX >= 16 ? 16 : 1)>(c, pr, g, init, ws, box, sig);
    else return false;
    return true;
}

// shared memory layout of a one-warp CTA: [sigma4 16][rowinfo maxrows][colinfo maxcols (bytes, padded)][arcbuf RING*32*LB_MAX_NC ... sized by host][box ...]
__device__ __forceinline__ void carve(const DevCtx &c, int *smem, int *&sig, WarpSmem &ws, int &box_words) {
    sig = smem;
    ws.rowinfo = (uint32_t *)(smem + 16);
    ws.colinfo = (uint8_t *)(ws.rowinfo + c.max_rows);
    ws.arcbuf = (int *)(ws.colinfo + c.max_cols_padded);
    ws.box = ws.arcbuf + c.arcbuf_words;
    box_words = c.smem_words - (int)(ws.box - smem);
}

// ------------------------------------------------------------------------------------------------
// D-fill kernel: persistent one-warp CTAs pull the tasks of one scheduling level from an atomic cursor.
// Tasks of a level are mutually independent: a task reads D only for arc matches strictly inside its
// box, whose left ends have a larger al+bl (aligner.cc:675-728; levels = al+bl descending, two at a time).
template <int NCMAX>
__global__ void __launch_bounds__(32) dfill_kernel(DevCtx c, int task_begin, int task_end, int *cursor) {
    extern __shared__ int smem[];
    const int lane = threadIdx.x;
    int *sig; WarpSmem ws; int box_words;
    carve(c, smem, sig, ws, box_words);
    if (lane < 16) sig[lane] = c.params.sigma4[lane];
    __syncwarp();
    int *scratch = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const bool nolp = c.params.no_lonely_pairs != 0;
    BoxInit init;
    init.col_base = c.params.open; init.col_step = c.params.gap; init.row_base = c.params.open; init.row_step = c.params.gap; init.clamp0 = 0;

    for (;;) {
        int t = 0;
        if (lane == 0) t = task_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= task_end) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        BoxGeom g;
        setup_box(c, pr, task.al, task.bl, task.R, task.C, g, ws.rowinfo, ws.colinfo);
        const int need = (g.Rn + 1) * g.strideD;
        int *box = (need <= box_words) ? ws.box : scratch;
        if (need > c.scratch_words && need > box_words) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box<NCMAX>(c, pr, g, init, ws, box, sig)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        // ---- D entries of all arc matches with these left ends (aligner.cc:574-657)
        const DevArcMatch *am = c.am + pr.am_base;
        int *dval = c.dval + pr.am_base;
        const int sh = nolp ? 2 : 1;
        for (int k = task.run_start + lane; k < task.run_start + task.run_count; k += 32) {
            const DevArcMatch x = am[k];
            if (nolp && x.inner < 0) continue;
            const int ar = (x.ends_a >> 12) & 0xfff, br = (x.ends_b >> 12) & 0xfff;
            const int mv = box_get(box, g, ar - sh - g.al, br - sh - g.bl);
            int d;
            if (nolp) {
                const DevArcMatch in = am[x.inner];
                const int a = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + in.score;
                const int y = max(a, dval[in.spos]);
                d = (y < LB_NEG_LIMIT) ? LB_NEG : y + x.score;
            } else d = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + x.score;
            dval[x.spos] = d;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Top level (aligner.cc:736-880): one box over the whole band with al = bl = 0, then the score.
template <int NCMAX>
__global__ void __launch_bounds__(32) toplevel_kernel(DevCtx c, int pair_begin, int pair_end, int *cursor) {
    extern __shared__ int smem[];
    const int lane = threadIdx.x;
    int *sig; WarpSmem ws; int box_words;
    carve(c, smem, sig, ws, box_words);
    if (lane < 16) sig[lane] = c.params.sigma4[lane];
    __syncwarp();
    int *scratch = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const DevParams &P = c.params;
    BoxInit init;
    // init_state(E_NO_NO, 0, lenA+1, 0, lenB+1, !allow_left_2, false, !allow_left_1, false) resp. all-local
    const bool globalA = !(P.sequ_local || P.fe_left2), globalB = !(P.sequ_local || P.fe_left1);
    init.col_base = globalA ? P.open : 0; init.col_step = globalA ? P.gap : 0;
    init.row_base = globalB ? P.open : 0; init.row_step = globalB ? P.gap : 0;
    init.clamp0 = P.sequ_local;
    for (;;) {
        int t = 0;
        if (lane == 0) t = pair_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= pair_end) break;
        const DevPair pr = c.pairs[t];
        BoxGeom g;
        setup_box(c, pr, 0, 0, pr.lenA, pr.lenB, g, ws.rowinfo, ws.colinfo);
        const int need = (g.Rn + 1) * g.strideD;
        int *box = (need <= box_words) ? ws.box : scratch;
        if (need > c.scratch_words && need > box_words) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box<NCMAX>(c, pr, g, init, ws, box, sig)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        const int n = pr.lenA, m = pr.lenB;
        auto cell = [&](int i, int j) -> int {
            const uint32_t ri = ws.rowinfo[i];
            const int jl = ri & 0xfff, jh = (ri >> 12) & 0xfff;
            if (j < jl || j > jh) return LB_NEG;
            const int v = box_get(box, g, i, j);
            return v < LB_NEG_LIMIT ? LB_NEG : v;
        };
        // candidates are ranked by (score, earlier position in the reference's scan order)
        int best = LB_NEG, bi = 0, bj = 0;
        long long bkey = 0x7fffffffffffffffLL;  // smaller = earlier in scan order
        auto consider = [&](int v, int i, int j, long long key) {
            if (v > best || (v == best && v > LB_NEG && key < bkey)) { best = v; bi = i; bj = j; bkey = key; }
        };
        if (P.sequ_local) {
            // first strict maximum in row-major order, initial best 0 at (0,0) (aligner.cc:829-876)
            best = 0; bkey = -1;
            for (int i = 1; i <= n; i++) {
                const uint32_t ri = ws.rowinfo[i];
                const int jl = max(1, (int)(ri & 0xfff)), jh = (ri >> 12) & 0xfff;
                for (int j = jl + lane; j <= jh; j += 32) consider(cell(i, j), i, j, (long long)i * (m + 1) + j);
            }
        } else if (P.fe_right1 || P.fe_right2) {
            // aligner.cc:775-816: last column scanned by rows first, then the last row by columns; strict >
            if (P.fe_right2) {
                for (int i = 1 + lane; i <= n; i += 32) {
                    const int jh = (ws.rowinfo[i] >> 12) & 0xfff;
                    if (jh >= m) consider(cell(i, m), i, m, (long long)i);
                }
            }
            if (P.fe_right1) {
                const uint32_t ri = ws.rowinfo[n];
                const int jl = max(1, (int)(ri & 0xfff)), jh = (ri >> 12) & 0xfff;
                for (int j = jl + lane; j <= jh; j += 32) consider(cell(n, j), n, j, (long long)n + 1 + j);
            }
        } else {
            if (lane == 0) { best = cell(n, m); bi = n; bj = m; bkey = 0; }
        }
        // warp arg-max with the scan-order tie break
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oj = __shfl_xor_sync(0xffffffffu, bj, o);
            const long long ok = __shfl_xor_sync(0xffffffffu, bkey, o);
            if (ob > best || (ob == best && ok < bkey)) { best = ob; bi = oi; bj = oj; bkey = ok; }
        }
        if (lane == 0) {
            DevTopResult r;
            r.score = best; r.max_i = bi; r.max_j = bj; r.pad = 0;
            if (!P.sequ_local && (P.fe_right1 || P.fe_right2) && best <= LB_NEG_LIMIT) { r.max_i = 0; r.max_j = 0; }
            c.top[t] = r;
        }
        __syncwarp();
    }
}

// host-side launchers; ncmax selects the instantiation (1, 2, 4, 8 or 16 diagonal pairs per lane)
#define LB_DISPATCH(ncmax, CALL)                                  \
    do {                                                          \
        if (ncmax <= 1) { CALL(1); }                              \
        else if (ncmax <= 2) { CALL(2); }                         \
        else if (ncmax <= 4) { CALL(4); }                         \
        else if (ncmax <= 8) { CALL(8); }                         \
        else { CALL(16); }                                        \
    } while (0)

void launch_dfill(const DevCtx &c, int ncmax, int grid, int smem_bytes, int task_begin, int task_end, int *cursor, cudaStream_t st) {
#define CALL(N) dfill_kernel<N><<<grid, 32, smem_bytes, st>>>(c, task_begin, task_end, cursor)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
void launch_toplevel(const DevCtx &c, int ncmax, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st) {
#define CALL(N) toplevel_kernel<N><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
cudaError_t configure_kernels(int ncmax, int smem_bytes, int *dfill_ctas_per_sm) {
    cudaError_t e = cudaSuccess;
#define CALL(N)                                                                                                         \
    e = cudaFuncSetAttribute(dfill_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);                 \
    if (e == cudaSuccess) e = cudaFuncSetAttribute(toplevel_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(dfill_ctas_per_sm, dfill_kernel<N>, 32, smem_bytes)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
    return e;
}

}  // namespace lb200
