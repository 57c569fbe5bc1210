// LocARNA-P reverse, outside and probability passes on the device (included by kernels.cu after pf_inside.cuh).
//
// Reference (file:line relative to /root/reference/src/LocARNA, T = double):
//   init_Mrev / comp_*rev_entry / align_reverse      aligner_p.icc:451-627    suffix partition functions
//   leftmost / rightmost_covering_arc                aligner_p.icc:651-714
//   virtual_Mprime / comp_Mprime_entry               aligner_p.icc:784-886    hole recursion, cases 4 and 5
//   align_outside_arcmatch / fill_Dprime / align_Dprime / align_outside   aligner_p.icc:894-1139
//   compute_arcmatch_probabilities                   aligner_p.icc:1150-1195
//   compute_basematch_probabilities                  aligner_p.icc:1201-1399
//
// First device version, written for exactness rather than speed: one CTA of 128 threads per left-end pair (task), dense
// (lenA+1) x (lenB+1) FP64 matrices exactly as in the reference (per pair: prefix table M of the whole sequences, suffix tables
// Mrev / Erev / Frev of the whole sequences, base-match accumulator; per CTA: four scratch matrices), thread per cell on an
// anti-diagonal, the reference's border / zero-guard initialisation transliterated. Arc-match terms are enumerated through the
// two sorted arc-match tables instead of the reference's adjacency-list double loops (which visit all arc pairs and rely on D = 0
// for invalid ones): terms whose arcs END at a cell come from the S-order list of its anti-diagonal, terms whose arcs START at a
// cell from the L-order run of that cell. The E / F rolling vectors of the reference become dense matrices; "not computed in this
// sweep" reads as 0, which is what the rolling vectors hold there (the band is monotone).
// Levels: Dprime(a,b) of a hole needs Dprime of arc matches with smaller left ends, so the level groups of the inside pass are run
// in ascending order (aligner_p.icc:1062-1064).
#ifndef LB200_PF_OUTSIDE_CUH
#define LB200_PF_OUTSIDE_CUH

namespace lb200 {

struct PfoCtx {
    PfCtx pc;
    const int *cell_start;   // builder: absolute L-order offset of the run of every band cell (ranked al desc, bl desc)
    const int *cell_rev;     // builder: rank prefix per row
    const int *arc_left, *arc_right;   // arcs in BasePairs index order (left end descending, right end ascending)
    const int *lptr, *lcount;          // per left end: first arc (sequence relative) and number of arcs
    double *dpp;             // outside value Dprime(a,b) per S-order entry
    double *amp;             // arc-match probability per S-order entry
    double *mats;            // per pair: 5 dense matrices (Mtop, Mrev00, Erev00, Frev00, bm), each mat_doubles
    double *cta;             // per CTA: 4 dense scratch matrices
    long long mat_doubles;   // >= (max lenA + 1) * (max lenB + 1)
    int acc_cols;            // >= max lenB + 2: columns of the two shared-memory arc-term accumulators
    double am_threshold;     // sqrt(min_am_prob) (aligner_p.icc:1213-1216)
};

struct PV {   // per-pair view
    int n, m, W;
    const int *lo, *hi;
    const uint8_t *ca, *cb;
    const DevEntry *ent; const int *sptr; const DevArcMatch *am;
    const double *dpf; double *dpp; double *amp;
    const int *cell; const int *rev;
    long long am_base;
    const int *arA, *arB, *lpA, *lcA, *lpB, *lcB;
    double g, open, inv_scale, scale, temp;
    const double *esig;
    __device__ bool valid(int i, int j) const { return lo[i] <= j && j <= hi[i]; }
    __device__ bool valid_match(int i, int j) const { return i >= 1 && j >= 1 && valid(i, j) && valid(i - 1, j - 1); }
    __device__ double sig(int i, int j) const { return esig[ca[i] * LB_NCODES + cb[j]]; }
    // L-order run (pair-relative) of the arc matches with left ends (al, bl); empty if the cell has none
    __device__ void run(int al, int bl, int &r0, int &r1) const {
        r0 = r1 = 0;
        if (al < 1 || al > n) return;
        const int hi_eff = min(hi[al], m), lo_eff = max(lo[al], 1);
        if (bl < lo_eff || bl > hi_eff) return;
        const int r = rev[al] + (hi_eff - bl);
        r0 = cell[r] - (int)am_base; r1 = cell[r + 1] - (int)am_base;
    }
};

__device__ PV make_pv(const DevCtx &c, const PfoCtx &o, const DevPair &pr) {
    PV v;
    v.n = pr.lenA; v.m = pr.lenB; v.W = pr.lenB + 1;
    v.lo = c.band_lo + pr.band; v.hi = c.band_hi + pr.band;
    v.ca = c.codes + pr.codesA; v.cb = c.codes + pr.codesB;
    v.ent = c.ent + pr.am_base; v.sptr = c.sptr + pr.sptr; v.am = c.am + pr.am_base;
    v.dpf = o.pc.dpf + pr.am_base; v.dpp = o.dpp + pr.am_base; v.amp = o.amp + pr.am_base;
    v.cell = o.cell_start + pr.cell_base; v.rev = o.cell_rev + pr.band;
    v.am_base = pr.am_base;
    v.arA = o.arc_right + pr.arcsA; v.arB = o.arc_right + pr.arcsB;
    v.lpA = o.lptr + pr.lptrA; v.lcA = o.lcount + pr.lptrA; v.lpB = o.lptr + pr.lptrB; v.lcB = o.lcount + pr.lptrB;
    v.g = o.pc.g; v.open = o.pc.open; v.inv_scale = o.pc.inv_scale; v.scale = o.pc.pf_scale; v.temp = o.pc.temp;
    v.esig = o.pc.esig;
    return v;
}

#define PFO_AT(Mx, i, j) (Mx)[(size_t)(i) * v.W + (j)]

// Arc-match terms of a dense sweep: the entries of the S-order list of one anti-diagonal are handled one per thread and summed per
// target column into a shared-memory accumulator (the cells of an anti-diagonal have distinct columns); the cell pass of that
// anti-diagonal adds and clears it. Two accumulators alternate, so the entries of the next anti-diagonal are folded in the same
// barrier interval in which the cells of the current one are computed (their sources lie >= 8 anti-diagonals away and are final).
struct PfoAcc { double *a[2]; };

// align_inside_arcmatch on a dense matrix (aligner_p.icc:148-312); arc terms from the S-order list of the cell's anti-diagonal
__device__ void dense_inside(const PV &v, int al, int ar, int bl, int br, double *M, double *Ed, double *Fd, const PfoAcc &acc, const double *bpow) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // init_M :148-190 (border values from the table of sequential products, zero guards around the band)
    for (int i = al + tid; i < ar; i += nt) {
        if (i == al) PFO_AT(M, al, bl) = v.inv_scale;
        else if (v.lo[i] <= bl) PFO_AT(M, i, bl) = bpow[i - al];
        else PFO_AT(M, i, v.lo[i] - 1) = 0;
    }
    const int max_col = min(br - 1, v.hi[al]);
    for (int j = bl + 1 + tid; j <= max_col; j += nt) PFO_AT(M, al, j) = bpow[j - bl];
    if (tid == 0) {
        int j = max(max_col, bl) + 1;
        for (int i2 = al + 1; i2 < ar; i2++)
            for (; j < min(br, v.hi[i2] + 1); ++j) PFO_AT(M, i2 - 1, j) = 0;
    }
    auto cmin = [&](int i) { return max(bl + 1, v.lo[i]); };
    auto cmax = [&](int i) { return min(br - 1, v.hi[i]); };
    auto computed = [&](int i, int j) { return i > al && i < ar && j >= cmin(i) && j <= cmax(i); };
    // entries of anti-diagonal d -> accumulator d & 1
    auto fold = [&](int d) {
        if (d < al + bl + 2 || d > ar + br - 2) return;
        double *a = acc.a[d & 1];
        for (int t = v.sptr[d] + tid; t < v.sptr[d + 1]; t += nt) {
            const DevEntry en = v.ent[t];
            const int i = LB_ENT_LO(en.y), j = LB_ENT_HI(en.y);
            const int p = LB_ENT_LO(en.x), q = LB_ENT_HI(en.x);   // al'-1, bl'-1
            if (computed(i, j) && p >= al && q >= bl) atomicAdd(a + j, PFO_AT(M, p, q) * v.dpf[t] * v.scale);
        }
    };
    __syncthreads();
    fold(al + bl + 2);
    for (int d = al + bl + 2; d <= ar + br - 2; d++) {
        __syncthreads();
        double *a = acc.a[d & 1];
        for (int i = max(al + 1, d - (br - 1)) + tid; i <= min(ar - 1, d - (bl + 1)); i += nt) {
            const int j = d - i;
            if (j < cmin(i) || j > cmax(i)) continue;
            const double eu = computed(i - 1, j) ? PFO_AT(Ed, i - 1, j) : 0.0;
            const double e = eu * v.g + (PFO_AT(M, i - 1, j) - eu) * v.g * v.open;
            const double fl = computed(i, j - 1) ? PFO_AT(Fd, i, j - 1) : 0.0;
            const double f = fl * v.g + (PFO_AT(M, i, j - 1) - fl) * v.g * v.open;
            const double pf = PFO_AT(M, i - 1, j - 1) * v.sig(i, j) + e + f + a[j];
            a[j] = 0.0;
            PFO_AT(M, i, j) = pf; PFO_AT(Ed, i, j) = e; PFO_AT(Fd, i, j) = f;
        }
        fold(d + 1);
    }
    __syncthreads();
}

// align_reverse on a dense matrix (aligner_p.icc:451-627); arc terms from the L-order run of cell (i+1, j+1)
__device__ void dense_reverse(const PV &v, int al, int ar, int bl, int br, double *Mr, double *Ed, double *Fd, const double *bpow) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // the reference refills the region with -1 "for debugging" (:586-590); kept because such a cell can be read (see dense_outside)
    if (ar >= al && br >= bl) {
        const int cols = br - bl + 1;
        for (int k = tid; k < (ar - al + 1) * cols; k += nt) PFO_AT(Mr, al + k / cols, bl + k % cols) = -1.0;
    }
    __syncthreads();
    // init_Mrev :451-510 (border values from the table of sequential products; the band is monotone, so "until the band ends" is a
    // per-row / per-column test)
    for (int i = al - 1 + tid; i < ar; i += nt) {
        if (v.hi[i] >= br) PFO_AT(Mr, i, br) = bpow[ar - i];
        else PFO_AT(Mr, i, v.hi[i] + 1) = 0;
    }
    const int min_col = max(bl - 1, v.lo[ar]);
    for (int j = min_col + tid; j < br; j += nt) PFO_AT(Mr, ar, j) = bpow[br - j];
    if (tid == 0) {
        PFO_AT(Mr, ar, br) = v.inv_scale;
        int j = min(min_col, br);
        for (int i2 = ar; i2 >= al;) {
            i2--;
            for (; j > max(bl - 1, v.lo[i2]);) { --j; PFO_AT(Mr, i2 + 1, j) = 0; }
        }
    }
    __syncthreads();
    // rows al-1 .. ar-1, columns [cmin(i), cmax(i)] (:598-606)
    auto cmin = [&](int i) { return max(bl, v.lo[i] + 1) - 1; };
    auto cmax = [&](int i) { return min(br, v.hi[i] + 1) - 1; };
    auto computed = [&](int i, int j) { return i >= al - 1 && i <= ar - 1 && j >= cmin(i) && j <= cmax(i); };
    for (int d = ar + br - 2; d >= al + bl - 2; d--) {
        for (int i = max(al - 1, d - (br - 1)) + tid; i <= min(ar - 1, d - (bl - 1)); i += nt) {
            const int j = d - i;
            if (j < cmin(i) || j > cmax(i)) continue;
            const double eu = computed(i + 1, j) ? PFO_AT(Ed, i + 1, j) : 0.0;
            const double e = eu * v.g + (PFO_AT(Mr, i + 1, j) - eu) * v.g * v.open;
            const double fr = computed(i, j + 1) ? PFO_AT(Fd, i, j + 1) : 0.0;
            const double f = fr * v.g + (PFO_AT(Mr, i, j + 1) - fr) * v.g * v.open;
            double pf = PFO_AT(Mr, i + 1, j + 1) * v.sig(i + 1, j + 1) + e + f;
            int r0, r1;
            v.run(i + 1, j + 1, r0, r1);
            for (int k = r0; k < r1; k++) {
                const DevArcMatch x = v.am[k];
                const int xr = x.ends_a >> 12, yr = x.ends_b >> 12;
                if (xr <= ar && yr <= br) pf += v.dpf[x.spos] * PFO_AT(Mr, xr, yr) * v.scale;
            }
            PFO_AT(Mr, i, j) = pf; PFO_AT(Ed, i, j) = e; PFO_AT(Fd, i, j) = f;
        }
        __syncthreads();
    }
}

// align_outside_arcmatch (aligner_p.icc:894-998) for the hole with left ends (al, bl): Mprime over rows ar..max_ar, columns br..max_br
__device__ void dense_outside(const PV &v, int al, int ar, int max_ar, int bl, int br, int max_br, double m0, double *Mp, double *Ed, double *Fd,
                              const double *Mrl, const double *Mr00, const double *Er00, const double *Fr00, const PfoAcc &acc) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // The reference refills Mprime with -1 before every hole, "only for debugging" (:918) - but the value is read: the cell diagonally
    // behind a band corner of the region (e.g. Mprime(i+1, max_br) with (i+1, max_br) outside the band) is neither initialised nor
    // guarded, so -1 * exp_basematch enters Mprime(i, max_br - 1). Reproduced here (only the rectangle that can be read is refilled).
    {
        const int cols = max_br - br + 1;
        for (int k = tid; k < (max_ar - ar + 1) * cols; k += nt) PFO_AT(Mp, ar + k / cols, br + k % cols) = -1.0;
    }
    __syncthreads();
    // :920-965
    for (int i = ar + tid; i <= max_ar; i += nt) {
        if (i == max_ar || v.hi[i] >= max_br) { if (v.valid(i, max_br)) PFO_AT(Mp, i, max_br) = m0 * PFO_AT(Mr00, i, max_br) * v.scale; }
        else if (v.hi[i] + 1 <= max_br) PFO_AT(Mp, i, v.hi[i] + 1) = 0;
    }
    const int min_col = max(br, v.lo[max_ar]), max_col = min(max_br - 1, v.hi[max_ar]);
    for (int j = min_col + tid; j <= max_col; j += nt) PFO_AT(Mp, max_ar, j) = m0 * PFO_AT(Mr00, max_ar, j) * v.scale;
    if (tid == 0) {
        int j = max_col + 1 > min_col ? min_col : max_col + 1;
        for (int i2 = max_ar; i2 > ar;) {
            i2--;
            for (; j > max(bl, v.lo[i2]);) { --j; PFO_AT(Mp, i2 + 1, j) = 0; }
        }
    }
    __syncthreads();
    auto cmin = [&](int i) { return max(br, v.lo[i]); };
    auto cmax = [&](int i) { return min(max_br - 1, v.hi[i]); };
    auto computed = [&](int i, int j) { return i >= ar && i <= max_ar - 1 && j >= cmin(i) && j <= cmax(i); };
    const int e_lo = max(br, v.lo[max_ar]), e_hi = min(max_br - 1, v.hi[max_ar]);   // Eprime of the border row max_ar (:948-953)
    auto vmp = [&](int i, int j) { return (i >= max_ar || j >= max_br) ? m0 * PFO_AT(Mr00, i, j) * v.scale : PFO_AT(Mp, i, j); };   // virtual_Mprime :784-796
    // case 4 (:836-858): arc matches with right ends (i+1, j+1) and left ends before (al, bl), from the list of anti-diagonal d + 2
    auto fold = [&](int d) {
        if (d < ar + br || d > max_ar + max_br - 2) return;
        double *a = acc.a[d & 1];
        for (int t = v.sptr[d + 2] + tid; t < v.sptr[d + 3]; t += nt) {
            const DevEntry en = v.ent[t];
            const int i = LB_ENT_LO(en.y) - 1, j = LB_ENT_HI(en.y) - 1;
            const int p = LB_ENT_LO(en.x) + 1, q = LB_ENT_HI(en.x) + 1;   // al', bl'
            if (computed(i, j) && p < al && q < bl) atomicAdd(a + j, v.dpp[t] * PFO_AT(Mrl, p, q) * v.scale);
        }
    };
    fold(max_ar + max_br - 2);
    for (int d = max_ar + max_br - 2; d >= ar + br; d--) {
        __syncthreads();
        double *a = acc.a[d & 1];
        for (int i = max(ar, d - (max_br - 1)) + tid; i <= min(max_ar - 1, d - br); i += nt) {
            const int j = d - i;
            if (j < cmin(i) || j > cmax(i)) continue;
            double fr;
            if (computed(i, j + 1)) fr = PFO_AT(Fd, i, j + 1);
            else fr = v.valid(i, max_br) ? m0 * PFO_AT(Fr00, i, max_br) * v.scale : 0.0;   // Fprime at the row start (:968-972)
            const double f = fr * v.g + (PFO_AT(Mp, i, j + 1) - fr) * v.g * v.open;
            double eu;
            if (i + 1 == max_ar) eu = (j >= e_lo && j <= e_hi) ? m0 * PFO_AT(Er00, max_ar, j) * v.scale : 0.0;
            else eu = computed(i + 1, j) ? PFO_AT(Ed, i + 1, j) : 0.0;
            const double e = eu * v.g + (PFO_AT(Mp, i + 1, j) - eu) * v.g * v.open;
            double pf = PFO_AT(Mp, i + 1, j + 1) * v.sig(i + 1, j + 1) + e + f + a[j];
            a[j] = 0.0;
            {   // case 5 (:861-881): arc matches with left ends (i+1, j+1)
                int r0, r1;
                v.run(i + 1, j + 1, r0, r1);
                for (int k = r0; k < r1; k++) {
                    const DevArcMatch x = v.am[k];
                    pf += vmp(x.ends_a >> 12, x.ends_b >> 12) * v.dpf[x.spos] * v.scale;
                }
            }
            PFO_AT(Mp, i, j) = pf; PFO_AT(Ed, i, j) = e; PFO_AT(Fd, i, j) = f;
        }
        fold(d - 1);
    }
    __syncthreads();
}

__device__ PfoAcc make_acc(const PfoCtx &o, double *smem) {
    PfoAcc acc;
    acc.a[0] = smem; acc.a[1] = smem + o.acc_cols;
    for (int k = threadIdx.x; k < 2 * o.acc_cols; k += blockDim.x) smem[k] = 0.0;
    __syncthreads();
    return acc;
}

// per pair: prefix table of the whole sequences (aligner_p.icc:428-433) and suffix tables with the E / F copies (:1133)
__global__ void __launch_bounds__(128) pfo_prepare_kernel(DevCtx c, PfoCtx o, int n_pairs) {
    extern __shared__ double pfo_smem[];
    const PfoAcc acc = make_acc(o, pfo_smem);
    for (int pk = blockIdx.x; pk < n_pairs; pk += gridDim.x) {
        const DevPair pr = c.pairs[pk];
        const PV v = make_pv(c, o, pr);
        double *mats = o.mats + (size_t)pk * 5 * o.mat_doubles;
        double *cta = o.cta + (size_t)blockIdx.x * 4 * o.mat_doubles;
        dense_inside(v, 0, v.n + 1, 0, v.m + 1, mats, cta, cta + o.mat_doubles, acc, o.pc.bpow);
        dense_reverse(v, 1, v.n, 1, v.m, mats + o.mat_doubles, mats + 2 * o.mat_doubles, mats + 3 * o.mat_doubles, o.pc.bpow);
        for (size_t k = threadIdx.x; k < (size_t)(v.n + 1) * v.W; k += blockDim.x) mats[4 * o.mat_doubles + k] = 0.0;
        __syncthreads();
    }
}

// one level group of holes: align_outside_arcmatch + fill_Dprime (aligner_p.icc:1048-1108)
__global__ void __launch_bounds__(128) pfo_outside_kernel(DevCtx c, PfoCtx o, int q, int *cursor) {
    extern __shared__ double pfo_smem[];
    __shared__ int s_task;
    __shared__ int s_red[4];
    const int task_begin = c.qstart[q], task_end = c.qstart[q + 1];
    if ((int)blockIdx.x >= task_end - task_begin) return;
    const PfoAcc acc = make_acc(o, pfo_smem);
    double *cta = o.cta + (size_t)blockIdx.x * 4 * o.mat_doubles;
    double *Mrl = cta, *Mp = cta + o.mat_doubles, *Ed = cta + 2 * o.mat_doubles, *Fd = cta + 3 * o.mat_doubles;
    for (;;) {
        if (threadIdx.x == 0) { s_task = task_begin + atomicAdd(cursor + q, 1); s_red[0] = 1 << 20; s_red[1] = 1 << 20; s_red[2] = 0; s_red[3] = 0; }
        __syncthreads();
        const int t = s_task;
        if (t >= task_end) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        const PV v = make_pv(c, o, pr);
        const double *mats = o.mats + (size_t)task.pair * 5 * o.mat_doubles;
        const double *Mtop = mats, *Mr00 = mats + o.mat_doubles, *Er00 = mats + 2 * o.mat_doubles, *Fr00 = mats + 3 * o.mat_doubles;
        const int al = task.al, bl = task.bl;
        // minimal right ends of the arc matches with these left ends (arc_matches.cc:358-370)
        int min_ar = v.n + 1, min_br = v.m + 1;
        for (int k = task.run_start + threadIdx.x; k < task.run_start + task.run_count; k += blockDim.x) {
            const DevArcMatch x = v.am[k];
            min_ar = min(min_ar, (int)(x.ends_a >> 12)); min_br = min(min_br, (int)(x.ends_b >> 12));
        }
        atomicMin(&s_red[0], min_ar); atomicMin(&s_red[1], min_br);
        __syncthreads();
        min_ar = s_red[0]; min_br = s_red[1];
        __syncthreads();
        // covering arcs (aligner_p.icc:651-714): leftmost start (rightmost end) of an arc that begins before the hole and ends after
        // min_ar / min_br; per left end only its longest arc matters (arcs of a left end are stored right end ascending)
        if (threadIdx.x == 0) { s_red[0] = al; s_red[1] = bl; s_red[2] = min_ar; s_red[3] = min_br; }
        __syncthreads();
        int sA, sB, max_ar, max_br;
        {
            int sa = al, ma = min_ar, sb = bl, mb = min_br;
            for (int l = 1 + threadIdx.x; l < al; l += blockDim.x) {
                const int cnt = v.lcA[l];
                if (cnt) { const int r = v.arA[v.lpA[l] + cnt - 1]; if (r > min_ar) { sa = min(sa, l); ma = max(ma, r); } }
            }
            for (int l = 1 + threadIdx.x; l < bl; l += blockDim.x) {
                const int cnt = v.lcB[l];
                if (cnt) { const int r = v.arB[v.lpB[l] + cnt - 1]; if (r > min_br) { sb = min(sb, l); mb = max(mb, r); } }
            }
            atomicMin(&s_red[0], sa); atomicMin(&s_red[1], sb); atomicMax(&s_red[2], ma); atomicMax(&s_red[3], mb);
            __syncthreads();
            sA = s_red[0]; sB = s_red[1]; max_ar = s_red[2]; max_br = s_red[3];
        }
        const double m0 = PFO_AT(Mtop, al - 1, bl - 1);
        dense_reverse(v, sA + 1, al - 1, sB + 1, bl - 1, Mrl, Ed, Fd, o.pc.bpow);
        dense_outside(v, al, min_ar, max_ar, bl, min_br, max_br, m0, Mp, Ed, Fd, Mrl, Mr00, Er00, Fr00, acc);
        // fill_Dprime (:1011-1043)
        for (int k = task.run_start + threadIdx.x; k < task.run_start + task.run_count; k += blockDim.x) {
            const DevArcMatch x = v.am[k];
            const int xr = x.ends_a >> 12, yr = x.ends_b >> 12;
            const double vm = (xr >= max_ar || yr >= max_br) ? m0 * PFO_AT(Mr00, xr, yr) * v.scale : PFO_AT(Mp, xr, yr);
            v.dpp[x.spos] = vm * exp((double)x.score / v.temp);
        }
        __syncthreads();
    }
}

// compute_arcmatch_probabilities (aligner_p.icc:1150-1195)
__global__ void pfo_amprob_kernel(DevCtx c, PfoCtx o, int n_pairs) {
    for (int pk = blockIdx.x; pk < n_pairs; pk += gridDim.x) {
        const DevPair pr = c.pairs[pk];
        const double Z = o.pc.ztop[pk];
        const DevArcMatch *am = c.am + pr.am_base;
        for (int k = threadIdx.x; k < pr.K; k += blockDim.x) {
            const DevArcMatch x = am[k];
            o.amp[pr.am_base + x.spos] = (o.pc.dpf[pr.am_base + x.spos] / Z) * o.dpp[pr.am_base + x.spos] * o.pc.pf_scale / exp((double)x.score / o.pc.temp);
        }
    }
}

// compute_basematch_probabilities, enclosed case (aligner_p.icc:1220-1330): one CTA per left-end pair
__global__ void __launch_bounds__(128) pfo_bm_enclosed_kernel(DevCtx c, PfoCtx o, int n_tasks, int *cursor) {
    extern __shared__ double pfo_smem[];
    __shared__ int s_task;
    __shared__ int s_red[2];
    const PfoAcc acc = make_acc(o, pfo_smem);
    double *cta = o.cta + (size_t)blockIdx.x * 4 * o.mat_doubles;
    double *M = cta, *Mr = cta + o.mat_doubles, *Ed = cta + 2 * o.mat_doubles, *Fd = cta + 3 * o.mat_doubles;
    for (;;) {
        if (threadIdx.x == 0) { s_task = atomicAdd(cursor, 1); s_red[0] = 0; s_red[1] = 0; }
        __syncthreads();
        const int t = s_task;
        if (t >= n_tasks) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        const PV v = make_pv(c, o, pr);
        double *bm = o.mats + (size_t)task.pair * 5 * o.mat_doubles + 4 * o.mat_doubles;
        const int al = task.al, bl = task.bl;
        int max_ar = al, max_br = bl;
        for (int k = task.run_start + threadIdx.x; k < task.run_start + task.run_count; k += blockDim.x) {
            const DevArcMatch x = v.am[k];
            if (v.amp[x.spos] > o.am_threshold) { max_ar = max(max_ar, (int)(x.ends_a >> 12)); max_br = max(max_br, (int)(x.ends_b >> 12)); }
        }
        atomicMax(&s_red[0], max_ar); atomicMax(&s_red[1], max_br);
        __syncthreads();
        max_ar = s_red[0]; max_br = s_red[1];
        __syncthreads();
        if (max_ar == al) continue;   // no arc match of this cell above the threshold
        dense_inside(v, al, max_ar, bl, max_br, M, Ed, Fd, acc, o.pc.bpow);
        for (int k = task.run_start; k < task.run_start + task.run_count; k++) {
            const DevArcMatch x = v.am[k];
            if (!(v.amp[x.spos] > o.am_threshold)) continue;
            const int ar = x.ends_a >> 12, br = x.ends_b >> 12;
            dense_reverse(v, al + 1, ar - 1, bl + 1, br - 1, Mr, Ed, Fd, o.pc.bpow);
            const double outside_pf = v.dpp[x.spos];
            const int rows = ar - al - 1, cols = br - bl - 1;
            for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
                const int i = al + 1 + idx / cols, j = bl + 1 + idx % cols;
                if (j < max(bl + 1, v.lo[i]) || j > min(br - 1, v.hi[i]) || !v.valid_match(i, j)) continue;
                atomicAdd(&PFO_AT(bm, i, j), PFO_AT(M, i - 1, j - 1) * v.sig(i, j) * PFO_AT(Mr, i, j) * v.scale * outside_pf * v.scale);
            }
            __syncthreads();
        }
    }
}

// unenclosed case and normalisation (aligner_p.icc:1332-1378)
__global__ void pfo_bm_final_kernel(DevCtx c, PfoCtx o, int n_pairs) {
    for (int pk = blockIdx.x; pk < n_pairs; pk += gridDim.x) {
        const DevPair pr = c.pairs[pk];
        const PV v = make_pv(c, o, pr);
        double *mats = o.mats + (size_t)pk * 5 * o.mat_doubles;
        const double *Mtop = mats, *Mr00 = mats + o.mat_doubles;
        double *bm = mats + 4 * o.mat_doubles;
        const double Z = o.pc.ztop[pk];
        for (int idx = threadIdx.x; idx < v.n * v.m; idx += blockDim.x) {
            const int i = 1 + idx / v.m, j = 1 + idx % v.m;
            if (j < max(1, v.lo[i]) || j > min(v.m, v.hi[i]) || !v.valid_match(i, j)) continue;
            PFO_AT(bm, i, j) = (PFO_AT(bm, i, j) + PFO_AT(Mtop, i - 1, j - 1) * v.sig(i, j) * PFO_AT(Mr00, i, j) * v.scale) / Z;
        }
    }
}

#undef PFO_AT

void launch_pfo_prepare(const DevCtx &c, const PfoCtx &o, int n_pairs, int grid, cudaStream_t st) { pfo_prepare_kernel<<<grid, 128, 2 * o.acc_cols * sizeof(double), st>>>(c, o, n_pairs); }
void launch_pfo_outside(const DevCtx &c, const PfoCtx &o, int q, int grid, int *cursor, cudaStream_t st) { pfo_outside_kernel<<<grid, 128, 2 * o.acc_cols * sizeof(double), st>>>(c, o, q, cursor); }
void launch_pfo_amprob(const DevCtx &c, const PfoCtx &o, int n_pairs, cudaStream_t st) { pfo_amprob_kernel<<<min(n_pairs, 1184), 256, 0, st>>>(c, o, n_pairs); }
void launch_pfo_bm(const DevCtx &c, const PfoCtx &o, int n_tasks, int n_pairs, int grid, int *cursor, cudaStream_t st) {
    pfo_bm_enclosed_kernel<<<grid, 128, 2 * o.acc_cols * sizeof(double), st>>>(c, o, n_tasks, cursor);
    pfo_bm_final_kernel<<<min(n_pairs, 1184), 256, 0, st>>>(c, o, n_pairs);
}

}  // namespace lb200
#endif
