// LocARNA-P inside pass on the device (included by kernels.cu; shares setup_box / BoxGeom / WarpSmem with the integer sweep).
//
// Reference (file:line relative to /root/reference/src/LocARNA, T = double as in locarna_p.cc:285-294):
//   PFScoring<T>                scoring.hh:744-931    Boltzmann weights exp(score / temperature_alipf)
//   init_M / init_E             aligner_p.icc:148-202 borders of one box, zero guards around the band
//   comp_E/F/M_entry            aligner_p.icc:209-279 sum-product cell recurrence
//   align_inside_arcmatch       aligner_p.icc:287-312 fill of one box
//   fill_D / align_D            aligner_p.icc:325-407 D(a,b) = M(ar-1, br-1) * exp_arcmatch(am)
//   align_inside                aligner_p.icc:413-438 top level box, Z = M(lenA, lenB)
//
// Same schedule as the integer D fill: one warp per left-end pair with arc matches, level groups (al+bl)>>1 descending, anti-diagonal
// sweep with lanes owning pairs of adjacent diagonals, the box kept in an L2-resident FP64 scratch. The reference visits every band
// cell (al, bl) (aligner_p.icc:380-403), but boxes of cells without arc matches define no D entry and are skipped here.
// Cells outside the band are 0 (the reference writes explicit zero guards, :166-189); E and F of border cells are 0 (:192-202, :297).
// The arc-match terms of one target cell are summed with shared-memory atomics, so their order (and the last bits of the result) can
// differ from the reference's loop order (:264-276); the parity bar for this path is 1e-6 relative.
#ifndef LB200_PF_INSIDE_CUH
#define LB200_PF_INSIDE_CUH

namespace lb200 {

struct PfCtx {
    const double *esig;      // 64: exp(sigma8 / temp)                                     (scoring.hh:903-915)
    const double *bpow;      // bpow[k] = (exp_open / pf_scale) * g^k by sequential products (aligner_p.icc:156-176), k >= 1
    double g;                // exp(gap / temp)             (position independent for single sequences, scoring.hh:917-931)
    double open;             // exp(indel_opening / temp)
    double inv_scale;        // 1 / pf_scale                (aligner_p.icc:151)
    double pf_scale;
    double temp;             // temperature_alipf
    double *dpf;             // inside value D(a,b) per S-order entry
    double *ztop;            // per pair: partition function
    double *scratch;         // per-CTA FP64 box
    long long scratch_dwords;
    int acc_doubles;         // 32 * NC
};

struct PfSmem { const double *esig; double *acc; };

__device__ __forceinline__ double pbox_get(const double *box, const BoxGeom &g, int ip, int jp) {
    return box[(ip + jp) * g.nslots + ((jp - ip - g.vmin) >> 1)];
}

template <int NC, int PAR>
__device__ __forceinline__ void dp_step_p(const BoxGeom &g, const WarpSmem &ws, const PfSmem &ps, const PfCtx &pc, double *box, int u, int lane,
                                          double (&mE)[NC], double (&eE)[NC], double (&fE)[NC], double (&mO)[NC], double (&eO)[NC], double (&fO)[NC]) {
    double xm, xo;
    if (PAR == 0) {
        xm = __shfl_up_sync(0xffffffffu, mO[NC - 1], 1);
        xo = __shfl_up_sync(0xffffffffu, fO[NC - 1], 1);
        if (lane == 0) { xm = 0.0; xo = 0.0; }
    } else {
        xm = __shfl_down_sync(0xffffffffu, mE[0], 1);
        xo = __shfl_down_sync(0xffffffffu, eE[0], 1);
        if (lane == 31) { xm = 0.0; xo = 0.0; }
    }
    const int U2 = (u - g.vmin - PAR) >> 1, J2 = (u + g.vmin + PAR) >> 1;
    double *boxrow = box + u * g.nslots;
    const double gg = pc.g, op = pc.open;
    double nm[NC], ne[NC], nf[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const int gidx = lane * NC + k;
        double m_up, e_up, m_left, f_left, m_diag;
        if (PAR == 0) {
            m_up = mO[k]; e_up = eO[k]; m_diag = mE[k];
            if (k == 0) { m_left = xm; f_left = xo; } else { m_left = mO[k > 0 ? k - 1 : 0]; f_left = fO[k > 0 ? k - 1 : 0]; }
        } else {
            m_left = mE[k]; f_left = fE[k]; m_diag = mO[k];
            if (k == NC - 1) { m_up = xm; e_up = xo; } else { m_up = mE[k < NC - 1 ? k + 1 : k]; e_up = eE[k < NC - 1 ? k + 1 : k]; }
        }
        const int ip = U2 - gidx, jp = J2 + gidx;
        const uint32_t ridx = min((uint32_t)(ip + 1), (uint32_t)(g.Rn + 2));
        const uint32_t cidx = min((uint32_t)jp, (uint32_t)(g.Cn + 1));
        const uint32_t rr = ws.rowrange[ridx];
        const double sg = ps.esig[ws.rowcode[ridx] + ws.colcode[cidx]];
        const bool ok = (uint32_t)(2 * gidx + PAR - (int)(rr & 0xffff)) <= (rr >> 16);
        // comp_E_entry / comp_F_entry / comp_M_entry (aligner_p.icc:209-279), same expression order
        double e = e_up * gg + (m_up - e_up) * gg * op;
        double f = f_left * gg + (m_left - f_left) * gg * op;
        double m = m_diag * sg + e + f;
        m += ps.acc[gidx];
        ps.acc[gidx] = 0.0;
        if (!ok) { m = 0.0; e = 0.0; f = 0.0; }
        else {
            if (ip == 0) { m = (jp == 0) ? pc.inv_scale : pc.bpow[jp]; e = 0.0; f = 0.0; }      // init_M row al (:170-177)
            else if (jp == 0) { m = pc.bpow[ip]; e = 0.0; f = 0.0; }                          // init_M column bl (:153-163)
            boxrow[gidx] = m;
        }
        nm[k] = m; ne[k] = e; nf[k] = f;
    }
#pragma unroll
    for (int k = 0; k < NC; k++) {
        if (PAR == 0) { mE[k] = nm[k]; eE[k] = ne[k]; fE[k] = nf[k]; }
        else { mO[k] = nm[k]; eO[k] = ne[k]; fO[k] = nf[k]; }
    }
}

template <int NC>
__device__ void fill_box_p(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const WarpSmem &ws, const PfSmem &ps, const PfCtx &pc, double *box) {
    const int lane = threadIdx.x & 31;
    for (int k = lane; k < 32 * NC; k += 32) ps.acc[k] = 0.0;
    const int s0 = g.al + g.bl;
    const int *q = c.sptr + pr.sptr + s0;
    const int q_cap = pr.lenA + pr.lenB + 2 - s0;
    const DevEntry *ent = c.ent + pr.am_base;
    const double *dpf = pc.dpf + pr.am_base;
    double mE[NC], eE[NC], fE[NC], mO[NC], eO[NC], fO[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) { mE[k] = eE[k] = fE[k] = mO[k] = eO[k] = fO[k] = 0.0; }
    const int par0 = (0 - g.vmin) & 1;
    {
        const int c0 = -g.vmin;   // the origin M(al, bl) = 1 / pf_scale (aligner_p.icc:151)
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (2 * (lane * NC + k) + par0 == c0) { if (par0 == 0) mE[k] = pc.inv_scale; else mO[k] = pc.inv_scale; box[c0 >> 1] = pc.inv_scale; }
        }
    }
    __syncwarp();
    for (int u = 1; u <= g.umax; u++) {
        // arc-match terms of anti-diagonal u (sources lie >= 8 anti-diagonals back): prefix of the list whose sources lie in the box
        if (u >= 8) {
            const int e0 = __ldg(q + min(u, q_cap)), e1 = __ldg(q + min(u + 1, q_cap));
            for (int base = e0; base < e1; base += 32) {
                const int e = base + lane;
                bool in_prefix = false;
                if (e < e1) {
                    const DevEntry en = ent[e];
                    in_prefix = en.s >= s0;
                    const int p = LB_ENT_LO(en.x) - g.al, qq = LB_ENT_HI(en.x) - g.bl;
                    const int ar = LB_ENT_LO(en.y) - g.al, br = LB_ENT_HI(en.y) - g.bl;
                    if ((p | qq | (g.Rn - ar) | (g.Cn - br)) >= 0) {
                        const double term = __ldcg(box + (p + qq) * g.nslots + ((qq - p - g.vmin) >> 1)) * dpf[e] * pc.pf_scale;   // :272-273
                        atomicAdd(&ps.acc[(br - ar - g.vmin) >> 1], term);
                    }
                }
                if ((__ballot_sync(0xffffffffu, in_prefix) >> 31) == 0) break;
            }
        }
        __syncwarp();
        if (((u + par0) & 1) == 0) dp_step_p<NC, 0>(g, ws, ps, pc, box, u, lane, mE, eE, fE, mO, eO, fO);
        else dp_step_p<NC, 1>(g, ws, ps, pc, box, u, lane, mE, eE, fE, mO, eO, fO);
        __syncwarp();
    }
}

template <int NCMAX>
__device__ bool run_box_p(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const WarpSmem &ws, const PfSmem &ps, const PfCtx &pc, double *box) {
    const int nc = (g.nslots + 31) >> 5;
    if (nc <= 1) fill_box_p<1>(c, pr, g, ws, ps, pc, box);
    else if (NCMAX >= 2 && nc <= 2) fill_box_p<(NCMAX >= 2 ? 2 : 1)>(c, pr, g, ws, ps, pc, box);
    else if (NCMAX >= 4 && nc <= 4) fill_box_p<(NCMAX >= 4 ? 4 : 1)>(c, pr, g, ws, ps, pc, box);
    else if (NCMAX >= 8 && nc <= 8) fill_box_p<(NCMAX >= 8 ? 8 : 1)>(c, pr, g, ws, ps, pc, box);
    else return false;
    return true;
}

// shared memory of a LocARNA-P warp: [esig 64 doubles][rowrange max_rows+2][rowcode][colcode][pad to 8][acc 32*NC doubles]
__device__ __forceinline__ void carve_p(const DevCtx &c, const PfCtx &pc, int *smem, WarpSmem &ws, PfSmem &ps) {
    double *es = (double *)smem;
    ws.sig = nullptr;
    ws.rowrange = (uint32_t *)(es + 64);
    ws.rowcode = (uint8_t *)(ws.rowrange + c.max_rows + 2);
    ws.colcode = ws.rowcode + c.rowcode_bytes;
    ws.roww = nullptr; ws.colc = nullptr; ws.arcbuf = nullptr;
    const size_t off = (512 + (size_t)(c.max_rows + 2) * 4 + c.rowcode_bytes + c.colcode_bytes + 7) & ~(size_t)7;
    ps.acc = (double *)((char *)smem + off);
    ps.esig = es;
    const int lane = threadIdx.x & 31;
    es[lane] = pc.esig[lane]; es[lane + 32] = pc.esig[lane + 32];
    __syncwarp();
}

template <int NCMAX>
__global__ void __launch_bounds__(32) pfill_kernel(DevCtx c, PfCtx pc, int q) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    const int task_begin = c.qstart[q], task_end = c.qstart[q + 1];
    if ((int)blockIdx.x >= task_end - task_begin) return;
    int *cursor = c.cursor + q;
    WarpSmem ws; PfSmem ps;
    carve_p(c, pc, smem, ws, ps);
    double *box = pc.scratch + (size_t)blockIdx.x * pc.scratch_dwords;
    for (;;) {
        int t = 0;
        if (lane == 0) t = task_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= task_end) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        BoxGeom g;
        setup_box(c, pr, task.al, task.bl, task.R, task.C, g, ws);
        if ((long long)(g.umax + 1) * g.nslots > pc.scratch_dwords) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box_p<NCMAX>(c, pr, g, ws, ps, pc, box)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        // fill_D (aligner_p.icc:325-353): D(am) = M(ar-1, br-1) * exp_arcmatch(am)
        const DevArcMatch *am = c.am + pr.am_base;
        double *dpf = pc.dpf + pr.am_base;
        for (int k = task.run_start + lane; k < task.run_start + task.run_count; k += 32) {
            const DevArcMatch x = am[k];
            const int ar = (x.ends_a >> 12) & 0xfff, br = (x.ends_b >> 12) & 0xfff;
            dpf[x.spos] = pbox_get(box, g, ar - 1 - g.al, br - 1 - g.bl) * exp((double)x.score / pc.temp);   // scoring.hh:795-798, :853-856
        }
        __syncwarp();
    }
}

template <int NCMAX>
__global__ void __launch_bounds__(32) ptop_kernel(DevCtx c, PfCtx pc, int pair_begin, int pair_end, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    WarpSmem ws; PfSmem ps;
    carve_p(c, pc, smem, ws, ps);
    double *box = pc.scratch + (size_t)blockIdx.x * pc.scratch_dwords;
    for (;;) {
        int t = 0;
        if (lane == 0) t = pair_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= pair_end) break;
        const DevPair pr = c.pairs[t];
        BoxGeom g;
        setup_box(c, pr, 0, 0, pr.lenA, pr.lenB, g, ws);
        if ((long long)(g.umax + 1) * g.nslots > pc.scratch_dwords) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box_p<NCMAX>(c, pr, g, ws, ps, pc, box)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        if (lane == 0) {
            const int n = pr.lenA, m = pr.lenB;
            const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
            // partFunc = M(lenA, lenB) (aligner_p.icc:433); a cell outside the band holds 0
            pc.ztop[t] = (m >= ((n == 0) ? 0 : lo[n]) && m <= hi[n]) ? pbox_get(box, g, n, m) : 0.0;
        }
        __syncwarp();
    }
}

void launch_pfill(const DevCtx &c, const PfCtx &pc, int ncmax, int grid, int smem_bytes, int q, cudaStream_t st) {
    if (ncmax <= 1) pfill_kernel<1><<<grid, 32, smem_bytes, st>>>(c, pc, q);
    else if (ncmax <= 2) pfill_kernel<2><<<grid, 32, smem_bytes, st>>>(c, pc, q);
    else if (ncmax <= 4) pfill_kernel<4><<<grid, 32, smem_bytes, st>>>(c, pc, q);
    else pfill_kernel<8><<<grid, 32, smem_bytes, st>>>(c, pc, q);
}
void launch_ptop(const DevCtx &c, const PfCtx &pc, int ncmax, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st) {
    if (ncmax <= 1) ptop_kernel<1><<<grid, 32, smem_bytes, st>>>(c, pc, pair_begin, pair_end, cursor);
    else if (ncmax <= 2) ptop_kernel<2><<<grid, 32, smem_bytes, st>>>(c, pc, pair_begin, pair_end, cursor);
    else if (ncmax <= 4) ptop_kernel<4><<<grid, 32, smem_bytes, st>>>(c, pc, pair_begin, pair_end, cursor);
    else ptop_kernel<8><<<grid, 32, smem_bytes, st>>>(c, pc, pair_begin, pair_end, cursor);
}
cudaError_t configure_pf(int ncmax, int smem_bytes, int *ctas_per_sm) {
    cudaError_t e = cudaSuccess;
#define LB_PF_SET(N)                                                                                                            \
    do {                                                                                                                        \
        const int cap = lb200_sticky_smem(300 + N, smem_bytes);                                                                 \
        e = cudaFuncSetAttribute(pfill_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);                            \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ptop_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);       \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, pfill_kernel<N>, 32, smem_bytes);   \
    } while (0)
    if (ncmax <= 1) LB_PF_SET(1); else if (ncmax <= 2) LB_PF_SET(2); else if (ncmax <= 4) LB_PF_SET(4); else LB_PF_SET(8);
#undef LB_PF_SET
    return e;
}

}  // namespace lb200
#endif
