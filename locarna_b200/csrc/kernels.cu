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
// instructions (VIADDMNMX / VIMNMX3). The box itself is written once per cell into an L2-resident
// per-warp scratch (arc-match terms read M(al'-1, bl'-1) at arbitrary earlier cells); shared memory
// holds only the per-row band/sequence info and the arc-term accumulators, which keeps 32 warps per SM
// resident - the sweep is a dependent chain, so throughput comes from warps in flight.
// Arc-match terms are not pulled per cell: the valid arc matches of a pair are kept sorted by the
// anti-diagonal of their right ends (S-order); the warp streams them in chunks of 32 with coalesced,
// prefetched loads and folds M(src)+D into a ring of per-diagonal accumulators with shared-memory atomicMax.
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
};

struct BoxGeom {
    int al, bl, Rn, Cn;      // origin and local extent (rows 0..Rn, cols 0..Cn; row/col 0 are borders)
    int vmin, umax;
    int nslots;              // number of diagonal pairs = row stride of the box, which is stored anti-diagonal major:
                             // M(i',j') lives at box[(i'+j') * nslots + ((j'-i'-vmin) >> 1)], so the cells of one step
                             // are contiguous (coalesced stores)
};

constexpr int RING = 8;      // arc-term accumulators for 8 consecutive anti-diagonals

// per-warp shared memory
struct WarpSmem {
    const int *sig;      // 8x8 base match scores by symbol code
    uint32_t *rowrange;  // [ip+1], ip = -1..Rn+1: first valid diagonal (c index) | number of further valid diagonals << 16
    uint8_t *rowcode;    // [ip+1]: 8 * symbol code of A[al+ip]
    uint8_t *colcode;    // [jp],  jp = 0..Cn+1: symbol code of B[bl+jp]
    int *arcbuf;         // RING * 32*NC accumulators (16-byte aligned)
    uint32_t *roww;      // padded row words of the single-state sweep (setup_box2); shares the region with rowrange/rowcode/colcode
    uint8_t *colc;       // padded column codes (x4)
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

constexpr uint32_t ROW_INVALID = 0x0000ffffu;  // first valid diagonal 65535, width 0: never matches

// Stage the per-row band/sequence info of a box into shared memory and derive its geometry.
// Row ip covers local columns [jl, jh]: row 0 and column 0 are the initialised borders
// (aligner.cc:307-357): column 0 belongs to row ip as long as min_col(al+ip) <= bl.
__device__ void setup_box(const DevCtx &c, const DevPair &pr, int al, int bl, int R, int C, BoxGeom &g, const WarpSmem &ws) {
    const int lane = threadIdx.x & 31;
    g.al = al; g.bl = bl; g.Rn = R - al; g.Cn = C - bl;
    int vmin = 1 << 20, vmax = -(1 << 20), umax = 0;
    const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
    const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
    for (int ip = lane; ip <= g.Rn; ip += 32) {
        const int i = al + ip;
        const int l = lo[i], h = hi[i];
        int jl = (ip == 0 || l <= bl) ? 0 : (l - bl);
        int jh = min(g.Cn, h - bl);
        if (jh < jl) { jl = 0xffff; jh = 0xffff; }
        else { vmin = min(vmin, jl - ip); vmax = max(vmax, jh - ip); umax = max(umax, ip + jh); }
        ws.rowrange[ip + 1] = (uint32_t)jl | ((uint32_t)jh << 16);
        ws.rowcode[ip + 1] = (uint8_t)(((i >= 1) ? ca[i] : 0) * 8);
    }
    for (int jp = lane; jp <= g.Cn; jp += 32) ws.colcode[jp] = (bl + jp >= 1) ? cb[bl + jp] : 0;
    if (lane == 0) {
        ws.rowrange[0] = ROW_INVALID; ws.rowrange[g.Rn + 2] = ROW_INVALID;
        ws.rowcode[0] = 0; ws.rowcode[g.Rn + 2] = 0; ws.colcode[g.Cn + 1] = 0;
    }
    g.vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    g.umax = warp_max(umax);
    const int wd = vmax - g.vmin + 1;
    g.nslots = (wd + 1) >> 1;
    __syncwarp();
    // convert column ranges to diagonal ranges relative to vmin
    for (int ip = lane; ip <= g.Rn; ip += 32) {
        const uint32_t r = ws.rowrange[ip + 1];
        const int jl = r & 0xffff, jh = r >> 16;
        ws.rowrange[ip + 1] = (jl == 0xffff) ? ROW_INVALID : ((uint32_t)(jl - ip - g.vmin) | ((uint32_t)(jh - jl) << 16));
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Single-state sweep (global / sequence-local / free end gaps / noLP boxes).
//
// Row / column tables in shared memory, padded so that the cell loop needs no index clamps:
//   roww[row_pad + 1 + ip]   one word per row ip = -1..Rn+1 (everything else = ROWW_INVALID):
//        bits  0..10  first valid diagonal index c (c = j' - i' - vmin), bit 11 guard
//        bits 12..22  2047 - last valid diagonal index, bit 23 guard
//        bits 24..31  32 * symbol code of A[al+ip]  (byte offset of the sigma row)
//   colc[col_pad + jp]       4 * symbol code of B[bl+jp]
// Band test of a cell on diagonal c in row word w:  z = K(c) - w  with  K(c) = (c | 0x800) | ((2047 - c) | 0x800) << 12;
// the guard bits of z survive iff first <= c and c <= last (no borrow crosses a field), i.e. one IADD3 + one LOP3.
constexpr uint32_t ROWW_INVALID = 0x007ff7ffu;   // first = 2047, last = 0
constexpr uint32_t ROWW_GUARDS = 0x00800800u;

template <int NC> struct RingStride { static constexpr int v = NC <= 1 ? 32 : NC <= 2 ? 64 : NC <= 4 ? 128 : NC <= 8 ? 256 : 512; };

__device__ void setup_box2(const DevCtx &c, const DevPair &pr, int al, int bl, int R, int C, BoxGeom &g, const WarpSmem &ws) {
    const int lane = threadIdx.x & 31;
    g.al = al; g.bl = bl; g.Rn = R - al; g.Cn = C - bl;
    for (int k = lane; k < c.row_words; k += 32) ws.roww[k] = ROWW_INVALID;
    for (int k = lane; k < (c.col_bytes >> 2); k += 32) ((uint32_t *)ws.colc)[k] = 0;
    __syncwarp();
    int vmin = 1 << 20, vmax = -(1 << 20), umax = 0;
    const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
    const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
    uint32_t *rw = ws.roww + c.row_pad + 1;
    for (int ip = lane; ip <= g.Rn; ip += 32) {
        const int i = al + ip;
        const int l = lo[i], h = hi[i];
        int jl = (ip == 0 || l <= bl) ? 0 : (l - bl);
        int jh = min(g.Cn, h - bl);
        if (jh < jl) { jl = 0xffff; jh = 0xffff; }
        else { vmin = min(vmin, jl - ip); vmax = max(vmax, jh - ip); umax = max(umax, ip + jh); }
        rw[ip] = (uint32_t)jl | ((uint32_t)jh << 16);
    }
    for (int jp = lane; jp <= g.Cn; jp += 32) ws.colc[c.col_pad + jp] = (uint8_t)(((bl + jp >= 1) ? cb[bl + jp] : 0) * 4);
    g.vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    g.umax = warp_max(umax);
    const int wd = vmax - g.vmin + 1;
    g.nslots = (((wd + 1) >> 1) + 3) & ~3;   // row stride of the box: diagonal pairs, rounded up for vector stores
    __syncwarp();
    for (int ip = lane; ip <= g.Rn; ip += 32) {
        const uint32_t r = rw[ip];
        const int jl = r & 0xffff, jh = r >> 16;
        const int i = al + ip;
        const uint32_t code = (uint32_t)((i >= 1) ? ca[i] : 0) << 29;   // 32 * code in bits 24..31
        rw[ip] = (jl == 0xffff) ? ROWW_INVALID : ((uint32_t)(jl - ip - g.vmin) | ((uint32_t)(2047 - (jh - ip - g.vmin)) << 12) | code);
    }
    __syncwarp();
}

// Register state of the sweep: per diagonal pair the cell last computed on the even / odd diagonal, the current row word and
// column code, and the per-lane table pointers.
template <int NC>
struct Sweep {
    int mE[NC], eE[NC], fE[NC], mO[NC], eO[NC], fO[NC];
    uint32_t w[NC];          // row words of the current rows (slot k: row U2 - gidx_k)
    int cc[NC];              // 4 * column code of the current columns
    uint32_t K0[NC];         // band-test constant of the even diagonal of slot k (odd: K0 - 4095)
    const uint32_t *rp;      // &roww[row_pad + 1 + U2 - lane*NC]      (slot k at rp[-k])
    const uint8_t *cp;       // &colc[col_pad + J2 + lane*NC]          (slot k at cp[k])
    int *ap;                 // &arcbuf[lane*NC]
    int *bp;                 // &box[u * stride + lane*NC]
    int ip0, jp0;            // row / column of slot 0 (generic borders only)
    const int *psig;         // profile pairs (generic borders only): &sigma'(al, bl), row length pld; nullptr: scores by symbol code
    int pld;
    bool st_ok;              // lane*NC < stride: this lane's slots lie inside the box row
};

// One anti-diagonal step. PAR = parity of (u - vmin): the active diagonal of pair g is c = 2g + PAR.
//   PAR == 0: new row (U2 advances): up neighbour = odd diagonal of the same pair, left = odd diagonal of the previous pair
//   PAR == 1: new column (J2 advances): left neighbour = even diagonal of the same pair, up = even diagonal of the next pair
// GB (generic borders): row 0 / column 0 get explicit values; otherwise they fall out of the recurrence, which is exact for
//   the all-global box init with indel_opening <= 0.   CLAMP: sequence-local top level (aligner.cc:866-868).
template <int NC, int PAR, bool GB, bool CLAMP>
__device__ __forceinline__ void dp_step2(Sweep<NC> &S, const BoxInit &init, const int *sig, int ringoff, int stride, int gap, int gap_open, int lane) {
    int xm, xo;
    if (PAR == 0) {
        xm = __shfl_up_sync(0xffffffffu, S.mO[NC - 1], 1);
        xo = __shfl_up_sync(0xffffffffu, S.fO[NC - 1], 1);
        if (lane == 0) { xm = LB_NEG; xo = LB_NEG; }
        S.rp += 1;
        if (GB) S.ip0 += 1;
    } else {
        xm = __shfl_down_sync(0xffffffffu, S.mE[0], 1);
        xo = __shfl_down_sync(0xffffffffu, S.eE[0], 1);
        if (lane == 31) { xm = LB_NEG; xo = LB_NEG; }
        S.cp += 1;
        if (GB) S.jp0 += 1;
    }
    int *arc = S.ap + ringoff;
    int nm[NC], ne[NC], nf[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) {
        int m_up, e_up, m_left, f_left, m_diag;
        if (PAR == 0) {
            S.w[k] = S.rp[-k];
            m_up = S.mO[k]; e_up = S.eO[k]; m_diag = S.mE[k];
            if (k == 0) { m_left = xm; f_left = xo; } else { m_left = S.mO[k > 0 ? k - 1 : 0]; f_left = S.fO[k > 0 ? k - 1 : 0]; }
        } else {
            S.cc[k] = S.cp[k];
            m_left = S.mE[k]; f_left = S.fE[k]; m_diag = S.mO[k];
            if (k == NC - 1) { m_up = xm; e_up = xo; } else { m_up = S.mE[k < NC - 1 ? k + 1 : k]; e_up = S.eE[k < NC - 1 ? k + 1 : k]; }
        }
        const uint32_t w = S.w[k];
        const uint32_t z = S.K0[k] - w - (PAR ? 4095u : 0u);
        const bool ok = (~z & ROWW_GUARDS) == 0;
        int sg = *(const int *)((const char *)sig + (w >> 24) + S.cc[k]);
        if (GB) {   // position-specific base match score of a profile pair (valid cells lie inside the box: 0 <= ip <= Rn, 0 <= jp <= Cn)
            if (S.psig != nullptr) sg = ok ? __ldg(S.psig + (S.ip0 - k) * S.pld + (S.jp0 + k)) : 0;
        }
        int e = addmax(e_up, gap, m_up + gap_open);
        int f = addmax(f_left, gap, m_left + gap_open);
        const int a = arc[k];
        int m = max3(addmax(m_diag, sg, e), f, a);
        if (CLAMP) m = max(m, 0);
        if (GB) {
            const int ip = S.ip0 - k, jp = S.jp0 + k;
            if (ip == 0) { m = (jp == 0) ? 0 : init.row_base + jp * init.row_step; e = LB_NEG; f = LB_NEG; }
            else if (jp == 0) { m = init.col_base + ip * init.col_step; e = LB_NEG; f = LB_NEG; }
        }
        nm[k] = ok ? m : LB_NEG; ne[k] = ok ? e : LB_NEG; nf[k] = ok ? f : LB_NEG;
    }
    // reset the consumed accumulators, store the row of the box (vector stores: lane*NC and the stride are multiples of the width)
    if (NC % 4 == 0) {
#pragma unroll
        for (int k = 0; k < NC; k += 4) {
            *(int4 *)(arc + k) = make_int4(LB_NEG, LB_NEG, LB_NEG, LB_NEG);
            if (NC == 4 ? S.st_ok : (lane * NC + k < stride)) *(int4 *)(S.bp + k) = make_int4(nm[k], nm[k + 1], nm[k + 2], nm[k + 3]);
        }
    } else if (NC % 2 == 0) {
#pragma unroll
        for (int k = 0; k < NC; k += 2) {
            *(int2 *)(arc + k) = make_int2(LB_NEG, LB_NEG);
            if (S.st_ok) *(int2 *)(S.bp + k) = make_int2(nm[k], nm[k + 1]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < NC; k++) {
            arc[k] = LB_NEG;
            if (S.st_ok && (NC == 1 || lane * NC + k < stride)) S.bp[k] = nm[k];
        }
    }
    S.bp += stride;
#pragma unroll
    for (int k = 0; k < NC; k++) {
        if (PAR == 0) { S.mE[k] = nm[k]; S.eE[k] = ne[k]; S.fE[k] = nf[k]; }
        else { S.mO[k] = nm[k]; S.eO[k] = ne[k]; S.fO[k] = nf[k]; }
    }
}

// The arc-match entry stream of one box. The pair's S-order entries are sorted by the anti-diagonal t of their right ends and,
// inside one anti-diagonal, by the anti-diagonal of their source cell M(al'-1, bl'-1) DESCENDING. A box with origin anti-diagonal
// s0 can only use entries whose source lies at or after s0, i.e. a prefix of every anti-diagonal's list - entries of long arcs that
// start before the box are never touched. Per cell step u the warp handles the list of target anti-diagonal u + 2: one 16-byte
// entry per lane (LDG.128, prefetched during the previous step), box filter, M(source) gather left in flight across the next
// cell step, then folded into the ring of per-anti-diagonal accumulators with a shared-memory atomicMax (sources lie >= 8
// anti-diagonals before their target because both arcs span >= 3 positions, so they are final). Lists whose prefix is longer
// than 32 entries continue with further (unprefetched) blocks.
struct Stream3 {
    const uint4 *entl;       // the pair's entries + lane
    const uint2 *entp;       // packed entries + lane (nullptr: the batch is not packed)
    const int *q;            // q[k] = S-order start of local anti-diagonal k (sptr + s0)
    int q_cap;               // largest valid index into q
    int qreg, qbase;         // lane l holds q[qbase + l]: the list bounds are read with one SHFL instead of a load per step
    int qx, qy;              // bounds q[t], q[t+1] of the list prefetched next
    uint4 nA, nB;            // prefetched lists of the next two cell steps (lanes beyond the list hold the poison entry)
    int pm, pd, ps;          // gather in flight: M(source), D, accumulator index (lanes without a hit: -inf into their own slot)
    int pm2, pd2, ps2;       // second gather in flight (packed batches: entries 32..63 of the list)
};

// filter one block of the list of target anti-diagonal t and start the gathers; returns whether lane 31 still belongs to the prefix.
// Lanes beyond the list hold x = 0xffffffff (fails the box filter) and s = -1 (outside every prefix).
template <int NC>
__device__ __forceinline__ bool stream_block(const uint4 v, const BoxGeom &g, const int *box, int ring_t, int s0, uint32_t org, uint32_t lim, int d0,
                                             int lane, int &pm, int &pd, int &ps) {
    const uint32_t t1 = v.x - org, t2 = lim - v.y;
    const bool in = ((t1 | t2) & 0x80008000u) == 0;
    const uint32_t p = t1 & 0xffffu, q = t1 >> 16;
    // branch-free: lanes without a hit fold a dummy value into their own junk slot behind the ring
    // (unconditional load from a safe address: a select after a predicated load would wait for the data inside this step)
    uint32_t idx = (p + q) * (uint32_t)g.nslots + ((q - p - (uint32_t)g.vmin) >> 1);
    idx = in ? idx : 0u;
    pm = __ldcg(box + idx);
    pd = in ? (int)v.z : LB_NEG;
    ps = in ? ring_t + ((LB_ENT_HI(v.y) - LB_ENT_LO(v.y) - d0) >> 1) : RING * RingStride<NC>::v + lane;
    return (__ballot_sync(0xffffffffu, (int)v.w >= s0) >> 31) != 0;
}

// Fill one M box. NC = diagonal pairs per lane (the warp covers 64*NC diagonals).
template <int NC, bool GB, bool CLAMP>
__device__ void fill_box(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const BoxInit &init, const WarpSmem &ws, int *box) {
    const int lane = threadIdx.x & 31;
    constexpr int NWP = RingStride<NC>::v;
    const int gap = c.params.gap, gap_open = c.params.gap_open;
    const int stride = g.nslots;

    for (int k = lane; k < RING * NWP; k += 32) ws.arcbuf[k] = LB_NEG;

    const int s0 = g.al + g.bl;
    Stream3 st;
    const bool packed = c.ent8 != nullptr;
    st.entl = (const uint4 *)(c.ent + pr.am_base) + lane;
    st.entp = packed ? (const uint2 *)(c.ent8 + pr.am_base) + lane : nullptr;
    if (packed) asm("" : "+l"(st.entp)); else asm("" : "+l"(st.entl));   // keep the lane's entry pointer in a register pair: one IMAD.WIDE per prefetch
    const int *boxr = box;
    asm("" : "+l"(boxr));      // likewise for the gathers (the CTA-uniform box address would be rebuilt with 64-bit adds)
    st.q = c.sptr + pr.sptr + s0;
    st.q_cap = pr.lenA + pr.lenB + 2 - s0;
    // list bounds: lane l keeps q[qbase + l]; stream(u) needs q[u+2 .. u+5], the window is moved when u+5 leaves it
    st.qbase = 3;
    st.qreg = __ldg(st.q + min(st.qbase + lane, st.q_cap));
    auto qget = [&](int k) { return __shfl_sync(0xffffffffu, st.qreg, k - st.qbase); };
    // entries e+lane of [e, e_end); lanes beyond the list get the poison fields x (fails the box filter) and w (outside every prefix)
    // (loads go to L2: the D field may have been written by another CTA of this launch). Packed batches prefetch TWO blocks per
    // list into the same four registers (entries e+lane -> v.x/v.y, e+32+lane -> v.z/v.w), so that only prefixes longer than 64
    // entries fall back to synchronous blocks; their poison entry has source (0, 0) and target (511, 511): it fails the box filter
    // and lies in a prefix only for s0 = 0.
    constexpr uint32_t POISON0 = (511u << 18) | (31u << 27), POISON1 = 15u;
    auto eld = [&](uint4 &v, int e, int e_end) {
        if (packed) {
            if (e + lane < e_end) { const uint2 t = __ldcg(st.entp + e); v.x = t.x; v.y = t.y; }
            else { v.x = POISON0; v.y = POISON1; }
            if (e + 32 + lane < e_end) { const uint2 t = __ldcg(st.entp + e + 32); v.z = t.x; v.w = t.y; }
            else { v.z = POISON0; v.w = POISON1; }
        } else {
            if (e + lane < e_end) v = __ldcg(st.entl + e);
            else { v.x = 0xffffffffu; v.w = 0xffffffffu; }
        }
    };
    // packed (w0, w1) -> the 16-byte entry fields
    auto unpack = [&](uint32_t w0, uint32_t w1) {
        uint4 r;
        r.x = (w0 & 0x1ffu) | ((w0 << 7) & 0x01ff0000u);
        r.y = ((w0 >> 18) & 0x1ffu) | ((w0 >> 27) << 16) | ((w1 & 15u) << 21);
        const int d = (int)w1 >> 4;
        r.z = (uint32_t)(d == LB_PACK_NEG ? LB_NEG : d);
        r.w = (r.x & 0xffffu) + (r.x >> 16);
        return r;
    };
    // the first list handled (after cell step 1) is the one of local anti-diagonal 3
    {
        const int q3 = qget(3), q4 = qget(4);
        st.qx = qget(5); st.qy = qget(6);
        st.nA = make_uint4(0xffffffffu, 0u, 0u, 0xffffffffu); st.nB = st.nA;
        eld(st.nA, q3, q4);
        eld(st.nB, q4, st.qx);
    }
    st.ps = RING * NWP + lane; st.pm = 0; st.pd = 0;
    st.ps2 = RING * NWP + lane; st.pm2 = 0; st.pd2 = 0;
    const uint32_t org = (uint32_t)g.al | ((uint32_t)g.bl << 16);
    const uint32_t lim = (uint32_t)(g.al + g.Rn) | ((uint32_t)(g.bl + g.Cn) << 16);
    const int d0 = g.bl - g.al + g.vmin;

    Sweep<NC> S;
#pragma unroll
    for (int k = 0; k < NC; k++) { S.mE[k] = S.eE[k] = S.fE[k] = S.mO[k] = S.eO[k] = S.fO[k] = LB_NEG; }
    const int par0 = (0 - g.vmin) & 1;   // parity class of anti-diagonal 0
    const int U2 = (-g.vmin - par0) >> 1, J2 = (g.vmin + par0) >> 1;   // row / column of diagonal pair 0 "at u = 0"
    S.rp = ws.roww + c.row_pad + 1 + U2 - lane * NC;
    S.cp = ws.colc + c.col_pad + J2 + lane * NC;
    S.ap = ws.arcbuf + lane * NC;
    S.bp = box + stride + lane * NC;     // row u = 1
    S.ip0 = U2 - lane * NC; S.jp0 = J2 + lane * NC;
    S.psig = (GB && c.ps_sig != nullptr && pr.ps_sig >= 0) ? c.ps_sig + pr.ps_sig + (long long)g.al * (pr.lenB + 1) + g.bl : nullptr;
    S.pld = pr.lenB + 1;
    S.st_ok = lane * NC < stride;
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const uint32_t cdiag = 2 * (lane * NC + k);
        S.K0[k] = (cdiag | 0x800u) | (((2047u - cdiag) | 0x800u) << 12);
        S.w[k] = S.rp[-k];
        S.cc[k] = S.cp[k];
    }
    // seed the origin M(al, bl) = 0 (aligner.cc:289): diagonal v = 0, i.e. c = -vmin
    {
        const int c0 = -g.vmin;
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (2 * (lane * NC + k) + par0 == c0) { if (par0 == 0) S.mE[k] = 0; else S.mO[k] = 0; box[c0 >> 1] = 0; }
        }
    }
    __syncwarp();

    int ringoff = ((s0 + 1) & (RING - 1)) * NWP;
    // after cell step u: land the gather of the previous step, handle the list of anti-diagonal u + 2 (buffer nx, loaded two steps
    // ago), refill nx with the list of anti-diagonal u + 4
    auto stream = [&](int u, uint4 &nx) {
        __syncwarp();
        atomicMax(&ws.arcbuf[st.ps], st.pm + st.pd);
        const int ring_t = ((s0 + u + 2) & (RING - 1)) * NWP;
        bool more;
        if (packed) {
            atomicMax(&ws.arcbuf[st.ps2], st.pm2 + st.pd2);
            more = stream_block<NC>(unpack(nx.x, nx.y), g, boxr, ring_t, s0, org, lim, d0, lane, st.pm, st.pd, st.ps);
            if (more) more = stream_block<NC>(unpack(nx.z, nx.w), g, boxr, ring_t, s0, org, lim, d0, lane, st.pm2, st.pd2, st.ps2);
            else st.ps2 = RING * NWP + lane;   // nothing in flight: the landing of the next stage goes to the junk slot
        } else more = stream_block<NC>(nx, g, boxr, ring_t, s0, org, lim, d0, lane, st.pm, st.pd, st.ps);
        if (more) {   // prefix longer than the prefetched blocks: further, synchronous blocks
            const int qa = qget(u + 2), qb = qget(u + 3);
            for (int e = qa + (packed ? 64 : 32); more && e < qb; e += 32) {
                uint4 v = make_uint4(0xffffffffu, 0u, 0u, 0xffffffffu);
                if (packed) {
                    uint2 t = make_uint2(POISON0, POISON1);
                    if (e + lane < qb) t = __ldcg(st.entp + e);
                    v = unpack(t.x, t.y);
                } else if (e + lane < qb) v = __ldcg(st.entl + e);
                atomicMax(&ws.arcbuf[st.ps], st.pm + st.pd);
                more = stream_block<NC>(v, g, boxr, ring_t, s0, org, lim, d0, lane, st.pm, st.pd, st.ps);
            }
        }
        eld(nx, st.qx, st.qy);                // list of anti-diagonal u + 4: [q[u+4], q[u+5])
        // bounds for the next call: q[u+5], q[u+6]; move the window (to start at u+3) when u+6 leaves it
        if (u + 6 - st.qbase >= 32) { st.qbase = u + 3; st.qreg = __ldg(st.q + min(st.qbase + lane, st.q_cap)); }
        st.qx = st.qy; st.qy = qget(u + 6);
        __syncwarp();
    };
    int u = 1;
    if (par0 == 0) {   // anti-diagonal 1 is of the odd class
        dp_step2<NC, 1, GB, CLAMP>(S, init, ws.sig, ringoff, stride, gap, gap_open, lane);
        ringoff = (ringoff + NWP) & (RING * NWP - 1);
        stream(u, st.nA);
        u = 2;
        // keep the buffer roles aligned with the unrolled loop below (nA first)
        const uint4 t = st.nA; st.nA = st.nB; st.nB = t;
    }
    for (; u + 1 <= g.umax; u += 2) {
        dp_step2<NC, 0, GB, CLAMP>(S, init, ws.sig, ringoff, stride, gap, gap_open, lane);
        ringoff = (ringoff + NWP) & (RING * NWP - 1);
        stream(u, st.nA);
        dp_step2<NC, 1, GB, CLAMP>(S, init, ws.sig, ringoff, stride, gap, gap_open, lane);
        ringoff = (ringoff + NWP) & (RING * NWP - 1);
        stream(u + 1, st.nB);
    }
    if (u <= g.umax) {
        dp_step2<NC, 0, GB, CLAMP>(S, init, ws.sig, ringoff, stride, gap, gap_open, lane);
        __syncwarp();
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Structure-local boxes (aligner.cc:398-418 init, :443-564 fill): eight matrices per box.
//   closed states 0..3 = E_NO_NO, E_X_NO, E_NO_X, E_X_X run the align_noex recurrence with their own E/F and their own
//   arc-match sources M_k(al'-1, bl'-1); open states E_OP_NO, E_NO_OP, E_OP_X, E_X_OP (here 0..3 of the o-arrays) are running
//   maxima along a column / row. All cross terms are cell-local or read the up / left neighbour, so one anti-diagonal sweep
//   computes the eight matrices together, per cell in the reference's matrix order
//   NO_NO, OP_NO, NO_OP, NO_X, OP_X, X_NO, X_OP, X_X.
// Boxes: box_k = boxes + k * box_words, k = 0..3 closed states, 4..7 = OP_NO, NO_OP, OP_X, X_OP (only if STORE_OPEN, for the
// traceback). Four accumulator rings (one per closed state) in ws.arcbuf.
struct SlBorder { int col[8], col_step[8], row[8], row_step[8]; };

__device__ __forceinline__ void sl_borders(const DevParams &P, SlBorder &b) {
    // init_state(state, globalA, exclA, globalB, exclB): aligner.cc:398-417
    const bool gA[8] = {true, true, true, true, false, true, false, true};     // NO_NO X_NO NO_X X_X OP_NO NO_OP OP_X X_OP
    const bool xA[8] = {false, true, false, true, false, false, false, true};
    const bool gB[8] = {true, true, true, true, true, false, true, false};
    const bool xB[8] = {false, false, true, true, false, false, true, false};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        b.col[k] = xA[k] ? P.exclusion : (gA[k] ? P.open : 0); b.col_step[k] = (!xA[k] && gA[k]) ? P.gap : 0;
        b.row[k] = xB[k] ? P.exclusion : (gB[k] ? P.open : 0); b.row_step[k] = (!xB[k] && gB[k]) ? P.gap : 0;
    }
}

template <int NC, int PAR, bool STORE_OPEN>
__device__ __forceinline__ void dp_step_sl(const BoxGeom &g, const SlBorder &bd, const WarpSmem &ws, int *boxes, int box_words,
                                           const DevParams &P, int u, int lane, int (&mE)[4][NC], int (&eE)[4][NC], int (&fE)[4][NC],
                                           int (&mO)[4][NC], int (&eO)[4][NC], int (&fO)[4][NC], int (&oE)[4][NC], int (&oO)[4][NC]) {
    constexpr int NW = 32 * NC;
    // open-state order in the o-arrays: 0 OP_NO (up), 1 NO_OP (left), 2 OP_X (up), 3 X_OP (left)
    int xm[4], xo[4], xop[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (PAR == 0) {
            xm[k] = __shfl_up_sync(0xffffffffu, mO[k][NC - 1], 1);
            xo[k] = __shfl_up_sync(0xffffffffu, fO[k][NC - 1], 1);
            xop[k] = __shfl_up_sync(0xffffffffu, oO[k][NC - 1], 1);
            if (lane == 0) { xm[k] = LB_NEG; xo[k] = LB_NEG; xop[k] = LB_NEG; }
        } else {
            xm[k] = __shfl_down_sync(0xffffffffu, mE[k][0], 1);
            xo[k] = __shfl_down_sync(0xffffffffu, eE[k][0], 1);
            xop[k] = __shfl_down_sync(0xffffffffu, oE[k][0], 1);
            if (lane == 31) { xm[k] = LB_NEG; xo[k] = LB_NEG; xop[k] = LB_NEG; }
        }
    }
    const int U2 = (u - g.vmin - PAR) >> 1, J2 = (u + g.vmin + PAR) >> 1;
    const int ring = (u & (RING - 1)) * NW;
    const int ex = P.exclusion;
    int nm[4][NC], ne[4][NC], nf[4][NC], no[4][NC];
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const int gidx = lane * NC + k;
        const int ip = U2 - gidx, jp = J2 + gidx;
        const uint32_t ridx = min((uint32_t)(ip + 1), (uint32_t)(g.Rn + 2));
        const uint32_t cidx = min((uint32_t)jp, (uint32_t)(g.Cn + 1));
        const uint32_t rr = ws.rowrange[ridx];
        const int sg = ws.sig[ws.rowcode[ridx] + ws.colcode[cidx]];
        const bool ok = (uint32_t)(2 * gidx + PAR - (int)(rr & 0xffff)) <= (rr >> 16);
        const bool brow = ip == 0, bcol = (jp == 0) & (ip != 0);
        int cm[4];  // closed states of this cell before the cross terms (align_noex)
#pragma unroll
        for (int st = 0; st < 4; st++) {
            int m_up, e_up, m_left, f_left, m_diag;
            if (PAR == 0) {
                m_up = mO[st][k]; e_up = eO[st][k]; m_diag = mE[st][k];
                if (k == 0) { m_left = xm[st]; f_left = xo[st]; } else { m_left = mO[st][k > 0 ? k - 1 : 0]; f_left = fO[st][k > 0 ? k - 1 : 0]; }
            } else {
                m_left = mE[st][k]; f_left = fE[st][k]; m_diag = mO[st][k];
                if (k == NC - 1) { m_up = xm[st]; e_up = xo[st]; } else { m_up = mE[st][k < NC - 1 ? k + 1 : k]; e_up = eE[st][k < NC - 1 ? k + 1 : k]; }
            }
            int e = addmax(e_up, P.gap, m_up + P.gap_open);
            int f = addmax(f_left, P.gap, m_left + P.gap_open);
            int *acc = ws.arcbuf + st * (RING * NW) + ring + gidx;
            const int arc = *acc;
            *acc = LB_NEG;
            cm[st] = max3(addmax(m_diag, sg, e), f, arc);
            ne[st][k] = e; nf[st][k] = f;
        }
        // neighbours of the open states
        int o_up0, o_up2, o_left1, o_left3;
        if (PAR == 0) {
            o_up0 = oO[0][k]; o_up2 = oO[2][k];
            if (k == 0) { o_left1 = xop[1]; o_left3 = xop[3]; } else { o_left1 = oO[1][k > 0 ? k - 1 : 0]; o_left3 = oO[3][k > 0 ? k - 1 : 0]; }
        } else {
            o_left1 = oE[1][k]; o_left3 = oE[3][k];
            if (k == NC - 1) { o_up0 = xop[0]; o_up2 = xop[2]; } else { o_up0 = oE[0][k < NC - 1 ? k + 1 : k]; o_up2 = oE[2][k < NC - 1 ? k + 1 : k]; }
        }
        // matrix order of the reference: NO_NO, OP_NO, NO_OP, NO_X, OP_X, X_NO, X_OP, X_X (aligner.cc:443-563)
        int m0 = cm[0];
        int op_no = max(o_up0, m0);
        int no_op = max(o_left1, m0);
        int m2 = max(cm[2], no_op + ex);       // E_NO_X
        int op_x = max(o_up2, m2);
        int m1 = max(cm[1], op_no + ex);       // E_X_NO
        int x_op = max(o_left3, m1);
        int m3 = max3(cm[3], op_x + ex, x_op + ex);  // E_X_X
        int mm[4] = {m0, m1, m2, m3};
        int oo[4] = {op_no, no_op, op_x, x_op};
        if (brow | bcol) {  // explicit borders (init_state); the origin is 0 in every state
            const int t = brow ? jp : ip;
#pragma unroll
            for (int st = 0; st < 4; st++) {
                mm[st] = (brow && jp == 0) ? 0 : (brow ? bd.row[st] + t * bd.row_step[st] : bd.col[st] + t * bd.col_step[st]);
                oo[st] = (brow && jp == 0) ? 0 : (brow ? bd.row[4 + st] + t * bd.row_step[4 + st] : bd.col[4 + st] + t * bd.col_step[4 + st]);
                ne[st][k] = LB_NEG; nf[st][k] = LB_NEG;
            }
        }
#pragma unroll
        for (int st = 0; st < 4; st++) {
            if (ok) {
                boxes[st * box_words + u * g.nslots + gidx] = mm[st];
                if (STORE_OPEN) boxes[(4 + st) * box_words + u * g.nslots + gidx] = oo[st];
            }
            nm[st][k] = ok ? mm[st] : LB_NEG; no[st][k] = ok ? oo[st] : LB_NEG;
            if (!ok) { ne[st][k] = LB_NEG; nf[st][k] = LB_NEG; }
        }
    }
#pragma unroll
    for (int st = 0; st < 4; st++)
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (PAR == 0) { mE[st][k] = nm[st][k]; eE[st][k] = ne[st][k]; fE[st][k] = nf[st][k]; oE[st][k] = no[st][k]; }
            else { mO[st][k] = nm[st][k]; eO[st][k] = ne[st][k]; fO[st][k] = nf[st][k]; oO[st][k] = no[st][k]; }
        }
}

// Fill the eight matrices of one structure-local box. The arc-match entries are folded synchronously per step (this path
// serves single long pairs, not the batched all-vs-all workload).
template <int NC, bool STORE_OPEN>
__device__ void fill_box_sl(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const WarpSmem &ws, int *boxes, int box_words) {
    const int lane = threadIdx.x & 31;
    constexpr int NW = 32 * NC;
    const DevParams &P = c.params;
    SlBorder bd;
    sl_borders(P, bd);
    for (int k = lane; k < 4 * RING * NW; k += 32) ws.arcbuf[k] = LB_NEG;
    const int *sptr = c.sptr + pr.sptr;
    const int s_base = g.al + g.bl, s_last = pr.lenA + pr.lenB + 1;
    auto sp = [&](int t) { return __ldg(sptr + min(s_base + min(t, g.umax + 1), s_last)); };
    const DevEntry *ent = c.ent + pr.am_base;
    int mE[4][NC], eE[4][NC], fE[4][NC], mO[4][NC], eO[4][NC], fO[4][NC], oE[4][NC], oO[4][NC];
#pragma unroll
    for (int st = 0; st < 4; st++)
#pragma unroll
        for (int k = 0; k < NC; k++) { mE[st][k] = eE[st][k] = fE[st][k] = mO[st][k] = eO[st][k] = fO[st][k] = oE[st][k] = oO[st][k] = LB_NEG; }
    const int par0 = (0 - g.vmin) & 1;
    {
        const int c0 = -g.vmin;
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (2 * (lane * NC + k) + par0 == c0) {
#pragma unroll
                for (int st = 0; st < 4; st++) {
                    if (par0 == 0) { mE[st][k] = 0; oE[st][k] = 0; } else { mO[st][k] = 0; oO[st][k] = 0; }
                    boxes[st * box_words + (c0 >> 1)] = 0;
                    if (STORE_OPEN) boxes[(4 + st) * box_words + (c0 >> 1)] = 0;
                }
            }
        }
    }
    __syncwarp();
    for (int u = 1; u <= g.umax; u++) {
        // arc-match terms of anti-diagonal u: sources lie >= 8 anti-diagonals back
        {
            const int e0 = sp(u), e1 = sp(u + 1);
            const int tring = (u & (RING - 1)) * NW;
            for (int e = e0 + lane; e < e1; e += 32) {
                const DevEntry en = ent[e];
                const int p = LB_ENT_LO(en.x) - g.al, q = LB_ENT_HI(en.x) - g.bl;
                const int ar = LB_ENT_LO(en.y) - g.al, br = LB_ENT_HI(en.y) - g.bl;
                if ((p | q | (g.Rn - ar) | (g.Cn - br)) >= 0) {
                    const int d = en.d;
                    const int src = (p + q) * g.nslots + ((q - p - g.vmin) >> 1);
                    const int slot = tring + ((br - ar - g.vmin) >> 1);
#pragma unroll
                    for (int st = 0; st < 4; st++) atomicMax(&ws.arcbuf[st * (RING * NW) + slot], boxes[st * box_words + src] + d);
                }
            }
        }
        __syncwarp();
        if (((u + par0) & 1) == 0) dp_step_sl<NC, 0, STORE_OPEN>(g, bd, ws, boxes, box_words, P, u, lane, mE, eE, fE, mO, eO, fO, oE, oO);
        else dp_step_sl<NC, 1, STORE_OPEN>(g, bd, ws, boxes, box_words, P, u, lane, mE, eE, fE, mO, eO, fO, oE, oO);
        __syncwarp();
    }
}

template <bool STORE_OPEN>
__device__ bool run_box_sl(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const WarpSmem &ws, int *boxes, int box_words) {
    const int nc = (g.nslots + 31) >> 5;
    if (nc <= 1) fill_box_sl<1, STORE_OPEN>(c, pr, g, ws, boxes, box_words);
    else if (nc <= 2) fill_box_sl<2, STORE_OPEN>(c, pr, g, ws, boxes, box_words);
    else if (nc <= 3) fill_box_sl<3, STORE_OPEN>(c, pr, g, ws, boxes, box_words);
    else if (nc <= 4) fill_box_sl<4, STORE_OPEN>(c, pr, g, ws, boxes, box_words);
    else return false;
    return true;
}

__device__ __forceinline__ int box_get(const int *box, const BoxGeom &g, int ip, int jp) {
    return box[(ip + jp) * g.nslots + ((jp - ip - g.vmin) >> 1)];
}

// Dispatch on the number of diagonal pairs per lane (NCMAX bounds the instantiated variants and thereby
// the register footprint of the kernel). Returns false if the band is wider than supported.
template <int NCMAX, bool GB, bool CLAMP>
__device__ bool run_box(const DevCtx &c, const DevPair &pr, const BoxGeom &g, const BoxInit &init, const WarpSmem &ws, int *box) {
    const int nc = (g.nslots + 31) >> 5;
    if (nc <= 1) fill_box<1, GB, CLAMP>(c, pr, g, init, ws, box);
    else if (NCMAX >= 2 && nc <= 2) fill_box<(NCMAX >= 2 ? 2 : 1), GB, CLAMP>(c, pr, g, init, ws, box);
    else if (NCMAX >= 3 && nc <= 3) fill_box<(NCMAX >= 3 ? 3 : 1), GB, CLAMP>(c, pr, g, init, ws, box);
    else if (NCMAX >= 4 && nc <= 4) fill_box<(NCMAX >= 4 ? 4 : 1), GB, CLAMP>(c, pr, g, init, ws, box);
    else if (NCMAX >= 8 && nc <= 8) fill_box<(NCMAX >= 8 ? 8 : 1), GB, CLAMP>(c, pr, g, init, ws, box);
    else if (NCMAX >= 16 && nc <= 16) fill_box<(NCMAX >= 16 ? 16 : 1), GB, CLAMP>(c, pr, g, init, ws, box);
    else return false;
    return true;
}

// shared memory layout of a one-warp CTA: [sig 64][rowrange max_rows+2][rowcode max_rows+2 (bytes, padded)][colcode max_cols+1 (bytes, padded)][arcbuf]
__device__ __forceinline__ void carve(const DevCtx &c, int *smem, WarpSmem &ws) {
    int *sig = smem;
    ws.sig = sig;
    ws.rowrange = (uint32_t *)(smem + 64);
    ws.rowcode = (uint8_t *)(ws.rowrange + c.max_rows + 2);
    ws.colcode = ws.rowcode + c.rowcode_bytes;
    ws.roww = (uint32_t *)(smem + 64);
    ws.colc = (uint8_t *)(ws.roww + c.row_words);
    ws.arcbuf = (int *)((char *)(smem + 64) + c.region_bytes);
    const int lane = threadIdx.x & 31;
    sig[lane] = c.params.sigma8[lane]; sig[lane + 32] = c.params.sigma8[lane + 32];
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// D-fill kernel: persistent one-warp CTAs pull the tasks of one scheduling level from an atomic cursor.
// Tasks of a level are mutually independent: a task reads D only for arc matches strictly inside its
// box, whose left ends have a larger al+bl (aligner.cc:675-728; levels = al+bl descending, two at a time).
// One D-fill task: the M box of the left ends (al, bl), then the D entries of all arc matches with these left ends
// (aligner.cc:574-657). D values are read and written through L2 (ld.cg / plain stores): with the dependency-driven
// schedule they are produced by other CTAs of the same launch.
template <int NCMAX, bool GB>
__device__ __forceinline__ void dfill_task(const DevCtx &c, const DevTask &task, const DevPair &pr, const BoxGeom &g, const BoxInit &init,
                                           const WarpSmem &ws, int *box, bool nolp, int lane) {
    if ((g.umax + 1) * g.nslots > c.scratch_words) { if (lane == 0) atomicExch(c.error_flag, 2); return; }
    if (!run_box<NCMAX, GB, false>(c, pr, g, init, ws, box)) { if (lane == 0) atomicExch(c.error_flag, 1); return; }
    const DevArcMatch *am = c.am + pr.am_base;
    DevEntry *ent = c.ent + pr.am_base;
    const int sh = nolp ? 2 : 1;
    const bool stacking = c.params.stacking != 0;
    for (int k = task.run_start + lane; k < task.run_start + task.run_count; k += 32) {
        const DevArcMatch x = am[k];
        if (nolp && (x.inner < 0 || (stacking && x.score_st == LB_NOSTACK))) continue;   // aligner.cc:628-629
        const int ar = (x.ends_a >> 12) & 0xfff, br = (x.ends_b >> 12) & 0xfff;
        const int mv = box_get(box, g, ar - sh - g.al, br - sh - g.bl);
        int d;
        if (nolp) {
            const DevArcMatch in = am[x.inner];
            const int a = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + in.score;
            const int y = max(a, __ldcg(&ent[in.spos].d));
            d = (y < LB_NEG_LIMIT) ? LB_NEG : y + (stacking ? x.score_st : x.score);
        } else {
            d = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + x.score;
            if (stacking && x.inner >= 0 && x.score_st != LB_NOSTACK) {   // aligner.cc:600-607
                const int di = __ldcg(&ent[am[x.inner].spos].d);
                if (di >= LB_NEG_LIMIT) d = max(d, di + x.score_st);
            }
        }
        ent[x.spos].d = d;
        if (c.ent8 != nullptr) c.ent8[pr.am_base + x.spos].y = LB_PACK_W1(br, d);
    }
}

template <int NCMAX, bool GB>
__global__ void __launch_bounds__(32, 32) dfill_kernel(DevCtx c, int q) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    const int task_begin = c.qstart[q], task_end = c.qstart[q + 1];
    if ((int)blockIdx.x >= task_end - task_begin) return;
    int *cursor = c.cursor + q;
    WarpSmem ws;
    carve(c, smem, ws);
    int *box = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const bool nolp = c.params.no_lonely_pairs != 0;
    BoxInit init;
    init.col_base = c.params.open; init.col_step = c.params.gap; init.row_base = c.params.open; init.row_step = c.params.gap;

    for (;;) {
        int t = 0;
        if (lane == 0) t = task_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= task_end) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        BoxGeom g;
        setup_box2(c, pr, task.al, task.bl, task.R, task.C, g, ws);
        dfill_task<NCMAX, GB>(c, task, pr, g, init, ws, box, nolp, lane);
        __syncwarp();
    }
}

// Dependency-driven D fill: ONE persistent launch per batch. Tasks are claimed in the order dep_order[] (blocks of pairs; inside a
// block level group descending, box area descending). A task of level group L reads D entries only of arc matches of ITS OWN PAIR
// in level groups > L, so instead of a grid-wide barrier per level it waits until dep_done[pair] reaches dep_need[pair][L]
// (all tasks of the pair in higher groups). Every task it waits for precedes it in the claim order and is therefore already held
// by a resident warp: no deadlock, no co-residency requirement. Grouping by pair blocks keeps the entry tables of the pairs in
// flight (a few hundred pairs x 0.4 MB) inside the 126 MB L2 instead of cycling through the whole batch once per level.
template <int NCMAX, bool GB>
__global__ void __launch_bounds__(32, 32) dfill_dep_kernel(DevCtx c, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    WarpSmem ws;
    carve(c, smem, ws);
    int *box = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const bool nolp = c.params.no_lonely_pairs != 0;
    BoxInit init;
    init.col_base = c.params.open; init.col_step = c.params.gap; init.row_base = c.params.open; init.row_step = c.params.gap;

    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(cursor, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= c.n_tasks) break;
        const DevTask task = c.tasks[c.dep_order[k]];
        const DevPair pr = c.pairs[task.pair];
        BoxGeom g;
        setup_box2(c, pr, task.al, task.bl, task.R, task.C, g, ws);   // reads only the band and the sequences: overlaps the wait
        if (lane == 0) {
            const int need = c.dep_need[(size_t)task.pair * c.n_groups + (((int)task.al + (int)task.bl) >> 1)];
            const int *done = c.dep_done + task.pair;
            int have;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(done) : "memory");
                if (have >= need) break;
                __nanosleep(200);
            }
        }
        __syncwarp();
        dfill_task<NCMAX, GB>(c, task, pr, g, init, ws, box, nolp, lane);
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(c.dep_done + task.pair, 1);
    }
}

// D-fill for --struct-local: D = max over the four closed states (aligner.cc:586-595, :632-640)
__global__ void __launch_bounds__(32) dfill_sl_kernel(DevCtx c, int q) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    const int task_begin = c.qstart[q], task_end = c.qstart[q + 1];
    if ((int)blockIdx.x >= task_end - task_begin) return;
    int *cursor = c.cursor + q;
    WarpSmem ws;
    carve(c, smem, ws);
    const int box_words = c.scratch_words / 8;
    int *boxes = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const bool nolp = c.params.no_lonely_pairs != 0;
    for (;;) {
        int t = 0;
        if (lane == 0) t = task_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= task_end) break;
        const DevTask task = c.tasks[t];
        const DevPair pr = c.pairs[task.pair];
        BoxGeom g;
        setup_box(c, pr, task.al, task.bl, task.R, task.C, g, ws);
        if ((g.umax + 1) * g.nslots > box_words) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box_sl<false>(c, pr, g, ws, boxes, box_words)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        const DevArcMatch *am = c.am + pr.am_base;
        DevEntry *ent = c.ent + pr.am_base;
        const int sh = nolp ? 2 : 1;
        const bool stacking = c.params.stacking != 0;
        for (int k = task.run_start + lane; k < task.run_start + task.run_count; k += 32) {
            const DevArcMatch x = am[k];
            if (nolp && (x.inner < 0 || (stacking && x.score_st == LB_NOSTACK))) continue;
            const int ar = (x.ends_a >> 12) & 0xfff, br = (x.ends_b >> 12) & 0xfff;
            int mv = LB_NEG;
#pragma unroll
            for (int st = 0; st < 4; st++) mv = max(mv, box_get(boxes + st * box_words, g, ar - sh - g.al, br - sh - g.bl));
            int d;
            if (nolp) {
                const DevArcMatch in = am[x.inner];
                const int a = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + in.score;
                const int y = max(a, ent[in.spos].d);
                d = (y < LB_NEG_LIMIT) ? LB_NEG : y + (stacking ? x.score_st : x.score);
            } else {
                d = (mv < LB_NEG_LIMIT) ? LB_NEG : mv + x.score;
                if (stacking && x.inner >= 0 && x.score_st != LB_NOSTACK) {
                    const int di = ent[am[x.inner].spos].d;
                    if (di >= LB_NEG_LIMIT) d = max(d, di + x.score_st);
                }
            }
            ent[x.spos].d = d;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Top level (aligner.cc:736-880): one box over the whole band with al = bl = 0, then the score.
template <int NCMAX, bool CLAMP>
__global__ void __launch_bounds__(32) toplevel_kernel(DevCtx c, int pair_begin, int pair_end, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    WarpSmem ws;
    carve(c, smem, ws);
    int *box = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const DevParams &P = c.params;
    BoxInit init;
    // init_state(E_NO_NO, 0, lenA+1, 0, lenB+1, !allow_left_2, false, !allow_left_1, false) resp. all-local
    const bool globalA = !(P.sequ_local || P.fe_left2), globalB = !(P.sequ_local || P.fe_left1);
    init.col_base = globalA ? P.open : 0; init.col_step = globalA ? P.gap : 0;
    init.row_base = globalB ? P.open : 0; init.row_step = globalB ? P.gap : 0;
    for (;;) {
        int t = 0;
        if (lane == 0) t = pair_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= pair_end) break;
        const DevPair pr = c.pairs[t];
        BoxGeom g;
        // the box of the (restricted) top level: origin (sa - 1, sb - 1), last row / column (ea, eb) (aligner.cc:743-745, :832-833)
        const int n = c.r_on ? c.r_ea : pr.lenA, m = c.r_on ? c.r_eb : pr.lenB;
        const int sa = c.r_on ? c.r_sa : 1, sb = c.r_on ? c.r_sb : 1, al0 = sa - 1, bl0 = sb - 1;
        setup_box2(c, pr, al0, bl0, n, m, g, ws);
        if ((g.umax + 1) * g.nslots > c.scratch_words) { if (lane == 0) atomicExch(c.error_flag, 2); continue; }
        if (!run_box<NCMAX, true, CLAMP>(c, pr, g, init, ws, box)) { if (lane == 0) atomicExch(c.error_flag, 1); continue; }
        const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
        // band cell of the top level box, normalised -inf
        auto cell = [&](int i, int j) -> int {
            if (j < ((i == al0) ? bl0 : max(lo[i], bl0)) || j > min(m, hi[i])) return LB_NEG;
            const int v = box_get(box, g, i - al0, j - bl0);
            return v < LB_NEG_LIMIT ? LB_NEG : v;
        };
        // candidates are ranked by (score, earlier position in the reference's scan order)
        int best = LB_NEG, bi = al0, bj = bl0;
        long long bkey = 0x7fffffffffffffffLL;  // smaller = earlier in scan order
        auto consider = [&](int v, int i, int j, long long key) {
            if (v > best || (v == best && v > LB_NEG && key < bkey)) { best = v; bi = i; bj = j; bkey = key; }
        };
        if (P.sequ_local) {
            // first strict maximum in row-major order, initial best 0 at (sa - 1, sb - 1) (aligner.cc:829-876)
            best = 0; bkey = -1;
            for (int i = sa; i <= n; i++) {
                const int jl = max(sb, lo[i]), jh = min(m, hi[i]);
                for (int j = jl + lane; j <= jh; j += 32) consider(cell(i, j), i, j, (long long)i * (pr.lenB + 1) + j);
            }
        } else if (P.fe_right1 || P.fe_right2) {
            // aligner.cc:775-816: last column scanned by rows first, then the last row by columns; strict >
            if (P.fe_right2) {
                for (int i = sa + lane; i <= n; i += 32)
                    if (hi[i] >= m) consider(cell(i, m), i, m, (long long)i);
            }
            if (P.fe_right1) {
                const int jl = max(sb, lo[n]), jh = min(m, hi[n]);
                for (int j = jl + lane; j <= jh; j += 32) consider(cell(n, j), n, j, (long long)pr.lenA + 1 + j);
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
            r.score = best; r.max_i = bi; r.max_j = bj; r.min_ij = 0;
            if (!P.sequ_local && (P.fe_right1 || P.fe_right2) && best <= LB_NEG_LIMIT) { r.max_i = al0; r.max_j = bl0; }
            c.top[t] = r;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback (aligner.cc:967-1358). One warp per pair works off a stack of boxes: the top level box first, then the
// box of every arc match found on the optimal path (the reference re-runs align_in_arcmatch for each of them,
// aligner.cc:1012/:1056). Inside a box the warp walks back from the last cell following trace_noex's case order
// (base match, deletion run, insertion run, arc matches in common_right_end_list order; first equality wins).
// Every alignment edge is written to slot i+j of the pair's edge array - i+j strictly increases along an alignment,
// so the order in which boxes are processed does not matter and the host only compacts the array.
#define LB_EDGE_MATCH 1
#define LB_EDGE_DEL 2   // (i, gap)
#define LB_EDGE_INS 3   // (gap, j)

template <int NCMAX, bool GBD, bool CLAMP>
__global__ void __launch_bounds__(32) trace_kernel(DevCtx c, int pair_begin, int pair_end, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    WarpSmem ws;
    carve(c, smem, ws);
    int *box = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const DevParams &P = c.params;
    const bool nolp = P.no_lonely_pairs != 0;
    const bool mod = c.use_tl != 0;            // normalized / penalized: the top level box uses the modified scoring
    const DevParams &PT = mod ? c.params_tl : c.params;
    BoxInit top_init, in_init;
    const bool globalA = !(P.sequ_local || P.fe_left2), globalB = !(P.sequ_local || P.fe_left1);
    top_init.col_base = globalA ? PT.open : 0; top_init.col_step = globalA ? PT.gap : 0;
    top_init.row_base = globalB ? PT.open : 0; top_init.row_step = globalB ? PT.gap : 0;
    in_init.col_base = P.open; in_init.col_step = P.gap; in_init.row_base = P.open; in_init.row_step = P.gap;
    for (;;) {
        int t = 0;
        if (lane == 0) t = pair_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= pair_end) break;
        const DevPair pr = c.pairs[t];
        const int n = pr.lenA, m = pr.lenB;
        const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
        const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
        const int *psig = (c.ps_sig != nullptr && pr.ps_sig >= 0) ? c.ps_sig + pr.ps_sig : nullptr;   // profile pair: sigma'(i, j) table
        const DevEntry *ent = c.ent + pr.am_base;
        const DevArcMatch *am = c.am + pr.am_base;
        const unsigned *lpos = c.lpos + pr.am_base;
        const int *sptr = c.sptr + pr.sptr;
        int *edges = c.trace_edges + pr.sptr;                // n + m + 3 slots
        char *strA = c.trace_str + pr.sptr, *strB = strA + n + 1;  // n + 1 and m + 1 bytes (sptr offsets are a superset)
        TraceJob *stack = c.trace_stack + (size_t)t * c.trace_stack_cap;
        for (int k = lane; k < n + m + 3; k += 32) edges[k] = 0;
        for (int k = lane; k <= n; k += 32) strA[k] = '.';
        for (int k = lane; k <= m; k += 32) strB[k] = '.';
        __syncwarp();
        auto valid = [&](int i, int j) { return i >= 0 && j >= 0 && lo[i] <= j && j <= hi[i]; };
        auto emit = [&](int i, int j, int kind) { if (lane == 0) edges[i + j] = (i << 2) | kind; };
        int sp = 0;
        const DevTopResult top = c.top[t];
        bool tl = true;
        TraceJob job;
        job.al = (short)(c.r_on ? c.r_sa - 1 : 0); job.bl = (short)(c.r_on ? c.r_sb - 1 : 0);
        job.R = (short)(c.r_on ? c.r_ea : n); job.C = (short)(c.r_on ? c.r_eb : m); job.am = -1;
        bool have = true;
        while (have) {
            BoxGeom g;
            setup_box2(c, pr, job.al, job.bl, job.R, job.C, g, ws);
            bool ok;
            if ((g.umax + 1) * g.nslots > c.scratch_words) ok = false;
            else if (tl && mod) {
                DevCtx ct = c;
                ct.params = c.params_tl; ct.ent = const_cast<DevEntry *>(c.ent_tl); ct.ent8 = const_cast<uint2 *>(c.ent8_tl);
                int *sg = const_cast<int *>(ws.sig);
                sg[lane] = ct.params.sigma8[lane]; sg[lane + 32] = ct.params.sigma8[lane + 32];
                __syncwarp();
                ok = run_box<NCMAX, true, CLAMP>(ct, pr, g, top_init, ws, box);
                __syncwarp();
                sg[lane] = c.params.sigma8[lane]; sg[lane + 32] = c.params.sigma8[lane + 32];
                __syncwarp();
            }
            else if (tl) ok = run_box<NCMAX, true, CLAMP>(c, pr, g, top_init, ws, box);
            else ok = run_box<NCMAX, GBD, false>(c, pr, g, in_init, ws, box);
            if (!ok) { if (lane == 0) atomicExch(c.error_flag, 3); break; }
            const int al = job.al, bl = job.bl;
            int i = tl ? top.max_i : job.R, j = tl ? top.max_j : job.C;
            const DevParams &PW = (tl && mod) ? c.params_tl : c.params;                      // scoring of this box's walk
            const DevEntry *entw = (tl && mod) ? c.ent_tl + pr.am_base : ent;                // D view of this box's walk
            // ---- walk (trace_in_arcmatch / trace_noex, state E_NO_NO)
            for (;;) {
                const int mij = box_get(box, g, i - al, j - bl);
                if (tl && P.sequ_local && mij == 0) { if (lane == 0) c.top[t].min_ij = (i & 0xffff) | (j << 16); break; }   // aligner.cc:1251-1255
                if (i <= al) {                                                               // :1257-1271
                    if (!(tl && (P.sequ_local || P.fe_left1))) for (int k = bl + 1 + lane; k <= j; k += 32) edges[al + k] = (al << 2) | LB_EDGE_INS;
                    break;
                }
                if (j <= bl) {                                                               // :1273-1287
                    if (!(tl && (P.sequ_local || P.fe_left2))) for (int k = al + 1 + lane; k <= i; k += 32) edges[k + bl] = (k << 2) | LB_EDGE_DEL;
                    break;
                }
                const bool vdiag = valid(i - 1, j - 1);
                const int sg_ij = psig != nullptr ? __ldg(psig + (long long)i * (m + 1) + j) : PW.sigma8[ca[i] * LB_NCODES + cb[j]];
                if (vdiag && mij == box_get(box, g, i - 1 - al, j - 1 - bl) + sg_ij) {   // :1099-1105
                    emit(i, j, LB_EDGE_MATCH);
                    i--; j--;
                    continue;
                }
                bool moved = false;
                if (PW.open == 0) {                                                           // :1107-1124 linear gap cost
                    if (valid(i - 1, j) && mij == box_get(box, g, i - 1 - al, j - bl) + PW.gap) { emit(i, j, LB_EDGE_DEL); i--; moved = true; }
                    else if (valid(i, j - 1) && mij == box_get(box, g, i - al, j - 1 - bl) + PW.gap) { emit(i, j, LB_EDGE_INS); j--; moved = true; }
                } else {                                                                     // :1125-1172 affine: gap runs
                    int cost = PW.open;
                    for (int k = 1; i >= al + k; k++) {
                        if (!valid(i - k, j)) break;
                        cost += PW.gap;
                        if (mij == box_get(box, g, i - k - al, j - bl) + cost) {
                            for (int l = lane; l < k; l += 32) edges[(i - l) + j] = ((i - l) << 2) | LB_EDGE_DEL;
                            i -= k; moved = true;
                            break;
                        }
                    }
                    if (!moved) {
                        cost = PW.open;
                        for (int k = 1; j >= bl + k; k++) {
                            if (!valid(i, j - k)) break;
                            cost += PW.gap;
                            if (mij == box_get(box, g, i - al, j - k - bl) + cost) {
                                for (int l = lane; l < k; l += 32) edges[i + (j - l)] = (i << 2) | LB_EDGE_INS;
                                j -= k; moved = true;
                                break;
                            }
                        }
                    }
                }
                if (moved) continue;
                if (!vdiag) break;                                                           // :1176-1178
                // arc matches with right ends (i, j), in common_right_end_list order (:1186-1228)
                const int e0 = sptr[i + j], e1 = sptr[i + j + 1];
                // the first hit in common_right_end_list order = the hit with the largest (al, bl) (arc_matches.hh:188-220);
                // the S-order inside an anti-diagonal is by source anti-diagonal, so all candidates are inspected
                int found = -1, fkey = -1;
                for (int base = e0; base < e1; base += 32) {
                    const int e = base + lane;
                    int key = -1;
                    if (e < e1) {
                        const DevEntry en = entw[e];
                        const int p = LB_ENT_LO(en.x), q = LB_ENT_HI(en.x);
                        if (LB_ENT_LO(en.y) == i && p >= al && q >= bl && mij == box_get(box, g, p - al, q - bl) + en.d) key = (p << 16) | q;
                    }
                    const int best = __reduce_max_sync(0xffffffffu, key);
                    if (best > fkey) { fkey = best; found = base + __ffs(__ballot_sync(0xffffffffu, key == best)) - 1; }
                }
                if (found < 0) { if (lane == 0) atomicExch(c.error_flag, 4); break; }
                const DevEntry en = ent[found];
                const int xl = LB_ENT_LO(en.x) + 1, yl = LB_ENT_HI(en.x) + 1;  // left ends of the arc match
                if (lane == 0) { strA[xl] = '('; strA[i] = ')'; strB[yl] = '('; strB[j] = ')'; }
                emit(xl, yl, LB_EDGE_MATCH);
                emit(i, j, LB_EDGE_MATCH);
                if (!nolp) {                                                                 // trace_arcmatch :967-1028
                    int cur = (int)lpos[found];
                    DevArcMatch x = am[cur];
                    // stacked arc matches (:984-1003): while D is explained by the inner arc match, descend without opening a box
                    while (P.stacking && x.inner >= 0 && x.score_st != LB_NOSTACK && ent[x.spos].d == ent[am[x.inner].spos].d + x.score_st &&
                           ent[am[x.inner].spos].d >= LB_NEG_LIMIT) {
                        cur = x.inner; x = am[cur];
                        const int ial = x.ends_a & 0xfff, iar = x.ends_a >> 12, ibl = x.ends_b & 0xfff, ibr = x.ends_b >> 12;
                        if (lane == 0) { strA[ial] = '('; strA[iar] = ')'; strB[ibl] = '('; strB[ibr] = ')'; }
                        emit(ial, ibl, LB_EDGE_MATCH);
                        emit(iar, ibr, LB_EDGE_MATCH);
                    }
                    const int bal = x.ends_a & 0xfff, bar = x.ends_a >> 12, bbl = x.ends_b & 0xfff, bbr = x.ends_b >> 12;
                    if (lane == 0) { TraceJob nj; nj.al = (short)bal; nj.bl = (short)bbl; nj.R = (short)(bar - 1); nj.C = (short)(bbr - 1); nj.am = -1; stack[sp] = nj; }
                    sp++;
                } else {                                                                     // trace_arcmatch_noLP :1030-1079
                    int cur = (int)lpos[found];
                    for (;;) {
                        const DevArcMatch x = am[cur];
                        const DevArcMatch in = am[x.inner];
                        const int ial = in.ends_a & 0xfff, iar = in.ends_a >> 12, ibl = in.ends_b & 0xfff, ibr = in.ends_b >> 12;
                        if (lane == 0) { strA[ial] = '('; strA[iar] = ')'; strB[ibl] = '('; strB[ibr] = ')'; }
                        emit(ial, ibl, LB_EDGE_MATCH);
                        emit(iar, ibr, LB_EDGE_MATCH);
                        if (ent[x.spos].d == ent[in.spos].d + (P.stacking ? x.score_st : x.score)) { cur = x.inner; continue; }
                        if (lane == 0) { TraceJob nj; nj.al = (short)ial; nj.bl = (short)ibl; nj.R = (short)(iar - 1); nj.C = (short)(ibr - 1); nj.am = -1; stack[sp] = nj; }
                        sp++;
                        break;
                    }
                }
                i = xl - 1; j = yl - 1;
            }
            __syncwarp();
            tl = false;
            have = sp > 0;
            if (have) { sp--; job = stack[sp]; }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Traceback for --struct-local (aligner.cc:1019-1026, :1065-1076, :1289-1342): the top level box has the single state
// E_NO_NO; every arc-match box is refilled with all eight matrices and walked with the reference's state machine.
template <bool CLAMP>
__global__ void __launch_bounds__(32) trace_sl_kernel(DevCtx c, int pair_begin, int pair_end, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    WarpSmem ws;
    carve(c, smem, ws);
    const int box_words = c.scratch_words / 8;
    int *boxes = c.scratch + (size_t)blockIdx.x * c.scratch_words;
    const DevParams &P = c.params;
    const bool nolp = P.no_lonely_pairs != 0;
    const int ex = P.exclusion;
    BoxInit top_init;
    const bool globalA = !(P.sequ_local || P.fe_left2), globalB = !(P.sequ_local || P.fe_left1);
    top_init.col_base = globalA ? P.open : 0; top_init.col_step = globalA ? P.gap : 0;
    top_init.row_base = globalB ? P.open : 0; top_init.row_step = globalB ? P.gap : 0;
    for (;;) {
        int t = 0;
        if (lane == 0) t = pair_begin + atomicAdd(cursor, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= pair_end) break;
        const DevPair pr = c.pairs[t];
        const int n = pr.lenA, m = pr.lenB;
        const int *lo = c.band_lo + pr.band, *hi = c.band_hi + pr.band;
        const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
        const DevEntry *ent = c.ent + pr.am_base;
        const DevArcMatch *am = c.am + pr.am_base;
        const unsigned *lpos = c.lpos + pr.am_base;
        const int *sptr = c.sptr + pr.sptr;
        int *edges = c.trace_edges + pr.sptr;
        char *strA = c.trace_str + pr.sptr, *strB = strA + n + 1;
        TraceJob *stack = c.trace_stack + (size_t)t * c.trace_stack_cap;
        for (int k = lane; k < n + m + 3; k += 32) edges[k] = 0;
        for (int k = lane; k <= n; k += 32) strA[k] = '.';
        for (int k = lane; k <= m; k += 32) strB[k] = '.';
        __syncwarp();
        auto valid = [&](int i, int j) { return i >= 0 && j >= 0 && lo[i] <= j && j <= hi[i]; };
        auto emit = [&](int i, int j, int kind) { if (lane == 0) edges[i + j] = (i << 2) | kind; };
        int sp = 0;
        const DevTopResult top = c.top[t];
        bool tl = true;
        TraceJob job;
        job.al = (short)(c.r_on ? c.r_sa - 1 : 0); job.bl = (short)(c.r_on ? c.r_sb - 1 : 0);
        job.R = (short)(c.r_on ? c.r_ea : n); job.C = (short)(c.r_on ? c.r_eb : m); job.am = -1;
        bool have = true;
        while (have) {
            BoxGeom g;
            if (tl) setup_box2(c, pr, job.al, job.bl, job.R, job.C, g, ws);
            else setup_box(c, pr, job.al, job.bl, job.R, job.C, g, ws);
            bool ok = (g.umax + 1) * g.nslots <= box_words;
            if (ok) ok = tl ? run_box<4, true, CLAMP>(c, pr, g, top_init, ws, boxes) : run_box_sl<true>(c, pr, g, ws, boxes, box_words);
            if (!ok) { if (lane == 0) atomicExch(c.error_flag, 3); break; }
            const int al = job.al, bl = job.bl;
            auto B = [&](int st, int i, int j) { return box_get(boxes + st * box_words, g, i - al, j - bl); };
            // M_st(i, j) or -inf if the cell is outside the band (such cells hold -inf in the reference, aligner.cc:322-367)
            auto Bv = [&](int st, int i, int j) { return valid(i, j) ? B(st, i, j) : LB_NEG; };
            int i = tl ? top.max_i : job.R, j = tl ? top.max_j : job.C;
            int st = 0;
            if (!tl) {
                // state in which the arc match closes: first closed state that explains D (aligner.cc:1019-1025, :1065-1075)
                const DevArcMatch x = am[job.am];
                const int add = nolp ? (P.stacking ? x.score_st : x.score) + am[x.inner].score : x.score;
                st = -1;
                for (int k = 0; k < 4; k++) if (ent[x.spos].d == B(k, i, j) + add) { st = k; break; }
                if (st < 0) { tl = false; have = sp > 0; if (have) { sp--; job = stack[sp]; } continue; }
            }
            for (;;) {
                const int mij = B(st, i, j);
                if (tl && P.sequ_local && mij == 0) break;
                if (i <= al) {
                    if (st == 0 || st == 4 || st == 1)
                        if (!(tl && (P.sequ_local || P.fe_left1))) for (int k = bl + 1 + lane; k <= j; k += 32) edges[al + k] = (al << 2) | LB_EDGE_INS;
                    break;
                }
                if (j <= bl) {
                    if (st == 0 || st == 5 || st == 2)
                        if (!(tl && (P.sequ_local || P.fe_left2))) for (int k = al + 1 + lane; k <= i; k += 32) edges[k + bl] = (k << 2) | LB_EDGE_DEL;
                    break;
                }
                // ---- open states and exclusion transitions (aligner.cc:1289-1342)
                bool noex = false, dead = false;
                switch (st) {
                    case 0: noex = true; break;
                    case 4: if (mij == Bv(4, i - 1, j)) i--; else if (mij == B(0, i, j)) st = 0; else dead = true; break;
                    case 5: if (mij == Bv(5, i, j - 1)) j--; else if (mij == B(0, i, j)) st = 0; else dead = true; break;
                    case 2: if (mij == B(5, i, j) + ex) st = 5; else noex = true; break;
                    case 6: if (mij == Bv(6, i - 1, j)) i--; else if (mij == B(2, i, j)) st = 2; else dead = true; break;
                    case 1: if (mij == B(4, i, j) + ex) st = 4; else noex = true; break;
                    case 7: if (mij == Bv(7, i, j - 1)) j--; else if (mij == B(1, i, j)) st = 1; else dead = true; break;
                    default: if (mij == B(6, i, j) + ex) st = 6; else if (mij == B(7, i, j) + ex) st = 7; else noex = true; break;
                }
                if (dead) break;
                if (!noex) continue;
                // ---- trace_noex in closed state st (aligner.cc:1082-1229)
                const bool vdiag = valid(i - 1, j - 1);
                if (vdiag && mij == B(st, i - 1, j - 1) + P.sigma8[ca[i] * LB_NCODES + cb[j]]) { emit(i, j, LB_EDGE_MATCH); i--; j--; continue; }
                bool moved = false;
                if (P.open == 0) {
                    if (valid(i - 1, j) && mij == B(st, i - 1, j) + P.gap) { emit(i, j, LB_EDGE_DEL); i--; moved = true; }
                    else if (valid(i, j - 1) && mij == B(st, i, j - 1) + P.gap) { emit(i, j, LB_EDGE_INS); j--; moved = true; }
                } else {
                    int cost = P.open;
                    for (int k = 1; i >= al + k; k++) {
                        if (!valid(i - k, j)) break;
                        cost += P.gap;
                        if (mij == B(st, i - k, j) + cost) {
                            for (int l = lane; l < k; l += 32) edges[(i - l) + j] = ((i - l) << 2) | LB_EDGE_DEL;
                            i -= k; moved = true;
                            break;
                        }
                    }
                    if (!moved) {
                        cost = P.open;
                        for (int k = 1; j >= bl + k; k++) {
                            if (!valid(i, j - k)) break;
                            cost += P.gap;
                            if (mij == B(st, i, j - k) + cost) {
                                for (int l = lane; l < k; l += 32) edges[i + (j - l)] = (i << 2) | LB_EDGE_INS;
                                j -= k; moved = true;
                                break;
                            }
                        }
                    }
                }
                if (moved) continue;
                if (!vdiag) break;
                const int e0 = sptr[i + j], e1 = sptr[i + j + 1];
                int found = -1, fkey = -1;   // hit with the largest (al, bl), see trace_kernel
                for (int base = e0; base < e1; base += 32) {
                    const int e = base + lane;
                    int key = -1;
                    if (e < e1) {
                        const DevEntry en = ent[e];
                        const int p = LB_ENT_LO(en.x), q = LB_ENT_HI(en.x);
                        if (LB_ENT_LO(en.y) == i && p >= al && q >= bl && mij == B(st, p, q) + en.d) key = (p << 16) | q;
                    }
                    const int best = __reduce_max_sync(0xffffffffu, key);
                    if (best > fkey) { fkey = best; found = base + __ffs(__ballot_sync(0xffffffffu, key == best)) - 1; }
                }
                if (found < 0) { if (lane == 0) atomicExch(c.error_flag, 4); break; }
                const DevEntry en = ent[found];
                const int xl = LB_ENT_LO(en.x) + 1, yl = LB_ENT_HI(en.x) + 1;
                if (lane == 0) { strA[xl] = '('; strA[i] = ')'; strB[yl] = '('; strB[j] = ')'; }
                emit(xl, yl, LB_EDGE_MATCH);
                emit(i, j, LB_EDGE_MATCH);
                int cur = (int)lpos[found];
                if (!nolp) {
                    DevArcMatch x = am[cur];
                    while (P.stacking && x.inner >= 0 && x.score_st != LB_NOSTACK && ent[x.spos].d == ent[am[x.inner].spos].d + x.score_st &&
                           ent[am[x.inner].spos].d >= LB_NEG_LIMIT) {
                        cur = x.inner; x = am[cur];
                        const int ial = x.ends_a & 0xfff, iar = x.ends_a >> 12, ibl = x.ends_b & 0xfff, ibr = x.ends_b >> 12;
                        if (lane == 0) { strA[ial] = '('; strA[iar] = ')'; strB[ibl] = '('; strB[ibr] = ')'; }
                        emit(ial, ibl, LB_EDGE_MATCH);
                        emit(iar, ibr, LB_EDGE_MATCH);
                    }
                    const int bal = x.ends_a & 0xfff, bar = x.ends_a >> 12, bbl = x.ends_b & 0xfff, bbr = x.ends_b >> 12;
                    if (lane == 0) { TraceJob nj; nj.al = (short)bal; nj.bl = (short)bbl; nj.R = (short)(bar - 1); nj.C = (short)(bbr - 1); nj.am = cur; stack[sp] = nj; }
                    sp++;
                } else {
                    for (;;) {
                        const DevArcMatch x = am[cur];
                        const DevArcMatch in = am[x.inner];
                        const int ial = in.ends_a & 0xfff, iar = in.ends_a >> 12, ibl = in.ends_b & 0xfff, ibr = in.ends_b >> 12;
                        if (lane == 0) { strA[ial] = '('; strA[iar] = ')'; strB[ibl] = '('; strB[ibr] = ')'; }
                        emit(ial, ibl, LB_EDGE_MATCH);
                        emit(iar, ibr, LB_EDGE_MATCH);
                        if (ent[x.spos].d == ent[in.spos].d + (P.stacking ? x.score_st : x.score)) { cur = x.inner; continue; }
                        if (lane == 0) { TraceJob nj; nj.al = (short)ial; nj.bl = (short)ibl; nj.R = (short)(iar - 1); nj.C = (short)(ibr - 1); nj.am = cur; stack[sp] = nj; }
                        sp++;
                        break;
                    }
                }
                i = xl - 1; j = yl - 1;
            }
            __syncwarp();
            tl = false;
            have = sp > 0;
            if (have) { sp--; job = stack[sp]; }
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

void launch_dfill(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int q, cudaStream_t st) {
#define CALL(N)                                                                                      \
    if (generic_borders) dfill_kernel<N, true><<<grid, 32, smem_bytes, st>>>(c, q); \
    else dfill_kernel<N, false><<<grid, 32, smem_bytes, st>>>(c, q)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
void launch_dfill_dep(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int *cursor, cudaStream_t st) {
#define CALL(N)                                                                                      \
    if (generic_borders) dfill_dep_kernel<N, true><<<grid, 32, smem_bytes, st>>>(c, cursor); \
    else dfill_dep_kernel<N, false><<<grid, 32, smem_bytes, st>>>(c, cursor)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
void launch_dfill_sl(const DevCtx &c, int grid, int smem_bytes, int q, cudaStream_t st) { dfill_sl_kernel<<<grid, 32, smem_bytes, st>>>(c, q); }
void launch_trace_sl(const DevCtx &c, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st) {
    if (c.params.sequ_local) trace_sl_kernel<true><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor);
    else trace_sl_kernel<false><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor);
}
cudaError_t configure_sl(int smem_bytes, int *ctas_per_sm) {
    const int cap = lb200_sticky_smem(200, smem_bytes);
    cudaError_t e = cudaFuncSetAttribute(dfill_sl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(trace_sl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(trace_sl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, dfill_sl_kernel, 32, smem_bytes);
    return e;
}
void launch_toplevel(const DevCtx &c, int ncmax, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st) {
    const bool clamp = c.params.sequ_local != 0;
#define CALL(N)                                                                                      \
    if (clamp) toplevel_kernel<N, true><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor); \
    else toplevel_kernel<N, false><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
void launch_trace(const DevCtx &c, int ncmax, bool generic_borders, int grid, int smem_bytes, int pair_begin, int pair_end, int *cursor, cudaStream_t st) {
    const bool clamp = c.params.sequ_local != 0;
#define CALL(N)                                                                                                  \
    if (clamp) { if (generic_borders) trace_kernel<N, true, true><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor); \
                 else trace_kernel<N, false, true><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor); }              \
    else { if (generic_borders) trace_kernel<N, true, false><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor);      \
           else trace_kernel<N, false, false><<<grid, 32, smem_bytes, st>>>(c, pair_begin, pair_end, cursor); }
    LB_DISPATCH(ncmax, CALL);
#undef CALL
}
cudaError_t configure_kernels(int ncmax, int smem_bytes, int *dfill_ctas_per_sm) {
    cudaError_t e = cudaSuccess;
    const int cap = lb200_sticky_smem(ncmax, smem_bytes);
#define SET(K) if (e == cudaSuccess) e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, cap)
#define CALL(N)                                                                                                         \
    SET((dfill_kernel<N, true>)); SET((dfill_kernel<N, false>)); SET((dfill_dep_kernel<N, true>)); SET((dfill_dep_kernel<N, false>)); SET((toplevel_kernel<N, true>)); SET((toplevel_kernel<N, false>)); \
    SET((trace_kernel<N, true, true>)); SET((trace_kernel<N, false, true>)); SET((trace_kernel<N, true, false>)); SET((trace_kernel<N, false, false>)); \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(dfill_ctas_per_sm, dfill_kernel<N, false>, 32, smem_bytes)
    LB_DISPATCH(ncmax, CALL);
#undef CALL
#undef SET
    return e;
}

}  // namespace lb200

#include "pf_inside.cuh"
#include "pf_outside.cuh"
