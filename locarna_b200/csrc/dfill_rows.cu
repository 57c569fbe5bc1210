// Row-grouped D fill: the dominant kernel of the batched path (sequences <= LB_PACK_MAXLEN, single-state boxes).
//
// Reference path (file:line relative to /root/reference/src/LocARNA): align_D aligner.cc:660-732 visits the left-end pairs (al, bl)
// one by one; for each it fills the M box from (al, bl) to the largest right ends (align_in_arcmatch :373-565, align_noex :153-233,
// init_state :264-368) and reads D(a, b) of every arc match with these left ends from the box (fill_D_entries[_noLP] :574-657).
//
// Here the tasks of one origin ROW al of a pair are handled together. Up to LB_GV of them form a group = one warp's work item: one
// anti-diagonal sweep computes LB_GV "layers" per cell, layer v being the box with origin (al, bl_v). A layer is -inf left of its
// origin column, so the recurrence itself keeps the layers apart; its origin is seeded with 0 and its first row / column then fall
// out of the E / F recurrences exactly as init_state writes them (all-global box, indel_opening <= 0). What the layers SHARE is
// everything that is not the nine integer operations of align_noex: the band test, the sigma lookup, the box addressing and the
// arc-match entries (one gather address and one 16-byte gather per entry for all layers).
// All groups of a row sweep the same geometry (the union box of the row), so they also share the arc-match entry LIST: the groups
// of a row first filter, each for its share of the target anti-diagonals, the pair's S-order entries down to those whose source and
// target lie in the row's box (with the box offset of the source, the accumulator slot of the target and D resolved), then every
// group streams that list: nothing is filtered, unpacked or addressed inside the sweep.
//
// Sweep geometry: lane l owns the columns j with (j / NC) mod 32 == l of the box (NC = 1 or 2 columns per lane). On anti-diagonal
// u = i + j it computes cell (u - j, j): the up neighbour is its own previous cell, the left / diagonal neighbours are the previous
// lane's last two cells (two SHFL per layer and step). A monotone band makes the active columns of an anti-diagonal a window that
// only moves right, so a lane whose column has left the band moves on to column j + 32 NC; the state it carries over is -inf because
// every out-of-band cell is forced to -inf. A box needs NC = 2 when more than ~30 of its columns are active at once.
// Arc-match terms: the entries of four target anti-diagonals 4g..4g+3 form one padded run of the list; it is folded after cell step
// 4g - 2 (all sources lie >= 8 anti-diagonals before their target, so they are final) into a ring of six per-anti-diagonal
// accumulators (shared memory, atomicMax) which the cell step consumes and resets.
//
// Schedule: groups are claimed in the order row al descending, the groups of a row consecutively. A group of row al reads D only of
// arc matches with left ends in rows > al of its own pair, so it waits until all groups of those rows are complete (per-pair
// counter); a second per-pair counter tells when all groups of its own row have contributed their share of the entry list. Every
// group a claimant waits for precedes it in the claim order or belongs to its own row, whose remaining groups are claimed next by
// the warps that run free: no deadlock, no co-residency requirement beyond one row's groups.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rows.h"

namespace lb200 {
namespace {

constexpr int V = LB_GV;
constexpr int RING = LB_ROWS_RING;   // accumulator rows: the anti-diagonal being consumed + the five a fold may target (see sweep)
constexpr int BIAS = 640;   // row indices are biased so that the bit-field band test never sees a negative row
constexpr uint32_t G_LO = 0x800u, G_HI = 0x800000u, GUARDS = G_LO | G_HI;
// column word: bits 0..10 first valid row + BIAS, bits 12..22 2047 - (last valid row + BIAS), bits 24..31 4 * symbol code of B
constexpr uint32_t COLW_NEVER = 2047u;                    // never valid, never "below the band"
constexpr uint32_t COLW_EMPTY = 2047u | (2047u << 12);    // never valid, below at once

static_assert(V == 4, "the vector loads / stores of the layers assume four layers per cell");

__device__ __forceinline__ int addmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }   // max(a + b, c)
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
// band-test constant of biased row ib: z = key - colw keeps both guard bits iff first <= ib <= last
__device__ __forceinline__ uint32_t band_key(int ib) { return ((uint32_t)ib | 0x800u) | ((((uint32_t)(2047 - ib)) | 0x800u) << 12); }

struct Smem {
    const int *sig;       // 8 x 8 base match scores
    int *acc;             // V x RING x (32 NC) accumulators, layer-major: the lanes of a fold hit different banks
    int *gstart;          // first block of every target group in the row's entry list (LB_ROWS_TG entries)
    uint32_t *colw;       // column words of the row's box
    uint8_t *rowcode;     // rowcode[row_pad + ip] = 32 * symbol code of A[al + ip]
};

struct Geom {
    int al, bl0, Rn, Cn;  // the row's union box: origin (al, bl0), local extent
    int umax;             // last anti-diagonal of the row's box
    int u0, u1;           // anti-diagonals this group sweeps: its first origin .. the last cell of its own tasks
    int useed;            // last anti-diagonal that holds an origin of this group
    int n_tg;             // target groups of the row's list
    int jv[V];            // local origin column of layer v (-1: layer not used)
};

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ int a_first(const Smem &sm, int j) { return (int)(sm.colw[j] & 0x7ffu) - BIAS + j; }
__device__ __forceinline__ int a_last(const Smem &sm, int j) { return 2047 - (int)((sm.colw[j] >> 12) & 0x7ffu) - BIAS + j; }

// Stage the column / row tables of the row's box and the block offsets of its entry list; decide how many columns per lane the box
// needs (0: not supported by this kernel).
__device__ int setup_group(const DevCtx &c, const RowsCtx &r, const DevGroup &grp, const DevPair &pr, Geom &g, const Smem &sm) {
    const int lane = threadIdx.x & 31;
    g.al = grp.al; g.bl0 = grp.bl0r; g.Rn = grp.Rr - grp.al; g.Cn = grp.Cr - grp.bl0r;
    int u0 = 1 << 20, useed = 0, own = 0;
#pragma unroll
    for (int v = 0; v < V; v++) {
        g.jv[v] = -1;
        if (v < grp.nmem) {
            const DevTask tk = c.tasks[grp.task[v]];
            g.jv[v] = (int)tk.bl - g.bl0;
            u0 = min(u0, g.jv[v]); useed = max(useed, g.jv[v]);
            own = max(own, ((int)tk.R - g.al) + ((int)tk.C - g.bl0));
        }
    }
    for (int k = lane; k < r.colw_words; k += 32) sm.colw[k] = COLW_NEVER;
    for (int k = lane; k < (r.rowcode_bytes >> 2); k += 32) ((uint32_t *)sm.rowcode)[k] = 0;
    __syncwarp();
    const int *cf = r.col_first + pr.sptr + g.bl0, *cl = r.col_last + pr.sptr + g.bl0;
    const uint8_t *ca = c.codes + pr.codesA, *cb = c.codes + pr.codesB;
    int umax = 0;
    bool supported = true;
    for (int j = lane; j <= g.Cn; j += 32) {
        const int first = max(cf[j], g.al), last = min(cl[j], (int)grp.Rr);
        const uint32_t code = (uint32_t)((g.bl0 + j >= 1) ? cb[g.bl0 + j] : 0) << 26;
        uint32_t w = COLW_EMPTY;
        if (last >= first) {
            w = (uint32_t)(first - g.al + BIAS) | ((uint32_t)(2047 - (last - g.al + BIAS)) << 12);
            umax = max(umax, last - g.al + j);
        } else supported = false;   // a gap in the band: left to the box-by-box kernels
        sm.colw[j] = w | code;
    }
    for (int ip = lane; ip <= g.Rn; ip += 32) sm.rowcode[r.row_pad + ip] = (uint8_t)(((g.al + ip >= 1) ? ca[g.al + ip] : 0) * 32);
    g.umax = warp_max(umax);
    g.u0 = u0; g.useed = useed; g.u1 = min(g.umax, own);
    __syncwarp();
    // a lane moves from column set s to s + 1 when its last column has left the band; this must happen before the first column of
    // the next set enters it (one idle step in between resets the carried state), and the previous lane must not have entered the
    // next set's band while this lane still reads it as its left neighbour
    bool ok1 = true, ok2 = true;
    for (int j = lane; j <= g.Cn; j += 32) {
        if (j >= 32 && a_first(sm, j) < a_last(sm, j - 32) + 2) ok1 = false;
        if (j >= 31 && a_first(sm, j) < a_last(sm, j - 31)) ok1 = false;
        if ((j & 1) == 0 && j >= 64 && a_first(sm, j) < a_last(sm, j - 63) + 2) ok2 = false;
        if ((j & 1) == 1 && j >= 63 && a_first(sm, j) < a_last(sm, j - 63)) ok2 = false;
    }
    supported = __all_sync(0xffffffffu, supported);
    ok1 = __all_sync(0xffffffffu, ok1);
    ok2 = __all_sync(0xffffffffu, ok2);
    // block offsets of the target groups: group t holds the entries of the local target anti-diagonals 4t .. 4t+3, i.e. at most the
    // S-order entries [q[4t], q[4t+4]), padded to whole blocks of 32
    const int s0 = g.al + g.bl0;
    const int *q = c.sptr + pr.sptr + s0;
    const int q_cap = pr.lenA + pr.lenB + 2 - s0;
    g.n_tg = (g.umax + 4) >> 2;
    int carry = 0;
    for (int base = 0; base < g.n_tg; base += 32) {
        const int t = base + lane;
        int blocks = 0;
        if (t < g.n_tg && t >= 2) blocks = (__ldg(q + min(4 * t + 4, q_cap)) - __ldg(q + min(4 * t, q_cap)) + 31) >> 5;   // targets < 8: no source inside the box
        int x = blocks;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (t < LB_ROWS_TG) sm.gstart[t] = carry + x - blocks;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    __syncwarp();
    if (!supported || g.Cn + 72 > r.colw_words || g.umax + BIAS > 2040 || g.n_tg + 1 > LB_ROWS_TG || (long long)carry * 32 > r.clist_cap) return 0;
    if (r.force_nc == 2 && ok2) return 2;
    return ok1 ? 1 : (ok2 ? 2 : 0);
}

// This group's share of the row's entry list: target groups t = gi, gi + G, ... (after all D values the row reads are final).
// Entry = (box offset of the source cell in int4 units | accumulator slot of the target << 16, D); entries whose D is -inf are dropped.
template <int NC>
__device__ void build_list(const DevCtx &c, const RowsCtx &r, const DevGroup &grp, const DevPair &pr, const Geom &g, const Smem &sm) {
    const int lane = threadIdx.x & 31;
    constexpr int LOGW = NC == 1 ? 5 : 6, W = 32 * NC;
    const int s0 = g.al + g.bl0;
    const int *q = c.sptr + pr.sptr + s0;
    const int q_cap = pr.lenA + pr.lenB + 2 - s0;
    const uint2 *ent8 = c.ent8 + pr.am_base;   // packed S-order (LB_PACK_*): 8 bytes per entry, D kept current by write_d
    uint2 *list = r.clist + (size_t)grp.pair * r.clist_cap;
    int *nblk = r.cnblk + (size_t)grp.pair * LB_ROWS_TG;
    const int a_lo = g.al, b_lo = g.bl0, a_hi = g.al + g.Rn, b_hi = g.bl0 + g.Cn;
    for (int t = 2 + grp.gi; t < g.n_tg; t += grp.G) {
        uint2 *out = list + (size_t)sm.gstart[t] * 32;
        int n = 0;
        const int qa = __ldg(q + min(4 * t, q_cap)), qb = __ldg(q + min(4 * t + 4, q_cap));
#pragma unroll 1
        for (int e0 = qa; e0 < qb; e0 += 128) {   // four independent loads per lane in flight
            uint2 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int e = e0 + 32 * k + lane;
                v[k] = make_uint2(0u, 0u);
                if (e < qb) v[k] = __ldcg(ent8 + e);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (e0 + 32 * k >= qb) break;
                const int e = e0 + 32 * k + lane;
                const int p1 = (int)(v[k].x & 511u), q1 = (int)((v[k].x >> 9) & 511u);          // al' - 1, bl' - 1
                const int ar = (int)((v[k].x >> 18) & 511u), br = (int)((v[k].x >> 27) | ((v[k].y & 15u) << 5));
                const int d = (int)v[k].y >> 4;
                const bool in = e < qb && p1 >= a_lo && q1 >= b_lo && ar <= a_hi && br <= b_hi && d != LB_PACK_NEG;
                const uint32_t p = (uint32_t)(p1 - a_lo), qq = (uint32_t)(q1 - b_lo);
                const uint32_t ta = (uint32_t)(ar + br);                                        // absolute target anti-diagonal
                const uint32_t tcol = (uint32_t)(br - b_lo);
                uint2 o;
                o.x = ((p + qq) << LOGW) | (qq & (W - 1)) | ((((ta % RING) << LOGW) | (tcol & (W - 1))) << 16);
                o.y = (uint32_t)d;
                const unsigned mask = __ballot_sync(0xffffffffu, in);
                if (in) out[n + __popc(mask & ((1u << lane) - 1u))] = o;
                n += __popc(mask);
            }
        }
        const int blocks = (n + 31) >> 5;
        if (n + lane < blocks * 32) out[n + lane] = make_uint2(0u, (uint32_t)LB_NEG);   // padding: folds -inf
        if (lane == 0) nblk[t] = blocks;
    }
}

template <int NC>
struct State {
    int m[NC][V], e[NC][V], f[NC][V], md[NC][V];   // last cell of every owned column: M, E, F, and the left neighbour's M one step back
    uint32_t w[NC];          // column words
    uint32_t K;              // band-test constant of slot 0's row (slot k: K + k * 4095)
    const uint32_t *cwp;     // &colw[first owned column]
    const uint8_t *rcp;      // &rowcode[row of slot 0] (slot k: rcp[-k])
    int *bp;                 // &box[u * 32 NC V + lane * NC * V]
    int jc;                  // first owned column
};

// One anti-diagonal step. The origins of the layers (cell (0, jv) on anti-diagonal jv) are seeded through the accumulators, see sweep().
template <int NC>
__device__ __forceinline__ void dp_step(State<NC> &S, const Geom &g, const int *sig, int *arcp, int gap, int gap_open, int left, int u) {
    constexpr int AW = 32 * NC * V;
    int mL0[V], fL0[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        mL0[v] = __shfl_sync(0xffffffffu, S.m[NC - 1][v], left);
        fL0[v] = __shfl_sync(0xffffffffu, S.f[NC - 1][v], left);
    }
    bool below = false;
    // the accumulators of the lane's NC cells are adjacent words: one access per layer (NC = 2: 64-bit, no bank conflict between lanes)
    int arc[NC][V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        if (NC == 2) {
            int2 *a2 = (int2 *)(arcp + v * (RING * 32 * NC));
            const int2 a = *a2;
            *a2 = make_int2(LB_NEG, LB_NEG);
            arc[0][v] = a.x; arc[NC - 1][v] = a.y;
        } else {
#pragma unroll
            for (int k = 0; k < NC; k++) { arc[k][v] = arcp[v * (RING * 32 * NC) + k]; arcp[v * (RING * 32 * NC) + k] = LB_NEG; }
        }
    }
#pragma unroll
    for (int k = NC - 1; k >= 0; k--) {   // right to left: slot k reads slot k - 1's previous cell
        const uint32_t z = S.K + (uint32_t)k * 4095u - S.w[k];
        const bool ok = (~z & GUARDS) == 0;
        if (k == NC - 1) below = (z & G_HI) == 0;
        const int sg = *(const int *)((const char *)sig + S.rcp[-k] + (S.w[k] >> 24));
        int nm[V];
#pragma unroll
        for (int v = 0; v < V; v++) {
            const int ml = k == 0 ? mL0[v] : S.m[k > 0 ? k - 1 : 0][v];
            const int fl = k == 0 ? fL0[v] : S.f[k > 0 ? k - 1 : 0][v];
            const int ee = addmax(S.e[k][v], gap, S.m[k][v] + gap_open);
            const int ff = addmax(fl, gap, ml + gap_open);
            int mm = max3(addmax(S.md[k][v], sg, ee), ff, arc[k][v]);
            nm[v] = ok ? mm : LB_NEG;
            S.m[k][v] = nm[v]; S.e[k][v] = ok ? ee : LB_NEG; S.f[k][v] = ok ? ff : LB_NEG;
            S.md[k][v] = ml;
        }
        if (ok) *(int4 *)(S.bp + k * V) = make_int4(nm[0], nm[1], nm[2], nm[3]);   // only band cells are ever read back
    }
    S.bp += AW;
    S.K -= 4095u;      // next row
    S.rcp += 1;
    if (below) {       // the lane's columns have left the band: on to the next column set (32 NC columns to the right, 32 NC rows up)
        S.cwp += 32 * NC;
#pragma unroll
        for (int k = 0; k < NC; k++) S.w[k] = S.cwp[k];
        S.K += (uint32_t)(32 * NC) * 4095u;
        S.rcp -= 32 * NC;
        S.jc += 32 * NC;
    }
}

template <int NC>
__device__ void sweep(const DevCtx &c, const RowsCtx &r, const DevGroup &grp, const Geom &g, const Smem &sm, int *box) {
    const int lane = threadIdx.x & 31;
    constexpr int LOGW = NC == 1 ? 5 : 6, W = 32 * NC, AW = W * V;
    const int gap = c.params.gap, gap_open = c.params.gap_open;
    for (int k = lane; k < RING * AW; k += 32) sm.acc[k] = LB_NEG;
    const int s0 = g.al + g.bl0;

    // every lane starts on the column of its residue class that is current at anti-diagonal u0
    State<NC> S;
#pragma unroll
    for (int k = 0; k < NC; k++) {
#pragma unroll
        for (int v = 0; v < V; v++) { S.m[k][v] = S.e[k][v] = S.f[k][v] = S.md[k][v] = LB_NEG; }
    }
    S.jc = lane * NC;
    while (S.jc + NC - 1 <= g.Cn && a_last(sm, S.jc + NC - 1) + 1 < g.u0) S.jc += W;
    S.cwp = sm.colw + S.jc;
#pragma unroll
    for (int k = 0; k < NC; k++) S.w[k] = S.cwp[k];
    S.K = band_key(BIAS + g.u0 - S.jc);
    S.rcp = sm.rowcode + r.row_pad + g.u0 - S.jc;
    S.bp = box + (size_t)g.u0 * AW + lane * NC * V;
    int *ap = sm.acc + lane * NC;
    const int left = (lane + 31) & 31;

    // entry list of the row: the blocks of target group t (targets 4t .. 4t+3) are folded after cell step 4t - 2: their sources lie at
    // least 8 anti-diagonals before the target, so they are final; the last gather lands after step 4t - 1, just before its first target
    // is consumed. The accumulators then cover the anti-diagonals 4t-1 .. 4t+3, which is why a ring of six suffices. The block counts move next to the block
    // offsets (the list is complete: all groups of the row have contributed), and the first block of the next target group is
    // requested one group ahead, so that a fold waits for one memory round trip (the gather of the sources), not for three in a row
    const uint2 *list = r.clist + (size_t)grp.pair * r.clist_cap + lane;
    {
        const int *nblk = r.cnblk + (size_t)grp.pair * LB_ROWS_TG;
        for (int t = lane; t < LB_ROWS_TG; t += 32) {
            const int nb = (t >= 2 && t < g.n_tg) ? min(__ldcg(nblk + t), 2047) : 0;
            sm.gstart[t] = (t < g.n_tg ? sm.gstart[t] & 0xfffff : 0) | (nb << 20);
        }
        __syncwarp();
    }
    const int4 *boxv = (const int4 *)box;
    int4 pm = make_int4(0, 0, 0, 0);
    int pd = LB_NEG, ps = 0;
    bool pending = false;
    {
        const int gs = sm.gstart[min((g.u0 + 5) >> 2, LB_ROWS_TG - 1)];   // the first group that falls due
        if ((gs >> 20) > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(list + (size_t)(gs & 0xfffff) * 32));
    }
    auto land = [&]() {
        // padding entries and entries with a source before u0 carry D = -inf: they fold nothing, and left in they would all hit the
        // same accumulator (one serialised wavefront per lane)
        if (pd >= LB_NEG_LIMIT) {
            atomicMax(sm.acc + ps, pm.x + pd);
            atomicMax(sm.acc + ps + RING * W, pm.y + pd);
            atomicMax(sm.acc + ps + 2 * RING * W, pm.z + pd);
            atomicMax(sm.acc + ps + 3 * RING * W, pm.w + pd);
        }
    };
    auto fold = [&](int u) {   // after cell step u
        __syncwarp();
        if (pending) { land(); pending = false; }
        if (((u + 2) & 3) == 0) {
            const int t = (u + 2) >> 2;
            if (t < g.n_tg) {
                const int gs = sm.gstart[t], gs1 = sm.gstart[t + 1];
                const int nb = gs >> 20;
                // the first block of the next group is pulled into L1 now and read from there at the next fold. (Carrying it in
                // registers does not work: ptxas loads into a temporary and copies it at the end of this fold, i.e. waits here.)
                // L1 is safe for the list: it is complete before any group of the row sweeps, and every group starts with an
                // acquire, which invalidates the SM's L1.
                if ((gs1 >> 20) > 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(list + (size_t)(gs1 & 0xfffff) * 32));
                uint2 en = make_uint2(0u, 0u);
                const uint2 *p = list + (size_t)(gs & 0xfffff) * 32;
#pragma unroll 1
                for (int b = 0; b < nb; b++) {
                    if (b == 0) asm volatile("ld.global.ca.v2.u32 {%0, %1}, [%2];" : "=r"(en.x), "=r"(en.y) : "l"(p) : "memory");
                    else en = __ldcg(p + b * 32);
                    if (pending) land();
                    // sources before this group's first anti-diagonal are -inf in all its layers (and not stored in its box)
                    const bool live = (int)((en.x & 0xffffu) >> LOGW) >= g.u0;
                    pm = __ldcg(boxv + (live ? (en.x & 0xffffu) : 0u));   // cell 0 holds a finite or -inf value (see below)
                    pd = live ? (int)en.y : LB_NEG;
                    ps = (int)(en.x >> 16);
                    pending = true;
                }
            }
        }
        __syncwarp();
    };
    // padding entries and entries whose source lies before u0 read cell 0 and fold it with D = -inf: the result is -inf-like as long
    // as the cell holds a score or -inf, so it must not be left uninitialised when the sweep starts later
    if (g.u0 > 0 && lane == 0) *(int4 *)box = make_int4(LB_NEG, LB_NEG, LB_NEG, LB_NEG);
    __syncwarp();
    int ringoff = ((s0 + g.u0) % RING) * W;
    int u = g.u0;
    int my_jv = -1;   // lane v seeds layer v
#pragma unroll
    for (int v = 0; v < V; v++) if (lane == v) my_jv = g.jv[v];
    for (; u <= g.useed && u <= g.u1; u++) {
        // origin of layer v: M = 0 (init_state, aligner.cc:282). At the origin every other term of the recurrence is -inf for this
        // layer (nothing of it lies above or to the left, no arc match ends there), so a 0 in its accumulator IS the cell value;
        // an origin outside the band stays -inf like any out-of-band cell.
        if (my_jv == u) sm.acc[lane * (RING * W) + ringoff + (u & (W - 1))] = 0;
        __syncwarp();
        dp_step<NC>(S, g, sm.sig, ap + ringoff, gap, gap_open, left, u);
        ringoff = ringoff + W == RING * W ? 0 : ringoff + W;
        fold(u);
    }
    for (; u <= g.u1; u++) {
        dp_step<NC>(S, g, sm.sig, ap + ringoff, gap, gap_open, left, u);
        ringoff = ringoff + W == RING * W ? 0 : ringoff + W;
        fold(u);
    }
    __syncwarp();
}

// D entries of all arc matches of the group's tasks (aligner.cc:574-657), layer v of the box for task v
template <int NC>
__device__ void write_d(const DevCtx &c, const DevPair &pr, const DevGroup &grp, const Geom &g, const int *box, bool nolp) {
    const int lane = threadIdx.x & 31;
    constexpr int W = 32 * NC;
    DevTask tk[V];
#pragma unroll
    for (int v = 0; v < V; v++) tk[v] = c.tasks[grp.task[v < grp.nmem ? v : 0]];
    const DevArcMatch *am = c.am + pr.am_base;
    DevEntry *ent = c.ent + pr.am_base;
    const int sh = nolp ? 2 : 1;
    const bool stacking = c.params.stacking != 0;
    int end[V], total = 0;
#pragma unroll
    for (int v = 0; v < V; v++) { total += v < grp.nmem ? tk[v].run_count : 0; end[v] = total; }
    for (int t = lane; t < total; t += 32) {
        int v = 0, base = 0;
#pragma unroll
        for (int k = 0; k < V - 1; k++) if (t >= end[k]) { v = k + 1; base = end[k]; }
        int run_start = tk[0].run_start;
#pragma unroll
        for (int k = 1; k < V; k++) if (v == k) run_start = tk[k].run_start;
        const int k = run_start + (t - base);
        const DevArcMatch x = am[k];
        if (nolp && (x.inner < 0 || (stacking && x.score_st == LB_NOSTACK))) continue;   // aligner.cc:628-629
        const int ar = (x.ends_a >> 12) & 0xfff, br = (x.ends_b >> 12) & 0xfff;
        const int ip = ar - sh - g.al, jp = br - sh - g.bl0;
        const int mv = __ldcg(box + ((ip + jp) * W + (jp & (W - 1))) * V + v);
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

__device__ __forceinline__ void wait_counter(const int *counter, int need) {
    int have;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(counter) : "memory");
        if (have >= need) break;
        __nanosleep(200);
    }
}

template <int NC>
__device__ void run_group(const DevCtx &c, const RowsCtx &r, const DevGroup &grp, const DevPair &pr, const Geom &g, const Smem &sm, int *box, int need, bool nolp) {
    const int lane = threadIdx.x & 31;
    build_list<NC>(c, r, grp, pr, g, sm);
    __threadfence();
    __syncwarp();
    if (lane == 0) { atomicAdd(r.row_built + grp.pair, 1); wait_counter(r.row_built + grp.pair, need + grp.G); }
    __syncwarp();
    sweep<NC>(c, r, grp, g, sm, box);
    write_d<NC>(c, pr, grp, g, box, nolp);
}

__global__ void __launch_bounds__(32, 18) dfill_rows_kernel(DevCtx c, RowsCtx r, int *cursor) {
    extern __shared__ __align__(16) int smem[];
    const int lane = threadIdx.x;
    Smem sm;
    {
        int *sig = smem;
        sig[lane] = c.params.sigma8[lane]; sig[lane + 32] = c.params.sigma8[lane + 32];
        sm.sig = sig;
        sm.acc = smem + 64;
        sm.gstart = sm.acc + r.acc_words;
        sm.colw = (uint32_t *)(sm.gstart + LB_ROWS_TG);
        sm.rowcode = (uint8_t *)(sm.colw + r.colw_words);
        __syncwarp();
    }
    int *box = r.scratch + (size_t)blockIdx.x * r.scratch_words;
    const bool nolp = c.params.no_lonely_pairs != 0;
    const int n_groups = *r.n_groups;
    for (;;) {
        int k = 0;
        if (lane == 0) k = atomicAdd(cursor, 1);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= n_groups) break;
        const DevGroup grp = r.groups[r.order[k]];
        const DevPair pr = c.pairs[grp.pair];
        Geom g;
        int nc = setup_group(c, r, grp, pr, g, sm);   // touches only the band and the sequences: overlaps the wait
        if (nc > r.nc_max) nc = 0;
        if (nc != 0 && (long long)(g.umax + 1) * (32 * nc * V) > r.scratch_words) nc = -1;
        const int need = r.dep_need[(size_t)grp.pair * r.n_levels + grp.al];
        if (lane == 0) wait_counter(r.dep_done + grp.pair, need);
        __syncwarp();
        if (nc == 1) run_group<1>(c, r, grp, pr, g, sm, box, need, nolp);
        else if (nc == 2) run_group<2>(c, r, grp, pr, g, sm, box, need, nolp);
        else {   // 5: box not supported by this kernel (the host falls back); keep the row's counters moving
            if (lane == 0) { atomicExch(c.error_flag, nc == 0 ? 5 : 2); atomicAdd(r.row_built + grp.pair, 1); }
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(r.dep_done + grp.pair, 1);
    }
}

}  // namespace

int rows_smem_bytes(const RowsCtx &r) { return 64 * 4 + r.acc_words * 4 + LB_ROWS_TG * 4 + r.colw_words * 4 + r.rowcode_bytes; }

cudaError_t rows_configure(int smem_bytes, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(dfill_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lb200_sticky_smem(100, smem_bytes));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, dfill_rows_kernel, 32, smem_bytes);
    return e;
}

void launch_dfill_rows(const DevCtx &c, const RowsCtx &r, int grid, int smem_bytes, int *cursor, cudaStream_t st) {
    dfill_rows_kernel<<<grid, 32, smem_bytes, st>>>(c, r, cursor);
}

}  // namespace lb200
