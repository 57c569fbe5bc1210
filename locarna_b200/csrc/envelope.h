// Device probability envelope (envelope.cu)
#ifndef LB200_ENVELOPE_H
#define LB200_ENVELOPE_H
#include <cuda_runtime.h>
#include <stdint.h>

namespace lb200 {

struct EnvPair {
    int lenA, lenB;
    int codesA, codesB;   // offsets into codes[]
    int probA, probB;     // offsets into p_up[] / p_down[] (entries 0..len)
    int band;             // offset into band_lo/band_hi/out_lo/out_hi (entries 0..lenA)
    int pad;
};

struct EnvCtx {
    const EnvPair *pairs;
    const uint8_t *codes;
    const double *p_up, *p_down;
    int *band_lo, *band_hi;       // band before the restriction (--max-diff or unrestricted)
    int *out_lo, *out_hi;         // restricted band
    int *out_flag;                // per pair: 1 = uncertain, recompute on the host in long double
    double *scratch;              // per CTA 6 matrices of (lenA+1) x (lenB+1) doubles
    size_t scratch_doubles;
    double bm[16];                // base similarity (ribosum base match scores or match/mismatch)
    double sw, open, ext, temp, min_prob;
    int local, fe_left1, fe_right1, fe_left2, fe_right2;
};

cudaError_t launch_envelope(const EnvCtx &e, int n_pairs, int grid, int *cursor, cudaStream_t st);

}  // namespace lb200
#endif
