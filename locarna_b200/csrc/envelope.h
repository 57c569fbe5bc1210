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
    double *scratch;              // per CTA 3 (fallback kernel: 6) matrices of (lenA+1) x (lenB+1) doubles
    size_t scratch_doubles;
    double bm[16];                // base similarity (ribosum base match scores or match/mismatch)
    double sw, open, ext, temp, min_prob;
    int local, fe_left1, fe_right1, fe_left2, fe_right2;
};

// max_rows / max_cols = longest first / second sequence + 1 of the batch (sizes of the shared-memory buffers)
cudaError_t launch_envelope(const EnvCtx &e, int n_pairs, int grid, int *cursor, int max_rows, int max_cols, cudaStream_t st);
int envelope_smem_bytes(int max_rows, int max_cols);
// doubles of scratch one CTA needs for matrices of max_cells cells (3 forward matrices; 6 for the global-memory fallback)
size_t envelope_scratch_doubles(size_t max_cells, int max_rows, int max_cols);

}  // namespace lb200
#endif
