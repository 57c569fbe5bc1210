/* Minimal stand-in for <ViennaRNA/data_structures.h>, written for the oracle build only.
 * ViennaRNA is not installed in this image; the reference's hot path (Aligner & co) does not
 * call into libRNA when its inputs are PP files, but a few headers mention these types:
 *   pfold_params.hh:61-94  (vrna_md_t, vrna_md_set_default, vrna_md_copy)
 *   rna_data.cc:1731-1773  (vrna_plist_t, MEA)
 * TEST INFRASTRUCTURE ONLY - never linked into the product library. */
#ifndef LB200_ORACLE_VRNA_SHIM_DATA_STRUCTURES_H
#define LB200_ORACLE_VRNA_SHIM_DATA_STRUCTURES_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct vrna_elem_prob_s { int i; int j; float p; int type; } vrna_plist_t;
typedef vrna_plist_t plist;
typedef struct vrna_md_s {
    double temperature; double betaScale; int pf_smooth; int dangles; int special_hp; int noLP;
    int noGU; int noGUclosure; int logML; int circ; int gquad; int uniq_ML; int energy_set;
    int backtrack; char backtrack_type; int compute_bpp; int max_bp_span; int min_loop_size;
    int window_size; int oldAliEn; int ribo; double cv_fact; double nc_fact; double sfact;
} vrna_md_t;
static inline void vrna_md_set_default(vrna_md_t *md) {
    md->temperature = 37.0; md->betaScale = 1.0; md->pf_smooth = 1; md->dangles = 2; md->special_hp = 1;
    md->noLP = 0; md->noGU = 0; md->noGUclosure = 0; md->logML = 0; md->circ = 0; md->gquad = 0;
    md->uniq_ML = 0; md->energy_set = 0; md->backtrack = 1; md->backtrack_type = 'F'; md->compute_bpp = 1;
    md->max_bp_span = -1; md->min_loop_size = 3; md->window_size = -1; md->oldAliEn = 0; md->ribo = 0;
    md->cv_fact = 1.0; md->nc_fact = 1.0; md->sfact = 1.07;
}
static inline vrna_md_t *vrna_md_copy(vrna_md_t *to, const vrna_md_t *from) { if (to && from) *to = *from; return to; }
#ifdef __cplusplus
}
#endif
#endif
