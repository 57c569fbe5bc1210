// Oracle-build replacement for the reference's rna_ensemble.cc (which needs libRNA).
// Every entry point that would fold a sequence throws LocARNA::failure, so the oracle binaries
// accept PP 2.0 / dot-plot input only (rna_data.cc:54-62 constructs an RnaEnsemble only for
// sequence-only input). TEST INFRASTRUCTURE ONLY.
#include <string>
#include "LocARNA/aux.hh"
#include "LocARNA/multiple_alignment.hh"
#include "LocARNA/pfold_params.hh"
#include "LocARNA/rna_ensemble.hh"
#include <ViennaRNA/MEA.h>

namespace LocARNA {
    class RnaEnsembleImpl {};
    static void nofold() { throw failure("oracle build: no ViennaRNA, cannot fold; use PP input"); }
    RnaEnsemble::RnaEnsemble(const MultipleAlignment &, const PFoldParams &, bool, bool) { nofold(); }
    RnaEnsemble::~RnaEnsemble() {}
    bool RnaEnsemble::has_base_pair_probs() const { return false; }
    bool RnaEnsemble::has_stacking_probs() const { return false; }
    bool RnaEnsemble::has_in_loop_probs() const { return false; }
    const MultipleAlignment &RnaEnsemble::multiple_alignment() const { nofold(); throw 0; }
    size_type RnaEnsemble::length() const { return 0; }
    double RnaEnsemble::min_free_energy() const { return 0; }
    std::string RnaEnsemble::min_free_energy_structure() const { return ""; }
    double RnaEnsemble::arc_prob(size_type, size_type) const { nofold(); return 0; }
    double RnaEnsemble::arc_2_prob(size_type, size_type) const { nofold(); return 0; }
    double RnaEnsemble::arc_in_loop_prob(size_type, size_type, size_type, size_type) const { nofold(); return 0; }
    double RnaEnsemble::arc_external_prob(size_type, size_type) const { nofold(); return 0; }
    double RnaEnsemble::unpaired_in_loop_prob(size_type, size_type, size_type) const { nofold(); return 0; }
    double RnaEnsemble::unpaired_external_prob(size_type) const { nofold(); return 0; }
}
extern "C" float MEA(plist *, char *, double) { throw LocARNA::failure("oracle build: MEA unavailable"); }
