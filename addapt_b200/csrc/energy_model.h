// Host side of the energy model: reads a ViennaRNA "parameter file v2.0", derives the integer (MFE)
// and Boltzmann (partition function) tables the kernels consume, evaluates a fixed structure, and
// prepares the ligand-motif record.
//
// Replaces, for the hot path, what the reference gets from libRNA through
//   vrna_md_set_default / vrna_fold_compound        reference src/scoring.cc:80-88
//   vrna_sc_add_hi_motif                            reference src/scoring.cc:94-99
#pragma once
#include <string>

#include "tables.h"

namespace adp {

class EnergyModel {
 public:
  // Throws std::string on malformed input, like the reference's own error convention.
  void load(const std::string& par_path);

  const IntTables& raw() const { return raw_; }   // file values
  const IntTables& mfe() const { return mfe_; }   // dangles / multi / exterior mismatches clamped <= 0
  const PfTables& pf() const { return pf_; }
  double kT_kcal() const { return pf_.kT / 1000.; }
  // Where the numbers came from: "file" for a parameter file as given, "reconstructed" when the file declares itself a
  // reconstruction (the one shipped in params/: int11 / int21 / int22 are rule-generated, not the measured Turner-2004 tables).
  const std::string& provenance() const { return provenance_; }
  const std::string& path() const { return path_; }

  // Free energy (dcal/mol) of a dot-bracket structure under the dangles=2 model (≙ vrna_eval_structure).
  int eval_structure(const std::string& seq, const std::string& db) const;

  // Soft-constraint record for a hairpin motif; energy_kcal = kT*ln(Kd_uM/1e6) (reference scoring.cc:98).
  MotifDev make_motif(const std::string& seq, const std::string& fold, double energy_kcal) const;

  static int encode(char c);                        // A1 C2 G3 U/T4, else 0
  static int pair_type(int a, int b);               // 0 = cannot pair

 private:
  int loop_energy(const int* S, const int* pt, int n, int i, int j) const;
  IntTables raw_{}, mfe_{};
  PfTables pf_{};
  std::string provenance_ = "file", path_;
};

}  // namespace adp
