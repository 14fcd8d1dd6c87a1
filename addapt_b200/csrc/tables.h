// Plain-data energy tables shared by the host loader (energy_model.cc) and the sm_100a kernels.
//
// These replace what the reference obtains from ViennaRNA when it calls
// vrna_md_set_default + vrna_fold_compound(..., VRNA_OPTION_PF)   (reference src/scoring.cc:80-88):
// the Turner-2004 nearest-neighbour set at 37 C with dangles=2, as integer dcal/mol (MFE) and as
// Boltzmann factors exp(-E/kT) (partition function).  Nucleotide codes: A1 C2 G3 U4 (0 = other);
// pair types: CG1 GC2 GU3 UG4 AU5 UA6 (7 = non-standard), as in libRNA.
#pragma once
#include <stdint.h>

namespace adp {

constexpr int kMaxLoop = 30;      // largest interior loop enumerated (libRNA MAXLOOP)
constexpr int kTurn = 3;          // minimum hairpin size
constexpr int kInf = 10000000;    // "no structure" for integer energies
constexpr int kMaxLen = 512;      // longest construct the tables are sized for
constexpr int kMaxSpecial = 64;   // special hairpins per class
constexpr int kMaxMotif = 64;     // longest ligand motif
constexpr int kMaxMotifPairs = 32;

// Integer energies, dcal/mol.  `mfe` view: dangles and multi/exterior mismatches clamped to <= 0.
struct IntTables {
  int32_t stack[8][8];
  int32_t hairpin[kMaxLen + 1];   // [u] for every loop size; u > 30 uses the logarithmic extrapolation
  int32_t bulge[31], interior[31];
  int32_t ninio[31];              // min(MAX_NINIO, d * ninio)
  int32_t mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5], mmM[8][5][5], mmExt[8][5][5];
  int32_t dangle5[8][5], dangle3[8][5];
  int32_t int11[8][8][5][5];
  int32_t int21[8][8][5][5][5];
  int32_t int22[8][8][5][5][5][5];
  int32_t ML_base, ML_closing, ML_intern, TerminalAU;
  int32_t n_tetra, n_tri, n_hex;
  uint32_t tetra_code[kMaxSpecial], tri_code[kMaxSpecial], hex_code[kMaxSpecial];  // 2 bits / nt, 5' first
  int32_t tetra_e[kMaxSpecial], tri_e[kMaxSpecial], hex_e[kMaxSpecial];
};

// Boltzmann factors (double).  Loop-length arrays are NOT yet multiplied by the pf_scale powers;
// the kernels apply scale[] explicitly, like McCaskill's recursion in libRNA does.
struct PfTables {
  double stack[8][8];
  double hairpin[kMaxLen + 1];
  double bulge[31], interior[31];
  double ninio[31];
  double mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5], mmM[8][5][5], mmExt[8][5][5];
  double dangle5[8][5], dangle3[8][5];
  double int11[8][8][5][5];
  double int21[8][8][5][5][5];
  double int22[8][8][5][5][5][5];
  double MLbase, MLclosing, MLintern, TermAU;
  double kT;                      // cal/mol
  double pf_scale;                // per-nucleotide scaling factor (no MFE estimate: exp(1.07*185/kT))
  double scale[kMaxLen + 4];      // pf_scale^-k
  double eMLb[kMaxLen + 4];       // (MLbase * 1/pf_scale)^k : k unpaired bases inside a multiloop
  int32_t n_tetra, n_tri, n_hex;
  uint32_t tetra_code[kMaxSpecial], tri_code[kMaxSpecial], hex_code[kMaxSpecial];
  double tetra_x[kMaxSpecial], tri_x[kMaxSpecial], hex_x[kMaxSpecial];
};

// A ligand hairpin motif staged for the device (≙ vrna_sc_add_hi_motif, reference src/scoring.cc:94-99).
struct MotifDev {
  int32_t len;                    // 0 = no motif
  int32_t contrib_dcal;           // (int)(energy*100) + (int)((E_motif - E_empty)*100)
  double boltz;                   // exp(-contrib*10/kT)
  uint8_t code[kMaxMotif];        // nucleotide codes of the motif sequence
  int32_t n_pairs;                // inner pairs of the motif fold (offsets from the 5' end)
  uint8_t pair_i[kMaxMotifPairs], pair_j[kMaxMotifPairs];
};

}  // namespace adp
