// See energy_model.h.  Host-only C++; no CUDA here.
#include "energy_model.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <vector>

namespace adp {
namespace {

constexpr double kGasConst = 1.98717;  // cal/(K mol), libRNA's GASCONST
constexpr double kK0 = 273.15;

struct Section {
  std::vector<double> nums;                                  // numeric sections
  std::vector<std::pair<std::string, int>> loops;            // special-hairpin sections
};

std::string strip_comments(const std::string& text) {
  std::string out;
  out.reserve(text.size());
  size_t p = 0;
  while (p < text.size()) {
    size_t a = text.find("/*", p);
    if (a == std::string::npos) { out.append(text, p, std::string::npos); break; }
    out.append(text, p, a - p);
    size_t b = text.find("*/", a + 2);
    if (b == std::string::npos) break;
    for (size_t q = a; q < b + 2; q++) if (text[q] == '\n') out.push_back('\n');
    p = b + 2;
  }
  return out;
}

std::map<std::string, Section> read_sections(const std::string& path) {
  std::ifstream in(path);
  if (!in) throw std::string("cannot open parameter file '") + path + "'";
  std::stringstream ss;
  ss << in.rdbuf();
  std::istringstream lines(strip_comments(ss.str()));
  std::map<std::string, Section> secs;
  Section* cur = nullptr;
  bool is_loops = false;
  std::string line;
  while (std::getline(lines, line)) {
    if (line.size() >= 2 && line[0] == '#' && line[1] == '#') continue;   // "## RNAfold parameter file"
    if (!line.empty() && line[0] == '#') {
      std::istringstream ls(line.substr(1));
      std::string name;
      ls >> name;
      if (name.empty() || name == "END") { cur = nullptr; continue; }
      cur = &secs[name];
      is_loops = (name == "Tetraloops" || name == "Triloops" || name == "Hexaloops");
      continue;
    }
    if (!cur) continue;
    std::istringstream ls(line);
    if (is_loops) {
      std::string w;
      int e;
      if (ls >> w >> e) cur->loops.emplace_back(w, e);
      continue;
    }
    std::string tok;
    while (ls >> tok) {
      if (tok == "INF") cur->nums.push_back(kInf);
      else if (tok == "DEF") cur->nums.push_back(-50);
      else {
        char* end = nullptr;
        double v = std::strtod(tok.c_str(), &end);
        if (end != tok.c_str()) cur->nums.push_back(v);
      }
    }
  }
  return secs;
}

const std::vector<double>& need(const std::map<std::string, Section>& s, const char* name, size_t at_least) {
  auto it = s.find(name);
  if (it == s.end()) throw std::string("parameter file: section '") + name + "' is missing";
  if (it->second.nums.size() < at_least)
    throw std::string("parameter file: section '") + name + "' is too short";
  return it->second.nums;
}

// tables [pair type 1..6|7][5^rank]
void fill_by_type(int32_t* dst, const std::vector<double>& v, int rank, const char* name) {
  size_t blk = 1;
  for (int r = 0; r < rank; r++) blk *= 5;
  size_t types = v.size() / blk;
  if (v.size() % blk || (types != 6 && types != 7)) throw std::string("parameter file: bad size for section ") + name;
  for (size_t t = 0; t < types; t++)
    for (size_t k = 0; k < blk; k++) dst[(t + 1) * blk + k] = (int32_t)v[t * blk + k];
}

// tables [type][type][5^rank]
void fill_by_type2(int32_t* dst, const std::vector<double>& v, int rank, const char* name) {
  size_t blk = 1;
  for (int r = 0; r < rank; r++) blk *= 5;
  size_t tt = v.size() / blk, types = tt == 36 ? 6 : tt == 49 ? 7 : 0;
  if (v.size() % blk || !types) throw std::string("parameter file: bad size for section ") + name;
  for (size_t a = 0; a < types; a++)
    for (size_t b = 0; b < types; b++)
      for (size_t k = 0; k < blk; k++) dst[((a + 1) * 8 + (b + 1)) * blk + k] = (int32_t)v[(a * types + b) * blk + k];
}

uint32_t pack_loop(const std::string& w) {
  uint32_t code = 0;
  for (size_t p = 0; p < w.size(); p++) {
    int c = EnergyModel::encode(w[p]);
    if (!c) throw std::string("parameter file: special hairpin '") + w + "' has a non-ACGU letter";
    code |= (uint32_t)(c - 1) << (2 * p);
  }
  return code;
}

void load_specials(const std::map<std::string, Section>& s, const char* name, size_t width, int32_t* n, uint32_t* code, int32_t* e) {
  *n = 0;
  auto it = s.find(name);
  if (it == s.end()) return;
  for (auto& le : it->second.loops) {
    if (le.first.size() != width || *n >= kMaxSpecial) throw std::string("parameter file: bad entry in ") + name;
    code[*n] = pack_loop(le.first);
    e[*n] = le.second;
    (*n)++;
  }
}

// libRNA's SMOOTH(): a differentiable stand-in for min(0, -E) used for dangles and multi/exterior mismatches
// in the partition function only.
double smooth(double x) {
  const double s = 10.;
  if (x / s < -1.2283697) return 0.;
  if (x / s > 0.8660254) return x;
  double w = std::sin(x / s - 0.34242663) + 1.;
  return s * 0.38490018 * w * w;
}

int bracket_partners(const std::string& db, std::vector<int>& pt) {   // 1-based partner table, 0 = unpaired
  int n = (int)db.size();
  pt.assign(n + 2, 0);
  std::vector<int> open;
  for (int i = 1; i <= n; i++) {
    char c = db[i - 1];
    if (c == '(') open.push_back(i);
    else if (c == ')') {
      if (open.empty()) return -1;
      pt[open.back()] = i;
      pt[i] = open.back();
      open.pop_back();
    }
  }
  return open.empty() ? 0 : -1;
}

}  // namespace

int EnergyModel::encode(char c) {
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'U': case 'u': case 'T': case 't': return 4;
    default: return 0;
  }
}

int EnergyModel::pair_type(int a, int b) {
  static const int T[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
  return T[a][b];
}

void EnergyModel::load(const std::string& path) {
  auto secs = read_sections(path);
  path_ = path;
  provenance_ = "file";
  {
    std::ifstream head(path);
    std::string line;
    for (int k = 0; k < 12 && std::getline(head, line); k++)
      if (line.find("RECONSTRUCTION") != std::string::npos) provenance_ = "reconstructed";
  }
  IntTables& R = raw_;
  std::memset(&R, 0, sizeof R);

  {
    const auto& v = need(secs, "stack", 36);
    size_t types = v.size() == 49 ? 7 : v.size() == 36 ? 6 : 0;
    if (!types) throw std::string("parameter file: bad size for section stack");
    for (size_t a = 0; a < types; a++)
      for (size_t b = 0; b < types; b++) R.stack[a + 1][b + 1] = (int32_t)v[a * types + b];
  }
  fill_by_type(&R.mmH[0][0][0], need(secs, "mismatch_hairpin", 150), 2, "mismatch_hairpin");
  fill_by_type(&R.mmI[0][0][0], need(secs, "mismatch_interior", 150), 2, "mismatch_interior");
  fill_by_type(&R.mm1nI[0][0][0], need(secs, "mismatch_interior_1n", 150), 2, "mismatch_interior_1n");
  fill_by_type(&R.mm23I[0][0][0], need(secs, "mismatch_interior_23", 150), 2, "mismatch_interior_23");
  fill_by_type(&R.mmM[0][0][0], need(secs, "mismatch_multi", 150), 2, "mismatch_multi");
  fill_by_type(&R.mmExt[0][0][0], need(secs, "mismatch_exterior", 150), 2, "mismatch_exterior");
  fill_by_type(&R.dangle5[0][0], need(secs, "dangle5", 30), 1, "dangle5");
  fill_by_type(&R.dangle3[0][0], need(secs, "dangle3", 30), 1, "dangle3");
  fill_by_type2(&R.int11[0][0][0][0], need(secs, "int11", 900), 2, "int11");
  fill_by_type2(&R.int21[0][0][0][0][0], need(secs, "int21", 4500), 3, "int21");
  {
    const auto& v = need(secs, "int22", 9216);
    if (v.size() != 9216) throw std::string("parameter file: int22 must hold 6*6*4^4 numbers");
    size_t q = 0;
    for (int a = 1; a <= 6; a++)
      for (int b = 1; b <= 6; b++)
        for (int w = 1; w <= 4; w++)
          for (int x = 1; x <= 4; x++)
            for (int y = 1; y <= 4; y++)
              for (int z = 1; z <= 4; z++) R.int22[a][b][w][x][y][z] = (int32_t)v[q++];
  }
  const auto& hp = need(secs, "hairpin", 31);
  const auto& bu = need(secs, "bulge", 31);
  const auto& in = need(secs, "interior", 31);
  const auto& ml = need(secs, "ML_params", 6);
  const auto& ni = need(secs, "NINIO", 3);
  const auto& mi = need(secs, "Misc", 5);
  R.ML_base = (int32_t)ml[0];
  R.ML_closing = (int32_t)ml[2];
  R.ML_intern = (int32_t)ml[4];
  R.TerminalAU = (int32_t)mi[2];
  const double lxc = mi[4];
  const int ninio = (int)ni[0], max_ninio = (int)ni[2];
  for (int u = 0; u <= 30; u++) {
    R.hairpin[u] = (int32_t)hp[u];
    R.bulge[u] = (int32_t)bu[u];
    R.interior[u] = (int32_t)in[u];
    R.ninio[u] = std::min(max_ninio, u * ninio);
  }
  for (int u = 31; u <= kMaxLen; u++) R.hairpin[u] = R.hairpin[30] + (int)(lxc * std::log(u / 30.));
  load_specials(secs, "Tetraloops", 6, &R.n_tetra, R.tetra_code, R.tetra_e);
  load_specials(secs, "Triloops", 5, &R.n_tri, R.tri_code, R.tri_e);
  load_specials(secs, "Hexaloops", 8, &R.n_hex, R.hex_code, R.hex_e);

  // MFE view
  mfe_ = R;
  for (int t = 0; t < 8; t++)
    for (int a = 0; a < 5; a++) {
      mfe_.dangle5[t][a] = std::min(0, R.dangle5[t][a]);
      mfe_.dangle3[t][a] = std::min(0, R.dangle3[t][a]);
      for (int b = 0; b < 5; b++) {
        mfe_.mmM[t][a][b] = std::min(0, R.mmM[t][a][b]);
        mfe_.mmExt[t][a][b] = std::min(0, R.mmExt[t][a][b]);
      }
    }

  // Boltzmann view
  PfTables& P = pf_;
  std::memset(&P, 0, sizeof P);
  const double kT = (37. + kK0) * kGasConst;
  P.kT = kT;
  P.pf_scale = std::exp(-(1.07 * (-185. + (37. - 37.) * 7.27)) / kT);
  if (P.pf_scale < 1.) P.pf_scale = 1.;
  auto boltz = [kT](double e) { return std::exp(-e * 10. / kT); };
  auto boltz_smooth = [kT](double e) { return std::exp(smooth(-e) * 10. / kT); };
  for (int u = 0; u <= 30; u++) {
    P.hairpin[u] = boltz(R.hairpin[u]);
    P.bulge[u] = boltz(R.bulge[u]);
    P.interior[u] = boltz(R.interior[u]);
    P.ninio[u] = boltz(R.ninio[u]);
  }
  for (int u = 31; u <= kMaxLen; u++) P.hairpin[u] = P.hairpin[30] * std::exp(-(lxc * std::log(u / 30.)) * 10. / kT);
  P.MLbase = boltz(R.ML_base);
  P.MLclosing = boltz(R.ML_closing);
  P.MLintern = boltz(R.ML_intern);
  P.TermAU = boltz(R.TerminalAU);
  for (int a = 0; a < 8; a++)
    for (int b = 0; b < 8; b++) {
      P.stack[a][b] = boltz(R.stack[a][b]);
      for (int w = 0; w < 5; w++)
        for (int x = 0; x < 5; x++) {
          P.int11[a][b][w][x] = boltz(R.int11[a][b][w][x]);
          for (int y = 0; y < 5; y++) {
            P.int21[a][b][w][x][y] = boltz(R.int21[a][b][w][x][y]);
            for (int z = 0; z < 5; z++) P.int22[a][b][w][x][y][z] = boltz(R.int22[a][b][w][x][y][z]);
          }
        }
    }
  for (int t = 0; t < 8; t++)
    for (int a = 0; a < 5; a++) {
      P.dangle5[t][a] = boltz_smooth(R.dangle5[t][a]);
      P.dangle3[t][a] = boltz_smooth(R.dangle3[t][a]);
      for (int b = 0; b < 5; b++) {
        P.mmH[t][a][b] = boltz(R.mmH[t][a][b]);
        P.mmI[t][a][b] = boltz(R.mmI[t][a][b]);
        P.mm1nI[t][a][b] = boltz(R.mm1nI[t][a][b]);
        P.mm23I[t][a][b] = boltz(R.mm23I[t][a][b]);
        P.mmM[t][a][b] = boltz_smooth(R.mmM[t][a][b]);
        P.mmExt[t][a][b] = boltz_smooth(R.mmExt[t][a][b]);
      }
    }
  P.scale[0] = 1.;
  P.eMLb[0] = 1.;
  for (int k = 1; k < kMaxLen + 4; k++) {
    P.scale[k] = P.scale[k - 1] / P.pf_scale;
    P.eMLb[k] = std::pow(P.MLbase, (double)k) * P.scale[k];
  }
  P.n_tetra = R.n_tetra; P.n_tri = R.n_tri; P.n_hex = R.n_hex;
  for (int k = 0; k < kMaxSpecial; k++) {
    P.tetra_code[k] = R.tetra_code[k]; P.tri_code[k] = R.tri_code[k]; P.hex_code[k] = R.hex_code[k];
    P.tetra_x[k] = boltz(R.tetra_e[k]); P.tri_x[k] = boltz(R.tri_e[k]); P.hex_x[k] = boltz(R.hex_e[k]);
  }
}

// ---------------------------------------------------------------------------------------------------
// structure evaluation (needed for the motif contribution; also a public utility)
namespace {

int hairpin_energy(const IntTables& P, const int* S, int i, int j, int type) {
  int u = j - i - 1;
  int e = P.hairpin[std::min(u, kMaxLen)];
  if (u < 3) return e;
  bool clean = true;
  uint32_t code = 0;
  if (u == 3 || u == 4 || u == 6) {
    for (int p = 0; p < u + 2; p++) {
      if (!S[i + p]) { clean = false; break; }
      code |= (uint32_t)(S[i + p] - 1) << (2 * p);
    }
    if (clean) {
      const uint32_t* codes = u == 4 ? P.tetra_code : u == 6 ? P.hex_code : P.tri_code;
      const int32_t* es = u == 4 ? P.tetra_e : u == 6 ? P.hex_e : P.tri_e;
      int cnt = u == 4 ? P.n_tetra : u == 6 ? P.n_hex : P.n_tri;
      for (int k = 0; k < cnt; k++) if (codes[k] == code) return es[k];
    }
  }
  if (u == 3) return e + (type > 2 ? P.TerminalAU : 0);
  return e + P.mmH[type][S[i + 1]][S[j - 1]];
}

int interior_energy(const IntTables& P, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  int big = std::max(u1, u2), small = std::min(u1, u2);
  if (big == 0) return P.stack[type][type2];
  if (small == 0) {
    int e = P.bulge[big];
    if (big == 1) return e + P.stack[type][type2];
    return e + (type > 2 ? P.TerminalAU : 0) + (type2 > 2 ? P.TerminalAU : 0);
  }
  if (small == 1) {
    if (big == 1) return P.int11[type][type2][si1][sj1];
    if (big == 2) return u1 == 1 ? P.int21[type][type2][si1][sq1][sj1] : P.int21[type2][type][sq1][si1][sp1];
    return P.interior[big + 1] + P.ninio[big - 1] + P.mm1nI[type][si1][sj1] + P.mm1nI[type2][sq1][sp1];
  }
  if (small == 2 && big == 2) return P.int22[type][type2][si1][sp1][sq1][sj1];
  if (small == 2 && big == 3) return P.interior[5] + P.ninio[1] + P.mm23I[type][si1][sj1] + P.mm23I[type2][sq1][sp1];
  return P.interior[u1 + u2] + P.ninio[big - small] + P.mmI[type][si1][sj1] + P.mmI[type2][sq1][sp1];
}

// a stem branching off a multiloop (ext=false) or the exterior loop (ext=true); s5/s3 < 0 = no neighbour
int stem_energy(const IntTables& P, int type, int s5, int s3, bool ext) {
  int e = 0;
  if (s5 >= 0 && s3 >= 0) e = ext ? P.mmExt[type][s5][s3] : P.mmM[type][s5][s3];
  else if (s5 >= 0) e = P.dangle5[type][s5];
  else if (s3 >= 0) e = P.dangle3[type][s3];
  if (type > 2) e += P.TerminalAU;
  if (!ext) e += P.ML_intern;
  return e;
}

const int kRtype[8] = {0, 2, 1, 4, 3, 6, 5, 7};

}  // namespace

int EnergyModel::loop_energy(const int* S, const int* pt, int n, int i, int j) const {
  const IntTables& P = mfe_;
  int type = pair_type(S[i], S[j]);
  if (!type) type = 7;
  std::vector<std::pair<int, int>> kids;
  for (int p = i + 1; p < j;) {
    if (pt[p] > p) { kids.emplace_back(p, pt[p]); p = pt[p] + 1; }
    else p++;
  }
  if (kids.empty()) return hairpin_energy(P, S, i, j, type);
  int total = 0;
  for (auto& kl : kids) total += loop_energy(S, pt, n, kl.first, kl.second);
  if (kids.size() == 1) {
    int k = kids[0].first, l = kids[0].second, t2 = pair_type(S[k], S[l]);
    if (!t2) t2 = 7;
    return total + interior_energy(P, k - i - 1, j - l - 1, type, kRtype[t2], S[i + 1], S[j - 1], S[k - 1], S[l + 1]);
  }
  int unpaired = j - i - 1;
  total += P.ML_closing + stem_energy(P, kRtype[type], S[j - 1], S[i + 1], false);
  for (auto& kl : kids) {
    int t2 = pair_type(S[kl.first], S[kl.second]);
    if (!t2) t2 = 7;
    total += stem_energy(P, t2, S[kl.first - 1], S[kl.second + 1], false);
    unpaired -= kl.second - kl.first + 1;
  }
  return total + unpaired * P.ML_base;
}

int EnergyModel::eval_structure(const std::string& seq, const std::string& db) const {
  int n = (int)seq.size();
  if ((int)db.size() != n) throw std::string("eval_structure: sequence and structure differ in length");
  std::vector<int> S(n + 2, 0), pt;
  for (int i = 1; i <= n; i++) S[i] = encode(seq[i - 1]);
  if (bracket_partners(db, pt)) throw std::string("eval_structure: unbalanced brackets");
  int e = 0;
  for (int p = 1; p <= n;) {
    if (pt[p] > p) {
      int q = pt[p], t = pair_type(S[p], S[q]);
      if (!t) t = 7;
      e += stem_energy(mfe_, t, p > 1 ? S[p - 1] : -1, q < n ? S[q + 1] : -1, true) + loop_energy(S.data(), pt.data(), n, p, q);
      p = q + 1;
    } else p++;
  }
  return e;
}

MotifDev EnergyModel::make_motif(const std::string& seq, const std::string& fold, double energy_kcal) const {
  MotifDev m;
  std::memset(&m, 0, sizeof m);
  int len = (int)seq.size();
  if (len < 5 || len > kMaxMotif) throw std::string("aptamer motif length must be 5..64");
  if ((int)fold.size() != len) throw std::string("aptamer sequence and fold differ in length");
  if (fold.find('&') != std::string::npos) throw std::string("only hairpin motifs (no '&') are supported");
  if (fold[0] != '(' || fold[len - 1] != ')') throw std::string("aptamer fold must be closed by its first and last base");
  std::vector<int> pt;
  if (bracket_partners(fold, pt) || pt[1] != len) throw std::string("aptamer fold is not a single closed hairpin motif");
  std::string empty(len, '.');
  empty[0] = '(';
  empty[len - 1] = ')';
  // vrna_sc_add_hi_motif works on float kcal/mol and truncates toward zero when converting to dcal/mol
  float e_alt = (float)((double)eval_structure(seq, empty) / 100.);
  float e_mot = (float)((double)eval_structure(seq, fold) / 100.);
  float diff = e_mot - e_alt;
  m.len = len;
  m.contrib_dcal = (int)(energy_kcal * 100.) + (int)(diff * 100.);
  m.boltz = std::exp(-(double)m.contrib_dcal * 10. / pf_.kT);
  for (int p = 0; p < len; p++) {
    m.code[p] = (uint8_t)encode(seq[p]);
    if (!m.code[p]) throw std::string("aptamer sequence has a non-ACGU letter");
  }
  for (int a = 2; a < len; a++)
    if (pt[a] > a) {
      if (m.n_pairs >= kMaxMotifPairs) throw std::string("aptamer fold has too many pairs");
      m.pair_i[m.n_pairs] = (uint8_t)(a - 1);
      m.pair_j[m.n_pairs] = (uint8_t)(pt[a] - 1);
      m.n_pairs++;
    }
  return m;
}

}  // namespace adp
