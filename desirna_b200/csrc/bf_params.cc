// bf_params.cc -- host-side parameter loader: "RNAfold parameter file v2.0" -> BfParams image.
//
// Stands in for RNA.params_load() as called at DesiRNA.py:455-456.  Sections and their
// shapes: SURVEY.md A.2 (stack 7x7; six mismatch tables 7x5x5; dangle5/3 7x5; int11
// 7x7x5x5; int21 7x7x5x5x5; int22 6x6x4x4x4x4; hairpin/bulge/interior[31]; NINIO;
// ML_params; Misc; Tri/Tetra/Hexaloops).  The *_enthalpies twins are skipped: the
// reference never changes the temperature (37 C), so they have no effect.
#include "bf_params.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

const uint32_t kImageMagic = 0x42465031u;  // "BFP1"

struct Section {
  std::vector<std::string> tok;
  size_t pos = 0;
  bool has(size_t k = 1) const { return pos + k <= tok.size(); }
  int next_int(bool *ok) {
    if (!has()) { *ok = false; return 0; }
    const std::string &s = tok[pos++];
    if (s == "INF") return BF_INF;
    if (s == "DEF") return -50;
    char *end = nullptr;
    long v = std::strtol(s.c_str(), &end, 10);
    if (end == s.c_str()) *ok = false;
    return (int)v;
  }
};

int base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'U': case 'u': case 'T': case 't': return 4;
    default: return 0;
  }
}

// 2 bits per base, first base in the most significant position; -1 if a base is not ACGU
int loop_key(const std::string &s) {
  int key = 0;
  for (char c : s) {
    int b = base_code(c);
    if (!b) return -1;
    key = (key << 2) | (b - 1);
  }
  return key;
}

double smooth_weight(int g, double kT) {
  // ViennaRNA SMOOTH(-G), SCALE = 10, pf_smooth on
  double x = -(double)g, xs = x / 10.0, y;
  if (xs < -1.2283697) y = 0.0;
  else if (xs > 0.8660254) y = x;
  else { double s = std::sin(xs - 0.34242663) + 1.0; y = 10.0 * 0.38490018 * s * s; }
  return std::exp(y * 10.0 / kT);
}

inline double bz(int e, double kT) { return std::exp(-(double)e * 10.0 / kT); }

void derive(BfParams *P, const int raw_mmM[8][5][5], const int raw_mmE[8][5][5], const int raw_d5[8][5], const int raw_d3[8][5]) {
  const double kT = P->kT;
  for (int t = 0; t < 8; t++)
    for (int a = 0; a < 5; a++) {
      P->si.dangle5[t][a] = raw_d5[t][a] > 0 ? 0 : raw_d5[t][a];
      P->si.dangle3[t][a] = raw_d3[t][a] > 0 ? 0 : raw_d3[t][a];
      P->sd.x_d5[t][a] = smooth_weight(raw_d5[t][a], kT);
      P->sd.x_d3[t][a] = smooth_weight(raw_d3[t][a], kT);
      for (int b = 0; b < 5; b++) {
        P->si.mmM[t][a][b] = raw_mmM[t][a][b] > 0 ? 0 : raw_mmM[t][a][b];
        P->si.mmE[t][a][b] = raw_mmE[t][a][b] > 0 ? 0 : raw_mmE[t][a][b];
        P->sd.x_mmM[t][a][b] = smooth_weight(raw_mmM[t][a][b], kT);
        P->sd.x_mmE[t][a][b] = smooth_weight(raw_mmE[t][a][b], kT);
        P->sd.x_mmH[t][a][b] = bz(P->si.mmH[t][a][b], kT);
        P->sd.x_mmI[t][a][b] = bz(P->si.mmI[t][a][b], kT);
        P->sd.x_mm1nI[t][a][b] = bz(P->si.mm1nI[t][a][b], kT);
        P->sd.x_mm23I[t][a][b] = bz(P->si.mm23I[t][a][b], kT);
      }
    }
  // non-standard pair type (7) rows of int22 = max over the six standard partners (A.2)
  for (int c = 1; c < 5; c++) for (int d = 1; d < 5; d++) for (int e = 1; e < 5; e++) for (int g = 1; g < 5; g++) {
    int mall = -BF_INF;
    for (int a = 1; a <= 6; a++) {
      int m1 = -BF_INF, m2 = -BF_INF;
      for (int b = 1; b <= 6; b++) {
        if (P->int22[a][b][c][d][e][g] > m1) m1 = P->int22[a][b][c][d][e][g];
        if (P->int22[b][a][c][d][e][g] > m2) m2 = P->int22[b][a][c][d][e][g];
      }
      P->int22[a][7][c][d][e][g] = m1;
      P->int22[7][a][c][d][e][g] = m2;
      if (m1 > mall) mall = m1;
    }
    P->int22[7][7][c][d][e][g] = mall;
  }
  for (int a = 0; a < 8; a++) for (int b = 0; b < 8; b++) {
    P->sd.x_stack[a][b] = bz(P->si.stack[a][b], kT);
    for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) {
      P->x_int11[a][b][c][d] = bz(P->int11[a][b][c][d], kT);
      for (int e = 0; e < 5; e++) {
        P->x_int21[a][b][c][d][e] = bz(P->int21[a][b][c][d][e], kT);
        for (int g = 0; g < 5; g++) P->x_int22[a][b][c][d][e][g] = bz(P->int22[a][b][c][d][e][g], kT);
      }
    }
  }
  for (int u = 0; u <= 30; u++) {
    P->sd.x_hairpin[u] = bz(P->si.hairpin[u], kT);
    P->sd.x_bulge[u] = bz(P->si.bulge[u], kT);
    P->sd.x_interior[u] = bz(P->si.interior[u], kT);
    int nin = u * P->si.ninio_m; if (nin > P->si.ninio_max) nin = P->si.ninio_max;
    P->si.ninio[u] = nin;
    P->sd.x_ninio[u] = bz(nin, kT);
  }
  for (int u = 0; u < BF_EXT_TAB; u++) {
    if (u <= 30) { P->ext_log[u] = 0; P->x_hp_big[u] = P->sd.x_hairpin[u]; continue; }
    double l = P->lxc * std::log((double)u / 30.0);
    P->ext_log[u] = (int)l;
    P->x_hp_big[u] = bz(P->si.hairpin[30], kT) * std::exp(-l * 10.0 / kT);
  }
  P->sd.x_MLbase = bz(P->si.MLbase, kT); P->sd.x_MLclosing = bz(P->si.MLclosing, kT); P->sd.x_MLintern = bz(P->si.MLintern, kT);
  P->sd.x_DuplexInit = bz(P->si.DuplexInit, kT); P->sd.x_TerminalAU = bz(P->si.TerminalAU, kT);
  for (int k = 0; k < 4096; k++) P->x_tetra[k] = P->tetra_e[k] == BF_NO_SPECIAL ? 0.0 : bz(P->tetra_e[k], kT);
  for (int k = 0; k < 1024; k++) P->x_tri[k] = P->tri_e[k] == BF_NO_SPECIAL ? 0.0 : bz(P->tri_e[k], kT);
  for (int k = 0; k < P->n_hexa; k++) P->x_hexa[k] = bz(P->hexa_e[k], kT);
}

}  // namespace

int bf_params_parse_file(const char *path, BfParams *P, std::string *err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { if (err) *err = std::string("cannot open parameter file: ") + path; return 1; }
  std::stringstream ss; ss << f.rdbuf();
  std::string text = ss.str();
  if (text.find("RNAfold parameter file v2.0") == std::string::npos) { if (err) *err = "not an 'RNAfold parameter file v2.0'"; return 2; }
  // strip /* ... */ comments
  for (size_t p = 0; (p = text.find("/*", p)) != std::string::npos;) {
    size_t e = text.find("*/", p + 2);
    if (e == std::string::npos) e = text.size() - 2;
    for (size_t k = p; k < e + 2; k++) if (text[k] != '\n') text[k] = ' ';
    p = e + 2;
  }
  std::map<std::string, Section> sec;
  std::string cur;
  std::istringstream in(text);
  for (std::string line; std::getline(in, line);) {
    size_t a = line.find_first_not_of(" \t\r");
    if (a == std::string::npos) continue;
    if (line[a] == '#') {
      if (a + 1 < line.size() && line[a + 1] == '#') continue;
      std::istringstream h(line.substr(a + 1)); h >> cur;
      continue;
    }
    if (cur.empty()) continue;
    std::istringstream ls(line);
    for (std::string t; ls >> t;) sec[cur].tok.push_back(t);
  }
  std::memset(P, 0, sizeof(BfParams));
  P->kT = (37.0 + 273.15) * 1.98717;
  P->sd.kT = P->kT;
  bool ok = true;
  static int raw_mmM[8][5][5], raw_mmE[8][5][5], raw_d5[8][5], raw_d3[8][5];
  std::memset(raw_mmM, 0, sizeof raw_mmM); std::memset(raw_mmE, 0, sizeof raw_mmE);
  std::memset(raw_d5, 0, sizeof raw_d5); std::memset(raw_d3, 0, sizeof raw_d3);
  const char *need[] = {"stack", "mismatch_hairpin", "mismatch_interior", "mismatch_interior_1n", "mismatch_interior_23", "mismatch_multi",
                        "mismatch_exterior", "dangle5", "dangle3", "int11", "int21", "int22", "hairpin", "bulge", "interior", "NINIO", "ML_params", "Misc"};
  for (const char *nm : need) if (!sec.count(nm)) { if (err) *err = std::string("missing section # ") + nm; return 3; }
  { Section &s = sec["stack"]; for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) P->si.stack[a][b] = s.next_int(&ok); }
  auto rd_mm = [&](const char *nm, int (*dst)[5][5]) { Section &s = sec[nm]; for (int t = 1; t <= 7; t++) for (int a = 0; a < 5; a++) for (int b = 0; b < 5; b++) dst[t][a][b] = s.next_int(&ok); };
  rd_mm("mismatch_hairpin", P->si.mmH); rd_mm("mismatch_interior", P->si.mmI); rd_mm("mismatch_interior_1n", P->si.mm1nI); rd_mm("mismatch_interior_23", P->si.mm23I);
  rd_mm("mismatch_multi", raw_mmM); rd_mm("mismatch_exterior", raw_mmE);
  { Section &s = sec["dangle5"]; for (int t = 1; t <= 7; t++) for (int a = 0; a < 5; a++) raw_d5[t][a] = s.next_int(&ok); }
  { Section &s = sec["dangle3"]; for (int t = 1; t <= 7; t++) for (int a = 0; a < 5; a++) raw_d3[t][a] = s.next_int(&ok); }
  { Section &s = sec["int11"]; for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) P->int11[a][b][c][d] = s.next_int(&ok); }
  { Section &s = sec["int21"]; for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) for (int e = 0; e < 5; e++) P->int21[a][b][c][d][e] = s.next_int(&ok); }
  { Section &s = sec["int22"]; for (int a = 1; a <= 6; a++) for (int b = 1; b <= 6; b++) for (int c = 1; c < 5; c++) for (int d = 1; d < 5; d++) for (int e = 1; e < 5; e++) for (int g = 1; g < 5; g++) P->int22[a][b][c][d][e][g] = s.next_int(&ok); }
  { Section &s = sec["hairpin"]; for (int u = 0; u <= 30; u++) P->si.hairpin[u] = s.next_int(&ok); }
  { Section &s = sec["bulge"]; for (int u = 0; u <= 30; u++) P->si.bulge[u] = s.next_int(&ok); }
  { Section &s = sec["interior"]; for (int u = 0; u <= 30; u++) P->si.interior[u] = s.next_int(&ok); }
  { Section &s = sec["NINIO"]; P->si.ninio_m = s.next_int(&ok); (void)s.next_int(&ok); P->si.ninio_max = s.next_int(&ok); }
  { Section &s = sec["ML_params"]; P->si.MLbase = s.next_int(&ok); (void)s.next_int(&ok); P->si.MLclosing = s.next_int(&ok); (void)s.next_int(&ok); P->si.MLintern = s.next_int(&ok); (void)s.next_int(&ok); }
  { Section &s = sec["Misc"]; P->si.DuplexInit = s.next_int(&ok); (void)s.next_int(&ok); P->si.TerminalAU = s.next_int(&ok); (void)s.next_int(&ok);
    if (s.has()) P->lxc = std::atof(s.tok[s.pos++].c_str()); else ok = false; }
  if (!ok) { if (err) *err = "parameter file: a section is shorter than its declared shape"; return 4; }
  for (int k = 0; k < 4096; k++) P->tetra_e[k] = BF_NO_SPECIAL;
  for (int k = 0; k < 1024; k++) P->tri_e[k] = BF_NO_SPECIAL;
  auto rd_loops = [&](const char *nm, size_t len) -> int {
    if (!sec.count(nm)) return 0;
    Section &s = sec[nm];
    for (; s.has(3); s.pos += 3) {
      const std::string &q = s.tok[s.pos];
      int e = (s.tok[s.pos + 1] == "INF") ? BF_INF : std::atoi(s.tok[s.pos + 1].c_str());
      int key = loop_key(q);
      if (q.size() != len || key < 0) { if (err) *err = std::string("bad special loop entry: ") + q; return 5; }
      if (len == 6) { if (P->tetra_e[key] == BF_NO_SPECIAL) { P->tetra_e[key] = e; P->n_tetra++; } }
      else if (len == 5) { if (P->tri_e[key] == BF_NO_SPECIAL) { P->tri_e[key] = e; P->n_tri++; } }
      else { if (P->n_hexa >= BF_MAX_HEXA) { if (err) *err = "too many hexaloops"; return 6; } P->hexa_key[P->n_hexa] = key; P->hexa_e[P->n_hexa++] = e; }
    }
    return 0;
  };
  int rc;
  if ((rc = rd_loops("Triloops", 5)) || (rc = rd_loops("Tetraloops", 6)) || (rc = rd_loops("Hexaloops", 8))) return rc;
  derive(P, raw_mmM, raw_mmE, raw_d5, raw_d3);
  return 0;
}

int bf_params_save_image(const char *path, const BfParams *p, std::string *err) {
  FILE *f = std::fopen(path, "wb");
  if (!f) { if (err) *err = std::string("cannot write ") + path; return 1; }
  uint32_t hdr[2] = {kImageMagic, (uint32_t)sizeof(BfParams)};
  bool ok = std::fwrite(hdr, sizeof hdr, 1, f) == 1 && std::fwrite(p, sizeof(BfParams), 1, f) == 1;
  std::fclose(f);
  if (!ok && err) *err = "short write";
  return ok ? 0 : 2;
}

int bf_params_load_image(const char *path, BfParams *out, std::string *err) {
  FILE *f = std::fopen(path, "rb");
  if (!f) { if (err) *err = std::string("cannot open ") + path; return 1; }
  uint32_t hdr[2] = {0, 0};
  bool ok = std::fread(hdr, sizeof hdr, 1, f) == 1 && hdr[0] == kImageMagic && hdr[1] == sizeof(BfParams) &&
            std::fread(out, sizeof(BfParams), 1, f) == 1;
  std::fclose(f);
  if (!ok && err) *err = std::string("not a parameter image of this build: ") + path;
  return ok ? 0 : 2;
}
