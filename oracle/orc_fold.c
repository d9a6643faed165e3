/*
 * orc_fold.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See orc_fold.h.
 *
 * Restates, in plain C, the ViennaRNA 2.6.4 model that DesiRNA's hot path calls:
 *   RNA.params_load            DesiRNA.py:455-456          -> orc_params_load
 *   fc.eval_structure          utils/energy_scores.py:75,99 -> orc_eval
 *   fc.mfe / fc.mfe_dimer      utils/energy_scores.py:151,156; sequence_utils.py:1183 -> orc_mfe
 *   fc.pf / fc.pf_dimer        utils/energy_scores.py:150,157; dimer_multichain_energy.py:47 -> orc_pf
 *   fc.ensemble_defect         utils/energy_scores.py:374   -> orc_ensemble_defect
 * Model details follow SURVEY.md Appendix A (validated there against the reference's
 * shipped trajectories): dangles=2, TURN=3, MAXLOOP=30, special hairpins, pf_smooth=1,
 * T=37C, kT = 310.15*1.98717 cal/mol.  Energies are int dcal/mol, INF = 10^7.
 *
 * Parity status: pinned by the tests/golden/ jsonl fixtures (Turner 1999).  bpp / ensemble defect
 * are pinned only by exhaustive enumeration under the same (validated) loop model.
 */
#include "orc_fold.h"
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define INF 10000000
#define TURN 3
#define MAXLOOP 30
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

struct orc_params {
  int stack[8][8];
  int mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5];
  int mmM[8][5][5], mmE[8][5][5];         /* clamped to <= 0 (dangles != 0) */
  int mmM_raw[8][5][5], mmE_raw[8][5][5]; /* file values, feed the smoothed PF weights */
  int dangle5[8][5], dangle3[8][5], d5_raw[8][5], d3_raw[8][5];
  int int11[8][8][5][5];
  int int21[8][8][5][5][5];
  int int22[8][8][5][5][5][5];
  int hairpin[31], bulge[31], interior[31];
  int ninio_m, ninio_max, MLbase, MLclosing, MLintern, DuplexInit, TerminalAU;
  double lxc;
  int n_tri, n_tetra, n_hexa;
  char tri[64][8], tetra[128][8], hexa[64][12];
  int tri_e[64], tetra_e[128], hexa_e[64];
  double kT; /* cal/mol */
  /* Boltzmann weights */
  double x_mmM[8][5][5], x_mmE[8][5][5], x_d5[8][5], x_d3[8][5];
  /* plain exp(-E*10/kT) twins of the integer tables (as ViennaRNA's vrna_exp_param_t) */
  double x_stack[8][8], x_mmI[8][5][5], x_mm1nI[8][5][5], x_mm23I[8][5][5];
  double x_int11[8][8][5][5], x_int21[8][8][5][5][5], x_int22[8][8][5][5][5][5];
  double x_bulge[31], x_interior[31], x_ninio[31], x_TerminalAU;
};

static const int RTYPE[8] = {0, 2, 1, 4, 3, 6, 5, 7};
/* pair type of (a,b), bases N=0 A=1 C=2 G=3 U=4: CG=1 GC=2 GU=3 UG=4 AU=5 UA=6 */
static const int PTYPE[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};

static int enc(char c) {
  switch (toupper((unsigned char)c)) {
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'U': case 'T': return 4;
    default: return 0;
  }
}

/* ------------------------------------------------------------------ parameter file */
typedef struct { char **tok; int n, cap; } toklist;
static void tl_push(toklist *t, const char *s, int len) {
  if (t->n == t->cap) { t->cap = t->cap ? t->cap * 2 : 256; t->tok = (char **)realloc(t->tok, sizeof(char *) * t->cap); }
  char *c = (char *)malloc(len + 1); memcpy(c, s, len); c[len] = 0; t->tok[t->n++] = c;
}
static void tl_free(toklist *t) { for (int i = 0; i < t->n; i++) free(t->tok[i]); free(t->tok); }
static int tok_int(const char *s) {
  if (!strcmp(s, "INF")) return INF;
  if (!strcmp(s, "DEF")) return -50;
  return atoi(s);
}

static double smooth_w(int G, double kT) {
  /* ViennaRNA SMOOTH(-G) with SCALE 10 (pf_smooth=1), SURVEY A.6 */
  double x = -(double)G, xs = x / 10.0, y;
  if (xs < -1.2283697) y = 0.0;
  else if (xs > 0.8660254) y = x;
  else { double s = sin(xs - 0.34242663) + 1.0; y = 10.0 * 0.38490018 * s * s; }
  return exp(y * 10.0 / kT);
}

orc_params *orc_params_load(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  char *buf = (char *)malloc(sz + 2);
  if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); return NULL; }
  fclose(f); buf[sz] = '\n'; buf[sz + 1] = 0;
  /* blank out comments */
  for (char *p = buf; (p = strstr(p, "/*"));) {
    char *e = strstr(p, "*/");
    if (!e) e = buf + sz - 2;
    for (char *q = p; q < e + 2; q++) if (*q != '\n') *q = ' ';
    p = e + 2;
  }
  orc_params *P = (orc_params *)calloc(1, sizeof(orc_params));
  P->kT = (37.0 + 273.15) * 1.98717;
  char *line = buf;
  char sec[64] = "";
  toklist T = {0};
  int ok = 1;
  /* flush helper is inlined via goto-free loop: we collect tokens of a section, then assign */
  for (;;) {
    char *nl = line ? strchr(line, '\n') : NULL;
    int at_end = (nl == NULL);
    int is_hdr = 0;
    char newsec[64] = "";
    if (!at_end) {
      *nl = 0;
      char *s = line; while (*s == ' ' || *s == '\t' || *s == '\r') s++;
      if (s[0] == '#' && s[1] != '#') { is_hdr = 1; sscanf(s + 1, " %63s", newsec); }
      else if (s[0] == '#') { /* file header */ }
      else {
        char *q = s;
        while (*q) {
          while (*q && isspace((unsigned char)*q)) q++;
          char *st = q;
          while (*q && !isspace((unsigned char)*q)) q++;
          if (q > st) tl_push(&T, st, (int)(q - st));
        }
      }
    }
    if (is_hdr || at_end) {
      /* assign collected tokens to the finished section */
      int k = 0;
#define NEXT() (k < T.n ? tok_int(T.tok[k++]) : (ok = 0, 0))
      if (!strcmp(sec, "stack")) { for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) P->stack[a][b] = NEXT(); }
      else if (!strcmp(sec, "mismatch_hairpin") || !strcmp(sec, "mismatch_interior") || !strcmp(sec, "mismatch_interior_1n") ||
               !strcmp(sec, "mismatch_interior_23") || !strcmp(sec, "mismatch_multi") || !strcmp(sec, "mismatch_exterior")) {
        int(*dst)[5][5] = !strcmp(sec, "mismatch_hairpin") ? P->mmH : !strcmp(sec, "mismatch_interior") ? P->mmI :
                          !strcmp(sec, "mismatch_interior_1n") ? P->mm1nI : !strcmp(sec, "mismatch_interior_23") ? P->mm23I :
                          !strcmp(sec, "mismatch_multi") ? P->mmM_raw : P->mmE_raw;
        for (int a = 1; a <= 7; a++) for (int b = 0; b < 5; b++) for (int c = 0; c < 5; c++) dst[a][b][c] = NEXT();
      }
      else if (!strcmp(sec, "dangle5")) { for (int a = 1; a <= 7; a++) for (int b = 0; b < 5; b++) P->d5_raw[a][b] = NEXT(); }
      else if (!strcmp(sec, "dangle3")) { for (int a = 1; a <= 7; a++) for (int b = 0; b < 5; b++) P->d3_raw[a][b] = NEXT(); }
      else if (!strcmp(sec, "int11")) { for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) P->int11[a][b][c][d] = NEXT(); }
      else if (!strcmp(sec, "int21")) { for (int a = 1; a <= 7; a++) for (int b = 1; b <= 7; b++) for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) for (int e = 0; e < 5; e++) P->int21[a][b][c][d][e] = NEXT(); }
      else if (!strcmp(sec, "int22")) { for (int a = 1; a <= 6; a++) for (int b = 1; b <= 6; b++) for (int c = 1; c < 5; c++) for (int d = 1; d < 5; d++) for (int e = 1; e < 5; e++) for (int g = 1; g < 5; g++) P->int22[a][b][c][d][e][g] = NEXT(); }
      else if (!strcmp(sec, "hairpin")) { for (int a = 0; a <= 30; a++) P->hairpin[a] = NEXT(); }
      else if (!strcmp(sec, "bulge")) { for (int a = 0; a <= 30; a++) P->bulge[a] = NEXT(); }
      else if (!strcmp(sec, "interior")) { for (int a = 0; a <= 30; a++) P->interior[a] = NEXT(); }
      else if (!strcmp(sec, "NINIO")) { P->ninio_m = NEXT(); (void)NEXT(); P->ninio_max = NEXT(); }
      else if (!strcmp(sec, "ML_params")) { P->MLbase = NEXT(); (void)NEXT(); P->MLclosing = NEXT(); (void)NEXT(); P->MLintern = NEXT(); (void)NEXT(); }
      else if (!strcmp(sec, "Misc")) {
        P->DuplexInit = NEXT(); (void)NEXT(); P->TerminalAU = NEXT(); (void)NEXT();
        if (k < T.n) P->lxc = atof(T.tok[k++]); else ok = 0;
      }
      else if (!strcmp(sec, "Triloops") || !strcmp(sec, "Tetraloops") || !strcmp(sec, "Hexaloops")) {
        for (; k + 3 <= T.n; k += 3) {
          const char *s = T.tok[k]; int e = tok_int(T.tok[k + 1]);
          if (!strcmp(sec, "Triloops") && P->n_tri < 64) { strncpy(P->tri[P->n_tri], s, 7); P->tri_e[P->n_tri++] = e; }
          if (!strcmp(sec, "Tetraloops") && P->n_tetra < 128) { strncpy(P->tetra[P->n_tetra], s, 7); P->tetra_e[P->n_tetra++] = e; }
          if (!strcmp(sec, "Hexaloops") && P->n_hexa < 64) { strncpy(P->hexa[P->n_hexa], s, 11); P->hexa_e[P->n_hexa++] = e; }
        }
      }
#undef NEXT
      for (int i = 0; i < T.n; i++) free(T.tok[i]);
      T.n = 0;
      strcpy(sec, newsec);
    }
    if (at_end) break;
    line = nl + 1;
  }
  tl_free(&T);
  free(buf);
  if (!ok || P->hairpin[3] == 0) { free(P); return NULL; }
  /* derived tables: clamp dangles / multi / exterior mismatches at 0 */
  for (int a = 0; a < 8; a++) for (int b = 0; b < 5; b++) {
    P->dangle5[a][b] = P->d5_raw[a][b] > 0 ? 0 : P->d5_raw[a][b];
    P->dangle3[a][b] = P->d3_raw[a][b] > 0 ? 0 : P->d3_raw[a][b];
    P->x_d5[a][b] = smooth_w(P->d5_raw[a][b], P->kT);
    P->x_d3[a][b] = smooth_w(P->d3_raw[a][b], P->kT);
    for (int c = 0; c < 5; c++) {
      P->mmM[a][b][c] = P->mmM_raw[a][b][c] > 0 ? 0 : P->mmM_raw[a][b][c];
      P->mmE[a][b][c] = P->mmE_raw[a][b][c] > 0 ? 0 : P->mmE_raw[a][b][c];
      P->x_mmM[a][b][c] = smooth_w(P->mmM_raw[a][b][c], P->kT);
      P->x_mmE[a][b][c] = smooth_w(P->mmE_raw[a][b][c], P->kT);
    }
  }
  /* int22 rows for the non-standard pair type 7 = max over the standard partners (A.2, recalled) */
  for (int c = 1; c < 5; c++) for (int d = 1; d < 5; d++) for (int e = 1; e < 5; e++) for (int g = 1; g < 5; g++) {
    int mall = -INF;
    for (int a = 1; a <= 6; a++) {
      int m1 = -INF, m2 = -INF;
      for (int b = 1; b <= 6; b++) { m1 = MAX2(m1, P->int22[a][b][c][d][e][g]); m2 = MAX2(m2, P->int22[b][a][c][d][e][g]); }
      P->int22[a][7][c][d][e][g] = m1; P->int22[7][a][c][d][e][g] = m2; mall = MAX2(mall, m1);
    }
    P->int22[7][7][c][d][e][g] = mall;
  }
  {
    const double kT = P->kT;
#define BZ(e) exp(-(double)(e) * 10.0 / kT)
    for (int a = 0; a < 8; a++) for (int b = 0; b < 8; b++) {
      P->x_stack[a][b] = BZ(P->stack[a][b]);
      for (int c = 0; c < 5; c++) for (int d = 0; d < 5; d++) {
        P->x_int11[a][b][c][d] = BZ(P->int11[a][b][c][d]);
        for (int e = 0; e < 5; e++) {
          P->x_int21[a][b][c][d][e] = BZ(P->int21[a][b][c][d][e]);
          for (int g = 0; g < 5; g++) P->x_int22[a][b][c][d][e][g] = BZ(P->int22[a][b][c][d][e][g]);
        }
      }
    }
    for (int a = 0; a < 8; a++) for (int b = 0; b < 5; b++) for (int c = 0; c < 5; c++) {
      P->x_mmI[a][b][c] = BZ(P->mmI[a][b][c]); P->x_mm1nI[a][b][c] = BZ(P->mm1nI[a][b][c]); P->x_mm23I[a][b][c] = BZ(P->mm23I[a][b][c]);
    }
    for (int u = 0; u <= 30; u++) {
      P->x_bulge[u] = BZ(P->bulge[u]); P->x_interior[u] = BZ(P->interior[u]);
      P->x_ninio[u] = BZ(MIN2(P->ninio_max, u * P->ninio_m));
    }
    P->x_TerminalAU = BZ(P->TerminalAU);
#undef BZ
  }
  return P;
}
void orc_params_free(orc_params *P) { free(P); }

int orc_params_get(const orc_params *P, const char *name, int i0, int i1, int i2, int i3, int i4, int i5) {
  if (!strcmp(name, "stack")) return P->stack[i0][i1];
  if (!strcmp(name, "mmH")) return P->mmH[i0][i1][i2];
  if (!strcmp(name, "mmI")) return P->mmI[i0][i1][i2];
  if (!strcmp(name, "mm1nI")) return P->mm1nI[i0][i1][i2];
  if (!strcmp(name, "mm23I")) return P->mm23I[i0][i1][i2];
  if (!strcmp(name, "mmM")) return P->mmM[i0][i1][i2];
  if (!strcmp(name, "mmE")) return P->mmE[i0][i1][i2];
  if (!strcmp(name, "mmM_raw")) return P->mmM_raw[i0][i1][i2];
  if (!strcmp(name, "mmE_raw")) return P->mmE_raw[i0][i1][i2];
  if (!strcmp(name, "dangle5")) return P->dangle5[i0][i1];
  if (!strcmp(name, "dangle3")) return P->dangle3[i0][i1];
  if (!strcmp(name, "d5_raw")) return P->d5_raw[i0][i1];
  if (!strcmp(name, "d3_raw")) return P->d3_raw[i0][i1];
  if (!strcmp(name, "int11")) return P->int11[i0][i1][i2][i3];
  if (!strcmp(name, "int21")) return P->int21[i0][i1][i2][i3][i4];
  if (!strcmp(name, "int22")) return P->int22[i0][i1][i2][i3][i4][i5];
  if (!strcmp(name, "hairpin")) return P->hairpin[i0];
  if (!strcmp(name, "bulge")) return P->bulge[i0];
  if (!strcmp(name, "interior")) return P->interior[i0];
  if (!strcmp(name, "ninio_m")) return P->ninio_m;
  if (!strcmp(name, "ninio_max")) return P->ninio_max;
  if (!strcmp(name, "MLbase")) return P->MLbase;
  if (!strcmp(name, "MLclosing")) return P->MLclosing;
  if (!strcmp(name, "MLintern")) return P->MLintern;
  if (!strcmp(name, "DuplexInit")) return P->DuplexInit;
  if (!strcmp(name, "TerminalAU")) return P->TerminalAU;
  if (!strcmp(name, "lxc1000")) return (int)lround(P->lxc * 1000.0);
  if (!strcmp(name, "n_tetra")) return P->n_tetra;
  if (!strcmp(name, "n_tri")) return P->n_tri;
  if (!strcmp(name, "n_hexa")) return P->n_hexa;
  if (!strcmp(name, "tetra_e")) return P->tetra_e[i0];
  return -INF;
}

/* ------------------------------------------------------------------ loop energies (A.3) */
static int loop_ext(int tab30, double lxc, int u) { return tab30 + (int)(lxc * log((double)u / 30.0)); }

/* seq0: 0-based ASCII (upper-cased copy), i,j 1-based */
static int special_hairpin(const orc_params *P, const char *seqU, int i, int j, int *found) {
  int u = j - i - 1;
  *found = 0;
  if (u == 4) { for (int k = 0; k < P->n_tetra; k++) if (!strncmp(P->tetra[k], seqU + i - 1, 6)) { *found = 1; return P->tetra_e[k]; } }
  else if (u == 6) { for (int k = 0; k < P->n_hexa; k++) if (!strncmp(P->hexa[k], seqU + i - 1, 8)) { *found = 1; return P->hexa_e[k]; } }
  else if (u == 3) { for (int k = 0; k < P->n_tri; k++) if (!strncmp(P->tri[k], seqU + i - 1, 5)) { *found = 1; return P->tri_e[k]; } }
  return 0;
}

static int E_hairpin(const orc_params *P, int u, int t, int si1, int sj1, const char *seqU, int i, int j) {
  int e = (u <= 30) ? P->hairpin[u] : loop_ext(P->hairpin[30], P->lxc, u);
  if (u < 3) return e;
  int found, es = special_hairpin(P, seqU, i, j, &found);
  if (found) return es;
  if (u == 3) return e + (t > 2 ? P->TerminalAU : 0);
  return e + P->mmH[t][si1][sj1];
}

static int E_intloop(const orc_params *P, int n1, int n2, int t, int t2, int si1, int sj1, int sp1, int sq1) {
  int nl = MAX2(n1, n2), ns = MIN2(n1, n2), e;
  if (nl == 0) return P->stack[t][t2];
  if (ns == 0) {
    e = (nl <= MAXLOOP) ? P->bulge[nl] : loop_ext(P->bulge[30], P->lxc, nl);
    if (nl == 1) e += P->stack[t][t2];
    else { if (t > 2) e += P->TerminalAU; if (t2 > 2) e += P->TerminalAU; }
    return e;
  }
  if (ns == 1) {
    if (nl == 1) return P->int11[t][t2][si1][sj1];
    if (nl == 2) return (n1 == 1) ? P->int21[t][t2][si1][sq1][sj1] : P->int21[t2][t][sq1][si1][sp1];
    e = (nl + 1 <= MAXLOOP) ? P->interior[nl + 1] : loop_ext(P->interior[30], P->lxc, nl + 1);
    e += MIN2(P->ninio_max, (nl - ns) * P->ninio_m);
    e += P->mm1nI[t][si1][sj1] + P->mm1nI[t2][sq1][sp1];
    return e;
  }
  if (ns == 2) {
    if (nl == 2) return P->int22[t][t2][si1][sp1][sq1][sj1];
    if (nl == 3) return P->interior[5] + P->ninio_m + P->mm23I[t][si1][sj1] + P->mm23I[t2][sq1][sp1];
  }
  {
    int u = nl + ns;
    e = (u <= MAXLOOP) ? P->interior[u] : loop_ext(P->interior[30], P->lxc, u);
    e += MIN2(P->ninio_max, (nl - ns) * P->ninio_m);
    e += P->mmI[t][si1][sj1] + P->mmI[t2][sq1][sp1];
    return e;
  }
}

/* a, b < 0 : neighbour absent */
static int E_mlstem(const orc_params *P, int t, int a, int b) {
  int e = 0;
  if (a >= 0 && b >= 0) e += P->mmM[t][a][b];
  else if (a >= 0) e += P->dangle5[t][a];
  else if (b >= 0) e += P->dangle3[t][b];
  if (t > 2) e += P->TerminalAU;
  return e + P->MLintern;
}
static int E_ext(const orc_params *P, int t, int a, int b) {
  int e = 0;
  if (a >= 0 && b >= 0) e += P->mmE[t][a][b];
  else if (a >= 0) e += P->dangle5[t][a];
  else if (b >= 0) e += P->dangle3[t][b];
  if (t > 2) e += P->TerminalAU;
  return e;
}
static double boltz(const orc_params *P, int e) { return exp(-(double)e * 10.0 / P->kT); }
static double X_mlstem(const orc_params *P, int t, int a, int b) {
  double w = 1.0;
  if (a >= 0 && b >= 0) w = P->x_mmM[t][a][b];
  else if (a >= 0) w = P->x_d5[t][a];
  else if (b >= 0) w = P->x_d3[t][b];
  if (t > 2) w *= boltz(P, P->TerminalAU);
  return w * boltz(P, P->MLintern);
}
static double X_ext(const orc_params *P, int t, int a, int b) {
  double w = 1.0;
  if (a >= 0 && b >= 0) w = P->x_mmE[t][a][b];
  else if (a >= 0) w = P->x_d5[t][a];
  else if (b >= 0) w = P->x_d3[t][b];
  if (t > 2) w *= boltz(P, P->TerminalAU);
  return w;
}
/* PF hairpin weight: u>30 keeps the un-truncated log term (A.6) */
static double X_hairpin(const orc_params *P, int u, int t, int si1, int sj1, const char *seqU, int i, int j) {
  if (u <= 30) return boltz(P, E_hairpin(P, u, t, si1, sj1, seqU, i, j));
  double w = boltz(P, P->hairpin[30]) * exp(-(P->lxc * log((double)u / 30.0)) * 10.0 / P->kT);
  return w * boltz(P, P->mmH[t][si1][sj1]);
}

/* ------------------------------------------------------------------ fold context */
typedef struct {
  const orc_params *P;
  int n, cp;     /* cp = first index of strand B (n+1 when single strand) */
  int *S;        /* 0..n+1 */
  char *seqU;    /* 0-based upper-case with T->U */
  const unsigned char *nopair;
  unsigned char *ptm; /* (n+2)^2 pair-type matrix honoured by the recursions */
} ctx_t;

static void ctx_make_ptm(ctx_t *X);
static void ctx_init(ctx_t *X, const orc_params *P, const char *seq, int n, int cut, const unsigned char *nopair) {
  X->P = P; X->n = n; X->cp = (cut > 1 && cut <= n) ? cut : n + 1; X->nopair = nopair;
  X->S = (int *)calloc(n + 2, sizeof(int));
  X->seqU = (char *)calloc(n + 16, 1);
  for (int i = 1; i <= n; i++) { X->S[i] = enc(seq[i - 1]); X->seqU[i - 1] = "NACGU"[X->S[i]]; }
  ctx_make_ptm(X);
}
static void ctx_free(ctx_t *X) { free(X->S); free(X->seqU); free(X->ptm); }
static inline int same(const ctx_t *X, int a, int b) { return (a >= X->cp) == (b >= X->cp); }
/* pair type honoured by the folding recursions (0 = may not pair) */
static inline int ptype_slow(const ctx_t *X, int i, int j) {
  int t = PTYPE[X->S[i]][X->S[j]];
  if (!t) return 0;
  if (same(X, i, j) && j - i <= TURN) return 0;
  if (X->nopair && (X->nopair[i - 1] || X->nopair[j - 1])) return 0;
  return t;
}
static void ctx_make_ptm(ctx_t *X) {
  int W = X->n + 2;
  X->ptm = (unsigned char *)calloc((size_t)W * W, 1);
  for (int i = 1; i <= X->n; i++) for (int j = i + 1; j <= X->n; j++) X->ptm[(size_t)i * W + j] = (unsigned char)ptype_slow(X, i, j);
}
static inline int ptype(const ctx_t *X, int i, int j) { return X->ptm[(size_t)i * (X->n + 2) + j]; }
/* exterior-stem energy with strand-aware neighbours */
static inline int ext_stem(const ctx_t *X, int i, int j, int t) {
  int a = (i > 1 && same(X, i - 1, i)) ? X->S[i - 1] : -1;
  int b = (j < X->n && same(X, j, j + 1)) ? X->S[j + 1] : -1;
  return E_ext(X->P, t, a, b);
}
static inline double x_ext_stem(const ctx_t *X, int i, int j, int t) {
  int a = (i > 1 && same(X, i - 1, i)) ? X->S[i - 1] : -1;
  int b = (j < X->n && same(X, j, j + 1)) ? X->S[j + 1] : -1;
  return X_ext(X->P, t, a, b);
}
/* closing pair (i,j) of a nick-containing loop seen from inside (A.7) */
static inline void nick_close_nb(const ctx_t *X, int i, int j, int *a, int *b) {
  *a = (j - 1 >= X->cp) ? X->S[j - 1] : -1;
  *b = (i + 1 < X->cp) ? X->S[i + 1] : -1;
}

/* ------------------------------------------------------------------ eval_structure (A.8) */
static int make_pt(const char *db, int n, int *pt) {
  int *st = (int *)malloc(sizeof(int) * (n + 1)), sp = 0;
  for (int i = 1; i <= n; i++) pt[i] = 0;
  for (int i = 1; i <= n; i++) {
    if (db[i - 1] == '(') st[sp++] = i;
    else if (db[i - 1] == ')') { if (!sp) { free(st); return -1; } int o = st[--sp]; pt[o] = i; pt[i] = o; }
  }
  free(st);
  return sp ? -1 : 0;
}

int orc_eval(const orc_params *P, const char *seq, int n, int cut, const char *db) {
  ctx_t X; ctx_init(&X, P, seq, n, cut, NULL);
  int *pt = (int *)calloc(n + 2, sizeof(int));
  int energy = 0;
  if (make_pt(db, n, pt)) { free(pt); ctx_free(&X); return INF; }
  int *S = X.S, cp = X.cp, connected = 0;
  /* exterior loop */
  for (int i = 1; i <= n; i++) {
    if (pt[i] > i) {
      int j = pt[i], t = PTYPE[S[i]][S[j]]; if (!t) t = 7;
      energy += ext_stem(&X, i, j, t);
      i = j;
    }
  }
  for (int i = 1; i <= n; i++) {
    int j = pt[i];
    if (j <= i) continue;
    if (i < cp && j >= cp) connected = 1;
    int t = PTYPE[S[i]][S[j]]; if (!t) t = 7;
    /* collect the loop closed by (i,j) */
    int nstem = 0, nick = 0, p1 = 0, q1 = 0, mlsum = 0, extsum = 0, unp = 0;
    int k = i + 1, prev = i;
    while (k < j) {
      if (pt[k] > k) {
        int p = k, q = pt[k], t2 = PTYPE[S[p]][S[q]]; if (!t2) t2 = 7;
        if (prev < cp && p >= cp) nick = 1;
        if (!nstem) { p1 = p; q1 = q; }
        nstem++;
        mlsum += E_mlstem(P, t2, S[p - 1], S[q + 1]);
        extsum += ext_stem(&X, p, q, t2);
        prev = q; k = q + 1;
      } else { unp++; k++; }
    }
    if (prev < cp && j >= cp) nick = 1;
    if (nick) {
      int a, b; nick_close_nb(&X, i, j, &a, &b);
      energy += E_ext(P, RTYPE[t], a, b) + extsum;
    } else if (nstem == 0) {
      energy += E_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X.seqU, i, j);
    } else if (nstem == 1) {
      int t2 = PTYPE[S[q1]][S[p1]]; if (!t2) t2 = 7;
      energy += E_intloop(P, p1 - i - 1, j - q1 - 1, t, t2, S[i + 1], S[j - 1], S[p1 - 1], S[q1 + 1]);
    } else {
      energy += P->MLclosing + E_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]) + mlsum + unp * P->MLbase;
    }
  }
  if (connected) energy += P->DuplexInit;
  free(pt); ctx_free(&X);
  return energy;
}

/* ------------------------------------------------------------------ MFE (A.4, A.5, A.7) */
typedef struct { int i, j, kind; } sector; /* kind 0 ext(f5 up to j), 1 ML, 2 pair, 3 fcA from i, 4 fcB up to j */

/* ---- backtrack (A.5): shared by orc_mfe and the tuned fill below; tables row-major (n+2)^2 ---- */
static void mfe_backtrack(const ctx_t *Xp, const int *c, const int *fML, const int *f5, const int *fcA, const int *fcB, char *ss_out) {
  const ctx_t X = *Xp;
  const orc_params *P = X.P;
  const int n = X.n, W = n + 2, cp = X.cp;
  const int *S = X.S;
#define C(i, j) c[(i) * W + (j)]
#define M(i, j) fML[(i) * W + (j)]
  if (ss_out) {
    memset(ss_out, '.', n); ss_out[n] = 0;
    sector *st = (sector *)malloc(sizeof(sector) * (4 * n + 16)); int sp = 0;
    st[sp++] = (sector){1, n, 0};
    while (sp > 0) {
      sector s = st[--sp];
      int i = s.i, j = s.j;
      if (s.kind == 0) {
        while (j > 0 && f5[j] == f5[j - 1]) j--;
        if (j <= 1) continue;
        int found = 0;
        for (int u = j - 1; u >= 1 && !found; u--) {
          int t = ptype(&X, u, j);
          if (!t || C(u, j) >= INF) continue;
          if (f5[j] == f5[u - 1] + C(u, j) + ext_stem(&X, u, j, t) + (same(&X, u, j) ? 0 : P->DuplexInit)) {
            st[sp++] = (sector){1, u - 1, 0};
            st[sp++] = (sector){u, j, 2};
            found = 1;
          }
        }
        if (!found) { fprintf(stderr, "orc_mfe: backtrack failed in f5 at %d\n", j); break; }
        continue;
      }
      if (s.kind == 3) { /* fcA: segment i..cp-1 */
        while (i < cp && fcA[i] == fcA[i + 1]) i++;
        if (i >= cp) continue;
        int found = 0;
        for (int q = i + 1; q <= cp - 1 && !found; q++) {
          int t = ptype(&X, i, q);
          if (!t || C(i, q) >= INF) continue;
          if (fcA[i] == C(i, q) + ext_stem(&X, i, q, t) + fcA[q + 1]) {
            st[sp++] = (sector){q + 1, 0, 3};
            st[sp++] = (sector){i, q, 2};
            found = 1;
          }
        }
        if (!found) { fprintf(stderr, "orc_mfe: backtrack failed in fcA at %d\n", i); break; }
        continue;
      }
      if (s.kind == 4) { /* fcB: segment cp..j */
        while (j >= cp && fcB[j] == fcB[j - 1]) j--;
        if (j < cp) continue;
        int found = 0;
        for (int p = j - 1; p >= cp && !found; p--) {
          int t = ptype(&X, p, j);
          if (!t || C(p, j) >= INF) continue;
          if (fcB[j] == fcB[p - 1] + C(p, j) + ext_stem(&X, p, j, t)) {
            st[sp++] = (sector){0, p - 1, 4};
            st[sp++] = (sector){p, j, 2};
            found = 1;
          }
        }
        if (!found) { fprintf(stderr, "orc_mfe: backtrack failed in fcB at %d\n", j); break; }
        continue;
      }
      if (s.kind == 1) {
        while (j > i && same(&X, j - 1, j) && M(i, j) == M(i, j - 1) + P->MLbase) j--;
        while (i < j && same(&X, i, i + 1) && M(i, j) == M(i + 1, j) + P->MLbase) i++;
        int t = ptype(&X, i, j);
        if (t && C(i, j) < INF && same(&X, i - 1, i) && same(&X, j, j + 1) && i > 1 && j < n &&
            M(i, j) == C(i, j) + E_mlstem(P, t, S[i - 1], S[j + 1])) {
          /* fall through to pair */
        } else {
          int found = 0;
          for (int u = i + 1; u <= j && !found; u++) {
            if (!same(&X, u - 1, u)) continue;
            int l = M(i, u - 1), r = M(u, j);
            if (l < INF && r < INF && M(i, j) == l + r) {
              st[sp++] = (sector){i, u - 1, 1};
              st[sp++] = (sector){u, j, 1};
              found = 1;
            }
          }
          if (!found) { fprintf(stderr, "orc_mfe: backtrack failed in fML at %d,%d\n", i, j); break; }
          continue;
        }
      }
      /* pair (i,j) */
      for (;;) {
        ss_out[i - 1] = '('; ss_out[j - 1] = ')';
        int t = ptype(&X, i, j), cij = C(i, j);
        if (i < cp && j >= cp) {
          /* hairpin-like loop containing the nick with nothing else inside: tried first, like a hairpin */
          int a, b; nick_close_nb(&X, i, j, &a, &b);
          if (cij == E_ext(P, RTYPE[t], a, b)) break;
        } else if (cij == E_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X.seqU, i, j)) break;
        int found = 0, pmax = MIN2(j - 2, i + MAXLOOP + 1);
        for (int p = i + 1; p <= pmax && !found; p++) {
          if (!same(&X, i, p)) break;
          int minq = j - i + p - MAXLOOP - 2; if (minq < p + 1) minq = p + 1;
          for (int q = j - 1; q >= minq; q--) {
            if (!same(&X, q, j)) break;
            int t2 = ptype(&X, p, q);
            if (!t2 || C(p, q) >= INF) continue;
            if (cij == C(p, q) + E_intloop(P, p - i - 1, j - q - 1, t, RTYPE[t2], S[i + 1], S[j - 1], S[p - 1], S[q + 1])) {
              i = p; j = q; found = 1; break;
            }
          }
        }
        if (found) continue;
        /* multiloop */
        int en = cij - P->MLclosing - E_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]);
        for (int u = i + 2; u <= j - 1 && !found && same(&X, i, i + 1) && same(&X, j - 1, j); u++) {
          if (!same(&X, u - 1, u)) continue;
          int l = M(i + 1, u - 1), r = M(u, j - 1);
          if (l < INF && r < INF && en == l + r) {
            st[sp++] = (sector){i + 1, u - 1, 1};
            st[sp++] = (sector){u, j - 1, 1};
            found = 1;
          }
        }
        /* nick-containing loop with stems inside is tried last (order pinned by 5 tied G5 goldens) */
        if (!found && i < cp && j >= cp) {
          int a, b; nick_close_nb(&X, i, j, &a, &b);
          if (cij == E_ext(P, RTYPE[t], a, b) + fcA[i + 1] + fcB[j - 1]) {
            st[sp++] = (sector){i + 1, 0, 3};
            st[sp++] = (sector){0, j - 1, 4};
            found = 1;
          }
        }
        if (!found) fprintf(stderr, "orc_mfe: backtrack failed at pair %d,%d\n", i, j);
        break;
      }
    }
    free(st);
  }
#undef C
#undef M
}


int orc_mfe(const orc_params *P, const char *seq, int n, int cut, const unsigned char *nopair, char *ss_out, long long *counts) {
  ctx_t X; ctx_init(&X, P, seq, n, cut, nopair);
  const int W = n + 2, cp = X.cp; int *S = X.S;
  int *c = (int *)malloc(sizeof(int) * W * W), *fML = (int *)malloc(sizeof(int) * W * W);
  int *f5 = (int *)calloc(n + 2, sizeof(int)), *fcA = (int *)calloc(n + 3, sizeof(int)), *fcB = (int *)calloc(n + 3, sizeof(int));
  long long cnt[4] = {0, 0, 0, 0};
#define C(i, j) c[(i) * W + (j)]
#define M(i, j) fML[(i) * W + (j)]
  for (int k = 0; k < W * W; k++) { c[k] = INF; fML[k] = INF; }
  int fcB_done = 0;
  for (int i = n; i >= 1; i--) {
    if (!fcB_done && cp <= n && i < cp) {
      /* all rows >= cp are final: best exterior-style decomposition of cp..k */
      fcB[cp - 1] = 0;
      for (int k = cp; k <= n; k++) {
        int e = fcB[k - 1];
        for (int p = k - 1; p >= cp; p--) { int t = ptype(&X, p, k); if (t && C(p, k) < INF) e = MIN2(e, fcB[p - 1] + C(p, k) + ext_stem(&X, p, k, t)); }
        fcB[k] = e;
      }
      fcB_done = 1;
    }
    for (int j = i + 1; j <= n; j++) {
      int t = ptype(&X, i, j);
      int e = INF;
      if (t) {
        if (i < cp && j >= cp) {
          int a, b; nick_close_nb(&X, i, j, &a, &b);
          e = E_ext(P, RTYPE[t], a, b) + fcA[i + 1] + fcB[j - 1];
        } else {
          e = E_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X.seqU, i, j);
        }
        /* interior loops */
        int pmax = MIN2(j - 2, i + MAXLOOP + 1);
        for (int p = i + 1; p <= pmax; p++) {
          if (!same(&X, i, p)) break;
          int minq = j - i + p - MAXLOOP - 2; if (minq < p + 1) minq = p + 1;
          for (int q = j - 1; q >= minq; q--) {
            if (!same(&X, q, j)) break;
            int t2 = ptype(&X, p, q);
            if (!t2) continue;
            cnt[0]++;
            int cc = C(p, q);
            if (cc >= INF) continue;
            int en = cc + E_intloop(P, p - i - 1, j - q - 1, t, RTYPE[t2], S[i + 1], S[j - 1], S[p - 1], S[q + 1]);
            e = MIN2(e, en);
          }
        }
        /* multiloop */
        if (same(&X, i, i + 1) && same(&X, j - 1, j)) {
          int dec = INF;
          for (int u = i + 2; u <= j - 1; u++) {
            if (!same(&X, u - 1, u)) continue;
            int l = M(i + 1, u - 1), r = M(u, j - 1);
            if (u - 1 - (i + 1) >= 1 && j - 1 - u >= 1) cnt[1]++;
            if (l < INF && r < INF) dec = MIN2(dec, l + r);
          }
          if (dec < INF) e = MIN2(e, dec + P->MLclosing + E_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]));
        }
        if (e > INF) e = INF;
      }
      C(i, j) = e;
      /* fML */
      int m = INF;
      if (e < INF && same(&X, i - 1, i) && same(&X, j, j + 1) && i > 1 && j < n) m = e + E_mlstem(P, t, S[i - 1], S[j + 1]);
      if (same(&X, i, i + 1) && M(i + 1, j) < INF) m = MIN2(m, M(i + 1, j) + P->MLbase);
      if (same(&X, j - 1, j) && M(i, j - 1) < INF) m = MIN2(m, M(i, j - 1) + P->MLbase);
      for (int u = i + 1; u <= j; u++) {
        if (!same(&X, u - 1, u)) continue;
        int l = M(i, u - 1), r = M(u, j);
        cnt[2]++;
        if (l < INF && r < INF) m = MIN2(m, l + r);
      }
      M(i, j) = m;
    }
    if (i < cp && cp <= n) {
      /* row i final: best exterior-style decomposition of i..cp-1 */
      int e = fcA[i + 1]; /* fcA[cp] = 0 */
      for (int q = i + 1; q <= cp - 1; q++) { int t = ptype(&X, i, q); if (t && C(i, q) < INF) e = MIN2(e, C(i, q) + ext_stem(&X, i, q, t) + fcA[q + 1]); }
      fcA[i] = e;
    }
  }
  f5[0] = 0;
  for (int j = 1; j <= n; j++) {
    int e = f5[j - 1];
    for (int i = j - 1; i >= 1; i--) {
      int t = ptype(&X, i, j);
      if (!t) continue;
      cnt[3]++;
      if (C(i, j) >= INF) continue;
      int en = f5[i - 1] + C(i, j) + ext_stem(&X, i, j, t) + (same(&X, i, j) ? 0 : P->DuplexInit);
      e = MIN2(e, en);
    }
    f5[j] = e;
  }
  int mfe = f5[n];
  mfe_backtrack(&X, c, fML, f5, fcA, fcB, ss_out);
#undef C
#undef M
  if (counts) for (int k = 0; k < 4; k++) counts[k] = cnt[k];
  free(c); free(fML); free(f5); free(fcA); free(fcB); ctx_free(&X);
  return mfe;
}

/* ------------------------------------------------------------------ partition function (A.6, A.7, A.10) */
static double X_intloop(const orc_params *P, int n1, int n2, int t, int t2, int si1, int sj1, int sp1, int sq1) {
  /* product of tabulated weights, as ViennaRNA's exp_E_IntLoop; u1+u2 <= MAXLOOP guaranteed by the callers */
  int nl = MAX2(n1, n2), ns = MIN2(n1, n2);
  if (nl == 0) return P->x_stack[t][t2];
  if (ns == 0) {
    double w = P->x_bulge[nl];
    if (nl == 1) return w * P->x_stack[t][t2];
    if (t > 2) w *= P->x_TerminalAU;
    if (t2 > 2) w *= P->x_TerminalAU;
    return w;
  }
  if (ns == 1) {
    if (nl == 1) return P->x_int11[t][t2][si1][sj1];
    if (nl == 2) return (n1 == 1) ? P->x_int21[t][t2][si1][sq1][sj1] : P->x_int21[t2][t][sq1][si1][sp1];
    return P->x_interior[nl + 1] * P->x_ninio[nl - ns] * P->x_mm1nI[t][si1][sj1] * P->x_mm1nI[t2][sq1][sp1];
  }
  if (ns == 2) {
    if (nl == 2) return P->x_int22[t][t2][si1][sp1][sq1][sj1];
    if (nl == 3) return P->x_interior[5] * P->x_ninio[1] * P->x_mm23I[t][si1][sj1] * P->x_mm23I[t2][sq1][sp1];
  }
  return P->x_interior[nl + ns] * P->x_ninio[nl - ns] * P->x_mmI[t][si1][sj1] * P->x_mmI[t2][sq1][sp1];
}

double orc_pf(const orc_params *P, const char *seq, int n, int cut, double *out5, double *bpp) {
  ctx_t X; ctx_init(&X, P, seq, n, cut, NULL);
  const int W = n + 2, cp = X.cp; int *S = X.S;
  const double kT = P->kT;
  const double pf_scale = exp(185.0 / kT); /* ViennaRNA default estimate: -185 cal/mol per nt */
  double *scl = (double *)malloc(sizeof(double) * (n + 3));
  scl[0] = 1.0; for (int k = 1; k <= n + 2; k++) scl[k] = scl[k - 1] / pf_scale;
  double *qb = (double *)calloc((size_t)W * W, sizeof(double)), *qm = (double *)calloc((size_t)W * W, sizeof(double));
  double *qm1 = (double *)calloc((size_t)W * W, sizeof(double));
  double *q5 = (double *)calloc(n + 3, sizeof(double)), *qA = (double *)calloc(n + 3, sizeof(double)), *qB = (double *)calloc(n + 3, sizeof(double));
  const double xMLb = boltz(P, P->MLbase), xMLc = boltz(P, P->MLclosing);
  double *bu = (double *)malloc(sizeof(double) * (n + 3)); /* (B(MLbase)/scale)^k */
  bu[0] = 1.0; for (int k = 1; k <= n + 2; k++) bu[k] = bu[k - 1] * xMLb / pf_scale;
#define QB(i, j) qb[(size_t)(i) * W + (j)]
#define QM(i, j) qm[(size_t)(i) * W + (j)]
#define QM1(i, j) qm1[(size_t)(i) * W + (j)]
  int qB_done = 0;
  for (int k = 0; k <= n + 2; k++) { qA[k] = 0; qB[k] = 0; }
  if (cp <= n) qA[cp] = 1.0;
  for (int i = n; i >= 1; i--) {
    if (!qB_done && cp <= n && i < cp) {
      qB[cp - 1] = 1.0;
      for (int k = cp; k <= n; k++) {
        double s = qB[k - 1] * scl[1];
        for (int p = cp; p < k; p++) { int t = ptype(&X, p, k); if (t) s += qB[p - 1] * QB(p, k) * x_ext_stem(&X, p, k, t); }
        qB[k] = s;
      }
      qB_done = 1;
    }
    for (int j = i + 1; j <= n; j++) {
      int t = ptype(&X, i, j);
      double s = 0.0;
      if (t) {
        if (i < cp && j >= cp) {
          int a, b; nick_close_nb(&X, i, j, &a, &b);
          s = X_ext(P, RTYPE[t], a, b) * qA[i + 1] * qB[j - 1] * scl[2];
        } else {
          s = X_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X.seqU, i, j) * scl[j - i + 1];
        }
        int pmax = MIN2(j - 2, i + MAXLOOP + 1);
        for (int p = i + 1; p <= pmax; p++) {
          if (!same(&X, i, p)) break;
          int minq = j - i + p - MAXLOOP - 2; if (minq < p + 1) minq = p + 1;
          for (int q = j - 1; q >= minq; q--) {
            if (!same(&X, q, j)) break;
            int t2 = ptype(&X, p, q);
            if (!t2) continue;
            s += QB(p, q) * X_intloop(P, p - i - 1, j - q - 1, t, RTYPE[t2], S[i + 1], S[j - 1], S[p - 1], S[q + 1]) * scl[p - i + j - q];
          }
        }
        if (same(&X, i, i + 1) && same(&X, j - 1, j)) {
          double dec = 0.0;
          for (int u = i + 2; u <= j - 1; u++) { if (!same(&X, u - 1, u)) continue; dec += QM(i + 1, u - 1) * QM1(u, j - 1); }
          s += dec * xMLc * X_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]) * scl[2];
        }
      }
      QB(i, j) = s;
      /* qm1[i][j]: exactly one stem starting at i, unpaired tail */
      double m1 = 0.0;
      if (same(&X, j - 1, j)) m1 = QM1(i, j - 1) * bu[1];
      if (t && same(&X, i - 1, i) && same(&X, j, j + 1) && i > 1 && j < n) m1 += s * X_mlstem(P, t, S[i - 1], S[j + 1]);
      QM1(i, j) = m1;
      /* qm[i][j] */
      double m = 0.0;
      for (int u = i; u <= j; u++) {
        double left = 0.0;
        if (u == i) left = 1.0;
        else {
          if (same(&X, i, u)) left = bu[u - i];
          if (same(&X, u - 1, u)) left += QM(i, u - 1);
        }
        m += left * QM1(u, j);
      }
      QM(i, j) = m;
    }
    if (i < cp && cp <= n) {
      double s = qA[i + 1] * scl[1];
      for (int q = i + 1; q <= cp - 1; q++) { int t = ptype(&X, i, q); if (t) s += QB(i, q) * x_ext_stem(&X, i, q, t) * qA[q + 1]; }
      qA[i] = s;
    }
  }
  q5[0] = 1.0;
  for (int j = 1; j <= n; j++) {
    double s = q5[j - 1] * scl[1];
    for (int i = 1; i < j; i++) { int t = ptype(&X, i, j); if (t) s += q5[i - 1] * QB(i, j) * x_ext_stem(&X, i, j, t); }
    q5[j] = s;
  }
  const double lnscale = log(pf_scale);
  double F0 = -kT * (log(q5[n]) + n * lnscale) / 1000.0;
  double o[5] = {0, 0, 0, 0, F0};
  if (cp <= n) {
    int nA = cp - 1, nB = n - cp + 1;
    double QA = qA[1], QBv = qB[n];
    /* Q_full/scale^n - QA/scale^nA * QB/scale^nB */
    double QAB = (q5[n] - QA * QBv) * boltz(P, P->DuplexInit);
    if (nA == nB && !strncmp(X.seqU, X.seqU + nA, nA)) QAB *= 0.5;
    double QT = QA * QBv + QAB;
    o[0] = -kT * (log(QA) + nA * lnscale) / 1000.0;
    o[1] = -kT * (log(QBv) + nB * lnscale) / 1000.0;
    /* ViennaRNA tests the *scaled* QAB against 1e-17 */
    o[2] = (QAB > 1e-17) ? -kT * (log(QAB) + n * lnscale) / 1000.0 : 999.0;
    o[3] = -kT * (log(QT) + n * lnscale) / 1000.0;
  }
  if (out5) for (int k = 0; k < 5; k++) out5[k] = o[k];

  /* ---- outside / base-pair probabilities (single strand), McCaskill with row accumulators (A.10) ---- */
  if (bpp && cp > n) {
    double *q3 = (double *)calloc(n + 3, sizeof(double));
    q3[n + 1] = 1.0;
    for (int i = n; i >= 1; i--) {
      double s = q3[i + 1] * scl[1];
      for (int j = i + 1; j <= n; j++) { int t = ptype(&X, i, j); if (t) s += QB(i, j) * x_ext_stem(&X, i, j, t) * q3[j + 1]; }
      q3[i] = s;
    }
    double *O = (double *)calloc((size_t)W * W, sizeof(double));  /* outside weight / Z */
    double *M1 = (double *)calloc((size_t)W * W, sizeof(double)); /* sum_j O[i][j]*close(i,j)*(bu[j-l-1]+qm[l+1][j-1]) */
    double *M2 = (double *)calloc((size_t)W * W, sizeof(double)); /* sum_j O[i][j]*close(i,j)*qm[l+1][j-1] */
#define OO(i, j) O[(size_t)(i) * W + (j)]
    const double Z = q5[n];
    for (int d = n - 1; d >= TURN + 1; d--) {
      for (int k = 1; k + d <= n; k++) {
        int l = k + d, t = ptype(&X, k, l);
        if (!t) continue;
        double o_ = q5[k - 1] * x_ext_stem(&X, k, l, t) * q3[l + 1] / Z;
        /* enclosing interior loops */
        for (int i = k - 1; i >= 1 && k - i - 1 <= MAXLOOP; i--) {
          int n1 = k - i - 1;
          for (int j = l + 1; j <= n && n1 + (j - l - 1) <= MAXLOOP; j++) {
            int to = ptype(&X, i, j);
            if (!to || OO(i, j) == 0.0) continue;
            o_ += OO(i, j) * X_intloop(P, n1, j - l - 1, to, RTYPE[t], S[i + 1], S[j - 1], S[k - 1], S[l + 1]) * scl[n1 + j - l - 1 + 2];
          }
        }
        /* enclosing multiloops */
        double ml = 0.0;
        for (int i = 1; i < k; i++) {
          double a = M1[(size_t)i * W + l], b2 = M2[(size_t)i * W + l];
          if (a == 0.0 && b2 == 0.0) continue;
          if (k - i - 1 >= 1) ml += QM(i + 1, k - 1) * a;
          ml += bu[k - i - 1] * b2;
        }
        if (k > 1 && l < n) o_ += ml * X_mlstem(P, t, S[k - 1], S[l + 1]);
        OO(k, l) = o_;
        /* feed accumulators with (k,l) acting as a closing pair */
        if (o_ != 0.0) {
          double cl = o_ * xMLc * X_mlstem(P, RTYPE[t], S[l - 1], S[k + 1]) * scl[2];
          for (int x = k + 1; x <= l - 1; x++) {
            /* inner stem ends at x; right part x+1..l-1 */
            double qmr = (l - 1 >= x + 1) ? QM(x + 1, l - 1) : 0.0;
            M1[(size_t)k * W + x] += cl * (bu[l - x - 1] + qmr);
            M2[(size_t)k * W + x] += cl * qmr;
          }
        }
      }
    }
    for (int i = 0; i < n * n; i++) bpp[i] = 0.0;
    for (int i = 1; i <= n; i++) for (int j = i + 1; j <= n; j++) bpp[(size_t)(i - 1) * n + (j - 1)] = OO(i, j) * QB(i, j);
#undef OO
    free(q3); free(O); free(M1); free(M2);
  }
#undef QB
#undef QM
#undef QM1
  free(scl); free(bu); free(qb); free(qm); free(qm1); free(q5); free(qA); free(qB); ctx_free(&X);
  return F0;
}

double orc_ensemble_defect(const double *bpp, int n, const char *db) {
  int *pt = (int *)calloc(n + 2, sizeof(int));
  make_pt(db, n, pt);
  double ed = 0.0;
  for (int i = 1; i <= n; i++) {
    if (pt[i]) {
      int a = MIN2(i, pt[i]), b = MAX2(i, pt[i]);
      ed += 1.0 - bpp[(size_t)(a - 1) * n + (b - 1)];
    } else {
      double pp = 0.0;
      for (int j = 1; j <= n; j++) if (j != i) { int a = MIN2(i, j), b = MAX2(i, j); pp += bpp[(size_t)(a - 1) * n + (b - 1)]; }
      ed += pp;
    }
  }
  free(pt);
  return ed / n;
}

/* ------------------------------------------------------------------ exhaustive enumeration (tests) */
typedef struct {
  ctx_t *X; int n; char *db; double Z; double *bpp; int e1, e2; int *pt;
  int bound, cap, cnt; int *band_e; char *band_ss; /* optional: every structure with energy <= bound */
} enum_t;

static double struct_weight(enum_t *E, int *e_out) {
  /* weight under the PF model: integer loop energies except smoothed dangle/mismatch families */
  ctx_t *X = E->X; const orc_params *P = X->P; int n = E->n, *S = X->S, *pt = E->pt;
  double w = 1.0; int en = 0;
  for (int i = 1; i <= n; i++) if (pt[i] > i) { int j = pt[i], t = PTYPE[S[i]][S[j]]; w *= x_ext_stem(X, i, j, t); en += ext_stem(X, i, j, t); i = j; }
  for (int i = 1; i <= n; i++) {
    int j = pt[i]; if (j <= i) continue;
    int t = PTYPE[S[i]][S[j]], nstem = 0, p1 = 0, q1 = 0, unp = 0; double xml = 1.0; int eml = 0;
    for (int k = i + 1; k < j;) {
      if (pt[k] > k) { int p = k, q = pt[k], t2 = PTYPE[S[p]][S[q]]; if (!nstem) { p1 = p; q1 = q; } nstem++; xml *= X_mlstem(P, t2, S[p - 1], S[q + 1]); eml += E_mlstem(P, t2, S[p - 1], S[q + 1]); k = q + 1; }
      else { unp++; k++; }
    }
    if (nstem == 0) { w *= X_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X->seqU, i, j); en += E_hairpin(P, j - i - 1, t, S[i + 1], S[j - 1], X->seqU, i, j); }
    else if (nstem == 1) { int t2 = PTYPE[S[q1]][S[p1]]; int e = E_intloop(P, p1 - i - 1, j - q1 - 1, t, t2, S[i + 1], S[j - 1], S[p1 - 1], S[q1 + 1]); w *= boltz(P, e); en += e; }
    else { w *= boltz(P, P->MLclosing + unp * P->MLbase) * X_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]) * xml; en += P->MLclosing + unp * P->MLbase + E_mlstem(P, RTYPE[t], S[j - 1], S[i + 1]) + eml; }
  }
  *e_out = en;
  return w;
}

static void enum_rec(enum_t *E, int pos) {
  int n = E->n;
  while (pos <= n && E->pt[pos] != 0) pos++;
  if (pos > n) {
    int en; double w = struct_weight(E, &en);
    E->Z += w;
    if (E->bpp) for (int i = 1; i <= n; i++) if (E->pt[i] > i) E->bpp[(size_t)(i - 1) * n + (E->pt[i] - 1)] += w;
    if (en < E->e1) { E->e2 = E->e1; E->e1 = en; } else if (en > E->e1 && en < E->e2) E->e2 = en;
    if (E->band_e && en <= E->bound) {
      if (E->cnt < E->cap) {
        E->band_e[E->cnt] = en;
        char *o = E->band_ss + (size_t)E->cnt * (n + 1);
        for (int i = 1; i <= n; i++) o[i - 1] = E->pt[i] > i ? '(' : (E->pt[i] > 0 ? ')' : '.');
        o[n] = 0;
      }
      E->cnt++;
    }
    return;
  }
  /* pos unpaired (mark with -1) */
  E->pt[pos] = -1; enum_rec(E, pos + 1); E->pt[pos] = 0;
  for (int j = pos + TURN + 1; j <= n; j++) {
    if (E->pt[j] != 0 || !ptype(E->X, pos, j)) continue;
    /* non-crossing, interior loop size <= MAXLOOP is NOT enforced here on purpose only for n<=MAXLOOP+? (n small) */
    int ok = 1;
    for (int k = pos + 1; k < j && ok; k++) if (E->pt[k] > 0 && (E->pt[k] > j)) ok = 0;
    if (!ok) continue;
    /* all of pos+1..j-1 currently undecided or decided-inside? positions < pos are decided; ensure none pairs into (pos,j) from outside */
    for (int k = 1; k < pos && ok; k++) if (E->pt[k] > pos && E->pt[k] < j) ok = 0;
    if (!ok) continue;
    E->pt[pos] = j; E->pt[j] = pos; enum_rec(E, pos + 1); E->pt[pos] = 0; E->pt[j] = 0;
  }
}

double orc_enumerate(const orc_params *P, const char *seq, int n, double *bpp, int *emin, int *e2nd) {
  ctx_t X; ctx_init(&X, P, seq, n, 0, NULL);
  enum_t E; E.X = &X; E.n = n; E.Z = 0; E.bpp = bpp; E.e1 = INF; E.e2 = INF; E.pt = (int *)calloc(n + 2, sizeof(int));
  E.band_e = NULL; E.band_ss = NULL; E.cnt = 0; E.cap = 0; E.bound = 0;
  if (bpp) for (int i = 0; i < n * n; i++) bpp[i] = 0;
  enum_rec(&E, 1);
  /* unpaired markers are -1 inside recursion only */
  if (bpp) for (int i = 0; i < n * n; i++) bpp[i] /= E.Z;
  if (emin) *emin = E.e1;
  if (e2nd) *e2nd = E.e2;
  double F = -P->kT * log(E.Z) / 1000.0;
  free(E.pt); ctx_free(&X);
  return F;
}

/* every structure with energy <= bound (dcal), brute force; returns how many there are (may exceed cap) */
int orc_enumerate_band(const orc_params *P, const char *seq, int n, int bound, int cap, int *energies, char *ss) {
  ctx_t X; ctx_init(&X, P, seq, n, 0, NULL);
  enum_t E; E.X = &X; E.n = n; E.Z = 0; E.bpp = NULL; E.e1 = INF; E.e2 = INF; E.pt = (int *)calloc(n + 2, sizeof(int));
  E.band_e = energies; E.band_ss = ss; E.cnt = 0; E.cap = cap; E.bound = bound;
  enum_rec(&E, 1);
  free(E.pt); ctx_free(&X);
  return E.cnt;
}

/* ------------------------------------------------------------------ batch helper (CPU baseline) */
/* ======================================================================================================================
 * Tuned CPU arm (bench.py: cpu_baseline.kind = "port-tuned", --impl reference).  TEST / BENCH INFRASTRUCTURE like the rest of
 * this file.  Same recurrences, same tables and the same backtrack as orc_mfe / orc_pf for ONE strand without constraints,
 * organised the way a CPU likes them, so that the GPU/CPU ratio is quoted against a fair CPU number and not against the
 * clarity-first restatement above:
 *   - interior loops decomposed: the inner pair's mismatch / terminal-AU term is folded into three copies of c (generic, 1xn,
 *     bulge), the nine non-decomposable shapes are evaluated explicitly, everything else is  min_q  copy[p][q] + pen[u1][u2]
 *     over a contiguous q range -- a loop gcc vectorises (AVX2: eight candidates per instruction);
 *   - the split loops read fML[i][u-1] and a transposed copy fT[j][u] = fML[u][j]: both contiguous, vectorised;
 *   - pair types from the precomputed matrix, no function call per candidate.
 * Bit-identical c / fML / f5 (a minimum does not depend on the order of its candidates), hence identical structures; the
 * partition function agrees to rounding (sums are re-associated).  tests/test_oracle_golden.py checks both.
 * ====================================================================================================================== */
static int mfe_fill_fast(const ctx_t *X, int *c, int *fML, int *f5) {
  const orc_params *P = X->P;
  const int n = X->n, W = n + 2;
  const int *S = X->S;
  const size_t WW = (size_t)W * W;
  int *fT = (int *)malloc(sizeof(int) * WW);                       /* fT[j][u] = fML[u][j] */
  int *cg = (int *)malloc(sizeof(int) * WW), *c1 = (int *)malloc(sizeof(int) * WW), *cb = (int *)malloc(sizeof(int) * WW);
  int *c1T = (int *)malloc(sizeof(int) * WW), *cbT = (int *)malloc(sizeof(int) * WW);   /* [q][p] copies for the families with q fixed */
  for (size_t k = 0; k < WW; k++) { c[k] = INF; fML[k] = INF; fT[k] = INF; cg[k] = INF; c1[k] = INF; cb[k] = INF; c1T[k] = INF; cbT[k] = INF; }
  int PG[31][32], pen1[32];
  for (int a = 0; a < 31; a++) for (int b = 0; b < 32; b++) PG[a][b] = INF;
  for (int u1 = 2; u1 <= 28; u1++) for (int u2 = 2; u1 + u2 <= MAXLOOP; u2++) PG[u1][u2] = P->interior[u1 + u2] + MIN2(P->ninio_max, abs(u1 - u2) * P->ninio_m);
  PG[2][2] = PG[2][3] = PG[3][2] = INF;   /* 2x2, 2x3, 3x2 have their own tables: among the nine explicit shapes */
  for (int sz = 0; sz < 32; sz++) pen1[sz] = (sz >= 4 && sz <= MAXLOOP) ? P->interior[sz] + MIN2(P->ninio_max, (sz - 2) * P->ninio_m) : INF;
  static const int SU1[9] = {0, 0, 1, 1, 1, 2, 2, 2, 3}, SU2[9] = {0, 1, 0, 1, 2, 1, 2, 3, 2};
  for (int i = n; i >= 1; i--) {
    for (int j = i + TURN + 1; j <= n; j++) {
      const int t = ptype(X, i, j);
      int e = INF;
      if (t) {
        const int si1 = S[i + 1], sj1 = S[j - 1], d = j - i;
        e = E_hairpin(P, d - 1, t, si1, sj1, X->seqU, i, j);
        for (int k = 0; k < 9; k++) {
          const int p = i + 1 + SU1[k], q = j - 1 - SU2[k];
          if (q <= p) continue;
          const int t2 = ptype(X, p, q);
          if (!t2) continue;
          const int cc = c[p * W + q];
          if (cc >= INF) continue;
          const int en = cc + E_intloop(P, SU1[k], SU2[k], t, RTYPE[t2], si1, sj1, S[p - 1], S[q + 1]);
          e = MIN2(e, en);
        }
        const int smax = MIN2(MAXLOOP, d - 6);   /* inner pair spans at least TURN + 1 */
        if (smax >= 2) {
          int ab = INF, a1 = INF, ag = INF;
          { /* bulges: (0, s) reads row i+1, (s, 0) reads column j-1 */
            const int *r1 = cb + (size_t)(i + 1) * W + (j - 1), *r2 = cbT + (size_t)(j - 1) * W + (i + 1);
            for (int sz = 2; sz <= smax; sz++) { const int v = MIN2(r1[-sz], r2[sz]) + P->bulge[sz]; ab = MIN2(ab, v); }
          }
          if (smax >= 4) { /* 1 x n: (1, s-1) reads row i+2, (s-1, 1) reads column j-2 */
            const int *r1 = c1 + (size_t)(i + 2) * W + j, *r2 = c1T + (size_t)(j - 2) * W + i;
            for (int sz = 4; sz <= smax; sz++) { const int v = MIN2(r1[-sz], r2[sz]) + pen1[sz]; a1 = MIN2(a1, v); }
          }
          for (int u1 = 2; u1 <= smax - 2; u1++) { /* generic: row p = i+1+u1, q = j-1-u2 */
            const int *row = cg + (size_t)(i + 1 + u1) * W + (j - 1);
            const int *pen = PG[u1];
            const int u2max = smax - u1;
            int m = INF;
            for (int u2 = 2; u2 <= u2max; u2++) { const int v = row[-u2] + pen[u2]; m = MIN2(m, v); }
            ag = MIN2(ag, m);
          }
          int v = MIN2(ag + P->mmI[t][si1][sj1], a1 + P->mm1nI[t][si1][sj1]);
          v = MIN2(v, ab + (t > 2 ? P->TerminalAU : 0));
          e = MIN2(e, v);
        }
        { /* multiloop closed by (i,j) */
          const int *l = fML + (size_t)(i + 1) * W, *r = fT + (size_t)(j - 1) * W;
          int dec = INF;
          for (int u = i + 2; u <= j - 1; u++) { const int v = l[u - 1] + r[u]; dec = MIN2(dec, v); }
          if (dec < INF / 2) e = MIN2(e, dec + P->MLclosing + E_mlstem(P, RTYPE[t], sj1, si1));
        }
        if (e >= INF / 2) e = INF;
      }
      c[i * W + j] = e;
      if (e < INF) {
        const int tr = RTYPE[t], x = S[j + 1], y = S[i - 1];
        const int vg = e + P->mmI[tr][x][y], v1 = e + P->mm1nI[tr][x][y], vb = e + (t > 2 ? P->TerminalAU : 0);
        cg[i * W + j] = vg; c1[i * W + j] = v1; cb[i * W + j] = vb; c1T[j * W + i] = v1; cbT[j * W + i] = vb;
      }
      int m = INF;
      if (e < INF && i > 1 && j < n) m = e + E_mlstem(P, t, S[i - 1], S[j + 1]);
      if (fML[(i + 1) * W + j] < INF) m = MIN2(m, fML[(i + 1) * W + j] + P->MLbase);
      if (fML[i * W + j - 1] < INF) m = MIN2(m, fML[i * W + j - 1] + P->MLbase);
      {
        const int *l = fML + (size_t)i * W, *r = fT + (size_t)j * W;
        int sp = INF;
        for (int u = i + 1; u <= j; u++) { const int v = l[u - 1] + r[u]; sp = MIN2(sp, v); }
        if (sp < INF / 2) m = MIN2(m, sp);
      }
      fML[i * W + j] = m;
      fT[j * W + i] = m;
    }
  }
  f5[0] = 0;
  for (int j = 1; j <= n; j++) {
    int e = f5[j - 1];
    for (int i = j - 1; i >= 1; i--) {
      const int t = ptype(X, i, j);
      if (!t || c[i * W + j] >= INF) continue;
      const int en = f5[i - 1] + c[i * W + j] + ext_stem(X, i, j, t);
      e = MIN2(e, en);
    }
    f5[j] = e;
  }
  free(fT); free(cg); free(c1); free(cb); free(c1T); free(cbT);
  return f5[n];
}

int orc_mfe_fast(const orc_params *P, const char *seq, int n, char *ss_out) {
  ctx_t X; ctx_init(&X, P, seq, n, 0, NULL);
  const int W = n + 2;
  int *c = (int *)malloc(sizeof(int) * W * W), *fML = (int *)malloc(sizeof(int) * W * W);
  int *f5 = (int *)calloc(n + 2, sizeof(int)), *fcA = (int *)calloc(n + 3, sizeof(int)), *fcB = (int *)calloc(n + 3, sizeof(int));
  const int mfe = mfe_fill_fast(&X, c, fML, f5);
  mfe_backtrack(&X, c, fML, f5, fcA, fcB, ss_out);
  free(c); free(fML); free(f5); free(fcA); free(fcB); ctx_free(&X);
  return mfe;
}

double orc_pf_fast(const orc_params *P, const char *seq, int n) {
  ctx_t Xs; ctx_init(&Xs, P, seq, n, 0, NULL);
  const ctx_t *X = &Xs;
  const int W = n + 2;
  const int *S = X->S;
  const size_t WW = (size_t)W * W;
  const double kT = P->kT, pf_scale = exp(185.0 / kT);
  double *scl = (double *)malloc(sizeof(double) * (n + 40)), *bu = (double *)malloc(sizeof(double) * (n + 40));
  scl[0] = 1.0; bu[0] = 1.0;
  const double xMLb = boltz(P, P->MLbase), xMLc = boltz(P, P->MLclosing), xtau = boltz(P, P->TerminalAU);
  for (int k = 1; k < n + 40; k++) { scl[k] = scl[k - 1] / pf_scale; bu[k] = bu[k - 1] * xMLb / pf_scale; }
  double *qb = (double *)calloc(WW, sizeof(double)), *qm = (double *)calloc(WW, sizeof(double)), *qm1 = (double *)calloc(WW, sizeof(double));
  double *q1T = (double *)calloc(WW, sizeof(double));   /* q1T[j][u] = qm1[u][j] */
  double *qg = (double *)calloc(WW, sizeof(double)), *q1 = (double *)calloc(WW, sizeof(double)), *qbb = (double *)calloc(WW, sizeof(double));
  double *q1nT = (double *)calloc(WW, sizeof(double)), *qbbT = (double *)calloc(WW, sizeof(double));
  double WG[31][32], w1[32], wb[32];
  for (int a = 0; a < 31; a++) for (int b = 0; b < 32; b++) WG[a][b] = 0.0;
  for (int u1 = 2; u1 <= 28; u1++) for (int u2 = 2; u1 + u2 <= MAXLOOP; u2++) WG[u1][u2] = P->x_interior[u1 + u2] * P->x_ninio[abs(u1 - u2)] * scl[u1 + u2 + 2];
  WG[2][2] = WG[2][3] = WG[3][2] = 0.0;   /* among the nine explicit shapes */
  for (int sz = 0; sz < 32; sz++) {
    w1[sz] = (sz >= 4 && sz <= MAXLOOP) ? P->x_interior[sz] * P->x_ninio[sz - 2] * scl[sz + 2] : 0.0;
    wb[sz] = (sz >= 2 && sz <= MAXLOOP) ? P->x_bulge[sz] * scl[sz + 2] : 0.0;
  }
  static const int SU1[9] = {0, 0, 1, 1, 1, 2, 2, 2, 3}, SU2[9] = {0, 1, 0, 1, 2, 1, 2, 3, 2};
  for (int i = n; i >= 1; i--) {
    for (int j = i + 1; j <= n; j++) {
      const int t = ptype(X, i, j);
      double s = 0.0;
      if (t) {
        const int si1 = S[i + 1], sj1 = S[j - 1], d = j - i;
        s = X_hairpin(P, d - 1, t, si1, sj1, X->seqU, i, j) * scl[d + 1];
        for (int k = 0; k < 9; k++) {
          const int p = i + 1 + SU1[k], q = j - 1 - SU2[k];
          if (q <= p) continue;
          const int t2 = ptype(X, p, q);
          if (!t2) continue;
          s += qb[(size_t)p * W + q] * X_intloop(P, SU1[k], SU2[k], t, RTYPE[t2], si1, sj1, S[p - 1], S[q + 1]) * scl[SU1[k] + SU2[k] + 2];
        }
        const int smax = MIN2(MAXLOOP, d - 6);
        if (smax >= 2) {
          double ab = 0.0, a1 = 0.0, ag = 0.0;
          {
            const double *r1 = qbb + (size_t)(i + 1) * W + (j - 1), *r2 = qbbT + (size_t)(j - 1) * W + (i + 1);
            for (int sz = 2; sz <= smax; sz++) ab += (r1[-sz] + r2[sz]) * wb[sz];
          }
          if (smax >= 4) {
            const double *r1 = q1 + (size_t)(i + 2) * W + j, *r2 = q1nT + (size_t)(j - 2) * W + i;
            for (int sz = 4; sz <= smax; sz++) a1 += (r1[-sz] + r2[sz]) * w1[sz];
          }
          for (int u1 = 2; u1 <= smax - 2; u1++) {
            const double *row = qg + (size_t)(i + 1 + u1) * W + (j - 1);
            const double *w = WG[u1];
            const int u2max = smax - u1;
            double m = 0.0;
            for (int u2 = 2; u2 <= u2max; u2++) m += row[-u2] * w[u2];
            ag += m;
          }
          s += ag * P->x_mmI[t][si1][sj1] + a1 * P->x_mm1nI[t][si1][sj1] + ab * (t > 2 ? xtau : 1.0);
        }
        {
          const double *l = qm + (size_t)(i + 1) * W, *r = q1T + (size_t)(j - 1) * W;
          double dec = 0.0;
          for (int u = i + 2; u <= j - 1; u++) dec += l[u - 1] * r[u];
          s += dec * xMLc * X_mlstem(P, RTYPE[t], sj1, si1) * scl[2];
        }
      }
      qb[(size_t)i * W + j] = s;
      if (s != 0.0) {
        const int tr = RTYPE[t], x = S[j + 1], y = S[i - 1];
        const double vg = s * P->x_mmI[tr][x][y], v1 = s * P->x_mm1nI[tr][x][y], vb = t > 2 ? s * xtau : s;
        qg[(size_t)i * W + j] = vg; q1[(size_t)i * W + j] = v1; qbb[(size_t)i * W + j] = vb;
        q1nT[(size_t)j * W + i] = v1; qbbT[(size_t)j * W + i] = vb;
      }
      double m1 = qm1[(size_t)i * W + j - 1] * bu[1];
      if (t && i > 1 && j < n) m1 += s * X_mlstem(P, t, S[i - 1], S[j + 1]);
      qm1[(size_t)i * W + j] = m1;
      q1T[(size_t)j * W + i] = m1;
      {
        /* qm[i][j] = qm1[i][j] + sum_{u>i} (bu[u-i] + qm[i][u-1]) qm1[u][j] */
        const double *l = qm + (size_t)i * W, *r = q1T + (size_t)j * W, *b = bu - i;
        double m = m1;
        for (int u = i + 1; u <= j; u++) m += (b[u] + l[u - 1]) * r[u];
        qm[(size_t)i * W + j] = m;
      }
    }
  }
  double *q5 = (double *)calloc(n + 3, sizeof(double));
  q5[0] = 1.0;
  for (int j = 1; j <= n; j++) {
    double s = q5[j - 1] * scl[1];
    for (int i = 1; i < j; i++) { const int t = ptype(X, i, j); if (t) s += q5[i - 1] * qb[(size_t)i * W + j] * x_ext_stem(X, i, j, t); }
    q5[j] = s;
  }
  const double F0 = -kT * (log(q5[n]) + n * log(pf_scale)) / 1000.0;
  free(scl); free(bu); free(qb); free(qm); free(qm1); free(q1T); free(qg); free(q1); free(qbb); free(q1nT); free(qbbT); free(q5);
  ctx_free(&Xs);
  return F0;
}


typedef struct {
  const orc_params *P; const char *seqs, *targets; int B, n; int *mfe; char *ss; double *epf; int *ed;
  volatile int *next;
  int fast;   /* 1: the tuned fill / partition function (single strand, no constraints) */
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *J = (batch_job *)arg;
  int n = J->n;
  char *o = (char *)malloc(n + 1), *tg = (char *)malloc(n + 1);
  for (;;) {
    int b = __sync_fetch_and_add(J->next, 1);
    if (b >= J->B) break;
    const char *s = J->seqs + (size_t)b * n;
    int e = J->fast ? orc_mfe_fast(J->P, s, n, o) : orc_mfe(J->P, s, n, 0, NULL, o, NULL);
    if (J->mfe) J->mfe[b] = e;
    if (J->ss) memcpy(J->ss + (size_t)b * (n + 1), o, n + 1);
    double f = J->fast ? orc_pf_fast(J->P, s, n) : orc_pf(J->P, s, n, 0, NULL, NULL);
    if (J->epf) J->epf[b] = f;
    if (J->ed) {
      if (J->targets) { memcpy(tg, J->targets + (size_t)b * n, n); tg[n] = 0; } else memcpy(tg, o, n + 1);
      J->ed[b] = orc_eval(J->P, s, n, 0, tg);
    }
  }
  free(o); free(tg);
  return NULL;
}

static int fold_batch_impl(const orc_params *P, const char *seqs, const char *targets, int B, int n, int nthreads,
                           int *mfe, char *ss, double *epf, int *ed, int fast) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  volatile int next = 0;
  batch_job J = {P, seqs, targets, B, n, mfe, ss, epf, ed, &next, fast};
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &J);
  batch_worker(&J);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  return 0;
}

int orc_fold_batch(const orc_params *P, const char *seqs, const char *targets, int B, int n, int nthreads,
                   int *mfe, char *ss, double *epf, int *ed) {
  return fold_batch_impl(P, seqs, targets, B, n, nthreads, mfe, ss, epf, ed, 0);
}
int orc_fold_batch_fast(const orc_params *P, const char *seqs, const char *targets, int B, int n, int nthreads,
                        int *mfe, char *ss, double *epf, int *ed) {
  return fold_batch_impl(P, seqs, targets, B, n, nthreads, mfe, ss, epf, ed, 1);
}
