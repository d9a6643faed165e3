// bf_params.h -- parameter image shared by the host loader and the sm_100a kernels.
//
// Replaces the process-global tables that RNA.params_load() fills in the reference
// (DesiRNA.py:455-456; file format: rna_turner1999.par, "RNAfold parameter file v2.0").
// The image is built once on the host (bf_params.cc), uploaded once to HBM and read by
// the kernels through the read-only path; the handful of scalars and the small loop
// tables used in every inner-loop iteration are additionally staged into shared memory
// by each CTA (see bf_kernels.cu, SmallTabs).
#pragma once
#include <stdint.h>

#define BF_INF 10000000
#define BF_MAXLOOP 30
#define BF_TURN 3
#define BF_NO_SPECIAL INT32_MIN
#define BF_EXT_TAB 4096  // (int)(lxc*ln(u/30)) table, u < BF_EXT_TAB
#define BF_MAX_HEXA 64

// Small tables + scalars touched in every inner-loop iteration: staged per CTA into shared memory.
struct BfSmallI {
  int32_t stack[8][8];
  int32_t mmH[8][5][5], mmI[8][5][5], mm1nI[8][5][5], mm23I[8][5][5];
  int32_t mmM[8][5][5], mmE[8][5][5];  // clamped at 0 (dangles=2)
  int32_t dangle5[8][5], dangle3[8][5];
  int32_t hairpin[31], bulge[31], interior[31], ninio[31];
  int32_t MLbase, MLclosing, MLintern, DuplexInit, TerminalAU, ninio_m, ninio_max, pad_;
};
struct BfSmallD {
  double x_stack[8][8];
  double x_mmH[8][5][5], x_mmI[8][5][5], x_mm1nI[8][5][5], x_mm23I[8][5][5];
  double x_mmM[8][5][5], x_mmE[8][5][5];  // pf_smooth weights of the un-clamped file values
  double x_d5[8][5], x_d3[8][5];
  double x_hairpin[31], x_bulge[31], x_interior[31], x_ninio[31];
  double x_MLbase, x_MLclosing, x_MLintern, x_DuplexInit, x_TerminalAU, kT;
};

struct BfParams {
  // ---- integer energies, dcal/mol.  pair types 1..7 (7 = non-standard, eval only), bases 0..4 (0 = N)
  BfSmallI si;
  int32_t int11[8][8][5][5];
  int32_t int21[8][8][5][5][5];
  int32_t int22[8][8][5][5][5][5];
  int32_t ext_log[BF_EXT_TAB];  // (int)(lxc*log(u/30.)) computed with the host libm (bit parity with the CPU)
  int32_t tetra_e[4096];        // key = 6 bases x 2 bit; BF_NO_SPECIAL if absent
  int32_t tri_e[1024];          // key = 5 bases x 2 bit
  int32_t n_hexa, hexa_key[BF_MAX_HEXA], hexa_e[BF_MAX_HEXA];  // key = 8 bases x 2 bit
  int32_t n_tri, n_tetra, year;
  double lxc, kT;
  // ---- Boltzmann weights exp(-E*10/kT)
  BfSmallD sd;
  double x_int11[8][8][5][5];
  double x_int21[8][8][5][5][5];
  double x_int22[8][8][5][5][5][5];
  double x_hp_big[BF_EXT_TAB];  // u > 30: exp(-(hairpin[30] + lxc*ln(u/30))*10/kT), un-truncated
  double x_tetra[4096], x_tri[1024], x_hexa[BF_MAX_HEXA];
};

#ifdef __cplusplus
#include <string>
// host-side loader (bf_params.cc)
int bf_params_parse_file(const char *path, BfParams *out, std::string *err);
int bf_params_save_image(const char *path, const BfParams *p, std::string *err);
int bf_params_load_image(const char *path, BfParams *out, std::string *err);
#endif
