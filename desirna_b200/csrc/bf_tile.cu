// bf_tile.cu -- tile-wavefront fill kernels for single-strand folds (sm_100a).
//
// Same recurrences and the same HBM tables as bf_fill.cu (the path behind fc.mfe() / fc.pf(),
// utils/energy_scores.py:150-151 in the reference; recurrences: SURVEY.md A.4-A.6), scheduled
// differently.  The (i,j) triangle is cut into 4x4 TILES and processed by TILE-DIAGONAL D = J - I:
//
//   * multiloop split  min_u fML[i][u-1] + fML[u][j]  is a (min,+) product of 4x4 blocks: one thread
//     owns the 16 cells of a tile and walks the intermediate tiles K, 8 x 128-bit loads for 64
//     relaxations (fML is kept TILE-MAJOR so every 4-row of a tile is one aligned vector).  Because
//     TURN = 3, blocks that would depend on the tile itself hold only +INF: the split of tile-diagonal
//     D needs tile-diagonals < D only.
//   * interior loops with a far inner pair (u1 >= 3 or u2 >= 3, "bulk" taps) also reach only earlier
//     tile-diagonals.  They are evaluated LANES-OVER-TAPS: a warp takes one pairable cell, each lane
//     owns a fixed (u1,u2) per pass with its ring offset and length penalty in registers, one
//     load + add-min per 32 candidates, one REDUX per cell.  Inner-pair terms are pre-folded into three
//     ring variants exactly as in bf_fill.cu.
//   * the 11 near taps (u1,u2 <= 2, plus 2x3 / 3x2), the hairpin, the multiloop closing and the fML
//     recurrences are the only in-tile dependencies: one warp owns 8 tiles and walks the 7 in-tile
//     sub-diagonals with __syncwarp() only.
//
// Two CTA barriers per TILE-diagonal (n/4 phases) instead of one per diagonal, ~2 instructions per 32
// interior candidates and ~1.1 per split relaxation.
#include "bf_kernels.h"

#include <cstdlib>
#include <cstring>

#include "bf_device.cuh"

namespace {

constexpr int kR = 40;          // ring depth: 7 diagonals in flight + 32 of look-back (+1)
constexpr int kNPass = 16;      // 12 generic + 2 (1xn) + 2 (bulge) passes of 32 taps
constexpr int kInfThr = BF_INF / 2;

__constant__ int c_tap[kNPass * 32];      // s | u1 << 8 ; s = 127: padding slot
__constant__ int c_pass_smin[kNPass];     // smallest loop size in the pass (skip the pass when smin > d-6)

__host__ __device__ __forceinline__ int tri_off(int n, int d) { return (d - 4) * n - (d * (d - 1) / 2 - 6); }
__host__ __device__ __forceinline__ size_t tri_size(int n) { return n >= 5 ? (size_t)tri_off(n, n) : 0; }
__host__ __device__ __forceinline__ int tile_off(int NT, int I) { return I * NT - I * (I - 1) / 2; }  // first tile of tile-row I
__host__ __device__ __forceinline__ int tile_idx(int NT, int I, int J) { return tile_off(NT, I) + (J - I); }

// units (a,b) of a tile sorted by in-tile sub-diagonal e = b - a
__device__ __forceinline__ int unit_a(int u) { return (int)((0x3231230123012010ull >> (4 * (15 - u))) & 15); }
__device__ __forceinline__ int unit_b(int u) { return (int)((0x0010120123123233ull >> (4 * (15 - u))) & 15); }

struct TilePlanI {
  int rs, nt;
  size_t o_S, o_SP, o_ring, o_fm, o_spl, o_ps, o_bi, total;
};
__host__ __device__ inline TilePlanI tile_plan_i(int nmax, int nws, bool fm_smem) {
  TilePlanI p;
  p.nt = (nmax + 3) / 4;
  p.rs = (nmax + 2 + 3) / 4 * 4;
  size_t o = 0;
  p.o_ring = o; o += ((size_t)3 * kR * p.rs + 64) * sizeof(int);
  p.o_fm = o; o += fm_smem ? (size_t)(p.nt * (p.nt + 1) / 2) * 16 * sizeof(int) : 0;
  p.o_spl = o; o += (size_t)3 * p.nt * 16 * sizeof(int);
  p.o_ps = o; o += (size_t)(nws > 1 ? nws : 0) * p.nt * 16 * sizeof(int);
  p.o_bi = o; o += (size_t)p.nt * 16 * sizeof(int);
  p.o_S = o; o += (nmax + 2 + 15) / 16 * 16;
  p.o_SP = o; o += (nmax + 2 + 15) / 16 * 16;
  p.total = o;
  return p;
}

// =====================================================================================================
//                                           MFE fill
// =====================================================================================================
template <int NW, int NWS, bool FM_SMEM>
__global__ void __launch_bounds__(NW * 32, (NW <= 8 ? 2 : 1)) bf_k_mfe_tile(const BfParams *__restrict__ P, BfBatchDev b, int *ctri, int *ftri, size_t tri_slot,
                                                         int *ws, size_t ws_slot, int *work_counter) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ int s_seq;
  constexpr int NWB = NW - NWS;  // warps of the bulk-interior part
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nmax = b.stride;
  const TilePlanI pl = tile_plan_i(nmax, NWS, FM_SMEM);
  const int RS = pl.rs;
  uint8_t *S = dyn + pl.o_S;
  uint8_t *SP = dyn + pl.o_SP;
  int *RING = reinterpret_cast<int *>(dyn + pl.o_ring);  // [variant][row][pos]: CG, C1, CB
  int *FM = FM_SMEM ? reinterpret_cast<int *>(dyn + pl.o_fm) : ws + (size_t)blockIdx.x * ws_slot;
  int *SPL = reinterpret_cast<int *>(dyn + pl.o_spl);    // split minima of the last 3 tile-diagonals
  int *PS = reinterpret_cast<int *>(dyn + pl.o_ps);      // per-split-warp partial minima (NWS > 1)
  int *BI = reinterpret_cast<int *>(dyn + pl.o_bi);      // bulk-interior minima of the current tile-diagonal
  const BfSmallI &T = P->si;

  // ---- per-lane tap constants: packed (pen0 | s << 16 | u1 << 24), pen0 = length penalty of the tap
  int tapk[kNPass];
#pragma unroll
  for (int p = 0; p < kNPass; p++) {
    const int v = c_tap[p * 32 + lane];
    const int s = v & 255, u1 = (v >> 8) & 255;
    int pen = 0;
    if (s != 127) {
      if (p < 12) pen = T.interior[s] + min(T.ninio_max, abs(s - 2 * u1) * T.ninio_m);
      else if (p < 14) pen = T.interior[s] + min(T.ninio_max, (s - 2) * T.ninio_m);
      else pen = T.bulge[s];
    }
    tapk[p] = pen | (s << 16) | (u1 << 24);
  }
  for (int k = tid; k < 3 * kR * RS + 64; k += blockDim.x) RING[k] = BF_INF;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_seq = atomicAdd(work_counter, 1);
    __syncthreads();
    const int sq = s_seq;
    if (sq >= b.B) break;
    const int n = b.len[sq];
    const int NT = (n + 3) >> 2;
    {
      const char *src = b.seq + (size_t)sq * b.stride;
      const uint8_t *np = b.nopair ? b.nopair + (size_t)sq * b.stride : nullptr;
      for (int k = tid; k <= n + 1; k += blockDim.x) {
        int code = (k >= 1 && k <= n) ? bf_base_code(src[k - 1]) : 0;
        S[k] = (uint8_t)code;
        SP[k] = (uint8_t)((np && k >= 1 && k <= n && np[k - 1]) ? 0 : code);
      }
    }
    for (int k = tid; k < (NT * (NT + 1) / 2) * 16; k += blockDim.x) FM[k] = BF_INF;
    for (int k = tid; k < 3 * NT * 16; k += blockDim.x) SPL[k] = BF_INF;
    int *cg_out = ctri + (size_t)sq * tri_slot;
    int *fg_out = ftri + (size_t)sq * tri_slot;
    int cur_d = -1;
    int off[kNPass], pen[kNPass];
#pragma unroll
    for (int p = 0; p < kNPass; p++) { off[p] = 0; pen[p] = BF_INF; }
    __syncthreads();

    for (int D = 1; D < NT; D++) {
      const int Tn = NT - D;  // tiles on this tile-diagonal: I = 0 .. Tn-1, J = I + D
      // =================================================================== step 1a: split (warps 0 .. NWS-1)
      if (warp < NWS) {
        int *dst = (NWS > 1) ? PS + (size_t)warp * NT * 16 : SPL + (size_t)(D % 3) * NT * 16;
        int g = 1;
        while (g * 2 * min(Tn, 32) <= 32 && g * 2 * NWS <= max(1, D - 1)) g *= 2;
        const int tpc = 32 / g;  // tiles per chunk
        for (int c0 = 0; c0 < Tn; c0 += tpc) {
          const int I = c0 + lane / g, J = I + D;
          const int sl = warp * g + (lane % g), nsl = NWS * g;
          int acc[16];
#pragma unroll
          for (int k = 0; k < 16; k++) acc[k] = BF_INF;
          if (I < Tn) {
            const int *rowI = FM + (size_t)tile_off(NT, I) * 16;
            for (int K = I + 1 + sl; K <= J - 1; K += nsl) {
              const int4 *lp = reinterpret_cast<const int4 *>(rowI + (K - I) * 16);
              const int4 *rp = reinterpret_cast<const int4 *>(FM + (size_t)tile_idx(NT, K, J) * 16);
              const int4 *rq = reinterpret_cast<const int4 *>(FM + (size_t)tile_idx(NT, K + 1, J) * 16);
              const int4 r0 = rp[1], r1 = rp[2], r2 = rp[3], r3 = rq[0];  // rows u = 4K+c+1, c = 0..3
#pragma unroll
              for (int a = 0; a < 4; a++) {
                const int4 l = lp[a];  // fML[4I+a][4K+c], c = 0..3
                acc[a * 4 + 0] = min(acc[a * 4 + 0], min(min(l.x + r0.x, l.y + r1.x), min(l.z + r2.x, l.w + r3.x)));
                acc[a * 4 + 1] = min(acc[a * 4 + 1], min(min(l.x + r0.y, l.y + r1.y), min(l.z + r2.y, l.w + r3.y)));
                acc[a * 4 + 2] = min(acc[a * 4 + 2], min(min(l.x + r0.z, l.y + r1.z), min(l.z + r2.z, l.w + r3.z)));
                acc[a * 4 + 3] = min(acc[a * 4 + 3], min(min(l.x + r0.w, l.y + r1.w), min(l.z + r2.w, l.w + r3.w)));
              }
            }
          }
          for (int o = g >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] = min(acc[k], __shfl_xor_sync(BF_FULL, acc[k], o));
          }
          if (I < Tn && (lane % g) == 0) {
            int4 *d4 = reinterpret_cast<int4 *>(dst + (size_t)I * 16);
            d4[0] = make_int4(acc[0], acc[1], acc[2], acc[3]);
            d4[1] = make_int4(acc[4], acc[5], acc[6], acc[7]);
            d4[2] = make_int4(acc[8], acc[9], acc[10], acc[11]);
            d4[3] = make_int4(acc[12], acc[13], acc[14], acc[15]);
          }
        }
      } else {
        // ================================================================= step 1b: bulk interior taps, lanes over taps
        const int wb = warp - NWS;
        const int nchunk = (Tn + 15) >> 4;  // items = (unit, 16-tile chunk)
        const int nitems = 16 * nchunk;
        const int it0 = wb * nitems / NWB, it1 = (wb + 1) * nitems / NWB;
        for (int item = it0; item < it1; item++) {
          const int u = item / nchunk, ch = item - u * nchunk;
          const int a = unit_a(u), bq = unit_b(u);
          const int d = 4 * D + bq - a;
          if (d < 9) continue;  // the smallest bulk tap (s = 3) needs an inner diagonal d-5 >= 4
          if (d != cur_d) {
            cur_d = d;
            const int r0 = (d - 2) % kR;
#pragma unroll
            for (int p = 0; p < kNPass; p++) {
              const int s = (tapk[p] >> 16) & 255, u1 = (tapk[p] >> 24) & 255;
              int row = r0 - (s & 31);
              if (row < 0) row += kR;
              const int var = p < 12 ? 0 : p < 14 ? 1 : 2;
              off[p] = (var * kR + row) * RS + 1 + u1;
              pen[p] = (s != 127 && s <= d - 6) ? (tapk[p] & 0xffff) : BF_INF;  // s = 127: padding slot
            }
          }
          const int Il = ch * 16 + (lane & 15);
          const int il = 4 * Il + a + 1, jl = 4 * (Il + D) + bq + 1;
          const bool okl = lane < 16 && Il < Tn && jl <= n && bf_ptype_bases(SP[il], SP[jl]) != 0;
          unsigned mk = __ballot_sync(BF_FULL, okl);
          while (mk) {
            const int l0 = __ffs(mk) - 1;
            mk &= mk - 1;
            const int l1 = mk ? __ffs(mk) - 1 : l0;
            mk &= mk - 1;
            const int I0 = ch * 16 + l0, I1 = ch * 16 + l1;
            const int i0 = 4 * I0 + a + 1, i1 = 4 * I1 + a + 1;
            const int *rg0 = RING + i0, *rg1 = RING + i1;
            int g0 = BF_INF, g1 = BF_INF, h0 = BF_INF, h1 = BF_INF, b0 = BF_INF, b1 = BF_INF;
            const int dm6 = d - 6;
#pragma unroll
            for (int p = 0; p < 12; p++)
              if (c_pass_smin[p] <= dm6) { g0 = min(g0, rg0[off[p]] + pen[p]); g1 = min(g1, rg1[off[p]] + pen[p]); }
#pragma unroll
            for (int p = 12; p < 14; p++)
              if (c_pass_smin[p] <= dm6) { h0 = min(h0, rg0[off[p]] + pen[p]); h1 = min(h1, rg1[off[p]] + pen[p]); }
#pragma unroll
            for (int p = 14; p < 16; p++)
              if (c_pass_smin[p] <= dm6) { b0 = min(b0, rg0[off[p]] + pen[p]); b1 = min(b1, rg1[off[p]] + pen[p]); }
            const int j0 = i0 + d, j1 = i1 + d;
            const int t0 = bf_ptype_bases(SP[i0], SP[j0]), t1 = bf_ptype_bases(SP[i1], SP[j1]);
            int v0 = min(g0 + T.mmI[t0][S[i0 + 1]][S[j0 - 1]], min(h0 + T.mm1nI[t0][S[i0 + 1]][S[j0 - 1]], b0 + (t0 > 2 ? T.TerminalAU : 0)));
            int v1 = min(g1 + T.mmI[t1][S[i1 + 1]][S[j1 - 1]], min(h1 + T.mm1nI[t1][S[i1 + 1]][S[j1 - 1]], b1 + (t1 > 2 ? T.TerminalAU : 0)));
            v0 = __reduce_min_sync(BF_FULL, v0);
            v1 = __reduce_min_sync(BF_FULL, v1);
            if (lane == 0) { BI[I0 * 16 + a * 4 + bq] = v0; BI[I1 * 16 + a * 4 + bq] = v1; }
          }
        }
      }
      __syncthreads();
      // =================================================================== step 2: near part, in-tile sub-diagonals
      for (int grp = warp; grp * 8 < Tn; grp += NW) {
        const int I = grp * 8 + (lane >> 2), r = lane & 3, J = I + D;
        const bool tile_ok = I < Tn;
        const int tix = tile_ok ? tile_idx(NT, I, J) : 0;
        for (int e = -3; e <= 3; e++) {
          const int d = 4 * D + e;
          if (d <= BF_TURN) continue;  // warp-uniform
          const int a = max(0, -e) + r, bq = a + e;
          const int i = 4 * I + a + 1, j = i + d;
          if (tile_ok && r < 4 - abs(e) && j <= n) {
            const int t = bf_ptype_bases(SP[i], SP[j]);
            const int ab = a * 4 + bq;
            int en = BF_INF;
            int sp = SPL[(D % 3) * NT * 16 + I * 16 + ab];
            if (NWS > 1) {
              sp = PS[I * 16 + ab];
#pragma unroll
              for (int w = 1; w < NWS; w++) sp = min(sp, PS[(size_t)w * NT * 16 + I * 16 + ab]);
            }
            if (sp >= kInfThr) sp = BF_INF;
            if (t) {
              const int si1 = S[i + 1], sj1 = S[j - 1];
              if (d >= 9) en = BI[I * 16 + ab];
              en = min(en, bf_e_hairpin(P, T, S, i, j, t));
#pragma unroll 1
              for (int k = 0; k < 11; k++) {
                const int u1 = (int)((0x20322211100ull >> (4 * k)) & 15), u2 = (int)((0x02232121010ull >> (4 * k)) & 15);
                const int p = i + 1 + u1, q = j - 1 - u2;
                if (q - p <= BF_TURN) continue;
                int cc = RING[(2 * kR + (q - p) % kR) * RS + p];  // c + terminalAU(inner pair)
                if (cc >= kInfThr) continue;
                const int t2 = bf_ptype_bases(SP[p], SP[q]);
                if (t2 > 2) cc -= T.TerminalAU;
                en = min(en, cc + bf_e_intloop(P, T, u1, u2, t, bf_rtype(t2), si1, sj1, S[p - 1], S[q + 1]));
              }
              // multiloop closing: split minimum of cell (i+1, j-1)
              {
                int I2 = I, a2 = a + 1, J2 = J, b2 = bq - 1;
                if (a2 == 4) { a2 = 0; I2++; }
                if (b2 < 0) { b2 = 3; J2--; }
                if (d >= 11) {  // a split of (i+1, j-1) needs j-1 - (i+1) >= 9
                  const int dm = SPL[((J2 - I2) % 3) * NT * 16 + I2 * 16 + a2 * 4 + b2];
                  if (dm < kInfThr) en = min(en, dm + T.MLclosing + bf_e_mlstem(T, bf_rtype(t), sj1, si1));
                }
              }
              if (en >= kInfThr) en = BF_INF;
            }
            int m = sp;
            if (en < BF_INF && i > 1 && j < n) m = min(m, en + bf_e_mlstem(T, t, S[i - 1], S[j + 1]));
            if (d > BF_TURN + 1) {
              const int fa = (a < 3) ? FM[(size_t)tix * 16 + ab + 4] : FM[(size_t)tile_idx(NT, I + 1, J) * 16 + bq];
              const int fb = (bq > 0) ? FM[(size_t)tix * 16 + ab - 1] : FM[(size_t)(tix - 1) * 16 + a * 4 + 3];
              m = min(m, min(fa, fb) + T.MLbase);
            }
            if (m >= kInfThr) m = BF_INF;
            const int o = tri_off(n, d) + i - 1;
            cg_out[o] = en;
            fg_out[o] = m;
            FM[(size_t)tix * 16 + ab] = m;
            SPL[(D % 3) * NT * 16 + I * 16 + ab] = sp;
            int eg = BF_INF, e1 = BF_INF, eb = BF_INF;
            if (en < BF_INF) {
              const int t2 = bf_rtype(t), x = S[j + 1], y = S[i - 1];  // as an inner pair: sq1 = S[q+1], sp1 = S[p-1]
              eg = en + T.mmI[t2][x][y];
              e1 = en + T.mm1nI[t2][x][y];
              eb = en + (t > 2 ? T.TerminalAU : 0);
            }
            const int row = (d % kR) * RS + i;
            RING[row] = eg;
            RING[kR * RS + row] = e1;
            RING[2 * kR * RS + row] = eb;
          }
          __syncwarp();
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool g_tap_uploaded = false;

cudaError_t upload_taps() {
  if (g_tap_uploaded) return cudaSuccess;
  int tap[kNPass * 32], smin[kNPass];
  for (int k = 0; k < kNPass * 32; k++) tap[k] = 127;
  for (int p = 0; p < kNPass; p++) smin[p] = 127;
  int k = 0;
  auto put = [&](int base_pass, int s, int u1) {
    const int slot = base_pass * 32 + k;
    tap[slot] = s | (u1 << 8);
    const int p = slot / 32;
    if (s < smin[p]) smin[p] = s;
    k++;
  };
  k = 0;  // generic: u1, u2 >= 2 except 2x2, 2x3, 3x2  ->  s = 6..30, u1 = 2..s-2
  for (int s = 6; s <= BF_MAXLOOP; s++)
    for (int u1 = 2; u1 <= s - 2; u1++) put(0, s, u1);
  k = 0;  // 1xn: (1, s-1) and (s-1, 1), s = 4..30
  for (int s = 4; s <= BF_MAXLOOP; s++) { put(12, s, 1); put(12, s, s - 1); }
  k = 0;  // bulges (0, s) and (s, 0), s = 3..30 (s = 1, 2 are near taps)
  for (int s = 3; s <= BF_MAXLOOP; s++) { put(14, s, 0); put(14, s, s); }
  cudaError_t e = cudaMemcpyToSymbol(c_tap, tap, sizeof(tap));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyToSymbol(c_pass_smin, smin, sizeof(smin));
  if (e != cudaSuccess) return e;
  g_tap_uploaded = true;
  return cudaSuccess;
}

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

constexpr size_t kSmemBudget = 232448 - 1024 - 256;

template <typename K>
cudaError_t set_smem(K kern, size_t sm) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sm > 1024 ? sm : 1024));
}

struct TileCfg { int nw, nws; bool fm_smem; };

TileCfg mfe_tile_cfg(int nmax) {
  TileCfg c;
  c.nw = env_int("BF_TILE_NW", 8);
  if (c.nw != 12 && c.nw != 16) c.nw = 8;
  c.nws = env_int("BF_TILE_NWS", nmax > 160 ? 2 : 1);
  if (c.nws != 2) c.nws = 1;
  c.fm_smem = tile_plan_i(nmax, c.nws, true).total <= (size_t)env_int("BF_TILE_FM_SMEM_MAX", 112 * 1024);
  return c;
}

template <int NW, int NWS, bool FM_SMEM>
cudaError_t mfe_tile_t(const BfParams *dP, const BfBatchDev &b, int *ctri, int *ftri, int *ws, size_t ws_slot, int sms, int *grid_out, bool launch,
                       int *counter, cudaStream_t st) {
  auto kern = bf_k_mfe_tile<NW, NWS, FM_SMEM>;
  const size_t sm = tile_plan_i(b.stride, NWS, FM_SMEM).total;
  if (sm > kSmemBudget) return cudaErrorInvalidConfiguration;
  cudaError_t e = set_smem(kern, sm);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, sm);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  const int grid = b.B < sms * occ ? b.B : sms * occ;
  if (grid_out) *grid_out = grid;
  if (!launch) return cudaSuccess;
  e = upload_taps();
  if (e != cudaSuccess) return e;
  kern<<<grid, NW * 32, sm, st>>>(dP, b, ctri, ftri, bf_tri_slot(b.stride), ws, ws_slot, counter);
  return cudaGetLastError();
}

template <int NW>
cudaError_t mfe_tile_nw(const TileCfg &c, const BfParams *dP, const BfBatchDev &b, int *ctri, int *ftri, int *ws, size_t ws_slot, int sms,
                        int *grid_out, bool launch, int *counter, cudaStream_t st) {
  if (c.nws == 2) {
    if (c.fm_smem) return mfe_tile_t<NW, 2, true>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
    return mfe_tile_t<NW, 2, false>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
  }
  if (c.fm_smem) return mfe_tile_t<NW, 1, true>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
  return mfe_tile_t<NW, 1, false>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
}

cudaError_t mfe_tile_dispatch(const BfParams *dP, const BfBatchDev &b, int *ctri, int *ftri, int *ws, int sms, int *grid_out, bool launch,
                              int *counter, cudaStream_t st) {
  const TileCfg c = mfe_tile_cfg(b.stride);
  const size_t ws_slot = bf_mfe_tile_ws_slot(b.stride);
  if (c.nw == 12) return mfe_tile_nw<12>(c, dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
  if (c.nw == 16) return mfe_tile_nw<16>(c, dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
  return mfe_tile_nw<8>(c, dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
}

}  // namespace

// 1 if the tile path covers this length (ring rows and tile tables must fit the CTA's shared memory)
int bf_tile_mfe_ok(int nmax) {
  if (nmax < 1 || nmax > 2000) return 0;
  const TileCfg c = mfe_tile_cfg(nmax);
  return tile_plan_i(nmax, c.nws, c.fm_smem).total <= kSmemBudget ? 1 : 0;
}
size_t bf_mfe_tile_ws_slot(int nmax) {  // ints of per-CTA HBM workspace (tile-major fML when it is not on chip)
  const TileCfg c = mfe_tile_cfg(nmax);
  if (c.fm_smem) return 0;
  const size_t nt = (nmax + 3) / 4;
  return (nt * (nt + 1) / 2 * 16 + 7) / 8 * 8;
}
cudaError_t bf_mfe_tile_grid(const BfBatchDev &b, int sms, int *grid) {
  return mfe_tile_dispatch(nullptr, b, nullptr, nullptr, nullptr, sms, grid, false, nullptr, nullptr);
}
cudaError_t bf_launch_mfe_tile(const BfParams *dP, const BfBatchDev &b, int *ctri, int *ftri, int *ws, int sms, int *work_counter,
                               cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return e;
  return mfe_tile_dispatch(dP, b, ctri, ftri, ws, sms, nullptr, true, work_counter, st);
}
