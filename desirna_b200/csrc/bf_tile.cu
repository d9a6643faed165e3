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
__constant__ int c_pass_smin[kNPass];
#ifdef BF_TILE_TRACE
__device__ long long g_trace[128 * 16 * 4];  // [phase][warp][4 time stamps], CTA 0, first sequence
#define TR(slot) do { if (blockIdx.x == 0 && first_seq && lane == 0 && D < 128) g_trace[(D * 16 + warp) * 4 + (slot)] = clock64(); } while (0)
#else
#define TR(slot) do {} while (0)
#endif     // smallest loop size in the pass (skip the pass when smin > d-6)

__host__ __device__ __forceinline__ int tri_off(int n, int d) { return (d - 4) * n - (d * (d - 1) / 2 - 6); }
__host__ __device__ __forceinline__ size_t tri_size(int n) { return n >= 5 ? (size_t)tri_off(n, n) : 0; }
__host__ __device__ __forceinline__ int tile_off(int NT, int I) { return I * NT - I * (I - 1) / 2; }  // first tile of tile-row I
__host__ __device__ __forceinline__ int tile_idx(int NT, int I, int J) { return tile_off(NT, I) + (J - I); }

// units (a,b) of a tile sorted by in-tile sub-diagonal e = b - a
__device__ __forceinline__ int unit_a(int u) { return (int)((0x3231230123012010ull >> (4 * (15 - u))) & 15); }
__device__ __forceinline__ int unit_b(int u) { return (int)((0x0010120123123233ull >> (4 * (15 - u))) & 15); }

struct TilePlanI {
  int rs, nt;
  size_t o_S, o_SP, o_ring, o_fm, o_spl, o_ps, o_bi, o_list, o_tab, total;
};
// big: the tile-major fML table and the per-cell sequence-only terms live in a per-CTA HBM workspace (long sequences)
__host__ __device__ inline TilePlanI tile_plan_i(int nmax, int nws, bool big) {
  TilePlanI p;
  p.nt = (nmax + 3) / 4;
  p.rs = (nmax + 2 + 3) / 4 * 4;
  size_t o = 0;
  p.o_ring = o; o += ((size_t)3 * kR * p.rs + 64) * sizeof(int);
  p.o_fm = o; o += big ? 0 : (size_t)(p.nt * (p.nt + 1) / 2) * 16 * sizeof(int);
  p.o_spl = o; o += (size_t)3 * p.nt * 16 * sizeof(int);
  p.o_ps = o; o += (size_t)(nws > 1 ? nws : 0) * p.nt * 16 * sizeof(int);
  p.o_bi = o; o += (size_t)p.nt * 16 * sizeof(int);
  p.o_tab = o; o += (sizeof(BfSmallI) + 15) / 16 * 16;
  p.o_list = o; o += (size_t)2 * p.nt * 16 * sizeof(unsigned short);
  p.o_S = o; o += (nmax + 2 + 15) / 16 * 16;
  p.o_SP = o; o += (nmax + 2 + 15) / 16 * 16;
  p.total = o;
  return p;
}
// ints of per-CTA HBM workspace: 32 bytes of sequence-only terms per cell (tile-major), then, in "big" mode, the tile-major fML table
__host__ __device__ inline size_t tile_ws_ints(int nmax, bool big) {
  const size_t nt = (nmax + 3) / 4, ntile = nt * (nt + 1) / 2;
  return (ntile * 16 * 8 + (big ? ntile * 16 : 0) + 7) / 8 * 8;
}

__device__ __forceinline__ int lds_s32(unsigned addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int clamp16(int v) { return max(-32768, min(32767, v)); }

// near taps (u1, u2), k = 0..10: 0x0 0x1 1x0 1x1 1x2 2x1 2x2 2x3 3x2 0x2 2x0
__host__ __device__ constexpr int near_u1(int k) { return (int)((0x20322211100ull >> (4 * k)) & 15); }
__host__ __device__ constexpr int near_u2(int k) { return (int)((0x02232121010ull >> (4 * k)) & 15); }

// lanes-over-taps minimum over the bulk taps of two cells of the same diagonal (interleaved for ILP);
// cellb = shared-window byte address of RING + i
template <int NG, bool FAR>
__device__ __forceinline__ void bulk_cell2(unsigned cb0, unsigned cb1, const int (&offb)[kNPass], const int (&pen)[kNPass], int &g0, int &h0, int &b0,
                                           int &g1, int &h1, int &b1) {
  g0 = BF_INF; g1 = BF_INF;
#pragma unroll
  for (int p = 0; p < NG; p++) {
    g0 = min(g0, lds_s32(cb0 + offb[p]) + pen[p]);
    g1 = min(g1, lds_s32(cb1 + offb[p]) + pen[p]);
  }
  h0 = lds_s32(cb0 + offb[12]) + pen[12];
  h1 = lds_s32(cb1 + offb[12]) + pen[12];
  b0 = lds_s32(cb0 + offb[14]) + pen[14];
  b1 = lds_s32(cb1 + offb[14]) + pen[14];
  if (FAR) {
    h0 = min(h0, lds_s32(cb0 + offb[13]) + pen[13]);
    h1 = min(h1, lds_s32(cb1 + offb[13]) + pen[13]);
    b0 = min(b0, lds_s32(cb0 + offb[15]) + pen[15]);
    b1 = min(b1, lds_s32(cb1 + offb[15]) + pen[15]);
  }
}

// Compact the pairable cells of tile-diagonal D into list[] (sorted by in-tile sub-diagonal, i.e. by d).  One warp.
__device__ __forceinline__ int build_cell_list(const uint8_t *SP, int n, int NT, int D, unsigned short *list, int lane) {
  const int Tn = NT - D;
  int count = 0;
  for (int u = 0; u < 16; u++) {
    const int a = unit_a(u), bq = unit_b(u);
    const int d = 4 * D + bq - a;
    if (d <= BF_TURN) continue;
    for (int c0 = 0; c0 < Tn; c0 += 32) {
      const int I = c0 + lane;
      const int i = 4 * I + a + 1, j = i + d;
      const bool ok = I < Tn && j <= n && bf_ptype_bases(SP[i], SP[j]) != 0;
      const unsigned mk = __ballot_sync(BF_FULL, ok);
      if (ok) {
        const int pos = count + __popc(mk & ((1u << lane) - 1));
        list[pos] = (unsigned short)(I * 16 + a * 4 + bq);
      }
      count += __popc(mk);
    }
  }
  return count;
}

// =====================================================================================================
//                                           MFE fill
// =====================================================================================================
template <int NW, int NWS, bool BIG>
__global__ void __launch_bounds__(NW * 32, (NW <= 8 ? 2 : 1)) bf_k_mfe_tile(const BfParams *__restrict__ P, BfBatchDev b, int *ctri, int *ftri,
                                                                           size_t tri_slot, int *ws, size_t ws_slot, int *work_counter) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ int s_seq, s_cnt[2];
  constexpr int NWB = NW - NWS;  // warps of the bulk-interior part
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nmax = b.stride;
  const TilePlanI pl = tile_plan_i(nmax, NWS, BIG);
  const int RS = pl.rs;
  uint8_t *S = dyn + pl.o_S;
  uint8_t *SP = dyn + pl.o_SP;
  int *RING = reinterpret_cast<int *>(dyn + pl.o_ring);  // [variant][row][pos]: CG, C1, CB
  int *wsp = ws + (size_t)blockIdx.x * ws_slot;
  // per cell (tile-major, 16 shorts): 0..10 near-tap loop energies minus terminalAU(inner pair), 11 hairpin, 12 ML closing stem,
  // 13 ML stem as a branch, 14 / 15 mismatch terms of the cell as the INNER pair of a generic / 1xn loop.  Sequence-only,
  // written once per sequence by the whole CTA, L2-resident.
  int4 *REC = reinterpret_cast<int4 *>(wsp);
  int *FM = BIG ? wsp + (size_t)(pl.nt * (pl.nt + 1) / 2) * 16 * 8 : reinterpret_cast<int *>(dyn + pl.o_fm);
  int *SPL = reinterpret_cast<int *>(dyn + pl.o_spl);    // split minima of the last 3 tile-diagonals
  int *PS = reinterpret_cast<int *>(dyn + pl.o_ps);      // per-split-warp partial minima (NWS > 1)
  int *BI = reinterpret_cast<int *>(dyn + pl.o_bi);      // bulk-interior minima of the current tile-diagonal
  unsigned short *LST = reinterpret_cast<unsigned short *>(dyn + pl.o_list);
  BfSmallI *Tsm = reinterpret_cast<BfSmallI *>(dyn + pl.o_tab);
  bf_stage(Tsm, &P->si);
  const BfSmallI &T = *Tsm;
  const unsigned ring_b = (unsigned)__cvta_generic_to_shared(RING);
  for (int k = tid; k < 3 * kR * RS + 64; k += blockDim.x) RING[k] = BF_INF;
  __syncthreads();

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

#ifdef BF_TILE_TRACE
  bool first_seq = true;
#endif
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
    __syncthreads();
    // ---- sequence-only terms of every pairable cell (all table look-ups of the near part happen here, off the wavefront)
    for (int I = warp; I < NT - 1; I += NW) {
      const int base = tile_off(NT, I);
      for (int x = 16 + lane; x < (NT - I) * 16; x += 32) {  // tile J = I + x/16 >= I+1, cell x%16
        const int J = I + (x >> 4), a = (x >> 2) & 3, bq = x & 3;
        const int i = 4 * I + a + 1, j = 4 * J + bq + 1;
        if (j > n || j - i <= BF_TURN) continue;
        const int t = bf_ptype_bases(SP[i], SP[j]);
        if (!t) continue;
        const int si1 = S[i + 1], sj1 = S[j - 1];
        int v[16];
#pragma unroll
        for (int k = 0; k < 11; k++) {
          const int u1 = near_u1(k), u2 = near_u2(k);
          const int p = i + 1 + u1, q = j - 1 - u2;
          int e = 0;
          if (q - p > BF_TURN) {
            const int t2 = bf_ptype_bases(SP[p], SP[q]);
            if (t2) e = bf_e_intloop(P, T, u1, u2, t, bf_rtype(t2), si1, sj1, S[p - 1], S[q + 1]) - (t2 > 2 ? T.TerminalAU : 0);
          }
          v[k] = clamp16(e);
        }
        v[11] = clamp16(bf_e_hairpin(P, T, S, i, j, t));
        v[12] = clamp16(T.MLclosing + bf_e_mlstem(T, bf_rtype(t), sj1, si1));
        v[13] = clamp16(bf_e_mlstem(T, t, S[i - 1], S[j + 1]));
        v[14] = T.mmI[bf_rtype(t)][S[j + 1]][S[i - 1]];
        v[15] = T.mm1nI[bf_rtype(t)][S[j + 1]][S[i - 1]];
        int w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = (v[2 * k] & 0xffff) | (v[2 * k + 1] << 16);
        int4 *dst = REC + (size_t)(base + (x >> 4)) * 32 + (x & 15) * 2;
        dst[0] = make_int4(w[0], w[1], w[2], w[3]);
        dst[1] = make_int4(w[4], w[5], w[6], w[7]);
      }
    }
    int *cg_out = ctri + (size_t)sq * tri_slot;
    int *fg_out = ftri + (size_t)sq * tri_slot;
    int cur_d = -1;
    int offb[kNPass], pen[kNPass];
#pragma unroll
    for (int p = 0; p < kNPass; p++) { offb[p] = 0; pen[p] = BF_INF; }
    __syncthreads();
    if (warp == 0 && NT > 1) {
      const int cnt = build_cell_list(SP, n, NT, 1, LST + 1 * pl.nt * 16, lane);
      if (lane == 0) s_cnt[1] = cnt;
    }
    __syncthreads();

    for (int D = 1; D < NT; D++) {
      TR(0);
      const int Tn = NT - D;  // tiles on this tile-diagonal: I = 0 .. Tn-1, J = I + D
      const unsigned short *list = LST + (D & 1) * pl.nt * 16;
      // prefetch the sequence-only terms of this warp's first near task (step 2) so that they arrive during step 1
      int4 pre0 = make_int4(0, 0, 0, 0), pre1 = pre0;
      {
        const int I = warp * 2 + (lane >> 4), ab = lane & 15;
        const int i = 4 * I + (ab >> 2) + 1, j = 4 * (I + D) + (ab & 3) + 1;
        if (I < Tn && j <= n && j - i > BF_TURN) {
          const int4 *src = REC + (size_t)tile_idx(NT, I, I + D) * 32 + ab * 2;
          pre0 = src[0]; pre1 = src[1];
        }
      }
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
        // ================================================================= step 1b: this warp's share of the pairable cells
        const int wb = warp - NWS;
        const int cnt = s_cnt[D & 1];
        const int c_lo = wb * cnt / NWB, c_hi = (wb + 1) * cnt / NWB;
        // ---- bulk interior taps, lanes over taps, two cells of the same diagonal at a time
        int c = c_lo;
        while (c < c_hi) {
          const int ent0 = list[c];
          const int d = 4 * D + (ent0 & 3) - ((ent0 >> 2) & 3);
          if (d < 9) { c++; continue; }  // the smallest bulk tap (s = 3) needs an inner diagonal d-5 >= 4
          int ent1 = ent0;
          bool two = false;
          if (c + 1 < c_hi) {
            const int e1 = list[c + 1];
            if (4 * D + (e1 & 3) - ((e1 >> 2) & 3) == d) { ent1 = e1; two = true; }
          }
          if (d != cur_d) {
            cur_d = d;
            const int r0 = (d - 2) % kR;
#pragma unroll
            for (int p = 0; p < kNPass; p++) {
              const int s = (tapk[p] >> 16) & 255, u1 = (tapk[p] >> 24) & 255;
              int row = r0 - (s & 31);
              if (row < 0) row += kR;
              const int var = p < 12 ? 0 : p < 14 ? 1 : 2;
              offb[p] = ((var * kR + row) * RS + 1 + u1) * 4;
              pen[p] = (s != 127 && s <= d - 6) ? (tapk[p] & 0xffff) : BF_INF;  // s = 127: padding slot
            }
          }
          const int i0 = 4 * (ent0 >> 4) + ((ent0 >> 2) & 3) + 1, i1 = 4 * (ent1 >> 4) + ((ent1 >> 2) & 3) + 1;
          const int j0 = i0 + d, j1 = i1 + d;
          const int t0 = bf_ptype_bases(SP[i0], SP[j0]), t1 = bf_ptype_bases(SP[i1], SP[j1]);
          const int fG0 = T.mmI[t0][S[i0 + 1]][S[j0 - 1]], f10 = T.mm1nI[t0][S[i0 + 1]][S[j0 - 1]], fB0 = (t0 > 2 ? T.TerminalAU : 0);
          const int fG1 = T.mmI[t1][S[i1 + 1]][S[j1 - 1]], f11 = T.mm1nI[t1][S[i1 + 1]][S[j1 - 1]], fB1 = (t1 > 2 ? T.TerminalAU : 0);
          const unsigned cb0 = ring_b + 4u * (unsigned)i0, cb1 = ring_b + 4u * (unsigned)i1;
          int g0, h0, b0, g1, h1, b1;
          // generic passes hold s <= 10 | 14 | 18 | 22 | 26 | 30 after 1 | 2 | 4 | 6 | 9 | 12 passes; a tap needs s <= d-6
          if (d >= 33) bulk_cell2<12, true>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          else if (d >= 29) bulk_cell2<9, true>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          else if (d >= 25) bulk_cell2<6, true>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          else if (d >= 21) bulk_cell2<4, false>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          else if (d >= 17) bulk_cell2<2, false>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          else bulk_cell2<1, false>(cb0, cb1, offb, pen, g0, h0, b0, g1, h1, b1);
          int v0 = min(g0 + fG0, min(h0 + f10, b0 + fB0));
          int v1 = min(g1 + fG1, min(h1 + f11, b1 + fB1));
          v0 = __reduce_min_sync(BF_FULL, v0);
          v1 = __reduce_min_sync(BF_FULL, v1);
          if (lane == 0) { BI[ent0] = v0; BI[ent1] = v1; }
          c += two ? 2 : 1;
        }
      }
      TR(1);
      __syncthreads();
      TR(2);
      // =================================================================== step 2: near part, in-tile sub-diagonals
      if (warp == NW - 1 && D + 1 < NT) {
        const int cnt = build_cell_list(SP, n, NT, D + 1, LST + ((D + 1) & 1) * pl.nt * 16, lane);
        if (lane == 0) s_cnt[(D + 1) & 1] = cnt;
      }
      for (int task = warp; task * 2 < Tn; task += NW) {
        // one lane per cell of two tiles; everything that does not depend on the tile itself is done up front
        const int I = task * 2 + (lane >> 4), J = I + D;
        const int ab = lane & 15, a = ab >> 2, bq = ab & 3, e = bq - a;
        const int i = 4 * I + a + 1, j = 4 * J + bq + 1, d = j - i;
        const bool valid = I < Tn && j <= n && d > BF_TURN;
        const int t = valid ? bf_ptype_bases(SP[i], SP[j]) : 0;
        const int tix = I < Tn ? tile_idx(NT, I, J) : 0;
        int en = BF_INF, mlclose = 0, mlout = 0, dG = 0, d1 = 0, tau = 0;
        int nw[8];
        {
          int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
          if (task == warp) { n0 = pre0; n1 = pre1; }
          else if (t) { n0 = REC[(size_t)tix * 32 + ab * 2]; n1 = REC[(size_t)tix * 32 + ab * 2 + 1]; }
          nw[0] = n0.x; nw[1] = n0.y; nw[2] = n0.z; nw[3] = n0.w; nw[4] = n1.x; nw[5] = n1.y; nw[6] = n1.z; nw[7] = n1.w;
        }
        unsigned same = 0;
        int rb[6];  // CB-ring address of row (d-2-s), position i+1
        {
          const int r0 = (d + 2 * kR - 2) % kR;
#pragma unroll
          for (int s2 = 0; s2 < 6; s2++) {
            int row = r0 - s2;
            if (row < 0) row += kR;
            rb[s2] = (2 * kR + row) * RS + i + 1;
          }
        }
#define BF_REC(k) (((k) & 1) ? (nw[(k) >> 1] >> 16) : ((nw[(k) >> 1] << 16) >> 16))
        if (t) {
          if (d >= 9) en = BI[I * 16 + ab];
          en = min(en, BF_REC(11));  // hairpin
#pragma unroll
          for (int k = 0; k < 11; k++) {
            const int u1 = near_u1(k), u2 = near_u2(k);
            if (d - 2 - u1 - u2 > BF_TURN) {
              if (a + 1 + u1 <= 3 && bq - 1 - u2 >= 0) same |= 1u << k;                      // inner cell in this tile: later
              else en = min(en, RING[rb[u1 + u2] + u1] + BF_REC(k));                           // c + terminalAU(inner) from an earlier tile-diagonal (INF if it cannot pair)
            }
          }
          mlclose = BF_REC(12);
          if (d >= 11 && (a == 3 || bq == 0)) {  // multiloop closing from cell (i+1, j-1) of an earlier tile-diagonal
            int I2 = I, a2 = a + 1, J2 = J, b2 = bq - 1;
            if (a2 == 4) { a2 = 0; I2++; }
            if (b2 < 0) { b2 = 3; J2--; }
            const int dm = SPL[((J2 - I2) % 3) * NT * 16 + I2 * 16 + a2 * 4 + b2];
            if (dm < kInfThr) en = min(en, dm + mlclose);
          }
          mlout = BF_REC(13);
          dG = BF_REC(14);
          d1 = BF_REC(15);
          tau = (t > 2 ? T.TerminalAU : 0);
        }
        int sp = BF_INF, fa = BF_INF, fb = BF_INF;
        if (valid) {
          sp = SPL[(D % 3) * NT * 16 + I * 16 + ab];
          if (NWS > 1) {
            sp = PS[I * 16 + ab];
#pragma unroll
            for (int w = 1; w < NWS; w++) sp = min(sp, PS[(size_t)w * NT * 16 + I * 16 + ab]);
          }
          if (sp >= kInfThr) sp = BF_INF;
          if (d > BF_TURN + 1) {  // fML neighbours that live in earlier tile-diagonals
            if (a == 3) fa = FM[(size_t)tile_idx(NT, I + 1, J) * 16 + bq];
            if (bq == 0) fb = FM[(size_t)(tix - 1) * 16 + a * 4 + 3];
          }
        }
        const int o = valid ? tri_off(n, d) + i - 1 : 0;
        const int rrow = valid ? (d % kR) * RS + i : 0;
        for (int e0 = -3; e0 <= 3; e0++) {
          if (valid && e == e0) {
            if (t) {
#pragma unroll
              for (int k = 0; k < 11; k++)
                if ((near_u1(k) <= 2 && near_u2(k) <= 2) && ((same >> k) & 1)) en = min(en, RING[rb[near_u1(k) + near_u2(k)] + near_u1(k)] + BF_REC(k));
              if (d >= 11 && a < 3 && bq > 0) {
                const int dm = SPL[(D % 3) * NT * 16 + I * 16 + ab + 3];  // cell (a+1, bq-1) of this tile
                if (dm < kInfThr) en = min(en, dm + mlclose);
              }
              if (en >= kInfThr) en = BF_INF;
            }
            int m = sp;
            if (en < BF_INF && i > 1 && j < n) m = min(m, en + mlout);
            if (d > BF_TURN + 1) {
              if (a < 3) fa = FM[(size_t)tix * 16 + ab + 4];
              if (bq > 0) fb = FM[(size_t)tix * 16 + ab - 1];
              m = min(m, min(fa, fb) + T.MLbase);
            }
            if (m >= kInfThr) m = BF_INF;
            cg_out[o] = en;
            fg_out[o] = m;
            FM[(size_t)tix * 16 + ab] = m;
            SPL[(D % 3) * NT * 16 + I * 16 + ab] = sp;
            const bool fin = en < BF_INF;
            RING[rrow] = fin ? en + dG : BF_INF;
            RING[kR * RS + rrow] = fin ? en + d1 : BF_INF;
            RING[2 * kR * RS + rrow] = fin ? en + tau : BF_INF;
          }
          __syncwarp();
        }
      }
      TR(3);
      __syncthreads();
    }
#ifdef BF_TILE_TRACE
    first_seq = false;
#endif
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

struct TileCfg { int nw, nws; bool big; };

TileCfg mfe_tile_cfg(int nmax) {
  TileCfg c;
  c.nw = env_int("BF_TILE_NW", 8);
  if (c.nw != 12 && c.nw != 16) c.nw = 8;
  c.nws = env_int("BF_TILE_NWS", nmax > 160 ? 2 : 1);
  if (c.nws != 2) c.nws = 1;
  c.big = tile_plan_i(nmax, c.nws, false).total > (size_t)env_int("BF_TILE_SMEM_MAX", 112 * 1024);
  return c;
}

template <int NW, int NWS, bool BIG>
cudaError_t mfe_tile_t(const BfParams *dP, const BfBatchDev &b, int *ctri, int *ftri, int *ws, size_t ws_slot, int sms, int *grid_out, bool launch,
                       int *counter, cudaStream_t st) {
  auto kern = bf_k_mfe_tile<NW, NWS, BIG>;
  const size_t sm = tile_plan_i(b.stride, NWS, BIG).total;
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
    if (c.big) return mfe_tile_t<NW, 2, true>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
    return mfe_tile_t<NW, 2, false>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
  }
  if (c.big) return mfe_tile_t<NW, 1, true>(dP, b, ctri, ftri, ws, ws_slot, sms, grid_out, launch, counter, st);
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

#ifdef BF_TILE_TRACE
extern "C" int bf_tile_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace)); }
#endif
// 1 if the tile path covers this length (ring rows and tile tables must fit the CTA's shared memory)
int bf_tile_mfe_ok(int nmax) {
  if (nmax < 1 || nmax > 2000) return 0;
  const TileCfg c = mfe_tile_cfg(nmax);
  return tile_plan_i(nmax, c.nws, c.big).total <= kSmemBudget ? 1 : 0;
}
size_t bf_mfe_tile_ws_slot(int nmax) {  // ints of per-CTA HBM workspace (tile-major fML when it is not on chip)
  const TileCfg c = mfe_tile_cfg(nmax);
  return tile_ws_ints(nmax, c.big);
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
