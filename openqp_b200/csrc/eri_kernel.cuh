// Rys-quadrature ERI + fused J/K digestion kernels for sm_100a.
//
// One kernel per canonical angular-momentum class (LA>=LB | LC>=LD), bra class >= ket class, s..f.
// Replaces, for one batch of surviving shell quartets, the reference's per-thread sequence
//   shellquartet -> int2_rys_compute (int_rys.F90:156-276) -> normalize/pure projection
//   (int2.F90:1187-1207, int_rys.F90:715-785) -> storeints (int2.F90:1741-1865) -> consumer update
//   (int2.F90:1414-1578, tdhf_lib.F90:140-224, tdhf_mrsf_lib.F90:218-333).
//
// Kernel families, chosen per class at compile time (launch_eri*) and per launch by the host (run_build):
//   eri_small_kernel   one thread = one quartet; classes with <= 36 Cartesian integrals keep everything in registers,
//                      classes with <= 150 write the gx/gy/gz tables of a root to shared memory ([entry][thread]; ptxas
//                      forwards the stored values, so they live in registers too); WPQ variant: one warp = one quartet;
//   eri_run_kernel     the same evaluation, but a warp walks a run of consecutive surviving kets of ONE bra (bra entry and
//                      D_ab loaded once, J_ab accumulated in registers over the run); SYM consumers with one Fock matrix;
//   eri_group_kernel   a quartet is owned by an aligned group of G = 4..32 lanes of one warp (__syncwarp only), lanes own
//                      BRA component pairs: roots, 2-D VRR + HRR per (root, direction) in registers, tables in shared
//                      memory ([c][d][a][b]), register-tiled assembly, block projected and digested from shared memory;
//   eri_kown_kernel    (eri_kown.cuh) groups of any size whose lanes own KET components with all bra components in
//                      registers: table rows read once (LDS.128, broadcast), bra projection and SYM digestion from registers;
//   eri_kernel         CTA teams (NA*NB*KS threads per quartet, CTA barriers) for the few largest classes.
// Per primitive quartet:  roots/weights (Chebyshev tables) -> 2-D VRR on centres A and C -> HRR to B and D ->
// I += gx*gy*gz.  Then the Cartesian block is normalised / projected to pure functions index by index (sparse
// compile-time tables), the element cutoff and the coincidence factors are applied, and the block is contracted with
// the density: SYM consumers (RHF/UROHF) reduce per output element in registers and leave as one FP64
// red.global.add per Fock element per quartet (after a segmented reduction over quartets of the warp that share the
// output); GEN consumers (TD/MRSF, many general densities) are digested warp-cooperatively with DMMA m8n8k4.
// mu2inv != 0 selects Erf-attenuated integrals (CAM second pass).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>
#include <utility>

// ---- tuning knobs (overridable with -D..., see tools/tune_variants.sh)
#ifndef OQPB_MEDIUM_MAX
#define OQPB_MEDIUM_MAX 150
#endif
#ifndef OQPB_SMALL_REGS
#define OQPB_SMALL_REGS 255
#endif
#ifndef OQPB_MED_REGS
#define OQPB_MED_REGS 255
#endif
#ifndef OQPB_GRP_REGS
#define OQPB_GRP_REGS 255
#endif
#ifndef OQPB_DMMA_MIN_FILL
#define OQPB_DMMA_MIN_FILL 0.35
#endif
#ifndef OQPB_SMALL_GRID
#define OQPB_SMALL_GRID 8
#endif
#ifndef OQPB_DEN_BATCH_MAX
#define OQPB_DEN_BATCH_MAX 96
#endif
#ifndef OQPB_DEN_EARLY_MAX
#define OQPB_DEN_EARLY_MAX 0
#endif
#ifndef OQPB_GRP_LIMIT
#define OQPB_GRP_LIMIT 56
#endif
#ifndef OQPB_RUN_M
#define OQPB_RUN_M 7
#endif
#ifndef OQPB_ROOT_UNROLL
#define OQPB_ROOT_UNROLL 2
#endif
#ifndef OQPB_ROOT_UNROLL_MAX
#define OQPB_ROOT_UNROLL_MAX 18
#endif
#ifndef OQPB_KP_UNROLL
#define OQPB_KP_UNROLL 1
#endif
#ifndef OQPB_KP_UNROLL_MAXR
#define OQPB_KP_UNROLL_MAXR 1
#endif
#ifndef OQPB_GRP_STATIC_WALK
#define OQPB_GRP_STATIC_WALK 0
#endif
#ifndef OQPB_MED_VOLATILE
#define OQPB_MED_VOLATILE 0
#endif
#ifndef OQPB_RUN_SEGC
#define OQPB_RUN_SEGC 1
#endif
#ifndef OQPB_REGVRR_MAX
#define OQPB_REGVRR_MAX 64
#endif

namespace oqpb {
__host__ __device__ constexpr int root_unroll() { return OQPB_ROOT_UNROLL; }
__host__ __device__ constexpr int root_unroll_max() { return OQPB_ROOT_UNROLL_MAX; }

struct alignas(16) PairEntry {
  int sa, sb;      // shells, am(sa) >= am(sb); equal am: sa is the canonical row shell (sa >= sb)
  int poff, pcnt;  // primitive-pair records, sorted by |K|/zeta descending
  double zmin;     // smallest zeta of the pair (lower bound of zeta+eta in the primitive-quartet test)
  double ax, ay, az;     // centre of shell sa
  double abx, aby, abz;  // A - B
  int oa, ob;            // first AO of sa, sb
};

// primitive pair record: Px Py Pz zeta da zinv, da = K/zeta with K = sqrt(2) pi^{5/4} c_a c_b exp(-ab R^2/zeta)
// (int2_pairs.F90:259, int_rys.F90:216); records of a pair are sorted by |da| descending
constexpr int PRIM_STRIDE = 6;

#include "proj_tables.inc"

// pure / Cartesian variant of a kernel: bit 0 = d shells are pure (5d), bit 1 = f shells are pure (7f)
template <int L, int PV>
struct Shell {
  static constexpr bool PURE = (L >= 2) && (((PV >> (L - 2)) & 1) != 0);
  static constexpr int NC = (L + 1) * (L + 2) / 2;
  static constexpr int NOUT = PROJC[L][PURE ? 1 : 0].nout;
};

enum Mode { MODE_SYM = 0, MODE_GEN = 1, MODE_SCHWARZ = 2, MODE_BLOCK = 3 };

constexpr int MAX_MATS = 8;

struct EriArgs {
  const PairEntry* bra;
  const PairEntry* ket;
  const double* prim;
  const double* xyz;
  const int* aooff;
  const int2* tasks;
  const unsigned* ntasks;  // device-resident count
  unsigned task_cap;       // capacity of `tasks` (0 = unbounded): k_enum counts past it, the kernels must not read past it
  unsigned* counter;       // dynamic task fetch
  // run kernels (eri_run_kernel): warp work items = up to RUN_LEN consecutive surviving kets of ONE bra, written by k_enum
  const int2* items;       // x = bra entry | length << 24, y = first task index
  const unsigned* nitems;
  unsigned item_cap;
  const double* rys_tab;   // table base for this nroots
  int rys_xmax;
  double herm_r[7], herm_w[7];
  double prim_cutoff;  // pair_cutoff^2 (int_rys.F90:74,232)
  double mu2inv;       // 1/mu^2 for Erf-attenuated integrals (CAM second pass, int_rys.F90:179-181, 225-227); 0 = regular
  double cutoff;       // element cutoff (int2.F90:1806-1812)
  int mode;
  int nbf;
  // MODE_SYM: packed Fock accumulation, out_m += 4*cj*Jtype[DJ_m] - ck*Ktype[DK_m]   (reference's 6 updates)
  int nmat;
  const double* DJ[MAX_MATS];
  const double* DK[MAX_MATS];
  double* F[MAX_MATS];
  double cj, ck;
  // MODE_GEN: square general densities P_m (row-major [a*nbf+b] = P(a,b)); Jout_m/Kout_m square accumulators
  //   Jout_m(a,b) += cj * v * (P(c,d)+P(d,c)), (c,d) likewise with (a,b);  Kout_m 8-target form with ck.
  //   gen_wantj[m]: whether matrix m takes the Coulomb part.
  int gen_wantj[MAX_MATS];
  // batched GEN: matrices stored interleaved [a*nbf+b][nmat] when gen_interleaved != 0 (MRSF layout)
  int gen_interleaved;
  int gen_nmat_total;  // nmat for interleaved layout (may exceed MAX_MATS): the stride between AO pairs
  int gen_mcount;      // matrices that take the exchange part, starting at Pgen / Fgen (<= gen_nmat_total; the CAM second
                       // pass of the MRSF consumer passes Pgen + 6 nvec and nvec: component 7 only)
  int gen_xoff;        // first matrix that takes the exchange part (matrices [gen_xoff, gen_xoff + gen_mcount)); the generic
                       // J/K entry keeps Coulomb-only matrices in front of it
  int gen_ncoul;       // interleaved: component index < gen_ncoul gets Coulomb (uses comp = m / nvec ... see kernel)
  int gen_nvec;
  const double* Pgen;  // interleaved density
  double* Fgen;        // interleaved output
  unsigned long long* stat;  // [0] += primitive quartets evaluated, [1] += 8 * sum(fac * surviving AO integrals)
  // MODE_SCHWARZ
  double* qout;  // per bra entry
  // MODE_BLOCK
  double* blockout;
};

__host__ __device__ constexpr int ncart(int l) { return (l + 1) * (l + 2) / 2; }

// internal Cartesian order: x descending, then y descending
template <int L>
struct Cart {
  static constexpr int N = (L + 1) * (L + 2) / 2;
  __host__ __device__ static constexpr int x(int c) {
    int k = 0;
    for (int xx = L; xx >= 0; --xx)
      for (int yy = L - xx; yy >= 0; --yy) {
        if (k == c) return xx;
        ++k;
      }
    return 0;
  }
  __host__ __device__ static constexpr int y(int c) {
    int k = 0;
    for (int xx = L; xx >= 0; --xx)
      for (int yy = L - xx; yy >= 0; --yy) {
        if (k == c) return yy;
        ++k;
      }
    return 0;
  }
  __host__ __device__ static constexpr int z(int c) { return L - x(c) - y(c); }
};

template <int I, class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
  (f(std::integral_constant<int, I + Is>{}), ...);
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl<I>(f, std::make_integer_sequence<int, (N > I ? N - I : 0)>{});
}

__device__ __forceinline__ void cart_xyz_rt(int l, int c, int& x, int& y, int& z) {
  int k = 0;
  x = y = z = 0;
  for (int xx = l; xx >= 0; --xx)
    for (int yy = l - xx; yy >= 0; --yy) {
      if (k == c) { x = xx; y = yy; z = l - xx - yy; }
      ++k;
    }
}

template <int LA, int LB, int LC, int LD>
struct ClassCfg {
  static constexpr int NA = ncart(LA), NB = ncart(LB), NC = ncart(LC), ND = ncart(LD);
  static constexpr int R = (LA + LB + LC + LD) / 2 + 1;
  static constexpr int NMAX = LA + LB + 1, MMAX = LC + LD + 1;
  static constexpr int NIJ1 = (LA + 1) * (LB + 1), NKL1 = (LC + 1) * (LD + 1);
  static constexpr int NKET = NC * ND;
  static constexpr int KS = (NKET + 39) / 40;           // ket slices (<= 40 accumulators / thread)
  static constexpr int TKC = (NC + KS - 1) / KS;        // ket-c components per slice
  static constexpr int NACC = TKC * ND;
  static constexpr int TS = NA * NB * KS;               // threads per quartet
  static constexpr int G1 = NMAX * MMAX, G2 = NMAX * NKL1, G3 = NIJ1 * NKL1;
  static constexpr int GSTR = (G1 + G2 + G3) | 1;       // odd stride: conflict-free across (root,dir) tasks
  static constexpr int GREG = 3 * R * GSTR;
  static constexpr int NCART4 = NA * NB * NC * ND;
  static constexpr int BLK = 2 * NCART4;                // ping-pong for the index-wise projection
  static constexpr int QSM0 = (GREG > BLK ? GREG : BLK) + 2 * R + 2;
  static constexpr int QSM = QSM0 | 1;                  // doubles per quartet, odd
  static constexpr int QPB_T = (64 / TS) > 0 ? (64 / TS) : 1;  // small CTAs: lock-step over few quartets, many CTAs/SM
  static constexpr int QPB_S = (12288 / QSM) > 0 ? (12288 / QSM) : 1;  // <= 96 KB of dynamic smem per CTA
  static constexpr int QPB = QPB_T < QPB_S ? QPB_T : QPB_S;
  static constexpr int NT = ((TS * QPB + 31) / 32) * 32;
  static constexpr size_t SMEM = (size_t)QPB * QSM * sizeof(double) + QPB * (sizeof(int) * 24 + 64 * sizeof(unsigned short));
};

struct QInfo {  // 64 ints per quartet in shared memory
  int sa, sb, sc, sd;
  int boff, bcnt, koff, kcnt;
  int oa, ob, oc, od;
  int valid, nonzero, imax, jmax;
  float fac;
  int bra_id, ket_id, lcount;
};

// ---------------------------------------------------------------------------------------------------
// roots / weights for function f (f < R: root t^2, f >= R: weight) at X
// Table format per nroots: unit X intervals with 12-term polynomial fits (Chebyshev fits of tools/gen_rys_tables.py,
// converted to monomial coefficients in t in [-1, 1] when the context is created); nroots = 1 ((ss|ss), (ps|ss): the most
// heavily contracted classes, where the interpolation is a third of the FP64 work) uses quarter-width intervals with
// 8 terms (tools/gen_rys_tables_fine.py, fit error < 2e-15): -4 % on those classes.  The same format for nroots = 2
// was measured slower: its 51 KB table per CTA eats the L1 of the 128-register classes.
constexpr int RYS_FINE_MAXR = 1;
template <int R>
struct RysFmt {
  static constexpr int NC = R <= RYS_FINE_MAXR ? 8 : 12;  // coefficients per function and interval
  static constexpr int DIV = R <= RYS_FINE_MAXR ? 4 : 1;  // intervals per unit of X
};
template <int R>
__device__ __forceinline__ double rys_eval(const EriArgs& a, double X, int f) {
  if (X >= (double)a.rys_xmax) {
    // half-range Gauss-Hermite asymptote (rys.F90:2711-2713)
    if (f < R) return a.herm_r[f] / X;
    return a.herm_w[f - R] * rsqrt(X);
  }
  constexpr int NCF = RysFmt<R>::NC;
  const double xs = X * RysFmt<R>::DIV;
  int iv = (int)xs;
  double t = 2.0 * (xs - (double)iv) - 1.0;
  const double* c = a.rys_tab + ((size_t)iv * (2 * R) + f) * NCF;
  // Horner on the monomial coefficients (converted from the Chebyshev fits at context creation, oqp_b200.cu
  // cheb_to_monomial): one DFMA per term instead of the DADD + DFMA of a Clenshaw step, same accuracy (2.3e-16 rel.)
  double b = __ldg(c + NCF - 1);
#pragma unroll
  for (int k = NCF - 2; k >= 0; --k) b = fma(b, t, __ldg(c + k));
  return b;
}

// ---------------------------------------------------------------------------------------------------
// Normalisation + Cartesian -> pure projection of ONE index of a 4-index block, tables known at compile time
// (int2.F90:1187-1207; int_rys.F90:715-785; proj_tables.inc).  Tensor layout (OUTER, NIN, INNER) -> (OUTER, NOUT, INNER).
// (1) block in registers (thread-per-quartet kernels): every index is a constant expression
template <int L, bool PURE, int OUTER, int INNER, int NI, int NO>
__device__ __forceinline__ void proj_reg(const double (&in)[NI], double (&out)[NO]) {
  constexpr int NIN = (L + 1) * (L + 2) / 2, NOUT = PROJC[L][PURE ? 1 : 0].nout;
  static_assert(NI == OUTER * NIN * INNER && NO == OUTER * NOUT * INNER, "proj_reg: shape");
  if constexpr (L < 2) {
#pragma unroll
    for (int e = 0; e < NO; ++e) out[e] = in[e];
  } else {
    static_for<0, NO>([&](auto E) {
      constexpr int e = decltype(E)::value;
      constexpr int i = e % INNER, o = (e / INNER) % NOUT, ou = e / (INNER * NOUT);
      constexpr int nt = PROJC[L][PURE ? 1 : 0].nterm[o];
      double sum = 0.0;
      static_for<0, nt>([&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int j = PROJC[L][PURE ? 1 : 0].idx[o][k];
        constexpr double c = PROJC[L][PURE ? 1 : 0].coef[o][k];
        sum = fma(c, in[(ou * NIN + j) * INNER + i], sum);
      });
      out[e] = sum;
    });
  }
}
// (2) block in shared memory (team kernels): a lane transforms whole NIN-vectors, coefficients are immediates
template <int L, bool PURE, int OUTER, int INNER>
__device__ __forceinline__ void proj_smem(const double* __restrict__ in, double* __restrict__ out, int t, int ts) {
  constexpr int NIN = (L + 1) * (L + 2) / 2, NOUT = PROJC[L][PURE ? 1 : 0].nout;
  for (int e = t; e < OUTER * INNER; e += ts) {
    const int i = e % INNER, ou = e / INNER;
    const double* src = in + (size_t)ou * NIN * INNER + i;
    double* dst = out + (size_t)ou * NOUT * INNER + i;
    double x[NIN];
#pragma unroll
    for (int j = 0; j < NIN; ++j) x[j] = src[j * INNER];
    static_for<0, NOUT>([&](auto O) {
      constexpr int o = decltype(O)::value;
      constexpr int nt = PROJC[L][PURE ? 1 : 0].nterm[o];
      double sum = 0.0;
      static_for<0, nt>([&](auto K) {
        constexpr int k = decltype(K)::value;
        constexpr int j = PROJC[L][PURE ? 1 : 0].idx[o][k];
        constexpr double c = PROJC[L][PURE ? 1 : 0].coef[o][k];
        sum = fma(c, x[j], sum);
      });
      dst[o * INNER] = sum;
    });
  }
}

__device__ __forceinline__ size_t tri_idx(int p, int q) {
  return p >= q ? (size_t)p * (p + 1) / 2 + q : (size_t)q * (q + 1) / 2 + p;
}
__device__ __forceinline__ unsigned tri_u(unsigned p, unsigned q) {  // nbf <= 46000: p(p+1) < 2^32
  return p >= q ? p * (p + 1) / 2 + q : q * (q + 1) / 2 + p;
}

// ---------------------------------------------------------------------------------------------------
// Digestion of one finished block blk[a][b][c][d] (d fastest), compile-time dims, AO offsets o0..o3.
// MODE_SYM: the reference's six packed updates (int2.F90:1414-1484 / 1488-1578) on the full block
// with the shell-level coincidence factor already applied (equivalent to the unique-AO walk + AO-level halving
// of storeints, int2.F90:1769-1851).  One FP64 red per Fock element per quartet.
// Warp-segmented reduction for the thread-per-quartet kernels.  The enumeration writes the surviving kets of a bra
// contiguously and the pair lists are ordered by the ket's first shell inside a Schwarz bin, so the 32 quartets of
// a warp form runs with the same bra (same J_ab targets) and, inside those, runs with the same ket shell c (same
// K_ac / K_bc targets).  A run is summed with shuffles and its first lane issues the one red.global.add.
struct SegMask {
  unsigned char up[5];  // lane + 2^k belongs to the same run
  bool head;            // first lane of its run
};
__device__ __forceinline__ SegMask seg_make(long long key, int lane) {
  SegMask m;
  const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
  m.head = lane == 0 || prev != key;
  // run index = number of heads at or before this lane: equal keys that are not adjacent stay separate runs
  const unsigned heads = __ballot_sync(0xffffffffu, m.head);
  const int rid = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int o = __shfl_down_sync(0xffffffffu, rid, 1 << k);
    m.up[k] = (lane + (1 << k) < 32) && o == rid;
  }
  return m;
}
__device__ __forceinline__ double seg_sum(double v, const SegMask& m) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double o = __shfl_down_sync(0xffffffffu, v, 1 << k);
    if (m.up[k]) v += o;
  }
  return v;  // run total in the head lane
}

// (1) block in registers, one thread: everything unrolled.  The six density sub-blocks a quartet needs are loaded
// in ONE batch (DenBlk) so that their L2 latencies overlap; the small classes issue the batch before the primitive loop.
template <int N0, int N1, int N2, int N3>
struct DenBlk {
  static constexpr int SIZE = N2 * N3 + N0 * N1 + N1 * N3 + N1 * N2 + N0 * N3 + N0 * N2;
  double cd[N2 * N3], ab[N0 * N1], bd[N1 * N3], bc[N1 * N2], ad[N0 * N3], ac[N0 * N2];
};
template <int N0, int N1, int N2, int N3>
__device__ __forceinline__ void den_load(const EriArgs& A, int m, int o0, int o1, int o2, int o3, DenBlk<N0, N1, N2, N3>& D) {
  const unsigned nbf = (unsigned)A.nbf;
  const double* __restrict__ DJ = A.DJ[m];
  const double* __restrict__ DK = A.DK[m];
#pragma unroll
  for (int c = 0; c < N2; ++c)
#pragma unroll
    for (int d = 0; d < N3; ++d) D.cd[c * N3 + d] = __ldg(DJ + ((unsigned)(o2 + c) * nbf + (unsigned)(o3 + d)));
#pragma unroll
  for (int a = 0; a < N0; ++a)
#pragma unroll
    for (int b = 0; b < N1; ++b) D.ab[a * N1 + b] = __ldg(DJ + ((unsigned)(o0 + a) * nbf + (unsigned)(o1 + b)));
#pragma unroll
  for (int b = 0; b < N1; ++b)
#pragma unroll
    for (int d = 0; d < N3; ++d) D.bd[b * N3 + d] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o3 + d)));
#pragma unroll
  for (int b = 0; b < N1; ++b)
#pragma unroll
    for (int c = 0; c < N2; ++c) D.bc[b * N2 + c] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o2 + c)));
#pragma unroll
  for (int a = 0; a < N0; ++a)
#pragma unroll
    for (int d = 0; d < N3; ++d) D.ad[a * N3 + d] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o3 + d)));
#pragma unroll
  for (int a = 0; a < N0; ++a)
#pragma unroll
    for (int c = 0; c < N2; ++c) D.ac[a * N2 + c] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o2 + c)));
}
// All 32 lanes of the warp must call this together (lanes without a quartet pass a zero block).
// BATCH: all six sub-blocks in registers at once (pre = batch of matrix 0 loaded by the caller when have_pre).
template <int N0, int N1, int N2, int N3, bool SEGC, bool BATCH>
__device__ __forceinline__ void digest_sym_reg(const EriArgs& A, const double (&v)[N0 * N1 * N2 * N3], int o0, int o1,
                                               int o2, int o3, const SegMask& mbra, const SegMask& mc,
                                               const DenBlk<N0, N1, N2, N3>& pre, bool have_pre) {
  const unsigned nbf = (unsigned)A.nbf;
  const double c4 = 4.0 * A.cj, c1 = A.ck;
#define VV(a, b, c, d) v[(((a)*N1 + (b)) * N2 + (c)) * N3 + (d)]
  for (int m = 0; m < A.nmat; ++m) {
    const double* __restrict__ DJ = A.DJ[m];
    const double* __restrict__ DK = A.DK[m];
    double* __restrict__ F = A.F[m];
    DenBlk<N0, N1, N2, N3> D;
    if constexpr (BATCH) {
      if (m == 0 && have_pre) D = pre;
      else den_load<N0, N1, N2, N3>(A, m, o0, o1, o2, o3, D);
    }
    {  // J_ab += 4 cj sum_cd v D_cd
      if constexpr (!BATCH) {
#pragma unroll
        for (int c = 0; c < N2; ++c)
#pragma unroll
          for (int d = 0; d < N3; ++d) D.cd[c * N3 + d] = __ldg(DJ + ((unsigned)(o2 + c) * nbf + (unsigned)(o3 + d)));
      }
#pragma unroll
      for (int a = 0; a < N0; ++a)
#pragma unroll
        for (int b = 0; b < N1; ++b) {
          double sum = 0.0;
#pragma unroll
          for (int c = 0; c < N2; ++c)
#pragma unroll
            for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), D.cd[c * N3 + d], sum);
          sum = seg_sum(sum, mbra);
          if (mbra.head && sum != 0.0) atomicAdd(F + tri_u(o0 + a, o1 + b), c4 * sum);
        }
    }
    {  // J_cd += 4 cj sum_ab v D_ab
      if constexpr (!BATCH) {
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int b = 0; b < N1; ++b) D.ab[a * N1 + b] = __ldg(DJ + ((unsigned)(o0 + a) * nbf + (unsigned)(o1 + b)));
      }
#pragma unroll
      for (int c = 0; c < N2; ++c)
#pragma unroll
        for (int d = 0; d < N3; ++d) {
          double sum = 0.0;
#pragma unroll
          for (int a = 0; a < N0; ++a)
#pragma unroll
            for (int b = 0; b < N1; ++b) sum = fma(VV(a, b, c, d), D.ab[a * N1 + b], sum);
          if (sum != 0.0) atomicAdd(F + tri_u(o2 + c, o3 + d), c4 * sum);
        }
    }
    {  // K_ac -= ck sum_bd v D_bd
      if constexpr (!BATCH) {
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
          for (int d = 0; d < N3; ++d) D.bd[b * N3 + d] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o3 + d)));
      }
#pragma unroll
      for (int a = 0; a < N0; ++a)
#pragma unroll
        for (int c = 0; c < N2; ++c) {
          double sum = 0.0;
#pragma unroll
          for (int b = 0; b < N1; ++b)
#pragma unroll
            for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), D.bd[b * N3 + d], sum);
          if constexpr (SEGC) sum = seg_sum(sum, mc);
          if ((!SEGC || mc.head) && sum != 0.0) atomicAdd(F + tri_u(o0 + a, o2 + c), -c1 * sum);
        }
    }
    {  // K_ad -= ck sum_bc v D_bc
      if constexpr (!BATCH) {
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
          for (int c = 0; c < N2; ++c) D.bc[b * N2 + c] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o2 + c)));
      }
#pragma unroll
      for (int a = 0; a < N0; ++a)
#pragma unroll
        for (int d = 0; d < N3; ++d) {
          double sum = 0.0;
#pragma unroll
          for (int b = 0; b < N1; ++b)
#pragma unroll
            for (int c = 0; c < N2; ++c) sum = fma(VV(a, b, c, d), D.bc[b * N2 + c], sum);
          if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o3 + d), -c1 * sum);
        }
    }
    {  // K_bc -= ck sum_ad v D_ad
      if constexpr (!BATCH) {
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int d = 0; d < N3; ++d) D.ad[a * N3 + d] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o3 + d)));
      }
#pragma unroll
      for (int b = 0; b < N1; ++b)
#pragma unroll
        for (int c = 0; c < N2; ++c) {
          double sum = 0.0;
#pragma unroll
          for (int a = 0; a < N0; ++a)
#pragma unroll
            for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), D.ad[a * N3 + d], sum);
          if constexpr (SEGC) sum = seg_sum(sum, mc);
          if ((!SEGC || mc.head) && sum != 0.0) atomicAdd(F + tri_u(o1 + b, o2 + c), -c1 * sum);
        }
    }
    {  // K_bd -= ck sum_ac v D_ac
      if constexpr (!BATCH) {
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int c = 0; c < N2; ++c) D.ac[a * N2 + c] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o2 + c)));
      }
#pragma unroll
      for (int b = 0; b < N1; ++b)
#pragma unroll
        for (int d = 0; d < N3; ++d) {
          double sum = 0.0;
#pragma unroll
          for (int a = 0; a < N0; ++a)
#pragma unroll
            for (int c = 0; c < N2; ++c) sum = fma(VV(a, b, c, d), D.ac[a * N2 + c], sum);
          if (sum != 0.0) atomicAdd(F + tri_u(o1 + b, o3 + d), -c1 * sum);
        }
    }
  }
#undef VV
}

// (2) block in shared memory, outputs strided over the ts lanes of a team
template <int N0, int N1, int N2, int N3>
__device__ __forceinline__ void digest_sym(const EriArgs& A, const double* blk, int o0, int o1, int o2, int o3, int t,
                                           int ts) {
  const unsigned nbf = (unsigned)A.nbf;
  constexpr int N23 = N2 * N3;
  const double c4 = 4.0 * A.cj, c1 = A.ck;
  for (int m = 0; m < A.nmat; ++m) {
    const double* __restrict__ DJ = A.DJ[m];
    const double* __restrict__ DK = A.DK[m];
    double* __restrict__ F = A.F[m];
    // J_ab += 4 cj sum_cd v D_cd
    for (int o = t; o < N0 * N1; o += ts) {
      const int a = o / N1, b = o % N1;
      const double* v = blk + o * N23;
      double sum = 0.0;
#pragma unroll
      for (int c = 0; c < N2; ++c) {
        const double* drow = DJ + ((unsigned)(o2 + c) * nbf + (unsigned)o3);
#pragma unroll
        for (int d = 0; d < N3; ++d) sum = fma(v[c * N3 + d], __ldg(drow + d), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o1 + b), c4 * sum);
    }
    // J_cd += 4 cj sum_ab v D_ab
    for (int o = t; o < N23; o += ts) {
      const int c = o / N3, d = o % N3;
      double sum = 0.0;
#pragma unroll
      for (int a = 0; a < N0; ++a) {
        const double* drow = DJ + ((unsigned)(o0 + a) * nbf + (unsigned)o1);
#pragma unroll
        for (int b = 0; b < N1; ++b) sum = fma(blk[(a * N1 + b) * N23 + o], __ldg(drow + b), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o2 + c, o3 + d), c4 * sum);
    }
    // K_ac -= ck sum_bd v D_bd
    for (int o = t; o < N0 * N2; o += ts) {
      const int a = o / N2, c = o % N2;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        const double* drow = DK + ((unsigned)(o1 + b) * nbf + (unsigned)o3);
        const double* v = blk + (a * N1 + b) * N23 + c * N3;
#pragma unroll
        for (int d = 0; d < N3; ++d) sum = fma(v[d], __ldg(drow + d), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o2 + c), -c1 * sum);
    }
    // K_ad -= ck sum_bc v D_bc
    for (int o = t; o < N0 * N3; o += ts) {
      const int a = o / N3, d = o % N3;
      double sum = 0.0;
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        const double* drow = DK + ((unsigned)(o1 + b) * nbf + (unsigned)o2);
        const double* v = blk + (a * N1 + b) * N23 + d;
#pragma unroll
        for (int c = 0; c < N2; ++c) sum = fma(v[c * N3], __ldg(drow + c), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o3 + d), -c1 * sum);
    }
    // K_bc -= ck sum_ad v D_ad
    for (int o = t; o < N1 * N2; o += ts) {
      const int b = o / N2, c = o % N2;
      double sum = 0.0;
#pragma unroll
      for (int a = 0; a < N0; ++a) {
        const double* drow = DK + ((unsigned)(o0 + a) * nbf + (unsigned)o3);
        const double* v = blk + (a * N1 + b) * N23 + c * N3;
#pragma unroll
        for (int d = 0; d < N3; ++d) sum = fma(v[d], __ldg(drow + d), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o1 + b, o2 + c), -c1 * sum);
    }
    // K_bd -= ck sum_ac v D_ac
    for (int o = t; o < N1 * N3; o += ts) {
      const int b = o / N3, d = o % N3;
      double sum = 0.0;
#pragma unroll
      for (int a = 0; a < N0; ++a) {
        const double* drow = DK + ((unsigned)(o0 + a) * nbf + (unsigned)o2);
        const double* v = blk + (a * N1 + b) * N23 + d;
#pragma unroll
        for (int c = 0; c < N2; ++c) sum = fma(v[c * N3], __ldg(drow + c), sum);
      }
      if (sum != 0.0) atomicAdd(F + tri_u(o1 + b, o3 + d), -c1 * sum);
    }
  }
}

// (3) group kernel: block in shared memory, a quartet is owned by G lanes and a warp holds 32/G quartets.  Same
// updates as (2); J_ab and the K_ac / K_bc sums of quartets of the warp that share the bra (and the ket shell c)
// are first added across the quartet slots with shuffles (stride G), so that one lane issues the red.
template <int G>
struct GroupSeg {
  static constexpr int QPW = 32 / G;
  static constexpr int NST = QPW == 1 ? 0 : (QPW == 2 ? 1 : (QPW == 4 ? 2 : 3));
  bool up[NST > 0 ? NST : 1];
  bool head;
};
template <int G>
__device__ __forceinline__ GroupSeg<G> gseg_make(long long key, int lane) {
  GroupSeg<G> m;
  m.head = true;
  if constexpr (GroupSeg<G>::NST > 0) {
    const int g = lane / G;
    const long long prev = __shfl_up_sync(0xffffffffu, key, G);
    m.head = g == 0 || prev != key;
    const unsigned heads = __ballot_sync(0xffffffffu, m.head && (lane % G) == 0);
    const int rid = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int k = 0; k < GroupSeg<G>::NST; ++k) {
      const int o = __shfl_down_sync(0xffffffffu, rid, G << k);
      m.up[k] = (g + (1 << k) < GroupSeg<G>::QPW) && o == rid;
    }
  }
  return m;
}
template <int G>
__device__ __forceinline__ double gseg_sum(double v, const GroupSeg<G>& m) {
  if constexpr (GroupSeg<G>::NST > 0) {
#pragma unroll
    for (int k = 0; k < GroupSeg<G>::NST; ++k) {
      const double o = __shfl_down_sync(0xffffffffu, v, G << k);
      if (m.up[k]) v += o;
    }
  }
  return v;
}
// all 32 lanes call this together; `work` = this lane's quartet has a non-zero block
template <int N0, int N1, int N2, int N3, int G>
__device__ __forceinline__ void digest_sym_group(const EriArgs& A, const double* blk, int o0, int o1, int o2, int o3,
                                                 int t, bool work, const GroupSeg<G>& mbra, const GroupSeg<G>& mc) {
  const unsigned nbf = (unsigned)A.nbf;
  constexpr int N23 = N2 * N3;
  const double c4 = 4.0 * A.cj, c1 = A.ck;
  for (int m = 0; m < A.nmat; ++m) {
    const double* __restrict__ DJ = A.DJ[m];
    const double* __restrict__ DK = A.DK[m];
    double* __restrict__ F = A.F[m];
    // J_ab += 4 cj sum_cd v D_cd
    for (int ob = 0; ob < N0 * N1; ob += G) {
      const int o = ob + t;
      const bool act = work && o < N0 * N1;
      const int a = o / N1, b = o % N1;
      double sum = 0.0;
      if (act) {
        const double* v = blk + o * N23;
#pragma unroll
        for (int c = 0; c < N2; ++c) {
          const double* drow = DJ + ((unsigned)(o2 + c) * nbf + (unsigned)o3);
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(v[c * N3 + d], __ldg(drow + d), sum);
        }
      }
      sum = gseg_sum<G>(sum, mbra);
      if (act && mbra.head && sum != 0.0) atomicAdd(F + tri_u(o0 + a, o1 + b), c4 * sum);
    }
    // J_cd += 4 cj sum_ab v D_ab
    for (int ob = 0; ob < N23; ob += G) {
      const int o = ob + t;
      if (work && o < N23) {
        const int c = o / N3, d = o % N3;
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < N0; ++a) {
          const double* drow = DJ + ((unsigned)(o0 + a) * nbf + (unsigned)o1);
#pragma unroll
          for (int b = 0; b < N1; ++b) sum = fma(blk[(a * N1 + b) * N23 + o], __ldg(drow + b), sum);
        }
        if (sum != 0.0) atomicAdd(F + tri_u(o2 + c, o3 + d), c4 * sum);
      }
    }
    // K_ac -= ck sum_bd v D_bd
    for (int ob = 0; ob < N0 * N2; ob += G) {
      const int o = ob + t;
      const bool act = work && o < N0 * N2;
      const int a = o / N2, c = o % N2;
      double sum = 0.0;
      if (act) {
#pragma unroll
        for (int b = 0; b < N1; ++b) {
          const double* drow = DK + ((unsigned)(o1 + b) * nbf + (unsigned)o3);
          const double* v = blk + (a * N1 + b) * N23 + c * N3;
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(v[d], __ldg(drow + d), sum);
        }
      }
      sum = gseg_sum<G>(sum, mc);
      if (act && mc.head && sum != 0.0) atomicAdd(F + tri_u(o0 + a, o2 + c), -c1 * sum);
    }
    // K_ad -= ck sum_bc v D_bc
    for (int ob = 0; ob < N0 * N3; ob += G) {
      const int o = ob + t;
      if (work && o < N0 * N3) {
        const int a = o / N3, d = o % N3;
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b) {
          const double* drow = DK + ((unsigned)(o1 + b) * nbf + (unsigned)o2);
          const double* v = blk + (a * N1 + b) * N23 + d;
#pragma unroll
          for (int c = 0; c < N2; ++c) sum = fma(v[c * N3], __ldg(drow + c), sum);
        }
        if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o3 + d), -c1 * sum);
      }
    }
    // K_bc -= ck sum_ad v D_ad
    for (int ob = 0; ob < N1 * N2; ob += G) {
      const int o = ob + t;
      const bool act = work && o < N1 * N2;
      const int b = o / N2, c = o % N2;
      double sum = 0.0;
      if (act) {
#pragma unroll
        for (int a = 0; a < N0; ++a) {
          const double* drow = DK + ((unsigned)(o0 + a) * nbf + (unsigned)o3);
          const double* v = blk + (a * N1 + b) * N23 + c * N3;
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(v[d], __ldg(drow + d), sum);
        }
      }
      sum = gseg_sum<G>(sum, mc);
      if (act && mc.head && sum != 0.0) atomicAdd(F + tri_u(o1 + b, o2 + c), -c1 * sum);
    }
    // K_bd -= ck sum_ac v D_ac
    for (int ob = 0; ob < N1 * N3; ob += G) {
      const int o = ob + t;
      if (work && o < N1 * N3) {
        const int b = o / N3, d = o % N3;
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < N0; ++a) {
          const double* drow = DK + ((unsigned)(o0 + a) * nbf + (unsigned)o2);
          const double* v = blk + (a * N1 + b) * N23 + d;
#pragma unroll
          for (int c = 0; c < N2; ++c) sum = fma(v[c * N3], __ldg(drow + c), sum);
        }
        if (sum != 0.0) atomicAdd(F + tri_u(o1 + b, o3 + d), -c1 * sum);
      }
    }
  }
}

// MODE_GEN: general (non-symmetric) densities, all 8 permutations (tdhf_lib.F90:173-186,
// tdhf_mrsf_lib.F90:279-310).  Matrices interleaved in the reference's MRSF layout d3(m, mu, nu):
// X[(nu*nbf + mu)*NM + m] = X_m(mu,nu).
//   Coulomb (m with comp < ncoul): F(a,b),F(b,a) += cj v (P(c,d)+P(d,c)); F(c,d),F(d,c) += cj v (P(a,b)+P(b,a))
//   Exchange (all m): F(a,c) -= ck v P(b,d); F(c,a) -= ck v P(d,b); F(a,d) -= ck v P(b,c); F(d,a) -= ck v P(c,b);
//                     F(b,c) -= ck v P(a,d); F(c,b) -= ck v P(d,a); F(b,d) -= ck v P(a,c); F(d,b) -= ck v P(c,a)
// Threads are spread over (output element, matrix): m fastest -> coalesced density reads and atomics.
template <int n0, int n1, int n2, int n3>
__device__ __forceinline__ void digest_gen(const EriArgs& A, const double* blk, int o0, int o1, int o2, int o3, int t,
                                           int ts) {
  const int nbf = A.nbf, NM = A.gen_nmat_total, MC = A.gen_mcount, nv = A.gen_nvec;
  constexpr int n23 = n2 * n3;
  const double* __restrict__ P = A.Pgen;
  double* __restrict__ F = A.Fgen;
  const int ncm = A.gen_ncoul * nv;  // interleaved index m = comp*nvec + v ... Coulomb for m < ncoul*nvec
  const double cj = A.cj, ck = A.ck;
#define PX(p, q) __ldg(P + ((size_t)(q)*nbf + (p)) * NM + m)
#define FX(p, q, val) atomicAdd(F + ((size_t)(q)*nbf + (p)) * NM + m, (val))
  if (ncm > 0 && cj != 0.0) {
    for (int e = t; e < n0 * n1 * ncm; e += ts) {
      int m = e % ncm, o = e / ncm;
      int a = o / n1, b = o % n1;
      const double* v = blk + (size_t)o * n23;
      double s = 0.0;
      for (int c = 0; c < n2; ++c)
        for (int d = 0; d < n3; ++d) s = fma(v[c * n3 + d], PX(o2 + c, o3 + d) + PX(o3 + d, o2 + c), s);
      if (s != 0.0) {
        FX(o0 + a, o1 + b, cj * s);
        FX(o1 + b, o0 + a, cj * s);
      }
    }
    for (int e = t; e < n23 * ncm; e += ts) {
      int m = e % ncm, o = e / ncm;
      int c = o / n3, d = o % n3;
      double s = 0.0;
      for (int a = 0; a < n0; ++a)
        for (int b = 0; b < n1; ++b)
          s = fma(blk[(size_t)(a * n1 + b) * n23 + o], PX(o0 + a, o1 + b) + PX(o1 + b, o0 + a), s);
      if (s != 0.0) {
        FX(o2 + c, o3 + d, cj * s);
        FX(o3 + d, o2 + c, cj * s);
      }
    }
  }
  if (ck != 0.0) {
    for (int e = t; e < n0 * n2 * MC; e += ts) {  // (a,c)
      int m = e % MC + A.gen_xoff, o = e / MC;
      int a = o / n2, c = o % n2;
      double s1 = 0.0, s2 = 0.0;
      for (int b = 0; b < n1; ++b) {
        const double* v = blk + (size_t)(a * n1 + b) * n23 + c * n3;
        for (int d = 0; d < n3; ++d) {
          s1 = fma(v[d], PX(o1 + b, o3 + d), s1);
          s2 = fma(v[d], PX(o3 + d, o1 + b), s2);
        }
      }
      if (s1 != 0.0) FX(o0 + a, o2 + c, -ck * s1);
      if (s2 != 0.0) FX(o2 + c, o0 + a, -ck * s2);
    }
    for (int e = t; e < n0 * n3 * MC; e += ts) {  // (a,d)
      int m = e % MC + A.gen_xoff, o = e / MC;
      int a = o / n3, d = o % n3;
      double s1 = 0.0, s2 = 0.0;
      for (int b = 0; b < n1; ++b) {
        const double* v = blk + (size_t)(a * n1 + b) * n23 + d;
        for (int c = 0; c < n2; ++c) {
          s1 = fma(v[c * n3], PX(o1 + b, o2 + c), s1);
          s2 = fma(v[c * n3], PX(o2 + c, o1 + b), s2);
        }
      }
      if (s1 != 0.0) FX(o0 + a, o3 + d, -ck * s1);
      if (s2 != 0.0) FX(o3 + d, o0 + a, -ck * s2);
    }
    for (int e = t; e < n1 * n2 * MC; e += ts) {  // (b,c)
      int m = e % MC + A.gen_xoff, o = e / MC;
      int b = o / n2, c = o % n2;
      double s1 = 0.0, s2 = 0.0;
      for (int a = 0; a < n0; ++a) {
        const double* v = blk + (size_t)(a * n1 + b) * n23 + c * n3;
        for (int d = 0; d < n3; ++d) {
          s1 = fma(v[d], PX(o0 + a, o3 + d), s1);
          s2 = fma(v[d], PX(o3 + d, o0 + a), s2);
        }
      }
      if (s1 != 0.0) FX(o1 + b, o2 + c, -ck * s1);
      if (s2 != 0.0) FX(o2 + c, o1 + b, -ck * s2);
    }
    for (int e = t; e < n1 * n3 * MC; e += ts) {  // (b,d)
      int m = e % MC + A.gen_xoff, o = e / MC;
      int b = o / n3, d = o % n3;
      double s1 = 0.0, s2 = 0.0;
      for (int a = 0; a < n0; ++a) {
        const double* v = blk + (size_t)(a * n1 + b) * n23 + d;
        for (int c = 0; c < n2; ++c) {
          s1 = fma(v[c * n3], PX(o0 + a, o2 + c), s1);
          s2 = fma(v[c * n3], PX(o2 + c, o0 + a), s2);
        }
      }
      if (s1 != 0.0) FX(o1 + b, o3 + d, -ck * s1);
      if (s2 != 0.0) FX(o3 + d, o1 + b, -ck * s2);
    }
  }
#undef PX
#undef FX
}

// ---------------------------------------------------------------------------------------------------
// MODE_GEN, warp-cooperative: one finished block in shared memory is digested by all 32 lanes for ALL matrices.
// A contraction  out(m, o) = sum_k P_m(k) V(k, o)  (o = output AO pair, k = contracted AO pair, m = matrix) is a small
// dense GEMM with M = number of matrices (nvec x 7 for MRSF, tdhf_mrsf_lib.F90:279-310): it runs on the FP64 tensor
// cores (DMMA m8n8k4; A = density rows, contiguous in m; B = block elements from shared memory) when the (o, k) tile
// is reasonably full, otherwise as a scalar loop with lanes over (o, m), m fastest (coalesced P reads and reds).
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// X < Y: positions (0..3 = a,b,c,d) of the output pair; the other two positions (U < V) are contracted.
template <int N0, int N1, int N2, int N3, int X, int Y>
__device__ __forceinline__ void gen_contract(const EriArgs& A, const double* __restrict__ blk, const int (&off)[4], int lane) {
  constexpr int N[4] = {N0, N1, N2, N3};
  constexpr int STR[4] = {N1 * N2 * N3, N2 * N3, N3, 1};
  constexpr bool COUL = (X == 0 && Y == 1) || (X == 2 && Y == 3);
  constexpr int U = (X != 0 && Y != 0) ? 0 : ((X != 1 && Y != 1) ? 1 : 2);
  constexpr int V = (X != 3 && Y != 3) ? 3 : ((X != 2 && Y != 2) ? 2 : 1);
  static_assert(U < V && U != X && U != Y && V != X && V != Y, "gen_contract: index positions");
  constexpr int NO = N[X] * N[Y], NK = N[U] * N[V];
  constexpr int KT = (NK + 3) / 4, OT = (NO + 7) / 8;
  constexpr bool USE_MMA = (double)(NO * NK) / (double)(OT * 8 * KT * 4) >= OQPB_DMMA_MIN_FILL;
  const int NM = A.gen_nmat_total;
  const int mrows = COUL ? A.gen_ncoul * A.gen_nvec : A.gen_mcount;
  const double scale = COUL ? A.cj : -A.ck;
  if (mrows <= 0 || scale == 0.0) return;
  const size_t nbf = (size_t)A.nbf;
  const double* __restrict__ P = A.Pgen + (COUL ? 0 : A.gen_xoff);
  double* __restrict__ F = A.Fgen + (COUL ? 0 : A.gen_xoff);
  if constexpr (USE_MMA) {
    const int kl = lane & 3, rl = lane >> 2;
    size_t pa1[KT], pa2[KT];
    int vofs[KT];
    bool kok[KT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const int k = kt * 4 + kl;
      kok[kt] = k < NK;
      const int kk = kok[kt] ? k : 0;
      const int iu = kk / N[V], iv = kk % N[V];
      const size_t p = (size_t)(off[U] + iu), q = (size_t)(off[V] + iv);
      pa1[kt] = (q * nbf + p) * NM;
      pa2[kt] = (p * nbf + q) * NM;
      vofs[kt] = iu * STR[U] + iv * STR[V];
    }
    const int MT = (mrows + 7) >> 3;
#pragma unroll 1
    for (int ot = 0; ot < OT; ++ot) {
      const int o = ot * 8 + rl;
      const bool ook = o < NO;
      const int oo = ook ? o : 0;
      const int obase = (oo / N[Y]) * STR[X] + (oo % N[Y]) * STR[Y];
      double bf[KT];
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) bf[kt] = (ook && kok[kt]) ? blk[obase + vofs[kt]] : 0.0;
      size_t f1[2], f2[2];
      bool fok[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int oc = ot * 8 + 2 * kl + j;
        fok[j] = oc < NO;
        const int occ = fok[j] ? oc : 0;
        const size_t p = (size_t)(off[X] + occ / N[Y]), q = (size_t)(off[Y] + occ % N[Y]);
        f1[j] = (q * nbf + p) * NM;
        f2[j] = (p * nbf + q) * NM;
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int m = mt * 8 + rl;
        const bool mok = m < mrows;
        double c10 = 0.0, c11 = 0.0, c20 = 0.0, c21 = 0.0;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const bool ld = mok && kok[kt];
          double a1 = ld ? __ldg(P + pa1[kt] + m) : 0.0;
          const double a2 = ld ? __ldg(P + pa2[kt] + m) : 0.0;
          if constexpr (COUL) {
            a1 += a2;
            dmma_m8n8k4(c10, c11, a1, bf[kt]);
          } else {
            dmma_m8n8k4(c10, c11, a1, bf[kt]);
            dmma_m8n8k4(c20, c21, a2, bf[kt]);
          }
        }
        if (mok) {
          if constexpr (COUL) {
            if (fok[0] && c10 != 0.0) { atomicAdd(F + f1[0] + m, scale * c10); atomicAdd(F + f2[0] + m, scale * c10); }
            if (fok[1] && c11 != 0.0) { atomicAdd(F + f1[1] + m, scale * c11); atomicAdd(F + f2[1] + m, scale * c11); }
          } else {
            if (fok[0]) {
              if (c10 != 0.0) atomicAdd(F + f1[0] + m, scale * c10);
              if (c20 != 0.0) atomicAdd(F + f2[0] + m, scale * c20);
            }
            if (fok[1]) {
              if (c11 != 0.0) atomicAdd(F + f1[1] + m, scale * c11);
              if (c21 != 0.0) atomicAdd(F + f2[1] + m, scale * c21);
            }
          }
        }
      }
    }
  } else {
    // lanes over the matrices (m fastest: coalesced P reads and reds), outputs one after the other
#pragma unroll 1
    for (int o = 0; o < NO; ++o) {
      const int ix = o / N[Y], iy = o % N[Y];
      const double* v = blk + ix * STR[X] + iy * STR[Y];
      const size_t p = (size_t)(off[X] + ix), q = (size_t)(off[Y] + iy);
#pragma unroll 1
      for (int m = lane; m < mrows; m += 32) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int iu = 0; iu < N[U]; ++iu)
#pragma unroll
          for (int iv = 0; iv < N[V]; ++iv) {
            const size_t pp = (size_t)(off[U] + iu), qq = (size_t)(off[V] + iv);
            const double x = v[iu * STR[U] + iv * STR[V]];
            s1 = fma(x, __ldg(P + (qq * nbf + pp) * NM + m), s1);
            s2 = fma(x, __ldg(P + (pp * nbf + qq) * NM + m), s2);
          }
        if constexpr (COUL) {
          const double sj = s1 + s2;
          if (sj != 0.0) { atomicAdd(F + (q * nbf + p) * NM + m, scale * sj); atomicAdd(F + (p * nbf + q) * NM + m, scale * sj); }
        } else {
          if (s1 != 0.0) atomicAdd(F + (q * nbf + p) * NM + m, scale * s1);
          if (s2 != 0.0) atomicAdd(F + (p * nbf + q) * NM + m, scale * s2);
        }
      }
    }
  }
}
// all 32 lanes together; blk[a][b][c][d] in shared memory (element cutoff and coincidence factor applied)
template <int N0, int N1, int N2, int N3>
__device__ __forceinline__ void digest_gen_warp(const EriArgs& A, const double* blk, const int (&off)[4], int lane) {
  gen_contract<N0, N1, N2, N3, 0, 1>(A, blk, off, lane);  // Coulomb on (a,b),(b,a)
  gen_contract<N0, N1, N2, N3, 2, 3>(A, blk, off, lane);  // Coulomb on (c,d),(d,c)
  gen_contract<N0, N1, N2, N3, 0, 2>(A, blk, off, lane);  // exchange (a,c),(c,a)
  gen_contract<N0, N1, N2, N3, 0, 3>(A, blk, off, lane);
  gen_contract<N0, N1, N2, N3, 1, 2>(A, blk, off, lane);
  gen_contract<N0, N1, N2, N3, 1, 3>(A, blk, off, lane);
}

// ---------------------------------------------------------------------------------------------------
template <int LA, int LB, int LC, int LD, int PV>
__global__ void __launch_bounds__(ClassCfg<LA, LB, LC, LD>::NT)
eri_kernel(const EriArgs A) {
  constexpr int N0 = Shell<LA, PV>::NOUT, N1 = Shell<LB, PV>::NOUT, N2 = Shell<LC, PV>::NOUT, N3 = Shell<LD, PV>::NOUT;
  constexpr int NTOT = N0 * N1 * N2 * N3;
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  constexpr int R = Cfg::R, TS = Cfg::TS, QPB = Cfg::QPB, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, ND = Cfg::ND;
  constexpr int NMAX = Cfg::NMAX, MMAX = Cfg::MMAX, NKL1 = Cfg::NKL1, KS = Cfg::KS, TKC = Cfg::TKC;
  constexpr int G1 = Cfg::G1, G2 = Cfg::G2, GSTR = Cfg::GSTR, QSM = Cfg::QSM;

  extern __shared__ double smem[];
  QInfo* qinfo = reinterpret_cast<QInfo*>(smem + (size_t)QPB * QSM);
  __shared__ int s_maxk, s_maxl;
  __shared__ unsigned s_base;
  constexpr int LCAP = 64;  // primitive-quartet list window

  const int tid = threadIdx.x;
  const int q = tid / TS;   // local quartet
  const int t = tid % TS;   // team lane
  const bool team_ok = q < QPB;
  double* qs = smem + (size_t)(team_ok ? q : 0) * QSM;
  double* rw = qs + (QSM - 2 * R - 1);   // roots/weights live at the tail of the quartet region
  QInfo& qi = qinfo[team_ok ? q : 0];
  unsigned short* plist = reinterpret_cast<unsigned short*>(qinfo + QPB) + (size_t)(team_ok ? q : 0) * LCAP;

  // thread's bra component
  const int tab = t / KS, slice = t % KS;
  const int ia = tab / NB, ib = tab % NB;
  int ax, ay, az, bx, by, bz;
  cart_xyz_rt(LA, ia, ax, ay, az);
  cart_xyz_rt(LB, ib, bx, by, bz);
  const int obx = (ax * (LB + 1) + bx) * NKL1, oby = (ay * (LB + 1) + by) * NKL1, obz = (az * (LB + 1) + bz) * NKL1;

  const unsigned ntasks = A.task_cap ? min(*A.ntasks, A.task_cap) : *A.ntasks;
  unsigned long long st_prim = 0, st_ints = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_base = atomicAdd(A.counter, (unsigned)QPB);
      s_maxk = 0;
    }
    __syncthreads();
    const unsigned base = s_base;
    if (base >= ntasks) break;

    // ---- setup
    if (team_ok && t == 0) {
      unsigned ti = base + q;
      qi.valid = ti < ntasks;
      qi.nonzero = 0;
      qi.imax = qi.jmax = 0;
      if (qi.valid) {
        int2 tk = A.tasks[ti];
        PairEntry pb = A.bra[tk.x], pk = A.ket[tk.y];
        qi.sa = pb.sa; qi.sb = pb.sb; qi.sc = pk.sa; qi.sd = pk.sb;
        qi.boff = pb.poff; qi.bcnt = pb.pcnt; qi.koff = pk.poff; qi.kcnt = pk.pcnt;
        qi.oa = A.aooff[pb.sa]; qi.ob = A.aooff[pb.sb]; qi.oc = A.aooff[pk.sa]; qi.od = A.aooff[pk.sb];
        qi.bra_id = tk.x; qi.ket_id = tk.y;
        float f = 1.0f;
        if (pb.sa == pb.sb) f *= 0.5f;
        if (pk.sa == pk.sb) f *= 0.5f;
        if ((pb.sa == pk.sa && pb.sb == pk.sb) || (pb.sa == pk.sb && pb.sb == pk.sa)) f *= 0.5f;
        qi.fac = f;
        // primitives are sorted by |K|/zeta: only the leading imax x jmax rectangle can pass the
        // primitive-quartet test (da db)^2 >= cut * (zeta + eta) >= cut * (zmin_bra + zmin_ket)
        int imax = 0, jmax = 0;
        if (pb.pcnt > 0 && pk.pcnt > 0) {
          const double thr = A.prim_cutoff * (1.0 - 1e-9) * (pb.zmin + pk.zmin);
          const double* p0 = A.prim + (size_t)pb.poff * PRIM_STRIDE;
          const double* q0 = A.prim + (size_t)pk.poff * PRIM_STRIDE;
          const double da0 = __ldg(p0 + 4), db0 = __ldg(q0 + 4);
          while (imax < pb.pcnt) {
            const double* pp = p0 + (size_t)imax * PRIM_STRIDE;
            double v = __ldg(pp + 4) * db0;
            if (v * v < thr) break;
            ++imax;
          }
          while (jmax < pk.pcnt) {
            const double* pq = q0 + (size_t)jmax * PRIM_STRIDE;
            double v = __ldg(pq + 4) * da0;
            if (v * v < thr) break;
            ++jmax;
          }
        }
        qi.imax = imax; qi.jmax = jmax;
        atomicMax(&s_maxk, imax * jmax);
      }
    }
    __syncthreads();
    const int maxk = s_maxk;
    const bool valid = team_ok && qi.valid;
    const int imax = valid ? qi.imax : 1, ncand = valid ? qi.imax * qi.jmax : 0;
    double Ax = 0, Ay = 0, Az = 0, Cx = 0, Cy = 0, Cz = 0, ABx = 0, ABy = 0, ABz = 0, CDx = 0, CDy = 0, CDz = 0;
    if (valid) {
      const double* xa = A.xyz + 3 * qi.sa; const double* xb = A.xyz + 3 * qi.sb;
      const double* xc = A.xyz + 3 * qi.sc; const double* xd = A.xyz + 3 * qi.sd;
      Ax = xa[0]; Ay = xa[1]; Az = xa[2]; Cx = xc[0]; Cy = xc[1]; Cz = xc[2];
      ABx = Ax - xb[0]; ABy = Ay - xb[1]; ABz = Az - xb[2];
      CDx = Cx - xd[0]; CDy = Cy - xd[1]; CDz = Cz - xd[2];
    }

    double acc[Cfg::NACC];
#pragma unroll
    for (int k = 0; k < Cfg::NACC; ++k) acc[k] = 0.0;
    bool any = false;

    for (int w0 = 0; w0 < maxk; w0 += LCAP) {
      // ---- window scan: compact the primitive quartets that pass the int_rys.F90:229-232 test
      __syncthreads();  // previous window fully consumed (list, s_maxl)
      if (team_ok && t == 0) qi.lcount = 0;
      if (tid == 0) s_maxl = 0;
      __syncthreads();
      if (valid) {
        const int wend = min(w0 + LCAP, ncand);
        for (int cand = w0 + t; cand < wend; cand += TS) {
          const int i = cand % imax, j = cand / imax;
          const double* pp = A.prim + (size_t)(qi.boff + i) * PRIM_STRIDE;
          const double* pq = A.prim + (size_t)(qi.koff + j) * PRIM_STRIDE;
          const double z = __ldg(pp + 3), e = __ldg(pq + 3);
          const double pf = __ldg(pp + 4) * __ldg(pq + 4);
          if (!(pf * pf < A.prim_cutoff * (z + e + z * e * A.mu2inv))) {
            int pos = atomicAdd(&qi.lcount, 1);
            plist[pos] = (unsigned short)(j * 128 + i);
          }
        }
      }
      __syncthreads();
      if (valid && t == 0 && qi.lcount > 0) atomicMax(&s_maxl, qi.lcount);
      __syncthreads();
      const int maxl = s_maxl;
      const int lcount = valid ? qi.lcount : 0;
    for (int ip = 0; ip < maxl; ++ip) {
      const bool act = ip < lcount;
      double Px = 0, Py = 0, Pz = 0, zeta = 1, Kp = 0, Qx = 0, Qy = 0, Qz = 0, eta = 1, Kq = 0, zinv = 1, einv = 1;
      if (act) {
        const int code = plist[ip];
        const double* pp = A.prim + (size_t)(qi.boff + (code & 127)) * PRIM_STRIDE;
        const double* pq = A.prim + (size_t)(qi.koff + (code >> 7)) * PRIM_STRIDE;
        Px = __ldg(pp); Py = __ldg(pp + 1); Pz = __ldg(pp + 2); zeta = __ldg(pp + 3); Kp = __ldg(pp + 4); zinv = __ldg(pp + 5);
        Qx = __ldg(pq); Qy = __ldg(pq + 1); Qz = __ldg(pq + 2); eta = __ldg(pq + 3); Kq = __ldg(pq + 4); einv = __ldg(pq + 5);
      }
      const double ab = zeta + eta + zeta * eta * A.mu2inv;
      const double pfac = Kp * Kq;
      const bool keep = act;
      const double abinv = 1.0 / ab;
      const double rho = zeta * eta * abinv;
      const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
      const double X = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
      // ---- B1: roots and weights
      if (keep) {
        for (int f = t; f < 2 * R; f += TS) rw[f] = rys_eval<R>(A, X, f);
      }
      __syncthreads();
      // ---- B2: 2-D recurrences, one (root, direction) per thread
      if (keep) {
        const double pref = pfac * sqrt(abinv);
        for (int task = t; task < 3 * R; task += TS) {
          const int r = task / 3, dir = task % 3;
          const double t2 = rw[r];
          const double PAd = dir == 0 ? Px - Ax : (dir == 1 ? Py - Ay : Pz - Az);
          const double QCd = dir == 0 ? Qx - Cx : (dir == 1 ? Qy - Cy : Qz - Cz);
          const double PQd = dir == 0 ? PQx : (dir == 1 ? PQy : PQz);
          const double ABd = dir == 0 ? ABx : (dir == 1 ? ABy : ABz);
          const double CDd = dir == 0 ? CDx : (dir == 1 ? CDy : CDz);
          const double t2r = t2 * rho;
          const double c00 = PAd - t2r * zinv * PQd;
          const double d00 = QCd + t2r * einv * PQd;
          const double b10 = 0.5 * zinv * (1.0 - t2r * zinv);
          const double b01 = 0.5 * einv * (1.0 - t2r * einv);
          const double b00 = 0.5 * t2 * abinv;
          double* S1 = qs + (size_t)task * GSTR;   // [n][m], m fastest
          double* S2 = S1 + G1;                    // [n][c][d]
          double* S3 = S2 + G2;                    // [a][b][c][d]
          // VRR (int_rys.F90:529-617 restated on centres A and C)
          S1[0] = dir == 0 ? rw[R + r] * pref : 1.0;
          if (NMAX > 1) S1[MMAX] = c00 * S1[0];
          for (int n = 1; n < NMAX - 1; ++n) S1[(n + 1) * MMAX] = c00 * S1[n * MMAX] + n * b10 * S1[(n - 1) * MMAX];
          for (int m = 0; m < MMAX - 1; ++m) {
            double v0 = d00 * S1[m];
            if (m > 0) v0 += m * b01 * S1[m - 1];
            S1[m + 1] = v0;
            for (int n = 1; n < NMAX; ++n) {
              double v = d00 * S1[n * MMAX + m] + n * b00 * S1[(n - 1) * MMAX + m];
              if (m > 0) v += m * b01 * S1[n * MMAX + m - 1];
              S1[n * MMAX + m + 1] = v;
            }
          }
          // ket HRR: (c, d+1) = (c+1, d) + (C-D)(c, d)     (int_rys.F90:636-648)
          for (int n = 0; n < NMAX; ++n) {
            double* w = S1 + n * MMAX;
            for (int c = 0; c <= LC; ++c) S2[(n * (LC + 1) + c) * (LD + 1)] = w[c];
            for (int d = 1; d <= LD; ++d) {
              for (int c = 0; c < MMAX - d; ++c) w[c] = w[c + 1] + CDd * w[c];
              for (int c = 0; c <= LC; ++c) S2[(n * (LC + 1) + c) * (LD + 1) + d] = w[c];
            }
          }
          // bra HRR: (a, b+1) = (a+1, b) + (A-B)(a, b)     (int_rys.F90:650-660)
          for (int k = 0; k < NKL1; ++k) {
            for (int a = 0; a <= LA; ++a) S3[(a * (LB + 1)) * NKL1 + k] = S2[a * NKL1 + k];
            for (int b = 1; b <= LB; ++b) {
              for (int n = 0; n < NMAX - b; ++n) S2[n * NKL1 + k] = S2[(n + 1) * NKL1 + k] + ABd * S2[n * NKL1 + k];
              for (int a = 0; a <= LA; ++a) S3[(a * (LB + 1) + b) * NKL1 + k] = S2[a * NKL1 + k];
            }
          }
        }
      }
      __syncthreads();
      // ---- B3: assembly  I(ab|cd) += sum_r gx gy gz      (int_rys.F90:677-713)
      if (keep) {
        any = true;
        if (t == 0) ++st_prim;
        for (int r = 0; r < R; ++r) {
          const double* gx = qs + (size_t)(3 * r + 0) * GSTR + G1 + G2 + obx;
          const double* gy = qs + (size_t)(3 * r + 1) * GSTR + G1 + G2 + oby;
          const double* gz = qs + (size_t)(3 * r + 2) * GSTR + G1 + G2 + obz;
          double X_[NKL1], Y_[NKL1], Z_[NKL1];
#pragma unroll
          for (int k = 0; k < NKL1; ++k) { X_[k] = gx[k]; Y_[k] = gy[k]; Z_[k] = gz[k]; }
          auto body = [&](auto S) {
            constexpr int s0 = decltype(S)::value * TKC;
            static_for<0, Cfg::NACC>([&](auto I) {
              constexpr int k = decltype(I)::value;
              constexpr int ic = s0 + k / ND, id = k % ND;
              if constexpr (ic < NC) {
                constexpr int ix = Cart<LC>::x(ic) * (LD + 1) + Cart<LD>::x(id);
                constexpr int iy = Cart<LC>::y(ic) * (LD + 1) + Cart<LD>::y(id);
                constexpr int iz = Cart<LC>::z(ic) * (LD + 1) + Cart<LD>::z(id);
                acc[k] = fma(X_[ix] * Y_[iy], Z_[iz], acc[k]);
              }
            });
          };
          if constexpr (KS == 1) body(std::integral_constant<int, 0>{});
          else if constexpr (KS == 2) { if (slice == 0) body(std::integral_constant<int, 0>{}); else body(std::integral_constant<int, 1>{}); }
          else { if (slice == 0) body(std::integral_constant<int, 0>{}); else if (slice == 1) body(std::integral_constant<int, 1>{}); else body(std::integral_constant<int, 2>{}); }
        }
      }
      // no barrier here: the next iteration's B1 only writes rw, and B2 (which overwrites the g tables read
      // above) is behind the barrier that follows B1
    }
    }
    __syncthreads();

    // ---- block to shared memory (Cartesian, raw), region 0
    if (valid) {
      if (any && t == 0) qi.nonzero = 1;  // `any` is uniform over the team
      double* blk0 = qs;
#pragma unroll
      for (int k = 0; k < Cfg::NACC; ++k) {
        int ic = slice * TKC + k / ND, id = k % ND;
        if (ic < NC) blk0[((size_t)(ia * NB + ib) * NC + ic) * ND + id] = acc[k];
      }
    }
    __syncthreads();
    if (!(valid && qi.nonzero)) {
      if (valid && A.mode == MODE_SCHWARZ && t == 0) A.qout[qi.bra_id] = 0.0;
      if (valid && A.mode == MODE_BLOCK) {
        for (int e = t; e < NTOT; e += TS) A.blockout[e] = 0.0;
      }
      // all teams still take part in the barriers below
    }
    // ---- normalisation + pure projection, index by index (int2.F90:1187-1207; int_rys.F90:715-785)
    // dims (a,b,c,d) -> transform d, c, b, a ; ping-pong between region 0 and region 1
    double* src = qs;
    double* dst = qs + Cfg::NCART4;
    const bool work = valid && qi.nonzero;
    if (LD >= 2) {
      if (work) proj_smem<LD, Shell<LD, PV>::PURE, NA * NB * NC, 1>(src, dst, t, TS);
      double* tmp = src; src = dst; dst = tmp;
      __syncthreads();
    }
    if (LC >= 2) {
      if (work) proj_smem<LC, Shell<LC, PV>::PURE, NA * NB, N3>(src, dst, t, TS);
      double* tmp = src; src = dst; dst = tmp;
      __syncthreads();
    }
    if (LB >= 2) {
      if (work) proj_smem<LB, Shell<LB, PV>::PURE, NA, N2 * N3>(src, dst, t, TS);
      double* tmp = src; src = dst; dst = tmp;
      __syncthreads();
    }
    if (LA >= 2) {
      if (work) proj_smem<LA, Shell<LA, PV>::PURE, 1, N1 * N2 * N3>(src, dst, t, TS);
      double* tmp = src; src = dst; dst = tmp;
      __syncthreads();
    }
    constexpr int ntot = NTOT;
    if (A.mode == MODE_SCHWARZ) {
      // Q = sqrt(max |(ij|ij)|), int2.F90:1727-1728
      if (work) {
        double mx = 0.0;
        for (int e = t; e < ntot; e += TS) mx = fmax(mx, fabs(src[e]));
        dst[t] = mx;
      }
      __syncthreads();
      if (work && t == 0) {
        double mx = 0.0;
        for (int k = 0; k < TS && k < ntot; ++k) mx = fmax(mx, dst[k]);
        A.qout[qi.bra_id] = sqrt(mx);
      }
      continue;
    }
    if (A.mode == MODE_BLOCK) {
      if (work) for (int e = t; e < ntot; e += TS) A.blockout[e] = src[e];
      continue;
    }
    // ---- element cutoff (int2.F90:1806-1812) and shell-level coincidence factor (int2.F90:1849-1851)
    if (work) {
      const double fac = (double)qi.fac, cut = A.cutoff;
      unsigned nz = 0;
      for (int e = t; e < ntot; e += TS) {
        double v = src[e];
        bool z = fabs(v) < cut;
        nz += !z;
        src[e] = z ? 0.0 : v * fac;
      }
      st_ints += (unsigned long long)nz * (unsigned)(8.0f * qi.fac);
    }
    __syncthreads();
    if (work) {
      if (A.mode == MODE_SYM) digest_sym<N0, N1, N2, N3>(A, src, qi.oa, qi.ob, qi.oc, qi.od, t, TS);
      else digest_gen<N0, N1, N2, N3>(A, src, qi.oa, qi.ob, qi.oc, qi.od, t, TS);
    }
  }
  if (A.stat) {
    if (st_prim) atomicAdd(A.stat, st_prim);
    if (st_ints) atomicAdd(A.stat + 1, st_ints);
  }
}

// ---------------------------------------------------------------------------------------------------
// Register kernel for the small classes (NCART4 <= SMALL_MAX): one thread = one shell quartet, no shared
// memory, no barriers.  Same arithmetic as eri_kernel (2-D VRR on A and C, HRR to B and D, I += gx gy gz);
// the finished block is kept in thread-local memory and digested by the same device functions (t=0, ts=1).
constexpr int SMALL_MAX = 36;
// Medium classes (SMALL_MAX < NCART4 <= MEDIUM_MAX): same one-thread-per-quartet kernel, but the finished
// gx/gy/gz tables of a root live in shared memory (column per thread, [entry][thread]: conflict-free) so the
// registers are left to the NCART4 accumulators.
constexpr int MEDIUM_MAX = OQPB_MEDIUM_MAX;
constexpr int SMALL_NT = 128, MEDIUM_NT = 64;
// register caps requested from ptxas through __launch_bounds__ (min CTAs/SM = 65536 / (threads * cap)); the kernels are
// latency bound at 8 warps/SM, so trading a few spills for occupancy pays for some families (tools/tune_variants.sh)
__host__ __device__ constexpr int min_ctas(int nt, int regcap) { return regcap >= 255 ? 1 : 65536 / (nt * regcap); }
// measured on (H2O)32/cc-pVTZ: these thread-per-quartet classes gain 4-29 % from a 128-register cap (4 CTAs/SM),
// the others lose to the spills
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int small_regcap() {
  constexpr int key = LA * 1000 + LB * 100 + LC * 10 + LD;
#ifndef OQPB_CAP_2000
#define OQPB_CAP_2000 128
#endif
  // ((fs|ps), (fp|ss), (dd|ss) at 168 / 128 instead of 255 registers: +6...+30 %)
  if (key == 2000) return OQPB_CAP_2000;  // (ds|ss): 150-180 registers uncapped (2 CTAs of 128 threads): 75.7 ms; 160: 71.1; 128: 66.6
  // ((pp|ps), (ds|ds) at 168 instead of ~250 registers: +12...+19 %)
  return (key == 2010 || key == 2100 || key == 1100 || key == 1010 || key == 3000) ? 128 : OQPB_SMALL_REGS;
}

// MODE_GEN staging of the thread-per-quartet kernels: blocks of a warp staged at a time (power of two, at most 96 KB / CTA)
__host__ __device__ constexpr int gen_sub(int ntot, int nwarps) {
  int sub = 32;
  while (sub > 4 && (size_t)nwarps * sub * (ntot | 1) * 8 > 96 * 1024) sub /= 2;
  return sub;
}
// software-pipelined primitive loads (24 more registers): measured per class on (H2O)32/cc-pVTZ
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr bool prim_pipe() {
#ifdef OQPB_PRIM_PIPE
  return OQPB_PRIM_PIPE != 0;
#else
  constexpr int key = LA * 1000 + LB * 100 + LC * 10 + LD;
  return key == 0 || key == 2000 || key == 1110 || key == 3010 || key == 3100 || key == 2020 || key == 2200;
#endif
}
// Rys evaluation state at X shared by all roots and weights of a primitive quartet
struct RysX {
  bool asym;
  int iv;
  double rx, rs, t;
};
// 1/sqrt(x) for normal positive x: hardware seed (rsqrt.approx.f64, ~2^-22) + two Newton steps; the library
// rsqrt() with its special-case handling was 16 % of the instructions of the contracted s/p launches
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hx * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}
template <int R>
__device__ __forceinline__ RysX rys_prepare(const EriArgs& a, double X) {
  RysX s;
  s.asym = X >= (double)a.rys_xmax;
  s.iv = 0;
  s.rx = s.rs = s.t = 0.0;
  if (s.asym) {
    s.rs = rsqrt_nr(X);  // half-range Gauss-Hermite asymptote (rys.F90:2711-2713): r = h_r / X, w = h_w / sqrt(X)
    s.rx = s.rs * s.rs;
  } else {
    const double xs = X * RysFmt<R>::DIV;
    s.iv = (int)xs;
    s.t = 2.0 * (xs - (double)s.iv) - 1.0;
  }
  return s;
}
// Chebyshev table of one nroots in shared memory: per interval 2R functions x NC coefficients, padded by
// 2 doubles so that the 16-byte reads of lanes in different intervals fall into different banks
template <int R>
struct RysSmem {
  static constexpr int ROW = 2 * R * RysFmt<R>::NC;
  static constexpr int STRIDE = ROW + 2;
  static constexpr bool USE = R <= 3;
  __host__ __device__ static constexpr int doubles(int xmax) { return xmax * RysFmt<R>::DIV * STRIDE; }
};
// root r (as t^2) and its weight: two interleaved Horner recurrences over 16-byte table loads
template <int R, bool SM>
__device__ __forceinline__ void rys_pair(const EriArgs& a, const double* __restrict__ stab, const RysX& s, int r,
                                         double& t2, double& w) {
  if (s.asym) {
    t2 = a.herm_r[r] * s.rx;
    w = a.herm_w[r] * s.rs;
    return;
  }
  constexpr int NCF = RysFmt<R>::NC, H = NCF / 2;
  double2 p[H], q[H];
  if constexpr (SM) {
    const double2* ct = reinterpret_cast<const double2*>(stab + s.iv * RysSmem<R>::STRIDE + r * NCF);
    const double2* cw = ct + H * R;
#pragma unroll
    for (int k = 0; k < H; ++k) { p[k] = ct[k]; q[k] = cw[k]; }
  } else {
    const double2* __restrict__ ct = reinterpret_cast<const double2*>(a.rys_tab + ((size_t)s.iv * (2 * R) + r) * NCF);
    const double2* __restrict__ cw = ct + H * R;
#pragma unroll
    for (int k = 0; k < H; ++k) { p[k] = __ldg(ct + k); q[k] = __ldg(cw + k); }
  }
  // two interleaved Horner chains on the monomial coefficients (see rys_eval)
  double b = p[H - 1].y, e = q[H - 1].y;
#pragma unroll
  for (int k = NCF - 2; k >= 0; --k) {
    const double ck = (k & 1) ? p[k >> 1].y : p[k >> 1].x;
    const double dk = (k & 1) ? q[k >> 1].y : q[k >> 1].x;
    b = fma(b, s.t, ck);
    e = fma(e, s.t, dk);
  }
  t2 = b;
  w = e;
}

// Primitive-quartet loops of ONE shell quartet in one thread (WPQ: the slice of this lane): roots, 2-D VRR, HRR and the
// assembly I += gx gy gz into acc[NCART4] (Cartesian, raw).  Shared by the task kernel and the run kernel below.
template <int LA, int LB, int LC, int LD, bool GS, bool WPQ, int NTH>
__device__ __forceinline__ void eval_quartet_thread(const EriArgs& A, const PairEntry& pb, const PairEntry& pk, double* gsm, int lane,
                                                    double (&acc)[ClassCfg<LA, LB, LC, LD>::NCART4], bool& any,
                                                    unsigned long long& st_prim) {
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  constexpr int R = Cfg::R, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, ND = Cfg::ND, NCART4 = Cfg::NCART4;
  constexpr int NMAX = Cfg::NMAX, MMAX = Cfg::MMAX, NKL1 = Cfg::NKL1, NIJ1 = Cfg::NIJ1;
  constexpr bool RSM = !GS && RysSmem<R>::USE;
  (void)NA; (void)lane;
  const double Ax = pb.ax, Ay = pb.ay, Az = pb.az, Cx = pk.ax, Cy = pk.ay, Cz = pk.az;
  const double AB[3] = {pb.abx, pb.aby, pb.abz};
  const double CD[3] = {pk.abx, pk.aby, pk.abz};
#pragma unroll
  for (int k = 0; k < NCART4; ++k) acc[k] = 0.0;
  any = false;
  // primitives are sorted by |K|/zeta: prune with (da db)^2 >= cut*(zeta+eta) >= cut*(zmin_bra+zmin_ket)
  const double thr = A.prim_cutoff * (1.0 - 1e-9) * (pb.zmin + pk.zmin);
  double da0 = 0.0;
  if (pb.pcnt > 0) { const double* p0 = A.prim + (size_t)pb.poff * PRIM_STRIDE; da0 = __ldg(p0 + 4); }
  // the primitive records of the NEXT iteration are loaded before the body of the current one (software pipeline:
  // the L1 latency of these small dependent loads was 30 % of the stall samples of the contracted launches)
  const double2* pq0 = reinterpret_cast<const double2*>(A.prim + (size_t)pk.poff * PRIM_STRIDE);
  const double2* pp0 = reinterpret_cast<const double2*>(A.prim + (size_t)pb.poff * PRIM_STRIDE);
  constexpr bool PIPE = prim_pipe<LA, LB, LC, LD>() && !WPQ;
  // WPQ: lane = (lq, lp); ket primitives kq = lq, lq + QS, ...; bra primitives kp = lp, lp + 32/QS, ...
  int kq0 = 0, kqs = 1, kp0 = 0, kps = 1;
  if constexpr (WPQ) {
    int qs = 1;
    while (qs < 32 && qs < pk.pcnt) qs <<= 1;  // warp-uniform: all lanes hold the same quartet
    kq0 = lane % qs; kqs = qs; kp0 = lane / qs; kps = 32 / qs;
  }
  double2 nq01 = make_double2(0, 0), nq23 = nq01, nq45 = nq01;
  if (PIPE && pk.pcnt > 0) { nq01 = __ldg(pq0); nq23 = __ldg(pq0 + 1); nq45 = __ldg(pq0 + 2); }
  for (int kq = kq0; kq < pk.pcnt; kq += kqs) {
    if constexpr (!PIPE) { const double2* pq = pq0 + 3 * kq; nq01 = __ldg(pq); nq23 = __ldg(pq + 1); nq45 = __ldg(pq + 2); }
    const double2 q01 = nq01, q23 = nq23, q45 = nq45;
    if (PIPE && kq + 1 < pk.pcnt) {
      const double2* pq = pq0 + 3 * (kq + 1);
      nq01 = __ldg(pq); nq23 = __ldg(pq + 1); nq45 = __ldg(pq + 2);
    }
    const double Qx = q01.x, Qy = q01.y, Qz = q23.x, eta = q23.y, db = q45.x, einv = q45.y;
    if ((da0 * db) * (da0 * db) < thr) break;
    double2 np01 = make_double2(0, 0), np23 = np01, np45 = np01;
    if (PIPE && pb.pcnt > 0) { np01 = __ldg(pp0); np23 = __ldg(pp0 + 1); np45 = __ldg(pp0 + 2); }
    // (two bra primitives per iteration evaluated branch-free side by side for the R = 1 classes: (ps|ss) 124 -> 157 ms,
    // (ss|ss) 49 -> 56 ms -- the work spent on primitives that fail the tests outweighs the hidden latency)
    constexpr int KP_UNROLL = (R <= OQPB_KP_UNROLL_MAXR && !WPQ) ? OQPB_KP_UNROLL : 1;
#pragma unroll KP_UNROLL
    for (int kp = kp0; kp < pb.pcnt; kp += kps) {
      double2 p01 = np01, p23 = np23, p45 = np45;
      if constexpr (PIPE) {
        if (kp + 1 < pb.pcnt) {
          const double2* pp = pp0 + 3 * (kp + 1);
          np01 = __ldg(pp); np23 = __ldg(pp + 1); np45 = __ldg(pp + 2);
        }
      } else {
        const double2* pp = pp0 + 3 * kp;
        p23 = __ldg(pp + 1); p45 = __ldg(pp + 2);
      }
      const double zeta = p23.y, zinv = p45.y;
      const double pfac = p45.x * db;
      if (pfac * pfac < thr) break;
      const double ab = zeta + eta + zeta * eta * A.mu2inv;  // 2nd term: attenuated integrals, int_rys.F90:225-227
      if (pfac * pfac < A.prim_cutoff * ab) continue;  // int_rys.F90:229-232
      if constexpr (!PIPE) p01 = __ldg(pp0 + 3 * kp);
      const double Px = p01.x, Py = p01.y, Pz = p23.x;
      any = true;
      ++st_prim;
      const double rsab = rsqrt_nr(ab);
      const double abinv = rsab * rsab;
      const double rho = zeta * eta * abinv;
      const double PQ[3] = {Px - Qx, Py - Qy, Pz - Qz};
      const double PA[3] = {Px - Ax, Py - Ay, Pz - Az};
      const double QC[3] = {Qx - Cx, Qy - Cy, Qz - Cz};
      const double X = rho * (PQ[0] * PQ[0] + PQ[1] * PQ[1] + PQ[2] * PQ[2]);
      const double pref = pfac * rsab;
      const double rz = rho * zinv, re = rho * einv, hz = 0.5 * zinv, he = 0.5 * einv;
      const RysX rx = rys_prepare<R>(A, X);
      // OQPB_ROOT_UNROLL > 1: the register classes unroll the root loop so that the next root's interpolation chain overlaps
      // the recurrences of the current one (the thread-per-quartet kernels are latency bound at 8-16 warps per SM)
      // measured on (H2O)32: -3...-8 % on (ps|ps), (ds|ss), (pp|ss), (dp|ss), (fs|ss), (pp|ps); neutral or slower above
      constexpr int ROOT_UNROLL = (GS || R == 1 || (NCART4 > root_unroll_max() && !(LA == 1 && LB == 1 && LC == 1 && LD == 0))) ? 1 : root_unroll();
#pragma unroll ROOT_UNROLL
      for (int r = 0; r < R; ++r) {
        double t2, w;
        rys_pair<R, RSM>(A, gsm, rx, r, t2, w);
        const double b10 = hz * (1.0 - t2 * rz), b01 = he * (1.0 - t2 * re), b00 = 0.5 * t2 * abinv;
        constexpr int G3 = NIJ1 * NKL1;
        double g[GS ? 1 : 3][GS ? 1 : G3];
        double* gcol = gsm + threadIdx.x;
        auto GSET = [&](int dir, int idx, double val) {
          if constexpr (GS) gcol[(dir * G3 + idx) * NTH] = val;
          else g[dir][idx] = val;
        };
#pragma unroll
        for (int dir = 0; dir < 3; ++dir) {
          const double c00 = PA[dir] - t2 * rz * PQ[dir];
          const double d00 = QC[dir] + t2 * re * PQ[dir];
          double v[NMAX][MMAX];
          v[0][0] = dir == 0 ? w * pref : 1.0;
#pragma unroll
          for (int n = 1; n < NMAX; ++n) v[n][0] = c00 * v[n - 1][0] + (n >= 2 ? (n - 1) * b10 * v[n >= 2 ? n - 2 : 0][0] : 0.0);
#pragma unroll
          for (int m = 1; m < MMAX; ++m) {
            v[0][m] = d00 * v[0][m - 1] + (m >= 2 ? (m - 1) * b01 * v[0][m >= 2 ? m - 2 : 0] : 0.0);
#pragma unroll
            for (int n = 1; n < NMAX; ++n)
              v[n][m] = d00 * v[n][m - 1] + n * b00 * v[n - 1][m - 1] + (m >= 2 ? (m - 1) * b01 * v[n][m >= 2 ? m - 2 : 0] : 0.0);
          }
          // ket HRR then bra HRR
          double h[NMAX][NKL1];
#pragma unroll
          for (int n = 0; n < NMAX; ++n) {
#pragma unroll
            for (int c = 0; c <= LC; ++c) h[n][c * (LD + 1)] = v[n][c];
#pragma unroll
            for (int d = 1; d <= LD; ++d) {
#pragma unroll
              for (int c = 0; c < MMAX - d; ++c) v[n][c] = v[n][c + 1] + CD[dir] * v[n][c];
#pragma unroll
              for (int c = 0; c <= LC; ++c) h[n][c * (LD + 1) + d] = v[n][c];
            }
          }
#pragma unroll
          for (int k = 0; k < NKL1; ++k) {
#pragma unroll
            for (int a = 0; a <= LA; ++a) GSET(dir, (a * (LB + 1)) * NKL1 + k, h[a][k]);
#pragma unroll
            for (int b = 1; b <= LB; ++b) {
#pragma unroll
              for (int n = 0; n < NMAX - b; ++n) h[n][k] = h[n + 1][k] + AB[dir] * h[n][k];
#pragma unroll
              for (int a = 0; a <= LA; ++a) GSET(dir, (a * (LB + 1) + b) * NKL1 + k, h[a][k]);
            }
          }
        }
        static_for<0, NCART4>([&](auto I) {
          constexpr int e = decltype(I)::value;
          constexpr int id = e % ND, ic = (e / ND) % NC, ib = (e / (ND * NC)) % NB, ia = e / (ND * NC * NB);
          constexpr int ix = (Cart<LA>::x(ia) * (LB + 1) + Cart<LB>::x(ib)) * NKL1 + Cart<LC>::x(ic) * (LD + 1) + Cart<LD>::x(id);
          constexpr int iy = (Cart<LA>::y(ia) * (LB + 1) + Cart<LB>::y(ib)) * NKL1 + Cart<LC>::y(ic) * (LD + 1) + Cart<LD>::y(id);
          constexpr int iz = (Cart<LA>::z(ia) * (LB + 1) + Cart<LB>::z(ib)) * NKL1 + Cart<LC>::z(ic) * (LD + 1) + Cart<LD>::z(id);
#if OQPB_MED_VOLATILE
          // real shared-memory loads: without `volatile` ptxas forwards the stored table values and keeps all 3*G3 of them
          // in registers next to the NCART4 accumulators (spills above ~100 accumulators)
          if constexpr (GS) { const volatile double* gv = gcol; acc[e] = fma(gv[ix * NTH] * gv[(G3 + iy) * NTH], gv[(2 * G3 + iz) * NTH], acc[e]); }
#else
          if constexpr (GS) acc[e] = fma(gcol[ix * NTH] * gcol[(G3 + iy) * NTH], gcol[(2 * G3 + iz) * NTH], acc[e]);
#endif
          else acc[e] = fma(g[0][ix] * g[1][iy], g[2][iz], acc[e]);
        });
      }
    }
  }
}

// WPQ (warp per quartet): all 32 lanes work on ONE quartet and split its primitive quartets (ket primitives over
// QS lanes x bra primitives over 32/QS lanes), the partial blocks are summed with shuffles and lane 0 carries on
// alone.  Used for launches with few, heavily contracted quartets (small molecules), where one thread per quartet
// would run thousands of primitive quartets serially while the rest of the GPU idles.
template <int LA, int LB, int LC, int LD, int PV, bool GS, bool WPQ = false>
__global__ void __launch_bounds__(GS ? MEDIUM_NT : SMALL_NT, GS ? min_ctas(MEDIUM_NT, OQPB_MED_REGS) : min_ctas(SMALL_NT, small_regcap<LA, LB, LC, LD>()))
eri_small_kernel(const EriArgs A) {
  constexpr int NTH = GS ? MEDIUM_NT : SMALL_NT;
  extern __shared__ double gsm[];
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  constexpr int R = Cfg::R, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, ND = Cfg::ND, NCART4 = Cfg::NCART4;
  constexpr int NMAX = Cfg::NMAX, MMAX = Cfg::MMAX, NKL1 = Cfg::NKL1, NIJ1 = Cfg::NIJ1;
  constexpr int N0 = Shell<LA, PV>::NOUT, N1 = Shell<LB, PV>::NOUT, N2 = Shell<LC, PV>::NOUT, N3 = Shell<LD, PV>::NOUT;
  constexpr int NTOT = N0 * N1 * N2 * N3;
  const unsigned ntasks = A.task_cap ? min(*A.ntasks, A.task_cap) : *A.ntasks;
  unsigned long long st_prim = 0, st_ints = 0;
  constexpr bool RSM = !GS && RysSmem<R>::USE;  // Rys table of this nroots staged in shared memory
  if constexpr (RSM) {
    const int nint = A.rys_xmax * RysFmt<R>::DIV;
    for (int i = threadIdx.x; i < nint * RysSmem<R>::ROW; i += NTH)
      gsm[(i / RysSmem<R>::ROW) * RysSmem<R>::STRIDE + i % RysSmem<R>::ROW] = A.rys_tab[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count: the digestion reduces across the lanes of a warp
  constexpr unsigned TPC = WPQ ? NTH / 32 : NTH;  // tasks per CTA and pass
  for (unsigned tb = blockIdx.x * TPC + (WPQ ? (threadIdx.x >> 5) : (threadIdx.x & ~31u)); tb < ntasks; tb += gridDim.x * TPC) {
    const unsigned ti = WPQ ? tb : tb + lane;
    const bool valid = ti < ntasks;
    const bool validp = valid && (!WPQ || lane == 0);  // the lane that owns the finished block
    const int2 tk = A.tasks[valid ? ti : ntasks - 1];
    PairEntry pb = A.bra[tk.x], pk = A.ket[tk.y];
    if (!valid) pb.pcnt = pk.pcnt = 0;  // no primitive work, zero block
    // task of this lane's NEXT iteration: its pair entries are pulled towards the SM while this quartet is computed
    const unsigned tnext = ti + gridDim.x * blockDim.x;
    int2 tkn = make_int2(-1, -1);
    if (!WPQ && tnext < ntasks) tkn = A.tasks[tnext];
    // density sub-blocks of matrix 0: issued now, consumed by the digestion after the primitive loop
    using Den = DenBlk<N0, N1, N2, N3>;
    constexpr bool DEN_BATCH = Den::SIZE + NTOT <= OQPB_DEN_BATCH_MAX;
    constexpr bool DEN_EARLY = DEN_BATCH && Den::SIZE <= OQPB_DEN_EARLY_MAX;
    Den den0;
    if constexpr (DEN_EARLY) {
      if (A.mode == MODE_SYM) den_load<N0, N1, N2, N3>(A, 0, pb.oa, pb.ob, pk.oa, pk.ob, den0);
    }
    double acc[NCART4];
    bool any = false;
    eval_quartet_thread<LA, LB, LC, LD, GS, WPQ, NTH>(A, pb, pk, gsm, lane, acc, any, st_prim);
    if constexpr (WPQ) {
      any = __any_sync(0xffffffffu, any);
#pragma unroll
      for (int e = 0; e < NCART4; ++e) {
        double v = acc[e];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[e] = lane == 0 ? v : 0.0;
      }
    }
    if (tkn.x >= 0) {
      const char* nb_ = reinterpret_cast<const char*>(A.bra + tkn.x);
      const char* nk_ = reinterpret_cast<const char*>(A.ket + tkn.y);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nb_));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nb_ + 64));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nk_));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(nk_ + 64));
    }
    if (A.mode != MODE_SYM && A.mode != MODE_GEN && !any) {  // (mode is uniform; SYM / GEN keep the warp together)
      if (validp && A.mode == MODE_SCHWARZ) A.qout[tk.x] = 0.0;
      if (validp && A.mode == MODE_BLOCK)
        for (int e = 0; e < NTOT; ++e) A.blockout[e] = 0.0;
      continue;
    }
    // normalisation / pure projection in registers, index by index: d, c, b, a
    double b3[NA * NB * NC * N3], b2[NA * NB * N2 * N3], b1[NA * N1 * N2 * N3], blk[NTOT];
    proj_reg<LD, Shell<LD, PV>::PURE, NA * NB * NC, 1>(acc, b3);
    proj_reg<LC, Shell<LC, PV>::PURE, NA * NB, N3>(b3, b2);
    proj_reg<LB, Shell<LB, PV>::PURE, NA, N2 * N3>(b2, b1);
    proj_reg<LA, Shell<LA, PV>::PURE, 1, N1 * N2 * N3>(b1, blk);
    if (A.mode == MODE_SCHWARZ) {
      double mx = 0.0;
#pragma unroll
      for (int e = 0; e < NTOT; ++e) mx = fmax(mx, fabs(blk[e]));
      if (validp) A.qout[tk.x] = sqrt(mx);
      continue;
    }
    if (A.mode == MODE_BLOCK) {
      if (validp) {
#pragma unroll
        for (int e = 0; e < NTOT; ++e) A.blockout[e] = blk[e];
      }
      continue;
    }
    // element cutoff (int2.F90:1806-1812) and shell-level coincidence factor (int2.F90:1849-1851)
    float facf = 1.0f;
    if (pb.sa == pb.sb) facf *= 0.5f;
    if (pk.sa == pk.sb) facf *= 0.5f;
    if (pb.sa == pk.sa && pb.sb == pk.sb) facf *= 0.5f;
    const double fac = (double)facf, cut = A.cutoff;
    unsigned nz = 0;
#pragma unroll
    for (int e = 0; e < NTOT; ++e) {
      const double v = blk[e];
      const bool z = fabs(v) < cut;
      nz += !z;
      blk[e] = z ? 0.0 : v * fac;
    }
    st_ints += (unsigned long long)nz * (unsigned)(8.0f * facf);
    if (A.mode == MODE_SYM) {
      // runs of equal bra / equal (bra, ket shell c) inside the warp; lanes without a quartet get unique keys
      const long long kbra = validp ? (long long)tk.x : -1 - (long long)lane;
      const SegMask mbra = seg_make(kbra, lane);
      const SegMask mc = seg_make(validp ? ((long long)tk.x << 20) | (long long)pk.sa : kbra, lane);
      digest_sym_reg<N0, N1, N2, N3, !GS, DEN_BATCH>(A, blk, pb.oa, pb.ob, pk.oa, pk.ob, mbra, mc, den0, DEN_EARLY);
    } else {
      // MODE_GEN: the 32 blocks of the warp go to shared memory and are digested one after the other by all lanes
      // (GEN_SUB blocks per warp at a time, so that the staging area stays small and the occupancy is kept)
      constexpr int GBS = NTOT | 1, SUB = gen_sub(NTOT, NTH / 32);
      double* wb = gsm + (GS ? 3 * NIJ1 * NKL1 * NTH : (RSM ? RysSmem<R>::doubles(A.rys_xmax) : 0)) +
                   (size_t)(threadIdx.x >> 5) * (SUB * GBS);
      const unsigned live_all = __ballot_sync(0xffffffffu, validp && any);
#pragma unroll 1
      for (int h = 0; h < 32 / SUB; ++h) {
        unsigned live = live_all & (unsigned)(((1ull << SUB) - 1ull) << (h * SUB));
        if (live == 0) continue;
        __syncwarp();
        if (lane / SUB == h) {
#pragma unroll
          for (int e = 0; e < NTOT; ++e) wb[(lane % SUB) * GBS + e] = blk[e];
        }
        __syncwarp();
        while (live) {
          const int q = __ffs(live) - 1;
          live &= live - 1;
          const int off[4] = {__shfl_sync(0xffffffffu, pb.oa, q), __shfl_sync(0xffffffffu, pb.ob, q),
                              __shfl_sync(0xffffffffu, pk.oa, q), __shfl_sync(0xffffffffu, pk.ob, q)};
          digest_gen_warp<N0, N1, N2, N3>(A, wb + (q % SUB) * GBS, off, lane);
        }
      }
    }
  }
  if (A.stat) {
    if (st_prim) atomicAdd(A.stat, st_prim);
    if (st_ints) atomicAdd(A.stat + 1, st_ints);
  }
}

// ---------------------------------------------------------------------------------------------------
// Run kernel (MODE_SYM, one Fock matrix): the thread-per-quartet evaluation above, but a WARP walks a run of up to
// RUN_LEN consecutive surviving kets of ONE bra (k_enum writes a bra's survivors contiguously and cuts them into warp
// items); lane l takes the kets l, l + 32, ... of the run, so the lanes of one step still hold neighbouring kets
// (coalesced pair entries / primitive records, shared density rows).  The bra entry and D_ab are loaded once per run,
// J_ab is accumulated in registers over the whole run (one butterfly + one red per element and run instead of a
// segmented shuffle reduction per step); K_ac / K_bc keep the per-step segmented reduction over the lanes that share the
// ket shell c.  A lane-private run (lane l = kets l*M .. l*M+M-1, K_ac / K_bc in registers too) removed 28-42 % of the
// instructions of the uncontracted launches but was 15 % SLOWER overall: lanes 16 kets apart share no cache lines.
constexpr int RUN_LEN = 32 * OQPB_RUN_M;  // <= 255 (the length shares a word with the bra index): OQPB_RUN_M <= 7
static_assert(RUN_LEN <= 255, "OQPB_RUN_M");
template <int N0, int N1, int N2, int N3, bool SEGC>
__device__ __forceinline__ void digest_run(const EriArgs& A, const double (&v)[N0 * N1 * N2 * N3], const double (&dab)[N0 * N1],
                                           int o0, int o1, int o2, int o3, double (&jab)[N0 * N1], const SegMask& mc) {
  const unsigned nbf = (unsigned)A.nbf;
  const double c4 = 4.0 * A.cj, c1 = A.ck;
  const double* __restrict__ DJ = A.DJ[0];
  const double* __restrict__ DK = A.DK[0];
  double* __restrict__ F = A.F[0];
#define VV(a, b, c, d) v[(((a)*N1 + (b)) * N2 + (c)) * N3 + (d)]
  {  // J_ab += sum_cd v D_cd   (4 cj at the flush)
    double dd[N2 * N3];
#pragma unroll
    for (int c = 0; c < N2; ++c)
#pragma unroll
      for (int d = 0; d < N3; ++d) dd[c * N3 + d] = __ldg(DJ + ((unsigned)(o2 + c) * nbf + (unsigned)(o3 + d)));
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        double sum = jab[a * N1 + b];
#pragma unroll
        for (int c = 0; c < N2; ++c)
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), dd[c * N3 + d], sum);
        jab[a * N1 + b] = sum;
      }
  }
  {  // J_cd += 4 cj sum_ab v D_ab
#pragma unroll
    for (int c = 0; c < N2; ++c)
#pragma unroll
      for (int d = 0; d < N3; ++d) {
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int b = 0; b < N1; ++b) sum = fma(VV(a, b, c, d), dab[a * N1 + b], sum);
        if (sum != 0.0) atomicAdd(F + tri_u(o2 + c, o3 + d), c4 * sum);
      }
  }
  {  // K_ac -= ck sum_bd v D_bd
    double dd[N1 * N3];
#pragma unroll
    for (int b = 0; b < N1; ++b)
#pragma unroll
      for (int d = 0; d < N3; ++d) dd[b * N3 + d] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o3 + d)));
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int c = 0; c < N2; ++c) {
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), dd[b * N3 + d], sum);
        if constexpr (SEGC) sum = seg_sum(sum, mc);
        if ((!SEGC || mc.head) && sum != 0.0) atomicAdd(F + tri_u(o0 + a, o2 + c), -c1 * sum);
      }
  }
  {  // K_ad -= ck sum_bc v D_bc
    double dd[N1 * N2];
#pragma unroll
    for (int b = 0; b < N1; ++b)
#pragma unroll
      for (int c = 0; c < N2; ++c) dd[b * N2 + c] = __ldg(DK + ((unsigned)(o1 + b) * nbf + (unsigned)(o2 + c)));
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int d = 0; d < N3; ++d) {
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < N1; ++b)
#pragma unroll
          for (int c = 0; c < N2; ++c) sum = fma(VV(a, b, c, d), dd[b * N2 + c], sum);
        if (sum != 0.0) atomicAdd(F + tri_u(o0 + a, o3 + d), -c1 * sum);
      }
  }
  {  // K_bc -= ck sum_ad v D_ad
    double dd[N0 * N3];
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int d = 0; d < N3; ++d) dd[a * N3 + d] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o3 + d)));
#pragma unroll
    for (int b = 0; b < N1; ++b)
#pragma unroll
      for (int c = 0; c < N2; ++c) {
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int d = 0; d < N3; ++d) sum = fma(VV(a, b, c, d), dd[a * N3 + d], sum);
        if constexpr (SEGC) sum = seg_sum(sum, mc);
        if ((!SEGC || mc.head) && sum != 0.0) atomicAdd(F + tri_u(o1 + b, o2 + c), -c1 * sum);
      }
  }
  {  // K_bd -= ck sum_ac v D_ac
    double dd[N0 * N2];
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int c = 0; c < N2; ++c) dd[a * N2 + c] = __ldg(DK + ((unsigned)(o0 + a) * nbf + (unsigned)(o2 + c)));
#pragma unroll
    for (int b = 0; b < N1; ++b)
#pragma unroll
      for (int d = 0; d < N3; ++d) {
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < N0; ++a)
#pragma unroll
          for (int c = 0; c < N2; ++c) sum = fma(VV(a, b, c, d), dd[a * N2 + c], sum);
        if (sum != 0.0) atomicAdd(F + tri_u(o1 + b, o3 + d), -c1 * sum);
      }
  }
#undef VV
}

// the run kernel carries 2 N0 N1 more doubles than the task kernel: keep the classes that sat just under 128 registers there
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int run_regcap() {
  constexpr int key = LA * 1000 + LB * 100 + LC * 10 + LD;
  return (key == 1000 || key == 0) ? 128 : small_regcap<LA, LB, LC, LD>();
}
template <int LA, int LB, int LC, int LD, int PV, bool GS>
__global__ void __launch_bounds__(GS ? MEDIUM_NT : SMALL_NT, GS ? min_ctas(MEDIUM_NT, OQPB_MED_REGS) : min_ctas(SMALL_NT, run_regcap<LA, LB, LC, LD>()))
eri_run_kernel(const EriArgs A) {
  constexpr int NTH = GS ? MEDIUM_NT : SMALL_NT;
  constexpr int WPC = NTH / 32;
  extern __shared__ double gsm[];
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  constexpr int R = Cfg::R, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, NCART4 = Cfg::NCART4;
  constexpr int N0 = Shell<LA, PV>::NOUT, N1 = Shell<LB, PV>::NOUT, N2 = Shell<LC, PV>::NOUT, N3 = Shell<LD, PV>::NOUT;
  constexpr int NTOT = N0 * N1 * N2 * N3;
  constexpr bool RSM = !GS && RysSmem<R>::USE;
  constexpr unsigned FULL = 0xffffffffu;
  unsigned long long st_prim = 0, st_ints = 0;
  if constexpr (RSM) {
    const int nint = A.rys_xmax * RysFmt<R>::DIV;
    for (int i = threadIdx.x; i < nint * RysSmem<R>::ROW; i += NTH)
      gsm[(i / RysSmem<R>::ROW) * RysSmem<R>::STRIDE + i % RysSmem<R>::ROW] = A.rys_tab[i];
    __syncthreads();
  }
  const unsigned nitems = min(*A.nitems, A.item_cap);
  const int lane = threadIdx.x & 31;
  double* __restrict__ F = A.F[0];
  const double c4 = 4.0 * A.cj, cut = A.cutoff;
  const unsigned nbf = (unsigned)A.nbf;
  for (unsigned it = blockIdx.x * WPC + (threadIdx.x >> 5); it < nitems; it += gridDim.x * WPC) {
    const int2 item = A.items[it];  // warp-uniform
    const int len = (int)((unsigned)item.x >> 24);
    if (len == 0) continue;
    const PairEntry pb = A.bra[item.x & 0xffffff];
    const unsigned start = (unsigned)item.y;
    double jab[N0 * N1], dab[N0 * N1];
#pragma unroll
    for (int a = 0; a < N0; ++a)
#pragma unroll
      for (int b = 0; b < N1; ++b) {
        jab[a * N1 + b] = 0.0;
        dab[a * N1 + b] = __ldg(A.DJ[0] + ((unsigned)(pb.oa + a) * nbf + (unsigned)(pb.ob + b)));
      }
#pragma unroll 1
    for (int i0 = 0; i0 < len; i0 += 32) {
      const int idx = i0 + lane;
      const bool act = idx < len;
      PairEntry pk = A.ket[A.tasks[start + (act ? idx : 0)].y];
      if (!act) pk.pcnt = 0;  // no primitive work, zero block
      double acc[NCART4];
      bool any = false;
      eval_quartet_thread<LA, LB, LC, LD, GS, false, NTH>(A, pb, pk, gsm, lane, acc, any, st_prim);
      if (!GS && !__any_sync(FULL, any)) continue;  // (the K reductions of the register classes are warp collectives)
      if (GS && !any) continue;
      // normalisation / pure projection in registers, index by index: d, c, b, a
      double b3[NA * NB * NC * N3], b2[NA * NB * N2 * N3], b1[NA * N1 * N2 * N3], blk[NTOT];
      proj_reg<LD, Shell<LD, PV>::PURE, NA * NB * NC, 1>(acc, b3);
      proj_reg<LC, Shell<LC, PV>::PURE, NA * NB, N3>(b3, b2);
      proj_reg<LB, Shell<LB, PV>::PURE, NA, N2 * N3>(b2, b1);
      proj_reg<LA, Shell<LA, PV>::PURE, 1, N1 * N2 * N3>(b1, blk);
      // element cutoff (int2.F90:1806-1812) and shell-level coincidence factor (int2.F90:1849-1851)
      float facf = 1.0f;
      if (pb.sa == pb.sb) facf *= 0.5f;
      if (pk.sa == pk.sb) facf *= 0.5f;
      if (pb.sa == pk.sa && pb.sb == pk.sb) facf *= 0.5f;
      const double fac = (double)facf;
      unsigned nz = 0;
#pragma unroll
      for (int e = 0; e < NTOT; ++e) {
        const double v = blk[e];
        const bool z = fabs(v) < cut;
        nz += !z;
        blk[e] = z ? 0.0 : v * fac;
      }
      st_ints += (unsigned long long)nz * (unsigned)(8.0f * facf);
      SegMask mc;
      if constexpr (!GS && OQPB_RUN_SEGC) mc = seg_make(act ? (long long)pk.sa : -1 - (long long)lane, lane);  // runs of equal ket shell c
      digest_run<N0, N1, N2, N3, !GS && OQPB_RUN_SEGC>(A, blk, dab, pb.oa, pb.ob, pk.oa, pk.ob, jab, mc);
    }
    // J_ab of the whole run: butterfly over the lanes, lane (e mod 32) issues the red of element e
#pragma unroll
    for (int e = 0; e < N0 * N1; ++e) {
      double v = jab[e];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == (e & 31) && v != 0.0) atomicAdd(F + tri_u(pb.oa + e / N1, pb.ob + e % N1), c4 * v);
    }
  }
  if (A.stat) {
    if (st_prim) atomicAdd(A.stat, st_prim);
    if (st_ints) atomicAdd(A.stat + 1, st_ints);
  }
}

// ---------------------------------------------------------------------------------------------------
// Group kernel for the large classes: a quartet is owned by an aligned group of G lanes of ONE warp
// (G = 4, 8, 16 or 32), a warp works on 32/G quartets at once, and every phase boundary is a __syncwarp():
// no CTA barrier anywhere, warps fetch their own work.  Each lane owns NVL = ceil(NA*NB/G) bra Cartesian
// component pairs ("virtual lanes") with all NC*ND ket components in registers.  Phases per primitive quartet
// as in eri_kernel: B1 roots (2R tasks over the G lanes), B2 2-D recurrences (3R tasks), B3 assembly.
template <int LA, int LB, int LC, int LD>
struct GroupCfg {
  using C = ClassCfg<LA, LB, LC, LD>;
  static constexpr int VL = C::NA * C::NB;
  static constexpr int NKET = C::NKET;
  static constexpr int acc_for(int g) { return ((VL + g - 1) / g) * NKET; }
  // accumulators per lane for G < 32; (fs|dd) measured 27 % faster with 8 lanes x 72 accumulators than 16 x 36
  static constexpr int LIMIT = (LA == 3 && LB == 0 && LC == 2 && LD == 2) ? 72 : OQPB_GRP_LIMIT;
  static constexpr int LIMIT32 = 72;    // ... and for full-warp groups
  static constexpr bool OK = (C::KS == 1) && acc_for(32) <= LIMIT32;
  static constexpr int G = acc_for(4) <= LIMIT ? 4 : (acc_for(8) <= LIMIT ? 8 : (acc_for(16) <= LIMIT ? 16 : 32));
  static constexpr int NVL = (VL + G - 1) / G;
  static constexpr int QPW = 32 / G;
  // B2 in registers: the 2-D VRR and both HRR transfers of a (root, direction) task run fully unrolled in registers
  // and only the final [a][b][c][d] table goes to shared memory (the shared-memory pipe is the group kernel's limiter)
  static constexpr bool REGVRR = C::NMAX * C::MMAX + C::NMAX * C::NKL1 <= OQPB_REGVRR_MAX;
  static constexpr int GSTR = REGVRR ? (C::G3 | 1) : C::GSTR;   // doubles per (root, direction) table
  static constexpr int GOFF = REGVRR ? 0 : C::G1 + C::G2;       // offset of the [a][b][c][d] table inside it
  static constexpr int GREG = 3 * C::R * GSTR;
  static constexpr int QSM0 = ((GREG > C::BLK ? GREG : C::BLK) + 2 * C::R + 2) | 1;
  // per-quartet stride in doubles: == G (mod 16) for G < 16, so that the 16 lanes of a half-warp (16/G quartets x G
  // lanes, lane stride odd) fall into 16 different 8-byte banks
  static constexpr int QSMG = G >= 16 ? QSM0 : ((QSM0 + 15 - G) / 16) * 16 + G;
  static_assert(QSMG >= QSM0, "QSMG");
  static constexpr int QBYTES = QSMG * 8 + 96 + 128;  // block/g region + QInfo + primitive list
  static constexpr int WPC = (2 * QPW * QBYTES <= 64 * 1024) ? 2 : 1;
  static constexpr int NT = 32 * WPC;
  static constexpr size_t SMEM = (size_t)WPC * QPW * QBYTES;
};

template <int LA, int LB, int LC, int LD, int PV>
__global__ void __launch_bounds__(GroupCfg<LA, LB, LC, LD>::NT, min_ctas(GroupCfg<LA, LB, LC, LD>::NT, OQPB_GRP_REGS))
eri_group_kernel(const EriArgs A) {
  constexpr int N0 = Shell<LA, PV>::NOUT, N1 = Shell<LB, PV>::NOUT, N2 = Shell<LC, PV>::NOUT, N3 = Shell<LD, PV>::NOUT;
  constexpr int NTOT = N0 * N1 * N2 * N3;
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  using GC = GroupCfg<LA, LB, LC, LD>;
  constexpr int R = Cfg::R, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, ND = Cfg::ND, NKET = Cfg::NKET;
  constexpr int NMAX = Cfg::NMAX, MMAX = Cfg::MMAX, NKL1 = Cfg::NKL1, NIJ1 = Cfg::NIJ1;
  constexpr int G1 = Cfg::G1, G2 = Cfg::G2, GSTR = GC::GSTR, GOFF = GC::GOFF, QSM = GC::QSMG;
  constexpr int G = GC::G, NVL = GC::NVL, QPW = GC::QPW, VL = GC::VL;
  constexpr int LCAP = 64;
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane / G, t = lane % G;
  const int qslot = w * QPW + g;
  double* qs = smem + (size_t)qslot * QSM;
  double* rw = qs + (QSM - 2 * R - 2);  // 2R roots/weights + abinv + prefactor
  char* tail = reinterpret_cast<char*>(smem + (size_t)GC::WPC * QPW * QSM);
  QInfo& qi = *reinterpret_cast<QInfo*>(tail + (size_t)qslot * 96);
  unsigned short* plist = reinterpret_cast<unsigned short*>(tail + (size_t)GC::WPC * QPW * 96) + (size_t)qslot * LCAP;

  // virtual lanes of this thread
  int obx[NVL], oby[NVL], obz[NVL];
#pragma unroll
  for (int j = 0; j < NVL; ++j) {
    int vt = t + G * j;
    int ia = (vt < VL ? vt : 0) / NB, ib = (vt < VL ? vt : 0) % NB;
    int ax, ay, az, bx, by, bz;
    cart_xyz_rt(LA, ia, ax, ay, az);
    cart_xyz_rt(LB, ib, bx, by, bz);
    obx[j] = ax * (LB + 1) + bx; oby[j] = ay * (LB + 1) + by; obz[j] = az * (LB + 1) + bz;  // table layout [c][d][a][b]
  }
  const unsigned ntasks = A.task_cap ? min(*A.ntasks, A.task_cap) : *A.ntasks;
  unsigned long long st_prim = 0, st_ints = 0;

#if OQPB_GRP_STATIC_WALK
  // static warp-strided walk over the task list (helped the ket-owner kernel; here measured 1.5-2.5 % SLOWER than the dynamic
  // fetch: the contracted quartets of this kernel's classes vary more in cost)
  for (unsigned base = (blockIdx.x * GC::WPC + w) * QPW; base < ntasks; base += gridDim.x * GC::WPC * QPW) {
#else
  for (;;) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(A.counter, (unsigned)QPW);
    base = __shfl_sync(FULL, base, 0);
    if (base >= ntasks) break;
#endif
    __syncwarp();
    if (t == 0) {
      unsigned ti = base + g;
      qi.valid = ti < ntasks;
      qi.nonzero = 0;
      qi.imax = qi.jmax = 0;
      qi.lcount = 0;
      if (qi.valid) {
        int2 tk = A.tasks[ti];
        PairEntry pb = A.bra[tk.x], pk = A.ket[tk.y];
        qi.sa = pb.sa; qi.sb = pb.sb; qi.sc = pk.sa; qi.sd = pk.sb;
        qi.boff = pb.poff; qi.bcnt = pb.pcnt; qi.koff = pk.poff; qi.kcnt = pk.pcnt;
        qi.oa = A.aooff[pb.sa]; qi.ob = A.aooff[pb.sb]; qi.oc = A.aooff[pk.sa]; qi.od = A.aooff[pk.sb];
        qi.bra_id = tk.x; qi.ket_id = tk.y;
        float f = 1.0f;
        if (pb.sa == pb.sb) f *= 0.5f;
        if (pk.sa == pk.sb) f *= 0.5f;
        if (pb.sa == pk.sa && pb.sb == pk.sb) f *= 0.5f;
        qi.fac = f;
        int imax = 0, jmax = 0;
        if (pb.pcnt > 0 && pk.pcnt > 0) {
          const double thr = A.prim_cutoff * (1.0 - 1e-9) * (pb.zmin + pk.zmin);
          const double* p0 = A.prim + (size_t)pb.poff * PRIM_STRIDE;
          const double* q0 = A.prim + (size_t)pk.poff * PRIM_STRIDE;
          const double da0 = __ldg(p0 + 4), db0 = __ldg(q0 + 4);
          while (imax < pb.pcnt) { double v = __ldg(p0 + (size_t)imax * PRIM_STRIDE + 4) * db0; if (v * v < thr) break; ++imax; }
          while (jmax < pk.pcnt) { double v = __ldg(q0 + (size_t)jmax * PRIM_STRIDE + 4) * da0; if (v * v < thr) break; ++jmax; }
        }
        qi.imax = imax; qi.jmax = jmax;
      }
    }
    __syncwarp();
    const bool valid = qi.valid;
    const int imax = valid ? max(qi.imax, 1) : 1, ncand = valid ? qi.imax * qi.jmax : 0;
    const int maxk = __reduce_max_sync(FULL, ncand);
    double Ax = 0, Ay = 0, Az = 0, Cx = 0, Cy = 0, Cz = 0, ABx = 0, ABy = 0, ABz = 0, CDx = 0, CDy = 0, CDz = 0;
    if (valid) {
      const double* xa = A.xyz + 3 * qi.sa; const double* xb = A.xyz + 3 * qi.sb;
      const double* xc = A.xyz + 3 * qi.sc; const double* xd = A.xyz + 3 * qi.sd;
      Ax = xa[0]; Ay = xa[1]; Az = xa[2]; Cx = xc[0]; Cy = xc[1]; Cz = xc[2];
      ABx = Ax - xb[0]; ABy = Ay - xb[1]; ABz = Az - xb[2];
      CDx = Cx - xd[0]; CDy = Cy - xd[1]; CDz = Cz - xd[2];
    }
    double acc[NVL][NKET];
#pragma unroll
    for (int j = 0; j < NVL; ++j)
#pragma unroll
      for (int k = 0; k < NKET; ++k) acc[j][k] = 0.0;
    bool any = false;

    for (int w0 = 0; w0 < maxk; w0 += LCAP) {
      __syncwarp();
      if (t == 0) qi.lcount = 0;
      __syncwarp();
      if (valid) {
        const int wend = min(w0 + LCAP, ncand);
        for (int cand = w0 + t; cand < wend; cand += G) {
          const int i = cand % imax, j = cand / imax;
          const double* pp = A.prim + (size_t)(qi.boff + i) * PRIM_STRIDE;
          const double* pq = A.prim + (size_t)(qi.koff + j) * PRIM_STRIDE;
          const double pf = __ldg(pp + 4) * __ldg(pq + 4);
          const double zz = __ldg(pp + 3), ee = __ldg(pq + 3);
          if (!(pf * pf < A.prim_cutoff * (zz + ee + zz * ee * A.mu2inv))) {
            int pos = atomicAdd(&qi.lcount, 1);
            plist[pos] = (unsigned short)(j * 128 + i);
          }
        }
      }
      __syncwarp();
      const int lcount = valid ? qi.lcount : 0;
      const int maxl = __reduce_max_sync(FULL, lcount);
      for (int ip = 0; ip < maxl; ++ip) {
        const bool keep = ip < lcount;
        double Px = 0, Py = 0, Pz = 0, zeta = 1, Kp = 0, Qx = 0, Qy = 0, Qz = 0, eta = 1, Kq = 0, zinv = 1, einv = 1;
        if (keep) {
          const int code = plist[ip];
          const double* pp = A.prim + (size_t)(qi.boff + (code & 127)) * PRIM_STRIDE;
          const double* pq = A.prim + (size_t)(qi.koff + (code >> 7)) * PRIM_STRIDE;
          Px = __ldg(pp); Py = __ldg(pp + 1); Pz = __ldg(pp + 2); zeta = __ldg(pp + 3); Kp = __ldg(pp + 4); zinv = __ldg(pp + 5);
          Qx = __ldg(pq); Qy = __ldg(pq + 1); Qz = __ldg(pq + 2); eta = __ldg(pq + 3); Kq = __ldg(pq + 4); einv = __ldg(pq + 5);
        }
        const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
        // ---- B1: lane 0 of the group prepares the primitive scalars, 2R lanes evaluate roots and weights
        if (keep) {
          const double abinv = 1.0 / (zeta + eta + zeta * eta * A.mu2inv);
          const double rho = zeta * eta * abinv;
          const double X = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
          if (t == 0) { rw[2 * R] = abinv; rw[2 * R + 1] = Kp * Kq * sqrt(abinv); }
          for (int f = t; f < 2 * R; f += G) rw[f] = rys_eval<R>(A, X, f);
        }
        __syncwarp();
        // ---- B2: 2-D recurrences, (root, direction) tasks over the G lanes
        if (keep) {
          const double abinv = rw[2 * R], pref = rw[2 * R + 1];
          const double rho = zeta * eta * abinv;
          for (int task = t; task < 3 * R; task += G) {
            const int r = task / 3, dir = task % 3;
            const double t2 = rw[r];
            const double PAd = dir == 0 ? Px - Ax : (dir == 1 ? Py - Ay : Pz - Az);
            const double QCd = dir == 0 ? Qx - Cx : (dir == 1 ? Qy - Cy : Qz - Cz);
            const double PQd = dir == 0 ? PQx : (dir == 1 ? PQy : PQz);
            const double ABd = dir == 0 ? ABx : (dir == 1 ? ABy : ABz);
            const double CDd = dir == 0 ? CDx : (dir == 1 ? CDy : CDz);
            const double t2r = t2 * rho;
            const double c00 = PAd - t2r * zinv * PQd;
            const double d00 = QCd + t2r * einv * PQd;
            const double b10 = 0.5 * zinv * (1.0 - t2r * zinv);
            const double b01 = 0.5 * einv * (1.0 - t2r * einv);
            const double b00 = 0.5 * t2 * abinv;
            if constexpr (GC::REGVRR) {
              double* S3 = qs + (size_t)task * GSTR;
              double v[NMAX][MMAX];
              v[0][0] = dir == 0 ? rw[R + r] * pref : 1.0;
#pragma unroll
              for (int n = 1; n < NMAX; ++n) v[n][0] = c00 * v[n - 1][0] + (n >= 2 ? (n - 1) * b10 * v[n >= 2 ? n - 2 : 0][0] : 0.0);
#pragma unroll
              for (int m = 1; m < MMAX; ++m) {
                v[0][m] = d00 * v[0][m - 1] + (m >= 2 ? (m - 1) * b01 * v[0][m >= 2 ? m - 2 : 0] : 0.0);
#pragma unroll
                for (int n = 1; n < NMAX; ++n)
                  v[n][m] = d00 * v[n][m - 1] + n * b00 * v[n - 1][m - 1] + (m >= 2 ? (m - 1) * b01 * v[n][m >= 2 ? m - 2 : 0] : 0.0);
              }
              double h[NMAX][NKL1];
#pragma unroll
              for (int n = 0; n < NMAX; ++n) {
#pragma unroll
                for (int c = 0; c <= LC; ++c) h[n][c * (LD + 1)] = v[n][c];
#pragma unroll
                for (int d = 1; d <= LD; ++d) {
#pragma unroll
                  for (int c = 0; c < MMAX - d; ++c) v[n][c] = v[n][c + 1] + CDd * v[n][c];
#pragma unroll
                  for (int c = 0; c <= LC; ++c) h[n][c * (LD + 1) + d] = v[n][c];
                }
              }
#pragma unroll
              for (int k = 0; k < NKL1; ++k) {
#pragma unroll
                for (int a = 0; a <= LA; ++a) S3[k * NIJ1 + a * (LB + 1)] = h[a][k];
#pragma unroll
                for (int b = 1; b <= LB; ++b) {
#pragma unroll
                  for (int n = 0; n < NMAX - b; ++n) h[n][k] = h[n + 1][k] + ABd * h[n][k];
#pragma unroll
                  for (int a = 0; a <= LA; ++a) S3[k * NIJ1 + a * (LB + 1) + b] = h[a][k];
                }
              }
            } else {
            double* S1 = qs + (size_t)task * GSTR;
            double* S2 = S1 + G1;
            double* S3 = S2 + G2;
            S1[0] = dir == 0 ? rw[R + r] * pref : 1.0;
            if (NMAX > 1) S1[MMAX] = c00 * S1[0];
            for (int n = 1; n < NMAX - 1; ++n) S1[(n + 1) * MMAX] = c00 * S1[n * MMAX] + n * b10 * S1[(n - 1) * MMAX];
            for (int m = 0; m < MMAX - 1; ++m) {
              double v0 = d00 * S1[m];
              if (m > 0) v0 += m * b01 * S1[m - 1];
              S1[m + 1] = v0;
              for (int n = 1; n < NMAX; ++n) {
                double v = d00 * S1[n * MMAX + m] + n * b00 * S1[(n - 1) * MMAX + m];
                if (m > 0) v += m * b01 * S1[n * MMAX + m - 1];
                S1[n * MMAX + m + 1] = v;
              }
            }
            for (int n = 0; n < NMAX; ++n) {
              double* wv = S1 + n * MMAX;
              for (int c = 0; c <= LC; ++c) S2[(n * (LC + 1) + c) * (LD + 1)] = wv[c];
              for (int d = 1; d <= LD; ++d) {
                for (int c = 0; c < MMAX - d; ++c) wv[c] = wv[c + 1] + CDd * wv[c];
                for (int c = 0; c <= LC; ++c) S2[(n * (LC + 1) + c) * (LD + 1) + d] = wv[c];
              }
            }
            for (int k = 0; k < NKL1; ++k) {
              for (int a = 0; a <= LA; ++a) S3[k * NIJ1 + a * (LB + 1)] = S2[a * NKL1 + k];
              for (int b = 1; b <= LB; ++b) {
                for (int n = 0; n < NMAX - b; ++n) S2[n * NKL1 + k] = S2[(n + 1) * NKL1 + k] + ABd * S2[n * NKL1 + k];
                for (int a = 0; a <= LA; ++a) S3[k * NIJ1 + a * (LB + 1) + b] = S2[a * NKL1 + k];
              }
            }
            }
          }
        }
        __syncwarp();
        // ---- B3: assembly
        if (keep) {
          any = true;
          if (t == 0) ++st_prim;
          for (int r = 0; r < R; ++r) {
            const double* gbase = qs + (size_t)(3 * r) * GSTR + GOFF;
#pragma unroll
            for (int j = 0; j < NVL; ++j) {
              if (t + G * j < VL) {
                const double* gx = gbase + obx[j];
                const double* gy = gbase + GSTR + oby[j];
                const double* gz = gbase + 2 * GSTR + obz[j];
                double X_[NKL1], Y_[NKL1], Z_[NKL1];
#pragma unroll
                // [c][d][a][b] layout: the lanes of a group (different a,b) read neighbouring words, no bank conflicts
                for (int k = 0; k < NKL1; ++k) { X_[k] = gx[k * NIJ1]; Y_[k] = gy[k * NIJ1]; Z_[k] = gz[k * NIJ1]; }
                static_for<0, NKET>([&](auto I) {
                  constexpr int k = decltype(I)::value;
                  constexpr int ic = k / ND, id = k % ND;
                  constexpr int ix = Cart<LC>::x(ic) * (LD + 1) + Cart<LD>::x(id);
                  constexpr int iy = Cart<LC>::y(ic) * (LD + 1) + Cart<LD>::y(id);
                  constexpr int iz = Cart<LC>::z(ic) * (LD + 1) + Cart<LD>::z(id);
                  acc[j][k] = fma(X_[ix] * Y_[iy], Z_[iz], acc[j][k]);
                });
              }
            }
          }
        }
        __syncwarp();  // g tables are overwritten by the next primitive's B2; rw by its B1
      }
    }
    __syncwarp();
    // ---- block to shared memory (aliases the g tables)
    if (valid) {
      if (any && t == 0) qi.nonzero = 1;
#pragma unroll
      for (int j = 0; j < NVL; ++j) {
        int vt = t + G * j;
        if (vt < VL) {
#pragma unroll
          for (int k = 0; k < NKET; ++k) qs[(size_t)vt * NKET + k] = acc[j][k];
        }
      }
    }
    __syncwarp();
    const bool work = valid && qi.nonzero;
    if (valid && !qi.nonzero) {
      if (A.mode == MODE_SCHWARZ && t == 0) A.qout[qi.bra_id] = 0.0;
      if (A.mode == MODE_BLOCK)
        for (int e = t; e < NTOT; e += G) A.blockout[e] = 0.0;
    }
    double* src = qs;
    double* dst = qs + Cfg::NCART4;
    if (LD >= 2) {
      if (work) proj_smem<LD, Shell<LD, PV>::PURE, NA * NB * NC, 1>(src, dst, t, G);
      double* tmp = src; src = dst; dst = tmp;
      __syncwarp();
    }
    if (LC >= 2) {
      if (work) proj_smem<LC, Shell<LC, PV>::PURE, NA * NB, N3>(src, dst, t, G);
      double* tmp = src; src = dst; dst = tmp;
      __syncwarp();
    }
    if (LB >= 2) {
      if (work) proj_smem<LB, Shell<LB, PV>::PURE, NA, N2 * N3>(src, dst, t, G);
      double* tmp = src; src = dst; dst = tmp;
      __syncwarp();
    }
    if (LA >= 2) {
      if (work) proj_smem<LA, Shell<LA, PV>::PURE, 1, N1 * N2 * N3>(src, dst, t, G);
      double* tmp = src; src = dst; dst = tmp;
      __syncwarp();
    }
    constexpr int ntot = NTOT;
    if (A.mode == MODE_SCHWARZ) {
      double mx = 0.0;
      if (work) for (int e = t; e < ntot; e += G) mx = fmax(mx, fabs(src[e]));
#pragma unroll
      for (int off = G / 2; off >= 1; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off));
      if (work && t == 0) A.qout[qi.bra_id] = sqrt(mx);
      continue;
    }
    if (A.mode == MODE_BLOCK) {
      if (work) for (int e = t; e < ntot; e += G) A.blockout[e] = src[e];
      continue;
    }
    if (work) {
      const double fac = (double)qi.fac, cut = A.cutoff;
      unsigned nz = 0;
      for (int e = t; e < ntot; e += G) {
        double v = src[e];
        bool z = fabs(v) < cut;
        nz += !z;
        src[e] = z ? 0.0 : v * fac;
      }
      st_ints += (unsigned long long)nz * (unsigned)(8.0f * qi.fac);
    }
    __syncwarp();
    if (A.mode == MODE_SYM) {
      // quartets of the warp with the same bra / the same (bra, ket shell c) are summed before the red
      const long long kbra = valid ? (long long)qi.bra_id : -1 - (long long)g;
      const GroupSeg<G> mbra = gseg_make<G>(kbra, lane);
      const GroupSeg<G> mc = gseg_make<G>(valid ? ((long long)qi.bra_id << 20) | (long long)qi.sc : kbra, lane);
      digest_sym_group<N0, N1, N2, N3, G>(A, src, qi.oa, qi.ob, qi.oc, qi.od, t, work, mbra, mc);
    } else {
      // MODE_GEN: the quartets of the warp one after the other, each digested by all 32 lanes (DMMA)
      const int srcoff = (int)(src - qs);
#pragma unroll 1
      for (int gq = 0; gq < QPW; ++gq) {
        if (!__shfl_sync(FULL, work ? 1 : 0, gq * G)) continue;
        const QInfo& qq = *reinterpret_cast<const QInfo*>(tail + (size_t)(w * QPW + gq) * 96);
        const int off[4] = {qq.oa, qq.ob, qq.oc, qq.od};
        digest_gen_warp<N0, N1, N2, N3>(A, smem + (size_t)(w * QPW + gq) * QSM + srcoff, off, lane);
      }
    }
  }
  if (A.stat) {
    if (st_prim) atomicAdd(A.stat, st_prim);
    if (st_ints) atomicAdd(A.stat + 1, st_ints);
  }
}

#include "eri_kown.cuh"

// cudaFuncSetAttribute is per device: remember per (kernel instantiation, device) whether the opt-in is done
struct DevFlags {
  bool done[64] = {};
  bool& cur() {
    int dev = 0;
    cudaGetDevice(&dev);
    return done[dev & 63];
  }
};

template <int LA, int LB, int LC, int LD, int PV>
cudaError_t launch_eri(const EriArgs& args, int nblocks, cudaStream_t st) {
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  if constexpr (Cfg::NCART4 <= SMALL_MAX) {
    constexpr int R = Cfg::R;
    constexpr int NTOT = Shell<LA, PV>::NOUT * Shell<LB, PV>::NOUT * Shell<LC, PV>::NOUT * Shell<LD, PV>::NOUT;
    constexpr size_t gen = (size_t)(SMALL_NT / 32) * gen_sub(NTOT, SMALL_NT / 32) * (NTOT | 1) * sizeof(double);  // MODE_GEN block staging
    const size_t smem = (RysSmem<R>::USE ? (size_t)RysSmem<R>::doubles(args.rys_xmax) * sizeof(double) : 0) +
                        (args.mode == MODE_GEN ? gen : 0);
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_small_kernel<LA, LB, LC, LD, PV, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_small_kernel<LA, LB, LC, LD, PV, false><<<nblocks, SMALL_NT, smem, st>>>(args);
    return cudaGetLastError();
  } else if constexpr (Cfg::NCART4 <= MEDIUM_MAX) {
    constexpr int NTOT = Shell<LA, PV>::NOUT * Shell<LB, PV>::NOUT * Shell<LC, PV>::NOUT * Shell<LD, PV>::NOUT;
    constexpr size_t gen = (size_t)(MEDIUM_NT / 32) * gen_sub(NTOT, MEDIUM_NT / 32) * (NTOT | 1) * sizeof(double);  // MODE_GEN block staging
    constexpr size_t smem0 = (size_t)3 * Cfg::NIJ1 * Cfg::NKL1 * MEDIUM_NT * sizeof(double);
    const size_t smem = smem0 + (args.mode == MODE_GEN ? gen : 0);
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && smem0 + gen > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_small_kernel<LA, LB, LC, LD, PV, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem0 + gen));
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_small_kernel<LA, LB, LC, LD, PV, true><<<nblocks, MEDIUM_NT, smem, st>>>(args);
    return cudaGetLastError();
  } else if constexpr (GroupCfg<LA, LB, LC, LD>::OK) {
    using GC = GroupCfg<LA, LB, LC, LD>;
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && GC::SMEM > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_group_kernel<LA, LB, LC, LD, PV>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GC::SMEM);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_group_kernel<LA, LB, LC, LD, PV><<<nblocks, GC::NT, GC::SMEM, st>>>(args);
    return cudaGetLastError();
  } else {
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(eri_kernel<LA, LB, LC, LD, PV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg::SMEM);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_kernel<LA, LB, LC, LD, PV><<<nblocks, Cfg::NT, Cfg::SMEM, st>>>(args);
    return cudaGetLastError();
  }
}

// warp-per-quartet launch of the register kernel (classes with <= SMALL_MAX Cartesian integrals); nullptr otherwise
template <int LA, int LB, int LC, int LD, int PV>
cudaError_t launch_eri_wpq(const EriArgs& args, int nblocks, cudaStream_t st) {
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  if constexpr (Cfg::NCART4 <= SMALL_MAX) {
    constexpr int R = Cfg::R;
    constexpr int NTOT = Shell<LA, PV>::NOUT * Shell<LB, PV>::NOUT * Shell<LC, PV>::NOUT * Shell<LD, PV>::NOUT;
    constexpr size_t gen = (size_t)(SMALL_NT / 32) * (gen_sub(NTOT, SMALL_NT / 32) * (NTOT | 1) + 2 * gen_sub(NTOT, SMALL_NT / 32)) * sizeof(double);
    const size_t smem = (RysSmem<R>::USE ? (size_t)RysSmem<R>::doubles(args.rys_xmax) * sizeof(double) : 0) +
                        (args.mode == MODE_GEN ? gen : 0);
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_small_kernel<LA, LB, LC, LD, PV, false, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_small_kernel<LA, LB, LC, LD, PV, false, true><<<nblocks, SMALL_NT, smem, st>>>(args);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}
// run-kernel launch (thread-per-quartet classes, MODE_SYM with one Fock matrix)
template <int LA, int LB, int LC, int LD, int PV>
cudaError_t launch_eri_run(const EriArgs& args, int nblocks, cudaStream_t st) {
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  if constexpr (Cfg::NCART4 <= SMALL_MAX) {
    constexpr int R = Cfg::R;
    const size_t smem = RysSmem<R>::USE ? (size_t)RysSmem<R>::doubles(args.rys_xmax) * sizeof(double) : 0;
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_run_kernel<LA, LB, LC, LD, PV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_run_kernel<LA, LB, LC, LD, PV, false><<<nblocks, SMALL_NT, smem, st>>>(args);
    return cudaGetLastError();
  } else if constexpr (Cfg::NCART4 <= MEDIUM_MAX) {
    constexpr size_t smem0 = (size_t)3 * Cfg::NIJ1 * Cfg::NKL1 * MEDIUM_NT * sizeof(double);
    // OQPB_MED_CTAS=n: pad the dynamic shared memory so that at most n CTAs share an SM (fewer resident threads -> the
    // register spills of the >= 90-Cartesian classes stay in L1; experiment knob)
    static const int want = getenv("OQPB_MED_CTAS") ? atoi(getenv("OQPB_MED_CTAS")) : 0;
    const size_t smem = want > 0 ? std::max(smem0, (size_t)(227 * 1024 / (want + 1) + 1024)) : smem0;
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_run_kernel<LA, LB, LC, LD, PV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_run_kernel<LA, LB, LC, LD, PV, true><<<nblocks, MEDIUM_NT, smem, st>>>(args);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}
// ket-owner group kernel launch; cudaErrorNotSupported for classes it does not cover
template <int LA, int LB, int LC, int LD, int PV>
cudaError_t launch_eri_kown(const EriArgs& args, int nblocks, cudaStream_t st) {
  using KC = KownCfg<LA, LB, LC, LD, PV>;
  if constexpr (KC::OK) {
    static DevFlags flags;
    bool& attr_set = flags.cur();
    if (!attr_set && KC::SMEM > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(eri_kown_kernel<LA, LB, LC, LD, PV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KC::SMEM);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    eri_kown_kernel<LA, LB, LC, LD, PV><<<nblocks, KC::NT, KC::SMEM, st>>>(args);
    return cudaGetLastError();
  } else {
    return cudaErrorNotSupported;
  }
}
template <int LA, int LB, int LC, int LD, int PV>
constexpr int class_kown_qpb() {
  using KC = KownCfg<LA, LB, LC, LD, PV>;
  return KC::OK ? KC::WPC * KC::QPW : 0;
}
// classes that use the ket-owner kernel by default (OQPB_KOWN=2: every class it covers, 0: none); measured on (H2O)32/cc-pVTZ
template <int LA, int LB, int LC, int LD>
constexpr bool class_kown_default() {
  constexpr int key = LA * 1000 + LB * 100 + LC * 10 + LD;
  // per-class A/B against the bra-owner group kernel / the thread-per-quartet kernels (gpurun_out/s8_class_w32_*.txt):
  // the classes with >= 160 Cartesian integrals and a ket of p or d shells win 12-60 %
  return key == 2111 || key == 2220 || key == 2221 || key == 3121 || key == 3131 || key == 3221 || key == 3220 || key == 3211 ||
         key == 3231 || key == 3210 || key == 3111 || key == 3120 || key == 3230 || key == 3130 || key == 2211 || key == 2121;
}

template <int LA, int LB, int LC, int LD>
constexpr bool class_has_run() {
#ifdef OQPB_RUN_SMALL_ONLY
  return ClassCfg<LA, LB, LC, LD>::NCART4 <= SMALL_MAX;
#else
  return ClassCfg<LA, LB, LC, LD>::NCART4 <= MEDIUM_MAX;
#endif
}
template <int LA, int LB, int LC, int LD>
constexpr bool class_has_wpq() { return ClassCfg<LA, LB, LC, LD>::NCART4 <= SMALL_MAX; }

template <int LA, int LB, int LC, int LD>
constexpr int class_tasks_per_cta() {
  using C = ClassCfg<LA, LB, LC, LD>;
  using GC = GroupCfg<LA, LB, LC, LD>;
  return C::NCART4 <= SMALL_MAX ? SMALL_NT : (C::NCART4 <= MEDIUM_MAX ? MEDIUM_NT : (GC::OK ? GC::WPC * GC::QPW : C::QPB));
}

// grid cap: the thread-per-quartet kernels with a Rys table prologue run about two waves of resident CTAs
template <int LA, int LB, int LC, int LD>
constexpr int class_max_ctas() {
  using C = ClassCfg<LA, LB, LC, LD>;
  return (C::NCART4 <= SMALL_MAX && RysSmem<C::R>::USE) ? 148 * OQPB_SMALL_GRID : 148 * 32;
}

using LaunchFn = cudaError_t (*)(const EriArgs&, int, cudaStream_t);
struct ClassEntry { LaunchFn launch; int qpb; int nt; size_t smem; int maxcta; LaunchFn launch_wpq; LaunchFn launch_run; LaunchFn launch_kown; int kown_qpb; bool kown_default; };
// class table: [pure variant PV = (d pure) | (f pure) << 1][quartet class]
const ClassEntry* class_table(int pv);

}  // namespace oqpb
