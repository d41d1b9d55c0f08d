// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
// (openqp_b200/*); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use it.
//
// CPU restatement (C++17 + OpenMP) of the reference's direct-SCF two-electron path
// (/root/reference/source/integrals/int2.F90 and the Rys engine behind it).  The reference itself
// cannot be compiled here (no Fortran compiler, network-only externals), so this is a
// line-by-line restatement of the *algorithm*, pinned by the reference's pure-HF golden energies
// (tests/test_oracle_golden.py: examples/HF/H2O_RHF-HF_ENERGY.json etc.).
//
// Engine note: the reference dispatches s/p/d quartets to its rotated-axis code and f quartets to
// libint2 (v2.7.1.1-am4, un-vendored); both are replaced here by the reference's own in-tree Rys
// engine (int_rys.F90), which is valid for every L and is the officially supported `rys_only` mode
// (int2.F90:154-157).  Rys roots/weights use the reference's general algorithm (discretised
// Stieltjes + implicit-QL Golub-Welsch, rys.F90:2697-2881) for EVERY nroots; the reference switches
// to polynomial fits for nroots<=5 (rys.F90:45-2695), agreement ~1e-13 relative.
// Parity status: SCF goldens pinned (s,p,d Cartesian); f-shell and spherical d parity vs the real
// binary "unpinned" (no pure-HF golden with those exists in the reference).
//
// Each function cites the reference file:line it follows.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pure_tables.inc"

namespace {

constexpr int MAXL = 4;
constexpr int MXCART = 15;
constexpr int MAXCONTR = 120;  // int_rys.F90:7

// Cartesian component exponents, constants.F90:33-60
const int CX[5][15] = {{0}, {1, 0, 0}, {2, 0, 0, 1, 1, 0}, {3, 0, 0, 2, 2, 1, 0, 1, 0, 1},
                       {4, 0, 0, 3, 3, 1, 0, 1, 0, 2, 2, 0, 2, 1, 1}};
const int CY[5][15] = {{0}, {0, 1, 0}, {0, 2, 0, 1, 0, 1}, {0, 3, 0, 1, 0, 2, 2, 0, 1, 1},
                       {0, 4, 0, 1, 0, 3, 3, 0, 1, 2, 0, 2, 1, 2, 1}};
const int CZ[5][15] = {{0}, {0, 0, 1}, {0, 0, 2, 0, 1, 1}, {0, 0, 3, 0, 1, 0, 1, 2, 2, 1},
                       {0, 0, 4, 0, 1, 0, 1, 3, 3, 0, 2, 2, 1, 1, 2}};

inline int ncart(int l) { return (l + 1) * (l + 2) / 2; }

// shells_pnrm2, constants.F90:121-164: sqrt((2l-1)!! / ((2lx-1)!!(2ly-1)!!(2lz-1)!!))
double pnrm2(int l, int c) {
  auto df = [](int n) { double r = 1; for (int k = n; k > 1; k -= 2) r *= k; return r; };
  return std::sqrt(df(2 * l - 1) / (df(2 * CX[l][c] - 1) * df(2 * CY[l][c] - 1) * df(2 * CZ[l][c] - 1)));
}

// ------------------------------------------------------------------------------------------------
// Rys roots and weights: rys.F90:2697-2881 (general path) with tables regenerated numerically
// (rys_lut.F90 holds Gauss-Legendre half-range nodes^2 / weights and Gauss-Hermite half-range
//  nodes^2 / weights; both are recomputed here by Newton iteration instead of being copied).
constexpr int MXRYS = 13;
const int NAUXS[MXRYS] = {20, 25, 30, 30, 35, 40, 40, 40, 45, 50, 50, 55, 55};  // rys_lut.F90:8-9
// rys_lut.F90:12-17 switch points + 10.  At the reference's own switch points the Hermite asymptote is
// only ~1.5e-12 accurate; the reference's nroots<=5 fits (not restated) are ~1e-14 accurate there, so the
// oracle delays the switch to stay at fit-level accuracy for every nroots.
const double XASYMP[MXRYS] = {39, 47, 53, 59, 65, 70, 75, 81, 86, 91, 96, 101, 106};

struct RysTables {
  std::vector<double> raux[MXRYS], waux[MXRYS];  // per nroots
  double rherm[MXRYS][MXRYS], wherm[MXRYS][MXRYS];
  RysTables() {
    for (int n = 1; n <= MXRYS; n++) {
      int na = NAUXS[n - 1], N = 2 * na;
      raux[n - 1].resize(na);
      waux[n - 1].resize(na);
      // positive Gauss-Legendre nodes of order N (Newton, long double)
      for (int i = 0; i < na; i++) {
        long double x = cosl(M_PIl * (i + 0.75L) / (N + 0.5L)), pp = 0;
        for (int it = 0; it < 100; it++) {
          long double p0 = 1, p1 = x;
          for (int k = 2; k <= N; k++) { long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
          pp = N * (x * p1 - p0) / (x * x - 1);
          long double dx = p1 / pp;
          x -= dx;
          if (fabsl(dx) < 1e-19L) break;
        }
        long double p0 = 1, p1 = x;
        for (int k = 2; k <= N; k++) { long double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
        pp = N * (x * p1 - p0) / (x * x - 1);
        raux[n - 1][i] = (double)(x * x);
        waux[n - 1][i] = (double)(2 / ((1 - x * x) * pp * pp));
      }
      // positive Gauss-Hermite nodes of order 2n: Golub-Welsch in long double via Newton on H_N
      int NH = 2 * n;
      std::vector<long double> xs;
      // Newton from a scan of sign changes of the orthonormal Hermite polynomial
      auto hval = [&](long double x, long double &dp) {
        long double p0 = 0.7511255444649425L, p1 = sqrtl(2.0L) * x * p0;  // pi^-1/4
        if (NH == 0) { dp = 0; return p0; }
        for (int k = 2; k <= NH; k++) {
          long double p2 = x * sqrtl(2.0L / k) * p1 - sqrtl((long double)(k - 1) / k) * p0;
          p0 = p1; p1 = p2;
        }
        dp = sqrtl(2.0L * NH) * p0;
        return p1;
      };
      long double xmax = sqrtl(2.0L * NH + 1) + 1, step = 1e-3L, dp;
      long double prev = hval(0, dp);
      for (long double x = step; x < xmax; x += step) {
        long double cur = hval(x, dp);
        if ((prev < 0) != (cur < 0)) {
          long double r = x - step / 2;
          for (int it = 0; it < 100; it++) { long double v = hval(r, dp); long double d = v / dp; r -= d; if (fabsl(d) < 1e-19L) break; }
          xs.push_back(r);
        }
        prev = cur;
      }
      for (int i = 0; i < n; i++) {
        long double v = hval(xs[i], dp);
        (void)v;
        rherm[n - 1][i] = (double)(xs[i] * xs[i]);
        wherm[n - 1][i] = (double)(2.0L / (dp * dp));
      }
    }
  }
};
const RysTables &rys_tables() { static RysTables t; return t; }

// rys.F90:2729-2789
void discretized_stieltjes(int n, int naux, const double *r, const double *w, double *alpha, double *beta,
                           double *p_old, double *p) {
  double pp_old = 0, pp = 0;
  for (int i = 0; i < naux; i++) pp_old += w[i];
  for (int i = 0; i < naux; i++) pp += w[i] * r[i];
  alpha[0] = pp / pp_old;
  beta[0] = pp_old;
  if (n == 1) return;
  for (int i = 0; i < naux; i++) { p_old[i] = 0; p[i] = 1; }
  for (int k = 0; k < n - 1; k++) {
    pp = 0;
    double xpp = 0;
    for (int i = 0; i < naux; i++) {
      double tmp = p[i];
      p[i] = (r[i] - alpha[k]) * p[i] - beta[k] * p_old[i];
      p_old[i] = tmp;
      pp += w[i] * p[i] * p[i];
      xpp += r[i] * w[i] * p[i] * p[i];
    }
    alpha[k + 1] = xpp / pp;
    beta[k + 1] = pp / pp_old;
    pp_old = pp;
  }
}

// rys.F90:2791-2881 (alpha 1-based -> a[0..n-1]; beta(0:n) -> b[0..n])
void golub_welsch(int n, double *a, double *b, double eps, double *wt) {
  if (n == 1) { wt[0] = b[0]; return; }
  double mu0 = b[0];
  for (int i = 1; i <= n - 1; i++) b[i] = std::sqrt(b[i]);
  b[n] = 0;
  wt[0] = 1;
  for (int i = 1; i < n; i++) wt[i] = 0;
  auto A = [&](int i) -> double & { return a[i - 1]; };
  auto W = [&](int i) -> double & { return wt[i - 1]; };
  for (int l = 1; l <= n; l++) {
    int j = 0;
    for (;;) {
      int m;
      for (m = l; m <= n - 1; m++)
        if (std::fabs(b[m]) <= eps * (std::fabs(A(m)) + std::fabs(A(m + 1)))) break;
      if (m == l) break;
      if (j == 30) return;
      j++;
      double g = (A(l + 1) - A(l)) / (2.0 * b[l]);
      double r = std::sqrt(g * g + 1.0);
      g = A(m) - A(l) + b[l] / (g + std::copysign(r, g));
      double s = 1, c = 1, p = 0;
      for (int i = m - 1; i >= l; i--) {
        double f = s * b[i], bb = c * b[i];
        if (std::fabs(f) < std::fabs(g)) {
          s = f / g; r = std::sqrt(s * s + 1.0); b[i + 1] = g * r; c = 1.0 / r; s = s * c;
        } else {
          c = g / f; r = std::sqrt(c * c + 1.0); b[i + 1] = f * r; s = 1.0 / r; c = c * s;
        }
        g = A(i + 1) - p;
        r = (A(i) - g) * s + 2.0 * c * bb;
        p = s * r;
        A(i + 1) = g + p;
        g = c * r - bb;
        f = W(i + 1);
        W(i + 1) = s * W(i) + c * f;
        W(i) = c * W(i) - s * f;
      }
      A(l) = A(l) - p;
      b[l] = g;
      b[m] = 0;
    }
  }
  for (int i = 0; i < n; i++) wt[i] = mu0 * wt[i] * wt[i];
}

// ---- optional table-driven roots, for the TIMED CPU baseline only (bench.py --impl reference / cpu_baseline) ----------
// The reference evaluates nroots <= 5 from polynomial fits (rys.F90:45-2695), two orders of magnitude cheaper per X than
// the general Stieltjes path restated above; those fit tables are not restated here.  With orc_set_fast_rys(1) the roots
// and weights of nroots <= 7 are interpolated instead from the Chebyshev tables of tools/gen_rys_tables.py (100-digit
// mpmath moments; unit X intervals, 12 terms, half-range Hermite asymptote beyond the table), so that the CPU arm of the
// benchmark is not handicapped by the root finder.  The parity tests never enable it (the checker stays independent of
// anything the CUDA path uses); tests/test_oracle_golden.py checks the two paths against each other.
#include "../openqp_b200/csrc/rys_tables.inc"
static int g_fast_rys = 0;
static inline void rys_fast(double x, int nroots, double *u, double *w) {
  const int R = nroots;
  if (x >= (double)RYS_XMAX_H[R - 1]) {
    const double xi = 1.0 / x, rs = std::sqrt(xi);
    for (int i = 0; i < R; i++) {
      const double r = RYS_HERM_R_H[R - 1][i] * xi;
      u[i] = r / (1.0 - r);
      w[i] = RYS_HERM_W_H[R - 1][i] * rs;
    }
    return;
  }
  const int iv = (int)x;
  const double t = 2.0 * (x - (double)iv) - 1.0, t2 = 2.0 * t;
  const double *c = RYS_TAB_H + RYS_OFF_H[R - 1] + (size_t)iv * (2 * R) * RYS_NCOEF;
  for (int f = 0; f < 2 * R; f++, c += RYS_NCOEF) {  // c0 + sum_k c_k T_k(t), Clenshaw
    double b1 = 0.0, b2 = 0.0;
    for (int k = RYS_NCOEF - 1; k >= 1; k--) { const double b0 = t2 * b1 - b2 + c[k]; b2 = b1; b1 = b0; }
    const double v = t * b1 - b2 + c[0];
    if (f < R) u[f] = v / (1.0 - v); else w[f - R] = v;
  }
}

// rys.F90:2697-2727; returns u = r/(1-r) and weights
void rys_general(double x, int nroots, double *u, double *w) {
  if (g_fast_rys && nroots <= RYS_MAXR) { rys_fast(x, nroots, u, w); return; }
  const RysTables &T = rys_tables();
  double r[MXRYS];
  if (x >= XASYMP[nroots - 1]) {
    for (int i = 0; i < nroots; i++) {
      r[i] = T.rherm[nroots - 1][i] / x;
      w[i] = T.wherm[nroots - 1][i] / std::sqrt(x);
    }
  } else {
    int naux = NAUXS[nroots - 1];
    double rg[55], wg[55], s1[55], s2[55], beta[MXRYS + 1];
    for (int i = 0; i < naux; i++) {
      rg[i] = T.raux[nroots - 1][i];
      wg[i] = T.waux[nroots - 1][i] * std::exp(-x * rg[i]);
    }
    discretized_stieltjes(nroots, naux, rg, wg, r, beta, s1, s2);
    golub_welsch(nroots, r, beta, 1.0e-14, w);
  }
  for (int i = 0; i < nroots; i++) u[i] = r[i] / (1 - r[i]);
}

// ------------------------------------------------------------------------------------------------
struct Cutoffs {  // int2_pairs.F90:27-36, 285-294
  double integral, pair, quartet, exponent, pair2, quartet2;
  void set(double ci, double cp, double cq, double ce) {
    integral = ci; pair = cp; pair2 = cp * cp; quartet = cq; quartet2 = cq * cq; exponent = ce;
  }
};

struct Pairs {  // int2_pairs.F90:14-25
  std::vector<double> aa, ab, g, ginv, k, p, pa, pb;
  std::vector<int> cnt, off;
};

struct Basis {
  int nshell = 0, nprim = 0, nbf = 0, harmonic_active = 0;
  std::vector<int> am, ncontr, goff, aooff, naos, harm;
  std::vector<double> ex, cc, cen;
};

inline int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // 0-based lower-triangular, i>=j

// int2_pairs.F90:179-266 (fill pass; the count pass :74-176 uses the log form of the same test and
// is folded into the fill: a pair entry exists iff the fill predicate keeps it)
void build_pairs(const Basis &b, const Cutoffs &c, Pairs &pp) {
  int n2 = b.nshell * (b.nshell + 1) / 2;
  pp.cnt.assign(n2, 0);
  pp.off.assign(n2, 0);
  pp.aa.clear(); pp.ab.clear(); pp.g.clear(); pp.ginv.clear(); pp.k.clear(); pp.p.clear(); pp.pa.clear(); pp.pb.clear();
  const double sqrtpito52 = std::sqrt(2.0) * std::pow(4.0 * std::atan(1.0), 1.25);
  for (int i = 0; i < b.nshell; i++)
    for (int j = 0; j <= i; j++) {
      int sha = i, shb = j;
      if (b.am[i] > b.am[j]) { sha = j; shb = i; }  // lower AM first, :202-212
      const double *A = &b.cen[3 * sha], *B = &b.cen[3 * shb];
      double ab2 = 0;
      for (int x = 0; x < 3; x++) ab2 += (A[x] - B[x]) * (A[x] - B[x]);
      int id = tri(i, j);
      pp.off[id] = (int)pp.g.size();
      int cntp = 0;
      for (int p1 = 0; p1 < b.ncontr[sha]; p1++)
        for (int p2 = 0; p2 < b.ncontr[shb]; p2++) {
          double a1 = b.ex[b.goff[sha] + p1], a2 = b.ex[b.goff[shb] + p2];
          double gam = a1 + a2;
          double e12 = a1 * a2 * ab2;
          if (e12 > gam * c.exponent) continue;
          double gi = 1 / gam;
          e12 = e12 * gi;
          double k1 = b.cc[b.goff[sha] + p1] * b.cc[b.goff[shb] + p2] * std::exp(-e12);
          if (std::fabs(k1) < c.quartet) continue;
          pp.aa.push_back(a1); pp.ab.push_back(a2); pp.g.push_back(gam); pp.ginv.push_back(gi);
          for (int x = 0; x < 3; x++) {
            double P = (a1 * A[x] + a2 * B[x]) * gi;
            pp.p.push_back(P);
            pp.pa.push_back(P - A[x]);
            pp.pb.push_back(P - B[x]);
          }
          pp.k.push_back(sqrtpito52 * k1);
          cntp++;
        }
      pp.cnt[id] = cntp;
    }
}

// ------------------------------------------------------------------------------------------------
// Rys ERI engine: int_rys.F90
struct Proj { int ncart, nout, nterm[MXCART], out[MXCART][4]; double coef[MXCART][4]; };
void init_proj(int l, int pure, Proj &p) {  // int2_pure_generated.F90:22-53
  p.ncart = ncart(l);
  if (pure && l >= 2) {
    p.nout = 2 * l + 1;
    for (int c = 0; c < p.ncart; c++) {
      p.nterm[c] = PURE_NTERM_H[l - 2][c];
      for (int t = 0; t < 4; t++) { p.out[c][t] = PURE_OUT_H[l - 2][c][t]; p.coef[c][t] = PURE_COEF_H[l - 2][c][t]; }
    }
  } else {
    p.nout = p.ncart;
    for (int c = 0; c < p.ncart; c++) { p.nterm[c] = 1; p.out[c][0] = c; p.coef[c][0] = 1.0; }
  }
}

struct Eri {
  int id[4], am[4], flips[4], nbf[4], nbf_cart[4], nroots;
  bool direct_pure;
  Proj proj[4];
  std::vector<double> ints, gijkl, gnkl, gnm, b00, b01, b10, c00, d00, abv, PQ, PB, QD, dij, dkl, rw;
  int idx[3][MXCART][4];
  double quartet_cutoff;
  double mu2_1 = 0.0;  // 1/mu^2 for Erf-attenuated integrals (int_rys.F90:179-181), 0 = regular
  void init(int maxang, const Cutoffs &c) {  // int_rys.F90:62-99
    quartet_cutoff = c.pair2;  // :74
    int mxrys = (4 * maxang + 2) / 2, mxcart = maxang + 1, mxbra = 2 * mxcart - 1;
    int nc = ncart(maxang);
    ints.assign((size_t)nc * nc * nc * nc, 0.0);
    gijkl.resize((size_t)mxcart * mxcart * mxcart * mxcart * MAXCONTR * 3);
    gnkl.resize((size_t)mxcart * mxcart * mxbra * MAXCONTR * 3);
    gnm.resize((size_t)mxbra * mxbra * MAXCONTR * 3);
    b00.resize(mxrys * MAXCONTR); b01 = b00; b10 = b00;
    c00.resize(mxrys * MAXCONTR * 3); d00 = c00;
    abv.resize(6 * MAXCONTR); PQ.resize(3 * MAXCONTR); PB = PQ; QD = PQ; dij = PQ; dkl = PQ;
    rw.resize(2 * mxrys * MAXCONTR);
  }
};

// int_rys.F90:123-154
void set_ids(Eri &g, const Basis &b, const int id[4]) {
  int fl[4] = {0, 1, 2, 3}, am[4];
  for (int s = 0; s < 4; s++) am[s] = b.am[id[s]];
  if (am[0] > am[1]) { std::swap(fl[0], fl[1]); std::swap(am[0], am[1]); }
  if (am[2] > am[3]) { std::swap(fl[2], fl[3]); std::swap(am[2], am[3]); }
  if (am[0] + am[1] > am[2] + am[3]) {
    int f2[4] = {fl[2], fl[3], fl[0], fl[1]}, a2[4] = {am[2], am[3], am[0], am[1]};
    for (int s = 0; s < 4; s++) { fl[s] = f2[s]; am[s] = a2[s]; }
  }
  for (int s = 0; s < 4; s++) { g.flips[s] = fl[s]; g.id[s] = id[fl[s]]; g.am[s] = am[s]; }
}

// one primitive batch: int_rys.F90:351-382 (compute) and the routines it calls
void compute_batch(Eri &g, int ng, int nmax, int mmax) {
  const int nr = g.nroots;
  // compute_rys_rw :449-469
  for (int ig = 0; ig < ng; ig++) {
    double u[MXRYS], w[MXRYS];
    rys_general(g.abv[6 * ig + 5], nr, u, w);
    for (int r = 0; r < nr; r++) { g.rw[(r * ng + ig) * 2] = u[r]; g.rw[(r * ng + ig) * 2 + 1] = w[r]; }
  }
  const int ngnr = ng * nr;
  // layout: gnm[(m*nmax + n)*3 + xyz][nr][ng] (Fortran gnm(ng,nr,3,nmax,mmax))
  auto GNM = [&](int ig, int r, int x, int n, int m) -> double & {
    return g.gnm[((size_t)((m * nmax + n) * 3 + x) * nr + r) * ng + ig];
  };
  // compute_coefficients :471-527
  for (int r = 0; r < nr; r++)
    for (int ig = 0; ig < ng; ig++) {
      const double *abv = &g.abv[6 * ig];
      double a1 = abv[0], b1 = abv[1], rho = abv[2], pfac = abv[3], ab1 = abv[4];
      double uu = g.rw[(r * ng + ig) * 2], ww = g.rw[(r * ng + ig) * 2 + 1];
      GNM(ig, r, 0, 0, 0) = ww * pfac;
      GNM(ig, r, 1, 0, 0) = 1.0;
      GNM(ig, r, 2, 0, 0) = 1.0;
      double t2 = uu / (uu + 1);
      double t2ar = t2 * rho * a1, t2br = t2 * rho * b1;
      int q = r * ng + ig;
      g.b00[q] = 0.5 * ab1 * t2;
      g.b01[q] = 0.5 * b1 * (1.0 - t2br);
      g.b10[q] = 0.5 * a1 * (1.0 - t2ar);
      for (int x = 0; x < 3; x++) {
        if (mmax > 1) g.d00[x * ngnr + q] = g.QD[3 * ig + x] + t2br * g.PQ[3 * ig + x];
        if (nmax > 1) g.c00[x * ngnr + q] = g.PB[3 * ig + x] - t2ar * g.PQ[3 * ig + x];
      }
    }
  // compute_xyz_p0q0 :529-617 (2-D VRR), vectorised over q = (r,ig)
  auto G = [&](int q, int x, int n, int m) -> double & { return g.gnm[(size_t)((m * nmax + n) * 3 + x) * ngnr + q]; };
  if (std::max(nmax, mmax) > 1) {
    for (int x = 0; x < 3; x++)
      for (int q = 0; q < ngnr; q++) {
        double c0 = nmax > 1 ? g.c00[x * ngnr + q] : 0, d0 = mmax > 1 ? g.d00[x * ngnr + q] : 0;
        double B00 = g.b00[q], B01 = g.b01[q], B10 = g.b10[q];
        if (nmax > 1) G(q, x, 1, 0) = c0 * G(q, x, 0, 0);
        if (mmax > 1) {
          G(q, x, 0, 1) = d0 * G(q, x, 0, 0);
          if (nmax > 1) G(q, x, 1, 1) = B00 * G(q, x, 0, 0) + d0 * G(q, x, 1, 0);
        }
        if (nmax > 2) {
          for (int n = 2; n <= nmax - 1; n++) G(q, x, n, 0) = (n - 1) * B10 * G(q, x, n - 2, 0) + c0 * G(q, x, n - 1, 0);
          if (mmax > 1)
            for (int n = 2; n <= nmax - 1; n++) G(q, x, n, 1) = n * B00 * G(q, x, n - 1, 0) + d0 * G(q, x, n, 0);
        }
        if (mmax >= 3) {
          for (int m = 2; m <= mmax - 1; m++) G(q, x, 0, m) = (m - 1) * B01 * G(q, x, 0, m - 2) + d0 * G(q, x, 0, m - 1);
          if (nmax >= 2) {
            for (int m = 2; m <= mmax - 1; m++) G(q, x, 1, m) = m * B00 * G(q, x, 0, m - 1) + c0 * G(q, x, 0, m);
            if (nmax >= 3)
              for (int m = 2; m <= mmax - 1; m++)
                for (int n = 2; n <= nmax - 1; n++)
                  G(q, x, n, m) = (n - 1) * B10 * G(q, x, n - 2, m) + c0 * G(q, x, n - 1, m) + m * B00 * G(q, x, n - 1, m - 1);
          }
        }
      }
  }
  // compute_xyz_ijkl :619-661 (HRR).  gnkl(q,xyz,l,k,n), ijkl(q,xyz,l,k,j,i)
  const int ni = g.am[0] + 1, nj = g.am[1] + 1, nk = g.am[2] + 1, nl = g.am[3] + 1;
  auto NKL = [&](int q, int x, int l, int k, int n) -> double & {
    return g.gnkl[(size_t)((((n * nk + k) * nl + l) * 3) + x) * ngnr + q];
  };
  auto IJKL = [&](int q, int x, int l, int k, int j, int i) -> double & {
    return g.gijkl[(size_t)(((((i * nj + j) * nk + k) * nl + l) * 3) + x) * ngnr + q];
  };
  for (int k = 0; k < nk; k++) {
    for (int l = 0; l < nl; l++)
      for (int n = 0; n < nmax; n++)
        for (int x = 0; x < 3; x++)
          for (int q = 0; q < ngnr; q++) NKL(q, x, l, k, n) = G(q, x, n, l);
    if (k == nk - 1) break;
    int m1 = mmax - (k + 1);
    for (int x = 0; x < 3; x++)
      for (int m = 0; m < m1; m++)
        for (int n = 0; n < nmax; n++)
          for (int q = 0; q < ngnr; q++) G(q, x, n, m) = g.dkl[3 * (q % ng) + x] * G(q, x, n, m) + G(q, x, n, m + 1);
  }
  for (int i = 0; i < ni; i++) {
    for (int j = 0; j < nj; j++)
      for (int k = 0; k < nk; k++)
        for (int l = 0; l < nl; l++)
          for (int x = 0; x < 3; x++)
            for (int q = 0; q < ngnr; q++) IJKL(q, x, l, k, j, i) = NKL(q, x, l, k, j);
    if (i == ni - 1) break;
    int n1 = nmax - (i + 1);
    for (int x = 0; x < 3; x++)
      for (int n = 0; n < n1; n++)
        for (int k = 0; k < nk; k++)
          for (int l = 0; l < nl; l++)
            for (int q = 0; q < ngnr; q++)
              NKL(q, x, l, k, n) = g.dij[3 * (q % ng) + x] * NKL(q, x, l, k, n) + NKL(q, x, l, k, n + 1);
  }
  // compute_ints :677-713 / compute_ints_direct_pure :715-785
  const int n1 = g.nbf_cart[0], n2 = g.nbf_cart[1], n3 = g.nbf_cart[2], n4 = g.nbf_cart[3];
  const int o2 = g.nbf[1], o3 = g.nbf[2], o4 = g.nbf[3];
  for (int i = 0; i < n1; i++)
    for (int j = 0; j < n2; j++)
      for (int k = 0; k < n3; k++)
        for (int l = 0; l < n4; l++) {
          int nx = g.idx[0][i][0] + g.idx[0][j][1] + g.idx[0][k][2] + g.idx[0][l][3];
          int ny = g.idx[1][i][0] + g.idx[1][j][1] + g.idx[1][k][2] + g.idx[1][l][3];
          int nz = g.idx[2][i][0] + g.idx[2][j][1] + g.idx[2][k][2] + g.idx[2][l][3];
          const double *X = &g.gijkl[(size_t)(nx * 3 + 0) * ngnr], *Y = &g.gijkl[(size_t)(ny * 3 + 1) * ngnr],
                       *Z = &g.gijkl[(size_t)(nz * 3 + 2) * ngnr];
          double s = 0;
          for (int q = 0; q < ngnr; q++) s += X[q] * Y[q] * Z[q];
          if (!g.direct_pure) {
            g.ints[((i * n2 + j) * n3 + k) * n4 + l] += s;
          } else {
            double val = s * pnrm2(g.am[0], i) * pnrm2(g.am[1], j) * pnrm2(g.am[2], k) * pnrm2(g.am[3], l);
            if (val == 0.0) continue;
            const Proj &p1 = g.proj[0], &p2 = g.proj[1], &p3 = g.proj[2], &p4 = g.proj[3];
            for (int ti = 0; ti < p1.nterm[i]; ti++) {
              double vi = val * p1.coef[i][ti];
              for (int tj = 0; tj < p2.nterm[j]; tj++) {
                double vij = vi * p2.coef[j][tj];
                for (int tk = 0; tk < p3.nterm[k]; tk++) {
                  double vijk = vij * p3.coef[k][tk];
                  for (int tl = 0; tl < p4.nterm[l]; tl++)
                    g.ints[((p1.out[i][ti] * o2 + p2.out[j][tj]) * o3 + p3.out[k][tk]) * o4 + p4.out[l][tl]] +=
                        vijk * p4.coef[l][tl];
                }
              }
            }
          }
        }
}

// int_rys.F90:156-276 (int2_rys_compute, direct_pure=.true. as called from int2.F90:1124-1131)
// Output block g.ints(l,k,j,i) [l fastest] in the permuted (flips) shell order, unit-normalised
// and pure-projected (normalisation for the non-pure case: normalize_ints int2.F90:1187-1207).
bool rys_compute(Eri &g, const Basis &b, const Pairs &pp) {
  for (int s = 0; s < 4; s++) g.nbf[s] = g.nbf_cart[s] = ncart(g.am[s]);
  g.nroots = (g.am[0] + g.am[1] + g.am[2] + g.am[3] + 2) / 2;  // :406
  const int nj = g.am[1] + 1, nk = g.am[2] + 1, nl = g.am[3] + 1;
  for (int c = 0; c < g.nbf[0]; c++) { g.idx[0][c][0] = CX[g.am[0]][c] * nj * nk * nl; g.idx[1][c][0] = CY[g.am[0]][c] * nj * nk * nl; g.idx[2][c][0] = CZ[g.am[0]][c] * nj * nk * nl; }
  for (int c = 0; c < g.nbf[1]; c++) { g.idx[0][c][1] = CX[g.am[1]][c] * nk * nl; g.idx[1][c][1] = CY[g.am[1]][c] * nk * nl; g.idx[2][c][1] = CZ[g.am[1]][c] * nk * nl; }
  for (int c = 0; c < g.nbf[2]; c++) { g.idx[0][c][2] = CX[g.am[2]][c] * nl; g.idx[1][c][2] = CY[g.am[2]][c] * nl; g.idx[2][c][2] = CZ[g.am[2]][c] * nl; }
  for (int c = 0; c < g.nbf[3]; c++) { g.idx[0][c][3] = CX[g.am[3]][c]; g.idx[1][c][3] = CY[g.am[3]][c]; g.idx[2][c][3] = CZ[g.am[3]][c]; }
  // prepare_direct_pure :319-349
  g.direct_pure = false;
  if (b.harmonic_active) {
    bool any = false;
    for (int s = 0; s < 4; s++) any |= (b.harm[g.id[s]] == 1 && g.am[s] >= 2);
    if (any) {
      g.direct_pure = true;
      for (int s = 0; s < 4; s++) { init_proj(g.am[s], b.harm[g.id[s]] == 1, g.proj[s]); g.nbf[s] = g.proj[s].nout; }
    }
  }
  int i1 = std::max(g.id[0], g.id[1]), i2 = std::min(g.id[0], g.id[1]);
  int npp_p = pp.cnt[tri(i1, i2)], ppid_p = pp.off[tri(i1, i2)];
  i1 = std::max(g.id[2], g.id[3]); i2 = std::min(g.id[2], g.id[3]);
  int npp_q = pp.cnt[tri(i1, i2)], ppid_q = pp.off[tri(i1, i2)];
  if (npp_p * npp_q == 0) return true;
  const int nmax = g.am[0] + g.am[1] + 1, mmax = g.am[2] + g.am[3] + 1;
  const int maxgg = MAXCONTR / g.nroots;
  bool first = true, zero = true;
  int ng = 0;
  auto flush = [&]() {
    if (ng == 0) return;
    if (first) std::fill(g.ints.begin(), g.ints.begin() + (size_t)g.nbf[0] * g.nbf[1] * g.nbf[2] * g.nbf[3], 0.0);
    first = false;
    compute_batch(g, ng, nmax, mmax);
    ng = 0;
    zero = false;
  };
  for (int klg = 0; klg < npp_q; klg++) {
    int q = ppid_q + klg;
    double db = pp.k[q] * pp.ginv[q], bb = pp.g[q];
    for (int ijg = 0; ijg < npp_p; ijg++) {
      int p = ppid_p + ijg;
      double da = pp.k[p] * pp.ginv[p], aa = pp.g[p];
      double ab = (aa + bb) + aa * bb * g.mu2_1;  // 2nd term: Erf-attenuated integrals, int_rys.F90:225-227
      double pfac = da * db, test = pfac * pfac;
      if (test < g.quartet_cutoff * ab) continue;
      double aandb1 = 1.0 / ab, rho = aa * bb * aandb1;
      double *abv = &g.abv[6 * ng];
      abv[0] = pp.ginv[p]; abv[1] = pp.ginv[q]; abv[2] = rho; abv[3] = pfac * std::sqrt(aandb1); abv[4] = aandb1;
      double r2 = 0;
      for (int x = 0; x < 3; x++) { double d = pp.p[3 * p + x] - pp.p[3 * q + x]; g.PQ[3 * ng + x] = d; r2 += d * d; }
      abv[5] = rho * r2;
      for (int x = 0; x < 3; x++) {
        if (nmax > 1) g.PB[3 * ng + x] = pp.pb[3 * p + x];
        if (mmax > 1) g.QD[3 * ng + x] = pp.pb[3 * q + x];
        g.dij[3 * ng + x] = pp.pa[3 * p + x] - pp.pb[3 * p + x];
        g.dkl[3 * ng + x] = pp.pa[3 * q + x] - pp.pb[3 * q + x];
      }
      ng++;
      if (ng == maxgg) flush();
    }
  }
  flush();
  if (zero) return true;
  if (!g.direct_pure) {  // normalize_ints, int2.F90:1187-1207
    int n1 = g.nbf[0], n2 = g.nbf[1], n3 = g.nbf[2], n4 = g.nbf[3];
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n2; j++)
        for (int k = 0; k < n3; k++)
          for (int l = 0; l < n4; l++)
            g.ints[((i * n2 + j) * n3 + k) * n4 + l] = g.ints[((i * n2 + j) * n3 + k) * n4 + l] * pnrm2(g.am[0], i) *
                                                        pnrm2(g.am[1], j) * pnrm2(g.am[2], k) * pnrm2(g.am[3], l);
  }
  return false;
}

// ------------------------------------------------------------------------------------------------
struct Oracle {
  Basis b;
  Cutoffs cut;
  Pairs pp;
  std::vector<double> schwarz;  // nshell x nshell
  double cutoff = 5e-11;
  // range-separated (CAM) second pass, int2.F90:538-584, 674-685: attenuated Schwarz matrix and mu of the active pass
  std::vector<double> schwarz_regular, schwarz_att;
  double mu = 0.0;  // > 0: attenuated integrals erf(mu r)/r
};

// ints_exchange, int2.F90:1582-1737 (Rys branch)
void ints_exchange(Oracle &o, double mu = 0.0) {
  Cutoffs c;
  c.set(1.0e-15, 1.0e-17, 1.0e-17, 50.0);  // :1600-1604
  Pairs pp;
  build_pairs(o.b, c, pp);
  int lmax = *std::max_element(o.b.am.begin(), o.b.am.end());
  int ns = o.b.nshell;
  o.schwarz.assign((size_t)ns * ns, 0.0);
#pragma omp parallel
  {
    Eri g;
    g.init(lmax, c);
    g.mu2_1 = mu > 0 ? 1.0 / (mu * mu) : 0.0;  // ints_exchange(..., mu2), int2.F90:678
#pragma omp for schedule(dynamic, 4)
    for (int ish = 0; ish < ns; ish++)
      for (int jsh = 0; jsh <= ish; jsh++) {
        int ids[4] = {ish, jsh, ish, jsh};
        set_ids(g, o.b, ids);
        bool zero = rys_compute(g, o.b, pp);
        double vmax = 0;
        if (!zero) {
          size_t n = (size_t)g.nbf[0] * g.nbf[1] * g.nbf[2] * g.nbf[3];
          for (size_t t = 0; t < n; t++) vmax = std::max(vmax, std::fabs(g.ints[t]));
        }
        o.schwarz[(size_t)ish * ns + jsh] = o.schwarz[(size_t)jsh * ns + ish] = std::sqrt(vmax);
      }
  }
}

// Consumers -------------------------------------------------------------------------------------
enum Kind { RHF = 0, UROHF = 1, TD = 2, MRSF = 3, COLLECT = 4, TDGRD = 5, RPAGRD = 6, UMRSF = 7 };

struct Consumer {
  int kind = RHF, nbf = 0, nfocks = 1;
  double se = 1, sc = 1;
  // RHF/UROHF: packed
  const double *d = nullptr;
  // TD: d2(nbf,nbf,nvec) column-major [mu + nbf*nu + nbf^2*v]
  int int_apb = 1, int_amb = 0, tda = 0, tda_coulomb = 0;
  // MRSF: d3(nvec, ncomp, nbf, nbf), v fastest
  int ncomp = 7;
  int cur_pass = 1;  // multipass (CAM): pass 2 of the MRSF consumer touches only component 7 (tdhf_mrsf_lib.F90:312-326)
  std::vector<double> ds;
  std::vector<double> dsh;
  double max_den = 0;
  // RPAGRD (tdhf_lib.F90:42-57): xpy/xmy/t (nbf,nbf,nspin,n*) column-major; outputs hpp | hpt | hmm packed behind each
  // other in the one accumulator f (offsets rp_off[0..2])
  const double *xpy = nullptr, *xmy = nullptr, *tt = nullptr;
  int nspin = 1, np = 0, nm = 0, nt = 0;
  size_t rp_off[3] = {0, 0, 0};
};

struct Buf {  // int2_storage_t, int2.F90:60-80
  std::vector<int16_t> ids;
  std::vector<double> ints;
  int ncur = 0, size = 50000;
  Buf() : ids(4 * 50000), ints(50000) {}
};

inline size_t tri1(int i, int j) { return (size_t)i * (i - 1) / 2 + j - 1; }  // 1-based ids -> 0-based offset

// int2.F90:1414-1484 / 1488-1578 / tdhf_lib.F90:140-224 / tdhf_mrsf_lib.F90:218-333
void update(const Consumer &c, Buf &buf, double *f, double *f2, std::vector<int16_t> *collect_ids,
            std::vector<double> *collect_vals) {
  const int nbf = c.nbf;
  const size_t ntri = (size_t)nbf * (nbf + 1) / 2, n2 = (size_t)nbf * nbf;
  if (c.kind == RHF) {
    double xval1 = c.se, xval4 = 4 * c.sc;
    for (int ifock = 0; ifock < c.nfocks; ifock++) {
      const double *d = c.d + ifock * ntri;
      double *F = f + ifock * ntri;
      for (int n = 0; n < buf.ncur; n++) {
        int ii = buf.ids[4 * n], jj = buf.ids[4 * n + 1], kk = buf.ids[4 * n + 2], ll = buf.ids[4 * n + 3];
        double val = buf.ints[n];
        size_t ij = tri1(ii, jj), ik = tri1(ii, kk), il = tri1(ii, ll), jk = tri1(jj, kk), jl = tri1(jj, ll), kl = tri1(kk, ll);
        if (jj < kk) jk = tri1(kk, jj);
        if (jj < ll) jl = tri1(ll, jj);
        double val1 = val * xval1, val4 = val * xval4;
        F[ij] += val4 * d[kl];
        F[kl] += val4 * d[ij];
        F[ik] -= val1 * d[jl];
        F[jl] -= val1 * d[ik];
        F[il] -= val1 * d[jk];
        F[jk] -= val1 * d[il];
      }
    }
  } else if (c.kind == UROHF) {
    double xval2 = 2 * c.se, xval4 = 4 * c.sc;
    const double *d1 = c.d, *d2 = c.d + ntri;
    double *F1 = f, *F2 = f + ntri;
    for (int n = 0; n < buf.ncur; n++) {
      int ii = buf.ids[4 * n], jj = buf.ids[4 * n + 1], kk = buf.ids[4 * n + 2], ll = buf.ids[4 * n + 3];
      double val = buf.ints[n];
      size_t ij = tri1(ii, jj), ik = tri1(ii, kk), il = tri1(ii, ll), jk = tri1(jj, kk), jl = tri1(jj, ll), kl = tri1(kk, ll);
      if (jj < kk) jk = tri1(kk, jj);
      if (jj < ll) jl = tri1(ll, jj);
      double val1 = val * xval2, val4 = val * xval4;
      double cij = val4 * (d1[ij] + d2[ij]), ckl = val4 * (d1[kl] + d2[kl]);
      F1[ij] += ckl; F1[kl] += cij;
      F1[ik] -= val1 * d1[jl]; F1[jl] -= val1 * d1[ik]; F1[il] -= val1 * d1[jk]; F1[jk] -= val1 * d1[il];
      F2[ij] += ckl; F2[kl] += cij;
      F2[ik] -= val1 * d2[jl]; F2[jl] -= val1 * d2[ik]; F2[il] -= val1 * d2[jk]; F2[jk] -= val1 * d2[il];
    }
  } else if (c.kind == TD) {
    double xval1 = c.se, cval2 = 2 * c.sc, cval4 = 4 * c.sc;
    for (int v = 0; v < c.nfocks; v++) {
      const double *P = c.d + v * n2;
      double *apb = f + v * n2, *amb = f2 + v * n2;
      auto D = [&](int a, int b) { return P[(a - 1) + (size_t)nbf * (b - 1)]; };
      auto A = [&](double *m, int a, int b) -> double & { return m[(a - 1) + (size_t)nbf * (b - 1)]; };
      for (int n = 0; n < buf.ncur; n++) {
        int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
        double val = buf.ints[n];
        if (c.tda) {
          double val1 = val * xval1, val2c = val * cval2;
          A(amb, i, k) -= val1 * D(j, l); A(amb, k, i) -= val1 * D(l, j);
          A(amb, i, l) -= val1 * D(j, k); A(amb, l, i) -= val1 * D(k, j);
          A(amb, j, k) -= val1 * D(i, l); A(amb, k, j) -= val1 * D(l, i);
          A(amb, j, l) -= val1 * D(i, k); A(amb, l, j) -= val1 * D(k, i);
          if (c.tda_coulomb) {
            A(amb, i, j) += val2c * (D(k, l) + D(l, k)); A(amb, j, i) += val2c * (D(k, l) + D(l, k));
            A(amb, k, l) += val2c * (D(i, j) + D(j, i)); A(amb, l, k) += val2c * (D(i, j) + D(j, i));
          }
        } else {
          double val1 = val * xval1, val4c = val * cval4;
          if (c.int_apb) {
            A(apb, i, j) += val4c * (D(k, l) + D(l, k));
            A(apb, k, l) += val4c * (D(i, j) + D(j, i));
            A(apb, i, k) -= val1 * (D(j, l) + D(l, j));
            A(apb, i, l) -= val1 * (D(j, k) + D(k, j));
            A(apb, j, k) -= val1 * (D(i, l) + D(l, i));
            A(apb, j, l) -= val1 * (D(i, k) + D(k, i));
          }
          if (c.int_amb) {
            A(amb, i, k) += val1 * (D(l, j) - D(j, l)); A(amb, i, l) += val1 * (D(k, j) - D(j, k));
            A(amb, j, k) += val1 * (D(l, i) - D(i, l)); A(amb, j, l) += val1 * (D(k, i) - D(i, k));
            A(amb, k, i) -= val1 * (D(l, j) - D(j, l)); A(amb, l, i) -= val1 * (D(k, j) - D(j, k));
            A(amb, k, j) -= val1 * (D(l, i) - D(i, l)); A(amb, l, j) -= val1 * (D(k, i) - D(i, k));
          }
        }
      }
    }
  } else if (c.kind == MRSF) {
    const int nf = c.nfocks, nc = c.ncomp;
    auto D3 = [&](int v, int cc, int a, int b) { return c.d[v + (size_t)nf * (cc + (size_t)nc * ((a - 1) + (size_t)nbf * (b - 1)))]; };
    auto DS = [&](int v, int cc, int a, int b) { return c.ds[v + (size_t)nf * (cc + (size_t)4 * ((a - 1) + (size_t)nbf * (b - 1)))]; };
    auto F3 = [&](int v, int cc, int a, int b) -> double & { return f[v + (size_t)nf * (cc + (size_t)nc * ((a - 1) + (size_t)nbf * (b - 1)))]; };
    for (int n = 0; n < buf.ncur; n++) {
      int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
      double val = buf.ints[n];
      double xval = val * c.se, cval = val * c.sc;
      if (c.cur_pass == 1)
      for (int cc = 0; cc < 4; cc++)
        for (int v = 0; v < nf; v++) {
          F3(v, cc, i, j) += cval * DS(v, cc, k, l); F3(v, cc, j, i) += cval * DS(v, cc, k, l);
          F3(v, cc, k, l) += cval * DS(v, cc, i, j); F3(v, cc, l, k) += cval * DS(v, cc, i, j);
        }
      for (int cc = (c.cur_pass == 2 ? 6 : 0); cc < nc; cc++)  // pass 2: exchange of component 7 only (:312-326)
        for (int v = 0; v < nf; v++) {
          F3(v, cc, i, k) -= xval * D3(v, cc, j, l); F3(v, cc, k, i) -= xval * D3(v, cc, l, j);
          F3(v, cc, i, l) -= xval * D3(v, cc, j, k); F3(v, cc, l, i) -= xval * D3(v, cc, k, j);
          F3(v, cc, j, k) -= xval * D3(v, cc, i, l); F3(v, cc, k, j) -= xval * D3(v, cc, l, i);
          F3(v, cc, j, l) -= xval * D3(v, cc, i, k); F3(v, cc, l, j) -= xval * D3(v, cc, k, i);
        }
    }
  } else if (c.kind == TDGRD) {
    // int2_tdgrd_data_t_update, tdhf_lib.F90:228-295: two spin blocks d2(:,:,1:2); A+B for both, A-B for the first
    double xval1 = 1 * c.se, xval2 = 2 * c.sc;
    const double *P1 = c.d, *P2 = c.d + n2;
    double *apb1 = f, *apb2 = f + n2, *amb1 = f2;
    auto D1 = [&](int a, int b) { return P1[(a - 1) + (size_t)nbf * (b - 1)]; };
    auto D2 = [&](int a, int b) { return P2[(a - 1) + (size_t)nbf * (b - 1)]; };
    auto A = [&](double *m, int a, int b) -> double & { return m[(a - 1) + (size_t)nbf * (b - 1)]; };
    for (int n = 0; n < buf.ncur; n++) {
      int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
      double val = buf.ints[n], val1 = val * xval1, val2 = val * xval2;
      if (c.int_apb) {
        double ckl = val2 * (D1(k, l) + D1(l, k) + D2(k, l) + D2(l, k)), cij = val2 * (D1(i, j) + D1(j, i) + D2(i, j) + D2(j, i));
        A(apb1, i, j) += ckl; A(apb1, k, l) += cij;
        A(apb2, i, j) += ckl; A(apb2, k, l) += cij;
        A(apb1, i, k) -= val1 * (D1(j, l) + D1(l, j)); A(apb1, i, l) -= val1 * (D1(j, k) + D1(k, j));
        A(apb1, j, k) -= val1 * (D1(i, l) + D1(l, i)); A(apb1, j, l) -= val1 * (D1(i, k) + D1(k, i));
        A(apb2, i, k) -= val1 * (D2(j, l) + D2(l, j)); A(apb2, i, l) -= val1 * (D2(j, k) + D2(k, j));
        A(apb2, j, k) -= val1 * (D2(i, l) + D2(l, i)); A(apb2, j, l) -= val1 * (D2(i, k) + D2(k, i));
      }
      if (c.int_amb) {
        A(amb1, i, k) += val1 * (D1(l, j) - D1(j, l)); A(amb1, i, l) += val1 * (D1(k, j) - D1(j, k));
        A(amb1, j, k) += val1 * (D1(l, i) - D1(i, l)); A(amb1, j, l) += val1 * (D1(k, i) - D1(i, k));
        A(amb1, k, i) -= val1 * (D1(l, j) - D1(j, l)); A(amb1, l, i) -= val1 * (D1(k, j) - D1(j, k));
        A(amb1, k, j) -= val1 * (D1(l, i) - D1(i, l)); A(amb1, l, j) -= val1 * (D1(k, i) - D1(i, k));
      }
    }
  } else if (c.kind == RPAGRD) {
    // int2_rpagrd_data_t_update / _hplus / _hminus, tdhf_lib.F90:1177-1320
    const size_t slab = n2 * c.nspin;
    auto hplus = [&](double *hp, const double *v) {
      auto V = [&](int a, int b, int s) { return v[(a - 1) + (size_t)nbf * (b - 1) + n2 * s]; };
      auto H = [&](int a, int b, int s) -> double & { return hp[(a - 1) + (size_t)nbf * (b - 1) + n2 * s]; };
      if (c.nspin == 1) {
        double xfact = 2 * c.se, cfact = 8 * c.sc;
        for (int n = 0; n < buf.ncur; n++) {
          int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
          double val = buf.ints[n], xval = val * xfact, cval = val * cfact;
          H(i, j, 0) += cval * V(l, k, 0); H(k, l, 0) += cval * V(j, i, 0);
          H(i, k, 0) -= xval * V(l, j, 0); H(i, l, 0) -= xval * V(k, j, 0);
          H(j, k, 0) -= xval * V(l, i, 0); H(j, l, 0) -= xval * V(k, i, 0);
        }
      } else {
        double xfact = 1 * c.se, cfact = 2 * c.sc;
        for (int n = 0; n < buf.ncur; n++) {
          int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
          double val = buf.ints[n], xval = val * xfact, cval = val * cfact;
          double ckl = cval * (V(k, l, 0) + V(l, k, 0) + V(k, l, 1) + V(l, k, 1));
          double cij = cval * (V(i, j, 0) + V(j, i, 0) + V(i, j, 1) + V(j, i, 1));
          for (int sp = 0; sp < 2; sp++) {
            H(i, j, sp) += ckl; H(k, l, sp) += cij;
            H(i, k, sp) -= xval * (V(j, l, sp) + V(l, j, sp)); H(i, l, sp) -= xval * (V(j, k, sp) + V(k, j, sp));
            H(j, k, sp) -= xval * (V(i, l, sp) + V(l, i, sp)); H(j, l, sp) -= xval * (V(i, k, sp) + V(k, i, sp));
          }
        }
      }
    };
    auto hminus = [&](double *hm, const double *v) {  // first spin block only, as written (:1297-1318)
      auto V = [&](int a, int b) { return v[(a - 1) + (size_t)nbf * (b - 1)]; };
      auto H = [&](int a, int b) -> double & { return hm[(a - 1) + (size_t)nbf * (b - 1)]; };
      double xfact = c.se;
      for (int n = 0; n < buf.ncur; n++) {
        int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
        double xval = buf.ints[n] * xfact;
        H(i, k) += xval * (V(l, j) - V(j, l)); H(i, l) += xval * (V(k, j) - V(j, k));
        H(j, k) += xval * (V(l, i) - V(i, l)); H(j, l) += xval * (V(k, i) - V(i, k));
        H(k, i) -= xval * (V(l, j) - V(j, l)); H(l, i) -= xval * (V(k, j) - V(j, k));
        H(k, j) -= xval * (V(l, i) - V(i, l)); H(l, j) -= xval * (V(k, i) - V(i, k));
      }
    };
    for (int q = 0; q < c.np; q++) hplus(f + c.rp_off[0] + q * slab, c.xpy + q * slab);
    for (int q = 0; q < c.nt; q++) hplus(f + c.rp_off[1] + q * slab, c.tt + q * slab);
    for (int q = 0; q < c.nm; q++) hminus(f + c.rp_off[2] + q * slab, c.xmy + q * slab);
  } else if (c.kind == UMRSF) {
    // int2_umrsf_data_t_update, tdhf_mrsf_lib.F90:337-426 (11 components; 1-8 Coulomb + exchange, 9-10 the mixed-spin
    // exchange permutation, 11 plain exchange; pass 2 = component 11 only)
    const int nf = c.nfocks, nc = c.ncomp;
    auto D3 = [&](int v, int cc, int a, int b) { return c.d[v + (size_t)nf * (cc + (size_t)nc * ((a - 1) + (size_t)nbf * (b - 1)))]; };
    auto F3 = [&](int v, int cc, int a, int b) -> double & { return f[v + (size_t)nf * (cc + (size_t)nc * ((a - 1) + (size_t)nbf * (b - 1)))]; };
    for (int n = 0; n < buf.ncur; n++) {
      int i = buf.ids[4 * n], j = buf.ids[4 * n + 1], k = buf.ids[4 * n + 2], l = buf.ids[4 * n + 3];
      double val = buf.ints[n], xval = val * c.se, cval = val * c.sc;
      for (int v = 0; v < nf; v++) {
        if (c.cur_pass == 1) {
          for (int cc = 0; cc < 8; cc++) {
            F3(v, cc, i, j) += cval * D3(v, cc, k, l); F3(v, cc, k, l) += cval * D3(v, cc, i, j);
            F3(v, cc, i, j) += cval * D3(v, cc, l, k); F3(v, cc, l, k) += cval * D3(v, cc, i, j);
            F3(v, cc, j, i) += cval * D3(v, cc, k, l); F3(v, cc, k, l) += cval * D3(v, cc, j, i);
            F3(v, cc, j, i) += cval * D3(v, cc, l, k); F3(v, cc, l, k) += cval * D3(v, cc, j, i);
            F3(v, cc, i, k) -= xval * D3(v, cc, j, l); F3(v, cc, k, i) -= xval * D3(v, cc, l, j);
            F3(v, cc, i, l) -= xval * D3(v, cc, j, k); F3(v, cc, l, i) -= xval * D3(v, cc, k, j);
            F3(v, cc, j, k) -= xval * D3(v, cc, i, l); F3(v, cc, k, j) -= xval * D3(v, cc, l, i);
            F3(v, cc, j, l) -= xval * D3(v, cc, i, k); F3(v, cc, l, j) -= xval * D3(v, cc, k, i);
          }
          for (int cc = 8; cc < 10; cc++) {
            F3(v, cc, i, l) -= xval * D3(v, cc, k, j); F3(v, cc, l, i) -= xval * D3(v, cc, j, k);
            F3(v, cc, k, j) -= xval * D3(v, cc, i, l); F3(v, cc, j, k) -= xval * D3(v, cc, l, i);
            F3(v, cc, i, k) -= xval * D3(v, cc, l, j); F3(v, cc, k, i) -= xval * D3(v, cc, j, l);
            F3(v, cc, l, j) -= xval * D3(v, cc, i, k); F3(v, cc, j, l) -= xval * D3(v, cc, k, i);
          }
        }
        const int cc = 10;
        F3(v, cc, i, k) -= xval * D3(v, cc, j, l); F3(v, cc, k, i) -= xval * D3(v, cc, l, j);
        F3(v, cc, i, l) -= xval * D3(v, cc, j, k); F3(v, cc, l, i) -= xval * D3(v, cc, k, j);
        F3(v, cc, j, k) -= xval * D3(v, cc, i, l); F3(v, cc, k, j) -= xval * D3(v, cc, l, i);
        F3(v, cc, j, l) -= xval * D3(v, cc, i, k); F3(v, cc, l, j) -= xval * D3(v, cc, k, i);
      }
    }
  } else if (c.kind == COLLECT) {
    for (int n = 0; n < buf.ncur; n++) {
      for (int t = 0; t < 4; t++) collect_ids->push_back(buf.ids[4 * n + t]);
      collect_vals->push_back(buf.ints[n]);
    }
  }
  buf.ncur = 0;
}

// storeints, int2.F90:1741-1865 (C1: weight = 1)
template <class Flush>
void storeints(const Basis &b, const Eri &g, const int ids_in[4], Buf &buf, double cutoff, long &nint, Flush flush) {
  int ids[4];
  for (int s = 0; s < 4; s++) ids[s] = ids_in[g.flips[s]];
  const int *nbf = g.nbf;
  bool same = ids[0] == ids[2] && ids[1] == ids[3], iandj = ids[0] == ids[1], kandl = ids[2] == ids[3];
  int loci = b.aooff[ids[0]], locj = b.aooff[ids[1]], lock = b.aooff[ids[2]], locl = b.aooff[ids[3]];  // 0-based offsets
  int nij = 0, maxj = nbf[1];
  for (int i = 1; i <= nbf[0]; i++) {
    if (iandj) maxj = i;
    for (int j = 1; j <= maxj; j++) {
      nij++;
      int nkl = nij, maxl = nbf[3];
      bool cycle_j = false;
      for (int k = 1; k <= nbf[2] && !cycle_j; k++) {
        if (kandl) maxl = k;
        if (same) {
          int itmp = std::min(maxl, nkl);
          if (itmp == 0) { cycle_j = true; break; }
          maxl = itmp;
          nkl -= itmp;
        }
        for (int l = 1; l <= maxl; l++) {
          double val = g.ints[(((size_t)(i - 1) * nbf[1] + (j - 1)) * nbf[2] + (k - 1)) * nbf[3] + (l - 1)];
          if (std::fabs(val) < cutoff) continue;
          nint++;
          int i1 = i + loci, j1 = j + locj, k1 = k + lock, l1 = l + locl;  // 1-based AO ids
          if (i1 < j1) std::swap(i1, j1);
          if (k1 < l1) std::swap(k1, l1);
          int ii = i1, jj = j1, kk = k1, ll = l1;
          if (ii < kk) { ii = k1; jj = l1; kk = i1; ll = j1; }
          else if (ii == kk && jj < ll) { ii = i1; jj = l1; kk = k1; ll = j1; }
          if (ii == jj) val *= 0.5;
          if (kk == ll) val *= 0.5;
          if (ii == kk && jj == ll) val *= 0.5;
          int n = buf.ncur++;
          buf.ids[4 * n] = (int16_t)ii; buf.ids[4 * n + 1] = (int16_t)jj; buf.ids[4 * n + 2] = (int16_t)kk; buf.ids[4 * n + 3] = (int16_t)ll;
          buf.ints[n] = val;
          if (buf.ncur == buf.size) flush();
        }
      }
    }
  }
}

// shlden int2.F90:999-1047 ; shltd tdhf_lib.F90:300-325 ; shell_den_screen_mrsf tdhf_mrsf_lib.F90:189-214
void init_screen(const Basis &b, Consumer &c) {
  int ns = b.nshell, nbf = b.nbf;
  c.dsh.assign((size_t)ns * ns, 0.0);
  size_t ntri = (size_t)nbf * (nbf + 1) / 2, n2 = (size_t)nbf * nbf;
  for (int si = 0; si < ns; si++)
    for (int sj = 0; sj <= si; sj++) {
      int mini = b.aooff[si], maxi = mini + b.naos[si] - 1, minj = b.aooff[sj], maxj = minj + b.naos[sj] - 1;
      double dmax = 0;
      if (c.kind == RHF || c.kind == UROHF) {
        for (int f = 0; f < c.nfocks; f++)
          for (int i = mini; i <= maxi; i++) {
            int mj = (si == sj) ? i : maxj;
            for (int j = minj; j <= mj; j++) dmax = std::max(dmax, std::fabs(c.d[f * ntri + tri(i, j)]));
          }
      } else if (c.kind == RPAGRD) {  // shlrpagrd over xpy, xmy (guarded by np, as written) and t: tdhf_lib.F90:1161-1173
        auto blockmax = [&](const double *v, int cnt) {
          for (size_t q = 0; q < (size_t)cnt * c.nspin; q++)
            for (int i = mini; i <= maxi; i++)
              for (int j = minj; j <= maxj; j++) dmax = std::max(dmax, std::fabs(v[q * n2 + j + (size_t)nbf * i]));
        };
        if (c.np > 0) blockmax(c.xpy, c.np);
        if (c.np > 0 && c.xmy) blockmax(c.xmy, c.nm);
        if (c.nt > 0) blockmax(c.tt, c.nt);
      } else if (c.kind == TD || c.kind == TDGRD) {  // da(minj:maxj, mini:maxi, :)
        for (int v = 0; v < c.nfocks; v++)
          for (int i = mini; i <= maxi; i++)
            for (int j = minj; j <= maxj; j++) dmax = std::max(dmax, std::fabs(c.d[v * n2 + j + (size_t)nbf * i]));
      } else if (c.kind == MRSF || c.kind == UMRSF) {  // da(:, minj:maxj, mini:maxi) with da = d3 viewed (nvec*ncomp, nbf, nbf)
        int nm = c.nfocks * c.ncomp;
        for (int i = mini; i <= maxi; i++)
          for (int j = minj; j <= maxj; j++)
            for (int m = 0; m < nm; m++) dmax = std::max(dmax, std::fabs(c.d[m + (size_t)nm * (j + (size_t)nbf * i)]));
      } else {
        dmax = 1.0;
      }
      c.dsh[(size_t)si * ns + sj] = c.dsh[(size_t)sj * ns + si] = dmax;
    }
  c.max_den = 0;
  for (double v : c.dsh) c.max_den = std::max(c.max_den, std::fabs(v));
}

// int2_build_shell_pair_map, int2.F90:864-921
void build_pair_map(const Oracle &o, std::vector<int> &pi, std::vector<int> &pj) {
  const int ns = o.b.nshell, NCLASS = 64, OFFS = 42;
  std::vector<long> cnt(NCLASS + 1, 0), off(NCLASS + 1, 0);
  auto cls = [&](int i, int j) {  // i,j 0-based; reference uses 1-based i in the cost
    double sw = std::max(o.schwarz[(size_t)i * ns + j], 1.0e-30);
    double cost = sw * double(ncart(o.b.am[i]) * ncart(o.b.am[j])) * double(i + 1);
    return NCLASS - std::max(0, std::min(NCLASS, (int)(std::log(cost) / std::log(2.0)) + OFFS));
  };
  for (int i = ns - 1; i >= 0; i--) for (int j = 0; j <= i; j++) cnt[cls(i, j)]++;
  for (int c = 1; c <= NCLASS; c++) off[c] = off[c - 1] + cnt[c - 1];
  pi.resize((size_t)ns * (ns + 1) / 2);
  pj.resize(pi.size());
  for (int i = ns - 1; i >= 0; i--)
    for (int j = 0; j <= i; j++) { long p = off[cls(i, j)]++; pi[p] = i; pj[p] = j; }
}

// int2_twoei, int2.F90:589-923 (C1, single pass, no CAM).  pair_lo/pair_hi select a contiguous
// slice of the cost-sorted bra-pair list (bounded samples for the timed CPU baseline); stride/offset
// reproduce the MPI cyclic split `mod(ij_pair, size) == rank` (int2.F90:759-761).
struct RunStats { long nschwz = 0, nshq = 0, nint = 0; double flops = 0; };

void twoei(Oracle &o, Consumer &c, double *f, double *f2, size_t fsize, size_t f2size, int nthreads, long pair_lo,
           long pair_hi, int stride, int offset, RunStats &st, std::vector<int> *qlist,
           std::vector<int16_t> *collect_ids, std::vector<double> *collect_vals) {
  const Basis &b = o.b;
  const int ns = b.nshell;
  init_screen(b, c);
  std::vector<int> pi, pj;
  build_pair_map(o, pi, pj);
  long npairs = (long)pi.size();
  if (pair_hi < 0 || pair_hi > npairs) pair_hi = npairs;
  if (pair_lo < 0) pair_lo = 0;
  int lmax = *std::max_element(b.am.begin(), b.am.end());
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
  if (qlist || collect_ids) nthreads = 1;
  std::vector<std::vector<double>> fth(nthreads), f2th(nthreads);  // thread-private Fock copies, int2.F90:1328
  long nschwz = 0, nshq = 0, nint = 0;
  const double cutoff = o.cut.integral;
  const double *Q = o.schwarz.data();
#pragma omp parallel num_threads(nthreads) reduction(+ : nschwz, nshq, nint)
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    double *F = f, *F2 = f2;
    if (tid > 0) {
      fth[tid].assign(fsize, 0.0); F = fth[tid].data();
      if (f2size) { f2th[tid].assign(f2size, 0.0); F2 = f2th[tid].data(); }
    }
    Eri g;
    g.init(lmax, o.cut);
    g.mu2_1 = o.mu > 0 ? 1.0 / (o.mu * o.mu) : 0.0;  // mu2 = this%mu**2 when attenuated (int2.F90:1098-1101)
    Buf buf;
    auto flush = [&]() { update(c, buf, F, F2, collect_ids, collect_vals); };
#pragma omp for schedule(dynamic, 1)
    for (long ijp = pair_lo; ijp < pair_hi; ijp++) {
      if (stride > 1 && ((ijp + 1) % stride) != offset) continue;
      int i = pi[ijp], j = pj[ijp];
      double test = Q[(size_t)i * ns + j] * c.max_den;  // screen_ij :963-971
      if (test < cutoff) { nschwz += (long)(i + 1) * i / 2 + (j + 1); continue; }
      for (int k = 0; k <= i; k++) {
        int jork = (i == k) ? j : k;
        for (int l = 0; l <= jork; l++) {
          // screen_ijkl :975-986
          double res = Q[(size_t)i * ns + j] * Q[(size_t)k * ns + l];
          const double *dsh = c.dsh.data();
          double m = std::max({4 * dsh[(size_t)i * ns + j], 4 * dsh[(size_t)k * ns + l], dsh[(size_t)j * ns + l],
                               dsh[(size_t)j * ns + k], dsh[(size_t)i * ns + l], dsh[(size_t)i * ns + k]});
          res = res * m;
          if (res < cutoff) { nschwz++; continue; }
          nshq++;
          if (qlist) { qlist->push_back(i); qlist->push_back(j); qlist->push_back(k); qlist->push_back(l); }
          int ids[4] = {i, j, k, l};
          set_ids(g, b, ids);
          bool zero = rys_compute(g, b, o.pp);
          if (zero) continue;
          storeints(b, g, ids, buf, cutoff, nint, flush);
        }
      }
    }
    flush();
  }
  for (int t = 1; t < nthreads; t++) {  // parallel_stop :1383-1400
    if (!fth[t].empty()) for (size_t q = 0; q < fsize; q++) f[q] += fth[t][q];
    if (!f2th[t].empty()) for (size_t q = 0; q < f2size; q++) f2[q] += f2th[t][q];
  }
  st.nschwz = nschwz; st.nshq = nshq; st.nint = nint;
}

// ------------------------------------------------------------------------------------------------
// One-electron integrals (NOT on the hot path; needed only so the oracle can run an SCF and be
// pinned against the reference's golden energies).  Obara-Saika 1-D recurrences for S and T, Rys
// quadrature for V.  Output: unit-normalised (and pure-projected) square matrices, row-major.
void shell_transform(const Basis &b, int sh, std::vector<double> &T, int &nout) {
  // T[out][cart]: unit normalisation x pure projection
  int l = b.am[sh], nc = ncart(l);
  Proj p;
  init_proj(l, b.harmonic_active && b.harm[sh] == 1, p);
  nout = p.nout;
  T.assign((size_t)nout * nc, 0.0);
  for (int c = 0; c < nc; c++)
    for (int t = 0; t < p.nterm[c]; t++) T[(size_t)p.out[c][t] * nc + c] += p.coef[c][t] * pnrm2(l, c);
}

void int1e(const Basis &b, int natom, const double *Z, const double *xyz, double *S, double *Tk, double *V) {
  const int nbf = b.nbf;
  for (int si = 0; si < b.nshell; si++)
    for (int sj = 0; sj < b.nshell; sj++) {
      int la = b.am[si], lb = b.am[sj], na = ncart(la), nb = ncart(lb);
      const double *A = &b.cen[3 * si], *B = &b.cen[3 * sj];
      std::vector<double> s(na * nb, 0.0), t(na * nb, 0.0), v(na * nb, 0.0);
      double ab2 = 0;
      for (int x = 0; x < 3; x++) ab2 += (A[x] - B[x]) * (A[x] - B[x]);
      for (int p1 = 0; p1 < b.ncontr[si]; p1++)
        for (int p2 = 0; p2 < b.ncontr[sj]; p2++) {
          double a = b.ex[b.goff[si] + p1], bb = b.ex[b.goff[sj] + p2], cc = b.cc[b.goff[si] + p1] * b.cc[b.goff[sj] + p2];
          double z = a + bb, zi = 1 / z, pref = std::exp(-a * bb * ab2 * zi) * cc;
          double P[3], PA[3], PB[3];
          for (int x = 0; x < 3; x++) { P[x] = (a * A[x] + bb * B[x]) * zi; PA[x] = P[x] - A[x]; PB[x] = P[x] - B[x]; }
          // 1-D overlaps up to (la, lb+2)
          double ov[3][8][8];
          for (int x = 0; x < 3; x++) {
            ov[x][0][0] = std::sqrt(M_PI * zi);
            for (int i = 0; i <= la; i++) {
              if (i > 0) ov[x][i][0] = PA[x] * ov[x][i - 1][0] + (i > 1 ? (i - 1) * 0.5 * zi * ov[x][i - 2][0] : 0);
              for (int j = 1; j <= lb + 2; j++)
                ov[x][i][j] = PB[x] * ov[x][i][j - 1] + (j > 1 ? (j - 1) * 0.5 * zi * ov[x][i][j - 2] : 0) +
                              (i > 0 ? i * 0.5 * zi * ov[x][i - 1][j - 1] : 0);
            }
          }
          auto kin = [&](int x, int i, int j) {  // -1/2 <i| d2/dx2 |j>
            double r = -2 * bb * bb * ov[x][i][j + 2] + bb * (2 * j + 1) * ov[x][i][j];
            if (j >= 2) r -= 0.5 * j * (j - 1) * ov[x][i][j - 2];
            return r;
          };
          for (int ca = 0; ca < na; ca++)
            for (int cb = 0; cb < nb; cb++) {
              int ax = CX[la][ca], ay = CY[la][ca], az = CZ[la][ca], bx = CX[lb][cb], by = CY[lb][cb], bz = CZ[lb][cb];
              double sx = ov[0][ax][bx], sy = ov[1][ay][by], sz = ov[2][az][bz];
              s[ca * nb + cb] += pref * sx * sy * sz;
              t[ca * nb + cb] += pref * (kin(0, ax, bx) * sy * sz + sx * kin(1, ay, by) * sz + sx * sy * kin(2, az, bz));
            }
          // nuclear attraction by Rys quadrature
          int nr = (la + lb) / 2 + 1;
          for (int at = 0; at < natom; at++) {
            double PC[3], r2 = 0;
            for (int x = 0; x < 3; x++) { PC[x] = P[x] - xyz[3 * at + x]; r2 += PC[x] * PC[x]; }
            double u[MXRYS], w[MXRYS];
            rys_general(z * r2, nr, u, w);
            for (int r = 0; r < nr; r++) {
              double t2 = u[r] / (1 + u[r]);
              double gx[3][10][8];
              for (int x = 0; x < 3; x++) {
                double c00 = PA[x] - t2 * PC[x], b10 = 0.5 * zi * (1 - t2);
                double gn[16];
                gn[0] = 1;
                if (la + lb > 0) gn[1] = c00;
                for (int n = 2; n <= la + lb; n++) gn[n] = c00 * gn[n - 1] + (n - 1) * b10 * gn[n - 2];
                // HRR: (i, j) = (i+1, j-1) + (A-B)(i, j-1)
                double AB = A[x] - B[x];
                for (int n = 0; n <= la + lb; n++) gx[x][n][0] = gn[n];
                for (int jj = 1; jj <= lb; jj++)
                  for (int n = 0; n <= la + lb - jj; n++) gx[x][n][jj] = gx[x][n + 1][jj - 1] + AB * gx[x][n][jj - 1];
              }
              double fac = -Z[at] * 2 * M_PI * zi * pref * w[r];
              for (int ca = 0; ca < na; ca++)
                for (int cb = 0; cb < nb; cb++)
                  v[ca * nb + cb] += fac * gx[0][CX[la][ca]][CX[lb][cb]] * gx[1][CY[la][ca]][CY[lb][cb]] * gx[2][CZ[la][ca]][CZ[lb][cb]];
            }
          }
        }
      std::vector<double> Ta, Tb;
      int oa, ob;
      shell_transform(b, si, Ta, oa);
      shell_transform(b, sj, Tb, ob);
      for (int m = 0; m < 3; m++) {
        std::vector<double> &src = (m == 0 ? s : (m == 1 ? t : v));
        double *dst = (m == 0 ? S : (m == 1 ? Tk : V));
        for (int ia = 0; ia < oa; ia++)
          for (int ib = 0; ib < ob; ib++) {
            double acc = 0;
            for (int ca = 0; ca < na; ca++)
              for (int cb = 0; cb < nb; cb++) acc += Ta[ia * na + ca] * Tb[ib * nb + cb] * src[ca * nb + cb];
            dst[(size_t)(b.aooff[si] + ia) * nbf + b.aooff[sj] + ib] = acc;
          }
      }
    }
}

}  // namespace

// ================================================================================================ C API
extern "C" {

void *orc_create(int nshell, const int *am, const int *ncontr, const int *goff, const int *aooff, const int *naos,
                 const int *harm, const double *ex, const double *cc, const double *centers, int harmonic_active) {
  Oracle *o = new Oracle;
  Basis &b = o->b;
  b.nshell = nshell;
  b.am.assign(am, am + nshell); b.ncontr.assign(ncontr, ncontr + nshell); b.goff.assign(goff, goff + nshell);
  b.aooff.assign(aooff, aooff + nshell); b.naos.assign(naos, naos + nshell); b.harm.assign(harm, harm + nshell);
  b.nprim = goff[nshell - 1] + ncontr[nshell - 1];
  b.ex.assign(ex, ex + b.nprim); b.cc.assign(cc, cc + b.nprim); b.cen.assign(centers, centers + 3 * nshell);
  b.nbf = aooff[nshell - 1] + naos[nshell - 1];
  b.harmonic_active = harmonic_active;
  return o;
}
void orc_destroy(void *h) { delete (Oracle *)h; }

// int2_compute_t%init cutoffs + pair table: int2.F90:245-289
void orc_set_cutoff(void *h, double cutoff) {
  Oracle *o = (Oracle *)h;
  o->cutoff = cutoff;
  o->cut.set(cutoff, 1.0e-2 * cutoff, 1.0e-4 * cutoff, 25.0 * std::log(10.0));
  build_pairs(o->b, o->cut, o->pp);
}
long orc_npairs_prim(void *h) { return (long)((Oracle *)h)->pp.g.size(); }

void orc_schwarz(void *h, double *out) {
  Oracle *o = (Oracle *)h;
  ints_exchange(*o);
  if (out) std::memcpy(out, o->schwarz.data(), o->schwarz.size() * sizeof(double));
}
void orc_set_schwarz(void *h, const double *in) {
  Oracle *o = (Oracle *)h;
  o->schwarz.assign(in, in + (size_t)o->b.nshell * o->b.nshell);
}

// int2_run_cam pass switch (int2.F90:538-584): mu > 0 selects attenuated integrals and the attenuated Schwarz matrix
// (computed on first use, :674-681); mu = 0 goes back to the regular pass.
void orc_set_attenuation(void *h, double mu) {
  Oracle *o = (Oracle *)h;
  if (o->schwarz_regular.empty() && o->mu == 0.0) o->schwarz_regular = o->schwarz;
  if (mu > 0) {
    if (o->schwarz_att.empty() || o->mu != mu) {
      if (o->mu == 0.0) o->schwarz_regular = o->schwarz;
      ints_exchange(*o, mu);  // fills o->schwarz
      o->schwarz_att = o->schwarz;
    }
    o->schwarz = o->schwarz_att;
    o->mu = mu;
  } else {
    if (o->mu > 0) o->schwarz = o->schwarz_regular;
    o->mu = 0.0;
  }
}

void orc_get_schwarz(void *h, double *out) {  // the matrix of the active pass
  Oracle *o = (Oracle *)h;
  std::memcpy(out, o->schwarz.data(), o->schwarz.size() * sizeof(double));
}

void orc_rys(int nroots, double x, double *u, double *w) { rys_general(x, nroots, u, w); }
// table-driven roots for the timed CPU baseline (see rys_fast); 0 = the restated general algorithm (default, parity)
void orc_set_fast_rys(int on) { g_fast_rys = on; }

// one shell quartet (0-based shells); out(l,k,j,i) in ORIGINAL shell order i,j,k,l (l fastest), nout[4]
int orc_eri_block(void *h, int i, int j, int k, int l, double *out, int *nout) {
  Oracle *o = (Oracle *)h;
  Eri g;
  g.init(*std::max_element(o->b.am.begin(), o->b.am.end()), o->cut);
  g.mu2_1 = o->mu > 0 ? 1.0 / (o->mu * o->mu) : 0.0;
  int ids[4] = {i, j, k, l};
  set_ids(g, o->b, ids);
  bool zero = rys_compute(g, o->b, o->pp);
  int n[4];
  for (int s = 0; s < 4; s++) n[g.flips[s]] = g.nbf[s];
  for (int s = 0; s < 4; s++) nout[s] = n[s];
  size_t tot = (size_t)n[0] * n[1] * n[2] * n[3];
  if (zero) { std::fill(out, out + tot, 0.0); return 1; }
  int c[4];
  for (c[0] = 0; c[0] < g.nbf[0]; c[0]++)
    for (c[1] = 0; c[1] < g.nbf[1]; c[1]++)
      for (c[2] = 0; c[2] < g.nbf[2]; c[2]++)
        for (c[3] = 0; c[3] < g.nbf[3]; c[3]++) {
          int a[4];
          for (int s = 0; s < 4; s++) a[g.flips[s]] = c[s];
          out[((a[0] * n[1] + a[1]) * n[2] + a[2]) * n[3] + a[3]] =
              g.ints[((c[0] * g.nbf[1] + c[1]) * g.nbf[2] + c[2]) * g.nbf[3] + c[3]];
        }
  return 0;
}

// kind: 0 RHF, 1 UROHF (packed d,f), 2 TD (d2 -> apb=f, amb=f2), 3 MRSF (d3 -> f3=f), 4 COLLECT
// flags bit0 int_apb, bit1 int_amb, bit2 tamm_dancoff, bit3 tamm_dancoff_coulomb
// stats[0..2] = nschwz, surviving shell quartets, AO integrals stored.  Raw accumulators are returned
// (no 0.5/diag scaling, no apb symmetrisation): see orc_fock_post.
void orc_run(void *h, int kind, const double *d, int nfocks, int ncomp, double se, double sc, int flags, double *f,
             double *f2, int nthreads, long pair_lo, long pair_hi, int stride, int offset, long *stats) {
  Oracle *o = (Oracle *)h;
  Consumer c;
  c.kind = kind; c.nbf = o->b.nbf; c.nfocks = nfocks; c.se = se; c.sc = sc; c.d = d; c.ncomp = ncomp;
  c.int_apb = flags & 1; c.int_amb = (flags >> 1) & 1; c.tda = (flags >> 2) & 1; c.tda_coulomb = (flags >> 3) & 1;
  c.cur_pass = (flags >> 4) & 1 ? 2 : 1;
  size_t nbf = c.nbf, ntri = nbf * (nbf + 1) / 2, fs = 0, f2s = 0;
  if (kind == RHF || kind == UROHF) fs = ntri * nfocks;
  if (kind == TD || kind == TDGRD) { fs = nbf * nbf * nfocks; f2s = fs; }
  if (kind == UMRSF) fs = nbf * nbf * nfocks * ncomp;
  if (kind == MRSF) {
    fs = nbf * nbf * nfocks * ncomp;
    c.ds.assign(nbf * nbf * nfocks * 4, 0.0);  // tdhf_mrsf_lib.F90:86-92
    for (size_t nu = 0; nu < nbf; nu++)
      for (size_t mu = 0; mu < nbf; mu++)
        for (int cc = 0; cc < 4; cc++)
          for (int v = 0; v < nfocks; v++)
            c.ds[v + (size_t)nfocks * (cc + 4 * (mu + nbf * nu))] =
                d[v + (size_t)nfocks * (cc + (size_t)ncomp * (mu + nbf * nu))] + d[v + (size_t)nfocks * (cc + (size_t)ncomp * (nu + nbf * mu))];
  }
  std::fill(f, f + fs, 0.0);
  if (f2s) std::fill(f2, f2 + f2s, 0.0);
  RunStats st;
  twoei(*o, c, f, f2, fs, f2s, nthreads, pair_lo, pair_hi, stride, offset, st, nullptr, nullptr, nullptr);
  if (stats) { stats[0] = st.nschwz; stats[1] = st.nshq; stats[2] = st.nint; }
}

// int2_rpagrd_data_t (tdhf_lib.F90:42-57, 1068-1320): H+[X+Y] -> hpp, H+[T] -> hpt (both symmetrised at stop, :1138-1141),
// H-[X-Y] -> hmm; arrays (nbf, nbf, nspin, n) column-major
void orc_run_rpagrd(void *h, int nspin, int np, int nm, int nt, const double *xpy, const double *xmy, const double *t,
                    double se, double sc, double *hpp, double *hpt, double *hmm, int nthreads, long *stats) {
  Oracle *o = (Oracle *)h;
  Consumer c;
  c.kind = RPAGRD; c.nbf = o->b.nbf; c.se = se; c.sc = sc;
  c.nspin = nspin; c.np = np; c.nm = nm; c.nt = nt; c.xpy = xpy; c.xmy = xmy; c.tt = t;
  const size_t n2 = (size_t)c.nbf * c.nbf, slab = n2 * nspin;
  c.rp_off[0] = 0; c.rp_off[1] = slab * np; c.rp_off[2] = slab * (np + nt);
  std::vector<double> f(slab * (np + nt + nm), 0.0);
  RunStats st;
  twoei(*o, c, f.data(), nullptr, f.size(), 0, nthreads, 0, -1, 1, 0, st, nullptr, nullptr, nullptr);
  for (size_t q = 0; q < slab * np; q++) hpp[q] = f[q];
  for (size_t q = 0; q < slab * nt; q++) hpt[q] = f[c.rp_off[1] + q];
  for (size_t q = 0; q < slab * nm; q++) hmm[q] = f[c.rp_off[2] + q];
  auto symm = [&](double *a, size_t cnt) {  // symmetrize_matrices
    for (size_t m = 0; m < cnt; m++) {
      double *x = a + m * n2;
      for (int i = 0; i < c.nbf; i++)
        for (int j = 0; j <= i; j++) { double s2 = x[i + (size_t)c.nbf * j] + x[j + (size_t)c.nbf * i]; x[i + (size_t)c.nbf * j] = x[j + (size_t)c.nbf * i] = s2; }
    }
  };
  symm(hpp, (size_t)np * nspin);
  symm(hpt, (size_t)nt * nspin);
  if (stats) { stats[0] = st.nschwz; stats[1] = st.nshq; stats[2] = st.nint; }
}

// fock_jk post-processing scf_addons.F90:1177-1185: f = 0.5 f, diagonal x2 (packed, per fock)
void orc_fock_post(int nbf, int nfocks, double *f) {
  size_t ntri = (size_t)nbf * (nbf + 1) / 2;
  for (int m = 0; m < nfocks; m++) {
    double *F = f + m * ntri;
    for (size_t q = 0; q < ntri; q++) F[q] *= 0.5;
    for (int i = 0; i < nbf; i++) F[tri(i, i)] *= 2.0;
  }
}
// symmetrize_matrix used by int2_td_data_t_parallel_stop (tdhf_lib.F90:107-109): a <- a + a^T
void orc_td_post(int nbf, int nvec, double *apb) {
  for (int v = 0; v < nvec; v++) {
    double *a = apb + (size_t)v * nbf * nbf;
    for (int i = 0; i < nbf; i++)
      for (int j = 0; j <= i; j++) { double s = a[i + (size_t)nbf * j] + a[j + (size_t)nbf * i]; a[i + (size_t)nbf * j] = a[j + (size_t)nbf * i] = s; }
  }
}

// surviving canonical shell-quartet list (0-based i,j,k,l) in loop order for a packed RHF density
long orc_quartet_list(void *h, const double *d, int nfocks, int *out, long maxq, long *nschwz) {
  Oracle *o = (Oracle *)h;
  Consumer c;
  c.kind = RHF; c.nbf = o->b.nbf; c.nfocks = nfocks; c.d = d;
  init_screen(o->b, c);
  const int ns = o->b.nshell;
  const double *Q = o->schwarz.data(), *dsh = c.dsh.data();
  long n = 0, skipped = 0;
  for (int i = 0; i < ns; i++)
    for (int j = 0; j <= i; j++) {
      if (Q[(size_t)i * ns + j] * c.max_den < o->cut.integral) { skipped += (long)(i + 1) * i / 2 + (j + 1); continue; }
      for (int k = 0; k <= i; k++) {
        int jork = (i == k) ? j : k;
        for (int l = 0; l <= jork; l++) {
          double res = Q[(size_t)i * ns + j] * Q[(size_t)k * ns + l];
          double m = std::max({4 * dsh[(size_t)i * ns + j], 4 * dsh[(size_t)k * ns + l], dsh[(size_t)j * ns + l],
                               dsh[(size_t)j * ns + k], dsh[(size_t)i * ns + l], dsh[(size_t)i * ns + k]});
          res = res * m;
          if (res < o->cut.integral) { skipped++; continue; }
          if (out && n < maxq) { out[4 * n] = i; out[4 * n + 1] = j; out[4 * n + 2] = k; out[4 * n + 3] = l; }
          n++;
        }
      }
    }
  if (nschwz) *nschwz = skipped;
  return n;
}

void orc_shlden(void *h, int kind, const double *d, int nfocks, int ncomp, double *dsh, double *max_den) {
  Oracle *o = (Oracle *)h;
  Consumer c;
  c.kind = kind; c.nbf = o->b.nbf; c.nfocks = nfocks; c.ncomp = ncomp; c.d = d;
  init_screen(o->b, c);
  std::memcpy(dsh, c.dsh.data(), c.dsh.size() * sizeof(double));
  *max_den = c.max_den;
}

// dense ERI tensor (ab|cd), nbf^4 doubles, from the storeints stream with the halving undone
// (semantics of modules/int2e.F90:181-194). Small systems only.
void orc_dense_eri(void *h, double *eri) {
  Oracle *o = (Oracle *)h;
  Consumer c;
  c.kind = COLLECT; c.nbf = o->b.nbf;
  size_t n = c.nbf;
  std::fill(eri, eri + n * n * n * n, 0.0);
  std::vector<int16_t> ids;
  std::vector<double> vals;
  RunStats st;
  std::vector<double> saveQ = o->schwarz;
  Cutoffs savec = o->cut;
  o->schwarz.assign((size_t)o->b.nshell * o->b.nshell, 1.0e10);  // no Schwarz skipping
  o->cut.integral = 0.0;                                         // no element cutoff
  twoei(*o, c, nullptr, nullptr, 0, 0, 1, 0, -1, 1, 0, st, nullptr, &ids, &vals);
  o->schwarz = saveQ;
  o->cut = savec;
  for (size_t q = 0; q < vals.size(); q++) {
    int i = ids[4 * q] - 1, j = ids[4 * q + 1] - 1, k = ids[4 * q + 2] - 1, l = ids[4 * q + 3] - 1;
    double v = vals[q];
    if (i == j) v *= 2;
    if (k == l) v *= 2;
    if (i == k && j == l) v *= 2;
    int p[8][4] = {{i, j, k, l}, {j, i, k, l}, {i, j, l, k}, {j, i, l, k}, {k, l, i, j}, {l, k, i, j}, {k, l, j, i}, {l, k, j, i}};
    for (auto &t : p) eri[((t[0] * n + t[1]) * n + t[2]) * n + t[3]] = v;
  }
}

void orc_int1e(void *h, int natom, const double *Z, const double *xyz, double *S, double *T, double *V) {
  int1e(((Oracle *)h)->b, natom, Z, xyz, S, T, V);
}

// cost-sorted bra shell-pair order of int2_build_shell_pair_map (int2.F90:864-921): pi[p] >= pj[p], 0-based
long orc_pair_order(void *h, int *pi_out, int *pj_out) {
  Oracle *o = (Oracle *)h;
  std::vector<int> pi, pj;
  build_pair_map(*o, pi, pj);
  if (pi_out) for (size_t k = 0; k < pi.size(); k++) { pi_out[k] = pi[k]; pj_out[k] = pj[k]; }
  return (long)pi.size();
}

int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
