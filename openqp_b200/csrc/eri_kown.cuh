// Ket-owner group kernel of the ERI family; included by eri_kernel.cuh (inside namespace oqpb)
// ---------------------------------------------------------------------------------------------------
// A quartet is owned by G consecutive lanes of one warp; G = ceil(NKET / KPL) is NOT restricted to powers of two
// (1, 2, 3, 5, 6, 9, 10 ... lanes; 32/G quartets per warp, the remaining lanes idle).  Lane t owns the ket Cartesian
// components t, t + G, ... (KPL of them) with ALL NA*NB bra components in registers, so that
//   * a lane reads whole rows [bra index] of a (root, direction) table [ket index][bra index] with LDS.128 and the lanes
//     of a group that share the ket index of a direction read the same words in the same instruction (broadcast);
//   * a lane carries KPL*NA*NB accumulators (<= OQPB_KOWN_ACC) instead of the whole block;
//   * the bra indices are normalised / projected in registers before anything goes to shared memory;
//   * the SYM digestion runs from REGISTERS: a lane holds V[all a,b][its (c,d)], loads each density element it needs
//     once, finishes J_cd alone and leaves per-(c,d) partial sums of J_ab, K_ac, K_ad, K_bc, K_bd in shared memory; the
//     outputs are then summed over the ket components by the lanes of the group (one red per Fock element as before).
//     The bra-owner group kernel re-reads the block from shared memory six times and loads a density element per FMA.
// Phases per primitive quartet as in eri_group_kernel (roots, (root, direction) recurrences in registers, assembly).
#ifndef OQPB_KOWN_ACC
#define OQPB_KOWN_ACC 60
#endif
#ifndef OQPB_KOWN_REGS
#define OQPB_KOWN_REGS 128
#endif
// per-class override of the ket components per lane (0 = heuristic)
// Measured on (H2O)32 (gpurun_out/s8, s21): one ket component per lane wins wherever the lane then holds <= 40 accumulators;
// the (fp| bra (30 components) takes two ket components per lane and the (fd| bra (60) one, both WITHOUT a register cap
// (168 registers cost (fd|ps) 40 %: 600-1600 bytes of spills).
template <int LA, int LB, int LC, int LD>
__host__ __device__ constexpr int kown_kpl_override() {
  // |dp) kets (18 components) on the (dp| and (dd| bras: two per lane = 9 lanes x 3 quartets per warp instead of 18 lanes x 1
  // ((dp|dp) 48.5 -> 44.2 ms, (dd|dp) 24.6 -> 21.4 ms)
  if (LC == 2 && LD == 1 && ((LA == 2 && LB == 1) || (LA == 2 && LB == 2))) return 2;
  return (LA == 3 && LB == 1) ? 0 : 1;  // (fp| bra: the heuristic below (two where they divide evenly); everything else: one
}

template <int LA, int LB, int LC, int LD, int PV>
struct KownCfg {
  using C = ClassCfg<LA, LB, LC, LD>;
  static constexpr int NBRA = C::NA * C::NB, NKET = C::NKET, R = C::R, NKL1 = C::NKL1, NIJ1 = C::NIJ1;
  static constexpr int N0 = Shell<LA, PV>::NOUT, N1 = Shell<LB, PV>::NOUT, N2 = Shell<LC, PV>::NOUT, N3 = Shell<LD, PV>::NOUT;
  static constexpr int N01 = N0 * N1, NKETP = N2 * N3;
  // ket components per lane: as many as the accumulator budget allows (fewer lanes per quartet = less replicated
  // per-quartet work), as long as the components divide evenly enough over the lanes (>= 80 % of the lane-passes useful)
  static constexpr int pick_kpl() {
    if (kown_kpl_override<LA, LB, LC, LD>() > 0) return kown_kpl_override<LA, LB, LC, LD>();
    int best = 1;
    for (int kpl = 1; kpl <= NKET && (kpl == 1 || kpl * NBRA <= OQPB_KOWN_ACC); ++kpl) {
      const int g = (NKET + kpl - 1) / kpl;
      if (g > 32) continue;
      const int lanes = (32 / g) * g;
      if (NKET * 100 >= 80 * g * kpl && lanes >= 24) best = kpl;
    }
    return best;
  }
  static constexpr int KPL = pick_kpl();
  static constexpr int G = (NKET + KPL - 1) / KPL;
  static constexpr int QPW = G <= 32 ? 32 / G : 1;
  static constexpr bool OK = C::NCART4 > SMALL_MAX && G <= 32 && KPL * NBRA <= (OQPB_KOWN_ACC > 72 ? OQPB_KOWN_ACC : 72);
  // row stride of a table (doubles): even (LDS.128) and such that up to 8 rows start in different 16-byte banks
  static constexpr int row_stride() {
    for (int p = (NIJ1 + 1) & ~1;; p += 2) {
      const int u = p / 2, rows = NKL1 < 8 ? NKL1 : 8;
      bool ok = true;
      for (int a = 0; a < rows && ok; ++a)
        for (int b = a + 1; b < rows && ok; ++b)
          if ((a * u) % 8 == (b * u) % 8) ok = false;
      if (ok) return p;
    }
  }
  static constexpr int ROW = row_stride();
  static constexpr int GSTR0 = ROW * NKL1;
  static constexpr int GSTR = (GSTR0 / 2) % 2 == 1 ? GSTR0 : GSTR0 + 2;  // doubles per (root, direction) table; GSTR/2 odd:
                                                                         // the task lanes store to different 16-byte banks
  static constexpr int GREG = 3 * R * GSTR;
  static constexpr bool KPROJ = LC >= 2 || LD >= 2;  // ket indices need the shared-memory projection passes
  static constexpr int BLK0 = N01 * NKET;            // block after the bra projection (ket still Cartesian)
  static constexpr int BLKSZ = KPROJ ? 2 * BLK0 : BLK0;
  // partial sums of the register digestion: J_ab per lane, K_ac / K_ad / K_bc / K_bd per projected ket component
  static constexpr int PAB = 0, PAC = PAB + N01 * G, PAD = PAC + N0 * NKETP, PBC = PAD + N0 * NKETP, PBD = PBC + N1 * NKETP;
  static constexpr int REDSZ = PBD + N1 * NKETP;
  static constexpr int RW = 2 * R + 12 + (2 * R) % 2;  // roots / weights + quartet geometry (A, A-B, C, C-D)
  static constexpr int max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
  static constexpr int RWOFF = (max3(GREG, BLKSZ, REDSZ) + 1) & ~1;
  static constexpr int QSM0 = RWOFF + RW;
  static constexpr int QSM = (QSM0 / 2) % 2 == 1 ? QSM0 : QSM0 + 2;  // per-quartet stride (doubles), QSM/2 odd
  static constexpr int QBYTES = QSM * 8 + 32;
#ifndef OQPB_KOWN_WPC
#define OQPB_KOWN_WPC 4
#endif
  static constexpr int WPC = (OQPB_KOWN_WPC >= 4 && 4 * QPW * QBYTES <= 48 * 1024) ? 4 : ((OQPB_KOWN_WPC >= 2 && 2 * QPW * QBYTES <= 64 * 1024) ? 2 : 1);
  static constexpr int NT = 32 * WPC;
  static constexpr size_t SMEM = (size_t)WPC * QPW * QBYTES;
  // register cap requested from ptxas
#ifndef OQPB_KOWN_REGS_MID
#define OQPB_KOWN_REGS_MID 255
#endif
#ifndef OQPB_KOWN_REGS_36
#define OQPB_KOWN_REGS_36 168
#endif
  // register cap by accumulators per lane (measured per class on (H2O)32): <= 20 -> 128 ((dp| bra: 168 is 9 % slower),
  // 21..40 -> 168 ((dd| bra: (dd|ds) 32.7 -> 22.0 ms, (dd|pp) 23.3 -> 15.9 ms against 128), above -> no cap
  static constexpr int REGCAP = KPL * NBRA <= 20 ? OQPB_KOWN_REGS : (KPL * NBRA <= 40 ? OQPB_KOWN_REGS_36 : (KPL * NBRA <= 60 ? OQPB_KOWN_REGS_MID : 255));
};

// segmented sums over the quartet slots of a warp for any group size (QPW not a power of two)
template <int G>
struct KSeg {
  static constexpr int QPW = 32 / G;
  static constexpr int NST = QPW <= 1 ? 0 : (QPW <= 2 ? 1 : (QPW <= 4 ? 2 : (QPW <= 8 ? 3 : (QPW <= 16 ? 4 : 5))));
  bool up[NST > 0 ? NST : 1];
  bool head;
};
template <int G>
__device__ __forceinline__ KSeg<G> kseg_make(long long key, int lane) {
  KSeg<G> m;
  m.head = true;
  if constexpr (KSeg<G>::NST > 0) {
    const int g = lane / G;
    const long long prev = __shfl_up_sync(0xffffffffu, key, G);
    m.head = g == 0 || prev != key;
    const unsigned heads = __ballot_sync(0xffffffffu, m.head && (lane % G) == 0 && g < KSeg<G>::QPW);
    const int rid = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int k = 0; k < KSeg<G>::NST; ++k) {
      const int o = __shfl_down_sync(0xffffffffu, rid, G << k);
      m.up[k] = (g + (1 << k) < KSeg<G>::QPW) && o == rid;
    }
  }
  return m;
}
template <int G>
__device__ __forceinline__ double kseg_sum(double v, const KSeg<G>& m) {
  if constexpr (KSeg<G>::NST > 0) {
#pragma unroll
    for (int k = 0; k < KSeg<G>::NST; ++k) {
      const double o = __shfl_down_sync(0xffffffffu, v, G << k);
      if (m.up[k]) v += o;
    }
  }
  return v;
}

// quartet descriptor of a group, 32 bytes in shared memory (read by the MODE_GEN digestion of the whole warp)
struct KQ {
  int oa, ob, oc, od;
  int work, pad0, pad1, pad2;
};

template <int LA, int LB, int LC, int LD, int PV>
__global__ void __launch_bounds__(KownCfg<LA, LB, LC, LD, PV>::NT, min_ctas(KownCfg<LA, LB, LC, LD, PV>::NT, KownCfg<LA, LB, LC, LD, PV>::REGCAP))
eri_kown_kernel(const EriArgs A) {
  using Cfg = ClassCfg<LA, LB, LC, LD>;
  using KC = KownCfg<LA, LB, LC, LD, PV>;
  constexpr int N0 = KC::N0, N1 = KC::N1, N2 = KC::N2, N3 = KC::N3, NTOT = N0 * N1 * N2 * N3, N01 = KC::N01, NKETP = KC::NKETP;
  constexpr int R = Cfg::R, NA = Cfg::NA, NB = Cfg::NB, NC = Cfg::NC, ND = Cfg::ND, NKET = Cfg::NKET, NBRA = KC::NBRA;
  constexpr int NMAX = Cfg::NMAX, MMAX = Cfg::MMAX, NKL1 = Cfg::NKL1, NIJ1 = Cfg::NIJ1;
  constexpr int G = KC::G, KPL = KC::KPL, QPW = KC::QPW, ROW = KC::ROW, GSTR = KC::GSTR, QSM = KC::QSM;
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane / G, t = lane % G;
  const bool lane_ok = g < QPW;  // lanes past the last whole group idle (they still take part in the warp collectives)
  const int qslot = w * QPW + (lane_ok ? g : 0);
  double* qs = smem + (size_t)qslot * QSM;
  double* rw = qs + KC::RWOFF;
  double* geo = rw + 2 * R + (2 * R) % 2;  // A (3), A-B (3), C (3), C-D (3)
  char* tail = reinterpret_cast<char*>(smem + (size_t)KC::WPC * QPW * QSM);
  KQ& kq = *reinterpret_cast<KQ*>(tail + (size_t)qslot * 32);

  // ket components of this lane: kc = t + G j; row offsets of its x / y / z ket indices inside a table
  int kc[KPL], rx[KPL], ry[KPL], rz[KPL];
  bool kok[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int k = t + G * j;
    kok[j] = k < NKET;
    kc[j] = kok[j] ? k : 0;
    int cx, cy, cz, dx, dy, dz;
    cart_xyz_rt(LC, kc[j] / ND, cx, cy, cz);
    cart_xyz_rt(LD, kc[j] % ND, dx, dy, dz);
    rx[j] = (cx * (LD + 1) + dx) * ROW; ry[j] = (cy * (LD + 1) + dy) * ROW; rz[j] = (cz * (LD + 1) + dz) * ROW;
  }
  const unsigned ntasks = A.task_cap ? min(*A.ntasks, A.task_cap) : *A.ntasks;
  unsigned long long st_prim = 0, st_ints = 0;

  // static warp-strided walk over the task list (a dynamic fetch counter was ~9 % of the stall samples: one atomic per
  // 32/G quartets on a single address); neighbouring warps take neighbouring quartets, which usually share the bra
  for (unsigned base = (blockIdx.x * KC::WPC + w) * QPW; base < ntasks; base += gridDim.x * KC::WPC * QPW) {
    // ---- every lane of a group reads the quartet's task and pair entries itself (same addresses: broadcast loads)
    const unsigned ti = base + (lane_ok ? g : 0);
    const bool valid = lane_ok && ti < ntasks;
    const int2 tk = A.tasks[valid ? ti : base];
    const PairEntry* __restrict__ pbp = A.bra + tk.x;
    const PairEntry* __restrict__ pkp = A.ket + tk.y;
    const int sa = pbp->sa, sb = pbp->sb, sc = pkp->sa, sd = pkp->sb;
    const int boff = pbp->poff, koff = pkp->poff, bcnt = valid ? pbp->pcnt : 0, kcnt = valid ? pkp->pcnt : 0;
    const int oa = pbp->oa, ob = pbp->ob, oc = pkp->oa, od = pkp->ob;
    float facf = 1.0f;
    if (sa == sb) facf *= 0.5f;
    if (sc == sd) facf *= 0.5f;
    if (sa == sc && sb == sd) facf *= 0.5f;
    // primitives are sorted by |K|/zeta: only the leading imax x jmax rectangle can pass the primitive-quartet test
    int imax = 0, jmax = 0;
    if (bcnt > 0 && kcnt > 0) {
      const double thr = A.prim_cutoff * (1.0 - 1e-9) * (pbp->zmin + pkp->zmin);
      const double* p0 = A.prim + (size_t)boff * PRIM_STRIDE;
      const double* q0 = A.prim + (size_t)koff * PRIM_STRIDE;
      const double da0 = __ldg(p0 + 4), db0 = __ldg(q0 + 4);
      while (imax < bcnt) { const double v = __ldg(p0 + (size_t)imax * PRIM_STRIDE + 4) * db0; if (v * v < thr) break; ++imax; }
      while (jmax < kcnt) { const double v = __ldg(q0 + (size_t)jmax * PRIM_STRIDE + 4) * da0; if (v * v < thr) break; ++jmax; }
    }
    const int ncand = imax * jmax;
    if (imax < 1) imax = 1;
    const int maxk = __reduce_max_sync(FULL, ncand);
    __syncwarp();  // previous pass fully done with this group's shared memory
    if (valid && t < 12) {
      // quartet geometry for the recurrence tasks: A, A-B, C, C-D
      const double* src = t < 6 ? &pbp->ax : &pkp->ax;
      geo[t] = src[t < 6 ? t : t - 6];
    }
    if constexpr (G < 12) {
      if (valid) for (int i = t + G; i < 12; i += G) { const double* src = i < 6 ? &pbp->ax : &pkp->ax; geo[i] = src[i < 6 ? i : i - 6]; }
    }
    double acc[KPL][NBRA];
#pragma unroll
    for (int j = 0; j < KPL; ++j)
#pragma unroll
      for (int e = 0; e < NBRA; ++e) acc[j][e] = 0.0;
    bool any = false;
    __syncwarp();

#pragma unroll 1
    for (int ip = 0; ip < maxk; ++ip) {
      // candidate ip of this group's rectangle (bra primitive fastest); the int_rys.F90:229-232 test decides
      bool keep = ip < ncand;
      double Px = 0, Py = 0, Pz = 0, zeta = 1, Kp = 0, Qx = 0, Qy = 0, Qz = 0, eta = 1, Kq = 0, zinv = 1, einv = 1;
      if (keep) {
        const double2* pp = reinterpret_cast<const double2*>(A.prim + (size_t)(boff + ip % imax) * PRIM_STRIDE);
        const double2* pq = reinterpret_cast<const double2*>(A.prim + (size_t)(koff + ip / imax) * PRIM_STRIDE);
        const double2 p01 = __ldg(pp), p23 = __ldg(pp + 1), p45 = __ldg(pp + 2);
        const double2 q01 = __ldg(pq), q23 = __ldg(pq + 1), q45 = __ldg(pq + 2);
        Px = p01.x; Py = p01.y; Pz = p23.x; zeta = p23.y; Kp = p45.x; zinv = p45.y;
        Qx = q01.x; Qy = q01.y; Qz = q23.x; eta = q23.y; Kq = q45.x; einv = q45.y;
        const double pf = Kp * Kq;
        keep = !(pf * pf < A.prim_cutoff * (zeta + eta + zeta * eta * A.mu2inv));
      }
      if (!__any_sync(FULL, keep)) continue;
      const double PQx = Px - Qx, PQy = Py - Qy, PQz = Pz - Qz;
      const double rsab = rsqrt_nr(zeta + eta + zeta * eta * A.mu2inv);
      const double abinv = rsab * rsab;
      const double rho = zeta * eta * abinv;
      // ---- B1: roots and weights, one root per lane
      if (keep) {
        const double X = rho * (PQx * PQx + PQy * PQy + PQz * PQz);
        const RysX sx = rys_prepare<R>(A, X);
        for (int r = t; r < R; r += G) {
          double t2, wt;
          rys_pair<R, false>(A, nullptr, sx, r, t2, wt);
          rw[r] = t2; rw[R + r] = wt;
        }
      }
      __syncwarp();
      // ---- B2: 2-D recurrences in registers, (root, direction) tasks over the G lanes; table [ket index][bra index]
      if (keep) {
        const double pref = Kp * Kq * rsab;
        for (int task = t; task < 3 * R; task += G) {
          const int r = task / 3, dir = task % 3;
          const double t2 = rw[r];
          const double Ad = geo[dir], ABd = geo[3 + dir], Cd = geo[6 + dir], CDd = geo[9 + dir];
          const double Pd = dir == 0 ? Px : (dir == 1 ? Py : Pz);
          const double Qd = dir == 0 ? Qx : (dir == 1 ? Qy : Qz);
          const double PQd = Pd - Qd;
          const double t2r = t2 * rho;
          const double c00 = (Pd - Ad) - t2r * zinv * PQd;
          const double d00 = (Qd - Cd) + t2r * einv * PQd;
          const double b10 = 0.5 * zinv * (1.0 - t2r * zinv);
          const double b01 = 0.5 * einv * (1.0 - t2r * einv);
          const double b00 = 0.5 * t2 * abinv;
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
            double col[ROW];
#pragma unroll
            for (int i = NIJ1; i < ROW; ++i) col[i] = 0.0;
#pragma unroll
            for (int a = 0; a <= LA; ++a) col[a * (LB + 1)] = h[a][k];
#pragma unroll
            for (int b = 1; b <= LB; ++b) {
#pragma unroll
              for (int n = 0; n < NMAX - b; ++n) h[n][k] = h[n + 1][k] + ABd * h[n][k];
#pragma unroll
              for (int a = 0; a <= LA; ++a) col[a * (LB + 1) + b] = h[a][k];
            }
            double2* dstrow = reinterpret_cast<double2*>(S3 + k * ROW);
#pragma unroll
            for (int p = 0; p < (NIJ1 + 1) / 2; ++p) dstrow[p] = make_double2(col[2 * p], col[2 * p + 1]);
          }
        }
      }
      __syncwarp();
      // ---- B3: assembly, this lane's ket components x all bra components
      if (keep) {
        any = true;
        if (t == 0) ++st_prim;
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
          const double* gbase = qs + (size_t)(3 * r) * GSTR;
#pragma unroll
          for (int j = 0; j < KPL; ++j) {
            const double2* gx = reinterpret_cast<const double2*>(gbase + rx[j]);
            const double2* gy = reinterpret_cast<const double2*>(gbase + GSTR + ry[j]);
            const double2* gz = reinterpret_cast<const double2*>(gbase + 2 * GSTR + rz[j]);
            constexpr int HL = (NIJ1 + 1) / 2;
            double2 X2[HL], Y2[HL], Z2[HL];
#pragma unroll
            for (int p = 0; p < HL; ++p) { X2[p] = gx[p]; Y2[p] = gy[p]; Z2[p] = gz[p]; }
            static_for<0, NBRA>([&](auto E) {
              constexpr int e = decltype(E)::value;
              constexpr int ia = e / NB, ib = e % NB;
              constexpr int ix = Cart<LA>::x(ia) * (LB + 1) + Cart<LB>::x(ib);
              constexpr int iy = Cart<LA>::y(ia) * (LB + 1) + Cart<LB>::y(ib);
              constexpr int iz = Cart<LA>::z(ia) * (LB + 1) + Cart<LB>::z(ib);
              const double xv = (ix & 1) ? X2[ix >> 1].y : X2[ix >> 1].x;
              const double yv = (iy & 1) ? Y2[iy >> 1].y : Y2[iy >> 1].x;
              const double zv = (iz & 1) ? Z2[iz >> 1].y : Z2[iz >> 1].x;
              acc[j][e] = fma(xv * yv, zv, acc[j][e]);
            });
          }
        }
      }
      __syncwarp();  // the tables are overwritten by the next primitive's B2, rw by its B1
    }
    // ---- bra indices normalised / projected in registers
    const bool work = valid && any;  // `any` is uniform over the group
    const double fac = (double)facf, cut = A.cutoff;
    unsigned nz = 0;
    double vv[KPL][N01];
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      double t1[NA * N1];
      proj_reg<LB, Shell<LB, PV>::PURE, NA, 1>(acc[j], t1);
      proj_reg<LA, Shell<LA, PV>::PURE, 1, N1>(t1, vv[j]);
    }
    const bool sym = A.mode == MODE_SYM;
    bool kpok[KPL];  // this lane's j-th (projected) ket component exists
    int kp[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) { kp[j] = kc[j]; kpok[j] = kok[j]; }
    if (KC::KPROJ || !sym) {
      // block [a][b][c][d] (ket Cartesian) to shared memory (aliases the tables: every lane is past its last assembly read)
      if (work) {
#pragma unroll
        for (int j = 0; j < KPL; ++j)
          if (kok[j]) {
#pragma unroll
            for (int e = 0; e < N01; ++e) qs[e * NKET + kc[j]] = vv[j][e];
          }
      }
      __syncwarp();
      if (valid && !any) {
        if (A.mode == MODE_SCHWARZ && t == 0) A.qout[tk.x] = 0.0;
        if (A.mode == MODE_BLOCK)
          for (int e = t; e < NTOT; e += G) A.blockout[e] = 0.0;
      }
      double* src = qs;
      double* dst = qs + KC::BLK0;
      if (LD >= 2) {
        if (work) proj_smem<LD, Shell<LD, PV>::PURE, N01 * NC, 1>(src, dst, t, G);
        double* tmp = src; src = dst; dst = tmp;
        __syncwarp();
      }
      if (LC >= 2) {
        if (work) proj_smem<LC, Shell<LC, PV>::PURE, N01, N3>(src, dst, t, G);
        double* tmp = src; src = dst; dst = tmp;
        __syncwarp();
      }
      if (A.mode == MODE_SCHWARZ) {
        double mx = 0.0;
        if (work) for (int e = t; e < NTOT; e += G) mx = fmax(mx, fabs(src[e]));
        // group maximum in lane t == 0 (G is not a power of two: plain gather)
        double gm = mx;
        for (int k = 1; k < G; ++k) gm = fmax(gm, __shfl_sync(FULL, mx, (lane_ok ? g * G : 0) + k));
        if (work && t == 0) A.qout[tk.x] = sqrt(gm);
        continue;
      }
      if (A.mode == MODE_BLOCK) {
        if (work) for (int e = t; e < NTOT; e += G) A.blockout[e] = src[e];
        continue;
      }
      if (!sym) {
        // MODE_GEN: element cutoff / coincidence factor in shared memory, then the quartets of the warp one after the
        // other, each digested by all 32 lanes (DMMA)
        if (work) {
          for (int e = t; e < NTOT; e += G) {
            const double v = src[e];
            const bool z = fabs(v) < cut;
            nz += !z;
            src[e] = z ? 0.0 : v * fac;
          }
          st_ints += (unsigned long long)nz * (unsigned)(8.0f * facf);
          if (t == 0) { kq.oa = oa; kq.ob = ob; kq.oc = oc; kq.od = od; }
        }
        if (lane_ok && t == 0) kq.work = work ? 1 : 0;
        __syncwarp();
        const int srcoff = (int)(src - qs);
#pragma unroll 1
        for (int gq = 0; gq < QPW; ++gq) {
          const KQ& qq = *reinterpret_cast<const KQ*>(tail + (size_t)(w * QPW + gq) * 32);
          if (!qq.work) continue;
          const int off[4] = {qq.oa, qq.ob, qq.oc, qq.od};
          digest_gen_warp<N0, N1, N2, N3>(A, smem + (size_t)(w * QPW + gq) * QSM + srcoff, off, lane);
        }
        __syncwarp();
        continue;
      }
      // SYM with a ket projection: the lane takes the projected ket components t, t + G, ... back into registers
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int k = t + G * j;
        kpok[j] = k < NKETP;
        kp[j] = kpok[j] ? k : 0;
        if (work && kpok[j]) {
#pragma unroll
          for (int e = 0; e < N01; ++e) vv[j][e] = src[e * NKETP + kp[j]];
        }
      }
      __syncwarp();  // the partial sums below overwrite the block
    }
    // ---- element cutoff (int2.F90:1806-1812) and shell-level coincidence factor (int2.F90:1849-1851), in registers
#pragma unroll
    for (int j = 0; j < KPL; ++j)
#pragma unroll
      for (int e = 0; e < N01; ++e) {
        const double v = vv[j][e];
        const bool z = fabs(v) < cut || !(work && kpok[j]);
        nz += !z;
        vv[j][e] = z ? 0.0 : v * fac;
      }
    if (work) st_ints += (unsigned long long)nz * (unsigned)(8.0f * facf);
    // ---- SYM digestion from registers (the reference's six packed updates, int2.F90:1414-1578)
    {
      const unsigned nbf = (unsigned)A.nbf;
      const double c4 = 4.0 * A.cj, c1 = A.ck;
      // quartets of the warp with the same bra / the same (bra, ket shell c) are summed before the red
      const long long kbra = valid ? (long long)tk.x : -1 - (long long)lane;
      const KSeg<G> mbra = kseg_make<G>(kbra, lane);
      const KSeg<G> mc = kseg_make<G>(valid ? ((long long)tk.x << 20) | (long long)sc : kbra, lane);
      for (int m = 0; m < A.nmat; ++m) {
        const double* __restrict__ DJ = A.DJ[m];
        const double* __restrict__ DK = A.DK[m];
        double* __restrict__ F = A.F[m];
        if (m > 0) __syncwarp();  // partial sums of the previous matrix consumed
        if (work) {
          double dab[N01];
#pragma unroll
          for (int a = 0; a < N0; ++a)
#pragma unroll
            for (int b = 0; b < N1; ++b) dab[a * N1 + b] = __ldg(DJ + ((unsigned)(oa + a) * nbf + (unsigned)(ob + b)));
          double dcd[KPL];
#pragma unroll
          for (int j = 0; j < KPL; ++j) {
            dcd[j] = 0.0;
            if (!kpok[j]) continue;
            const int c = kp[j] / N3, d = kp[j] % N3;
            const unsigned rc = (unsigned)(oc + c), rd = (unsigned)(od + d);
            dcd[j] = __ldg(DJ + (rc * nbf + rd));
            double dbd[N1], dbc[N1], dad[N0], dac[N0];
#pragma unroll
            for (int b = 0; b < N1; ++b) {
              dbd[b] = __ldg(DK + ((unsigned)(ob + b) * nbf + rd));
              dbc[b] = __ldg(DK + ((unsigned)(ob + b) * nbf + rc));
            }
#pragma unroll
            for (int a = 0; a < N0; ++a) {
              dad[a] = __ldg(DK + ((unsigned)(oa + a) * nbf + rd));
              dac[a] = __ldg(DK + ((unsigned)(oa + a) * nbf + rc));
            }
            double jcd = 0.0;
            double pbc[N1], pbd[N1];
#pragma unroll
            for (int b = 0; b < N1; ++b) pbc[b] = pbd[b] = 0.0;
#pragma unroll
            for (int a = 0; a < N0; ++a) {
              double pac = 0.0, pad = 0.0;
#pragma unroll
              for (int b = 0; b < N1; ++b) {
                const double v = vv[j][a * N1 + b];
                jcd = fma(v, dab[a * N1 + b], jcd);
                pac = fma(v, dbd[b], pac);
                pad = fma(v, dbc[b], pad);
                pbc[b] = fma(v, dad[a], pbc[b]);
                pbd[b] = fma(v, dac[a], pbd[b]);
              }
              qs[KC::PAC + a * NKETP + kp[j]] = pac;
              qs[KC::PAD + a * NKETP + kp[j]] = pad;
            }
#pragma unroll
            for (int b = 0; b < N1; ++b) {
              qs[KC::PBC + b * NKETP + kp[j]] = pbc[b];
              qs[KC::PBD + b * NKETP + kp[j]] = pbd[b];
            }
            if (jcd != 0.0) atomicAdd(F + tri_u(rc, rd), c4 * jcd);  // J_cd += 4 cj sum_ab v D_ab
          }
#pragma unroll
          for (int e = 0; e < N01; ++e) {  // J_ab partial of this lane: sum over its ket components
            double pab = 0.0;
#pragma unroll
            for (int j = 0; j < KPL; ++j) pab = fma(vv[j][e], dcd[j], pab);
            qs[KC::PAB + e * G + t] = pab;
          }
        }
        __syncwarp();
        // J_ab += 4 cj sum_cd v D_cd
        for (int ob_ = 0; ob_ < N01; ob_ += G) {
          const int o = ob_ + t;
          const bool act = work && o < N01;
          double sum = 0.0;
          if (act) {
            const double* p = qs + KC::PAB + o * G;
#pragma unroll
            for (int l = 0; l < G; ++l) sum += p[l];
          }
          sum = kseg_sum<G>(sum, mbra);
          if (act && mbra.head && sum != 0.0) atomicAdd(F + tri_u(oa + o / N1, ob + o % N1), c4 * sum);
        }
        // K_ac -= ck sum_bd v D_bd
        for (int ob_ = 0; ob_ < N0 * N2; ob_ += G) {
          const int o = ob_ + t;
          const bool act = work && o < N0 * N2;
          const int a = o / N2, c = o % N2;
          double sum = 0.0;
          if (act) {
            const double* p = qs + KC::PAC + a * NKETP + c * N3;
#pragma unroll
            for (int d = 0; d < N3; ++d) sum += p[d];
          }
          sum = kseg_sum<G>(sum, mc);
          if (act && mc.head && sum != 0.0) atomicAdd(F + tri_u(oa + a, oc + c), -c1 * sum);
        }
        // K_ad -= ck sum_bc v D_bc
        for (int ob_ = 0; ob_ < N0 * N3; ob_ += G) {
          const int o = ob_ + t;
          if (work && o < N0 * N3) {
            const int a = o / N3, d = o % N3;
            const double* p = qs + KC::PAD + a * NKETP + d;
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < N2; ++c) sum += p[c * N3];
            if (sum != 0.0) atomicAdd(F + tri_u(oa + a, od + d), -c1 * sum);
          }
        }
        // K_bc -= ck sum_ad v D_ad
        for (int ob_ = 0; ob_ < N1 * N2; ob_ += G) {
          const int o = ob_ + t;
          const bool act = work && o < N1 * N2;
          const int b = o / N2, c = o % N2;
          double sum = 0.0;
          if (act) {
            const double* p = qs + KC::PBC + b * NKETP + c * N3;
#pragma unroll
            for (int d = 0; d < N3; ++d) sum += p[d];
          }
          sum = kseg_sum<G>(sum, mc);
          if (act && mc.head && sum != 0.0) atomicAdd(F + tri_u(ob + b, oc + c), -c1 * sum);
        }
        // K_bd -= ck sum_ac v D_ac
        for (int ob_ = 0; ob_ < N1 * N3; ob_ += G) {
          const int o = ob_ + t;
          if (work && o < N1 * N3) {
            const int b = o / N3, d = o % N3;
            const double* p = qs + KC::PBD + b * NKETP + d;
            double sum = 0.0;
#pragma unroll
            for (int c = 0; c < N2; ++c) sum += p[c * N3];
            if (sum != 0.0) atomicAdd(F + tri_u(ob + b, od + d), -c1 * sum);
          }
        }
      }
    }
  }
  if (A.stat) {
    if (st_prim) atomicAdd(A.stat, st_prim);
    if (st_ints) atomicAdd(A.stat + 1, st_ints);
  }
}
