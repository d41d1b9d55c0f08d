// MRSF sigma session on the device; included by oqp_b200.cu (needs oqpb_ctx, DevBuf, mrsf_core, g_default_ctx).
//
// Replaces, for the new Davidson trial vectors of one iteration, the reference's whole triple
//   6a  iatogen + mrsfcbc      MO amplitudes -> seven AO densities             tdhf_lib.F90:480-498, tdhf_mrsf_lib.F90:940-1273
//   6b  int2_mrsf_data_t run   the J/K build (mrsf_core above)                  tdhf_mrsf_lib.F90:218-333
//   6c  mrsfmntoia + mrsfesum  AO Fock-like matrices -> MO amplitudes + Fock    tdhf_mrsf_lib.F90:1463-1735, 1918-2036
// behind the reference's own session ABI  routec_sig_init / _set_scale / _iter / _free  (source/modules/routec_sig.F90:28-56;
// caller and gate: modules/tdhf_mrsf_energy.F90:648-713).  Everything between the H2D copy of the trial amplitudes and the
// D2H copy of sigma stays in HBM: the seven densities of all vectors are written directly in the interleaved layout
// d3(v, c, mu, nu) the J/K kernels read, and the back-transformation reads f3 in place.
//
// 6a and 6c are chains of small dense products (MO coefficient panels x amplitude blocks, rank-1 updates); they run as ONE
// batched, arbitrarily strided FP64 GEMM kernel (batch = trial vector) so that a step of the reference (one dgemm per
// vector) is one launch for all vectors and no operand is ever re-packed.  They are < 2 % of the J/K build's time.
// The CPU restatement the tests check this against (mrsf_sigma.py of the test infrastructure) is pinned to the reference's
// XC-free CH2O MRSF golden.

namespace {

struct GemmOp {
  int M, N, K, batch;
  double alpha, beta;
  const double* alpha_b;  // optional per-batch factor (alpha_b[b * sab])
  long sab;
  const double* A; long sAi, sAk, sAb;
  const double* B; long sBk, sBj, sBb;
  double* C; long sCi, sCj, sCb;
};
// C[b](i,j) = alpha * alpha_b[b] * sum_k A[b](i,k) B[b](k,j) + beta * C[b](i,j); 16x16 tiles through shared memory
__global__ void __launch_bounds__(256) k_gemm_strided(const GemmOp g) {
  __shared__ double sa[16][17], sb[16][17];
  const int b = blockIdx.z, tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  const double* A = g.A + (long)b * g.sAb;
  const double* B = g.B + (long)b * g.sBb;
  double acc = 0.0;
  for (int k0 = 0; k0 < g.K; k0 += 16) {
    // sa[ty][tx] = A(i0 + ty, k0 + tx), sb[ty][tx] = B(k0 + ty, j0 + tx)
    const int ka = k0 + tx, kb = k0 + ty;
    sa[ty][tx] = (i < g.M && ka < g.K) ? A[(long)i * g.sAi + (long)ka * g.sAk] : 0.0;
    sb[ty][tx] = (kb < g.K && j < g.N) ? B[(long)kb * g.sBk + (long)j * g.sBj] : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fma(sa[ty][k], sb[k][tx], acc);
    __syncthreads();
  }
  if (i < g.M && j < g.N) {
    double* c = g.C + (long)b * g.sCb + (long)i * g.sCi + (long)j * g.sCj;
    const double al = g.alpha * (g.alpha_b ? g.alpha_b[(long)b * g.sab] : 1.0);
    *c = g.beta == 0.0 ? al * acc : al * acc + g.beta * *c;
  }
}

// special amplitudes after mrsfmntoia (tdhf_mrsf_lib.F90:1677-1685); W: (ntrial, nv), amplitude (i,j) at i + na (j - nb)
__global__ void k_sig_mntoia_fix(double* W, long ntrial, int nv, int na, int nb, int kind) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  double* w = W + (long)v * ntrial;
  const int lr1 = na - 2, lr2 = na - 1;
  auto at = [&](int i, int j) -> double& { return w[i + (long)na * (j - nb)]; };
  const double s11 = at(lr1, lr1), s22 = at(lr2, lr2);
  const double isq2 = 0.70710678118654752440;
  if (kind == 1) {
    at(lr1, lr1) = (s11 - s22) * isq2;
    at(lr2, lr2) = 0.0;
  } else {
    at(lr1, lr1) = (s11 + s22) * isq2;
    at(lr2, lr1) = 0.0; at(lr1, lr2) = 0.0; at(lr2, lr2) = 0.0;
  }
}
// Xs = X with the (O1,O1) and (O2,O2) amplitudes zeroed (mrsfesum :1946-1948)
__global__ void k_sig_scr(const double* X, double* Xs, long ntrial, int nv, int na, int nb) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ntrial * nv) return;
  const long a = e % ntrial;
  const int i = (int)(a % na), j = (int)(a / na) + nb;
  const int lr1 = na - 2, lr2 = na - 1;
  Xs[e] = ((i == lr1 && j == lr1) || (i == lr2 && j == lr2)) ? 0.0 : X[e];
}
// the O1 / O2 rows and columns and the special amplitudes of mrsfesum (:1962-2020); E holds tmp1 on entry.  One CTA per vector.
__global__ void k_sig_esum_fix(double* E, const double* X, const double* Xs, const double* fij, const double* fab, long ntrial,
                               int n, int na, int nb, int kind) {
  const int v = blockIdx.x;
  double* e = E + (long)v * ntrial;
  const double* x = X + (long)v * ntrial;
  const double* xs = Xs + (long)v * ntrial;
  const int lr1 = na - 2, lr2 = na - 1;
  const double isq2 = 0.70710678118654752440;
  const double s2 = kind == 3 ? 1.0 : -1.0;
  auto idx = [&](int i, int j) { return i + (long)na * (j - nb); };
  const double xlr = x[idx(lr1, lr1)];
  __shared__ double red[256];
  // dumn (uses scr = Xs): -fij(O1,:) scr(:,O1) - s2 fij(O2,:) scr(:,O2) + fab(O1,:) scr(O1,:) + s2 fab(O2,:) scr(O2,:)
  double part = 0.0;
  for (int i = threadIdx.x; i < na; i += blockDim.x)
    part += -fij[lr1 + (long)n * i] * xs[idx(i, lr1)] - s2 * fij[lr2 + (long)n * i] * xs[idx(i, lr2)];
  for (int j = nb + threadIdx.x; j < n; j += blockDim.x)
    part += fab[lr1 + (long)n * j] * xs[idx(lr1, j)] + s2 * fab[lr2 + (long)n * j] * xs[idx(lr2, j)];
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double dumn = red[0];
  // rows O1, O2 (j = nb .. n-1) and columns O1, O2 (i = 0 .. na-1); the four corner amplitudes are overwritten below
  for (int j = nb + threadIdx.x; j < n; j += blockDim.x) {
    e[idx(lr1, j)] += fab[j + (long)n * lr1] * xlr * isq2;
    e[idx(lr2, j)] += s2 * fab[j + (long)n * lr2] * xlr * isq2;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    e[idx(i, lr1)] -= fij[i + (long)n * lr1] * xlr * isq2;
    e[idx(i, lr2)] -= s2 * fij[i + (long)n * lr2] * xlr * isq2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    e[idx(lr1, lr1)] = dumn * isq2 + xlr * (fab[lr1 + (long)n * lr1] + fab[lr2 + (long)n * lr2] - fij[lr1 + (long)n * lr1] -
                                           fij[lr2 + (long)n * lr2]) * 0.5;
    if (kind == 1) {
      e[idx(lr2, lr2)] = 0.0;
    } else {
      e[idx(lr2, lr1)] = 0.0; e[idx(lr1, lr2)] = 0.0; e[idx(lr2, lr2)] = 0.0;
    }
  }
}
__global__ void k_sig_add(double* a, const double* b, long n) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) a[e] += b[e];
}

struct SigSession {
  oqpb_ctx* ctx = nullptr;
  int n = 0, na = 0, nb = 0, kind = 0;
  double scale = 1.0;
  DevBuf va, vb, fa, fb, X, Xs, W, E, d3, f3, TV, TC, CV, T1, tmp;
  void release() {
    for (DevBuf* b : {&va, &vb, &fa, &fb, &X, &Xs, &W, &E, &d3, &f3, &TV, &TC, &CV, &T1, &tmp}) b->release();
    ctx = nullptr;
  }
};
SigSession g_sig;

int sig_gemm(cudaStream_t st, int M, int N, int K, int batch, double alpha, const double* A, long sAi, long sAk, long sAb,
             const double* B, long sBk, long sBj, long sBb, double beta, double* C, long sCi, long sCj, long sCb,
             const double* alpha_b = nullptr, long sab = 0) {
  if (M <= 0 || N <= 0 || batch <= 0) return OQPB_OK;
  GemmOp g{M, N, K, batch, alpha, beta, alpha_b, sab, A, sAi, sAk, sAb, B, sBk, sBj, sBb, C, sCi, sCj, sCb};
  dim3 grid((N + 15) / 16, (M + 15) / 16, batch), block(16, 16);
  k_gemm_strided<<<grid, block, 0, st>>>(g);
  return cudaGetLastError() == cudaSuccess ? OQPB_OK : OQPB_ERR_CUDA;
}
#define SG(...) do { int rc_ = sig_gemm(st, __VA_ARGS__); if (rc_) return rc_; } while (0)

// one Davidson step for nv trial vectors: X (ntrial, nv) on the device -> W = (A-B) X
int sig_apply_dev(SigSession& s, int nv) {
  oqpb_ctx* ctx = s.ctx;
  cudaStream_t st = ctx->stream;
  const int n = s.n, na = s.na, nb = s.nb, kind = s.kind;
  const int lr1 = na - 2, lr2 = na - 1, nvir = n - na, nvb = n - nb;
  const long ntrial = (long)na * nvb, n2 = (long)n * n;
  const int NM = 7 * nv;
  const double* va = s.va.as<double>();
  const double* vb = s.vb.as<double>();
  const double* fa = s.fa.as<double>();
  const double* fb = s.fb.as<double>();
  const double* X = s.X.as<double>();
  // X_v(i, j) = X[xo(i, j) + ntrial v], j >= nb   (iatogen, tdhf_lib.F90:480-498)
  auto xo = [na, nb](int i, int j) { return (long)i + (long)na * (j - nb); };
  CK(s.d3.ensure((size_t)n2 * NM * sizeof(double)));
  CK(s.f3.ensure((size_t)n2 * NM * sizeof(double)));
  CK(s.TV.ensure((size_t)2 * nv * n * sizeof(double)));
  CK(s.TC.ensure((size_t)2 * nv * n * sizeof(double)));
  CK(s.CV.ensure((size_t)std::max(1, nb) * nv * n * sizeof(double)));
  CK(s.T1.ensure((size_t)na * n * nv * sizeof(double)));
  CK(s.tmp.ensure((size_t)4 * nv * n * sizeof(double)));
  CK(s.Xs.ensure((size_t)ntrial * nv * sizeof(double)));
  CK(s.W.ensure((size_t)ntrial * nv * sizeof(double)));
  CK(s.E.ensure((size_t)ntrial * nv * sizeof(double)));
  double* d3 = s.d3.as<double>();
  double* f3 = s.f3.as<double>();
  double* TV = s.TV.as<double>();   // [w][v][mu]: w = 0 row O2 (tv2), w = 1 row O1 (tv1)
  double* TC = s.TC.as<double>();   // [w][v][mu]: w = 0 column O1 (tc1), w = 1 column O2 (tc2)
  double* CV = s.CV.as<double>();   // [v][i][nu]
  double* tv2 = TV, *tv1 = TV + (long)nv * n, *tc1 = TC, *tc2 = TC + (long)nv * n;
  // ---- 6a: mrsfcbc for all vectors (component c of vector v at d3[((nu n + mu) NM + c nv + v])
  CK(cudaMemsetAsync(d3, 0, (size_t)n2 * NM * sizeof(double), st));
  CK(cudaMemsetAsync(TC, 0, (size_t)2 * nv * n * sizeof(double), st));
  const long cI = NM, cJ = (long)n * NM;  // strides of mu, nu inside a component
  auto comp = [&](int c) { return d3 + (long)c * nv; };
  // tv_w(mu) = sum_a C^b(mu, a) X(O_w, a), a virtual                     (:1001-1004, 1028-1031)
  SG(n, 1, nvir, nv, 1.0, vb + (long)na * n, 1, n, 0, X + xo(lr2, na), na, 0, ntrial, 0.0, tv2, 1, 0, n);
  SG(n, 1, nvir, nv, 1.0, vb + (long)na * n, 1, n, 0, X + xo(lr1, na), na, 0, ntrial, 0.0, tv1, 1, 0, n);
  if (nb > 0) {
    // tc_w(mu) = sum_i C^a(mu, i) X(i, O_w), i doubly occupied           (:1061-1064, 1088-1091)
    SG(n, 1, nb, nv, 1.0, va, 1, n, 0, X + xo(0, lr1), 1, 0, ntrial, 0.0, tc1, 1, 0, n);
    SG(n, 1, nb, nv, 1.0, va, 1, n, 0, X + xo(0, lr2), 1, 0, ntrial, 0.0, tc2, 1, 0, n);
  }
  // rank-1 updates: C(mu, nu) += alpha a(mu) b(nu)
  auto outer = [&](double* C, double alpha, const double* a, long sAb, const double* b, long sBb, const double* ab = nullptr,
                   long sab = 0) { return sig_gemm(st, n, n, 1, nv, alpha, a, 1, 0, sAb, b, 0, 1, sBb, 1.0, C, cI, cJ, 1, ab, sab); };
  int rc;
  for (int c : {0, 6}) if ((rc = outer(comp(c), 1.0, va + (long)n * lr2, 0, tv2, n))) return rc;        // bo2v  (:1008-1011), ball (:1164)
  for (int c : {1, 6}) if ((rc = outer(comp(c), 1.0, va + (long)n * lr1, 0, tv1, n))) return rc;        // bo1v  (:1035-1038)
  if (nb > 0) {
    for (int c : {2, 6}) if ((rc = outer(comp(c), 1.0, tc1, n, vb + (long)n * lr1, 0))) return rc;      // bco1  (:1068-1071)
    for (int c : {3, 6}) if ((rc = outer(comp(c), 1.0, tc2, n, vb + (long)n * lr2, 0))) return rc;      // bco2  (:1095-1098)
  }
  if ((rc = outer(comp(4), 1.0, tv2, n, va + (long)n * lr1, 0))) return rc;                             // o21v  (:1119-1138)
  if ((rc = outer(comp(4), -1.0, tv1, n, va + (long)n * lr2, 0))) return rc;
  if (nb > 0) {
    if ((rc = outer(comp(5), 1.0, vb + (long)n * lr2, 0, tc1, n))) return rc;                           // co12  (:1141-1161)
    if ((rc = outer(comp(5), -1.0, vb + (long)n * lr1, 0, tc2, n))) return rc;
    // ball += C^a_c (X_cv C^b_v^T):  CV_v(nu, i) = sum_a C^b(nu, a) X(i, a);  ball(mu, nu) += sum_i C^a(mu, i) CV_v(nu, i)   (:1167-1175)
    SG(n, nb, nvir, nv, 1.0, vb + (long)na * n, 1, n, 0, X + xo(0, na), na, 1, ntrial, 0.0, CV, 1, n, (long)n * nb);
    SG(n, n, nb, nv, 1.0, va, 1, n, 0, CV, n, 1, (long)n * nb, 1.0, comp(6), cI, cJ, 1);
  }
  const double isq2 = 0.70710678118654752440;
  const double* x11 = X + xo(lr1, lr1);  // X_v(O1, O1), stride ntrial over v
  if (kind == 1) {  // :1178-1186
    if ((rc = outer(comp(6), 1.0, va + (long)n * lr2, 0, vb + (long)n * lr1, 0, X + xo(lr2, lr1), ntrial))) return rc;
    if ((rc = outer(comp(6), 1.0, va + (long)n * lr1, 0, vb + (long)n * lr2, 0, X + xo(lr1, lr2), ntrial))) return rc;
    if ((rc = outer(comp(6), isq2, va + (long)n * lr1, 0, vb + (long)n * lr1, 0, x11, ntrial))) return rc;
    if ((rc = outer(comp(6), -isq2, va + (long)n * lr2, 0, vb + (long)n * lr2, 0, x11, ntrial))) return rc;
  } else {  // :1187-1193
    if ((rc = outer(comp(6), isq2, va + (long)n * lr1, 0, vb + (long)n * lr1, 0, x11, ntrial))) return rc;
    if ((rc = outer(comp(6), isq2, va + (long)n * lr2, 0, vb + (long)n * lr2, 0, x11, ntrial))) return rc;
  }
  // ---- 6b: the J/K build (int2_mrsf_data_t, scale_exchange = scale_coulomb = scale: tdhf_mrsf_energy.F90:739-752)
  if ((rc = mrsf_core(ctx, d3, f3, nv, 7, s.scale, s.scale))) return rc;
  // ---- 6c: mrsfmntoia.  Triplet: components 1..6 change sign (tdhf_mrsf_energy.F90:762-763) -> sg on every term that reads them
  const double sg = kind == 3 ? -1.0 : 1.0;
  auto fcomp = [&](int c) { return f3 + (long)c * nv; };
  double* W = s.W.as<double>();
  double* T1 = s.T1.as<double>();
  // W(i, j) = [C^a^T agdlr C^b](i, j), i < na, j >= nb, written straight into the amplitude layout      (:1604-1614)
  SG(na, n, n, nv, 1.0, va, n, 1, 0, fcomp(6), cI, cJ, 1, 0.0, T1, 1, na, (long)na * n);
  SG(na, nvb, n, nv, 1.0, T1, 1, na, (long)na * n, vb + (long)n * nb, 1, n, 0, 0.0, W, 1, na, ntrial);
  double* tA = s.tmp.as<double>(), *tB = tA + (long)nv * n, *tC = tB + (long)nv * n, *tD = tC + (long)nv * n;
  // tA = ado1v C^b(:,O2) + aco12 C^b(:,O1);  tB = ado2v C^b(:,O1) - aco12 C^b(:,O2)                       (:1617-1646)
  SG(n, 1, n, nv, sg, fcomp(1), cI, cJ, 1, vb + (long)n * lr2, 1, 0, 0, 0.0, tA, 1, 0, n);
  SG(n, 1, n, nv, sg, fcomp(5), cI, cJ, 1, vb + (long)n * lr1, 1, 0, 0, 1.0, tA, 1, 0, n);
  SG(n, 1, n, nv, sg, fcomp(0), cI, cJ, 1, vb + (long)n * lr1, 1, 0, 0, 0.0, tB, 1, 0, n);
  SG(n, 1, n, nv, -sg, fcomp(5), cI, cJ, 1, vb + (long)n * lr2, 1, 0, 0, 1.0, tB, 1, 0, n);
  if (na > 2) {
    SG(na - 2, 1, n, nv, 1.0, va, n, 1, 0, tA, 1, 0, n, 1.0, W + (long)na * (lr2 - nb), 1, 0, ntrial);
    SG(na - 2, 1, n, nv, 1.0, va, n, 1, 0, tB, 1, 0, n, 1.0, W + (long)na * (lr1 - nb), 1, 0, ntrial);
  }
  // tC = adco2^T C^a(:,O1) + ao21v^T C^a(:,O2);  tD = adco1^T C^a(:,O2) - ao21v^T C^a(:,O1)               (:1649-1674)
  SG(n, 1, n, nv, sg, fcomp(3), cJ, cI, 1, va + (long)n * lr1, 1, 0, 0, 0.0, tC, 1, 0, n);
  SG(n, 1, n, nv, sg, fcomp(4), cJ, cI, 1, va + (long)n * lr2, 1, 0, 0, 1.0, tC, 1, 0, n);
  SG(n, 1, n, nv, sg, fcomp(2), cJ, cI, 1, va + (long)n * lr2, 1, 0, 0, 0.0, tD, 1, 0, n);
  SG(n, 1, n, nv, -sg, fcomp(4), cJ, cI, 1, va + (long)n * lr1, 1, 0, 0, 1.0, tD, 1, 0, n);
  SG(nvir, 1, n, nv, 1.0, vb + (long)n * na, n, 1, 0, tC, 1, 0, n, 1.0, W + lr1 + (long)na * (na - nb), na, 0, ntrial);
  SG(nvir, 1, n, nv, 1.0, vb + (long)n * na, n, 1, 0, tD, 1, 0, n, 1.0, W + lr2 + (long)na * (na - nb), na, 0, ntrial);
  k_sig_mntoia_fix<<<(nv + 63) / 64, 64, 0, st>>>(W, ntrial, nv, na, nb, kind);
  CK(cudaGetLastError());
  // ---- 6c: mrsfesum   E = scr fab^T - fij scr on the occupied-alpha x virtual-beta block                 (:1951-1960)
  double* Xs = s.Xs.as<double>();
  double* E = s.E.as<double>();
  k_sig_scr<<<(unsigned)((ntrial * nv + 255) / 256), 256, 0, st>>>(X, Xs, ntrial, nv, na, nb);
  CK(cudaGetLastError());
  SG(na, nvb, nvb, nv, 1.0, Xs, 1, na, ntrial, fb + nb + (long)n * nb, n, 1, 0, 0.0, E, 1, na, ntrial);
  SG(na, nvb, na, nv, -1.0, fa, 1, n, 0, Xs, 1, na, ntrial, 1.0, E, 1, na, ntrial);
  k_sig_esum_fix<<<nv, 256, 0, st>>>(E, X, Xs, fa, fb, ntrial, n, na, nb, kind);
  CK(cudaGetLastError());
  k_sig_add<<<(unsigned)((ntrial * nv + 255) / 256), 256, 0, st>>>(W, E, ntrial * nv);
  CK(cudaGetLastError());
  return OQPB_OK;
}
#undef SG

}  // namespace

extern "C" {

// routec_sig.F90:28-37.  mo_a, mo_b: MO coefficients (nbf, nbf) column-major; fmo_a, fmo_b: alpha / beta Fock matrices in
// the MO basis (unpacked square).  kind: 1 singlet, 3 triplet.  Runs on the default context (oqpb_set_default_ctx), whose
// basis, cutoff and screening must be set (the reference raises the response cutoff to 1e-8 before, tdhf_mrsf_energy.F90:516-519).
int routec_sig_init(const int* nbf, const double* mo_a, const double* mo_b, const double* fmo_a, const double* fmo_b,
                    const int* nocca, const int* noccb, const int* kind) {
  oqpb_ctx* ctx = g_default_ctx;
  if (!ctx || !nbf || !mo_a || !mo_b || !fmo_a || !fmo_b || !nocca || !noccb || !kind) return OQPB_ERR_BAD_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (*nbf != ctx->nbf || *nocca < 2 || *noccb < 0 || *noccb != *nocca - 2 || *nocca > *nbf || (*kind != 1 && *kind != 3))
    return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  g_sig.release();
  g_sig.ctx = ctx; g_sig.n = *nbf; g_sig.na = *nocca; g_sig.nb = *noccb; g_sig.kind = *kind; g_sig.scale = 1.0;
  const size_t b = (size_t)*nbf * *nbf * sizeof(double);
  for (auto pr : {std::make_pair(&g_sig.va, mo_a), std::make_pair(&g_sig.vb, mo_b), std::make_pair(&g_sig.fa, fmo_a),
                  std::make_pair(&g_sig.fb, fmo_b)}) {
    CK(pr.first->ensure(b));
    CK(cudaMemcpyAsync(pr.first->p, pr.second, b, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return OQPB_OK;
}

// routec_sig.F90:38-42: the exact-exchange scale of the response (scale_exchange = scale_coulomb of int2_mrsf_data_t)
void routec_sig_set_scale(const double* s) {
  if (s) g_sig.scale = *s;
}

// routec_sig.F90:43-51: bvec_mo (ntrial, nv_new) in, sigma_mo (ntrial, nv_new) = (A-B) X out, ntrial = nocca (nbf - noccb),
// column-major, i fastest.  info = 0 on success; anything else makes the caller fall back to its native path.
void routec_sig_iter(const double* bvec_mo, const int* nv_new, double* sigma_mo, int* info) {
  if (info) *info = 1;
  SigSession& s = g_sig;
  if (!s.ctx || !bvec_mo || !nv_new || !sigma_mo || *nv_new < 1) return;
  oqpb_ctx* ctx = s.ctx;
  cudaSetDevice(ctx->device);
  const int nv = *nv_new;
  const size_t bytes = (size_t)s.na * (s.n - s.nb) * nv * sizeof(double);
  auto fail = [&](int rc) { if (info) *info = rc ? rc : 1; };
  if (s.X.ensure(bytes) != cudaSuccess) return fail(OQPB_ERR_CUDA);
  if (cudaMemcpyAsync(s.X.p, bvec_mo, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return fail(OQPB_ERR_CUDA);
  int rc = sig_apply_dev(s, nv);
  if (rc) return fail(rc);
  if (cudaMemcpyAsync(sigma_mo, s.W.p, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return fail(OQPB_ERR_CUDA);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(OQPB_ERR_CUDA);
  if (info) *info = 0;
}

// routec_sig.F90:52-55
void routec_sig_free(void) {
  if (g_sig.ctx) cudaSetDevice(g_sig.ctx->device);
  g_sig.release();
}

}  // extern "C"
