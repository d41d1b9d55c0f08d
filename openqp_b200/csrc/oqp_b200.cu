// libopenqp_b200: host side of the C ABI (include/oqp_b200.h) and the small device kernels around the
// ERI/digestion kernel family (eri_kernel.cuh): pair table, Schwarz, shell densities, quartet enumeration.
//
// Reference map (what each piece replaces):
//   k_pair_count / k_pair_fill   int2_pair_storage%alloc / int2_prepare_pair     int2_pairs.F90:74-266
//   schwarz()                    ints_exchange                                    int2.F90:1582-1737
//   k_shlden_*                   shlden / shltd / shell_den_screen_mrsf           int2.F90:999-1047, tdhf_lib.F90:300-325,
//                                                                                  tdhf_mrsf_lib.F90:189-214
//   k_enum                       the i,j,k,l loops + screen_ij / screen_ijkl      int2.F90:756-805, 963-986
//   run_build()                  int2_twoei                                       int2.F90:589-923
//   k_fock_post                  fock_jk post-scaling                             scf_addons.F90:1177-1185
// There is no CPU fallback: every compute entry needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen()ed when a multi-device context is created
#include <chrono>
#include <thread>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/oqp_b200.h"
#include "eri_kernel.cuh"
#include "rys_tables.inc"
#include "rys_tables_fine.inc"

using namespace oqpb;

namespace oqpb {
const ClassEntry* class_table(int pv);  // eri_inst_*.cu: [pure variant][quartet class]
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char buf_[512];                                                                              \
      snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                             \
      return OQPB_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

namespace {

constexpr int NPC = 10;  // pair classes ss ps pp ds dp dd fs fp fd ff
// Each pair class is split into contraction buckets (number of primitive pairs) so that the quartets of one
// launch have similar primitive-loop lengths (one thread = one quartet in the small kernels: less divergence).
constexpr int NBK = 4;
constexpr int NL = NPC * NBK;  // pair lists
inline int bucket_of(int pcnt) { return pcnt <= 1 ? 0 : (pcnt <= 6 ? 1 : (pcnt <= 24 ? 2 : 3)); }
inline int pc_of(int list) { return list / NBK; }
inline int pair_class(int la, int lb) { return la * (la + 1) / 2 + lb; }
inline int quartet_class(int pa, int pb) { return pa * (pa + 1) / 2 + pb; }
const int PC_LA[NPC] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3};
const int PC_LB[NPC] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3};

struct Cutoffs {
  double integral, pair, quartet, exponent;
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t b) {
    if (b <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, b);
    if (e == cudaSuccess) bytes = b;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return (T*)p; }
};

struct PairTable {
  // entries grouped by pair class; within a class sorted by Q descending once Schwarz is known
  std::vector<PairEntry> ent;
  std::vector<int> canon;  // canonical pair id tri(i,j), i>=j by shell index
  std::vector<double> Q;
  std::vector<double> Qsuf;  // per list: max of Q over the entries at or after this one
  int cls_off[NL + 1] = {0};  // offsets of the NL pair lists (class-major, bucket-minor)
  DevBuf d_ent, d_prim, d_Q, d_canon;
  long nprim = 0;
  // CAM second pass (int2.F90:674-685): Schwarz bounds of the Erf-attenuated integrals for the same entry order
  std::vector<double> Qatt, Qsufatt;
  DevBuf d_Qatt;
  double att_mu = 0.0;
};

// FLOP model of SURVEY.md 8(d-1) (restates int_rys.F90:406, 471-713)
double fprim_model(int l1, int l2, int l3, int l4) {
  int a[4] = {l1, l2, l3, l4};
  if (a[0] > a[1]) std::swap(a[0], a[1]);
  if (a[2] > a[3]) std::swap(a[2], a[3]);
  if (a[0] + a[1] > a[2] + a[3]) { std::swap(a[0], a[2]); std::swap(a[1], a[3]); }
  l1 = a[0]; l2 = a[1]; l3 = a[2]; l4 = a[3];
  int R = (l1 + l2 + l3 + l4) / 2 + 1, n = l1 + l2 + 1, m = l3 + l4 + 1;
  int coef = 15 + (m > 1 ? 6 : 0) + (n > 1 ? 6 : 0), vrr = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < m; j++) {
      if (i + j == 0) continue;
      else if (i + j == 1) vrr += 1;
      else if (i == 1 && j == 1) vrr += 3;
      else if (i < 2 || j < 2) vrr += 4;
      else vrr += 7;
    }
  int hrr = 0;
  for (int k = 1; k <= l3; k++) hrr += 2 * n * (m - k);
  for (int i = 1; i <= l1; i++) hrr += 2 * (l3 + 1) * (l4 + 1) * (n - i);
  int N = ncart(l1) * ncart(l2) * ncart(l3) * ncart(l4);
  return 40.0 * R + R * (double)(coef + 3 * vrr + 3 * hrr + 3 * N);
}

}  // namespace

// Work plan of a build: per (bra list, ket list) the ket bound index of every bra of this rank and the launch chunks.
// It depends on the density only through 4*max|D| rounded UP to a power of two (a larger bound only admits more
// candidates to the exact test in k_enum), so consecutive SCF iterations reuse it: no host planning, no upload.
struct PlanChunk { int pca, pcb, p0, p1; size_t cand; int nr, rk; };  // nr / rk: bra stride and offset of this chunk (rank split, or 1 / 0 for a
                                                                       // list pair given to one rank as a whole)
struct BuildPlan {
  bool valid = false;
  double bound4 = -1.0;
  long gen = -1;
  std::vector<PlanChunk> chunks;
  std::vector<size_t> km_off;
  std::vector<std::pair<int, int>> cps;
  long long total_local = 0;
  DevBuf d_km;  // this plan's own ket-bound arrays (a shared buffer let the attenuated pass overwrite the regular plan's)
};

struct oqpb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  // basis (host copy)
  int nshell = 0, nprim = 0, nbf = 0, harmonic_active = 0, lmax = 0;
  std::vector<int> am, harm, ncontr, goff, aooff, naos;
  std::vector<double> ex, cc, cen;
  int pure_l[4] = {0, 0, 0, 0};
  bool have_basis = false, have_cutoff = false, have_screen = false;
  // device basis
  DevBuf d_am, d_ncontr, d_goff, d_aooff, d_naos, d_ex, d_cc, d_xyz;
  DevBuf d_rys;  // Rys tables
  double cutoff = 5e-11;
  Cutoffs cut{};
  PairTable run;
  std::vector<double> Qmat;  // nshell x nshell (host)
  std::vector<double> Qmat_att;  // same for the attenuated integrals (CAM)
  DevBuf d_Qmat, d_dsh, d_maxden, d_ok, d_d4, d_rowsbuf;
  // work
  static constexpr int NSTREAM = 8;  // launch lanes (nlanes of them used): chunk c runs on lane c % NSTREAM (own task buffer) so that the
                                     // tail of one class kernel overlaps the next enumeration / class kernel
  DevBuf d_tasks[NSTREAM], d_items[NSTREAM], d_counters, d_Dsq, d_F, d_Din, d_stats, d_gen_in, d_gen_out;
  cudaStream_t lane[NSTREAM] = {};
  cudaEvent_t lane_ev[NSTREAM] = {};
  int nlanes = 0;      // OQPB_NLANES (0 = by build size: 4 or 8, see run_build)
  int grid_pct = 50;   // OQPB_GRID_PCT: scales the per-class grid caps (100 -> 50 with 2^25 tasks per chunk: w32 1352 -> 1318 ms)
  bool use_run = true;   // OQPB_RUN=0: task kernels only
  bool use_graph = false;            // OQPB_GRAPH=1: replay the launch section as a CUDA graph (measured neutral, see run_build)
  size_t graph_max_chunks = 600;     // OQPB_GRAPH_MAX: builds with more chunks are launched eagerly (launch latency is hidden there)
  struct GraphSlot {
    std::string key;
    cudaGraphExec_t exec = nullptr;
    void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; key.clear(); }
  } graph[2];                        // regular / attenuated plan
  double whole_ms = 2.5;   // OQPB_WHOLE_MS: a list pair is shared by round(estimated ms / whole_ms) ranks (0 = always by all ranks)
  int use_kown = 1;      // OQPB_KOWN: 0 = never the ket-owner group kernel, 1 = the classes it wins (default), 2 = every class it covers
  int run_max_bucket_sum = 2;  // OQPB_RUN_BUCKETS
  size_t wpq_max_tasks = 16384;  // OQPB_WPQ_MAX: largest launch (candidate quartets) that uses the warp-per-quartet kernels
  cudaEvent_t fork_ev = nullptr;
  size_t task_cap = (size_t)1 << 25;  // quartets per chunk and stream lane (256 MB each); 2^23: +2.6 % launches/tails on w32
  int rank = 0, nranks = 1;
  // stats of the last build
  long long st_survivors = 0, st_skipped = 0, st_launches = 0;
  double st_flops = 0, st_kernel_ms = 0;
  bool record = false;
  bool profile = false;
  double prof[55][4] = {};  // per quartet class: ms, quartets, primitive quartets, model flops
  std::vector<int> rec;  // recorded shell quadruples
  unsigned* h_counts = nullptr;  // pinned
  size_t h_counts_cap = 0;
  double fp64_peak = 0;
  // multi-device context (oqpb_ctx_create_multi): the master holds the peers; every member has its communicator
  std::vector<oqpb_ctx*> peers;  // master only: devices 1 .. ndev-1
  ncclComm_t comm = nullptr;
  int mdev = 0, mndev = 1;       // index of this member / number of members
  int base_rank = 0, base_nranks = 1;  // partition requested by the caller (MPI rank split), before the device split
  DevBuf d_mask;  // optional bra shell-pair mask (oqpb_set_bra_mask)
  bool have_mask = false;
  BuildPlan plan[2];  // [0] regular, [1] attenuated pass
  long plan_gen = 0;  // bumped by set_basis / set_cutoff / set_screening / set_partition
};

// ===================================================================================== device kernels
namespace {

__device__ __forceinline__ void pair_from_index(long id, int& i, int& j) {
  // id = i(i+1)/2 + j, j <= i
  i = (int)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
  while ((long)i * (i + 1) / 2 > id) --i;
  while ((long)(i + 1) * (i + 2) / 2 <= id) ++i;
  j = (int)(id - (long)i * (i + 1) / 2);
}

// int2_pairs.F90:179-266: predicate of the fill pass; order: lower-AM shell's primitives outermost
template <bool FILL>
__global__ void k_pairs(int nshell, long npairs, const int* __restrict__ am, const int* __restrict__ ncontr,
                        const int* __restrict__ goff, const double* __restrict__ ex, const double* __restrict__ cc,
                        const double* __restrict__ xyz, double exponent_cutoff, double quartet_cutoff,
                        int* __restrict__ cnt, PairEntry* __restrict__ ent, double* __restrict__ prim, long nent) {
  long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int i, j, poff = 0;
  if (FILL) {
    if (id >= nent) return;
    i = ent[id].sa;
    j = ent[id].sb;
    poff = ent[id].poff;
    if (ent[id].pcnt == 0) return;
  } else {
    if (id >= npairs) return;
    pair_from_index(id, i, j);
  }
  int sha = i, shb = j;  // reference order: lower AM first (int2_pairs.F90:202-212); equal AM: (i, j)
  if (FILL) {
    // entries store the higher-AM shell first; the reference's outer loop is the lower-AM shell
    if (am[i] > am[j]) { sha = j; shb = i; }
  } else {
    if (am[i] > am[j]) { sha = j; shb = i; }
  }
  const double ax = xyz[3 * sha], ay = xyz[3 * sha + 1], az = xyz[3 * sha + 2];
  const double bx = xyz[3 * shb], by = xyz[3 * shb + 1], bz = xyz[3 * shb + 2];
  const double ab2 = (ax - bx) * (ax - bx) + (ay - by) * (ay - by) + (az - bz) * (az - bz);
  const double sqrtpito52 = 5.914967172795612486;  // sqrt(2) * pi^(5/4)
  int n = 0;
  for (int p1 = 0; p1 < ncontr[sha]; ++p1)
    for (int p2 = 0; p2 < ncontr[shb]; ++p2) {
      double a1 = ex[goff[sha] + p1], a2 = ex[goff[shb] + p2];
      double gam = a1 + a2;
      double e12 = a1 * a2 * ab2;
      if (e12 > gam * exponent_cutoff) continue;
      double gi = 1.0 / gam;
      e12 = e12 * gi;
      double k1 = cc[goff[sha] + p1] * cc[goff[shb] + p2] * exp(-e12);
      if (fabs(k1) < quartet_cutoff) continue;
      if (FILL) {
        double* o = prim + (size_t)(poff + n) * PRIM_STRIDE;
        o[0] = (a1 * ax + a2 * bx) * gi;
        o[1] = (a1 * ay + a2 * by) * gi;
        o[2] = (a1 * az + a2 * bz) * gi;
        o[3] = gam;
        o[4] = sqrtpito52 * k1 * gi;  // da = K * ginv (int_rys.F90:216)
        o[5] = gi;
      }
      ++n;
    }
  if (!FILL) {
    cnt[id] = n;
  } else {
    // sort the pair's primitives by |K|/zeta descending (lets the kernels prune the primitive-quartet loops)
    // and record the smallest zeta
    double* o = prim + (size_t)poff * PRIM_STRIDE;
    double zmin = 1e300;
    for (int a = 0; a < n; ++a) zmin = fmin(zmin, o[a * PRIM_STRIDE + 3]);
    for (int a = 1; a < n; ++a) {
      double rec[PRIM_STRIDE];
      for (int k = 0; k < PRIM_STRIDE; ++k) rec[k] = o[a * PRIM_STRIDE + k];
      const double key = fabs(rec[4]);
      int b = a - 1;
      while (b >= 0 && fabs(o[b * PRIM_STRIDE + 4]) < key) {
        for (int k = 0; k < PRIM_STRIDE; ++k) o[(b + 1) * PRIM_STRIDE + k] = o[b * PRIM_STRIDE + k];
        --b;
      }
      for (int k = 0; k < PRIM_STRIDE; ++k) o[(b + 1) * PRIM_STRIDE + k] = rec[k];
    }
    ent[id].zmin = n > 0 ? zmin : 0.0;
  }
}

__global__ void k_expand_packed(const double* __restrict__ dp, double* __restrict__ dsq, int nbf) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)nbf * nbf) return;
  int a = (int)(e / nbf), b = (int)(e % nbf);
  dsq[e] = dp[tri_idx(a, b)];
}
__global__ void k_add(const double* a, const double* b, double* c, long n) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) c[e] = a[e] + b[e];
}

// shlden int2.F90:999-1047 (packed densities, all focks)
__global__ void k_shlden_packed(int nshell, long npairs, const int* __restrict__ aooff, const int* __restrict__ naos,
                                const double* __restrict__ d, int nfocks, long ntri, double* __restrict__ dsh,
                                unsigned long long* maxden) {
  long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= npairs) return;
  int si, sj;
  pair_from_index(id, si, sj);
  int mini = aooff[si], maxi = mini + naos[si] - 1, minj = aooff[sj], maxj0 = minj + naos[sj] - 1;
  double dmax = 0.0;
  for (int f = 0; f < nfocks; ++f)
    for (int i = mini; i <= maxi; ++i) {
      int maxj = (si == sj) ? i : maxj0;
      for (int j = minj; j <= maxj; ++j) dmax = fmax(dmax, fabs(d[(size_t)f * ntri + tri_idx(i, j)]));
    }
  dsh[(size_t)si * nshell + sj] = dmax;
  dsh[(size_t)sj * nshell + si] = dmax;
  atomicMax(maxden, (unsigned long long)__double_as_longlong(dmax));
}
// shltd / shell_den_screen_mrsf: dsh(I,J) (I>=J) = max |X(m, mu in J, nu in I)| over the interleaved
// layout X[(nu*nbf + mu)*NM + m]
__global__ void k_shlden_gen(int nshell, long npairs, const int* __restrict__ aooff, const int* __restrict__ naos,
                             const double* __restrict__ X, int NM, int nbf, double* __restrict__ dsh,
                             unsigned long long* maxden) {
  long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= npairs) return;
  int si, sj;
  pair_from_index(id, si, sj);
  double dmax = 0.0;
  for (int nu = aooff[si]; nu < aooff[si] + naos[si]; ++nu)
    for (int mu = aooff[sj]; mu < aooff[sj] + naos[sj]; ++mu) {
      const double* p = X + ((size_t)nu * nbf + mu) * NM;
      for (int m = 0; m < NM; ++m) dmax = fmax(dmax, fabs(p[m]));
    }
  dsh[(size_t)si * nshell + sj] = dmax;
  dsh[(size_t)sj * nshell + si] = dmax;
  atomicMax(maxden, (unsigned long long)__double_as_longlong(dmax));
}

// per-entry build data: ok = bra-level test passes (screen_ij, int2.F90:763-772); d4 = 4*dsh(sa,sb)
// mask (optional): bra shell pairs (canonical id i(i+1)/2+j) taking part in this build -- the sampled-parity hook
__global__ void k_entry_screen(long nent, const PairEntry* __restrict__ ent, const double* __restrict__ Q,
                               const double* __restrict__ dsh, int nshell, const unsigned long long* maxden,
                               double cutoff, int* __restrict__ ok, double* __restrict__ d4,
                               const int* __restrict__ canon, const unsigned char* __restrict__ mask) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nent) return;
  double md = __longlong_as_double((long long)*maxden);
  ok[e] = !(__dmul_rn(Q[e], md) < cutoff) && (mask == nullptr || mask[canon[e]] != 0);
  d4[e] = 4.0 * dsh[(size_t)ent[e].sa * nshell + ent[e].sb];
}

// Quartet enumeration for one (bra list, ket list) chunk: CTA per bra entry, threads over ket entries.
// Predicate = screen_ijkl, int2.F90:975-986, evaluated in the reference's operation order without FMA.
// Kets at or beyond kmax[bra] cannot survive (suffix maximum of the Schwarz bounds with 4*max_den) and are not
// visited.  Two passes (count, write) so that the survivors of a bra occupy ONE contiguous segment of the task
// buffer in ket-list order: a warp of the ERI kernels then sees runs of quartets with the same bra and the same
// ket shell c.
constexpr int ENUM_NT = 256;
constexpr int ENUM_BITS_MAX = 1 << 17;  // kets per bra whose test outcomes are kept as bits in shared memory (16 KB)
__global__ void __launch_bounds__(ENUM_NT)
k_enum(const PairEntry* __restrict__ bra, const PairEntry* __restrict__ ket, const double* __restrict__ Qb,
       const double* __restrict__ Qk, const double* __restrict__ d4b, const double* __restrict__ d4k,
       const int* __restrict__ okb, const int* __restrict__ okk, const int* __restrict__ canb,
       const int* __restrict__ cank, const int* __restrict__ kmax, int p0, int p1, int pstride, int diag,
       const double* __restrict__ dsh, int nshell, double cutoff, int2* __restrict__ tasks, unsigned* __restrict__ count,
       unsigned cap, int use_smem, int2* __restrict__ items, unsigned* __restrict__ nitems, unsigned item_cap, int bits_cap) {
  extern __shared__ double rows[];
  __shared__ unsigned s_w[ENUM_NT / 32];
  __shared__ unsigned s_base, s_ibase;
  int p = p0 + blockIdx.x * pstride;
  if (p >= p1) return;
  const PairEntry eb = bra[p];
  const int nk = diag ? min(kmax[p], p + 1) : kmax[p];
  if (nk <= 0) return;
  const double* rowa = dsh + (size_t)eb.sa * nshell;
  const double* rowb = dsh + (size_t)eb.sb * nshell;
  if (use_smem) {
    for (int s = threadIdx.x; s < nshell; s += blockDim.x) {
      rows[s] = rowa[s];
      rows[nshell + s] = rowb[s];
    }
    __syncthreads();
    rowa = rows;
    rowb = rows + nshell;
  }
  const double qb = Qb[p], db4 = d4b[p];
  const int okp = okb[p], canp = canb[p];
  auto test = [&](int q) {
    const PairEntry& ek = ket[q];
    const int sc = ek.sa, sd = ek.sb;
    double m = fmax(fmax(fmax(db4, d4k[q]), fmax(rowb[sd], rowb[sc])), fmax(rowa[sd], rowa[sc]));
    double res = __dmul_rn(__dmul_rn(qb, Qk[q]), m);
    int bra_ok = (canp >= cank[q]) ? okp : okk[q];  // the canonically larger pair is the reference's bra
    return bra_ok && !(res < cutoff);
  };
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // pass 1: count; the outcome of every test is kept as one bit (one word per 32 consecutive kets) so that pass 2 does not
  // load the ket entries, bounds and density rows a second time
  unsigned* bits = reinterpret_cast<unsigned*>(rows + (use_smem ? 2 * nshell : 0));
  const bool keep_bits = nk <= bits_cap;
  unsigned cnt = 0;
  for (int q0 = 0; q0 < nk; q0 += ENUM_NT) {
    const int q = q0 + threadIdx.x;
    const bool sv = q < nk && test(q);
    const unsigned m = __ballot_sync(0xffffffffu, sv);
    if (keep_bits && lane == 0 && q < nk) bits[q >> 5] = m;  // (lane 0 holds the first ket of the warp's 32)
    cnt += sv ? 1u : 0u;
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0) s_w[w] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int k = 0; k < ENUM_NT / 32; ++k) tot += s_w[k];
    s_base = tot ? atomicAdd(count, tot) : 0u;
    s_w[0] = tot;  // reused as flag below
    // warp work items of the run kernels: the bra's contiguous survivors cut into pieces of RUN_LEN kets
    if (items && tot) s_ibase = atomicAdd(nitems, (tot + RUN_LEN - 1) / RUN_LEN);
  }
  __syncthreads();
  if (s_w[0] == 0) return;
  unsigned running = s_base;
  if (items) {
    const unsigned tot = s_w[0], nit = (tot + RUN_LEN - 1) / RUN_LEN;
    for (unsigned k = threadIdx.x; k < nit; k += ENUM_NT) {
      const unsigned start = running + k * RUN_LEN, len = min((unsigned)RUN_LEN, tot - k * RUN_LEN);
      if (s_ibase + k < item_cap)  // a piece past the task buffer gets length 0 (the host reports the overflow)
        items[s_ibase + k] = make_int2((int)((unsigned)p | ((start + len <= cap ? len : 0u) << 24)), (int)start);
    }
  }
  __syncthreads();
  // pass 2: write in ket-list order
  for (int q0 = 0; q0 < nk; q0 += ENUM_NT) {
    const int q = q0 + threadIdx.x;
    unsigned mask;
    bool surv;
    if (keep_bits) {
      mask = (q0 + (w << 5)) < nk ? bits[(q0 >> 5) + w] : 0u;
      surv = (mask >> lane) & 1u;
    } else {
      surv = q < nk && test(q);
      mask = __ballot_sync(0xffffffffu, surv);
    }
    if (lane == 0) s_w[w] = __popc(mask);
    __syncthreads();
    unsigned off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < ENUM_NT / 32; ++k) {
      const unsigned c = s_w[k];
      if (k < w) off += c;
      tot += c;
    }
    if (surv) {
      const unsigned pos = running + off + __popc(mask & ((1u << lane) - 1));
      if (pos < cap) tasks[pos] = make_int2(p, q);
    }
    running += tot;
    __syncthreads();
  }
}

__global__ void k_fock_post(double* f, int nbf, long ntri, int nfocks) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ntri * nfocks) return;
  long t = e % ntri;
  int i = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long)i * (i + 1) / 2 > t) --i;
  while ((long)(i + 1) * (i + 2) / 2 <= t) ++i;
  long j = t - (long)i * (i + 1) / 2;
  double v = 0.5 * f[e];
  if (j == i) v *= 2.0;
  f[e] = v;
}

// TD helpers: d2(mu,nu,v) column-major -> interleaved X[(nu*nbf+mu)*NM + m] with m = v (+ nvec for the
// antisymmetric copy):  comp 0: P + P^T, comp 1: P - P^T   (see oqpb_jk_td)
__global__ void k_td_pack(const double* __restrict__ d2, double* __restrict__ X, int nbf, int nvec, int ncomp, int mode) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n2 = (long)nbf * nbf;
  if (e >= n2 * nvec) return;
  int v = (int)(e / n2);
  long r = e % n2;
  int nu = (int)(r / nbf), mu = (int)(r % nbf);
  double p = d2[e], pt = d2[(size_t)v * n2 + (size_t)mu * nbf + nu];
  int NM = nvec * ncomp;
  if (mode == 0) {  // plain (TDA)
    X[(size_t)r * NM + v] = p;
  } else {
    X[(size_t)r * NM + v] = p + pt;
    if (ncomp > 1) X[(size_t)r * NM + nvec + v] = p - pt;
  }
}
__global__ void k_td_unpack(const double* __restrict__ X, double* __restrict__ out, int nbf, int nvec, int NM, int comp) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long n2 = (long)nbf * nbf;
  if (e >= n2 * nvec) return;
  int v = (int)(e / n2);
  long r = e % n2;
  out[e] = X[(size_t)r * NM + comp * nvec + v];
}

// Generic J/K: slab m of the interleaved X[(nu*nbf+mu)*NM + m] from a column-major nbf x nbf matrix P (mu fastest):
// op 0: P   1: P + P^T   2: P^T - P   3: P^T ; accumulate != 0 adds to the slab (sums of matrices)
__global__ void k_slab_pack(const double* __restrict__ P, double* __restrict__ X, int nbf, int NM, int m, int op, int accumulate) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)nbf * nbf) return;
  int nu = (int)(e / nbf), mu = (int)(e % nbf);
  double p = P[e], pt = P[(size_t)mu * nbf + nu];
  double v = op == 0 ? p : (op == 1 ? p + pt : (op == 2 ? pt - p : pt));
  double* x = X + (size_t)e * NM + m;
  *x = accumulate ? *x + v : v;
}
__global__ void k_slab_unpack(const double* __restrict__ X, double* __restrict__ out, int nbf, int NM, int m, double scale, int accumulate) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)nbf * nbf) return;
  double v = scale * X[(size_t)e * NM + m];
  out[e] = accumulate ? out[e] + v : v;
}
// UMRSF: components 9 and 10 (0-based 8, 9) of d3(v, c, mu, nu) enter the exchange transposed (tdhf_mrsf_lib.F90:393-400)
__global__ void k_umrsf_prepare(const double* __restrict__ d3, double* __restrict__ X, int nbf, int nvec, int ncomp) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int NM = nvec * ncomp;
  if (e >= (long)nbf * nbf * NM) return;
  int m = (int)(e % NM);
  long r = e / NM;
  int nu = (int)(r / nbf), mu = (int)(r % nbf);
  int c = m / nvec;
  X[e] = (c == 8 || c == 9) ? d3[((size_t)mu * nbf + nu) * NM + m] : d3[e];
}

__global__ void k_rys_test(EriArgs A, int nroots, int npts, const double* x, double* t2, double* w) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  for (int f = 0; f < 2 * nroots; ++f) {
    double v = 0;
    switch (nroots) {
      case 1: v = rys_eval<1>(A, x[i], f); break;
      case 2: v = rys_eval<2>(A, x[i], f); break;
      case 3: v = rys_eval<3>(A, x[i], f); break;
      case 4: v = rys_eval<4>(A, x[i], f); break;
      case 5: v = rys_eval<5>(A, x[i], f); break;
      case 6: v = rys_eval<6>(A, x[i], f); break;
      case 7: v = rys_eval<7>(A, x[i], f); break;
    }
    if (f < nroots) t2[(size_t)i * nroots + f] = v;
    else w[(size_t)i * nroots + f - nroots] = v;
  }
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

// ===================================================================================== host helpers
namespace {

template <class T>
int upload(oqpb_ctx* ctx, DevBuf& b, const std::vector<T>& v) {
  CK(b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return OQPB_OK;
}

int free_pairtable(PairTable& t) {
  t.d_ent.release(); t.d_prim.release(); t.d_Q.release(); t.d_canon.release(); t.d_Qatt.release();
  t.ent.clear(); t.canon.clear(); t.Q.clear(); t.Qatt.clear(); t.Qsufatt.clear(); t.att_mu = 0.0;
  return 0;
}

// Build a pair table for the given cutoffs.  Entries: every shell pair i>=j (zero-primitive pairs kept so the
// quartet bookkeeping -- nschwz -- matches the reference's loops exactly), grouped by pair class.
int build_pairtable(oqpb_ctx* ctx, const Cutoffs& c, PairTable& T, const std::vector<double>* Qmat) {
  const int ns = ctx->nshell;
  const long npairs = (long)ns * (ns + 1) / 2;
  DevBuf d_cnt;
  CK(d_cnt.ensure(npairs * sizeof(int)));
  int thr = 128;
  k_pairs<false><<<(unsigned)((npairs + thr - 1) / thr), thr, 0, ctx->stream>>>(
      ns, npairs, ctx->d_am.as<int>(), ctx->d_ncontr.as<int>(), ctx->d_goff.as<int>(), ctx->d_ex.as<double>(),
      ctx->d_cc.as<double>(), ctx->d_xyz.as<double>(), c.exponent, c.quartet, d_cnt.as<int>(), nullptr, nullptr, 0);
  CK(cudaGetLastError());
  std::vector<int> cnt(npairs);
  CK(cudaMemcpyAsync(cnt.data(), d_cnt.p, npairs * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  d_cnt.release();
  // group by class
  std::vector<std::vector<int>> bycls(NL);
  {
    long id = 0;
    for (int i = 0; i < ns; ++i)
      for (int j = 0; j <= i; ++j, ++id) {
        int la = ctx->am[i], lb = ctx->am[j];
        int pc = la >= lb ? pair_class(la, lb) : pair_class(lb, la);
        bycls[pc * NBK + bucket_of(cnt[id])].push_back((int)id);
      }
  }
  if (Qmat) {
    for (int pc = 0; pc < NL; ++pc) {
      auto& v = bycls[pc];
      auto qof = [&](int id) {
        int i = (int)((std::sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
        while ((long)i * (i + 1) / 2 > id) --i;
        while ((long)(i + 1) * (i + 2) / 2 <= id) ++i;
        int j = id - i * (i + 1) / 2;
        return (*Qmat)[(size_t)i * ns + j];
      };
      // order: Schwarz bound in descending factor-of-4 bins (keeps the enumeration's early cut-off), inside a bin by
      // the pair's first (higher angular momentum) shell, so that the surviving kets of a bra come in runs with
      // the same shell c: the digestion reduces J_ab and the K_ac / K_bc updates of a run inside the warp
      auto first_shell = [&](int id) {
        int i = (int)((std::sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
        while ((long)i * (i + 1) / 2 > id) --i;
        while ((long)(i + 1) * (i + 2) / 2 <= id) ++i;
        int j = id - i * (i + 1) / 2;
        return ctx->am[i] >= ctx->am[j] ? i : j;
      };
      struct Key { int bin; int pc; int sa; double q; int id; };
      std::vector<Key> key(v.size());
      // ... and inside a bin first by primitive-pair count: the lanes of a warp (one quartet each) then run primitive
      // loops of similar length (w32: -4 % build time); OQPB_SORT_PCNT=0 restores the pure shell order
      static const int sort_pcnt = getenv("OQPB_SORT_PCNT") ? atoi(getenv("OQPB_SORT_PCNT")) : 1;
      for (size_t k = 0; k < v.size(); ++k) {
        double q = qof(v[k]);
        int e = q > 0 ? std::ilogb(q) : -100000;
        key[k] = {e >= 0 ? e / 2 : -((-e + 1) / 2), sort_pcnt ? cnt[v[k]] : 0, first_shell(v[k]), q, v[k]};
      }
      std::sort(key.begin(), key.end(), [](const Key& a, const Key& b) {
        if (a.bin != b.bin) return a.bin > b.bin;
        if (a.pc != b.pc) return a.pc > b.pc;
        if (a.sa != b.sa) return a.sa < b.sa;
        if (a.q != b.q) return a.q > b.q;
        return a.id < b.id;
      });
      for (size_t k = 0; k < v.size(); ++k) v[k] = key[k].id;
    }
  }
  T.ent.clear(); T.canon.clear(); T.Q.clear();
  T.Qatt.clear(); T.Qsufatt.clear(); T.att_mu = 0.0;
  long poff = 0;
  for (int pc = 0; pc < NL; ++pc) {
    T.cls_off[pc] = (int)T.ent.size();
    for (int id : bycls[pc]) {
      int i = (int)((std::sqrt(8.0 * id + 1.0) - 1.0) * 0.5);
      while ((long)i * (i + 1) / 2 > id) --i;
      while ((long)(i + 1) * (i + 2) / 2 <= id) ++i;
      int j = id - i * (i + 1) / 2;
      PairEntry e;
      if (ctx->am[i] >= ctx->am[j]) { e.sa = i; e.sb = j; } else { e.sa = j; e.sb = i; }
      e.poff = (int)poff;
      e.pcnt = cnt[id];
      e.zmin = 0.0;
      const double* xa = &ctx->cen[3 * e.sa];
      const double* xb = &ctx->cen[3 * e.sb];
      e.ax = xa[0]; e.ay = xa[1]; e.az = xa[2];
      e.abx = xa[0] - xb[0]; e.aby = xa[1] - xb[1]; e.abz = xa[2] - xb[2];
      e.oa = ctx->aooff[e.sa]; e.ob = ctx->aooff[e.sb];
      poff += cnt[id];
      T.ent.push_back(e);
      T.canon.push_back(id);
      T.Q.push_back(Qmat ? (*Qmat)[(size_t)i * ns + j] : 0.0);
    }
  }
  T.cls_off[NL] = (int)T.ent.size();
  T.Qsuf = T.Q;
  for (int pc = 0; pc < NL; ++pc)
    for (int k = T.cls_off[pc + 1] - 2; k >= T.cls_off[pc]; --k) T.Qsuf[k] = std::max(T.Qsuf[k], T.Qsuf[k + 1]);
  T.nprim = poff;
  if (poff > 2000000000L) { ctx->err = "pair table too large"; return OQPB_ERR_UNSUPPORTED; }
  int rc;
  if ((rc = upload(ctx, T.d_ent, T.ent))) return rc;
  if ((rc = upload(ctx, T.d_canon, T.canon))) return rc;
  if ((rc = upload(ctx, T.d_Q, T.Q))) return rc;
  CK(T.d_prim.ensure(std::max<size_t>(poff, 1) * PRIM_STRIDE * sizeof(double)));
  long nent = (long)T.ent.size();
  k_pairs<true><<<(unsigned)((nent + thr - 1) / thr), thr, 0, ctx->stream>>>(
      ns, npairs, ctx->d_am.as<int>(), ctx->d_ncontr.as<int>(), ctx->d_goff.as<int>(), ctx->d_ex.as<double>(),
      ctx->d_cc.as<double>(), ctx->d_xyz.as<double>(), c.exponent, c.quartet, nullptr, T.d_ent.as<PairEntry>(),
      T.d_prim.as<double>(), nent);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return OQPB_OK;
}

// c0 + sum_k c_k T_k(t)  ->  sum_j a_j t^j for every block of nc coefficients, accumulated in long double (the integer
// Chebyshev coefficients grow to 2^(nc-2): 64-bit mantissas keep the result correctly rounded to double)
void cheb_to_monomial(const double* in, size_t n, int nc, double* out) {
  std::vector<std::vector<long double>> T(nc, std::vector<long double>(nc, 0.0L));  // T[k][j]: coefficient of t^j in T_k
  T[0][0] = 1.0L;
  if (nc > 1) T[1][1] = 1.0L;
  for (int k = 2; k < nc; ++k)
    for (int j = 0; j < nc; ++j) T[k][j] = (j > 0 ? 2.0L * T[k - 1][j - 1] : 0.0L) - T[k - 2][j];
  for (size_t b = 0; b + nc <= n; b += nc)
    for (int j = 0; j < nc; ++j) {
      long double a = 0.0L;
      for (int k = nc - 1; k >= j; --k) a += (long double)in[b + k] * T[k][j];
      out[b + j] = (double)a;
    }
}

// table of one nroots: the fine format (RysFmt<R>, quarter intervals x 8 terms) for nroots <= RYSF_MAXR
static_assert(RYS_FINE_MAXR <= RYSF_MAXR && RYSF_NCOEF == RysFmt<1>::NC && RYSF_DIV == RysFmt<1>::DIV &&
                  RYS_NCOEF == RysFmt<3>::NC,
              "rys_tables*.inc and RysFmt disagree");
const double* rys_table(oqpb_ctx* ctx, int R) {
  return R <= RYS_FINE_MAXR ? ctx->d_rys.as<double>() + RYS_NTAB + RYSF_OFF_H[R - 1] : ctx->d_rys.as<double>() + RYS_OFF_H[R - 1];
}

void fill_common_args(oqpb_ctx* ctx, const PairTable& T, int la_, int lb_, EriArgs& A) {
  memset(&A, 0, sizeof A);
  A.bra = T.d_ent.as<PairEntry>() + T.cls_off[la_];
  A.ket = T.d_ent.as<PairEntry>() + T.cls_off[lb_];
  const int pca = pc_of(la_), pcb = pc_of(lb_);
  A.prim = T.d_prim.as<double>();
  A.xyz = ctx->d_xyz.as<double>();
  A.aooff = ctx->d_aooff.as<int>();
  int R = (PC_LA[pca] + PC_LB[pca] + PC_LA[pcb] + PC_LB[pcb]) / 2 + 1;
  A.rys_tab = rys_table(ctx, R);
  A.rys_xmax = RYS_XMAX_H[R - 1];
  for (int k = 0; k < 7; ++k) { A.herm_r[k] = RYS_HERM_R_H[R - 1][k]; A.herm_w[k] = RYS_HERM_W_H[R - 1][k]; }
  A.nbf = ctx->nbf;
  A.mu2inv = 0.0;
}

int ensure_counts(oqpb_ctx* ctx, size_t n) {
  if (n <= ctx->h_counts_cap) return OQPB_OK;
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  ctx->h_counts = nullptr;
  size_t cap = std::max<size_t>(n * 2, 4096);
  CK(cudaMallocHost((void**)&ctx->h_counts, cap * sizeof(unsigned)));
  ctx->h_counts_cap = cap;
  return OQPB_OK;
}

// Schwarz matrix on the device: ints_exchange, int2.F90:1582-1737
// mu > 0: bounds of the Erf-attenuated integrals (ints_exchange(..., mu2), int2.F90:678); result in Qout (nshell^2)
int schwarz(oqpb_ctx* ctx, double mu, std::vector<double>& Qout) {
  Cutoffs c{1.0e-15, 1.0e-17, 1.0e-17, 50.0};  // int2.F90:1600-1604
  PairTable T;
  int rc = build_pairtable(ctx, c, T, nullptr);
  if (rc) return rc;
  const int ns = ctx->nshell;
  size_t nent = T.ent.size();
  DevBuf d_q, d_tasks, d_cnt;
  CK(d_q.ensure(nent * sizeof(double)));
  CK(cudaMemsetAsync(d_q.p, 0, nent * sizeof(double), ctx->stream));
  CK(d_cnt.ensure(2 * NL * sizeof(unsigned)));
  std::vector<unsigned> hc(2 * NL, 0);
  const ClassEntry* tab = class_table(ctx->pure_l[2] | (ctx->pure_l[3] << 1));
  for (int pc = 0; pc < NL; ++pc) {
    int n = T.cls_off[pc + 1] - T.cls_off[pc];
    hc[2 * pc] = n;
    hc[2 * pc + 1] = 0;
  }
  CK(cudaMemcpyAsync(d_cnt.p, hc.data(), hc.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
  for (int pc = 0; pc < NL; ++pc) {
    int n = T.cls_off[pc + 1] - T.cls_off[pc];
    if (n == 0) continue;
    std::vector<int2> tasks(n);
    for (int k = 0; k < n; ++k) tasks[k] = make_int2(k, k);
    CK(d_tasks.ensure(n * sizeof(int2)));
    CK(cudaMemcpyAsync(d_tasks.p, tasks.data(), n * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    EriArgs A;
    fill_common_args(ctx, T, pc, pc, A);
    A.tasks = d_tasks.as<int2>();
    A.ntasks = d_cnt.as<unsigned>() + 2 * pc;
    A.counter = d_cnt.as<unsigned>() + 2 * pc + 1;
    A.prim_cutoff = c.pair * c.pair;
    A.cutoff = 0.0;
    A.mu2inv = mu > 0 ? 1.0 / (mu * mu) : 0.0;
    A.mode = MODE_SCHWARZ;
    A.qout = d_q.as<double>() + T.cls_off[pc];
    const ClassEntry& ce = tab[quartet_class(pc_of(pc), pc_of(pc))];
    const bool kown = ce.launch_kown != nullptr && (ctx->use_kown >= 2 || (ctx->use_kown == 1 && ce.kown_default));
    const int qpb = kown ? ce.kown_qpb : ce.qpb;
    int nb = std::min((n + qpb - 1) / qpb, 148 * 16);
    CK((kown ? ce.launch_kown : ce.launch)(A, nb, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // tasks vector lifetime
  }
  std::vector<double> q(nent);
  CK(cudaMemcpyAsync(q.data(), d_q.p, nent * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  Qout.assign((size_t)ns * ns, 0.0);
  for (size_t e = 0; e < nent; ++e) {
    int i = T.ent[e].sa, j = T.ent[e].sb;
    Qout[(size_t)i * ns + j] = Qout[(size_t)j * ns + i] = q[e];
  }
  free_pairtable(T);
  d_q.release(); d_tasks.release(); d_cnt.release();
  return OQPB_OK;
}

struct BuildSpec {
  int mode;            // MODE_SYM / MODE_GEN
  // SYM
  int nmat = 0;
  const double* DJ[MAX_MATS];
  const double* DK[MAX_MATS];
  double* F[MAX_MATS];
  // GEN
  const double* Pgen = nullptr;
  double* Fgen = nullptr;
  int gen_nm = 0, gen_ncoul = 0, gen_nvec = 0;
  int gen_mcount = -1;  // matrices taking the exchange part (default: all gen_nm)
  int gen_xoff = 0;     // ... starting at this matrix
  double cj = 0, ck = 0;
  double digest_flops_per_int = 0;
  bool attenuated = false;  // CAM second pass: attenuated integrals + attenuated Schwarz bounds (ctx->run.att_mu)
};

// int2_twoei (int2.F90:589-923): screening data must already be in d_dsh / d_maxden.
int run_build(oqpb_ctx* ctx, const BuildSpec& S) {
  const PairTable& T = ctx->run;
  const int ns = ctx->nshell;
  const long nent = (long)T.ent.size();
  const double cutoff = ctx->cut.integral;
  if (S.attenuated && !(T.att_mu > 0 && T.Qatt.size() == T.ent.size())) {
    ctx->err = "attenuated pass without oqpb_set_screening_cam";
    return OQPB_ERR_STATE;
  }
  const std::vector<double>& hQ = S.attenuated ? T.Qatt : T.Q;
  const std::vector<double>& hQsuf = S.attenuated ? T.Qsufatt : T.Qsuf;
  const double* dQ = S.attenuated ? T.d_Qatt.as<double>() : T.d_Q.as<double>();
  CK(ctx->d_ok.ensure(nent * sizeof(int)));
  CK(ctx->d_d4.ensure(nent * sizeof(double)));
  k_entry_screen<<<(unsigned)((nent + 255) / 256), 256, 0, ctx->stream>>>(
      nent, T.d_ent.as<PairEntry>(), dQ, ctx->d_dsh.as<double>(), ns,
      ctx->d_maxden.as<unsigned long long>(), cutoff, ctx->d_ok.as<int>(), ctx->d_d4.as<double>(),
      T.d_canon.as<int>(), ctx->have_mask ? ctx->d_mask.as<unsigned char>() : nullptr);
  CK(cudaGetLastError());
  static const bool timing = getenv("OQPB_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = tnow();
  double maxden = 0;
  CK(cudaMemcpyAsync(&maxden, ctx->d_maxden.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const double t_synced = tnow();
  const double bound4 = 4.0 * maxden;

  // ---- plan (cached, see BuildPlan): per list pair, kmax[p] by binary search over the suffix maxima of the ket
  // list's Schwarz bounds; chunk boundaries
  using Chunk = PlanChunk;
  using GraphSlot = oqpb_ctx::GraphSlot;
  const int nr = ctx->nranks, rk = ctx->rank;
  double bound4p = bound4;
  if (bound4 > 0) { int e; std::frexp(bound4, &e); bound4p = std::ldexp(1.0, e); }  // next power of two >= bound4
  BuildPlan& P = ctx->plan[S.attenuated ? 1 : 0];
  DevBuf& d_km = P.d_km;
  int rc;
  double t_planned = tnow(), t_uploaded = t_planned;
  size_t km_count = 0;
  // A cached plan stays usable while its density bound is an UPPER bound that is not too loose: kets beyond the bound
  // index of a larger bound cannot survive a smaller one either, and k_enum applies the exact test.  Within 4 binades the
  // extra candidates cost < 2 % of a build; the default incremental SCF (dD shrinks every iteration) then re-plans every
  // third or fourth iteration instead of every iteration.
  if (!(P.valid && P.gen == ctx->plan_gen && P.bound4 >= bound4p && P.bound4 <= 16.0 * bound4p)) {
    P.valid = false;
    P.chunks.clear(); P.cps.clear(); P.km_off.clear(); P.total_local = 0;
    std::vector<std::vector<int>> kmax_all;  // index by list pair order
    std::vector<double> rank_load(std::max(1, ctx->nranks), 0.0);
    // OQPB_ONLY="a,b": profiling knob, build only the (bra list a, ket list b) launches (list = class * 4 + bucket)
    static const char* only_env = getenv("OQPB_ONLY");
    int only_a = -1, only_b = -1;
    if (only_env) sscanf(only_env, "%d,%d", &only_a, &only_b);
    // Several ranks: every launch carries a fixed tail (the last, partly filled wave), so dealing 1/N of EVERY list pair to
    // every rank makes N times as many small launches: measured 4 % of the per-rank time at N = 8 on (H2O)32.  A list
    // pair is therefore shared by only s = round(estimated ms / whole_ms) ranks (1 <= s <= N: short pairs go to ONE rank as
    // a whole, long ones to all), namely the s ranks with the least estimated load so far; among them the bras are dealt
    // cyclically.  Every rank computes the same assignment (the estimate uses the candidates of every 8th bra of the list,
    // whatever the rank).
    std::vector<int> share_s((size_t)NL * NL, nr), share_j((size_t)NL * NL, rk);  // ranks sharing the pair; my index among them (-1: none)
    if (nr > 1 && ctx->whole_ms > 0) {
      // cost model fitted to the per-(class, contraction bucket pair) profile of (H2O)32/cc-pVTZ (profiles/): SM-ns per
      // quartet = 15 + 0.8 N + prims (2 + 0.35 N), N = Cartesian integrals of the class, prims = primitive quartets that
      // pass the int_rys.F90:232 test (typical value per bucket pair)
      static const double prims_tab[4][4] = {{1, 3.1, 8, 28}, {3.1, 7.4, 18.4, 67}, {8, 18.4, 49, 183}, {28, 67, 183, 745}};
      std::vector<int> order(nr);
      // (pairs in list order: sorting longest-first was measured WORSE, 1.04-1.10 max/mean instead of 1.01-1.02 -- the model's
      // errors are correlated inside a class, and the list order interleaves the classes)
      for (int pca = 0; pca < NL; ++pca) {
        const int na = T.cls_off[pca + 1] - T.cls_off[pca];
        for (int pcb = 0; pcb <= pca && na > 0; ++pcb) {
          const int nb = T.cls_off[pcb + 1] - T.cls_off[pcb];
          if (nb == 0) continue;
          const double* Qa = hQ.data() + T.cls_off[pca];
          const double* Qs = hQsuf.data() + T.cls_off[pcb];
          double est = 0;
          for (int p = 0; p < na; p += 8) {
            int lo = 0, hi = nb;
            while (lo < hi) { int mid = (lo + hi) / 2; if ((Qa[p] * Qs[mid]) * bound4p < cutoff) hi = mid; else lo = mid + 1; }
            est += pca == pcb ? std::min(lo, p + 1) : lo;
          }
          const int pa_ = pc_of(pca), pb_ = pc_of(pcb);
          const double ncart4 = (double)ncart(PC_LA[pa_]) * ncart(PC_LB[pa_]) * ncart(PC_LA[pb_]) * ncart(PC_LB[pb_]);
          const double prims = prims_tab[pca % NBK][pcb % NBK];
          const double est_ms = est * std::min(8, na) * (15.0 + 0.8 * ncart4 + prims * (2.0 + 0.35 * ncart4)) / 148.0 * 1e-6;
          const int sgl = (int)std::min<double>(nr, std::max(1.0, std::floor(est_ms / ctx->whole_ms + 0.5)));
          const size_t id = (size_t)pca * NL + pcb;
          if (sgl >= nr) {
            for (int r = 0; r < nr; ++r) rank_load[r] += est_ms / nr;
            continue;
          }
          for (int r = 0; r < nr; ++r) order[r] = r;
          std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rank_load[a] < rank_load[b]; });
          share_s[id] = sgl;
          share_j[id] = -1;
          for (int j = 0; j < sgl; ++j) {
            rank_load[order[j]] += est_ms / sgl;
            if (order[j] == rk) share_j[id] = j;
          }
        }
      }
    }
    for (int pca = 0; pca < NL; ++pca) {  // pca / pcb are pair LISTS here (class x contraction bucket)
      int na = T.cls_off[pca + 1] - T.cls_off[pca];
      if (na == 0) continue;
      for (int pcb = 0; pcb <= pca; ++pcb) {
        int nb = T.cls_off[pcb + 1] - T.cls_off[pcb];
        if (nb == 0) continue;
        if (only_a >= 0 && (pca != only_a || pcb != only_b)) continue;
        const double* Qa = hQ.data() + T.cls_off[pca];
        std::vector<int> km(na, 0);
        // kets at or beyond km[p] cannot survive: suffix maxima of the ket list's bounds are monotone
        const double* Qs = hQsuf.data() + T.cls_off[pcb];
        const bool diag = pca == pcb;
        auto kbound = [&](int p) {
          int lo = 0, hi = nb;  // first k with Qa[p] * Qs[k] * bound4p < cutoff
          while (lo < hi) {
            int mid = (lo + hi) / 2;
            if ((Qa[p] * Qs[mid]) * bound4p < cutoff) hi = mid; else lo = mid + 1;
          }
          return lo;
        };
        // the ranks that share this list pair and this rank's place among them (see above)
        const int nr_ = share_s[(size_t)pca * NL + pcb], rk_ = share_j[(size_t)pca * NL + pcb];
        if (rk_ < 0) continue;
#pragma omp parallel for schedule(static) if (na > 4096)
        for (int p = rk_; p < na; p += nr_) km[p] = kbound(p);  // this rank's bras only
        // chunking over this rank's bras (p % nranks == rank)
        size_t cand = 0;
        int p0 = 0;
        for (int p = 0; p < na; ++p) {
          if (p % nr_ != rk_) continue;
          P.total_local += diag ? (p + 1) : nb;
          size_t c = diag ? (size_t)std::min(km[p], p + 1) : (size_t)km[p];
          if (cand + c > ctx->task_cap && cand > 0) {
            P.chunks.push_back({pca, pcb, p0, p, cand, nr_, rk_});
            p0 = p;
            cand = 0;
          }
          cand += c;
        }
        if (cand > 0) P.chunks.push_back({pca, pcb, p0, na, cand, nr_, rk_});
        kmax_all.push_back(std::move(km));
        P.cps.push_back({pca, pcb});
      }
    }
    // upload kmax arrays (concatenated)
    std::vector<int> km_cat;
    P.km_off.resize(P.cps.size());
    for (size_t c = 0; c < P.cps.size(); ++c) {
      P.km_off[c] = km_cat.size();
      km_cat.insert(km_cat.end(), kmax_all[c].begin(), kmax_all[c].end());
    }
    km_count = km_cat.size();
    t_planned = tnow();
    if ((rc = upload(ctx, d_km, km_cat))) return rc;
    t_uploaded = tnow();
    P.bound4 = bound4p; P.gen = ctx->plan_gen; P.valid = true;
  }
  const std::vector<Chunk>& chunks = P.chunks;
  const std::vector<std::pair<int, int>>& cps = P.cps;
  const std::vector<size_t>& km_off = P.km_off;
  const long long total_local = P.total_local;
  auto cp_index = [&](int pca, int pcb) {
    for (size_t c = 0; c < cps.size(); ++c) if (cps[c].first == pca && cps[c].second == pcb) return c;
    return (size_t)0;
  };

  size_t nch = chunks.size();
  if ((rc = ensure_counts(ctx, 4 * nch + 4))) return rc;
  CK(ctx->d_counters.ensure((4 * nch + 4) * sizeof(unsigned)));
  unsigned* d_cnt = ctx->d_counters.as<unsigned>();  // [4*c] = ntasks, [4*c+1] = fetch counter, [4*c+2] = run items
  // stream lanes: 4 by default; 8 for builds of 1e5 .. 1e8 candidate quartets (benzene/cc-pVDZ -18 %, n-C20H42 -1 %: their
  // device time is the serial depth of dependent enumerate -> evaluate pairs per lane), measured neutral on (H2O)32 and
  // 12 % slower on the 1.6e3 quartets of H2O/6-31G(d)
  size_t cand_total = 0;
  for (const Chunk& ch : chunks) cand_total += ch.cand;
  const int nlane = ctx->profile || ctx->record ? 1 : (ctx->nlanes > 0 ? ctx->nlanes : (cand_total >= 100000 && cand_total < 100000000 ? 8 : 4));
  // run kernels (one Fock matrix, SYM consumers): warp items = pieces of RUN_LEN kets, at most one short piece per bra
  const bool run_ok = ctx->use_run && S.mode == MODE_SYM && S.nmat == 1 && !ctx->record;
  size_t max_list = 0;
  for (int l = 0; l < NL; ++l) max_list = std::max<size_t>(max_list, T.cls_off[l + 1] - T.cls_off[l]);
  const size_t item_cap = ctx->task_cap / RUN_LEN + max_list + 1;
  for (int l = 0; l < nlane; ++l) {
    CK(ctx->d_tasks[l].ensure(ctx->task_cap * sizeof(int2)));
    if (run_ok) CK(ctx->d_items[l].ensure(item_cap * sizeof(int2)));
  }
  CK(ctx->d_stats.ensure((2 * nch + 2) * sizeof(unsigned long long)));
  std::vector<unsigned long long> h_stats(2 * nch + 2, 0);
  const ClassEntry* tab = class_table(ctx->pure_l[2] | (ctx->pure_l[3] << 1));
  ctx->rec.clear();
  ctx->st_launches = 0;
  std::vector<int2> rec_tmp;
  std::vector<float> chunk_ms;
  const size_t smem_rows = (size_t)2 * ns * sizeof(double);
  const int use_smem = smem_rows <= 80 * 1024;
  const size_t smem_enum_max = (use_smem ? smem_rows : 0) + ENUM_BITS_MAX / 8;  // density rows + one bit per candidate ket
  static DevFlags enum_flags;  // per device
  bool& enum_attr = enum_flags.cur();
  if (smem_enum_max > 48 * 1024 && !enum_attr) {
    CK(cudaFuncSetAttribute(k_enum, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    enum_attr = true;
  }
  // ---- the launch section of a build: counters cleared, then per chunk one enumeration and one ERI launch, on nlane
  // streams forked from / joined into ctx->stream.  Run eagerly, or captured into a CUDA graph and replayed (below).
  auto launch_all = [&]() -> int {
  CK(cudaMemsetAsync(ctx->d_counters.p, 0, (4 * nch + 4) * sizeof(unsigned), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_stats.p, 0, (2 * nch + 2) * sizeof(unsigned long long), ctx->stream));
  if (nlane > 1) {
    CK(cudaEventRecord(ctx->fork_ev, ctx->stream));
    for (int l = 0; l < nlane; ++l) CK(cudaStreamWaitEvent(ctx->lane[l], ctx->fork_ev, 0));
  }
  for (size_t c = 0; c < nch; ++c) {
    const Chunk& ch = chunks[c];
    const int ln = (int)(c % nlane);
    cudaStream_t cs = nlane > 1 ? ctx->lane[ln] : ctx->stream;
    int2* d_tasks = ctx->d_tasks[ln].as<int2>();
    const int qcls = quartet_class(pc_of(ch.pca), pc_of(ch.pcb));
    const ClassEntry& ce = tab[qcls];
    // few, heavily contracted quartets (small molecules): warp per quartet, lanes over the primitive quartets
    const bool wpq = ce.launch_wpq != nullptr && ch.cand <= ctx->wpq_max_tasks && (ch.pca % NBK >= 2 || ch.pcb % NBK >= 2);
    // measured on (H2O)32/cc-pVTZ per contraction bucket pair: the run kernels win 9-10 % where both pairs have few
    // primitives (buckets 0/1), are neutral at (0,2) / (2,0) / (1,1) and lose 7-40 % on the heavily contracted launches
    const bool kown = !wpq && ce.launch_kown != nullptr && (ctx->use_kown >= 2 || (ctx->use_kown == 1 && ce.kown_default));
    const bool run = run_ok && !wpq && !kown && ce.launch_run != nullptr && (ch.pca % NBK) + (ch.pcb % NBK) <= ctx->run_max_bucket_sum;
    int2* d_items = run ? ctx->d_items[ln].as<int2>() : nullptr;
    size_t ci = cp_index(ch.pca, ch.pcb);
    int offa = T.cls_off[ch.pca], offb = T.cls_off[ch.pcb];
    int nbra = (ch.p1 - ch.p0 + ch.nr - 1) / ch.nr + 1;
    // first bra of this rank at or after p0
    int pstart = ch.p0 + ((ch.rk - ch.p0 % ch.nr) % ch.nr + ch.nr) % ch.nr;
    const int bits_cap = std::min(((T.cls_off[ch.pcb + 1] - T.cls_off[ch.pcb]) + 31) & ~31, ENUM_BITS_MAX);
    k_enum<<<nbra, ENUM_NT, (use_smem ? smem_rows : 0) + (size_t)bits_cap / 8, cs>>>(
        T.d_ent.as<PairEntry>() + offa, T.d_ent.as<PairEntry>() + offb, dQ + offa,
        dQ + offb, ctx->d_d4.as<double>() + offa, ctx->d_d4.as<double>() + offb,
        ctx->d_ok.as<int>() + offa, ctx->d_ok.as<int>() + offb, T.d_canon.as<int>() + offa,
        T.d_canon.as<int>() + offb, d_km.as<int>() + km_off[ci], pstart, ch.p1, ch.nr, ch.pca == ch.pcb,
        ctx->d_dsh.as<double>(), ns, cutoff, d_tasks, d_cnt + 4 * c, (unsigned)ctx->task_cap, use_smem, d_items,
        d_cnt + 4 * c + 2, (unsigned)item_cap, bits_cap);
    CK(cudaGetLastError());
    EriArgs A;
    fill_common_args(ctx, T, ch.pca, ch.pcb, A);
    A.tasks = d_tasks;
    A.ntasks = d_cnt + 4 * c;
    A.task_cap = (unsigned)ctx->task_cap;
    A.counter = d_cnt + 4 * c + 1;
    A.items = d_items;
    A.nitems = d_cnt + 4 * c + 2;
    A.item_cap = (unsigned)item_cap;
    A.prim_cutoff = ctx->cut.pair * ctx->cut.pair;
    A.cutoff = cutoff;
    A.mu2inv = S.attenuated ? 1.0 / (T.att_mu * T.att_mu) : 0.0;
    A.stat = ctx->d_stats.as<unsigned long long>() + 2 * c;
    A.mode = S.mode;
    A.nmat = S.nmat;
    for (int m = 0; m < S.nmat; ++m) { A.DJ[m] = S.DJ[m]; A.DK[m] = S.DK[m]; A.F[m] = S.F[m]; }
    A.cj = S.cj; A.ck = S.ck;
    A.Pgen = S.Pgen; A.Fgen = S.Fgen; A.gen_nmat_total = S.gen_nm; A.gen_ncoul = S.gen_ncoul; A.gen_nvec = S.gen_nvec;
    A.gen_mcount = S.gen_mcount >= 0 ? S.gen_mcount : S.gen_nm;
    A.gen_xoff = S.gen_xoff;
    const size_t tasks_per_cta = wpq ? 4 : (kown ? (size_t)ce.kown_qpb : (size_t)ce.qpb);
    // run kernels: a warp = an item; the item count is bounded by candidates / RUN_LEN + one per bra, a warp's mean load by
    // the survivor count: size the grid like the task kernel's (one thread per candidate quartet)
    size_t nb = std::min<size_t>((ch.cand + tasks_per_cta - 1) / tasks_per_cta, (size_t)ce.maxcta * ctx->grid_pct / 100);
    if (run) nb = std::min<size_t>(nb, (ch.cand / RUN_LEN + (size_t)nbra + ce.qpb / 32 - 1) / (ce.qpb / 32) + 1);
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    if (ctx->profile) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); cudaEventRecord(pe0, cs); }
    CK((kown ? ce.launch_kown : (run ? ce.launch_run : (wpq ? ce.launch_wpq : ce.launch)))(A, (int)std::max<size_t>(nb, 1), cs));
    if (ctx->profile) {
      cudaEventRecord(pe1, cs);
      cudaEventSynchronize(pe1);
      float pms = 0;
      cudaEventElapsedTime(&pms, pe0, pe1);
      ctx->prof[qcls][0] += pms;
      chunk_ms.push_back(pms);
      cudaEventDestroy(pe0); cudaEventDestroy(pe1);
    }
    ctx->st_launches += 2;
    if (ctx->record) {
      unsigned n = 0;
      CK(cudaMemcpyAsync(&n, d_cnt + 4 * c, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      rec_tmp.resize(n);
      CK(cudaMemcpy(rec_tmp.data(), d_tasks, n * sizeof(int2), cudaMemcpyDeviceToHost));
      for (unsigned k = 0; k < n; ++k) {
        const PairEntry& eb = T.ent[offa + rec_tmp[k].x];
        const PairEntry& ek = T.ent[offb + rec_tmp[k].y];
        int i = std::max(eb.sa, eb.sb), j = std::min(eb.sa, eb.sb), kk = std::max(ek.sa, ek.sb), l = std::min(ek.sa, ek.sb);
        if ((long)i * (i + 1) / 2 + j < (long)kk * (kk + 1) / 2 + l) { std::swap(i, kk); std::swap(j, l); }
        ctx->rec.push_back(i); ctx->rec.push_back(j); ctx->rec.push_back(kk); ctx->rec.push_back(l);
      }
    }
  }
  if (nlane > 1) {
    for (int l = 0; l < nlane; ++l) {
      CK(cudaEventRecord(ctx->lane_ev[l], ctx->lane[l]));
      CK(cudaStreamWaitEvent(ctx->stream, ctx->lane_ev[l], 0));
    }
  }
  return OQPB_OK;
  };
  // Optional CUDA-graph replay of the cached plan (OQPB_GRAPH=1).  The key holds everything the kernel arguments are made
  // of; the first build with a new key runs eagerly (it also performs the one-time cudaFuncSetAttribute calls), the second
  // is captured, later ones are one cudaGraphLaunch.  Measured NEUTRAL on every configuration (c1 1.40 vs 1.41 ms, c2 6.90
  // vs 6.86, c3 19.2 vs 19.3): the small molecules are not bound by the host's launch rate but by the serial depth of the
  // ~30 dependent enumerate -> evaluate pairs per stream lane, ~40 us each (a few heavily contracted quartets per launch);
  // hence off by default.
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  {
    GraphSlot& G = ctx->graph[S.attenuated ? 1 : 0];
    const bool graph_ok = ctx->use_graph && !ctx->profile && !ctx->record && nch > 0 && nch <= ctx->graph_max_chunks;
    std::string key;
    if (graph_ok) {
      auto put = [&](const void* p_, size_t n_) { key.append(reinterpret_cast<const char*>(p_), n_); };
#define KPUT(x) do { auto v_ = (x); put(&v_, sizeof v_); } while (0)
      KPUT(P.gen); KPUT(P.bound4); KPUT(nch); KPUT(S.mode); KPUT(S.nmat); KPUT(S.cj); KPUT(S.ck); KPUT(S.Pgen); KPUT(S.Fgen);
      KPUT(S.gen_nm); KPUT(S.gen_ncoul); KPUT(S.gen_nvec); KPUT(S.gen_mcount); KPUT(S.gen_xoff); KPUT(S.attenuated);
      for (int m = 0; m < S.nmat; ++m) { KPUT(S.DJ[m]); KPUT(S.DK[m]); KPUT(S.F[m]); }
      KPUT(cutoff); KPUT(ctx->cut.pair); KPUT(T.att_mu); KPUT(ctx->task_cap); KPUT(nlane); KPUT(ctx->use_kown); KPUT(run_ok);
      KPUT(ctx->run_max_bucket_sum); KPUT(ctx->wpq_max_tasks); KPUT(ctx->grid_pct); KPUT(ctx->d_counters.p); KPUT(ctx->d_stats.p);
      KPUT(ctx->d_dsh.p); KPUT(ctx->d_d4.p); KPUT(ctx->d_ok.p); KPUT(d_km.p); KPUT(T.d_ent.p); KPUT(dQ); KPUT(T.d_canon.p);
      for (int l = 0; l < nlane; ++l) { KPUT(ctx->d_tasks[l].p); KPUT(ctx->d_items[l].p); }
#undef KPUT
    }
    if (graph_ok && G.exec && G.key == key) {
      CK(cudaGraphLaunch(G.exec, ctx->stream));
      ctx->st_launches = 2 * (long long)nch;
    } else if (graph_ok && !G.exec && G.key == key) {
      bool captured = false;
      if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        const int lrc = launch_all();
        cudaGraph_t graph = nullptr;
        const cudaError_t ce_ = cudaStreamEndCapture(ctx->stream, &graph);
        if (lrc == OQPB_OK && ce_ == cudaSuccess && graph && cudaGraphInstantiate(&G.exec, graph, 0) == cudaSuccess) captured = true;
        if (graph) cudaGraphDestroy(graph);
      }
      if (captured) {
        CK(cudaGraphLaunch(G.exec, ctx->stream));
      } else {  // capture not possible here: no graphs for this context, run eagerly
        cudaGetLastError();
        G.reset();
        ctx->use_graph = false;
        ctx->st_launches = 0;
        if ((rc = launch_all())) return rc;
      }
    } else {
      if (graph_ok) { G.reset(); G.key = key; }
      if ((rc = launch_all())) return rc;
    }
  }
  const double t_launched = tnow();
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  if (nch) CK(cudaMemcpyAsync(ctx->h_counts, d_cnt, 4 * nch * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
  if (nch) CK(cudaMemcpyAsync(h_stats.data(), ctx->d_stats.p, 2 * nch * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  ctx->st_kernel_ms = ms;
  if (timing)
    fprintf(stderr, "[oqpb timing] wait maxden %.2f ms, plan %.2f, upload kmax (%zu ints) %.2f, launch loop %.2f, drain %.2f; device %.2f ms, %zu chunks\n",
            t_synced - t_begin, t_planned - t_synced, km_count, t_uploaded - t_planned, t_launched - t_uploaded,
            tnow() - t_launched, ms, nch);
  long long surv = 0;
  double flops = 0;
  for (size_t c = 0; c < nch; ++c) {
    unsigned n = ctx->h_counts[4 * c];
    if (n > ctx->task_cap) { ctx->err = "task buffer overflow"; return OQPB_ERR_STATE; }
    surv += n;
    // algorithmic FLOPs: primitive quartets past the int_rys.F90:232 test x F_prim(class)
    // + surviving (de-duplicated) AO integrals x digestion cost per integral
    const Chunk& ch = chunks[c];
    const int pa_ = pc_of(ch.pca), pb_ = pc_of(ch.pcb);
    double fl_c = (double)h_stats[2 * c] * fprim_model(PC_LA[pa_], PC_LB[pa_], PC_LA[pb_], PC_LB[pb_]);
    if (ctx->profile) {
      double* pr = ctx->prof[quartet_class(pa_, pb_)];
      pr[1] += n; pr[2] += (double)h_stats[2 * c]; pr[3] += fl_c + (double)h_stats[2 * c + 1] / 8.0 * S.digest_flops_per_int;
    }
    flops += fl_c;
    flops += (double)h_stats[2 * c + 1] / 8.0 * S.digest_flops_per_int;
  }
  if (ctx->profile && getenv("OQPB_PROF_FILE") && chunk_ms.size() == nch) {
    // per-launch dump: pair list (class*4 + contraction bucket) of bra and ket, ms, quartets, primitive quartets
    if (FILE* fp = fopen(getenv("OQPB_PROF_FILE"), "a")) {
      for (size_t c = 0; c < nch; ++c)
        fprintf(fp, "%d %d %.4f %u %llu\n", chunks[c].pca, chunks[c].pcb, chunk_ms[c], ctx->h_counts[4 * c],
                (unsigned long long)h_stats[2 * c]);
      fclose(fp);
    }
  }
  ctx->st_flops = flops;
  ctx->st_survivors = surv;
  ctx->st_skipped = total_local - surv;
  return OQPB_OK;
}

int check_ready(oqpb_ctx* ctx) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  if (!ctx->have_basis || !ctx->have_cutoff || !ctx->have_screen) {
    ctx->err = "call order: set_basis, set_cutoff, set_screening";
    return OQPB_ERR_STATE;
  }
  return OQPB_OK;
}

oqpb_ctx* g_default_ctx = nullptr;
int g_default_urohf = -1;

// ---- NCCL, loaded on demand: a host without NCCL (or with a single GPU) never needs it
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api(std::string& err) {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("OQPB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      api.CommInitAll = (decltype(api.CommInitAll))dlsym(api.h, "ncclCommInitAll");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
    }
  }
  if (!api.h || !api.CommInitAll || !api.CommDestroy || !api.AllReduce) {
    err = "NCCL not available (libnccl.so.2; set OQPB_NCCL_LIB)";
    return nullptr;
  }
  return &api;
}

// member d of a multi-device context (0 = the master itself)
inline oqpb_ctx* member(oqpb_ctx* ctx, int d) { return d == 0 ? ctx : ctx->peers[d - 1]; }

// run fn(member, d) on one host thread per device (run_build blocks its host thread); first non-zero return wins
template <class F>
int for_each_member(oqpb_ctx* ctx, F&& fn) {
  const int n = ctx->mndev;
  std::vector<int> rc(n, 0);
  std::vector<std::thread> th;
  for (int d = 1; d < n; ++d) th.emplace_back([&, d] { rc[d] = fn(member(ctx, d), d); });
  rc[0] = fn(ctx, 0);
  for (auto& t : th) t.join();
  for (int d = 0; d < n; ++d)
    if (rc[d]) {
      if (d > 0) ctx->err = "device " + std::to_string(member(ctx, d)->device) + ": " + member(ctx, d)->err;
      return rc[d];
    }
  return OQPB_OK;
}

// sum the partial results of the members in place: ONE ncclAllReduce on each member's compute stream
// (replaces pe%allreduce, int2.F90:1392-1397 / parallel.F90:429-440)
int member_allreduce(oqpb_ctx* c, double* buf, size_t count) {
  std::string err;
  NcclApi* api = nccl_api(err);
  if (!api) { c->err = err; return OQPB_ERR_STATE; }
  ncclResult_t r = api->AllReduce(buf, buf, count, ncclDouble, ncclSum, c->comm, c->stream);
  if (r != ncclSuccess) {
    c->err = std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error");
    return OQPB_ERR_CUDA;
  }
  return OQPB_OK;
}

void apply_partition(oqpb_ctx* c) {
  c->rank = c->base_rank * c->mndev + c->mdev;
  c->nranks = c->base_nranks * c->mndev;
  ++c->plan_gen;
}

}  // namespace

// ===================================================================================== C ABI
extern "C" {

int oqpb_ctx_create(oqpb_ctx** out, int device) {
  if (!out) return OQPB_ERR_BAD_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return OQPB_ERR_NO_DEVICE;
  if (device < 0 || device >= ndev) return OQPB_ERR_BAD_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return OQPB_ERR_NO_DEVICE;
  oqpb_ctx* ctx = new oqpb_ctx;
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return OQPB_ERR_CUDA; }
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
  for (int l = 0; l < oqpb_ctx::NSTREAM; ++l) {
    cudaStreamCreateWithFlags(&ctx->lane[l], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->lane_ev[l], cudaEventDisableTiming);
  }
  // tuning knobs (tools/sweep_knobs.sh)
  if (const char* e = getenv("OQPB_NLANES")) ctx->nlanes = std::max(1, std::min((int)oqpb_ctx::NSTREAM, atoi(e)));
  if (const char* e = getenv("OQPB_GRID_PCT")) ctx->grid_pct = std::max(10, atoi(e));
  if (const char* e = getenv("OQPB_RUN")) ctx->use_run = atoi(e) != 0;
  if (const char* e = getenv("OQPB_KOWN")) ctx->use_kown = atoi(e);
  if (const char* e = getenv("OQPB_WHOLE_MS")) ctx->whole_ms = atof(e);
  if (const char* e = getenv("OQPB_GRAPH")) ctx->use_graph = atoi(e) != 0;
  if (const char* e = getenv("OQPB_GRAPH_MAX")) ctx->graph_max_chunks = (size_t)std::max(0, atoi(e));
  if (const char* e = getenv("OQPB_RUN_BUCKETS")) ctx->run_max_bucket_sum = atoi(e);
  if (const char* e = getenv("OQPB_WPQ_MAX")) ctx->wpq_max_tasks = (size_t)std::max(0, atoi(e));
  if (const char* e = getenv("OQPB_TASK_CAP_LOG2")) ctx->task_cap = (size_t)1 << std::max(16, std::min(28, atoi(e)));
  // Rys tables: the generated Chebyshev fits are converted to monomial coefficients (Horner evaluation in the kernels)
  if (ctx->d_rys.ensure(sizeof(RYS_TAB_H) + sizeof(RYSF_TAB_H)) != cudaSuccess) { delete ctx; return OQPB_ERR_CUDA; }
  {
    std::vector<double> mono(RYS_NTAB + RYSF_NTAB);
    cheb_to_monomial(RYS_TAB_H, RYS_NTAB, RYS_NCOEF, mono.data());
    cheb_to_monomial(RYSF_TAB_H, RYSF_NTAB, RYSF_NCOEF, mono.data() + RYS_NTAB);  // nroots 1, 2: fine intervals
    cudaMemcpy(ctx->d_rys.p, mono.data(), mono.size() * sizeof(double), cudaMemcpyHostToDevice);
  }
  ctx->d_maxden.ensure(16);
  *out = ctx;
  return OQPB_OK;
}

// One context driving ndev GPUs of this node from a single process: replicated basis / pair table / Schwarz matrix /
// density, the bra shell-pair list split over the devices (on top of the caller's rank split), partial results summed by
// ONE ncclAllReduce on the compute streams.  Every host-pointer entry (oqpb_fock, oqpb_fock_cam, oqpb_jk_mrsf[_cam],
// routec_fock_jk) then uses all the devices; a non-MPI OpenQP reaches GPUs 1..7 this way.
int oqpb_ctx_create_multi(oqpb_ctx** out, int ndev, const int* devices) {
  if (!out || ndev < 1) return OQPB_ERR_BAD_ARG;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) return OQPB_ERR_NO_DEVICE;
  if (ndev > have) return OQPB_ERR_BAD_ARG;
  std::vector<int> devs(ndev);
  for (int d = 0; d < ndev; ++d) devs[d] = devices ? devices[d] : d;
  std::vector<oqpb_ctx*> c(ndev, nullptr);
  auto cleanup = [&] { for (oqpb_ctx* x : c) if (x) { x->peers.clear(); oqpb_ctx_destroy(x); } };
  for (int d = 0; d < ndev; ++d) {
    int rc = oqpb_ctx_create(&c[d], devs[d]);
    if (rc) { cleanup(); return rc; }
    c[d]->mdev = d;
    c[d]->mndev = ndev;
    apply_partition(c[d]);
  }
  if (ndev > 1) {
    std::string err;
    NcclApi* api = nccl_api(err);
    std::vector<ncclComm_t> comms(ndev, nullptr);
    if (!api || api->CommInitAll(comms.data(), ndev, devs.data()) != ncclSuccess) { cleanup(); return OQPB_ERR_STATE; }
    for (int d = 0; d < ndev; ++d) c[d]->comm = comms[d];
  }
  for (int d = 1; d < ndev; ++d) c[0]->peers.push_back(c[d]);
  cudaSetDevice(devs[0]);
  *out = c[0];
  return OQPB_OK;
}

int oqpb_ctx_ndevices(const oqpb_ctx* ctx) { return ctx ? ctx->mndev : 0; }

void oqpb_ctx_destroy(oqpb_ctx* ctx) {
  if (!ctx) return;
  for (oqpb_ctx* p : ctx->peers) oqpb_ctx_destroy(p);
  ctx->peers.clear();
  cudaSetDevice(ctx->device);
  if (ctx->comm) {
    std::string err;
    if (NcclApi* api = nccl_api(err)) api->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  if (g_default_ctx == ctx) g_default_ctx = nullptr;
  free_pairtable(ctx->run);
  for (DevBuf* b : {&ctx->d_am, &ctx->d_ncontr, &ctx->d_goff, &ctx->d_aooff, &ctx->d_naos, &ctx->d_ex, &ctx->d_cc,
                    &ctx->d_xyz, &ctx->d_rys, &ctx->d_Qmat, &ctx->d_dsh, &ctx->d_maxden, &ctx->d_ok,
                    &ctx->d_d4, &ctx->d_rowsbuf, &ctx->plan[0].d_km, &ctx->plan[1].d_km, &ctx->d_counters, &ctx->d_Dsq, &ctx->d_F, &ctx->d_Din,
                    &ctx->d_stats, &ctx->d_gen_in, &ctx->d_gen_out, &ctx->d_mask})
    b->release();
  ctx->graph[0].reset(); ctx->graph[1].reset();
  for (int l = 0; l < oqpb_ctx::NSTREAM; ++l) { ctx->d_tasks[l].release(); ctx->d_items[l].release(); }
  if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaEventDestroy(ctx->fork_ev);
  for (int l = 0; l < oqpb_ctx::NSTREAM; ++l) { cudaStreamDestroy(ctx->lane[l]); cudaEventDestroy(ctx->lane_ev[l]); }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* oqpb_last_error(const oqpb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int oqpb_set_basis(oqpb_ctx* ctx, int nshell, int nprim, const int* am, const int* harmonic, const int* ncontr,
                   const int* g_offset, const int* ao_offset, const int* naos, const double* ex, const double* cc,
                   const double* centers, int harmonic_active) {
  if (!ctx || nshell <= 0 || nprim <= 0) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  ctx->nshell = nshell; ctx->nprim = nprim; ctx->harmonic_active = harmonic_active;
  ctx->am.assign(am, am + nshell); ctx->harm.assign(harmonic, harmonic + nshell);
  ctx->ncontr.assign(ncontr, ncontr + nshell); ctx->goff.assign(g_offset, g_offset + nshell);
  ctx->aooff.assign(ao_offset, ao_offset + nshell); ctx->naos.assign(naos, naos + nshell);
  ctx->ex.assign(ex, ex + nprim); ctx->cc.assign(cc, cc + nprim); ctx->cen.assign(centers, centers + 3 * nshell);
  ctx->nbf = ao_offset[nshell - 1] + naos[nshell - 1];
  ctx->lmax = 0;
  int flag[4] = {-1, -1, -1, -1};
  for (int s = 0; s < nshell; ++s) {
    int l = am[s];
    if (l < 0 || l > 3) { ctx->err = "angular momentum > f not supported"; return OQPB_ERR_UNSUPPORTED; }
    ctx->lmax = std::max(ctx->lmax, l);
    int pure = (harmonic_active && harmonic[s] == 1 && l >= 2) ? 1 : 0;
    if (flag[l] >= 0 && flag[l] != pure) { ctx->err = "mixed harmonic flags within one angular momentum"; return OQPB_ERR_UNSUPPORTED; }
    flag[l] = pure;
    int expect = pure ? 2 * l + 1 : ncart(l);
    if (naos[s] != expect) { ctx->err = "naos inconsistent with am/harmonic"; return OQPB_ERR_BAD_ARG; }
  }
  for (int l = 0; l < 4; ++l) ctx->pure_l[l] = flag[l] > 0;
  if (ctx->nbf > 46000) { ctx->err = "nbf too large"; return OQPB_ERR_UNSUPPORTED; }
  int rc;
  if ((rc = upload(ctx, ctx->d_am, ctx->am))) return rc;
  if ((rc = upload(ctx, ctx->d_ncontr, ctx->ncontr))) return rc;
  if ((rc = upload(ctx, ctx->d_goff, ctx->goff))) return rc;
  if ((rc = upload(ctx, ctx->d_aooff, ctx->aooff))) return rc;
  if ((rc = upload(ctx, ctx->d_naos, ctx->naos))) return rc;
  if ((rc = upload(ctx, ctx->d_ex, ctx->ex))) return rc;
  if ((rc = upload(ctx, ctx->d_cc, ctx->cc))) return rc;
  if ((rc = upload(ctx, ctx->d_xyz, ctx->cen))) return rc;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_basis = true;
  ctx->have_cutoff = ctx->have_screen = false;
  ++ctx->plan_gen;
  for (oqpb_ctx* pr : ctx->peers) {  // multi-device context: replicate
    rc = oqpb_set_basis(pr, nshell, nprim, am, harmonic, ncontr, g_offset, ao_offset, naos, ex, cc, centers, harmonic_active);
    if (rc) { ctx->err = pr->err; return rc; }
  }
  cudaSetDevice(ctx->device);
  return OQPB_OK;
}

int oqpb_set_cutoff(oqpb_ctx* ctx, double cutoff) {
  if (!ctx || !ctx->have_basis) return OQPB_ERR_STATE;
  for (oqpb_ctx* pr : ctx->peers) {
    int rcp = oqpb_set_cutoff(pr, cutoff);
    if (rcp) { ctx->err = pr->err; return rcp; }
  }
  cudaSetDevice(ctx->device);
  ctx->cutoff = cutoff;
  ctx->cut = Cutoffs{cutoff, 1.0e-2 * cutoff, 1.0e-4 * cutoff, 25.0 * std::log(10.0)};  // int2.F90:260-272
  ctx->have_cutoff = true;
  ++ctx->plan_gen;
  if (ctx->have_screen) {  // re-sort not needed, but the primitive table depends on the cutoffs
    int rc = build_pairtable(ctx, ctx->cut, ctx->run, &ctx->Qmat);
    if (rc) return rc;
  }
  return OQPB_OK;
}

int oqpb_set_screening(oqpb_ctx* ctx, const double* schwarz_in) {
  if (!ctx || !ctx->have_basis || !ctx->have_cutoff) return OQPB_ERR_STATE;
  cudaSetDevice(ctx->device);
  const int ns = ctx->nshell;
  if (schwarz_in) {
    ctx->Qmat.assign(schwarz_in, schwarz_in + (size_t)ns * ns);
  } else {
    int rc = schwarz(ctx, 0.0, ctx->Qmat);
    if (rc) return rc;
  }
  int rc = build_pairtable(ctx, ctx->cut, ctx->run, &ctx->Qmat);
  if (rc) return rc;
  CK(ctx->d_dsh.ensure((size_t)ns * ns * sizeof(double)));
  ctx->have_screen = true;
  ++ctx->plan_gen;
  for (oqpb_ctx* pr : ctx->peers) {  // the Schwarz matrix is computed once and replicated
    rc = oqpb_set_screening(pr, ctx->Qmat.data());
    if (rc) { ctx->err = pr->err; return rc; }
  }
  cudaSetDevice(ctx->device);
  return OQPB_OK;
}

// Schwarz bounds of the Erf-attenuated integrals erf(mu r)/r for the CAM second pass (int2.F90:674-685); the pair
// lists keep the order of the regular bounds, the enumeration's cut-off uses the suffix maxima of the new bounds.
int oqpb_set_screening_cam(oqpb_ctx* ctx, double mu, const double* schwarz_att_in) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!(mu > 0)) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const int ns = ctx->nshell;
  PairTable& T = ctx->run;
  if (T.att_mu == mu && T.Qatt.size() == T.ent.size() && !schwarz_att_in) return OQPB_OK;
  std::vector<double> Qm;
  if (schwarz_att_in) Qm.assign(schwarz_att_in, schwarz_att_in + (size_t)ns * ns);
  else if ((rc = schwarz(ctx, mu, Qm))) return rc;
  ctx->Qmat_att = Qm;
  T.Qatt.resize(T.ent.size());
  for (size_t e = 0; e < T.ent.size(); ++e) T.Qatt[e] = Qm[(size_t)T.ent[e].sa * ns + T.ent[e].sb];
  T.Qsufatt = T.Qatt;
  for (int pc = 0; pc < NL; ++pc)
    for (int k = T.cls_off[pc + 1] - 2; k >= T.cls_off[pc]; --k) T.Qsufatt[k] = std::max(T.Qsufatt[k], T.Qsufatt[k + 1]);
  if ((rc = upload(ctx, T.d_Qatt, T.Qatt))) return rc;
  T.att_mu = mu;
  ctx->plan[1].valid = false;
  for (oqpb_ctx* pr : ctx->peers) {
    rc = oqpb_set_screening_cam(pr, mu, ctx->Qmat_att.data());
    if (rc) { ctx->err = pr->err; return rc; }
  }
  cudaSetDevice(ctx->device);
  return OQPB_OK;
}

int oqpb_get_schwarz_cam(oqpb_ctx* ctx, double* out) {
  if (!ctx || ctx->Qmat_att.empty()) return OQPB_ERR_STATE;
  memcpy(out, ctx->Qmat_att.data(), ctx->Qmat_att.size() * sizeof(double));
  return OQPB_OK;
}

int oqpb_get_schwarz(oqpb_ctx* ctx, double* out) {
  if (!ctx || !ctx->have_screen) return OQPB_ERR_STATE;
  memcpy(out, ctx->Qmat.data(), ctx->Qmat.size() * sizeof(double));
  return OQPB_OK;
}

// Restrict the builds to the quartets whose reference bra pair (the canonically larger shell pair of the quartet,
// int2.F90:756-780) is flagged in mask[i(i+1)/2 + j] (0-based, i >= j); NULL removes the restriction.  Used to
// compare a sample of a large build with the oracle run on the same bra subset.
int oqpb_set_bra_mask(oqpb_ctx* ctx, const unsigned char* mask, long long npairs) {
  if (!ctx || !ctx->have_basis) return OQPB_ERR_STATE;
  cudaSetDevice(ctx->device);
  for (oqpb_ctx* pr : ctx->peers) {
    int rcp = oqpb_set_bra_mask(pr, mask, npairs);
    if (rcp) return rcp;
  }
  cudaSetDevice(ctx->device);
  if (!mask) { ctx->have_mask = false; return OQPB_OK; }
  if (npairs != (long long)ctx->nshell * (ctx->nshell + 1) / 2) return OQPB_ERR_BAD_ARG;
  CK(ctx->d_mask.ensure((size_t)npairs));
  CK(cudaMemcpy(ctx->d_mask.p, mask, (size_t)npairs, cudaMemcpyHostToDevice));
  ctx->have_mask = true;
  return OQPB_OK;
}

int oqpb_set_partition(oqpb_ctx* ctx, int rank, int nranks) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return OQPB_ERR_BAD_ARG;
  // a multi-device context splits the caller's slice once more over its devices
  for (int d = 0; d < ctx->mndev; ++d) {
    oqpb_ctx* c = member(ctx, d);
    c->base_rank = rank;
    c->base_nranks = nranks;
    apply_partition(c);
  }
  return OQPB_OK;
}

// npass = 1: regular build.  npass = 2: int2_run_cam (int2.F90:538-584): pass 1 regular integrals with (sc[0], se[0]),
// pass 2 Erf-attenuated integrals with (sc[1], se[1]), both accumulated into the same Fock matrices.
static int fock_core(oqpb_ctx* ctx, int urohf, const double* d_dev, double* f_dev, int nfocks, int npass, const double* se,
                     const double* sc) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  if (nfocks < 1 || nfocks > MAX_MATS - 1 || (urohf && nfocks != 2)) { ctx->err = "bad nfocks"; return OQPB_ERR_BAD_ARG; }
  const int nbf = ctx->nbf, ns = ctx->nshell;
  const long ntri = (long)nbf * (nbf + 1) / 2, n2 = (long)nbf * nbf;
  const long npairs = (long)ns * (ns + 1) / 2;
  // screening density: shlden (int2.F90:999-1047)
  CK(cudaMemsetAsync(ctx->d_maxden.p, 0, 8, ctx->stream));
  k_shlden_packed<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(
      ns, npairs, ctx->d_aooff.as<int>(), ctx->d_naos.as<int>(), d_dev, nfocks, ntri, ctx->d_dsh.as<double>(),
      ctx->d_maxden.as<unsigned long long>());
  CK(cudaGetLastError());
  // square densities
  int nsq = nfocks + (urohf ? 1 : 0);
  CK(ctx->d_Dsq.ensure((size_t)nsq * n2 * sizeof(double)));
  double* Dsq = ctx->d_Dsq.as<double>();
  for (int m = 0; m < nfocks; ++m)
    k_expand_packed<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(d_dev + (size_t)m * ntri, Dsq + (size_t)m * n2, nbf);
  if (urohf) k_add<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(Dsq, Dsq + n2, Dsq + 2 * n2, n2);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(f_dev, 0, (size_t)nfocks * ntri * sizeof(double), ctx->stream));
  double kernel_ms = 0, flops = 0;
  long long surv = 0, launches = 0;
  for (int pass = 0; pass < npass; ++pass) {
    BuildSpec S;
    S.mode = MODE_SYM;
    S.nmat = nfocks;
    for (int m = 0; m < nfocks; ++m) {
      S.DJ[m] = urohf ? Dsq + 2 * n2 : Dsq + (size_t)m * n2;
      S.DK[m] = Dsq + (size_t)m * n2;
      S.F[m] = f_dev + (size_t)m * ntri;
    }
    S.cj = sc[pass];                          // 4*sc applied in the kernel (xval4, int2.F90:1423 / 1499)
    S.ck = urohf ? 2.0 * se[pass] : se[pass];  // xval1 (int2.F90:1422) / xval2 (int2.F90:1498)
    S.digest_flops_per_int = urohf ? 26.0 : 14.0 * nfocks;
    S.attenuated = pass == 1;
    if ((rc = run_build(ctx, S))) return rc;
    kernel_ms += ctx->st_kernel_ms; flops += ctx->st_flops; surv += ctx->st_survivors; launches += ctx->st_launches;
  }
  // `skipped` stays what the last run_generic left (the reference overwrites it per pass); the rest is summed
  ctx->st_kernel_ms = kernel_ms; ctx->st_flops = flops; ctx->st_survivors = surv; ctx->st_launches = launches;
  return OQPB_OK;
}

int oqpb_fock_dev(oqpb_ctx* ctx, int urohf, const double* d_dev, double* f_dev, int nfocks, double se, double sc) {
  return fock_core(ctx, urohf, d_dev, f_dev, nfocks, 1, &se, &sc);
}

int oqpb_fock_cam_dev(oqpb_ctx* ctx, int urohf, const double* d_dev, double* f_dev, int nfocks, double alpha, double beta,
                      double mu, double alpha_coulomb, double beta_coulomb) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if ((rc = oqpb_set_screening_cam(ctx, mu, nullptr))) return rc;  // no-op when the bounds for this mu are cached
  const double se[2] = {alpha, beta}, sc[2] = {alpha_coulomb, beta_coulomb};
  return fock_core(ctx, urohf, d_dev, f_dev, nfocks, 2, se, sc);
}

int oqpb_fock_post_dev(oqpb_ctx* ctx, double* f_dev, int nfocks) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  const long ntri = (long)ctx->nbf * (ctx->nbf + 1) / 2;
  k_fock_post<<<(unsigned)((ntri * nfocks + 255) / 256), 256, 0, ctx->stream>>>(f_dev, ctx->nbf, ntri, nfocks);
  CK(cudaGetLastError());
  return OQPB_OK;
}

int oqpb_synchronize(oqpb_ctx* ctx) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  return OQPB_OK;
}
void* oqpb_stream(oqpb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int oqpb_set_stream(oqpb_ctx* ctx, void* stream) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)stream;
  ctx->own_stream = false;
  return OQPB_OK;
}

// Host-pointer Fock build on every member of the context: H2D of the packed densities, the member's slice of the
// build, (multi-device) ONE ncclAllReduce of the packed partial Fock matrices on the compute streams, post-scaling
// and D2H on the master.
static int fock_host(oqpb_ctx* ctx, int urohf, const double* d, double* f, int nfocks, int npass, const double* se,
                     const double* sc, double mu, int post, long long* nskipped) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (npass == 2 && (rc = oqpb_set_screening_cam(ctx, mu, nullptr))) return rc;  // no-op when cached for this mu
  const long ntri = (long)ctx->nbf * (ctx->nbf + 1) / 2;
  const size_t bytes = (size_t)nfocks * ntri * sizeof(double);
  auto body = [&](oqpb_ctx* c, int dev) -> int {
    oqpb_ctx* ctx = c;  // CK reports into the member
    cudaSetDevice(c->device);
    CK(c->d_Din.ensure(bytes));
    CK(c->d_F.ensure(bytes));
    CK(cudaMemcpyAsync(c->d_Din.p, d, bytes, cudaMemcpyHostToDevice, c->stream));
    int r = fock_core(c, urohf, c->d_Din.as<double>(), c->d_F.as<double>(), nfocks, npass, se, sc);
    if (r) return r;
    if (c->mndev > 1 && (r = member_allreduce(c, c->d_F.as<double>(), (size_t)nfocks * ntri))) return r;
    if (dev == 0) {
      if (post && (r = oqpb_fock_post_dev(c, c->d_F.as<double>(), nfocks))) return r;
      CK(cudaMemcpyAsync(f, c->d_F.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return OQPB_OK;
  };
  rc = ctx->mndev > 1 ? for_each_member(ctx, body) : body(ctx, 0);
  if (rc) return rc;
  for (oqpb_ctx* pr : ctx->peers) {  // whole-context statistics
    ctx->st_survivors += pr->st_survivors; ctx->st_skipped += pr->st_skipped; ctx->st_flops += pr->st_flops;
    ctx->st_launches += pr->st_launches; ctx->st_kernel_ms = std::max(ctx->st_kernel_ms, pr->st_kernel_ms);
  }
  cudaSetDevice(ctx->device);
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

int oqpb_fock(oqpb_ctx* ctx, int urohf, const double* d, double* f, int nfocks, double se, double sc, int post,
              long long* nskipped) {
  return fock_host(ctx, urohf, d, f, nfocks, 1, &se, &sc, 0.0, post, nskipped);
}

int oqpb_fock_cam(oqpb_ctx* ctx, int urohf, const double* d, double* f, int nfocks, double alpha, double beta, double mu,
                  double alpha_coulomb, double beta_coulomb, int post, long long* nskipped) {
  const double se[2] = {alpha, beta}, sc[2] = {alpha_coulomb, beta_coulomb};
  return fock_host(ctx, urohf, d, f, nfocks, 2, se, sc, mu, post, nskipped);
}

// shared driver for the general-density consumers: X interleaved [(nu*nbf+mu)*NM + m]
static int gen_build(oqpb_ctx* ctx, int NM, int ncoul, int nvec, double cj, double ck, bool attenuated = false) {
  const int nbf = ctx->nbf;
  if (!attenuated) CK(cudaMemsetAsync(ctx->d_gen_out.p, 0, (size_t)nbf * nbf * NM * sizeof(double), ctx->stream));
  BuildSpec S;
  S.mode = MODE_GEN;
  S.attenuated = attenuated;  // CAM second pass: accumulates on top of the first
  S.Pgen = ctx->d_gen_in.as<double>();
  S.Fgen = ctx->d_gen_out.as<double>();
  S.gen_nm = NM; S.gen_ncoul = ncoul; S.gen_nvec = nvec;
  S.cj = cj; S.ck = ck;
  S.digest_flops_per_int = (ncoul > 0 ? 144.0 / 7.0 : 16.0) * NM;  // MRSF: (4*4+8*7)*2 per vector; TD: 24 per vector
  return run_build(ctx, S);
}

// npass = 2: int2_run_cam with the TD consumer -- the same update in both passes with the pass's scale factors
static int td_core(oqpb_ctx* ctx, const double* d2, int nvec, int flags, int npass, const double* sev, const double* scv,
                   double mu, double* apb, double* amb, long long* nskipped) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  if (npass == 2 && (rc = oqpb_set_screening_cam(ctx, mu, nullptr))) return rc;
  const double se = sev[0], sc = scv[0];
  if (nvec < 1) return OQPB_ERR_BAD_ARG;
  // multi-device context: the TD consumer is not split over the devices yet -- the master does this rank's whole slice
  struct WholeSlice {
    oqpb_ctx* c;
    explicit WholeSlice(oqpb_ctx* x) : c(x) { if (c->mndev > 1) { c->rank = c->base_rank; c->nranks = c->base_nranks; ++c->plan_gen; } }
    ~WholeSlice() { if (c->mndev > 1) apply_partition(c); }
  } whole(ctx);
  const int ns = ctx->nshell, nbf = ctx->nbf;
  const long n2 = (long)nbf * nbf, npairs = (long)ns * (ns + 1) / 2;
  const bool tda = flags & OQPB_TD_TDA;
  const bool want_apb = !tda && (flags & OQPB_TD_APB), want_amb = tda || (flags & OQPB_TD_AMB);
  size_t bytes = (size_t)n2 * nvec * sizeof(double);
  CK(ctx->d_Din.ensure(bytes));
  CK(cudaMemcpyAsync(ctx->d_Din.p, d2, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // screening density from the caller's d2 itself: shltd (tdhf_lib.F90:300-325); d2 (mu,nu,v) has v slowest,
  // so screen on a plain interleaved copy first
  int ncomp = tda ? 1 : ((want_apb ? 1 : 0) + (want_amb ? 1 : 0));
  if (ncomp == 0) return OQPB_ERR_BAD_ARG;
  int NM = nvec * (tda ? 1 : 2);
  CK(ctx->d_gen_in.ensure((size_t)n2 * NM * sizeof(double)));
  CK(ctx->d_gen_out.ensure((size_t)n2 * NM * sizeof(double)));
  unsigned gb = (unsigned)((n2 * nvec + 255) / 256);
  // plain copy for screening (stored in d_gen_out temporarily)
  k_td_pack<<<gb, 256, 0, ctx->stream>>>(ctx->d_Din.as<double>(), ctx->d_gen_out.as<double>(), nbf, nvec, 1, 0);
  CK(cudaMemsetAsync(ctx->d_maxden.p, 0, 8, ctx->stream));
  k_shlden_gen<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(
      ns, npairs, ctx->d_aooff.as<int>(), ctx->d_naos.as<int>(), ctx->d_gen_out.as<double>(), nvec, nbf,
      ctx->d_dsh.as<double>(), ctx->d_maxden.as<unsigned long long>());
  CK(cudaGetLastError());
  if (tda) {
    // amb = -se K[P] (+ 2 sc J[P+P^T] on 4 targets): tdhf_lib.F90:169-186
    k_td_pack<<<gb, 256, 0, ctx->stream>>>(ctx->d_Din.as<double>(), ctx->d_gen_in.as<double>(), nbf, nvec, 1, 0);
    rc = gen_build(ctx, NM, (flags & OQPB_TD_TDA_COULOMB) ? 1 : 0, nvec, 2.0 * sc, se);
    if (rc) return rc;
    if (npass == 2 && (rc = gen_build(ctx, NM, (flags & OQPB_TD_TDA_COULOMB) ? 1 : 0, nvec, 2.0 * scv[1], sev[1], true))) return rc;
    k_td_unpack<<<gb, 256, 0, ctx->stream>>>(ctx->d_gen_out.as<double>(), ctx->d_Din.as<double>(), nbf, nvec, NM, 0);
    CK(cudaMemcpyAsync(amb, ctx->d_Din.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  } else {
    // comp 0: Ps = P+P^T -> apb (Coulomb 4 sc, exchange se, symmetrised); comp 1: Pa = P-P^T -> amb (exchange only)
    k_td_pack<<<gb, 256, 0, ctx->stream>>>(ctx->d_Din.as<double>(), ctx->d_gen_in.as<double>(), nbf, nvec, 2, 1);
    rc = gen_build(ctx, NM, want_apb ? 1 : 0, nvec, 2.0 * sc, se);
    if (rc) return rc;
    if (npass == 2 && (rc = gen_build(ctx, NM, want_apb ? 1 : 0, nvec, 2.0 * scv[1], sev[1], true))) return rc;
    if (want_apb) {
      k_td_unpack<<<gb, 256, 0, ctx->stream>>>(ctx->d_gen_out.as<double>(), ctx->d_Din.as<double>(), nbf, nvec, NM, 0);
      CK(cudaMemcpyAsync(apb, ctx->d_Din.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    if (want_amb) {
      k_td_unpack<<<gb, 256, 0, ctx->stream>>>(ctx->d_gen_out.as<double>(), ctx->d_Din.as<double>(), nbf, nvec, NM, 1);
      CK(cudaMemcpyAsync(amb, ctx->d_Din.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

int oqpb_jk_td(oqpb_ctx* ctx, const double* d2, int nvec, int flags, double se, double sc, double* apb, double* amb,
               long long* nskipped) {
  return td_core(ctx, d2, nvec, flags, 1, &se, &sc, 0.0, apb, amb, nskipped);
}
int oqpb_jk_td_cam(oqpb_ctx* ctx, const double* d2, int nvec, int flags, double alpha, double beta, double mu,
                   double alpha_coulomb, double beta_coulomb, double* apb, double* amb, long long* nskipped) {
  const double se[2] = {alpha, beta}, sc[2] = {alpha_coulomb, beta_coulomb};
  return td_core(ctx, d2, nvec, flags, 2, se, sc, mu, apb, amb, nskipped);
}

// d3 / f3 on the device (interleaved d3(v, c, mu, nu), v fastest).  npass = 2: int2_run_cam, pass 2 = attenuated
// integrals, exchange of component 7 only (tdhf_mrsf_lib.F90:312-326)
static int mrsf_core(oqpb_ctx* ctx, const double* d3_dev, double* f3_dev, int nvec, int ncomp, double se, double sc,
                     int npass = 1, double se2 = 0.0, double mu = 0.0) {
  if (npass == 2) {
    if (ncomp < 7) return OQPB_ERR_BAD_ARG;
    int rc0 = oqpb_set_screening_cam(ctx, mu, nullptr);
    if (rc0) return rc0;
  }
  const int ns = ctx->nshell, nbf = ctx->nbf;
  const long npairs = (long)ns * (ns + 1) / 2;
  const int NM = nvec * ncomp;
  CK(cudaMemsetAsync(ctx->d_maxden.p, 0, 8, ctx->stream));
  k_shlden_gen<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(
      ns, npairs, ctx->d_aooff.as<int>(), ctx->d_naos.as<int>(), d3_dev, NM, nbf, ctx->d_dsh.as<double>(),
      ctx->d_maxden.as<unsigned long long>());
  CK(cudaGetLastError());
  // Coulomb: cval * ds with ds = d3 + d3^T on (i,j),(j,i),(k,l),(l,k)  -> cj = sc; exchange xval -> ck = se
  CK(cudaMemsetAsync(f3_dev, 0, (size_t)nbf * nbf * NM * sizeof(double), ctx->stream));
  BuildSpec S;
  S.mode = MODE_GEN;
  S.Pgen = d3_dev;
  S.Fgen = f3_dev;
  S.gen_nm = NM; S.gen_ncoul = 4; S.gen_nvec = nvec;
  S.cj = sc; S.ck = se;
  S.digest_flops_per_int = 144.0 / 7.0 * NM;  // (4*4 + 8*7) * 2 flops per integral and vector
  int rc = run_build(ctx, S);
  if (rc || npass == 1) return rc;
  const double kernel_ms = ctx->st_kernel_ms, flops = ctx->st_flops;
  const long long surv = ctx->st_survivors, launches = ctx->st_launches;
  S.attenuated = true;
  S.Pgen = d3_dev + (size_t)6 * nvec;  // component 7 (m = 6 nvec .. 7 nvec - 1), same AO-pair stride NM
  S.Fgen = f3_dev + (size_t)6 * nvec;
  S.gen_ncoul = 0;
  S.gen_mcount = nvec;
  S.cj = 0.0; S.ck = se2;
  S.digest_flops_per_int = 16.0 * nvec;
  if ((rc = run_build(ctx, S))) return rc;
  ctx->st_kernel_ms += kernel_ms; ctx->st_flops += flops; ctx->st_survivors += surv; ctx->st_launches += launches;
  return OQPB_OK;
}

int oqpb_jk_mrsf_dev(oqpb_ctx* ctx, const double* d3_dev, int nvec, int ncomp, double se, double sc, double* f3_dev) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  if (nvec < 1 || ncomp < 4 || !d3_dev || !f3_dev) return OQPB_ERR_BAD_ARG;
  return mrsf_core(ctx, d3_dev, f3_dev, nvec, ncomp, se, sc);
}

int oqpb_jk_mrsf_cam_dev(oqpb_ctx* ctx, const double* d3_dev, int nvec, int ncomp, double alpha, double beta, double mu,
                         double alpha_coulomb, double* f3_dev) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  if (nvec < 1 || ncomp < 7 || !d3_dev || !f3_dev) return OQPB_ERR_BAD_ARG;
  return mrsf_core(ctx, d3_dev, f3_dev, nvec, ncomp, alpha, alpha_coulomb, 2, beta, mu);
}

static int mrsf_host(oqpb_ctx* ctx, const double* d3, int nvec, int ncomp, double se, double sc, double* f3,
                     long long* nskipped, int npass, double se2, double mu) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (nvec < 1 || ncomp < 4) return OQPB_ERR_BAD_ARG;
  if (npass == 2 && (rc = oqpb_set_screening_cam(ctx, mu, nullptr))) return rc;
  const long n2 = (long)ctx->nbf * ctx->nbf;
  const int NM = nvec * ncomp;
  const size_t bytes = (size_t)n2 * NM * sizeof(double);
  auto body = [&](oqpb_ctx* c, int dev) -> int {
    oqpb_ctx* ctx = c;
    cudaSetDevice(c->device);
    CK(c->d_gen_in.ensure(bytes));
    CK(c->d_gen_out.ensure(bytes));
    CK(cudaMemcpyAsync(c->d_gen_in.p, d3, bytes, cudaMemcpyHostToDevice, c->stream));
    int r = mrsf_core(c, c->d_gen_in.as<double>(), c->d_gen_out.as<double>(), nvec, ncomp, se, sc, npass, se2, mu);
    if (r) return r;
    if (c->mndev > 1 && (r = member_allreduce(c, c->d_gen_out.as<double>(), (size_t)n2 * NM))) return r;
    if (dev == 0) CK(cudaMemcpyAsync(f3, c->d_gen_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OQPB_OK;
  };
  rc = ctx->mndev > 1 ? for_each_member(ctx, body) : body(ctx, 0);
  if (rc) return rc;
  for (oqpb_ctx* pr : ctx->peers) {
    ctx->st_survivors += pr->st_survivors; ctx->st_skipped += pr->st_skipped; ctx->st_flops += pr->st_flops;
    ctx->st_launches += pr->st_launches; ctx->st_kernel_ms = std::max(ctx->st_kernel_ms, pr->st_kernel_ms);
  }
  cudaSetDevice(ctx->device);
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

int oqpb_jk_mrsf(oqpb_ctx* ctx, const double* d3, int nvec, int ncomp, double se, double sc, double* f3,
                 long long* nskipped) {
  return mrsf_host(ctx, d3, nvec, ncomp, se, sc, f3, nskipped, 1, 0.0, 0.0);
}
int oqpb_jk_mrsf_cam(oqpb_ctx* ctx, const double* d3, int nvec, int ncomp, double alpha, double beta, double mu,
                     double alpha_coulomb, double* f3, long long* nskipped) {
  return mrsf_host(ctx, d3, nvec, ncomp, alpha, alpha_coulomb, f3, nskipped, 2, beta, mu);
}

// ---- generic J/K engine on the GEN path (oqpb_jk) and the response / gradient consumers built on it ----------
namespace {
struct Slab { const double* src; int op; };  // host column-major nbf x nbf matrix and the k_slab_pack operation
// Upload `mats` (host, nbf^2 each), screen on them exactly like shltd / shlrpagrd (tdhf_lib.F90:300-325, 1324-1350:
// dsh(I,J) = max |P(mu in J, nu in I)| over all matrices, I >= J), then build  X_m = sum of its slabs'  op(P)  and run ONE
// GEN build:  Fgen_m = cj * J[X_m] (m < ncoul)  -  ck * K[X_m] (m >= ncoul).   J[P](a,b) = sum_cd (ab|cd) P(c,d),
// K[P](a,c) = sum_bd (ab|cd) P(b,d).  Results stay in ctx->d_gen_out (interleaved, NM slabs).
int jk_slabs(oqpb_ctx* ctx, const std::vector<const double*>& mats, const std::vector<std::vector<Slab>>& jslabs,
             const std::vector<std::vector<Slab>>& kslabs, double cj, double ck, double flops_per_int) {
  const int nbf = ctx->nbf, ns = ctx->nshell;
  const long n2 = (long)nbf * nbf, npairs = (long)ns * (ns + 1) / 2;
  const int nin = (int)mats.size(), nJ = (int)jslabs.size(), nK = (int)kslabs.size(), NM = nJ + nK;
  if (nin < 1 || NM < 1) return OQPB_ERR_BAD_ARG;
  const unsigned gb = (unsigned)((n2 + 255) / 256);
  CK(ctx->d_Din.ensure((size_t)n2 * nin * sizeof(double)));
  CK(ctx->d_gen_in.ensure((size_t)n2 * std::max(NM, nin) * sizeof(double)));
  CK(ctx->d_gen_out.ensure((size_t)n2 * std::max(NM, nin) * sizeof(double)));
  std::vector<const double*> dev(nin);
  for (int q = 0; q < nin; ++q) {
    dev[q] = ctx->d_Din.as<double>() + (size_t)q * n2;
    CK(cudaMemcpyAsync((void*)dev[q], mats[q], (size_t)n2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_slab_pack<<<gb, 256, 0, ctx->stream>>>(dev[q], ctx->d_gen_out.as<double>(), nbf, nin, q, 0, 0);  // plain copy for screening
  }
  CK(cudaMemsetAsync(ctx->d_maxden.p, 0, 8, ctx->stream));
  k_shlden_gen<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(
      ns, npairs, ctx->d_aooff.as<int>(), ctx->d_naos.as<int>(), ctx->d_gen_out.as<double>(), nin, nbf, ctx->d_dsh.as<double>(),
      ctx->d_maxden.as<unsigned long long>());
  CK(cudaGetLastError());
  auto index_of = [&](const double* p) { for (int q = 0; q < nin; ++q) if (mats[q] == p) return q; return -1; };
  for (int m = 0; m < NM; ++m) {
    const std::vector<Slab>& sl = m < nJ ? jslabs[m] : kslabs[m - nJ];
    for (size_t t = 0; t < sl.size(); ++t) {
      int q = index_of(sl[t].src);
      if (q < 0) return OQPB_ERR_BAD_ARG;
      k_slab_pack<<<gb, 256, 0, ctx->stream>>>(dev[q], ctx->d_gen_in.as<double>(), nbf, NM, m, sl[t].op, t > 0);
    }
  }
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(ctx->d_gen_out.p, 0, (size_t)n2 * NM * sizeof(double), ctx->stream));
  BuildSpec S;
  S.mode = MODE_GEN;
  S.Pgen = ctx->d_gen_in.as<double>();
  S.Fgen = ctx->d_gen_out.as<double>();
  S.gen_nm = NM; S.gen_ncoul = nJ; S.gen_nvec = 1; S.gen_xoff = nJ; S.gen_mcount = nK;
  S.cj = nJ > 0 ? cj : 0.0; S.ck = nK > 0 ? ck : 0.0;
  S.digest_flops_per_int = flops_per_int;
  return run_build(ctx, S);
}
// out (host) = sum_t scale_t * slab m_t of d_gen_out
int jk_fetch(oqpb_ctx* ctx, int NM, const std::vector<std::pair<int, double>>& terms, double* out) {
  const int nbf = ctx->nbf;
  const long n2 = (long)nbf * nbf;
  CK(ctx->d_F.ensure((size_t)n2 * sizeof(double)));
  for (size_t t = 0; t < terms.size(); ++t)
    k_slab_unpack<<<(unsigned)((n2 + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_gen_out.as<double>(), ctx->d_F.as<double>(), nbf, NM,
                                                                         terms[t].first, terms[t].second, t > 0);
  if (terms.empty()) CK(cudaMemsetAsync(ctx->d_F.p, 0, (size_t)n2 * sizeof(double), ctx->stream));
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, ctx->d_F.p, (size_t)n2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return OQPB_OK;
}
// a multi-device context runs these consumers on its first device over this rank's whole slice
struct WholeSliceGuard {
  oqpb_ctx* c;
  explicit WholeSliceGuard(oqpb_ctx* x) : c(x) { if (c->mndev > 1) { c->rank = c->base_rank; c->nranks = c->base_nranks; ++c->plan_gen; } }
  ~WholeSliceGuard() { if (c->mndev > 1) apply_partition(c); }
};
}  // namespace

// Generic J/K for n general (non-symmetric) AO matrices P_m, column-major (nbf, nbf, n):
//   J_m(a,b) = sum_cd (ab|cd) P_m(c,d)  if want_j[m],   K_m(a,c) = sum_bd (ab|cd) P_m(b,d)  if want_k[m]
// (slabs of J / K that are not wanted are left untouched).  Screening density = max |P_m| per shell block over all m, the
// rule of shltd (tdhf_lib.F90:300-325).  Every J/K consumer of the reference is a linear combination of these.
int oqpb_jk(oqpb_ctx* ctx, int n, const double* P, const int* want_j, const int* want_k, double* J, double* K,
            long long* nskipped) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (n < 1 || !P || !want_j || !want_k) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  WholeSliceGuard whole(ctx);
  const size_t n2 = (size_t)ctx->nbf * ctx->nbf;
  std::vector<const double*> mats(n);
  std::vector<std::vector<Slab>> js, ks;
  std::vector<int> jm, km;
  for (int m = 0; m < n; ++m) {
    mats[m] = P + m * n2;
    if (want_j[m]) { js.push_back({Slab{mats[m], 0}}); jm.push_back(m); }
    if (want_k[m]) { ks.push_back({Slab{mats[m], 0}}); km.push_back(m); }
  }
  if (js.empty() && ks.empty()) return OQPB_ERR_BAD_ARG;
  if ((rc = jk_slabs(ctx, mats, js, ks, 1.0, 1.0, 4.0 * js.size() + 16.0 * ks.size()))) return rc;
  const int NM = (int)(js.size() + ks.size());
  for (size_t t = 0; t < jm.size(); ++t)
    if ((rc = jk_fetch(ctx, NM, {{(int)t, 1.0}}, J + jm[t] * n2))) return rc;
  for (size_t t = 0; t < km.size(); ++t)
    if ((rc = jk_fetch(ctx, NM, {{(int)(js.size() + t), -1.0}}, K + km[t] * n2))) return rc;
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

// int2_tdgrd_data_t (tdhf_lib.F90:33-36, update :228-295, stop-time symmetrisation :107-109): two spin blocks
// d2(nbf,nbf,2);  apb_s = 2 sc J[P_1 + P_2] - se K[P_s + P_s^T]  (s = 1, 2),  amb_1 = se K[P_1^T - P_1],  amb_2 = 0.
int oqpb_jk_tdgrd(oqpb_ctx* ctx, const double* d2, int flags, double se, double sc, double* apb, double* amb,
                  long long* nskipped) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!d2) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  WholeSliceGuard whole(ctx);
  const size_t n2 = (size_t)ctx->nbf * ctx->nbf;
  const bool want_apb = flags & OQPB_TD_APB, want_amb = flags & OQPB_TD_AMB;
  if (!want_apb && !want_amb) return OQPB_ERR_BAD_ARG;
  const double *P1 = d2, *P2 = d2 + n2;
  std::vector<std::vector<Slab>> js, ks;
  if (want_apb) {
    js.push_back({Slab{P1, 0}, Slab{P2, 0}});
    ks.push_back({Slab{P1, 1}});
    ks.push_back({Slab{P2, 1}});
  }
  if (want_amb) ks.push_back({Slab{P1, 2}});
  if ((rc = jk_slabs(ctx, {P1, P2}, js, ks, 1.0, 1.0, 4.0 * js.size() + 16.0 * ks.size()))) return rc;
  const int NM = (int)(js.size() + ks.size());
  if (want_apb && apb) {
    if ((rc = jk_fetch(ctx, NM, {{0, 2.0 * sc}, {1, se}}, apb))) return rc;      // slabs hold +J and -K
    if ((rc = jk_fetch(ctx, NM, {{0, 2.0 * sc}, {2, se}}, apb + n2))) return rc;
  }
  if (amb) {
    if (want_amb) { if ((rc = jk_fetch(ctx, NM, {{NM - 1, -se}}, amb))) return rc; }
    else memset(amb, 0, n2 * sizeof(double));
    memset(amb + n2, 0, n2 * sizeof(double));
  }
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

// int2_rpagrd_data_t (tdhf_lib.F90:42-57, 1068-1320): xpy, t -> H+ ; xmy -> H- ; arrays (nbf, nbf, nspin, n) column-major.
//   nspin = 1:  H+[V] = 4 sc J[V] - 2 se K[V]           (V symmetric, as every caller passes it: the reference reads only
//                                                        one triangle of V in the Coulomb term and is order-dependent otherwise)
//   nspin = 2:  H+[V]_s = 2 sc J[V_1 + V_2] - se K[V_s + V_s^T]
//   H-[V] = se K[V_1^T - V_1]  (first spin block only, as written :1297-1318; the other block stays zero)
int oqpb_jk_rpagrd(oqpb_ctx* ctx, int nspin, int np, int nm, int nt, const double* xpy, const double* xmy, const double* t,
                   double se, double sc, double* hpp, double* hpt, double* hmm, long long* nskipped) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if ((nspin != 1 && nspin != 2) || np < 0 || nm < 0 || nt < 0 || np + nm + nt == 0) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  WholeSliceGuard whole(ctx);
  const size_t n2 = (size_t)ctx->nbf * ctx->nbf, slab = n2 * nspin;
  std::vector<const double*> mats;
  std::vector<std::vector<Slab>> js, ks;
  struct Out { double* dst; std::vector<std::pair<int, double>> terms; };
  std::vector<Out> outs;
  std::vector<std::pair<int, int>> pending;  // (index into outs, k-slab ordinal) resolved once nJ is known
  std::vector<int> pend_j;
  auto add_plus = [&](const double* v, int cnt, double* dst) {
    for (int q = 0; q < cnt; ++q) {
      const double* v1 = v + q * slab;
      for (int s2 = 0; s2 < nspin; ++s2) mats.push_back(v1 + s2 * n2);
      if (nspin == 1) {
        js.push_back({Slab{v1, 0}});
        ks.push_back({Slab{v1, 0}});
        outs.push_back({dst + q * slab, {{(int)js.size() - 1, 4.0 * sc}}});
        pending.push_back({(int)outs.size() - 1, (int)ks.size() - 1});
        pend_j.push_back(2);  // coefficient code: -2 se K  ->  slab holds -K: + 2 se
      } else {
        js.push_back({Slab{v1, 0}, Slab{v1 + n2, 0}});
        for (int s2 = 0; s2 < 2; ++s2) {
          ks.push_back({Slab{v1 + s2 * n2, 1}});
          outs.push_back({dst + q * slab + s2 * n2, {{(int)js.size() - 1, 2.0 * sc}}});
          pending.push_back({(int)outs.size() - 1, (int)ks.size() - 1});
          pend_j.push_back(1);  // + se * slab
        }
      }
    }
  };
  add_plus(xpy, np, hpp);
  add_plus(t, nt, hpt);
  for (int q = 0; q < nm; ++q) {
    const double* v1 = xmy + q * slab;
    for (int s2 = 0; s2 < nspin; ++s2) mats.push_back(v1 + s2 * n2);
    ks.push_back({Slab{v1, 2}});
    outs.push_back({hmm + q * slab, {}});
    pending.push_back({(int)outs.size() - 1, (int)ks.size() - 1});
    pend_j.push_back(-1);  // se K[V^T - V]: slab holds -K -> - se
    if (nspin == 2) outs.push_back({hmm + q * slab + n2, {}});  // stays zero
  }
  const int nJ = (int)js.size();
  for (size_t k = 0; k < pending.size(); ++k)
    outs[pending[k].first].terms.push_back({nJ + pending[k].second, pend_j[k] == 2 ? 2.0 * se : (pend_j[k] == 1 ? se : -se)});
  if ((rc = jk_slabs(ctx, mats, js, ks, 1.0, 1.0, 4.0 * js.size() + 16.0 * ks.size()))) return rc;
  const int NM = nJ + (int)ks.size();
  for (const Out& o : outs)
    if ((rc = jk_fetch(ctx, NM, o.terms, o.dst))) return rc;
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}

// int2_umrsf_data_t (tdhf_mrsf_lib.F90:28-32, update :337-426): d3, f3 (nvec, 11, nbf, nbf), nvec fastest.
//   c = 1..8:  f3 = sc J[d3] - se K[d3];   c = 9, 10:  f3 = - se K[d3^T] (the mixed-spin permutation);   c = 11:  f3 = - se K[d3]
// npass = 2 (int2_run_cam): pass 2 = Erf-attenuated exchange of component 11 only with `beta`.
static int umrsf_core(oqpb_ctx* ctx, const double* d3, int nvec, double se, double sc, double* f3, long long* nskipped,
                      int npass, double se2, double mu) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (nvec < 1 || !d3 || !f3) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  if (npass == 2 && (rc = oqpb_set_screening_cam(ctx, mu, nullptr))) return rc;
  WholeSliceGuard whole(ctx);
  const int nbf = ctx->nbf, ns = ctx->nshell, ncomp = 11, NM = nvec * ncomp;
  const long n2 = (long)nbf * nbf, npairs = (long)ns * (ns + 1) / 2;
  const size_t bytes = (size_t)n2 * NM * sizeof(double);
  CK(ctx->d_Din.ensure(bytes));
  CK(ctx->d_gen_in.ensure(bytes));
  CK(ctx->d_gen_out.ensure(bytes));
  CK(cudaMemcpyAsync(ctx->d_Din.p, d3, bytes, cudaMemcpyHostToDevice, ctx->stream));
  // screening on the caller's d3 (shell_den_screen_mrsf, tdhf_mrsf_lib.F90:189-214), before components 9/10 are transposed
  CK(cudaMemsetAsync(ctx->d_maxden.p, 0, 8, ctx->stream));
  k_shlden_gen<<<(unsigned)((npairs + 127) / 128), 128, 0, ctx->stream>>>(
      ns, npairs, ctx->d_aooff.as<int>(), ctx->d_naos.as<int>(), ctx->d_Din.as<double>(), NM, nbf, ctx->d_dsh.as<double>(),
      ctx->d_maxden.as<unsigned long long>());
  k_umrsf_prepare<<<(unsigned)(((size_t)n2 * NM + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_Din.as<double>(), ctx->d_gen_in.as<double>(),
                                                                                     nbf, nvec, ncomp);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(ctx->d_gen_out.p, 0, bytes, ctx->stream));
  BuildSpec S;
  S.mode = MODE_GEN;
  S.Pgen = ctx->d_gen_in.as<double>();
  S.Fgen = ctx->d_gen_out.as<double>();
  S.gen_nm = NM; S.gen_ncoul = 8; S.gen_nvec = nvec;
  S.cj = sc; S.ck = se;
  S.digest_flops_per_int = (8.0 * 4 + 8.0 * 11) * 2 * nvec;
  if ((rc = run_build(ctx, S))) return rc;
  if (npass == 2) {
    S.attenuated = true;
    S.gen_ncoul = 0; S.gen_xoff = 10 * nvec; S.gen_mcount = nvec;
    S.cj = 0.0; S.ck = se2;
    S.digest_flops_per_int = 16.0 * nvec;
    if ((rc = run_build(ctx, S))) return rc;
  }
  CK(cudaMemcpyAsync(f3, ctx->d_gen_out.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (nskipped) *nskipped = ctx->st_skipped;
  return OQPB_OK;
}
int oqpb_jk_umrsf(oqpb_ctx* ctx, const double* d3, int nvec, double se, double sc, double* f3, long long* nskipped) {
  return umrsf_core(ctx, d3, nvec, se, sc, f3, nskipped, 1, 0.0, 0.0);
}
int oqpb_jk_umrsf_cam(oqpb_ctx* ctx, const double* d3, int nvec, double alpha, double beta, double mu, double alpha_coulomb,
                      double* f3, long long* nskipped) {
  return umrsf_core(ctx, d3, nvec, alpha, alpha_coulomb, f3, nskipped, 2, beta, mu);
}

int oqpb_last_stats(oqpb_ctx* ctx, long long* s) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  s[0] = ctx->st_survivors; s[1] = ctx->st_skipped; s[2] = 0; s[3] = ctx->st_launches;
  return OQPB_OK;
}
double oqpb_last_flops(oqpb_ctx* ctx) { return ctx ? ctx->st_flops : 0.0; }
double oqpb_last_kernel_ms(oqpb_ctx* ctx) { return ctx ? ctx->st_kernel_ms : 0.0; }

int oqpb_profile(oqpb_ctx* ctx, int enable, double* out /* 55*4 or NULL */) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  if (out) memcpy(out, ctx->prof, sizeof ctx->prof);
  ctx->profile = enable != 0;
  memset(ctx->prof, 0, sizeof ctx->prof);
  return OQPB_OK;
}

int oqpb_record_quartets(oqpb_ctx* ctx, int enable) {
  if (!ctx) return OQPB_ERR_BAD_ARG;
  ctx->record = enable != 0;
  return OQPB_OK;
}
long long oqpb_get_quartets(oqpb_ctx* ctx, int* ijkl, long long maxq) {
  if (!ctx) return -1;
  long long n = (long long)ctx->rec.size() / 4;
  if (ijkl) memcpy(ijkl, ctx->rec.data(), (size_t)std::min(n, maxq) * 4 * sizeof(int));
  return n;
}

int oqpb_get_shell_density(oqpb_ctx* ctx, double* dsh, double* max_den) {
  if (!ctx || !ctx->have_screen) return OQPB_ERR_STATE;
  cudaSetDevice(ctx->device);
  CK(cudaMemcpy(dsh, ctx->d_dsh.p, (size_t)ctx->nshell * ctx->nshell * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(max_den, ctx->d_maxden.p, sizeof(double), cudaMemcpyDeviceToHost));
  return OQPB_OK;
}

int oqpb_eri_block(oqpb_ctx* ctx, int i, int j, int k, int l, double* out, int* nout) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  cudaSetDevice(ctx->device);
  const PairTable& T = ctx->run;
  auto find = [&](int a, int b, int& pc, int& idx) {
    int la = ctx->am[a], lb = ctx->am[b];
    pc = la >= lb ? pair_class(la, lb) : pair_class(lb, la);
    int hi = std::max(a, b), lo = std::min(a, b);
    int canon = hi * (hi + 1) / 2 + lo;
    for (int lst = pc * NBK; lst < (pc + 1) * NBK; ++lst)
      for (int e = T.cls_off[lst]; e < T.cls_off[lst + 1]; ++e)
        if (T.canon[e] == canon) { idx = e - T.cls_off[lst]; pc = lst; return; }
    idx = -1;
  };
  int pca, pcb, ea, eb;  // pair lists
  find(i, j, pca, ea);
  find(k, l, pcb, eb);
  bool swapped = pca < pcb;
  if (swapped) { std::swap(pca, pcb); std::swap(ea, eb); }
  DevBuf d_task, d_cnt, d_out;
  int2 task = make_int2(ea, eb);
  unsigned cnt[2] = {1, 0};
  CK(d_task.ensure(sizeof(int2)));
  CK(d_cnt.ensure(sizeof cnt));
  CK(d_out.ensure(10000 * sizeof(double)));
  CK(cudaMemcpy(d_task.p, &task, sizeof task, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cnt.p, cnt, sizeof cnt, cudaMemcpyHostToDevice));
  EriArgs A;
  fill_common_args(ctx, T, pca, pcb, A);
  A.tasks = d_task.as<int2>();
  A.ntasks = d_cnt.as<unsigned>();
  A.counter = d_cnt.as<unsigned>() + 1;
  A.prim_cutoff = ctx->cut.pair * ctx->cut.pair;
  A.mode = MODE_BLOCK;
  A.blockout = d_out.as<double>();
  const ClassEntry& ce = class_table(ctx->pure_l[2] | (ctx->pure_l[3] << 1))[quartet_class(pc_of(pca), pc_of(pcb))];
  const bool kown = ce.launch_kown != nullptr && (ctx->use_kown >= 2 || (ctx->use_kown == 1 && ce.kown_default));
  CK((kown ? ce.launch_kown : ce.launch)(A, 1, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // kernel block order: (A,B,C,D) = (bra.sa, bra.sb, ket.sa, ket.sb); map back to the caller's (i,j,k,l)
  const PairEntry& pb = T.ent[T.cls_off[pca] + ea];
  const PairEntry& pk = T.ent[T.cls_off[pcb] + eb];
  int sh[4] = {pb.sa, pb.sb, pk.sa, pk.sb};
  int dims[4];
  for (int s = 0; s < 4; ++s) dims[s] = ctx->naos[sh[s]];
  std::vector<double> blk((size_t)dims[0] * dims[1] * dims[2] * dims[3]);
  CK(cudaMemcpy(blk.data(), d_out.p, blk.size() * sizeof(double), cudaMemcpyDeviceToHost));
  int want[4] = {i, j, k, l};
  // position of each caller index in kernel order
  int pos[4];
  bool used[4] = {false, false, false, false};
  int braw[2] = {swapped ? 2 : 0, swapped ? 3 : 1};  // caller positions forming the kernel's bra
  int ketw[2] = {swapped ? 0 : 2, swapped ? 1 : 3};
  // bra: kernel slots 0,1 ; ket: slots 2,3
  auto assign = [&](int w0, int w1, int s0, int s1) {
    if (want[w0] == sh[s0] && want[w1] == sh[s1] && !(want[w0] == want[w1] && false)) { pos[w0] = s0; pos[w1] = s1; }
    else { pos[w0] = s1; pos[w1] = s0; }
  };
  assign(braw[0], braw[1], 0, 1);
  assign(ketw[0], ketw[1], 2, 3);
  (void)used;
  int nd[4];
  for (int w = 0; w < 4; ++w) nd[w] = dims[pos[w]];
  for (int w = 0; w < 4; ++w) nout[w] = nd[w];
  int c[4];
  for (c[0] = 0; c[0] < nd[0]; ++c[0])
    for (c[1] = 0; c[1] < nd[1]; ++c[1])
      for (c[2] = 0; c[2] < nd[2]; ++c[2])
        for (c[3] = 0; c[3] < nd[3]; ++c[3]) {
          int kidx[4];
          for (int w = 0; w < 4; ++w) kidx[pos[w]] = c[w];
          out[((c[0] * nd[1] + c[1]) * nd[2] + c[2]) * nd[3] + c[3]] =
              blk[(((size_t)kidx[0] * dims[1] + kidx[1]) * dims[2] + kidx[2]) * dims[3] + kidx[3]];
        }
  d_task.release(); d_cnt.release(); d_out.release();
  return OQPB_OK;
}

int oqpb_rys(oqpb_ctx* ctx, int nroots, int npts, const double* x, double* t2, double* w) {
  if (!ctx || nroots < 1 || nroots > RYS_MAXR) return OQPB_ERR_BAD_ARG;
  cudaSetDevice(ctx->device);
  DevBuf dx, dt, dw;
  CK(dx.ensure(npts * sizeof(double)));
  CK(dt.ensure((size_t)npts * nroots * sizeof(double)));
  CK(dw.ensure((size_t)npts * nroots * sizeof(double)));
  CK(cudaMemcpy(dx.p, x, npts * sizeof(double), cudaMemcpyHostToDevice));
  EriArgs A;
  memset(&A, 0, sizeof A);
  A.rys_tab = rys_table(ctx, nroots);
  A.rys_xmax = RYS_XMAX_H[nroots - 1];
  for (int k = 0; k < 7; ++k) { A.herm_r[k] = RYS_HERM_R_H[nroots - 1][k]; A.herm_w[k] = RYS_HERM_W_H[nroots - 1][k]; }
  k_rys_test<<<(npts + 127) / 128, 128, 0, ctx->stream>>>(A, nroots, npts, dx.as<double>(), dt.as<double>(), dw.as<double>());
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(t2, dt.p, (size_t)npts * nroots * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(w, dw.p, (size_t)npts * nroots * sizeof(double), cudaMemcpyDeviceToHost));
  dx.release(); dt.release(); dw.release();
  return OQPB_OK;
}

double oqpb_fp64_peak_tflops(oqpb_ctx* ctx) {
  if (!ctx) return 0.0;
  if (ctx->fp64_peak > 0) return ctx->fp64_peak;
  cudaSetDevice(ctx->device);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, ctx->device);
  int nb = prop.multiProcessorCount * 8, nt = 256, iters = 1 << 16;
  double* d = nullptr;
  if (cudaMalloc(&d, (size_t)nb * nt * sizeof(double)) != cudaSuccess) return 0.0;
  k_fp64_peak<<<nb, nt, 0, ctx->stream>>>(d, 1024);
  cudaStreamSynchronize(ctx->stream);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(ctx->ev0, ctx->stream);
    k_fp64_peak<<<nb, nt, 0, ctx->stream>>>(d, iters);
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    double tf = 2.0 * 8.0 * (double)iters * nb * nt / (ms * 1e-3) / 1e12;
    best = std::max(best, tf);
  }
  cudaFree(d);
  ctx->fp64_peak = best;
  return best;
}

int oqpb_set_default_ctx(oqpb_ctx* ctx) { g_default_ctx = ctx; return OQPB_OK; }
int oqpb_set_default_scftype(int urohf) { g_default_urohf = urohf; return OQPB_OK; }

// routec_fock_jk: signature and semantics of routec_bridge.F90:33-40 (f returned ready to use, info != 0 -> native)
void routec_fock_jk(const double* d, double* f, const int* nbf, const int* nfocks, const double* se, const double* sc,
                    int* info) {
  if (info) *info = 1;
  oqpb_ctx* ctx = g_default_ctx;
  if (!ctx || !d || !f || !nbf || !nfocks || !se || !sc) return;
  if (*nbf != ctx->nbf) return;
  int urohf = g_default_urohf >= 0 ? g_default_urohf : (*nfocks == 2 ? 1 : 0);
  int rc = oqpb_fock(ctx, urohf, d, f, *nfocks, *se, *sc, 1, nullptr);
  if (info) *info = rc;
}

}  // extern "C"

#include "sigma_session.cuh"
