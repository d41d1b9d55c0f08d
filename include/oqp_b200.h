/*
 * oqp_b200.h -- C ABI of libopenqp_b200.so, a B200 (sm_100a) direct-SCF two-electron J/K Fock builder
 * that drops in for OpenQP's `int2` driver and its J/K consumers.
 *
 * Conventions (identical to the reference's existing GPU seam, source/modules/routec_bridge.F90:30-41):
 *   - plain C, `extern "C"`, host pointers unless the name ends in `_dev`, caller owns every array,
 *     nothing is retained past return except what the ctx caches (basis, pair table, Schwarz matrix);
 *   - every entry returns `info` (also stored through the trailing `int* info` of the legacy symbol):
 *     0 = success, anything else = "not handled" -> the Fortran caller silently runs its native path
 *     (routec_bridge.F90:258-263).  There is NO CPU fallback inside this library: without a CUDA device
 *     every compute entry returns OQPB_ERR_NO_DEVICE.
 *   - matrices: packed lower-triangular `ij = i(i-1)/2 + j` (1-based, i >= j) for SCF densities/Focks
 *     (int2.F90:1436-1447); column-major nbf x nbf for response densities (tdhf_lib.F90:13-15);
 *     d3/f3(nvec, ncomp, nbf, nbf) with nvec fastest for MRSF (tdhf_mrsf_lib.F90:10-11).
 *   - shell/AO indices in `oqpb_set_basis` are 0-based offsets (subtract 1 from the Fortran arrays).
 *
 * Each entry cites the reference interface it replaces.
 */
#ifndef OQP_B200_H
#define OQP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oqpb_ctx oqpb_ctx;

enum {
  OQPB_OK = 0,
  OQPB_ERR_NO_DEVICE = 1,   /* no CUDA device / driver: caller must use its native path            */
  OQPB_ERR_BAD_ARG = 2,
  OQPB_ERR_UNSUPPORTED = 3, /* L > 3, mixed harmonic flags inside one L, nbf too large, ...         */
  OQPB_ERR_STATE = 4,       /* call order violated (basis / cutoff / screening not set)            */
  OQPB_ERR_CUDA = 5         /* a CUDA runtime call failed; see oqpb_last_error()                    */
};

/* flags of oqpb_jk_td: fields of int2_td_data_t, tdhf_lib.F90:11-31 */
enum { OQPB_TD_APB = 1, OQPB_TD_AMB = 2, OQPB_TD_TDA = 4, OQPB_TD_TDA_COULOMB = 8 };

/* ---- lifetime: replaces int2_compute_t construction/clean (int2.F90:137-185, 245-289) ------------- */
int  oqpb_ctx_create(oqpb_ctx** ctx, int device);
/* One context over ndev GPUs of this node, driven from a single process (devices == NULL: 0 .. ndev-1): basis, pair
 * table, Schwarz matrix and densities are replicated, the bra shell-pair list is split over the devices (on top of
 * oqpb_set_partition's rank split), and the partial results are summed by ONE ncclAllReduce(ncclDouble, ncclSum) on
 * the compute streams -- the replacement of `pe%allreduce` at the end of int2_twoei (int2.F90:1392-1397,
 * parallel.F90:429-440).  A non-MPI OpenQP (ENABLE_MPI is OFF by default, CMakeLists.txt:76) reaches every GPU of the
 * box this way: oqpb_fock, oqpb_fock_cam, oqpb_jk_mrsf[_cam] and routec_fock_jk use all members; oqpb_jk_td runs on the
 * first device.  The *_dev entries act on the first device's slice only.  NCCL (libnccl.so.2) is loaded on demand.   */
int  oqpb_ctx_create_multi(oqpb_ctx** ctx, int ndev, const int* devices);
int  oqpb_ctx_ndevices(const oqpb_ctx* ctx);
void oqpb_ctx_destroy(oqpb_ctx* ctx);
const char* oqpb_last_error(const oqpb_ctx* ctx);

/* The arrays of `basis_set` (basis_tools.F90:32-51) after normalize_primitives (:277-303):
 * am, harmonic, ncontr, g_offset, ao_offset, naos [nshell]; ex, cc [nprim]; centers [3*nshell] (Bohr,
 * = shell_centers, basis_tools.F90:1316-1326).  harmonic_active = constants.F90 HARMONIC_ACTIVE.      */
int oqpb_set_basis(oqpb_ctx* ctx, int nshell, int nprim, const int* am, const int* harmonic,
                   const int* ncontr, const int* g_offset, const int* ao_offset, const int* naos,
                   const double* ex, const double* cc, const double* centers, int harmonic_active);

/* int2_compute_t%init cutoffs + %set_cutoff (int2.F90:260-272; int2_pairs.F90:285-294):
 * integral = c, pair = 1e-2 c, quartet = 1e-4 c, exponent = 25 ln 10.  (Re)builds the device pair table. */
int oqpb_set_cutoff(oqpb_ctx* ctx, double cutoff);

/* int2_compute_t%set_screening -> ints_exchange (int2.F90:473-477, 1582-1737).
 * schwarz_in == NULL: compute Q_ij = sqrt(max|(ij|ij)|) on the device with the reference's tight cutoffs
 * (1e-15, 1e-17, 1e-17, 50); otherwise upload the caller's nshell x nshell matrix (bit-exact screening
 * against a host-computed Q).  Sorts the pair lists by Q for the quartet enumeration.                   */
int oqpb_set_screening(oqpb_ctx* ctx, const double* schwarz_in);
int oqpb_get_schwarz(oqpb_ctx* ctx, double* schwarz_out /* nshell*nshell */);

/* Multi-GPU / MPI replicated-data split of the work, the role of the reference's `mod(ij_pair, size) == rank`
 * (int2.F90:759-761).  Every pair list (angular-momentum class x contraction bucket, Schwarz-sorted) pair is shared by
 * s = round(estimated ms / 2.5) of the ranks (1 <= s <= nranks, the least loaded ones by a per-class cost model that every
 * rank evaluates identically) and the bras of the list are dealt cyclically among those, p % s == j: long list pairs are
 * split over all ranks with the same class mix, short ones run as one launch on one rank instead of nranks tiny ones.  A
 * rank's slice (and its nskipped) therefore differs from the slice a native OpenQP rank would take: all ranks of a job
 * must use this library; the slices are disjoint and their nskipped add up to the single-rank value.  The caller sums
 * the partial results (pe%allreduce, int2.F90:1396) -- or uses a multi-device context.                               */
int oqpb_set_partition(oqpb_ctx* ctx, int rank, int nranks);

/* Test hook: restrict the following builds to the quartets whose reference bra pair (the canonically larger
 * shell pair, the `ij_pair` of int2.F90:756-780) is flagged in mask[i(i+1)/2+j] (0-based, i >= j, npairs =
 * nshell(nshell+1)/2 bytes); NULL lifts the restriction.  Lets a strided sample of a large build be compared
 * with the CPU oracle run on the same bra subset (the oracle's stride/offset over its cost-sorted list).   */
int oqpb_set_bra_mask(oqpb_ctx* ctx, const unsigned char* mask, long long npairs);

/* fock_jk (scf_addons.F90:1063-1214) = int2_rhf_data_t / int2_urohf_data_t consumers
 * (int2.F90:1414-1578) + 0.5/diagonal post-scaling (:1177-1185) when `post` != 0:
 *   RHF  (urohf = 0): f_m = sc*J[d_m] - 1/2 se*K[d_m],  m = 1..nfocks
 *   UROHF(urohf = 1): f_s = sc*J[d_1+d_2] - se*K[d_s],   nfocks = 2
 * d, f: packed (ntri, nfocks).  nskipped = int2_compute_t%skipped (nschwz).  Host pointers.           */
int oqpb_fock(oqpb_ctx* ctx, int urohf, const double* d, double* f, int nfocks, double scale_exchange,
              double scale_coulomb, int post, long long* nskipped);
/* same with DEVICE pointers, no host copies: the result stays in HBM for the caller's NCCL all-reduce;
 * oqpb_fock_post_dev applies the 0.5/diag scaling afterwards.  The kernels run on the ctx stream (and its helper
 * streams, joined back into it), but the call BLOCKS the host until the build's statistics are back.               */
int oqpb_fock_dev(oqpb_ctx* ctx, int urohf, const double* d_dev, double* f_dev, int nfocks,
                  double scale_exchange, double scale_coulomb);
int oqpb_fock_post_dev(oqpb_ctx* ctx, double* f_dev, int nfocks);

/* Range-separated (CAM) build = int2_compute_t%run(consumer, cam=.true., alpha, beta, mu), i.e. int2_run_cam
 * (int2.F90:538-584): pass 1 regular integrals with (scale_coulomb, scale_exchange) = (alpha_coulomb, alpha), pass 2
 * Erf-attenuated integrals erf(mu r)/r (int_rys.F90:179-181, 225-227) with (beta_coulomb, beta), screened with the
 * attenuated Schwarz matrix (int2.F90:674-685), both into the same Fock matrices.  The legacy seam declines CAM
 * (scf_addons.F90:1141); fock_jk passes alpha_coulomb = 1, beta_coulomb = 0 (scf_addons.F90:1167-1173).
 * oqpb_set_screening_cam computes (or takes) the attenuated Schwarz matrix once per geometry and mu; the fock entries
 * call it themselves when the cached mu differs.  nskipped = the second pass's count, as the reference leaves it.   */
int oqpb_set_screening_cam(oqpb_ctx* ctx, double mu, const double* schwarz_att_in /* nshell*nshell or NULL */);
int oqpb_get_schwarz_cam(oqpb_ctx* ctx, double* out /* nshell*nshell */);
int oqpb_fock_cam(oqpb_ctx* ctx, int urohf, const double* d, double* f, int nfocks, double alpha, double beta,
                  double mu, double alpha_coulomb, double beta_coulomb, int post, long long* nskipped);
int oqpb_fock_cam_dev(oqpb_ctx* ctx, int urohf, const double* d_dev, double* f_dev, int nfocks, double alpha,
                      double beta, double mu, double alpha_coulomb, double beta_coulomb);
int oqpb_synchronize(oqpb_ctx* ctx);
void* oqpb_stream(oqpb_ctx* ctx); /* cudaStream_t the ctx launches on */
/* launch on the caller's stream instead (e.g. the stream the caller's NCCL communicator is ordered on) */
int oqpb_set_stream(oqpb_ctx* ctx, void* cuda_stream);

/* int2_td_data_t (tdhf_lib.F90:11-31, update :140-224, parallel_stop symmetrisation :107-109):
 * d2, apb, amb: column-major (nbf, nbf, nvec).  apb is returned symmetrised (apb + apb^T).            */
int oqpb_jk_td(oqpb_ctx* ctx, const double* d2, int nvec, int flags, double scale_exchange,
               double scale_coulomb, double* apb, double* amb, long long* nskipped);

/* int2_mrsf_data_t (tdhf_mrsf_lib.F90:8-26, update :218-333), pass 1:
 * f3(v,c,:,:) = sc*J[d3(v,c)] (c <= 4) - se*K[d3(v,c)] (all c); d3, f3: (nvec, ncomp, nbf, nbf).      */
int oqpb_jk_mrsf(oqpb_ctx* ctx, const double* d3, int nvec, int ncomp, double scale_exchange,
                 double scale_coulomb, double* f3, long long* nskipped);
/* same with DEVICE pointers (d3_dev, f3_dev: nvec*ncomp*nbf*nbf doubles each, f3_dev is overwritten): the Davidson
 * trial densities and their Fock-like matrices stay in HBM between iterations (the sigma session of
 * routec_sig_iter, routec_sig.F90:28-56, would sit on this entry).                                     */
int oqpb_jk_mrsf_dev(oqpb_ctx* ctx, const double* d3_dev, int nvec, int ncomp, double scale_exchange,
                     double scale_coulomb, double* f3_dev);

/* Generic J/K for n general (non-symmetric) AO matrices P (nbf, nbf, n), column-major:
 *   J_m(a,b) = sum_cd (ab|cd) P_m(c,d)  when want_j[m];   K_m(a,c) = sum_bd (ab|cd) P_m(b,d)  when want_k[m]
 * screened with max|P_m| per shell block over all m (shltd, tdhf_lib.F90:300-325).  J, K: (nbf, nbf, n); slabs that were
 * not asked for are left untouched.  Every J/K consumer of the reference is a linear combination of these; the Z-vector and
 * gradient drivers (modules/tdhf_mrsf_z_vector.F90, tdhf_sf_z_vector.F90, tdhf_z_vector.F90) can call it directly.        */
int oqpb_jk(oqpb_ctx* ctx, int n, const double* P, const int* want_j, const int* want_k, double* J, double* K,
            long long* nskipped);
/* int2_tdgrd_data_t (tdhf_lib.F90:33-36, update :228-295): two spin blocks d2(nbf,nbf,2); flags = OQPB_TD_APB | OQPB_TD_AMB.
 *   apb(:,:,s) = 2 sc J[P_1+P_2] - se K[P_s+P_s^T] (symmetrised as parallel_stop does), amb(:,:,1) = se K[P_1^T-P_1], amb(:,:,2) = 0 */
int oqpb_jk_tdgrd(oqpb_ctx* ctx, const double* d2, int flags, double scale_exchange, double scale_coulomb, double* apb,
                  double* amb, long long* nskipped);
/* int2_rpagrd_data_t (tdhf_lib.F90:42-57, 1068-1320): xpy, xmy, t (nbf, nbf, nspin, np | nm | nt) -> hpp = H+[X+Y],
 * hpt = H+[T] (both symmetrised, :1138-1141), hmm = H-[X-Y].  X+Y and T must be symmetric matrices (they are at every call
 * site: the reference's update reads one triangle of them).  Pointers may be NULL when the count is 0.                   */
int oqpb_jk_rpagrd(oqpb_ctx* ctx, int nspin, int np, int nm, int nt, const double* xpy, const double* xmy, const double* t,
                   double scale_exchange, double scale_coulomb, double* hpp, double* hpt, double* hmm, long long* nskipped);
/* int2_umrsf_data_t (tdhf_mrsf_lib.F90:28-32, update :337-426): d3, f3 (nvec, 11, nbf, nbf), nvec fastest; the _cam variant
 * is int2_run_cam (pass 2 = attenuated exchange of component 11 only).                                                  */
int oqpb_jk_umrsf(oqpb_ctx* ctx, const double* d3, int nvec, double scale_exchange, double scale_coulomb, double* f3,
                  long long* nskipped);
int oqpb_jk_umrsf_cam(oqpb_ctx* ctx, const double* d3, int nvec, double alpha, double beta, double mu, double alpha_coulomb,
                      double* f3, long long* nskipped);

/* The response consumers through int2_run_cam (range-separated functionals):
 *   TD   (tdhf_lib.F90:140-224): the same update in both passes, pass 2 with Erf-attenuated integrals and
 *        (beta_coulomb, beta);
 *   MRSF (tdhf_mrsf_lib.F90:279-326): pass 1 all components with (alpha_coulomb, alpha), pass 2 the exchange of
 *        component 7 only with beta (ncomp must be 7).                                                    */
int oqpb_jk_td_cam(oqpb_ctx* ctx, const double* d2, int nvec, int flags, double alpha, double beta, double mu,
                   double alpha_coulomb, double beta_coulomb, double* apb, double* amb, long long* nskipped);
int oqpb_jk_mrsf_cam(oqpb_ctx* ctx, const double* d3, int nvec, int ncomp, double alpha, double beta, double mu,
                     double alpha_coulomb, double* f3, long long* nskipped);
int oqpb_jk_mrsf_cam_dev(oqpb_ctx* ctx, const double* d3_dev, int nvec, int ncomp, double alpha, double beta,
                         double mu, double alpha_coulomb, double* f3_dev);

/* ---- introspection used by the parity tests and the benchmark ------------------------------------- */
/* statistics of the last build: [0] surviving shell quartets, [1] skipped (nschwz), [2] primitive
 * quartets evaluated is not tracked (0), [3] kernel launches, [4] algorithmic FLOPs (SURVEY 8d-1 model,
 * as a double bit-cast is avoided: returned through oqpb_last_flops).                                  */
int oqpb_last_stats(oqpb_ctx* ctx, long long* stats4);
double oqpb_last_flops(oqpb_ctx* ctx);
double oqpb_last_kernel_ms(oqpb_ctx* ctx); /* CUDA-event time of the ERI/digest kernels of the last build */
/* surviving canonical shell quartets (i>=j, k>=l, (ij)>=(kl), 0-based) of the last build, unordered;
 * returns the count, writes at most maxq quadruples.  Must be enabled before the build.               */
int oqpb_record_quartets(oqpb_ctx* ctx, int enable);
/* per-class profiling (serialises launches): out[55][4] = ms, quartets, primitive quartets, model FLOPs of the
 * builds since the last call; class index = pa*(pa+1)/2+pb over pair classes ss ps pp ds dp dd fs fp fd ff */
int oqpb_profile(oqpb_ctx* ctx, int enable, double* out);
long long oqpb_get_quartets(oqpb_ctx* ctx, int* ijkl, long long maxq);
/* shell-block max|D| of the last build (shlden, int2.F90:999-1047), nshell*nshell                     */
int oqpb_get_shell_density(oqpb_ctx* ctx, double* dsh, double* max_den);
/* one shell quartet (0-based, any order) -> unit-normalised, pure-projected block out(i,j,k,l), l fastest
 * (shellquartet, int2.F90:1051-1183); nout[4] receives the block dimensions.                          */
int oqpb_eri_block(oqpb_ctx* ctx, int i, int j, int k, int l, double* out, int* nout);
/* Rys roots (as t^2) and weights from the device tables, for n <= 7 (rys_root_t%evaluate, rys.F90:24-36) */
int oqpb_rys(oqpb_ctx* ctx, int nroots, int npts, const double* x, double* t2, double* w);
/* measured FP64 FMA peak of this device in TFLOP/s (roofline denominator; MEASURED_PEAKS.json has none) */
double oqpb_fp64_peak_tflops(oqpb_ctx* ctx);

/* ---- legacy seam, identical signature to routec_bridge.F90:33-40 / 81-89 --------------------------- */
/* Uses the process-global context registered with oqpb_set_default_ctx; RHF semantics for nfocks = 1,
 * UROHF for nfocks = 2 unless overridden by oqpb_set_default_scftype.                                  */
void routec_fock_jk(const double* d, double* f, const int* nbf, const int* nfocks, const double* scale_exchange,
                    const double* scale_coulomb, int* info);
int oqpb_set_default_ctx(oqpb_ctx* ctx);
int oqpb_set_default_scftype(int urohf);

/* ---- MRSF sigma session, identical signatures to source/modules/routec_sig.F90:28-56 ---------------- */
/* Replaces the reference's per-iteration triple  mrsfcbc -> int2_mrsf_data_t run -> mrsfmntoia + mrsfesum
 * (tdhf_mrsf_lib.F90:940-1273, 218-333, 1463-1735, 1918-2036; caller and gate modules/tdhf_mrsf_energy.F90:648-713)
 * with one device call: trial amplitudes in, (A-B) X out, nothing else crosses the bus.  The session runs on the context
 * registered with oqpb_set_default_ctx (basis, cutoff and screening set; the reference raises the response cutoff to
 * max(int2e_cutoff, 1e-8) first, tdhf_mrsf_energy.F90:516-519).
 *   init: nbf, MO coefficients mo_a / mo_b (nbf, nbf) and MO-basis Fock matrices fmo_a / fmo_b (nbf, nbf), column-major;
 *         nocca, noccb = nocca - 2; kind 1 = singlet, 3 = triplet response.  Returns 0 when the session is ready.
 *   set_scale: exact-exchange scale of the response (scale_exchange = scale_coulomb of int2_mrsf_data_t).
 *   iter: bvec_mo (ntrial, nv_new) -> sigma_mo (ntrial, nv_new), ntrial = nocca (nbf - noccb), occupied index fastest;
 *         info = 0 on success, anything else tells the caller to take its native path (as the reference does).         */
int  routec_sig_init(const int* nbf, const double* mo_a, const double* mo_b, const double* fmo_a, const double* fmo_b,
                     const int* nocca, const int* noccb, const int* kind);
void routec_sig_set_scale(const double* s);
void routec_sig_iter(const double* bvec_mo, const int* nv_new, double* sigma_mo, int* info);
void routec_sig_free(void);

#ifdef __cplusplus
}
#endif
#endif
