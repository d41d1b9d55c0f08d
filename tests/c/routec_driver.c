/* C99 host of the drop-in boundary: what a non-Python, non-Fortran caller (or the Fortran shim, through ISO_C_BINDING) does.
 * Reads a basis + packed density dump written by tests/test_gpu_parity.py::test_c_host_program, registers a context and
 * calls the LEGACY SEAM routec_fock_jk (routec_bridge.F90:33-40) by reference, then oqpb_fock and the multi-device
 * context when more than one GPU is visible.  Writes the Fock matrices back for the Python side to compare with the oracle.
 *
 *   gcc -std=c99 -I include tests/c/routec_driver.c -o /tmp/routec_driver -L openqp_b200 -lopenqp_b200 -Wl,-rpath,$PWD/openqp_b200
 *   /tmp/routec_driver dump.bin out.bin
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oqp_b200.h"

static void die(const char* what, int rc, oqpb_ctx* ctx) {
  fprintf(stderr, "%s failed: rc = %d (%s)\n", what, rc, ctx ? oqpb_last_error(ctx) : "");
  exit(2);
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s dump.bin out.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("dump"); return 1; }
  int hdr[4]; /* nshell, nprim, nbf, harmonic_active */
  if (fread(hdr, sizeof(int), 4, f) != 4) return 1;
  const int nshell = hdr[0], nprim = hdr[1], nbf = hdr[2], harmonic_active = hdr[3];
  const long ntri = (long)nbf * (nbf + 1) / 2;
  int* iarr = (int*)malloc(sizeof(int) * 6 * nshell);
  double* ex = (double*)malloc(sizeof(double) * nprim);
  double* cc = (double*)malloc(sizeof(double) * nprim);
  double* cen = (double*)malloc(sizeof(double) * 3 * nshell);
  double* d = (double*)malloc(sizeof(double) * ntri);
  double* fock = (double*)calloc(3 * ntri, sizeof(double));
  if (fread(iarr, sizeof(int), 6 * nshell, f) != (size_t)(6 * nshell) || fread(ex, sizeof(double), nprim, f) != (size_t)nprim ||
      fread(cc, sizeof(double), nprim, f) != (size_t)nprim || fread(cen, sizeof(double), 3 * nshell, f) != (size_t)(3 * nshell) ||
      fread(d, sizeof(double), ntri, f) != (size_t)ntri) { fprintf(stderr, "short dump\n"); return 1; }
  fclose(f);
  const int *am = iarr, *harm = iarr + nshell, *ncontr = iarr + 2 * nshell, *goff = iarr + 3 * nshell, *aooff = iarr + 4 * nshell,
            *naos = iarr + 5 * nshell;

  oqpb_ctx* ctx = NULL;
  int rc = oqpb_ctx_create(&ctx, 0);
  if (rc) die("oqpb_ctx_create", rc, NULL);
  if ((rc = oqpb_set_basis(ctx, nshell, nprim, am, harm, ncontr, goff, aooff, naos, ex, cc, cen, harmonic_active))) die("oqpb_set_basis", rc, ctx);
  if ((rc = oqpb_set_cutoff(ctx, 5e-11))) die("oqpb_set_cutoff", rc, ctx);
  if ((rc = oqpb_set_screening(ctx, NULL))) die("oqpb_set_screening", rc, ctx);

  /* 1. the legacy seam, all scalars by reference; info = 0 on success, f ready to use */
  oqpb_set_default_ctx(ctx);
  oqpb_set_default_scftype(0);
  int info = 7, nfocks = 1;
  double se = 1.0, sc = 1.0;
  routec_fock_jk(d, fock, &nbf, &nfocks, &se, &sc, &info);
  if (info != 0) die("routec_fock_jk", info, ctx);
  int bad = nbf + 1;
  routec_fock_jk(d, fock + 2 * ntri, &bad, &nfocks, &se, &sc, &info);
  if (info == 0) { fprintf(stderr, "routec_fock_jk accepted a wrong nbf\n"); return 3; }

  /* 2. the explicit entry, raw accumulator + nskipped */
  long long nskipped = -1;
  if ((rc = oqpb_fock(ctx, 0, d, fock + ntri, 1, 1.0, 1.0, 1, &nskipped))) die("oqpb_fock", rc, ctx);
  oqpb_set_default_ctx(NULL);
  oqpb_ctx_destroy(ctx);

  /* 3. one process, every visible GPU (skipped on a single-GPU box): same call, same answer */
  int ndev_used = 1;
  oqpb_ctx* mctx = NULL;
  if (oqpb_ctx_create_multi(&mctx, 2, NULL) == 0) {
    ndev_used = oqpb_ctx_ndevices(mctx);
    if ((rc = oqpb_set_basis(mctx, nshell, nprim, am, harm, ncontr, goff, aooff, naos, ex, cc, cen, harmonic_active))) die("multi set_basis", rc, mctx);
    if ((rc = oqpb_set_cutoff(mctx, 5e-11))) die("multi set_cutoff", rc, mctx);
    if ((rc = oqpb_set_screening(mctx, NULL))) die("multi set_screening", rc, mctx);
    if ((rc = oqpb_fock(mctx, 0, d, fock + 2 * ntri, 1, 1.0, 1.0, 1, NULL))) die("multi oqpb_fock", rc, mctx);
    oqpb_ctx_destroy(mctx);
  } else {
    memcpy(fock + 2 * ntri, fock + ntri, sizeof(double) * ntri);
  }

  f = fopen(argv[2], "wb");
  if (!f) { perror("out"); return 1; }
  fwrite(&nskipped, sizeof nskipped, 1, f);
  fwrite(&ndev_used, sizeof ndev_used, 1, f);
  fwrite(fock, sizeof(double), 3 * ntri, f);
  fclose(f);
  printf("routec_driver: nbf %d, nskipped %lld, devices %d\n", nbf, nskipped, ndev_used);
  return 0;
}
