/* A plain C host against include/fatode_b200.h (no torch, no Python): NRANKS processes, one GPU each, shard a batch of CBM-IV
 * cells, integrate forward + adjoint, and let the library sum Mu over all ranks (NCCL inside libfatode_b200.so).
 * Checks: rc / IERR, the ISTATUS closed forms, and mu_sum(all ranks) == sum over the per-cell Mu of every rank.
 *   usage: abi_two_ranks <nranks> <cells_per_rank> [hours]        (nranks = 1 works on a single-GPU box)
 * Built by tests/test_c_host.py: gcc -Iinclude tests/c/abi_two_ranks.c -Lfatode_b200/lib -lfatode_b200 -lcudart */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include "fatode_b200.h"

extern int cudaSetDevice(int);
enum { N = 32, NP = 29, NADJ = 32 };

static int run_rank(int nranks, int rank, long ncells, double hours, const char* idfile, const char* outfile) {
  if (cudaSetDevice(rank) != 0) { fprintf(stderr, "rank %d: cudaSetDevice failed\n", rank); return 2; }
  unsigned char id[128];
  FILE* f = fopen(idfile, "rb");
  if (!f || fread(id, 1, 128, f) != 128) { fprintf(stderr, "rank %d: cannot read the NCCL id\n", rank); return 2; }
  fclose(f);
  if (fatode_comm_init(nranks, rank, id) != 0) { fprintf(stderr, "rank %d: %s\n", rank, fatode_comm_last_error()); return 3; }
  double base[N];
  fatode_model_initial_state(FATODE_MODEL_CBM4, base);
  double* y = malloc(sizeof(double) * N * ncells);
  double* lam = calloc((size_t)N * NADJ * ncells, sizeof(double));
  double* mu = calloc((size_t)NP * NADJ * ncells, sizeof(double));
  int* ist = malloc(sizeof(int) * 20 * ncells);
  double* rst = malloc(sizeof(double) * 20 * ncells);
  int* ierr = malloc(sizeof(int) * ncells);
  double musum[NP * NADJ];
  for (long c = 0; c < ncells; ++c)          /* this rank's shard: a deterministic perturbation per GLOBAL cell index */
    for (int i = 0; i < N; ++i) y[c * N + i] = base[i] * (1.0 + 0.05 * sin(1.0 + 0.37 * i + 0.11 * (double)(rank * ncells + c)));
  const float f01 = 0.1f;
  double rtol[N], atol[N];
  for (int i = 0; i < N; ++i) { rtol[i] = (double)f01 * 1.0e-4; atol[i] = (double)f01 * 1.0e-6; }
  int icntrl[20] = {0};
  icntrl[2] = 5; /* Rodas4 */
  const double t0 = 12.0 * 3600.0, t1 = t0 + hours * 3600.0;
  int rc = fatode_ros_integrate_adj_batch(FATODE_MODEL_CBM4, ncells, N, NP, NADJ, y, lam, mu, t0, t1, NULL, NULL, atol, rtol, icntrl,
                                          NULL, ist, rst, ierr, musum, NULL, FATODE_MEM_HOST, NULL);
  if (rc != 0) { fprintf(stderr, "rank %d: rc %d: %s\n", rank, rc, fatode_last_error()); return 4; }
  int bad = 0;
  double local[NP * NADJ];
  memset(local, 0, sizeof local);
  for (long c = 0; c < ncells; ++c) {
    if (ierr[c] != 1) ++bad;
    const int* I = ist + 20 * c;
    const int A = I[3], att = I[2] - A;
    if (I[0] != 2 * A + 5 * att || I[6] != att * 6 + A * 6 * NADJ) ++bad;      /* Nfun, Nsol closed forms */
    for (int k = 0; k < NP * NADJ; ++k) local[k] += mu[(size_t)c * NP * NADJ + k];
  }
  /* the same sum by hand: every rank contributes its local sum through the library's own all-reduce */
  if (fatode_comm_allreduce_sum(local, NP * NADJ, FATODE_MEM_HOST, NULL) != 0) { fprintf(stderr, "%s\n", fatode_comm_last_error()); return 5; }
  double worst = 0.0, scale = 0.0;
  for (int k = 0; k < NP * NADJ; ++k) { scale = fmax(scale, fabs(local[k])); worst = fmax(worst, fabs(local[k] - musum[k])); }
  if (!(worst <= 1e-12 * scale) || scale == 0.0) ++bad;
  f = fopen(outfile, "w");
  fprintf(f, "rank %d of %d: cells %ld bad %d comm_size %d mu_sum_scale %.6e mu_sum_diff %.3e\n", rank, nranks, ncells, bad,
          fatode_comm_size(), scale, worst);
  fclose(f);
  fatode_comm_finalize();
  fatode_release_workspace();
  return bad ? 6 : 0;
}

int main(int argc, char** argv) {
  const int nranks = argc > 1 ? atoi(argv[1]) : 1;
  const long ncells = argc > 2 ? atol(argv[2]) : 16;
  const double hours = argc > 3 ? atof(argv[3]) : 2.0;
  char idfile[200], outfile[256];
  snprintf(idfile, sizeof idfile, "/tmp/fatode_nccl_id_%d", (int)getpid());
  unsigned char id[128];
  if (fatode_comm_unique_id(id) != 0) { fprintf(stderr, "%s\n", fatode_comm_last_error()); return 1; }
  FILE* f = fopen(idfile, "wb");
  fwrite(id, 1, 128, f);
  fclose(f);
  pid_t pids[16];
  for (int r = 0; r < nranks; ++r) {
    pids[r] = fork();
    if (pids[r] == 0) {
      snprintf(outfile, sizeof outfile, "%s.rank%d", idfile, r);
      _exit(run_rank(nranks, r, ncells, hours, idfile, outfile));
    }
  }
  int fail = 0;
  for (int r = 0; r < nranks; ++r) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { fprintf(stderr, "rank %d failed (status %d)\n", r, st); fail = 1; }
    snprintf(outfile, sizeof outfile, "%s.rank%d", idfile, r);
    FILE* g = fopen(outfile, "r");
    char line[512];
    if (g && fgets(line, sizeof line, g)) fputs(line, stdout);
    if (g) fclose(g);
    remove(outfile);
  }
  remove(idfile);
  puts(fail ? "FAILED" : "OK");
  return fail;
}
