/* abi_driver.c -- replays, from plain C, the exact call sequence of wrf_b200/fortran/solve_em_gpu.F90 against
 * libwrfgpu.so: wrfgpu_device_count, wrfgpu_create with a filled wrfgpu_dims / wrfgpu_config, the (name, pointer) tables of
 * wrfgpu_upload (the boundary arrays of a specified / nested domain among them), wrfgpu_start_domain, wrfgpu_set_dtbc and
 * wrfgpu_step one step at a time, wrfgpu_download of the shim's field list, and
 * wrfgpu_last_error on failure.  The Fortran shim cannot be compiled in the development image; this driver pins the
 * boundary the shim binds (include/wrfgpu.h) with a caller that is not Python.
 *
 *   abi_driver <state.bin> <nsteps> <out.bin>
 * state.bin (written by tests/test_abi.py): wrfgpu_dims, wrfgpu_config, int nfields, then per field char name[24],
 * long long n, float data[n].  out.bin: per downloaded field the same record.  Exit 3 + message when there is no device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/wrfgpu.h"

#define MAXF 384
static void die(const char *where) {
  fprintf(stderr, "abi_driver: %s: %s\n", where, wrfgpu_last_error());
  exit(2);
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "usage: abi_driver state.bin nsteps out.bin\n"); return 64; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 66; }
  wrfgpu_dims d;
  wrfgpu_config c;
  int nf = 0;
  if (fread(&d, sizeof d, 1, f) != 1 || fread(&c, sizeof c, 1, f) != 1 || fread(&nf, sizeof nf, 1, f) != 1 || nf > MAXF) return 65;
  static char name[MAXF][24];
  static float *data[MAXF];
  static long long count[MAXF];
  const char *names[MAXF];
  const float *host[MAXF];
  for (int i = 0; i < nf; i++) {
    if (fread(name[i], 24, 1, f) != 1 || fread(&count[i], sizeof(long long), 1, f) != 1) return 65;
    data[i] = (float *)malloc(sizeof(float) * (size_t)count[i]);
    if (fread(data[i], sizeof(float), (size_t)count[i], f) != (size_t)count[i]) return 65;
    names[i] = name[i];
    host[i] = data[i];
  }
  fclose(f);
  if (wrfgpu_device_count() < 1) {      /* the shim's wrf_error_fatal */
    fprintf(stderr, "abi_driver: no CUDA device on this node (there is no CPU fallback)\n");
    wrfgpu_solver *s = NULL;
    if (wrfgpu_create(&d, &c, 0, &s) == 0) { fprintf(stderr, "abi_driver: wrfgpu_create succeeded without a device\n"); return 1; }
    fprintf(stderr, "abi_driver: wrfgpu_create: %s\n", wrfgpu_last_error());
    return 3;
  }
  wrfgpu_solver *s = NULL;
  if (wrfgpu_create(&d, &c, 0 % wrfgpu_device_count(), &s)) die("wrfgpu_create");
  for (int i = 0; i < nf; i++)
    if (wrfgpu_field_size(s, names[i]) != (size_t)count[i]) { fprintf(stderr, "abi_driver: size of %s\n", names[i]); return 1; }
  if (wrfgpu_upload(s, nf, names, host)) die("wrfgpu_upload");
  if (wrfgpu_start_domain(s)) die("wrfgpu_start_domain");
  const int nsteps = atoi(argv[2]);
  float dtbc = 0.f;                      /* grid%dtbc: the boundary arrays above were just read */
  for (int n = 0; n < nsteps; n++) {
    if ((c.specified || c.nested) && wrfgpu_set_dtbc(s, dtbc)) die("wrfgpu_set_dtbc");   /* the shim, before every step */
    dtbc = dtbc + c.dt;                  /* solve_em.F:377 on the host copy */
    if (wrfgpu_step(s, 1)) die("wrfgpu_step");
  }
  /* the shim's download list (history / restart alarm) */
  const char *want[] = {"u_1", "u_2", "v_1", "v_2", "w_1", "w_2", "t_1", "t_2", "ph_1", "ph_2", "p", "al", "alt", "ww", "h_diabatic",
                        "moist_qv", "moist_qc", "moist_qr", "scalar_1", "scalar_2", "tke_1", "tke_2", "mu_1", "mu_2", "rainnc", "rainncv"};
  const char *dn[MAXF];
  float *dh[MAXF];
  size_t dc[MAXF];
  int nd = 0;
  for (unsigned i = 0; i < sizeof want / sizeof want[0]; i++) {
    size_t n = wrfgpu_field_size(s, want[i]);
    if (!n && (!strcmp(want[i], "tke_1") || !strcmp(want[i], "tke_2"))) {   /* Registry aliases of the tke species */
      /* wrfgpu_field_size only knows resident names; the alias is resolved by upload / download: ask through a download */
      continue;
    }
    if (!n) continue;
    dn[nd] = want[i]; dc[nd] = n; dh[nd] = (float *)malloc(sizeof(float) * n); nd++;
  }
  if (wrfgpu_download(s, nd, dn, dh)) die("wrfgpu_download");
  FILE *o = fopen(argv[3], "wb");
  if (!o) { perror(argv[3]); return 73; }
  fwrite(&nd, sizeof nd, 1, o);
  for (int i = 0; i < nd; i++) {
    char nm[24];
    memset(nm, 0, 24);
    strncpy(nm, dn[i], 23);
    long long n = (long long)dc[i];
    fwrite(nm, 24, 1, o); fwrite(&n, sizeof n, 1, o); fwrite(dh[i], sizeof(float), dc[i], o);
  }
  fclose(o);
  printf("abi_driver: %d fields up, %d steps, %d fields down, %lld launches\n", nf, nsteps, nd, wrfgpu_launch_count(s));
  wrfgpu_destroy(s);
  return 0;
}
