/*
 * cafcube_driver.c -- the step loop of `program cafcube` (CUBE/main/cafcube.f90:5-52) over the C ABI of libcubegpu.so, in plain C:
 * what the Fortran driver does once its five hot-path calls are replaced, without Python in between.
 *
 *   cafcube_driver <kernel_dir> <checkpoint_dir> <z_in> <nsteps> <dt> <a_mid> [z_out]
 *
 *   kernel_dir      wfxyzf.3.ascii and wfxyzc.2.ascii (CUBE/kernels, read like kernel_f.f90:14-23 / kernel_c.f90:30-38)
 *   checkpoint_dir  image1/<z>zip2_1.bin (168-byte sim_header + rhoc), <z>zip0_1.bin (xp), <z>zip1_1.bin (vp), <z>vfield_1.bin
 *                   (checkpoint.f90:33-70, file names parameters.f90:244-257); single image
 *   nsteps, dt, a_mid   a fixed-step replay of the call order (the adaptive controller `timestep` stays host code and is not
 *                   part of this example); the limits particle_mesh returns are printed every step
 *   z_out           if given, the final state is written back as a checkpoint at that redshift
 *
 * Build:  gcc -O2 -o cafcube_driver examples/cafcube_driver.c -Iinclude -Lcafproject_b200 -lcubegpu -lm -Wl,-rpath,$PWD/cafproject_b200
 * There is no CPU path: without a usable CUDA device cube_gpu_init fails and the driver stops with the library's message.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cube_gpu.h"

/* sim_header, parameters.f90:119-140: 13 x int64 then 16 x float32, no padding (168 bytes) */
#pragma pack(push, 1)
typedef struct {
  int64_t nplocal, izipx, izipv, image, nn, nnt, nt, ncell, ncb, istep, cur_checkpoint, cur_proj, cur_halo;
  float a, t, tau, dt_f_acc, dt_pp_acc, dt_c_acc, mass_p, box, h0, omega_m, omega_l, s8, vsim2phys, sigma_vres, sigma_vi, z_i;
} sim_header;
#pragma pack(pop)

static void die(const char *what) {
  fprintf(stderr, "cafcube_driver: %s\n", what);
  exit(1);
}
static void check(int rc) {  /* the reference's `print*` + `stop` convention */
  if (rc != 0) die(cube_gpu_last_error());
}
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) die("out of host memory");
  return p;
}
static void read_exact(const char *path, long offset, void *dst, size_t bytes) {
  FILE *f = fopen(path, "rb");
  if (!f) { perror(path); exit(1); }
  if (fseek(f, offset, SEEK_SET) != 0 || fread(dst, 1, bytes, f) != bytes) { fprintf(stderr, "cafcube_driver: short read of %s\n", path); exit(1); }
  fclose(f);
}
static void write_two(const char *path, const void *a, size_t na, const void *b, size_t nb) {
  FILE *f = fopen(path, "wb");
  if (!f) { perror(path); exit(1); }
  if ((na && fwrite(a, 1, na, f) != na) || (nb && fwrite(b, 1, nb, f) != nb)) { fprintf(stderr, "cafcube_driver: short write of %s\n", path); exit(1); }
  fclose(f);
}
/* write(str,'(f7.3)') z ; trim(adjustl(str))  (parameters.f90:213-219) */
static void z2str(double z, char *out) {
  char buf[32];
  snprintf(buf, sizeof buf, "%7.3f", z);
  const char *p = buf;
  while (*p == ' ') p++;
  strcpy(out, p);
}
static void ckpt_name(char *out, size_t n, const char *dir, double z, const char *zip) {
  char zs[32];
  z2str(z, zs);
  snprintf(out, n, "%s/image1/%s%s_1.bin", dir, zs, zip);
}
/* the ASCII kernel tables: rows "i j k fx fy fz", i fastest (kernel_f.f90:19) */
static void read_table(const char *path, int n, float *fx_fy_fz /* [k][j][i][3] */) {
  FILE *f = fopen(path, "r");
  if (!f) { perror(path); exit(1); }
  for (int q = 0; q < n * n * n; q++) {
    int i, j, k; double a, b, c;
    if (fscanf(f, "%d %d %d %lf %lf %lf", &i, &j, &k, &a, &b, &c) != 6) { fprintf(stderr, "cafcube_driver: bad row %d of %s\n", q, path); exit(1); }
    float *o = fx_fy_fz + 3 * ((((size_t)k - 1) * n + (j - 1)) * n + (i - 1));
    o[0] = (float)a; o[1] = (float)b; o[2] = (float)c;
  }
  fclose(f);
}

int main(int argc, char **argv) {
  if (argc < 7) die("usage: cafcube_driver <kernel_dir> <checkpoint_dir> <z_in> <nsteps> <dt> <a_mid> [z_out]");
  const char *kdir = argv[1], *cdir = argv[2];
  const double z_in = atof(argv[3]);
  const int nsteps = atoi(argv[4]);
  const float dt = (float)atof(argv[5]), a_mid = (float)atof(argv[6]);
  char path[1024];

  /* ---- particle_initialization.f90:11-64: header, then the four arrays --------------------- */
  sim_header hd;
  ckpt_name(path, sizeof path, cdir, z_in, "zip2");
  read_exact(path, 0, &hd, sizeof hd);
  if (hd.izipx != 2 || hd.izipv != 2) die("zip format incompatable");          /* particle_initialization.f90:14-18 */
  if (hd.nn != 1) die("this example drives a single image");
  const int nnt = (int)hd.nnt, nt = (int)hd.nt, nc = nnt * nt;
  const size_t ncell = (size_t)nc * nc * nc, np = (size_t)hd.nplocal;
  int32_t *rhoc = xmalloc(4 * ncell);
  float *vfield = xmalloc(12 * ncell);
  int16_t *xp = xmalloc(6 * np), *vp = xmalloc(6 * np);
  read_exact(path, sizeof hd, rhoc, 4 * ncell);
  ckpt_name(path, sizeof path, cdir, z_in, "vfield"); read_exact(path, 0, vfield, 12 * ncell);
  ckpt_name(path, sizeof path, cdir, z_in, "zip0");   read_exact(path, 0, xp, 6 * np);
  ckpt_name(path, sizeof path, cdir, z_in, "zip1");   read_exact(path, 0, vp, 6 * np);
  /* -DPID runs also have <z>zipid_1.bin, integer(8) IDs (particle_initialization.f90:56-59); optional here */
  int64_t *pid = NULL;
  ckpt_name(path, sizeof path, cdir, z_in, "zipid");
  { FILE *f = fopen(path, "rb"); if (f) { fclose(f); pid = xmalloc(8 * np); read_exact(path, 0, pid, 8 * np); } }

  /* ---- initialize.f90: kernel tables; the host's own tanf table (pm.f90:102) ---------------- */
  float *fk_kji = xmalloc(sizeof(float) * 16 * 16 * 16 * 3), *fk = xmalloc(sizeof(float) * 3 * 16 * 16 * 16);
  float *ck_kji = xmalloc(sizeof(float) * 4 * 4 * 4 * 3), *ck = xmalloc(sizeof(float) * 3 * 4 * 4 * 4);
  snprintf(path, sizeof path, "%s/wfxyzf.3.ascii", kdir); read_table(path, 16, fk_kji);
  snprintf(path, sizeof path, "%s/wfxyzc.2.ascii", kdir); read_table(path, 4, ck_kji);
  for (int d = 0; d < 3; d++)          /* fk_table(i,j,k,dim): dim slowest */
    for (size_t q = 0; q < 16 * 16 * 16; q++) fk[(size_t)d * 4096 + q] = fk_kji[3 * q + d];
  memcpy(ck, ck_kji, sizeof(float) * 192); /* ck_table(dim,i,j,k): dim fastest = the row layout */
  float *lut = xmalloc(sizeof(float) * 65536);
  const float pi = 4.0f * atanf(1.0f);
  for (int u = 0; u < 65536; u++) lut[u] = tanf((pi * (float)(int16_t)(uint16_t)u) / 65535.0f);

  cube_params p;
  memset(&p, 0, sizeof p);
  p.nn[0] = p.nn[1] = p.nn[2] = 1;
  p.nnt = nnt; p.nc = nc; p.ncell = 4; p.ncb = 6; p.izipx = p.izipv = 2;
  p.np_nc = (int)lround(cbrt((double)np / (double)ncell));
  if (p.np_nc < 1) p.np_nc = 1;
  p.image_buffer = 1.5f; p.tile_buffer = 2.5f;                                 /* parameters.f90:59-60 */
  cube_handle *h = NULL;
  check(cube_gpu_init(&p, fk, ck, lut, NULL, &h));
  check(cube_gpu_upload(h, xp, vp, rhoc, vfield, (int64_t)np, (int64_t)np, hd.sigma_vi));
  if (pid) check(cube_gpu_upload_pid(h, pid));

  /* ---- cafcube.f90:16-20 then the loop :25-46 ---------------------------------------------- */
  float ovh_image = 0.f;
  check(cube_gpu_buffer(h, 1, 0, 0, &ovh_image));   /* call buffer_density */
  check(cube_gpu_buffer(h, 0, 1, 0, NULL));         /* call buffer_x       */
  check(cube_gpu_buffer(h, 0, 0, 1, NULL));         /* call buffer_v       */
  float dt_old = 0.f;
  for (int istep = 1; istep <= nsteps; istep++) {
    int64_t nplocal = 0; float sigma_new = 0.f, ovh_tile = 0.f; double std_vsim[3];
    float dt_fine = 0.f, dt_coarse = 0.f, dt_vmax = 0.f, vmax = 0.f;
    check(cube_gpu_update_x(h, dt_old, dt, &nplocal, &sigma_new, std_vsim, &ovh_tile));   /* call update_particle */
    check(cube_gpu_buffer(h, 1, 0, 0, &ovh_image));                                       /* call buffer_density  */
    check(cube_gpu_buffer(h, 0, 1, 0, NULL));                                             /* call buffer_x        */
    check(cube_gpu_particle_mesh(h, a_mid, dt, &dt_fine, &dt_coarse, &dt_vmax, &vmax));   /* call particle_mesh   */
    check(cube_gpu_buffer(h, 0, 0, 1, NULL));                                             /* call buffer_v        */
    printf("step %d: nplocal %lld sigma_vi %g overhead_tile %g overhead_image %g dt_fine %g dt_coarse %g dt_vmax %g\n", istep,
           (long long)nplocal, sigma_new, ovh_tile, ovh_image, dt_fine, dt_coarse, dt_vmax);
    dt_old = dt;
  }

  /* ---- checkpoint.f90:33-70 ---------------------------------------------------------------- */
  int64_t nplocal = 0; float sigma = 0.f;
  check(cube_gpu_download(h, NULL, NULL, NULL, NULL, &nplocal, &sigma));
  if ((size_t)nplocal > np) { free(xp); free(vp); xp = xmalloc(6 * (size_t)nplocal); vp = xmalloc(6 * (size_t)nplocal); }
  check(cube_gpu_download(h, xp, vp, rhoc, vfield, &nplocal, &sigma));
  if (pid) {
    if ((size_t)nplocal > np) { free(pid); pid = xmalloc(8 * (size_t)nplocal); }
    check(cube_gpu_download_pid(h, pid));
  }
  if (argc > 7) {
    const double z_out = atof(argv[7]);
    hd.nplocal = nplocal; hd.sigma_vi = sigma; hd.istep += nsteps;
    ckpt_name(path, sizeof path, cdir, z_out, "zip2");   write_two(path, &hd, sizeof hd, rhoc, 4 * ncell);
    ckpt_name(path, sizeof path, cdir, z_out, "vfield"); write_two(path, vfield, 12 * ncell, NULL, 0);
    ckpt_name(path, sizeof path, cdir, z_out, "zip0");   write_two(path, xp, 6 * (size_t)nplocal, NULL, 0);
    ckpt_name(path, sizeof path, cdir, z_out, "zip1");   write_two(path, vp, 6 * (size_t)nplocal, NULL, 0);
    if (pid) { ckpt_name(path, sizeof path, cdir, z_out, "zipid"); write_two(path, pid, 8 * (size_t)nplocal, NULL, 0); }
  }
  printf("done: %lld particles, sigma_vi %g\n", (long long)nplocal, sigma);
  check(cube_gpu_finalize(h));
  free(xp); free(vp); free(pid); free(rhoc); free(vfield); free(fk); free(fk_kji); free(ck); free(ck_kji); free(lut);
  return 0;
}
