/*
 * cube_oracle.c -- CPU restatement of CUBE's per-timestep particle-mesh step.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle and the CPU baseline
 * ("port") of the reference; the product (cafproject_b200/csrc) never links,
 * imports or calls it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * PARITY STATUS: *unpinned by the reference*.  yuhaoran/cafproject ships no golden
 * vectors, no known-answer tests and cannot be compiled here (no Fortran compiler,
 * no FFTW, no coarray runtime).  The oracle is a line-by-line restatement of
 *   CUBE/main/update_particle.f90, buffer_density.f90, buffer_x.f90, buffer_v.f90,
 *   pm.f90, variables.f90 (cumsum3/cumsum6), parameters.f90
 * with Fortran's type-promotion rules applied by hand (default real = f32, default
 * integer = i32, `pi` is f32, mixed expressions promote per binary operation, no
 * FMA contraction: build with -ffp-contract=off).  The weak pins that do exist
 * (paper decode example ms_caf.tex:78, kernel tables, conservation invariants) are
 * exercised in tests/test_oracle_pins.py.
 *
 * Zip formats: izipx, izipv in {1,2} bytes per code (CUBE/main/universe*.fh:2-3), chosen per run with oracle_set_zip;
 * the default is x2v2, the format the GPU library is built for.
 *
 * Generalisation beyond the reference: the image grid is (nnx,nny,nnz) instead of
 * nn^3 so that 2- and 4-GPU weak-scaling points exist; with nnx=nny=nnz=nn it is the
 * reference's geometry (parameters.f90:178-203).  All images of a run live in one
 * process; a coarray GET `a(..)[img]` is a read of images[img].  `sync all` phases
 * are respected by looping over images inside each phase.
 *
 * FFTs are not done here: the Python driver (oracle/cube_oracle.py) performs them with
 * scipy.fft (pocketfft, f32) between the deposit and kick calls below.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int16_t i16;
typedef int32_t i32;
typedef int64_t i64;

/* parameters.f90:73  real,parameter :: pi=4*atan(1.)  -> f32 */
static const float PI_F = 3.14159274101257324f; /* 0x40490FDB */

typedef struct {
  i64 nn[3];       /* images per dim */
  i64 nnt, nc, nt; /* tiles/image/dim, coarse cells/image/dim, coarse cells/tile/dim */
  i64 ncell, ncb;  /* 4, 6 */
  i64 nte, nft, nfb, nfe;
  i64 np_image_max, np_tile_max;
  i64 nimg;
  i64 izipx, izipv; /* bytes per position / velocity code: 1 or 2 (universe*.fh) */
  i64 nvbin;        /* 2^(8*izipv)        parameters.f90:13 */
  i64 ishift;       /* -2^(8*izipx-1)     parameters.f90:14 */
  double x_resolution; /* 1/2^(8*izipx) */
} geom_t;

typedef struct {
  i16 *xp, *vp;   /* (3, np_image_max) */
  i64 *pid;       /* (np_image_max) particle IDs, -DPID (CUBE/main variables.f90:44 integer(8); CUBEnu variables.f90:47 integer(4)) */
  i32 *rhoc;      /* (nte,nte,nte,nnt,nnt,nnt), lower bound 1-ncb */
  float *vfield;  /* (3,nte,nte,nte,nnt,nnt,nnt) */
  i64 *cum;       /* same shape as rhoc */
  i64 nplocal;
  i64 icx, icy, icz, inx, iny, inz, ipx, ipy, ipz; /* 1-based, parameters.f90:181-194 */
  double std_vsim, std_vsim_c, std_vsim_res;
  float overhead_tile, overhead_image;
  float vmax;
  float vmax3[3]; /* CUBEnu pm.f90:398: vmax(3)=max(vmax,abs(v)) per component */
  float f2_max_coarse;
  float *f2_max_fine; /* (nnt,nnt,nnt) */
} image_t;

typedef struct {
  geom_t g;
  image_t *im;
  float sigma_vi, sigma_vi_new;
  float mass_p;
  i64 npglobal;
  int error; /* set instead of Fortran `stop` */
  i64 nlayer; /* 1 = CUBE/main source order; >1 = CUBEnu's colour passes over k (CUBEnu update_particle.f90:37,55-58) */
  char errmsg[256];
  /* scratch of update_particle (update_particle.f90:8-12) */
  i32 *rhoce, *rholocal;
  float *vfield_new;
  i64 *cume;
  i16 *xp_new, *vp_new;
  i64 *pid_new;
  int has_pid; /* -DPID: IDs ride with vp through buffer_density, buffer_v and update_particle */
} ctx_t;

/* ---- Fortran-style index helpers ------------------------------------------------ */
/* rhoc(i,j,k,itx,ity,itz), i,j,k in 1-ncb..nt+ncb ; it* in 1..nnt */
static inline i64 RH(const geom_t *g, i64 i, i64 j, i64 k, i64 tx, i64 ty, i64 tz) {
  const i64 b = g->ncb - 1, e = g->nte, t = g->nnt;
  return (i + b) + e * ((j + b) + e * ((k + b) + e * ((tx - 1) + t * ((ty - 1) + t * (tz - 1)))));
}
/* rhoce(i,j,k), i,j,k in 1-2ncb..nt+2ncb */
static inline i64 RE(const geom_t *g, i64 i, i64 j, i64 k) {
  const i64 b = 2 * g->ncb - 1, e = g->nt + 4 * g->ncb;
  return (i + b) + e * ((j + b) + e * (k + b));
}
static inline i64 image1d(const geom_t *g, i64 cx, i64 cy, i64 cz) {
  /* parameters.f90:200-203 generalised to a non-cubic image grid */
  return (cx - 1) + g->nn[0] * (cy - 1) + g->nn[0] * g->nn[1] * (cz - 1); /* 0-based here */
}
static inline i64 modulo_i(i64 a, i64 n) { i64 r = a % n; return r < 0 ? r + n : r; }

/* ---- codes -------------------------------------------------------- */
/* Zip formats (parameters.f90:10-15, universe*.fh): izipx, izipv in {1,2} bytes.  The oracle keeps every code in an int16
 * slot; a 1-byte code sits there sign-extended and each site that Fortran evaluates in kind=izipx / kind=izipv wraps to
 * that kind explicitly (to_kind), so the x2v2 arithmetic is unchanged and x1/v1 follow the same source lines. */
static inline i16 to_kind(i64 v, i64 izip) { return izip == 1 ? (i16)(int8_t)(uint8_t)v : (i16)(uint16_t)v; }
/* int(xp+ishift,izipx)+rshift   with ishift=-2^(8*izipx-1), rshift=0.5-ishift  (parameters.f90:14-15) */
static inline double xp_decode(const geom_t *g, i16 xp) {
  const i64 ishift = g->ishift;
  const double rshift = 0.5 - (double)ishift; /* 32768.5 (x2), 128.5 (x1) */
  i16 t = to_kind((i64)xp + ishift, g->izipx); /* int(...,izipx): wraps */
  return (double)t + rshift;
}
/* tan((pi*real(vp))/real(nvbin-1))  -- all f32, libm tanf */
static inline float vp_tan(const geom_t *g, i16 vp) { return tanf((PI_F * (float)vp) / (float)(g->nvbin - 1)); }
/* sqrt(pi/2)/(sigma_vi*vrel_boost)  -> f64   (vrel_boost is real(8)=2.5, parameters.f90:103) */
static inline double vscale(float sigma) { return (double)sqrtf(PI_F / 2) / ((double)sigma * 2.5); }
/* nint(real(nvbin-1)*atan(S*v)/pi,kind=izipv) */
static inline i16 vp_encode(const geom_t *g, double v, double S) {
  return to_kind(llround((double)(float)(g->nvbin - 1) * atan(S * v) / (double)PI_F), g->izipv);
}

/* ---- lifecycle --------------------------------------------------------------------- */
ctx_t *oracle_create(i64 nnx, i64 nny, i64 nnz, i64 nnt, i64 nc, i64 np_nc,
                     float image_buffer, float tile_buffer) {
  ctx_t *c = (ctx_t *)calloc(1, sizeof(ctx_t));
  geom_t *g = &c->g;
  g->nn[0] = nnx; g->nn[1] = nny; g->nn[2] = nnz;
  g->nnt = nnt; g->nc = nc; g->nt = nc / nnt;
  g->ncell = 4; g->ncb = 6;
  g->nte = g->nt + 2 * g->ncb;
  g->nft = g->nt * g->ncell;
  g->nfb = g->ncb * g->ncell;
  g->nfe = g->nft + 2 * g->nfb;
  g->izipx = g->izipv = 2; g->nvbin = 65536; g->ishift = -32768; g->x_resolution = 1.0 / 65536.0; /* oracle_set_zip changes them */
  g->nimg = nnx * nny * nnz;
  /* variables.f90:7-9 (real(4) arithmetic, truncated to integer(8)) */
  i64 np_image = (nc * np_nc) * (nc * np_nc) * (nc * np_nc);
  float r = ((float)g->nte * 1.f) / (float)g->nt;
  float r3 = r * r * r;
  g->np_image_max = (i64)((float)np_image * r3 * image_buffer);
  g->np_tile_max = (i64)((float)(np_image / (nnt * nnt * nnt)) * r3 * tile_buffer);
  c->im = (image_t *)calloc(g->nimg, sizeof(image_t));
  i64 ncell_e = g->nte * g->nte * g->nte * nnt * nnt * nnt;
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    im->xp = (i16 *)calloc(3 * g->np_image_max, sizeof(i16));
    im->vp = (i16 *)calloc(3 * g->np_image_max, sizeof(i16));
    im->pid = NULL; /* allocated by oracle_load_pid */
    im->rhoc = (i32 *)calloc(ncell_e, sizeof(i32));
    im->vfield = (float *)calloc(3 * ncell_e, sizeof(float));
    im->cum = (i64 *)calloc(ncell_e, sizeof(i64));
    im->f2_max_fine = (float *)calloc(nnt * nnt * nnt, sizeof(float));
    /* parameters.f90:180-194 */
    i64 rank = m;
    im->icz = rank / (nnx * nny) + 1;
    im->icy = (rank - nnx * nny * (im->icz - 1)) / nnx + 1;
    im->icx = rank % nnx + 1;
    im->inx = modulo_i(im->icx - 2, nnx) + 1;
    im->iny = modulo_i(im->icy - 2, nny) + 1;
    im->inz = modulo_i(im->icz - 2, nnz) + 1;
    im->ipx = modulo_i(im->icx, nnx) + 1;
    im->ipy = modulo_i(im->icy, nny) + 1;
    im->ipz = modulo_i(im->icz, nnz) + 1;
  }
  i64 ne2 = g->nt + 4 * g->ncb, n2 = ne2 * ne2 * ne2;
  c->rhoce = (i32 *)calloc(n2, sizeof(i32));
  c->rholocal = (i32 *)calloc(n2, sizeof(i32));
  c->vfield_new = (float *)calloc(3 * n2, sizeof(float));
  c->cume = (i64 *)calloc(n2, sizeof(i64));
  c->xp_new = (i16 *)calloc(3 * g->np_tile_max, sizeof(i16));
  c->vp_new = (i16 *)calloc(3 * g->np_tile_max, sizeof(i16));
  c->pid_new = NULL;
  return c;
}

void oracle_destroy(ctx_t *c) {
  for (i64 m = 0; m < c->g.nimg; m++) {
    image_t *im = &c->im[m];
    free(im->xp); free(im->vp); free(im->pid); free(im->rhoc); free(im->vfield); free(im->cum); free(im->f2_max_fine);
  }
  free(c->im); free(c->rhoce); free(c->rholocal); free(c->vfield_new); free(c->cume);
  free(c->xp_new); free(c->vp_new); free(c->pid_new); free(c);
}

/* accessors for the Python side */
i64 oracle_np_image_max(ctx_t *c) { return c->g.np_image_max; }
i64 oracle_np_tile_max(ctx_t *c) { return c->g.np_tile_max; }
i16 *oracle_xp(ctx_t *c, i64 m) { return c->im[m].xp; }
i16 *oracle_vp(ctx_t *c, i64 m) { return c->im[m].vp; }
i64 *oracle_pid(ctx_t *c, i64 m) { return c->im[m].pid; }
i32 *oracle_rhoc(ctx_t *c, i64 m) { return c->im[m].rhoc; }
float *oracle_vfield(ctx_t *c, i64 m) { return c->im[m].vfield; }
i64 *oracle_cum(ctx_t *c, i64 m) { return c->im[m].cum; }
i64 oracle_nplocal(ctx_t *c, i64 m) { return c->im[m].nplocal; }
float *oracle_f2_max_fine(ctx_t *c, i64 m) { return c->im[m].f2_max_fine; }
float oracle_get_sigma_vi(ctx_t *c) { return c->sigma_vi; }
float oracle_get_sigma_vi_new(ctx_t *c) { return c->sigma_vi_new; }
void oracle_set_sigma_vi(ctx_t *c, float s) { c->sigma_vi = s; }
void oracle_set_sigma_vi_new(ctx_t *c, float s) { c->sigma_vi_new = s; }
float oracle_mass_p(ctx_t *c) { return c->mass_p; }
/* a species of a multi-species run carries its own particle mass (CUBEnu pm.f90:79-99 deposits mass_p_cdm / mass_p_nu) */
void oracle_set_mass_p(ctx_t *c, float mp) { c->mass_p = mp; }
i64 oracle_npglobal(ctx_t *c) { return c->npglobal; }
int oracle_error(ctx_t *c) { return c->error; }
const char *oracle_errmsg(ctx_t *c) { return c->errmsg; }
double oracle_std_vsim(ctx_t *c, int which) {
  return which == 0 ? c->im[0].std_vsim : which == 1 ? c->im[0].std_vsim_c : c->im[0].std_vsim_res;
}
float oracle_overhead_tile(ctx_t *c) { return c->im[0].overhead_tile; }
float oracle_overhead_image(ctx_t *c) { return c->im[0].overhead_image; }
float oracle_vmax(ctx_t *c, i64 m) { return c->im[m].vmax; }
float oracle_vmax3(ctx_t *c, i64 m, int d) { return c->im[m].vmax3[d]; }
/* CUBEnu update_particle.f90:37: nlayer=2*ceiling(dt_mid*sim%vz_max/ncell)+1 (f32 product and quotient); 0 or 1 = CUBE/main */
void oracle_set_nlayer(ctx_t *c, i64 nlayer) { c->nlayer = nlayer; }
i64 oracle_nlayer_for(ctx_t *c, float dt_old, float dt, float vz_max) {
  const float dt_mid = (dt_old + dt) / 2;
  return 2 * (i64)ceilf(dt_mid * vz_max / (float)c->g.ncell) + 1;
}

/* host-libm tanf table, indexed by the raw 16-bit code (SURVEY sec. 7 hard part 1) */
void oracle_tanf_lut(float *lut) {
  geom_t g2; g2.nvbin = 65536;
  for (int u = 0; u < 65536; u++) lut[u] = vp_tan(&g2, (i16)(uint16_t)u);
}
/* the same for either velocity format: 2^(8*izipv) entries indexed by the raw unsigned code */
void oracle_tanf_lut_zip(float *lut, i64 izipv) {
  geom_t g2; g2.nvbin = (i64)1 << (8 * izipv);
  for (i64 u = 0; u < g2.nvbin; u++) lut[u] = vp_tan(&g2, to_kind(u, izipv));
}
/* choose the zip formats of a run (universe*.fh:2-3); call before oracle_load_image */
int oracle_set_zip(ctx_t *c, i64 izipx, i64 izipv) {
  if ((izipx != 1 && izipx != 2) || (izipv != 1 && izipv != 2)) return 1;
  geom_t *g = &c->g;
  g->izipx = izipx; g->izipv = izipv;
  g->nvbin = (i64)1 << (8 * izipv);
  g->ishift = -((i64)1 << (8 * izipx - 1));
  g->x_resolution = (double)(1.0f / (float)((i64)1 << (8 * izipx)));
  return 0;
}

/* probes of the code arithmetic for the pin tests: xq of update_particle.f90:41 for coarse cell index `cell` (1-based),
 * the decoded velocity of :42 (without vfield) and the encoder of :86 */
double oracle_probe_xq(ctx_t *c, i64 cell, i64 code) {
  const geom_t *g = &c->g;
  return ((double)cell - 1.0) + xp_decode(g, to_kind(code, g->izipx)) * g->x_resolution;
}
double oracle_probe_vdecode(ctx_t *c, i64 code, float sigma) {
  const geom_t *g = &c->g;
  return (double)vp_tan(g, to_kind(code, g->izipv)) / vscale(sigma);
}
i64 oracle_probe_vencode(ctx_t *c, double v, float sigma) { return (i64)vp_encode(&c->g, v, vscale(sigma)); }

/* ---- cumsum (variables.f90:74-110) ------------------------------------------------- */
static void cumsum6(const geom_t *g, const i32 *rho, i64 *cum) {
  i64 n = g->nte * g->nte * g->nte * g->nnt * g->nnt * g->nnt, nsum = 0;
  for (i64 q = 0; q < n; q++) { nsum += rho[q]; cum[q] = nsum; } /* storage order == loop order */
}
static void cumsum3(const geom_t *g, const i32 *rho, i64 *cum) {
  i64 e = g->nt + 4 * g->ncb, n = e * e * e, nsum = 0;
  for (i64 q = 0; q < n; q++) { nsum += rho[q]; cum[q] = nsum; }
}

/* ---- particle_initialization.f90:11-72 (disjoint state in, from caller arrays) ------ */
void oracle_load_image(ctx_t *c, i64 m, const i16 *xp, const i16 *vp, const i32 *rhoc_phys,
                       const float *vfield_phys, i64 nplocal) {
  const geom_t *g = &c->g;
  image_t *im = &c->im[m];
  i64 ncell_e = g->nte * g->nte * g->nte * g->nnt * g->nnt * g->nnt;
  memset(im->rhoc, 0, ncell_e * sizeof(i32));
  memset(im->vfield, 0, 3 * ncell_e * sizeof(float));
  memset(im->xp, 0, 3 * g->np_image_max * sizeof(i16));
  memset(im->vp, 0, 3 * g->np_image_max * sizeof(i16));
  i64 q = 0;
  for (i64 tz = 1; tz <= g->nnt; tz++) for (i64 ty = 1; ty <= g->nnt; ty++) for (i64 tx = 1; tx <= g->nnt; tx++)
    for (i64 k = 1; k <= g->nt; k++) for (i64 j = 1; j <= g->nt; j++) for (i64 i = 1; i <= g->nt; i++, q++) {
      i64 r = RH(g, i, j, k, tx, ty, tz);
      im->rhoc[r] = rhoc_phys[q];
      for (int d = 0; d < 3; d++) im->vfield[3 * r + d] = vfield_phys[3 * q + d];
    }
  memcpy(im->xp, xp, 3 * nplocal * sizeof(i16));
  memcpy(im->vp, vp, 3 * nplocal * sizeof(i16));
  im->nplocal = nplocal;
}
/* -DPID (particle_initialization.f90, `read(14) pid(:nplocal)`): IDs of the nplocal particles just loaded, file order.
 * Every image of the run must be given IDs before oracle_finish_load. */
void oracle_load_pid(ctx_t *c, i64 m, const i64 *pid) {
  const geom_t *g = &c->g;
  image_t *im = &c->im[m];
  if (!im->pid) im->pid = (i64 *)calloc(g->np_image_max, sizeof(i64));
  if (!c->pid_new) c->pid_new = (i64 *)calloc(g->np_tile_max, sizeof(i64));
  memset(im->pid, 0, g->np_image_max * sizeof(i64));
  memcpy(im->pid, pid, im->nplocal * sizeof(i64));
  c->has_pid = 1;
}
/* particle_initialization.f90:65-72: npglobal, mass_p = real((nf*nn)**3)/npglobal */
void oracle_finish_load(ctx_t *c, float sigma_vi) {
  const geom_t *g = &c->g;
  c->npglobal = 0;
  for (i64 m = 0; m < g->nimg; m++) c->npglobal += c->im[m].nplocal;
  /* generalised: nf_global^3 -> product over dims of nf*nn_d */
  i64 nf = g->nc * g->ncell;
  c->mass_p = (float)((nf * g->nn[0]) * (nf * g->nn[1]) * (nf * g->nn[2])) / (float)c->npglobal;
  c->sigma_vi = sigma_vi;
  c->sigma_vi_new = sigma_vi;
}
/* checkpoint.f90:35,40: the physical sub-blocks as written to zip2/vfield */
void oracle_store_image(ctx_t *c, i64 m, i32 *rhoc_phys, float *vfield_phys) {
  const geom_t *g = &c->g;
  image_t *im = &c->im[m];
  i64 q = 0;
  for (i64 tz = 1; tz <= g->nnt; tz++) for (i64 ty = 1; ty <= g->nnt; ty++) for (i64 tx = 1; tx <= g->nnt; tx++)
    for (i64 k = 1; k <= g->nt; k++) for (i64 j = 1; j <= g->nt; j++) for (i64 i = 1; i <= g->nt; i++, q++) {
      i64 r = RH(g, i, j, k, tx, ty, tz);
      rhoc_phys[q] = im->rhoc[r];
      for (int d = 0; d < 3; d++) vfield_phys[3 * q + d] = im->vfield[3 * r + d];
    }
}

/* ---- buffer_density.f90 ------------------------------------------------------------ */
/* copy a block of ghost layers of rhoc and vfield: dst cells (di..,dj..,dk..) of tile (dtx,dty,dtz) of
   image md <- src cells of tile (stx,sty,stz) of image ms; block extents (ni,nj,nk). */
static void halo_copy(ctx_t *c, i64 md, i64 ms, i64 di, i64 dj, i64 dk, i64 dtx, i64 dty, i64 dtz,
                      i64 si, i64 sj, i64 sk, i64 stx, i64 sty, i64 stz, i64 ni, i64 nj, i64 nk) {
  const geom_t *g = &c->g;
  image_t *D = &c->im[md], *S = &c->im[ms];
  for (i64 k = 0; k < nk; k++) for (i64 j = 0; j < nj; j++) for (i64 i = 0; i < ni; i++) {
    i64 rd = RH(g, di + i, dj + j, dk + k, dtx, dty, dtz), rs = RH(g, si + i, sj + j, sk + k, stx, sty, stz);
    D->rhoc[rd] = S->rhoc[rs];
    for (int d = 0; d < 3; d++) D->vfield[3 * rd + d] = S->vfield[3 * rs + d];
  }
}

void oracle_buffer_density(ctx_t *c) {
  const geom_t *g = &c->g;
  const i64 nt = g->nt, ncb = g->ncb, nnt = g->nnt, nte = g->nte, lo = 1 - ncb;
  /* x (buffer_density.f90:11-26): only physical y,z rows are meaningful yet, but the reference copies
     the full (:,:) extent, so do we.  Within one phase sources (physical-in-that-dim layers) and
     destinations (ghost layers in that dim) are disjoint, so image order does not matter. */
  for (int dim = 0; dim < 3; dim++) {
    for (i64 m = 0; m < g->nimg; m++) {
      image_t *im = &c->im[m];
      i64 mneg, mpos;
      if (dim == 0) { mneg = image1d(g, im->inx, im->icy, im->icz); mpos = image1d(g, im->ipx, im->icy, im->icz); }
      else if (dim == 1) { mneg = image1d(g, im->icx, im->iny, im->icz); mpos = image1d(g, im->icx, im->ipy, im->icz); }
      else { mneg = image1d(g, im->icx, im->icy, im->inz); mpos = image1d(g, im->icx, im->icy, im->ipz); }
      for (i64 tb = 1; tb <= nnt; tb++) for (i64 ta = 1; ta <= nnt; ta++) for (i64 t = 1; t <= nnt; t++) {
        /* tile index along `dim` is t; the other two are (ta,tb) */
        i64 T[3], Tm[3], Tp[3];
        if (dim == 0) { T[0] = t; T[1] = ta; T[2] = tb; }
        else if (dim == 1) { T[0] = ta; T[1] = t; T[2] = tb; }
        else { T[0] = ta; T[1] = tb; T[2] = t; }
        memcpy(Tm, T, sizeof T); memcpy(Tp, T, sizeof T);
        i64 srcm_img = m, srcp_img = m;
        if (t == 1) { Tm[dim] = nnt; srcm_img = mneg; } else Tm[dim] = t - 1;
        if (t == nnt) { Tp[dim] = 1; srcp_img = mpos; } else Tp[dim] = t + 1;
        i64 d0[3] = {lo, lo, lo}, s0[3] = {lo, lo, lo}, n[3] = {nte, nte, nte};
        /* ghost layer (:0) <- (nt-ncb+1:nt) of the lower neighbour */
        d0[dim] = lo; s0[dim] = nt - ncb + 1; n[dim] = ncb;
        halo_copy(c, m, srcm_img, d0[0], d0[1], d0[2], T[0], T[1], T[2], s0[0], s0[1], s0[2], Tm[0], Tm[1], Tm[2], n[0], n[1], n[2]);
        /* ghost layer (nt+1:) <- (1:ncb) of the upper neighbour */
        d0[dim] = nt + 1; s0[dim] = 1;
        halo_copy(c, m, srcp_img, d0[0], d0[1], d0[2], T[0], T[1], T[2], s0[0], s0[1], s0[2], Tp[0], Tp[1], Tp[2], n[0], n[1], n[2]);
      }
    }
  }
  /* buffer_density.f90:75-93 */
  float ovh = 0;
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    i64 n = nte * nte * nte * nnt * nnt * nnt, s = 0;
    for (i64 q = 0; q < n; q++) s += im->rhoc[q];
    im->overhead_image = (float)((double)s / (double)g->np_image_max);
    if (im->overhead_image > ovh) ovh = im->overhead_image;
    if ((double)im->overhead_image > 1.0) {
      c->error = 2;
      snprintf(c->errmsg, sizeof c->errmsg, "error: too many particles in this image+buffer: %lld > %lld on image %lld; please set image_buffer larger",
               (long long)s, (long long)g->np_image_max, (long long)(m + 1));
      return;
    }
  }
  for (i64 m = 0; m < g->nimg; m++) c->im[m].overhead_image = ovh;
  /* buffer_density.f90:97-141: move particles to the top, then row by row to their cum slots */
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    i64 nshift = g->np_image_max - im->nplocal;
    memmove(im->xp + 3 * nshift, im->xp, 3 * im->nplocal * sizeof(i16));
    memmove(im->vp + 3 * nshift, im->vp, 3 * im->nplocal * sizeof(i16));
    memset(im->xp, 0, 3 * nshift * sizeof(i16));
    memset(im->vp, 0, 3 * nshift * sizeof(i16));
    if (c->has_pid) { /* buffer_density.f90:111-114 */
      memmove(im->pid + nshift, im->pid, im->nplocal * sizeof(i64));
      memset(im->pid, 0, nshift * sizeof(i64));
    }
    cumsum6(g, im->rhoc, im->cum);
    i64 ifrom = nshift;
    for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++)
      for (i64 iz = 1; iz <= nt; iz++) for (i64 iy = 1; iy <= nt; iy++) {
        i64 nlast = im->cum[RH(g, nt, iy, iz, tx, ty, tz)];
        i64 nlen = nlast - im->cum[RH(g, 0, iy, iz, tx, ty, tz)];
        /* xp(:,nlast-nlen+1:nlast)=xp(:,ifrom+1:ifrom+nlen) ; then zero the source.  Fortran array
           assignment has copy semantics for overlap -> memmove; the zeroing comes after the copy. */
        memmove(im->xp + 3 * (nlast - nlen), im->xp + 3 * ifrom, 3 * nlen * sizeof(i16));
        memmove(im->vp + 3 * (nlast - nlen), im->vp + 3 * ifrom, 3 * nlen * sizeof(i16));
        /* xp(:,ifrom+1:ifrom+nlen)=0 -- literal; if source and destination ever overlapped (image buffer
           nearly full) the reference would destroy data here and report it through its checksum. */
        memset(im->xp + 3 * ifrom, 0, 3 * nlen * sizeof(i16));
        memset(im->vp + 3 * ifrom, 0, 3 * nlen * sizeof(i16));
        if (c->has_pid) { /* buffer_density.f90:132-135 */
          memmove(im->pid + (nlast - nlen), im->pid + ifrom, nlen * sizeof(i64));
          memset(im->pid + ifrom, 0, nlen * sizeof(i64));
        }
        ifrom += nlen;
      }
  }
}

/* ---- buffer_x.f90 / buffer_v.f90: identical index logic, different array --------------- */
static void buffer_particles(ctx_t *c, int which /*0=xp,1=vp*/) {
  const geom_t *g = &c->g;
  const i64 nt = g->nt, ncb = g->ncb, nnt = g->nnt, lo = 1 - ncb, hi = nt + ncb;
#define ARR(im_) (which == 0 ? (im_)->xp : (im_)->vp)
  /* staged copies so that every "remote GET" of a phase sees pre-phase data (all images are in one
     process here).  Sources and destinations of one phase are disjoint index ranges in the reference
     (ghost slots vs. slots filled in earlier phases), so in-place memcpy is equivalent. */
  /* x- then x+ (buffer_x.f90:12-75) */
  for (int side = 0; side < 2; side++)
    for (i64 m = 0; m < g->nimg; m++) {
      image_t *im = &c->im[m];
      for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++)
        for (i64 iz = 1; iz <= nt; iz++) for (i64 iy = 1; iy <= nt; iy++) {
          i64 nlast, nlen, mlast; image_t *src = im;
          if (side == 0) {
            nlast = im->cum[RH(g, 0, iy, iz, tx, ty, tz)];
            nlen = nlast - im->cum[RH(g, lo, iy, iz, tx, ty, tz)] + im->rhoc[RH(g, lo, iy, iz, tx, ty, tz)];
            if (tx == 1) { src = &c->im[image1d(g, im->inx, im->icy, im->icz)]; mlast = src->cum[RH(g, nt, iy, iz, nnt, ty, tz)]; }
            else mlast = im->cum[RH(g, nt, iy, iz, tx - 1, ty, tz)];
          } else {
            nlast = im->cum[RH(g, hi, iy, iz, tx, ty, tz)];
            nlen = nlast - im->cum[RH(g, nt, iy, iz, tx, ty, tz)];
            if (tx == nnt) { src = &c->im[image1d(g, im->ipx, im->icy, im->icz)]; mlast = src->cum[RH(g, ncb, iy, iz, 1, ty, tz)]; }
            else mlast = im->cum[RH(g, ncb, iy, iz, tx + 1, ty, tz)];
          }
          memcpy(ARR(im) + 3 * (nlast - nlen), ARR(src) + 3 * (mlast - nlen), 3 * nlen * sizeof(i16));
          if (which == 1 && c->has_pid) memcpy(im->pid + (nlast - nlen), src->pid + (mlast - nlen), nlen * sizeof(i64)); /* buffer_v.f90:22-23,41-42 */
        }
    }
  /* y- then y+ (buffer_x.f90:82-137) */
  for (int side = 0; side < 2; side++)
    for (i64 m = 0; m < g->nimg; m++) {
      image_t *im = &c->im[m];
      for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++)
        for (i64 iz = 1; iz <= nt; iz++) {
          i64 nlast, nlen, mlast; image_t *src = im;
          if (side == 0) {
            nlast = im->cum[RH(g, hi, 0, iz, tx, ty, tz)];
            nlen = nlast - im->cum[RH(g, lo, lo, iz, tx, ty, tz)] + im->rhoc[RH(g, lo, lo, iz, tx, ty, tz)];
            if (ty == 1) { src = &c->im[image1d(g, im->icx, im->iny, im->icz)]; mlast = src->cum[RH(g, hi, nt, iz, tx, nnt, tz)]; }
            else mlast = im->cum[RH(g, hi, nt, iz, tx, ty - 1, tz)];
          } else {
            nlast = im->cum[RH(g, hi, hi, iz, tx, ty, tz)];
            nlen = nlast - im->cum[RH(g, hi, nt, iz, tx, ty, tz)];
            if (ty == nnt) { src = &c->im[image1d(g, im->icx, im->ipy, im->icz)]; mlast = src->cum[RH(g, hi, ncb, iz, tx, 1, tz)]; }
            else mlast = im->cum[RH(g, hi, ncb, iz, tx, ty + 1, tz)];
          }
          memcpy(ARR(im) + 3 * (nlast - nlen), ARR(src) + 3 * (mlast - nlen), 3 * nlen * sizeof(i16));
          if (which == 1 && c->has_pid) memcpy(im->pid + (nlast - nlen), src->pid + (mlast - nlen), nlen * sizeof(i64)); /* buffer_v.f90:22-23,41-42 */
        }
    }
  /* z- then z+ (buffer_x.f90:144-191) */
  for (int side = 0; side < 2; side++)
    for (i64 m = 0; m < g->nimg; m++) {
      image_t *im = &c->im[m];
      for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++) {
        i64 nlast, nlen, mlast; image_t *src = im;
        if (side == 0) {
          nlast = im->cum[RH(g, hi, hi, 0, tx, ty, tz)];
          nlen = nlast - im->cum[RH(g, lo, lo, lo, tx, ty, tz)] + im->rhoc[RH(g, lo, lo, lo, tx, ty, tz)];
          if (tz == 1) { src = &c->im[image1d(g, im->icx, im->icy, im->inz)]; mlast = src->cum[RH(g, hi, hi, nt, tx, ty, nnt)]; }
          else mlast = im->cum[RH(g, hi, hi, nt, tx, ty, tz - 1)];
        } else {
          nlast = im->cum[RH(g, hi, hi, hi, tx, ty, tz)];
          nlen = nlast - im->cum[RH(g, hi, hi, nt, tx, ty, tz)];
          if (tz == nnt) { src = &c->im[image1d(g, im->icx, im->icy, im->ipz)]; mlast = src->cum[RH(g, hi, hi, ncb, tx, ty, 1)]; }
          else mlast = im->cum[RH(g, hi, hi, ncb, tx, ty, tz + 1)];
        }
        memcpy(ARR(im) + 3 * (nlast - nlen), ARR(src) + 3 * (mlast - nlen), 3 * nlen * sizeof(i16));
          if (which == 1 && c->has_pid) memcpy(im->pid + (nlast - nlen), src->pid + (mlast - nlen), nlen * sizeof(i64)); /* buffer_v.f90:22-23,41-42 */
      }
    }
#undef ARR
}
void oracle_buffer_x(ctx_t *c) { buffer_particles(c, 0); }
void oracle_buffer_v(ctx_t *c) { buffer_particles(c, 1); }

/* ---- update_particle.f90 ----------------------------------------------------------- */
void oracle_update_particle(ctx_t *c, float dt_old, float dt) {
  const geom_t *g = &c->g;
  const i64 nt = g->nt, ncb = g->ncb, nnt = g->nnt, lo = 1 - ncb, hi = nt + ncb;
  const i64 lo2 = 1 - 2 * ncb, hi2 = nt + 2 * ncb, ne2 = nt + 4 * ncb, n2 = ne2 * ne2 * ne2;
  const double weight_v = (double)0.1f; /* real(8),parameter :: weight_v=0.1  (f32 literal) :10 */
  const double x_resolution = g->x_resolution; /* 2^-(8*izipx), parameters.f90:101 */
  const float dt_mid = (dt_old + dt) / 2; /* :16 */
  const double S = vscale(c->sigma_vi);
  const i64 nlayer = c->nlayer > 1 ? c->nlayer : 1;
  float ovh_all = 0;
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    im->overhead_tile = 0;
    i64 iright = 0;
    for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++) {
      memset(c->rhoce, 0, n2 * sizeof(i32));
      memset(c->rholocal, 0, n2 * sizeof(i32));
      memset(c->vfield_new, 0, 3 * n2 * sizeof(float));
      for (i64 k = lo; k <= hi; k++) for (i64 j = lo; j <= hi; j++) for (i64 i = lo; i <= hi; i++) /* :27 */
        for (int d = 0; d < 3; d++)
          c->vfield_new[3 * RE(g, i, j, k) + d] = (float)((double)im->vfield[3 * RH(g, i, j, k, tx, ty, tz) + d] * weight_v);
      /* pass 1 :34-51; CUBEnu visits the k planes in nlayer colour passes (k = 1-ncb+ilayer step nlayer), which fixes the
       * order of the f32 additions into vfield_new and of the particles inside a destination cell; nlayer = 1 is CUBE/main */
      for (i64 ilayer = 0; ilayer < nlayer; ilayer++)
      for (i64 k = lo + ilayer; k <= hi; k += nlayer) for (i64 j = lo; j <= hi; j++) for (i64 i = lo; i <= hi; i++) {
        i64 r = RH(g, i, j, k, tx, ty, tz);
        i64 nlast = im->cum[r], np = im->rhoc[r];
        const i64 cell[3] = {i, j, k};
        for (i64 l = 1; l <= np; l++) {
          i64 ip = nlast - np + l - 1; /* 0-based */
          i64 gg[3]; double vreal[3];
          for (int d = 0; d < 3; d++) {
            double xq = ((double)cell[d] - 1.0) + xp_decode(g, im->xp[3 * ip + d]) * x_resolution;
            vreal[d] = (double)vp_tan(g, im->vp[3 * ip + d]) / S;
            vreal[d] = vreal[d] + (double)im->vfield[3 * r + d];
            double deltax = ((double)dt_mid * vreal[d]) / (double)g->ncell;
            gg[d] = (i64)ceil(xq + deltax);
          }
          if (gg[0] < lo2 || gg[0] > hi2 || gg[1] < lo2 || gg[1] > hi2 || gg[2] < lo2 || gg[2] > hi2) {
            c->error = 3; snprintf(c->errmsg, sizeof c->errmsg, "particle left the double buffer (reference would write out of bounds)"); return;
          }
          i64 e = RE(g, gg[0], gg[1], gg[2]);
          c->rhoce[e] += 1;
          for (int d = 0; d < 3; d++) c->vfield_new[3 * e + d] = (float)((double)c->vfield_new[3 * e + d] + vreal[d]);
        }
      }
      for (i64 e = 0; e < n2; e++) /* :55-57 */
        for (int d = 0; d < 3; d++)
          c->vfield_new[3 * e + d] = (float)((double)c->vfield_new[3 * e + d] / ((double)c->rhoce[e] + weight_v));
      cumsum3(g, c->rhoce, c->cume);
      i64 ntot = c->cume[n2 - 1];
      float ov = (float)ntot / (float)g->np_tile_max; /* integer(8)/real(4) -> real(4) :60 */
      if (ov > im->overhead_tile) im->overhead_tile = ov;
      if (ntot > g->np_tile_max) {
        c->error = 1;
        snprintf(c->errmsg, sizeof c->errmsg, "error: too many particles in this tile+buffer: %lld > %lld on image %lld tile %lld %lld %lld; please set tile_buffer larger",
                 (long long)ntot, (long long)g->np_tile_max, (long long)(m + 1), (long long)tx, (long long)ty, (long long)tz);
        return;
      }
      /* pass 2 :70-93 (CUBEnu update_particle.f90:97-103: the same colour passes) */
      for (i64 ilayer = 0; ilayer < nlayer; ilayer++)
      for (i64 k = lo + ilayer; k <= hi; k += nlayer) for (i64 j = lo; j <= hi; j++) for (i64 i = lo; i <= hi; i++) {
        i64 r = RH(g, i, j, k, tx, ty, tz);
        i64 nlast = im->cum[r], np = im->rhoc[r];
        const i64 cell[3] = {i, j, k};
        for (i64 l = 1; l <= np; l++) {
          i64 ip = nlast - np + l - 1;
          i64 gg[3]; double vreal[3];
          for (int d = 0; d < 3; d++) {
            double xq = ((double)cell[d] - 1.0) + xp_decode(g, im->xp[3 * ip + d]) * x_resolution;
            vreal[d] = (double)vp_tan(g, im->vp[3 * ip + d]) / S;
            vreal[d] = vreal[d] + (double)im->vfield[3 * r + d];
            double deltax = ((double)dt_mid * vreal[d]) / (double)g->ncell;
            gg[d] = (i64)ceil(xq + deltax);
          }
          i64 e = RE(g, gg[0], gg[1], gg[2]);
          c->rholocal[e] += 1;
          i64 idx = c->cume[e] - c->rhoce[e] + c->rholocal[e] - 1; /* 0-based */
          for (int d = 0; d < 3; d++) {
            /* xp+nint(dt_mid*vreal/(x_resolution*ncell)) : int(izipx) + i32 -> i32 -> int(izipx) (wraps) :84 */
            i32 dxi = (i32)lround((double)dt_mid * vreal[d] / (x_resolution * (double)g->ncell));
            c->xp_new[3 * idx + d] = to_kind((i64)((i32)im->xp[3 * ip + d] + dxi), g->izipx);
            double vr = vreal[d] - (double)c->vfield_new[3 * e + d];
            c->vp_new[3 * idx + d] = vp_encode(g, vr, S);
          }
          if (c->has_pid) c->pid_new[idx] = im->pid[ip]; /* :88 */
        }
      }
      /* delete buffer particles :97-109 */
      for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) {
        i64 nlast = c->cume[RE(g, nt, j, k)];
        i64 nlen = nlast - c->cume[RE(g, 0, j, k)];
        memcpy(im->xp + 3 * iright, c->xp_new + 3 * (nlast - nlen), 3 * nlen * sizeof(i16));
        memcpy(im->vp + 3 * iright, c->vp_new + 3 * (nlast - nlen), 3 * nlen * sizeof(i16));
        if (c->has_pid) memcpy(im->pid + iright, c->pid_new + (nlast - nlen), nlen * sizeof(i64)); /* :106 */
        iright += nlen;
      }
      /* :111-112 */
      for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) for (i64 i = 1; i <= nt; i++) {
        i64 r = RH(g, i, j, k, tx, ty, tz), e = RE(g, i, j, k);
        im->rhoc[r] = c->rhoce[e];
        for (int d = 0; d < 3; d++) im->vfield[3 * r + d] = c->vfield_new[3 * e + d];
      }
    }
    im->nplocal = iright; /* :117-119 */
    memset(im->xp + 3 * iright, 0, 3 * (g->np_image_max - iright) * sizeof(i16));
    memset(im->vp + 3 * iright, 0, 3 * (g->np_image_max - iright) * sizeof(i16));
    if (c->has_pid) memset(im->pid + iright, 0, (g->np_image_max - iright) * sizeof(i64)); /* :121 */
    if (im->overhead_tile > ovh_all) ovh_all = im->overhead_tile;
  }
  /* velocity statistics :129-175 */
  double sv = 0, svc = 0, svr = 0;
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    i64 ip = 0; double a = 0, b = 0, r_ = 0;
    for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++)
      for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) for (i64 i = 1; i <= nt; i++) {
        i64 r = RH(g, i, j, k, tx, ty, tz);
        const float *vf = &im->vfield[3 * r];
        /* sum(vfield**2): f32 squares, f32 sum, then promoted */
        b = b + (double)((vf[0] * vf[0] + vf[1] * vf[1]) + vf[2] * vf[2]);
        for (i64 l = 1; l <= im->rhoc[r]; l++, ip++) {
          double v[3];
          for (int d = 0; d < 3; d++) v[d] = (double)vp_tan(g, im->vp[3 * ip + d]) / S;
          r_ = r_ + ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
          for (int d = 0; d < 3; d++) v[d] = v[d] + (double)vf[d];
          a = a + ((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
        }
      }
    im->std_vsim = a; im->std_vsim_c = b; im->std_vsim_res = r_;
  }
  /* co_sum on head :154-166 */
  sv = c->im[0].std_vsim; svc = c->im[0].std_vsim_c; svr = c->im[0].std_vsim_res;
  for (i64 m = 1; m < g->nimg; m++) { svc += c->im[m].std_vsim_c; svr += c->im[m].std_vsim_res; sv += c->im[m].std_vsim; }
  /* :170-175.  std_vsim_c/nc/nc/nc/nn/nn/nn generalised to the per-dim image counts */
  sv = sqrt(sv / (double)c->npglobal);
  svc = sqrt(svc / (double)g->nc / (double)g->nc / (double)g->nc / (double)g->nn[0] / (double)g->nn[1] / (double)g->nn[2]);
  svr = sqrt(svr / (double)c->npglobal);
  for (i64 m = 0; m < g->nimg; m++) { c->im[m].std_vsim = sv; c->im[m].std_vsim_c = svc; c->im[m].std_vsim_res = svr; c->im[m].overhead_tile = ovh_all; }
  c->sigma_vi_new = (float)(svr / (double)sqrtf(3.f));
  /* clean up buffer region of rhoc :197-202 */
  for (i64 m = 0; m < g->nimg; m++) {
    image_t *im = &c->im[m];
    for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++)
      for (i64 k = lo; k <= hi; k++) for (i64 j = lo; j <= hi; j++) for (i64 i = lo; i <= hi; i++)
        if (i < 1 || i > nt || j < 1 || j > nt || k < 1 || k > nt) im->rhoc[RH(g, i, j, k, tx, ty, tz)] = 0;
  }
}

/* ---- particle_mesh (pm.f90) pieces; FFTs happen in Python between them ---------------- */
void oracle_pm_begin(ctx_t *c) { /* pm.f90:27-35 */
  for (i64 m = 0; m < c->g.nimg; m++) {
    image_t *im = &c->im[m];
    cumsum6(&c->g, im->rhoc, im->cum);
    im->vmax = 0; im->f2_max_coarse = 0; im->vmax3[0] = im->vmax3[1] = im->vmax3[2] = 0;
    for (i64 q = 0; q < c->g.nnt * c->g.nnt * c->g.nnt; q++) im->f2_max_fine[q] = 0;
  }
}

/* CIC helper (pm.f90:54-60): tempx (f32) -> idx1, idx2, dx1, dx2 */
static inline void cic(float tempx, i64 *idx1, float *dx1, float *dx2) {
  *idx1 = (i64)floorf(tempx) + 1;
  *dx1 = (float)(*idx1) - tempx;
  *dx2 = 1 - *dx1;
}

/* fine_cic_mass pm.f90:44-72.  rho_f(nfe+2,nfe,nfe) column-major, caller-zeroed NOT required. */
void oracle_fine_deposit(ctx_t *c, i64 m, i64 tx, i64 ty, i64 tz, float *rho_f) {
  const geom_t *g = &c->g; image_t *im = &c->im[m];
  const i64 nt = g->nt, ncb = g->ncb, nfe = g->nfe, nfb = g->nfb, ld = nfe + 2;
  const float mass_p = c->mass_p;
  const double x_resolution = g->x_resolution; /* 2^-(8*izipx), parameters.f90:101 */
  memset(rho_f, 0, ld * nfe * nfe * sizeof(float));
#define RF(a, b, cc) rho_f[((a)-1) + ld * (((b)-1) + nfe * ((cc)-1))]
  for (i64 k = 2 - ncb; k <= nt + ncb - 1; k++) for (i64 j = 2 - ncb; j <= nt + ncb - 1; j++) for (i64 i = 2 - ncb; i <= nt + ncb - 1; i++) {
    i64 nlast = im->cum[RH(g, i - 1, j, k, tx, ty, tz)], np = im->rhoc[RH(g, i, j, k, tx, ty, tz)];
    const i64 cell[3] = {i, j, k};
    for (i64 l = 1; l <= np; l++) {
      i64 ip = nlast + l - 1;
      i64 i1[3], i2[3]; float d1[3], d2[3];
      for (int d = 0; d < 3; d++) {
        /* tempx=4.*((/i,j,k/)-1)+4*(int(xp+ishift,izipx)+rshift)*x_resolution  (f32 + f64 -> f32) */
        float tempx = (float)((double)(4.f * (float)(cell[d] - 1)) + 4 * xp_decode(g, im->xp[3 * ip + d]) * x_resolution);
        cic(tempx, &i1[d], &d1[d], &d2[d]);
        i1[d] += nfb; i2[d] = i1[d] + 1;
      }
      RF(i1[0], i1[1], i1[2]) += d1[0] * d1[1] * d1[2] * mass_p;
      RF(i2[0], i1[1], i1[2]) += d2[0] * d1[1] * d1[2] * mass_p;
      RF(i1[0], i2[1], i1[2]) += d1[0] * d2[1] * d1[2] * mass_p;
      RF(i1[0], i1[1], i2[2]) += d1[0] * d1[1] * d2[2] * mass_p;
      RF(i1[0], i2[1], i2[2]) += d1[0] * d2[1] * d2[2] * mass_p;
      RF(i2[0], i1[1], i2[2]) += d2[0] * d1[1] * d2[2] * mass_p;
      RF(i2[0], i2[1], i1[2]) += d2[0] * d2[1] * d1[2] * mass_p;
      RF(i2[0], i2[1], i2[2]) += d2[0] * d2[1] * d2[2] * mass_p;
    }
  }
#undef RF
}

/* 8-corner gather shared by the two kicks: v += F*a_mid*dt/6/pi*wx*wy*wz in the order of pm.f90:104-111 */
#define KICK8(F3)                                                                          \
  do {                                                                                     \
    static const int cx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, cy[8] = {0, 0, 1, 0, 1, 0, 1, 1},   \
                     cz[8] = {0, 0, 0, 1, 1, 1, 0, 1};                                     \
    for (int q = 0; q < 8; q++) {                                                          \
      const float *F = F3(cx[q] ? i2[0] : i1[0], cy[q] ? i2[1] : i1[1], cz[q] ? i2[2] : i1[2]); \
      float wx = cx[q] ? d2[0] : d1[0], wy = cy[q] ? d2[1] : d1[1], wz = cz[q] ? d2[2] : d1[2]; \
      for (int d = 0; d < 3; d++)                                                          \
        vreal[d] = vreal[d] + (double)(F[d] * a_mid * dt / 6 / PI_F * wx * wy * wz);       \
    }                                                                                      \
  } while (0)

/* fine velocity pm.f90:85-118.  force_f(3, nfb:nfe-nfb+1, same, same) column-major. */
void oracle_fine_kick(ctx_t *c, i64 m, i64 tx, i64 ty, i64 tz, const float *force_f, float a_mid, float dt) {
  const geom_t *g = &c->g; image_t *im = &c->im[m];
  const i64 nt = g->nt, nfb = g->nfb, nff = g->nft + 2;
  const double x_resolution = g->x_resolution; /* 2^-(8*izipx), parameters.f90:101 */
  const double S = vscale(c->sigma_vi), Snew = vscale(c->sigma_vi_new);
#define FF(a, b, cc) (&force_f[3 * (((a)-nfb) + nff * (((b)-nfb) + nff * ((cc)-nfb)))])
  /* f2_max_fine(itx,ity,itz)=maxval(sum(force_f**2,1)) :85 */
  float f2 = 0; /* maxval of non-negative numbers; array is non-empty */
  f2 = -INFINITY;
  for (i64 q = 0; q < nff * nff * nff; q++) {
    const float *F = &force_f[3 * q];
    float s = (F[0] * F[0] + F[1] * F[1]) + F[2] * F[2];
    if (s > f2) f2 = s;
  }
  im->f2_max_fine[(tx - 1) + g->nnt * ((ty - 1) + g->nnt * (tz - 1))] = f2;
  for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) for (i64 i = 1; i <= nt; i++) {
    i64 nlast = im->cum[RH(g, i - 1, j, k, tx, ty, tz)], np = im->rhoc[RH(g, i, j, k, tx, ty, tz)];
    const i64 cell[3] = {i, j, k};
    for (i64 l = 1; l <= np; l++) {
      i64 ip = nlast + l - 1;
      i64 i1[3], i2[3]; float d1[3], d2[3]; double vreal[3];
      for (int d = 0; d < 3; d++) {
        float tempx = (float)((double)(4.f * (float)(cell[d] - 1)) + 4 * xp_decode(g, im->xp[3 * ip + d]) * x_resolution);
        cic(tempx, &i1[d], &d1[d], &d2[d]);
        i1[d] += nfb; i2[d] = i1[d] + 1;
        vreal[d] = (double)vp_tan(g, im->vp[3 * ip + d]) / S;
      }
      KICK8(FF);
      for (int d = 0; d < 3; d++) im->vp[3 * ip + d] = vp_encode(g, vreal[d], Snew);
    }
  }
#undef FF
}
void oracle_pm_fine_end(ctx_t *c) { c->sigma_vi = c->sigma_vi_new; } /* pm.f90:122 */

/* coarse_cic_mass pm.f90:130-163 -> r3(nc,nc,nc) of image m */
void oracle_coarse_deposit(ctx_t *c, i64 m, float *r3) {
  const geom_t *g = &c->g; image_t *im = &c->im[m];
  const i64 nt = g->nt, nnt = g->nnt, nc = g->nc, e = nt + 4;
  const float mass_p = c->mass_p;
  const double x_resolution = g->x_resolution; /* 2^-(8*izipx), parameters.f90:101 */
  float *r3t = (float *)malloc(e * e * e * sizeof(float));
#define RT(a, b, cc) r3t[((a) + 1) + e * (((b) + 1) + e * ((cc) + 1))]
  for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++) {
    memset(r3t, 0, e * e * e * sizeof(float));
    for (i64 k = 0; k <= nt + 1; k++) for (i64 j = 0; j <= nt + 1; j++) for (i64 i = 0; i <= nt + 1; i++) {
      i64 nlast = im->cum[RH(g, i - 1, j, k, tx, ty, tz)], np = im->rhoc[RH(g, i, j, k, tx, ty, tz)];
      const i64 cell[3] = {i, j, k};
      for (i64 l = 1; l <= np; l++) {
        i64 ip = nlast + l - 1;
        i64 i1[3], i2[3]; float d1[3], d2[3];
        for (int d = 0; d < 3; d++) {
          /* tempx=((/i,j,k/)-1)+(int(xp+ishift,izipx)+rshift)*x_resolution-0.5  (i64 + f64 - f32 -> f64 -> f32) */
          float tempx = (float)((double)(cell[d] - 1) + xp_decode(g, im->xp[3 * ip + d]) * x_resolution - (double)0.5f);
          cic(tempx, &i1[d], &d1[d], &d2[d]);
          i2[d] = i1[d] + 1;
        }
        RT(i1[0], i1[1], i1[2]) += d1[0] * d1[1] * d1[2] * mass_p;
        RT(i2[0], i1[1], i1[2]) += d2[0] * d1[1] * d1[2] * mass_p;
        RT(i1[0], i2[1], i1[2]) += d1[0] * d2[1] * d1[2] * mass_p;
        RT(i1[0], i1[1], i2[2]) += d1[0] * d1[1] * d2[2] * mass_p;
        RT(i1[0], i2[1], i2[2]) += d1[0] * d2[1] * d2[2] * mass_p;
        RT(i2[0], i1[1], i2[2]) += d2[0] * d1[1] * d2[2] * mass_p;
        RT(i2[0], i2[1], i1[2]) += d2[0] * d2[1] * d1[2] * mass_p;
        RT(i2[0], i2[1], i2[2]) += d2[0] * d2[1] * d2[2] * mass_p;
      }
    }
    for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) for (i64 i = 1; i <= nt; i++)
      r3[((tx - 1) * nt + i - 1) + nc * (((ty - 1) * nt + j - 1) + nc * ((tz - 1) * nt + k - 1))] = RT(i, j, k);
  }
#undef RT
  free(r3t);
}

/* coarse velocity pm.f90:192-228.  force_c(3,0:nc+1,0:nc+1,0:nc+1) of image m, halo already filled. */
void oracle_coarse_kick(ctx_t *c, i64 m, const float *force_c, float a_mid, float dt) {
  const geom_t *g = &c->g; image_t *im = &c->im[m];
  const i64 nt = g->nt, nnt = g->nnt, nc = g->nc, e = nc + 2;
  const double x_resolution = g->x_resolution; /* 2^-(8*izipx), parameters.f90:101 */
  const double S = vscale(c->sigma_vi);
#define FC(a, b, cc) (&force_c[3 * ((a) + e * ((b) + e * (cc)))])
  float f2 = -INFINITY; /* f2_max_coarse=maxval(sum(force_c**2,1)) :192 */
  for (i64 q = 0; q < e * e * e; q++) {
    const float *F = &force_c[3 * q];
    float s = (F[0] * F[0] + F[1] * F[1]) + F[2] * F[2];
    if (s > f2) f2 = s;
  }
  im->f2_max_coarse = f2;
  float vmax = im->vmax;
  for (i64 tz = 1; tz <= nnt; tz++) for (i64 ty = 1; ty <= nnt; ty++) for (i64 tx = 1; tx <= nnt; tx++) {
    const i64 tile[3] = {tx, ty, tz};
    for (i64 k = 1; k <= nt; k++) for (i64 j = 1; j <= nt; j++) for (i64 i = 1; i <= nt; i++) {
      i64 r = RH(g, i, j, k, tx, ty, tz);
      i64 nlast = im->cum[RH(g, i - 1, j, k, tx, ty, tz)], np = im->rhoc[r];
      const i64 cell[3] = {i, j, k};
      for (i64 l = 1; l <= np; l++) {
        i64 ip = nlast + l - 1;
        i64 i1[3], i2[3]; float d1[3], d2[3]; double vreal[3];
        for (int d = 0; d < 3; d++) {
          /* tempx=((/itx,ity,itz/)-1)*nt+((/i,j,k/)-1)+(...)*x_resolution-0.5 */
          float tempx = (float)((double)((tile[d] - 1) * nt + (cell[d] - 1)) + xp_decode(g, im->xp[3 * ip + d]) * x_resolution - (double)0.5f);
          cic(tempx, &i1[d], &d1[d], &d2[d]);
          i2[d] = i1[d] + 1;
          vreal[d] = (double)vp_tan(g, im->vp[3 * ip + d]) / S;
        }
        KICK8(FC);
        /* vmax=max(vmax,maxval(vreal+vfield(:,i,j,k,...)))  f64 max assigned to f32 :220 */
        double mx = vreal[0] + (double)im->vfield[3 * r + 0];
        for (int d = 1; d < 3; d++) { double t = vreal[d] + (double)im->vfield[3 * r + d]; if (t > mx) mx = t; }
        if (mx > (double)vmax) vmax = (float)mx;
        for (int d = 0; d < 3; d++) { /* CUBEnu pm.f90:349: vmax=max(vmax,abs(vreal+vfield)) per component, f64 max -> f32 */
          double t = fabs(vreal[d] + (double)im->vfield[3 * r + d]);
          if (t > (double)im->vmax3[d]) im->vmax3[d] = (float)t;
        }
        for (int d = 0; d < 3; d++) im->vp[3 * ip + d] = vp_encode(g, vreal[d], S);
      }
    }
  }
  im->vmax = vmax;
#undef FC
}
float oracle_f2_max_coarse(ctx_t *c, i64 m) { return c->im[m].f2_max_coarse; }
