/*
 * cube_gpu.h -- C ABI of the B200-native CUBE particle-mesh step (libcubegpu.so).
 *
 * The reference (yuhaoran/cafproject, Coarray Fortran) has no plugin/FFI interface: its step loop
 * calls argument-less subroutines that work on module globals (CUBE/main/cafcube.f90:25-46,
 * CUBE/main/variables.f90:17-68).  This header is the boundary *cut* at those calls; each entry
 * point names the reference routine it replaces.  The Fortran side binds them with ISO_C_BINDING
 * (fortran/cube_gpu.f90) and passes its globals explicitly.
 *
 * Conventions
 *  - plain pointers and sizes only; all arrays are host memory owned by the caller and touched
 *    only inside the call; the authoritative state lives in HBM between upload and download.
 *  - arrays use the reference's own memory layout (Fortran column-major, first index fastest):
 *      xp                integer(izipx) (3, nplocal)   izipx, izipv = 1 or 2 bytes per code, the run's zip format
 *      vp                integer(izipv) (3, nplocal)   (CUBE/main/universe*.fh:2-3, variables.f90:41-42, parameters.f90:13-15)
 *      rhoc_phys         integer(4)  (nt,nt,nt,nnt,nnt,nnt)   = rhoc(1:nt,1:nt,1:nt,:,:,:)   checkpoint.f90:35
 *      vfield_phys       real(4)     (3,nt,nt,nt,nnt,nnt,nnt) = vfield(:,1:nt,1:nt,1:nt,:,:,:) checkpoint.f90:40
 *    i.e. exactly what the reference writes to zip0/zip1/zip2/vfield ("disjoint state").
 *  - every function returns 0 on success; non-zero = error, message in cube_gpu_last_error().
 *    Capacity overflows are reported with the reference's own wording (update_particle.f90:61-67,
 *    buffer_density.f90:87-93) instead of `stop`.
 *  - single caller per handle; calls are not re-entrant (the reference's step calls never overlap).
 *  - there is no CPU fallback: every call fails if no CUDA device is usable.
 */
#ifndef CUBE_GPU_H
#define CUBE_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Compile-time constants of CUBE/main/parameters.f90:13-60 passed at run time. */
typedef struct cube_params {
  int32_t nn[3];        /* images per dimension (reference: nn,nn,nn; parameters.f90:20)            */
  int32_t rank;         /* this image, 0-based = this_image()-1 (parameters.f90:180)               */
  int32_t nnt;          /* tiles / image / dim                         parameters.f90:22           */
  int32_t nc;           /* coarse cells / image / dim                  parameters.f90:23           */
  int32_t ncell;        /* fine cells per coarse cell / dim (must be 4) parameters.f90:21           */
  int32_t ncb;          /* buffer depth in coarse cells (must be 6)    parameters.f90:46           */
  int32_t izipx, izipv; /* bytes per position / velocity code, each 1 or 2   universe*.fh:2-3      */
  int32_t np_nc;        /* particles / coarse cell / dim (capacity sizing) parameters.f90:55       */
  float image_buffer;   /* parameters.f90:59 */
  float tile_buffer;    /* parameters.f90:60 */
  int32_t device;       /* CUDA device ordinal */
  int32_t fine_batch;   /* tiles per batched fine-mesh FFT, 0 = choose automatically */
  int32_t local_group;  /* 0: one process per image (NCCL when nn>1); k>0: all images of the run are host threads of THIS
                           process and share in-process group k (several images per GPU, `-fcoarray=single`-style runs) */
  int32_t reserved[3];  /* reserved[0] bit 0: this handle is a further species of a two-species run (cube_gpu_particle_mesh_species):
                           its own fine-mesh pipeline is allocated for one tile only */
} cube_params;

typedef struct cube_handle cube_handle;

/* initialize.f90:1-56: geometry, FFT plans, kernel_f (kernel_f.f90), kernel_c (kernel_c.f90).
 *   fk_table  real(4) (16,16,16,3)  = fk_table(i,j,k,dim) as read from ../kernels/wfxyzf.3.ascii
 *   ck_table  real(4) (3,4,4,4)     = ck_table(dim,i,j,k) as read from ../kernels/wfxyzc.2.ascii
 *   tanf_lut  real(4) (0:nvbin-1)   tan((pi*real(v))/real(nvbin-1)) evaluated BY THE HOST's libm for
 *                                   v = int(u,izipv) (u = raw pattern of the code, nvbin = 2**(8*izipv):
 *                                   65536 or 256 entries) -- makes velocity decoding bit-identical to
 *                                   the host build (pm.f90:102, update_particle.f90:42)
 *   nccl_unique_id  NULL for a single image; otherwise the 128-byte ncclUniqueId shared by all images. */
int cube_gpu_init(const cube_params *p, const float *fk_table, const float *ck_table, const float *tanf_lut,
                  const void *nccl_unique_id, cube_handle **h);

/* Replaces the coarray runtime's bootstrap (cafcube.f90:6-14, `sync all`): image 1 calls this and broadcasts the 128
 * bytes to every image (coarray assignment, MPI_Bcast, torch.distributed...), which pass them to cube_gpu_init.
 * The image grid is nn[0] x nn[1] x nn[2] (the reference hard-wires nn^3, parameters.f90:181-183); image
 * rank = icx-1 + nn[0]*((icy-1) + nn[1]*(icz-1)) as in parameters.f90:200-203. */
int cube_gpu_nccl_unique_id(void *id128);

/* particle_initialization.f90:11-72: take the disjoint state (file order). mass_p = nf_global^3/npglobal. */
int cube_gpu_upload(cube_handle *h, const void *xp, const void *vp, const int32_t *rhoc_phys,
                    const float *vfield_phys, int64_t nplocal, int64_t npglobal, float sigma_vi);

/* The same, returning while xp and vp are still on the bus (the cell arrays and their scan are done): xp and vp must be page-locked
 * and stay unchanged until the next call that uses the particles has returned.  cube_gpu_update_x keys the particles chunk by
 * chunk as they land (its first pass then runs under the transfer); every other entry point first waits for the whole upload. */
int cube_gpu_upload_begin(cube_handle *h, const void *xp, const void *vp, const int32_t *rhoc_phys,
                          const float *vfield_phys, int64_t nplocal, int64_t npglobal, float sigma_vi);

/* update_particle (update_particle.f90:1-213): drift + cell re-sort + vfield rebuild + sigma statistics.
 * in: buffered state; out: disjoint state.  std_vsim = {std_vsim, std_vsim_c, std_vsim_res}. */
int cube_gpu_update_x(cube_handle *h, float dt_old, float dt, int64_t *nplocal, float *sigma_vi_new,
                      double std_vsim[3], float *overhead_tile);

/* buffer_density / buffer_x / buffer_v (buffer_density.f90, buffer_x.f90, buffer_v.f90):
 * disjoint -> buffered state.  Flags select which of the three reference calls this stands for. */
int cube_gpu_buffer(cube_handle *h, int do_density, int do_x, int do_v, float *overhead_image);

/* particle_mesh (pm.f90:1-247): fine + coarse PM force, both kicks, time-step limits. */
int cube_gpu_particle_mesh(cube_handle *h, float a_mid, float dt, float *dt_fine, float *dt_coarse,
                           float *dt_vmax, float *vmax);

/* checkpoint.f90:33-70: bring the disjoint state back.  Pass NULL for anything not wanted.
 * Capacity of xp/vp must be >= nplocal as returned by the last update_x/upload. */
int cube_gpu_download(cube_handle *h, void *xp, void *vp, int32_t *rhoc_phys, float *vfield_phys,
                      int64_t *nplocal, float *sigma_vi);

/* -DPID (CUBE/main variables.f90:44, particle_initialization.f90:56-59; on by default in CUBEnu's Makefile): optional particle
 * IDs.  upload_pid gives the IDs of the nplocal particles of the LAST cube_gpu_upload, in the same file order; they then take the
 * permutation of every cube_gpu_update_x (update_particle.f90:88,106) and download_pid returns them in the order of the current
 * disjoint state (what checkpoint.f90 writes to `zipid`).  With several images the IDs of the ghost particles travel with vp in
 * cube_gpu_buffer(do_v) (buffer_v.f90:23,42,62,81,104), so a particle that crosses an image boundary keeps its ID.  A new
 * cube_gpu_upload drops the IDs. */
int cube_gpu_upload_pid(cube_handle *h, const int64_t *pid);
int cube_gpu_download_pid(cube_handle *h, int64_t *pid);

/* Streamed checkpoint: start the device->host copy of xp and/or vp (NULL = skip) of the current disjoint state behind the
 * work already queued and return at once; cube_gpu_download (with NULL for what was streamed) waits for it.  Host buffers
 * should be page-locked.  Typical use: xp right after cube_gpu_update_x -- particle_mesh does not move particles, so the
 * position traffic of checkpoint.f90:43-50 overlaps the force computation. */
int cube_gpu_download_async(cube_handle *h, void *xp, void *vp);
/* The same for the per-cell arrays rhoc(nt,nt,nt,nnt,nnt,nnt) and vfield(3,...) of checkpoint.f90:35,40: they are final once
 * cube_gpu_update_x has returned (particle_mesh changes neither), so they too can leave under the force computation.
 * cube_gpu_download with NULL for them waits for the stream. */
int cube_gpu_download_cells_async(cube_handle *h, int32_t *rhoc_phys, float *vfield_phys);
/* Velocities: registers a page-locked buffer for the NEXT cube_gpu_particle_mesh call only.  That call then works through the
 * tiles in (at least four) batches, gives every batch its coarse kick right after its fine kick (the coarse force only needs
 * positions and is ready by then; per particle the two kicks are the same operations in the same order as pm.f90:88-228) and
 * streams the batch's final velocities into vp while the next batch's force is computed -- the velocity half of
 * checkpoint.f90:51-58 leaves under the computation.  cube_gpu_download with NULL for vp waits for the stream. */
int cube_gpu_stream_vp(cube_handle *h, void *vp);

/* CUBEnu keeps the same arithmetic with other bookkeeping (CUBEnu/work/main/update_particle.f90:37,55-58, pm.f90:349,398):
 * update_xp visits the source planes in `nlayer = 2*ceiling(dt_mid*sim%vz_max/ncell)+1` colour passes, which fixes the order of the
 * particles inside a destination cell (and of the f32 additions into vfield_new), and particle_mesh keeps vmax(3) = max|v| per
 * component.  set_drift_layers(nlayer) selects that order for the following cube_gpu_update_x calls (0 or 1 = CUBE/main's, the default);
 * get_vmax3 returns vmax(3) of the last cube_gpu_particle_mesh (whose `vmax` argument stays CUBE/main's scalar, pm.f90:220). */
int cube_gpu_set_drift_layers(cube_handle *h, int nlayer);
int cube_gpu_get_vmax3(cube_handle *h, float vmax3[3]);

/* Two particle species sharing the meshes (CUBEnu -DNEUTRINOS: pm.f90:79-99,160,235,356; update_particle.f90:326-353).  Every species is
 * a handle of its own: its zip format, codes, rhoc/vfield, sigma_vi, capacities and ghost exchange -- cube_gpu_init / upload / update_x /
 * buffer / download are called per species exactly as for one.  set_mass_p gives a species its particle mass (sim%mass_p_cdm,
 * sim%mass_p_nu; the single-species default is nf_global^3/npglobal).  particle_mesh_species is particle_mesh for both: the particles
 * of `h2` are deposited into `h`'s fine and coarse meshes after h's own, one convolution, and both species are kicked by its forces,
 * each with its own sigma_vi / sigma_vi_new; dt_vmax2, vmax2 are the second species' (dt_vmax_nu).  Same geometry and device required. */
int cube_gpu_set_mass_p(cube_handle *h, float mass_p);
int cube_gpu_particle_mesh_species(cube_handle *h, cube_handle *h2, float a_mid, float dt, float *dt_fine, float *dt_coarse,
                                   float *dt_vmax, float *vmax, float *dt_vmax2, float *vmax2);

int cube_gpu_finalize(cube_handle *h);
const char *cube_gpu_last_error(void);

/* cicpower + powerspectrum (CUBE/utilities/cicpower.f90:70-167, powerspectrum.f90:21-108 with linear_kbin) of the state resident
 * on the device -- the z = 0 P(k) gate without bringing the particles back: cell-centred CIC on ng = nf = 4*nc nodes with the
 * library's own deposit kernel, density contrast, r2c (cuFFT), shell sums.  xi = the reference's xi(10,nbin), stored row by row
 * (xi[r*nbin + i]): r = 0 mode count, 1 k [h/Mpc], 2 (= 3 = 4) Delta^2(k), 5, 6 the sinc kernels, 7 r, 8 b, 9 reconstructed power;
 * nbin = nint(nyquist*sqrt(3)) is returned in *nbin (call with nbin_cap = 0 to ask).  Needs the buffered state; single image. */
int cube_gpu_power_spectrum(cube_handle *h, float box, double *xi, int nbin_cap, int *nbin);

/* ---- diagnostics used by the parity tests and the benchmark (not part of the step loop) -------- */

/* derived sizes: what[] = "np_image_max","np_tile_max","nfe","nft","nt","fine_batch","kernel_launches" */
int64_t cube_gpu_query(cube_handle *h, const char *what);
/* kern_f(nfe/2+1,nfe,nfe,3) and kern_c(nc*nn/2+1,nc,nc,3) (single image) as built at init, unscaled */
int cube_gpu_get_kern_f(cube_handle *h, float *out);
int cube_gpu_get_kern_c(cube_handle *h, float *out);
/* rho_f(nfe+2,nfe,nfe) of tile (itx,ity,itz) (1-based) from the current buffered state  pm.f90:44-72 */
int cube_gpu_fine_density(cube_handle *h, int itx, int ity, int itz, float *rho_f);
/* force_f(3,nft+2,nft+2,nft+2) of that tile  pm.f90:75-84 */
int cube_gpu_fine_force(cube_handle *h, int itx, int ity, int itz, float *force_f);
/* fine kick of one tile with a caller-supplied force_f (parity: identical F => identical vp) pm.f90:88-118 */
int cube_gpu_fine_kick_with(cube_handle *h, int itx, int ity, int itz, const float *force_f, float a_mid,
                            float dt, float sigma_vi, float sigma_vi_new, float *f2_max);
/* r3(nc,nc,nc)  pm.f90:130-163 */
int cube_gpu_coarse_density(cube_handle *h, float *r3);
/* force_c(3,0:nc+1,0:nc+1,0:nc+1) incl. halo  pm.f90:168-189 */
int cube_gpu_coarse_force(cube_handle *h, float *force_c);
/* coarse kick with a caller-supplied force_c  pm.f90:192-228 */
int cube_gpu_coarse_kick_with(cube_handle *h, const float *force_c, float a_mid, float dt, float sigma_vi,
                              float *vmax, float *f2_max);
/* self-test of the table-driven velocity-code conversions against their defining formulas (pm.f90:102,113): the encoder
 * on both sides of every code threshold plus `nsweep` values over [1e-12,1e8], the shared-memory decoder on all 65536
 * codes for this sigma_vi.  Returns the mismatch counts (0, 0 expected) and whether the FMA form of the division by S
 * reproduced t/S for every table entry (else the kernels use a true division). */
int cube_gpu_selftest_codes(cube_handle *h, float sigma_vi, int64_t nsweep, int64_t *bad_encode, int64_t *bad_decode,
                            int *fma_division);
/* message plan of one image, host only (no device needed): rows of 8 int64, returns the row count (-1 = bad params)
 *   {0, rx, ry, rz, src_rank, dst_rank, ncell, cell0}  ghost direction: I receive my ghost box from src_rank and send
 *                                                      the opposite physical box to dst_rank (buffer_density/x/v)
 *   {1, rank, nplanes, ...} / {2, rank, nplanes, ...}  force_c planes sent to / received from `rank` (pm.f90:176-189)
 *   {3, R, Gx, Gy, Gz, sz, nyl, grp0}                  distributed coarse FFT geometry (replaces pencil_fft.f90) */
int cube_gpu_exchange_plan(const cube_params *p, int64_t *out, int cap_rows);
/* per-phase device times (ms) of the last update_x / particle_mesh, CUBEnu-style phase brackets
 * (CUBEnu/work/main/pm.f90:35,195,268,299,388): names returned via cube_gpu_phase_name(i) */
int cube_gpu_phase_count(void);
const char *cube_gpu_phase_name(int i);
int cube_gpu_phase_times(cube_handle *h, float *ms);
/* device-side stopwatch on the library's own launch stream (CUDA events): start=1 records the start event,
 * start=0 records the stop event, waits for it and returns the elapsed milliseconds in *ms. */
int cube_gpu_timer(cube_handle *h, int start, float *ms);
/* enable (1) / disable (0) per-phase event timing (off by default; costs a few event records) */
int cube_gpu_set_profiling(cube_handle *h, int on);

#ifdef __cplusplus
}
#endif
#endif /* CUBE_GPU_H */
