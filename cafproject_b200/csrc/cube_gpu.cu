// cube_gpu.cu -- host side of libcubegpu.so: the C ABI of include/cube_gpu.h.
//
// State lives in HBM between calls:
//   disjoint state  : xp, vp (file order), rhoc_p, vfield_p, cstart_p           (what update_x leaves)
//   buffered state  : + rhoc_e, cstart_e, vfield_e on the extended image grid   (what buffer builds)
// See DESIGN.md for the layout and the per-kernel roofline.
#include <cuda_runtime.h>
#include <cufft.h>

#include <cmath>
#include <cstdarg>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cube_gpu.h"
#include "cube_kernels.cuh"
#include "cube_fft.cuh"
#include "cube_particles.cuh"
#include "cube_kick.cuh"
#include "cube_comm.cuh"
#include "cube_exchange.cuh"
#include "cube_coarse.cuh"

#include <memory>

using namespace cube;

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return 1;
}
// a CUDA failure of an image that shares a communicator: abort it (cube_comm.cuh) -- an image must not return between two collectives
// and leave the others waiting in the next one
static void abort_current_comm();
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) { abort_current_comm(); return fail("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } \
  } while (0)
#define CF(call)                                                                                     \
  do {                                                                                               \
    cufftResult r_ = (call);                                                                         \
    if (r_ != CUFFT_SUCCESS) return fail("%s:%d cuFFT error %d in %s", __FILE__, __LINE__, (int)r_, #call); \
  } while (0)
#define CKL() CK(cudaGetLastError())
// one instantiation of the particle kernels per zip format (izipx, izipv in {1,2}; CUBE/main/universe*.fh)
#define FMT_SWITCH(h, ...)                                                       \
  do {                                                                           \
    if ((h)->zx == 2 && (h)->zv == 2) { using F = Fmt<2, 2>; __VA_ARGS__; }      \
    else if ((h)->zx == 1 && (h)->zv == 2) { using F = Fmt<1, 2>; __VA_ARGS__; } \
    else if ((h)->zx == 2 && (h)->zv == 1) { using F = Fmt<2, 1>; __VA_ARGS__; } \
    else { using F = Fmt<1, 1>; __VA_ARGS__; }                                   \
  } while (0)
#define XPC(p) ((const typename F::XT*)(p))
#define VPC(p) ((const typename F::VT*)(p))
#define XPM(p) ((typename F::XT*)(p))
#define VPM(p) ((typename F::VT*)(p))

enum Phase { PH_KEY, PH_COUNT, PH_SCAN, PH_PLACE, PH_BUFFER, PH_FDEP, PH_FFTX, PH_FFTY, PH_FFTZ, PH_IFFTY, PH_IFFTX, PH_FMAX, PH_FKICK,
             PH_CDEP, PH_CFFT, PH_CKICK, PH_N };
static const char* kPhaseNames[PH_N] = {"drift_key", "drift_count", "drift_scan", "drift_place", "buffer", "fine_deposit",
                                        "fine_fft_x", "fine_fft_y", "fine_fft_z_green", "fine_ifft_y", "fine_ifft_x", "fine_f2max",
                                        "fine_kick", "coarse_deposit", "coarse_fft_green", "coarse_kick"};

// one instantiation of the hand-written fine-mesh FFT kernels per supported transform length N = R1*R2
struct FftPlan {
  int R1, R2;
  void (*x_fwd)(FftGeom, RhoView, float2*, const float2*);
  void (*y_fwd)(FftGeom, float2*, const float2*);
  void (*y_inv)(FftGeom, float2*, const float2*);
  void (*z_green)(FftGeom, const float2*, float2*, const float*, float, const float2*);
  void (*z_green_pk)(FftGeom, const float2*, float2*, const float*, float, const float2*);  // A/B variants of the z pass's arithmetic
  void (*z_green_mx)(FftGeom, const float2*, float2*, const float*, float, const float2*);  // single-buffered, three CTAs per SM ("sb")
  void (*x_inv)(FftGeom, const float2*, float*, const float2*, unsigned*);
  int x_inv_threads; size_t x_inv_smem;
  int N() const { return R1 * R2; }
  int threads() const { return FL * (R1 > R2 ? R1 : R2); }
};
template <int R1, int R2> static FftPlan make_plan() {
  return {R1, R2, k_fft_x_fwd<R1, R2>, k_fft_y<R1, R2, -1>, k_fft_y<R1, R2, +1, (R1 * R2 <= 320 ? 4 : 0)>,
          k_fft_z_green<R1, R2>, k_fft_z_green<R1, R2, Pk>, k_fft_z_green<R1, R2, Sc, 1>, k_fft_x_inv3<R1, R2>,
          X3Cfg<R1, R2>::NT, X3Cfg<R1, R2>::SMEM};
}
// N must be >= nft + 32 (see cube_fft.cuh); nt = 12,16,24,32,48,64,128 map to 80,96,128,160,256,288,576
static const FftPlan kPlans[] = {make_plan<8, 10>(), make_plan<8, 12>(), make_plan<8, 16>(), make_plan<10, 16>(), make_plan<12, 16>(),
                                 make_plan<16, 16>(), make_plan<16, 18>(), make_plan<18, 20>(), make_plan<20, 24>(), make_plan<24, 24>()};

struct cube_handle {
  cube_params p;
  Geom g;
  cudaStream_t st = nullptr;
  cudaStream_t st_copy = nullptr; cudaEvent_t ev_copy[2] = {}; bool copy_pending = false, copy_reads_vp = false;  // cube_gpu_download_async
  HostStage rb;  // small read-backs (counts, maxima) through mapped memory: never queued behind a checkpoint on the copy engines
  int zx = 2, zv = 2;                 // bytes per position / velocity code (izipx, izipv)
  int nvbin = 65536;                  // 2^(8 izipv): size of the velocity tables
  void* vp_stream_host = nullptr;     // cube_gpu_stream_vp: where the next particle_mesh streams the final velocities
  double* dvlut2 = nullptr; int* divok2 = nullptr;  // decode table of sigma_vi_new while the main one still serves the fine kick
  unsigned long long* kick_next = nullptr; int kb_hot = 0; bool old_kick = false;  // merged brick kick (cube_kick.cuh)
  CUtensorMap fmap = {}; int kick_stage = 2;  // TMA view of F[batch][M][M][3][FP]
  cudaStream_t st_coarse = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool overlap_coarse = true;  // coarse mesh under the fine mesh
  cudaEvent_t ev_vpack = nullptr, ev_vghost = nullptr; bool vghost_pending = false, async_vghost = true;  // buffer_v's exchange under the next drift's key pass
  cudaEvent_t ev_xghost = nullptr; bool xghost_pending = false, async_xghost = true;  // buffer_x's exchange under the fine deposit of the bricks that read no ghost cell
  cudaStream_t st_main = nullptr;  // h->st, also while particle_mesh points h->st at the side stream
  // cube_gpu_upload_begin: the particles arrive in UP_N chunks of up_stride cells on the copy stream; update_x keys each chunk as it lands
  static constexpr int UP_N = 8;
  cudaEvent_t ev_up[UP_N] = {}; int up_n = 0; long long up_stride = 0;
  long long np_image_max = 0, np_tile_max = 0;
  long long nplocal = 0, npglobal = 0;
  float sigma_vi = 0, sigma_vi_new = 0, mass_p = 0;
  bool buffered = false;
  // particles (double buffered)
  void *xp = nullptr, *vp = nullptr, *xp2 = nullptr, *vp2 = nullptr;  // integer(izipx) xp(3,np), integer(izipv) vp(3,np)
  long long *pid = nullptr, *pid2 = nullptr; bool pid_valid = false;  // -DPID: optional particle IDs (cube_gpu_upload_pid)
  long long* pid_send = nullptr; long long pid_sendcap = 0;           // message buffer of the IDs that travel with vp (buffer_v.f90)
  unsigned short* key = nullptr;
  // coarse-cell arrays, file order
  int *rhoc_p = nullptr, *rhoc_p2 = nullptr;
  float *vfield_p = nullptr, *vfield_p2 = nullptr;
  long long *cstart_p = nullptr, *cstart_p2 = nullptr;
  // extended image grid
  int* rhoc_e = nullptr; long long* cstart_e = nullptr; float* vfield_e = nullptr;
  int* sid_e = nullptr; unsigned *mask_s = nullptr, *mask_e = nullptr; int* farblk = nullptr; float* csum = nullptr;  // source-cell mover summaries of the drift (cube_particles.cuh)
  unsigned char* inflag = nullptr; int *flist = nullptr, *nflag = nullptr; bool count_all = false;  // destination cells with arrivals (pass B's work list)
  int nlayer = 1;              // 1: CUBE/main's in-cell order; > 1: CUBEnu's colour passes (cube_gpu_set_drift_layers)
  float vmax3[3] = {0, 0, 0};  // CUBEnu's vmax(3) of the last particle_mesh
  // scan scratch, reductions
  long long* bsum = nullptr; int nscan_blocks = 0;
  double* stat_partial = nullptr; double* stat3 = nullptr;
  unsigned* rank = nullptr;
  long long* tile_count = nullptr;
  int* maxoff = nullptr; unsigned* f2max = nullptr; unsigned long long* vmax_bits = nullptr;
  // LUTs
  float* tanlut = nullptr; double* dvlut = nullptr; float lut_sigma = -1.f; double* enc = nullptr;
  // cells above these particle counts are processed by a whole warp instead of one thread (cube_kernels.cuh, cube_particles.cuh)
  int heavy_deposit = 32, heavy_count = 64, count_minb = 8;
  int fd_brick = 844;       // brick of the fine deposit: 888 | 884 | 844 coarse cells (cube_kernels.cuh)
  bool shared_region = true;  // one fine-density grid per batch of tiles (nt <= 123); else one window per tile in the tile's own frame
  float* tanh = nullptr; int* divok = nullptr; int vt_hot = 0;  // shared-memory copies of the tables (cube_particles.cuh)
  int nsm = 1;
  // fine mesh (cube_fft.cuh)
  int batch = 1;
  const FftPlan* plan = nullptr; FftGeom fg = {};
  size_t rho_n = 0;                    // elements of the fine-density buffer (the largest batch's region)
  size_t A_n = 0, B_n = 0, F_n = 0;  // elements per tile
  float* rho = nullptr;      // [batch][N][N][N]            (aliases the head of Bk: dead before Bk is written)
  float* rho_own = nullptr;  // its own allocation when the alias does not fit (CUBE_GPU_NFFT on small tiles)
  float2* Ak = nullptr;      // [batch][N][N][P]
  float2* Bk = nullptr;      // [3][batch][M][N][P]
  float* F = nullptr;        // [batch][M][M][3][FP]
  float* kern_f = nullptr;   // [3][N][N][P]  unscaled Im(FFT(kernel)) on the N grid
  float2* tw = nullptr;      // [N] exp(-2 pi i t/N)
  // coarse mesh
  long long cvol = 0, cnk = 0;
  float* r3 = nullptr; float* cforce = nullptr; float* kern_c = nullptr; float* fc = nullptr;
  cufftHandle cplan_r2c = 0, cplan_c2r = 0;
  // ---- more than one image (cube_comm.cuh, cube_exchange.cuh, cube_coarse.cuh) ----
  int nimg = 1;
  std::unique_ptr<Comm> comm;
  ExPlan ex;
  int *gcell_ext = nullptr, *scell_L = nullptr, *gcnt = nullptr, *scnt = nullptr;
  long long *gstart = nullptr, *sstart = nullptr, *dir_cell0 = nullptr, *dir_bounds = nullptr;
  HaloRec *hsend = nullptr, *hrecv = nullptr;
  void* psend = nullptr; long long sendcap = 0;
  std::vector<long long> gbound, sbound;  // host copies of the message offsets [ndir+1]
  long long nghost = 0;
  double* stat_partial_g = nullptr;
  CoarseGeom cg = {};
  float *stageA = nullptr, *slabR = nullptr, *kernT = nullptr, *sendF = nullptr, *recvF = nullptr;
  float2 *slabC = nullptr, *packT = nullptr, *T = nullptr, *T3 = nullptr;
  cufftHandle p2d_r2c = 0, p2d_c2r = 0, pz = 0;
  struct FTarget { int rank, icx, icy; std::vector<int> zloc; long long off; int* d_zloc; };
  struct FSource { int rank; long long off, nplanes; };
  std::vector<FTarget> ftargets;   // images that need planes of my z-slab (interior + halo)
  std::vector<FSource> fsources;   // slab owners my force_c planes come from
  long long *zzoff = nullptr, *zzcs = nullptr;
  // profiling
  cudaEvent_t tev[2] = {};
  bool prof = false; cudaEvent_t ev[2 * PH_N] = {}; float phase_ms[PH_N] = {}; long long launches = 0;
  float last_f2max_fine = 0; int last_radius = 0;
};

struct PhaseTimer {  // CUBEnu-style phase bracket (pm.f90:35,195,...) with CUDA events on the launch stream
  cube_handle* h; int ph;
  PhaseTimer(cube_handle* h_, int ph_) : h(h_), ph(ph_) { if (h->prof) cudaEventRecord(h->ev[2 * ph], h->st); }
  ~PhaseTimer() {
    if (h->prof) {
      cudaEventRecord(h->ev[2 * ph + 1], h->st);
      cudaEventSynchronize(h->ev[2 * ph + 1]);
      float ms = 0; cudaEventElapsedTime(&ms, h->ev[2 * ph], h->ev[2 * ph + 1]);
      h->phase_ms[ph] += ms;
    }
  }
};

static thread_local cube_handle* g_cur = nullptr;  // the handle of the API call in progress on this thread
static void abort_current_comm() { if (g_cur && g_cur->comm) g_cur->comm->abort_group(); }
template <class T> static cudaError_t dmalloc(T** p, long long n) { return cudaMalloc((void**)p, (size_t)(n > 0 ? n : 1) * sizeof(T)); }
static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// exclusive scan int32 -> int64, total returned in h->bsum[nb] (device)
static int scan_counts(cube_handle* h, const int* in, long long n, long long* out) {
  int nb = (int)((n + SCAN_B - 1) / SCAN_B);
  k_scan_blocksum<<<nb, SCAN_T, 0, h->st>>>(in, n, h->bsum); CKL();
  k_scan_bsums<<<1, 1024, 0, h->st>>>(h->bsum, nb); CKL();
  k_scan_final<<<nb, SCAN_T, 0, h->st>>>(in, n, h->bsum, out); CKL();
  CK(cudaMemcpyAsync(out + n, h->bsum + nb, sizeof(long long), cudaMemcpyDeviceToDevice, h->st));  // sentinel = total
  h->launches += 3;
  return 0;
}

static int build_dvlut(cube_handle* h, float sigma) {
  if (h->lut_sigma == sigma) return 0;
  CK(cudaMemsetAsync(h->divok, 0xff, sizeof(int), h->st));  // cleared by the kernel if the FMA division misses t/S anywhere
  k_build_dvlut<<<nblk(h->nvbin, 256), 256, 0, h->st>>>(h->nvbin, h->tanlut, vscale(sigma), 1.0 / vscale(sigma), h->dvlut, h->divok); CKL();
  h->launches++;
  h->lut_sigma = sigma;
  return 0;
}
static VTab vtab(const cube_handle* h) { return VTab{h->tanlut, h->tanh, h->enc, h->dvlut, h->divok, h->vt_hot}; }
// one 1024-thread CTA per SM, fewer when there are not enough warp chunks
static unsigned pw_grid(const cube_handle* h, long long ncells) {
  const long long nwc = (ncells + WC - 1) / WC;
  return (unsigned)std::max<long long>(1, std::min<long long>(h->nsm, (nwc + PW_W - 1) / PW_W));
}


using Fd888 = FdCfg<8, 8, 8, 1024>;
using Fd884 = FdCfg<8, 8, 4, 512>;
using Fd844 = FdCfg<8, 4, 4, 256>;

// A batch of tiles [tile0, tile0+nb) must be a box of the tile grid (x fastest): part of a row, whole rows of one layer, or whole
// layers.  Largest batch <= want with that property for every tile0 that is a multiple of it.
static int align_batch(int nnt, int want) {
  const int layer = nnt * nnt;
  if (want >= layer) return want / layer * layer;
  int best = 1;
  if (want >= nnt) { for (int k = 1; k <= want / nnt; k++) if (nnt % k == 0) best = k * nnt; return best; }
  for (int k = 1; k <= want; k++) if (nnt % k == 0) best = k;
  return best;
}
// tile box of an aligned batch: first tile and extent per dimension
static void batch_box(const Geom& g, int tile0, int nb, int t0[3], int k[3]) {
  const int nnt = g.nnt, layer = nnt * nnt;
  t0[0] = tile0 % nnt; t0[1] = (tile0 / nnt) % nnt; t0[2] = tile0 / layer;
  if (nb >= layer) { k[0] = nnt; k[1] = nnt; k[2] = (nb + layer - 1) / layer; }
  else if (nb >= nnt) { k[0] = nnt; k[1] = nb / nnt; k[2] = 1; }
  else { k[0] = nb; k[1] = 1; k[2] = 1; }
}
// fine-density region of a batch: the tiles' FFT windows (N nodes from 16 nodes below each tile) on one grid
static FineRegion batch_region(const cube_handle* h, int tile0, int nb) {
  const Geom& g = h->g;
  int t0[3], k[3];
  batch_box(g, tile0, nb, t0, k);
  FineRegion R;
  for (int d = 0; d < 3; d++) {
    R.c0[d] = t0[d] * g.nt - (NCB - 1); R.c1[d] = (t0[d] + k[d]) * g.nt + (NCB - 1);  // cells 2-ncb .. nt+ncb-1 of pm.f90:50-52
    R.f0[d] = 4 * (t0[d] * g.nt) - 16;
    R.n[d] = 4 * g.nt * (k[d] - 1) + h->fg.N;
  }
  R.ldy = R.n[0]; R.ldz = R.ldy * R.n[1];
  return R;
}
static size_t region_elems(const cube_handle* h, int batch) {
  if (!h->shared_region) return (size_t)h->fg.N * h->fg.N * h->fg.N * batch;
  const FineRegion R = batch_region(h, 0, batch);
  return (size_t)R.ldz * R.n[2];
}


// TMA descriptor of the fine force F[batch][M][M][3][FP] (f32): the kick fetches a brick's 36 x 3 x 9 x 9 nodes with one box copy
static int make_force_map(cube_handle* h) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    if (h->kick_stage == 2) h->kick_stage = 1;  // no TMA descriptors on this driver: cp.async staging
    return 0;
  }
  const FftGeom& f = h->fg;
  const cuuint64_t dims[5] = {(cuuint64_t)f.FP, 3, (cuuint64_t)f.M, (cuuint64_t)f.M, (cuuint64_t)h->batch};
  const cuuint64_t strides[4] = {(cuuint64_t)f.FP * 4, (cuuint64_t)f.FP * 12, (cuuint64_t)f.M * f.FP * 12, (cuuint64_t)f.M * f.M * f.FP * 12};
  const cuuint32_t box[5] = {(cuuint32_t)KB_NX, 3, (cuuint32_t)KB_NY, (cuuint32_t)KB_NZ, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUresult r = ((EncodeFn)fn)(&h->fmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, h->F, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cube_gpu_init: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// fine kick of the tiles [tile0, tile0+nb) with the force in h->F (pm.f90:88-118); codes come in with h->dvlut's sigma
// `hm`: whose force mesh (a second species is kicked by the first species' h->F / h->fc)
static int run_fine_kick(cube_handle* h, int tile0, int nb, double S_new, cube_handle* hm = nullptr) {
  if (!hm) hm = h;
  const Geom& g = h->g;
  const long long nt3 = (long long)g.nt * g.nt * g.nt;
  cudaStream_t st = hm->st;
  if (hm == h && h->kick_stage == 2 && h->zx == 2 && h->zv == 2 && getenv("CUBE_GPU_FKICK_BRICK")) {
    const int bpt = ((g.nt + KB_X - 1) / KB_X) * ((g.nt + KB_Y - 1) / KB_Y) * ((g.nt + KB_Z - 1) / KB_Z);
    k_fine_kick_brick<<<dim3(bpt, nb), FKB_T, 0, h->st>>>(g, tile0, h->fg.M, (const short*)h->xp, (short*)h->vp, h->cstart_p, h->dvlut, h->enc, S_new, h->fmap);
  } else {
    FMT_SWITCH(h, k_fine_kick_p<F><<<dim3(nblk(nt3, PC_CELLS), nb), PC_T, 0, st>>>(g, tile0, hm->fg.M, hm->fg.FP, XPC(h->xp), VPM(h->vp), h->cstart_p, hm->F, h->dvlut, h->enc, S_new));
  }
  CKL();
  h->launches++;
  return 0;
}
// coarse kick of the file-order cells [c_begin, c_end) with the force in h->fc (pm.f90:196-228)
static int run_coarse_kick(cube_handle* h, const VTab& vt, double S, long long c_begin, long long c_end, cube_handle* hm = nullptr) {
  if (!hm) hm = h;
  FMT_SWITCH(h, k_coarse_kick_w<F><<<pw_grid(h, c_end - c_begin), PW_T, pw_smem_bytes(h->vt_hot), hm->st>>>(h->g, vt, S, XPC(h->xp), VPM(h->vp), h->cstart_p, h->vfield_p, hm->fc,
                                                                                                        h->vmax_bits, c_begin, c_end));
  CKL();
  h->launches++;
  return 0;
}

// fine and/or coarse kick of the tiles [tile0, tile0+nb) in one pass (cube_kick.cuh); F = force_f of those tiles (nullptr: none),
// Gc = force_c (nullptr: none), both already multiplied by a_mid*dt/6/pi
static int launch_kick(cube_handle* h, int tile0, int nb, const float* F, const float* Gc, double S_in, double S_out) {
  const Geom& g = h->g;
  KickArgs A;
  A.g = g; A.tile0 = tile0; A.nb = nb; A.M = h->fg.M; A.FP = h->fg.FP; A.F = F; A.Gc = Gc;
  A.xp = (const short*)h->xp; A.vp = (short*)h->vp; A.cstart_p = h->cstart_p; A.vfield_p = h->vfield_p;
  A.vt_in = VTab{h->tanlut, h->tanh, h->enc, h->dvlut, h->divok, h->kb_hot}; A.S_in = S_in;
  A.vt_out = VTab{h->tanlut, h->tanh, h->enc, h->dvlut2, h->divok2, h->kb_hot}; A.S_out = S_out;
  A.vmax_bits = h->vmax_bits; A.next_brick = h->kick_next;
  const long long nbrick = (long long)nb * ((g.nt + KB_X - 1) / KB_X) * ((g.nt + KB_Y - 1) / KB_Y) * ((g.nt + KB_Z - 1) / KB_Z);
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(h->nsm, (nbrick + KB_G - 1) / KB_G));
  CK(cudaMemsetAsync(h->kick_next, 0, sizeof(unsigned long long), h->st));
  if (h->kick_stage == 2) k_kick_brick<2><<<grid, KB_T, kb_smem_bytes(h->kb_hot), h->st>>>(A, h->fmap);
  else if (h->kick_stage == 1) k_kick_brick<1><<<grid, KB_T, kb_smem_bytes(h->kb_hot), h->st>>>(A, h->fmap);
  else k_kick_brick<0><<<grid, KB_T, kb_smem_bytes(h->kb_hot), h->st>>>(A, h->fmap);
  CKL();
  h->launches++;
  return 0;
}
// decode table of `sigma` in dvlut2/divok2 (the codes the fine kick writes)
static int build_dvlut2(cube_handle* h, float sigma) {
  const double S = vscale(sigma);
  CK(cudaMemsetAsync(h->divok2, 0xff, sizeof(int), h->st));
  k_build_dvlut<<<nblk(h->nvbin, 256), 256, 0, h->st>>>(h->nvbin, h->tanlut, S, 1.0 / S, h->dvlut2, h->divok2); CKL();
  h->launches++;
  return 0;
}

extern "C" const char* cube_gpu_last_error(void) { return g_err.c_str(); }

extern "C" int cube_gpu_nccl_unique_id(void* id128) {
  if (!id128) return fail("null argument");
  NcclApi& a = nccl_api();
  if (!a.load()) return fail("cube_gpu_nccl_unique_id: %s", a.err.c_str());
  ncclUniqueId id;
  ncclResult_t r = a.GetUniqueId(&id);
  if (r != ncclSuccess) return fail("ncclGetUniqueId: %s", a.GetErrorString(r));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id128, &id, 128);
  return 0;
}



// parameters.f90:178-203 (geometry) generalised to an nn[0] x nn[1] x nn[2] image grid
static void fill_geom(const cube_params* p, Geom& g) {
  for (int d = 0; d < 3; d++) { g.nn[d] = p->nn[d]; }
  g.ic[0] = p->rank % p->nn[0]; g.ic[1] = (p->rank / p->nn[0]) % p->nn[1]; g.ic[2] = p->rank / (p->nn[0] * p->nn[1]);
  g.nnt = p->nnt; g.nc = p->nc; g.nt = p->nc / p->nnt;
  g.nte = g.nt + 2 * NCB; g.nft = g.nt * NCELL; g.nfe = g.nft + 2 * NFB; g.ne = g.nc + 2 * NCB;
  g.ncell_p = (long long)g.nc * g.nc * g.nc; g.ncell_e = (long long)g.ne * g.ne * g.ne;
}
static void fill_coarse_geom(const Geom& g, CoarseGeom& c) {
  c.R = g.nn[0] * g.nn[1] * g.nn[2]; c.nc = g.nc;
  c.Gx = g.nc * g.nn[0]; c.Gy = g.nc * g.nn[1]; c.Gz = g.nc * g.nn[2]; c.KX = c.Gx / 2 + 1;
  c.grp = g.nn[0] * g.nn[1]; c.grp0 = g.ic[2] * c.grp;
  c.sz = c.Gz / c.R; c.nyl = c.Gy / c.R;
}
// last hop of the inverse coarse transform: planes zz = -1..nc (halo included) of image m live on slab owner Z/sz
struct FPlane { int rank; std::vector<int> zz; };
static void force_plane_lists(const Geom& g, const CoarseGeom& c, int my_rank, std::vector<FPlane>& targets, std::vector<FPlane>& sources) {
  targets.clear(); sources.clear();
  for (int m = 0; m < c.R; m++) {  // what image m needs from my slab (stored as my local plane index)
    const int mz = m / (g.nn[0] * g.nn[1]);
    FPlane t; t.rank = m;
    for (int zz = -1; zz <= g.nc; zz++) {
      const int Z = ((mz * g.nc + zz) % c.Gz + c.Gz) % c.Gz;
      if (Z / c.sz == my_rank) t.zz.push_back(Z - my_rank * c.sz);
    }
    if (!t.zz.empty()) targets.push_back(t);
  }
  for (int q = 0; q < c.R; q++) {  // where my own planes come from (stored as zz)
    FPlane sr; sr.rank = q;
    for (int zz = -1; zz <= g.nc; zz++) {
      const int Z = ((g.ic[2] * g.nc + zz) % c.Gz + c.Gz) % c.Gz;
      if (Z / c.sz == q) sr.zz.push_back(zz);
    }
    if (!sr.zz.empty()) sources.push_back(sr);
  }
}

// ---------------------------------------------------------------------------------------------
// more than one image: exchange plan, distributed coarse FFT
// ---------------------------------------------------------------------------------------------
static int comm_fail(cube_handle* h) {
  h->comm->abort_group();
  return fail("image %d: %s", h->p.rank + 1, h->comm->err.c_str());
}
#define CC(call) do { if ((call)) return comm_fail(h); } while (0)

static int init_exchange(cube_handle* h) {
  const Geom& g = h->g;
  build_exchange_plan(g, h->p.rank, h->ex, true);
  const long long ng = h->ex.ng;
  const int nd = (int)h->ex.dirs.size();
  if (ng == 0) return 0;
  CK(dmalloc(&h->gcell_ext, ng)); CK(dmalloc(&h->scell_L, ng)); CK(dmalloc(&h->gcnt, ng)); CK(dmalloc(&h->scnt, ng));
  CK(dmalloc(&h->gstart, ng + 1)); CK(dmalloc(&h->sstart, ng + 1));
  CK(dmalloc(&h->hsend, ng)); CK(dmalloc(&h->hrecv, ng));
  CK(dmalloc(&h->dir_cell0, nd + 1)); CK(dmalloc(&h->dir_bounds, 2 * (nd + 1)));
  CK(cudaMemcpyAsync(h->gcell_ext, h->ex.gcell_ext.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->scell_L, h->ex.scell_L.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, h->st));
  std::vector<long long> c0(nd + 1);
  for (int i = 0; i < nd; i++) c0[i] = h->ex.dirs[i].cell0;
  c0[nd] = ng;
  CK(cudaMemcpyAsync(h->dir_cell0, c0.data(), sizeof(long long) * (nd + 1), cudaMemcpyHostToDevice, h->st));
  CK(h->rb.sync(h->st));
  h->ex.gcell_ext.clear(); h->ex.gcell_ext.shrink_to_fit(); h->ex.scell_L.clear(); h->ex.scell_L.shrink_to_fit();
  h->gbound.assign(nd + 1, 0); h->sbound.assign(nd + 1, 0);
  // message buffer for the particles I send: mean occupancy of the send cells with the image_buffer margin, x2
  const double mean = (double)h->p.np_nc * h->p.np_nc * h->p.np_nc;
  h->sendcap = (long long)((double)ng * mean * (double)h->p.image_buffer * 2.0) + 4096;
  CK(cudaMalloc(&h->psend, (size_t)3 * std::max(h->zx, h->zv) * h->sendcap + 16));
  CK(dmalloc(&h->stat_partial_g, 2 * (long long)nblk(ng, PC_CELLS) + 2));
  return 0;
}

static int init_coarse_dist(cube_handle* h) {
  const Geom& g = h->g;
  CoarseGeom& c = h->cg;
  fill_coarse_geom(g, c);
  if (c.Gz % c.R || c.Gy % c.R) return fail("cube_gpu_init: nc=%d must be a multiple of nnx*nny=%d and of nnx*nnz=%d for the distributed coarse FFT", g.nc, g.nn[0] * g.nn[1], g.nn[0] * g.nn[2]);
  const long long slab = (long long)c.sz * c.Gy * c.Gx, slabk = (long long)c.sz * c.Gy * c.KX, nk = (long long)c.Gz * c.nyl * c.KX;
  CK(dmalloc(&h->stageA, slab)); CK(dmalloc(&h->slabR, 3 * slab)); CK(dmalloc(&h->slabC, 3 * slabk)); CK(dmalloc(&h->packT, 3 * slabk));
  CK(dmalloc(&h->T, nk)); CK(dmalloc(&h->T3, 3 * nk)); CK(dmalloc(&h->kernT, 3 * nk));
  {
    int n2[2] = {c.Gy, c.Gx};
    CF(cufftPlanMany(&h->p2d_r2c, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_R2C, c.sz));
    CF(cufftPlanMany(&h->p2d_c2r, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_C2R, 3 * c.sz));
    int n1[1] = {c.Gz};
    const int lines = c.nyl * c.KX;
    CF(cufftPlanMany(&h->pz, 1, n1, n1, lines, 1, n1, lines, 1, CUFFT_C2C, lines));
    CF(cufftSetStream(h->p2d_r2c, h->st)); CF(cufftSetStream(h->p2d_c2r, h->st)); CF(cufftSetStream(h->pz, h->st));
  }
  const int m2 = g.nc + 2;
  std::vector<FPlane> targets, sources;
  force_plane_lists(g, c, h->p.rank, targets, sources);
  long long off = 0;
  for (const FPlane& fp : targets) {
    cube_handle::FTarget t; t.rank = fp.rank; t.icx = fp.rank % g.nn[0]; t.icy = (fp.rank / g.nn[0]) % g.nn[1]; t.off = off; t.d_zloc = nullptr;
    t.zloc = fp.zz;
    CK(dmalloc(&t.d_zloc, (long long)t.zloc.size()));
    CK(cudaMemcpyAsync(t.d_zloc, t.zloc.data(), sizeof(int) * t.zloc.size(), cudaMemcpyHostToDevice, h->st));
    off += 3LL * (long long)t.zloc.size() * m2 * m2;
    h->ftargets.push_back(t);
  }
  CK(dmalloc(&h->sendF, off));
  std::vector<long long> zzoff(m2), zzcs(m2);
  off = 0;
  for (const FPlane& fp : sources) {
    cube_handle::FSource sr; sr.rank = fp.rank; sr.off = off; sr.nplanes = (long long)fp.zz.size();
    for (size_t i = 0; i < fp.zz.size(); i++) { zzoff[fp.zz[i] + 1] = off + (long long)i * m2 * m2; zzcs[fp.zz[i] + 1] = sr.nplanes * m2 * m2; }
    off += 3LL * sr.nplanes * m2 * m2;
    h->fsources.push_back(sr);
  }
  CK(dmalloc(&h->recvF, off));
  CK(dmalloc(&h->zzoff, m2)); CK(dmalloc(&h->zzcs, m2));
  CK(cudaMemcpyAsync(h->zzoff, zzoff.data(), sizeof(long long) * m2, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->zzcs, zzcs.data(), sizeof(long long) * m2, cudaMemcpyHostToDevice, h->st));
  CK(h->rb.sync(h->st));
  return 0;
}

// forward transform of this image's block in[nc][nc][nc] (unpadded) -> h->T[kz][kyl][kx]
static int coarse_forward(cube_handle* h, const float* in) {
  const CoarseGeom& c = h->cg;
  const Geom& g = h->g;
  Comm* cm = h->comm.get();
  const size_t blk = (size_t)c.sz * g.nc * g.nc;
  CC(cm->begin(h->st));
  for (int j = 0; j < c.grp; j++) CC(cm->send(in + (size_t)j * blk, blk * sizeof(float), c.grp0 + j));
  for (int j = 0; j < c.grp; j++) CC(cm->recv(h->stageA + (size_t)j * blk, blk * sizeof(float), c.grp0 + j));
  CC(cm->end());
  k_slab_assemble<<<1184, 256, 0, h->st>>>(c, g.nn[0], h->stageA, h->slabR); CKL();
  CF(cufftExecR2C(h->p2d_r2c, h->slabR, (cufftComplex*)h->slabC));
  k_pack_T<<<1184, 256, 0, h->st>>>(c, h->slabC, h->packT); CKL();
  const size_t seg = (size_t)c.sz * c.nyl * c.KX;
  CC(cm->begin(h->st));
  for (int q = 0; q < c.R; q++) CC(cm->send(h->packT + (size_t)q * seg, seg * sizeof(float2), q));
  for (int q = 0; q < c.R; q++) CC(cm->recv(h->T + (size_t)q * seg, seg * sizeof(float2), q));
  CC(cm->end());
  CF(cufftExecC2C(h->pz, (cufftComplex*)h->T, (cufftComplex*)h->T, CUFFT_FORWARD));
  h->launches += 2;
  return 0;
}

// T -> i*kern_c*T/(Gx Gy Gz) -> three inverse transforms -> force planes with halo in h->recvF
static int coarse_backward3(cube_handle* h) {
  const CoarseGeom& c = h->cg;
  const Geom& g = h->g;
  Comm* cm = h->comm.get();
  const long long nk = (long long)c.Gz * c.nyl * c.KX;
  const float scale = 1.0f / ((float)g.nc * g.nn[0]) / ((float)g.nc * g.nn[1]) / ((float)g.nc * g.nn[2]);
  k_green_T<<<nblk(nk, 256), 256, 0, h->st>>>(nk, h->T, h->kernT, scale, h->T3); CKL();
  for (int d = 0; d < 3; d++) CF(cufftExecC2C(h->pz, (cufftComplex*)(h->T3 + d * nk), (cufftComplex*)(h->T3 + d * nk), CUFFT_INVERSE));
  const size_t seg = (size_t)c.sz * c.nyl * c.KX;
  CC(cm->begin(h->st));
  for (int q = 0; q < c.R; q++)
    for (int d = 0; d < 3; d++) CC(cm->send(h->T3 + (size_t)d * nk + (size_t)q * seg, seg * sizeof(float2), q));
  for (int q = 0; q < c.R; q++)
    for (int d = 0; d < 3; d++) CC(cm->recv(h->packT + ((size_t)q * 3 + d) * seg, seg * sizeof(float2), q));
  CC(cm->end());
  k_unpack_T<<<1184, 256, 0, h->st>>>(c, h->packT, h->slabC); CKL();
  CF(cufftExecC2R(h->p2d_c2r, (cufftComplex*)h->slabC, h->slabR));
  const int m2 = g.nc + 2;
  for (auto& t : h->ftargets) {
    k_pack_F<<<592, 256, 0, h->st>>>(c, t.icx, t.icy, (int)t.zloc.size(), t.d_zloc, h->slabR, h->sendF + t.off); CKL();
    h->launches++;
  }
  CC(cm->begin(h->st));
  for (auto& t : h->ftargets) CC(cm->send(h->sendF + t.off, 3 * t.zloc.size() * (size_t)m2 * m2 * sizeof(float), t.rank));
  for (auto& sr : h->fsources) CC(cm->recv(h->recvF + sr.off, 3 * (size_t)sr.nplanes * m2 * m2 * sizeof(float), sr.rank));
  CC(cm->end());
  h->launches += 2;
  return 0;
}

// kernel_c.f90:16-117 on the global lattice, k-space in the transposed layout
static int build_kernel_c_dist(cube_handle* h, const float* d_ck) {
  const Geom& g = h->g;
  const CoarseGeom& c = h->cg;
  const long long n = (long long)g.nc * g.nc * g.nc, nk = (long long)c.Gz * c.nyl * c.KX;
  float* blk = h->r3;  // [nc]^3 scratch
  const int ox0 = g.ic[0] * g.nc, oy0 = g.ic[1] * g.nc, oz0 = g.ic[2] * g.nc;
  for (int d = 0; d < 3; d++) {
    k_kernc_fill_dist<<<nblk(n, 256), 256, 0, h->st>>>(c, ox0, oy0, oz0, d_ck, d, 1, blk); CKL();
    if (coarse_forward(h, blk)) return 1;
    k_take_imag_scaled<<<nblk(nk, 256), 256, 0, h->st>>>(nk, h->T, h->kernT + d * nk); CKL();
    k_kernc_fill_dist<<<nblk(n, 256), 256, 0, h->st>>>(c, ox0, oy0, oz0, d_ck, d, 0, blk); CKL();
    if (coarse_forward(h, blk)) return 1;
    k_kernc_lrck_T<<<nblk(nk, 256), 256, 0, h->st>>>(c, h->p.rank, d, h->T, h->kernT + d * nk); CKL();
  }
  CK(h->rb.sync(h->st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
static int build_kernels(cube_handle* h, const float* fk_table, const float* ck_table) {
  const Geom& g = h->g;
  float *d_fk = nullptr, *d_ck = nullptr;
  CK(dmalloc(&d_fk, 16 * 16 * 16 * 3)); CK(dmalloc(&d_ck, 3 * 64));
  CK(cudaMemcpyAsync(d_fk, fk_table, sizeof(float) * 16 * 16 * 16 * 3, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(d_ck, ck_table, sizeof(float) * 3 * 64, cudaMemcpyHostToDevice, h->st));
  // kernel_f.f90:32-41 on the N-point window: mirrored 16^3 table, r2c (cuFFT, init only), keep Im
  {
    const int N = h->fg.N;
    const long long vol = (long long)N * N * (N + 2), nk = (long long)N * N * (N / 2 + 1);
    float* tmp = nullptr; CK(dmalloc(&tmp, vol));
    cufftHandle pl = 0;
    CF(cufftPlan3d(&pl, N, N, N, CUFFT_R2C)); CF(cufftSetStream(pl, h->st));
    CK(cudaMemsetAsync(h->kern_f, 0, sizeof(float) * 3 * (size_t)N * N * h->fg.P, h->st));
    for (int d = 0; d < 3; d++) {
      k_kernf_fill<<<nblk(vol, 256), 256, 0, h->st>>>(N, d_fk, d, tmp); CKL();
      CF(cufftExecR2C(pl, tmp, (cufftComplex*)tmp));
      k_take_imag_pitched<<<nblk(nk, 256), 256, 0, h->st>>>(N, h->fg.P, (const float2*)tmp, h->kern_f + (size_t)d * N * N * h->fg.P); CKL();
    }
    CK(h->rb.sync(h->st));
    cufftDestroy(pl); cudaFree(tmp);
  }
  if (h->nimg > 1) {
    const int rc = build_kernel_c_dist(h, d_ck);
    cudaFree(d_fk); cudaFree(d_ck);
    return rc;
  }
  // kernel_c.f90:16-117 (single image: the coarse lattice is this image's nc^3)
  float* pure = h->cforce;  // scratch [cvol]
  for (int d = 0; d < 3; d++) {
    k_kernc_fill<<<nblk(h->cvol, 256), 256, 0, h->st>>>(g.nc, d_ck, d, 1, h->r3); CKL();
    CF(cufftExecR2C(h->cplan_r2c, h->r3, (cufftComplex*)h->r3));
    k_take_imag<<<nblk(h->cnk, 256), 256, 0, h->st>>>(h->cnk, (const float2*)h->r3, h->kern_c + d * h->cnk); CKL();
    k_kernc_fill<<<nblk(h->cvol, 256), 256, 0, h->st>>>(g.nc, d_ck, d, 0, pure); CKL();
    CF(cufftExecR2C(h->cplan_r2c, pure, (cufftComplex*)pure));
    k_kernc_lrck<<<nblk(h->cnk, 256), 256, 0, h->st>>>(g.nc, d, (const float2*)pure, h->kern_c + d * h->cnk); CKL();
  }
  CK(h->rb.sync(h->st));
  cudaFree(d_fk); cudaFree(d_ck);
  return 0;
}

extern "C" int cube_gpu_init(const cube_params* p, const float* fk_table, const float* ck_table, const float* tanf_lut,
                             const void* nccl_unique_id, cube_handle** out) {
  if (!p || !fk_table || !ck_table || !tanf_lut || !out) return fail("cube_gpu_init: null argument");
  if ((p->izipx != 1 && p->izipx != 2) || (p->izipv != 1 && p->izipv != 2))
    return fail("zip format incompatable: izipx and izipv are 1 or 2 bytes (got %d,%d)", p->izipx, p->izipv);
  if (p->ncell != NCELL || p->ncb != NCB) return fail("cube_gpu_init: ncell must be 4 and ncb 6");
  const int nimg = p->nn[0] * p->nn[1] * p->nn[2];
  if (p->nn[0] < 1 || p->nn[1] < 1 || p->nn[2] < 1 || p->rank < 0 || p->rank >= nimg) return fail("cube_gpu_init: bad image grid / rank");
  if (nimg > 1 && !nccl_unique_id && p->local_group <= 0)
    return fail("cube_gpu_init: %d images need either an ncclUniqueId (one process per GPU) or local_group > 0 (all images are threads of this process)", nimg);
  if (p->nnt < 1 || p->nc % p->nnt) return fail("cube_gpu_init: nc must be a multiple of nnt");
  if (p->nc / p->nnt < 12 || p->nc < 24) return fail("cube_gpu_init: need nc>=24 and nt>=12 (parameters.f90:23-24)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("cube_gpu_init: no CUDA device (there is no CPU fallback)");
  CK(cudaSetDevice(p->device));
  cube_handle* h = new cube_handle();
  h->p = *p;
  h->zx = p->izipx; h->zv = p->izipv; h->nvbin = 1 << (8 * p->izipv);
  h->nimg = nimg;
  Geom& g = h->g;
  fill_geom(p, g);
  // variables.f90:7-9 (real(4) arithmetic)
  long long np_image = (long long)g.nc * p->np_nc; np_image = np_image * np_image * np_image;
  float r = ((float)g.nte * 1.f) / (float)g.nt, r3 = r * r * r;
  h->np_image_max = (long long)((float)np_image * r3 * p->image_buffer);
  h->np_tile_max = (long long)((float)(np_image / ((long long)g.nnt * g.nnt * g.nnt)) * r3 * p->tile_buffer);
  CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking));
  CK(h->rb.init());
  {  // highest priority: the coarse stream's small kernels (and exchange kernels) take the next CTA slots that free up
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->st_coarse, cudaStreamNonBlocking, getenv("CUBE_GPU_COARSE_PRIO0") ? lo : hi));
  }
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_vpack, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_vghost, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_xghost, cudaEventDisableTiming));
  for (auto& e : h->ev_up) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  h->st_main = h->st;
  h->async_vghost = getenv("CUBE_GPU_SYNC_BUFFER_V") == nullptr;
  h->async_xghost = getenv("CUBE_GPU_SYNC_BUFFER_X") == nullptr;
  h->overlap_coarse = getenv("CUBE_GPU_NO_OVERLAP") == nullptr;  // multi-image runs: decided after the communicator exists (below)
  CK(cudaEventCreateWithFlags(&h->ev_copy[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_copy[1], cudaEventDisableTiming));
  for (int i = 0; i < 2 * PH_N; i++) CK(cudaEventCreate(&h->ev[i]));
  CK(cudaEventCreate(&h->tev[0])); CK(cudaEventCreate(&h->tev[1]));
  const long long cap = h->np_image_max;
  CK(cudaMalloc(&h->xp, (size_t)3 * h->zx * cap + 16)); CK(cudaMalloc(&h->vp, (size_t)3 * h->zv * cap + 16));
  CK(cudaMalloc(&h->xp2, (size_t)3 * h->zx * cap + 16)); CK(cudaMalloc(&h->vp2, (size_t)3 * h->zv * cap + 16));
  CK(dmalloc(&h->key, cap));
  CK(dmalloc(&h->rhoc_p, g.ncell_p)); CK(dmalloc(&h->rhoc_p2, g.ncell_p));
  CK(dmalloc(&h->vfield_p, 3 * g.ncell_p)); CK(dmalloc(&h->vfield_p2, 3 * g.ncell_p));
  CK(dmalloc(&h->cstart_p, g.ncell_p + 1)); CK(dmalloc(&h->cstart_p2, g.ncell_p + 1));  // + sentinel = nplocal
  CK(dmalloc(&h->rhoc_e, g.ncell_e)); CK(dmalloc(&h->cstart_e, g.ncell_e)); CK(dmalloc(&h->vfield_e, 3 * g.ncell_e));
  if (nimg > 1) {
    if (nccl_unique_id) {
      auto c = std::make_unique<NcclComm>();
      if (c->init(p->rank, nimg, nccl_unique_id)) return fail("cube_gpu_init: %s", c->err.c_str());
      h->comm = std::move(c);
    } else {
      auto c = std::make_unique<LocalComm>();
      if (c->init(p->rank, nimg, p->local_group)) return fail("cube_gpu_init: %s", c->err.c_str());
      h->comm = std::move(c);
    }
    if (init_exchange(h)) return 1;
  }
  CK(dmalloc(&h->sid_e, g.ncell_e)); CK(dmalloc(&h->mask_e, MASK_W * g.ncell_e)); CK(dmalloc(&h->mask_s, MASK_W * (g.ncell_p + h->ex.ng)));
  CK(dmalloc(&h->farblk, (long long)farblk_dim(g) * farblk_dim(g) * farblk_dim(g)));
  CK(dmalloc(&h->inflag, g.ncell_p)); CK(dmalloc(&h->flist, g.ncell_p)); CK(dmalloc(&h->nflag, 1));
  h->count_all = getenv("CUBE_GPU_COUNT_ALL") != nullptr;

  h->nscan_blocks = (int)((std::max(g.ncell_p, h->ex.ng) + SCAN_B - 1) / SCAN_B);
  CK(dmalloc(&h->bsum, h->nscan_blocks + 1));
  CK(dmalloc(&h->stat_partial, std::max<long long>(2 * 4096 * PW_W, (long long)nblk(g.ncell_p, 128)))); CK(dmalloc(&h->stat3, 8));
  CK(dmalloc(&h->rank, cap));
  CK(dmalloc(&h->tile_count, (long long)g.nnt * g.nnt * g.nnt));
  CK(dmalloc(&h->maxoff, 1)); CK(dmalloc(&h->vmax_bits, 4));
  const int nvb = h->nvbin, vh = nvb / 2;  // table sizes of this velocity format
  CK(dmalloc(&h->tanlut, nvb)); CK(dmalloc(&h->dvlut, nvb)); CK(dmalloc(&h->enc, vh));
  CK(dmalloc(&h->tanh, vh + 4)); CK(dmalloc(&h->divok, 1)); CK(dmalloc(&h->dvlut2, nvb)); CK(dmalloc(&h->divok2, 1));
  if (h->zv == 2) k_build_enc<16><<<nblk(vh, 256), 256, 0, h->st>>>(h->enc); else k_build_enc<8><<<nblk(vh, 256), 256, 0, h->st>>>(h->enc);
  CKL();
  CK(cudaMemcpyAsync(h->tanlut, tanf_lut, nvb * sizeof(float), cudaMemcpyHostToDevice, h->st));
  {
    // half table for shared memory: index = |code|.  tanf_lut is indexed by the code's raw pattern, so -c sits at nvbin-c.
    // The host tanf must be odd for this (glibc's is); if it is not, hot = 0 sends every lookup to the full global tables.
    std::vector<float> half(vh + 4, 0.f);
    bool odd = true;
    for (int c = 0; c <= vh - 1; c++) half[c] = tanf_lut[c];
    half[vh] = -tanf_lut[vh];
    for (int c = 1; c <= vh - 1; c++) { const float neg = -tanf_lut[nvb - c]; if (memcmp(&neg, &tanf_lut[c], 4) != 0) { odd = false; break; } }
    if (tanf_lut[0] != 0.f || std::signbit(tanf_lut[0])) odd = false;
    h->vt_hot = (odd && !getenv("CUBE_GPU_GLOBAL_TABLES")) ? vh : 0;
    if (const char* e = getenv("CUBE_GPU_VT_HOT")) { if (h->vt_hot) h->vt_hot = std::min(vh, std::max(4, atoi(e) & ~3)); }
    if (const char* e = getenv("CUBE_GPU_HEAVY_DEPOSIT")) h->heavy_deposit = atoi(e);
    if (const char* e = getenv("CUBE_GPU_HEAVY_COUNT")) h->heavy_count = atoi(e);
    if (const char* e = getenv("CUBE_GPU_FD_BRICK")) h->fd_brick = atoi(e);
    if (const char* e = getenv("CUBE_GPU_COUNT_MINB")) h->count_minb = atoi(e);
    CK(cudaMemcpyAsync(h->tanh, half.data(), (vh + 4) * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CK(h->rb.sync(h->st));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, p->device));
    h->nsm = prop.multiProcessorCount;
    FMT_SWITCH(h, CK(cudaFuncSetAttribute((const void*)k_drift_place_w<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_MAX));
               CK(cudaFuncSetAttribute((const void*)k_coarse_kick_w<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_MAX));
               CK(cudaFuncSetAttribute((const void*)k_drift_key_chain<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC_SMEM)));
    CK(cudaFuncSetAttribute((const void*)k_selftest_decode<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_MAX));
    CK(cudaFuncSetAttribute((const void*)k_selftest_decode<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, PW_SMEM_MAX));
    h->kb_hot = h->vt_hot ? std::min(KB_HOT_DEFAULT, vh) : 0;
    if (const char* e = getenv("CUBE_GPU_KICK_HOT")) { if (h->kb_hot) h->kb_hot = std::min(VT_ALL, std::max(4, atoi(e) & ~3)); }
    h->old_kick = getenv("CUBE_GPU_MERGED_KICK") == nullptr || h->zx != 2 || h->zv != 2;  // measured (profiles/r02_notes.md): the two separate kicks are faster
    if (const char* e = getenv("CUBE_GPU_KICK_STAGE")) h->kick_stage = atoi(e);
    CK(cudaFuncSetAttribute((const void*)k_kick_brick<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb_smem_bytes(h->kb_hot)));
    CK(cudaFuncSetAttribute((const void*)k_kick_brick<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb_smem_bytes(h->kb_hot)));
    CK(cudaFuncSetAttribute((const void*)k_kick_brick<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kb_smem_bytes(h->kb_hot)));
    CK(dmalloc(&h->kick_next, 1));
  }
  // fine mesh: pick the transform length, size the batch, allocate the pipeline arrays
  const int ntile = g.nnt * g.nnt * g.nnt;
  {
    const int need = g.nft + 32;  // M + 2*(nf_cutoff-1), M = nft+2
    // test hook: CUBE_GPU_NFFT=N forces the window length (any built N >= need gives the same forces; lets the parity tests
    // run every instantiation of the line-FFT kernels at a size the oracle finishes in seconds)
    const int forced = getenv("CUBE_GPU_NFFT") ? atoi(getenv("CUBE_GPU_NFFT")) : 0;
    for (const FftPlan& pl : kPlans)
      if (forced ? pl.N() == forced && forced >= need : (pl.N() >= need && (!h->plan || pl.N() < h->plan->N()))) h->plan = &pl;
    if (forced && !h->plan) return fail("cube_gpu_init: CUBE_GPU_NFFT=%d is not a built transform length >= %d", forced, need);
    if (!h->plan) return fail("cube_gpu_init: no fine-mesh FFT plan for nt=%d (needs N>=%d; largest built N is 576, i.e. nt<=136)", g.nt, need);
    FftGeom& f = h->fg;
    f.N = h->plan->N(); f.NH = f.N / 2 + 1; f.P = (f.NH + FL - 1) / FL * FL; f.M = g.nft + 2; f.off = 15; f.FP = (f.M + 7) / 8 * 8;
    h->rho_n = 0; h->A_n = (size_t)f.N * f.N * f.P; h->B_n = 3 * (size_t)f.M * f.N * f.P;
    h->F_n = (size_t)f.M * f.M * 3 * f.FP;
  }
  CK(dmalloc(&h->csum, 27LL * (g.nc + 2) * (g.nc + 2) * (g.nc + 2)));  // partial sums of the coarse deposit (cube_kernels.cuh)
  int batch = p->fine_batch;
  {
    const size_t per = h->A_n * sizeof(float2) + h->B_n * sizeof(float2) + h->F_n * sizeof(float);
    if (batch <= 0) {
      size_t fr = 0, tot = 0; CK(cudaMemGetInfo(&fr, &tot));
      batch = (int)std::max<long long>(1, std::min<long long>(64, (long long)(fr * 0.6) / (long long)per));
    }
  }
  if (p->reserved[0] & 1) batch = 1;  // a further species of a two-species run: kicked from the first species' meshes, never uses its own
  batch = align_batch(g.nnt, std::min(batch, ntile));
  h->batch = batch;
  // pm.f90:54-58 in f32 is exact below 512 fine cells of tile-local coordinate (cells up to nt+5): then the weights do not depend
  // on the tile frame and one density grid serves a whole batch of tiles
  h->shared_region = g.nt + 5 <= 128 && getenv("CUBE_GPU_TILE_REGIONS") == nullptr;
  h->rho_n = std::max(h->rho_n, region_elems(h, batch));
  // (the tensor map of F is encoded below, once F exists)
  h->fg.nbatch = batch;
  CK(dmalloc(&h->f2max, batch + 1));
  CK(dmalloc(&h->Ak, (long long)(h->A_n * batch))); CK(dmalloc(&h->Bk, (long long)(h->B_n * batch))); CK(dmalloc(&h->F, (long long)(h->F_n * batch) + 1024));  // + slack: the kick's 144-byte row copies may run past the last kept point
  if (h->rho_n * sizeof(float) <= h->B_n * batch * sizeof(float2)) h->rho = reinterpret_cast<float*>(h->Bk);  // dead before Bk is written
  else { CK(dmalloc(&h->rho_own, (long long)h->rho_n)); h->rho = h->rho_own; }  // forced long windows on small tiles (test hook)
  if (make_force_map(h)) return 1;
  CK(dmalloc(&h->kern_f, 3LL * h->fg.N * h->fg.N * h->fg.P));
  CK(dmalloc(&h->tw, h->fg.N));
  {
    const int N = h->fg.N;
    std::vector<float2> tw(N);
    for (int t = 0; t < N; t++) { double a = -2.0 * M_PI * t / N; tw[t] = make_float2((float)cos(a), (float)sin(a)); }
    CK(cudaMemcpyAsync(h->tw, tw.data(), sizeof(float2) * N, cudaMemcpyHostToDevice, h->st));
    CK(h->rb.sync(h->st));
    const int smem_x = (N * (FL + 1) + N) * (int)sizeof(float2), smem_y = (N * FL + N) * (int)sizeof(float2);
    const int smem_z = (2 * N * FL + N) * (int)sizeof(float2) + 3 * (N / 2 + 1) * FL * (int)sizeof(float);
    CK(cudaFuncSetAttribute((const void*)h->plan->x_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_x));
    CK(cudaFuncSetAttribute((const void*)h->plan->x_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->plan->x_inv_smem));
    CK(cudaFuncSetAttribute((const void*)h->plan->y_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_y));
    CK(cudaFuncSetAttribute((const void*)h->plan->y_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_y));
    // the inverse y pass is register-capped for four CTAs per SM (measured 6.1 -> 5.5 ms at cfg 2; the forward pass does not gain)
    CK(cudaFuncSetAttribute((const void*)h->plan->y_inv, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (auto zk : {h->plan->z_green, h->plan->z_green_pk, h->plan->z_green_mx}) {
      CK(cudaFuncSetAttribute((const void*)zk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_z));
      CK(cudaFuncSetAttribute((const void*)zk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }

  }
  // coarse mesh
  h->cvol = (long long)g.nc * g.nc * (g.nc + 2);
  h->cnk = (long long)g.nc * g.nc * (g.nc / 2 + 1);
  CK(dmalloc(&h->r3, h->cvol));
  CK(dmalloc(&h->fc, 3LL * (g.nc + 2) * (g.nc + 2) * (g.nc + 2)));
  if (nimg > 1) {
    if (init_coarse_dist(h)) return 1;
  } else {
    CK(dmalloc(&h->cforce, 3 * h->cvol)); CK(dmalloc(&h->kern_c, 3 * h->cnk));
    int n[3] = {g.nc, g.nc, g.nc};
    int rembed[3] = {g.nc, g.nc, g.nc + 2}, cembed[3] = {g.nc, g.nc, g.nc / 2 + 1};
    CF(cufftPlanMany(&h->cplan_r2c, 3, n, rembed, 1, (int)h->cvol, cembed, 1, (int)h->cnk, CUFFT_R2C, 1));
    CF(cufftPlanMany(&h->cplan_c2r, 3, n, cembed, 1, (int)h->cnk, rembed, 1, (int)h->cvol, CUFFT_C2R, 3));
    CF(cufftSetStream(h->cplan_r2c, h->st)); CF(cufftSetStream(h->cplan_c2r, h->st));
  }
#define FD_ATTR(C, XT)                                                                                                                          \
  CK(cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, false, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM)); \
  CK(cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, true, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM))
  if (h->zx == 2) { FD_ATTR(Fd888, short); FD_ATTR(Fd884, short); FD_ATTR(Fd844, short); }
  else { FD_ATTR(Fd888, signed char); FD_ATTR(Fd884, signed char); FD_ATTR(Fd844, signed char); }
#undef FD_ATTR
  if (build_kernels(h, fk_table, ck_table)) { return 1; }
  *out = h;
  return 0;
}

extern "C" int cube_gpu_finalize(cube_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->p.device);
  cudaStreamSynchronize(h->st);
  void* ptrs[] = {h->xp, h->vp, h->xp2, h->vp2, h->key, h->rhoc_p, h->rhoc_p2, h->vfield_p, h->vfield_p2, h->cstart_p, h->cstart_p2,
                  h->rhoc_e, h->cstart_e, h->vfield_e, h->sid_e, h->mask_s, h->mask_e, h->farblk, h->inflag, h->flist, h->nflag, h->csum, h->bsum, h->stat_partial, h->stat3, h->rank, h->tile_count, h->maxoff, h->f2max,
                  h->vmax_bits, h->tanlut, h->dvlut, h->enc, h->tanh, h->divok, h->dvlut2, h->divok2, h->kick_next, h->Ak, h->Bk, h->F, h->kern_f, h->tw, h->r3, h->cforce, h->kern_c, h->fc, h->pid, h->pid2, h->pid_send, h->rho_own};
  for (void* q : ptrs) if (q) cudaFree(q);
  void* mptrs[] = {h->gcell_ext, h->scell_L, h->gcnt, h->scnt, h->gstart, h->sstart, h->dir_cell0, h->dir_bounds, h->hsend, h->hrecv, h->psend,
                   h->stat_partial_g, h->stageA, h->slabR, h->kernT, h->sendF, h->recvF, h->slabC, h->packT, h->T, h->T3, h->zzoff, h->zzcs};
  for (void* q : mptrs) if (q) cudaFree(q);
  for (auto& t : h->ftargets) if (t.d_zloc) cudaFree(t.d_zloc);
  h->comm.reset();
  cufftHandle plans[] = {h->cplan_r2c, h->cplan_c2r, h->p2d_r2c, h->p2d_c2r, h->pz};
  for (cufftHandle pl : plans) if (pl) cufftDestroy(pl);
  for (int i = 0; i < 2 * PH_N; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->st_copy) { cudaStreamSynchronize(h->st_copy); cudaStreamDestroy(h->st_copy); }
  h->rb.destroy();
  if (h->st_coarse) { cudaStreamSynchronize(h->st_coarse); cudaStreamDestroy(h->st_coarse); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_vpack) cudaEventDestroy(h->ev_vpack);
  if (h->ev_vghost) cudaEventDestroy(h->ev_vghost);
  if (h->ev_xghost) cudaEventDestroy(h->ev_xghost);
  for (auto e : h->ev_up) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_copy) if (e) cudaEventDestroy(e);
  cudaStreamDestroy(h->st);
  delete h;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Every entry point that reads or writes the particle arrays: select the device and, after a cube_gpu_upload_begin, let the main
// stream wait for the last chunk (the chunks are in order on the copy stream).  cube_gpu_update_x waits chunk by chunk instead.
static int enter(cube_handle* h, bool particles = true) {
  CK(cudaSetDevice(h->p.device));
  if (particles && h->up_n) { CK(cudaStreamWaitEvent(h->st_main, h->ev_up[h->up_n - 1], 0)); h->up_n = 0; }
  return 0;
}

// the ghost positions of `hp` (buffer_x's messages may still be in flight on its side stream): every later launch on h's main
// stream sees them
static int join_xghost(cube_handle* h, cube_handle* hp = nullptr) {
  if (!hp) hp = h;
  if (hp->xghost_pending) { CK(cudaStreamWaitEvent(h->st_main, hp->ev_xghost, 0)); if (hp == h) hp->xghost_pending = false; }
  return 0;
}

static int upload_impl(cube_handle* h, const void* xp, const void* vp, const int32_t* rhoc_phys, const float* vfield_phys, int64_t nplocal,
                       int64_t npglobal, float sigma_vi, bool streamed) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  h->rb.drop();
  const Geom& g = h->g;
  if (nplocal > h->np_image_max)
    return fail("error: too many particles in this image+buffer: %lld > %lld; please set image_buffer larger", (long long)nplocal, h->np_image_max);
  if (h->copy_pending) { CK(cudaStreamWaitEvent(h->st, h->ev_copy[1], 0)); }  // a streamed download still reads the arrays overwritten here
  if (h->vghost_pending) { CK(cudaStreamWaitEvent(h->st, h->ev_vghost, 0)); h->vghost_pending = false; }
  if (join_xghost(h)) return 1;
  h->vp_stream_host = nullptr;
  h->pid_valid = false;  // a new state: its IDs, if any, come with cube_gpu_upload_pid
  // streamed: the cell arrays first, then the particles in UP_N chunks of whole key-pass CTAs on the copy stream, so that
  // cube_gpu_update_x keys chunk c while chunk c+1 is on the bus
  constexpr int UP_N = cube_handle::UP_N;
  const long long stride = g.ncell_p / UP_N / PC_CELLS * PC_CELLS;
  if (stride == 0) streamed = false;
  if (!streamed) {
    CK(cudaMemcpyAsync(h->xp, xp, (size_t)3 * h->zx * nplocal, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->vp, vp, (size_t)3 * h->zv * nplocal, cudaMemcpyHostToDevice, h->st));
  }
  CK(cudaMemcpyAsync(h->rhoc_p, rhoc_phys, sizeof(int) * g.ncell_p, cudaMemcpyHostToDevice, h->st));
  CK(cudaMemcpyAsync(h->vfield_p, vfield_phys, sizeof(float) * 3 * g.ncell_p, cudaMemcpyHostToDevice, h->st));
  if (scan_counts(h, h->rhoc_p, g.ncell_p, h->cstart_p)) return 1;
  long long tot = 0, first[UP_N + 1];
  CK(h->rb.read(&tot, h->cstart_p + g.ncell_p, sizeof(long long), h->st));
  if (streamed) CK(h->rb.read(first, h->cstart_p, sizeof(long long) * UP_N, h->st, stride));
  CK(h->rb.sync(h->st));
  if (tot != nplocal) return fail("cube_gpu_upload: sum(rhoc)=%lld differs from nplocal=%lld", tot, (long long)nplocal);
  if (streamed) {  // the main stream is idle (synchronised above) and so is the copy stream (a pending download was waited for)
    first[UP_N] = nplocal;
    for (int c = 0; c < UP_N; c++) {
      const size_t p0 = (size_t)first[c], n = (size_t)(first[c + 1] - first[c]);
      if (n) {
        CK(cudaMemcpyAsync((char*)h->xp + 3 * h->zx * p0, (const char*)xp + 3 * h->zx * p0, 3 * h->zx * n, cudaMemcpyHostToDevice, h->st_copy));
        CK(cudaMemcpyAsync((char*)h->vp + 3 * h->zv * p0, (const char*)vp + 3 * h->zv * p0, 3 * h->zv * n, cudaMemcpyHostToDevice, h->st_copy));
      }
      CK(cudaEventRecord(h->ev_up[c], h->st_copy));
    }
    h->up_n = UP_N; h->up_stride = stride;
  }
  h->nplocal = nplocal; h->npglobal = npglobal;
  h->sigma_vi = h->sigma_vi_new = sigma_vi;
  // mass_p=real((nf*nn)**3)/npglobal  (particle_initialization.f90:72)
  long long nf = (long long)g.nc * NCELL;
  h->mass_p = (float)((nf * g.nn[0]) * (nf * g.nn[1]) * (nf * g.nn[2])) / (float)npglobal;
  h->buffered = false;
  return 0;
}
extern "C" int cube_gpu_upload(cube_handle* h, const void* xp, const void* vp, const int32_t* rhoc_phys,
                               const float* vfield_phys, int64_t nplocal, int64_t npglobal, float sigma_vi) {
  return upload_impl(h, xp, vp, rhoc_phys, vfield_phys, nplocal, npglobal, sigma_vi, false);
}
// The same, returning while xp and vp are still on their way: both host arrays must be page-locked and stay untouched until the
// next call that uses the particles has returned (cube_gpu_update_x; any other entry point waits for the whole upload first).
extern "C" int cube_gpu_upload_begin(cube_handle* h, const void* xp, const void* vp, const int32_t* rhoc_phys,
                                     const float* vfield_phys, int64_t nplocal, int64_t npglobal, float sigma_vi) {
  return upload_impl(h, xp, vp, rhoc_phys, vfield_phys, nplocal, npglobal, sigma_vi, true);
}

// -DPID: the IDs of the nplocal particles of the last cube_gpu_upload, file order (particle_initialization.f90:56-59).  They ride
// through update_x (the buffers of a single image hold no ghost copies, and particle_mesh does not move particles).
extern "C" int cube_gpu_upload_pid(cube_handle* h, const int64_t* pid) {
  if (!h) return fail("null handle");
  if (!pid) return fail("cube_gpu_upload_pid: null array");
  if (enter(h)) return 1;
  if (!h->pid) { CK(dmalloc(&h->pid, h->np_image_max)); CK(dmalloc(&h->pid2, h->np_image_max)); }
  CK(cudaMemcpyAsync(h->pid, pid, sizeof(long long) * h->nplocal, cudaMemcpyHostToDevice, h->st));
  CK(h->rb.sync(h->st));
  h->pid_valid = true;
  return 0;
}
extern "C" int cube_gpu_download_pid(cube_handle* h, int64_t* pid) {
  if (!h) return fail("null handle");
  if (!h->pid_valid) return fail("no particle IDs were uploaded for this state");
  if (enter(h)) return 1;
  CK(cudaMemcpyAsync(pid, h->pid, sizeof(long long) * h->nplocal, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  return 0;
}

// Start copying xp and/or vp (whichever is not NULL) of the current disjoint state to the host on a second stream, behind
// everything already queued; returns at once.  Lets the checkpoint's device->host traffic overlap the rest of the step
// (e.g. xp right after update_x: particle_mesh does not change positions).  cube_gpu_download waits for it.
extern "C" int cube_gpu_download_async(cube_handle* h, void* xp, void* vp) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  CK(cudaEventRecord(h->ev_copy[0], h->st));
  CK(cudaStreamWaitEvent(h->st_copy, h->ev_copy[0], 0));
  if (xp) CK(cudaMemcpyAsync(xp, h->xp, (size_t)3 * h->zx * h->nplocal, cudaMemcpyDeviceToHost, h->st_copy));
  if (vp) { CK(cudaMemcpyAsync(vp, h->vp, (size_t)3 * h->zv * h->nplocal, cudaMemcpyDeviceToHost, h->st_copy)); h->copy_reads_vp = true; }
  CK(cudaEventRecord(h->ev_copy[1], h->st_copy));
  h->copy_pending = true;
  return 0;
}

extern "C" int cube_gpu_download_cells_async(cube_handle* h, int32_t* rhoc_phys, float* vfield_phys) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  const Geom& g = h->g;
  CK(cudaEventRecord(h->ev_copy[0], h->st));
  CK(cudaStreamWaitEvent(h->st_copy, h->ev_copy[0], 0));
  if (rhoc_phys) CK(cudaMemcpyAsync(rhoc_phys, h->rhoc_p, sizeof(int) * g.ncell_p, cudaMemcpyDeviceToHost, h->st_copy));
  if (vfield_phys) CK(cudaMemcpyAsync(vfield_phys, h->vfield_p, sizeof(float) * 3 * g.ncell_p, cudaMemcpyDeviceToHost, h->st_copy));
  CK(cudaEventRecord(h->ev_copy[1], h->st_copy));
  h->copy_pending = true;
  return 0;
}

extern "C" int cube_gpu_stream_vp(cube_handle* h, void* vp) {
  if (!h) return fail("null handle");
  h->vp_stream_host = vp;
  return 0;
}

extern "C" int cube_gpu_download(cube_handle* h, void* xp, void* vp, int32_t* rhoc_phys, float* vfield_phys,
                                 int64_t* nplocal, float* sigma_vi) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  const Geom& g = h->g;
  if (h->copy_pending) { CK(cudaEventSynchronize(h->ev_copy[1])); h->copy_pending = false; h->copy_reads_vp = false; }
  if (xp) CK(cudaMemcpyAsync(xp, h->xp, (size_t)3 * h->zx * h->nplocal, cudaMemcpyDeviceToHost, h->st));
  if (vp) CK(cudaMemcpyAsync(vp, h->vp, (size_t)3 * h->zv * h->nplocal, cudaMemcpyDeviceToHost, h->st));
  if (rhoc_phys) CK(cudaMemcpyAsync(rhoc_phys, h->rhoc_p, sizeof(int) * g.ncell_p, cudaMemcpyDeviceToHost, h->st));
  if (vfield_phys) CK(cudaMemcpyAsync(vfield_phys, h->vfield_p, sizeof(float) * 3 * g.ncell_p, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  if (nplocal) *nplocal = h->nplocal;
  if (sigma_vi) *sigma_vi = h->sigma_vi;
  return 0;
}

// rhoc + vfield ghost layers from the other images (buffer_density.f90:11-68), message offsets of the particle exchange
static int exchange_density(cube_handle* h, int* status) {
  const long long ng = h->ex.ng;
  const int nd = (int)h->ex.dirs.size();
  Comm* cm = h->comm.get();
  k_halo_pack<<<nblk(ng, 256), 256, 0, h->st>>>(ng, h->scell_L, h->rhoc_p, h->vfield_p, h->hsend, h->scnt); CKL();
  CC(cm->begin(h->st));
  for (const ExDir& d : h->ex.dirs) CC(cm->send(h->hsend + d.cell0, (size_t)d.ncell * sizeof(HaloRec), d.dst_rank));
  for (const ExDir& d : h->ex.dirs) CC(cm->recv(h->hrecv + d.cell0, (size_t)d.ncell * sizeof(HaloRec), d.src_rank));
  CC(cm->end());
  k_halo_unpack<<<nblk(ng, 256), 256, 0, h->st>>>(ng, h->gcell_ext, h->hrecv, h->rhoc_e, h->vfield_e, h->gcnt, h->sid_e, h->g.ncell_p); CKL();
  if (scan_counts(h, h->gcnt, ng, h->gstart)) return 1;
  if (scan_counts(h, h->scnt, ng, h->sstart)) return 1;
  k_ghost_cstart<<<nblk(ng, 256), 256, 0, h->st>>>(ng, h->gcell_ext, h->gstart, h->nplocal, h->cstart_e); CKL();
  k_dir_bounds<<<1, 64, 0, h->st>>>(nd, h->dir_cell0, h->gstart, h->sstart, h->dir_bounds); CKL();
  h->launches += 4;
  std::vector<long long> b(2 * (nd + 1));
  CK(h->rb.read(b.data(), h->dir_bounds, sizeof(long long) * b.size(), h->st));
  CK(h->rb.sync(h->st));
  for (int i = 0; i <= nd; i++) { h->gbound[i] = b[i]; h->sbound[i] = b[nd + 1 + i]; }
  h->nghost = h->gbound[nd];
  if (h->sbound[nd] > h->sendcap) {  // the message buffer follows the state (the reference's only limit is np_image_max, below)
    CK(h->rb.sync(h->st));
    CK(cudaFree(h->psend)); h->psend = nullptr;
    h->sendcap = h->sbound[nd] + h->sbound[nd] / 4 + 4096;
    CK(cudaMalloc(&h->psend, (size_t)3 * std::max(h->zx, h->zv) * h->sendcap + 16));
  }
  if (h->nplocal + h->nghost > h->np_image_max) *status = 1;
  return 0;
}

// ghost particles (buffer_x.f90 for xp, buffer_v.f90 for vp): received behind the physical particles
// `cst`: the stream the messages travel on (the pack kernel always runs on h->st, the messages wait for it)
static int exchange_particles(cube_handle* h, void* arr_v, int z /* bytes per code */, cudaStream_t cst) {
  char* arr = (char*)arr_v; char* psend = (char*)h->psend;
  const long long ng = h->ex.ng;
  const int nd = (int)h->ex.dirs.size();
  Comm* cm = h->comm.get();
  if (z == 2) k_particle_pack<short><<<nblk(ng, PC_CELLS), PC_T, 0, h->st>>>(ng, h->scell_L, h->sstart, h->cstart_p, (const short*)arr, (short*)psend);
  else k_particle_pack<signed char><<<nblk(ng, PC_CELLS), PC_T, 0, h->st>>>(ng, h->scell_L, h->sstart, h->cstart_p, (const signed char*)arr, (signed char*)psend);
  CKL();
  h->launches++;
  if (cst != h->st) { CK(cudaEventRecord(h->ev_vpack, h->st)); CK(cudaStreamWaitEvent(cst, h->ev_vpack, 0)); }
  CC(cm->begin(cst));
  for (int i = 0; i < nd; i++) {
    const size_t n = (size_t)(h->sbound[i + 1] - h->sbound[i]);
    if (n) CC(cm->send(psend + (size_t)3 * z * h->sbound[i], n * 3 * z, h->ex.dirs[i].dst_rank));
  }
  for (int i = 0; i < nd; i++) {
    const size_t n = (size_t)(h->gbound[i + 1] - h->gbound[i]);
    if (n) CC(cm->recv(arr + (size_t)3 * z * (h->nplocal + h->gbound[i]), n * 3 * z, h->ex.dirs[i].src_rank));
  }
  CC(cm->end());
  return 0;
}

// -DPID: the IDs of the ghost particles (buffer_v.f90:23,42,62,81,104 moves pid with vp), received behind the physical ones
static int exchange_pid(cube_handle* h, cudaStream_t cst) {
  const long long ng = h->ex.ng;
  const int nd = (int)h->ex.dirs.size();
  Comm* cm = h->comm.get();
  if (h->sbound[nd] > h->pid_sendcap) {
    if (h->pid_send) { CK(h->rb.sync(h->st)); CK(cudaFree(h->pid_send)); h->pid_send = nullptr; }
    h->pid_sendcap = h->sbound[nd] + h->sbound[nd] / 4 + 4096;
    CK(dmalloc(&h->pid_send, h->pid_sendcap));
  }
  k_pid_pack<<<nblk(ng, PC_CELLS), PC_T, 0, h->st>>>(ng, h->scell_L, h->sstart, h->cstart_p, h->pid, h->pid_send); CKL();
  h->launches++;
  if (cst != h->st) { CK(cudaEventRecord(h->ev_vpack, h->st)); CK(cudaStreamWaitEvent(cst, h->ev_vpack, 0)); }
  CC(cm->begin(cst));
  for (int i = 0; i < nd; i++) {
    const size_t n = (size_t)(h->sbound[i + 1] - h->sbound[i]);
    if (n) CC(cm->send(h->pid_send + h->sbound[i], n * sizeof(long long), h->ex.dirs[i].dst_rank));
  }
  for (int i = 0; i < nd; i++) {
    const size_t n = (size_t)(h->gbound[i + 1] - h->gbound[i]);
    if (n) CC(cm->recv(h->pid + h->nplocal + h->gbound[i], n * sizeof(long long), h->ex.dirs[i].src_rank));
  }
  CC(cm->end());
  return 0;
}

extern "C" int cube_gpu_buffer(cube_handle* h, int do_density, int do_x, int do_v, float* overhead_image) {
  if (!h) return fail("null handle");
  g_cur = h; h->rb.drop();
  if (enter(h, h->nimg > 1 && (do_x || do_v))) return 1;  // buffer_density reads the cell arrays only; one image's ghosts are aliases
  const Geom& g = h->g;
  const bool multi = h->nimg > 1;
  if (do_density) {
    PhaseTimer pt(h, PH_BUFFER);
    int status = 0;
    k_build_ext<<<nblk(g.ncell_e, 256), 256, 0, h->st>>>(g, h->rhoc_p, h->cstart_p, h->vfield_p, h->rhoc_e, h->cstart_e, h->vfield_e, h->sid_e); CKL();
    if (multi && exchange_density(h, &status)) return 1;
    CK(cudaMemsetAsync(h->tile_count, 0, sizeof(long long) * g.nnt * g.nnt * g.nnt, h->st));
    k_tile_counts<<<dim3(32, g.nnt * g.nnt * g.nnt), 256, 0, h->st>>>(g, h->rhoc_e, (unsigned long long*)h->tile_count); CKL();
    h->launches += 2;
    // overhead_image=sum(rhoc)/np_image_max over the buffered rhoc of all tiles (buffer_density.f90:75)
    const int ntile = g.nnt * g.nnt * g.nnt;
    std::vector<long long> tc(ntile);
    if (multi || overhead_image) {
      CK(h->rb.read(tc.data(), h->tile_count, sizeof(long long) * ntile, h->st));
      CK(h->rb.sync(h->st));
    }
    long long s = 0; for (long long v : tc) s += v;
    float ovh = (float)((double)s / (double)h->np_image_max);
    int bad_image = (double)ovh > 1.0 || status ? h->p.rank + 1 : 0;
    long long bad_count = s;
    if (multi) {  // overhead_image is the maximum over the images (buffer_density.f90:78-86); every image stops together
      struct Rec { float ovh; int bad; long long s; } mine = {ovh, bad_image, s};
      std::vector<Rec> all(h->nimg);
      CC(h->comm->allgather_host(&mine, all.data(), sizeof(Rec), h->st));
      bad_image = 0;
      for (int m = 0; m < h->nimg; m++) {
        ovh = std::max(ovh, all[m].ovh);
        if (all[m].bad && !bad_image) { bad_image = all[m].bad; bad_count = all[m].s; }
      }
    }
    if (overhead_image) *overhead_image = ovh;
    if (bad_image)
      return fail("error: too many particles in this image+buffer: %lld > %lld on image %d; please set image_buffer larger", bad_count, h->np_image_max, bad_image);
    h->buffered = true;
  }
  // one image: ghost particles alias the periodic image, nothing to copy
  if (do_x && multi) {
    // The ghosts' positions are first read by the fine deposit's bricks next to the image boundary and by the coarse deposit: the
    // messages travel on the side stream while particle_mesh deposits the bricks that read no ghost cell (fine_deposit below).
    PhaseTimer pt(h, PH_BUFFER);
    if (join_xghost(h)) return 1;
    if (h->vghost_pending) { CK(cudaStreamWaitEvent(h->st, h->ev_vghost, 0)); }  // buffer_v's messages leave from the same send buffer
    cudaStream_t cst = (h->async_xghost && !h->prof) ? h->st_coarse : h->st;
    if (exchange_particles(h, h->xp, h->zx, cst)) return 1;
    if (cst != h->st) { CK(cudaEventRecord(h->ev_xghost, cst)); h->xghost_pending = true; }
  }
  if (do_v && multi) {
    // The ghosts' velocities (and IDs) are first read by the next update_particle's pass over the GHOST cells: the messages travel on
    // the high-priority side stream and that pass waits for them, so they cross under the key pass of the physical cells (which
    // only reads physical particles).  Phase profiling keeps everything on one stream.
    PhaseTimer pt(h, PH_BUFFER);
    if (join_xghost(h)) return 1;  // buffer_x's messages leave from the same send buffer
    cudaStream_t cst = (h->async_vghost && !h->prof) ? h->st_coarse : h->st;
    if (exchange_particles(h, h->vp, h->zv, cst)) return 1;
    if (h->pid_valid && exchange_pid(h, cst)) return 1;
    if (cst != h->st) { CK(cudaEventRecord(h->ev_vghost, cst)); h->vghost_pending = true; }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" int cube_gpu_update_x(cube_handle* h, float dt_old, float dt, int64_t* nplocal, float* sigma_vi_new,
                                 double std_vsim[3], float* overhead_tile) {
  if (!h) return fail("null handle");
  g_cur = h; h->rb.drop();
  if (enter(h, false)) return 1;  // a streamed upload is waited for chunk by chunk, below
  if (!h->buffered) return fail("cube_gpu_update_x: state is not buffered (call cube_gpu_buffer first, cafcube.f90:17-19)");
  if (h->copy_pending) { CK(cudaStreamWaitEvent(h->st, h->ev_copy[1], 0)); }  // an asynchronous download still reads the particle arrays
  const Geom& g = h->g;
  const bool multi = h->nimg > 1;
  const long long ng = h->ex.ng;
  const float dt_mid_f = (dt_old + dt) / 2;  // update_particle.f90:16
  const double dt_mid = (double)dt_mid_f, S = vscale(h->sigma_vi);
  if (build_dvlut(h, h->sigma_vi)) return 1;
  // tile overflow check (update_particle.f90:60-67): particles in each tile's extended region
  const int ntile = g.nnt * g.nnt * g.nnt;
  std::vector<long long> tc(ntile);
  CK(h->rb.read(tc.data(), h->tile_count, sizeof(long long) * ntile, h->st));
  CK(cudaMemsetAsync(h->maxoff, 0, sizeof(int), h->st));
  int maxoff = 0;
  const unsigned npw = pw_grid(h, g.ncell_p), nchunk_g = nblk(ng, PC_CELLS);
  {
    PhaseTimer pt(h, PH_KEY);
    CK(cudaMemsetAsync(h->inflag, 0, (size_t)g.ncell_p, h->st));
    CK(cudaMemsetAsync(h->nflag, 0, sizeof(int), h->st));
    const int nup = std::max(1, h->up_n);
    for (int c = 0; c < nup; c++) {  // one launch, or one per chunk of a streamed upload as it lands (cube_gpu_upload_begin)
      const long long cb = h->up_n ? c * h->up_stride : 0, ce = (h->up_n && c + 1 < nup) ? cb + h->up_stride : g.ncell_p;
      if (h->up_n) CK(cudaStreamWaitEvent(h->st, h->ev_up[c], 0));
      FMT_SWITCH(h, k_drift_key_chain<F><<<nblk(ce - cb, PC_CELLS), PC_T, KC_SMEM, h->st>>>(g, XPC(h->xp), VPC(h->vp), h->cstart_p, h->vfield_p, h->dvlut, dt_mid, h->key,
                                                                                       h->rank, h->maxoff, h->mask_s, h->inflag, h->rhoc_p2, h->vfield_p2, cb));
      CKL();
      h->launches++;
    }
    h->up_n = 0;
    if (h->vghost_pending) { CK(cudaStreamWaitEvent(h->st, h->ev_vghost, 0)); h->vghost_pending = false; }  // ghost velocities from buffer_v
    if (join_xghost(h)) return 1;
    if (multi && ng) {
      FMT_SWITCH(h, k_drift_key_g<F><<<nchunk_g, PC_T, 0, h->st>>>(g, ng, h->gcell_ext, h->gstart, h->nplocal, XPC(h->xp), VPC(h->vp), h->vfield_e, h->dvlut, dt_mid,
                                                                   h->key, h->rank, h->maxoff, h->mask_s + MASK_W * g.ncell_p, h->inflag));
      CKL();
      h->launches++;
    }
    CK(cudaMemsetAsync(h->farblk, 0, sizeof(int) * (size_t)farblk_dim(g) * farblk_dim(g) * farblk_dim(g), h->st));
    k_mask_ext<<<nblk(g.ncell_e, 256), 256, 0, h->st>>>(g, h->sid_e, h->mask_s, h->rhoc_e, std::max(1, h->heavy_count / 4), h->mask_e, h->farblk); CKL();
    h->launches++;
    CK(h->rb.read(&maxoff, h->maxoff, sizeof(int), h->st));
    CK(h->rb.sync(h->st));
  }
  // local failures are agreed on by all images before anyone returns (the reference `stop`s every image)
  int status = 0; std::string msg;
  float ovh = 0;
  for (int t = 0; t < ntile; t++) {
    ovh = std::max(ovh, (float)tc[t] / (float)h->np_tile_max);
    if (tc[t] > h->np_tile_max && !status) {
      status = 1;
      char buf[512];
      snprintf(buf, sizeof buf, "error: too many particles in this tile+buffer: %lld > %lld on image %d tile %d %d %d; please set tile_buffer larger",
               tc[t], h->np_tile_max, h->p.rank + 1, t % g.nnt + 1, (t / g.nnt) % g.nnt + 1, t / (g.nnt * g.nnt) + 1);
      msg = buf;
    }
  }
  if (maxoff > NCB && !status) {
    status = 2;
    char buf[256];
    snprintf(buf, sizeof buf, "cube_gpu_update_x: a particle of image %d moves %d coarse cells in one step (> ncb=%d): outside the tile buffer", h->p.rank + 1, maxoff, NCB);
    msg = buf;
  }
  long long tot = 0;
  double st[3] = {0, 0, 0};
  if (!status) {
    const int r = maxoff;
    h->last_radius = r;
    const unsigned nb = nblk(g.ncell_p, 128);
    {
      PhaseTimer pt(h, PH_COUNT);
      // pass B only redoes the cells that receive somebody; the others were finished by k_drift_key_chain
      const int* flist = h->count_all ? nullptr : h->flist;
      if (flist) { k_flag_compact<<<nblk(g.ncell_p, 256), 256, 0, h->st>>>(g.ncell_p, h->inflag, h->flist, h->nflag); CKL(); h->launches++; }
      FMT_SWITCH(h, const DriftCountArgs<F> A{XPC(h->xp), VPC(h->vp), h->key, h->rhoc_e, h->cstart_e, h->vfield_e, h->dvlut, h->mask_e, h->farblk, h->rank, dt_mid, r, h->nlayer};
                 if (h->count_minb == 8) k_drift_count<8, F><<<nb, DC_T, 0, h->st>>>(g, A, h->heavy_count, flist, h->nflag, h->rhoc_p2, h->vfield_p2);
                 else k_drift_count<5, F><<<nb, DC_T, 0, h->st>>>(g, A, h->heavy_count, flist, h->nflag, h->rhoc_p2, h->vfield_p2));
      CKL();
      k_vfield_sq<<<nb, 128, 0, h->st>>>(g.ncell_p, h->vfield_p2, h->stat_partial); CKL();
      k_reduce_strided<<<1, 1024, 0, h->st>>>(h->stat_partial, nb, 1, 0, h->stat3 + 1); CKL();
      h->launches += 3;
    }
    {
      PhaseTimer pt(h, PH_SCAN);
      if (scan_counts(h, h->rhoc_p2, g.ncell_p, h->cstart_p2)) return 1;
    }
    CK(h->rb.read(&tot, h->cstart_p2 + g.ncell_p, sizeof(long long), h->st));
    CK(h->rb.sync(h->st));
    if (tot > h->np_image_max) {
      status = 3;
      char buf[256];
      snprintf(buf, sizeof buf, "error: too many particles in this image+buffer: %lld > %lld on image %d; please set image_buffer larger", tot, h->np_image_max, h->p.rank + 1);
      msg = buf;
    }
  }
  if (!status) {
    PhaseTimer pt(h, PH_PLACE);
    FMT_SWITCH(h, k_drift_place_w<F><<<npw, PW_T, pw_smem_bytes(h->vt_hot), h->st>>>(g, vtab(h), S, XPC(h->xp), VPC(h->vp), h->rank, h->cstart_p, h->vfield_p, h->cstart_p2,
                                                                                     h->vfield_p2, dt_mid, XPM(h->xp2), VPM(h->vp2), h->stat_partial));
    CKL();
    k_reduce_strided<<<1, 1024, 0, h->st>>>(h->stat_partial, (long long)npw * PW_W, 2, 0, h->stat3); CKL();
    k_reduce_strided<<<1, 1024, 0, h->st>>>(h->stat_partial, (long long)npw * PW_W, 2, 1, h->stat3 + 2); CKL();
    h->launches += 3;
    if (h->pid_valid) {  // update_particle.f90:88,106: the IDs take the same permutation
      k_pid_place<<<nblk(g.ncell_p, 256), 256, 0, h->st>>>(g, h->pid, h->rank, h->cstart_p, h->cstart_p2, h->pid2); CKL();
      h->launches += 1;
      if (multi && ng) {
        k_pid_place_g<<<nblk(ng, 256), 256, 0, h->st>>>(g, ng, h->gcell_ext, h->gstart, h->nplocal, h->pid, h->rank, h->cstart_p2, h->pid2); CKL();
        h->launches += 1;
      }
    }
    double stg[2] = {0, 0};
    if (multi && ng) {
      FMT_SWITCH(h, k_drift_place_g<F><<<nchunk_g, PC_T, 0, h->st>>>(g, ng, h->gcell_ext, h->gstart, h->nplocal, XPC(h->xp), VPC(h->vp), h->rank, h->vfield_e, h->cstart_p2,
                                                                     h->vfield_p2, h->dvlut, h->enc, dt_mid, S, XPM(h->xp2), VPM(h->vp2), h->stat_partial_g));
      CKL();
      k_reduce_strided<<<1, 1024, 0, h->st>>>(h->stat_partial_g, (long long)nchunk_g, 2, 0, h->stat3 + 3); CKL();
      k_reduce_strided<<<1, 1024, 0, h->st>>>(h->stat_partial_g, (long long)nchunk_g, 2, 1, h->stat3 + 4); CKL();
      h->launches += 3;
      CK(h->rb.read(stg, h->stat3 + 3, sizeof stg, h->st));
    }
    CK(h->rb.read(st, h->stat3, sizeof st, h->st));
    CK(h->rb.sync(h->st));
    st[0] += stg[0]; st[2] += stg[1];
  }
  long long npsum = tot;
  if (multi) {  // co_sum / co_max of update_particle.f90:154-166,186-193 in image order
    struct Rec { double st[3]; long long np; float ovh; int status; } mine = {{st[0], st[1], st[2]}, tot, ovh, status};
    std::vector<Rec> all(h->nimg);
    CC(h->comm->allgather_host(&mine, all.data(), sizeof(Rec), h->st));
    int bad = -1;
    st[0] = all[0].st[0]; st[1] = all[0].st[1]; st[2] = all[0].st[2]; npsum = all[0].np; ovh = all[0].ovh;
    if (all[0].status) bad = 0;
    for (int m = 1; m < h->nimg; m++) {
      st[0] += all[m].st[0]; st[1] += all[m].st[1]; st[2] += all[m].st[2]; npsum += all[m].np; ovh = std::max(ovh, all[m].ovh);
      if (all[m].status && bad < 0) bad = m;
    }
    if (bad >= 0 && !status) return fail("update_particle stopped: image %d reported an error (status %d)", bad + 1, all[bad].status);
  }
  if (status) return fail("%s", msg.c_str());
  if (multi && npsum != h->npglobal)  // update_particle.f90:205-211
    return fail("np check failed: sum(nplocal)=%lld differs from npglobal=%lld", npsum, h->npglobal);
  std::swap(h->xp, h->xp2); std::swap(h->vp, h->vp2);
  if (h->pid_valid) std::swap(h->pid, h->pid2);
  std::swap(h->rhoc_p, h->rhoc_p2); std::swap(h->vfield_p, h->vfield_p2); std::swap(h->cstart_p, h->cstart_p2);
  h->nplocal = tot;
  h->buffered = false;
  // update_particle.f90:170-175
  const double nglob = (double)h->npglobal;
  const double sv = std::sqrt(st[0] / nglob);
  const double svc = std::sqrt(st[1] / (double)g.nc / (double)g.nc / (double)g.nc / (double)g.nn[0] / (double)g.nn[1] / (double)g.nn[2]);
  const double svr = std::sqrt(st[2] / nglob);
  h->sigma_vi_new = (float)(svr / (double)sqrtf(3.f));
  if (nplocal) *nplocal = tot;
  if (sigma_vi_new) *sigma_vi_new = h->sigma_vi_new;
  if (std_vsim) { std_vsim[0] = sv; std_vsim[1] = svc; std_vsim[2] = svr; }
  if (overhead_tile) *overhead_tile = ovh;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// fine mesh of tiles [tile0, tile0+nb): deposit -> x,y forward -> z forward * i kern_f, z inverse (x3) -> y,x inverse
// (pm.f90:44-84).  Leaves force_f of the nb tiles in h->F.
// deposit the particles of source cells R.c0..R.c1 onto the fine-grid region R (frame: FRAME_NONE, or the tile's first cell)
// `hp`: whose particles (a further species of a two-species run deposits into the first species' meshes, accumulate = 1)
static int fine_deposit(cube_handle* h, const FineRegion& R, int3 frame, float* out, cube_handle* hp = nullptr, int accumulate = 0) {
  if (!hp) hp = h;
  PhaseTimer pt(h, PH_FDEP);
  // ghost positions still in flight (buffer_x on the side stream): first the bricks that read no ghost cell, then, behind the
  // messages, the others.  Anything else waits for the messages first.
  const bool two_phase = hp->xghost_pending && h->st == h->st_main && frame.x == FRAME_NONE && !accumulate;
  if (!two_phase && hp->xghost_pending && h->st == h->st_main && join_xghost(h, hp)) return 1;
  int part = two_phase ? 1 : 0;
  auto launch = [&](auto cfg) {
    using C = decltype(cfg);
    const unsigned nbx = (R.n[0] + C::NX - 1) / C::NX, nby = (R.n[1] + C::NY - 1) / C::NY, nbz = (R.n[2] + C::NZ - 1) / C::NZ;
    auto go = [&](auto xt) {
      using XT = decltype(xt);
#define FD_GO(FR, AC) k_fine_deposit_r<C, FR, XT, false, AC><<<nbx * nby * nbz, C::NT, C::SMEM, h->st>>>(h->g, R, frame, (const XT*)hp->xp, hp->rhoc_e, hp->cstart_e, hp->mass_p, out, part)
      if (accumulate) {
        if (cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, false, XT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, true, XT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) return;
        if (frame.x == FRAME_NONE) FD_GO(false, true); else FD_GO(true, true);
      } else if (frame.x == FRAME_NONE) FD_GO(false, false); else FD_GO(true, false);
#undef FD_GO
    };
    if (hp->zx == 2) go((short)0); else go((signed char)0);
  };
  if (h->fd_brick == 888) launch(Fd888{}); else if (h->fd_brick == 844) launch(Fd844{}); else launch(Fd884{});
  CKL();
  h->launches++;
  if (two_phase) {
    if (join_xghost(h, hp)) return 1;
    part = 2;
    if (h->fd_brick == 888) launch(Fd888{}); else if (h->fd_brick == 844) launch(Fd844{}); else launch(Fd884{});
    CKL();
    h->launches++;
  }
  return 0;
}
// fine density of the tiles [tile0, tile0+nb) in h->rho; returns how the x pass finds each tile's window
static int fine_density_batch(cube_handle* h, int tile0, int nb, RhoView& v, cube_handle* h2 = nullptr) {
  const Geom& g = h->g;
  const int N = h->fg.N;
  v.p = h->rho; v.nnt = g.nnt; v.tile0 = tile0; v.tstep = 4 * g.nt;
  if (h->shared_region) {
    const FineRegion R = batch_region(h, tile0, nb);
    int k[3];
    batch_box(g, tile0, nb, v.t0, k);
    v.ldy = R.ldy; v.ldz = R.ldz; v.tvol = 0;
    if (fine_deposit(h, R, make_int3(FRAME_NONE, 0, 0), h->rho)) return 1;
    return h2 ? fine_deposit(h, R, make_int3(FRAME_NONE, 0, 0), h->rho, h2, 1) : 0;  // pm.f90:79-99 (NEUTRINOS): the same rho_f
  }
  v.ldy = N; v.ldz = (long long)N * N; v.tvol = (long long)N * N * N; v.t0[0] = v.t0[1] = v.t0[2] = 0;
  for (int b = 0; b < nb; b++) {  // large tiles: each in its own frame (f32 rounding of tempx, cube_kernels.cuh)
    const int t = tile0 + b, tc[3] = {t % g.nnt, (t / g.nnt) % g.nnt, t / (g.nnt * g.nnt)};
    FineRegion R;
    for (int d = 0; d < 3; d++) { R.c0[d] = tc[d] * g.nt - (NCB - 1); R.c1[d] = (tc[d] + 1) * g.nt + (NCB - 1); R.f0[d] = 4 * tc[d] * g.nt - 16; R.n[d] = N; }
    R.ldy = N; R.ldz = (long long)N * N;
    if (fine_deposit(h, R, make_int3(tc[0] * g.nt, tc[1] * g.nt, tc[2] * g.nt), h->rho + (size_t)b * v.tvol)) return 1;
    if (h2 && fine_deposit(h, R, make_int3(tc[0] * g.nt, tc[1] * g.nt, tc[2] * g.nt), h->rho + (size_t)b * v.tvol, h2, 1)) return 1;
  }
  return 0;
}

// leaves force_f of the nb tiles in h->F (multiplied by the kick prefix a_mid*dt/6/pi when `prefix`) and the per-tile
// f2_max_fine in h->f2max[0..nb)
// `pre`: the density is already there (a region deposited for a larger group of tiles that contains these)
static int fine_mesh(cube_handle* h, int tile0, int nb, bool prefix, float a_mid, float dt, cube_handle* h2 = nullptr, const RhoView* pre = nullptr) {
  FftGeom f = h->fg;
  f.nbatch = nb;
  const FftPlan& pl = *h->plan;
  const int N = f.N, T = pl.threads();
  const size_t smem_x = (size_t)(N * (FL + 1) + N) * sizeof(float2), smem_y = (size_t)(N * FL + N) * sizeof(float2);
  const size_t smem_z = (size_t)(2 * N * FL + N) * sizeof(float2) + (size_t)3 * (N / 2 + 1) * FL * sizeof(float);
  RhoView rv;
  if (pre) { rv = *pre; rv.tile0 = tile0; }
  else if (fine_density_batch(h, tile0, nb, rv, h2)) return 1;
  {
    PhaseTimer pt(h, PH_FFTX);
    pl.x_fwd<<<dim3((N + 31) / 32, N, nb), T, smem_x, h->st>>>(f, rv, h->Ak, h->tw); CKL();
  }
  {
    PhaseTimer pt(h, PH_FFTY);
    pl.y_fwd<<<dim3(f.P / FL, N, nb), T, smem_y, h->st>>>(f, h->Ak, h->tw); CKL();
  }
  {
    PhaseTimer pt(h, PH_FFTZ);
    // 1/N^3 (pm.f90:82) and, with `prefix`, the kick's per-node factor a_mid*dt/6/pi (pm.f90:104) ride on the Green multiply
    float scale = 1.0f / ((float)N * (float)N * (float)N);
    if (prefix) scale *= ((1.0f * a_mid) * dt) / 6.0f / PI_F;
    const char* zg = getenv("CUBE_GPU_ZG");
    const bool sb = zg && !strcmp(zg, "sb");
    auto zk = zg && !strcmp(zg, "pk") ? pl.z_green_pk : sb ? pl.z_green_mx : pl.z_green;
    zk<<<dim3(f.P / FL, N), T, sb ? smem_z - (size_t)N * FL * sizeof(float2) : smem_z, h->st>>>(f, h->Ak, h->Bk, h->kern_f, scale, h->tw); CKL();
  }
  {
    PhaseTimer pt(h, PH_IFFTY);
    pl.y_inv<<<dim3(f.P / FL, f.M, 3 * nb), T, smem_y, h->st>>>(f, h->Bk, h->tw); CKL();
  }
  {
    PhaseTimer pt(h, PH_IFFTX);  // x inverse of the three components + f2_max_fine
    CK(cudaMemsetAsync(h->f2max, 0, sizeof(unsigned) * nb, h->st));
    pl.x_inv<<<dim3((f.M + 15) / 16, f.M, nb), pl.x_inv_threads, pl.x_inv_smem, h->st>>>(f, h->Bk, h->F, h->tw, h->f2max); CKL();
  }
  h->launches += 5;
  return 0;
}

// the stream the coarse-mesh cuFFT plans execute on
static int set_coarse_streams(cube_handle* h, cudaStream_t st) {
  if (h->nimg > 1) { CF(cufftSetStream(h->p2d_r2c, st)); CF(cufftSetStream(h->p2d_c2r, st)); CF(cufftSetStream(h->pz, st)); }
  else { CF(cufftSetStream(h->cplan_r2c, st)); CF(cufftSetStream(h->cplan_c2r, st)); }
  return 0;
}

// coarse mesh (pm.f90:127-189): deposit -> r2c -> i kern_c -> 3 c2r -> force_c with halo.  Leaves the kick prefix
// force_c*a_mid*dt/6/pi in h->fc, f2_max_coarse in h->f2max[batch]; raw (optional, device) gets force_c itself.
static int coarse_mesh(cube_handle* h, bool through_force, float a_mid, float dt, float* raw, cube_handle* h2 = nullptr) {
  const Geom& g = h->g;
  const bool multi = h->nimg > 1;
  // the coarse deposit reads the first layer of ghost cells: on the side stream it is queued behind buffer_x's messages already
  for (cube_handle* hp : {h, h2})
    if (hp && hp->xghost_pending && (h->st == h->st_main || hp != h)) CK(cudaStreamWaitEvent(h->st, hp->ev_xghost, 0));
  {
    PhaseTimer pt(h, PH_CDEP);
    const long long nbox = (long long)(g.nc + 2) * (g.nc + 2) * (g.nc + 2);
    for (cube_handle* hp : {h, h2}) {  // every species into the same r3 (pm.f90:130-163 with NEUTRINOS)
      if (!hp) continue;
      const int heavy = hp->mass_p <= 8.f ? h->heavy_deposit : INT_MAX;  // REDUX sums of 32 terms stay below 2^32
      if (hp->zx == 2) k_coarse_cell_sums<short><<<nblk(nbox, CD_T), CD_T, 0, h->st>>>(g, heavy, (const short*)hp->xp, hp->rhoc_e, hp->cstart_e, hp->mass_p, nbox, h->csum);
      else k_coarse_cell_sums<signed char><<<nblk(nbox, CD_T), CD_T, 0, h->st>>>(g, heavy, (const signed char*)hp->xp, hp->rhoc_e, hp->cstart_e, hp->mass_p, nbox, h->csum);
      CKL();
      k_coarse_gather27<<<nblk(g.ncell_p, 256), 256, 0, h->st>>>(g, nbox, h->csum, h->r3, multi ? g.nc : g.nc + 2, hp != h); CKL();
      h->launches += 2;
    }
  }
  if (!through_force) return 0;
  PhaseTimer pt(h, PH_CFFT);
  CK(cudaMemsetAsync(h->f2max + h->batch, 0, sizeof(unsigned), h->st));
  if (multi) {  // distributed transform (cube_coarse.cuh) replacing pencil_fft.f90 + the halo GETs of pm.f90:182-189
    if (coarse_forward(h, h->r3)) return 1;
    if (coarse_backward3(h)) return 1;
    k_force_c_finish_dist<<<1184, 256, 0, h->st>>>(g.nc, h->recvF, h->zzoff, h->zzcs, a_mid, dt, h->fc, raw, h->f2max + h->batch); CKL();
    h->launches++;
    return 0;
  }
  CF(cufftExecR2C(h->cplan_r2c, h->r3, (cufftComplex*)h->r3));
  // rxyz = i*kern_c*crho_c ; r3=r3/ng_global^3 (pm.f90:172-175, pencil_fft.f90:58)
  const float scale = 1.0f / ((float)g.nc * g.nn[0]) / ((float)g.nc * g.nn[1]) / ((float)g.nc * g.nn[2]);
  k_green<<<nblk(h->cnk, 256), 256, 0, h->st>>>(h->cnk, 1, (const float2*)h->r3, h->kern_c, scale, (float2*)h->cforce); CKL();
  CF(cufftExecC2R(h->cplan_c2r, (cufftComplex*)h->cforce, h->cforce));
  k_force_c_finish<<<1184, 256, 0, h->st>>>(g, h->cforce, a_mid, dt, h->fc, raw, h->f2max + h->batch); CKL();
  h->launches += 2;
  return 0;
}

extern "C" int cube_gpu_particle_mesh(cube_handle* h, float a_mid, float dt, float* dt_fine, float* dt_coarse,
                                      float* dt_vmax, float* vmax_out) {
  if (!h) return fail("null handle");
  g_cur = h; h->rb.drop();
  if (enter(h)) return 1;
  if (!h->buffered) return fail("cube_gpu_particle_mesh: state is not buffered (call cube_gpu_buffer first)");
  // a cube_gpu_download_async of vp still reads the velocities the kicks rewrite in place (positions may keep streaming)
  if (h->copy_pending && h->copy_reads_vp) { CK(cudaStreamWaitEvent(h->st, h->ev_copy[1], 0)); h->copy_reads_vp = false; }
  const Geom& g = h->g;
  const int ntile = g.nnt * g.nnt * g.nnt;
  const long long nt3 = (long long)g.nt * g.nt * g.nt;
  if (build_dvlut(h, h->sigma_vi)) return 1;
  const double S_new = vscale(h->sigma_vi_new);
  std::vector<float> f2(ntile, 0.f);
  // The x inverse stores force_f*a_mid*dt/6/pi (the kick's per-node prefix, pm.f90:104); f2_max_fine is then taken over
  // the prefixed mesh and scaled back by the square of the same f32 factor: equal to maxval(sum(force_f**2,1)) (pm.f90:85)
  // to 3e-7 relative, below the difference between any two FFT implementations.  A degenerate factor (dt = 0) takes
  // the unfused route: raw forces, exact f2_max, separate prefix pass.
  const float pscale = ((1.0f * a_mid) * dt) / 6.0f / PI_F;
  const bool pre_in_fft = pscale > 1e-12f && pscale < 1e12f;
  // The coarse mesh only needs the positions: it runs on a second stream under the fine mesh (its small FFTs and, with several
  // images, its all-to-all exchanges then cost nothing).  Phase profiling keeps everything on one stream, in the reference's order.
  // The coarse stream has the highest priority: with several images the exchange kernels of the distributed coarse FFT must
  // get the next CTA slots that free up, or every rank waits for the slowest peer's exchange kernel to be scheduled behind a
  // GPU full of fine-mesh CTAs (measured at default priority: 65.1 ms/step overlapped against 56.8 in sequence on 8 B200s,
  // 63.4 against 52.8 on 2; at the highest priority 51.8 against 53.5 on 2).
  const bool overlap = h->overlap_coarse && !h->prof;
  // streamed velocities (cube_gpu_stream_vp): smaller batches, coarse kick per batch, the batch's vp out under the next batch
  void* const vp_host = h->vp_stream_host;
  h->vp_stream_host = nullptr;
  // streamed velocities: batches of a quarter of the tiles, then 1/8, then 1/16 + 1/16 -- a batch's velocities leave under the next
  // batch, the last batch's are the exposed tail, and every launch of the z pass reloads its Green table (0.75 tiles' worth)
  const int step_batch = vp_host ? align_batch(g.nnt, std::max(1, std::min(h->batch, (ntile + 3) / 4))) : h->batch;
  const int step_min = vp_host ? align_batch(g.nnt, std::max(1, std::min(step_batch, (ntile + 15) / 16))) : h->batch;
  std::vector<long long> tile_start(ntile + 1, 0);
  const bool merged = !h->old_kick;  // one pass per batch does both kicks (cube_kick.cuh); else fine kick per batch, coarse kick at the end
  const bool kick_c_per_batch = merged || vp_host;
  if (build_dvlut2(h, h->sigma_vi_new)) return 1;
  CK(cudaMemsetAsync(h->vmax_bits, 0, 4 * sizeof(unsigned long long), h->st));
  if (vp_host) {
    CK(h->rb.read(tile_start.data(), h->cstart_p, sizeof(long long) * (ntile + 1), h->st, nt3));
    CK(h->rb.sync(h->st));
  }
  if (!overlap && kick_c_per_batch && coarse_mesh(h, true, a_mid, dt, nullptr)) return 1;
  if (overlap) {
    cudaStream_t main_st = h->st;
    CK(cudaEventRecord(h->ev_fork, main_st));
    CK(cudaStreamWaitEvent(h->st_coarse, h->ev_fork, 0));
    h->st = h->st_coarse;
    int rc = set_coarse_streams(h, h->st_coarse);
    if (!rc) rc = coarse_mesh(h, true, a_mid, dt, nullptr);
    h->st = main_st;
    if (set_coarse_streams(h, main_st) || rc) return 1;
    CK(cudaEventRecord(h->ev_join, h->st_coarse));
  }
  // Streaming in small batches would deposit each batch's own region (more overlap between regions: 1.75x of the particles for a
  // layer of tiles against 1.19x for the image).  Instead the density of a whole group of h->batch tiles is deposited once, behind the
  // part of Bk the small batches use, and each small batch transforms its windows of it.
  struct RhoSwap { cube_handle* h; float* keep; ~RhoSwap() { h->rho = keep; } } swap_back{h, h->rho};
  const bool split = vp_host && h->shared_region && step_batch < h->batch && !h->rho_own &&
                     h->rho_n * sizeof(float) + h->B_n * step_batch * sizeof(float2) <= h->B_n * h->batch * sizeof(float2);
  if (split) h->rho = reinterpret_cast<float*>(h->Bk + h->B_n * step_batch);
  RhoView group;
  for (int t0 = 0, nb = 0; t0 < ntile; t0 += nb) {
    nb = std::min(step_batch, ntile - t0);
    if (vp_host && ntile - t0 <= step_batch && nb > step_min) nb = std::max(step_min, align_batch(g.nnt, nb / 2));  // the tapering tail
    if (split && t0 % h->batch == 0 && fine_density_batch(h, t0, std::min(h->batch, ntile - t0), group)) return 1;
    if (fine_mesh(h, t0, nb, pre_in_fft, a_mid, dt, nullptr, split ? &group : nullptr)) return 1;
    CK(h->rb.read(f2.data() + t0, h->f2max, sizeof(float) * nb, h->st));
    if (!pre_in_fft) {
      FftGeom fb = h->fg; fb.nbatch = nb;
      k_prefix_rows<<<dim3(592, nb), 256, 0, h->st>>>(fb, h->F, a_mid, dt); CKL();
      h->launches++;
    }
    if (kick_c_per_batch && overlap && t0 == 0) CK(cudaStreamWaitEvent(h->st, h->ev_join, 0));
    if (merged) {
      PhaseTimer pt(h, PH_FKICK);  // both kicks
      if (launch_kick(h, t0, nb, h->F, h->fc, vscale(h->sigma_vi), S_new)) return 1;
    } else {
      {
        PhaseTimer pt(h, PH_FKICK);
        if (run_fine_kick(h, t0, nb, S_new)) return 1;
      }
      if (vp_host) {
        const long long c_begin = (long long)t0 * nt3, c_end = (long long)(t0 + nb) * nt3;
        PhaseTimer pc(h, PH_CKICK);
        if (run_coarse_kick(h, VTab{h->tanlut, h->tanh, h->enc, h->dvlut2, h->divok2, h->vt_hot}, S_new, c_begin, c_end)) return 1;
      }
    }
    if (vp_host) {
      CK(cudaEventRecord(h->ev_copy[0], h->st));
      CK(cudaStreamWaitEvent(h->st_copy, h->ev_copy[0], 0));
      const long long p_begin = tile_start[t0], p_end = tile_start[t0 + nb];
      if (p_end > p_begin)
        CK(cudaMemcpyAsync((char*)vp_host + (size_t)3 * h->zv * p_begin, (char*)h->vp + (size_t)3 * h->zv * p_begin, (size_t)3 * h->zv * (p_end - p_begin), cudaMemcpyDeviceToHost, h->st_copy));
      CK(cudaEventRecord(h->ev_copy[1], h->st_copy));
      h->copy_pending = true;
    }
  }
  h->sigma_vi = h->sigma_vi_new;  // pm.f90:122
  if (build_dvlut(h, h->sigma_vi)) return 1;
  if (overlap) { CK(cudaStreamWaitEvent(h->st, h->ev_join, 0)); }
  else if (!kick_c_per_batch && coarse_mesh(h, true, a_mid, dt, nullptr)) return 1;
  float f2c = 0; unsigned long long vb4[4] = {0, 0, 0, 0};
  if (!kick_c_per_batch) {
    PhaseTimer pt(h, PH_CKICK);
    if (run_coarse_kick(h, vtab(h), vscale(h->sigma_vi), 0, g.ncell_p)) return 1;
  }
  CK(h->rb.read(&f2c, h->f2max + h->batch, sizeof(float), h->st));
  CK(h->rb.read(vb4, h->vmax_bits, sizeof vb4, h->st));
  CK(h->rb.sync(h->st));
  double vmd; memcpy(&vmd, &vb4[0], sizeof vmd);
  const float vmax = (float)vmd;  // f32 <- max(f32, f64) is monotone, so one final rounding is the same
  for (int d = 0; d < 3; d++) { double t; memcpy(&t, &vb4[1 + d], sizeof t); h->vmax3[d] = (float)t; }
  float f2f = 0; for (float v : f2) f2f = std::max(f2f, v);
  if (pre_in_fft) f2f = f2f / pscale / pscale;
  float vmax_all = vmax;
  if (h->nimg > 1) {  // pm.f90:239-244: every image takes the minimum of every image's dt = the dt of the maxima
    struct Rec { float f2f, f2c, vmax; } mine = {f2f, f2c, vmax};
    std::vector<Rec> all(h->nimg);
    CC(h->comm->allgather_host(&mine, all.data(), sizeof(Rec), h->st));
    for (int m = 0; m < h->nimg; m++) { f2f = std::max(f2f, all[m].f2f); f2c = std::max(f2c, all[m].f2c); vmax_all = std::max(vmax_all, all[m].vmax); }
  }
  h->last_f2max_fine = f2f;
  // pm.f90:233-236, all f32
  const float GG = 1.0f / 6.0f / PI_F;
  if (dt_fine) *dt_fine = sqrtf(1.0f / (sqrtf(f2f) * a_mid * GG));
  if (dt_coarse) *dt_coarse = sqrtf((float)NCELL / (sqrtf(f2c) * a_mid * GG));
  if (dt_vmax) *dt_vmax = 0.9f * 20 / vmax_all;
  if (vmax_out) *vmax_out = vmax;
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Two particle species sharing the meshes (CUBEnu -DNEUTRINOS: pm.f90:79-99,160,235,356 deposit xp_nu with mass_p_nu into the same
// rho_f / r3 and kick vp_nu with the same forces; update_particle.f90:326-353 drifts them through their own rhoc_nu / vfield_nu with
// sigma_vi_nu).  Every species is a handle of its own -- its own codes, zip format, cell arrays, sigma_vi, capacities, ghost exchange:
// cube_gpu_upload / update_x / buffer / download act per handle exactly as for one species.  This call is particle_mesh for both:
// `h` owns the meshes (fine pipeline, coarse FFT), `h2`'s particles are deposited into them after h's and kicked from them.
//   mass_p of a species: cube_gpu_set_mass_p (sim%mass_p_cdm, sim%mass_p_nu; the single-species default nf_global^3/npglobal would
//   count the mean density twice).
extern "C" int cube_gpu_set_mass_p(cube_handle* h, float mass_p) {
  if (!h) return fail("null handle");
  if (!(mass_p > 0.f)) return fail("cube_gpu_set_mass_p: mass_p must be positive");
  h->mass_p = mass_p;
  return 0;
}

extern "C" int cube_gpu_particle_mesh_species(cube_handle* h, cube_handle* h2, float a_mid, float dt, float* dt_fine, float* dt_coarse,
                                              float* dt_vmax, float* vmax_out, float* dt_vmax2, float* vmax2_out) {
  if (!h || !h2) return fail("null handle");
  if (h == h2) return fail("cube_gpu_particle_mesh_species: the two species must be different handles");
  g_cur = h; h->rb.drop();
  if (enter(h)) return 1;
  if (h2->p.device != h->p.device) return fail("cube_gpu_particle_mesh_species: both species must live on the same device");
  const Geom& g = h->g;
  if (memcmp(&h2->g, &g, sizeof(Geom)) != 0) return fail("cube_gpu_particle_mesh_species: the species differ in geometry (nn, nnt, nc, image)");
  if (!h->buffered || !h2->buffered) return fail("cube_gpu_particle_mesh_species: both states must be buffered (call cube_gpu_buffer on each)");
  for (cube_handle* q : {h, h2}) {
    if (q->copy_pending && q->copy_reads_vp) { CK(cudaStreamWaitEvent(h->st, q->ev_copy[1], 0)); q->copy_reads_vp = false; }
    q->vp_stream_host = nullptr;
  }
  // everything below runs on h's stream: wait for what the second species still has in flight on its own, and lend it ours
  CK(cudaEventRecord(h2->ev_fork, h2->st)); CK(cudaStreamWaitEvent(h->st, h2->ev_fork, 0));
  if (h2->vghost_pending) { CK(cudaStreamWaitEvent(h->st, h2->ev_vghost, 0)); }
  cudaStream_t st2 = h2->st;
  h2->st = h->st;
  int rc = 0;
  float f2f = 0, f2c = 0, vmax[2] = {0, 0};
  const int ntile = g.nnt * g.nnt * g.nnt;
  std::vector<float> f2(ntile, 0.f);
  const float pscale = ((1.0f * a_mid) * dt) / 6.0f / PI_F;
  const bool pre_in_fft = pscale > 1e-12f && pscale < 1e12f;
  do {
    for (cube_handle* q : {h, h2}) {
      if ((rc = build_dvlut(q, q->sigma_vi))) break;
      if ((rc = build_dvlut2(q, q->sigma_vi_new))) break;
      if (cudaMemsetAsync(q->vmax_bits, 0, 4 * sizeof(unsigned long long), h->st) != cudaSuccess) { rc = fail("memset"); break; }
    }
    if (rc) break;
    if ((rc = coarse_mesh(h, true, a_mid, dt, nullptr, h2))) break;
    for (int t0 = 0; t0 < ntile && !rc; t0 += h->batch) {
      const int nb = std::min(h->batch, ntile - t0);
      if ((rc = fine_mesh(h, t0, nb, pre_in_fft, a_mid, dt, h2))) break;
      if (h->rb.read(f2.data() + t0, h->f2max, sizeof(float) * nb, h->st) != cudaSuccess) { rc = fail("f2max copy"); break; }
      if (!pre_in_fft) {
        FftGeom fb = h->fg; fb.nbatch = nb;
        k_prefix_rows<<<dim3(592, nb), 256, 0, h->st>>>(fb, h->F, a_mid, dt);
        h->launches++;
      }
      PhaseTimer pt(h, PH_FKICK);
      if ((rc = run_fine_kick(h, t0, nb, vscale(h->sigma_vi_new)))) break;                 // pm.f90:88-118
      if ((rc = run_fine_kick(h2, t0, nb, vscale(h2->sigma_vi_new), h))) break;            // ... and its NEUTRINOS block
    }
    if (rc) break;
    for (cube_handle* q : {h, h2}) {
      q->sigma_vi = q->sigma_vi_new;  // pm.f90:122
      if ((rc = build_dvlut(q, q->sigma_vi))) break;
    }
    if (rc) break;
    {
      PhaseTimer pt(h, PH_CKICK);
      if ((rc = run_coarse_kick(h, vtab(h), vscale(h->sigma_vi), 0, g.ncell_p))) break;     // pm.f90:196-228
      if ((rc = run_coarse_kick(h2, vtab(h2), vscale(h2->sigma_vi), 0, g.ncell_p, h))) break;
    }
    unsigned long long vb[2] = {0, 0};
    if (h->rb.read(&f2c, h->f2max + h->batch, sizeof(float), h->st) != cudaSuccess ||
        h->rb.read(&vb[0], h->vmax_bits, sizeof(unsigned long long), h->st) != cudaSuccess ||
        h->rb.read(&vb[1], h2->vmax_bits, sizeof(unsigned long long), h->st) != cudaSuccess ||
        h->rb.sync(h->st) != cudaSuccess) { rc = fail("cube_gpu_particle_mesh_species: %s", cudaGetErrorString(cudaGetLastError())); break; }
    for (int q = 0; q < 2; q++) { double t; memcpy(&t, &vb[q], sizeof t); vmax[q] = (float)t; }
    for (float v : f2) f2f = std::max(f2f, v);
    if (pre_in_fft) f2f = f2f / pscale / pscale;
  } while (0);
  h2->st = st2;
  if (rc) { abort_current_comm(); return 1; }
  float vall[2] = {vmax[0], vmax[1]};
  if (h->nimg > 1) {  // pm.f90:239-244
    struct Rec { float f2f, f2c, v0, v1; } mine = {f2f, f2c, vmax[0], vmax[1]};
    std::vector<Rec> all(h->nimg);
    CC(h->comm->allgather_host(&mine, all.data(), sizeof(Rec), h->st));
    for (int m = 0; m < h->nimg; m++) { f2f = std::max(f2f, all[m].f2f); f2c = std::max(f2c, all[m].f2c); vall[0] = std::max(vall[0], all[m].v0); vall[1] = std::max(vall[1], all[m].v1); }
  }
  h->last_f2max_fine = f2f;
  const float GG = 1.0f / 6.0f / PI_F;
  if (dt_fine) *dt_fine = sqrtf(1.0f / (sqrtf(f2f) * a_mid * GG));
  if (dt_coarse) *dt_coarse = sqrtf((float)NCELL / (sqrtf(f2c) * a_mid * GG));
  if (dt_vmax) *dt_vmax = 0.9f * 20 / vall[0];
  if (dt_vmax2) *dt_vmax2 = 0.9f * 20 / vall[1];
  if (vmax_out) *vmax_out = vmax[0];
  if (vmax2_out) *vmax2_out = vmax[1];
  return 0;
}

// ---------------------------------------------------------------------------------------------
// cicpower + powerspectrum (CUBE/utilities/cicpower.f90:70-167, powerspectrum.f90:21-108, linear_kbin) of the resident state, with
// the library's own deposit kernel (cell-centred variant) and cuFFT.  xi(10,nbin) row-major like the reference's xi(10,nbin):
// rows 0 count, 1 k [h/Mpc], 2 = 3 = 4 Delta^2 (auto power), 5-6 kernels, 7 r = 1, 8 b = 1, 9 = row 2.  Single image.
extern "C" int cube_gpu_power_spectrum(cube_handle* h, float box, double* xi, int nbin_cap, int* nbin_out) {
  if (!h || !xi) return fail("null argument");
  if (enter(h)) return 1;
  if (!h->buffered) return fail("cube_gpu_power_spectrum: state is not buffered (call cube_gpu_buffer first)");
  if (h->nimg > 1) return fail("cube_gpu_power_spectrum: single image (a distributed transform of the global fine grid is not built)");
  const Geom& g = h->g;
  const int n = NCELL * g.nc, m = n + 4;  // ng = nf (cicpower.f90: ng=nf); deposit grid with the one-node frame shift
  const int nbin = (int)lround((double)(n / 2) * sqrt(3.0));
  if (nbin_out) *nbin_out = nbin;
  if (nbin_cap < nbin) return fail("cube_gpu_power_spectrum: xi holds %d bins, %d needed", nbin_cap, nbin);
  const long long tot = (long long)n * n * n, vol = (long long)n * n * (n + 2);
  float *dep = nullptr, *rho = nullptr; double *part = nullptr, *bins = nullptr;
  const unsigned nb = nblk(tot, 256);
  CK(dmalloc(&dep, (long long)m * m * m)); CK(dmalloc(&rho, vol)); CK(dmalloc(&part, (long long)nb + 1)); CK(dmalloc(&bins, PS_Q * nbin));
  int rc = 0;
  cufftHandle pl = 0;
  do {
    FineRegion R;
    for (int d = 0; d < 3; d++) { R.c0[d] = -1; R.c1[d] = g.nc + 1; R.f0[d] = 0; R.n[d] = m; }
    R.ldy = m; R.ldz = (long long)m * m;
    {  // cell-centred CIC of all particles (ghost cells alias the periodic image: the wrap of cicpower.f90:118-126 comes for free)
      using C = Fd844;
      const unsigned nbx = (m + C::NX - 1) / C::NX, nby = (m + C::NY - 1) / C::NY, nbz = (m + C::NZ - 1) / C::NZ;
      if (h->zx == 2) {
        if (cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, false, short, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) { rc = fail("power spectrum: kernel attribute"); break; }
        k_fine_deposit_r<C, false, short, true><<<nbx * nby * nbz, C::NT, C::SMEM, h->st>>>(g, R, make_int3(FRAME_NONE, 0, 0), (const short*)h->xp, h->rhoc_e, h->cstart_e, h->mass_p, dep);
      } else {
        if (cudaFuncSetAttribute((const void*)k_fine_deposit_r<C, false, signed char, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM) != cudaSuccess) { rc = fail("power spectrum: kernel attribute"); break; }
        k_fine_deposit_r<C, false, signed char, true><<<nbx * nby * nbz, C::NT, C::SMEM, h->st>>>(g, R, make_int3(FRAME_NONE, 0, 0), (const signed char*)h->xp, h->rhoc_e, h->cstart_e, h->mass_p, dep);
      }
    }
    k_ps_extract<<<nb, 256, 0, h->st>>>(n, dep, rho, part);
    k_reduce_strided<<<1, 1024, 0, h->st>>>(part, nb, 1, 0, part + nb);
    k_ps_contrast<<<nb, 256, 0, h->st>>>(n, rho, part + nb);
    if (cudaGetLastError() != cudaSuccess) { rc = fail("power spectrum: kernel launch failed"); break; }
    int dims[3] = {n, n, n}, rembed[3] = {n, n, n + 2}, cembed[3] = {n, n, n / 2 + 1};
    if (cufftPlanMany(&pl, 3, dims, rembed, 1, (int)vol, cembed, 1, (int)(vol / 2), CUFFT_R2C, 1) != CUFFT_SUCCESS || cufftSetStream(pl, h->st) != CUFFT_SUCCESS ||
        cufftExecR2C(pl, rho, (cufftComplex*)rho) != CUFFT_SUCCESS) { rc = fail("power spectrum: cuFFT failed (n=%d)", n); break; }
    if (cudaMemsetAsync(bins, 0, sizeof(double) * PS_Q * nbin, h->st) != cudaSuccess) { rc = fail("power spectrum: memset"); break; }
    const size_t smem = sizeof(double) * PS_Q * nbin;
    if (cudaFuncSetAttribute((const void*)k_ps_bin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { rc = fail("power spectrum: %d bins do not fit shared memory", nbin); break; }
    k_ps_bin<<<2 * h->nsm, 256, smem, h->st>>>(n, nbin, (const float2*)rho, bins);
    std::vector<double> b(PS_Q * (size_t)nbin);
    if (cudaMemcpyAsync(b.data(), bins, sizeof(double) * b.size(), cudaMemcpyDeviceToHost, h->st) != cudaSuccess || cudaStreamSynchronize(h->st) != cudaSuccess) {
      rc = fail("power spectrum: %s", cudaGetErrorString(cudaGetLastError())); break;
    }
    h->launches += 5;
    for (int i = 0; i < nbin; i++) {  // powerspectrum.f90:98-106
      const double cnt = b[i];
      double* col = xi + i;
      col[0] = cnt;
      col[1 * nbin] = b[nbin + i] / cnt * (2.0 * (double)PI_F) / (double)box;
      const double p11 = b[2 * nbin + i] / cnt;
      col[2 * nbin] = col[3 * nbin] = col[4 * nbin] = p11;
      col[5 * nbin] = b[3 * nbin + i] / cnt; col[6 * nbin] = b[4 * nbin + i] / cnt;
      col[7 * nbin] = p11 / sqrt(p11 * p11); col[8 * nbin] = sqrt(p11 / p11); col[9 * nbin] = p11;
    }
  } while (0);
  if (pl) cufftDestroy(pl);
  cudaFree(dep); cudaFree(rho); cudaFree(part); cudaFree(bins);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// diagnostics
extern "C" int64_t cube_gpu_query(cube_handle* h, const char* what) {
  if (!h || !what) return -1;
  std::string w(what);
  if (w == "np_image_max") return h->np_image_max;
  if (w == "np_tile_max") return h->np_tile_max;
  if (w == "nfe") return h->g.nfe;
  if (w == "nft") return h->g.nft;
  if (w == "nt") return h->g.nt;
  if (w == "fine_batch") return h->batch;
  if (w == "drift_radius") return h->last_radius;
  if (w == "nfft") return h->fg.N;
  if (w == "nfft_pitch") return h->fg.P;
  if (w == "kernel_launches") return h->launches;
  if (w == "nplocal") return h->nplocal;
  return -1;
}
extern "C" int cube_gpu_get_kern_f(cube_handle* h, float* out) {
  if (enter(h)) return 1;
  const FftGeom& f = h->fg;  // strip the kx pitch: out(N/2+1, N, N, 3)
  CK(cudaMemcpy2D(out, sizeof(float) * f.NH, h->kern_f, sizeof(float) * f.P, sizeof(float) * f.NH, (size_t)3 * f.N * f.N,
                  cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int cube_gpu_get_kern_c(cube_handle* h, float* out) {
  if (enter(h)) return 1;
  if (h->nimg > 1) return fail("cube_gpu_get_kern_c: single-image diagnostic (the distributed kern_c lives in the transposed k-space layout)");
  CK(cudaMemcpy(out, h->kern_c, sizeof(float) * 3 * h->cnk, cudaMemcpyDeviceToHost));
  return 0;
}
static int tile_index(cube_handle* h, int itx, int ity, int itz, int* t) {
  const int n = h->g.nnt;
  if (itx < 1 || ity < 1 || itz < 1 || itx > n || ity > n || itz > n) return fail("tile index out of range");
  *t = ((itz - 1) * n + (ity - 1)) * n + (itx - 1);
  return 0;
}
extern "C" int cube_gpu_fine_density(cube_handle* h, int itx, int ity, int itz, float* rho_f) {
  if (enter(h)) return 1;
  if (!h->buffered) return fail("state is not buffered");
  int t; if (tile_index(h, itx, ity, itz, &t)) return 1;
  // diagnostic: deposit on the reference's whole padded grid rho_f(nfe+2,nfe,nfe) instead of the FFT window
  const Geom& g = h->g;
  const long long vol = (long long)g.nfe * g.nfe * (g.nfe + 2);
  float* tmp = nullptr; CK(dmalloc(&tmp, vol));
  CK(cudaMemsetAsync(tmp, 0, sizeof(float) * vol, h->st));
  const int tc[3] = {t % g.nnt, (t / g.nnt) % g.nnt, t / (g.nnt * g.nnt)};
  FineRegion R;
  for (int d = 0; d < 3; d++) { R.c0[d] = tc[d] * g.nt - (NCB - 1); R.c1[d] = (tc[d] + 1) * g.nt + (NCB - 1); R.f0[d] = 4 * tc[d] * g.nt - NFB; R.n[d] = g.nfe; }
  R.ldy = g.nfe + 2; R.ldz = R.ldy * g.nfe;
  const int3 frame = h->shared_region ? make_int3(FRAME_NONE, 0, 0) : make_int3(tc[0] * g.nt, tc[1] * g.nt, tc[2] * g.nt);
  if (fine_deposit(h, R, frame, tmp)) { cudaFree(tmp); return 1; }
  CK(cudaMemcpyAsync(rho_f, tmp, sizeof(float) * vol, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  cudaFree(tmp);
  return 0;
}
extern "C" int cube_gpu_fine_force(cube_handle* h, int itx, int ity, int itz, float* force_f) {
  if (enter(h)) return 1;
  if (!h->buffered) return fail("state is not buffered");
  int t; if (tile_index(h, itx, ity, itz, &t)) return 1;
  if (fine_mesh(h, t, 1, false, 0.f, 0.f)) return 1;
  const long long m = h->fg.M, n = m * m * m;
  float* tmp = nullptr; CK(dmalloc(&tmp, 3 * n));
  k_force_to_ref<<<nblk(n, 256), 256, 0, h->st>>>((int)m, h->fg.FP, h->F, tmp); CKL();
  CK(cudaMemcpyAsync(force_f, tmp, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  cudaFree(tmp);
  return 0;
}
extern "C" int cube_gpu_fine_kick_with(cube_handle* h, int itx, int ity, int itz, const float* force_f, float a_mid, float dt,
                                       float sigma_vi, float sigma_vi_new, float* f2_max) {
  if (enter(h)) return 1;
  if (!h->buffered) return fail("state is not buffered");
  int t; if (tile_index(h, itx, ity, itz, &t)) return 1;
  const long long m = h->fg.M, n = m * m * m;
  float* tmp = nullptr; CK(dmalloc(&tmp, 3 * n));
  CK(cudaMemcpyAsync(tmp, force_f, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, h->st));
  k_force_from_ref<<<nblk(n, 256), 256, 0, h->st>>>((int)m, h->fg.FP, tmp, h->F); CKL();
  if (build_dvlut(h, sigma_vi)) return 1;
  float f2 = 0;
  CK(cudaMemsetAsync(h->f2max, 0, sizeof(unsigned), h->st));
  FftGeom f1 = h->fg; f1.nbatch = 1;
  k_f2max_rows<<<dim3(592, 1), 256, 0, h->st>>>(f1, h->F, h->f2max); CKL();
  CK(cudaMemcpyAsync(&f2, h->f2max, sizeof(float), cudaMemcpyDeviceToHost, h->st));
  k_prefix_rows<<<dim3(592, 1), 256, 0, h->st>>>(f1, h->F, a_mid, dt); CKL();
  if (h->old_kick) { if (run_fine_kick(h, t, 1, vscale(sigma_vi_new))) return 1; }
  else { if (build_dvlut2(h, sigma_vi_new) || launch_kick(h, t, 1, h->F, nullptr, vscale(sigma_vi), vscale(sigma_vi_new))) return 1; }
  CK(h->rb.sync(h->st));
  cudaFree(tmp);
  if (f2_max) *f2_max = f2;
  return 0;
}
extern "C" int cube_gpu_coarse_density(cube_handle* h, float* r3) {
  if (enter(h)) return 1;
  if (!h->buffered) return fail("state is not buffered");
  const Geom& g = h->g;
  if (coarse_mesh(h, false, 0.f, 0.f, nullptr)) return 1;
  CK(cudaMemcpy2DAsync(r3, sizeof(float) * g.nc, h->r3, sizeof(float) * (h->nimg > 1 ? g.nc : g.nc + 2), sizeof(float) * g.nc, (size_t)g.nc * g.nc,
                       cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  return 0;
}
extern "C" int cube_gpu_coarse_force(cube_handle* h, float* force_c) {
  if (enter(h)) return 1;
  if (!h->buffered) return fail("state is not buffered");
  const long long m = h->g.nc + 2;
  float* raw = nullptr; CK(dmalloc(&raw, 3 * m * m * m));
  if (coarse_mesh(h, true, 0.f, 0.f, raw)) { cudaFree(raw); return 1; }
  CK(cudaMemcpyAsync(force_c, raw, sizeof(float) * 3 * m * m * m, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  cudaFree(raw);
  return 0;
}
extern "C" int cube_gpu_coarse_kick_with(cube_handle* h, const float* force_c, float a_mid, float dt, float sigma_vi, float* vmax,
                                         float* f2_max) {
  if (enter(h)) return 1;
  h->rb.drop();
  if (!h->buffered) return fail("state is not buffered");
  const Geom& g = h->g;
  const long long m = g.nc + 2;
  CK(cudaMemcpyAsync(h->fc, force_c, sizeof(float) * 3 * m * m * m, cudaMemcpyHostToDevice, h->st));
  if (build_dvlut(h, sigma_vi)) return 1;
  CK(cudaMemsetAsync(h->f2max + h->batch, 0, sizeof(unsigned), h->st));
  CK(cudaMemsetAsync(h->vmax_bits, 0, 4 * sizeof(unsigned long long), h->st));
  k_force_c_prefix<<<1184, 256, 0, h->st>>>(m * m * m, h->fc, a_mid, dt, h->f2max + h->batch); CKL();
  if (h->old_kick) {
    if (run_coarse_kick(h, vtab(h), vscale(sigma_vi), 0, g.ncell_p)) return 1;
  } else if (build_dvlut2(h, sigma_vi) || launch_kick(h, 0, g.nnt * g.nnt * g.nnt, nullptr, h->fc, vscale(sigma_vi), vscale(sigma_vi))) return 1;
  float f2c = 0; unsigned long long vb = 0;
  CK(h->rb.read(&f2c, h->f2max + h->batch, sizeof(float), h->st));
  CK(h->rb.read(&vb, h->vmax_bits, sizeof vb, h->st));
  CK(h->rb.sync(h->st));
  double vmd; memcpy(&vmd, &vb, sizeof vmd);
  if (vmax) *vmax = (float)vmd;
  if (f2_max) *f2_max = f2c;
  return 0;
}
// message plan of one image (host only, no device needed): rows of 8 int64
//   {0, rx, ry, rz, src_rank, dst_rank, ncell, cell0}   ghost direction (halo cells; particle messages follow the same pairs)
//   {1, rank, nplanes, 0...}                            force_c planes I send to `rank` (last hop of the coarse inverse)
//   {2, rank, nplanes, 0...}                            force_c planes I receive from `rank`
//   {3, R, Gx, Gy, Gz, sz, nyl, grp0}                   coarse transform geometry
extern "C" int cube_gpu_exchange_plan(const cube_params* p, int64_t* out, int cap_rows) {
  if (!p || !out) return -1;
  const int nimg = p->nn[0] * p->nn[1] * p->nn[2];
  if (p->nn[0] < 1 || p->nn[1] < 1 || p->nn[2] < 1 || p->rank < 0 || p->rank >= nimg || p->nnt < 1 || p->nc % p->nnt) return -1;
  Geom g; fill_geom(p, g);
  ExPlan P; build_exchange_plan(g, p->rank, P, false);
  CoarseGeom c; fill_coarse_geom(g, c);
  std::vector<FPlane> targets, sources;
  if (nimg > 1 && c.Gz % c.R == 0 && c.Gy % c.R == 0) force_plane_lists(g, c, p->rank, targets, sources);
  int n = 0;
  auto row = [&](long long a0, long long a1, long long a2, long long a3, long long a4, long long a5, long long a6, long long a7) {
    if (n < cap_rows) { int64_t* r = out + 8 * n; r[0] = a0; r[1] = a1; r[2] = a2; r[3] = a3; r[4] = a4; r[5] = a5; r[6] = a6; r[7] = a7; }
    n++;
  };
  for (const ExDir& d : P.dirs) row(0, d.r[0], d.r[1], d.r[2], d.src_rank, d.dst_rank, d.ncell, d.cell0);
  for (const FPlane& t : targets) row(1, t.rank, (long long)t.zz.size(), 0, 0, 0, 0, 0);
  for (const FPlane& t : sources) row(2, t.rank, (long long)t.zz.size(), 0, 0, 0, 0, 0);
  row(3, c.R, c.Gx, c.Gy, c.Gz, c.sz, c.nyl, c.grp0);
  return n;
}

// table-driven velocity code conversions against their defining formulas (see k_selftest_encode / k_selftest_decode)
extern "C" int cube_gpu_selftest_codes(cube_handle* h, float sigma_vi, int64_t nsweep, int64_t* bad_encode, int64_t* bad_decode, int* fma_division) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  if (build_dvlut(h, sigma_vi)) return 1;
  unsigned long long* cnt = nullptr; CK(dmalloc(&cnt, 2));
  CK(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), h->st));
  const long long n = 3LL * (h->nvbin / 2 - 1) + std::max<long long>(0, nsweep);
  if (h->zv == 2) {
    k_selftest_encode<16><<<nblk(n, 256), 256, 0, h->st>>>(h->enc, std::max<long long>(0, nsweep), cnt); CKL();
    k_selftest_decode<16><<<1, PW_T, pw_smem_bytes(h->vt_hot), h->st>>>(vtab(h), vscale(sigma_vi), cnt + 1); CKL();
  } else {
    k_selftest_encode<8><<<nblk(n, 256), 256, 0, h->st>>>(h->enc, std::max<long long>(0, nsweep), cnt); CKL();
    k_selftest_decode<8><<<1, PW_T, pw_smem_bytes(h->vt_hot), h->st>>>(vtab(h), vscale(sigma_vi), cnt + 1); CKL();
  }
  unsigned long long out[2] = {0, 0}; int ok = 0;
  CK(cudaMemcpyAsync(out, cnt, sizeof out, cudaMemcpyDeviceToHost, h->st));
  CK(cudaMemcpyAsync(&ok, h->divok, sizeof ok, cudaMemcpyDeviceToHost, h->st));
  CK(h->rb.sync(h->st));
  cudaFree(cnt);
  if (bad_encode) *bad_encode = (int64_t)out[0];
  if (bad_decode) *bad_decode = (int64_t)out[1];
  if (fma_division) *fma_division = (ok != 0 && h->vt_hot) ? 1 : 0;
  return 0;
}

// CUBEnu's bookkeeping of the same arithmetic: the order in which update_xp visits the source planes and the
// per-component vmax.  nlayer = 2*ceiling(dt_mid*sim%vz_max/ncell)+1 (CUBEnu update_particle.f90:37), 0 or 1 = CUBE/main.
extern "C" int cube_gpu_set_drift_layers(cube_handle* h, int nlayer) {
  if (!h) return fail("null handle");
  if (nlayer < 0) return fail("cube_gpu_set_drift_layers: nlayer must be >= 0");
  h->nlayer = nlayer < 1 ? 1 : nlayer;
  return 0;
}
extern "C" int cube_gpu_get_vmax3(cube_handle* h, float vmax3[3]) {
  if (!h || !vmax3) return fail("null argument");
  for (int d = 0; d < 3; d++) vmax3[d] = h->vmax3[d];
  return 0;
}

extern "C" int cube_gpu_phase_count(void) { return PH_N; }
extern "C" const char* cube_gpu_phase_name(int i) { return (i >= 0 && i < PH_N) ? kPhaseNames[i] : ""; }
extern "C" int cube_gpu_phase_times(cube_handle* h, float* ms) {
  for (int i = 0; i < PH_N; i++) { ms[i] = h->phase_ms[i]; h->phase_ms[i] = 0; }
  return 0;
}
extern "C" int cube_gpu_timer(cube_handle* h, int start, float* ms) {
  if (!h) return fail("null handle");
  if (enter(h)) return 1;
  if (start) { CK(cudaEventRecord(h->tev[0], h->st)); return 0; }
  if (h->vghost_pending) CK(cudaStreamWaitEvent(h->st, h->ev_vghost, 0));  // the stopwatch covers a ghost exchange still in flight on the side stream
  if (h->xghost_pending) CK(cudaStreamWaitEvent(h->st, h->ev_xghost, 0));
  CK(cudaEventRecord(h->tev[1], h->st));
  CK(cudaEventSynchronize(h->tev[1]));
  if (ms) CK(cudaEventElapsedTime(ms, h->tev[0], h->tev[1]));
  return 0;
}
extern "C" int cube_gpu_set_profiling(cube_handle* h, int on) { h->prof = on != 0; return 0; }
