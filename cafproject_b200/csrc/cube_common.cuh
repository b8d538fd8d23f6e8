// cube_common.cuh -- geometry, code conversions and small device helpers shared by all kernels.
//
// Exact-arithmetic rules: every operation that feeds an integer code or a
// cell index is written with the round-to-nearest intrinsics (__dadd_rn, __dmul_rn, __fmul_rn, ...)
// so that nvcc can never contract a multiply-add into an FMA; the reference build (gfortran -O3,
// x86-64 baseline) has no FMA.  f64 division, sqrt, ceil, floor and llround are IEEE-exact on the
// device; f32 division uses __fdiv_rn.
#pragma once
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <cuda_runtime.h>

namespace cube {

constexpr float PI_F = 3.14159274101257324f;  // parameters.f90:73  pi=4*atan(1.) (f32, 0x40490FDB)
constexpr int NCB = 6;                        // parameters.f90:46
constexpr int NCELL = 4;                      // parameters.f90:21
constexpr int NFB = NCB * NCELL;              // parameters.f90:50

struct Geom {
  int nn[3];   // images per dim
  int ic[3];   // this image's coordinates, 0-based (icx-1, icy-1, icz-1)
  int nnt;     // tiles / image / dim
  int nc;      // coarse cells / image / dim
  int nt;      // coarse cells / tile / dim
  int nte;     // nt + 2 ncb
  int nft;     // 4 nt
  int nfe;     // nft + 2 nfb
  int ne;      // nc + 2 ncb : extended image grid (ghost layers alias the periodic image when nn_d == 1)
  long long ncell_p;  // nc^3
  long long ncell_e;  // ne^3
};

// ---- index helpers ---------------------------------------------------------------------------
// file-order ("disjoint state") linear index of physical cell: tile-major, then k, j, i (0-based)
__host__ __device__ inline long long phys_index(const Geom& g, int tx, int ty, int tz, int i, int j, int k) {
  long long nt = g.nt;
  return (((long long)(tz * g.nnt + ty) * g.nnt + tx) * nt * nt * nt) + ((long long)k * nt + j) * nt + i;
}
// inverse of phys_index
__host__ __device__ inline void phys_decompose(const Geom& g, long long L, int& tx, int& ty, int& tz, int& i, int& j, int& k) {
  long long nt = g.nt, nt3 = nt * nt * nt;
  long long t = L / nt3, c = L - t * nt3;
  tx = (int)(t % g.nnt); ty = (int)((t / g.nnt) % g.nnt); tz = (int)(t / ((long long)g.nnt * g.nnt));
  i = (int)(c % nt); j = (int)((c / nt) % nt); k = (int)(c / (nt * nt));
}
// extended image grid index; x,y,z image-local 0-based in [-ncb, nc+ncb)
__host__ __device__ inline long long ext_index(const Geom& g, int x, int y, int z) {
  return ((long long)(z + NCB) * g.ne + (y + NCB)) * g.ne + (x + NCB);
}

// true if extended-grid cell (x,y,z) is owned by another image (a real ghost), false if it aliases this image
__host__ __device__ inline bool ext_is_remote(const Geom& g, int x, int y, int z) {
  return ((x < 0 || x >= g.nc) && g.nn[0] > 1) || ((y < 0 || y >= g.nc) && g.nn[1] > 1) || ((z < 0 || z >= g.nc) && g.nn[2] > 1);
}

// ---- codes -----------------------------------------------------------------------------------
// The zip format of a run (CUBE/main/universe*.fh, parameters.f90:13-15): izipx, izipv = bytes per position / velocity code.
//   x_resolution = 2^-(8 izipx), ishift = -2^(8 izipx - 1), rshift = 0.5 - ishift;  nvbin = 2^(8 izipv)
// Kernels are instantiated per format; a code travels as a sign-extended `short` in registers whatever its storage type.
template <int ZX, int ZV>
struct Fmt {
  static_assert((ZX == 1 || ZX == 2) && (ZV == 1 || ZV == 2), "izipx, izipv are 1 or 2 (universe*.fh)");
  static constexpr int ZXB = ZX, ZVB = ZV;
  static constexpr int XB = 8 * ZX, VB = 8 * ZV;      // bits per code
  using XT = typename std::conditional<ZX == 1, signed char, short>::type;
  using VT = typename std::conditional<ZV == 1, signed char, short>::type;
  static constexpr int NV = (1 << VB) - 1;           // nvbin-1
  static constexpr int VHALF = 1 << (VB - 1);        // velocity codes are -VHALF .. VHALF-1; size of the half tables
};
// raw B-bit pattern of a sign-extended code
template <int B> __host__ __device__ __forceinline__ unsigned upat(short c) { return (unsigned)(int)c & ((1u << B) - 1u); }
// int(xp+ishift,izipx)+rshift == u + 0.5 with u the raw pattern (parameters.f90:14-15)
template <int XB> __device__ __forceinline__ double xp_frac(short xp) {  // (u+0.5)*x_resolution, exact
  return ((double)upat<XB>(xp) + 0.5) * (1.0 / (double)(1 << XB));
}
// nint(real(nvbin-1)*atan(S*v)/pi,kind=izipv)  (pm.f90:113, update_particle.f90:86)
template <int VB> __device__ __forceinline__ short vp_encode(double v, double S) {
  double t = __dmul_rn((double)((1 << VB) - 1), atan(__dmul_rn(S, v)));
  return (short)llround(t / (double)PI_F);
}
// sqrt(pi/2)/(sigma_vi*vrel_boost): f32 sqrt promoted, f64 product (update_particle.f90:42)
__host__ __device__ inline double vscale(float sigma) {
  return (double)sqrtf(PI_F / 2) / ((double)sigma * 2.5);
}

// CIC split of an f32 coordinate (pm.f90:55-58): idx1=floor(x)+1, dx1=idx1-x, dx2=1-dx1
__device__ __forceinline__ void cic_split(float tempx, int& idx1, float& dx1, float& dx2) {
  idx1 = (int)floorf(tempx) + 1;
  dx1 = __fsub_rn((float)idx1, tempx);
  dx2 = __fsub_rn(1.0f, dx1);
}

// one CIC term of the kick: F*a_mid*dt/6/pi*wx*wy*wz, all f32, left to right (pm.f90:104)
__device__ __forceinline__ float kick_term(float F, float a_mid, float dt, float wx, float wy, float wz) {
  float t = __fmul_rn(F, a_mid);
  t = __fmul_rn(t, dt);
  t = __fdiv_rn(t, 6.0f);
  t = __fdiv_rn(t, PI_F);
  t = __fmul_rn(t, wx);
  t = __fmul_rn(t, wy);
  return __fmul_rn(t, wz);
}

// IEEE-correct a/c for the constants c = 6 and c = pi_f without the generic division routine: q0 = RN(a*rc),
// e = a - c*q0 (exact with an FMA), q = RN(q0 + e*rc) with rc = RN(1/c).  Verified exhaustively against a/c for
// all 2^23 mantissas (tests/test_oracle_pins.py::test_fma_division_by_constants); tiny/huge operands take __fdiv_rn.
__device__ __forceinline__ float div_const_rn(float a, float c, float rc) {
  const float aa = fabsf(a);
  if (aa != 0.f && (aa < 1e-30f || aa > 1e30f)) return __fdiv_rn(a, c);
  const float q0 = __fmul_rn(a, rc);
  const float e = __fmaf_rn(-c, q0, a);
  return __fmaf_rn(e, rc, q0);
}
// F*a_mid*dt/6/pi in the reference's order (pm.f90:104): the per-node prefix of every kick term.  One range test covers
// both divisions: with |t| in [1e-28, 1e28] (or 0) t/6 is still inside div_const_rn's verified range.
__device__ __forceinline__ float kick_prefix(float F, float a_mid, float dt) {
  const float t = __fmul_rn(__fmul_rn(F, a_mid), dt);
  const float aa = fabsf(t);
  if (aa != 0.f && (aa < 1e-28f || aa > 1e28f)) return __fdiv_rn(__fdiv_rn(t, 6.0f), PI_F);
  const float q0 = __fmul_rn(t, 1.0f / 6.0f);
  const float q = __fmaf_rn(__fmaf_rn(-6.0f, q0, t), 1.0f / 6.0f, q0);
  const float r0 = __fmul_rn(q, 1.0f / PI_F);
  return __fmaf_rn(__fmaf_rn(-PI_F, r0, q), 1.0f / PI_F, r0);
}
__device__ __forceinline__ float kick_weight(float G, float wx, float wy, float wz) {
  return __fmul_rn(__fmul_rn(__fmul_rn(G, wx), wy), wz);
}

// the three codes of one particle (AoS: only the code's own alignment is guaranteed), sign-extended
struct Code3 { short x, y, z; };
template <class T> __device__ __forceinline__ Code3 load_code3(const T* __restrict__ a, long long p) {
  const T* q = a + 3 * p;
  Code3 c; c.x = (short)__ldg(q); c.y = (short)__ldg(q + 1); c.z = (short)__ldg(q + 2);
  return c;
}
template <class T> __device__ __forceinline__ void store_code3(T* a, long long p, short x, short y, short z) {
  T* q = a + 3 * p; q[0] = (T)x; q[1] = (T)y; q[2] = (T)z;
}

// drift key: destination-minus-source cell offset per dim (5 bits each, biased by 16) + near-tie flag
constexpr unsigned KEY_FLAG = 0x8000u;
__host__ __device__ inline unsigned key_pack(int dx, int dy, int dz) {
  return (unsigned)(dx + 16) | ((unsigned)(dy + 16) << 5) | ((unsigned)(dz + 16) << 10);
}

// ---------------------------------------------------------------------------------------------
// Small device<->host messages that do not use the copy engines.  A few bytes copied with cudaMemcpyAsync wait in the copy engine's
// queue behind whatever large transfer is in flight on ANY stream -- with a checkpoint streaming out under particle_mesh, each
// 4-byte f2_max read-back waited for the 0.8 GB of positions in front of it and the step lost all of the overlap (measured: 105 ms
// per end-to-end step, equal to the sum of its serialised parts).  These go through mapped pinned memory instead: a one-warp kernel
// on the compute stream moves the words, the host reads them after the stream synchronisation it does anyway.
// ---------------------------------------------------------------------------------------------
__global__ void k_copy_words(unsigned* __restrict__ dst, const unsigned* __restrict__ src, int nwords) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
}
__global__ void k_copy_ll_strided(long long* __restrict__ dst, const long long* __restrict__ src, long long stride, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[(long long)i * stride];
}
struct HostStage {
  static constexpr size_t CAP = 1 << 16;
  char* buf = nullptr;  // cudaHostAllocMapped: the same address on both sides (unified addressing)
  size_t used = 0;
  struct Pending { void* dst; size_t off, bytes; cudaStream_t st; };
  Pending pend[256];
  int npend = 0;
  cudaError_t init() { return buf ? cudaSuccess : cudaHostAlloc((void**)&buf, CAP, cudaHostAllocMapped | cudaHostAllocPortable); }
  void destroy() { if (buf) cudaFreeHost(buf); buf = nullptr; }
  void drop() { npend = 0; used = 0; }  // at an entry point: forget read-backs of a call that returned on an error before its sync()
  // device -> host, words of 4 bytes; `dst` is valid after sync()
  cudaError_t read(void* dst, const void* src, size_t bytes, cudaStream_t st, long long stride_ll = 0) {
    const size_t need = (bytes + 15) & ~(size_t)15;
    if (!buf || (bytes & 3) || npend == 256 || used + need > CAP) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
    if (stride_ll) k_copy_ll_strided<<<1, 128, 0, st>>>((long long*)(buf + used), (const long long*)src, stride_ll, (int)(bytes / 8));
    else k_copy_words<<<1, 128, 0, st>>>((unsigned*)(buf + used), (const unsigned*)src, (int)(bytes / 4));
    pend[npend++] = Pending{dst, used, bytes, st};
    used += need;
    return cudaGetLastError();
  }
  cudaError_t sync(cudaStream_t st) {
    cudaError_t e = cudaStreamSynchronize(st);
    for (int i = 0; i < npend && e == cudaSuccess; i++)
      if (pend[i].st != st) e = cudaStreamSynchronize(pend[i].st);
    if (e == cudaSuccess)
      for (int i = 0; i < npend; i++) memcpy(pend[i].dst, buf + pend[i].off, pend[i].bytes);
    npend = 0; used = 0;
    return e;
  }
  // host -> device through the same memory (the caller's bytes are copied now; the kernel reads them in stream order)
  cudaError_t write(void* dst_dev, const void* src, size_t bytes, cudaStream_t st) {
    const size_t need = (bytes + 15) & ~(size_t)15;
    if (!buf || (bytes & 3) || used + need > CAP) return cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st);
    memcpy(buf + used, src, bytes);
    k_copy_words<<<1, 128, 0, st>>>((unsigned*)dst_dev, (const unsigned*)(buf + used), (int)(bytes / 4));
    used += need;
    return cudaGetLastError();
  }
};

}  // namespace cube
