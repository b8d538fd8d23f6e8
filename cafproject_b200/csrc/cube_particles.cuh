// cube_particles.cuh -- particle-parallel kernels (one thread per particle) on CUBE's cell-ordered storage.
//
// A CTA owns PC_CELLS consecutive coarse cells in file order; their particles are one contiguous run of the
// int16 arrays, so a warp's loads/stores of xp/vp are coalesced.  Each thread finds the cell of its particle by a
// binary search in the CTA's prefix offsets (shared memory).  All arithmetic that feeds an integer code keeps the
// reference's operation order and rounding (cube_common.cuh).
#pragma once
#include "cube_common.cuh"

namespace cube {

constexpr int PC_CELLS = 128;  // file-order coarse cells per CTA
constexpr int PC_T = 256;      // threads per CTA

// prefix offsets of the CTA's cells relative to its first particle; cstart has ncell+1 entries (sentinel = total)
__device__ __forceinline__ int chunk_setup(const long long* __restrict__ cstart, long long c0, long long ncell, int* soff) {
  const long long base = cstart[c0];
  for (int t = threadIdx.x; t <= PC_CELLS; t += blockDim.x) {
    const long long c = c0 + t < ncell ? c0 + t : ncell;
    soff[t] = (int)(cstart[c] - base);
  }
  __syncthreads();
  return soff[PC_CELLS];
}
// cell (0..PC_CELLS-1) of particle q: largest c with soff[c] <= q  (empty cells have soff[c] == soff[c+1])
__device__ __forceinline__ int chunk_find(const int* soff, int q) {
  int lo = 0;
#pragma unroll
  for (int step = PC_CELLS / 2; step > 0; step >>= 1)
    if (soff[lo + step] <= q) lo += step;
  return lo;
}

// ---------------------------------------------------------------------------------------------
// velocity encode without atan:  code = nint(65535*atan(X)/pi_f) is a monotone step function of X = S*v; the
// table holds, for c = 0..32767, the smallest double X with code(X) >= c+1, found at start-up by bisection on
// the double bit pattern with the exact formula (same device atan), so the lookup reproduces it for every X.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long enc_exact(double X) {
  return llround(__dmul_rn(65535.0, atan(X)) / (double)PI_F);
}
__global__ void k_build_enc(double* __restrict__ B) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > 32767) return;
  if (c == 32767) { B[c] = __longlong_as_double(0x7ff0000000000000LL); return; }
  unsigned long long lo = 0, hi = (unsigned long long)__double_as_longlong(1e300);
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (enc_exact(__longlong_as_double((long long)mid)) >= c + 1) hi = mid; else lo = mid + 1;
  }
  B[c] = __longlong_as_double((long long)hi);
}
// nint(real(nvbin-1)*atan(S*v)/pi,kind=izipv)  (pm.f90:113, update_particle.f90:86)
__device__ __forceinline__ short vp_encode_lut(double v, double S, const double* __restrict__ B) {
  const double X = __dmul_rn(S, v);
  const double a = fabs(X);
  int c = __float2int_rn(atanf((float)a) * (65535.0f / PI_F));
  c = min(max(c, 0), 32767);
  while (c < 32767 && a >= __ldg(B + c)) c++;
  while (c > 0 && a < __ldg(B + c - 1)) c--;
  return (short)(X < 0.0 ? -c : c);
}

// ---------------------------------------------------------------------------------------------
// fine kick (pm.f90:88-118) for the tiles [tile0, tile0+nb): grid = (chunks per tile, nb)
// F[b][z'][y'][d][x'] = force_f on the M kept points
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PC_T) k_fine_kick_p(Geom g, int tile0, int M, int FP, const short* __restrict__ xp, short* __restrict__ vp,
                                                     const long long* __restrict__ cstart_p, const float* __restrict__ G,
                                                     const double* __restrict__ dvlut, const double* __restrict__ enc, double S_new,
                                                     float a_mid, float dt) {
  __shared__ int soff[PC_CELLS + 1];
  const long long nt3 = (long long)g.nt * g.nt * g.nt;
  const int b = blockIdx.y;
  const long long tbase = (long long)(tile0 + b) * nt3;
  const long long c0 = tbase + (long long)blockIdx.x * PC_CELLS;
  const int np = chunk_setup(cstart_p, c0, tbase + nt3, soff);
  const long long p0 = cstart_p[c0];
  const float* Gb = G + (long long)b * M * M * 3 * FP;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const long long c = (long long)blockIdx.x * PC_CELLS + chunk_find(soff, q);
    const int i = (int)(c % g.nt) + 1, j = (int)((c / g.nt) % g.nt) + 1, k = (int)(c / ((long long)g.nt * g.nt)) + 1;
    const long long p = p0 + q;
    const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
    int i1, j1, k1; float ax[2], ay[2], az[2];
    cic_split(fine_tempx(i, xc.x), i1, ax[0], ax[1]);  // idx1 of pm.f90:96 = 0-based kept index
    cic_split(fine_tempx(j, xc.y), j1, ay[0], ay[1]);
    cic_split(fine_tempx(k, xc.z), k1, az[0], az[1]);
    double v0 = dvlut[(unsigned short)vc.x], v1 = dvlut[(unsigned short)vc.y], v2 = dvlut[(unsigned short)vc.z];
    const int qx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, qy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, qz[8] = {0, 0, 0, 1, 1, 1, 0, 1};  // pm.f90:104-111
    float f0[8], f1[8], f2[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float* f = Gb + ((long long)(k1 + qz[t]) * M + (j1 + qy[t])) * 3 * FP + (i1 + qx[t]);
      f0[t] = __ldg(f); f1[t] = __ldg(f + FP); f2[t] = __ldg(f + 2 * FP);
    }
#pragma unroll
    for (int t = 0; t < 8; t++) { f0[t] = kick_prefix(f0[t], a_mid, dt); f1[t] = kick_prefix(f1[t], a_mid, dt); f2[t] = kick_prefix(f2[t], a_mid, dt); }
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float wx = ax[qx[t]], wy = ay[qy[t]], wz = az[qz[t]];
      v0 = __dadd_rn(v0, (double)kick_weight(f0[t], wx, wy, wz));
      v1 = __dadd_rn(v1, (double)kick_weight(f1[t], wx, wy, wz));
      v2 = __dadd_rn(v2, (double)kick_weight(f2[t], wx, wy, wz));
    }
    store_code3(vp, p, vp_encode_lut(v0, S_new, enc), vp_encode_lut(v1, S_new, enc), vp_encode_lut(v2, S_new, enc));
  }
}

// ---------------------------------------------------------------------------------------------
// coarse kick (pm.f90:196-228); Gc(3,0:nc+1,0:nc+1,0:nc+1) = kick_prefix(force_c), vmax over v+vfield (no abs)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PC_T) k_coarse_kick_p(Geom g, const short* __restrict__ xp, short* __restrict__ vp,
                                                       const long long* __restrict__ cstart_p, const float* __restrict__ vfield_p,
                                                       const float* __restrict__ Gc, const double* __restrict__ dvlut,
                                                       const double* __restrict__ enc, double S, unsigned long long* __restrict__ vmax_bits) {
  __shared__ int soff[PC_CELLS + 1];
  const long long c0 = (long long)blockIdx.x * PC_CELLS;
  const int np = chunk_setup(cstart_p, c0, g.ncell_p, soff);
  const long long p0 = cstart_p[c0];
  const int m = g.nc + 2;
  double vm = 0.0;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const long long L = c0 + chunk_find(soff, q);
    int tx, ty, tz, i, j, k;
    phys_decompose(g, L, tx, ty, tz, i, j, k);
    const int X = tx * g.nt + i, Y = ty * g.nt + j, Z = tz * g.nt + k;  // ((itx-1)*nt + (i-1)) of pm.f90:206
    const long long p = p0 + q;
    const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
    int i1, j1, k1; float ax[2], ay[2], az[2];
    cic_split(coarse_tempx(X, xc.x), i1, ax[0], ax[1]);
    cic_split(coarse_tempx(Y, xc.y), j1, ay[0], ay[1]);
    cic_split(coarse_tempx(Z, xc.z), k1, az[0], az[1]);
    double v0 = dvlut[(unsigned short)vc.x], v1 = dvlut[(unsigned short)vc.y], v2 = dvlut[(unsigned short)vc.z];
    const int qx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, qy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, qz[8] = {0, 0, 0, 1, 1, 1, 0, 1};
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float* f = Gc + 3 * (((long long)(k1 + qz[t]) * m + (j1 + qy[t])) * m + (i1 + qx[t]));
      const float wx = ax[qx[t]], wy = ay[qy[t]], wz = az[qz[t]];
      v0 = __dadd_rn(v0, (double)kick_weight(__ldg(f), wx, wy, wz));
      v1 = __dadd_rn(v1, (double)kick_weight(__ldg(f + 1), wx, wy, wz));
      v2 = __dadd_rn(v2, (double)kick_weight(__ldg(f + 2), wx, wy, wz));
    }
    const double vf0 = vfield_p[3 * L], vf1 = vfield_p[3 * L + 1], vf2 = vfield_p[3 * L + 2];
    vm = fmax(vm, fmax(__dadd_rn(v0, vf0), fmax(__dadd_rn(v1, vf1), __dadd_rn(v2, vf2))));  // pm.f90:220
    store_code3(vp, p, vp_encode_lut(v0, S, enc), vp_encode_lut(v1, S, enc), vp_encode_lut(v2, S, enc));
  }
  for (int o = 16; o; o >>= 1) vm = fmax(vm, __shfl_down_sync(0xffffffffu, vm, o));
  if ((threadIdx.x & 31) == 0 && vm > 0.0) atomicMax(vmax_bits, (unsigned long long)__double_as_longlong(vm));
}

// force_c(3,0:nc+1,...) from the three inverse transforms + periodic 1-cell halo (pm.f90:176-189, single image),
// f2_max_coarse = maxval(sum(force_c**2,1)) (pm.f90:192) and the kick prefix, in one pass.
// raw != nullptr additionally stores the unscaled force (diagnostics).
__global__ void __launch_bounds__(256) k_force_c_finish(Geom g, const float* __restrict__ F /*[3][nc][nc][nc+2]*/, float a_mid, float dt,
                                                        float* __restrict__ Gc, float* __restrict__ raw, unsigned* __restrict__ f2max) {
  const int m = g.nc + 2;
  const long long n = (long long)m * m * m;
  const long long vol = (long long)g.nc * g.nc * (g.nc + 2);
  float best = 0.f;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    int x = (int)(q % m) - 1, y = (int)((q / m) % m) - 1, z = (int)(q / ((long long)m * m)) - 1;
    x = (x + g.nc) % g.nc; y = (y + g.nc) % g.nc; z = (z + g.nc) % g.nc;
    const long long o = ((long long)z * g.nc + y) * (g.nc + 2) + x;
    const float f0 = F[o], f1 = F[vol + o], f2 = F[2 * vol + o];
    best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
    Gc[3 * q] = kick_prefix(f0, a_mid, dt); Gc[3 * q + 1] = kick_prefix(f1, a_mid, dt); Gc[3 * q + 2] = kick_prefix(f2, a_mid, dt);
    if (raw) { raw[3 * q] = f0; raw[3 * q + 1] = f1; raw[3 * q + 2] = f2; }
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if ((threadIdx.x & 31) == 0) atomicMax(f2max, __float_as_uint(best));
}
// same for a caller-supplied force_c (diagnostics): in place
__global__ void __launch_bounds__(256) k_force_c_prefix(long long n, float* __restrict__ fc, float a_mid, float dt, unsigned* __restrict__ f2max) {
  float best = 0.f;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const float f0 = fc[3 * q], f1 = fc[3 * q + 1], f2 = fc[3 * q + 2];
    best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
    fc[3 * q] = kick_prefix(f0, a_mid, dt); fc[3 * q + 1] = kick_prefix(f1, a_mid, dt); fc[3 * q + 2] = kick_prefix(f2, a_mid, dt);
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if ((threadIdx.x & 31) == 0) atomicMax(f2max, __float_as_uint(best));
}
}  // namespace cube
