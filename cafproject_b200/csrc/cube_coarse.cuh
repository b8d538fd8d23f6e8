// cube_coarse.cuh -- coarse-mesh force on more than one image: the distributed FFT that replaces
// CUBE/main/pencil_fft.f90 (cube -> x pencils -> y -> z with two `sync all` per slab) and the force_c halo of
// pm.f90:182-189.
//
// Global coarse grid (Gx,Gy,Gz) = nc*(nnx,nny,nnz), R = nnx*nny*nnz images, image rank = icx + nnx*(icy + nny*icz).
//   forward   cube block [nc]^3  --(all-to-all inside the nnx*nny images that share icz)-->  z-slab [sz][Gy][Gx],
//             sz = Gz/R planes, owned by image q = Z/sz;  2-D r2c per plane (cuFFT);
//             --(all-to-all over all images)-->  T[kz][kyl][kx], kyl = Gy/R rows of ky per image;  1-D FFT along kz.
//   k-space   T3[d] = i * kern_c[d] * T / (Gx*Gy*Gz)           (pm.f90:172-175, pencil_fft.f90:58)
//   backward  the mirror image, the three components batched; the last hop sends every image its (nc+2)^2 x planes
//             block INCLUDING the one-cell halo, so the six halo GETs of pm.f90:182-189 need no extra exchange.
// Only real-space results are pinned by the reference; the k-space distribution is free.
#pragma once
#include "cube_common.cuh"

namespace cube {

struct CoarseGeom {
  int R, Gx, Gy, Gz, KX;  // images, global grid, Gx/2+1
  int sz, nyl;            // planes per z-slab, ky rows per image in the transposed layout
  int grp, grp0;          // images per xy group (nnx*nny), first rank of my group
  int nc;
};

// ck(dim) on this image's block of the global lattice (kernel_c.f90:16-72): -r/r^3 with r = ncell*offset, the 4^3
// table (and its mirrored 3-wide far corners) over the eight corners of the GLOBAL lattice when `corrected`
__device__ __forceinline__ int signed_index_n(int gidx, int n) { return (gidx + n / 2) % n - n / 2; }
__global__ void k_kernc_fill_dist(CoarseGeom c, int ox0, int oy0, int oz0, const float* __restrict__ tab /*(3,4,4,4)*/, int dim, int corrected,
                                  float* __restrict__ out /*[nc][nc][nc]*/) {
  const long long n = (long long)c.nc * c.nc * c.nc;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int x = (int)(q % c.nc), y = (int)((q / c.nc) % c.nc), z = (int)(q / ((long long)c.nc * c.nc));
  const int o[3] = {signed_index_n(ox0 + x, c.Gx), signed_index_n(oy0 + y, c.Gy), signed_index_n(oz0 + z, c.Gz)};
  const float rx = 4.f * o[0], ry = 4.f * o[1], rz = 4.f * o[2];
  const float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
  const float rr[3] = {rx, ry, rz};
  float v = (r == 0.f) ? 0.f : __fdiv_rn(-rr[dim], __fmul_rn(__fmul_rn(r, r), r));
  if (corrected && o[0] > -4 && o[0] < 4 && o[1] > -4 && o[1] < 4 && o[2] > -4 && o[2] < 4) {
    v = tab[dim + 3 * (abs(o[0]) + 4 * (abs(o[1]) + 4 * abs(o[2])))];
    if (o[dim] < 0) v = -v;
  }
  out[q] = v;
}

// z-slab from the blocks of the nnx*nny images of my xy group: stage[j][zl][y][x], j = icx' + nnx*icy'
__global__ void __launch_bounds__(256) k_slab_assemble(CoarseGeom c, int nnx, const float* __restrict__ stage, float* __restrict__ slab) {
  const long long n = (long long)c.sz * c.Gy * c.Gx;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(q % c.Gx), Y = (int)((q / c.Gx) % c.Gy), zl = (int)(q / ((long long)c.Gx * c.Gy));
    const int j = X / c.nc + nnx * (Y / c.nc);
    slab[q] = stage[(((long long)j * c.sz + zl) * c.nc + (Y % c.nc)) * c.nc + (X % c.nc)];
  }
}
// slabC[zl][ky][kx] -> pack[q][zl][kyl][kx], ky = q*nyl + kyl
__global__ void __launch_bounds__(256) k_pack_T(CoarseGeom c, const float2* __restrict__ slabC, float2* __restrict__ pack) {
  const long long n = (long long)c.sz * c.Gy * c.KX;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i % c.KX), ky = (int)((i / c.KX) % c.Gy), zl = (int)(i / ((long long)c.KX * c.Gy));
    const int q = ky / c.nyl, kyl = ky - q * c.nyl;
    pack[(((long long)q * c.sz + zl) * c.nyl + kyl) * c.KX + kx] = slabC[i];
  }
}
// pack[q][d][zl][kyl][kx] -> slabC[d][zl][ky][kx]
__global__ void __launch_bounds__(256) k_unpack_T(CoarseGeom c, const float2* __restrict__ pack, float2* __restrict__ slabC) {
  const long long n = 3LL * c.sz * c.Gy * c.KX;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i % c.KX), ky = (int)((i / c.KX) % c.Gy);
    const int zl = (int)((i / ((long long)c.KX * c.Gy)) % c.sz), d = (int)(i / ((long long)c.KX * c.Gy * c.sz));
    const int q = ky / c.nyl, kyl = ky - q * c.nyl;
    slabC[i] = pack[((((long long)q * 3 + d) * c.sz + zl) * c.nyl + kyl) * c.KX + kx];
  }
}
// F_d = i * kern_d * rho_k * scale on the transposed layout (pm.f90:172-175)
__global__ void __launch_bounds__(256) k_green_T(long long nk, const float2* __restrict__ T, const float* __restrict__ kern /*[3][nk]*/, float scale,
                                                 float2* __restrict__ T3 /*[3][nk]*/) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nk) return;
  const float2 c = T[q];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const float k = kern[d * nk + q] * scale;
    T3[d * nk + q] = make_float2(-c.y * k, c.x * k);
  }
}
// LRCKCORR (kernel_c.f90:76-117) on the transposed layout T[kz][kyl][kx]
__global__ void k_kernc_lrck_T(CoarseGeom c, int rank, int dim, const float2* __restrict__ cpure, float* __restrict__ kern) {
  const long long nk = (long long)c.Gz * c.nyl * c.KX;
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nk) return;
  const int x = (int)(q % c.KX), kyl = (int)((q / c.KX) % c.nyl), z = (int)(q / ((long long)c.KX * c.nyl));
  const int n[3] = {c.Gx, c.Gy, c.Gz};
  const float kx[3] = {(float)signed_index_n(x, c.Gx), (float)signed_index_n(rank * c.nyl + kyl, c.Gy), (float)signed_index_n(z, c.Gz)};
  const float kr = sqrtf(kx[0] * kx[0] + kx[1] * kx[1] + kx[2] * kx[2]);
  if (kr > 8.0f || kx[dim] == 0.f) return;
  float ks[3];
  for (int d = 0; d < 3; d++) ks[d] = 2.f * sinf(PI_F * kx[d] / (float)n[d]);
  const float ssum = ks[0] * ks[0] + ks[1] * ks[1] + ks[2] * ks[2];
  kern[q] = kern[q] * 0.25f * PI_F * ks[dim] / ssum / cpure[q].y;
}
__global__ void k_take_imag_scaled(long long nk, const float2* __restrict__ c, float* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nk) out[q] = c[q].y;
}

// slab -> (nc+2)^2 x nplanes block of image (icx,icy) with the periodic one-cell halo in x and y:
// out[d][idx][yy][xx] = slab[d][zloc[idx]][(icy*nc-1+yy) mod Gy][(icx*nc-1+xx) mod Gx]
__global__ void __launch_bounds__(256) k_pack_F(CoarseGeom c, int icx, int icy, int nplanes, const int* __restrict__ zloc,
                                                const float* __restrict__ slab /*[3][sz][Gy][Gx]*/, float* __restrict__ out) {
  const int m = c.nc + 2;
  const long long n = 3LL * nplanes * m * m;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % m), yy = (int)((i / m) % m), idx = (int)((i / ((long long)m * m)) % nplanes), d = (int)(i / ((long long)m * m * nplanes));
    const int X = (icx * c.nc - 1 + xx + c.Gx) % c.Gx, Y = (icy * c.nc - 1 + yy + c.Gy) % c.Gy;
    out[i] = slab[(((long long)d * c.sz + zloc[idx]) * c.Gy + Y) * c.Gx + X];
  }
}
// force_c(3,0:nc+1,0:nc+1,0:nc+1) from the received plane groups: plane zz of component d starts at
// recv[zzoff[zz] + d*zzcs[zz]];  f2_max_coarse (pm.f90:192) and the kick prefix in the same pass
__global__ void __launch_bounds__(256) k_force_c_finish_dist(int nc, const float* __restrict__ recv, const long long* __restrict__ zzoff,
                                                             const long long* __restrict__ zzcs, float a_mid, float dt, float* __restrict__ Gc,
                                                             float* __restrict__ raw, unsigned* __restrict__ f2max) {
  const int m = nc + 2;
  const long long n = (long long)m * m * m;
  float best = 0.f;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const int zz = (int)(q / ((long long)m * m));
    const long long o = zzoff[zz] + (q - (long long)zz * m * m), cs = zzcs[zz];
    const float f0 = recv[o], f1 = recv[o + cs], f2 = recv[o + 2 * cs];
    best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
    Gc[3 * q] = kick_prefix(f0, a_mid, dt); Gc[3 * q + 1] = kick_prefix(f1, a_mid, dt); Gc[3 * q + 2] = kick_prefix(f2, a_mid, dt);
    if (raw) { raw[3 * q] = f0; raw[3 * q + 1] = f1; raw[3 * q + 2] = f2; }
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if ((threadIdx.x & 31) == 0) atomicMax(f2max, __float_as_uint(best));
}

}  // namespace cube
