// cube_fft2d.cuh -- plane-fused passes of the fine-mesh convolution: a 2-CTA thread-block cluster owns one z plane,
// keeps it in (distributed) shared memory between the x and the y transform, so that k-space makes one HBM trip per
// plane instead of two (cube_fft.cuh has the line kernels and the remaining z pass).
//
//   k_fft_xy_fwd   rho[b][z][y][x] (real)  ->  A[b][z][ky][kx]          replaces k_fft_x_fwd + k_fft_y<-1>
//   k_fft_yx_inv   B[d][b][z'][ky][kx]     ->  F[b][z'][y'][d][x'] + f2_max_fine   replaces k_fft_y<+1> + k_fft_x_inv
//                                                                                  + k_f2max_rows
// Shared memory per CTA: col[N][CP] complex = all N rows of HALF the kx columns (CTA h holds kx in [h*K0, ...)),
// a 16-line work buffer s[N][17] for the x transforms and the twiddles: 212 KB at N = 288.  Each CTA x-transforms
// half of the rows and y-transforms half of the columns; the hand-over between the two phases goes through the
// peer's shared memory (DSMEM stores in the forward kernel, DSMEM loads in the inverse one) around one cluster barrier.
#pragma once
#include <cooperative_groups.h>

#include "cube_fft.cuh"

namespace cube {
namespace cg = cooperative_groups;

template <int R1, int R2>
struct Fft2dCfg {
  static constexpr int N = R1 * R2, NH = N / 2 + 1;
  static constexpr int K0 = (NH / 2) & ~1;                 // CTA 0: kx 0..K0-1, CTA 1: kx K0..NH-1 (even split point: 16-byte copies)
  static constexpr int CP = ((NH - K0) + 1) & ~1;          // column-buffer pitch (complex), even
  static constexpr int NG = (NH - K0 + FL - 1) / FL;       // 16-column groups per CTA
  static constexpr int LWX = FL + 1;
  static constexpr int NT = FL * (R1 > R2 ? R1 : R2);
  static constexpr size_t SMEM = (size_t)(N * CP + N * LWX + N) * sizeof(float2);
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
template <int NKEEP_MAX, int I>
__device__ __forceinline__ void cp_async_wait_dyn(int keep) {  // wait_group needs an immediate
  if constexpr (I <= NKEEP_MAX) {
    if (keep == I) cp_async_wait<I>();
    else cp_async_wait_dyn<NKEEP_MAX, I + 1>(keep);
  } else cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// forward: grid = (2, N, nbatch), cluster (2,1,1)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FL * (R1 > R2 ? R1 : R2))
    k_fft_xy_fwd(FftGeom g, const float* __restrict__ rho, float2* __restrict__ A, const float2* __restrict__ tw_g) {
  using C = Fft2dCfg<R1, R2>;
  constexpr int N = C::N, NH = C::NH, K0 = C::K0, CP = C::CP, LW = C::LWX, NT = C::NT;
  extern __shared__ float2 smem[];
  float2* col = smem;                 // [N][CP]
  float2* s = smem + N * CP;          // [N][LW]
  float2* tw = s + N * LW;            // [N]
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  float2* const colpeer = cluster.map_shared_rank(col, c ^ 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, line = tid % FL, idx = tid / FL;
  const int z = blockIdx.y, b = blockIdx.z;
  load_tw(tw, tw_g, N);
  cluster.sync();  // the peer CTA is running: its shared memory may be written
  // ---- x phase: my half of the rows, 32 real rows (16 complex lines) per batch
  const int rows_half = (N / 2 + 1) & ~1;                       // even number of rows for CTA 0
  const int y_lo = c ? rows_half : 0, y_hi = c ? N : rows_half;
  const float* src = rho + ((size_t)b * N + z) * (size_t)N * N;
  for (int y0 = y_lo; y0 < y_hi; y0 += 32) {
    __syncthreads();  // work buffer free (previous batch's scatter done); also orders the twiddle fill
    for (int r = warp; r < 32; r += NT / 32) {
      const int y = y0 + r, l = r >> 1, im = r & 1;
      float* dst = reinterpret_cast<float*>(s) + im;
      if (y < y_hi) {
        const float* row = src + (size_t)y * N;
        for (int x = lane; x < N; x += 32) dst[(x * LW + l) * 2] = row[x];
      } else {
        for (int x = lane; x < N; x += 32) dst[(x * LW + l) * 2] = 0.f;
      }
    }
    __syncthreads();
    fft_step_a<R1, R2, -1, LW>(s, tw, line, idx);
    __syncthreads();
    float2 v[R2];
    if (idx < R1) fft_step_b<R1, R2, -1, LW>(s, line, idx, v);
    __syncthreads();
    if (idx < R1) {
#pragma unroll
      for (int k2 = 0; k2 < R2; k2++) s[(idx + R1 * k2) * LW + line] = v[k2];
    }
    __syncthreads();
    // separate the two real rows (Xa[k] = (Z[k] + conj Z[N-k])/2, Xb[k] = (Z[k] - conj Z[N-k])/(2i)) and hand every
    // kx to the CTA that owns its column
    for (int r = warp; r < 32; r += NT / 32) {
      const int y = y0 + r;
      if (y >= y_hi) continue;
      const int l = r >> 1, im = r & 1;
      for (int k = lane; k < NH; k += 32) {
        const float2 zk = s[k * LW + l], zn = s[(k ? N - k : 0) * LW + l];
        float2 o;
        if (!im) o = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
        else o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
        const int h = k >= K0;
        (h == c ? col : colpeer)[y * CP + (k - h * K0)] = o;
      }
    }
  }
  cluster.sync();  // both halves of the rows are in my columns
  // ---- y phase: my columns, 16 per group, in place; results go straight to A
  const int ncol = c ? NH - K0 : K0;
  float2* dstA = A + ((size_t)b * N + z) * (size_t)N * g.P + (c ? K0 : 0);
  for (int gq = 0; gq * FL < ncol; gq++) {
    float2* sc = col + gq * FL;
    const bool act = gq * FL + line < ncol;
    if (act) fft_step_a<R1, R2, -1, CP>(sc, tw, line, idx);
    __syncthreads();
    if (act && idx < R1) {
      float2 v[R2];
      fft_step_b<R1, R2, -1, CP>(sc, line, idx, v);
      float2* o = dstA + gq * FL + line;
#pragma unroll
      for (int k2 = 0; k2 < R2; k2++) o[(size_t)(idx + R1 * k2) * g.P] = v[k2];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// inverse: grid = (2, M, nbatch), cluster (2,1,1); the three force components of one plane are done in turn so that
// f2_max_fine = maxval(sum(force_f**2,1)) (pm.f90:85) is taken here (components 0,1 re-read through L2)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FL * (R1 > R2 ? R1 : R2))
    k_fft_yx_inv(FftGeom g, const float2* __restrict__ B, float* __restrict__ F, unsigned* __restrict__ f2max, const float2* __restrict__ tw_g) {
  using C = Fft2dCfg<R1, R2>;
  constexpr int N = C::N, NH = C::NH, K0 = C::K0, CP = C::CP, LW = C::LWX, NT = C::NT, NG = C::NG;
  extern __shared__ float2 smem[];
  float2* col = smem;
  float2* s = smem + N * CP;
  float2* tw = s + N * LW;
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const float2* const colpeer = cluster.map_shared_rank(col, c ^ 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, line = tid % FL, idx = tid / FL;
  const int zp = blockIdx.y, b = blockIdx.z;
  const int ncol = c ? NH - K0 : K0, k0 = c ? K0 : 0;
  const int npair = (g.M + 1) / 2, p_half = (npair + 1) / 2;
  const int p_lo = c ? p_half : 0, p_hi = c ? npair : p_half;
  load_tw(tw, tw_g, N);
  cluster.sync();  // the peer CTA is running: its shared memory may be read
  float best = 0.f;
  for (int d = 0; d < 3; d++) {
    // ---- load my columns of the plane, one commit group per 16-column group
    const float2* src = B + (((size_t)d * g.nbatch + b) * g.M + zp) * (size_t)N * g.P + k0;
#pragma unroll
    for (int gq = 0; gq < NG; gq++) {
      const int cbase = gq * FL;
      for (int e = tid; e < N * (FL / 2); e += NT) {
        const int n = e / (FL / 2), cc = cbase + 2 * (e - n * (FL / 2));
        if (cc + 1 < ncol) cp_async16(col + n * CP + cc, src + (size_t)n * g.P + cc);
        else if (cc < ncol) cp_async8(col + n * CP + cc, src + (size_t)n * g.P + cc);
      }
      cp_async_commit();
    }
    // ---- y inverse, in place (natural order back into the column buffer)
#pragma unroll
    for (int gq = 0; gq < NG; gq++) {
      cp_async_wait_dyn<NG - 1, 0>(NG - 1 - gq);
      __syncthreads();
      float2* sc = col + gq * FL;
      const bool act = gq * FL + line < ncol;
      if (act) fft_step_a<R1, R2, +1, CP>(sc, tw, line, idx);
      __syncthreads();
      float2 v[R2];
      if (act && idx < R1) fft_step_b<R1, R2, +1, CP>(sc, line, idx, v);
      __syncthreads();
      if (act && idx < R1) {
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) sc[(idx + R1 * k2) * CP + line] = v[k2];
      }
    }
    cluster.sync();  // all columns of the plane are y-transformed (mine here, the rest in the peer's shared memory)
    // ---- x inverse (c2r), my half of the kept rows, 16 row pairs per batch
    float* dstF = F + (((size_t)b * g.M + zp) * g.M) * 3 * (size_t)g.FP;
    for (int pb = p_lo; pb < p_hi; pb += FL) {
      // Z[k] = Xa[k] + i Xb[k], Z[N-k] = conj(Xa[k]) + i conj(Xb[k])
      for (int e = tid; e < FL * NH; e += NT) {
        const int l = e / NH, k = e - l * NH;
        const int p = pb + l, ya = 2 * p, yb = ya + 1;
        float2 a = make_float2(0.f, 0.f), cc = a;
        const int h = k >= K0;
        const float2* ch = (h == c ? col : colpeer) + (k - h * K0);
        if (p < p_hi) {
          a = ch[(ya + g.off) * CP];
          if (yb < g.M) cc = ch[(yb + g.off) * CP];
        }
        s[k * LW + l] = make_float2(a.x - cc.y, a.y + cc.x);
        if (k && 2 * k != N) s[(N - k) * LW + l] = make_float2(a.x + cc.y, cc.x - a.y);
      }
      __syncthreads();
      fft_step_a<R1, R2, +1, LW>(s, tw, line, idx);
      __syncthreads();
      float2 v[R2];
      if (idx < R1) fft_step_b<R1, R2, +1, LW>(s, line, idx, v);
      __syncthreads();
      if (idx < R1) {
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) s[(idx + R1 * k2) * LW + line] = v[k2];
      }
      __syncthreads();
      for (int r = warp; r < 2 * FL; r += NT / 32) {
        const int l = r >> 1, im = r & 1, p = pb + l, yp = 2 * p + im;
        if (p >= p_hi || yp >= g.M) continue;
        float* row = dstF + ((size_t)yp * 3 + d) * g.FP;
        const float* sp = reinterpret_cast<const float*>(s) + im;
        if (d < 2) {
#pragma unroll 4
          for (int x = lane; x < g.M; x += 32) row[x] = sp[((x + g.off) * LW + l) * 2];
        } else {
#pragma unroll 4
          for (int x = lane; x < g.M; x += 32) {
            const float f2 = sp[((x + g.off) * LW + l) * 2];
            row[x] = f2;
            const float f0 = __ldcg(row - 2 * g.FP + x), f1 = __ldcg(row - g.FP + x);  // written by this thread for d = 0, 1
            best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
          }
        }
      }
      __syncthreads();  // work buffer free for the next batch
    }
    cluster.sync();  // the peer has finished reading my columns: they may be overwritten (or the CTA may exit)
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if (lane == 0) atomicMax(&f2max[b], __float_as_uint(best));
}

}  // namespace cube
