// cube_fft.cuh -- hand-written fine-mesh force convolution for sm_100a (replaces FFTW's
// plan_fft_fine / plan_ifft_fine + the kern_f multiply of CUBE/main/pm.f90:75-84).
//
// The reference transforms the whole padded tile rho_f(nfe,nfe,nfe), nfe = nft + 2*nfb (304 for nt = 64),
// multiplies by kern_f and transforms back three times, but it only keeps force_f on the M = nft+2 points
// nfb..nfe-nfb+1 (pm.f90:83), and the real-space kernel has support |offset| <= nf_cutoff-1 = 15
// (kernel_f.f90:32-38).  A circular convolution of length N >= M + 30 on the window that starts 15 cells
// before the kept region therefore gives the same forces (no wrap-around reaches the kept points); for
// nt = 64 that is N = 288 = 16*18 instead of 304 = 16*19: 15% fewer cells and radix-friendly.
//
// Pipeline per batch of tiles (all arrays in HBM, x fastest):
//   rho [b][z][y][x]          real   N^3                       (fine CIC deposit)
//   A   [b][z][ky][kx]        complex, kx pitch P >= N/2+1      k_fft_x_fwd (r2c, two real rows per complex line)
//                                                              k_fft_y (in place)
//   B   [d][b][z'][ky][kx]    complex, z' = M kept planes      k_fft_z_green: z forward, x i*K_d, z inverse, d=1..3
//                                                              k_fft_y (inverse, in place, keeps M rows)
//   F   [b][z'][y'][d][x']    real, x' pitch FP                k_fft_x_inv3 (c2r, 3 components + f2_max) -> read by the fine kick
// Every kernel moves 16 lines x N points through shared memory in the layout s[n][line] (line fastest):
// a warp then touches 32 consecutive float2 per access (conflict-free) in both Cooley-Tukey steps
// N = R1*R2, each thread doing one R-point DFT in registers with compile-time twiddles.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <type_traits>
#include "cube_common.cuh"

namespace cube {

// ---------------------------------------------------------------------------------------------
// compile-time trigonometry: cos/sin(2 pi a / R) to double precision
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr double ct_sin_small(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int n = 1; n <= 12; n++) { term *= -x2 / (double)((2 * n) * (2 * n + 1)); sum += term; }
  return sum;
}
__host__ __device__ constexpr double ct_cos_small(double x) {
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int n = 1; n <= 12; n++) { term *= -x2 / (double)((2 * n - 1) * (2 * n)); sum += term; }
  return sum;
}
struct ct_cs { double c, s; };
__host__ __device__ constexpr ct_cs ct_cossin_turn(int a, int R) {  // angle = 2 pi a / R
  a %= R; if (a < 0) a += R;
  const int q = (8 * a) / R, r = 8 * a - q * R;  // octant and remainder: angle = (pi/4)(q + r/R)
  const double qp = 0.785398163397448309615660845819875721;
  const double t = qp * (double)r / (double)R, u = qp * (double)(R - r) / (double)R;  // theta', pi/4 - theta'
  switch (q) {
    case 0: return {ct_cos_small(t), ct_sin_small(t)};
    case 1: return {ct_sin_small(u), ct_cos_small(u)};
    case 2: return {-ct_sin_small(t), ct_cos_small(t)};
    case 3: return {-ct_cos_small(u), ct_sin_small(u)};
    case 4: return {-ct_cos_small(t), -ct_sin_small(t)};
    case 5: return {-ct_sin_small(u), -ct_cos_small(u)};
    case 6: return {ct_sin_small(t), -ct_cos_small(t)};
    default: return {ct_cos_small(u), -ct_sin_small(u)};
  }
}

template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Packed complex arithmetic.  sm_100a has two-wide f32 instructions (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2 on a
// 64-bit register pair): a complex add is ONE instruction instead of two, a multiply by a twiddle TWO (FMUL2 + FFMA2) instead of
// four, and ptxas folds the (re,im) swap, per-half negation and splat of an operand into the instruction's own operand modifiers
// (R.F32x2.LO_HI.NP ...), so multiplying by +-i costs nothing extra.  The line-FFT kernels are bound by instruction issue, most of
// it this arithmetic (profiles/r02_notes.md).  Each half is an IEEE f32 operation: same numerics as the scalar form.
// ---------------------------------------------------------------------------------------------
typedef unsigned long long cpk;
__device__ __forceinline__ cpk c_pack(float x, float y) { cpk r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ cpk c_pack(float2 a) { return c_pack(a.x, a.y); }
__device__ __forceinline__ float2 c_unpack(cpk v) { float2 a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(v)); return a; }
// the four two-wide operations, packed (Pk) or as pairs of scalar instructions (Sc: same values; the z pass is faster with it --
// measured 6.8 against 7.8 ms at cfg 2, its twiddle constants then ride as immediates of scalar FFMAs -- every other pass with Pk)
struct Pk {
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { cpk r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_pack(a)), "l"(c_pack(b))); return c_unpack(r); }
  static __device__ __forceinline__ float2 sub(float2 a, float2 b) { cpk r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_pack(a)), "l"(c_pack(b))); return c_unpack(r); }
  static __device__ __forceinline__ float2 mul2(float2 a, float2 b) { cpk r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c_pack(a)), "l"(c_pack(b))); return c_unpack(r); }
  static __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    cpk r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(c_pack(a)), "l"(c_pack(b)), "l"(c_pack(c))); return c_unpack(r);
  }
};
struct Sc {
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
  static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
  static __device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
  static __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
};
struct Mx {  // packed additions, scalar multiplies (constants stay immediates)
  static __device__ __forceinline__ float2 add(float2 a, float2 b) { return Pk::add(a, b); }
  static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return Pk::sub(a, b); }
  static __device__ __forceinline__ float2 mul2(float2 a, float2 b) { return Sc::mul2(a, b); }
  static __device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return Sc::fma2(a, b, c); }
};
template <class M = Pk> __device__ __forceinline__ float2 c_add(float2 a, float2 b) { return M::add(a, b); }
template <class M = Pk> __device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return M::sub(a, b); }
template <class M = Pk> __device__ __forceinline__ float2 c_mul2(float2 a, float2 b) { return M::mul2(a, b); }
template <class M = Pk> __device__ __forceinline__ float2 c_fma2(float2 a, float2 b, float2 c) { return M::fma2(a, b, c); }
// a + S*i*b  (S = +1 or -1): (a.x - S b.y, a.y + S b.x)
template <int S, class M = Pk> __device__ __forceinline__ float2 c_add_i(float2 a, float2 b) {
  return M::fma2(make_float2(b.y, b.x), make_float2(S > 0 ? -1.f : 1.f, S > 0 ? 1.f : -1.f), a);
}
// a * (c + i s)
template <class M = Pk> __device__ __forceinline__ float2 c_mul(float2 a, float c, float s) {
  return M::fma2(make_float2(a.y, a.x), make_float2(-s, s), M::mul2(a, make_float2(c, c)));
}

// a * exp(DIR * 2 pi i T / R) with the root folded to immediates; trivial roots cost nothing
template <int DIR, int T, int R, class M = Pk>
__device__ __forceinline__ float2 mul_root(float2 a) {
  constexpr int t = ((T % R) + R) % R;
  if constexpr (t == 0) return a;
  else if constexpr (2 * t == R) return make_float2(-a.x, -a.y);
  else if constexpr (4 * t == R) return DIR > 0 ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);      // * (+-i)
  else if constexpr (4 * t == 3 * R) return DIR > 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);  // * (-+i)
  else {
    constexpr ct_cs w = ct_cossin_turn(t, R);
    constexpr float c = (float)w.c, s = (float)(DIR > 0 ? w.s : -w.s);
    return c_mul<M>(a, c, s);
  }
}

__host__ __device__ constexpr int fft_radix_of(int R) { return R % 4 == 0 ? 4 : R % 2 == 0 ? 2 : R % 3 == 0 ? 3 : R % 5 == 0 ? 5 : R; }

// p-point butterfly, natural order in and out, root exp(DIR*2 pi i/p)
template <int P, int DIR, class M = Pk>
__device__ __forceinline__ void butterfly(float2 (&t)[P]) {
  if constexpr (P == 2) {
    const float2 a = t[0], b = t[1];
    t[0] = c_add<M>(a, b);
    t[1] = c_sub<M>(a, b);
  } else if constexpr (P == 4) {
    const float2 a = c_add<M>(t[0], t[2]), b = c_sub<M>(t[0], t[2]);
    const float2 c = c_add<M>(t[1], t[3]), d = c_sub<M>(t[1], t[3]);
    t[0] = c_add<M>(a, c);
    t[2] = c_sub<M>(a, c);
    t[1] = c_add_i<(DIR > 0 ? 1 : -1), M>(b, d);   // b +- i d
    t[3] = c_add_i<(DIR > 0 ? -1 : 1), M>(b, d);   // b -+ i d
  } else if constexpr (P == 3) {
    constexpr float s0 = (float)(DIR > 0 ? 0.866025403784438646763723170752936183 : -0.866025403784438646763723170752936183);
    const float2 s = c_add<M>(t[1], t[2]), d = c_sub<M>(t[1], t[2]);
    const float2 m = c_fma2<M>(s, make_float2(-0.5f, -0.5f), t[0]);
    t[0] = c_add<M>(t[0], s);
    t[1] = c_fma2<M>(make_float2(d.y, d.x), make_float2(-s0, s0), m);   // m + i s0 d
    t[2] = c_fma2<M>(make_float2(d.y, d.x), make_float2(s0, -s0), m);   // m - i s0 d
  } else {  // generic small prime (5): direct DFT with immediate roots
    float2 o[P];
    static_for<0, P>([&](auto Q) {
      float2 acc = t[0];
      static_for<1, P>([&](auto J) { acc = c_add<M>(acc, mul_root<DIR, (J.value * Q.value) % P, P, M>(t[J.value])); });
      o[Q.value] = acc;
    });
    static_for<0, P>([&](auto Q) { t[Q.value] = o[Q.value]; });
  }
}

// In-register DFT of the R elements x[OFF + S*i] (natural order in and out), decimation in time.
template <int R, int DIR, int S, int OFF, int NT, class M = Pk>
__device__ __forceinline__ void dft_rec(float2 (&x)[NT]) {
  if constexpr (R > 1) {
    constexpr int p = fft_radix_of(R), m = R / p;
    static_for<0, p>([&](auto J) { dft_rec<m, DIR, S * p, OFF + S * J.value, NT, M>(x); });
    float2 y[R];
    static_for<0, m>([&](auto K) {
      float2 t[p];
      static_for<0, p>([&](auto J) { t[J.value] = mul_root<DIR, J.value * K.value, R, M>(x[OFF + S * (J.value + p * K.value)]); });
      butterfly<p, DIR, M>(t);
      static_for<0, p>([&](auto Q) { y[K.value + m * Q.value] = t[Q.value]; });
    });
    static_for<0, R>([&](auto I) { x[OFF + S * I.value] = y[I.value]; });
  }
}
template <int R, int DIR, class M = Pk>
__device__ __forceinline__ void dft(float2 (&x)[R]) { dft_rec<R, DIR, 1, 0, R, M>(x); }

// a * tw or a * conj(tw)
template <int DIR, class M = Pk>
__device__ __forceinline__ float2 mul_tw(float2 a, float2 w) { return c_mul<M>(a, w.x, DIR < 0 ? w.y : -w.y); }

// ---------------------------------------------------------------------------------------------
// CTA-level line FFT, N = R1*R2, LW lines interleaved: s[n*LW + line].  tw[t] = exp(-2 pi i t/N).
// thread -> (line = tid % 16, idx = tid / 16); blockDim.x = 16*max(R1,R2).
// ---------------------------------------------------------------------------------------------
constexpr int FL = 16;  // lines per CTA

// step A: R1-point DFTs over n1 for fixed n2 = idx, twiddle, back to the same slots (holds Y[k1][n2])
template <int R1, int R2, int DIR, int LW, class M = Pk>
__device__ __forceinline__ void fft_step_a(float2* s, const float2* tw, int line, int idx) {
  if (idx < R2) {
    float2 v[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; n1++) v[n1] = s[(n1 * R2 + idx) * LW + line];
    dft<R1, DIR, M>(v);
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) s[(k1 * R2 + idx) * LW + line] = k1 ? mul_tw<DIR, M>(v[k1], tw[idx * k1]) : v[0];
  }
}
// step B: R2-point DFT over n2 for fixed k1 = idx; v[k2] = X[k1 + R1*k2]
template <int R1, int R2, int DIR, int LW, class M = Pk>
__device__ __forceinline__ void fft_step_b(const float2* s, int line, int idx, float2 (&v)[R2]) {
#pragma unroll
  for (int n2 = 0; n2 < R2; n2++) v[n2] = s[(idx * R2 + n2) * LW + line];
  dft<R2, DIR, M>(v);
}
// mirrored inverse, first half: from v[k2] = X[k1 + R1*k2] (registers of thread k1 = idx) to s[(k1*R2+n2)]
template <int R1, int R2, int LW, class M = Pk>
__device__ __forceinline__ void ifft_step_a(float2 (&v)[R2], float2* s, const float2* tw, int line, int idx) {
  dft<R2, +1, M>(v);
#pragma unroll
  for (int n2 = 0; n2 < R2; n2++) s[(idx * R2 + n2) * LW + line] = n2 ? mul_tw<+1, M>(v[n2], tw[n2 * idx]) : v[0];
}
// mirrored inverse, second half: thread n2 = idx gets v[n1] = x[n1*R2 + n2]
template <int R1, int R2, int LW, class M = Pk>
__device__ __forceinline__ void ifft_step_b(const float2* s, int line, int idx, float2 (&v)[R1]) {
#pragma unroll
  for (int k1 = 0; k1 < R1; k1++) v[k1] = s[(k1 * R2 + idx) * LW + line];
  dft<R1, +1, M>(v);
}

struct FftGeom {
  int N;     // transform length
  int NH;    // N/2+1
  int P;     // kx pitch of the complex arrays (multiple of 16)
  int M;     // kept points per dim (nft+2)
  int off;   // first kept point on the FFT grid (15)
  int FP;    // x' pitch of the force rows
  int nbatch;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int NKEEP> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(NKEEP) : "memory"); }
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void load_tw(float2* tw_s, const float2* __restrict__ tw_g, int N) {
  for (int t = threadIdx.x; t < N; t += blockDim.x) tw_s[t] = tw_g[t];
}

// Where tile b of a batch finds its N^3 window of fine density: inside one region grid shared by the batch (tvol == 0; the
// tiles' windows overlap there, cube_kernels.cuh) or as the b-th of nb separate windows.
struct RhoView {
  const float* p; long long ldy, ldz, tvol;
  int t0[3];  // first tile of the batch's box
  int nnt, tile0, tstep /* 4 nt */;
};
__device__ __forceinline__ const float* rho_window(const RhoView& v, int b) {
  if (v.tvol) return v.p + (size_t)b * v.tvol;
  const int t = v.tile0 + b;
  const int tx = t % v.nnt - v.t0[0], ty = (t / v.nnt) % v.nnt - v.t0[1], tz = t / (v.nnt * v.nnt) - v.t0[2];
  return v.p + ((size_t)tz * v.ldz + (size_t)ty * v.ldy + tx) * v.tstep;
}

// ---------------------------------------------------------------------------------------------
// x forward (r2c): window of rho [z][y][x] -> A[b][z][y][kx].  One CTA = 32 real rows = 16 complex lines.
// grid = (ceil(N/32), N, nbatch)
// ---------------------------------------------------------------------------------------------
template <int R1, int R2>
__global__ void __launch_bounds__(FL * (R1 > R2 ? R1 : R2)) k_fft_x_fwd(FftGeom g, RhoView rho, float2* __restrict__ A,
                                                                       const float2* __restrict__ tw_g) {
  constexpr int N = R1 * R2, LW = FL + 1, NT = FL * (R1 > R2 ? R1 : R2);
  extern __shared__ float2 smem[];
  float2* s = smem;            // [N][LW]
  float2* tw = smem + N * LW;  // [N]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int y0 = blockIdx.x * 32, z = blockIdx.y, b = blockIdx.z;
  load_tw(tw, tw_g, N);
  const float* src = rho_window(rho, b) + (size_t)z * rho.ldz;
  // rows y0+2l (re), y0+2l+1 (im); lanes along x
  for (int r = warp; r < 32; r += NT / 32) {
    const int y = y0 + r;
    const int l = r >> 1, im = r & 1;
    float* dst = reinterpret_cast<float*>(s) + im;
    if (y < N) {  // asynchronous copies: every load of the CTA is in flight at once (the pass was bound by global-load latency)
      const float* row = src + (size_t)y * rho.ldy;
      for (int x = lane; x < N; x += 32) cp_async4(&dst[(x * LW + l) * 2], row + x);
    } else {
      for (int x = lane; x < N; x += 32) dst[(x * LW + l) * 2] = 0.f;
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  const int line = tid % FL, idx = tid / FL;
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
  // separate the two real rows: Xa[k] = (Z[k] + conj Z[N-k])/2, Xb[k] = (Z[k] - conj Z[N-k])/(2i)
  float2* dstA = A + ((size_t)b * N + z) * (size_t)N * g.P;
  for (int r = warp; r < 32; r += NT / 32) {
    const int y = y0 + r;
    if (y >= N) continue;
    const int l = r >> 1, im = r & 1;
    float2* row = dstA + (size_t)y * g.P;
    for (int k = lane; k < g.NH; k += 32) {
      const float2 zk = s[k * LW + l], zn = s[(k ? N - k : 0) * LW + l];
      float2 o;
      if (!im) o = c_mul2(c_add(zk, make_float2(zn.x, -zn.y)), make_float2(0.5f, 0.5f));  // (Z[k] + conj Z[N-k]) / 2
      else { const float2 d = c_sub(zk, make_float2(zn.x, -zn.y)); o = c_mul2(make_float2(d.y, d.x), make_float2(0.5f, -0.5f)); }  // (Z[k] - conj Z[N-k]) / 2i
      row[k] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// y transform, in place, lines along ky for 16 adjacent kx.  DIR=-1: A (all rows, N planes);
// DIR=+1: B (writes only the M kept rows y' -> row index y'+off stays in place), planes = M per (d,b).
// grid = (P/16, planes, nslab) ; plane stride = N*P, slab stride = planes*N*P
// ---------------------------------------------------------------------------------------------
template <int R1, int R2, int DIR, int MINB = 0>
__global__ void __launch_bounds__(FL * (R1 > R2 ? R1 : R2), MINB) k_fft_y(FftGeom g, float2* __restrict__ A, const float2* __restrict__ tw_g) {
  constexpr int N = R1 * R2, LW = FL, NT = FL * (R1 > R2 ? R1 : R2);
  extern __shared__ float2 smem[];
  float2* s = smem;
  float2* tw = smem + N * LW;
  const int tid = threadIdx.x, line = tid % FL, idx = tid / FL;
  const int kx = blockIdx.x * FL + line;
  load_tw(tw, tw_g, N);
  float2* base = A + ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * (size_t)N * g.P + kx;
  const bool act = kx < g.NH;
  {  // all N x 16 elements as 16-byte asynchronous copies issued at once (rows of A are 16-byte aligned: P, kx0 multiples of 16)
    const int kx0 = blockIdx.x * FL;
    const float2* src = base - line;
    for (int e = tid; e < N * (FL / 2); e += NT) {
      const int n = e / (FL / 2), pr = e - n * (FL / 2);
      if (kx0 + 2 * pr < g.NH) cp_async16(s + n * LW + 2 * pr, src + (size_t)n * g.P + 2 * pr);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
  if (act) fft_step_a<R1, R2, DIR, LW>(s, tw, line, idx);
  __syncthreads();
  if (act && idx < R1) {
    float2 v[R2];
    fft_step_b<R1, R2, DIR, LW>(s, line, idx, v);
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) {
      const int y = idx + R1 * k2;
      if (DIR < 0 || (y >= g.off && y < g.off + g.M)) base[(size_t)y * g.P] = v[k2];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// z forward + Green multiply + z inverse for the three force components.
// A[b][z][ky][kx] -> B[d][b][z'][ky][kx];  kern[d][kz][ky][kx] real, multiplied by `scale` (1/N^3) on load.
// out = i*K*c  (pm.f90:79-80: re' = -im*K, im' = re*K).   grid = (P/16, N); the CTA loops over the batch with the
// next tile's 16 lines prefetched (cp.async, 16-byte copies) into the other half of a double buffer while the
// current tile is transformed.  K is kept in shared memory for kz <= N/2 only: K_d(N-kz) = +-K_d(kz) (odd in its
// own axis, even in the others: kernel_f.f90:32-38 mirrors the table that way).
// ---------------------------------------------------------------------------------------------

// NB = 2: the next tile's lines are prefetched into the other half of a double buffer; NB = 1: one buffer, so that three CTAs fit an
// SM (N <= 320) and cover each other's loads instead
template <int R1, int R2, class ZM = Sc, int NB = 2>
__global__ void __launch_bounds__(FL * (R1 > R2 ? R1 : R2), (R1 * R2 <= 320 ? (NB == 1 ? 3 : 2) : 1)) k_fft_z_green(FftGeom g, const float2* __restrict__ A, float2* __restrict__ B,
                                                                            const float* __restrict__ kern, float scale,
                                                                            const float2* __restrict__ tw_g) {
  constexpr int N = R1 * R2, LW = FL, NT = FL * (R1 > R2 ? R1 : R2), NHZ = N / 2 + 1;
  extern __shared__ float2 smem[];
  float2* sbuf = smem;                                           // [NB][N][16]
  float2* tw = smem + NB * N * LW;                               // [N]
  float* ks = reinterpret_cast<float*>(smem + NB * N * LW + N);  // [3][NHZ][16]
  const int tid = threadIdx.x, line = tid % FL, idx = tid / FL;
  const int kx0 = blockIdx.x * FL, kx = kx0 + line, ky = blockIdx.y;
  const bool act = kx < g.NH;
  const size_t plane = (size_t)N * g.P;
  const size_t colo = (size_t)ky * g.P + kx;
  // prefetch: thread -> (n, pair of lines); rows of A are 16-byte aligned (P and kx0 are multiples of 16)
  auto prefetch = [&](int b, float2* dst) {
    const float2* src = A + (size_t)b * N * plane + (size_t)ky * g.P + kx0;
    for (int e = tid; e < N * (FL / 2); e += NT) {
      const int n = e / (FL / 2), pr = e - n * (FL / 2);
      if (kx0 + 2 * pr < g.NH) cp_async16(dst + n * LW + 2 * pr, src + (size_t)n * plane + 2 * pr);
    }
    cp_async_commit();
  };
  if (NB == 2) prefetch(0, sbuf);
  load_tw(tw, tw_g, N);
  for (int q = idx; q < 3 * NHZ; q += NT / FL) {
    const int d = q / NHZ, kz = q - d * NHZ;
    ks[q * FL + line] = act ? kern[((size_t)(d * N + kz) * N + ky) * g.P + kx] * scale : 0.f;
  }
  for (int b = 0; b < g.nbatch; b++) {
    float2* s = sbuf + (NB == 2 ? (b & 1) : 0) * N * LW;
    __syncthreads();  // the other buffer's readers (tile b-1) are done; also orders the ks/tw fill
    if (NB == 2) {
      if (b + 1 < g.nbatch) { prefetch(b + 1, sbuf + ((b + 1) & 1) * N * LW); cp_async_wait<1>(); }
      else cp_async_wait<0>();
    } else {
      prefetch(b, s);
      cp_async_wait<0>();
    }
    __syncthreads();
    if (act) fft_step_a<R1, R2, -1, LW, ZM>(s, tw, line, idx);
    __syncthreads();
    float2 X[R2];
    if (act && idx < R1) fft_step_b<R1, R2, -1, LW, ZM>(s, line, idx, X);
#pragma unroll 1
    for (int d = 0; d < 3; d++) {
      __syncthreads();
      if (act && idx < R1) {
        float2 w[R2];
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) {
          const int kz = idx + R1 * k2;
          const int kf = kz < NHZ ? kz : N - kz;
          float K = ks[(d * NHZ + kf) * FL + line];
          if (d == 2 && kz >= NHZ) K = -K;
          w[k2] = make_float2(-X[k2].y * K, X[k2].x * K);  // i K X
        }
        ifft_step_a<R1, R2, LW, ZM>(w, s, tw, line, idx);
      }
      __syncthreads();
      if (act && idx < R2) {
        float2 v[R1];
        ifft_step_b<R1, R2, LW, ZM>(s, line, idx, v);
        float2* dst = B + ((size_t)d * g.nbatch + b) * (size_t)g.M * plane + colo;
#pragma unroll
        for (int n1 = 0; n1 < R1; n1++) {
          const int zi = n1 * R2 + idx - g.off;
          if (zi >= 0 && zi < g.M) dst[(size_t)zi * plane] = v[n1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// x inverse (c2r) of the THREE force components in one CTA + f2_max_fine (pm.f90:83-85).
// B[d][b][z'][y][kx] -> F[b][z'][y'][d][x'].  One CTA = 16 kept rows x 3 components = 24 complex lines (two real rows per
// complex line); grid = (ceil(M/16), M, nbatch), 24*max(R1,R2) threads.  Having the three components of a mesh node in
// one CTA lets the epilogue take |force|^2 for f2_max_fine without another pass over F (the separate pass cost 1.9 ms at
// cfg 2), and the epilogue reads a transposed copy T[line][x] of the result: one 8-byte shared load yields both rows of a
// line, every global store is a full 128-byte row segment.  (The one-component-per-CTA kernel it replaces spent half of
// its instructions in its epilogue and was issue-bound: ncu 82 % issue-active, profiles/r01h_ncu_full_cfg1.csv.)
// The kick's per-node prefix a_mid*dt/6/pi (pm.f90:104) is folded into the Green multiply of k_fft_z_green by the caller,
// so the values stored here are already what the kick gathers.
// Work layout s[n*25 + (n/R2)*6 + line]: odd line pitch keeps the Hermitian unpack (lanes along k) conflict-free, the
// +6 per R2 rows keeps step B's half-warps that straddle two idx values on disjoint banks.
// ---------------------------------------------------------------------------------------------
constexpr int X3L = 24, X3LW = 25, X3PAD = 6;
template <int R1, int R2> struct X3Cfg {
  static constexpr int N = R1 * R2, NT = X3L * (R1 > R2 ? R1 : R2);
  static constexpr int WORK = N * X3LW + R1 * X3PAD, TP = N + 1, TRANS = X3L * TP;
  static constexpr int BUF = WORK > TRANS ? WORK : TRANS;
  static constexpr size_t SMEM = (size_t)(BUF + N) * sizeof(float2);
};
template <int R1, int R2>
__global__ void __launch_bounds__(X3Cfg<R1, R2>::NT, 3) k_fft_x_inv3(FftGeom g, const float2* __restrict__ B, float* __restrict__ F,
                                                                     const float2* __restrict__ tw_g, unsigned* __restrict__ f2max) {
  using C = X3Cfg<R1, R2>;
  constexpr int N = C::N, NT = C::NT, NHC = N / 2 + 1, TP = C::TP;
  extern __shared__ float2 smem[];
  float2* s = smem;
  float2* tw = smem + C::BUF;
  __shared__ float wmax[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * 16, zp = blockIdx.y, b = blockIdx.z;
  const size_t plane = (size_t)N * g.P;
  // Z[k] = Xa[k] + i Xb[k], Z[N-k] = conj(Xa[k]) + i conj(Xb[k]).  Thread -> (component grp, kx = kk [+ KQ]); it walks
  // the 8 line pairs with constant pointer increments (no per-element index arithmetic), 16 loads in flight.
  constexpr int KQ = NT / 3;
  const int grp = tid / KQ, kk = tid - grp * KQ;
  const int nrow = g.M - r0;  // kept rows left from r0 on (>= 1)
  load_tw(tw, tw_g, N);
  const bool full = nrow >= 16;  // all but the last row block: no per-row predicates on the hot path
  for (int k = kk; k < NHC; k += KQ) {
    const float2* src = B + (((size_t)grp * g.nbatch + b) * g.M + zp) * plane + (size_t)(g.off + r0) * g.P + k;
    float2 va[8], vb[8];
    if (full) {
#pragma unroll
      for (int l = 0; l < 8; l++) { va[l] = __ldg(src + (size_t)(2 * l) * g.P); vb[l] = __ldg(src + (size_t)(2 * l + 1) * g.P); }
    } else {
#pragma unroll
      for (int l = 0; l < 8; l++) {
        va[l] = 2 * l < nrow ? __ldg(src + (size_t)(2 * l) * g.P) : make_float2(0.f, 0.f);
        vb[l] = 2 * l + 1 < nrow ? __ldg(src + (size_t)(2 * l + 1) * g.P) : make_float2(0.f, 0.f);
      }
    }
    float2* sk = s + k * X3LW + (k / R2) * X3PAD + grp * 8;
    const int m = N - k;
    float2* sm = s + m * X3LW + (m / R2) * X3PAD + grp * 8;
    if (k && 2 * k != N) {
#pragma unroll
      for (int l = 0; l < 8; l++) {
        const float2 a = va[l], c = vb[l];
        sk[l] = c_add_i<1>(a, c);                                                              // Xa + i Xb
        sm[l] = c_add(make_float2(a.x, -a.y), make_float2(c.y, c.x));                          // conj(Xa) + i conj(Xb)
      }
    } else {
#pragma unroll
      for (int l = 0; l < 8; l++) sk[l] = c_add_i<1>(va[l], vb[l]);
    }
  }
  __syncthreads();
  const int line = tid % X3L, idx = tid / X3L;
  if (idx < R2) {  // R1-point DFTs over n1 for fixed n2 = idx, twiddle, back to the same slots
    float2 v[R1];
#pragma unroll
    for (int n1 = 0; n1 < R1; n1++) v[n1] = s[(n1 * R2 + idx) * X3LW + n1 * X3PAD + line];
    dft<R1, +1>(v);
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) s[(k1 * R2 + idx) * X3LW + k1 * X3PAD + line] = k1 ? mul_tw<+1>(v[k1], tw[idx * k1]) : v[0];
  }
  __syncthreads();
  float2 v[R2];
  if (idx < R1) {  // R2-point DFT over n2 for fixed k1 = idx; v[k2] = x[k1 + R1*k2]
#pragma unroll
    for (int n2 = 0; n2 < R2; n2++) v[n2] = s[(idx * R2 + n2) * X3LW + idx * X3PAD + line];
    dft<R2, +1>(v);
  }
  __syncthreads();
  if (idx < R1) {  // transposed: T[line][x], pitch N+1 (odd: the 16 lines of a half-warp land on 16 banks)
#pragma unroll
    for (int k2 = 0; k2 < R2; k2++) s[line * TP + idx + R1 * k2] = v[k2];
  }
  __syncthreads();
  // epilogue: thread -> (x = kk [+ KQ ...], row pairs l = grp, grp+3, grp+6); .x of T is row r0+2l, .y is row r0+2l+1
  float best = 0.f;
  float* dst = F + (((size_t)b * g.M + zp) * g.M + r0) * 3 * (size_t)g.FP;
  const size_t rp = 3 * (size_t)g.FP;
  const int fp = g.FP;
  for (int x = kk; x < g.M; x += KQ) {
    const float2* t = s + x + g.off;
    if (full) {
#pragma unroll
      for (int l = grp; l < 8; l += 3) {
        const float2 f0 = t[l * TP], f1 = t[(8 + l) * TP], f2 = t[(16 + l) * TP];
        float* row = dst + (size_t)(2 * l) * rp + x;
        row[0] = f0.x; row[fp] = f1.x; row[2 * fp] = f2.x;
        row[3 * fp] = f0.y; row[4 * fp] = f1.y; row[5 * fp] = f2.y;
        best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0.x, f0.x), __fmul_rn(f1.x, f1.x)), __fmul_rn(f2.x, f2.x)));
        best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0.y, f0.y), __fmul_rn(f1.y, f1.y)), __fmul_rn(f2.y, f2.y)));
      }
    } else {
      for (int l = grp; l < 8; l += 3) {
        if (2 * l < nrow) {
          const float2 f0 = t[l * TP], f1 = t[(8 + l) * TP], f2 = t[(16 + l) * TP];
          float* row = dst + (size_t)(2 * l) * rp + x;
          row[0] = f0.x; row[fp] = f1.x; row[2 * fp] = f2.x;
          best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0.x, f0.x), __fmul_rn(f1.x, f1.x)), __fmul_rn(f2.x, f2.x)));
          if (2 * l + 1 < nrow) {
            row[3 * fp] = f0.y; row[4 * fp] = f1.y; row[5 * fp] = f2.y;
            best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0.y, f0.y), __fmul_rn(f1.y, f1.y)), __fmul_rn(f2.y, f2.y)));
          }
        }
      }
    }
  }
  best = __uint_as_float(__reduce_max_sync(__activemask(), __float_as_uint(best)));  // the last warp may be partial
  if (lane == 0) wmax[warp] = best;
  __syncthreads();
  if (tid == 0) {
    float m = 0.f;
    for (int w = 0; w < (NT + 31) / 32; w++) m = fmaxf(m, wmax[w]);
    if (__float_as_uint(m) > f2max[b]) atomicMax(&f2max[b], __float_as_uint(m));
  }
}

// f2_max_fine(tile) = maxval(sum(force_f**2,1))  (pm.f90:85) on F[b][z'][y'][d][x'];  grid = (blocks, nbatch)
__global__ void __launch_bounds__(256) k_f2max_rows(FftGeom g, const float* __restrict__ F, unsigned* __restrict__ f2max) {
  const int b = blockIdx.y;
  const size_t nrow = (size_t)g.M * g.M;
  const float* base = F + (size_t)b * nrow * 3 * g.FP;
  float best = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (size_t r = (size_t)blockIdx.x * nw + warp; r < nrow; r += (size_t)gridDim.x * nw) {
    const float* row = base + r * 3 * g.FP;
    for (int x = lane; x < g.M; x += 32) {
      const float f0 = row[x], f1 = row[g.FP + x], f2 = row[2 * g.FP + x];
      best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
    }
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if (lane == 0) atomicMax(&f2max[b], __float_as_uint(best));
}

// F <- F*a_mid*dt/6/pi in place (diagnostic path: a caller-supplied force_f goes through the same kick as the step's)
__global__ void __launch_bounds__(256) k_prefix_rows(FftGeom g, float* __restrict__ F, float a_mid, float dt) {
  const size_t nrow = (size_t)g.M * g.M * 3;
  float* base = F + (size_t)blockIdx.y * nrow * g.FP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (size_t r = (size_t)blockIdx.x * nw + warp; r < nrow; r += (size_t)gridDim.x * nw)
    for (int x = lane; x < g.M; x += 32) base[r * g.FP + x] = kick_prefix(base[r * g.FP + x], a_mid, dt);
}

}  // namespace cube
