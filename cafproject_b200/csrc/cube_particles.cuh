// cube_particles.cuh -- particle-parallel kernels (one thread per particle) on CUBE's cell-ordered storage.
//
// A CTA owns PC_CELLS consecutive coarse cells in file order; their particles are one contiguous run of the
// int16 arrays, so a warp's loads/stores of xp/vp are coalesced.  Each thread finds the cell of its particle by a
// binary search in the CTA's prefix offsets (shared memory).  All arithmetic that feeds an integer code keeps the
// reference's operation order and rounding (cube_common.cuh).
#pragma once
#include "cube_common.cuh"

namespace cube {

constexpr int PC_CELLS = 128;  // file-order coarse cells per CTA (ghost-particle kernels)
constexpr int PC_T = 256;      // threads per CTA (ghost-particle kernels)

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
template <int VB> __device__ __forceinline__ long long enc_exact(double X) {
  return llround(__dmul_rn((double)((1 << VB) - 1), atan(X)) / (double)PI_F);
}
template <int VB> __global__ void k_build_enc(double* __restrict__ B) {
  constexpr int VH = 1 << (VB - 1);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > VH - 1) return;
  if (c == VH - 1) { B[c] = __longlong_as_double(0x7ff0000000000000LL); return; }
  unsigned long long lo = 0, hi = (unsigned long long)__double_as_longlong(1e300);
  while (lo < hi) {
    const unsigned long long mid = lo + ((hi - lo) >> 1);
    if (enc_exact<VB>(__longlong_as_double((long long)mid)) >= c + 1) hi = mid; else lo = mid + 1;
  }
  B[c] = __longlong_as_double((long long)hi);
}
// nint(real(nvbin-1)*atan(S*v)/pi,kind=izipv)  (pm.f90:113, update_particle.f90:86).
// cr = 65535*atanf(a)/pi_f in f32 is within 0.01 of the exact real-valued code (atanf: 1 ulp, CUDA math API; the f32
// roundings of a and of the product add < 0.005), and the code is its rounding to nearest: unless cr lies within 1/16 of a
// half-integer that rounding is already decided; otherwise (1 encode in 8) the exact threshold T[c] between codes c and
// c+1, c = floor(cr), decides it.  Same result as the formula for every input, 1/8 table load per encode instead of 2.
template <int VB> __device__ __forceinline__ short vp_encode_lut(double v, double S, const double* __restrict__ B) {
  constexpr int CMAX = (1 << (VB - 1)) - 1;
  const double X = __dmul_rn(S, v);
  const double a = fabs(X);
  const float cr = atanf((float)a) * ((float)((1 << VB) - 1) / PI_F);
  const float fl = floorf(cr), fr = cr - fl;
  int c = (int)fl;
  if (fabsf(fr - 0.5f) <= 0.0625f) { c = min(c, CMAX); c += (int)(a >= __ldg(B + c)); }
  else c += (int)(fr > 0.5f);
  c = min(c, CMAX);
  return (short)(X < 0.0 ? -c : c);
}

// =============================================================================================
// Velocity-code tables in SHARED memory.
// The particle kernels were bound by the L1 data pipe (ncu: l1tex__data_pipe_lsu_wavefronts 77-90 % of peak,
// profiles/r01h_ncu_full_cfg1.csv): every decode/encode gathered from the 512 KB f64 tables in global memory, ~23
// wavefronts per warp-wide gather.  A random 4-byte gather from shared memory costs ~3.  So the drift's placement pass and
// the coarse kick run one 1024-thread CTA per SM, keep the "hot" part of the decode table in shared memory as f32 and stay
// bit-identical (measured at cfg 2: place 5.4 -> 3.6 ms, coarse kick 4.5 -> 3.3 ms; the key pass and the fine kick, which
// need the L1 cache that the shared-memory table takes away, are faster in the small-CTA form and keep it):
//   decode  dv = dble(tanf(pi*vp/N)) / S : s_tan[|vp|] is the host tanf table (odd: checked at init) for |vp| < VT_HOT,
//           the f64 division is done as q0 = t*rS, q = fma(fma(-S,q0,t), rS, q0) with rS = 1/S, which k_build_dvlut
//           checks against t/S for every table entry each time S changes (vt.divok; else a true division);
//   encode  needs its table once in 8 calls (vp_encode_lut) and reads it from global memory;
//   codes at or beyond VT_HOT (|v| > 4.8 sigma) use the global f64 decode table.
// A warp owns WC consecutive cells of the file order at a time (their particles are one contiguous run): warp-private
// prefix offsets, no block barriers after the table fill.
// =============================================================================================
constexpr int VT_HOT = 24576;   // |code| < VT_HOT = 4.8 sigma: the part of the half table the kernels were first given
constexpr int VT_ALL = 32768;   // the whole half table (2-byte codes; 128 entries for 1-byte codes), the default: the velocity distribution of
                                // a clustered state has long tails (z = 0 state of cfg 2: place 7.28 -> 6.78 ms, coarse kick 5.11 -> 4.99)
constexpr int WC = 32;         // cells per warp chunk
constexpr int PW_T = 1024;     // threads per CTA
constexpr int PW_W = PW_T / 32;
constexpr unsigned FULL = 0xffffffffu;

struct VTab {
  const float* tanlut;  // [nvbin] host tanf, indexed by the code's raw pattern
  const float* tanh;    // [nvbin/2+1] host tanf of codes 0..nvbin/2 (the last one = -tanf(most negative code))
  const double* thr;    // [nvbin/2] exact encoder thresholds
  const double* dvlut;  // [nvbin] f64 decode table of the current S
  const int* divok;     // != 0: the FMA division reproduces t/S for every table entry (current S)
  int hot;              // VT_HOT, or 0 when the host tanf table is not odd (everything takes the global tables)
};

// tile and tile-local coordinates of a warp's cells, worked out once per cell
struct CellPos { short tx, ty, tz, i, j, k, tile, pad; };
struct WarpScratch { int soff[WC + 1]; unsigned mask[WC]; CellPos pos[WC]; };
constexpr int PW_SMEM_FULL = VT_HOT * 4 + PW_W * (int)sizeof(WarpScratch);
constexpr int PW_SMEM_MAX = VT_ALL * 4 + PW_W * (int)sizeof(WarpScratch);
__host__ __device__ inline int pw_smem_bytes(int hot) { return hot * 4 + PW_W * (int)sizeof(WarpScratch); }

__device__ __forceinline__ void fill_tab(float* dst, const float* __restrict__ src, int n) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

struct VDec { const float* s_tan; const double* dvlut; double S, rS; int hot, fast; };
__device__ __forceinline__ VDec make_dec(const VTab& vt, const float* s_tan, double S) {
  VDec d; d.s_tan = s_tan; d.dvlut = vt.dvlut; d.S = S; d.rS = 1.0 / S; d.hot = vt.hot; d.fast = *vt.divok;
  return d;
}
template <int VB> __device__ __forceinline__ double v_decode(const VDec& d, short c) {
  const int a = abs((int)c);
  if (a < d.hot) {
    const float t = d.s_tan[a];
    const double td = (double)(c < 0 ? -t : t);
    if (d.fast) { const double q0 = __dmul_rn(td, d.rS); return __fma_rn(__fma_rn(-d.S, q0, td), d.rS, q0); }
    return __ddiv_rn(td, d.S);
  }
  return __ldg(d.dvlut + upat<VB>(c));
}
// host tanf value of a code: shared-memory half table below `hot`, the full global table beyond
template <int VB> __device__ __forceinline__ float tan_value(const VDec& d, const float* __restrict__ tanlut, short c) {
  const int a = abs((int)c);
  if (a < d.hot) { const float t = d.s_tan[a]; return c < 0 ? -t : t; }
  return __ldg(tanlut + upat<VB>(c));
}
// prefix offsets and coordinates of the warp's cells [c0, c0+WC) (clipped at cend); returns the number of particles,
// p0 = index of the first one.  cstart[cend] must be readable (next cell's start or the sentinel).
__device__ __forceinline__ int warp_chunk_setup(const Geom& g, const long long* __restrict__ cstart, long long c0, long long cend,
                                                WarpScratch* ws, int lane, long long& p0) {
  const long long c = c0 + lane < cend ? c0 + lane : cend;
  const long long cs = cstart[c];
  const long long ce = cstart[c0 + WC < cend ? c0 + WC : cend];
  p0 = __shfl_sync(FULL, cs, 0);
  __syncwarp();  // the previous chunk's readers are done
  ws->soff[lane] = (int)(cs - p0);
  if (lane == 0) ws->soff[WC] = (int)(ce - p0);
  ws->mask[lane] = 0u;
  CellPos q = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 + lane < cend) {
    int tx, ty, tz, i, j, k;
    phys_decompose(g, c0 + lane, tx, ty, tz, i, j, k);
    q.tx = (short)tx; q.ty = (short)ty; q.tz = (short)tz; q.i = (short)i; q.j = (short)j; q.k = (short)k;
    q.tile = (short)((tz * g.nnt + ty) * g.nnt + tx);
  }
  ws->pos[lane] = q;
  __syncwarp();
  return (int)(ce - p0);
}
// cell (0..WC-1) of particle q of the chunk: largest c with soff[c] <= q
__device__ __forceinline__ int warp_chunk_find(const WarpScratch* ws, int q) {
  int lo = 0;
#pragma unroll
  for (int step = WC / 2; step > 0; step >>= 1)
    if (ws->soff[lo + step] <= q) lo += step;
  return lo;
}

// tile and tile-local coordinates of the CTA's cells, worked out once per cell (the integer divisions of phys_decompose
// cost more than the rest of a particle's index arithmetic when they are redone per particle)
__device__ __forceinline__ void chunk_cells(const Geom& g, long long c0, long long ncell, CellPos* sp) {
  for (int t = threadIdx.x; t < PC_CELLS; t += blockDim.x) {
    CellPos q = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 + t < ncell) {
      int tx, ty, tz, i, j, k;
      phys_decompose(g, c0 + t, tx, ty, tz, i, j, k);
      q.tx = (short)tx; q.ty = (short)ty; q.tz = (short)tz; q.i = (short)i; q.j = (short)j; q.k = (short)k;
    }
    sp[t] = q;
  }
}

// ---------------------------------------------------------------------------------------------
// self-test of the table-driven code conversions against their defining formulas (cube_gpu_selftest_codes):
//   encode: vp_encode_lut(X, 1, T) vs vp_encode(X, 1) (cube_common.cuh: f64 atan formula) for X just below, at and just
//           above every threshold, both signs, plus a geometric sweep of `nsweep` values over [1e-12, 1e8];
//   decode: v_decode (shared-memory f32 table + FMA division) vs dvlut[] for all 65536 codes.
// ---------------------------------------------------------------------------------------------
template <int VB>
__global__ void __launch_bounds__(256) k_selftest_encode(const double* __restrict__ T, long long nsweep, unsigned long long* __restrict__ bad) {
  constexpr long long NT3 = 3LL * ((1 << (VB - 1)) - 1);
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double X;
  if (q < NT3) {
    const int c = (int)(q / 3), w = (int)(q % 3);
    const long long bits = __double_as_longlong(T[c]);
    X = __longlong_as_double(bits + (w - 1));
  } else if (q < NT3 + nsweep) {
    const double f = (double)(q - NT3) / (double)nsweep;
    X = exp(-27.631021115928547 + f * 46.051701859880914);  // 1e-12 .. 1e8
  } else return;
  int n = 0;
  if (vp_encode_lut<VB>(X, 1.0, T) != vp_encode<VB>(X, 1.0)) n++;
  if (vp_encode_lut<VB>(-X, 1.0, T) != vp_encode<VB>(-X, 1.0)) n++;
  if (n) atomicAdd(bad, (unsigned long long)n);
}
template <int VB>
__global__ void __launch_bounds__(PW_T, 1) k_selftest_decode(VTab vt, double S, unsigned long long* __restrict__ bad) {
  extern __shared__ __align__(16) unsigned char pw_smem[];
  float* s_tan = reinterpret_cast<float*>(pw_smem);
  if (vt.hot) fill_tab(s_tan, vt.tanh, vt.hot);
  const VDec dec = make_dec(vt, s_tan, S);
  __syncthreads();
  int n = 0;
  for (int u = threadIdx.x; u < (1 << VB); u += PW_T) {
    const short c = (short)(u >= (1 << (VB - 1)) ? u - (1 << VB) : u);
    if (__double_as_longlong(v_decode<VB>(dec, c)) != __double_as_longlong(vt.dvlut[u])) n++;
  }
  if (n) atomicAdd(bad, (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------------
// fine kick (pm.f90:88-118) for the tiles [tile0, tile0+nb): grid = (chunks per tile, nb)
// Stays on the small-CTA / global-table form: its 24 force gathers per particle live off the L1 cache, and the 221 KB of
// shared-memory tables of the warp kernels leave 28 KB of L1 (measured: 10.8 ms against 5.2 ms at cfg 2).
// G[b][z'][y'][d][x'] = force_f*a_mid*dt/6/pi (the per-node prefix of every kick term, applied once per mesh node in
// the epilogue of the x inverse, cube_fft.cuh) on the M kept points
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(PC_T) k_fine_kick_p(Geom g, int tile0, int M, int FP, const typename F::XT* __restrict__ xp, typename F::VT* __restrict__ vp,
                                                     const long long* __restrict__ cstart_p, const float* __restrict__ G,
                                                     const double* __restrict__ dvlut, const double* __restrict__ enc, double S_new) {
  __shared__ int soff[PC_CELLS + 1];
  __shared__ CellPos spos[PC_CELLS];
  const long long nt3 = (long long)g.nt * g.nt * g.nt;
  const int b = blockIdx.y;
  const long long tbase = (long long)(tile0 + b) * nt3;
  const long long c0 = tbase + (long long)blockIdx.x * PC_CELLS;
  chunk_cells(g, c0, tbase + nt3, spos);
  const int np = chunk_setup(cstart_p, c0, tbase + nt3, soff);
  const long long p0 = cstart_p[c0];
  const float* Gb = G + (long long)b * M * M * 3 * FP;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const CellPos cp = spos[chunk_find(soff, q)];
    const int i = cp.i + 1, j = cp.j + 1, k = cp.k + 1;
    const long long p = p0 + q;
    const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
    int i1, j1, k1; float ax[2], ay[2], az[2];
    cic_split(fine_tempx<F::XB>(i, xc.x), i1, ax[0], ax[1]);  // idx1 of pm.f90:96 = 0-based kept index
    cic_split(fine_tempx<F::XB>(j, xc.y), j1, ay[0], ay[1]);
    cic_split(fine_tempx<F::XB>(k, xc.z), k1, az[0], az[1]);
    double v0 = dvlut[upat<F::VB>(vc.x)], v1 = dvlut[upat<F::VB>(vc.y)], v2 = dvlut[upat<F::VB>(vc.z)];
    const int qx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, qy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, qz[8] = {0, 0, 0, 1, 1, 1, 0, 1};  // pm.f90:104-111
    float f0[8], f1[8], f2[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float* f = Gb + ((long long)(k1 + qz[t]) * M + (j1 + qy[t])) * 3 * FP + (i1 + qx[t]);
      f0[t] = __ldg(f); f1[t] = __ldg(f + FP); f2[t] = __ldg(f + 2 * FP);
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
      const float wx = ax[qx[t]], wy = ay[qy[t]], wz = az[qz[t]];
      v0 = __dadd_rn(v0, (double)kick_weight(f0[t], wx, wy, wz));
      v1 = __dadd_rn(v1, (double)kick_weight(f1[t], wx, wy, wz));
      v2 = __dadd_rn(v2, (double)kick_weight(f2[t], wx, wy, wz));
    }
    store_code3(vp, p, vp_encode_lut<F::VB>(v0, S_new, enc), vp_encode_lut<F::VB>(v1, S_new, enc), vp_encode_lut<F::VB>(v2, S_new, enc));
  }
}

// ---------------------------------------------------------------------------------------------
// coarse kick (pm.f90:196-228); Gc(3,0:nc+1,0:nc+1,0:nc+1) = kick_prefix(force_c), vmax over v+vfield (no abs)
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(PW_T, 1) k_coarse_kick_w(Geom g, VTab vt, double S, const typename F::XT* __restrict__ xp, typename F::VT* __restrict__ vp,
                                                          const long long* __restrict__ cstart_p, const float* __restrict__ vfield_p,
                                                          const float* __restrict__ Gc, unsigned long long* __restrict__ vmax_bits /* [4]: scalar, |.| per component */,
                                                          long long c_begin, long long c_end /* file-order cell range */) {
  extern __shared__ __align__(16) unsigned char pw_smem[];
  float* s_tan = reinterpret_cast<float*>(pw_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(pw_smem + vt.hot * 4) + warp;
  if (vt.hot) fill_tab(s_tan, vt.tanh, vt.hot);
  const VDec dec = make_dec(vt, s_tan, S);
  __syncthreads();
  const int m = g.nc + 2;
  double vm = 0.0, va0 = 0.0, va1 = 0.0, va2 = 0.0;
  for (long long c0 = c_begin + ((long long)blockIdx.x * PW_W + warp) * WC; c0 < c_end; c0 += (long long)gridDim.x * PW_W * WC) {
    long long p0;
    const int np = warp_chunk_setup(g, cstart_p, c0, c_end, ws, lane, p0);
    for (int q = lane; q < np; q += 32) {
      const int cl = warp_chunk_find(ws, q);
      const long long L = c0 + cl;
      const CellPos cp = ws->pos[cl];
      const int X = cp.tx * g.nt + cp.i, Y = cp.ty * g.nt + cp.j, Z = cp.tz * g.nt + cp.k;  // ((itx-1)*nt + (i-1)) of pm.f90:206
      const long long p = p0 + q;
      const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
      int i1, j1, k1; float ax[2], ay[2], az[2];
      cic_split(coarse_tempx<F::XB>(X, xc.x), i1, ax[0], ax[1]);
      cic_split(coarse_tempx<F::XB>(Y, xc.y), j1, ay[0], ay[1]);
      cic_split(coarse_tempx<F::XB>(Z, xc.z), k1, az[0], az[1]);
      double v0 = v_decode<F::VB>(dec, vc.x), v1 = v_decode<F::VB>(dec, vc.y), v2 = v_decode<F::VB>(dec, vc.z);
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
      const double t0 = __dadd_rn(v0, vf0), t1 = __dadd_rn(v1, vf1), t2 = __dadd_rn(v2, vf2);
      vm = fmax(vm, fmax(t0, fmax(t1, t2)));                                            // pm.f90:220
      va0 = fmax(va0, fabs(t0)); va1 = fmax(va1, fabs(t1)); va2 = fmax(va2, fabs(t2));  // CUBEnu pm.f90:349: vmax(3), with abs
      store_code3(vp, p, vp_encode_lut<F::VB>(v0, S, vt.thr), vp_encode_lut<F::VB>(v1, S, vt.thr), vp_encode_lut<F::VB>(v2, S, vt.thr));
    }
  }
  for (int o = 16; o; o >>= 1) {
    vm = fmax(vm, __shfl_down_sync(FULL, vm, o));
    va0 = fmax(va0, __shfl_down_sync(FULL, va0, o)); va1 = fmax(va1, __shfl_down_sync(FULL, va1, o)); va2 = fmax(va2, __shfl_down_sync(FULL, va2, o));
  }
  if (lane == 0) {  // non-negative doubles order like their bit patterns
    if (vm > 0.0) atomicMax(vmax_bits, (unsigned long long)__double_as_longlong(vm));
    if (va0 > 0.0) atomicMax(vmax_bits + 1, (unsigned long long)__double_as_longlong(va0));
    if (va1 > 0.0) atomicMax(vmax_bits + 2, (unsigned long long)__double_as_longlong(va1));
    if (va2 > 0.0) atomicMax(vmax_bits + 3, (unsigned long long)__double_as_longlong(va2));
  }
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

// =============================================================================================
// drift (update_particle.f90) in three particle-parallel / cell-parallel passes
//   A  k_drift_key_p    per particle : destination offset key (+ near-tie flag), max |offset|
//   B  k_drift_count    per destination cell: visits its source cells in the reference's traversal order
//                       (tile-local k,j,i, then storage order), counts, chains vfield_new (order-dependent f32
//                       rounding, update_particle.f90:47), and writes every accepted particle's rank in its cell
//   C  k_drift_place_w  per particle : pos = cstart_new[dest] + rank ; xp_new, vp_new, velocity statistics
// Near-tie particles (flagged in A) are decided in B in the destination tile's frame, like the reference.
// rank[p] = rank in the destination cell (20 bits) | destination offset (3 x 4 bits, biased by 8) << 20; A presets
// 0xFFFFFFFF for flagged particles so that one no destination accepts is dropped, as the reference would.
// (No atomics inside the gather loop: they make nvcc give up warp reconvergence -- 4x the instructions.)
// =============================================================================================
constexpr double TIE_EPS = 1e-9;
constexpr unsigned RANK_LOST = 0xFFFFFFFFu;
constexpr int RANK_BITS = 20;
__device__ __forceinline__ unsigned off_pack12(int dx, int dy, int dz) { return (unsigned)(dx + 8) | ((unsigned)(dy + 8) << 4) | ((unsigned)(dz + 8) << 8); }

// per-source-cell summary of pass A, four words per cell: bit (dx+2)+5(dy+2)+25(dz+2) = "some particle of this cell moves by
// (dx,dy,dz)" for offsets up to two cells, bit 125 = "some particle moves farther".  A particle that sits on a cell boundary
// (near-tie: its destination is re-decided in the destination tile's frame and may come out one cell off) sets the bits of
// all 27 offsets around its own.  Pass B only walks the particles of a source cell when the bit of the offset it is looking
// for is set: at small time steps almost every particle stays in its cell, and inside a halo (10^3-10^4 particles per cell,
// velocity dispersion of a cell per step) a destination still skips the source cells that send it nothing.
constexpr int MASK_W = 4;
constexpr int MASK_FAR_BIT = 125;
__device__ __forceinline__ int offset_bit(int dx, int dy, int dz) { return (dx + 2) + 5 * (dy + 2) + 25 * (dz + 2); }
// does the summary `m` (pointer to the cell's four words) announce movers by (dx,dy,dz)?
__device__ __forceinline__ bool mask_hit(const unsigned* __restrict__ m, int dx, int dy, int dz) {
  const int b = max(abs(dx), max(abs(dy), abs(dz))) <= 2 ? offset_bit(dx, dy, dz) : MASK_FAR_BIT;
  return (m[b >> 5] >> (b & 31)) & 1u;
}
__device__ __forceinline__ void mask_set(unsigned* m, int ox, int oy, int oz, bool tie) {
  if (!tie) {
    const int b = max(abs(ox), max(abs(oy), abs(oz))) <= 2 ? offset_bit(ox, oy, oz) : MASK_FAR_BIT;
    // almost every particle of a cell sets the same bit: look before the atomic (a stale read only costs a redundant atomicOr)
    if (!((*(volatile unsigned*)(m + (b >> 5)) >> (b & 31)) & 1u)) atomicOr(m + (b >> 5), 1u << (b & 31));
    return;
  }
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int x = ox + dx, y = oy + dy, z = oz + dz;
        const int b = max(abs(x), max(abs(y), abs(z))) <= 2 ? offset_bit(x, y, z) : MASK_FAR_BIT;
        atomicOr(m + (b >> 5), 1u << (b & 31));
      }
}
// the bits of the offsets within one cell, per word (everything else = "moves two cells or more")
__host__ __device__ constexpr unsigned mask_inner_word(int w) {
  unsigned v = 0;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++) {
        const int b = (dx + 2) + 5 * (dy + 2) + 25 * (dz + 2);
        if ((b >> 5) == w) v |= 1u << (b & 31);
      }
  return v;
}
// Block flags: 4^3 blocks of the extended grid.  FLAG_FAR = the block holds a source cell with movers beyond one cell: a
// destination whose (2r+1)^3 neighbourhood touches no such block only has to visit its 27 nearest source cells, whatever the
// step radius r.  FLAG_CROWD = the block holds a cell with more than `crowd` particles: a destination with no such block
// nearby cannot have many candidates and skips the candidate count that decides between the thread and the warp path.
constexpr int FARB = 4;
constexpr int FLAG_FAR = 1, FLAG_CROWD = 2;
__host__ __device__ inline int farblk_dim(const Geom& g) { return (g.ne + FARB - 1) / FARB; }

// destination cell of one coordinate, tile-local Fortran index `cell1` (update_particle.f90:41-45)
template <int XB> __device__ __forceinline__ int drift_dest(int cell1, short xp, double v, double dt_mid, bool& tie) {
  double xq = __dadd_rn((double)(cell1 - 1), xp_frac<XB>(xp));
  double dx = __dmul_rn(__dmul_rn(dt_mid, v), 0.25);  // (dt_mid*vreal)/ncell, ncell=4: exact scaling
  double s = __dadd_rn(xq, dx);
  double c = ceil(s);
  tie = tie || (c - s < TIE_EPS) || (s - (c - 1.0) < TIE_EPS);
  return (int)c;
}

// file-order index of the physical cell at offset (ox,oy,oz) from the cell `cp`; false when that cell belongs to another image
// (|offset| <= ncb < nt: at most one tile step per dimension, no divisions)
__device__ __forceinline__ bool dest_cell(const Geom& g, const CellPos& cp, int ox, int oy, int oz, long long& D) {
  const int nt = g.nt, nnt = g.nnt;
  int i = cp.i + ox, j = cp.j + oy, k = cp.k + oz, tx = cp.tx, ty = cp.ty, tz = cp.tz;
  if (i < 0) { i += nt; tx--; } else if (i >= nt) { i -= nt; tx++; }
  if (j < 0) { j += nt; ty--; } else if (j >= nt) { j -= nt; ty++; }
  if (k < 0) { k += nt; tz--; } else if (k >= nt) { k -= nt; tz++; }
  // nn_d == 1: the neighbour image is this image (periodic wrap); nn_d > 1: the cell is another image's
  if (g.nn[0] == 1) tx = tx < 0 ? tx + nnt : (tx >= nnt ? tx - nnt : tx);
  if (g.nn[1] == 1) ty = ty < 0 ? ty + nnt : (ty >= nnt ? ty - nnt : ty);
  if (g.nn[2] == 1) tz = tz < 0 ? tz + nnt : (tz >= nnt ? tz - nnt : tz);
  if ((unsigned)tx >= (unsigned)nnt || (unsigned)ty >= (unsigned)nnt || (unsigned)tz >= (unsigned)nnt) return false;
  D = phys_index(g, tx, ty, tz, i, j, k);
  return true;
}
// a particle of cell `cp` heads for offset (ox,oy,oz): mark the destination "has arrivals" (pass B then walks it); a near-tie may
// come out one cell off in the destination tile's frame: mark the 27 cells around (its own cell is among them or marked too)
__device__ __noinline__ void flag_arrival_tie(const Geom& g, const CellPos& cp, int ox, int oy, int oz, unsigned char* __restrict__ inflag) {
  long long D;
  for (int dz = -1; dz <= 1; dz++)
    for (int dy = -1; dy <= 1; dy++)
      for (int dx = -1; dx <= 1; dx++)
        if (dest_cell(g, cp, ox + dx, oy + dy, oz + dz, D)) inflag[D] = 1;
  if (dest_cell(g, cp, 0, 0, 0, D)) inflag[D] = 1;
}
__device__ __forceinline__ void flag_arrival(const Geom& g, const CellPos& cp, int ox, int oy, int oz, bool tie, unsigned char* __restrict__ inflag) {
  long long D;
  if (tie) flag_arrival_tie(g, cp, ox, oy, oz, inflag);
  else if (dest_cell(g, cp, ox, oy, oz, D)) inflag[D] = 1;
}

// Pass A (+ the bulk of pass B).  Per particle: destination offset key (+ near-tie flag), max |offset|, per-source-cell mover
// summary -- and every mover marks its destination cell in `inflag`.  At the time steps of a run almost every particle stays in
// its cell and most cells receive nobody: for such a cell the reference's arrival order IS its own storage order, so this kernel
// also walks every cell's stayers (their velocities are still in shared memory), ranks them, chains vfield_new
// (update_particle.f90:47) and writes the cell's count and mean velocity.  Cells that do receive somebody (inflag) are redone by
// pass B from scratch -- only those.  A chunk with more particles than the velocity stash holds marks all its cells for pass B.
constexpr int KC_CAPV = 1280;  // particles whose velocities a CTA stashes (a uniform chunk of 128 cells holds ~1024)
constexpr int KC_SMEM = KC_CAPV * (3 * 8 + 1);
template <class F>
__global__ void __launch_bounds__(PC_T, 4) k_drift_key_chain(Geom g, const typename F::XT* __restrict__ xp, const typename F::VT* __restrict__ vp,
                                                            const long long* __restrict__ cstart_p, const float* __restrict__ vfield_p,
                                                            const double* __restrict__ dvlut, double dt_mid, unsigned short* __restrict__ key,
                                                            unsigned* __restrict__ rank, int* __restrict__ maxoff, unsigned* __restrict__ mask_s,
                                                            unsigned char* __restrict__ inflag, int* __restrict__ rhoc_new, float* __restrict__ vfield_new,
                                                            long long cell_base /* first cell of this launch: a streamed upload is keyed chunk by chunk */) {
  extern __shared__ __align__(16) unsigned char kc_smem[];
  double* sv = reinterpret_cast<double*>(kc_smem);                               // [3][KC_CAPV] velocities of the chunk's particles
  unsigned char* scl = reinterpret_cast<unsigned char*>(sv + 3 * KC_CAPV);       // [KC_CAPV] cell of the particle | 0x80 = leaves its cell
  __shared__ int soff[PC_CELLS + 1];
  __shared__ unsigned smask[PC_CELLS * MASK_W];
  __shared__ CellPos spos[PC_CELLS];
  __shared__ float svf[PC_CELLS * 3];
  __shared__ int s_cnt[PC_CELLS];  // movers of the cell
  static_assert(PC_CELLS <= 128, "cell index and mover flag share a byte");
  const long long c0 = cell_base + (long long)blockIdx.x * PC_CELLS;
  for (int t = threadIdx.x; t < PC_CELLS * MASK_W; t += PC_T) smask[t] = 0u;
  for (int t = threadIdx.x; t < PC_CELLS * 3; t += PC_T) svf[t] = c0 * 3 + t < g.ncell_p * 3 ? vfield_p[c0 * 3 + t] : 0.f;
  for (int t = threadIdx.x; t < PC_CELLS; t += PC_T) s_cnt[t] = 0;
  chunk_cells(g, c0, g.ncell_p, spos);
  const int np = chunk_setup(cstart_p, c0, g.ncell_p, soff);
  const long long p0 = cstart_p[c0];
  const bool stash = np <= KC_CAPV;
  int m = 0;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const int cl = chunk_find(soff, q);
    const CellPos cp = spos[cl];
    const int i = cp.i, j = cp.j, k = cp.k;
    const long long p = p0 + q;
    const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
    const double v0 = __dadd_rn(dvlut[upat<F::VB>(vc.x)], (double)svf[3 * cl]);
    const double v1 = __dadd_rn(dvlut[upat<F::VB>(vc.y)], (double)svf[3 * cl + 1]);
    const double v2 = __dadd_rn(dvlut[upat<F::VB>(vc.z)], (double)svf[3 * cl + 2]);
    bool tie = false;
    int ox = drift_dest<F::XB>(i + 1, xc.x, v0, dt_mid, tie) - (i + 1);
    int oy = drift_dest<F::XB>(j + 1, xc.y, v1, dt_mid, tie) - (j + 1);
    int oz = drift_dest<F::XB>(k + 1, xc.z, v2, dt_mid, tie) - (k + 1);
    const int far = max(abs(ox), max(abs(oy), abs(oz)));
    m = max(m, far);
    const bool mover = tie || far != 0;
    if (mover) {
      mask_set(smask + cl * MASK_W, ox, oy, oz, tie);
      // movers beyond the tile buffer stop the step (maxoff); their clamped offsets are never used
      if (far <= NCB) flag_arrival(g, cp, ox, oy, oz, tie, inflag);
      if (stash) atomicAdd(&s_cnt[cl], 1);
    }
    if (stash) {
      scl[q] = (unsigned char)(cl | (mover ? 0x80 : 0));
      // a mover's slot holds -0.0: x + (-0.0) == x for every x, zeros of either sign included, so the chains below need no test
      sv[q] = mover ? -0.0 : v0; sv[KC_CAPV + q] = mover ? -0.0 : v1; sv[2 * KC_CAPV + q] = mover ? -0.0 : v2;
    } else rank[p] = RANK_LOST;  // pass B overwrites it for the one destination cell of this image that accepts the particle
    ox = min(max(ox, -15), 15); oy = min(max(oy, -15), 15); oz = min(max(oz, -15), 15);
    key[p] = (unsigned short)(key_pack(ox, oy, oz) | (tie ? KEY_FLAG : 0u));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxoff, m);
  __syncthreads();
  for (int t = threadIdx.x; t < PC_CELLS * MASK_W; t += PC_T) {
    const int cl = t / MASK_W;
    if (c0 + cl < g.ncell_p) {
      unsigned w = smask[t];
      // stayers are not in the shared summary (they never needed an atomic): the "offset 0" bit is set for every occupied cell,
      // which is all pass B's mask test needs (it re-checks every candidate's key)
      if ((t % MASK_W) == (offset_bit(0, 0, 0) >> 5) && soff[cl + 1] > soff[cl]) w |= 1u << (offset_bit(0, 0, 0) & 31);
      mask_s[c0 * MASK_W + t] = w;
    }
  }
  if (!stash) {  // crowded chunk: all of it goes to pass B (its warp path)
    for (int t = threadIdx.x; t < PC_CELLS; t += PC_T) if (c0 + t < g.ncell_p) inflag[c0 + t] = 1;
    return;
  }
  // ranks of the stayers among the stayers of their cell (= their arrival order when nobody else arrives), coalesced
  const unsigned o12 = off_pack12(0, 0, 0) << RANK_BITS;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const unsigned c = scl[q];
    unsigned r = RANK_LOST;
    if (!(c & 0x80u)) {
      int before = q - soff[c];
      if (s_cnt[c]) for (int q2 = soff[c]; q2 < q; q2++) before -= scl[q2] >> 7;
      r = (unsigned)before | o12;
    }
    rank[p0 + q] = r;
  }
  // vfield_new chains (update_particle.f90:47): one thread per cell, its three components side by side
  const double weight_v = (double)0.1f;  // update_particle.f90:10
  const int cl = threadIdx.x;
  if (cl < PC_CELLS && c0 + cl < g.ncell_p) {
    const long long L = c0 + cl;
    float vfn0 = (float)__dmul_rn((double)svf[3 * cl], weight_v), vfn1 = (float)__dmul_rn((double)svf[3 * cl + 1], weight_v),
          vfn2 = (float)__dmul_rn((double)svf[3 * cl + 2], weight_v);  // :27
    const int qe = soff[cl + 1];
    for (int q = soff[cl]; q < qe; q++) {  // :47, f32 store after each f64 add
      vfn0 = (float)__dadd_rn((double)vfn0, sv[q]);
      vfn1 = (float)__dadd_rn((double)vfn1, sv[KC_CAPV + q]);
      vfn2 = (float)__dadd_rn((double)vfn2, sv[2 * KC_CAPV + q]);
    }
    const int cnt = qe - soff[cl] - s_cnt[cl];
    const double den = __dadd_rn((double)cnt, weight_v);  // :55-57
    rhoc_new[L] = cnt;
    vfield_new[3 * L] = (float)((double)vfn0 / den); vfield_new[3 * L + 1] = (float)((double)vfn1 / den); vfield_new[3 * L + 2] = (float)((double)vfn2 / den);
  }
}

// pass A for the ghost particles received from other images (cube_exchange.cuh): cells in message order,
// gstart = exclusive prefix of their counts (+ sentinel), particles at base + gstart[.]
template <class F>
__global__ void __launch_bounds__(PC_T) k_drift_key_g(Geom g, long long ng, const int* __restrict__ gcell_ext, const long long* __restrict__ gstart,
                                                     long long base, const typename F::XT* __restrict__ xp, const typename F::VT* __restrict__ vp,
                                                     const float* __restrict__ vfield_e, const double* __restrict__ dvlut, double dt_mid,
                                                     unsigned short* __restrict__ key, unsigned* __restrict__ rank, int* __restrict__ maxoff,
                                                     unsigned* __restrict__ mask_g, unsigned char* __restrict__ inflag) {
  __shared__ int soff[PC_CELLS + 1];
  __shared__ unsigned smask[PC_CELLS * MASK_W];
  const long long c0 = (long long)blockIdx.x * PC_CELLS;
  for (int t = threadIdx.x; t < PC_CELLS * MASK_W; t += PC_T) smask[t] = 0u;
  const int np = chunk_setup(gstart, c0, ng, soff);
  const long long p0 = base + gstart[c0];
  int m = 0;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const int cl = chunk_find(soff, q);
    const long long e = gcell_ext[c0 + cl];
    const int i = (int)(e % g.ne) - NCB, j = (int)((e / g.ne) % g.ne) - NCB, k = (int)(e / ((long long)g.ne * g.ne)) - NCB;
    const long long p = p0 + q;
    const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
    const double vf0 = vfield_e[3 * e], vf1 = vfield_e[3 * e + 1], vf2 = vfield_e[3 * e + 2];
    bool tie = false;
    int ox = drift_dest<F::XB>(i + 1, xc.x, __dadd_rn(dvlut[upat<F::VB>(vc.x)], vf0), dt_mid, tie) - (i + 1);
    int oy = drift_dest<F::XB>(j + 1, xc.y, __dadd_rn(dvlut[upat<F::VB>(vc.y)], vf1), dt_mid, tie) - (j + 1);
    int oz = drift_dest<F::XB>(k + 1, xc.z, __dadd_rn(dvlut[upat<F::VB>(vc.z)], vf2), dt_mid, tie) - (k + 1);
    // a ghost can only enter from at most ncb cells away; its owner image checks the full offset of the same particle
    m = max(m, min(NCB, max(abs(ox), max(abs(oy), abs(oz)))));
    mask_set(smask + cl * MASK_W, ox, oy, oz, tie);
    {  // physical cells this ghost may arrive in (a near-tie: the 27 around its computed destination)
      const int tr = tie ? 1 : 0;
      for (int dz = -tr; dz <= tr; dz++)
        for (int dy = -tr; dy <= tr; dy++)
          for (int dx = -tr; dx <= tr; dx++) {
            const int X = i + ox + dx, Y = j + oy + dy, Z = k + oz + dz;
            if ((unsigned)X < (unsigned)g.nc && (unsigned)Y < (unsigned)g.nc && (unsigned)Z < (unsigned)g.nc)
              inflag[phys_index(g, X / g.nt, Y / g.nt, Z / g.nt, X % g.nt, Y % g.nt, Z % g.nt)] = 1;
          }
    }
    ox = min(max(ox, -15), 15); oy = min(max(oy, -15), 15); oz = min(max(oz, -15), 15);
    key[p] = (unsigned short)(key_pack(ox, oy, oz) | (tie ? KEY_FLAG : 0u));
    rank[p] = RANK_LOST;
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(maxoff, m);
  __syncthreads();
  for (int t = threadIdx.x; t < PC_CELLS * MASK_W; t += PC_T)
    if (c0 + t / MASK_W < ng) mask_g[c0 * MASK_W + t] = smask[t];
}

// source-cell summaries on the extended grid: sid_e[e] = file-order index of an aliased cell, ncell_p + q of ghost cell q;
// also sets the block flags (farblk zeroed by the caller)
__global__ void __launch_bounds__(256) k_mask_ext(Geom g, const int* __restrict__ sid_e, const unsigned* __restrict__ mask_s,
                                                  const int* __restrict__ rhoc_e, int crowd, unsigned* __restrict__ mask_e, int* __restrict__ farblk) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ncell_e) return;
  const uint4 m = reinterpret_cast<const uint4*>(mask_s)[sid_e[e]];
  reinterpret_cast<uint4*>(mask_e)[e] = m;
  int f = ((m.x & ~mask_inner_word(0)) | (m.y & ~mask_inner_word(1)) | (m.z & ~mask_inner_word(2)) | (m.w & ~mask_inner_word(3))) ? FLAG_FAR : 0;
  if (rhoc_e[e] > crowd) f |= FLAG_CROWD;
  if (f) {
    const int nb = farblk_dim(g);
    const int x = (int)(e % g.ne), y = (int)((e / g.ne) % g.ne), z = (int)(e / ((long long)g.ne * g.ne));
    atomicOr(&farblk[((z / FARB) * nb + y / FARB) * nb + x / FARB], f);
  }
}
// flags of the blocks touched by the (2r+1)^3 neighbourhood of the destination at extended-grid coordinates (x,y,z)
// (0-based, ghost layers included)
__device__ __forceinline__ int dest_flags(const Geom& g, const int* __restrict__ farblk, int r, int x, int y, int z) {
  const int nb = farblk_dim(g);
  int f = 0;
  for (int bz = (z - r) / FARB; bz <= (z + r) / FARB; bz++)
    for (int by = (y - r) / FARB; by <= (y + r) / FARB; by++)
      for (int bx = (x - r) / FARB; bx <= (x + r) / FARB; bx++) f |= farblk[(bz * nb + by) * nb + bx];
  return f;
}

// (Measured and dropped, profiles/r01i_notes.md: staging the warp's own keys/codes or host-tanf values in shared memory and
//  batching the 27 summary loads at radius 1 did not help -- 3.0 -> 3.0 / 3.4 ms.  ncu: 18 of 32 lanes active on average,
//  25 inner iterations per warp: the cost is the divergence of per-cell particle counts and of the neighbour visits.)
template <class F> struct DriftCountArgs {
  const typename F::XT* xp; const typename F::VT* vp; const unsigned short* key; const int* rhoc_e; const long long* cstart_e; const float* vfield_e;
  const double* dvlut; const unsigned* mask_e; const int* farblk; unsigned* rank; double dt_mid; int r;
  int nlayer;  // 1: CUBE/main (source planes in storage order); > 1: CUBEnu's colour passes over k (update_particle.f90:37,55-58)
};
// Source planes sk in [k-r, k+r] of a destination at tile-local plane k (0-based), in the reference's traversal order.  CUBE/main
// walks k upwards; CUBEnu walks `do ilayer=0,nlayer-1; do k=1-ncb+ilayer,nt+ncb,nlayer`: colour (sk0 + ncb) mod nlayer first, then k.
// Usage: for (PlaneOrder po(k, r, nlayer); po.valid(); po.next()) { const int sk = po.sk; ... }
struct PlaneOrder {
  int sk, last, nlayer, c, c_first, first;
  __device__ __forceinline__ PlaneOrder(int k, int r, int nl) : last(k + r), nlayer(nl < 1 ? 1 : nl), c(0), first(k - r) {
    c_first = ((first + NCB) % nlayer + nlayer) % nlayer;
    seek();
  }
  __device__ __forceinline__ void seek() {  // first plane of colour c (or of the next colour that has one)
    for (; c < nlayer; c++) {
      sk = first + ((c - c_first) % nlayer + nlayer) % nlayer;
      if (sk <= last) return;
    }
  }
  __device__ __forceinline__ bool valid() const { return c < nlayer; }
  __device__ __forceinline__ void next() { sk += nlayer; if (sk > last) { c++; seek(); } }
};
// one candidate particle of source cell (si,sj,sk) for destination (i,j,k): accepted? and its velocity
template <class F>
__device__ __forceinline__ bool drift_accept(const DriftCountArgs<F>& A, long long p, unsigned want, int si, int sj, int sk, int i, int j, int k,
                                             double vf0, double vf1, double vf2, double& v0, double& v1, double& v2) {
  const unsigned kk = A.key[p];
  if (kk == want) {  // common case: one predictable branch, the body is straight-line code
    const Code3 vc = load_code3(A.vp, p);
    v0 = __dadd_rn(A.dvlut[upat<F::VB>(vc.x)], vf0); v1 = __dadd_rn(A.dvlut[upat<F::VB>(vc.y)], vf1); v2 = __dadd_rn(A.dvlut[upat<F::VB>(vc.z)], vf2);
    return true;
  }
  if (kk & KEY_FLAG) {  // near a cell boundary: redo the ceiling in THIS tile's frame
    const Code3 vc = load_code3(A.vp, p), xc = load_code3(A.xp, p);
    v0 = __dadd_rn(A.dvlut[upat<F::VB>(vc.x)], vf0); v1 = __dadd_rn(A.dvlut[upat<F::VB>(vc.y)], vf1); v2 = __dadd_rn(A.dvlut[upat<F::VB>(vc.z)], vf2);
    bool t = false;
    return (drift_dest<F::XB>(si + 1, xc.x, v0, A.dt_mid, t) == i + 1) & (drift_dest<F::XB>(sj + 1, xc.y, v1, A.dt_mid, t) == j + 1) &
           (drift_dest<F::XB>(sk + 1, xc.z, v2, A.dt_mid, t) == k + 1);
  }
  return false;
}

// pass B: one thread per destination (physical) cell, file order.  A destination with more than `heavy` candidate particles
// (everything in the source cells whose summary says somebody may come its way) is not walked by its thread -- a warp would
// wait for its fullest lane, and at late times haloes put 10^3-10^4 particles into one coarse cell -- but by a whole warp:
// 32 candidates are tested at once, ranks come from a ballot prefix, and only the vfield_new chain (f32 rounding after every
// add, update_particle.f90:47) stays serial, fed by shuffles in storage order.  Same traversal order, same results.
constexpr int DC_T = 128;
// the cells marked in `inflag`, in any order (each is processed independently; which thread gets which cell changes nothing)
__global__ void __launch_bounds__(256) k_flag_compact(long long ncell, const unsigned char* __restrict__ inflag, int* __restrict__ flist, int* __restrict__ nflag) {
  const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool f = L < ncell && inflag[L];
  const unsigned b = __ballot_sync(FULL, f);
  if (!b) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == 0) base = atomicAdd(nflag, __popc(b));
  base = __shfl_sync(FULL, base, 0);
  if (f) flist[base + __popc(b & ((1u << lane) - 1u))] = (int)L;
}

// flist == nullptr: every physical cell; else the cells flist[0 .. *nflag)
template <int MINB, class F>
__global__ void __launch_bounds__(DC_T, MINB) k_drift_count(Geom g, DriftCountArgs<F> A, int heavy, const int* __restrict__ flist, const int* __restrict__ nflag,
                                                            int* __restrict__ rhoc_new, float* __restrict__ vfield_new) {
  const double weight_v = (double)0.1f;  // update_particle.f90:10
  __shared__ unsigned s_hm[DC_T / 32];
  __shared__ double s_v[DC_T / 32][3][32];  // a warp's accepted velocities of the current 32 candidates (crowded path)
  __shared__ int s_dest[DC_T];
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nslot = flist ? (long long)*nflag : g.ncell_p;
  if ((long long)blockIdx.x * blockDim.x >= nslot) return;
  const long long L = slot < nslot ? (flist ? (long long)flist[slot] : slot) : g.ncell_p;
  s_dest[threadIdx.x] = (int)L;
  bool crowded = false;
  if (L < g.ncell_p) {
    int tx, ty, tz, i, j, k;
    phys_decompose(g, L, tx, ty, tz, i, j, k);
    const int X0 = tx * g.nt, Y0 = ty * g.nt, Z0 = tz * g.nt;
    const int flags = dest_flags(g, A.farblk, A.r, X0 + i + NCB, Y0 + j + NCB, Z0 + k + NCB);
    const int r = (flags & FLAG_FAR) ? A.r : min(A.r, 1);
    // candidates (only counted when a crowded cell is near)
    int cand = 0;
    if (flags & FLAG_CROWD)
    for (int sk = k - r; sk <= k + r; sk++)
      for (int sj = j - r; sj <= j + r; sj++) {
        long long e = ext_index(g, X0 + i - r, Y0 + sj, Z0 + sk);
        for (int si = i - r; si <= i + r; si++, e++) {
          const int ddx = i - si, ddy = j - sj, ddz = k - sk;
          if (mask_hit(A.mask_e + MASK_W * e, ddx, ddy, ddz)) cand += A.rhoc_e[e];
        }
      }
    crowded = cand > heavy;
    if (!crowded) {
      int cnt = 0;
      const long long e0 = ext_index(g, X0 + i, Y0 + j, Z0 + k);
      float vfn0 = (float)__dmul_rn((double)A.vfield_e[3 * e0], weight_v);  // :27
      float vfn1 = (float)__dmul_rn((double)A.vfield_e[3 * e0 + 1], weight_v);
      float vfn2 = (float)__dmul_rn((double)A.vfield_e[3 * e0 + 2], weight_v);
      for (PlaneOrder po(k, r, A.nlayer); po.valid(); po.next())
        for (int sj = j - r, sk = po.sk; sj <= j + r; sj++) {
          long long e = ext_index(g, X0 + i - r, Y0 + sj, Z0 + sk);
          for (int si = i - r; si <= i + r; si++, e++) {
            const int ddx = i - si, ddy = j - sj, ddz = k - sk;
            if (!mask_hit(A.mask_e + MASK_W * e, ddx, ddy, ddz)) continue;  // nobody in this source cell comes my way
            const int n = A.rhoc_e[e];
            const long long s = A.cstart_e[e];
            const unsigned want = key_pack(ddx, ddy, ddz), o12 = off_pack12(ddx, ddy, ddz) << RANK_BITS;
            const double vf0 = A.vfield_e[3 * e], vf1 = A.vfield_e[3 * e + 1], vf2 = A.vfield_e[3 * e + 2];
            for (int l = 0; l < n; l++) {
              double v0, v1, v2;
              if (drift_accept(A, s + l, want, si, sj, sk, i, j, k, vf0, vf1, vf2, v0, v1, v2)) {
                A.rank[s + l] = (unsigned)cnt | o12;
                cnt++;
                vfn0 = (float)__dadd_rn((double)vfn0, v0);  // :47, f32 store after each f64 add
                vfn1 = (float)__dadd_rn((double)vfn1, v1);
                vfn2 = (float)__dadd_rn((double)vfn2, v2);
              }
            }
          }
        }
      const double den = __dadd_rn((double)cnt, weight_v);  // :55-57
      rhoc_new[L] = cnt;
      vfield_new[3 * L] = (float)((double)vfn0 / den); vfield_new[3 * L + 1] = (float)((double)vfn1 / den); vfield_new[3 * L + 2] = (float)((double)vfn2 / den);
    }
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  {
    const unsigned hm = __ballot_sync(FULL, crowded);
    if (lane == 0) s_hm[wp] = hm;
  }
  __syncthreads();
  int ord = 0;
  for (int wd = 0; wd < DC_T / 32; wd++) {
    unsigned bits = s_hm[wd];
    while (bits) {
      const long long D = s_dest[wd * 32 + __ffs(bits) - 1];
      bits &= bits - 1;
      if ((ord++ & (DC_T / 32 - 1)) != wp) continue;
      int tx, ty, tz, i, j, k;
      phys_decompose(g, D, tx, ty, tz, i, j, k);
      const int X0 = tx * g.nt, Y0 = ty * g.nt, Z0 = tz * g.nt;
      const int r = (dest_flags(g, A.farblk, A.r, X0 + i + NCB, Y0 + j + NCB, Z0 + k + NCB) & FLAG_FAR) ? A.r : min(A.r, 1);
      int cnt = 0;
      const long long e0 = ext_index(g, X0 + i, Y0 + j, Z0 + k);
      // the three components' chains run side by side in lanes 0, 1, 2 (the other lanes repeat them): an instruction costs a
      // warp the same for one lane as for 32, so this is a third of the chain instructions of "every lane all three"
      const int dcomp = lane % 3;
      float vfn = (float)__dmul_rn((double)A.vfield_e[3 * e0 + dcomp], weight_v);
      double(*sv)[32] = s_v[wp];
      for (PlaneOrder po(k, r, A.nlayer); po.valid(); po.next())
        for (int sj = j - r, sk = po.sk; sj <= j + r; sj++) {
          long long e = ext_index(g, X0 + i - r, Y0 + sj, Z0 + sk);
          for (int si = i - r; si <= i + r; si++, e++) {
            const int ddx = i - si, ddy = j - sj, ddz = k - sk;
            if (!mask_hit(A.mask_e + MASK_W * e, ddx, ddy, ddz)) continue;
            const int n = A.rhoc_e[e];
            const long long s = A.cstart_e[e];
            const unsigned want = key_pack(ddx, ddy, ddz), o12 = off_pack12(ddx, ddy, ddz) << RANK_BITS;
            const double vf0 = A.vfield_e[3 * e], vf1 = A.vfield_e[3 * e + 1], vf2 = A.vfield_e[3 * e + 2];
            for (int base = 0; base < n; base += 32) {
              double v0 = 0, v1 = 0, v2 = 0;
              const bool acc = base + lane < n && drift_accept(A, s + base + lane, want, si, sj, sk, i, j, k, vf0, vf1, vf2, v0, v1, v2);
              const unsigned b = __ballot_sync(FULL, acc);
              if (!b) continue;
              if (acc) {  // accepted velocities, compacted in storage order
                const int pos = __popc(b & ((1u << lane) - 1u));
                A.rank[s + base + lane] = (unsigned)(cnt + pos) | o12;
                sv[0][pos] = v0; sv[1][pos] = v1; sv[2][pos] = v2;
              }
              const int na = __popc(b);
              cnt += na;
              __syncwarp();
              for (int q = 0; q < na; q++) vfn = (float)__dadd_rn((double)vfn, sv[dcomp][q]);  // :47, f32 store after each f64 add
              __syncwarp();
            }
          }
        }
      const double den = __dadd_rn((double)cnt, weight_v);
      if (lane == 0) rhoc_new[D] = cnt;
      if (lane < 3) vfield_new[3 * D + lane] = (float)((double)vfn / den);
    }
  }
}
// sum over the cells of vfield_new^2 (std_vsim_c, update_particle.f90:133-136): one partial per 128 cells, fixed order
__global__ void __launch_bounds__(128) k_vfield_sq(long long ncell, const float* __restrict__ vfield_new, double* __restrict__ stc_partial) {
  const long long L = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double st_c = 0;
  if (L < ncell) {
    const float a = vfield_new[3 * L], b = vfield_new[3 * L + 1], c = vfield_new[3 * L + 2];
    st_c = (double)__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
  }
  __shared__ double sm[4];
  for (int o = 16; o; o >>= 1) st_c += __shfl_down_sync(0xffffffffu, st_c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = st_c;
  __syncthreads();
  if (threadIdx.x == 0) stc_partial[blockIdx.x] = ((sm[0] + sm[1]) + sm[2]) + sm[3];
}

// one particle: new codes at slot `pos` of the re-sorted arrays + its terms of the velocity statistics
template <class F>
__device__ __forceinline__ void drift_move(long long p, long long pos, const typename F::XT* __restrict__ xp, const typename F::VT* __restrict__ vp,
                                           const float* __restrict__ vf_src, const float* __restrict__ vf_new,
                                           const double* __restrict__ dvlut, const double* __restrict__ enc, double dt_mid, double S,
                                           typename F::XT* __restrict__ xp_new, typename F::VT* __restrict__ vp_new, double& st_tot, double& st_res) {
  constexpr double XSCALE = (double)(1 << (F::XB - 2));  // 1/(x_resolution*ncell): an exact scaling
  const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);
  const double v0 = __dadd_rn(dvlut[upat<F::VB>(vc.x)], (double)vf_src[0]);
  const double v1 = __dadd_rn(dvlut[upat<F::VB>(vc.y)], (double)vf_src[1]);
  const double v2 = __dadd_rn(dvlut[upat<F::VB>(vc.z)], (double)vf_src[2]);
  // xp_new=xp+nint(dt_mid*vreal/(x_resolution*ncell))  :84 (the store wraps to the code's width)
  const short x0 = (short)((int)xc.x + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v0), XSCALE)));
  const short x1 = (short)((int)xc.y + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v1), XSCALE)));
  const short x2 = (short)((int)xc.z + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v2), XSCALE)));
  const float n0 = vf_new[0], n1 = vf_new[1], n2 = vf_new[2];
  const short w0 = vp_encode_lut<F::VB>(__dsub_rn(v0, (double)n0), S, enc);  // :85-86
  const short w1 = vp_encode_lut<F::VB>(__dsub_rn(v1, (double)n1), S, enc);
  const short w2 = vp_encode_lut<F::VB>(__dsub_rn(v2, (double)n2), S, enc);
  store_code3(xp_new, pos, x0, x1, x2);
  store_code3(vp_new, pos, w0, w1, w2);
  // velocity statistics, update_particle.f90:140-143 (decoded with the old sigma_vi)
  double a0 = dvlut[upat<F::VB>(w0)], a1 = dvlut[upat<F::VB>(w1)], a2 = dvlut[upat<F::VB>(w2)];
  st_res += __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
  a0 = __dadd_rn(a0, (double)n0); a1 = __dadd_rn(a1, (double)n1); a2 = __dadd_rn(a2, (double)n2);
  st_tot += __dadd_rn(__dadd_rn(__dmul_rn(a0, a0), __dmul_rn(a1, a1)), __dmul_rn(a2, a2));
}

__device__ __forceinline__ void block_sum2(double a, double b, double* __restrict__ out2) {  // fixed-order block reduction
  __shared__ double sm[2][PC_T / 32];
  for (int o = 16; o; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0;
    for (int w = 0; w < PC_T / 32; w++) t += sm[threadIdx.x][w];
    out2[threadIdx.x] = t;
  }
}

// pass C: one thread per particle; single image: destinations wrap periodically.  stat_partial gets one (total, residual)
// pair per warp of the launch (fixed order for a fixed grid).
template <class F>
__global__ void __launch_bounds__(PW_T, 1) k_drift_place_w(Geom g, VTab vt, double S, const typename F::XT* __restrict__ xp, const typename F::VT* __restrict__ vp,
                                                          const unsigned* __restrict__ rank, const long long* __restrict__ cstart_p,
                                                          const float* __restrict__ vfield_p, const long long* __restrict__ cstart_new,
                                                          const float* __restrict__ vfield_new, double dt_mid, typename F::XT* __restrict__ xp_new,
                                                          typename F::VT* __restrict__ vp_new, double* __restrict__ stat_partial) {
  constexpr double XSCALE = (double)(1 << (F::XB - 2));  // 1/(x_resolution*ncell): an exact scaling
  extern __shared__ __align__(16) unsigned char pw_smem[];
  float* s_tan = reinterpret_cast<float*>(pw_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(pw_smem + vt.hot * 4) + warp;
  if (vt.hot) fill_tab(s_tan, vt.tanh, vt.hot);
  const VDec dec = make_dec(vt, s_tan, S);
  __syncthreads();
  double s_t2 = 0, s_tn = 0, s_n2 = 0;
  for (long long c0 = ((long long)blockIdx.x * PW_W + warp) * WC; c0 < g.ncell_p; c0 += (long long)gridDim.x * PW_W * WC) {
    long long p0;
    const int np = warp_chunk_setup(g, cstart_p, c0, g.ncell_p, ws, lane, p0);
    for (int q = lane; q < np; q += 32) {
      const long long p = p0 + q;
      const unsigned rk = rank[p];
      const Code3 xc = load_code3(xp, p), vc = load_code3(vp, p);  // issued with the rank load: one memory latency, not two
      if (rk == RANK_LOST) continue;
      const int cl = warp_chunk_find(ws, q);
      const long long L = c0 + cl;
      const CellPos cp = ws->pos[cl];
      const unsigned o = rk >> RANK_BITS;
      long long D = L;
      if (o != off_pack12(0, 0, 0)) {  // a mover (a few per cent): destination = source + offset
        if (!dest_cell(g, cp, (int)(o & 15u) - 8, (int)((o >> 4) & 15u) - 8, (int)((o >> 8) & 15u) - 8, D)) continue;  // cannot happen for an accepted particle
      }
      const long long pos = cstart_new[D] + (rk & ((1u << RANK_BITS) - 1));
      const float* vf_src = vfield_p + 3 * L;
      const float* vf_new = vfield_new + 3 * D;
      const double v0 = __dadd_rn(v_decode<F::VB>(dec, vc.x), (double)vf_src[0]);
      const double v1 = __dadd_rn(v_decode<F::VB>(dec, vc.y), (double)vf_src[1]);
      const double v2 = __dadd_rn(v_decode<F::VB>(dec, vc.z), (double)vf_src[2]);
      // xp_new=xp+nint(dt_mid*vreal/(x_resolution*ncell))  :84 (the store wraps to the code's width)
      const short x0 = (short)((int)xc.x + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v0), XSCALE)));
      const short x1 = (short)((int)xc.y + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v1), XSCALE)));
      const short x2 = (short)((int)xc.z + (int)llround(__dmul_rn(__dmul_rn(dt_mid, v2), XSCALE)));
      const float n0 = vf_new[0], n1 = vf_new[1], n2 = vf_new[2];
      const short w0 = vp_encode_lut<F::VB>(__dsub_rn(v0, (double)n0), S, vt.thr);  // :85-86
      const short w1 = vp_encode_lut<F::VB>(__dsub_rn(v1, (double)n1), S, vt.thr);
      const short w2 = vp_encode_lut<F::VB>(__dsub_rn(v2, (double)n2), S, vt.thr);
      store_code3(xp_new, pos, x0, x1, x2);
      store_code3(vp_new, pos, w0, w1, w2);
      // velocity statistics, update_particle.f90:140-143 (the new codes decoded with the old sigma_vi: dv = t/S, t the host tanf
      // value).  sum dv^2 = (sum t^2)/S^2 and sum (dv+n)^2 = sum dv^2 + 2 (sum t n)/S + sum n^2: the sums of t^2 (exact products in
      // f64), t*n and n^2 are taken here and scaled once per warp -- equal to the term-by-term form to a few 1e-16, which is the
      // size of its own rounding; the statistics are only ever used rounded to f32.
      const double t0 = (double)tan_value<F::VB>(dec, vt.tanlut, w0), t1 = (double)tan_value<F::VB>(dec, vt.tanlut, w1),
                   t2 = (double)tan_value<F::VB>(dec, vt.tanlut, w2);
      s_t2 = __fma_rn(t0, t0, __fma_rn(t1, t1, __fma_rn(t2, t2, s_t2)));
      s_tn = __fma_rn(t0, (double)n0, __fma_rn(t1, (double)n1, __fma_rn(t2, (double)n2, s_tn)));
      s_n2 = __fma_rn((double)n0, (double)n0, __fma_rn((double)n1, (double)n1, __fma_rn((double)n2, (double)n2, s_n2)));
    }
  }
  for (int o = 16; o; o >>= 1) {
    s_t2 += __shfl_down_sync(FULL, s_t2, o); s_tn += __shfl_down_sync(FULL, s_tn, o); s_n2 += __shfl_down_sync(FULL, s_n2, o);
  }
  const double st_res = s_t2 / (S * S), st_tot = st_res + 2.0 * s_tn / S + s_n2;
  if (lane == 0) {
    stat_partial[2 * ((long long)blockIdx.x * PW_W + warp)] = st_tot;
    stat_partial[2 * ((long long)blockIdx.x * PW_W + warp) + 1] = st_res;
  }
}

// pass C for the ghost particles that enter this image
template <class F>
__global__ void __launch_bounds__(PC_T) k_drift_place_g(Geom g, long long ng, const int* __restrict__ gcell_ext, const long long* __restrict__ gstart,
                                                       long long base, const typename F::XT* __restrict__ xp, const typename F::VT* __restrict__ vp,
                                                       const unsigned* __restrict__ rank, const float* __restrict__ vfield_e,
                                                       const long long* __restrict__ cstart_new, const float* __restrict__ vfield_new,
                                                       const double* __restrict__ dvlut, const double* __restrict__ enc, double dt_mid, double S,
                                                       typename F::XT* __restrict__ xp_new, typename F::VT* __restrict__ vp_new, double* __restrict__ stat_partial) {
  __shared__ int soff[PC_CELLS + 1];
  const long long c0 = (long long)blockIdx.x * PC_CELLS;
  const int np = chunk_setup(gstart, c0, ng, soff);
  const long long p0 = base + gstart[c0];
  double st_tot = 0, st_res = 0;
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const long long p = p0 + q;
    const unsigned rk = rank[p];
    if (rk == RANK_LOST) continue;
    const long long e = gcell_ext[c0 + chunk_find(soff, q)];
    const unsigned o = rk >> RANK_BITS;
    const int X = (int)(e % g.ne) - NCB + (int)(o & 15u) - 8, Y = (int)((e / g.ne) % g.ne) - NCB + (int)((o >> 4) & 15u) - 8,
              Z = (int)(e / ((long long)g.ne * g.ne)) - NCB + (int)((o >> 8) & 15u) - 8;
    if ((unsigned)X >= (unsigned)g.nc || (unsigned)Y >= (unsigned)g.nc || (unsigned)Z >= (unsigned)g.nc) continue;  // cannot happen for an accepted particle
    const long long D = phys_index(g, X / g.nt, Y / g.nt, Z / g.nt, X % g.nt, Y % g.nt, Z % g.nt);
    drift_move<F>(p, cstart_new[D] + (rk & ((1u << RANK_BITS) - 1)), xp, vp, vfield_e + 3 * e, vfield_new + 3 * D, dvlut, enc, dt_mid, S, xp_new,
               vp_new, st_tot, st_res);
  }
  block_sum2(st_tot, st_res, stat_partial + 2 * (long long)blockIdx.x);
}

// fixed-order final reduction of n partial sums laid out with stride `stride`, component `comp`
__global__ void __launch_bounds__(1024) k_reduce_strided(const double* __restrict__ part, long long n, int stride, int comp, double* __restrict__ out) {
  __shared__ double sm[32];
  double s = 0;
  for (long long b = threadIdx.x; b < n; b += blockDim.x) s += part[b * stride + comp];
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 32; w++) t += sm[w]; *out = t; }
}

// -DPID (update_particle.f90:88 `pid_new(idx)=pid(ip)`): the particle IDs follow the permutation of pass C.  IDs are optional and off
// the hot path: one thread per source cell, the destination arithmetic of pass C (dest_cell: periodic wrap for nn_d == 1, particles
// that leave the image are another image's).
__global__ void __launch_bounds__(256) k_pid_place(Geom g, const long long* __restrict__ pid, const unsigned* __restrict__ rank,
                                                   const long long* __restrict__ cstart_p, const long long* __restrict__ cstart_new,
                                                   long long* __restrict__ pid_new) {
  const long long L = (long long)blockIdx.x * 256 + threadIdx.x;
  if (L >= g.ncell_p) return;
  int tx0, ty0, tz0, i0, j0, k0;
  phys_decompose(g, L, tx0, ty0, tz0, i0, j0, k0);
  CellPos cp = {(short)tx0, (short)ty0, (short)tz0, (short)i0, (short)j0, (short)k0, 0, 0};
  const long long pend = cstart_p[L + 1];
  for (long long p = cstart_p[L]; p < pend; p++) {
    const unsigned rk = rank[p];
    if (rk == RANK_LOST) continue;
    const unsigned o = rk >> RANK_BITS;
    long long D;
    if (!dest_cell(g, cp, (int)(o & 15u) - 8, (int)((o >> 4) & 15u) - 8, (int)((o >> 8) & 15u) - 8, D)) continue;
    pid_new[cstart_new[D] + (rk & ((1u << RANK_BITS) - 1))] = pid[p];
  }
}
// the same for the ghost particles received from other images (buffer_v.f90:23,42,... carries pid with vp): one thread per ghost cell
__global__ void __launch_bounds__(256) k_pid_place_g(Geom g, long long ng, const int* __restrict__ gcell_ext, const long long* __restrict__ gstart,
                                                     long long base, const long long* __restrict__ pid, const unsigned* __restrict__ rank,
                                                     const long long* __restrict__ cstart_new, long long* __restrict__ pid_new) {
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  if (q >= ng) return;
  const long long e = gcell_ext[q];
  const int x0 = (int)(e % g.ne) - NCB, y0 = (int)((e / g.ne) % g.ne) - NCB, z0 = (int)(e / ((long long)g.ne * g.ne)) - NCB;
  const long long pend = base + gstart[q + 1];
  for (long long p = base + gstart[q]; p < pend; p++) {
    const unsigned rk = rank[p];
    if (rk == RANK_LOST) continue;
    const unsigned o = rk >> RANK_BITS;
    const int X = x0 + (int)(o & 15u) - 8, Y = y0 + (int)((o >> 4) & 15u) - 8, Z = z0 + (int)((o >> 8) & 15u) - 8;
    if ((unsigned)X >= (unsigned)g.nc || (unsigned)Y >= (unsigned)g.nc || (unsigned)Z >= (unsigned)g.nc) continue;
    const long long D = phys_index(g, X / g.nt, Y / g.nt, Z / g.nt, X % g.nt, Y % g.nt, Z % g.nt);
    pid_new[cstart_new[D] + (rk & ((1u << RANK_BITS) - 1))] = pid[p];
  }
}

}  // namespace cube
