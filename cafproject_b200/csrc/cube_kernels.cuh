// cube_kernels.cuh -- sm_100a kernels of the CUBE particle-mesh step.
//
// All particle kernels exploit CUBE's cell-ordered storage: a coarse cell's particles are the
// contiguous run [cstart, cstart+rhoc) of the AoS int16 arrays.  Every mesh value is produced by
// exactly one thread that pulls ("gathers") its contributions in the reference's traversal order
// (tile-local k, j, i, then storage order inside a cell), so there are no float atomics and the
// results do not depend on scheduling; for the densities they are bit-identical to the reference's
// sequential scatter loop.
#pragma once
#include <climits>
#include "cube_common.cuh"

namespace cube {

// =============================================================================================
// scan: exclusive prefix sum int32 -> int64 (cumsum6 of variables.f90:90-110, in file order)
// =============================================================================================
constexpr int SCAN_T = 256, SCAN_I = 8, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* smem /*>=33*/, long long& total) {
  // inclusive warp scan
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[w] = x;
  __syncthreads();
  if (w == 0) {
    long long s = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    smem[lane] = s;  // inclusive over warps
  }
  __syncthreads();
  long long woff = w ? smem[w - 1] : 0;
  total = smem[(blockDim.x >> 5) - 1];
  __syncthreads();
  return woff + x - v;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_blocksum(const int* __restrict__ in, long long n, long long* __restrict__ bsum) {
  __shared__ long long sm[33];
  long long base = (long long)blockIdx.x * SCAN_B + (long long)threadIdx.x * SCAN_I;
  long long s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_I; q++) if (base + q < n) s += in[base + q];
  long long tot;
  block_exclusive_scan(s, sm, tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
// single block: exclusive scan of the block sums in place; total -> bsum[nb]
__global__ void __launch_bounds__(1024) k_scan_bsums(long long* bsum, int nb) {
  __shared__ long long sm[33];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += blockDim.x) {
    int q = b0 + threadIdx.x;
    long long v = q < nb ? bsum[q] : 0, tot;
    long long ex = block_exclusive_scan(v, sm, tot);
    if (q < nb) bsum[q] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}
__global__ void __launch_bounds__(SCAN_T) k_scan_final(const int* __restrict__ in, long long n, const long long* __restrict__ bsum,
                                                       long long* __restrict__ out) {
  __shared__ long long sm[33];
  long long base = (long long)blockIdx.x * SCAN_B + (long long)threadIdx.x * SCAN_I;
  int v[SCAN_I];
  long long s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_I; q++) { v[q] = (base + q < n) ? in[base + q] : 0; s += v[q]; }
  long long tot;
  long long ex = block_exclusive_scan(s, sm, tot) + bsum[blockIdx.x];
#pragma unroll
  for (int q = 0; q < SCAN_I; q++) { if (base + q < n) out[base + q] = ex; ex += v[q]; }
}

// =============================================================================================
// buffered state: extended image grid (buffer_density.f90 for one image: ghost layers alias the
// periodic image, so ghost particles are never copied)
// =============================================================================================
__global__ void k_build_ext(Geom g, const int* __restrict__ rhoc_p, const long long* __restrict__ cstart_p,
                            const float* __restrict__ vfield_p, int* __restrict__ rhoc_e, long long* __restrict__ cstart_e,
                            float* __restrict__ vfield_e, int* __restrict__ sid_e) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.ncell_e) return;
  int x = (int)(e % g.ne) - NCB, y = (int)((e / g.ne) % g.ne) - NCB, z = (int)(e / ((long long)g.ne * g.ne)) - NCB;
  if (ext_is_remote(g, x, y, z)) return;  // owned by another image: filled by the halo exchange (cube_exchange.cuh)
  // nn_d == 1: neighbour image is this image (inx=ipx=icx), parameters.f90:189-194
  int xw = (x + g.nc) % g.nc, yw = (y + g.nc) % g.nc, zw = (z + g.nc) % g.nc;
  long long L = phys_index(g, xw / g.nt, yw / g.nt, zw / g.nt, xw % g.nt, yw % g.nt, zw % g.nt);
  rhoc_e[e] = rhoc_p[L];
  sid_e[e] = (int)L;
  cstart_e[e] = cstart_p[L];
  vfield_e[3 * e + 0] = vfield_p[3 * L + 0];
  vfield_e[3 * e + 1] = vfield_p[3 * L + 1];
  vfield_e[3 * e + 2] = vfield_p[3 * L + 2];
}

// particles in each tile's extended region (cume(nt+2ncb,...) of update_particle.f90:60); grid = (splits, tiles),
// tile_count zeroed by the caller (integer atomics: order-independent)
__global__ void __launch_bounds__(256) k_tile_counts(Geom g, const int* __restrict__ rhoc_e, unsigned long long* __restrict__ tile_count) {
  __shared__ long long sm[33];
  const int t = blockIdx.y;
  const int tx = t % g.nnt, ty = (t / g.nnt) % g.nnt, tz = t / (g.nnt * g.nnt);
  const long long nrow = (long long)g.nte * g.nte;
  long long s = 0;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < nrow; row += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int j = (int)(row % g.nte), k = (int)(row / g.nte);
    const int* r = rhoc_e + ext_index(g, tx * g.nt - NCB, ty * g.nt + j - NCB, tz * g.nt + k - NCB);
    for (int i = threadIdx.x & 31; i < g.nte; i += 32) s += r[i];
  }
  long long tot;
  block_exclusive_scan(s, sm, tot);
  if (threadIdx.x == 0 && tot) atomicAdd(&tile_count[t], (unsigned long long)tot);
}

// dv table: dble(tan((pi*real(vp))/real(nvbin-1))) / (sqrt(pi/2)/(sigma_vi*vrel_boost))
// Also checks, entry by entry, that t/S equals its FMA form q0 = t*rS, q = fma(fma(-S,q0,t), rS, q0) (rS = 1/S): the
// particle kernels then divide that way (cube_particles.cuh, v_decode); *divok is cleared on any mismatch.
__global__ void k_build_dvlut(int ncode /* nvbin */, const float* __restrict__ tanlut, double S, double rS, double* __restrict__ dvlut,
                              int* __restrict__ divok) {
  int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= ncode) return;
  const double t = (double)tanlut[u], q = t / S;
  dvlut[u] = q;
  const double q0 = __dmul_rn(t, rS);
  const double qf = __fma_rn(__fma_rn(-S, q0, t), rS, q0);
  if (__double_as_longlong(qf) != __double_as_longlong(q)) atomicExch(divok, 0);
}

// fixed-order final reduction of per-block partial sums (3 interleaved series)
__global__ void __launch_bounds__(1024) k_reduce3(const double* __restrict__ part, long long nb, double* __restrict__ out) {
  __shared__ double sm[3][32];
  double s[3] = {0, 0, 0};
  for (long long b = threadIdx.x; b < nb; b += blockDim.x) {
    s[0] += part[3 * b]; s[1] += part[3 * b + 1]; s[2] += part[3 * b + 2];
  }
  for (int c = 0; c < 3; c++) {
    for (int o = 16; o; o >>= 1) s[c] += __shfl_down_sync(0xffffffffu, s[c], o);
    if ((threadIdx.x & 31) == 0) sm[c][threadIdx.x >> 5] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sm[threadIdx.x][w];
    out[threadIdx.x] = t;
  }
}

// =============================================================================================
// fine mesh (pm.f90:44-118)
// =============================================================================================

// tempx=4.*((/i,j,k/)-1)+4*(int(xp+ishift,izipx)+rshift)*x_resolution, rounded to f32 (pm.f90:54).
// = (2K+1)/2^(XB-1) with K = 2^XB*(cell-1) + u an integer (XB = 8 izipx bits): the f64 expression is exact, so its f32 rounding
// equals the round-to-nearest int->float conversion of 2K+1 scaled by a power of two (no f64 instructions)
template <int XB> __device__ __forceinline__ float fine_tempx(int cell1, short xp) {
  return __int2float_rn(2 * ((1 << XB) * (cell1 - 1) + (int)upat<XB>(xp)) + 1) * (1.0f / (float)(1 << (XB - 1)));
}

// ---------------------------------------------------------------------------------------------
// Fine CIC deposit (pm.f90:44-72) on a REGION of the image's fine grid.
//
// The reference deposits every tile's particles (cells 2-ncb..nt+ncb-1) into that tile's padded grid rho_f.  A fine node's
// value is the sum of the terms of ALL particles within one fine cell of it, whichever tile's grid it is read from, and for
// tile-local coordinates below 512 fine cells (cells up to nt + 5 <= 128) every f32 step of pm.f90:54-58 is exact: the weights are
//     upper node: h = (2 (u mod 2^(XB-2)) + 1) 2^-(XB-1),  lower node: 1 - h,   lower node index = 4 cell + (u >> (XB-2))
// (u = raw XB-bit position code), independent of the tile frame.  So one deposit of the image's cells onto one grid that spans
// a whole batch of tiles serves every tile of the batch: the FFT's x pass reads each tile's window out of it, and the 1.42x
// of work that overlapping windows cost disappears.  (Larger tiles are deposited one by one in their own frame, with the
// reference's f32 rounding of tempx reproduced: `frame0` = image-local cell index of the tile's cell 1.)
//
// One CTA = one brick of BX x BY x BZ coarse cells of output = 4BX x 4BY x 4BZ nodes held in shared memory as 32-bit
// FIXED-POINT accumulators; one thread per particle of the brick's (BX+1)(BY+1)(BZ+1) source cells (its own cells plus the
// low-side neighbour layer whose upper nodes land in the brick), native 32-bit integer shared-memory atomics.  Integer sums do
// not depend on the order in which the hardware serialises them: deterministic, bit-identical run to run, and equal to the
// reference's sequential f32 scatter to round-off.  The scale 2^S is chosen per brick from its fullest source cell so that no
// node can overflow (a node receives terms from at most 8 coarse cells, each term <= mass_p): S = 21 for the uniform z = 49
// state (resolution 5e-7 of a mass unit, per-term rounding 3e-8 of the mean node), smaller inside haloes where the f32 sum it
// replaces has long lost those bits.  Accumulator layout x + 36 y + (36*4BY+16) z: the 4x4x4 nodes of a coarse cell fall on 32
// different banks (a dense x + 32 y + .. layout puts them on 4).
// ---------------------------------------------------------------------------------------------
struct FineRegion {
  int c0[3], c1[3];    // source coarse cells [c0, c1) per dimension, image-local 0-based (ghost layers: < 0 or >= nc)
  int f0[3];           // image-local fine node index of output element (0,0,0), a multiple of 4
  int n[3];            // output extent in nodes
  long long ldy, ldz;  // output pitches
};
constexpr int FRAME_NONE = INT_MIN;

template <int BX, int BY, int BZ, int NT_>
struct FdCfg {
  static constexpr int NT = NT_;
  static constexpr int SX = BX + 1, SY = BY + 1, SZ = BZ + 1, NS = SX * SY * SZ;
  static constexpr int NX = 4 * BX, NY = 4 * BY, NZ = 4 * BZ;
  static constexpr int PY = NX + 4, PZ = PY * NY + 16;
  static constexpr int P2 = NS < 256 ? 256 : NS < 512 ? 512 : 1024;
  static constexpr int ACC = PZ * NZ;  // words
  static constexpr int CAP = NS <= 256 ? 4096 : NS <= 512 ? 8192 : 16384;  // particle slots of the cell-id list (~2.5x the mean brick)
  static constexpr int EXPAND_MAX = 64;  // bricks whose fullest cell holds more find each particle's cell by binary search instead
  static constexpr size_t SMEM = (size_t)ACC * 4 + (size_t)P2 * 4 + (size_t)NS * 8 + (size_t)CAP * 2 + 64;
  static_assert(NX == 32, "a brick row is one warp-wide 128-byte store");
  static_assert(PZ % 32 == 16 && PY % 32 == 4, "bank layout");
  static_assert(NS < P2 && NS <= NT_, "one thread per source cell in the set-up");
};

// lower node (image-local fine index) and the two weights of one coordinate
// HALF: the cell-centred assignment of the power-spectrum estimator (cicpower.f90:84-90: pos*ng/nc - 0.5), in a node frame shifted by
// one so that a cell still only reaches its own nodes and the next cell's: returned L = (true lower node) + 1 = 4 cell + ((u + 2^(XB-3)) >> (XB-2))
template <int XB, bool HALF = false> __device__ __forceinline__ void fine_cic(int cell, short xp, int frame0, int& L, float& w0, float& w1) {
  if (frame0 == FRAME_NONE) {
    const unsigned u = upat<XB>(xp) + (HALF ? (1u << (XB - 3)) : 0u);
    L = 4 * cell + (int)(u >> (XB - 2));
    w1 = __int2float_rn(2 * (int)(u & ((1u << (XB - 2)) - 1u)) + 1) * (1.0f / (float)(1 << (XB - 1)));
    w0 = 1.0f - w1;  // exact
  } else {  // the tile's own frame: tempx may round (pm.f90:54 in f32 beyond 512 fine cells)
    int idx1;
    cic_split(fine_tempx<XB>(cell - frame0 + 1, xp), idx1, w0, w1);
    L = 4 * frame0 + idx1 - 1;
  }
}

// ACC: add to `out` instead of overwriting it (a further species of a two-species run, pm.f90:79-99 NEUTRINOS): the lines already there
// are prefetched to L2 before the particle loop and the write-out reads four rows ahead of its stores
template <class C, bool FRAME, class XT, bool HALF = false, bool ACC = false>
__global__ void __launch_bounds__(C::NT) k_fine_deposit_r(Geom g, FineRegion R, int3 frame0, const XT* __restrict__ xp,
                                                          const int* __restrict__ rhoc_e, const long long* __restrict__ cstart_e,
                                                          float mass_p, float* __restrict__ out,
                                                          int part = 0 /* 1: only bricks that read no ghost cell, 2: only the others */) {
  constexpr int NIT = C::NX * C::NY * C::NZ / C::NT;
  static_assert(C::NX * C::NY * C::NZ % C::NT == 0, "write-out loop");
  extern __shared__ __align__(16) unsigned fd_smem[];
  unsigned* acc = fd_smem;                                             // [NZ][PZ]
  int* pref = reinterpret_cast<int*>(fd_smem + C::ACC);                // [P2]
  long long* start = reinterpret_cast<long long*>(pref + C::P2);       // [NS]
  unsigned short* cid = reinterpret_cast<unsigned short*>(start + C::NS);  // [CAP] source cell of every particle slot (sparse bricks)
  __shared__ int s_w[32], s_m[32];
  const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
  const int nbx = (R.n[0] + C::NX - 1) / C::NX, nby = (R.n[1] + C::NY - 1) / C::NY;
  const int bx = blockIdx.x % nbx, by = (blockIdx.x / nbx) % nby, bz = blockIdx.x / (nbx * nby);
  // image-local node index of the brick's first node; its first own coarse cell
  const int N0x = R.f0[0] + bx * C::NX, N0y = R.f0[1] + by * C::NY, N0z = R.f0[2] + bz * C::NZ;
  const int cbx = N0x >> 2, cby = N0y >> 2, cbz = N0z >> 2;
  if (part) {  // several images: the bricks away from the image boundary run while the ghost positions are still on the wire
    const bool edge = cbx < 1 || cbx - 1 + C::SX > g.nc || cby < 1 || cby - 1 + C::SY > g.nc || cbz < 1 || cbz - 1 + C::SZ > g.nc;
    if (edge != (part == 2)) return;
  }
  // --- the brick's source cells: counts, starts, prefix, fullest cell
  int n = 0; long long s = 0;
  if (t < C::NS) {
    const int sx = t % C::SX, sy = (t / C::SX) % C::SY, sz = t / (C::SX * C::SY);
    const int cx = cbx - 1 + sx, cy = cby - 1 + sy, cz = cbz - 1 + sz;
    if (cx >= R.c0[0] && cx < R.c1[0] && cy >= R.c0[1] && cy < R.c1[1] && cz >= R.c0[2] && cz < R.c1[2]) {
      const long long e = ext_index(g, cx, cy, cz);
      n = rhoc_e[e]; s = cstart_e[e];
    }
    start[t] = s;
  }
  int incl = n, mx = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 31) { s_w[wp] = incl; s_m[wp] = mx; }
  __syncthreads();
  int woff = 0, total = 0, cmax = 0;
#pragma unroll
  for (int q = 0; q < C::NT / 32; q++) { const int v = s_w[q]; if (q < wp) woff += v; total += v; cmax = max(cmax, s_m[q]); }
  const int ox = bx * C::NX, oy = by * C::NY, oz = bz * C::NZ;  // brick offset inside the region
  if (total == 0) {  // empty brick: zeros, no accumulators
    if (!ACC)
      for (int o = t; o < C::NX * C::NY * C::NZ; o += C::NT) {
        const int X = o % C::NX, Y = (o / C::NX) % C::NY, Z = o / (C::NX * C::NY);
        if (ox + X < R.n[0] && oy + Y < R.n[1] && oz + Z < R.n[2]) out[(long long)(oz + Z) * R.ldz + (long long)(oy + Y) * R.ldy + ox + X] = 0.f;
      }
    return;
  }
  if (ACC)
    for (int r = t; r < 2 * C::NY * C::NZ; r += C::NT) {  // both ends of every row: a row of 32 floats touches at most two lines
      const int Y = (r >> 1) % C::NY, Z = (r >> 1) / C::NY, X = (r & 1) * (C::NX - 1);
      if (ox + X < R.n[0] && oy + Y < R.n[1] && oz + Z < R.n[2])
        asm volatile("prefetch.global.L2 [%0];" ::"l"(out + (long long)(oz + Z) * R.ldz + (long long)(oy + Y) * R.ldy + ox + X));
    }
  if (t < C::NS) pref[t] = woff + incl - n;
  for (int e = C::NS + t; e < C::P2; e += C::NT) pref[e] = total;  // sentinels: the search never steps onto them (q < total)
  for (int e = t; e < C::ACC / 4; e += C::NT) reinterpret_cast<uint4*>(acc)[e] = make_uint4(0u, 0u, 0u, 0u);
  // a node gets terms from the particles of at most 8 coarse cells, each term <= mass_p: 2^S mass_p 8 cmax <= 2^31, which leaves
  // 2^31 units for the half unit of rounding per term
  int S = 31 - (int)ceilf(log2f(__fmul_rn(__fmul_rn(mass_p, 8.0f), (float)cmax)));
  S = min(max(S, -32), 40);
  const float scale = __fmul_rn(mass_p, exp2f((float)S)), inv = exp2f(-(float)S);
  // Sparse bricks (the uniform state): every source cell's thread writes its index into the slots of its particles, one shared
  // store per particle instead of a ten-step search; bricks with a crowded cell keep the search (a thread would loop for long).
  const bool expand = cmax <= C::EXPAND_MAX && total <= C::CAP;
  if (expand && t < C::NS) {
    const int b0 = woff + incl - n;
    for (int k = 0; k < n; k++) cid[b0 + k] = (unsigned short)t;
  }
  __syncthreads();
  // --- one thread per particle
  for (int q = t; q < total; q += C::NT) {
    int c = 0;
    if (expand) c = cid[q];
    else {
#pragma unroll
      for (int step = C::P2 / 2; step > 0; step >>= 1)
        if (pref[c + step] <= q) c += step;  // largest c with pref[c] <= q (empty cells share their successor's offset)
    }
    const long long p = start[c] + (q - pref[c]);
    const Code3 cur = load_code3(xp, p);
    const int sx = c % C::SX, sy = (c / C::SX) % C::SY, sz = c / (C::SX * C::SY);
    int lx, ly, lz; float ax[2], ay[2], az[2];
    constexpr int XB = 8 * (int)sizeof(XT);
    fine_cic<XB, HALF>(cbx - 1 + sx, cur.x, FRAME ? frame0.x : FRAME_NONE, lx, ax[0], ax[1]);
    fine_cic<XB, HALF>(cby - 1 + sy, cur.y, FRAME ? frame0.y : FRAME_NONE, ly, ay[0], ay[1]);
    fine_cic<XB, HALF>(cbz - 1 + sz, cur.z, FRAME ? frame0.z : FRAME_NONE, lz, az[0], az[1]);
    lx -= N0x; ly -= N0y; lz -= N0z;  // brick-local lower node, -4 .. 4B-1 (the upper node is the next one)
    if (lx < -1 || ly < -1 || lz < -1) continue;  // low-side layer: only particles next to the brick reach into it
    const bool vx[2] = {(unsigned)lx < (unsigned)C::NX, (unsigned)(lx + 1) < (unsigned)C::NX};
    const bool vy[2] = {(unsigned)ly < (unsigned)C::NY, (unsigned)(ly + 1) < (unsigned)C::NY};
    const bool vz[2] = {(unsigned)lz < (unsigned)C::NZ, (unsigned)(lz + 1) < (unsigned)C::NZ};
    unsigned* a0 = acc + lz * C::PZ + ly * C::PY + lx;
#pragma unroll
    for (int qq = 0; qq < 8; qq++) {
      const int qa = qq & 1, qb = (qq >> 1) & 1, qc = qq >> 2;
      if (vx[qa] && vy[qb] && vz[qc]) {
        // dx(1)*dx(2)*dx(3)*mass_p (pm.f90:61-68) with the power-of-two scale riding on mass_p
        const unsigned wf = __float2uint_rn(__fmul_rn(__fmul_rn(__fmul_rn(ax[qa], ay[qb]), az[qc]), scale));
        atomicAdd(a0 + qc * C::PZ + qb * C::PY + qa, wf);
      }
    }
  }
  __syncthreads();
  if constexpr (!ACC) {
    for (int o = t; o < C::NX * C::NY * C::NZ; o += C::NT) {  // one 32-float row per warp and iteration; one writer per node
      const int X = o % C::NX, Y = (o / C::NX) % C::NY, Z = o / (C::NX * C::NY);
      if (ox + X < R.n[0] && oy + Y < R.n[1] && oz + Z < R.n[2])
        out[(long long)(oz + Z) * R.ldz + (long long)(oy + Y) * R.ldy + ox + X] = __fmul_rn(__uint2float_rn(acc[Z * C::PZ + Y * C::PY + X]), inv);
    }
  } else {
    constexpr int AH = 4;  // rows whose old values are read ahead of the stores
    static_assert(NIT % AH == 0, "write-out loop");
#pragma unroll 1
    for (int it = 0; it < NIT; it += AH) {
      float* dst[AH];
      float prev[AH];
#pragma unroll
      for (int u = 0; u < AH; u++) {
        const int o = t + (it + u) * C::NT;
        const int X = o % C::NX, Y = (o / C::NX) % C::NY, Z = o / (C::NX * C::NY);
        dst[u] = (ox + X < R.n[0] && oy + Y < R.n[1] && oz + Z < R.n[2]) ? out + (long long)(oz + Z) * R.ldz + (long long)(oy + Y) * R.ldy + ox + X : nullptr;
        prev[u] = dst[u] ? *dst[u] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < AH; u++) {
        const int o = t + (it + u) * C::NT;
        const int X = o % C::NX, Y = (o / C::NX) % C::NY, Z = o / (C::NX * C::NY);
        if (dst[u]) *dst[u] = __fadd_rn(prev[u], __fmul_rn(__uint2float_rn(acc[Z * C::PZ + Y * C::PY + X]), inv));
      }
    }
  }
}

// k-space: F_d = i * kern_d * rho_k (pm.f90:79-80), with the 1/nfe^3 of pm.f90:82 folded in.
// One thread per k-space element, looping over the tiles of the batch so that kern is read once.
__global__ void __launch_bounds__(256) k_green(long long nk, int nbatch, const float2* __restrict__ crho,
                                               const float* __restrict__ kern /*[3][nk]*/, float scale,
                                               float2* __restrict__ out /*[3][batch][nk]*/) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nk) return;
  const float k0 = kern[q] * scale, k1 = kern[nk + q] * scale, k2 = kern[2 * nk + q] * scale;
  for (int b = 0; b < nbatch; b++) {
    const float2 c = crho[(long long)b * nk + q];
    out[((long long)0 * nbatch + b) * nk + q] = make_float2(-c.y * k0, c.x * k0);
    out[((long long)1 * nbatch + b) * nk + q] = make_float2(-c.y * k1, c.x * k1);
    out[((long long)2 * nbatch + b) * nk + q] = make_float2(-c.y * k2, c.x * k2);
  }
}

// force_f(3,nft+2,nft+2,nft+2) [z][y][x][3] <-> F[z'][y'][d][x'] of batch slot 0 (diagnostics only)
__global__ void k_force_to_ref(int M, int FP, const float* __restrict__ F, float* __restrict__ out) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)M * M * M) return;
  const int x = (int)(q % M); const long long r = q / M;
  for (int d = 0; d < 3; d++) out[3 * q + d] = F[(r * 3 + d) * FP + x];
}
__global__ void k_force_from_ref(int M, int FP, const float* __restrict__ in, float* __restrict__ F) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)M * M * M) return;
  const int x = (int)(q % M); const long long r = q / M;
  for (int d = 0; d < 3; d++) F[(r * 3 + d) * FP + x] = in[3 * q + d];
}

// =============================================================================================
// coarse mesh (pm.f90:127-228)
// =============================================================================================
// tempx=((/i,j,k/)-1)+(...)*x_resolution-0.5 -> f32 (pm.f90:142); cell0 = Fortran index - 1.
// = (2K+1)/2^(XB+1) with K = 2^XB*cell0 + u - 2^(XB-1) (every f64 step of the reference expression is exact)
template <int XB> __device__ __forceinline__ float coarse_tempx(int cell0, short xp) {
  return __int2float_rn(2 * ((1 << XB) * cell0 + (int)upat<XB>(xp) - (1 << (XB - 1))) + 1) * (1.0f / (float)(1 << (XB + 1)));
}

// Coarse CIC deposit (pm.f90:130-163) in two passes over the image, no atomics, deterministic:
//   k_coarse_cell_sums   every source cell of the image and of its one-cell rim (extended-grid cells -1..nc per dimension; a rim
//                        cell aliases the periodic image or holds received ghosts) adds its particles, in storage order, into its
//                        27 partial sums S[T][cell], T = target cell-1..cell+1 per dimension -- once per source cell;
//   k_coarse_gather27    every target cell is the sum of the 27 partial sums aimed at it, in the reference's k, j, i source order.
// (The one-kernel form this replaces kept the partial sums in shared memory per 8x8x4 brick and had to redo the source layer
// around every brick: 2.3 visits per particle.)  The weights only depend on the position inside the cell -- every f32 step of
// pm.f90:142-146 is exact for tile-local indices -- so the source cells need no tile frame.
// Crowded source cells (more than `heavy` particles; a warp would wait for its fullest lane, and haloes put 10^3-10^4 particles
// into a coarse cell at late times) are summed by a whole warp instead of their own thread: lane = particle, the 27 products
// wx_a wy_b wz_c mass_p (a particle has two non-zero weights per dimension) are taken in f32 like the reference's, converted to
// fixed point (2^-23) and added over the warp with the integer warp reduction (REDUX); lane T keeps the 64-bit total of
// accumulator T.  Integer sums: independent of any order, deterministic.
constexpr int CD_T = 256;
template <class XT>
__device__ __forceinline__ void coarse_sum_warp(float* __restrict__ my /*[27] stride CD_T*/, const XT* __restrict__ xp, long long s, int n,
                                                float mass_p, int lane) {
  constexpr int XB = 8 * (int)sizeof(XT);
  unsigned long long tot = 0;
  const float scale = __fmul_rn(mass_p, 8388608.0f);
  for (int base = 0; base < n; base += 32) {
    float wx[3] = {0.f, 0.f, 0.f}, wy[3] = {0.f, 0.f, 0.f}, wz[3] = {0.f, 0.f, 0.f};
    if (base + lane < n) {
      const Code3 c = load_code3(xp, s + base + lane);
      int i1, j1, k1; float d1, d2;
      cic_split(coarse_tempx<XB>(1, c.x), i1, d1, d2);  // i1 = 1 or 2: lower target = cell-1 or cell
      if (i1 == 1) { wx[0] = d1; wx[1] = d2; } else { wx[1] = d1; wx[2] = d2; }
      cic_split(coarse_tempx<XB>(1, c.y), j1, d1, d2);
      if (j1 == 1) { wy[0] = d1; wy[1] = d2; } else { wy[1] = d1; wy[2] = d2; }
      cic_split(coarse_tempx<XB>(1, c.z), k1, d1, d2);
      if (k1 == 1) { wz[0] = d1; wz[1] = d2; } else { wz[1] = d1; wz[2] = d2; }
    }
#pragma unroll
    for (int T = 0; T < 27; T++) {
      const unsigned wf = __float2uint_rn(__fmul_rn(__fmul_rn(__fmul_rn(wx[T % 3], wy[(T / 3) % 3]), wz[T / 9]), scale));
      const unsigned sum = __reduce_add_sync(0xffffffffu, wf);  // 32 terms below 2^26 each (mass_p <= 8): no overflow
      if (lane == T) tot += sum;
    }
  }
  if (lane < 27) my[lane * CD_T] = __fmul_rn(__ull2float_rn(tot), 0x1p-23f);
}

// S[T][nbox], box cell b = ((z+1)*(nc+2) + (y+1))*(nc+2) + (x+1), x,y,z = -1..nc
template <class XT>
__global__ void __launch_bounds__(CD_T) k_coarse_cell_sums(Geom g, int heavy, const XT* __restrict__ xp, const int* __restrict__ rhoc_e,
                                                           const long long* __restrict__ cstart_e, float mass_p, long long nbox,
                                                           float* __restrict__ S) {
  constexpr int XB = 8 * (int)sizeof(XT);
  __shared__ float acc[27 * CD_T];  // [T][thread]: a thread's 27 sums sit in one bank
  __shared__ unsigned s_hm[CD_T / 32];
  const int t = threadIdx.x, lane = t & 31, wp = t >> 5;
  const long long b = (long long)blockIdx.x * CD_T + t;
  const int m = g.nc + 2;
  int n = 0;
  long long s = 0;
  if (b < nbox) {
    const long long e = ext_index(g, (int)(b % m) - 1, (int)((b / m) % m) - 1, (int)(b / ((long long)m * m)) - 1);
    n = rhoc_e[e];
    s = cstart_e[e];
  }
#pragma unroll
  for (int T = 0; T < 27; T++) acc[T * CD_T + t] = 0.f;
  {
    const unsigned hm = __ballot_sync(0xffffffffu, n > heavy);
    if (lane == 0) s_hm[wp] = hm;
  }
  __syncthreads();
  if (n <= heavy) {
    float* my = acc + t;
    for (int l = 0; l < n; l++) {
      const Code3 c = load_code3(xp, s + l);
      int i1, j1, k1; float ax[2], ay[2], az[2];
      cic_split(coarse_tempx<XB>(1, c.x), i1, ax[0], ax[1]);  // lower target: i1 - 1 = 0 (cell-1) or 1 (cell)
      cic_split(coarse_tempx<XB>(1, c.y), j1, ay[0], ay[1]);
      cic_split(coarse_tempx<XB>(1, c.z), k1, az[0], az[1]);
      const int ra = i1 - 1, rb = j1 - 1, rc = k1 - 1;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const int qa = q & 1, qb = (q >> 1) & 1, qc = q >> 2;
        const float wgt = __fmul_rn(__fmul_rn(__fmul_rn(ax[qa], ay[qb]), az[qc]), mass_p);  // pm.f90:147-154
        float* p = my + (((rc + qc) * 3 + (rb + qb)) * 3 + (ra + qa)) * CD_T;
        *p = __fadd_rn(*p, wgt);
      }
    }
  }
  {  // crowded cells, dealt round-robin to the warps
    int ord = 0;
    for (int wd = 0; wd < CD_T / 32; wd++) {
      unsigned bits = s_hm[wd];
      while (bits) {
        const int th = wd * 32 + __ffs(bits) - 1;
        bits &= bits - 1;
        if ((ord++ & (CD_T / 32 - 1)) != wp) continue;
        const long long bh = (long long)blockIdx.x * CD_T + th;
        const long long e = ext_index(g, (int)(bh % m) - 1, (int)((bh / m) % m) - 1, (int)(bh / ((long long)m * m)) - 1);
        coarse_sum_warp(acc + th, xp, cstart_e[e], rhoc_e[e], mass_p, lane);
      }
    }
  }
  __syncthreads();
  if (b < nbox) {
#pragma unroll
    for (int T = 0; T < 27; T++) S[T * nbox + b] = acc[T * CD_T + t];
  }
}

// r3(X,Y,Z) = sum over the 27 source cells (X+dx, Y+dy, Z+dz), in k, j, i order, of the partial sum each aims at this target
__global__ void __launch_bounds__(256) k_coarse_gather27(Geom g, long long nbox, const float* __restrict__ S, float* __restrict__ r3 /*[nc][nc][ld]*/, int ld,
                                                         int accumulate /* add to r3: a further species (pm.f90:160, NEUTRINOS) */) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= g.ncell_p) return;
  const int X = (int)(q % g.nc), Y = (int)((q / g.nc) % g.nc), Z = (int)(q / ((long long)g.nc * g.nc));
  const int m = g.nc + 2;
  float v = 0.f;
#pragma unroll
  for (int dz = -1; dz <= 1; dz++)
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        // source S = target + d holds this target at relative index 1 - d
        const long long b = ((long long)(Z + dz + 1) * m + (Y + dy + 1)) * m + (X + dx + 1);
        v = __fadd_rn(v, S[(((1 - dz) * 3 + (1 - dy)) * 3 + (1 - dx)) * nbox + b]);
      }
  float* dst = r3 + ((long long)Z * g.nc + Y) * ld + X;
  *dst = accumulate ? __fadd_rn(*dst, v) : v;
}

// force_c(3,0:nc+1,0:nc+1,0:nc+1) from the three inverse transforms + periodic 1-cell halo
// (pm.f90:176-189, single image: neighbour = self)
__global__ void k_force_c_assemble(Geom g, const float* __restrict__ F /*[3][nc][nc][nc+2]*/, float* __restrict__ fc) {
  const int m = g.nc + 2;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)m * m * m) return;
  int x = (int)(q % m) - 1, y = (int)((q / m) % m) - 1, z = (int)(q / ((long long)m * m)) - 1;
  x = (x + g.nc) % g.nc; y = (y + g.nc) % g.nc; z = (z + g.nc) % g.nc;
  const long long vol = (long long)g.nc * g.nc * (g.nc + 2), o = ((long long)z * g.nc + y) * (g.nc + 2) + x;
  fc[3 * q] = F[o]; fc[3 * q + 1] = F[vol + o]; fc[3 * q + 2] = F[2 * vol + o];
}

__global__ void __launch_bounds__(256) k_f2max_aos(long long n, const float* __restrict__ f, unsigned* __restrict__ f2max) {
  float best = 0.f;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    float f0 = f[3 * q], f1 = f[3 * q + 1], f2 = f[3 * q + 2];
    best = fmaxf(best, __fadd_rn(__fadd_rn(__fmul_rn(f0, f0), __fmul_rn(f1, f1)), __fmul_rn(f2, f2)));
  }
  best = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(best)));
  if ((threadIdx.x & 31) == 0) atomicMax(f2max, __float_as_uint(best));
}

// =============================================================================================
// matter power spectrum of the resident state (CUBE/utilities/cicpower.f90:70-140, powerspectrum.f90:47-108, linear_kbin)
// =============================================================================================
// density on the (n+4)^3 deposit grid (element e = true node e-1; k_fine_deposit_r<.., HALF>) -> rho(n,n,n) in the in-place r2c
// layout [n][n][n+2]; one partial sum per block, fixed order
__global__ void __launch_bounds__(256) k_ps_extract(int n, const float* __restrict__ dep, float* __restrict__ rho, double* __restrict__ part) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long tot = (long long)n * n * n;
  const int m = n + 4;
  double v = 0;
  if (q < tot) {
    const int x = (int)(q % n), y = (int)((q / n) % n), z = (int)(q / ((long long)n * n));
    const float f = dep[((long long)(z + 1) * m + (y + 1)) * m + (x + 1)];
    rho[((long long)z * n + y) * (n + 2) + x] = f;
    v = (double)f;
  }
  __shared__ double sm[8];
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; w++) t += sm[w]; part[blockIdx.x] = t; }
}
// rho1=rho1/(rho8/ng_global^3)-1 (cicpower.f90:134)
__global__ void __launch_bounds__(256) k_ps_contrast(int n, float* __restrict__ rho, const double* __restrict__ sum) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n * n * n) return;
  const int x = (int)(q % n);
  const long long r = q / n;
  const float mean = (float)(*sum / (double)n / (double)n / (double)n);
  float* p = rho + r * (n + 2) + x;
  *p = *p / mean - 1.0f;
}
// shell sums of powerspectrum.f90:47-86 (linear_kbin: ibin = nint(kr)) for the auto power of c = FFT(delta): per bin count, sum kr,
// sum |c|^2 4 pi kr^3 / n^6 / sinc^4, sum 1/sinc^2, sum 1/sinc^4.  Accumulated in f64 (shared-memory bins per CTA, then global).
constexpr int PS_Q = 5;
__global__ void __launch_bounds__(256) k_ps_bin(int n, int nbin, const float2* __restrict__ c, double* __restrict__ bins /*[PS_Q][nbin]*/) {
  extern __shared__ double sb[];  // [PS_Q][nbin]
  for (int t = threadIdx.x; t < PS_Q * nbin; t += blockDim.x) sb[t] = 0.0;
  __syncthreads();
  const int nyq = n / 2, nh = nyq + 1;
  const long long tot = (long long)n * n * nh;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < tot; q += (long long)gridDim.x * blockDim.x) {
    const int ig = (int)(q % nh), jg = (int)((q / nh) % n), kg = (int)(q / ((long long)nh * n));  // 0-based
    if (ig == 0 && jg == 0 && kg == 0) continue;                                   // zero frequency
    const bool edge = ig == 0 || ig == nyq;
    if (edge && jg > nyq) continue;                                                // :57
    if (edge && (jg == 0 || jg == nyq) && kg > nyq) continue;                      // :58
    const float kx = (float)ig, ky = (float)((jg + nyq) % n - nyq), kz = (float)((kg + nyq) % n - nyq);
    const float kr = sqrtf(kx * kx + ky * ky + kz * kz);
    const float ax = PI_F * kx / (float)n, ay = PI_F * ky / (float)n, az = PI_F * kz / (float)n;
    const float sinc = (kx == 0.f ? 1.f : sinf(ax) / ax) * (ky == 0.f ? 1.f : sinf(ay) / ay) * (kz == 0.f ? 1.f : sinf(az) / az);
    const int ibin = (int)lrintf(kr);  // kr^2 is an integer: never a tie
    if (ibin < 1 || ibin > nbin) continue;
    const float2 v = c[q];
    const double n3 = (double)n * n * n, s2 = (double)sinc * sinc, s4 = s2 * s2;
    const double amp = ((double)v.x * v.x + (double)v.y * v.y) / n3 / n3 / s4 * 4.0 * (double)PI_F * (double)kr * kr * kr;
    double* b = sb + (ibin - 1);
    atomicAdd(b, 1.0); atomicAdd(b + nbin, (double)kr); atomicAdd(b + 2 * nbin, amp); atomicAdd(b + 3 * nbin, 1.0 / s2); atomicAdd(b + 4 * nbin, 1.0 / s4);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < PS_Q * nbin; t += blockDim.x) if (sb[t] != 0.0) atomicAdd(bins + t, sb[t]);
}

// =============================================================================================
// kernel construction (kernel_f.f90, kernel_c.f90) -- init only
// =============================================================================================
// rho_f <- 16^3 table mirrored into 8 octants, odd along the force's own axis (kernel_f.f90:32-38)
__global__ void k_kernf_fill(int nfe, const float* __restrict__ tab /*(16,16,16,3) Fortran*/, int dim, float* __restrict__ rho) {
  const long long ld = nfe + 2, n = (long long)nfe * nfe * ld;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int x = (int)(q % ld), y = (int)((q / ld) % nfe), z = (int)(q / (ld * nfe));
  float v = 0.f;
  if (x < nfe) {
    int ox = x < 16 ? x : x - nfe, oy = y < 16 ? y : y - nfe, oz = z < 16 ? z : z - nfe;
    if (ox > -16 && oy > -16 && oz > -16 && ox < 16 && oy < 16 && oz < 16) {
      int o[3] = {ox, oy, oz};
      v = tab[(((long long)dim * 16 + abs(oz)) * 16 + abs(oy)) * 16 + abs(ox)];
      if (o[dim] < 0) v = -v;
    }
  }
  rho[q] = v;
}
// Im of an r2c result c[z][y][n/2+1] into a kx-pitched real array out[z][y][P]
__global__ void k_take_imag_pitched(int n, int P, const float2* __restrict__ c, float* __restrict__ out) {
  const int nh = n / 2 + 1;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)n * n * nh) return;
  out[(q / nh) * P + (q % nh)] = c[q].y;
}
__global__ void k_take_imag(long long nk, const float2* __restrict__ c, float* __restrict__ out) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nk) out[q] = c[q].y;
}

__device__ __forceinline__ int signed_index(int g, int n) { return (g + n / 2) % n - n / 2; }

// ck(dim) on the (single-image) coarse lattice: -r/r^3 with the 4^3 table in the 8 corners when
// `corrected` (kernel_c.f90:16-72)
__global__ void k_kernc_fill(int nc, const float* __restrict__ tab /*(3,4,4,4) Fortran*/, int dim, int corrected, float* __restrict__ r3) {
  const long long ld = nc + 2, n = (long long)nc * nc * ld;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int x = (int)(q % ld), y = (int)((q / ld) % nc), z = (int)(q / (ld * nc));
  float v = 0.f;
  if (x < nc) {
    int o[3] = {signed_index(x, nc), signed_index(y, nc), signed_index(z, nc)};
    float rx = 4.f * o[0], ry = 4.f * o[1], rz = 4.f * o[2];
    float r = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    float rr[3] = {rx, ry, rz};
    v = (r == 0.f) ? 0.f : __fdiv_rn(-rr[dim], __fmul_rn(__fmul_rn(r, r), r));
    if (corrected && o[0] > -4 && o[0] < 4 && o[1] > -4 && o[1] < 4 && o[2] > -4 && o[2] < 4) {
      // positive side uses table index 0..3, negative side (offsets -3..-1) mirrors 3..1
      v = tab[dim + 3 * (abs(o[0]) + 4 * (abs(o[1]) + 4 * abs(o[2])))];
      if (o[dim] < 0) v = -v;
    }
  }
  r3[q] = v;
}
// LRCKCORR (kernel_c.f90:76-117) on the half-x k-space grid of a single image
__global__ void k_kernc_lrck(int nc, int dim, const float2* __restrict__ cpure, float* __restrict__ kern) {
  const int nh = nc / 2 + 1;
  long long nk = (long long)nh * nc * nc;
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nk) return;
  int x = (int)(q % nh), y = (int)((q / nh) % nc), z = (int)(q / ((long long)nh * nc));
  float kx[3] = {(float)signed_index(x, nc), (float)signed_index(y, nc), (float)signed_index(z, nc)};
  float kr = sqrtf(kx[0] * kx[0] + kx[1] * kx[1] + kx[2] * kx[2]);
  if (kr > 8.0f || kx[dim] == 0.f) return;
  float ks[3];
  for (int d = 0; d < 3; d++) ks[d] = 2.f * sinf(PI_F * kx[d] / (float)nc);
  float ssum = ks[0] * ks[0] + ks[1] * ks[1] + ks[2] * ks[2];
  kern[q] = kern[q] * 0.25f * PI_F * ks[dim] / ssum / cpure[q].y;
}

}  // namespace cube
