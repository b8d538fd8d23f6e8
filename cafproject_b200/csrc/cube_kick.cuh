// cube_kick.cuh -- fine kick (pm.f90:88-118) and coarse kick (pm.f90:196-228) of a batch of tiles in ONE pass over the particles.
//
// Per particle the two kicks are the reference's operations in the reference's order: decode with sigma_vi, add the eight fine
// CIC terms, re-quantise with sigma_vi_new (pm.f90:113), decode that code again (pm.f90:209), add the eight coarse terms,
// vmax, re-quantise (pm.f90:221).  What the merge saves is everything around them: one read of xp, one read and one write of
// vp instead of two each, one cell search, one set of index arithmetic.
//
// A persistent CTA (one per SM) holds the host tanf table of the velocity decode in shared memory (|code| < hot, cube_particles.cuh
// v_decode) and runs KB_G independent groups of KB_GT threads.  A group takes bricks of 8 x 2 x 2 coarse cells (about 256
// particles, one per thread; a group is wider than the mean so that a brick rarely needs a second, nearly empty, round) from a
// global counter, and for each brick
//   * brings the force nodes its particles can touch -- 36 x 9 x 9 nodes x 3 components of F[b][z'][y'][d][x'] -- into shared
//     memory by BULK ASYNCHRONOUS COPIES (cp.async.bulk, one 144-byte row per copy, completion counted on an mbarrier): rows start
//     at x' = 4*cx0, a 16-byte boundary, one node before the first one needed.  The 24 force gathers per particle then hit
//     shared memory (~3 wavefronts per warp-wide gather of scattered rows instead of ~16 through L1: the kernel this replaces
//     ran at 88 % of the L1 data pipe, profiles/r01k_ncu_full_cfg1.csv);
//   * prefetches the rows of the brick it will take next into L2 (cp.async.bulk.prefetch.L2);
//   * stages the brick's coarse force nodes (10 x 4 x 4 x 3 floats) with plain coalesced loads.
// Groups synchronise among themselves with named barriers only: while one waits for its copies the others compute.
// Built for 2-byte codes (izipx = izipv = 2).  OPT-IN (CUBE_GPU_MERGED_KICK): measured on B200 at cfg 2 it takes 7.85 ms against
// 3.95 + 2.96 ms for the two separate kicks (k_fine_kick_p, k_coarse_kick_w) -- with 64 registers x 960 threads and one force
// brick per group the SM holds 30 warps, and the per-brick chain (brick id, cell table, box copy, codes) is exposed; see
// profiles/r02_notes.md for the variants (row-wise bulk copies 11.7 ms, cp.async 12.8 ms, one CTA per brick 4.9 ms fine-only).
#pragma once
#include <cuda.h>  // CUtensorMap (the encode function is fetched at run time: no link against libcuda)
#include "cube_fft.cuh"
#include "cube_particles.cuh"

namespace cube {

constexpr int KB_X = 8, KB_Y = 2, KB_Z = 2, KB_CELLS = KB_X * KB_Y * KB_Z;  // 32 cells: one warp scans their counts
constexpr int KB_G = 3, KB_GT = 320, KB_T = KB_G * KB_GT;
constexpr int KB_NX = 4 * KB_X + 4, KB_NY = 4 * KB_Y + 1, KB_NZ = 4 * KB_Z + 1;  // staged nodes per dimension (x: 16-byte rows)
constexpr int KB_ROWS = KB_NZ * KB_NY * 3, KB_ROWB = KB_NX * 4;                 // 243 rows of 144 bytes
constexpr int KB_FW = KB_ROWS * KB_NX;                                          // floats of the fine brick
constexpr int KB_CX = KB_X + 2, KB_CY = KB_Y + 2, KB_CZ = KB_Z + 2, KB_CW = KB_CX * KB_CY * KB_CZ * 3;  // coarse brick
static_assert(KB_CELLS == 32 && KB_ROWS <= KB_GT, "one warp scans the cells; one copy per thread");
// per group: fine brick, coarse brick, prefix offsets (+pad), first particles, vfield, mbarrier, next brick
constexpr int KB_GROUP_BYTES = (KB_FW * 4 + KB_CW * 4 + (KB_CELLS + 4) * 4 + KB_CELLS * 8 + KB_CELLS * 12 + 16 + 127) / 128 * 128;
static_assert(KB_FW * 4 % 16 == 0, "group buffers stay aligned (a tensor-map copy wants 128 bytes)");
constexpr int KB_HOT_DEFAULT = 24576;
__host__ __device__ inline size_t kb_smem_bytes(int hot) { return (size_t)hot * 4 + (size_t)KB_G * KB_GROUP_BYTES; }

struct KickArgs {
  Geom g;
  int tile0, nb;            // tiles [tile0, tile0+nb): F holds their force_f (x a_mid dt/6/pi) in this order
  int M, FP;
  const float* F;           // nullptr: no fine kick
  const float* Gc;          // (3,0:nc+1,0:nc+1,0:nc+1) x a_mid dt/6/pi; nullptr: no coarse kick
  const short* xp; short* vp;
  const long long* cstart_p; const float* vfield_p;
  VTab vt_in; double S_in;    // codes as they come in (sigma_vi)
  VTab vt_out; double S_out;  // codes after the fine kick and after the coarse kick (sigma_vi_new)
  unsigned long long* vmax_bits;
  unsigned long long* next_brick;  // work counter, zeroed by the caller
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 16-byte aligned global -> shared bulk copy; its bytes are counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}
// one box of the 5-d tensor F[b][z'][y'][d][x'] -> shared memory (TMA); out-of-range parts of the box arrive as zeros
__device__ __forceinline__ void tma_box5(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, void* bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_prefetch5(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// brick -> batch slot and first cell
struct BrickPos { int b, cx0, cy0, cz0; };
__device__ __forceinline__ BrickPos brick_pos(const Geom& g, long long brick) {
  const int nbx = (g.nt + KB_X - 1) / KB_X, nby = (g.nt + KB_Y - 1) / KB_Y, nbz = (g.nt + KB_Z - 1) / KB_Z;
  const int bpt = nbx * nby * nbz;
  BrickPos p;
  p.b = (int)(brick / bpt);
  const int r = (int)(brick - (long long)p.b * bpt);
  p.cx0 = (r % nbx) * KB_X; p.cy0 = ((r / nbx) % nby) * KB_Y; p.cz0 = (r / (nbx * nby)) * KB_Z;
  return p;
}
// global address of staged row r = (z*KB_NY + y)*3 + d of the brick (rows beyond the kept points repeat the last one: never read)
__device__ __forceinline__ const float* brick_row(const KickArgs& A, const BrickPos& bp, int r) {
  const int d = r % 3, y = (r / 3) % KB_NY, z = r / (3 * KB_NY);
  const int yy = min(4 * bp.cy0 + 1 + y, A.M - 1), zz = min(4 * bp.cz0 + 1 + z, A.M - 1);
  return A.F + ((((size_t)bp.b * A.M + zz) * A.M + yy) * 3 + d) * (size_t)A.FP + 4 * bp.cx0;
}
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;\n" ::"r"(grp + 1), "n"(KB_GT) : "memory"); }

// STAGE: how a brick's force nodes reach shared memory -- 2: one TMA box copy per brick (tensor map), 0: one bulk copy per row,
// 1: 16-byte cp.async copies issued by all threads of the group
template <int STAGE>
__global__ void __launch_bounds__(KB_T, 1) k_kick_brick(KickArgs A, const __grid_constant__ CUtensorMap fmap) {
  extern __shared__ __align__(128) unsigned char kb_smem[];
  float* s_tan = reinterpret_cast<float*>(kb_smem);  // [hot]
  const int grp = threadIdx.x / KB_GT, t = threadIdx.x - grp * KB_GT, lane = t & 31;
  unsigned char* gb = kb_smem + (size_t)A.vt_in.hot * 4 + (size_t)grp * KB_GROUP_BYTES;
  float* sF = reinterpret_cast<float*>(gb);                                        // [KB_ROWS][KB_NX]
  float* sC = sF + KB_FW;                                                          // [KB_CZ][KB_CY][KB_CX][3]
  int* soff = reinterpret_cast<int*>(sC + KB_CW);                                  // [KB_CELLS+1] (+3 pad)
  long long* sstart = reinterpret_cast<long long*>(soff + KB_CELLS + 4);           // [KB_CELLS] first particle of the cell
  float* svf = reinterpret_cast<float*>(sstart + KB_CELLS);                        // [KB_CELLS][3] vfield of the cell (vmax)
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(svf + 3 * KB_CELLS);
  long long* s_next = reinterpret_cast<long long*>(bar + 1);                       // the brick this group takes next
  const Geom& g = A.g;
  const int nt = g.nt;
  const long long nt3 = (long long)nt * nt * nt;
  const long long nbrick = (long long)A.nb * ((nt + KB_X - 1) / KB_X) * ((nt + KB_Y - 1) / KB_Y) * ((nt + KB_Z - 1) / KB_Z);
  if (A.vt_in.hot) fill_tab(s_tan, A.vt_in.tanh, A.vt_in.hot);
  const VDec dec_in = make_dec(A.vt_in, s_tan, A.S_in), dec_out = make_dec(A.vt_out, s_tan, A.S_out);
  if (t == 0) { mbar_init(bar, 1); *s_next = (long long)atomicAdd(A.next_brick, 1ULL); }
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  const int mc = g.nc + 2;
  double vm = 0.0;
  unsigned parity = 0;
  long long brick = *s_next;
  while (brick < nbrick) {
    const BrickPos bp = brick_pos(g, brick);
    const int tile = A.tile0 + bp.b;
    const int tx = tile % g.nnt, ty = (tile / g.nnt) % g.nnt, tz = tile / (g.nnt * g.nnt);
    group_sync(grp);  // everybody has read s_next and is done with the previous brick's buffers
    // --- fine force brick: bulk copies, all in flight at once
    if (t == 0) {
      if (A.F && STAGE != 1) mbar_expect_tx(bar, KB_ROWS * KB_ROWB);
      if (A.F && STAGE == 2) tma_box5(sF, &fmap, 4 * bp.cx0, 0, 4 * bp.cy0 + 1, 4 * bp.cz0 + 1, bp.b, bar);
      *s_next = (long long)atomicAdd(A.next_brick, 1ULL);
    }
    if (A.F && STAGE == 0 && t < KB_ROWS) bulk_g2s(sF + t * KB_NX, brick_row(A, bp, t), KB_ROWB, bar);
    if (A.F && STAGE == 1) {
      for (int e = t; e < KB_ROWS * (KB_NX / 4); e += KB_GT) {
        const int r = e / (KB_NX / 4), ch = e - r * (KB_NX / 4);
        cp_async16(sF + r * KB_NX + 4 * ch, brick_row(A, bp, r) + 4 * ch);
      }
      cp_async_commit();
    }
    // --- the brick's cells: first particle, count, prefix (one warp)
    if (t < KB_CELLS) {
      const int x = t % KB_X, y = (t / KB_X) % KB_Y, z = t / (KB_X * KB_Y);
      const int i = bp.cx0 + x, j = bp.cy0 + y, k = bp.cz0 + z;
      long long s = 0; int n = 0;
      if (i < nt && j < nt && k < nt) {
        const long long L = (long long)tile * nt3 + ((long long)k * nt + j) * nt + i;
        s = A.cstart_p[L];
        n = (int)(A.cstart_p[L + 1] - s);
        svf[3 * t] = A.vfield_p[3 * L]; svf[3 * t + 1] = A.vfield_p[3 * L + 1]; svf[3 * t + 2] = A.vfield_p[3 * L + 2];
      }
      sstart[t] = s;
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
      soff[t] = incl - n;
      if (t == KB_CELLS - 1) soff[KB_CELLS] = incl;
    }
    // --- coarse force brick (image-local nodes X0 .. X0+KB_X+1 of the (0:nc+1) array)
    if (A.Gc) {
      const int X0 = tx * nt + bp.cx0, Y0 = ty * nt + bp.cy0, Z0 = tz * nt + bp.cz0;
      for (int e = t; e < KB_CW; e += KB_GT) {
        const int c = e % (3 * KB_CX), y = (e / (3 * KB_CX)) % KB_CY, z = e / (3 * KB_CX * KB_CY);
        const int Y = min(Y0 + y, mc - 1), Z = min(Z0 + z, mc - 1), XC = min(3 * X0 + c, 3 * mc - 1);
        sC[e] = __ldg(A.Gc + ((size_t)Z * mc + Y) * (3 * mc) + XC);
      }
    }
    if (STAGE == 1) cp_async_wait<0>();
    group_sync(grp);
    // the group's next brick goes to L2 while this one is worked on
    const long long nxt = *s_next;
    if (A.F && nxt < nbrick) {
      if (STAGE == 2) {
        if (t == 0) { const BrickPos bn = brick_pos(g, nxt); tma_prefetch5(&fmap, 4 * bn.cx0, 0, 4 * bn.cy0 + 1, 4 * bn.cz0 + 1, bn.b); }
      } else if (t < KB_ROWS) bulk_prefetch_l2(brick_row(A, brick_pos(g, nxt), t), KB_ROWB);
    }
    const int np = soff[KB_CELLS];
    if (A.F && STAGE != 1) mbar_wait(bar, parity);
    parity ^= 1u;
    for (int q = t; q < np; q += KB_GT) {
      int c = 0;
#pragma unroll
      for (int step = KB_CELLS / 2; step > 0; step >>= 1)
        if (soff[c + step] <= q) c += step;
      const long long p = sstart[c] + (q - soff[c]);
      const int x = c % KB_X, y = (c / KB_X) % KB_Y, z = c / (KB_X * KB_Y);
      const Code3 xc = load_code3(A.xp, p), vc = load_code3(A.vp, p);
      const int qx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, qy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, qz[8] = {0, 0, 0, 1, 1, 1, 0, 1};  // pm.f90:104-111
      double v0, v1, v2;
      short w0 = vc.x, w1 = vc.y, w2 = vc.z;
      if (A.F) {
        v0 = v_decode<16>(dec_in, vc.x); v1 = v_decode<16>(dec_in, vc.y); v2 = v_decode<16>(dec_in, vc.z);
        int i1, j1, k1; float ax[2], ay[2], az[2];
        cic_split(fine_tempx<16>(bp.cx0 + x + 1, xc.x), i1, ax[0], ax[1]);  // idx1 = 0-based kept index
        cic_split(fine_tempx<16>(bp.cy0 + y + 1, xc.y), j1, ay[0], ay[1]);
        cic_split(fine_tempx<16>(bp.cz0 + z + 1, xc.z), k1, az[0], az[1]);
        // brick-local node: x from 4*cx0 (the aligned row start), y and z from the first node needed, 4*c0+1
        const float* f = sF + (((k1 - 4 * bp.cz0 - 1) * KB_NY + (j1 - 4 * bp.cy0 - 1)) * 3) * KB_NX + (i1 - 4 * bp.cx0);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float* fu = f + ((qz[u] * KB_NY + qy[u]) * 3) * KB_NX + qx[u];
          const float wx = ax[qx[u]], wy = ay[qy[u]], wz = az[qz[u]];
          v0 = __dadd_rn(v0, (double)kick_weight(fu[0], wx, wy, wz));
          v1 = __dadd_rn(v1, (double)kick_weight(fu[KB_NX], wx, wy, wz));
          v2 = __dadd_rn(v2, (double)kick_weight(fu[2 * KB_NX], wx, wy, wz));
        }
        w0 = vp_encode_lut<16>(v0, A.S_out, A.vt_out.thr); w1 = vp_encode_lut<16>(v1, A.S_out, A.vt_out.thr); w2 = vp_encode_lut<16>(v2, A.S_out, A.vt_out.thr);
      }
      if (A.Gc) {
        v0 = v_decode<16>(dec_out, w0); v1 = v_decode<16>(dec_out, w1); v2 = v_decode<16>(dec_out, w2);
        const int X = tx * nt + bp.cx0 + x, Y = ty * nt + bp.cy0 + y, Z = tz * nt + bp.cz0 + z;  // ((itx-1)*nt + (i-1)) of pm.f90:206
        int i1, j1, k1; float ax[2], ay[2], az[2];
        cic_split(coarse_tempx<16>(X, xc.x), i1, ax[0], ax[1]);
        cic_split(coarse_tempx<16>(Y, xc.y), j1, ay[0], ay[1]);
        cic_split(coarse_tempx<16>(Z, xc.z), k1, az[0], az[1]);
        const float* f = sC + (((k1 - (Z - z)) * KB_CY + (j1 - (Y - y))) * KB_CX + (i1 - (X - x))) * 3;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float* fu = f + ((qz[u] * KB_CY + qy[u]) * KB_CX + qx[u]) * 3;
          const float wx = ax[qx[u]], wy = ay[qy[u]], wz = az[qz[u]];
          v0 = __dadd_rn(v0, (double)kick_weight(fu[0], wx, wy, wz));
          v1 = __dadd_rn(v1, (double)kick_weight(fu[1], wx, wy, wz));
          v2 = __dadd_rn(v2, (double)kick_weight(fu[2], wx, wy, wz));
        }
        const double vf0 = svf[3 * c], vf1 = svf[3 * c + 1], vf2 = svf[3 * c + 2];
        vm = fmax(vm, fmax(__dadd_rn(v0, vf0), fmax(__dadd_rn(v1, vf1), __dadd_rn(v2, vf2))));  // pm.f90:220
        w0 = vp_encode_lut<16>(v0, A.S_out, A.vt_out.thr); w1 = vp_encode_lut<16>(v1, A.S_out, A.vt_out.thr); w2 = vp_encode_lut<16>(v2, A.S_out, A.vt_out.thr);
      }
      store_code3(A.vp, p, w0, w1, w2);
    }
    brick = nxt;
  }
  if (A.Gc && A.vmax_bits) {
    for (int o = 16; o; o >>= 1) vm = fmax(vm, __shfl_down_sync(FULL, vm, o));
    if (lane == 0 && vm > 0.0) atomicMax(A.vmax_bits, (unsigned long long)__double_as_longlong(vm));
  }
}

// ---------------------------------------------------------------------------------------------
// Fine kick alone, one CTA per brick (not persistent): a 256-thread CTA fetches its brick of force nodes with one TMA box copy
// (35 KB of shared memory: six CTAs per SM hide each other's copy and setup latencies) and decodes through the global f64 table
// like the kernel it replaces (k_fine_kick_p) -- what changes is where the 24 force gathers per particle go.
// grid = (bricks per tile, nb)
// ---------------------------------------------------------------------------------------------
constexpr int FKB_T = 256;
__global__ void __launch_bounds__(FKB_T) k_fine_kick_brick(Geom g, int tile0, int M, const short* __restrict__ xp, short* __restrict__ vp,
                                                           const long long* __restrict__ cstart_p, const double* __restrict__ dvlut,
                                                           const double* __restrict__ enc, double S_new, const __grid_constant__ CUtensorMap fmap) {
  __shared__ __align__(128) float sF[KB_FW];
  __shared__ int soff[KB_CELLS + 1];
  __shared__ long long sstart[KB_CELLS];
  __shared__ __align__(8) unsigned long long bar;
  const int t = threadIdx.x, lane = t & 31;
  const int nt = g.nt;
  const long long nt3 = (long long)nt * nt * nt;
  const int nbx = (nt + KB_X - 1) / KB_X, nby = (nt + KB_Y - 1) / KB_Y;
  const int r = blockIdx.x, b = blockIdx.y;
  const int cx0 = (r % nbx) * KB_X, cy0 = ((r / nbx) % nby) * KB_Y, cz0 = (r / (nbx * nby)) * KB_Z;
  const int tile = tile0 + b;
  if (t == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  __syncthreads();
  if (t == 0) {
    mbar_expect_tx(&bar, KB_ROWS * KB_ROWB);
    tma_box5(sF, &fmap, 4 * cx0, 0, 4 * cy0 + 1, 4 * cz0 + 1, b, &bar);
  }
  if (t < KB_CELLS) {
    const int x = t % KB_X, y = (t / KB_X) % KB_Y, z = t / (KB_X * KB_Y);
    const int i = cx0 + x, j = cy0 + y, k = cz0 + z;
    long long s = 0; int n = 0;
    if (i < nt && j < nt && k < nt) {
      const long long L = (long long)tile * nt3 + ((long long)k * nt + j) * nt + i;
      s = cstart_p[L];
      n = (int)(cstart_p[L + 1] - s);
    }
    sstart[t] = s;
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
    soff[t] = incl - n;
    if (t == KB_CELLS - 1) soff[KB_CELLS] = incl;
  }
  __syncthreads();
  const int np = soff[KB_CELLS];
  // the particles' codes are fetched while the box is still on its way
  int c = 0; long long p = 0; Code3 xc = {0, 0, 0}, vc = {0, 0, 0};
  if (t < np) {
#pragma unroll
    for (int step = KB_CELLS / 2; step > 0; step >>= 1)
      if (soff[c + step] <= t) c += step;
    p = sstart[c] + (t - soff[c]);
    xc = load_code3(xp, p); vc = load_code3(vp, p);
  }
  mbar_wait(&bar, 0);
  for (int q = t; q < np; q += FKB_T) {
    if (q != t) {
      c = 0;
#pragma unroll
      for (int step = KB_CELLS / 2; step > 0; step >>= 1)
        if (soff[c + step] <= q) c += step;
      p = sstart[c] + (q - soff[c]);
      xc = load_code3(xp, p); vc = load_code3(vp, p);
    }
    const int x = c % KB_X, y = (c / KB_X) % KB_Y, z = c / (KB_X * KB_Y);
    const int qx[8] = {0, 1, 0, 0, 0, 1, 1, 1}, qy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, qz[8] = {0, 0, 0, 1, 1, 1, 0, 1};  // pm.f90:104-111
    double v0 = dvlut[upat<16>(vc.x)], v1 = dvlut[upat<16>(vc.y)], v2 = dvlut[upat<16>(vc.z)];
    int i1, j1, k1; float ax[2], ay[2], az[2];
    cic_split(fine_tempx<16>(cx0 + x + 1, xc.x), i1, ax[0], ax[1]);  // idx1 = 0-based kept index
    cic_split(fine_tempx<16>(cy0 + y + 1, xc.y), j1, ay[0], ay[1]);
    cic_split(fine_tempx<16>(cz0 + z + 1, xc.z), k1, az[0], az[1]);
    const float* f = sF + (((k1 - 4 * cz0 - 1) * KB_NY + (j1 - 4 * cy0 - 1)) * 3) * KB_NX + (i1 - 4 * cx0);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const float* fu = f + ((qz[u] * KB_NY + qy[u]) * 3) * KB_NX + qx[u];
      const float wx = ax[qx[u]], wy = ay[qy[u]], wz = az[qz[u]];
      v0 = __dadd_rn(v0, (double)kick_weight(fu[0], wx, wy, wz));
      v1 = __dadd_rn(v1, (double)kick_weight(fu[KB_NX], wx, wy, wz));
      v2 = __dadd_rn(v2, (double)kick_weight(fu[2 * KB_NX], wx, wy, wz));
    }
    store_code3(vp, p, vp_encode_lut<16>(v0, S_new, enc), vp_encode_lut<16>(v1, S_new, enc), vp_encode_lut<16>(v2, S_new, enc));
  }
}

}  // namespace cube
