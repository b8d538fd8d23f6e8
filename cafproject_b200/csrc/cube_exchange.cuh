// cube_exchange.cuh -- tile-buffer (ghost) exchange between images: buffer_density.f90 / buffer_x.f90 / buffer_v.f90
// for nn > 1, with the coarray GETs replaced by packed per-direction messages (cube_comm.cuh).
//
// The reference syncs ghost layers x, then y, then z so that edges and corners arrive transitively
// (buffer_density.f90:11-68, buffer_x.f90:12-191).  Here every one of the 26 directions r = (rx,ry,rz) is one
// message: the receiver's ghost box in direction r (ncb cells deep where r_d != 0, nc cells where r_d == 0) is the
// sender's physical box on the opposite side.  A direction whose source image is this image itself (nn_d == 1 in
// every dim with r_d != 0) is not a message at all: those ghost cells alias the physical particles (k_build_ext).
// Ghost particles are received straight into the tail of xp / vp behind the nplocal physical ones, in message
// order = direction order, then box cells in z,y,x order, then the sender's in-cell order (which is what the
// reference's row copies preserve).
#pragma once
#include <vector>

#include "cube_common.cuh"
#include "cube_particles.cuh"

namespace cube {

struct ExDir {
  int r[3];
  int src_rank, dst_rank;   // I receive my ghost box(r) from src_rank; I send my physical box(r) to dst_rank
  long long cell0, ncell;   // position of this direction's cells in the ghost / send cell lists
};

struct ExPlan {
  std::vector<ExDir> dirs;   // real (non-alias) directions, increasing direction index
  long long ng = 0;          // ghost cells = send cells
  std::vector<int> gcell_ext;   // [ng] extended-grid index of each ghost cell, message order
  std::vector<int> scell_L;     // [ng] file-order index of each cell I send, message order
};

inline int image_rank(const Geom& g, int cx, int cy, int cz) {
  cx = ((cx % g.nn[0]) + g.nn[0]) % g.nn[0]; cy = ((cy % g.nn[1]) + g.nn[1]) % g.nn[1]; cz = ((cz % g.nn[2]) + g.nn[2]) % g.nn[2];
  return cx + g.nn[0] * (cy + g.nn[1] * cz);
}

inline void build_exchange_plan(const Geom& g, int my_rank, ExPlan& P, bool with_lists) {
  P.dirs.clear(); P.ng = 0; P.gcell_ext.clear(); P.scell_L.clear();
  for (int rz = -1; rz <= 1; rz++)
    for (int ry = -1; ry <= 1; ry++)
      for (int rx = -1; rx <= 1; rx++) {
        if (!rx && !ry && !rz) continue;
        ExDir d;
        d.r[0] = rx; d.r[1] = ry; d.r[2] = rz;
        d.src_rank = image_rank(g, g.ic[0] + rx, g.ic[1] + ry, g.ic[2] + rz);
        d.dst_rank = image_rank(g, g.ic[0] - rx, g.ic[1] - ry, g.ic[2] - rz);
        const bool alias = (rx == 0 || g.nn[0] == 1) && (ry == 0 || g.nn[1] == 1) && (rz == 0 || g.nn[2] == 1);
        if (alias) continue;  // source image == this image: the ghost cells alias the physical particles
        int n[3], glo[3], slo[3];
        for (int a = 0; a < 3; a++) {
          n[a] = d.r[a] ? NCB : g.nc;
          glo[a] = d.r[a] < 0 ? -NCB : (d.r[a] == 0 ? 0 : g.nc);   // my ghost box
          slo[a] = d.r[a] < 0 ? g.nc - NCB : 0;                    // the box I send for this direction
        }
        d.cell0 = P.ng; d.ncell = (long long)n[0] * n[1] * n[2];
        if (with_lists)
          for (int z = 0; z < n[2]; z++)
            for (int y = 0; y < n[1]; y++)
              for (int x = 0; x < n[0]; x++) {
                P.gcell_ext.push_back((int)ext_index(g, glo[0] + x, glo[1] + y, glo[2] + z));
                const int X = slo[0] + x, Y = slo[1] + y, Z = slo[2] + z;
                P.scell_L.push_back((int)phys_index(g, X / g.nt, Y / g.nt, Z / g.nt, X % g.nt, Y % g.nt, Z % g.nt));
              }
        P.ng += d.ncell;
        P.dirs.push_back(d);
      }
}

struct HaloRec { int rho; float v[3]; };  // 16 bytes per coarse cell: rhoc + vfield travel together

__global__ void __launch_bounds__(256) k_halo_pack(long long ng, const int* __restrict__ scell_L, const int* __restrict__ rhoc_p,
                                                   const float* __restrict__ vfield_p, HaloRec* __restrict__ out, int* __restrict__ scnt) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ng) return;
  const long long L = scell_L[s];
  HaloRec r; r.rho = rhoc_p[L]; r.v[0] = vfield_p[3 * L]; r.v[1] = vfield_p[3 * L + 1]; r.v[2] = vfield_p[3 * L + 2];
  out[s] = r;
  scnt[s] = r.rho;
}
__global__ void __launch_bounds__(256) k_halo_unpack(long long ng, const int* __restrict__ gcell_ext, const HaloRec* __restrict__ in,
                                                     int* __restrict__ rhoc_e, float* __restrict__ vfield_e, int* __restrict__ gcnt,
                                                     int* __restrict__ sid_e, long long ncell_p) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= ng) return;
  const long long e = gcell_ext[q];
  const HaloRec r = in[q];
  rhoc_e[e] = r.rho; vfield_e[3 * e] = r.v[0]; vfield_e[3 * e + 1] = r.v[1]; vfield_e[3 * e + 2] = r.v[2];
  gcnt[q] = r.rho;
  sid_e[e] = (int)(ncell_p + q);
}
// ghost cells point into the received segment behind the physical particles
__global__ void __launch_bounds__(256) k_ghost_cstart(long long ng, const int* __restrict__ gcell_ext, const long long* __restrict__ gstart,
                                                      long long base, long long* __restrict__ cstart_e) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < ng) cstart_e[gcell_ext[q]] = base + gstart[q];
}
// prefix values at the direction boundaries (message offsets), for the host
__global__ void k_dir_bounds(int ndir, const long long* __restrict__ cell0 /*[ndir+1]*/, const long long* __restrict__ gstart,
                             const long long* __restrict__ sstart, long long* __restrict__ out /*[2][ndir+1]*/) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d > ndir) return;
  out[d] = gstart[cell0[d]];
  out[ndir + 1 + d] = sstart[cell0[d]];
}

// pack the particles of the cells I send: a CTA owns PC_CELLS consecutive send cells = one contiguous run of the
// message buffer (coalesced stores); loads are contiguous per cell
template <class T>
__global__ void __launch_bounds__(PC_T) k_particle_pack(long long ng, const int* __restrict__ scell_L, const long long* __restrict__ sstart,
                                                        const long long* __restrict__ cstart_p, const T* __restrict__ arr,
                                                        T* __restrict__ out) {
  __shared__ int soff[PC_CELLS + 1];
  __shared__ long long ssrc[PC_CELLS];
  const long long c0 = (long long)blockIdx.x * PC_CELLS;
  for (int t = threadIdx.x; t < PC_CELLS; t += blockDim.x) ssrc[t] = c0 + t < ng ? cstart_p[scell_L[c0 + t]] : 0;
  const int np = chunk_setup(sstart, c0, ng, soff);
  const long long p0 = sstart[c0];
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const int c = chunk_find(soff, q);
    const Code3 v = load_code3(arr, ssrc[c] + (q - soff[c]));
    store_code3(out, p0 + q, v.x, v.y, v.z);
  }
}

// the IDs of the particles I send (-DPID: buffer_v.f90 carries pid with vp), one 8-byte word per particle
__global__ void __launch_bounds__(PC_T) k_pid_pack(long long ng, const int* __restrict__ scell_L, const long long* __restrict__ sstart,
                                                   const long long* __restrict__ cstart_p, const long long* __restrict__ pid,
                                                   long long* __restrict__ out) {
  __shared__ int soff[PC_CELLS + 1];
  __shared__ long long ssrc[PC_CELLS];
  const long long c0 = (long long)blockIdx.x * PC_CELLS;
  for (int t = threadIdx.x; t < PC_CELLS; t += blockDim.x) ssrc[t] = c0 + t < ng ? cstart_p[scell_L[c0 + t]] : 0;
  const int np = chunk_setup(sstart, c0, ng, soff);
  const long long p0 = sstart[c0];
  for (int q = threadIdx.x; q < np; q += PC_T) {
    const int c = chunk_find(soff, q);
    out[p0 + q] = pid[ssrc[c] + (q - soff[c])];
  }
}

}  // namespace cube
