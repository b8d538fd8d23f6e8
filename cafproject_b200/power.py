"""CIC density contrast and (cross) power spectrum of checkpoints: CUBE/utilities/cicpower.f90:70-140 and
CUBE/utilities/powerspectrum.f90:21-108 (``linear_kbin``), for all images of a run at once.  Analysis utility ("next"
row f1 of SURVEY.md sec. 8): NumPy, sized for the parity tests' grids; used for the z=0 P(k) gate.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def cic_delta(states, nn, nc, nnt, ng_per_nc=4):
    """Density contrast ``delta_N`` on the global grid ``ng_global = ng_per_nc*nc*nn`` from the disjoint states
    (``states[m] = dict(xp int16 (n,3), rhoc (nnt,nnt,nnt,nt,nt,nt))``, image order x fastest).  numpy ``[z][y][x]``."""
    nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
    nt = nc // nnt
    G = [ng_per_nc * nc * n for n in nn]
    rho = np.zeros(G[2] * G[1] * G[0], np.float64)
    for m, st in enumerate(states):
        ic = (m % nn[0], (m // nn[0]) % nn[1], m // (nn[0] * nn[1]))
        cnt = np.asarray(st["rhoc"]).reshape(-1).astype(np.int64)
        L = np.repeat(np.arange(cnt.size, dtype=np.int64), cnt)            # file-order cell of every particle
        nt3 = nt ** 3
        t, c = L // nt3, L % nt3
        cell = [(t % nnt) * nt + c % nt, ((t // nnt) % nnt) * nt + (c // nt) % nt, (t // (nnt * nnt)) * nt + c // (nt * nt)]
        u = np.asarray(st["xp"]).astype(np.int64) & 0xFFFF                 # int(xp+ishift,izipx)+rshift = u + 0.5
        idx, w = [], []
        for d in range(3):
            pos = (cell[d] + ic[d] * nc + (u[:, d] + 0.5) / 65536.0) * ng_per_nc - 0.5     # cicpower.f90:84-85, global
            i1 = np.floor(pos).astype(np.int64)
            dx1 = (i1 + 1) - pos
            idx.append((np.mod(i1, G[d]), np.mod(i1 + 1, G[d])))
            w.append((dx1, 1.0 - dx1))
        for qz in (0, 1):
            for qy in (0, 1):
                for qx in (0, 1):
                    flat = (idx[2][qz] * G[1] + idx[1][qy]) * G[0] + idx[0][qx]
                    rho += np.bincount(flat, weights=w[0][qx] * w[1][qy] * w[2][qz], minlength=rho.size)
    rho = rho.reshape(G[2], G[1], G[0])
    return (rho / rho.mean() - 1.0).astype(F32)                            # cicpower.f90:140


def cross_power(d1, d2, box):
    """``xi(10,nbin)`` of powerspectrum.f90:47-108 for two density contrasts on the same (cubic) grid: rows
    0 count, 1 k [h/Mpc], 2 Delta^2_11, 3 Delta^2_22, 4 Delta^2_12, 5-6 kernels, 7 r, 8 b, 9 reco power."""
    n = d1.shape[0]
    assert d1.shape == d2.shape == (n, n, n)
    nyq = n // 2
    nbin = int(round(nyq * np.sqrt(3.0)))
    c1, c2 = np.fft.rfftn(d1.astype(np.float64)), np.fft.rfftn(d2.astype(np.float64))
    kf = np.mod(np.arange(n) + nyq, n) - nyq                                # mod((/ig,jg,kg/)+nyquest-1,ng_global)-nyquest
    kx = np.arange(nyq + 1, dtype=np.float64)[None, None, :]
    ky = kf.astype(np.float64)[None, :, None]
    kz = kf.astype(np.float64)[:, None, None]
    kr = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    keep = np.ones(kr.shape, bool)
    ig = np.arange(nyq + 1)[None, None, :]; jg = np.arange(n)[None, :, None]; kg = np.arange(n)[:, None, None]
    edge = (ig == 0) | (ig == nyq)
    keep &= ~((ig == 0) & (jg == 0) & (kg == 0))                            # powerspectrum.f90:56-58
    keep &= ~(edge & (jg > nyq))
    keep &= ~(edge & ((jg == 0) | (jg == nyq)) & (kg > nyq))
    sinc = np.sinc(kx / n) * np.sinc(ky / n) * np.sinc(kz / n)              # np.sinc(x) = sin(pi x)/(pi x)
    ibin = np.rint(kr).astype(np.int64)                                      # linear_kbin: ibin=nint(kr)
    norm = 4 * np.pi * kr ** 3 / float(n) ** 6 / sinc ** 4
    sel = keep & (ibin >= 1) & (ibin <= nbin)
    b = ibin[sel] - 1
    xi = np.zeros((10, nbin))
    xi[0] = np.bincount(b, minlength=nbin)
    xi[1] = np.bincount(b, weights=kr[sel], minlength=nbin)
    xi[2] = np.bincount(b, weights=(c1 * np.conj(c1)).real[sel] * norm[sel], minlength=nbin)
    xi[3] = np.bincount(b, weights=(c2 * np.conj(c2)).real[sel] * norm[sel], minlength=nbin)
    xi[4] = np.bincount(b, weights=(c1 * np.conj(c2)).real[sel] * norm[sel], minlength=nbin)
    xi[5] = np.bincount(b, weights=(1 / sinc ** 2)[sel], minlength=nbin)
    xi[6] = np.bincount(b, weights=(1 / sinc ** 4)[sel], minlength=nbin)
    with np.errstate(divide="ignore", invalid="ignore"):
        cnt = xi[0]
        xi[1] = xi[1] / cnt * (2 * np.pi) / box
        for r in (2, 3, 4, 5, 6):
            xi[r] = xi[r] / cnt
        xi[7] = xi[4] / np.sqrt(xi[2] * xi[3])
        xi[8] = np.sqrt(xi[3] / xi[2])
        xi[9] = xi[7] ** 4 / xi[8] ** 2 * xi[3]
    return xi


# ---------------------------------------------------------------------------------------------
# The same estimator with torch as the array library: runs on the GPU at the bench scale (ng_global = 1024 at cfg 2, where the
# numpy version needs minutes and 70 GB of f64 temporaries); on the CPU it is checked against the numpy functions above
# (tests/test_oracle_pins.py).  SURVEY.md sec. 8 row f1.
# ---------------------------------------------------------------------------------------------
def cic_delta_torch(states, nn, nc, nnt, ng_per_nc=4, device="cpu"):
    """``cic_delta`` on ``device``: f64 accumulation by ``index_add_`` per corner, f32 contrast out (torch tensor [z][y][x])."""
    import torch
    nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
    nt = nc // nnt
    G = [ng_per_nc * nc * n for n in nn]
    dev = torch.device(device)
    rho = torch.zeros(G[2] * G[1] * G[0], dtype=torch.float64, device=dev)
    nt3 = nt ** 3
    for m, st in enumerate(states):
        ic = (m % nn[0], (m // nn[0]) % nn[1], m // (nn[0] * nn[1]))
        cnt = torch.as_tensor(np.ascontiguousarray(st["rhoc"]).reshape(-1), device=dev).to(torch.int64)
        L = torch.repeat_interleave(torch.arange(cnt.numel(), device=dev, dtype=torch.int64), cnt)
        t, c = L // nt3, L % nt3
        cell = [(t % nnt) * nt + c % nt, ((t // nnt) % nnt) * nt + (c // nt) % nt, (t // (nnt * nnt)) * nt + c // (nt * nt)]
        del L, t, c
        u = torch.as_tensor(np.ascontiguousarray(st["xp"]), device=dev).to(torch.int64) & 0xFFFF
        idx, w = [], []
        for d in range(3):
            pos = (cell[d].double() + float(ic[d] * nc) + (u[:, d].double() + 0.5) / 65536.0) * ng_per_nc - 0.5
            i1 = torch.floor(pos).to(torch.int64)
            dx1 = (i1 + 1).double() - pos
            idx.append((torch.remainder(i1, G[d]), torch.remainder(i1 + 1, G[d])))
            w.append((dx1, 1.0 - dx1))
        del cell, u
        for qz in (0, 1):
            for qy in (0, 1):
                for qx in (0, 1):
                    flat = (idx[2][qz] * G[1] + idx[1][qy]) * G[0] + idx[0][qx]
                    rho.index_add_(0, flat, w[0][qx] * w[1][qy] * w[2][qz])
    rho = rho.view(G[2], G[1], G[0])
    return (rho / rho.mean() - 1.0).float()


def cross_power_torch(d1, d2, box):
    """``cross_power`` for torch tensors on any device (f32 transforms, f64 binning); returns the numpy ``xi(10,nbin)``."""
    import torch
    n = d1.shape[0]
    assert tuple(d1.shape) == tuple(d2.shape) == (n, n, n)
    dev = d1.device
    nyq = n // 2
    nbin = int(round(nyq * np.sqrt(3.0)))
    same = d2 is d1
    c1 = torch.fft.rfftn(d1.float())
    c2 = c1 if same else torch.fft.rfftn(d2.float())
    ar = torch.arange(n, device=dev)
    kf = (torch.remainder(ar + nyq, n) - nyq).double()
    kx = torch.arange(nyq + 1, device=dev, dtype=torch.float64)[None, None, :]
    ky, kz = kf[None, :, None], kf[:, None, None]
    ig = torch.arange(nyq + 1, device=dev)[None, None, :]; jg = ar[None, :, None]; kg = ar[:, None, None]
    edge = (ig == 0) | (ig == nyq)
    keep = ~((ig == 0) & (jg == 0) & (kg == 0))                             # powerspectrum.f90:56-58
    keep = keep & ~(edge & (jg > nyq)) & ~(edge & ((jg == 0) | (jg == nyq)) & (kg > nyq))
    xi = np.zeros((10, nbin))

    def binsum(wgt, b, sel):
        out = torch.zeros(nbin, dtype=torch.float64, device=dev)
        out.index_add_(0, b, wgt[sel].double())
        return out.cpu().numpy()

    kr = torch.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
    ibin = torch.round(kr).to(torch.int64)                                   # linear_kbin: ibin=nint(kr) (no exact .5: kr^2 is an integer)
    sel = keep & (ibin >= 1) & (ibin <= nbin)
    b = ibin[sel] - 1
    del ibin, keep
    sinc = torch.sinc(kx / n) * torch.sinc(ky / n) * torch.sinc(kz / n)
    norm = 4 * np.pi * kr ** 3 / float(n) ** 6 / sinc ** 4
    xi[0] = binsum(torch.ones_like(kr), b, sel)
    xi[1] = binsum(kr, b, sel)
    xi[2] = binsum((c1.real.double() ** 2 + c1.imag.double() ** 2) * norm, b, sel)
    xi[3] = xi[2] if same else binsum((c2.real.double() ** 2 + c2.imag.double() ** 2) * norm, b, sel)
    xi[4] = xi[2] if same else binsum((c1.real.double() * c2.real.double() + c1.imag.double() * c2.imag.double()) * norm, b, sel)
    xi[5] = binsum(1 / sinc ** 2, b, sel)
    xi[6] = binsum(1 / sinc ** 4, b, sel)
    with np.errstate(divide="ignore", invalid="ignore"):
        cnt = xi[0].copy()
        xi[1] = xi[1] / cnt * (2 * np.pi) / box
        for r in (2, 3, 4, 5, 6):
            xi[r] = xi[r] / cnt
        xi[7] = xi[4] / np.sqrt(xi[2] * xi[3])
        xi[8] = np.sqrt(xi[3] / xi[2])
        xi[9] = xi[7] ** 4 / xi[8] ** 2 * xi[3]
    return xi
