"""Synthetic LCDM-shaped initial conditions in CUBE's integer checkpoint format.

The reference's ``ic.x`` (CUBE/utilities/initial_conditions.f90) cannot be run here (no Fortran, and its
``seed_N.bin`` files replay gfortran's RNG), so *these* files are the "identical inputs" both the oracle
and the GPU step consume.  Same conventions as initial_conditions.f90:509-580:

* particles start on a simple-cubic lattice, ``np_nc`` per coarse cell per dim, at
  ``q=(i-1)/np_nc + 0.5/ncell`` (:519), displaced by a Zel'dovich field;
* ``xp = floor(frac(x)/x_resolution)`` truncated to int16 (:554);
* ``vfield`` = per-coarse-cell mean velocity, ``vp = nint(N*atan(S*(v-vfield))/pi)`` (:556-557);
* particles are packed tile-major, then k, j, i, in file order (Appendix B of SURVEY.md).

Uses torch only as an array library (CPU for tests, CUDA for the large bench inputs).
"""
from __future__ import annotations

import math

import numpy as np
import torch

PI_F = float(np.float32(4) * np.arctan(np.float32(1)))


def _bbks(k, gamma):
    q = k / gamma
    q = torch.clamp(q, min=1e-12)
    t = torch.log(1 + 2.34 * q) / (2.34 * q) * (1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
    return t


def vfactor(a, omega_m=0.32, omega_l=0.68):
    """initial_conditions.f90:754-763."""
    lm = omega_l / omega_m
    km = (1 - omega_m - omega_l) / omega_m
    H = 2 / (3 * math.sqrt(a ** 3)) * math.sqrt(1 + a * km + a ** 3 * lm)
    return a ** 2 * H


def pack_states(xs, vels, nn, nc, nnt, izipx=2, izipv=2):
    """Particles at global positions ``xs[d]`` (coarse cells, any real; wrapped periodically) with velocities ``vels[d]``
    (f64 torch vectors) -> ``(states, sigma_vi)`` in CUBE's cell-ordered integer format, one state per image.
    ``izipx``/``izipv`` = bytes per position / velocity code (1 or 2; CUBE/main/universe*.fh:2-3): ``xp`` comes back
    int8 or int16, likewise ``vp``."""
    assert izipx in (1, 2) and izipv in (1, 2)
    nxbin, nvbin = 1 << (8 * izipx), 1 << (8 * izipv)
    xdt, vdt = (torch.int8, torch.int16)[izipx - 1], (torch.int8, torch.int16)[izipv - 1]
    nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
    nt = nc // nnt
    dev = xs[0].device
    ncg = [nc * n for n in nn]
    cells, codes = [], []
    for d in range(3):
        x = torch.remainder(xs[d].double(), float(ncg[d]))
        c = torch.floor(x).clamp_(0, ncg[d] - 1)
        u = torch.floor((x - c) * float(nxbin)).clamp_(0, nxbin - 1).to(torch.int64)
        cells.append(c.to(torch.int64))
        codes.append(u)
    vels = [v.double() for v in vels]
    # linear key: image (x fastest), tile (x fastest), k, j, i
    img = [cells[d] // nc for d in range(3)]
    loc = [cells[d] % nc for d in range(3)]
    til = [loc[d] // nt for d in range(3)]
    cel = [loc[d] % nt for d in range(3)]
    m = img[0] + nn[0] * (img[1] + nn[1] * img[2])
    t = til[0] + nnt * (til[1] + nnt * til[2])
    cc = cel[0] + nt * (cel[1] + nt * cel[2])
    ncell_img = nc ** 3
    key = (m * (nnt ** 3) + t) * (nt ** 3) + cc
    del img, loc, til, cel, m, t, cc, cells
    nimg = nn[0] * nn[1] * nn[2]
    order = torch.argsort(key, stable=True)
    key_s = key[order]
    counts = torch.bincount(key_s, minlength=nimg * ncell_img)
    v_s = torch.stack([v[order] for v in vels], 1)          # (N,3) f64
    u_s = torch.stack([c[order] for c in codes], 1)
    del vels, codes, key
    vsum = torch.zeros((nimg * ncell_img, 3), dtype=torch.float64, device=dev)
    vsum.index_add_(0, key_s, v_s)
    vfield = (vsum / counts.clamp(min=1)[:, None].double()).float()
    res = v_s - vfield[key_s].double()
    sigma_vi = np.float32(math.sqrt(float((res ** 2).sum(1).mean())) / math.sqrt(3.0))
    S = float(np.float64(np.sqrt(np.float32(PI_F / 2), dtype=np.float32)) / (np.float64(sigma_vi) * 2.5))
    vp = torch.round(float(nvbin - 1) * torch.atan(S * res) / PI_F).clamp_(-(nvbin // 2 - 1), nvbin // 2 - 1).to(vdt)
    xp = u_s.to(torch.int32)
    xp = torch.where(xp >= nxbin // 2, xp - nxbin, xp).to(xdt)
    bounds = torch.cumsum(counts.view(nimg, -1).sum(1), 0).cpu().numpy()
    starts = np.concatenate([[0], bounds[:-1]])
    states = []
    counts_c = counts.view(nimg, nnt, nnt, nnt, nt, nt, nt).to(torch.int32).cpu().numpy()
    vfield_c = vfield.view(nimg, nnt, nnt, nnt, nt, nt, nt, 3).cpu().numpy()
    xp_c = xp.cpu().numpy(); vp_c = vp.cpu().numpy()
    for mi in range(nimg):
        s, e = int(starts[mi]), int(bounds[mi])
        states.append(dict(xp=np.ascontiguousarray(xp_c[s:e]), vp=np.ascontiguousarray(vp_c[s:e]),
                           rhoc=np.ascontiguousarray(counts_c[mi]), vfield=np.ascontiguousarray(vfield_c[mi])))
    return states, sigma_vi


def make_ic(nn=(1, 1, 1), nc=32, nnt=2, np_nc=2, seed=1, disp_rms=0.6, box_per_image=200.0, z_i=49.0,
            n_s=0.9619, h=0.67, omega_m=0.32, device="cpu", velocity_boost=1.0, izipx=2, izipv=2):
    """Return ``(states, sigma_vi, info)``; ``states[m]`` = dict(xp, vp, rhoc, vfield) numpy arrays in
    file order for image ``m`` (image order x fastest, parameters.f90:200-203).

    ``disp_rms``: rms Zel'dovich displacement per dimension in *fine* cells.
    """
    nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
    ncell = 4
    nt = nc // nnt
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    npd = [np_nc * nc * n for n in nn]            # particles per dim (x,y,z)
    shape = (npd[2], npd[1], npd[0])
    white = torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)
    dk = torch.fft.rfftn(white)
    del white
    # wavenumbers in h/Mpc ; box length per dim = box_per_image*nn_d
    def kvec(n, L, half=False):
        f = torch.fft.rfftfreq(n, d=1.0 / n, device=dev) if half else torch.fft.fftfreq(n, d=1.0 / n, device=dev)
        return (2 * math.pi / L) * f
    kx = kvec(npd[0], box_per_image * nn[0], True)[None, None, :]
    ky = kvec(npd[1], box_per_image * nn[1])[None, :, None]
    kz = kvec(npd[2], box_per_image * nn[2])[:, None, None]
    k2 = kx ** 2 + ky ** 2 + kz ** 2
    k = torch.sqrt(k2)
    pk = torch.where(k > 0, k ** n_s * _bbks(k, omega_m * h) ** 2, torch.zeros_like(k))
    dk = dk * torch.sqrt(pk)
    k2 = torch.where(k2 > 0, k2, torch.ones_like(k2))
    psi = []
    for kd in (kx, ky, kz):
        psi.append(torch.fft.irfftn(dk * (1j * kd / k2), s=shape))
    del dk
    rms = torch.sqrt(sum((p.double() ** 2).mean() for p in psi) / 3).item()
    scale = disp_rms / rms
    a = 1.0 / (1.0 + z_i)
    vf = vfactor(a, omega_m, 1 - omega_m) * velocity_boost
    # positions in coarse cells (global), float64
    idx = [torch.arange(npd[d], device=dev, dtype=torch.float64) for d in range(3)]
    q = [idx[d] / np_nc + 0.5 / ncell for d in range(3)]
    qb = (q[0][None, None, :], q[1][None, :, None], q[2][:, None, None])
    xs = [(qb[d] + psi[d].double() * (scale / ncell)).reshape(-1) for d in range(3)]
    vels = [(psi[d].double() * (scale * vf)).reshape(-1) for d in range(3)]
    del psi
    states, sigma_vi = pack_states(xs, vels, nn, nc, nnt, izipx, izipv)
    npglobal = sum(int(st["xp"].shape[0]) for st in states)
    info = dict(nn=nn, nc=nc, nnt=nnt, np_nc=np_nc, a=a, vf=vf, npglobal=npglobal, seed=seed,
                disp_rms=disp_rms)
    return states, sigma_vi, info


def make_clustered_ic(nn=(1, 1, 1), nc=32, nnt=2, np_nc=2, seed=1, nblob=6, blob_fraction=0.5, blob_sigma=0.35, v_rms=0.3,
                      device="cpu", izipx=2, izipv=2):
    """A late-time-like state for the tests of the crowded-cell paths: ``blob_fraction`` of the ``(np_nc nc)^3`` particles per
    image sit in ``nblob`` Gaussian clumps per image (``blob_sigma`` coarse cells wide, so single coarse cells hold hundreds
    of particles, next to empty ones), the rest are uniform; velocities = a per-clump bulk flow + Gaussian dispersion
    ``v_rms``.  Returns ``(states, sigma_vi, info)`` like ``make_ic``."""
    nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    nimg = nn[0] * nn[1] * nn[2]
    ntot = (np_nc * nc) ** 3 * nimg
    nb = int(ntot * blob_fraction)
    L = [float(nc * n) for n in nn]
    which = torch.randint(0, nblob * nimg, (nb,), generator=gen, device=dev)
    xs, vels = [], []
    for d in range(3):
        centre = torch.rand(nblob * nimg, generator=gen, device=dev, dtype=torch.float64) * L[d]
        bulk = torch.randn(nblob * nimg, generator=gen, device=dev, dtype=torch.float64) * v_rms
        xb = centre[which] + torch.randn(nb, generator=gen, device=dev, dtype=torch.float64) * blob_sigma
        xu = torch.rand(ntot - nb, generator=gen, device=dev, dtype=torch.float64) * L[d]
        xs.append(torch.cat([xb, xu]))
        v = torch.randn(ntot, generator=gen, device=dev, dtype=torch.float64) * v_rms
        v[:nb] += bulk[which]
        vels.append(v)
    states, sigma_vi = pack_states(xs, vels, nn, nc, nnt, izipx, izipv)
    info = dict(nn=nn, nc=nc, nnt=nnt, np_nc=np_nc, npglobal=ntot, seed=seed, rhoc_max=max(int(st["rhoc"].max()) for st in states))
    return states, sigma_vi, info


def tile_state(st, nnt, reps):
    """Periodic replication of one image's state ``reps`` times per dimension: the state of an image with ``reps*nnt`` tiles
    per dimension (same ``nt``), built on the host from per-tile runs.  Lets the HBM-sized configurations (1024^3 particles per
    GPU) be loaded without generating a 1024^3 field first (the torch generator above needs ~170 GB of temporaries for that)."""
    rhoc, vfield = st["rhoc"], st["vfield"]
    per_tile = rhoc.reshape(nnt ** 3, -1).sum(1, dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(per_tile)])
    big = nnt * reps
    order = []
    for tz in range(big):
        for ty in range(big):
            for tx in range(big):
                order.append(((tz % nnt) * nnt + ty % nnt) * nnt + tx % nnt)
    n = int(per_tile[order].sum())
    xp = np.empty((n, 3), st["xp"].dtype); vp = np.empty((n, 3), st["vp"].dtype)
    o = 0
    for t in order:
        m = int(per_tile[t])
        xp[o:o + m] = st["xp"][start[t]:start[t] + m]; vp[o:o + m] = st["vp"][start[t]:start[t] + m]
        o += m
    return dict(xp=xp, vp=vp, rhoc=np.ascontiguousarray(np.tile(rhoc, (reps, reps, reps, 1, 1, 1))),
                vfield=np.ascontiguousarray(np.tile(vfield, (reps, reps, reps, 1, 1, 1, 1))))
