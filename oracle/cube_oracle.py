"""CPU oracle for CUBE's particle-mesh step: Python driver around ``cube_oracle.c``.

TEST INFRASTRUCTURE ONLY -- see the header of ``cube_oracle.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this module.  Parity is *unpinned by the reference* (no golden vectors exist upstream; the Fortran
cannot be built here); the restatement follows, with file:line citations in the code:

* ``CUBE/main/kernel_f.f90:14-41``      -> :func:`kernel_f`
* ``CUBE/main/kernel_c.f90:16-125``     -> :func:`kernel_c`
* ``CUBE/main/pm.f90:27-245``           -> :meth:`Oracle.particle_mesh`
* ``CUBE/main/update_particle.f90``     -> :meth:`Oracle.update_particle` (C)
* ``CUBE/main/buffer_density.f90``, ``buffer_x.f90``, ``buffer_v.f90`` -> C
* ``CUBE/main/timestep.f90:1-133``      -> :class:`TimeStepper`
* ``CUBE/main/pencil_fft.f90:31-60``    -> plain 3-D r2c/c2r of the *global* coarse grid (all images
  live in this process), unnormalised forward, ``/ng_global`` three times after the inverse.

FFTW is replaced by ``scipy.fft`` (pocketfft, single precision): the DFT definition is the same, the
round-off is not, hence the 1e-5 norm-relative tolerance on densities/forces.

Array convention: a Fortran array ``A(i,j,k)`` is a C-ordered numpy array ``A[k-1,j-1,i-1]`` (same
memory layout); a leading component index ``F(3,i,j,k)`` becomes ``F[k,j,i,3]``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

try:  # pocketfft, f32-preserving, multi-threaded
    import scipy.fft as _fft
except Exception:  # pragma: no cover
    _fft = None

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libcube_oracle.so")

F32 = np.float32
PI_F = F32(4) * np.arctan(F32(1.0), dtype=F32)  # parameters.f90:73 -> 0x40490FDB
assert PI_F.view(np.uint32) == 0x40490FDB


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "cube_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        i64, f32, vp = C.c_int64, C.c_float, C.c_void_p
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i64] * 6 + [f32, f32]
        L.oracle_destroy.argtypes = [vp]
        for name in ("np_image_max", "np_tile_max", "npglobal"):
            getattr(L, "oracle_" + name).restype = i64
            getattr(L, "oracle_" + name).argtypes = [vp]
        for name in ("xp", "vp", "rhoc", "vfield", "cum", "f2_max_fine"):
            getattr(L, "oracle_" + name).restype = vp
            getattr(L, "oracle_" + name).argtypes = [vp, i64]
        L.oracle_nplocal.restype = i64
        L.oracle_nplocal.argtypes = [vp, i64]
        for name in ("get_sigma_vi", "get_sigma_vi_new", "mass_p", "overhead_tile", "overhead_image"):
            getattr(L, "oracle_" + name).restype = f32
            getattr(L, "oracle_" + name).argtypes = [vp]
        L.oracle_set_sigma_vi.argtypes = [vp, f32]
        L.oracle_set_mass_p.argtypes = [vp, f32]
        L.oracle_set_sigma_vi_new.argtypes = [vp, f32]
        L.oracle_vmax.restype = f32
        L.oracle_vmax.argtypes = [vp, i64]
        L.oracle_vmax3.restype = f32
        L.oracle_vmax3.argtypes = [vp, i64, C.c_int]
        L.oracle_set_nlayer.argtypes = [vp, i64]
        L.oracle_nlayer_for.restype = i64
        L.oracle_nlayer_for.argtypes = [vp, f32, f32, f32]
        L.oracle_f2_max_coarse.restype = f32
        L.oracle_f2_max_coarse.argtypes = [vp, i64]
        L.oracle_error.restype = C.c_int
        L.oracle_error.argtypes = [vp]
        L.oracle_errmsg.restype = C.c_char_p
        L.oracle_errmsg.argtypes = [vp]
        L.oracle_std_vsim.restype = C.c_double
        L.oracle_std_vsim.argtypes = [vp, C.c_int]
        L.oracle_tanf_lut.argtypes = [vp]
        L.oracle_tanf_lut_zip.argtypes = [vp, i64]
        L.oracle_set_zip.restype = C.c_int
        L.oracle_set_zip.argtypes = [vp, i64, i64]
        L.oracle_probe_xq.restype = C.c_double
        L.oracle_probe_xq.argtypes = [vp, i64, i64]
        L.oracle_probe_vdecode.restype = C.c_double
        L.oracle_probe_vdecode.argtypes = [vp, i64, f32]
        L.oracle_probe_vencode.restype = i64
        L.oracle_probe_vencode.argtypes = [vp, C.c_double, f32]
        L.oracle_load_image.argtypes = [vp, i64, vp, vp, vp, vp, i64]
        L.oracle_finish_load.argtypes = [vp, f32]
        L.oracle_load_pid.argtypes = [vp, i64, vp]
        L.oracle_pid.restype = vp
        L.oracle_pid.argtypes = [vp, i64]
        L.oracle_store_image.argtypes = [vp, i64, vp, vp]
        for name in ("buffer_density", "buffer_x", "buffer_v", "pm_begin", "pm_fine_end"):
            getattr(L, "oracle_" + name).argtypes = [vp]
        L.oracle_update_particle.argtypes = [vp, f32, f32]
        L.oracle_fine_deposit.argtypes = [vp, i64, i64, i64, i64, vp]
        L.oracle_fine_kick.argtypes = [vp, i64, i64, i64, i64, vp, f32, f32]
        L.oracle_coarse_deposit.argtypes = [vp, i64, vp]
        L.oracle_coarse_kick.argtypes = [vp, i64, vp, f32, f32]
        _lib = L
    return _lib


def tanf_lut(izipv: int = 2) -> np.ndarray:
    """``2^(8*izipv)`` host-libm values ``tanf((pi_f*float(code))/float(nvbin-1))`` indexed by the raw unsigned code
    (65536 entries for the 2-byte format the product is built for)."""
    out = np.empty(1 << (8 * izipv), F32)
    if izipv == 2:
        lib().oracle_tanf_lut(out.ctypes.data)
    else:
        lib().oracle_tanf_lut_zip(out.ctypes.data, izipv)
    return out


def _ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _workers() -> int:
    return int(os.environ.get("CUBE_ORACLE_THREADS", os.cpu_count() or 1))


def rfftn(a):
    return _fft.rfftn(a, workers=_workers())


def irfftn_unnorm(c, shape):
    # FFTW c2r is unnormalised: norm="forward" leaves the backward transform unscaled
    return _fft.irfftn(c, s=shape, norm="forward", workers=_workers())


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
def kernel_f(fk_table: np.ndarray, nfe: int) -> np.ndarray:
    """``kern_f(nfe/2+1,nfe,nfe,3)`` -> numpy ``[3][kz][ky][kx]`` (CUBE/main/kernel_f.f90:32-41).

    ``fk_table`` is ``wfxyzf.3.ascii`` as ``[k][j][i][dim]`` (16,16,16,3) f32.
    """
    ncut = 16  # nf_cutoff, parameters.f90:51
    out = np.empty((3, nfe, nfe, nfe // 2 + 1), F32)
    for d in range(3):
        rho = np.zeros((nfe, nfe, nfe + 2), F32)
        mf = [F32(-1) if dd == d else F32(1) for dd in range(3)]  # mfactor :34
        rho[:ncut, :ncut, :ncut] = fk_table[:, :, :, d]
        # rho_f(nfe-nf_cutoff+2:nfe,:,:)=mfactor(1)*rho_f(nf_cutoff:2:-1,:,:)  (x is the last numpy axis)
        rho[:, :, nfe - ncut + 1:nfe] = mf[0] * rho[:, :, ncut - 1:0:-1]
        rho[:, nfe - ncut + 1:nfe, :] = mf[1] * rho[:, ncut - 1:0:-1, :]
        rho[nfe - ncut + 1:nfe, :, :] = mf[2] * rho[ncut - 1:0:-1, :, :]
        out[d] = rfftn(rho[:, :, :nfe]).imag  # kern_f(:,:,:,i_dim)=rho_f(2::2,:,:)
    return out


def _signed_index(n: int) -> np.ndarray:
    # mod((/ig/)+ncglobal/2-1,ncglobal)-ncglobal/2 with 1-based ig  == mod(g+n/2,n)-n/2 with 0-based g
    g = np.arange(n)
    return (np.mod(g + n // 2, n) - n // 2)


def kernel_c(ck_table: np.ndarray, ncg, ncell: int = 4, lrckcorr: bool = True) -> np.ndarray:
    """Global coarse kernel ``kern_c`` as ``[3][kz][ky][kx]`` on the full ``ncg`` grid, half x.

    Restates CUBE/main/kernel_c.f90:16-117 for all images at once (the per-image ``if (icx==..)``
    blocks place the 4^3 corrections at the eight corners of the *global* lattice).  ``ncg`` may be a
    scalar (cubic, the reference) or (ncgx, ncgy, ncgz).  ``ck_table`` is ``wfxyzc.2.ascii`` as
    ``[k][j][i][dim]`` (4,4,4,3).
    """
    if np.isscalar(ncg):
        ncg = (int(ncg),) * 3
    nx, ny, nz = (int(v) for v in ncg)
    rx = (F32(ncell) * _signed_index(nx).astype(F32))[None, None, :]
    ry = (F32(ncell) * _signed_index(ny).astype(F32))[None, :, None]
    rz = (F32(ncell) * _signed_index(nz).astype(F32))[:, None, None]
    r = np.sqrt(rx ** 2 + ry ** 2 + rz ** 2, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        r3 = r ** 3
        base = np.stack([np.where(r == 0, F32(0), -np.broadcast_to(c, r.shape) / r3) for c in (rx, ry, rz)]).astype(F32)
    ck = base.copy()  # [d][z][y][x]
    t = np.ascontiguousarray(np.moveaxis(ck_table.astype(F32), 3, 0))  # [d][k][j][i]
    lo = slice(0, 4)
    hx, hy, hz = slice(nx - 3, nx), slice(ny - 3, ny), slice(nz - 3, nz)
    rev = slice(3, 0, -1)  # 4:2:-1
    sgn = lambda flip: np.array([-1 if f else 1 for f in flip], F32)[:, None, None, None]
    # octants, kernel_c.f90:43-72; sign flips on the components whose axis is mirrored
    ck[:, lo, lo, lo] = t
    ck[:, lo, lo, hx] = sgn((1, 0, 0)) * t[:, :, :, rev]
    ck[:, lo, hy, lo] = sgn((0, 1, 0)) * t[:, :, rev, :]
    ck[:, hz, lo, lo] = sgn((0, 0, 1)) * t[:, rev, :, :]
    ck[:, hz, hy, lo] = sgn((0, 1, 1)) * t[:, rev, rev, :]
    ck[:, hz, lo, hx] = sgn((1, 0, 1)) * t[:, rev, :, rev]
    ck[:, lo, hy, hx] = sgn((1, 1, 0)) * t[:, :, rev, rev]
    ck[:, hz, hy, hx] = sgn((1, 1, 1)) * t[:, rev, rev, rev]
    kern = np.stack([rfftn(ck[d]).imag.astype(F32) for d in range(3)])
    if not lrckcorr:
        return np.stack([rfftn(base[d]).imag.astype(F32) for d in range(3)])
    # LRCKCORR, kernel_c.f90:76-117
    kx = _signed_index(nx).astype(F32)[: nx // 2 + 1][None, None, :]
    ky = _signed_index(ny).astype(F32)[None, :, None]
    kz = _signed_index(nz).astype(F32)[:, None, None]
    kr = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2, dtype=F32)
    ks = [F32(2) * np.sin(PI_F * k / F32(n), dtype=F32) for k, n in ((kx, nx), (ky, ny), (kz, nz))]
    ssum = (ks[0] ** 2 + ks[1] ** 2 + ks[2] ** 2).astype(F32)
    kk = (kx, ky, kz)
    for d in range(3):
        im0 = rfftn(base[d]).imag.astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            corr = kern[d] * F32(0.25) * PI_F * ks[d] / ssum / im0
        keep = (kr > F32(8.0)) | (np.broadcast_to(kk[d], kr.shape) == 0)
        kern[d] = np.where(keep, kern[d], corr).astype(F32)
    return kern


# ---------------------------------------------------------------------------------------------
# timestep.f90
# ---------------------------------------------------------------------------------------------
@dataclass
class Cosmology:
    """parameters.f90:63-86 (f32 parameters)."""
    z_i: float = 49.0
    box: float = 200.0
    h0: float = 67.0
    omega_c: float = 0.27
    omega_b: float = 0.05
    wde: float = -1.0
    ra_max: float = 0.2
    dt_max: float = 1.0

    @property
    def omega_m(self):
        return F32(self.omega_c) + F32(self.omega_b)

    @property
    def omega_l(self):
        return F32(1) - self.omega_m


def expansion(cos: Cosmology, a0, dt0):
    """timestep.f90:89-133.  f32 in/out, f64 inside."""
    a0, dt0 = F32(a0), F32(dt0)
    om, ol, wde = cos.omega_m, cos.omega_l, F32(cos.wde)
    dt_x = F32(dt0 / F32(2))
    dt_x2 = F32(dt_x * dt_x)          # dt_x**2 and dt_x**3 are real(4) expressions
    dt_x3 = F32(dt_x2 * dt_x)

    def half(a_x):
        omHsq = np.float64(F32(4.0) / F32(9.0))
        a3rlm = a_x ** np.float64(F32(-3) * wde) * np.float64(ol) / np.float64(om)
        arkm = a_x * np.float64(F32(1.0) - om - ol) / np.float64(om)
        adot = np.sqrt(omHsq * (a_x * a_x * a_x) * (1.0 + arkm + a3rlm))
        addot = (a_x * a_x) * omHsq * (1.5 + 2.0 * arkm + np.float64(F32(1.5) * (F32(1.0) - wde)) * a3rlm)
        atdot = a_x * adot * omHsq * (3.0 + 6.0 * arkm + np.float64(F32(1.5) * (F32(2.0) - F32(3.0) * wde) * (F32(1.0) - wde)) * a3rlm)
        return F32(adot * np.float64(dt_x) + (addot * np.float64(dt_x2)) / 2.0 + (atdot * np.float64(dt_x3)) / 6.0)

    da1 = half(np.float64(a0))
    da2 = half(np.float64(F32(a0 + da1)))  # a_x=a0+da1 evaluated in f32 then promoted
    return da1, da2


class TimeStepper:
    """Host-side scalar controller, CUBE/main/timestep.f90:1-86 + initialize.f90:25-37."""

    def __init__(self, cos: Cosmology, z_checkpoint):
        self.cos = cos
        self.z_checkpoint = [F32(z) for z in z_checkpoint]
        self.a = F32(1) / (F32(1) + F32(cos.z_i))
        self.dt = F32(0); self.dt_old = F32(0); self.da = F32(0)
        self.tau = F32(-3) / np.sqrt(self.a); self.t = F32(0)
        self.dt_fine = self.dt_coarse = self.dt_pp = self.dt_vmax = F32(1000)
        self.cur_checkpoint = 0
        self.checkpoint_step = False
        self.final_step = False
        self.a_mid = self.a
        self.istep = 0

    def step(self):
        c = self.cos
        self.dt_old = self.dt
        dt_e = F32(c.dt_max)
        ntemp = 0
        while True:
            ntemp += 1
            da1, da2 = expansion(c, self.a, dt_e)
            da = F32(da1 + da2)
            ra = F32(da / F32(self.a + da))
            if ra > F32(c.ra_max):
                dt_e = F32(dt_e * F32(F32(c.ra_max) / ra))
            else:
                break
            if ntemp > 10:
                break
        dt = min(dt_e, self.dt_fine, self.dt_coarse, self.dt_pp, self.dt_vmax)
        da1, da2 = expansion(c, self.a, dt)
        da = F32(da1 + da2)
        self.checkpoint_step = False
        a_chk = F32(1.0) / F32(F32(1) + self.z_checkpoint[self.cur_checkpoint])
        if da >= F32(a_chk - self.a):
            self.checkpoint_step = True
            if self.cur_checkpoint == len(self.z_checkpoint) - 1:
                self.final_step = True
            guard = 0
            while abs(F32(F32(self.a + da) / a_chk) - F32(1)) >= F32(1e-6) and guard < 100:
                dt = F32(F32(dt * F32(a_chk - self.a)) / da)
                da1, da2 = expansion(c, self.a, dt)
                da = F32(da1 + da2)
                guard += 1
        self.a_mid = F32(self.a + F32(da / F32(2)))
        self.dt = F32(dt)
        self.da = da
        self.tau = F32(self.tau + dt); self.t = F32(self.t + dt)
        self.a = F32(self.a + da)
        self.istep += 1
        return self.dt_old, self.dt, self.a_mid

    def after_checkpoint(self):  # cafcube.f90:40-42
        self.cur_checkpoint += 1
        self.checkpoint_step = False
        self.dt = F32(0)


# ---------------------------------------------------------------------------------------------
# the simulation object
# ---------------------------------------------------------------------------------------------
class Oracle:
    """All images of one CUBE run in one process.  Geometry mirrors parameters.f90:20-60."""

    def __init__(self, nn=1, nnt=2, nc=32, np_nc=2, image_buffer=1.5, tile_buffer=2.5,
                 fk_table=None, ck_table=None, izipx=2, izipv=2):
        self.izipx, self.izipv = int(izipx), int(izipv)   # universe*.fh:2-3: bytes per position / velocity code
        self.nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
        self.nnt, self.nc, self.np_nc = int(nnt), int(nc), int(np_nc)
        assert nc % nnt == 0
        self.nt = nc // nnt
        self.ncell, self.ncb = 4, 6
        self.nte = self.nt + 2 * self.ncb
        self.nft = self.nt * self.ncell
        self.nfb = self.ncb * self.ncell
        self.nfe = self.nft + 2 * self.nfb
        self.nimg = self.nn[0] * self.nn[1] * self.nn[2]
        L = lib()
        self.h = L.oracle_create(self.nn[0], self.nn[1], self.nn[2], self.nnt, self.nc, self.np_nc,
                                 F32(image_buffer), F32(tile_buffer))
        if L.oracle_set_zip(self.h, self.izipx, self.izipv):
            raise ValueError("izipx, izipv must be 1 or 2")
        self.np_image_max = L.oracle_np_image_max(self.h)
        self.np_tile_max = L.oracle_np_tile_max(self.h)
        self.kern_f = None if fk_table is None else kernel_f(fk_table, self.nfe)
        ncg = tuple(self.nc * n for n in self.nn)
        self.ncg = ncg
        self.kern_c = None if ck_table is None else kernel_c(ck_table, ncg, self.ncell)
        self.dt_fine = self.dt_coarse = self.dt_vmax = F32(1000)
        self.dt_pp = F32(1000)
        self.last = {}
        self.has_pid = False

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- views --------------------------------------------------------------------------------
    def _view(self, name, m, shape, dtype):
        p = getattr(lib(), "oracle_" + name)(self.h, m)
        n = int(np.prod(shape))
        return np.ctypeslib.as_array((np.ctypeslib.as_ctypes_type(dtype) * n).from_address(p)).reshape(shape)

    def xp(self, m=0):
        return self._view("xp", m, (self.np_image_max, 3), np.int16)

    def vp(self, m=0):
        return self._view("vp", m, (self.np_image_max, 3), np.int16)

    def pid(self, m=0):
        """Particle IDs (-DPID), same slots as ``vp``; only after a ``load`` whose states carried ``pid``."""
        assert self.has_pid
        return self._view("pid", m, (self.np_image_max,), np.int64)

    def rhoc(self, m=0):
        t, e = self.nnt, self.nte
        return self._view("rhoc", m, (t, t, t, e, e, e), np.int32)

    def vfield(self, m=0):
        t, e = self.nnt, self.nte
        return self._view("vfield", m, (t, t, t, e, e, e, 3), np.float32)

    def cum(self, m=0):
        t, e = self.nnt, self.nte
        return self._view("cum", m, (t, t, t, e, e, e), np.int64)

    def nplocal(self, m=0):
        return lib().oracle_nplocal(self.h, m)

    @property
    def sigma_vi(self):
        return F32(lib().oracle_get_sigma_vi(self.h))

    @property
    def sigma_vi_new(self):
        return F32(lib().oracle_get_sigma_vi_new(self.h))

    @property
    def mass_p(self):
        return F32(lib().oracle_mass_p(self.h))

    @property
    def npglobal(self):
        return lib().oracle_npglobal(self.h)

    def _check(self):
        if lib().oracle_error(self.h):
            raise RuntimeError(lib().oracle_errmsg(self.h).decode())

    # ---- state in/out (disjoint state, file order) ------------------------------------------
    def load(self, states, sigma_vi):
        """``states[m] = dict(xp, vp, rhoc, vfield)`` in file order (particle_initialization.f90)."""
        for m, s in enumerate(states):
            for name, z in (("xp", self.izipx), ("vp", self.izipv)):   # particle_initialization.f90:14 "zip format incompatable"
                assert np.asarray(s[name]).dtype == (np.int8, np.int16)[z - 1], f"{name} must be int{8 * z} for izip={z}"
            # 1-byte codes are held sign-extended in the oracle's int16 slots (cube_oracle.c, to_kind)
            xp = np.ascontiguousarray(s["xp"], np.int16); vp = np.ascontiguousarray(s["vp"], np.int16)
            rc = np.ascontiguousarray(s["rhoc"], np.int32); vf = np.ascontiguousarray(s["vfield"], np.float32)
            n = xp.shape[0]
            assert n == int(rc.sum()) and n <= self.np_image_max
            lib().oracle_load_image(self.h, m, _ptr(xp), _ptr(vp), _ptr(rc), _ptr(vf), n)
            if "pid" in s:   # -DPID: integer IDs that ride with vp (buffer_density/buffer_v/update_particle `#ifdef PID` lines)
                pid = np.ascontiguousarray(s["pid"], np.int64)
                assert pid.shape == (n,)
                lib().oracle_load_pid(self.h, m, _ptr(pid))
        self.has_pid = all("pid" in s for s in states)
        assert self.has_pid or not any("pid" in s for s in states), "give every image IDs or none"
        lib().oracle_finish_load(self.h, F32(sigma_vi))

    def store(self, m=0):
        n = self.nplocal(m)
        shp = (self.nnt,) * 3 + (self.nt,) * 3
        rc = np.empty(shp, np.int32); vf = np.empty(shp + (3,), np.float32)
        lib().oracle_store_image(self.h, m, _ptr(rc), _ptr(vf))
        xp, vp = self.xp(m)[:n], self.vp(m)[:n]
        if self.izipx == 1:
            assert xp.min(initial=0) >= -128 and xp.max(initial=0) <= 127
        if self.izipv == 1:
            assert vp.min(initial=0) >= -128 and vp.max(initial=0) <= 127
        out = dict(xp=xp.astype((np.int8, np.int16)[self.izipx - 1]), vp=vp.astype((np.int8, np.int16)[self.izipv - 1]),
                   rhoc=rc, vfield=vf)
        if self.has_pid:
            out["pid"] = self.pid(m)[:n].copy()
        return out

    # ---- step subroutines ---------------------------------------------------------------------
    def buffer_density(self):
        lib().oracle_buffer_density(self.h); self._check()
        return F32(lib().oracle_overhead_image(self.h))

    def buffer_x(self):
        lib().oracle_buffer_x(self.h)

    def buffer_v(self):
        lib().oracle_buffer_v(self.h)

    def update_particle(self, dt_old, dt, vz_max=None):
        """``vz_max=None``: CUBE/main (source cells in storage order).  With ``vz_max`` (CUBEnu's ``sim%vz_max`` of the
        previous ``particle_mesh``, pm.f90:398) the k planes are visited in ``nlayer = 2*ceiling(dt_mid*vz_max/ncell)+1``
        colour passes as CUBEnu's ``update_xp`` does (update_particle.f90:37,55-58,97-103): same arithmetic, another
        order of the particles inside a destination cell and of the f32 sums into ``vfield_new``."""
        L = lib()
        self.nlayer = 1 if vz_max is None else int(L.oracle_nlayer_for(self.h, F32(dt_old), F32(dt), F32(vz_max)))
        L.oracle_set_nlayer(self.h, self.nlayer)
        L.oracle_update_particle(self.h, F32(dt_old), F32(dt)); self._check()
        return dict(sigma_vi_new=self.sigma_vi_new,
                    std_vsim=lib().oracle_std_vsim(self.h, 0), std_vsim_c=lib().oracle_std_vsim(self.h, 1),
                    std_vsim_res=lib().oracle_std_vsim(self.h, 2),
                    overhead_tile=F32(lib().oracle_overhead_tile(self.h)))

    # ---- particle_mesh (pm.f90) -----------------------------------------------------------------
    def fine_density(self, m, tx, ty, tz):
        """rho_f(nfe+2,nfe,nfe) of one tile as numpy [z][y][x+2] (1-based tile indices)."""
        rho = np.empty((self.nfe, self.nfe, self.nfe + 2), F32)
        lib().oracle_fine_deposit(self.h, m, tx, ty, tz, _ptr(rho))
        return rho

    def fine_force(self, rho):
        """pm.f90:75-84 -> force_f(3,nfb:nfe-nfb+1,...) as numpy [z][y][x][3]."""
        nfe, nfb = self.nfe, self.nfb
        c = rfftn(rho[:, :, :nfe])
        s = slice(nfb - 1, nfe - nfb + 1)
        ff = np.empty((self.nft + 2,) * 3 + (3,), F32)
        for d in range(3):
            k = self.kern_f[d]
            out = np.empty_like(c)
            out.real = -c.imag * k   # rho_f(::2)=-crho_f(2::2)*kern_f
            out.imag = c.real * k    # rho_f(2::2)=crho_f(::2)*kern_f
            r = irfftn_unnorm(out, (nfe,) * 3).astype(F32, copy=False)
            r = r / F32(nfe) / F32(nfe) / F32(nfe)
            ff[..., d] = r[s, s, s]
        return ff

    def coarse_density(self):
        """Global r3 assembled from all images, numpy [z][y][x]."""
        nc = self.nc
        g = np.empty((self.ncg[2], self.ncg[1], self.ncg[0]), F32)
        r3 = np.empty((nc, nc, nc), F32)
        for m in range(self.nimg):
            lib().oracle_coarse_deposit(self.h, m, _ptr(r3))
            ix, iy, iz = self.image_coords(m)
            g[iz * nc:(iz + 1) * nc, iy * nc:(iy + 1) * nc, ix * nc:(ix + 1) * nc] = r3
        return g

    def coarse_force(self, r3g):
        """pm.f90:168-178 on the global grid -> [z][y][x][3] (no halo)."""
        c = rfftn(r3g)
        out_f = np.empty(r3g.shape + (3,), F32)
        for d in range(3):
            k = self.kern_c[d]
            out = np.empty_like(c)
            out.real = -c.imag * k
            out.imag = c.real * k
            r = irfftn_unnorm(out, r3g.shape).astype(F32, copy=False)
            # r3=r3/ng_global/ng_global/ng_global  (pencil_fft.f90:58), per-dim for non-cubic grids
            r = r / F32(self.ncg[0]) / F32(self.ncg[1]) / F32(self.ncg[2])
            out_f[..., d] = r
        return out_f

    def image_coords(self, m):
        nx, ny, _ = self.nn
        return m % nx, (m // nx) % ny, m // (nx * ny)

    def force_c_image(self, fcg, m):
        """force_c(3,0:nc+1,...) of image m incl. the 1-cell halo (pm.f90:182-189)."""
        nc = self.nc
        ix, iy, iz = self.image_coords(m)
        zi = np.arange(iz * nc - 1, (iz + 1) * nc + 1) % self.ncg[2]
        yi = np.arange(iy * nc - 1, (iy + 1) * nc + 1) % self.ncg[1]
        xi = np.arange(ix * nc - 1, (ix + 1) * nc + 1) % self.ncg[0]
        return np.ascontiguousarray(fcg[np.ix_(zi, yi, xi)])

    def particle_mesh(self, a_mid, dt, keep=False):
        """One call of ``particle_mesh`` (pm.f90:1-247).  Returns dt limits; with ``keep`` also the
        meshes (for parity tests)."""
        L = lib()
        a_mid, dt = F32(a_mid), F32(dt)
        L.oracle_pm_begin(self.h)
        kept = dict(rho_f={}, force_f={}) if keep else None
        for m in range(self.nimg):
            for tz in range(1, self.nnt + 1):
                for ty in range(1, self.nnt + 1):
                    for tx in range(1, self.nnt + 1):
                        rho = self.fine_density(m, tx, ty, tz)
                        ff = self.fine_force(rho)
                        L.oracle_fine_kick(self.h, m, tx, ty, tz, _ptr(ff), a_mid, dt)
                        if keep:
                            kept["rho_f"][(m, tx, ty, tz)] = rho
                            kept["force_f"][(m, tx, ty, tz)] = ff
        L.oracle_pm_fine_end(self.h)
        r3g = self.coarse_density()
        fcg = self.coarse_force(r3g)
        for m in range(self.nimg):
            fc = self.force_c_image(fcg, m)
            L.oracle_coarse_kick(self.h, m, _ptr(fc), a_mid, dt)
        # pm.f90:233-244 (f32)
        GG = F32(1.0) / F32(6.0) / PI_F
        t = self.nnt ** 3
        f2f = max(float(np.ctypeslib.as_array((C.c_float * t).from_address(L.oracle_f2_max_fine(self.h, m))).max())
                  for m in range(self.nimg))
        dtf, dtc, dtv = [], [], []
        for m in range(self.nimg):
            f2m = np.ctypeslib.as_array((C.c_float * t).from_address(L.oracle_f2_max_fine(self.h, m))).max()
            dtf.append(np.sqrt(F32(1.0) / (np.sqrt(F32(f2m)) * a_mid * GG)))
            dtc.append(np.sqrt(F32(self.ncell) / (np.sqrt(F32(L.oracle_f2_max_coarse(self.h, m))) * a_mid * GG)))
            with np.errstate(divide="ignore"):
                dtv.append(F32(0.9) * F32(20) / F32(L.oracle_vmax(self.h, m)))
        self.dt_fine, self.dt_coarse, self.dt_vmax = F32(min(dtf)), F32(min(dtc)), F32(min(dtv))
        out = dict(dt_fine=self.dt_fine, dt_coarse=self.dt_coarse, dt_vmax=self.dt_vmax, dt_pp=F32(1000),
                   vmax=[F32(L.oracle_vmax(self.h, m)) for m in range(self.nimg)],
                   vmax3=[[F32(L.oracle_vmax3(self.h, m, d)) for d in range(3)] for m in range(self.nimg)],  # CUBEnu pm.f90:349,398
                   f2_max_fine=F32(f2f),
                   f2_max_coarse=[F32(L.oracle_f2_max_coarse(self.h, m)) for m in range(self.nimg)])
        if keep:
            kept["r3"] = r3g; kept["force_c"] = fcg
            out["meshes"] = kept
        return out

    def set_mass_p(self, mass_p):
        lib().oracle_set_mass_p(self.h, F32(mass_p))

    # ---- whole step, cafcube.f90:26-31 --------------------------------------------------------
    def step(self, dt_old, dt, a_mid, keep=False):
        up = self.update_particle(dt_old, dt)
        self.buffer_density(); self.buffer_x()
        pm = self.particle_mesh(a_mid, dt, keep=keep)
        self.buffer_v()
        return up, pm


def particle_mesh_two_species(A, B, a_mid, dt):
    """``particle_mesh`` of a two-species run as CUBEnu's pm.f90 reads with NEUTRINOS (pm.f90:79-99: both species are deposited
    into the same rho_f; :130-163: and into the same r3; :196-228 and its neutrino twin: each species is kicked from the common
    force with its own sigma_vi), composed from the one-species restatement: ``A`` and ``B`` are two :class:`Oracle` objects on
    the same geometry holding one species each (own ``mass_p``, own ``sigma_vi``).  The reference adds the second species'
    terms into the array that already holds the first one's; here the two deposits are summed afterwards (f32 round-off of a
    different grouping, far below the 1e-6 density gate).  That build does not compile upstream, so this is a
    reading of the code, unpinned.  Returns the dt limits over both species."""
    L = lib()
    a_mid, dt = F32(a_mid), F32(dt)
    assert (A.nn, A.nnt, A.nc) == (B.nn, B.nnt, B.nc)
    for O in (A, B):
        L.oracle_pm_begin(O.h)
    for m in range(A.nimg):
        for tz in range(1, A.nnt + 1):
            for ty in range(1, A.nnt + 1):
                for tx in range(1, A.nnt + 1):
                    rho = A.fine_density(m, tx, ty, tz) + B.fine_density(m, tx, ty, tz)
                    ff = A.fine_force(rho)
                    for O in (A, B):
                        L.oracle_fine_kick(O.h, m, tx, ty, tz, _ptr(ff), a_mid, dt)
    for O in (A, B):
        L.oracle_pm_fine_end(O.h)
    fcg = A.coarse_force(A.coarse_density() + B.coarse_density())
    for m in range(A.nimg):
        fc = A.force_c_image(fcg, m)
        for O in (A, B):
            L.oracle_coarse_kick(O.h, m, _ptr(fc), a_mid, dt)
    GG = F32(1.0) / F32(6.0) / PI_F
    t = A.nnt ** 3
    f2f = max(float(np.ctypeslib.as_array((C.c_float * t).from_address(L.oracle_f2_max_fine(A.h, m))).max()) for m in range(A.nimg))
    f2c = max(float(L.oracle_f2_max_coarse(A.h, m)) for m in range(A.nimg))
    vmax = [max(float(L.oracle_vmax(O.h, m)) for m in range(O.nimg)) for O in (A, B)]
    return dict(dt_fine=np.sqrt(F32(1.0) / (np.sqrt(F32(f2f)) * a_mid * GG)), dt_coarse=np.sqrt(F32(A.ncell) / (np.sqrt(F32(f2c)) * a_mid * GG)),
                dt_vmax=F32(0.9) * F32(20) / F32(max(vmax)), vmax=F32(vmax[0]), vmax2=F32(vmax[1]))
