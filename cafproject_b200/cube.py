"""Host-side mirror of CUBE's step loop on top of the C ABI (``include/cube_gpu.h``).

The reference's "operator interface" for the hot path is the sequence of argument-less subroutine
calls in ``CUBE/main/cafcube.f90:25-46`` working on module globals.  :class:`CubeGPU` offers the same
names (``update_particle``, ``buffer_density``, ``buffer_x``, ``buffer_v``, ``particle_mesh``,
``checkpoint``) over ``libcubegpu.so``; the state lives in HBM between calls.  There is no CPU
fallback: importing works without a GPU (so that the symbol table can be checked), every compute
call raises if the CUDA library or a device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcubegpu.so")

F32 = np.float32
PI_F = F32(4) * np.arctan(F32(1.0), dtype=F32)


class CubeParams(C.Structure):
    """``cube_params`` of include/cube_gpu.h."""
    _fields_ = [("nn", C.c_int32 * 3), ("rank", C.c_int32), ("nnt", C.c_int32), ("nc", C.c_int32),
                ("ncell", C.c_int32), ("ncb", C.c_int32), ("izipx", C.c_int32), ("izipv", C.c_int32),
                ("np_nc", C.c_int32), ("image_buffer", C.c_float), ("tile_buffer", C.c_float),
                ("device", C.c_int32), ("fine_batch", C.c_int32), ("local_group", C.c_int32), ("reserved", C.c_int32 * 3)]


class CubeGPUError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen libcubegpu.so and declare every prototype of include/cube_gpu.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CubeGPUError("libcubegpu.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, f32, i32 = C.c_void_p, C.c_int64, C.c_float, C.c_int
    L.cube_gpu_init.argtypes = [C.POINTER(CubeParams), vp, vp, vp, vp, C.POINTER(vp)]
    L.cube_gpu_upload.argtypes = [vp, vp, vp, vp, vp, i64, i64, f32]
    L.cube_gpu_upload_begin.argtypes = [vp, vp, vp, vp, vp, i64, i64, f32]
    L.cube_gpu_update_x.argtypes = [vp, f32, f32, C.POINTER(i64), C.POINTER(f32), C.POINTER(C.c_double * 3), C.POINTER(f32)]
    L.cube_gpu_buffer.argtypes = [vp, i32, i32, i32, C.POINTER(f32)]
    L.cube_gpu_particle_mesh.argtypes = [vp, f32, f32] + [C.POINTER(f32)] * 4
    L.cube_gpu_download.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(f32)]
    L.cube_gpu_finalize.argtypes = [vp]
    L.cube_gpu_download_async.argtypes = [vp, vp, vp]
    L.cube_gpu_download_cells_async.argtypes = [vp, vp, vp]
    L.cube_gpu_stream_vp.argtypes = [vp, vp]
    L.cube_gpu_upload_pid.argtypes = [vp, vp]
    L.cube_gpu_download_pid.argtypes = [vp, vp]
    L.cube_gpu_last_error.restype = C.c_char_p
    L.cube_gpu_query.restype = i64
    L.cube_gpu_query.argtypes = [vp, C.c_char_p]
    L.cube_gpu_get_kern_f.argtypes = [vp, vp]
    L.cube_gpu_get_kern_c.argtypes = [vp, vp]
    L.cube_gpu_fine_density.argtypes = [vp, i32, i32, i32, vp]
    L.cube_gpu_fine_force.argtypes = [vp, i32, i32, i32, vp]
    L.cube_gpu_fine_kick_with.argtypes = [vp, i32, i32, i32, vp, f32, f32, f32, f32, C.POINTER(f32)]
    L.cube_gpu_coarse_density.argtypes = [vp, vp]
    L.cube_gpu_coarse_force.argtypes = [vp, vp]
    L.cube_gpu_coarse_kick_with.argtypes = [vp, vp, f32, f32, f32, C.POINTER(f32), C.POINTER(f32)]
    L.cube_gpu_phase_count.restype = i32
    L.cube_gpu_phase_name.restype = C.c_char_p
    L.cube_gpu_phase_name.argtypes = [i32]
    L.cube_gpu_phase_times.argtypes = [vp, vp]
    L.cube_gpu_set_profiling.argtypes = [vp, i32]
    L.cube_gpu_timer.argtypes = [vp, i32, C.POINTER(f32)]
    L.cube_gpu_nccl_unique_id.argtypes = [vp]
    L.cube_gpu_exchange_plan.argtypes = [C.POINTER(CubeParams), vp, i32]
    L.cube_gpu_selftest_codes.argtypes = [vp, f32, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.cube_gpu_power_spectrum.argtypes = [vp, f32, vp, i32, C.POINTER(i32)]
    L.cube_gpu_set_mass_p.argtypes = [vp, f32]
    L.cube_gpu_particle_mesh_species.argtypes = [vp, vp, f32, f32] + [C.POINTER(f32)] * 6
    L.cube_gpu_set_drift_layers.argtypes = [vp, i32]
    L.cube_gpu_get_vmax3.argtypes = [vp, C.POINTER(C.c_float * 3)]
    _lib = L
    return L


#: every symbol include/cube_gpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "cube_gpu_init", "cube_gpu_upload", "cube_gpu_update_x", "cube_gpu_buffer", "cube_gpu_particle_mesh",
    "cube_gpu_download", "cube_gpu_finalize", "cube_gpu_last_error", "cube_gpu_query", "cube_gpu_get_kern_f",
    "cube_gpu_get_kern_c", "cube_gpu_fine_density", "cube_gpu_fine_force", "cube_gpu_fine_kick_with",
    "cube_gpu_coarse_density", "cube_gpu_coarse_force", "cube_gpu_coarse_kick_with", "cube_gpu_phase_count",
    "cube_gpu_phase_name", "cube_gpu_phase_times", "cube_gpu_set_profiling", "cube_gpu_timer", "cube_gpu_nccl_unique_id",
    "cube_gpu_exchange_plan", "cube_gpu_download_async", "cube_gpu_download_cells_async", "cube_gpu_stream_vp", "cube_gpu_selftest_codes",
    "cube_gpu_upload_pid", "cube_gpu_download_pid", "cube_gpu_power_spectrum", "cube_gpu_set_drift_layers", "cube_gpu_get_vmax3", "cube_gpu_set_mass_p", "cube_gpu_particle_mesh_species",
    "cube_gpu_upload_begin",
]


def exchange_plan(nn, rank, nc, nnt):
    """Message plan of image ``rank`` on the image grid ``nn`` (host only; see include/cube_gpu.h)."""
    L = load_library()
    p = CubeParams()
    p.nn[:] = tuple(int(v) for v in nn)
    p.rank, p.nnt, p.nc, p.ncell, p.ncb, p.izipx, p.izipv = rank, nnt, nc, 4, 6, 2, 2
    out = np.zeros((128, 8), np.int64)
    n = L.cube_gpu_exchange_plan(C.byref(p), out.ctypes.data, 128)
    if n < 0:
        raise CubeGPUError("bad parameters")
    rows = out[:n]
    return dict(ghost=[tuple(int(v) for v in r[1:]) for r in rows if r[0] == 0],
                force_send=[(int(r[1]), int(r[2])) for r in rows if r[0] == 1],
                force_recv=[(int(r[1]), int(r[2])) for r in rows if r[0] == 2],
                coarse=[tuple(int(v) for v in r[1:]) for r in rows if r[0] == 3][0])


def nccl_unique_id() -> bytes:
    """128-byte ncclUniqueId made by the library's own NCCL (image 1 calls this and broadcasts it)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.cube_gpu_nccl_unique_id(buf) != 0:
        raise CubeGPUError(L.cube_gpu_last_error().decode())
    return buf.raw


def image_grid(n: int):
    """Image grid (nnx,nny,nnz) used for n = 1, 2, 4, 8 GPUs: 1x1x1, 2x1x1, 2x2x1, 2x2x2 (the reference only has nn^3)."""
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    if n not in grids:
        raise ValueError("supported image counts: 1, 2, 4, 8")
    return grids[n]


def host_tanf_lut(izipv: int = 2) -> np.ndarray:
    """``tan((pi*real(v))/real(nvbin-1))`` for all ``nvbin = 2**(8*izipv)`` codes (indexed by the raw pattern) with the
    *host's* libm ``tanf`` (what the Fortran host passes to ``cube_gpu_init``).  Uses libm through ctypes, not the oracle."""
    libm = C.CDLL("libm.so.6")
    libm.tanf.restype = C.c_float
    libm.tanf.argtypes = [C.c_float]
    nvbin = 1 << (8 * izipv)
    if izipv == 2:
        codes = np.arange(nvbin, dtype=np.uint16).view(np.int16).astype(F32)
    else:
        codes = np.arange(nvbin, dtype=np.uint8).view(np.int8).astype(F32)
    arg = (PI_F * codes) / F32(nvbin - 1)
    return np.array([libm.tanf(float(a)) for a in arg], dtype=F32)


def code_dtypes(izipx: int, izipv: int):
    """numpy dtypes of xp and vp (integer(izipx), integer(izipv): CUBE/main/variables.f90:41-42)."""
    return (np.int8, np.int16)[izipx - 1], (np.int8, np.int16)[izipv - 1]


def _p(a):
    return a.ctypes.data if a is not None else None


class CubeGPU:
    """One image of a CUBE run on one B200.  Mirrors the step subroutines of CUBE/main."""

    def __init__(self, nc, nnt, fk_table, ck_table, nn=(1, 1, 1), rank=0, np_nc=2, image_buffer=1.5, tile_buffer=2.5,
                 device=0, fine_batch=0, tanf_lut=None, nccl_id=None, local_group=0, izipx=2, izipv=2, secondary=False):
        """``nn``: image grid; ``rank``: this image (0-based, x fastest).  More than one image needs either ``nccl_id``
        (one process per GPU; the 128 bytes of :func:`nccl_unique_id` broadcast from image 1) or ``local_group`` > 0
        (every image is a host thread of this process, e.g. several images per GPU)."""
        L = load_library()
        self.L = L
        nn = (int(nn),) * 3 if np.isscalar(nn) else tuple(int(v) for v in nn)
        p = CubeParams()
        p.nn[:] = nn
        p.rank, p.nnt, p.nc, p.ncell, p.ncb = rank, nnt, nc, 4, 6
        p.izipx, p.izipv = int(izipx), int(izipv)    # the run's zip format (CUBE/main/universe*.fh:2-3)
        self.izipx, self.izipv = int(izipx), int(izipv)
        self.xdt, self.vdt = code_dtypes(self.izipx, self.izipv) if izipx in (1, 2) and izipv in (1, 2) else (np.int16, np.int16)
        p.np_nc, p.image_buffer, p.tile_buffer, p.device, p.fine_batch = np_nc, image_buffer, tile_buffer, device, fine_batch
        p.local_group = int(local_group)
        p.reserved[0] = 1 if secondary else 0     # a further species: kicked from another handle's meshes (particle_mesh_species)
        self.params = p
        self.nn, self.nc, self.nnt, self.nt = nn, nc, nnt, nc // nnt
        self.nft = 4 * self.nt
        self.nfe = self.nft + 48
        # reference layouts: fk_table(16,16,16,3) [dim slowest], ck_table(3,4,4,4) [dim fastest]
        fk = np.ascontiguousarray(np.moveaxis(np.asarray(fk_table, F32), 3, 0))  # fixture is [k][j][i][dim]
        ck = np.ascontiguousarray(np.asarray(ck_table, F32))
        lut = np.ascontiguousarray(host_tanf_lut(p.izipv if p.izipv in (1, 2) else 2) if tanf_lut is None else tanf_lut, F32)
        assert p.izipv not in (1, 2) or lut.shape == (1 << (8 * p.izipv),)
        h = C.c_void_p()
        self.h = None
        idbuf = C.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
        self._ck(L.cube_gpu_init(C.byref(p), _p(fk), _p(ck), _p(lut), idbuf, C.byref(h)))
        self.h = h
        self.nplocal = 0
        self.sigma_vi = F32(0)
        self.dt_fine = self.dt_coarse = self.dt_vmax = self.dt_pp = F32(1000)

    def _ck(self, rc):
        if rc != 0:
            raise CubeGPUError(self.L.cube_gpu_last_error().decode())

    def close(self):
        if self.h is not None:
            self.L.cube_gpu_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, what):
        return int(self.L.cube_gpu_query(self.h, what.encode()))

    def selftest_codes(self, sigma_vi, nsweep=1 << 22):
        """(bad_encode, bad_decode, fma_division): table-driven code conversions vs their formulas (include/cube_gpu.h)."""
        be, bd, ok = C.c_int64(0), C.c_int64(0), C.c_int(0)
        self._ck(self.L.cube_gpu_selftest_codes(self.h, F32(sigma_vi), int(nsweep), C.byref(be), C.byref(bd), C.byref(ok)))
        return be.value, bd.value, ok.value

    # ---- particle_initialization / checkpoint ------------------------------------------------
    def particle_initialization(self, state, sigma_vi, npglobal=None, streamed=False):
        """``streamed``: return while xp and vp (page-locked arrays, left untouched until the next ``update_particle`` has returned)
        are still being copied; ``update_particle`` then keys them chunk by chunk as they land (cube_gpu_upload_begin)."""
        if np.asarray(state["xp"]).dtype != self.xdt or np.asarray(state["vp"]).dtype != self.vdt:
            # particle_initialization.f90:14-18: a run built for (izipx, izipv) stops on a state in another format
            raise RuntimeError("zip format incompatable: this run has izipx=%d izipv=%d, the state holds %s/%s"
                               % (self.izipx, self.izipv, np.asarray(state["xp"]).dtype, np.asarray(state["vp"]).dtype))
        xp = np.ascontiguousarray(state["xp"], self.xdt); vp = np.ascontiguousarray(state["vp"], self.vdt)
        rc = np.ascontiguousarray(state["rhoc"], np.int32); vf = np.ascontiguousarray(state["vfield"], np.float32)
        n = xp.shape[0]
        up = self.L.cube_gpu_upload_begin if streamed else self.L.cube_gpu_upload
        self._ck(up(self.h, _p(xp), _p(vp), _p(rc), _p(vf), n, npglobal or n, F32(sigma_vi)))
        self._upload_keep = (xp, vp) if streamed else None     # the arrays the copy engine is still reading
        self.nplocal = n
        self.sigma_vi = F32(sigma_vi)
        self.has_pid = "pid" in state
        if self.has_pid:   # -DPID: IDs ride with the particles through update_particle (single image; include/cube_gpu.h)
            pid = np.ascontiguousarray(state["pid"], np.int64)
            assert pid.shape == (n,)
            self._ck(self.L.cube_gpu_upload_pid(self.h, _p(pid)))

    def checkpoint(self, out=None, skip=()):
        """Disjoint state back on the host (what checkpoint.f90 writes).  ``out`` may hold preallocated
        (e.g. pinned) arrays xp, vp (capacity >= nplocal rows), rhoc, vfield; names in ``skip`` were already streamed by
        :meth:`checkpoint_begin` and are only waited for."""
        n = self.query("nplocal")
        shp = (self.nnt,) * 3 + (self.nt,) * 3
        if out is None:
            out = dict(xp=np.empty((n, 3), self.xdt), vp=np.empty((n, 3), self.vdt), rhoc=np.empty(shp, np.int32),
                       vfield=np.empty(shp + (3,), np.float32))
        assert out["xp"].shape[0] >= n and out["vp"].shape[0] >= n and out["xp"].dtype == self.xdt and out["vp"].dtype == self.vdt
        npl = C.c_int64(); sig = C.c_float()
        self._ck(self.L.cube_gpu_download(self.h, None if "xp" in skip else _p(out["xp"]), None if "vp" in skip else _p(out["vp"]),
                                          None if "rhoc" in skip else _p(out["rhoc"]), None if "vfield" in skip else _p(out["vfield"]),
                                          C.byref(npl), C.byref(sig)))
        if out["xp"].shape[0] != n:
            out = dict(out, xp=out["xp"][:n], vp=out["vp"][:n])
        if getattr(self, "has_pid", False):
            pid = np.empty(n, np.int64)
            self._ck(self.L.cube_gpu_download_pid(self.h, _p(pid)))
            out = dict(out, pid=pid)
        return out, F32(sig.value)

    def checkpoint_begin(self, out, xp=True, vp=False, cells=False, vp_during_pm=False):
        """Start streaming xp and/or vp (and, with ``cells``, rhoc and vfield) of the current disjoint state into ``out``
        (page-locked arrays; xp, vp with capacity >= nplocal rows) while later calls run; finish with
        ``checkpoint(out=out, skip=...)``.  ``vp_during_pm``: the next ``particle_mesh`` streams every tile batch's final
        velocities into ``out["vp"]`` as soon as that batch has had its kicks (cube_gpu_stream_vp)."""
        n = self.query("nplocal")
        assert out["xp"].shape[0] >= n and out["vp"].shape[0] >= n
        if xp or vp:
            self._ck(self.L.cube_gpu_download_async(self.h, _p(out["xp"]) if xp else None, _p(out["vp"]) if vp else None))
        if cells:
            self._ck(self.L.cube_gpu_download_cells_async(self.h, _p(out["rhoc"]), _p(out["vfield"])))
        if vp_during_pm:
            self._ck(self.L.cube_gpu_stream_vp(self.h, _p(out["vp"])))

    # ---- step subroutines -------------------------------------------------------------------
    def update_particle(self, dt_old, dt, vz_max=None):
        """``vz_max=None``: CUBE/main's in-cell order.  With CUBEnu's ``sim%vz_max`` the source planes are visited in
        ``nlayer = 2*ceiling(dt_mid*vz_max/ncell)+1`` colour passes (CUBEnu update_particle.f90:37,55-58)."""
        nlayer = 1
        if vz_max is not None:
            dt_mid = F32(F32(F32(dt_old) + F32(dt)) / F32(2))
            nlayer = 2 * int(np.ceil(F32(F32(dt_mid * F32(vz_max)) / F32(4)))) + 1
        self._ck(self.L.cube_gpu_set_drift_layers(self.h, int(nlayer)))
        npl = C.c_int64(); sig = C.c_float(); ovh = C.c_float(); st = (C.c_double * 3)()
        self._ck(self.L.cube_gpu_update_x(self.h, F32(dt_old), F32(dt), C.byref(npl), C.byref(sig), C.byref(st), C.byref(ovh)))
        self.nplocal = npl.value
        return dict(nplocal=npl.value, sigma_vi_new=F32(sig.value), std_vsim=st[0], std_vsim_c=st[1], std_vsim_res=st[2],
                    overhead_tile=F32(ovh.value))

    def buffer_density(self):
        ovh = C.c_float()
        self._ck(self.L.cube_gpu_buffer(self.h, 1, 0, 0, C.byref(ovh)))
        return F32(ovh.value)

    def buffer_x(self):
        self._ck(self.L.cube_gpu_buffer(self.h, 0, 1, 0, None))

    def buffer_v(self):
        self._ck(self.L.cube_gpu_buffer(self.h, 0, 0, 1, None))

    def particle_mesh(self, a_mid, dt):
        o = [C.c_float() for _ in range(4)]
        self._ck(self.L.cube_gpu_particle_mesh(self.h, F32(a_mid), F32(dt), *[C.byref(v) for v in o]))
        self.dt_fine, self.dt_coarse, self.dt_vmax = (F32(v.value) for v in o[:3])
        v3 = (C.c_float * 3)()
        self._ck(self.L.cube_gpu_get_vmax3(self.h, C.byref(v3)))
        return dict(dt_fine=self.dt_fine, dt_coarse=self.dt_coarse, dt_vmax=self.dt_vmax, dt_pp=F32(1000), vmax=F32(o[3].value),
                    vmax3=[F32(v3[d]) for d in range(3)])   # vmax3: CUBEnu's vmax(3) (pm.f90:349,398)

    def set_mass_p(self, mass_p):
        """Particle mass of this species (CUBEnu sim%mass_p_cdm / sim%mass_p_nu); call after particle_initialization."""
        self._ck(self.L.cube_gpu_set_mass_p(self.h, F32(mass_p)))

    def particle_mesh_species(self, other, a_mid, dt):
        """particle_mesh for two species (CUBEnu -DNEUTRINOS): ``other``'s particles are deposited into this handle's meshes after
        its own, both are kicked by the one force field.  Returns this species' limits plus ``dt_vmax2``/``vmax2`` of ``other``."""
        o = [C.c_float() for _ in range(6)]
        self._ck(self.L.cube_gpu_particle_mesh_species(self.h, other.h, F32(a_mid), F32(dt), *[C.byref(v) for v in o]))
        self.dt_fine, self.dt_coarse, self.dt_vmax = (F32(v.value) for v in o[:3])
        other.dt_fine, other.dt_coarse, other.dt_vmax = self.dt_fine, self.dt_coarse, F32(o[4].value)
        return dict(dt_fine=self.dt_fine, dt_coarse=self.dt_coarse, dt_vmax=self.dt_vmax, dt_pp=F32(1000), vmax=F32(o[3].value),
                    dt_vmax2=F32(o[4].value), vmax2=F32(o[5].value))

    def step(self, dt_old, dt, a_mid):
        """cafcube.f90:27-31."""
        up = self.update_particle(dt_old, dt)
        self.buffer_density(); self.buffer_x()
        pm = self.particle_mesh(a_mid, dt)
        self.buffer_v()
        return up, pm

    # ---- diagnostics -------------------------------------------------------------------------
    def kern_f(self):
        """Im(FFT(fine force kernel)) on the N = query("nfft") window grid (kernel_f.f90:32-41 with nfe -> N)."""
        n = self.query("nfft")
        out = np.empty((3, n, n, n // 2 + 1), F32)
        self._ck(self.L.cube_gpu_get_kern_f(self.h, _p(out)))
        return out

    def kern_c(self):
        out = np.empty((3, self.nc, self.nc, self.nc // 2 + 1), F32)
        self._ck(self.L.cube_gpu_get_kern_c(self.h, _p(out)))
        return out

    def fine_density(self, itx, ity, itz):
        out = np.empty((self.nfe, self.nfe, self.nfe + 2), F32)
        self._ck(self.L.cube_gpu_fine_density(self.h, itx, ity, itz, _p(out)))
        return out

    def fine_force(self, itx, ity, itz):
        m = self.nft + 2
        out = np.empty((m, m, m, 3), F32)
        self._ck(self.L.cube_gpu_fine_force(self.h, itx, ity, itz, _p(out)))
        return out

    def fine_kick_with(self, itx, ity, itz, force_f, a_mid, dt, sigma_vi, sigma_vi_new):
        f = np.ascontiguousarray(force_f, F32); f2 = C.c_float()
        self._ck(self.L.cube_gpu_fine_kick_with(self.h, itx, ity, itz, _p(f), F32(a_mid), F32(dt), F32(sigma_vi), F32(sigma_vi_new), C.byref(f2)))
        return F32(f2.value)

    def coarse_density(self):
        out = np.empty((self.nc,) * 3, F32)
        self._ck(self.L.cube_gpu_coarse_density(self.h, _p(out)))
        return out

    def coarse_force(self):
        m = self.nc + 2
        out = np.empty((m, m, m, 3), F32)
        self._ck(self.L.cube_gpu_coarse_force(self.h, _p(out)))
        return out

    def coarse_kick_with(self, force_c, a_mid, dt, sigma_vi):
        f = np.ascontiguousarray(force_c, F32); vm = C.c_float(); f2 = C.c_float()
        self._ck(self.L.cube_gpu_coarse_kick_with(self.h, _p(f), F32(a_mid), F32(dt), F32(sigma_vi), C.byref(vm), C.byref(f2)))
        return F32(vm.value), F32(f2.value)

    def power_spectrum(self, box=200.0):
        """``xi(10, nbin)`` of CUBE/utilities/powerspectrum.f90 (linear_kbin) for the density contrast of the resident state
        (cicpower.f90), computed on the device by the library's own kernels; needs the buffered state, single image."""
        nbin = int(round((4 * self.nc // 2) * np.sqrt(3.0)))
        xi = np.zeros((10, nbin), np.float64)
        nb = C.c_int(0)
        self._ck(self.L.cube_gpu_power_spectrum(self.h, F32(box), _p(xi), nbin, C.byref(nb)))
        assert nb.value == nbin
        return xi

    def timer_start(self):
        self._ck(self.L.cube_gpu_timer(self.h, 1, None))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.cube_gpu_timer(self.h, 0, C.byref(ms)))
        return float(ms.value)

    def set_profiling(self, on=True):
        self.L.cube_gpu_set_profiling(self.h, int(on))

    def phase_times(self):
        n = self.L.cube_gpu_phase_count()
        ms = np.zeros(n, F32)
        self.L.cube_gpu_phase_times(self.h, _p(ms))
        return {self.L.cube_gpu_phase_name(i).decode(): float(ms[i]) for i in range(n)}
