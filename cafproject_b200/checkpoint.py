"""CUBE checkpoint files (CUBE/main/checkpoint.f90:33-70, particle_initialization.f90:11-64).

Little-endian stream binary, no record markers, one set per image under ``<opath>/image<N>/``:

=========  ==============================================  =======================
``zip2``   168-byte ``sim_header`` + ``rhoc(nt,nt,nt,nnt,nnt,nnt)`` int32   parameters.f90:119-140
``zip0``   ``xp int16(3,nplocal)``
``zip1``   ``vp int16(3,nplocal)``
``vfield`` ``f32(3,nt,nt,nt,nnt,nnt,nnt)``
=========  ==============================================  =======================

File names follow parameters.f90:221-257: ``<z as f7.3, trimmed><name>_<image>.bin``.

The CUBEnu variant of the same state (CUBEnu/work/main/checkpoint.f90:10-50, parameters.f90:136-160,
basic_functions.fh:36-87) is written with ``convention="cubenu"``: a 224-byte header alone in ``info``, the counts in
``np``, and the names ``<z>_<name>_<image>.bin`` with ``xp, vp, np, vc`` (``id`` for particle IDs).
"""
from __future__ import annotations

import os

import numpy as np

#: sim_header, parameters.f90:119-140: 13 x int64 then 16 x float32, no padding (168 bytes)
HEADER_DTYPE = np.dtype(
    [(n, "<i8") for n in ("nplocal", "izipx", "izipv", "image", "nn", "nnt", "nt", "ncell", "ncb", "istep",
                          "cur_checkpoint", "cur_proj", "cur_halo")]
    + [(n, "<f4") for n in ("a", "t", "tau", "dt_f_acc", "dt_pp_acc", "dt_c_acc", "mass_p", "box", "h0",
                            "omega_m", "omega_l", "s8", "vsim2phys", "sigma_vres", "sigma_vi", "z_i")]
)
assert HEADER_DTYPE.itemsize == 168


#: CUBEnu sim_header, CUBEnu/work/main/parameters.f90:136-160: 17 x int64 then 22 x float32 (224 bytes)
HEADER_NU_DTYPE = np.dtype(
    [(n, "<i8") for n in ("nplocal", "npglobal", "nplocal_nu", "npglobal_nu", "izipx", "izipv", "izipx_nu", "izipv_nu", "image",
                          "nn", "nnt", "nt", "ncell", "ncb", "timestep", "cur_checkpoint", "cur_halofind")]
    + [(n, "<f4") for n in ("a", "t", "tau", "dt_pp", "dt_fine", "dt_coarse", "dt_vmax", "dt_vmax_nu", "mass_p_cdm", "mass_p_nu",
                            "box", "h0", "omega_m", "omega_l", "s8", "vsim2phys", "sigma_vres", "sigma_vi", "sigma_vi_nu", "z_i",
                            "z_i_nu", "vz_max")]
)
assert HEADER_NU_DTYPE.itemsize == 224
#: CUBE name -> CUBEnu name of the same content
NU_NAMES = {"zip2": "info", "zip0": "xp", "zip1": "vp", "rhoc": "np", "vfield": "vc", "zipid": "id"}


PID_DTYPE = {"cube": np.dtype("<i8"), "cubenu": np.dtype("<i4")}


def z2str(z: float) -> str:
    """parameters.f90:213-219: write(str,'(f7.3)') z ; trim(adjustl(str))."""
    return ("%7.3f" % z).strip()


def file_name(opath: str, z: float, image: int, zipname: str, convention: str = "cube") -> str:
    """parameters.f90:244-257 (image is 1-based); CUBEnu: basic_functions.fh:52-87 with the CUBEnu name of ``zipname``."""
    if convention == "cubenu":
        return os.path.join(opath, "image%d" % image, "%s_%s_%d.bin" % (z2str(z), NU_NAMES.get(zipname, zipname), image))
    return os.path.join(opath, "image%d" % image, "%s%s_%d.bin" % (z2str(z), zipname, image))


def make_header(convention: str = "cube", **kw) -> np.ndarray:
    h = np.zeros((), HEADER_NU_DTYPE if convention == "cubenu" else HEADER_DTYPE)
    for k, v in kw.items():
        h[k] = v
    return h


def write_checkpoint(opath: str, z: float, image: int, header: np.ndarray, state: dict, convention: str = "cube") -> None:
    """``state``: xp (n,3) int8|int16, vp (n,3) int8|int16 (the header's izipx/izipv bytes per code, variables.f90:41-42),
    rhoc [tz][ty][tx][k][j][i] i32, vfield [...][3] f32."""
    os.makedirs(os.path.join(opath, "image%d" % image), exist_ok=True)
    header = header.copy()
    header["nplocal"] = state["xp"].shape[0]
    xdt, vdt = _code_dtypes(header)
    if state["xp"].dtype != xdt or state["vp"].dtype != vdt:
        raise ValueError("zip format incompatable: header says izipx=%d izipv=%d, state holds %s/%s"
                         % (int(header["izipx"]), int(header["izipv"]), state["xp"].dtype, state["vp"].dtype))
    if convention == "cubenu":
        if header.dtype != HEADER_NU_DTYPE:
            raise ValueError("convention='cubenu' needs a HEADER_NU_DTYPE header (make_header('cubenu', ...))")
        with open(file_name(opath, z, image, "zip2", convention), "wb") as f:
            f.write(header.tobytes())
        np.ascontiguousarray(state["rhoc"], "<i4").tofile(file_name(opath, z, image, "rhoc", convention))
    else:
        with open(file_name(opath, z, image, "zip2"), "wb") as f:
            f.write(header.tobytes())
            f.write(np.ascontiguousarray(state["rhoc"], "<i4").tobytes())
    np.ascontiguousarray(state["vfield"], "<f4").tofile(file_name(opath, z, image, "vfield", convention))
    np.ascontiguousarray(state["xp"], xdt).tofile(file_name(opath, z, image, "zip0", convention))
    np.ascontiguousarray(state["vp"], vdt).tofile(file_name(opath, z, image, "zip1", convention))
    if "pid" in state:   # -DPID: CUBE/main `zipid`, integer(8) (variables.f90:44); CUBEnu `id`, integer(4) (variables.f90:47, checkpoint.f90:47)
        np.ascontiguousarray(state["pid"], PID_DTYPE[convention]).tofile(file_name(opath, z, image, "zipid", convention))


def _code_dtypes(header):
    """numpy dtypes of xp and vp for a header's izipx, izipv (1 or 2 bytes; CUBE/main/universe*.fh:2-3)."""
    zx, zv = int(header["izipx"]), int(header["izipv"])
    if zx not in (1, 2) or zv not in (1, 2):
        raise ValueError("zip format incompatable: izipx=%d izipv=%d" % (zx, zv))
    return np.dtype("<i%d" % zx), np.dtype("<i%d" % zv)


def _with_pid(state, opath, z, image, convention):
    """Adds ``pid`` when the checkpoint has an ID file (runs built with -DPID; particle_initialization.f90:56)."""
    fn = file_name(opath, z, image, "zipid", convention)
    if os.path.exists(fn):
        state["pid"] = np.fromfile(fn, PID_DTYPE[convention])
        if state["pid"].shape[0] != state["xp"].shape[0]:
            raise ValueError("ID file holds %d IDs for %d particles" % (state["pid"].shape[0], state["xp"].shape[0]))
    return state


def _check_zip(header, expect):
    """particle_initialization.f90:14-18: a run built for (izipx, izipv) stops on a file written in another format."""
    if expect is not None and (int(header["izipx"]), int(header["izipv"])) != tuple(expect):
        raise ValueError("zip format incompatable")


def read_checkpoint(opath: str, z: float, image: int, convention: str = "cube", expect_zip=None):
    """Returns ``(header, state)``; ``xp``/``vp`` come back in the file's own 1- or 2-byte format.  ``expect_zip=(izipx,
    izipv)`` makes a mismatch the reference's "zip format incompatable" stop (the GPU step is built for (2, 2))."""
    if convention == "cubenu":
        header = np.fromfile(file_name(opath, z, image, "zip2", convention), HEADER_NU_DTYPE)[0]
        nnt, nt = int(header["nnt"]), int(header["nt"])
        rhoc = np.fromfile(file_name(opath, z, image, "rhoc", convention), "<i4").reshape((nnt,) * 3 + (nt,) * 3)
        _check_zip(header, expect_zip)
        xdt, vdt = _code_dtypes(header)
        n = int(header["nplocal"])
        vfield = np.fromfile(file_name(opath, z, image, "vfield", convention), "<f4").reshape(rhoc.shape + (3,))
        xp = np.fromfile(file_name(opath, z, image, "zip0", convention), xdt).reshape(n, 3)
        vp = np.fromfile(file_name(opath, z, image, "zip1", convention), vdt).reshape(n, 3)
        return header, _with_pid(dict(xp=xp, vp=vp, rhoc=rhoc, vfield=vfield), opath, z, image, convention)
    with open(file_name(opath, z, image, "zip2"), "rb") as f:
        header = np.frombuffer(f.read(HEADER_DTYPE.itemsize), HEADER_DTYPE)[0]
        nnt, nt = int(header["nnt"]), int(header["nt"])
        rhoc = np.frombuffer(f.read(), "<i4").reshape((nnt,) * 3 + (nt,) * 3).copy()
    _check_zip(header, expect_zip)
    xdt, vdt = _code_dtypes(header)
    n = int(header["nplocal"])
    vfield = np.fromfile(file_name(opath, z, image, "vfield"), "<f4").reshape(rhoc.shape + (3,))
    xp = np.fromfile(file_name(opath, z, image, "zip0"), xdt).reshape(n, 3)
    vp = np.fromfile(file_name(opath, z, image, "zip1"), vdt).reshape(n, 3)
    return header, _with_pid(dict(xp=xp, vp=vp, rhoc=rhoc, vfield=vfield), opath, z, image, convention)
