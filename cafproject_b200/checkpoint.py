"""CUBE checkpoint files (CUBE/main/checkpoint.f90:33-70, particle_initialization.f90:11-64).

Little-endian stream binary, no record markers, one set per image under ``<opath>/image<N>/``:

=========  ==============================================  =======================
``zip2``   168-byte ``sim_header`` + ``rhoc(nt,nt,nt,nnt,nnt,nnt)`` int32   parameters.f90:119-140
``zip0``   ``xp int16(3,nplocal)``
``zip1``   ``vp int16(3,nplocal)``
``vfield`` ``f32(3,nt,nt,nt,nnt,nnt,nnt)``
=========  ==============================================  =======================

File names follow parameters.f90:221-257: ``<z as f7.3, trimmed><name>_<image>.bin``.
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


def z2str(z: float) -> str:
    """parameters.f90:213-219: write(str,'(f7.3)') z ; trim(adjustl(str))."""
    return ("%7.3f" % z).strip()


def file_name(opath: str, z: float, image: int, zipname: str) -> str:
    """parameters.f90:244-257 (image is 1-based)."""
    return os.path.join(opath, "image%d" % image, "%s%s_%d.bin" % (z2str(z), zipname, image))


def make_header(**kw) -> np.ndarray:
    h = np.zeros((), HEADER_DTYPE)
    for k, v in kw.items():
        h[k] = v
    return h


def write_checkpoint(opath: str, z: float, image: int, header: np.ndarray, state: dict) -> None:
    """``state``: xp (n,3) i16, vp (n,3) i16, rhoc [tz][ty][tx][k][j][i] i32, vfield [...][3] f32."""
    os.makedirs(os.path.join(opath, "image%d" % image), exist_ok=True)
    header = header.copy()
    header["nplocal"] = state["xp"].shape[0]
    with open(file_name(opath, z, image, "zip2"), "wb") as f:
        f.write(header.tobytes())
        f.write(np.ascontiguousarray(state["rhoc"], "<i4").tobytes())
    np.ascontiguousarray(state["vfield"], "<f4").tofile(file_name(opath, z, image, "vfield"))
    np.ascontiguousarray(state["xp"], "<i2").tofile(file_name(opath, z, image, "zip0"))
    np.ascontiguousarray(state["vp"], "<i2").tofile(file_name(opath, z, image, "zip1"))


def read_checkpoint(opath: str, z: float, image: int):
    with open(file_name(opath, z, image, "zip2"), "rb") as f:
        header = np.frombuffer(f.read(HEADER_DTYPE.itemsize), HEADER_DTYPE)[0]
        nnt, nt = int(header["nnt"]), int(header["nt"])
        rhoc = np.frombuffer(f.read(), "<i4").reshape((nnt,) * 3 + (nt,) * 3).copy()
    if int(header["izipx"]) != 2 or int(header["izipv"]) != 2:
        raise ValueError("zip format incompatable")  # particle_initialization.f90:14-18
    n = int(header["nplocal"])
    vfield = np.fromfile(file_name(opath, z, image, "vfield"), "<f4").reshape(rhoc.shape + (3,))
    xp = np.fromfile(file_name(opath, z, image, "zip0"), "<i2").reshape(n, 3)
    vp = np.fromfile(file_name(opath, z, image, "zip1"), "<i2").reshape(n, 3)
    return header, dict(xp=xp, vp=vp, rhoc=rhoc, vfield=vfield)
