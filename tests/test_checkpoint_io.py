"""Checkpoint files keep the reference's byte layout (checkpoint.f90:33-70, parameters.f90:119-140,213-257)."""
import os

import numpy as np

from cafproject_b200 import checkpoint as ck


def test_header_is_168_bytes_in_reference_order():
    assert ck.HEADER_DTYPE.itemsize == 168
    names = list(ck.HEADER_DTYPE.names)
    assert names[:13] == ["nplocal", "izipx", "izipv", "image", "nn", "nnt", "nt", "ncell", "ncb", "istep", "cur_checkpoint",
                          "cur_proj", "cur_halo"]
    assert names[13:] == ["a", "t", "tau", "dt_f_acc", "dt_pp_acc", "dt_c_acc", "mass_p", "box", "h0", "omega_m", "omega_l",
                          "s8", "vsim2phys", "sigma_vres", "sigma_vi", "z_i"]
    assert ck.HEADER_DTYPE.fields["a"][1] == 13 * 8


def test_file_names():
    assert ck.z2str(49.0) == "49.000" and ck.z2str(0.0) == "0.000" and ck.z2str(100.0) == "100.000"
    assert ck.file_name("/o", 49.0, 3, "zip0") == "/o/image3/49.000zip0_3.bin"


def test_round_trip(tmp_path):
    from cafproject_b200.synthetic_ic import make_ic
    states, sig, info = make_ic(nn=1, nc=24, nnt=2, np_nc=1, seed=4)
    s = states[0]
    h = ck.make_header(izipx=2, izipv=2, image=1, nn=1, nnt=2, nt=12, ncell=4, ncb=6, a=0.02, sigma_vi=sig, z_i=49, mass_p=64.0)
    ck.write_checkpoint(str(tmp_path), 49.0, 1, h, s)
    n = s["xp"].shape[0]
    d = tmp_path / "image1"
    assert os.path.getsize(d / "49.000zip0_1.bin") == 6 * n
    assert os.path.getsize(d / "49.000zip1_1.bin") == 6 * n
    assert os.path.getsize(d / "49.000zip2_1.bin") == 168 + 4 * 24 ** 3
    assert os.path.getsize(d / "49.000vfield_1.bin") == 12 * 24 ** 3
    h2, s2 = ck.read_checkpoint(str(tmp_path), 49.0, 1)
    assert int(h2["nplocal"]) == n and np.float32(h2["sigma_vi"]) == sig
    for k in ("xp", "vp", "rhoc", "vfield"):
        assert np.array_equal(s[k], s2[k])
    # Fortran order check: rhoc(i,j,k,itx,ity,itz) first index fastest == numpy [tz][ty][tx][k][j][i]
    raw = np.fromfile(d / "49.000zip2_1.bin", "<i4", offset=168)
    assert raw[1] == s["rhoc"][0, 0, 0, 0, 0, 1]


def test_cubenu_convention(tmp_path):
    """Same state in CUBEnu's files: 224-byte header alone in `info`, counts in `np`, names <z>_<name>_<image>.bin
    (CUBEnu/work/main/checkpoint.f90:10-50, parameters.f90:136-160, basic_functions.fh:52-87)."""
    from cafproject_b200.synthetic_ic import make_ic
    assert ck.HEADER_NU_DTYPE.itemsize == 224
    names = list(ck.HEADER_NU_DTYPE.names)
    assert names[:4] == ["nplocal", "npglobal", "nplocal_nu", "npglobal_nu"] and names[14:17] == ["timestep", "cur_checkpoint", "cur_halofind"]
    assert names[17:20] == ["a", "t", "tau"] and names[-1] == "vz_max" and ck.HEADER_NU_DTYPE.fields["a"][1] == 17 * 8
    assert ck.file_name("/o", 0.5, 2, "zip0", "cubenu") == "/o/image2/0.500_xp_2.bin"
    assert ck.file_name("/o", 0.5, 2, "vfield", "cubenu") == "/o/image2/0.500_vc_2.bin"
    states, sig, info = make_ic(nn=1, nc=24, nnt=2, np_nc=1, seed=5)
    s = states[0]
    n = s["xp"].shape[0]
    h = ck.make_header("cubenu", npglobal=n, izipx=2, izipv=2, image=1, nn=1, nnt=2, nt=12, ncell=4, ncb=6, a=0.02, sigma_vi=sig, z_i=49)
    ck.write_checkpoint(str(tmp_path), 49.0, 1, h, s, convention="cubenu")
    d = tmp_path / "image1"
    assert os.path.getsize(d / "49.000_info_1.bin") == 224
    assert os.path.getsize(d / "49.000_np_1.bin") == 4 * 24 ** 3
    assert os.path.getsize(d / "49.000_xp_1.bin") == 6 * n and os.path.getsize(d / "49.000_vc_1.bin") == 12 * 24 ** 3
    h2, s2 = ck.read_checkpoint(str(tmp_path), 49.0, 1, convention="cubenu")
    assert int(h2["nplocal"]) == n and int(h2["npglobal"]) == n and np.float32(h2["sigma_vi"]) == sig
    for k in ("xp", "vp", "rhoc", "vfield"):
        assert np.array_equal(s[k], s2[k])
