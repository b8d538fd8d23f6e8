"""examples/cafcube_driver.c: the reference's step loop over the C ABI from compiled C (no Python in between).  Here: it builds
against include/cube_gpu.h, links against libcubegpu.so, reads a CUBE checkpoint + the ASCII kernel tables, and -- on a machine
without a CUDA device -- stops with the library's message (the product has no CPU path)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_ascii_table(path, tab):
    """rows "i j k fx fy fz", i fastest (CUBE/kernels/wfxyz*.ascii, kernel_f.f90:19); ``tab`` is [k][j][i][3]."""
    n = tab.shape[0]
    with open(path, "w") as f:
        for k in range(n):
            for j in range(n):
                for i in range(n):
                    f.write("%d %d %d %.9e %.9e %.9e\n" % ((i + 1, j + 1, k + 1) + tuple(float(v) for v in tab[k, j, i])))


def _build_driver(tmp_path, tables):
    lib = os.path.join(ROOT, "cafproject_b200", "libcubegpu.so")
    if not os.path.exists(lib):
        pytest.skip("libcubegpu.so not built (run __graft_entry__.build())")
    exe = str(tmp_path / "cafcube_driver")
    cmd = ["gcc", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "examples", "cafcube_driver.c"), "-I" + os.path.join(ROOT, "include"),
           "-L" + os.path.join(ROOT, "cafproject_b200"), "-lcubegpu", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "cafproject_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    fk, ckt = tables
    kdir = tmp_path / "kernels"; kdir.mkdir()
    _write_ascii_table(str(kdir / "wfxyzf.3.ascii"), fk)
    _write_ascii_table(str(kdir / "wfxyzc.2.ascii"), ckt)
    return exe, kdir


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_driver_two_steps_equal_the_python_path(tmp_path, tables):
    """The only compiled-language caller of the ABI, run on the device: checkpoint in -> 2 fixed steps -> checkpoint out must be,
    byte for byte, what the Python mirror (cafproject_b200.cube.CubeGPU over the same library) produces from the same files."""
    from cafproject_b200 import checkpoint as ck
    from cafproject_b200.cube import CubeGPU, host_tanf_lut
    from cafproject_b200.synthetic_ic import make_ic
    exe, kdir = _build_driver(tmp_path, tables)
    fk, ckt = tables
    nc, nnt, dt, a_mid = 32, 2, np.float32(0.5), np.float32(0.0205)
    states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=3)
    hdr = ck.make_header(izipx=2, izipv=2, image=1, nn=1, nnt=nnt, nt=nc // nnt, ncell=4, ncb=6, sigma_vi=sig, mass_p=8.0, box=200.0)
    ck.write_checkpoint(str(tmp_path / "out"), 49.0, 1, hdr, states[0])
    run = subprocess.run([exe, str(kdir), str(tmp_path / "out"), "49.0", "2", repr(float(dt)), repr(float(a_mid)), "48.0"],
                         capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "done: %d particles" % info["npglobal"] in run.stdout, run.stdout + run.stderr
    h_c, st_c = ck.read_checkpoint(str(tmp_path / "out"), 48.0, 1)
    # the same two steps through the Python mirror, from the same checkpoint files
    h_in, st_in = ck.read_checkpoint(str(tmp_path / "out"), 49.0, 1)
    # the table the ASCII files hold (9 significant digits: exact for f32)
    G = CubeGPU(nc, nnt, fk, ckt, np_nc=2, tanf_lut=host_tanf_lut())
    try:
        G.particle_initialization(st_in, np.float32(h_in["sigma_vi"]))
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        dt_old = np.float32(0)
        for _ in range(2):
            G.step(dt_old, dt, a_mid)
            dt_old = dt
        st_p, sig_p = G.checkpoint()
    finally:
        G.close()
    for k in ("xp", "vp", "rhoc"):
        assert st_c[k].tobytes() == st_p[k].tobytes(), k
    assert st_c["vfield"].tobytes() == st_p["vfield"].tobytes()
    assert np.float32(h_c["sigma_vi"]) == sig_p and int(h_c["nplocal"]) == st_p["xp"].shape[0]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_driver_builds_links_and_fails_loudly_without_a_gpu(tmp_path, tables):
    from cafproject_b200 import checkpoint as ck
    from cafproject_b200.synthetic_ic import make_ic
    exe, kdir = _build_driver(tmp_path, tables)
    nc, nnt = 24, 2          # cube_gpu_init wants nc >= 24, nt >= 12 (parameters.f90:23-24)
    states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=1)
    hdr = ck.make_header(izipx=2, izipv=2, image=1, nn=1, nnt=nnt, nt=nc // nnt, ncell=4, ncb=6, sigma_vi=sig, mass_p=8.0, box=200.0)
    ck.write_checkpoint(str(tmp_path / "out"), 49.0, 1, hdr, states[0])
    run = subprocess.run([exe, str(kdir), str(tmp_path / "out"), "49.0", "1", "0.5", "0.0205", "48.0"], capture_output=True, text=True, timeout=300)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "done: %d particles" % info["npglobal"] in run.stdout, run.stdout + run.stderr
        _, st = ck.read_checkpoint(str(tmp_path / "out"), 48.0, 1)
        assert st["xp"].shape[0] == info["npglobal"] and int(st["rhoc"].sum()) == info["npglobal"]
    else:
        assert run.returncode != 0
        assert "no CUDA device" in run.stderr or "CUDA" in run.stderr, run.stderr
