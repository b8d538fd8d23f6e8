"""CPU tests of the oracle's variants beyond CUBE/main x2v2: (1) the reference's other zip formats and (2) CUBEnu's order of
the particles inside a destination cell (bottom of the file).

(1) izipx, izipv in {1, 2} bytes per code
(CUBE/main/universe6.fh: x1v2, universe7.fh: x2v1, universe8.fh: x1v1; parameters.f90:13-15,101; variables.f90:41-42).

The GPU library is built for x2v2 only and rejects the rest at init ("zip format incompatable",
particle_initialization.f90:14-18); these tests make the checker for the 1-byte formats ready and pin it on what the
reference offers: the paper's 4-particle decode example (which *is* a 1-byte example, ms_caf/ms_caf.tex:78), the decode /
encode identities of parameters.f90:14-15, and the run-time invariants the Fortran code stops on.  Parity otherwise unpinned
by the reference (no golden outputs exist).
"""
import os

import numpy as np
import pytest

from conftest import physical

F32 = np.float32
FORMATS = [(1, 1), (1, 2), (2, 1)]


@pytest.fixture(scope="module")
def co():
    from oracle import cube_oracle
    cube_oracle.build()
    return cube_oracle


def make(co, tables, izipx, izipv, nc=24, nnt=2, np_nc=2, seed=3, disp_rms=0.7, nn=1):
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, info = make_ic(nn=nn, nc=nc, nnt=nnt, np_nc=np_nc, seed=seed, disp_rms=disp_rms, izipx=izipx, izipv=izipv)
    O = co.Oracle(nn=nn, nnt=nnt, nc=nc, np_nc=np_nc, fk_table=fk, ck_table=ck, izipx=izipx, izipv=izipv)
    O.load(states, sig)
    return O, states, sig


def test_paper_decode_example_through_the_oracle(co):
    """ms_caf.tex:66-78 evaluated by the C oracle's own xq expression (update_particle.f90:41) in the 1-byte format:
    chi=(-128,127,0,60) offset-binary, rho_c=(1,0,2,1) -> x=(0.001953125, 2.998046875, 2.501953125, 3.736328125)."""
    O = co.Oracle(nc=24, nnt=2, izipx=1, izipv=1)
    L = co.lib()
    chi = [-128, 127, 0, 60]
    cell = [1, 3, 3, 4]                       # n_c from the prefix sum of rho_c = (1,0,2,1)
    want = [0.001953125, 2.998046875, 2.501953125, 3.736328125]
    for c, x, w in zip(cell, chi, want):
        u = x + 128                           # fraction bin; the code stores its two's-complement wrap
        code = u - 256 if u >= 128 else u
        assert L.oracle_probe_xq(O.h, c, code) == w
    O.close()


@pytest.mark.parametrize("izipx", [1, 2])
def test_decode_identity_every_code(co, izipx):
    """int(xp+ishift,izipx)+rshift == u+0.5 for every code of the format (parameters.f90:14-15)."""
    O = co.Oracle(nc=24, nnt=2, izipx=izipx, izipv=2)
    L = co.lib()
    n = 1 << (8 * izipx)
    for u in range(0, n, 1 if izipx == 1 else 97):
        code = u - n if u >= n // 2 else u
        assert L.oracle_probe_xq(O.h, 5, code) == 4.0 + (u + 0.5) / n
    O.close()


@pytest.mark.parametrize("izipv", [1, 2])
def test_velocity_codec_round_trip_and_table(co, izipv):
    """encode(decode(c)) == c for every code (pm.f90:102,113), the table is host tanf of the same expression and odd."""
    O = co.Oracle(nc=24, nnt=2, izipx=2, izipv=izipv)
    L = co.lib()
    n = 1 << (8 * izipv)
    lut = co.tanf_lut(izipv)
    assert lut.shape == (n,) and lut.dtype == F32
    codes = np.arange(n, dtype=np.int64)
    signed = np.where(codes >= n // 2, codes - n, codes)
    want = np.tan((F32(co.PI_F) * signed.astype(F32)) / F32(n - 1), dtype=F32)
    assert float(np.abs(lut - want).max()) <= 1e-6 * float(np.abs(want[np.abs(signed) < n // 2 - 1]).max())
    assert np.array_equal(lut[1:n // 2], -lut[:n // 2:-1])
    sig = F32(0.37)
    S = float(np.float64(np.sqrt(F32(co.PI_F / F32(2)))) / (np.float64(sig) * 2.5))
    for c in range(-(n // 2 - 1), n // 2, 1 if izipv == 1 else 61):
        v = L.oracle_probe_vdecode(O.h, c, sig)
        assert v == float(np.float64(lut[c % n]) / S)
        assert L.oracle_probe_vencode(O.h, v, sig) == c
    O.close()


@pytest.mark.parametrize("izipx,izipv", FORMATS)
def test_buffer_and_update_keep_particles(co, tables, izipx, izipv):
    """buffer_density.f90:99-109,143-146 and update_particle.f90:205-211 in the 1-byte formats; the state comes back in
    its own dtype."""
    O, states, sig = make(co, tables, izipx, izipv)
    s0 = states[0]
    assert s0["xp"].dtype == (np.int8, np.int16)[izipx - 1] and s0["vp"].dtype == (np.int8, np.int16)[izipv - 1]
    n0 = s0["xp"].shape[0]
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    assert np.array_equal(physical(O, "xp"), s0["xp"])
    assert np.array_equal(physical(O, "vp"), s0["vp"])
    up = O.update_particle(F32(0), F32(1.0))
    s1 = O.store(0)
    assert O.nplocal(0) == n0 == int(s1["rhoc"].sum())
    assert s1["xp"].dtype == s0["xp"].dtype and s1["vp"].dtype == s0["vp"].dtype
    assert up["sigma_vi_new"] > 0
    O.close()


@pytest.mark.parametrize("izipx,izipv", FORMATS)
def test_drift_is_pure_integer_given_v(co, tables, izipx, izipv):
    """xp_new = xp + nint(dt_mid*v/(x_resolution*ncell)) wraps mod 2^(8*izipx) and the destination cell is the carry
    (update_particle.f90:44-45,84): same multiset of global fixed-point positions as the integer prediction."""
    O, states, sig = make(co, tables, izipx, izipv, disp_rms=0.5)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    s0 = states[0]
    bits = 8 * izipx
    udt = (np.uint8, np.uint16)[izipx - 1]
    nv = 1 << (8 * izipv)
    lut = co.tanf_lut(izipv)
    S = float(np.float64(np.sqrt(F32(co.PI_F / F32(2)))) / (np.float64(sig) * 2.5))
    rho = s0["rhoc"]
    nt, nc = O.nt, O.nc
    tz, ty, tx, k, j, i = np.meshgrid(*[np.arange(n) for n in rho.shape], indexing="ij")
    cells = lambda r: np.stack([np.repeat((t * nt + c).ravel(), r.ravel()) for t, c in ((tx, i), (ty, j), (tz, k))], 1).astype(np.int64)
    vf = np.repeat(s0["vfield"].reshape(-1, 3), rho.ravel(), axis=0).astype(np.float64)
    v = lut[s0["vp"].astype(np.int64) % nv].astype(np.float64) / S + vf
    dt_mid = np.float64(F32((F32(0) + F32(1.0)) / F32(2)))
    inc = np.rint(np.abs(dt_mid * v * float((1 << bits) // 4))) * np.sign(v)    # nint: half away from zero
    pos0 = (cells(rho) << bits) + s0["xp"].view(udt).astype(np.int64)
    pos1 = (pos0 + inc.astype(np.int64)) % (nc << bits)
    O.update_particle(F32(0), F32(1.0))
    s1 = O.store(0)
    got = (cells(s1["rhoc"]) << bits) + s1["xp"].view(udt).astype(np.int64)
    key = lambda p: np.sort(p[:, 0] * (nc << bits) ** 2 + p[:, 1] * (nc << bits) + p[:, 2])
    assert int((key(pos1) != key(got)).sum()) == 0
    O.close()


@pytest.mark.parametrize("izipx", [1, 2])
def test_coarse_deposit_against_an_independent_numpy_cic(co, tables, izipx):
    """pm.f90:136-160 restated a second time with numpy in f64 (x = cell-1 + (u+0.5)*2^-8izipx - 0.5, periodic CIC):
    the oracle's f32 mesh agrees to single-precision round-off and carries N*mass_p."""
    O, states, sig = make(co, tables, izipx, 2)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    s0 = states[0]
    rho = s0["rhoc"]
    nt, nc = O.nt, O.nc
    udt = (np.uint8, np.uint16)[izipx - 1]
    tz, ty, tx, k, j, i = np.meshgrid(*[np.arange(n) for n in rho.shape], indexing="ij")
    cell = np.stack([np.repeat((t * nt + c).ravel(), rho.ravel()) for t, c in ((tx, i), (ty, j), (tz, k))], 1).astype(np.float64)
    x = cell + (s0["xp"].view(udt).astype(np.float64) + 0.5) / float(1 << (8 * izipx)) - 0.5
    i1 = np.floor(x).astype(np.int64)
    w2 = x - i1; w1 = 1.0 - w2
    want = np.zeros((nc, nc, nc))
    mp = float(O.mass_p)
    for a, wx in ((0, w1[:, 0]), (1, w2[:, 0])):
        for b, wy in ((0, w1[:, 1]), (1, w2[:, 1])):
            for c, wz in ((0, w1[:, 2]), (1, w2[:, 2])):
                np.add.at(want, ((i1[:, 2] + c) % nc, (i1[:, 1] + b) % nc, (i1[:, 0] + a) % nc), wx * wy * wz * mp)
    r3 = O.coarse_density()
    n = s0["xp"].shape[0]
    assert abs(float(r3.sum(dtype=np.float64)) - n * mp) < 1e-5 * n * mp
    assert float(np.abs(r3 - want).max()) < 1e-5 * float(want.max())
    O.close()


@pytest.mark.parametrize("izipx,izipv", [(1, 1)])
def test_full_step_and_multi_image_equals_single_image(co, tables, izipx, izipv):
    """One whole step (cafcube.f90:27-31) in x1v1: particle count kept, time-step limits finite, and two images per
    dimension in x give the same integer state as the single image (the buffers cross images by copy only)."""
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    O, states, sig = make(co, tables, izipx, izipv, nc=24, nnt=2)
    n0 = states[0]["xp"].shape[0]
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    up, pm = O.step(F32(0), F32(0.5), F32(0.021))
    assert O.nplocal(0) == n0
    assert all(np.isfinite(float(pm[k])) and float(pm[k]) > 0 for k in ("dt_fine", "dt_coarse", "dt_vmax"))
    O.close()
    nn = (2, 1, 1)
    st2, sig2, _ = make_ic(nn=nn, nc=24, nnt=2, np_nc=1, seed=9, disp_rms=0.6, izipx=izipx, izipv=izipv)
    O2 = co.Oracle(nn=nn, nnt=2, nc=24, np_nc=1, fk_table=fk, ck_table=ck, izipx=izipx, izipv=izipv)
    O2.load(st2, sig2)
    O2.buffer_density(); O2.buffer_x(); O2.buffer_v()
    O2.update_particle(F32(0), F32(1.0))
    tot = sum(O2.nplocal(m) for m in range(2))
    assert tot == sum(s["xp"].shape[0] for s in st2)
    for m in range(2):
        s = O2.store(m)
        assert s["xp"].dtype == np.int8 and int(s["rhoc"].sum()) == O2.nplocal(m)
    O2.close()


@pytest.mark.parametrize("izipx,izipv", FORMATS)
def test_checkpoint_files_in_one_byte_formats(tmp_path, izipx, izipv):
    """checkpoint.f90:33-70 writes xp/vp in their own kind: 3*izip bytes per particle; a run built for another format stops
    with "zip format incompatable" (particle_initialization.f90:14-18)."""
    from cafproject_b200 import checkpoint as ck
    from cafproject_b200.synthetic_ic import make_ic
    states, sig, info = make_ic(nn=1, nc=24, nnt=2, np_nc=1, seed=4, izipx=izipx, izipv=izipv)
    s = states[0]
    n = s["xp"].shape[0]
    h = ck.make_header(izipx=izipx, izipv=izipv, image=1, nn=1, nnt=2, nt=12, ncell=4, ncb=6, a=0.02, sigma_vi=sig, z_i=49, mass_p=64.0)
    ck.write_checkpoint(str(tmp_path), 49.0, 1, h, s)
    d = tmp_path / "image1"
    assert os.path.getsize(d / "49.000zip0_1.bin") == 3 * izipx * n
    assert os.path.getsize(d / "49.000zip1_1.bin") == 3 * izipv * n
    h2, s2 = ck.read_checkpoint(str(tmp_path), 49.0, 1)
    for k in ("xp", "vp", "rhoc", "vfield"):
        assert s2[k].dtype == s[k].dtype and np.array_equal(s[k], s2[k])
    with pytest.raises(ValueError, match="zip format incompatable"):
        ck.read_checkpoint(str(tmp_path), 49.0, 1, expect_zip=(2, 2))
    bad = ck.make_header(izipx=2, izipv=2, image=1, nn=1, nnt=2, nt=12, ncell=4, ncb=6)
    with pytest.raises(ValueError, match="zip format incompatable"):
        ck.write_checkpoint(str(tmp_path), 48.0, 1, bad, s)


# ---- CUBEnu's order of the particles inside a destination cell ---------------------------------------
def _predicted_order(co, O, s0, sig, nlayer):
    """xp after the drift predicted without the oracle's loops: integer increments give every particle its destination cell;
    inside a destination cell the particles arrive in the order the tile's double loop visits their source cells --
    colour pass ilayer = (k_rel-(1-ncb)) mod nlayer first (CUBEnu update_particle.f90:55-58; nlayer = 1: CUBE/main
    update_particle.f90:70-75), then k, j, i of the source cell relative to the destination's tile, then storage order."""
    rho = s0["rhoc"]
    nt, nc, nnt, ncb = O.nt, O.nc, O.nnt, O.ncb
    lut = co.tanf_lut(2)
    S = float(np.float64(np.sqrt(F32(co.PI_F / F32(2)))) / (np.float64(sig) * 2.5))
    tz, ty, tx, k, j, i = np.meshgrid(*[np.arange(n) for n in rho.shape], indexing="ij")
    src = np.stack([np.repeat((t * nt + c).ravel(), rho.ravel()) for t, c in ((tx, i), (ty, j), (tz, k))], 1).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(rho.ravel())[:-1]])
    l = np.arange(src.shape[0]) - np.repeat(start, rho.ravel())
    vf = np.repeat(s0["vfield"].reshape(-1, 3), rho.ravel(), axis=0).astype(np.float64)
    v = lut[s0["vp"].view(np.uint16)].astype(np.float64) / S + vf
    dt_mid = np.float64(F32((F32(0) + F32(1.0)) / F32(2)))
    inc = (np.rint(np.abs(dt_mid * v * 16384.0)) * np.sign(v)).astype(np.int64)
    pos1 = ((src << 16) + s0["xp"].view(np.uint16).astype(np.int64) + inc) % (nc << 16)
    dst = pos1 >> 16
    dt_, dc = dst // nt, dst % nt                      # destination tile and cell in it
    rel = (src - dt_ * nt + ncb - 1) % nc - (ncb - 1)   # source cell relative to the destination's tile, in 1-ncb..nt+ncb (0-based: -ncb..)
    ilayer = (rel[:, 2] + ncb) % nlayer                 # 0-based rel = Fortran k-1, so k-(1-ncb) = rel+ncb
    dkey = ((((dt_[:, 2] * nnt + dt_[:, 1]) * nnt + dt_[:, 0]) * nt + dc[:, 2]) * nt + dc[:, 1]) * nt + dc[:, 0]
    order = np.lexsort((l, rel[:, 0], rel[:, 1], rel[:, 2], ilayer, dkey))
    xp1 = (pos1 & 0xFFFF).astype(np.uint16).view(np.int16)
    return xp1[order], np.bincount(dkey, minlength=rho.size).reshape(rho.shape)


@pytest.mark.parametrize("vz_max", [None, 0.0, 1.0, 7.0])
def test_in_cell_order_main_and_cubenu(co, tables, vz_max):
    """The oracle's re-sorted positions equal an independent prediction of (destination cell, arrival order) for CUBE/main
    (vz_max None: storage order) and for CUBEnu's colour passes (nlayer = 1, 3, 9 here)."""
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=24, nnt=2, np_nc=2, seed=11, disp_rms=0.8)
    O = co.Oracle(nn=1, nnt=2, nc=24, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    O.update_particle(F32(0), F32(1.0), vz_max=vz_max)
    want_nlayer = 1 if vz_max is None else 2 * int(np.ceil(F32(0.5) * F32(vz_max) / F32(4))) + 1
    assert O.nlayer == want_nlayer
    s1 = O.store(0)
    xp_pred, rho_pred = _predicted_order(co, O, states[0], sig, O.nlayer)
    assert np.array_equal(s1["rhoc"], rho_pred)
    assert np.array_equal(s1["xp"], xp_pred)
    O.close()


def test_cubenu_order_changes_only_the_order(co, tables):
    """Colour passes permute particles inside cells and reorder the f32 sums of vfield_new: counts are identical, vfield agrees to
    f32 round-off, per-cell multisets of positions are identical."""
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=24, nnt=2, np_nc=2, seed=12, disp_rms=0.8)
    out = []
    for vz in (None, 9.0):
        O = co.Oracle(nn=1, nnt=2, nc=24, np_nc=2, fk_table=fk, ck_table=ck)
        O.load(states, sig)
        O.buffer_density(); O.buffer_x(); O.buffer_v()
        up = O.update_particle(F32(0), F32(1.0), vz_max=vz)
        out.append((O.store(0), up, O.nlayer))
        O.close()
    (a, ua, la), (b, ub, lb) = out
    assert la == 1 and lb == 5
    assert np.array_equal(a["rhoc"], b["rhoc"])
    assert not np.array_equal(a["xp"], b["xp"])
    cell = np.repeat(np.arange(a["rhoc"].size), a["rhoc"].ravel())
    canon = lambda s: s["xp"][np.lexsort((s["xp"][:, 2], s["xp"][:, 1], s["xp"][:, 0], cell))]
    assert np.array_equal(canon(a), canon(b))
    scale = float(np.abs(a["vfield"]).max())
    assert float(np.abs(a["vfield"] - b["vfield"]).max()) < 1e-5 * scale
    assert abs(float(ua["sigma_vi_new"]) - float(ub["sigma_vi_new"])) < 1e-5 * float(ua["sigma_vi_new"])


def test_cubenu_vmax3(co, tables):
    """CUBEnu pm.f90:349,398: vmax(3) = max |v| per component (CUBE/main: one scalar, no abs, pm.f90:220)."""
    O, states, sig = make(co, tables, 2, 2)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    up, pm = O.step(F32(0), F32(0.5), F32(0.021))
    v3 = np.array(pm["vmax3"][0], np.float64)
    assert (v3 > 0).all() and float(pm["vmax"][0]) <= v3.max() * (1 + 1e-6)
    O.close()


# ---- -DPID: particle IDs ride with vp (CUBE/main buffer_density.f90:111-135, buffer_v.f90:22-42, update_particle.f90:88,106) ------
def _with_pid(states):
    out, base = [], 0
    for s in states:
        n = s["xp"].shape[0]
        out.append(dict(s, pid=np.arange(base + 1, base + n + 1, dtype=np.int64)))   # all positive (checkpoint.f90 checks that)
        base += n
    return out


@pytest.mark.parametrize("vz_max", [None, 9.0])
def test_pid_tracks_every_particle_through_the_drift(co, tables, vz_max):
    """With IDs the drift can be checked particle by particle instead of as a multiset: the particle with ID p ends in the
    cell and with the position code that the integer increment predicts for it, buffers keep the IDs of the physical
    particles, ghosts carry their source's ID, and particle_mesh does not touch them."""
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=24, nnt=2, np_nc=2, seed=21, disp_rms=0.8)
    states = _with_pid(states)
    s0 = states[0]
    n = s0["xp"].shape[0]
    O = co.Oracle(nn=1, nnt=2, nc=24, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    assert np.array_equal(physical(O, "pid"), s0["pid"])
    # every slot of the buffered layout that holds a particle holds an ID in 1..n, and a ghost's codes are its source's
    pid_all, vp_all, xp_all = O.pid(0), O.vp(0), O.xp(0)
    used = pid_all > 0
    assert int(used.sum()) == int(O.rhoc(0).sum()) and int(pid_all.max()) == n
    assert np.array_equal(vp_all[used], s0["vp"][pid_all[used] - 1])
    assert np.array_equal(xp_all[used], s0["xp"][pid_all[used] - 1])
    # prediction per ID
    rho, nt, nc = s0["rhoc"], O.nt, O.nc
    tz, ty, tx, k, j, i = np.meshgrid(*[np.arange(m) for m in rho.shape], indexing="ij")
    cells = lambda r: np.stack([np.repeat((t * nt + c).ravel(), r.ravel()) for t, c in ((tx, i), (ty, j), (tz, k))], 1).astype(np.int64)
    lut = co.tanf_lut(2)
    S = float(np.float64(np.sqrt(F32(co.PI_F / F32(2)))) / (np.float64(sig) * 2.5))
    vf = np.repeat(s0["vfield"].reshape(-1, 3), rho.ravel(), axis=0).astype(np.float64)
    v = lut[s0["vp"].view(np.uint16)].astype(np.float64) / S + vf
    inc = (np.rint(np.abs(0.5 * v * 16384.0)) * np.sign(v)).astype(np.int64)
    pos1 = ((cells(rho) << 16) + s0["xp"].view(np.uint16).astype(np.int64) + inc) % (nc << 16)
    O.update_particle(F32(0), F32(1.0), vz_max=vz_max)
    s1 = O.store(0)
    assert np.array_equal(np.sort(s1["pid"]), s0["pid"])                      # a permutation
    got = (cells(s1["rhoc"]) << 16) + s1["xp"].view(np.uint16).astype(np.int64)
    assert np.array_equal(got, pos1[s1["pid"] - 1])                           # particle by particle
    O.buffer_density(); O.buffer_x()
    O.particle_mesh(F32(0.021), F32(1.0))
    O.buffer_v()
    assert np.array_equal(physical(O, "pid"), s1["pid"])
    O.close()


def test_pid_across_images(co, tables):
    """Two images in x: IDs are global, the drift hands particles (and their IDs) across the image boundary through the
    ghost zones, nothing is lost or duplicated (update_particle.f90:205-211 with IDs)."""
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    nn = (2, 1, 1)
    states, sig, _ = make_ic(nn=nn, nc=24, nnt=2, np_nc=1, seed=22, disp_rms=1.5)
    states = _with_pid(states)
    ntot = sum(s["xp"].shape[0] for s in states)
    O = co.Oracle(nn=nn, nnt=2, nc=24, np_nc=1, fk_table=fk, ck_table=ck)
    O.load(states, sig)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    O.update_particle(F32(0), F32(1.0))
    after = [O.store(m) for m in range(2)]
    allp = np.concatenate([s["pid"] for s in after])
    assert np.array_equal(np.sort(allp), np.arange(1, ntot + 1))
    n0 = states[0]["xp"].shape[0]
    moved = int((after[0]["pid"] > n0).sum()) + int((after[1]["pid"] <= n0).sum())
    assert moved > 0                                                           # some particles did change image
    O.close()


@pytest.mark.parametrize("convention,idbytes,idname", [("cube", 8, "49.000zipid_1.bin"), ("cubenu", 4, "49.000_id_1.bin")])
def test_checkpoint_id_files(tmp_path, convention, idbytes, idname):
    """-DPID checkpoints: CUBE/main `zipid` integer(8) (particle_initialization.f90:56), CUBEnu `id` integer(4)
    (checkpoint.f90:17,47)."""
    from cafproject_b200 import checkpoint as ck
    from cafproject_b200.synthetic_ic import make_ic
    states, sig, _ = make_ic(nn=1, nc=24, nnt=2, np_nc=1, seed=4)
    s = _with_pid(states)[0]
    n = s["xp"].shape[0]
    h = ck.make_header(convention, izipx=2, izipv=2, image=1, nn=1, nnt=2, nt=12, ncell=4, ncb=6, sigma_vi=sig)
    ck.write_checkpoint(str(tmp_path), 49.0, 1, h, s, convention=convention)
    assert os.path.getsize(tmp_path / "image1" / idname) == idbytes * n
    _, s2 = ck.read_checkpoint(str(tmp_path), 49.0, 1, convention=convention)
    assert np.array_equal(s2["pid"], s["pid"])


# ---- edge cases of the domain -----------------------------------------------------------------------------------------
def test_zero_time_step_keeps_positions_and_counts(co, tables):
    """dt_mid = 0: nobody moves -- counts, order and position codes are unchanged (the velocities are only re-expressed against
    the rebuilt vfield); a second zero step changes nothing in them either (idempotence)."""
    O, states, sig = make(co, tables, 2, 2)
    s0 = states[0]
    for _ in range(2):
        O.buffer_density(); O.buffer_x(); O.buffer_v()
        O.update_particle(F32(0), F32(0))
        s1 = O.store(0)
        assert np.array_equal(s1["rhoc"], s0["rhoc"]) and np.array_equal(s1["xp"], s0["xp"])
    O.close()


def test_empty_tiles_and_one_crowded_cell(co, tables):
    """Ragged input: every particle of the image in one coarse cell of one tile (7 tiles empty, 13 823 empty cells), at rest
    relative to the cell flow.  Buffers, drift and both meshes go through; the count stays in the cell (or moves as a block
    with the common velocity), the coarse mesh carries the whole mass on the 8 nodes around it."""
    fk, ck = tables
    nc, nnt, n = 24, 2, 500
    nt = nc // nnt
    rng = np.random.default_rng(5)
    rhoc = np.zeros((nnt,) * 3 + (nt,) * 3, np.int32)
    rhoc[1, 0, 1, 3, 4, 5] = n
    vfield = np.zeros(rhoc.shape + (3,), np.float32)
    state = dict(xp=rng.integers(-32768, 32768, (n, 3)).astype(np.int16), vp=np.zeros((n, 3), np.int16), rhoc=rhoc, vfield=vfield)
    O = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    O.load([state], F32(0.1))
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    r3 = O.coarse_density()
    mp = float(O.mass_p)
    assert abs(float(r3.sum(dtype=np.float64)) - n * mp) < 1e-5 * n * mp
    assert int((r3 != 0).sum()) <= 27 and int((r3 != 0).sum()) >= 8
    up, pm = O.step(F32(0), F32(0.01), F32(0.02))
    s1 = O.store(0)
    assert int(s1["rhoc"].sum()) == n and O.nplocal(0) == n
    assert int((s1["rhoc"] != 0).sum()) <= 8                                   # a tiny step: at most the neighbouring cells
    assert all(np.isfinite(float(pm[k])) and float(pm[k]) > 0 for k in ("dt_fine", "dt_coarse"))
    O.close()


def test_extreme_velocity_codes_round_trip(co):
    """The largest codes (+-32767 at v2, +-127 at v1) decode to finite velocities and encode back to themselves; the code
    -32768 / -128 is never produced by the encoder (|nint(N*atan(x)/pi)| <= (N-1)/2)."""
    for izipv in (1, 2):
        O = co.Oracle(nc=24, nnt=2, izipx=2, izipv=izipv)
        L = co.lib()
        top = (1 << (8 * izipv - 1)) - 1
        for c in (top, -top):
            v = L.oracle_probe_vdecode(O.h, c, F32(0.2))
            assert np.isfinite(v) and L.oracle_probe_vencode(O.h, v, F32(0.2)) == c
        assert L.oracle_probe_vencode(O.h, 1e300, F32(0.2)) == top and L.oracle_probe_vencode(O.h, -1e300, F32(0.2)) == -top
        O.close()


def test_two_species_composition_reduces_to_one_species(tables):
    """oracle.particle_mesh_two_species (CUBEnu pm.f90 with NEUTRINOS, composed from the one-species restatement) on a one-species
    state dealt alternately to two species of the same particle mass and sigma_vi: the same time-step limits and, up to rare flips of
    codes on a quantiser boundary (the two deposits are summed in another order), the same kicked velocities as particle_mesh."""
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    from test_gpu_two_species import _split
    fk, ck = tables
    nc, nnt = 24, 2
    st, sig, _ = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=5, disp_rms=0.7)
    a_mid, dt = np.float32(0.021), np.float32(0.8)
    O = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    A = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    B = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    try:
        O.load(st, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
        p1 = O.particle_mesh(a_mid, dt)
        (ma, sa), (mb, sb) = _split(st[0])
        for X, s in ((A, sa), (B, sb)):
            X.load([s], sig); X.set_mass_p(O.mass_p)
            assert X.mass_p == O.mass_p
            X.buffer_density(); X.buffer_x(); X.buffer_v()
        p2 = co.particle_mesh_two_species(A, B, a_mid, dt)
        for k in ("dt_fine", "dt_coarse", "dt_vmax"):
            assert abs(float(p1[k]) - float(p2[k])) <= 1e-6 * float(p1[k]), k
        assert max(float(p2["vmax"]), float(p2["vmax2"])) == float(max(p1["vmax"]))
        v1 = physical(O, "vp")
        for m, X in ((ma, A), (mb, B)):
            assert np.array_equal(physical(X, "xp"), physical(O, "xp")[m])
            dv = np.abs(v1[m].astype(np.int32) - physical(X, "vp").astype(np.int32))
            assert dv.max() <= 2 and (dv != 0).mean() < 2e-3
        assert A.sigma_vi == O.sigma_vi and B.sigma_vi == O.sigma_vi     # pm.f90:122 on both
    finally:
        O.close(); A.close(); B.close()
