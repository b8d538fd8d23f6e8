"""A second particle species on the same meshes (CUBEnu -DNEUTRINOS: pm.f90:79-99,160,235,356; cube_gpu_particle_mesh_species).

The reference's two-species build does not compile upstream and the oracle restates the one-species path, so this is a test by
construction: a one-species state the oracle and the other GPU tests cover is SPLIT into two species -- the particles of every
cell dealt alternately to A and B, same particle mass, same sigma_vi and vfield -- each species living in its own handle.  The sum
of the two deposits is the one-species density (each term is the same f32 value; only the fixed-point grouping differs), so the
forces, the time-step limits and every particle's kicked velocity code must come out as in the one-species run, up to the rare
one-unit flips of codes that sit on a quantiser boundary.  Then each species drifts on its own and the counts add up.
"""
import numpy as np
import pytest

from conftest import physical

pytestmark = pytest.mark.gpu

NC, NNT, NP_NC = 24, 2, 2


def _split(state):
    cnt = state["rhoc"].reshape(-1).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(cnt)])
    cell = np.repeat(np.arange(cnt.size), cnt)
    within = np.arange(cell.size) - start[cell]
    out = []
    for parity in (0, 1):
        m = (within % 2) == parity
        rc = np.bincount(cell[m], minlength=cnt.size).astype(np.int32).reshape(state["rhoc"].shape)
        out.append((m, dict(xp=np.ascontiguousarray(state["xp"][m]), vp=np.ascontiguousarray(state["vp"][m]), rhoc=rc, vfield=state["vfield"])))
    return out


def test_two_species_equal_the_one_species_run(tables):
    from cafproject_b200.cube import CubeGPU, host_tanf_lut
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, info = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=61, disp_rms=0.9)
    st = states[0]
    n = st["xp"].shape[0]
    lut = host_tanf_lut()
    a_mid, dt = np.float32(0.021), np.float32(0.8)
    mass_p = float((4 * NC) ** 3) / n
    # one species
    G = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=lut)
    G.particle_initialization(st, sig); G.buffer_density(); G.buffer_x(); G.buffer_v()
    pm1 = G.particle_mesh(a_mid, dt)
    one, _ = G.checkpoint()
    one = {k: np.array(v, copy=True) for k, v in one.items()}
    G.close()
    # the same particles as two species
    (ma, sa), (mb, sb) = _split(st)
    assert sa["xp"].shape[0] + sb["xp"].shape[0] == n and min(sa["xp"].shape[0], sb["xp"].shape[0]) > n // 3
    GA = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=lut)
    GB = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=lut)
    try:
        for Gs, s in ((GA, sa), (GB, sb)):
            Gs.particle_initialization(s, sig, npglobal=s["xp"].shape[0])
            Gs.set_mass_p(mass_p)
            Gs.buffer_density(); Gs.buffer_x(); Gs.buffer_v()
        pm2 = GA.particle_mesh_species(GB, a_mid, dt)
        for Gs in (GA, GB):
            Gs.buffer_v()
        ca, _ = GA.checkpoint(); cb, _ = GB.checkpoint()
        for k in ("dt_fine", "dt_coarse"):
            assert abs(float(pm2[k]) - float(pm1[k])) <= 1e-5 * float(pm1[k]), k
        assert max(float(pm2["vmax"]), float(pm2["vmax2"])) == float(pm1["vmax"])
        for m, c in ((ma, ca), (mb, cb)):
            assert np.array_equal(c["xp"], one["xp"][m])                     # particle_mesh moves nobody
            dv = np.abs(c["vp"].astype(np.int32) - one["vp"][m].astype(np.int32))
            assert dv.max() <= 2 and (dv != 0).mean() < 1e-3, (dv.max(), (dv != 0).mean())
        # deterministic: the same call on the same states again gives the same codes
        for Gs, s in ((GA, sa), (GB, sb)):
            Gs.particle_initialization(s, sig, npglobal=s["xp"].shape[0]); Gs.set_mass_p(mass_p)
            Gs.buffer_density(); Gs.buffer_x(); Gs.buffer_v()
        GA.particle_mesh_species(GB, a_mid, dt)
        ca2, _ = GA.checkpoint(); cb2, _ = GB.checkpoint()
        assert np.array_equal(ca2["vp"], ca["vp"]) and np.array_equal(cb2["vp"], cb["vp"])
        # each species drifts through its own cell arrays; nobody is lost
        for Gs in (GA, GB):
            Gs.buffer_density(); Gs.buffer_x(); Gs.buffer_v()
        ua = GA.update_particle(dt, dt); ub = GB.update_particle(dt, dt)
        assert ua["nplocal"] + ub["nplocal"] == n
    finally:
        GA.close(); GB.close()


def test_two_species_against_the_composed_oracle(tables):
    """Two DIFFERENT species (own particles, own mass -- 90 % / 10 % of the total -- own velocity dispersion and sigma_vi, as CDM and
    a hot light species) against CUBEnu's two-species particle_mesh as composed from the one-species restatement
    (oracle/cube_oracle.py::particle_mesh_two_species; that build does not compile upstream: parity unpinned).  Each side computes
    its own FFTs, so the gates are those of a full step (test_gpu_parity.py::test_full_steps): time-step limits to 1e-4, velocity
    codes equal except for rare flips of codes on a quantiser boundary."""
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    fk, ck = tables
    sa, sig_a, ia = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=71, disp_rms=0.8)
    sb, sig_b, ib = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=72, disp_rms=0.8, velocity_boost=3.0)
    assert sig_b > 2 * sig_a
    na, nb = sa[0]["xp"].shape[0], sb[0]["xp"].shape[0]
    nf3 = float((4 * NC) ** 3)
    mass_a, mass_b = np.float32(0.9 * nf3 / na), np.float32(0.1 * nf3 / nb)
    a_mid, dt = np.float32(0.021), np.float32(0.8)
    OA = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=NP_NC, fk_table=fk, ck_table=ck)
    OB = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=NP_NC, fk_table=fk, ck_table=ck)
    lut = co.tanf_lut()
    GA = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=lut)
    GB = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=lut, secondary=True)
    try:
        for O, st, sig, mp in ((OA, sa, sig_a, mass_a), (OB, sb, sig_b, mass_b)):
            O.load(st, sig); O.set_mass_p(mp)
            O.buffer_density(); O.buffer_x(); O.buffer_v()
        po = co.particle_mesh_two_species(OA, OB, a_mid, dt)
        for G, st, sig, mp in ((GA, sa, sig_a, mass_a), (GB, sb, sig_b, mass_b)):
            G.particle_initialization(st[0], sig); G.set_mass_p(mp)
            G.buffer_density(); G.buffer_x(); G.buffer_v()
        pg = GA.particle_mesh_species(GB, a_mid, dt)
        for k in ("dt_fine", "dt_coarse"):
            assert abs(float(pg[k]) - float(po[k])) <= 1e-4 * float(po[k]), (k, pg[k], po[k])
        for k in ("vmax", "vmax2"):
            assert abs(float(pg[k]) - float(po[k])) <= 1e-3 * float(po[k]), (k, pg[k], po[k])
        assert float(pg["vmax2"]) > 1.5 * float(pg["vmax"])        # the second species is the hot one
        for G, O, name in ((GA, OA, "A"), (GB, OB, "B")):
            got, _ = G.checkpoint()
            vo = physical(O, "vp", 0)
            assert np.array_equal(got["xp"], physical(O, "xp", 0)), name
            dv = np.abs(got["vp"].astype(np.int32) - vo.astype(np.int32))
            assert dv.max() <= 2 and (dv != 0).mean() < 2e-3, (name, dv.max(), (dv != 0).mean())
        # the deposit did weigh the species: with equal masses instead the forces (and dt_fine) differ
        for G, st, sig in ((GA, sa, sig_a), (GB, sb, sig_b)):
            G.particle_initialization(st[0], sig); G.set_mass_p(np.float32(0.5 * nf3 / na))
            G.buffer_density(); G.buffer_x(); G.buffer_v()
        pe = GA.particle_mesh_species(GB, a_mid, dt)
        assert abs(float(pe["dt_fine"]) - float(po["dt_fine"])) > 1e-3 * float(po["dt_fine"])
    finally:
        GA.close(); GB.close(); OA.close(); OB.close()


def test_species_must_share_the_geometry(tables):
    from cafproject_b200.cube import CubeGPU, CubeGPUError
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=62)
    s2, sig2, _ = make_ic(nn=1, nc=NC, nnt=1, np_nc=NP_NC, seed=63)
    GA = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC)
    GB = CubeGPU(NC, 1, fk, ck, np_nc=NP_NC)
    try:
        GA.particle_initialization(states[0], sig); GA.buffer_density(); GA.buffer_x(); GA.buffer_v()
        GB.particle_initialization(s2[0], sig2); GB.buffer_density(); GB.buffer_x(); GB.buffer_v()
        with pytest.raises(CubeGPUError, match="geometry"):
            GA.particle_mesh_species(GB, np.float32(0.021), np.float32(0.5))
    finally:
        GA.close(); GB.close()
