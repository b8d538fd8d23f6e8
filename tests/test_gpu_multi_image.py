"""More than one image on ONE GPU: every image is a host thread of this process
(``local_group``), ghost exchange / distributed coarse FFT / scalar reductions go through the same code as the
NCCL path, only the transport differs (cube_comm.cuh).  Oracle = all images in one process
(oracle/cube_oracle.c), image grids 2x1x1, 2x2x1 and the reference's own 2x2x2.

Parity is unpinned by the reference (no golden vectors upstream); gates as in test_gpu_parity.py.
"""
import threading

import numpy as np
import pytest

from conftest import norm_rel, physical

pytestmark = pytest.mark.gpu

NP_NC = 2
_group = [100]


def run_images(nimg, fn):
    """fn(m) on one thread per image; re-raises the first failure."""
    out, err = [None] * nimg, [None] * nimg

    def work(m):
        try:
            out[m] = fn(m)
        except BaseException as e:  # noqa: BLE001
            err[m] = e

    th = [threading.Thread(target=work, args=(m,)) for m in range(nimg)]
    for t in th: t.start()
    for t in th: t.join()
    for e in err:
        if e is not None:
            raise e
    return out


class Run:
    def __init__(self, tables, nn, nc, nnt, seed, disp_rms=0.8):
        from cafproject_b200.cube import CubeGPU
        from cafproject_b200.synthetic_ic import make_ic
        from oracle import cube_oracle as co
        fk, ck = tables
        self.nn, self.nimg = nn, nn[0] * nn[1] * nn[2]
        self.states, self.sig, info = make_ic(nn=nn, nc=nc, nnt=nnt, np_nc=NP_NC, seed=seed, disp_rms=disp_rms)
        self.npglobal = info["npglobal"]
        self.O = co.Oracle(nn=nn, nnt=nnt, nc=nc, np_nc=NP_NC, fk_table=fk, ck_table=ck)
        self.O.load(self.states, self.sig)
        self.O.buffer_density(); self.O.buffer_x(); self.O.buffer_v()
        _group[0] += 1
        grp = _group[0]
        lut = co.tanf_lut()

        def mk(m):
            G = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=m, np_nc=NP_NC, tanf_lut=lut, local_group=grp, fine_batch=2)
            G.particle_initialization(self.states[m], self.sig, npglobal=self.npglobal)
            G.buffer_density(); G.buffer_x(); G.buffer_v()
            return G
        self.G = run_images(self.nimg, mk)

    def each(self, fn):
        return run_images(self.nimg, lambda m: fn(m, self.G[m]))

    def close(self):
        self.each(lambda m, G: G.close())
        self.O.close()


@pytest.mark.parametrize("nn,nc,nnt", [((2, 1, 1), 24, 2), ((2, 2, 1), 24, 1), ((2, 2, 2), 24, 2)])
def test_drift_bit_exact_and_meshes(tables, nn, nc, nnt):
    """buffer_density/x/v + update_particle across images: every code bit-exact; then densities, forces and both
    kicks (with the oracle's forces) image by image."""
    R = Run(tables, nn, nc, nnt, seed=21 + nn[1] + nn[2])
    O = R.O
    try:
        dt_old, dt, a_mid = np.float32(0.0), np.float32(1.0), np.float32(0.021)
        uo = O.update_particle(dt_old, dt)
        ug = R.each(lambda m, G: G.update_particle(dt_old, dt))
        for m in range(R.nimg):
            so = O.store(m)
            sg, _ = R.G[m].checkpoint()
            assert ug[m]["nplocal"] == O.nplocal(m)
            assert np.array_equal(so["rhoc"], sg["rhoc"]), m
            assert np.array_equal(so["xp"], sg["xp"]), m
            assert np.array_equal(so["vfield"].view(np.uint32), sg["vfield"].view(np.uint32)), m
            assert np.array_equal(so["vp"], sg["vp"]), m
            assert ug[m]["sigma_vi_new"] == uo["sigma_vi_new"]
            assert ug[m]["overhead_tile"] == uo["overhead_tile"]
            for k in ("std_vsim", "std_vsim_c", "std_vsim_res"):
                assert abs(ug[m][k] - uo[k]) <= 1e-12 * abs(uo[k])
        assert sum(u["nplocal"] for u in ug) == R.npglobal
        ovo = O.buffer_density(); O.buffer_x()
        ovg = R.each(lambda m, G: (G.buffer_density(), G.buffer_x())[0])
        assert all(v == ovo for v in ovg)
        # fine density of a tile that touches an image boundary, coarse density, coarse force
        t = (1, 1, 1)
        for m in range(R.nimg):
            ro = O.fine_density(m, *t)
            rg = R.G[m].fine_density(*t)
            assert norm_rel(rg[:, :, :O.nfe], ro[:, :, :O.nfe]) < 1e-6, m
            assert not np.any(rg[:, :, :O.nfe][ro[:, :, :O.nfe] == 0]) and np.all(rg[:, :, :O.nfe][ro[:, :, :O.nfe] > 1e-4] > 0)
        r3o = O.coarse_density()
        r3g = R.each(lambda m, G: G.coarse_density())
        for m in range(R.nimg):
            ix, iy, iz = O.image_coords(m)
            blk = r3o[iz * nc:(iz + 1) * nc, iy * nc:(iy + 1) * nc, ix * nc:(ix + 1) * nc]
            assert norm_rel(r3g[m], blk) < 1e-6, m
        fco = O.coarse_force(r3o)
        fcg = R.each(lambda m, G: G.coarse_force())
        for m in range(R.nimg):
            assert norm_rel(fcg[m], O.force_c_image(fco, m)) < 1e-5, m
        # kicks with the oracle's forces: velocity codes bit-exact
        sig_old, sig_new = O.sigma_vi, O.sigma_vi_new
        pm = O.particle_mesh(a_mid, dt, keep=True)

        def kicks(m, G):
            f2 = []
            for tz in range(1, nnt + 1):
                for ty in range(1, nnt + 1):
                    for tx in range(1, nnt + 1):
                        f2.append(G.fine_kick_with(tx, ty, tz, pm["meshes"]["force_f"][(m, tx, ty, tz)], a_mid, dt, sig_old, sig_new))
            vmax, f2c = G.coarse_kick_with(O.force_c_image(pm["meshes"]["force_c"], m), a_mid, dt, sig_new)
            return max(f2), vmax, f2c
        kg = R.each(kicks)
        assert np.float32(max(k[0] for k in kg)) == pm["f2_max_fine"]
        for m in range(R.nimg):
            assert kg[m][1] == pm["vmax"][m]
            assert kg[m][2] == pm["f2_max_coarse"][m]
            sg, _ = R.G[m].checkpoint()
            assert np.array_equal(physical(O, "xp", m), sg["xp"])
            assert np.array_equal(physical(O, "vp", m), sg["vp"])
    finally:
        R.close()


def test_full_steps_two_images(tables):
    """Three full steps on 2x1x1 images, each side with its own FFTs (test_gpu_parity.py::test_full_steps rules)."""
    from oracle import cube_oracle as co
    nn, nc, nnt = (2, 1, 1), 24, 2
    R = Run(tables, nn, nc, nnt, seed=33, disp_rms=0.6)
    O = R.O
    try:
        ts = co.TimeStepper(co.Cosmology(), [0.0])
        for it in range(3):
            dt_old, dt, a_mid = ts.step()
            uo, po = O.step(dt_old, dt, a_mid)
            res = R.each(lambda m, G: G.step(dt_old, dt, a_mid))
            for m in range(R.nimg):
                ug, pg = res[m]
                sg, _ = R.G[m].checkpoint()
                xp_o, vp_o = physical(O, "xp", m), physical(O, "vp", m)
                same_cells = np.array_equal(O.store(m)["rhoc"], sg["rhoc"])
                if it == 0:
                    assert same_cells and ug["nplocal"] == O.nplocal(m)
                    assert np.array_equal(xp_o, sg["xp"])
                if same_cells:
                    dv = np.abs(vp_o.astype(np.int32) - sg["vp"].astype(np.int32))
                    assert dv.max() <= (2 if it == 0 else 4)
                    assert (dv != 0).mean() < (2e-3 if it == 0 else 2e-2)
                else:
                    assert int(np.abs(O.store(m)["rhoc"] - sg["rhoc"]).sum()) < 1e-4 * xp_o.shape[0]
                for k in ("dt_fine", "dt_coarse", "dt_vmax"):
                    assert abs(float(pg[k]) - float(po[k])) <= 1e-4 * abs(float(po[k])), k
            ts.dt_fine, ts.dt_coarse, ts.dt_vmax = po["dt_fine"], po["dt_coarse"], po["dt_vmax"]
    finally:
        R.close()


@pytest.mark.parametrize("nn,nc,nnt", [((2, 1, 1), 24, 2), ((2, 2, 2), 24, 1)])
def test_particle_ids_cross_image_boundaries(tables, nn, nc, nnt):
    """-DPID with several images: the IDs travel with vp in buffer_v (buffer_v.f90:23,42,62,81,104) and take the permutation of
    update_particle (update_particle.f90:88) -- after two drifts every image holds the IDs the oracle holds, slot for slot, and
    all images together still hold every ID exactly once (a large step makes many particles change image)."""
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    fk, ck = tables
    nimg = nn[0] * nn[1] * nn[2]
    states, sig, info = make_ic(nn=nn, nc=nc, nnt=nnt, np_nc=NP_NC, seed=90, disp_rms=1.5)
    base = 0
    for s in states:
        n = s["xp"].shape[0]
        s["pid"] = np.arange(base + 1, base + n + 1, dtype=np.int64)
        base += n
    O = co.Oracle(nn=nn, nnt=nnt, nc=nc, np_nc=NP_NC, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    _group[0] += 1
    grp = _group[0]
    lut = co.tanf_lut()

    def mk(m):
        G = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=m, np_nc=NP_NC, tanf_lut=lut, local_group=grp, fine_batch=2)
        G.particle_initialization(states[m], sig, npglobal=info["npglobal"])
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        return G
    Gs = run_images(nimg, mk)
    try:
        dt = np.float32(1.5)
        moved = 0
        for it in range(2):
            O.update_particle(np.float32(0) if it == 0 else dt, dt)
            run_images(nimg, lambda m: Gs[m].update_particle(np.float32(0) if it == 0 else dt, dt))
            allids = []
            for m in range(nimg):
                so = O.store(m)
                sg, _ = Gs[m].checkpoint()
                assert np.array_equal(so["xp"], sg["xp"]) and np.array_equal(so["vp"], sg["vp"]), (it, m)
                assert np.array_equal(so["pid"], sg["pid"]), (it, m)
                allids.append(sg["pid"])
                lo = sum(s["xp"].shape[0] for s in states[:m])
                moved += int(((sg["pid"] <= lo) | (sg["pid"] > lo + states[m]["xp"].shape[0])).sum())
            assert np.array_equal(np.sort(np.concatenate(allids)), np.arange(1, info["npglobal"] + 1))
            O.buffer_density(); O.buffer_x(); O.buffer_v()
            run_images(nimg, lambda m: (Gs[m].buffer_density(), Gs[m].buffer_x(), Gs[m].buffer_v()))
        assert moved > 0          # the test did move particles (and their IDs) from image to image
    finally:
        run_images(nimg, lambda m: Gs[m].close())
        O.close()


def test_two_species_on_two_images(tables):
    """cube_gpu_particle_mesh_species with several images (BASELINE cfg 4's shape): every image holds both species, each species
    has its own ghost exchange, the coarse density of both goes through one distributed transform.  By construction as in
    test_gpu_two_species.py: the particles of a one-species 2x1x1 state dealt alternately to two species reproduce the
    one-species kick (codes within the rare one-unit flips) and time-step limits; then both species drift and nobody is lost."""
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    from test_gpu_two_species import _split
    fk, ck = tables
    nn, nc, nnt = (2, 1, 1), 24, 2
    nimg = 2
    states, sig, info = make_ic(nn=nn, nc=nc, nnt=nnt, np_nc=NP_NC, seed=77, disp_rms=0.8)
    npg = info["npglobal"]
    lut = co.tanf_lut()
    a_mid, dt = np.float32(0.021), np.float32(0.8)
    mass_p = float((4 * nc) ** 3 * nimg) / npg
    _group[0] += 3
    g1, ga, gb = _group[0] - 2, _group[0] - 1, _group[0]

    def one(m):
        G = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=m, np_nc=NP_NC, tanf_lut=lut, local_group=g1, fine_batch=2)
        try:
            G.particle_initialization(states[m], sig, npglobal=npg)
            G.buffer_density(); G.buffer_x(); G.buffer_v()
            pm = G.particle_mesh(a_mid, dt)
            st, _ = G.checkpoint()
            return pm, {k: np.array(v, copy=True) for k, v in st.items()}
        finally:
            G.close()
    ref = run_images(nimg, one)
    halves = [_split(states[m]) for m in range(nimg)]
    npa = sum(h[0][1]["xp"].shape[0] for h in halves); npb = sum(h[1][1]["xp"].shape[0] for h in halves)

    def two(m):
        GA = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=m, np_nc=NP_NC, tanf_lut=lut, local_group=ga, fine_batch=2)
        GB = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=m, np_nc=NP_NC, tanf_lut=lut, local_group=gb, fine_batch=2, secondary=True)
        try:
            for Gs, (mask, s), npgs in ((GA, halves[m][0], npa), (GB, halves[m][1], npb)):
                Gs.particle_initialization(s, sig, npglobal=npgs)
                Gs.set_mass_p(mass_p)
                Gs.buffer_density(); Gs.buffer_x(); Gs.buffer_v()
            pm = GA.particle_mesh_species(GB, a_mid, dt)
            for Gs in (GA, GB):
                Gs.buffer_v()
            ca, _ = GA.checkpoint(); cb, _ = GB.checkpoint()
            ca = {k: np.array(v, copy=True) for k, v in ca.items()}; cb = {k: np.array(v, copy=True) for k, v in cb.items()}
            for Gs in (GA, GB):
                Gs.buffer_density(); Gs.buffer_x(); Gs.buffer_v()
            ua = GA.update_particle(dt, dt); ub = GB.update_particle(dt, dt)
            return pm, ca, cb, ua["nplocal"], ub["nplocal"]
        finally:
            GA.close(); GB.close()
    got = run_images(nimg, two)
    for m in range(nimg):
        pm1, st1 = ref[m]
        pm2, ca, cb, _, _ = got[m]
        for k in ("dt_fine", "dt_coarse"):
            assert abs(float(pm2[k]) - float(pm1[k])) <= 1e-5 * float(pm1[k]), (m, k)
        for (mask, _), c in ((halves[m][0], ca), (halves[m][1], cb)):
            assert np.array_equal(c["xp"], st1["xp"][mask])
            dv = np.abs(c["vp"].astype(np.int32) - st1["vp"][mask].astype(np.int32))
            assert dv.max() <= 2 and (dv != 0).mean() < 1e-3, (m, dv.max(), (dv != 0).mean())
    assert sum(g[3] for g in got) == npa and sum(g[4] for g in got) == npb
