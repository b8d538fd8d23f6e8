"""Host-side input generators (cafproject_b200/synthetic_ic.py): the states they return must be valid CUBE checkpoints
(cell-ordered, counts consistent) that the CPU oracle steps without complaint -- the GPU parity tests feed on them."""
import numpy as np


def _check_state(st, nc, nnt):
    nt = nc // nnt
    assert st["rhoc"].shape == (nnt,) * 3 + (nt,) * 3 and st["rhoc"].dtype == np.int32
    assert st["vfield"].shape == (nnt,) * 3 + (nt,) * 3 + (3,) and st["vfield"].dtype == np.float32
    n = int(st["rhoc"].sum(dtype=np.int64))
    assert st["xp"].shape == (n, 3) and st["vp"].shape == (n, 3)
    assert st["xp"].dtype == np.int16 and st["vp"].dtype == np.int16
    assert int(np.abs(st["vp"].astype(np.int32)).max()) <= 32767


def test_make_ic_lattice_state():
    from cafproject_b200.synthetic_ic import make_ic
    states, sig, info = make_ic(nn=(2, 1, 1), nc=16, nnt=2, np_nc=2, seed=3)
    assert len(states) == 2 and sig > 0
    for st in states:
        _check_state(st, 16, 2)
    assert sum(int(st["xp"].shape[0]) for st in states) == info["npglobal"] == 2 * 32 ** 3
    again, sig2, _ = make_ic(nn=(2, 1, 1), nc=16, nnt=2, np_nc=2, seed=3)
    assert sig2 == sig and all(np.array_equal(a[k], b[k]) for a, b in zip(states, again) for k in a)   # seeded: reproducible


def test_clustered_state_is_crowded_and_steps_through_the_oracle(tables):
    """The late-time-like state of the crowded-cell tests: single coarse cells with 10^3 particles next to empty ones."""
    from cafproject_b200.synthetic_ic import make_clustered_ic
    from oracle import cube_oracle as co
    fk, ck = tables
    nc, nnt = 16, 2
    states, sig, info = make_clustered_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=21, nblob=3, blob_sigma=0.5)
    _check_state(states[0], nc, nnt)
    assert info["rhoc_max"] > 500 and (states[0]["rhoc"] == 0).any()
    O = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    u = O.update_particle(np.float32(0.0), np.float32(1.0))
    assert O.nplocal(0) == info["npglobal"]                      # nobody lost in the re-sort
    assert np.isfinite(u["sigma_vi_new"]) and u["sigma_vi_new"] > 0
    st = O.store(0)
    assert int(st["rhoc"].sum(dtype=np.int64)) == info["npglobal"]
    O.close()


def test_tile_state_is_the_periodic_replica(tables):
    """tile_state(st, nnt, R): the R^3-fold periodic replica as ONE image with R*nnt tiles per dimension -- every big tile
    holds exactly the particles of the small tile it copies, and the oracle accepts it as a checkpoint."""
    from cafproject_b200.synthetic_ic import make_ic, tile_state
    from oracle import cube_oracle as co
    fk, ck = tables
    nc, nnt, R = 16, 2, 2          # nt = 8 >= ncb: the tile buffer does not wrap around the small image
    states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=4)
    small, big = states[0], tile_state(states[0], nnt, R)
    _check_state(big, nc * R, nnt * R)
    assert big["xp"].shape[0] == R ** 3 * small["xp"].shape[0]
    ps = small["rhoc"].reshape(nnt ** 3, -1).sum(1); ss = np.concatenate([[0], np.cumsum(ps)])
    pb = big["rhoc"].reshape((nnt * R) ** 3, -1).sum(1); sb = np.concatenate([[0], np.cumsum(pb)])
    B = nnt * R
    for tz in range(B):
        for ty in range(B):
            for tx in range(B):
                tb = (tz * B + ty) * B + tx
                tsm = ((tz % nnt) * nnt + ty % nnt) * nnt + tx % nnt
                assert np.array_equal(big["rhoc"].reshape(B ** 3, -1)[tb], small["rhoc"].reshape(nnt ** 3, -1)[tsm])
                assert np.array_equal(big["xp"][sb[tb]:sb[tb + 1]], small["xp"][ss[tsm]:ss[tsm + 1]])
                assert np.array_equal(big["vp"][sb[tb]:sb[tb + 1]], small["vp"][ss[tsm]:ss[tsm + 1]])
    # one drift of the replica = the replica of one drift (periodic box tiled: same physics)
    O1 = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    O1.load([small], sig); O1.buffer_density(); O1.buffer_x(); O1.buffer_v()
    O1.update_particle(np.float32(0.0), np.float32(1.0))
    O2 = co.Oracle(nn=1, nnt=nnt * R, nc=nc * R, np_nc=2, fk_table=fk, ck_table=ck)
    O2.load([big], sig); O2.buffer_density(); O2.buffer_x(); O2.buffer_v()
    O2.update_particle(np.float32(0.0), np.float32(1.0))
    s1, s2 = O1.store(0), O2.store(0)
    rep = tile_state(dict(xp=s1["xp"], vp=s1["vp"], rhoc=s1["rhoc"], vfield=s1["vfield"]), nnt, R)
    assert np.array_equal(rep["rhoc"], s2["rhoc"]) and np.array_equal(rep["xp"], s2["xp"])
    O1.close(); O2.close()
