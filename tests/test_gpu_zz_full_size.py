"""Size-independent properties of the GPU step at the bench's full size (BASELINE.json configs[1]: 512^3 particles in one image,
nc = 256, 64 tiles of nt = 64) -- a size the CPU oracle does not finish in seconds, so parity proper is left to the small cases
of test_gpu_parity.py and this file checks what must hold at any size: the upload/download round trip, conservation of the
particle number through drift and re-sort (update_particle.f90:205-211), the mass on the coarse mesh (CUBEnu pm.f90:267), the
time-step limits, and run-to-run determinism bit for bit (no floating-point atomics anywhere on the path).  Named to sort
after the other GPU tests.

Status: written at the end of round 1 after the round's GPU minutes were spent -- every call in it is one bench.py or the
small parity tests already make on a B200, but this file itself has not run on hardware yet.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NC, NNT, NP_NC = 256, 4, 2


def _run(G, st, sig, dt, a_mid):
    G.particle_initialization(st, sig)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    up, pm = G.step(dt, dt, a_mid)
    out, sig_out = G.checkpoint()
    return up, pm, {k: np.array(v, copy=True) for k, v in out.items()}, sig_out


def test_full_size_step_properties(tables):
    import torch
    from cafproject_b200.cube import CubeGPU, host_tanf_lut
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, info = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=2000, device="cuda")
    torch.cuda.empty_cache()
    st = states[0]
    n = st["xp"].shape[0]
    assert n == (NP_NC * NC) ** 3 == int(st["rhoc"].sum(dtype=np.int64))
    G = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=host_tanf_lut())
    try:
        # upload -> buffered state -> download gives the state back (particle_initialization / checkpoint round trip)
        G.particle_initialization(st, sig)
        ovh = G.buffer_density(); G.buffer_x(); G.buffer_v()
        assert 0 < float(ovh) <= 1
        back, sig_back = G.checkpoint()
        assert sig_back == np.float32(sig)
        for k in ("xp", "vp", "rhoc", "vfield"):
            assert np.array_equal(back[k], st[k]), k
        del back
        # all the mass is on the coarse mesh: sum(r3) = N*mass_p = nf_global^3
        r3 = G.coarse_density()
        total = float((4 * NC) ** 3)
        assert abs(float(r3.sum(dtype=np.float64)) - total) < 1e-5 * total
        assert float(r3.min()) >= 0
        del r3
        dt, a_mid = np.float32(0.5), np.float32(0.0205)
        up1, pm1, s1, sig1 = _run(G, st, sig, dt, a_mid)
        # drift + re-sort keep every particle; counts are a partition of them
        assert up1["nplocal"] == n == s1["xp"].shape[0]
        assert int(s1["rhoc"].sum(dtype=np.int64)) == n and int(s1["rhoc"].min()) >= 0
        assert 0 < float(up1["overhead_tile"]) <= 1
        assert float(up1["sigma_vi_new"]) > 0 and np.isfinite(s1["vfield"]).all()
        assert not np.array_equal(s1["xp"], st["xp"])                        # the particles did move
        for k in ("dt_fine", "dt_coarse", "dt_vmax"):
            assert np.isfinite(float(pm1[k])) and float(pm1[k]) > 0, k
        # the same step from the same state again: identical bits
        up2, pm2, s2, sig2 = _run(G, st, sig, dt, a_mid)
        for k in ("xp", "vp", "rhoc", "vfield"):
            assert np.array_equal(s1[k].view(np.uint8), s2[k].view(np.uint8)), k
        assert sig1 == sig2 and up1["sigma_vi_new"] == up2["sigma_vi_new"]
        for k in ("dt_fine", "dt_coarse", "dt_vmax", "vmax"):
            assert pm1[k] == pm2[k], k
    finally:
        G.close()
