"""The 1-byte zip formats through the GPU kernels (CUBE/main/universe6.fh: izipx=1 izipv=2; universe7.fh: 2,1; universe8.fh: 1,1;
parameters.f90:13-15: nvbin = 2^(8 izipv), x_resolution = 2^-(8 izipx), ishift, rshift).  Every particle kernel is instantiated
per format (cube_common.cuh::Fmt); here each format runs the same gates as x2v2 in test_gpu_parity.py against the oracle built
for that format: code tables against their formulas, densities, drift + both kicks bit for bit, one full step.
"""
import numpy as np
import pytest

from conftest import norm_rel, physical
from test_gpu_parity import _check_drift_then_kicks

pytestmark = pytest.mark.gpu

NC, NNT, NP_NC = 24, 2, 2
FORMATS = [(1, 2), (2, 1), (1, 1)]


def _pair(tables, izipx, izipv, seed=41, disp_rms=0.8):
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=seed, disp_rms=disp_rms, izipx=izipx, izipv=izipv)
    O = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=NP_NC, fk_table=fk, ck_table=ck, izipx=izipx, izipv=izipv)
    O.load(states, sig)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    G = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=co.tanf_lut(izipv), izipx=izipx, izipv=izipv)
    G.particle_initialization(states[0], sig)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    return O, G, states, sig


@pytest.mark.parametrize("izipx,izipv", FORMATS)
def test_zip_format_meshes_and_bit_exact_drift_and_kicks(tables, izipx, izipv):
    O, G, states, sig = _pair(tables, izipx, izipv)
    try:
        bad_enc, bad_dec, _ = G.selftest_codes(float(sig))
        assert bad_enc == 0 and bad_dec == 0
        back, _ = G.checkpoint()
        assert back["xp"].dtype == states[0]["xp"].dtype and back["vp"].dtype == states[0]["vp"].dtype
        assert np.array_equal(back["xp"], states[0]["xp"]) and np.array_equal(back["vp"], states[0]["vp"])
        for t in [(1, 1, 1), (2, 1, 2)]:
            ro, rg = O.fine_density(0, *t), G.fine_density(*t)
            assert norm_rel(rg[:, :, :O.nfe], ro[:, :, :O.nfe]) < 1e-6, t
            assert norm_rel(G.fine_force(*t), O.fine_force(ro)) < 1e-5, t
        assert norm_rel(G.coarse_density(), O.coarse_density()) < 1e-6
        _check_drift_then_kicks(O, G, nnt=NNT)
    finally:
        G.close(); O.close()


@pytest.mark.parametrize("izipx,izipv", FORMATS)
def test_zip_format_full_step(tables, izipx, izipv):
    O, G, _, _ = _pair(tables, izipx, izipv, seed=42, disp_rms=0.6)
    try:
        dt_old, dt, a_mid = np.float32(0), np.float32(0.7), np.float32(0.021)
        uo, po = O.step(dt_old, dt, a_mid)
        ug, pg = G.step(dt_old, dt, a_mid)
        sg, _ = G.checkpoint()
        so = O.store(0)
        assert ug["nplocal"] == O.nplocal(0)
        assert np.array_equal(so["rhoc"], sg["rhoc"])
        assert np.array_equal(physical(O, "xp"), sg["xp"])
        assert np.array_equal(so["vfield"].view(np.uint32), sg["vfield"].view(np.uint32))
        assert ug["sigma_vi_new"] == uo["sigma_vi_new"]
        dv = np.abs(physical(O, "vp").astype(np.int32) - sg["vp"].astype(np.int32))
        assert dv.max() <= 2 and (dv != 0).mean() < 5e-3
        for k in ("dt_fine", "dt_coarse", "dt_vmax"):
            assert abs(float(pg[k]) - float(po[k])) <= 1e-4 * abs(float(po[k])), k
    finally:
        G.close(); O.close()


def test_state_in_another_format_is_refused(tables):
    """particle_initialization.f90:14-18: "zip format incompatable"."""
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=1, izipx=1, izipv=2)
    G = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC)
    try:
        with pytest.raises(RuntimeError, match="zip format incompatable"):
            G.particle_initialization(states[0], sig)
    finally:
        G.close()
