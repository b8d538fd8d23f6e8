"""Every instantiation of the hand-written line-FFT kernels (kPlans of cafproject_b200/csrc/cube_gpu.cu) against the oracle.

The fine-mesh force is a circular convolution on a window of length N >= nft + 32; any such N gives the reference's force_f
(cube_fft.cuh).  CUBE_GPU_NFFT forces the window, so each <R1,R2> kernel set (its dft<R1>, dft<R2> butterflies, shared-memory
pitches and launch-bound branches) runs on a 12^3-cell tile the oracle convolves in a fraction of a second: kernel table,
force of two tiles at the 1e-5 gate, and a full step's time-step limits.  The plans a run picks by itself are additionally
covered at their own tile size in test_gpu_parity.py (nt = 12, 16, 24 -> N = 80, 96, 128) and test_gpu_bench_tile.py
(nt = 64 -> 288).
"""
import numpy as np
import pytest

from conftest import norm_rel

pytestmark = pytest.mark.gpu

NC, NNT, NP_NC = 24, 2, 2
PLANS = [80, 96, 128, 160, 192, 256, 288, 360, 480, 576]      # R1*R2 of every kPlans entry


@pytest.fixture(scope="module")
def oracle_side(tables):
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=NP_NC, seed=31, disp_rms=0.8)
    O = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=NP_NC, fk_table=fk, ck_table=ck)
    O.load(states, sig)
    O.buffer_density(); O.buffer_x(); O.buffer_v()
    forces = {t: O.fine_force(O.fine_density(0, *t)) for t in [(1, 1, 1), (2, 1, 2)]}
    pm = O.particle_mesh(np.float32(0.021), np.float32(0.7))
    yield states, sig, forces, pm
    O.close()


@pytest.mark.parametrize("n", PLANS)
def test_every_fft_plan_gives_the_reference_force(n, oracle_side, tables, monkeypatch):
    from cafproject_b200.cube import CubeGPU
    from oracle import cube_oracle as co
    states, sig, forces, pm = oracle_side
    fk, ck = tables
    monkeypatch.setenv("CUBE_GPU_NFFT", str(n))
    G = CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC, tanf_lut=co.tanf_lut())
    try:
        assert G.query("nfft") == n
        assert norm_rel(G.kern_f(), co.kernel_f(fk, n)) < 1e-5
        G.particle_initialization(states[0], sig)
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        for t, fo in forces.items():
            fg = G.fine_force(*t)
            assert norm_rel(fg, fo) < 1e-5, (n, t)
            for d in range(3):
                assert norm_rel(fg[..., d], fo[..., d]) < 1e-5, (n, t, d)
        # the step's own route (prefix folded into the Green multiply, f2_max in the x-inverse epilogue)
        pg = G.particle_mesh(np.float32(0.021), np.float32(0.7))
        assert abs(float(pg["dt_fine"]) - float(pm["dt_fine"])) <= 1e-4 * float(pm["dt_fine"])
    finally:
        G.close()


def test_forced_window_below_the_minimum_is_refused(tables, monkeypatch):
    from cafproject_b200.cube import CubeGPU, CubeGPUError
    fk, ck = tables
    monkeypatch.setenv("CUBE_GPU_NFFT", "77")
    with pytest.raises(CubeGPUError, match="not a built transform length"):
        CubeGPU(NC, NNT, fk, ck, np_nc=NP_NC)
