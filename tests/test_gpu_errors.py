"""Error behaviour of the C ABI = the reference's `print*` + `stop` conditions, reported instead of silently truncated:
update_particle.f90:61-67 (tile capacity), buffer_density.f90:87-93 / particle_initialization (image capacity), and a particle
that outruns the tile buffer (the reference would index out of bounds)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NC, NNT = 32, 2


def _loaded(tables, states, sig, **kw):
    from cafproject_b200.cube import CubeGPU
    fk, ck = tables
    G = CubeGPU(NC, NNT, fk, ck, np_nc=2, **kw)
    G.particle_initialization(states[0], sig)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    return G


def test_tile_capacity_overflow_is_reported(tables):
    """Half of the particles in three clumps: one tile + buffer holds far more than np_tile_max at tile_buffer = 1."""
    from cafproject_b200.synthetic_ic import make_clustered_ic
    states, sig, _ = make_clustered_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=5, nblob=1, blob_fraction=0.9, blob_sigma=0.5)
    G = _loaded(tables, states, sig, tile_buffer=0.15)
    with pytest.raises(RuntimeError, match="please set tile_buffer larger"):
        G.update_particle(np.float32(0.0), np.float32(0.5))
    G.close()


def test_image_capacity_overflow_is_reported(tables):
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=6)
    G = CubeGPU(NC, NNT, fk, ck, np_nc=2, image_buffer=0.1)
    with pytest.raises(RuntimeError, match="please set image_buffer larger"):
        G.particle_initialization(states[0], sig)
    G.close()


def test_step_beyond_the_tile_buffer_is_reported(tables):
    """dt so large that particles move more than ncb = 6 coarse cells: outside what the ghost layers cover."""
    from cafproject_b200.synthetic_ic import make_ic
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=7, disp_rms=0.8, velocity_boost=50.0)
    G = _loaded(tables, states, sig)
    with pytest.raises(RuntimeError, match="outside the tile buffer"):
        G.update_particle(np.float32(0.0), np.float32(4000.0))
    G.close()


def test_particle_mesh_needs_buffered_state(tables):
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables
    states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=8)
    G = CubeGPU(NC, NNT, fk, ck, np_nc=2)
    G.particle_initialization(states[0], sig)
    with pytest.raises(RuntimeError, match="not buffered"):
        G.particle_mesh(np.float32(0.02), np.float32(0.5))
    G.close()
