"""The step loop of ``program cafcube`` (CUBE/main/cafcube.f90:16-46) over the C ABI: what the Fortran driver does once
the five hot-path calls are replaced by ``libcubegpu.so``.  One image per caller; in a
multi-image run every image executes this loop (the library's calls are collective where the reference has ``sync all``).
"""
from __future__ import annotations

import numpy as np

from .timestep import Cosmology, TimeStepper


def cafcube(G, ts: TimeStepper, on_checkpoint=None, istep_max=100000, log=None):
    """Run ``G`` (a :class:`cafproject_b200.cube.CubeGPU` whose state is already buffered, cafcube.f90:16-20) until the
    last redshift of ``ts.z_checkpoint``.  ``on_checkpoint(z, state, sigma_vi)`` gets the disjoint state that
    ``checkpoint`` would write (checkpoint.f90:33-70).  Returns the number of steps taken."""
    for _ in range(istep_max):
        dt_old, dt, a_mid = ts.step()                       # call timestep
        up = G.update_particle(dt_old, dt)                  # call update_particle
        G.buffer_density(); G.buffer_x()                    # call buffer_density ; call buffer_x
        pm = G.particle_mesh(a_mid, dt)                     # call particle_mesh
        G.buffer_v()                                        # call buffer_v
        ts.limits(pm)
        if log:
            log(ts, up, pm)
        if ts.checkpoint_step:                              # cafcube.f90:32-43
            G.update_particle(np.float32(0), ts.dt)         # dt_old=0 ; call update_particle (half drift)
            state, sig = G.checkpoint()
            if on_checkpoint:
                on_checkpoint(float(ts.z_checkpoint[ts.cur_checkpoint]), state, sig)
            if ts.final_step:
                break
            G.buffer_density(); G.buffer_x(); G.buffer_v()
            ts.after_checkpoint()
    return ts.istep
