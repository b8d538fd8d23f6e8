#!/usr/bin/env python
"""The z = 0 P(k) gate (within 0.1 % for k < k_Nyquist/2) at a size beyond the unit tests: the GPU library and the
CPU oracle evolve the same z = 49 initial conditions to z = 0 through their own step loops (cafcube.f90:25-46); both final states
go through the SAME estimator -- cube_gpu_power_spectrum on the device (cicpower.f90 + powerspectrum.f90) -- and the ratio of the
two spectra is printed per shell.  Test infrastructure (runs the oracle); not part of the product path.

usage: python scripts/pk_gate.py [--nc 64 --nnt 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, default=64)
    ap.add_argument("--nnt", type=int, default=2)
    args = ap.parse_args()
    from cafproject_b200.cube import CubeGPU
    from cafproject_b200.run import cafcube
    from cafproject_b200.synthetic_ic import make_ic
    from cafproject_b200.timestep import Cosmology, TimeStepper
    from oracle import cube_oracle as co
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    fk, ck = np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))
    nc, nnt = args.nc, args.nnt
    states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=49, disp_rms=0.5)
    # ---- oracle
    t0 = time.perf_counter()
    O = co.Oracle(nn=1, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    ts = co.TimeStepper(co.Cosmology(), [0.0])
    nso = 0
    while True:
        dt_old, dt, a_mid = ts.step()
        _, pm = O.step(dt_old, dt, a_mid)
        ts.dt_fine, ts.dt_coarse, ts.dt_vmax = pm["dt_fine"], pm["dt_coarse"], pm["dt_vmax"]
        nso += 1
        if ts.checkpoint_step:
            O.update_particle(np.float32(0), ts.dt)
            break
    final_o = {k: np.array(v, copy=True) for k, v in O.store(0).items()}
    sig_o = O.sigma_vi
    O.close()
    t_or = time.perf_counter() - t0
    # ---- GPU
    t0 = time.perf_counter()
    G = CubeGPU(nc, nnt, fk, ck, np_nc=2, tanf_lut=co.tanf_lut())
    G.particle_initialization(states[0], sig); G.buffer_density(); G.buffer_x(); G.buffer_v()
    got = {}
    nsg = cafcube(G, TimeStepper(Cosmology(), [0.0]), on_checkpoint=lambda z, st, s: got.update(z=z, state=st, sig=s))
    t_gpu = time.perf_counter() - t0

    def spectrum(state, s):
        G.particle_initialization(state, s); G.buffer_density(); G.buffer_x(); G.buffer_v()
        return G.power_spectrum(200.0)

    xg, xo, xi = spectrum(got["state"], got["sig"]), spectrum(final_o, sig_o), spectrum(states[0], sig)
    G.close()
    nyq = 2 * nc
    k = np.arange(1, xg.shape[1] + 1)
    low = (k < nyq / 2) & (xo[0] > 0)
    ratio = xg[2][low] / xo[2][low]
    growth = xo[2][low] / xi[2][low]
    same_cells = float((got["state"]["rhoc"] == final_o["rhoc"]).mean())
    out = dict(nc=nc, nnt=nnt, nparticles=int(info["npglobal"]), steps_oracle=nso, steps_gpu=int(nsg), seconds_oracle=t_or, seconds_gpu=t_gpu,
               shells=int(low.sum()), max_abs_ratio_minus_1=float(np.abs(ratio - 1).max()), gate="0.1 % for k < k_Nyquist/2",
               passed=bool(np.abs(ratio - 1).max() < 1e-3), growth_min=float(growth.min()), growth_max=float(growth.max()),
               fraction_of_cells_with_equal_counts=same_cells,
               k_h_per_Mpc=[float(v) for v in xo[1][low][:: max(1, int(low.sum()) // 16)]],
               ratio=[float(v) for v in ratio[:: max(1, int(low.sum()) // 16)]])
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
