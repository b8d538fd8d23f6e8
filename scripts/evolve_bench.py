#!/usr/bin/env python
"""Evolve the bench workload from z=49 towards z=0 with the product's own step loop (cafproject_b200.run.cafcube over the
C ABI) and time the PM step along the way: the z=49 state is nearly uniform, load imbalance between cells, bricks and tiles
only appears once the particles cluster.  One JSON line per redshift window on stdout.

usage: python scripts/evolve_bench.py [--nc 256 --nnt 4 --z-end 0 --max-steps 2000 --window 25]
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
    ap.add_argument("--nc", type=int, default=256)
    ap.add_argument("--nnt", type=int, default=4)
    ap.add_argument("--z-end", type=float, default=0.0)
    ap.add_argument("--max-steps", type=int, default=2000)
    ap.add_argument("--window", type=int, default=25, help="steps per reported timing window")
    ap.add_argument("--max-seconds", type=float, default=600.0)
    ap.add_argument("--pk", action="store_true", help="matter power spectrum of the final state (cicpower/powerspectrum estimator, torch on the GPU)")
    ap.add_argument("--sweep", default="", help='";"-separated environment settings ("A=1,B=2;A=3") re-timed on the final state')
    args = ap.parse_args()
    import torch
    from cafproject_b200.cube import CubeGPU, host_tanf_lut
    from cafproject_b200.synthetic_ic import make_ic
    from cafproject_b200.timestep import Cosmology, TimeStepper

    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    fk, ck = np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))
    states, sig, info = make_ic(nn=1, nc=args.nc, nnt=args.nnt, np_nc=2, seed=2000, device="cuda")
    torch.cuda.empty_cache()
    st = states[0]
    npart = st["xp"].shape[0]
    G = CubeGPU(args.nc, args.nnt, fk, ck, np_nc=2, tanf_lut=host_tanf_lut())
    G.particle_initialization(st, sig)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    ts = TimeStepper(Cosmology(), [args.z_end])
    t_start = time.perf_counter()
    win = dict(n=0, ms=0.0, ovh=0.0, radius=0)
    nstep = 0
    while nstep < args.max_steps and time.perf_counter() - t_start < args.max_seconds:
        dt_old, dt, a_mid = ts.step()
        G.timer_start()
        up = G.update_particle(dt_old, dt)
        G.buffer_density(); G.buffer_x()
        pm = G.particle_mesh(a_mid, dt)
        G.buffer_v()
        ms = G.timer_stop()
        ts.limits(pm)
        nstep += 1
        assert up["nplocal"] == npart, "particle count changed: %d != %d" % (up["nplocal"], npart)
        win["n"] += 1; win["ms"] += ms; win["ovh"] = max(win["ovh"], float(up["overhead_tile"])); win["radius"] = max(win["radius"], G.query("drift_radius"))
        if win["n"] == args.window or ts.checkpoint_step:
            z = 1.0 / float(ts.a) - 1.0
            print(json.dumps(dict(step=nstep, z=round(z, 3), a=float(ts.a), dt=float(dt), ms_per_step=win["ms"] / win["n"],
                                  particle_updates_per_s=npart / (win["ms"] / win["n"] * 1e-3), overhead_tile=win["ovh"],
                                  drift_radius=win["radius"], sigma_vi=float(up["sigma_vi_new"]),
                                  limits=dict(fine=float(pm["dt_fine"]), coarse=float(pm["dt_coarse"]), vmax=float(pm["dt_vmax"])))), flush=True)
            win = dict(n=0, ms=0.0, ovh=0.0, radius=0)
        if ts.checkpoint_step:
            break
    # phase split of the last state (profiling brackets synchronise, so outside the timed windows)
    G.set_profiling(True)
    for _ in range(2):
        dt_old, dt, a_mid = ts.dt, ts.dt, ts.a_mid
        G.update_particle(dt_old, dt); G.buffer_density(); G.buffer_x(); G.particle_mesh(a_mid, dt); G.buffer_v()
    ph = {k: v / 2 for k, v in G.phase_times().items() if v > 0}
    final_state, final_sig = G.checkpoint()
    rc = final_state["rhoc"]
    edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 1 << 30]
    flat = rc.reshape(-1).astype(np.int64)
    hist = {("%d-%d" % (lo, hi - 1)): dict(cells=int(((flat >= lo) & (flat < hi)).sum()), particles=int(flat[(flat >= lo) & (flat < hi)].sum()))
            for lo, hi in zip(edges[:-1], edges[1:])}
    print(json.dumps(dict(rhoc_histogram=hist)), flush=True)
    print(json.dumps(dict(final=True, steps=nstep, z=1.0 / float(ts.a) - 1.0, wall_s=time.perf_counter() - t_start, phases_ms=ph,
                          rhoc_max=int(rc.max()), rhoc_mean=float(rc.mean()), empty_cell_fraction=float((rc == 0).mean()), nparticles=int(npart))), flush=True)
    xi = None
    if args.pk:   # cicpower + powerspectrum of the resident state, on the device with the library's own kernels
        t0 = time.perf_counter()
        xi = G.power_spectrum(200.0)
        pk_s = time.perf_counter() - t0
    G.close()
    if args.pk:
        sel = [i for i in range(xi.shape[1]) if xi[0][i] > 0][:: max(1, xi.shape[1] // 24)]
        print(json.dumps(dict(power_spectrum_seconds=pk_s)), flush=True)
        print(json.dumps(dict(power_spectrum=dict(k_h_per_Mpc=[float(xi[1][i]) for i in sel], Delta2=[float(xi[2][i]) for i in sel],
                                                  ng=4 * args.nc))), flush=True)
    # the same final state under other settings of the library's environment switches (read at init)
    dts = (ts.dt, ts.a_mid)
    for setting in [x for x in args.sweep.split(";") if x]:
        for kv in setting.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
        G = CubeGPU(args.nc, args.nnt, fk, ck, np_nc=2, tanf_lut=host_tanf_lut())
        G.particle_initialization(final_state, final_sig)
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        G.set_profiling(True)
        ph = None
        for it in range(3):
            G.update_particle(dts[0], dts[0]); G.buffer_density(); G.buffer_x(); G.particle_mesh(dts[1], dts[0]); G.buffer_v()
            if it == 0:
                G.phase_times()    # discard the first (warm-up) step
        ph = {k: v / 2 for k, v in G.phase_times().items() if v > 0}
        print(json.dumps(dict(sweep=setting, ms_per_step=sum(ph.values()),
                              phases_ms={k: round(v, 3) for k, v in ph.items()})), flush=True)
        G.close()


if __name__ == "__main__":
    main()
