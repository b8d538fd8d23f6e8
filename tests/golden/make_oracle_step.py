"""Golden vector of ONE PM step of the CPU oracle on a seeded 64^3-particle state (nc = 32, 2^3 tiles): md5 of the integer
outputs (counts, position codes, velocity codes, vfield bits) after update_particle and after the whole step, plus a few
scalars.  The reference ships no stored outputs for this path and cannot run here ("parity unpinned"), so
this file pins the ORACLE against accidental change (and, through the bit-exact GPU parity tests, the product): it is
regenerated only deliberately, with this script, in the build container.

  python tests/golden/make_oracle_step.py            # rewrites tests/golden/oracle_step_nc32.json
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NC, NNT, SEED = 32, 2, 12345
DT_OLD, DT, A_MID = 0.0, 1.0, 0.021


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def run():
    from cafproject_b200.synthetic_ic import make_ic
    from conftest import physical
    from oracle import cube_oracle as co
    fk, ck = np.load(os.path.join(HERE, "fk_table.npy")), np.load(os.path.join(HERE, "ck_table.npy"))
    states, sig, info = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=SEED, disp_rms=0.8)
    out = dict(config=dict(nc=NC, nnt=NNT, np_nc=2, seed=SEED, disp_rms=0.8, dt_old=DT_OLD, dt=DT, a_mid=A_MID),
               input=dict(xp=md5(states[0]["xp"]), vp=md5(states[0]["vp"]), rhoc=md5(states[0]["rhoc"]), vfield=md5(states[0]["vfield"]),
                          sigma_vi=float(sig), npart=int(states[0]["xp"].shape[0])))
    O = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    u = O.update_particle(np.float32(DT_OLD), np.float32(DT))
    st = O.store(0)
    out["after_update_particle"] = dict(xp=md5(st["xp"]), vp=md5(st["vp"]), rhoc=md5(st["rhoc"]), vfield=md5(st["vfield"].view(np.uint32)),
                                        nplocal=int(O.nplocal(0)), sigma_vi_new=float(u["sigma_vi_new"]), overhead_tile=float(u["overhead_tile"]),
                                        rhoc_max=int(st["rhoc"].max()))
    O.buffer_density(); O.buffer_x()
    pm = O.particle_mesh(np.float32(A_MID), np.float32(DT))
    O.buffer_v()
    # velocity codes after the kicks go through an f32 FFT (pocketfft: SIMD-path dependent in the last bit): a sum, not a hash
    out["after_particle_mesh"] = dict(xp=md5(physical(O, "xp")), vp_abs_sum=int(np.abs(physical(O, "vp").astype(np.int64)).sum()),
                                      dt_fine=float(pm["dt_fine"]), dt_coarse=float(pm["dt_coarse"]), dt_vmax=float(pm["dt_vmax"]))
    O.close()
    return out


VARIANTS = {"x1v1": dict(izipx=1, izipv=1), "x1v2": dict(izipx=1, izipv=2), "x2v1": dict(izipx=2, izipv=1),
            "cubenu_order_nlayer5": dict(vz_max=9.0)}


def run_variants():
    """The drift (update_particle) of the same seeded field in the reference's other zip formats (CUBE/main/universe6-8.fh)
    and in CUBEnu's colour-pass order (CUBEnu update_particle.f90:37,55-58): md5 of the integer outputs."""
    from cafproject_b200.synthetic_ic import make_ic
    from oracle import cube_oracle as co
    fk, ck = np.load(os.path.join(HERE, "fk_table.npy")), np.load(os.path.join(HERE, "ck_table.npy"))
    out = {}
    for name, v in VARIANTS.items():
        zx, zv = v.get("izipx", 2), v.get("izipv", 2)
        states, sig, _ = make_ic(nn=1, nc=NC, nnt=NNT, np_nc=2, seed=SEED, disp_rms=0.8, izipx=zx, izipv=zv)
        O = co.Oracle(nn=1, nnt=NNT, nc=NC, np_nc=2, fk_table=fk, ck_table=ck, izipx=zx, izipv=zv)
        O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
        u = O.update_particle(np.float32(DT_OLD), np.float32(DT), vz_max=v.get("vz_max"))
        st = O.store(0)
        out[name] = dict(input=dict(xp=md5(states[0]["xp"]), vp=md5(states[0]["vp"]), sigma_vi=float(sig)), nlayer=int(O.nlayer),
                         xp=md5(st["xp"]), vp=md5(st["vp"]), rhoc=md5(st["rhoc"]), vfield=md5(st["vfield"].view(np.uint32)),
                         nplocal=int(O.nplocal(0)), sigma_vi_new=float(u["sigma_vi_new"]))
        O.close()
    return out


if __name__ == "__main__":
    res = run()
    res["variants_after_update_particle"] = run_variants()
    json.dump(res, open(os.path.join(HERE, "oracle_step_nc32.json"), "w"), indent=1)
    print(json.dumps(res, indent=1))
