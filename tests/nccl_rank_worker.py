"""Worker of tests/test_gpu_nccl.py: one process per GPU (torchrun), the library's NCCL transport.
Every rank runs the whole oracle (all images) and compares its own image."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from cafproject_b200.cube import CubeGPU, image_grid
    from cafproject_b200.dist import env_rank, shared_nccl_id
    from cafproject_b200.synthetic_ic import make_ic
    from conftest import norm_rel, physical
    from oracle import cube_oracle as co

    rank, world, local = env_rank()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = os.path.join(ROOT, "tests", "golden")
    fk, ck = np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))
    nn, nc, nnt = image_grid(world), 24, 2
    states, sig, info = make_ic(nn=nn, nc=nc, nnt=nnt, np_nc=2, seed=77, disp_rms=0.8)
    O = co.Oracle(nn=nn, nnt=nnt, nc=nc, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    G = CubeGPU(nc, nnt, fk, ck, nn=nn, rank=rank, np_nc=2, device=local, tanf_lut=co.tanf_lut(), nccl_id=shared_nccl_id(device="cuda"))
    G.particle_initialization(states[rank], sig, npglobal=info["npglobal"])
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    dt_old, dt, a_mid = np.float32(0.0), np.float32(1.0), np.float32(0.021)
    uo = O.update_particle(dt_old, dt)
    ug = G.update_particle(dt_old, dt)
    so = O.store(rank); sg, _ = G.checkpoint()
    assert ug["nplocal"] == O.nplocal(rank)
    assert np.array_equal(so["rhoc"], sg["rhoc"]) and np.array_equal(so["xp"], sg["xp"]) and np.array_equal(so["vp"], sg["vp"])
    assert np.array_equal(so["vfield"].view(np.uint32), sg["vfield"].view(np.uint32))
    assert ug["sigma_vi_new"] == uo["sigma_vi_new"] and ug["overhead_tile"] == uo["overhead_tile"]
    assert O.buffer_density() == G.buffer_density()
    O.buffer_x(); G.buffer_x()
    r3o = O.coarse_density()
    fcg = G.coarse_force()
    assert norm_rel(fcg, O.force_c_image(O.coarse_force(r3o), rank)) < 1e-5
    po = O.particle_mesh(a_mid, dt)
    pg = G.particle_mesh(a_mid, dt)
    O.buffer_v(); G.buffer_v()
    for k in ("dt_fine", "dt_coarse", "dt_vmax"):
        assert abs(float(pg[k]) - float(po[k])) <= 1e-4 * abs(float(po[k])), k
    sg, _ = G.checkpoint()
    dv = np.abs(physical(O, "vp", rank).astype(np.int32) - sg["vp"].astype(np.int32))
    assert dv.max() <= 2 and (dv != 0).mean() < 2e-3
    # second step: exercises the vp ghosts received by buffer_v
    O.update_particle(dt, dt); u2 = G.update_particle(dt, dt)
    assert abs(u2["nplocal"] - O.nplocal(rank)) <= 1e-4 * O.nplocal(rank) + 2
    G.close(); O.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK world=%d grid=%s" % (world, nn))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
