"""Timeline of the end-to-end step of bench.py (upload -> buffers -> update_particle -> streamed checkpoint under particle_mesh):
runs two e2e steps under the CUPTI activity tracer that torch.profiler drives (there is no nsys in the image) and prints, per CUDA
stream, when it was busy, plus every memcpy with start and duration.  Numbers taken under the tracer are for reading the overlap,
never bench values.

    python scripts/e2e_trace.py [--nc 256 --nnt 4] [--out gpurun_out/e2e_trace.json]
    torchrun --nproc-per-node 8 ... scripts/e2e_trace.py --mode step     # resident steps of an 8-image run, traced on rank 0
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nc", type=int, default=256)
    ap.add_argument("--nnt", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_trace.json"))
    ap.add_argument("--no-stream-vp", action="store_true")
    ap.add_argument("--mode", choices=("e2e", "step"), default="e2e")
    ap.add_argument("--min-us", type=float, default=300.0, help="print kernels and gaps longer than this")
    args = ap.parse_args()
    rank, world, local_rank = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    from cafproject_b200.cube import CubeGPU, host_tanf_lut, image_grid
    from cafproject_b200.dist import shared_nccl_id
    from cafproject_b200.synthetic_ic import make_ic
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = os.path.join(ROOT, "tests", "golden")
    fk, ck = np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))
    states, sig, _ = make_ic(nn=1, nc=args.nc, nnt=args.nnt, np_nc=2, seed=2000, device="cuda")
    torch.cuda.empty_cache()
    npart = states[0]["xp"].shape[0]
    G = CubeGPU(args.nc, args.nnt, fk, ck, nn=image_grid(world), rank=rank, np_nc=2, device=local_rank, tanf_lut=host_tanf_lut(),
                nccl_id=shared_nccl_id(device="cuda") if world > 1 else None)
    G.particle_initialization(states[0], sig, npglobal=world * npart)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    dt, a_mid = np.float32(0.5), np.float32(0.021)
    G.step(np.float32(0), dt, a_mid)
    for _ in range(3):
        G.step(dt, dt, a_mid)
    cur, sig_cur = G.checkpoint()
    n0 = cur["xp"].shape[0]
    pin = {k: (torch.empty((n0 + 1024, 3), dtype=torch.int16).pin_memory() if k in ("xp", "vp") else torch.from_numpy(v).pin_memory())
           for k, v in cur.items()}
    host = {k: v.numpy() for k, v in pin.items()}
    host["xp"][:n0] = cur["xp"]; host["vp"][:n0] = cur["vp"]
    inp = dict(host, xp=host["xp"][:n0], vp=host["vp"][:n0])

    def step(inp, sig_cur):
        if args.mode == "step":
            G.step(dt, dt, a_mid)
            return inp, sig_cur
        G.particle_initialization(inp, sig_cur, npglobal=world * npart)
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        G.update_particle(dt, dt)
        G.checkpoint_begin(host, xp=True, cells=True, vp_during_pm=not args.no_stream_vp)
        G.buffer_density(); G.buffer_x()
        G.particle_mesh(a_mid, dt)
        G.buffer_v()
        return G.checkpoint(out=host, skip=("xp", "rhoc", "vfield") + (() if args.no_stream_vp else ("vp",)))

    inp, sig_cur = step(inp, sig_cur)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            inp, sig_cur = step(inp, sig_cur)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        prof.export_chrome_trace(args.out)
    G.close()
    if rank != 0:
        return
    ev = [e for e in json.load(open(args.out))["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    print("# %d device activities over %.2f ms (2 steps)" % (len(ev), (ev[-1]["ts"] + ev[-1]["dur"] - t0) / 1e3))
    print("# memcpy > 0.2 ms: start_ms dur_ms stream name")
    for e in ev:
        if e["cat"] == "gpu_memcpy" and e["dur"] > min(200, args.min_us):
            print("copy %9.2f %8.2f  s%-3s %s %s" % ((e["ts"] - t0) / 1e3, e["dur"] / 1e3, e["args"].get("stream"), e["name"][:24], e["args"].get("bytes", "")))
    print("# kernels > 0.3 ms and every gap > 0.3 ms on their stream: start_ms dur_ms stream name")
    last = {}
    for e in ev:
        if e["cat"] != "kernel":
            continue
        s = e["args"].get("stream")
        if s in last and e["ts"] - last[s] > args.min_us:
            print("gap  %9.2f %8.2f  s%-3s" % ((last[s] - t0) / 1e3, (e["ts"] - last[s]) / 1e3, s))
        last[s] = e["ts"] + e["dur"]
        if e["dur"] > args.min_us:
            print("kern %9.2f %8.2f  s%-3s %s" % ((e["ts"] - t0) / 1e3, e["dur"] / 1e3, s, e["name"][:60]))


if __name__ == "__main__":
    main()
