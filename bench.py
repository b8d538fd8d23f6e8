#!/usr/bin/env python
"""bench.py -- particle-updates/s of one CUBE PM step (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--nc NC --nnt NNT]

A "step" is one pass of the hot path over the resident state (cafcube.f90:27-31):
update_particle -> buffer_density -> buffer_x -> particle_mesh -> buffer_v.  Workload at N=1 is
BASELINE.json configs[1]: CUBE LCDM 512^3 particles, 1 image (nc=256, nnt=4 => nt=64, nfe=304, np_nc=2),
synthetic Zel'dovich LCDM-shaped initial conditions at z=49.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with the state resident in HBM;
`e2e` = same metric through the C ABI with the disjoint state in pinned host memory, upload and download
inside the timed region; `roofline` = the dominant phase of the step against MEASURED_PEAKS.json;
`cpu_baseline` = the CPU oracle (restated reference path, kind "port") on a bounded sample.
`--impl reference` times that CPU port alone, with all host threads, on the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-updates/sec per PM step"
UNIT = "particle-updates/s"


def tables():
    g = os.path.join(ROOT, "tests", "golden")
    return np.load(os.path.join(g, "fk_table.npy")), np.load(os.path.join(g, "ck_table.npy"))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), sampled in-process through NVML
    (nvidia_ml_py) every 20 ms so that even a sub-second timed region gets samples; falls back to one nvidia-smi query."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device=0):
        self.device, self.rows, self.stop_flag, self.thread, self.nv = device, [], False, None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device])
            except Exception:
                pass
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.hd = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.hd, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.hd, nv.NVML_CLOCK_SM))
        try:
            rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.hd))
        except Exception:
            rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.hd))
        try:
            pw = nv.nvmlDeviceGetPowerUsage(self.hd) / 1000.0
        except Exception:
            pw = None
        self.rows.append((sm, rs, pw))

    def _loop(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=20).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["sampled once after the timed region (NVML unavailable)"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join()
        if not self.rows:
            try:
                self._sample()
            except Exception:
                pass
        reasons = set()
        for _, rs, _ in self.rows:
            for name, bit in self.BITS.items():
                if rs & bit:
                    reasons.add(name)
        sm = [r[0] for r in self.rows]
        pw = [r[2] for r in self.rows if r[2] is not None]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": float(np.median(pw)) if pw else None}


# ---------------------------------------------------------------------------------------------
# CPU port (the oracle): bounded sample, one independent single-tile image per host thread --
# the way the reference is deployed (one coarray image per core, CUBE/main/run.sh)
# ---------------------------------------------------------------------------------------------
def cpu_port_throughput(nt, steps=1, warmup=0, max_threads=32):
    from oracle import cube_oracle as co
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables()
    ncores = os.cpu_count() or 1
    nth = max(1, min(ncores, max_threads))
    os.environ["CUBE_ORACLE_THREADS"] = "1"
    states, sig, _ = make_ic(nn=1, nc=nt, nnt=1, np_nc=2, seed=7)
    npart = states[0]["xp"].shape[0]
    kf = co.kernel_f(fk, 4 * nt + 48)
    kc = co.kernel_c(ck, nt)
    sims = []
    for _ in range(nth):
        O = co.Oracle(nn=1, nnt=1, nc=nt, np_nc=2)
        O.kern_f, O.kern_c = kf, kc
        O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
        sims.append(O)
    dt_old, dt, a_mid = np.float32(0.0), np.float32(1.0), np.float32(0.021)

    def run(O, n):
        for _ in range(n):
            O.step(dt_old, dt, a_mid)

    def timed(n):
        th = [threading.Thread(target=run, args=(O, n)) for O in sims]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        return time.perf_counter() - t0

    if warmup:
        timed(warmup)
    sec = timed(steps)
    for O in sims:
        O.close()
    value = nth * npart * steps / sec
    sample = ("%d independent single-tile images (nc=nt=%d, nfe=%d, %d particles each = 1/%s of the GPU workload's tiles), "
              "one per host thread, %d step(s)" % (nth, nt, 4 * nt + 48, npart, "64", steps))
    return dict(value=value, unit=UNIT, cores=nth, kind="port", sample=sample, seconds=sec, ms_per_step=1e3 * sec / steps)


def cpu_port_cfg1_image(steps=1):
    """BASELINE.json configs[0] as CUBE/main runs it: ONE image (nc=128, nnt=2: 8 tiles of nt=64, 256^3 particles), its particle
    loops serial (CUBE/main has no OpenMP; one coarray image = one core), the FFTs through pocketfft with every host core.  Phase
    seconds at the boundaries CUBEnu prints (update_particle | buffers | particle_mesh)."""
    from oracle import cube_oracle as co
    from cafproject_b200.synthetic_ic import make_ic
    fk, ck = tables()
    os.environ.pop("CUBE_ORACLE_THREADS", None)
    states, sig, _ = make_ic(nn=1, nc=128, nnt=2, np_nc=2, seed=1000)
    npart = states[0]["xp"].shape[0]
    O = co.Oracle(nn=1, nnt=2, nc=128, np_nc=2, fk_table=fk, ck_table=ck)
    O.load(states, sig); O.buffer_density(); O.buffer_x(); O.buffer_v()
    dt_old, dt, a_mid = np.float32(0.0), np.float32(1.0), np.float32(0.021)
    ph = {"update_particle": 0.0, "buffer": 0.0, "particle_mesh": 0.0}
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter(); O.update_particle(dt_old, dt); t1 = time.perf_counter()
        O.buffer_density(); O.buffer_x(); t2 = time.perf_counter()
        O.particle_mesh(a_mid, dt); t3 = time.perf_counter()
        O.buffer_v(); t4 = time.perf_counter()
        ph["update_particle"] += t1 - t0; ph["buffer"] += (t2 - t1) + (t4 - t3); ph["particle_mesh"] += t3 - t2
        dt_old = dt
    sec = time.perf_counter() - t_all
    O.close()
    return dict(value=npart * steps / sec, unit=UNIT, cores=1, fft_workers=os.cpu_count() or 1, kind="port",
                sample="%d step(s) of ONE cfg-1 image (nc=128 nnt=2 nt=64 nfe=304, %d particles), serial particle loops as in CUBE/main, "
                       "pocketfft on all host cores" % (steps, npart),
                seconds=sec, phase_seconds={k: v / steps for k, v in ph.items()})


def run_reference(args, rank, world):
    if rank != 0:
        return
    nt = args.nc // args.nnt
    r = cpu_port_throughput(nt, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    cfg = workload_config(args, world)
    # what this arm ran (the GPU arm's workload is named beside it: the metric is a throughput, comparable across the two)
    cfg = dict(cfg, workload="CPU port, the reference's deployment of one coarray image per core: " + r["sample"]
               + "; GPU arm's workload: " + cfg["workload"])
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 mesh / f64 particle update / int16 codes", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "restated CPU path (C + pocketfft port of CUBE/main), not the coarray-Fortran binary: no Fortran compiler exists here"}
    if not args.no_cfg1:
        try:   # config 1 as one image, with phase timers and the core count
            line["cfg1_single_image"] = cpu_port_cfg1_image(1)
        except Exception as e:   # never lose the arm's line over the extra sample
            line["cfg1_single_image"] = {"error": str(e)}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    nt = args.nc // args.nnt
    two = getattr(args, "species", 1) == 2
    return {"workload": "CUBE LCDM %d^3 particles per image%s, %d image(s), nc=%d nnt=%d nt=%d nfe=%d np_nc=2 izipx=izipv=2, z=49 Zel'dovich ICs"
                        % (2 * args.nc, " + as many of a second (hot, light) species on the same meshes (CUBEnu -DNEUTRINOS)" if two else "",
                           world, args.nc, args.nnt, nt, 4 * nt + 48),
            "step": "update_particle+buffer_density+buffer_x+particle_mesh+buffer_v (cafcube.f90:27-31)",
            "l2": "state and meshes (>1.5 GB) exceed the 126 MB L2; no flush needed",
            "parallelism": "1 image" if world == 1 else
                           "%d images on a %dx%dx%d image grid, one per GPU; ghost-cell/particle exchange, distributed coarse FFT and scalar "
                           "reductions over NCCL (global box = the tiling of one image's ICs)" % ((world,) + tuple(_grid(world)))}


def _grid(world):
    from cafproject_b200.cube import image_grid
    return image_grid(world)


def bind_near_gpu(device):
    """Pin this process to the CPUs NVML names as closest to its GPU, before any pinned host memory is allocated: the e2e leg moves
    3.8 GB per step and rank over PCIe, and with 8 ranks an allocation on the far socket crosses the inter-socket link twice.
    Returns what was done, for the JSON line."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        hd = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(device).uuid)).encode())
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(hd)
        return "nvml: %d of %d cpus" % (len(os.sched_getaffinity(0)), before)
    except Exception as e:  # noqa: BLE001
        return "unchanged (%s)" % type(e).__name__


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--nc", type=int, default=256)
    ap.add_argument("--no-stream-vp", action="store_true", help="e2e leg: download the velocities after particle_mesh instead of streaming them per tile batch")
    ap.add_argument("--ic-tile", type=int, default=1, help="build the image by replicating an (nc/R)^3 image R times per dimension")
    ap.add_argument("--nnt", type=int, default=4)
    ap.add_argument("--fine-batch", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stream-upload", action="store_true", help="e2e leg: cube_gpu_upload instead of cube_gpu_upload_begin")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the CPUs next to its GPU")
    ap.add_argument("--species", type=int, default=1, help="2: BASELINE.json configs[3], a second (hot, light) species of as many particles on the same meshes")
    ap.add_argument("--no-cfg1", action="store_true", help="reference arm: skip the one-step sample of a whole cfg-1 image")
    ap.add_argument("--no-late", action="store_true", help="skip the late-time leg (evolve the ICs to z=0 and time the clustered state)")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    # library chatter (e.g. NCCL's version banner) must not share stdout with the one JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from cafproject_b200.cube import CubeGPU, host_tanf_lut, image_grid
    from cafproject_b200.dist import shared_nccl_id
    from cafproject_b200.synthetic_ic import make_ic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    affinity = bind_near_gpu(local_rank) if not args.no_bind else "unchanged (--no-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fk, ck = tables()
    nc, nnt = args.nc, args.nnt
    # every image starts from the same periodic ICs: the global box is their tiling (continuous across image boundaries,
    # identical work per GPU = weak scaling)
    if args.ic_tile > 1:    # HBM-sized images: a smaller periodic image replicated (the field generator's temporaries would not fit)
        from cafproject_b200.synthetic_ic import tile_state
        states, sig, info = make_ic(nn=1, nc=nc // args.ic_tile, nnt=nnt // args.ic_tile, np_nc=2, seed=2000, device="cuda")
        states = [tile_state(states[0], nnt // args.ic_tile, args.ic_tile)]
    else:
        states, sig, info = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=2000, device="cuda")
    torch.cuda.empty_cache()
    st = states[0]
    npart = st["xp"].shape[0]
    nccl_id = shared_nccl_id(device="cuda") if world > 1 else None
    G = CubeGPU(nc, nnt, fk, ck, nn=image_grid(world), rank=rank, np_nc=2, device=local_rank, fine_batch=args.fine_batch,
                tanf_lut=host_tanf_lut(), nccl_id=nccl_id)
    G.particle_initialization(st, sig, npglobal=world * npart)
    G.buffer_density(); G.buffer_x(); G.buffer_v()
    # fixed small time step so that every timed step does the same work (dt from the first PM limits)
    dt, a_mid = np.float32(0.5), np.float32(0.0205)
    G2 = None
    if args.species == 2:   # cfg 4: CDM + a hot light species (5 % of the mass, 3x the velocity dispersion), each in its own handle
        s2, sig2, _ = make_ic(nn=1, nc=nc, nnt=nnt, np_nc=2, seed=4000, device="cuda", velocity_boost=3.0)
        torch.cuda.empty_cache()
        nccl_id2 = shared_nccl_id(device="cuda") if world > 1 else None
        G2 = CubeGPU(nc, nnt, fk, ck, nn=image_grid(world), rank=rank, np_nc=2, device=local_rank, tanf_lut=host_tanf_lut(), nccl_id=nccl_id2, secondary=True)
        G2.particle_initialization(s2[0], sig2, npglobal=world * npart)
        mass = float((4 * nc) ** 3) / npart
        G.set_mass_p(0.95 * mass); G2.set_mass_p(0.05 * mass)
        G2.buffer_density(); G2.buffer_x(); G2.buffer_v()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(dt_old):
        if G2 is None:
            G.step(dt_old, dt, a_mid)
            return
        for H in (G, G2):                    # CUBEnu main.f90:102-118 with NEUTRINOS: every call for both species
            H.update_particle(dt_old, dt)
        for H in (G, G2):
            H.buffer_density(); H.buffer_x()
        G.particle_mesh_species(G2, a_mid, dt)
        for H in (G, G2):
            H.buffer_v()

    one_step(np.float32(0.0))
    for _ in range(max(0, args.warmup - 1)):
        one_step(dt)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = G.query("kernel_launches")
    barrier()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStart()
    G.timer_start()
    for _ in range(args.steps):
        one_step(dt)
    ms = G.timer_stop()
    if args.profile:
        torch.cuda.cudart().cudaProfilerStop()
    barrier()
    launches = G.query("kernel_launches") - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms_per_step = ms / args.steps
    nspec = 2 if G2 is not None else 1
    value = nspec * world * npart / (ms_per_step * 1e-3)

    # ---- per-phase timing (CUDA events on the library stream) -> dominant kernel + roofline ----
    G.set_profiling(True); G.phase_times()
    if G2 is not None:
        G2.set_profiling(True); G2.phase_times()
    nprof = 2
    for _ in range(nprof):
        one_step(dt)
    phases = {k: v / nprof for k, v in G.phase_times().items()}
    G.set_profiling(False)
    if G2 is not None:   # the second species' own drift and exchange phases (its mesh phases are in the first handle's)
        for k, v in G2.phase_times().items():
            if v > 0:
                phases[k + "_species2"] = v / nprof
        G2.set_profiling(False)
    nt = nc // nnt; nfe = 4 * nt + 48
    ntile = nnt ** 3
    batch = G.query("fine_batch")
    nfine = ntile * nfe ** 3
    # compulsory bytes per step of each phase: every input read once, every output written once (DESIGN.md "roofline")
    N = G.query("nfft"); NH = N // 2 + 1; M = 4 * nt + 2
    reg = (4 * nc + N - 4 * nt) ** 3 if batch == ntile else ntile * N ** 3   # nodes of the fine-density region(s) written per step
    np_reg = npart * ((nc + 10) / nc) ** 3      # particles deposited once: the image plus five ghost cells each side
    alg = {"drift_key": 18 * npart + 28 * nc ** 3, "drift_count": 8 * npart, "drift_place": 28 * npart, "drift_scan": 12 * nc ** 3,
           "buffer": 32 * nc ** 3, "fine_deposit": 6 * np_reg + 4 * reg,
           "fine_fft_x": ntile * (4 * N ** 3 + 8 * N * N * NH), "fine_fft_y": ntile * 16 * N * N * NH,
           "fine_fft_z_green": ntile * (8 * N * N * NH + 24 * M * N * NH) + 12 * N * N * NH,
           "fine_ifft_y": ntile * 3 * (8 * M * N * NH + 8 * M * M * NH), "fine_ifft_x": ntile * 3 * (8 * M * M * NH + 4 * M ** 3),
           "fine_f2max": ntile * 12 * M ** 3, "fine_fft_xy": ntile * (4 * N ** 3 + 8 * N * N * NH),
           "fine_ifft_yx_f2max": ntile * 3 * (8 * M * N * NH + 4 * M ** 3), "fine_kick": 18 * npart + 12 * ntile * M ** 3,
           "coarse_deposit": 6 * npart + 4 * nc ** 3, "coarse_fft_green": 50 * nc ** 3, "coarse_kick": 18 * npart + 12 * nc ** 3}
    phases = {k: v for k, v in phases.items() if v > 0}
    launches_per_step = {k: (ntile + batch - 1) // batch if k.startswith("fine") else 1 for k in phases}
    for k in list(phases):
        alg.setdefault(k, alg.get(k.replace("_species2", ""), 0))
    dom = max(phases, key=lambda k: phases[k])
    peak, which = measured_peak()
    dom_ms_per_launch = phases[dom] / launches_per_step[dom]
    achieved = alg[dom] / launches_per_step[dom] / (dom_ms_per_launch * 1e-3) / 1e9
    # DRAM traffic of the dominant phase per launch, from the committed ncu capture of this workload (profiles/traffic_cfg2.json,
    # made by scripts/traffic_summary.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); null for other workloads
    traffic, traffic_src = None, None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_cfg2.json")
    if (nc, nnt, world) == (256, 4, 1) and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if dom in tj["phases"]:
            traffic = tj["phases"][dom]["dram_bytes"] / launches_per_step[dom]
            traffic_src = "profiles/traffic_cfg2.json (" + tj["source"] + ")"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "ms_per_launch": dom_ms_per_launch,
                "algorithmic_bytes_per_launch": alg[dom] / launches_per_step[dom]}
    bytes_step = 48 * npart + 42 * nfine + 82 * nc ** 3
    step_roof = {"bytes_step": bytes_step, "achieved": bytes_step / (ms_per_step * 1e-3) / 1e9,
                 "frac": bytes_step / (ms_per_step * 1e-3) / 1e9 / peak, "formula": "48*Np+42*Nfine_ext+82*Ncoarse (BASELINE.md sec.2)"}

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    e2e = None
    if G2 is not None:
        G2.close()
    if not args.no_e2e and G2 is None:
        cur, sig_cur = G.checkpoint()
        cap = int(1.25 * cur["xp"].shape[0]) + 1024      # nplocal of an image changes from step to step when nn > 1
        pin = {k: (torch.empty((cap, 3), dtype=torch.int16).pin_memory() if k in ("xp", "vp") else torch.from_numpy(v).pin_memory())
               for k, v in cur.items()}
        host = {k: v.numpy() for k, v in pin.items()}
        n0 = cur["xp"].shape[0]
        host["xp"][:n0] = cur["xp"]; host["vp"][:n0] = cur["vp"]
        inp = dict(host, xp=host["xp"][:n0], vp=host["vp"][:n0])
        n_e2e = max(3, min(args.steps, 5))
        h2d = d2h = 0

        def e2e_step(inp, sig_cur):
            G.particle_initialization(inp, sig_cur, npglobal=world * npart, streamed=not args.no_stream_upload)
            G.buffer_density(); G.buffer_x(); G.buffer_v()
            G.update_particle(dt, dt)      # keys each chunk of a streamed upload as it lands
            # positions, rhoc and vfield are final for this step: stream them out under particle_mesh; the velocities follow
            # tile batch by tile batch as their kicks are done
            G.checkpoint_begin(host, xp=True, cells=True, vp_during_pm=not args.no_stream_vp)
            G.buffer_density(); G.buffer_x()
            G.particle_mesh(a_mid, dt)
            G.buffer_v()
            # the result lands in the same pinned buffers = the next step's input
            return G.checkpoint(out=host, skip=("xp", "rhoc", "vfield") + (() if args.no_stream_vp else ("vp",)))

        inp, sig_cur = e2e_step(inp, sig_cur)   # warm-up: the streamed-checkpoint path's first use (kernel modules, its density buffer)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            h2d = sum(v.nbytes for v in inp.values())
            inp, sig_cur = e2e_step(inp, sig_cur)
            d2h = sum(v.nbytes for v in inp.values())
        barrier()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); sec = float(t.item())
        e2e = {"value": world * npart * n_e2e / sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * sec / n_e2e, "steps": n_e2e, "cpu_affinity": affinity}
        # one more step, untimed for the metric, with a device synchronisation after every call: where the wall clock of an e2e step goes
        marks = []

        def mark(name, t_prev):
            torch.cuda.synchronize()
            t = time.perf_counter()
            marks.append((name, 1e3 * (t - t_prev)))
            return t
        barrier()
        t = time.perf_counter()
        G.particle_initialization(inp, sig_cur, npglobal=world * npart); t = mark("upload", t)
        G.buffer_density(); G.buffer_x(); G.buffer_v(); t = mark("buffer", t)
        G.update_particle(dt, dt); t = mark("update_particle", t)
        G.checkpoint_begin(host, xp=True, cells=True, vp_during_pm=not args.no_stream_vp)
        G.buffer_density(); G.buffer_x(); t = mark("buffer_after_drift+xp_cells_download_so_far", t)
        G.particle_mesh(a_mid, dt); t = mark("particle_mesh+streamed_downloads", t)
        G.buffer_v(); t = mark("buffer_v", t)
        inp, sig_cur = G.checkpoint(out=host, skip=("xp", "rhoc", "vfield") + (() if args.no_stream_vp else ("vp",))); t = mark("checkpoint_tail", t)
        e2e["serialised_breakdown_ms"] = {k: round(v, 2) for k, v in marks}
        e2e["h2d_GBps_upload_alone"] = round(h2d / 1e9 / (marks[0][1] * 1e-3), 1)
    radius = G.query("drift_radius")
    G.close()

    # ---- late-time (clustered) state: the same ICs evolved to z=0 by the product's own step loop, then timed ----------------
    late = None
    if not args.no_late and args.species == 1:
        from cafproject_b200.timestep import Cosmology, TimeStepper
        G = CubeGPU(nc, nnt, fk, ck, nn=image_grid(world), rank=rank, np_nc=2, device=local_rank, fine_batch=args.fine_batch, tanf_lut=host_tanf_lut(),
                    nccl_id=shared_nccl_id(device="cuda") if world > 1 else None)
        G.particle_initialization(st, sig, npglobal=world * npart)
        G.buffer_density(); G.buffer_x(); G.buffer_v()
        ts = TimeStepper(Cosmology(), [0.0])
        t0 = time.perf_counter()
        nev = 0
        while nev < 400 and (world > 1 or time.perf_counter() - t0 < 120.0):   # (every image must take the same number of steps)
            dto, dtn, am = ts.step()
            G.update_particle(dto, dtn); G.buffer_density(); G.buffer_x()
            ts.limits(G.particle_mesh(am, dtn)); G.buffer_v()
            nev += 1
            if ts.checkpoint_step:
                break
        dtl, aml = ts.dt, ts.a_mid
        for _ in range(2):
            G.update_particle(dtl, dtl); G.buffer_density(); G.buffer_x(); G.particle_mesh(aml, dtl); G.buffer_v()
        torch.cuda.synchronize()
        G.timer_start()
        nl = 3
        for _ in range(nl):
            G.update_particle(dtl, dtl); G.buffer_density(); G.buffer_x(); G.particle_mesh(aml, dtl); G.buffer_v()
        msl = G.timer_stop() / nl
        if world > 1:
            t = torch.tensor([msl], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); msl = float(t.item())
        G.set_profiling(True); G.phase_times()
        G.update_particle(dtl, dtl); G.buffer_density(); G.buffer_x(); G.particle_mesh(aml, dtl); G.buffer_v()
        phl = {k: v for k, v in G.phase_times().items() if v > 0}
        late = {"z": 1.0 / float(ts.a) - 1.0, "steps_evolved": nev, "ms_per_step": msl, "value": world * npart / (msl * 1e-3), "unit": UNIT,
                "drift_radius": G.query("drift_radius"), "phases_ms_per_step": phl,
                "note": "the bench ICs evolved from z=49 by the adaptive step loop (cafcube.f90:25-46), then timed: haloes, empty cells"}
        G.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_port_throughput(nt, steps=1, warmup=0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 mesh / f64 particle update / int16 codes", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "step_roofline": step_roof, "phases_ms_per_step": phases, "cpu_baseline": cpu,
                "particles_per_gpu": int(npart), "drift_radius": radius, "late_time": late}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
