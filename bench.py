#!/usr/bin/env python
"""Benchmark of the explicit structured-block update (BASELINE.json metric:
cell-updates/s in FP64 and fraction of the HBM roofline, beside the CPU path).

    python bench.py --gpus N --steps K --warmup W             (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2], the synthetic 3D 512^3 ideal-air box in
64 structured blocks of 128^3, MUSCL (l2r2 + van Albada) + AUSMDV, predictor-corrector;
the 64 blocks are divided over the N GPUs (strong scaling: total work fixed), ghost cells
between GPUs are exchanged every stage.  `--workload ffs` runs configs[1] (2D Mach-3
forward-facing step, 4096 x 1024) instead; its single-GPU number is also reported under
"also" in the default run.

One step = one full predictor-corrector time step (2 stages) of every cell.
value   = cells x K / device time of K steps, state resident in HBM (CUDA events on the
          library's stream, max over ranks).
e2e     = same metric through the C ABI with HOST buffers: eb200_upload_flow (pinned host ->
          device, encode/decode) + eb200_step + eb200_download_flow every step.
roofline= algorithmic bytes (280 B per cell-update in 3D, 224 B in 2D)
          / device time of the fused flux+update kernel, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline = the CPU oracle (restated reference, oracle/) on this box's host cores on a
          bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self, name):
        """Wall-clock stamp of the start / end of the timed region (nvidia-smi stamps its lines with local time)."""
        setattr(self, name, time.time())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        rows, smax = [], None
        t0, t1 = getattr(self, "t0", None), getattr(self, "t1", None)
        try:
            with open(self.path) as f:
                for line in f:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 10:
                        continue
                    try:
                        mhz = float(p[1])
                        smax = float(p[2])
                    except ValueError:
                        continue
                    try:
                        stamp = datetime.datetime.strptime(p[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except ValueError:
                        stamp = None
                    rs = {name for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9])
                          if v.lower().startswith("active")}
                    rows.append((stamp, mhz, rs))
            os.unlink(self.path)
        except Exception:
            pass
        # the samples taken inside the timed region (the sampler starts before the job is set up, so that it is running
        # by then); if the stamps cannot be matched, the upper half of all samples (those under load)
        inside = [r for r in rows if r[0] is not None and t0 is not None and t1 is not None and t0 - 0.02 <= r[0] <= t1 + 0.02]
        if inside:
            use, out["window"] = inside, "timed region"
        else:
            use = sorted(rows, key=lambda r: r[1])[len(rows) // 2:]
            out["window"] = "whole run (no sample stamped inside the timed region)"
        if use:
            out["sm_mhz"] = statistics.median(r[1] for r in use)
            out["sm_max_mhz"] = smax
            out["samples"] = len(use)
        out["reasons"] = sorted(set().union(*[r[2] for r in use])) if use else []
        return out


def spec_of(args, workload=None, **over):
    """A workload description: the command-line defaults with overrides (used for the `also` entries)."""
    d = {"workload": workload or args.workload, "n": args.n, "nb": args.nb, "flux": args.flux, "sheared": args.sheared,
         "dt_scale": args.dt_scale, "M_inf": None, "ffs_nx": args.ffs_nx, "ffs_ny": args.ffs_ny, "kernel": args.kernel}
    d.update(over)
    return d


def build_case(spec):
    from gdtk_b200 import cases
    w = spec["workload"]
    if w == "ffs":
        i_step = 2 * int(round(0.5 * 0.6 * spec["ffs_nx"] / 3.0))      # nearest even cell index to x = 0.6: even padded widths (TMA rows)
        cfg, gm, blocks = cases.ffs(nx=spec["ffs_nx"], ny=spec["ffs_ny"], flux_calculator=spec["flux"], i_step=i_step)
        name = (f"synthetic 2D Mach-3 forward-facing step {spec['ffs_nx']}x{spec['ffs_ny']}, 3 blocks (step face at cell {i_step}), "
                f"ideal air, l2r2+van Albada, {spec['flux']}, pc")
        balg = 224.0
    elif w == "tpg":
        cfg, gm, blocks = cases.tpg_box3d(n=spec["n"], nb=spec["nb"], flux_calculator=spec["flux"])
        name = (f"synthetic 3D {spec['n']}^3 thermally perfect 5-species air box (frozen chemistry), {spec['nb'] ** 3} blocks of "
                f"{spec['n'] // spec['nb']}^3, l2r2+van Albada, {spec['flux']}, pc")
        balg = 560.0
    else:
        extra = {"M_inf": spec["M_inf"]} if spec.get("M_inf") else {}
        cfg, gm, blocks = cases.box3d(n=spec["n"], nb=spec["nb"], flux_calculator=spec["flux"], sheared=spec["sheared"], **extra)
        name = (f"synthetic 3D {spec['n']}^3 ideal-air box, {spec['nb'] ** 3} blocks of {spec['n'] // spec['nb']}^3, "
                f"{'k-lines sheared by 10 degrees (general-metric path)' if spec['sheared'] else 'uniform Cartesian'}, "
                f"l2r2+van Albada, {spec['flux']}, pc" + (f", config.M_inf = {spec['M_inf']}" if spec.get("M_inf") else ""))
        # 280 B per cell-update; general-metric blocks read 272 B of metrics per stage on top
        balg = 824.0 if spec["sheared"] else 280.0
    cfg.force_generic_kernel = {"auto": 0, "generic": 1, "v2": 2}[spec["kernel"]]
    return cfg, gm, blocks, name, balg


def kernel_name(spec):
    """Which fused flux+update kernel the library picks for this workload (same rule as gdtk_b200/csrc/flux_inst.cu)."""
    if spec["kernel"] == "generic":
        return "flux_update_kernel (generic)"
    if spec["workload"] == "tpg":
        return "flux_update_kernel_tp (cell-centred, thermally perfect mixture, uniform Cartesian)"
    if spec["workload"] == "box3d" and spec["sheared"]:
        return "flux_update_kernel_v2 (face-centred, general metric)"
    return "flux_update_kernel_v2 (face-centred)" if spec["kernel"] == "v2" else "flux_update_kernel_v3 (cell-centred, uniform Cartesian)"


def describe(sim):
    """eb200_describe: what the library set up on this rank (blocks per kernel, tiles, TMA, halo transport)."""
    try:
        buf = C.create_string_buffer(512)
        sim.lib.describe(sim.handle, buf, 512)
        return buf.value.decode()
    except Exception:
        return None


def cfl_dt(sim, scale=1.0):
    dt_allow, _ = sim.compute_dt(False)
    dt_allow, _ = sim.reduce_dt(dt_allow, 0.0)
    return dt_allow * scale


def make_sim(spec, world, local_rank):
    from gdtk_b200 import Simulation
    from gdtk_b200.distributed import DistributedSimulation, octant_owner, distribute_blocks
    cfg, gm, blocks, name, balg = build_case(spec)
    if world > 1:
        if spec["workload"] in ("box3d", "tpg") and cfg.block_index:
            owner = octant_owner({v: next(b for b in blocks if b.id == k) for k, v in cfg.block_index.items()}, spec["nb"], world)
        else:
            owner = distribute_blocks(blocks, world)
        sim = DistributedSimulation(cfg, gm, blocks, owner, device=local_rank)
    else:
        sim = Simulation(cfg, gm, blocks, device=local_rank)
    return sim, name, balg


def run_gpu_workload(args, spec, rank, world, local_rank, with_e2e=False, with_real_loop=False, steps=None):
    import torch
    steps = steps or args.steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # long before the timed region: nvidia-smi needs a moment to come up
    t_setup = time.time()
    sim, name, balg = make_sim(spec, world, local_rank)
    t_setup = time.time() - t_setup
    lib, h = sim.lib, sim.handle
    ncells_local = int(sim.n_local_cells)
    ncells = ncells_local
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ncells_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        ncells = int(t.item())
    dt = cfl_dt(sim, spec["dt_scale"])
    ext = torch.cuda.ExternalStream(int(lib.cuda_stream(h)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up
    sim.run_fixed(args.warmup, dt)
    sim.flux_kernel_time(reset=True)
    launches0 = sim.kernel_launches()
    barrier()
    sampler.mark("t0")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    sim.run_fixed(steps, dt)
    e1.record(ext)
    e1.synchronize()
    sampler.mark("t1")
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms = e0.elapsed_time(e1)
    flux_ms, flux_n = sim.flux_kernel_time(reset=True)
    launches = sim.kernel_launches() - launches0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms, flux_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, flux_ms = float(t[0]), float(t[1])
        t = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches = int(t.item())
    value = ncells * steps / (ms * 1e-3)
    peak, peak_src = measured_peak()
    # dominant kernel: the fused flux+update kernel, one launch per stage and rank
    flux_bytes = balg * ncells_local * steps                 # algorithmic bytes this rank's launches moved
    achieved = flux_bytes / (flux_ms * 1e-3) / 1e9 if flux_ms > 0 else 0.0
    # DRAM traffic and FP64-pipe instructions of the dominant kernel per cell and launch, from the committed
    # `ncu --set full` capture of the metric configuration (profiles/r2_traffic.json) x the cells one launch processes
    traffic = None
    fp64 = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    is_metric_config = spec["workload"] == "box3d" and spec["flux"] == "ausmdv" and not spec["sheared"] and spec["kernel"] == "auto"
    if is_metric_config and os.path.exists(tpath):
        with open(tpath) as f:
            prof = json.load(f)
        traffic = prof["dram_bytes_per_cell_per_launch"] * ncells_local
        # the kernel is bound by the FP64 pipe / instruction issue, not by HBM: place it against the FP64 pipe too
        # (peak measured with profiles/micro/fp64_peak.cu on this pool's B200, thread-level DFMA/s)
        inst = prof["fp64_pipe_inst_per_cell_per_launch"] * ncells_local * steps * 2
        fp64 = {"achieved": inst / (flux_ms * 1e-3) / 1e12, "peak": prof["fp64_pipe_peak_tinst_s"],
                "unit": "T inst/s (thread-level FP64-pipe instructions)",
                "frac": inst / (flux_ms * 1e-3) / 1e12 / prof["fp64_pipe_peak_tinst_s"],
                "note": "instructions per cell from the ncu capture in profiles/ (DFMA+DMUL+DADD+DSETP); peak measured with profiles/micro/fp64_peak.cu"}
    lib_line = describe(sim) or ""
    kname = kernel_name(spec)
    if "flux_update_kernel_v3" in kname and "(0 run by flux_update_kernel_v3)" in lib_line:
        kname = "flux_update_kernel_v2 (face-centred; the cell-centred kernel needs TMA-stageable rows)"
    result = {
        "name": name, "value": value, "ms": ms, "steps": steps, "ncells": ncells, "dt": dt, "launches": launches,
        "setup_s": t_setup, "clocks": clocks, "halo": getattr(sim, "halo_transport", None), "library": lib_line,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per cell (256^3 capture, profiles/r2_traffic.json) x cells per launch; "
                                     "algorithmic = half of algorithmic_bytes_per_cell_update per launch (two stage launches per step)",
                     "peak_source": peak_src, "kernel": kname,
                     "algorithmic_bytes_per_cell_update": balg, "kernel_ms_per_launch": flux_ms / max(1, flux_n),
                     "kernel_share_of_step": flux_ms / ms if ms > 0 else None},
    }
    if fp64:
        result["fp64_pipe"] = fp64
    if with_real_loop:
        result["real_loop"] = run_real_loop(args, sim, dt, ncells, world, ext, steps)
    if with_e2e:
        result["e2e"] = run_e2e(args, sim, dt, ncells, world)
    sim.close()
    return result


def run_real_loop(args, sim, dt, ncells, world, ext, steps):
    """The loop the D shim runs (integrate_in_time): determine_time_step_size every cfl_count = 10 steps
    (eb200_compute_dt + the min over ranks) and eb200_step with its status read-back every step, instead of
    eb200_run_steps' fixed dt without host synchronisation."""
    import torch
    nbad = C.c_int(0)
    lib, h = sim.lib, sim.handle
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for n in range(steps):
        if n % 10 == 0:
            dt = min(dt, cfl_dt(sim))
        rc = lib.step(h, 0.0, dt, C.byref(nbad))
        if rc != 0:
            raise RuntimeError(f"real-loop step returned {rc}: {lib.error()}")
    e1.record(ext)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return {"value": ncells * steps / (ms * 1e-3), "unit": "cell-updates/s", "steps": steps,
            "note": "eb200_compute_dt (+ min over ranks) every 10 steps and eb200_step with status read-back every step"}


def run_e2e(args, sim, dt, ncells, world):
    """Same metric through the C ABI with host buffers: per step upload (pinned host -> device),
    step, download (device -> pinned host).  A single-species job sends the five independent FlowState variables
    (eb200_upload_flow's short form: rho, u, velocity; p, T, a follow on the device) and reads back the five conserved
    quantities (eb200_download_conserved); a multi-species job moves whole FlowStates both ways."""
    import torch
    lib, h = sim.lib, sim.handle
    nprim = sim.nprim
    short = nprim == 8
    ncq = 5 if sim.config.dimensions == 3 else 4
    host = {}
    h2d = d2h = 0

    def pinned(n, count):
        bufs = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(count)]
        return bufs, (C.POINTER(C.c_double) * count)(*[C.cast(t.data_ptr(), C.POINTER(C.c_double)) for t in bufs])

    for b in sim.local_blocks:
        g = b.geom
        n = g.NK * g.NJ * g.NI
        bufs, ptrs = pinned(n, nprim)
        lib.check(lib.download_flow(h, b.id, ptrs, nprim), "download_flow")
        if short:
            up = [bufs[0], bufs[1], bufs[5], bufs[6], bufs[7]]
            up_ptrs = (C.POINTER(C.c_double) * 5)(*[C.cast(t.data_ptr(), C.POINTER(C.c_double)) for t in up])
            down, down_ptrs = pinned(n, ncq)
            host[b.id] = (up, up_ptrs, 5, down, down_ptrs, ncq)
            h2d += n * 5 * 8
            d2h += n * ncq * 8
        else:
            host[b.id] = (bufs, ptrs, nprim, bufs, ptrs, nprim)
            h2d += n * nprim * 8
            d2h += n * nprim * 8
    nbad = C.c_int(0)
    nsteps = max(1, min(args.steps, args.e2e_steps))

    def one_step():
        for bid, (_, up_ptrs, nup, _, _, _) in host.items():
            lib.check(lib.upload_flow(h, bid, up_ptrs, nup), "upload_flow")
        rc = lib.step(h, 0.0, dt, C.byref(nbad))
        if rc != 0:
            raise RuntimeError(f"e2e step returned {rc}: {lib.error()}")
        for bid, (_, _, _, _, down_ptrs, ndown) in host.items():
            if short:
                # asynchronous: these device -> host copies run beside the next step's host -> device copies
                lib.check(lib.download_conserved_async(h, bid, down_ptrs, ndown), "download_conserved_async")
            else:
                lib.check(lib.download_flow(h, bid, down_ptrs, ndown), "download_flow")

    one_step()          # warm
    lib.check(lib.wait_downloads(h), "wait_downloads")
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        one_step()
    lib.check(lib.wait_downloads(h), "wait_downloads")        # the last step's results are in host memory
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([el], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t[0])
    if world > 1:
        h2d_t = torch.tensor([h2d, d2h], dtype=torch.int64, device="cuda")
        import torch.distributed as dist
        dist.all_reduce(h2d_t)
        h2d, d2h = int(h2d_t[0]), int(h2d_t[1])
    return {"value": ncells * nsteps / el, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": nsteps,
            "note": ("eb200_upload_flow (rho, u, velocity) + eb200_step + eb200_download_conserved_async (+ eb200_wait_downloads after the last step; the downloads of a step overlap the uploads of the next)" if short else
                     "eb200_upload_flow + eb200_step + eb200_download_flow") + " per step, pinned host buffers (bytes summed over ranks)"}


def run_parity_check(args, rank, world, local_rank):
    """N-GPU == 1-GPU, bit for bit: a small box (8 blocks of 32^3, one or more per rank) advanced 10 steps with the
    blocks spread over the ranks, against the same job with all blocks on this rank's GPU; every rank compares the
    conserved quantities of its own blocks.  Untimed; the halo exchange is a pure copy, so equality is exact."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from gdtk_b200 import Simulation
    spec = spec_of(args, "box3d", n=64, nb=2, flux=args.flux, sheared=False)
    nsteps = 10
    dsim, _, _ = make_sim(spec, world, local_rank)
    dt = cfl_dt(dsim)
    dsim.run_fixed(nsteps, dt)
    mine = {b.id: [a.copy() for a in dsim.download_conserved(b.id)] for b in dsim.local_blocks}
    dsim.close()
    cfg, gm, blocks, _, _ = build_case(spec)
    ssim = Simulation(cfg, gm, blocks, device=local_rank)
    ssim.run_fixed(nsteps, dt)
    ok = 1
    worst = 0.0
    for bid, arrs in mine.items():
        for a, r in zip(arrs, ssim.download_conserved(bid)):
            ia, ir = ssim.interior(bid, a), ssim.interior(bid, r)
            if not np.array_equal(ia, ir):
                ok = 0
                worst = max(worst, float(np.max(np.abs(ia - ir) / np.maximum(np.abs(ir), 1e-300))))
    ssim.close()
    t = torch.tensor([ok, -worst], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"case": f"3D 64^3 ideal-air box, 8 blocks of 32^3 over {world} ranks, {args.flux}, pc, {nsteps} steps at the CFL step",
            "compared": "conserved quantities of every block, multi-rank run vs all blocks on one GPU", "bit_equal": bool(t[0] > 0.5),
            "max_rel_diff": -float(t[1])}


def run_chicken(args, n=256, steps=30):
    """The reference's own CUDA solver (src/chicken, compiled for sm_100a into oracle/_ref/chkn-run by oracle/Makefile)
    on the same box: 3D ideal-air box, 2 x 2 x 2 blocks, AUSMDV, second-order van Albada reconstruction, TVD-RK3
    (three stages per step; the product's predictor-corrector has two -- compare per stage).  Timed from the wall
    clock chicken prints with its step count."""
    import re
    import shutil
    exe = os.path.join(ROOT, "oracle", "_ref", "chkn-run")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/chkn-run is not built (needs the reference sources: make -C oracle chicken)"}
    sys.path.insert(0, os.path.join(ROOT, "oracle", "chicken"))
    from make_job import make_job
    last_err = "no attempt"
    for size in (n, n // 2):
        tmp = tempfile.mkdtemp(prefix="chkn_")
        try:
            cells = make_job(os.path.join(tmp, "box"), size, max_step=steps, print_count=max(1, steps // 3))
            r = subprocess.run([exe, "--job=box", "--binary"], cwd=tmp, capture_output=True, text=True, timeout=600)
            marks = [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"Step=(\d+) .*? WC=([0-9.eE+-]+)s", r.stdout)]
            if len(marks) >= 2 and marks[-1][1] > marks[0][1]:
                (s0, w0), (s1, w1) = marks[0], marks[-1]
                per_step = cells * (s1 - s0) / (w1 - w0)
                return {"value": per_step, "unit": "cell-updates/s", "stages_per_step": 3, "cell_stage_updates_per_s": 3.0 * per_step,
                        "workload": f"chicken (reference src/chicken, -arch=sm_100a, FP64): 3D {size}^3 ideal-air box, 8 blocks of {size // 2}^3, "
                                    f"ausmdv, x_order 2, TVD-RK3, steps {s0}..{s1} by its own wall clock",
                        "n_gpus": 1}
            last_err = (r.stderr or r.stdout)[-300:].strip().replace("\n", " | ")
        except Exception as e:
            last_err = f"{type(e).__name__}: {e}"
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    return {"unavailable": f"chkn-run did not produce two progress lines: {last_err}"}


def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this rank's threads (and with them its pinned host buffers, first touch) on the CPUs
    the driver names as closest to its GPU, so that the host<->device copies of the e2e figure do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hdl, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def oracle_library():
    from gdtk_b200 import _abi
    so = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _abi.load_library(so, "orc_")


def run_cpu_sample(args, target_seconds, steps=None, warmup=1):
    """The CPU oracle (restated reference, one block per OpenMP thread like the reference's
    parallel foreach / one block per MPI rank) on a bounded sample of the 3D workload."""
    from gdtk_b200 import Simulation, cases
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    # torchrun exports OMP_NUM_THREADS=1: set the thread count explicitly, before libgomp reads the environment,
    # and again through the library; what the library then reports is what goes into the JSON line
    os.environ["OMP_NUM_THREADS"] = str(cores)
    lib = oracle_library()
    fn = lib.cdll.orc_omp_threads
    fn.restype = C.c_int
    fn.argtypes = [C.c_int]
    n, nb = args.cpu_n, args.cpu_nb
    omp_threads = int(fn(cores))
    cfg, gm, blocks = cases.box3d(n=n, nb=nb, flux_calculator=args.flux)
    sim = Simulation(cfg, gm, blocks, lib=lib)
    dt = cfl_dt(sim)
    ncells = n ** 3
    t0 = time.perf_counter()
    sim.run_fixed(max(1, warmup), dt)
    t1 = (time.perf_counter() - t0) / max(1, warmup)
    if steps is None:
        steps = int(max(1, min(50, target_seconds / max(t1, 1e-3))))
    t0 = time.perf_counter()
    sim.run_fixed(steps, dt)
    el = time.perf_counter() - t0
    sim.close()
    threads = min(omp_threads, nb ** 3)         # the oracle never uses more threads than blocks
    return {"value": ncells * steps / el, "unit": "cell-updates/s", "cores": threads, "kind": "port",
            "sample": f"{n}^3 cells in {nb ** 3} blocks of {n // nb}^3 (same job as the GPU workload at reduced size), "
                      f"{steps} predictor-corrector steps, {el:.1f} s, OpenMP over blocks: {threads} threads "
                      f"(omp_get_max_threads = {omp_threads}, {cores} host cores available to this process)",
            "ms_per_step": el / steps * 1e3, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="box3d", choices=["box3d", "ffs", "tpg"])
    ap.add_argument("--size", dest="n", type=int, default=512)
    ap.add_argument("--blocks-per-dim", dest="nb", type=int, default=4)
    ap.add_argument("--ffs-nx", type=int, default=4096)
    ap.add_argument("--ffs-ny", type=int, default=1024)
    ap.add_argument("--flux", default="ausmdv")
    ap.add_argument("--sheared", action="store_true", help="box3d on a sheared grid: every block takes the general-metric path")
    ap.add_argument("--dt-scale", type=float, default=1.0,
                    help="fraction of the CFL time step to run at (ausm_plus_up is not stable at the full CFL step on the noisy box)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "v2"],
                    help="A/B testing: force the generic fused kernel, or the face-centred tuned kernel (v2) where the "
                         "cell-centred one (v3) would run")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-n", type=int, default=128)
    ap.add_argument("--cpu-nb", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE configurations and the multi-GPU parity check")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        # The reference's CPU implementation of the path: no D toolchain exists here, so this is the
        # oracle port (restated reference) on all host cores.  Rank 0 only.
        if rank != 0:
            return
        cb = run_cpu_sample(args, 0.0, steps=max(1, args.steps), warmup=max(1, args.warmup))
        line = {
            "impl": "reference", "metric": "cell-updates/s (FP64)", "value": cb["value"], "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": cb["steps"], "warmup": max(1, args.warmup), "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 3D ideal-air box (bounded sample of the 512^3 job): " + cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # keep this process (and with it the page-locked host buffers of the e2e figure: first touch) on the CPUs next to
    # its GPU; the CPU arm below gets the full affinity mask back
    try:
        affinity0 = os.sched_getaffinity(0)
    except Exception:
        affinity0 = None
    numa = bind_near_gpu(local_rank)
    if world > 1:
        import torch.distributed as dist
        try:     # halo messages on a high-priority NCCL stream: they run beside the interior tiles
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
        except Exception:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    main_spec = spec_of(args)
    main_res = run_gpu_workload(args, main_spec, rank, world, local_rank, with_e2e=True, with_real_loop=True)
    also = {}

    def entry(r, spec):
        e = {"workload": r["name"], "value": r["value"], "unit": "cell-updates/s", "n_gpus": world,
             "ms_per_step": r["ms"] / r["steps"], "steps": r["steps"], "kernel": r["roofline"]["kernel"],
             "algorithmic_bytes_per_cell_update": r["roofline"]["algorithmic_bytes_per_cell_update"],
             "roofline_frac": r["roofline"]["frac"], "kernel_ms_per_launch": r["roofline"]["kernel_ms_per_launch"],
             "library": r.get("library")}
        if spec["dt_scale"] != 1.0:
            e["dt_scale"] = spec["dt_scale"]
        return e

    if args.workload == "box3d" and not args.no_also and args.n == 512 and not args.sheared:
        # the other BASELINE.json configurations, each a short run of the same kind
        others = []
        if world == 1:
            others.append(("ffs_4096x1024", spec_of(args, "ffs", flux="ausmdv")))
            others.append(("box3d_256_general_metric", spec_of(args, "box3d", n=256, nb=2, sheared=True, flux="ausmdv")))
            for fx in ("hanel", "ldfss0", "ldfss2", "roe", "ausm_plus_up", "adaptive_hanel_ausmdv"):
                # ausm_plus_up needs its representative Mach number: with the default config.M_inf = 0.01 the Mach-1.5 box
                # blows up within eight steps (in the oracle too); with the inflow Mach number it runs at the CFL step
                others.append((f"box3d_256_{fx}", spec_of(args, "box3d", n=256, nb=2, flux=fx, M_inf=1.5 if fx == "ausm_plus_up" else None)))
        # configs[4]: thermally perfect 5-species air 256^3, on one GPU and spread over the GPUs of the run
        others.append(("tpg_256", spec_of(args, "tpg", n=256, nb=(2 if world == 1 else 4), flux="ausmdv")))
        for key, sp in others:
            try:
                r = run_gpu_workload(args, sp, rank, world, local_rank, steps=min(args.steps, 5))
                also[key] = entry(r, sp)
            except Exception as e:       # an `also` entry must never take the headline line down with it
                also[key] = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and rank == 0 and args.workload == "box3d" and not args.no_also and args.n == 512 and not args.sheared:
        also["chicken"] = run_chicken(args)
        pc_stages = 2.0
        also["chicken"]["ours_cell_stage_updates_per_s"] = pc_stages * main_res["value"]
    parity = None
    if world > 1 and not args.no_also:
        parity = run_parity_check(args, rank, world, local_rank)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if affinity0:
            try:
                os.sched_setaffinity(0, affinity0)       # all host cores for the CPU arm
            except Exception:
                pass
        cpu = run_cpu_sample(args, args.cpu_seconds)
    if rank == 0:
        line = {
            "metric": "cell-updates/s (FP64)", "value": main_res["value"], "unit": "cell-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms"] / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": main_res["name"], "cells": main_res["ncells"], "dt": main_res["dt"],
                       "parallelism": (f"blocks over {world} GPU(s), halo exchange per stage: {main_res.get('halo') or 'exchange callback (NCCL send/recv)'}"
                                       if world > 1 else "single GPU"),
                       "timed_region": "eb200_run_steps: fixed dt, no host synchronisation between steps; "
                                       "`real_loop` is the same steps with eb200_compute_dt every 10 steps and eb200_step's status read-back",
                       "l2_policy": "state arrays (tens of GB) far exceed the 126 MB L2; no flush needed",
                       "setup_s": round(main_res["setup_s"], 1), "library": main_res.get("library")},
            "roofline": main_res["roofline"],
            "e2e": main_res.get("e2e"),
            "real_loop": main_res.get("real_loop"),
            "gpu_launches": main_res["launches"],
            "clocks": main_res["clocks"],
        }
        if main_res.get("fp64_pipe"):
            line["fp64_pipe"] = main_res["fp64_pipe"]
        if cpu:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if parity:
            line["parity_check"] = parity
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
