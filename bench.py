#!/usr/bin/env python
"""bench.py -- cell-stage updates/s of the MUSCL residual + RK time-march hot path on B200.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W
prints ONE JSON line on rank 0.  A "step" is one full time step of the named workload:
CFL reduction + num_stages x (fused stage kernel + ghost/BC refresh [+ NCCL halo exchange]).

Workload (BASELINE.json configs[4], the configuration the metric is quoted on): synthetic
weak-scaling explosion -- 8 blocks of 2048x2048 cells per GPU (8 wide x N high for N GPUs), fp64,
Roe + Venkatakrishnan + Green-Gauss, conservative reconstruction, RK4, CFL 0.7, reflection BCs.

`--impl reference` times the CPU restatement of the reference (oracle/, "port") on the host
cores for the same metric on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

GAMMA = 1.4
A_INF = 343.0
BLOCK_LEN = 1.25  # length units per block side


# ---------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# ---------------------------------------------------------------------------------------------
def ws_mesh(n_gpus, blocks_per_gpu):
    from pyhype_b200.mesh.rectangular import RectagularMeshGenerator

    return RectagularMeshGenerator.generate(
        BCE=["Reflection"], BCW=["Reflection"], BCN=["Reflection"], BCS=["Reflection"],
        east=BLOCK_LEN * blocks_per_gpu, west=0.0, north=BLOCK_LEN * n_gpus, south=0.0,
        n_blocks_horizontal=blocks_per_gpu, n_blocks_vertical=n_gpus,
    ).dict


def ws_ic(x, y, width, height):
    """Explosion box over the central 40 % of the domain (explosion_multi states), conservative,
    non-dimensional (examples/explosion/initial_condition.py:35-60)."""
    inside = (x >= 0.3 * width) & (x <= 0.7 * width) & (y >= 0.3 * height) & (y <= 0.7 * height)

    def cons(rho, p):
        e = p / (GAMMA - 1) + 0.0
        return np.array([rho / 1.0, 0.0, 0.0, e / (1.0 * A_INF**2)])

    hi, lo = cons(4.6968, 404400.0), cons(1.1742, 101100.0)
    return np.where(inside[..., None], hi, lo)


def ws_ic_smooth(x, y, width, height):
    """Rounding-robust smooth field (IC-B): every face carries a genuine Riemann problem
    (diagnostic runs with --ic smooth; the headline workload is the explosion box)."""
    rho = 1.2 + 0.3 * np.sin(0.7 * x + 0.3) * np.cos(0.45 * y + 0.1)
    u = 30 * np.cos(0.5 * x) * np.sin(0.35 * y + 0.2)
    v = -25 * np.sin(0.4 * x + 0.5) * np.cos(0.3 * y)
    p = 101325 * (1 + 0.2 * np.cos(0.6 * x - 0.2) * np.sin(0.5 * y + 0.4))
    ek = 0.5 * rho * (u * u + v * v)
    U = np.stack((rho, rho * u, rho * v, p / (GAMMA - 1) + ek), axis=-1)
    return U / np.array([1.0, A_INF, A_INF, A_INF**2])


BYTES_PER_CELL_STEP = {"RK4": 512.0, "RK2": 160.0, "ExplicitEuler1": 64.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU, so that the page-locked host buffers of
    the end-to-end leg are first-touched on the GPU's NUMA node (PCIe copies from the far socket run at about half
    rate).  Returns the previous affinity (restored before the CPU baseline runs) or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
            return prev
    except Exception:
        pass
    return None


def run_ours(args):
    import torch

    from pyhype_b200.distributed import HaloExchanger, advance, distribute_blocks, world
    from pyhype_b200.engine import Engine, SIDES
    from pyhype_b200.mesh.quad_mesh import QuadMesh
    from pyhype_b200.time_marching import get_tableau

    rank, wsize, lrank = world()
    n_gpus = args.gpus
    if wsize != n_gpus:
        if wsize == 1 and n_gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
        n_gpus = wsize
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(lrank)
    prev_affinity = bind_to_gpu_numa_node(lrank)
    dist = None
    if wsize > 1:
        import torch.distributed as dist

        from pyhype_b200.distributed import init_nccl

        init_nccl(lrank)

    nb, n = args.blocks_per_gpu, args.block
    integ = args.integrator
    tab = get_tableau(integ)
    nstages = len(tab)
    blocks = ws_mesh(n_gpus, nb)
    owner = distribute_blocks(len(blocks), wsize)
    mine = sorted(g for g, r in owner.items() if r == rank)
    width, height = BLOCK_LEN * nb, BLOCK_LEN * n_gpus

    eng = Engine(n, n, args.flux, "Venkatakrishnan", args.recon, tab, GAMMA, 0.7, device=lrank)
    host_states = {}
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(mine)))) as pool:   # host geometry: numpy-bound, releases the GIL
        meshes = dict(zip(mine, pool.map(lambda g: QuadMesh(n, n, NE=blocks[g]["NE"], NW=blocks[g]["NW"], SE=blocks[g]["SE"], SW=blocks[g]["SW"]), mine)))
    for gid in mine:
        b = blocks[gid]
        m = meshes.pop(gid)
        eng.add_block(gid, m, {s: b["Neighbor" + s] for s in SIDES}, {s: b["BCType" + s] for s in SIDES}, local_gids=set(mine))
        U = (ws_ic_smooth if args.ic == "smooth" else ws_ic)(m.x[:, :, 0], m.y[:, :, 0], width, height)
        pinned = torch.empty((n, n, 4), dtype=torch.float64, pin_memory=True)
        pinned.numpy()[...] = U
        host_states[gid] = pinned
        del m, U
    eng.finalize()
    stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", lrank))
    hx = HaloExchanger(eng, owner, rank) if wsize > 1 else None

    def upload_all():
        for gid in mine:
            eng.upload(gid, host_states[gid].numpy())

    def refresh_ghosts():
        if hx is not None:
            hx.exchange()
        eng.apply_bc()

    def one_step():
        """get_dt + integrate (Euler2D._solve body, pyhype/solvers/Euler2D.py:199-204), dt stays on the device"""
        if hx is not None:
            ptr = hx.global_dt().data_ptr()
        else:
            eng.local_dt(dt_dev.data_ptr())
            ptr = dt_dev.data_ptr()
        advance(eng, hx, nstages, dt_dev_ptr=ptr)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        dt_dev = torch.zeros(1, dtype=torch.float64, device=f"cuda:{lrank}")
        upload_all()
        refresh_ghosts()
        for _ in range(args.warmup):
            one_step()
        barrier()
        sampler = ClockSampler(lrank)
        if rank == 0:
            sampler.start()
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            one_step()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        ok = eng.realizable()

        # per-kernel timing of the dominant (stage) kernel, live, with events on the launching stream
        stage_ms = []
        for _ in range(2):
            if hx is not None:
                eng.step_begin_dev(hx.global_dt().data_ptr())
            else:
                eng.local_dt(dt_dev.data_ptr()); eng.step_begin_dev(dt_dev.data_ptr())
            for s in range(nstages):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); eng.stage(s); b_.record(stream)
                refresh_ghosts()
                stage_ms.append((a, b_))
        torch.cuda.synchronize()
        stage_ms = [a.elapsed_time(b_) for a, b_ in stage_ms]

        # end-to-end through the host-buffer API: EVERY step uploads its input state from pinned host
        # memory, refreshes the ghosts, advances one time step and reads the result back into pinned
        # host memory.  The steps are independent jobs, so the streaming entry points overlap the H2D
        # copy of step n+1 and the D2H copy of step n-1 with the compute of step n (three streams).
        e2e_steps = max(1, args.e2e_steps)
        out_host = {gid: torch.empty((n, n, 4), dtype=torch.float64, pin_memory=True).numpy() for gid in mine}
        in_host = {gid: host_states[gid].numpy() for gid in mine}

        def stage_inputs():
            for gid in mine:
                eng.upload_async(gid, in_host[gid])

        barrier()
        t0 = time.perf_counter()
        stage_inputs()
        for i in range(e2e_steps):
            eng.commit_uploads()
            if i + 1 < e2e_steps:
                stage_inputs()          # waits (on the copy stream) until the conversion above has consumed the staging area
            refresh_ghosts()
            one_step()
            for gid in mine:
                eng.download_async(gid, out_host[gid])
        eng.transfers_sync()
        barrier()
        e2e_s = time.perf_counter() - t0

    # max over ranks
    cells_local = len(mine) * n * n
    if dist is not None:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device=f"cuda:{lrank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        cc = torch.tensor([cells_local, launches], dtype=torch.float64, device=f"cuda:{lrank}")
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        cells_total, launches_total = int(cc[0]), int(cc[1])
    else:
        cells_total, launches_total = cells_local, launches
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = cells_total * nstages * args.steps / (ms * 1e-3)
    e2e_value = cells_total * nstages * e2e_steps / e2e_s
    peak, peak_src = load_peaks()
    bytes_per_cell_stage = BYTES_PER_CELL_STEP.get(integ, 128.0 * nstages) / nstages
    kern_ms = float(np.mean(stage_ms))
    achieved = cells_local * bytes_per_cell_stage / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("block") == n and tj.get("blocks_per_gpu") == nb:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    fp64_peak = 18.27e12  # measured DFMA/DADD/DMUL issue rate, profiles/r01_fp64_microbench.txt
    out = {
        "metric": "cell-stage updates/sec", "value": value, "unit": "cell-stage updates/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"synthetic weak-scaling explosion: {nb}x{n_gpus} blocks of {n}x{n} cells, {args.flux} + Venkatakrishnan + GreenGauss, {args.recon} reconstruction, {integ}, CFL 0.7, reflection BCs" + ("" if args.ic == "explosion" else f", IC {args.ic}"),
            "blocks_per_gpu": nb, "block": n, "cells_total": cells_total, "stages_per_step": nstages,
            "parallelism": f"block-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2_policy": f"inputs larger than L2 ({cells_local * 32 / 1e6:.0f} MB per state array per GPU vs 126 MB L2)",
            "realizable_after_run": bool(ok),
        },
        "e2e": {
            "value": e2e_value, "unit": "cell-stage updates/s",
            "h2d_bytes_per_step": cells_total * 32, "d2h_bytes_per_step": cells_total * 32,
            "steps": e2e_steps,
            "how": "per step: H2D of all block states from pinned host memory (pyh_upload_state_async + pyh_commit_uploads), ghost refresh, one time step, D2H of all block states into pinned host memory (pyh_download_state_async); steps are independent jobs pipelined over copy-in / compute / copy-out streams; wall clock from the first H2D to the last D2H, max over ranks",
        },
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "kernel": "k_stage_march", "kernel_ms_avg": kern_ms,
            "algorithmic_bytes_per_cell_stage": bytes_per_cell_stage, "cells_per_launch": cells_local,
            "fp64_pipe": {
                "note": "bit-faithful fp64 (no FMA contraction): the FP64 pipe binds before HBM (DESIGN.md)",
                "peak_issue_per_s": fp64_peak,
                "cell_stage_per_s_kernel": cells_local / (kern_ms * 1e-3),
                # FP64 instructions per cell-stage of the headline scheme, counted by ncu (profiles/r01n_summary.md);
                # ncu (profiles/r01s_*): 1152 with the exact power-of-two scalings folded; the literal operation list of SURVEY.md section 8A needs 1187
                "fp64_instr_per_cell_stage": 1152 if (args.flux == "Roe" and args.recon == "conservative") else None,
                "frac": (1152 * cells_local / (kern_ms * 1e-3) / fp64_peak) if (args.flux == "Roe" and args.recon == "conservative") else None,
            },
        },
    }
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)   # the CPU arm may use every core again
    if not args.no_cpu_baseline and n_gpus == 1:
        out["cpu_baseline"] = cpu_baseline(args, seconds=args.cpu_seconds)
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_sample_problem(args, nblk_side):
    from oracle import muscl_oracle as mo

    nb = args.blocks_per_gpu
    blocks = ws_mesh(1, nb)
    prob = mo.Problem(blocks, nblk_side, nblk_side, flux=args.flux, limiter="Venkatakrishnan", recon="conservative",
                      integrator=args.integrator, CFL=0.7)
    for b in prob.blocks.values():
        b.U = ws_ic(b.geom.xc, b.geom.yc, BLOCK_LEN * nb, BLOCK_LEN)
    prob.apply_bc()
    return prob


def _cpu_worker(rank, world, port, argd, seconds, steps, ret):
    import torch
    import torch.distributed as dist

    from oracle.sharded import OracleShardEngine
    from pyhype_b200.distributed import HaloExchanger, advance, distribute_blocks

    torch.set_num_threads(1)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nb, side = argd["blocks_per_gpu"], argd["cpu_block"]
    blocks = ws_mesh(1, nb)
    owner = distribute_blocks(len(blocks), world)
    eng = OracleShardEngine(blocks, side, side, owner, rank, lambda x, y: ws_ic(x, y, BLOCK_LEN * nb, BLOCK_LEN),
                            flux=argd["flux"], limiter="Venkatakrishnan", recon="conservative",
                            integrator=argd["integrator"], CFL=0.7)
    hx = HaloExchanger(eng, owner, rank, backend_device=torch.device("cpu"))

    def one_step():
        dt = hx.global_dt()
        advance(eng, hx, eng.num_stages, dt_dev_ptr=dt.data_ptr())

    hx.exchange()
    eng.apply_bc()
    one_step()  # warm-up
    dist.barrier()
    if steps is None:  # calibrate the step count for ~`seconds` of work
        t0 = time.perf_counter()
        one_step()
        dist.barrier()
        el = torch.tensor([time.perf_counter() - t0])
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        steps = max(2, min(400, int(seconds / max(float(el[0]), 1e-3))))
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dist.barrier()
    el = torch.tensor([time.perf_counter() - t0])
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret["seconds"] = float(el[0])
        ret["steps"] = steps
        ret["stages"] = eng.num_stages
    dist.destroy_process_group()


def cpu_baseline(args, seconds=15.0, steps=None):
    """Times the numpy port of the reference (oracle/, kind "port") on the host cores, on a bounded
    sample of the workload (same mesh / IC / scheme at reduced block size).  Like the reference under
    `mpiexec -n P` it runs one single-threaded process per block group (P = min(cores, blocks)),
    exchanging ghost strips every RK stage and reducing dt every step -- over gloo instead of MPI
    (oracle/sharded.py)."""
    import socket

    import torch.multiprocessing as mp

    side, nb = args.cpu_block, args.blocks_per_gpu
    nproc = max(1, min(os.cpu_count() or 1, nb, args.cpu_procs if args.cpu_procs > 0 else 1 << 30))
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    argd = dict(blocks_per_gpu=nb, cpu_block=side, flux=args.flux, integrator=args.integrator)
    mp.spawn(_cpu_worker, args=(nproc, port, argd, seconds, steps, ret), nprocs=nproc, join=True)
    cells = nb * side * side
    n, el, nstages = ret["steps"], ret["seconds"], ret["stages"]
    return {
        "value": cells * nstages * n / el, "unit": "cell-stage updates/s", "cores": nproc, "kind": "port",
        "sample": f"{nb} blocks of {side}x{side} (same mesh/IC/scheme as the workload at reduced block size), {n} {args.integrator} steps, {el:.1f} s, {nproc} single-threaded processes (block-sharded like mpiexec -n {nproc}, ghost exchange + dt reduction over gloo)",
        "host_cores_available": os.cpu_count(),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args, steps=args.steps)  # each worker does its own untimed warm-up step
    nstages = 4 if args.integrator == "RK4" else len(__import__("oracle.muscl_oracle", fromlist=["TABLEAUX"]).TABLEAUX[args.integrator])
    side = args.cpu_block
    cells = args.blocks_per_gpu * side * side
    out = {
        "impl": "reference", "metric": "cell-stage updates/sec", "value": cb["value"], "unit": "cell-stage updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cells * nstages / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"synthetic weak-scaling explosion: {args.blocks_per_gpu}x1 blocks, {args.flux} + Venkatakrishnan + GreenGauss, conservative reconstruction, {args.integrator}, CFL 0.7, reflection BCs; CPU arm runs a bounded sample at {side}x{side} cells per block",
            "blocks_per_gpu": args.blocks_per_gpu, "block": side,
        },
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "cell-stage updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, default=2048, help="cells per block side")
    ap.add_argument("--blocks-per-gpu", type=int, default=8)
    ap.add_argument("--flux", default="Roe")
    ap.add_argument("--integrator", default="RK4")
    ap.add_argument("--recon", default="conservative", choices=["conservative", "primitive"])
    ap.add_argument("--ic", default="explosion", choices=["explosion", "smooth"], help="diagnostics; the headline workload is 'explosion'")
    ap.add_argument("--e2e-steps", type=int, default=24, help="pipelined end-to-end steps (fill + drain of the three-stream pipeline cost ~2 steps)")
    ap.add_argument("--cpu-block", type=int, default=192, help="block side of the CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes of the CPU arm (0 = min(cores, blocks))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
