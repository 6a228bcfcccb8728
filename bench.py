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


# ---------------------------------------------------------------------------------------------
# workload definition (shared by both arms)
# ---------------------------------------------------------------------------------------------
from pyhype_b200.examples import WS_BLOCK_LEN as BLOCK_LEN, ws_ic, ws_ic_smooth, ws_mesh  # noqa: E402  (the inputs tests/ pins to the reference)


BYTES_PER_CELL_STEP = {"RK4": 512.0, "RK2": 160.0, "ExplicitEuler1": 64.0}
# FP64 instructions (DFMA + DMUL + DADD + DSETP) per cell-stage of the headline scheme, counted by ncu on the shipped build
# (profiles/r02d_stage_march_ncu_summary.txt: 44.5 % of 2475 warp instructions per cell; r01s had 1152 before the Harten certificate
# and the min/max network; the literal operation list of SURVEY.md section 8A needs 1187)
FP64_INSTR_PER_CELL_STAGE = 1101


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


def value_digest(U):
    """sha256 of a state BY VALUE (-0.0 folded onto +0.0), the form tests/golden/named/*.npz store"""
    import hashlib

    return hashlib.sha256((np.ascontiguousarray(U) + 0.0).tobytes()).hexdigest()


def build_engine(blocks, mine, nx, ny, scheme, device, ic=None, pin=False, timing=None):
    """Engine with the local blocks `mine` of the block dictionary `blocks`; returns (engine, {gid: host state}).
    scheme: dict(flux, limiter, recon, integrator, CFL).  Host geometry is built in parallel threads (numpy-bound)."""
    from concurrent.futures import ThreadPoolExecutor

    from pyhype_b200.engine import Engine, SIDES
    from pyhype_b200.mesh.quad_mesh import QuadMesh
    from pyhype_b200.time_marching import get_tableau

    tab = get_tableau(scheme["integrator"])
    eng = Engine(nx, ny, scheme["flux"], scheme.get("limiter", "Venkatakrishnan"), scheme["recon"], tab, GAMMA, scheme["CFL"], device=device)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(mine)))) as pool:
        meshes = dict(zip(mine, pool.map(lambda g: QuadMesh(nx, ny, NE=blocks[g]["NE"], NW=blocks[g]["NW"], SE=blocks[g]["SE"], SW=blocks[g]["SW"]), mine)))
    t1 = time.perf_counter()
    states = {}
    for gid in mine:
        b = blocks[gid]
        m = meshes.pop(gid)
        eng.add_block(gid, m, {s: b["Neighbor" + s] for s in SIDES}, {s: b["BCType" + s] for s in SIDES}, local_gids=set(mine))
        if ic is not None:
            U = np.ascontiguousarray(ic(m.x[:, :, 0], m.y[:, :, 0]))
            if pin:
                buf = eng.pinned_state_buffer()
                buf[...] = U
                U = buf
            states[gid] = U
        del m
    eng.finalize()
    if timing is not None:
        timing["host_geometry_s"] = t1 - t0
        timing["add_blocks_and_ic_s"] = time.perf_counter() - t1
    return eng, states, len(tab)


class StreamTimer:
    """CUDA events on the engine's own stream (torch.cuda.Event only sees torch's current stream)."""

    def __init__(self, torch, eng, device):
        self.torch = torch
        self.stream = torch.cuda.ExternalStream(eng.stream(), device=torch.device("cuda", device))

    def event(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        return e


def selfcheck_sharded(torch, dist, rank, wsize, lrank, args):
    """The sharded run must produce the single-GPU bits: the weak-scaling grid at a tiny block size, a few time steps on
    `wsize` ranks through pyh_run (NCCL strip exchange + dt all-reduce inside), every block's state digest gathered on
    rank 0 and compared with the same run done on rank 0's GPU alone (all blocks local).  That single-GPU path is what
    the -m gpu suite pins to the reference."""
    from pyhype_b200.distributed import distribute_blocks, share_unique_id
    from pyhype_b200.engine import Engine

    nb, n, steps = args.blocks_per_gpu, 48, 6
    blocks = ws_mesh(wsize, nb)
    owner = distribute_blocks(len(blocks), wsize)
    mine = sorted(g for g, r in owner.items() if r == rank)
    width, height = BLOCK_LEN * nb, BLOCK_LEN * wsize
    scheme = dict(flux=args.flux, recon=args.recon, integrator=args.integrator, CFL=0.7)
    ic = lambda x, y: ws_ic(x, y, width, height) + 0.05 * (ws_ic_smooth(x, y, width, height) - ws_ic_smooth(0 * x, 0 * y, width, height))  # noqa: E731
    eng, states, _ = build_engine(blocks, mine, n, n, scheme, lrank, ic=ic)
    eng.comm_init(rank, wsize, share_unique_id(Engine.comm_unique_id, rank, wsize), owner)
    for g in mine:
        eng.upload(g, states[g])
    eng.apply_bc()
    t, nsteps, bad, dts = eng.run(0.0, 1e9, max_steps=steps, poll_every=4, record_dts=steps)
    mine_dig = {g: value_digest(eng.download(g)) for g in mine}
    eng.close()
    gathered = [None] * wsize
    dist.gather_object((mine_dig, list(dts), nsteps, bool(bad)), gathered if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    ref, rstates, _ = build_engine(blocks, sorted(blocks), n, n, scheme, lrank, ic=ic)
    for g in sorted(blocks):
        ref.upload(g, rstates[g])
    ref.apply_bc()
    t1, n1, bad1, dts1 = ref.run(0.0, 1e9, max_steps=steps, poll_every=4, record_dts=steps)
    ref_dig = {g: value_digest(ref.download(g)) for g in sorted(blocks)}
    ref.close()
    ok, seen = True, {}
    for dig, d, ns, b in gathered:
        seen.update(dig)
        ok = ok and d == list(dts1) and ns == n1 and not b
    ok = ok and seen == ref_dig and not bad1
    return {"sharded_equals_single_gpu": bool(ok), "ranks": wsize, "blocks": len(blocks), "block": n, "steps": steps,
            "what": "per-block sha256 of the state by value + the dt sequence, N ranks vs one GPU"}


NAMED = {
    # BASELINE.json configs[0]: examples/explosion_multi exactly as shipped (config.py:8-34; CFL 0.7)
    "explosion_multi": dict(fingerprint="em", title="examples/explosion_multi: 2x4 blocks of 150x150, Roe + Venkatakrishnan + GreenGauss, RK4, CFL 0.7 (as shipped), reflection BCs, t_final=0.07"),
    # BASELINE.json configs[1]: the DMR scheme (examples/dmr/config.py:8-33) at the README's 500x500 blocks; the shipped mesh has 4 blocks
    "dmr": dict(fingerprint="dmr", title="examples/dmr: Mach 10 double Mach reflection, 4 blocks (as shipped) of 500x500, HLLL + Venkatakrishnan, primitive reconstruction, RK2 (midpoint, the factory's SSP-RK2 stand-in), CFL 0.4"),
}


def run_named(torch, name, device, steps, peak):
    """One of the reference's own named configurations at its named size on one GPU, through the device-resident loop
    (pyh_run), replayed against the fingerprints of the unmodified reference (tests/golden/named/*.npz: data files written
    by oracle/make_named_fingerprints.py; no oracle code runs here)."""
    from pyhype_b200 import examples as ex

    fpz = np.load(os.path.join(ROOT, "tests", "golden", "named", NAMED[name]["fingerprint"] + ".npz"))
    meta = json.loads(str(fpz["meta"]))
    blocks = getattr(ex, meta["mesh"])(*meta.get("mesh_args", []))
    ic = getattr(ex, meta["ic"])
    nx, ny = meta["nx"], meta["ny"]
    scheme = dict(flux=meta["flux"], limiter=meta["limiter"], recon=meta["recon"], integrator=meta["integrator"], CFL=meta["CFL"])
    gids = sorted(blocks)
    eng, states, nstages = build_engine(blocks, gids, nx, ny, scheme, device, ic=ic, pin=True)
    cells = len(gids) * nx * ny
    shape = eng.march_shape()
    timer = StreamTimer(torch, eng, device)

    def load_ic():
        for g in gids:
            eng.upload(g, states[g])
        eng.apply_bc()

    # 1) parity replay (also the warm-up: captures the CUDA graph of one step)
    load_ic()
    done, t, parity_ok = 0, 0.0, True
    for n in meta["checkpoints"]:
        t, k, bad, dts = eng.run(t, meta["t_final_nd"], max_steps=n - done, poll_every=64, record_dts=n - done)
        parity_ok = parity_ok and (k == n - done) and not bad and np.array_equal(np.asarray(dts), fpz["dts"][done:n])
        done = n
        for g in gids:
            parity_ok = parity_ok and value_digest(eng.download(g)) == meta["digests"][f"{n}_{g}"]
    # 2) timed: the same run again from the initial condition, `steps` steps (0 = as far as the fingerprint goes)
    K = min(steps, done) if steps > 0 else done
    load_ic()
    eng.sync()
    l0 = eng.launch_count()
    e0 = timer.event()
    t, k, bad, _ = eng.run(0.0, meta["t_final_nd"], max_steps=K, poll_every=max(K, 1))
    e1 = timer.event()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    path, tuning = eng.stage_path(), eng.stage_path_tuning()   # chosen by measurement at the first run() (pyh_stage_path)
    # 3) end to end: host state in (pinned) -> H2D -> ghost refresh -> K steps -> D2H into pinned host buffers
    outs = {g: eng.pinned_state_buffer() for g in gids}
    t0 = time.perf_counter()
    for g in gids:
        eng.upload_async(g, states[g])
    eng.commit_uploads()
    eng.apply_bc()
    eng.run(0.0, meta["t_final_nd"], max_steps=K, poll_every=max(K, 1))
    for g in gids:
        eng.download_async(g, outs[g])
    eng.transfers_sync()
    e2e_s = time.perf_counter() - t0
    eng.close()
    value = cells * nstages * k / (ms * 1e-3)
    bpcs = BYTES_PER_CELL_STEP.get(meta["integrator"], 128.0 * nstages) / nstages
    return {
        "workload": NAMED[name]["title"], "cells_total": cells, "stages_per_step": nstages, "steps": int(k),
        "value": value, "unit": "cell-stage updates/s", "ms_per_step": ms / max(k, 1), "gpu_launches": int(launches),
        "stage_path": path, "stage_path_tuning_ms": {"fused": tuning[0], "split": tuning[1], "what": "ms per stage launch measured at the first pyh_run; the faster path runs"},
        "stage_kernel_shape": {"lanes": shape[0], "rows_per_strip": shape[1]} if path == "fused" else {"kernels": "k_split_recon (32 x 8 cells per thread block) -> k_split_flux (one thread per face) -> k_split_update (one thread per cell)"},
        "parity": {"bit_identical_to_reference": bool(parity_ok), "checkpoints": meta["checkpoints"],
                   "what": "every dt and the sha256-by-value of every block state at each checkpoint vs the unmodified reference's fingerprint (" + meta["generator"] + ")"},
        "e2e": {"value": cells * nstages * k / e2e_s, "unit": "cell-stage updates/s", "h2d_bytes_per_step": cells * 32 / max(k, 1),
                "d2h_bytes_per_step": cells * 32 / max(k, 1), "how": f"one job = upload the initial state from pinned host memory, ghost refresh, {int(k)} steps in the device-resident loop, download the state; bytes averaged over the steps"},
        "roofline": {"bound": "hbm", "achieved": value * bpcs / 1e9, "peak": peak, "unit": "GB/s", "frac": value * bpcs / 1e9 / peak, "traffic": None,
                     "algorithmic_bytes_per_cell_stage": bpcs,
                     "note": "whole-step rate (all kernels of the captured step graph), not a single launch: a stage launch covers only %d cells here (%.2f waves of 592 resident thread blocks)" % (cells, cells / (124.0 * 64 * 592))},
    }


def run_ours(args):
    import torch

    from pyhype_b200.distributed import distribute_blocks, share_unique_id, world
    from pyhype_b200.engine import Engine

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
    peak, peak_src = load_peaks()
    dist = None
    if wsize > 1:
        import torch.distributed as dist

        # control plane only (barriers, the max over ranks, the NCCL id): the data path is the library's own communicator
        dist.init_process_group("gloo", rank=rank, world_size=wsize)

    if args.config != "ws":   # one of the reference's own small configurations as the main line (single GPU)
        if wsize > 1:
            raise SystemExit("--config explosion_multi|dmr are single-GPU lines (8 and 4 small blocks)")
        sampler = ClockSampler(lrank)
        sampler.start()
        r = run_named(torch, args.config, lrank, args.steps if args.steps_given else 0, peak)
        clocks = sampler.stop()
        out = {"metric": "cell-stage updates/sec", "value": r["value"], "unit": r["unit"], "n_gpus": 1, "steps": r["steps"], "warmup": args.warmup,
               "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": r["workload"], "cells_total": r["cells_total"], "stages_per_step": r["stages_per_step"], "parity": r["parity"],
                          "l2_policy": "state (%.1f MB) fits the 126 MB L2: this is the reference's own size" % (r["cells_total"] * 32 / 1e6)},
               "stage_path": r["stage_path"], "stage_kernel_shape": r["stage_kernel_shape"],
               "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks, "roofline": dict(r["roofline"], peak_source=peak_src)}
        print(json.dumps(out), flush=True)
        return

    selfcheck = None
    if wsize > 1 and not args.no_selfcheck:
        selfcheck = selfcheck_sharded(torch, dist, rank, wsize, lrank, args)

    nb, n = args.blocks_per_gpu, args.block
    integ = args.integrator
    blocks = ws_mesh(n_gpus, nb)
    owner = distribute_blocks(len(blocks), wsize)
    mine = sorted(g for g, r in owner.items() if r == rank)
    width, height = BLOCK_LEN * nb, BLOCK_LEN * n_gpus
    scheme = dict(flux=args.flux, recon=args.recon, integrator=integ, CFL=0.7)
    setup = {}
    def ws_box_states():
        """the two conservative states of ws_ic and its box, for the device-side fill (pyh_fill_box)"""
        hi = ws_ic(np.array([[0.5 * width]]), np.array([[0.5 * height]]), width, height)[0, 0]
        lo = ws_ic(np.array([[0.0]]), np.array([[0.0]]), width, height)[0, 0]
        return (0.3 * width, 0.7 * width, 0.3 * height, 0.7 * height), hi, lo

    def set_initial_state(e):
        """explosion box: evaluated on the device's centroids (no (ny, nx, 4) host array, no upload); smooth field: host numpy + upload"""
        if args.ic == "explosion":
            box, hi, lo = ws_box_states()
            for gid in mine:
                e.fill_box(gid, *box, hi, lo)
        else:
            for gid in mine:
                m = e_meshes[gid]
                e.upload(gid, np.ascontiguousarray(ws_ic_smooth(m[0], m[1], width, height)))
        e.apply_bc()

    e_meshes = {}
    if args.ic == "smooth":   # centroids for the host-evaluated field
        from pyhype_b200.mesh.quad_mesh import QuadMesh

        for gid in mine:
            b = blocks[gid]
            m = QuadMesh(n, n, NE=b["NE"], NW=b["NW"], SE=b["SE"], SW=b["SW"])
            e_meshes[gid] = (m.x[:, :, 0].copy(), m.y[:, :, 0].copy())
            del m
    eng, _, nstages = build_engine(blocks, mine, n, n, scheme, lrank, ic=None, timing=setup)
    if wsize > 1:
        eng.comm_init(rank, wsize, share_unique_id(Engine.comm_unique_id, rank, wsize), owner)
    timer = StreamTimer(torch, eng, lrank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    set_initial_state(eng)
    eng.sync()
    setup["initial_state_s"] = time.perf_counter() - t0
    # the end-to-end leg starts from HOST buffers: page-locked copies of the initial state
    host_states = {gid: eng.pinned_state_buffer() for gid in mine}
    for gid in mine:
        eng.download(gid, out=host_states[gid])

    # ---- device-timed region: exactly K steps of the loop Euler2D._solve runs (pyh_run: CFL reduction [+ all-reduce], stages,
    # [strip exchange,] ghost refresh; one captured CUDA graph per step), state resident in HBM
    t, _, bad, _ = eng.run(0.0, 1e9, max_steps=args.warmup, poll_every=args.warmup)
    barrier()
    sampler = ClockSampler(lrank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    e0 = timer.event()
    t, k, bad, _ = eng.run(t, 1e9, max_steps=args.steps, poll_every=args.steps)
    e1 = timer.event()
    barrier()
    ms = e0.elapsed_time(e1)
    assert k == args.steps, (k, args.steps)
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ok = eng.realizable()

    # ---- the dominant kernel alone: events on the launching stream around every stage launch of two more steps
    stage_ev = []
    for _ in range(2):
        eng.local_dt()          # dt stays in the context's device scratch
        eng.step_begin_dev()
        for s in range(nstages):
            a = timer.event(); eng.stage(s); b_ = timer.event()
            eng.apply_bc()
            stage_ev.append((a, b_))
    torch.cuda.synchronize()
    stage_ms = [a.elapsed_time(b_) for a, b_ in stage_ev]

    # ---- sustained window: a second, long timed region (>= 5 s at the default size) with the clock sampler on
    sustained = None
    if args.sustain_steps > 0:
        barrier()
        s2 = ClockSampler(lrank)
        if rank == 0:
            s2.start()
        f0 = timer.event()
        t, k2, bad2, _ = eng.run(t, 1e9, max_steps=args.sustain_steps, poll_every=args.sustain_steps)
        f1 = timer.event()
        barrier()
        sus_ms = f0.elapsed_time(f1)
        sus_clocks = s2.stop() if rank == 0 else None
        sustained = (sus_ms, int(k2), sus_clocks)

    # ---- end to end through the host-buffer API: EVERY step uploads its input state from pinned host memory, refreshes the
    # ghosts, advances one time step (pyh_run, 1 step) and reads the result back into pinned host memory.  The steps are
    # independent jobs, so the streaming entry points overlap the H2D copy of step n+1 and the D2H copy of step n-1 with the
    # compute of step n (three streams).
    e2e_steps = max(1, args.e2e_steps)
    out_host = {gid: eng.pinned_state_buffer() for gid in mine}

    def stage_inputs():
        for gid in mine:
            eng.upload_async(gid, host_states[gid])

    barrier()
    t0 = time.perf_counter()
    stage_inputs()
    for i in range(e2e_steps):
        eng.commit_uploads()
        if i + 1 < e2e_steps:
            stage_inputs()          # waits (on the copy stream) until the conversion above has consumed the staging area
        eng.apply_bc()
        eng.run(0.0, 1e9, max_steps=1, poll_every=1)
        for gid in mine:
            eng.download_async(gid, out_host[gid])
    eng.transfers_sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    stage_path_info = (eng.stage_path(), dict(zip(("fused", "split"), eng.stage_path_tuning())))   # which kernels ran the stages (pyh_stage_path)
    eng.close()

    # max over ranks
    cells_local = len(mine) * n * n
    sus_ms = sustained[0] if sustained else 0.0
    kern_per_rank = None
    if dist is not None:
        tt = torch.tensor([ms, e2e_s, sus_ms], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s, sus_ms = float(tt[0]), float(tt[1]), float(tt[2])
        kk = [torch.zeros(1, dtype=torch.float64) for _ in range(wsize)]
        dist.all_gather(kk, torch.tensor([float(np.mean(stage_ms))], dtype=torch.float64))
        kern_per_rank = [float(x[0]) for x in kk]   # chip-to-chip spread: every step waits for the slowest rank
        cc = torch.tensor([cells_local, launches], dtype=torch.float64)
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
    bytes_per_cell_stage = BYTES_PER_CELL_STEP.get(integ, 128.0 * nstages) / nstages
    kern_ms = float(np.mean(stage_ms)) if kern_per_rank is None else max(kern_per_rank)
    achieved = cells_local * bytes_per_cell_stage / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("block") == n and tj.get("blocks_per_gpu") == nb:
                traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
    fp64_peak = 18.27e12  # measured DFMA/DADD/DMUL issue rate, profiles/r01_fp64_microbench.txt
    headline = args.flux == "Roe" and args.recon == "conservative"
    fp64_per_cell = FP64_INSTR_PER_CELL_STAGE if headline else None
    out = {
        "metric": "cell-stage updates/sec", "value": value, "unit": "cell-stage updates/s",
        "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"synthetic weak-scaling explosion: {nb}x{n_gpus} blocks of {n}x{n} cells, {args.flux} + Venkatakrishnan + GreenGauss, {args.recon} reconstruction, {integ}, CFL 0.7, reflection BCs" + ("" if args.ic == "explosion" else f", IC {args.ic}"),
            "blocks_per_gpu": nb, "block": n, "cells_total": cells_total, "stages_per_step": nstages,
            "parallelism": f"block-sharded x{n_gpus}: NCCL strip exchange per stage + dt all-reduce per step inside the captured step graph (pyh_comm_init)" if n_gpus > 1 else "single GPU",
            "l2_policy": f"inputs larger than L2 ({cells_local * 32 / 1e6:.0f} MB per state array per GPU vs 126 MB L2)",
            "timed_call": "pyh_run (the device-resident loop Euler2D._solve drives): one CUDA graph per time step",
            "realizable_after_run": bool(ok),
            "setup_s": {k_: round(v, 3) for k_, v in setup.items()},
            "stage_path": stage_path_info[0],   # 'fused' at the headline size (33.5 M cells per GPU are outside the split stage's range)
            "stage_path_tuning_ms": stage_path_info[1],
        },
        "e2e": {
            "value": e2e_value, "unit": "cell-stage updates/s",
            "h2d_bytes_per_step": cells_total * 32, "d2h_bytes_per_step": cells_total * 32,
            "steps": e2e_steps,
            "how": "per step: H2D of all block states from pinned host memory (pyh_upload_state_async + pyh_commit_uploads), ghost refresh, one time step (pyh_run), D2H of all block states into pinned host memory (pyh_download_state_async); steps are independent jobs pipelined over copy-in / compute / copy-out streams; wall clock from the first H2D to the last D2H, max over ranks",
        },
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": peak_src, "kernel": "k_stage_march", "kernel_ms_avg": kern_ms, "kernel_ms_avg_per_rank": kern_per_rank,
            "algorithmic_bytes_per_cell_stage": bytes_per_cell_stage, "cells_per_launch": cells_local,
            "binding_resource": "fp64_pipe",
            "fp64_pipe": {
                "note": "what actually binds: bit-faithful fp64 (no FMA contraction, IEEE division / sqrt) needs ~1.1 k FP64 instructions per cell-stage, so the FP64 pipe saturates long before HBM (DESIGN.md section 4); `frac` above is the contract's HBM fraction, this is the pipe's",
                "peak_issue_per_s": fp64_peak,
                "cell_stage_per_s_kernel": cells_local / (kern_ms * 1e-3),
                "fp64_instr_per_cell_stage": fp64_per_cell,
                "frac": (fp64_per_cell * cells_local / (kern_ms * 1e-3) / fp64_peak) if fp64_per_cell else None,
            },
        },
    }
    if sustained:
        out["sustained"] = {"steps": sustained[1], "seconds": sus_ms * 1e-3, "value": cells_total * nstages * sustained[1] / (sus_ms * 1e-3),
                            "unit": "cell-stage updates/s", "clocks": sustained[2]}
    if selfcheck is not None:
        out["selfcheck"] = selfcheck
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)   # the CPU arm may use every core again
    if n_gpus == 1 and not args.no_named:
        out["named_configs"] = {}
        for name in NAMED:
            try:
                out["named_configs"][name] = run_named(torch, name, lrank, 0, peak)
            except Exception as e:   # a missing fingerprint file must not take the headline line down
                out["named_configs"][name] = {"error": repr(e)}
    if n_gpus == 1 and not args.no_named and headline and integ == "RK4":
        # the DMR scheme (HLLL + primitive reconstruction + RK2) on the SAME weak-scaling blocks: how the other flux family runs at
        # a size that fills the GPU (its named-size line above is a 0.2-wave problem)
        try:
            sch2 = dict(flux="HLLL", recon="primitive", integrator="RK2", CFL=0.4)
            eng2, st2, ns2 = build_engine(blocks, mine, n, n, sch2, lrank, ic=None)
            set_initial_state(eng2)
            t2, _, _, _ = eng2.run(0.0, 1e9, max_steps=args.warmup, poll_every=args.warmup)
            tm2 = StreamTimer(torch, eng2, lrank)
            g0 = tm2.event()
            t2, k2_, bad2_, _ = eng2.run(t2, 1e9, max_steps=args.steps, poll_every=args.steps)
            g1 = tm2.event()
            torch.cuda.synchronize()
            ms2 = g0.elapsed_time(g1)
            eng2.close()
            v2 = cells_local * ns2 * k2_ / (ms2 * 1e-3)
            out["other_schemes"] = {"hlll_primitive_rk2": {
                "workload": f"same {nb} blocks of {n}x{n}, HLLL + Venkatakrishnan, primitive reconstruction, RK2 (the DMR scheme), CFL 0.4",
                "value": v2, "unit": "cell-stage updates/s", "ms_per_step": ms2 / max(k2_, 1), "steps": int(k2_), "realizable": not bad2_,
                "roofline": {"bound": "hbm", "achieved": v2 * 80.0 / 1e9, "peak": peak, "unit": "GB/s", "frac": v2 * 80.0 / 1e9 / peak,
                             "algorithmic_bytes_per_cell_stage": 80.0, "note": "whole-step rate; HLLL needs OpenBLAS dnrm2's x87 arithmetic emulated in double-double (DESIGN.md section 4)"}}}
        except Exception as e:
            out["other_schemes"] = {"error": repr(e)}
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
    # under torch.distributed.run the parent's TORCHELASTIC_* / rendezvous variables would make this private gloo group
    # connect as a CLIENT to a store nobody serves (TORCHELASTIC_USE_AGENT_STORE): drop them and rendezvous explicitly
    for k in [k for k in os.environ if k.startswith("TORCHELASTIC_") or k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK", "MASTER_ADDR", "MASTER_PORT")]:
        del os.environ[k]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nb, side, rows = argd["blocks_per_gpu"], argd["cpu_block"], argd["rows"]
    blocks = ws_mesh(rows, nb)
    owner = distribute_blocks(len(blocks), world)
    eng = OracleShardEngine(blocks, side, side, owner, rank, lambda x, y: ws_ic(x, y, BLOCK_LEN * nb, BLOCK_LEN * rows),
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


def cpu_baseline(args, seconds=15.0, steps=None, rows=1):
    """Times the numpy port of the reference (oracle/, kind "port") on the host cores, on a bounded
    sample of the workload (same mesh / IC / scheme at reduced block size).  Like the reference under
    `mpiexec -n P` it runs one single-threaded process per block group (P = min(cores, blocks)),
    exchanging ghost strips every RK stage and reducing dt every step -- over gloo instead of MPI
    (oracle/sharded.py)."""
    import socket

    import torch.multiprocessing as mp

    side, nb = args.cpu_block, args.blocks_per_gpu
    nproc = max(1, min(os.cpu_count() or 1, nb * rows, args.cpu_procs if args.cpu_procs > 0 else 1 << 30))
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    argd = dict(blocks_per_gpu=nb, cpu_block=side, flux=args.flux, integrator=args.integrator, rows=rows)
    mp.spawn(_cpu_worker, args=(nproc, port, argd, seconds, steps, ret), nprocs=nproc, join=True)
    cells = nb * rows * side * side
    n, el, nstages = ret["steps"], ret["seconds"], ret["stages"]
    return {
        "value": cells * nstages * n / el, "unit": "cell-stage updates/s", "cores": nproc, "kind": "port",
        "sample": f"{nb}x{rows} blocks of {side}x{side} (same mesh/IC/scheme as the workload at reduced block size), {n} {args.integrator} steps, {el:.1f} s, {nproc} single-threaded processes (block-sharded like mpiexec -n {nproc}, ghost exchange + dt reduction over gloo)",
        "host_cores_available": os.cpu_count(),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = max(1, int(os.environ.get("WORLD_SIZE", args.gpus)))   # the N-GPU workload is an 8 x N grid of blocks
    cb = cpu_baseline(args, steps=args.steps, rows=rows)  # each worker does its own untimed warm-up step
    nstages = 4 if args.integrator == "RK4" else len(__import__("oracle.muscl_oracle", fromlist=["TABLEAUX"]).TABLEAUX[args.integrator])
    side = args.cpu_block
    cells = args.blocks_per_gpu * rows * side * side
    out = {
        "impl": "reference", "metric": "cell-stage updates/sec", "value": cb["value"], "unit": "cell-stage updates/s",
        "n_gpus": rows, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cells * nstages / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"synthetic weak-scaling explosion: {args.blocks_per_gpu}x{rows} blocks, {args.flux} + Venkatakrishnan + GreenGauss, conservative reconstruction, {args.integrator}, CFL 0.7, reflection BCs; CPU arm runs a bounded sample at {side}x{side} cells per block",
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
    ap.add_argument("--config", default="ws", choices=["ws", "explosion_multi", "dmr"],
                    help="ws: the weak-scaling workload the metric is quoted on (default); explosion_multi / dmr: the reference's own configurations at their named sizes as the main line (1 GPU)")
    ap.add_argument("--no-named", action="store_true", help="skip the named-configuration sub-lines of the default N=1 run")
    ap.add_argument("--no-selfcheck", action="store_true", help="skip the sharded == single-GPU bit check at the start of an N>1 run")
    ap.add_argument("--sustain-steps", type=int, default=300, help="second, long timed window (0 = off)")
    args = ap.parse_args()
    args.steps_given = any(a == "--steps" or a.startswith("--steps=") for a in sys.argv[1:])
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
