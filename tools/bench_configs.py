#!/usr/bin/env python
"""Throughput of the reference's own named configurations (BASELINE.json configs[0], configs[1]) on one GPU,
through the device-resident loop (pyh_run): cell-stage updates/s, wall clock around the run (one host sync at
each end), after a warm-up run.  These are small problems (180 k and 1 M cells): they measure launch / latency
behaviour, not the roofline -- the headline number is bench.py's.

    python tools/bench_configs.py [--steps 200]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402

CONFIGS = {
    # examples/explosion_multi: 2x4 blocks of 150x150, Roe + Venkatakrishnan + GreenGauss, RK4, CFL 0.7 (as shipped)
    "explosion_multi": dict(mesh=cases.em_mesh, n=150, ic=cases.explosion_ic, flux="Roe", recon="conservative", integrator="RK4", CFL=0.7),
    # examples/dmr at the README size: 4 blocks of 500x500, HLLL + Venkatakrishnan, primitive reconstruction, RK2 (midpoint), CFL 0.4
    "dmr": dict(mesh=cases.dmr_mesh, n=500, ic=cases.dmr_ic, flux="HLLL", recon="primitive", integrator="RK2", CFL=0.4),
    "dmr_ssprk2": dict(mesh=cases.dmr_mesh, n=500, ic=cases.dmr_ic, flux="HLLL", recon="primitive", integrator="SSPRK2", CFL=0.4),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from pyhype_b200.time_marching import get_tableau

    for name, c in CONFIGS.items():
        if args.only and name != args.only:
            continue
        blocks = c["mesh"]()
        tab = get_tableau(c["integrator"])
        eng = cases.build_engine(blocks, c["n"], c["n"], c["ic"], flux=c["flux"], recon=c["recon"], integrator=tab, CFL=c["CFL"])
        cells = len(blocks) * c["n"] ** 2
        t, n, bad, _ = eng.run(0.0, 1e9, max_steps=20)          # warm-up (also advances past the very first transients)
        eng.sync()
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        t, n, bad, _ = eng.run(t, 1e9, max_steps=args.steps, poll_every=args.steps)
        eng.sync()
        dt = time.perf_counter() - t0
        print(json.dumps({
            "config": name, "cells": cells, "stages": len(tab), "steps": int(n), "unrealizable": bool(bad),
            "ms_per_step": dt / n * 1e3, "cell_stage_updates_per_s": cells * len(tab) * n / dt,
            "launches_per_step": (eng.launch_count() - l0) / n,
        }), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
