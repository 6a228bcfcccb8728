#!/usr/bin/env python
"""CPU replays of BASELINE.json's named configurations at their named sizes against the fingerprints of the unmodified
reference (tests/golden/named/, oracle/make_named_fingerprints.py) -- longer than the test suite affords:

    python tools/named_replay.py oracle em          # the numpy restatement, every checkpoint up to t_final (~35 min)
    python tools/named_replay.py twin em 200        # the stage kernel's source on the thread-block emulator, 200 steps (~20 min)
    python tools/named_replay.py twin dmr 20
    python tools/named_replay.py twin em 50 uniform_shortcut      # an opt-in build of tests/test_kernel_twin.py BUILDS

Test infrastructure (it runs oracle/ and tests/host_twin/); results of the last runs: profiles/r01u_named_config_parity.md."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import test_named_configs as N  # noqa: E402


def main():
    mode, name = sys.argv[1], sys.argv[2]
    fp = N.Named(name)
    t0 = time.time()
    if mode == "oracle":
        last = int(sys.argv[3]) if len(sys.argv) > 3 else fp.meta["checkpoints"][-1]
        prob = cases.build_oracle(fp.blocks, fp.nx, fp.ny, fp.ic, **fp.scheme())
        t, done = 0.0, 0
        for n in [c for c in fp.meta["checkpoints"] if c <= last]:
            t, dts = prob.run(t, fp.meta["t_final_nd"], max_steps=n - done)
            assert np.array_equal(np.asarray(dts), fp.dts[done:n]), ("dt sequence differs", name, n)
            done = n
            fp.check(n, {g: prob.blocks[g].U for g in fp.gids})
            print(f"oracle == reference: {name}, step {n}, t = {t!r} ({time.time() - t0:.0f} s)", flush=True)
        if done == fp.meta["checkpoints"][-1]:
            assert t == fp.meta["t_end"]
    elif mode == "twin":
        import test_kernel_twin as T

        n = int(sys.argv[3])
        build = sys.argv[4] if len(sys.argv) > 4 else "default"     # a key of tests/test_kernel_twin.py BUILDS (opt-in variants)
        assert n in fp.meta["checkpoints"], fp.meta["checkpoints"]
        idx, Uout, dts, t, nsteps, bad = T.run_loop(T.build(build), fp, 0.0, fp.meta["t_final_nd"], n, nt=128, tys=64)
        assert nsteps == n and not bad
        assert np.array_equal(dts, fp.dts[:n]), ("dt sequence differs", name, n)
        fp.check(n, {g: Uout[idx[g]] for g in fp.gids})
        print(f"kernel twin ({build} build) == reference: {name}, step {n} ({time.time() - t0:.0f} s)", flush=True)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
