#!/usr/bin/env python
"""Golden outputs for the randomly generated CLI cases of tests/test_differential_cpu.py, printed by the REFERENCE
ITSELF (build container only).  The inputs are a pure function of the seed (numpy default_rng), so the GPU box
regenerates them and compares the real kernels' output with what the reference printed here:

    make -C oracle ref && python tests/golden/make_golden_fuzz.py        -> tests/golden/fuzz/<seed>.stdout, cases.json
"""
import json
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_golden as mg  # noqa: E402
import test_differential_cpu as diff  # noqa: E402

N_SEEDS = 90


def main():
    out_dir = os.path.join(HERE, "fuzz")
    os.makedirs(out_dir, exist_ok=True)
    meta = {}
    for seed in range(N_SEEDS):
        with tempfile.TemporaryDirectory() as root:
            rng = np.random.default_rng(9000 + seed)
            paths = diff.draw_inputs(rng, root)
            mode, argv, compat = diff.draw_argv(rng, paths)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out, _, code = mg.run_cli(argv)
            except Exception as exc:
                print("seed %d: reference raised %s -- no golden" % (seed, type(exc).__name__))
                continue
            inv = {v: "{%s}" % k for k, v in paths.items()}
            meta[str(seed)] = {"mode": mode, "exit": code, "argv": [inv.get(a, a) for a in argv] + compat}
            with open(os.path.join(out_dir, "%d.stdout" % seed), "w") as fh:
                fh.write(out)
    with open(os.path.join(out_dir, "cases.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("%d cases written" % len(meta))


if __name__ == "__main__":
    main()
