#!/usr/bin/env python
"""Runs the reference's whole stage sequence (3 data associations + S0..S5, L/test/lvi_initialize_surfel_orb.cpp:539-708) through the
CPU ORACLE on the synthetic sequence and writes the per-stage results (iterations, costs, extrinsics, wall time) as JSON.

Two uses (test infrastructure, never the product path):
  * the C2 / C5 result files under tests/golden/ are what bench.py's `extrinsic_err_vs_oracle` and the C5 parity test compare the
    CUDA path against (the oracle needs minutes per configuration, so it is run once and committed);
  * `--time` on the GPU box gives the CPU stage-sequence wall time that the north star's ">= 50x calibration-solve" is quoted against.

    python tools/oracle_calibration.py --duration 60 --out tests/golden/oracle_calibration_c2.json
    python tools/oracle_calibration.py --duration 60 --degenerate --out tests/golden/oracle_calibration_c5.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--duration", type=float, default=60.0)
    ap.add_argument("--degenerate", action="store_true", help="C5: planar, low-excitation trajectory (SURVEY §8d)")
    ap.add_argument("--assoc-mode", type=int, default=1, help="0 = reference-faithful O(P*W*H) sweep, 1 = voxel lookup (identical results)")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from lvi_exc_b200 import pipeline, synth
    from tests import oracle_binding as ob
    from tests.oracle_backend import OracleBackend
    threads = args.threads or os.cpu_count()
    ob.set_num_threads(threads)
    cfg = synth.default_config(duration=args.duration, degenerate=int(args.degenerate))
    t0 = time.perf_counter()
    seq = synth.make_sequence(cfg)
    t_gen = time.perf_counter() - t0
    backend = OracleBackend(assoc_mode=args.assoc_mode)
    timed = {}

    class Timed:
        """wall time per backend call, by name"""
        def __init__(self, inner):
            self._i = inner
        def __getattr__(self, name):
            f = getattr(self._i, name)
            if not callable(f):
                return f
            def g(*a, **k):
                t = time.perf_counter()
                r = f(*a, **k)
                timed[name] = timed.get(name, 0.0) + time.perf_counter() - t
                return r
            return g

    t0 = time.perf_counter()
    res = pipeline.run_calibration(seq, Timed(backend), verbose=True)
    wall = time.perf_counter() - t0
    c = res["calib"]
    out = {
        "config": {"seconds": args.duration, "degenerate": bool(args.degenerate), "assoc_mode": args.assoc_mode},
        "threads": ob.lib().orc_num_threads(), "generator_s": t_gen, "wall_s": wall, "backend_s": timed,
        "stages": [{k: (st[k] if not isinstance(st[k], dict) else {kk: float(vv) for kk, vv in st[k].items()})
                    for k in ("name", "iterations", "initial_cost", "final_cost", "termination", "errors", "time_ms", "n_res")} for st in res["stages"]],
        "assoc_counts": res["assoc_counts"], "n_lm_plane": res.get("n_lm_plane"),
        "calib": {k: np.asarray(getattr(c, k)).tolist() for k in ("q_LtoI", "p_LinI", "q_CtoI", "p_CinI", "gravity_rp", "acc_bias", "gyr_bias")},
        "extrinsic_err_vs_gt": {k: float(v) for k, v in pipeline.extrinsic_errors(c, seq.gt).items()},
    }
    s = json.dumps(out, indent=1)
    if args.out:
        Path(args.out).write_text(s + "\n")
    print(s)


if __name__ == "__main__":
    main()
