"""Dev script (GPU box): dense (random-model) kernel, ROS5L vs ROS6L at several tolerances: throughput, steps, error vs a tight run."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, ns, B in (("randmod", 6, 16384), ("randmod", 4, 65536)):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(3)
    y0s = np.asarray(initial_condition(ns, model))
    for name, p, y0 in (("U steady-y0", rng.uniform(0.05, 3.0, (B, P)), y0s),
                        ("logU steady-y0", np.exp(rng.uniform(np.log(0.01), np.log(20.0), (B // 4, P))), y0s),
                        ("U random-y0", rng.uniform(0.05, 3.0, (B // 2, P)), None)):
        y0 = rng.uniform(0.1, 1.1, (p.shape[0], n)) if y0 is None else y0
        nref = min(4096, p.shape[0])
        ref = eng.solve_local_batch(model, p[:nref], y0 if y0.ndim == 1 else y0[:nref], ns, T, want=("sol",), rtol=1e-10, atol=1e-14, method="ros5l")
        for method, rtol in (("ros5l", 2e-6), ("ros6l", 2e-5), ("ros6l", 1e-5), ("ros6l", 5e-6)):
            a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), method=method, rtol=rtol, atol=2e-9)
            a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), method=method, rtol=rtol, atol=2e-9)
            ms = eng.last_launch_info()[1]
            ok = (a["status"][:nref] == 0) & (ref["status"] == 0)
            ratio = (np.abs(a["sol"][:nref][ok] - ref["sol"][ok]) / (1e-6 * np.abs(ref["sol"][ok]) + 1e-9)).max(axis=(1, 2))
            print(f"{model}-{ns} {name:15s} {method} rtol {rtol:g}: {p.shape[0] / ms * 1e3:10.4g} solves/s steps {a['nsteps'].mean():6.1f} rej {a['nrej'].mean():5.2f} fail {int((a['status'] != 0).sum())} "
                  f"err/bound max {ratio.max():.3f} p99.9 {np.percentile(ratio, 99.9):.3f}", flush=True)
