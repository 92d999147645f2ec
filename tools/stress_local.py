"""Dev script (GPU box): local kernels at the library defaults vs a tight-tolerance run of the same kernels over
harsh parameter draws (log-uniform over the reference's fit bounds [1e-2, 20], config.toml:189-195)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
from phoskintime_b200.steady import initial_condition
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
for model, ns, B in (("distmod", 3, 100000), ("succmod", 5, 100000), ("distmod", 8, 50000), ("randmod", 4, 20000), ("randmod", 6, 4000)):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(3)
    y0 = np.asarray(initial_condition(ns, model))
    for name, draw in (("U(0.05,3)", lambda: rng.uniform(0.05, 3.0, (B, P))), ("logU(0.01,20)", lambda: np.exp(rng.uniform(np.log(0.01), np.log(20.0), (B, P))))):
        p = draw()
        ref = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), rtol=1e-10, atol=1e-14)
        for rtol, atol in ((None, None), (2e-6, 2e-10), (1e-6, 1e-10), (5e-7, 5e-10)):
            a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), rtol=rtol, atol=atol)
            ok = (a["status"] == 0) & (ref["status"] == 0)
            ratio = (np.abs(a["sol"][ok] - ref["sol"][ok]) / (1e-6 * np.abs(ref["sol"][ok]) + 1e-9)).max(axis=(1, 2))
            print(f"{model}-{ns} {name:14s} rtol {rtol} atol {atol}: fail {int((a['status'] != 0).sum())} steps {a['nsteps'].mean():.0f} | vs tight: max {ratio.max():.3g} p99.9 {np.percentile(ratio, 99.9):.3g} median {np.median(ratio):.3g}", flush=True)
