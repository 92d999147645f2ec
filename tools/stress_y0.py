"""Dev script (GPU box): local kernels vs a tight run of the same kernels for RANDOM initial conditions (large
transients, components decaying towards zero) — the tail that decides the default tolerances."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phoskintime_b200 as pk
eng = pk.get_engine(0)
T = np.array([0.0, 0.5, 0.75, 1.0, 2.0, 4.0, 8.0, 16.0, 30.0, 60.0, 120.0, 240.0, 480.0, 960.0])
B = 200000
for model, ns in (("succmod", 5), ("distmod", 3)):
    n, P, L = pk.local_dims(model, ns, 14)
    rng = np.random.default_rng(11)
    for name, p in (("U(0.05,3)", rng.uniform(0.05, 3.0, (B, P))), ("logU(0.01,20)", np.exp(rng.uniform(np.log(0.01), np.log(20.0), (B, P))))):
        y0 = rng.uniform(0.1, 1.1, (B, n))
        ref = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), rtol=1e-10, atol=1e-14, method="ros5l")
        for method, rtol, atol in (("ros5l", 2e-6, 2e-9), ("ros6l", 2e-5, 2e-9), ("ros6l", 1e-5, 2e-9), ("ros6l", 5e-6, 2e-9), ("ros6l", 5e-6, 1e-9), ("ros6l", 2e-6, 2e-9)):
            a = eng.solve_local_batch(model, p, y0, ns, T, want=("sol",), rtol=rtol, atol=atol, method=method)
            ok = (a["status"] == 0) & (ref["status"] == 0)
            ratio = (np.abs(a["sol"][ok] - ref["sol"][ok]) / (1e-6 * np.abs(ref["sol"][ok]) + 1e-9)).max(axis=(1, 2))
            print(f"{model}-{ns} {name:14s} {method} rtol {rtol:g} atol {atol:g}: fail {int((~ok).sum())} steps {a['nsteps'].mean():.1f} rej {a['nrej'].mean():.2f} | "
                  f"vs tight: max {ratio.max():.3g} p99.99 {np.percentile(ratio, 99.99):.3g} p99.9 {np.percentile(ratio, 99.9):.3g} median {np.median(ratio):.3g}", flush=True)
